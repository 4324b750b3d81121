// 3x3 64->64 convolutions of the Real NVP s/t ResNet (nn_extra_nvp.py:108-141, conv2d_wn :378-404) as tcgen05
// implicit GEMMs on sm_100a, forward / dgrad (one kernel, different epilogues) and wgrad.
//
//   forward / dgrad:  D[128 positions][64 out-ch] = sum over 9 taps, 4 K-steps of  A_tap[128][16] * W_tap[16][64]
//     A_tap is a ROW-SHIFTED view of one shared-memory slab (see slab.cuh): no im2col, one 1-D bulk copy per tile.
//     Warp roles: warp 0 = bulk-copy producer, warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = epilogue
//     (tcgen05.ld -> bias / ELU / residual / ELU' -> bf16 -> swizzled store).  TMEM accumulators are double buffered so
//     the epilogue of tile i overlaps the MMAs of tile i+1.  Persistent: one CTA per SM, weights loaded once.
//   wgrad:  dW[tap][ci][co] = sum over positions  X[pos + shift(tap)][ci] * G[pos][co]
//     both operands MN-major (contraction over positions = rows); two taps share one M=128 instruction (the two
//     64-channel atoms are LBO bytes apart = the row shift between the taps); the bias gradient rides along as the
//     spare atom of the ninth tap (a constant "ones" operand).  fp32 accumulation in TMEM over all tiles of the CTA.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "slab.cuh"

namespace bruno {

constexpr int W_TAP_BYTES = 64 * 128;        // one tap: 64 rows (N) x 64 bf16 (K)
constexpr int W_BYTES = 9 * W_TAP_BYTES;     // 73728
constexpr int NSTAGE = 3;
constexpr int CONV_THREADS = 192;

enum ConvMode { MODE_ELU = 0, MODE_RES_ELU = 1, MODE_MUL = 2, MODE_ADD_MUL = 3, MODE_LINEAR = 4 };

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }
// derivative of ELU expressed through its OUTPUT o: 1 if o > 0 else o + 1
__device__ __forceinline__ float elu_grad_from_out(float o) { return o > 0.f ? 1.f : o + 1.f; }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// Optional pipeline trace (debug): CTA 0 records clock64() at role events into a caller-provided device buffer
// [tile][16]: 0 slab load issued, 1 slab landed (MMA thread), 2 MMAs issued, 3 accumulator ready (epilogue warp 2),
// 4 epilogue math done, 5 store issued, 6 store smem-read done.
// The buffer pointer is a kernel parameter (a uniform compare, not a global load on the issue path).
static long long* g_conv_trace_host = nullptr;
__device__ __forceinline__ void trace_at(long long* buf, int it, int slot) {
  if (buf != nullptr && blockIdx.x == 0 && it < 64) buf[it * 16 + slot] = clock64();
}

constexpr int FWD_EPI_WARPS = 16;            // 4 per TMEM lane quarter: each handles 32 rows x 16 channels
constexpr int EPI_CH = 16;                   // channels per epilogue pass
constexpr int FWD_THREADS = 32 * (2 + FWD_EPI_WARPS + 1);   // producer, MMA issuer, epilogue warps, aux/store warp
constexpr int TILE_BYTES = TILE_M * SLAB_ROW_BYTES;          // 16 KB: one aux / gradient tile of the wgrad kernel
constexpr int OUT_ROWS = 126;                // output rows per tile of the dx-fused forward kernel (see below)
constexpr int STG_ROWS = 136;                // staging rows: 126 + up to 7 rows of 8-row phase alignment, rounded
constexpr int STG_BYTES = STG_ROWS * SLAB_ROW_BYTES;
constexpr int OUT_BYTES = OUT_ROWS * SLAB_ROW_BYTES;

struct ConvSmem {
  uint64_t full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2], wbar;
  uint64_t auxfull[2], outready[2];
  uint32_t tmem_base;
  alignas(16) float bias[64];
  alignas(16) float xch[2][2][4][4][EPI_CH];   // [tile parity][dx block 0 / 2][lane quarter][channel group][channels]
};

__host__ __device__ constexpr bool mode_has_aux1(int m) { return m == MODE_RES_ELU || m == MODE_MUL || m == MODE_ADD_MUL; }
__host__ __device__ constexpr bool mode_has_aux2(int m) { return m == MODE_ADD_MUL; }

__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ELU with one MUFU: exp(x) - 1 = 2^(x log2 e) - 1 (abs. error ~1e-7, far below the bf16 output quantum)
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : ex2_ftz(x * 1.4426950408889634f) - 1.f; }

// ---- dx-fused implicit GEMM --------------------------------------------------------------------------------------
// A 64 -> 64 convolution issued as nine M=128, N=64 MMAs per K step reads 6 KB of shared memory per 128x64x16 MMA and is
// bound by the 128 B/clk shared-memory port at 49 cycles per MMA = 65 % of the tensor peak (tools/umma_bench.cu).
// Instead, for each vertical tap dy ONE MMA with N = 192 computes the three horizontal taps at once from the SAME
// activation rows:   E_dx[q] = sum_dy in[q + dy P] W[dy, dx]      (D[q][dx*64 + co], 12 MMAs of N = 192 per tile,
// 10 KB of shared memory per 128x192x16 MMA = 96 cycles = tensor-bound), and the epilogue applies the horizontal shift
//                    out[p] = E_-1[p - 1] + E_0[p] + E_+1[p + 1]
// with one warp shuffle per value (neighbouring positions are neighbouring TMEM lanes = neighbouring threads); the two
// boundary lanes of every warp are exchanged through shared memory, and consecutive tiles overlap by two rows
// (128 computed rows, 126 stored) so no value is ever needed from another tile.
// Warp roles: 0 slab producer, 1 MMA issuer, 2..9 epilogue, 10 aux prefetch / output drain; all global traffic is bulk
// async copies (TMA engine); TMEM accumulators (2 x 192 columns) are double buffered.
template <int MODE>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv3x3_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
               const uint8_t* __restrict__ aux1, const uint8_t* __restrict__ aux2, uint8_t* __restrict__ out,
               SlabGeom g, int ntiles, int bias_img_stride, long long* trace_buf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int slab_rows = TILE_M + 2 * g.HALO + 8;
  const uint32_t slab_bytes = static_cast<uint32_t>(slab_rows) * SLAB_ROW_BYTES;
  uint8_t* wsm = smem;
  uint8_t* slabs = smem + W_BYTES;
  uint8_t* stage1 = slabs + NSTAGE * slab_bytes;                       // [2][STG_BYTES] aux1 in / output out (in place)
  uint8_t* stage2 = stage1 + 2 * STG_BYTES;                            // [2][STG_BYTES] aux2 (MODE_ADD_MUL only)
  ConvSmem* cs = reinterpret_cast<ConvSmem*>(stage2 + (mode_has_aux2(MODE) ? 2 * STG_BYTES : 0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool HAS_BIAS = MODE == MODE_ELU || MODE == MODE_RES_ELU || MODE == MODE_LINEAR;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&cs->full[s], 1);
      mbar_init(&cs->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&cs->tfull[a], 1);
      mbar_init(&cs->tempty[a], FWD_EPI_WARPS / 2);
      mbar_init(&cs->auxfull[a], 1);
      mbar_init(&cs->outready[a], FWD_EPI_WARPS * 16);
    }
    mbar_init(&cs->wbar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) cs->bias[threadIdx.x - 64] = HAS_BIAS ? bias[threadIdx.x - 64] : 0.f;
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = cs->tmem_base;
  // Programmatic dependent launch: let the next kernel of the stream be scheduled now (its prologue and weight load
  // overlap our tail); our own reads / writes of activation slabs wait for the previous kernel (griddepcontrol.wait).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // geometry of tile t: computed rows = data rows [126 t - 1, 126 t + 127); stored rows = data rows [126 t, 126 t + 126)
  //   first computed buffer row  Rc = HALO + 126 t - 1;  slab load starts at LS = floor8(Rc - P)
  //   first stored buffer row    Rf = HALO + 126 t;      staging row 0 <-> buffer row RB = floor8(Rf)
  if (warp == 0) {
    // ===== producer: weights once, then the activation slab of every tile =====
    if (lane == 0) {
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, wpack, W_BYTES, &cs->wbar);      // weights do not depend on the previous kernel
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int LS = (g.HALO + OUT_ROWS * tile - 1 - g.P) & ~7;
        mbar_wait(&cs->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&cs->full[s], slab_bytes);
        bulk_g2s(slabs + s * slab_bytes, in + static_cast<size_t>(LS) * SLAB_ROW_BYTES, slab_bytes, &cs->full[s]);
        trace_at(trace_buf, it, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: 3 vertical taps x 4 K steps, N = 192 =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 0, 0);
      mbar_wait(&cs->wbar, 0);
      const uint64_t bdesc0 = umma_desc_sw128(smem_u32(wsm), 16, 1024);
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        const int Rc = g.HALO + OUT_ROWS * tile - 1;
        const int LS = (Rc - g.P) & ~7;
        mbar_wait(&cs->tempty[a], aph ^ 1);
        mbar_wait(&cs->full[s], ph);
        tc_fence_after();
        trace_at(trace_buf, it, 1);
        const uint64_t adesc0 = umma_desc_sw128(smem_u32(slabs + s * slab_bytes), 16, 1024);
        const uint32_t d_tmem = tmem + a * 256;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const uint32_t a_off = static_cast<uint32_t>(Rc + (dy - 1) * g.P - LS) * 8u;   // rows * 128 B, in 16-byte units
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d_tmem, adesc0 + (a_off + 2u * k), bdesc0 + static_cast<uint32_t>((dy * 3 * W_TAP_BYTES + k * 32) >> 4),
                      idesc, (dy | k) != 0);
          }
        }
        umma_commit(&cs->empty[s]);
        umma_commit(&cs->tfull[a]);
        trace_at(trace_buf, it, 2);
      }
    }
  } else if (warp < 2 + FWD_EPI_WARPS) {
    // ===== epilogue: two groups of 8 warps alternate tiles (group = accumulator / staging buffer index), so that the
    // TMEM-load + shuffle phase of one tile overlaps the math + store phase of the previous one.  Inside a group,
    // warp -> TMEM lane quarter (w & 3) and channel half; each half is processed as two passes of EPI_CH channels.
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;              // computed row inside the tile (0..127); rows 1..126 are stored
    constexpr int NCH = EPI_CH / 8;             // 8-channel chunks per pass
    const float m_up = lane > 0 ? 1.f : 0.f, m_dn = lane < 31 ? 1.f : 0.f;   // warp-edge lanes get theirs via smem
    const bool stored = row >= 1 && row <= OUT_ROWS;
    const int a = grp;
    // row bookkeeping without integer division in the loop: q_img = data_row mod IMG advances by a constant per tile
    const int tile_stride = 2 * gridDim.x;
    const int step_mod = (OUT_ROWS * tile_stride) % g.IMG;
    const int step_div = (OUT_ROWS * tile_stride) / g.IMG;
    int tile = blockIdx.x + grp * gridDim.x;
    int data_row = OUT_ROWS * tile - 1 + row;
    int q_img = ((data_row % g.IMG) + g.IMG) % g.IMG;
    int n_img = (data_row - q_img) / g.IMG;     // image index of the row (for per-image biases)
    const float inv_p = 1.0f / static_cast<float>(g.P);
    int it = grp;
    for (; tile < ntiles; tile += tile_stride, it += 2) {
      const uint32_t aph = (it >> 1) & 1;
      const int yy = static_cast<int>((static_cast<float>(q_img) + 0.5f) * inv_p);   // exact for q_img < 2^20
      const int xx = q_img - yy * g.P;
      const bool valid = stored && data_row >= 0 && data_row < g.data_rows && yy >= 1 && xx >= 1;
      const int Rf = g.HALO + OUT_ROWS * tile;
      const int srow = (Rf & 7) + row - 1;      // staging row (only meaningful when stored)
      const int R7 = (g.HALO + data_row) & 7;   // swizzle phase of the buffer row
      const uint32_t st1 = smem_u32(stage1 + a * STG_BYTES + srow * SLAB_ROW_BYTES);
      const uint32_t st2 = smem_u32(stage2 + a * STG_BYTES + srow * SLAB_ROW_BYTES);
      mbar_wait(&cs->tfull[a], aph);
      tc_fence_after();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 3);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int cg = half * 2 + p;
        const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + a * 256 + cg * EPI_CH;
        uint32_t e0[NCH][8], e1[NCH][8], e2[NCH][8];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_ld8(tbase + 0 * 64 + c * 8, e0[c]);
          tmem_ld8(tbase + 1 * 64 + c * 8, e1[c]);
          tmem_ld8(tbase + 2 * 64 + c * 8, e2[c]);
        }
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cs->tempty[a]);
          if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 6);
        }
        const uint32_t xw0 = smem_u32(cs->xch[a][0][q][cg]);   // my dx = -1 block, needed by row + 1 -> written by lane 31
        const uint32_t xw2 = smem_u32(cs->xch[a][1][q][cg]);   // my dx = +1 block, needed by row - 1 -> written by lane 0
        float acc[EPI_CH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (lane == 31) {
            sts_u4(xw0 + c * 32, make_uint4(e0[c][0], e0[c][1], e0[c][2], e0[c][3]));
            sts_u4(xw0 + c * 32 + 16, make_uint4(e0[c][4], e0[c][5], e0[c][6], e0[c][7]));
          }
          if (lane == 0) {
            sts_u4(xw2 + c * 32, make_uint4(e2[c][0], e2[c][1], e2[c][2], e2[c][3]));
            sts_u4(xw2 + c * 32 + 16, make_uint4(e2[c][4], e2[c][5], e2[c][6], e2[c][7]));
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(e0[c][e]), 1);      // dx = -1 term of row - 1
            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(e2[c][e]), 1);    // dx = +1 term of row + 1
            acc[c * 8 + e] = fmaf(dn, m_dn, fmaf(up, m_up, __uint_as_float(e1[c][e])));
          }
        }
        // boundary exchange between the four lane quarters of this group (barrier id 1 + group, 256 threads)
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
        if ((lane == 0 && q > 0) || (lane == 31 && q < 3)) {
          const uint32_t xr = smem_u32(lane == 0 ? cs->xch[a][0][q - 1][cg] : cs->xch[a][1][q + 1][cg]);
#pragma unroll
          for (int e = 0; e < EPI_CH / 4; ++e) {
            const uint4 t = lds_u4(xr + e * 16);
            acc[4 * e] += __uint_as_float(t.x); acc[4 * e + 1] += __uint_as_float(t.y);
            acc[4 * e + 2] += __uint_as_float(t.z); acc[4 * e + 3] += __uint_as_float(t.w);
          }
        }
        if (p == 0) mbar_wait(&cs->auxfull[a], aph);          // aux tile(s) landed / staging buffer free
        if (stored) {
#pragma unroll
          for (int jj = 0; jj < NCH; ++jj) {
            const int j = cg * NCH + jj;
            const int pos = (j ^ R7) << 4;
            float o[8], bj[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = acc[jj * 8 + e];
            if (HAS_BIAS) {
              // shared bias from smem, or the per-image bias table of the label-conditioned layers [N][64]
              const float* bsrc = bias_img_stride ? bias + static_cast<size_t>(valid ? n_img : 0) * bias_img_stride + j * 8
                                                  : &cs->bias[j * 8];
              const float4 b0 = *reinterpret_cast<const float4*>(bsrc);
              const float4 b1 = *reinterpret_cast<const float4*>(bsrc + 4);
              bj[0] = b0.x; bj[1] = b0.y; bj[2] = b0.z; bj[3] = b0.w; bj[4] = b1.x; bj[5] = b1.y; bj[6] = b1.z; bj[7] = b1.w;
            }
            if (MODE == MODE_ELU || MODE == MODE_LINEAR) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float t = o[e] + bj[e];
                o[e] = MODE == MODE_ELU ? elu_fast(t) : t;
              }
            } else if (MODE == MODE_RES_ELU) {
              float sk[8];
              unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = elu_fast(o[e] + bj[e] + sk[e]);
            } else if (MODE == MODE_MUL) {
              float sk[8];
              unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = o[e] * elu_grad_from_out(sk[e]);
            } else {
              float sk[8], t[8];
              unpack8(lds_u4(st1 + pos), sk);
              unpack8(lds_u4(st2 + pos), t);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = (o[e] + t[e]) * elu_grad_from_out(sk[e]);
            }
            sts_u4(st1 + pos, valid ? pack8(o) : make_uint4(0, 0, 0, 0));
          }
        }
      }
      fence_proxy_async_smem();                 // my generic-proxy writes -> visible to the bulk store
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 4);
      mbar_arrive(&cs->outready[a]);
      data_row += OUT_ROWS * tile_stride;
      q_img += step_mod;
      n_img += step_div;
      if (q_img >= g.IMG) {
        q_img -= g.IMG;
        ++n_img;
      }
    }
  } else {
    // ===== aux / store warp: prefetch the aux tile(s) of tile `it`, then drain the output tile of tile `it - 1` =====
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      int prev_tile = -1;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        const bool live = tile < ntiles;
        if (live) {
          const int b = it & 1;
          const int Rf = g.HALO + OUT_ROWS * tile;
          bulk_wait_read<0>();                  // store of tile it-2 (same staging buffer) has finished reading it
          if (mode_has_aux1(MODE)) {
            const size_t off = static_cast<size_t>(Rf) * SLAB_ROW_BYTES;
            const uint32_t soff = static_cast<uint32_t>(Rf & 7) * SLAB_ROW_BYTES;
            mbar_arrive_expect_tx(&cs->auxfull[b], mode_has_aux2(MODE) ? 2 * OUT_BYTES : OUT_BYTES);
            bulk_g2s(stage1 + b * STG_BYTES + soff, aux1 + off, OUT_BYTES, &cs->auxfull[b]);
            if (mode_has_aux2(MODE)) bulk_g2s(stage2 + b * STG_BYTES + soff, aux2 + off, OUT_BYTES, &cs->auxfull[b]);
          } else {
            mbar_arrive(&cs->auxfull[b]);
          }
        }
        if (prev_tile >= 0) {
          const int pit = it - 1;
          const int pa = pit & 1;
          const int Rf = g.HALO + OUT_ROWS * prev_tile;
          mbar_wait(&cs->outready[pa], (pit >> 1) & 1);
          bulk_s2g(out + static_cast<size_t>(Rf) * SLAB_ROW_BYTES, stage1 + pa * STG_BYTES + (Rf & 7) * SLAB_ROW_BYTES, OUT_BYTES);
          bulk_commit();
          trace_at(trace_buf, pit, 5);
        }
        if (!live) break;
        prev_tile = tile;
      }
      bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ---- per-tap implicit GEMM (variant 1) ---------------------------------------------------------------------------
// Nine row-shifted views x 4 K steps = 36 MMAs of N = 64 accumulate into ONE 64-column accumulator, so the epilogue is
// a plain row-wise map (no horizontal shuffles, no boundary exchange, no overlapping tiles): a third of the TMEM loads
// and ~40 % of the instructions of the dx-fused epilogue.  The MMAs themselves are shared-memory-port bound (6 KB per
// 128x64x16 instruction = 49 cycles, 65 % of the tensor peak), so this variant wins only while the fused kernel is
// bound by its epilogue; both are kept and selected per process (bruno_debug_conv_variant) so they can be A/B timed.
// Four 64-column TMEM accumulators decouple the issuer from the two epilogue groups.
constexpr int TAP_ACC = 4;
// (dgrad + addend) epilogue: take the addend tile global -> registers instead of staging it through shared memory?  On paper
// it saves 32 KB of port traffic per tile and frees a slab stage; measured it DOUBLES the 28x28 launch (83.9 vs 45.7 us):
// a warp's 16-byte loads touch 32 different lines, the LSU path costs the tensor pipe far more than the staging did.
#ifdef BRUNO_TAP_AUX2_DIRECT
constexpr bool TAP_AUX2_DIRECT = true;
#else
constexpr bool TAP_AUX2_DIRECT = false;
#endif
constexpr int TAP_MAX_STAGES = 4;          // slab ring depth: as many as fit next to the staging buffers (host decides)

struct TapSmem {
  uint64_t full[TAP_MAX_STAGES], empty[TAP_MAX_STAGES], tfull[TAP_ACC], tempty[TAP_ACC], wbar;
  uint64_t auxfull[2], outready[2];
  uint64_t flagok[8];                        // chain kernel: neighbour flags of iteration it seen (ring of 8)
  volatile int prod_it;                      // chain kernel: iterations whose slab load the producer has issued
  uint32_t tmem_base;
  alignas(16) float bias[64];
};

template <int MODE, int NS>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv3x3_tap_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                   const uint8_t* __restrict__ aux1, const uint8_t* __restrict__ aux2, uint8_t* __restrict__ out,
                   SlabGeom g, int ntiles, int bias_img_stride, long long* trace_buf) {
  constexpr int nstage = NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int slab_rows = TILE_M + 2 * g.HALO + 8;
  const uint32_t slab_bytes = static_cast<uint32_t>(slab_rows) * SLAB_ROW_BYTES;
  uint8_t* wsm = smem;
  uint8_t* slabs = smem + W_BYTES;
  uint8_t* stage1 = slabs + nstage * slab_bytes;                       // [2][TILE_BYTES] aux1 in / output out (in place)
  uint8_t* stage2 = stage1 + 2 * TILE_BYTES;                           // [2][TILE_BYTES] aux2 (MODE_ADD_MUL only)
  TapSmem* cs = reinterpret_cast<TapSmem*>(stage2 + ((mode_has_aux2(MODE) && !TAP_AUX2_DIRECT) ? 2 * TILE_BYTES : 0));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  constexpr bool HAS_BIAS = MODE == MODE_ELU || MODE == MODE_RES_ELU || MODE == MODE_LINEAR;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TAP_MAX_STAGES; ++s) {
      mbar_init(&cs->full[s], 1);
      mbar_init(&cs->empty[s], 1);
    }
    for (int a = 0; a < TAP_ACC; ++a) {
      mbar_init(&cs->tfull[a], 1);
      mbar_init(&cs->tempty[a], FWD_EPI_WARPS / 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&cs->auxfull[a], 1);
      mbar_init(&cs->outready[a], FWD_EPI_WARPS * 16);
    }
    mbar_init(&cs->wbar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) cs->bias[threadIdx.x - 64] = HAS_BIAS ? bias[threadIdx.x - 64] : 0.f;
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, 64 * TAP_ACC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = cs->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile t = data rows [128 t, 128 t + 128) = buffer rows from Rf = HALO + 128 t (a multiple of 8); the slab load starts
  // at LS = floor8(Rf - P - 1) and covers every tap's view.
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, wpack, W_BYTES, &cs->wbar);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % nstage;
        const uint32_t ph = (it / nstage) & 1;
        const int LS = (g.HALO + TILE_M * tile - g.P - 1) & ~7;
        mbar_wait(&cs->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&cs->full[s], slab_bytes);
        bulk_g2s(slabs + s * slab_bytes, in + static_cast<size_t>(LS) * SLAB_ROW_BYTES, slab_bytes, &cs->full[s]);
        trace_at(trace_buf, it, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The whole warp runs the loop (uniform control flow: descriptors live in uniform registers and the
    // issue path is two or three instructions per MMA); one elected lane executes tcgen05.mma / commit.
    // The tile loop is unrolled over lcm(NS, TAP_ACC) tiles so that the slab stage and the accumulator of every unrolled
    // body are compile-time: no descriptor or TMEM address is recomputed between tiles (rewriting a uniform register that
    // queued MMAs still read stalls the issuer until the tensor pipe has drained -- a bubble per tile). =====
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    constexpr int U = (NS == 3) ? 12 : 4;
    mbar_wait(&cs->wbar, 0);
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(wsm), 16, 1024);
    uint64_t adesc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) adesc[s] = umma_desc_sw128(smem_u32(slabs + s * slab_bytes), 16, 1024);
    // Rf - LS is the same for every tile (HALO and 128 t are multiples of 8)
    const uint32_t a_base = static_cast<uint32_t>(g.HALO - ((g.HALO - g.P - 1) & ~7)) * 8u;
    const uint32_t p8 = static_cast<uint32_t>(g.P) * 8u;
    // The barrier waits of tile it+1 (each a ~100-cycle round trip even when the barrier completed long ago) are taken
    // by the elected lane BETWEEN the last taps of tile it, while the tensor pipe still holds queued MMAs.
    int tile = blockIdx.x;
    bool more = tile < ntiles;
    const bool leader = elect_one();
    if (more && leader) {
      mbar_wait(&cs->tempty[0], 1u);
      mbar_wait(&cs->full[0], 0u);
      tc_fence_after();
    }
    for (uint32_t r = 0; more; ++r) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (more) {
          const int s = u % NS, a = u % TAP_ACC;
          const int sn = (u + 1) % NS, an = (u + 1) % TAP_ACC;
          const int it = static_cast<int>(r) * U + u;
          tile += gridDim.x;
          more = tile < ntiles;
          if (leader) {
            trace_at(trace_buf, it, 1);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              const uint32_t a_off = a_base + static_cast<uint32_t>(dy) * p8 + static_cast<uint32_t>(dx * 8);
              if (more) {
                if (tap == 5) mbar_wait(&cs->tempty[an], ((r * (U / TAP_ACC) + (u + 1) / TAP_ACC) & 1u) ^ 1u);
                if (tap == 6) mbar_wait(&cs->full[sn], (r * (U / NS) + (u + 1) / NS) & 1u);
                if (tap == 7) tc_fence_after();
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem + a * 64, adesc[s] + (a_off + 2u * k),
                          bdesc0 + static_cast<uint32_t>((tap * W_TAP_BYTES + k * 32) >> 4), idesc, (tap | k) != 0);
              }
            }
            umma_commit(&cs->empty[s]);
            umma_commit(&cs->tfull[a]);
            trace_at(trace_buf, it, 2);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + FWD_EPI_WARPS) {
    // two groups of 8 warps alternate tiles (group = staging buffer); warp -> TMEM lane quarter and 32-channel half
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;
    const int R7 = lane & 7;                    // HALO and the tile origin are multiples of 8
    const int tile_stride = 2 * gridDim.x;
    const int step_mod = (TILE_M * tile_stride) % g.IMG;
    const int step_div = (TILE_M * tile_stride) / g.IMG;
    int tile = blockIdx.x + grp * gridDim.x;
    int data_row = TILE_M * tile + row;
    int q_img = data_row % g.IMG;
    int n_img = data_row / g.IMG;
    const float inv_p = 1.0f / static_cast<float>(g.P);
    const uint32_t st1 = smem_u32(stage1 + grp * TILE_BYTES + row * SLAB_ROW_BYTES);
    const uint32_t st2 = smem_u32(stage2 + grp * TILE_BYTES + row * SLAB_ROW_BYTES);
    if (mode_has_aux2(MODE) && TAP_AUX2_DIRECT) asm volatile("griddepcontrol.wait;" ::: "memory");   // reads aux2 from global
    int it = grp;
    for (; tile < ntiles; tile += tile_stride, it += 2) {
      const int a = it % TAP_ACC;
      // the addend tile (aux2) goes global -> registers: staging it costs 32 KB of shared-memory port traffic per tile and
      // a fourth pair of staging buffers (= one slab stage).  Issued before the accumulator wait: its latency hides there.
      uint4 t2[4];
      if (mode_has_aux2(MODE) && TAP_AUX2_DIRECT) {
        const uint8_t* arow = aux2 + static_cast<size_t>(g.HALO + TILE_M * tile + row) * SLAB_ROW_BYTES;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) t2[jj] = ldg_nc_u4(arow + (((half * 4 + jj) ^ R7) << 4));
      }
      const int yy = static_cast<int>((static_cast<float>(q_img) + 0.5f) * inv_p);
      const int xx = q_img - yy * g.P;
      const bool valid = data_row < g.data_rows && yy >= 1 && xx >= 1;
      mbar_wait(&cs->tfull[a], (it / TAP_ACC) & 1);
      tc_fence_after();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 3);
      const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + a * 64 + half * 32;
      uint32_t acc[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld8(tbase + c * 8, acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&cs->tempty[a]);
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 6);
      mbar_wait(&cs->auxfull[grp], (it >> 1) & 1);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = half * 4 + jj;
        const int pos = (j ^ R7) << 4;
        float o[8], bj[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(acc[jj][e]);
        if (HAS_BIAS) {
          const float* bsrc = bias_img_stride ? bias + static_cast<size_t>(valid ? n_img : 0) * bias_img_stride + j * 8
                                              : &cs->bias[j * 8];
          const float4 b0 = *reinterpret_cast<const float4*>(bsrc);
          const float4 b1 = *reinterpret_cast<const float4*>(bsrc + 4);
          bj[0] = b0.x; bj[1] = b0.y; bj[2] = b0.z; bj[3] = b0.w; bj[4] = b1.x; bj[5] = b1.y; bj[6] = b1.z; bj[7] = b1.w;
        }
        if (MODE == MODE_ELU || MODE == MODE_LINEAR) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = o[e] + bj[e];
            o[e] = MODE == MODE_ELU ? elu_fast(t) : t;
          }
        } else if (MODE == MODE_RES_ELU) {
          float sk[8];
          unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = elu_fast(o[e] + bj[e] + sk[e]);
        } else if (MODE == MODE_MUL) {
          float sk[8];
          unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = o[e] * elu_grad_from_out(sk[e]);
        } else {
          float sk[8], t[8];
          unpack8(lds_u4(st1 + pos), sk);
          unpack8(TAP_AUX2_DIRECT ? t2[jj] : lds_u4(st2 + pos), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = (o[e] + t[e]) * elu_grad_from_out(sk[e]);
        }
        sts_u4(st1 + pos, valid ? pack8(o) : make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 4);
      mbar_arrive(&cs->outready[grp]);
      data_row += TILE_M * tile_stride;
      q_img += step_mod;
      n_img += step_div;
      if (q_img >= g.IMG) {
        q_img -= g.IMG;
        ++n_img;
      }
    }
  } else {
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      int prev_tile = -1;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        const bool live = tile < ntiles;
        if (live) {
          const int b = it & 1;
          bulk_wait_read<0>();                  // the store of tile it-2 has finished reading this staging buffer
          if (mode_has_aux1(MODE)) {
            const size_t off = static_cast<size_t>(g.HALO + TILE_M * tile) * SLAB_ROW_BYTES;
            constexpr bool STAGE_AUX2 = mode_has_aux2(MODE) && !TAP_AUX2_DIRECT;
            mbar_arrive_expect_tx(&cs->auxfull[b], STAGE_AUX2 ? 2 * TILE_BYTES : TILE_BYTES);
            bulk_g2s(stage1 + b * TILE_BYTES, aux1 + off, TILE_BYTES, &cs->auxfull[b]);
            if (STAGE_AUX2) bulk_g2s(stage2 + b * TILE_BYTES, aux2 + off, TILE_BYTES, &cs->auxfull[b]);
          } else {
            mbar_arrive(&cs->auxfull[b]);
          }
        }
        if (prev_tile >= 0) {
          const int pit = it - 1;
          const int pa = pit & 1;
          mbar_wait(&cs->outready[pa], (pit >> 1) & 1);
          bulk_s2g(out + static_cast<size_t>(g.HALO + TILE_M * prev_tile) * SLAB_ROW_BYTES, stage1 + pa * TILE_BYTES, TILE_BYTES);
          bulk_commit();
          trace_at(trace_buf, pit, 5);
        }
        if (!live) break;
        prev_tile = tile;
      }
      bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64 * TAP_ACC);
  }
}

// ---- layer chain: the 2R convolutions of a coupling's s/t ResNet (or of its backward) in ONE persistent launch --------
// A 14x14 layer is 7.6 tiles per CTA = 8.4 us of pipeline, but a launch costs ~17 us: barrier / TMEM set-up, the wait for
// the previous kernel's flush, the first slab's latency, and the drain of the last tile (tools/dbg_trace.py).  The chain
// kernel keeps the CTA, its barriers, its TMEM and its pipeline state alive across layers.  Tile t of every layer belongs
// to CTA t mod grid; tile t of layer l needs the tiles t-1, t, t+1 of layer l-1 (halo < 128 rows), which are published
// through per-tile EPOCH flags in global memory (set after the tile's bulk store has completed, read with acquire by the
// slab producer).  All CTAs are co-resident (grid <= #SMs, one CTA per SM), and a CTA only ever waits for tiles of the
// PREVIOUS layer, so the waits cannot cycle.  The weights (one 72 KB buffer) are reloaded at each layer boundary as soon
// as the CTA's last MMA of the previous layer has completed.
// Same arithmetic as conv3x3_tap_kernel (bit-identical results, tests/test_conv_gpu.py).
// STATUS: correct, within 3.5 % of the per-layer launches but NOT faster yet -- bn2_omniglot_tp step 15.0 ms with the chain vs
// 14.5 ms with one launch per layer on the same box (619 -> 218 launches per step).  What the timeline (tools/
// dbg_trace_chain.py) showed and what was fixed: (i) ONE aux/store thread that waits for each store to land before the
// next is issued runs at 3000 cycles per tile (store latency ~1.4 us) against 2150 for the tensor pipe -> two store
// threads, one per staging buffer; (ii) a gpu-scope fence issued by the PRODUCER after the flag poll waits for its slab
// loads in flight (~2500 cycles per tile, loads serialised) -> a separate flag warp that owns no asynchronous copies polls,
// fences and releases the producer through an mbarrier ring; (iii) serial acquire polls -> three relaxed loads + one fence.
// Left: the tile period inside a layer is ~2230 cycles (per-tap kernel: 2100) and every layer boundary costs ~2200 cycles
// (the weights are swapped only after the layer's last MMA has retired).  Chaining ONLY the 14x14 layers
// (BRUNO_CHAIN_MAX_TILES=2000) gives 14.17 vs 14.25 ms per step and 4.83 vs 4.91 ms per evaluation: inside the noise.
// Off by default (bruno_debug_conv_chain).
constexpr int CHAIN_MAX_LAYERS = 16;
constexpr int CHAIN_THREADS = FWD_THREADS + 64;   // + second aux / store warp + flag warp
struct ChainLayer {
  const uint8_t* in;
  const uint8_t* w;
  const float* bias;
  const uint8_t* aux1;
  const uint8_t* aux2;
  uint8_t* out;
  int mode;
  int pad_;
};
struct ChainArgs {
  ChainLayer layer[CHAIN_MAX_LAYERS];
  int L;
  int bias_img_stride;
  unsigned epoch;           // value the flags of THIS launch take (monotonic over launches: no clearing)
  unsigned* flags;          // [L][ntiles]
  unsigned* error;          // set to 1 if a flag wait gave up (never in a correct run; avoids a hung GPU)
  long long* trace;         // debug timeline of CTA 0 (bruno_debug_conv_trace), or null
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool BWD, int NS>
__global__ void __launch_bounds__(CHAIN_THREADS, 1)
conv3x3_chain_kernel(const __grid_constant__ ChainArgs args, SlabGeom g, int ntiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int slab_rows = TILE_M + 2 * g.HALO + 8;
  const uint32_t slab_bytes = static_cast<uint32_t>(slab_rows) * SLAB_ROW_BYTES;
  uint8_t* wsm = smem;
  uint8_t* slabs = smem + W_BYTES;
  uint8_t* stage1 = slabs + NS * slab_bytes;                           // [2][TILE_BYTES] aux1 in / output out (in place)
  uint8_t* stage2 = stage1 + 2 * TILE_BYTES;                           // [2][TILE_BYTES] aux2 (backward only)
  TapSmem* cs = reinterpret_cast<TapSmem*>(stage2 + (BWD ? 2 * TILE_BYTES : 0));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int L = args.L;
  const int grid = gridDim.x;
  const int cnt = (ntiles - static_cast<int>(blockIdx.x) + grid - 1) / grid;     // tiles of this CTA per layer (>= 1)
  const int total = L * cnt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TAP_MAX_STAGES; ++s) {
      mbar_init(&cs->full[s], 1);
      mbar_init(&cs->empty[s], 1);
    }
    for (int a = 0; a < TAP_ACC; ++a) {
      mbar_init(&cs->tfull[a], 1);
      mbar_init(&cs->tempty[a], FWD_EPI_WARPS / 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&cs->auxfull[a], 1);
      mbar_init(&cs->outready[a], FWD_EPI_WARPS * 16);
    }
    mbar_init(&cs->wbar, 1);
    for (int f = 0; f < 8; ++f) mbar_init(&cs->flagok[f], 1);
    cs->prod_it = 0;
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, 64 * TAP_ACC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = cs->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== producer: per layer the weights, per tile the activation slab (after the neighbours' flags) =====
    if (lane == 0) {
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, args.layer[0].w, W_BYTES, &cs->wbar);           // weights do not depend on the previous kernel
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int l = 0, j = 0;
      for (int it = 0; it < total; ++it) {
        const int tile = blockIdx.x + j * grid;
        const int s = it % NS;
        if (j == 0 && l > 0) {
          // every MMA of layer l-1 in this CTA has completed -> the weight buffer may be overwritten
          mbar_wait(&cs->empty[(it - 1) % NS], ((it - 1) / NS) & 1);
          mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
          bulk_g2s(wsm, args.layer[l].w, W_BYTES, &cs->wbar);
        }
        mbar_wait(&cs->empty[s], ((it / NS) & 1) ^ 1);
        trace_at(args.trace, it, 7);
        // the neighbour tiles of the previous layer have been seen published by the flag warp (which holds no
        // outstanding asynchronous copies: a gpu-scope fence issued HERE waited for the slab loads in flight, ~2500 cycles
        // per tile, and serialised them)
        mbar_wait(&cs->flagok[it & 7], (it >> 3) & 1);            // (arrives for layer 0 too: uniform phases)
        const int LS = (g.HALO + TILE_M * tile - g.P - 1) & ~7;
        mbar_arrive_expect_tx(&cs->full[s], slab_bytes);
        bulk_g2s(slabs + s * slab_bytes, args.layer[l].in + static_cast<size_t>(LS) * SLAB_ROW_BYTES, slab_bytes, &cs->full[s]);
        trace_at(args.trace, it, 0);
        cs->prod_it = it + 1;
        if (++j == cnt) {
          j = 0;
          ++l;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (see conv3x3_tap_kernel); the iteration counter runs on across layers =====
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    constexpr int U = (NS == 3) ? 12 : 4;
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(wsm), 16, 1024);
    uint64_t adesc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) adesc[s] = umma_desc_sw128(smem_u32(slabs + s * slab_bytes), 16, 1024);
    const uint32_t a_base = static_cast<uint32_t>(g.HALO - ((g.HALO - g.P - 1) & ~7)) * 8u;
    const uint32_t p8 = static_cast<uint32_t>(g.P) * 8u;
    const bool leader = elect_one();
    int j = 0, l = 0;                                 // position of iteration `it` inside its layer
    if (leader) {
      mbar_wait(&cs->wbar, 0u);
      mbar_wait(&cs->tempty[0], 1u);
      mbar_wait(&cs->full[0], 0u);
      tc_fence_after();
    }
    bool more = total > 0;
    for (uint32_t r = 0; more; ++r) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (more) {
          const int s = u % NS, a = u % TAP_ACC;
          const int sn = (u + 1) % NS, an = (u + 1) % TAP_ACC;
          const int it = static_cast<int>(r) * U + u;
          more = it + 1 < total;
          const bool next_new_layer = (j + 1 == cnt);
          if (leader) {
            trace_at(args.trace, it, 1);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              const uint32_t a_off = a_base + static_cast<uint32_t>(dy) * p8 + static_cast<uint32_t>(dx * 8);
              if (more && !next_new_layer) {
                if (tap == 5) mbar_wait(&cs->tempty[an], ((r * (U / TAP_ACC) + (u + 1) / TAP_ACC) & 1u) ^ 1u);
                if (tap == 6) mbar_wait(&cs->full[sn], (r * (U / NS) + (u + 1) / NS) & 1u);
                if (tap == 7) tc_fence_after();
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem + a * 64, adesc[s] + (a_off + 2u * k),
                          bdesc0 + static_cast<uint32_t>((tap * W_TAP_BYTES + k * 32) >> 4), idesc, (tap | k) != 0);
              }
            }
            umma_commit(&cs->empty[s]);
            umma_commit(&cs->tfull[a]);
            trace_at(args.trace, it, 2);
            if (more && next_new_layer) {
              // layer boundary: the new weights land only after these MMAs have completed (the producer waits for them)
              mbar_wait(&cs->wbar, static_cast<uint32_t>(l + 1) & 1u);
              mbar_wait(&cs->tempty[an], ((r * (U / TAP_ACC) + (u + 1) / TAP_ACC) & 1u) ^ 1u);
              mbar_wait(&cs->full[sn], (r * (U / NS) + (u + 1) / NS) & 1u);
              tc_fence_after();
            }
          }
          if (++j == cnt) {
            j = 0;
            ++l;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + FWD_EPI_WARPS) {
    // ===== epilogue: two groups of 8 warps alternate iterations; warp -> TMEM lane quarter and 32-channel half =====
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;
    const int R7 = lane & 7;
    const float inv_p = 1.0f / static_cast<float>(g.P), inv_img = 1.0f / static_cast<float>(g.IMG);
    const bool small = g.NT * TILE_M < (1 << 22);
    const uint32_t st1 = smem_u32(stage1 + grp * TILE_BYTES + row * SLAB_ROW_BYTES);
    const uint32_t st2 = smem_u32(stage2 + grp * TILE_BYTES + row * SLAB_ROW_BYTES);
    for (int it = grp; it < total; it += 2) {
      const int l = it / cnt, j = it - l * cnt;
      const int tile = blockIdx.x + j * grid;
      const int mode = args.layer[l].mode;
      const float* __restrict__ bias = args.layer[l].bias;
      const int a = it % TAP_ACC;
      const int data_row = TILE_M * tile + row;
      const int n_img = small ? static_cast<int>((static_cast<float>(data_row) + 0.5f) * inv_img) : data_row / g.IMG;
      const int q_img = data_row - n_img * g.IMG;
      const int yy = static_cast<int>((static_cast<float>(q_img) + 0.5f) * inv_p);
      const int xx = q_img - yy * g.P;
      const bool valid = data_row < g.data_rows && yy >= 1 && xx >= 1;
      mbar_wait(&cs->tfull[a], (it / TAP_ACC) & 1);
      tc_fence_after();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(args.trace, it, 3);
      const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + a * 64 + half * 32;
      uint32_t acc[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld8(tbase + c * 8, acc[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&cs->tempty[a]);
      mbar_wait(&cs->auxfull[grp], (it >> 1) & 1);
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(args.trace, it, 6);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int jc = half * 4 + jj;
        const int pos = (jc ^ R7) << 4;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(acc[jj][e]);
        if (!BWD) {
          float bj[8];
          const float* bsrc = bias + (args.bias_img_stride ? static_cast<size_t>(valid ? n_img : 0) * args.bias_img_stride : 0) + jc * 8;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bsrc));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bsrc + 4));
          bj[0] = b0.x; bj[1] = b0.y; bj[2] = b0.z; bj[3] = b0.w; bj[4] = b1.x; bj[5] = b1.y; bj[6] = b1.z; bj[7] = b1.w;
          if (mode == MODE_ELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = elu_fast(o[e] + bj[e]);
          } else {        // MODE_RES_ELU
            float sk[8];
            unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = elu_fast(o[e] + bj[e] + sk[e]);
          }
        } else {
          float sk[8];
          unpack8(lds_u4(st1 + pos), sk);
          if (mode == MODE_MUL) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = o[e] * elu_grad_from_out(sk[e]);
          } else {        // MODE_ADD_MUL
            float t[8];
            unpack8(lds_u4(st2 + pos), t);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (o[e] + t[e]) * elu_grad_from_out(sk[e]);
          }
        }
        sts_u4(st1 + pos, valid ? pack8(o) : make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(args.trace, it, 4);
      mbar_arrive(&cs->outready[grp]);
    }
  } else if (warp < 2 + FWD_EPI_WARPS + 2) {
    // ===== two aux / store threads, one per staging buffer (iteration parity): aux tiles in, output tile out, and the
    // tile's flag once its store has LANDED.  One thread doing all three waited ~1.4 us for every store before issuing
    // the next: 3000 cycles per tile against 2150 for the tensor pipe.  Per thread and iteration:
    //   staging buffer free (own store of it-2 read) -> prefetch aux(it) -> store(it-2) landed -> publish(it-2)
    //   -> epilogue(it) done -> store(it).
    // A flag never waits for a later iteration's epilogue: the next layer's producers are waiting for it. =====
    if (lane == 0) {
      const int par = warp - (2 + FWD_EPI_WARPS);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      auto publish = [&](int it) {
        const int l = it / cnt, j = it - l * cnt;
        __threadfence();
        st_release_u32(args.flags + static_cast<size_t>(l) * ntiles + blockIdx.x + j * grid, args.epoch);
      };
      int last = -1;
      for (int it = par; it < total; it += 2) {
        const int l = it / cnt, j = it - l * cnt;
        const int tile = blockIdx.x + j * grid;
        const int mode = args.layer[l].mode;
        // with one tile per layer the aux tile of iteration it was WRITTEN by the store of it-2: it must have landed
        if (cnt < 2) bulk_wait_all<0>();
        else bulk_wait_read<0>();
        const bool a1 = mode == MODE_RES_ELU || mode == MODE_MUL || mode == MODE_ADD_MUL, a2 = mode == MODE_ADD_MUL;
        if (a1) {
          const size_t off = static_cast<size_t>(g.HALO + TILE_M * tile) * SLAB_ROW_BYTES;
          mbar_arrive_expect_tx(&cs->auxfull[par], a2 ? 2 * TILE_BYTES : TILE_BYTES);
          bulk_g2s(stage1 + par * TILE_BYTES, args.layer[l].aux1 + off, TILE_BYTES, &cs->auxfull[par]);
          if (a2) bulk_g2s(stage2 + par * TILE_BYTES, args.layer[l].aux2 + off, TILE_BYTES, &cs->auxfull[par]);
        } else {
          mbar_arrive(&cs->auxfull[par]);
        }
        if (last >= 0) {
          bulk_wait_all<0>();
          publish(last);
        }
        mbar_wait(&cs->outready[par], (it >> 1) & 1);
        bulk_s2g(args.layer[l].out + static_cast<size_t>(g.HALO + TILE_M * tile) * SLAB_ROW_BYTES, stage1 + par * TILE_BYTES,
                 TILE_BYTES);
        bulk_commit();
        trace_at(args.trace, it, 5);
        last = it;
      }
      bulk_wait_all<0>();
      if (last >= 0) publish(last);
    }
  } else {
    // ===== flag warp: sees the three neighbour tiles of the previous layer published, then releases the producer =====
    if (lane == 0) {
      int l = 0, j = 0;
      for (int it = 0; it < total; ++it) {
        // ring of 8: this warp runs up to 8 iterations ahead of the producer, whose loads must not wait for the poll +
        // fences (~2000 cycles).  Entry it & 7 may be re-armed once the producer has consumed it for iteration it - 8.
        // (Not an mbarrier parity wait on empty[]: eight iterations are two phases of a 4-stage ring -- parity aliases.)
        while (it - cs->prod_it >= 8) {
        }
        if (l > 0) {
          const int tile = blockIdx.x + j * grid;
          const unsigned* f = args.flags + static_cast<size_t>(l - 1) * ntiles;
          const int t0 = tile > 0 ? tile - 1 : 0, t1 = tile + 1 < ntiles ? tile + 1 : ntiles - 1;
          unsigned spins = 0;
          for (;;) {
            const unsigned f0 = ld_relaxed_u32(f + t0), f1 = ld_relaxed_u32(f + tile), f2 = ld_relaxed_u32(f + t1);
            if (f0 == args.epoch && f1 == args.epoch && f2 == args.epoch) break;
            if (++spins > (1u << 21)) {
              *args.error = 1u;
              break;
            }
          }
          asm volatile("fence.acq_rel.gpu;" ::: "memory");          // relaxed loads that saw the release stores + fence = acquire
          asm volatile("fence.proxy.async;" ::: "memory");          // the tiles were written / will be read by the async proxy
        }
        mbar_arrive(&cs->flagok[it & 7]);
        if (++j == cnt) {
          j = 0;
          ++l;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64 * TAP_ACC);
  }
}

// ---- pair-fused implicit GEMM (variant 2) ------------------------------------------------------------------------
// The per-tap kernel is bound by the shared-memory port: 36 MMAs x 6 KB of operand reads + ~58 KB of slab / staging
// traffic per 128-row tile = 2140 cycles at 128 B/clk (measured: 2100).  Here the taps (dy, 0) and (dy, +1) share ONE
// activation view in an N = 128 MMA (8 KB per instruction, at the port limit AND the tensor rate), and the tap (dy, -1)
// is an N = 64 MMA on the view shifted by one row that accumulates into the same 64 columns as (dy, 0):
//     D0[q] = sum_dy in[q + dy P - 1] W[dy,-1] + in[q + dy P] W[dy,0]        (columns 0..63)
//     D1[q] = sum_dy in[q + dy P]     W[dy,+1]                               (columns 64..127)
//     out[p] = D0[p] + D1[p + 1]
// 24 MMAs / 168 KB per tile instead of 36 / 216 KB; the epilogue pays ONE warp shuffle per value (half of the dx-fused
// kernel, a third fewer TMEM loads) and tiles overlap by one row (128 computed, 127 stored).
// Measured (tools/bench_conv.py, 640 x 28 x 28): the issuer needs 1250 cycles per tile (per-tap: 1700) but the two-pass
// shuffle epilogue needs ~3700 cycles per tile and group, so the kernel is epilogue-bound at ~1950 cycles per tile:
// 38.8 us in mode 0 (per-tap 39.7) and SLOWER in the modes with auxiliary tiles (44-55 us vs 40-46).  Kept as variant 2
// for A/B work on its epilogue; the per-tap kernel stays the default.
constexpr int PR_OUT_ROWS = 127;
constexpr int PR_OUT_BYTES = PR_OUT_ROWS * SLAB_ROW_BYTES;
constexpr int PR_ACC = 4;                    // 4 x 128 TMEM columns

struct PairSmem {
  uint64_t full[TAP_MAX_STAGES], empty[TAP_MAX_STAGES], tfull[PR_ACC], tempty[PR_ACC], wbar;
  uint64_t auxfull[2], outready[2];
  uint32_t tmem_base;
  alignas(16) float bias[64];
  alignas(16) float xch[2][4][4][EPI_CH];    // [group][lane quarter][channel group][channels]: D1 of the quarter's row 0
};

template <int MODE, int NS>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv3x3_pair_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                    const uint8_t* __restrict__ aux1, const uint8_t* __restrict__ aux2, uint8_t* __restrict__ out,
                    SlabGeom g, int ntiles, int bias_img_stride, long long* trace_buf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int slab_rows = TILE_M + 2 * g.HALO + 8;
  const uint32_t slab_bytes = static_cast<uint32_t>(slab_rows) * SLAB_ROW_BYTES;
  uint8_t* wsm = smem;
  uint8_t* slabs = smem + W_BYTES;
  uint8_t* stage1 = slabs + NS * slab_bytes;                           // [2][STG_BYTES] aux1 in / output out (in place)
  uint8_t* stage2 = stage1 + 2 * STG_BYTES;                            // [2][STG_BYTES] aux2 (MODE_ADD_MUL only)
  PairSmem* cs = reinterpret_cast<PairSmem*>(stage2 + (mode_has_aux2(MODE) ? 2 * STG_BYTES : 0));
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr bool HAS_BIAS = MODE == MODE_ELU || MODE == MODE_RES_ELU || MODE == MODE_LINEAR;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TAP_MAX_STAGES; ++s) {
      mbar_init(&cs->full[s], 1);
      mbar_init(&cs->empty[s], 1);
    }
    for (int a = 0; a < PR_ACC; ++a) {
      mbar_init(&cs->tfull[a], 1);
      mbar_init(&cs->tempty[a], FWD_EPI_WARPS / 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&cs->auxfull[a], 1);
      mbar_init(&cs->outready[a], FWD_EPI_WARPS * 16);
    }
    mbar_init(&cs->wbar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) cs->bias[threadIdx.x - 64] = HAS_BIAS ? bias[threadIdx.x - 64] : 0.f;
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = cs->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile t: computed rows = data rows [127 t, 127 t + 128), stored rows [127 t, 127 t + 127).  First computed = first
  // stored buffer row Rf = HALO + 127 t; slab load starts at LS = floor8(Rf - P - 1); staging row 0 <-> floor8(Rf).
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, wpack, W_BYTES, &cs->wbar);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % NS;
        const uint32_t ph = (it / NS) & 1;
        const int LS = (g.HALO + PR_OUT_ROWS * tile - g.P - 1) & ~7;
        mbar_wait(&cs->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&cs->full[s], slab_bytes);
        bulk_g2s(slabs + s * slab_bytes, in + static_cast<size_t>(LS) * SLAB_ROW_BYTES, slab_bytes, &cs->full[s]);
        trace_at(trace_buf, it, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (uniform loop, elected lane issues; stage / accumulator compile-time, see the per-tap kernel) =====
    constexpr uint32_t idesc128 = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0);
    constexpr int U = (NS == 3) ? 12 : 4;
    mbar_wait(&cs->wbar, 0);
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(wsm), 16, 1024);
    uint64_t adesc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) adesc[s] = umma_desc_sw128(smem_u32(slabs + s * slab_bytes), 16, 1024);
    const uint32_t p8 = static_cast<uint32_t>(g.P) * 8u;
    int tile = blockIdx.x;
    bool more = tile < ntiles;
    const bool leader = elect_one();
    if (more && leader) {
      mbar_wait(&cs->tempty[0], 1u);
      mbar_wait(&cs->full[0], 0u);
      tc_fence_after();
    }
    for (uint32_t r = 0; more; ++r) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (more) {
          const int s = u % NS, a = u % PR_ACC;
          const int sn = (u + 1) % NS, an = (u + 1) % PR_ACC;
          const int it = static_cast<int>(r) * U + u;
          const int Rf = g.HALO + PR_OUT_ROWS * tile;
          const uint32_t a_base = static_cast<uint32_t>(Rf - ((Rf - g.P - 1) & ~7)) * 8u;
          tile += gridDim.x;
          more = tile < ntiles;
          if (leader) {
            trace_at(trace_buf, it, 1);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t a_off = a_base + static_cast<uint32_t>(dy - 1) * p8;
              if (more) {
                if (dy == 1) mbar_wait(&cs->tempty[an], ((r * (U / PR_ACC) + (u + 1) / PR_ACC) & 1u) ^ 1u);
                if (dy == 2) {
                  mbar_wait(&cs->full[sn], (r * (U / NS) + (u + 1) / NS) & 1u);
                  tc_fence_after();
                }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)      // taps (dy, 0) | (dy, +1): columns 0..127
                umma_bf16(tmem + a * 128, adesc[s] + (a_off + 2u * k),
                          bdesc0 + static_cast<uint32_t>(((dy * 3 + 1) * W_TAP_BYTES + k * 32) >> 4), idesc128, (dy | k) != 0);
#pragma unroll
              for (int k = 0; k < 4; ++k)      // tap (dy, -1) on the view one row up: columns 0..63
                umma_bf16(tmem + a * 128, adesc[s] + (a_off - 8u + 2u * k),
                          bdesc0 + static_cast<uint32_t>(((dy * 3) * W_TAP_BYTES + k * 32) >> 4), idesc64, 1u);
            }
            umma_commit(&cs->empty[s]);
            umma_commit(&cs->tfull[a]);
            trace_at(trace_buf, it, 2);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + FWD_EPI_WARPS) {
    // ===== epilogue: two groups of 8 warps alternate tiles; warp -> TMEM lane quarter and 32-channel half, processed as
    // two passes of 16 channels =====
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;              // computed row inside the tile; rows 0..126 are stored
    constexpr int NCH = EPI_CH / 8;
    const float m_dn = lane < 31 ? 1.f : 0.f;   // lane 31 takes row + 1 from the next lane quarter through smem
    const bool stored = row < PR_OUT_ROWS;
    const int tile_stride = 2 * gridDim.x;
    const int step_mod = (PR_OUT_ROWS * tile_stride) % g.IMG;
    const int step_div = (PR_OUT_ROWS * tile_stride) / g.IMG;
    int tile = blockIdx.x + grp * gridDim.x;
    int data_row = PR_OUT_ROWS * tile + row;
    int q_img = data_row % g.IMG;
    int n_img = data_row / g.IMG;
    const float inv_p = 1.0f / static_cast<float>(g.P);
    int it = grp;
    for (; tile < ntiles; tile += tile_stride, it += 2) {
      const int a = it % PR_ACC;
      const int yy = static_cast<int>((static_cast<float>(q_img) + 0.5f) * inv_p);
      const int xx = q_img - yy * g.P;
      const bool valid = stored && data_row < g.data_rows && yy >= 1 && xx >= 1;
      const int Rf = g.HALO + PR_OUT_ROWS * tile;
      const int srow = (Rf & 7) + row;
      const int R7 = (Rf + row) & 7;
      const uint32_t st1 = smem_u32(stage1 + grp * STG_BYTES + srow * SLAB_ROW_BYTES);
      const uint32_t st2 = smem_u32(stage2 + grp * STG_BYTES + srow * SLAB_ROW_BYTES);
      mbar_wait(&cs->tfull[a], (it / PR_ACC) & 1);
      tc_fence_after();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 3);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int cg = half * 2 + p;
        const uint32_t tbase = tmem + (static_cast<uint32_t>(q * 32) << 16) + a * 128 + cg * EPI_CH;
        uint32_t d0[NCH][8], d1[NCH][8];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_ld8(tbase + c * 8, d0[c]);
          tmem_ld8(tbase + 64 + c * 8, d1[c]);
        }
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&cs->tempty[a]);
          if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 6);
        }
        const uint32_t xw = smem_u32(cs->xch[grp][q][cg]);
        float acc[EPI_CH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (lane == 0) {
            sts_u4(xw + c * 32, make_uint4(d1[c][0], d1[c][1], d1[c][2], d1[c][3]));
            sts_u4(xw + c * 32 + 16, make_uint4(d1[c][4], d1[c][5], d1[c][6], d1[c][7]));
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(d1[c][e]), 1);    // dx = +1 term of row + 1
            acc[c * 8 + e] = fmaf(dn, m_dn, __uint_as_float(d0[c][e]));
          }
        }
        // the last row of a lane quarter takes row + 1 from the next quarter (another warp): group barrier, then read
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
        if (lane == 31 && q < 3) {
          const uint32_t xr = smem_u32(cs->xch[grp][q + 1][cg]);
#pragma unroll
          for (int e = 0; e < EPI_CH / 4; ++e) {
            const uint4 t = lds_u4(xr + e * 16);
            acc[4 * e] += __uint_as_float(t.x); acc[4 * e + 1] += __uint_as_float(t.y);
            acc[4 * e + 2] += __uint_as_float(t.z); acc[4 * e + 3] += __uint_as_float(t.w);
          }
        }
        if (p == 0) mbar_wait(&cs->auxfull[grp], (it >> 1) & 1);
        if (stored) {
#pragma unroll
          for (int jj = 0; jj < NCH; ++jj) {
            const int j = cg * NCH + jj;
            const int pos = (j ^ R7) << 4;
            float o[8], bj[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = acc[jj * 8 + e];
            if (HAS_BIAS) {
              const float* bsrc = bias_img_stride ? bias + static_cast<size_t>(valid ? n_img : 0) * bias_img_stride + j * 8
                                                  : &cs->bias[j * 8];
              const float4 b0 = *reinterpret_cast<const float4*>(bsrc);
              const float4 b1 = *reinterpret_cast<const float4*>(bsrc + 4);
              bj[0] = b0.x; bj[1] = b0.y; bj[2] = b0.z; bj[3] = b0.w; bj[4] = b1.x; bj[5] = b1.y; bj[6] = b1.z; bj[7] = b1.w;
            }
            if (MODE == MODE_ELU || MODE == MODE_LINEAR) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float t = o[e] + bj[e];
                o[e] = MODE == MODE_ELU ? elu_fast(t) : t;
              }
            } else if (MODE == MODE_RES_ELU) {
              float sk[8];
              unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = elu_fast(o[e] + bj[e] + sk[e]);
            } else if (MODE == MODE_MUL) {
              float sk[8];
              unpack8(lds_u4(st1 + pos), sk);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = o[e] * elu_grad_from_out(sk[e]);
            } else {
              float sk[8], t[8];
              unpack8(lds_u4(st1 + pos), sk);
              unpack8(lds_u4(st2 + pos), t);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = (o[e] + t[e]) * elu_grad_from_out(sk[e]);
            }
            sts_u4(st1 + pos, valid ? pack8(o) : make_uint4(0, 0, 0, 0));
          }
        }
      }
      fence_proxy_async_smem();
      if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 4);
      mbar_arrive(&cs->outready[grp]);
      data_row += PR_OUT_ROWS * tile_stride;
      q_img += step_mod;
      n_img += step_div;
      if (q_img >= g.IMG) {
        q_img -= g.IMG;
        ++n_img;
      }
    }
  } else {
    // ===== aux / store warp =====
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int it = 0;
      int prev_tile = -1;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        const bool live = tile < ntiles;
        if (live) {
          const int b = it & 1;
          const int Rf = g.HALO + PR_OUT_ROWS * tile;
          bulk_wait_read<0>();
          if (mode_has_aux1(MODE)) {
            const size_t off = static_cast<size_t>(Rf) * SLAB_ROW_BYTES;
            const uint32_t soff = static_cast<uint32_t>(Rf & 7) * SLAB_ROW_BYTES;
            mbar_arrive_expect_tx(&cs->auxfull[b], mode_has_aux2(MODE) ? 2 * PR_OUT_BYTES : PR_OUT_BYTES);
            bulk_g2s(stage1 + b * STG_BYTES + soff, aux1 + off, PR_OUT_BYTES, &cs->auxfull[b]);
            if (mode_has_aux2(MODE)) bulk_g2s(stage2 + b * STG_BYTES + soff, aux2 + off, PR_OUT_BYTES, &cs->auxfull[b]);
          } else {
            mbar_arrive(&cs->auxfull[b]);
          }
        }
        if (prev_tile >= 0) {
          const int pit = it - 1;
          const int pa = pit & 1;
          const int Rf = g.HALO + PR_OUT_ROWS * prev_tile;
          mbar_wait(&cs->outready[pa], (pit >> 1) & 1);
          bulk_s2g(out + static_cast<size_t>(Rf) * SLAB_ROW_BYTES, stage1 + pa * STG_BYTES + (Rf & 7) * SLAB_ROW_BYTES, PR_OUT_BYTES);
          bulk_commit();
          trace_at(trace_buf, pit, 5);
        }
        if (!live) break;
        prev_tile = tile;
      }
      bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 3;
constexpr int ONES_BYTES = 16 * 128;

struct WgradSmem {
  uint64_t full[WG_STAGES], empty[WG_STAGES], done;
  uint32_t tmem_base;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Batched over L layers that share one geometry: layer l reads x + l * slab_stride, gr + l * slab_stride and
// accumulates into dW + l * 9*64*64, db + l * 64.  The grid is split among the layers (ctas_per_layer CTAs each), so a
// whole coupling (16 convolutions) is ONE launch: the per-launch fixed cost (~19 us: launch, pipeline fill, atomic
// tail) is paid once instead of 16 times and each dW sees ctas_per_layer partial sums instead of 148.
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_wgrad_kernel(const uint8_t* __restrict__ x_base, const uint8_t* __restrict__ gr_base, float* __restrict__ dW_base,
                     float* __restrict__ db_base, SlabGeom g, size_t slab_stride, int ctas_per_layer) {
  const int layer = blockIdx.x / ctas_per_layer;
  const int part = blockIdx.x % ctas_per_layer;
  const uint8_t* __restrict__ x = x_base + static_cast<size_t>(layer) * slab_stride;
  const uint8_t* __restrict__ gr = gr_base + static_cast<size_t>(layer) * slab_stride;
  float* __restrict__ dW = dW_base + static_cast<size_t>(layer) * 9 * 64 * 64;
  float* __restrict__ db = db_base + static_cast<size_t>(layer) * 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t xs_bytes = static_cast<uint32_t>(TILE_M + 2 * g.HALO) * SLAB_ROW_BYTES;
  const uint32_t gs_bytes = TILE_M * SLAB_ROW_BYTES;
  const uint32_t stage_bytes = xs_bytes + gs_bytes;       // multiple of 1024
  uint8_t* stages = smem;
  uint8_t* ones = smem + WG_STAGES * stage_bytes;         // 1024-aligned, AFTER the stages (positive LBO)
  WgradSmem* ws = reinterpret_cast<WgradSmem*>(ones + ONES_BYTES);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform

  // constant operand: logical channel 0 of every row = 1.0, everything else 0 (swizzled like a slab row)
  for (int i = threadIdx.x; i < ONES_BYTES / 16; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (c == (r & 7)) val.x = 0x00003f80u;  // bf16 1.0 in element 0
    reinterpret_cast<uint4*>(ones)[i] = val;
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&ws->full[s], 1);
      mbar_init(&ws->empty[s], 1);
    }
    mbar_init(&ws->done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&ws->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ws->tmem_base;

  // contiguous tile range per CTA (keeps the fp32 accumulation order fixed for a given grid)
  const int per = (g.NT + ctas_per_layer - 1) / ctas_per_layer;
  const int t0 = part * per;
  const int t1 = min(t0 + per, g.NT);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = t0; tile < t1; ++tile, ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&ws->empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&ws->full[s], stage_bytes);
        uint8_t* st = stages + s * stage_bytes;
        bulk_g2s(st, x + static_cast<size_t>(tile) * TILE_M * SLAB_ROW_BYTES, xs_bytes, &ws->full[s]);
        bulk_g2s(st + xs_bytes, gr + (static_cast<size_t>(tile) * TILE_M + g.HALO) * SLAB_ROW_BYTES, gs_bytes, &ws->full[s]);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform control flow for the whole warp, one elected lane issues.  Stage, K step and tap pair are all
    // compile-time in the unrolled body and every descriptor is (loop-invariant base) + constant: the issue path is a
    // couple of uniform-datapath instructions per MMA.
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t ones_lo = (smem_u32(ones) & 0x3FFFFu) >> 4;
    uint32_t a_lo[WG_STAGES][5], b_lo[WG_STAGES];   // low descriptor words at K step 0: address | LBO << 16
#pragma unroll
    for (int s = 0; s < WG_STAGES; ++s) {
      const uint32_t xs_lo = (smem_u32(stages + s * stage_bytes) & 0x3FFFFu) >> 4;
      b_lo[s] = (xs_lo + (xs_bytes >> 4)) | (1u << 16);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int tapA = 2 * c, tapB = 2 * c + 1;
        const int offA = g.HALO + (tapA / 3 - 1) * g.P + (tapA % 3 - 1);
        const int offB = g.HALO + (tapB / 3 - 1) * g.P + (tapB % 3 - 1);
        const uint32_t lo = xs_lo + static_cast<uint32_t>(offA) * 8u;
        // LBO (16-byte units): distance from the tap A atom to the tap B atom; for the ninth tap, to the "ones" atom
        // (that distance shrinks by 128 units per K step, see below)
        const uint32_t lbo = c < 4 ? static_cast<uint32_t>(offB - offA) * 8u : ones_lo - lo;
        a_lo[s][c] = lo | (lbo << 16);
      }
    }
    const bool leader = elect_one();
    int tile = t0;
    bool more = tile < t1;
    for (uint32_t r = 0; more; ++r) {
#pragma unroll
      for (int s = 0; s < WG_STAGES; ++s) {
        if (more) {
          if (leader) {
            mbar_wait(&ws->full[s], r & 1u);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TILE_M / 16; ++ks) {
              const uint64_t bdesc = (static_cast<uint64_t>(DESC_HI) << 32) | (b_lo[s] + ks * 128u);
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                // address += 128 units per K step; the ones atom stays put, so its LBO shrinks by the same amount
                const uint32_t lo = a_lo[s][c] + (c < 4 ? ks * 128u : ks * 128u - ((ks * 128u) << 16));
                umma_bf16(tmem + c * 64, (static_cast<uint64_t>(DESC_HI) << 32) | lo, bdesc, idesc, (r | s | ks) != 0);
              }
            }
            umma_commit(&ws->empty[s]);
          }
          ++tile;
          more = tile < t1;
        }
      }
    }
    // the issuer itself waits for the last commit: no asynchronous mbarrier arrive may still be in flight when the CTA
    // exits (it would land in the shared memory of whichever CTA gets this SM next)
    if (leader && t1 > t0) {
      umma_commit(&ws->done);
      mbar_wait(&ws->done, 0);
    }
    __syncwarp();
  } else if (t1 > t0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row: [0,64) = tap A channels, [64,128) = tap B channels
    mbar_wait(&ws->done, 0);
    tc_fence_after();
    for (int c = 0; c < 5; ++c) {
      const int tap = 2 * c + (m >> 6);
      const int ci = m & 63;
      uint32_t v[32];
      for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c * 64 + half * 32, v);
        tmem_ld_wait();
        float* dst = nullptr;
        if (tap < 9) dst = dW + (static_cast<size_t>(tap) * 64 + ci) * 64 + half * 32;
        else if (ci == 0) dst = db + half * 32;
        if (dst) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            red_add_v4(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                       __uint_as_float(v[e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------
// weight-norm pack (forward + transposed/flipped operand for dgrad) and weight-norm backward
//   W[ky,kx,ci,co] = g[co] * V[ky,kx,ci,co] / sqrt(max(sum_{ky,kx,ci} V^2, 1e-12))      nn_extra_nvp.py:389
// ------------------------------------------------------------------------------------------------------------
// grid (L, 3): block (l, t) normalises layer l (every block recomputes the 64 column norms: 147 KB from L2) and packs
// taps 3t .. 3t+2; 1024 threads, four independent accumulators per thread in the norm pass (the single 256-thread block
// per layer walked 144 dependent loads per thread: 21 us, 15 % of a few-shot episode).
__global__ void __launch_bounds__(1024)
conv_wn_pack_kernel(const float* __restrict__ V, const float* __restrict__ gain, uint8_t* __restrict__ wfwd,
                    uint8_t* __restrict__ wbwd) {
  __shared__ float scale[64];
  __shared__ float part[16][64];
  const int l = blockIdx.x;
  const float* v = V + static_cast<size_t>(l) * 9 * 64 * 64;
  const int co = threadIdx.x & 63, grp = threadIdx.x >> 6;       // 16 row groups x 36 rows
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 36; i += 4) {
    const float t0 = v[(grp + 16 * i) * 64 + co], t1 = v[(grp + 16 * (i + 1)) * 64 + co];
    const float t2 = v[(grp + 16 * (i + 2)) * 64 + co], t3 = v[(grp + 16 * (i + 3)) * 64 + co];
    s0 = fmaf(t0, t0, s0); s1 = fmaf(t1, t1, s1); s2 = fmaf(t2, t2, s2); s3 = fmaf(t3, t3, s3);
  }
  part[grp][co] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.x < 64) {
    float n2 = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) n2 += part[q][co];
    scale[co] = gain[l * 64 + co] * rsqrtf(fmaxf(n2, 1e-12f));
  }
  __syncthreads();
  uint8_t* wf = wfwd + static_cast<size_t>(l) * W_BYTES;
  uint8_t* wb = wbwd ? wbwd + static_cast<size_t>(l) * W_BYTES : nullptr;
  // forward operand: [tap][row = co][k = ci]; one thread packs 8 consecutive k
  for (int i = blockIdx.y * 3 * 512 + threadIdx.x; i < (blockIdx.y + 1) * 3 * 512; i += blockDim.x) {
    const int tap = i / 512, rem = i % 512, rowi = rem >> 3, j = rem & 7;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = v[(tap * 64 + (j * 8 + e)) * 64 + rowi] * scale[rowi];
    *reinterpret_cast<uint4*>(wf + tap * W_TAP_BYTES + rowi * 128 + ((j ^ (rowi & 7)) << 4)) = pack8(f);
    if (wb) {
      // dgrad operand: tap' = 8 - tap, [row = ci][k = co]
      float h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = v[(tap * 64 + rowi) * 64 + (j * 8 + e)] * scale[j * 8 + e];
      *reinterpret_cast<uint4*>(wb + (8 - tap) * W_TAP_BYTES + rowi * 128 + ((j ^ (rowi & 7)) << 4)) = pack8(h);
    }
  }
}

// dV, dg from dW (all [L][9][64][64] / [L][64]); K = rows per output channel (576 for 3x3, C for the 1x1 c1)
constexpr int WNB_GROUPS = 16;
__global__ void __launch_bounds__(64 * WNB_GROUPS)
conv_wn_bwd_kernel(const float* __restrict__ V, const float* __restrict__ gain, const float* __restrict__ dW,
                   float* __restrict__ dV, float* __restrict__ dgain, int K) {
  __shared__ float p_n2[WNB_GROUPS][64], p_dot[WNB_GROUPS][64], s_a[64], s_b[64];
  const int l = blockIdx.x;
  const float* v = V + static_cast<size_t>(l) * K * 64;
  const float* dw = dW + static_cast<size_t>(l) * K * 64;
  const int co = threadIdx.x & 63, grp = threadIdx.x >> 6;
  float n2 = 0.f, dot = 0.f;
  for (int r = grp; r < K; r += WNB_GROUPS) {
    const float t = v[r * 64 + co];
    n2 = fmaf(t, t, n2);
    dot = fmaf(t, dw[r * 64 + co], dot);
  }
  p_n2[grp][co] = n2;
  p_dot[grp][co] = dot;
  __syncthreads();
  if (threadIdx.x < 64) {
    float N2 = 0.f, DOT = 0.f;
#pragma unroll
    for (int i = 0; i < WNB_GROUPS; ++i) {
      N2 += p_n2[i][co];
      DOT += p_dot[i][co];
    }
    N2 = fmaxf(N2, 1e-12f);
    const float inv = rsqrtf(N2);
    const float gg = gain[l * 64 + co];
    dgain[l * 64 + co] = DOT * inv;
    s_a[co] = gg * inv;               // dV = a * dW - b * V
    s_b[co] = gg * inv * DOT / N2;
  }
  __syncthreads();
  float* dv = dV + static_cast<size_t>(l) * K * 64;
  for (int r = grp; r < K; r += WNB_GROUPS) dv[r * 64 + co] = s_a[co] * dw[r * 64 + co] - s_b[co] * v[r * 64 + co];
}

static size_t conv_smem_bytes(const SlabGeom& g, int mode = MODE_ADD_MUL) {
  return 1024 + W_BYTES + static_cast<size_t>(NSTAGE) * (TILE_M + 2 * g.HALO + 8) * SLAB_ROW_BYTES +
         (mode_has_aux2(mode) ? 4 : 2) * STG_BYTES + sizeof(ConvSmem) + 64;
}
static size_t wgrad_smem_bytes(const SlabGeom& g) {
  return 1024 + static_cast<size_t>(WG_STAGES) * ((TILE_M + 2 * g.HALO) + TILE_M) * SLAB_ROW_BYTES + ONES_BYTES +
         sizeof(WgradSmem) + 64;
}

static int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

static int g_conv_variant = 1;   // 0: dx-fused N=192 kernel, 1: per-tap N=64 kernel, 2: pair-fused N=128 + N=64 kernel

template <int MODE, int NS>
static int launch_conv_tap_ns(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                              const void* a2, void* out, const SlabGeom& g, size_t smem, cudaStream_t st) {
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_tap_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const int ntiles = g.NT;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  ProfSpan span(BRUNO_PROF_CONV, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FWD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv3x3_tap_kernel<MODE, NS>, static_cast<const uint8_t*>(in), static_cast<const uint8_t*>(w), bias,
                     static_cast<const uint8_t*>(a1), static_cast<const uint8_t*>(a2), static_cast<uint8_t*>(out), g, ntiles,
                     bias_img_stride, g_conv_trace_host);
  return check_launch("conv3x3_tap_kernel");
}

static size_t tap_fixed_smem(int mode) {
  return 1024 + W_BYTES + ((mode_has_aux2(mode) && !TAP_AUX2_DIRECT) ? 4 : 2) * TILE_BYTES + sizeof(TapSmem) + 64;
}
static int tap_stages(const SlabGeom& g, int mode) {
  const size_t slab = static_cast<size_t>(TILE_M + 2 * g.HALO + 8) * SLAB_ROW_BYTES;
  const size_t fixed = tap_fixed_smem(mode);
  if (fixed + 2 * slab > 227 * 1024) return 0;
  const int n = static_cast<int>((227 * 1024 - fixed) / slab);
  return n > TAP_MAX_STAGES ? TAP_MAX_STAGES : n;
}

template <int MODE>
static int launch_conv_tap(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                           const void* a2, void* out, const SlabGeom& g, cudaStream_t st) {
  const int ns = tap_stages(g, MODE);
  BRUNO_REQUIRE(ns >= 2, "bruno_conv3x3: W=%d too wide for the shared-memory slab ring", g.W);
  const size_t smem = tap_fixed_smem(MODE) + static_cast<size_t>(ns) * (TILE_M + 2 * g.HALO + 8) * SLAB_ROW_BYTES;
  if (ns == 4) return launch_conv_tap_ns<MODE, 4>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
  if (ns == 3) return launch_conv_tap_ns<MODE, 3>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
  return launch_conv_tap_ns<MODE, 2>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
}

template <int MODE, int NS>
static int launch_conv_pair_ns(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                               const void* a2, void* out, const SlabGeom& g, size_t smem, cudaStream_t st) {
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_pair_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const int ntiles = (g.data_rows + PR_OUT_ROWS - 1) / PR_OUT_ROWS;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  ProfSpan span(BRUNO_PROF_CONV, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FWD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel<MODE, NS>, static_cast<const uint8_t*>(in), static_cast<const uint8_t*>(w), bias,
                     static_cast<const uint8_t*>(a1), static_cast<const uint8_t*>(a2), static_cast<uint8_t*>(out), g, ntiles,
                     bias_img_stride, g_conv_trace_host);
  return check_launch("conv3x3_pair_kernel");
}

template <int MODE>
static int launch_conv_pair(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                            const void* a2, void* out, const SlabGeom& g, cudaStream_t st) {
  const size_t fixed = 1024 + W_BYTES + (mode_has_aux2(MODE) ? 4 : 2) * STG_BYTES + sizeof(PairSmem) + 64;
  const size_t slab = static_cast<size_t>(TILE_M + 2 * g.HALO + 8) * SLAB_ROW_BYTES;
  int ns = fixed + 2 * slab > 227 * 1024 ? 0 : static_cast<int>((227 * 1024 - fixed) / slab);
  if (ns > TAP_MAX_STAGES) ns = TAP_MAX_STAGES;
  BRUNO_REQUIRE(ns >= 2, "bruno_conv3x3: W=%d too wide for the shared-memory slab ring", g.W);
  const size_t smem = fixed + static_cast<size_t>(ns) * slab;
  if (ns == 4) return launch_conv_pair_ns<MODE, 4>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
  if (ns == 3) return launch_conv_pair_ns<MODE, 3>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
  return launch_conv_pair_ns<MODE, 2>(in, w, bias, bias_img_stride, a1, a2, out, g, smem, st);
}

// ---- chain launcher ----------------------------------------------------------------------------------------------
static int g_conv_chain = 0;                 // bruno_debug_conv_chain: OFF by default (measured slower, see above)
static unsigned* g_chain_flags = nullptr;    // [CHAIN_MAX_LAYERS][g_chain_tiles] epoch flags + 1 error word
static int g_chain_tiles = 0;
static unsigned g_chain_epoch = 0;
static unsigned* g_chain_error_host = nullptr;   // pinned copy of the error word of the previous launch

template <bool BWD, int NS>
static int launch_chain_ns(const ChainArgs& a, const SlabGeom& g, size_t smem, cudaStream_t st) {
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_chain_kernel<BWD, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const int ntiles = g.NT;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  ProfSpan span(BRUNO_PROF_CONV, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv3x3_chain_kernel<BWD, NS>, a, g, ntiles);
  return check_launch("conv3x3_chain_kernel");
}

// L stacked convolutions of one geometry in one launch.  Returns BRUNO_OK, an error, or +1 when the chain path is not
// available (switched off, too many layers, flags cannot be allocated under stream capture, slab too wide): the caller
// then issues the layers one by one.
int conv3x3_chain(bool backward, int L, const void* const* in, const void* const* w, const float* const* bias,
                  const void* const* aux1, const void* const* aux2, void* const* out, const int* mode, int bias_img_stride,
                  int N, int H, int W, cudaStream_t st) {
  if (!g_conv_chain || g_conv_variant != 1 || L < 1 || L > CHAIN_MAX_LAYERS) return 1;
  const SlabGeom g = slab_geom(N, H, W);
  // with fewer than two tiles per CTA every layer is one lock-step round of flag hand-offs: mode 1 leaves those shapes to
  // separate launches, mode 2 chains everything (tests)
  if (g_conv_chain == 1 && g.NT < 2 * num_sms()) return 1;
  static const char* max_tiles_env = getenv("BRUNO_CHAIN_MAX_TILES");     // experiment: chain only the small layers
  if (g_conv_chain == 1 && max_tiles_env && g.NT > atoi(max_tiles_env)) return 1;
  if (g_chain_flags == nullptr || g.NT > g_chain_tiles) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return 1;
    }
    cudaDeviceSynchronize();                 // the old buffer may still be in use
    if (g_chain_flags) cudaFree(g_chain_flags);
    g_chain_tiles = g.NT > 8192 ? g.NT : 8192;
    const size_t words = static_cast<size_t>(CHAIN_MAX_LAYERS) * g_chain_tiles + 1;
    if (cudaMalloc(&g_chain_flags, words * sizeof(unsigned)) != cudaSuccess) {
      cudaGetLastError();
      g_chain_flags = nullptr;
      g_chain_tiles = 0;
      return 1;
    }
    cudaMemset(g_chain_flags, 0, words * sizeof(unsigned));
    if (!g_chain_error_host) {
      cudaMallocHost(&g_chain_error_host, sizeof(unsigned));
      *g_chain_error_host = 0u;
    }
  }
  if (g_chain_error_host && *g_chain_error_host) {
    set_last_error("conv3x3_chain: a tile-flag wait of an earlier launch gave up (inter-layer dependency never published)");
    return BRUNO_ECUDA;
  }
  const size_t fixed = 1024 + W_BYTES + (backward ? 4 : 2) * TILE_BYTES + sizeof(TapSmem) + 64;
  const size_t slab = static_cast<size_t>(TILE_M + 2 * g.HALO + 8) * SLAB_ROW_BYTES;
  if (fixed + 2 * slab > 227 * 1024) return 1;
  int ns = static_cast<int>((227 * 1024 - fixed) / slab);
  if (ns > TAP_MAX_STAGES) ns = TAP_MAX_STAGES;
  ChainArgs a{};
  for (int l = 0; l < L; ++l) {
    a.layer[l].in = static_cast<const uint8_t*>(in[l]);
    a.layer[l].w = static_cast<const uint8_t*>(w[l]);
    a.layer[l].bias = bias ? bias[l] : nullptr;
    a.layer[l].aux1 = aux1 ? static_cast<const uint8_t*>(aux1[l]) : nullptr;
    a.layer[l].aux2 = aux2 ? static_cast<const uint8_t*>(aux2[l]) : nullptr;
    a.layer[l].out = static_cast<uint8_t*>(out[l]);
    a.layer[l].mode = mode[l];
  }
  a.L = L;
  a.bias_img_stride = bias_img_stride;
  a.epoch = ++g_chain_epoch;
  if (g_chain_epoch == 0xFFFFFFFFu) {        // wrap: clear the flags once every 4 billion launches
    cudaMemsetAsync(g_chain_flags, 0, (static_cast<size_t>(CHAIN_MAX_LAYERS) * g_chain_tiles) * sizeof(unsigned), st);
    g_chain_epoch = 0;
    a.epoch = ++g_chain_epoch;
  }
  a.flags = g_chain_flags;
  a.error = g_chain_flags + static_cast<size_t>(CHAIN_MAX_LAYERS) * g_chain_tiles;
  a.trace = g_conv_trace_host;
  // the flag rows are [L][ntiles] with the CURRENT ntiles as pitch (a launch only reads what it wrote)
  const size_t smem = fixed + static_cast<size_t>(ns) * slab;
  int rc;
  if (backward) rc = ns >= 4 ? launch_chain_ns<true, 4>(a, g, smem, st) : ns == 3 ? launch_chain_ns<true, 3>(a, g, smem, st)
                                                                                  : launch_chain_ns<true, 2>(a, g, smem, st);
  else rc = ns >= 4 ? launch_chain_ns<false, 4>(a, g, smem, st) : ns == 3 ? launch_chain_ns<false, 3>(a, g, smem, st)
                                                                         : launch_chain_ns<false, 2>(a, g, smem, st);
  // the error word is looked at by the NEXT call (no synchronisation here)
  if (rc == BRUNO_OK) cudaMemcpyAsync(g_chain_error_host, a.error, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
  return rc;
}

template <int MODE>
static int launch_conv(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                       const void* a2, void* out, const SlabGeom& g, cudaStream_t st) {
  if (g_conv_variant == 1) return launch_conv_tap<MODE>(in, w, bias, bias_img_stride, a1, a2, out, g, st);
  if (g_conv_variant == 2) return launch_conv_pair<MODE>(in, w, bias, bias_img_stride, a1, a2, out, g, st);
  const size_t smem = conv_smem_bytes(g, MODE);
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const int ntiles = (g.data_rows + OUT_ROWS - 1) / OUT_ROWS;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  ProfSpan span(BRUNO_PROF_CONV, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FWD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv3x3_kernel<MODE>, static_cast<const uint8_t*>(in), static_cast<const uint8_t*>(w), bias,
                     static_cast<const uint8_t*>(a1), static_cast<const uint8_t*>(a2), static_cast<uint8_t*>(out), g, ntiles,
                     bias_img_stride, g_conv_trace_host);
  return check_launch("conv3x3_kernel");
}

}  // namespace bruno

using namespace bruno;

extern "C" {

size_t bruno_slab_bytes(int N, int H, int W) {
  const SlabGeom g = slab_geom(N, H, W);
  return static_cast<size_t>(g.rows) * SLAB_ROW_BYTES;
}

size_t bruno_conv_wpack_bytes(void) { return W_BYTES; }

int bruno_debug_conv_trace(long long* device_buffer) {
  g_conv_trace_host = device_buffer;
  return BRUNO_OK;
}

int bruno_debug_conv_chain(int on) {
  if (on < 0) return g_conv_chain;
  g_conv_chain = on > 2 ? 1 : on;            // 0 off, 1 on for >= 2 tiles per CTA (default), 2 on for every shape (tests)
  return BRUNO_OK;
}

int bruno_debug_conv_variant(int variant) {
  if (variant < 0) return g_conv_variant;
  BRUNO_REQUIRE(variant >= 0 && variant <= 2, "bruno_debug_conv_variant: 0 (dx-fused), 1 (per-tap) or 2 (pair-fused)");
  g_conv_variant = variant;
  return BRUNO_OK;
}

int bruno_conv_wn_pack(const float* V, const float* g, void* wfwd, void* wbwd, int L, void* stream) {
  BRUNO_REQUIRE(V && g && wfwd && L > 0, "bruno_conv_wn_pack: bad arguments");
  BRUNO_REQUIRE(aligned16(wfwd) && (!wbwd || aligned16(wbwd)), "bruno_conv_wn_pack: packed operands must be 16-byte aligned");
  conv_wn_pack_kernel<<<dim3(L, 3), 1024, 0, static_cast<cudaStream_t>(stream)>>>(V, g, static_cast<uint8_t*>(wfwd),
                                                                        static_cast<uint8_t*>(wbwd));
  return check_launch("conv_wn_pack_kernel");
}

int bruno_conv_wn_bwd(const float* V, const float* g, const float* dW, float* dV, float* dg, int L, int K, void* stream) {
  BRUNO_REQUIRE(V && g && dW && dV && dg && L > 0 && K > 0, "bruno_conv_wn_bwd: bad arguments");
  conv_wn_bwd_kernel<<<L, 64 * WNB_GROUPS, 0, static_cast<cudaStream_t>(stream)>>>(V, g, dW, dV, dg, K);
  return check_launch("conv_wn_bwd_kernel");
}

int bruno_conv3x3(int mode, const void* in_slab, const void* wpack, const float* bias, int bias_img_stride,
                  const void* aux1, const void* aux2, void* out_slab, int N, int H, int W, void* stream) {
  BRUNO_REQUIRE(bias_img_stride == 0 || bias_img_stride >= 64, "bruno_conv3x3: bias_img_stride must be 0 or >= 64");
  BRUNO_REQUIRE(in_slab && wpack && out_slab && N > 0 && H > 0 && W > 0, "bruno_conv3x3: bad arguments");
  BRUNO_REQUIRE(aligned16(in_slab) && aligned16(wpack) && aligned16(out_slab), "bruno_conv3x3: 16-byte alignment required");
  BRUNO_REQUIRE(in_slab != out_slab, "bruno_conv3x3: in-place convolution is not possible");
  const SlabGeom g = slab_geom(N, H, W);
  BRUNO_REQUIRE(g_conv_variant != 0 || conv_smem_bytes(g) <= 227 * 1024, "bruno_conv3x3: W=%d too wide for the shared-memory slab", W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (mode) {
    case MODE_ELU:
      BRUNO_REQUIRE(bias, "bruno_conv3x3: bias required");
      return launch_conv<MODE_ELU>(in_slab, wpack, bias, bias_img_stride, nullptr, nullptr, out_slab, g, st);
    case MODE_RES_ELU:
      BRUNO_REQUIRE(bias && aux1, "bruno_conv3x3: bias and skip slab required");
      return launch_conv<MODE_RES_ELU>(in_slab, wpack, bias, bias_img_stride, aux1, nullptr, out_slab, g, st);
    case MODE_MUL:
      BRUNO_REQUIRE(aux1, "bruno_conv3x3: saved-activation slab required");
      return launch_conv<MODE_MUL>(in_slab, wpack, nullptr, 0, aux1, nullptr, out_slab, g, st);
    case MODE_ADD_MUL:
      BRUNO_REQUIRE(aux1 && aux2, "bruno_conv3x3: saved-activation and addend slabs required");
      return launch_conv<MODE_ADD_MUL>(in_slab, wpack, nullptr, 0, aux1, aux2, out_slab, g, st);
    case MODE_LINEAR:
      BRUNO_REQUIRE(bias, "bruno_conv3x3: bias required");
      return launch_conv<MODE_LINEAR>(in_slab, wpack, bias, bias_img_stride, nullptr, nullptr, out_slab, g, st);
    default:
      set_last_error("bruno_conv3x3: unknown mode %d", mode);
      return BRUNO_EINVAL;
  }
}

int bruno_conv3x3_wgrad(const void* x_slab, const void* g_slab, float* dW, float* db, int N, int H, int W, void* stream) {
  return bruno_conv3x3_wgrad_batched(x_slab, g_slab, dW, db, 1, N, H, W, stream);
}

int bruno_conv3x3_wgrad_batched(const void* x_slab, const void* g_slab, float* dW, float* db, int L, int N, int H, int W,
                                void* stream) {
  BRUNO_REQUIRE(x_slab && g_slab && dW && db && L > 0 && N > 0 && H > 0 && W > 0, "bruno_conv3x3_wgrad: bad arguments");
  BRUNO_REQUIRE(aligned16(x_slab) && aligned16(g_slab) && aligned16(dW) && aligned16(db), "bruno_conv3x3_wgrad: alignment");
  const SlabGeom g = slab_geom(N, H, W);
  const size_t smem = wgrad_smem_bytes(g);
  BRUNO_REQUIRE(smem <= 227 * 1024, "bruno_conv3x3_wgrad: W=%d too wide", W);
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  int per_layer = num_sms() / L;
  if (per_layer < 1) per_layer = 1;
  if (per_layer > g.NT) per_layer = g.NT;
  const size_t slab_stride = static_cast<size_t>(g.rows) * SLAB_ROW_BYTES;
  ProfSpan span(BRUNO_PROF_WGRAD, static_cast<cudaStream_t>(stream));
  conv3x3_wgrad_kernel<<<per_layer * L, CONV_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(x_slab), static_cast<const uint8_t*>(g_slab), dW, db, g, slab_stride, per_layer);
  return check_launch("conv3x3_wgrad_kernel");
}

}  // extern "C"
