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
// Packed operand order: tap (dy, dx) = (tap / 3 - 1, tap % 3 - 1) lives in 8 KB block (dx + 1) * 3 + (1 - dy), i.e. the
// three vertical taps of one horizontal offset are 192 consecutive operand rows ordered dy = +1, 0, -1: the line-tile
// kernel multiplies one input line by all three in ONE N = 192 instruction (output lines y-1, y, y+1 = ascending
// accumulator slots).
__host__ __device__ constexpr int wpack_block(int tap) { return (tap % 3) * 3 + (2 - tap / 3); }
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
  if (buf != nullptr && blockIdx.x == 0 && it < 64) {
    buf[it * 16 + slot] = clock64();
    if (slot == 8 || slot == 12) {      // kernel entry / end also in ns: (cycles / ns) = the SM clock the launch really ran at
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      buf[it * 16 + (slot == 8 ? 13 : 14)] = static_cast<long long>(ns);
    }
  }
}

constexpr int FWD_EPI_WARPS = 16;            // 4 per TMEM lane quarter: each handles 32 rows x 16 channels
constexpr int FWD_THREADS = 32 * (2 + FWD_EPI_WARPS + 1);   // producer, MMA issuer, epilogue warps, aux/store warp
constexpr int TILE_BYTES = TILE_M * SLAB_ROW_BYTES;          // 16 KB: one aux / gradient tile of the wgrad kernel
__host__ __device__ constexpr bool mode_has_aux1(int m) { return m == MODE_RES_ELU || m == MODE_MUL || m == MODE_ADD_MUL; }
__host__ __device__ constexpr bool mode_has_aux2(int m) { return m == MODE_ADD_MUL; }

__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ELU with one MUFU: exp(x) - 1 = 2^(x log2 e) - 1 (abs. error ~1e-7, far below the bf16 output quantum)
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : ex2_ftz(x * 1.4426950408889634f) - 1.f; }

// ---- per-tap implicit GEMM ----------------------------------------------------------------------------------------
// Nine row-shifted views x 4 K steps = 36 MMAs of N = 64 accumulate into ONE 64-column accumulator, so the epilogue is
// a plain row-wise map (no horizontal shuffles, no boundary exchange, no overlapping tiles).  The MMAs themselves are
// shared-memory-port bound (6 KB per 128x64x16 instruction = 49 cycles, 65 % of the tensor peak).  Two
// variants that trade MMA count for epilogue work (three horizontal taps in one N = 192 MMA; two taps in one N = 128 MMA)
// and a flag-synchronised multi-layer launch were built, measured slower in round 1 and removed in round 2
// (profiles/r01_*; git history before the round-2 clean-up).
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
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 8);                       // kernel entry

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
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 9);                       // barriers + TMEM set up

  // tile t = data rows [128 t, 128 t + 128) = buffer rows from Rf = HALO + 128 t (a multiple of 8); the slab load starts
  // at LS = floor8(Rf - P - 1) and covers every tap's view.
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, wpack, W_BYTES, &cs->wbar);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      trace_at(trace_buf, 0, 10);                                        // previous grid complete
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
    if (lane == 0) trace_at(trace_buf, 0, 11);                           // weights landed
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
                          bdesc0 + static_cast<uint32_t>((wpack_block(tap) * W_TAP_BYTES + k * 32) >> 4), idesc, (tap | k) != 0);
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
      const bool valid = data_row < g.data_rows && yy >= 1 && xx >= 1 && xx <= g.W;
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
      // (waiting only for the staging-buffer READS here, cp.async.bulk.wait_group.read, was measured: no change -- the
      // dependent grid's griddepcontrol.wait waits for the writes anyway)
      bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 12);                      // every role done (stores landed)
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64 * TAP_ACC);
  }
}

// ---- line-tile implicit GEMM (default) ----------------------------------------------------------------------------
// The per-tap kernel above is bound by the shared-memory port: every M128 x N64 x K16 instruction re-reads 4 KB of A and
// 2 KB of B for 32 cycles of tensor work, and N cannot grow because Cout = 64 -- unless several TAPS share one A view.
// Taps that share a view produce outputs for DIFFERENT positions, which the earlier fused variants paid for with warp
// shuffles in the epilogue.  Here the shift is absorbed by the tile geometry instead:
//   * an MMA tile is ONE IMAGE LINE of G = 128 / P consecutive images (P = slab row pitch, a multiple of 8: G bulk copies
//     of P rows each land at rows i * P of the stage, swizzle phase preserved), so "one line down" is "the next tile, same
//     TMEM lanes";
//   * input line y times the 192-row operand [W(dy=+1, dx); W(dy=0, dx); W(dy=-1, dx)] is ONE N = 192 instruction whose
//     three 64-column blocks are the contributions to output lines y-1, y, y+1;
//   * output lines own consecutive 64-column TMEM slots (ring of 8), so those three blocks land in three consecutive
//     slots: each slot accumulates its line from the three input lines around it, in the order dy = -1, 0, +1, dx, k --
//     the SAME summation order as the per-tap kernel (the two kernels agree bit for bit), and the epilogue stays a plain
//     row-wise map.
// 12 instructions of N = 192 (96 tensor cycles, 10 KB of operands = 80 port cycles each) replace 36 of N = 64 (32 tensor
// cycles, 6 KB = 48 port cycles): the MMAs of a tile are tensor-bound, no slab halo is loaded (a line needs only its
// left / right neighbour rows = the zero padding columns), and zero padding lines cost no MMA at all.
// A CTA walks a contiguous range of output lines; where its range starts / ends inside an image group it also multiplies
// the neighbouring input line by the one block it needs (N = 64).  The first instruction of an input line is split where
// a slot sees its first contribution (overwrite) next to slots that accumulate, and where the slot ring wraps.
constexpr int LT_SLOTS = 8;                                       // 64-column accumulators = all 512 TMEM columns
constexpr int LT_GUARD = 8;                                       // zero rows before / after the 128 tile rows of a stage
constexpr int LT_STAGE_BYTES = (TILE_M + 2 * LT_GUARD) * SLAB_ROW_BYTES;   // 18 KB
constexpr int LT_MAX_STAGES = 4;

constexpr int LT_THREADS = 32 * (2 + FWD_EPI_WARPS + 4);   // producer, MMA issuer, epilogue warps, aux1 loader, output drain, gatekeeper, aux2 loader
constexpr int LT_GATES = 4;
constexpr int LT_MAX_NB = 4;                                // ring of "tile may be issued" barriers (gatekeeper -> issuer)

struct LineSmem {
  uint64_t full[LT_MAX_STAGES], empty[LT_MAX_STAGES], tfull[LT_SLOTS], tempty[LT_SLOTS], wbar;
  uint64_t auxfull[LT_MAX_NB], outready[LT_MAX_NB], outfree[LT_MAX_NB], go[LT_GATES], aux2full[2], aux2free[2];
  uint32_t tmem_base;
  alignas(16) float bias[64];
};
// Staging ring: NB buffers of one tile per stream (aux1 / output IN PLACE, aux2).  An aux tile is requested as soon as the
// store of the tile NB back has read the buffer, i.e. NB - 1 tile periods before its epilogue needs it: with the HBM
// latency at 2-3 tile periods, NB = 2 (one buffer per epilogue group) exposes it in every epilogue.
// NB must be a multiple of the number of epilogue groups (2): the groups wait on auxfull[it % NB] by phase PARITY, which is
// only safe if consecutive users of one buffer are the same group -- with NB = 3 a buffer alternates between the groups,
// a group that runs a tile ahead can reach the buffer while its previous phase is still open, its parity wait falls
// through, and two epilogues share a buffer (observed: MODE_ADD_MUL at 640x28x28, where HBM-bound aux tiles land out of
// order, failed on back-to-back launches with NB = 3 and only then).
// MODE_ADD_MUL: three input stages, four aux1 / output buffers and two addend buffers (one per epilogue group): an epilogue
// copies its addend tile to registers first thing and frees the buffer, so the addend of the group's next tile streams in
// during the whole epilogue.
__host__ __device__ constexpr int line_nb(int mode) { return mode_has_aux1(mode) ? 4 : 2; }
__host__ __device__ constexpr int line_ns(int mode) { return mode_has_aux2(mode) ? 3 : LT_MAX_STAGES; }
static_assert(line_nb(MODE_ADD_MUL) % 2 == 0 && line_nb(MODE_MUL) % 2 == 0 && line_nb(MODE_ELU) % 2 == 0, "see above");
__host__ __device__ constexpr int line_stage_tiles(int mode) { return line_nb(mode) + (mode_has_aux2(mode) ? 2 : 0); }

// This CTA's share [t, t1) of the output line tiles (linear index = image group * H + data line - 1), walked segment by
// segment; a segment = consecutive output lines yo0..yo1 (1-based data lines) of ONE image group j.  Its input lines are
// yi0()..yi1(): one more on either side unless that side is the zero padding line.
struct LineWalk {
  int H, t, t1, j, yo0, yo1;
  __device__ LineWalk(int H_, int t0_, int t1_) : H(H_), t(t0_), t1(t1_), j(0), yo0(0), yo1(0) {}
  __device__ bool next() {
    if (t >= t1) return false;
    j = t / H;
    yo0 = t - j * H + 1;
    const int n = min(H - yo0 + 1, t1 - t);
    yo1 = yo0 + n - 1;
    t += n;
    return true;
  }
  __device__ int yi0() const { return max(1, yo0 - 1); }
  __device__ int yi1() const { return min(H, yo1 + 1); }
};

template <int MODE, int NS>
__global__ void __launch_bounds__(LT_THREADS, 1)
conv3x3_line_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                    const uint8_t* __restrict__ aux1, const uint8_t* __restrict__ aux2, uint8_t* __restrict__ out,
                    SlabGeom g, int G, int n_lines, int bias_img_stride, long long* trace_buf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* wsm = smem;
  uint8_t* slabs = smem + W_BYTES;                                     // [NS][guard | 128 tile rows | guard]
  constexpr int NB = line_nb(MODE);
  uint8_t* stage1 = slabs + NS * LT_STAGE_BYTES;                       // [NB][TILE_BYTES] aux1 tile in / output tile out (in place)
  uint8_t* stage2 = stage1 + NB * TILE_BYTES;                          // [2][TILE_BYTES] addend tiles (MODE_ADD_MUL only), one per epilogue group
  LineSmem* cs = reinterpret_cast<LineSmem*>(stage1 + line_stage_tiles(MODE) * TILE_BYTES);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  constexpr bool HAS_BIAS = MODE == MODE_ELU || MODE == MODE_RES_ELU || MODE == MODE_LINEAR;
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 8);                       // kernel entry

  // barriers initialised by five threads of different warps in parallel (41 serial mbarrier.init in one thread were ~800
  // cycles of every launch's prologue)
  {
    const int t = threadIdx.x;
    if (t == 0) {
      for (int i = 0; i < LT_MAX_STAGES; ++i) {
        mbar_init(&cs->full[i], 1);
        mbar_init(&cs->empty[i], 1);
      }
      mbar_init(&cs->wbar, 1);
      fence_mbar_init();
    } else if (t == 32) {
      for (int i = 0; i < LT_SLOTS; ++i) mbar_init(&cs->tfull[i], 1);
      fence_mbar_init();
    } else if (t == 64) {
      for (int i = 0; i < LT_SLOTS; ++i) mbar_init(&cs->tempty[i], FWD_EPI_WARPS / 2);
      fence_mbar_init();
    } else if (t == 96) {
      for (int i = 0; i < LT_MAX_NB; ++i) {
        mbar_init(&cs->auxfull[i], 1);
        mbar_init(&cs->outready[i], FWD_EPI_WARPS * 16);
        mbar_init(&cs->outfree[i], 1);
      }
      fence_mbar_init();
    } else if (t == 128) {
      for (int i = 0; i < LT_GATES; ++i) mbar_init(&cs->go[i], 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&cs->aux2full[i], 1);
        mbar_init(&cs->aux2free[i], FWD_EPI_WARPS * 16);
      }
      fence_mbar_init();
    }
  }
  if (threadIdx.x == 0) trace_at(trace_buf, 62, 0);                      // barriers initialised (this thread's share)
  if (threadIdx.x >= 192 && threadIdx.x < 256) cs->bias[threadIdx.x - 192] = HAS_BIAS ? bias[threadIdx.x - 192] : 0.f;
  // Rows of a stage that no copy ever writes stay zero for the whole kernel: the guards, the padding columns past W + 1
  // of every image segment (only the rows a line's neighbours can read are copied) and the rows past the last segment.
  for (int i = threadIdx.x; i < NS * (LT_STAGE_BYTES / 16); i += blockDim.x) reinterpret_cast<uint4*>(slabs)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) trace_at(trace_buf, 62, 1);                      // stage rows zeroed
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, 64 * LT_SLOTS);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) trace_at(trace_buf, 62, 2);                      // TMEM allocated
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = cs->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 9);                       // barriers + TMEM set up

  const int t0 = static_cast<int>(static_cast<long long>(blockIdx.x) * n_lines / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * n_lines / gridDim.x);
  const uint32_t seg_bytes = static_cast<uint32_t>(g.P) * SLAB_ROW_BYTES;
  // HBM traffic: an input line needs its W pixels and one padding row on either side (rows 0 .. W + 1 of the segment);
  // aux tiles and the output need the W pixel rows only (the padding rows of a slab are never written: they stay zero)
  const uint32_t in_bytes = static_cast<uint32_t>(min(g.P, g.W + 2)) * SLAB_ROW_BYTES;
  const uint32_t px_bytes = static_cast<uint32_t>(g.W) * SLAB_ROW_BYTES;
  const size_t img_bytes = static_cast<size_t>(g.IMG) * SLAB_ROW_BYTES;
  // byte offset of line `line` (0 = the zero line above the image) of image n in a slab
  auto line_off = [&](int n, int line) {
    return (static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + static_cast<size_t>(line) * g.P) * SLAB_ROW_BYTES;
  };

  if (warp == 0) {
    // ===== producer: one input line of G images per stage = G bulk copies of P rows =====
    if (lane == 0) {
      const uint64_t pol_first = l2_policy_evict_first();
      mbar_arrive_expect_tx(&cs->wbar, W_BYTES);
      bulk_g2s(wsm, wpack, W_BYTES, &cs->wbar);
      LineWalk w(g.H, t0, t1);
      int ii = 0;
      bool waited = false;                                               // the segment walk / addresses of the first tile are
      while (w.next()) {                                                 // computed BEFORE waiting for the previous grid
        for (int yi = w.yi0(); yi <= w.yi1(); ++yi, ++ii) {
          const int s = ii % NS;
          const int geff = min(G, g.N - w.j * G);
          uint8_t* dst = slabs + s * LT_STAGE_BYTES + LT_GUARD * SLAB_ROW_BYTES;
          const uint8_t* src = in + line_off(w.j * G, yi);
          if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            trace_at(trace_buf, 0, 10);                                  // previous grid complete
            waited = true;
          }
          mbar_wait(&cs->empty[s], ((ii / NS) & 1) ^ 1);
          mbar_arrive_expect_tx(&cs->full[s], static_cast<uint32_t>(geff) * in_bytes + (geff < G ? seg_bytes : 0u));
          trace_at(trace_buf, ii, 0);
          // MODE_RES_ELU / MODE_ADD_MUL: this launch is the last reader of its input slab for now (c3's input u is next needed
          // in the backward pass, the gradient g2 never again): evict first, keep L2 for the slabs the next launches re-read
          if (MODE == MODE_RES_ELU || MODE == MODE_ADD_MUL) {
            for (int i = 0; i < geff; ++i, dst += seg_bytes, src += img_bytes) bulk_g2s_hint(dst, src, in_bytes, &cs->full[s], pol_first);
          } else {
            for (int i = 0; i < geff; ++i, dst += seg_bytes, src += img_bytes) bulk_g2s(dst, src, in_bytes, &cs->full[s]);
          }
          // the segment after the last image of a partial group holds that image's right-hand neighbours: zeros (the
          // slab's leading guard rows), whatever an earlier full group left there
          if (geff < G) bulk_g2s(dst, in, seg_bytes, &cs->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one elected lane.  Measured (tools/umma_bench.cu): the MMAs of a tile take 1250 cycles in
    // isolation and every instruction the issuing thread executes between them adds its full latency (branches ~8 cycles)
    // -- the tensor pipe does not run ahead of the issuer far enough to hide them.  So the issuer does the bare minimum:
    // ONE barrier wait per tile (the gatekeeper thread has already waited for the input line and for the accumulator
    // slots), incremental uniform-register bookkeeping, and for the common case -- a line with both neighbours inside this
    // CTA's segment -- straight-line code per accumulator-ring position class (no wrap / wrap after two blocks / wrap
    // after one). =====
    mbar_wait(&cs->wbar, 0);
    if (lane == 0) trace_at(trace_buf, 0, 11);                           // weights landed
    if (elect_one()) {
      constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0), idesc128 = umma_idesc_bf16(128, 128, 0, 0),
                         idesc192 = umma_idesc_bf16(128, 192, 0, 0);
      constexpr uint32_t BLK = W_TAP_BYTES >> 4;                         // one 64-row operand block in descriptor units
      const uint64_t bdesc0 = umma_desc_sw128(smem_u32(wsm), 16, 1024);
      const uint64_t adesc0 = umma_desc_sw128(smem_u32(slabs + LT_GUARD * SLAB_ROW_BYTES), 16, 1024);
      const uint32_t empty0 = smem_u32(&cs->empty[0]), tfull0 = smem_u32(&cs->tfull[0]), go0 = smem_u32(&cs->go[0]);
      auto commit_addr = [](uint32_t bar) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      };
      auto wait_addr = [](uint32_t bar, uint32_t parity) {
        uint32_t ok;
        do {
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(ok)
              : "r"(bar), "r"(parity)
              : "memory");
        } while (!ok);
      };
      // steps 1..11 of a tile: (dx, k) = A view shifted by dx rows (8 descriptor units) at K offset 32 k bytes (2 units)
      // times the operand of dx at the same K offset; run 0 = n0 blocks from block b_lo into d0, run 1 (n1 > 0: the slot
      // ring wrapped) = the following n1 blocks into slot 0
      auto steps = [&](uint64_t adesc, uint64_t bq, uint32_t d0, uint32_t id0, uint32_t n0, uint32_t id1, bool two) {
#pragma unroll
        for (int st = 1; st < 12; ++st) {
          const int dxi = st >> 2, k = st & 3;
          const uint64_t a = adesc + static_cast<uint64_t>(static_cast<int64_t>((dxi - 1) * 8 + 2 * k));
          const uint32_t bo = static_cast<uint32_t>((dxi * 3 * W_TAP_BYTES + k * 32) >> 4);
          umma_bf16(d0, a, bq + bo, id0, 1u);
          if (two) umma_bf16(tmem, a, bq + (bo + n0 * BLK), id1, 1u);
        }
      };
      LineWalk w(g.H, t0, t1);
      int ii = 0, it_seg = 0, sidx = 0, tr_i = 0;
      uint32_t a_off = 0;                                                // descriptor offset of the current stage
      while (w.next()) {
        const int yo0 = w.yo0, yo1 = w.yo1, yi1 = w.yi1();
        int yi = w.yi0();
        int it0 = it_seg + (yi - 1 - yo0);                               // output-tile index (within the CTA) of block 0
        for (; yi <= yi1; ++yi, ++it0, ++ii) {
          wait_addr(go0 + 8u * (ii & (LT_GATES - 1)), (ii / LT_GATES) & 1);
          tc_fence_after();
          trace_at(trace_buf, tr_i, 1);
          const uint64_t adesc = adesc0 + a_off;
          const uint64_t a1 = adesc - 8u;                                // first step: dx = -1, k = 0
          const uint32_t slot = it0 & 7;
          const uint32_t d0 = tmem + slot * 64;
          if (yi > yo0 && yi < yo1 && yi > 1) {
            // interior line: blocks 0, 1 accumulate, block 2 opens a fresh slot; all three valid; line yi - 1 completes
            if (slot <= 5) {
              umma_bf16(d0, a1, bdesc0, idesc128, 1u);
              umma_bf16(d0 + 128, a1, bdesc0 + 2 * BLK, idesc64, 0u);
              steps(adesc, bdesc0, d0, idesc192, 3, 0, false);
            } else if (slot == 6) {
              umma_bf16(d0, a1, bdesc0, idesc128, 1u);
              umma_bf16(tmem, a1, bdesc0 + 2 * BLK, idesc64, 0u);
              steps(adesc, bdesc0, d0, idesc128, 2, idesc64, true);
            } else {
              umma_bf16(d0, a1, bdesc0, idesc64, 1u);
              umma_bf16(tmem, a1, bdesc0 + BLK, idesc64, 1u);
              umma_bf16(tmem + 64, a1, bdesc0 + 2 * BLK, idesc64, 0u);
              steps(adesc, bdesc0, d0, idesc64, 1, idesc128, true);
            }
            commit_addr(empty0 + 8u * sidx);
            commit_addr(tfull0 + 8u * slot);
          } else {
            // edge line: block b = output line yi - 1 + b is valid if that line is in this segment
            const int b_lo = yi - 1 >= yo0 ? 0 : (yi >= yo0 ? 1 : 2);
            const int b_hi = yi + 1 <= yo1 ? 2 : (yi <= yo1 ? 1 : 0);
            const int bf = yi == 1 ? 1 : 2;                              // blocks >= bf see their FIRST contribution here
            const uint32_t slot0 = (it0 + b_lo) & 7, nbt = b_hi - b_lo + 1;
            const uint32_t n0 = min(nbt, 8u - slot0), n1 = nbt - n0;     // blocks before / after the slot ring wraps
#pragma unroll
            for (int b = 0; b < 3; ++b)
              if (b >= b_lo && b <= b_hi)
                umma_bf16(tmem + static_cast<uint32_t>(((it0 + b) & 7) * 64), a1, bdesc0 + b * BLK, idesc64, b < bf ? 1u : 0u);
            steps(adesc, bdesc0 + b_lo * BLK, tmem + slot0 * 64, idesc64 + (n0 - 1) * (8u << 17), n0, idesc64 + (n1 - 1) * (8u << 17),
                  n1 != 0);
            commit_addr(empty0 + 8u * sidx);
            if (b_lo == 0) commit_addr(tfull0 + 8u * slot);              // line yi - 1 has all three contributions
            if (yi == g.H && b_lo <= 1 && b_hi >= 1) commit_addr(tfull0 + 8u * ((it0 + 1) & 7));   // last line of the image
          }
          trace_at(trace_buf, tr_i, 2);
          ++tr_i;
          if (++sidx == NS) {
            sidx = 0;
            a_off = 0;
          } else {
            a_off += LT_STAGE_BYTES >> 4;
          }
        }
        it_seg += yo1 - yo0 + 1;
      }
    }
    __syncwarp();
    // no asynchronous mbarrier arrive may still be in flight when the CTA exits: the last commit is a tfull that an
    // epilogue warp waits for before the final __syncthreads, and commits complete in order
  } else if (warp < 2 + FWD_EPI_WARPS) {
    // ===== epilogue: two groups of 8 warps alternate output tiles (group = staging buffer) =====
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;                                       // tile row = TMEM lane
    const int R7 = lane & 7;                                             // P and the segment origins are multiples of 8
    const int img_i = row / g.P, x1 = row - img_i * g.P;                 // image within the group, column + 1
    const bool valid_x = img_i < G && x1 >= 1 && x1 <= g.W;
    const uint32_t st1_0 = smem_u32(stage1 + row * SLAB_ROW_BYTES);
    const uint32_t st2_0 = smem_u32(stage2 + row * SLAB_ROW_BYTES);
    LineWalk w(g.H, t0, t1);
    int it = 0;
    while (w.next()) {
      const int n_img = w.j * G + img_i;
      const bool valid = valid_x && n_img < g.N;
      for (int yo = w.yo0; yo <= w.yo1; ++yo, ++it) {
        if ((it & 1) != grp) continue;
        const int a = it & 7;
        uint4 t2[4];
        if (mode_has_aux2(MODE)) {            // addend tile -> registers, buffer free for this group's next tile
          mbar_wait(&cs->aux2full[grp], (it >> 1) & 1);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) t2[jj] = lds_u4(st2_0 + grp * TILE_BYTES + (((half * 4 + jj) ^ R7) << 4));
          mbar_arrive(&cs->aux2free[grp]);
        }
        mbar_wait(&cs->tfull[a], (it >> 3) & 1);
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
        const int sb = it % NB;                                          // staging buffer of this tile
        const uint32_t st1 = st1_0 + sb * TILE_BYTES;
        if (mode_has_aux1(MODE)) mbar_wait(&cs->auxfull[sb], (it / NB) & 1);           // (the loader waited for the buffer)
        else mbar_wait(&cs->outfree[sb], ((it / NB) & 1) ^ 1);           // the store of tile it - NB has read the buffer
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
            unpack8(t2[jj], t);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (o[e] + t[e]) * elu_grad_from_out(sk[e]);
          }
          sts_u4(st1 + pos, valid ? pack8(o) : make_uint4(0, 0, 0, 0));
        }
        fence_proxy_async_smem();
        if ((warp == 2 || warp == 10) && lane == 0) trace_at(trace_buf, it, 4);
        mbar_arrive(&cs->outready[sb]);
      }
    }
  } else if (warp == 2 + FWD_EPI_WARPS) {
    // ===== aux1 loader: the saved-activation tile of output tile it, up to NB - 1 tiles ahead of its epilogue.  (One bulk
    // copy costs its issuing thread ~110 cycles -- 8 segments per tile at 14x14 are most of a tile period -- so the addend
    // stream of MODE_ADD_MUL has its own thread and buffers, below.) =====
    if (mode_has_aux1(MODE) && lane == 0) {
      const uint64_t pol_first = l2_policy_evict_first();
      asm volatile("griddepcontrol.wait;" ::: "memory");
      LineWalk w(g.H, t0, t1);
      int it = 0;
      while (w.next()) {
        const int geff = min(G, g.N - w.j * G);
        for (int yo = w.yo0; yo <= w.yo1; ++yo, ++it) {
          const int b = it % NB;
          if (it >= NB) mbar_wait(&cs->outfree[b], ((it / NB) - 1) & 1);   // the store of tile it - NB has read buffer b
          mbar_arrive_expect_tx(&cs->auxfull[b], static_cast<uint32_t>(geff) * px_bytes);
          size_t off = line_off(w.j * G, yo) + SLAB_ROW_BYTES;          // pixel rows 1 .. W of each segment
          uint32_t so = b * TILE_BYTES + SLAB_ROW_BYTES;
          // aux1 is a saved activation: after this launch it is not read again (backward) or not before the backward pass
          // (the skip tile of MODE_RES_ELU) -- it must not evict the slabs the next launches re-read from L2
          for (int i = 0; i < geff; ++i, off += img_bytes, so += seg_bytes)
            bulk_g2s_hint(stage1 + so, aux1 + off, px_bytes, &cs->auxfull[b], pol_first);
        }
      }
    }
  } else if (warp == 2 + FWD_EPI_WARPS + 3) {
    // ===== addend loader (MODE_ADD_MUL): tile it goes to the buffer of epilogue group it & 1 as soon as that group has
    // copied the addend of its previous tile (it - 2) to registers =====
    if (mode_has_aux2(MODE) && lane == 0) {
      const uint64_t pol_first = l2_policy_evict_first();
      asm volatile("griddepcontrol.wait;" ::: "memory");
      LineWalk w(g.H, t0, t1);
      int it = 0;
      while (w.next()) {
        const int geff = min(G, g.N - w.j * G);
        for (int yo = w.yo0; yo <= w.yo1; ++yo, ++it) {
          const int b = it & 1;
          if (it >= 2) mbar_wait(&cs->aux2free[b], ((it >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&cs->aux2full[b], static_cast<uint32_t>(geff) * px_bytes);
          size_t off = line_off(w.j * G, yo) + SLAB_ROW_BYTES;
          uint32_t so = b * TILE_BYTES + SLAB_ROW_BYTES;
          for (int i = 0; i < geff; ++i, off += img_bytes, so += seg_bytes)   // last reader of the addend slab
            bulk_g2s_hint(stage2 + so, aux2 + off, px_bytes, &cs->aux2full[b], pol_first);
        }
      }
    }
  } else if (warp == 2 + FWD_EPI_WARPS + 1) {
    // ===== output drain: G bulk stores per tile =====
    if (lane == 0) {
      const uint64_t pol_last = l2_policy_evict_last();
      LineWalk w(g.H, t0, t1);
      int it = 0;
      while (w.next()) {
        const int geff = min(G, g.N - w.j * G);
        for (int yo = w.yo0; yo <= w.yo1; ++yo, ++it) {
          const int b = it % NB;
          mbar_wait(&cs->outready[b], (it / NB) & 1);
          uint8_t* dst = out + line_off(w.j * G, yo) + SLAB_ROW_BYTES;   // pixel rows 1 .. W of each segment
          const uint8_t* src = stage1 + b * TILE_BYTES + SLAB_ROW_BYTES;
          for (int i = 0; i < geff; ++i, dst += img_bytes, src += seg_bytes)   // the next launch(es) re-read the output slab
            bulk_s2g_hint(dst, src, px_bytes, pol_last);
          bulk_commit();
          trace_at(trace_buf, it, 5);
          bulk_wait_read<0>();                // the store has read its staging buffer (a few hundred cycles; the next
          mbar_arrive(&cs->outfree[b]);       // tile's outready is a tile period away): free it for tile it + NB
        }
      }
      bulk_wait_all<0>();
    }
  } else if (warp == 2 + FWD_EPI_WARPS + 2) {
    // ===== gatekeeper: takes the barrier waits of the issuer (input line landed, fresh accumulator slots drained) and
    // opens ONE gate per input tile.  A gate of the ring is reused LT_GATES tiles later; by then its tile has been issued
    // and completed (this tile's input stage was freed by the commit of a tile at most LT_GATES - 1 back). =====
    if (lane == 0) {
      LineWalk w(g.H, t0, t1);
      int ii = 0, it_seg = 0;
      while (w.next()) {
        const int yo0 = w.yo0, yo1 = w.yo1;
        for (int yi = w.yi0(); yi <= w.yi1(); ++yi, ++ii) {
          const int b_lo = yi - 1 >= yo0 ? 0 : (yi >= yo0 ? 1 : 2);
          const int b_hi = yi + 1 <= yo1 ? 2 : (yi <= yo1 ? 1 : 0);
          const int bf = yi == 1 ? 1 : 2;
          const int it0 = it_seg + (yi - 1 - yo0);
          mbar_wait(&cs->full[ii % NS], (ii / NS) & 1);
          for (int b = max(bf, b_lo); b <= b_hi; ++b)
            mbar_wait(&cs->tempty[(it0 + b) & 7], (((it0 + b) >> 3) & 1) ^ 1);
          mbar_arrive(&cs->go[ii & (LT_GATES - 1)]);
        }
        it_seg += yo1 - yo0 + 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_at(trace_buf, 0, 12);                      // every role done (stores landed)
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64 * LT_SLOTS);
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

// ---- line-tile wgrad (default) --------------------------------------------------------------------------------------
// Same tile geometry as the line-tile forward kernel: a tile = one image line of G images = 128 contraction rows.
//   dW[dy][dx][ci][co] = sum over tiles  X_line(y + dy)[row + dx][ci] * G_line(y)[row][co]
// For the X line yi ONE N = 192 instruction pairs two horizontal views of it (M = 128: dx = -1, 0; second instruction:
// dx = +1 and the constant "ones" atom whose row 64 collects the bias gradient) with the THREE gradient lines yi-1, yi,
// yi+1 -- three 64-column atoms LBO = one ring slot apart -- so the accumulator block b is the vertical tap dy = 1 - b:
// 16 instructions of 96 tensor cycles per tile (10 KB of operands each: tensor-bound) instead of 40 of N = 64 at 49
// port-bound cycles, the zero padding lines are never touched, and every X / G line is loaded exactly once per CTA
// range.  Gradient lines live in a ring of WL_GSLOTS slots; where three consecutive lines wrap around the ring the
// instruction is split in two.  Accumulators (2 x 192 columns) are zeroed once and summed over all tiles of the CTA.
constexpr int WL_XSTAGES = 3;
constexpr int WL_GSLOTS = 6;
constexpr int WL_GATES = 4;
constexpr int WL_THREADS = 32 * 8;      // X producer, issuer, 4 epilogue warps (one per TMEM lane quarter), gatekeeper, G producer

struct WlineSmem {
  uint64_t xfull[WL_XSTAGES], xempty[WL_XSTAGES], gfull[WL_GSLOTS], gempty[WL_GSLOTS], go[WL_GATES], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WL_THREADS, 1)
conv3x3_wgrad_line_kernel(const uint8_t* __restrict__ x_base, const uint8_t* __restrict__ gr_base, float* __restrict__ dW_base,
                          float* __restrict__ db_base, SlabGeom g, size_t slab_stride, int ctas_per_layer, int G, int n_lines) {
  const int layer = blockIdx.x / ctas_per_layer;
  const int part = blockIdx.x % ctas_per_layer;
  const uint8_t* __restrict__ x = x_base + static_cast<size_t>(layer) * slab_stride;
  const uint8_t* __restrict__ gr = gr_base + static_cast<size_t>(layer) * slab_stride;
  float* __restrict__ dW = dW_base + static_cast<size_t>(layer) * 9 * 64 * 64;
  float* __restrict__ db = db_base + static_cast<size_t>(layer) * 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* xst = smem;                                                  // [WL_XSTAGES][guard | 128 rows | guard]
  uint8_t* gst = xst + WL_XSTAGES * LT_STAGE_BYTES;                     // [WL_GSLOTS][128 rows]
  uint8_t* ones = gst + WL_GSLOTS * TILE_BYTES;                         // AFTER every X stage (positive LBO)
  WlineSmem* ws = reinterpret_cast<WlineSmem*>(ones + ONES_BYTES);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform

  // everything no copy writes stays zero: guards, padding columns, rows past the last segment
  for (int i = threadIdx.x; i < (WL_XSTAGES * LT_STAGE_BYTES + WL_GSLOTS * TILE_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(xst)[i] = make_uint4(0, 0, 0, 0);
  // constant operand: logical channel 0 of every row = 1.0, everything else 0 (swizzled like a slab row)
  for (int i = threadIdx.x; i < ONES_BYTES / 16; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (c == (r & 7)) val.x = 0x00003f80u;  // bf16 1.0 in element 0
    reinterpret_cast<uint4*>(ones)[i] = val;
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < WL_XSTAGES; ++s) {
      mbar_init(&ws->xfull[s], 1);
      mbar_init(&ws->xempty[s], 1);
    }
    for (int s = 0; s < WL_GSLOTS; ++s) {
      mbar_init(&ws->gfull[s], 1);
      mbar_init(&ws->gempty[s], 1);
    }
    for (int s = 0; s < WL_GATES; ++s) mbar_init(&ws->go[s], 1);
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
  if (warp >= 2 && warp < 6) {                                           // zero the two 192-column accumulators
    const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = 0; c < 384; c += 8) tmem_st8_fill(tl + c, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // contiguous range of X line tiles per CTA (keeps the fp32 accumulation order fixed for a given grid)
  const int t0 = static_cast<int>(static_cast<long long>(part) * n_lines / ctas_per_layer);
  const int t1 = static_cast<int>(static_cast<long long>(part + 1) * n_lines / ctas_per_layer);
  const uint32_t seg_bytes = static_cast<uint32_t>(g.P) * SLAB_ROW_BYTES;
  const uint32_t in_bytes = static_cast<uint32_t>(min(g.P, g.W + 2)) * SLAB_ROW_BYTES;   // pixels + one padding row either side
  const uint32_t px_bytes = static_cast<uint32_t>(g.W) * SLAB_ROW_BYTES;
  const size_t img_bytes = static_cast<size_t>(g.IMG) * SLAB_ROW_BYTES;
  auto line_off = [&](int n, int line) {
    return (static_cast<size_t>(g.HALO) + static_cast<size_t>(n) * g.IMG + static_cast<size_t>(line) * g.P) * SLAB_ROW_BYTES;
  };

  if (warp == 0) {
    // ===== X producer: one X line of G images per stage =====
    if (lane == 0) {
      LineWalk w(g.H, t0, t1);
      int xi = 0;
      while (w.next()) {
        const int geff = min(G, g.N - w.j * G);
        for (int yi = w.yo0; yi <= w.yo1; ++yi, ++xi) {
          const int s = xi % WL_XSTAGES;
          mbar_wait(&ws->xempty[s], ((xi / WL_XSTAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&ws->xfull[s], static_cast<uint32_t>(G) * in_bytes);
          uint8_t* dst = xst + s * LT_STAGE_BYTES + LT_GUARD * SLAB_ROW_BYTES;
          const uint8_t* src = x + line_off(w.j * G, yi);
          for (int i = 0; i < geff; ++i, dst += seg_bytes, src += img_bytes) bulk_g2s(dst, src, in_bytes, &ws->xfull[s]);
          for (int i = geff; i < G; ++i, dst += seg_bytes) bulk_g2s(dst, x, in_bytes, &ws->xfull[s]);       // zeros (slab guard rows)
        }
      }
    }
  } else if (warp == 7) {
    // ===== G producer: the gradient lines of every segment, each once, into the slot ring =====
    if (lane == 0) {
      LineWalk w(g.H, t0, t1);
      int gc = 0;
      while (w.next()) {
        const int geff = min(G, g.N - w.j * G);
        const int gb = min(g.H, w.yo1 + 1);
        for (int y = max(1, w.yo0 - 1); y <= gb; ++y, ++gc) {
          const int sl = gc % WL_GSLOTS;
          mbar_wait(&ws->gempty[sl], ((gc / WL_GSLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&ws->gfull[sl], static_cast<uint32_t>(G) * px_bytes);
          uint8_t* dst = gst + sl * TILE_BYTES + SLAB_ROW_BYTES;          // pixel rows 1 .. W of each segment
          const uint8_t* src = gr + line_off(w.j * G, y) + SLAB_ROW_BYTES;
          for (int i = 0; i < geff; ++i, dst += seg_bytes, src += img_bytes) bulk_g2s(dst, src, px_bytes, &ws->gfull[sl]);
          for (int i = geff; i < G; ++i, dst += seg_bytes) bulk_g2s(dst, gr, px_bytes, &ws->gfull[sl]);     // zeros
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one wait per tile (the gatekeeper has seen the X line and its gradient lines land) =====
    if (elect_one()) {
      constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 1, 1);
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t GSLOT = TILE_BYTES >> 4;                          // one gradient-line slot in descriptor units
      const uint32_t x_lo0 = ((smem_u32(xst) & 0x3FFFFu) >> 4) + LT_GUARD * 8;   // tile row 0 of X stage 0
      const uint32_t g_lo0 = (smem_u32(gst) & 0x3FFFFu) >> 4;
      const uint32_t ones_lo = (smem_u32(ones) & 0x3FFFFu) >> 4;
      const uint32_t go0 = smem_u32(&ws->go[0]), xempty0 = smem_u32(&ws->xempty[0]), gempty0 = smem_u32(&ws->gempty[0]);
      auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      auto commit_addr = [](uint32_t bar) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      };
      LineWalk w(g.H, t0, t1);
      int xi = 0, gc_seg = 0, sidx = 0;
      uint32_t x_off = 0;
      while (w.next()) {
        const int ga = max(1, w.yo0 - 1), gb = min(g.H, w.yo1 + 1);
        for (int yi = w.yo0; yi <= w.yo1; ++yi, ++xi) {
          mbar_wait(&ws->go[xi & (WL_GATES - 1)], (xi / WL_GATES) & 1);
          tc_fence_after();
          // accumulator block b = gradient line yi - 1 + b = vertical tap dy = 1 - b
          const int b_lo = yi - 1 >= 1 ? 0 : 1, b_hi = yi + 1 <= g.H ? 2 : 1;
          const int gcnt = gc_seg + (yi - 1 + b_lo - ga);                  // ring counter of the first gradient line used
          const uint32_t slot0 = static_cast<uint32_t>(gcnt % WL_GSLOTS), nbt = b_hi - b_lo + 1;
          const uint32_t n0 = min(nbt, WL_GSLOTS - slot0), n1 = nbt - n0;  // lines before / after the ring wraps
          const uint32_t d1 = tmem + b_lo * 64, d2 = tmem + 192 + b_lo * 64;
          const uint32_t id0 = idesc64 + (n0 - 1) * (8u << 17), id1 = idesc64 + (n1 - 1) * (8u << 17);
          // A: two atoms LBO apart.  dx = -1, 0: rows -1 and 0 of the tile (LBO = one row = 8 units);  dx = +1 and "ones":
          // LBO = distance to the constant atom, shrinking by the 128 units the address advances per K step
          const uint32_t xa = x_lo0 + x_off;
          const uint32_t a1 = (xa - 8u) | (8u << 16);
          const uint32_t a2 = (xa + 8u) | ((ones_lo - (xa + 8u)) << 16);
          const uint64_t A1 = desc(a1), A2 = desc(a2);
          const uint64_t B0 = desc((g_lo0 + slot0 * GSLOT) | (GSLOT << 16)), B1 = desc(g_lo0 | (GSLOT << 16));
          // per K step (16 rows = 128 descriptor units): every address + 128 ks; the LBO of A2 - 128 ks (constant atom)
          if (n1 == 0) {
#pragma unroll
            for (int ks = 0; ks < TILE_M / 16; ++ks) {
              const uint64_t k = ks * 128u;
              umma_bf16(d1, A1 + k, B0 + k, id0, 1u);
              umma_bf16(d2, A2 + k - (k << 16), B0 + k, id0, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < TILE_M / 16; ++ks) {
              const uint64_t k = ks * 128u;
              umma_bf16(d1, A1 + k, B0 + k, id0, 1u);
              umma_bf16(d1 + n0 * 64, A1 + k, B1 + k, id1, 1u);
              umma_bf16(d2, A2 + k - (k << 16), B0 + k, id0, 1u);
              umma_bf16(d2 + n0 * 64, A2 + k - (k << 16), B1 + k, id1, 1u);
            }
          }
          commit_addr(xempty0 + 8u * sidx);
          // gradient line yi - 1 is not needed after this tile; the last tile of a segment frees the rest
          if (yi - 1 >= ga) commit_addr(gempty0 + 8u * ((gc_seg + (yi - 1 - ga)) % WL_GSLOTS));
          if (yi == w.yo1)
            for (int y = yi; y <= gb; ++y) commit_addr(gempty0 + 8u * ((gc_seg + (y - ga)) % WL_GSLOTS));
          if (++sidx == WL_XSTAGES) {
            sidx = 0;
            x_off = 0;
          } else {
            x_off += LT_STAGE_BYTES >> 4;
          }
        }
        gc_seg += gb - ga + 1;
      }
      // the issuer itself waits for the last commit: no asynchronous mbarrier arrive may still be in flight when the CTA
      // exits (it would land in the shared memory of whichever CTA gets this SM next)
      umma_commit(&ws->done);
      mbar_wait(&ws->done, 0);
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 6) {
    // ===== final epilogue: the accumulated gradient of this CTA's range -> global (fp32 atomics) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;                // accumulator row: view (m >> 6) x input channel (m & 63)
    mbar_wait(&ws->done, 0);
    tc_fence_after();
    if (t1 > t0) {
      const int ci = m & 63;
      for (int half = 0; half < 2; ++half) {    // 0: dx = -1 / 0 (rows 0..63 / 64..127);  1: dx = +1 (rows 0..63), bias (row 64)
        const int dx = half == 0 ? (m >> 6) - 1 : 1;
        for (int b = 0; b < 3; ++b) {
          const int tap = (2 - b) * 3 + (dx + 1);                          // dy = 1 - b
          for (int h32 = 0; h32 < 2; ++h32) {
            uint32_t v[32];
            tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + half * 192 + b * 64 + h32 * 32, v);
            tmem_ld_wait();
            float* dst = nullptr;
            if (half == 0 || m < 64) dst = dW + (static_cast<size_t>(tap) * 64 + ci) * 64 + h32 * 32;
            else if (m == 64 && b == 1) dst = db + h32 * 32;               // sum of the gradient line itself (dy = 0 block)
            if (dst) {
#pragma unroll
              for (int e = 0; e < 32; e += 4)
                red_add_v4(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                           __uint_as_float(v[e + 3]));
            }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ===== gatekeeper: waits for the X line and for the gradient lines it needs, opens one gate per tile =====
    if (lane == 0) {
      LineWalk w(g.H, t0, t1);
      int xi = 0, gc = 0;
      while (w.next()) {
        const int gb = min(g.H, w.yo1 + 1);
        int gnext = max(1, w.yo0 - 1);
        for (int yi = w.yo0; yi <= w.yo1; ++yi, ++xi) {
          for (; gnext <= min(gb, yi + 1); ++gnext, ++gc) mbar_wait(&ws->gfull[gc % WL_GSLOTS], (gc / WL_GSLOTS) & 1);
          mbar_wait(&ws->xfull[xi % WL_XSTAGES], (xi / WL_XSTAGES) & 1);
          mbar_arrive(&ws->go[xi & (WL_GATES - 1)]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static size_t wgrad_line_smem_bytes() {
  return 1024 + static_cast<size_t>(WL_XSTAGES) * LT_STAGE_BYTES + static_cast<size_t>(WL_GSLOTS) * TILE_BYTES + ONES_BYTES +
         sizeof(WlineSmem) + 64;
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
    *reinterpret_cast<uint4*>(wf + wpack_block(tap) * W_TAP_BYTES + rowi * 128 + ((j ^ (rowi & 7)) << 4)) = pack8(f);
    if (wb) {
      // dgrad operand: tap' = 8 - tap, [row = ci][k = co]
      float h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = v[(tap * 64 + rowi) * 64 + (j * 8 + e)] * scale[j * 8 + e];
      *reinterpret_cast<uint4*>(wb + wpack_block(8 - tap) * W_TAP_BYTES + rowi * 128 + ((j ^ (rowi & 7)) << 4)) = pack8(h);
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

static size_t line_fixed_smem(int mode) {
  return 1024 + W_BYTES + line_stage_tiles(mode) * TILE_BYTES + sizeof(LineSmem) + 64;
}

template <int MODE, int NS>
static int launch_conv_line_ns(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                               const void* a2, void* out, const SlabGeom& g, cudaStream_t st) {
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(conv3x3_line_kernel<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const int G = TILE_M / g.P;                                   // images per line tile
  const int n_groups = (g.N + G - 1) / G;
  const long long n_lines = static_cast<long long>(n_groups) * g.H;
  BRUNO_REQUIRE(n_lines < (1ll << 30), "bruno_conv3x3: too many line tiles");
  const int grid = n_lines < num_sms() ? static_cast<int>(n_lines) : num_sms();
  const size_t smem = line_fixed_smem(MODE) + static_cast<size_t>(NS) * LT_STAGE_BYTES;
  ProfSpan span(BRUNO_PROF_CONV, st);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, conv3x3_line_kernel<MODE, NS>, static_cast<const uint8_t*>(in), static_cast<const uint8_t*>(w), bias,
                     static_cast<const uint8_t*>(a1), static_cast<const uint8_t*>(a2), static_cast<uint8_t*>(out), g, G,
                     static_cast<int>(n_lines), bias_img_stride, g_conv_trace_host);
  return check_launch("conv3x3_line_kernel");
}

// 0 = automatic: the line-tile kernel when every CTA gets at least 8 output lines, else the per-tap kernel.  A CTA range
// that starts / ends inside an image pays two halo input lines, and the line kernel's set-up (736 threads, 8 TMEM slots,
// zeroed stages) is longer: measured cross-over (tools/bench_conv.py, both kernels, same build) 320x28x28 (15 lines per CTA)
// 18.7 vs 22.0 us, 160x28x28 / 640x14x14 (7.6) a tie in modes 0-2 and 14.6 vs 12.8 us in MODE_ADD_MUL, 320x14x14 (3.8)
// 10.0 vs 9.0 us, the 40-image few-shot episode 1.61 vs 1.33 ms per episode.
// 1 = per-tap kernel (also the fallback for a row pitch that is not a multiple of 8 or exceeds 128); 2 = line-tile kernel
static int g_conv_variant = 0;

template <int MODE>
static int launch_conv(const void* in, const void* w, const float* bias, int bias_img_stride, const void* a1,
                       const void* a2, void* out, const SlabGeom& g, cudaStream_t st) {
  const int G = g.P <= TILE_M ? TILE_M / g.P : 1;
  const long long n_lines = static_cast<long long>((g.N + G - 1) / G) * g.H;
  const bool line = g.P <= TILE_M && g.P % 8 == 0 &&
                    (g_conv_variant == 2 || (g_conv_variant == 0 && (n_lines >= 8ll * num_sms() || tap_stages(g, MODE) < 2)));
  if (line) return launch_conv_line_ns<MODE, line_ns(MODE)>(in, w, bias, bias_img_stride, a1, a2, out, g, st);
  return launch_conv_tap<MODE>(in, w, bias, bias_img_stride, a1, a2, out, g, st);
}

}  // namespace bruno

using namespace bruno;

extern "C" {

size_t bruno_slab_bytes(int N, int H, int W) {
  const SlabGeom g = slab_geom(N, H, W);
  return static_cast<size_t>(g.rows) * SLAB_ROW_BYTES;
}

size_t bruno_conv_wpack_bytes(void) { return W_BYTES; }

#ifndef BRUNO_NO_DEBUG
int bruno_debug_conv_trace(long long* device_buffer) {
  g_conv_trace_host = device_buffer;
  return BRUNO_OK;
}
int bruno_debug_conv_variant(int variant) {
  BRUNO_REQUIRE(variant >= 0 && variant <= 2, "bruno_debug_conv_variant: 0 = automatic, 1 = per-tap kernel, 2 = line-tile kernel");
  g_conv_variant = variant;
  return BRUNO_OK;
}
#endif

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
  if (g_conv_variant != 1 && g.P <= TILE_M && g.P % 8 == 0) {   // line-tile kernel (the per-tap one: odd pitches, wide images, debug)
    static OncePerDevice once_line;
    once_line.run([&] {
      cudaFuncSetAttribute(conv3x3_wgrad_line_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    const int G = TILE_M / g.P;
    const long long n_lines = static_cast<long long>((N + G - 1) / G) * H;
    BRUNO_REQUIRE(n_lines < (1ll << 30), "bruno_conv3x3_wgrad: too many line tiles");
    int per_layer = num_sms() / L;
    if (per_layer < 1) per_layer = 1;
    if (per_layer > n_lines) per_layer = static_cast<int>(n_lines);
    ProfSpan span(BRUNO_PROF_WGRAD, static_cast<cudaStream_t>(stream));
    conv3x3_wgrad_line_kernel<<<per_layer * L, WL_THREADS, wgrad_line_smem_bytes(), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(x_slab), static_cast<const uint8_t*>(g_slab), dW, db, g,
        static_cast<size_t>(g.rows) * SLAB_ROW_BYTES, per_layer, G, static_cast<int>(n_lines));
    return check_launch("conv3x3_wgrad_line_kernel");
  }
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
