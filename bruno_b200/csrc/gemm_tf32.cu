// fp32-accurate GEMM on the tensor cores for the dense coupling layers (CouplingLayerDense, nn_extra_nvp.py:190-274;
// dense_wn :408-432): C[M,N] = op(A)[M,K] op(B)[K,N] with the fused prologue / epilogue of GemmEpi.
//
// The dense s/t networks are the parity-critical fp32 part of the flow (en2_noconv_mnist is dense only, tolerance 1e-4),
// but their GEMMs are skinny (640 x 512 x 784 in training, 40 x 512 x 784 when scoring a few-shot episode): on CUDA cores
// they are bound by FFMA issue and shared-memory traffic at ~19 % of the fp32 peak.  Here every operand is split on the
// fly into two TF32 terms, a = hi + lo (hi = tf32(a), lo = tf32(a - hi), |a - hi - lo| <= 2^-22 |a|), and the product is
// accumulated as  hi*hi + lo*hi + hi*lo  by three tcgen05.mma.kind::tf32 per K step into one fp32 TMEM accumulator
// (the dropped lo*lo term is 2^-22 relative: the result is as accurate as an fp32 FFMA chain of length K).
//
//   * no transposes: a row-major [M,K] / [N,K] source is a K-major UMMA operand, a row-major [K,M] / [K,N] source an
//     MN-major one; both are written by the loader warps straight into the 128-byte-swizzled canonical layouts;
//   * 128 x 128 output tile per CTA, K tiles of 32 through a 3-stage ring (4 loader warps -> 1 issuer warp);
//   * split-K over gridDim.z so that ~one CTA per SM is busy even for 8..20 output tiles; partials go to a workspace
//     and the last CTA of a tile (ticket) adds them in z order (deterministic) and runs the epilogue.
#include "dense_gemm.cuh"
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "ptx.cuh"

namespace bruno {

namespace {

constexpr int TG_BM = 128, TG_BK = 32, TG_STAGES = 3;
constexpr int TG_TILE_BYTES = 128 * TG_BK * 4;            // 16 KB: one 128-row operand term of one stage
constexpr int TG_LOADER_WARPS = 4 * TG_STAGES;            // one group of 4 warps per ring stage: three K tiles in flight
constexpr int TG_THREADS = 32 * (TG_LOADER_WARPS + 1);    // + issuer warp; warps 0..3 also run the epilogue
constexpr int TG_MAX_SPLIT = 8;

struct TgSmem {
  uint64_t full[TG_STAGES], empty[TG_STAGES], done;
  uint64_t red;       // cluster split-K: counts the bytes of partial rows the peers (and this CTA) push into this CTA
  uint32_t tmem_base;
  int last;
};

// a = hi + lo with hi = a rounded to TF32 (11 significant bits) and lo = a - hi (exact, <= 13 significant bits).
// cvt.rna.tf32.f32 issues at a fraction of the ALU rate and made the loaders the bottleneck; the same rounding is two
// integer ops: add half an ulp of the TF32 mantissa to the magnitude bits, clear the 13 low bits.  lo is stored as it is:
// the tensor core ignores its 13 low mantissa bits, |error| <= 2^-10 |lo| <= 2^-21 |a|.
__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  uint4 h, l;
  h.x = (__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u;
  h.y = (__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u;
  h.z = (__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u;
  h.w = (__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u;
  l.x = __float_as_uint(v.x - __uint_as_float(h.x));
  l.y = __float_as_uint(v.y - __uint_as_float(h.y));
  l.z = __float_as_uint(v.z - __uint_as_float(h.z));
  l.w = __float_as_uint(v.w - __uint_as_float(h.w));
  sts_u4(hi_addr, h);
  sts_u4(lo_addr, l);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::tf32 instruction descriptor: TF32 A/B (format 2), fp32 D, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// One 128 x 32 operand tile (rows = M or N index r0.., K tile at k0) of a row-major source, held in registers between the
// global loads (fetch) and the split into two TF32 terms written in the canonical swizzled layout (store).  128 loader
// threads, 8 x float4 each.
//   KMAJOR  (source [rows, K], K contiguous): smem row r (128 B = 32 k), 16-byte chunk c at (c ^ (r & 7)); a warp
//            instruction covers 4 full 128-byte source lines.
//   MNMAJOR (source [K, rows], rows contiguous): 32-bit MN-major operands use the "128-byte swizzle with 32-byte
//            atomicity" (descriptor layout type 1, cute Layout_MN_SW128_32B_Atom: Swizzle<2,5,2>): four 32-row atom
//            columns 4 KB apart (LBO); inside one, smem row k (128 B = 32 consecutive rows), atoms of 4 k rows
//            (SBO = 512 B), 32-BYTE chunk cc at (cc ^ (k & 3)); a warp instruction covers 512 contiguous source bytes.
// Rows beyond `rows_total` are neither loaded nor stored: accumulator row m (column n) depends only on A row m (B row
// n), and those outputs are never read.  The K tail IS zero-filled (it feeds every output).
template <bool MNMAJOR, int ROWS>
struct OperandTile {
  float4 v[8];
  const float* p0;   // this thread's element 0 of the K tile at k = 0
  size_t istride;    // floats from element i to element i + 1
  uint32_t base;     // shared-memory offset of element 0 before the per-element swizzle term
  uint32_t mask;     // bit i: element i is inside the operand (loaded and stored)
  int kthr;          // K index of this thread's data inside a tile (element 0)

  // Everything that does not depend on the K tile is computed once per thread (the loaders are instruction-latency
  // bound: four warps work on a tile and nothing else hides their dependent chains).
  __device__ __forceinline__ void init(const float* __restrict__ src, int ld, int rows_total, int r0, int tid) {
    const int w = tid >> 5, l = tid & 31, c = l & 7;
    mask = 0;
    if (!MNMAJOR) {
      const int rb = w * 32 + (l >> 3);                       // element i: smem row rb + 4 i
      p0 = src + static_cast<size_t>(r0 + rb) * ld + 4 * c;
      istride = static_cast<size_t>(4) * ld;
      base = static_cast<uint32_t>(rb * 128 + ((c ^ (l >> 3)) << 4));   // (rb + 4 i) & 7 = (l >> 3) ^ 4 (i & 1)
      kthr = 4 * c;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (rb + 4 * i < ROWS && r0 + rb + 4 * i < rows_total) mask |= 1u << i;
    } else {
      const int a = l >> 3;
      p0 = src + static_cast<size_t>(w * 8) * ld + r0 + a * 32 + 4 * c;
      istride = static_cast<size_t>(ld);
      base = static_cast<uint32_t>(a * 4096 + w * 8 * 128 + (c << 4));    // k & 3 = i & 3 flips chunk bits 1..2
      kthr = w * 8;
      if (a * 32 < ROWS && r0 + a * 32 + 4 * c < rows_total) mask = 0xFFu;
    }
  }

  __device__ __forceinline__ uint32_t offset(int i) const {
    if (!MNMAJOR) return (base ^ (static_cast<uint32_t>(i & 1) << 6)) + static_cast<uint32_t>(i) * 512u;
    return (base ^ (static_cast<uint32_t>(i & 3) << 5)) + static_cast<uint32_t>(i) * 128u;
  }

  __device__ __forceinline__ void fetch(int ld, int K, int k0, const float* __restrict__ kscale) {
    if (!MNMAJOR) {
      const float* p = p0 + k0;
      const bool kin = k0 + kthr < K;
      if (mask == 0xFFu && kin) {                             // full tile: eight independent loads, no predicates
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(p + i * istride));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (((mask >> i) & 1u) && kin) v[i] = __ldg(reinterpret_cast<const float4*>(p + i * istride));
        }
      }
      if (kscale && kin) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(kscale + k0 + kthr));
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i].x *= s.x; v[i].y *= s.y; v[i].z *= s.z; v[i].w *= s.w; }
      }
    } else {
      const float* p = p0 + static_cast<size_t>(k0) * ld;
      if (mask == 0xFFu && k0 + TG_BK <= K) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(p + i * istride));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mask && k0 + kthr + i < K) v[i] = __ldg(reinterpret_cast<const float4*>(p + i * istride));
        }
      }
      if (kscale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int gk = k0 + kthr + i;
          const float s = gk < K ? __ldg(kscale + gk) : 0.f;
          v[i].x *= s; v[i].y *= s; v[i].z *= s; v[i].w *= s;
        }
      }
    }
  }

  __device__ __forceinline__ void store(uint32_t hi_base, uint32_t lo_base) const {
    if (mask == 0xFFu) {                                      // straight-line: the eight splits interleave
#pragma unroll
      for (int i = 0; i < 8; ++i) split_store(hi_base + offset(i), lo_base + offset(i), v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mask & (1u << i)) split_store(hi_base + offset(i), lo_base + offset(i), v[i]);
    }
  }
};

// debug timeline (bruno_debug_gemm_trace): CTA 0 records clock64() per K tile [64][8]:
// 0 stage free (loader), 1 tile stored, 2 arrived, 3 issuer saw it, 4 MMAs issued; row 63: 5 accumulator done, 6 epilogue end
static long long* g_gemm_trace_host = nullptr;
__device__ __forceinline__ void gtrace(long long* buf, int it, int slot) {
  if (buf != nullptr && blockIdx.x + blockIdx.y + blockIdx.z == 0 && it < 64) buf[it * 8 + slot] = clock64();
}

template <int ACT>
__device__ __forceinline__ float act_apply(float x) {
  if (ACT == 1) return x > 0.f ? x : expm1f(x);
  if (ACT == 2) return x > 0.f ? x : 0.2f * x;      // tf.nn.leaky_relu default alpha
  if (ACT == 3) return tanhf(x);
  return x;
}
// runs f(integral_constant<ACT>) for the runtime activation code: the activation is compile-time inside the row loops
// (left as a runtime test there, both transcendental paths are evaluated for every element: ~250 cycles per row)
template <class F>
__device__ __forceinline__ void with_act(int act, F f) {
  if (act == 1) f(std::integral_constant<int, 1>{});
  else if (act == 2) f(std::integral_constant<int, 2>{});
  else if (act == 3) f(std::integral_constant<int, 3>{});
  else f(std::integral_constant<int, 0>{});
}

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(TG_THREADS, 1)
tf32x3_gemm_kernel(int M, int N, int K, const float* __restrict__ A_, int lda, const float* __restrict__ B_, int ldb,
                   float* __restrict__ C_, int ldc, GemmEpi ep, int kchunk, float* __restrict__ ws, unsigned* __restrict__ tickets,
                   int cluster_reduce, GemmPair pair, int nx1, long long* trace_buf) {
  constexpr int B_TILE_BYTES = BN * TG_BK * 4;
  constexpr int STAGE_BYTES = 2 * TG_TILE_BYTES + 2 * B_TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
  extern __shared__ uint8_t tg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  TgSmem* cs = reinterpret_cast<TgSmem*>(smem + TG_STAGES * STAGE_BYTES);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // pair.mode 1: the upper half of gridDim.x is the second problem; mode 2: the upper half of the split-K slices multiplies
  // the second operand pair into the same tile (GemmPair, dense_gemm.cuh)
  const float* __restrict__ A = A_;
  const float* __restrict__ B = B_;
  float* __restrict__ C = C_;
  int bx = blockIdx.x, bz = blockIdx.z;
  if (pair.mode == 1 && bx >= nx1) {
    bx -= nx1;
    A = pair.A2;
    B = pair.B2;
    C = pair.C2;
    ep.bias = pair.bias2;
  } else if (pair.mode == 2 && bz >= static_cast<int>(gridDim.z >> 1)) {
    bz -= static_cast<int>(gridDim.z >> 1);
    A = pair.A2;
    B = pair.B2;
  }
  const int m0 = blockIdx.y * TG_BM, n0 = bx * BN;
  const int nk = (K + TG_BK - 1) / TG_BK;
  const int kt0 = bz * kchunk, kt1 = min(nk, kt0 + kchunk);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(&cs->full[s], 128);
      mbar_init(&cs->empty[s], 1);
    }
    mbar_init(&cs->done, 1);
    mbar_init(&cs->red, 1);
    fence_mbar_init();
  }
  if (warp == TG_LOADER_WARPS) {
    __syncwarp();
    tmem_alloc(&cs->tmem_base, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();      // the next launch may set itself up; its griddepcontrol.wait sees this grid complete
  if (gridDim.z > 1 && cluster_reduce) {
    // every CTA of the cluster runs and has initialised its barriers before anyone pushes into it.  Here, where the
    // peers start together and nothing is in flight: the acquire side of a cluster barrier invalidates L1 (CCTL.IVALL).
    if (threadIdx.x == 0) {                                   // bytes of partial rows this CTA will receive (see the epilogue)
      const int Sz = gridDim.z, rws = min(TG_BM, M - m0), per = (rws + Sz - 1) / Sz, rb = static_cast<int>(cluster_ctarank()) * per;
      const int nmine = max(0, min(rws, rb + per) - rb);
      if (nmine > 0) mbar_arrive_expect_tx(&cs->red, static_cast<uint32_t>(Sz * nmine * BN * 4));
    }
    __syncwarp();
    cluster_arrive_relaxed();
    cluster_wait();
  }
  const uint32_t tmem = cs->tmem_base;
  const uint32_t stage0 = smem_u32(smem);

  if (warp < TG_LOADER_WARPS) {
    // ===== loaders: group g owns ring stage g and K tiles kt0 + g, kt0 + g + 3, ... =====
    const int grp = warp >> 2;
    const int tid = threadIdx.x & 127;
    const uint32_t st = stage0 + grp * STAGE_BYTES;
    // A is stored [K, M] when transposed (M contiguous -> MN-major), B is stored [K, N] when NOT transposed.
    // The loads of the group's NEXT tile are issued right after the current one is stored, so they are in flight while
    // the tensor core works on this stage and while the group waits for it to drain.
    OperandTile<TA, TG_BM> ta_;
    OperandTile<!TB, BN> tb_;
    ta_.init(A, lda, M, m0, tid);
    tb_.init(B, ldb, N, n0, tid);
    int kt = kt0 + grp;
    griddep_wait();                                           // first global read of this thread
    if (kt < kt1) {
      ta_.fetch(lda, K, kt * TG_BK, ep.a_kscale);
      tb_.fetch(ldb, K, kt * TG_BK, nullptr);
    }
    uint32_t j = 0;
    for (; kt < kt1; kt += TG_STAGES, ++j) {
      mbar_wait(&cs->empty[grp], (j & 1u) ^ 1u);
      if (tid == 0) gtrace(trace_buf, kt - kt0, 0);
      ta_.store(st, st + TG_TILE_BYTES);
      tb_.store(st + 2 * TG_TILE_BYTES, st + 2 * TG_TILE_BYTES + B_TILE_BYTES);
      if (tid == 0) gtrace(trace_buf, kt - kt0, 1);
      fence_proxy_async_smem();
      mbar_arrive(&cs->full[grp]);
      if (tid == 0) gtrace(trace_buf, kt - kt0, 2);
      if (kt + TG_STAGES < kt1) {
        ta_.fetch(lda, K, (kt + TG_STAGES) * TG_BK, ep.a_kscale);
        tb_.fetch(ldb, K, (kt + TG_STAGES) * TG_BK, nullptr);
      }
    }
  } else {
    // ===== issuer: uniform loop, one elected lane issues =====
    constexpr uint32_t idesc = idesc_tf32(BN, TA ? 1 : 0, TB ? 0 : 1);
    constexpr uint32_t HI_K = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO 1024 B, version, SWIZZLE_128B
    constexpr uint32_t HI_MN = (512u >> 4) | (1u << 14) | (1u << 29);        // SBO 512 B, version, SWIZZLE_128B_BASE32B
    constexpr uint32_t A_HI = TA ? HI_MN : HI_K, B_HI = TB ? HI_K : HI_MN;
    constexpr uint32_t A_LBO = TA ? (4096u >> 4) : 1u, B_LBO = TB ? 1u : (4096u >> 4);
    constexpr uint32_t A_STEP = TA ? 64u : 2u, B_STEP = TB ? 2u : 64u;       // per K = 8: 8 smem rows / 32 bytes
    const bool leader = elect_one();
    int it = 0;
    for (int kt = kt0; kt < kt1; ++kt, ++it) {
      const int s = it % TG_STAGES;
      if (leader) {
        mbar_wait(&cs->full[s], (it / TG_STAGES) & 1);
        tc_fence_after();
        gtrace(trace_buf, it, 3);
        const uint32_t st = (stage0 + s * STAGE_BYTES) >> 4;
        const uint32_t ahi = (st & 0x3FFFu) | (A_LBO << 16), alo = ((st + (TG_TILE_BYTES >> 4)) & 0x3FFFu) | (A_LBO << 16);
        const uint32_t bhi = ((st + 2 * (TG_TILE_BYTES >> 4)) & 0x3FFFu) | (B_LBO << 16);
        const uint32_t blo = ((st + 2 * (TG_TILE_BYTES >> 4) + (B_TILE_BYTES >> 4)) & 0x3FFFu) | (B_LBO << 16);
#pragma unroll
        for (int kk = 0; kk < TG_BK / 8; ++kk) {
          const uint64_t dah = (static_cast<uint64_t>(A_HI) << 32) | (ahi + kk * A_STEP);
          const uint64_t dal = (static_cast<uint64_t>(A_HI) << 32) | (alo + kk * A_STEP);
          const uint64_t dbh = (static_cast<uint64_t>(B_HI) << 32) | (bhi + kk * B_STEP);
          const uint64_t dbl = (static_cast<uint64_t>(B_HI) << 32) | (blo + kk * B_STEP);
          umma_tf32(tmem, dal, dbh, idesc, (it | kk) != 0);     // small terms first
          umma_tf32(tmem, dah, dbl, idesc, 1u);
          umma_tf32(tmem, dah, dbh, idesc, 1u);
        }
        umma_commit(&cs->empty[s]);
        gtrace(trace_buf, it, 4);
      }
    }
    if (leader) {
      umma_commit(&cs->done);
      mbar_wait(&cs->done, 0);           // no asynchronous arrive may outlive the CTA
    }
    __syncwarp();
  }

  // ===== epilogue.
  // Split-K: the S CTAs of a tile are ONE THREAD-BLOCK CLUSTER (1, 1, S).  CTA z owns rows [z rows_per, (z + 1) rows_per) of
  // the tile.  Warps 0..3 of every CTA read their TMEM lane quarter (lane = row) and PUSH each row of their partial into
  // its owner's shared memory (st.async.shared::cluster, slot [source z][row]: fire and forget, no exposed round trip),
  // counting the bytes on the OWNER's mbarrier (complete_tx).  When its count is complete a CTA adds the S partials of
  // its rows in z order (deterministic) from LOCAL shared memory, applies the epilogue and writes C.  No cluster barrier
  // in the epilogue: its release side is a GPU-scope MEMBAR, its acquire side an L1 invalidate per warp (measured: 1100
  // cycles for the barrier and a clogged load/store pipe behind it).  No partials in global memory, no fence / ticket round trip, every CTA shares the
  // reduction.  (Round 1 wrote the partials to a workspace and let the last CTA of a tile -- ticket -- add them: half of
  // the 15 us launch; pulling the peers' tiles with ld.shared::cluster cost ~2000 cycles per batch of loads.  The
  // workspace path stays for a launch without cluster.)  The push buffer is outside the operand ring: a peer may
  // still be multiplying when the first rows arrive.
  const int S = gridDim.z;
  const bool loader_warp = warp < TG_LOADER_WARPS;
  constexpr int NCH = BN / 32;
  const int rows = min(TG_BM, M - m0);
  if (S > 1 && cluster_reduce) {
    if constexpr (BN == 64) {
      constexpr int PP = BN + 4;                              // 16-byte aligned rows; 16-byte stores at this pitch are conflict-free
      float* P = reinterpret_cast<float*>(smem + TG_STAGES * STAGE_BYTES + 128);     // [S][rows_per][PP]
      const int rank = static_cast<int>(cluster_ctarank());   // = blockIdx.z: cluster (1, 1, S), gridDim.z = S
      const int rows_per = (rows + S - 1) / S;
      static_assert(TG_LOADER_WARPS % NCH == 0, "a warp keeps its column chunk over the reduction tasks");
      const int rb = rank * rows_per;
      const int nmine = max(0, min(rows, rb + rows_per) - rb);          // rows this CTA owns
      if (warp < 4) {
        mbar_wait(&cs->done, 0);
        tc_fence_after();
        if (threadIdx.x == 0) gtrace(trace_buf, 63, 5);
        const int r = warp * 32 + lane;
        const int owner = r / rows_per, lr = r - owner * rows_per;
        const bool push = r < rows;                           // owner < S then
        const uint32_t dst = mapa_shared(smem_u32(P), static_cast<uint32_t>(push ? owner : rank)) +
                             static_cast<uint32_t>(((rank * rows_per + lr) * PP) * 4);
        const uint32_t dbar = mapa_shared(smem_u32(&cs->red), static_cast<uint32_t>(push ? owner : rank));
        uint32_t v[NCH][32];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v[ch]);
        tmem_ld_wait();
        if (threadIdx.x == 0) gtrace(trace_buf, 62, 6);
        if (push) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              st_async_v4(dst + static_cast<uint32_t>((ch * 32 + e) * 4), v[ch][e], v[ch][e + 1], v[ch][e + 2], v[ch][e + 3], dbar);
        }
        if (threadIdx.x == 0) gtrace(trace_buf, 60, 5);
      }
      if (loader_warp && nmine > 0) {
        // This code runs once per CTA: its cost is instruction fetch and dependent-issue latency as much as memory
        // (an unrolled 4-task scalar version took ~2300 cycles per iteration on LOCAL shared memory), so the loops
        // are short and not unrolled.  Fast path: a lane owns 4 consecutive columns (16 lanes per row, two rows per
        // warp and iteration, LDS.128 / STG.128); column scale / bias are fetched before the wait.
        const int zstride = rows_per * PP;
        const bool vec = (ldc & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                         (ep.pre_out == nullptr || (reinterpret_cast<uintptr_t>(ep.pre_out) & 15) == 0);
        const int half = lane >> 4, c4 = (lane & 15) * 4;
        const bool col4_on = vec && n0 + c4 < N;
        float cs4[4], cb4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          cs4[e] = (col4_on && ep.col_scale) ? ep.col_scale[n0 + c4 + e] : 1.f;
          cb4[e] = (col4_on && ep.bias) ? ep.bias[n0 + c4 + e] : 0.f;
        }
        mbar_wait(&cs->red, 0);                               // every partial row of this CTA's rows has landed
        if (threadIdx.x == 0) gtrace(trace_buf, 61, 5);
        with_act(ep.act, [&](auto act) {
          if (vec) {
            const uint32_t psm = smem_u32(P) + static_cast<uint32_t>(c4 * 4);
#pragma unroll 1
            for (int task = warp; 2 * task < nmine; task += TG_LOADER_WARPS) {
              const int lr = 2 * task + half;
              if (col4_on && lr < nmine) {
                float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int z = 0; z < TG_MAX_SPLIT; ++z)
                  if (z < S) {
                    const uint4 q = lds_u4(psm + static_cast<uint32_t>((z * zstride + lr * PP) * 4));
                    x[0] += __uint_as_float(q.x); x[1] += __uint_as_float(q.y);
                    x[2] += __uint_as_float(q.z); x[3] += __uint_as_float(q.w);
                  }
                const size_t o = static_cast<size_t>(m0 + rb + lr) * ldc + n0 + c4;
                float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ep.accumulate) old = *reinterpret_cast<const float4*>(C + o);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (ep.col_scale) x[e] = __fmul_rn(x[e], cs4[e]);
                  if (ep.bias) x[e] = __fadd_rn(x[e], cb4[e]);
                }
                if (ep.pre_out) *reinterpret_cast<float4*>(ep.pre_out + o) = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = act_apply<decltype(act)::value>(x[e]);
                if (ep.accumulate) { x[0] += old.x; x[1] += old.y; x[2] += old.z; x[3] += old.w; }
                *reinterpret_cast<float4*>(C + o) = make_float4(x[0], x[1], x[2], x[3]);
              }
              if (threadIdx.x == 0 && task == warp) gtrace(trace_buf, 61, 6);
            }
          } else {                                            // any N / ldc / alignment: one (row, 32-column chunk) per warp
            const int ntask = nmine * NCH;
            const int n = n0 + (warp % NCH) * 32 + lane;      // 12 % NCH == 0: a warp keeps its chunk
#pragma unroll 1
            for (int task = warp; task < ntask; task += TG_LOADER_WARPS) {
              if (n < N) {
                const float* src = P + (task / NCH) * PP + (task % NCH) * 32 + lane;
                float x = 0.f;
#pragma unroll
                for (int z = 0; z < TG_MAX_SPLIT; ++z)
                  if (z < S) x += src[z * zstride];
                const size_t o = static_cast<size_t>(m0 + rb + task / NCH) * ldc + n;
                if (ep.col_scale) x = __fmul_rn(x, ep.col_scale[n]);
                if (ep.bias) x = __fadd_rn(x, ep.bias[n]);
                if (ep.pre_out) ep.pre_out[o] = x;
                x = act_apply<decltype(act)::value>(x);
                if (ep.accumulate) x += C[o];
                C[o] = x;
              }
            }
          }
        });
        if (threadIdx.x == 0) gtrace(trace_buf, 60, 6);
      }
      if (threadIdx.x == 0) gtrace(trace_buf, 63, 6);
    }
  } else if (loader_warp) {
    // Un-split, or split-K through the workspace: warps 0..3 copy their TMEM lane quarter into a row-major shared tile
    // (the ring is idle once `done` fired; pitch BN + 1: conflict-free both ways), then all 12 warps handle rows (a warp
    // instruction = 32 consecutive columns of one row); with split-K the tile goes to the workspace and the LAST CTA
    // of the tile (ticket) adds the partials in z order.
    constexpr int TP = BN + 1;
    float* T = reinterpret_cast<float*>(smem);
    if (warp < 4) {
      mbar_wait(&cs->done, 0);
      tc_fence_after();
      if (threadIdx.x == 0) gtrace(trace_buf, 63, 5);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) T[(warp * 32 + lane) * TP + ch * 32 + e] = __uint_as_float(v[e]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * TG_LOADER_WARPS) : "memory");
    if (threadIdx.x == 0) gtrace(trace_buf, 60, 5);
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    float* wt = ws + static_cast<size_t>(tile) * S * (TG_BM * BN);        // [z][row][BN]
    // phase B: all 12 warps write rows (a warp instruction = 32 consecutive columns of one row)
    float* mine = wt + static_cast<size_t>(blockIdx.z) * (TG_BM * BN);
    if (S > 1) {
      for (int task = warp; task < rows * NCH; task += 2 * TG_LOADER_WARPS) {
        const int t2 = task + TG_LOADER_WARPS;
        const int r_a = task / NCH, c_a = task % NCH, r_b = t2 / NCH, c_b = t2 % NCH;
        const float va = T[r_a * TP + c_a * 32 + lane];
        const float vb = t2 < rows * NCH ? T[r_b * TP + c_b * 32 + lane] : 0.f;
        __stcg(mine + r_a * BN + c_a * 32 + lane, va);
        if (t2 < rows * NCH) __stcg(mine + r_b * BN + c_b * 32 + lane, vb);
      }
    } else {
      with_act(ep.act, [&](auto act) {
        constexpr int TU = 4;
        const int ntask = rows * NCH;
        for (int task0 = warp; task0 < ntask; task0 += TU * TG_LOADER_WARPS) {
          float val[TU], old[TU];
#pragma unroll
          for (int u = 0; u < TU; ++u) {
            const int task = task0 + u * TG_LOADER_WARPS;
            const int row = task / NCH, ch = task % NCH;
            const bool on = task < ntask && n0 + ch * 32 + lane < N;
            val[u] = on ? T[row * TP + ch * 32 + lane] : 0.f;
            old[u] = (on && ep.accumulate) ? C[static_cast<size_t>(m0 + row) * ldc + n0 + ch * 32 + lane] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < TU; ++u) {
            const int task = task0 + u * TG_LOADER_WARPS;
            const int row = task / NCH, ch = task % NCH;
            const int n = n0 + ch * 32 + lane;
            if (task < ntask && n < N) {
              const size_t o = static_cast<size_t>(m0 + row) * ldc + n;
              float x = val[u];
              if (ep.col_scale) x = __fmul_rn(x, ep.col_scale[n]);
              if (ep.bias) x = __fadd_rn(x, ep.bias[n]);
              if (ep.pre_out) ep.pre_out[o] = x;
              C[o] = act_apply<decltype(act)::value>(x) + old[u];
            }
          }
        }
      });
    }
    if (threadIdx.x == 0) gtrace(trace_buf, 60, 6);
    if (S > 1) {
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TG_LOADER_WARPS) : "memory");
      if (threadIdx.x == 0) cs->last = (atomicAdd(tickets + tile, 1u) == static_cast<unsigned>(S - 1));
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TG_LOADER_WARPS) : "memory");
      if (cs->last) {
        __threadfence();
        // task = (row, 32-column chunk); each lane adds the S partials of its column in z order.  Four tasks per
        // iteration: all their partial loads are in flight together (one L2 round trip per iteration, not per task).
        with_act(ep.act, [&](auto act) {
          constexpr int TU = 4;
          const int ntask = rows * NCH;
          for (int task0 = warp; task0 < ntask; task0 += TU * TG_LOADER_WARPS) {
            float part[TU][TG_MAX_SPLIT];
#pragma unroll
            for (int u = 0; u < TU; ++u) {
              const int task = task0 + u * TG_LOADER_WARPS;
              const float* src = wt + (task / NCH) * BN + (task % NCH) * 32 + lane;
#pragma unroll
              for (int z = 0; z < TG_MAX_SPLIT; ++z)
                part[u][z] = (task < ntask && z < S) ? __ldcg(src + static_cast<size_t>(z) * (TG_BM * BN)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < TU; ++u) {
              const int task = task0 + u * TG_LOADER_WARPS;
              if (task >= ntask) break;
              const int row = task / NCH, ch = task % NCH;
              const int n = n0 + ch * 32 + lane;
              float x = 0.f;
#pragma unroll
              for (int z = 0; z < TG_MAX_SPLIT; ++z) x += part[u][z];
              if (n < N) {
                const size_t o = static_cast<size_t>(m0 + row) * ldc + n;
                if (ep.col_scale) x = __fmul_rn(x, ep.col_scale[n]);
                if (ep.bias) x = __fadd_rn(x, ep.bias[n]);
                if (ep.pre_out) ep.pre_out[o] = x;
                x = act_apply<decltype(act)::value>(x);
                if (ep.accumulate) x += C[o];
                C[o] = x;
              }
            }
          }
        });
        if (threadIdx.x == 0) tickets[tile] = 0u;
      }
    }
    if (threadIdx.x == 0) gtrace(trace_buf, 63, 6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) gtrace(trace_buf, 62, 5);
  if (warp == TG_LOADER_WARPS) {
    __syncwarp();
    tmem_dealloc(tmem, BN);
  }
}

// ---- launch geometry (pure host arithmetic: also exposed as bruno_debug_gemm_plan and tested without a GPU) ----
// 64-column tiles while 128-column tiles would leave most SMs without work
static int plan_bn(int M, int N, int sms, int pair_mode) {
  static const char* force_bn = getenv("BRUNO_TF32_BN");
  if (force_bn) return atoi(force_bn) == 64 ? 64 : 128;
  const int tiles128 = ceil_div(N, 128) * ceil_div(M, TG_BM) * (pair_mode == 1 ? 2 : 1);
  return (tiles128 * 2 <= sms || N <= 64) ? 64 : 128;
}
// split-K so that ~one CTA per SM is busy, at least two K tiles per CTA, at most TG_MAX_SPLIT slices (= the cluster size);
// pair mode 2 gives each of the two products half of the slices.  `split_ok`: a cluster launch, or a workspace that holds
// the partial tiles (`ws_floats`; 0 = cluster).  Returns gridDim.z and *kchunk = K tiles per slice.
static int plan_split(int nk, int tiles, int sms, int pair_mode, bool split_ok, size_t ws_floats, int bn, int* kchunk) {
  static const char* force_s = getenv("BRUNO_TF32_SPLIT");
  int S = 1;
  if (split_ok) {
    S = force_s ? atoi(force_s) : sms / tiles;
    const int smax = nk / 2 < TG_MAX_SPLIT ? nk / 2 : TG_MAX_SPLIT;
    S = S > smax ? smax : S;
    S = S < 1 ? 1 : S;
    while (ws_floats && S > 1 && static_cast<size_t>(tiles) * S * TG_BM * bn > ws_floats) --S;
  }
  if (pair_mode == 2) S = S / 2 < 1 ? 1 : S / 2;                       // slices per product
  *kchunk = ceil_div(nk, S);
  S = ceil_div(nk, *kchunk);
  return pair_mode == 2 ? 2 * S : S;
}

// cluster sizes whose launch failed on this process's device (bit S): never tried again (keeps stream captures clean)
static std::atomic<unsigned> g_cluster_refused{0u};

template <bool TA, bool TB, int BN>
int launch_tf32_bn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const GemmEpi& ep,
                   float* ws, size_t ws_floats, unsigned* tickets, int max_tiles, int sms, bool cluster, cudaStream_t st,
                   const GemmPair* pair) {
  // + the push buffer of the cluster split-K reduction (64-column tiles only): <= TG_BM + TG_MAX_SPLIT rows of BN + 4 floats
  constexpr size_t push_bytes = BN == 64 ? static_cast<size_t>(TG_BM + TG_MAX_SPLIT) * (BN + 4) * 4 : 0;
  constexpr size_t smem = 1024 + TG_STAGES * (2 * TG_TILE_BYTES + 2 * BN * TG_BK * 4) + 128 + push_bytes + 64;
  static_assert(sizeof(TgSmem) <= 128, "control block");
  cluster = cluster && BN == 64;                                       // 128-column tiles only run un-split in practice
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(tf32x3_gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  });
  // two GEMMs in one launch (GemmPair) only on the cluster path; otherwise 1 = "not taken": the caller launches them one by one
  const int mode = pair ? pair->mode : 0;
  if (mode && !cluster) return 1;
  const int nx = ceil_div(N, BN), ny = ceil_div(M, TG_BM), nxg = mode == 1 ? 2 * nx : nx, tiles = nxg * ny, nk = ceil_div(K, TG_BK);
  const bool have_ws = ws && tickets && tiles <= max_tiles;
  auto launch = [&](bool with_cluster) {
    int kchunk = nk;
    const int S = plan_split(nk, tiles, sms, mode, with_cluster || have_ws, with_cluster ? 0 : ws_floats, BN, &kchunk);
    if (with_cluster && S > 1 && (g_cluster_refused.load(std::memory_order_relaxed) >> S & 1u)) return cudaErrorInvalidClusterSize;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nxg, ny, S);
    cfg.blockDim = dim3(TG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the set-up overlaps the previous launch's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
    if (with_cluster && S > 1) {
      attr[1].id = cudaLaunchAttributeClusterDimension;                // the S CTAs of a tile reduce through peer shared memory
      attr[1].val.clusterDim.x = 1;
      attr[1].val.clusterDim.y = 1;
      attr[1].val.clusterDim.z = static_cast<unsigned>(S);
      cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, tf32x3_gemm_kernel<TA, TB, BN>, M, N, K, A, lda, B, ldb, C, ldc, ep, kchunk, ws,
                                             tickets, with_cluster ? 1 : 0, pair ? *pair : GemmPair{0, nullptr, nullptr, nullptr, nullptr},
                                             nx, g_gemm_trace_host);
    if (e != cudaSuccess && with_cluster && S > 1) {
      cudaGetLastError();
      g_cluster_refused.fetch_or(1u << S, std::memory_order_relaxed);
    }
    return e;
  };
  if (cluster && launch(true) == cudaSuccess) return check_launch("tf32x3_gemm_kernel");
  if (mode) return 1;
  launch(false);                                                       // workspace + ticket fix-up (or un-split)
  return check_launch("tf32x3_gemm_kernel");
}

template <bool TA, bool TB>
int launch_tf32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const GemmEpi& ep,
                float* ws, size_t ws_floats, unsigned* tickets, int max_tiles, int sms, bool cluster, cudaStream_t st, int* taken,
                const GemmPair* pair) {
  *taken = 1;
  const bool want64 = plan_bn(M, N, sms, pair ? pair->mode : 0) == 64;
  const int rc = want64 ? launch_tf32_bn<TA, TB, 64>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, pair)
                        : launch_tf32_bn<TA, TB, 128>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, pair);
  if (rc == 1) {                                                       // a pair that cannot run as one launch
    *taken = 0;
    return 0;
  }
  return rc;
}

}  // namespace

void gemm_tf32x3_set_trace(long long* device_buffer) { g_gemm_trace_host = device_buffer; }

void gemm_tf32x3_plan(int M, int N, int K, int pair_mode, int sms, int out[6]) {
  const int bn = plan_bn(M, N, sms, pair_mode);
  const bool cluster = bn == 64;                                        // 128-column tiles: no cluster, no pairs
  const int nx = ceil_div(N, bn), ny = ceil_div(M, TG_BM), nxg = pair_mode == 1 ? 2 * nx : nx, nk = ceil_div(K, TG_BK);
  int kchunk = nk;
  const int S = (pair_mode && !cluster) ? 0 : plan_split(nk, nxg * ny, sms, pair_mode, cluster, 0, bn, &kchunk);
  out[0] = bn; out[1] = nxg; out[2] = ny; out[3] = S; out[4] = kchunk; out[5] = (cluster && S > 1) ? S : 1;
}

int gemm_tf32x3(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const GemmEpi& ep, float* ws, size_t ws_floats, unsigned* tickets, int max_tiles, int sms, bool cluster,
                cudaStream_t st, int* taken, const GemmPair* pair) {
  *taken = 0;
  if (!ta && !tb) return launch_tf32<false, false>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, taken, pair);
  if (ta && !tb) return launch_tf32<true, false>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, taken, pair);
  if (!ta && tb) return launch_tf32<false, true>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, taken, pair);
  return launch_tf32<true, true>(M, N, K, A, lda, B, ldb, C, ldc, ep, ws, ws_floats, tickets, max_tiles, sms, cluster, st, taken, pair);
}

}  // namespace bruno
