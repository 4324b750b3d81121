// Dense (fully connected) coupling layers: CouplingLayerDense (nn_extra_nvp.py:190-274) with dense_wn (:408-432).
//   h1 = ELU((x b) V1 * g1/|V1| + b1);  h2 = ELU(h1 V2 * g2/|V2| + b2)
//   s = tanh(h2 Ks + bs)(1-b);  t = (h2 Kt + bt)(1-b);  y = x b + (1-b)(x e^s + t);  ld += sum s
// fp32 accuracy throughout (these layers carry 0.4 % of the FLOPs of bn2_omniglot and all of the parity budget of the
// dense en2 config).  GEMMs: the tcgen05 3xTF32 kernel of gemm_tf32.cu by default (fp32-accurate split on the tensor
// cores), or the register-tiled CUDA-core SGEMM below (strict fp32 FFMA; bruno_debug_gemm_variant(0), and the fallback for
// unaligned operands), both with the fused column-scale / bias / activation epilogue; plus the elementwise coupling
// kernels and the hand-written backward, orchestrated in C so that one coupling is ONE host call.
#include "common.cuh"
#include "dense_gemm.cuh"

namespace bruno {

constexpr int GB = 64;   // block tile M and N
constexpr int GK = 16;

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const int sz = valid ? 4 : 0;   // src-size 0 -> the 4 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int GSTAGES = 4;

// C[M,N] = op(A)[M,K] * op(B)[K,N];  transA: A stored [K,M];  transB: B stored [N,K].  Row-major, leading dims given.
// 64x64x16 block tile, 4x4 register tile, 4-stage cp.async (LDGSTS) pipeline: the K loop is latency bound otherwise
// (these GEMMs are 640 x 512 x 784: one partial wave of 80 CTAs).
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, GemmEpi ep) {
  __shared__ __align__(16) float As[GSTAGES][GK][GB + 4];
  __shared__ __align__(16) float Bs[GSTAGES][GK][GB + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  const int tx = tid & 15, ty = tid >> 4;  // thread computes rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + GK - 1) / GK;
  auto load_stage = [&](int kt) {
    const int st = kt % GSTAGES;
    const int k0 = kt * GK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (TA) { m = tid & 63; k = (tid >> 6) + 4 * i; }
      else    { k = tid & 15; m = (tid >> 4) + 16 * i; }
      const bool va = (m0 + m < M) && (k0 + k < K);
      const float* pa = TA ? A + static_cast<size_t>(k0 + k) * lda + m0 + m : A + static_cast<size_t>(m0 + m) * lda + k0 + k;
      cp_async4(&As[st][k][m], va ? pa : A, va);
      int n, kb;
      if (TB) { kb = tid & 15; n = (tid >> 4) + 16 * i; }
      else    { n = tid & 63; kb = (tid >> 6) + 4 * i; }
      const bool vb = (n0 + n < N) && (k0 + kb < K);
      const float* pb = TB ? B + static_cast<size_t>(n0 + n) * ldb + k0 + kb : B + static_cast<size_t>(k0 + kb) * ldb + n0 + n;
      cp_async4(&Bs[st][kb][n], vb ? pb : B, vb);
    }
  };
#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nk) load_stage(s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    if (kt + GSTAGES - 1 < nk) load_stage(kt + GSTAGES - 1);
    cp_async_commit();
    const int st = kt % GSTAGES;
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
      if (ep.a_kscale) {
        const float sc = (kt * GK + k < K) ? __ldg(ep.a_kscale + kt * GK + k) : 0.f;
        a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc;
      }
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.col_scale) v *= ep.col_scale[n];
      if (ep.bias) v += ep.bias[n];
      const size_t o = static_cast<size_t>(m) * ldc + n;
      if (ep.pre_out) ep.pre_out[o] = v;
      if (ep.act == 1) v = v > 0.f ? v : expm1f(v);
      else if (ep.act == 2) v = v > 0.f ? v : 0.2f * v;   // tf.nn.leaky_relu default alpha
      else if (ep.act == 3) v = tanhf(v);
      if (ep.accumulate) v += C[o];
      C[o] = v;
    }
  }
}

// Fast path for 16-byte aligned operands (all shapes of the target configs).
//   * both operands are staged in their NATURAL orientation by 16-byte cp.async (LDGSTS.128, zero-fill at the edges)
//     through a 4-stage ring: the K loop of these skinny GEMMs (640 x 512 x 784; 40 x 512 x 784 for few-shot scoring) is
//     bound by global-load latency, not FFMA issue, so three K tiles are kept in flight;
//   * block tile (16 MI) x 64 x 16 with an MI x 4 register tile (MI = 8 for M >= 192, else 4);
//   * split-K over gridDim.z so that the grid covers ~2 CTAs per SM even when the output has only 8-80 tiles.  Partials
//     go to a per-tile workspace; the LAST CTA to finish a tile (ticket counter) sums them in z order -- deterministic --
//     and runs the fused epilogue.  The counter is reset by that CTA, so launches are self-cleaning and graph-safe.
//     The workspace is shared by all launches of this process: SGEMMs must be issued on one stream (they are).
// Requires lda, ldb % 4 == 0, 16-byte aligned A, B, K % 4 == 0, and M % 4 (transA) / N % 4 (!transB).
constexpr int SK_STAGES = 4;
constexpr int SK_BN = 64;

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <bool TA, int MI> struct SkShape {
  static constexpr int BM = 16 * MI;
  static constexpr int A_FLOATS = TA ? GK * (BM + 4) : BM * (GK + 4);
};
template <bool TB> struct SkShapeB {
  static constexpr int B_FLOATS = TB ? SK_BN * (GK + 4) : GK * (SK_BN + 4);
};

template <bool TA, bool TB, int MI>
__global__ void __launch_bounds__(256, 2)
sgemm_sk_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                float* __restrict__ C, int ldc, GemmEpi ep, int kchunk, float* __restrict__ ws, unsigned* __restrict__ tickets) {
  constexpr int BM = SkShape<TA, MI>::BM;
  constexpr int A_FLOATS = SkShape<TA, MI>::A_FLOATS, B_FLOATS = SkShapeB<TB>::B_FLOATS;
  constexpr int APITCH = TA ? BM + 4 : GK + 4, BPITCH = TB ? GK + 4 : SK_BN + 4;
  extern __shared__ __align__(16) float sk_smem[];
  float* As = sk_smem;                          // [SK_STAGES][A_FLOATS]
  float* Bs = sk_smem + SK_STAGES * A_FLOATS;   // [SK_STAGES][B_FLOATS]
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * SK_BN;
  const int tx = tid & 15, ty = tid >> 4;
  const int nk = (K + GK - 1) / GK;
  const int kt0 = blockIdx.z * kchunk, kt1 = min(nk, kt0 + kchunk);

  float acc[MI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto load_stage = [&](int kt) {
    const int st = kt % SK_STAGES, k0 = kt * GK;
    float* as = As + st * A_FLOATS;
    float* bs = Bs + st * B_FLOATS;
#pragma unroll
    for (int v = tid; v < BM * GK / 4; v += 256) {
      if (TA) {   // A stored [K, M]: vectors along m
        const int k = v / (BM / 4), m4 = (v % (BM / 4)) * 4;
        const bool ok = (k0 + k < K) && (m0 + m4 < M);
        cp_async16(as + k * APITCH + m4, ok ? A + static_cast<size_t>(k0 + k) * lda + m0 + m4 : A, ok);
      } else {    // A stored [M, K]: vectors along k
        const int m = v >> 2, k4 = (v & 3) * 4;
        const bool ok = (m0 + m < M) && (k0 + k4 < K);
        cp_async16(as + m * APITCH + k4, ok ? A + static_cast<size_t>(m0 + m) * lda + k0 + k4 : A, ok);
      }
    }
    {
      const int v = tid;
      if (TB) {   // B stored [N, K]
        const int n = v >> 2, k4 = (v & 3) * 4;
        const bool ok = (n0 + n < N) && (k0 + k4 < K);
        cp_async16(bs + n * BPITCH + k4, ok ? B + static_cast<size_t>(n0 + n) * ldb + k0 + k4 : B, ok);
      } else {    // B stored [K, N]
        const int k = v >> 4, n4 = (v & 15) * 4;
        const bool ok = (k0 + k < K) && (n0 + n4 < N);
        cp_async16(bs + k * BPITCH + n4, ok ? B + static_cast<size_t>(k0 + k) * ldb + n0 + n4 : B, ok);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < SK_STAGES - 1; ++s) {
    if (kt0 + s < kt1) load_stage(kt0 + s);
    cp_async_commit();
  }
  for (int kt = kt0; kt < kt1; ++kt) {
    cp_async_wait<SK_STAGES - 2>();
    __syncthreads();
    if (kt + SK_STAGES - 1 < kt1) load_stage(kt + SK_STAGES - 1);
    cp_async_commit();
    const float* as = As + (kt % SK_STAGES) * A_FLOATS;
    const float* bs = Bs + (kt % SK_STAGES) * B_FLOATS;
#pragma unroll
    for (int k4 = 0; k4 < GK; k4 += 4) {
      float a[MI][4], b[4][4];   // [row or col][k within the group of 4]
      if (TA) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int h = 0; h < MI / 4; ++h) {
            const float4 v = *reinterpret_cast<const float4*>(as + (k4 + kk) * APITCH + ty * 4 + 64 * h);
            a[4 * h][kk] = v.x; a[4 * h + 1][kk] = v.y; a[4 * h + 2][kk] = v.z; a[4 * h + 3][kk] = v.w;
          }
      } else {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(as + (ty * 4 + (i & 3) + 64 * (i >> 2)) * APITCH + k4);
          a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
        }
      }
      if (TB) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(bs + (tx + 16 * j) * BPITCH + k4);
          b[j][0] = v.x; b[j][1] = v.y; b[j][2] = v.z; b[j][3] = v.w;
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 v = *reinterpret_cast<const float4*>(bs + (k4 + kk) * BPITCH + tx * 4);
          b[0][kk] = v.x; b[1][kk] = v.y; b[2][kk] = v.z; b[3][kk] = v.w;
        }
      }
      if (ep.a_kscale) {   // (A diag(s)) B == A (diag(s) B): scale the 16 b values instead of the 4 MI a values
        const int kg = kt * GK + k4;
        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kg < K) sc = __ldg(reinterpret_cast<const float4*>(ep.a_kscale + kg));
#pragma unroll
        for (int j = 0; j < 4; ++j) { b[j][0] *= sc.x; b[j][1] *= sc.y; b[j][2] *= sc.z; b[j][3] *= sc.w; }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
    }
  }
  cp_async_wait<0>();

  const int S = gridDim.z;
  if (S > 1) {
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    float4* wt = reinterpret_cast<float4*>(ws) + static_cast<size_t>(tile) * S * (MI * 256);
    float4* mine = wt + static_cast<size_t>(blockIdx.z) * (MI * 256);
#pragma unroll
    for (int i = 0; i < MI; ++i) __stcg(mine + i * 256 + tid, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(tickets + tile, 1u) == static_cast<unsigned>(S - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < MI; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    constexpr int ZU = MI == 4 ? 4 : 2;   // partials fetched per round trip (loads first, then the ordered adds)
    for (int z = 0; z < S; z += ZU) {
      float4 v[ZU][MI];
#pragma unroll
      for (int u = 0; u < ZU; ++u)
#pragma unroll
        for (int i = 0; i < MI; ++i)
          v[u][i] = z + u < S ? __ldcg(wt + static_cast<size_t>(z + u) * (MI * 256) + i * 256 + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < ZU; ++u)
#pragma unroll
        for (int i = 0; i < MI; ++i) { acc[i][0] += v[u][i].x; acc[i][1] += v[u][i].y; acc[i][2] += v[u][i].z; acc[i][3] += v[u][i].w; }
    }
    if (tid == 0) tickets[tile] = 0u;
  }

#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + (TB ? tx + 16 * j : tx * 4 + j);
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.col_scale) v *= ep.col_scale[n];
      if (ep.bias) v += ep.bias[n];
      const size_t o = static_cast<size_t>(m) * ldc + n;
      if (ep.pre_out) ep.pre_out[o] = v;
      if (ep.act == 1) v = v > 0.f ? v : expm1f(v);
      else if (ep.act == 2) v = v > 0.f ? v : 0.2f * v;
      else if (ep.act == 3) v = tanhf(v);
      if (ep.accumulate) v += C[o];
      C[o] = v;
    }
  }
}

// Split-K / column-reduction workspace.  The library never allocates device memory: the caller allocates
// bruno_workspace_bytes() zero-initialised bytes and binds them to the stream it launches on (bruno_workspace_bind).  A launch
// on a stream without a bound workspace runs un-split (correct, slower).  One workspace per stream: the tickets and the
// partial tiles of two GEMMs in flight must not alias.
constexpr size_t SK_WS_FLOATS = size_t(4) << 20;   // 16 MB
constexpr int SK_MAX_TILES = 4096;
constexpr size_t CR_WS_FLOATS = size_t(64) << 10;  // column-reduction partials, after the split-K region
constexpr int CR_MAX_GROUPS = 1024;                // column-reduction tickets, after the split-K tickets
constexpr size_t WS_BYTES = (SK_WS_FLOATS + CR_WS_FLOATS) * sizeof(float) + (SK_MAX_TILES + CR_MAX_GROUPS) * sizeof(unsigned);

struct Workspace {
  float* ws = nullptr;           // [SK_WS_FLOATS + CR_WS_FLOATS]
  unsigned* tickets = nullptr;   // [SK_MAX_TILES + CR_MAX_GROUPS], zero between launches
  int sms = 0;
  explicit operator bool() const { return ws != nullptr; }
};
struct WsEntry { cudaStream_t stream; int device; Workspace w; bool used; };
constexpr int WS_MAX_BOUND = 64;
static WsEntry g_ws_table[WS_MAX_BOUND];         // caller-owned buffers, keyed by (device, stream)
static std::mutex g_ws_mutex;

static Workspace sk_workspace(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (int i = 0; i < WS_MAX_BOUND; ++i)
    if (g_ws_table[i].used && g_ws_table[i].stream == st && g_ws_table[i].device == dev) return g_ws_table[i].w;
  return Workspace{};
}

#ifndef BRUNO_NO_DEBUG
static int g_gemm_variant = 1;   // 0: CUDA-core SGEMM, 1: tcgen05 3xTF32 (gemm_tf32.cu), 2: the same with the workspace split-K fix-up
#else
constexpr int g_gemm_variant = 1;
#endif

template <bool TA, bool TB, int MI>
static int launch_sk(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const GemmEpi& ep,
                     cudaStream_t st) {
  constexpr int BM = SkShape<TA, MI>::BM;
  constexpr size_t smem = size_t(SK_STAGES) * (SkShape<TA, MI>::A_FLOATS + SkShapeB<TB>::B_FLOATS) * sizeof(float);
  static OncePerDevice once;
  once.run([&] {
    cudaFuncSetAttribute(sgemm_sk_kernel<TA, TB, MI>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  });
  const int nx = ceil_div(N, SK_BN), ny = ceil_div(M, BM), tiles = nx * ny, nk = ceil_div(K, GK);
  int S = 1;
  const Workspace wk = sk_workspace(st);
  if (tiles <= SK_MAX_TILES && wk) {
    S = (2 * wk.sms + tiles / 2) / tiles;
    S = max(1, min(S, nk / 4));
    const size_t per_split = static_cast<size_t>(tiles) * BM * SK_BN;
    while (S > 1 && per_split * S > SK_WS_FLOATS) --S;
  }
  const int kchunk = ceil_div(nk, S);
  S = ceil_div(nk, kchunk);
  sgemm_sk_kernel<TA, TB, MI><<<dim3(nx, ny, S), 256, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, kchunk, wk.ws, wk.tickets);
  return check_launch("sgemm_sk_kernel");
}

static int gemm(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const GemmEpi& ep, cudaStream_t st) {
  const bool vec_ok = aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0 && K % 4 == 0 &&
                      (ta ? M % 4 == 0 : true) && (tb ? true : N % 4 == 0) && (!ep.a_kscale || aligned16(ep.a_kscale));
  const Workspace wk = sk_workspace(st);
  if (vec_ok && g_gemm_variant >= 1) {
    // the cluster split-K of the tcgen05 GEMM needs no workspace; without one the fix-up variant (2) runs un-split
    static int sms_cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = wk ? wk.sms : sms_cached[dev & 63];
    if (sms == 0) {
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      sms_cached[dev & 63] = sms;
    }
    int taken = 0;
    const int rc = gemm_tf32x3(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, ep, wk.ws, SK_WS_FLOATS, wk.tickets, SK_MAX_TILES,
                               sms, g_gemm_variant == 1, st, &taken);
    if (taken) return rc;
  }
  if (vec_ok) {
    const bool big = M >= 192;
#define SK_DISPATCH(TA_, TB_)                                                                            \
  return big ? launch_sk<TA_, TB_, 8>(M, N, K, A, lda, B, ldb, C, ldc, ep, st)                           \
             : launch_sk<TA_, TB_, 4>(M, N, K, A, lda, B, ldb, C, ldc, ep, st)
    if (!ta && !tb) { SK_DISPATCH(false, false); }
    else if (ta && !tb) { SK_DISPATCH(true, false); }
    else if (!ta && tb) { SK_DISPATCH(false, true); }
    else { SK_DISPATCH(true, true); }
#undef SK_DISPATCH
  }
  dim3 grid(ceil_div(N, GB), ceil_div(M, GB));
  if (!ta && !tb) sgemm_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else if (ta && !tb) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else if (!ta && tb) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  else sgemm_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep);
  return check_launch("sgemm_kernel");
}

// Two GEMMs of one shape as ONE launch of the tcgen05 kernel (GemmPair, dense_gemm.cuh): mode 1 = side by side (same flags,
// own A / B / C / bias), mode 2 = C = op(A) op(B) + op(A2) op(B2).  Returns 1 when launched, 0 when the caller has to run
// the two GEMMs one after the other (operands not 16-byte aligned, another GEMM variant selected, cluster launch refused),
// a negative code on error.
static int gemm_pair(int mode, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                     int ldc, const GemmEpi& ep, const float* A2, const float* B2, float* C2, const float* bias2, cudaStream_t st) {
  const bool vec_ok = aligned16(A) && aligned16(B) && aligned16(A2) && aligned16(B2) && lda % 4 == 0 && ldb % 4 == 0 && K % 4 == 0 &&
                      (ta ? M % 4 == 0 : true) && (tb ? true : N % 4 == 0) && !ep.a_kscale;
  if (!vec_ok || g_gemm_variant != 1) return 0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  const Workspace wk = sk_workspace(st);
  if (wk) sms = wk.sms;
  else cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const GemmPair pair{mode, A2, B2, C2, bias2};
  int taken = 0;
  const int rc = gemm_tf32x3(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, ep, nullptr, 0, nullptr, 0, sms, true, st, &taken, &pair);
  if (rc) return rc < 0 ? rc : -1;
  return taken;
}

__device__ __forceinline__ float dense_mask_b(int mask_type, int d) {
  // nn_extra_nvp.py:199-221: 'even' -> b = idx % 2, 'odd' -> 1 - idx % 2 on the flattened (h, w, c) index
  return mask_type == BRUNO_MASK_EVEN ? static_cast<float>(d & 1) : static_cast<float>(1 - (d & 1));
}

// Column reductions: block = 32 columns x 8 row groups (256 threads); a warp reads 32 consecutive columns of one row
// (coalesced), the 8 row groups are combined in shared memory in a fixed order (deterministic).
__device__ __forceinline__ float col_block_reduce(float v, float (*sh)[33]) {
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  sh[r][c] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += sh[i][c];
  __syncthreads();
  return s;
}

// sc[j] = g[j] / ||V[:, j]||_2   (tf.norm, no epsilon: nvp:419);  inv_nrm[j] = 1 / ||V[:, j]||
__global__ void __launch_bounds__(1024)
dense_wn_scale_kernel(const float* __restrict__ V, const float* __restrict__ g, float* __restrict__ sc,
                      float* __restrict__ inv_nrm, int K, int N) {
  // 32 columns x 32 row groups per block, four loads in flight per thread, fixed-order combine
  __shared__ float sh[32][33];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
  if (j < N) {
    for (int k = rg; k < K; k += 128) {
      const float t0 = V[static_cast<size_t>(k) * N + j];
      const float t1 = k + 32 < K ? V[static_cast<size_t>(k + 32) * N + j] : 0.f;
      const float t2 = k + 64 < K ? V[static_cast<size_t>(k + 64) * N + j] : 0.f;
      const float t3 = k + 96 < K ? V[static_cast<size_t>(k + 96) * N + j] : 0.f;
      n0 = fmaf(t0, t0, n0); n1 = fmaf(t1, t1, n1); n2 = fmaf(t2, t2, n2); n3 = fmaf(t3, t3, n3);
    }
  }
  sh[rg][c] = (n0 + n1) + (n2 + n3);
  __syncthreads();
  if (j < N && threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += sh[r][c];
    const float inv = rsqrtf(t);
    sc[j] = g[j] * inv;
    if (inv_nrm) inv_nrm[j] = inv;
  }
}

__global__ void dense_mask_mul_kernel(const float* __restrict__ x, float* __restrict__ xm, size_t total, int D, int mask_type) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) xm[i] = x[i] * dense_mask_b(mask_type, static_cast<int>(i % D));
}

__device__ __forceinline__ float block_sum_256d(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// one block per row
__global__ void __launch_bounds__(256)
dense_coupling_kernel(const float* __restrict__ x, const float* __restrict__ s_pre, const float* __restrict__ t_pre,
                      const float* __restrict__ ld_in, float* __restrict__ y, float* __restrict__ ld_out, int D,
                      int mask_type, int inverse) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const size_t o = static_cast<size_t>(n) * D;
  float ssum = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float xv = x[o + d];
    float out = xv;
    if (dense_mask_b(mask_type, d) == 0.f) {
      const float s = tanhf(s_pre[o + d]);
      ssum += s;
      out = inverse ? (xv - t_pre[o + d]) * expf(-s) : fmaf(xv, expf(s), t_pre[o + d]);
    }
    y[o + d] = out;
  }
  const float tot = block_sum_256d(ssum, red);
  if (threadIdx.x == 0) ld_out[n] = ld_in[n] + (inverse ? -tot : tot);
}

__global__ void dense_coupling_bwd_kernel(const float* __restrict__ x, const float* __restrict__ s_pre,
                                          const float* __restrict__ dy, const float* __restrict__ dld,
                                          float* __restrict__ dx, float* __restrict__ ds_pre, float* __restrict__ dt_pre,
                                          size_t total, int D, int mask_type) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = static_cast<int>(i % D);
  const size_t n = i / D;
  const float g = dy[i];
  float dxv = g, dsp = 0.f, dtp = 0.f;
  if (dense_mask_b(mask_type, d) == 0.f) {
    const float s = tanhf(s_pre[i]);
    const float es = expf(s);
    dxv = g * es;
    dtp = g;
    dsp = (g * x[i] * es + dld[n]) * (1.f - s * s);
  }
  dx[i] = dxv;
  ds_pre[i] = dsp;
  dt_pre[i] = dtp;
}

// dpre = dh * ELU'(pre)
__global__ void elu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dh, float* __restrict__ dpre,
                               size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float p = pre[i];
  dpre[i] = dh[i] * (p > 0.f ? 1.f : expf(p));
}

// out1[j] = sum_m A[m,j];  out2[j] = sum_m A[m,j] * Bm[m,j] (if Bm)
// grid (N/32, MS): block (x, y) reduces rows [y M/MS, (y+1) M/MS) of 32 columns with four loads in flight per thread;
// with MS > 1 the partials go to the workspace and the last block of a column group (ticket) adds them in y order.
constexpr int CR_MAX_SPLIT = 16;
// The column reduction also PRODUCES its operand where that is an elementwise map (one pass over the matrix, one launch
// instead of two or three), same summation order in every mode (bit-identical sums):
//   CR_PLAIN   : v = A[m,j], w = Bm[m,j]                                        out1 = sum v, out2 = sum v w
//   CR_ELU_BWD : v = A[m,j] ELU'(Bm[m,j]) (A = dh, Bm = pre), stored to st0     out1 = sum v (db), out2 = sum v pre (r1)
//   CR_COUPLING: the dense coupling backward (dense_coupling_bwd_kernel) at (m,j); dx -> st0, ds_pre -> st1, dt_pre -> st2
//                                                                               out1 = sum ds_pre (dbs), out2 = sum dt_pre (dbt)
enum { CR_PLAIN = 0, CR_ELU_BWD = 1, CR_COUPLING = 2 };
struct ColReduceIo {
  const float* A;      // CR_PLAIN: the matrix; CR_ELU_BWD: dh; CR_COUPLING: dy
  const float* Bm;     // CR_PLAIN: second factor or null; CR_ELU_BWD: pre; CR_COUPLING: s_pre
  const float* x;      // CR_COUPLING
  const float* dld;    // CR_COUPLING: [M]
  float *st0, *st1, *st2;
  int mask_type;
};
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(ColReduceIo io, float* __restrict__ out1, float* __restrict__ out2, int M, int N, float* __restrict__ part,
                  unsigned* __restrict__ tickets) {
  __shared__ float sh[8][33];
  __shared__ int s_last;
  const float* __restrict__ A = io.A;
  const float* __restrict__ Bm = io.Bm;
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  const int MS = gridDim.y;
  const int rows = (M + MS - 1) / MS;
  const int m_begin = blockIdx.y * rows, m_end = min(M, m_begin + rows);
  const bool two = MODE != CR_PLAIN || Bm != nullptr;
  const bool masked = MODE == CR_COUPLING && dense_mask_b(io.mask_type, j) == 0.f;   // the transformed half of the columns
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (j < N) {
    for (int m = m_begin + rg; m < m_end; m += 32) {
      float v[4], w[4], xv[4], dl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mm = m + 8 * u;
        const bool on = mm < m_end;
        const size_t o = static_cast<size_t>(mm) * N + j;
        v[u] = on ? A[o] : 0.f;
        w[u] = (on && (MODE != CR_PLAIN || Bm) && (MODE != CR_COUPLING || masked)) ? Bm[o] : 0.f;
        if (MODE == CR_COUPLING) {
          xv[u] = (on && masked) ? io.x[o] : 0.f;
          dl[u] = (on && masked) ? io.dld[mm] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mm = m + 8 * u;
        const size_t o = static_cast<size_t>(mm) * N + j;
        if (MODE == CR_PLAIN) {
          a[u] += v[u];
          b[u] = fmaf(v[u], w[u], b[u]);
        } else if (MODE == CR_ELU_BWD) {
          const float d = v[u] * (w[u] > 0.f ? 1.f : expf(w[u]));          // elu_bwd_kernel
          if (mm < m_end) io.st0[o] = d;
          a[u] += d;
          b[u] = fmaf(d, w[u], b[u]);
        } else {
          const float g = v[u];                                              // dense_coupling_bwd_kernel
          float dxv = g, dsp = 0.f, dtp = 0.f;
          if (masked) {
            const float s_ = tanhf(w[u]);
            const float es = expf(s_);
            dxv = g * es;
            dtp = g;
            dsp = (g * xv[u] * es + dl[u]) * (1.f - s_ * s_);
          }
          if (mm < m_end) {
            io.st0[o] = dxv;
            io.st1[o] = dsp;
            io.st2[o] = dtp;
          }
          a[u] += dsp;                                                       // rows past m_end: g = 0 -> both 0
          b[u] += dtp;
        }
      }
    }
  }
  float sa = col_block_reduce((a[0] + a[1]) + (a[2] + a[3]), sh);
  float sb = two ? col_block_reduce((b[0] + b[1]) + (b[2] + b[3]), sh) : 0.f;
  if (MS > 1) {
    // partial layout [column group][y][2][32]
    float* mine = part + (static_cast<size_t>(blockIdx.x) * MS + blockIdx.y) * 64;
    if (threadIdx.x < 32) {
      __stcg(mine + c, sa);
      __stcg(mine + 32 + c, sb);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.x, 1u) == static_cast<unsigned>(MS - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 32) {
      sa = 0.f;
      sb = 0.f;
      const float* base = part + static_cast<size_t>(blockIdx.x) * MS * 64;
      for (int y = 0; y < MS; ++y) {
        sa += __ldcg(base + y * 64 + c);
        sb += __ldcg(base + y * 64 + 32 + c);
      }
      if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
    }
  }
  if (j < N && threadIdx.x < 32) {
    out1[j] = sa;
    if (out2) out2[j] = sb;
  }
}

template <int MODE>
static void col_reduce_io(const ColReduceIo& io, float* out1, float* out2, int M, int N, cudaStream_t st) {
  const int groups = ceil_div(N, 32);
  int ms = 1;
  const Workspace wk = sk_workspace(st);
  if (groups <= CR_MAX_GROUPS && wk) {
    ms = max(1, min(CR_MAX_SPLIT, min(M / 32, (2 * wk.sms) / groups)));
    while (ms > 1 && static_cast<size_t>(groups) * ms * 64 > CR_WS_FLOATS) --ms;
  }
  col_reduce_kernel<MODE><<<dim3(groups, ms), 256, 0, st>>>(io, out1, out2, M, N, wk ? wk.ws + SK_WS_FLOATS : nullptr,
                                                            wk ? wk.tickets + SK_MAX_TILES : nullptr);
}
static void col_reduce(const float* A, const float* Bm, float* out1, float* out2, int M, int N, cudaStream_t st) {
  col_reduce_io<CR_PLAIN>(ColReduceIo{A, Bm, nullptr, nullptr, nullptr, nullptr, nullptr, 0}, out1, out2, M, N, st);
}
// dpre = dh ELU'(pre), db = colsum(dpre), r1 = colsum(dpre pre) in one pass
static void elu_bwd_col_reduce(const float* pre, const float* dh, float* dpre, float* db, float* r1, int M, int N, cudaStream_t st) {
  col_reduce_io<CR_ELU_BWD>(ColReduceIo{dh, pre, nullptr, nullptr, dpre, nullptr, nullptr, 0}, db, r1, M, N, st);
}
// the dense coupling's elementwise backward with dbs = colsum(ds_pre), dbt = colsum(dt_pre) in one pass
static void coupling_bwd_col_reduce(const float* x, const float* s_pre, const float* dy, const float* dld, float* dx, float* ds_pre,
                                    float* dt_pre, float* dbs, float* dbt, int M, int N, int mask_type, cudaStream_t st) {
  col_reduce_io<CR_COUPLING>(ColReduceIo{dy, s_pre, x, dld, dx, ds_pre, dt_pre, mask_type}, dbs, dbt, M, N, st);
}

// weight-norm backward for dense_wn.  dVraw = x^T dpre;  r1 = colsum(dpre * pre);  db = colsum(dpre).
//   u = (pre - b)/sc  =>  dsc = (r1 - b db)/sc;  dg = dsc/nrm;  dV = dVraw sc - dsc g / nrm^3 V
// sc and inv_nrm = 1/nrm come from the forward pass.  Elementwise over [K, N] (coalesced).
__global__ void __launch_bounds__(256)
dense_wn_bwd_kernel(const float* __restrict__ V, const float* __restrict__ g, const float* __restrict__ b,
                    const float* __restrict__ sc, const float* __restrict__ inv_nrm, const float* __restrict__ r1,
                    const float* __restrict__ db, float* __restrict__ dV /*in: dVraw*/, float* __restrict__ dg, int K, int N) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(K) * N) return;
  const int j = static_cast<int>(i % N);
  const float s_ = sc[j], inv = inv_nrm[j];
  const float dsc = fabsf(s_) > 1e-30f ? (r1[j] - b[j] * db[j]) / s_ : 0.f;
  if (i < static_cast<size_t>(N)) dg[j] = dsc * inv;
  dV[i] = dV[i] * s_ - dsc * g[j] * inv * inv * inv * V[i];
}

__global__ void masked_accumulate_kernel(float* __restrict__ dx, const float* __restrict__ dxm, size_t total, int D, int mask_type) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) dx[i] += dxm[i] * dense_mask_b(mask_type, static_cast<int>(i % D));
}

static inline unsigned nblk(size_t total) { return static_cast<unsigned>((total + 255) / 256); }

// ---- label-conditioned dense coupling (nn_extra_nvp_conditional.py:162-266, 350-369, 437-471) -----------------------
// dst[m, off + j] = src[m, j] (* dense mask b(j) if mask_type >= 0), dst row pitch ld
__global__ void copy_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t total, int n, int ld, int off,
                                 int mask_type) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / n;
  const int j = static_cast<int>(i - m * n);
  const float b = mask_type >= 0 ? dense_mask_b(mask_type, j) : 1.f;
  dst[m * ld + off + j] = src[i] * b;
}
// dst[m, j] = src[m, off + j], src row pitch ld
__global__ void slice_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t total, int n, int ld, int off) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t m = i / n;
  dst[i] = src[m * ld + off + static_cast<int>(i - m * n)];
}
// dpre = dh * leaky_relu'(h): the sign of the output is the sign of the pre-activation (alpha = 0.2)
__global__ void leaky_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dh, float* __restrict__ dpre, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) dpre[i] = dh[i] * (h[i] > 0.f ? 1.f : 0.2f);
}
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) a[i] += b[i];
}

struct CondDenseWs {
  float *scL1, *inL1, *scL2, *inL2, *sc1, *in1, *sc2, *in2;
  float *hL1, *preL1, *o1, *h1, *pre1, *hL2, *preL2, *o2, *h2, *pre2, *hs, *ht, *s_pre, *t_pre;
};
static size_t cond_dense_ws_floats(int N, int D, int U, int L) {
  const size_t n = N, nh1 = D / 2, nh2 = U / 2;
  auto r4 = [](size_t v) { return (v + 3) / 4 * 4; };
  return 2 * r4(nh1) + 2 * r4(nh2) + 4 * r4(U) + 2 * r4(n * nh1) + r4(n * (nh1 + D)) + 2 * r4(n * U) + 2 * r4(n * nh2) +
         r4(n * (nh2 + U)) + 2 * r4(n * U) + 2 * r4(n * nh2) + 2 * r4(n * D);
}
static CondDenseWs cond_carve(float* ws, int N, int D, int U) {
  CondDenseWs w;
  const size_t n = N, nh1 = D / 2, nh2 = U / 2;
  size_t o = 0;
  auto take = [&](size_t v) { float* p = ws + o; o += (v + 3) / 4 * 4; return p; };
  w.scL1 = take(nh1); w.inL1 = take(nh1); w.scL2 = take(nh2); w.inL2 = take(nh2);
  w.sc1 = take(U); w.in1 = take(U); w.sc2 = take(U); w.in2 = take(U);
  w.hL1 = take(n * nh1); w.preL1 = take(n * nh1); w.o1 = take(n * (nh1 + D)); w.h1 = take(n * U); w.pre1 = take(n * U);
  w.hL2 = take(n * nh2); w.preL2 = take(n * nh2); w.o2 = take(n * (nh2 + U)); w.h2 = take(n * U); w.pre2 = take(n * U);
  w.hs = take(n * nh2); w.ht = take(n * nh2); w.s_pre = take(n * D); w.t_pre = take(n * D);
  return w;
}
// order of the 20 variables of a label-conditioned dense coupling (construction order of the reference graph)
enum CondVar { cLV1, cLg1, cLb1, cV1, cg1, cb1, cLV2, cLg2, cLb2, cV2, cg2, cb2, cWs, cbws, cKs, cbs, cWt, cbwt, cKt, cbt, cNVAR };

struct DenseWs {
  float *sc1, *sc2, *in1, *in2, *xm, *pre1, *pre2, *h1, *h2, *s_pre, *t_pre;
};
static DenseWs carve(float* ws, int N, int D, int U) {
  DenseWs w;
  size_t o = 0;
  auto take = [&](size_t n) { float* p = ws + o; o += (n + 3) / 4 * 4; return p; };
  w.sc1 = take(U); w.sc2 = take(U); w.in1 = take(U); w.in2 = take(U);
  w.xm = take(static_cast<size_t>(N) * D);
  w.pre1 = take(static_cast<size_t>(N) * U); w.pre2 = take(static_cast<size_t>(N) * U);
  w.h1 = take(static_cast<size_t>(N) * U); w.h2 = take(static_cast<size_t>(N) * U);
  w.s_pre = take(static_cast<size_t>(N) * D); w.t_pre = take(static_cast<size_t>(N) * D);
  return w;
}

}  // namespace bruno

using namespace bruno;

extern "C" {

size_t bruno_coupling_dense_ws_floats(int N, int D, int U) {
  return 4 * (static_cast<size_t>(U) + 4) + 3 * (static_cast<size_t>(N) * D + 4) + 4 * (static_cast<size_t>(N) * U + 4);
}
size_t bruno_coupling_dense_bwd_ws_floats(int N, int D, int U) {
  return 3 * (static_cast<size_t>(N) * D + 4) + 2 * (static_cast<size_t>(N) * U + 4) + 4 * (static_cast<size_t>(U) + 4);
}

#ifndef BRUNO_NO_DEBUG
int bruno_debug_gemm_variant(int variant) {
  if (variant < 0) return g_gemm_variant;
  BRUNO_REQUIRE(variant >= 0 && variant <= 2,
                "bruno_debug_gemm_variant: 0 (CUDA-core SGEMM), 1 (tcgen05 3xTF32, cluster split-K) or 2 (tcgen05, workspace split-K)");
  g_gemm_variant = variant;
  return BRUNO_OK;
}

int bruno_debug_gemm_plan(int M, int N, int K, int pair_mode, int sms, int* out6) {
  BRUNO_REQUIRE(M > 0 && N > 0 && K > 0 && pair_mode >= 0 && pair_mode <= 2 && sms > 0 && out6, "bruno_debug_gemm_plan: bad arguments");
  gemm_tf32x3_plan(M, N, K, pair_mode, sms, out6);
  return BRUNO_OK;
}

int bruno_debug_gemm_trace(long long* device_buffer) {
  gemm_tf32x3_set_trace(device_buffer);
  return BRUNO_OK;
}
#endif

size_t bruno_workspace_bytes(void) { return WS_BYTES; }

int bruno_workspace_bind(void* stream, void* workspace, size_t bytes) {
  BRUNO_REQUIRE(workspace && bytes >= WS_BYTES && aligned16(workspace), "bruno_workspace_bind: need >= %zu zero-initialised, 16-byte aligned bytes",
                WS_BYTES);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  int slot = -1;
  for (int i = 0; i < WS_MAX_BOUND; ++i) {
    if (g_ws_table[i].used && g_ws_table[i].stream == st && g_ws_table[i].device == dev) { slot = i; break; }
    if (!g_ws_table[i].used && slot < 0) slot = i;
  }
  BRUNO_REQUIRE(slot >= 0, "bruno_workspace_bind: more than %d bound workspaces", WS_MAX_BOUND);
  Workspace w;
  w.ws = static_cast<float*>(workspace);
  w.tickets = reinterpret_cast<unsigned*>(w.ws + SK_WS_FLOATS + CR_WS_FLOATS);
  w.sms = sms;
  g_ws_table[slot] = WsEntry{st, dev, w, true};
  return BRUNO_OK;
}

int bruno_workspace_unbind(void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (int i = 0; i < WS_MAX_BOUND; ++i)
    if (g_ws_table[i].used && g_ws_table[i].stream == static_cast<cudaStream_t>(stream) && g_ws_table[i].device == dev)
      g_ws_table[i].used = false;
  return BRUNO_OK;
}

int bruno_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                int ldc, const float* col_scale, const float* bias, int act, int accumulate, void* stream) {
  BRUNO_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "bruno_sgemm: bad arguments");
  GemmEpi ep{col_scale, bias, nullptr, act, accumulate, nullptr};
  return gemm(transA != 0, transB != 0, M, N, K, A, lda, B, ldb, C, ldc, ep, static_cast<cudaStream_t>(stream));
}

int bruno_dense_wn_scale(const float* V, const float* g, float* sc, float* inv_nrm, int K, int N, void* stream) {
  BRUNO_REQUIRE(V && g && sc && K > 0 && N > 0, "bruno_dense_wn_scale: bad arguments");
  dense_wn_scale_kernel<<<ceil_div(N, 32), 1024, 0, static_cast<cudaStream_t>(stream)>>>(V, g, sc, inv_nrm, K, N);
  return check_launch("dense_wn_scale_kernel");
}

int bruno_dense_coupling(const float* x, const float* s_pre, const float* t_pre, const float* ld_in, float* y, float* ld_out,
                         int N, int D, int mask_type, int inverse, void* stream) {
  BRUNO_REQUIRE(x && s_pre && t_pre && ld_in && y && ld_out && N > 0 && D > 0, "bruno_dense_coupling: bad arguments");
  BRUNO_REQUIRE(mask_type == BRUNO_MASK_EVEN || mask_type == BRUNO_MASK_ODD, "bruno_dense_coupling: bad mask");
  dense_coupling_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, s_pre, t_pre, ld_in, y, ld_out, D, mask_type, inverse);
  return check_launch("dense_coupling_kernel");
}

int bruno_dense_coupling_bwd(const float* x, const float* s_pre, const float* dy, const float* dld, float* dx, float* ds_pre,
                             float* dt_pre, int N, int D, int mask_type, void* stream) {
  BRUNO_REQUIRE(x && s_pre && dy && dld && dx && ds_pre && dt_pre && N > 0 && D > 0, "bruno_dense_coupling_bwd: bad arguments");
  BRUNO_REQUIRE(mask_type == BRUNO_MASK_EVEN || mask_type == BRUNO_MASK_ODD, "bruno_dense_coupling_bwd: bad mask");
  const size_t ND = static_cast<size_t>(N) * D;
  dense_coupling_bwd_kernel<<<nblk(ND), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, s_pre, dy, dld, dx, ds_pre, dt_pre, ND, D,
                                                                                      mask_type);
  return check_launch("dense_coupling_bwd_kernel");
}

int bruno_coupling_dense_fwd(const float* x, const float* ld_in, float* y, float* ld_out, int N, int D, int U, int mask_type,
                             int inverse, const float* V1, const float* g1, const float* b1, const float* V2,
                             const float* g2, const float* b2, const float* Ks, const float* bs, const float* Kt,
                             const float* bt, float* ws, void* stream) {
  BRUNO_REQUIRE(x && ld_in && y && ld_out && V1 && g1 && b1 && V2 && g2 && b2 && Ks && bs && Kt && bt && ws,
                "bruno_coupling_dense_fwd: null pointer");
  BRUNO_REQUIRE(N > 0 && D > 0 && U > 0 && (mask_type == BRUNO_MASK_EVEN || mask_type == BRUNO_MASK_ODD),
                "bruno_coupling_dense_fwd: bad sizes / mask");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DenseWs w = carve(ws, N, D, U);
  const size_t ND = static_cast<size_t>(N) * D;
  dense_wn_scale_kernel<<<ceil_div(U, 32), 1024, 0, st>>>(V1, g1, w.sc1, w.in1, D, U);
  dense_wn_scale_kernel<<<ceil_div(U, 32), 1024, 0, st>>>(V2, g2, w.sc2, w.in2, U, U);
  dense_mask_mul_kernel<<<nblk(ND), 256, 0, st>>>(x, w.xm, ND, D, mask_type);
  int rc;
  if ((rc = gemm(false, false, N, U, D, w.xm, D, V1, U, w.h1, U, GemmEpi{w.sc1, b1, w.pre1, 1, 0, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, U, U, w.h1, U, V2, U, w.h2, U, GemmEpi{w.sc2, b2, w.pre2, 1, 0, nullptr}, st))) return rc;
  // the s and t heads read the same h2: one launch (130 tiles: no split-K at all), else two
  rc = gemm_pair(1, false, false, N, D, U, w.h2, U, Ks, D, w.s_pre, D, GemmEpi{nullptr, bs, nullptr, 0, 0, nullptr}, w.h2, Kt, w.t_pre,
                 bt, st);
  if (rc < 0) return rc;
  if (rc == 0) {
    if ((rc = gemm(false, false, N, D, U, w.h2, U, Ks, D, w.s_pre, D, GemmEpi{nullptr, bs, nullptr, 0, 0, nullptr}, st))) return rc;
    if ((rc = gemm(false, false, N, D, U, w.h2, U, Kt, D, w.t_pre, D, GemmEpi{nullptr, bt, nullptr, 0, 0, nullptr}, st))) return rc;
  }
  dense_coupling_kernel<<<N, 256, 0, st>>>(x, w.s_pre, w.t_pre, ld_in, y, ld_out, D, mask_type, inverse);
  return check_launch("dense_coupling_kernel");
}

int bruno_coupling_dense_bwd(const float* x, const float* dy, const float* dld, int N, int D, int U, int mask_type,
                             const float* V1, const float* g1, const float* b1, const float* V2, const float* g2,
                             const float* b2, const float* Ks, const float* Kt, const float* ws_fwd, float* ws_bwd,
                             float* dx, float* dV1, float* dg1, float* db1, float* dV2, float* dg2, float* db2, float* dKs,
                             float* dbs, float* dKt, float* dbt, void* stream) {
  BRUNO_REQUIRE(x && dy && dld && V1 && g1 && b1 && V2 && g2 && b2 && Ks && Kt && ws_fwd && ws_bwd && dx && dV1 && dg1 &&
                    db1 && dV2 && dg2 && db2 && dKs && dbs && dKt && dbt, "bruno_coupling_dense_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DenseWs w = carve(const_cast<float*>(ws_fwd), N, D, U);
  const size_t ND = static_cast<size_t>(N) * D, NU = static_cast<size_t>(N) * U;
  size_t o = 0;
  auto take = [&](size_t n) { float* p = ws_bwd + o; o += (n + 3) / 4 * 4; return p; };
  float* ds_pre = take(ND);
  float* dt_pre = take(ND);
  float* dxm = take(ND);
  float* dh = take(NU);
  float* dpre = take(NU);
  float* r1 = take(U);
  int rc;
  const GemmEpi plain{nullptr, nullptr, nullptr, 0, 0, nullptr};
  const GemmEpi accum{nullptr, nullptr, nullptr, 0, 1, nullptr};
  // elementwise coupling backward + the head-bias gradients (column sums of ds_pre, dt_pre) in one pass
  coupling_bwd_col_reduce(x, w.s_pre, dy, dld, dx, ds_pre, dt_pre, dbs, dbt, N, D, mask_type, st);
  // heads
  // dKs and dKt (same h2^T) side by side in one launch; dh = ds Ks^T + dt Kt^T as one launch whose split-K slices come from
  // the two products; each falls back to two launches when the pair cannot run (gemm_pair)
  rc = gemm_pair(1, true, false, U, D, N, w.h2, U, ds_pre, D, dKs, D, plain, w.h2, dt_pre, dKt, nullptr, st);
  if (rc < 0) return rc;
  if (rc == 0) {
    if ((rc = gemm(true, false, U, D, N, w.h2, U, ds_pre, D, dKs, D, plain, st))) return rc;
    if ((rc = gemm(true, false, U, D, N, w.h2, U, dt_pre, D, dKt, D, plain, st))) return rc;
  }
  rc = gemm_pair(2, false, true, N, U, D, ds_pre, D, Ks, D, dh, U, plain, dt_pre, Kt, nullptr, nullptr, st);
  if (rc < 0) return rc;
  if (rc == 0) {
    if ((rc = gemm(false, true, N, U, D, ds_pre, D, Ks, D, dh, U, plain, st))) return rc;
    if ((rc = gemm(false, true, N, U, D, dt_pre, D, Kt, D, dh, U, accum, st))) return rc;
  }
  // d2
  elu_bwd_col_reduce(w.pre2, dh, dpre, db2, r1, N, U, st);          // dpre = dh ELU'(pre2), db2, r1 in one pass
  if ((rc = gemm(true, false, U, U, N, w.h1, U, dpre, U, dV2, U, plain, st))) return rc;
  dense_wn_bwd_kernel<<<nblk(static_cast<size_t>(U) * U), 256, 0, st>>>(V2, g2, b2, w.sc2, w.in2, r1, db2, dV2, dg2, U, U);
  {
    GemmEpi e = plain;
    e.a_kscale = w.sc2;
    if ((rc = gemm(false, true, N, U, U, dpre, U, V2, U, dh, U, e, st))) return rc;
  }
  // d1
  elu_bwd_col_reduce(w.pre1, dh, dpre, db1, r1, N, U, st);
  if ((rc = gemm(true, false, D, U, N, w.xm, D, dpre, U, dV1, U, plain, st))) return rc;
  dense_wn_bwd_kernel<<<nblk(static_cast<size_t>(D) * U), 256, 0, st>>>(V1, g1, b1, w.sc1, w.in1, r1, db1, dV1, dg1, D, U);
  {
    GemmEpi e = plain;
    e.a_kscale = w.sc1;
    if ((rc = gemm(false, true, N, D, U, dpre, U, V1, U, dxm, D, e, st))) return rc;
  }
  masked_accumulate_kernel<<<nblk(ND), 256, 0, st>>>(dx, dxm, ND, D, mask_type);
  return check_launch("bruno_coupling_dense_bwd");
}

size_t bruno_coupling_dense_cond_ws_floats(int N, int D, int U, int L) { return cond_dense_ws_floats(N, D, U, L); }
size_t bruno_coupling_dense_cond_bwd_ws_floats(int N, int D, int U, int L) {
  const size_t n = N, k1 = D / 2 + D, k2 = U / 2 + U;
  auto r4 = [](size_t v) { return (v + 3) / 4 * 4; };
  return 2 * r4(n * D) + r4(n * (k1 > k2 ? k1 : k2)) + 2 * r4(n * U) + 2 * r4(n * (D / 2 > U / 2 ? D / 2 : U / 2)) +
         r4(n * L) + r4(U > D / 2 ? U : D / 2) + 16;
}

int bruno_coupling_dense_cond_fwd(const float* x, const float* yl, const float* ld_in, float* y, float* ld_out, int N, int D,
                                  int U, int L, int mask_type, int inverse, const float* const* vars, float* ws, void* stream) {
  BRUNO_REQUIRE(x && yl && ld_in && y && ld_out && vars && ws, "bruno_coupling_dense_cond_fwd: null pointer");
  BRUNO_REQUIRE(N > 0 && D > 0 && U > 0 && L > 0 && D % 2 == 0 && U % 2 == 0 &&
                    (mask_type == BRUNO_MASK_EVEN || mask_type == BRUNO_MASK_ODD), "bruno_coupling_dense_cond_fwd: bad sizes / mask");
  for (int i = 0; i < cNVAR; ++i) BRUNO_REQUIRE(vars[i] != nullptr, "bruno_coupling_dense_cond_fwd: variable %d is null", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nh1 = D / 2, nh2 = U / 2, K1 = nh1 + D, K2 = nh2 + U;
  const size_t n = N;
  CondDenseWs w = cond_carve(ws, N, D, U);
  dense_wn_scale_kernel<<<ceil_div(nh1, 32), 1024, 0, st>>>(vars[cLV1], vars[cLg1], w.scL1, w.inL1, L, nh1);
  dense_wn_scale_kernel<<<ceil_div(nh2, 32), 1024, 0, st>>>(vars[cLV2], vars[cLg2], w.scL2, w.inL2, L, nh2);
  dense_wn_scale_kernel<<<ceil_div(U, 32), 1024, 0, st>>>(vars[cV1], vars[cg1], w.sc1, w.in1, K1, U);
  dense_wn_scale_kernel<<<ceil_div(U, 32), 1024, 0, st>>>(vars[cV2], vars[cg2], w.sc2, w.in2, K2, U);
  int rc;
  // d1 = dense_wn(concat(leaky(dense_wn(y_label)), x * b))   (cond:461-471)
  if ((rc = gemm(false, false, N, nh1, L, yl, L, vars[cLV1], nh1, w.hL1, nh1, GemmEpi{w.scL1, vars[cLb1], w.preL1, 2, 0, nullptr}, st))) return rc;
  copy_cols_kernel<<<nblk(n * nh1), 256, 0, st>>>(w.hL1, w.o1, n * nh1, nh1, K1, 0, -1);
  copy_cols_kernel<<<nblk(n * D), 256, 0, st>>>(x, w.o1, n * D, D, K1, nh1, mask_type);
  if ((rc = gemm(false, false, N, U, K1, w.o1, K1, vars[cV1], U, w.h1, U, GemmEpi{w.sc1, vars[cb1], w.pre1, 1, 0, nullptr}, st))) return rc;
  // d2
  if ((rc = gemm(false, false, N, nh2, L, yl, L, vars[cLV2], nh2, w.hL2, nh2, GemmEpi{w.scL2, vars[cLb2], w.preL2, 2, 0, nullptr}, st))) return rc;
  copy_cols_kernel<<<nblk(n * nh2), 256, 0, st>>>(w.hL2, w.o2, n * nh2, nh2, K2, 0, -1);
  copy_cols_kernel<<<nblk(n * U), 256, 0, st>>>(w.h1, w.o2, n * U, U, K2, nh2, -1);
  if ((rc = gemm(false, false, N, U, K2, w.o2, K2, vars[cV2], U, w.h2, U, GemmEpi{w.sc2, vars[cb2], w.pre2, 1, 0, nullptr}, st))) return rc;
  // heads: plain dense(concat(leaky(dense(y_label)), h2))   (cond:350-369) = hs Ks[:nh2] + h2 Ks[nh2:] + bs
  if ((rc = gemm(false, false, N, nh2, L, yl, L, vars[cWs], nh2, w.hs, nh2, GemmEpi{nullptr, vars[cbws], nullptr, 2, 0, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, nh2, L, yl, L, vars[cWt], nh2, w.ht, nh2, GemmEpi{nullptr, vars[cbwt], nullptr, 2, 0, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, D, nh2, w.hs, nh2, vars[cKs], D, w.s_pre, D, GemmEpi{nullptr, vars[cbs], nullptr, 0, 0, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, D, U, w.h2, U, vars[cKs] + static_cast<size_t>(nh2) * D, D, w.s_pre, D, GemmEpi{nullptr, nullptr, nullptr, 0, 1, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, D, nh2, w.ht, nh2, vars[cKt], D, w.t_pre, D, GemmEpi{nullptr, vars[cbt], nullptr, 0, 0, nullptr}, st))) return rc;
  if ((rc = gemm(false, false, N, D, U, w.h2, U, vars[cKt] + static_cast<size_t>(nh2) * D, D, w.t_pre, D, GemmEpi{nullptr, nullptr, nullptr, 0, 1, nullptr}, st))) return rc;
  dense_coupling_kernel<<<N, 256, 0, st>>>(x, w.s_pre, w.t_pre, ld_in, y, ld_out, D, mask_type, inverse);
  return check_launch("bruno_coupling_dense_cond_fwd");
}

int bruno_coupling_dense_cond_bwd(const float* x, const float* yl, const float* dy, const float* dld, int N, int D, int U, int L,
                                  int mask_type, const float* const* vars, const float* ws_fwd, float* ws_bwd, float* dx,
                                  float* dyl, float* const* grads, void* stream) {
  BRUNO_REQUIRE(x && yl && dy && dld && vars && ws_fwd && ws_bwd && dx && dyl && grads, "bruno_coupling_dense_cond_bwd: null pointer");
  for (int i = 0; i < cNVAR; ++i) BRUNO_REQUIRE(vars[i] && grads[i], "bruno_coupling_dense_cond_bwd: variable / gradient %d is null", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nh1 = D / 2, nh2 = U / 2, K1 = nh1 + D, K2 = nh2 + U;
  const size_t n = N, ND = n * D, NU = n * U;
  CondDenseWs w = cond_carve(const_cast<float*>(ws_fwd), N, D, U);
  size_t o = 0;
  auto take = [&](size_t v) { float* p = ws_bwd + o; o += (v + 3) / 4 * 4; return p; };
  float* ds_pre = take(ND);
  float* dt_pre = take(ND);
  float* dob = take(n * (K1 > K2 ? K1 : K2));      // gradient of a concat buffer
  float* dh = take(NU);
  float* dpre = take(NU);
  float* dhl = take(n * (nh1 > nh2 ? nh1 : nh2));  // gradient of a label branch output
  float* dprel = take(n * (nh1 > nh2 ? nh1 : nh2));
  float* dyl_t = take(n * L);
  float* r1 = take(U > nh1 ? U : nh1);
  int rc;
  const GemmEpi plain{nullptr, nullptr, nullptr, 0, 0, nullptr};
  const GemmEpi accum{nullptr, nullptr, nullptr, 0, 1, nullptr};
  bool dyl_set = false;
  // dyl (+)= dpre_label @ W^T (optionally with the weight-norm column scale folded into the rows of dpre_label)
  auto label_back = [&](const float* dpl, const float* W, int ncol, const float* kscale) -> int {
    GemmEpi e = dyl_set ? accum : plain;
    e.a_kscale = kscale;
    dyl_set = true;
    return gemm(false, true, N, L, ncol, dpl, ncol, W, ncol, dyl, L, e, st);
  };
  (void)dyl_t;
  dense_coupling_bwd_kernel<<<nblk(ND), 256, 0, st>>>(x, w.s_pre, dy, dld, dx, ds_pre, dt_pre, ND, D, mask_type);
  // ---- heads
  struct Head { const float* dpre_out; int W, bw, K, b; const float* hl; };
  const Head heads[2] = {{ds_pre, cWs, cbws, cKs, cbs, w.hs}, {dt_pre, cWt, cbwt, cKt, cbt, w.ht}};
  for (int hidx = 0; hidx < 2; ++hidx) {
    const Head& H = heads[hidx];
    if ((rc = gemm(true, false, nh2, D, N, H.hl, nh2, H.dpre_out, D, grads[H.K], D, plain, st))) return rc;
    if ((rc = gemm(true, false, U, D, N, w.h2, U, H.dpre_out, D, grads[H.K] + static_cast<size_t>(nh2) * D, D, plain, st))) return rc;
    col_reduce(H.dpre_out, nullptr, grads[H.b], nullptr, N, D, st);
    if ((rc = gemm(false, true, N, U, D, H.dpre_out, D, vars[H.K] + static_cast<size_t>(nh2) * D, D, dh, U, hidx ? accum : plain, st))) return rc;
    // label branch of the head: plain dense + leaky_relu
    if ((rc = gemm(false, true, N, nh2, D, H.dpre_out, D, vars[H.K], D, dhl, nh2, plain, st))) return rc;
    leaky_bwd_kernel<<<nblk(n * nh2), 256, 0, st>>>(H.hl, dhl, dprel, n * nh2);
    if ((rc = gemm(true, false, L, nh2, N, yl, L, dprel, nh2, grads[H.W], nh2, plain, st))) return rc;
    col_reduce(dprel, nullptr, grads[H.bw], nullptr, N, nh2, st);
    if ((rc = label_back(dprel, vars[H.W], nh2, nullptr))) return rc;
  }
  // ---- d2 then d1: dense_wn(concat(label branch, input))
  struct Lay { int LV, Lg, Lb, V, g, b, nh, K, din; const float *scL, *inL, *sc, *in, *hL, *preL, *o, *pre; };
  const Lay lays[2] = {{cLV2, cLg2, cLb2, cV2, cg2, cb2, nh2, K2, U, w.scL2, w.inL2, w.sc2, w.in2, w.hL2, w.preL2, w.o2, w.pre2},
                       {cLV1, cLg1, cLb1, cV1, cg1, cb1, nh1, K1, D, w.scL1, w.inL1, w.sc1, w.in1, w.hL1, w.preL1, w.o1, w.pre1}};
  for (int li = 0; li < 2; ++li) {
    const Lay& Y = lays[li];
    elu_bwd_kernel<<<nblk(NU), 256, 0, st>>>(Y.pre, dh, dpre, NU);
    col_reduce(dpre, Y.pre, grads[Y.b], r1, N, U, st);
    if ((rc = gemm(true, false, Y.K, U, N, Y.o, Y.K, dpre, U, grads[Y.V], U, plain, st))) return rc;
    dense_wn_bwd_kernel<<<nblk(static_cast<size_t>(Y.K) * U), 256, 0, st>>>(vars[Y.V], vars[Y.g], vars[Y.b], Y.sc, Y.in, r1, grads[Y.b],
                                                                          grads[Y.V], grads[Y.g], Y.K, U);
    {
      GemmEpi e = plain;
      e.a_kscale = Y.sc;
      if ((rc = gemm(false, true, N, Y.K, U, dpre, U, vars[Y.V], U, dob, Y.K, e, st))) return rc;
    }
    // label branch: dense_wn + leaky_relu
    slice_cols_kernel<<<nblk(n * Y.nh), 256, 0, st>>>(dob, dhl, n * Y.nh, Y.nh, Y.K, 0);
    leaky_bwd_kernel<<<nblk(n * Y.nh), 256, 0, st>>>(Y.hL, dhl, dprel, n * Y.nh);
    col_reduce(dprel, Y.preL, grads[Y.Lb], r1, N, Y.nh, st);
    if ((rc = gemm(true, false, L, Y.nh, N, yl, L, dprel, Y.nh, grads[Y.LV], Y.nh, plain, st))) return rc;
    dense_wn_bwd_kernel<<<nblk(static_cast<size_t>(L) * Y.nh), 256, 0, st>>>(vars[Y.LV], vars[Y.Lg], vars[Y.Lb], Y.scL, Y.inL, r1,
                                                                            grads[Y.Lb], grads[Y.LV], grads[Y.Lg], L, Y.nh);
    if ((rc = label_back(dprel, vars[Y.LV], Y.nh, Y.scL))) return rc;
    // input branch
    if (li == 0) {
      slice_cols_kernel<<<nblk(NU), 256, 0, st>>>(dob, dh, NU, U, Y.K, Y.nh);           // gradient of h1
    } else {
      slice_cols_kernel<<<nblk(ND), 256, 0, st>>>(dob, ds_pre, ND, D, Y.K, Y.nh);        // gradient of x * b (ds_pre is free now)
      masked_accumulate_kernel<<<nblk(ND), 256, 0, st>>>(dx, ds_pre, ND, D, mask_type);
    }
  }
  return check_launch("bruno_coupling_dense_cond_bwd");
}

// ---- one dense layer, forward + hand-written backward (label paths of the conditional model) --------------------------
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a ? a[i] : 0.f) + (b ? b[i] : 0.f);
}

size_t bruno_dense_layer_bwd_ws_floats(int N, int U) { return static_cast<size_t>(N) * U + U + 8; }

int bruno_dense_layer_fwd(const float* x, const float* W, const float* g, const float* b, const float* extra_bias, int N, int K,
                          int U, int act, float* out, float* pre, float* sc, float* inv_nrm, float* btot, void* stream) {
  BRUNO_REQUIRE(x && W && out && btot && N > 0 && K > 0 && U > 0 && act >= 0 && act <= 3, "bruno_dense_layer_fwd: bad arguments");
  BRUNO_REQUIRE(!g || (sc && inv_nrm), "bruno_dense_layer_fwd: weight-norm needs the sc / inv_nrm outputs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  add_vec_kernel<<<ceil_div(U, 256), 256, 0, st>>>(b, extra_bias, btot, U);
  if (g) dense_wn_scale_kernel<<<ceil_div(U, 32), 1024, 0, st>>>(W, g, sc, inv_nrm, K, U);
  return gemm(false, false, N, U, K, x, K, W, U, out, U, GemmEpi{g ? sc : nullptr, btot, pre, act, 0, nullptr}, st);
}

int bruno_dense_layer_bwd(const float* x, const float* W, const float* g, const float* btot, const float* sc,
                          const float* inv_nrm, const float* pre, const float* out, const float* dout, int N, int K, int U,
                          int act, float* ws, float* dx, float* dW, float* dg, float* db, void* stream) {
  BRUNO_REQUIRE(x && W && dout && ws && dW && db && N > 0 && K > 0 && U > 0, "bruno_dense_layer_bwd: bad arguments");
  BRUNO_REQUIRE(!g || (btot && sc && inv_nrm && pre && dg), "bruno_dense_layer_bwd: weight-norm needs btot / sc / inv_nrm / pre / dg");
  BRUNO_REQUIRE(act == 0 || (act == 1 && pre) || (act == 2 && out), "bruno_dense_layer_bwd: act 1 needs pre, act 2 needs out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t NU = static_cast<size_t>(N) * U;
  float* dpre_buf = ws;
  float* r1 = ws + (NU + 3) / 4 * 4;
  const float* dpre = dout;
  if (act == 1) {
    elu_bwd_kernel<<<nblk(NU), 256, 0, st>>>(pre, dout, dpre_buf, NU);
    dpre = dpre_buf;
  } else if (act == 2) {
    leaky_bwd_kernel<<<nblk(NU), 256, 0, st>>>(out, dout, dpre_buf, NU);
    dpre = dpre_buf;
  }
  col_reduce(dpre, g ? pre : nullptr, db, g ? r1 : nullptr, N, U, st);
  int rc;
  const GemmEpi plain{nullptr, nullptr, nullptr, 0, 0, nullptr};
  if ((rc = gemm(true, false, K, U, N, x, K, dpre, U, dW, U, plain, st))) return rc;
  if (g) dense_wn_bwd_kernel<<<nblk(static_cast<size_t>(K) * U), 256, 0, st>>>(W, g, btot, sc, inv_nrm, r1, db, dW, dg, K, U);
  if (dx) {
    GemmEpi e = plain;
    e.a_kscale = g ? sc : nullptr;
    if ((rc = gemm(false, true, N, K, U, dpre, U, W, U, dx, K, e, st))) return rc;
  }
  return check_launch("bruno_dense_layer_bwd");
}

}  // extern "C"
