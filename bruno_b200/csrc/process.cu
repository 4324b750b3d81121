// Exchangeable Student-t / Gaussian process kernels (replaces nn_extra_student.py / nn_extra_gauss.py of the
// reference, whose every method expands into ~45 tiny TF ops per sequence step).
//
// Design (HBM-bound scan):
//   * The reference recursion (student:85-106) has O(1) state, and unrolled it is a function of the prefix sums
//     S1 = sum(x - mu0), S2 = sum((x - mu0)^2) only (utils_stats.py:87-95 states the same closed form).  Everything
//     else -- dd_n, k_n, b_n, nu_n, the lgamma quotient -- depends on (step, dim) but not on the batch element, so a
//     tiny fp64 "table" kernel evaluates it once per (step, dim) and the scan kernel reads it through L1/L2.
//   * scan kernel: one thread = 4 consecutive latent dims (128-bit loads along D) x NB batch rows; the sequence is
//     walked in registers; per step the per-row partial sums are reduced over D with a value-halving butterfly of
//     warp shuffles, then across warps through shared memory in a fixed order (bit-reproducible).
//   * hand-written backward: pass 1 re-reads z to form the total sums, pass 2 walks the sequence in reverse,
//     carrying the suffix sums of dL/dS1, dL/dS2; coefficient gradients are chained to the raw variables with the
//     fp64 forward-mode derivatives stored by the table kernel.
//   * default kernels (process_fwd_ring_kernel / process_bwd_ring_kernel, D % 4 == 0 and aligned pointers): persistent
//     blocks, a producer warp streaming z -- and the step's table rows -- through a cp.async.bulk + mbarrier
//     shared-memory ring, the same thread <-> (4 dims x NB rows) mapping in the consumer warps; the forward carries
//     q_n = 1 + beta_n / (nu0 - 2) recursively (q_{n+1} = q_n + r_n^2: one lg2 per element).  The register-staged
//     kernels below them in this file are the fallback for other shapes and the A/B partners
//     (bruno_process_set_fwd_variant / _bwd_variant).  Measured: forward 97 % of the HBM copy peak (was 44.5 %).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include <mutex>
#include "ptx.cuh"

namespace bruno {

// ------------------------------------------------------------------------------------------------------------
// forward-mode dual numbers in fp64 (3 tangents: raw var, raw nu, raw corr)
// ------------------------------------------------------------------------------------------------------------
struct D3 {
  double v, d[3];
};
__device__ inline D3 mk(double v) { return D3{v, {0.0, 0.0, 0.0}}; }
__device__ inline D3 seed(double v, int i) {
  D3 r = mk(v);
  r.d[i] = 1.0;
  return r;
}
__device__ inline D3 operator+(D3 a, D3 b) { return D3{a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
__device__ inline D3 operator-(D3 a, D3 b) { return D3{a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
__device__ inline D3 operator*(D3 a, D3 b) {
  return D3{a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
__device__ inline D3 operator/(D3 a, D3 b) {
  double q = a.v / b.v, ib = 1.0 / b.v;
  return D3{q, {(a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib}};
}
__device__ inline D3 operator+(D3 a, double s) { a.v += s; return a; }
__device__ inline D3 operator-(D3 a, double s) { a.v -= s; return a; }
__device__ inline D3 operator*(D3 a, double s) { return D3{a.v * s, {a.d[0] * s, a.d[1] * s, a.d[2] * s}}; }
__device__ inline D3 chain(D3 a, double f, double df) { return D3{f, {a.d[0] * df, a.d[1] * df, a.d[2] * df}}; }
__device__ inline D3 dlog(D3 a) { return chain(a, log(a.v), 1.0 / a.v); }
__device__ inline D3 dsqrt(D3 a) {
  double s = sqrt(a.v);
  return chain(a, s, 0.5 / s);
}
__device__ inline D3 dexp(D3 a) {
  double e = exp(a.v);
  return chain(a, e, e);
}
__device__ inline double digamma(double x) {
  double r = 0.0;
  while (x < 8.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0)))));
  return r + log(x) - 0.5 / x + t;
}
__device__ inline D3 dlgamma(D3 a) { return chain(a, lgamma(a.v), digamma(a.v)); }
__device__ inline D3 dsoftplus(D3 a) {  // tf.nn.softplus
  double x = a.v;
  double sp = x > 30.0 ? x : log1p(exp(x));
  double sg = 1.0 / (1.0 + exp(-x));
  return chain(a, sp, sg);
}
__device__ inline D3 dsigmoid(D3 a) {
  double sg = 1.0 / (1.0 + exp(-a.v));
  return chain(a, sg, sg * (1.0 - sg));
}

constexpr int TP_NCF = 7;  // w, ddw, e, f, C, A2, B2
constexpr int GP_NCF = 3;  // dd, I, C
constexpr int N_PRIOR = 3;
__host__ __device__ constexpr int ncf_of(int kind) { return kind == BRUNO_TP ? TP_NCF : GP_NCF; }

#define LN2_D 0.69314718055994530942
#define LNPI_D 1.14472988584940017414

// One thread per (row-group n in [0, T], dim d); n == T builds the prior constants.
__global__ void process_table_kernel(int kind, const float* __restrict__ var_vbl, const float* __restrict__ nu_vbl,
                                     const float* __restrict__ corr_vbl, const float* __restrict__ mask_dim,
                                     float min_nu, int n0, int T, int D, int with_grad, float* __restrict__ table) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;  // 0..T
  if (d >= D) return;
  const int NCF = ncf_of(kind);
  const bool is_prior = row == T;
  const double n = is_prior ? 0.0 : static_cast<double>(n0 + row);
  const double m = mask_dim ? static_cast<double>(mask_dim[d]) : 1.0;

  D3 sp = dsoftplus(seed(static_cast<double>(var_vbl[d]), 0));
  D3 v = sp * sp;                                                       // student:38
  D3 rho = dsigmoid(seed(static_cast<double>(corr_vbl[d]), 2));         // student:60
  D3 c = rho * v;                                                       // student:61
  D3 dd = c / (v + c * (n - 1.0));                                      // student:93
  D3 k = v - dd * c * n;                                                // closed form of student:101
  D3 out[TP_NCF];
  int nout;
  if (kind == BRUNO_TP) {
    D3 nu0 = dexp(seed(static_cast<double>(nu_vbl[d]), 1)) + static_cast<double>(min_nu);  // student:46
    D3 c0 = k * (nu0 - 2.0);          // (nu_n - 2) * sigma_n at beta = 0
    D3 s = mk(1.0) / c0;
    D3 w = dsqrt(s);
    D3 iv = mk(1.0) / (v - c);        // a_i - b_i of student:97-98, constant in i
    D3 b = (mk(0.0) - c) / ((v - c) * (c * (n - 1.0) + v));             // student:98
    D3 A = (nu0 + (n + 1.0)) * 0.5;   // (1 + nu_n) / 2
    D3 C = dlgamma(A) - dlgamma((nu0 + n) * 0.5) - dlog(c0) * 0.5 - 0.5 * LNPI_D;
    if (!is_prior) {
      out[0] = w;
      out[1] = dd * w;
      out[2] = iv / (nu0 - 2.0);
      out[3] = b / (nu0 - 2.0);
      out[4] = C * m;
      out[5] = A * (LN2_D * m);
      out[6] = (A - 0.5) * (LN2_D * m);
      nout = TP_NCF;
    } else {
      out[0] = s;
      out[1] = C * m;
      out[2] = A * (LN2_D * m);
      nout = N_PRIOR;
    }
  } else {
    // gauss:75-96: sigma_n follows the same recursion as k_n.
    D3 I = mk(0.5) / k;
    D3 C = (dlog(k) + log(2.0 * 3.14159265358979323846)) * (-0.5);
    if (!is_prior) {
      out[0] = dd;
      out[1] = I * m;
      out[2] = C * m;
      nout = GP_NCF;
    } else {
      out[0] = I * m;
      out[1] = C * m;
      out[2] = mk(0.0);
      nout = N_PRIOR;
    }
  }
  const int row0 = is_prior ? T * NCF : row * NCF;
  const size_t gbase = static_cast<size_t>(T * NCF + N_PRIOR) * D;
  for (int i = 0; i < nout; ++i) {
    table[static_cast<size_t>(row0 + i) * D + d] = static_cast<float>(out[i].v);
    if (with_grad)
      for (int p = 0; p < 3; ++p)
        table[gbase + (static_cast<size_t>(row0 + i) * 3 + p) * D + d] = static_cast<float>(out[i].d[p]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
struct Vf {
  float x[VEC];
};
template <int VEC>
__device__ __forceinline__ Vf<VEC> ldv(const float* p) {
  Vf<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.x[0] = t.x; r.x[1] = t.y;
  } else {
    r.x[0] = __ldg(p);
  }
  return r;
}
// streaming (read-once) variant for z
template <int VEC>
__device__ __forceinline__ Vf<VEC> ldv_stream(const float* p) {
  Vf<VEC> r;
  if constexpr (VEC == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x[0]), "=f"(r.x[1]), "=f"(r.x[2]), "=f"(r.x[3])
                 : "l"(p));
  } else if constexpr (VEC == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x[0]), "=f"(r.x[1]) : "l"(p));
  } else {
    r.x[0] = __ldg(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Vf<VEC>& v) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v.x[0], v.x[1], v.x[2], v.x[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v.x[0], v.x[1]);
  } else {
    p[0] = v.x[0];
  }
}
template <int VEC>
__device__ __forceinline__ Vf<VEC> zerov() {
  Vf<VEC> r;
#pragma unroll
  for (int i = 0; i < VEC; ++i) r.x[i] = 0.f;
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Reduce NV (power of two <= 32) values per lane across the warp with NV-1 + log2(32/NV) shuffles.
// On return lane L holds in v[0] the warp total of value index (L >> (5 - log2 NV)) ... see slot_of_lane.
template <int NV>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[NV], int lane) {
  int cnt = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      cnt >>= 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < cnt) {
          float send = upper ? v[i] : v[i + cnt];
          float keep = upper ? v[i + cnt] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}
template <int NV>
__device__ __forceinline__ int slot_of_lane(int lane) {
  // value index held by this lane after warp_multi_reduce: built from the lane bits used while halving
  int idx = 0, cnt = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      cnt >>= 1;
      if (lane & off) idx += cnt;
    }
  }
  return idx;
}
template <int NV>
__device__ __forceinline__ bool lane_is_writer(int lane) {
  // lanes whose un-used (fully reduced) bits are zero
  int used = 0, cnt = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      cnt >>= 1;
      used |= off;
    }
  }
  return (lane & ~used & 31) == 0;
}

// ------------------------------------------------------------------------------------------------------------
// forward scan
// ------------------------------------------------------------------------------------------------------------
template <int KIND, int VEC, int NB, bool PRIOR, int PF, int MINB, int RS, int MAXT>
__global__ void __launch_bounds__(MAXT, MINB)
process_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mu0, const float* __restrict__ tab,
                   float* __restrict__ s1, float* __restrict__ s2, float* __restrict__ lp,
                   float* __restrict__ lp_prior, float* __restrict__ scratch, int B, int T, int D) {
  constexpr int NCF = ncf_of(KIND);
  constexpr int NV = NB * 2;
  constexpr int NVAL = NV * RS;    // RS steps' partial sums are reduced together: ONE NVAL-value butterfly per RS steps
  static_assert(NVAL <= 32, "at most 32 values per butterfly");
  (void)PF;
  extern __shared__ float sred[];  // [T][nwarps][NV]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int d0 = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
  const bool active = d0 < D;
  const int b0 = blockIdx.x * NB;
  const size_t rowTD = static_cast<size_t>(T) * D;

  Vf<VEC> S1[NB], S2[NB], mu = zerov<VEC>(), pc[N_PRIOR];
  Vf<VEC> c_e = zerov<VEC>(), c_a0 = zerov<VEC>(), c_da = zerov<VEC>();   // TP: coefficients that need no per-step load
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    S1[j] = zerov<VEC>();
    S2[j] = zerov<VEC>();
  }
#pragma unroll
  for (int i = 0; i < N_PRIOR; ++i) pc[i] = zerov<VEC>();
  if (active) {
    mu = ldv<VEC>(mu0 + d0);
    if (PRIOR) {
#pragma unroll
      for (int i = 0; i < N_PRIOR; ++i) pc[i] = ldv<VEC>(tab + static_cast<size_t>(T * NCF + i) * D + d0);
    }
    if (s1 != nullptr) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (b0 + j < B) {
          S1[j] = ldv<VEC>(s1 + static_cast<size_t>(b0 + j) * D + d0);
          if (KIND == BRUNO_TP) S2[j] = ldv<VEC>(s2 + static_cast<size_t>(b0 + j) * D + d0);
        }
    }
  }
  // Hot loop without guards: out-of-range lanes / rows read CLAMPED (valid) addresses and are masked once per reduce
  // batch; rows >= B are simply not written out.  Running pointers instead of per-load 64-bit index arithmetic.
  // (ncu, first version: 23 instructions per element at 53 % issue utilisation = 42 % of HBM peak; the arithmetic is
  // 13 FP + 2 MUFU per element, the rest was address math, guards and register moves.)
  const int d0c = active ? d0 : 0;
  const float lane_mask = active ? 1.f : 0.f;
  if (KIND == BRUNO_TP) {
    // e' = 1 / ((v - c)(nu0 - 2)) does not depend on the step; A2_n = A2_0 + n dA, B2_n = A2_n - dA (dA = m ln2 / 2).
    // Inactive lanes take the (consistent) coefficient set of dim 0: their values stay finite and are masked out.
    c_e = ldv<VEC>(tab + static_cast<size_t>(2) * D + d0c);
    c_a0 = ldv<VEC>(tab + static_cast<size_t>(5) * D + d0c);
    const Vf<VEC> b0v = ldv<VEC>(tab + static_cast<size_t>(6) * D + d0c);
#pragma unroll
    for (int e = 0; e < VEC; ++e) c_da.x[e] = c_a0.x[e] - b0v.x[e];
  }
  if (PRIOR && !active) {
#pragma unroll
    for (int i = 0; i < N_PRIOR; ++i) pc[i] = ldv<VEC>(tab + static_cast<size_t>(T * NCF + i) * D + d0c);
  }
  const float* zp[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) zp[j] = z + static_cast<size_t>(min(b0 + j, B - 1)) * rowTD + d0c;
  const float* tp = tab + d0c;
  const size_t tstep = static_cast<size_t>(NCF) * D;
  const bool mu_zero = __all_sync(0xffffffffu, [&] { bool zr = true;
#pragma unroll
    for (int e = 0; e < VEC; ++e) zr = zr && (mu.x[e] == 0.f);
    return zr; }());

  auto do_step = [&](int n, float* vals_pp) {
    Vf<VEC> x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      x[j] = ldv_stream<VEC>(zp[j]);
      zp[j] += D;
    }
    Vf<VEC> cw, cdd, cf3, cC;   // TP: w, dd w, f', C;   GP: dd, I, C in cw, cdd, cC
    if (KIND == BRUNO_TP) {
      cw = ldv<VEC>(tp);
      cdd = ldv<VEC>(tp + D);
      cf3 = ldv<VEC>(tp + 3 * static_cast<size_t>(D));
      cC = ldv<VEC>(tp + 4 * static_cast<size_t>(D));
    } else {
      cw = ldv<VEC>(tp);
      cdd = ldv<VEC>(tp + D);
      cC = ldv<VEC>(tp + 2 * static_cast<size_t>(D));
      cf3 = zerov<VEC>();
    }
    tp += tstep;
    const float fn = static_cast<float>(n);
    Vf<VEC> a2, b2;
    if (KIND == BRUNO_TP) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        a2.x[e] = fmaf(fn, c_da.x[e], c_a0.x[e]);
        b2.x[e] = a2.x[e] - c_da.x[e];
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float acc = 0.f, accp = 0.f;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float xz = mu_zero ? x[j].x[e] : x[j].x[e] - mu.x[e];
        if constexpr (KIND == BRUNO_TP) {
          const float s1v = S1[j].x[e];
          const float r = fmaf(-cdd.x[e], s1v, cw.x[e] * xz);
          const float q = fmaf(cf3.x[e], s1v * s1v, fmaf(c_e.x[e], S2[j].x[e], 1.f));
          const float u = fmaf(r, r, q);
          acc += fmaf(b2.x[e], fast_lg2(q), fmaf(-a2.x[e], fast_lg2(u), cC.x[e]));
          if (PRIOR) accp += fmaf(-pc[2].x[e], fast_lg2(fmaf(xz * xz, pc[0].x[e], 1.f)), pc[1].x[e]);
          S1[j].x[e] = s1v + xz;
          S2[j].x[e] = fmaf(xz, xz, S2[j].x[e]);
        } else {
          const float r = fmaf(-cw.x[e], S1[j].x[e], xz);
          acc += fmaf(-cdd.x[e], r * r, cC.x[e]);
          if (PRIOR) accp += fmaf(-pc[0].x[e], xz * xz, pc[1].x[e]);
          S1[j].x[e] += xz;
        }
      }
      vals_pp[2 * j] = acc;
      vals_pp[2 * j + 1] = accp;
    }
  };

  int n0 = 0;
  for (; n0 + RS <= T; n0 += RS) {
    float vals[NVAL];
#pragma unroll
    for (int pp = 0; pp < RS; ++pp) do_step(n0 + pp, vals + pp * NV);
#pragma unroll
    for (int i = 0; i < NVAL; ++i) vals[i] *= lane_mask;
    warp_multi_reduce<NVAL>(vals, lane);
    if (lane_is_writer<NVAL>(lane)) {
      const int slot = slot_of_lane<NVAL>(lane);   // = pp * NV + idx
      sred[((n0 + slot / NV) * nwarps + warp) * NV + (slot % NV)] = vals[0];
    }
  }
  for (; n0 < T; ++n0) {                         // tail: fewer than RS steps left, one small butterfly per step
    float vals[NV];
    do_step(n0, vals);
#pragma unroll
    for (int i = 0; i < NV; ++i) vals[i] *= lane_mask;
    warp_multi_reduce<NV>(vals, lane);
    if (lane_is_writer<NV>(lane)) sred[(n0 * nwarps + warp) * NV + slot_of_lane<NV>(lane)] = vals[0];
  }
  __syncthreads();
  const int nseg = gridDim.y;
  for (int i = threadIdx.x; i < T * NV; i += blockDim.x) {
    const int n = i / NV, idx = i % NV, j = idx >> 1, which = idx & 1;
    if (b0 + j >= B) continue;
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += sred[(n * nwarps + w) * NV + idx];
    if (nseg == 1) {
      if (which == 0) lp[static_cast<size_t>(b0 + j) * T + n] = s;
      else if (PRIOR) lp_prior[static_cast<size_t>(b0 + j) * T + n] = s;
    } else {
      scratch[((static_cast<size_t>(blockIdx.y) * B + b0 + j) * T + n) * 2 + which] = s;
    }
  }
  if (s1 != nullptr && active) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (b0 + j < B) {
        stv<VEC>(s1 + static_cast<size_t>(b0 + j) * D + d0, S1[j]);
        if (KIND == BRUNO_TP) stv<VEC>(s2 + static_cast<size_t>(b0 + j) * D + d0, S2[j]);
      }
  }
}

__global__ void process_fwd_finalize(const float* __restrict__ scratch, float* __restrict__ lp,
                                     float* __restrict__ lp_prior, int nseg, int BT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT) return;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < nseg; ++s) {
    a += scratch[(static_cast<size_t>(s) * BT + i) * 2];
    b += scratch[(static_cast<size_t>(s) * BT + i) * 2 + 1];
  }
  lp[i] = a;
  if (lp_prior) lp_prior[i] = b;
}

// ------------------------------------------------------------------------------------------------------------
// forward scan, persistent + bulk-copy ring (the default when D % 4 == 0 and the pointers are 16-byte aligned)
// ------------------------------------------------------------------------------------------------------------
// The register-staged kernel above is latency-bound (ncu: 38 % issue utilisation, long-scoreboard the top stall, 14
// warps per SM at 128 registers, and a block lives for only T / RS = 5 loop iterations, so its prologue -- dependent
// L2 loads of the coefficients -- and its epilogue -- block barrier, cross-warp reduce -- leave the memory pipe idle for
// a sizeable share of its life).  This variant keeps HBM busy independently of what the math warps are doing:
//   * persistent blocks (grid.x = resident blocks), each walking row groups g = blockIdx.x, += gridDim.x;
//   * one producer warp: an elected lane streams z[rows of the group][step][segment] with 1-D bulk copies (UBLKCP; a
//     row's step is one contiguous D*4-byte run) into an NSTAGE-deep shared-memory ring, full/empty mbarriers per
//     stage, running ahead across group boundaries so the next group's first steps land while this group reduces;
//   * consumer warps: same thread <-> (4 dims x NB rows) mapping, same butterfly and same cross-warp order as the
//     kernel above, z from LDS.128; the three per-step coefficient rows either ride in the same ring stage (CPF = 2,
//     default: three more bulk copies per stage from the L2-resident table, no LDG in the hot loop) or are prefetched
//     one step ahead with LDG (CPF = 1) / loaded in place (CPF = 0); the [T][warps][NV] partials are double-buffered
//     by group parity so one consumer-only named barrier per group suffices.
template <int KIND, int NB, bool PRIOR, int NSTAGE, int RS, int MINB, int CPF, bool STATE>
__global__ void __launch_bounds__(256, MINB)
process_fwd_ring_kernel(const float* __restrict__ z, const float* __restrict__ mu0, const float* __restrict__ tab,
                        float* __restrict__ s1, float* __restrict__ s2, float* __restrict__ lp,
                        float* __restrict__ lp_prior, float* __restrict__ scratch, int B, int T, int D) {
  constexpr int VEC = 4;
  constexpr int NCF = ncf_of(KIND);
  constexpr int NV = PRIOR ? NB * 2 : NB;
  constexpr int NVAL = NV * RS;
  static_assert(NVAL <= 32, "at most 32 values per butterfly");
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int ncons = blockDim.x - 32, ncw = ncons >> 5;       // consumer threads / warps; the last warp produces
  const int seg_cap = ncons * VEC;                            // dims per segment (row slot length in the ring)
  const int seg_d0 = blockIdx.y * seg_cap;
  const int seg_len = min(D - seg_d0, seg_cap);
  constexpr int NROW = CPF == 2 ? NB + 3 : NB;               // CPF == 2: the step's 3 coefficient rows ride in the stage
  float* stage = reinterpret_cast<float*>(ring_smem);                                  // [NSTAGE][NROW][seg_cap]
  float* sred = stage + static_cast<size_t>(NSTAGE) * NROW * seg_cap;                   // [2][T][ncw][NV]
  uint64_t* full = reinterpret_cast<uint64_t*>(sred + static_cast<size_t>(2) * T * ncw * NV);
  uint64_t* empty = full + NSTAGE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = ceil_div(B, NB);
  const size_t rowTD = static_cast<size_t>(T) * D;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncw);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == ncw) {                                          // ---------------- producer warp
    if (lane == 0) {
      const uint32_t seg_bytes = static_cast<uint32_t>(seg_len) * 4u;
      int s = 0;
      uint32_t ph = 1;                                        // fresh barrier: waiting on parity 1 falls through
      for (int g = blockIdx.x; g < groups; g += gridDim.x) {
        const float* src[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) src[j] = z + static_cast<size_t>(min(g * NB + j, B - 1)) * rowTD + seg_d0;
        for (int n = 0; n < T; ++n) {
          mbar_wait(&empty[s], ph);
          mbar_arrive_expect_tx(&full[s], seg_bytes * NROW);
          float* dst = stage + static_cast<size_t>(s) * NROW * seg_cap;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            bulk_g2s(dst + j * seg_cap, src[j], seg_bytes, &full[s]);
            src[j] += D;
          }
          if (CPF == 2) {                                     // w / dd, dd w / I, C of this step (L2-resident table)
            const float* trow = tab + static_cast<size_t>(n) * NCF * D + seg_d0;
            bulk_g2s(dst + (NB + 0) * seg_cap, trow, seg_bytes, &full[s]);
            bulk_g2s(dst + (NB + 1) * seg_cap, trow + D, seg_bytes, &full[s]);
            bulk_g2s(dst + (NB + 2) * seg_cap, trow + (KIND == BRUNO_TP ? 4 : 2) * static_cast<size_t>(D), seg_bytes, &full[s]);
          }
          if (++s == NSTAGE) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
  const int dl = threadIdx.x * VEC;                           // dim inside the segment
  const bool active = dl < seg_len;
  const int d0 = seg_d0 + dl;
  const int d0c = active ? d0 : seg_d0;                       // inactive lanes compute on dim seg_d0 and are masked
  const int dlc = active ? dl : 0;
  const float lane_mask = active ? 1.f : 0.f;

  Vf<VEC> pc[N_PRIOR];
  Vf<VEC> c_e = zerov<VEC>(), c_f0 = zerov<VEC>(), c_a0 = zerov<VEC>(), c_da = zerov<VEC>();
  const Vf<VEC> mu = ldv<VEC>(mu0 + d0c);
#pragma unroll
  for (int i = 0; i < N_PRIOR; ++i) pc[i] = PRIOR ? ldv<VEC>(tab + static_cast<size_t>(T * NCF + i) * D + d0c) : zerov<VEC>();
  if (KIND == BRUNO_TP) {
    c_a0 = ldv<VEC>(tab + static_cast<size_t>(5) * D + d0c);
    const Vf<VEC> b0v = ldv<VEC>(tab + static_cast<size_t>(6) * D + d0c);
#pragma unroll
    for (int e = 0; e < VEC; ++e) c_da.x[e] = c_a0.x[e] - b0v.x[e];
  }
  const bool mu_zero = __all_sync(0xffffffffu, [&] { bool zr = true;
#pragma unroll
    for (int e = 0; e < VEC; ++e) zr = zr && (mu.x[e] == 0.f);
    return zr; }());
  float psum = 0.f;                                           // the prior's additive constant over this thread's dims
#pragma unroll
  for (int e = 0; e < VEC; ++e) psum += pc[1].x[e];

  struct Coef {
    Vf<VEC> w, dd, C;                                         // TP: w, dd w, C;   GP: dd, I, C
  };
  const float* const tp0 = tab + d0c;
  const size_t tstep = static_cast<size_t>(NCF) * D;
  auto load_coef = [&](const float* tp) {
    Coef c;
    c.w = ldv<VEC>(tp);
    c.dd = ldv<VEC>(tp + D);
    c.C = ldv<VEC>(tp + (KIND == BRUNO_TP ? 4 : 2) * static_cast<size_t>(D));
    return c;
  };
  const float* tpn = tp0;                                     // table row of the NEXT step (CPF) / of this step
  Coef cn;
  if (CPF == 1) cn = load_coef(tpn);
  int s = 0;
  uint32_t ph = 0;
  int par = 0;

  // The Student-t scan carries q_n = 1 + beta_n / (nu0 - 2) recursively: the Mahalanobis term of the posterior obeys
  // beta_{n+1} = beta_n + (x - mu_n)^2 / sigma_n (the literal update of student:100-106), i.e. q_{n+1} = q_n + r_n^2 = u_n,
  // so lg2(q_n) is the previous step's lg2(u_{n-1}): ONE MUFU and 6 FP instructions per element instead of 2 + 11 for
  // the closed form q_n = 1 + e S2 + f_n S1^2, no S2 in the hot loop, and none of the closed form's cancellation.
  // STATE = running prefix sums are read and written back (the reference's step-by-step API); the model builders'
  // sequence-at-once path runs the !STATE instantiation, which never touches S2.
  {
    constexpr bool FAST = !STATE;
    Vf<VEC> S1[NB], S2[NB], Q[NB], LQ[NB];
    for (int g = blockIdx.x; g < groups; g += gridDim.x, par ^= 1) {
      const int b0 = g * NB;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        S1[j] = zerov<VEC>();
        S2[j] = zerov<VEC>();
        LQ[j] = zerov<VEC>();
#pragma unroll
        for (int e = 0; e < VEC; ++e) Q[j].x[e] = 1.f;
      }
      if (STATE) {
        if (KIND == BRUNO_TP) {
          c_e = ldv<VEC>(tab + static_cast<size_t>(2) * D + d0c);
          c_f0 = ldv<VEC>(tab + static_cast<size_t>(3) * D + d0c);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (active && b0 + j < B) {
            S1[j] = ldv<VEC>(s1 + static_cast<size_t>(b0 + j) * D + d0);
            if (KIND == BRUNO_TP) {
              S2[j] = ldv<VEC>(s2 + static_cast<size_t>(b0 + j) * D + d0);
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                Q[j].x[e] = fmaf(c_f0.x[e], S1[j].x[e] * S1[j].x[e], fmaf(c_e.x[e], S2[j].x[e], 1.f));
                LQ[j].x[e] = fast_lg2(Q[j].x[e]);
              }
            }
          }
      }
      float* sr = sred + static_cast<size_t>(par) * T * ncw * NV;

      auto do_step = [&](int n, float* vals_pp) {
        Coef c;
        if (CPF == 1) {                                       // coefficients one step ahead (12 more registers)
          c = cn;
          tpn = (n + 1 == T) ? tp0 : tpn + tstep;
          cn = load_coef(tpn);
        } else if (CPF == 0) {
          c = load_coef(tpn);
          tpn = (n + 1 == T) ? tp0 : tpn + tstep;
        }
        mbar_wait(&full[s], ph);
        const float* xs = stage + static_cast<size_t>(s) * NROW * seg_cap + dlc;
        Vf<VEC> x[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(xs + j * seg_cap);
          x[j].x[0] = t.x; x[j].x[1] = t.y; x[j].x[2] = t.z; x[j].x[3] = t.w;
        }
        if (CPF == 2) {                                       // the step's coefficients came with the stage
          const float4 t0 = *reinterpret_cast<const float4*>(xs + (NB + 0) * seg_cap);
          const float4 t1 = *reinterpret_cast<const float4*>(xs + (NB + 1) * seg_cap);
          const float4 t2 = *reinterpret_cast<const float4*>(xs + (NB + 2) * seg_cap);
          c.w.x[0] = t0.x; c.w.x[1] = t0.y; c.w.x[2] = t0.z; c.w.x[3] = t0.w;
          c.dd.x[0] = t1.x; c.dd.x[1] = t1.y; c.dd.x[2] = t1.z; c.dd.x[3] = t1.w;
          c.C.x[0] = t2.x; c.C.x[1] = t2.y; c.C.x[2] = t2.z; c.C.x[3] = t2.w;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NSTAGE) {
          s = 0;
          ph ^= 1u;
        }
        if (!mu_zero) {                                       // warp-uniform; mu0 is 0 in every reference config
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int e = 0; e < VEC; ++e) x[j].x[e] -= mu.x[e];
        }
        const float fn = static_cast<float>(n);
        Vf<VEC> a2, b2;
        float csum = 0.f;                                     // the batch-independent additive constant of this step
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (KIND == BRUNO_TP) {
            a2.x[e] = fmaf(fn, c_da.x[e], c_a0.x[e]);
            b2.x[e] = a2.x[e] - c_da.x[e];
          }
          csum += c.C.x[e];
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          float acc = csum, accp = psum;
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float xz = x[j].x[e];
            if constexpr (KIND == BRUNO_TP) {
              const float r = fmaf(-c.dd.x[e], S1[j].x[e], c.w.x[e] * xz);
              const float u = fmaf(r, r, Q[j].x[e]);
              const float lu = fast_lg2(u);
              acc = fmaf(-a2.x[e], lu, fmaf(b2.x[e], LQ[j].x[e], acc));
              if (PRIOR) accp = fmaf(-pc[2].x[e], fast_lg2(fmaf(xz * xz, pc[0].x[e], 1.f)), accp);
              S1[j].x[e] += xz;
              if (!FAST) S2[j].x[e] = fmaf(xz, xz, S2[j].x[e]);
              Q[j].x[e] = u;
              LQ[j].x[e] = lu;
            } else {
              const float r = fmaf(-c.w.x[e], S1[j].x[e], xz);
              acc = fmaf(-c.dd.x[e], r * r, acc);
              if (PRIOR) accp = fmaf(-pc[0].x[e], xz * xz, accp);
              S1[j].x[e] += xz;
            }
          }
          if (PRIOR) {
            vals_pp[2 * j] = acc;
            vals_pp[2 * j + 1] = accp;
          } else {
            vals_pp[j] = acc;
          }
        }
      };

      int n0 = 0;
      for (; n0 + RS <= T; n0 += RS) {
        float vals[NVAL];
#pragma unroll
        for (int pp = 0; pp < RS; ++pp) do_step(n0 + pp, vals + pp * NV);
#pragma unroll
        for (int i = 0; i < NVAL; ++i) vals[i] *= lane_mask;
        warp_multi_reduce<NVAL>(vals, lane);
        if (lane_is_writer<NVAL>(lane)) {
          const int slot = slot_of_lane<NVAL>(lane);
          sr[((n0 + slot / NV) * ncw + warp) * NV + (slot % NV)] = vals[0];
        }
      }
      for (; n0 < T; ++n0) {
        float vals[NV];
        do_step(n0, vals);
#pragma unroll
        for (int i = 0; i < NV; ++i) vals[i] *= lane_mask;
        warp_multi_reduce<NV>(vals, lane);
        if (lane_is_writer<NV>(lane)) sr[(n0 * ncw + warp) * NV + slot_of_lane<NV>(lane)] = vals[0];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");   // consumers only; the producer keeps streaming
      const int nseg = gridDim.y;
      for (int i = threadIdx.x; i < T * NV; i += ncons) {
        const int n = i / NV, idx = i % NV, j = PRIOR ? idx >> 1 : idx, which = PRIOR ? idx & 1 : 0;
        if (b0 + j >= B) continue;
        float sum = 0.f;
        for (int w = 0; w < ncw; ++w) sum += sr[(n * ncw + w) * NV + idx];
        if (nseg == 1) {
          if (which == 0) lp[static_cast<size_t>(b0 + j) * T + n] = sum;
          else lp_prior[static_cast<size_t>(b0 + j) * T + n] = sum;
        } else {
          scratch[((static_cast<size_t>(blockIdx.y) * B + b0 + j) * T + n) * 2 + which] = sum;
        }
      }
      if (STATE && active) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (b0 + j < B) {
            stv<VEC>(s1 + static_cast<size_t>(b0 + j) * D + d0, S1[j]);
            if (KIND == BRUNO_TP) stv<VEC>(s2 + static_cast<size_t>(b0 + j) * D + d0, S2[j]);
          }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// backward scan
// ------------------------------------------------------------------------------------------------------------
// double-single accumulation of S2 = sum (x - mu0)^2 in the backward: the reverse pass recovers the prefix sums by
// SUBTRACTING observations from the total, and q = 1 + e S2 + f S1^2 needs S2 to an absolute accuracy << 1/e even when
// the total is ~1e8 (untrained flows produce |z| ~ 1e4): plain fp32 leaves a residue of +-32 there and q < 0 -> NaN.
__device__ __forceinline__ void ds_add(float& hi, float& lo, float v, float ve) {
  const float s = hi + v;
  const float bb = s - hi;
  const float err = (hi - (s - bb)) + (v - bb);
  hi = s;
  lo += err + ve;
}

template <int KIND, int VEC, int NB, bool PRIOR>
__global__ void __launch_bounds__(256)
process_bwd_kernel(const float* __restrict__ z, const float* __restrict__ mu0, const float* __restrict__ tab,
                   const float* __restrict__ dlp, const float* __restrict__ dlpp, float* __restrict__ dz,
                   float* __restrict__ partial, int B, int T, int D, int groups_per_block) {
  constexpr int NCF = ncf_of(KIND);
  constexpr float INV_LN2 = 1.4426950408889634f;
  const int d0 = (blockIdx.y * blockDim.x + threadIdx.x) * VEC;
  if (d0 >= D) return;
  const size_t rowTD = static_cast<size_t>(T) * D;
  const float* gtab = tab + static_cast<size_t>(T * NCF + N_PRIOR) * D;
  const Vf<VEC> mu = ldv<VEC>(mu0 + d0);
  Vf<VEC> pc[N_PRIOR];
#pragma unroll
  for (int i = 0; i < N_PRIOR; ++i) pc[i] = ldv<VEC>(tab + static_cast<size_t>(T * NCF + i) * D + d0);

  Vf<VEC> pg[3], pcg[N_PRIOR];
#pragma unroll
  for (int p = 0; p < 3; ++p) pg[p] = zerov<VEC>();
#pragma unroll
  for (int i = 0; i < N_PRIOR; ++i) pcg[i] = zerov<VEC>();

  for (int g = 0; g < groups_per_block; ++g) {
    const int b0 = (blockIdx.x * groups_per_block + g) * NB;
    if (b0 >= B) break;
    Vf<VEC> S1[NB], S2[NB], S2lo[NB], G1[NB], G2[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      S1[j] = zerov<VEC>(); S2[j] = zerov<VEC>(); S2lo[j] = zerov<VEC>(); G1[j] = zerov<VEC>(); G2[j] = zerov<VEC>();
    }
    // pass 1: totals
    for (int n = 0; n < T; ++n) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (b0 + j < B) {
          Vf<VEC> x = ldv<VEC>(z + static_cast<size_t>(b0 + j) * rowTD + static_cast<size_t>(n) * D + d0);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float xz = x.x[e] - mu.x[e];
            S1[j].x[e] += xz;
            if (KIND == BRUNO_TP) {
              const float p2 = xz * xz;
              ds_add(S2[j].x[e], S2lo[j].x[e], p2, fmaf(xz, xz, -p2));
            }
          }
        }
    }
    // pass 2: reverse
    for (int n = T - 1; n >= 0; --n) {
      Vf<VEC> cf[NCF], cg[NCF];
#pragma unroll
      for (int i = 0; i < NCF; ++i) {
        cf[i] = ldv<VEC>(tab + static_cast<size_t>(n * NCF + i) * D + d0);
        cg[i] = zerov<VEC>();
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (b0 + j >= B) continue;
        const size_t off = static_cast<size_t>(b0 + j) * rowTD + static_cast<size_t>(n) * D + d0;
        Vf<VEC> x = ldv<VEC>(z + off), gx;
        const float dl = __ldg(dlp + static_cast<size_t>(b0 + j) * T + n);
        const float dlq = PRIOR ? __ldg(dlpp + static_cast<size_t>(b0 + j) * T + n) : 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float xz = x.x[e] - mu.x[e];
          float dxz, dS1, dS2 = 0.f;
          if constexpr (KIND == BRUNO_TP) {
            const float s1v = S1[j].x[e] - xz;            // prefix sums BEFORE observation n
            const float p2 = xz * xz;
            ds_add(S2[j].x[e], S2lo[j].x[e], -p2, -fmaf(xz, xz, -p2));
            const float s2v = fmaxf(S2[j].x[e] + S2lo[j].x[e], 0.f);
            S1[j].x[e] = s1v;
            const float s1sq = s1v * s1v;
            const float r = fmaf(-cf[1].x[e], s1v, cf[0].x[e] * xz);
            const float q = fmaf(cf[3].x[e], s1sq, fmaf(cf[2].x[e], s2v, 1.f));
            const float u = fmaf(r, r, q);
            const float gu = -cf[5].x[e] * INV_LN2 * fast_rcp(u) * dl;
            const float gq = fmaf(cf[6].x[e] * INV_LN2 * dl, fast_rcp(q), gu);
            const float gr = 2.f * r * gu;
            dxz = gr * cf[0].x[e];
            dS1 = fmaf(gq * cf[3].x[e], 2.f * s1v, -gr * cf[1].x[e]);
            dS2 = gq * cf[2].x[e];
            cg[0].x[e] = fmaf(gr, xz, cg[0].x[e]);
            cg[1].x[e] = fmaf(-gr, s1v, cg[1].x[e]);
            cg[2].x[e] = fmaf(gq, s2v, cg[2].x[e]);
            cg[3].x[e] = fmaf(gq, s1sq, cg[3].x[e]);
            cg[4].x[e] += dl;
            cg[5].x[e] = fmaf(-fast_lg2(u), dl, cg[5].x[e]);
            cg[6].x[e] = fmaf(fast_lg2(q), dl, cg[6].x[e]);
            if (PRIOR) {
              const float x2 = xz * xz;
              const float up = fmaf(x2, pc[0].x[e], 1.f);
              const float gup = -pc[2].x[e] * INV_LN2 * fast_rcp(up) * dlq;
              dxz = fmaf(gup * 2.f * pc[0].x[e], xz, dxz);
              pcg[0].x[e] = fmaf(gup, x2, pcg[0].x[e]);
              pcg[1].x[e] += dlq;
              pcg[2].x[e] = fmaf(-fast_lg2(up), dlq, pcg[2].x[e]);
            }
          } else {
            const float s1v = S1[j].x[e] - xz;
            S1[j].x[e] = s1v;
            const float r = fmaf(-cf[0].x[e], s1v, xz);
            const float gr = -2.f * cf[1].x[e] * r * dl;
            dxz = gr;
            dS1 = -gr * cf[0].x[e];
            cg[0].x[e] = fmaf(-gr, s1v, cg[0].x[e]);
            cg[1].x[e] = fmaf(-r * r, dl, cg[1].x[e]);
            cg[2].x[e] += dl;
            if (PRIOR) {
              dxz = fmaf(-2.f * pc[0].x[e] * dlq, xz, dxz);
              pcg[0].x[e] = fmaf(-xz * xz, dlq, pcg[0].x[e]);
              pcg[1].x[e] += dlq;
            }
          }
          gx.x[e] = dxz + G1[j].x[e] + 2.f * xz * G2[j].x[e];
          G1[j].x[e] += dS1;
          G2[j].x[e] += dS2;
        }
        stv<VEC>(dz + off, gx);
      }
      // chain the coefficient gradients of this step to the raw variables
#pragma unroll
      for (int i = 0; i < NCF; ++i)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          if (KIND == BRUNO_GP && p == 1) continue;
          Vf<VEC> dc = ldv<VEC>(gtab + (static_cast<size_t>(n * NCF + i) * 3 + p) * D + d0);
#pragma unroll
          for (int e = 0; e < VEC; ++e) pg[p].x[e] = fmaf(cg[i].x[e], dc.x[e], pg[p].x[e]);
        }
    }
  }
  if (PRIOR) {
#pragma unroll
    for (int i = 0; i < N_PRIOR; ++i)
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        Vf<VEC> dc = ldv<VEC>(gtab + (static_cast<size_t>(T * NCF + i) * 3 + p) * D + d0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) pg[p].x[e] = fmaf(pcg[i].x[e], dc.x[e], pg[p].x[e]);
      }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) stv<VEC>(partial + (static_cast<size_t>(blockIdx.x) * 3 + p) * D + d0, pg[p]);
}

// ------------------------------------------------------------------------------------------------------------
// backward scan, persistent + bulk-copy ring (default when D % 4 == 0 and the pointers are 16-byte aligned)
// ------------------------------------------------------------------------------------------------------------
// Same arithmetic per element as process_bwd_kernel.  What changes is where the bytes come from:
//   * z through the producer warp's bulk-copy ring, forward for the totals and again in reverse (the second read of a
//     group is an L2 hit: NB * T * D * 4 bytes per block are in flight);
//   * the coefficient tables were the real bound (28 vector loads per thread-step for NB * 4 elements, ~14x the z bytes
//     through L2): e, A2 and B2 have step-independent derivatives, so their accumulators run over all steps and groups of
//     the block and are chained to the raw variables ONCE at the end; A2_n / B2_n themselves are a0 + n da; C's
//     accumulator is a scalar per step.  Per step: 3 coefficient + 12 derivative vectors (TP), 2 + 6 (GP);
//   * coefficients and dlp one step ahead in registers.
template <int KIND, int NB, bool PRIOR, int NSTAGE, int MINB, bool CT>
__global__ void __launch_bounds__(256, MINB)
process_bwd_ring_kernel(const float* __restrict__ z, const float* __restrict__ mu0, const float* __restrict__ tab,
                        const float* __restrict__ dlp, const float* __restrict__ dlpp, float* __restrict__ dz,
                        float* __restrict__ partial, int B, int T, int D) {
  constexpr int VEC = 4;
  constexpr int NCF = ncf_of(KIND);
  constexpr float INV_LN2 = 1.4426950408889634f;
  extern __shared__ __align__(16) unsigned char ring_smem[];
  const int ncons = blockDim.x - 32, ncw = ncons >> 5;
  const int seg_cap = ncons * VEC;
  const int seg_d0 = blockIdx.y * seg_cap;
  const int seg_len = min(D - seg_d0, seg_cap);
  // CT: the reverse pass's stages also carry the step's coefficient rows (TP w, dd w, f'; GP dd, I) and the derivative
  // rows of its step-dependent chains (TP 4 x 3, GP 3 x 2): no LDG and no 64-bit address arithmetic in the hot loop
  constexpr int NCO = KIND == BRUNO_TP ? 3 : 2, NDE = KIND == BRUNO_TP ? 12 : 6;
  constexpr int NROW = CT ? NB + NCO + NDE : NB;
  float* stage = reinterpret_cast<float*>(ring_smem);                                  // [NSTAGE][NROW][seg_cap]
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + static_cast<size_t>(NSTAGE) * NROW * seg_cap);
  uint64_t* empty = full + NSTAGE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = ceil_div(B, NB);
  const size_t rowTD = static_cast<size_t>(T) * D;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncw);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == ncw) {                                          // ---------------- producer warp
    if (lane == 0) {
      const uint32_t seg_bytes = static_cast<uint32_t>(seg_len) * 4u;
      int s = 0;
      uint32_t ph = 1;
      for (int g = blockIdx.x; g < groups; g += gridDim.x) {
        const float* src[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) src[j] = z + static_cast<size_t>(min(g * NB + j, B - 1)) * rowTD + seg_d0;
        for (int it = 0; it < 2 * T; ++it) {                  // steps 0..T-1 (totals), then T-1..0 (reverse pass)
          const int n = it < T ? it : 2 * T - 1 - it;
          mbar_wait(&empty[s], ph);
          const bool tables = CT && it >= T;
          mbar_arrive_expect_tx(&full[s], seg_bytes * (tables ? NROW : NB));
          float* dst = stage + static_cast<size_t>(s) * NROW * seg_cap;
#pragma unroll
          for (int j = 0; j < NB; ++j) bulk_g2s(dst + j * seg_cap, src[j] + static_cast<size_t>(n) * D, seg_bytes, &full[s]);
          if (tables) {
            const float* trow = tab + static_cast<size_t>(n) * NCF * D + seg_d0;
            const float* grow = tab + static_cast<size_t>(T * NCF + N_PRIOR) * D + static_cast<size_t>(n) * NCF * 3 * D + seg_d0;
            bulk_g2s(dst + (NB + 0) * seg_cap, trow, seg_bytes, &full[s]);
            bulk_g2s(dst + (NB + 1) * seg_cap, trow + D, seg_bytes, &full[s]);
            if (KIND == BRUNO_TP) bulk_g2s(dst + (NB + 2) * seg_cap, trow + 3 * static_cast<size_t>(D), seg_bytes, &full[s]);
            int r = NB + NCO;
#pragma unroll
            for (int k = 0; k < (KIND == BRUNO_TP ? 4 : 3); ++k) {
              const int ci = KIND == BRUNO_TP ? (k < 2 ? k : k + 1) : k;       // TP: w, dd w, f', C = rows 0, 1, 3, 4
#pragma unroll
              for (int p = 0; p < 3; ++p) {
                if (KIND == BRUNO_GP && p == 1) continue;
                bulk_g2s(dst + r * seg_cap, grow + static_cast<size_t>(ci * 3 + p) * D, seg_bytes, &full[s]);
                ++r;
              }
            }
          }
          if (++s == NSTAGE) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
  const int dl_ = threadIdx.x * VEC;
  const bool active = dl_ < seg_len;
  const int d0 = seg_d0 + dl_;
  const int d0c = active ? d0 : seg_d0;
  const int dlc = active ? dl_ : 0;
  const float* gtab = tab + static_cast<size_t>(T * NCF + N_PRIOR) * D;
  const Vf<VEC> mu = ldv<VEC>(mu0 + d0c);
  Vf<VEC> pc[N_PRIOR], pcg[N_PRIOR], pg[3];
#pragma unroll
  for (int i = 0; i < N_PRIOR; ++i) {
    pc[i] = PRIOR ? ldv<VEC>(tab + static_cast<size_t>(T * NCF + i) * D + d0c) : zerov<VEC>();
    pcg[i] = zerov<VEC>();
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) pg[p] = zerov<VEC>();
  Vf<VEC> c_e = zerov<VEC>(), c_a0 = zerov<VEC>(), c_da = zerov<VEC>();
  Vf<VEC> cgE = zerov<VEC>(), cg5 = zerov<VEC>(), cg6 = zerov<VEC>();      // step-independent derivatives: chained at the end
  if (KIND == BRUNO_TP) {
    c_e = ldv<VEC>(tab + static_cast<size_t>(2) * D + d0c);
    c_a0 = ldv<VEC>(tab + static_cast<size_t>(5) * D + d0c);
    const Vf<VEC> b0v = ldv<VEC>(tab + static_cast<size_t>(6) * D + d0c);
#pragma unroll
    for (int e = 0; e < VEC; ++e) c_da.x[e] = c_a0.x[e] - b0v.x[e];
  }
  struct Coef {
    Vf<VEC> a, b, f;                                          // TP: w, dd w, f';   GP: dd, I, -
  };
  auto load_coef = [&](int n) {
    const float* tp = tab + static_cast<size_t>(n) * NCF * D + d0c;
    Coef c;
    c.a = ldv<VEC>(tp);
    c.b = ldv<VEC>(tp + D);
    c.f = KIND == BRUNO_TP ? ldv<VEC>(tp + 3 * static_cast<size_t>(D)) : zerov<VEC>();
    return c;
  };
  int s = 0;
  uint32_t ph = 0;
  auto next_stage = [&]() -> const float* {
    mbar_wait(&full[s], ph);
    return stage + static_cast<size_t>(s) * NROW * seg_cap + dlc;
  };
  auto lds4 = [](const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    Vf<VEC> r;
    r.x[0] = t.x; r.x[1] = t.y; r.x[2] = t.z; r.x[3] = t.w;
    return r;
  };
  auto release_stage = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == NSTAGE) {
      s = 0;
      ph ^= 1u;
    }
  };

  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int b0 = g * NB;
    Vf<VEC> S1[NB], S2[NB], S2lo[NB], G1[NB], G2[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      S1[j] = zerov<VEC>(); S2[j] = zerov<VEC>(); S2lo[j] = zerov<VEC>(); G1[j] = zerov<VEC>(); G2[j] = zerov<VEC>();
    }
    // coefficients and upstream gradients of the LAST step: in flight while pass 1 runs
    Coef cn;
    if (!CT) cn = load_coef(T - 1);
    float dln[NB], dqn[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const bool valid = b0 + j < B;
      dln[j] = valid ? __ldg(dlp + static_cast<size_t>(b0 + j) * T + (T - 1)) : 0.f;
      dqn[j] = (PRIOR && valid) ? __ldg(dlpp + static_cast<size_t>(b0 + j) * T + (T - 1)) : 0.f;
    }
    // pass 1: totals
    for (int n = 0; n < T; ++n) {
      const float* xs = next_stage();
      Vf<VEC> x[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(xs + j * seg_cap);
        x[j].x[0] = t.x; x[j].x[1] = t.y; x[j].x[2] = t.z; x[j].x[3] = t.w;
      }
      release_stage();
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float xz = x[j].x[e] - mu.x[e];
          S1[j].x[e] += xz;
          if (KIND == BRUNO_TP) {
            const float p2 = xz * xz;
            ds_add(S2[j].x[e], S2lo[j].x[e], p2, fmaf(xz, xz, -p2));
          }
        }
    }
    // pass 2: reverse
    for (int n = T - 1; n >= 0; --n) {
      Coef c;
      if (!CT) c = cn;
      float dlv[NB], dqv[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        dlv[j] = dln[j];
        dqv[j] = dqn[j];
      }
      if (n > 0) {
        if (!CT) cn = load_coef(n - 1);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const bool valid = b0 + j < B;
          dln[j] = valid ? __ldg(dlp + static_cast<size_t>(b0 + j) * T + (n - 1)) : 0.f;
          dqn[j] = (PRIOR && valid) ? __ldg(dlpp + static_cast<size_t>(b0 + j) * T + (n - 1)) : 0.f;
        }
      }
      const float* xs = next_stage();
      Vf<VEC> x[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(xs + j * seg_cap);
        x[j].x[0] = t.x; x[j].x[1] = t.y; x[j].x[2] = t.z; x[j].x[3] = t.w;
      }
      if (CT) {
        c.a = lds4(xs + (NB + 0) * seg_cap);
        c.b = lds4(xs + (NB + 1) * seg_cap);
        c.f = KIND == BRUNO_TP ? lds4(xs + (NB + 2) * seg_cap) : zerov<VEC>();
      } else {
        release_stage();
      }
      const float fn = static_cast<float>(n);
      Vf<VEC> a2 = zerov<VEC>(), b2 = zerov<VEC>();           // the table's A2_n, B2_n
      if (KIND == BRUNO_TP) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          a2.x[e] = fmaf(fn, c_da.x[e], c_a0.x[e]);
          b2.x[e] = a2.x[e] - c_da.x[e];
        }
      }
      Vf<VEC> cga = zerov<VEC>(), cgb = zerov<VEC>(), cgf = zerov<VEC>();
      float cgC = 0.f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float dl = dlv[j], dlq = dqv[j];                // 0 for rows past B: no gradient contribution
        cgC += dl;
        Vf<VEC> gx;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float xz = x[j].x[e] - mu.x[e];
          float dxz, dS1, dS2 = 0.f;
          if constexpr (KIND == BRUNO_TP) {
            const float s1v = S1[j].x[e] - xz;
            const float p2 = xz * xz;
            ds_add(S2[j].x[e], S2lo[j].x[e], -p2, -fmaf(xz, xz, -p2));
            const float s2v = fmaxf(S2[j].x[e] + S2lo[j].x[e], 0.f);
            S1[j].x[e] = s1v;
            const float s1sq = s1v * s1v;
            const float r = fmaf(-c.b.x[e], s1v, c.a.x[e] * xz);
            const float q = fmaf(c.f.x[e], s1sq, fmaf(c_e.x[e], s2v, 1.f));
            const float u = fmaf(r, r, q);
            const float gu = -a2.x[e] * INV_LN2 * fast_rcp(u) * dl;
            const float gq = fmaf(b2.x[e] * INV_LN2 * dl, fast_rcp(q), gu);
            const float gr = 2.f * r * gu;
            dxz = gr * c.a.x[e];
            dS1 = fmaf(gq * c.f.x[e], 2.f * s1v, -gr * c.b.x[e]);
            dS2 = gq * c_e.x[e];
            cga.x[e] = fmaf(gr, xz, cga.x[e]);
            cgb.x[e] = fmaf(-gr, s1v, cgb.x[e]);
            cgE.x[e] = fmaf(gq, s2v, cgE.x[e]);
            cgf.x[e] = fmaf(gq, s1sq, cgf.x[e]);
            cg5.x[e] = fmaf(-fast_lg2(u), dl, cg5.x[e]);
            cg6.x[e] = fmaf(fast_lg2(q), dl, cg6.x[e]);
            if (PRIOR) {
              const float x2 = xz * xz;
              const float up = fmaf(x2, pc[0].x[e], 1.f);
              const float gup = -pc[2].x[e] * INV_LN2 * fast_rcp(up) * dlq;
              dxz = fmaf(gup * 2.f * pc[0].x[e], xz, dxz);
              pcg[0].x[e] = fmaf(gup, x2, pcg[0].x[e]);
              pcg[1].x[e] += dlq;
              pcg[2].x[e] = fmaf(-fast_lg2(up), dlq, pcg[2].x[e]);
            }
          } else {
            const float s1v = S1[j].x[e] - xz;
            S1[j].x[e] = s1v;
            const float r = fmaf(-c.a.x[e], s1v, xz);
            const float gr = -2.f * c.b.x[e] * r * dl;
            dxz = gr;
            dS1 = -gr * c.a.x[e];
            cga.x[e] = fmaf(-gr, s1v, cga.x[e]);
            cgb.x[e] = fmaf(-r * r, dl, cgb.x[e]);
            if (PRIOR) {
              dxz = fmaf(-2.f * pc[0].x[e] * dlq, xz, dxz);
              pcg[0].x[e] = fmaf(-xz * xz, dlq, pcg[0].x[e]);
              pcg[1].x[e] += dlq;
            }
          }
          gx.x[e] = dxz + G1[j].x[e] + 2.f * xz * G2[j].x[e];
          G1[j].x[e] += dS1;
          G2[j].x[e] += dS2;
        }
        if (active && b0 + j < B) stv<VEC>(dz + static_cast<size_t>(b0 + j) * rowTD + static_cast<size_t>(n) * D + d0, gx);
      }
      // chain this step's accumulators (the coefficients whose derivatives depend on the step) to the raw variables
      const float* gt = gtab + static_cast<size_t>(n) * NCF * 3 * D + d0c;
      constexpr int NP = KIND == BRUNO_TP ? 3 : 2;            // derivative rows per coefficient in the stage
      const float* gs = xs + (NB + NCO) * seg_cap;            // [k][pi] rows of the stage (CT)
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        if (KIND == BRUNO_GP && p == 1) continue;
        const int pi = KIND == BRUNO_TP ? p : (p >> 1);
        const Vf<VEC> da_ = CT ? lds4(gs + (0 * NP + pi) * seg_cap) : ldv<VEC>(gt + static_cast<size_t>(0 * 3 + p) * D);
        const Vf<VEC> db_ = CT ? lds4(gs + (1 * NP + pi) * seg_cap) : ldv<VEC>(gt + static_cast<size_t>(1 * 3 + p) * D);
        const Vf<VEC> dC_ = CT ? lds4(gs + ((KIND == BRUNO_TP ? 3 : 2) * NP + pi) * seg_cap)
                               : ldv<VEC>(gt + static_cast<size_t>((KIND == BRUNO_TP ? 4 : 2) * 3 + p) * D);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          pg[p].x[e] = fmaf(cgC, dC_.x[e], fmaf(cgb.x[e], db_.x[e], fmaf(cga.x[e], da_.x[e], pg[p].x[e])));
        if (KIND == BRUNO_TP) {
          const Vf<VEC> df_ = CT ? lds4(gs + (2 * NP + pi) * seg_cap) : ldv<VEC>(gt + static_cast<size_t>(3 * 3 + p) * D);
#pragma unroll
          for (int e = 0; e < VEC; ++e) pg[p].x[e] = fmaf(cgf.x[e], df_.x[e], pg[p].x[e]);
        }
      }
      if (CT) release_stage();
    }
  }
  if (KIND == BRUNO_TP) {                                     // e, A2, B2: d/d(raw variable) does not depend on the step
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const Vf<VEC> dE = ldv<VEC>(gtab + static_cast<size_t>(2 * 3 + p) * D + d0c);
      const Vf<VEC> d5 = ldv<VEC>(gtab + static_cast<size_t>(5 * 3 + p) * D + d0c);
      const Vf<VEC> d6 = ldv<VEC>(gtab + static_cast<size_t>(6 * 3 + p) * D + d0c);
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        pg[p].x[e] = fmaf(cg6.x[e], d6.x[e], fmaf(cg5.x[e], d5.x[e], fmaf(cgE.x[e], dE.x[e], pg[p].x[e])));
    }
  }
  if (PRIOR) {
#pragma unroll
    for (int i = 0; i < N_PRIOR; ++i)
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const Vf<VEC> dc = ldv<VEC>(gtab + (static_cast<size_t>(T * NCF + i) * 3 + p) * D + d0c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) pg[p].x[e] = fmaf(pcg[i].x[e], dc.x[e], pg[p].x[e]);
      }
  }
  if (active) {
#pragma unroll
    for (int p = 0; p < 3; ++p) stv<VEC>(partial + (static_cast<size_t>(blockIdx.x) * 3 + p) * D + d0, pg[p]);
  }
}

__global__ void process_bwd_finalize(const float* __restrict__ partial, int nblk, int D, float* __restrict__ dvar,
                                     float* __restrict__ dnu, float* __restrict__ dcorr) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float a[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < nblk; ++k)
    for (int p = 0; p < 3; ++p) a[p] += partial[(static_cast<size_t>(k) * 3 + p) * D + d];
  dvar[d] = a[0];
  if (dnu) dnu[d] = a[1];
  dcorr[d] = a[2];
}

// ------------------------------------------------------------------------------------------------------------
// posterior after each prefix + sampler (sampling branch only; follows the reference recursion literally)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void process_posterior_kernel(int kind, const float* __restrict__ z, const float* __restrict__ var_vbl,
                                         const float* __restrict__ nu_vbl, const float* __restrict__ corr_vbl,
                                         const float* __restrict__ mu0, float min_nu, float* __restrict__ mu_out,
                                         float* __restrict__ var_out, float* __restrict__ nu_out, int B, int T, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (d >= D) return;
  const float sp = softplus_f(var_vbl[d]);
  const float v = sp * sp;
  const float c = v / (1.f + expf(-corr_vbl[d]));
  const float nu0 = kind == BRUNO_TP ? expf(nu_vbl[d]) + min_nu : 0.f;
  const float m0 = mu0[d];
  float mu = m0, sigma = v, nu = nu0, beta = 0.f, xs = 0.f, k = v;
  const size_t o = (static_cast<size_t>(b) * (T + 1)) * D + d;
  mu_out[o] = mu;
  var_out[o] = sigma;
  if (kind == BRUNO_TP && b == 0) nu_out[d] = nu;
  for (int t = 0; t < T; ++t) {
    const float x = z[(static_cast<size_t>(b) * T + t) * D + d];
    const float i = static_cast<float>(t + 1);
    const float xz = x - m0;
    const float dd = c / (v + c * (i - 1.f));
    mu = (1.f - dd) * mu + x * dd;
    if (kind == BRUNO_TP) {
      const float a_i = (c * (i - 2.f) + v) / ((v - c) * (c * (i - 1.f) + v));
      const float b_i = -c / ((v - c) * (c * (i - 1.f) + v));
      const float b_p = -c / ((v - c) * (c * (i - 2.f) + v));
      beta = beta + (a_i - b_i) * xz * xz + b_i * (xs + xz) * (xs + xz) - b_p * xs * xs;
      xs += xz;
      k = (1.f - dd) * k + (v - c) * dd;
      nu = nu + 1.f;
      sigma = (nu0 + beta - 2.f) / (nu - 2.f) * k;
      if (b == 0) nu_out[static_cast<size_t>(t + 1) * D + d] = nu;
    } else {
      sigma = (1.f - dd) * sigma + (v - c) * dd;
    }
    mu_out[o + static_cast<size_t>(t + 1) * D] = mu;
    var_out[o + static_cast<size_t>(t + 1) * D] = sigma;
  }
}

__global__ void process_sample_kernel(int kind, const float* __restrict__ rvs, const float* __restrict__ mu,
                                      const float* __restrict__ var, const float* __restrict__ nu,
                                      float* __restrict__ out, int n, int B, int D, long stride_b) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n) * B * D;
  if (i >= total) return;
  const int d = static_cast<int>(i % D);
  const int b = static_cast<int>((i / D) % B);
  const float m = mu[static_cast<size_t>(b) * stride_b + d], s = var[static_cast<size_t>(b) * stride_b + d];
  if (kind == BRUNO_TP) {
    const float r0 = rvs[i], r1 = rvs[total + i];
    const float a = fminf(r0, r1), bb = fmaxf(r0, r1);
    const float ang = 2.f * 3.14159265358979323846f * a / bb;
    const float u = bb * cosf(ang), v = bb * sinf(ang);
    const float w = u * u + v * v;
    const float nv = nu[d];
    // pow(w, -2/nu) - 1 evaluated as expm1(-2/nu * ln w): same value, without the fp32 cancellation
    const float t = u * sqrtf(nv * expm1f(-2.f / nv * logf(w)) / w);
    out[i] = m + sqrtf(s * (nv - 2.f) / nv) * t;
  } else {
    out[i] = m + sqrtf(s) * rvs[i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------------
struct ScanGeom {
  int vec, threads, nseg;
};
// Forward-scan variant: 0 = default (bulk-copy ring kernel when D % 4 == 0 and the pointers are aligned, else the
// register-staged kernel), 1-4 / 9 = register-staged variants (9 = its default shape), 10-13 = ring variants.
// BRUNO_TP_VARIANT in the environment or bruno_process_set_fwd_variant() at run time (A/B benches and tests).
static int g_fwd_variant = -1;
static int fwd_variant() {
  if (g_fwd_variant < 0) {
    const char* e = getenv("BRUNO_TP_VARIANT");
    g_fwd_variant = e ? atoi(e) : 0;
  }
  return g_fwd_variant;
}

static ScanGeom scan_geom(int D, bool ptr_ok, bool fwd = false) {
  ScanGeom g;
  g.vec = (D % 4 == 0 && ptr_ok) ? 4 : 1;
  int maxt = 256;
  if (fwd && fwd_variant() == 1 && D % 4 == 0 && ptr_ok && D / 2 <= 512) {   // 2-wide layout: more, lighter threads
    g.vec = 2;
    maxt = 512;
  }
  const int chunks = D / g.vec;
  g.threads = chunks >= maxt ? maxt : ((chunks + 31) / 32) * 32;
  g.nseg = (chunks + g.threads - 1) / g.threads;
  return g;
}

constexpr int BWD_NB = 2;

// Variants of the forward scan (rows per thread NB, prefetch depth PF, resident blocks MINB), picked for occupancy /
// memory-level parallelism: the first version (NB 4, PF 1, 128 registers, 14 warps per SM) reached 72 % of HBM peak for
// the GP but only 36 % for the TP, whose 3 MUFU + ~17 FP ops per element need more warps to hide latency.
template <int KIND, int VEC, int NB, int PF, int MINB, int RS, int MAXT>
static int launch_fwd_variant(const float* z, const float* mu0, const float* table, float* s1, float* s2, float* lp,
                              float* lp_prior, float* scratch, int B, int T, int D, const ScanGeom& g, cudaStream_t st) {
  dim3 grid(ceil_div(B, NB), g.nseg);
  const size_t smem = static_cast<size_t>(T) * (g.threads / 32) * NB * 2 * sizeof(float);
  auto kp = process_fwd_kernel<KIND, VEC, NB, true, PF, MINB, RS, MAXT>;
  auto kn = process_fwd_kernel<KIND, VEC, NB, false, PF, MINB, RS, MAXT>;
  auto k = lp_prior ? kp : kn;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) {
      set_last_error("bruno_process_fwd: T=%d too long for one launch (smem %zu)", T, smem);
      return BRUNO_EINVAL;
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  int rc;
  {
    ProfSpan span(BRUNO_PROF_PROCESS_FWD, st);
    k<<<grid, g.threads, smem, st>>>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D);
    rc = check_launch("process_fwd_kernel");
  }
  if (rc) return rc;
  if (g.nseg > 1) {
    process_fwd_finalize<<<ceil_div(B * T, 256), 256, 0, st>>>(scratch, lp, lp_prior, g.nseg, B * T);
    rc = check_launch("process_fwd_finalize");
  }
  return rc;
}

// Ring kernel launch.  Returns +1 when this shape cannot take the ring path (caller falls back to the register-staged
// kernel): T too long for the shared-memory partials.
// Per-device launch geometry of a ring kernel instantiation: {dynamic smem opt-in, resident blocks per SM, SM count} are
// properties of the DEVICE, and two host threads may reach a call site together: keyed by (device, slot), under a lock.
struct RingLaunchCache {
  struct Entry { size_t smem = 0; int threads = 0, per_sm = 0, sms = 0; };
  Entry e[16][4];
  std::mutex mu;
  template <class K>
  Entry get(K kernel, int slot, size_t smem, int threads) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Entry& c = e[dev & 15][slot];
    if (c.smem != smem || c.threads != threads || c.sms == 0) {
      Entry n;
      n.smem = smem;
      n.threads = threads;
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n.per_sm, kernel, threads + 32, smem) != cudaSuccess ||
          cudaDeviceGetAttribute(&n.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        n.per_sm = 0;                                          // the caller falls back to the register-staged scan
        n.sms = 0;
      }
      c = n;
    }
    return c;
  }
};

constexpr int RING_CONS = 224;   // consumer threads per block at most (7 warps + 1 producer warp = 256 threads)
static ScanGeom ring_geom(int D) {
  ScanGeom g;
  g.vec = 4;
  const int chunks = D / 4;
  g.threads = chunks >= RING_CONS ? RING_CONS : ((chunks + 31) / 32) * 32;
  g.nseg = (chunks + g.threads - 1) / g.threads;
  return g;
}
template <int KIND, int NB, int NSTAGE, int RS, int MINB, int CPF>
static int launch_fwd_ring(const float* z, const float* mu0, const float* table, float* s1, float* s2, float* lp,
                           float* lp_prior, float* scratch, int B, int T, int D, cudaStream_t st) {
  const ScanGeom g = ring_geom(D);
  const int ncw = g.threads / 32;
  const size_t smem = static_cast<size_t>(NSTAGE) * (CPF == 2 ? NB + 3 : NB) * g.threads * 4 * sizeof(float) +
                      static_cast<size_t>(2) * T * ncw * NB * 2 * sizeof(float) + 2 * NSTAGE * sizeof(uint64_t);
  if (smem > 112 * 1024 || (g.nseg > 1 && scratch == nullptr)) return 1;
  auto k = s1 ? (lp_prior ? process_fwd_ring_kernel<KIND, NB, true, NSTAGE, RS, MINB, CPF, true>
                          : process_fwd_ring_kernel<KIND, NB, false, NSTAGE, RS, MINB, CPF, true>)
              : (lp_prior ? process_fwd_ring_kernel<KIND, NB, true, NSTAGE, RS, MINB, CPF, false>
                          : process_fwd_ring_kernel<KIND, NB, false, NSTAGE, RS, MINB, CPF, false>);
  // resident blocks per SM of this instantiation at this (threads, smem): queried once per change and device, not per launch
  static RingLaunchCache cache;
  const RingLaunchCache::Entry ce = cache.get(k, (s1 ? 2 : 0) + (lp_prior ? 1 : 0), smem, g.threads);
  const int per_sm = ce.per_sm, sm_count = ce.sms;
  if (per_sm < 1) return 1;
  const int groups = ceil_div(B, NB);
  const int resident = ceil_div(per_sm * sm_count, g.nseg);
  dim3 grid(groups < resident ? groups : resident, g.nseg);
  int rc;
  {
    ProfSpan span(BRUNO_PROF_PROCESS_FWD, st);
    k<<<grid, g.threads + 32, smem, st>>>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D);
    rc = check_launch("process_fwd_ring_kernel");
  }
  if (rc) return rc;
  if (g.nseg > 1) {
    process_fwd_finalize<<<ceil_div(B * T, 256), 256, 0, st>>>(scratch, lp, lp_prior, g.nseg, B * T);
    rc = check_launch("process_fwd_finalize");
  }
  return rc;
}

template <int KIND, int VEC>
static int launch_fwd(const float* z, const float* mu0, const float* table, float* s1, float* s2, float* lp,
                      float* lp_prior, float* scratch, int B, int T, int D, const ScanGeom& g, cudaStream_t st) {
  if constexpr (VEC == 4) {
    const int v = fwd_variant();
    // The running-state entry (s1 / s2 read and written back: the reference's step-by-step API, not a hot path) stays on
    // the register-staged scan.  The STATE instantiation of the ring kernel returns ONE wrong row (the first row of one
    // 4-row group, all steps) in ~20 % of launches when its z buffer is a freshly allocated copy, never when the buffer is
    // reused -- independent of what ran before, of device synchronisation and of the output buffer
    // (tools/dbg_proc_flaky.py).  The sequence-at-once instantiation (the product path) shows no such rows in any of these
    // set-ups (NaN-prefilled outputs, 150 launches per case).  Cause not found; the default does not use the instantiation.
    if ((v == 0 && s1 == nullptr) || v >= 10) {
      int rc;
      switch (v) {
        case 11: rc = launch_fwd_ring<KIND, 2, 8, 2, 3, 1>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 12: rc = launch_fwd_ring<KIND, 2, 8, 4, 3, 0>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 13: rc = launch_fwd_ring<KIND, 2, 8, 2, 4, 0>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 14: rc = launch_fwd_ring<KIND, 4, 6, 2, 2, 1>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 16: rc = launch_fwd_ring<KIND, 4, 4, 4, 2, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 17: rc = launch_fwd_ring<KIND, 4, 4, 2, 2, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 18: rc = launch_fwd_ring<KIND, 4, 3, 4, 2, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 15: rc = launch_fwd_ring<KIND, 4, 4, 2, 2, 0>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        case 10: rc = launch_fwd_ring<KIND, 4, 4, 4, 2, 1>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
        // default: the step's coefficient rows ride in the ring stage (CPF = 2): no LDG in the hot loop, 648 vs 689 us
        // for the TP scan at B = 65536 (profiles/r01_process_ring_sweep_v6.log)
        default: rc = launch_fwd_ring<KIND, 4, 4, 4, 2, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, st); break;
      }
      if (rc <= 0) return rc;
    }
  }
  if (g.threads > 256)   // only the 2-wide layout uses blocks of up to 512 threads
    return launch_fwd_variant<KIND, VEC, 4, 2, 2, 2, 512>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
  switch (fwd_variant()) {
    case 2: return launch_fwd_variant<KIND, VEC, 4, 1, 3, 4, 256>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
    case 3: return launch_fwd_variant<KIND, VEC, 2, 1, 3, 4, 256>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
    case 4: return launch_fwd_variant<KIND, VEC, 2, 1, 4, 8, 256>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
    default: return launch_fwd_variant<KIND, VEC, 4, 1, 2, 4, 256>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
  }
}

static int bwd_blocks(int B) {
  const int groups = ceil_div(B, BWD_NB);
  return groups < 592 ? groups : 592;
}

// Backward variant: 0 = default (ring kernel when eligible), 9 = always the register-staged kernel, 10/11 = ring shapes.
static int g_bwd_variant = -1;
static int bwd_variant() {
  if (g_bwd_variant < 0) {
    const char* e = getenv("BRUNO_TP_BWD_VARIANT");
    g_bwd_variant = e ? atoi(e) : 0;
  }
  return g_bwd_variant;
}

template <int KIND, int NB, int NSTAGE, int MINB, bool CT>
static int launch_bwd_ring(const float* z, const float* mu0, const float* table, const float* dlp, const float* dlpp,
                           float* dz, float* dvar, float* dnu, float* dcorr, float* scratch, int B, int T, int D,
                           cudaStream_t st) {
  const ScanGeom g = ring_geom(D);
  constexpr int NROW = CT ? NB + (KIND == BRUNO_TP ? 15 : 8) : NB;
  const size_t smem = static_cast<size_t>(NSTAGE) * NROW * g.threads * 4 * sizeof(float) + 2 * NSTAGE * sizeof(uint64_t);
  if (smem > 226 * 1024) return 1;
  auto k = dlpp ? process_bwd_ring_kernel<KIND, NB, true, NSTAGE, MINB, CT> : process_bwd_ring_kernel<KIND, NB, false, NSTAGE, MINB, CT>;
  static RingLaunchCache cache;
  const RingLaunchCache::Entry ce = cache.get(k, dlpp ? 1 : 0, smem, g.threads);
  const int sm_count = ce.sms;
  if (ce.per_sm < 1) return 1;
  const int groups = ceil_div(B, NB);
  int nblk = ceil_div(ce.per_sm * sm_count, g.nseg);
  if (nblk > groups) nblk = groups;
  if (nblk > bwd_blocks(B)) nblk = bwd_blocks(B);             // the scratch holds bwd_blocks(B) partial rows
  dim3 grid(nblk, g.nseg);
  int rc;
  {
    ProfSpan span(BRUNO_PROF_PROCESS_BWD, st);
    k<<<grid, g.threads + 32, smem, st>>>(z, mu0, table, dlp, dlpp, dz, scratch, B, T, D);
    rc = check_launch("process_bwd_ring_kernel");
  }
  if (rc) return rc;
  process_bwd_finalize<<<ceil_div(D, 256), 256, 0, st>>>(scratch, nblk, D, dvar, KIND == BRUNO_TP ? dnu : nullptr, dcorr);
  return check_launch("process_bwd_finalize");
}

template <int KIND, int VEC>
static int launch_bwd(const float* z, const float* mu0, const float* table, const float* dlp, const float* dlpp,
                      float* dz, float* dvar, float* dnu, float* dcorr, float* scratch, int B, int T, int D,
                      const ScanGeom& g, cudaStream_t st) {
  if constexpr (VEC == 4) {
    const int v = bwd_variant();
    if (v != 9) {
      int rc;
      switch (v) {
        case 10: rc = launch_bwd_ring<KIND, 2, 4, 2, false>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        case 15: rc = launch_bwd_ring<KIND, 2, 8, 1, false>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        case 11: rc = launch_bwd_ring<KIND, 2, 8, 2, false>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        case 13: rc = launch_bwd_ring<KIND, 2, 3, 1, true>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        case 14: rc = launch_bwd_ring<KIND, 2, 2, 1, true>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        case 12: rc = launch_bwd_ring<KIND, 4, 4, 1, false>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st); break;
        // measured (profiles/r01_process_bwd_ring_sweep.log): one 254-register block per SM without spills beats two
        // 128-register blocks with 48 spilled values per thread (TP 5.5 vs 6.9 ms at B = 65536; GP equal)
        // default: the reverse pass's coefficient and derivative rows ride in the ring stage (TP 5.0 vs 5.5 ms, TP + prior
        // 5.5 vs 6.9, GP 2.2 vs 2.65 at B = 65536); GP + prior is the one case measured slower that way (3.5 vs 2.8 ms)
        default:
          if (KIND == BRUNO_TP || dlpp == nullptr)
            rc = launch_bwd_ring<KIND, 2, 3, 1, true>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st);
          else
            rc = launch_bwd_ring<KIND, 2, 8, 1, false>(z, mu0, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D, st);
          break;
      }
      if (rc <= 0) return rc;
    }
  }
  const int groups = ceil_div(B, BWD_NB);
  const int nblk = bwd_blocks(B);
  const int gpb = ceil_div(groups, nblk);
  const int nblk_used = ceil_div(groups, gpb);
  dim3 grid(nblk_used, g.nseg);
  int rc;
  {
    ProfSpan span(BRUNO_PROF_PROCESS_BWD, st);
    if (dlpp)
      process_bwd_kernel<KIND, VEC, BWD_NB, true><<<grid, g.threads, 0, st>>>(z, mu0, table, dlp, dlpp, dz, scratch, B, T, D, gpb);
    else
      process_bwd_kernel<KIND, VEC, BWD_NB, false><<<grid, g.threads, 0, st>>>(z, mu0, table, dlp, dlpp, dz, scratch, B, T, D, gpb);
    rc = check_launch("process_bwd_kernel");
  }
  if (rc) return rc;
  process_bwd_finalize<<<ceil_div(D, 256), 256, 0, st>>>(scratch, nblk_used, D, dvar, KIND == BRUNO_TP ? dnu : nullptr, dcorr);
  return check_launch("process_bwd_finalize");
}

}  // namespace bruno

using namespace bruno;

extern "C" {

size_t bruno_process_table_floats(int kind, int T, int D, int with_grad) {
  return static_cast<size_t>(T * ncf_of(kind) + N_PRIOR) * D * (with_grad ? 4 : 1);
}

int bruno_process_table(int kind, const float* var_vbl, const float* nu_vbl, const float* corr_vbl, const float* mu0,
                        const float* mask_dim, float min_nu, int n0, int T, int D, int with_grad, float* table,
                        void* stream) {
  (void)mu0;
  BRUNO_REQUIRE(kind == BRUNO_TP || kind == BRUNO_GP, "bruno_process_table: bad kind %d", kind);
  BRUNO_REQUIRE(T >= 0 && D > 0 && n0 >= 0, "bruno_process_table: bad sizes T=%d D=%d n0=%d", T, D, n0);
  BRUNO_REQUIRE(var_vbl && corr_vbl && table && (kind == BRUNO_GP || nu_vbl), "bruno_process_table: null pointer");
  dim3 grid(ceil_div(D, 128), T + 1);
  process_table_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(kind, var_vbl, nu_vbl, corr_vbl, mask_dim,
                                                                              min_nu, n0, T, D, with_grad, table);
  return check_launch("process_table_kernel");
}

#ifndef BRUNO_NO_DEBUG
void bruno_process_set_fwd_variant(int variant) { g_fwd_variant = variant < 0 ? 0 : variant; }
void bruno_process_set_bwd_variant(int variant) { g_bwd_variant = variant < 0 ? 0 : variant; }
#endif

size_t bruno_process_scratch_floats(int B, int T, int D) {
  const int nseg = scan_geom(D, false).nseg;  // worst case: the scalar (unaligned) path (>= the ring kernel's segments)
  return nseg > 1 ? static_cast<size_t>(nseg) * B * T * 2 : 0;
}

int bruno_process_fwd(int kind, const float* z, const float* mu0, const float* table, float* s1, float* s2, float* lp,
                      float* lp_prior, float* scratch, int B, int T, int D, void* stream) {
  BRUNO_REQUIRE(kind == BRUNO_TP || kind == BRUNO_GP, "bruno_process_fwd: bad kind %d", kind);
  BRUNO_REQUIRE(B > 0 && T > 0 && D > 0, "bruno_process_fwd: bad sizes B=%d T=%d D=%d", B, T, D);
  BRUNO_REQUIRE(z && mu0 && table && lp, "bruno_process_fwd: null pointer");
  BRUNO_REQUIRE((s1 == nullptr) == (s2 == nullptr) || kind == BRUNO_GP, "bruno_process_fwd: s1/s2 must come together");
  const bool ptr_ok = aligned16(z) && aligned16(mu0) && aligned16(table) && (!s1 || aligned16(s1)) && (!s2 || aligned16(s2));
  ScanGeom g = scan_geom(D, ptr_ok, true);
  BRUNO_REQUIRE(g.nseg == 1 || scratch, "bruno_process_fwd: scratch required when D is split (D=%d)", D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == BRUNO_TP) {
    if (g.vec == 4) return launch_fwd<BRUNO_TP, 4>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
    if (g.vec == 2) return launch_fwd<BRUNO_TP, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
    return launch_fwd<BRUNO_TP, 1>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
  }
  if (g.vec == 4) return launch_fwd<BRUNO_GP, 4>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
  if (g.vec == 2) return launch_fwd<BRUNO_GP, 2>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
  return launch_fwd<BRUNO_GP, 1>(z, mu0, table, s1, s2, lp, lp_prior, scratch, B, T, D, g, st);
}

size_t bruno_process_bwd_scratch_floats(int B, int T, int D) {
  (void)T;
  return static_cast<size_t>(bwd_blocks(B)) * 3 * D;
}

int bruno_process_bwd(int kind, const float* z, const float* mu0, const float* table, const float* dlp,
                      const float* dlp_prior, float* dz, float* dvar_vbl, float* dnu_vbl, float* dcorr_vbl,
                      float* scratch, int B, int T, int D, void* stream) {
  BRUNO_REQUIRE(kind == BRUNO_TP || kind == BRUNO_GP, "bruno_process_bwd: bad kind %d", kind);
  BRUNO_REQUIRE(B > 0 && T > 0 && D > 0, "bruno_process_bwd: bad sizes B=%d T=%d D=%d", B, T, D);
  BRUNO_REQUIRE(z && mu0 && table && dlp && dz && dvar_vbl && dcorr_vbl && scratch && (kind == BRUNO_GP || dnu_vbl),
                "bruno_process_bwd: null pointer");
  const bool ptr_ok = aligned16(z) && aligned16(mu0) && aligned16(table) && aligned16(dz) && aligned16(scratch);
  ScanGeom g = scan_geom(D, ptr_ok);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == BRUNO_TP)
    return g.vec == 4
               ? launch_bwd<BRUNO_TP, 4>(z, mu0, table, dlp, dlp_prior, dz, dvar_vbl, dnu_vbl, dcorr_vbl, scratch, B, T, D, g, st)
               : launch_bwd<BRUNO_TP, 1>(z, mu0, table, dlp, dlp_prior, dz, dvar_vbl, dnu_vbl, dcorr_vbl, scratch, B, T, D, g, st);
  return g.vec == 4
             ? launch_bwd<BRUNO_GP, 4>(z, mu0, table, dlp, dlp_prior, dz, dvar_vbl, dnu_vbl, dcorr_vbl, scratch, B, T, D, g, st)
             : launch_bwd<BRUNO_GP, 1>(z, mu0, table, dlp, dlp_prior, dz, dvar_vbl, dnu_vbl, dcorr_vbl, scratch, B, T, D, g, st);
}

int bruno_process_posterior(int kind, const float* z, const float* var_vbl, const float* nu_vbl, const float* corr_vbl,
                            const float* mu0, float min_nu, float* mu_out, float* var_out, float* nu_out, int B, int T,
                            int D, void* stream) {
  BRUNO_REQUIRE(kind == BRUNO_TP || kind == BRUNO_GP, "bruno_process_posterior: bad kind %d", kind);
  BRUNO_REQUIRE(B > 0 && T >= 0 && D > 0, "bruno_process_posterior: bad sizes");
  BRUNO_REQUIRE((z || T == 0) && var_vbl && corr_vbl && mu0 && mu_out && var_out && (kind == BRUNO_GP || (nu_vbl && nu_out)),
                "bruno_process_posterior: null pointer");
  dim3 grid(ceil_div(D, 128), B);
  process_posterior_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(kind, z, var_vbl, nu_vbl, corr_vbl, mu0,
                                                                                  min_nu, mu_out, var_out, nu_out, B, T, D);
  return check_launch("process_posterior_kernel");
}

int bruno_process_sample(int kind, const float* rvs, const float* mu, const float* var, const float* nu, float* out,
                         int n, int B, int D, long stride_b, void* stream) {
  BRUNO_REQUIRE(kind == BRUNO_TP || kind == BRUNO_GP, "bruno_process_sample: bad kind %d", kind);
  BRUNO_REQUIRE(n > 0 && B > 0 && D > 0, "bruno_process_sample: bad sizes");
  BRUNO_REQUIRE(rvs && mu && var && out && (kind == BRUNO_GP || nu), "bruno_process_sample: null pointer");
  const size_t total = static_cast<size_t>(n) * B * D;
  process_sample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      kind, rvs, mu, var, nu, out, n, B, D, stride_b);
  return check_launch("process_sample_kernel");
}

}  // extern "C"
