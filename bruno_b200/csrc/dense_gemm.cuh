// Shared declarations of the dense-path GEMMs (dense.cu: CUDA-core SGEMM; gemm_tf32.cu: tcgen05 3xTF32 GEMM).
#pragma once
#include "common.cuh"

namespace bruno {

// fused epilogue / prologue of   C[M,N] = op(A)[M,K] * op(B)[K,N]
struct GemmEpi {
  const float* col_scale;  // out = acc * col_scale[n]          (may be null)
  const float* bias;       // + bias[n]                         (may be null)
  float* pre_out;          // store the pre-activation here too (may be null)
  int act;                 // 0 none, 1 ELU, 2 leaky ReLU (0.2), 3 tanh
  int accumulate;          // C += result
  const float* a_kscale;   // A[m,k] *= a_kscale[k]             (may be null)
};

__device__ __forceinline__ float gemm_epilogue_value(float v, int n, const GemmEpi& ep, float* pre_slot) {
  if (ep.col_scale) v *= ep.col_scale[n];
  if (ep.bias) v += ep.bias[n];
  if (pre_slot) *pre_slot = v;
  if (ep.act == 1) v = v > 0.f ? v : expm1f(v);
  else if (ep.act == 2) v = v > 0.f ? v : 0.2f * v;   // tf.nn.leaky_relu default alpha
  else if (ep.act == 3) v = tanhf(v);
  return v;
}

// Two GEMMs of the same shape and flags in ONE launch of the tcgen05 kernel:
//   mode 1 (side by side): C  = op(A) op(B)  + epilogue(ep),   C2 = op(A2) op(B2) + epilogue(ep with bias2)
//                          -- the second problem takes the upper half of gridDim.x (the s / t heads, dKs / dKt);
//   mode 2 (K-concatenated): C = op(A) op(B) + op(A2) op(B2)  -- the second product takes the upper half of the split-K
//                          slices, the cluster reduction adds all of them in order (dh = ds Ks^T + dt Kt^T).
struct GemmPair {
  int mode;            // 0: single GEMM
  const float* A2;
  const float* B2;
  float* C2;           // mode 1
  const float* bias2;  // mode 1
};

// tcgen05 path.  `cluster`: the split-K CTAs of a tile are one thread-block cluster and reduce through peer shared memory
// (needs no workspace).  Otherwise / if the cluster launch is refused: `ws` / `tickets` = split-K workspace (>= ws_floats
// floats) and zeroed ticket counters (>= max_tiles), both may be null (un-split then).
// Returns BRUNO_OK, or a negative code; *taken = 0 when the shape is not eligible (caller falls back to the SIMT kernel).
int gemm_tf32x3(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                const GemmEpi& ep, float* ws, size_t ws_floats, unsigned* tickets, int max_tiles, int sms, bool cluster,
                cudaStream_t st, int* taken, const GemmPair* pair = nullptr);

void gemm_tf32x3_set_trace(long long* device_buffer);
// launch geometry of the cluster path for a shape: out = {BN, gridDim.x, gridDim.y, gridDim.z (0: the pair cannot run as one
// launch), K tiles per split-K slice, cluster size}
void gemm_tf32x3_plan(int M, int N, int K, int pair_mode, int sms, int out[6]);

}  // namespace bruno
