// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, SS mode) for the operand layouts the
// conv kernels use.  One CTA per SM, operands resident in shared memory, NISSUE MMAs back to back, one commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../bruno_b200/csrc/ptx.cuh"
using namespace bruno;

__global__ void __launch_bounds__(128, 1)
bench_kernel(int N, int a_mn, int b_mn, int niter, int shift_rows, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 64);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < niter; ++it) {
      // 36 MMAs like one conv tile: 9 "taps" x 4 k-steps, A shifted by rows per tap
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t aa = a0 + tap * shift_rows * 128 + (a_mn ? k * 2048 : k * 32);
          uint32_t bb = b0 + tap * 8192 + (b_mn ? k * 2048 : k * 32);
          umma_bf16(tmem + (it & 1) * 256, umma_desc_sw128(aa, a_mn ? 3712 : 16, 1024), umma_desc_sw128(bb, 16, 1024), idesc, (tap | k) != 0);
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
  const int niter = 200;
  struct { int N, a, b, shift; const char* name; } cases[] = {
      {64, 0, 0, 29, "K-major A (row-shifted taps) x K-major B, N=64   [conv fwd/dgrad]"},
      {64, 0, 0, 0, "K-major A (no shift)           x K-major B, N=64"},
      {128, 0, 0, 29, "K-major A x K-major B, N=128"},
      {256, 0, 0, 29, "K-major A x K-major B, N=256"},
      {64, 1, 1, 29, "MN-major A x MN-major B, N=64                    [conv wgrad]"},
      {128, 1, 1, 29, "MN-major A x MN-major B, N=128"},
      {32, 0, 0, 29, "K-major A x K-major B, N=32"},
  };
  for (auto& c : cases) {
    for (int grid : {1, 148}) {
      bench_kernel<<<grid, 128, 204 * 1024>>>(c.N, c.a, c.b, niter, c.shift, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      double per = double(cyc) / (niter * 36.0);
      printf("%-70s grid %3d: %7.1f cycles/MMA  (ideal %5.1f)  -> %5.1f%% of tensor peak  %s\n", c.name, grid, per,
             128.0 * c.N / 256.0, 100.0 * (128.0 * c.N / 256.0) / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
