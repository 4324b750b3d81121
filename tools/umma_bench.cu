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

// The MMA stream of the line-tile convolution kernel: per tile 3 x N=64 (first step, block by block) + 11 x N=192 into
// accumulator slots (slot, slot+1, slot+2) of a ring of 8 (64 columns each); `dstep` = slot advance per tile (1 = the real
// kernel, 0 = always the same columns), `dalign` = first slot.
__global__ void __launch_bounds__(128, 1)
line_kernel(int niter, int dstep, int dalign, int split_first, long long* cycles, int var = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 128);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); mbar_init(bar + 3, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t id64 = umma_idesc_bf16(128, 64, 0, 0), id128 = umma_idesc_bf16(128, 128, 0, 0), id192 = umma_idesc_bf16(128, 192, 0, 0);
    const uint64_t a0 = umma_desc_sw128(smem_u32(smem) + 1024, 16, 1024), b0 = umma_desc_sw128(smem_u32(smem + 96 * 1024), 16, 1024);
    long long t0 = clock64();
    int sl = dalign;
    for (int it = 0; it < niter; ++it) {
      const uint64_t ad = a0 + ((var & 2) ? 0u : static_cast<uint32_t>(((it & 3) * 18432) >> 4));
      const int n0 = min(3, 8 - sl);
      if (split_first) {
#pragma unroll
        for (int b = 0; b < 3; ++b) umma_bf16(tmem + ((sl + b) & 7) * 64, ad - 8u, b0 + b * 512u, id64, b < 2);
      }
#pragma unroll
      for (int st = split_first ? 1 : 0; st < 12; ++st) {
        const int dxi = st >> 2, k = st & 3;
        const uint64_t a = ad + static_cast<uint64_t>(static_cast<int64_t>(((var & 1) ? 0 : (dxi - 1) * 8) + ((var & 8) ? 0 : 2 * k)));
        const uint32_t bo = (((var & 4) ? 0 : dxi * 3 * 8192) + ((var & 8) ? 0 : k * 32)) >> 4;
        if ((var & 32) && st == 6) { while (!mbar_try_wait(bar + 2, 1)) {} }
        if ((var & 128) && (st == 3 || st == 9)) { while (!mbar_try_wait(bar + 2, 1)) {} }
        if (n0 == 3) umma_bf16(tmem + sl * 64, a, b0 + bo, id192, 1);
        else {
          umma_bf16(tmem + sl * 64, a, b0 + bo, n0 == 2 ? id128 : id64, 1);
          umma_bf16(tmem, a, b0 + bo + n0 * 512u, n0 == 2 ? id64 : id128, 1);
        }
      }
      if (var & 16) { umma_commit(bar + 1); umma_commit(bar + 3); }
      if (var & 64) tc_fence_after();
      sl = (sl + dstep) & 7;
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
  {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(line_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
    struct { int dstep, dalign, split; const char* name; } lc[] = {
        {1, 0, 1, "line-tile stream: slot ring (advance 1 per tile), first step block by block"},
        {1, 0, 0, "line-tile stream: slot ring, 12 x N=192 (no first-step split)"},
        {0, 0, 1, "same accumulator columns 0..191 every tile, first step split"},
        {0, 0, 0, "same accumulator columns 0..191 every tile, 12 x N=192"},
        {0, 1, 0, "columns 64..255 every tile, 12 x N=192"},
        {0, 2, 0, "columns 128..319 every tile, 12 x N=192"},
        {0, 4, 0, "columns 256..447 every tile, 12 x N=192"},
        {4, 0, 0, "alternating columns 0..191 / 256..447, 12 x N=192"},
    };
    for (auto& c : lc) {
      line_kernel<<<148, 128, 204 * 1024>>>(400, c.dstep, c.dalign, c.split, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("%-80s %7.1f cycles/tile %s\n", c.name, double(cyc) / 400.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    for (int var : {0, 16, 32, 64, 128, 16 + 32, 16 + 32 + 64, 16 + 128 + 64}) {
      line_kernel<<<148, 128, 204 * 1024>>>(400, 1, 0, 1, d, var);
      cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("line-tile stream +%s%s%s%s  %7.1f cycles/tile\n", (var & 16) ? " 2 commits/tile" : "", (var & 32) ? " 1 try_wait mid-tile" : "",
             (var & 128) ? " 2 try_waits mid-tile" : "", (var & 64) ? " fence::after_thread_sync" : "", double(cyc) / 400.0);
    }
  }
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
      {192, 0, 0, 1, "K-major A (shift 1 row) x K-major B, N=192      [line-tile conv]"},
      {192, 0, 0, 29, "K-major A x K-major B, N=192"},
      {96, 0, 0, 29, "K-major A x K-major B, N=96"},
      {160, 0, 0, 29, "K-major A x K-major B, N=160"},
      {192, 1, 1, 29, "MN-major A x MN-major B, N=192"},
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
