// Launch-to-launch period of a chain of small dependent kernels replayed from a CUDA graph, with and without
// programmatic dependent launch (griddepcontrol): what converting the helper kernels of a step to PDL would buy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_bench tools/pdl_bench.cu && tools/pdl_bench
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void small_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * 1.0001f + 1.f;
}

template <bool PDL>
float run(int blocks, int n, float* a, float* b, int chain, int reps) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  auto launch = [&](int i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = PDL ? 1 : 0;
    const float* src = (i & 1) ? b : a;
    float* dst = (i & 1) ? a : b;
    cudaLaunchKernelEx(&cfg, small_kernel<PDL>, src, dst, n);
  };
  for (int i = 0; i < chain; ++i) launch(i);
  cudaStreamSynchronize(st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) launch(i);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (reps * chain);
}

int main() {
  const int n = 640 * 784;
  float *a, *b;
  cudaMalloc(&a, n * sizeof(float));
  cudaMalloc(&b, n * sizeof(float));
  cudaMemset(a, 0, n * sizeof(float));
  for (int blocks : {8, 148, (n + 255) / 256}) {
    const float plain = run<false>(blocks, n, a, b, 200, 20);
    const float pdl = run<true>(blocks, n, a, b, 200, 20);
    printf("chain of 200 kernels, %5d blocks x 256 threads (640x784 floats): %.2f us per launch plain, %.2f us with PDL\n", blocks,
           plain, pdl);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
