// Shared host-side plumbing for the C-ABI: error codes, last-error string, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "../../include/bruno_sm100.h"

namespace bruno {

void set_last_error(const char* fmt, ...);

extern std::atomic<long> g_launches;

// RAII CUDA-event span around one kernel launch (no-op unless bruno_profile_enable(1))
class ProfSpan {
 public:
  ProfSpan(int kind, cudaStream_t st);
  ~ProfSpan();

 private:
  int kind_;
  cudaStream_t st_;
  bool active_;
  void *begin_ = nullptr, *end_ = nullptr;
};

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return BRUNO_ECUDA;
  }
  return BRUNO_OK;
}

#define BRUNO_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      bruno::set_last_error(__VA_ARGS__); \
      return BRUNO_EINVAL;            \
    }                                 \
  } while (0)

// Runs `f` once per (call site, CUDA device), under a lock: kernel attributes such as the dynamic-shared-memory opt-in are
// per device, and two host threads (e.g. the autograd thread and the main thread) may reach a call site at the same time.
class OncePerDevice {
 public:
  template <class F>
  void run(F f) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0;
    if (bit && (done_.load(std::memory_order_acquire) & bit)) return;
    std::lock_guard<std::mutex> lock(mu_);
    if (bit && (done_.load(std::memory_order_relaxed) & bit)) return;
    f();
    done_.fetch_or(bit, std::memory_order_release);
  }

 private:
  std::atomic<uint64_t> done_{0};
  std::mutex mu_;
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace bruno
