// Process-wide plumbing of the C ABI: last-error string, launch counter, optional per-kernel-family CUDA-event timing
// (used by bench.py to measure the live duration of the dominant kernel inside the timed region).
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace bruno {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

std::atomic<long> g_launches{0};

struct ProfState {
  bool on = false;
  std::mutex mu;
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[BRUNO_PROF_KINDS];
  size_t next = 0;
};
static ProfState g_prof;

static cudaEvent_t prof_event() {
  if (g_prof.next == g_prof.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof.pool.push_back(e);
  }
  return g_prof.pool[g_prof.next++];
}

ProfSpan::ProfSpan(int kind, cudaStream_t st) : kind_(kind), st_(st), active_(false) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  begin_ = prof_event();
  end_ = prof_event();
  active_ = true;
  cudaEventRecord(static_cast<cudaEvent_t>(begin_), st_);
}
ProfSpan::~ProfSpan() {
  if (!active_) return;
  cudaEventRecord(static_cast<cudaEvent_t>(end_), st_);
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.spans[kind_].emplace_back(static_cast<cudaEvent_t>(begin_), static_cast<cudaEvent_t>(end_));
}

}  // namespace bruno

using namespace bruno;

extern "C" {

const char* bruno_last_error(void) { return g_err; }
int bruno_abi_version(void) { return 1; }
long bruno_launch_count(void) { return g_launches.load(); }

int bruno_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.on = on != 0;
  for (int k = 0; k < BRUNO_PROF_KINDS; ++k) g_prof.spans[k].clear();
  g_prof.next = 0;
  return BRUNO_OK;
}

int bruno_profile_read(int kind, double* total_ms, long* launches) {
  BRUNO_REQUIRE(kind >= 0 && kind < BRUNO_PROF_KINDS && total_ms && launches, "bruno_profile_read: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof.mu);
  double tot = 0.0;
  for (auto& s : g_prof.spans[kind]) {
    cudaEventSynchronize(s.second);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) tot += ms;
  }
  *total_ms = tot;
  *launches = static_cast<long>(g_prof.spans[kind].size());
  return BRUNO_OK;
}

}  // extern "C"
