// Library-wide runtime pieces of libsln_b200.so: thread-local error string, launch accounting and the optional
// CUDA-event profiler that bench.py uses to time each kernel class inside a real step (include/sln_b200.h).
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/sln_b200.h"
#include "common.cuh"

namespace sln {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfRec { cudaEvent_t a, b; int cls; double work; };
struct Prof {
  std::mutex mu;
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
Prof g_prof;
}  // namespace

bool prof_enabled() { return g_prof.on; }
void prof_begin(cudaStream_t st, int cls, double work) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  ProfRec r; r.a = g_prof.get(); r.b = g_prof.get(); r.cls = cls; r.work = work;
  cudaEventRecord(r.a, st);
  g_prof.recs.push_back(r);
}
void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.recs.empty()) cudaEventRecord(g_prof.recs.back().b, st);
}

}  // namespace sln

using namespace sln;

extern "C" {

int sln_version(void) { return SLN_ABI_VERSION; }
const char* sln_last_error(void) { return get_error(); }
int64_t sln_launch_count(void) { return (int64_t)g_launches.load(); }

int sln_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (auto& r : g_prof.recs) { g_prof.pool.push_back(r.a); g_prof.pool.push_back(r.b); }
  g_prof.recs.clear();
  g_prof.on = on != 0;
  return SLN_OK;
}

int sln_prof_read(int cls, double* ms, double* work, int64_t* launches) {
  SLN_CHECK_ARG(cls >= 0 && cls < PROF_NUM && ms && work && launches, "bad profiler class / null output");
  std::lock_guard<std::mutex> lk(g_prof.mu);
  double t = 0.0, w = 0.0; int64_t n = 0;
  for (auto& r : g_prof.recs) {
    if (r.cls != cls) continue;
    SLN_CUDA_TRY(cudaEventSynchronize(r.b));
    float x = 0.f;
    SLN_CUDA_TRY(cudaEventElapsedTime(&x, r.a, r.b));
    t += x; w += r.work; ++n;
  }
  *ms = t; *work = w; *launches = n;
  return SLN_OK;
}

}  // extern "C"
