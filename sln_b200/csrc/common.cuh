// Common device/host helpers for the 3D_SLN B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "sln_b200 kernels are written for sm_100a only"
#endif

namespace sln {

// ---------------------------------------------------------------- error plumbing
// Thread-local last-error string; C-ABI entry points return 0 / negative SLN_E* / cudaError_t.
void set_error(const char* fmt, ...);
const char* get_error();

#define SLN_OK 0
#define SLN_EINVAL (-1)      // bad argument (shape / alignment / null pointer)
#define SLN_EWORKSPACE (-2)  // workspace too small
#define SLN_EUNSUPPORTED (-3)
#define SLN_ECUDA (-4)       // CUDA runtime error; message holds cudaGetErrorString

#define SLN_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::sln::set_error(__VA_ARGS__);             \
      return SLN_EINVAL;                         \
    }                                            \
  } while (0)

#define SLN_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::sln::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                       __FILE__, __LINE__);                                       \
      return SLN_ECUDA;                                                           \
    }                                                                             \
  } while (0)

#define SLN_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != SLN_OK) return _r; \
  } while (0)

// Launch accounting + optional per-class CUDA-event profiler (runtime.cu).  Both are host-side only.
void count_launch();
enum ProfClass { PROF_GEMM_FWD = 0, PROF_GEMM_BWD_X = 1, PROF_GEMM_BWD_W = 2, PROF_POOL = 3, PROF_PREP = 4, PROF_MISC = 5,
                 PROF_RASTER_FWD = 6, PROF_RASTER_BWD = 7, PROF_SPADE_CONV = 8, PROF_SPADE_MISC = 9, PROF_NUM = 10 };
bool prof_enabled();
void prof_begin(cudaStream_t st, int cls, double work);
void prof_end(cudaStream_t st);
struct ProfScope {  // wraps one launch (or a short launch group) in an event pair when profiling is on
  cudaStream_t st; bool on;
  ProfScope(cudaStream_t s, int cls, double work) : st(s), on(prof_enabled()) { if (on) prof_begin(st, cls, work); }
  ~ProfScope() { if (on) prof_end(st); }
};

inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return SLN_ECUDA;
  }
  return SLN_OK;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- bump allocator over caller workspace
// cudaFuncSetAttribute is per DEVICE: `mask` (one static per call site / template instantiation) remembers which devices of this
// process have been configured.  Returns true when the current device still needs the attribute set.
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool dry;  // size query only
  Arena(void* b, size_t c) : base((char*)b), cap(c), off(0), dry(b == nullptr) {}
  template <class T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T), 256);
    size_t o = off;
    off += bytes;
    if (dry) return (T*)(uintptr_t)(256 + o);  // non-null fake address (never dereferenced)
    return (T*)(base + o);
  }
  bool ok() const { return dry || off <= cap; }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 make4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// fire-and-forget fp32 reduction into global memory (SASS: RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
#endif

}  // namespace sln
