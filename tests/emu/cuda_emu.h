// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: run the .cu kernel sources on the CPU.
//
// There is no GPU in the build container, so the `-m "not gpu"` test-suite compiles the very same
// kernel sources with g++ against this ~300-line shim and checks indexing / halo / scan logic against
// the oracle before any GPU minute is spent.  Each CUDA block runs on one OS thread; its CUDA threads
// are ucontext fibers scheduled round-robin, so __syncthreads(), __syncwarp() and warp shuffles have
// their real semantics.  The product never loads the library built from this header
// (gcmiipy_b200/_lib.py refuses unless GCMIIPY_B200_EMULATE=1 is set by the tests).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
// a block runs on ONE OS thread (its CUDA threads are fibers of it), so block-shared = thread_local
#define __shared__ static thread_local

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }


// ---- atomics and bit casts (blocks run on several OS threads, so these are real atomics) ----------
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val) {
  __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline unsigned long long atomicExch(unsigned long long* a, unsigned long long val) {
  return __atomic_exchange_n(a, val, __ATOMIC_SEQ_CST);
}
static inline double atomicAdd(double* a, double v) {
  unsigned long long* u = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old = __atomic_load_n(u, __ATOMIC_SEQ_CST);
  for (;;) {
    double d; memcpy(&d, &old, 8); d += v;
    unsigned long long nu; memcpy(&nu, &d, 8);
    if (__atomic_compare_exchange_n(u, &old, nu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { memcpy(&d, &old, 8); return d; }
  }
}
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline long long __double_as_longlong(double d) { long long x; memcpy(&x, &d, 8); return x; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
using std::isinf;
using std::isnan;

template <class T> static inline T __ldg(const T* p) { return *p; }

namespace gcm_emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = 0;  // 0 runnable, 1 waiting at block barrier, 2 waiting at warp barrier, 3 done
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = 0;
  int nthreads = 0;
  char* smem = nullptr;
  const std::function<void()>* body = nullptr;
  alignas(16) unsigned char shfl[64][32][16];  // per warp, per lane exchange slots
};

extern thread_local Block* tl_block;
extern thread_local uint3 tl_threadIdx, tl_blockIdx;
extern thread_local dim3 tl_blockDim, tl_gridDim;

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
void yield_state(int state);

static inline void* dyn_smem() { return tl_block->smem; }
static inline int lane_id() { return tl_block->cur & 31; }
static inline int warp_id() { return tl_block->cur >> 5; }

}  // namespace gcm_emu

#define threadIdx (gcm_emu::tl_threadIdx)
#define blockIdx (gcm_emu::tl_blockIdx)
#define blockDim (gcm_emu::tl_blockDim)
#define gridDim (gcm_emu::tl_gridDim)

static inline void __syncthreads() { gcm_emu::yield_state(1); }
static inline void __syncwarp(unsigned = 0xffffffffu) { gcm_emu::yield_state(2); }

template <class T> static inline T emu_shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  gcm_emu::Block* b = gcm_emu::tl_block;
  int w = gcm_emu::warp_id(), l = gcm_emu::lane_id();
  memcpy(b->shfl[w][l], &v, sizeof(T));
  gcm_emu::yield_state(2);
  T r = v;
  int first = w * 32;
  if (src_lane >= 0 && src_lane < 32 && first + src_lane < b->nthreads) memcpy(&r, b->shfl[w][src_lane], sizeof(T));
  gcm_emu::yield_state(2);
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int l = gcm_emu::lane_id();
  return emu_shfl_from(v, (l / width) * width + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  (void)width;
  return emu_shfl_from(v, gcm_emu::lane_id() ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int l = gcm_emu::lane_id();
  int s = l + (int)d;
  return emu_shfl_from(v, (s / width == l / width) ? s : l);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int l = gcm_emu::lane_id();
  int s = l - (int)d;
  return emu_shfl_from(v, (s >= 0 && s / width == l / width) ? s : l);
}
