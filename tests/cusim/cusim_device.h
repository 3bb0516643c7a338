// cusim_device.h -- device-side half of the CPU SIMT emulator used by the CPU test suite.
//
// TEST INFRASTRUCTURE ONLY.  tests/cusim compiles the product's CUDA sources (csrc/nd_b200.cu + nd_b200_kernels.cuh) with
// g++ against these headers so that `pytest -m "not gpu"` executes the real kernel code thread by thread: every CUDA thread
// of a block is a fibre, __syncthreads / warp collectives are scheduling points with the hardware's semantics (a divergent
// barrier or a collective that not every named lane reaches is reported as an error instead of hanging).  Nothing here is
// part of libnd_b200.so, and the package never loads the emulated library -- there is no CPU fallback in the product.
#pragma once
#include <math.h>
#include <sched.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <utility>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local

struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace cusim {
struct Ctx { dim3 tid, bid, bdim, gdim; };
extern thread_local Ctx ctx;
enum { OP_BALLOT = 1, OP_SHFL_DOWN = 2, OP_SYNCWARP = 3 };
void barrier_block();                                                        // __syncthreads
unsigned long long warp_collective(int op, unsigned mask, unsigned long long in, int arg);
long long clock_ns();
}  // namespace cusim

#define threadIdx (cusim::ctx.tid)
#define blockIdx (cusim::ctx.bid)
#define blockDim (cusim::ctx.bdim)
#define gridDim (cusim::ctx.gdim)

static inline void __syncthreads() { cusim::barrier_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cusim::warp_collective(cusim::OP_SYNCWARP, mask, 0, 0); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  return (unsigned)cusim::warp_collective(cusim::OP_BALLOT, mask, pred ? 1ull : 0ull, 0);
}
static inline double __shfl_down_sync(unsigned mask, double v, int delta) {
  unsigned long long bits;
  memcpy(&bits, &v, 8);
  bits = cusim::warp_collective(cusim::OP_SHFL_DOWN, mask, bits, delta);
  memcpy(&v, &bits, 8);
  return v;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline long long clock64() { return cusim::clock_ns(); }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// cudaLaunchKernel-style argument unpacking for kernels compiled at run time (see nvrtc.h)
namespace cusim {
template <class... A, size_t... I>
inline void invoke_impl(void (*f)(A...), void** a, std::index_sequence<I...>) {
  f(*reinterpret_cast<typename std::remove_reference<A>::type*>(a[I])...);
}
template <class... A>
inline void invoke(void (*f)(A...), void** a) { invoke_impl(f, a, std::index_sequence_for<A...>{}); }
}  // namespace cusim
