// cuda_runtime.h stand-in for tests/cc/wide_emul.cc -- TEST INFRASTRUCTURE ONLY.
//
// Lets g++ compile svinet_b200/csrc/svi_ls_kernels.cuh + svi_ls_wide.cuh as host code: a block is a team of host
// threads (one per CUDA thread), threadIdx / blockIdx are thread-local, __syncthreads is a barrier of the team,
// block-shared arrays (SVI_BLOCK_SHARED) are function-local statics (blocks run one after another).  Only what the
// wide kernels use is functional; the warp-level intrinsics of the register-tiled kernels merely have to parse.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__
#define SVI_BLOCK_SHARED static

struct emu_idx { unsigned x, y, z; };
extern thread_local emu_idx threadIdx, blockIdx;
extern emu_idx blockDim, gridDim;
void __syncthreads();

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { std::abort(); }
template <class T> static inline T __shfl_sync(unsigned, T, int, int = 32) { std::abort(); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
// parse-only (draw / eager-blend kernels of the register tiles, never run in the emulation)
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __ballot_sync(unsigned, int) { std::abort(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
using std::max;
using std::min;
