// svi_wide_reduce.cuh -- what the block-per-row kernels (svi_ls_wide.cuh, svi_fa2_wide.cuh) share: the block size and
// block-wide reductions through a shared-memory tree in a fixed order (the same bits on every run).
//
// Nothing here but threadIdx, __syncthreads and caller-provided block-shared arrays, so the kernels built on it also
// run as host code under tests/cc/*_emul.cc (a block = kWideT host threads, __syncthreads = a barrier).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#ifndef SVI_WIDE_T
#define SVI_WIDE_T 256
#endif
#ifndef SVI_BLOCK_SHARED
#define SVI_BLOCK_SHARED __shared__
#endif

namespace svi {

constexpr uint32_t kWideT = SVI_WIDE_T;   // threads per block (a power of two)

// per-block slot width of the column partials: k_reduce_kpart's `cap` (= 2 * Ops::lanes * Ops::vec on the host)
static inline uint32_t wide_cap(uint32_t ld) { return ((ld + 2u * kWideT - 1u) / (2u * kWideT)) * (2u * kWideT); }

// ---- block-wide reductions: every thread of the block calls them, every thread gets the result --------------------
// (leading barrier: the previous result has been read by everybody before `red` is written again)
__device__ __forceinline__ double wide_sum(double v, double *red) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  red[t] = v;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  return red[0];
}
__device__ __forceinline__ double wide_max(double v, double *red) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  red[t] = v;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) red[t] = fmax(red[t], red[t + s]);
    __syncthreads();
  }
  return red[0];
}
// largest value; among equal values the smallest index (D1Array::max, src/matrix.hh:521-532: the first maximum)
__device__ __forceinline__ void wide_argmax(double &best, uint32_t &bestk, double *red, uint32_t *redk) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  red[t] = best;
  redk[t] = bestk;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) {
      const double ob = red[t + s];
      const uint32_t ok = redk[t + s];
      if (ob > red[t] || (ob == red[t] && ok < redk[t])) { red[t] = ob; redk[t] = ok; }
    }
    __syncthreads();
  }
  best = red[0];
  bestk = redk[0];
}
// sum of `cnt` and maximum of `mx` over the block
__device__ __forceinline__ void wide_count_max(uint32_t &cnt, uint32_t &mx, uint32_t *redc, uint32_t *redm) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  redc[t] = cnt;
  redm[t] = mx;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) {
      redc[t] += redc[t + s];
      redm[t] = max(redm[t], redm[t + s]);
    }
    __syncthreads();
  }
  cnt = redc[0];
  mx = redm[0];
}

// two sums / two maxima at once (half the barriers of two calls)
__device__ __forceinline__ void wide_sum2(double &a, double &b, double *red, double *red2) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  red[t] = a;
  red2[t] = b;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) {
      red[t] += red[t + s];
      red2[t] += red2[t + s];
    }
    __syncthreads();
  }
  a = red[0];
  b = red2[0];
}
__device__ __forceinline__ void wide_max2(double &a, double &b, double *red, double *red2) {
  const uint32_t t = threadIdx.x;
  __syncthreads();
  red[t] = a;
  red2[t] = b;
  __syncthreads();
  for (uint32_t s = kWideT / 2; s > 0; s >>= 1) {
    if (t < s) {
      red[t] = fmax(red[t], red[t + s]);
      red2[t] = fmax(red2[t], red2[t + s]);
    }
    __syncthreads();
  }
  a = red[0];
  b = red2[0];
}

}  // namespace svi
