// svi_ls_build.cuh -- device-side construction of the half-edge CSR that the sweeps read.
//
// Replaces the host loop that turned LinkSampling::assign_training_links' list `_links` (src/linksampling.cc:493-523)
// into adjacency: 2.8 s on one host thread at 1e8 links (5.7 s per rank in a sharded run, where each rank also
// filtered the list in numpy).  Here the link list is uploaded once and every half-edge becomes one 64-bit key
//     (local source node) << 33 | (1 if the source OWNS the link for the s3 sweep) << 32 | neighbour
// (half-edges whose source lies outside the handle's node block get the key nlocal << 33 and sort to the end); one radix
// sort of the keys yields, per node, [neighbours it does not own | neighbours it owns], each part ordered by
// neighbour id -- a canonical order, independent of the order of the input list, so results do not depend on it.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svi {

// Which endpoint sweeps a link in the s3 pass.  The pass is symmetric in its endpoints (src/linksampling.cc:
// 731-746: the full product commutes, and the shortcut always reads "the other endpoint's row at the converged
// endpoint's community"), so any rule works; the parity rule gives every contiguous node block about half of
// its links, whereas "the smaller id owns" hands the low blocks of a sharded run most of the s3 work.
__host__ __device__ inline uint32_t s3_owner(uint32_t p, uint32_t q) {
  const uint32_t lo = p < q ? p : q, hi = p < q ? q : p;
  return ((lo ^ hi) & 1u) ? lo : hi;
}

// err[0] = number of invalid links, err[1] = index of one of them
static __global__ void k_build_keys(const uint32_t *links, uint64_t nlinks, uint32_t n, uint32_t nb, uint32_t ne,
                                    uint64_t *keys, uint32_t *deg_lo, uint32_t *deg_up, unsigned long long *err) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nlinks; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint2 l = reinterpret_cast<const uint2 *>(links)[e];
    const uint32_t p = l.x, q = l.y;
    uint64_t k_up = (uint64_t)(ne - nb) << 33, k_lo = k_up;
    if (p >= n || q >= n || p == q) {
      atomicAdd(err, 1ull);
      err[1] = e;
    } else {
      const uint32_t own = s3_owner(p, q), oth = own == p ? q : p;
      if (own >= nb && own < ne) {
        k_up = ((uint64_t)(own - nb) << 33) | (1ull << 32) | oth;
        atomicAdd(deg_up + (own - nb), 1u);
      }
      if (oth >= nb && oth < ne) {
        k_lo = ((uint64_t)(oth - nb) << 33) | own;
        atomicAdd(deg_lo + (oth - nb), 1u);
      }
    }
    keys[2 * e] = k_up;
    keys[2 * e + 1] = k_lo;
  }
}

static __global__ void k_keys_to_col(const uint64_t *keys, uint64_t he, uint32_t *col) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < he; i += (uint64_t)gridDim.x * blockDim.x)
    col[i] = (uint32_t)keys[i];
}

}  // namespace svi
