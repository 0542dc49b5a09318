// svi_ls_mg.cuh -- device side of the multi-GPU exchange over PEER MEMORY.
//
// Every handle keeps its exchange buffers (the replicated-layout matrices b = exp(Elogpi - max), mphi, gamma, the
// converged flags, the active masks, K-vector slots and a flag table) in one arena; the arenas of all shards are
// mapped into every process (CUDA IPC across processes, plain peer access inside one).  A shard PUSHES the rows it
// produced into every peer's arena (copy engines over NVLink, on a side stream, beside the sweeps) and then raises
// an epoch-stamped flag in the peer's flag table; consumers wait for the flags on their own stream.  The K-vector
// all-reduces (sum, s1, s2 after the node pass; s3 after its sweep) are a push of 3 (1) x K doubles into per-source
// slots plus a fixed-order sum on every shard -- bit-identical on all shards, no NCCL on the data path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svi {

constexpr uint32_t kMaxWorld = 16;
constexpr uint32_t kFlagKinds = 8;
enum MgFlag : uint32_t { FLAG_B = 0, FLAG_M = 1, FLAG_KXN = 2, FLAG_KXS = 3, FLAG_G = 4 };

struct Peers {
  uint32_t world, rank;
  unsigned char *arena[kMaxWorld];   // arena base of every shard as mapped in THIS process (own arena included)
  uint64_t flags_off, kx_off;        // byte offsets of the flag table / K-vector slots inside an arena
  uint64_t gamma_off;                // ... of the gamma matrix
  uint32_t kx_stride;                // doubles per K-vector slot (4 * ld)
  uint32_t bounds[kMaxWorld + 1];    // node blocks of the shards
  // the gamma row of `node` in the arena of the shard that owns it (a peer load over NVLink when it is not ours)
  __device__ __forceinline__ const double *gamma_row(uint32_t node, uint32_t ld) const {
    uint32_t r = 0;
    while (r + 1 < world && node >= bounds[r + 1]) ++r;
    return reinterpret_cast<const double *>(arena[r] + gamma_off) + (size_t)node * ld;
  }
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// flag (source = this shard, kind) := epoch in every peer's table.  Launched on the stream that carried the
// copies it announces: stream order puts it after their completion.
static __global__ void k_mg_signal(const Peers pr, const uint32_t kind, const uint32_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < pr.world && t != pr.rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t *>(pr.arena[t] + pr.flags_off) + pr.rank * kFlagKinds + kind, epoch);
  }
}

// wait until every peer's flag `kind` in OUR table has reached `epoch`.  Bounded: after `timeout_ns` the kernel
// gives up and records the failure in *err (svi_ls_sync reports it) instead of hanging the GPU.
__device__ __forceinline__ void mg_wait_flags(const Peers &pr, uint32_t kind, uint32_t epoch, uint32_t *err,
                                              uint64_t timeout_ns) {
  const uint32_t t = threadIdx.x;
  if (t < pr.world && t != pr.rank) {
    const uint32_t *f = reinterpret_cast<const uint32_t *>(pr.arena[pr.rank] + pr.flags_off) + t * kFlagKinds + kind;
    const uint64_t t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if (global_ns() - t0 > timeout_ns) {
        atomicExch(err, 0x100u + kind * 16u + t);
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}
static __global__ void k_mg_wait(const Peers pr, const uint32_t kind, const uint32_t epoch, uint32_t *err,
                                 const uint64_t timeout_ns) {
  mg_wait_flags(pr, kind, epoch, err, timeout_ns);
}

// Row push by the SMs: every 16-byte piece of [off, off + bytes) of OUR arena is read once and stored into the same
// place of every peer's arena (posted NVLink writes).  A few small blocks on a high-priority stream: they slip onto
// the SMs as sweep blocks retire.  (The alternative is one copy-engine transfer per peer, see svi_ls.cu.)
template <class T>
static __global__ void __launch_bounds__(256) k_mg_push(const Peers pr, const size_t off, const size_t count) {
  const T *src = reinterpret_cast<const T *>(pr.arena[pr.rank] + off);
  T *dst[kMaxWorld];
  uint32_t nd = 0;
  for (uint32_t d = 1; d < pr.world; ++d) dst[nd++] = reinterpret_cast<T *>(pr.arena[(pr.rank + d) % pr.world] + off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const T v = __ldcs(src + i);
    for (uint32_t d = 0; d < nd; ++d) dst[d][i] = v;
  }
  __threadfence_system();
}

// Link-community membership words of OUR rows: OR of every shard's replica (a shard sets the bits of both endpoints of
// the links it owns in its own replica, svi_ls_ring.cuh `publish`).  Peer loads, ~n * words * 4 / world bytes per peer.
static __global__ void __launch_bounds__(256) k_mg_or_rows(const Peers pr, const size_t off, const size_t first,
                                                           const size_t count) {
  uint32_t *own = reinterpret_cast<uint32_t *>(pr.arena[pr.rank] + off) + first;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = own[i];
    for (uint32_t d = 1; d < pr.world; ++d)
      v |= __ldcg(reinterpret_cast<const uint32_t *>(pr.arena[(pr.rank + d) % pr.world] + off) + first + i);
    own[i] = v;
  }
}

// all-reduce, first half: our `count` doubles (count <= kx_stride) go into slot [parity][which][rank] of every
// shard's arena (our own included), then the flag `kind` is raised.  One block.
//   extra (may be null): one more value rides along at index `count` -- the shard's "a node newly converged" flag
static __global__ void k_mg_kx_push(const Peers pr, const double *src, const uint32_t count, const uint32_t parity,
                                    const uint32_t which, const uint32_t kind, const uint32_t epoch,
                                    const uint32_t *extra) {
  const size_t slot = ((size_t)(parity * 2u + which) * kMaxWorld + pr.rank) * pr.kx_stride;
  for (uint32_t t = 0; t < pr.world; ++t) {
    double *dst = reinterpret_cast<double *>(pr.arena[t] + pr.kx_off) + slot;
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
    if (extra && threadIdx.x == 0) dst[count] = (double)*extra;
  }
  __threadfence_system();
  __syncthreads();
  const uint32_t t = threadIdx.x;
  if (t < pr.world && t != pr.rank)
    st_release_sys(reinterpret_cast<uint32_t *>(pr.arena[t] + pr.flags_off) + pr.rank * kFlagKinds + kind, epoch);
}

// all-reduce, second half: wait for every shard's slot, then sum them in rank order (the same order on every
// shard: the result is bit-identical everywhere).  One block.
//   extra_out (may be null): receives 1 if any shard's extra value was non-zero
static __global__ void k_mg_kx_sum(const Peers pr, double *dst, const uint32_t count, const uint32_t parity,
                                   const uint32_t which, const uint32_t kind, const uint32_t epoch, uint32_t *err,
                                   const uint64_t timeout_ns, uint32_t *extra_out) {
  mg_wait_flags(pr, kind, epoch, err, timeout_ns);
  __syncthreads();
  const double *base = reinterpret_cast<const double *>(pr.arena[pr.rank] + pr.kx_off) +
                       (size_t)(parity * 2u + which) * kMaxWorld * pr.kx_stride;
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) {
    double s = 0.0;
    for (uint32_t r = 0; r < pr.world; ++r) s += __ldcg(base + (size_t)r * pr.kx_stride + i);
    dst[i] = s;
  }
  if (extra_out && threadIdx.x == 0) {
    double any = 0.0;
    for (uint32_t r = 0; r < pr.world; ++r) any += __ldcg(base + (size_t)r * pr.kx_stride + count);
    *extra_out = any > 0.0 ? 1u : 0u;
  }
}

}  // namespace svi
