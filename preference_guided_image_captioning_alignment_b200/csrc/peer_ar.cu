// Progress-gated all-reduce over NVLink peer memory: the data-parallel sum of the LM-head weight gradient, overlapped
// with the kernel that produces it.
//
// The dual backward kernel (sgg_f.cu) finishes dW in vocabulary order and publishes per-segment progress counters
// (release at GPU scope).  This kernel is launched on a second stream next to it: one small CTA per SM (256 threads,
// ~60 registers, no shared memory — it fits beside the persistent CTA the dual kernel keeps on every SM).  For each
// segment it waits for the local counter, meets the other ranks at a flag barrier in peer memory, and then every rank
// reduces ITS 1/world slice of the segment: 16-byte loads of the slice from all `world` buffers (one outstanding load
// per peer and thread: 148 x 256 x world x 16 B in flight covers the NVLink latency-bandwidth product), a sum in fixed
// rank order (deterministic, bit-identical on every rank) and 16-byte stores of the result into all `world` buffers.
// Reduce-scatter and all-gather of a two-shot all-reduce in one pass; per GPU and direction it moves (world-1)/world of
// the buffer, like a ring.  Only the last segment's share is exposed after the producer ends.
//
// With an NVSwitch multicast mapping of the buffer (NVLS) the slice is summed INSIDE the switch: one
// multimem.ld_reduce returns the sum over all ranks, one multimem.st stores it into all of them — a `world`-th of
// the loads, stores and adds on the SMs this kernel shares with the producer.
//
// What DDP's bucket all-reduce does for the reference (pkg/training/trainer.py:201,492,616: accelerator.prepare wraps
// the model, gradients are averaged in backward) — here for the one gradient that dominates the Stage-2 head's traffic.
#include "common.h"
#include "ptx.cuh"

#include <cudaTypedefs.h>

#include <mutex>

namespace pgica {
namespace {

// cuStreamWaitValue32 through the runtime's driver entry point (the library links against cudart only)
PFN_cuStreamWaitValue32 resolve_wait_value() {
  static PFN_cuStreamWaitValue32 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuStreamWaitValue32>(p);
  });
  return fn;
}

constexpr int kMaxWorld = 16;
constexpr int kMaxSeg = 32;
constexpr int kArThreads = 256;

struct PeerArParams {
  float4* bufs[kMaxWorld];
  uint32_t* flags[kMaxWorld];   // rank r's flag area: [(nseg + 1)][world] uint32; slot [s][q] is written by rank q
  const uint32_t* progress;     // local producer's counters (may be null)
  uint32_t target[kMaxSeg];
  long long seg_begin4[kMaxSeg + 1];  // segment boundaries in float4 units
  float4* mc;                   // multicast mapping of the same buffer (null: unicast loads / stores through bufs[])
  uint32_t* local_sync;         // [0] segments released to this rank's CTAs (monotonic), [1] CTAs finished (monotonic)
  int world, rank, nseg;
  uint32_t epoch;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// NVLS: the load is answered by the switch with the element-wise fp32 sum over every GPU of the multicast group, the
// store is replicated to all of them.
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Spin until pred() holds; traps (instead of wedging the GPU) when a peer never shows up.
template <class Pred>
__device__ __forceinline__ void spin_until(Pred&& pred, const char* what) {
  if (pred()) return;
  const long long t0 = clock64();
  while (!pred()) {
    __nanosleep(200);
    if (clock64() - t0 > 3 * PGICA_WATCHDOG_CYCLES) {
      printf("pgica: peer all-reduce watchdog waiting for %s (block %d thread %d)\n", what, (int)blockIdx.x,
             (int)threadIdx.x);
      __trap();
    }
  }
}

// Flag barrier over peer memory, executed by the first `world` threads of ONE CTA: thread q tells rank q "rank `rank`
// has reached point (slot, epoch)" and waits until rank q has told us the same.
__device__ __forceinline__ void peer_barrier(const PeerArParams& p, int slot, const char* what) {
  const int q = threadIdx.x;
  if (q < p.world) {
    st_release_sys(p.flags[q] + (size_t)slot * p.world + p.rank, p.epoch);
    const uint32_t* mine = p.flags[p.rank] + (size_t)slot * p.world + q;
    spin_until([&] { return ld_acquire_sys(mine) >= p.epoch; }, what);
  }
}

template <int W>
__device__ __forceinline__ void reduce_slice(const PeerArParams& p, long long lo, long long hi) {
  for (long long i = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += (long long)gridDim.x * kArThreads) {
    float4 v[W];
#pragma unroll
    for (int r = 0; r < W; ++r) v[r] = ld_sys_v4(p.bufs[r] + i);  // W loads in flight per thread
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < W; ++r) {
      acc.x += v[r].x;
      acc.y += v[r].y;
      acc.z += v[r].z;
      acc.w += v[r].w;
    }
#pragma unroll
    for (int r = 0; r < W; ++r) st_sys_v4(p.bufs[r] + i, acc);
  }
}

__device__ __forceinline__ void reduce_slice_multicast(const PeerArParams& p, long long lo, long long hi) {
  constexpr int U = 8;  // loads in flight per thread (a round trip through the switch each): as many bytes as unicast at 8 ranks
  const long long stride = (long long)gridDim.x * kArThreads;
  for (long long i = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < hi) v[u] = multimem_ld_reduce_v4(p.mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < hi) multimem_st_v4(p.mc + i + u * stride, v[u]);
  }
}

__device__ __forceinline__ void reduce_slice_any(const PeerArParams& p, long long lo, long long hi) {
  for (long long i = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += (long long)gridDim.x * kArThreads) {
    float4 acc = ld_sys_v4(p.bufs[0] + i);
    for (int r = 1; r < p.world; ++r) {
      const float4 v = ld_sys_v4(p.bufs[r] + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    for (int r = 0; r < p.world; ++r) st_sys_v4(p.bufs[r] + i, acc);
  }
}

__global__ void __launch_bounds__(kArThreads) peer_allreduce_progress_kernel(const PeerArParams p) {
  const uint32_t released_before = (p.epoch - 1u) * (uint32_t)p.nseg;
  for (int s = 0; s < p.nseg; ++s) {
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0 && p.progress != nullptr) {
        const uint32_t* c = p.progress + s;
        const uint32_t want = p.target[s];
        spin_until([&] { return ld_acquire_sys(c) >= want; }, "the local producer");
      }
      __syncthreads();
      peer_barrier(p, s, "a peer's segment");  // every rank's segment s is final
      __syncthreads();
      if (threadIdx.x == 0) st_release_gpu_u32(p.local_sync, released_before + (uint32_t)s + 1u);
    } else {
      if (threadIdx.x == 0) {
        const uint32_t want = released_before + (uint32_t)s + 1u;
        spin_until([&] { return ld_acquire_gpu_u32(p.local_sync) >= want; }, "segment release");
      }
      __syncthreads();
    }
    const long long b = p.seg_begin4[s], e = p.seg_begin4[s + 1];
    const long long len = (e - b) / p.world;
    const long long lo = b + (long long)p.rank * len, hi = lo + len;
    if (p.mc != nullptr) {
      reduce_slice_multicast(p, lo, hi);
      continue;
    }
    switch (p.world) {
      case 2: reduce_slice<2>(p, lo, hi); break;
      case 4: reduce_slice<4>(p, lo, hi); break;
      case 8: reduce_slice<8>(p, lo, hi); break;
      default: reduce_slice_any(p, lo, hi); break;
    }
  }
  // every CTA's stores are out -> CTA 0 meets the peers once more: all owners' results have landed in every buffer
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(p.local_sync + 1, 1u);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const uint32_t want = p.epoch * gridDim.x;
      spin_until([&] { return ld_acquire_gpu_u32(p.local_sync + 1) >= want; }, "this rank's CTAs");
      __threadfence_system();
    }
    __syncthreads();
    peer_barrier(p, p.nseg, "a peer's completion");
  }
}

}  // namespace
}  // namespace pgica

extern "C" int pgica_peer_allreduce_progress(const void* const* bufs_host, const void* const* flags_host,
                                             void* multicast_buf, int world, int rank, const uint32_t* progress, const uint32_t* progress_target_host,
                                             const int64_t* seg_begin_host, int nseg, uint32_t epoch,
                                             uint32_t* local_sync, int max_ctas, void* stream) {
  using namespace pgica;
  int rc = pgica_device_check();
  if (rc != PGICA_OK) return rc;
  PGICA_REQUIRE(bufs_host && flags_host && seg_begin_host && local_sync, "peer_allreduce_progress: null pointer");
  PGICA_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
                "peer_allreduce_progress: bad rank %d of %d (at most %d ranks)", rank, world, kMaxWorld);
  PGICA_REQUIRE(nseg >= 1 && nseg <= kMaxSeg, "peer_allreduce_progress: 1..%d segments (got %d)", kMaxSeg, nseg);
  PGICA_REQUIRE(epoch >= 1, "peer_allreduce_progress: epochs start at 1");
  PGICA_REQUIRE(progress == nullptr || progress_target_host != nullptr,
                "peer_allreduce_progress: progress counters without targets");
  PeerArParams p{};
  for (int r = 0; r < world; ++r) {
    PGICA_REQUIRE(bufs_host[r] && flags_host[r] && (reinterpret_cast<uintptr_t>(bufs_host[r]) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(flags_host[r]) & 3u) == 0,
                  "peer_allreduce_progress: buffer / flag pointer of rank %d missing or unaligned", r);
    p.bufs[r] = static_cast<float4*>(const_cast<void*>(bufs_host[r]));
    p.flags[r] = static_cast<uint32_t*>(const_cast<void*>(flags_host[r]));
  }
  for (int s = 0; s <= nseg; ++s) {
    PGICA_REQUIRE(seg_begin_host[s] >= 0 && seg_begin_host[s] % (4 * world) == 0 &&
                      (s == 0 || seg_begin_host[s] >= seg_begin_host[s - 1]),
                  "peer_allreduce_progress: segment boundary %d (%lld) must be a non-decreasing multiple of 4 * world", s,
                  (long long)seg_begin_host[s]);
    p.seg_begin4[s] = seg_begin_host[s] / 4;
  }
  for (int s = 0; s < nseg; ++s) p.target[s] = progress ? progress_target_host[s] : 0u;
  PGICA_REQUIRE((reinterpret_cast<uintptr_t>(multicast_buf) & 15u) == 0, "peer_allreduce_progress: multicast pointer unaligned");
  p.mc = static_cast<float4*>(multicast_buf);
  p.progress = progress;
  p.local_sync = local_sync;
  p.world = world;
  p.rank = rank;
  p.nseg = nseg;
  p.epoch = epoch;
  const int grid = max_ctas > 0 ? max_ctas : device_sm_count();
  if (progress != nullptr) {
    // The kernel must not become resident BEFORE the producer: its CTAs spin on the producer's counters, and several
    // of them packed onto one SM would keep a CTA of the producer's cooperative grid from ever fitting there (the
    // producer's roles wait on each other: deadlock).  The stream therefore first waits, without occupying an SM,
    // until the producer has finished its first segment — by then all of its CTAs are resident, and this kernel's
    // CTAs can only go where there is room beside them.
    PFN_cuStreamWaitValue32 wait_value = resolve_wait_value();
    if (!wait_value) {
      set_error("cuStreamWaitValue32 is not available from this driver");
      return PGICA_ERR_CUDA;
    }
    const CUresult r = wait_value(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(progress), p.target[0],
                                  CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
      set_error("cuStreamWaitValue32 failed with CUresult %d", (int)r);
      return PGICA_ERR_CUDA;
    }
  }
  peer_allreduce_progress_kernel<<<(unsigned)grid, kArThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}
