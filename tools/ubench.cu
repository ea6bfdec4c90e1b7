// Micro-benchmarks that pin the B200 numbers the softmax-gradient GEMM is designed around:
//   (1) tcgen05.mma issue/execute rate per shape (N = 128 / 256), B-operand major, number of accumulators
//   (2) cp.async.bulk shared::cta -> shared::cluster (DSMEM) copy bandwidth inside a 2/4-CTA cluster
//   (3) tcgen05.ld epilogue read rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc> tools/ubench.cu -o tools/ubench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"

using namespace pgica;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e = (x);                                                                        \
    if (e != cudaSuccess) {                                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);            \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

// ------------------------------------------------------------------------------------------------ (1)
// One thread issues `iters` MMAs (M=128, N, K=16) on fixed smem operands, then commits; cycles from the first issue
// to the commit arrival.  b_mn: B operand read MN-major (as MMA2 of the sgg kernel does).  n_acc accumulators are
// used round-robin.  a_tmem: A operand from TMEM (columns 256..).
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int b_mn, int n_acc, int iters, int a_tmem, int noise,
                                                          long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    *stop = 0;
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, b_mn);
    const uint32_t a_addr = smem_u32(smem);              // A: [128][64] K-major, 16 KB (4 k-steps)
    const uint32_t b_addr = smem_u32(smem + 32 * 1024);  // B: up to 256 x 64, 32 KB
    const uint64_t da0 = make_smem_desc(a_addr, 16, 1024);
    const uint64_t db0 = b_mn ? make_smem_desc(b_addr, 16 * 1024, 1024) : make_smem_desc(b_addr, 16, 1024);
    const uint32_t bstep = b_mn ? (16 * 128 >> 4) : 2;
    const uint32_t d0 = tmem_base, d1 = tmem_base + (n_acc > 1 ? N : 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = (u & 1) ? d1 : d0;
          if (a_tmem) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                "r"(tmem_base + 448u), "l"(db0 + (u & 3) * bstep), "r"(idesc), "r"(1u)
                : "memory");
          } else {
            umma_bf16_ss(d, da0 + (u & 3) * 2, db0 + (u & 3) * bstep, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 2 + 0] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      *stop = 1;
    }
  } else if (warp >= 2 && noise) {
    // background shared-memory traffic (stand-in for TMA fills / epilogue stores): 16-byte stores + loads
    uint32_t addr = smem_u32(smem + 96 * 1024) + (threadIdx.x - 64) * 16;
    uint32_t acc = 0;
    while (!*stop) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (noise == 1) {
          st_smem_v4(addr + u * 1024, acc, u, 1, 2);
        } else {
          uint32_t x, y, z, w;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr + u * 1024));
          acc += x + y + z + w;
        }
      }
    }
    if (acc == 0x12345) out[4000] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ (1b)
// cta_group::2: a CTA pair issues M=256 (128 rows per CTA) x N x 16 MMAs; the leader CTA issues, B is split across
// the two CTAs' shared memories by the hardware.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma2_rate_kernel(int N, int b_mn, int iters, int a_tmem, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N, 0, b_mn);
      const uint32_t a_addr = smem_u32(smem);
      const uint32_t b_addr = smem_u32(smem + 32 * 1024);
      const uint64_t da0 = make_smem_desc(a_addr, 16, 1024);
      const uint64_t db0 = b_mn ? make_smem_desc(b_addr, 16 * 1024, 1024) : make_smem_desc(b_addr, 16, 1024);
      const uint32_t bstep = b_mn ? (16 * 128 >> 4) : 2;
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (a_tmem) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(
                      tmem_base),
                  "r"(tmem_base + 448u), "l"(db0 + (u & 3) * bstep), "r"(idesc), "r"(1u), "r"(0u)
                  : "memory");
            } else {
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(
                      tmem_base),
                  "l"(da0 + (u & 3) * 2), "l"(db0 + (u & 3) * bstep), "r"(idesc), "r"(1u), "r"(0u)
                  : "memory");
            }
          }
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
      __syncwarp();
      mbar_wait(bar, 0);
      const long long t2 = clock64();
      if (lane == 0) {
        out[(blockIdx.x >> 1) * 2 + 0] = t1 - t0;
        out[(blockIdx.x >> 1) * 2 + 1] = t2 - t0;
      }
    } else {
      mbar_wait(bar, 0);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ (2)
// Every CTA of the cluster sends `bytes` to each of its C-1 peers, `iters` times (waiting for its own inbound copies
// of the round before the next round).
template <int C>
__global__ void __launch_bounds__(128, 1) dsmem_kernel(int bytes, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* src = smem;                 // 64 KB
  uint8_t* dst = smem + 64 * 1024;     // C slots x 32 KB max
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  const uint32_t q = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(src)[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint64_t* b = &bar[it & 1];  // a peer is at most one round ahead: alternate barriers, no phase mixing
      mbar_expect_tx(b, (uint32_t)bytes * (C - 1));  // my inbound traffic this round
      for (int c = 1; c < C; ++c) {
        const uint32_t peer = (q + c) % C;
        const uint32_t rbar = mapa_u32(smem_u32(b), peer);
        dsmem_bulk_copy(mapa_u32(smem_u32(dst + (size_t)q * 32 * 1024), peer), smem_u32(src), (uint32_t)bytes, rbar);
      }
      mbar_wait_cluster(b, (it >> 1) & 1);
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------ (3)
__global__ void __launch_bounds__(128, 1) tmem_ld_kernel(int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int ch = 0; ch < 16; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + lane_addr + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int C>
void run_dsmem(int bytes, int iters, int clusters, long long* d_out) {
  auto kern = dsmem_kernel<C>;
  const size_t smem = 1024 + 201 * 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * C);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, bytes, iters, d_out));
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(clusters * C);
  CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (auto v : h) mean += (double)v;
  mean /= h.size();
  printf("{\"bench\": \"dsmem_bulk\", \"C\": %d, \"clusters\": %d, \"bytes\": %d, \"iters\": %d, \"cycles_per_round\": %.1f, "
         "\"out_bytes_per_cycle_per_cta\": %.2f}\n",
         C, clusters, bytes, iters, mean / iters, (double)bytes * (C - 1) / (mean / iters));
}

int main() {
  long long* d_out;
  float* d_sink;
  CK(cudaMalloc(&d_out, 4096 * 8));
  CK(cudaMalloc(&d_sink, 16));
  const size_t smem = 1024 + 161 * 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4096;
  for (int grid : {1, 148}) {
    for (int noise : {0}) {
      for (int a_tmem : {0, 1}) {
        for (int N : {128, 256}) {
          for (int b_mn : {0, 1}) {
            for (int n_acc : {1, 2}) {
              if (n_acc * N > 448) continue;
              if (grid == 1 && noise) continue;
              mma_rate_kernel<<<grid, 128, smem>>>(N, b_mn, n_acc, iters, a_tmem, noise, d_out);
              CK(cudaDeviceSynchronize());
              std::vector<long long> h(grid * 2);
              CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
              double issue = 0, total = 0;
              for (int i = 0; i < grid; ++i) {
                issue += (double)h[2 * i];
                total += (double)h[2 * i + 1];
              }
              issue /= grid;
              total /= grid;
              printf("{\"bench\": \"mma_rate\", \"grid\": %d, \"noise\": %d, \"a_tmem\": %d, \"N\": %d, \"b_mn_major\": %d, "
                     "\"n_acc\": %d, \"issue_cycles_per_mma\": %.1f, \"cycles_per_mma\": %.1f, \"ideal\": %d, "
                     "\"flop_per_cycle\": %.0f}\n",
                     grid, noise, a_tmem, N, b_mn, n_acc, issue / iters, total / iters, N / 2,
                     2.0 * 128 * N * 16 * iters / total);
            }
          }
        }
      }
    }
  }
  CK(cudaFuncSetAttribute(mma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {2, 148}) {
    for (int a_tmem : {0, 1}) {
      for (int N : {64, 128, 256}) {
        for (int b_mn : {0, 1}) {
          mma2_rate_kernel<<<grid, 128, smem>>>(N, b_mn, iters, a_tmem, d_out);
          CK(cudaDeviceSynchronize());
          const int pairs = grid / 2;
          std::vector<long long> h(pairs * 2);
          CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
          double issue = 0, total = 0;
          for (int i = 0; i < pairs; ++i) {
            issue += (double)h[2 * i];
            total += (double)h[2 * i + 1];
          }
          issue /= pairs;
          total /= pairs;
          printf("{\"bench\": \"mma2_rate\", \"grid\": %d, \"a_tmem\": %d, \"N\": %d, \"b_mn_major\": %d, "
                 "\"issue_cycles_per_mma\": %.1f, \"cycles_per_mma\": %.1f, \"ideal\": %d, \"flop_per_cycle_per_sm\": %.0f}\n",
                 grid, a_tmem, N, b_mn, issue / iters, total / iters, N / 2, 2.0 * 128 * N * 16 * iters / total);
        }
      }
    }
  }
  for (int clusters : {1, 32}) {
    for (int bytes : {8192, 32768}) {
      run_dsmem<2>(bytes, 64, clusters, d_out);
      run_dsmem<4>(bytes, 64, clusters, d_out);
    }
  }
  for (int grid : {1, 148}) {
    tmem_ld_kernel<<<grid, 128>>>(64, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (auto v : h) mean += (double)v;
    mean /= grid;
    // per iteration: 4 warps x 16 loads x (32 lanes x 32 cols x 4 B) = 256 KB
    printf("{\"bench\": \"tmem_ld_32x32b_x32\", \"grid\": %d, \"cycles_per_128x512_tile\": %.1f, \"bytes_per_cycle\": %.1f}\n",
           grid, mean / 64, 262144.0 / (mean / 64));
  }
  return 0;
}
