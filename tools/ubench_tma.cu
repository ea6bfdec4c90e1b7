// How fast can ONE SM pull operand tiles out of L2?  A single warp keeps a ring of 16 KB stages full with either
//   mode 0: cp.async.bulk.tensor.2d  [128 rows][64 bf16] boxes, SWIZZLE_128B, out of a row-major [rows][k] matrix
//           (128 row requests of 128 B, 2 KB apart — what the GEMM kernels do), or
//   mode 1: cp.async.bulk 1-D copies of 16 KB contiguous bytes (what they could do if the operand were pre-tiled),
// and nothing consumes the data.  Prints bytes / cycle / SM for a few grid sizes and ring depths.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc> tools/ubench_tma.cu <csrc>/common.cu -o tools/ubench_tma
#include <cstdio>
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

using namespace pgica;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr uint32_t kTile = 16384;

__global__ void __launch_bounds__(128, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int rows, int k, int mode, int stages,
                int iters, int warps, int batch, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bar_all = reinterpret_cast<uint64_t*>(smem + 12 * kTile);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar_all[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w < warps) {
    // each issuing warp owns stages / warps ring slots
    stages /= warps * batch;  // ring slots (of `batch` tiles each) per issuing warp
    smem += (size_t)w * stages * batch * kTile;
    uint64_t* bar = bar_all + w * stages;
    const int kbs = k / 64, rbs = rows / 128;
    int slot = 0;
    uint32_t phase = 0;
    // every SM walks its own sequence of tiles (row block, k chunk), spread over the matrix
    uint32_t t = blockIdx.x * 977u + w * 131u;
    const long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      if (i >= stages) {
        mbar_wait(&bar[slot], phase);
      }
      if (i < iters && elect_one()) {
        mbar_expect_tx(&bar[slot], kTile * batch);
        for (int b = 0; b < batch; ++b) {
          // tile coordinates without divisions (kbs = 16, rbs = 32 are powers of two): the loop must cost what a
          // GEMM's loader pays, not integer division
          const uint32_t tt = t + b * 37u;
          const int kb = (int)(tt & (uint32_t)(kbs - 1)), rb = (int)((tt >> 4) & (uint32_t)(rbs - 1));
          uint8_t* dst = smem + (size_t)(slot * batch + b) * kTile;
          if (mode == 0) {
            tma_load_2d(dst, &tm, &bar[slot], kb * 64, rb * 128);
          } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(dst)),
                         "l"(base + (size_t)(rb * kbs + kb) * kTile), "r"(kTile), "r"(smem_u32(&bar[slot]))
                         : "memory");
          }
        }
      }
      __syncwarp();
      t += 1;
      if (i >= stages - 1 || true) {
        if (++slot == stages) {
          slot = 0;
          if (i >= stages) phase ^= 1;
        }
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && w == 0) out[blockIdx.x] = t1 - t0;
  }
}

int main() {
  const int rows = 4096, k = 1024;  // 8 MB: L2-resident
  uint8_t* d;
  CK(cudaMalloc(&d, (size_t)rows * k * 2));
  CK(cudaMemset(d, 0, (size_t)rows * k * 2));
  long long* d_out;
  CK(cudaMalloc(&d_out, 1024 * sizeof(long long)));
  CUtensorMap tm;
  if (make_tmap_bf16(&tm, d, rows, k, k, 128) != 0) {
    printf("tensor map failed: %s\n", get_error());
    return 1;
  }
  const size_t smem = 1024 + 12 * kTile + 256;
  CK(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4000;
  for (int grid : {1, 148})
    for (int warps : {1, 2, 4})
      for (int batch : {1, 2, 3})
        for (int mode : {0, 1}) {
          const int stages = 12;
          if (stages % (warps * batch) != 0) continue;
          for (int rep = 0; rep < 2; ++rep) {
            tma_rate_kernel<<<grid, 128, smem>>>(tm, d, rows, k, mode, stages, iters, warps, batch, d_out);
            CK(cudaDeviceSynchronize());
          }
          long long h[148];
          CK(cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
          double avg = 0;
          for (int i = 0; i < grid; ++i) avg += (double)h[i];
          avg /= grid;
          printf("{\"bench\": \"tma_rate\", \"grid\": %d, \"issuing_warps\": %d, \"tiles_per_barrier\": %d, \"mode\": \"%s\", "
                 "\"bytes_per_cycle_per_sm\": %.1f, \"cycles_per_iteration_per_warp\": %.0f}\n",
                 grid, warps, batch, mode == 0 ? "tensor_2d_128x64_sw128" : "bulk_1d_16KB",
                 (double)warps * batch * iters * kTile / avg, avg / iters);
        }
  return 0;
}
