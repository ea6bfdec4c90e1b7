// Self-test kernel: a single CTA performs one 128 x n x k tcgen05 product with the operand layouts the
// production kernels rely on (K-major TMA tiles, MN-major TMA tiles, a hand-swizzled A tile) and dumps
// the accumulator.  tests/ compares it against a host product, which pins the descriptor conventions.
#include "common.h"
#include "ptx.cuh"

namespace pgica {
namespace {

struct ProbeParams {
  int n, k, b_mn_major, a_manual;
  uint32_t b_lbo, b_sbo;
  const __nv_bfloat16* a;
  float* d;
};

constexpr size_t kProbeSmem = 1024 + 64 * 1024 + 128 * 1024 + 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + 64 * 1024;
  uint64_t* load_bar = reinterpret_cast<uint64_t*>(smem + 192 * 1024);
  uint64_t* mma_bar = load_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = p.k / 64;
  const int nchunks = p.n / 64;

  if (threadIdx.x == 0) {
    mbar_init(load_bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (p.a_manual) {
    // each thread owns one row: 16-byte chunks go through the 128-B swizzle
    const int row = threadIdx.x;
    for (int kc = 0; kc < kchunks; ++kc) {
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.a + (size_t)row * p.k + kc * 64 + c * 8);
        *reinterpret_cast<uint4*>(smem_a + kc * 16384 + sw128_offset(row, c)) = v;
      }
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (!p.a_manual) bytes += kchunks * 16384u;
    bytes += p.b_mn_major ? nchunks * (uint32_t)p.k * 128u : kchunks * (uint32_t)p.n * 128u;
    mbar_expect_tx(load_bar, bytes);
    if (!p.a_manual)
      for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(smem_a + kc * 16384, &tm_a, load_bar, kc * 64, 0);
    if (p.b_mn_major) {
      for (int nc = 0; nc < nchunks; ++nc) tma_load_2d(smem_b + nc * p.k * 128, &tm_b, load_bar, nc * 64, 0);
    } else {
      for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(smem_b + kc * p.n * 128, &tm_b, load_bar, kc * 64, 0);
    }
    mbar_wait(load_bar, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, p.n, 0, p.b_mn_major);
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
    for (int kk = 0; kk < p.k / 16; ++kk) {
      const uint64_t da = make_smem_desc(a_base + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
      uint64_t db;
      if (p.b_mn_major)
        db = make_smem_desc(b_base + kk * 16 * 128, p.b_lbo, p.b_sbo);
      else
        db = make_smem_desc(b_base + (kk >> 2) * p.n * 128 + (kk & 3) * 32, 16, 1024);
      umma_bf16_ss(tmem_base, da, db, idesc, kk != 0);
    }
    umma_commit(mma_bar);
  }
  __syncthreads();
  mbar_wait(mma_bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int ch = 0; ch < p.n / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[(size_t)row * p.n + ch * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace
}  // namespace pgica

extern "C" int pgica_probe_umma(const void* a, const void* b, int64_t n, int64_t k, int b_mn_major, int a_manual,
                                uint32_t b_lbo_bytes, uint32_t b_sbo_bytes, float* d, void* stream) {
  using namespace pgica;
  int rc = pgica_device_check();
  if (rc != PGICA_OK) return rc;
  PGICA_REQUIRE(n >= 64 && n <= 256 && n % 64 == 0, "probe: n must be 64, 128, 192 or 256");
  PGICA_REQUIRE(k >= 64 && k <= 256 && k % 64 == 0, "probe: k must be 64..256, multiple of 64");
  CUtensorMap tm_a, tm_b;
  rc = make_tmap_bf16(&tm_a, a, 128, k, k, 128);
  if (rc != PGICA_OK) return rc;
  if (b_mn_major)
    rc = make_tmap_bf16(&tm_b, b, k, n, n, (uint32_t)k);
  else
    rc = make_tmap_bf16(&tm_b, b, n, k, k, (uint32_t)n);
  if (rc != PGICA_OK) return rc;
  ProbeParams p{(int)n, (int)k, b_mn_major, a_manual, b_lbo_bytes, b_sbo_bytes,
                static_cast<const __nv_bfloat16*>(a), d};
  PGICA_CUDA_OK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProbeSmem));
  probe_kernel<<<1, 128, kProbeSmem, static_cast<cudaStream_t>(stream)>>>(tm_a, tm_b, p);
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}
