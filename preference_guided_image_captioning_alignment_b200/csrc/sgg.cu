// K2/K4: the backward of both loss heads as one "softmax-gradient GEMM":
//
//   Out[i, :] = sum_j G(i, j) * Y[j, :]         G(i, j) = rC[i] * (exp(s*z_ij - rL[i]) - [j == rT[i]])
//                                                       + cC[j] * (exp(s*z_ij - cL[j]) - [i == cT[j]])
//   z_ij = <X[i, :], Y[j, :]>   recomputed tile by tile on the tensor cores; G never leaves the SM.
//
//   DPO dH : X = H (tokens),  Y = W (vocab),  row term only   (rC = -coef, rL = lse, rT = label)
//   DPO dW : X = W (vocab),   Y = H (tokens), column term only
//   NT-Xent dA / dB : X = a / b, Y = b / a, both terms (row and column log-sum-exp)
//
// One CTA owns a 128-row x 256-column block of Out, resident in TMEM (256 columns) for the whole loop over
// the 128-row tiles of Y.  Per tile: MMA1 (128x128xK) -> Z in TMEM (double-buffered) -> epilogue warps turn Z
// into the bf16 G tile in shared memory (K-major, 128-B swizzle) -> MMA2 (128x256x128) accumulates into Out
// with the Y tile read MN-major.  Operands stream through one TMA ring of 32-KB slots.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace pgica {
namespace {

constexpr int kBM = 128;        // rows of X / Out per CTA
constexpr int kBT = 128;        // rows of Y per tile (N of MMA1, K of MMA2)
constexpr int kBD = 256;        // Out columns per CTA (N of MMA2)
constexpr int kBK = 64;         // k-chunk (one 128-B swizzle row)
constexpr int kSlots = 6;       // ring depth (even: a two-slot item never wraps)
constexpr uint32_t kChunkBytes = 128 * kBK * 2;   // 16 KB: a [128][64] bf16 box
constexpr uint32_t kSlotBytes = 2 * kChunkBytes;  // 32 KB
constexpr uint32_t kPBytes = kBM * kBT * 2;       // 32 KB
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

constexpr uint32_t kTmemOut = 0, kTmemZ = 256;  // column offsets: Out [0,256), Z0 [256,384), Z1 [384,512)

struct SggParams {
  int mx, my, k, num_tiles, passes, ldo, out_bf16;
  float c;  // scale * log2(e)
  const float* r_lse;
  const float* r_coef;
  const int* r_tgt;
  const float* c_lse;
  const float* c_coef;
  const int* c_tgt;
  void* out;
};

constexpr size_t kSggSmem = 1024 + kSlots * kSlotBytes + kPBytes + 3 * kBT * 4 + 256;

template <bool kRow, bool kCol>
__global__ void __launch_bounds__(kThreads, 1)
sgg_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, const SggParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* ring = smem;
  uint8_t* p_tile = smem + kSlots * kSlotBytes;
  float* s_cl = reinterpret_cast<float*>(p_tile + kPBytes);
  float* s_cc = s_cl + kBT;
  int* s_ct = reinterpret_cast<int*>(s_cc + kBT);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ct + kBT);
  uint64_t* empty_bar = full_bar + kSlots;
  uint64_t* zfull_bar = empty_bar + kSlots;   // [2]
  uint64_t* zempty_bar = zfull_bar + 2;       // [2]
  uint64_t* pfull_bar = zempty_bar + 2;
  uint64_t* pfree_bar = pfull_bar + 1;
  uint64_t* out_bar = pfree_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x / p.passes;
  const int q = blockIdx.x - m_blk * p.passes;
  const int num_kb = (p.k + kBK - 1) / kBK;
  const bool pad = (num_kb & 1) != 0;
  const int J = p.num_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&zfull_bar[i], 1);
      mbar_init(&zempty_bar[i], 128);
    }
    mbar_init(pfull_bar, 128);
    mbar_init(pfree_bar, 1);
    mbar_init(out_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp walks the schedule so that control flow stays warp-uniform; one elected lane issues the copies)
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == kSlots) {
        slot = 0;
        phase ^= 1;
      }
    };
    auto load_mma1 = [&](int j) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[slot], kSlotBytes);
          uint8_t* dst = ring + slot * kSlotBytes;
          tma_load_2d(dst, &tm_x, &full_bar[slot], kb * kBK, m_blk * kBM);
          tma_load_2d(dst + kChunkBytes, &tm_y, &full_bar[slot], kb * kBK, j * kBT);
        }
        __syncwarp();
        advance();
      }
      if (pad) {  // keep every item an even number of slots
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (elect_one()) mbar_expect_tx(&full_bar[slot], 0);
        __syncwarp();
        advance();
      }
    };
    auto load_mma2 = [&](int j) {  // Y[j tile, q*256 .. +256) as four [128][64] boxes in two slots
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[slot], kSlotBytes);
          uint8_t* dst = ring + slot * kSlotBytes;
          tma_load_2d(dst, &tm_y, &full_bar[slot], q * kBD + (2 * h) * kBK, j * kBT);
          tma_load_2d(dst + kChunkBytes, &tm_y, &full_bar[slot], q * kBD + (2 * h + 1) * kBK, j * kBT);
        }
        __syncwarp();
        advance();
      }
    };
    for (int j = 0; j < J; ++j) {
      load_mma1(j);
      if (j > 0) load_mma2(j - 1);
    }
    load_mma2(J - 1);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc1 = make_idesc_bf16(kBM, kBT, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(kBM, kBD, 0, 1);
    const uint64_t desc_k = make_smem_desc(0, 16, 1024);            // K-major operand, start address 0
    const uint64_t desc_mn = make_smem_desc(0, kChunkBytes, 1024);  // MN-major operand (Y tile of MMA2)
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == kSlots) {
        slot = 0;
        phase ^= 1;
      }
    };
    uint32_t pphase = 0;
    const uint32_t p_addr = smem_u32(p_tile);
    auto mma2 = [&](int j) {
      mbar_wait(pfull_bar, pphase);
      pphase ^= 1;
      mbar_wait(&full_bar[slot], phase);
      const int slot2 = slot + 1;  // never wraps: items are even-sized, ring is even
      mbar_wait(&full_bar[slot2], phase);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t y_addr = smem_u32(ring + slot * kSlotBytes);
        const uint64_t dg = desc_k | ((p_addr >> 4) & 0x3FFF);
        const uint64_t dy = desc_mn | ((y_addr >> 4) & 0x3FFF);
#pragma unroll
        for (int ks = 0; ks < kBT / 16; ++ks)
          umma_bf16_ss(tmem_base + kTmemOut, dg + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2,
                       dy + ks * (16 * 128 >> 4), idesc2, (j | ks) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[slot]);
        umma_commit(&empty_bar[slot2]);
        umma_commit(pfree_bar);
      }
      __syncwarp();
      advance();
      advance();
    };
    int zb = 0;
    uint32_t zphase = 0;
    for (int j = 0; j < J; ++j) {
      mbar_wait(&zempty_bar[zb], zphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + kTmemZ + zb * kBT;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t x_addr = smem_u32(ring + slot * kSlotBytes);
          const uint64_t da = desc_k | ((x_addr >> 4) & 0x3FFF);
          const uint64_t db = desc_k | (((x_addr + kChunkBytes) >> 4) & 0x3FFF);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[slot]);
          if (kb == num_kb - 1) umma_commit(&zfull_bar[zb]);
        }
        __syncwarp();
        advance();
      }
      if (pad) {
        mbar_wait(&full_bar[slot], phase);
        if (elect_one()) mbar_arrive(&empty_bar[slot]);
        __syncwarp();
        advance();
      }
      if (++zb == 2) {
        zb = 0;
        zphase ^= 1;
      }
      if (j > 0) mma2(j - 1);
    }
    mma2(J - 1);
    if (elect_one()) umma_commit(out_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: Z -> G (bf16, smem)
    const int quarter = warp & 3;
    const int et = threadIdx.x - 128;  // 0..127
    const int row_in_blk = quarter * 32 + lane;
    const int row = m_blk * kBM + row_in_blk;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float rl = 0.f, rc = 0.f;
    int rt = -1;
    if (kRow && row < p.mx) {
      rl = p.r_lse[row] * kLog2e;
      rc = p.r_coef[row];
      rt = p.r_tgt ? p.r_tgt[row] : -1;
    }
    auto load_col = [&](int j, float& l, float& cf, int& tg) {
      const int col = j * kBT + et;
      l = 0.f;
      cf = 0.f;
      tg = -1;
      if (kCol && col < p.my) {
        l = p.c_lse[col] * kLog2e;
        cf = p.c_coef[col];
        tg = p.c_tgt ? p.c_tgt[col] : -1;
      }
    };
    float nl, nc;
    int nt;
    load_col(0, nl, nc, nt);
    int zb = 0;
    uint32_t zphase = 0, fphase = 0;
    const uint32_t p_addr = smem_u32(p_tile);
    for (int j = 0; j < J; ++j) {
      if (kCol) {
        s_cl[et] = nl;
        s_cc[et] = nc;
        s_ct[et] = nt;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (j + 1 < J) load_col(j + 1, nl, nc, nt);
      }
      mbar_wait(&zfull_bar[zb], zphase);
      tc_fence_after_sync();
      mbar_wait(pfree_bar, fphase ^ 1);  // MMA2(j-1) has finished reading the G tile
      fphase ^= 1;
      const int col0 = j * kBT;
      const int rrel = rt - col0;
#pragma unroll 1
      for (int ch = 0; ch < kBT / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + lane_addr + kTmemZ + zb * kBT + ch * 32, r);
        tmem_ld_wait();
        float g[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float t = __uint_as_float(r[jj]) * p.c;
          float v = 0.f;
          if (kRow) v = rc * fast_exp2(t - rl);
          if (kCol) {
            const int cj = ch * 32 + jj;
            const float ccj = s_cc[cj];
            v = fmaf(ccj, fast_exp2(t - s_cl[cj]), v);
            if (s_ct[cj] == row) v -= ccj;
          }
          g[jj] = v;
        }
        if (kRow && rrel >= 0 && (rrel >> 5) == ch) {
          const int jj0 = rrel & 31;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (jj == jj0) g[jj] -= rc;
        }
        // 32 columns = 4 16-byte chunks of this row inside k-chunk (ch >> 1)
        const uint32_t base = p_addr + (ch >> 1) * kChunkBytes;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t off = sw128_offset(row_in_blk, (ch & 1) * 4 + c4);
          st_smem_v4(base + off, pack_bf16x2(g[c4 * 8 + 0], g[c4 * 8 + 1]), pack_bf16x2(g[c4 * 8 + 2], g[c4 * 8 + 3]),
                     pack_bf16x2(g[c4 * 8 + 4], g[c4 * 8 + 5]), pack_bf16x2(g[c4 * 8 + 6], g[c4 * 8 + 7]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(pfull_bar);
      mbar_arrive(&zempty_bar[zb]);
      if (++zb == 2) {
        zb = 0;
        zphase ^= 1;
      }
      if (kCol) asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ------------------------------------------------------------------ final: Out (TMEM) -> global
    mbar_wait(out_bar, 0);
    tc_fence_after_sync();
    const int ocol0 = q * kBD;
#pragma unroll 1
    for (int ch = 0; ch < kBD / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + lane_addr + kTmemOut + ch * 32, r);
      tmem_ld_wait();
      const int col = ocol0 + ch * 32;
      if (row < p.mx && col < p.k) {
        if (col + 32 <= p.k) {
          if (p.out_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(r[c4 * 8 + 0]), __uint_as_float(r[c4 * 8 + 1]));
              v.y = pack_bf16x2(__uint_as_float(r[c4 * 8 + 2]), __uint_as_float(r[c4 * 8 + 3]));
              v.z = pack_bf16x2(__uint_as_float(r[c4 * 8 + 4]), __uint_as_float(r[c4 * 8 + 5]));
              v.w = pack_bf16x2(__uint_as_float(r[c4 * 8 + 6]), __uint_as_float(r[c4 * 8 + 7]));
              dst[c4] = v;
            }
          } else {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<float*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_uint4(r[c4 * 4], r[c4 * 4 + 1], r[c4 * 4 + 2], r[c4 * 4 + 3]);
          }
        } else {
          for (int jj = 0; jj < 32 && col + jj < p.k; ++jj) {
            if (p.out_bf16)
              static_cast<__nv_bfloat16*>(p.out)[(size_t)row * p.ldo + col + jj] =
                  __float2bfloat16(__uint_as_float(r[jj]));
            else
              static_cast<float*>(p.out)[(size_t)row * p.ldo + col + jj] = __uint_as_float(r[jj]);
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace

int sggx_dispatch(int cluster, const void* x, const void* y, int64_t mx, int64_t my, int64_t k, float scale,
                  const float* r_lse, const float* r_coef, const int32_t* r_tgt, const float* c_lse,
                  const float* c_coef, const int32_t* c_tgt, void* out, int out_is_bf16, void* workspace,
                  size_t workspace_bytes, cudaStream_t st);
size_t sggx_workspace_bytes();

// Cluster size for the shared-recompute kernel (sgg_x.cu): the largest of {4, 2} whose 256-column slices tile k exactly,
// unless option sgg_cluster caps it (1 = the single-CTA kernel in this file).
static int choose_cluster(int64_t k) {
  int want = (int)get_option(kOptSggCluster);
  if (want <= 0) want = 4;
  if (want >= 4 && k % 1024 == 0) return 4;
  if (want >= 2 && k % 512 == 0) return 2;
  return 1;
}
}  // namespace pgica

extern "C" int pgica_softmax_grad_gemm_workspace_bytes(int64_t mx, int64_t my, int64_t k, size_t* bytes_host) {
  PGICA_REQUIRE(bytes_host, "workspace query: null result pointer");
  (void)mx;
  (void)my;
  (void)k;
  *bytes_host = pgica::sggx_workspace_bytes();
  return PGICA_OK;
}

extern "C" int pgica_softmax_grad_gemm(const void* x, const void* y, int64_t mx, int64_t my, int64_t k, float scale,
                                       const float* r_lse, const float* r_coef, const int32_t* r_tgt,
                                       const float* c_lse, const float* c_coef, const int32_t* c_tgt, void* out,
                                       int out_is_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pgica;
  int rc = pgica_device_check();
  if (rc != PGICA_OK) return rc;
  PGICA_REQUIRE(x && y && out, "softmax_grad_gemm: null operand");
  PGICA_REQUIRE(mx > 0 && my > 0 && k > 0 && k % 8 == 0, "softmax_grad_gemm: bad shape (mx %lld my %lld k %lld)",
                (long long)mx, (long long)my, (long long)k);
  PGICA_REQUIRE(mx < (1ll << 30) && my < (1ll << 30) && k <= (1 << 20), "softmax_grad_gemm: dimension too large");
  const bool row = r_lse != nullptr, col = c_lse != nullptr;
  PGICA_REQUIRE(row || col, "softmax_grad_gemm: need row statistics, column statistics or both");
  PGICA_REQUIRE(!row || r_coef, "softmax_grad_gemm: r_coef missing");
  PGICA_REQUIRE(!col || c_coef, "softmax_grad_gemm: c_coef missing");
  PGICA_REQUIRE(scale > 0.f, "softmax_grad_gemm: scale must be positive");
  const int cluster = choose_cluster(k);
  if (cluster > 1 && workspace != nullptr && workspace_bytes > 0)
    return sggx_dispatch(cluster, x, y, mx, my, k, scale, r_lse, r_coef, r_tgt, c_lse, c_coef, c_tgt, out, out_is_bf16,
                         workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  SggParams p{};
  p.mx = (int)mx;
  p.my = (int)my;
  p.k = (int)k;
  p.num_tiles = (int)ceil_div(my, kBT);
  p.passes = (int)ceil_div(k, kBD);
  p.ldo = (int)k;
  p.out_bf16 = out_is_bf16;
  p.c = scale * kLog2e;
  p.r_lse = r_lse;
  p.r_coef = r_coef;
  p.r_tgt = r_tgt;
  p.c_lse = c_lse;
  p.c_coef = c_coef;
  p.c_tgt = c_tgt;
  p.out = out;
  CUtensorMap tm_x, tm_y;
  rc = make_tmap_bf16(&tm_x, x, mx, k, k, 128);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_y, y, my, k, k, 128);
  if (rc != PGICA_OK) return rc;
  const int64_t grid = ceil_div(mx, kBM) * p.passes;
  PGICA_REQUIRE(grid < (1ll << 31), "softmax_grad_gemm: grid too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define PGICA_LAUNCH_SGG(R, C)                                                                                    \
  do {                                                                                                            \
    PGICA_CUDA_OK(cudaFuncSetAttribute(sgg_kernel<R, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                       (int)kSggSmem));                                                           \
    sgg_kernel<R, C><<<(unsigned)grid, kThreads, kSggSmem, st>>>(tm_x, tm_y, p);                                   \
  } while (0)
  if (row && col)
    PGICA_LAUNCH_SGG(true, true);
  else if (row)
    PGICA_LAUNCH_SGG(true, false);
  else
    PGICA_LAUNCH_SGG(false, true);
#undef PGICA_LAUNCH_SGG
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}
