// Cluster form of the softmax-gradient GEMM (see sgg.cu for the maths): the C CTAs of a thread-block cluster own
// the C 256-column slices of one 128-row block of Out, and SHARE the recomputed G tiles through distributed shared
// memory, so the logits of a (128 x 128) tile are recomputed once per cluster instead of once per 256 output
// columns.  Tile j is produced by CTA (j mod C): MMA1 -> Z in TMEM -> epilogue -> bf16 G tile written into slot
// (j mod C) of EVERY CTA's G buffers (st.shared::cluster), signalled with cluster-scope mbarrier arrives.  Every CTA
// then runs MMA2 (Out_slice += G * Y[tile, slice], N = 256 in one instruction, the tile taken in two 64-row halves)
// for every tile.  A round = C consecutive tiles; a CTA issues MMA1 for its tile of round r, then the MMA2s of round
// r-1, so the tensor pipe never waits for an epilogue.  The epilogue keeps the freshly computed G tile in registers
// until its slot is free everywhere, so the exp/pack work overlaps the wait instead of following it.
#include "common.h"
#include "ptx.cuh"

namespace pgica {
namespace {

constexpr int kBM = 128;
constexpr int kBT = 128;
constexpr int kBD = 256;
constexpr int kBK = 64;
constexpr uint32_t kChunkBytes = 128 * kBK * 2;   // 16 KB
constexpr uint32_t kSlotBytes = 2 * kChunkBytes;  // 32 KB
constexpr uint32_t kHalfBoxBytes = 64 * kBK * 2;  // 8 KB: a [64][64] bf16 box (MMA2 operand, one 64-column chunk)
constexpr uint32_t kPBytes = kBM * kBT * 2;       // 32 KB
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr uint32_t kTmemOut = 0, kTmemZ = 256;

#ifdef PGICA_TRACE
// Debug build only (-DPGICA_TRACE): per-CTA cycle accounting of the three roles, [grid][16] int64.
__device__ long long* g_sgg_trace = nullptr;
__device__ __forceinline__ long long trace_now() {
#ifdef __CUDA_ARCH__
  return clock64();
#else
  return 0;
#endif
}
struct Lap {
  long long t, acc[6];
  __device__ Lap() : t(trace_now()) {
    for (int i = 0; i < 6; ++i) acc[i] = 0;
  }
  __device__ void operator()(int i) {
    const long long n = trace_now();
    acc[i] += n - t;
    t = n;
  }
  __device__ void flush(int base, int n) {
    if (g_sgg_trace)
      for (int i = 0; i < n; ++i) g_sgg_trace[(size_t)blockIdx.x * 16 + base + i] = acc[i];
  }
};
#define LAP(i) lap(i)
#define LAP_DECL Lap lap
#define LAP_FLUSH(base, n) lap.flush(base, n)
#else
#define LAP(i) ((void)0)
#define LAP_DECL ((void)0)
#define LAP_FLUSH(base, n) ((void)0)
#endif

struct SggcParams {
  int mx, my, k, num_tiles, passes, ldo, out_bf16, prefetch_y;
  float c;
  const float* r_lse;
  const float* r_coef;
  const int* r_tgt;
  const float* c_lse;
  const float* c_coef;
  const int* c_tgt;
  void* out;
};

template <int C>
struct Cfg {
  static constexpr int kRing = 7 - C;  // ring + G slots = 7 x 32 KB
  static constexpr size_t kSmem = 1024 + 7 * kSlotBytes + 3 * kBT * 4 + 256;
};

template <int C, bool kRow, bool kCol>
__global__ void __launch_bounds__(kThreads, 1)
sgg_cluster_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                   const __grid_constant__ CUtensorMap tm_y2, const SggcParams p) {
  constexpr int kRing = Cfg<C>::kRing;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* ring = smem;
  uint8_t* g_tiles = smem + kRing * kSlotBytes;  // C slots, slot s holds the tile produced by CTA s
  float* s_cl = reinterpret_cast<float*>(g_tiles + C * kPBytes);
  float* s_cc = s_cl + kBT;
  int* s_ct = reinterpret_cast<int*>(s_cc + kBT);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ct + kBT);
  uint64_t* empty_bar = full_bar + kRing;
  uint64_t* zfull_bar = empty_bar + kRing;  // [2]
  uint64_t* zempty_bar = zfull_bar + 2;     // [2]
  uint64_t* gfull_bar = zempty_bar + 2;     // [C]  tile in slot s has landed (1 arrival + 32 KB of copy bytes)
  uint64_t* gfree_bar = gfull_bar + C;      // [1]  every CTA has consumed MY slot (C commit arrivals)
  uint64_t* out_bar = gfree_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = cluster_ctarank();
  const int cluster_id = blockIdx.x / C;
  const int m_blk = cluster_id / p.passes;
  const int pass = cluster_id - m_blk * p.passes;
  const int out_col0 = (pass * C + (int)q) * kBD;
  const int num_kb = (p.k + kBK - 1) / kBK;
  const int J = p.num_tiles;
  const int rounds = (J + C - 1) / C;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_y2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&zfull_bar[i], 1);
      mbar_init(&zempty_bar[i], 128);
    }
    for (int i = 0; i < C; ++i) mbar_init(&gfull_bar[i], 1);
    mbar_init(gfree_bar, C);
    mbar_init(out_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anyone arrives on / writes into a peer
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the schedule (warp-uniform control flow, so addresses and coordinates stay in uniform
    // registers); one elected lane issues the copies.
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == kRing) {
        slot = 0;
        phase ^= 1;
      }
    };
    LAP_DECL;
    for (int r = 0; r <= rounds; ++r) {
      const int own = r * C + (int)q;
      if (own < J) {
        // A large Y (the LM-head weight in dH) streams from HBM: one cluster in eight pulls the tile this CTA will
        // need two rounds from now into L2, so that nobody's ring stalls on a DRAM round trip.
        const int ahead = own + 2 * C;
        const bool pf = p.prefetch_y && ahead < J && (((own / C) ^ cluster_id) & 7) == 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          LAP(0);
          mbar_wait(&empty_bar[slot], phase ^ 1);
          LAP(1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[slot], kSlotBytes);
            uint8_t* dst = ring + slot * kSlotBytes;
            tma_load_2d(dst, &tm_x, &full_bar[slot], kb * kBK, m_blk * kBM);
            tma_load_2d(dst + kChunkBytes, &tm_y, &full_bar[slot], kb * kBK, own * kBT);
            if (pf) tma_prefetch_2d(&tm_y, kb * kBK, ahead * kBT);
          }
          __syncwarp();
          advance();
        }
      }
      if (r > 0) {
        const int t_end = min(r * C, J);
        for (int t = (r - 1) * C; t < t_end; ++t) {
          for (int h = 0; h < 2; ++h) {  // Y[64-row half h of tile t, out_col0 .. +256): four [64][64] boxes per slot
            LAP(0);
            mbar_wait(&empty_bar[slot], phase ^ 1);
            LAP(2);
            if (elect_one()) {
              mbar_expect_tx(&full_bar[slot], kSlotBytes);
              uint8_t* dst = ring + slot * kSlotBytes;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                tma_load_2d(dst + i * kHalfBoxBytes, &tm_y2, &full_bar[slot], out_col0 + i * kBK, t * kBT + h * 64);
            }
            __syncwarp();
            advance();
          }
        }
      }
    }
    LAP(0);
    if (lane == 0) LAP_FLUSH(0, 3);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc1 = make_idesc_bf16(kBM, kBT, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(kBM, kBD, 0, 1);
    const uint64_t desc_k = make_smem_desc(0, 16, 1024);              // K-major operand, start address 0
    const uint64_t desc_mn = make_smem_desc(0, kHalfBoxBytes, 1024);  // MN-major operand: 64-column chunks 8 KB apart
    const uint32_t gfree_remote0 = smem_u32(gfree_bar);
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++slot == kRing) {
        slot = 0;
        phase ^= 1;
      }
    };
    int zb = 0;
    uint32_t zphase = 0;
    LAP_DECL;
    for (int r = 0; r <= rounds; ++r) {
      const int own = r * C + (int)q;
      if (own < J) {
        LAP(0);
        mbar_wait(&zempty_bar[zb], zphase ^ 1);
        LAP(1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + kTmemZ + zb * kBT;
        for (int kb = 0; kb < num_kb; ++kb) {
          LAP(0);
          mbar_wait(&full_bar[slot], phase);
          LAP(2);
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
        if (++zb == 2) {
          zb = 0;
          zphase ^= 1;
        }
      }
      if (r > 0) {
        const int t_end = min(r * C, J);
        const uint32_t gphase = (uint32_t)(r - 1) & 1u;
        for (int t = (r - 1) * C; t < t_end; ++t) {
          const int s = t - (r - 1) * C;  // producer CTA == G slot
          LAP(0);
          mbar_wait_cluster(&gfull_bar[s], gphase);
          LAP(3);
          const uint32_t g_addr = smem_u32(g_tiles + s * kPBytes);
          for (int h = 0; h < 2; ++h) {
            LAP(0);
            mbar_wait(&full_bar[slot], phase);
            LAP(4);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t y_addr = smem_u32(ring + slot * kSlotBytes);
              // half h of the G tile = its k-chunk h ([128 rows][64 vocab]); four K=16 steps, N = 256 each
              const uint64_t dg = desc_k | (((g_addr + h * kChunkBytes) >> 4) & 0x3FFF);
              const uint64_t dy = desc_mn | ((y_addr >> 4) & 0x3FFF);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16_ss(tmem_base + kTmemOut, dg + ks * 2, dy + ks * (16 * 128 >> 4), idesc2,
                             (t | h | ks) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[slot]);
              // after the second half: tell the producer of slot s that this CTA is done reading it
              if (h == 1) umma_commit_remote(mapa_u32(gfree_remote0, (uint32_t)s));
            }
            __syncwarp();
            advance();
          }
        }
      }
    }
    if (elect_one()) umma_commit(out_bar);
    __syncwarp();
    LAP(0);
    if (lane == 0) LAP_FLUSH(3, 5);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue for this CTA's own tiles
    const int quarter = warp & 3;
    const int et = threadIdx.x - 128;
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
    float nl = 0.f, nc = 0.f;
    int nt = -1;
    if ((int)q < J) load_col((int)q, nl, nc, nt);
    // my G slot (index q): written locally, then bulk-copied over DSMEM into the same slot of every peer
    const uint32_t g_local = smem_u32(g_tiles + q * kPBytes);
    int zb = 0;
    uint32_t zphase = 0, fphase = 0;
    LAP_DECL;
    for (int r = 0; r < rounds; ++r) {
      const int own = r * C + (int)q;
      if (own >= J) break;
      LAP(0);
      if (kCol) {
        s_cl[et] = nl;
        s_cc[et] = nc;
        s_ct[et] = nt;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (own + C < J) load_col(own + C, nl, nc, nt);
      }
      mbar_wait(&zfull_bar[zb], zphase);
      LAP(1);
      tc_fence_after_sync();
      const int col0 = own * kBT;
      const int rrel = rt - col0;
      uint32_t gp[kBT / 2];  // this thread's row of the G tile, bf16 pairs, held in registers until the slot is free
#pragma unroll
      for (int ch = 0; ch < kBT / 32; ++ch) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + lane_addr + kTmemZ + zb * kBT + ch * 32, rr);
        tmem_ld_wait();
        float g[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float t = __uint_as_float(rr[jj]) * p.c;
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
#pragma unroll
        for (int i = 0; i < 16; ++i) gp[ch * 16 + i] = pack_bf16x2(g[2 * i], g[2 * i + 1]);
      }
      tc_fence_before_sync();
      mbar_arrive(&zempty_bar[zb]);  // Z buffer is free for the MMA1 after next
      LAP(3);
      mbar_wait_cluster(gfree_bar, fphase ^ 1);  // every CTA has finished MMA2 on my previous tile
      LAP(2);
      fphase ^= 1;
#pragma unroll
      for (int ch = 0; ch < kBT / 32; ++ch) {
        const uint32_t chunk_off = (ch >> 1) * kChunkBytes;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t off = chunk_off + sw128_offset(row_in_blk, (ch & 1) * 4 + c4);
          st_smem_v4(g_local + off, gp[ch * 16 + c4 * 4 + 0], gp[ch * 16 + c4 * 4 + 1], gp[ch * 16 + c4 * 4 + 2],
                     gp[ch * 16 + c4 * 4 + 3]);
        }
      }
      fence_proxy_async_smem();  // my generic-proxy writes -> visible to the async proxy (UMMA reads, bulk copy)
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the whole tile is in shared memory
      LAP(4);
      if (et == 0) {
        mbar_arrive(&gfull_bar[q]);  // local consumer
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (c == (int)q) continue;
          const uint32_t rbar = mapa_u32(smem_u32(&gfull_bar[q]), (uint32_t)c);
          mbar_expect_tx_remote(rbar, kPBytes);
          dsmem_bulk_copy(mapa_u32(g_local, (uint32_t)c), g_local, kPBytes, rbar);
        }
      }
      if (++zb == 2) {
        zb = 0;
        zphase ^= 1;
      }
      if (kCol) asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // every peer's "done with your slot" commit must have landed before this CTA may exit
    if ((int)q < J) mbar_wait_cluster(gfree_bar, fphase ^ 1);
    LAP(0);
    if (et == 0) LAP_FLUSH(8, 5);
    // ------------------------------------------------------------------ final: Out slice (TMEM) -> global
    mbar_wait(out_bar, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int ch = 0; ch < kBD / 32; ++ch) {
      uint32_t rr[32];
      tmem_ld_32x32(tmem_base + lane_addr + kTmemOut + ch * 32, rr);
      tmem_ld_wait();
      const int col = out_col0 + ch * 32;
      if (row < p.mx) {
        if (p.out_bf16) {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(rr[c4 * 8 + 0]), __uint_as_float(rr[c4 * 8 + 1]));
            v.y = pack_bf16x2(__uint_as_float(rr[c4 * 8 + 2]), __uint_as_float(rr[c4 * 8 + 3]));
            v.z = pack_bf16x2(__uint_as_float(rr[c4 * 8 + 4]), __uint_as_float(rr[c4 * 8 + 5]));
            v.w = pack_bf16x2(__uint_as_float(rr[c4 * 8 + 6]), __uint_as_float(rr[c4 * 8 + 7]));
            dst[c4] = v;
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<float*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_uint4(rr[c4 * 4], rr[c4 * 4 + 1], rr[c4 * 4 + 2], rr[c4 * 4 + 3]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit while a peer can still arrive on its barriers
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int C, bool kRow, bool kCol>
int launch(const CUtensorMap& tm_x, const CUtensorMap& tm_y, const CUtensorMap& tm_y2, const SggcParams& p,
           int64_t clusters, cudaStream_t st) {
  auto kern = sgg_cluster_kernel<C, kRow, kCol>;
  PGICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<C>::kSmem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * C));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<C>::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PGICA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm_x, tm_y, tm_y2, p));
  count_launches(1);
  return PGICA_OK;
}

}  // namespace

#ifdef PGICA_TRACE
extern "C" int pgica_debug_set_sgg_trace(void* buf) {
  long long* p = static_cast<long long*>(buf);
  return cudaMemcpyToSymbol(g_sgg_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1;
}
#endif

// Called by pgica_softmax_grad_gemm when k is a multiple of 256*C.  Returns PGICA_OK or an error code.
int sgg_cluster_dispatch(int cluster, const void* x, const void* y, int64_t mx, int64_t my, int64_t k, float scale,
                         const float* r_lse, const float* r_coef, const int32_t* r_tgt, const float* c_lse,
                         const float* c_coef, const int32_t* c_tgt, void* out, int out_is_bf16, cudaStream_t st) {
  SggcParams p{};
  p.mx = (int)mx;
  p.my = (int)my;
  p.k = (int)k;
  p.num_tiles = (int)ceil_div(my, kBT);
  p.passes = (int)(k / (kBD * cluster));
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
  p.prefetch_y = (my * k * 2 > (int64_t)(32 << 20)) ? 1 : 0;  // only an operand that cannot sit in L2
  CUtensorMap tm_x, tm_y, tm_y2;
  int rc = make_tmap_bf16(&tm_x, x, mx, k, k, 128);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_y, y, my, k, k, 128);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_y2, y, my, k, k, 64);  // 64-row boxes: one K-half of a tile for MMA2
  if (rc != PGICA_OK) return rc;
  const int64_t clusters = ceil_div(mx, kBM) * p.passes;
  PGICA_REQUIRE(clusters * cluster < (1ll << 31), "softmax_grad_gemm: grid too large");
  const bool row = r_lse != nullptr, col = c_lse != nullptr;
#define PGICA_SGGC(CC)                                                        \
  do {                                                                        \
    if (row && col) return launch<CC, true, true>(tm_x, tm_y, tm_y2, p, clusters, st); \
    if (row) return launch<CC, true, false>(tm_x, tm_y, tm_y2, p, clusters, st);     \
    return launch<CC, false, true>(tm_x, tm_y, tm_y2, p, clusters, st);              \
  } while (0)
  if (cluster == 2) PGICA_SGGC(2);
  if (cluster == 4) PGICA_SGGC(4);
#undef PGICA_SGGC
  set_error("softmax_grad_gemm: unsupported cluster size %d", cluster);
  return PGICA_ERR_INVALID_ARGUMENT;
}

}  // namespace pgica
