// Dual softmax-gradient GEMM: BOTH products of a backward pass from ONE recomputation of the logits.
//
//   OutX[i, :] = sum_j G(i, j) * Y[j, :]        (DPO: dH, X = hidden, Y = LM-head weight;  NT-Xent: dA)
//   OutY[j, :] = sum_i G(i, j) * X[i, :]        (DPO: dW;                                   NT-Xent: dB)
//   G(i, j) as in sgg.cu (row term and/or column term), z_ij = <X_i, Y_j> recomputed on the tensor cores.
//
// sggx_kernel (sgg_x.cu) runs once per product and therefore recomputes every 128 x 128 tile of logits twice:
// 4 GEMM-units of tensor work for 2 units of result.  Here a tile of G is produced once and consumed twice, 3 units.
// TMEM is what makes that possible and what shapes the kernel: an SM holds 128 x 512 fp32 accumulators, so the
// 148 SMs together hold 74 slices of (128 rows x 1024 columns).  The whole GPU runs ONE persistent cooperative
// grid whose CTAs take one of three roles for the lifetime of the launch:
//
//   X-holders  R*S CTAs   (S = k/512) keep OutX of R row blocks resident in TMEM for a whole chunk of R row
//                         blocks and accumulate OutX[r] += G(r, c) * Y[c]  for every column tile c,
//   Y-holders  Cw*S CTAs  keep OutY of Cw column tiles resident for one pass over the chunk's R row blocks and
//                         accumulate OutY[c] += G(r, c)^T * X[r]  (G^T is the same tile read MN-major),
//   producers  the rest   recompute Z = X[r] Y[c, c+1]^T (N = 256), turn it into two bf16 G tiles and publish them
//                         with a TMA store into an exchange ring in global memory (a few MB, overwritten every few
//                         tiles, L2-resident) followed by a release store to a per-slot flag.
//
// Consumers poll the flag (acquire), pull the tile through TMA into shared memory and feed tcgen05.mma; when the
// tile has landed they bump a per-slot counter that lets the producer reuse the slot.  Within a pass the tiles are
// produced along diagonals (time step t: row block (cp + t) mod Rc for column pair cp), so every Y-holder consumes
// one tile per time step and every X-holder at most two: consumers run at a steady rate and the exchange ring
// stays short.  Nothing tile-sized beyond that ring ever reaches memory.  No reduction through memory is needed:
// OutX is complete when its chunk ends, OutY when its pass ends (and is accumulated in place, fp32, when X needs
// more than one chunk).
//
// Progress: every role walks the tiles in the same global production order q.  The lowest-numbered unfinished tile
// can always be produced (its slot's previous tenant has a lower number) and consumed (its consumers have nothing
// older to wait for), so the grid cannot deadlock provided all CTAs are co-resident — hence the cooperative launch.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace pgica {
namespace {

constexpr int kBM = 128;
constexpr int kBT = 128;
constexpr int kBK = 64;
constexpr int kNC = 512;                              // output columns per consumer CTA (all of its TMEM)
constexpr uint32_t kChunkBytes = 128 * kBK * 2;       // 16 KB: a [128][64] bf16 box
constexpr uint32_t kPBytes = kBM * kBT * 2;           // 32 KB: one G tile
constexpr uint32_t kPStageBytes = 3 * kChunkBytes;    // producer ring slot: X chunk + 256-row Y chunk
constexpr int kPRing = 4;
constexpr uint32_t kBox32Bytes = 32 * kBK * 2;        // 4 KB: a [32][64] box
constexpr uint32_t kCStageBytes = 8 * kBox32Bytes;    // consumer ring slot: 32 operand rows x 512 columns
constexpr int kCRing = 5;
constexpr int kStagesPerTile = kBT / 32;              // 4
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr size_t kSmem = 1024 + 7 * 32768 + 3 * kBT * 4 + 512;
static_assert(kPRing * kPStageBytes + kPBytes == 7 * 32768, "producer shared-memory plan");
static_assert(2 * kPBytes + kCRing * kCStageBytes == 7 * 32768, "consumer shared-memory plan");

#ifdef PGICA_TRACE
__device__ long long* g_sggf_trace = nullptr;
__device__ __forceinline__ long long ftrace_now() {
#ifdef __CUDA_ARCH__
  return clock64();
#else
  return 0;
#endif
}
struct FLap {
  long long t, acc[6];
  __device__ FLap() : t(ftrace_now()) {
    for (int i = 0; i < 6; ++i) acc[i] = 0;
  }
  __device__ void operator()(int i) {
    const long long n = ftrace_now();
    acc[i] += n - t;
    t = n;
  }
  __device__ void flush(int base, int n) {
    if (g_sggf_trace)
      for (int i = 0; i < n; ++i) g_sggf_trace[(size_t)blockIdx.x * 24 + base + i] = acc[i];
  }
};
#define LAP(i) lap(i)
#define LAP_DECL FLap lap
#define LAP_FLUSH(base, n) lap.flush(base, n)
#else
#define LAP(i) ((void)0)
#define LAP_DECL ((void)0)
#define LAP_FLUSH(base, n) ((void)0)
#endif

struct SggfParams {
  int mx, my, k;
  int RB, J;              // row blocks of X, column tiles of Y
  int R, Cw, S;           // row blocks per chunk, column tiles per pass, 512-column splits of k
  int nH, nW, nP, D;      // CTAs per role; exchange double-slots per producer
  int outx_bf16, outy_bf16;
  float c;                // scale * log2(e)
  const float* r_lse;
  const float* r_coef;
  const int* r_tgt;
  const float* c_lse;
  const float* c_coef;
  const int* c_tgt;
  void* out_x;
  void* out_y;
  uint32_t* ready;        // [nP*D*2] use count + 1 of the tile that is complete in the slot
  uint32_t* done;         // [nP*D*2] consumers that have pulled a tile out of the slot, ever
};

__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until *p >= v (monotonic counters); traps instead of hanging if the protocol is broken.
__device__ __forceinline__ void wait_flag_ge(const uint32_t* p, uint32_t v) {
  if (ld_acquire_gpu(p) >= v) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < v) {
    __nanosleep(40);
    if (clock64() - t0 > PGICA_WATCHDOG_CYCLES) {
      printf("pgica: exchange-flag watchdog (block %d thread %d want %u have %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, v, ld_acquire_gpu(p));
      __trap();
    }
  }
}

// Visits every double tile in production order.  f(q, chunk, pass, r0, r, c0, cp, Cc): sequence number q; the
// tile pair covers row block r0 + r and column tiles c0 + 2cp and c0 + 2cp + 1 (the latter only if 2cp + 1 < Cc).
template <class F>
__device__ __forceinline__ void for_each_dt(const SggfParams& p, F&& f) {
  int q = 0;
  for (int r0 = 0, chunk = 0; r0 < p.RB; r0 += p.R, ++chunk) {
    const int Rc = min(p.R, p.RB - r0);
    for (int c0 = 0, pass = 0; c0 < p.J; c0 += p.Cw, ++pass) {
      const int Cc = min(p.Cw, p.J - c0);
      const int Cwp = (Cc + 1) >> 1;
      for (int t = 0; t < Rc; ++t) {
        int r = t;
        for (int cp = 0; cp < Cwp; ++cp, ++q) {
          f(q, chunk, pass, r0, r, c0, cp, Cc);
          if (++r == Rc) r = 0;
        }
      }
    }
  }
}

// Visits, in production order, the tiles one consumer CTA accumulates.  is_y = false: X-holder of row block
// `idx` (chunk-relative); true: Y-holder of column tile `idx` (pass-relative).  f(q, half, rblk, ctile, first,
// period) with absolute block / tile numbers; `first` marks the first tile of an accumulation period.  g(period,
// chunk, pass) is called after the last tile of every period that had tiles.
template <class F, class G>
__device__ __forceinline__ void for_each_consumer_tile(const SggfParams& p, bool is_y, int idx, F&& f, G&& g) {
  int qbase = 0, period = 0;
  for (int r0 = 0, chunk = 0; r0 < p.RB; r0 += p.R, ++chunk) {
    const int Rc = min(p.R, p.RB - r0);
    bool first = true;
    for (int c0 = 0, pass = 0; c0 < p.J; c0 += p.Cw, ++pass) {
      const int Cc = min(p.Cw, p.J - c0);
      const int Cwp = (Cc + 1) >> 1;
      if (is_y) {
        if (idx < Cc) {
          const int cp = idx >> 1;
          int r = cp % Rc;
          for (int t = 0; t < Rc; ++t) {
            f(qbase + t * Cwp + cp, idx & 1, r0 + r, c0 + idx, t == 0, period);
            if (++r == Rc) r = 0;
          }
          g(period, chunk, pass);
          ++period;
        }
      } else if (idx < Rc) {
        for (int t = 0; t < Rc; ++t) {
          // column pairs cp with (cp + t) mod Rc == idx, increasing
          int cp = idx - t;
          if (cp < 0) cp += Rc;
          for (; cp < Cwp; cp += Rc) {
            f(qbase + t * Cwp + cp, 0, r0 + idx, c0 + 2 * cp, first, period);
            first = false;
            if (2 * cp + 1 < Cc) f(qbase + t * Cwp + cp, 1, r0 + idx, c0 + 2 * cp + 1, false, period);
          }
        }
      }
      qbase += Rc * Cwp;
    }
    if (!is_y && idx < Rc) {
      g(period, chunk, 0);
      ++period;
    }
  }
}

template <bool kRow, bool kCol>
__global__ void __launch_bounds__(kThreads, 1)
sggf_kernel(const __grid_constant__ CUtensorMap tm_x128, const __grid_constant__ CUtensorMap tm_y128,
            const __grid_constant__ CUtensorMap tm_x32, const __grid_constant__ CUtensorMap tm_y32,
            const __grid_constant__ CUtensorMap tm_s, const SggfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  float* s_cl = reinterpret_cast<float*>(smem + 7 * 32768);
  float* s_cc = s_cl + kBT;
  int* s_ct = reinterpret_cast<int*>(s_cc + kBT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ct + kBT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bid = (int)blockIdx.x;
  const bool is_producer = bid >= p.nH + p.nW;
  const int num_kb = p.k / kBK;
  const uint32_t n_consumers = 2u * (uint32_t)p.S;  // CTAs that pull each G tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x128);
    tma_prefetch_desc(&tm_y128);
    tma_prefetch_desc(&tm_x32);
    tma_prefetch_desc(&tm_y32);
    tma_prefetch_desc(&tm_s);
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);

  if (is_producer) {
    // =================================================================================================== producer
    uint8_t* ring = smem;
    uint8_t* staging = smem + kPRing * kPStageBytes;
    uint64_t* full_bar = bars;                 // [kPRing]
    uint64_t* empty_bar = full_bar + kPRing;   // [kPRing]
    uint64_t* zfull_bar = empty_bar + kPRing;  // [2]
    uint64_t* zempty_bar = zfull_bar + 2;      // [2]
    uint64_t* stfull_bar = zempty_bar + 2;
    uint64_t* stfree_bar = stfull_bar + 1;
    if (warp == 1 && lane == 0) {
      for (int i = 0; i < kPRing; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&zfull_bar[i], 1);
        mbar_init(&zempty_bar[i], 128);
      }
      mbar_init(stfull_bar, 128);
      mbar_init(stfree_bar, 1);
      fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int pid = bid - p.nH - p.nW;

    if (warp == 0) {
      // ---------------------------------------------------------------- TMA loads of the MMA1 operands
      int slot = 0, mine = pid;
      uint32_t phase = 0;
      LAP_DECL;
      for_each_dt(p, [&](int q, int, int, int r0, int r, int c0, int cp, int) {
        if (q != mine) return;
        mine += p.nP;
        const int xrow = (r0 + r) * kBM, yrow = (c0 + 2 * cp) * kBT;
        for (int kb = 0; kb < num_kb; ++kb) {
          LAP(0);
          mbar_wait(&empty_bar[slot], phase ^ 1);
          LAP(1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[slot], kPStageBytes);
            uint8_t* dst = ring + slot * kPStageBytes;
            tma_load_2d(dst, &tm_x128, &full_bar[slot], kb * kBK, xrow);
            tma_load_2d(dst + kChunkBytes, &tm_y128, &full_bar[slot], kb * kBK, yrow);
            tma_load_2d(dst + 2 * kChunkBytes, &tm_y128, &full_bar[slot], kb * kBK, yrow + kBT);
          }
          __syncwarp();
          if (++slot == kPRing) {
            slot = 0;
            phase ^= 1;
          }
        }
      });
      LAP(0);
      if (lane == 0) LAP_FLUSH(4, 2);
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA1: Z[128 x 256] = X[r] Y[c, c+1]^T
      constexpr uint32_t idesc1 = make_idesc_bf16(kBM, 2 * kBT, 0, 0);
      const uint64_t desc_k = make_smem_desc(0, 16, 1024);
      int slot = 0, mine = pid;
      uint32_t phase = 0, n = 0;
      LAP_DECL;
      for_each_dt(p, [&](int q, int, int, int, int, int, int, int) {
        if (q != mine) return;
        mine += p.nP;
        const uint32_t buf = n & 1u;
        LAP(0);
        mbar_wait(&zempty_bar[buf], ((n >> 1) & 1u) ^ 1u);  // the epilogue has read this Z buffer's previous tile
        LAP(1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * 256u;
        for (int kb = 0; kb < num_kb; ++kb) {
          LAP(0);
          mbar_wait(&full_bar[slot], phase);
          LAP(2);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t x_addr = smem_u32(ring + slot * kPStageBytes);
            const uint64_t da = desc_k | ((x_addr >> 4) & 0x3FFF);
            const uint64_t db = desc_k | (((x_addr + kChunkBytes) >> 4) & 0x3FFF);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[slot]);
            if (kb == num_kb - 1) umma_commit(&zfull_bar[buf]);
          }
          __syncwarp();
          if (++slot == kPRing) {
            slot = 0;
            phase ^= 1;
          }
        }
        ++n;
      });
      LAP(0);
      if (lane == 0) LAP_FLUSH(0, 3);
    } else if (warp == 3) {
      // ---------------------------------------------------------------- exchange: staging -> ring slot -> flag
      int mine = pid;
      uint32_t n = 0, ntile = 0;
      LAP_DECL;
      for_each_dt(p, [&](int q, int, int, int, int, int, int cp, int Cc) {
        if (q != mine) return;
        mine += p.nP;
        const uint32_t ds = n % (uint32_t)p.D, use = n / (uint32_t)p.D;
        for (int half = 0; half < 2; ++half) {
          const uint32_t tslot = ((uint32_t)pid * p.D + ds) * 2u + half;
          if (2 * cp + half >= Cc) {
            // a pass with an odd number of column tiles: nobody will pull this use of the slot; keep its counter in step
            if (lane == 0) red_release_gpu_add(p.done + tslot, n_consumers);
            break;
          }
          LAP(0);
          mbar_wait(stfull_bar, ntile & 1u);
          LAP(1);
          if (use > 0) wait_flag_ge(p.done + tslot, n_consumers * use);  // every consumer has pulled the old tile
          LAP(2);
          if (lane == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");  // acquire above before the async-proxy write
            tma_store_2d(&tm_s, staging, 0, (int)tslot * kBM);
            tma_store_2d(&tm_s, staging + kChunkBytes, kBK, (int)tslot * kBM);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging has been read
            mbar_arrive(stfree_bar);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the tile is complete in global memory
            asm volatile("fence.proxy.async.global;" ::: "memory");
            st_release_gpu(p.ready + tslot, use + 1u);
          }
          __syncwarp();
          LAP(3);
          ++ntile;
        }
        ++n;
      });
      LAP(0);
      if (lane == 0) LAP_FLUSH(6, 4);
    } else if (warp >= 4) {
      // ---------------------------------------------------------------- epilogue: Z -> G (bf16) -> staging
      const int quarter = warp & 3;
      const int et = threadIdx.x - 128;
      const int row_in_blk = quarter * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t g_local = smem_u32(staging);
      int mine = pid;
      uint32_t n = 0, ntile = 0;
      LAP_DECL;
      for_each_dt(p, [&](int q, int, int, int r0, int r, int c0, int cp, int Cc) {
        if (q != mine) return;
        mine += p.nP;
        const int row = (r0 + r) * kBM + row_in_blk;
        float rl = 0.f, rc = 0.f;
        int rt = -1;
        if (kRow && row < p.mx) {
          rl = p.r_lse[row] * kLog2e;
          rc = p.r_coef[row];
          rt = p.r_tgt ? p.r_tgt[row] : -1;
        }
        const uint32_t buf = n & 1u;
        LAP(0);
        mbar_wait(&zfull_bar[buf], (n >> 1) & 1u);
        LAP(1);
        tc_fence_after_sync();
        for (int half = 0; half < 2; ++half) {
          if (2 * cp + half >= Cc) break;
          const int col0 = (c0 + 2 * cp + half) * kBT;
          if (kCol) {
            const int col = col0 + et;
            float l = 0.f, cf = 0.f;
            int tg = -1;
            if (col < p.my) {
              l = p.c_lse[col] * kLog2e;
              cf = p.c_coef[col];
              tg = p.c_tgt ? p.c_tgt[col] : -1;
            }
            s_cl[et] = l;
            s_cc[et] = cf;
            s_ct[et] = tg;
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
          const int rrel = rt - col0;
          uint32_t gp[kBT / 2];
#pragma unroll
          for (int ch = 0; ch < kBT / 32; ++ch) {
            uint32_t rr[32];
            tmem_ld_32x32(tmem_base + lane_addr + buf * 256u + half * kBT + ch * 32, rr);
            tmem_ld_wait();
            float g[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const float tz = __uint_as_float(rr[jj]) * p.c;
              float v = 0.f;
              if (kRow) v = rc * fast_exp2(tz - rl);
              if (kCol) {
                const int cj = ch * 32 + jj;
                const float ccj = s_cc[cj];
                v = fmaf(ccj, fast_exp2(tz - s_cl[cj]), v);
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
          if (half == 1 || 2 * cp + 1 >= Cc) {
            tc_fence_before_sync();
            mbar_arrive(&zempty_bar[buf]);  // this Z buffer may be overwritten by the MMA1 after next
          }
          LAP(2);
          mbar_wait(stfree_bar, (ntile & 1u) ^ 1u);  // the exchange warp's TMA store has read the previous tile
          LAP(3);
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
          fence_proxy_async_smem();
          mbar_arrive(stfull_bar);
          ++ntile;
          if (kCol) asm volatile("bar.sync 3, 128;" ::: "memory");  // the column statistics may be overwritten
          LAP(4);
        }
        ++n;
      });
      LAP(0);
      if (et == 0) LAP_FLUSH(10, 5);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
    return;
  }

  // ===================================================================================================== consumers
  const bool is_y = bid >= p.nH;
  const int cidx = (is_y ? bid - p.nH : bid) / p.S;  // row block in chunk (X-holder) / column tile in pass (Y-holder)
  const int split = (is_y ? bid - p.nH : bid) % p.S;  // which 512 output columns
  uint8_t* gbuf = smem;                       // two G tiles
  uint8_t* ring = smem + 2 * kPBytes;         // kCRing operand stages
  uint64_t* full_bar = bars;                  // [kCRing]
  uint64_t* empty_bar = full_bar + kCRing;    // [kCRing]
  uint64_t* gfull_bar = empty_bar + kCRing;   // [2]
  uint64_t* gempty_bar = gfull_bar + 2;       // [2]
  uint64_t* outfull_bar = gempty_bar + 2;
  uint64_t* outfree_bar = outfull_bar + 1;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kCRing; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&gfull_bar[i], 1);
      mbar_init(&gempty_bar[i], 1);
    }
    mbar_init(outfull_bar, 1);
    mbar_init(outfree_bar, 128);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const CUtensorMap* tm_op = is_y ? &tm_x32 : &tm_y32;  // the other operand of this CTA's product
  const int op_col0 = split * kNC;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: G tiles from the ring + operand rows
    int slot = 0;
    uint32_t phase = 0, n = 0;
    LAP_DECL;
    for_each_consumer_tile(
        p, is_y, cidx,
        [&](int q, int half, int rblk, int ctile, bool, int) {
          const uint32_t prod = (uint32_t)q % (uint32_t)p.nP, i = (uint32_t)q / (uint32_t)p.nP;
          const uint32_t ds = i % (uint32_t)p.D, use = i / (uint32_t)p.D;
          const uint32_t tslot = (prod * p.D + ds) * 2u + (uint32_t)half;
          const uint32_t gb = n & 1u;
          LAP(0);
          mbar_wait(&gempty_bar[gb], ((n >> 1) & 1u) ^ 1u);  // the MMAs of the tile before last are complete
          LAP(1);
          wait_flag_ge(p.ready + tslot, use + 1u);           // the tile is complete in the ring
          LAP(2);
          if (elect_one()) {
            asm volatile("fence.proxy.async.global;" ::: "memory");  // acquire above before the async-proxy read
            mbar_expect_tx(&gfull_bar[gb], kPBytes);
            tma_load_2d(gbuf + gb * kPBytes, &tm_s, &gfull_bar[gb], 0, (int)tslot * kBM);
            tma_load_2d(gbuf + gb * kPBytes + kChunkBytes, &tm_s, &gfull_bar[gb], kBK, (int)tslot * kBM);
          }
          __syncwarp();
          const int op_row0 = (is_y ? rblk : ctile) * kBM;
          for (int st = 0; st < kStagesPerTile; ++st) {
            LAP(0);
            mbar_wait(&empty_bar[slot], phase ^ 1);
            LAP(3);
            if (elect_one()) {
              mbar_expect_tx(&full_bar[slot], kCStageBytes);
              uint8_t* dst = ring + slot * kCStageBytes;
#pragma unroll
              for (int b = 0; b < 8; ++b)
                tma_load_2d(dst + b * kBox32Bytes, tm_op, &full_bar[slot], op_col0 + b * kBK, op_row0 + st * 32);
            }
            __syncwarp();
            if (++slot == kCRing) {
              slot = 0;
              phase ^= 1;
            }
          }
          ++n;
        },
        [&](int, int, int) {});
    LAP(0);
    if (lane == 0) LAP_FLUSH(4, 4);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA2: Out[128 x 512] += G(^T) * operand rows
    const uint32_t idesc2 = make_idesc_bf16(kBM, 256, is_y ? 1 : 0, 1);
    const uint64_t desc_ak = make_smem_desc(0, 16, 1024);              // G as A, K-major (X-holder)
    const uint64_t desc_amn = make_smem_desc(0, kChunkBytes, 1024);    // G^T as A: the same tile read MN-major
    const uint64_t desc_b = make_smem_desc(0, kBox32Bytes, 1024);      // operand rows, MN-major, 64-col boxes 4 KB apart
    int slot = 0;
    uint32_t phase = 0, n = 0;
    LAP_DECL;
    for_each_consumer_tile(
        p, is_y, cidx,
        [&](int q, int half, int, int, bool first, int period) {
          const uint32_t gb = n & 1u;
          if (first && period > 0) {
            LAP(0);
            mbar_wait(outfree_bar, (uint32_t)(period - 1) & 1u);  // the previous accumulator has left TMEM
            LAP(1);
            tc_fence_after_sync();
          }
          LAP(0);
          mbar_wait(&gfull_bar[gb], (n >> 1) & 1u);
          LAP(2);
          tc_fence_after_sync();
          if (elect_one()) {
            // the ring slot may be overwritten: this CTA has its copy
            const uint32_t prod = (uint32_t)q % (uint32_t)p.nP, i = (uint32_t)q / (uint32_t)p.nP;
            red_release_gpu_add(p.done + (prod * p.D + i % (uint32_t)p.D) * 2u + (uint32_t)half, 1u);
          }
          __syncwarp();
          const uint32_t g_addr = smem_u32(gbuf + gb * kPBytes);
          const uint64_t dg = (is_y ? desc_amn : desc_ak) | ((g_addr >> 4) & 0x3FFF);
          for (int st = 0; st < kStagesPerTile; ++st) {
            LAP(0);
            mbar_wait(&full_bar[slot], phase);
            LAP(3);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t y_addr = smem_u32(ring + slot * kCStageBytes);
              const uint64_t dy = desc_b | ((y_addr >> 4) & 0x3FFF);
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const int ks = st * 2 + kk;  // K step of 16 within the tile's 128
                const uint64_t da = is_y ? dg + ks * (2048 >> 4) : dg + (ks >> 2) * (kChunkBytes >> 4) + (ks & 3) * 2;
                const uint32_t acc = (first && ks == 0) ? 0u : 1u;
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  umma_bf16_ss(tmem_base + h * 256u, da, dy + kk * (2048 >> 4) + h * (4 * kBox32Bytes >> 4), idesc2, acc);
              }
              umma_commit(&empty_bar[slot]);
              if (st == kStagesPerTile - 1) umma_commit(&gempty_bar[gb]);
            }
            __syncwarp();
            if (++slot == kCRing) {
              slot = 0;
              phase ^= 1;
            }
          }
          ++n;
        },
        [&](int, int, int) {
          if (elect_one()) umma_commit(outfull_bar);
          __syncwarp();
        });
    LAP(0);
    if (lane == 0) LAP_FLUSH(0, 4);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ drain: accumulator -> OutX / OutY
    const int quarter = warp & 3;
    const int row_in_blk = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    LAP_DECL;
    for_each_consumer_tile(
        p, is_y, cidx, [&](int, int, int, int, bool, int) {},
        [&](int period, int chunk, int pass) {
          const int blk = is_y ? pass * p.Cw + cidx : chunk * p.R + cidx;
          const int row = blk * kBM + row_in_blk;
          const int limit = is_y ? p.my : p.mx;
          void* out = is_y ? p.out_y : p.out_x;
          const bool bf16 = (is_y ? p.outy_bf16 : p.outx_bf16) != 0;
          const bool accumulate = is_y && chunk > 0;
          LAP(0);
          mbar_wait(outfull_bar, (uint32_t)period & 1u);
          LAP(1);
          tc_fence_after_sync();
#pragma unroll 1
          for (int ch = 0; ch < kNC / 32; ++ch) {
            uint32_t rr[32];
            tmem_ld_32x32(tmem_base + lane_addr + ch * 32, rr);
            tmem_ld_wait();
            const int col = op_col0 + ch * 32;
            if (row < limit) {
              if (bf16) {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + (size_t)row * p.k + col);
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
                float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + (size_t)row * p.k + col);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                  float4 v = make_float4(__uint_as_float(rr[c4 * 4]), __uint_as_float(rr[c4 * 4 + 1]),
                                         __uint_as_float(rr[c4 * 4 + 2]), __uint_as_float(rr[c4 * 4 + 3]));
                  if (accumulate) {
                    const float4 o = dst[c4];
                    v.x += o.x;
                    v.y += o.y;
                    v.z += o.z;
                    v.w += o.w;
                  }
                  dst[c4] = v;
                }
              }
            }
          }
          tc_fence_before_sync();
          mbar_arrive(outfree_bar);
          LAP(2);
        });
    if (threadIdx.x == 128) LAP_FLUSH(10, 3);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------- host side
struct Plan {
  int R, Cw, nH, nW, nP;
  double cost;
};

// Tensor-pipe time model in units of one 128 x 256 x 16 instruction (~171 cycles): a consumer spends 16 per tile,
// a producer k/16 per double tile; a drain costs about 24.  The slowest role sets the pace of a chunk.
Plan choose_plan(int RB, int J, int k, int nsm) {
  const int S = k / kNC;
  const double per_dt = k / 16.0;
  Plan best{0, 0, 0, 0, 0, 1e300};
  for (int R = 1; R <= RB && R * S <= nsm - S - 1; ++R) {
    const int nH = R * S;
    for (int Cw = 2; Cw * S <= nsm - nH - 1; Cw += 2) {
      if (Cw > J + 1 && Cw > 2) break;
      const int nW = Cw * S, nP = nsm - nH - nW;
      const int passes = (J + Cw - 1) / Cw;
      const int last = J - (passes - 1) * Cw;
      const long pairs = (long)(passes - 1) * (Cw / 2) + (last + 1) / 2;
      double total = 0;
      for (int r0 = 0; r0 < RB; r0 += R) {
        const int Rc = RB - r0 < R ? RB - r0 : R;
        const double tH = 16.0 * J + 24;
        const double tW = passes * (16.0 * Rc + 24);
        const double tP = (double)((pairs * Rc + nP - 1) / nP) * per_dt;
        double t = tH > tW ? tH : tW;
        if (tP > t) t = tP;
        total += t + per_dt + 48;  // pipeline fill / drain of a chunk
      }
      if (total < best.cost) best = Plan{R, Cw, nH, nW, nP, total};
    }
  }
  return best;
}

int plan_override(Plan* pl, int nsm, int S) {
  // PGICA_SGGF_PLAN="R,Cw" pins the role split (tuning / tests)
  const char* e = getenv("PGICA_SGGF_PLAN");
  if (!e) return 0;
  int R = 0, Cw = 0;
  if (sscanf(e, "%d,%d", &R, &Cw) != 2 || R < 1 || Cw < 1) return 0;
  if (R * S + Cw * S >= nsm) return 0;
  pl->R = R;
  pl->Cw = Cw;
  pl->nH = R * S;
  pl->nW = Cw * S;
  pl->nP = nsm - pl->nH - pl->nW;
  return 1;
}

constexpr int kSlotsPerProducer = 4;  // double slots (two G tiles each)

template <bool kRow, bool kCol>
int launch(const CUtensorMap& tm_x128, const CUtensorMap& tm_y128, const CUtensorMap& tm_x32,
           const CUtensorMap& tm_y32, const CUtensorMap& tm_s, const SggfParams& p, int grid, cudaStream_t st) {
  auto kern = sggf_kernel<kRow, kCol>;
  static bool configured = false;
  if (!configured) {
    PGICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    int per_sm = 0;
    PGICA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, kSmem));
    if (per_sm < 1) {
      set_error("softmax_grad_gemm_dual: the kernel does not fit on an SM");
      return PGICA_ERR_CUDA;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident or the launch fails: the roles wait on each other
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PGICA_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm_x128, tm_y128, tm_x32, tm_y32, tm_s, p));
  count_launches(1);
  return PGICA_OK;
}

}  // namespace

bool sggf_supported(int64_t mx, int64_t my, int64_t k) {
  // opt-in for now (PGICA_SGG_FUSED=1): on cfg2 the two single-product launches are still faster (1.34 vs 1.43 ms)
  const char* e = getenv("PGICA_SGG_FUSED");
  if (!e || atoi(e) == 0) return false;
  return k % kNC == 0 && k / kNC <= 4 && mx >= 1 && my >= 1 && device_sm_count() >= 3 * (int)(k / kNC) + 1;
}

size_t sggf_workspace_bytes() {
  // exchange ring for the largest producer count + flags
  const size_t np = 160;
  return np * kSlotsPerProducer * 2 * (size_t)kPBytes + 2 * align_up(np * kSlotsPerProducer * 2 * sizeof(uint32_t), 256);
}

int sggf_dispatch(const void* x, const void* y, int64_t mx, int64_t my, int64_t k, float scale, const float* r_lse,
                  const float* r_coef, const int32_t* r_tgt, const float* c_lse, const float* c_coef,
                  const int32_t* c_tgt, void* out_x, int out_x_is_bf16, void* out_y, int out_y_is_bf16,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  PGICA_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                "softmax_grad_gemm_dual: workspace missing or not 256-byte aligned");
  const int nsm = device_sm_count();
  const int S = (int)(k / kNC);
  const int RB = (int)ceil_div(mx, kBM), J = (int)ceil_div(my, kBT);
  PGICA_REQUIRE((int64_t)RB * J < (1ll << 30), "softmax_grad_gemm_dual: problem too large");
  Plan pl = choose_plan(RB, J, (int)k, nsm);
  plan_override(&pl, nsm, S);
  PGICA_REQUIRE(pl.R >= 1 && pl.nP >= 1, "softmax_grad_gemm_dual: no role split for %d SMs", nsm);
  PGICA_REQUIRE(!(out_y_is_bf16 && RB > pl.R), "softmax_grad_gemm_dual: a bf16 OutY cannot be accumulated over chunks");
  SggfParams p{};
  p.mx = (int)mx;
  p.my = (int)my;
  p.k = (int)k;
  p.RB = RB;
  p.J = J;
  p.R = pl.R;
  p.Cw = pl.Cw;
  p.S = S;
  p.nH = pl.nH;
  p.nW = pl.nW;
  p.nP = pl.nP;
  p.D = kSlotsPerProducer;
  p.outx_bf16 = out_x_is_bf16;
  p.outy_bf16 = out_y_is_bf16;
  p.c = scale * kLog2e;
  p.r_lse = r_lse;
  p.r_coef = r_coef;
  p.r_tgt = r_tgt;
  p.c_lse = c_lse;
  p.c_coef = c_coef;
  p.c_tgt = c_tgt;
  p.out_x = out_x;
  p.out_y = out_y;
  const size_t nslots = (size_t)p.nP * p.D * 2;
  const size_t ring_bytes = nslots * kPBytes;
  const size_t flag_bytes = align_up(nslots * sizeof(uint32_t), 256);
  if (workspace_bytes < ring_bytes + 2 * flag_bytes) {
    set_error("softmax_grad_gemm_dual: workspace too small (%zu < %zu)", workspace_bytes, ring_bytes + 2 * flag_bytes);
    return PGICA_ERR_WORKSPACE_TOO_SMALL;
  }
  uint8_t* w = static_cast<uint8_t*>(workspace);
  p.ready = reinterpret_cast<uint32_t*>(w + ring_bytes);
  p.done = reinterpret_cast<uint32_t*>(w + ring_bytes + flag_bytes);
  PGICA_CUDA_OK(cudaMemsetAsync(p.ready, 0, 2 * flag_bytes, st));
  CUtensorMap tm_x128, tm_y128, tm_x32, tm_y32, tm_s;
  int rc = make_tmap_bf16(&tm_x128, x, mx, k, k, 128);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_y128, y, my, k, k, 128);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_x32, x, mx, k, k, 32);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_y32, y, my, k, k, 32);
  if (rc != PGICA_OK) return rc;
  rc = make_tmap_bf16(&tm_s, workspace, nslots * kBM, kBT, kBT, 128);
  if (rc != PGICA_OK) return rc;
  const int grid = p.nH + p.nW + p.nP;
  const bool row = r_lse != nullptr, col = c_lse != nullptr;
  if (row && col) return launch<true, true>(tm_x128, tm_y128, tm_x32, tm_y32, tm_s, p, grid, st);
  if (row) return launch<true, false>(tm_x128, tm_y128, tm_x32, tm_y32, tm_s, p, grid, st);
  return launch<false, true>(tm_x128, tm_y128, tm_x32, tm_y32, tm_s, p, grid, st);
}

}  // namespace pgica

#ifdef PGICA_TRACE
extern "C" int pgica_debug_set_sggf_trace(void* buf) {
  long long* p = static_cast<long long*>(buf);
  return cudaMemcpyToSymbol(pgica::g_sggf_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int pgica_softmax_grad_gemm_dual_workspace_bytes(int64_t mx, int64_t my, int64_t k, size_t* bytes_host) {
  PGICA_REQUIRE(bytes_host, "workspace query: null result pointer");
  (void)mx;
  (void)my;
  (void)k;
  *bytes_host = pgica::sggf_workspace_bytes();
  return PGICA_OK;
}

extern "C" int pgica_softmax_grad_gemm_dual(const void* x, const void* y, int64_t mx, int64_t my, int64_t k,
                                            float scale, const float* r_lse, const float* r_coef,
                                            const int32_t* r_tgt, const float* c_lse, const float* c_coef,
                                            const int32_t* c_tgt, void* out_x, int out_x_is_bf16, void* out_y,
                                            int out_y_is_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pgica;
  int rc = pgica_device_check();
  if (rc != PGICA_OK) return rc;
  PGICA_REQUIRE(x && y && out_x && out_y, "softmax_grad_gemm_dual: null operand");
  PGICA_REQUIRE(mx > 0 && my > 0 && k > 0, "softmax_grad_gemm_dual: bad shape (mx %lld my %lld k %lld)", (long long)mx,
                (long long)my, (long long)k);
  PGICA_REQUIRE(k % kNC == 0 && k / kNC <= 4, "softmax_grad_gemm_dual: k must be 512, 1024, 1536 or 2048 (got %lld)",
                (long long)k);
  PGICA_REQUIRE(mx < (1ll << 30) && my < (1ll << 30), "softmax_grad_gemm_dual: dimension too large");
  const bool row = r_lse != nullptr, col = c_lse != nullptr;
  PGICA_REQUIRE(row || col, "softmax_grad_gemm_dual: need row statistics, column statistics or both");
  PGICA_REQUIRE(!row || r_coef, "softmax_grad_gemm_dual: r_coef missing");
  PGICA_REQUIRE(!col || c_coef, "softmax_grad_gemm_dual: c_coef missing");
  PGICA_REQUIRE(scale > 0.f, "softmax_grad_gemm_dual: scale must be positive");
  return sggf_dispatch(x, y, mx, my, k, scale, r_lse, r_coef, r_tgt, c_lse, c_coef, c_tgt, out_x, out_x_is_bf16, out_y,
                       out_y_is_bf16, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
