// the finite-check + global gradient norm + clip that follows the loss heads in every micro-step
// (NaNSafeGradientNorm, pkg/models/components.py:283-318; the per-parameter isfinite() scan and clip_grad_norm_ of
// pkg/training/trainer.py:494-515 / 619-628) as three launches over ALL gradient tensors at once instead of two or
// three small kernels and a host synchronisation per parameter.  HBM-bound streaming work:
//   sumsq     one pass over every gradient (fp32 or bf16): per-chunk sum of squares; Inf / NaN propagate into the sum,
//             so "all gradients finite" == "the total norm is finite", which is exactly the reference's test
//   finalize  one block folds the per-chunk partials in a fixed order (deterministic, double accumulation):
//             out = {total_norm, clip_coef = min(1, max_norm / (total_norm + 1e-6)), is_finite}
//   scale     g *= clip_coef in place, only if finite and clip_coef < 1 (torch multiplies by 1.0 otherwise: same bits)
// Algorithmic bytes: sum(numel * sizeof) read by sumsq; read + write of the same by scale when it clips.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"

namespace pgica {
namespace {

constexpr int kChunk = 32768;  // elements per block
constexpr int kThreads = 256;

struct GradChunk {
  const void* ptr;  // first element of the chunk
  int count;        // elements in it (<= kChunk)
  int is_bf16;
};

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < kThreads / 32 ? s_red[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const GradChunk* __restrict__ chunks,
                                                              float* __restrict__ partial) {
  __shared__ float s_red[kThreads / 32];
  const GradChunk c = chunks[blockIdx.x];
  float acc = 0.f;
  if (c.is_bf16) {
    const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(c.ptr);
    const int n8 = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? c.count / 8 : 0;
    const uint4* g8 = reinterpret_cast<const uint4*>(g);
    for (int i = threadIdx.x; i < n8; i += kThreads) {
      const uint4 v = g8[i];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
        acc = fmaf(lo, lo, acc);
        acc = fmaf(hi, hi, acc);
      }
    }
    for (int i = n8 * 8 + threadIdx.x; i < c.count; i += kThreads) {
      const float x = __bfloat162float(g[i]);
      acc = fmaf(x, x, acc);
    }
  } else {
    const float* g = static_cast<const float*>(c.ptr);
    const int n4 = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? c.count / 4 : 0;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    // four independent 16-byte loads in flight per thread (a full chunk is 32 per thread)
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = threadIdx.x;
    for (; i + 3 * kThreads < n4; i += 4 * kThreads) {
      const float4 v0 = g4[i], v1 = g4[i + kThreads], v2 = g4[i + 2 * kThreads], v3 = g4[i + 3 * kThreads];
      acc = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, acc))));
      a1 = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, a1))));
      a2 = fmaf(v2.x, v2.x, fmaf(v2.y, v2.y, fmaf(v2.z, v2.z, fmaf(v2.w, v2.w, a2))));
      a3 = fmaf(v3.x, v3.x, fmaf(v3.y, v3.y, fmaf(v3.z, v3.z, fmaf(v3.w, v3.w, a3))));
    }
    for (; i < n4; i += kThreads) {
      const float4 v = g4[i];
      acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    acc = (acc + a1) + (a2 + a3);
    for (int j = n4 * 4 + threadIdx.x; j < c.count; j += kThreads) acc = fmaf(g[j], g[j], acc);
  }
  const float t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) grad_norm_finalize_kernel(const float* __restrict__ partial, int n,
                                                                  float max_norm, float* __restrict__ out) {
  __shared__ double s[1024];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) a += (double)partial[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(s[0]);
    const bool finite = isfinite(norm);
    out[0] = norm;
    out[1] = finite ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;  // torch.nn.utils.clip_grad_norm_
    out[2] = finite ? 1.0f : 0.0f;
  }
}

__global__ void __launch_bounds__(kThreads) grad_scale_kernel(const GradChunk* __restrict__ chunks,
                                                              const float* __restrict__ stats) {
  const float coef = stats[1];
  if (!(stats[2] != 0.f) || !(coef < 1.0f)) return;  // non-finite: the reference leaves the gradients alone
  const GradChunk c = chunks[blockIdx.x];
  if (c.is_bf16) {
    __nv_bfloat16* g = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(c.ptr));
    for (int i = threadIdx.x; i < c.count; i += kThreads) g[i] = __float2bfloat16(__bfloat162float(g[i]) * coef);
  } else {
    float* g = const_cast<float*>(static_cast<const float*>(c.ptr));
    const int n4 = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? c.count / 4 : 0;
    float4* g4 = reinterpret_cast<float4*>(g);
    int i = threadIdx.x;
    for (; i + 3 * kThreads < n4; i += 4 * kThreads) {
      float4 v0 = g4[i], v1 = g4[i + kThreads], v2 = g4[i + 2 * kThreads], v3 = g4[i + 3 * kThreads];
      g4[i] = make_float4(v0.x * coef, v0.y * coef, v0.z * coef, v0.w * coef);
      g4[i + kThreads] = make_float4(v1.x * coef, v1.y * coef, v1.z * coef, v1.w * coef);
      g4[i + 2 * kThreads] = make_float4(v2.x * coef, v2.y * coef, v2.z * coef, v2.w * coef);
      g4[i + 3 * kThreads] = make_float4(v3.x * coef, v3.y * coef, v3.z * coef, v3.w * coef);
    }
    for (; i < n4; i += kThreads) {
      float4 v = g4[i];
      g4[i] = make_float4(v.x * coef, v.y * coef, v.z * coef, v.w * coef);
    }
    for (int j = n4 * 4 + threadIdx.x; j < c.count; j += kThreads) g[j] *= coef;
  }
}

int64_t count_chunks(const int64_t* numels, int n) {
  int64_t c = 0;
  for (int i = 0; i < n; ++i) c += ceil_div(numels[i], kChunk);
  return c;
}

}  // namespace
}  // namespace pgica

extern "C" {

using namespace pgica;

int pgica_grad_norm_clip_workspace_bytes(const int64_t* numels_host, int n_tensors, size_t* bytes_host) {
  PGICA_REQUIRE(numels_host && bytes_host && n_tensors >= 1, "grad_norm_clip: bad argument");
  const int64_t chunks = count_chunks(numels_host, n_tensors);
  *bytes_host = align_up((size_t)chunks * sizeof(GradChunk), 256) + align_up((size_t)chunks * sizeof(float), 256);
  return PGICA_OK;
}

int pgica_grad_norm_clip(const void* const* grads_host, const int64_t* numels_host, const int32_t* is_bf16_host,
                         int n_tensors, float max_norm, int clip, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = pgica_device_check();
  if (rc != PGICA_OK) return rc;
  PGICA_REQUIRE(grads_host && numels_host && is_bf16_host && n_tensors >= 1 && stats && workspace,
                "grad_norm_clip: null pointer");
  PGICA_REQUIRE(max_norm > 0.f, "grad_norm_clip: max_norm must be positive");
  const int64_t chunks = count_chunks(numels_host, n_tensors);
  PGICA_REQUIRE(chunks >= 1 && chunks < (1ll << 31), "grad_norm_clip: nothing to do / too many elements");
  const size_t table_bytes = align_up((size_t)chunks * sizeof(GradChunk), 256);
  if (workspace_bytes < table_bytes + align_up((size_t)chunks * sizeof(float), 256)) {
    set_error("grad_norm_clip: workspace too small");
    return PGICA_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(clip & 2)) {
    // chunk table, built on the host and copied ahead of the launches (pageable source: staged synchronously).
    // Flag bit 1 of `clip` skips this when the caller kept the workspace of its previous call on the SAME tensors.
    GradChunk* table = static_cast<GradChunk*>(malloc((size_t)chunks * sizeof(GradChunk)));
    PGICA_REQUIRE(table, "grad_norm_clip: out of host memory");
    int64_t k = 0;
    for (int i = 0; i < n_tensors; ++i) {
      const size_t esz = is_bf16_host[i] ? 2 : 4;
      for (int64_t off = 0; off < numels_host[i]; off += kChunk) {
        table[k].ptr = static_cast<const uint8_t*>(grads_host[i]) + (size_t)off * esz;
        table[k].count = (int)(numels_host[i] - off < kChunk ? numels_host[i] - off : kChunk);
        table[k].is_bf16 = is_bf16_host[i] ? 1 : 0;
        ++k;
      }
    }
    cudaError_t e = cudaMemcpyAsync(workspace, table, (size_t)chunks * sizeof(GradChunk), cudaMemcpyHostToDevice, st);
    free(table);
    PGICA_CUDA_OK(e);
  }
  const GradChunk* d_table = static_cast<const GradChunk*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + table_bytes);
  grad_sumsq_kernel<<<(unsigned)chunks, kThreads, 0, st>>>(d_table, partial);
  PGICA_CUDA_OK(cudaGetLastError());
  grad_norm_finalize_kernel<<<1, 1024, 0, st>>>(partial, (int)chunks, max_norm, stats);
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(2);
  if (clip & 1) {
    grad_scale_kernel<<<(unsigned)chunks, kThreads, 0, st>>>(d_table, stats);
    PGICA_CUDA_OK(cudaGetLastError());
    count_launches(1);
  }
  return PGICA_OK;
}

}  // extern "C"
