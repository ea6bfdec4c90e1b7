// Valid-row compaction for the Stage-2 head.  Rows whose mask weight is zero (and the never-scored last position of
// every sequence) contribute nothing to the loss or to any gradient; with the reference's
// tokenisation (padding="max_length", 128 positions, 10-20 real tokens: pkg/data/preprocessing.py:223-231) that is
// > 80 % of the rows of a real Stage-2 batch.  These kernels build the list of scored rows, gather them (casting to the
// bf16 tensor-core operand on the way) into a dense [n][d] matrix, and scatter per-row results back.  All HBM-bound.
#include "common.h"
#include "ptx.cuh"

namespace pgica {
namespace {

// One block: index[0..count) = the rows r (ascending) with row_weight[r] != 0 (NaN weights count as scored: they must
// poison the sequence like in the reference), count[0] = how many.  Ballot + popc scan, 1024 rows per iteration.
__global__ void compact_rows_kernel(const float* __restrict__ row_weight, int rows, int* __restrict__ index,
                                    int* __restrict__ count) {
  __shared__ int warp_total[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < rows; start += 1024) {
    const int r = start + threadIdx.x;
    const bool valid = r < rows && row_weight[r] != 0.f;
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_total[warp] = __popc(ballot);
    __syncthreads();
    int offset = base_s;
    for (int w = 0; w < warp; ++w) offset += warp_total[w];
    if (valid) index[offset + before] = r;
    __syncthreads();
    if (threadIdx.x == 1023) base_s = offset + before + (valid ? 1 : 0);
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = base_s;
}

// dst[i][:] = bf16(src[index[i]][:]); optionally label_out[i] = label_in[index[i]].  One block per output row.
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, const int* __restrict__ index, int d,
                                   __nv_bfloat16* __restrict__ dst, const int* __restrict__ label_in,
                                   int* __restrict__ label_out) {
  const int i = blockIdx.x;
  const int r = index[i];
  if (threadIdx.x == 0 && label_in) label_out[i] = label_in[r];
  const T* s = src + (size_t)r * d;
  __nv_bfloat16* o = dst + (size_t)i * d;
  if constexpr (sizeof(T) == 4) {
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(s + c);
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 p;
      p.x = *reinterpret_cast<unsigned*>(&lo);
      p.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(o + c) = p;
    }
  } else {
    for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8)
      *reinterpret_cast<uint4*>(o + c) = *reinterpret_cast<const uint4*>(s + c);
  }
}

// dst[index[i]][:] = src[i][:] with a dtype change if asked (dst was zero-filled by the caller of the kernel)
template <typename S, typename D>
__global__ void scatter_rows_kernel(const S* __restrict__ src, const int* __restrict__ index, int d,
                                    D* __restrict__ dst) {
  const int i = blockIdx.x;
  const S* s = src + (size_t)i * d;
  D* o = dst + (size_t)index[i] * d;
  for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
    float v[4];
    if constexpr (sizeof(S) == 4) {
      const float4 q = *reinterpret_cast<const float4*>(s + c);
      v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
      const uint2 q = *reinterpret_cast<const uint2*>(s + c);
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
      v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
    }
    if constexpr (sizeof(D) == 4) {
      *reinterpret_cast<float4*>(o + c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 p;
      p.x = *reinterpret_cast<unsigned*>(&lo);
      p.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(o + c) = p;
    }
  }
}

// 4-byte elements (float or int32 bit patterns): dst[i] = src[index[i]]  /  dst[index[i]] = src[i]
__global__ void gather_u32_kernel(const unsigned* __restrict__ src, const int* __restrict__ index, int n,
                                  unsigned* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[index[i]];
}
__global__ void scatter_u32_kernel(const unsigned* __restrict__ src, const int* __restrict__ index, int n,
                                   unsigned* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[index[i]] = src[i];
}

}  // namespace
}  // namespace pgica

extern "C" {

using namespace pgica;

int pgica_compact_rows(const float* row_weight, int64_t rows, int32_t* index, int32_t* count, void* stream) {
  PGICA_REQUIRE(row_weight && index && count, "compact_rows: null pointer");
  PGICA_REQUIRE(rows > 0 && rows < (1ll << 31), "compact_rows: bad row count");
  compact_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_weight, (int)rows, index, count);
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}

int pgica_gather_rows_bf16(const void* src, int src_is_bf16, const int32_t* index, int64_t n, int64_t d,
                           void* dst_bf16, const int32_t* label_in, int32_t* label_out, void* stream) {
  PGICA_REQUIRE(src && index && dst_bf16, "gather_rows: null pointer");
  PGICA_REQUIRE((label_in == nullptr) == (label_out == nullptr), "gather_rows: labels come as an in/out pair");
  PGICA_REQUIRE(n >= 0 && n < (1ll << 31) && d > 0 && d % 8 == 0, "gather_rows: need d %% 8 == 0 (got n %lld, d %lld)",
                (long long)n, (long long)d);
  PGICA_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_bf16)) & 15u) == 0,
                "gather_rows: pointers must be 16-byte aligned");
  if (n == 0) return PGICA_OK;
  const int threads = d >= 1024 ? 256 : 128;
  if (src_is_bf16)
    gather_rows_kernel<__nv_bfloat16><<<(unsigned)n, threads, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(src), index, (int)d, static_cast<__nv_bfloat16*>(dst_bf16), label_in,
        label_out);
  else
    gather_rows_kernel<float><<<(unsigned)n, threads, 0, (cudaStream_t)stream>>>(
        static_cast<const float*>(src), index, (int)d, static_cast<__nv_bfloat16*>(dst_bf16), label_in, label_out);
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}

int pgica_scatter_rows(const void* src, int src_is_bf16, const int32_t* index, int64_t n, int64_t d, void* dst,
                       int dst_is_bf16, int64_t dst_rows, void* stream) {
  PGICA_REQUIRE(src && index && dst, "scatter_rows: null pointer");
  PGICA_REQUIRE(n >= 0 && n <= dst_rows && dst_rows < (1ll << 31) && d > 0 && d % 4 == 0,
                "scatter_rows: bad shape (n %lld, rows %lld, d %lld)", (long long)n, (long long)dst_rows, (long long)d);
  PGICA_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0,
                "scatter_rows: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PGICA_CUDA_OK(cudaMemsetAsync(dst, 0, (size_t)dst_rows * d * (dst_is_bf16 ? 2 : 4), st));
  if (n == 0) return PGICA_OK;
  const unsigned g = (unsigned)n;
  const int dd = (int)d;
  if (src_is_bf16 && dst_is_bf16)
    scatter_rows_kernel<<<g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), index, dd,
                                           static_cast<__nv_bfloat16*>(dst));
  else if (src_is_bf16)
    scatter_rows_kernel<<<g, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), index, dd, static_cast<float*>(dst));
  else if (dst_is_bf16)
    scatter_rows_kernel<<<g, 256, 0, st>>>(static_cast<const float*>(src), index, dd, static_cast<__nv_bfloat16*>(dst));
  else
    scatter_rows_kernel<<<g, 256, 0, st>>>(static_cast<const float*>(src), index, dd, static_cast<float*>(dst));
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}

int pgica_gather_u32(const void* src, const int32_t* index, int64_t n, void* dst, void* stream) {
  PGICA_REQUIRE(src && index && dst && n >= 0 && n < (1ll << 31), "gather_u32: bad argument");
  if (n == 0) return PGICA_OK;
  gather_u32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const unsigned*>(src), index, (int)n, static_cast<unsigned*>(dst));
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}

int pgica_scatter_u32(const void* src, const int32_t* index, int64_t n, void* dst, int64_t dst_n, void* stream) {
  PGICA_REQUIRE(src && index && dst && n >= 0 && n <= dst_n && dst_n < (1ll << 31), "scatter_u32: bad argument");
  PGICA_CUDA_OK(cudaMemsetAsync(dst, 0, (size_t)dst_n * 4, (cudaStream_t)stream));
  if (n == 0) return PGICA_OK;
  scatter_u32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const unsigned*>(src), index, (int)n, static_cast<unsigned*>(dst));
  PGICA_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return PGICA_OK;
}

}  // extern "C"
