#include "common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace pgica {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_error; }

static PFN_cuTensorMapEncodeTiled resolve_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride,
                   uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = resolve_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return PGICA_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (row_stride * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (base %p, pitch %llu B)", base,
              (unsigned long long)(row_stride * 2));
    return PGICA_ERR_INVALID_ARGUMENT;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estride[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu pitch %llu box_rows %u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride, box_rows);
    return PGICA_ERR_CUDA;
  }
  return PGICA_OK;
}

int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride,
                  uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled enc = resolve_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return PGICA_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (row_stride * 4) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (base %p, pitch %llu B)", base,
              (unsigned long long)(row_stride * 4));
    return PGICA_ERR_INVALID_ARGUMENT;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride * 4};
  cuuint32_t box[2] = {32u, box_rows};
  cuuint32_t estride[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows %llu cols %llu pitch %llu box_rows %u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride, box_rows);
    return PGICA_ERR_CUDA;
  }
  return PGICA_OK;
}

// ------------------------------------------------------------------------------------------------ options
namespace {
struct OptionTable {
  std::atomic<int64_t> v[kOptCount];
  OptionTable() {
    static const int64_t defaults[kOptCount] = {1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    for (int i = 0; i < kOptCount; ++i) v[i].store(defaults[i]);
    auto env_int = [](const char* name, Option o, OptionTable* t) {
      if (const char* e = getenv(name)) t->v[o].store(atoll(e));
    };
    // a profiler whose kernel replay cannot re-issue cooperative cluster launches is attached: plain launches by
    // default (PGICA_SGGF_COOP=1 keeps the cooperative launch, e.g. under `ncu --replay-mode application`)
    if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("CUDA_INJECTION64_PATH")) v[kOptSggfCoop].store(0);
    env_int("PGICA_SGG_FUSED", kOptSggFused, this);
    env_int("PGICA_SGGF_COOP", kOptSggfCoop, this);
    env_int("PGICA_SGGF_SPREAD", kOptSggfSpread, this);
    env_int("PGICA_SGGF_SLOTS", kOptSggfSlots, this);
    env_int("PGICA_SGGF_DEBUG_PRODUCERS_ONLY", kOptSggfProducersOnly, this);
    env_int("PGICA_SGG_CLUSTER", kOptSggCluster, this);
    env_int("PGICA_SGGF_SINGLE_CHUNK", kOptSggfSingleChunk, this);
    env_int("PGICA_SGGF_COL_GROUPS", kOptSggfColGroups, this);
    if (const char* e = getenv("PGICA_SGGF_PLAN")) {
      int r2 = 0, c2 = 0;
      if (sscanf(e, "%d,%d", &r2, &c2) == 2 && r2 >= 1 && c2 >= 1) {
        v[kOptSggfPlanR2].store(r2);
        v[kOptSggfPlanC2].store(c2);
      }
    }
  }
};
OptionTable& options() {
  static OptionTable t;
  return t;
}
const char* const kOptionNames[kOptCount] = {"sgg_fused",  "sggf_plan_r2", "sggf_plan_c2",        "sggf_coop", "sggf_spread",
                                             "sggf_slots", "sggf_producers_only", "sgg_cluster", "sggf_single_chunk",
                                             "sggf_col_groups"};
int option_index(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < kOptCount; ++i)
    if (strcmp(name, kOptionNames[i]) == 0) return i;
  return -1;
}
}  // namespace
int64_t get_option(Option o) { return options().v[o].load(std::memory_order_relaxed); }

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace pgica

extern "C" {

int pgica_abi_version(void) { return PGICA_ABI_VERSION; }
const char* pgica_last_error(void) { return pgica::get_error(); }
int pgica_sm_count(void) { return pgica::device_sm_count(); }
int64_t pgica_kernel_launches(void) { return (int64_t)pgica::launches_so_far(); }

int pgica_set_option(const char* name, int64_t value) {
  const int i = pgica::option_index(name);
  if (i < 0) {
    pgica::set_error("unknown option '%s'", name ? name : "(null)");
    return PGICA_ERR_INVALID_ARGUMENT;
  }
  pgica::options().v[i].store(value, std::memory_order_relaxed);
  return PGICA_OK;
}

int64_t pgica_get_option(const char* name) {
  const int i = pgica::option_index(name);
  return i < 0 ? INT64_MIN : pgica::options().v[i].load(std::memory_order_relaxed);
}

int pgica_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  PGICA_CUDA_OK(cudaGetDevice(&dev));
  PGICA_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  PGICA_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    pgica::set_error("device %d is sm_%d%d; this library only contains sm_100a code (no fallback)", dev, major, minor);
    return PGICA_ERR_UNSUPPORTED_DEVICE;
  }
  return PGICA_OK;
}

}  // extern "C"
