// Host-side helpers shared by the C-ABI translation units: thread-local error string,
// CUDA error mapping, TMA tensor-map construction (driver entry point resolved at run time,
// so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/pgica.h"
#include "../../include/pgica_debug.h"

namespace pgica {

void set_error(const char* fmt, ...);
const char* get_error();

#define PGICA_CUDA_OK(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::pgica::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PGICA_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

#define PGICA_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::pgica::set_error(__VA_ARGS__);    \
      return PGICA_ERR_INVALID_ARGUMENT;  \
    }                                     \
  } while (0)

// Row-major [rows][cols] bf16 matrix with `row_stride` elements between rows; box = box_rows x 64
// columns, SWIZZLE_128B, out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride,
                   uint32_t box_rows);

// Row-major [rows][cols] fp32 matrix; box = box_rows x 32 columns (128 B), SWIZZLE_128B (TMA stores / reductions).
int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride,
                  uint32_t box_rows);

int device_sm_count();

// process-wide count of kernels this library has launched (bench.py reports it as `gpu_launches`)
void count_launches(int n);

// Process-wide tuning options (pgica_set_option / pgica_get_option).  Defaults are seeded ONCE, when the library is
// loaded, from environment variables PGICA_<NAME>; nothing reads the environment on the call path.
enum Option {
  kOptSggFused = 0,          // "sgg_fused": 1 = dual backward kernel when both gradients are wanted (default), 0 = one launch per product
  kOptSggfPlanR2,            // "sggf_plan_r2", "sggf_plan_c2": pin the dual kernel's role split (row pairs per chunk,
  kOptSggfPlanC2,            //   column pairs per pass); 0 = planner decides
  kOptSggfCoop,              // "sggf_coop": 1 = cooperative launch, refused launch is an error (default); 0 = plain launch
  kOptSggfSpread,            // "sggf_spread": alternative diagonal schedule (tuning)
  kOptSggfSlots,             // "sggf_slots": exchange double-slots per producer CTA (0 = default 4)
  kOptSggfProducersOnly,     // "sggf_producers_only": diagnostics, holders idle
  kOptSggCluster,            // "sgg_cluster": cluster size of the one-product kernel (4, 2 or 1; 0 = largest that tiles k)
  kOptSggfSingleChunk,       // "sggf_single_chunk": 1 = keep all of X in one chunk whenever it fits (W streamed once)
  kOptSggfColGroups,         // "sggf_col_groups": 1 = the planner may split an X-holder's column sweep over several pairs (default),
                             //   0 = never, n > 1 = exactly n groups when they fit
  kOptCount
};
int64_t get_option(Option o);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pgica
