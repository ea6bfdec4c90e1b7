/*
 * pgica_debug.h — bring-up and test hooks of libpgica.so.  NOT part of the product ABI (include/pgica.h): nothing
 * on the loss-head path calls these; tests/ and tools/ do.
 */
#ifndef PGICA_DEBUG_H_
#define PGICA_DEBUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Debug / self-test: one CTA, one 128 x n x k tcgen05 product, D written to `d` (fp32 [128][n]).
 *   b_mn_major = 0: b is [n][k] (K-major);  1: b is [k][n] (MN-major, the layout the backward's second
 *   product reads W / H tiles in).  a_manual = 1 writes the A tile with st.shared through the
 *   sw128 swizzle formula instead of TMA (the way the backward stages its probability tile).
 * ---------------------------------------------------------------------------------------------- */
int pgica_probe_umma(const void* a, const void* b, int64_t n, int64_t k, int b_mn_major, int a_manual,
                     uint32_t b_lbo_bytes, uint32_t b_sbo_bytes, float* d, void* stream);

/* Host replay of the dual kernel's tile schedule (the same enumerators the device code runs; test hook, no device):
 * role 0: every quad in production order, 3 ints each (q, row pair, column pair); role 1 / 2: the pair-tiles X- /
 * Y-holder pair `idx` accumulates, 6 ints each (q, sel, row pair, column pair, first-of-period, period).  Returns the
 * number of records (only the first `capacity` are written), -1 on a bad argument.  Bits 8.. of `spread` carry the
 * column-group count of the X-holders (0 / 1 = none); with groups, X-holder idx = (row pair) * groups + group. */
int64_t pgica_debug_dual_schedule(int row_pairs, int col_pairs, int row_pairs_per_chunk, int col_pairs_per_pass,
                                  int spread, int role, int idx, int32_t* out_host, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* PGICA_DEBUG_H_ */
