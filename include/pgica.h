/*
 * pgica.h — C ABI of the B200 (sm_100a) alignment-loss-head library.
 *
 * The reference (A-SHOJAEI/preference-guided-image-captioning-alignment) is pure Python/PyTorch and
 * has no FFI of its own; each entry point below therefore cites the reference *Python* function whose
 * arithmetic it replaces (paths relative to the reference repository root, "pkg/" =
 * src/preference_guided_image_captioning_alignment/).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative PGICA_ERR_* code otherwise; the message for the calling thread is
 *     available from pgica_last_error().  Nothing throws or exits across this boundary.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; the caller owns every buffer,
 *     including `workspace`; the library allocates nothing that outlives a call.
 *   - `stream` is a cudaStream_t (0 = legacy default stream); kernels are enqueued on it, the call does
 *     not synchronise.
 *   - matrices are row-major, contiguous, base pointers 16-byte aligned; bf16 = __nv_bfloat16.
 *   - there is no CPU fallback: on a device that is not sm_100 every compute entry point fails with
 *     PGICA_ERR_UNSUPPORTED_DEVICE.
 */
#ifndef PGICA_H_
#define PGICA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGICA_ABI_VERSION 1

#define PGICA_OK 0
#define PGICA_ERR_INVALID_ARGUMENT (-1)
#define PGICA_ERR_CUDA (-2)
#define PGICA_ERR_UNSUPPORTED_DEVICE (-3)
#define PGICA_ERR_WORKSPACE_TOO_SMALL (-4)

/* dtype tags for mask / float inputs */
#define PGICA_MASK_NONE 0 /* no mask: every position valid (components.py:341-344)  */
#define PGICA_MASK_I64 1  /* int64 {0,1}          (tokenizer attention_mask)        */
#define PGICA_MASK_F32 2  /* float32, MULTIPLIED into the log-prob like the reference */
#define PGICA_MASK_U8 3   /* bool / uint8                                            */
#define PGICA_MASK_I32 4

int pgica_abi_version(void);
const char* pgica_last_error(void);
/* 0 when the current CUDA device is sm_100 (B200); PGICA_ERR_UNSUPPORTED_DEVICE otherwise. */
int pgica_device_check(void);
int pgica_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1/K3 core — tensor-core GEMM with a fused online log-sum-exp + target-gather epilogue.
 *
 *   z[i][j] = scale * <a[i,:], b[j,:]>          (never written to memory)
 *   lse[i]  = log sum_j exp(z[i][j])            tgt[i] = z[i][label(i)]
 *
 * label(i) = labels[i] when `labels` != NULL (a negative or >= cols entry yields tgt[i] = 0), else
 * label(i) = i + diag_offset (NT-Xent positives on the diagonal).
 *
 * Replaces, for the LM head: transformers GPT2LMHeadModel.lm_head (modeling_gpt2.py:706, called from
 * pkg/models/model.py:604-610) + F.log_softmax + gather of pkg/models/components.py:346-352 and
 * pkg/models/model.py:1074-1079; for NT-Xent: torch.matmul(...)/temperature + the logsumexp inside
 * F.cross_entropy of pkg/models/model.py:988-995 and pkg/models/components.py:78-81,135-136.
 *
 * a: bf16 [rows][k], b: bf16 [cols][k], k % 8 == 0, scale > 0.  lse, tgt: fp32 [rows].
 * ---------------------------------------------------------------------------------------------- */
int pgica_gemm_lse_workspace_bytes(int64_t rows, int64_t cols, int64_t k, size_t* bytes_host);
int pgica_gemm_lse(const void* a, const void* b, int64_t rows, int64_t cols, int64_t k, float scale,
                   const int32_t* labels, int64_t diag_offset, float* lse, float* tgt, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Debug / self-test: one CTA, one 128 x n x k tcgen05 product, D written to `d` (fp32 [128][n]).
 *   b_mn_major = 0: b is [n][k] (K-major);  1: b is [k][n] (MN-major, the layout the backward's second
 *   product reads W / H tiles in).  a_manual = 1 writes the A tile with st.shared through the
 *   sw128 swizzle formula instead of TMA (the way the backward stages its probability tile).
 * ---------------------------------------------------------------------------------------------- */
int pgica_probe_umma(const void* a, const void* b, int64_t n, int64_t k, int b_mn_major, int a_manual,
                     uint32_t b_lbo_bytes, uint32_t b_sbo_bytes, float* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGICA_H_ */
