/*
 * ttts_b200 -- C ABI of the Blackwell (sm_100a) hot-path library `libttts_b200.so`.
 *
 * The reference (adelacvg/ttts) is pure Python/PyTorch and has NO FFI of its own; the boundary a
 * maintainer binds is therefore the set of torch ops its nn.Modules dispatch on the hot path.
 * Each entry point below names the reference op site (file:line under /root/reference, or HF:
 * transformers 5.5 modeling_gpt2.py) whose arithmetic it replaces.  See INTEGRATION.md for the
 * ctypes stubs the reference side would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ttts_status otherwise; the text of the last
 *     error on the calling thread is available from ttts_last_error();
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch tensors); the library
 *     allocates no persistent device memory;
 *   - every call takes the cudaStream_t to launch on (as void*), performs no host sync;
 *   - bf16 = raw 16-bit brain float, row-major unless stated.
 */
#ifndef TTTS_B200_H
#define TTTS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TTTS_OK = 0,
    TTTS_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    TTTS_ERR_CUDA = -2,      /* CUDA runtime / driver error      */
    TTTS_ERR_ARCH = -3       /* device is not sm_100             */
} ttts_status;

int ttts_version(void);
const char* ttts_last_error(void);
/* 1 if the current device can run the sm_100a kernels */
int ttts_device_ok(void);

/* --------------------------------------------------------------------------------------------
 * GEMM on the 5th-gen tensor cores (tcgen05.mma, TMA-fed, TMEM accumulators).
 *   D[M,N] = A[M,K] * B[K,N]   (bf16 x bf16 -> fp32 accumulate) followed by a fused epilogue.
 * Replaces the cuBLAS addmm / linear calls of HF Conv1D (HF: pytorch_utils.py:119-123) and
 * nn.Linear heads (ttts/gpt/model.py:348-349,432-437) and their autograd backward.
 *   a_mn = 0: A stored [M][K] (K contiguous, row stride lda);  a_mn = 1: A stored [K][M] (M contiguous)
 *   b_mn = 0: B stored [N][K] (K contiguous, row stride ldb);  b_mn = 1: B stored [K][N] (N contiguous)
 * ------------------------------------------------------------------------------------------ */
typedef enum {
    TTTS_EPI_BF16 = 0,       /* out_bf16 = acc (+bias[n])                                              */
    TTTS_EPI_GELU = 1,       /* aux_out_bf16 = pre = bf16(acc+bias) ; out_bf16 = gelu_new(pre)         */
    TTTS_EPI_RESID = 2,      /* out_f32 = resid_f32 + dropout(bf16(acc+bias))   (out may alias resid)  */
    TTTS_EPI_DGELU = 3,      /* out_bf16 = acc * gelu_new'(aux_bf16[m,n])                              */
    TTTS_EPI_F32_ADD = 4,    /* out_f32 += acc   (red.global.add; used by split-K weight gradients)    */
    TTTS_EPI_F32 = 5         /* out_f32 = acc (+bias)                                                  */
} ttts_epilogue;

typedef struct {
    int32_t M, N, K;
    const void* A; int32_t lda; int32_t a_mn;
    const void* B; int32_t ldb; int32_t b_mn;
    int32_t epi;
    void* out; int32_t ldo;
    const float* bias;          /* [N] or NULL */
    const void* aux; int32_t ldaux;   /* RESID: fp32 residual in ; DGELU: bf16 pre-activation */
    void* aux_out; int32_t ldaux_out; /* GELU: bf16 pre-activation out (may be NULL) */
    int32_t split_k;            /* >=1; only with TTTS_EPI_F32_ADD */
    /* dropout for TTTS_EPI_RESID (p = drop_thresh16 / 65536; 0 disables) */
    uint32_t drop_thresh16; float drop_scale; uint64_t drop_seed;
} ttts_gemm_args;

int ttts_gemm_bf16(const ttts_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTTS_B200_H */
