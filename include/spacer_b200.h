/* spacer_b200 -- C ABI of the B200-native SG-RLVR hot path (libspacer_b200.so).
 *
 * The reference (OuyangKun10/SpaceR) has no FFI of its own: its hot path is Python
 * (`SGRLVRTrainer.compute_loss`, SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py:384-686)
 * that reaches the GPU through transformers' `Qwen2VLForConditionalGeneration`
 * (transformers 5.5.0, models/qwen2_vl/modeling_qwen2_vl.py -- "MQ2" below), `generate()`
 * (generation/utils.py:2658-2830) and ATen.  This header is the boundary a maintainer binds instead:
 * every entry point names the reference call it replaces.  The ctypes binding that ships with this
 * repo is spacer_b200/_lib.py; INTEGRATION.md shows the stub on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - bf16 tensors are `void*` to 2-byte elements, row-major, 16-byte aligned;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises with the host unless the
 *     comment says so;
 *   - return 0 on success; non-zero on error with the message available from sb_last_error()
 *     (thread-local).  There is no CPU fallback anywhere in this library.
 */
#ifndef SPACER_B200_H
#define SPACER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sb_stream_t;

#define SB_ABI_VERSION 1

const char* sb_last_error(void);
int sb_abi_version(void);
/* fills SM count and compute capability of the current device; error if it is not sm_100 */
int sb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * GEMM  D[M,N] = epi(A[M,K] * B[N,K]^T)   (tcgen05 + TMEM + TMA)
 * replaces: torch.nn.functional.linear / Conv3d-as-GEMM under MQ2:304-310 (patch embed), :401-405,
 * :329-337 (ViT), :559-593 (attention projections), :502-504 (SwiGLU MLP), :1437-1438 (lm_head),
 * and autograd's backward GEMMs for the same.
 * ------------------------------------------------------------------------------------------------ */
enum {
  SB_EPI_STORE = 0,     /* D = bf16(acc + bias) [+ residual]                                        */
  SB_EPI_QUICKGELU = 1, /* z = acc + bias; aux = z (optional); D = z * sigmoid(1.702 z)   (ViT fc1) */
  SB_EPI_GELU = 2,      /* z = acc + bias; aux = z (optional); D = gelu_erf(z)       (PatchMerger)  */
  SB_EPI_SWIGLU = 3,    /* B rows interleaved [64 gate | 64 up]; D[M,N/2] = silu(g)*u; aux = raw    */
  SB_EPI_F32T = 4,      /* D_f32[split][n][m] = partial acc (swap-AB decode GEMV with split-K)      */
  SB_EPI_LMHEAD = 5,    /* per-row (max, sumexp) per N tile of bf16-rounded logits + target gather  */
  SB_EPI_DLOGITS = 6    /* D = bf16(coef[m] * (onehot(target[m]) - exp(logit - lse[m])))            */
};

typedef struct sb_gemm_args {
  int M, N, K;
  const void* A; long long lda; int a_mn; /* a_mn=0: A is [M,K] row-major; 1: A stored as [K,M]     */
  const void* B; long long ldb; int b_mn; /* b_mn=0: B is [N,K] row-major; 1: B stored as [K,N]     */
  void* D; long long ldd;
  int epilogue;
  int k_splits;                           /* >1 only with SB_EPI_F32T                               */
  int bn;                                 /* force tile N (0 = auto)                                */
  const void* bias;                       /* bf16 [N] or NULL                                       */
  const void* residual; long long ldr;    /* bf16 [M,N] or NULL (may alias D for accumulation)      */
  void* aux; long long ldaux;             /* bf16 pre-activation / raw gate-up output or NULL       */
  const int* targets;                     /* [M] LMHEAD / DLOGITS                                   */
  float* lse_part;                        /* [M][ceil(N/256)][2] LMHEAD out: (max, sumexp)          */
  float* tgt_logit;                       /* [M] LMHEAD out                                         */
  const float* lse;                       /* [M] DLOGITS in                                         */
  const float* coef;                      /* [M] DLOGITS in: dLoss/dlogprob                         */
} sb_gemm_args;

int sb_gemm(const sb_gemm_args* args, sb_stream_t stream);
/* number of non-empty K splits sb_gemm will use for (K, k_splits) */
int sb_gemm_effective_splits(int K, int k_splits);

#ifdef __cplusplus
}
#endif
#endif /* SPACER_B200_H */
