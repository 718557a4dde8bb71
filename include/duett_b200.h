/*
 * duett_b200 — C ABI of the B200-native DuETT hot path (sm_100a).
 *
 * This header is the drop-in boundary.  The reference
 * (lastdancewithyou/multimodal_edema_prediction) has no FFI: its hot path is reached through Python
 * nn.Module calls that bottom out in stock ATen/cuBLAS kernels.  Each entry point below therefore cites
 * the reference Python site whose arithmetic it replaces; the Python host layer
 * (multimodal_edema_prediction_b200/) binds these symbols with ctypes and keeps the reference's
 * nn.Module / trainer surface.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (dx_last_error() gives the message);
 *  - all pointers are DEVICE pointers unless the name ends in _host; no allocation, no host sync,
 *    re-entrant per stream; `stream` is a cudaStream_t passed as void*;
 *  - dtype codes: DX_F32 = 0, DX_BF16 = 1.  "act dtype" is the storage type of activations
 *    (bf16 in the bf16 mode, f32 in the fp32 mode); statistics, parameters and gradients are f32.
 */
#ifndef DUETT_B200_H
#define DUETT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DX_F32 0
#define DX_BF16 1

/* activation codes for dx_gemm_desc.act */
#define DX_ACT_NONE 0
#define DX_ACT_GELU 1     /* out = gelu(pre), out2 = pre                     (x_transformers FeedForward, duett/duett.py:95-105) */
#define DX_ACT_RELU 2     /* out = relu(pre)                                  (simple_mlp, duett/duett.py:24-39) */
#define DX_ACT_TANH 3     /* out = tanh(pre)                                  (cve, duett/duett.py:151-157) */
#define DX_ACT_GELU_BWD 4 /* v = pre * gelu'(aux); out2 = v; out = v*row_scale2; row_dot += v*(aux-aux_bias) */
#define DX_ACT_RELU_BWD 5 /* v = pre * (aux > 0) */
#define DX_ACT_TANH_BWD 6 /* v = pre * (1 - aux^2), aux = tanh output */

const char* dx_last_error(void);
int dx_version(void);
/* 1 when the running device is sm_100 (tcgen05/TMEM/TMA paths usable) */
int dx_device_ok(void);

/*
 * Fused GEMM:  acc[m,n] = sum_k A(m,k) * B(n,k)     (i.e. A @ B^T, the nn.Linear contraction)
 *   a_mn = 0: A(m,k) = A[m*lda + k] (K-major)   a_mn = 1: A(m,k) = A[k*lda + m] (MN-major)
 *   b_mn = 0: B(n,k) = B[n*ldb + k]             b_mn = 1: B(n,k) = B[k*ldb + n]
 * in_dtype DX_BF16 -> tcgen05.mma kind::f16 with TMEM fp32 accumulators, TMA-fed (dx_gemm_tc.cu);
 * in_dtype DX_F32  -> fp32 FFMA kernel (dx_gemm_simt.cu), used by the fp32 precision mode.
 * Epilogue (all optional, null pointer = skipped), in this order:
 *   v = acc * row_scale[m] + bias[n];  act (see DX_ACT_*);  v += res[m,n];
 *   v -= cx[m,n] * coef_num[m] / max(coef_den[m], 1e-24);   (ScaleNorm backward projection)
 *   row_sumsq[m] += sum_n v^2;  out[m,n] (=|+=) v
 * Replaces: every nn.Linear / x_transformers projection on the path — duett/duett.py:95-105 (Encoder
 * to_q/to_k/to_v/to_out/ff), :84-86,106,110-125 (MLPs), models/main_architecture_duett.py:566,1027,1216-1219.
 */
typedef struct dx_gemm_desc {
  int32_t M, N, K;
  int32_t in_dtype;   /* dtype of A and B */
  int32_t a_mn, b_mn;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* out; int64_t ldo; int32_t out_dtype; int32_t accumulate; /* accumulate: out += v (f32 out only) */
  void* out2; int64_t ldo2;          /* act dtype */
  int32_t act; int32_t act_dtype;    /* act dtype = dtype of out2/res/cx/aux */
  const float* row_scale;            /* [M] */
  const float* row_scale2;           /* [M] (DX_ACT_GELU_BWD only) */
  const float* bias;                 /* [N] */
  const void* res; int64_t ldr;      /* [M,N] residual */
  const void* aux; int64_t ldx;      /* [M,N] saved pre-activation / activation for *_BWD */
  const float* aux_bias;             /* [N] subtracted from aux in row_dot */
  const void* cx; int64_t ldc;       /* [M,N] */
  const float* coef_num; const float* coef_den; /* [M] */
  float* row_sumsq;                  /* [M] atomically accumulated */
  float* row_dot;                    /* [M] atomically accumulated (DX_ACT_GELU_BWD) */
  int32_t force_simt;                /* test hook: run the FFMA kernel even for bf16 inputs */
  int32_t reserved;
} dx_gemm_desc;

int dx_gemm(const dx_gemm_desc* d, void* stream);

/*
 * Test hook for bring-up: same as dx_gemm on the tcgen05 path but with the shared-memory
 * matrix-descriptor fields overridden (lbo/sbo in bytes for A and B; <0 keeps the built-in value).
 */
int dx_gemm_tc_debug(const dx_gemm_desc* d, int32_t block_n, int32_t stages, int32_t a_lbo, int32_t a_sbo,
                     int32_t b_lbo, int32_t b_sbo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUETT_B200_H */
