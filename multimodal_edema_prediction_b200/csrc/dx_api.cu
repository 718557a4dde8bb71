// extern "C" entry points of libduett_b200.so that are not kernels themselves: error reporting, device probe,
// GEMM dispatch (tcgen05 for bf16 inputs, FFMA for fp32 inputs).
#include "dx_common.cuh"
#include "../../include/duett_b200.h"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void dx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int dx_gemm_simt_launch(const dx_gemm_desc* d, cudaStream_t stream);
int dx_gemm_tc_launch(const dx_gemm_desc* d, int bn, int stages, int a_lbo, int a_sbo, int b_lbo, int b_sbo,
                      cudaStream_t stream);

static int g_reserved_sms = -1;
int dx_gemm_reserved_sms() {
  if (g_reserved_sms < 0) {
    const char* env = getenv("DX_GEMM_SM_RESERVE");
    g_reserved_sms = env ? atoi(env) : 0;
    if (g_reserved_sms < 0) g_reserved_sms = 0;
  }
  return g_reserved_sms;
}

extern "C" {

/* Persistent tcgen05 GEMM grids use (#SMs - n) CTAs (n even) from now on: leaves SMs to a collective that runs concurrently
 * with the backward pass (ddp.GradReducer sets it for world > 1).  Returns the previous value. */
int dx_gemm_reserve_sms(int n) {
  const int prev = dx_gemm_reserved_sms();
  g_reserved_sms = n < 0 ? 0 : (n & ~1);
  return prev;
}

const char* dx_last_error(void) { return g_err; }
int dx_version(void) { return 100; }

int dx_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

static int check_gemm(const dx_gemm_desc* d) {
  DX_CHECK_ARG(d != nullptr, "dx_gemm: null descriptor");
  DX_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "dx_gemm: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  DX_CHECK_ARG(d->A && d->B, "dx_gemm: null operand");
  DX_CHECK_ARG(d->in_dtype == DX_F32 || d->in_dtype == DX_BF16, "dx_gemm: bad in_dtype %d", d->in_dtype);
  DX_CHECK_ARG(!d->accumulate || d->out_dtype == DX_F32, "dx_gemm: accumulate needs an f32 output");
  DX_CHECK_ARG(!d->cx || (d->coef_num && d->coef_den), "dx_gemm: cx needs coef_num/coef_den");
  DX_CHECK_ARG(d->act < DX_ACT_GELU_BWD || d->aux, "dx_gemm: backward activation needs aux");
  return DX_OK;
}

int dx_gemm(const dx_gemm_desc* d, void* stream) {
  int rc = check_gemm(d);
  if (rc) return rc;
  if (d->in_dtype == DX_BF16 && !d->force_simt) return dx_gemm_tc_launch(d, 0, 0, -1, -1, -1, -1, (cudaStream_t)stream);
  // fp32 operands: exact FFMA products by default; allow_tf32 opts into the tensor cores (kind::tf32: 10-bit mantissa
  // products, fp32 accumulate — the reference's own SSL / fine-tune precision, torch.set_float32_matmul_precision('high'),
  // duett/duett.py:9).  Layouts the TMA cannot address (row pitch not a multiple of 16 B, split_k) stay on FFMA.
  if (d->in_dtype == DX_F32 && d->allow_tf32 && !d->force_simt && d->split_k <= 1 && (d->lda % 4 == 0) && (d->ldb % 4 == 0) &&
      ((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0) && d->K >= 32 && (d->batch <= 1 || ((d->a_bs % 4 == 0) && (d->b_bs % 4 == 0))))
    return dx_gemm_tc_launch(d, 0, 0, -1, -1, -1, -1, (cudaStream_t)stream);
  return dx_gemm_simt_launch(d, (cudaStream_t)stream);
}

int dx_gemm_tc_debug(const dx_gemm_desc* d, int32_t block_n, int32_t stages, int32_t a_lbo, int32_t a_sbo,
                     int32_t b_lbo, int32_t b_sbo, void* stream) {
  int rc = check_gemm(d);
  if (rc) return rc;
  return dx_gemm_tc_launch(d, block_n, stages, a_lbo, a_sbo, b_lbo, b_sbo, (cudaStream_t)stream);
}

}  // extern "C"
