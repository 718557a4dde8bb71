// AUROC / AUPRC of a binary scorer on the device (training_duett/evaluator.py:10-37: sigmoid -> sklearn roc_auc_score /
// average_precision_score on the host).  Same definitions, evaluated from the score-sorted label sequence:
//   thresholds = distinct scores, descending; TP_g / FP_g = cumulative positives / negatives down to threshold g
//   AUROC = sum_g (FP_g - FP_{g-1}) (TP_g + TP_{g-1}) / (2 P N)          (trapezoid over the ROC points = sklearn's auc)
//   AUPRC = sum_g (TP_g - TP_{g-1}) / P * TP_g / (TP_g + FP_g)           (sklearn's step-wise average precision)
// Counts are exact integers; the two sums are accumulated in double in a fixed order (deterministic).
// One CTA: bitonic sort of (score, label) in global scratch, then a chunked scan.  Evaluation sets are 1e3..1e6 samples
// and this runs once per epoch, so a single CTA (no inter-CTA synchronisation) is the simple, sufficient shape.
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr int NT = 1024;

__global__ void __launch_bounds__(NT) binary_auc_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                       long long n, long long npad, float* __restrict__ key,
                                                       float* __restrict__ lab, int apply_sigmoid, double* __restrict__ out) {
  __shared__ long long s_pos[NT];      // positives per chunk -> exclusive prefix
  __shared__ long long s_gtp[NT];      // (tp, fp) at the last threshold inside the chunk, -1 = none
  __shared__ long long s_gfp[NT];
  __shared__ double s_a[NT], s_b[NT];
  const int tid = threadIdx.x;
  for (long long i = tid; i < npad; i += NT) {
    float k = -INFINITY, l = 0.f;
    if (i < n) {
      const float x = logits[i];
      k = apply_sigmoid ? 1.f / (1.f + expf(-x)) : x;
      l = labels[i] > 0.5f ? 1.f : 0.f;
    }
    key[i] = k;
    lab[i] = l;
  }
  __syncthreads();
  // bitonic sort, descending by key (padding = -inf sinks to the end; NaN scores are not supported)
  for (long long k2 = 2; k2 <= npad; k2 <<= 1) {
    for (long long j = k2 >> 1; j > 0; j >>= 1) {
      for (long long i = tid; i < npad; i += NT) {
        const long long ixj = i ^ j;
        if (ixj > i) {
          const float a = key[i], b = key[ixj];
          const bool desc = (i & k2) == 0;
          if (desc ? (a < b) : (a > b)) {
            key[i] = b; key[ixj] = a;
            const float la = lab[i];
            lab[i] = lab[ixj]; lab[ixj] = la;
          }
        }
      }
      __syncthreads();
    }
  }
  // chunked scan over the n sorted samples
  const long long chunk = (n + NT - 1) / NT;
  const long long i0 = (long long)tid * chunk, i1 = i0 + chunk < n ? i0 + chunk : n;
  long long cpos = 0;
  for (long long i = i0; i < i1; ++i) cpos += lab[i] > 0.5f;
  s_pos[tid] = cpos;
  __syncthreads();
  if (tid == 0) {   // exclusive prefix (1024 adds)
    long long run = 0;
    for (int t = 0; t < NT; ++t) { const long long c = s_pos[t]; s_pos[t] = run; run += c; }
    out[2] = (double)run;        // P
    out[3] = (double)n;
  }
  __syncthreads();
  const long long P = (long long)out[2], Nn = n - P;
  // pass 1: (tp, fp) at the last threshold (group end) inside each chunk
  {
    long long tp = s_pos[tid], fp = i0 - s_pos[tid], gtp = -1, gfp = -1;
    for (long long i = i0; i < i1; ++i) {
      if (lab[i] > 0.5f) ++tp; else ++fp;
      if (i == n - 1 || key[i] != key[i + 1]) { gtp = tp; gfp = fp; }
    }
    s_gtp[tid] = gtp; s_gfp[tid] = gfp;
  }
  __syncthreads();
  // pass 2: contributions of the thresholds inside the chunk
  double a = 0.0, b = 0.0;
  if (i0 < i1) {
    long long ptp = 0, pfp = 0;   // previous threshold: nearest earlier chunk that holds one
    for (int t = tid - 1; t >= 0; --t)
      if (s_gtp[t] >= 0) { ptp = s_gtp[t]; pfp = s_gfp[t]; break; }
    long long tp = s_pos[tid], fp = i0 - s_pos[tid];
    for (long long i = i0; i < i1; ++i) {
      if (lab[i] > 0.5f) ++tp; else ++fp;
      if (i == n - 1 || key[i] != key[i + 1]) {
        a += (double)(fp - pfp) * (double)(tp + ptp);
        if (tp > ptp) b += (double)(tp - ptp) * ((double)tp / (double)(tp + fp));
        ptp = tp; pfp = fp;
      }
    }
  }
  s_a[tid] = a; s_b[tid] = b;
  __syncthreads();
  if (tid == 0) {
    double sa = 0.0, sb = 0.0;
    for (int t = 0; t < NT; ++t) { sa += s_a[t]; sb += s_b[t]; }
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (P > 0 && Nn > 0) ? sa / (2.0 * (double)P * (double)Nn) : nanv;   // roc_auc_score raises on a single class
    out[1] = P > 0 ? sb / (double)P : nanv;
  }
}

}  // namespace

extern "C" {

/* out[0] = AUROC, out[1] = AUPRC (average precision), out[2] = #positives, out[3] = n; key_ws / lab_ws: npad floats each,
 * npad = n rounded up to a power of two.  apply_sigmoid = 1 ranks sigmoid(logits) in fp32 like the reference does. */
int dx_binary_auc(const float* logits, const float* labels, int64_t n, float* key_ws, float* lab_ws, int64_t npad,
                  int apply_sigmoid, double* out, void* stream) {
  DX_CHECK_ARG(logits && labels && key_ws && lab_ws && out && n > 0, "dx_binary_auc: bad arguments");
  DX_CHECK_ARG(npad >= n && (npad & (npad - 1)) == 0, "dx_binary_auc: npad must be a power of two >= n");
  binary_auc_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(logits, labels, n, npad, key_ws, lab_ws, apply_sigmoid, out);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
