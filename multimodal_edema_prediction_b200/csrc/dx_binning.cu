// Input binning of the MIMIC stays (duett/mimic_dataset.py:33-46, build_stay_tensor): scatter of the hourly event rows of
// a stay into the dense [T, 2V] grid  x[t, j] = (value - mean_j) / (std_j + 1e-7),  x[t, V + j] = count  for count > 0.
// The reference walks the rows of a pandas frame in python (iterrows x V per sample); here one launch bins a whole
// batch of stays: thread (stay, slot, variable) scans that stay's rows IN ORDER, so a later row of the same slot
// overwrites an earlier one exactly like the reference's sequential assignment.  Arithmetic is float64 -> one rounding to
// float32 (the reference computes in python floats and stores into a float32 tensor): results are bit-identical.
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

__global__ void __launch_bounds__(256) bin_events_kernel(const int* __restrict__ slot, const double* __restrict__ vals,
                                                        const double* __restrict__ cnts, const long long* __restrict__ row_start,
                                                        const double* __restrict__ means, const double* __restrict__ stds,
                                                        int B, int T, int V, float* __restrict__ x) {
  const long long n = (long long)B * T * V;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int j = (int)(i % V);
    const int t = (int)((i / V) % T);
    const int b = (int)(i / ((long long)V * T));
    float xv = 0.f, xc = 0.f;
    const double mu = means[j], sd = stds[j] + 1e-7;
    for (long long r = row_start[b]; r < row_start[b + 1]; ++r) {
      if (slot[r] != t) continue;
      const double c = cnts[r * V + j];
      if (c > 0.0) {   // NaN counts compare false, like `if count > 0` in the reference
        xv = (float)((vals[r * V + j] - mu) / sd);
        xc = (float)c;
      }
    }
    float* row = x + ((long long)b * T + t) * (2 * V);
    row[j] = xv;
    row[V + j] = xc;
  }
}

// SSL masking of Model.pretrain_prep_batch (duett/duett.py:189-237) driven by the host numpy-RNG draws: step[b, 0..K) = the
// masked timesteps (K = pretrain_masked_steps; drawn with replacement, so they may repeat), ev[b] = the masked variable
// (null: predict_events off), keep[b,v] = variable-dropout draw (null: pretrain_dropout == 0).  One thread per cell of xs
// [B,T,C] (C = 2V+1) writes x_c and, for the cells of the masked rows / column, the targets (y_ts / y_mask [B,K,V] in draw
// order, a repeated step fills each of its slots).  Pure selection: bit-exact against the reference's index chain.
__global__ void __launch_bounds__(256) ssl_mask_kernel(const float* __restrict__ xs, const int* __restrict__ step,
                                                       const int* __restrict__ ev, const unsigned char* __restrict__ keep, int B,
                                                       int T, int V, int K, float* __restrict__ xc, float* __restrict__ y_ts,
                                                       float* __restrict__ y_mask, float* __restrict__ y_ev,
                                                       float* __restrict__ y_ev_mask) {
  const int C = 2 * V + 1;
  const long long n = (long long)B * T * C;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int b = (int)(i / ((long long)C * T));
    const float x = xs[i];
    const int* st = step + (long long)b * K;
    const int e = ev ? ev[b] : -1;
    float o = x;
    for (int j = 0; j < K; ++j) {
      if (t != st[j]) continue;                      // x_c[b, st, :] = 0, flag column = 1; targets of the masked timestep
      o = c == C - 1 ? 1.f : 0.f;
      if (c < V) y_ts[((long long)b * K + j) * V + c] = x;
      else if (c < 2 * V) y_mask[((long long)b * K + j) * V + (c - V)] = fminf(fmaxf(x, 0.f), 1.f);
    }
    if (e >= 0) {                                    // x_c[b, :, e] = 0, x_c[b, :, e+V] = -1; targets of the masked variable
      if (c == e) { o = 0.f; y_ev[(long long)b * T + t] = x; }
      else if (c == e + V) { o = -1.f; y_ev_mask[(long long)b * T + t] = fminf(fmaxf(x, 0.f), 1.f); }
    }
    if (keep && c < 2 * V && o != -1.f) {            // variable dropout: observed-at-a-masked-step variables may be dropped
      const int v = c < V ? c : c - V;
      // K == 1: 1 - y_ts_masks;  K > 1: 1 - y_ts_masks.sum(dim=1).clip(0, 1)   (duett/duett.py:229-232)
      float m = 0.f;
      for (int j = 0; j < K; ++j) m += fminf(fmaxf(xs[((long long)b * T + st[j]) * C + V + v], 0.f), 1.f);
      if (K > 1) m = fminf(fmaxf(m, 0.f), 1.f);
      const bool kp = (1.f - m) != 0.f || keep[(long long)b * V + v] != 0;
      if (!kp) o = 0.f * o;                          // the reference multiplies by the boolean (keeps -0 / NaN semantics)
    }
    xc[i] = o;
  }
}

}  // namespace

extern "C" {

int dx_ssl_mask(const float* xs, const int* step, const int* ev, const unsigned char* keep, int B, int T, int V, int K,
                float* xc, float* y_ts, float* y_mask, float* y_ev, float* y_ev_mask, void* stream) {
  DX_CHECK_ARG(xs && step && xc && y_ts && y_mask && B > 0 && T > 0 && V > 0 && K > 0, "dx_ssl_mask: bad arguments");
  DX_CHECK_ARG(!ev || (y_ev && y_ev_mask), "dx_ssl_mask: event targets missing");
  const long long n = (long long)B * T * (2 * V + 1);
  long long gl = (n + 255) / 256;
  const int grid = (int)(gl < 148 * 16 ? gl : 148 * 16);
  ssl_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xs, step, ev, keep, B, T, V, K, xc, y_ts, y_mask, y_ev, y_ev_mask);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_bin_events(const int* slot, const double* vals, const double* cnts, const int64_t* row_start, const double* means,
                  const double* stds, int B, int T, int V, float* x, void* stream) {
  DX_CHECK_ARG(slot && vals && cnts && row_start && means && stds && x && B > 0 && T > 0 && V > 0, "dx_bin_events: bad arguments");
  const long long n = (long long)B * T * V;
  long long gl = (n + 255) / 256;
  const int grid = (int)(gl < 148 * 16 ? gl : 148 * 16);
  bin_events_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slot, vals, cnts, reinterpret_cast<const long long*>(row_start),
                                                            means, stds, B, T, V, x);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
