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

}  // namespace

extern "C" {

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
