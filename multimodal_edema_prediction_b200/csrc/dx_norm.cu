// BatchNorm1d over the last dim of a [R,C] fp32 matrix (BatchNormLastDim, duett/duett.py:11-22: tab_encoder, cve time
// embedding, supervised head) and LayerNorm over the last dim (perceiver blocks,
// models/main_architecture_duett.py:745-774) — forward and backward, fused with their reductions.
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr float BN_EPS = 1e-5f;
constexpr float LN_EPS = 1e-5f;

// Two launches each way so that small-C problems still fill the machine: (1) partial sums over row chunks
// (grid = channel tiles x row chunks, atomics into a [C,2] workspace), (2) apply.  block = 32 channels x 8 row lanes.
__global__ void __launch_bounds__(256) bn2d_stats_kernel(const float* __restrict__ x, int R, int C, int rows_per_block,
                                                        double* __restrict__ ws) {
  __shared__ float sh[2][8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s = 0.f, ss = 0.f;
  if (c < C)
    for (int r = r0 + rl; r < r1; r += 8) {
      const float v = x[(long long)r * C + c];
      s += v;
      ss = fmaf(v, v, ss);
    }
  sh[0][rl][cl] = s;
  sh[1][rl][cl] = ss;
  __syncthreads();
  if (rl == 0 && c < C) {
    s = 0.f; ss = 0.f;
    for (int i = 0; i < 8; ++i) { s += sh[0][i][cl]; ss += sh[1][i][cl]; }
    atomicAdd(ws + 2 * c, (double)s);
    atomicAdd(ws + 2 * c + 1, (double)ss);
  }
}

__global__ void __launch_bounds__(256) bn2d_apply_kernel(const float* __restrict__ x, int R, int C, int rows_per_block,
                                                        const double* __restrict__ ws, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ run_mean,
                                                        float* __restrict__ run_var, float* __restrict__ y,
                                                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int training) {
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (c >= C) return;
  float m, rs;
  if (training) {
    const double mm = ws[2 * c] / R;
    double var = ws[2 * c + 1] / R - mm * mm;
    if (var < 0) var = 0;
    m = (float)mm;
    rs = (float)(1.0 / sqrt(var + (double)BN_EPS));
    if (blockIdx.y == 0 && rl == 0 && run_mean) {
      const double unb = R > 1 ? var * R / (R - 1) : var;
      run_mean[c] = 0.9f * run_mean[c] + 0.1f * m;
      run_var[c] = 0.9f * run_var[c] + 0.1f * (float)unb;
    }
  } else {
    m = run_mean[c];
    rs = rsqrtf(run_var[c] + BN_EPS);
  }
  if (blockIdx.y == 0 && rl == 0) { mean_out[c] = m; rstd_out[c] = rs; }
  const float ww = w[c], bb = b[c];
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int r = r0 + rl; r < r1; r += 8) y[(long long)r * C + c] = (x[(long long)r * C + c] - m) * rs * ww + bb;
}

// ws[c] = {sum dy, sum dy*xhat}
__global__ void __launch_bounds__(256) bn2d_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, int R, int C,
                                                             int rows_per_block, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ ws) {
  __shared__ float sh[2][8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float a = 0.f, bsum = 0.f;
  if (c < C) {
    const float m = mean[c], rs = rstd[c];
    for (int r = r0 + rl; r < r1; r += 8) {
      const float g = dy[(long long)r * C + c];
      a += g;
      bsum = fmaf(g, (x[(long long)r * C + c] - m) * rs, bsum);
    }
  }
  sh[0][rl][cl] = a;
  sh[1][rl][cl] = bsum;
  __syncthreads();
  if (rl == 0 && c < C) {
    a = 0.f; bsum = 0.f;
    for (int i = 0; i < 8; ++i) { a += sh[0][i][cl]; bsum += sh[1][i][cl]; }
    atomicAdd(ws + 2 * c, a);
    atomicAdd(ws + 2 * c + 1, bsum);
  }
}

__global__ void __launch_bounds__(256) bn2d_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, int R, int C,
                                                            int rows_per_block, const float* __restrict__ w,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ ws, float* __restrict__ dx,
                                                            float* __restrict__ dw, float* __restrict__ db, int training) {
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (c >= C) return;
  const float s1 = ws[2 * c], s2 = ws[2 * c + 1];
  if (blockIdx.y == 0 && rl == 0) { atomicAdd(db + c, s1); atomicAdd(dw + c, s2); }
  if (!dx) return;
  const float m = mean[c], rs = rstd[c], ww = w[c];
  const float mg = training ? s1 / R : 0.f, mgx = training ? s2 / R : 0.f;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int r = r0 + rl; r < r1; r += 8) {
    const float g = dy[(long long)r * C + c];
    const float xh = (x[(long long)r * C + c] - m) * rs;
    dx[(long long)r * C + c] = ww * rs * (g - mg - xh * mgx);
  }
}

static inline int bn_chunks(int R, int C, int& rpb) {
  const int ct = (C + 31) / 32;
  int nch = (2 * 148 + ct - 1) / ct;
  if (nch > (R + 63) / 64) nch = (R + 63) / 64;
  if (nch < 1) nch = 1;
  rpb = (R + nch - 1) / nch;
  rpb = ((rpb + 7) / 8) * 8;
  return (R + rpb - 1) / rpb;
}

// LayerNorm: one warp per row
template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, int R, int C, const float* __restrict__ w,
                                                    const float* __restrict__ b, T* __restrict__ y, float* __restrict__ mean,
                                                    float* __restrict__ rstd) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const T* xr = x + (long long)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += dx_ld(xr + c);
  const float m = dx_warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float t = dx_ld(xr + c) - m; v = fmaf(t, t, v); }
  const float rs = rsqrtf(dx_warp_sum(v) / C + LN_EPS);
  T* yr = y + (long long)row * C;
  for (int c = lane; c < C; c += 32) dx_st(yr + c, (dx_ld(xr + c) - m) * rs * w[c] + b[c]);
  if (lane == 0) { mean[row] = m; rstd[row] = rs; }
}

// dx per row (warp), dw/db accumulated per block in shared memory then atomics. C <= 1024.
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int R, int C,
                                                    const float* __restrict__ w, const float* __restrict__ mean,
                                                    const float* __restrict__ rstd, T* __restrict__ dx, float* __restrict__ dw,
                                                    float* __restrict__ db, int rows_per_block) {
  extern __shared__ float smem[];  // [2][C]
  float* sdw = smem;
  float* sdb = smem + C;
  for (int c = threadIdx.x; c < 2 * C; c += 256) smem[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int row = r0 + warp; row < r1; row += 8) {
    const T* xr = x + (long long)row * C;
    const T* gr = dy + (long long)row * C;
    const float m = mean[row], rs = rstd[row];
    float a = 0.f, b2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dx_ld(gr + c) * w[c];
      const float xh = (dx_ld(xr + c) - m) * rs;
      a += g;
      b2 = fmaf(g, xh, b2);
    }
    a = dx_warp_sum(a) / C;
    b2 = dx_warp_sum(b2) / C;
    T* dr = dx + (long long)row * C;
    for (int c = lane; c < C; c += 32) {
      const float g0 = dx_ld(gr + c);
      const float xh = (dx_ld(xr + c) - m) * rs;
      if (dx) dx_st(dr + c, rs * (g0 * w[c] - a - xh * b2));
      atomicAdd(&sdw[c], g0 * xh);
      atomicAdd(&sdb[c], g0);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(dw + c, sdw[c]);
    atomicAdd(db + c, sdb[c]);
  }
}

}  // namespace

extern "C" {

int dx_bn2d_fwd(const float* x, int R, int C, const float* w, const float* b, float* run_mean, float* run_var, float* y,
                float* mean, float* rstd, double* stats_ws, int training, void* stream) {
  DX_CHECK_ARG(x && w && b && y && mean && rstd && R > 0 && C > 0, "dx_bn2d_fwd: bad arguments");
  DX_CHECK_ARG(training || (run_mean && run_var), "dx_bn2d_fwd: eval mode needs running statistics");
  DX_CHECK_ARG(!training || stats_ws, "dx_bn2d_fwd: training mode needs the [C,2] double workspace");
  cudaStream_t st = (cudaStream_t)stream;
  int rpb;
  const int nch = bn_chunks(R, C, rpb);
  dim3 grid((C + 31) / 32, nch);
  if (training) {
    DX_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * C, st));
    bn2d_stats_kernel<<<grid, 256, 0, st>>>(x, R, C, rpb, stats_ws);
    DX_LAUNCH_CHECK();
  }
  bn2d_apply_kernel<<<grid, 256, 0, st>>>(x, R, C, rpb, stats_ws, w, b, run_mean, run_var, y, mean, rstd, training);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* dw, db are accumulated; dx may be NULL. ws: [C,2] f32 scratch. */
int dx_bn2d_bwd(const float* dy, const float* x, int R, int C, const float* w, const float* mean, const float* rstd,
                float* dx, float* dw, float* db, float* ws, int training, void* stream) {
  DX_CHECK_ARG(dy && x && w && mean && rstd && dw && db && ws, "dx_bn2d_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int rpb;
  const int nch = bn_chunks(R, C, rpb);
  dim3 grid((C + 31) / 32, nch);
  DX_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, st));
  bn2d_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dy, x, R, C, rpb, mean, rstd, ws);
  DX_LAUNCH_CHECK();
  bn2d_bwd_apply_kernel<<<grid, 256, 0, st>>>(dy, x, R, C, rpb, w, mean, rstd, ws, dx, dw, db, training);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_layernorm_fwd(const void* x, int R, int C, const float* w, const float* b, void* y, float* mean, float* rstd,
                     int dtype, void* stream) {
  DX_CHECK_ARG(x && w && b && y && mean && rstd, "dx_layernorm_fwd: bad arguments");
  const int grid = dx_ceil_div((long long)R * 32, 256);
  if (dtype == DX_BF16) ln_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, R, C, w, b, (bf16*)y, mean, rstd);
  else ln_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, R, C, w, b, (float*)y, mean, rstd);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* dw, db accumulated; dx may be NULL. */
int dx_layernorm_bwd(const void* dy, const void* x, int R, int C, const float* w, const float* mean, const float* rstd,
                     void* dx, float* dw, float* db, int dtype, void* stream) {
  DX_CHECK_ARG(dy && x && w && mean && rstd && dw && db && C <= 4096, "dx_layernorm_bwd: bad arguments");
  int nblk = 148 * 2;
  if (nblk > dx_ceil_div(R, 8)) nblk = dx_ceil_div(R, 8);
  const int rpb = dx_ceil_div(R, nblk);
  nblk = dx_ceil_div(R, rpb);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (dtype == DX_BF16)
    ln_bwd_kernel<bf16><<<nblk, 256, smem, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)x, R, C, w, mean, rstd, (bf16*)dx, dw, db, rpb);
  else
    ln_bwd_kernel<float><<<nblk, 256, smem, (cudaStream_t)stream>>>((const float*)dy, (const float*)x, R, C, w, mean, rstd, (float*)dx, dw, db, rpb);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
