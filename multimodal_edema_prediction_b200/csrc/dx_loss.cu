// Fused loss kernels: every loss on the path is a small reduction over [B] or [B,K]; each kernel produces the scalar
// loss terms AND the gradient w.r.t. the logits in one launch (single block, warp-shuffle reduction).
//   dx_kd_loss            loss/losses_duett.py:8-25,39-57   (VanillaKLKD + StudentKDLoss)
//   dx_bce_logits         duett/duett.py:360-365            (supervised BCE with class-balance weights)
//   dx_masked_mse_bce     duett/duett.py:337-358            (SSL: masked MSE + 0.2 * presence BCE)
//   dx_masked_bce_cols    loss/losses_duett.py:152-194      (per-pathology masked BCE, DualPathologyLoss / PathologyMultiLabelLoss)
//   dx_aux_residual_kl    training_duett/engine.py:149-165  (label-smoothed Bernoulli KL on the residual correction)
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float bce_logits(float z, float y, float pw) {
  // -[pw*y*log(sigmoid z) + (1-y)*log(1-sigmoid z)] = pw*y*softplus(-z) + (1-y)*softplus(z)
  return pw * y * dx_softplus(-z) + (1.f - y) * dx_softplus(z);
}
__device__ __forceinline__ float bce_logits_grad(float z, float y, float pw) {
  const float s = dx_sigmoid(z);
  return -pw * y * (1.f - s) + (1.f - y) * s;
}

__global__ void __launch_bounds__(NT) kd_loss_kernel(const float* __restrict__ zs, const float* __restrict__ zt,
                                                    const float* __restrict__ y, int B, float T, float alpha, float pos_weight,
                                                    float eps, float* __restrict__ out /*[3]: total,bce,kd*/,
                                                    float* __restrict__ dz) {
  __shared__ float sh[33];
  float a_bce = 0.f, a_kd = 0.f;
  const float pw = pos_weight > 0.f ? pos_weight : 1.f;
  for (int i = threadIdx.x; i < B; i += NT) {
    const float z = zs[i], yy = y[i];
    a_bce += bce_logits(z, yy, pw);
    const float pt_raw = dx_sigmoid(zt[i] / T), ps_raw = dx_sigmoid(z / T);
    const float pt = fminf(fmaxf(pt_raw, eps), 1.f - eps);
    const float ps = fminf(fmaxf(ps_raw, eps), 1.f - eps);
    a_kd += pt * (logf(pt) - logf(ps)) + (1.f - pt) * (logf(1.f - pt) - logf(1.f - ps));
    if (dz) {
      // d kd_i / d ps = -pt/ps + (1-pt)/(1-ps); clamp passes gradient only strictly inside (eps, 1-eps)
      const bool inside = ps_raw > eps && ps_raw < 1.f - eps;
      const float dps = inside ? (-pt / ps + (1.f - pt) / (1.f - ps)) * ps_raw * (1.f - ps_raw) / T : 0.f;
      dz[i] = (alpha * bce_logits_grad(z, yy, pw) + (1.f - alpha) * T * T * dps) / B;
    }
  }
  a_bce = dx_block_sum(a_bce, sh);
  a_kd = dx_block_sum(a_kd, sh);
  if (threadIdx.x == 0) {
    const float bce = a_bce / B, kd = T * T * a_kd / B;
    out[0] = alpha * bce + (1.f - alpha) * kd;
    out[1] = bce;
    out[2] = kd;
  }
}

// loss = mean_i w_i * bce(z_i, y_i) with w_i = y_i > 0 ? w_pos : w_neg
__global__ void __launch_bounds__(NT) bce_logits_kernel(const float* __restrict__ z, const float* __restrict__ y, int n,
                                                       float w_pos, float w_neg, float* __restrict__ out,
                                                       float* __restrict__ dz) {
  __shared__ float sh[33];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) {
    const float w = y[i] > 0.f ? w_pos : w_neg;
    a += w * bce_logits(z[i], y[i], 1.f);
    if (dz) dz[i] = w * bce_logits_grad(z[i], y[i], 1.f) / n;
  }
  a = dx_block_sum(a, sh);
  if (threadIdx.x == 0) out[0] = a / n;
}

// out[0] += mean((yhat*m - y*m)^2) ; out[1] += w_presence * mean(bce(phat, m))     over n elements
__global__ void __launch_bounds__(NT) masked_mse_bce_kernel(const float* __restrict__ yhat, const float* __restrict__ phat,
                                                           const float* __restrict__ y, const float* __restrict__ m, int n,
                                                           float w_presence, float* __restrict__ out,
                                                           float* __restrict__ d_yhat, float* __restrict__ d_phat) {
  __shared__ float sh[33];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) {
    const float mm = m[i];
    const float diff = (yhat[i] - y[i]) * mm;
    a = fmaf(diff, diff, a);
    b += bce_logits(phat[i], mm, 1.f);
    if (d_yhat) d_yhat[i] = 2.f * diff * mm / n;
    if (d_phat) d_phat[i] = w_presence * bce_logits_grad(phat[i], mm, 1.f) / n;
  }
  a = dx_block_sum(a, sh);
  b = dx_block_sum(b, sh);
  if (threadIdx.x == 0) {
    out[0] += a / n;
    out[1] += w_presence * b / n;
  }
}

// per[k] = sum_b bce(z[b,k], y[b,k]; pw[k]) * m[b,k] / (sum_b m[b,k] + eps);  dz[b,k] = coef[k] * d per[k] / d z[b,k]
// one warp per column k (K is tiny)
__global__ void __launch_bounds__(NT) masked_bce_cols_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                                            const float* __restrict__ m, const float* __restrict__ pw,
                                                            const float* __restrict__ coef, int B, int K, float eps,
                                                            float* __restrict__ per, float* __restrict__ dz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += NT / 32) {
    const float p = pw ? pw[k] : 1.f;
    float a = 0.f, ms = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float mm = m[b * K + k];
      a += bce_logits(z[b * K + k], y[b * K + k], p) * mm;
      ms += mm;
    }
    a = dx_warp_sum(a);
    ms = dx_warp_sum(ms);
    const float inv = 1.f / (ms + eps);
    if (lane == 0) per[k] = a * inv;
    if (dz) {
      const float c = coef ? coef[k] : 1.f;
      for (int b = lane; b < B; b += 32)
        dz[b * K + k] = c * inv * m[b * K + k] * bce_logits_grad(z[b * K + k], y[b * K + k], p);
    }
  }
}

__global__ void __launch_bounds__(NT) aux_residual_kl_kernel(const float* __restrict__ img, const float* __restrict__ corr,
                                                            const float* __restrict__ y, const float* __restrict__ m, int n,
                                                            float eps, float* __restrict__ out, float* __restrict__ dcorr) {
  __shared__ float sh[33];
  float ms = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) ms += m[i];
  ms = dx_block_sum(ms, sh);
  const float denom = fmaxf(ms, 1.f);
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += NT) {
    const float ys = y[i] * (1.f - eps) + (1.f - y[i]) * eps;
    const float praw = dx_sigmoid(img[i] + corr[i]);
    const float p = fminf(fmaxf(praw, 1e-6f), 1.f - 1e-6f);
    const float kl = ys * (logf(ys) - logf(p)) + (1.f - ys) * (logf(1.f - ys) - logf(1.f - p));
    a += kl * m[i];
    if (dcorr) {
      const bool inside = praw > 1e-6f && praw < 1.f - 1e-6f;
      dcorr[i] = inside ? m[i] * (-ys / p + (1.f - ys) / (1.f - p)) * praw * (1.f - praw) / denom : 0.f;
    }
  }
  a = dx_block_sum(a, sh);
  if (threadIdx.x == 0) out[0] = a / denom;
}

}  // namespace

extern "C" {

int dx_kd_loss(const float* zs, const float* zt, const float* y, int B, float T, float alpha, float pos_weight, float eps,
               float* out3, float* dz, void* stream) {
  DX_CHECK_ARG(zs && zt && y && out3 && B > 0, "dx_kd_loss: bad arguments");
  kd_loss_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(zs, zt, y, B, T, alpha, pos_weight, eps, out3, dz);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_bce_logits(const float* z, const float* y, int n, float w_pos, float w_neg, float* out, float* dz, void* stream) {
  DX_CHECK_ARG(z && y && out && n > 0, "dx_bce_logits: bad arguments");
  bce_logits_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(z, y, n, w_pos, w_neg, out, dz);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* out[0] += masked MSE, out[1] += w_presence * presence BCE (caller zeroes out[0..1] first). */
int dx_masked_mse_bce(const float* yhat, const float* phat, const float* y, const float* m, int n, float w_presence,
                      float* out2, float* d_yhat, float* d_phat, void* stream) {
  DX_CHECK_ARG(yhat && phat && y && m && out2 && n > 0, "dx_masked_mse_bce: bad arguments");
  masked_mse_bce_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(yhat, phat, y, m, n, w_presence, out2, d_yhat, d_phat);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_masked_bce_cols(const float* z, const float* y, const float* m, const float* pos_weight, const float* coef, int B,
                       int K, float eps, float* per, float* dz, void* stream) {
  DX_CHECK_ARG(z && y && m && per && B > 0 && K > 0, "dx_masked_bce_cols: bad arguments");
  masked_bce_cols_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(z, y, m, pos_weight, coef, B, K, eps, per, dz);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_aux_residual_kl(const float* img_logits, const float* scaled_corr, const float* y, const float* mask, int n,
                       float eps_smooth, float* out, float* dcorr, void* stream) {
  DX_CHECK_ARG(img_logits && scaled_corr && y && mask && out && n > 0, "dx_aux_residual_kl: bad arguments");
  aux_residual_kl_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(img_logits, scaled_corr, y, mask, n, eps_smooth, out, dcorr);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
