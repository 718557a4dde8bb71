// Small data-movement kernels around the heads, and the fused optimizer.
//   dx_mean_rows / _bwd : StudentModel mean pooling over the T hourly tokens (models/main_architecture_duett.py:1228-1231),
//                         Model 'averaging' fusion (duett/duett.py:297-298)
//   dx_gather_vec / dx_scatter_vec : masked-step row gather and masked-variable column gather of the SSL heads
//                         (duett/duett.py:291-296, 310-313) and their backward scatter
//   dx_adamw, dx_sumsq  : fused AdamW over flat parameter/gradient buffers + global-norm clip factor
//                         (training_duett/trainer.py:383,902; duett/duett.py:325-327; train_duett_ssl.py:191)
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr int NT = 256;

// y[b, e] = (1/T) sum_{t<T} x[b, t, e]      x: [B, T1, E] (T1 >= T rows per sample)
template <typename TI>
__global__ void __launch_bounds__(NT) mean_rows_kernel(const TI* __restrict__ x, float* __restrict__ y, int B, int T1, int T,
                                                      long long E) {
  const long long nvec = (long long)B * (E >> 3);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
    const int b = (int)(i / (E >> 3));
    const long long e = (i % (E >> 3)) << 3;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < T; ++t) {
      float v[8];
      dx_ld8(x + ((long long)b * T1 + t) * E + e, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    const float inv = 1.f / T;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    dx_st8(y + (long long)b * E + e, acc);
  }
}

// dx[b, t, e] = dy[b, e] / T for t < T, 0 for T <= t < T1
template <typename TO>
__global__ void __launch_bounds__(NT) mean_rows_bwd_kernel(const float* __restrict__ dy, TO* __restrict__ dx, int B, int T1, int T,
                                                          long long E) {
  const long long nvec = (long long)B * T1 * (E >> 3);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
    const long long e = (i % (E >> 3)) << 3;
    const long long row = i / (E >> 3);
    const int b = (int)(row / T1), t = (int)(row % T1);
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (t < T) {
      dx_ld8(dy + (long long)b * E + e, v);
      const float inv = 1.f / T;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= inv;
    }
    dx_st8(dx + row * E + e, v);
  }
}

// out[i, 0:L] = src[off[i] : off[i]+L]   (TI -> f32); off[i] < 0: a zero row
template <typename TI>
__global__ void __launch_bounds__(NT) gather_vec_kernel(const TI* __restrict__ src, const long long* __restrict__ off,
                                                       float* __restrict__ out, int n, int L) {
  const long long tot = (long long)n * L;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < tot; i += (long long)gridDim.x * NT) {
    const int r = (int)(i / L), j = (int)(i % L);
    const long long o = off[r];
    out[i] = o < 0 ? 0.f : dx_ld(src + o + j);
  }
}
// dst[off[i] : off[i]+L] (+)= src[i, 0:L]   (f32 -> TO); offsets must not overlap; off[i] < 0: row skipped
template <typename TO>
__global__ void __launch_bounds__(NT) scatter_vec_kernel(const float* __restrict__ src, const long long* __restrict__ off,
                                                        TO* __restrict__ dst, int n, int L, int accumulate) {
  const long long tot = (long long)n * L;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < tot; i += (long long)gridDim.x * NT) {
    const int r = (int)(i / L), j = (int)(i % L);
    if (off[r] < 0) continue;
    TO* p = dst + off[r] + j;
    dx_st(p, accumulate ? dx_ld(p) + src[i] : src[i]);
  }
}

__global__ void __launch_bounds__(NT) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                  float wd, float bc1, float bc2, const float* __restrict__ gscale_ptr,
                                                  float gscale, const int* __restrict__ step_dev,
                                                  const float* __restrict__ lr_scale_dev, bf16* __restrict__ shadow) {
  const float gs = gscale_ptr ? gscale_ptr[0] * gscale : gscale;
  if (step_dev) {   // CUDA-graph friendly: the step counter (hence the bias correction) lives on the device
    const float st = (float)step_dev[0];
    bc1 = 1.f - powf(b1, st);
    bc2 = 1.f - powf(b2, st);
  }
  if (lr_scale_dev) lr *= lr_scale_dev[0];
  const float inv_bc1 = 1.f / bc1, inv_sbc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  // main part: float4 (flat buffers are 32 B aligned slices); every stream is touched exactly once
  const long long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0 ? (n >> 2) : 0;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    const float ge[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
    float pe[4] = {p4.x, p4.y, p4.z, p4.w}, me[4] = {m4.x, m4.y, m4.z, m4.w}, ve[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      me[j] = b1 * me[j] + (1.f - b1) * ge[j];
      ve[j] = b2 * ve[j] + (1.f - b2) * ge[j] * ge[j];
      pe[j] = pe[j] * decay - (lr * inv_bc1) * me[j] / (sqrtf(ve[j]) * inv_sbc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    if (shadow) {   // bf16 mirror of the updated weights (the tcgen05 GEMMs' operands): 2 B/param here instead of a cast pass
      __nv_bfloat162 lo = __floats2bfloat162_rn(pe[0], pe[1]), hi = __floats2bfloat162_rn(pe[2], pe[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float gi = g[i] * gs;
    float pi = p[i];
    pi -= lr * wd * pi;  // decoupled weight decay (torch.optim.AdamW)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * mi / denom;
    if (shadow) shadow[i] = __float2bfloat16_rn(p[i]);
  }
}

__global__ void __launch_bounds__(NT) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[33];
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) a = fmaf(x[i], x[i], a);
  a = dx_block_sum(a, sh);
  if (threadIdx.x == 0) atomicAdd(out, a);
}

// clip[0] = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_factor_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ clip) {
  clip[0] = fminf(1.f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
}


// out = g * act'(aux): RELU_BWD (aux = activation output or pre-activation, >0 test), TANH_BWD (aux = tanh output),
// GELU_BWD (aux = pre-activation)
template <typename T>
__global__ void __launch_bounds__(NT) act_bwd_kernel(const T* __restrict__ g, const T* __restrict__ aux, T* __restrict__ out,
                                                    long long n, int act) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float a = dx_ld(aux + i);
    float v = dx_ld(g + i);
    if (act == DX_ACT_RELU_BWD) v = a > 0.f ? v : 0.f;
    else if (act == DX_ACT_TANH_BWD) v *= (1.f - a * a);
    else v *= dx_gelu_grad(a);
    dx_st(out + i, v);
  }
}

// Pathology-query logits of PatchDualPathologyPerceiver (models/main_architecture_duett.py:631-639):
//   img = hi + bias_i ; ts = ht + bias_t ; scaled = beta * corr ; fusion = img.detach() + scaled
__global__ void __launch_bounds__(NT) fusion_logits_kernel(const float* __restrict__ hi, const float* __restrict__ ht,
                                                          const float* __restrict__ corr, const float* __restrict__ bi,
                                                          const float* __restrict__ bt, const float* __restrict__ beta,
                                                          float* __restrict__ img, float* __restrict__ ts,
                                                          float* __restrict__ scaled, float* __restrict__ fusion, int B, int K) {
  for (int i = blockIdx.x * NT + threadIdx.x; i < B * K; i += gridDim.x * NT) {
    const int k = i % K;
    const float im = hi[i] + bi[k];
    const float sc = beta[k] * corr[i];
    img[i] = im;
    ts[i] = ht[i] + bt[k];
    scaled[i] = sc;
    fusion[i] = im + sc;
  }
}
// one warp per k: d_corr = (d_fus + d_scaled) * beta ; dbeta += sum_b (d_fus+d_scaled)*corr ; dbi += sum d_img ; dbt += sum d_ts
__global__ void __launch_bounds__(NT) fusion_logits_bwd_kernel(const float* __restrict__ d_img, const float* __restrict__ d_ts,
                                                              const float* __restrict__ d_scaled, const float* __restrict__ d_fus,
                                                              const float* __restrict__ corr, const float* __restrict__ beta,
                                                              float* __restrict__ d_corr, float* __restrict__ dbeta,
                                                              float* __restrict__ dbi, float* __restrict__ dbt, int B, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += NT / 32) {
    float a = 0.f, b1 = 0.f, b2 = 0.f;
    for (int b = lane; b < B; b += 32) {
      const int i = b * K + k;
      const float gs = (d_scaled ? d_scaled[i] : 0.f) + (d_fus ? d_fus[i] : 0.f);
      d_corr[i] = gs * beta[k];
      a = fmaf(gs, corr[i], a);
      b1 += d_img ? d_img[i] : 0.f;
      b2 += d_ts ? d_ts[i] : 0.f;
    }
    a = dx_warp_sum(a); b1 = dx_warp_sum(b1); b2 = dx_warp_sum(b2);
    if (lane == 0) { atomicAdd(dbeta + k, a); atomicAdd(dbi + k, b1); atomicAdd(dbt + k, b2); }
  }
}

inline int grid_for(long long n) {
  long long g = (n + NT - 1) / NT;
  if (g < 1) g = 1;
  return (int)(g < 148 * 8 ? g : 148 * 8);
}

}  // namespace

// y = x * keep / (1 - p), keep regenerated from (seed, flat index): forward and backward of nn.Dropout are the same call
template <typename T>
__global__ void __launch_bounds__(NT) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, DxDrop d,
                                                    const unsigned long long* __restrict__ seed_dev) {
  d = dx_drop_resolve(d, seed_dev);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT)
    dx_st(y + i, dx_ld(x + i) * dx_drop_factor(d, (unsigned long long)i));
}

// out[n] = sum_c a[n,c] * (b[n,c] - bias[c])   (FFN ScaleNorm-backward row dot when dropout sits between GELU and W2)
template <typename T>
__global__ void __launch_bounds__(256) rowdot_bias_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                         const float* __restrict__ bias, float* __restrict__ out, int N, int C) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const T* ar = a + (long long)warp * C;
  const T* br = b + (long long)warp * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot = fmaf(dx_ld(ar + c), dx_ld(br + c) - (bias ? bias[c] : 0.f), dot);
  dot = dx_warp_sum(dot);
  if (lane == 0) out[warp] = dot;
}

extern "C" {

int dx_mean_rows(const void* x, float* y, int B, int T1, int T, int64_t E, int dtype, void* stream) {
  DX_CHECK_ARG(x && y && E % 8 == 0 && T > 0 && T <= T1, "dx_mean_rows: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((long long)B * (E >> 3));
  if (dtype == DX_BF16) mean_rows_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)x, y, B, T1, T, E);
  else mean_rows_kernel<float><<<grid, NT, 0, st>>>((const float*)x, y, B, T1, T, E);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_mean_rows_bwd(const float* dy, void* dx, int B, int T1, int T, int64_t E, int dtype, void* stream) {
  DX_CHECK_ARG(dy && dx && E % 8 == 0 && T > 0 && T <= T1, "dx_mean_rows_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((long long)B * T1 * (E >> 3));
  if (dtype == DX_BF16) mean_rows_bwd_kernel<bf16><<<grid, NT, 0, st>>>(dy, (bf16*)dx, B, T1, T, E);
  else mean_rows_bwd_kernel<float><<<grid, NT, 0, st>>>(dy, (float*)dx, B, T1, T, E);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_gather_vec(const void* src, const int64_t* offsets, float* out, int n, int L, int dtype, void* stream) {
  DX_CHECK_ARG(src && offsets && out && n > 0 && L > 0, "dx_gather_vec: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((long long)n * L);
  if (dtype == DX_BF16) gather_vec_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)src, (const long long*)offsets, out, n, L);
  else gather_vec_kernel<float><<<grid, NT, 0, st>>>((const float*)src, (const long long*)offsets, out, n, L);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_scatter_vec(const float* src, const int64_t* offsets, void* dst, int n, int L, int accumulate, int dtype, void* stream) {
  DX_CHECK_ARG(src && offsets && dst && n > 0 && L > 0, "dx_scatter_vec: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((long long)n * L);
  if (dtype == DX_BF16) scatter_vec_kernel<bf16><<<grid, NT, 0, st>>>(src, (const long long*)offsets, (bf16*)dst, n, L, accumulate);
  else scatter_vec_kernel<float><<<grid, NT, 0, st>>>(src, (const long long*)offsets, (float*)dst, n, L, accumulate);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* out[i] = x[i] * s[i % ns]  (ns = 1: device scalar, e.g. the upstream loss gradient; ns = K: per-column scale) */
__global__ void __launch_bounds__(256) scale_dev_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                       float* __restrict__ out, long long n, int ns) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    out[i] = x[i] * s[ns == 1 ? 0 : (int)(i % ns)];
}

int dx_scale_dev(const float* x, const float* s, float* out, int64_t n, int ns, void* stream) {
  DX_CHECK_ARG(x && s && out && n > 0 && ns > 0, "dx_scale_dev: bad arguments");
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  scale_dev_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(x, s, out, n, ns);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

extern "C++" {
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, long long n, int act) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = dx_ld(x + i);
    dx_st(out + i, act == DX_ACT_RELU ? fmaxf(v, 0.f) : (act == DX_ACT_TANH ? tanhf(v) : dx_gelu(v)));
  }
}
}  // extern "C++"

int dx_act_fwd(const void* x, void* out, int64_t n, int act, int dtype, void* stream) {
  DX_CHECK_ARG(x && out && n > 0 && act >= DX_ACT_GELU && act <= DX_ACT_TANH, "dx_act_fwd: bad arguments");
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (dtype == DX_BF16) act_fwd_kernel<bf16><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, n, act);
  else act_fwd_kernel<float><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)out, n, act);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* sink[0] += (sum_i x[i]) / g[0]   — ScaleNorm gain gradient from the per-row dots (backbone.py) */
__global__ void __launch_bounds__(256) sum_div_acc_kernel(const float* __restrict__ x, long long n, const float* __restrict__ g,
                                                         float* __restrict__ sink) {
  __shared__ float sh[33];
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) a += x[i];
  a = dx_block_sum(a, sh);
  if (threadIdx.x == 0) atomicAdd(sink, a / g[0]);
}

int dx_sum_div_acc(const float* x, int64_t n, const float* g, float* sink, void* stream) {
  DX_CHECK_ARG(x && g && sink && n > 0, "dx_sum_div_acc: bad arguments");
  long long nb = (n + 255) / 256;
  if (nb > 64) nb = 64;
  sum_div_acc_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(x, n, g, sink);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_act_bwd(const void* g, const void* aux, void* out, int64_t n, int act, int dtype, void* stream) {
  DX_CHECK_ARG(g && aux && out && n > 0 && act >= DX_ACT_GELU_BWD && act <= DX_ACT_TANH_BWD, "dx_act_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DX_BF16) act_bwd_kernel<bf16><<<grid_for(n), NT, 0, st>>>((const bf16*)g, (const bf16*)aux, (bf16*)out, n, act);
  else act_bwd_kernel<float><<<grid_for(n), NT, 0, st>>>((const float*)g, (const float*)aux, (float*)out, n, act);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_fusion_logits(const float* hi, const float* ht, const float* corr, const float* bias_i, const float* bias_t,
                     const float* beta, float* img, float* ts, float* scaled, float* fusion, int B, int K, void* stream) {
  DX_CHECK_ARG(hi && ht && corr && bias_i && bias_t && beta && img && ts && scaled && fusion, "dx_fusion_logits: null argument");
  fusion_logits_kernel<<<grid_for((long long)B * K), NT, 0, (cudaStream_t)stream>>>(hi, ht, corr, bias_i, bias_t, beta, img, ts, scaled, fusion, B, K);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* d_img/d_ts/d_scaled/d_fus may be NULL (treated as zero); dbeta/dbias_i/dbias_t are accumulated; d_corr is written.
 * d(hi) = d_img and d(ht) = d_ts pass through unchanged (fusion uses img.detach()). */
int dx_fusion_logits_bwd(const float* d_img, const float* d_ts, const float* d_scaled, const float* d_fus, const float* corr,
                         const float* beta, float* d_corr, float* dbeta, float* dbias_i, float* dbias_t, int B, int K,
                         void* stream) {
  DX_CHECK_ARG(corr && beta && d_corr && dbeta && dbias_i && dbias_t, "dx_fusion_logits_bwd: null argument");
  fusion_logits_bwd_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(d_img, d_ts, d_scaled, d_fus, corr, beta, d_corr, dbeta, dbias_i, dbias_t, B, K);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* AdamW step on flat f32 buffers. Effective gradient = g * grad_scale * (grad_scale_dev ? *grad_scale_dev : 1). */
int dx_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int step, const float* grad_scale_dev, float grad_scale, const int* step_dev,
             const float* lr_scale_dev, void* shadow_bf16, void* stream) {
  DX_CHECK_ARG(p && g && m && v && n > 0 && (step >= 1 || step_dev), "dx_adamw: bad arguments");
  DX_CHECK_ARG(!shadow_bf16 || ((uintptr_t)shadow_bf16 % 8 == 0) || ((uintptr_t)p % 16 != 0), "dx_adamw: shadow must be 8 B aligned");
  const float bc1 = 1.f - powf(beta1, (float)(step >= 1 ? step : 1)), bc2 = 1.f - powf(beta2, (float)(step >= 1 ? step : 1));
  adamw_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                             grad_scale_dev, grad_scale, step_dev, lr_scale_dev,
                                                             reinterpret_cast<bf16*>(shadow_bf16));
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* out[0] += sum x^2 (caller zeroes). */
int dx_sumsq(const float* x, int64_t n, float* out, void* stream) {
  DX_CHECK_ARG(x && out && n > 0, "dx_sumsq: bad arguments");
  sumsq_kernel<<<grid_for(n), NT, 0, (cudaStream_t)stream>>>(x, n, out);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_clip_factor(const float* sumsq, float max_norm, float* clip, void* stream) {
  DX_CHECK_ARG(sumsq && clip, "dx_clip_factor: bad arguments");
  clip_factor_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, clip);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, int dtype, void* stream) {
  DX_CHECK_ARG(x && y && n > 0 && p >= 0.f && p < 1.f, "dx_dropout: bad arguments (0 <= p < 1)");
  const DxDrop d = dx_make_drop(p, seed);
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  if (dtype == DX_BF16) dropout_kernel<bf16><<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n, d, sd);
  else dropout_kernel<float><<<grid_for(n), NT, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, n, d, sd);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_rowdot_bias(const void* a, const void* b, const float* bias, float* out, int N, int C, int dtype, void* stream) {
  DX_CHECK_ARG(a && b && out && N > 0 && C > 0, "dx_rowdot_bias: bad arguments");
  const int grid = dx_ceil_div((long long)N * 32, 256);
  if (dtype == DX_BF16) rowdot_bias_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, bias, out, N, C);
  else rowdot_bias_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, (const float*)b, bias, out, N, C);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
