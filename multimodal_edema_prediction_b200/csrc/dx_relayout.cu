// T<->V axis re-layout fused with ScaleNorm bookkeeping (HBM-bound, one pass per direction).
//
// psi lives as [B, P, Q, d]: tokens of the *source* axis are rows (b,p) with features (q,dd).  The destination
// axis needs rows (b,q) with features (p,dd).  Because dd is contiguous in both views the permutation moves whole
// d-vectors, so both the gather and the store are coalesced 16/32 B vector accesses without a shared-memory tile.
//
// forward  (duett/duett.py:274-279): dst[b,q,p,:] = src[b,p,q,:] * s_src[b,p] + pos        (s_src = final ScaleNorm
//          of the previous Encoder: sqrt(Q*d) * g / ||src row||), and the squared norm of every dst row is produced
//          for the next pre-norm ScaleNorm — the x_transformers final_norm, the axis transpose, the positional
//          add and the next norm's reduction in one read + one write of psi.
// backward: dsrc[b,p,q,:] = s * (gy - src * <src,gy>/||src||^2), gy[b,p,q,:] = gdst[b,q,p,:]; dg += sum <gy,src> * c/||src||.
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr int NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) relayout_fwd_kernel(const T* __restrict__ src, const float* __restrict__ src_rowsq,
                                                         const float* __restrict__ g, const float* __restrict__ pos_b,
                                                         const T* __restrict__ pos_n, T* __restrict__ dst,
                                                         float* __restrict__ dst_rowsq, int B, int P, int Q, int d) {
  __shared__ float sh[33];
  const int row = blockIdx.x;  // b*Q + q
  const int b = row / Q, q = row % Q;
  const int nvec = (P * d) >> 3;
  const float c = g ? sqrtf((float)Q * (float)d) * g[0] : 1.f;
  float ss = 0.f;
  T* drow = dst + (long long)row * P * d;
  for (int i = threadIdx.x; i < nvec; i += NT) {
    const int e = i << 3;
    const int p = e / d, dd = e - p * d;
    float v[8];
    dx_ld8(src + (((long long)b * P + p) * Q + q) * d + dd, v);
    if (src_rowsq) {
      const float s = c / fmaxf(sqrtf(src_rowsq[b * P + p]), 1e-12f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= s;
    }
    if (pos_b) {
      float pv[8];
      dx_ld8(pos_b + (long long)q * P * d + e, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += pv[j];
    }
    if (pos_n) {
      float pv[8];
      dx_ld8(pos_n + (long long)row * P * d + e, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += pv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
    dx_st8(drow + e, v);
  }
  if (dst_rowsq) {
    ss = dx_block_sum(ss, sh);
    if (threadIdx.x == 0) dst_rowsq[row] = ss;
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) relayout_bwd_kernel(const T* __restrict__ gdst, const T* __restrict__ src,
                                                         const float* __restrict__ src_rowsq, const float* __restrict__ g,
                                                         T* __restrict__ dsrc, float* __restrict__ dg, int B, int P, int Q,
                                                         int d) {
  __shared__ float sh[33];
  const int row = blockIdx.x;  // b*P + p
  const int b = row / P, p = row % P;
  const int nvec = (Q * d) >> 3;
  const long long roff = (long long)row * Q * d;
  const bool norm = src_rowsq != nullptr;
  float dot = 0.f;
  if (norm) {
    for (int i = threadIdx.x; i < nvec; i += NT) {
      const int e = i << 3;
      const int q = e / d, dd = e - q * d;
      float gy[8], x[8];
      dx_ld8(gdst + (((long long)b * Q + q) * P + p) * d + dd, gy);
      dx_ld8(src + roff + e, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += gy[j] * x[j];
    }
    dot = dx_block_sum(dot, sh);
  }
  float s = 1.f, k = 0.f;
  if (norm) {
    const float nsq = fmaxf(src_rowsq[row], 1e-24f);
    const float c = sqrtf((float)Q * (float)d);
    const float inv_n = rsqrtf(nsq);
    s = c * g[0] * inv_n;
    k = dot / nsq;
    if (threadIdx.x == 0 && dg) atomicAdd(dg, dot * c * inv_n);
  }
  for (int i = threadIdx.x; i < nvec; i += NT) {
    const int e = i << 3;
    const int q = e / d, dd = e - q * d;
    float gy[8];
    dx_ld8(gdst + (((long long)b * Q + q) * P + p) * d + dd, gy);
    if (norm) {
      float x[8];
      dx_ld8(src + roff + e, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) gy[j] = s * (gy[j] - x[j] * k);
    }
    dx_st8(dsrc + roff + e, gy);
  }
}

// out[n] (+)= sum_m X[m*ld + n]   (bias grads, positional-embedding grads)
template <typename T>
__global__ void __launch_bounds__(NT) colsum_kernel(const T* __restrict__ X, long long ld, int M, long long N,
                                                   float* __restrict__ out, int accumulate, int rows_per_block) {
  const long long n = (long long)blockIdx.x * NT + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += dx_ld(X + (long long)m * ld + n);
  if (gridDim.y == 1 && !accumulate) out[n] = s;
  else atomicAdd(out + n, s);
}

// vector variant: each thread owns 8 adjacent columns (16 B bf16 / 32 B f32 loads; a warp reads 512 contiguous bytes per row)
template <typename T>
__global__ void __launch_bounds__(NT) colsum_vec_kernel(const T* __restrict__ X, long long ld, int M, long long N,
                                                       float* __restrict__ out, int accumulate, int rows_per_block) {
  const long long n = ((long long)blockIdx.x * NT + threadIdx.x) * 8;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const T* p = X + (long long)m0 * ld + n;
  for (int m = m0; m < m1; ++m, p += ld) {
    float v[8];
    dx_ld8(p, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  if (gridDim.y == 1 && !accumulate) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[n + j] = acc[j];
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(out + n + j, acc[j]);
  }
}

// y (+)= alpha * x  over n elements (n % 8 == 0)
template <typename T>
__global__ void __launch_bounds__(NT) axpy_kernel(const T* __restrict__ x, T* __restrict__ y, long long nvec, float alpha,
                                                 int accumulate) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
    float a[8];
    dx_ld8(x + i * 8, a);
    if (accumulate) {
      float b[8];
      dx_ld8(y + i * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = b[j] + alpha * a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] *= alpha;
    }
    dx_st8(y + i * 8, a);
  }
}

// y = x0 + x1 + ... (up to 8 addends, summed in fp32 in argument order, rounded once)
struct SumArgs {
  const void* x[8];
  int n;
};
template <typename T>
__global__ void __launch_bounds__(NT) sum_n_kernel(SumArgs a, T* __restrict__ y, long long nvec) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
    float acc[8];
    dx_ld8(reinterpret_cast<const T*>(a.x[0]) + i * 8, acc);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      if (k < a.n) {
        float b[8];
        dx_ld8(reinterpret_cast<const T*>(a.x[k]) + i * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += b[j];
      }
    }
    dx_st8(y + i * 8, acc);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  const long long nvec = n >> 3;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < nvec; i += (long long)gridDim.x * NT) {
    float a[8];
    dx_ld8(x + i * 8, a);
    dx_st8(y + i * 8, a);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nvec << 3) + threadIdx.x;
    dx_st(y + i, dx_ld(x + i));
  }
}

}  // namespace

extern "C" {

int dx_relayout_fwd(const void* src, const float* src_rowsq, const float* g, const float* pos_bcast,
                    const void* pos_batched, void* dst, float* dst_rowsq, int B, int P, int Q, int d, int act_dtype,
                    void* stream) {
  DX_CHECK_ARG(src && dst, "dx_relayout_fwd: null tensor");
  DX_CHECK_ARG(d % 8 == 0, "dx_relayout_fwd: d_embedding must be a multiple of 8 (got %d)", d);
  DX_CHECK_ARG(!src_rowsq || g, "dx_relayout_fwd: src_rowsq needs g");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * Q;
  if (act_dtype == DX_BF16)
    relayout_fwd_kernel<bf16><<<rows, NT, 0, st>>>((const bf16*)src, src_rowsq, src_rowsq ? g : nullptr, pos_bcast,
                                                   (const bf16*)pos_batched, (bf16*)dst, dst_rowsq, B, P, Q, d);
  else
    relayout_fwd_kernel<float><<<rows, NT, 0, st>>>((const float*)src, src_rowsq, src_rowsq ? g : nullptr, pos_bcast,
                                                    (const float*)pos_batched, (float*)dst, dst_rowsq, B, P, Q, d);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_relayout_bwd(const void* gdst, const void* src, const float* src_rowsq, const float* g, void* dsrc, float* dg,
                    int B, int P, int Q, int d, int act_dtype, void* stream) {
  DX_CHECK_ARG(gdst && dsrc, "dx_relayout_bwd: null tensor");
  DX_CHECK_ARG(d % 8 == 0, "dx_relayout_bwd: d_embedding must be a multiple of 8 (got %d)", d);
  DX_CHECK_ARG(!src_rowsq || (g && src), "dx_relayout_bwd: src_rowsq needs g and src");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * P;
  if (act_dtype == DX_BF16)
    relayout_bwd_kernel<bf16><<<rows, NT, 0, st>>>((const bf16*)gdst, (const bf16*)src, src_rowsq, g, (bf16*)dsrc, dg, B,
                                                   P, Q, d);
  else
    relayout_bwd_kernel<float><<<rows, NT, 0, st>>>((const float*)gdst, (const float*)src, src_rowsq, g, (float*)dsrc, dg,
                                                    B, P, Q, d);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_colsum(const void* X, int64_t ld, int M, int64_t N, float* out, int accumulate, int dtype, void* stream) {
  DX_CHECK_ARG(X && out && M > 0 && N > 0, "dx_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long esz = dtype == DX_BF16 ? 2 : 4;
  const bool vec = (N % 8 == 0) && ((uintptr_t)X % 16 == 0) && ((ld * esz) % 16 == 0);
  const int gx = vec ? dx_ceil_div(N / 8, NT) : dx_ceil_div(N, NT);
  // enough row-splits to fill the machine when N is small
  int gy = 1;
  if (gx < 296) { const int a = dx_ceil_div(M, 64), b = dx_ceil_div(592, gx); gy = a < b ? a : b; }
  const int rpb = dx_ceil_div(M, gy);
  gy = dx_ceil_div(M, rpb);
  if (gy > 1 && !accumulate) DX_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  dim3 grid(gx, gy);
  if (vec) {
    if (dtype == DX_BF16) colsum_vec_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)X, ld, M, N, out, accumulate, rpb);
    else colsum_vec_kernel<float><<<grid, NT, 0, st>>>((const float*)X, ld, M, N, out, accumulate, rpb);
  } else if (dtype == DX_BF16) colsum_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)X, ld, M, N, out, accumulate, rpb);
  else colsum_kernel<float><<<grid, NT, 0, st>>>((const float*)X, ld, M, N, out, accumulate, rpb);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_axpy(const void* x, void* y, int64_t n, float alpha, int accumulate, int dtype, void* stream) {
  DX_CHECK_ARG(x && y && n % 8 == 0, "dx_axpy: n must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nvec = n >> 3;
  long long gl = (nvec + NT - 1) / NT; const int grid = (int)(gl < 148 * 8 ? gl : 148 * 8);
  if (grid <= 0) return DX_OK;
  if (dtype == DX_BF16) axpy_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)x, (bf16*)y, nvec, alpha, accumulate);
  else axpy_kernel<float><<<grid, NT, 0, st>>>((const float*)x, (float*)y, nvec, alpha, accumulate);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_sum_n(const void* const* xs, int count, void* y, int64_t n, int dtype, void* stream) {
  DX_CHECK_ARG(xs && y && count >= 1 && count <= 8 && n % 8 == 0, "dx_sum_n: 1..8 addends, n %% 8 == 0");
  SumArgs a;
  a.n = count;
  for (int k = 0; k < 8; ++k) {
    a.x[k] = k < count ? xs[k] : nullptr;
    DX_CHECK_ARG(k >= count || (xs[k] && (uintptr_t)xs[k] % 16 == 0), "dx_sum_n: addends must be non-null and 16 B aligned");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long nvec = n >> 3;
  long long gl = (nvec + NT - 1) / NT; const int grid = (int)(gl < 148 * 8 ? gl : 148 * 8);
  if (grid <= 0) return DX_OK;
  if (dtype == DX_BF16) sum_n_kernel<bf16><<<grid, NT, 0, st>>>(a, (bf16*)y, nvec);
  else sum_n_kernel<float><<<grid, NT, 0, st>>>(a, (float*)y, nvec);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, void* stream) {
  DX_CHECK_ARG(x && y && n > 0, "dx_cast: bad arguments");
  DX_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "dx_cast: pointers must be 16 B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  long long gl = ((n >> 3) + NT - 1) / NT; if (gl < 1) gl = 1; const int grid = (int)(gl < 148 * 8 ? gl : 148 * 8);
  if (x_dtype == DX_F32 && y_dtype == DX_BF16) cast_kernel<float, bf16><<<grid, NT, 0, st>>>((const float*)x, (bf16*)y, n);
  else if (x_dtype == DX_BF16 && y_dtype == DX_F32) cast_kernel<bf16, float><<<grid, NT, 0, st>>>((const bf16*)x, (float*)y, n);
  else if (x_dtype == DX_F32 && y_dtype == DX_F32) cast_kernel<float, float><<<grid, NT, 0, st>>>((const float*)x, (float*)y, n);
  else cast_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)x, (bf16*)y, n);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
