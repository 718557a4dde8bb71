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
#include <cstdlib>
#include "../../include/duett_b200.h"

namespace {

constexpr int NT = 256;

// raw 16 B (bf16) / 32 B (f32) vector of 8 elements, kept packed until it is used (so that several loads can be in flight)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void raw_ld(const bf16* p, Raw8<bf16>& r) { r.u = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void raw_ld(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void raw_unpack(const Raw8<bf16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
  v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

// One CTA per destination row.  Every thread keeps TWO independent vectors in flight (loads of both issued before either
// is used): with one load per thread the kernel is bound by bytes in flight per SM (2048 threads x 16 B), not by HBM.
template <typename T>
__global__ void __launch_bounds__(NT) relayout_fwd_kernel(const T* __restrict__ src, const float* __restrict__ src_rowsq,
                                                         const float* __restrict__ g, const float* __restrict__ pos_b,
                                                         const T* __restrict__ pos_n, T* __restrict__ dst,
                                                         float* __restrict__ dst_rowsq, int B, int P, int Q, int d) {
  __shared__ float sh[33];
  const int nt = blockDim.x;
  const int row = blockIdx.x;  // b*Q + q
  const int b = row / Q, q = row % Q;
  const int nvec = (P * d) >> 3;
  const float c = g ? sqrtf((float)Q * (float)d) * g[0] : 1.f;
  float ss = 0.f;
  T* drow = dst + (long long)row * P * d;
  const T* sbase = src + ((long long)b * P * Q + q) * d;
  const float* pb = pos_b ? pos_b + (long long)q * P * d : nullptr;
  const T* pn = pos_n ? pos_n + (long long)row * P * d : nullptr;
  const float* rsq = src_rowsq ? src_rowsq + b * P : nullptr;

  auto finish = [&](int e, const Raw8<T>& rv, float rq, const Raw8<float>& rpb, const Raw8<T>& rpn) {
    float v[8];
    raw_unpack(rv, v);
    if (rsq) {
      const float s = c / fmaxf(sqrtf(rq), 1e-12f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= s;
    }
    if (pb) {
      float pv[8];
      raw_unpack(rpb, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += pv[j];
    }
    if (pn) {
      float pv[8];
      raw_unpack(rpn, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += pv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
    dx_st8(drow + e, v);
  };
  auto fetch = [&](int i, Raw8<T>& rv, float& rq, Raw8<float>& rpb, Raw8<T>& rpn) {
    const int e = i << 3;
    const int p = e / d, dd = e - p * d;
    raw_ld(sbase + (long long)p * Q * d + dd, rv);
    rq = rsq ? rsq[p] : 1.f;
    if (pb) raw_ld(pb + e, rpb);
    if (pn) raw_ld(pn + e, rpn);
  };

  int i = threadIdx.x;
  for (; i + nt < nvec; i += 2 * nt) {
    Raw8<T> v0, v1, n0, n1;
    Raw8<float> p0, p1;
    float q0, q1;
    fetch(i, v0, q0, p0, n0);
    fetch(i + nt, v1, q1, p1, n1);
    finish(i << 3, v0, q0, p0, n0);
    finish((i + nt) << 3, v1, q1, p1, n1);
  }
  if (i < nvec) {
    Raw8<T> v0, n0;
    Raw8<float> p0;
    float q0;
    fetch(i, v0, q0, p0, n0);
    finish(i << 3, v0, q0, p0, n0);
  }
  if (dst_rowsq) {
    ss = dx_block_sum(ss, sh);
    if (threadIdx.x == 0) dst_rowsq[row] = ss;
  }
}

// Backward: dsrc[b,p,q,:] = s * (gy - k x) with gy = gdst[b,q,p,:] (final-ScaleNorm backward of the source row) or a plain
// transpose.  The row of gy and x is staged in shared memory by the dot-product pass (when it fits), so the second pass
// does not go back to L2/HBM.
template <typename T>
__global__ void __launch_bounds__(NT) relayout_bwd_kernel(const T* __restrict__ gdst, const T* __restrict__ src,
                                                         const float* __restrict__ src_rowsq, const float* __restrict__ g,
                                                         T* __restrict__ dsrc, float* __restrict__ dg, int B, int P, int Q,
                                                         int d, int stage_row) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ float sh[33];
  const int nt = blockDim.x;
  const int row = blockIdx.x;  // b*P + p
  const int b = row / P, p = row % P;
  const int nvec = (Q * d) >> 3;
  const long long roff = (long long)row * Q * d;
  const T* gbase = gdst + ((long long)b * Q * P + p) * d;
  const bool norm = src_rowsq != nullptr;
  Raw8<T>* sg = reinterpret_cast<Raw8<T>*>(smraw);   // [nvec] gy, then [nvec] x  (stage_row only)
  Raw8<T>* sx = sg + nvec;
  float dot = 0.f;
  if (norm) {
    int i = threadIdx.x;
    for (; i + nt < nvec; i += 2 * nt) {
      Raw8<T> g0, g1, x0, x1;
      const int e0 = i << 3, e1 = (i + nt) << 3;
      const int q0 = e0 / d, q1 = e1 / d;
      raw_ld(gbase + (long long)q0 * P * d + (e0 - q0 * d), g0);
      raw_ld(gbase + (long long)q1 * P * d + (e1 - q1 * d), g1);
      raw_ld(src + roff + e0, x0);
      raw_ld(src + roff + e1, x1);
      float a[8], c[8];
      raw_unpack(g0, a); raw_unpack(x0, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += a[j] * c[j];
      raw_unpack(g1, a); raw_unpack(x1, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += a[j] * c[j];
      if (stage_row) { sg[i] = g0; sx[i] = x0; sg[i + nt] = g1; sx[i + nt] = x1; }
    }
    if (i < nvec) {
      Raw8<T> g0, x0;
      const int e0 = i << 3, q0 = e0 / d;
      raw_ld(gbase + (long long)q0 * P * d + (e0 - q0 * d), g0);
      raw_ld(src + roff + e0, x0);
      float a[8], c[8];
      raw_unpack(g0, a); raw_unpack(x0, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot += a[j] * c[j];
      if (stage_row) { sg[i] = g0; sx[i] = x0; }
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
  if (norm && stage_row) {
    // every thread re-reads exactly the slots it wrote: no barrier needed beyond the one inside dx_block_sum
    for (int i = threadIdx.x; i < nvec; i += nt) {
      float gy[8], x[8];
      raw_unpack(sg[i], gy);
      raw_unpack(sx[i], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) gy[j] = s * (gy[j] - x[j] * k);
      dx_st8(dsrc + roff + (i << 3), gy);
    }
    return;
  }
  int i = threadIdx.x;
  for (; i + nt < nvec; i += 2 * nt) {
    Raw8<T> g0, g1, x0, x1;
    const int e0 = i << 3, e1 = (i + nt) << 3;
    const int q0 = e0 / d, q1 = e1 / d;
    raw_ld(gbase + (long long)q0 * P * d + (e0 - q0 * d), g0);
    raw_ld(gbase + (long long)q1 * P * d + (e1 - q1 * d), g1);
    if (norm) { raw_ld(src + roff + e0, x0); raw_ld(src + roff + e1, x1); }
    float gy[8], x[8];
    raw_unpack(g0, gy);
    if (norm) {
      raw_unpack(x0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) gy[j] = s * (gy[j] - x[j] * k);
    }
    dx_st8(dsrc + roff + e0, gy);
    raw_unpack(g1, gy);
    if (norm) {
      raw_unpack(x1, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) gy[j] = s * (gy[j] - x[j] * k);
    }
    dx_st8(dsrc + roff + e1, gy);
  }
  if (i < nvec) {
    Raw8<T> g0, x0;
    const int e0 = i << 3, q0 = e0 / d;
    raw_ld(gbase + (long long)q0 * P * d + (e0 - q0 * d), g0);
    float gy[8], x[8];
    raw_unpack(g0, gy);
    if (norm) {
      raw_ld(src + roff + e0, x0);
      raw_unpack(x0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) gy[j] = s * (gy[j] - x[j] * k);
    }
    dx_st8(dsrc + roff + e0, gy);
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

// strip variant (round 2): a warp owns a 32-vector (256-column) strip, the 8 warps of a CTA interleave over the rows of the
// CTA's row chunk, 4 rows in flight per thread; the 8 partial strips are summed through shared memory and leave as one
// atomicAdd per column.  Narrow matrices (N = 512: two strips) keep every thread busy and the grid is sized by rows, not by
// columns — the single-row-per-iteration kernel above ran the path's bias / positional sums at 3.4 TB/s.
constexpr int CS_WARPS = 8, CS_UNROLL = 4;
template <typename T>
__global__ void __launch_bounds__(32 * CS_WARPS) colsum_strip_kernel(const T* __restrict__ X, long long ld, int M, long long N,
                                                                    float* __restrict__ out, int direct, int rows_per_block) {
  __shared__ float part[CS_WARPS][32][9];   // +1: conflict-free column reads
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long n = ((long long)blockIdx.x * 32 + lane) * 8;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const T* p = X + n;
    int m = m0 + w;
    for (; m + (CS_UNROLL - 1) * CS_WARPS < m1; m += CS_UNROLL * CS_WARPS) {
      float v[CS_UNROLL][8];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) dx_ld8(p + (long long)(m + u * CS_WARPS) * ld, v[u]);
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; m < m1; m += CS_WARPS) {
      float v[8];
      dx_ld8(p + (long long)m * ld, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[w][lane][j] = acc[j];
  __syncthreads();
  // thread t sums column t of the strip over the 8 warps
  const int c = threadIdx.x;               // 0 .. 255
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CS_WARPS; ++k) s += part[k][c >> 3][c & 7];
  const long long col = (long long)blockIdx.x * 256 + c;
  if (col < N) {
    if (direct) out[col] = s;
    else atomicAdd(out + col, s);
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

// short rows: smaller CTAs keep more rows in flight per SM (the per-CTA latency chain load -> reduce -> store, not HBM,
// bounds a one-row-per-CTA kernel).  DX_RELAYOUT_NT overrides.
static int relayout_threads(int nvec, bool bwd) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DX_RELAYOUT_NT");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  // measured on B200 at the bench shapes (tools/relayout_bench.py): rows of 528 vectors: fwd 123 / 135 / 179 us and bwd
  // 172 / 160 / 212 us with 64 / 128 / 256 threads; rows of 2064 vectors are insensitive (fwd) or best at 256 (bwd)
  if (nvec <= 640) return bwd ? 128 : 64;
  return 256;
}

extern "C" {

int dx_relayout_fwd(const void* src, const float* src_rowsq, const float* g, const float* pos_bcast,
                    const void* pos_batched, void* dst, float* dst_rowsq, int B, int P, int Q, int d, int act_dtype,
                    void* stream) {
  DX_CHECK_ARG(src && dst, "dx_relayout_fwd: null tensor");
  DX_CHECK_ARG(d % 8 == 0, "dx_relayout_fwd: d_embedding must be a multiple of 8 (got %d)", d);
  DX_CHECK_ARG(!src_rowsq || g, "dx_relayout_fwd: src_rowsq needs g");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * Q;
  const int nt = relayout_threads((P * d) >> 3, false);
  if (act_dtype == DX_BF16)
    relayout_fwd_kernel<bf16><<<rows, nt, 0, st>>>((const bf16*)src, src_rowsq, src_rowsq ? g : nullptr, pos_bcast,
                                                   (const bf16*)pos_batched, (bf16*)dst, dst_rowsq, B, P, Q, d);
  else
    relayout_fwd_kernel<float><<<rows, nt, 0, st>>>((const float*)src, src_rowsq, src_rowsq ? g : nullptr, pos_bcast,
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
  // stage the (gy, x) row in shared memory when it is short enough not to limit the CTAs per SM
  const size_t esz = act_dtype == DX_BF16 ? 2 : 4;
  const size_t row_bytes = 2 * (size_t)Q * d * esz;
  const int stage_row = (src_rowsq && row_bytes <= 24 * 1024) ? 1 : 0;   // long rows: staging would cost occupancy
  const size_t smem = stage_row ? row_bytes : 0;
  const int nt = relayout_threads((Q * d) >> 3, true);
  if (act_dtype == DX_BF16) {
    auto kern = relayout_bwd_kernel<bf16>;
    static size_t attr = 48 * 1024;
    if (smem > attr) { DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
    kern<<<rows, nt, smem, st>>>((const bf16*)gdst, (const bf16*)src, src_rowsq, g, (bf16*)dsrc, dg, B, P, Q, d, stage_row);
  } else {
    auto kern = relayout_bwd_kernel<float>;
    static size_t attr = 48 * 1024;
    if (smem > attr) { DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
    kern<<<rows, nt, smem, st>>>((const float*)gdst, (const float*)src, src_rowsq, g, (float*)dsrc, dg, B, P, Q, d, stage_row);
  }
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_colsum(const void* X, int64_t ld, int M, int64_t N, float* out, int accumulate, int dtype, void* stream) {
  DX_CHECK_ARG(X && out && M > 0 && N > 0, "dx_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long esz = dtype == DX_BF16 ? 2 : 4;
  const bool vec = (N % 8 == 0) && ((uintptr_t)X % 16 == 0) && ((ld * esz) % 16 == 0);
  static const bool v1 = [] { const char* e = getenv("DX_COLSUM_V1"); return e && atoi(e) != 0; }();   // A/B: the round-1 kernel
  if (vec && !v1) {
    const int strips = dx_ceil_div(N / 8, 32);
    // row chunks: ~8 CTAs per SM over the whole grid, at least CS_UNROLL * CS_WARPS rows each
    int gy = dx_ceil_div(148 * 8, strips);
    const int max_gy = dx_ceil_div(M, CS_UNROLL * CS_WARPS);
    if (gy > max_gy) gy = max_gy;
    const int rpb = dx_ceil_div(M, gy);
    gy = dx_ceil_div(M, rpb);
    const int direct = (gy == 1 && !accumulate) ? 1 : 0;
    if (gy > 1 && !accumulate) DX_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
    dim3 grid(strips, gy);
    if (dtype == DX_BF16) colsum_strip_kernel<bf16><<<grid, 32 * CS_WARPS, 0, st>>>((const bf16*)X, ld, M, N, out, direct, rpb);
    else colsum_strip_kernel<float><<<grid, 32 * CS_WARPS, 0, st>>>((const float*)X, ld, M, N, out, direct, rpb);
    DX_LAUNCH_CHECK();
    return DX_OK;
  }
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
