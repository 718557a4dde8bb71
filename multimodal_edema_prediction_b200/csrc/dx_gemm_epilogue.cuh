// Fused GEMM epilogue shared by the tcgen05 kernel (dx_gemm_tc.cu) and the FFMA kernel (dx_gemm_simt.cu).
// A caller owns a row `m` and CW consecutive accumulator columns starting at n0.
#pragma once
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

struct DxEpi {
  int M, N;
  void* out; long long ldo; int out_dtype; int accumulate;
  void* out2; long long ldo2;
  int act; int act_dtype;
  const float* row_scale; const float* row_scale2; const float* bias;
  const void* res; long long ldr;
  const void* aux; long long ldx; const float* aux_bias;
  const void* cx; long long ldc; const float* coef_num; const float* coef_den;
  float* row_sumsq; float* row_dot;
  int vec_ok;  // all leading dims and base pointers allow 16 B vector access at 8-column granularity
};

static inline DxEpi dx_make_epi(const dx_gemm_desc* d) {
  DxEpi e;
  e.M = d->M; e.N = d->N;
  e.out = d->out; e.ldo = d->ldo; e.out_dtype = d->out_dtype; e.accumulate = d->accumulate;
  e.out2 = d->out2; e.ldo2 = d->ldo2;
  e.act = d->act; e.act_dtype = d->act_dtype;
  e.row_scale = d->row_scale; e.row_scale2 = d->row_scale2; e.bias = d->bias;
  e.res = d->res; e.ldr = d->ldr;
  e.aux = d->aux; e.ldx = d->ldx; e.aux_bias = d->aux_bias;
  e.cx = d->cx; e.ldc = d->ldc; e.coef_num = d->coef_num; e.coef_den = d->coef_den;
  e.row_sumsq = d->row_sumsq; e.row_dot = d->row_dot;
  auto ok = [](const void* p, long long ld, int dtype) {
    if (!p) return true;
    const long long esz = dtype == DX_BF16 ? 2 : 4;
    return ((uintptr_t)p % 16 == 0) && ((ld * esz) % 16 == 0);
  };
  e.vec_ok = ok(d->out, d->ldo, d->out_dtype) && ok(d->out2, d->ldo2, d->act_dtype) &&
             ok(d->res, d->ldr, d->act_dtype) && ok(d->aux, d->ldx, d->act_dtype) &&
             ok(d->cx, d->ldc, d->act_dtype);
  return e;
}

// Load CW (multiple of 8) values of a [.., ld] matrix row as float.
template <int CW>
__device__ __forceinline__ void dx_epi_load(const void* base, long long ld, int dtype, int m, int n0, int nvalid,
                                            bool vec, float (&v)[CW]) {
  if (dtype == DX_BF16) {
    const bf16* p = reinterpret_cast<const bf16*>(base) + (long long)m * ld + n0;
    if (vec && nvalid == CW) {
#pragma unroll
      for (int i = 0; i < CW; i += 8) {
        float t[8];
        dx_ld8(p + i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i + j] = t[j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = (i < nvalid) ? __bfloat162float(p[i]) : 0.f;
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + (long long)m * ld + n0;
    if (vec && nvalid == CW) {
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = (i < nvalid) ? p[i] : 0.f;
    }
  }
}

template <int CW>
__device__ __forceinline__ void dx_epi_store(void* base, long long ld, int dtype, int accumulate, int m, int n0,
                                             int nvalid, bool vec, const float (&v)[CW]) {
  if (dtype == DX_BF16) {
    bf16* p = reinterpret_cast<bf16*>(base) + (long long)m * ld + n0;
    if (vec && nvalid == CW) {
#pragma unroll
      for (int i = 0; i < CW; i += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = v[i + j];
        dx_st8(p + i, t);
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + (long long)m * ld + n0;
    if (vec && nvalid == CW) {
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (accumulate) {
          float4 o = *reinterpret_cast<const float4*>(p + i);
          t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *reinterpret_cast<float4*>(p + i) = t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (i < nvalid) p[i] = accumulate ? p[i] + v[i] : v[i];
    }
  }
}

// Applies the epilogue to CW accumulator columns of row m (m < M guaranteed by the caller).
// rs_acc / rd_acc collect the row reductions; the caller flushes them with one atomicAdd per row.
template <int CW>
__device__ __forceinline__ void dx_epilogue_chunk(const DxEpi& e, int m, int n0, float (&v)[CW], float& rs_acc,
                                                  float& rd_acc) {
  const int nvalid = min(CW, e.N - n0);
  if (nvalid <= 0) return;
  const bool vec = e.vec_ok != 0;
  if (e.row_scale) {
    const float s = e.row_scale[m];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] *= s;
  }
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] += (i < nvalid) ? __ldg(e.bias + n0 + i) : 0.f;
  }
  if (e.act == DX_ACT_GELU) {
    if (e.out2) dx_epi_store<CW>(e.out2, e.ldo2, e.act_dtype, 0, m, n0, nvalid, vec, v);
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = dx_gelu(v[i]);
  } else if (e.act == DX_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (e.act == DX_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = tanhf(v[i]);
  } else if (e.act >= DX_ACT_GELU_BWD) {
    float a[CW];
    dx_epi_load<CW>(e.aux, e.ldx, e.act_dtype, m, n0, nvalid, vec, a);
    if (e.act == DX_ACT_GELU_BWD) {
      float rd = 0.f;
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        v[i] *= dx_gelu_grad(a[i]);
        const float ab = (e.aux_bias && i < nvalid) ? __ldg(e.aux_bias + n0 + i) : 0.f;
        rd += (i < nvalid) ? v[i] * (a[i] - ab) : 0.f;
      }
      rd_acc += rd;
      if (e.out2) dx_epi_store<CW>(e.out2, e.ldo2, e.act_dtype, 0, m, n0, nvalid, vec, v);
      if (e.row_scale2) {
        const float s2 = e.row_scale2[m];
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] *= s2;
      }
    } else if (e.act == DX_ACT_RELU_BWD) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = a[i] > 0.f ? v[i] : 0.f;
    } else {  // TANH_BWD
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] *= (1.f - a[i] * a[i]);
    }
  }
  if (e.res) {
    float r[CW];
    dx_epi_load<CW>(e.res, e.ldr, e.act_dtype, m, n0, nvalid, vec, r);
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] += r[i];
  }
  if (e.cx) {
    float c[CW];
    dx_epi_load<CW>(e.cx, e.ldc, e.act_dtype, m, n0, nvalid, vec, c);
    const float coef = e.coef_num[m] / fmaxf(e.coef_den[m], 1e-24f);
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] -= c[i] * coef;
  }
  if (e.row_sumsq) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) s += (i < nvalid) ? v[i] * v[i] : 0.f;
    rs_acc += s;
  }
  if (e.out) dx_epi_store<CW>(e.out, e.ldo, e.out_dtype, e.accumulate, m, n0, nvalid, vec, v);
}

__device__ __forceinline__ void dx_epilogue_flush_row(const DxEpi& e, int m, float rs_acc, float rd_acc) {
  if (e.row_sumsq) atomicAdd(e.row_sumsq + m, rs_acc);
  if (e.row_dot) atomicAdd(e.row_dot + m, rd_acc);
}
