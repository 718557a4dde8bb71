// Fused GEMM epilogue shared by the tcgen05 kernel (dx_gemm_tc.cu) and the FFMA kernel (dx_gemm_simt.cu).
// A caller owns a row `m` and CW consecutive accumulator columns starting at n0.  The arithmetic lives in
// dx_epilogue_math (registers only); the two front-ends differ in how the [M,N] side tensors reach registers:
//   dx_epilogue_chunk        : direct global loads/stores (thread-per-row vectors; FFMA kernel, unaligned fallbacks)
//   tcgen05 staged epilogue  : warp-staged through swizzled shared memory so every global access is a full 128 B row
//                              segment (dx_gemm_tc.cu)
#pragma once
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

int dx_gemm_reserved_sms();   // dx_api.cu

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
  // batched mode (blockIdx.z = batch index): element offsets added per batch
  long long out_bs, out2_bs, res_bs, aux_bs, cx_bs; int bias_bs; int rowvec_bs;
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
  auto ok = [](const void* p, long long ld, long long bs, int dtype) {
    if (!p) return true;
    const long long esz = dtype == DX_BF16 ? 2 : 4;
    return ((uintptr_t)p % 16 == 0) && ((ld * esz) % 16 == 0) && ((bs * esz) % 16 == 0);
  };
  e.out_bs = d->out_bs; e.out2_bs = d->out2_bs; e.res_bs = d->res_bs; e.aux_bs = d->aux_bs; e.cx_bs = d->cx_bs;
  e.bias_bs = d->bias_bs; e.rowvec_bs = d->rowvec_bs;
  e.vec_ok = ok(d->out, d->ldo, d->out_bs, d->out_dtype) && ok(d->out2, d->ldo2, d->out2_bs, d->act_dtype) &&
             ok(d->res, d->ldr, d->res_bs, d->act_dtype) && ok(d->aux, d->ldx, d->aux_bs, d->act_dtype) &&
             ok(d->cx, d->ldc, d->cx_bs, d->act_dtype);
  return e;
}

// Shift every pointer of the epilogue to batch z (called once per CTA).
__device__ __forceinline__ void dx_epi_select_batch(DxEpi& e, int z) {
  if (z == 0) return;
  const long long osz = e.out_dtype == DX_BF16 ? 2 : 4, asz = e.act_dtype == DX_BF16 ? 2 : 4;
  if (e.out) e.out = (char*)e.out + z * e.out_bs * osz;
  if (e.out2) e.out2 = (char*)e.out2 + z * e.out2_bs * asz;
  if (e.res) e.res = (const char*)e.res + z * e.res_bs * asz;
  if (e.aux) e.aux = (const char*)e.aux + z * e.aux_bs * asz;
  if (e.cx) e.cx = (const char*)e.cx + z * e.cx_bs * asz;
  if (e.bias) e.bias += (long long)z * e.bias_bs;
  if (e.aux_bias) e.aux_bias += (long long)z * e.bias_bs;
  const long long rv = (long long)z * e.rowvec_bs;
  if (e.row_scale) e.row_scale += rv;
  if (e.row_scale2) e.row_scale2 += rv;
  if (e.coef_num) e.coef_num += rv;
  if (e.coef_den) e.coef_den += rv;
  if (e.row_sumsq) e.row_sumsq += rv;
  if (e.row_dot) e.row_dot += rv;
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

// Per-row constants of the epilogue, loaded / computed once per (tile,row) instead of once per 8-column piece.
struct DxRowConst {
  float rs, rs2, coef;
};
// Raw per-row operands of DxRowConst: loading them one tile ahead (and dividing only when the tile starts) keeps the
// dependent global loads off the critical path of the persistent epilogue.
struct DxRowRaw {
  float rs, rs2, num, den;
};
__device__ __forceinline__ DxRowRaw dx_row_raw(const DxEpi& e, int m) {
  DxRowRaw r;
  r.rs = e.row_scale ? e.row_scale[m] : 1.f;
  r.rs2 = e.row_scale2 ? e.row_scale2[m] : 1.f;
  r.num = e.cx ? e.coef_num[m] : 0.f;
  r.den = e.cx ? e.coef_den[m] : 1.f;
  return r;
}
__device__ __forceinline__ DxRowConst dx_row_finish(const DxRowRaw& r) {
  DxRowConst c;
  c.rs = r.rs;
  c.rs2 = r.rs2;
  c.coef = r.num / fmaxf(r.den, 1e-24f);
  return c;
}
__device__ __forceinline__ DxRowConst dx_row_const(const DxEpi& e, int m) {
  DxRowConst c;
  c.rs = e.row_scale ? e.row_scale[m] : 1.f;
  c.rs2 = e.row_scale2 ? e.row_scale2[m] : 1.f;
  c.coef = e.cx ? e.coef_num[m] / fmaxf(e.coef_den[m], 1e-24f) : 0.f;
  return c;
}

// split-K partial sums: out (f32) += v with red.global.add (the destination is a gradient accumulator)
__device__ __forceinline__ void dx_epi_atomic_add8(void* base, long long ld, int m, int n0, int N, const float (&v)[8]) {
  float* p = reinterpret_cast<float*>(base) + (long long)m * ld + n0;
  if (n0 + 8 <= N && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {   // red.global.add.v4.f32
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (n0 + i < N) atomicAdd(p + i, v[i]);
}

// Register-only arithmetic of the epilogue.  v: accumulator columns (in) -> final `out` values (out).
// r / a / c: preloaded residual, aux and cx values (ignored when the corresponding pointer in `e` is null).
// o2: receives the `out2` values when has_out2() (pre-activation for GELU, unscaled dpre for GELU_BWD).
template <int CW>
__device__ __forceinline__ void dx_epilogue_math(const DxEpi& e, const DxRowConst& rc, int n0, int nvalid, float (&v)[CW],
                                                 const float (&r)[CW], const float (&a)[CW], const float (&c)[CW],
                                                 float (&o2)[CW], float& rs_acc, float& rd_acc) {
  if (e.row_scale) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] *= rc.rs;
  }
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] += (i < nvalid) ? __ldg(e.bias + n0 + i) : 0.f;
  }
  if (e.act == DX_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < CW; ++i) { o2[i] = v[i]; v[i] = dx_gelu(v[i]); }
  } else if (e.act == DX_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (e.act == DX_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = tanhf(v[i]);
  } else if (e.act == DX_ACT_GELU_BWD) {
    float rd = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      v[i] *= dx_gelu_grad(a[i]);
      const float ab = (e.aux_bias && i < nvalid) ? __ldg(e.aux_bias + n0 + i) : 0.f;
      rd += (i < nvalid) ? v[i] * (a[i] - ab) : 0.f;
      o2[i] = v[i];
    }
    rd_acc += rd;
    if (e.row_scale2) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] *= rc.rs2;
    }
  } else if (e.act == DX_ACT_RELU_BWD) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = a[i] > 0.f ? v[i] : 0.f;
  } else if (e.act == DX_ACT_TANH_BWD) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] *= (1.f - a[i] * a[i]);
  }
  if (e.res) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] += r[i];
  }
  if (e.cx) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] -= c[i] * rc.coef;
  }
  if (e.row_sumsq) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) s += (i < nvalid) ? v[i] * v[i] : 0.f;
    rs_acc += s;
  }
}

// Compile-time specialised arithmetic: MASK is a bit set of the features a launch uses (the host computes it once), so the
// per-piece code has no feature branches.  Every combination the DuETT path produces has an instantiation in
// dx_gemm_tc.cu; anything else falls back to the runtime-flag version above.
enum : int { DX_M_RS = 1, DX_M_BIAS = 2, DX_M_GELU = 4, DX_M_RES = 8, DX_M_ROWSQ = 16, DX_M_CX = 32, DX_M_GELUBWD = 64 };

static inline int dx_epi_mask(const dx_gemm_desc* d) {
  // returns -1 when the launch uses a feature combination without a specialisation
  int m = 0;
  if (d->row_scale) m |= DX_M_RS;
  if (d->bias) m |= DX_M_BIAS;
  if (d->res) m |= DX_M_RES;
  if (d->row_sumsq) m |= DX_M_ROWSQ;
  if (d->cx) m |= DX_M_CX;
  if (d->act == DX_ACT_GELU) { if (!d->out2) return -1; m |= DX_M_GELU; }
  else if (d->act == DX_ACT_GELU_BWD) { if (!d->out2 || !d->aux_bias || !d->row_scale2 || !d->row_dot) return -1; m |= DX_M_GELUBWD; }
  else if (d->act != DX_ACT_NONE) return -1;
  else if (d->aux || d->row_scale2 || d->row_dot) return -1;
  // the specialised path stages bias / aux_bias through shared memory with 16 B cp.async
  if ((d->bias && ((uintptr_t)d->bias & 15)) || (d->aux_bias && ((uintptr_t)d->aux_bias & 15)) || (d->bias_bs & 3)) return -1;
  switch (m) {
    case 0: case DX_M_BIAS: case DX_M_RS: case DX_M_RS | DX_M_BIAS | DX_M_GELU: case DX_M_RES | DX_M_ROWSQ:
    case DX_M_BIAS | DX_M_RES | DX_M_ROWSQ: case DX_M_GELUBWD: case DX_M_RES | DX_M_CX:
      return m;
    default:
      return -1;
  }
}

// bv: the 8 per-column bias values of this piece (bias for BIAS, aux_bias for GELUBWD), staged by the caller.
template <int MASK>
__device__ __forceinline__ void dx_epilogue_math_c(const DxRowConst& rc, float (&v)[8], const float (&r)[8], const float (&a)[8],
                                                   const float (&bv)[8], float (&o2)[8], float& rs_acc, float& rd_acc) {
  if constexpr ((MASK & DX_M_RS) != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= rc.rs;
  }
  if constexpr ((MASK & DX_M_BIAS) != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += bv[i];
  }
  if constexpr ((MASK & DX_M_GELU) != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { o2[i] = v[i]; v[i] = dx_gelu(v[i]); }
  }
  if constexpr ((MASK & DX_M_GELUBWD) != 0) {
    float rd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] *= dx_gelu_grad(a[i]);
      rd = fmaf(v[i], a[i] - bv[i], rd);
      o2[i] = v[i];
      v[i] *= rc.rs2;
    }
    rd_acc += rd;
  }
  if constexpr ((MASK & DX_M_RES) != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  }
  if constexpr ((MASK & DX_M_CX) != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(-a[i], rc.coef, v[i]);
  }
  if constexpr ((MASK & DX_M_ROWSQ) != 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(v[i], v[i], s);
    rs_acc += s;
  }
}

__device__ __forceinline__ bool dx_epi_has_out2(const DxEpi& e) {
  return e.out2 != nullptr && (e.act == DX_ACT_GELU || e.act == DX_ACT_GELU_BWD);
}

// Direct-global front-end: applies the epilogue to CW accumulator columns of row m (m < M guaranteed by the caller),
// 8 columns (one 16 B bf16 vector) at a time to keep the live register set small.
// rs_acc / rd_acc collect the row reductions; the caller flushes them with one atomicAdd per row.
__device__ __forceinline__ void dx_epilogue_piece(const DxEpi& e, const DxRowConst& rc, int m, int n0, float (&v)[8],
                                                  float& rs_acc, float& rd_acc) {
  const int nvalid = min(8, e.N - n0);
  if (nvalid <= 0) return;
  const bool vec = e.vec_ok != 0;
  float r[8], a[8], c[8], o2[8];
  if (e.res) dx_epi_load<8>(e.res, e.ldr, e.act_dtype, m, n0, nvalid, vec, r);
  if (e.aux) dx_epi_load<8>(e.aux, e.ldx, e.act_dtype, m, n0, nvalid, vec, a);
  if (e.cx) dx_epi_load<8>(e.cx, e.ldc, e.act_dtype, m, n0, nvalid, vec, c);
  dx_epilogue_math<8>(e, rc, n0, nvalid, v, r, a, c, o2, rs_acc, rd_acc);
  if (dx_epi_has_out2(e)) dx_epi_store<8>(e.out2, e.ldo2, e.act_dtype, 0, m, n0, nvalid, vec, o2);
  if (e.out) dx_epi_store<8>(e.out, e.ldo, e.out_dtype, e.accumulate, m, n0, nvalid, vec, v);
}

template <int CW>
__device__ __forceinline__ void dx_epilogue_chunk(const DxEpi& e, int m, int n0, float (&v)[CW], float& rs_acc,
                                                  float& rd_acc) {
  const DxRowConst rc = dx_row_const(e, m);
#pragma unroll
  for (int i = 0; i < CW; i += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[i + j];
    dx_epilogue_piece(e, rc, m, n0 + i, t, rs_acc, rd_acc);
  }
}

__device__ __forceinline__ void dx_epilogue_flush_row(const DxEpi& e, int m, float rs_acc, float rd_acc) {
  if (e.row_sumsq) atomicAdd(e.row_sumsq + m, rs_acc);
  if (e.row_dot) atomicAdd(e.row_dot + m, rd_acc);
}
