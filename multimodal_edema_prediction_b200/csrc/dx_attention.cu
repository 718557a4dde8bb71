// Unmasked multi-head softmax attention for tiny sequences (DuETT: S = V+1 event tokens or T+1 time tokens,
// head dim d/2; perceiver: 7 pathology queries over 1369 patch / 24 hour tokens).  x_transformers Attention core
// (duett/duett.py:95-105) and nn.MultiheadAttention core (models/main_architecture_duett.py:752): softmax(q k^T / sqrt(dh)) v
// with fp32 softmax.  The whole problem lives in shared memory/registers; FLOPs are <0.3 % of the step,
// so this is a warp-shuffle FFMA kernel, not a tensor-core one.
//
// forward : one thread group (TPQ lanes) per query row, keys/values streamed through shared memory in tiles of 64,
//           online softmax; writes o and the row log-sum-exp.
// backward: kernel A (per query) -> D = <do,o>, dq ; kernel B (per key) -> dk, dv, streaming query tiles.
#include "dx_common.cuh"
#include "../../include/duett_b200.h"
#include <cstdlib>

// bf16 tensor-core path for short self-attention (dx_attention_mma.cu)
bool dx_attn_mma_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh);
bool dx_attn_mmat_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh);
// tcgen05 / TMEM forward (dx_attention_tc.cu): dh = 64, Sk <= 256, Sq >= 64 — the event axis of the DuETT blocks
bool dx_attn_tc_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh);
int dx_attn_tc_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                   long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                   int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st);
int dx_attn_mmat_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                     long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                     int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st);
int dx_attn_mmat_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                     long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs, const void* go, long long go_bs,
                     long long go_rs, void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                     void* dv, long long dv_bs, long long dv_rs, const float* lse, float* Dws, int B, int H, int Sq, int Sk, int dh,
                     DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st);
int dx_attn_mma_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                    long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                    int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st);
int dx_attn_mma_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                    long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs, const void* go, long long go_bs,
                    long long go_rs, void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                    void* dv, long long dv_bs, long long dv_rs, const float* lse, int B, int H, int Sq, int Sk, int dh,
                    DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st);

namespace {

// DX_ATTN_SIMT=1 forces the SIMT kernels (A/B timing, debugging)
bool attn_force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DX_ATTN_SIMT");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

constexpr int KT = 64;    // keys (or queries) per shared-memory tile
constexpr int NTH = 128;  // threads per block

struct AttnView {      // element (b, s, head h, i) at base + b*bs + s*rs + h*dh + i
  const void* p;
  long long bs, rs;
};
struct AttnViewW {
  void* p;
  long long bs, rs;
};

template <typename T>
__device__ __forceinline__ float ldv(const void* base, long long off) {
  return dx_ld(reinterpret_cast<const T*>(base) + off);
}

// Loads `rows` rows x DH of a [S, DH]-strided head slice into shared memory as fp32 (zero padded).
template <typename T, int DH>
__device__ __forceinline__ void load_tile(float* sm, const AttnView& v, int b, int h, int s0, int S) {
  for (int idx = threadIdx.x; idx < KT * DH; idx += NTH) {
    const int r = idx / DH, i = idx - r * DH;
    const int s = s0 + r;
    sm[idx] = (s < S) ? ldv<T>(v.p, (long long)b * v.bs + (long long)s * v.rs + h * DH + i) : 0.f;
  }
}

// 16 B shared-memory reads: a key/value row slice of DHT floats is consumed as DHT/4 LDS.128 (every DHT is a multiple of 4)
template <int DHT>
__device__ __forceinline__ float dot4(const float (&q)[DHT], const float* __restrict__ k) {
  const float4* k4 = reinterpret_cast<const float4*>(k);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DHT / 4; ++i) {
    const float4 kk = k4[i];
    s = fmaf(q[4 * i], kk.x, s);
    s = fmaf(q[4 * i + 1], kk.y, s);
    s = fmaf(q[4 * i + 2], kk.z, s);
    s = fmaf(q[4 * i + 3], kk.w, s);
  }
  return s;
}
template <int DHT>
__device__ __forceinline__ void axpy4(float (&acc)[DHT], float p, const float* __restrict__ v) {
  const float4* v4 = reinterpret_cast<const float4*>(v);
#pragma unroll
  for (int i = 0; i < DHT / 4; ++i) {
    const float4 vv = v4[i];
    acc[4 * i] = fmaf(p, vv.x, acc[4 * i]);
    acc[4 * i + 1] = fmaf(p, vv.y, acc[4 * i + 1]);
    acc[4 * i + 2] = fmaf(p, vv.z, acc[4 * i + 2]);
    acc[4 * i + 3] = fmaf(p, vv.w, acc[4 * i + 3]);
  }
}

template <typename T, int DH, int TPQ>
__global__ void __launch_bounds__(NTH) attn_fwd_kernel(AttnView q, AttnView k, AttnView v, AttnViewW o, float* lse,
                                                      int H, int Sq, int Sk, float scale, DxDrop drop,
                                                      const unsigned long long* seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  constexpr int DHT = DH / TPQ;
  extern __shared__ __align__(16) float smem[];
  float* sk = smem;
  float* sv = smem + KT * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int qi = blockIdx.y * (NTH / TPQ) + threadIdx.x / TPQ;
  const int part = threadIdx.x % TPQ;
  const bool active = qi < Sq;
  float qr[DHT], acc[DHT];
#pragma unroll
  for (int i = 0; i < DHT; ++i) {
    qr[i] = active ? ldv<T>(q.p, (long long)b * q.bs + (long long)qi * q.rs + h * DH + part * DHT + i) * scale : 0.f;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const unsigned long long drow = ((unsigned long long)blockIdx.x * Sq + (unsigned)qi) * (unsigned long long)Sk;   // (b*H+h, q) row
  for (int k0 = 0; k0 < Sk; k0 += KT) {
    __syncthreads();
    load_tile<T, DH>(sk, k, b, h, k0, Sk);
    load_tile<T, DH>(sv, v, b, h, k0, Sk);
    __syncthreads();
    const int kn = min(KT, Sk - k0);
    for (int j = 0; j < kn; ++j) {
      const float* kj = sk + j * DH + part * DHT;
      float s = dot4<DHT>(qr, kj);
#pragma unroll
      for (int off = 1; off < TPQ; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (s > m) {
        const float corr = __expf(m - s);
        l *= corr;
#pragma unroll
        for (int i = 0; i < DHT; ++i) acc[i] *= corr;
        m = s;
      }
      const float p = __expf(s - m);
      l += p;   // softmax normaliser: before dropout
      const float pd = drop.thresh ? p * dx_drop_factor(drop, drow + (unsigned)(k0 + j)) : p;
      axpy4<DHT>(acc, pd, sv + j * DH + part * DHT);
    }
  }
  if (active) {
    const float inv = 1.f / l;
    T* op = reinterpret_cast<T*>(o.p) + (long long)b * o.bs + (long long)qi * o.rs + h * DH + part * DHT;
#pragma unroll
    for (int i = 0; i < DHT; ++i) dx_st(op + i, acc[i] * inv);
    if (part == 0 && lse) lse[((long long)b * H + h) * Sq + qi] = m + __logf(l);
  }
}

// Kernel A: per query -> D_i = <do_i, o_i>, dq_i = scale * sum_j p_ij (do_i.v_j - D_i) k_j
template <typename T, int DH, int TPQ>
__global__ void __launch_bounds__(NTH) attn_bwd_q_kernel(AttnView q, AttnView k, AttnView v, AttnView o, AttnView go,
                                                        AttnViewW dq, const float* lse, float* Dv, int H, int Sq, int Sk,
                                                        float scale, DxDrop drop, const unsigned long long* seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  constexpr int DHT = DH / TPQ;
  extern __shared__ __align__(16) float smem[];
  float* sk = smem;
  float* sv = smem + KT * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int qi = blockIdx.y * (NTH / TPQ) + threadIdx.x / TPQ;
  const int part = threadIdx.x % TPQ;
  const bool active = qi < Sq;
  float qr[DHT], gr[DHT], acc[DHT];
  float D = 0.f;
#pragma unroll
  for (int i = 0; i < DHT; ++i) {
    const long long hoff = h * DH + part * DHT + i;
    qr[i] = active ? ldv<T>(q.p, (long long)b * q.bs + (long long)qi * q.rs + hoff) * scale : 0.f;
    gr[i] = active ? ldv<T>(go.p, (long long)b * go.bs + (long long)qi * go.rs + hoff) : 0.f;
    const float ov = active ? ldv<T>(o.p, (long long)b * o.bs + (long long)qi * o.rs + hoff) : 0.f;
    D = fmaf(gr[i], ov, D);
    acc[i] = 0.f;
  }
#pragma unroll
  for (int off = 1; off < TPQ; off <<= 1) D += __shfl_xor_sync(0xffffffffu, D, off);
  const float L = active ? lse[((long long)b * H + h) * Sq + qi] : 0.f;
  const unsigned long long drow = ((unsigned long long)blockIdx.x * Sq + (unsigned)qi) * (unsigned long long)Sk;
  for (int k0 = 0; k0 < Sk; k0 += KT) {
    __syncthreads();
    load_tile<T, DH>(sk, k, b, h, k0, Sk);
    load_tile<T, DH>(sv, v, b, h, k0, Sk);
    __syncthreads();
    const int kn = min(KT, Sk - k0);
    for (int j = 0; j < kn; ++j) {
      const float* kj = sk + j * DH + part * DHT;
      const float* vj = sv + j * DH + part * DHT;
      float s = dot4<DHT>(qr, kj), dp = dot4<DHT>(gr, vj);
#pragma unroll
      for (int off = 1; off < TPQ; off <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        dp += __shfl_xor_sync(0xffffffffu, dp, off);
      }
      if (drop.thresh) dp *= dx_drop_factor(drop, drow + (unsigned)(k0 + j));   // dP = dP_dropped * mask / (1-p)
      const float ds = __expf(s - L) * (dp - D);
      axpy4<DHT>(acc, ds, kj);
    }
  }
  if (active) {
    T* dp_ = reinterpret_cast<T*>(dq.p) + (long long)b * dq.bs + (long long)qi * dq.rs + h * DH + part * DHT;
#pragma unroll
    for (int i = 0; i < DHT; ++i) dx_st(dp_ + i, acc[i] * scale);
    if (part == 0) Dv[((long long)b * H + h) * Sq + qi] = D;
  }
}

// Kernel B: per key -> dv_j = sum_i p_ij do_i ; dk_j = scale * sum_i p_ij (do_i.v_j - D_i) q_i
template <typename T, int DH, int TPQ>
__global__ void __launch_bounds__(NTH) attn_bwd_kv_kernel(AttnView q, AttnView k, AttnView v, AttnView go, AttnViewW dk,
                                                         AttnViewW dv, const float* lse, const float* Dv, int H, int Sq,
                                                         int Sk, float scale, DxDrop drop, const unsigned long long* seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  constexpr int DHT = DH / TPQ;
  extern __shared__ __align__(16) float smem[];
  float* sq = smem;
  float* sg = smem + KT * DH;
  float* sl = smem + 2 * KT * DH;  // lse tile
  float* sd = sl + KT;             // D tile
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int kj = blockIdx.y * (NTH / TPQ) + threadIdx.x / TPQ;
  const int part = threadIdx.x % TPQ;
  const bool active = kj < Sk;
  float kr[DHT], vr[DHT], ak[DHT], av[DHT];
#pragma unroll
  for (int i = 0; i < DHT; ++i) {
    const long long hoff = h * DH + part * DHT + i;
    kr[i] = active ? ldv<T>(k.p, (long long)b * k.bs + (long long)kj * k.rs + hoff) : 0.f;
    vr[i] = active ? ldv<T>(v.p, (long long)b * v.bs + (long long)kj * v.rs + hoff) : 0.f;
    ak[i] = 0.f;
    av[i] = 0.f;
  }
  for (int q0 = 0; q0 < Sq; q0 += KT) {
    __syncthreads();
    load_tile<T, DH>(sq, q, b, h, q0, Sq);
    load_tile<T, DH>(sg, go, b, h, q0, Sq);
    if (threadIdx.x < KT) {
      const int s = q0 + threadIdx.x;
      sl[threadIdx.x] = s < Sq ? lse[((long long)b * H + h) * Sq + s] : 0.f;
      sd[threadIdx.x] = s < Sq ? Dv[((long long)b * H + h) * Sq + s] : 0.f;
    }
    __syncthreads();
    const int qn = min(KT, Sq - q0);
    for (int i2 = 0; i2 < qn; ++i2) {
      const float* qi = sq + i2 * DH + part * DHT;
      const float* gi = sg + i2 * DH + part * DHT;
      float s = dot4<DHT>(kr, qi), dp = dot4<DHT>(vr, gi);
#pragma unroll
      for (int off = 1; off < TPQ; off <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        dp += __shfl_xor_sync(0xffffffffu, dp, off);
      }
      const float p = __expf(s * scale - sl[i2]);
      const float mk = drop.thresh
                           ? dx_drop_factor(drop, ((unsigned long long)blockIdx.x * Sq + (unsigned)(q0 + i2)) * (unsigned long long)Sk + (unsigned)kj)
                           : 1.f;
      const float ds = p * (dp * mk - sd[i2]);
      axpy4<DHT>(av, p * mk, gi);
      axpy4<DHT>(ak, ds, qi);
    }
  }
  if (active) {
    T* dkp = reinterpret_cast<T*>(dk.p) + (long long)b * dk.bs + (long long)kj * dk.rs + h * DH + part * DHT;
    T* dvp = reinterpret_cast<T*>(dv.p) + (long long)b * dv.bs + (long long)kj * dv.rs + h * DH + part * DHT;
#pragma unroll
    for (int i = 0; i < DHT; ++i) {
      dx_st(dkp + i, ak[i] * scale);
      dx_st(dvp + i, av[i]);
    }
  }
}

// rowdot[n] = <a[n,:], g[n,:]>;  g[n,:] *= row_scale[n]   (ScaleNorm-backward bookkeeping on the small [N,3d] tensors)
template <typename T>
__global__ void __launch_bounds__(256) rowdot_scale_kernel(const T* __restrict__ a, T* __restrict__ g,
                                                          const float* __restrict__ row_scale, float* __restrict__ rowdot,
                                                          int N, int C) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const T* ar = a + (long long)warp * C;
  T* gr = g + (long long)warp * C;
  const float s = row_scale ? row_scale[warp] : 1.f;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float gv = dx_ld(gr + c);
    dot = fmaf(dx_ld(ar + c), gv, dot);
    if (row_scale) dx_st(gr + c, gv * s);
  }
  dot = dx_warp_sum(dot);
  if (lane == 0 && rowdot) rowdot[warp] = dot;
}

// row_scale[n] = c * g / max(sqrt(rowsq[n]), eps)
__global__ void scalenorm_scale_kernel(const float* __restrict__ rowsq, const float* __restrict__ g, float c,
                                       float* __restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = c * g[0] / fmaxf(sqrtf(rowsq[i]), 1e-12f);
}

template <typename T, int DH, int TPQ>
int launch_fwd(const AttnView& q, const AttnView& k, const AttnView& v, const AttnViewW& o, float* lse, int B, int H,
               int Sq, int Sk, float scale, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  const size_t smem = 2 * KT * DH * sizeof(float);
  auto kern = attn_fwd_kernel<T, DH, TPQ>;
  if (smem > 48 * 1024) DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(B * H, dx_ceil_div(Sq, NTH / TPQ));
  kern<<<grid, NTH, smem, st>>>(q, k, v, o, lse, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

template <typename T, int DH, int TPQ>
int launch_bwd(const AttnView& q, const AttnView& k, const AttnView& v, const AttnView& o, const AttnView& go,
               const AttnViewW& dq, const AttnViewW& dk, const AttnViewW& dv, const float* lse, float* Dv, int B, int H,
               int Sq, int Sk, float scale, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  const size_t smem_q = 2 * KT * DH * sizeof(float);
  const size_t smem_kv = (2 * KT * DH + 2 * KT) * sizeof(float);
  auto kq = attn_bwd_q_kernel<T, DH, TPQ>;
  auto kkv = attn_bwd_kv_kernel<T, DH, TPQ>;
  if (smem_kv > 48 * 1024) {
    DX_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    DX_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
  }
  dim3 gq(B * H, dx_ceil_div(Sq, NTH / TPQ));
  kq<<<gq, NTH, smem_q, st>>>(q, k, v, o, go, dq, lse, Dv, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  dim3 gk(B * H, dx_ceil_div(Sk, NTH / TPQ));
  kkv<<<gk, NTH, smem_kv, st>>>(q, k, v, go, dk, dv, lse, Dv, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

// Head-averaged attention probabilities, the second return value of nn.MultiheadAttention(need_weights=True,
// average_attn_weights=True) in _PerceiverBlock.forward(return_attn=True) (models/main_architecture_duett.py:762,772):
//   out[b,i,j] = 1/H * sum_h exp(scale * <q[b,i,h,:], k[b,j,h,:]> - lse[b,h,i])
// with lse the row log-sum-exp dx_attn_fwd wrote.  Inference-time visualisation (7 pathology queries over 1369 patch or 24
// hour tokens): one CTA per (query, 128 keys), the query row in shared memory, one key per thread.
template <typename T>
__global__ void __launch_bounds__(NTH) attn_probs_mean_kernel(AttnView Q, AttnView K, const float* __restrict__ lse,
                                                             float* __restrict__ out, int H, int Sq, int Sk, int dh,
                                                             float scale) {
  extern __shared__ float sq[];   // H * dh
  const int b = blockIdx.z, i = blockIdx.y;
  for (int t = threadIdx.x; t < H * dh; t += NTH) sq[t] = ldv<T>(Q.p, (long long)b * Q.bs + (long long)i * Q.rs + t);
  __syncthreads();
  const int j = blockIdx.x * NTH + threadIdx.x;
  if (j >= Sk) return;
  const long long koff = (long long)b * K.bs + (long long)j * K.rs;
  float acc = 0.f;
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int e = 0; e < dh; ++e) s = fmaf(sq[h * dh + e], ldv<T>(K.p, koff + h * dh + e), s);
    acc += expf(s * scale - lse[((long long)b * H + h) * Sq + i]);
  }
  out[((long long)b * Sq + i) * Sk + j] = acc / (float)H;
}

#define DX_ATTN_DISPATCH(T, CALL)                                                     \
  switch (dh) {                                                                       \
    case 4: return CALL(T, 4, 1);                                                     \
    case 8: return CALL(T, 8, 1);                                                     \
    case 12: return CALL(T, 12, 1);                                                   \
    case 16: return CALL(T, 16, 1);                                                   \
    case 32: return CALL(T, 32, 1);                                                   \
    case 64: return CALL(T, 64, 2);                                                   \
    case 128: return CALL(T, 128, 4);                                                 \
    default:                                                                          \
      dx_set_error("dx_attn: unsupported head dim %d (supported 4,8,12,16,32,64,128)", dh); \
      return DX_ERR_UNSUPPORTED;                                                      \
  }

}  // namespace

extern "C" {

/* q/k/v/o: element (b,s,h,i) at ptr + b*bs + s*rs + h*dh + i (strides in elements). */
int dx_attn_fwd(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const void* v,
                int64_t v_bs, int64_t v_rs, void* o, int64_t o_bs, int64_t o_rs, float* lse, int B, int H, int Sq, int Sk,
                int dh, int dtype, float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_dev, void* stream) {
  DX_CHECK_ARG(q && k && v && o, "dx_attn_fwd: null tensor");
  DX_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "dx_attn_fwd: dropout probability must be in [0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  const DxDrop drop = dx_make_drop(drop_p, drop_seed);
  const unsigned long long* seed_dev = reinterpret_cast<const unsigned long long*>(drop_seed_dev);
  if (dtype == DX_BF16 && !attn_force_simt()) {
    const void* ptrs[4] = {q, k, v, o};
    const long long bs[4] = {q_bs, k_bs, v_bs, o_bs}, rs[4] = {q_rs, k_rs, v_rs, o_rs};
    if (dx_attn_tc_supported(ptrs, bs, rs, 4, Sq, Sk, dh))
      return dx_attn_tc_fwd(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, B, H, Sq, Sk, dh, drop, seed_dev, st);
    if (dx_attn_mma_supported(ptrs, bs, rs, 4, Sq, Sk, dh))
      return dx_attn_mma_fwd(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, B, H, Sq, Sk, dh, drop, seed_dev, st);
    // long sequences / dh = 128 (stress shape): tiled tensor-core kernels; few queries (perceiver latents) stay on the SIMT path
    if (Sq >= 16 && lse && dx_attn_mmat_supported(ptrs, bs, rs, 4, Sq, Sk, dh))
      return dx_attn_mmat_fwd(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, lse, B, H, Sq, Sk, dh, drop, seed_dev, st);
  }
  AttnView Q{q, q_bs, q_rs}, K{k, k_bs, k_rs}, V{v, v_bs, v_rs};
  AttnViewW O{o, o_bs, o_rs};
  const float scale = 1.f / sqrtf((float)dh);
#define CALL_FWD(T, DH, TPQ) launch_fwd<T, DH, TPQ>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st)
  if (dtype == DX_BF16) { DX_ATTN_DISPATCH(bf16, CALL_FWD) } else { DX_ATTN_DISPATCH(float, CALL_FWD) }
#undef CALL_FWD
}

int dx_attn_bwd(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const void* v,
                int64_t v_bs, int64_t v_rs, const void* o, int64_t o_bs, int64_t o_rs, const void* go, int64_t go_bs,
                int64_t go_rs, void* dq, int64_t dq_bs, int64_t dq_rs, void* dk, int64_t dk_bs, int64_t dk_rs, void* dv,
                int64_t dv_bs, int64_t dv_rs, const float* lse, float* D_ws, int B, int H, int Sq, int Sk, int dh,
                int dtype, float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_dev, void* stream) {
  DX_CHECK_ARG(q && k && v && o && go && dq && dk && dv && lse && D_ws, "dx_attn_bwd: null tensor");
  DX_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "dx_attn_bwd: dropout probability must be in [0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  const DxDrop drop = dx_make_drop(drop_p, drop_seed);
  const unsigned long long* seed_dev = reinterpret_cast<const unsigned long long*>(drop_seed_dev);
  if (dtype == DX_BF16 && !attn_force_simt()) {
    const void* ptrs[8] = {q, k, v, o, go, dq, dk, dv};
    const long long bs[8] = {q_bs, k_bs, v_bs, o_bs, go_bs, dq_bs, dk_bs, dv_bs};
    const long long rs[8] = {q_rs, k_rs, v_rs, o_rs, go_rs, dq_rs, dk_rs, dv_rs};
    if (dx_attn_mma_supported(ptrs, bs, rs, 8, Sq, Sk, dh))
      return dx_attn_mma_bwd(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, go, go_bs, go_rs, dq, dq_bs, dq_rs, dk,
                             dk_bs, dk_rs, dv, dv_bs, dv_rs, lse, B, H, Sq, Sk, dh, drop, seed_dev, st);
    if (Sq >= 16 && dx_attn_mmat_supported(ptrs, bs, rs, 8, Sq, Sk, dh))
      return dx_attn_mmat_bwd(q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, go, go_bs, go_rs, dq, dq_bs, dq_rs, dk,
                              dk_bs, dk_rs, dv, dv_bs, dv_rs, lse, D_ws, B, H, Sq, Sk, dh, drop, seed_dev, st);
  }
  AttnView Q{q, q_bs, q_rs}, K{k, k_bs, k_rs}, V{v, v_bs, v_rs}, O{o, o_bs, o_rs}, GO{go, go_bs, go_rs};
  AttnViewW DQ{dq, dq_bs, dq_rs}, DK{dk, dk_bs, dk_rs}, DV{dv, dv_bs, dv_rs};
  const float scale = 1.f / sqrtf((float)dh);
#define CALL_BWD(T, DH, TPQ) launch_bwd<T, DH, TPQ>(Q, K, V, O, GO, DQ, DK, DV, lse, D_ws, B, H, Sq, Sk, scale, drop, seed_dev, st)
  if (dtype == DX_BF16) { DX_ATTN_DISPATCH(bf16, CALL_BWD) } else { DX_ATTN_DISPATCH(float, CALL_BWD) }
#undef CALL_BWD
}

int dx_attn_probs_mean(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const float* lse,
                       float* out, int B, int H, int Sq, int Sk, int dh, int dtype, void* stream) {
  DX_CHECK_ARG(q && k && lse && out, "dx_attn_probs_mean: null tensor");
  DX_CHECK_ARG(B > 0 && H > 0 && Sq > 0 && Sk > 0 && dh > 0, "dx_attn_probs_mean: empty problem");
  DX_CHECK_ARG(B <= 65535 && Sq <= 65535, "dx_attn_probs_mean: B and Sq must be <= 65535");
  DX_CHECK_ARG((size_t)H * dh * sizeof(float) <= 48 * 1024, "dx_attn_probs_mean: H*dh must be <= 12288");
  cudaStream_t st = (cudaStream_t)stream;
  AttnView Q{q, q_bs, q_rs}, K{k, k_bs, k_rs};
  const float scale = 1.f / sqrtf((float)dh);
  const dim3 grid(dx_ceil_div(Sk, NTH), Sq, B);
  const size_t smem = (size_t)H * dh * sizeof(float);
  if (dtype == DX_BF16) attn_probs_mean_kernel<bf16><<<grid, NTH, smem, st>>>(Q, K, lse, out, H, Sq, Sk, dh, scale);
  else attn_probs_mean_kernel<float><<<grid, NTH, smem, st>>>(Q, K, lse, out, H, Sq, Sk, dh, scale);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_rowdot_scale(const void* a, void* g, const float* row_scale, float* rowdot, int N, int C, int dtype, void* stream) {
  DX_CHECK_ARG(a && g, "dx_rowdot_scale: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = dx_ceil_div((long long)N * 32, 256);
  if (dtype == DX_BF16) rowdot_scale_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)a, (bf16*)g, row_scale, rowdot, N, C);
  else rowdot_scale_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (float*)g, row_scale, rowdot, N, C);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

int dx_scalenorm_scale(const float* rowsq, const float* g, float c, float* out, int N, void* stream) {
  DX_CHECK_ARG(rowsq && g && out, "dx_scalenorm_scale: null tensor");
  scalenorm_scale_kernel<<<dx_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(rowsq, g, c, out, N);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
