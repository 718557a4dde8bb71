// Tensor-core self-attention for the DuETT axis encoders in bf16 (duett/duett.py:95-105: x_transformers Attention core,
// softmax(q k^T / sqrt(dh)) v, softmax in fp32).  The sequences are tiny (S = V+1 = 129 event tokens or T+1 = 33 time
// tokens, head dim d/heads = 32..64, B*heads = 512 problems), so one CTA owns one (sample, head) problem entirely:
// Q, K, V (and dO) live in shared memory, every warp owns 16 query rows (forward, dQ) or 16 key rows (dK, dV), and the
// 16-row x 8/16-column products run on warp-level mma.sync.m16n8k16 bf16 -> fp32.  The whole step spends < 0.3 % of its
// FLOPs here; a 128-row tcgen05 tile would be 87 % padding at S = 129/33, which is why this kernel uses the
// 16-row warp MMA instead of the UMMA path of dx_gemm_tc.cu.  The SIMT kernels in dx_attention.cu remain the fp32 path
// and the path for long key sets (perceiver cross-attention).
//
// forward : streaming over 16-key blocks with an online softmax; writes o and the row log-sum-exp (natural log).
// backward: D = <dO, o> per query, then  dQ = scale * dS K   (warps own queries)
//                                        dV = P^T dO, dK = scale * dS^T Q   (warps own keys),  dS = P o (dP - D).
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace dx_attn_mma {

constexpr int PAD = 8;          // bf16 elements of row padding: row pitch (DH+8)*2 B keeps ldmatrix conflict-free
constexpr int MAX_S = 144;      // 9 row blocks of 16
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct View {      // element (b, s, head h, i) at p + b*bs + s*rs + h*dh + i  (strides in elements)
  const bf16* p;
  long long bs, rs;
};
struct ViewW {
  bf16* p;
  long long bs, rs;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// rows [0, SP) of a [S, DH] head slice -> shared memory (pitch DH+PAD), zero-filled past S; 16 B per thread per step
template <int DH>
__device__ __forceinline__ void load_rows(bf16* sm, const View& v, int b, int h, int S, int SP) {
  constexpr int CPR = DH / 8;   // 16 B chunks per row
  const bf16* base = v.p + (long long)b * v.bs + (long long)h * DH;
  for (int idx = threadIdx.x; idx < SP * CPR; idx += blockDim.x) {
    const int r = idx / CPR, c = idx - r * CPR;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r < S) u = *reinterpret_cast<const uint4*>(base + (long long)r * v.rs + c * 8);
    *reinterpret_cast<uint4*>(sm + r * (DH + PAD) + c * 8) = u;
  }
}

// A fragments (16 rows x DH) of the row block starting at m0
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t sbase, int m0, int lane, uint32_t (&a)[DH / 16][4]) {
  const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) ldsm_x4(sbase + (uint32_t)((row * (DH + PAD) + kk * 16 + (lane >> 4) * 8) * 2), a[kk]);
}

// c[2][4] (16 x 16 block) = A(16 x DH) . X[n0..n0+16, :]^T   with X row-major [n][k] in shared memory
template <int DH>
__device__ __forceinline__ void mma_a_xt(float (&c)[2][4], const uint32_t (&a)[DH / 16][4], uint32_t sbase, int n0, int lane) {
  const int row = n0 + (lane & 7) + (lane >> 4) * 8;
  const int col = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    uint32_t r[4];
    ldsm_x4(sbase + (uint32_t)((row * (DH + PAD) + kk * 16 + col) * 2), r);
    mma16816(c[0], a[kk], r[0], r[1]);
    mma16816(c[1], a[kk], r[2], r[3]);
  }
}

// acc[DH/8][4] (16 x DH) += P(16 x 16, A fragment pa) . X[k0..k0+16, :]   with X row-major [k][n] in shared memory
template <int DH>
__device__ __forceinline__ void mma_p_x(float (&acc)[DH / 8][4], const uint32_t (&pa)[4], uint32_t sbase, int k0, int lane) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int col = (lane >> 4) * 8;
#pragma unroll
  for (int nt = 0; nt < DH / 16; ++nt) {
    uint32_t r[4];
    ldsm_x4_t(sbase + (uint32_t)((row * (DH + PAD) + nt * 16 + col) * 2), r);
    mma16816(acc[2 * nt], pa, r[0], r[1]);
    mma16816(acc[2 * nt + 1], pa, r[2], r[3]);
  }
}

// fragment (16 x DH fp32 accumulators, rows m0+g / m0+g+8) -> global bf16 rows, scaled
template <int DH>
__device__ __forceinline__ void store_frag(const ViewW& o, int b, int h, int m0, int S, int lane, const float (&acc)[DH / 8][4],
                                           float s0, float s1) {
  const int g = lane >> 2, t = lane & 3;
  bf16* base = o.p + (long long)b * o.bs + (long long)h * DH + 2 * t;
  const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
    if (r0 < S) *reinterpret_cast<uint32_t*>(base + (long long)r0 * o.rs + nt * 8) = pack2(acc[nt][0] * s0, acc[nt][1] * s0);
    if (r1 < S) *reinterpret_cast<uint32_t*>(base + (long long)r1 * o.rs + nt * 8) = pack2(acc[nt][2] * s1, acc[nt][3] * s1);
  }
}

template <int DH>
__global__ void __launch_bounds__(288) attn_mma_fwd_kernel(View q, View k, View v, ViewW o, float* __restrict__ lse, int H,
                                                          int Sq, int Sk, float scale, DxDrop drop,
                                                          const unsigned long long* __restrict__ seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int SPq = (Sq + 15) & ~15, SPk = (Sk + 15) & ~15;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + SPq * (DH + PAD);
  bf16* sV = sK + SPk * (DH + PAD);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  load_rows<DH>(sQ, q, b, h, Sq, SPq);
  load_rows<DH>(sK, k, b, h, Sk, SPk);
  load_rows<DH>(sV, v, b, h, Sk, SPk);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const int m0 = warp * 16;
  if (m0 >= SPq) return;
  const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);
  uint32_t qa[DH / 16][4];
  load_a_frags<DH>(uQ, m0, lane, qa);
  float acc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, ls[2] = {0.f, 0.f};   // running row max (log2 domain) / thread-partial row sums
  const float sc2 = scale * LOG2E;
  const unsigned long long dbase = (unsigned long long)blockIdx.x * Sq;   // (b*H + h) * Sq
  for (int k0 = 0; k0 < SPk; k0 += 16) {
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_a_xt<DH>(s, qa, uK, k0, lane);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + nt * 8 + 2 * t + (e & 1);
        s[nt][e] = key < Sk ? s[nt][e] * sc2 : -INFINITY;
      }
    uint32_t pa[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {   // r = 0: row g (elements 0,1), r = 1: row g+8 (elements 2,3)
      float m = fmaxf(fmaxf(s[0][2 * r], s[0][2 * r + 1]), fmaxf(s[1][2 * r], s[1][2 * r + 1]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      const float mnew = fmaxf(mx[r], m);       // finite: key block 0 always holds a valid key
      const float corr = ex2(mx[r] - mnew);
      mx[r] = mnew;
      float p00 = ex2(s[0][2 * r] - mnew), p01 = ex2(s[0][2 * r + 1] - mnew);
      float p10 = ex2(s[1][2 * r] - mnew), p11 = ex2(s[1][2 * r + 1] - mnew);
      ls[r] = ls[r] * corr + (p00 + p01) + (p10 + p11);   // softmax normaliser: before dropout
      if (drop.thresh) {
        const unsigned long long di = (dbase + (unsigned)(m0 + (lane >> 2) + 8 * r)) * (unsigned long long)Sk + (unsigned)(k0 + 2 * t);
        p00 *= dx_drop_factor(drop, di);
        p01 *= dx_drop_factor(drop, di + 1);
        p10 *= dx_drop_factor(drop, di + 8);
        p11 *= dx_drop_factor(drop, di + 9);
      }
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        acc[i][2 * r] *= corr;
        acc[i][2 * r + 1] *= corr;
      }
      pa[r] = pack2(p00, p01);          // a0 (row g, keys 2t..) / a1 (row g+8)
      pa[2 + r] = pack2(p10, p11);      // a2 (row g, keys 8+2t..) / a3
    }
    mma_p_x<DH>(acc, pa, uV, k0, lane);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 1);
    ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 2);
  }
  store_frag<DH>(o, b, h, m0, Sq, lane, acc, 1.f / ls[0], 1.f / ls[1]);
  if (lse && t == 0) {
    const int g = lane >> 2;
    float* lp = lse + ((long long)b * H + h) * Sq;
    if (m0 + g < Sq) lp[m0 + g] = (mx[0] + log2f(ls[0])) * LN2;
    if (m0 + g + 8 < Sq) lp[m0 + g + 8] = (mx[1] + log2f(ls[1])) * LN2;
  }
}

template <int DH>
__global__ void __launch_bounds__(288) attn_mma_bwd_kernel(View q, View k, View v, View o, View go, ViewW dq, ViewW dk, ViewW dv,
                                                          const float* __restrict__ lse, int H, int Sq, int Sk, float scale,
                                                          DxDrop drop, const unsigned long long* __restrict__ seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  const unsigned long long dbase = (unsigned long long)blockIdx.x * Sq;   // (b*H + h) * Sq
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int SPq = (Sq + 15) & ~15, SPk = (Sk + 15) & ~15;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sG = sQ + SPq * (DH + PAD);
  bf16* sK = sG + SPq * (DH + PAD);
  bf16* sV = sK + SPk * (DH + PAD);
  float* sL = reinterpret_cast<float*>(sV + SPk * (DH + PAD));   // lse * log2(e)
  float* sD = sL + SPq;                                          // D = <dO, o>
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nwarps = blockDim.x >> 5;
  load_rows<DH>(sQ, q, b, h, Sq, SPq);
  load_rows<DH>(sG, go, b, h, Sq, SPq);
  load_rows<DH>(sK, k, b, h, Sk, SPk);
  load_rows<DH>(sV, v, b, h, Sk, SPk);
  for (int i = threadIdx.x; i < SPq; i += blockDim.x) sL[i] = i < Sq ? lse[((long long)b * H + h) * Sq + i] * LOG2E : 0.f;
  __syncthreads();
  // D: one warp per query row, dO from shared memory, o from global (2 bf16 per lane and step)
  for (int r = warp; r < SPq; r += nwarps) {
    float d = 0.f;
    if (r < Sq) {
      const bf16* op = o.p + (long long)b * o.bs + (long long)r * o.rs + (long long)h * DH;
      for (int c = 2 * lane; c < DH; c += 64) {
        const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(op + c));
        const float2 gv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sG + r * (DH + PAD) + c));
        d = fmaf(ov.x, gv.x, fmaf(ov.y, gv.y, d));
      }
    }
    d = dx_warp_sum(d);
    if (lane == 0) sD[r] = d;
  }
  __syncthreads();
  const uint32_t uQ = smem_u32(sQ), uG = smem_u32(sG), uK = smem_u32(sK), uV = smem_u32(sV);
  const float sc2 = scale * LOG2E;
  const int m0 = warp * 16;
  // ---- dQ: this warp owns query rows [m0, m0+16) ----------------------------------------------------------------------
  if (m0 < SPq) {
    uint32_t qa[DH / 16][4], ga[DH / 16][4];
    load_a_frags<DH>(uQ, m0, lane, qa);
    load_a_frags<DH>(uG, m0, lane, ga);
    const float L[2] = {sL[m0 + g], sL[m0 + g + 8]}, Dr[2] = {sD[m0 + g], sD[m0 + g + 8]};
    float acc[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int k0 = 0; k0 < SPk; k0 += 16) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_a_xt<DH>(s, qa, uK, k0, lane);
      mma_a_xt<DH>(dp, ga, uV, k0, lane);
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = k0 + nt * 8 + 2 * t + (e & 1);
          const float p = key < Sk ? ex2(s[nt][e] * sc2 - L[e >> 1]) : 0.f;
          float dpe = dp[nt][e];
          if (drop.thresh)
            dpe *= dx_drop_factor(drop, (dbase + (unsigned)(m0 + g + 8 * (e >> 1))) * (unsigned long long)Sk + (unsigned)key);
          ds[nt][e] = p * (dpe - Dr[e >> 1]);
        }
      const uint32_t dsa[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                               pack2(ds[1][2], ds[1][3])};
      mma_p_x<DH>(acc, dsa, uK, k0, lane);
    }
    store_frag<DH>(dq, b, h, m0, Sq, lane, acc, scale, scale);
  }
  // ---- dK, dV: this warp owns key rows [m0, m0+16) ---------------------------------------------------------------------
  if (m0 < SPk) {
    uint32_t ka[DH / 16][4], va[DH / 16][4];
    load_a_frags<DH>(uK, m0, lane, ka);
    load_a_frags<DH>(uV, m0, lane, va);
    float ak[DH / 8][4], av[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      ak[i][0] = ak[i][1] = ak[i][2] = ak[i][3] = 0.f;
      av[i][0] = av[i][1] = av[i][2] = av[i][3] = 0.f;
    }
    for (int q0 = 0; q0 < SPq; q0 += 16) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_a_xt<DH>(s, ka, uQ, q0, lane);     // S^T block: rows = keys, columns = queries
      mma_a_xt<DH>(dp, va, uG, q0, lane);    // dP^T block
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = q0 + nt * 8 + 2 * t + (e & 1);
          const float pe = qi < Sq ? ex2(s[nt][e] * sc2 - sL[qi]) : 0.f;
          const float mk = drop.thresh
                               ? dx_drop_factor(drop, (dbase + (unsigned)qi) * (unsigned long long)Sk + (unsigned)(m0 + g + 8 * (e >> 1)))
                               : 1.f;
          ds[nt][e] = pe * (dp[nt][e] * mk - sD[qi]);
          p[nt][e] = pe * mk;   // dV uses the dropped probabilities
        }
      const uint32_t pa[4] = {pack2(p[0][0], p[0][1]), pack2(p[0][2], p[0][3]), pack2(p[1][0], p[1][1]), pack2(p[1][2], p[1][3])};
      const uint32_t dsa[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                               pack2(ds[1][2], ds[1][3])};
      mma_p_x<DH>(av, pa, uG, q0, lane);     // dV += P^T dO
      mma_p_x<DH>(ak, dsa, uQ, q0, lane);    // dK += dS^T Q
    }
    store_frag<DH>(dk, b, h, m0, Sk, lane, ak, scale, scale);
    store_frag<DH>(dv, b, h, m0, Sk, lane, av, 1.f, 1.f);
  }
}

// =========================================================================================================================
// Tiled variants: any sequence length and head dims up to 128 (the stress shape of BASELINE.json configs[4]: S = 513 / 129,
// dh = 128, where the one-CTA-per-(sample, head) kernels above do not fit in shared memory and the SIMT fallback took 31 % of
// the step).  Same warp-level mma.sync.m16n8k16 fragments and helpers; a CTA owns a 128-row tile of queries (forward, dQ) or
// of keys (dK, dV) and streams the other operand through shared memory in 64-row tiles.  A fragments are re-read from shared
// memory with ldmatrix inside the loops (keeping Q/dO/K/V fragments AND two 16 x 128 accumulators in registers would spill).
// The backward is two launches (FlashAttention-2 style): the dQ kernel also produces D = <dO, o> for the dK/dV kernel.
// =========================================================================================================================
constexpr int TQ = 128;   // resident rows per CTA (8 warps x 16)
constexpr int TS = 64;    // streamed rows per step

// rows [r0, r0 + nrows) of a head slice -> shared rows [0, nrows), zero-filled past S
template <int DH>
__device__ __forceinline__ void load_rows_off(bf16* sm, const View& v, int b, int h, int r0, int nrows, int S) {
  constexpr int CPR = DH / 8;
  const bf16* base = v.p + (long long)b * v.bs + (long long)h * DH;
  for (int idx = threadIdx.x; idx < nrows * CPR; idx += blockDim.x) {
    const int r = idx / CPR, c = idx - r * CPR;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r0 + r < S) u = *reinterpret_cast<const uint4*>(base + (long long)(r0 + r) * v.rs + c * 8);
    *reinterpret_cast<uint4*>(sm + r * (DH + PAD) + c * 8) = u;
  }
}

// c[2][4] (16 x 16) = A[m0..m0+16, :] . X[n0..n0+16, :]^T, both row-major [row][k] in shared memory (A fragments loaded per step)
template <int DH>
__device__ __forceinline__ void mma_xa_xt(float (&c)[2][4], uint32_t abase, int m0, uint32_t xbase, int n0, int lane) {
  const int arow = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int xrow = n0 + (lane & 7) + (lane >> 4) * 8;
  const int xcol = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    uint32_t a[4], r[4];
    ldsm_x4(abase + (uint32_t)((arow * (DH + PAD) + kk * 16 + (lane >> 4) * 8) * 2), a);
    ldsm_x4(xbase + (uint32_t)((xrow * (DH + PAD) + kk * 16 + xcol) * 2), r);
    mma16816(c[0], a, r[0], r[1]);
    mma16816(c[1], a, r[2], r[3]);
  }
}

template <int DH>
__global__ void __launch_bounds__(256) attn_mmat_fwd_kernel(View q, View k, View v, ViewW o, float* __restrict__ lse, int H, int Sq,
                                                           int Sk, float scale, DxDrop drop,
                                                           const unsigned long long* __restrict__ seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + TQ * (DH + PAD);
  bf16* sV = sK + TS * (DH + PAD);
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int q0 = blockIdx.y * TQ;
  load_rows_off<DH>(sQ, q, b, h, q0, TQ, Sq);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const int m0 = warp * 16;                          // local row of this warp inside the tile
  const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);
  float acc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float mx[2] = {-INFINITY, -INFINITY}, ls[2] = {0.f, 0.f};
  const float sc2 = scale * LOG2E;
  const unsigned long long dbase = (unsigned long long)blockIdx.x * Sq;   // (b*H + h) * Sq
  for (int kt = 0; kt < Sk; kt += TS) {
    __syncthreads();                                  // the previous tile is no longer read
    load_rows_off<DH>(sK, k, b, h, kt, TS, Sk);
    load_rows_off<DH>(sV, v, b, h, kt, TS, Sk);
    __syncthreads();
#pragma unroll 1
    for (int k0 = 0; k0 < TS; k0 += 16) {
      if (kt + k0 >= Sk) break;                       // uniform per CTA
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_xa_xt<DH>(s, uQ, m0, uK, k0, lane);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kt + k0 + nt * 8 + 2 * t + (e & 1);
          s[nt][e] = key < Sk ? s[nt][e] * sc2 : -INFINITY;
        }
      uint32_t pa[4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float m = fmaxf(fmaxf(s[0][2 * r], s[0][2 * r + 1]), fmaxf(s[1][2 * r], s[1][2 * r + 1]));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        const float mnew = fmaxf(mx[r], m);           // finite: the first key block always holds a valid key
        const float corr = ex2(mx[r] - mnew);
        mx[r] = mnew;
        float p00 = ex2(s[0][2 * r] - mnew), p01 = ex2(s[0][2 * r + 1] - mnew);
        float p10 = ex2(s[1][2 * r] - mnew), p11 = ex2(s[1][2 * r + 1] - mnew);
        ls[r] = ls[r] * corr + (p00 + p01) + (p10 + p11);
        if (drop.thresh) {
          const unsigned long long di =
              (dbase + (unsigned)(q0 + m0 + (lane >> 2) + 8 * r)) * (unsigned long long)Sk + (unsigned)(kt + k0 + 2 * t);
          p00 *= dx_drop_factor(drop, di);
          p01 *= dx_drop_factor(drop, di + 1);
          p10 *= dx_drop_factor(drop, di + 8);
          p11 *= dx_drop_factor(drop, di + 9);
        }
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) {
          acc[i][2 * r] *= corr;
          acc[i][2 * r + 1] *= corr;
        }
        pa[r] = pack2(p00, p01);
        pa[2 + r] = pack2(p10, p11);
      }
      mma_p_x<DH>(acc, pa, uV, k0, lane);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 1);
    ls[r] += __shfl_xor_sync(0xffffffffu, ls[r], 2);
  }
  store_frag<DH>(o, b, h, q0 + m0, Sq, lane, acc, 1.f / ls[0], 1.f / ls[1]);
  if (lse && t == 0) {
    const int g = lane >> 2;
    float* lp = lse + ((long long)b * H + h) * Sq;
    if (q0 + m0 + g < Sq) lp[q0 + m0 + g] = (mx[0] + log2f(ls[0])) * LN2;
    if (q0 + m0 + g + 8 < Sq) lp[q0 + m0 + g + 8] = (mx[1] + log2f(ls[1])) * LN2;
  }
}

// dQ for a 128-row query tile; also writes D[b,h,row] = <dO[row], o[row]> for the dK/dV kernel
template <int DH>
__global__ void __launch_bounds__(256) attn_mmat_bwd_dq_kernel(View q, View k, View v, View o, View go, ViewW dq,
                                                              const float* __restrict__ lse, float* __restrict__ Dws, int H, int Sq,
                                                              int Sk, float scale, DxDrop drop,
                                                              const unsigned long long* __restrict__ seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  const unsigned long long dbase = (unsigned long long)blockIdx.x * Sq;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sG = sQ + TQ * (DH + PAD);
  bf16* sK = sG + TQ * (DH + PAD);
  bf16* sV = sK + TS * (DH + PAD);
  float* sL = reinterpret_cast<float*>(sV + TS * (DH + PAD));
  float* sD = sL + TQ;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int q0 = blockIdx.y * TQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  load_rows_off<DH>(sQ, q, b, h, q0, TQ, Sq);
  load_rows_off<DH>(sG, go, b, h, q0, TQ, Sq);
  for (int i = threadIdx.x; i < TQ; i += blockDim.x) sL[i] = q0 + i < Sq ? lse[((long long)b * H + h) * Sq + q0 + i] * LOG2E : 0.f;
  __syncthreads();
  for (int r = warp; r < TQ; r += 8) {               // D: one warp per query row, dO from shared memory, o from global
    float d = 0.f;
    if (q0 + r < Sq) {
      const bf16* op = o.p + (long long)b * o.bs + (long long)(q0 + r) * o.rs + (long long)h * DH;
      for (int c = 2 * lane; c < DH; c += 64) {
        const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(op + c));
        const float2 gv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sG + r * (DH + PAD) + c));
        d = fmaf(ov.x, gv.x, fmaf(ov.y, gv.y, d));
      }
    }
    d = dx_warp_sum(d);
    if (lane == 0) {
      sD[r] = d;
      if (q0 + r < Sq) Dws[((long long)b * H + h) * Sq + q0 + r] = d;
    }
  }
  __syncthreads();
  const uint32_t uQ = smem_u32(sQ), uG = smem_u32(sG), uK = smem_u32(sK), uV = smem_u32(sV);
  const float sc2 = scale * LOG2E;
  const int m0 = warp * 16;
  const float L[2] = {sL[m0 + g], sL[m0 + g + 8]}, Dr[2] = {sD[m0 + g], sD[m0 + g + 8]};
  float acc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int kt = 0; kt < Sk; kt += TS) {
    __syncthreads();
    load_rows_off<DH>(sK, k, b, h, kt, TS, Sk);
    load_rows_off<DH>(sV, v, b, h, kt, TS, Sk);
    __syncthreads();
#pragma unroll 1
    for (int k0 = 0; k0 < TS; k0 += 16) {
      if (kt + k0 >= Sk) break;
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_xa_xt<DH>(s, uQ, m0, uK, k0, lane);
      mma_xa_xt<DH>(dp, uG, m0, uV, k0, lane);
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kt + k0 + nt * 8 + 2 * t + (e & 1);
          const float p = key < Sk ? ex2(s[nt][e] * sc2 - L[e >> 1]) : 0.f;
          float dpe = dp[nt][e];
          if (drop.thresh)
            dpe *= dx_drop_factor(drop, (dbase + (unsigned)(q0 + m0 + g + 8 * (e >> 1))) * (unsigned long long)Sk + (unsigned)key);
          ds[nt][e] = p * (dpe - Dr[e >> 1]);
        }
      const uint32_t dsa[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                               pack2(ds[1][2], ds[1][3])};
      mma_p_x<DH>(acc, dsa, uK, k0, lane);
    }
  }
  store_frag<DH>(dq, b, h, q0 + m0, Sq, lane, acc, scale, scale);
}

// dK, dV for a 128-row key tile; queries (Q, dO, lse, D) streamed in 64-row tiles
template <int DH>
__global__ void __launch_bounds__(256) attn_mmat_bwd_kv_kernel(View q, View k, View v, View go, ViewW dk, ViewW dv,
                                                              const float* __restrict__ lse, const float* __restrict__ Dws, int H,
                                                              int Sq, int Sk, float scale, DxDrop drop,
                                                              const unsigned long long* __restrict__ seed_dev) {
  drop = dx_drop_resolve(drop, seed_dev);
  const unsigned long long dbase = (unsigned long long)blockIdx.x * Sq;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + TQ * (DH + PAD);
  bf16* sQ = sV + TQ * (DH + PAD);
  bf16* sG = sQ + TS * (DH + PAD);
  float* sL = reinterpret_cast<float*>(sG + TS * (DH + PAD));
  float* sD = sL + TS;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int kbase = blockIdx.y * TQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  load_rows_off<DH>(sK, k, b, h, kbase, TQ, Sk);
  load_rows_off<DH>(sV, v, b, h, kbase, TQ, Sk);
  const uint32_t uQ = smem_u32(sQ), uG = smem_u32(sG), uK = smem_u32(sK), uV = smem_u32(sV);
  const float sc2 = scale * LOG2E;
  const int m0 = warp * 16;                          // local key row of this warp
  float ak[DH / 8][4], av[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    ak[i][0] = ak[i][1] = ak[i][2] = ak[i][3] = 0.f;
    av[i][0] = av[i][1] = av[i][2] = av[i][3] = 0.f;
  }
  for (int qt = 0; qt < Sq; qt += TS) {
    __syncthreads();
    load_rows_off<DH>(sQ, q, b, h, qt, TS, Sq);
    load_rows_off<DH>(sG, go, b, h, qt, TS, Sq);
    for (int i = threadIdx.x; i < TS; i += blockDim.x) {
      const bool ok = qt + i < Sq;
      sL[i] = ok ? lse[((long long)b * H + h) * Sq + qt + i] * LOG2E : 0.f;
      sD[i] = ok ? Dws[((long long)b * H + h) * Sq + qt + i] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int qq = 0; qq < TS; qq += 16) {
      if (qt + qq >= Sq) break;
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_xa_xt<DH>(s, uK, m0, uQ, qq, lane);      // S^T block: rows = keys, columns = queries
      mma_xa_xt<DH>(dp, uV, m0, uG, qq, lane);     // dP^T block
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ql = qq + nt * 8 + 2 * t + (e & 1);      // local query index in the streamed tile
          const int qi = qt + ql;
          const float pe = qi < Sq ? ex2(s[nt][e] * sc2 - sL[ql]) : 0.f;
          const float mk = drop.thresh ? dx_drop_factor(drop, (dbase + (unsigned)qi) * (unsigned long long)Sk +
                                                                  (unsigned)(kbase + m0 + g + 8 * (e >> 1)))
                                       : 1.f;
          ds[nt][e] = pe * (dp[nt][e] * mk - sD[ql]);
          p[nt][e] = pe * mk;
        }
      const uint32_t pa[4] = {pack2(p[0][0], p[0][1]), pack2(p[0][2], p[0][3]), pack2(p[1][0], p[1][1]), pack2(p[1][2], p[1][3])};
      const uint32_t dsa[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                               pack2(ds[1][2], ds[1][3])};
      mma_p_x<DH>(av, pa, uG, qq, lane);     // dV += P^T dO
      mma_p_x<DH>(ak, dsa, uQ, qq, lane);    // dK += dS^T Q
    }
  }
  store_frag<DH>(dk, b, h, kbase + m0, Sk, lane, ak, scale, scale);
  store_frag<DH>(dv, b, h, kbase + m0, Sk, lane, av, 1.f, 1.f);
}

template <int DH>
int launch_fwd_tiled(const View& q, const View& k, const View& v, const ViewW& o, float* lse, int B, int H, int Sq, int Sk,
                     float scale, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  const size_t smem = (size_t)(TQ + 2 * TS) * (DH + PAD) * 2;
  auto kern = attn_mmat_fwd_kernel<DH>;
  static bool attr = false;
  if (!attr) {
    DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(B * H, (Sq + TQ - 1) / TQ);
  kern<<<grid, 256, smem, st>>>(q, k, v, o, lse, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

template <int DH>
int launch_bwd_tiled(const View& q, const View& k, const View& v, const View& o, const View& go, const ViewW& dq, const ViewW& dk,
                     const ViewW& dv, const float* lse, float* Dws, int B, int H, int Sq, int Sk, float scale, DxDrop drop,
                     const unsigned long long* seed_dev, cudaStream_t st) {
  const size_t smem = (size_t)(2 * TQ + 2 * TS) * (DH + PAD) * 2 + 2 * TQ * sizeof(float);
  auto kq = attn_mmat_bwd_dq_kernel<DH>;
  auto kkv = attn_mmat_bwd_kv_kernel<DH>;
  static bool attr = false;
  if (!attr) {
    DX_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DX_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 gq(B * H, (Sq + TQ - 1) / TQ), gk(B * H, (Sk + TQ - 1) / TQ);
  kq<<<gq, 256, smem, st>>>(q, k, v, o, go, dq, lse, Dws, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  kkv<<<gk, 256, smem, st>>>(q, k, v, go, dk, dv, lse, Dws, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

inline bool view_ok(const void* p, long long bs, long long rs, int dh) {
  return ((uintptr_t)p % 16 == 0) && (bs % 8 == 0) && (rs % 8 == 0) && (dh % 8 == 0);
}

template <int DH>
int launch_fwd(const View& q, const View& k, const View& v, const ViewW& o, float* lse, int B, int H, int Sq, int Sk,
               float scale, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  const int SPq = (Sq + 15) & ~15, SPk = (Sk + 15) & ~15;
  const size_t smem = (size_t)(SPq + 2 * SPk) * (DH + PAD) * 2;
  auto kern = attn_mma_fwd_kernel<DH>;
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nw = (SPq > SPk ? SPq : SPk) / 16;
  kern<<<B * H, 32 * nw, smem, st>>>(q, k, v, o, lse, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

template <int DH>
int launch_bwd(const View& q, const View& k, const View& v, const View& o, const View& go, const ViewW& dq, const ViewW& dk,
               const ViewW& dv, const float* lse, int B, int H, int Sq, int Sk, float scale, DxDrop drop,
               const unsigned long long* seed_dev, cudaStream_t st) {
  const int SPq = (Sq + 15) & ~15, SPk = (Sk + 15) & ~15;
  const size_t smem = (size_t)(2 * SPq + 2 * SPk) * (DH + PAD) * 2 + 2 * SPq * sizeof(float);
  auto kern = attn_mma_bwd_kernel<DH>;
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nw = (SPq > SPk ? SPq : SPk) / 16;
  kern<<<B * H, 32 * nw, smem, st>>>(q, k, v, o, go, dq, dk, dv, lse, H, Sq, Sk, scale, drop, seed_dev);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // namespace dx_attn_mma

// Entry points used by dx_attention.cu.  Return DX_ERR_UNSUPPORTED (without touching dx_last_error) when the problem does
// not fit this path; the caller then runs the SIMT kernels.
bool dx_attn_mma_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh) {
  if (!(dh == 16 || dh == 32 || dh == 64)) return false;
  if (Sq < 1 || Sk < 1 || Sq > dx_attn_mma::MAX_S || Sk > dx_attn_mma::MAX_S) return false;
  for (int i = 0; i < n; ++i)
    if (!dx_attn_mma::view_ok(ptrs[i], bs[i], rs[i], dh)) return false;
  return true;
}

// tiled kernels: dh 64 / 128, any sequence length
bool dx_attn_mmat_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh) {
  if (!(dh == 64 || dh == 128)) return false;
  if (Sq < 1 || Sk < 1) return false;
  for (int i = 0; i < n; ++i)
    if (!dx_attn_mma::view_ok(ptrs[i], bs[i], rs[i], dh)) return false;
  return true;
}

int dx_attn_mmat_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                     long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                     int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  using namespace dx_attn_mma;
  View Q{(const bf16*)q, q_bs, q_rs}, K{(const bf16*)k, k_bs, k_rs}, V{(const bf16*)v, v_bs, v_rs};
  ViewW O{(bf16*)o, o_bs, o_rs};
  const float scale = 1.f / sqrtf((float)dh);
  if (dh == 64) return launch_fwd_tiled<64>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
  return launch_fwd_tiled<128>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
}

int dx_attn_mmat_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                     long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs, const void* go, long long go_bs,
                     long long go_rs, void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                     void* dv, long long dv_bs, long long dv_rs, const float* lse, float* Dws, int B, int H, int Sq, int Sk, int dh,
                     DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  using namespace dx_attn_mma;
  View Q{(const bf16*)q, q_bs, q_rs}, K{(const bf16*)k, k_bs, k_rs}, V{(const bf16*)v, v_bs, v_rs}, O{(const bf16*)o, o_bs, o_rs},
      GO{(const bf16*)go, go_bs, go_rs};
  ViewW DQ{(bf16*)dq, dq_bs, dq_rs}, DK{(bf16*)dk, dk_bs, dk_rs}, DV{(bf16*)dv, dv_bs, dv_rs};
  const float scale = 1.f / sqrtf((float)dh);
  if (dh == 64) return launch_bwd_tiled<64>(Q, K, V, O, GO, DQ, DK, DV, lse, Dws, B, H, Sq, Sk, scale, drop, seed_dev, st);
  return launch_bwd_tiled<128>(Q, K, V, O, GO, DQ, DK, DV, lse, Dws, B, H, Sq, Sk, scale, drop, seed_dev, st);
}

int dx_attn_mma_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                    long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                    int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  using namespace dx_attn_mma;
  View Q{(const bf16*)q, q_bs, q_rs}, K{(const bf16*)k, k_bs, k_rs}, V{(const bf16*)v, v_bs, v_rs};
  ViewW O{(bf16*)o, o_bs, o_rs};
  const float scale = 1.f / sqrtf((float)dh);
  switch (dh) {
    case 16: return launch_fwd<16>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
    case 32: return launch_fwd<32>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
    default: return launch_fwd<64>(Q, K, V, O, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
  }
}

int dx_attn_mma_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                    long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs, const void* go, long long go_bs,
                    long long go_rs, void* dq, long long dq_bs, long long dq_rs, void* dk, long long dk_bs, long long dk_rs,
                    void* dv, long long dv_bs, long long dv_rs, const float* lse, int B, int H, int Sq, int Sk, int dh,
                    DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  using namespace dx_attn_mma;
  View Q{(const bf16*)q, q_bs, q_rs}, K{(const bf16*)k, k_bs, k_rs}, V{(const bf16*)v, v_bs, v_rs}, O{(const bf16*)o, o_bs, o_rs},
      GO{(const bf16*)go, go_bs, go_rs};
  ViewW DQ{(bf16*)dq, dq_bs, dq_rs}, DK{(bf16*)dk, dk_bs, dk_rs}, DV{(bf16*)dv, dv_bs, dv_rs};
  const float scale = 1.f / sqrtf((float)dh);
  switch (dh) {
    case 16: return launch_bwd<16>(Q, K, V, O, GO, DQ, DK, DV, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
    case 32: return launch_bwd<32>(Q, K, V, O, GO, DQ, DK, DV, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
    default: return launch_bwd<64>(Q, K, V, O, GO, DQ, DK, DV, lse, B, H, Sq, Sk, scale, drop, seed_dev, st);
  }
}
