// tcgen05 / TMEM attention forward for the event axis of the DuETT blocks (x_transformers Attention restated:
// softmax(q k^T / sqrt(dh)) v with attention dropout, reference call site duett/duett.py:95-105,276).
//
// One CTA per (batch row, head, 128-query tile), 128 threads:
//   thread 0     TMA: Q tile {64 dh, 128 rows}, K and V {64 dh, ceil16(Sk) rows} straight out of the packed qkv tensor into
//                128B-swizzled shared memory (rows past the sequence are zero-filled by TMA; issued before the TMEM allocation
//                so the two overlap), then S = Q K^T as 4 tcgen05.mma (M=128, N=ceil16(Sk), k=16 each; both operands K-major)
//   128 threads  thread r owns TMEM lane r = query row r.  Sk <= 160 (the path: 129 events): the whole S row is pulled into
//                registers by one burst of tcgen05.ld (max, exp2, sum, dropout on 4 independent accumulators); longer rows
//                make two passes over TMEM.  P is written as bf16 into shared memory in the K-major A-operand layout
//                (64-key chunks of 128 rows x 128 B)
//   (extra row)  when Sq = 128 k + 1 (the path: 128 events + the [REP] token) the last row is computed on CUDA cores from the
//                K / V tiles already in shared memory, by the same four warps in the bubbles where they would wait for the
//                two MMA batches — a second 128-row tile for one row would cost a whole TMA -> MMA -> softmax -> MMA chain
//   thread 0     O = P V as ceil16(Sk)/16 tcgen05.mma (M=128, N=64): A = P (K-major), B = V used MN-major (V is [key][dh],
//                i.e. N contiguous — the very bytes TMA already delivered); O reuses the TMEM columns of S
//   128 threads  tcgen05.ld of the O row, * 1/sum, bf16 stores; lse
// Shapes: dh = 64, Sk <= 256 (one key pass; S fits 256 TMEM columns -> 2 CTAs per SM), any Sq.  The backward stays on the
// mma.sync kernels (dx_attention_mma.cu); both draw the same counter-based dropout mask ((b,h,row) * Sk + key).
#include "dx_gemm_tc_impl.cuh"
#include <cstdio>

using namespace dx_tc;

namespace {

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr int DH = 64;
constexpr int TQ = 128;
constexpr int NT = 128;   // one thread per query row (TMEM lane)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

PFN_cuTensorMapEncodeTiled_v12000 g_enc = nullptr;

int get_enc() {
  if (g_enc) return DX_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    dx_set_error("cuTensorMapEncodeTiled not available from the driver");
    return DX_ERR_CUDA;
  }
  g_enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return DX_OK;
}

// view [B][S][H*64] (row stride rs, batch stride bs, elements) -> 3-D map, box {64, rows, 1}
int head_map(CUtensorMap* map, const void* base, int H, int S, int B, long long rs, long long bs, int rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)H * DH, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)rs * 2, (cuuint64_t)(B > 1 ? bs : rs) * 2};
  cuuint32_t box[3] = {(cuuint32_t)DH, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dx_set_error("dx_attn_tc: cuTensorMapEncodeTiled failed (%d): H=%d S=%d B=%d rs=%lld bs=%lld rows=%d", (int)r, H, S, B, rs, bs, rows);
    return DX_ERR_CUDA;
  }
  return DX_OK;
}

struct AttnTcParams {
  const bf16* q;   // the extra-row warp reads its query row directly
  long long q_bs, q_rs;
  bf16* o;
  long long o_bs, o_rs;
  float* lse;
  int H, Sq, Sk;
  int SPk;         // ceil16(Sk): instruction N of S = Q K^T, K extent of O = P V
  int tmem_cols;   // power of two >= ceil32(SPk), >= 64
  int extra_row;   // Sq % 128 == 1: index of the one query row past the last full tile (computed on CUDA cores by that tile's CTA), else -1
  float sc2;       // scale * log2(e)
  DxDrop drop;
  const unsigned long long* seed_dev;
  long long* trace;   // debug (DX_ATTN_TC_TRACE): 8 clock64 stamps per CTA
};

// zero-cost ordering point: the value is "produced" here, i.e. after the preceding tcgen05.wait::ld (volatile asm statements keep
// their order), so arithmetic on TMEM-loaded registers cannot be hoisted above the wait
__device__ __forceinline__ float after_wait(uint32_t r) {
  asm volatile("" : "+r"(r));
  return __uint_as_float(r);
}

// One query row on CUDA cores: the path's sequences are 128 events + 1 [REP] token, and a second 128-row tile for that one row
// would cost a whole TMA -> MMA -> softmax -> MMA latency chain.  The row is computed by the CTA's four warps from the K / V
// tiles TMA delivered, in the bubbles where they would otherwise wait for the tensor core:
//   scores (while S = Q K^T runs): thread t takes keys t and t + 128
//   softmax (while O = P V runs):  every warp reduces the 129 scores itself (same order -> same max / sum in all warps)
//   p v:                           warp w takes the 8-key groups g = w mod 4; partial rows summed through shared memory
template <int NB>
__device__ __forceinline__ void extra_row_scores(const AttnTcParams& p, int b, int h, int tid, const uint8_t* sK, float* sXs) {
  const bf16* qp = p.q + (long long)b * p.q_bs + (long long)p.extra_row * p.q_rs + h * DH;
  float qf[DH];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    float t8[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(qp + i * 8)), t8);
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[i * 8 + j] = t8[j];
  }
  const uint32_t uk = smem_u32(sK);
#pragma unroll
  for (int half = 0; half < (NB > 4 ? 2 : 1); ++half) {
    const int key = tid + half * TQ;
    if (key < NB * 32) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if (key < p.Sk) {
        uint4 u[8];
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) lds_raw(uk + stg_off(key, pc), u[pc]);
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) {
          float k8[8];
          unpack8(u[pc], k8);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j & 3] = fmaf(qf[pc * 8 + j], k8[j], a[j & 3]);
        }
      }
      sXs[key] = key < p.Sk ? ((a[0] + a[1]) + (a[2] + a[3])) * p.sc2 : -INFINITY;
    }
  }
}

template <int NB>
__device__ __forceinline__ void extra_row_softmax(const AttnTcParams& p, const DxDrop& drop, int bh, int warp, int lane,
                                                  const float* sXs, float* sXp, float& mx, float& sum) {
  float sc[NB];
  mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    sc[i] = sXs[lane + 32 * i];
    mx = fmaxf(mx, sc[i]);
  }
  mx = dx_warp_max(mx);
  sum = 0.f;
  const unsigned long long drow = ((unsigned long long)bh * p.Sq + (unsigned)p.extra_row) * (unsigned long long)p.Sk;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int key = lane + 32 * i;
    float e = ex2f(sc[i] - mx);            // masked keys: exp2(-inf) = 0
    sum += e;
    if ((i & 3) == warp) {
      if (drop.thresh && key < p.Sk) e *= dx_drop_factor(drop, drow + (unsigned)key);
      sXp[key] = e;
    }
  }
  sum = dx_warp_sum(sum);
}

// partial o[2 lane], o[2 lane + 1] over this warp's 8-key groups: 4 B per lane of each swizzled V row, 8 rows in flight; the rows
// in [Sk, SPk) are zero in V and in p
__device__ __forceinline__ void extra_row_pv(const AttnTcParams& p, int warp, int lane, const uint8_t* sV, const float* sXp, float* sXo) {
  const uint32_t uv = smem_u32(sV);
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t voff = (uint32_t)(lane & 3) * 4, vpc = (uint32_t)lane >> 2;
  for (int j = warp * 8; j < p.SPk; j += 32) {
    uint32_t w[8];
    float pj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {      // (j + u) & 7 == u: j is a multiple of 8
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[u]) : "r"(uv + (uint32_t)(j + u) * 128 + ((vpc ^ (uint32_t)u) << 4) + voff));
      pj[u] = sXp[j + u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      o0[u & 3] = fmaf(pj[u], __uint_as_float(w[u] << 16), o0[u & 3]);
      o1[u & 3] = fmaf(pj[u], __uint_as_float(w[u] & 0xffff0000u), o1[u & 3]);
    }
  }
  sXo[warp * DH + 2 * lane] = (o0[0] + o0[1]) + (o0[2] + o0[3]);
  sXo[warp * DH + 2 * lane + 1] = (o1[0] + o1[1]) + (o1[2] + o1[3]);
}

// NB = ceil32(Sk) / 32 column blocks of S per query row.  NB <= 5 (Sk <= 160, the path's 129 events): the whole row lives in
// registers — one burst of tcgen05.ld, one wait; longer rows make two passes over TMEM.
template <int NB>
__global__ void __launch_bounds__(NT, 2) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                            const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
  constexpr bool RES = NB <= 5;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KV_BYTES = p.SPk * 128;                       // multiple of 2048
  const int KV_STRIDE = (KV_BYTES + 1023) & ~1023;
  uint8_t* sQ = smem;                                     // 128 rows x 128 B
  uint8_t* sK = sQ + TQ * 128;
  uint8_t* sV = sK + KV_STRIDE;
  uint8_t* sP = sV + KV_STRIDE;                           // ceil64(Sk)/64 chunks of 128 rows x 128 B
  constexpr int nchunk = (NB + 1) / 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + nchunk * (TQ * 128));   // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* sXs = reinterpret_cast<float*>(bars + 4);        // extra row: [NB * 32] scores (log2 domain), [NB * 32] probabilities,
  float* sXp = sXs + NB * 32;                             // [4][64] partial outputs of the four warps
  float* sXo = sXp + NB * 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H, q0 = blockIdx.y * TQ;
  const DxDrop drop = dx_drop_resolve(p.drop, p.seed_dev);
  long long* tr = p.trace ? p.trace + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (tr && tid == 0) tr[0] = clock64();

  if (tid == 0) {   // the loads are in flight while warp 1 allocates TMEM
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bars + 0, (uint32_t)(TQ * 128 + 2 * KV_BYTES));
    tma_load_3d(sQ, &tmQ, bars + 0, h * DH, q0, b);
    tma_load_3d(sK, &tmK, bars + 0, h * DH, 0, b);
    tma_load_3d(sV, &tmV, bars + 0, h * DH, 0, b);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr && tid == 0) tr[1] = clock64();     // barriers + TMEM allocation done

  // instruction descriptors: D=f32, A=B=bf16, M=128; S: both K-major, N=SPk; O: B (=V) MN-major, N=64
  const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.SPk >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DH >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);

  const bool extra = p.extra_row >= 0 && (int)blockIdx.y == (int)gridDim.y - 1;   // CTA-uniform
  if (tid == 0 || extra) mbar_wait(bars + 0, 0);   // TMA writes are visible after the barrier wait
  if (tid == 0) {
    tc_fence_after();
    if (tr) tr[2] = clock64();                 // Q, K, V landed
    const uint32_t uq = smem_u32(sQ), uk = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < DH / 16; ++k)
      umma_f16(tmem_base, make_smem_desc(uq + k * 32, 16, 1024), make_smem_desc(uk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
    umma_commit(bars + 1);
  }
  __syncwarp();
  if (extra) extra_row_scores<NB>(p, b, h, tid, sK, sXs);   // while the tensor core computes S

  const int r = tid;                                    // row within the tile (warps 0-3: TMEM lane)
  const int row = q0 + r;
  const bool live = q0 + warp * 32 < p.Sq;              // warp-uniform: warps whose rows are all past Sq skip the math
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t up = smem_u32(sP);
  float mx = -INFINITY, sum = 0.f;
  {
    mbar_wait(bars + 1, 0);
    tc_fence_after();
    if (tr && tid == 0) tr[3] = clock64();     // S in TMEM
    if (live) {
      const unsigned long long drow = ((unsigned long long)bh * p.Sq + (unsigned)row) * (unsigned long long)p.Sk;
      if constexpr (RES) {
        uint32_t sr[NB][32];
#pragma unroll
        for (int c = 0; c < NB; ++c) tmem_ld32_issue(trow + c * 32, sr[c]);
        tmem_ld_wait();
        float s[NB][32];
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[c][i] = after_wait(sr[c][i]);
            if (c == NB - 1) s[c][i] = (c * 32 + i) < p.Sk ? s[c][i] : -INFINITY;    // only the last block can hold masked keys
            m4[i & 3] = fmaxf(m4[i & 3], s[c][i]);
          }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.sc2;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NB; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s[c][i] = ex2f(s[c][i] * p.sc2 - mx);      // masked keys: exp2(-inf) = 0
            s4[i & 3] += s[c][i];
          }
          if (drop.thresh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) s[c][i] *= dx_drop_factor(drop, drow + (unsigned)(c * 32 + i));
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = s[c][g * 8 + i];
            const int kc = c * 32 + g * 8;                  // first key of this 16 B piece
            sts_raw(up + (kc >> 6) * (TQ * 128) + stg_off(r, (kc & 63) >> 3), pack8(pv));
          }
        }
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      } else {
        const int c_full = p.Sk >> 5;                       // column blocks below c_full hold 32 valid keys: no masking
        for (int c = 0; c < NB; ++c) {
          float s[32];
          tmem_ld32(trow + c * 32, s);
          if (c < c_full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, s[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i) < p.Sk ? s[i] : -INFINITY);
          }
        }
        mx *= p.sc2;
        for (int c = 0; c < NB; ++c) {
          float s[32];
          tmem_ld32(trow + c * 32, s);
          if (c < c_full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = ex2f(s[i] * p.sc2 - mx);
              sum += s[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = (c * 32 + i) < p.Sk ? ex2f(s[i] * p.sc2 - mx) : 0.f;
              sum += s[i];
            }
          }
          if (drop.thresh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] *= dx_drop_factor(drop, drow + (unsigned)(c * 32 + i));   // masked keys are 0 already
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = s[g * 8 + i];
            const int kc = c * 32 + g * 8;
            sts_raw(up + (kc >> 6) * (TQ * 128) + stg_off(r, (kc & 63) >> 3), pack8(pv));
          }
        }
      }
    }
  }
  if (tr && tid == 0) tr[7] = clock64();       // warp 0 through its softmax (the "softmax" phase ends when the slowest warp is)
  // P (generic-proxy stores) must be visible to the tensor core's async-proxy reads; S has been read by everybody
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    if (tr) tr[4] = clock64();                 // softmax done, P in shared memory
    const uint32_t uv = smem_u32(sV);
    const int nk = p.SPk >> 4;
    for (int k = 0; k < nk; ++k) {
      const uint32_t pa = up + (k >> 2) * (TQ * 128) + (k & 3) * 32;   // 64-key chunk, 32 B per 16 keys inside the swizzle row
      umma_f16(tmem_base, make_smem_desc(pa, 16, 1024), make_smem_desc(uv + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
    }
    umma_commit(bars + 2);
  }
  __syncwarp();
  if (extra) {                                 // while the tensor core computes O
    float xm, xs;
    extra_row_softmax<NB>(p, drop, bh, warp, lane, sXs, sXp, xm, xs);
    __syncthreads();
    extra_row_pv(p, warp, lane, sV, sXp, sXo);
    __syncthreads();
    if (warp == 0) {
      const float inv = 1.f / xs;
      const float e0 = (sXo[2 * lane] + sXo[DH + 2 * lane]) + (sXo[2 * DH + 2 * lane] + sXo[3 * DH + 2 * lane]);
      const float e1 = (sXo[2 * lane + 1] + sXo[DH + 2 * lane + 1]) + (sXo[2 * DH + 2 * lane + 1] + sXo[3 * DH + 2 * lane + 1]);
      bf16* op = p.o + (long long)b * p.o_bs + (long long)p.extra_row * p.o_rs + h * DH + 2 * lane;
      *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(e0 * inv, e1 * inv);
      if (p.lse && lane == 0) p.lse[(long long)bh * p.Sq + p.extra_row] = (xm + log2f(xs)) * LN2;
    }
  }
  {
    mbar_wait(bars + 2, 0);
    tc_fence_after();
    if (tr && tid == 0) tr[5] = clock64();     // O in TMEM
    if (live) {
      const float inv = 1.f / sum;
      uint32_t orw[DH / 32][32];
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) tmem_ld32_issue(trow + c * 32, orw[c]);
      tmem_ld_wait();
      if (row < p.Sq) {
        bf16* op = p.o + (long long)b * p.o_bs + (long long)row * p.o_rs + h * DH;
#pragma unroll
        for (int c = 0; c < DH / 32; ++c)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float t8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = after_wait(orw[c][g * 8 + i]) * inv;
            *reinterpret_cast<uint4*>(op + c * 32 + g * 8) = pack8(t8);
          }
        if (p.lse) p.lse[(long long)bh * p.Sq + row] = (mx + log2f(sum)) * LN2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr && tid == 0) tr[6] = clock64();     // outputs stored
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

template <int NB>
int launch_nb(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnTcParams& p, dim3 grid, size_t smem, int B,
              cudaStream_t st) {
  auto kern = attn_tc_fwd_kernel<NB>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  p.trace = nullptr;
  const char* trace_path = getenv("DX_ATTN_TC_TRACE");   // debug: per-phase clock64 averages of this launch -> file (synchronises)
  if (trace_path && *trace_path) {
    const size_t n = (size_t)grid.x * grid.y * 8;
    DX_CUDA(cudaMalloc(&p.trace, n * sizeof(long long)));
    DX_CUDA(cudaMemsetAsync(p.trace, 0, n * sizeof(long long), st));
    kern<<<grid, NT, smem, st>>>(tq, tk, tv, p);
    DX_CUDA(cudaStreamSynchronize(st));
    long long* hbuf = (long long*)malloc(n * sizeof(long long));
    DX_CUDA(cudaMemcpy(hbuf, p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
    DX_CUDA(cudaFree(p.trace));
    if (FILE* f = fopen(trace_path, "a")) {
      static const char* names[6] = {"setup", "loads", "qk_mma", "softmax", "pv_mma", "epilogue"};
      for (unsigned y = 0; y < grid.y; ++y) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, tot = 0, w0 = 0;
        for (unsigned x = 0; x < grid.x; ++x) {
          const long long* t = hbuf + ((size_t)y * grid.x + x) * 8;
          for (int i = 0; i < 6; ++i) acc[i] += (double)(t[i + 1] - t[i]);
          tot += (double)(t[6] - t[0]);
          w0 += (double)(t[7] - t[3]);
        }
        fprintf(f, "{\"B\": %d, \"H\": %d, \"Sq\": %d, \"Sk\": %d, \"q_tile\": %u, \"cycles\": {", B, p.H, p.Sq, p.Sk, y);
        for (int i = 0; i < 6; ++i) fprintf(f, "\"%s\": %.0f, ", names[i], acc[i] / grid.x);
        fprintf(f, "\"softmax_warp0\": %.0f, \"total\": %.0f}}\n", w0 / grid.x, tot / grid.x);
      }
      fclose(f);
    }
    free(hbuf);
    return DX_OK;
  }
  kern<<<grid, NT, smem, st>>>(tq, tk, tv, p);
  DX_CUDA(cudaGetLastError());
  return DX_OK;
}

}  // namespace

// DX_ATTN_TC=0 keeps the event axis on the mma.sync kernels (A/B timing)
static bool attn_tc_enabled() {
  const char* e = getenv("DX_ATTN_TC");
  return !(e && atoi(e) == 0);
}

bool dx_attn_tc_supported(const void* const* ptrs, const long long* bs, const long long* rs, int n, int Sq, int Sk, int dh) {
  if (!attn_tc_enabled()) return false;
  if (dh != DH || Sk < 1 || Sk > 256 || Sq < 64) return false;
  for (int i = 0; i < n; ++i)
    if (((uintptr_t)ptrs[i] % 16) || (bs[i] % 8) || (rs[i] % 8)) return false;
  return true;
}

int dx_attn_tc_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs, const void* v,
                   long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs, float* lse, int B, int H, int Sq,
                   int Sk, int dh, DxDrop drop, const unsigned long long* seed_dev, cudaStream_t st) {
  if (dh != DH || Sk > 256) {
    dx_set_error("dx_attn_tc_fwd: unsupported shape dh=%d Sk=%d", dh, Sk);
    return DX_ERR_UNSUPPORTED;
  }
  int rc = get_enc();
  if (rc) return rc;
  const int SPk = (Sk + 15) & ~15;
  CUtensorMap tq, tk, tv;
  if ((rc = head_map(&tq, q, H, Sq, B, q_rs, q_bs, TQ))) return rc;
  if ((rc = head_map(&tk, k, H, Sk, B, k_rs, k_bs, SPk))) return rc;
  if ((rc = head_map(&tv, v, H, Sk, B, v_rs, v_bs, SPk))) return rc;
  AttnTcParams p;
  p.q = (const bf16*)q;
  p.q_bs = q_bs;
  p.q_rs = q_rs;
  p.o = (bf16*)o;
  p.o_bs = o_bs;
  p.o_rs = o_rs;
  p.lse = lse;
  p.H = H;
  p.Sq = Sq;
  p.Sk = Sk;
  p.SPk = SPk;
  const int need = ((SPk + 31) & ~31) > DH ? ((SPk + 31) & ~31) : DH;
  p.tmem_cols = need <= 64 ? 64 : (need <= 128 ? 128 : 256);
  p.extra_row = (Sq > TQ && Sq % TQ == 1) ? Sq - 1 : -1;
  p.sc2 = LOG2E / sqrtf((float)dh);
  p.drop = drop;
  p.seed_dev = seed_dev;
  const int NB = (Sk + 31) >> 5;
  const int kv_stride = (SPk * 128 + 1023) & ~1023;
  const size_t smem = 1024 + TQ * 128 + 2 * (size_t)kv_stride + (size_t)((NB + 1) / 2) * TQ * 128 + 32 + (size_t)(2 * NB * 32 + 4 * DH) * sizeof(float);
  dim3 grid((unsigned)(B * H), (unsigned)(p.extra_row >= 0 ? Sq / TQ : (Sq + TQ - 1) / TQ));
  switch (NB) {
    case 1: return launch_nb<1>(tq, tk, tv, p, grid, smem, B, st);
    case 2: return launch_nb<2>(tq, tk, tv, p, grid, smem, B, st);
    case 3: return launch_nb<3>(tq, tk, tv, p, grid, smem, B, st);
    case 4: return launch_nb<4>(tq, tk, tv, p, grid, smem, B, st);
    case 5: return launch_nb<5>(tq, tk, tv, p, grid, smem, B, st);
    case 6: return launch_nb<6>(tq, tk, tv, p, grid, smem, B, st);
    case 7: return launch_nb<7>(tq, tk, tv, p, grid, smem, B, st);
    default: return launch_nb<8>(tq, tk, tv, p, grid, smem, B, st);
  }
}
