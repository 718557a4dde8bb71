// fp32 FFMA GEMM with the shared fused epilogue.  This is the fp32 precision mode's contraction kernel
// (parity at 1e-3 against the CPU oracle needs true fp32 products; tcgen05 kind::tf32 would not give it),
// the kernel of the small fp32 MLPs around the backbone (tab_encoder, cve front, heads) and the on-device cross-check
// for the tcgen05 kernel.  Tile 64x128x16, 256 threads, 4x8 micro-tile.
// blockIdx.z = batch index (grouped mode) or K-split index (split_k > 1: skinny outputs with a huge K such as the
// [B, (V+1)d] x [(V+1)d, 64] head projection; partial sums are red.add'ed into a zero-initialised fp32 output).
#include "dx_gemm_epilogue.cuh"

namespace {

constexpr int BM = 64, BN = 128, BK = 16, NT = 256;

template <typename TI>
__global__ void __launch_bounds__(NT) dx_gemm_simt_kernel(const TI* __restrict__ A, long long sam, long long sak,
                                                         const TI* __restrict__ B, long long sbn, long long sbk,
                                                         int K, DxEpi e, long long a_bs, long long b_bs, int splits) {
  int k_lo = 0, k_hi = K;
  if (splits > 1) {
    const int per = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    k_lo = blockIdx.z * per;
    k_hi = min(K, k_lo + per);
  } else {
    A += (long long)blockIdx.z * a_bs;
    B += (long long)blockIdx.z * b_bs;
    dx_epi_select_batch(e, blockIdx.z);
  }
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = t >> 4, tx = t & 15;  // rows ty*4.., cols tx*8..
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_kmajor = (sak == 1), b_kmajor = (sbk == 1);
  // register-staged software pipeline: the global loads of k-block i+1 are in flight while k-block i is multiplied
  constexpr int AE = (BM * BK) / NT, BE = (BN * BK) / NT;
  float ra[AE], rb[BE];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < AE; ++i) {
      const int idx = t + i * NT;
      int m, k;
      if (a_kmajor) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < e.M && gk < k_hi) ? dx_ld(A + (long long)gm * sam + (long long)gk * sak) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < BE; ++i) {
      const int idx = t + i * NT;
      int n, k;
      if (b_kmajor) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < e.N && gk < k_hi) ? dx_ld(B + (long long)gn * sbn + (long long)gk * sbk) : 0.f;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < AE; ++i) {
      const int idx = t + i * NT;
      int m, k;
      if (a_kmajor) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < BE; ++i) {
      const int idx = t + i * NT;
      int n, k;
      if (b_kmajor) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      Bs[k][n] = rb[i];
    }
  };
  if (k_lo < k_hi) gload(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
    sstore();
    __syncthreads();
    if (k0 + BK < k_hi) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= e.M) continue;
    const int nn = n0 + tx * 8;
    if (splits > 1) {
      if (blockIdx.z == 0 && e.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += (nn + j < e.N) ? e.bias[nn + j] : 0.f;
      }
      dx_epi_atomic_add8(e.out, e.ldo, m, nn, e.N, acc[i]);
      continue;
    }
    float rs = 0.f, rd = 0.f;
    dx_epilogue_chunk<8>(e, m, nn, acc[i], rs, rd);
    if (nn < e.N) dx_epilogue_flush_row(e, m, rs, rd);
  }
}

}  // namespace

int dx_gemm_simt_launch(const dx_gemm_desc* d, cudaStream_t stream) {
  DxEpi e = dx_make_epi(d);
  const long long sam = d->a_mn ? 1 : d->lda, sak = d->a_mn ? d->lda : 1;
  const long long sbn = d->b_mn ? 1 : d->ldb, sbk = d->b_mn ? d->ldb : 1;
  const int batch = d->batch > 1 ? d->batch : 1;
  int splits = d->split_k > 1 ? d->split_k : 1;
  if (splits > 1) {
    DX_CHECK_ARG(batch == 1 && d->accumulate && d->out && d->out_dtype == DX_F32 && !d->out2 && !d->res && !d->aux && !d->cx &&
                     !d->row_scale && !d->row_sumsq && !d->row_dot && d->act == DX_ACT_NONE,
                 "dx_gemm: split_k needs a plain fp32 accumulate epilogue (bias allowed), batch == 1");
    const int per = ((d->K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (d->K + per - 1) / per;
  }
  dim3 grid(dx_ceil_div(d->N, BN), dx_ceil_div(d->M, BM), splits > 1 ? splits : batch);
  if (d->in_dtype == DX_F32) {
    dx_gemm_simt_kernel<float><<<grid, NT, 0, stream>>>((const float*)d->A, sam, sak, (const float*)d->B, sbn, sbk,
                                                        d->K, e, d->a_bs, d->b_bs, splits);
  } else {
    dx_gemm_simt_kernel<bf16><<<grid, NT, 0, stream>>>((const bf16*)d->A, sam, sak, (const bf16*)d->B, sbn, sbk,
                                                       d->K, e, d->a_bs, d->b_bs, splits);
  }
  DX_LAUNCH_CHECK();
  return DX_OK;
}
