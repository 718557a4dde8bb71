// Host-side dispatch of the tcgen05 GEMM (kernel + launcher: dx_gemm_tc_impl.cuh; instantiations: dx_gemm_tc_inst_*.cu).
#include "dx_gemm_tc_impl.cuh"

namespace dx_tc {
DX_TC_ALL_GROUPS(DX_TC_DECLARE)
DX_TC_GROUP_8(DX_TC32_DECLARE)
}

using namespace dx_tc;

namespace {

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode_fn() {
  if (g_encode) return DX_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    dx_set_error("cuTensorMapEncodeTiled not available from the driver");
    return DX_ERR_CUDA;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return DX_OK;
}

// 3-D bf16 tensor map: dim0 = contiguous (size inner), dim1 = rows (size outer, stride ld elements),
// dim2 = batch (stride bs elements).
int make_tmap(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int batch, long long bs,
              int box_inner, int box_outer, int esz = 2, bool atom32 = false) {
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)(batch > 1 ? batch : 1)};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * esz, (cuuint64_t)(batch > 1 ? bs * esz : ld * esz)};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                        const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dx_set_error("cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld ld=%lld batch=%d bs=%lld box=%dx%d base=%p",
                 (int)r, inner, outer, ld, batch, bs, box_inner, box_outer, base);
    return DX_ERR_CUDA;
  }
  return DX_OK;
}

// CTA-pair instantiations exist for the three operand layouts of the deep-K GEMMs of the path
template <bool A_MN, bool B_MN, bool STAGED>
constexpr bool pair_instantiated() { return (!A_MN && STAGED) || (A_MN && B_MN && !STAGED); }

template <bool A_MN, bool B_MN, bool STAGED>
int launch_major(const dx_gemm_desc* d, int bn, int stages, int ctas, const CUtensorMap& ta, const CUtensorMap& tb,
                 const CUtensorMap& tr, const CUtensorMap& tx, const TcParams& p, const DxEpi& e, cudaStream_t stream) {
  if (ctas == 3) {   // B-multicast cluster of two single-CTA MMAs: HBM-bound staged shapes
    if constexpr (!A_MN && STAGED) {
      if (bn == 256 && stages == 3) return launch_cfg<256, 3, A_MN, B_MN, STAGED, 3>(d, ta, tb, tr, tx, p, e, stream);
      if (bn == 256 && stages == 2) return launch_cfg<256, 2, A_MN, B_MN, STAGED, 3>(d, ta, tb, tr, tx, p, e, stream);
    }
    dx_set_error("dx_gemm_tc: no B-multicast instance for BN=%d stages=%d a_mn=%d staged=%d", bn, stages, (int)A_MN, (int)STAGED);
    return DX_ERR_UNSUPPORTED;
  }
  if (ctas == 2) {
    if constexpr (pair_instantiated<A_MN, B_MN, STAGED>()) {
      if (bn == 256 && stages == 6) return launch_cfg<256, 6, A_MN, B_MN, STAGED, 2>(d, ta, tb, tr, tx, p, e, stream);
      if (bn == 256 && stages == 4) return launch_cfg<256, 4, A_MN, B_MN, STAGED, 2>(d, ta, tb, tr, tx, p, e, stream);
      if (bn == 192 && stages == 6) return launch_cfg<192, 6, A_MN, B_MN, STAGED, 2>(d, ta, tb, tr, tx, p, e, stream);
    }
    dx_set_error("dx_gemm_tc: no CTA-pair instance for BN=%d stages=%d a_mn=%d b_mn=%d staged=%d", bn, stages, (int)A_MN,
                 (int)B_MN, (int)STAGED);
    return DX_ERR_UNSUPPORTED;
  }
  if (bn == 256 && stages == 4) return launch_cfg<256, 4, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 256 && stages == 3) return launch_cfg<256, 3, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 256 && stages == 2) return launch_cfg<256, 2, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 192 && stages == 4) return launch_cfg<192, 4, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 192 && stages == 3) return launch_cfg<192, 3, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 128 && stages == 2) return launch_cfg<128, 2, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 128 && stages == 3) return launch_cfg<128, 3, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 128 && stages == 4) return launch_cfg<128, 4, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 128 && stages == 6) return launch_cfg<128, 6, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 64 && stages == 4) return launch_cfg<64, 4, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  if (bn == 64 && stages == 6) return launch_cfg<64, 6, A_MN, B_MN, STAGED>(d, ta, tb, tr, tx, p, e, stream);
  dx_set_error("dx_gemm_tc: unsupported tile config BN=%d stages=%d", bn, stages);
  return DX_ERR_UNSUPPORTED;
}

// fp32 operands on kind::tf32 (direct epilogue, single CTA)
template <bool A_MN, bool B_MN>
int launch_tf32_major(const dx_gemm_desc* d, int bn, int stages, const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p,
                      const DxEpi& e, cudaStream_t stream) {
  if (bn == 256 && stages == 4) return launch_cfg<256, 4, A_MN, B_MN, false, 1, true>(d, ta, tb, ta, ta, p, e, stream);
  if (bn == 192 && stages == 4) return launch_cfg<192, 4, A_MN, B_MN, false, 1, true>(d, ta, tb, ta, ta, p, e, stream);
  if (bn == 128 && stages == 6) return launch_cfg<128, 6, A_MN, B_MN, false, 1, true>(d, ta, tb, ta, ta, p, e, stream);
  if (bn == 64 && stages == 6) return launch_cfg<64, 6, A_MN, B_MN, false, 1, true>(d, ta, tb, ta, ta, p, e, stream);
  dx_set_error("dx_gemm_tc: no tf32 instance for BN=%d stages=%d", bn, stages);
  return DX_ERR_UNSUPPORTED;
}

int launch_tf32(const dx_gemm_desc* d, cudaStream_t stream) {
  const int batch = d->batch > 1 ? d->batch : 1;
  DX_CHECK_ARG(((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0), "dx_gemm_tc(tf32): A/B must be 16 B aligned");
  DX_CHECK_ARG((d->lda % 4 == 0) && (d->ldb % 4 == 0), "dx_gemm_tc(tf32): lda/ldb must be multiples of 4 (got %lld, %lld)",
               (long long)d->lda, (long long)d->ldb);
  DX_CHECK_ARG(batch == 1 || ((d->a_bs % 4 == 0) && (d->b_bs % 4 == 0)), "dx_gemm_tc(tf32): batch strides must be multiples of 4");
  DX_CHECK_ARG(batch <= 65535, "dx_gemm_tc: batch too large");
  DxEpi e = dx_make_epi(d);
  int bn = d->N <= 64 ? 64 : (d->N >= 256 ? 256 : 128);
  if (d->N > 128 && d->N < 1024 && (d->N % 256) > 128 && (d->N % 256) <= 192 && (d->N % 192) == 0) bn = 192;
  const int stages = bn >= 192 ? 4 : 6;
  constexpr int BKE = 32, CHE = 32;    // fp32 elements per 128 B swizzle row
  CUtensorMap ta, tb;
  int rc;
  if (!d->a_mn) rc = make_tmap(&ta, d->A, d->K, d->M, d->lda, batch, d->a_bs, BKE, BM, 4);
  else rc = make_tmap(&ta, d->A, d->M, d->K, d->lda, batch, d->a_bs, CHE, BKE, 4, true);
  if (rc) return rc;
  if (!d->b_mn) rc = make_tmap(&tb, d->B, d->K, d->N, d->ldb, batch, d->b_bs, BKE, bn, 4);
  else rc = make_tmap(&tb, d->B, d->N, d->K, d->ldb, batch, d->b_bs, CHE, BKE, 4, true);
  if (rc) return rc;
  TcParams p;
  p.K = d->K;
  // K-major tf32: SWIZZLE_128B like bf16 (LBO unused, SBO = 8 rows x 128 B).
  // MN-major tf32: the tensor core only takes the 128B swizzle with 32 B atomicity (cute: Swizzle<2,5,2>, descriptor layout
  // type SWIZZLE_128B_BASE32B, TMA mode 128B_ATOM_32B): atoms of 4 k-rows x 128 B, so SBO = 512 B between atoms and
  // LBO = distance between 32-element MN chunks (one {32 mn, 32 k} box = 4096 B).
  p.a_lbo = d->a_mn ? BKE * 128 : 16;
  p.a_sbo = d->a_mn ? 512 : 1024;
  p.b_lbo = d->b_mn ? BKE * 128 : 16;
  p.b_sbo = d->b_mn ? 512 : 1024;
  if (const char* env = getenv("DX_TF32_SBO")) { if (d->a_mn) p.a_sbo = atoi(env); if (d->b_mn) p.b_sbo = atoi(env); }
  p.stage_bufs = 0;
  p.stage_ring = 1;
  p.epi_mask = -1;
  if (!d->a_mn && !d->b_mn) return launch_tf32_major<false, false>(d, bn, stages, ta, tb, p, e, stream);
  if (!d->a_mn && d->b_mn) return launch_tf32_major<false, true>(d, bn, stages, ta, tb, p, e, stream);
  if (d->a_mn && !d->b_mn) return launch_tf32_major<true, false>(d, bn, stages, ta, tb, p, e, stream);
  return launch_tf32_major<true, true>(d, bn, stages, ta, tb, p, e, stream);
}

template <bool STAGED>
int launch_staged(const dx_gemm_desc* d, int bn, int stages, int ctas, const CUtensorMap& ta, const CUtensorMap& tb,
                  const CUtensorMap& tr, const CUtensorMap& tx, const TcParams& p, const DxEpi& e, cudaStream_t stream) {
  if (!d->a_mn && !d->b_mn) return launch_major<false, false, STAGED>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
  if (!d->a_mn && d->b_mn) return launch_major<false, true, STAGED>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
  if constexpr (!STAGED) {   // MN-major A only occurs in dW GEMMs (fp32 accumulate, direct epilogue): no staged instances
    if (d->a_mn && !d->b_mn) return launch_major<true, false, STAGED>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
    return launch_major<true, true, STAGED>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
  }
  dx_set_error("dx_gemm_tc: internal: staged epilogue with MN-major A");
  return DX_ERR_UNSUPPORTED;
}

}  // namespace

// Wave-quantisation fix for the CTA-pair GEMMs: the rows covered by whole waves of 256-row pair tiles run on the pair kernel,
// the remaining rows as a second launch of 128-row single-CTA tiles (twice as many, half as long: 258 pair tiles on 74 pairs
// = 3.49 -> 4 waves become 3 + ~0.55).  Only the M extent and the row-indexed pointers of the descriptor change.
static thread_local int g_tail_depth = 0;      // > 0: inside a split launch (no further splitting)
static thread_local int g_force_single = 0;    // the tail part: single-CTA tiles

static dx_gemm_desc rows_from(const dx_gemm_desc& d, int m1) {
  dx_gemm_desc t = d;
  const long long osz = d.out_dtype == DX_BF16 ? 2 : 4, asz = d.act_dtype == DX_BF16 ? 2 : 4;
  auto adv = [](const void* p, long long bytes) -> void* { return p ? (void*)((const char*)p + bytes) : nullptr; };
  t.M = d.M - m1;
  t.A = adv(d.A, (d.a_mn ? (long long)m1 : (long long)m1 * d.lda) * 2);
  t.out = adv(d.out, (long long)m1 * d.ldo * osz);
  t.out2 = adv(d.out2, (long long)m1 * d.ldo2 * asz);
  t.res = adv(d.res, (long long)m1 * d.ldr * asz);
  t.aux = adv(d.aux, (long long)m1 * d.ldx * asz);
  t.cx = adv(d.cx, (long long)m1 * d.ldc * asz);
  t.row_scale = (const float*)adv(d.row_scale, 4LL * m1);
  t.row_scale2 = (const float*)adv(d.row_scale2, 4LL * m1);
  t.coef_num = (const float*)adv(d.coef_num, 4LL * m1);
  t.coef_den = (const float*)adv(d.coef_den, 4LL * m1);
  t.row_sumsq = (float*)adv(d.row_sumsq, 4LL * m1);
  t.row_dot = (float*)adv(d.row_dot, 4LL * m1);
  return t;
}

int dx_gemm_tc_launch(const dx_gemm_desc* d, int bn, int stages, int a_lbo, int a_sbo, int b_lbo, int b_sbo,
                      cudaStream_t stream) {
  int rc = get_encode_fn();
  if (rc) return rc;
  if (d->in_dtype == DX_F32) return launch_tf32(d, stream);
  const int batch = d->batch > 1 ? d->batch : 1;
  DX_CHECK_ARG(d->in_dtype == DX_BF16, "dx_gemm_tc: inputs must be bf16");
  DX_CHECK_ARG(((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0), "dx_gemm_tc: A/B must be 16 B aligned");
  DX_CHECK_ARG((d->lda % 8 == 0) && (d->ldb % 8 == 0), "dx_gemm_tc: lda/ldb must be multiples of 8 (got %lld, %lld)",
               (long long)d->lda, (long long)d->ldb);
  DX_CHECK_ARG(batch == 1 || ((d->a_bs % 8 == 0) && (d->b_bs % 8 == 0)), "dx_gemm_tc: batch strides must be multiples of 8");
  DX_CHECK_ARG(batch <= 65535, "dx_gemm_tc: batch too large");
  DxEpi e = dx_make_epi(d);
  // Staged (coalesced) epilogue whenever every [M,N] side tensor is bf16 and 16 B aligned.
  const bool has_o2 = d->out2 != nullptr && (d->act == DX_ACT_GELU || d->act == DX_ACT_GELU_BWD);
  const bool any_side = d->out || d->res || d->aux || d->cx;
  const bool staged = any_side && e.vec_ok && d->act_dtype == DX_BF16 && (!d->out || d->out_dtype == DX_BF16) &&
                      !d->accumulate && (d->N % 8 == 0) && !(d->aux && d->cx) && !d->a_mn;
  // staging blocks per epilogue warp: R[ring] (res -> out in place), X[ring] (aux|cx), O (out2).  Shallow-K GEMMs with [M,N]
  // side inputs are HBM-bound on those inputs: they get the 2-deep ring (prefetch one item ahead); deep-K ones keep the
  // shared memory for the operand ring.
  const bool side_in = d->res || d->aux || d->cx;
  int ring = (staged && side_in && d->K <= 1024) ? 2 : 1;
  if (const char* env = getenv("DX_GEMM_STAGE_RING")) ring = (atoi(env) == 2 && staged && side_in) ? 2 : 1;
  const bool user_cfg = bn > 0;
  int nbufs = 0;
  for (;; ring = 1) {
    nbufs = staged ? (ring + ((d->aux || d->cx) ? ring : (has_o2 ? 1 : 0))) : 0;
    const int budget = 232448 - 1536 - (staged ? NEPI * (nbufs * STG_BYTES + 512) : 0);
    if (user_cfg) {
      if (ring == 1 || stages * (BM + bn) * BK * 2 <= budget) break;
      continue;
    }
    // 128x256 tiles halve the L2 operand re-reads of 128x128 ones (every shape with N >= 256 uses them); narrow outputs get
    // 128x128 / 128x64.  The operand ring takes the deepest instantiated depth that fits next to the epilogue staging.
    bn = d->N <= 64 ? 64 : (d->N >= 256 ? 256 : 128);
    // narrow outputs whose last 256-column tile would be half empty (N = 384: QKV with 2 heads of 64) use 128x192 tiles
    if (d->N > 128 && d->N < 1024 && (d->N % 256) > 128 && (d->N % 256) <= 192 && (d->N % 192) == 0) bn = 192;
    const int stage_bytes = (BM + bn) * BK * 2;
    const int cand[4][3] = {{6, 4, 4}, {6, 4, 3}, {4, 3, 2}, {4, 3, 3}};   // 64-wide tiles stream A: the deeper ring keeps more bytes in flight
    const int* c = cand[bn == 64 ? 0 : (bn == 128 ? 1 : (bn == 256 ? 2 : 3))];
    stages = 0;
    for (int i = 0; i < 3; ++i)
      if (c[i] * stage_bytes <= budget) { stages = c[i]; break; }
    if (stages || ring == 1) {
      if (!stages) stages = c[2];   // reported as unsupported by launch_cfg
      if (const char* env = getenv("DX_GEMM_FORCE_STAGES")) {   // experiments
        const int fs = atoi(env);
        if (fs > 0) stages = fs;
      }
      break;
    }
  }
  // CTA pairs (256 x 256 tiles over two SMs) for the deep-K contractions: they are bound by the L2 -> SM operand feed with
  // single-CTA tiles.  DX_GEMM_PAIR=0 disables, =1 forces the pair kernel wherever an instance exists.
  int ctas = 1;
  {
    const bool inst = (!d->a_mn && staged) || (d->a_mn && d->b_mn && !staged);
    // a half-empty last 256-row tile wastes a quarter of the MMA work of a short M (dW of the 384-wide QKV weight)
    const bool m_ok = d->M >= 2048 || (d->M % 256) == 0 || (d->M % 256) > 128;
    const bool legal = !user_cfg && inst && (bn == 256 || bn == 192) && batch == 1 && d->M > 128 && m_ok;
    int mode = -1;
    if (const char* env = getenv("DX_GEMM_PAIR")) mode = atoi(env);
    if (legal && (mode == 1 || (mode != 0 && d->K >= 1024))) {
      const int budget = 232448 - 1536 - (staged ? NEPI * (nbufs * STG_BYTES + 512) : 0);
      const int stage_bytes = (BM + bn / 2) * BK * 2;
      const int st2 = 6 * stage_bytes <= budget ? 6 : ((bn == 256 && 4 * stage_bytes <= budget) ? 4 : 0);
      if (st2 && !g_force_single) { ctas = 2; stages = st2; }
    }
  }
  if (ctas == 2 && g_tail_depth == 0) {
    // Measured on B200 (profiles/r02_gemm_exp_tailsplit.json): NOT a win — a 128-row single-CTA tile stages 48 KB per k-block
    // against the pair's 32 KB per SM, and these GEMMs are bound by the per-SM operand ingest, so the "half-height" tail
    // tiles take ~0.8x (not 0.5x) of a pair tile: FFN-in 142 -> 161 us, dW 137 -> 181 us.  Off unless DX_GEMM_TAILSPLIT=1.
    static int dev_sms = 0;
    const char* ts_env = getenv("DX_GEMM_TAILSPLIT");
    const int tail_split = ts_env ? (atoi(ts_env) != 0) : 0;
    if (!dev_sms) {
      int dev = 0;
      DX_CUDA(cudaGetDevice(&dev));
      DX_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int sms = dev_sms - dx_gemm_reserved_sms() > 16 ? dev_sms - dx_gemm_reserved_sms() : dev_sms;
    const int workers = sms / 2, tiles_n = dx_ceil_div(d->N, bn), tiles_m = dx_ceil_div(d->M, 2 * BM);
    const int total = tiles_m * tiles_n, full = total / workers, rem = total - full * workers;
    const int m1 = (full * workers / tiles_n) * 2 * BM;      // rows covered by whole waves (whole 256-row blocks only)
    if (tail_split && full >= 1 && rem > 0 && rem * 10 <= workers * 8 && m1 > 0 && m1 < d->M &&
        (long long)dx_ceil_div(d->M - m1, BM) * tiles_n <= sms) {
      dx_gemm_desc head = *d, tail = rows_from(*d, m1);
      head.M = m1;
      ++g_tail_depth;
      rc = dx_gemm_tc_launch(&head, 0, 0, a_lbo, a_sbo, b_lbo, b_sbo, stream);
      if (!rc) {
        g_force_single = 1;
        rc = dx_gemm_tc_launch(&tail, 0, 0, a_lbo, a_sbo, b_lbo, b_sbo, stream);
        g_force_single = 0;
      }
      --g_tail_depth;
      return rc;
    }
  }
  // B-multicast cluster for the HBM-bound staged shapes (shallow K, many M tiles re-reading the same B block).  Measured
  // on B200: it removes a quarter of their L2 -> SM traffic (2.29 -> 1.72 GB per launch) but not a microsecond of their
  // time (dX 187.4 us either way), i.e. those kernels are not bound by L2 throughput; it stays opt-in (DX_GEMM_MCAST=1,
  // covered by the kernel tests) and off by default.
  if (ctas == 1 && !user_cfg && staged && !d->a_mn && bn == 256 && batch == 1 && (stages == 3 || stages == 2)) {
    if (const char* env = getenv("DX_GEMM_MCAST"))
      if (atoi(env) == 1) ctas = 3;
  }
  CUtensorMap ta, tb;
  if (!d->a_mn) rc = make_tmap(&ta, d->A, d->K, d->M, d->lda, batch, d->a_bs, BK, BM);
  else rc = make_tmap(&ta, d->A, d->M, d->K, d->lda, batch, d->a_bs, 64, BK);
  if (rc) return rc;
  if (!d->b_mn) rc = make_tmap(&tb, d->B, d->K, d->N, d->ldb, batch, d->b_bs, BK, ctas == 1 ? bn : bn / 2);
  else rc = make_tmap(&tb, d->B, d->N, d->K, d->ldb, batch, d->b_bs, 64, BK);
  if (rc) return rc;
  TcParams p;
  p.K = d->K;
  // K-major SW128: LBO unused (canonical value 1 -> 16 B), SBO = 8 rows * 128 B.
  // MN-major SW128: LBO = distance between 64-element MN chunks (one 64x64 TMA box = 8192 B),
  //                 SBO = distance between 8-row K groups (1024 B).
  p.a_lbo = a_lbo >= 0 ? a_lbo : (d->a_mn ? 8192 : 16);
  p.a_sbo = a_sbo >= 0 ? a_sbo : 1024;
  p.b_lbo = b_lbo >= 0 ? b_lbo : (d->b_mn ? 8192 : 16);
  p.b_sbo = b_sbo >= 0 ? b_sbo : 1024;
  p.stage_bufs = nbufs;
  p.stage_ring = ring;
  p.epi_mask = staged ? dx_epi_mask(d) : -1;
  // side tensors of the staged epilogue: [32 rows x 64 columns] boxes of the [M,N] matrices (batch = third dimension)
  CUtensorMap tr = ta, tx = ta;
  if (staged && d->res) {
    rc = make_tmap(&tr, d->res, d->N, d->M, d->ldr, batch, d->res_bs, 64, 32);
    if (rc) return rc;
  }
  if (staged && (d->aux || d->cx)) {
    rc = d->aux ? make_tmap(&tx, d->aux, d->N, d->M, d->ldx, batch, d->aux_bs, 64, 32)
                : make_tmap(&tx, d->cx, d->N, d->M, d->ldc, batch, d->cx_bs, 64, 32);
    if (rc) return rc;
  }
  if (staged) return launch_staged<true>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
  return launch_staged<false>(d, bn, stages, ctas, ta, tb, tr, tx, p, e, stream);
}
