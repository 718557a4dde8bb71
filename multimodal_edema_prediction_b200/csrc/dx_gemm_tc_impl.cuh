// bf16 GEMM on 5th-gen tensor cores: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma kind::f16 (one elected thread) -> fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   acc[m,n] = sum_k A(m,k) * B(n,k); A and B each K-major or MN-major (so X@W^T, dY@W and dY^T@X all run
//   from the tensors as they sit in HBM, without transposed copies).  Grouped mode: blockIdx.z selects one of
//   `batch` independent problems through the third dimension of the tensor maps.
//
// Persistent kernel: every CTA (or CTA pair) loops over output tiles.  CTA tile = 128 x BN (BN in 64/128/192/256),
// BLOCK_K = 64 (one 128 B swizzle row of bf16), STAGES-deep mbarrier operand ring, two TMEM accumulators (2*BN columns) so
// the epilogue of tile i overlaps the mainloop of tile i+1.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = epilogue (TMEM lane quadrant = warp % 4; the two
// warps of a quadrant take alternate 64-column blocks, so every SM scheduler has two epilogue warps to interleave).
// CTAS = 2 (deep-K shapes): a cluster of two CTAs shares a 256 x BN tile through tcgen05.mma.cta_group::2 (see the kernel).
// Work units run N-fastest or M-fastest (TcParams::raster_m), whichever keeps the larger operand block hot in L2; dW
// GEMMs with too few tiles split K (red.global.add partial sums).
//
// Epilogue front-ends:
//   STAGED  (bf16 side tensors, 16 B aligned): each epilogue warp moves 32 rows x 64 columns at a time through its own
//           XOR-swizzled shared-memory staging blocks — the res/cx/aux blocks arrive as TMA boxes one item ahead, stores of
//           out/out2 are full 128 B row segments (4 rows per warp instruction), the thread-per-row TMEM layout only ever
//           touches smem.  The feature set is a compile-time mask (staged_epilogue<BN, MASK, CTAS>).
//   direct  : thread-per-row 16 B vectors straight to global (fp32 outputs such as dW accumulation, unaligned shapes).
#pragma once
#include "dx_gemm_epilogue.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>


// This header holds the kernel template and its launcher; the instantiations are spread over dx_gemm_tc_inst_*.cu (one
// translation unit per operand-layout / mode group, so the library builds in parallel) and dispatched from dx_gemm_tc.cu.
namespace dx_tc {


constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NEPI = 8;                 // epilogue warps: 2 per TMEM lane quadrant (they split the column blocks)
constexpr int NTHREADS = 64 + 32 * NEPI;
constexpr int STG_BYTES = 32 * 128;   // one staging block: 32 rows x 64 bf16

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("dx_gemm_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA prefetch of one box into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// kind::tf32: fp32 operands in shared memory (the tensor core reads the upper 19 bits), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two SMs of a cluster share one 256 x BN tile -----------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose bytes are counted on the LEADER CTA's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA load written into the same shared-memory offset of every CTA in `mask` (and counted on each one's mbarrier at the same
// offset): two CTAs working on different M tiles of the same N tile fetch each half of the B block from L2 only once
__device__ __forceinline__ void tma_load_3d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// single-CTA MMAs, but the completion is signalled on the barrier at this offset in BOTH CTAs of the cluster
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B).
// layout: 2 = SWIZZLE_128B (16 B chunks XOR row%8, 8-row atoms); 1 = SWIZZLE_128B_BASE32B (32 B chunks XOR row%4, 4-row
// atoms) — the only layout the tensor core accepts for MN-major (transposed) tf32 operands; TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

struct TcParams {
  int K;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // descriptor byte offsets (test-overridable)
  int stage_bufs;                        // staging blocks per epilogue warp (STAGED only), see staged_epilogue
  int stage_ring;                        // 1 or 2: ring depth of the res / aux|cx staging blocks (2 = prefetch one item ahead)
  int tiles_m, tiles_n, total_tiles;     // persistent work loop: unit -> (k split, batch z, m block, n block), n fastest
  int splits, kb_per_split;              // split-K (dW GEMMs with few output tiles): partial sums are red.add'ed into out
  int epi_mask;                          // staged epilogue: compile-time feature mask (dx_epi_mask), -1 = runtime flags
  int raster_m;                          // 1: consecutive work units walk M first (few M tiles sharing a large B column block)
  int chunked;                           // 1: every worker owns a contiguous run of work units (unit_range)
  int l2pf;                              // experiment (default 0 = off): every l2pf k-blocks the producer prefetches the NEXT l2pf boxes of its
                                         // K-major A rows into L2 in one burst (cp.async.bulk.prefetch.tensor)
  int fault;                             // test hook (DX_GEMM_FAULT): bit 0 = from a CTA's third tile on the accumulator is NOT reset, i.e. a
                                         // TMEM slot-reuse bug (tests/test_gemm_production_gpu.py proves its checks catch it); bits 1 / 2 =
                                         // the staged epilogue skips its side-tensor loads / its stores (timing experiments, wrong results)
};

// Work units of one persistent worker (CTA or CTA pair): strided (unit = w, w + W, ...: neighbouring workers run neighbouring
// tiles at the same time) or chunked (a contiguous run per worker: with N running fastest one worker walks all N tiles of an
// M block, so only its first pass over the A block comes from DRAM).
__device__ __forceinline__ void unit_range(const TcParams& p, int w, int W, int& first, int& end, int& step);

// work unit -> tile indices.  The fast-running index is the one whose tiles share the LARGER operand block, so that block
// is fetched from DRAM once and hit in L2 by the neighbouring tiles (ncu: 2.1x the algorithmic DRAM bytes for the
// 384 x 16512 dW GEMMs with N running fastest, every M tile re-streaming all of B).
__device__ __forceinline__ void tile_decode(const TcParams& p, int tile, int& mi, int& ni, int& z) {
  const int mn = p.tiles_m * p.tiles_n;
  z = tile / mn;
  const int r = tile - z * mn;
  if (p.raster_m) {
    ni = r / p.tiles_m;
    mi = r - ni * p.tiles_m;
  } else {
    mi = r / p.tiles_n;
    ni = r - mi * p.tiles_n;
  }
}

__device__ __forceinline__ void unit_range(const TcParams& p, int w, int W, int& first, int& end, int& step) {
  if (p.chunked) {
    const int base = p.total_tiles / W, rem = p.total_tiles - base * W;
    first = w * base + (w < rem ? w : rem);
    end = first + base + (w < rem ? 1 : 0);
    step = 1;
  } else {
    first = w;
    end = p.total_tiles;
    step = W;
  }
}

// ---- warp-staged tile movement (STAGED epilogue) -----------------------------------------------------------
// A staging block holds 32 rows x 8 pieces of 16 B; piece p of row r lives at r*128 + ((p ^ (r & 7)) << 4), which is
// conflict-free both for the row-per-lane view (epilogue math) and for the 8-lanes-per-row view (global traffic).
__device__ __forceinline__ uint32_t stg_off(int r, int p) { return (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)); }

__device__ __forceinline__ void stage_store(uint32_t buf, void* base, long long ld, int m_base, int n0, int M, int N, int lane) {
  const int piece = lane & 7, rsub = lane >> 3;
  const int col = n0 + piece * 8;
  if (col >= N) return;
  bf16* gp = reinterpret_cast<bf16*>(base) + (long long)(m_base + rsub) * ld + col;
  const long long gstep = 4 * ld;
  const uint32_t sp = buf + rsub * 128;
  const uint32_t pe = (uint32_t)((piece ^ rsub) << 4), po = (uint32_t)((piece ^ (rsub + 4)) << 4);
  const int rows = M - m_base - rsub;   // rows r = 4*i + rsub valid while 4*i < rows
  if (rows > 28) {   // whole block inside the matrix (the common case): no per-row predicates
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(sp + i * 512 + ((i & 1) ? po : pe)));
      *reinterpret_cast<uint4*>(gp) = u;
      gp += gstep;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (4 * i < rows) {
      uint4 u;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(sp + i * 512 + ((i & 1) ? po : pe)));
      *reinterpret_cast<uint4*>(gp + i * gstep) = u;
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// explicit shared-space 16 B accesses on staging blocks (generic pointers made ptxas emit LD.E/ST.E here)
__device__ __forceinline__ void lds8(uint32_t addr, float (&v)[8]) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  // bf16 -> f32 is a 16-bit left shift: low element = word << 16, high element = word & 0xffff0000 (one op each)
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void lds_raw(uint32_t addr, uint4& u) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void sts_raw(uint32_t addr, const uint4& u) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t addr, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void lds_f8(uint32_t addr, float (&v)[8]) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr + 16));
}

// The staged epilogue of one epilogue warp over ALL of its CTA's tiles.  MASK >= 0: compile-time feature set; -1: runtime.
// The warp's work is a stream of items (tile, 64-column super-chunk).  Everything an item needs from global memory is
// fetched with cp.async one item AHEAD (the first chunk of the NEXT tile included) while the current item is computed and
// stored: always the 64 per-column bias values (a dependent LDG per 8-column piece showed up as 60 % long-scoreboard
// stalls in ncu), and with p.stage_ring == 2 also the [32 x 64] blocks of the side tensors (res, aux|cx).
// Staging blocks of a warp: R[ring] (res, reused in place for out), X[ring] (aux|cx, only if present; reused in place for
// out2), O (out2 when there is no X block); wbias: 2 x 64 floats.
template <int BN, int MASK, int CTAS>
__device__ __forceinline__ void staged_epilogue(const DxEpi& e0, const TcParams& p, const CUtensorMap* tmR, const CUtensorMap* tmX,
                                                uint32_t tmem_base, uint8_t* wstg, uint8_t* wbias, uint64_t* sfull, int q,
                                                int chalf, int lane, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar) {
  constexpr bool CT = MASK >= 0;
  const bool has_res = CT ? ((MASK & DX_M_RES) != 0) : (e0.res != nullptr);
  const bool has_x = CT ? ((MASK & (DX_M_CX | DX_M_GELUBWD)) != 0) : (e0.aux != nullptr || e0.cx != nullptr);
  const bool has_o2 = CT ? ((MASK & (DX_M_GELU | DX_M_GELUBWD)) != 0) : dx_epi_has_out2(e0);
  constexpr bool has_b = CT && ((MASK & (DX_M_BIAS | DX_M_GELUBWD)) != 0);   // bias (or aux_bias) staged in wbias
  const bool xp_noload = (p.fault & 2) != 0, xp_nostore = (p.fault & 4) != 0;   // timing experiments only (DX_GEMM_FAULT bits)
  const bool any_in = (has_res || has_x) && !xp_noload;
  const int ring = p.stage_ring;
  const bool pf_side = any_in && ring == 2;
  const uint32_t stg = smem_u32(wstg), sbias = smem_u32(wbias);
  const uint32_t bufO_own = stg + (p.stage_bufs - 1) * STG_BYTES;   // used when there is no aux|cx block to reuse
  const uint32_t lsw = (uint32_t)(lane & 7);
  // [32 x 64] blocks of res / aux|cx: one TMA box each (128B swizzle = the staging layout), completion on sfull[b].  The
  // bulk-async path keeps whole blocks in flight without occupying L1 miss slots the way 16 B cp.async requests did.
  auto issue_side = [&](int b, int m_base_, int nc, int z_) {
    if (lane == 0) {
      mbar_arrive_expect_tx(sfull + b, (uint32_t)((has_res ? STG_BYTES : 0) + (has_x ? STG_BYTES : 0)));
      if (has_res) tma_load_3d(wstg + b * STG_BYTES, tmR, sfull + b, nc, m_base_, z_);
      if (has_x) tma_load_3d(wstg + (ring + b) * STG_BYTES, tmX, sfull + b, nc, m_base_, z_);
    }
  };
  uint32_t sphase = 0;   // bit b: parity of the next completion of sfull[b]
  auto issue_bias = [&](const DxEpi& e_, int b, int nc) {   // 64 floats = 16 lanes x 16 B, zero-filled past N (N % 8 == 0)
    if (lane < 16) {
      const float* src = (MASK >= 0 && (MASK & DX_M_GELUBWD)) ? e_.aux_bias : e_.bias;
      const int col = nc + lane * 4;
      const bool ok = col < e_.N;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sbias + b * 256 + lane * 16), "l"(ok ? src + col : src),
                   "r"(ok ? 16 : 0)
                   : "memory");
    }
  };
  int bi = 0, bb = 0;                         // ring slots of the current item (side blocks / bias)
  bool side_loaded = false, bias_loaded = false;   // the current item's loads were issued while the previous item ran
  uint32_t tcount = 0;
  // CTA pair: both CTAs walk the same 256-row tiles; this CTA owns rows [rank*128, +128) of each
  constexpr int CL = CTAS == 1 ? 1 : 2;   // CTAs per cluster (CTAS: 1 single, 2 cta_group::2 pair, 3 B-multicast pair)
  const int m_off = CL == 2 ? (int)cluster_ctarank() * BM : 0;
  const int ustep = (int)gridDim.x / CL;
  const uint32_t empty_remote = CTAS == 2 ? mapa_u32(smem_u32(tmem_empty_bar), 0) : 0;   // the leader's tmem_empty_bar[0]
  // the staged epilogue never runs with split-K (work unit == tile); each tile's coordinates are decoded once, one tile
  // ahead (integer divisions), and serve both the cross-tile prefetch and the next iteration
  int mi = 0, ni = 0, z = 0;
  DxRowRaw raw_next{1.f, 1.f, 0.f, 1.f};
  bool raw_ok = false;
  int ufirst, uend, ustride;
  unit_range(p, (int)blockIdx.x / CL, ustep, ufirst, uend, ustride);
  if (ufirst < uend) tile_decode(p, ufirst, mi, ni, z);
  for (int unit = ufirst; unit < uend; unit += ustride, ++tcount) {
    int mi2 = 0, ni2 = 0, z2 = -1;
    if (unit + ustride < uend) tile_decode(p, unit + ustride, mi2, ni2, z2);
    const int n0 = ni * BN;
    const int m0 = mi * (BM * CL) + m_off;
    const uint32_t slot = tcount & 1, use = tcount >> 1;
    const uint32_t acc = tmem_base + slot * BN + ((uint32_t)(q * 32) << 16);
    DxEpi e = e0;
    dx_epi_select_batch(e, z);
    const int m_base = m0 + q * 32;
    const int m = m_base + lane;
    const bool row_ok = m < e.M;
    // per-row constants: fetched one tile ahead when the next tile lies in the same batch slice
    const DxRowConst rc = raw_ok ? dx_row_finish(raw_next) : (row_ok ? dx_row_const(e, m) : DxRowConst{1.f, 1.f, 0.f});
    raw_ok = false;
    if (z2 == z) {
      const int mn = mi2 * (BM * CL) + m_off + q * 32 + lane;
      raw_next = mn < e.M ? dx_row_raw(e, mn) : DxRowRaw{1.f, 1.f, 0.f, 1.f};
      raw_ok = true;
    }
    float rs = 0.f, rd = 0.f;
    const int nsc = min(BN / 64, (e.N - n0 + 63) / 64);
    bool acc_ready = false;
#pragma unroll 1
    for (int sc = chalf; sc < nsc; sc += 2) {
      const int nc = n0 + sc * 64;
      if (any_in && !side_loaded) issue_side(bi, m_base, nc, z);
      if (has_b && !bias_loaded) {
        issue_bias(e, bb, nc);
        cp_async_commit();
      }
      // prefetch the next item of this warp's stream
      bool next = false;
      if (pf_side || has_b) {
        int m_base2 = m_base, nc2 = nc + 128;
        if (sc + 2 < nsc) {
          next = true;
        } else {
          if (z2 == z) {   // next tile exists and lies in the same batch slice: the pointers of `e` are valid for it
            nc2 = ni2 * BN + chalf * 64;
            m_base2 = mi2 * (BM * CL) + m_off + q * 32;
            next = nc2 < e.N;
          }
        }
        if (next) {
          if (pf_side) issue_side(bi ^ 1, m_base2, nc2, z);
          if (has_b) {
            issue_bias(e, bb ^ 1, nc2);
            cp_async_commit();
          }
        }
      }
      if (!acc_ready) {   // side tensors do not depend on the accumulator: their loads run under the mainloop
        mbar_wait(tmem_full_bar + slot, use & 1);
        tc_fence_after();
        acc_ready = true;
      }
      if (has_b) {
        if (next) cp_async_wait<1>();
        else cp_async_wait<0>();
      }
      if (any_in) {
        mbar_wait(sfull + bi, (sphase >> bi) & 1u);
        sphase ^= 1u << bi;
      }
      __syncwarp();
      // piece p of this lane's row lives at row + ((p ^ (lane & 7)) << 4); the blocks are 1024 B aligned, so with the lane's
      // swizzle term folded into the row address the piece address is one XOR with a compile-time constant
      const uint32_t bufR = stg + bi * STG_BYTES;
      const uint32_t rowR = bufR + lane * 128 + (lsw << 4);
      const uint32_t rowX = stg + (ring + bi) * STG_BYTES + lane * 128 + (lsw << 4);
      // out2 is staged in place over the aux block when there is one (each lane rewrites the 16 B piece it has just read)
      const uint32_t bufO = has_x ? stg + (ring + bi) * STG_BYTES : bufO_own;
      const uint32_t rowO = bufO + lane * 128 + (lsw << 4);
      const uint32_t biasS = sbias + bb * 256;
      // both 32-column halves of the item are requested before the first is used (one TMEM round trip per item)
      uint32_t vr[2][32];
      tmem_ld32_issue(acc + (uint32_t)(sc * 64), vr[0]);        // warp-collective
      tmem_ld32_issue(acc + (uint32_t)(sc * 64 + 32), vr[1]);
      tmem_ld_wait();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if constexpr (CT) {
          // Three phases per 32-column half so that the shared-memory latencies overlap instead of chaining piece by piece
          // (a load of piece j+1 cannot be hoisted above the store of piece j by ptxas: possible aliasing): (1) all side /
          // bias loads of the 4 pieces, (2) arithmetic in registers, (3) all stores.
          uint4 ur[4], ux[4];
          float bv[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t px = (uint32_t)(half * 4 + j) << 4;
            if (has_res) lds_raw(rowR ^ px, ur[j]);
            if (has_x) lds_raw(rowX ^ px, ux[j]);
            if (has_b) lds_f8(biasS + (half * 4 + j) * 32, bv[j]);
          }
          uint4 po[4], po2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float t[8], r[8], a[8], o2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = __uint_as_float(vr[half][j * 8 + k]);
            if (has_res) unpack8(ur[j], r);
            if (has_x) unpack8(ux[j], a);
            // no bounds branch: out-of-range rows / columns hold zeros everywhere (operands, side blocks and bias are
            // zero-filled by TMA / cp.async), contribute nothing to the row sums, and are never stored or flushed
            dx_epilogue_math_c<MASK>(rc, t, r, a, bv[j], o2, rs, rd);
            po[j] = pack8(t);
            if (has_o2) po2[j] = pack8(o2);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t px = (uint32_t)(half * 4 + j) << 4;
            sts_raw(rowR ^ px, po[j]);
            if (has_o2) sts_raw(rowO ^ px, po2[j]);
          }
        } else {
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(vr[half][k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int piece = half * 4 + j;
            const uint32_t px = (uint32_t)piece << 4;
            const int ncol = nc + piece * 8;
            float t[8], r[8], a[8], o2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = v[j * 8 + k];
            if (has_res) lds8(rowR ^ px, r);
            if (has_x) lds8(rowX ^ px, a);
            if (row_ok && ncol < e.N) {   // N % 8 == 0 on this path: pieces are whole
              dx_epilogue_math<8>(e, rc, ncol, 8, t, r, a, a, o2, rs, rd);   // aux and cx are mutually exclusive: `a` is both
            }
            sts8(rowR ^ px, t);
            if (has_o2) sts8(rowO ^ px, o2);
          }
        }
      }
      __syncwarp();
      if (e.out && !xp_nostore) stage_store(bufR, e.out, e.ldo, m_base, nc, e.M, e.N, lane);
      if (has_o2 && !xp_nostore) stage_store(bufO, e.out2, e.ldo2, m_base, nc, e.M, e.N, lane);
      if (any_in) fence_proxy_async();   // generic-proxy accesses of these blocks are ordered before the next TMA write into them
      __syncwarp();
      side_loaded = next && pf_side;
      bias_loaded = next && has_b;
      if (pf_side) bi ^= 1;
      if (has_b) bb ^= 1;
    }
    if (!acc_ready) {   // this warp owns no column block of a narrow tile: still take part in the accumulator hand-shake
      mbar_wait(tmem_full_bar + slot, use & 1);
      tc_fence_after();
    }
    // this warp has finished reading the accumulator: hand the TMEM slot back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CTAS == 2) mbar_arrive_cluster(empty_remote + slot * 8);
      else mbar_arrive(tmem_empty_bar + slot);
    }
    if (row_ok) dx_epilogue_flush_row(e, m, rs, rd);
    mi = mi2; ni = ni2; z = z2;
  }
}

// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool A_MN, bool B_MN, bool STAGED, int CTAS, bool TF32 = false>
__global__ void __launch_bounds__(NTHREADS) dx_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmR,
                                                             const __grid_constant__ CUtensorMap tmX, TcParams p,
                                                             DxEpi e0) {
  // Persistent CTA: loops over output tiles; the TMA producer and the MMA issuer run ahead of the epilogue warps
  // through a STAGES-deep operand ring and a 2-deep ring of TMEM accumulators (2*BN columns).
  // CTAS == 2: a cluster of two CTAs (one SM pair) computes a 256 x BN tile with tcgen05.mma.cta_group::2 issued by the
  // leader (cluster rank 0).  Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), so the L2 -> SM
  // operand traffic per FLOP drops by a third against 128 x BN single-CTA tiles; each CTA's TMEM holds its 128 accumulator
  // rows and its own epilogue warps drain them.
  // CTAS == 3: the same cluster of two, but each CTA runs its own 128 x BN MMAs (cta_group::1) on its own M tile; the two
  // tiles share the N tile, and each CTA fetches HALF of the B block and multicasts it into both CTAs' rings.  This is the
  // mode of the HBM-bound shapes (K <= 1024, N = dim): their L2 -> SM traffic (B re-read by every M tile) sat at the L2
  // throughput cap.
  // TF32: fp32 operands (kind::tf32, direct epilogue, single CTA).  A k-block is still one 128 B swizzle row = 32 elements,
  // an MN-major tile is made of 32-element (128 B) column chunks of BKE rows, one MMA consumes 8 k.
  static_assert(!TF32 || (!STAGED && CTAS == 1), "tf32 instances use the direct epilogue on single CTAs");
  constexpr int BKE = TF32 ? 32 : 64;          // elements per k-block
  constexpr int CHE = TF32 ? 32 : 64;          // elements per 128 B chunk of an MN-major tile
  constexpr int CHB = BKE * 128;               // bytes of one MN-major chunk (BKE rows x 128 B)
  constexpr int MNK = TF32 ? 1024 : 2048;      // MN-major: smem advance per MMA (8 / 16 k-rows)
  constexpr bool PAIR = CTAS == 2, MC = CTAS == 3;
  constexpr int CL = CTAS == 1 ? 1 : 2;
  constexpr int BNL = PAIR ? BN / 2 : BN;   // B rows held in this CTA's ring
  constexpr int BNH = BN / 2;               // B rows FETCHED by this CTA in the two cluster modes
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BNL * BK * 2;
  constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // allocation must be a power of two
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment is required by the 128B swizzle atoms; align manually as well.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // layout: operand ring | staging blocks (1024 B aligned: TMA 128B-swizzle destinations) | bias blocks | barriers
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
  uint8_t* bias_base = stg_base + (STAGED ? NEPI * p.stage_bufs * STG_BYTES : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_base + (STAGED ? NEPI * 512 : 0));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* side_full_bar = tmem_empty_bar + 2;   // [NEPI][2]: side-tensor blocks of each epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(side_full_bar + 2 * NEPI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BKE - 1) / BKE;
  const uint32_t cta_rank = CL == 2 ? cluster_ctarank() : 0;
  int unit0, uend, ustep;
  unit_range(p, (int)blockIdx.x / CL, (int)gridDim.x / CL, unit0, uend, ustep);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, MC ? 2 : 1);   // B-multicast pair: both CTAs' MMAs must have released the stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar + s, 1);
      mbar_init(tmem_empty_bar + s, NEPI * (PAIR ? 2 : 1));   // one arrival per epilogue warp (of both CTAs of a pair)
    }
    for (int s = 0; s < 2 * NEPI; ++s) mbar_init(side_full_bar + s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CL == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything is signalled remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) ran while the
  // previous kernel of the stream was still draining; from here on its results are read.  The dependents of THIS kernel may
  // start their own prologue as soon as SMs free up (they block at the same point until this grid has completed).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;   // running k-block counter across tiles
      const uint32_t leader_full = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0;
      for (int unit = unit0; unit < uend; unit += ustep) {
        const int tile = unit / p.splits, split = unit - tile * p.splits;
        int mi, ni, z;
        tile_decode(p, tile, mi, ni, z);
        const int n0 = ni * BN + (PAIR ? (int)cta_rank * BNL : 0);
        const int m0 = mi * (BM * CL) + (int)cta_rank * BM;
        const int kb_lo = split * p.kb_per_split, kb_hi = min(num_kb, kb_lo + p.kb_per_split);
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kb * BKE;
          if (!A_MN && p.l2pf > 0 && (kb - kb_lo) % p.l2pf == 0) {
            for (int j = 0; j < p.l2pf; ++j) {
              const int kk = kb + p.l2pf + j;
              if (kk < kb_hi) tma_prefetch_l2_3d(&tmA, kk * BKE, m0, z);
            }
          }
          if (MC) {
            // own A tile locally; own half of the B block into BOTH CTAs (the other half arrives from the peer)
            mbar_arrive_expect_tx(full_bar + s, STAGE_BYTES);
            if (!A_MN) {
              tma_load_3d(sa, &tmA, full_bar + s, k0, m0, z);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_3d(sa + c * 8192, &tmA, full_bar + s, m0 + c * 64, k0, z);
            }
            if (!B_MN) {
              tma_load_3d_mcast(sb + cta_rank * (BNH * BK * 2), &tmB, full_bar + s, k0, n0 + (int)cta_rank * BNH, z, 3);
            } else {
#pragma unroll
              for (int c = 0; c < BNH / 64; ++c) {
                const int cc = (int)cta_rank * (BNH / 64) + c;
                tma_load_3d_mcast(sb + cc * 8192, &tmB, full_bar + s, n0 + cc * 64, k0, z, 3);
              }
            }
          } else if (PAIR) {
            // both CTAs' bytes are counted on the leader's barrier (its MMA thread consumes both halves)
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar + s, 2 * STAGE_BYTES);
            const uint32_t lb = leader_full + s * 8;
            if (!A_MN) {
              tma_load_3d_pair(sa, &tmA, lb, k0, m0, z);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_3d_pair(sa + c * 8192, &tmA, lb, m0 + c * 64, k0, z);
            }
            if (!B_MN) {
              tma_load_3d_pair(sb, &tmB, lb, k0, n0, z);  // box {64 k, BN/2 n, 1}
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_3d_pair(sb + c * 8192, &tmB, lb, n0 + c * 64, k0, z);
            }
          } else {
            mbar_arrive_expect_tx(full_bar + s, STAGE_BYTES);
            if (!A_MN) {
              tma_load_3d(sa, &tmA, full_bar + s, k0, m0, z);  // box {BKE k, 128 m, 1}
            } else {
#pragma unroll
              for (int c = 0; c < BM / CHE; ++c) tma_load_3d(sa + c * CHB, &tmA, full_bar + s, m0 + c * CHE, k0, z);  // {CHE m, BKE k, 1}
            }
            if (!B_MN) {
              tma_load_3d(sb, &tmB, full_bar + s, k0, n0, z);  // box {BKE k, BN n, 1}
            } else {
#pragma unroll
              for (int c = 0; c < BN / CHE; ++c) tma_load_3d(sb + c * CHB, &tmB, full_bar + s, n0 + c * CHE, k0, z);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (cta_rank == 0 || !PAIR)) {
      // ===== MMA issuer (the leader CTA of a pair issues for both) =====
      // Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
      // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29).
      constexpr uint32_t FMT = TF32 ? 2u : 1u;   // a_format / b_format: 1 = bf16, 2 = tf32
      constexpr uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * (PAIR ? 2 : 1)) >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int unit = unit0; unit < uend; unit += ustep, ++tcount) {
        const int split = unit % p.splits;
        const int kb_lo = split * p.kb_per_split, kb_hi = min(num_kb, kb_lo + p.kb_per_split);
        const uint32_t slot = tcount & 1, use = tcount >> 1;
        mbar_wait(tmem_empty_bar + slot, (use & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + slot * BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // K-major: advance 16 bf16 (8 tf32) = 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 (8) k-rows = two (one) 1024 B swizzle atoms.
            const uint64_t ad = make_smem_desc(sa + (A_MN ? k * MNK : k * 32), p.a_lbo, p.a_sbo, (TF32 && A_MN) ? 1u : 2u);
            const uint64_t bd = make_smem_desc(sb + (B_MN ? k * MNK : k * 32), p.b_lbo, p.b_sbo, (TF32 && B_MN) ? 1u : 2u);
            const uint32_t accf = (kb > kb_lo || k > 0 || ((p.fault & 1) && tcount >= 2)) ? 1u : 0u;
            if (TF32) umma_tf32(acc, ad, bd, idesc, accf);
            else if (PAIR) umma_f16_pair(acc, ad, bd, idesc, accf);
            else umma_f16(acc, ad, bd, idesc, accf);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (PAIR) umma_commit_pair(empty_bar + s);
          else if (MC) umma_commit_mcast(empty_bar + s);
          else umma_commit(empty_bar + s);
        }
        // accumulator complete
        if (PAIR) umma_commit_pair(tmem_full_bar + slot);
        else umma_commit(tmem_full_bar + slot);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w reads TMEM lanes [32*(w%4), +32); column blocks alternate between the quadrant's 2 warps =====
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;   // 0 or 1
    if constexpr (STAGED) {
      uint8_t* wstg = stg_base + (warp - 2) * p.stage_bufs * STG_BYTES;
      uint8_t* wbias = bias_base + (warp - 2) * 512;   // 2 x 64 floats
      uint64_t* sfull = side_full_bar + (warp - 2) * 2;
      if (lane == 0) {
        prefetch_tmap(&tmR);
        prefetch_tmap(&tmX);
      }
#define DX_EPI(MASKV) \
  staged_epilogue<BN, MASKV, CTAS>(e0, p, &tmR, &tmX, tmem_base, wstg, wbias, sfull, q, chalf, lane, tmem_full_bar, tmem_empty_bar)
      switch (p.epi_mask) {
        case 0: DX_EPI(0); break;
        case DX_M_BIAS: DX_EPI(DX_M_BIAS); break;
        case DX_M_RS: DX_EPI(DX_M_RS); break;
        case DX_M_RS | DX_M_BIAS | DX_M_GELU: DX_EPI(DX_M_RS | DX_M_BIAS | DX_M_GELU); break;
        case DX_M_RES | DX_M_ROWSQ: DX_EPI(DX_M_RES | DX_M_ROWSQ); break;
        case DX_M_BIAS | DX_M_RES | DX_M_ROWSQ: DX_EPI(DX_M_BIAS | DX_M_RES | DX_M_ROWSQ); break;
        case DX_M_GELUBWD: DX_EPI(DX_M_GELUBWD); break;
        case DX_M_RES | DX_M_CX: DX_EPI(DX_M_RES | DX_M_CX); break;
        default: DX_EPI(-1); break;
      }
#undef DX_EPI
    } else {
      uint32_t tcount = 0;
      const uint32_t empty_remote = CTAS == 2 ? mapa_u32(smem_u32(tmem_empty_bar), 0) : 0;
      for (int unit = unit0; unit < uend; unit += ustep, ++tcount) {
        const int tile = unit / p.splits;
        int mi, ni, z;
        tile_decode(p, tile, mi, ni, z);
        const int n0 = ni * BN;
        const int m0 = mi * (BM * CL) + (int)cta_rank * BM;
        const uint32_t slot = tcount & 1, use = tcount >> 1;
        const uint32_t acc = tmem_base + slot * BN + ((uint32_t)(q * 32) << 16);
        DxEpi e = e0;
        dx_epi_select_batch(e, z);
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < e.M;
        const DxRowConst rc = row_ok ? dx_row_const(e, m) : DxRowConst{1.f, 1.f, 0.f};
        float rs = 0.f, rd = 0.f;
        mbar_wait(tmem_full_bar + slot, use & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          float v[32];
          tmem_ld32(acc + (uint32_t)(c * 32), v);  // warp-collective
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = v[j * 8 + k];
              if (p.splits > 1) dx_epi_atomic_add8(e.out, e.ldo, m, n0 + c * 32 + j * 8, e.N, t);
              else dx_epilogue_piece(e, rc, m, n0 + c * 32 + j * 8, t, rs, rd);
            }
          }
        }
        // this warp has finished reading the accumulator: hand the TMEM slot back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(empty_remote + slot * 8);
          else mbar_arrive(tmem_empty_bar + slot);
        }
        if (row_ok) dx_epilogue_flush_row(e, m, rs, rd);
      }
    }
  }
  tc_fence_before();
  if (CL == 2) {
    cluster_sync_all();   // the peer may still be reading this CTA's operands / signalling its barriers / writing its ring
    if (warp == 1) {
      if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
      else tmem_dealloc(tmem_base, TMEM_COLS);
    }
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// Launcher (one instantiation per tile configuration)
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool A_MN, bool B_MN, bool STAGED, int CTAS = 1, bool TF32 = false>
int launch_cfg(const dx_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr,
               const CUtensorMap& tx, const TcParams& p, const DxEpi& e, cudaStream_t stream) {
  constexpr int CL = CTAS == 1 ? 1 : 2;
  const int smem = STAGES * (BM * BK * 2 + (CTAS == 2 ? BN / 2 : BN) * BK * 2) + 1024 /*align slack*/ + 512 /*barriers*/ +
                   (STAGED ? NEPI * (p.stage_bufs * STG_BYTES + 512) : 0);
  if (smem > 232448) {
    dx_set_error("dx_gemm_tc: tile config BN=%d stages=%d needs %d B of shared memory", BN, STAGES, smem);
    return DX_ERR_UNSUPPORTED;
  }
  auto kern = dx_gemm_tc_kernel<BN, STAGES, A_MN, B_MN, STAGED, CTAS, TF32>;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    DX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  TcParams pp = p;
  pp.tiles_n = dx_ceil_div(d->N, BN);
  pp.tiles_m = dx_ceil_div(d->M, BM * CL);
  pp.raster_m = pp.tiles_m < pp.tiles_n ? 1 : 0;
  if (const char* env = getenv("DX_GEMM_RASTER_M")) pp.raster_m = atoi(env) != 0;
  pp.fault = 0;
  if (const char* env = getenv("DX_GEMM_FAULT")) pp.fault = atoi(env);   // bit 0: injected bug; bits 1, 2: skip the side loads / the stores
  // measured on B200 (profiles/r02_gemm_exp_l2pf.json): prefetching 2..16 boxes ahead makes the deep-K GEMMs 10-30 % SLOWER
  // (the prefetches compete for the same per-SM TMA request capacity as the loads), so it is off; DX_GEMM_L2PF=n enables it
  pp.l2pf = 0;
  if (const char* env = getenv("DX_GEMM_L2PF")) pp.l2pf = atoi(env);
  pp.chunked = 0;
  if (const char* env = getenv("DX_GEMM_CHUNK")) pp.chunked = atoi(env) != 0;
  long long total = (long long)pp.tiles_n * pp.tiles_m * (d->batch > 1 ? d->batch : 1);
  static int dev_sms = 0;
  if (!dev_sms) {
    int dev = 0;
    DX_CUDA(cudaGetDevice(&dev));
    DX_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // SMs left free for a concurrently running collective (dx_gemm_reserve_sms): a persistent grid of exactly #SMs CTAs would
  // otherwise wait for the SMs NCCL holds and run its last CTAs as a second wave
  const int num_sms = dev_sms - dx_gemm_reserved_sms() > 16 ? dev_sms - dx_gemm_reserved_sms() : dev_sms;
  // split-K: only for pure fp32 accumulation (dW += A^T B) when the output tiles cannot fill the machine
  pp.splits = 1;
  const int num_kb = dx_ceil_div(d->K, TF32 ? 32 : BK);
  const bool pure_acc = d->accumulate && d->out_dtype == DX_F32 && !d->out2 && !d->res && !d->aux && !d->cx && !d->bias &&
                        !d->row_scale && !d->row_sumsq && !d->row_dot && d->act == DX_ACT_NONE;
  const int workers = num_sms / CL;   // CTAs, or CTA pairs
  const double eff1 = (double)total / ((double)((total + workers - 1) / workers) * workers);   // wave efficiency unsplit
  if (!STAGED && pure_acc && eff1 < 0.85 && num_kb >= 64) {
    // smallest split whose work units fill >= 85 % of whole waves (atomic traffic grows with the split), else the best
    double best = eff1;
    for (int sp = 2; sp <= 8 && num_kb / sp >= 32; ++sp) {
      const long long u = total * sp;
      const double eff = (double)u / ((double)((u + workers - 1) / workers) * workers);
      if (eff > best + 0.04) { best = eff; pp.splits = sp; }
      if (eff >= 0.85) break;
    }
  }
  pp.kb_per_split = dx_ceil_div(num_kb, pp.splits);
  pp.splits = dx_ceil_div(num_kb, pp.kb_per_split);
  total *= pp.splits;
  if (total > 0x7fffffffLL) {
    dx_set_error("dx_gemm_tc: too many tiles");
    return DX_ERR_ARG;
  }
  pp.total_tiles = (int)total;
  // programmatic stream serialization (PDL): the kernel's prologue overlaps the tail of its predecessor in the stream
  // (griddepcontrol.wait in the kernel); DX_GEMM_PDL=0 launches with full serialization
  static int pdl = -1;
  if (pdl < 0) {
    const char* env = getenv("DX_GEMM_PDL");
    pdl = env ? (atoi(env) != 0) : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CL == 2) {
    // one CTA pair (cluster of 2 = the two SMs of a TPC) per 256-row tile; persistent over min(pairs, tiles)
    const int pairs = (int)(total < workers ? total : workers);
    cfg.gridDim = dim3(2 * pairs);
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  } else {
    const int ctas_per_sm = smem <= 113 * 1024 ? 2 : 1;   // persistent grid: fill every SM, no more
    cfg.gridDim = dim3((unsigned)(total < (long long)num_sms * ctas_per_sm ? total : (long long)num_sms * ctas_per_sm));
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DX_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, tx, pp, e));
  return DX_OK;
}


// ---- instantiation lists: X(BN, STAGES, A_MN, B_MN, STAGED, CTAS) ---------------------------------------------------
#define DX_TC_1CTA_BIG(X, A, B, S) X(256, 4, A, B, S, 1) X(256, 3, A, B, S, 1) X(256, 2, A, B, S, 1)
#define DX_TC_1CTA_REST(X, A, B, S)                                                                              \
  X(192, 4, A, B, S, 1) X(192, 3, A, B, S, 1) X(128, 2, A, B, S, 1) X(128, 3, A, B, S, 1) X(128, 4, A, B, S, 1) \
      X(128, 6, A, B, S, 1) X(64, 4, A, B, S, 1) X(64, 6, A, B, S, 1)
#define DX_TC_PAIR(X, A, B, S) X(256, 6, A, B, S, 2) X(256, 4, A, B, S, 2) X(192, 6, A, B, S, 2)
#define DX_TC_MCAST(X, A, B, S) X(256, 3, A, B, S, 3) X(256, 2, A, B, S, 3)
// groups = translation units
#define DX_TC_GROUP_0(X) DX_TC_1CTA_BIG(X, false, false, false) DX_TC_1CTA_REST(X, false, false, false) \
                         DX_TC_1CTA_BIG(X, false, true, false) DX_TC_1CTA_REST(X, false, true, false)
#define DX_TC_GROUP_1(X) DX_TC_1CTA_BIG(X, true, false, false) DX_TC_1CTA_REST(X, true, false, false) \
                         DX_TC_1CTA_BIG(X, true, true, false) DX_TC_1CTA_REST(X, true, true, false) DX_TC_PAIR(X, true, true, false)
#define DX_TC_GROUP_2(X) DX_TC_1CTA_BIG(X, false, false, true)
#define DX_TC_GROUP_3(X) DX_TC_1CTA_REST(X, false, false, true)
#define DX_TC_GROUP_4(X) DX_TC_1CTA_BIG(X, false, true, true)
#define DX_TC_GROUP_5(X) DX_TC_1CTA_REST(X, false, true, true)
#define DX_TC_GROUP_6(X) DX_TC_PAIR(X, false, false, true) DX_TC_PAIR(X, false, true, true)
#define DX_TC_GROUP_7(X) DX_TC_MCAST(X, false, false, true) DX_TC_MCAST(X, false, true, true)
// fp32-operand (kind::tf32) instances: direct epilogue, single CTA, every operand layout
#define DX_TC32_LAYOUT(X, A, B) X(256, 4, A, B) X(192, 4, A, B) X(128, 6, A, B) X(64, 6, A, B)
#define DX_TC_GROUP_8(X) DX_TC32_LAYOUT(X, false, false) DX_TC32_LAYOUT(X, false, true) DX_TC32_LAYOUT(X, true, false) \
                         DX_TC32_LAYOUT(X, true, true)
#define DX_TC32_INSTANTIATE(BN, ST, A, B) template int launch_cfg<BN, ST, A, B, false, 1, true> DX_TC_SIG;
#define DX_TC32_DECLARE(BN, ST, A, B) extern template int launch_cfg<BN, ST, A, B, false, 1, true> DX_TC_SIG;
#define DX_TC_ALL_GROUPS(X) DX_TC_GROUP_0(X) DX_TC_GROUP_1(X) DX_TC_GROUP_2(X) DX_TC_GROUP_3(X) DX_TC_GROUP_4(X) \
                            DX_TC_GROUP_5(X) DX_TC_GROUP_6(X) DX_TC_GROUP_7(X)
#define DX_TC_SIG                                                                                                      \
  (const dx_gemm_desc*, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const TcParams&, \
   const DxEpi&, cudaStream_t)
#define DX_TC_INSTANTIATE(BN, ST, A, B, S, C) template int launch_cfg<BN, ST, A, B, S, C> DX_TC_SIG;
#define DX_TC_DECLARE(BN, ST, A, B, S, C) extern template int launch_cfg<BN, ST, A, B, S, C> DX_TC_SIG;

}  // namespace dx_tc
