// gemm_topk.cu -- K2: tcgen05/TMEM bf16 GEMM scoring with a fused threshold top-k epilogue.
//
// Large query batches really are a dense contraction S = Q . E^T (B x N, K = D), so this
// path runs on the 5th-generation tensor cores.  It replaces the body of HnswIndex::search
// (reference src/vector.rs:195-202, a stub) for nq >= 16 over a bf16 index.
//
// Per CTA (one per SM, 576 threads; above 128 queries two CTAs form a cta_group::2 pair):
//   * 128 queries (one UMMA M tile per CTA) stay resident in shared memory for the whole kernel
//     as K/64 swizzle-128B K-major tiles (TMA, 96 KB at D = 384);
//   * corpus tiles of 256 rows stream through a ring of 64-element k-blocks (TMA 2-D tensor map
//     over the row-major matrix, swizzle 128B): a pair's CTAs stage 128 rows each (16 KB per
//     stage, 8 stages), an independent CTA all 256 (32 KB, 4 stages);
//   * one elected thread (of a pair: the leader's) issues tcgen05.mma -- M256 N256 K16 across
//     the pair, M128 N256 K16 alone -- with fp32 accumulation into one of two 256-column TMEM
//     accumulators; tcgen05.commit frees the smem stage / publishes the tile (in both CTAs);
//   * sixteen epilogue warps read the accumulator with tcgen05.ld (lane = query, column =
//     corpus row), scale by the row's 1/norm and either
//       mode 0: keep the per-tile maximum  (threshold pass over a strided tile sample), or
//       mode 1: append (score,row) keys that reach the query's threshold to a global list.
//     The epilogue of tile t overlaps the MMAs of tile t+1 (double-buffered TMEM).
//
// The threshold of a query is the k-th largest of >= k per-tile maxima: k distinct rows score
// at least that much, so it is a valid lower bound of the global k-th best score and no row
// below it can be in the top-k.  The exact top-k is then selected from the (few thousand)
// survivors.  The N x B score matrix is never materialised.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "gemm_topk.cuh"

namespace tss {

namespace {

constexpr int kBlockM = 128;      // queries per CTA (UMMA M)
constexpr int kBlockN = 256;      // corpus rows per tile (UMMA N)
constexpr int kBlockK = 64;       // bf16 elements per k-block = 128 bytes = one swizzle atom row
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kBTileBytes = kBlockN * kBlockK * 2;  // 32 KB
constexpr int kColSplit = 4;      // epilogue warps per TMEM lane quadrant (each takes 256/4 columns)
constexpr int kEpiThreads = 128 * kColSplit;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr int kBarSlots = 24;     // full[stages] + empty[stages] + queries + 2 tmem full + 2 tmem empty
constexpr int kThreads = 64 + kEpiThreads;  // warp 0 TMA, warp 1 MMA + TMEM alloc, rest epilogue
constexpr int kTmemCols = 512;    // two 256-column fp32 accumulators

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---- cta_group::2 (a CTA pair drives one M256 MMA; SASS UTCHMMA.2CTA) ----
// shared::cluster address of `addr` (an address in this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA box into THIS CTA's shared memory whose bytes are counted on a barrier that may live in
// the pair's other CTA (the leader's, where the MMA thread waits); SASS UTMALDG.2D.2CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map,
                                                 uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same, delivered to every CTA of cta_mask at the same CTA-relative offset (a quad = two
// pairs sharing corpus tiles: the CTAs of equal parity in both pairs need the same half tile);
// each destination's bytes are counted on the leader barrier of ITS OWN pair.
// SASS UTMALDG.2D.2CTA.MULTICAST
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap* map,
                                                    uint32_t cluster_bar, int c0, int c1,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// arrives (once the MMAs issued so far retire) on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
// (default .release.cta: what it orders here are TMEM reads, which tcgen05.fence covers; the
// .release.cluster form costs a MEMBAR.ALL.GPU per arrival -- 30 % of all stall samples)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// K-major, swizzle-128B shared-memory matrix descriptor (sm_100 format):
// start address >> 4 | SBO (8 rows x 128 B = 1024) >> 4 at bit 32 | version 1 at bit 46 |
// layout SWIZZLE_128B (2) at bit 61.  LBO is unused for a single swizzle atom along K.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBlockN >> 3) << 17) |
                            ((uint32_t)(kBlockM >> 4) << 24);

// the pair's instruction: M = 256 (128 queries from each CTA), N = 256 (128 corpus rows from each)
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) |
                                ((uint32_t)(kBlockN >> 3) << 17) | ((uint32_t)((2 * kBlockM) >> 4) << 24);
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescPair), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// two fp32 multiplies in one instruction (SASS FMUL2): v = {bits(a0)*b.x, bits(a1)*b.y}
__device__ __forceinline__ void mul2(uint32_t a0, uint32_t a1, float bx, float by, float& v0,
                                     float& v1) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(rd));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace

// KB: k-blocks of 64 elements (D padded to KB*64).
// KB <= 6 (D <= 384): the 128 queries stay resident in shared memory (KB x 16 KB).  Larger D
// does not leave room for that next to the corpus ring, so the query tile of each k-block
// streams through the ring together with the corpus box (16 + 32 KB per stage, L2-resident).
// PAIR: launched as clusters of two CTAs that form one cta_group::2 pair and score two
// different query blocks against the same corpus tiles.  The leader (cluster rank 0)
// issues M256 N256 K16 MMAs (128 clocks each) whose A rows are the 128 queries of EACH CTA and
// whose B rows are the 128 corpus rows EACH CTA staged.  Shared-memory traffic per SM drops
// from 96 B/clk of operand reads (4 KB A + 8 KB B per MMA) + 64 B/clk of TMA writes (a 32 KB
// stage per 4 MMAs) -- more than the 128 B/clk an SM has -- to 64 + 32 B/clk, and the L2 -> SM
// traffic halves without any multicast: every CTA loads only its own half of the tile (16 KB
// per stage, so the ring is 8 deep) with the cta_group::2 form of the TMA load, which counts
// its bytes on the leader's barrier.  Each CTA's TMEM receives its own 128 queries x all 256
// rows, so the epilogue is the same in both modes.
// CL: CTAs per cluster -- 1 independent CTAs, 2 one cta_group::2 pair, 4 a QUAD = two pairs that
// score four query blocks against the same corpus tiles: every CTA fetches a QUARTER of each
// tile (64 rows) and multicasts it to the CTA of its parity in the other pair, so each corpus
// byte crosses L2 -> SM twice per 512 queries instead of four times.
template <int KB, int CL>
__host__ __device__ constexpr int gemm_stages() { return CL >= 2 ? (KB <= 6 ? 8 : 6) : 4; }
// floats of staged 1/|row|: a pair's epilogue warps each keep the 32 columns' worth of the chunk
// they are on (no block barrier); otherwise one double-buffered tile's worth shared by all
// epilogue warps
__host__ __device__ constexpr int gemm_ninv_floats(bool pair) {
  return pair ? kEpiWarps * 32 : 2 * kBlockN;
}

// W: the epilogue scales every score by a per-row weight (1/|row|, or NaN for a masked row).
// W = false: the matrix the tensor cores read holds UNIT rows (launch_normalize_rows) and there
// is no mask, so the accumulator already is the score: the epilogue is a bare maximum + compare
// -- no weight fetch, no staging, no multiply (the multiplies and their shared-memory loads were
// half of the epilogue's instructions, and the epilogue math 15 % of a power-capped batch).
template <int KB, int CL, bool W>
__global__ void __launch_bounds__(kThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_e,
                 const GemmParams p) {
  constexpr bool PAIR = CL >= 2;
  constexpr bool QUAD = CL == 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // swizzle-128B tiles need 1024-byte alignment; the kernel has no static shared memory, so
  // the dynamic window starts on its own allocation boundary (checked, not assumed)
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  uint8_t* smem = smem_raw;
  constexpr bool ARES = KB <= 6;                         // queries resident (else streamed)
  constexpr int kStages = gemm_stages<KB, CL>();
  constexpr int kBBytes = PAIR ? kBTileBytes / 2 : kBTileBytes;  // corpus rows staged per CTA
  constexpr int kABytes = ARES ? KB * kATileBytes : 0;
  constexpr int kStageBytes = kBBytes + (ARES ? 0 : kATileBytes);
  constexpr uint32_t kTxCtas = PAIR ? 2 : 1;             // CTAs whose bytes one full barrier counts
  const uint32_t sA = smem_base;                         // ARES: KB tiles of 16 KB
  const uint32_t sB = sA + kABytes;                      // kStages stages: [B][A 16 KB if !ARES]
  uint8_t* tail = smem + kABytes + kStages * kStageBytes;
  float* s_ninv = reinterpret_cast<float*>(tail);        // inverse row norms (gemm_ninv_floats)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + gemm_ninv_floats(PAIR) * sizeof(float));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kBarSlots);
  static_assert(2 * kStages + 6 <= kBarSlots, "barrier slots");
  const uint32_t bar_full = smem_u32(&bars[0]);        // [kStages]
  const uint32_t bar_empty = smem_u32(&bars[kStages]); // [kStages]
  const uint32_t bar_a = smem_u32(&bars[2 * kStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kStages + 1]);   // [2]
  const uint32_t bar_tempty = smem_u32(&bars[2 * kStages + 3]);  // [2]
  const uint32_t bar_afree = smem_u32(&bars[2 * kStages + 5]);   // resident queries may be replaced

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
  // ---- work assignment, balanced over ALL clusters of the launch --------------------------
  // C clusters serve G query groups of CL blocks (G = mb / CL).  F = C / G clusters belong to
  // one group each; the R = C % G left over serve EVERY group in turn (G segments, reloading
  // their queries in between).  Inside each period of C consecutive items (tiles, or sampled
  // tiles) a group's own cluster i takes items [i*G, (i+1)*G) and left-over cluster j takes
  // item F*G + j: every cluster does G items per period -- no SM idles because 148 is not a
  // multiple of the query blocks -- and all clusters sweep the corpus together, so L2 sees
  // each tile once per period.  Survivor lists are indexed by (query, slice of its group).
  const uint32_t cluster_id = blockIdx.x / CL;
  const uint32_t C = gridDim.x / CL, G = p.mb / CL, F = C / G, R = C - F * G;
  if (F == 0) __trap();  // the host launches at least one cluster per query group
  const bool left_over = cluster_id >= F * G;
  const uint32_t nseg = left_over ? G : 1u;
  const uint32_t run = left_over ? 1u : G;
  // (own clusters are numbered group-fastest: the G clusters that read the SAME tiles for the
  // G query groups have consecutive ids, so they are scheduled side by side and the second
  // reader of a tile finds it in L2 -- numbered group-slowest, 1.5x the corpus came from DRAM)
  const uint32_t own_i = cluster_id / G, own_g = cluster_id % G;
  const uint32_t item_base = left_over ? F * G + (cluster_id - F * G) : own_i * G;
  const uint32_t slice = left_over ? F + (cluster_id - F * G) : own_i;
  const uint32_t nslices = F + R;
  auto item = [&](uint32_t n) -> uint32_t { return (n / run) * C + item_base + (n % run); };
  auto seg_m_blk = [&](uint32_t seg) -> uint32_t {
    return (left_over ? seg : own_g) * CL + rank;
  };
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);   // every CTA of the cluster
  const uint32_t pair_rank = rank & 1u, pair_id = rank >> 1;  // (QUAD: which pair, which half)
  const uint32_t lead_rank = rank & ~1u;                      // this CTA's pair leader
  const uint16_t pair_mask = (uint16_t)(3u << lead_rank);     // the two CTAs of this pair
  // this cluster's items: item(0), item(1), ... < count; tile = item * stride
  const uint32_t count = p.mode == 0 ? p.sample_count : p.num_tiles;
  const uint32_t stride = p.mode == 0 ? p.sample_stride : 1u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_e);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      // a stage is refilled by loads that land in BOTH pairs of a quad: it is free only when
      // both leaders' MMAs have read it
      mbar_init(bar_empty + 8 * s, QUAD ? 2 : 1);
    }
    mbar_init(bar_a, 1);
    mbar_init(bar_afree, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      // one arrival per epilogue warp; the leader's MMA thread also waits for the peer's warps
      mbar_init(bar_tempty + 8 * a, kEpiWarps * kTxCtas);
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-wide (a pair: the same warp of both CTAs)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(s_tmem)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(s_tmem)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peer barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // ring stages in use (diagnostic knob: fewer stages = fewer bytes in flight)
  const uint32_t nst = p.ring_stages && p.ring_stages < (uint32_t)kStages ? p.ring_stages : kStages;
  // PAIR: the barriers TMA bytes and drained accumulators are reported to are the leader's
  const bool leader = !PAIR || pair_rank == 0;
  const uint32_t lead_full = PAIR ? map_to_rank(bar_full, lead_rank) : bar_full;
  const uint32_t lead_a = PAIR ? map_to_rank(bar_a, lead_rank) : bar_a;
  const uint32_t lead_tempty = PAIR ? map_to_rank(bar_tempty, lead_rank) : bar_tempty;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      // (A contiguous L2 prefetch of whole tiles ahead of these strided boxes used to live here;
      // measured with a 7- or 8-deep ring it changes nothing at 512+ queries and costs 25-40 % at
      // <= 256, where every SM streams its own tiles.)
      for (uint32_t seg = 0; seg < nseg; ++seg) {
      const uint32_t m_blk = seg_m_blk(seg);
      if (ARES) {
        // (a left-over cluster's next group: the MMAs that read the old queries have retired)
        if (seg) mbar_wait(bar_afree, (seg - 1) & 1u);
        if (leader) mbar_arrive_expect_tx(bar_a, kTxCtas * KB * kATileBytes);
        for (int kb = 0; kb < KB; ++kb) {
          if (PAIR)
            tma_load_2d_pair(sA + kb * kATileBytes, &tmap_q, lead_a, kb * kBlockK,
                             (int)(m_blk * kBlockM));
          else
            tma_load_2d(sA + kb * kATileBytes, &tmap_q, bar_a, kb * kBlockK, (int)(m_blk * kBlockM));
        }
      }
      for (uint32_t n = 0;; ++n) {
        const uint32_t i = item(n);
        if (i >= count) break;
        const int row0 = (int)(i * stride * kBlockN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (p.debug & 4) {
            if (leader) mbar_arrive(bar_full + 8 * s);
          } else if (PAIR) {
            // the leader's barrier counts both CTAs' stages: (query tile,) 128 corpus rows each
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, kTxCtas * kStageBytes);
            const uint32_t stage = sB + s * kStageBytes;
            if (!ARES)
              tma_load_2d_pair(stage + kBBytes, &tmap_q, lead_full + 8 * s, kb * kBlockK,
                               (int)(m_blk * kBlockM));
            if (QUAD)  // my quarter of the tile, into my pair AND the other pair's CTA of my parity
              tma_load_2d_pair_mc(stage + pair_id * (kBBytes / 2), &tmap_e, lead_full + 8 * s,
                                  kb * kBlockK,
                                  row0 + (int)pair_rank * (kBlockN / 2) + (int)pair_id * (kBlockN / 4),
                                  (uint16_t)(5u << pair_rank));
            else
              tma_load_2d_pair(stage, &tmap_e, lead_full + 8 * s, kb * kBlockK,
                               row0 + (int)rank * (kBlockN / 2));
          } else {
            // (query tile and) the 256 corpus rows of this k-block
            mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
            const uint32_t stage = sB + s * kStageBytes;
            if (!ARES)
              tma_load_2d(stage + kBTileBytes, &tmap_q, bar_full + 8 * s, kb * kBlockK,
                          (int)(m_blk * kBlockM));
            tma_load_2d(stage, &tmap_e, bar_full + 8 * s, kb * kBlockK, row0);
          }
          if (++s == nst) s = 0, ph ^= 1;
        }
      }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane; of a pair, the leader's) =====
    if (lane == 0 && leader) {
      uint32_t s = 0, ph = 0, it = 0;
      for (uint32_t seg = 0; seg < nseg; ++seg) {
      if (ARES) mbar_wait(bar_a, seg & 1u);
      for (uint32_t n = 0; item(n) < count; ++n, ++it) {
        const uint32_t acc = it & 1u, use = it >> 1;
        mbar_wait(bar_tempty + 8 * acc, (use & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t stage = sB + s * kStageBytes;
          // (a pair's descriptors name the same offsets in both CTAs' shared memory)
          const uint64_t adesc = make_desc(ARES ? sA + kb * kATileBytes : stage + kBBytes);
          const uint64_t bdesc = make_desc(stage);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {  // +32 bytes along K per step (>> 4 = 2)
            if (p.debug & 2) continue;
            if (PAIR) umma_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, (kb | k) != 0 ? 1u : 0u);
            else umma(tmem_d, adesc + 2 * k, bdesc + 2 * k, (kb | k) != 0 ? 1u : 0u);
          }
          // smem stage reusable once these MMAs retire (a pair's: in both CTAs; a quad's: every
          // CTA hears from both leaders)
          if (PAIR) umma_commit_pair(bar_empty + 8 * s, kMask);
          else umma_commit(bar_empty + 8 * s);
          if (++s == nst) s = 0, ph ^= 1;
        }
        // accumulator complete (a pair's: in both CTAs' TMEM)
        if (PAIR) umma_commit_pair(bar_tfull + 8 * acc, pair_mask);
        else umma_commit(bar_tfull + 8 * acc);
      }
      if (ARES && seg + 1 < nseg) {  // the resident queries may be replaced once these retire
        if (PAIR) umma_commit_pair(bar_afree, pair_mask);
        else umma_commit(bar_afree);
      }
      }
    }
  } else {
    // ===== epilogue: warps 2...  Warp w may touch TMEM lanes 32*(w%4)..+31 only, so
    // kColSplit warps share each lane quadrant and split the tile's 256 columns between them:
    // thread = (query, column part).  Several warps per scheduler hide each other's latency.
    constexpr uint32_t kColsPer = kBlockN / kColSplit;
    const uint32_t lane_base = 32u * (warp & 3);
    const uint32_t part = (uint32_t)(warp - 2) >> 2;      // which kColsPer-column part
    const uint32_t q_local = lane_base + lane;            // query within the block
    const uint32_t et = threadIdx.x - 64;                 // 0..kEpiThreads-1
    // survivors of (query, slice, part) go to a list only this thread writes: no atomics
    const uint32_t sub = slice * kColSplit + part, nsub = nslices * kColSplit;
    uint32_t it = 0;
    // 1/|row| of local row r as the epilogue wants it
    auto row_weight = [&](uint64_t r) -> float {
      if (r >= p.n_rows) return 0.f;
      float w = __ldg(p.inv_norm + r);
      if (p.mask) {
        uint32_t bit = (__ldg(p.mask + (r >> 5)) >> (uint32_t)(r & 31)) & 1u;
        if (p.mask_mode == 2) bit ^= 1u;  // TSS_MASK_EXCLUDE
        // a masked row's score becomes NaN: fmax drops it from every maximum and
        // `v >= thr` is false, at no cost in the per-element code
        if (!bit) w = __uint_as_float(0x7FC00000u);
      }
      return w;
    };
    // The weights of a tile are fetched one tile ahead into registers, so their global-memory
    // latency never sits between two tiles.  PAIR: lane l of a warp fetches the weights of
    // columns part*64 + l and + 32 + l, and the warp stages the 32 of the chunk it is about to
    // read in its own 128 bytes (warp-level sync only).  Otherwise thread et < 256 fetches column et and all epilogue
    // warps meet at a named barrier (double-buffered by accumulator).
    const uint64_t my_col = PAIR ? part * kColsPer + lane : et;
    auto fetch_weights = [&](uint32_t i, float& a, float& b) {
      if (!W || i >= count) return;
      const uint64_t r = (uint64_t)i * stride * kBlockN + my_col;
      if (PAIR) {
        a = row_weight(r);
        b = row_weight(r + 32);
      } else if (et < (uint32_t)kBlockN) {
        a = row_weight(r);
      }
    };
    float* warp_ninv = s_ninv + (warp - 2) * 32;
    for (uint32_t seg = 0; seg < nseg; ++seg) {
    const uint32_t q = seg_m_blk(seg) * kBlockM + q_local;  // query within the batch (may be >= nq)
    const float thr = (p.mode == 1) ? p.thr[q] : 0.f;
    uint64_t* my_cand = p.cand + ((size_t)q * nsub + sub) * p.cand_cap;
    uint32_t my_count = 0;
    float w_a = 0.f, w_b = 0.f;
    fetch_weights(item(0), w_a, w_b);
    for (uint32_t n = 0;; ++n, ++it) {
      const uint32_t i = item(n);
      if (i >= count) break;
      const uint32_t acc = it & 1u, use = it >> 1;
      const uint64_t row0 = (uint64_t)i * stride * kBlockN;
      // ninv[c] = weight of the tile's column c (for the columns this thread visits)
      float* ninv = s_ninv + acc * kBlockN;
      if (!W) {
      } else if (PAIR) {
        __syncwarp();  // the previous tile's reads of the warp's buffer are done
        warp_ninv[lane] = w_a;
        __syncwarp();
      } else {
        if (et < (uint32_t)kBlockN) ninv[et] = w_a;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      float n_a = 0.f, n_b = 0.f;  // the next tile's weights, in flight while this one is scored
      fetch_weights(item(n + 1), n_a, n_b);
      mbar_wait(bar_tfull + 8 * acc, use & 1u);
      tc_fence_after();
      const uint64_t left = p.n_rows - row0;
      const uint32_t ncols = left >= (uint64_t)kBlockN ? kBlockN : (uint32_t)left;
      float mx = -INFINITY;
#pragma unroll 1  // (fully unrolled the two chunks cost 20 %: 7.25 vs 6.0 ms per batch)
      for (uint32_t cb = part * kColsPer; cb < (part + 1) * kColsPer; cb += 32) {
        if (W && PAIR && cb != part * kColsPer) {  // second chunk: its weights replace the first's
          __syncwarp();
          warp_ninv[lane] = w_b;
          __syncwarp();
        }
        uint32_t r[32];
        tmem_ld32(tmem_base + (lane_base << 16) + acc * kBlockN + cb, r);
        tmem_ld_wait();
        if (cb >= ncols || (p.debug & 1)) continue;  // warp-uniform: the load above stays converged
        // branch-free common case: scale by 1/|row|, maxima of the four groups of 8 and of
        // the chunk; only a chunk (then a group) whose maximum reaches the threshold is walked
        float v[32];
        if (W) {
          const float4* nv = reinterpret_cast<const float4*>(PAIR ? warp_ninv : ninv + cb);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 w = nv[j4];
            mul2(r[4 * j4 + 0], r[4 * j4 + 1], w.x, w.y, v[4 * j4 + 0], v[4 * j4 + 1]);
            mul2(r[4 * j4 + 2], r[4 * j4 + 3], w.z, w.w, v[4 * j4 + 2], v[4 * j4 + 3]);
          }
        } else {  // unit rows: the accumulator is the score
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (cb + 32 > ncols) {  // last, partial tile of the corpus only
          // columns past the last row are marked like masked rows: NaN, which fmax drops and
          // `>= thr` rejects even when thr is -inf (fewer than k live maxima) -- -inf here
          // would pass `-inf >= -inf` and push row ids that do not exist
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cb + j >= ncols) v[j] = __uint_as_float(0x7FC00000u);
        }
        // chunk maximum as a tree of 3-input maxima (16 FMNMX3 for 32 values); the per-group
        // maxima are only formed on the rare path that found a survivor
        float t1[11];
#pragma unroll
        for (int i = 0; i < 10; ++i) t1[i] = fmaxf(fmaxf(v[3 * i], v[3 * i + 1]), v[3 * i + 2]);
        t1[10] = fmaxf(v[30], v[31]);
        const float t2a = fmaxf(fmaxf(t1[0], t1[1]), t1[2]), t2b = fmaxf(fmaxf(t1[3], t1[4]), t1[5]);
        const float t2c = fmaxf(fmaxf(t1[6], t1[7]), t1[8]), t2d = fmaxf(t1[9], t1[10]);
        const float m = fmaxf(fmaxf(fmaxf(t2a, t2b), t2c), t2d);
        if (p.mode == 0) {
          mx = fmaxf(mx, m);
        } else if (m >= thr) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float a = fmaxf(fmaxf(v[8 * g + 0], v[8 * g + 1]), fmaxf(v[8 * g + 2], v[8 * g + 3]));
            const float b = fmaxf(fmaxf(v[8 * g + 4], v[8 * g + 5]), fmaxf(v[8 * g + 6], v[8 * g + 7]));
            if (fmaxf(a, b) >= thr) {
#pragma unroll
              for (int j = 8 * g; j < 8 * g + 8; ++j) {
                if (v[j] >= thr) {
                  if (my_count < p.cand_cap)
                    my_cand[my_count] = pack_key(v[j], p.row_base + (uint32_t)(row0 + cb + j));
                  ++my_count;
                }
              }
            }
          }
        }
      }
      w_a = n_a;
      w_b = n_b;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(lead_tempty + 8 * acc);
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      // each (tile, part) is its own sample for the threshold: kColsPer distinct rows
      if (p.mode == 0)
        p.tile_max[(size_t)q * (p.sample_count * kColSplit) + (size_t)i * kColSplit + part] = mx;
    }
    if (p.mode == 1) p.cand_count[(size_t)q * nsub + sub] = my_count;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(kTmemCols)
                   : "memory");
  }
}

// ---- small kernels around it ------------------------------------------------------------
// fp32 queries -> bf16 [mb*128][kpad] (zero padded) + per-query 1/|q| (of the bf16-rounded query)
// + the rescoring margin (see select_kernel): a bound on 2 * |A(e) - B(e)| over all rows e, where
// A is what the tensor cores see and B = q . e / |e| what the scan computes from the stored row.
//   A = q_bf16 . e_t [/ |e_t|], e_t the bf16 row the tensor cores read.  Then
//   A - B = (q_bf16 - q) . u  +  q . (u - e/|e|),  u = e_t [/ |e_t|],
//   |A - B| <= dq * |u| + |q| * de,   dq = |q_bf16 - q|,  de = |u - e/|e||.
// dq is computed HERE, exactly, for this query (Cauchy-Schwarz needs nothing more) instead of its
// worst case 2^-8 |q| -- on ordinary data it is ~0.41 * 2^-8 |q|.  de: zero when the tensor cores
// read a bf16 index's own rows (u = e/|e|); for the unit-row shadow the largest rounding distance
// over all rows was recorded when the shadow was built (*shadow_err, launch_normalize_rows;
// ~0.45 * 2^-8, at most 2^-8): |u| <= 1 + de, and dividing by |e_t| moves u by at most de more
// (chord <= de (1 + de)).  Plus the fp32 accumulation errors of both paths and the rsqrt behind
// the weights (together under (D + 4) * 2^-22 |q|).  tests/test_margin_bound.py checks the bound
// in float64.
__global__ void prep_queries_kernel(const float* q, uint32_t nq, uint32_t dim, uint32_t kpad,
                                    uint32_t nq_pad, uint16_t* out, float* inv_qnorm, float* margin,
                                    const unsigned int* shadow_err) {
  const uint32_t qi = blockIdx.x;
  if (qi >= nq_pad) return;
  float ss = 0.f, sf = 0.f, sd = 0.f;
  for (uint32_t j = threadIdx.x; j < kpad; j += blockDim.x) {
    float v = (qi < nq && j < dim) ? q[(size_t)qi * dim + j] : 0.f;
    uint32_t u = __float_as_uint(v);
    u += 0x7FFFu + ((u >> 16) & 1u);
    uint16_t h = (uint16_t)(u >> 16);
    out[(size_t)qi * kpad + j] = h;
    float w = __uint_as_float((uint32_t)h << 16);
    ss = fmaf(w, w, ss);
    sf = fmaf(v, v, sf);  // (a component that rounds to inf, or a norm that overflows, makes the
                          // margin inf: every row survives, the list overflows, the scan answers)
    const float d = w - v;  // exact: both are fp32 and w is v with low mantissa bits rounded
    sd = fmaf(d, d, sd);
  }
  __shared__ float red[3][32];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    ss += __shfl_xor_sync(FULL_MASK, ss, m);
    sf += __shfl_xor_sync(FULL_MASK, sf, m);
    sd += __shfl_xor_sync(FULL_MASK, sd, m);
  }
  if ((threadIdx.x & 31) == 0)
    red[0][threadIdx.x >> 5] = ss, red[1][threadIdx.x >> 5] = sf, red[2][threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, tf = 0.f, td = 0.f;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[0][w], tf += red[1][w], td += red[2][w];
    inv_qnorm[qi] = t > 0.f ? rsqrtf(t) : 0.f;
    const float de = shadow_err ? __uint_as_float(*shadow_err) : 0.f;
    const float qn = sqrtf(tf), dq = sqrtf(td);
    // 1.002: the fp32 sums of squares above; inf/NaN inputs propagate to an infinite margin
    margin[qi] = 2.f * 1.002f *
                 (dq * (1.f + de) + qn * (de * (1.f + de) + (float)(dim + 4) * 0x1p-22f));
    if (!(tf < INFINITY)) margin[qi] = INFINITY;
  }
}

// 1/|e| of every stored bf16 row (one warp per row)
__global__ void row_inv_norm_kernel(const uint16_t* rows, uint64_t n_rows, uint32_t stride_elems,
                                    float* out) {
  const uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const uint16_t* row = rows + r * stride_elems;
  float ss = 0.f;
  for (uint32_t j = lane; j < stride_elems; j += 32) {
    float w = __uint_as_float((uint32_t)row[j] << 16);
    ss = fmaf(w, w, ss);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, m);
  if (lane == 0) out[r] = ss > 0.f ? rsqrtf(ss) : 0.f;
}

// block-wide bitonic sort (descending) of n (power of two) u64 keys in shared memory
__device__ void block_bitonic_desc(uint64_t* sk, uint32_t n) {
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        uint32_t lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        uint32_t hi = lo | stride;
        bool desc = (lo & size) == 0;
        uint64_t a = sk[lo], b = sk[hi];
        if ((a < b) == desc) {
          sk[lo] = b;
          sk[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// k-th largest (1-based) of a set of 32-bit keys by a 4-pass 8-bit radix select.  for_each(f)
// must call f(key, valid) in lock-step across each warp (same trip count in every lane; lanes
// without an element pass valid = false); it is invoked once per pass.  hist: 256 shared counters,
// s_sel: 2 shared words.  All 256 threads of the block call it.
constexpr uint32_t kRadixThreads = 256;
constexpr uint32_t kThresholdStage = 8192;  // maxima per query threshold_kernel stages in shared memory
template <class ForEach>
__device__ uint32_t block_radix_kth(uint32_t k, ForEach for_each, uint32_t* hist, uint32_t* s_sel) {
  __shared__ uint32_t s_wtot[kRadixThreads / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t prefix = 0, mask = 0, remaining = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    // Sign, exponent and leading mantissa bits are shared by almost all keys, so the first two
    // passes would hammer one counter: there a warp adds once per distinct bin.  The low bits
    // spread over the bins and plain atomics are cheaper than the match.
    if (shift >= 16) {
      for_each([&](uint32_t v, bool valid) {
        const bool mine = valid && (v & mask) == prefix;
        const uint32_t bin = mine ? (v >> shift) & 255u : 256u;
        const uint32_t peers = __match_any_sync(FULL_MASK, bin);
        if (mine && lane == (uint32_t)__ffs(peers) - 1u)
          atomicAdd(&hist[bin], (uint32_t)__popc(peers));
      });
    } else {
      for_each([&](uint32_t v, bool valid) {
        if (valid && (v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
      });
    }
    __syncthreads();
    // one bin per thread: how many keys sit in higher bins (suffix sums by shuffles)
    const uint32_t h = hist[threadIdx.x];
    uint32_t incl = h;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_down_sync(FULL_MASK, incl, d);
      if (lane + d < 32) incl += o;
    }
    if (lane == 0) s_wtot[warp] = incl;
    __syncthreads();
    uint32_t above = incl - h;
    for (uint32_t w = warp + 1; w < kRadixThreads / 32; ++w) above += s_wtot[w];
    if (above < remaining && remaining <= above + h) {  // exactly one bin holds the k-th key
      s_sel[0] = threadIdx.x;
      s_sel[1] = remaining - above;
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    mask |= 255u << shift;
    remaining = s_sel[1];
    __syncthreads();
  }
  return prefix;
}

// thr[q] = k-th largest of the `count` per-tile-part maxima of query q (+inf for padding
// queries, -inf when there are fewer than k maxima: keep everything).  tile_max is [nq_pad][count].
// margin (nullable): the survivors will be re-scored, so everything within the margin below the
// bound is collected too.
__global__ void __launch_bounds__(kRadixThreads)
threshold_kernel(const float* tile_max, uint32_t count, uint32_t nq, uint32_t k,
                 const float* margin, float* thr) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_sel[2];
  // the maxima are read from global memory once (as orderable words) when they fit: the four
  // radix passes then run out of shared memory (19.6 -> 13.3 us for a 16-query batch, which waits for
  // this kernel between its two tensor-core passes)
  __shared__ uint32_t s_vals[kThresholdStage];
  const uint32_t q = blockIdx.x;
  const float* mine = tile_max + (size_t)q * count;
  float t;
  if (q >= nq) {
    t = INFINITY;
  } else if (k > count) {
    t = -INFINITY;
  } else {
    const bool staged = count <= kThresholdStage;
    if (staged) {
      for (uint32_t i = threadIdx.x; i < count; i += kRadixThreads)
        s_vals[i] = orderable_bits(__float_as_uint(mine[i]));
      __syncthreads();
    }
    auto for_each = [&](auto f) {
#pragma unroll 4
      for (uint32_t i0 = 0; i0 < count; i0 += kRadixThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const bool valid = i < count;
        f(valid ? (staged ? s_vals[i] : orderable_bits(__float_as_uint(mine[i]))) : 0u, valid);
      }
    };
    uint32_t o = block_radix_kth(k, for_each, hist, s_sel);
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    t = __uint_as_float(u);
    if (margin) t -= margin[q];
  }
  if (threadIdx.x == 0) thr[q] = t;
}

// exact top-k of each query's survivors (nsub private lists of <= cap_s keys each).  A radix
// select over the score words finds the k-th best tensor-core score t; overflow[q] = 1 when some
// list was too short to hold its survivors, or the kept keys do not fit the sorter: the query is
// then redone by the exact scan.
//   rs.queries == null: the keys at or above t (about k of the few thousand survivors) are scaled
//     by 1/|q|, sorted and written -- the exact top-k of the bf16 x bf16 tensor-core scores.
//   rs.queries != null (default): every key within margin[q] below t is RE-SCORED with the scan
//     kernel's arithmetic (fp32 query x stored bf16 row, canonical summation order, DESIGN.md
//     section 3) and the top-k of those scores is written.  With A(e) the tensor-core score of row
//     e and B(e) the scan's, |A - B| <= margin/2 =: eps.  The k rows with A >= t have B >= t - eps,
//     so the scan's k-th best score is >= t - eps, so every row of the scan's top-k has
//     A >= t - 2 eps and is among the re-scored keys: the result is bit-identical to the scan's
//     (and the oracle's) top-k on the same bf16 index, ties included.
constexpr uint32_t kSelectSortMax = 4096;  // keys the sorter can take (dynamic shared memory)
// power of two >= 8k (the keys within the re-scoring margin of the k-th are 2-6x k on ordinary
// data), at least 1024, at most kSelectSortMax
__host__ __device__ inline uint32_t select_sort_cap(uint32_t k) {
  uint32_t c = 1024;
  while (c < 8 * k && c < kSelectSortMax) c <<= 1;
  return c;
}
constexpr uint32_t kSelectMaxLists = 1024;
constexpr uint32_t kSelectThreads = kRadixThreads;
struct Rescore {
  const float* queries;     // [nq][dim] fp32 as the caller passed them, or null
  const float* margin;      // [nq] (prep_queries_kernel)
  const void* rows;         // the STORED matrix (bf16 or fp32), stride_elems per row (k * 128)
  int rows_f32;
  uint32_t dim, stride_elems, row_base;
  uint64_t n_rows;          // rows of the shard: a key naming a row beyond them is dropped
};
// Re-score the keys sk[0..m) in place (low word = ~global row) with the scan's arithmetic
// (scan.cuh): "lane" l of 32 owns elements 4l..4l+3 of every 128-element
// stripe with four sub-accumulators, combines them ((a0+a1)+(a2+a3)), then an xor butterfly
// 16, 8, 4, 2, 1 over the 32 lane partials.  The query norm is done that way by a whole warp.
// Rows are done FOUR per warp for memory parallelism: thread t of an 8-thread group plays lanes
// t, t+8, t+16, t+24 one after the other; butterfly steps 16 and 8 pair exactly those four
// partials inside the thread ((p_t + p_t+16) + (p_t+8 + p_t+24)), steps 4, 2, 1 are shuffles
// within the group -- the same additions in the same order, so the same bits.
// s_q: the fp32 query in shared memory, zero padded to the row stride.  All kSelectThreads
// threads call it; it ends with a block barrier.
__device__ void rescore_keys(uint64_t* sk, uint32_t m, const Rescore& rs, const float* s_q) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t kWarps = kSelectThreads / 32;
  const uint32_t ns = rs.stride_elems / 128;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (uint32_t st = 0; st < ns; ++st) {
    const float4 qv = *reinterpret_cast<const float4*>(s_q + st * 128 + lane * 4);
    a0 = __fmaf_rn(qv.x, qv.x, a0);
    a1 = __fmaf_rn(qv.y, qv.y, a1);
    a2 = __fmaf_rn(qv.z, qv.z, a2);
    a3 = __fmaf_rn(qv.w, qv.w, a3);
  }
  const float sq_nq = __fsqrt_rn(butterfly_sum(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3))));
  const uint32_t grp = lane >> 3, t8 = lane & 7;
  for (uint32_t i0 = warp * 4; i0 < m; i0 += kWarps * 4) {
    const uint32_t i = i0 + grp;
    const bool live = i < m;
    const uint32_t row_g = 0xFFFFFFFFu - (uint32_t)sk[live ? i : i0];
    const size_t roff = (size_t)(row_g - rs.row_base) * rs.stride_elems;
    const uint16_t* rp = reinterpret_cast<const uint16_t*>(rs.rows) + roff;
    const float* rpf = reinterpret_cast<const float*>(rs.rows) + roff;
    float pd[4], pn[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
      const uint32_t l = t8 + 8 * j;  // the lane being played
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
      for (uint32_t st = 0; st < ns; ++st) {
        const float4 qv = *reinterpret_cast<const float4*>(s_q + st * 128 + l * 4);
        float ex, ey, ez, ew;
        if (rs.rows_f32) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(rpf + st * 128 + l * 4));
          ex = e.x, ey = e.y, ez = e.z, ew = e.w;
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(rp + st * 128 + l * 4));
          ex = __uint_as_float(u.x << 16), ey = __uint_as_float(u.x & 0xFFFF0000u);
          ez = __uint_as_float(u.y << 16), ew = __uint_as_float(u.y & 0xFFFF0000u);
        }
        d0 = __fmaf_rn(qv.x, ex, d0);
        d1 = __fmaf_rn(qv.y, ey, d1);
        d2 = __fmaf_rn(qv.z, ez, d2);
        d3 = __fmaf_rn(qv.w, ew, d3);
        n0 = __fmaf_rn(ex, ex, n0);
        n1 = __fmaf_rn(ey, ey, n1);
        n2 = __fmaf_rn(ez, ez, n2);
        n3 = __fmaf_rn(ew, ew, n3);
      }
      pd[j] = __fadd_rn(__fadd_rn(d0, d1), __fadd_rn(d2, d3));
      pn[j] = __fadd_rn(__fadd_rn(n0, n1), __fadd_rn(n2, n3));
    }
    float dot = __fadd_rn(__fadd_rn(pd[0], pd[2]), __fadd_rn(pd[1], pd[3]));  // xor 16, then 8
    float ne2 = __fadd_rn(__fadd_rn(pn[0], pn[2]), __fadd_rn(pn[1], pn[3]));
#pragma unroll
    for (int mm = 4; mm >= 1; mm >>= 1) {
      dot = __fadd_rn(dot, __shfl_xor_sync(FULL_MASK, dot, mm));
      ne2 = __fadd_rn(ne2, __shfl_xor_sync(FULL_MASK, ne2, mm));
    }
    if (live && t8 == 0) sk[i] = pack_key(finish_score(dot, sq_nq, ne2), row_g);
  }
  __syncthreads();
}

// One block per query; the kernel is a single wave of latency-bound blocks (8 per SM), so it is
// written for few dependent memory round trips: lists are walked by warps (no index search),
// four keys per lane are in flight, and the re-scoring runs 32 rows at a time.
#ifndef TSS_SELECT_MINB
#define TSS_SELECT_MINB 8
#endif
__global__ void __launch_bounds__(kSelectThreads, TSS_SELECT_MINB)
select_kernel(const uint64_t* cand, const uint32_t* cand_count, uint32_t nsub, uint32_t cap_s,
              const float* inv_qnorm, const Rescore rs, uint32_t k, uint64_t* out,
              uint32_t* overflow) {
  extern __shared__ __align__(16) uint64_t sk[];  // select_sort_cap(k) keys
  const uint32_t kSelectSort = select_sort_cap(k);
  __shared__ uint32_t s_cnt[kSelectMaxLists];
  __shared__ __align__(16) float s_q[1024];  // the fp32 query, zero padded to the row stride
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_sel[2];
  __shared__ uint32_t s_over, s_n, s_tot;
  const uint32_t q = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t kWarps = kSelectThreads / 32;
  const bool rescore = rs.queries != nullptr;
  if (threadIdx.x == 0) s_over = 0, s_n = 0, s_tot = 0;
  __syncthreads();
  {
    uint32_t tot = 0, over = 0;
    for (uint32_t sl = threadIdx.x; sl < nsub; sl += kSelectThreads) {
      uint32_t c = cand_count[(size_t)q * nsub + sl];
      if (c > cap_s) c = cap_s, over = 1;
      s_cnt[sl] = c;
      tot += c;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      tot += __shfl_xor_sync(FULL_MASK, tot, m);
      over |= __shfl_xor_sync(FULL_MASK, over, m);
    }
    if (lane == 0) {
      atomicAdd(&s_tot, tot);
      if (over) s_over = 1;
    }
  }
  if (rescore)
    for (uint32_t j = threadIdx.x; j < rs.stride_elems; j += kSelectThreads)
      s_q[j] = j < rs.dim ? rs.queries[(size_t)q * rs.dim + j] : 0.f;
  __syncthreads();
  const uint32_t cnt = s_tot;
  const uint64_t* base = cand + (size_t)q * nsub * cap_s;
  // f(key, valid) over all survivors, lock-step per warp.  Many short lists (a small batch spread
  // over every cluster: hundreds of lists of a few keys): a LANE per list, the warp stepping to
  // its longest one -- a warp per list would pay a memory round trip for two or three keys.
  // Few long lists (large batches): warp w walks lists w, w + 8, ...
  auto for_each_key = [&](auto f) {
    if (nsub >= 128) {
      for (uint32_t sl0 = warp * 32; sl0 < nsub; sl0 += kWarps * 32) {
        const uint32_t sl = sl0 + lane;
        const uint32_t c = sl < nsub ? s_cnt[sl] : 0u;
        const uint32_t cmax = __reduce_max_sync(FULL_MASK, c);
        const uint64_t* lp = base + (size_t)sl * cap_s;
        for (uint32_t j0 = 0; j0 < cmax; j0 += 4) {
          uint64_t key[4];
#pragma unroll
          for (uint32_t u = 0; u < 4; ++u) key[u] = j0 + u < c ? lp[j0 + u] : 0ull;
#pragma unroll
          for (uint32_t u = 0; u < 4; ++u)
            if (j0 + u < cmax) f(key[u], j0 + u < c);
        }
      }
      return;
    }
    for (uint32_t sl = warp; sl < nsub; sl += kWarps) {
      const uint32_t c = s_cnt[sl];
      const uint64_t* lp = base + (size_t)sl * cap_s;
      for (uint32_t j0 = 0; j0 < c; j0 += 128) {
        uint64_t key[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
          const uint32_t j = j0 + 32 * u + lane;
          key[u] = j < c ? lp[j] : 0ull;
        }
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u)
          if (j0 + 32 * u < c) f(key[u], j0 + 32 * u + lane < c);
      }
    }
  };
  // The survivors are walked in global memory ONCE when they fit the sorter's shared memory (the
  // ordinary case): they are staged in sk, and the four radix passes and the collection below
  // read them from there.  Otherwise (crowded scores) every pass walks the lists again.
  const bool staged = cnt <= kSelectSort;  // (block-uniform)
  if (staged) {
    for_each_key([&](uint64_t key, bool valid) {
      if (valid) sk[atomicAdd(&s_n, 1u)] = key;
    });
    __syncthreads();  // s_n == cnt
  }
  auto for_each_staged = [&](auto f) {
    for (uint32_t i0 = 0; i0 < cnt; i0 += kSelectThreads) {
      const uint32_t i = i0 + threadIdx.x;
      f(i < cnt ? sk[i] : 0ull, i < cnt);
    }
  };
  uint32_t kth = 0;  // orderable score word of the k-th best survivor (0: keep all)
  if (cnt > k) {
    auto for_each_score = [&](auto f) {
      if (staged) for_each_staged([&](uint64_t key, bool valid) { f((uint32_t)(key >> 32), valid); });
      else for_each_key([&](uint64_t key, bool valid) { f((uint32_t)(key >> 32), valid); });
    };
    kth = block_radix_kth(k, for_each_score, hist, s_sel);
  }
  uint32_t keep_from = kth;  // orderable score word from which keys are kept
  if (rescore && kth) {
    const uint32_t u = (kth & 0x80000000u) ? (kth & 0x7FFFFFFFu) : ~kth;
    keep_from = orderable_bits(__float_as_uint(__uint_as_float(u) - rs.margin[q]));
  }
  const float iq = inv_qnorm[q];
  auto keep = [&](uint64_t key, bool valid) {
    const uint32_t ob = (uint32_t)(key >> 32);
    if (!valid || ob < keep_from) return;
    // (belt and braces: the collect pass never emits a row beyond the shard, and re-scoring
    // one would read past the matrix)
    if ((uint64_t)(0xFFFFFFFFu - (uint32_t)key - rs.row_base) >= rs.n_rows) return;
    const uint32_t pos = atomicAdd(&s_n, 1u);
    if (pos >= kSelectSort) return;
    if (rescore) {
      sk[pos] = key;
    } else {
      uint32_t u = (ob & 0x80000000u) ? (ob & 0x7FFFFFFFu) : ~ob;
      float sc = __uint_as_float(u) * iq;
      if (!isfinite(sc)) sc = 0.f;
      if (sc == 0.f) sc = 0.f;
      sk[pos] = ((uint64_t)orderable_bits(__float_as_uint(sc)) << 32) | (key & 0xFFFFFFFFull);
    }
  };
  if (staged) {
    // compaction in place, 256 staged keys per round: every thread reads its key, the block
    // meets, the kept keys are appended from sk[0].  A round writes below 256 * (round + 1) --
    // no more keys are kept than were read -- i.e. only where every key has been read already.
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < cnt; i0 += kSelectThreads) {
      const uint32_t i = i0 + threadIdx.x;
      const uint64_t key = i < cnt ? sk[i] : 0ull;
      __syncthreads();
      keep(key, i < cnt);
    }
  } else {
    for_each_key(keep);
  }
  __syncthreads();
  uint32_t m = s_n;
  if (m > kSelectSort) m = kSelectSort, s_over = 1;  // (benign race: every writer stores 1)
  if (rescore) rescore_keys(sk, m, rs, s_q);
  uint32_t npad = 2;
  while (npad < m) npad <<= 1;
  for (uint32_t i = m + threadIdx.x; i < npad; i += blockDim.x) sk[i] = 0;
  __syncthreads();
  block_bitonic_desc(sk, npad);
  for (uint32_t e = threadIdx.x; e < k; e += blockDim.x)
    out[(size_t)q * k + e] = e < npad ? sk[e] : 0;
  if (threadIdx.x == 0) overflow[q] = s_over;
}

// Shadow prefilter for single queries on an fp32 index (tss_index_set_batch_policy(1)): the scan
// kernel has streamed the bf16 SHADOW of the matrix and left its top-kc keys (cosine of the fp32
// query with the bf16 rows, sorted) in cand[q][kc], kc > k.  With A the shadow score and B the
// score of the stored fp32 row, |A - B| <= eps := 2^-8 (1 + 2^-8) + (D + 4) * 2^-22 (cosine units:
// the unit vectors of a row and of its bf16 rounding are at most a 2^-8 chord apart, see
// prep_queries_kernel; plus both accumulations).  If the kc-th
// shadow score lies more than 2 eps below the k-th, every row of the fp32 top-k is among the kc
// candidates (the argument of select_kernel); they are re-scored from the fp32 rows and the top-k
// written.  Otherwise incomplete[q] = 1 and the host redoes the query with the fp32 scan.
constexpr uint32_t kRefineMax = 128;
__global__ void __launch_bounds__(kSelectThreads)
refine_kernel(const uint64_t* cand, uint32_t kc, const Rescore rs, const unsigned int* shadow_err,
              uint32_t k, uint64_t* out, uint32_t* incomplete) {
  // |A - B| <= eps: A = cos(q, e_n) over the unit-row shadow, B = cos(q, e); the two unit
  // vectors are a chord <= de (1 + de) apart, de the shadow's recorded rounding distance, plus
  // both fp32 accumulations
  const float de = __uint_as_float(*shadow_err);
  const float two_eps = 2.f * 1.002f * (de * (1.f + de) + (float)(rs.dim + 4) * 0x1p-22f);
  __shared__ uint64_t sk[kRefineMax];
  __shared__ __align__(16) float s_q[1024];
  const uint32_t q = blockIdx.x;
  for (uint32_t j = threadIdx.x; j < rs.stride_elems; j += kSelectThreads)
    s_q[j] = j < rs.dim ? rs.queries[(size_t)q * rs.dim + j] : 0.f;
  if (threadIdx.x < kRefineMax) sk[threadIdx.x] = threadIdx.x < kc ? cand[(size_t)q * kc + threadIdx.x] : 0;
  __syncthreads();
  const uint32_t m = (uint32_t)__syncthreads_count(threadIdx.x < kc && sk[threadIdx.x] != 0);
  if (threadIdx.x == 0) {
    uint32_t bad = 0;
    if (m == kc && m > k) {  // a full list: rows beyond it exist and must be provably too low
      auto score = [&](uint64_t key) {
        const uint32_t o = (uint32_t)(key >> 32);
        return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
      };
      bad = !(score(sk[kc - 1]) < score(sk[k - 1]) - two_eps);
    } else if (m == kc) {
      bad = 1;  // kc <= k: nothing to spare
    }
    incomplete[q] = bad;
  }
  __syncthreads();
  rescore_keys(sk, m, rs, s_q);
  block_bitonic_desc(sk, kRefineMax);  // unused slots are 0 = lowest key
  for (uint32_t e = threadIdx.x; e < k; e += kSelectThreads)
    out[(size_t)q * k + e] = e < kRefineMax ? sk[e] : 0;
}

// ---- host side ------------------------------------------------------------------------------
int gemm_col_split() { return kColSplit; }
int gemm_max_quads(int kb) {
  // clusters of four 1-CTA/SM blocks the device can keep resident at once (GPC shapes decide)
  static int cached[17] = {};
  if (kb < 0 || kb > 16) return 0;
  if (cached[kb]) return cached[kb] < 0 ? 0 : cached[kb];
  int n = 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4 * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = gemm_smem_bytes(kb, true);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaErrorInvalidValue;
#define TSS_QUAD_OCC(KBV)                                                                        \
  case KBV:                                                                                       \
    e = cudaFuncSetAttribute(gemm_topk_kernel<KBV, 4, true>,                                      \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes); \
    if (e == cudaSuccess)                                                                         \
      e = cudaOccupancyMaxActiveClusters(&n, gemm_topk_kernel<KBV, 4, true>, &cfg);               \
    break;
  switch (kb) {
    TSS_QUAD_OCC(2)
    TSS_QUAD_OCC(4)
    TSS_QUAD_OCC(6)
    TSS_QUAD_OCC(8)
    TSS_QUAD_OCC(12)
    TSS_QUAD_OCC(16)
    default: break;
  }
#undef TSS_QUAD_OCC
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cached[kb] = n > 0 ? n : -1;
  return n;
}
size_t gemm_smem_bytes(int kb, bool pair) {
  const size_t stages = pair ? (kb <= 6 ? 8 : 6) : 4;
  const size_t b = pair ? kBTileBytes / 2 : kBTileBytes;
  const size_t ring = kb <= 6 ? (size_t)kb * kATileBytes + stages * b : stages * (b + kATileBytes);
  return ring + gemm_ninv_floats(pair) * sizeof(float) + kBarSlots * 8 + 16;
}

template <int KB, int CL, bool W>
static cudaError_t launch_gemm_w(const CUtensorMap& tmap_q, const CUtensorMap& tmap_e,
                                 const GemmParams& p, int grid, size_t smem, cudaStream_t st) {
  static_assert(gemm_stages<KB, CL>() == (CL >= 2 ? (KB <= 6 ? 8 : 6) : 4), "gemm_smem_bytes");
  auto kern = gemm_topk_kernel<KB, CL, W>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmap_q, tmap_e, p);
}

template <int KB, int CL>
static cudaError_t launch_gemm_inst(const CUtensorMap& tmap_q, const CUtensorMap& tmap_e,
                                    const GemmParams& p, int grid, size_t smem, cudaStream_t st) {
  // no weights <=> unit rows and no mask (GemmParams::inv_norm == null)
  return p.inv_norm ? launch_gemm_w<KB, CL, true>(tmap_q, tmap_e, p, grid, smem, st)
                    : launch_gemm_w<KB, CL, false>(tmap_q, tmap_e, p, grid, smem, st);
}

cudaError_t launch_gemm_topk(int kb, int cluster, const CUtensorMap& tmap_q,
                             const CUtensorMap& tmap_e, const GemmParams& p, int grid,
                             cudaStream_t st) {
  const size_t smem = gemm_smem_bytes(kb, cluster != TSS_GEMM_SINGLE);
#define TSS_GEMM_CASE(KBV)                                                                   \
  case KBV:                                                                                   \
    return cluster == TSS_GEMM_QUAD   ? launch_gemm_inst<KBV, 4>(tmap_q, tmap_e, p, grid, smem, st) \
           : cluster == TSS_GEMM_PAIR ? launch_gemm_inst<KBV, 2>(tmap_q, tmap_e, p, grid, smem, st) \
                                      : launch_gemm_inst<KBV, 1>(tmap_q, tmap_e, p, grid, smem, st);
  switch (kb) {
    TSS_GEMM_CASE(2)
    TSS_GEMM_CASE(4)
    TSS_GEMM_CASE(6)
    TSS_GEMM_CASE(8)
    TSS_GEMM_CASE(12)
    TSS_GEMM_CASE(16)
    default: return cudaErrorInvalidValue;
  }
#undef TSS_GEMM_CASE
}

cudaError_t launch_prep_queries(const float* q, uint32_t nq, uint32_t dim, uint32_t kpad,
                                uint32_t nq_pad, uint16_t* out, float* inv_qnorm, float* margin,
                                const unsigned int* shadow_err, cudaStream_t st) {
  prep_queries_kernel<<<nq_pad, 128, 0, st>>>(q, nq, dim, kpad, nq_pad, out, inv_qnorm, margin,
                                              shadow_err);
  return cudaGetLastError();
}
cudaError_t launch_row_inv_norm(const void* rows, uint64_t n_rows, uint32_t stride_elems, float* out,
                                cudaStream_t st) {
  if (!n_rows) return cudaSuccess;
  const uint64_t blocks = (n_rows * 32 + 255) / 256;
  row_inv_norm_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(rows),
                                                        n_rows, stride_elems, out);
  return cudaGetLastError();
}
cudaError_t launch_threshold(const float* tile_max, uint32_t count, uint32_t nq_pad, uint32_t nq,
                             uint32_t k, const float* margin, float* thr, cudaStream_t st) {
  threshold_kernel<<<nq_pad, 256, 0, st>>>(tile_max, count, nq, k, margin, thr);
  return cudaGetLastError();
}
cudaError_t launch_select(const uint64_t* cand, const uint32_t* cand_count, uint32_t nslices,
                          uint32_t cap_s, const float* inv_qnorm, uint32_t nq, uint32_t k,
                          const float* queries, const float* margin, const void* rows,
                          bool rows_f32, uint32_t dim, uint32_t stride_elems, uint32_t row_base,
                          uint64_t n_rows, uint64_t* out, uint32_t* overflow, cudaStream_t st) {
  if (nslices > kSelectMaxLists || k > kSelectSortMax / 2 || stride_elems > 1024 ||
      stride_elems % 128)
    return cudaErrorInvalidConfiguration;
  Rescore rs{queries, margin, rows, rows_f32 ? 1 : 0, dim, stride_elems, row_base, n_rows};
  select_kernel<<<nq, kSelectThreads, select_sort_cap(k) * sizeof(uint64_t), st>>>(
      cand, cand_count, nslices, cap_s, inv_qnorm, rs, k, out, overflow);
  return cudaGetLastError();
}

cudaError_t launch_refine(const uint64_t* cand, uint32_t kc, const float* queries, const void* rows_f32,
                          uint32_t dim, uint32_t stride_elems, uint32_t row_base, uint32_t nq,
                          uint64_t n_rows, const unsigned int* shadow_err, uint32_t k, uint64_t* out,
                          uint32_t* incomplete, cudaStream_t st) {
  if (!shadow_err) return cudaErrorInvalidValue;
  if (kc > kRefineMax || k > kRefineMax || stride_elems > 1024 || stride_elems % 128)
    return cudaErrorInvalidConfiguration;
  Rescore rs{queries, nullptr, rows_f32, 1, dim, stride_elems, row_base, n_rows};
  refine_kernel<<<nq, kSelectThreads, 0, st>>>(cand, kc, rs, shadow_err, k, out, incomplete);
  return cudaGetLastError();
}

}  // namespace tss
