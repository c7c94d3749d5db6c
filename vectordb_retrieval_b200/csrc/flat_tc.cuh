// Flat (exact) scan on tcgen05 tensor cores: query x base contraction in 3xTF32 with the
// per-query running top-k' bound applied straight out of TMEM - the nq x n score matrix is
// never written to memory.
//
// Replaces the arithmetic of faiss.IndexFlat.search (reference src/algorithms/exact_search.py:78)
// and LinearSearcher.batch_search (reference src/algorithms/modular.py:336-387).
//
// One work item = (query tile, base chunk), handed out chunk-major to a persistent grid (one cluster
// per SM pair): the clusters of one wave read the same base tiles at about the same time, so the
// base streams from HBM once per wave and is served from L2 otherwise.  Per CTA: 128 queries = 128
// TMEM lanes, so each epilogue thread owns one query and sees that query's keys as a register
// stream.  Candidate pools are per (query, lineage), see the epilogue.
//   warp 0      TMA producer  : base tiles (hi, lo) -> 128B-swizzled smem ring
//   warp 1      MMA issuer    : per 32-wide k-block  hi*hi + hi*lo + lo*hi  (kind::tf32, fp32 acc in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue      : (TMEM lane quarter = warp % 4)
//                               tcgen05.ld 32 columns, key = acc + |x|^2 (queries are pre-scaled by
//                               -2, the tile's norms are bulk-copied to smem by the producer),
//                               compare against the query's bound, append the rare survivors to
//                               the (query, chunk, group) pool
// kCtaGroup == 2: two CTAs of a cluster form one 256 x 256 UMMA (cta_group::2); each loads half
// of every base tile, which halves L2->smem traffic per SM.
//
// The same pipeline runs with four other epilogues (template flags; the host side is in flat.cu):
//   kSeed   seeding pre-pass over a strided sample of base tiles: one TF32 product, no candidates,
//           the 16 smallest 32-row chunk minima per (query, item) in registers -> starting bounds
//   (redo)  P.qtile_active: a launch that runs only the query tiles flagged by the verify kernel
//   kDense  writes the key matrix; with P.dense_only it keeps no candidates (small bases: the
//           selection is a separate kernel) - otherwise the test hook for the 3xTF32 keys
//   kHam    Hamming scan: bf16 +-1 code rows (the byte geometry of a 128-float row), one kind::f16
//           MMA group per k-block, every key within the query's sampled bound appended to a
//           (segment, query) list
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vdb {

struct FlatScanParams {
  const float* norms;   // [n_pad]: |x|^2 (L2) / 0 (IP) / +inf (padding rows)
  int64_t nq;           // live queries
  int n_tiles;          // base tiles of TILE_N rows
  int tiles_per_chunk;  // base tiles per work item
  int n_chunks;         // S
  int n_pools;          // candidate pools per query (tcgen05: lineages L, SIMT: S)
  int n_qtiles;         // query tiles of 128*kCtaGroup rows
  uint64_t* trash;      // tcgen05 kernel: [nq_pad] write-only slots for keys that miss the bound
  int* handover;        // tcgen05 kernel: [n_qtiles][n_pools] hand-over counters between the chunks of a lineage
  int kb;               // k-blocks of 32 (kpad / 32)
  int kslices;          // K = 8 MMA slices that hold data: ceil(d / 8) (<= 4 * kb); the zero slices of the last k-block are not issued
  uint64_t* pools;      // [nq_pad][n_pools][pool_cap(KP)]
  int* pool_cnt;        // [nq_pad][n_pools]
  uint32_t* thr;        // [nq_pad] ordered-uint running bounds shared by all CTAs
  float* dense;         // kDense: dense keys [nq_pad][dense_ld] (test hook, and the small-base path of flat.cu) or nullptr
  int64_t dense_ld;
  int dense_only;       // kDense: write the keys and keep no candidates (the selection runs as a separate kernel)
  int tile_stride;      // base tile visited by step t is t * tile_stride (1 = every tile; > 1 = the strided sample of the pre-pass)
  const int* qtile_active;  // redo pass: only query tiles flagged here are processed (nullptr = all)
  float* seed_out;      // seeding pre-pass (kSeed): [nq_pad][n_chunks][kSeed] smallest chunk minima per (query, item)
  int seed_keep;        // ... kSeed of the variant to launch (16 or 32); host side only
  // Hamming scan (kHam, flat.cu): operands are bf16 +-1 codes (queries negated, so key = -dot and
  // ham = (nbits + key) / 2); every key within the query's bound is appended to the (segment, query) list.
  // A chunk is cut into two segments: its first ceil(n/2) tiles are drained by epilogue group 0 (warps 4..7, TMEM
  // buffer 0), the rest by group 1 (warps 8..11, buffer 1); the issuer alternates between the two tile streams.
  int ham_nbits;
  const int* ham_bound; // [nq] bound T (-1 = query off)
  uint32_t* ham_list;   // [2 * n_chunks][nq][ham_cap] (distance << 23 | row), rows ascending inside a segment and across segments
  int* ham_cnt;         // [2 * n_chunks][nq] entries offered to the list (> ham_cap: the list overflowed)
  int ham_cap;
  int dbg;              // bring-up knob (vdb_set_debug_mode): 0 normal, 2 no appends, 3 no tcgen05.ld, 5 keep the previous call's bounds, 8/9 = 0/5 + counters
};

// bring-up counters (vdb_debug_read_prof): see include/vdb_cuda.h for the slots
__device__ unsigned long long g_prof[8];

constexpr int kSeedKeep = 32;   // most chunk minima a query keeps per item of the seeding pre-pass (kSeed = 16 or 32: the largest usable rank)

namespace tc {
constexpr int kThreads = 256;
constexpr int kThreadsHam = 384;                     // the Hamming scan runs a second epilogue group (warps 8..11)
template <int kHam> constexpr int threads() { return kHam != 0 ? kThreadsHam : kThreads; }
constexpr int kStages = 3;
constexpr int kBlockRows = 128;                      // rows per operand block (A or B half)
constexpr int kBlockBytes = kBlockRows * 128;        // 16 KB: 128 rows x 32 fp32, SWIZZLE_128B
constexpr int kMaxResidentKb = 4;                    // query tile stays in smem when kpad <= 128
constexpr int kNormRingBytes = 2 * 256 * 4;        // |x|^2 of two tiles (<= 256 rows each)
constexpr int kKeyStageBytes = 0;
constexpr int kBarrierBytes = 256;

template <bool kAResident>
constexpr int smem_bytes() {
  // resident: A_hi/A_lo [4] + ring of {B_hi, B_lo};  streamed: ring of {A_hi, A_lo, B_hi, B_lo}
  return (kAResident ? 2 * kMaxResidentKb * kBlockBytes + kStages * 2 * kBlockBytes
                     : kStages * 4 * kBlockBytes) +
         kNormRingBytes + kKeyStageBytes + kBarrierBytes;   // dynamic smem must start 1024-byte aligned (checked)
}
}  // namespace tc

// kSeed: the seeding pre-pass (flat.cu).  Same pipeline, but the epilogue keeps no candidates: per
// (query, item) it tracks the kSeed (16 or 32) smallest *minima of 32-row chunks* in registers (a branch-free
// insertion network, so the pass runs at the contraction's speed) and writes them out at the end of
// the item; the r-th smallest of them over the sample is the query's starting bound for the main pass.
template <int kCtaGroup, bool kAResident, int KP, bool kDense = false, int kSeedN = 0, int kHam = 0>
__global__ void __launch_bounds__(tc::threads<kHam>(), 1)
flat_scan_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    const FlatScanParams P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_SPECIFIC__)
  using namespace ptx;
  constexpr int UMMA_M = 128 * kCtaGroup;
  constexpr int UMMA_N = 128 * kCtaGroup;     // base rows per tile (each CTA loads 128 of them)
  constexpr int TMEM_COLS = 2 * UMMA_N;       // double-buffered accumulator
  constexpr int CAP = pool_cap(KP);
  constexpr uint32_t IDESC = make_idesc_tf32(UMMA_M, UMMA_N);
  constexpr uint32_t IDESC_BF16 = kHam == 2 ? make_idesc_f16_acc16(UMMA_M, UMMA_N) : make_idesc_bf16(UMMA_M, UMMA_N);
  constexpr bool kSeed = kSeedN != 0;         // kSeedN = minima kept per (query, item): 16 or 32
  static_assert(kSeedN == 0 || kSeedN == 16 || kSeedN == 32, "seeding keeps 16 or 32 minima");
  static_assert(kHam == 0 || (kAResident && !kDense && !kSeed), "the Hamming scan uses the resident query tile");
  constexpr int kOps = kHam != 0 ? 1 : 2;      // operand arrays in use: hi only / hi and lo

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) {   // SWIZZLE_128B operand blocks need 1024-byte alignment
    if (threadIdx.x == 0) printf("vdb: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  // operand blocks
  uint8_t* a_res = smem;                                                        // [2][kMaxResidentKb][16K] (hi, lo)
  uint8_t* ring = smem + (kAResident ? 2 * tc::kMaxResidentKb * tc::kBlockBytes : 0);
  constexpr int kStageBytes = (kAResident ? 2 : 4) * tc::kBlockBytes;
  float* norm_ring = reinterpret_cast<float*>(ring + tc::kStages * kStageBytes);   // [2][UMMA_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(norm_ring) + tc::kNormRingBytes + tc::kKeyStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA          (leader's are used)
  uint64_t* empty_bar = bars + tc::kStages;        // [kStages]  MMA -> TMA          (every CTA)
  uint64_t* a_full_bar = bars + 2 * tc::kStages;   //            resident A landed   (leader)
  uint64_t* a_empty_bar = a_full_bar + 1;          //            resident A consumed (every CTA)
  uint64_t* tmem_full_bar = a_empty_bar + 1;       // [2]        MMA -> epilogue     (every CTA)
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]        epilogue -> MMA     (leader)
  uint64_t* norm_full_bar = tmem_empty_bar + 2;    // [2]        norms landed        (every CTA, local)
  uint64_t* norm_empty_bar = norm_full_bar + 2;    // [2]        norms consumed      (every CTA, local)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(norm_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCtaGroup == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int n_clusters = gridDim.x / kCtaGroup;
  const int n_items = P.n_qtiles * P.n_chunks;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&map_q_hi); prefetch_tensormap(&map_q_lo);
    prefetch_tensormap(&map_b_hi); prefetch_tensormap(&map_b_lo);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < tc::kStages; ++s) { mbar_init(full_bar + s, kCtaGroup); mbar_init(empty_bar + s, 1); }
    mbar_init(a_full_bar, kCtaGroup);
    mbar_init(a_empty_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar + b, 1); mbar_init(tmem_empty_bar + b, 4 * kCtaGroup);
      mbar_init(norm_full_bar + b, 1); mbar_init(norm_empty_bar + b, 4);
    }
    fence_barrier_init();
  }
  if (kCtaGroup == 2) cluster_sync_all();   // peer barriers must exist before any remote arrive
  if (warp == 2) tmem_alloc<kCtaGroup>(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0; int item_iter = 0; uint32_t tile_iter = 0;
      uint32_t ham_use[2] = {0u, 0u};             // kHam: uses of each norm / TMEM buffer so far
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int chunk = item / P.n_qtiles, qt = item % P.n_qtiles;
        if (P.qtile_active != nullptr && __ldg(P.qtile_active + qt) == 0) continue;   // every role skips the same items
        const int t0 = chunk * P.tiles_per_chunk;
        const int t1 = min(t0 + P.tiles_per_chunk, P.n_tiles);
        const int q_row0 = (qt * kCtaGroup + cta_rank) * 128;
        if (kAResident) {
          if (item_iter > 0) mbar_wait(a_empty_bar, (item_iter - 1) & 1);
          for (int kbi = 0; kbi < P.kb; ++kbi) {
            tma_load_2d<kCtaGroup>(a_res + kbi * tc::kBlockBytes, &map_q_hi, a_full_bar, kbi * 32, q_row0);
            if (kOps == 2)
              tma_load_2d<kCtaGroup>(a_res + (tc::kMaxResidentKb + kbi) * tc::kBlockBytes, &map_q_lo, a_full_bar, kbi * 32, q_row0);
          }
          if (is_leader) mbar_arrive_expect_tx(a_full_bar, kCtaGroup * P.kb * kOps * tc::kBlockBytes);
          else mbar_arrive_cluster(a_full_bar, 0);
        }
        const int n_item = t1 - t0, n_first = (n_item + 1) >> 1;      // kHam: tiles of the item, of its first segment
        for (int ts = 0; ts < n_item; ++ts, ++tile_iter) {
          // kHam: alternate between the item's two segments (even step: group 0 / buffer 0, odd: group 1 / buffer 1)
          const int t = kHam != 0 ? ((ts & 1) ? t0 + n_first + (ts >> 1) : t0 + (ts >> 1)) : t0 + ts;
          const int b_row0 = t * P.tile_stride * UMMA_N + cta_rank * 128;
          for (int kbi = 0; kbi < P.kb; ++kbi) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* st = ring + stage * kStageBytes;
            if (!kAResident) {
              tma_load_2d<kCtaGroup>(st + 2 * tc::kBlockBytes, &map_q_hi, full_bar + stage, kbi * 32, q_row0);
              tma_load_2d<kCtaGroup>(st + 3 * tc::kBlockBytes, &map_q_lo, full_bar + stage, kbi * 32, q_row0);
            }
            tma_load_2d<kCtaGroup>(st, &map_b_hi, full_bar + stage, kbi * 32, b_row0);
            if (kOps == 2) tma_load_2d<kCtaGroup>(st + tc::kBlockBytes, &map_b_lo, full_bar + stage, kbi * 32, b_row0);
            if (is_leader) mbar_arrive_expect_tx(full_bar + stage, kCtaGroup * (kOps == 2 ? kStageBytes : tc::kBlockBytes));
            else mbar_arrive_cluster(full_bar + stage, 0);
            if (++stage == tc::kStages) { stage = 0; phase ^= 1; }
          }
          // |x|^2 of the whole tile for this CTA's epilogue (issued after the operand loads, so the
          // wait for the previous user of this buffer never delays them)
          uint32_t nbuf = tile_iter & 1, nuse = tile_iter >> 1;       // buffer, and how often it was used before
          if constexpr (kHam != 0) { nbuf = ts & 1; nuse = ham_use[nbuf]++; }
          mbar_wait(norm_empty_bar + nbuf, (nuse & 1) ^ 1);
          mbar_arrive_expect_tx(norm_full_bar + nbuf, UMMA_N * 4);
          bulk_load_1d(norm_ring + nbuf * UMMA_N, P.norms + static_cast<int64_t>(t) * P.tile_stride * UMMA_N, UMMA_N * 4,
                       norm_full_bar + nbuf);
        }
        ++item_iter;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) ==========================
    if (is_leader) {
      int stage = 0; uint32_t phase = 0; int item_iter = 0; uint32_t tile_iter = 0;
      uint32_t ham_use[2] = {0u, 0u};
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int chunk = item / P.n_qtiles;
        if (P.qtile_active != nullptr && __ldg(P.qtile_active + item % P.n_qtiles) == 0) continue;
        const int t0 = chunk * P.tiles_per_chunk;
        const int t1 = min(t0 + P.tiles_per_chunk, P.n_tiles);
        if (kAResident) { mbar_wait(a_full_bar, item_iter & 1); tc_fence_after(); }
        for (int t = t0; t < t1; ++t, ++tile_iter) {
          uint32_t buf = tile_iter & 1, use = tile_iter >> 1;
          if constexpr (kHam != 0) { buf = (t - t0) & 1; use = ham_use[buf]++; }    // only the step parity matters here
          mbar_wait(tmem_empty_bar + buf, (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * UMMA_N;
          for (int kbi = 0; kbi < P.kb; ++kbi) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint8_t* st = ring + stage * kStageBytes;
              const uint32_t b_hi = smem_u32(st), b_lo = smem_u32(st + tc::kBlockBytes);
              const uint32_t a_hi = kAResident ? smem_u32(a_res + kbi * tc::kBlockBytes) : smem_u32(st + 2 * tc::kBlockBytes);
              const uint32_t a_lo = kAResident ? smem_u32(a_res + (tc::kMaxResidentKb + kbi) * tc::kBlockBytes)
                                               : smem_u32(st + 3 * tc::kBlockBytes);
              const uint64_t da_hi = make_sw128_kmajor_desc(a_hi), da_lo = make_sw128_kmajor_desc(a_lo);
              const uint64_t db_hi = make_sw128_kmajor_desc(b_hi), db_lo = make_sw128_kmajor_desc(b_lo);
              // slices of this k-block that hold data (d = 50: the second block has 3 of 4 - the padding
              // columns 56..63 are zero in both operands, so skipping their MMAs changes nothing but the time)
              const int ns = min(4, P.kslices - kbi * 4);
              if constexpr (kHam != 0) {   // bf16 codes: 64 elements per 128-byte block, K = 16 per MMA
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<kCtaGroup>(tmem_d, da_hi + 2 * k, db_hi + 2 * k, IDESC_BF16, (kbi | k) != 0);
              } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)   // 4 x K=8 per 32-wide block: +32 bytes inside the swizzle atom
                if (k < ns) umma_tf32<kCtaGroup>(tmem_d, da_hi + 2 * k, db_hi + 2 * k, IDESC, (kbi | k) != 0);
              }
              if constexpr (!kSeed && kHam == 0) {   // the seeding pre-pass only guesses: plain TF32 keys (1e-3 relative) do
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < ns) umma_tf32<kCtaGroup>(tmem_d, da_hi + 2 * k, db_lo + 2 * k, IDESC, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < ns) umma_tf32<kCtaGroup>(tmem_d, da_lo + 2 * k, db_hi + 2 * k, IDESC, 1);
              }
              umma_commit<kCtaGroup>(empty_bar + stage);
              if (kbi == P.kb - 1) umma_commit<kCtaGroup>(tmem_full_bar + buf);
            }
            __syncwarp();
            if (++stage == tc::kStages) { stage = 0; phase ^= 1; }
          }
        }
        if (kAResident && elect_one()) umma_commit<kCtaGroup>(a_empty_bar);
        __syncwarp();
        ++item_iter;
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue: filter keys against the running bound ==
    // Queries arrive scaled by -2, so key_j = acc_j + |x_j|^2.  The tile's norms were put into
    // shared memory by the producer (bulk copy, one barrier), every lane reads them as broadcast
    // 128-bit loads.  Per 32-column chunk a branch-free min tree decides whether any key of this
    // lane beats its bound; only then is the (rare) append code entered.  The chunk loop is a
    // runtime loop over two register buffers so the steady state stays small in the I-cache.
    const int ew = warp & 3;                       // TMEM lane quarter == warp % 4
    constexpr int NCH = UMMA_N / 32;               // 32-column chunks per tile
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    uint32_t va[32], vb[32];
    uint32_t tile_iter = 0;
    const bool prof = P.dbg >= 8;
    long long pf_wait = 0, pf_norm = 0, pf_comp = 0, pf_app = 0, pf_ncomp = 0, pf_hits = 0, pf_hitcyc = 0, pf_wait0 = 0;
    const long long pf_t0 = clock64();
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int chunk = item / P.n_qtiles, qt = item % P.n_qtiles;
      if (P.qtile_active != nullptr && __ldg(P.qtile_active + qt) == 0) continue;
      const int t0 = chunk * P.tiles_per_chunk;
      const int t1 = min(t0 + P.tiles_per_chunk, P.n_tiles);
      // Pool lineage: chunks c, c + L, c + 2L, ... of a query tile run in different waves (L * n_qtiles
      // >= number of clusters), so they share one pool per query: the successor waits for the
      // predecessor's hand-over flag and continues with its count, entries and bound.  A query's
      // bound therefore keeps tightening over ~n / L rows instead of restarting with every chunk.
      const int slot = chunk % P.n_pools, gen = chunk / P.n_pools;
      const int64_t q = static_cast<int64_t>(qt * kCtaGroup + cta_rank) * 128 + ew * 32 + lane;
      const bool live = q < P.nq;
      const int64_t pool_id = q * P.n_pools + slot;
      uint64_t* pool = P.pools + pool_id * CAP;
      uint32_t* thr_g = P.thr + q;
      int* handover = P.handover + qt * P.n_pools + slot;   // completed (warp, item) pairs of this lineage
      int cnt = 0;
      float thr = -CUDART_INF_F;
      float top[kSeedN > 0 ? kSeedN : 1];
      int ham_t = -1;
      uint32_t* ham_out = nullptr;
      // kHam: this warp's epilogue group (0: warps 4..7, 1: warps 8..11) drains one of the item's two segments
      const int n_item = t1 - t0, n_first = (n_item + 1) >> 1;
      const int grp = kHam != 0 ? ((warp - 4) >> 2) : 0;
      const int tb = kHam != 0 ? (grp != 0 ? t0 + n_first : t0) : t0;
      const int te = kHam != 0 ? (grp != 0 ? t1 : t0 + n_first) : t1;
      const int64_t ham_seg = static_cast<int64_t>(chunk) * 2 + grp;
      if constexpr (kHam != 0) {
        // keys are -dot (integers): ham <= b  <=>  key <= 2b - nbits; the filter compares with `<`
        if (live) ham_t = P.ham_bound[q];
        if (ham_t >= 0) thr = static_cast<float>(2 * ham_t - P.ham_nbits) + 0.5f;
        ham_out = P.ham_list + (ham_seg * P.nq + (live ? q : 0)) * P.ham_cap;
      } else if constexpr (kSeed) {
#pragma unroll
        for (int i = 0; i < kSeedN; ++i) top[i] = CUDART_INF_F;
      } else if (!(kDense && P.dense_only)) {
        if (gen > 0) {
          if (lane == 0) wait_counter(handover, gen * 4 * kCtaGroup);
          __syncwarp();
        }
        // redo pass: the pools of the tile's other queries keep what the main pass left in them
        if (gen > 0 || P.qtile_active != nullptr) cnt = __ldcg(P.pool_cnt + pool_id);
        if (live) thr = ld_volatile_thr(thr_g);
      }
      for (int t = tb; t < te; ++t, ++tile_iter) {
        uint32_t buf = tile_iter & 1, ph = (tile_iter >> 1) & 1;
        if constexpr (kHam != 0) { buf = grp; ph = tile_iter & 1; }   // the group owns its buffer: tile_iter counts its uses
        // the shared bound is read here and folded in after the tile: its latency hides behind the tile
        const uint32_t thr_seen = (kSeed || kHam != 0) ? 0u : ld_volatile_thr_raw(thr_g);   // branch-free; converted where it is used
        const uint32_t row0 = static_cast<uint32_t>(t * P.tile_stride) * UMMA_N;
        const float* nb = norm_ring + buf * UMMA_N;
        long long pf_a = 0, pf_b = 0;
        if (prof) pf_a = clock64();
        mbar_wait(norm_full_bar + buf, ph);
        if (prof) pf_b = clock64();
        mbar_wait(tmem_full_bar + buf, ph);
        if (prof) { pf_wait += clock64() - pf_b; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + buf * UMMA_N;
        auto release_tmem = [&]() {                // every column of this accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(tmem_empty_bar + buf);
            else mbar_arrive_cluster(tmem_empty_bar + buf, 0);
          }
        };
        if (P.dbg == 3) {   // pipeline-only timing: skip the accumulator read-out entirely
          release_tmem();
          if (lane == 0) mbar_arrive(norm_empty_bar + buf);
          continue;
        }
        if constexpr (kHam == 2) {
          // ---- fp16 accumulators: 64 columns per load (.pack::16b), keys filtered two per instruction -------------
          // keys are -dot, integers in [-nbits, nbits], exact in fp16.  Padding rows (zero codes, key 0) are NOT
          // masked here: the selection kernel drops rows >= n.
          const uint32_t thr2 = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(thr))) * 0x10001u;
          // ham = (key + nbits) / 2; adding 1024 puts the integer into the low mantissa bits of the half
          const uint32_t half2c = 0x38003800u;                                                  // (0.5, 0.5)
          const uint32_t bias2 = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(0.5f * P.ham_nbits + 1024.f))) * 0x10001u;
          auto consume16 = [&](uint32_t (&v)[32], int c) {
            const uint32_t rbase = row0 + c * 64;
            unsigned gmask = 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t m = hmin2_u32(hmin2_u32(v[g * 4 + 0], v[g * 4 + 1]), hmin2_u32(v[g * 4 + 2], v[g * 4 + 3]));
              gmask |= hany_lt2_u32(m, thr2) ? (1u << g) : 0u;
            }
            const unsigned um = __reduce_or_sync(0xffffffffu, gmask);
            if (um == 0u) return;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (um & (1u << g)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t k2 = v[g * 4 + u];
                  uint32_t hit_lo, hit_hi;
                  hlt2_u32(k2, thr2, hit_lo, hit_hi);
                  const uint32_t h2 = hfma2_u32(k2, half2c, bias2);            // low 10 bits of each half = ham
                  const uint32_t row = rbase + g * 8 + u * 2;
                  const uint32_t e_lo = (h2 << 23) | row, e_hi = ((h2 >> 16) << 23) | (row + 1u);
                  if (hit_lo != 0u && cnt < P.ham_cap) ham_out[cnt] = e_lo;
                  cnt += static_cast<int>(hit_lo);
                  if (hit_hi != 0u && cnt < P.ham_cap) ham_out[cnt] = e_hi;
                  cnt += static_cast<int>(hit_hi);
                }
              }
            }
          };
          constexpr int NCH16 = UMMA_N / 64;
          tmem_ld_32x32_pack16(taddr, va);
#pragma unroll 1
          for (int c = 0; c < NCH16; c += 2) {
            tmem_ld_wait();
            tmem_ld_32x32_pack16(taddr + (c + 1) * 64, vb);
            consume16(va, c);
            tmem_ld_wait();
            if (c + 2 < NCH16) tmem_ld_32x32_pack16(taddr + (c + 2) * 64, va);
            else release_tmem();
            consume16(vb, c + 1);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(norm_empty_bar + buf);
          continue;
        }
        tmem_ld_32x32(taddr, va);

        auto consume = [&](uint32_t (&v)[32], int c) {
          const float4* n4 = reinterpret_cast<const float4*>(nb + c * 32);
          const uint32_t rbase = row0 + c * 32;
          float gm[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 nn = n4[g];
            const float k0 = __uint_as_float(v[g * 4 + 0]) + nn.x, k1 = __uint_as_float(v[g * 4 + 1]) + nn.y;
            const float k2 = __uint_as_float(v[g * 4 + 2]) + nn.z, k3 = __uint_as_float(v[g * 4 + 3]) + nn.w;
            v[g * 4 + 0] = __float_as_uint(k0); v[g * 4 + 1] = __float_as_uint(k1);
            v[g * 4 + 2] = __float_as_uint(k2); v[g * 4 + 3] = __float_as_uint(k3);
            gm[g] = fminf(fminf(k0, k1), fminf(k2, k3));
          }
          if constexpr (kSeed) {
            float x = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
#pragma unroll
            for (int i = 0; i < kSeedN; ++i) {     // sorted insertion, branch-free
              const float lo_v = fminf(top[i], x);
              x = fmaxf(top[i], x);
              top[i] = lo_v;
            }
            return;
          }
          if constexpr (kDense) {
            if (live) {
              uint4* dst = reinterpret_cast<uint4*>(P.dense + q * P.dense_ld + rbase);   // 128-byte aligned
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (P.dense_only) return;
          }
          // Append path.  The lanes OR their 8-bit masks of 4-column groups that beat the bound (one
          // redux); the warp then tests the bits of that warp-uniform mask and visits only flagged
          // groups, where the per-key test merely predicates the store.  This role runs one warp per
          // scheduler, so branch / scoreboard latency is exposed; measured cycles per chunk with a hit:
          // divergent per-key branches ~760, one vote per group ~670, 32 straight-line predicated
          // stores ~1550, smem transpose + ballot per hitting lane ~1100, redux + switch loop ~1090.
          unsigned gmask = 0;
#pragma unroll
          for (int g = 0; g < 8; ++g) gmask |= (kHam != 0 ? gm[g] < thr : gm[g] <= thr) ? (1u << g) : 0u;   // flat: ties compete on the row id
          const unsigned um = __reduce_or_sync(0xffffffffu, gmask);
          if constexpr (kHam != 0) {
            if (um != 0u) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                if (um & (1u << g)) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    // branch-free like the flat scan's appends: the per-key test only predicates the store
                    const float key = __uint_as_float(v[g * 4 + u]);
                    const bool hit = key < thr;                        // ham <= bound (padding rows carry +inf)
                    // key + nbits = 2 * ham (an even integer): shifted by 22 it is ham << 23, the row takes bits 0..22
                    const uint32_t entry = (static_cast<uint32_t>(__float2int_rn(key) + P.ham_nbits) << 22) | (rbase + g * 4 + u);
                    if (hit && cnt < P.ham_cap) ham_out[cnt] = entry;
                    cnt += hit ? 1 : 0;
                  }
                }
              }
            }
            return;
          }
          if (um != 0u && P.dbg != 2) {
            long long th0 = 0;
            if (prof) { th0 = clock64(); pf_hits += 1; }
            pool_maintain<KP>(thr, cnt, pool, lane, thr_g);       // room for up to 32 appends per lane
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (um & (1u << g)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float key = __uint_as_float(v[g * 4 + u]);
                  const uint64_t packed = pack_key(key, rbase + g * 4 + u);
                  const bool hit = key <= thr;
                  if (hit) pool[cnt] = packed;
                  cnt += hit ? 1 : 0;
                }
              }
            }
            if (prof) { const long long dt = clock64() - th0; pf_hitcyc += dt; if (gen == 0) { pf_norm += dt; pf_wait0 += 1; } }
          }
        };

#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
          // chunk c is in flight in va: wait, start chunk c+1 into vb, filter va under that latency
          tmem_ld_wait();
          tmem_ld_32x32(taddr + (c + 1) * 32, vb);
          { const int before = cnt; consume(va, c); pf_app += max(cnt - before, 0); }
          tmem_ld_wait();
          if (c + 2 < NCH) tmem_ld_32x32(taddr + (c + 2) * 32, va);
          else release_tmem();
          { const int before = cnt; consume(vb, c + 1); pf_app += max(cnt - before, 0); }
        }
        __syncwarp();                              // all lanes are done with this tile's norms
        if (lane == 0) mbar_arrive(norm_empty_bar + buf);
        if (!kSeed && kHam == 0 && live) thr = fminf(thr, ord2f(thr_seen));
      }
      if constexpr (kHam != 0) {
        if (live) P.ham_cnt[ham_seg * P.nq + q] = cnt;
      }
      if constexpr (kSeed) {
        float4* out = reinterpret_cast<float4*>(P.seed_out + (q * P.n_chunks + chunk) * kSeedN);
#pragma unroll
        for (int i = 0; i < kSeedN; i += 4) out[i / 4] = make_float4(top[i], top[i + 1], top[i + 2], top[i + 3]);
      } else if (kHam == 0 && !(kDense && P.dense_only)) {
        __stcg(P.pool_cnt + pool_id, cnt);
        __threadfence();                           // pool entries + count before the hand-over flag
        __syncwarp();
        if (lane == 0) atomicAdd(handover, 1);
      }
    }
    if (prof) {
      if (lane == 0) {
        atomicAdd(&g_prof[0], static_cast<unsigned long long>(clock64() - pf_t0));
        atomicAdd(&g_prof[1], static_cast<unsigned long long>(pf_wait));
        atomicAdd(&g_prof[5], static_cast<unsigned long long>(pf_norm));     // hit-path cycles in generation-0 items
        atomicAdd(&g_prof[4], static_cast<unsigned long long>(pf_wait0));    // hit chunks in generation-0 items
        atomicAdd(&g_prof[6], 1ull);
        atomicAdd(&g_prof[7], static_cast<unsigned long long>(pf_hits));
        atomicAdd(&g_prof[2], static_cast<unsigned long long>(pf_hitcyc) << 0);
      }
      atomicAdd(&g_prof[3], static_cast<unsigned long long>(pf_app));
      (void)pf_ncomp; (void)pf_comp;
    }
  }

  // ===================================== teardown ===========================================
  tc_fence_before();
  if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, TMEM_COLS);
#else
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("vdb: flat_scan_tc_kernel needs sm_100a\n");
  __trap();
#endif
}

}  // namespace vdb
