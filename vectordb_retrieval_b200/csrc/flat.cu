// Flat (exact) search: planning, the CUDA-core scan (checker / fallback shape coverage),
// pool finalisation with exact re-scoring, shard merge, and the C entry points.
//   reference call sites replaced: src/algorithms/exact_search.py:38-39,78 (faiss.IndexFlat),
//   src/algorithms/modular.py:336-387 (LinearSearcher.batch_search)
#include <algorithm>
#include <cstdio>

#include "flat_tc.cuh"
#include "select.cuh"

namespace vdb {

// --------------------------------------------------------------------------------------------
// workspace layout: [thr u32 nq_pad][pool_cnt i32 nq_pad*S][pools u64 nq_pad*S*CAP]
struct FlatPlan {
  int cta_group;       // 0 = SIMT
  int n_qtiles;
  int n_tiles;
  int tile_rows;
  int n_chunks;
  int n_pools;         // pools per query
  int tiles_per_chunk;
  int slots;
  int clusters;        // tcgen05: persistent clusters launched
};

// Number of base chunks S.  Two constraints: (a) enough (query tile, chunk) items to fill the
// persistent grid evenly - pick the S with the best wave efficiency; (b) a chunk must fit the L2
// (`max_tiles`): the clusters of a wave read the same chunk, and only while it is L2-resident does
// each base tile leave HBM once per wave instead of once per query tile (at 100M x 128 an
// unbounded chunk of 9 GB let the clusters drift apart and the scan became HBM-bound).
static int pick_chunks(int n_qtiles, int n_tiles, int slots, int waves, int max_tiles) {
  int s_max = std::max(1, (slots * waves + n_qtiles - 1) / n_qtiles);
  s_max = std::min(s_max, std::max(slots, 16));   // few query tiles: one chunk per slot is enough
  const int s_min = std::min(std::max(1, n_tiles), (n_tiles + max_tiles - 1) / max_tiles);
  s_max = std::max(s_max, s_min + s_min / 4 + 1);
  s_max = std::min(s_max, std::max(1, n_tiles));
  int best = s_min;
  double best_eff = -1.0;
  for (int s = s_min; s <= s_max; ++s) {
    const long items = static_cast<long>(n_qtiles) * s;
    const long rounds = (items + slots - 1) / slots;
    const double eff = static_cast<double>(items) / static_cast<double>(rounds * slots);
    if (eff >= best_eff - 1e-9) { best_eff = std::max(eff, best_eff); best = s; }
  }
  return best;
}

static int max_pools(int64_t nq, int sm) {
  // upper bound of the pools per query over every implementation, independent of the base size
  const int qt128 = static_cast<int>((nq + 127) / 128), qt64 = static_cast<int>((nq + 63) / 64);
  const int a = sm / qt128 + 3;                                      // tcgen05: lineages of one query tile
  const int b = std::max(1, (2 * sm * 8 + qt64 - 1) / qt64);         // SIMT chunks (slots = 2*sm)
  return std::min(std::max(a, b), 2 * sm) + 2;
}

// Clusters of the persistent tcgen05 grid that can be CO-RESIDENT on the current device.  The lineage hand-over
// (flat_tc.cuh) makes a running cluster wait for an item owned by another cluster of the same grid, which is only
// deadlock-free when every cluster of the grid is on an SM at the same time; the raw SM count over-states that when
// SMs are shared or withheld (MPS, green contexts, unpaired SMs), so the occupancy calculator has the last word.
static int coresident_clusters(int cta_group, bool a_resident, int sm);

static FlatPlan make_plan(int impl, int64_t nq, int64_t n_pad, int kpad, int sm) {
  FlatPlan p{};
  if (impl == VDB_IMPL_SIMT) {
    p.cta_group = 0; p.tile_rows = 64; p.slots = 2 * sm;
    p.n_qtiles = static_cast<int>((nq + 63) / 64);
  } else {
    p.cta_group = impl == VDB_IMPL_TCGEN05_1CTA ? 1 : 2;
    p.tile_rows = 128 * p.cta_group;
    p.slots = std::max(1, std::min(sm / p.cta_group, coresident_clusters(p.cta_group, kpad / 32 <= tc::kMaxResidentKb, sm)));
    p.n_qtiles = static_cast<int>((nq + p.tile_rows - 1) / p.tile_rows);
  }
  p.n_tiles = static_cast<int>(n_pad / p.tile_rows);
  const int64_t tile_bytes = static_cast<int64_t>(p.tile_rows) * kpad * 8;        // hi + lo
  const int max_tiles = p.cta_group == 0 ? (1 << 30) : static_cast<int>(std::max<int64_t>(8, (48ll << 20) / tile_bytes));
  const int s = pick_chunks(p.n_qtiles, p.n_tiles, p.slots, 8, max_tiles);
  p.tiles_per_chunk = (p.n_tiles + s - 1) / s;
  p.n_chunks = (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  p.n_pools = p.n_chunks;
  if (p.cta_group != 0) {
    // chunks c and c + L of a query tile are >= one wave apart when L * n_qtiles >= clusters: they
    // share a pool (lineage), see flat_tc.cuh
    p.clusters = std::max(1, std::min(p.slots, p.n_qtiles * p.n_chunks));
    p.n_pools = std::min(p.n_chunks, (p.clusters + p.n_qtiles - 1) / p.n_qtiles);
  }
  return p;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
// Bring-up / measurement knobs are PER THREAD: a debug mode or a seeding override set by one host thread (a test, a
// profiling script) cannot change the results of an index another thread is searching.
static thread_local int g_debug_mode = 0;

// Seeded bounds (tcgen05 path, large shards).  A lineage that starts with an infinite bound accepts
// every key until its pool has seen ~k'/p rows for a hit rate p - roughly the first 200k rows of
// every lineage are epilogue-bound.  So a pre-pass scans a strided sample of S base tiles (kSeed
// variant of the scan kernel: no pools, just the smallest chunk minima per query in registers) and
// takes about the r-th smallest sampled key of every query as that query's *guess* tau (expected
// rank in the whole shard: r * N / S); the main pass starts from tau, so hits are rare from the
// first tile on.  tau is not a proven bound: the verify kernel counts the candidates of
// every query after the main pass, and a query with fewer than k' of them (tau was too tight) is
// reset and re-scanned from an infinite bound by a third launch that skips every query tile
// without such a query (normally all of them).  S is capped so that r * N / S >= margin * k':
// for a random sample P(fewer than k' rows below tau) = P(Poisson(r / margin) >= r) per query - and ONE such
// query costs a whole wave of the redo launch (its query tile's items run as long as any item of the main pass),
// so the product with 10k queries has to stay far below 1:
//     r = 16, margin 4: 5e-6      r = 16, margin 3: 1e-4 (a redo in most 10k-query searches: 2.4 ms instead of
//     r = 32, margin 3: 7e-8                               1.46 ms on a 125k-row shard - measured, rejected)
// The rank is chosen per shard (g_pre_rank = 0): 16 with margin 4 while the 64-tile sample is the binding limit
// (large shards; the guess's expected rank is 16 N / 64 tiles), 32 with margin 3 where the margin cap binds and
// that gives a clearly tighter guess (shards under ~200k rows: the sample grows from 3 % to 8 % of the rows, and
// the main scan of a 125k-row shard drops from 1.33 to 1.18 ms).
static thread_local int g_pre_tiles = 64;     // sample size in base tiles (vdb_flat_set_seeding)
static thread_local int g_pre_rank = 0;       // r (1..kSeedKeep); 0 = per shard, see above
static thread_local int g_pre_margin = 0;     // r * N / S >= margin * k'; 0 = default: 3 at r >= 32, else max(4, 64 / r)
__device__ unsigned long long g_redo_queries;

struct PrePlan {
  bool on;
  int stride;          // sampled tile t is base tile t * stride
  int rank;            // tau = the rank-th smallest sampled chunk minimum
  FlatPlan plan;       // decomposition of the sample
};

// bench.py's roofline leg: scan-kernel durations measured with CUDA events on the launching stream
constexpr int kTimingSlots = 512;
static thread_local bool g_timing_on = false;
static thread_local int g_timing_n = 0;
static thread_local cudaEvent_t g_ev0[kTimingSlots], g_ev1[kTimingSlots];
static thread_local bool g_ev_made = false;

// --------------------------------------------------------------------------------------------
__global__ void flat_init_kernel(uint32_t* thr, int64_t nq_pad, int* pool_cnt, int64_t n_cnt, int* handover,
                                 int64_t n_hand, int keep_thr, int* active, int64_t n_active) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nq_pad && !keep_thr) thr[i] = f2ord(kOpenBound);
  if (i < n_cnt) pool_cnt[i] = 0;
  if (i < n_hand) handover[i] = 0;
  if (i < n_active) active[i] = 0;
}

static int default_margin(int r) { return g_pre_margin > 0 ? g_pre_margin : (r >= 32 ? 3 : std::max(4, 64 / r)); }

static PrePlan make_pre_plan(int impl, int64_t nq, const FlatPlan& main_plan, int kpad, int sm, int kp) {
  PrePlan pp{};
  if (main_plan.cta_group == 0 || g_pre_tiles <= 0) return pp;
  auto sample_for = [&](int r) {      // tiles: the configured sample, capped so that the guess keeps its margin
    const int64_t cap = static_cast<int64_t>(r) * main_plan.n_tiles / (static_cast<int64_t>(default_margin(r)) * kp);
    return static_cast<int>(std::min<int64_t>(g_pre_tiles, cap));
  };
  int r = std::min(g_pre_rank, kSeedKeep);
  if (r <= 0) {                       // per shard: the guess's expected rank in the shard is r * n_tiles / S rows-per-tile units
    const int s16 = sample_for(16), s32 = sample_for(32);
    r = (s32 >= 4 && (s16 < 4 || 32.0 * s16 < 0.85 * 16.0 * s32)) ? 32 : 16;
  }
  const int s_tiles = sample_for(r);
  if (s_tiles < 4) return pp;                       // small shard (< ~50k rows at k' = 128): the sample would not pay for itself
  pp.on = true;
  pp.rank = r;
  pp.stride = main_plan.n_tiles / s_tiles;
  pp.plan = make_plan(impl, nq, static_cast<int64_t>(s_tiles) * main_plan.tile_rows, kpad, sm);
  return pp;
}

// tau[q] = the `rank`-th smallest of the chunk minima the pre-pass kept for query q (kSeedKeep per
// item, so the union holds the kSeedKeep smallest of the whole sample), as the ordered-uint image
// the scan compares with.  A chunk minimum is a real key, so tau's rank among the sampled rows is
// at least `rank`: slightly looser than the exact order statistic, never tighter.
__global__ void __launch_bounds__(128)
flat_tau_kernel(const float* __restrict__ seed, int n_chunks, int keep, int64_t nq, int rank, int force_fail,
                uint32_t* __restrict__ thr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= nq) return;
  const int total = n_chunks * keep;
  const float* v = seed + q * total;
  uint64_t best[1] = {kEmpty};
  for (int base = 0; base < total; base += 128) {            // four independent loads in flight per lane
    float x[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = base + t * 32 + lane;
      x[t] = i < total ? __ldcg(v + i) : CUDART_INF_F;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = base + t * 32 + lane;
      if (base + t * 32 >= total) break;                        // warp-uniform
      uint64_t fresh[1] = {i < total ? pack_key(x[t], static_cast<uint32_t>(i)) : kEmpty};
      warp_merge_keep<1>(best, fresh, lane);
    }
  }
  const uint64_t w = __shfl_sync(0xffffffffu, best[0], rank - 1);
  if (lane == 0) {
    uint32_t t = w == kEmpty ? f2ord(kOpenBound) : min(static_cast<uint32_t>(w >> 32), f2ord(kOpenBound));   // never +inf: padding rows must not pass
    if (force_fail) t = f2ord(-CUDART_INF_F);       // test hook: every query must take the redo pass
    thr[q] = t;
  }
}

// After the main pass: a query whose pools hold fewer than kp candidates in total was started from
// too tight a guess.  Such queries are reset (empty pools, infinite bound) and their query tile is
// flagged for the redo pass; every other query gets the bound -inf, so the redo pass - which runs
// whole query tiles - cannot append to its pools again.
__global__ void flat_verify_kernel(int* pool_cnt, int n_pools, int64_t nq, int64_t nq_pad, int kp, int tile_rows,
                                   uint32_t* thr, int* qtile_active, int* handover, int64_t n_hand) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_hand) handover[i] = 0;
  if (i >= nq_pad) return;
  bool failed = false;
  if (i < nq) {
    int total = 0;
    for (int s = 0; s < n_pools; ++s) total += pool_cnt[i * n_pools + s];
    failed = total < kp;
  }
  if (failed) {
    for (int s = 0; s < n_pools; ++s) pool_cnt[i * n_pools + s] = 0;
    thr[i] = f2ord(kOpenBound);
    qtile_active[i / tile_rows] = 1;
    atomicAdd(&g_redo_queries, 1ull);
  } else {
    thr[i] = f2ord(-CUDART_INF_F);
  }
}

// --------------------------------------------------------------------------------------------
// CUDA-core scan with the same key / pool / bound protocol as the tcgen05 kernel.  64 queries x
// 64 base rows per step, fp32 FFMA on x = hi + lo (exact fp32 inputs).  Used as the on-device
// checker for the tensor-core kernel and for debugging; correctness first.
template <int KP>
__global__ void __launch_bounds__(256)
flat_scan_simt_kernel(const float* __restrict__ b_hi, const float* __restrict__ b_lo,
                      const float* __restrict__ q_hi, const float* __restrict__ q_lo, int kpad,
                      const FlatScanParams P) {
  constexpr int CAP = pool_cap(KP);
  __shared__ float Qs[64][33];
  __shared__ float Bs[64][33];
  __shared__ float Ss[64][65];
  const int tid = threadIdx.x, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
  const int qt = blockIdx.x, chunk = blockIdx.y;
  const int t0 = chunk * P.tiles_per_chunk;
  const int t1 = min(t0 + P.tiles_per_chunk, P.n_tiles);
  const int64_t q_row0 = static_cast<int64_t>(qt) * 64;

  const int64_t q = q_row0 + tid;              // epilogue role: threads 0..63 own one query each
  const bool epi = tid < 64;
  const bool live = epi && q < P.nq;
  uint64_t* pool = epi ? P.pools + (q * P.n_pools + chunk) * CAP : nullptr;
  uint32_t* thr_g = epi ? P.thr + q : nullptr;
  int cnt = 0;
  float thr = live ? ld_volatile_thr(thr_g) : -CUDART_INF_F;

  for (int t = t0; t < t1; ++t) {
    const int64_t b_row0 = static_cast<int64_t>(t) * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < kpad; k0 += 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * 256, r = idx >> 5, c = idx & 31;
        const int64_t qo = (q_row0 + r) * kpad + k0 + c, bo = (b_row0 + r) * kpad + k0 + c;
        Qs[r][c] = q_hi[qo] + q_lo[qo];
        Bs[r][c] = b_hi[bo] + b_lo[bo];
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = Qs[ty * 4 + i][kk]; b[i] = Bs[tx * 4 + i][kk]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ss[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();
    if (epi) {   // warps 0 and 1, fully populated
      if (live) thr = fminf(thr, ld_volatile_thr(thr_g));
      for (int c = 0; c < 2; ++c) {
        pool_maintain<KP>(thr, cnt, pool, lane, thr_g);
        for (int i = 0; i < 32; ++i) {
          const uint32_t row = static_cast<uint32_t>(b_row0) + c * 32 + i;
          const float key = Ss[tid][c * 32 + i] + __ldg(P.norms + row);   // queries carry the -2
          if (P.dense != nullptr && live) P.dense[q * P.dense_ld + row] = key;
          if (key <= thr) { pool[cnt] = pack_key(key, row); ++cnt; }
        }
      }
    }
    __syncthreads();
  }
  if (epi) P.pool_cnt[q * P.n_pools + chunk] = cnt;
}

// --------------------------------------------------------------------------------------------
// Finalisation: one warp per query.  Streams the query's pools through a selection pool in shared memory
// (radix-select compaction, no sorting: ~400 instructions per cut where the first version merged with 256-element
// bitonic sorts, 2 600 each), re-scores the KP survivors exactly (fp32 inputs x = hi + lo, fp64 accumulation;
// difference form for L2, like LinearSearcher) in place, then sorts once by (distance, row) and writes the best k.
// The candidate list lives in shared memory, not registers, so the gather loop runs at 8 blocks per SM for KP <= 128
// (the kernel is bound by the latency of its dependent gather rounds: occupancy is what hides it).
template <int KP>
__global__ void __launch_bounds__(128, KP <= 128 ? 8 : (KP == 256 ? 4 : 2))
flat_finalize_kernel(int metric, const float* __restrict__ b_hi, const float* __restrict__ b_lo, int kpad,
                     int64_t n, int64_t id_offset, const float* __restrict__ q_hi, const float* __restrict__ q_lo,
                     int64_t nq, int n_chunks, const uint64_t* __restrict__ pools, const int* __restrict__ pool_cnt,
                     int k, int flags, float pad_value, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  constexpr int CAP = pool_cap(KP);
  constexpr int E = KP / 32;
  extern __shared__ __align__(16) uint8_t smem_fin[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= nq) return;
  WarpTopK<KP> sel;
  sel.init(reinterpret_cast<uint64_t*>(smem_fin) + warp * CAP);

  // ---- selection: every pool entry is offered once; four coalesced loads in flight
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int c = min(pool_cnt[q * n_chunks + chunk], CAP);
    const uint64_t* p = pools + (q * n_chunks + chunk) * CAP;
    for (int base = 0; base < c; base += 128) {
      uint64_t w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = base + t * 32 + lane;
        w[t] = i < c ? __ldcg(p + i) : kEmpty;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (base + t * 32 >= c) break;                          // warp-uniform
        sel.push_packed(base + t * 32 + lane < c, w[t], lane);
      }
    }
  }
  if (sel.cnt > KP) sel.compact(lane);
  __syncwarp();
  const int m = min(sel.cnt, KP);
  uint64_t* list = sel.pool;                                    // m unsorted candidates

  // ---- exact re-scoring in place: the warp walks the candidates four at a time (four independent row gathers in
  // flight), each row read coalesced over the lanes with 128-bit loads
  const float* qh = q_hi + q * kpad;
  const float* ql = q_lo + q * kpad;
  const bool one_pass = kpad <= 128;                            // the query fits one float4 per lane: keep it in registers
  float xq0[4] = {0.f, 0.f, 0.f, 0.f};
  if (one_pass && lane * 4 < kpad) {
    const float4 a4 = *reinterpret_cast<const float4*>(qh + lane * 4), b4 = *reinterpret_cast<const float4*>(ql + lane * 4);
    xq0[0] = -0.5f * (a4.x + b4.x); xq0[1] = -0.5f * (a4.y + b4.y); xq0[2] = -0.5f * (a4.z + b4.z); xq0[3] = -0.5f * (a4.w + b4.w);
  }
  for (int g0 = 0; g0 < m; g0 += 4) {
    uint32_t row[4];
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      row[u] = g0 + u < m ? packed_row(list[g0 + u]) : 0u;      // broadcast reads
      acc[u] = 0.0;
    }
    for (int j = lane * 4; j < kpad; j += 128) {
      float xq[4];
      if (one_pass) {
#pragma unroll
        for (int t = 0; t < 4; ++t) xq[t] = xq0[t];
      } else {
        const float4 a4 = *reinterpret_cast<const float4*>(qh + j), b4 = *reinterpret_cast<const float4*>(ql + j);
        xq[0] = -0.5f * (a4.x + b4.x); xq[1] = -0.5f * (a4.y + b4.y); xq[2] = -0.5f * (a4.z + b4.z); xq[3] = -0.5f * (a4.w + b4.w);
      }                                                         // operands hold -2q (exact scaling)
      float4 h4[4], l4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        h4[u] = __ldg(reinterpret_cast<const float4*>(b_hi + static_cast<int64_t>(row[u]) * kpad + j));
        l4[u] = __ldg(reinterpret_cast<const float4*>(b_lo + static_cast<int64_t>(row[u]) * kpad + j));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xb[4] = {h4[u].x + l4[u].x, h4[u].y + l4[u].y, h4[u].z + l4[u].z, h4[u].w + l4[u].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (metric == VDB_METRIC_L2) {
            const float df = xb[t] - xq[t];
            acc[u] += static_cast<double>(df) * static_cast<double>(df);
          } else {
            acc[u] += static_cast<double>(xq[t]) * static_cast<double>(xb[t]);
          }
        }
      }
    }
    __syncwarp();                                               // every lane has read this group's rows from the list
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double a = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      const float key = metric == VDB_METRIC_L2 ? static_cast<float>(a) : -static_cast<float>(a);
      if (lane == u && g0 + u < m) list[g0 + u] = pack_key(key, row[u]);
    }
  }
  __syncwarp();

  // ---- one sort by (distance, row), then the output conventions
  uint64_t exact[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    exact[e] = i < m ? list[i] : kEmpty;
  }
  warp_sort_noinline<E>(exact, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int r = e * 32 + lane;
    if (r < k) {
      float dv = pad_value;
      int64_t iv = -1;
      if (exact[e] != kEmpty) {
        const float key = packed_key(exact[e]);
        iv = static_cast<int64_t>(packed_row(exact[e])) + id_offset;
        if (metric == VDB_METRIC_L2) dv = (flags & VDB_OUT_SQRT) ? sqrtf(key) : key;
        else dv = (flags & VDB_OUT_NEGATE) ? key : ((flags & VDB_OUT_ONE_MINUS) ? 1.f + key : -key);
      }
      out_d[q * k + r] = dv;
      out_i[q * k + r] = iv;
    }
  }
}

// --------------------------------------------------------------------------------------------
// Merge of per-shard sorted lists after the allgather: one warp per query, (distance, id) order.
template <int KP>
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ d_all, const int64_t* __restrict__ i_all, int64_t stride_d, int64_t stride_i,
                  int parts, int64_t nq, int k, int descending, float pad_value, float* __restrict__ out_d,
                  int64_t* __restrict__ out_i) {
  // Candidates are keyed (value, position) with position = part*k + rank.  Parts arrive in
  // ascending id-range order (row-sharded base, allgather in rank order) and every part is
  // already sorted by (value, id), so (value, position) order IS (value, id) order: the merged
  // list does not depend on how many parts the base was split into.
  constexpr int E = KP / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= nq) return;
  uint64_t best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = kEmpty;
  const int total = parts * k;
  for (int base = 0; base < total; base += KP) {
    uint64_t fresh[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int pos = base + e * 32 + lane;
      uint64_t w = kEmpty;
      if (pos < total) {
        const int p = pos / k, r = pos % k;
        const int64_t off = q * k + r;
        if (i_all[p * stride_i + off] >= 0) {
          const float val = d_all[p * stride_d + off];
          w = pack_key(descending ? -val : val, static_cast<uint32_t>(pos));
        }
      }
      fresh[e] = w;
    }
    warp_merge_keep<E>(best, fresh, lane);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int r = e * 32 + lane;
    if (r < k) {
      float dv = pad_value;
      int64_t iv = -1;
      if (best[e] != kEmpty) {
        const int pos = static_cast<int>(packed_row(best[e]));
        const int64_t off = q * k + pos % k;
        dv = d_all[(pos / k) * stride_d + off];
        iv = i_all[(pos / k) * stride_i + off];
      }
      out_d[q * k + r] = dv;
      out_i[q * k + r] = iv;
    }
  }
}

// --------------------------------------------------------------------------------------------
// Small bases (a coarse quantiser's centroids, the 10k-20k row sets of the reference's published
// runs): every lineage would spend its whole life in the loose-bound phase - with 4 096 rows and
// k' = 32 the pools of a query were compacted ~20 times, 0.5 ms for what is 40 us of contraction.
// There the nq x n key matrix is small enough to exist: the scan writes it (dense-only epilogue,
// 128-bit stores) and this kernel selects per query with a warp pool in shared memory, leaving
// the k' smallest (key, row) words as the query's single pool for the usual finalize step.
constexpr int64_t kDenseMaxRows = 32768;
constexpr int64_t kDenseMaxBytes = 192ll << 20;

template <int KP>
__global__ void __launch_bounds__(128)
dense_select_kernel(const float* __restrict__ keys, int64_t ld, int64_t n, int64_t nq, uint64_t* __restrict__ pools,
                    int* __restrict__ pool_cnt) {
  constexpr int CAP = pool_cap(KP);
  extern __shared__ __align__(16) uint8_t smem_sel[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= nq) return;
  WarpTopK<KP> sel;
  sel.init(reinterpret_cast<uint64_t*>(smem_sel) + warp * CAP);
  const float* row = keys + q * ld;
  for (int64_t base = 0; base < n; base += 128) {            // four coalesced 128-byte loads in flight
    float x[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t i = base + t * 32 + lane;
      x[t] = i < n ? __ldcs(row + i) : CUDART_INF_F;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t i = base + t * 32 + lane;
      if (base + t * 32 >= n) break;                            // warp-uniform
      sel.push(i < n, x[t], static_cast<uint32_t>(i), lane);
    }
  }
  if (sel.cnt > KP) sel.compact(lane);
  __syncwarp();
  uint64_t* out = pools + q * CAP;
  for (int i = lane; i < sel.cnt; i += 32) out[i] = sel.pool[i];
  if (lane == 0) pool_cnt[q] = sel.cnt;
}

// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Descriptors are pure functions of (pointer, rows, kpad): a search encodes twelve of them (three
// launches x four operands), mostly for the same buffers as the previous call, so the last few are
// kept per thread.
struct MapCacheEntry { const float* ptr; int64_t rows; int kpad; CUtensorMap map; };
static int make_operand_map_uncached(CUtensorMap* map, const float* ptr, int64_t rows, int kpad);
static int make_operand_map(CUtensorMap* map, const float* ptr, int64_t rows, int kpad) {
  constexpr int kSlots = 16;
  thread_local MapCacheEntry cache[kSlots] = {};
  thread_local int next = 0;
  for (int i = 0; i < kSlots; ++i) {
    if (cache[i].ptr == ptr && cache[i].rows == rows && cache[i].kpad == kpad) { *map = cache[i].map; return 0; }
  }
  const int rc = make_operand_map_uncached(map, ptr, rows, kpad);
  if (rc) return rc;
  cache[next] = MapCacheEntry{ptr, rows, kpad, *map};
  next = (next + 1) % kSlots;
  return 0;
}

static int make_operand_map_uncached(CUtensorMap* map, const float* ptr, int64_t rows, int kpad) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VDB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    VDB_REQUIRE(sym != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kpad) * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VDB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld kpad=%d", (int)r, (long long)rows, kpad);
  return 0;
}

template <int CG, bool ARES, int KP, bool DENSE, int SEED = 0, int HAM = 0>
static int launch_tc(const CUtensorMap& mqh, const CUtensorMap& mql, const CUtensorMap& mbh, const CUtensorMap& mbl,
                     const FlatScanParams& P, int clusters, cudaStream_t stream) {
  auto kern = flat_scan_tc_kernel<CG, ARES, KP, DENSE, SEED, HAM>;
  constexpr int smem = tc::smem_bytes<ARES>();
  static int configured[64] = {};     // per device; the attribute call is idempotent, so a race only repeats it
  int dev = 0;
  VDB_CHECK_CUDA(cudaGetDevice(&dev));
  if (!__atomic_load_n(&configured[dev & 63], __ATOMIC_ACQUIRE)) {
    VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    __atomic_store_n(&configured[dev & 63], 1, __ATOMIC_RELEASE);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(tc::threads<HAM>());
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mqh, mql, mbh, mbl, P));
  return 0;
}

template <int CG, bool ARES>
static int query_coresident(int sm) {
  auto kern = flat_scan_tc_kernel<CG, ARES, 128, false>;
  constexpr int smem = tc::smem_bytes<ARES>();
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return sm / CG; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sm / CG * CG);
  cfg.blockDim = dim3(tc::kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); return sm / CG; }
  return n;
}

static int coresident_clusters(int cta_group, bool a_resident, int sm) {
  // cached per (thread, device, variant): the answer is a property of the device / context partition
  thread_local int cache[64][4] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return sm / cta_group; }
  int& slot = cache[dev & 63][(cta_group == 2 ? 2 : 0) + (a_resident ? 1 : 0)];
  if (slot == 0) {
    slot = cta_group == 2 ? (a_resident ? query_coresident<2, true>(sm) : query_coresident<2, false>(sm))
                          : (a_resident ? query_coresident<1, true>(sm) : query_coresident<1, false>(sm));
  }
  return slot;
}

template <int KP>
static int launch_finalize(int metric, const float* hi, const float* lo, int kpad, int64_t n, int64_t id_offset, const float* q_hi,
                           const float* q_lo, int64_t nq, int n_pools, const uint64_t* pools, const int* pool_cnt, int k, int flags,
                           float pad_value, float* out_d, int64_t* out_i, cudaStream_t stream) {
  auto kern = flat_finalize_kernel<KP>;
  const size_t smem = static_cast<size_t>(4) * pool_cap(KP) * 8;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>((nq + 3) / 4), 128, smem, stream>>>(metric, hi, lo, kpad, n, id_offset, q_hi, q_lo, nq, n_pools, pools,
                                                                    pool_cnt, k, flags, pad_value, out_d, out_i);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int KP>
static int run_scan(int impl, const float* hi, const float* lo, int64_t n_pad, int kpad, const float* q_hi,
                    const float* q_lo, int64_t nq_pad, const FlatPlan& plan, const FlatScanParams& P, int sm,
                    cudaStream_t stream) {
  if (plan.cta_group == 0) {
    dim3 grid(plan.n_qtiles, plan.n_chunks);
    flat_scan_simt_kernel<KP><<<grid, 256, 0, stream>>>(hi, lo, q_hi, q_lo, kpad, P);
    VDB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  CUtensorMap mqh, mql, mbh, mbl;
  if (make_operand_map(&mqh, q_hi, nq_pad, kpad) || make_operand_map(&mql, q_lo, nq_pad, kpad) ||
      make_operand_map(&mbh, hi, n_pad, kpad) || make_operand_map(&mbl, lo, n_pad, kpad))
    return 3;
  const bool ares = P.kb <= tc::kMaxResidentKb;
  if constexpr (KP == 32) {   // the seeding pre-pass is instantiated once, under the smallest pool size
    if (P.seed_out != nullptr && P.seed_keep > 16) {       // kept minima per item: 16 or 32 (the insertion network's depth)
      if (plan.cta_group == 2)
        return ares ? launch_tc<2, true, KP, false, 32>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                    : launch_tc<2, false, KP, false, 32>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
      return ares ? launch_tc<1, true, KP, false, 32>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                  : launch_tc<1, false, KP, false, 32>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
    }
    if (P.seed_out != nullptr) {
      if (plan.cta_group == 2)
        return ares ? launch_tc<2, true, KP, false, 16>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                    : launch_tc<2, false, KP, false, 16>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
      return ares ? launch_tc<1, true, KP, false, 16>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                  : launch_tc<1, false, KP, false, 16>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
    }
  }
  if constexpr (KP == 32) {   // the dense-key test hook exists for the smallest pool size only
    if (P.dense != nullptr) {
      if (plan.cta_group == 2)
        return ares ? launch_tc<2, true, KP, true>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                    : launch_tc<2, false, KP, true>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
      return ares ? launch_tc<1, true, KP, true>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                  : launch_tc<1, false, KP, true>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
    }
  }
  if (plan.cta_group == 2) {
    return ares ? launch_tc<2, true, KP, false>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
                : launch_tc<2, false, KP, false>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
  }
  return ares ? launch_tc<1, true, KP, false>(mqh, mql, mbh, mbl, P, plan.clusters, stream)
              : launch_tc<1, false, KP, false>(mqh, mql, mbh, mbl, P, plan.clusters, stream);
}

template <int KP>
static int flat_topk_impl(int metric, const float* hi, const float* lo, const float* norms, int64_t n, int d,
                          int64_t id_offset, const float* q_hi, const float* q_lo, int64_t nq, int k, int flags,
                          float pad_value, int impl, float* out_d, int64_t* out_i, void* ws, size_t ws_bytes,
                          float* dense, cudaStream_t stream) {
  int sm = 0;
  if (vdb_sm_count(&sm)) return 1;
  const int kpad = vdb_flat_kpad(d);
  const int64_t n_pad = vdb_flat_npad(n), nq_pad = vdb_flat_nqpad(nq);
  const FlatPlan plan = make_plan(impl, nq, n_pad, kpad, sm);
  PrePlan pre{};
  if (dense == nullptr && g_debug_mode != 5 && g_debug_mode != 9 && g_debug_mode != 7)
    pre = make_pre_plan(impl, nq, plan, kpad, sm, KP);
  constexpr int CAP = pool_cap(KP);
  const size_t off_cnt = align256(static_cast<size_t>(nq_pad) * 4);
  const size_t off_hand = off_cnt + align256(static_cast<size_t>(nq_pad) * plan.n_pools * 4);
  const size_t off_pool = off_hand + align256(static_cast<size_t>(plan.n_qtiles + 1) * plan.n_pools * 4);
  const size_t off_trash = off_pool + static_cast<size_t>(nq_pad) * plan.n_pools * CAP * 8;
  size_t need = align256(off_trash + static_cast<size_t>(nq_pad) * 8);
  // seeding region: [kept chunk minima nq_pad x n_chunks x kSeedKeep f32][query-tile flags of the redo pass]
  const size_t off_seed = need;
  const size_t off_act = off_seed + align256(static_cast<size_t>(nq_pad) * (pre.on ? pre.plan.n_chunks : 0) * kSeedKeep * 4);
  if (pre.on) need = off_act + align256(static_cast<size_t>(plan.n_qtiles + 1) * 4);
  VDB_REQUIRE(ws != nullptr && ws_bytes >= need, "vdb_flat_topk: workspace too small (%zu < %zu)", ws_bytes, need);
  uint8_t* w = static_cast<uint8_t*>(ws);
  FlatScanParams P{};
  P.norms = norms; P.nq = nq; P.n_tiles = plan.n_tiles; P.tiles_per_chunk = plan.tiles_per_chunk;
  P.n_chunks = plan.n_chunks; P.n_pools = plan.n_pools; P.n_qtiles = plan.n_qtiles; P.kb = kpad / 32;
  P.kslices = (d + 7) / 8;
  P.handover = reinterpret_cast<int*>(w + off_hand);
  P.trash = reinterpret_cast<uint64_t*>(w + off_trash);
  P.thr = reinterpret_cast<uint32_t*>(w);
  P.pool_cnt = reinterpret_cast<int*>(w + off_cnt);
  P.pools = reinterpret_cast<uint64_t*>(w + off_pool);
  P.dense = dense; P.dense_ld = n_pad; P.dbg = g_debug_mode;
  P.tile_stride = 1; P.qtile_active = nullptr; P.seed_out = nullptr;
  P.dense_only = 0;
  // small base: dense keys + a selection kernel instead of pools filled from a loose bound
  const size_t off_dense = align256(need);
  const size_t dense_bytes = static_cast<size_t>(nq_pad) * n_pad * 4;
  if (plan.cta_group != 0 && dense == nullptr && !pre.on && (g_debug_mode == 0 || g_debug_mode == 7) &&
      n_pad <= kDenseMaxRows && static_cast<int64_t>(dense_bytes) <= kDenseMaxBytes && ws_bytes >= off_dense + dense_bytes) {
    FlatScanParams Dn = P;
    Dn.dense = reinterpret_cast<float*>(w + off_dense);
    Dn.dense_only = 1;
    Dn.n_pools = plan.n_chunks;                      // no lineages: nothing is handed over
    const bool timed_d = g_timing_on && g_timing_n < kTimingSlots;
    if (timed_d) VDB_CHECK_CUDA(cudaEventRecord(g_ev0[g_timing_n], stream));
    const int rcd = run_scan<32>(impl, hi, lo, n_pad, kpad, q_hi, q_lo, nq_pad, plan, Dn, sm, stream);
    if (rcd) return rcd;
    if (timed_d) VDB_CHECK_CUDA(cudaEventRecord(g_ev1[g_timing_n++], stream));
    auto sel = dense_select_kernel<KP>;
    const size_t sel_smem = static_cast<size_t>(4) * CAP * 8;
    if (sel_smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(sel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sel_smem)));
    sel<<<static_cast<unsigned>((nq + 3) / 4), 128, sel_smem, stream>>>(Dn.dense, n_pad, n, nq, P.pools, P.pool_cnt);
    VDB_CHECK_CUDA(cudaGetLastError());
    count_launches(2 + (out_d != nullptr ? 1 : 0));
    if (out_d != nullptr) {
      if (launch_finalize<KP>(metric, hi, lo, kpad, n, id_offset, q_hi, q_lo, nq, 1, P.pools, P.pool_cnt, k, flags, pad_value, out_d,
                              out_i, stream)) return 1;
      VDB_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
  }
  int* qtile_active = pre.on ? reinterpret_cast<int*>(w + off_act) : nullptr;
  const int64_t n_cnt = nq_pad * plan.n_pools;
  const int64_t n_hand = static_cast<int64_t>(plan.n_qtiles + 1) * plan.n_pools;
  const int64_t n_act = pre.on ? plan.n_qtiles + 1 : 0;
  flat_init_kernel<<<static_cast<unsigned>((std::max(n_cnt, n_hand) + 255) / 256), 256, 0, stream>>>(
      P.thr, nq_pad, P.pool_cnt, n_cnt, P.handover, n_hand, g_debug_mode == 5 || g_debug_mode == 9, qtile_active, n_act);
  VDB_CHECK_CUDA(cudaGetLastError());
  int launches = 2 + (out_d != nullptr ? 1 : 0);
  if (pre.on) {
    // pre-pass over the strided sample, then tau -> the main pass's starting bounds
    FlatScanParams Q = P;
    Q.n_tiles = pre.plan.n_tiles; Q.tiles_per_chunk = pre.plan.tiles_per_chunk; Q.n_chunks = pre.plan.n_chunks;
    Q.n_pools = pre.plan.n_pools; Q.tile_stride = pre.stride;
    Q.seed_out = reinterpret_cast<float*>(w + off_seed);
    Q.seed_keep = pre.rank > 16 ? 32 : 16;
    Q.dbg = 0;
    const int rc0 = run_scan<32>(impl, hi, lo, n_pad, kpad, q_hi, q_lo, nq_pad, pre.plan, Q, sm, stream);
    if (rc0) return rc0;
    flat_tau_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, 0, stream>>>(
        Q.seed_out, Q.n_chunks, Q.seed_keep, nq, pre.rank, g_debug_mode == 6, P.thr);
    VDB_CHECK_CUDA(cudaGetLastError());
    launches += 2;
  }
  const bool timed = g_timing_on && g_timing_n < kTimingSlots;
  if (timed) VDB_CHECK_CUDA(cudaEventRecord(g_ev0[g_timing_n], stream));
  const int rc = run_scan<KP>(impl, hi, lo, n_pad, kpad, q_hi, q_lo, nq_pad, plan, P, sm, stream);
  if (rc) return rc;
  if (timed) VDB_CHECK_CUDA(cudaEventRecord(g_ev1[g_timing_n++], stream));
  if (pre.on) {
    // verify the guesses; re-scan the query tiles that hold a failed query (normally none: the
    // third launch then finds every tile unflagged and returns)
    flat_verify_kernel<<<static_cast<unsigned>((std::max(nq_pad, n_hand) + 255) / 256), 256, 0, stream>>>(
        P.pool_cnt, plan.n_pools, nq, nq_pad, static_cast<int>(std::min<int64_t>(KP, n)), plan.tile_rows, P.thr,
        qtile_active, P.handover, n_hand);
    VDB_CHECK_CUDA(cudaGetLastError());
    FlatScanParams R = P;
    R.qtile_active = qtile_active;
    R.dbg = 0;
    const int rc2 = run_scan<KP>(impl, hi, lo, n_pad, kpad, q_hi, q_lo, nq_pad, plan, R, sm, stream);
    if (rc2) return rc2;
    launches += 2;
  }
  count_launches(launches);
  if (out_d != nullptr) {
    if (launch_finalize<KP>(metric, hi, lo, kpad, n, id_offset, q_hi, q_lo, nq, plan.n_pools, P.pools, P.pool_cnt, k, flags,
                            pad_value, out_d, out_i, stream)) return 1;
  }
  return 0;
}


// --------------------------------------------------------------------------------------------
// Hamming top-k on the tensor pipe (replaces faiss.IndexLSH.search for candidate generation,
// reference src/algorithms/modular.py:477): codes as bf16 +-1 rows, distances from one kind::f16
// contraction, selection by counting (distances are small integers):
//   sample bound T (popc kernels, lsh.cu) -> count keys <= T per (segment, query, bin) -> cut bin t*,
//   rows taken from it per segment, list offsets -> collect the entries <= t* (segment-major, rows
//   ascending) -> stable counting sort per query = (distance, id) order.  Queries whose sampled bound
//   turns out too small are recounted without a bound by a second launch over their query tiles only.
__global__ void ham_expand_kernel(const uint32_t* __restrict__ codes, int64_t n, int words, int nbits, int kwords,
                                  int negate, int fp16, uint4* __restrict__ out, float* __restrict__ norms, int64_t rows_pad) {
  // thread = (row, 32-bit word): writes 32 bf16 values (64 bytes)
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows_pad * kwords) return;
  const int64_t row = i / kwords;
  const int w = static_cast<int>(i % kwords);
  const uint32_t bits = row < n && w < words ? codes[row * words + w] : 0u;
  const uint32_t p1 = fp16 ? 0x3C00u : 0x3F80u, m1 = fp16 ? 0xBC00u : 0xBF80u;            // +1.0 / -1.0 as fp16 or bf16
  const uint32_t one = negate ? m1 : p1, minus = negate ? p1 : m1;
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    uint32_t v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = w * 32 + 2 * j + h;
      v[h] = row < n && b < nbits ? (((bits >> (2 * j + h)) & 1u) ? one : minus) : 0u;
    }
    pk[j] = v[0] | (v[1] << 16);
  }
  uint4* dst = out + i * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  if (norms != nullptr && w == 0) norms[row] = row < n ? 0.f : CUDART_INF_F;
}

// Selection over what the scan collected for a query (lists per segment, rows ascending, every key
// within the sampled bound; entry = distance << 23 | row): histogram of the distances in shared memory, cut bin and tie budget, then
// a stable placement pass - segments and rows arrive ascending, so equal distances keep id order.
// A query whose lists overflowed or hold fewer than `need` entries is flagged for the exact popc path.
__global__ void __launch_bounds__(128)
ham_select_kernel(const uint32_t* __restrict__ list, const int* __restrict__ lcnt, int segs, int cap, int64_t nq, int nbits,
                  int k, int need, uint32_t n_rows, int64_t id_offset, float* __restrict__ out_d, int64_t* __restrict__ out_i,
                  uint8_t* __restrict__ fallback) {
  extern __shared__ int cur_all[];                     // [4][nbits + 2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= nq) return;
  const int bins = nbits + 1;
  int* cur = cur_all + warp * (bins + 1);
  for (int b = lane; b <= bins; b += 32) cur[b] = 0;
  __syncwarp();
  bool over = false;
  for (int s = 0; s < segs; ++s) over |= lcnt[static_cast<int64_t>(s) * nq + q] > cap;
  if (over) {
    if (lane == 0) fallback[q] = 1;
    return;
  }
  // histogram of the collected distances; padding rows (row >= n_rows: the fp16 scan does not mask them) are skipped
  int mine = 0;
  for (int s = 0; s < segs; ++s) {
    const int c = lcnt[static_cast<int64_t>(s) * nq + q];
    const uint32_t* src = list + (static_cast<int64_t>(s) * nq + q) * cap;
    for (int i = lane; i < c; i += 32) {
      const uint32_t w = src[i];
      if ((w & 0x7fffffu) < n_rows) { atomicAdd(cur + static_cast<int>(w >> 23), 1); ++mine; }
    }
  }
  const int total = __reduce_add_sync(0xffffffffu, mine);
  if (total < need) {                                 // the sampled bound was too tight: exact popc path for this query
    if (lane == 0) fallback[q] = 1;
    return;
  }
  if (lane == 0) fallback[q] = 0;
  __syncwarp();
  int t = bins, tie_budget = 0;
  if (lane == 0) {                                     // bins -> output offsets; cut bin and its budget
    int run = 0;
    t = nbits; tie_budget = 1 << 30;
    bool found = false;
    for (int b = 0; b < bins; ++b) {
      const int c = cur[b];
      if (!found && run + c >= k) { t = b; tie_budget = k - run; found = true; }
      cur[b] = run;
      run += c;
    }
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  tie_budget = __shfl_sync(0xffffffffu, tie_budget, 0);
  __syncwarp();
  const unsigned lt_mask = (1u << lane) - 1u;
  int seen = 0;                                        // ties of the cut bin met so far (warp-uniform)
  for (int s = 0; s < segs; ++s) {
    const int c = lcnt[static_cast<int64_t>(s) * nq + q];
    const uint32_t* src = list + (static_cast<int64_t>(s) * nq + q) * cap;
    for (int base = 0; base < c; base += 32) {
      const int i = base + lane;
      const uint32_t w = i < c ? src[i] : 0xffffffffu;
      const bool valid = i < c && (w & 0x7fffffu) < n_rows;
      const int dist = valid ? static_cast<int>(w >> 23) : -1;
      const bool tie = valid && dist == t;
      const unsigned tie_m = __ballot_sync(0xffffffffu, tie);
      const bool takes = valid && (dist < t || (tie && seen + __popc(tie_m & lt_mask) < tie_budget));
      seen += __popc(tie_m);
      const unsigned peers = __match_any_sync(0xffffffffu, takes ? dist : -2 - lane);
      if (takes) {
        const int slot = cur[dist] + __popc(peers & lt_mask);
        if (slot < k) {
          out_d[q * k + slot] = static_cast<float>(dist);
          out_i[q * k + slot] = static_cast<int64_t>(w & 0x7fffffu) + id_offset;
        }
      }
      __syncwarp();
      if (takes && (peers & lt_mask) == 0u) cur[dist] += __popc(peers);
      __syncwarp();
    }
  }
}

}  // namespace vdb

using namespace vdb;

extern "C" {

int vdb_set_debug_mode(int mode) {
  const int old = g_debug_mode;
  g_debug_mode = mode;
  return old;
}

int vdb_flat_grid_clusters(int impl, int d, int* out) {
  int sm = 0;
  if (vdb_sm_count(&sm)) return 1;
  if (impl == VDB_IMPL_AUTO) impl = VDB_IMPL_TCGEN05;
  VDB_REQUIRE((impl == VDB_IMPL_TCGEN05 || impl == VDB_IMPL_TCGEN05_1CTA) && d > 0 && out != nullptr, "vdb_flat_grid_clusters: bad arguments");
  const int cg = impl == VDB_IMPL_TCGEN05_1CTA ? 1 : 2;
  *out = std::max(1, std::min(sm / cg, coresident_clusters(cg, vdb_flat_kpad(d) / 32 <= tc::kMaxResidentKb, sm)));
  return 0;
}

int vdb_debug_read_prof(uint64_t* out8) {
  unsigned long long h[8] = {};
  VDB_CHECK_CUDA(cudaMemcpyFromSymbol(h, g_prof, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  unsigned long long z[8] = {};
  VDB_CHECK_CUDA(cudaMemcpyToSymbol(g_prof, z, sizeof(z)));
  return 0;
}

int vdb_flat_set_seeding(int sample_tiles, int rank) {
  VDB_REQUIRE(sample_tiles >= 0 && rank >= 0 && rank <= kSeedKeep, "vdb_flat_set_seeding: sample_tiles >= 0, 0 (per shard) <= rank <= %d", kSeedKeep);
  g_pre_tiles = sample_tiles;
  g_pre_rank = rank;
  return 0;
}

int vdb_flat_set_seeding_margin(int margin) {
  VDB_REQUIRE(margin >= 0 && margin <= 64, "vdb_flat_set_seeding_margin: 0 (default) .. 64");
  g_pre_margin = margin;
  return 0;
}

int vdb_debug_redo_queries(uint64_t* out) {
  unsigned long long h = 0, z = 0;
  VDB_CHECK_CUDA(cudaMemcpyFromSymbol(&h, g_redo_queries, sizeof(h)));
  VDB_CHECK_CUDA(cudaMemcpyToSymbol(g_redo_queries, &z, sizeof(z)));
  *out = h;
  return 0;
}

int vdb_flat_timing_enable(int on) {
  if (on && !g_ev_made) {
    for (int i = 0; i < kTimingSlots; ++i) {
      VDB_CHECK_CUDA(cudaEventCreate(&g_ev0[i]));
      VDB_CHECK_CUDA(cudaEventCreate(&g_ev1[i]));
    }
    g_ev_made = true;
  }
  g_timing_on = on != 0;
  g_timing_n = 0;
  return 0;
}

int vdb_flat_timing_read(float* ms, int max_records, int* n_out) {
  const int n = g_timing_n < max_records ? g_timing_n : max_records;
  for (int i = 0; i < n; ++i) {
    VDB_CHECK_CUDA(cudaEventSynchronize(g_ev1[i]));
    VDB_CHECK_CUDA(cudaEventElapsedTime(ms + i, g_ev0[i], g_ev1[i]));
  }
  if (n_out != nullptr) *n_out = n;
  g_timing_n = 0;
  return 0;
}

size_t vdb_flat_topk_workspace_bytes(int64_t nq, int k) {
  int sm = 148;
  vdb_sm_count(&sm);
  const int kp = keep_for_k(k);
  if (kp == 0 || nq <= 0) return 0;
  const int64_t nq_pad = vdb_flat_nqpad(nq);
  const int s = max_pools(nq, sm);
  const size_t main_part = align256(static_cast<size_t>(nq_pad) * 4) + align256(static_cast<size_t>(nq_pad) * s * 4) +
                           align256(static_cast<size_t>(nq_pad / 64 + 2) * s * 4) +
                           static_cast<size_t>(nq_pad) * s * pool_cap(kp) * 8 + align256(static_cast<size_t>(nq_pad) * 8);
  // seeding pre-pass: kSeedKeep chunk minima per (query, item of the sample), plus the redo pass's tile flags
  const int64_t qt_min = std::max<int64_t>(1, nq_pad / 256);
  const int64_t chunks_max = std::min<int64_t>(sm, (static_cast<int64_t>(sm) * 8 + qt_min - 1) / qt_min + 2);
  const size_t pre_part = align256(static_cast<size_t>(nq_pad) * chunks_max * kSeedKeep * 4) +
                          align256(static_cast<size_t>(nq_pad / 64 + 2) * 4);
  // small-base path: the dense key matrix (used only while it fits these caps)
  const size_t dense_part = static_cast<size_t>(std::min<int64_t>(kDenseMaxBytes, nq_pad * kDenseMaxRows * 4)) + 256;
  return main_part + pre_part + dense_part + 256;
}

int vdb_flat_topk(int metric, const float* hi, const float* lo, const float* norms, int64_t n, int d,
                  int64_t id_offset, const float* q_hi, const float* q_lo, int64_t nq, int k, int flags,
                  float pad_value, int impl, float* out_d, int64_t* out_i, void* workspace, size_t workspace_bytes,
                  void* stream) {
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_flat_topk: bad metric %d", metric);
  VDB_REQUIRE(n > 0 && nq > 0 && d > 0, "vdb_flat_topk: empty problem n=%lld nq=%lld d=%d", (long long)n, (long long)nq, d);
  VDB_REQUIRE(n < (int64_t(1) << 32), "vdb_flat_topk: shard too large (%lld rows; shard the base)", (long long)n);
  const int kp = keep_for_k(k);
  VDB_REQUIRE(k >= 1 && kp != 0, "vdb_flat_topk: k=%d unsupported (1..504)", k);
  if (impl == VDB_IMPL_AUTO) impl = VDB_IMPL_TCGEN05;
  VDB_REQUIRE(impl >= VDB_IMPL_TCGEN05 && impl <= VDB_IMPL_SIMT, "vdb_flat_topk: bad impl %d", impl);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define VDB_GO(KP)                                                                                              \
  return flat_topk_impl<KP>(metric, hi, lo, norms, n, d, id_offset, q_hi, q_lo, nq, k, flags, pad_value, impl,  \
                            out_d, out_i, workspace, workspace_bytes, nullptr, s)
  switch (kp) {
    case 32: VDB_GO(32);
    case 128: VDB_GO(128);
    case 256: VDB_GO(256);
    default: VDB_GO(512);
  }
#undef VDB_GO
}

int vdb_flat_dense_keys(const float* hi, const float* lo, const float* norms, int64_t n, int d, const float* q_hi,
                        const float* q_lo, int64_t nq, int impl, float* keys, void* stream) {
  VDB_REQUIRE(n > 0 && nq > 0 && d > 0, "vdb_flat_dense_keys: empty problem");
  if (impl == VDB_IMPL_AUTO) impl = VDB_IMPL_TCGEN05;
  const size_t bytes = vdb_flat_topk_workspace_bytes(nq, 24);
  void* ws = nullptr;
  VDB_CHECK_CUDA(cudaMalloc(&ws, bytes));   // test hook only: the product path never allocates
  const int rc = flat_topk_impl<32>(VDB_METRIC_L2, hi, lo, norms, n, d, 0, q_hi, q_lo, nq, 24, 0, 0.f, impl, nullptr,
                                    nullptr, ws, bytes, keys, static_cast<cudaStream_t>(stream));
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  cudaFree(ws);
  if (rc) return rc;
  VDB_CHECK_CUDA(e);
  return 0;
}

int vdb_merge_topk_strided(const float* d_all, const int64_t* i_all, int64_t stride_d, int64_t stride_i, int parts,
                           int64_t nq, int k, int descending, float pad_value, float* out_d, int64_t* out_i, void* stream) {
  VDB_REQUIRE(parts >= 1 && nq > 0 && k >= 1, "vdb_merge_topk: bad shape");
  VDB_REQUIRE(stride_d >= nq * k && stride_i >= nq * k, "vdb_merge_topk: part stride smaller than nq * k");
  const int kp = k <= 32 ? 32 : k <= 128 ? 128 : k <= 256 ? 256 : k <= 512 ? 512 : 0;
  VDB_REQUIRE(kp != 0, "vdb_merge_topk: k=%d unsupported (<= 512)", k);
  VDB_REQUIRE(static_cast<int64_t>(parts) * k < (int64_t(1) << 31), "vdb_merge_topk: parts*k too large");
  const unsigned blocks = static_cast<unsigned>((nq + 3) / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define VDB_GO(KP) merge_topk_kernel<KP><<<blocks, 128, 0, s>>>(d_all, i_all, stride_d, stride_i, parts, nq, k, descending, pad_value, out_d, out_i)
  switch (kp) {
    case 32: VDB_GO(32); break;
    case 128: VDB_GO(128); break;
    case 256: VDB_GO(256); break;
    default: VDB_GO(512); break;
  }
#undef VDB_GO
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_merge_topk(const float* d_all, const int64_t* i_all, int parts, int64_t nq, int k, int descending,
                   float pad_value, float* out_d, int64_t* out_i, void* stream) {
  return vdb_merge_topk_strided(d_all, i_all, nq * k, nq * k, parts, nq, k, descending, pad_value, out_d, out_i, stream);
}

int vdb_hamming_tc_row_bytes(int nbits) { return nbits <= 0 || nbits > 256 ? 0 : (nbits + 63) / 64 * 128; }

int vdb_hamming_tc_expand(const uint32_t* codes, int64_t n, int words, int nbits, int negate, void* out, float* norms,
                          int64_t rows_pad, void* stream) {
  const int rb = vdb_hamming_tc_row_bytes(nbits);
  VDB_REQUIRE(rb != 0 && n > 0 && rows_pad >= n && words >= (nbits + 31) / 32, "vdb_hamming_tc_expand: bad shape (nbits <= 256)");
  const int kwords = rb / 64;
  const int64_t threads = rows_pad * kwords;
  ham_expand_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      codes, n, words, nbits, kwords, negate & 1, (negate >> 1) & 1, static_cast<uint4*>(out), norms, rows_pad);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int ham_list_cap(int64_t n, int k, int segs) {
  // expected entries per (segment, query): the sampled bound aims at lam + 4 sqrt(lam) + 4 of 32 768 sample
  // rows, and rounds up to a whole distance bin; four times the per-segment average, plus slack
  const double s = 32768.0, lam = static_cast<double>(k) * s / static_cast<double>(n);
  const double target = (lam + 4.0 * sqrt(lam) + 4.0) * static_cast<double>(n) / s;
  return static_cast<int>(4.0 * target / segs) + 128;
}

size_t vdb_hamming_tc_workspace_bytes(int64_t nq, int nbits, int k, int64_t n) {
  const int rb = vdb_hamming_tc_row_bytes(nbits);
  if (nq <= 0 || rb == 0 || k < 1 || n <= 0) return 0;
  int sm = 148;
  vdb_sm_count(&sm);
  const FlatPlan plan = make_plan(VDB_IMPL_TCGEN05, nq, vdb_flat_npad(n), rb / 4, sm);
  const int64_t segs = 2 * plan.n_chunks;                     // every chunk is drained as two segments (flat_tc.cuh)
  const int cap = ham_list_cap(n, k, static_cast<int>(segs));
  return align256(static_cast<size_t>(segs) * nq * cap * 4) + align256(static_cast<size_t>(segs) * nq * 4) +
         2 * align256(static_cast<size_t>(nq) * 4) + align256(static_cast<size_t>(nq) * (nbits + 1) * 4) +
         align256(vdb_hamming_topk_workspace_bytes(nq, nbits)) + 256;
}

static int hamming_topk_tc_impl(bool fp16, const void* base_bf16, const float* norms, const uint32_t* codes, int64_t n, const void* q_bf16,
                                const uint32_t* qcodes, int64_t nq, int nbits, int k, int64_t id_offset, float* out_d,
                                int64_t* out_i, void* workspace, size_t workspace_bytes, void* stream) {
  const int rb = vdb_hamming_tc_row_bytes(nbits);
  VDB_REQUIRE(rb != 0 && n > 65536 && nq > 0 && k >= 1 && vdb_flat_npad(n) <= (int64_t(1) << 23),
              "vdb_hamming_topk_tc: bad shape (nbits <= 256, 65536 < n <= 8388608: list entries hold a 23-bit row)");
  VDB_REQUIRE(workspace != nullptr && workspace_bytes >= vdb_hamming_tc_workspace_bytes(nq, nbits, k, n),
              "vdb_hamming_topk_tc: workspace too small");
  int sm = 0;
  if (vdb_sm_count(&sm)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kpad = rb / 4;                                     // the code rows seen as fp32 rows: same bytes, same tiles
  const int64_t n_pad = vdb_flat_npad(n), nq_pad = vdb_flat_nqpad(nq);
  const FlatPlan plan = make_plan(VDB_IMPL_TCGEN05, nq, n_pad, kpad, sm);
  const int segs = 2 * plan.n_chunks, bins = nbits + 1;
  const int cap = ham_list_cap(n, k, segs);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  size_t off = 0;
  uint32_t* list = reinterpret_cast<uint32_t*>(w + off); off += align256(static_cast<size_t>(segs) * nq * cap * 4);
  int* lcnt = reinterpret_cast<int*>(w + off);           off += align256(static_cast<size_t>(segs) * nq * 4);
  int* T = reinterpret_cast<int*>(w + off);              off += align256(static_cast<size_t>(nq) * 4);
  uint8_t* fallback = w + off;                           off += align256(static_cast<size_t>(nq) * 4);
  int* sample_hist = reinterpret_cast<int*>(w + off);    off += align256(static_cast<size_t>(nq) * bins * 4);
  void* popc_ws = w + off;

  if (hamming_sample_bound(codes, n, qcodes, nq, nbits, k, sample_hist, T, s)) return 3;
  FlatScanParams P{};
  P.norms = norms; P.nq = nq; P.n_tiles = plan.n_tiles; P.tiles_per_chunk = plan.tiles_per_chunk;
  P.n_chunks = plan.n_chunks; P.n_pools = plan.n_chunks; P.n_qtiles = plan.n_qtiles; P.kb = kpad / 32;
  P.kslices = 4 * P.kb;
  P.tile_stride = 1; P.ham_nbits = nbits; P.ham_bound = T; P.ham_list = list; P.ham_cnt = lcnt; P.ham_cap = cap;
  CUtensorMap mq, mb;
  if (make_operand_map(&mq, static_cast<const float*>(q_bf16), nq_pad, kpad) ||
      make_operand_map(&mb, static_cast<const float*>(base_bf16), n_pad, kpad))
    return 3;
  if (fp16) {
    VDB_REQUIRE(nbits % 2 == 0, "vdb_hamming_topk_tc_f16: nbits must be even (the fp16 epilogue halves key + nbits exactly)");
    if (launch_tc<2, true, 32, false, 0, 2>(mq, mq, mb, mb, P, plan.clusters, s)) return 3;
  } else {
    if (launch_tc<2, true, 32, false, 0, 1>(mq, mq, mb, mb, P, plan.clusters, s)) return 3;
  }
  ham_select_kernel<<<static_cast<unsigned>((nq + 3) / 4), 128, static_cast<size_t>(4) * (bins + 1) * sizeof(int), s>>>(
      list, lcnt, segs, cap, nq, nbits, k, static_cast<int>(std::min<int64_t>(k, n)), static_cast<uint32_t>(n), id_offset, out_d,
      out_i, fallback);
  VDB_CHECK_CUDA(cudaGetLastError());
  count_launches(2);
  // queries whose bound was short or whose lists overflowed: the exact popc path, for them alone
  return hamming_topk_subset(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, popc_ws, fallback, s);
}

int vdb_hamming_topk_tc(const void* base_bf16, const float* norms, const uint32_t* codes, int64_t n, const void* q_bf16,
                        const uint32_t* qcodes, int64_t nq, int nbits, int k, int64_t id_offset, float* out_d,
                        int64_t* out_i, void* workspace, size_t workspace_bytes, void* stream) {
  return hamming_topk_tc_impl(false, base_bf16, norms, codes, n, q_bf16, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace,
                              workspace_bytes, stream);
}

int vdb_hamming_topk_tc_f16(const void* base_f16, const float* norms, const uint32_t* codes, int64_t n, const void* q_f16,
                            const uint32_t* qcodes, int64_t nq, int nbits, int k, int64_t id_offset, float* out_d,
                            int64_t* out_i, void* workspace, size_t workspace_bytes, void* stream) {
  return hamming_topk_tc_impl(true, base_f16, norms, codes, n, q_f16, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace,
                              workspace_bytes, stream);
}

}  // extern "C"
