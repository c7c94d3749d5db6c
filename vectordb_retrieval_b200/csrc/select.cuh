// Warp-private running top-k' in shared memory, shared by the IVF list scan and the LSH rerank.
#pragma once
#include "common.cuh"

namespace vdb {

template <int KP>
struct WarpTopK {
  static constexpr int CAP = pool_cap(KP);
  static constexpr int E = CAP / 32;
  uint64_t* pool;   // CAP slots, warp-private shared memory
  int cnt;          // warp-uniform
  float thr;        // warp-uniform: KP-th smallest key seen so far (+inf until KP keys were seen)

  __device__ __forceinline__ void init(uint64_t* p) { pool = p; cnt = 0; thr = CUDART_INF_F; }

  // one out-of-line copy per kernel; state travels by value so the hot loop keeps it in registers.
  // Selection, not a sort (the pool is unordered anyway): radix-select the KP-th key, keep the KP
  // smallest (key, row) words in place - ~400 instructions where a 256-element bitonic sort took 2 600
  // (18 % of the rerank kernel's executed instructions, profiles/rerank_kernel_r1.md).
  static __device__ __noinline__ PoolState compact_impl(uint64_t* pool, int cnt, float thr, int lane) {
    __syncwarp();
    if (cnt <= KP) return PoolState{thr, cnt};
    uint32_t hi[E], lo[E];
    bool live[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      live[e] = i < cnt;
      const uint64_t w = live[e] ? pool[i] : kEmpty;
      hi[e] = static_cast<uint32_t>(w >> 32);
      lo[e] = static_cast<uint32_t>(w);
    }
    const uint32_t kth = warp_select_kth<E>(hi, live, KP);
    int below = 0, equal = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      below += (live[e] && hi[e] < kth) ? 1 : 0;
      equal += (live[e] && hi[e] == kth) ? 1 : 0;
    }
    below = __reduce_add_sync(0xffffffffu, below);
    equal = __reduce_add_sync(0xffffffffu, equal);
    uint32_t row_cut = 0xffffffffu;                         // keys equal to the bound: keep the lowest rows
    if (below + equal > KP) {
      bool tie[E];
#pragma unroll
      for (int e = 0; e < E; ++e) tie[e] = live[e] && hi[e] == kth;
      row_cut = warp_select_kth<E>(lo, tie, KP - below);
    }
    __syncwarp();                                           // every slot is in registers before the in-place rewrite
    const unsigned lt_mask = (1u << lane) - 1u;
    int off = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool keep = live[e] && (hi[e] < kth || (hi[e] == kth && lo[e] <= row_cut));
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) pool[off + __popc(m & lt_mask)] = (static_cast<uint64_t>(hi[e]) << 32) | lo[e];
      off += __popc(m);
    }
    thr = fminf(thr, ord2f(kth));
    __syncwarp();
    return PoolState{thr, off};
  }
  __device__ __forceinline__ void compact(int lane) {
    const PoolState st = compact_impl(pool, cnt, thr, lane);
    thr = st.thr;
    cnt = st.cnt;
  }

  // the same for words that are already packed (ordered key << 32 | row)
  __device__ __forceinline__ void push_packed(bool valid, uint64_t w, int lane) {
    if (cnt > CAP - 32) compact(lane);
    const bool pass = valid && static_cast<uint32_t>(w >> 32) <= f2ord(thr);
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (pass) pool[cnt + __popc(m & ((1u << lane) - 1u))] = w;
    cnt += __popc(m);
  }

  // every lane of the warp calls; `valid` lanes offer one (key,row) each
  __device__ __forceinline__ void push(bool valid, float key, uint32_t row, int lane) {
    if (cnt > CAP - 32) compact(lane);
    // `<=`: a key equal to the bound may still belong to the best KP by (key, row) order - the compaction
    // keeps the lowest rows of a tie group - so the result never depends on the order rows are met in
    const bool pass = valid && key <= thr;
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (pass) pool[cnt + __popc(m & ((1u << lane) - 1u))] = pack_key(key, row);
    cnt += __popc(m);
  }
};

// Barrier over the NT threads that serve one query (a CTA holds 256 / NT such groups; id 1..8, 0 is __syncthreads).
template <int NT>
__device__ __forceinline__ void group_sync(int id) {
  if constexpr (NT == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NT) : "memory");
}

// Merge the W warp pools of one query (the group's warp 0 does it) and write the best k as (distance, id).
//   key -> distance: L2: key (or sqrt), IP: key = -score.
// W warps serve one query; a CTA may hold several such groups (small probe counts / candidate sets: one
// query per 8 warps leaves every warp a handful of rows, no pool ever fills, and the group's single merging
// warp walks W nearly empty pools).  pools_smem / cnts_smem / thr_s are the GROUP's arrays, `warp` the warp's
// index inside the group, bar_id the group's barrier.
template <int KP, int W>
__device__ __forceinline__ void cta_write_topk(WarpTopK<KP>& mine, uint64_t* pools_smem, int* cnts_smem, float* thr_s, int warp,
                                               int lane, int bar_id, int metric, int k, int flags, float pad_value,
                                               int64_t id_offset, float* out_d, int64_t* out_i) {
  constexpr int CAP = pool_cap(KP);
  constexpr int E = KP / 32;
  // The tightest of the warps' bounds is a bound for the whole query (that warp alone holds KP keys
  // below it), so every warp first drops what lies above it from its own pool - in parallel - and
  // the serial merge below sees little more than KP entries instead of up to W * CAP.
  if constexpr (W > 1) {
    if (lane == 0) thr_s[warp] = mine.thr;
    group_sync<W * 32>(bar_id);
    float gthr = thr_s[0];
#pragma unroll
    for (int w = 1; w < W; ++w) gthr = fminf(gthr, thr_s[w]);
    const uint32_t cut = f2ord(gthr);
    int kept = 0;
    for (int base = 0; base < mine.cnt; base += 32) {
      const int i = base + lane;
      const uint64_t v = i < mine.cnt ? mine.pool[i] : kEmpty;
      const bool keep = i < mine.cnt && static_cast<uint32_t>(v >> 32) <= cut;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) mine.pool[kept + __popc(m & ((1u << lane) - 1u))] = v;   // kept <= base: never overtakes the reads
      kept += __popc(m);
      __syncwarp();
    }
    mine.cnt = kept;
  }
  if (lane == 0) cnts_smem[warp] = mine.cnt;
  group_sync<W * 32>(bar_id);
  if (warp != 0) return;
  // the W pools are walked as one virtual list, KP entries per merge step: short lists (small
  // nprobe, few candidates) then cost one or two merges instead of one per warp
  uint64_t best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = kEmpty;
  int start[W + 1];
  start[0] = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) start[w + 1] = start[w] + cnts_smem[w];
  const int total = start[W];
  for (int base = 0; base < total; base += KP) {
    uint64_t fresh[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int g = base + e * 32 + lane;
      uint64_t v = kEmpty;
      if (g < total) {
        int w = 0;
#pragma unroll
        for (int t = 1; t < W; ++t) w += g >= start[t] ? 1 : 0;
        v = pools_smem[w * CAP + (g - start[w])];
      }
      fresh[e] = v;
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
        const float key = packed_key(best[e]);
        iv = static_cast<int64_t>(packed_row(best[e])) + id_offset;
        if (metric == VDB_METRIC_L2) dv = (flags & VDB_OUT_SQRT) ? sqrtf(key) : key;
        else dv = (flags & VDB_OUT_NEGATE) ? key : ((flags & VDB_OUT_ONE_MINUS) ? 1.f + key : -key);
      }
      out_d[r] = dv;
      out_i[r] = iv;
    }
  }
}

}  // namespace vdb
