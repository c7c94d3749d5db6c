// Shared device helpers: packed (key,row) candidates, warp-wide bitonic sort, candidate pools.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math_constants.h>

#include "../../include/vdb_cuda.h"

namespace vdb {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
void count_launches(int n);   // every kernel launched by the library is counted (vdb_launch_count)
#define VDB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::vdb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                       __LINE__);                                                         \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)
#define VDB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vdb::set_error(__VA_ARGS__);    \
      return 2;                         \
    }                                   \
  } while (0)

// lsh.cu: bound T[q] such that, judging by a 32k-row sample, about 1.25 k rows of the n lie within
// Hamming distance T[q] (popc kernels; hist_scratch [nq][nbits + 1] ints)
int hamming_sample_bound(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                         int* hist_scratch, int* T, cudaStream_t stream);

// lsh.cu: the exact popc path (count every distance, cut, emit) for the queries flagged in `only`;
// ws as for vdb_hamming_topk
int hamming_topk_subset(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                        int64_t id_offset, float* out_d, int64_t* out_i, void* ws, const uint8_t* only, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Candidate encoding: one 64-bit word, high half = order-preserving image of the fp32 key,
// low half = row index inside the shard.  Unsigned compare == (key, row) lexicographic compare,
// which makes every selection deterministic and independent of how the base is sharded.
constexpr uint64_t kEmpty = ~0ull;
// "No bound yet": the largest finite float, NOT +inf.  Keys are admitted with `key <= bound` so that a row whose key
// EQUALS the running k'-th key still competes on its row id (the compaction keeps the lowest rows of a tie group;
// with `<` a late low-id duplicate was dropped and the result depended on the scan order, i.e. on the sharding).
// Padding rows carry +inf norms, hence +inf keys, and must never pass: +inf <= FLT_MAX is false.
constexpr float kOpenBound = 3.402823466e+38f;

__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return u ^ (static_cast<uint32_t>(static_cast<int32_t>(u) >> 31) | 0x80000000u);   // negative: ~u, else u | sign
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
__device__ __forceinline__ uint64_t pack_key(float key, uint32_t row) {
  return (static_cast<uint64_t>(f2ord(key)) << 32) | row;
}
__device__ __forceinline__ float packed_key(uint64_t p) { return ord2f(static_cast<uint32_t>(p >> 32)); }
__device__ __forceinline__ uint32_t packed_row(uint64_t p) { return static_cast<uint32_t>(p); }

// k' (kept candidates) for a requested k: at least 8 spare slots so that a 1e-6-level error
// of the 3xTF32 key cannot push a true top-k row out before the exact re-scoring.
__host__ __device__ constexpr int keep_for_k(int k) {
  return k + 8 <= 32 ? 32 : k + 8 <= 128 ? 128 : k + 8 <= 256 ? 256 : k + 8 <= 512 ? 512 : 0;
}

// Slots of a candidate pool that keeps the best kp: the owner may append until fewer than 32 slots
// are free, then the pool is cut back to kp.  2*kp leaves kp - 32 appends between two cuts; for
// kp = 32 that would be none (a cut after every append), so small pools get 4*kp.
__host__ __device__ constexpr int pool_cap(int kp) { return kp < 64 ? 4 * kp : 2 * kp; }

// ---------------------------------------------------------------------------------------------
// Bitonic sort of 32*E packed words held by one warp, element i = e*32 + lane, ascending.
template <int E>
__device__ __forceinline__ void warp_sort(uint64_t (&v)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int je = j >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int pe = e ^ je;
          if (pe > e) {
            const bool asc = ((e * 32) & k) == 0;
            const uint64_t a = v[e], b = v[pe];
            const bool sw = (a > b) == asc;
            v[e] = sw ? b : a;
            v[pe] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const bool asc = (((e * 32) | lane) & k) == 0;
          const uint64_t a = v[e];
          const uint64_t b = __shfl_xor_sync(0xffffffffu, a, j);
          const bool keep_min = ((lane & j) == 0) == asc;
          const uint64_t mn = a < b ? a : b, mx = a < b ? b : a;
          v[e] = keep_min ? mn : mx;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Per-query candidate pool in global memory (L2 resident): CAP = pool_cap(KP) slots.  The owner
// thread appends every key that beats its running bound `thr`; when fewer than 32 free slots
// remain the whole warp selects that pool's best KP (radix select, no sort) and tightens `thr`
// to the KP-th key - a valid bound because at least KP scanned rows are <= it.  All 32 lanes call.
struct PoolState {
  float thr;
  int cnt;
};

// Warp-wide selection: the value of rank `kth` (1-indexed, ascending) among the 32*E words
// held as w[e] by the lanes, counting only words whose `live` flag is set.  MSB-first binary
// search on the value bits with one warp-sum per bit; the bits all live words share are skipped.
template <int E>
__device__ __forceinline__ uint32_t warp_select_kth(const uint32_t (&w)[E], const bool (&live)[E], int kth) {
  uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (live[e]) { mn = min(mn, w[e]); mx = max(mx, w[e]); }
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  const uint32_t diff = mn ^ mx;
  if (diff == 0u) return mn;
  const int top = 31 - __clz(diff);                         // highest bit where live words differ
  uint32_t prefix = top == 31 ? 0u : (mn >> (top + 1)) << (top + 1);
  for (int bit = top; bit >= 0; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (live[e] && w[e] < cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < kth) prefix = cand;                             // fewer than kth words below cand: the answer is >= cand
  }
  return prefix;
}

// slow path, one copy per kernel: compacts the pool of every lane flagged in `need`.
// Keeps exactly the KP smallest (key, row) words (selection, not a sort: the pool is unordered)
// and returns the KP-th key as the new bound.
template <int KP>
__device__ __noinline__ PoolState pool_compact(unsigned need, float thr, int cnt, uint64_t* pool, int lane,
                                               uint32_t* thr_shared) {
  constexpr int CAP = pool_cap(KP);
  constexpr int E = CAP / 32;
  const unsigned lt_mask = (1u << lane) - 1u;
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    uint64_t* p = reinterpret_cast<uint64_t*>(
        __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(pool), src));
    const int c = __shfl_sync(0xffffffffu, cnt, src);       // CAP - 32 < c <= CAP, so c >= KP
    __syncwarp();
    uint32_t hi[E], lo[E];
    bool live[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      live[e] = i < c;
      const uint64_t w = live[e] ? __ldcg(p + i) : kEmpty;
      hi[e] = static_cast<uint32_t>(w >> 32);
      lo[e] = static_cast<uint32_t>(w);
    }
    const uint32_t kth = warp_select_kth<E>(hi, live, KP);   // ordered image of the KP-th smallest key
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
    int off = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool keep = live[e] && (hi[e] < kth || (hi[e] == kth && lo[e] <= row_cut));
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcg(p + off + __popc(m & lt_mask), (static_cast<uint64_t>(hi[e]) << 32) | lo[e]);
      off += __popc(m);
    }
    if (lane == src) {
      cnt = KP;
      thr = fminf(thr, ord2f(kth));
      if (thr_shared) atomicMin(thr_shared, kth);
    }
    __syncwarp();
  }
  return PoolState{thr, cnt};
}

template <int KP>
__device__ __forceinline__ void pool_maintain(float& thr, int& cnt, uint64_t* pool, int lane,
                                              uint32_t* thr_shared /* this lane's global bound, may be null */) {
  const unsigned need = __ballot_sync(0xffffffffu, cnt > pool_cap(KP) - 32);
  if (need) {
    const PoolState st = pool_compact<KP>(need, thr, cnt, pool, lane, thr_shared);
    thr = st.thr;
    cnt = st.cnt;
  }
}

// merge step shared by the finalisers: v[0..E) (sorted best so far) with E fresh words -> best E
template <int E>
__device__ __noinline__ void warp_merge_keep(uint64_t* best /* [E] in local */, const uint64_t* fresh, int lane) {
  uint64_t v[2 * E];
#pragma unroll
  for (int e = 0; e < E; ++e) { v[e] = best[e]; v[E + e] = fresh[e]; }
  warp_sort<2 * E>(v, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = v[e];
}

template <int E>
__device__ __noinline__ void warp_sort_noinline(uint64_t* v, int lane) {
  uint64_t w[E];
#pragma unroll
  for (int e = 0; e < E; ++e) w[e] = v[e];
  warp_sort<E>(w, lane);
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = w[e];
}

// Spin until *p >= target (a counter another CTA of this co-resident persistent grid increments
// with a release pattern: data, __threadfence, atomicAdd).  Bounded: a protocol bug traps.
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) return;
    if (clock64() - t0 > 8000000000ll) {
      printf("vdb: hand-over wait timed out (block %d, have %d, want %d)\n", blockIdx.x, v, target);
      __trap();
    }
    __nanosleep(100);
  }
}

// raw (ordered-uint) form: convert with ord2f at the point of use, so that the scoreboard wait
// for this L2 round trip lands there and not right behind the load
__device__ __forceinline__ uint32_t ld_volatile_thr_raw(const uint32_t* p) {
  uint32_t o;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(o) : "l"(p));
  return o;
}
__device__ __forceinline__ float ld_volatile_thr(const uint32_t* p) {
  uint32_t o;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(o) : "l"(p));
  return ord2f(o);
}

}  // namespace vdb
