// Sign-projection codes and Hamming top-k: the candidate generator in front of the LSH rerank.
//   reference: faiss.IndexLSH(d, nbits).add / .search reached from
//   src/algorithms/modular.py:215-216 (FaissLSHIndexer.build) and :477 (candidate search).
//
// Hamming distances are small integers, so selection is a counting problem, not a sort:
//   1. count   per query, a histogram of the distances <= a bound T (bound from a row sample, so
//              almost every row fails one compare and touches nothing)
//   2. cut     smallest distance t* whose cumulative count reaches k, the number r of rows to take
//              from bin t*, and the output offset of every bin
//   3. emit    second pass in row order: rows below t* and the first r rows of bin t* are written
//              at offset[bin] + rank - a stable counting sort, so the output is ordered by
//              (distance, id) without sorting anything.
// One warp owns QT queries (codes in registers) and streams the row codes of one row segment with
// 128-bit loads, one row per lane (next 32 rows prefetched); the base is cut into RS segments so
// that (nq / QT) * RS warps fill the machine, and every (segment, query, bin) keeps its own count,
// so the output offsets - and with them the row order and the result - stay deterministic.
// Integer / byte work: bound by the popc + compare issue rate and L2 bandwidth.
#include <algorithm>

#include "common.cuh"

namespace vdb {

// ---------------------------------------------------------------------------------------------
// codes[row, w] bit b = [ sum_j x[row, j] * projT[j, w*32 + b] >= 0 ]; one warp per row, lane = bit.
__global__ void lsh_encode_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld,
                                  const float* __restrict__ projT /* [d][nbits_pad] */, int nbits, int nbits_pad,
                                  uint32_t* __restrict__ codes /* [n][nbits_pad / 32] */) {
  extern __shared__ float xs[];   // [warps][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= n) return;
  float* xr = xs + warp * d;
  for (int j = lane; j < d; j += 32) xr[j] = x[row * ld + j];
  __syncwarp();
  const int words = nbits_pad >> 5;
  for (int w = 0; w < words; ++w) {
    const int b = w * 32 + lane;
    float acc = 0.f;
    for (int j = 0; j < d; ++j) acc = fmaf(xr[j], __ldg(projT + static_cast<int64_t>(j) * nbits_pad + b), acc);
    const unsigned word = __ballot_sync(0xffffffffu, b < nbits && acc >= 0.f);
    if (lane == 0) codes[row * words + w] = word;
  }
}

// ---------------------------------------------------------------------------------------------
template <int W4>
__device__ __forceinline__ int hamming(const uint4 (&a)[W4], const uint4 (&b)[W4]) {
  int dist = 0;
#pragma unroll
  for (int w = 0; w < W4; ++w) {
    const unsigned long long x0 = (static_cast<unsigned long long>(a[w].x ^ b[w].x) << 32) | (a[w].y ^ b[w].y);
    const unsigned long long x1 = (static_cast<unsigned long long>(a[w].z ^ b[w].z) << 32) | (a[w].w ^ b[w].w);
    dist += __popcll(x0) + __popcll(x1);
  }
  return dist;
}

template <int W4>
constexpr int queries_per_warp() { return W4 <= 2 ? 8 : (W4 == 4 ? 4 : 2); }

// Histogram of the distances <= T[q] over the rows of segment blockIdx.y (rows [row_begin, row_end)
// cut into gridDim.y segments of seg_len rows) -> hist[seg][q][bin].  T == nullptr: every distance.
// `only` (nullable): process just the flagged queries.
template <int W4>
__global__ void __launch_bounds__(128)
hamming_count_kernel(const uint4* __restrict__ codes, int64_t row_begin, int64_t row_end, int64_t seg_len,
                     const uint4* __restrict__ qcodes, int64_t nq, int nbits, const int* __restrict__ T,
                     const uint8_t* __restrict__ only, int* __restrict__ hist /* [segs][nq][nbits + 1] */) {
  constexpr int QT = queries_per_warp<W4>();
  extern __shared__ int smem_i[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bins = nbits + 1;
  const int64_t q0 = (static_cast<int64_t>(blockIdx.x) * 4 + warp) * QT;
  if (q0 >= nq) return;
  const int64_t r0 = row_begin + static_cast<int64_t>(blockIdx.y) * seg_len;
  const int64_t r1 = min(row_end, r0 + seg_len);
  int* h = smem_i + warp * QT * bins;
  for (int i = lane; i < QT * bins; i += 32) h[i] = 0;
  uint4 qc[QT][W4];
  int tq[QT];
  bool any = false;
#pragma unroll
  for (int qi = 0; qi < QT; ++qi) {
    const int64_t q = q0 + qi;
    const bool live = q < nq && (only == nullptr || only[q] != 0);
    tq[qi] = live ? (T != nullptr ? T[q] : nbits) : -1;
    any |= live;
#pragma unroll
    for (int w = 0; w < W4; ++w) qc[qi][w] = live ? qcodes[q * W4 + w] : make_uint4(0, 0, 0, 0);
  }
  if (!any) return;
  __syncwarp();
  uint4 nxt[W4];
#pragma unroll
  for (int w = 0; w < W4; ++w) nxt[w] = r0 + lane < r1 ? __ldg(codes + (r0 + lane) * W4 + w) : make_uint4(0, 0, 0, 0);
  for (int64_t row0 = r0; row0 < r1; row0 += 32) {
    const int64_t row = row0 + lane;
    const bool valid = row < r1;
    uint4 c[W4];
#pragma unroll
    for (int w = 0; w < W4; ++w) c[w] = nxt[w];
#pragma unroll
    for (int w = 0; w < W4; ++w) nxt[w] = row + 32 < r1 ? __ldg(codes + (row + 32) * W4 + w) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
      const int dist = hamming<W4>(c, qc[qi]);
      if (valid && dist <= tq[qi]) atomicAdd(h + qi * bins + dist, 1);
    }
  }
  __syncwarp();
  int* out = hist + static_cast<int64_t>(blockIdx.y) * nq * bins;
#pragma unroll
  for (int qi = 0; qi < QT; ++qi) {
    if (tq[qi] >= 0) {
      for (int b = lane; b < bins; b += 32) out[(q0 + qi) * bins + b] = h[qi * bins + b];
    }
  }
}

// Bound from a sample of `s` rows: smallest t whose sample count, scaled to n rows, covers k with
// a 25% margin.  An estimate only - the cut kernel verifies it and asks for a recount if short.
__global__ void hamming_bound_kernel(const int* __restrict__ hist, int64_t nq, int nbits, int64_t s, int64_t n, int k,
                                     int* __restrict__ T) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int bins = nbits + 1;
  // expected sample rows among the k nearest, plus four standard deviations of that (Poisson) count:
  // a flat 25 % margin left 13 % of the queries short at k = 800 (22 expected sample rows)
  const double lam = static_cast<double>(k) * static_cast<double>(s) / static_cast<double>(n);
  const double need = lam + 4.0 * sqrt(lam) + 4.0;
  long long cum = 0;
  int t = nbits;
  for (int b = 0; b < bins; ++b) {
    cum += hist[q * bins + b];
    if (static_cast<double>(cum) >= need) { t = b; break; }
  }
  T[q] = t;
}

// Per-segment counts -> cut bin t*, and for every segment the rows it takes from bin t* (the first
// `k - (rows below t*)` rows of that bin in row order) and its output offset per bin (written back
// over the counts).  A query whose bound turned out too small is flagged for a full recount.
__global__ void hamming_cut_kernel(int* __restrict__ hist /* [segs][nq][bins] */, int segs, int64_t nq, int nbits, int k,
                                   int* __restrict__ T, const uint8_t* __restrict__ only, int* __restrict__ cut,
                                   int* __restrict__ take /* [segs][nq] */, uint8_t* __restrict__ redo) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  if (only != nullptr && only[q] == 0) { if (redo != nullptr) redo[q] = 0; return; }
  const int bins = nbits + 1;
  const int64_t seg_stride = nq * bins;
  const int bound = T != nullptr ? T[q] : nbits;
  long long cum = 0;
  int t = -1, need = 0;
  for (int b = 0; b <= bound && t < 0; ++b) {
    long long c = 0;
    for (int s = 0; s < segs; ++s) c += hist[s * seg_stride + q * bins + b];
    if (cum + c >= k) { t = b; need = static_cast<int>(k - cum); }
    else cum += c;
  }
  if (t < 0) {
    if (bound < nbits && redo != nullptr) {            // the estimate was too tight: count everything
      redo[q] = 1;
      T[q] = nbits;
      cut[q] = -1;
      for (int s = 0; s < segs; ++s) take[s * nq + q] = 0;
      return;
    }
    t = bound;                                         // fewer than k rows exist: take them all
    need = 1 << 30;
  }
  // offsets: bins in ascending order, inside a bin the segments in row order
  long long run = 0;
  for (int b = 0; b <= t; ++b) {
    for (int s = 0; s < segs; ++s) {
      int* p = hist + s * seg_stride + q * bins + b;
      const int c = *p;
      *p = static_cast<int>(min(run, static_cast<long long>(k)));
      if (b == t) {
        const int tk = min(need, c);
        take[s * nq + q] = tk;
        need -= tk;
        run += tk;
      } else {
        run += c;
      }
    }
  }
  cut[q] = t;
  if (redo != nullptr) redo[q] = 0;
}

// Emission in row order within segment blockIdx.y.  offs = hist after the cut kernel.
template <int W4>
__global__ void __launch_bounds__(128)
hamming_emit_kernel(const uint4* __restrict__ codes, int64_t n, int64_t seg_len, const uint4* __restrict__ qcodes,
                    int64_t nq, int nbits, const int* __restrict__ cut, const int* __restrict__ take,
                    const int* __restrict__ offs, int k, int64_t id_offset, float* __restrict__ out_d,
                    int64_t* __restrict__ out_i, const uint8_t* __restrict__ only) {
  constexpr int QT = queries_per_warp<W4>();
  extern __shared__ int smem_i[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int bins = nbits + 1;
  const int64_t q0 = (static_cast<int64_t>(blockIdx.x) * 4 + warp) * QT;
  if (q0 >= nq) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * seg_len;
  const int64_t r1 = min(n, r0 + seg_len);
  const int* seg_offs = offs + static_cast<int64_t>(blockIdx.y) * nq * bins;
  const int* seg_take = take + static_cast<int64_t>(blockIdx.y) * nq;
  int* cursor = smem_i + warp * QT * bins;
  uint4 qc[QT][W4];
  int cq[QT], rq[QT], seen[QT];
  if (only != nullptr) {                               // subset run: warps without a flagged query have nothing to do
    bool any = false;
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) any |= q0 + qi < nq && only[q0 + qi] != 0;
    if (!any) return;
  }
#pragma unroll
  for (int qi = 0; qi < QT; ++qi) {
    const int64_t q = q0 + qi;
    const bool live = q < nq && (only == nullptr || only[q] != 0);
    cq[qi] = live ? cut[q] : -1;
    rq[qi] = live ? seg_take[q] : 0;
    seen[qi] = 0;
    for (int b = lane; b < bins; b += 32) cursor[qi * bins + b] = live ? seg_offs[q * bins + b] : 0;
#pragma unroll
    for (int w = 0; w < W4; ++w) qc[qi][w] = live ? qcodes[q * W4 + w] : make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  uint4 nxt[W4];
#pragma unroll
  for (int w = 0; w < W4; ++w) nxt[w] = r0 + lane < r1 ? __ldg(codes + (r0 + lane) * W4 + w) : make_uint4(0, 0, 0, 0);
  for (int64_t row0 = r0; row0 < r1; row0 += 32) {
    const int64_t row = row0 + lane;
    const bool valid = row < r1;
    uint4 c[W4];
#pragma unroll
    for (int w = 0; w < W4; ++w) c[w] = nxt[w];
#pragma unroll
    for (int w = 0; w < W4; ++w) nxt[w] = row + 32 < r1 ? __ldg(codes + (row + 32) * W4 + w) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
      const int dist = hamming<W4>(c, qc[qi]);
      const bool below = valid && dist < cq[qi];
      const bool tie = valid && dist == cq[qi];
      const unsigned tie_m = __ballot_sync(0xffffffffu, tie);
      const bool takes = below || (tie && seen[qi] + __popc(tie_m & lt_mask) < rq[qi]);
      seen[qi] += __popc(tie_m);
      const unsigned take_m = __ballot_sync(0xffffffffu, takes);
      if (take_m == 0u) continue;
      int peers_n = 0, rank = 0, slot = 0;
      if (takes) {
        const unsigned peers = __match_any_sync(take_m, dist);
        rank = __popc(peers & lt_mask);
        peers_n = __popc(peers);
        slot = cursor[qi * bins + dist] + rank;
        if (slot < k) {
          out_d[(q0 + qi) * k + slot] = static_cast<float>(dist);
          out_i[(q0 + qi) * k + slot] = row + id_offset;
        }
      }
      __syncwarp();
      if (takes && rank == 0) cursor[qi * bins + dist] += peers_n;
      __syncwarp();
    }
  }
}

static int hamming_segments(int64_t n, int64_t nq, int qt) {
  const int64_t warps = (nq + qt - 1) / qt;
  int64_t segs = (8192 + warps - 1) / warps;                 // aim at ~8k warps in flight
  segs = std::min<int64_t>(segs, std::max<int64_t>(1, n / 32768));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(segs, 32)));
}

template <int W4>
static int run_hamming(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                       int64_t id_offset, float* out_d, int64_t* out_i, void* ws, cudaStream_t s) {
  constexpr int QT = queries_per_warp<W4>();
  const int bins = nbits + 1;
  const int segs = hamming_segments(n, nq, QT);
  const int64_t seg_len = ((n + segs - 1) / segs + 31) / 32 * 32;
  uint8_t* w = static_cast<uint8_t*>(ws);
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  int* hist = reinterpret_cast<int*>(w);                     // [segs][nq][bins]
  size_t off = al(static_cast<size_t>(segs) * nq * bins * 4);
  int* take = reinterpret_cast<int*>(w + off); off += al(static_cast<size_t>(segs) * nq * 4);
  int* T = reinterpret_cast<int*>(w + off);    off += al(static_cast<size_t>(nq) * 4);
  int* cut = reinterpret_cast<int*>(w + off);  off += al(static_cast<size_t>(nq) * 4);
  uint8_t* redo = w + off;
  const uint4* c4 = reinterpret_cast<const uint4*>(codes);
  const uint4* q4 = reinterpret_cast<const uint4*>(qcodes);
  const unsigned blocks = static_cast<unsigned>((nq + 4 * QT - 1) / (4 * QT));
  const unsigned qblocks = static_cast<unsigned>((nq + 255) / 256);
  const size_t smem = static_cast<size_t>(4) * QT * bins * sizeof(int);
  auto count = hamming_count_kernel<W4>;
  auto emit = hamming_emit_kernel<W4>;
  if (smem > 48 * 1024) {
    VDB_CHECK_CUDA(cudaFuncSetAttribute(count, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    VDB_CHECK_CUDA(cudaFuncSetAttribute(emit, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const dim3 grid(blocks, segs);
  constexpr int64_t kSample = 32768;
  if (n <= 2 * kSample) {            // small base: one exact count
    count<<<grid, 128, smem, s>>>(c4, 0, n, seg_len, q4, nq, nbits, nullptr, nullptr, hist);
    hamming_cut_kernel<<<qblocks, 256, 0, s>>>(hist, segs, nq, nbits, k, nullptr, nullptr, cut, take, nullptr);
    count_launches(2);
  } else {
    count<<<dim3(blocks, 1), 128, smem, s>>>(c4, 0, kSample, kSample, q4, nq, nbits, nullptr, nullptr, hist);
    hamming_bound_kernel<<<qblocks, 256, 0, s>>>(hist, nq, nbits, kSample, n, k, T);
    count<<<grid, 128, smem, s>>>(c4, 0, n, seg_len, q4, nq, nbits, T, nullptr, hist);
    hamming_cut_kernel<<<qblocks, 256, 0, s>>>(hist, segs, nq, nbits, k, T, nullptr, cut, take, redo);
    // queries whose bound was short are recounted without a bound (warps without such a query exit at once)
    count<<<grid, 128, smem, s>>>(c4, 0, n, seg_len, q4, nq, nbits, T, redo, hist);
    hamming_cut_kernel<<<qblocks, 256, 0, s>>>(hist, segs, nq, nbits, k, T, redo, cut, take, nullptr);
    count_launches(6);
  }
  emit<<<grid, 128, smem, s>>>(c4, n, seg_len, q4, nq, nbits, cut, take, hist, k, id_offset, out_d, out_i, nullptr);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int W4>
static int sample_bound_impl(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                             int* hist, int* T, cudaStream_t s) {
  constexpr int QT = queries_per_warp<W4>();
  constexpr int64_t kSample = 32768;
  const int bins = nbits + 1;
  const unsigned blocks = static_cast<unsigned>((nq + 4 * QT - 1) / (4 * QT));
  const size_t smem = static_cast<size_t>(4) * QT * bins * sizeof(int);
  auto count = hamming_count_kernel<W4>;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(count, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t rows = std::min<int64_t>(n, kSample);
  count<<<dim3(blocks, 1), 128, smem, s>>>(reinterpret_cast<const uint4*>(codes), 0, rows, rows,
                                           reinterpret_cast<const uint4*>(qcodes), nq, nbits, nullptr, nullptr, hist);
  hamming_bound_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, s>>>(hist, nq, nbits, rows, n, k, T);
  count_launches(2);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// exact popc path for the queries flagged in `only` (all others are left untouched): full count, cut, emission
template <int W4>
static int subset_impl(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                       int64_t id_offset, float* out_d, int64_t* out_i, void* ws, const uint8_t* only, cudaStream_t s) {
  constexpr int QT = queries_per_warp<W4>();
  const int bins = nbits + 1;
  const int segs = hamming_segments(n, nq, QT);
  const int64_t seg_len = ((n + segs - 1) / segs + 31) / 32 * 32;
  uint8_t* w = static_cast<uint8_t*>(ws);
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  int* hist = reinterpret_cast<int*>(w);
  size_t off = al(static_cast<size_t>(segs) * nq * bins * 4);
  int* take = reinterpret_cast<int*>(w + off); off += al(static_cast<size_t>(segs) * nq * 4);
  off += al(static_cast<size_t>(nq) * 4);
  int* cut = reinterpret_cast<int*>(w + off);
  const uint4* c4 = reinterpret_cast<const uint4*>(codes);
  const uint4* q4 = reinterpret_cast<const uint4*>(qcodes);
  const unsigned blocks = static_cast<unsigned>((nq + 4 * QT - 1) / (4 * QT));
  const unsigned qblocks = static_cast<unsigned>((nq + 255) / 256);
  const size_t smem = static_cast<size_t>(4) * QT * bins * sizeof(int);
  auto count = hamming_count_kernel<W4>;
  auto emit = hamming_emit_kernel<W4>;
  if (smem > 48 * 1024) {
    VDB_CHECK_CUDA(cudaFuncSetAttribute(count, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    VDB_CHECK_CUDA(cudaFuncSetAttribute(emit, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const dim3 grid(blocks, segs);
  count<<<grid, 128, smem, s>>>(c4, 0, n, seg_len, q4, nq, nbits, nullptr, only, hist);
  hamming_cut_kernel<<<qblocks, 256, 0, s>>>(hist, segs, nq, nbits, k, nullptr, only, cut, take, nullptr);
  emit<<<grid, 128, smem, s>>>(c4, n, seg_len, q4, nq, nbits, cut, take, hist, k, id_offset, out_d, out_i, only);
  count_launches(3);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int hamming_topk_subset(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                        int64_t id_offset, float* out_d, int64_t* out_i, void* ws, const uint8_t* only, cudaStream_t stream) {
  switch (vdb_lsh_code_words(nbits) / 4) {
    case 1: return subset_impl<1>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, ws, only, stream);
    case 2: return subset_impl<2>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, ws, only, stream);
    default: set_error("hamming_topk_subset: nbits %d not supported", nbits); return 2;
  }
}

int hamming_sample_bound(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                         int* hist_scratch, int* T, cudaStream_t stream) {
  switch (vdb_lsh_code_words(nbits) / 4) {
    case 1: return sample_bound_impl<1>(codes, n, qcodes, nq, nbits, k, hist_scratch, T, stream);
    case 2: return sample_bound_impl<2>(codes, n, qcodes, nq, nbits, k, hist_scratch, T, stream);
    default: set_error("hamming_sample_bound: nbits %d not supported", nbits); return 2;
  }
}

}  // namespace vdb

using namespace vdb;

extern "C" {

int vdb_lsh_code_words(int nbits) { return (nbits + 127) / 128 * 4; }

int vdb_lsh_encode(const float* x, int64_t n, int d, int64_t ld, const float* proj_t, int nbits, uint32_t* codes,
                   void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && ld >= d && nbits > 0 && nbits <= 1024, "vdb_lsh_encode: bad shape n=%lld d=%d nbits=%d",
              (long long)n, d, nbits);
  const int nbits_pad = vdb_lsh_code_words(nbits) * 32;
  const int warps = 8;
  const size_t smem = static_cast<size_t>(warps) * d * sizeof(float);
  VDB_REQUIRE(smem <= 48 * 1024, "vdb_lsh_encode: d=%d too large", d);
  lsh_encode_kernel<<<static_cast<unsigned>((n + warps - 1) / warps), warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      x, n, d, ld, proj_t, nbits, nbits_pad, codes);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t vdb_hamming_topk_workspace_bytes(int64_t nq, int nbits) {
  if (nq <= 0 || nbits <= 0) return 0;
  const size_t segs = 32;                                     // upper bound of hamming_segments()
  const size_t a = (segs * static_cast<size_t>(nq) * (nbits + 1) * 4 + 255) & ~size_t(255);
  const size_t t = (segs * static_cast<size_t>(nq) * 4 + 255) & ~size_t(255);
  const size_t b = (static_cast<size_t>(nq) * 4 + 255) & ~size_t(255);
  return a + t + 2 * b + ((static_cast<size_t>(nq) + 255) & ~size_t(255));
}

int vdb_hamming_topk(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits, int k,
                     int64_t id_offset, float* out_d, int64_t* out_i, void* workspace, size_t workspace_bytes,
                     void* stream) {
  VDB_REQUIRE(n > 0 && nq > 0 && nbits > 0 && nbits <= 1024 && k >= 1, "vdb_hamming_topk: bad shape");
  VDB_REQUIRE(workspace != nullptr && workspace_bytes >= vdb_hamming_topk_workspace_bytes(nq, nbits),
              "vdb_hamming_topk: workspace too small");
  VDB_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0 && (reinterpret_cast<uintptr_t>(qcodes) & 15) == 0,
              "vdb_hamming_topk: codes must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (vdb_lsh_code_words(nbits) / 4) {
    case 1: return run_hamming<1>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace, s);
    case 2: return run_hamming<2>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace, s);
    case 3: case 4: {
      VDB_REQUIRE(vdb_lsh_code_words(nbits) == 16, "vdb_hamming_topk: nbits in (256, 512] must be padded to 512 by the caller");
      return run_hamming<4>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace, s);
    }
    default: {
      VDB_REQUIRE(vdb_lsh_code_words(nbits) == 32, "vdb_hamming_topk: nbits in (512, 1024] must be padded to 1024 by the caller");
      return run_hamming<8>(codes, n, qcodes, nq, nbits, k, id_offset, out_d, out_i, workspace, s);
    }
  }
}

}  // extern "C"
