// Product quantisation: "PQ<m>" and "IVF<n>,PQ<m>" (8 bits per sub-quantiser): encoding, list scan with per-list
// look-up tables (asymmetric distance computation) + top-k.
//   reference: FaissFactoryIndexer(index_key="IVF256,PQ64" | "PQ64" | ...) + FaissSearcher
//   src/algorithms/modular.py:224-286,536-548; configs/benchmark_config.yaml:36-50,61-72
// FAISS semantics restated [FAISS-upstream, parity unpinned]: the d dimensions are cut into M sub-vectors of dsub = d / M
// dimensions, each with a codebook of 256 centroids (k-means, trained by the host side with the IVF k-means recipe); a row is
// the M bytes of its nearest sub-centroids.  IndexIVFPQ encodes the residual x - centroid(x) (by_residual), IndexPQ the row
// itself (here: one list, zero centroid).  Search (asymmetric distance computation): |q - c - r^|^2 = |q - c|^2 + (|r^|^2 + 2 c.r^)
// - 2 q.r^: the bracket is one float per row, fixed at build time; -2 q.r^ = sum_m T[m][code_m] with ONE table per query,
// T[m][j] = -2 q_m . cb[m][j]; |q - c|^2 is a scalar per probed list.  Inner product: q.c + sum_m q_m . cb[m][code_m].
// Lists use the byte layout of sq8.cu with d := M (block = 32 rows as uint4 [ceil(M/16)][32 lanes]).  The scan is bound by
// shared-memory gathers (one 4-byte look-up per code byte), not by HBM.
#include "select.cuh"

namespace vdb {

// codes[row, m] = argmin_j |x[row, m*dsub : (m+1)*dsub] - cb[m][j]|^2 (ties: lowest j).  grid (row slabs of 256, M).
__global__ void __launch_bounds__(256)
pq_encode_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, const float* __restrict__ cb /* [M][256][dsub] */, int M,
                 int dsub, uint8_t* __restrict__ codes /* [n][M] */) {
  extern __shared__ float cbs[];                               // [256][dsub] of sub-quantiser m
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < 256 * dsub; i += blockDim.x) cbs[i] = cb[static_cast<int64_t>(m) * 256 * dsub + i];
  __syncthreads();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (row >= n) return;
  const float* v = x + row * ld + m * dsub;
  float best = CUDART_INF_F;
  int arg = 0;
  for (int j = 0; j < 256; ++j) {
    float acc = 0.f;
    for (int t = 0; t < dsub; ++t) {
      const float df = v[t] - cbs[j * dsub + t];
      acc = fmaf(df, df, acc);
    }
    if (acc < best) { best = acc; arg = j; }
  }
  codes[row * M + m] = static_cast<uint8_t>(arg);
}

// bias[row] = |r^|^2 + 2 c . r^ for the row's reconstructed residual r^ = decode(code) and its list centroid c (zero for
// IndexPQ): with it  |q - c - r^|^2 = |q - c|^2 + bias - 2 q . r^,  so the scan needs ONE table per query (-2 q_s . cb[s][j])
// and a scalar per probed list instead of a table per (query, list) - FAISS's precomputed-table decomposition with the
// list-dependent term folded into a per-row float.  One warp per row, fp64 accumulation.
__global__ void pq_bias_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int dsub, const float* __restrict__ cb,
                               const float* __restrict__ cent, const int32_t* __restrict__ assign, float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int d = M * dsub;
  const float* c = cent != nullptr ? cent + static_cast<int64_t>(assign[row]) * d : nullptr;
  double acc = 0.0;
  for (int s = lane; s < M; s += 32) {
    const float* v = cb + (static_cast<int64_t>(s) * 256 + codes[row * M + s]) * dsub;
    for (int t = 0; t < dsub; ++t) {
      const double r = v[t], cc = c != nullptr ? c[s * dsub + t] : 0.0;
      acc += r * r + 2.0 * cc * r;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bias[row] = static_cast<float>(acc);
}

// out[row, j] = cb[s][codes[row, s]][j - s * dsub] (+ cent[assign[row]][j]): the vector a code stands for.  thread = (row, j);
// the codebooks (m * 256 * dsub floats: 51 KB at d = 50) stay in L1 / L2.
__global__ void pq_decode_kernel(const uint8_t* __restrict__ codes, int64_t n, int d, int M, int dsub, const float* __restrict__ cb,
                                 const float* __restrict__ cent, const int32_t* __restrict__ assign, float* __restrict__ out,
                                 int64_t ld_out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int64_t row = i / d;
  const int j = static_cast<int>(i - row * d);
  const int s = j / dsub;
  float v = __ldg(cb + (static_cast<int64_t>(s) * 256 + codes[row * M + s]) * dsub + (j - s * dsub));
  if (cent != nullptr) v += __ldg(cent + static_cast<int64_t>(assign[row]) * d + j);
  out[row * ld_out + j] = v;
}

// scatter precomputed byte rows (and, optionally, one float per row) into the interleaved byte lists (layout of sq8.cu with
// d := M); one warp per row
__global__ void bytes_fill_kernel(const uint8_t* __restrict__ rows, int64_t n, int M, const int32_t* __restrict__ assign,
                                  const int32_t* __restrict__ blk_off, int nlist, int32_t* cursor, uint8_t* __restrict__ lists,
                                  int32_t* __restrict__ ids, const float* __restrict__ row_values, float* __restrict__ list_values) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int l = assign[row];
  if (l < 0 || l >= nlist) return;
  int slot = 0;
  if (lane == 0) slot = atomicAdd(cursor + l, 1);
  slot = __shfl_sync(0xffffffffu, slot, 0);
  const int64_t b = blk_off[l] + (slot >> 5);
  const int v = slot & 31;
  const int m16 = (M + 15) / 16;
  for (int c = lane; c < m16 * 16; c += 32)
    lists[((b * m16 + (c >> 4)) * 32 + v) * 16 + (c & 15)] = c < M ? rows[row * M + c] : 0;
  if (lane == 0) {
    ids[b * 32 + v] = static_cast<int32_t>(row);
    if (row_values != nullptr) list_values[b * 32 + v] = row_values[row];
  }
}

// One query per CTA (TW warps).  The query's table T[s][j] = -2 q_s . cb[s][j] (inner product: q_s . cb[s][j]) is built ONCE
// in shared memory ([m16*16][256] floats; the padding sub-quantisers hold zeros, as do their code bytes); per probed list
// every warp computes the scalar |q - c|^2 (q . c) for itself - no CTA barrier inside the probe loop - and a row scores
//   L2: |q - c|^2 + bias_row + sum_s T[s][code_s]        IP: -(q . c + sum_s T[s][code_s])
template <int KP, int TW>
__global__ void __launch_bounds__(TW * 32)
ivf_pq_scan_kernel(int metric, const uint4* __restrict__ lists, const int32_t* __restrict__ ids, const float* __restrict__ list_bias,
                   const int32_t* __restrict__ blk_off, int nlist, int d, int M, int dsub,
                   const float* __restrict__ cent /* [nlist][d] or nullptr = zero */, const float* __restrict__ cb,
                   const int64_t* __restrict__ probes, int nprobe, const float* __restrict__ qmat, int64_t ld_q, int k, int flags,
                   float pad_value, int64_t id_offset, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  constexpr int CAP = pool_cap(KP);
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint64_t* pools = reinterpret_cast<uint64_t*>(smem_dyn);
  int* cnts = reinterpret_cast<int*>(pools + TW * CAP);
  float* thr_s = reinterpret_cast<float*>(cnts + TW);
  float* qs = thr_s + TW;                                      // [d] the query
  const int m16 = (M + 15) / 16;
  float* lut = qs + ((d + 3) & ~3);                            // [m16 * 16][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = blockIdx.x;
  const bool l2 = metric == VDB_METRIC_L2;
  for (int j = threadIdx.x; j < d; j += TW * 32) qs[j] = qmat[q * ld_q + j];
  for (int i = threadIdx.x; i < (m16 * 16 - M) * 256; i += TW * 32) lut[M * 256 + i] = 0.f;   // padding sub-quantisers
  __syncthreads();
  for (int e = threadIdx.x; e < M * 256; e += TW * 32) {       // table entry (s, j)
    const float* c = cb + static_cast<int64_t>(e) * dsub;
    const float* qq = qs + (e >> 8) * dsub;
    float acc = 0.f;
    for (int t = 0; t < dsub; ++t) acc = fmaf(qq[t], c[t], acc);
    lut[e] = l2 ? -2.f * acc : acc;
  }
  __syncthreads();
  WarpTopK<KP> sel;
  sel.init(pools + warp * CAP);
  int turn = 0;
  for (int pi = 0; pi < nprobe; ++pi) {
    const int64_t l = probes != nullptr ? probes[q * nprobe + pi] : 0;
    if (l < 0 || l >= nlist) continue;
    const int b0 = blk_off[l], b1 = blk_off[l + 1];
    const int first = b0 + ((warp - turn) % TW + TW) % TW;     // blocks of all probed lists are dealt round-robin to the warps
    turn = (turn + (b1 - b0)) % TW;
    if (first >= b1) continue;                                 // nothing of this list for this warp
    float part = 0.f;                                          // |q - c|^2 (L2) or q . c (inner product), per warp
    for (int j = lane; j < d; j += 32) {
      const float cv = cent != nullptr ? __ldg(cent + l * d + j) : 0.f;
      if (l2) { const float df = qs[j] - cv; part = fmaf(df, df, part); }
      else part = fmaf(qs[j], cv, part);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    for (int b = first; b < b1; b += TW) {
      const int id = ids[static_cast<int64_t>(b) * 32 + lane];
      const float bias = l2 ? list_bias[static_cast<int64_t>(b) * 32 + lane] : 0.f;
      const uint4* p = lists + static_cast<int64_t>(b) * m16 * 32 + lane;
      float acc = 0.f;
      for (int c0 = 0; c0 < m16; c0 += 4) {                    // four 128-bit loads (64 sub-quantisers) in flight per lane
        uint4 u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = c0 + i < m16 ? __ldg(p + (c0 + i) * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + i >= m16) break;
          const uint32_t wds[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
          const float* t0 = lut + (c0 + i) * 16 * 256;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float* th = t0 + h * 4 * 256;
            acc += th[wds[h] & 0xffu] + th[256 + ((wds[h] >> 8) & 0xffu)] + th[512 + ((wds[h] >> 16) & 0xffu)] + th[768 + (wds[h] >> 24)];
          }
        }
      }
      const float key = l2 ? fmaxf(part + bias + acc, 0.f) : -(part + acc);
      sel.push(id >= 0, key, static_cast<uint32_t>(id), lane);
    }
  }
  cta_write_topk<KP, TW>(sel, pools, cnts, thr_s, warp, lane, 1, metric, k, flags, pad_value, id_offset, out_d + q * k, out_i + q * k);
}

template <int KP, int TW>
static int launch_pq_scan(int metric, const uint8_t* lists, const int32_t* ids, const float* list_bias, const int32_t* blk_off,
                          int nlist, int d, int M, const float* cent, const float* cb, const int64_t* probes, int nprobe, const float* q, int64_t ld_q,
                          int64_t nq, int k, int flags, float pad_value, int64_t id_offset, float* out_d, int64_t* out_i,
                          cudaStream_t stream) {
  const int m16 = (M + 15) / 16;
  const size_t smem = static_cast<size_t>(TW) * pool_cap(KP) * 8 + TW * 8 + static_cast<size_t>((d + 3) & ~3) * 4 +
                      static_cast<size_t>(m16) * 16 * 256 * 4;
  VDB_REQUIRE(smem <= 220 * 1024, "vdb_ivf_pq_scan_topk: look-up table does not fit shared memory (M = %d, k = %d)", M, k);
  auto kern = ivf_pq_scan_kernel<KP, TW>;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>(nq), TW * 32, smem, stream>>>(metric, reinterpret_cast<const uint4*>(lists), ids, list_bias, blk_off, nlist,
                                                              d, M, d / M, cent, cb, probes, nprobe, q, ld_q, k, flags, pad_value,
                                                              id_offset, out_d, out_i);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vdb

using namespace vdb;

extern "C" {

int vdb_pq_encode(const float* x, int64_t n, int d, int64_t ld, const float* codebooks, int m, uint8_t* codes, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && m > 0 && d % m == 0 && ld >= d, "vdb_pq_encode: bad shape (d must be a multiple of m)");
  const int dsub = d / m;
  VDB_REQUIRE(256 * dsub * 4 <= 96 * 1024, "vdb_pq_encode: sub-vector too long (dsub = %d)", dsub);
  const size_t smem = static_cast<size_t>(256) * dsub * 4;
  auto kern = pq_encode_kernel;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(m));
  kern<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, n, d, ld, codebooks, m, dsub, codes);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_pq_bias(const uint8_t* codes, int64_t n, int d, int m, const float* codebooks, const float* centroids, const int32_t* assign,
                float* bias, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && m > 0 && d % m == 0 && (centroids == nullptr || assign != nullptr), "vdb_pq_bias: bad arguments");
  pq_bias_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, n, m, d / m, codebooks,
                                                                                                        centroids, assign, bias);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_pq_decode(const uint8_t* codes, int64_t n, int d, int m, const float* codebooks, const float* centroids, const int32_t* assign,
                  float* out, int64_t ld_out, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && m > 0 && d % m == 0 && ld_out >= d && (centroids == nullptr || assign != nullptr),
              "vdb_pq_decode: bad arguments");
  pq_decode_kernel<<<static_cast<unsigned>((n * d + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      codes, n, d, m, d / m, codebooks, centroids, assign, out, ld_out);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_bytes_fill(const uint8_t* rows, int64_t n, int m, const int32_t* assign, const int32_t* blk_off, int nlist, int32_t* cursor,
                   uint8_t* lists, int32_t* list_ids, const float* row_values, float* list_values, void* stream) {
  VDB_REQUIRE(n > 0 && n < (int64_t(1) << 31) && m > 0 && nlist > 0, "vdb_bytes_fill: bad shape");
  VDB_REQUIRE((row_values == nullptr) == (list_values == nullptr), "vdb_bytes_fill: row_values and list_values go together");
  bytes_fill_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rows, n, m, assign, blk_off, nlist, cursor, lists, list_ids, row_values, list_values);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_ivf_pq_scan_topk(int metric, const uint8_t* lists, const int32_t* list_ids, const float* list_bias, const int32_t* blk_off,
                         int nlist, int d, int m, const float* centroids, const float* codebooks, const int64_t* probes, int nprobe, const float* q,
                         int64_t ld_q, int64_t nq, int k, int flags, float pad_value, int64_t id_offset, float* out_d,
                         int64_t* out_i, void* stream) {
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_ivf_pq_scan_topk: bad metric %d", metric);
  VDB_REQUIRE(nq > 0 && d > 0 && m > 0 && d % m == 0 && nlist > 0 && nprobe >= 1 && ld_q >= d, "vdb_ivf_pq_scan_topk: bad shape");
  VDB_REQUIRE(probes != nullptr || (nlist == 1 && nprobe == 1), "vdb_ivf_pq_scan_topk: probes may be null only for a single list");
  VDB_REQUIRE(metric != VDB_METRIC_L2 || list_bias != nullptr, "vdb_ivf_pq_scan_topk: L2 needs the per-row bias (vdb_pq_bias)");
  VDB_REQUIRE((reinterpret_cast<uintptr_t>(lists) & 15) == 0, "vdb_ivf_pq_scan_topk: lists must be 16-byte aligned");
  const int kp = k <= 32 ? 32 : k <= 128 ? 128 : k <= 256 ? 256 : k <= 512 ? 512 : 0;
  VDB_REQUIRE(k >= 1 && kp != 0, "vdb_ivf_pq_scan_topk: k=%d unsupported (1..512)", k);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define VDB_GO(KP, TW)                                                                                                               \
  return launch_pq_scan<KP, TW>(metric, lists, list_ids, list_bias, blk_off, nlist, d, m, centroids, codebooks, probes, nprobe, q, ld_q, \
                                nq, k, flags, pad_value, id_offset, out_d, out_i, s)
  switch (kp) {
    case 32: VDB_GO(32, 8);
    case 128: VDB_GO(128, 8);
    case 256: VDB_GO(256, 8);
    default: VDB_GO(512, 4);
  }
#undef VDB_GO
}

}  // extern "C"
