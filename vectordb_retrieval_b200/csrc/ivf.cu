// IVF-Flat: inverted-list construction and the list-scan + top-k kernel.
//   reference: faiss.index_factory("IVF<n>,Flat") add/search reached from
//   src/algorithms/modular.py:277-286,536-548 and src/algorithms/approximate_search.py:39-51,87
// Coarse assignment (nearest / top-nprobe centroids) is vdb_flat_topk with base := centroids.
//
// List layout "interleaved-32": block = 32 vectors stored as float4 [d4][32]; a warp reads one
// 512-byte line per 4 dimensions with 128-bit loads, one vector per lane, no cross-lane
// reduction.  HBM-bound: algorithmic bytes = scanned rows * d * 4.
#include <cstdlib>

#include "select.cuh"

namespace vdb {

__global__ void ivf_count_kernel(const int32_t* __restrict__ assign, int64_t n, int nlist, int32_t* counts) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const int l = assign[i];
    if (l >= 0 && l < nlist) atomicAdd(counts + l, 1);
  }
}

// one warp per row; slot order inside a list is arbitrary (results never depend on it: the
// scan orders by (distance, id))
__global__ void ivf_fill_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, const int32_t* __restrict__ assign,
                                const int32_t* __restrict__ blk_off, int nlist, int32_t* cursor, float* __restrict__ vecs,
                                int32_t* __restrict__ ids) {
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
  const int d4 = (d + 3) / 4;
  for (int c = lane; c < d4 * 4; c += 32)
    vecs[((b * d4 + (c >> 2)) * 32 + v) * 4 + (c & 3)] = c < d ? x[row * ld + c] : 0.f;
  if (lane == 0) ids[b * 32 + v] = static_cast<int32_t>(row);
}

// k-means centroid update: sums[l, :] += x[row, :], counts[l] += 1 for l = assign[row]; one warp per row.
__global__ void kmeans_accumulate_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld,
                                         const int64_t* __restrict__ assign, int nlist, float* sums, int32_t* counts) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int64_t l = assign[row];
  if (l < 0 || l >= nlist) return;
  for (int j = lane; j < d; j += 32) atomicAdd(sums + l * d + j, x[row * ld + j]);
  if (lane == 0) atomicAdd(counts + l, 1);
}

// W warps serve one query and a CTA of TW warps holds TW / W queries: one query per 8 warps (W = TW) suits long
// scans; at nprobe 1-8 over ~300-row lists that leaves a warp one or two 32-row blocks, no pool ever fills and the
// per-query fixed work (staging the query, two barriers, the merge by one warp) dominates - 0.30 of the HBM rate at
// nprobe 1.  The host picks W from the expected rows per query (~512 rows per warp).
template <int KP, int W, int TW>
__global__ void __launch_bounds__(TW * 32, 1024 / (TW * 32))
ivf_scan_kernel(int metric, const float4* __restrict__ vecs, const int32_t* __restrict__ ids,
                const int32_t* __restrict__ blk_off, int nlist, int d4, const int64_t* __restrict__ probes, int nprobe,
                const float* __restrict__ qmat, int64_t ld_q, int64_t nq, int d, int k, int flags, float pad_value,
                int64_t id_offset, float* __restrict__ out_d, int64_t* __restrict__ out_i, unsigned long long* scanned) {
  constexpr int CAP = pool_cap(KP);
  constexpr int QPC = TW / W;                 // queries per CTA
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint64_t* pools_all = reinterpret_cast<uint64_t*>(smem_dyn);
  int* cnts_all = reinterpret_cast<int*>(pools_all + TW * CAP);
  float* thr_all = reinterpret_cast<float*>(cnts_all + TW);
  const int d4p = (d4 + 7) & ~7;
  float4* qs_all = reinterpret_cast<float4*>(thr_all + TW);      // TW is a multiple of 4: stays 16-byte aligned
  const int warp_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp_cta / W, warp = warp_cta % W;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * QPC + group;
  if (q >= nq) return;                        // whole groups leave together: the group barriers below stay balanced
  uint64_t* pools = pools_all + group * W * CAP;
  int* cnts = cnts_all + group * W;
  float* thr_s = thr_all + group * W;
  float4* qs = qs_all + group * d4p;
  const int bar_id = 1 + group;
  for (int j = warp * 32 + lane; j < d4p * 4; j += W * 32) reinterpret_cast<float*>(qs)[j] = j < d ? qmat[q * ld_q + j] : 0.f;
  group_sync<W * 32>(bar_id);
  WarpTopK<KP> sel;
  sel.init(pools + warp * CAP);
  unsigned rows_seen = 0;
  int turn = 0;   // blocks of all probed lists are dealt round-robin to the W warps
  for (int pi = 0; pi < nprobe; ++pi) {
    const int64_t l = probes[q * nprobe + pi];
    if (l < 0 || l >= nlist) continue;
    const int b0 = blk_off[l], b1 = blk_off[l + 1];
    int b = b0 + ((warp - turn) % W + W) % W;
    turn = (turn + (b1 - b0)) % W;
    for (; b < b1; b += W) {
      const int id = ids[static_cast<int64_t>(b) * 32 + lane];
      const float4* p = vecs + static_cast<int64_t>(b) * d4 * 32 + lane;
      double acc = 0.0;
      // batches of 8 independent 128-bit loads per lane (4 KB of the block in flight per warp); the
      // tail batch is predicated, the query tile in smem is zero-padded to a multiple of 8
      for (int c0 = 0; c0 < d4; c0 += 8) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = c0 + u < d4 ? __ldg(p + (c0 + u) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 y = qs[c0 + u];
          if (metric == VDB_METRIC_L2) {
            // four terms in fp32 (fused multiply-adds), then into the fp64 accumulator: one conversion
            // per 16 bytes instead of four, relative error of the sum stays ~1e-7
            const float d0 = x[u].x - y.x, d1 = x[u].y - y.y, d2 = x[u].z - y.z, d3 = x[u].w - y.w;
            acc += static_cast<double>(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3))));
          } else {
            acc += static_cast<double>(fmaf(x[u].x, y.x, fmaf(x[u].y, y.y, fmaf(x[u].z, y.z, x[u].w * y.w))));
          }
        }
      }
      const bool valid = id >= 0;
      rows_seen += __popc(__ballot_sync(0xffffffffu, valid));
      const float key = metric == VDB_METRIC_L2 ? static_cast<float>(acc) : -static_cast<float>(acc);
      sel.push(valid, key, static_cast<uint32_t>(id), lane);
    }
  }
  if (scanned != nullptr && lane == 0 && rows_seen) atomicAdd(scanned, static_cast<unsigned long long>(rows_seen));
  cta_write_topk<KP, W>(sel, pools, cnts, thr_s, warp, lane, bar_id, metric, k, flags, pad_value, id_offset, out_d + q * k,
                        out_i + q * k);
}

template <int KP, int W, int TW>
static int launch_scan(int metric, const float* vecs, const int32_t* ids, const int32_t* blk_off, int nlist, int d,
                       const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq, int k, int flags,
                       float pad_value, int64_t id_offset, float* out_d, int64_t* out_i, int64_t* scanned,
                       cudaStream_t stream) {
  const int d4 = (d + 3) / 4;
  constexpr int QPC = TW / W;
  const size_t smem = static_cast<size_t>(TW) * pool_cap(KP) * 8 + TW * 8 + static_cast<size_t>(QPC) * ((d4 + 7) & ~7) * 16;
  auto kern = ivf_scan_kernel<KP, W, TW>;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>((nq + QPC - 1) / QPC), TW * 32, smem, stream>>>(
      metric, reinterpret_cast<const float4*>(vecs), ids, blk_off, nlist, d4, probes, nprobe, q, ld_q, nq, d, k, flags,
      pad_value, id_offset, out_d, out_i, reinterpret_cast<unsigned long long*>(scanned));
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// warps per query for an expected number of scanned rows per query: about 4 096 rows per warp, a power of two <= tw.
// Measured on 1.2M x 50, nlist 4096, 10k queries (scripts/perf_ivf.py): one warp per query is fastest up to nprobe 8
// (2 400 rows: 0.77 ms against 1.09 ms with eight), two at nprobe 16-32, four at nprobe 64.
static int warps_for_rows(int64_t rows, int tw) {
  int w = 1;
  while (w < tw && static_cast<int64_t>(w) * 4096 < rows) w *= 2;
  return w;
}

}  // namespace vdb

using namespace vdb;

extern "C" {

int vdb_ivf_d4(int d) { return (d + 3) / 4; }

int vdb_ivf_count(const int32_t* assign, int64_t n, int nlist, int32_t* counts, void* stream) {
  VDB_REQUIRE(n > 0 && nlist > 0, "vdb_ivf_count: bad shape");
  ivf_count_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(assign, n, nlist, counts);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_kmeans_accumulate(const float* x, int64_t n, int d, int64_t ld, const int64_t* assign, int nlist, float* sums,
                          int32_t* counts, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && ld >= d && nlist > 0, "vdb_kmeans_accumulate: bad shape");
  kmeans_accumulate_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, d, ld, assign, nlist, sums, counts);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_ivf_fill(const float* x, int64_t n, int d, int64_t ld, const int32_t* assign, const int32_t* blk_off, int nlist,
                 int32_t* cursor, float* list_vecs, int32_t* list_ids, void* stream) {
  VDB_REQUIRE(n > 0 && n < (int64_t(1) << 31) && d > 0 && ld >= d && nlist > 0, "vdb_ivf_fill: bad shape");
  ivf_fill_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, d, ld, assign, blk_off, nlist, cursor, list_vecs, list_ids);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_ivf_scan_topk_ex(int metric, const float* list_vecs, const int32_t* list_ids, const int32_t* blk_off, int nlist,
                         int d, const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq, int k,
                         int flags, float pad_value, int64_t id_offset, float* out_d, int64_t* out_i,
                         int64_t* scanned_rows, int64_t rows_per_query_hint, void* stream) {
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_ivf_scan_topk: bad metric %d", metric);
  VDB_REQUIRE(nq > 0 && d > 0 && nlist > 0 && nprobe >= 1 && ld_q >= d, "vdb_ivf_scan_topk: bad shape");
  VDB_REQUIRE((reinterpret_cast<uintptr_t>(list_vecs) & 15) == 0, "vdb_ivf_scan_topk: list_vecs must be 16-byte aligned");
  const int kp = k <= 32 ? 32 : k <= 128 ? 128 : k <= 256 ? 256 : k <= 512 ? 512 : 0;
  VDB_REQUIRE(k >= 1 && kp != 0, "vdb_ivf_scan_topk: k=%d unsupported (1..512)", k);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tw = kp == 512 ? 4 : 8;
  int w = rows_per_query_hint > 0 ? warps_for_rows(rows_per_query_hint, tw) : tw;
  if (const char* e = getenv("VDB_IVF_WPQ")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) w = v < tw ? v : tw; }   // tuning override
  while (w < tw && static_cast<size_t>(tw / w) * (((d + 3) / 4 + 7) & ~7) * 16 > 64 * 1024) w *= 2;   // staged queries must fit shared memory
#define VDB_GO(KP, W, TW)                                                                                              \
  return launch_scan<KP, W, TW>(metric, list_vecs, list_ids, blk_off, nlist, d, probes, nprobe, q, ld_q, nq, k, flags, \
                                pad_value, id_offset, out_d, out_i, scanned_rows, s)
#define VDB_PICK(KP, TW)                        \
  switch (w) {                                  \
    case 1: VDB_GO(KP, 1, TW);                  \
    case 2: VDB_GO(KP, 2, TW);                  \
    case 4: VDB_GO(KP, 4, TW);                  \
    default: VDB_GO(KP, TW, TW);                \
  }
  switch (kp) {
    case 32: VDB_PICK(32, 8)
    case 128: VDB_PICK(128, 8)
    case 256: VDB_PICK(256, 8)
    default: VDB_PICK(512, 4)
  }
#undef VDB_PICK
#undef VDB_GO
}

int vdb_ivf_scan_topk(int metric, const float* list_vecs, const int32_t* list_ids, const int32_t* blk_off, int nlist,
                      int d, const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq, int k,
                      int flags, float pad_value, int64_t id_offset, float* out_d, int64_t* out_i,
                      int64_t* scanned_rows, void* stream) {
  return vdb_ivf_scan_topk_ex(metric, list_vecs, list_ids, blk_off, nlist, d, probes, nprobe, q, ld_q, nq, k, flags, pad_value,
                              id_offset, out_d, out_i, scanned_rows, 0, stream);
}

}  // extern "C"
