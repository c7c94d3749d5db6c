// Candidate generation of the Python LSH on the device: the union of a query's T hash buckets ordered like
// `Counter.most_common()` - votes descending, first seen first among equal votes - and cut at the candidate budget.
//   reference: LSHSearcher._gather_candidates / _select_candidates   src/algorithms/lsh.py:219-240
// The reference walks the buckets query by query in Python (a Counter over every bucket entry); on 1.2M x 50 rows with
// 12 tables a query unions 20-50 thousand entries and that loop, not the rerank kernel, bounds the pipeline at a few
// hundred queries per second.  Hashing stays on the host in NumPy (a GPU projection would round differently and move rows
// across bucket borders; the published recall is reproduced bit for bit), the host also resolves each (query, table) key
// to its bucket's (offset, length) in the table arrays.  Everything from there on runs here:
//   expand   every bucket entry -> key (query, id), value = its position in the query's concatenated bucket walk
//   sort #1  by (query, id), stable: a run = one candidate, its length = votes, its first value = first-seen position
//   heads    run heads -> key (query, T - votes, first seen), value = id; other entries get an all-ones key
//   sort #2  by that key: every query's candidates in most_common() order, contiguous
//   gather   the first `cap` of every query -> cand [nq, cap] int64 (-1 = padding), ready for vdb_rerank_topk
// The two sorts are cub::DeviceRadixSort over the significant key bits only (library code, like a cuBLAS call);
// integer work, bound by the sort's memory traffic: 2 x ~5 passes over 12 bytes per bucket entry.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace vdb {

__global__ void __launch_bounds__(128)
lsh_expand_kernel(const int32_t* __restrict__ tbl_ids, const int64_t* __restrict__ seg_off, const int32_t* __restrict__ seg_len,
                  const int64_t* __restrict__ elem_off, int T, int id_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t q = blockIdx.x;
  const int t = blockIdx.y;
  const int len = seg_len[q * T + t];
  if (len <= 0) return;
  int before = 0;                                   // entries of the query's earlier tables: the walk order of the reference
  for (int u = 0; u < t; ++u) before += max(seg_len[q * T + u], 0);
  const int32_t* src = tbl_ids + seg_off[q * T + t];
  const int64_t dst0 = elem_off[q] + before;
  const uint64_t qk = static_cast<uint64_t>(q) << id_bits;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    keys[dst0 + i] = qk | static_cast<uint32_t>(src[i]);
    vals[dst0 + i] = static_cast<uint32_t>(before + i);
  }
}

__global__ void lsh_heads_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ first_seen, int64_t m, int T,
                                 int id_bits, int first_bits, uint64_t* __restrict__ keys2, int32_t* __restrict__ ids2,
                                 unsigned long long* __restrict__ uniq_cnt /* [nq + 1], pre-zeroed */) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t k = keys[i];
  const bool head = i == 0 || keys[i - 1] != k;
  if (!head) {
    keys2[i] = ~0ull;                               // sorts behind every candidate
    ids2[i] = -1;
    return;
  }
  int votes = 1;                                    // an id sits at most once in a bucket: a run is at most T long
  while (votes < T && i + votes < m && keys[i + votes] == k) ++votes;
  const uint64_t q = k >> id_bits;
  const uint32_t id = static_cast<uint32_t>(k & ((1ull << id_bits) - 1ull));
  keys2[i] = (((q << 6) | static_cast<uint64_t>(T - votes)) << first_bits) | first_seen[i];
  ids2[i] = static_cast<int32_t>(id);
  atomicAdd(uniq_cnt + q, 1ull);
}

__global__ void lsh_gather_kernel(const int32_t* __restrict__ ids_sorted, const unsigned long long* __restrict__ uniq_start,
                                  int64_t nq, int cap, int64_t* __restrict__ cand, int32_t* __restrict__ cand_cnt) {
  const int64_t q = blockIdx.x;
  if (q >= nq) return;
  const unsigned long long s0 = uniq_start[q], s1 = uniq_start[q + 1];
  const int have = static_cast<int>(min(static_cast<unsigned long long>(cap), s1 - s0));
  for (int j = threadIdx.x; j < cap; j += blockDim.x)
    cand[q * cap + j] = j < have ? static_cast<int64_t>(ids_sorted[s0 + j]) : -1;
  if (threadIdx.x == 0 && cand_cnt != nullptr) cand_cnt[q] = have;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct CandLayout {
  size_t keys_a, keys_b, vals_a, vals_b, uniq, cub, total;
};

static int cand_layout(int64_t m, int64_t nq, CandLayout* L) {
  size_t sort_bytes = 0, scan_bytes = 0;
  VDB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                                 static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), m, 0, 64));
  VDB_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<const unsigned long long*>(nullptr),
                                               static_cast<unsigned long long*>(nullptr), static_cast<int>(nq + 1)));
  size_t off = 0;
  L->keys_a = off; off += align256(static_cast<size_t>(m) * 8);
  L->keys_b = off; off += align256(static_cast<size_t>(m) * 8);
  L->vals_a = off; off += align256(static_cast<size_t>(m) * 4);
  L->vals_b = off; off += align256(static_cast<size_t>(m) * 4);
  L->uniq = off;   off += 2 * align256(static_cast<size_t>(nq + 1) * 8);
  L->cub = off;    off += align256(std::max(sort_bytes, scan_bytes));
  L->total = off + 256;
  return 0;
}

static int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

}  // namespace vdb

using namespace vdb;

extern "C" {

size_t vdb_lsh_candidates_workspace_bytes(int64_t m_total, int64_t nq) {
  if (m_total <= 0 || nq <= 0) return 0;
  CandLayout L{};
  if (cand_layout(m_total, nq, &L)) return 0;
  return L.total;
}

int vdb_lsh_candidates(const int32_t* tbl_ids, int64_t n_rows, const int64_t* seg_off, const int32_t* seg_len,
                       const int64_t* elem_off, int64_t nq, int n_tables, int64_t m_total, int64_t max_per_query, int cap,
                       int64_t* cand, int32_t* cand_cnt, void* workspace, size_t workspace_bytes, void* stream) {
  VDB_REQUIRE(nq > 0 && n_tables >= 1 && n_tables < 64 && cap >= 1 && n_rows > 0 && n_rows < (int64_t(1) << 31),
              "vdb_lsh_candidates: bad shape (1 <= tables < 64)");
  VDB_REQUIRE(m_total >= 0 && m_total < (int64_t(1) << 31) && max_per_query >= 0 && max_per_query < (int64_t(1) << 32),
              "vdb_lsh_candidates: too many bucket entries in one call (%lld; split the query batch)", (long long)m_total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m_total == 0) {                                 // no bucket hit at all
    VDB_CHECK_CUDA(cudaMemsetAsync(cand, 0xff, static_cast<size_t>(nq) * cap * 8, s));
    if (cand_cnt != nullptr) VDB_CHECK_CUDA(cudaMemsetAsync(cand_cnt, 0, static_cast<size_t>(nq) * 4, s));
    return 0;
  }
  CandLayout L{};
  if (cand_layout(m_total, nq, &L)) return 1;
  VDB_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, "vdb_lsh_candidates: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  uint64_t* keys_a = reinterpret_cast<uint64_t*>(w + L.keys_a);
  uint64_t* keys_b = reinterpret_cast<uint64_t*>(w + L.keys_b);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(w + L.vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(w + L.vals_b);
  unsigned long long* uniq_cnt = reinterpret_cast<unsigned long long*>(w + L.uniq);
  unsigned long long* uniq_start = uniq_cnt + (align256(static_cast<size_t>(nq + 1) * 8) / 8);
  size_t cub_bytes = workspace_bytes - L.cub;
  const int id_bits = bits_for(static_cast<uint64_t>(n_rows - 1));
  const int q_bits = bits_for(static_cast<uint64_t>(nq - 1));
  const int first_bits = bits_for(static_cast<uint64_t>(std::max<int64_t>(max_per_query, 1) - 1));
  VDB_REQUIRE(id_bits + q_bits <= 64 && first_bits + 6 + q_bits <= 63, "vdb_lsh_candidates: key does not fit 64 bits");

  dim3 grid(static_cast<unsigned>(nq), static_cast<unsigned>(n_tables));
  lsh_expand_kernel<<<grid, 128, 0, s>>>(tbl_ids, seg_off, seg_len, elem_off, n_tables, id_bits, keys_a, vals_a);
  VDB_CHECK_CUDA(cudaGetLastError());
  VDB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w + L.cub, cub_bytes, keys_a, keys_b, vals_a, vals_b, m_total, 0, id_bits + q_bits, s));
  VDB_CHECK_CUDA(cudaMemsetAsync(uniq_cnt, 0, static_cast<size_t>(nq + 1) * 8, s));
  lsh_heads_kernel<<<static_cast<unsigned>((m_total + 255) / 256), 256, 0, s>>>(
      keys_b, vals_b, m_total, n_tables, id_bits, first_bits, keys_a, reinterpret_cast<int32_t*>(vals_a), uniq_cnt);
  VDB_CHECK_CUDA(cudaGetLastError());
  cub_bytes = workspace_bytes - L.cub;
  VDB_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w + L.cub, cub_bytes, keys_a, keys_b, vals_a, vals_b, m_total, 0,
                                                 std::min(64, first_bits + 6 + q_bits + 1), s));
  cub_bytes = workspace_bytes - L.cub;
  VDB_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w + L.cub, cub_bytes, uniq_cnt, uniq_start, static_cast<int>(nq + 1), s));
  lsh_gather_kernel<<<static_cast<unsigned>(nq), 128, 0, s>>>(reinterpret_cast<const int32_t*>(vals_b), uniq_start, nq, cap, cand, cand_cnt);
  VDB_CHECK_CUDA(cudaGetLastError());
  count_launches(3);
  return 0;
}

}  // extern "C"
