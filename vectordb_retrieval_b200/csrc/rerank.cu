// Exact re-scoring of per-query candidate sets + top-k (LSH rerank).
//   reference: FaissSearcher._batch_search_lsh_rerank  src/algorithms/modular.py:483-532
//              LSHSearcher._compute_distances + argsort  src/algorithms/lsh.py:242-283
// HBM/L2-bound random row gather: 8 lanes fetch one candidate row with 128-bit loads (a full
// 128-byte line per step), 4 candidates per warp step; fp32 difference, fp64 accumulation.
#include <cstdlib>

#include "select.cuh"

namespace vdb {

// W warps serve one query, a CTA of TW warps holds TW / W queries (see ivf.cu): with C = 800 candidates one query per
// 8 warps gives a warp 100 candidates - no pool ever reaches its KP-th key, nothing is filtered, and the group's
// merging warp sorts all 800 entries serially (7 x 256-element bitonic merges, ~40 us per query): 0.31 of the HBM rate.
// About 512 candidates per warp keeps the pools filtering and the merge short.
template <int KP, int W, int TW>
__global__ void __launch_bounds__(TW * 32)
rerank_topk_kernel(int metric, const float* __restrict__ base, int64_t n, int dpad, int64_t ld,
                   const int64_t* __restrict__ cand, int64_t nq, int c, const float* __restrict__ qmat, int64_t ld_q,
                   int k, int flags, float pad_value, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  constexpr int CAP = pool_cap(KP);
  constexpr int QPC = TW / W;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint64_t* pools_all = reinterpret_cast<uint64_t*>(smem_dyn);
  int* cnts_all = reinterpret_cast<int*>(pools_all + TW * CAP);
  float* thr_all = reinterpret_cast<float*>(cnts_all + TW);
  float* qs_all = thr_all + TW;                 // TW is a multiple of 4: stays 16-byte aligned
  const int warp_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp_cta / W, warp = warp_cta % W;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * QPC + group;
  if (q >= nq) return;                          // whole groups leave together
  uint64_t* pools = pools_all + group * W * CAP;
  int* cnts = cnts_all + group * W;
  float* thr_s = thr_all + group * W;
  float* qs = qs_all + group * dpad;
  const int bar_id = 1 + group;
  for (int j = warp * 32 + lane; j < dpad; j += W * 32) qs[j] = qmat[q * ld_q + j];
  group_sync<W * 32>(bar_id);
  WarpTopK<KP> sel;
  sel.init(pools + warp * CAP);
  const int sub = lane >> 3, sl = lane & 7;
  const int64_t* cq = cand + q * c;
  // Each 8-lane group scores U candidates per step, all of their row loads issued before the first is
  // consumed: the gather is latency bound (random 4*d-byte rows), so bytes in flight are what counts.
  constexpr int U = 4;
  for (int c0 = warp * 4 * U; c0 < c; c0 += W * 4 * U) {
    int64_t id[U];
    bool valid[U];
    const float* row[U];
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ci = c0 + u * 4 + sub;
      id[u] = ci < c ? cq[ci] : -1;
      valid[u] = id[u] >= 0 && id[u] < n;
      row[u] = base + (valid[u] ? id[u] : 0) * ld;
      acc[u] = 0.0;
    }
    for (int j = sl * 4; j < dpad; j += 32) {
      float4 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = valid[u] ? __ldg(reinterpret_cast<const float4*>(row[u] + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 y = *reinterpret_cast<const float4*>(qs + j);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (metric == VDB_METRIC_L2) {
          // four terms in fp32 (fused multiply-adds), then into the fp64 accumulator: one conversion
          // per 16 bytes instead of four, relative error of the sum stays ~1e-7
          const float d0 = x[u].x - y.x, d1 = x[u].y - y.y, d2 = x[u].z - y.z, d3 = x[u].w - y.w;
          acc[u] += static_cast<double>(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3))));
        } else {
          acc[u] += static_cast<double>(fmaf(x[u].x, y.x, fmaf(x[u].y, y.y, fmaf(x[u].z, y.z, x[u].w * y.w))));
        }
      }
    }
    // every lane of a group ends up with the group's four sums; lane u of the group offers candidate
    // u, so the step costs one push (16 offers) instead of four pushes of 4 offers each
    float my_key = 0.f;
    uint32_t my_id = 0u;
    bool my_valid = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double a = acc[u];
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (sl == u) {
        my_key = metric == VDB_METRIC_L2 ? static_cast<float>(a) : -static_cast<float>(a);
        my_id = static_cast<uint32_t>(id[u]);
        my_valid = valid[u];
      }
    }
    sel.push(my_valid, my_key, my_id, lane);
  }
  cta_write_topk<KP, W>(sel, pools, cnts, thr_s, warp, lane, bar_id, metric, k, flags, pad_value, 0, out_d + q * k, out_i + q * k);
}

template <int KP, int W, int TW>
static int launch_rerank(int metric, const float* base, int64_t n, int dpad, int64_t ld, const int64_t* cand, int64_t nq,
                         int c, const float* q, int64_t ld_q, int k, int flags, float pad_value, float* out_d,
                         int64_t* out_i, cudaStream_t stream) {
  constexpr int QPC = TW / W;
  const size_t smem = static_cast<size_t>(TW) * pool_cap(KP) * 8 + TW * 8 + static_cast<size_t>(QPC) * dpad * 4;
  auto kern = rerank_topk_kernel<KP, W, TW>;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>((nq + QPC - 1) / QPC), TW * 32, smem, stream>>>(metric, base, n, dpad, ld, cand, nq, c, q, ld_q, k,
                                                                             flags, pad_value, out_d, out_i);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vdb

using namespace vdb;

extern "C" int vdb_rerank_topk(int metric, const float* base, int64_t n, int d, int64_t ld, const int64_t* cand,
                               int64_t nq, int c, const float* q, int64_t ld_q, int k, int flags, float pad_value,
                               float* out_d, int64_t* out_i, void* stream) {
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_rerank_topk: bad metric %d", metric);
  VDB_REQUIRE(n > 0 && n < (int64_t(1) << 32) && nq > 0 && c >= 1 && d > 0, "vdb_rerank_topk: bad shape");
  const int dpad = (d + 3) / 4 * 4;
  VDB_REQUIRE(ld % 4 == 0 && ld >= dpad && ld_q % 4 == 0 && ld_q >= dpad,
              "vdb_rerank_topk: rows must be zero-padded to a multiple of 4 floats (d=%d ld=%lld ld_q=%lld)", d,
              (long long)ld, (long long)ld_q);
  VDB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
              "vdb_rerank_topk: base/q must be 16-byte aligned");
  VDB_REQUIRE(dpad <= 16384, "vdb_rerank_topk: d too large");
  const int kp = k <= 32 ? 32 : k <= 128 ? 128 : k <= 256 ? 256 : k <= 512 ? 512 : 0;
  VDB_REQUIRE(k >= 1 && kp != 0, "vdb_rerank_topk: k=%d unsupported (1..512)", k);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tw = kp == 512 ? 4 : 8;
  int w = 1;                                      // about 4 096 candidates per warp (measured on 1.2M x 50: one warp per query
  while (w < tw && w * 4096 < c) w *= 2;          // wins up to C = 3 200 - 0.57 / 1.66 ms at C = 800 / 3 200 -, two at C = 6 400)
  if (const char* e = getenv("VDB_RERANK_WPQ")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) w = v < tw ? v : tw; }   // tuning override
  while (w < tw && static_cast<size_t>(tw / w) * dpad * 4 > 64 * 1024) w *= 2;      // staged queries must fit shared memory
#define VDB_GO(KP, W, TW) \
  return launch_rerank<KP, W, TW>(metric, base, n, dpad, ld, cand, nq, c, q, ld_q, k, flags, pad_value, out_d, out_i, s)
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
