// IVF + 8-bit scalar quantiser ("IVF<n>,SQ8"): per-dimension range training, residual encoding, list scan + top-k.
//   reference: FaissFactoryIndexer(index_key="IVF256,SQ8") + FaissSearcher
//   src/algorithms/modular.py:224-286,536-548; configs/benchmark_config.yaml:51-60
// FAISS semantics restated [FAISS-upstream, parity unpinned]: IndexIVFScalarQuantizer with QT_8bit and by_residual:
//   train   per dimension vmin = min, vdiff = max - min over the training residuals x - centroid(x)
//   encode  code_j = min(255, int(255 * clamp((r_j - vmin_j) / vdiff_j, 0, 1)))           (truncation)
//   decode  r^_j = vmin_j + vdiff_j * (code_j + 0.5) / 255
//   search  top-nprobe lists by the coarse quantiser; L2: |q - c - r^|^2, inner product: q.c + q.r^
// List layout "interleaved-32 bytes": block = 32 vectors as uint4 [d16][32 lanes], d16 = ceil(d / 16): a warp reads one
// 512-byte line per 16 dimensions, one vector per lane.  One byte per dimension instead of four: the scan is bound by the
// byte -> float conversion and the FMA issue rate about as much as by HBM (about 4 instructions per byte).
#include <cstdlib>

#include "select.cuh"

namespace vdb {

// column-wise min / max over rows: one block of 256 threads per 256-row slab, ordered-uint atomics on the result
__global__ void sq8_minmax_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, uint32_t* __restrict__ omin,
                                  uint32_t* __restrict__ omax) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 256, r1 = min(n, r0 + 256);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float lo = CUDART_INF_F, hi = -CUDART_INF_F;
    for (int64_t r = r0; r < r1; ++r) {
      const float v = x[r * ld + j];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    atomicMin(omin + j, f2ord(lo));
    atomicMax(omax + j, f2ord(hi));
  }
}

__global__ void sq8_range_kernel(const uint32_t* __restrict__ omin, const uint32_t* __restrict__ omax, int d, float* __restrict__ vmin,
                                 float* __restrict__ vdiff) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const float lo = ord2f(omin[j]), hi = ord2f(omax[j]);
  vmin[j] = lo;
  vdiff[j] = hi - lo;
}

// out[i, :] = x[i, :] - centroids[assign[i], :]   (assign < 0: plain copy)
__global__ void sq8_residual_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, const float* __restrict__ cent,
                                    const int32_t* __restrict__ assign, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int l = assign[row];
  for (int j = lane; j < d; j += 32) out[row * d + j] = x[row * ld + j] - (l >= 0 ? cent[static_cast<int64_t>(l) * d + j] : 0.f);
}

__device__ __forceinline__ uint32_t sq8_code(float r, float vmin, float vdiff) {
  float xi = vdiff != 0.f ? (r - vmin) / vdiff : 0.f;
  xi = fminf(fmaxf(xi, 0.f), 1.f);
  return min(255u, static_cast<uint32_t>(static_cast<int>(255.f * xi)));
}

// scatter rows into the interleaved byte layout, encoding the residual against the row's centroid; one warp per row
__global__ void sq8_fill_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, const float* __restrict__ cent,
                                const int32_t* __restrict__ assign, const int32_t* __restrict__ blk_off, int nlist,
                                int32_t* cursor, const float* __restrict__ vmin, const float* __restrict__ vdiff,
                                uint8_t* __restrict__ codes, int32_t* __restrict__ ids) {
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
  const int d16 = (d + 15) / 16;
  for (int c = lane; c < d16 * 16; c += 32) {
    uint32_t code = 0;
    if (c < d) code = sq8_code(x[row * ld + c] - cent[static_cast<int64_t>(l) * d + c], vmin[c], vdiff[c]);
    codes[((b * d16 + (c >> 4)) * 32 + v) * 16 + (c & 15)] = static_cast<uint8_t>(code);
  }
  if (lane == 0) ids[b * 32 + v] = static_cast<int32_t>(row);
}

// Byte B of w as a float, exactly, without the conversion unit: PRMT drops the byte into the mantissa of 2^23 (ALU pipe), one
// FADD removes the 2^23.  I2F runs on the quarter-rate XU pipe, which the first version of the scan kept 69 % busy
// (profiles/r2/sq8_scan_np32_r3p.md) - its top pipe, ahead of FMA and LSU.
template <int B>
__device__ __forceinline__ float byte_as_float(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | B)) - 8388608.f;
}

// W warps per query, TW / W queries per CTA (see ivf.cu).  Per probed list the group stages, in shared memory,
//   L2: t_j = (q_j - c_j) - (vmin_j + 0.5 vdiff_j / 255)  so that  dist = sum_j (t_j - s_j code_j)^2,  s_j = vdiff_j / 255
//   IP: w_j = q_j s_j (once per query) and the scalar  q.c + sum_j q_j (vmin_j + 0.5 s_j)  so that  score = scalar + sum_j w_j code_j
template <int KP, int W, int TW>
__global__ void __launch_bounds__(TW * 32, 1024 / (TW * 32))
ivf_sq8_scan_kernel(int metric, const uint4* __restrict__ codes, const int32_t* __restrict__ ids, const int32_t* __restrict__ blk_off,
                    int nlist, int d, const float* __restrict__ cent, const float* __restrict__ vmin, const float* __restrict__ vdiff,
                    const int64_t* __restrict__ probes, int nprobe, const float* __restrict__ qmat, int64_t ld_q, int64_t nq, int k,
                    int flags, float pad_value, int64_t id_offset, float* __restrict__ out_d, int64_t* __restrict__ out_i) {
  constexpr int CAP = pool_cap(KP);
  constexpr int QPC = TW / W;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint64_t* pools_all = reinterpret_cast<uint64_t*>(smem_dyn);
  int* cnts_all = reinterpret_cast<int*>(pools_all + TW * CAP);
  float* thr_all = reinterpret_cast<float*>(cnts_all + TW);
  const int d16 = (d + 15) / 16, dp = d16 * 16;
  float* vec_all = thr_all + TW;                               // [QPC][2][dp]: per-dimension scale s (or w), and t
  float* red_all = vec_all + QPC * 2 * dp;                     // [TW] partial sums of the IP scalar
  const int warp_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp_cta / W, warp = warp_cta % W;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * QPC + group;
  if (q >= nq) return;                                         // whole groups leave together
  uint64_t* pools = pools_all + group * W * CAP;
  int* cnts = cnts_all + group * W;
  float* thr_s = thr_all + group * W;
  float* sv = vec_all + group * 2 * dp;                        // s_j (L2) or w_j (IP)
  float* tv = sv + dp;                                         // t_j (L2 only)
  float* red = red_all + group * W;
  const int bar_id = 1 + group;
  const int gt = warp * 32 + lane;                             // thread index inside the group
  const bool l2 = metric == VDB_METRIC_L2;
  for (int j = gt; j < dp; j += W * 32) {
    const float s = j < d ? vdiff[j] * (1.f / 255.f) : 0.f;
    sv[j] = l2 ? s : (j < d ? qmat[q * ld_q + j] * s : 0.f);
    tv[j] = 0.f;
  }
  group_sync<W * 32>(bar_id);
  WarpTopK<KP> sel;
  sel.init(pools + warp * CAP);
  for (int pi = 0; pi < nprobe; ++pi) {
    const int64_t l = probes[q * nprobe + pi];
    if (l < 0 || l >= nlist) continue;                         // group-uniform
    float base = 0.f;                                          // IP: q.c + sum q_j (vmin_j + 0.5 s_j)
    if (l2) {
      for (int j = gt; j < d; j += W * 32)
        tv[j] = (qmat[q * ld_q + j] - cent[l * d + j]) - (vmin[j] + 0.5f * vdiff[j] * (1.f / 255.f));
    } else {
      float part = 0.f;
      for (int j = gt; j < d; j += W * 32)
        part = fmaf(qmat[q * ld_q + j], cent[l * d + j] + vmin[j] + 0.5f * vdiff[j] * (1.f / 255.f), part);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) red[warp] = part;
    }
    group_sync<W * 32>(bar_id);
    if (!l2) {
#pragma unroll
      for (int w = 0; w < W; ++w) base += red[w];
    }
    const int b0 = blk_off[l], b1 = blk_off[l + 1];
    for (int b = b0 + warp; b < b1; b += W) {
      const int id = ids[static_cast<int64_t>(b) * 32 + lane];
      const uint4* p = codes + static_cast<int64_t>(b) * d16 * 32 + lane;
      float acc = 0.f;
      for (int c0 = 0; c0 < d16; c0 += 4) {                    // four 128-bit loads (64 dimensions) in flight per lane
        uint4 u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = c0 + i < d16 ? __ldg(p + (c0 + i) * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + i >= d16) break;
          const uint32_t wds[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int j = (c0 + i) * 16 + h * 4;
            const float4 s4 = *reinterpret_cast<const float4*>(sv + j);
            const float c0f = byte_as_float<0>(wds[h]), c1f = byte_as_float<1>(wds[h]);
            const float c2f = byte_as_float<2>(wds[h]), c3f = byte_as_float<3>(wds[h]);
            if (l2) {
              const float4 t4 = *reinterpret_cast<const float4*>(tv + j);
              const float e0 = fmaf(-s4.x, c0f, t4.x), e1 = fmaf(-s4.y, c1f, t4.y), e2 = fmaf(-s4.z, c2f, t4.z), e3 = fmaf(-s4.w, c3f, t4.w);
              acc = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, acc))));
            } else {
              acc = fmaf(s4.x, c0f, fmaf(s4.y, c1f, fmaf(s4.z, c2f, fmaf(s4.w, c3f, acc))));
            }
          }
        }
      }
      const float key = l2 ? acc : -(base + acc);
      sel.push(id >= 0, key, static_cast<uint32_t>(id), lane);
    }
    group_sync<W * 32>(bar_id);                                // every warp is done with t / red before the next list overwrites them
  }
  cta_write_topk<KP, W>(sel, pools, cnts, thr_s, warp, lane, bar_id, metric, k, flags, pad_value, id_offset, out_d + q * k,
                        out_i + q * k);
}

template <int KP, int W, int TW>
static int launch_sq8_scan(int metric, const uint8_t* codes, const int32_t* ids, const int32_t* blk_off, int nlist, int d,
                           const float* cent, const float* vmin, const float* vdiff, const int64_t* probes, int nprobe,
                           const float* q, int64_t ld_q, int64_t nq, int k, int flags, float pad_value, int64_t id_offset,
                           float* out_d, int64_t* out_i, cudaStream_t stream) {
  constexpr int QPC = TW / W;
  const int dp = (d + 15) / 16 * 16;
  const size_t smem = static_cast<size_t>(TW) * pool_cap(KP) * 8 + TW * 8 + static_cast<size_t>(QPC) * 2 * dp * 4 + TW * 4;
  auto kern = ivf_sq8_scan_kernel<KP, W, TW>;
  if (smem > 48 * 1024) VDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<static_cast<unsigned>((nq + QPC - 1) / QPC), TW * 32, smem, stream>>>(
      metric, reinterpret_cast<const uint4*>(codes), ids, blk_off, nlist, d, cent, vmin, vdiff, probes, nprobe, q, ld_q, nq, k, flags,
      pad_value, id_offset, out_d, out_i);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vdb

using namespace vdb;

extern "C" {

int vdb_sq8_d16(int d) { return (d + 15) / 16; }

int vdb_sq8_residuals(const float* x, int64_t n, int d, int64_t ld, const float* centroids, const int32_t* assign, float* out,
                      void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && ld >= d, "vdb_sq8_residuals: bad shape");
  sq8_residual_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, d, ld, centroids,
                                                                                                             assign, out);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_sq8_train(const float* x, int64_t n, int d, int64_t ld, float* vmin, float* vdiff, void* scratch_2d_u32, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && ld >= d && scratch_2d_u32 != nullptr, "vdb_sq8_train: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* omin = static_cast<uint32_t*>(scratch_2d_u32);
  uint32_t* omax = omin + d;
  VDB_CHECK_CUDA(cudaMemsetAsync(omin, 0xff, static_cast<size_t>(d) * 4, s));
  VDB_CHECK_CUDA(cudaMemsetAsync(omax, 0x00, static_cast<size_t>(d) * 4, s));
  sq8_minmax_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x, n, d, ld, omin, omax);
  sq8_range_kernel<<<(d + 127) / 128, 128, 0, s>>>(omin, omax, d, vmin, vdiff);
  count_launches(2);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_sq8_fill(const float* x, int64_t n, int d, int64_t ld, const float* centroids, const int32_t* assign,
                 const int32_t* blk_off, int nlist, int32_t* cursor, const float* vmin, const float* vdiff, uint8_t* list_codes,
                 int32_t* list_ids, void* stream) {
  VDB_REQUIRE(n > 0 && n < (int64_t(1) << 31) && d > 0 && ld >= d && nlist > 0, "vdb_sq8_fill: bad shape");
  sq8_fill_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, d, ld, centroids, assign, blk_off, nlist, cursor, vmin, vdiff, list_codes, list_ids);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_ivf_sq8_scan_topk(int metric, const uint8_t* list_codes, const int32_t* list_ids, const int32_t* blk_off, int nlist, int d,
                          const float* centroids, const float* vmin, const float* vdiff, const int64_t* probes, int nprobe,
                          const float* q, int64_t ld_q, int64_t nq, int k, int flags, float pad_value, int64_t id_offset,
                          float* out_d, int64_t* out_i, int64_t rows_per_query_hint, void* stream) {
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_ivf_sq8_scan_topk: bad metric %d", metric);
  VDB_REQUIRE(nq > 0 && d > 0 && d <= 4096 && nlist > 0 && nprobe >= 1 && ld_q >= d, "vdb_ivf_sq8_scan_topk: bad shape (d <= 4096)");
  VDB_REQUIRE((reinterpret_cast<uintptr_t>(list_codes) & 15) == 0, "vdb_ivf_sq8_scan_topk: list_codes must be 16-byte aligned");
  const int kp = k <= 32 ? 32 : k <= 128 ? 128 : k <= 256 ? 256 : k <= 512 ? 512 : 0;
  VDB_REQUIRE(k >= 1 && kp != 0, "vdb_ivf_sq8_scan_topk: k=%d unsupported (1..512)", k);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tw = kp == 512 ? 4 : 8;
  int w = 1;                                                   // about 4 096 rows per warp, like the fp32 list scan
  while (w < tw && static_cast<int64_t>(w) * 4096 < rows_per_query_hint) w *= 2;
  if (rows_per_query_hint <= 0) w = tw;
  while (w < tw && static_cast<size_t>(tw / w) * 2 * ((d + 15) / 16 * 16) * 4 > 64 * 1024) w *= 2;
#define VDB_GO(KP, W, TW)                                                                                                       \
  return launch_sq8_scan<KP, W, TW>(metric, list_codes, list_ids, blk_off, nlist, d, centroids, vmin, vdiff, probes, nprobe, q, ld_q, \
                                    nq, k, flags, pad_value, id_offset, out_d, out_i, s)
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

}  // extern "C"
