// Row utilities and operand preparation for the flat (exact) scan.
//   row norms / row normalisation   <- _safe_normalize (reference src/algorithms/modular.py:109-111)
//   TF32 hi/lo split of base rows and queries (operands of the 3xTF32 tcgen05 contraction)
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace vdb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, static_cast<unsigned long long>(n), __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// one warp per row; fp64 accumulation so the result is the correctly rounded |x|^2
__global__ void row_norms_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* r = x + row * ld;
  double acc = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = r[j];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = static_cast<float>(acc);
}

// y = x / |x| with |x| = sqrtf(float(sum x^2)); zero rows -> 0 (reference: np.divide(..., where=norms > 0))
__global__ void normalize_rows_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ld,
                                      float* __restrict__ y, int64_t ld_y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* r = x + row * ld;
  double acc = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = r[j];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float nrm = sqrtf(static_cast<float>(acc));
  float* w = y + row * ld_y;
  for (int j = lane; j < d; j += 32) w[j] = nrm > 0.f ? r[j] / nrm : 0.f;
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// hi/lo split into the padded operand layout [n_pad, kpad]; also |x|^2 (L2) / 0 (IP) / +inf (pad rows).
// One warp per output row; rows >= n and columns >= d are written as zero.
__global__ void split_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d, int64_t ld,
                                  int kpad, int metric, float scale, float* __restrict__ hi, float* __restrict__ lo,
                                  float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_pad) return;
  float* h = hi + row * kpad;
  float* l = lo + row * kpad;
  double acc = 0.0;
  const bool live = row < n;
  const float* r = x + (live ? row : 0) * ld;
  for (int j = lane; j < kpad; j += 32) {
    const float v = (live && j < d) ? r[j] * scale : 0.f;   // scale is a power of two: exact
    const float vh = tf32_round(v);
    h[j] = vh;
    l[j] = v - vh;
    acc += static_cast<double>(v) * v;
  }
  if (norms != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norms[row] = !live ? CUDART_INF_F : (metric == VDB_METRIC_L2 ? static_cast<float>(acc) : 0.f);
  }
}

}  // namespace vdb

using namespace vdb;

extern "C" {

const char* vdb_last_error(void) { return vdb::last_error(); }
int vdb_abi_version(void) { return VDB_ABI_VERSION; }
int64_t vdb_launch_count(void) { return static_cast<int64_t>(vdb::launches()); }

int vdb_sm_count(int* out) {
  int dev = 0;
  VDB_CHECK_CUDA(cudaGetDevice(&dev));
  VDB_CHECK_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}

int vdb_row_norms(const float* x, int64_t n, int d, int64_t ld, float* out, void* stream) {
  VDB_REQUIRE(n >= 0 && d > 0 && ld >= d, "vdb_row_norms: bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
  if (n == 0) return 0;
  const int64_t blocks = (n * 32 + 255) / 256;
  row_norms_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, d, ld, out);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_normalize_rows(const float* x, int64_t n, int d, int64_t ld, float* y, int64_t ld_y, void* stream) {
  VDB_REQUIRE(n >= 0 && d > 0 && ld >= d && ld_y >= d, "vdb_normalize_rows: bad shape");
  if (n == 0) return 0;
  const int64_t blocks = (n * 32 + 255) / 256;
  normalize_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, d, ld, y, ld_y);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_flat_kpad(int d) { return (d + 31) / 32 * 32; }
int64_t vdb_flat_npad(int64_t n) { return (n + 255) / 256 * 256; }
int64_t vdb_flat_nqpad(int64_t nq) { return (nq + 255) / 256 * 256; }

int vdb_flat_prepare(const float* x, int64_t n, int d, int64_t ld, int metric, float* hi, float* lo,
                     float* norms, void* stream) {
  VDB_REQUIRE(n > 0 && d > 0 && ld >= d, "vdb_flat_prepare: bad shape n=%lld d=%d", (long long)n, d);
  VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "vdb_flat_prepare: bad metric %d", metric);
  const int64_t n_pad = vdb_flat_npad(n);
  const int64_t blocks = (n_pad * 32 + 255) / 256;
  split_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, n_pad, d, ld, vdb_flat_kpad(d), metric, 1.f, hi, lo, norms);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int vdb_flat_prepare_queries(const float* q, int64_t nq, int d, int64_t ld, float* q_hi, float* q_lo, void* stream) {
  VDB_REQUIRE(nq > 0 && d > 0 && ld >= d, "vdb_flat_prepare_queries: bad shape nq=%lld d=%d", (long long)nq, d);
  const int64_t nq_pad = vdb_flat_nqpad(nq);
  const int64_t blocks = (nq_pad * 32 + 255) / 256;
  split_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q, nq, nq_pad, d, ld, vdb_flat_kpad(d), VDB_METRIC_IP, -2.f, q_hi, q_lo, nullptr);
  count_launches(1);
  VDB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
