/* libvdbcuda - C ABI of the B200 scan + top-k hot path.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point replaces the arithmetic behind one reference call site
 * (Human-Augment-Analytics/vectordb-retrieval, paths relative to the reference root):
 *
 *   vdb_flat_*       faiss.IndexFlat.add/search   src/algorithms/exact_search.py:38-39,58,78
 *                    LinearSearcher.batch_search   src/algorithms/modular.py:336-387
 *                    ground-truth flat search      src/benchmark/dataset.py:915-950
 *   vdb_normalize_rows / vdb_row_norms
 *                    _safe_normalize               src/algorithms/modular.py:109-111
 *                    _normalize_rows               src/algorithms/lsh.py:13-16
 *   vdb_ivf_*        index_factory("IVFn,Flat") train/add/search
 *                                                  src/algorithms/modular.py:277-286,536-548
 *                                                  src/algorithms/approximate_search.py:39-51,87
 *   vdb_rerank_topk  FaissSearcher._batch_search_lsh_rerank  src/algorithms/modular.py:483-532
 *                    LSHSearcher._compute_distances + argsort src/algorithms/lsh.py:242-283
 *   vdb_lsh_encode / vdb_hamming_topk
 *                    faiss.IndexLSH add/search     src/algorithms/modular.py:215-216,477
 *   vdb_merge_topk   (new) merge of per-GPU top-k lists after the NCCL allgather
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says host; row-major, `ld` in elements
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and allocates
 *     nothing: the caller provides workspaces sized by the matching *_bytes() query
 *   - return value: 0 = ok, non-zero = error; text via vdb_last_error() (thread-local)
 *   - no C++ exception crosses this boundary
 *   - distances out: float32; ids out: int64 (row index + id_offset), -1 = padding
 *   - tuning overrides read from the environment at call time (launch shape only, never results):
 *     VDB_IVF_WPQ / VDB_RERANK_WPQ = 1 | 2 | 4 | 8 warps per query for the IVF list scan / the rerank kernel
 */
#ifndef VDB_CUDA_H_
#define VDB_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDB_ABI_VERSION 1

/* metric of the scoring kernels (cosine = IP on rows normalised by vdb_normalize_rows) */
enum { VDB_METRIC_L2 = 0, VDB_METRIC_IP = 1 };

/* output convention flags (the reference's classes disagree on these) */
enum {
  VDB_OUT_SQRT   = 1, /* L2: report sqrt(d2) (LinearSearcher, LSH rerank) instead of d2 (FAISS) */
  VDB_OUT_NEGATE = 2, /* IP: report -score (LinearSearcher/FaissSearcher) instead of +score (raw FAISS) */
  VDB_OUT_ONE_MINUS = 4 /* IP: report 1 - score (LSHSearcher cosine distance, lsh.py:246) */
};

/* which scan kernel vdb_flat_topk uses */
enum { VDB_IMPL_AUTO = 0, VDB_IMPL_TCGEN05 = 1, VDB_IMPL_TCGEN05_1CTA = 2, VDB_IMPL_SIMT = 3 };

const char* vdb_last_error(void);
int vdb_abi_version(void);
/* kernels launched by this library since it was loaded (bench.py reports the delta) */
int64_t vdb_launch_count(void);
/* number of SMs of the current device (grid sizing / reporting) */
int vdb_sm_count(int* out);

/* ---- row utilities ------------------------------------------------------------------ */
/* out[i] = sum_j x[i,j]^2 (fp64 accumulate, fp32 result) */
int vdb_row_norms(const float* x, int64_t n, int d, int64_t ld, float* out, void* stream);
/* y[i,:] = x[i,:] / |x[i,:]|, zero rows stay zero; y may alias x */
int vdb_normalize_rows(const float* x, int64_t n, int d, int64_t ld, float* y, int64_t ld_y, void* stream);

/* ---- flat (exact) index ------------------------------------------------------------- */
/* Device layout of a prepared flat shard (all owned by the caller):
 *   hi, lo : [n_pad, kpad] fp32, kpad = round_up(d, 32), n_pad = round_up(n, 256);
 *            hi = x rounded to TF32, lo = x - hi (exact), so hi + lo == x bit-exactly;
 *            padding rows / columns are zero.  These are the TMA / tcgen05 operands.
 *   norms  : [n_pad] fp32, |x|^2 for L2, 0 for IP, +inf for padding rows.            */
int     vdb_flat_kpad(int d);
int64_t vdb_flat_npad(int64_t n);
int vdb_flat_prepare(const float* x, int64_t n, int d, int64_t ld, int metric,
                     float* hi, float* lo, float* norms, void* stream);

/* Split queries the same way, pre-scaled by -2 (exact): q_hi + q_lo == -2q; [nq_pad, kpad],
 * nq_pad = round_up(nq, 256).  The contraction then yields -2 q.x and the epilogue only adds
 * |x|^2 to obtain the ranking key. */
int64_t vdb_flat_nqpad(int64_t nq);
int vdb_flat_prepare_queries(const float* q, int64_t nq, int d, int64_t ld,
                             float* q_hi, float* q_lo, void* stream);

/* Workspace for one vdb_flat_topk call (pools of threshold-passing candidates, counters,
 * per-query shared thresholds).  Depends on nq, k and the SM count only - never on n. */
size_t vdb_flat_topk_workspace_bytes(int64_t nq, int k);

/* Fused scoring + selection + exact re-scoring of the winners.
 *   keys   n_j - 2 q.x_j (L2) / -2 q.x_j (IP) from a 3xTF32 tcgen05 contraction (or the SIMT
 *          kernel), filtered against a per-query running k'-th bound inside the epilogue;
 *          the nq x n matrix is never written.
 *   finish the k' >= k + 8 survivors are re-scored exactly (fp64 accumulate over hi+lo, the
 *          difference form for L2), sorted by (distance, id) and the best k written.
 *   out_d [nq,k] float32, out_i [nq,k] int64 = row + id_offset; rows beyond n: pad_value / -1.
 *   flags  VDB_OUT_*; impl VDB_IMPL_*.  1 <= k <= 504.                                 */
int vdb_flat_topk(int metric, const float* hi, const float* lo, const float* norms,
                  int64_t n, int d, int64_t id_offset,
                  const float* q_hi, const float* q_lo, int64_t nq,
                  int k, int flags, float pad_value, int impl,
                  float* out_d, int64_t* out_i,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Size of the persistent tcgen05 grid (clusters of 1 or 2 CTAs) the scan launches for dimension d on the current
 * device: min(SMs / cluster size, clusters the occupancy calculator reports as co-resident).  The pool hand-over
 * between work items needs every cluster of the grid resident at once, so the grid never exceeds this. */
int vdb_flat_grid_clusters(int impl, int d, int* out);

/* Debug / test hook: dense key matrix of the tcgen05 contraction for a small problem,
 * keys [nq_pad, n_pad] fp32 (n_pad, nq_pad as above).  impl: VDB_IMPL_TCGEN05 / _1CTA / SIMT. */
int vdb_flat_dense_keys(const float* hi, const float* lo, const float* norms, int64_t n, int d,
                        const float* q_hi, const float* q_lo, int64_t nq, int impl,
                        float* keys, void* stream);

/* The knobs below (debug mode, seeding override, timing log) are PER HOST THREAD: set by one thread, they affect
 * only the calls that thread makes.
 * Bring-up knob for kernel timing experiments (results are WRONG for mode != 0): 2 = filter
 * but never append a candidate, 3 = do not read the accumulators at all (contraction pipeline
 * only), 5 = start from the bounds the previous call left in the workspace (perfect warm start).
 * Returns the previous mode. */
int vdb_set_debug_mode(int mode);
/* mode 8 (and 9 = 8 with kept bounds): the tcgen05 scan also accumulates epilogue counters; this
 * reads and clears them: [0] epilogue-warp cycles, [1] of which waiting for an accumulator, [2] in
 * the append path, [3] appended candidates, [4] chunks with a hit in first-generation items and
 * [5] their cycles, [6] epilogue warps, [7] chunks with a hit. */
int vdb_debug_read_prof(uint64_t* out8);

/* Seeded bounds of the tcgen05 scan (see flat.cu): a pre-pass over `sample_tiles` strided base
 * tiles (256 rows each; 0 disables seeding) gives every query the `rank`-th smallest sampled key
 * as its starting bound; queries whose guess turns out too tight are detected after the main pass
 * and re-scanned from an infinite bound, so results never depend on these knobs.  Defaults 64, 0 (rank chosen per shard: 16, or 32 on
 * small shards where the margin cap below binds; 1 <= rank <= 32 pins it).
 * Debug mode 6 forces every query through the re-scan (test hook), mode 7 disables seeding. */
int vdb_flat_set_seeding(int sample_tiles, int rank);
/* The sample is capped so that the guess's expected rank in the shard, rank * N / S rows, stays >= margin * k'
 * (k' = kept candidates, 128 for k = 100): a smaller margin allows a larger sample and a tighter guess on small
 * shards at a higher (still verified and repaired) chance of a redo.  0 = default: 3 at rank 32, else max(4, 64 / rank) - one failed guess costs a whole wave of the redo launch, so the
 * failure probability per query, P(Poisson(rank / margin) >= rank), has to stay far below 1 / nq. */
int vdb_flat_set_seeding_margin(int margin);
/* Number of queries re-scanned since the last call (reads and clears a device counter). */
int vdb_debug_redo_queries(uint64_t* out);

/* Measurement hook for bench.py's roofline leg: while enabled, every vdb_flat_topk call brackets
 * its scan kernel with a pair of CUDA events on the launching stream (up to 512 calls).
 * vdb_flat_timing_read waits for the recorded events, writes the scan durations in milliseconds
 * in call order, stores their number in *n_out and clears the log. */
int vdb_flat_timing_enable(int on);
int vdb_flat_timing_read(float* ms, int max_records, int* n_out);

/* ---- multi-GPU merge ----------------------------------------------------------------- */
/* d_all/i_all: [parts, nq, k] as written by an allgather of per-shard results (each sorted
 * best-first by (distance, id), padding id -1).  Parts must arrive in ascending id-range order
 * (row-sharded base, allgather in rank order): then the merge is on (distance, id) and the
 * result does not depend on the number of parts.  `descending` for raw-IP (+score) lists. */
int vdb_merge_topk(const float* d_all, const int64_t* i_all, int parts, int64_t nq, int k,
                   int descending, float pad_value, float* out_d, int64_t* out_i, void* stream);
/* Same merge over parts that are not densely packed: part p's lists start at d_all + p * stride_d and
 * i_all + p * stride_i (strides in elements, >= nq * k).  This is what lets ONE collective carry both
 * arrays (each rank's block = [distances | ids]) and lets a rank merge only a slice of the queries
 * (pass pointers offset to the slice's first query and keep the strides). */
int vdb_merge_topk_strided(const float* d_all, const int64_t* i_all, int64_t stride_d, int64_t stride_i,
                           int parts, int64_t nq, int k, int descending, float pad_value,
                           float* out_d, int64_t* out_i, void* stream);

/* ---- candidate re-ranking (LSH) ------------------------------------------------------- */
/* base [n,d] fp32 row-major (ld elements, ld % 4 == 0, 16-byte aligned), cand [nq,C] int64
 * (-1 = invalid), q [nq,d].  Exact fp32-input / fp64-accumulate scoring, difference form
 * for L2.  Rows with fewer than k valid candidates are padded (pad_value, -1). */
int vdb_rerank_topk(int metric, const float* base, int64_t n, int d, int64_t ld,
                    const int64_t* cand, int64_t nq, int c, const float* q, int64_t ld_q,
                    int k, int flags, float pad_value, float* out_d, int64_t* out_i, void* stream);

/* ---- LSH sign codes + Hamming top-k (candidate generator in front of the rerank) ------- */
/* Code layout: [n, words] uint32, words = vdb_lsh_code_words(nbits) = 4 * ceil(nbits / 128)
 * (16-byte aligned rows, unused high bits zero); bit b of a row = [ x . P[b] >= 0 ].
 * proj_t: the projection transposed and padded, [d, words * 32] fp32 (column b = P[b]). */
int vdb_lsh_code_words(int nbits);
int vdb_lsh_encode(const float* x, int64_t n, int d, int64_t ld, const float* proj_t, int nbits,
                   uint32_t* codes, void* stream);
/* k smallest Hamming distances per query, ordered by (distance, id): out_d [nq,k] float32
 * (the integer distance), out_i [nq,k] int64 = row + id_offset.  out_* must be pre-filled with
 * the padding the caller wants for k > n.  nbits <= 1024.  Replaces faiss.IndexLSH.search. */
size_t vdb_hamming_topk_workspace_bytes(int64_t nq, int nbits);
int vdb_hamming_topk(const uint32_t* codes, int64_t n, const uint32_t* qcodes, int64_t nq, int nbits,
                     int k, int64_t id_offset, float* out_d, int64_t* out_i,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- Python-LSH candidate generation (bucket union + vote order) ----------------------------------------- */
/* Replaces LSHSearcher._gather_candidates / _select_candidates (src/algorithms/lsh.py:219-240): the union of a
 * query's n_tables hash buckets in Counter.most_common() order (votes descending, first seen first), cut at `cap`.
 *   tbl_ids   [n_tables * n_rows] int32: table t's rows grouped by bucket, insertion (= ascending row) order inside
 *   seg_off   [nq, n_tables] int64: start of the query's bucket in tbl_ids (table offset included); seg_len [nq, n_tables]
 *             int32: its length, 0 = no such bucket.  Both come from the HOST: hashing stays in NumPy so that keys
 *             match the reference bit for bit.
 *   elem_off  [nq + 1] int64: exclusive prefix sum of a query's total bucket entries; m_total = elem_off[nq]
 *             (< 2^31 per call: split large batches), max_per_query = the largest single total
 *   cand      [nq, cap] int64 out, -1 = padding; cand_cnt [nq] int32 out (nullable): candidates found (<= cap)
 * Two cub::DeviceRadixSort passes over 12 bytes per bucket entry; workspace from the _bytes query. */
size_t vdb_lsh_candidates_workspace_bytes(int64_t m_total, int64_t nq);
int vdb_lsh_candidates(const int32_t* tbl_ids, int64_t n_rows, const int64_t* seg_off, const int32_t* seg_len,
                       const int64_t* elem_off, int64_t nq, int n_tables, int64_t m_total, int64_t max_per_query,
                       int cap, int64_t* cand, int32_t* cand_cnt, void* workspace, size_t workspace_bytes, void* stream);

/* ---- IVF-Flat -------------------------------------------------------------------------- */
/* Inverted lists, device layout ("interleaved-32"): list l occupies blocks
 * [blk_off[l], blk_off[l+1]) ; block b holds 32 vectors as float4 [d4][32 lanes]
 * (d4 = ceil(d/4), component c of slot v at ((b*d4 + c/4)*32 + v)*4 + c%4), so a warp reads
 * one 512-byte line per 4 dimensions with 128-bit loads.  list_ids: [n_blocks*32] int32 row
 * index inside the shard, -1 = empty slot (caller pre-fills with -1, list_vecs with 0). */
int vdb_ivf_d4(int d);
/* counts[l] += number of rows assigned to list l (assign in [0,nlist); counts pre-zeroed) */
int vdb_ivf_count(const int32_t* assign, int64_t n, int nlist, int32_t* counts, void* stream);
/* k-means centroid update (IVF training): sums[l,:] += x[i,:] and counts[l] += 1 with
 * l = assign[i] (int64, as written by vdb_flat_topk with k = 1; out-of-range = skipped).
 * sums [nlist,d] fp32 and counts [nlist] int32 are pre-zeroed by the caller. */
int vdb_kmeans_accumulate(const float* x, int64_t n, int d, int64_t ld, const int64_t* assign,
                          int nlist, float* sums, int32_t* counts, void* stream);
/* scatter rows into the interleaved layout; blk_off [nlist+1] int32 (prefix sum of
 * ceil(count/32)), cursor [nlist] int32 zeroed scratch.  Slot order inside a list is
 * arbitrary; search results never depend on it (the scan orders by (distance, id)). */
int vdb_ivf_fill(const float* x, int64_t n, int d, int64_t ld, const int32_t* assign,
                 const int32_t* blk_off, int nlist, int32_t* cursor,
                 float* list_vecs, int32_t* list_ids, void* stream);
/* probes [nq,nprobe] int64 as written by vdb_flat_topk over the centroids (-1 = skip); scans
 * those lists with exact scoring (fp32 difference, fp64 accumulation) and keeps the top k.
 * out_i = row + id_offset.  scanned_rows (nullable, pre-zeroed): += scanned list lengths. */
int vdb_ivf_scan_topk(int metric, const float* list_vecs, const int32_t* list_ids,
                      const int32_t* blk_off, int nlist, int d,
                      const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq,
                      int k, int flags, float pad_value, int64_t id_offset,
                      float* out_d, int64_t* out_i, int64_t* scanned_rows, void* stream);

/* Same scan with a hint for the launch shape: rows_per_query_hint ~ nprobe * (rows / nlist), the rows a query is
 * expected to scan.  The kernel gives a query one warp per ~4 096 expected rows (1, 2, 4 or 8 warps; a 256-thread CTA
 * then serves 8, 4, 2 or 1 queries), so short scans (nprobe 1..8 over ~300-row lists) are not dominated by per-query
 * fixed work.  0 = one query per CTA (what vdb_ivf_scan_topk does).  Results never depend on the hint. */
int vdb_ivf_scan_topk_ex(int metric, const float* list_vecs, const int32_t* list_ids,
                         const int32_t* blk_off, int nlist, int d,
                         const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq,
                         int k, int flags, float pad_value, int64_t id_offset,
                         float* out_d, int64_t* out_i, int64_t* scanned_rows,
                         int64_t rows_per_query_hint, void* stream);

/* ---- IVF + 8-bit scalar quantiser ("IVF<n>,SQ8") ------------------------------------------------------------ */
/* faiss.index_factory(d, "IVF<n>,SQ8") as reached from FaissFactoryIndexer (src/algorithms/modular.py:224-286;
 * configs/benchmark_config.yaml:51-60): IndexIVFScalarQuantizer, QT_8bit, by_residual.  Coarse quantiser, k-means and
 * assignment are the IVF-Flat ones (vdb_flat_topk over the centroids, vdb_kmeans_accumulate, vdb_ivf_count).
 *   vdb_sq8_residuals  out[i,:] = x[i,:] - centroids[assign[i],:]             (training input of the quantiser)
 *   vdb_sq8_train      vmin[j] = min_i r[i,j], vdiff[j] = max_i r[i,j] - vmin[j]; scratch = 2 * d uint32
 *   vdb_sq8_fill       code = min(255, int(255 * clamp((r - vmin) / vdiff, 0, 1))) of the residual, scattered into the
 *                      byte lists: block b holds 32 vectors as uint4 [d16][32 lanes], d16 = ceil(d / 16), byte c of slot
 *                      v at ((b*d16 + c/16)*32 + v)*16 + c%16; blk_off / cursor / list_ids as for vdb_ivf_fill
 *   vdb_ivf_sq8_scan_topk  decode r^ = vmin + vdiff * (code + 0.5) / 255 on the fly; L2: |q - c - r^|^2, inner product
 *                      q.c + q.r^; top k per query, (distance, id) order; conventions and hint as vdb_ivf_scan_topk_ex */
int vdb_sq8_d16(int d);
int vdb_sq8_residuals(const float* x, int64_t n, int d, int64_t ld, const float* centroids, const int32_t* assign,
                      float* out, void* stream);
int vdb_sq8_train(const float* x, int64_t n, int d, int64_t ld, float* vmin, float* vdiff, void* scratch_2d_u32, void* stream);
int vdb_sq8_fill(const float* x, int64_t n, int d, int64_t ld, const float* centroids, const int32_t* assign,
                 const int32_t* blk_off, int nlist, int32_t* cursor, const float* vmin, const float* vdiff,
                 uint8_t* list_codes, int32_t* list_ids, void* stream);
int vdb_ivf_sq8_scan_topk(int metric, const uint8_t* list_codes, const int32_t* list_ids, const int32_t* blk_off,
                          int nlist, int d, const float* centroids, const float* vmin, const float* vdiff,
                          const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq,
                          int k, int flags, float pad_value, int64_t id_offset, float* out_d, int64_t* out_i,
                          int64_t rows_per_query_hint, void* stream);

/* ---- product quantisation ("PQ<m>", "IVF<n>,PQ<m>", 8 bits per sub-quantiser) --------------------------------- */
/* faiss.index_factory(d, "IVF<n>,PQ<m>" | "PQ<m>") as reached from FaissFactoryIndexer (src/algorithms/modular.py:224-286;
 * configs/benchmark_config.yaml:36-50,61-72): IndexIVFPQ (by_residual) / IndexPQ.  codebooks [m][256][d/m] fp32 come
 * from the IVF k-means recipe run per sub-space by the host side.
 *   vdb_pq_encode         codes[i, s] = nearest of the 256 centroids of sub-quantiser s to x[i, s*dsub : (s+1)*dsub]
 *   vdb_pq_bias           bias[i] = |r^_i|^2 + 2 c_i . r^_i of the row's decoded residual and its list centroid (NULL = zero)
 *   vdb_pq_decode         out[i, :] = the vector row i's code stands for (+ its list centroid; NULL = zero), fp32 [n, ld_out]:
 *                         "PQ<m>" with many queries is searched as a flat scan over the decoded rows - the same distances
 *                         (|q - x^|^2 / q . x^) from the tensor pipe instead of m table look-ups per (query, row)
 *   vdb_bytes_fill        scatter byte rows [n, m] - and optionally one float per row - into interleaved-32 byte lists
 *                         (layout of vdb_sq8_fill with d := m; list_values [n_blocks * 32] float)
 *   vdb_ivf_pq_scan_topk  one table per query, T[s][j] = -2 q_s . cb[s][j] (inner product: q_s . cb[s][j]), in shared memory;
 *                         L2: |q - c|^2 + bias_row + sum_s T[s][code_s] (clamped at 0), inner product: q . c + sum_s T[s][code_s];
 *                         centroids == NULL = zero centroid, probes == NULL = the single list 0 (IndexPQ); list_bias may be
 *                         NULL for inner product; conventions as vdb_ivf_scan_topk */
int vdb_pq_encode(const float* x, int64_t n, int d, int64_t ld, const float* codebooks, int m, uint8_t* codes, void* stream);
int vdb_pq_bias(const uint8_t* codes, int64_t n, int d, int m, const float* codebooks, const float* centroids,
                const int32_t* assign, float* bias, void* stream);
int vdb_pq_decode(const uint8_t* codes, int64_t n, int d, int m, const float* codebooks, const float* centroids,
                  const int32_t* assign, float* out, int64_t ld_out, void* stream);
int vdb_bytes_fill(const uint8_t* rows, int64_t n, int m, const int32_t* assign, const int32_t* blk_off, int nlist,
                   int32_t* cursor, uint8_t* lists, int32_t* list_ids, const float* row_values, float* list_values,
                   void* stream);
int vdb_ivf_pq_scan_topk(int metric, const uint8_t* lists, const int32_t* list_ids, const float* list_bias,
                         const int32_t* blk_off, int nlist, int d, int m, const float* centroids, const float* codebooks,
                         const int64_t* probes, int nprobe, const float* q, int64_t ld_q, int64_t nq, int k, int flags,
                         float pad_value, int64_t id_offset, float* out_d, int64_t* out_i, void* stream);

/* ---- Hamming top-k on the tensor pipe (large bases, nbits <= 256) ----------------------------- */
/* Same contract as vdb_hamming_topk ((distance, id) order, out_d / out_i [nq,k] pre-filled by the
 * caller with padding), distances from a bf16 +-1 contraction on the tcgen05 pipeline of the flat
 * scan.  Operands: vdb_hamming_tc_expand turns packed codes [n, words] into bf16 rows of
 * vdb_hamming_tc_row_bytes(nbits) bytes for rows_pad >= n rows (rows_pad = vdb_flat_npad(n) for the
 * base together with norms [rows_pad], vdb_flat_nqpad(nq) and negate = 1 for the queries).  The packed
 * codes are still needed for the sampled bound.  65536 < n <= 8388608 (2^23). */
int vdb_hamming_tc_row_bytes(int nbits);
int vdb_hamming_tc_expand(const uint32_t* codes, int64_t n, int words, int nbits, int negate, void* out, float* norms,
                          int64_t rows_pad, void* stream);
size_t vdb_hamming_tc_workspace_bytes(int64_t nq, int nbits, int k, int64_t n);
int vdb_hamming_topk_tc(const void* base_bf16, const float* norms, const uint32_t* codes, int64_t n, const void* q_bf16,
                        const uint32_t* qcodes, int64_t nq, int nbits, int k, int64_t id_offset, float* out_d,
                        int64_t* out_i, void* workspace, size_t workspace_bytes, void* stream);
/* The same scan with fp16 operands (vdb_hamming_tc_expand with bit 1 of `negate` set: 2 = fp16 rows, 3 = fp16 and
 * negated) and fp16 ACCUMULATORS: sums of +-1 products are integers of magnitude <= nbits <= 256, exact in fp16, and
 * the epilogue reads the accumulators packed two per register (tcgen05.ld .pack::16b) and filters two keys per
 * instruction.  Same results bit for bit.  nbits even. */
int vdb_hamming_topk_tc_f16(const void* base_f16, const float* norms, const uint32_t* codes, int64_t n, const void* q_f16,
                            const uint32_t* qcodes, int64_t nq, int nbits, int k, int64_t id_offset, float* out_d,
                            int64_t* out_i, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDB_CUDA_H_ */
