/*
 * ppb.h — C ABI of the B200-native core/accessory sketch-distance engine (libppb.so).
 *
 * This is the drop-in boundary for ONE path of bacpop/PopPUNK: the all-vs-all and
 * query-vs-ref (core pi, accessory a) distance calculation that PopPUNK reaches through
 *     PopPUNK/sketchlib.py:475-632   queryDatabase()
 *         -> pp_sketchlib.queryDatabase(ref_db_name, query_db_name, rList, qList, klist,
 *                random_correct, jaccard, num_threads, use_gpu, device_id)
 *            (call sites PopPUNK/sketchlib.py:528-537, 547-564, 584-593, 601-618;
 *             positional order pinned by test/test-update-gpu.py:85-86)
 * and the step immediately after it,
 *     src/boundary.cpp:42-80         line_dist() / assign_threshold()
 *            (bound at src/python_bindings.cpp:18-25, 79-83; called models.py:1085-1089).
 *
 * Everything here is plain C: pointers, sizes, scalars.  No torch / pybind / Eigen types.
 * Functions never throw and never allocate caller-visible output; every output buffer is owned
 * by the caller.  Return value: 0 = ok, nonzero = error code (message via ppb_last_error()).
 *
 * ---- data conventions -------------------------------------------------------------------
 * Sketch (reference type: HDF5 dataset /sketches/<name>/<k>, schema PopPUNK/web.py:14-61;
 * fixture test/json_sketch.txt: sketchsize64=156, bbits=14, 2184 = 156*14 words per k):
 *   one genome, one k  = W = sketchsize64 * bbits  uint64 words, bindash bit-sliced:
 *   word [s*bbits + b] holds bit b of the bbits-bit signatures of bins 64s .. 64s+63.
 *   A "sketch array" is  uint64 [n][K][W]  row-major (genome-major, then k in klist order).
 *   bbits must be 14 (the only value pp-sketchlib writes).
 * Output row order (reference: PopPUNK/utils.py:199-226, src/boundary.cpp:22-37,97-123):
 *   self  (qry == NULL): condensed upper triangle, row(i<j) = n*i - i*(i+1)/2 + j - 1 - i
 *   non-self            : query-major rectangle,   row(q,r) = q*n_ref + r
 *   [row_begin,row_end) selects a contiguous shard of rows; out holds row_end-row_begin rows.
 * Random-match table (pp_sketchlib random_correct=True, sketchlib.py:533,589):
 *   rand_table float32 [n_clusters][n_clusters][K], entry (ref_cluster, qry_cluster, k);
 *   cluster ids uint16 per genome.  rand_table == NULL  <=>  random_correct=False.
 */
#ifndef PPB_H
#define PPB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPB_VERSION 100          /* 0.1.0 */
#define PPB_BBITS 14             /* signature bits per bin (pp-sketchlib constant) */
#define PPB_MAX_K 32             /* max number of k-mer lengths in one call */
#define PPB_MAX_PEERS 16         /* max peer output buffers of the fused multi-GPU form */
#define PPB_MIN_JACCARD_BINS 5   /* J < 5/S ends the regression series (docs/sketching.rst:161-165) */

/* output modes: what one output row holds */
#define PPB_OUT_DISTS   0        /* float32 [2]  (core, accessory)   jaccard=False */
#define PPB_OUT_JACCARD 1        /* float32 [K]  per-k Jaccard        jaccard=True  */
#define PPB_OUT_COUNTS  2        /* uint32  [K]  per-k matching-bin counts c_k (bit-exact probe) */

/* error codes */
#define PPB_OK            0
#define PPB_ERR_ARG       1
#define PPB_ERR_CUDA      2
#define PPB_ERR_NO_DEVICE 3
#define PPB_ERR_NOMEM     4

/* Decision boundary for the fused assign_threshold epilogue.
 * Replaces models.py:1085-1089 `assignThreshold(X/self.scale, slope, x_max, y_max)`:
 * x0 = core/scale_x, y0 = acc/scale_y (float32 division), then src/boundary.cpp:42-58. */
typedef struct ppb_boundary {
    int32_t slope;               /* 0 vertical, 1 horizontal, 2 sloped (boundary.cpp:45-55) */
    float   x_max, y_max;        /* narrowed double->float as in python_bindings.cpp:19-23 */
    float   scale_x, scale_y;    /* 1.0f for none */
} ppb_boundary;

int         ppb_version(void);
const char *ppb_last_error(void);
/* number of visible CUDA devices; <0 on error. */
int         ppb_device_count(void);

/* ---------------- index maps (src/boundary.cpp:18-37; utils.py:199-261) ---------------- */
int64_t ppb_square_to_condensed(int64_t i, int64_t j, int64_t n);
int64_t ppb_calc_row_idx(int64_t k, int64_t n);
int64_t ppb_calc_col_idx(int64_t k, int64_t i, int64_t n);
int64_t ppb_num_rows(int64_t n_ref, int64_t n_qry, int self);

/* ---------------- device-pointer entry points (caller owns all device memory) -----------
 * stream is a cudaStream_t passed as void*. All work is enqueued on it; no host sync unless
 * stated. Pointers are device pointers (torch tensor.data_ptr()).                           */

/* Size in bytes of the packed (lane-sliced) sketch buffer for n genomes. */
size_t ppb_packed_bytes(int64_t n, int32_t K, int32_t sketchsize64);

/* Re-lay a canonical sketch array uint64 [n_src][K][W] into the engine's packed layout.
 * idx (device int64 [n], may be NULL = identity with n == n_src) gathers genomes into list
 * order — the rList/qList subset+order semantics of pp_sketchlib.queryDatabase.             */
int ppb_pack_dev(const uint64_t *d_sketch, int64_t n_src, const int64_t *d_idx, int64_t n,
                 int32_t K, int32_t sketchsize64, uint32_t *d_packed, void *stream);

/* The same for a PART of the genomes, written to several packed arrays at once: genomes [g_begin, g_end) of the
 * n-genome layout (g_end may run into the padding, i.e. up to n rounded up to 128) are read from d_sketch_part
 * (row g - g_begin, or d_idx[g - g_begin]) and stored into each of the n_dst packed arrays d_packed[0..n_dst) —
 * this device's and, in a single-process multi-GPU call, the peers' (NVLink peer-mapped) — so every device uploads
 * and packs 1/G of the sketches instead of all of them.                                                        */
int ppb_pack_part_dev(const uint64_t *d_sketch_part, const int64_t *d_idx, int64_t g_begin, int64_t g_end, int64_t n,
                      int32_t K, int32_t sketchsize64, uint32_t *const *d_packed, int32_t n_dst, void *stream);

/* The hot path.  Replaces pp_sketchlib.queryDatabase(...) after the sketches are on the device.
 *   d_qry_packed == NULL  => self mode (n_qry ignored)
 *   kmers        host int32 [K], ascending k-mer lengths (x of the regression)
 *   d_rand_table device float32 [C][C][K] or NULL; d_ref_cluster/d_qry_cluster device uint16
 *   out_mode     PPB_OUT_*; d_out device buffer of (row_end-row_begin) rows of that mode,
 *                may be NULL when only labels are wanted (PPB_OUT_DISTS only)
 *   boundary/d_labels  optional fused assign_threshold: int8 label in {-1,0,+1} per row
 *   d_n_degenerate     optional device counter (uint64), incremented once per row whose
 *                      series has < 2 usable k (row gets (0,0)); caller zeroes it.           */
int ppb_query_dev(const uint32_t *d_ref_packed, int64_t n_ref,
                  const uint32_t *d_qry_packed, int64_t n_qry,
                  const int32_t *kmers, int32_t K, int32_t sketchsize64,
                  const float *d_rand_table, int32_t n_clusters,
                  const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                  int64_t row_begin, int64_t row_end,
                  int32_t out_mode, void *d_out,
                  const ppb_boundary *boundary, int8_t *d_labels,
                  unsigned long long *d_n_degenerate, void *stream);

/* Fused compute + exchange for multi-GPU runs (PPB_OUT_DISTS only).  Same as ppb_query_dev, but every result
 * row is additionally stored — from inside the kernel's epilogue, over NVLink — into each of the n_peers FULL
 * result buffers d_peer_out[g] (float32 [total_rows][2] on every GPU, indexed by GLOBAL row; e.g. the
 * buffer_ptrs of a torch symmetric-memory allocation, this GPU's own buffer included), or, when d_mc_out is not
 * NULL, once through that NVSwitch multicast address (multimem.st).  After all ranks have run it and passed a
 * barrier, every GPU holds the whole row-ordered result: no all-gather.  d_out (local shard) may be NULL.    */
int ppb_query_dev_fused(const uint32_t *d_ref_packed, int64_t n_ref,
                        const uint32_t *d_qry_packed, int64_t n_qry,
                        const int32_t *kmers, int32_t K, int32_t sketchsize64,
                        const float *d_rand_table, int32_t n_clusters,
                        const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                        int64_t row_begin, int64_t row_end, void *d_out,
                        void *const *d_peer_out, int32_t n_peers, void *d_mc_out,
                        unsigned long long *d_n_degenerate, void *stream);

/* Standalone assign_threshold over an existing (n,2) float32 row-major array
 * (src/boundary.cpp:60-80). d_out float32 [n] in {-1,0,+1}.                                  */
int ppb_assign_threshold_dev(const float *d_dists, int64_t n, int32_t slope,
                             float x_max, float y_max, float *d_out, void *stream);

/* ---------------- "next" rows: consumers of the path's output ----------------------
 *
 * N1 — edge lists after the threshold.
 * ppb_query_edges_dev: the distance kernel with the boundary test fused AND the within-boundary pairs appended,
 *   as GLOBAL row indices, to d_edge_rows (unordered, warp-aggregated atomics; *d_edge_count may exceed
 *   capacity — then only `capacity` rows were stored).  include_boundary = 0: line_dist < 0 (what
 *   generate_tuples does with assign_threshold labels == -1, network.py:1180-1184); 1: line_dist <= 0 (what
 *   edge_iterate does, src/boundary.cpp:86-87).  d_out / d_labels are optional (NULL = not written).
 * ppb_rows_to_pairs_dev: rows -> (i, j) sample pairs with generate_tuples' conventions
 *   (src/boundary.cpp:97-123: self -> calc_row_idx/calc_col_idx + int_offset; non-self -> i = row % num_ref +
 *   int_offset, j = row / num_ref + num_ref + int_offset; swapped so that i <= j).
 * ppb_edges_from_dists_dev / ppb_edges_from_labels_dev: ORDERED (ascending row, like the reference loops)
 *   compaction of an existing (n,2) distance array (edge_iterate, src/boundary.cpp:82-95) or label array
 *   (generate_tuples; label_dtype 0 = int8, 1 = int32, 2 = float32).  d_scratch: ppb_edges_scratch_bytes(n_rows).
 *   *d_count receives the number of edges found (pairs beyond `capacity` are not written).                  */
int ppb_query_edges_dev(const uint32_t *d_ref_packed, int64_t n_ref,
                        const uint32_t *d_qry_packed, int64_t n_qry,
                        const int32_t *kmers, int32_t K, int32_t sketchsize64,
                        const float *d_rand_table, int32_t n_clusters,
                        const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                        int64_t row_begin, int64_t row_end,
                        const ppb_boundary *boundary, int32_t include_boundary,
                        int64_t *d_edge_rows, int64_t capacity, unsigned long long *d_edge_count,
                        void *d_out, int8_t *d_labels,
                        unsigned long long *d_n_degenerate, void *stream);
int ppb_rows_to_pairs_dev(const int64_t *d_rows, int64_t n, int32_t self, int64_t n_samples_or_num_ref,
                          int64_t int_offset, int64_t *d_i, int64_t *d_j, void *stream);
/* Ascending in-place sort of n row indices in [0, max_row] (what brings ppb_query_edges_dev's unordered rows into the
 * reference's row order): least-significant-digit radix sort, 8 bits per pass, hand-written (csrc/ppb_sort.cuh).   */
int ppb_sort_rows_dev(int64_t *d_rows, int64_t n, int64_t max_row, void *stream);
size_t ppb_edges_scratch_bytes(int64_t n_rows);
int ppb_edges_from_dists_dev(const float *d_dists, int64_t n_rows, int64_t n_samples, int32_t slope,
                             float x_max, float y_max, int64_t *d_i, int64_t *d_j, int64_t capacity,
                             int64_t *d_count, void *d_scratch, void *stream);
int ppb_edges_from_labels_dev(const void *d_labels, int32_t label_dtype, int64_t n_rows, int32_t within_label,
                              int32_t self, int64_t n_samples_or_num_ref, int64_t int_offset,
                              int64_t *d_i, int64_t *d_j, int64_t capacity, int64_t *d_count,
                              void *d_scratch, void *stream);

/* N1 (rest) — src/boundary.cpp:125-237, bound at src/python_bindings.cpp:42-77.
 * ppb_generate_all_tuples_dev: generate_all_tuples — every pair of the triangle (self) or rectangle, as (i, j);
 *   d_i / d_j hold n(n-1)/2 (self) or num_ref*num_queries entries.
 * ppb_threshold_iterate_1d_dev: threshold_iterate_1D — the boundary moves along the line (x0,y0)->(x1,y1) by the
 *   (sorted, HOST) offsets; rows are ranked once by their signed distance to the first boundary (stable) and each
 *   offset admits the next rows of that order with line_dist <= 0.  Outputs (i, j, index of the admitting offset)
 *   in admission order; *d_count = number admitted (<= n_rows; rows beyond `capacity` are not written).
 *   Synchronises the stream once (it has to know whether the exact sequential walk is needed).
 * ppb_threshold_iterate_2d_dev: threshold_iterate_2D — sloped boundaries (x_max[o], y_max), x_max sorted, HOST array
 *   of at most 1024 entries; step o admits, in row order, the rows inside boundary o and outside boundary o-1.    */
int ppb_generate_all_tuples_dev(int64_t num_ref, int64_t num_queries, int32_t self, int64_t int_offset,
                                int64_t *d_i, int64_t *d_j, void *stream);
int ppb_threshold_iterate_1d_dev(const float *d_dists, int64_t n_rows, const double *offsets, int32_t n_off,
                                 int32_t slope, float x0, float y0, float x1, float y1,
                                 int64_t *d_i, int64_t *d_j, int64_t *d_off, int64_t capacity, int64_t *d_count,
                                 void *stream);
int ppb_threshold_iterate_2d_dev(const float *d_dists, int64_t n_rows, const float *x_max, int32_t n_off, float y_max,
                                 int64_t *d_i, int64_t *d_j, int64_t *d_off, int64_t capacity, int64_t *d_count,
                                 void *stream);

/* N3 — nearest-neighbour extraction for the lineage models (src/extend.cpp:52-289, bound at
 * src/python_bindings.cpp:109-136; callers PopPUNK/models.py:1177-1184, 1215-1222, 1366-1372, assign.py:680-686,
 * mandrake.py:67).  Sparse matrices are COO triples (i sorted ascending, as every producer here emits them).
 * ppb_knn_dev: get_kNN_distances — per row of a dense rows x cols float32 matrix the kNN smallest entries, ties to
 *   the lower column, never column == row; outputs rows*kNN long (zeros where a row has fewer candidates).
 * ppb_lower_rank_dev: lower_rank — per sample its sparse neighbours in ascending (stable) order, kept while the
 *   reference's rank rule holds (plain: kNN+1 entries; count_unique_distances: distances differing by >= epsilon
 *   count as new ranks), optionally only reciprocal (i<j) pairs.  Outputs sized nnz; *d_count = entries written.
 *   At most 1024 sparse neighbours per sample.  Synchronises the stream.
 * ppb_extend_dev: extend — the kNN of every reference (sparse ref-ref row + dense row of qr [nr x nq]) and every
 *   query (column of qr + row of qq [nq x nq]); the dense query part wins ties; outputs sized (nr+nq)*kNN.
 * kNN <= 2048.                                                                                                */
int ppb_knn_dev(const float *d_mat, int64_t rows, int64_t cols, int32_t knn,
                int64_t *d_i, int64_t *d_j, float *d_d, void *stream);
int ppb_lower_rank_dev(const int64_t *d_sp_i, const int64_t *d_sp_j, const float *d_sp_d, int64_t nnz,
                       int64_t n_samples, int64_t knn, int32_t reciprocal_only, int32_t count_unique_distances,
                       float epsilon, int64_t *d_i, int64_t *d_j, float *d_d, int64_t *d_count, void *stream);
int ppb_extend_dev(const int64_t *d_sp_i, const int64_t *d_sp_j, const float *d_sp_d, int64_t nnz,
                   const float *d_qq, const float *d_qr, int64_t nr, int64_t nq, int32_t knn,
                   int64_t *d_i, int64_t *d_j, float *d_d, int64_t *d_count, void *stream);

/* N2 — long <-> square reshapes (pp_sketchlib.longToSquare / squareToLong / longToSquareMulti; call sites
 * PopPUNK/utils.py:393-405, network.py:2133-2134, models.py:1217,1357, mandrake.py:165).  float32; the
 * condensed / rectangular vectors are read with an element stride so a column of the (n_pairs,2) distance
 * array can be passed in place (stride 2).  Squares are row-major n x n, symmetric, zero diagonal.          */
int ppb_long_to_square_dev(const float *d_vec, int64_t stride, int64_t n, float *d_square, void *stream);
int ppb_square_to_long_dev(const float *d_square, int64_t n, float *d_vec, void *stream);
int ppb_long_to_square_multi_dev(const float *d_rr, int64_t stride_rr, const float *d_qr, int64_t stride_qr,
                                 const float *d_qq, int64_t stride_qq, int64_t n_ref, int64_t n_qry,
                                 float *d_square, void *stream);

/* ---------------- host-buffer entry points (what a pybind/ctypes shim of PopPUNK calls) --
 * Same semantics as above with HOST pointers: the library stages host->device copies,
 * packs, runs the kernels in row chunks and copies results back, overlapping copy and
 * compute on internal streams.  Blocking.  out may be pinned or pageable.                    */
int ppb_query_host(const uint64_t *ref, int64_t n_ref,
                   const uint64_t *qry, int64_t n_qry,
                   const int32_t *kmers, int32_t K, int32_t sketchsize64, int32_t bbits,
                   const float *rand_table, int32_t n_clusters,
                   const uint16_t *ref_cluster, const uint16_t *qry_cluster,
                   int64_t row_begin, int64_t row_end,
                   int32_t out_mode, void *out,
                   const ppb_boundary *boundary, int8_t *labels,
                   int64_t *n_degenerate, int32_t device_id);

/* The same call on SEVERAL devices of one process — what the drop-in queryDatabase() uses (PopPUNK is one process:
 * PopPUNK/utils.py:118-126, __main__.py:220): the row range is cut into n_devices contiguous shards of equal pair
 * count on row-tile boundaries (ppb_plan_device_shards), one worker thread per device uploads and packs 1/G of the
 * reference sketches into every device's packed array (peer stores over NVLink; without peer access every device
 * uploads all of them), uploads only ITS OWN queries in non-self mode (rows are query-major), runs its shard in row
 * chunks and copies them back straight into the ONE caller buffer.  No exchange between devices.  A job with fewer
 * than ~16 Mi rows per device uses fewer devices, and so does a PAGEABLE destination (it is drained by the host cores,
 * which one or two devices already out-produce; PPB_STAGED_ALL_DEVICES=1 overrides).  Page-locked destinations
 * (cudaHostAlloc / cudaHostRegister / ppb_host_alloc blocks after their second reuse) are written by direct DMA.
 * ppb_query_host(…, device_id) is this call with one device.                                                     */
int ppb_query_host_multi(const uint64_t *ref, int64_t n_ref,
                         const uint64_t *qry, int64_t n_qry,
                         const int32_t *kmers, int32_t K, int32_t sketchsize64, int32_t bbits,
                         const float *rand_table, int32_t n_clusters,
                         const uint16_t *ref_cluster, const uint16_t *qry_cluster,
                         int64_t row_begin, int64_t row_end,
                         int32_t out_mode, void *out,
                         const ppb_boundary *boundary, int8_t *labels,
                         int64_t *n_degenerate, const int32_t *device_ids, int32_t n_devices);
/* The n_devices+1 row boundaries ppb_query_host_multi uses (cuts[g] .. cuts[g+1] is device g's shard). Host-only. */
int64_t ppb_plan_device_shards(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end,
                               int32_t n_devices, int64_t *cuts);

/* Host memory for results the LIBRARY's caller hands on as its own (the drop-in returns a NumPy array it owns):
 * ppb_host_alloc returns a block of at least `bytes` (anonymous mapping, transparent huge pages requested) from a
 * small pool; ppb_host_free hands it back.  A reused block has been touched (a staged copy into it no longer
 * page-faults) and is page-locked (cudaHostRegister) on its second reuse, after which results are DMA-ed straight
 * into the array the caller receives.  PPB_HOST_PIN_AFTER=n changes that count (0 = never page-lock);
 * PPB_HOST_POOL_MAX_GB bounds the idle bytes kept (default: half of physical memory); ppb_release_workspace() drops
 * idle blocks.                                                                                                   */
void *ppb_host_alloc(size_t bytes);
int   ppb_host_free(void *p);
int   ppb_host_pool_stats(size_t *bytes_held, size_t *bytes_in_use, size_t *bytes_pinned);

/* How ppb_query_host cuts [row_begin,row_end) into launches: chunks of at most cap_rows rows that end on row-tile
 * boundaries (64 genomes of the row side) whenever a whole row tile fits.  Writes up to max_chunks (begin,end)
 * pairs into bounds (may be NULL) and returns the number of chunks, or -1 on a bad argument.  Host-only.     */
int64_t ppb_plan_host_chunks(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end,
                             int64_t cap_rows, int64_t *bounds, int64_t max_chunks);

/* The tile schedule of one ppb_query_dev launch over [row_begin,row_end): (row tile, column tile) index pairs in
 * execution order — row tiles are PPB_TILE_ROWS genomes of the row side, column tiles tile_cols reference genomes;
 * bands of band_tiles row tiles, the column tile the slow index inside a band; self mode skips tiles with no j > i.
 * Writes up to max_tiles pairs into tiles (int32 [2*max_tiles], may be NULL), returns the tile count or -1.  Host-only. */
#define PPB_TILE_ROWS 64
int64_t ppb_plan_tiles(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end,
                       int32_t tile_cols, int32_t band_tiles, int32_t *tiles, int64_t max_tiles);

/* Frees the device workspace the host-buffer calls keep between invocations (grow-only, per device). */
int ppb_release_workspace(void);

int ppb_assign_threshold_host(const float *dists, int64_t n, int32_t slope,
                              float x_max, float y_max, float *out, int32_t device_id);

/* ---------------- measurement helpers ---------------------------------------------------
 * Integer-pipe micro-roofline: runs a LOP3-only (mode 0), POPC-only (mode 1) or
 * LOP3+POPC mixed (mode 2) kernel on `stream` and returns lane-ops executed; caller times it. */
int ppb_microbench_dev(int32_t mode, int64_t iters, uint32_t *d_sink, int64_t *lane_ops, void *stream);

/* The distance kernel's column body (LDS + 14 LOP3 per row and 32 bins + POPC + pack + REDUX + STS) as a stand-alone
 * loop with rows_per_warp (4, 5 or 8) register-stationary rows and warps_per_scheduler (1..4) warps per SM scheduler,
 * one CTA per SM, no barriers: how much of the LOP3 pipe that instruction mix can reach.  Returns LOP3 lane-ops.   */
int ppb_microbench_mix_dev(int32_t rows_per_warp, int32_t warps_per_scheduler, int32_t with_lds, int64_t iters,
                           uint32_t *d_sink, int64_t *lop3_lane_ops, void *stream);

/* kernel launch counter since library load (for bench.py's gpu_launches). */
int64_t ppb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PPB_H */
