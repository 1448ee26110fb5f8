// The rest of poppunk_refine's consumers of the distance path:
//
//   N1  threshold_iterate_1D / threshold_iterate_2D / generate_all_tuples   (src/boundary.cpp:125-237)
//   N3  get_kNN_distances / lower_rank / extend                              (src/extend.cpp:52-289)
//
// All of it is index / compare work on the (n_pairs, 2) distance array or on dense distance rows: memory-bound,
// one pass where the reference makes one, results in the reference's own order (stable sorts: ties keep index
// order), bit-exact.  Checked against the reference's own sources compiled into oracle/_ref.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ppb_next.cuh"

namespace ppb {

// float -> uint32 whose unsigned order is the float order; -0 and +0 tie (the reference compares with `<`)
__device__ __forceinline__ uint32_t float_key(float v) {
    uint32_t b = __float_as_uint(v);
    if (v == 0.0f) b = 0;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ------------------------------------------------------------------------------------------------------------
// generate_all_tuples (boundary.cpp:125-149)
// ------------------------------------------------------------------------------------------------------------
__global__ void all_tuples_kernel(int64_t num_ref, int64_t num_queries, int32_t self, int64_t int_offset, int64_t total,
                                  int64_t *__restrict__ out_i, int64_t *__restrict__ out_j) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i, j;
        if (self) {
            row_to_pair(PairMap{1, num_ref, int_offset}, t, i, j);
        } else {  // refs outer, queries inner; int_offset is not applied on this branch (:143-147)
            i = t % num_queries;
            j = t / num_queries + num_ref;
        }
        out_i[t] = i;
        out_j[t] = j;
    }
}

// ------------------------------------------------------------------------------------------------------------
// threshold_iterate_2D (boundary.cpp:211-237): step o admits, in row order, rows inside boundary o that were
// outside boundary o-1.  Output order is (o, row) — a stable counting sort by o.  Pass 0 counts per (o, block)
// into cnt[o * n_blocks + block]; ONE exclusive scan of that o-major array gives every final position; pass 1
// recomputes and writes.  A thread keeps its 16 rows in registers and tests them against every boundary.
// ------------------------------------------------------------------------------------------------------------
constexpr int kIterMaxOffsets = 1024;

__global__ void __launch_bounds__(kSelThreads) iterate2d_kernel(const float2 *__restrict__ d, int64_t n_rows,
                                                                const float *__restrict__ x_max, int32_t n_off,
                                                                float y_max, int pass, int64_t *__restrict__ cnt,
                                                                int64_t n_blocks, int64_t capacity,
                                                                int64_t *__restrict__ out_i, int64_t *__restrict__ out_j,
                                                                int64_t *__restrict__ out_o) {
    __shared__ uint32_t warp_tot[kIterMaxOffsets][kSelThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * kSelBlockRows + (int64_t)warp * (32 * kSelIters);
    const int64_t n_samples = (int64_t)(0.5 * (1.0 + sqrt(1.0 + 8.0 * (double)n_rows)));
    float2 v[kSelIters];
#pragma unroll
    for (int m = 0; m < kSelIters; m++) {
        const int64_t row = base + m * 32 + lane;
        v[m] = row < n_rows ? d[row] : make_float2(0.f, 0.f);
    }
    // in(o, row) is evaluated once per step: "outside boundary o-1" is the previous step's result, kept as one bit per
    // row; the sloped form of line_dist is written out with its uniform product hoisted (same operations, same
    // order as boundary.cpp:48-50), the sqrt form (a zero intercept) stays behind a uniform branch.
    auto inside = [&](const float2 &p2, float xm, float prod, bool degenerate) -> bool {
        const float side = degenerate ? line_dist(p2.x, p2.y, xm, y_max, 2)
                                      : __fsub_rn(__fadd_rn(__fmul_rn(p2.y, xm), __fmul_rn(p2.x, y_max)), prod);
        return side <= 0.0f;
    };
    uint32_t valid = 0;
#pragma unroll
    for (int m = 0; m < kSelIters; m++) valid |= (base + m * 32 + lane < n_rows ? 1u : 0u) << m;
    uint32_t prev = 0;
    for (int o = 0; o < n_off; o++) {
        const float xm = x_max[o], prod = __fmul_rn(xm, y_max);
        const bool degenerate = xm == 0.0f || y_max == 0.0f;
        uint32_t cur = 0;
#pragma unroll
        for (int m = 0; m < kSelIters; m++) cur |= (inside(v[m], xm, prod, degenerate) ? 1u : 0u) << m;
        const uint32_t admit = cur & ~prev & valid;
        const uint32_t tot = __reduce_add_sync(0xffffffffu, __popc(admit));  // the warp's admissions at this step
        prev = cur;
        if (lane == 0) warp_tot[o][warp] = tot;
    }
    __syncthreads();
    if (pass == 0) {
        for (int o = threadIdx.x; o < n_off; o += kSelThreads) {
            uint32_t s = 0;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; w++) s += warp_tot[o][w];
            cnt[(int64_t)o * n_blocks + blockIdx.x] = s;
        }
        return;
    }
    prev = 0;
    for (int o = 0; o < n_off; o++) {
        int64_t pos = cnt[(int64_t)o * n_blocks + blockIdx.x];
        for (int w = 0; w < warp; w++) pos += warp_tot[o][w];
        const float xm = x_max[o], prod = __fmul_rn(xm, y_max);
        const bool degenerate = xm == 0.0f || y_max == 0.0f;
        uint32_t cur = 0;
#pragma unroll
        for (int m = 0; m < kSelIters; m++) cur |= (inside(v[m], xm, prod, degenerate) ? 1u : 0u) << m;
        const uint32_t admit = cur & ~prev & valid;
        prev = cur;
        if (warp_tot[o][warp] == 0) continue;  // nothing of this warp is admitted at this step (the common case)
#pragma unroll
        for (int m = 0; m < kSelIters; m++) {
            const bool sel = (admit >> m) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, sel);
            if (sel) {
                const int64_t at = pos + __popc(b & ((1u << lane) - 1));
                if (at < capacity) {
                    const int64_t row = base + m * 32 + lane;
                    const int64_t i = dev_row_idx(row, n_samples);
                    out_i[at] = i;
                    out_j[at] = dev_col_idx(row, i, n_samples);
                    out_o[at] = o;
                }
            }
            pos += __popc(b);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// threshold_iterate_1D (boundary.cpp:151-209).  The reference ranks all rows once by their signed distance to the
// FIRST boundary (stable sort), then walks that order: offset o admits rows while line_dist_o(row) <= 0.  So the
// row at sorted position p is admitted by  o_p = min{o >= o_{p-1} : line_dist_o(row_p) <= 0}  and the walk ends
// at the first row no offset admits.  When each row's test is monotone in o (false..false true..true — always,
// up to float rounding) this is a running maximum of the rows' first admitting offsets t_p: a scan.  Rows whose
// test is not monotone are counted; if there are any, the exact sequential walk is run instead (one thread).
// ------------------------------------------------------------------------------------------------------------
__global__ void iterate1d_keys_kernel(const float2 *__restrict__ d, int64_t n_rows, int32_t slope, float x_max, float y_max,
                                      uint32_t *__restrict__ keys, int64_t *__restrict__ rows) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const float2 v = d[r];
        keys[r] = float_key(line_dist(v.x, v.y, x_max, y_max, slope));
        rows[r] = r;
    }
}

// The fast path sorts only the rows that are EVER admitted.  iterate1d_classify_kernel makes one coalesced pass over
// the unsorted array: key (distance to the first boundary), first admitting offset (n_off = never), the count of rows
// whose test is not monotone in the offset, and per block the smallest (key, row) among the never-admitted rows — the
// reference's walk stops at the first such row of the sorted order, so exactly the admitted rows that sort BEFORE that
// (key, row) are emitted.  The admitted rows are then compacted (row order), sorted by key (stable), and walked.
constexpr int kCls1dRows = 16;
__global__ void __launch_bounds__(256) iterate1d_classify_kernel(const float2 *__restrict__ d, int64_t n_rows, int32_t slope,
                                                                 const float2 *__restrict__ bnd, int32_t n_off,
                                                                 uint32_t *__restrict__ key, uint16_t *__restrict__ first,
                                                                 unsigned long long *__restrict__ n_irregular,
                                                                 uint32_t *__restrict__ blk_key, int64_t *__restrict__ blk_row,
                                                                 const StepSearch search) {
    __shared__ uint32_t wk[8];
    __shared__ int64_t wr[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * (256 * kCls1dRows) + (int64_t)warp * (32 * kCls1dRows);
    float2 v[kCls1dRows];
#pragma unroll
    for (int m = 0; m < kCls1dRows; m++) v[m] = __ldg(d + min(base + m * 32 + lane, n_rows - 1));
    uint32_t kmin = 0xffffffffu;
    int64_t rmin = INT64_MAX;
    uint32_t irregular = 0;
#pragma unroll 2
    for (int m = 0; m < kCls1dRows; m++) {
        const int64_t row = base + m * 32 + lane;
        if (row >= n_rows) continue;
        int32_t t = first_admitting_step(search, v[m]);   // bisection where every test's sign is certain
        bool bad = false;
        if (t < 0) {
            t = n_off;
            for (int o = 0; o < n_off; o++) {
                const float2 b = __ldg(bnd + o);
                const bool in = line_dist(v[m].x, v[m].y, b.x, b.y, slope) <= 0.0f;
                if (in && t == n_off) t = o;
                if (!in && t != n_off) bad = true;
            }
        }
        const float2 b0 = __ldg(bnd);
        const uint32_t k = float_key(line_dist(v[m].x, v[m].y, b0.x, b0.y, slope));
        key[row] = k;
        first[row] = (uint16_t)t;
        irregular += bad;
        if (t == n_off && (k < kmin || (k == kmin && row < rmin))) {
            kmin = k;
            rmin = row;
        }
    }
    irregular = __reduce_add_sync(0xffffffffu, irregular);
    if (lane == 0 && irregular) atomicAdd(n_irregular, (unsigned long long)irregular);
    // lexicographic minimum of (key, row): over the warp, then over the block
    const uint32_t k_w = __reduce_min_sync(0xffffffffu, kmin);
    const unsigned long long r_mine = kmin == k_w ? (unsigned long long)rmin : ~0ull;
    const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(r_mine >> 32));
    const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(r_mine >> 32) == hi ? (unsigned)r_mine : 0xffffffffu);
    if (lane == 0) {
        wk[warp] = k_w;
        wr[warp] = (int64_t)(((unsigned long long)hi << 32) | lo);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bk = 0xffffffffu;
        int64_t br = INT64_MAX;
        for (int w = 0; w < 8; w++)
            if (wk[w] < bk || (wk[w] == bk && wr[w] < br)) {
                bk = wk[w];
                br = wr[w];
            }
        blk_key[blockIdx.x] = bk;
        blk_row[blockIdx.x] = br;
    }
}
// (key, row) where the walk stops = the minimum over the blocks; cut[0] = key, cut[1] = row (INT64_MAX: no never-admitted row)
__global__ void __launch_bounds__(1024) iterate1d_cut_kernel(const uint32_t *__restrict__ blk_key, const int64_t *__restrict__ blk_row,
                                                             int64_t n_blocks, unsigned long long *__restrict__ cut) {
    __shared__ uint32_t sk[1024];
    __shared__ int64_t sr[1024];
    uint32_t bk = 0xffffffffu;
    int64_t br = INT64_MAX;
    for (int64_t b = threadIdx.x; b < n_blocks; b += 1024) {
        const uint32_t k = blk_key[b];
        const int64_t r = blk_row[b];
        if (k < bk || (k == bk && r < br)) {
            bk = k;
            br = r;
        }
    }
    sk[threadIdx.x] = bk;
    sr[threadIdx.x] = br;
    __syncthreads();
    for (int s2 = 512; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) {
            const uint32_t k = sk[threadIdx.x + s2];
            const int64_t r = sr[threadIdx.x + s2];
            if (k < sk[threadIdx.x] || (k == sk[threadIdx.x] && r < sr[threadIdx.x])) {
                sk[threadIdx.x] = k;
                sr[threadIdx.x] = r;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cut[0] = sk[0];
        cut[1] = (unsigned long long)sr[0];
    }
}
struct PredAdmitted {  // rows some offset admits
    typedef uint16_t Value;
    const uint16_t *first;
    int32_t n_off;
    __device__ __forceinline__ uint16_t load(int64_t row) const { return __ldg(first + row); }
    __device__ __forceinline__ bool test(const uint16_t &t) const { return (int32_t)t < n_off; }
};
struct OutKeyed {  // compacted (key, row | first << 48) pairs, in row order
    const uint32_t *key;
    const uint16_t *first;
    uint32_t *ck;
    int64_t *cv;
    __device__ __forceinline__ void write(int64_t at, int64_t row) const {
        ck[at] = key[row];
        cv[at] = row | ((int64_t)first[row] << 48);
    }
};
// sorted admitted rows -> order / first arrays + block maxima of `first` (the walk's running maximum is scanned from them)
__global__ void __launch_bounds__(1024) iterate1d_unpack_kernel(const int64_t *__restrict__ cv, int64_t m, int64_t *__restrict__ order,
                                                                int32_t *__restrict__ first, int32_t *__restrict__ block_max) {
    __shared__ int32_t smax[32];
    const int64_t p = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int32_t t = -1;
    if (p < m) {
        const int64_t v = cv[p];
        t = (int32_t)(v >> 48);
        order[p] = v & (((int64_t)1 << 48) - 1);
        first[p] = t;
    }
    const int32_t wm = __reduce_max_sync(0xffffffffu, t);
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t mx = -1;
        for (int w = 0; w < 32; w++) mx = max(mx, smax[w]);
        block_max[blockIdx.x] = mx;
    }
}

constexpr int kScanBlock = 1024;

// first admitting offset of every sorted row (n_off = none) + block maxima + count of non-monotone rows
__global__ void __launch_bounds__(kScanBlock) iterate1d_first_kernel(const float2 *__restrict__ d,
                                                                     const int64_t *__restrict__ order, int64_t n_rows,
                                                                     int32_t slope, const float2 *__restrict__ bnd,
                                                                     int32_t n_off, int32_t *__restrict__ first,
                                                                     int32_t *__restrict__ block_max,
                                                                     unsigned long long *__restrict__ n_irregular) {
    __shared__ int32_t smax[kScanBlock / 32];
    const int64_t p = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int32_t t = -1;
    if (p < n_rows) {
        const float2 v = d[order[p]];
        bool irregular = false;
        t = n_off;
        for (int o = 0; o < n_off; o++) {
            const bool in = line_dist(v.x, v.y, bnd[o].x, bnd[o].y, slope) <= 0.0f;
            if (in && t == n_off) t = o;
            if (!in && t != n_off) irregular = true;
        }
        first[p] = t;
        if (irregular) atomicAdd(n_irregular, 1ull);
    }
    const int32_t wm = __reduce_max_sync(0xffffffffu, t);
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t m = -1;
        for (int w = 0; w < kScanBlock / 32; w++) m = max(m, smax[w]);
        block_max[blockIdx.x] = m;
    }
}

// exclusive running maximum over the block maxima, in place (one CTA)
__global__ void __launch_bounds__(1024) max_scan_kernel(int32_t *__restrict__ a, int64_t n) {
    __shared__ int32_t wmax[32];
    __shared__ int32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = -1;
    __syncthreads();
    for (int64_t start = 0; start < n; start += 1024) {
        const int64_t idx = start + threadIdx.x;
        const int32_t v = idx < n ? a[idx] : -1;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x = max(x, y);
        }
        if (lane == 31) wmax[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int32_t s = wmax[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s = max(s, y);
            }
            wmax[lane] = s;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        const int32_t incl = max(max(x, warp ? wmax[warp - 1] : -1), carry);
        // exclusive value = max of everything strictly before idx
        int32_t excl = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) excl = -1;
        excl = max(max(excl, warp ? wmax[warp - 1] : -1), carry);
        if (idx < n) a[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

// o_p = running max; rows are emitted while o_p < n_off.  The number emitted = first p with o_p == n_off.
__global__ void __launch_bounds__(kScanBlock) iterate1d_emit_kernel(const int64_t *__restrict__ order,
                                                                    const int32_t *__restrict__ first,
                                                                    const int32_t *__restrict__ block_excl, int64_t n_rows,
                                                                    int32_t n_off, int64_t capacity,
                                                                    int64_t *__restrict__ out_i, int64_t *__restrict__ out_j,
                                                                    int64_t *__restrict__ out_o,
                                                                    unsigned long long *__restrict__ n_emit,
                                                                    int64_t n_rows_total, const uint32_t *__restrict__ keys,
                                                                    const unsigned long long *__restrict__ cut) {
    __shared__ int32_t wmax[kScanBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int64_t n_samples = (int64_t)(0.5 * (1.0 + sqrt(1.0 + 8.0 * (double)n_rows_total)));
    int32_t x = p < n_rows ? first[p] : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = max(x, y);
    }
    if (lane == 31) wmax[warp] = x;
    __syncthreads();
    int32_t run = block_excl[blockIdx.x];
    for (int w = 0; w < warp; w++) run = max(run, wmax[w]);
    x = max(x, run);  // o_p
    bool admitted = p < n_rows && x < n_off;
    if (admitted && cut) {  // only what sorts before the first never-admitted row: that is where the reference's walk stops
        const uint32_t ck = (uint32_t)cut[0], k = keys[p];
        admitted = k < ck || (k == ck && order[p] < (int64_t)cut[1]);
    }
    if (admitted && p < capacity) {
        const int64_t row = order[p];
        const int64_t i = dev_row_idx(row, n_samples);
        out_i[p] = i;
        out_j[p] = dev_col_idx(row, i, n_samples);
        out_o[p] = x;
    }
    // admitted rows are a prefix of the order: the count is the largest admitted position + 1.  One atomic per BLOCK
    // (one per row on a single address serialised the whole kernel: 76 of its 76 ms at 400 M rows).
    const unsigned long long mine = admitted ? (unsigned long long)(p + 1) : 0ull;
    const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(mine >> 32));
    const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(mine >> 32) == hi ? (unsigned)mine : 0u);
    __shared__ unsigned long long wbest[kScanBlock / 32];
    if (lane == 0) wbest[warp] = ((unsigned long long)hi << 32) | lo;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long best = 0;
        for (int w = 0; w < kScanBlock / 32; w++) best = max(best, wbest[w]);
        if (best) atomicMax(n_emit, best);
    }
}

// the reference's walk, verbatim in one thread: only used when some row's test is not monotone in the offset
__global__ void iterate1d_sequential_kernel(const float2 *__restrict__ d, const int64_t *__restrict__ order, int64_t n_rows,
                                            int32_t slope, const float2 *__restrict__ bnd, int32_t n_off, int64_t capacity,
                                            int64_t *__restrict__ out_i, int64_t *__restrict__ out_j,
                                            int64_t *__restrict__ out_o, unsigned long long *__restrict__ n_emit) {
    if (blockIdx.x || threadIdx.x) return;
    const int64_t n_samples = (int64_t)(0.5 * (1.0 + sqrt(1.0 + 8.0 * (double)n_rows)));
    int64_t p = 0;
    for (int o = 0; o < n_off && p < n_rows; o++) {
        while (p < n_rows) {
            const int64_t row = order[p];
            const float2 v = d[row];
            if (!(line_dist(v.x, v.y, bnd[o].x, bnd[o].y, slope) <= 0.0f)) break;
            if (p < capacity) {
                const int64_t i = dev_row_idx(row, n_samples);
                out_i[p] = i;
                out_j[p] = dev_col_idx(row, i, n_samples);
                out_o[p] = o;
            }
            p++;
        }
    }
    *n_emit = (unsigned long long)p;
}

// ------------------------------------------------------------------------------------------------------------
// k smallest of a per-row candidate list, stable (ties: lower list position first), a given j excluded.
// One CTA per row.  (1) radix select, most significant byte first, over the 32-bit float keys (4 histogram passes
// over the row, which stays in L2) -> the kNN-th smallest key T and how many keys are below it; (2) collect all
// candidates below T plus the first (kNN - below) with key == T in position order (block-ordered ballot scan);
// (3) bitonic sort of the <= kNN (key, position) pairs in shared memory; (4) write (row, j, dist).
// The candidate list is a functor, so get_kNN_distances (a dense matrix row) and extend (dense query part
// followed by the sparse / dense reference part) share the kernel.
// ------------------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 256;  // also the number of radix buckets: one thread folds one bucket
constexpr int kKnnMax = 2048;
static_assert(kKnnThreads == 256, "one thread per radix bucket when the per-warp histograms are folded");

struct DenseRowCands {  // get_kNN_distances (extend.cpp:245-289): row r of a rows x cols matrix, j = column
    const float *mat;
    int64_t cols;
    int64_t rows;
    __device__ __forceinline__ int64_t length(int64_t) const { return cols; }
    __device__ __forceinline__ float dist(int64_t r, int64_t pos) const { return mat[r * cols + pos]; }
    __device__ __forceinline__ int64_t j_of(int64_t, int64_t pos) const { return pos; }
    __device__ __forceinline__ int64_t out_pos(int64_t r, int32_t knn) const { return r * knn; }
    __device__ __forceinline__ bool pad() const { return true; }  // rows*kNN outputs, zero where candidates run out
};

struct ExtendCands {  // extend (extend.cpp:52-136): sample s < nr is a reference, s >= nr a query
    const int64_t *row_start;  // CSR starts of the sparse ref-ref list (nr + 1)
    const int64_t *sp_j;
    const float *sp_d;
    const float *qq, *qr;      // (nq x nq), (nr x nq)
    int64_t nr, nq;
    const int64_t *out_start;  // exclusive scan of the per-sample output counts
    __device__ __forceinline__ int64_t length(int64_t s) const {
        return nq + (s < nr ? row_start[s + 1] - row_start[s] : nr);
    }
    // positions 0..nq-1: the dense query part (wins ties, as the reference's merge does); then the ref part
    __device__ __forceinline__ float dist(int64_t s, int64_t pos) const {
        if (pos < nq) return s < nr ? qr[s * nq + pos] : qq[(s - nr) * nq + pos];
        pos -= nq;
        return s < nr ? sp_d[row_start[s] + pos] : qr[pos * nq + (s - nr)];
    }
    __device__ __forceinline__ int64_t j_of(int64_t s, int64_t pos) const {
        if (pos < nq) return nr + pos;
        pos -= nq;
        return s < nr ? sp_j[row_start[s] + pos] : pos;
    }
    __device__ __forceinline__ int64_t out_pos(int64_t s, int32_t) const { return out_start[s]; }
    __device__ __forceinline__ bool pad() const { return false; }
};

// (Staging each row in shared memory with one TMA bulk copy, so that the four re-reads do not go to L2, was measured
//  and lost: 8.2 ms vs 5.5 ms at 28 k x 28 k — a 113 KB row leaves one 256-thread CTA per SM, too few warps for the
//  shared-memory atomics of the histogram passes.  kNN <= 32 takes knn_small_kernel (one pass) instead.)
template <typename Cands>
__global__ void __launch_bounds__(kKnnThreads) knn_kernel(Cands c, int64_t n_rows, int32_t knn, int64_t *__restrict__ out_i,
                                                          int64_t *__restrict__ out_j, float *__restrict__ out_d) {
    __shared__ uint32_t hist[kKnnThreads / 32][256];  // one histogram per warp: 8x less contention on hot buckets
    __shared__ unsigned long long sel[kKnnMax];  // (key << 32 | position): positions < 2^32
    __shared__ uint32_t warp_cnt[kKnnThreads / 32];
    __shared__ uint32_t s_prefix, s_need, s_below, s_taken_eq, s_n_sel, s_eq_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kBatch = 4;  // loads of a batch are issued together, before the first shared-memory atomic
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t len = c.length(r);
        auto dist_at = [&](int64_t pos) -> float { return c.dist(r, pos); };
        // ---- (1) radix select of the knn-th smallest key among candidates with j != r
        uint32_t prefix = 0, need = (uint32_t)knn, below = 0;  // keys matching `prefix` in the decided bytes
        bool enough = true;
        for (int byte = 3; byte >= 0; byte--) {
            for (int b = threadIdx.x; b < 256 * (kKnnThreads / 32); b += kKnnThreads) (&hist[0][0])[b] = 0;
            __syncthreads();
            const uint32_t decided = byte == 3 ? 0u : (0xffffffffu << (8 * (byte + 1)));
            for (int64_t p0 = threadIdx.x; p0 < len; p0 += (int64_t)kBatch * kKnnThreads) {
                uint32_t key[kBatch];
                bool use[kBatch];
#pragma unroll
                for (int q = 0; q < kBatch; q++) {
                    const int64_t pos = p0 + (int64_t)q * kKnnThreads;
                    use[q] = pos < len && c.j_of(r, pos) != r;
                    key[q] = float_key(dist_at(min(pos, len - 1)));
                }
#pragma unroll
                for (int q = 0; q < kBatch; q++)
                    if (use[q] && (key[q] & decided) == (prefix & decided)) atomicAdd(&hist[warp][(key[q] >> (8 * byte)) & 255u], 1u);
            }
            __syncthreads();
            {   // fold the per-warp histograms into hist[0]
                uint32_t t = 0;
#pragma unroll
                for (int w = 0; w < kKnnThreads / 32; w++) t += hist[w][threadIdx.x];
                __syncthreads();
                hist[0][threadIdx.x] = t;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t acc = 0, b = 0;
                for (; b < 256; b++) {
                    if (acc + hist[0][b] >= need) break;
                    acc += hist[0][b];
                }
                if (b == 256) {  // fewer than `need` candidates in total: take everything
                    s_need = 0xffffffffu;
                } else {
                    s_need = need - acc;
                    s_prefix = prefix | (b << (8 * byte));
                    s_below = below + acc;
                    s_eq_total = hist[0][b];  // after the last byte: how many keys equal the threshold key
                }
            }
            __syncthreads();
            if (s_need == 0xffffffffu) {
                enough = false;
                break;
            }
            need = s_need;
            prefix = s_prefix;
            below = s_below;
            __syncthreads();
        }
        // enough: T = prefix, `below` keys < T, take `need` of the keys == T in position order.
        // not enough: every candidate is taken.
        if (threadIdx.x == 0) {
            s_n_sel = 0;
            s_taken_eq = 0;
        }
        __syncthreads();
        // ---- (2) collect.  Usually every key equal to the threshold is needed (no tie is cut), or everything is
        // taken: then order does not matter before the sort and one batched pass with a slot counter does it.  Only
        // when the threshold cuts THROUGH a run of ties are the first `need` of them, in position order, picked with
        // the block-ordered ballot scan (three barriers per 256 candidates).
        if (!enough || s_eq_total == need) {
            for (int64_t p0 = threadIdx.x; p0 < len; p0 += (int64_t)kBatch * kKnnThreads) {
                uint32_t key[kBatch];
                bool use[kBatch];
#pragma unroll
                for (int q = 0; q < kBatch; q++) {
                    const int64_t pos = p0 + (int64_t)q * kKnnThreads;
                    use[q] = pos < len && c.j_of(r, pos) != r;
                    key[q] = float_key(dist_at(min(pos, len - 1)));
                }
#pragma unroll
                for (int q = 0; q < kBatch; q++)
                    if (use[q] && (!enough || key[q] <= prefix)) {
                        const uint32_t at = atomicAdd(&s_n_sel, 1u);
                        if (at < kKnnMax) sel[at] = ((unsigned long long)key[q] << 32) | (unsigned long long)(uint32_t)(p0 + (int64_t)q * kKnnThreads);
                    }
            }
            __syncthreads();
        } else {
            for (int64_t start = 0; start < len; start += kKnnThreads) {
                const int64_t pos = start + threadIdx.x;
                bool lt = false, eq = false;
                uint32_t key = 0;
                if (pos < len && c.j_of(r, pos) != r) {
                    key = float_key(dist_at(pos));
                    lt = key < prefix;
                    eq = key == prefix;
                }
                const uint32_t beq = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) warp_cnt[warp] = __popc(beq);
                __syncthreads();
                uint32_t before = s_taken_eq;
                for (int w = 0; w < warp; w++) before += warp_cnt[w];
                const uint32_t rank = before + __popc(beq & ((1u << lane) - 1));
                if (lt || (eq && rank < need)) {
                    const uint32_t at = atomicAdd(&s_n_sel, 1u);
                    if (at < kKnnMax) sel[at] = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)pos;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    uint32_t t = 0;
                    for (int w = 0; w < kKnnThreads / 32; w++) t += warp_cnt[w];
                    s_taken_eq += t;
                }
                __syncthreads();
            }
        }
        const uint32_t n_sel = min(s_n_sel, (uint32_t)knn);
        // ---- (3) bitonic sort of sel[0 .. n_pow2)
        uint32_t n_pow2 = 1;
        while (n_pow2 < n_sel) n_pow2 <<= 1;
        for (uint32_t t = n_sel + threadIdx.x; t < n_pow2; t += kKnnThreads) sel[t] = ~0ull;
        __syncthreads();
        for (uint32_t k2 = 2; k2 <= n_pow2; k2 <<= 1)
            for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                for (uint32_t t = threadIdx.x; t < n_pow2; t += kKnnThreads) {
                    const uint32_t partner = t ^ j2;
                    if (partner > t) {
                        const unsigned long long a = sel[t], b = sel[partner];
                        const bool up = (t & k2) == 0;
                        if ((a > b) == up) {
                            sel[t] = b;
                            sel[partner] = a;
                        }
                    }
                }
                __syncthreads();
            }
        // ---- (4) write
        const int64_t o0 = c.out_pos(r, knn);
        for (uint32_t t = threadIdx.x; t < (c.pad() ? (uint32_t)knn : n_sel); t += kKnnThreads) {
            int64_t j = 0;
            float dv = 0.0f;
            if (t < n_sel) {
                const int64_t pos = (int64_t)(uint32_t)sel[t];
                j = c.j_of(r, pos);
                dv = dist_at(pos);
            }
            out_i[o0 + t] = r;
            out_j[o0 + t] = j;
            out_d[o0 + t] = dv;
        }
        __syncthreads();
    }
}

// kNN <= 32 (the lineage models' ranks: PopPUNK/models.py:1215-1222 asks for max(ranks), a handful): ONE pass over the
// candidates.  Each of the 8 warps streams every 8th batch of the row and keeps its kNN smallest (key, position) pairs
// sorted across its lanes; a candidate is inserted only if it beats the warp's current kNN-th (after the first few
// batches almost none does), so the pass is a coalesced read and a compare.  The warps' lists meet in shared memory,
// are sorted, and the first kNN are written.  (key << 32 | position) orders ties by position, as the reference does.
template <typename Cands>
__global__ void __launch_bounds__(256) knn_small_kernel(Cands c, int64_t n_rows, int32_t knn, int64_t *__restrict__ out_i,
                                                        int64_t *__restrict__ out_j, float *__restrict__ out_d) {
    __shared__ unsigned long long sel[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kU = 8;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t len = c.length(r);
        unsigned long long best = ~0ull;  // lane l: the l-th smallest pair this warp has seen
        for (int64_t b0 = (int64_t)warp * (32 * kU); b0 < len; b0 += 8 * 32 * kU) {
            float v[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) v[u] = c.dist(r, min(b0 + u * 32 + lane, len - 1));
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int64_t pos = b0 + u * 32 + lane;
                const bool ok = pos < len && c.j_of(r, pos) != r;
                const unsigned long long x = ok ? (((unsigned long long)float_key(v[u]) << 32) | (unsigned long long)(uint32_t)pos) : ~0ull;
                unsigned m = __ballot_sync(0xffffffffu, x < __shfl_sync(0xffffffffu, best, knn - 1));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long xs = __shfl_sync(0xffffffffu, x, src);
                    if (xs < __shfl_sync(0xffffffffu, best, knn - 1)) {   // (uniform) still among the kNN smallest
                        const int at = __ffs(__ballot_sync(0xffffffffu, best > xs)) - 1;
                        const unsigned long long up = __shfl_up_sync(0xffffffffu, best, 1);
                        if (lane > at) best = up;
                        else if (lane == at) best = xs;
                    }
                }
            }
        }
        sel[threadIdx.x] = lane < knn ? best : ~0ull;
        __syncthreads();
        for (uint32_t k2 = 2; k2 <= 256; k2 <<= 1)
            for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                const uint32_t t = threadIdx.x, partner = t ^ j2;
                if (partner > t) {
                    const unsigned long long a = sel[t], b = sel[partner];
                    const bool up = (t & k2) == 0;
                    if ((a > b) == up) {
                        sel[t] = b;
                        sel[partner] = a;
                    }
                }
                __syncthreads();
            }
        const int64_t o0 = c.out_pos(r, knn);
        if ((int)threadIdx.x < knn) {
            const unsigned long long e = sel[threadIdx.x];
            if (e != ~0ull) {
                const int64_t pos = (int64_t)(uint32_t)e;
                out_i[o0 + threadIdx.x] = r;
                out_j[o0 + threadIdx.x] = c.j_of(r, pos);
                out_d[o0 + threadIdx.x] = c.dist(r, pos);
            } else if (c.pad()) {  // fewer candidates than kNN: zero padded (extend.cpp:262-274 sizes rows*kNN)
                out_i[o0 + threadIdx.x] = r;
                out_j[o0 + threadIdx.x] = 0;
                out_d[o0 + threadIdx.x] = 0.0f;
            }
        }
        __syncthreads();
    }
}

// CSR starts of a COO list sorted by i: start[r] = first p with i[p] >= r (extend.cpp:14-38), start[n] = nnz
__global__ void row_starts_kernel(const int64_t *__restrict__ i_vec, int64_t nnz, int64_t n, int64_t *__restrict__ start) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (i_vec[mid] < r) lo = mid + 1; else hi = mid;
        }
        start[r] = r == n ? nnz : lo;
    }
}

// extend: how many entries sample s will emit = min(kNN, candidates that are not s itself)
__global__ void extend_count_kernel(ExtendCands c, int64_t n_total, int32_t knn, int64_t *__restrict__ cnt) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_total; s += (int64_t)gridDim.x * blockDim.x) {
        int64_t avail;
        if (s < c.nr) {
            avail = c.nq;
            for (int64_t p = c.row_start[s]; p < c.row_start[s + 1]; p++) avail += c.sp_j[p] != s;
        } else {
            avail = c.nr + c.nq - 1;
        }
        cnt[s] = min((int64_t)knn, avail);
    }
}

// ------------------------------------------------------------------------------------------------------------
// lower_rank (extend.cpp:146-243).  One warp per sample: its (<= kLowerMaxRow) sparse neighbours are sorted by
// (distance, list position) in shared memory, lane 0 walks them with the reference's keep rule, and the kept
// entries go to a staging area at the row's own CSR offset (kept <= row length).  An optional reciprocal filter
// then keeps (i < j) entries whose mirror (j, i) was also kept; a scan of the per-row counts places the output.
// ------------------------------------------------------------------------------------------------------------
constexpr int kLowerMaxRow = 1024;
constexpr int kLowerWarps = 4;

__global__ void __launch_bounds__(kLowerWarps * 32) lower_rank_keep_kernel(
    const int64_t *__restrict__ row_start, const int64_t *__restrict__ sp_j, const float *__restrict__ sp_d, int64_t n,
    int64_t knn, int32_t count_unique, float epsilon, int64_t *__restrict__ st_j, float *__restrict__ st_d,
    int64_t *__restrict__ kept, int32_t *__restrict__ too_long) {
    __shared__ unsigned long long key[kLowerWarps][kLowerMaxRow];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * kLowerWarps + warp;
    if (r >= n) return;
    const int64_t b = row_start[r];
    const int64_t len = max((int64_t)0, row_start[r + 1] - b);
    if (len > kLowerMaxRow) {
        if (lane == 0) {
            atomicExch(too_long, 1);
            kept[r] = 0;
        }
        return;
    }
    uint32_t n_pow2 = 1;
    while (n_pow2 < (uint32_t)len) n_pow2 <<= 1;
    for (uint32_t t = lane; t < n_pow2; t += 32)
        key[warp][t] = t < len ? (((unsigned long long)float_key(sp_d[b + t]) << 32) | t) : ~0ull;
    __syncwarp();
    for (uint32_t k2 = 2; k2 <= n_pow2; k2 <<= 1)
        for (uint32_t j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (uint32_t t = lane; t < n_pow2; t += 32) {
                const uint32_t partner = t ^ j2;
                if (partner > t) {
                    const unsigned long long x = key[warp][t], y = key[warp][partner];
                    if ((x > y) == ((t & k2) == 0)) {
                        key[warp][t] = y;
                        key[warp][partner] = x;
                    }
                }
            }
            __syncwarp();
        }
    if (lane == 0) {
        int64_t unique = 0, n_kept = 0;
        float prev = 0.0f;
        for (int64_t t = 0; t < len; t++) {
            const int64_t pos = (int64_t)(uint32_t)key[warp][t];
            const int64_t j = sp_j[b + pos];
            const float dv = sp_d[b + pos];
            if (j == r) continue;
            if (count_unique) {
                if (fabsf(__fsub_rn(dv, prev)) >= epsilon) {
                    unique++;
                    prev = dv;
                }
            } else {
                unique = n_kept;
            }
            if (unique > knn) break;
            st_j[b + n_kept] = j;
            st_d[b + n_kept] = dv;
            n_kept++;
        }
        kept[r] = n_kept;
    }
}

// reciprocal filter: compact each row's staged entries to those with i < j whose mirror was kept; in place per row
// is unsafe (other rows read this row), so the survivors are flagged in `keep_flag` and counted in `cnt`.
__global__ void lower_rank_reciprocal_kernel(const int64_t *__restrict__ row_start, const int64_t *__restrict__ st_j,
                                             const int64_t *__restrict__ kept, int64_t n, uint8_t *__restrict__ keep_flag,
                                             int64_t *__restrict__ cnt) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = row_start[r];
        int64_t c = 0;
        for (int64_t t = 0; t < kept[r]; t++) {
            const int64_t j = st_j[b + t];
            bool ok = false;
            if (r < j && j < n) {
                const int64_t bj = row_start[j];
                for (int64_t u = 0; u < kept[j] && !ok; u++) ok = st_j[bj + u] == r;
            }
            keep_flag[b + t] = ok;
            c += ok;
        }
        cnt[r] = c;
    }
}

__global__ void lower_rank_write_kernel(const int64_t *__restrict__ row_start, const int64_t *__restrict__ st_j,
                                        const float *__restrict__ st_d, const int64_t *__restrict__ kept,
                                        const uint8_t *__restrict__ keep_flag, const int64_t *__restrict__ out_start,
                                        int64_t n, int64_t *__restrict__ out_i, int64_t *__restrict__ out_j,
                                        float *__restrict__ out_d) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = row_start[r];
        int64_t o = out_start[r];
        for (int64_t t = 0; t < kept[r]; t++) {
            if (keep_flag && !keep_flag[b + t]) continue;
            out_i[o] = r;
            out_j[o] = st_j[b + t];
            out_d[o] = st_d[b + t];
            o++;
        }
    }
}

}  // namespace ppb
