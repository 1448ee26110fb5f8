// "Next" rows of SURVEY.md section 8(f): the immediate consumers of the distance path.
//
//   N1  edge lists after the threshold (src/boundary.cpp:82-123: edge_iterate, generate_tuples):
//       ordered stream compaction of the rows that pass a predicate, mapped to (i, j) sample pairs
//   N2  long <-> square reshapes of the path's output (pp_sketchlib.longToSquare / squareToLong /
//       longToSquareMulti; call sites PopPUNK/utils.py:393-405, network.py:2133-2134, models.py:1217,1357)
//
// All memory-bound index work: coalesced loads, warp ballots for ranking, no atomics on the ordered path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ppb_kernels.cuh"

namespace ppb {

// ---- row -> (i, j): src/boundary.cpp:22-31 with an exact integer fix-up of the double sqrt ---------------
__device__ __forceinline__ int64_t dev_sq2cond(int64_t i, int64_t j, int64_t n) {
    return n * i - ((i * (i + 1)) >> 1) + j - 1 - i;
}
__device__ __forceinline__ int64_t dev_row_idx(int64_t k, int64_t n) {
    const double d = sqrt((double)(-8 * k + 4 * n * (n - 1) - 7));
    int64_t i = n - 2 - (int64_t)floor(d / 2.0 - 0.5);
    i = max((int64_t)0, min(i, n - 2));
    while (i > 0 && dev_sq2cond(i, i + 1, n) > k) i--;
    while (i < n - 2 && dev_sq2cond(i + 1, i + 2, n) <= k) i++;
    return i;
}
__device__ __forceinline__ int64_t dev_col_idx(int64_t k, int64_t i, int64_t n) {
    return k + i + 1 - n * (n - 1) / 2 + (n - i) * ((n - i) - 1) / 2;
}

struct PairMap {
    int32_t self;
    int64_t n;           // self: number of samples; non-self: num_ref
    int64_t int_offset;  // boundary.cpp:97-123 generate_tuples
};
__device__ __forceinline__ void row_to_pair(const PairMap &m, int64_t row, int64_t &i, int64_t &j) {
    if (m.self) {
        i = dev_row_idx(row, m.n);
        j = dev_col_idx(row, i, m.n) + m.int_offset;
        i += m.int_offset;
    } else {  // boundary.cpp:112-113
        i = row % m.n + m.int_offset;
        j = row / m.n + m.n + m.int_offset;
    }
    if (i > j) {
        const int64_t t = i;
        i = j;
        j = t;
    }
}

// ---- first admitting boundary of a row ------------------------------------------------------------------------
// threshold_iterate_1D / 2D test a row against 20-40 boundaries that move outward (src/boundary.cpp:171-185, 211-237;
// PopPUNK/refine.py:116-123,190-191).  In exact arithmetic a row's test flips once, from outside to inside, so the
// first admitting boundary can be found by bisection — PROVIDED the float32 line_dist has the true sign at every
// boundary.  That is certain wherever |line_dist| exceeds twice a bound on its rounding error taken over ALL
// boundaries: |fl(fl(fl(y0 xm) + fl(x0 ym)) - fl(xm ym)) - exact| <= 4 u (|y0| XM + |x0| YM + XM YM), u = 2^-24
// (slopes 0 and 1 are a single subtraction, whose sign is always exact).  A probe inside that band, a degenerate
// boundary (an intercept of 0: the sqrt form) or boundaries that do not move outward send the row to the full scan.
struct StepSearch {
    const float2 *step;   // device: (x_max[o], y_max[o])
    int32_t n_off, slope;
    int32_t bisect;       // host-verified: intercepts positive and non-decreasing in o
    float XM, YM;         // max |x_max|, max |y_max|
};
// returns the first o with line_dist_o(v) <= 0 (n_off: none), or -1 when the row needs the full scan
template <int kSlope>
__device__ __forceinline__ int32_t bisect_steps(const StepSearch &S, const float2 v) {
    // bisection is only allowed without degenerate boundaries, so slope 2 is line_dist's three-product form
    // (boundary.cpp:48-50, same operations in the same order), slopes 0 / 1 one subtraction (:51-54)
    const float guard = kSlope == 2 ? 9.6e-7f * (fabsf(v.y) * S.XM + fabsf(v.x) * S.YM + S.XM * S.YM) : -1.0f;  // 2 * 8 u * (...)
    auto side_at = [&](int o) -> float {
        const float2 b = __ldg(S.step + o);
        if (kSlope == 2) return __fsub_rn(__fadd_rn(__fmul_rn(v.y, b.x), __fmul_rn(v.x, b.y)), __fmul_rn(b.x, b.y));
        return kSlope == 0 ? __fsub_rn(v.x, b.x) : __fsub_rn(v.y, b.y);
    };
    float s = side_at(S.n_off - 1);
    if (!(fabsf(s) > guard)) return -1;
    if (s > 0.0f) return S.n_off;
    s = side_at(0);
    if (!(fabsf(s) > guard)) return -1;
    if (s <= 0.0f) return 0;
    int lo = 0, hi = S.n_off - 1;  // outside at lo, inside at hi
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        s = side_at(mid);
        if (!(fabsf(s) > guard)) return -1;
        if (s <= 0.0f) hi = mid; else lo = mid;
    }
    return hi;
}
__device__ __noinline__ int32_t first_admitting_step(const StepSearch &S, const float2 v) {   // (one copy: callers unroll 16x)
    if (!S.bisect) return -1;
    if (S.slope == 2) return bisect_steps<2>(S, v);
    return S.slope == 0 ? bisect_steps<0>(S, v) : bisect_steps<1>(S, v);
}

// ---- predicates -------------------------------------------------------------------------------------------
// A predicate is split into load(row) and test(value) so that a thread can put all its loads in flight before the
// first ballot (a ballot is a convergence point: written as one call, every load would wait for the previous vote).
struct PredDists {  // edge_iterate: line_dist(...) <= 0   (boundary.cpp:82-95)
    typedef float2 Value;
    const float2 *d;
    int32_t slope;
    float x_max, y_max;
    __device__ __forceinline__ float2 load(int64_t row) const { return __ldg(d + row); }
    __device__ __forceinline__ bool test(const float2 &v) const { return line_dist(v.x, v.y, x_max, y_max, slope) <= 0.0f; }
};
template <typename T>
struct PredLabels {  // generate_tuples: assignments[row] == within_label   (boundary.cpp:106)
    typedef T Value;
    const T *labels;
    int32_t within;
    __device__ __forceinline__ T load(int64_t row) const { return __ldg(labels + row); }
    __device__ __forceinline__ bool test(const T &v) const { return (int32_t)v == within; }
};

// ---- ordered compaction -----------------------------------------------------------------------------------
// A block owns kSelBlockRows consecutive rows; a warp owns kSelIters runs of 32 consecutive rows, so one ballot
// per run already is the order.  Pass 0 writes the block's count; after an exclusive scan of the counts
// (scan_kernel) pass 1 writes the pairs at their final position.  The predicate is evaluated once per pass.
constexpr int kSelThreads = 256;
constexpr int kSelIters = 16;
constexpr int kSelBlockRows = kSelThreads * kSelIters;  // 4096

template <typename Pred>
__global__ void __launch_bounds__(kSelThreads) select_kernel(Pred pred, int64_t n_rows, PairMap map, int pass,
                                                             int64_t *__restrict__ block_off, int64_t capacity,
                                                             int64_t *__restrict__ out_i, int64_t *__restrict__ out_j) {
    __shared__ uint32_t warp_tot[kSelThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * kSelBlockRows + (int64_t)warp * (32 * kSelIters);
    uint32_t ballots[kSelIters];
    uint32_t tot = 0;
    typename Pred::Value vals[kSelIters];
#pragma unroll
    for (int m = 0; m < kSelIters; m++)  // 16 independent coalesced loads in flight per thread (clamped, branch-free)
        vals[m] = pred.load(min(base + m * 32 + lane, n_rows - 1));
#pragma unroll
    for (int m = 0; m < kSelIters; m++) {
        const bool sel = base + m * 32 + lane < n_rows && pred.test(vals[m]);
        ballots[m] = __ballot_sync(0xffffffffu, sel);
        tot += __popc(ballots[m]);
    }
    if (lane == 0) warp_tot[warp] = tot;
    __syncthreads();
    uint32_t before = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; w++) {
        before += w < warp ? warp_tot[w] : 0;
        block_total += warp_tot[w];
    }
    if (pass == 0) {
        if (threadIdx.x == 0) block_off[blockIdx.x] = block_total;
        return;
    }
    int64_t pos = block_off[blockIdx.x] + before;
#pragma unroll
    for (int m = 0; m < kSelIters; m++) {
        const uint32_t b = ballots[m];
        if (b & (1u << lane)) {
            const int64_t at = pos + __popc(b & ((1u << lane) - 1));
            if (at < capacity) {
                int64_t i, j;
                row_to_pair(map, base + m * 32 + lane, i, j);
                out_i[at] = i;
                out_j[at] = j;
            }
        }
        pos += __popc(b);
    }
}

// exclusive scan of int64 counts in place (one CTA; the array is n_rows/4096 long), total -> *d_total
__global__ void __launch_bounds__(1024) scan_kernel(int64_t *__restrict__ a, int64_t n, int64_t *__restrict__ d_total) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < n; start += 1024) {
        const int64_t idx = start + threadIdx.x;
        const int64_t v = idx < n ? a[idx] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t incl = x + (warp ? warp_sum[warp - 1] : 0) + carry;
        if (idx < n) a[idx] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_total = carry_s;
}

__global__ void rows_to_pairs_kernel(const int64_t *__restrict__ rows, int64_t n, PairMap map,
                                     int64_t *__restrict__ out_i, int64_t *__restrict__ out_j) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t i, j;
        row_to_pair(map, rows[r], i, j);
        out_i[r] = i;
        out_j[r] = j;
    }
}

// ---- N2: long <-> square -----------------------------------------------------------------------------------
// vec is read with an element stride (PopPUNK passes column views distMat[:, [c]] of the (n_pairs, 2) array).
// A condensed row is contiguous in j, so the upper triangle is read and written in coalesced runs; the lower
// triangle is its transpose, produced through a 32 x 33 shared-memory tile so that it is written in coalesced runs
// too (a direct gather would fetch one 32-byte sector per element).  Grid: 32 x 32 tiles with tile_c >= tile_r.
constexpr int kSqTile = 32;

__global__ void __launch_bounds__(kSqTile * 8) long_to_square_kernel(const float *__restrict__ vec, int64_t stride, int64_t n,
                                                                     float *__restrict__ sq) {
    __shared__ float tile[kSqTile][kSqTile + 1];
    const int64_t nt = (n + kSqTile - 1) / kSqTile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads, 4 rows each
    for (int64_t t = blockIdx.x; t < nt * (nt + 1) / 2; t += gridDim.x) {
        // t -> (tr, tc) with tc >= tr: row tr of the upper-triangular tile grid starts at tr*nt - tr(tr-1)/2
        int64_t tr = (int64_t)((2.0 * (double)nt + 1.0 - sqrt((2.0 * (double)nt + 1.0) * (2.0 * (double)nt + 1.0) - 8.0 * (double)t)) * 0.5);
        while (tr > 0 && tr * nt - tr * (tr - 1) / 2 > t) tr--;
        while ((tr + 1) * nt - (tr + 1) * tr / 2 <= t) tr++;
        const int64_t tc = tr + (t - (tr * nt - tr * (tr - 1) / 2));
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = tr * kSqTile + ty * 4 + k, c = tc * kSqTile + tx;
            float v = 0.0f;
            if (r < n && c < n && r != c) v = r < c ? vec[dev_sq2cond(r, c, n) * stride] : 0.0f;
            if (r < n && c < n) {
                if (r < c) sq[r * n + c] = v;
                else if (r == c) sq[r * n + c] = 0.0f;
            }
            tile[ty * 4 + k][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {  // transposed tile: element (c, r) of the square, c in this tile's columns
            const int64_t c = tc * kSqTile + ty * 4 + k, r = tr * kSqTile + tx;
            if (c < n && r < n && r < c) sq[c * n + r] = tile[tx][ty * 4 + k];
        }
    }
}
// one block per square row i: its j > i entries are one contiguous run of the condensed vector
__global__ void square_to_long_kernel(const float *__restrict__ sq, int64_t n, float *__restrict__ vec) {
    for (int64_t i = blockIdx.x; i < n - 1; i += gridDim.x) {
        const int64_t base = dev_sq2cond(i, i + 1, n) - (i + 1);
        for (int64_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x) vec[base + j] = sq[i * n + j];
    }
}
// (R+Q)^2 square from: condensed ref-ref, query-major query-ref rectangle, condensed query-query
__global__ void long_to_square_multi_kernel(const float *__restrict__ rr, int64_t s_rr, const float *__restrict__ qr,
                                            int64_t s_qr, const float *__restrict__ qq, int64_t s_qq, int64_t R,
                                            int64_t Q, float *__restrict__ sq) {
    const int64_t n = R + Q, total = n * n;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = o / n, c = o - r * n;
        const int64_t lo = min(r, c), hi = max(r, c);
        float v = 0.0f;
        if (lo != hi) {
            if (hi < R)
                v = rr[dev_sq2cond(lo, hi, R) * s_rr];
            else if (lo >= R)
                v = qq[dev_sq2cond(lo - R, hi - R, Q) * s_qq];
            else
                v = qr[((hi - R) * R + lo) * s_qr];  // row = q*R + r (utils.py:224-226)
        }
        sq[o] = v;
    }
}

}  // namespace ppb
