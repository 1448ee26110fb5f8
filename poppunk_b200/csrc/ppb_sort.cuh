// Ordered stream primitives for the consumers of the distance array, hand-written:
//
//   select_mark/emit      ordered compaction that reads its input once (bit mask + unit counts, scan, emit from the
//                         mask): edge_iterate / generate_tuples (src/boundary.cpp:82-123)
//   bucket_*_kernel       stable scatter of items into <= 256 buckets (count per chunk -> scans -> emit at the final
//                         position, ranks inside a chunk from warp match masks).  Used as
//                           - threshold_iterate_2D (src/boundary.cpp:211-237): bucket = admitting step; the classify
//                             pass reads the distances ONCE and leaves one byte per row for the emit pass
//                           - the passes of a least-significant-digit radix sort (8-bit digits): the stable sort of
//                             threshold_iterate_1D (boundary.cpp:190, sort_indexes) and the row order of the fused
//                             edge list — no library sort anywhere on the path
//
// All of it is memory-bound index work: coalesced 16-row-per-thread loads, no global atomics on the ordered paths.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ppb_next.cuh"

namespace ppb {

// ------------------------------------------------------------------------------------------------------------
// ordered compaction that reads its input ONCE:  mark (streaming: one bit per row + a count per 1024-row unit)
// -> exclusive scan of the unit counts -> emit (reads only the bit mask: n_rows / 8 bytes).
// (A single-pass variant with decoupled look-back between 4096-row chunks was measured first: 4.4 ms for 400 M rows —
//  with ~450 chunks in flight every chunk walks ~14 look-back windows of ~1 us, which is longer than its own loads.)
// ------------------------------------------------------------------------------------------------------------
constexpr int kUnitRows = 1024;  // one warp-iteration: 32 ballots of 32 rows

template <typename Pred>
__global__ void __launch_bounds__(256) select_mark_kernel(Pred pred, int64_t n_rows, uint32_t *__restrict__ bits,
                                                          uint32_t *__restrict__ unit_count) {
    const int lane = threadIdx.x & 31;
    const int64_t n_units = (n_rows + kUnitRows - 1) / kUnitRows;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp0; u < n_units; u += n_warps) {
        const int64_t base = u * kUnitRows;
        uint32_t mine = 0;  // lane m keeps the ballot of rows base + 32 m .. + 31
#pragma unroll
        for (int half = 0; half < 2; half++) {
            typename Pred::Value vals[16];
#pragma unroll
            for (int m = 0; m < 16; m++) vals[m] = pred.load(min(base + (half * 16 + m) * 32 + lane, n_rows - 1));
#pragma unroll
            for (int m = 0; m < 16; m++) {
                const int mm = half * 16 + m;
                const uint32_t b = __ballot_sync(0xffffffffu, base + mm * 32 + lane < n_rows && pred.test(vals[m]));
                if (lane == mm) mine = b;
            }
        }
        bits[u * 32 + lane] = mine;                                   // 128 coalesced bytes per unit
        const uint32_t tot = __reduce_add_sync(0xffffffffu, __popc(mine));
        if (lane == 0) unit_count[u] = tot;
    }
}

// exclusive scan of uint32 unit counts into int64 offsets (one CTA, 8 entries per thread per round), total -> *d_total
__global__ void __launch_bounds__(1024) unit_scan_kernel(const uint32_t *__restrict__ cnt, int64_t n, int64_t *__restrict__ off,
                                                         int64_t *__restrict__ d_total) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < n; start += 8192) {
        const int64_t i0 = start + (int64_t)threadIdx.x * 8;
        uint32_t v[8];
        int64_t mine = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            v[q] = i0 + q < n ? cnt[i0 + q] : 0u;
            mine += v[q];
        }
        int64_t x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t s2 = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, s2, o);
                if (lane >= o) s2 += y;
            }
            warp_sum[lane] = s2;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        int64_t run = x - mine + (warp ? warp_sum[warp - 1] : 0) + carry;  // exclusive prefix of this thread's 8 entries
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (i0 + q < n) off[i0 + q] = run;
            run += v[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_total = carry_s;
}

struct OutPairs {  // (i, j) sample pairs of the selected rows (generate_tuples' conventions)
    PairMap map;
    int64_t *out_i, *out_j;
    __device__ __forceinline__ void write(int64_t at, int64_t row) const {
        int64_t i, j;
        row_to_pair(map, row, i, j);
        out_i[at] = i;
        out_j[at] = j;
    }
};

template <typename Out>
__global__ void __launch_bounds__(256) select_emit_kernel(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ unit_count,
                                                          const int64_t *__restrict__ unit_off, int64_t n_rows, Out out,
                                                          int64_t capacity) {
    const int lane = threadIdx.x & 31;
    const int64_t n_units = (n_rows + kUnitRows - 1) / kUnitRows;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp0; u < n_units; u += n_warps) {
        if (unit_count[u] == 0) continue;                             // (uniform across the warp)
        uint32_t w = bits[u * 32 + lane];
        uint32_t incl = __popc(w);                                    // rows selected in words 0..lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t before = incl - __popc(w);                     // ... in words 0..lane-1
        const int64_t off = unit_off[u];
        if (unit_count[u] >= 96) {
            // dense unit: the warp walks the non-empty words together — lane l takes row 32 m + l of word m, so the rows
            // read and the positions written by one step are consecutive (a lane walking its own word writes 32 runs)
            uint32_t nz = __ballot_sync(0xffffffffu, w != 0);
            while (nz) {
                const int m = __ffs(nz) - 1;
                nz &= nz - 1;
                const uint32_t wm = __shfl_sync(0xffffffffu, w, m), bm = __shfl_sync(0xffffffffu, before, m);
                if (wm & (1u << lane)) {
                    const int64_t at = off + bm + __popc(wm & ((1u << lane) - 1));
                    if (at < capacity) out.write(at, u * kUnitRows + m * 32 + lane);
                }
            }
        } else {
            int64_t at = off + before;
            const int64_t row0 = u * kUnitRows + lane * 32;
            while (w) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                if (at < capacity) out.write(at, row0 + bit);
                at++;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// stable bucket scatter
// ------------------------------------------------------------------------------------------------------------
constexpr int kBktMax = 256;
constexpr uint32_t kNoBucket = 0xffffffffu;

// A classifier says which bucket(s) an item belongs to:
//   Value load(item)                                  the item's input (all 16 loads of a thread are issued together)
//   kMulti == false : uint32_t classify(item, value) / recall(item)   -> bucket or kNoBucket
//   kMulti == true  : uint64_t classify(item, value) / recall(item)   -> bit b set = bucket b (<= 64 buckets)
// classify() runs in the counting pass (and may leave notes for recall(), which runs in the emit pass).
template <typename C>
__global__ void __launch_bounds__(kSelThreads) bucket_count_kernel(C cls, int64_t n, int32_t n_buckets,
                                                                   int64_t *__restrict__ hist) {
    __shared__ uint32_t cnt[kBktMax];
    for (int b = threadIdx.x; b < kBktMax; b += kSelThreads) cnt[b] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * kSelBlockRows + (int64_t)warp * (32 * kSelIters);
    typename C::Value vals[kSelIters];
#pragma unroll
    for (int m = 0; m < kSelIters; m++) vals[m] = cls.load(min(base + m * 32 + lane, n - 1));  // 16 loads in flight
#pragma unroll
    for (int m = 0; m < kSelIters; m++) {
        const int64_t item = base + m * 32 + lane;
        if (item < n) {
            if constexpr (C::kMulti) {
                uint64_t mask = cls.classify(item, vals[m]);
                while (mask) {
                    atomicAdd(&cnt[__ffsll((long long)mask) - 1], 1u);
                    mask &= mask - 1;
                }
            } else {
                const uint32_t code = cls.classify(item, vals[m]);
                if (code != kNoBucket) atomicAdd(&cnt[code], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_buckets; b += kSelThreads) hist[(int64_t)blockIdx.x * n_buckets + b] = cnt[b];
}

// CTA b: exclusive scan of bucket b's per-chunk counts (in place, stride n_buckets), its total -> totals[b]
__global__ void __launch_bounds__(1024) bucket_scan_chunks_kernel(int64_t *__restrict__ hist, int64_t n_chunks,
                                                                  int32_t n_buckets, int64_t *__restrict__ totals) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < n_chunks; start += 1024) {
        const int64_t idx = start + threadIdx.x;
        const int64_t v = idx < n_chunks ? hist[idx * n_buckets + b] : 0;
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
            warp_sum[lane] = s;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t incl = x + (warp ? warp_sum[warp - 1] : 0) + carry;
        if (idx < n_chunks) hist[idx * n_buckets + b] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[b] = carry_s;
}
// exclusive scan over the (<= 256) bucket totals: where each bucket starts in the output; grand total -> *d_total
__global__ void __launch_bounds__(kBktMax) bucket_scan_totals_kernel(const int64_t *__restrict__ totals, int32_t n_buckets,
                                                                     int64_t *__restrict__ bucket_base,
                                                                     int64_t *__restrict__ d_total) {
    __shared__ int64_t s[kBktMax];
    s[threadIdx.x] = (int)threadIdx.x < n_buckets ? totals[threadIdx.x] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int b = 0; b < n_buckets; b++) {
            const int64_t t = s[b];
            bucket_base[b] = acc;
            acc += t;
        }
        if (d_total) *d_total = acc;
    }
}

// E: void emit(item, bucket, position)
// Chunks go through a staging step: every item's (bucket, item) is first put at its
// chunk-local sorted slot in shared memory, then consecutive threads write consecutive slots — runs of one bucket go to
// consecutive positions, so the stores coalesce (written directly, a warp's 32 stores hit ~32 buckets: 15.7 ms per
// pass of 400 M pairs, all of it in the store path).
template <typename C, typename E>
__global__ void __launch_bounds__(kSelThreads) bucket_emit_kernel(C cls, E em, int64_t n, int32_t n_buckets,
                                                                  const int64_t *__restrict__ hist,
                                                                  const int64_t *__restrict__ bucket_base, int64_t capacity) {
    constexpr uint32_t kNone16 = 0xFFFFu, kSeveral16 = 0xFFFEu;
    __shared__ uint32_t wcnt[kSelThreads / 32][kBktMax];
    __shared__ uint32_t run[kSelThreads / 32][kBktMax];   // items of bucket b placed so far, counted from the chunk's first
    __shared__ uint32_t staged[kSelBlockRows];            // (bucket << 16 | item - chunk base) at its local slot
    __shared__ uint16_t scode[kSelBlockRows];             // every item's bucket (kNone16 / kSeveral16)
    __shared__ uint32_t lstart[kBktMax];                  // first local slot of each bucket
    __shared__ int64_t gstart[kBktMax];                   // final position of that slot
    __shared__ uint32_t scan_w[kSelThreads / 32];
    __shared__ uint32_t s_total;
    for (int b = threadIdx.x; b < (kSelThreads / 32) * kBktMax; b += kSelThreads) (&wcnt[0][0])[b] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk0 = (int64_t)blockIdx.x * kSelBlockRows;
    const int64_t base = chunk0 + (int64_t)warp * (32 * kSelIters);
    // the buckets go to shared memory (two bytes per item), so the placement loop below needs no unrolling — with the
    // codes in registers the 16x unrolled body was instruction-cache bound
    bool any_several = false;
#pragma unroll
    for (int m = 0; m < kSelIters; m++) {
        const int64_t item = base + m * 32 + lane;
        uint32_t code = kNone16;
        if constexpr (C::kMulti) {
            uint64_t mask = item < n ? cls.recall(item) : 0ull;
            const bool several = (mask & (mask - 1)) != 0;
            any_several |= several;
            if (mask) code = several ? kSeveral16 : (uint32_t)(__ffsll((long long)mask) - 1);
            while (mask) {
                atomicAdd(&wcnt[warp][__ffsll((long long)mask) - 1], 1u);
                mask &= mask - 1;
            }
        } else {
            const uint32_t c = item < n ? cls.recall(item) : kNoBucket;
            if (c != kNoBucket) {
                code = c;
                atomicAdd(&wcnt[warp][c], 1u);
            }
        }
        scode[warp * (32 * kSelIters) + m * 32 + lane] = (uint16_t)code;
    }
    // a chunk with an item in several buckets (threshold_iterate_2D, float rounding at a boundary) is written directly
    const bool kStaged = !__syncthreads_or(any_several);
    uint32_t my_total = 0;
    for (int b = threadIdx.x; b < n_buckets; b += kSelThreads) {  // where warp w's items of bucket b start
        gstart[b] = bucket_base[b] + hist[(int64_t)blockIdx.x * n_buckets + b];
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kSelThreads / 32; w++) {
            run[w][b] = acc;
            acc += wcnt[w][b];
        }
        my_total = acc;
    }
    {   // exclusive scan of the chunk's bucket totals (bucket b = thread b; n_buckets <= 256 = block size)
        uint32_t x = my_total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) scan_w[warp] = x;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += scan_w[w];
        if ((int)threadIdx.x < n_buckets) lstart[threadIdx.x] = before + x - my_total;
        if (threadIdx.x == kSelThreads - 1) s_total = before + x;
    }
    __syncthreads();
    // items are visited warp-major, round-major, lane-major = item order; peers of a round share a bucket
    auto place = [&](uint32_t code, int64_t item) {
        const uint32_t peers = __match_any_sync(0xffffffffu, code);
        const int leader = __ffs(peers) - 1;
        uint32_t p0 = 0;
        if (lane == leader && code != kNoBucket) {
            p0 = run[warp][code];
            run[warp][code] = p0 + __popc(peers);
        }
        p0 = __shfl_sync(0xffffffffu, p0, leader);
        __syncwarp();
        if (code != kNoBucket) {
            const uint32_t local = p0 + __popc(peers & ((1u << lane) - 1));
            if (kStaged) {
                staged[lstart[code] + local] = (code << 16) | (uint32_t)(item - chunk0);
            } else {
                const int64_t pos = gstart[code] + local;
                if (pos < capacity) em.emit(item, code, pos);
            }
        }
    };
#pragma unroll 1
    for (int m = 0; m < kSelIters; m++) {
        const int64_t item = base + m * 32 + lane;
        const uint32_t c16 = scode[warp * (32 * kSelIters) + m * 32 + lane];
        if constexpr (C::kMulti) {
            if (!__any_sync(0xffffffffu, c16 == kSeveral16)) {
                place(c16 == kNone16 ? kNoBucket : c16, item);
            } else {  // an item in several buckets: bucket by bucket keeps the item order exact
                const uint64_t mask = c16 == kSeveral16 ? cls.recall(item) : (c16 == kNone16 ? 0ull : 1ull << c16);
#pragma unroll 1
                for (int b = 0; b < n_buckets; b++) place(((mask >> b) & 1ull) ? (uint32_t)b : kNoBucket, item);
            }
        } else {
            place(c16 == kNone16 ? kNoBucket : c16, item);
        }
    }
    if (kStaged) {
        __syncthreads();
        for (uint32_t slot = threadIdx.x; slot < s_total; slot += kSelThreads) {
            const uint32_t v = staged[slot], code = v >> 16;
            const int64_t pos = gstart[code] + (slot - lstart[code]);
            if (pos < capacity) em.emit(chunk0 + (v & 0xffffu), code, pos);
        }
    }
}

// ---- classifiers / emitters -------------------------------------------------------------------------------------
// threshold_iterate_2D: bucket = step o with  inside(o) && (o == 0 || !inside(o-1))   (boundary.cpp:219-223).
// classify() evaluates every boundary once on the row's distances and leaves one byte: the step, 0xFF = never,
// 0xFE = several steps (the test is not monotone in o for this row: float rounding) -> recall() recomputes those.
struct Iter2dClass {
    static constexpr bool kMulti = true;
    typedef uint64_t Code;
    typedef float2 Value;
    const float2 *d;
    const float2 *step;   // device, n_off <= 64 entries (x_max[o], x_max[o] * y_max): the uniform product is formed once
    int32_t n_off;
    float y_max;
    uint8_t *note;
    StepSearch search;    // bisection over the boundaries for rows that are clear of all of them (almost all rows)
    __device__ __forceinline__ float2 load(int64_t row) const { return __ldg(d + row); }
    __device__ __noinline__ uint64_t steps_of(const float2 v) const {   // (the rare full scan: one copy)
        uint64_t mask = 0;
        bool prev = false;
        const float xy = __fmul_rn(v.x, y_max);   // line_dist (boundary.cpp:48-50): y0*x_max + x0*y_max - x_max*y_max
        for (int o = 0; o < n_off; o++) {
            const float2 s = __ldg(step + o);
            const bool in = (s.x == 0.0f || y_max == 0.0f)
                                ? line_dist(v.x, v.y, s.x, y_max, 2) <= 0.0f
                                : __fsub_rn(__fadd_rn(__fmul_rn(v.y, s.x), xy), s.y) <= 0.0f;
            if (in && !prev) mask |= 1ull << o;
            prev = in;
        }
        return mask;
    }
    __device__ __forceinline__ uint64_t classify(int64_t row, const float2 v) const {
        const int32_t t = first_admitting_step(search, v);
        if (t >= 0) {  // the sign of every test is certain: one admission (or none)
            note[row] = t == n_off ? 0xFF : (uint8_t)t;
            return t == n_off ? 0ull : 1ull << t;
        }
        const uint64_t mask = steps_of(v);
        note[row] = mask == 0 ? 0xFF : ((mask & (mask - 1)) ? 0xFE : (uint8_t)(__ffsll((long long)mask) - 1));
        return mask;
    }
    __device__ __forceinline__ uint64_t recall(int64_t row) const {
        const uint8_t c = note[row];
        if (c == 0xFF) return 0;
        if (c == 0xFE) return steps_of(__ldg(d + row));
        return 1ull << c;
    }
};
struct Iter2dEmit {
    int64_t n_samples;
    int64_t *out_i, *out_j, *out_o;
    __device__ __forceinline__ void emit(int64_t row, uint32_t step, int64_t pos) const {
        const int64_t i = dev_row_idx(row, n_samples);
        out_i[pos] = i;
        out_j[pos] = dev_col_idx(row, i, n_samples);
        out_o[pos] = step;
    }
};

// radix-sort passes (least significant digit first, 8 bits per pass): stable, so equal keys keep their input order
struct DigitOfU32 {
    static constexpr bool kMulti = false;
    typedef uint32_t Code;
    typedef uint32_t Value;
    const uint32_t *keys;
    int32_t shift;
    __device__ __forceinline__ uint32_t load(int64_t i) const { return __ldg(keys + i); }
    __device__ __forceinline__ uint32_t classify(int64_t, uint32_t k) const { return (k >> shift) & 0xffu; }
    __device__ __forceinline__ uint32_t recall(int64_t i) const { return (__ldg(keys + i) >> shift) & 0xffu; }
};
struct DigitOfU64 {
    static constexpr bool kMulti = false;
    typedef uint32_t Code;
    typedef unsigned long long Value;
    const unsigned long long *keys;
    int32_t shift;
    __device__ __forceinline__ unsigned long long load(int64_t i) const { return __ldg(keys + i); }
    __device__ __forceinline__ uint32_t classify(int64_t, unsigned long long k) const { return (uint32_t)(k >> shift) & 0xffu; }
    __device__ __forceinline__ uint32_t recall(int64_t i) const { return (uint32_t)(__ldg(keys + i) >> shift) & 0xffu; }
};
struct MovePair {  // (uint32 key, int64 value)
    const uint32_t *k_in;
    const int64_t *v_in;
    uint32_t *k_out;
    int64_t *v_out;
    __device__ __forceinline__ void emit(int64_t i, uint32_t, int64_t pos) const {
        k_out[pos] = k_in[i];
        v_out[pos] = v_in[i];
    }
};
struct MoveKey64 {
    const unsigned long long *in;
    unsigned long long *out;
    __device__ __forceinline__ void emit(int64_t i, uint32_t, int64_t pos) const { out[pos] = in[i]; }
};

}  // namespace ppb
