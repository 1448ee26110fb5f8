// Device kernels of the B200 sketch-distance engine (sm_100a only).
//
//   pack_kernel      canonical uint64 [n][K][W] bindash sketches  ->  lane-sliced layout (below)
//   query_kernel     the hot path: per-k matching-bin counts (XOR/AND-NOT + POPC, a4), random-match
//                    correction (a5), log-linear fit (a6), optional assign_threshold epilogue (a7);
//                    one coalesced float2 store per pair in PopPUNK's row order (a2/a8)
//   threshold_kernel standalone assign_threshold (src/boundary.cpp:60-80)
//
// PACKED ("lane-sliced") LAYOUT.  A sketch for one k is viewed as G32 = 2*sketchsize64 groups of 32 bins
// (group g = 2*s + h is the low (h=0) / high (h=1) half of the 64-bit column s); a group has 14 plane
// words.  Groups are cut into slices of 32: lane l of slice t owns group 32*t + l.  One (genome, k, slice)
// is 448 uint32 = 1792 B:
//     words [q*128 + l*4 + e], q=0..2, e=0..3 : plane 4q+e of lane l      (three conflict-free LDS.128)
//     words [384 + l*2 + e],   e=0..1         : plane 12+e  of lane l      (one LDS.64)
// and the whole array is  uint32 [K*n_slices][n_pad][448]  (k-slice-major, genome-minor), so the JB
// consecutive genomes a pipeline stage needs are ONE contiguous TMA bulk copy (4 genomes = 7 KB).
//
// MAPPING (why it looks like this — DESIGN.md has the numbers).  The work per pair is 14 LOP3 per group of
// 32 bins and nothing else of weight, so the kernel is bound by the INT32 logic pipe, not by HBM.  Each
// warp keeps 8 "row" genomes (i) of the current (k, slice) stationary in registers (8 x 14 words per lane:
// the register file is the A-tile), streams "column" genomes (j) through shared memory (TMA-fed ring, one
// LDS.128 feeds 4 x 8 LOP3), and gets the per-pair count with a POPC per lane and one warp REDUX per two
// pairs.  Per-k counts wait in shared memory (uint16) until every k is done, then the same CTA runs the
// regression in float64 and writes the row-ordered outputs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ppb.h"
#include "ppb_ptx.cuh"

namespace ppb {

constexpr int kBbits = 14;
constexpr int kSliceWords = 448;              // 32 lanes x 14 planes
constexpr int kSliceBytes = kSliceWords * 4;  // 1792
constexpr int kRowsPerWarp = 8;               // register-stationary genomes per warp
constexpr int kComputeWarps = 8;
constexpr int kTI = kRowsPerWarp * kComputeWarps;  // 64 rows per tile
#ifndef PPB_JB
#define PPB_JB 4
#endif
#ifndef PPB_STAGES
#define PPB_STAGES 6
#endif
#ifndef PPB_EPI_WARPS
#define PPB_EPI_WARPS 4
#endif
#ifndef PPB_JJ_UNROLL
#define PPB_JJ_UNROLL 4
#endif
constexpr int kJB = PPB_JB;                        // column genomes per pipeline stage
constexpr int kStages = PPB_STAGES;
// (Measured and dropped, profiles/r01_experiments.md + r02_experiments.md section 3: one TMA ring per warp group walking the
//  k-slices out of phase (4 % slower), incrementally kept ring position, early probe of the next stage's barrier, deferred
//  packing of the popcounts, rotated warp roles, staggered warp start — all within noise.)
constexpr int kJJUnroll = PPB_JJ_UNROLL;           // column-loop unroll inside a stage
constexpr int kStageBytes = kJB * kSliceBytes;     // 7168
constexpr int kCntRowWords = kTI / 2 + 4;          // 64 uint16 counts + pad: rows stay 16-B aligned (STS.128)
constexpr int kCntRowWordsWide = kTI + 4;          // 64 uint32 counts + pad: sketches of more than 65535 bins (sketchsize64 >= 1024)
__host__ __device__ constexpr int cnt_row_words(bool wide) { return wide ? kCntRowWordsWide : kCntRowWords; }
constexpr int kEpiWarps = PPB_EPI_WARPS;           // epilogue warps (fit + stores): one per scheduler, so all four are loaded alike
constexpr int kProducerWarp = kComputeWarps + kEpiWarps;  // last warp: TMA producer
constexpr int kThreads = (kComputeWarps + 2 * 4) * 32;
static_assert(kThreads == 512, "4 full warpgroups: setmaxnreg is a warpgroup-wide operation");
// 13 warps put four on scheduler 0, i.e. 128 registers per thread at launch; the warpgroups then trade registers
// (setmaxnreg): helpers shrink, the two compute warpgroups grow back to what the register-stationary tile needs.
// Conservation inside the CTA's pool: 8 x 200 + 4 x 88 + 4 x 24 = 16 x 128 exactly (warps 13-15 only give registers
// back).  The compute branch holds ~150 live values; ptxas schedules against the setmaxnreg budget, and the slack up to
// 200 is what lets it load the next column's plane words before the current column's LOP3s have drained.
#ifndef PPB_REGS_COMPUTE
#define PPB_REGS_COMPUTE 200  // measured at N=100k: 168 -> 741.7 ms, 192 -> 737.3, 200 -> 719.9 (profiles/r01_variants_regs_100k.log)
#endif
#ifndef PPB_REGS_HELPER
#define PPB_REGS_HELPER 88
#endif
constexpr int kRegsCompute = PPB_REGS_COMPUTE, kRegsHelper = PPB_REGS_HELPER, kRegsProducer = 24;
static_assert(kEpiWarps == 4 && kComputeWarps * kRegsCompute + 4 * kRegsHelper + 4 * kRegsProducer <= 16 * 128,
              "setmaxnreg: the compute warps can only grow by what the helper warpgroups give back");
constexpr int kPad = 128;                          // genome padding of packed arrays
constexpr int kMaxTJ = 128;
constexpr int kCtasPerSM = 1;                      // measured: 2 x (4 warps, 32-row tiles) is slower (profiles/)
constexpr int kCntBufs = 2;                        // count tiles: one being filled, one being fitted

struct QueryParams {
    const uint32_t *A;  // packed rows   (queries; == B in self mode)
    const uint32_t *B;  // packed columns (refs)
    int64_t nA, nB, nA_pad, nB_pad;
    int32_t K, n_slices, KS, G32;
    int32_t self, tj;
    int32_t wide;         // per-k counts are uint32 in the count tile (S > 65535); the fit then computes its logs in place
    const int2 *tiles;
    int64_t n_tiles;
    int64_t row_begin, row_end;
    int32_t out_mode;
    void *out;            // row r of the shard is written at index r - row_begin
    // fused multi-GPU exchange (PPB_OUT_DISTS): every row is ALSO stored into these peer buffers (NVLink-mapped,
    // indexed by global row), or once through an NVSwitch multicast address — replaces the all-gather
    void *peer_out[PPB_MAX_PEERS];
    int32_t n_peer_out;
    void *mc_out;
    int32_t stream_stores;  // 1: outputs are written evict-first (st.global.cs)
    // fused edge list (N1): global row indices of pairs on the within side of the boundary, appended unordered
    int32_t edge_mode;      // 0 off, 1: line_dist < 0 (assign == -1), 2: line_dist <= 0 (edge_iterate)
    long long *edge_rows;
    long long edge_cap;
    unsigned long long *edge_count;
    int32_t debug_skip_epilogue;  // measurement only (PPB_DEBUG_SKIP_EPILOGUE): epilogue warps do no work
    int32_t a_policy, b_policy;  // L2 eviction priority of row-genome loads / column-genome TMA (0 normal, 1 last, 2 first)
    int8_t *labels;
    int32_t has_boundary;
    ppb_boundary bnd;
    const float *rand_table;
    int32_t C;
    const uint16_t *clA, *clB;
    unsigned long long *n_degenerate;
    const double *ytab;  // y-table (see ytab_kernel) or nullptr: compute logs in place
    double S, inv_S, tol;
    int32_t S_pow2;
    double x[PPB_MAX_K];
    double xbar[PPB_MAX_K + 1], inv_sxx[PPB_MAX_K + 1], inv_n[PPB_MAX_K + 1];
};

// ------------------------------------------------------------------------------------------------
// pack: one warp per (k-slice, genome) unit of 1792 B.  The unit's source words (its 16 uint64 columns x 14 planes,
// contiguous in the canonical array when the slice is full) are read with coalesced 8-byte loads into shared
// memory, permuted there, and written as one coalesced 1792-byte run: HBM-bound, 2 x 8960 B per genome.
// ------------------------------------------------------------------------------------------------
constexpr int kPackWarps = 8;

// Where a packed part goes: this device's buffer and, in a single-process multi-GPU call, every peer's (NVLink-mapped)
// buffer — each device packs 1/G of the genomes and stores them everywhere, so nobody uploads or packs the whole array.
struct PackDsts {
    uint32_t *p[PPB_MAX_PEERS];
    int32_t n;
};

// Genomes [g_begin, g_end) of the full packed layout (n real genomes padded to n_pad; g >= n are zero sketches);
// `src` holds only this part (genome g is src row idx[g - g_begin], or g - g_begin without an index list).
__global__ void __launch_bounds__(kPackWarps * 32) pack_kernel(const uint64_t *__restrict__ src,
                                                               const int64_t *__restrict__ idx, int64_t g_begin,
                                                               int64_t g_end, int64_t n, int64_t n_pad, int32_t K,
                                                               int32_t ss64, int32_t n_slices, const PackDsts dsts) {
    __shared__ uint32_t stage[kPackWarps][kSliceWords];  // [column-in-slice (16)][plane (14)][half (2)] as uint32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t span = g_end - g_begin;
    const int64_t units = (int64_t)K * n_slices * span;
    const int64_t W = (int64_t)ss64 * kBbits;
    uint32_t *sm = stage[warp];
    for (int64_t u = (int64_t)blockIdx.x * kPackWarps + warp; u < units; u += (int64_t)gridDim.x * kPackWarps) {
        const int64_t gl = u % span, g = g_begin + gl;
        const int32_t ks = (int32_t)(u / span);
        const int32_t k = ks / n_slices, sl = ks - k * n_slices;
        // uint64 columns 16*sl .. 16*sl+15 of this (genome, k): 224 consecutive words (fewer in the last slice)
        const int32_t col0 = sl * 16, n_cols = max(0, min(16, ss64 - col0));
        const int32_t n_words = g < n ? n_cols * kBbits : 0;
        const uint64_t *base = nullptr;
        if (g < n) base = src + ((idx ? idx[gl] : gl) * K + k) * W + (int64_t)col0 * kBbits;
        __syncwarp();
        for (int w = lane; w < kSliceWords / 2; w += 32) {
            const uint64_t v = w < n_words ? __ldg(base + w) : 0ull;
            sm[2 * w] = (uint32_t)v;
            sm[2 * w + 1] = (uint32_t)(v >> 32);
        }
        __syncwarp();
        const int64_t out_off = ((int64_t)ks * n_pad + g) * kSliceWords;
        for (int o = lane; o < kSliceWords; o += 32) {
            int32_t l, plane;  // output word o = plane `plane` of group (lane) `l`, see the layout comment above
            if (o < 384) {
                l = (o & 127) >> 2;
                plane = ((o >> 7) << 2) + (o & 3);
            } else {
                l = (o - 384) >> 1;
                plane = 12 + (o & 1);
            }
            // group l = half (l & 1) of column (l >> 1); staged word index = (column * 14 + plane) * 2 + half
            const uint32_t v = sm[(((l >> 1) * kBbits + plane) << 1) + (l & 1)];
            for (int d = 0; d < dsts.n; d++) dsts.p[d][out_off + o] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a7: src/boundary.cpp:42-58 line_dist — float32, the reference's operation order, no FMA contraction.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float line_dist(float x0, float y0, float x_max, float y_max, int slope) {
    float side = 0.0f;
    if (slope == 2) {
        if (x_max == 0.0f || y_max == 0.0f) {
            side = __fsqrt_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(y0, y0)));
        } else {
            side = __fsub_rn(__fadd_rn(__fmul_rn(y0, x_max), __fmul_rn(x0, y_max)), __fmul_rn(x_max, y_max));
        }
    } else if (slope == 0) {
        side = __fsub_rn(x0, x_max);
    } else if (slope == 1) {
        side = __fsub_rn(y0, y_max);
    }
    return side;
}
__device__ __forceinline__ float boundary_side(float in_tri) {  // boundary.cpp:68-76
    return in_tri == 0.0f ? 0.0f : (in_tri > 0.0f ? 1.0f : -1.0f);
}

__global__ void threshold_kernel(const float2 *__restrict__ d, int64_t n, int32_t slope, float x_max,
                                 float y_max, float *__restrict__ out) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const float2 v = d[r];
        out[r] = boundary_side(line_dist(v.x, v.y, x_max, y_max, slope));
    }
}

// ------------------------------------------------------------------------------------------------
// y-table: y(c) = ln( max(0, c/S - r) / (1 - r) ) for every count c in [0, S], every k and every
// (ref cluster, query cluster) pair — the only transcendental of the per-pair fit becomes one cached
// 8-byte load.  Entries with J < 5/S hold the sentinel +1 (ln J <= 0 always), which ends the series.
// Layout: double [max(C,1)^2][K][S + 1].
// ------------------------------------------------------------------------------------------------
constexpr double kYSentinel = 1.0;

__device__ __forceinline__ double jaccard_of_count(const QueryParams &p, double c, const float *rt, int t) {
    double jac = p.S_pow2 ? c * p.inv_S : c / p.S;
    if (rt) {  // observed_excess(obs, r, 1): max(0, obs - r) / (1 - r)
        const double r = (double)__ldg(rt + t);
        double diff = jac - r;
        if (diff < 0.0) diff = 0.0;
        jac = diff / (1.0 - r);
    }
    return jac;
}

__global__ void ytab_kernel(const QueryParams p, double *__restrict__ ytab) {
    const int S1 = (int)p.S + 1;
    const int64_t per_cp = (int64_t)p.K * S1;
    const int64_t total = per_cp * (p.rand_table ? (int64_t)p.C * p.C : 1);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cp = o / per_cp;
        const int t = (int)((o - cp * per_cp) / S1), c = (int)(o % S1);
        const float *rt = p.rand_table ? p.rand_table + cp * p.K : nullptr;
        const double jac = jaccard_of_count(p, (double)c, rt, t);
        ytab[o] = jac < p.tol ? kYSentinel : log(jac);
    }
}

// ------------------------------------------------------------------------------------------------
// Per-pair epilogue: counts (shared memory) -> Jaccard -> truncated log-linear fit -> outputs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t read_count(const uint32_t *cnt, int t, int tj, int jl, int il, int wide) {
    if (wide) return cnt[(t * tj + jl) * kCntRowWordsWide + il];
    const uint32_t w = cnt[(t * tj + jl) * kCntRowWords + (il >> 1)];
    return (il & 1) ? (w >> 16) : (w & 0xffffu);
}


// exp(x) for x <= 0 in float64 without the special-case handling of the library version: Cody-Waite
// reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18), exponent insert.
__constant__ double kExpPoly[12] = {1.6059043836821613e-10, 2.08767569878681e-09,  2.505210838544172e-08,
                                    2.755731922398589e-07,  2.7557319223985893e-06, 2.48015873015873e-05,
                                    1.984126984126984e-04,  1.388888888888889e-03,  8.333333333333333e-03,
                                    4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
__device__ __forceinline__ double exp_nonpos(double x) {
    x = fmax(x, -700.0);
    const int n = __double2int_rn(x * 1.4426950408889634);
    const double fn = (double)n;
    double r = fma(-fn, 6.93147180369123816490e-01, x);
    r = fma(-fn, 1.90821492927058770002e-10, r);
    double q = kExpPoly[0];
#pragma unroll
    for (int d = 1; d < 12; d++) q = fma(q, r, kExpPoly[d]);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    return q * __hiloint2double((n + 1023) << 20, 0);  // 2^n, n in [-1010, 0]
}

// Per-row facts the epilogue warps precompute once per tile (shared memory).
struct RowInfo {
    long long row_base;  // output row of (i, j) is row_base + j   (already minus row_begin)
    int32_t ytab_off;    // this row's cluster offset into the y-table (elements)
    int32_t i_ok;        // row genome exists
};

// returns true when the pair is on the within side of the boundary (fused edge list)
__device__ __forceinline__ bool store_pair(const QueryParams &p, double sy, double sxy, int n, long long row) {
    float core = 0.0f, acc = 0.0f;
    if (n >= 2) {
        const double beta = (sxy - p.xbar[n] * sy) * p.inv_sxx[n];  // slope     = log(1 - core)
        const double alpha = sy * p.inv_n[n] - beta * p.xbar[n];     // intercept = log(1 - acc)
        core = beta < 0.0 ? (float)(1.0 - exp_nonpos(beta)) : 0.0f;
        acc = alpha < 0.0 ? (float)(1.0 - exp_nonpos(alpha)) : 0.0f;
    }  // else D3: fewer than two usable k -> (0, 0), counted by the caller
    // streaming store (evict-first): 8 B/pair of output must not push the sketch tiles out of L2
    if (p.out) {
        if (p.stream_stores)
            __stcs(reinterpret_cast<float2 *>(p.out) + row, make_float2(core, acc));
        else
            reinterpret_cast<float2 *>(p.out)[row] = make_float2(core, acc);
    }
    if (p.mc_out) {  // one store, replicated to every GPU by the NVSwitch (multimem)
        float2 *dst = reinterpret_cast<float2 *>(p.mc_out) + (row + p.row_begin);
        asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(core), "f"(acc) : "memory");
    } else {
        for (int g = 0; g < p.n_peer_out; g++)  // peer stores over NVLink (write-only, coalesced 256 B per warp row)
            reinterpret_cast<float2 *>(p.peer_out[g])[row + p.row_begin] = make_float2(core, acc);
    }
    if (p.has_boundary) {
        // models.py:1085-1089: assignThreshold(X / self.scale, slope, x_max, y_max)
        const float x0 = __fdiv_rn(core, p.bnd.scale_x), y0 = __fdiv_rn(acc, p.bnd.scale_y);
        const float side = line_dist(x0, y0, p.bnd.x_max, p.bnd.y_max, p.bnd.slope);
        if (p.labels) __stcs(reinterpret_cast<signed char *>(p.labels) + row, (signed char)boundary_side(side));
        return p.edge_mode == 2 ? side <= 0.0f : side < 0.0f;
    }
    return false;
}

// generic path: any output mode, logs computed in place (also the fallback when the y-table would be huge)
__device__ __forceinline__ bool pair_epilogue(const QueryParams &p, const uint32_t *cnt, int jl, int il,
                                              int64_t i, int64_t j, int64_t row, bool &degenerate) {
    const int K = p.K;
    const float *rt = nullptr;
    if (p.rand_table) rt = p.rand_table + ((int64_t)p.clB[j] * p.C + p.clA[i]) * K;

    if (p.out_mode == PPB_OUT_COUNTS) {
        uint32_t *o = reinterpret_cast<uint32_t *>(p.out) + row * K;
        for (int t = 0; t < K; t++) o[t] = read_count(cnt, t, p.tj, jl, il, p.wide);
        return false;
    }
    double sy = 0.0, sxy = 0.0;
    int n = 0;
    bool open = true;
    for (int t = 0; t < K; t++) {
        const double jac = jaccard_of_count(p, (double)read_count(cnt, t, p.tj, jl, il, p.wide), rt, t);
        if (p.out_mode == PPB_OUT_JACCARD) {
            reinterpret_cast<float *>(p.out)[row * K + t] = (float)jac;
            continue;
        }
        if (open) {
            if (jac < p.tol) {
                open = false;
            } else {
                const double y = log(jac);
                sy += y;
                sxy = fma(p.x[t], y, sxy);
                n++;
            }
        }
    }
    if (p.out_mode == PPB_OUT_JACCARD) return false;
    degenerate = n < 2;
    return store_pair(p, sy, sxy, n, row);
}


// fused edge list: warp-aggregated append of the global row index of every pair on the within side (called by all 32 lanes)
__device__ __forceinline__ void append_edge(const QueryParams &p, bool within, long long global_row, int lane) {
    const uint32_t b = __ballot_sync(0xffffffffu, within);
    if (b) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(p.edge_count, (unsigned long long)__popc(b));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(b & ((1u << lane) - 1));
        if (within && at < (unsigned long long)p.edge_cap) p.edge_rows[at] = global_row;
    }
}

// Epilogue of one tile, run by the kEpiWarps epilogue warps (et = 0..95).  Work unit = 4 consecutive rows x
// 32 column slots (lane = column: coalesced 256-B row-order stores); the 4 rows' counts of one k are one LDS.64.
__device__ __forceinline__ void tile_epilogue(const QueryParams &p, const uint32_t *cnt, RowInfo *rinfo, int64_t i0,
                                              int64_t j0, int et, int lane) {
    const int K = p.K, S1 = (int)p.S + 1, tj = p.tj;
    const int tj_shift = 31 - __clz(tj);
    // ---- per-row facts
    for (int il = et; il < kTI; il += kEpiWarps * 32) {
        const int64_t i = i0 + il;
        RowInfo ri;
        ri.i_ok = i < p.nA;
        ri.row_base = (p.self ? p.nB * i - ((i * (i + 1)) >> 1) - 1 - i   // boundary.cpp:33-37
                              : i * p.nB) - p.row_begin;                  // utils.py:224-226
        ri.ytab_off = (p.rand_table && ri.i_ok) ? (int32_t)p.clA[i] * K * S1 : 0;
        rinfo[il] = ri;
    }
    bar_sync(2, kEpiWarps * 32);
    const int ewarp = et >> 5;
    const int n_units = (kTI / 4) * tj / 32;
    const bool fast = p.ytab != nullptr;
    // interior tile: every (i, j) of it is a real pair inside the row shard -> no per-pair tests
    bool interior = i0 + kTI <= p.nA && j0 + tj <= p.nB && (!p.self || i0 + kTI - 1 < j0);
    if (interior) {
        const long long first = rinfo[0].row_base + j0, last = rinfo[kTI - 1].row_base + j0 + tj - 1;
        interior = first >= 0 && last < p.row_end - p.row_begin;
    }
    uint32_t n_deg = 0;
    for (int u = ewarp; u < n_units; u += kEpiWarps) {
        const int flat = u * 32 + lane;
        const int jl = flat & (tj - 1), il0 = (flat >> tj_shift) * 4;
        const int64_t j = j0 + jl;
        const bool j_ok = j < p.nB;
        const uint32_t *cnt_col = cnt + jl * kCntRowWords + (il0 >> 1);
        if (fast) {
            // Written to stay off the ALU pipe (it belongs to the compute warps sharing the scheduler): counts
            // come in as 16-bit loads, table addresses are formed with IMAD / IMAD.WIDE, the series is folded with
            // predicated FP64 adds, and interior tiles skip every per-pair bounds test.
            const int32_t col_off = (p.rand_table && j_ok) ? (int32_t)p.clB[j] * p.C * K * S1 : 0;
            double sy[4] = {0, 0, 0, 0}, sxy[4] = {0, 0, 0, 0};
            int n[4] = {0, 0, 0, 0};
            bool open[4] = {true, true, true, true};
            const double *yrow[4];
#pragma unroll
            for (int r = 0; r < 4; r++) yrow[r] = p.ytab + (col_off + rinfo[il0 + r].ytab_off);
            const uint16_t *c16 = reinterpret_cast<const uint16_t *>(cnt_col);
#pragma unroll 5
            for (int t = 0; t < K; t++) {
                const double x = p.x[t];
                const uint32_t tS1 = (uint32_t)(t * S1);
                double y[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    uint32_t e;  // c + t*(S+1) as an IMAD (FMA pipe)
                    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(e) : "r"((uint32_t)c16[t * tj * (kCntRowWords * 2) + r]), "r"(tS1));
                    y[r] = __ldg(yrow[r] + e);
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    open[r] = open[r] && (y[r] <= 0.0);  // the first k with J < 5/S ends the series
                    if (open[r]) {
                        sy[r] += y[r];
                        sxy[r] = fma(x, y[r], sxy[r]);
                        n[r]++;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const RowInfo ri = rinfo[il0 + r];
                const long long row = ri.row_base + j;
                const bool ok = interior || (ri.i_ok && j_ok && (!p.self || (i0 + il0 + r) < j) && row >= 0 &&
                                             row < p.row_end - p.row_begin);
                bool within = false;
                if (ok) {
                    within = store_pair(p, sy[r], sxy[r], n[r], row);
                    n_deg += n[r] < 2;
                }
                if (p.edge_mode) append_edge(p, within, row + p.row_begin, lane);
            }
        } else {
            for (int r = 0; r < 4; r++) {
                const RowInfo ri = rinfo[il0 + r];
                const long long row = ri.row_base + j;
                const bool ok = ri.i_ok && j_ok && (!p.self || (i0 + il0 + r) < j) && row >= 0 &&
                                row < p.row_end - p.row_begin;
                bool deg = false, within = false;
                if (ok) within = pair_epilogue(p, cnt, jl, il0 + r, i0 + il0 + r, j, row, deg);
                n_deg += deg;
                if (p.edge_mode) append_edge(p, within, row + p.row_begin, lane);   // outside the divergent branch
            }
        }
    }
    if (p.n_degenerate) {
        const uint32_t total = __reduce_add_sync(0xffffffffu, n_deg);
        if (total && lane == 0) atomicAdd(p.n_degenerate, (unsigned long long)total);
    }
}

// ------------------------------------------------------------------------------------------------
// The hot path.  Persistent CTAs (one per SM), warp-specialised:
//   warps 0-7   compute: LOP3/POPC/REDUX stream, never leave it (count tile double-buffered)
//   warps 8-11  epilogue (one per scheduler): counts -> fit -> row-ordered stores of the previous tile, concurrently
//   warp  12    TMA producer: 1-D bulk copies of column-genome slices into the kStages-deep ring (13-15: fillers)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t redux_add(uint32_t v) {
    uint32_t r;
    asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack2(uint32_t lo, uint32_t hi) {  // lo + hi * 65536 on the FMA pipe (IMAD)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}
// lane 0 stores this warp's 8 uint16 counts of one column: one predicated 16-byte shared store.  When a sketch has
// several 32-group slices per k (S > 1024) the counts of the later slices ADD to what the earlier ones stored: every
// lane has read the old value with a broadcast LDS.128 at the top of the column (`old`, latency hidden under the LOP3
// stream) and folds it in with IMADs (FMA pipe) — no branch, so the LOP3 stream keeps its shape.
template <bool kAccumulate>
__device__ __forceinline__ void store_counts(uint32_t dst, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                             uint32_t lane, uint32_t accumulate, const uint4 &old) {
    if (kAccumulate) {
        const uint32_t f = accumulate ? 1u : 0u;
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r0) : "r"(old.x), "r"(f));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r1) : "r"(old.y), "r"(f));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r2) : "r"(old.z), "r"(f));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r3) : "r"(old.w), "r"(f));
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"r"(dst),
        "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(lane)
        : "memory");
}

// uint32 counts (sketches of more than 65535 bins): a slice's partial counts still fit the two 16-bit fields of a REDUX
// word (<= 1024 per slice); they are unpacked and added to the 8 uint32 the earlier slices left (two broadcast LDS.128).
__device__ __forceinline__ void store_counts_wide(uint32_t dst, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3,
                                                  uint32_t lane, uint32_t accumulate, const uint4 &old_lo, const uint4 &old_hi) {
    const uint32_t f = accumulate ? 1u : 0u;
    uint32_t c0 = r0 & 0xffffu, c1 = r0 >> 16, c2 = r1 & 0xffffu, c3 = r1 >> 16;
    uint32_t c4 = r2 & 0xffffu, c5 = r2 >> 16, c6 = r3 & 0xffffu, c7 = r3 >> 16;
    c0 += old_lo.x * f, c1 += old_lo.y * f, c2 += old_lo.z * f, c3 += old_lo.w * f;
    c4 += old_hi.x * f, c5 += old_hi.y * f, c6 += old_hi.z * f, c7 += old_hi.w * f;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %9, 0;\n\t@p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n\t"
        "@p st.shared.v4.u32 [%0+16], {%5,%6,%7,%8};\n\t}" ::"r"(dst),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(c5), "r"(c6), "r"(c7), "r"(lane)
        : "memory");
}

struct SmemLayout {
    uint32_t cnt_bytes;   // one count tile
    uint32_t off_cnt, off_rinfo, off_bar, off_trash, total;
};
__host__ __device__ inline SmemLayout smem_layout(int K, int tj, bool wide) {
    SmemLayout L;
    L.cnt_bytes = ((uint32_t)K * tj * cnt_row_words(wide) * 4 + 127u) & ~127u;
    L.off_cnt = kStages * kStageBytes;
    L.off_rinfo = L.off_cnt + kCntBufs * L.cnt_bytes;
    L.off_bar = L.off_rinfo + kTI * (uint32_t)sizeof(RowInfo);
    L.off_trash = L.off_bar + (2 * kStages + 2 * kCntBufs) * 8;
    L.total = L.off_trash + kComputeWarps * 32;
    return L;
}

// kMode 0: S <= 1024 (one 32-group slice per k) — the common case; drops the accumulate path.
// kMode 1: several slices per k accumulate into uint16 counts.  kMode 2: ... into uint32 counts (S > 65535).
template <int kMode>
__global__ void __launch_bounds__(kThreads, kCtasPerSM) query_kernel(const __grid_constant__ QueryParams p) {
    constexpr bool kSingleSlice = kMode == 0, kWide = kMode == 2;
    constexpr int kRowW = cnt_row_words(kWide);
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemLayout L = smem_layout(p.K, p.tj, kWide);
    uint8_t *stage_base = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    uint64_t *empty = full + kStages;
    uint64_t *cfull = empty + kStages;    // count tile b complete (all compute warps arrived)
    uint64_t *cempty = cfull + kCntBufs;  // count tile b consumed (all epilogue warps arrived)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeWarps);
        }
        for (int b = 0; b < kCntBufs; b++) {
            mbar_init(&cfull[b], kComputeWarps);
            mbar_init(&cempty[b], kEpiWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int tj = p.tj, n_jb = tj / kJB, KS = p.KS;

    if (warp >= kProducerWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
        if (warp >= kProducerWarp + 1) return;  // filler warps of the producer's warpgroup
        // ===== TMA producer: streams column-genome slices of every (tile, k, slice) into the ring =====
        if (lane == 0) {
            const uint64_t pol_b = l2_policy(p.b_policy);
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int2 tc = p.tiles[tile];
                const int64_t j0 = (int64_t)tc.y * tj;
                for (int ks = 0; ks < KS; ks++) {
                    const uint32_t *src = p.B + ((int64_t)ks * p.nB_pad + j0) * kSliceWords;
                    for (int jb = 0; jb < n_jb; jb++, it++) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait_relaxed(&empty[s], ph ^ 1, 100);
                        mbar_arrive_expect_tx(&full[s], kStageBytes);
                        tma_load_1d_hint(stage_base + s * kStageBytes, src + (int64_t)jb * kJB * kSliceWords,
                                         kStageBytes, &full[s], pol_b);
                    }
                }
            }
        }
        return;
    }

    if (warp >= kComputeWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsHelper));
        // ===== epilogue warps: fit + stores of tile t while the compute warps are already in tile t+1 =====
        const int et = (warp - kComputeWarps) * 32 + lane;
        RowInfo *rinfo = reinterpret_cast<RowInfo *>(smem + L.off_rinfo);
        uint32_t lt = 0;
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, lt++) {
            const int2 tc = p.tiles[tile];
            const uint32_t b = lt % kCntBufs, ph = (lt / kCntBufs) & 1;
            mbar_wait_relaxed(&cfull[b], ph, 500);
            const uint32_t *cnt = reinterpret_cast<const uint32_t *>(smem + L.off_cnt + b * L.cnt_bytes);
            if (!p.debug_skip_epilogue) tile_epilogue(p, cnt, rinfo, (int64_t)tc.x * kTI, (int64_t)tc.y * tj, et, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&cempty[b]);
            bar_sync(2, kEpiWarps * 32);  // rinfo is rewritten by the next tile
        }
        return;
    }

    // ===== compute warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    // Software pipeline over columns: while the LOP3 stream of column c runs, the packed partial counts of
    // column c-1 go through REDUX and are stored at the end — no POPC/REDUX latency is ever waited for.
    const uint32_t trash_addr = smem_u32(smem + L.off_trash) + warp * 32;
    const uint64_t pol_a = l2_policy(p.a_policy);  // the band's row genomes are re-read by every column tile: keep them
    uint32_t it = 0, lt = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, lt++) {
        const int2 tc = p.tiles[tile];
        const int64_t i0 = (int64_t)tc.x * kTI;
        const uint32_t cb = lt % kCntBufs, cph = (lt / kCntBufs) & 1;
        const uint32_t cnt_addr = smem_u32(smem + L.off_cnt + cb * L.cnt_bytes) + warp * (kWide ? kRowsPerWarp : kRowsPerWarp / 2) * 4;
        mbar_wait(&cempty[cb], cph ^ 1);  // the epilogue warps are done with this count tile (2 tiles ago)

        uint32_t pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0;  // packed (2 x uint16) partial counts of the previous column
        uint32_t pdst = trash_addr, pacc = 0;          // where they go; whether they add to an earlier slice

        for (int ks = 0; ks < KS; ks++) {
            const int k = kSingleSlice ? ks : ks / p.n_slices;
            const int sl = kSingleSlice ? 0 : ks - k * p.n_slices;
            const uint32_t valid = (sl * 32 + lane < p.G32) ? 0xffffffffu : 0u;

            // register-stationary row genomes: 8 x 14 plane words of this lane's group
            uint32_t a[kRowsPerWarp][kBbits];
            {
                const uint32_t *ap = p.A + ((int64_t)ks * p.nA_pad + i0 + warp * kRowsPerWarp) * kSliceWords;
#pragma unroll
                for (int g = 0; g < kRowsPerWarp; g++) {
                    const uint4 *q4 = reinterpret_cast<const uint4 *>(ap + g * kSliceWords);
                    const uint4 v0 = ldg128_hint(q4 + lane, pol_a), v1 = ldg128_hint(q4 + 32 + lane, pol_a),
                                v2 = ldg128_hint(q4 + 64 + lane, pol_a);
                    const uint2 v3 = ldg64_hint(reinterpret_cast<const uint2 *>(ap + g * kSliceWords + 384) + lane, pol_a);
                    a[g][0] = v0.x, a[g][1] = v0.y, a[g][2] = v0.z, a[g][3] = v0.w;
                    a[g][4] = v1.x, a[g][5] = v1.y, a[g][6] = v1.z, a[g][7] = v1.w;
                    a[g][8] = v2.x, a[g][9] = v2.y, a[g][10] = v2.z, a[g][11] = v2.w;
                    a[g][12] = v3.x, a[g][13] = v3.y;
                }
            }
            const uint32_t cnt_k = cnt_addr + k * tj * kRowW * 4;

            for (int jb = 0; jb < n_jb; jb++, it++) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&full[s], ph);
                const uint8_t *sb = stage_base + s * kStageBytes;
                uint32_t dst = cnt_k + jb * kJB * kRowW * 4;
#pragma unroll kJJUnroll
                for (int jj = 0; jj < kJB; jj++, dst += kRowW * 4) {
                    uint4 old = make_uint4(0u, 0u, 0u, 0u), old_hi = make_uint4(0u, 0u, 0u, 0u);
                    if (!kSingleSlice) old = lds128(pdst);  // what earlier slices of this k stored for the previous column
                    if (kWide) old_hi = lds128(pdst + 16);
                    const uint4 *b4 = reinterpret_cast<const uint4 *>(sb + jj * kSliceBytes);
                    const uint4 b0 = b4[lane], b1 = b4[32 + lane], b2 = b4[64 + lane];
                    const uint2 b3 = reinterpret_cast<const uint2 *>(sb + jj * kSliceBytes + 1536)[lane];
                    const uint32_t bw[kBbits] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z,
                                                 b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y};
                    uint32_t c[kRowsPerWarp];
                    // four AND-chains advance together, plane by plane: a dependent LOP3 is always >= 4
                    // instructions behind its producer, so one warp alone can keep the ALU pipe full
                    {
                        uint32_t x0 = valid, x1 = valid, x2 = valid, x3 = valid;
#pragma unroll
                        for (int q = 0; q < kBbits; q++) {
                            x0 = and_xnor(x0, a[0][q], bw[q]);
                            x1 = and_xnor(x1, a[1][q], bw[q]);
                            x2 = and_xnor(x2, a[2][q], bw[q]);
                            x3 = and_xnor(x3, a[3][q], bw[q]);
                        }
                        c[0] = __popc(x0), c[1] = __popc(x1), c[2] = __popc(x2), c[3] = __popc(x3);
                    }
                    // previous column: warp-sum of its packed counts (inputs were ready an iteration ago)
                    const uint32_t r0 = redux_add(pk0), r1 = redux_add(pk1), r2 = redux_add(pk2), r3 = redux_add(pk3);
                    {
                        uint32_t x0 = valid, x1 = valid, x2 = valid, x3 = valid;
#pragma unroll
                        for (int q = 0; q < kBbits; q++) {
                            x0 = and_xnor(x0, a[4][q], bw[q]);
                            x1 = and_xnor(x1, a[5][q], bw[q]);
                            x2 = and_xnor(x2, a[6][q], bw[q]);
                            x3 = and_xnor(x3, a[7][q], bw[q]);
                        }
                        c[4] = __popc(x0), c[5] = __popc(x1), c[6] = __popc(x2), c[7] = __popc(x3);
                    }
                    if (kWide)
                        store_counts_wide(pdst, r0, r1, r2, r3, lane, pacc, old, old_hi);
                    else
                        store_counts<!kSingleSlice>(pdst, r0, r1, r2, r3, lane, pacc, old);
                    // two 16-bit partial counts per REDUX; a slice contributes <= 1024 per pair
                    pk0 = pack2(c[0], c[1]), pk1 = pack2(c[2], c[3]);
                    pk2 = pack2(c[4], c[5]), pk3 = pack2(c[6], c[7]);
                    pdst = dst;
                    pacc = sl;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        }
        {  // drain the pipeline: the tile's last column
            const uint32_t r0 = redux_add(pk0), r1 = redux_add(pk1), r2 = redux_add(pk2), r3 = redux_add(pk3);
            uint4 old = make_uint4(0u, 0u, 0u, 0u), old_hi = make_uint4(0u, 0u, 0u, 0u);
            if (!kSingleSlice) old = lds128(pdst);
            if (kWide) {
                old_hi = lds128(pdst + 16);
                store_counts_wide(pdst, r0, r1, r2, r3, lane, pacc, old, old_hi);
            } else {
                store_counts<!kSingleSlice>(pdst, r0, r1, r2, r3, lane, pacc, old);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&cfull[cb]);  // release: this warp's counts of the tile are visible
    }
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe micro-roofline kernels (bench.py reports the hot kernel against these).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 2) microbench_kernel(int64_t iters, uint32_t *sink, uint32_t seed) {
    uint32_t x[8], av[14], acc = 0, y[4] = {1, 2, 3, 4};
    float f[4] = {1.f, 2.f, 3.f, 4.f};
#pragma unroll
    for (int g = 0; g < 8; g++) x[g] = seed * (threadIdx.x + 1) + g * 0x9e3779b9u;
#pragma unroll
    for (int r = 0; r < 14; r++) av[r] = (seed ^ (blockIdx.x * 2654435761u)) + r * 0x85ebca6bu;
    for (int64_t i = 0; i < iters; i++) {
        if (MODE == 0) {  // LOP3 only: 8 independent chains x 14 (the hot loop's dependency shape)
#pragma unroll
            for (int r = 0; r < 14; r++)
#pragma unroll
                for (int g = 0; g < 8; g++) x[g] = and_xnor(x[g], av[r], av[(r + g) % 14]);
        } else if (MODE == 1) {  // POPC only
#pragma unroll
            for (int r = 0; r < 14; r++)
#pragma unroll
                for (int g = 0; g < 8; g++) x[g] = __popc(x[g]);
        } else if (MODE == 2) {  // the hot loop's mix: 14 LOP3 : 1 POPC : 1 IADD
#pragma unroll
            for (int g = 0; g < 8; g++) {
                uint32_t bits = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < 14; r++) bits = and_xnor(bits, av[r], x[g]);
                x[g] += __popc(bits);
            }
        } else if (MODE == 3) {  // warp REDUX
#pragma unroll
            for (int g = 0; g < 8; g++) x[g] = __reduce_add_sync(0xffffffffu, x[g]);
        } else {
            // MODE 4/5/6: does a non-ALU instruction issued between LOP3s cost LOP3 throughput?
            // 4 interleaved chains x 14 LOP3 (x2), plus per 14 LOP3: 4 IMAD (mode 4), 4 LDS (mode 5), 4 FFMA (mode 6)
            __shared__ uint32_t sm[256 * 4];
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int r = 0; r < 14; r++)
#pragma unroll
                    for (int g = 0; g < 4; g++) x[half * 4 + g] = and_xnor(x[half * 4 + g], av[r], av[(r + g) % 14]);
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    if (MODE == 4) y[e & 3] = y[e & 3] * 3u + av[e & 7];
                    if (MODE == 5) y[e & 3] = reinterpret_cast<volatile uint32_t *>(sm)[threadIdx.x + (e & 3) * 256];
                    if (MODE == 6) f[e & 3] = fmaf(f[e & 3], 1.0001f, 0.5f);
                }
            }
        }
    }
    x[0] ^= y[0] ^ y[1] ^ y[2] ^ y[3] ^ __float_as_uint(f[0] + f[1] + f[2] + f[3]);
#pragma unroll
    for (int g = 0; g < 8; g++) acc ^= x[g];
    if (acc == 0x12345678u) sink[0] = acc;  // keep the work alive
}


// The compute warps' column body as a stand-alone loop (no barriers, no TMA, no epilogue warps): 8 register-stationary
// rows x 14 planes, per column 3 LDS.128 + 1 LDS.64, 112 LOP3 in four interleaved chains, 8 POPC, 4 IMAD packs,
// 4 REDUX and one predicated STS.128 — launched with kRows-row tiles and W warps per scheduler.  It answers "what
// fraction of the LOP3 pipe can THIS instruction mix reach with W in-order warps per scheduler", i.e. how much of the
// distance kernel's gap to the LOP3-only peak is the mix itself and how much is synchronisation.
template <int kRows, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads, 1) mixbench_kernel(int64_t iters, uint32_t *sink, uint32_t seed, int with_lds) {
    extern __shared__ __align__(128) uint8_t msm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *cols = reinterpret_cast<uint32_t *>(msm);                 // 4 column slices (7 KB), shared by all warps
    for (int w = threadIdx.x; w < 4 * kSliceWords; w += blockDim.x) cols[w] = seed * (w + 1) + blockIdx.x;
    __syncthreads();
    uint32_t a[kRows][kBbits];
#pragma unroll
    for (int g = 0; g < kRows; g++)
#pragma unroll
        for (int q = 0; q < kBbits; q++) a[g][q] = (seed ^ (blockIdx.x * 2654435761u)) + (g * 14 + q) * 0x85ebca6bu + threadIdx.x;
    const uint32_t base = smem_u32(msm), trash = base + 4 * kSliceBytes + warp * 16;
    uint32_t pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0;
    uint32_t b_reg[kBbits];
#pragma unroll
    for (int q = 0; q < kBbits; q++) b_reg[q] = seed + q * 0x9e3779b9u + lane;
    for (int64_t i = 0; i < iters; i++) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            uint32_t bw[kBbits];
            if (with_lds) {
                const uint32_t cb = base + jj * kSliceBytes + lane * 16;
                const uint4 b0 = lds128(cb), b1 = lds128(cb + 512), b2 = lds128(cb + 1024);
                const uint2 b3 = lds64(base + jj * kSliceBytes + 1536 + lane * 8);
                bw[0] = b0.x, bw[1] = b0.y, bw[2] = b0.z, bw[3] = b0.w, bw[4] = b1.x, bw[5] = b1.y, bw[6] = b1.z;
                bw[7] = b1.w, bw[8] = b2.x, bw[9] = b2.y, bw[10] = b2.z, bw[11] = b2.w, bw[12] = b3.x, bw[13] = b3.y;
            } else {
#pragma unroll
                for (int q = 0; q < kBbits; q++) bw[q] = b_reg[q] + pk0;  // keeps the values loop-variant
            }
            uint32_t c[kRows];
#pragma unroll
            for (int h = 0; h < kRows; h += 4) {
                uint32_t x0 = 0xffffffffu, x1 = 0xffffffffu, x2 = 0xffffffffu, x3 = 0xffffffffu;
#pragma unroll
                for (int q = 0; q < kBbits; q++) {
                    x0 = and_xnor(x0, a[h][q], bw[q]);
                    if (h + 1 < kRows) x1 = and_xnor(x1, a[h + 1][q], bw[q]);
                    if (h + 2 < kRows) x2 = and_xnor(x2, a[h + 2][q], bw[q]);
                    if (h + 3 < kRows) x3 = and_xnor(x3, a[h + 3][q], bw[q]);
                }
                c[h] = __popc(x0);
                if (h + 1 < kRows) c[h + 1] = __popc(x1);
                if (h + 2 < kRows) c[h + 2] = __popc(x2);
                if (h + 3 < kRows) c[h + 3] = __popc(x3);
            }
            const uint32_t r0 = redux_add(pk0), r1 = redux_add(pk1), r2 = redux_add(pk2), r3 = redux_add(pk3);
            store_counts<false>(trash, r0, r1, r2, r3, lane, 0, make_uint4(0, 0, 0, 0));
            pk0 = pack2(c[0], c[1 % kRows]), pk1 = pack2(c[2 % kRows], c[3 % kRows]);
            pk2 = pack2(c[4 % kRows], c[5 % kRows]), pk3 = pack2(c[6 % kRows], c[7 % kRows]);
        }
    }
    if ((pk0 ^ pk1 ^ pk2 ^ pk3) == 0x12345678u) sink[0] = pk0;
}

}  // namespace ppb
