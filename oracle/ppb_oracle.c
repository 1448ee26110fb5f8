/*
 * ppb_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm for the one hot path this repo replaces
 * (core/accessory sketch distances behind PopPUNK/sketchlib.py:475-632 queryDatabase()).
 * It is the checker for the CUDA engine and the CPU baseline arm of bench.py.  Nothing in
 * the product path (poppunk_b200/, libppb.so) may import, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: "parity unpinned" for the (pi, a) VALUES.
 *   The arithmetic lives in the third-party dependency pp-sketchlib (bacpop/pp-sketchlib,
 *   required >= 2.0.1: PopPUNK/__init__.py:9-11, environment.yml:22, setup.py:113).  Its source
 *   is not under /root/reference, it is not installed in this image, and there is no network,
 *   so it cannot be built into oracle/_ref or imported to generate vectors.  PopPUNK's own tests
 *   pin no distance values (test/run_test.py checks exit codes; test/test-update-gpu.py:24-29,
 *   85-90 accepts R^2 >= 0.99 between two runs of the same library).  This file therefore
 *   restates pp-sketchlib's PUBLISHED algorithm (BinDash b-bit one-permutation MinHash,
 *   citation.py:35-38; the docs cited below) and is anchored on what the reference does pin:
 *     (i)   sketch schema: W = sketchsize64*bbits uint64 words per k, bbits = 14
 *           (test/json_sketch.txt: sketchsize64=156, bbits=14, 2184 words; web.py:14-61)
 *     (ii)  output row order (utils.py:199-226; src/boundary.cpp:22-37,97-123)
 *     (iii) model  pr(a,b) = (1-a)(1-c)^k, log-linear fit, clamps at 0, output (core, acc)
 *           (sketchlib.py:482, 635-670 fitKmerCurve — checked against scipy in tests/)
 *     (iv)  "Jaccard distances < 5/s are ignored in the fit" (docs/sketching.rst:161-165)
 *     (v)   random-match correction enabled on every production call (sketchlib.py:533,589)
 *     (vi)  assign_threshold / line_dist — src/boundary.cpp:42-80 IS in the reference tree and
 *           is restated here operation for operation; its known-answer test
 *           (test/test-refine.py:46-61) is a golden vector in tests/golden/.
 *
 * Decisions that upstream does not let us verify here:
 *   D1 bbits = 14; word [s*bbits+b] = bit b of bins 64s..64s+63 (bindash fillusigs layout).
 *   D2 the first k (ascending) with J_k < 5/S ends the series; it and all larger k are dropped.
 *   D3 fewer than 2 usable k: upstream aborts ("Fitting k-mer gradient failed",
 *      docs/troubleshooting.rst:176-191); here the row is (0,0) — PopPUNK's own failure
 *      convention, sketchlib.py:662-667 — and is counted in *n_degenerate.
 *   D4 regression in float64, cast to float32 last.
 *   D5 random_correct=False  <=>  r_k = 0.
 *   D6 klist ascending, subset of the DB's kmers.
 *   D7 no b-bit collision correction: intersize == samebits for every S (bindash's
 *      expected_samebits = S >> bbits branch returns samebits unchanged when non-zero, and is
 *      the identity when zero).  Kept behind PPO_BBIT_CORRECTION so it can be flipped.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PPO_BBITS 14
#define PPO_MAX_K 32
#define PPO_MIN_JACCARD_BINS 5.0
#define PPO_BBIT_CORRECTION 0 /* D7 */

#define PPO_OUT_DISTS 0
#define PPO_OUT_JACCARD 1
#define PPO_OUT_COUNTS 2

typedef struct ppo_boundary {
    int32_t slope;
    float x_max, y_max;
    float scale_x, scale_y;
} ppo_boundary;

/* ------------------------------------------------------------------------------------------
 * Index maps — src/boundary.cpp:18-37 (rows_to_samples, calc_row_idx, calc_col_idx,
 * square_to_condensed); Python mirror utils.py:199-261.
 * ---------------------------------------------------------------------------------------- */
int64_t ppo_square_to_condensed(int64_t i, int64_t j, int64_t n) {
    /* boundary.cpp:33-37 */
    return n * i - ((i * (i + 1)) >> 1) + j - 1 - i;
}

int64_t ppo_calc_row_idx(int64_t k, int64_t n) {
    /* boundary.cpp:22-27: n - 2 - floor(sqrt(-8k + 4n(n-1) - 7)/2 - 0.5).  The double sqrt is
     * followed by an exact integer fix-up so the map is right for every n (the reference's
     * formula alone is exact for the n PopPUNK uses; validated for n <= 2001 in tests/). */
    double d = sqrt((double)(-8 * k + 4 * n * (n - 1) - 7));
    int64_t i = n - 2 - (int64_t)floor(d / 2.0 - 0.5);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && ppo_square_to_condensed(i, i + 1, n) > k) i--;
    while (i < n - 2 && ppo_square_to_condensed(i + 1, i + 2, n) <= k) i++;
    return i;
}

int64_t ppo_calc_col_idx(int64_t k, int64_t i, int64_t n) {
    /* boundary.cpp:29-31 */
    return k + i + 1 - n * (n - 1) / 2 + (n - i) * ((n - i) - 1) / 2;
}

int64_t ppo_num_rows(int64_t n_ref, int64_t n_qry, int self) {
    return self ? n_ref * (n_ref - 1) / 2 : n_ref * n_qry;
}

/* ------------------------------------------------------------------------------------------
 * a4: per-k bindash Jaccard numerator.  [UPSTREAM-RECALL: pp-sketchlib src/sketch/bitfuncs.cpp
 * calc_intersize; method = BinDash, citation.py:35-38.]
 *   samebits = sum_s popcount( AND_b ~(A[s*bbits+b] ^ B[s*bbits+b]) )
 * = number of bins whose bbits-bit signatures agree.
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t ppo_intersize(const uint64_t *a, const uint64_t *b, int ss64, int bbits) {
    uint32_t samebits = 0;
    for (int s = 0; s < ss64; s++) {
        uint64_t bits = ~(uint64_t)0;
        for (int p = 0; p < bbits; p++) bits &= ~(a[s * bbits + p] ^ b[s * bbits + p]);
        samebits += (uint32_t)__builtin_popcountll(bits);
    }
#if PPO_BBIT_CORRECTION
    const uint64_t maxnbits = (uint64_t)ss64 * 64;
    const uint64_t expected = maxnbits >> bbits;
    if (expected) {
        uint64_t ret = samebits > expected ? samebits - expected : 0;
        return (uint32_t)(ret * maxnbits / (maxnbits - expected));
    }
#endif
    return samebits;
}

/* a5: random-match correction.  [UPSTREAM-RECALL observed_excess(obs, exp, max=1);
 * docs/sketching.rst:107-118 for what r is.]  J = max(0, J_obs - r) / (1 - r). */
static inline double ppo_observed_excess(double obs, double r) {
    double diff = obs - r;
    if (diff < 0) diff = 0;
    return diff * 1.0 / (1.0 - r);
}

/* a6: across-k regression.  Model sketchlib.py:482; clamp = bounds <= 0 at sketchlib.py:660;
 * return order (core, accessory) sketchlib.py:669-670; truncation docs/sketching.rst:161-165.
 * Returns 1 if the pair was degenerate (D3). */
/* the fit itself, on y[t] = ln J_t for the n usable k (shared with the tuned arm, ppb_oracle_tuned.inc) */
static inline int ppo_fit_logs(const double *y, int n, const int32_t *kmers, float *core, float *acc) {
    if (n < 2) {
        *core = 0.0f;
        *acc = 0.0f;
        return 1;
    }
    double xbar = 0, ybar = 0;
    for (int t = 0; t < n; t++) {
        xbar += (double)kmers[t];
        ybar += y[t];
    }
    xbar /= n;
    ybar /= n;
    double sxx = 0, sxy = 0;
    for (int t = 0; t < n; t++) {
        double dx = (double)kmers[t] - xbar;
        sxx += dx * dx;
        sxy += dx * (y[t] - ybar);
    }
    double beta = sxy / sxx;             /* slope      = log(1 - core) */
    double alpha = ybar - beta * xbar;   /* intercept  = log(1 - acc)  */
    *core = beta < 0 ? (float)(1.0 - exp(beta)) : 0.0f;
    *acc = alpha < 0 ? (float)(1.0 - exp(alpha)) : 0.0f;
    return 0;
}

static inline int ppo_regress(const double *jac, const int32_t *kmers, int K, double S, float *core,
                              float *acc) {
    const double tolerance = PPO_MIN_JACCARD_BINS / S;
    int n = K;
    for (int t = 0; t < K; t++) {
        if (jac[t] < tolerance) {
            n = t;
            break;
        }
    }
    double y[PPO_MAX_K];
    for (int t = 0; t < n; t++) y[t] = log(jac[t]);
    return ppo_fit_logs(y, n, kmers, core, acc);
}

/* Regression only, on caller-supplied per-k Jaccards double [rows][K] (pins a6 against the golden
 * vectors made from PopPUNK/sketchlib.py:635-670 fitKmerCurve). Returns the degenerate-row count. */
int64_t ppo_regress_rows(const double *jac, int64_t rows, const int32_t *kmers, int32_t K, double S,
                         float *out) {
    int64_t deg = 0;
    for (int64_t r = 0; r < rows; r++)
        deg += ppo_regress(jac + r * K, kmers, K, S, out + 2 * r, out + 2 * r + 1);
    return deg;
}

/* a7: src/boundary.cpp:42-58 line_dist, float32, same operation order, no FMA contraction
 * (the reference is built for baseline x86-64: CMakeLists.txt has no -march / -mfma). */
static inline float ppo_line_dist(float x0, float y0, float x_max, float y_max, int slope) {
    volatile float boundary_side = 0;
    if (slope == 2) {
        if (x_max == 0 || y_max == 0) {
            volatile float xx = x0 * x0;
            volatile float yy = y0 * y0;
            boundary_side = sqrtf(xx + yy);
        } else {
            volatile float t1 = y0 * x_max;
            volatile float t2 = x0 * y_max;
            volatile float t3 = x_max * y_max;
            volatile float s = t1 + t2;
            boundary_side = s - t3;
        }
    } else if (slope == 0) {
        boundary_side = x0 - x_max;
    } else if (slope == 1) {
        boundary_side = y0 - y_max;
    }
    return boundary_side;
}

static inline float ppo_side(float in_tri) {
    /* boundary.cpp:68-76 */
    if (in_tri == 0) return 0.0f;
    return in_tri > 0 ? 1.0f : -1.0f;
}

/* src/boundary.cpp:60-80 assign_threshold */
int ppo_assign_threshold(const float *dists, int64_t n, int32_t slope, float x_max, float y_max,
                         float *out, int32_t threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t r = 0; r < n; r++)
        out[r] = ppo_side(ppo_line_dist(dists[2 * r], dists[2 * r + 1], x_max, y_max, slope));
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a2: the whole call.  Host-buffer twin of ppb_query_host (include/ppb.h) so tests can feed
 * both the same arguments.  ref/qry: uint64 [n][K][W], W = sketchsize64*bbits.
 * ---------------------------------------------------------------------------------------- */
int ppo_query_host(const uint64_t *ref, int64_t n_ref, const uint64_t *qry, int64_t n_qry,
                   const int32_t *kmers, int32_t K, int32_t sketchsize64, int32_t bbits,
                   const float *rand_table, int32_t n_clusters, const uint16_t *ref_cluster,
                   const uint16_t *qry_cluster, int64_t row_begin, int64_t row_end,
                   int32_t out_mode, void *out, const ppo_boundary *boundary, int8_t *labels,
                   int64_t *n_degenerate, int32_t threads) {
    if (!ref || K < 1 || K > PPO_MAX_K || bbits != PPO_BBITS || sketchsize64 < 1) return 1;
    const int self = (qry == NULL);
    const int64_t n_rows = ppo_num_rows(n_ref, self ? n_ref : n_qry, self);
    if (row_begin < 0 || row_end > n_rows || row_begin > row_end) return 1;
    if (threads < 1) threads = 1;
    const int64_t W = (int64_t)sketchsize64 * bbits;
    const int64_t stride = (int64_t)K * W;
    const double S = 64.0 * sketchsize64;
    int64_t degenerate = 0;

#pragma omp parallel num_threads(threads) reduction(+ : degenerate)
    {
#ifdef _OPENMP
        const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
        const int nt = 1, tid = 0;
#endif
        const int64_t total = row_end - row_begin;
        const int64_t lo = row_begin + total * tid / nt;
        const int64_t hi = row_begin + total * (tid + 1) / nt;
        int64_t i = 0, j = 0; /* self: i<j genome indices; non-self: i = query, j = ref */
        if (lo < hi) {
            if (self) {
                i = ppo_calc_row_idx(lo, n_ref);
                j = ppo_calc_col_idx(lo, i, n_ref);
            } else {
                i = lo / n_ref;
                j = lo % n_ref;
            }
        }
        for (int64_t row = lo; row < hi; row++) {
            /* self row (i<j): the reference yields (refSeqs[j], refSeqs[i]) — utils.py:220-222;
             * non-self row q*R + r — utils.py:224-226. */
            const uint64_t *A = self ? ref + i * stride : qry + i * stride;
            const uint64_t *B = ref + j * stride;
            const int cq = self ? (ref_cluster ? ref_cluster[i] : 0) : (qry_cluster ? qry_cluster[i] : 0);
            const int cr = ref_cluster ? ref_cluster[j] : 0;
            double jac[PPO_MAX_K];
            uint32_t cnt[PPO_MAX_K];
            for (int t = 0; t < K; t++) {
                cnt[t] = ppo_intersize(A + t * W, B + t * W, sketchsize64, bbits);
                const double r = rand_table ? (double)rand_table[((int64_t)cr * n_clusters + cq) * K + t] : 0.0;
                jac[t] = ppo_observed_excess((double)cnt[t] / S, r);
            }
            const int64_t o = row - row_begin;
            if (out_mode == PPO_OUT_COUNTS) {
                for (int t = 0; t < K; t++) ((uint32_t *)out)[o * K + t] = cnt[t];
            } else if (out_mode == PPO_OUT_JACCARD) {
                for (int t = 0; t < K; t++) ((float *)out)[o * K + t] = (float)jac[t];
            } else {
                float core, acc;
                degenerate += ppo_regress(jac, kmers, K, S, &core, &acc);
                if (out) {
                    ((float *)out)[2 * o] = core;
                    ((float *)out)[2 * o + 1] = acc;
                }
                if (boundary && labels) {
                    /* models.py:1085-1089: X/self.scale in float32, then assignThreshold */
                    volatile float x0 = core / boundary->scale_x;
                    volatile float y0 = acc / boundary->scale_y;
                    labels[o] = (int8_t)ppo_side(
                        ppo_line_dist(x0, y0, boundary->x_max, boundary->y_max, boundary->slope));
                }
            }
            /* advance (i,j) in output row order */
            if (self) {
                if (++j >= n_ref) {
                    i++;
                    j = i + 1;
                }
            } else {
                if (++j >= n_ref) {
                    i++;
                    j = 0;
                }
            }
        }
    }
    if (n_degenerate) *n_degenerate = degenerate;
    return 0;
}

int ppo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * "Next" rows — restated from src/boundary.cpp, which IS in the reference tree.
 * ---------------------------------------------------------------------------------------- */

/* N1: src/boundary.cpp:82-95 edge_iterate — rows with line_dist <= 0, in row order, as (i, j). Returns count. */
int64_t ppo_edge_iterate(const float *dists, int64_t n_rows, int32_t slope, float x_max, float y_max,
                         int64_t *out_i, int64_t *out_j) {
    const int64_t n_samples = (int64_t)(0.5 * (1 + sqrt(1 + 8.0 * (double)n_rows))); /* rows_to_samples :18-20 */
    int64_t cnt = 0;
    for (int64_t row = 0; row < n_rows; row++) {
        if (ppo_line_dist(dists[2 * row], dists[2 * row + 1], x_max, y_max, slope) <= 0) {
            const int64_t i = ppo_calc_row_idx(row, n_samples);
            out_i[cnt] = i;
            out_j[cnt] = ppo_calc_col_idx(row, i, n_samples);
            cnt++;
        }
    }
    return cnt;
}

/* N1: src/boundary.cpp:97-123 generate_tuples. Returns count. */
int64_t ppo_generate_tuples(const int32_t *assignments, int64_t n_rows, int32_t within_label, int32_t self,
                            int64_t num_ref, int64_t int_offset, int64_t *out_i, int64_t *out_j) {
    const int64_t n_samples = (int64_t)(0.5 * (1 + sqrt(1 + 8.0 * (double)n_rows)));
    int64_t cnt = 0;
    for (int64_t row = 0; row < n_rows; row++) {
        if (assignments[row] == within_label) {
            int64_t i, j;
            if (self) {
                i = ppo_calc_row_idx(row, n_samples);
                j = ppo_calc_col_idx(row, i, n_samples) + int_offset;
                i = i + int_offset;
            } else {
                i = row % num_ref + int_offset;
                j = row / num_ref + num_ref + int_offset;
            }
            if (i > j) {
                int64_t t = i;
                i = j;
                j = t;
            }
            out_i[cnt] = i;
            out_j[cnt] = j;
            cnt++;
        }
    }
    return cnt;
}

/* N2: pp_sketchlib.longToSquare / squareToLong / longToSquareMulti [UPSTREAM-RECALL for the bodies; semantics
 * from the call sites PopPUNK/utils.py:393-405 (square of N = R (+ Q) samples from the condensed ref-ref vector,
 * the query-major query-ref rectangle and the condensed query-query vector), network.py:2133-2134]. */
void ppo_long_to_square(const float *vec, int64_t stride, int64_t n, float *sq) {
    for (int64_t r = 0; r < n; r++) {
        sq[r * n + r] = 0.0f;
        for (int64_t c = r + 1; c < n; c++) {
            const float v = vec[ppo_square_to_condensed(r, c, n) * stride];
            sq[r * n + c] = v;
            sq[c * n + r] = v;
        }
    }
}
void ppo_square_to_long(const float *sq, int64_t n, float *vec) {
    for (int64_t r = 0; r < n; r++)
        for (int64_t c = r + 1; c < n; c++) vec[ppo_square_to_condensed(r, c, n)] = sq[r * n + c];
}
void ppo_long_to_square_multi(const float *rr, int64_t s_rr, const float *qr, int64_t s_qr, const float *qq,
                              int64_t s_qq, int64_t R, int64_t Q, float *sq) {
    const int64_t n = R + Q;
    for (int64_t r = 0; r < n; r++) {
        sq[r * n + r] = 0.0f;
        for (int64_t c = r + 1; c < n; c++) {
            float v;
            if (c < R)
                v = rr[ppo_square_to_condensed(r, c, R) * s_rr];
            else if (r >= R)
                v = qq[ppo_square_to_condensed(r - R, c - R, Q) * s_qq];
            else
                v = qr[((c - R) * R + r) * s_qr];
            sq[r * n + c] = v;
            sq[c * n + r] = v;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * More of N1 and N3.  The sources of these ARE in the reference tree (src/boundary.cpp,
 * src/extend.cpp) and build here into oracle/_ref/libpprefine_ref.so (Makefile target `ref`); tests/ checks
 * every function below against that build, so these restatements are PINNED by the reference itself.
 * ---------------------------------------------------------------------------------------- */

/* index order of a stable ascending sort of v (boundary.hpp:27-42 sort_indexes): ties keep index order */
typedef struct {
    float v;
    int64_t idx;
} ppo_keyed;
static int ppo_keyed_cmp(const void *a, const void *b) {
    const ppo_keyed *x = (const ppo_keyed *)a, *y = (const ppo_keyed *)b;
    if (x->v < y->v) return -1;
    if (y->v < x->v) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
static void ppo_sort_indexes(const float *v, int64_t stride, int64_t n, ppo_keyed *scratch) {
    for (int64_t t = 0; t < n; t++) {
        scratch[t].v = v[t * stride];
        scratch[t].idx = t;
    }
    qsort(scratch, (size_t)n, sizeof(ppo_keyed), ppo_keyed_cmp);
}

/* N1: src/boundary.cpp:125-149 generate_all_tuples. Returns count. */
int64_t ppo_generate_all_tuples(int64_t num_ref, int64_t num_queries, int32_t self, int64_t int_offset, int64_t *out_i,
                                int64_t *out_j) {
    int64_t cnt = 0;
    if (self) {
        const double w = 2.0 * (double)num_ref - 1.0;
        const int64_t n_rows = (int64_t)((w * w - 1.0) / 8.0); /* (pow(2n-1, 2) - 1) / 8 == n(n-1)/2 */
        for (int64_t row = 0; row < n_rows; row++, cnt++) {
            int64_t i = ppo_calc_row_idx(row, num_ref);
            int64_t j = ppo_calc_col_idx(row, i, num_ref) + int_offset;
            i += int_offset;
            out_i[cnt] = i < j ? i : j;
            out_j[cnt] = i < j ? j : i;
        }
    } else {
        for (int64_t j = 0; j < num_ref; j++)
            for (int64_t i = 0; i < num_queries; i++, cnt++) {
                out_i[cnt] = i; /* int_offset is not applied on this branch (:143-147) */
                out_j[cnt] = j + num_ref;
            }
    }
    return cnt;
}

/* The boundary of one step of threshold_iterate_1D (src/boundary.cpp:171-185): the point `offset` along the line
 * (x0,y0)->(x1,y1), turned into axis intercepts.  offsets are double, everything else float: the products are
 * formed in double and narrowed on assignment, exactly as the mixed-type expressions of the reference do. */
void ppo_iterate_1d_boundary(double offset, int32_t slope, float x0, float y0, float x1, float y1, float *x_max,
                             float *y_max) {
    const float dx = x1 - x0, dy = y1 - y0;
    const float ds = sqrtf(dx * dx + dy * dy);
    const float gradient = dy / dx;
    const float xi = (float)((double)x0 + offset * (double)(dx / ds));
    const float yi = (float)((double)y0 + offset * (double)(dy / ds));
    if (slope == 2) {
        *x_max = xi + yi * gradient;
        *y_max = yi + xi / gradient;
    } else if (slope == 0) {
        *x_max = xi;
        *y_max = 0;
    } else {
        *x_max = 0;
        *y_max = yi;
    }
}

/* N1: src/boundary.cpp:151-209 threshold_iterate_1D.  Rows are ranked once by their signed distance to the FIRST
 * boundary; each later (sorted) offset then admits rows in that fixed order while line_dist <= 0.  Outputs
 * (i, j, index of the offset that admitted the row).  Returns count. */
int64_t ppo_threshold_iterate_1d(const float *dists, int64_t n_rows, const double *offsets, int64_t n_off, int32_t slope,
                                 float x0, float y0, float x1, float y1, int64_t *out_i, int64_t *out_j,
                                 int64_t *out_off) {
    if (n_rows == 0 || n_off == 0) return 0;
    const int64_t n_samples = (int64_t)(0.5 * (1 + sqrt(1 + 8.0 * (double)n_rows)));
    ppo_keyed *order = (ppo_keyed *)malloc(sizeof(ppo_keyed) * (size_t)n_rows);
    float *first = (float *)malloc(sizeof(float) * (size_t)n_rows);
    int64_t cnt = 0, pos = 0;
    for (int64_t o = 0; o < n_off; o++) {
        float x_max, y_max;
        ppo_iterate_1d_boundary(offsets[o], slope, x0, y0, x1, y1, &x_max, &y_max);
        if (o == 0) {
            for (int64_t r = 0; r < n_rows; r++) first[r] = ppo_line_dist(dists[2 * r], dists[2 * r + 1], x_max, y_max, slope);
            ppo_sort_indexes(first, 1, n_rows, order);
        }
        while (pos < n_rows) {
            const int64_t r = order[pos].idx;
            if (!(ppo_line_dist(dists[2 * r], dists[2 * r + 1], x_max, y_max, slope) <= 0)) break;
            const int64_t i = ppo_calc_row_idx(r, n_samples);
            out_i[cnt] = i;
            out_j[cnt] = ppo_calc_col_idx(r, i, n_samples);
            out_off[cnt] = o;
            cnt++;
            pos++;
        }
    }
    free(order);
    free(first);
    return cnt;
}

/* N1: src/boundary.cpp:211-237 threshold_iterate_2D: step o admits, in row order, the rows inside the sloped
 * boundary (x_max[o], y_max) that were outside the previous one.  Returns count. */
int64_t ppo_threshold_iterate_2d(const float *dists, int64_t n_rows, const float *x_max, int64_t n_off, float y_max,
                                 int64_t *out_i, int64_t *out_j, int64_t *out_off) {
    const int64_t n_samples = (int64_t)(0.5 * (1 + sqrt(1 + 8.0 * (double)n_rows)));
    int64_t cnt = 0;
    for (int64_t o = 0; o < n_off; o++)
        for (int64_t r = 0; r < n_rows; r++) {
            const float x = dists[2 * r], y = dists[2 * r + 1];
            if (ppo_line_dist(x, y, x_max[o], y_max, 2) <= 0 && (o == 0 || ppo_line_dist(x, y, x_max[o - 1], y_max, 2) > 0)) {
                const int64_t i = ppo_calc_row_idx(r, n_samples);
                out_i[cnt] = i;
                out_j[cnt] = ppo_calc_col_idx(r, i, n_samples);
                out_off[cnt] = o;
                cnt++;
            }
        }
    return cnt;
}

/* N3: src/extend.cpp:245-289 get_kNN_distances.  Row r of a dense rows x cols matrix -> its kNN smallest entries
 * (ties: lower column first), never column r itself.  Outputs are rows*kNN long, zero where a row has fewer
 * than kNN candidates (the reference's vectors are value-initialised). */
void ppo_get_knn_distances(const float *mat, int64_t rows, int64_t cols, int64_t knn, int64_t *out_i, int64_t *out_j,
                           float *out_d) {
#pragma omp parallel
    {
        ppo_keyed *order = (ppo_keyed *)malloc(sizeof(ppo_keyed) * (size_t)(cols > 0 ? cols : 1));
#pragma omp for schedule(static)
        for (int64_t r = 0; r < rows; r++) {
            ppo_sort_indexes(mat + r * cols, 1, cols, order);
            int64_t k = 0;
            for (int64_t t = 0; t < knn; t++) {
                out_i[r * knn + t] = r;
                out_j[r * knn + t] = 0;
                out_d[r * knn + t] = 0.0f;
            }
            for (int64_t t = 0; t < cols && k < knn; t++) {
                if (order[t].idx == r) continue;
                out_j[r * knn + k] = order[t].idx;
                out_d[r * knn + k] = order[t].v;
                k++;
            }
        }
        free(order);
    }
}

/* CSR row pointers of a COO list sorted by i (src/extend.cpp:14-38 row_start_indices) */
static void ppo_row_starts(const int64_t *i_vec, int64_t nnz, int64_t n, int64_t *start) {
    int64_t p = 0;
    for (int64_t r = 0; r <= n; r++) {
        while (p < nnz && i_vec[p] < r) p++;
        start[r] = p;
    }
    start[n] = nnz;
}

/* N3: src/extend.cpp:146-243 lower_rank.  Per row: neighbours in ascending distance (stable), self links dropped;
 * keep while fewer than... exactly: plain mode keeps an entry while the number already kept is <= kNN (so kNN+1
 * entries — the reference's own off-by-one, preserved); unique-distance mode counts a new distance whenever it
 * differs from the previous counted one by >= epsilon and keeps entries while that count is <= kNN.
 * reciprocal_only then keeps (i<j) entries whose mirror (j,i) was also kept.  Returns count (outputs sized nnz). */
int64_t ppo_lower_rank(const int64_t *i_vec, const int64_t *j_vec, const float *d_vec, int64_t nnz, int64_t n,
                       int64_t knn, int32_t reciprocal_only, int32_t count_unique, float epsilon, int64_t *out_i,
                       int64_t *out_j, float *out_d) {
    int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 2));
    int64_t *kept_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 2));
    ppo_keyed *order = (ppo_keyed *)malloc(sizeof(ppo_keyed) * (size_t)(nnz > 0 ? nnz : 1));
    ppo_row_starts(i_vec, nnz, n, start);
    int64_t cnt = 0;
    for (int64_t r = 0; r < n; r++) {
        kept_start[r] = cnt;
        const int64_t b = start[r], len = start[r + 1] - start[r];
        if (len <= 0) continue;
        ppo_sort_indexes(d_vec + b, 1, len, order);
        int64_t unique = 0, kept = 0;
        float prev = 0.0f;
        for (int64_t t = 0; t < len; t++) {
            const int64_t j = j_vec[b + order[t].idx];
            const float d = order[t].v;
            if (j == r) continue;
            if (count_unique) {
                if (fabsf(d - prev) >= epsilon) {
                    unique++;
                    prev = d;
                }
            } else {
                unique = kept;
            }
            if (unique > knn) break;
            out_i[cnt] = r;
            out_j[cnt] = j;
            out_d[cnt] = d;
            cnt++;
            kept++;
        }
    }
    kept_start[n] = cnt;
    if (reciprocal_only) {
        int64_t w = 0;
        int64_t *ti = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt > 0 ? cnt : 1));
        int64_t *tj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt > 0 ? cnt : 1));
        float *td = (float *)malloc(sizeof(float) * (size_t)(cnt > 0 ? cnt : 1));
        for (int64_t t = 0; t < cnt; t++) {
            const int64_t i = out_i[t], j = out_j[t];
            if (!(i < j) || j >= n) continue;
            int found = 0;
            for (int64_t u = kept_start[j]; u < kept_start[j + 1] && !found; u++) found = out_j[u] == i;
            if (found) {
                ti[w] = i;
                tj[w] = j;
                td[w] = out_d[t];
                w++;
            }
        }
        memcpy(out_i, ti, sizeof(int64_t) * (size_t)w);
        memcpy(out_j, tj, sizeof(int64_t) * (size_t)w);
        memcpy(out_d, td, sizeof(float) * (size_t)w);
        free(ti);
        free(tj);
        free(td);
        cnt = w;
    }
    free(start);
    free(kept_start);
    free(order);
    return cnt;
}

/* N3: src/extend.cpp:52-136 extend.  Sample s < nr is a reference: candidates are its sparse ref-ref row and its
 * dense row of qr (nr x nq); sample s >= nr is a query: candidates are its column of qr (the refs) and its row of
 * qq (nq x nq).  Both lists are sorted (stable) and merged, the dense-query list winning ties; j of a query
 * candidate is nr + its index; self links are skipped; the first kNN survive.  Returns count (outputs sized
 * (nr+nq)*kNN).  A row that runs out of candidates before kNN simply ends (the reference would raise there). */
int64_t ppo_extend(const int64_t *i_vec, const int64_t *j_vec, const float *d_vec, int64_t nnz, const float *qq,
                   const float *qr, int64_t nr, int64_t nq, int64_t knn, int64_t *out_i, int64_t *out_j, float *out_d) {
    int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nr + 2));
    const int64_t big = (nnz > nr ? nnz : nr) + nq + 1;
    ppo_keyed *oq = (ppo_keyed *)malloc(sizeof(ppo_keyed) * (size_t)big);
    ppo_keyed *orr = (ppo_keyed *)malloc(sizeof(ppo_keyed) * (size_t)big);
    ppo_row_starts(i_vec, nnz, nr, start);
    int64_t cnt = 0;
    for (int64_t s = 0; s < nr + nq; s++) {
        int64_t n_rr, n_qr = nq;
        if (s < nr) {
            n_rr = start[s + 1] - start[s];
            if (n_rr < 0) n_rr = 0;
            ppo_sort_indexes(qr + s * nq, 1, nq, oq);
            ppo_sort_indexes(d_vec + start[s], 1, n_rr, orr);
        } else {
            n_rr = nr;
            ppo_sort_indexes(qq + (s - nr) * nq, 1, nq, oq);
            ppo_sort_indexes(qr + (s - nr), nq, nr, orr);
        }
        int64_t a = 0, b = 0, kept = 0;
        while (a < n_qr || b < n_rr) {
            int64_t j;
            float d;
            if (b == n_rr || (a < n_qr && oq[a].v <= orr[b].v)) {
                j = oq[a].idx + nr;
                d = oq[a].v;
                a++;
            } else {
                j = s < nr ? j_vec[start[s] + orr[b].idx] : orr[b].idx;
                d = orr[b].v;
                b++;
            }
            if (j == s) continue;
            if (kept >= knn) break;
            out_i[cnt] = s;
            out_j[cnt] = j;
            out_d[cnt] = d;
            cnt++;
            kept++;
        }
    }
    free(start);
    free(oq);
    free(orr);
    return cnt;
}

#include "ppb_oracle_tuned.inc"
