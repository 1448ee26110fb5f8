// extern "C" entry points over the reference's OWN poppunk_refine sources — TEST INFRASTRUCTURE.
//
// oracle/_ref/libpprefine_ref.so = /root/reference/src/boundary.cpp + /root/reference/src/extend.cpp compiled
// unchanged from where they lie (oracle/Makefile, target `ref`) against the container stand-ins in oracle/shim/,
// plus this file, which only converts between C arrays and the reference's argument types
// (prototypes: src/boundary.hpp:42-69, src/extend.hpp:10-25).  Used by tests/ to validate the C restatement
// (oracle/ppb_oracle.c) and to generate tests/golden/; never by the product.
#include <cstdint>
#include <cstring>
#include <tuple>
#include <vector>

#include "boundary.hpp"
#include "extend.hpp"

namespace {
template <class A, class B> int64_t emit_pairs(const std::vector<std::tuple<A, B>> &e, int64_t *oi, int64_t *oj, int64_t cap) {
    for (size_t t = 0; t < e.size() && (int64_t)t < cap; t++) {
        oi[t] = (int64_t)std::get<0>(e[t]);
        oj[t] = (int64_t)std::get<1>(e[t]);
    }
    return (int64_t)e.size();
}
template <class V> int64_t emit_coo(const std::tuple<std::vector<long>, std::vector<long>, std::vector<V>> &c, int64_t *oi,
                                    int64_t *oj, V *ov, int64_t cap) {
    const auto &i = std::get<0>(c);
    const auto &j = std::get<1>(c);
    const auto &v = std::get<2>(c);
    for (size_t t = 0; t < i.size() && (int64_t)t < cap; t++) {
        oi[t] = i[t];
        oj[t] = j[t];
        ov[t] = v[t];
    }
    return (int64_t)i.size();
}
sparse_coo make_coo(const int64_t *i, const int64_t *j, const float *d, int64_t n) {
    return std::make_tuple(std::vector<long>(i, i + n), std::vector<long>(j, j + n), std::vector<float>(d, d + n));
}
}  // namespace

extern "C" {

void ppr_assign_threshold(const float *dists, int64_t n, int slope, float x_max, float y_max, int threads, float *out) {
    NumpyMatrix m(dists, n, 2);
    Eigen::VectorXf r = assign_threshold(m, slope, x_max, y_max, (unsigned)threads);
    std::memcpy(out, r.data(), sizeof(float) * (size_t)n);
}
int64_t ppr_edge_iterate(const float *dists, int64_t n, int slope, float x_max, float y_max, int64_t *oi, int64_t *oj,
                         int64_t cap) {
    return emit_pairs(edge_iterate(NumpyMatrix(dists, n, 2), slope, x_max, y_max), oi, oj, cap);
}
int64_t ppr_generate_tuples(const int32_t *assign, int64_t n, int within_label, int self, int num_ref, int int_offset,
                            int64_t *oi, int64_t *oj, int64_t cap) {
    return emit_pairs(generate_tuples(std::vector<int>(assign, assign + n), within_label, self != 0, num_ref, int_offset),
                      oi, oj, cap);
}
int64_t ppr_generate_all_tuples(int num_ref, int num_queries, int self, int int_offset, int64_t *oi, int64_t *oj,
                                int64_t cap) {
    return emit_pairs(generate_all_tuples(num_ref, num_queries, self != 0, int_offset), oi, oj, cap);
}
int64_t ppr_threshold_iterate_1d(const float *dists, int64_t n, const double *offsets, int64_t n_off, int slope, float x0,
                                 float y0, float x1, float y1, int threads, int64_t *oi, int64_t *oj, int64_t *oo,
                                 int64_t cap) {
    return emit_coo(threshold_iterate_1D(NumpyMatrix(dists, n, 2), std::vector<double>(offsets, offsets + n_off), slope,
                                         x0, y0, x1, y1, threads),
                    oi, oj, (long *)oo, cap);
}
int64_t ppr_threshold_iterate_2d(const float *dists, int64_t n, const float *x_max, int64_t n_off, float y_max,
                                 int64_t *oi, int64_t *oj, int64_t *oo, int64_t cap) {
    return emit_coo(threshold_iterate_2D(NumpyMatrix(dists, n, 2), std::vector<float>(x_max, x_max + n_off), y_max), oi,
                    oj, (long *)oo, cap);
}
int64_t ppr_get_knn_distances(const float *mat, int64_t rows, int64_t cols, int knn, int64_t dist_col, int threads,
                              int64_t *oi, int64_t *oj, float *od, int64_t cap) {
    return emit_coo(get_kNN_distances(NumpyMatrix(mat, rows, cols), knn, (size_t)dist_col, (size_t)threads), oi, oj, od,
                    cap);
}
int64_t ppr_lower_rank(const int64_t *i, const int64_t *j, const float *d, int64_t nnz, int64_t n_samples, int64_t knn,
                       int reciprocal_only, int count_unique, float epsilon, int threads, int64_t *oi, int64_t *oj,
                       float *od, int64_t cap) {
    return emit_coo(lower_rank(make_coo(i, j, d, nnz), (size_t)n_samples, (size_t)knn, reciprocal_only != 0,
                               count_unique != 0, epsilon, (size_t)threads),
                    oi, oj, od, cap);
}
int64_t ppr_extend(const int64_t *i, const int64_t *j, const float *d, int64_t nnz, const float *qq, const float *qr,
                   int64_t nr, int64_t nq, int64_t knn, int threads, int64_t *oi, int64_t *oj, float *od, int64_t cap) {
    return emit_coo(extend(make_coo(i, j, d, nnz), NumpyMatrix(qq, nq, nq), NumpyMatrix(qr, nr, nq), (size_t)knn,
                           (size_t)threads),
                    oi, oj, od, cap);
}
}
