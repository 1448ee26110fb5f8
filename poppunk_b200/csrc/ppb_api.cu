// C ABI of libppb.so (declared in include/ppb.h).  Host-side orchestration only: argument checks,
// tile scheduling, launches, and the host-buffer convenience path with overlapped copies.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "ppb_kernels.cuh"
#include "ppb_next.cuh"
#include "ppb_refine.cuh"
#include "ppb_sort.cuh"

namespace {

thread_local std::string g_err;
// A host-buffer call launches the distance kernel once per row chunk with identical fit parameters: it lends the
// launches ONE y-table buffer (filled by the first launch) instead of each launch allocating and filling its own.
struct YtabLease {
    double *buf = nullptr;
    size_t capacity = 0;   // entries
    bool filled = false;
};
thread_local YtabLease *g_ytab_lease = nullptr;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define PPB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(PPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
inline int32_t n_slices_of(int32_t ss64) { return (2 * ss64 + 31) / 32; }

// ---- index maps: src/boundary.cpp:18-37 ----------------------------------------------------
inline int64_t sq2cond(int64_t i, int64_t j, int64_t n) { return n * i - ((i * (i + 1)) >> 1) + j - 1 - i; }
inline int64_t row_idx(int64_t k, int64_t n) {
    // boundary.cpp:22-27 plus an exact integer fix-up (the double sqrt alone drifts for huge n)
    double d = std::sqrt((double)(-8 * k + 4 * n * (n - 1) - 7));
    int64_t i = n - 2 - (int64_t)std::floor(d / 2.0 - 0.5);
    i = std::max<int64_t>(0, std::min<int64_t>(i, n - 2));
    while (i > 0 && sq2cond(i, i + 1, n) > k) i--;
    while (i < n - 2 && sq2cond(i + 1, i + 2, n) <= k) i++;
    return i;
}
inline int64_t col_idx(int64_t k, int64_t i, int64_t n) {
    return k + i + 1 - n * (n - 1) / 2 + (n - i) * ((n - i) - 1) / 2;
}

// ---- tile schedule -----------------------------------------------------------------------
// Tiles are (kTI rows) x (tj columns).  Order: bands of kBand row-tiles; inside a band the column tile is
// the slow index, so the CTAs running at any moment share a handful of column tiles (L2 hits) and the
// band's row genomes stay L2-resident while the columns stream past once per band.
// Band height is chosen per call so that a band's row genomes are ~36 MB (4096 genomes at S=1024, K=5):
// measured DRAM traffic is flat between 18 and 73 MB bands (profiles/), so the smaller footprint is used.
constexpr size_t kBandBytes = (size_t)36 << 20;

struct TileKey {
    int dev;
    int64_t nA, nB;
    int self, tj;
    int64_t i_lo, i_hi;
    int band;  // row tiles per band
    bool operator<(const TileKey &o) const {
        return std::tie(dev, nA, nB, self, tj, i_lo, i_hi, band) <
               std::tie(o.dev, o.nA, o.nB, o.self, o.tj, o.i_lo, o.i_hi, o.band);
    }
};
struct TileList {
    int2 *d = nullptr;
    int64_t n = 0;
};
std::mutex g_tile_mu;
std::map<TileKey, TileList> g_tiles;

static_assert(ppb::kTI == PPB_TILE_ROWS, "include/ppb.h documents the row-tile height");
// the (row tile, column tile) list of one launch, in execution order (host only: also behind ppb_plan_tiles)
void plan_tiles(const TileKey &key, std::vector<int2> *v) {
    const int64_t nTi = (key.nA + ppb::kTI - 1) / ppb::kTI, nTj = (key.nB + key.tj - 1) / key.tj;
    const int64_t it_lo = key.i_lo / ppb::kTI, it_hi = key.i_hi / ppb::kTI;
    const int64_t kBand = std::max(1, key.band);
    for (int64_t b0 = it_lo / kBand * kBand; b0 <= it_hi && b0 < nTi; b0 += kBand) {
        const int64_t b1 = std::min<int64_t>({b0 + kBand, nTi, it_hi + 1});
        for (int64_t jt = 0; jt < nTj; jt++)
            for (int64_t ti = std::max(b0, it_lo); ti < b1; ti++) {
                if (key.self && jt * key.tj + key.tj - 1 <= ti * ppb::kTI) continue;  // no j > i in this tile
                v->push_back(make_int2((int)ti, (int)jt));
            }
    }
}

// A host-buffer call launches one kernel per row chunk: it plans the tile lists of all its chunks up front, uploads them
// with one copy and lends them to the launches of its thread, so no launch allocates, copies or synchronises.
struct TileLease {
    std::map<TileKey, TileList> lists;
};
thread_local TileLease *g_tile_lease = nullptr;

int get_tiles(const TileKey &key, cudaStream_t stream, TileList *out) {
    if (g_tile_lease) {
        auto it = g_tile_lease->lists.find(key);
        if (it != g_tile_lease->lists.end()) {
            *out = it->second;
            return PPB_OK;
        }
    }
    {
        std::lock_guard<std::mutex> lk(g_tile_mu);
        auto it = g_tiles.find(key);
        if (it != g_tiles.end()) {
            *out = it->second;
            return PPB_OK;
        }
    }
    // built outside the lock: the copy below waits for the stream, and launches on other devices must not wait with it
    std::vector<int2> v;
    plan_tiles(key, &v);
    TileList tl;
    tl.n = (int64_t)v.size();
    if (tl.n) {
        PPB_CUDA(cudaMalloc(&tl.d, v.size() * sizeof(int2)));
        PPB_CUDA(cudaMemcpyAsync(tl.d, v.data(), v.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
        PPB_CUDA(cudaStreamSynchronize(stream));  // v dies at scope exit
    }
    std::lock_guard<std::mutex> lk(g_tile_mu);
    auto it = g_tiles.find(key);
    if (it != g_tiles.end()) {  // another thread planned the same launch meanwhile
        if (tl.d) cudaFree(tl.d);
        *out = it->second;
        return PPB_OK;
    }
    // bounded cache; evict oldest first (cudaFree waits for the launches that still read the list)
    static std::vector<TileKey> order;
    if (g_tiles.size() >= 256) {
        cudaFree(g_tiles[order.front()].d);
        g_tiles.erase(order.front());
        order.erase(order.begin());
    }
    g_tiles[key] = tl;
    order.push_back(key);
    *out = tl;
    return PPB_OK;
}

// Column-tile width and band height of a launch (per device, K and sketch size; the same for every row range).
struct TileShape {
    int tj = 0, band = 0, max_smem = 0;
};
int tile_shape(int dev, int K, int sketchsize64, TileShape *out) {
    const bool wide = sketchsize64 > 1023;
    // column-tile width: the widest whose uint16 count buffer fits beside the TMA ring
    int max_smem = 0, sm_smem = 0;
    PPB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PPB_CUDA(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    auto smem_need = [&](int tjv) { return (size_t)ppb::smem_layout(K, tjv, wide).total; };
    const size_t per_cta_budget = std::min<size_t>((size_t)max_smem, (size_t)sm_smem / ppb::kCtasPerSM - 1024);
    int tj = ppb::kMaxTJ;
    while (tj > ppb::kJB && smem_need(tj) > per_cta_budget) tj >>= 1;
    if (smem_need(tj) > (size_t)max_smem) return fail(PPB_ERR_ARG, "ppb_query_dev: K too large for shared memory");
    const size_t genome_bytes = (size_t)K * n_slices_of(sketchsize64) * ppb::kSliceBytes;
    // at least 12 row tiles per band: at S = 16384 (143 KB per genome) the byte rule alone gives 3, and the column stream is
    // then re-read per band — measured on a 1/8 slice of cfg5: band 3 / 6 / 12 -> 181 / 82 / 55 GB of DRAM reads, 356.7 / 354.8 / 352.7 ms
    int band = (int)std::min<size_t>(4096, std::max<size_t>(12, kBandBytes / (genome_bytes * ppb::kTI)));
    if (const char *e = std::getenv("PPB_BAND_TILES")) band = std::max(1, atoi(e));  // tuning experiments
    out->tj = tj, out->band = band, out->max_smem = max_smem;
    return PPB_OK;
}
TileKey make_tile_key(int dev, int64_t n_ref, int64_t n_qry, int self, int64_t row_begin, int64_t row_end,
                      const TileShape &shape) {
    TileKey key{dev, self ? n_ref : n_qry, n_ref, self, shape.tj, 0, 0, shape.band};
    if (self) {
        key.i_lo = row_idx(row_begin, n_ref);
        key.i_hi = row_idx(row_end - 1, n_ref);
    } else {
        key.i_lo = row_begin / n_ref;
        key.i_hi = (row_end - 1) / n_ref;
    }
    return key;
}

int num_sms(int dev, int *out) {
    static std::mutex mu;
    static std::map<int, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(dev);
    if (it == cache.end()) {
        int n = 0;
        PPB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        it = cache.emplace(dev, n).first;
    }
    *out = it->second;
    return PPB_OK;
}

int out_row_bytes(int mode, int K) { return mode == PPB_OUT_DISTS ? 8 : 4 * K; }

// Row chunks of the host-buffer path: cut [row_begin, row_end) into pieces of at most `cap` rows that end on
// row-TILE boundaries (kTI genomes of the row side) wherever a whole row tile fits under the cap.
void plan_chunks(int64_t n_ref, int64_t n_qry, int self, int64_t row_begin, int64_t row_end, int64_t cap,
                 std::vector<std::pair<int64_t, int64_t>> *chunks) {
    const int64_t n_side = self ? n_ref : n_qry;
    const int64_t total_rows = self ? n_ref * (n_ref - 1) / 2 : n_ref * n_qry;
    // first output row of row-side genome g (self: condensed row of (g, g+1); non-self: g * n_ref)
    auto first_row_of = [&](int64_t g) -> int64_t {
        if (g >= (self ? n_side - 1 : n_side)) return total_rows;
        return self ? sq2cond(g, g + 1, n_ref) : g * n_ref;
    };
    const int64_t n_row_tiles = (n_side + ppb::kTI - 1) / ppb::kTI;
    auto tile_end = [&](int64_t t) { return std::min(row_end, first_row_of(t * ppb::kTI)); };  // monotonic in t
    for (int64_t r0 = row_begin; r0 < row_end;) {
        const int64_t g0 = self ? row_idx(r0, n_ref) : r0 / n_ref;
        int64_t lo = g0 / ppb::kTI + 1, r1;
        if (tile_end(lo) > r0 + cap) {
            r1 = std::min(row_end, r0 + cap);  // one row tile alone exceeds the cap (n > ~1M): cut inside it
        } else {
            int64_t hi = std::max(lo, n_row_tiles);  // last boundary t with tile_end(t) <= r0 + cap
            while (lo < hi) {
                const int64_t mid = lo + (hi - lo + 1) / 2;
                if (tile_end(mid) <= r0 + cap) lo = mid; else hi = mid - 1;
            }
            r1 = tile_end(lo);
        }
        chunks->emplace_back(r0, r1);
        r0 = r1;
    }
}

}  // namespace

extern "C" {

int ppb_version(void) { return PPB_VERSION; }
const char *ppb_last_error(void) { return g_err.c_str(); }
int64_t ppb_launch_count(void) { return g_launches.load(); }

int ppb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return n;
}

int64_t ppb_square_to_condensed(int64_t i, int64_t j, int64_t n) { return sq2cond(i, j, n); }
int64_t ppb_calc_row_idx(int64_t k, int64_t n) { return row_idx(k, n); }
int64_t ppb_calc_col_idx(int64_t k, int64_t i, int64_t n) { return col_idx(k, i, n); }
int64_t ppb_num_rows(int64_t n_ref, int64_t n_qry, int self) {
    return self ? n_ref * (n_ref - 1) / 2 : n_ref * n_qry;
}

size_t ppb_packed_bytes(int64_t n, int32_t K, int32_t sketchsize64) {
    return (size_t)K * n_slices_of(sketchsize64) * round_up(std::max<int64_t>(n, 1), ppb::kPad) * ppb::kSliceBytes;
}

int ppb_pack_part_dev(const uint64_t *d_sketch_part, const int64_t *d_idx, int64_t g_begin, int64_t g_end, int64_t n,
                      int32_t K, int32_t sketchsize64, uint32_t *const *d_packed, int32_t n_dst, void *stream) {
    const int64_t n_pad = round_up(std::max<int64_t>(n, 1), ppb::kPad);
    if (!d_packed || n_dst < 1 || n_dst > PPB_MAX_PEERS || n < 0 || K < 1 || K > PPB_MAX_K || sketchsize64 < 1 ||
        g_begin < 0 || g_end < g_begin || g_end > n_pad || (!d_sketch_part && g_begin < std::min(g_end, n)))
        return fail(PPB_ERR_ARG, "ppb_pack_part_dev: bad argument");
    if (g_begin == g_end) return PPB_OK;
    ppb::PackDsts dsts;
    dsts.n = n_dst;
    for (int d = 0; d < n_dst; d++) {
        if (!d_packed[d]) return fail(PPB_ERR_ARG, "ppb_pack_part_dev: null destination");
        dsts.p[d] = d_packed[d];
    }
    const int64_t units = (int64_t)K * n_slices_of(sketchsize64) * (g_end - g_begin);  // one warp per 1792-byte unit
    const int64_t blocks = std::min<int64_t>((units + ppb::kPackWarps - 1) / ppb::kPackWarps, 148 * 32);
    ppb::pack_kernel<<<(unsigned)blocks, ppb::kPackWarps * 32, 0, (cudaStream_t)stream>>>(
        d_sketch_part, d_idx, g_begin, g_end, n, n_pad, K, sketchsize64, n_slices_of(sketchsize64), dsts);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_pack_dev(const uint64_t *d_sketch, int64_t n_src, const int64_t *d_idx, int64_t n, int32_t K,
                 int32_t sketchsize64, uint32_t *d_packed, void *stream) {
    if (!d_sketch || !d_packed || n < 0 || K < 1 || K > PPB_MAX_K || sketchsize64 < 1)
        return fail(PPB_ERR_ARG, "ppb_pack_dev: bad argument");
    if (!d_idx && n != n_src) return fail(PPB_ERR_ARG, "ppb_pack_dev: n != n_src without an index list");
    const int64_t n_pad = round_up(std::max<int64_t>(n, 1), ppb::kPad);
    return ppb_pack_part_dev(d_sketch, d_idx, 0, n_pad, n, K, sketchsize64, &d_packed, 1, stream);
}

static int query_dev_impl(const uint32_t *d_ref_packed, int64_t n_ref, const uint32_t *d_qry_packed, int64_t n_qry,
                          const int32_t *kmers, int32_t K, int32_t sketchsize64, const float *d_rand_table,
                          int32_t n_clusters, const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                          int64_t row_begin, int64_t row_end, int32_t out_mode, void *d_out,
                          const ppb_boundary *boundary, int8_t *d_labels, unsigned long long *d_n_degenerate,
                          void *stream, void *const *d_peer_out, int32_t n_peers, void *d_mc_out,
                          int32_t edge_mode = 0, int64_t *d_edge_rows = nullptr, int64_t edge_cap = 0,
                          unsigned long long *d_edge_count = nullptr) {
    const int self = d_qry_packed == nullptr;
    if (!d_ref_packed || !kmers || K < 1 || K > PPB_MAX_K || n_ref < 0 || (!self && n_qry < 0))
        return fail(PPB_ERR_ARG, "ppb_query_dev: bad argument");
    if (sketchsize64 < 1 || sketchsize64 > (1 << 24))
        return fail(PPB_ERR_ARG, "ppb_query_dev: sketchsize64 must be in [1, 2^24]");
    const bool wide = sketchsize64 > 1023;   // more than 65535 bins: per-k counts no longer fit uint16 (PopPUNK accepts
                                             // --sketch-size up to 10^6 bins, __main__.py:310)
    if (out_mode < PPB_OUT_DISTS || out_mode > PPB_OUT_COUNTS) return fail(PPB_ERR_ARG, "ppb_query_dev: bad out_mode");
    {   // an empty row range is a no-op whatever the buffers are (torch hands out null pointers for empty tensors, and
        // a rank with an empty shard must still reach the collectives that follow)
        const int64_t all_rows = ppb_num_rows(n_ref, n_qry, self);
        if (row_begin >= 0 && row_begin == row_end && row_end <= all_rows) return PPB_OK;
    }
    const bool has_peers = (n_peers > 0 && d_peer_out) || d_mc_out;
    if (n_peers < 0 || n_peers > PPB_MAX_PEERS) return fail(PPB_ERR_ARG, "ppb_query_dev_fused: bad n_peers");
    if (has_peers && out_mode != PPB_OUT_DISTS) return fail(PPB_ERR_ARG, "ppb_query_dev_fused: PPB_OUT_DISTS only");
    const bool has_edges = edge_mode != 0 && d_edge_rows && d_edge_count && boundary;
    if (edge_mode != 0 && !has_edges) return fail(PPB_ERR_ARG, "ppb_query_edges_dev: edge buffers and a boundary are required");
    if (!d_out && !has_peers && !has_edges && !(out_mode == PPB_OUT_DISTS && boundary && d_labels))
        return fail(PPB_ERR_ARG, "ppb_query_dev: no output buffer");
    if (!has_edges && (boundary != nullptr) != (d_labels != nullptr))
        return fail(PPB_ERR_ARG, "ppb_query_dev: boundary and d_labels go together");
    if (d_labels && !boundary) return fail(PPB_ERR_ARG, "ppb_query_dev: labels need a boundary");
    if (boundary && out_mode != PPB_OUT_DISTS) return fail(PPB_ERR_ARG, "ppb_query_dev: labels need PPB_OUT_DISTS");
    if (boundary && (boundary->slope < 0 || boundary->slope > 2)) return fail(PPB_ERR_ARG, "ppb_query_dev: bad slope");
    if (d_rand_table && (n_clusters < 1 || !d_ref_cluster || (!self && !d_qry_cluster)))
        return fail(PPB_ERR_ARG, "ppb_query_dev: random table without cluster ids");
    for (int t = 1; t < K; t++)
        if (kmers[t] <= kmers[t - 1]) return fail(PPB_ERR_ARG, "ppb_query_dev: kmers must be ascending");
    const int64_t total_rows = ppb_num_rows(n_ref, n_qry, self);
    if (row_begin < 0 || row_end > total_rows || row_begin > row_end)
        return fail(PPB_ERR_ARG, "ppb_query_dev: bad row range");
    if (row_begin == row_end) return PPB_OK;

    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    PPB_CUDA(cudaGetDevice(&dev));

    ppb::QueryParams p;
    std::memset(&p, 0, sizeof(p));
    p.self = self;
    p.A = self ? d_ref_packed : d_qry_packed;
    p.B = d_ref_packed;
    p.nA = self ? n_ref : n_qry;
    p.nB = n_ref;
    p.nA_pad = round_up(std::max<int64_t>(p.nA, 1), ppb::kPad);
    p.nB_pad = round_up(std::max<int64_t>(p.nB, 1), ppb::kPad);
    p.K = K;
    p.n_slices = n_slices_of(sketchsize64);
    p.KS = K * p.n_slices;
    p.G32 = 2 * sketchsize64;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.out_mode = out_mode;
    p.out = d_out;
    p.labels = d_labels;
    p.mc_out = d_mc_out;
    p.edge_mode = has_edges ? edge_mode : 0;
    p.edge_rows = (long long *)d_edge_rows;
    p.edge_cap = edge_cap;
    p.edge_count = d_edge_count;
    if (!d_mc_out && d_peer_out)
        for (int g = 0; g < n_peers; g++) p.peer_out[p.n_peer_out++] = d_peer_out[g];
    p.has_boundary = boundary != nullptr;
    if (boundary) p.bnd = *boundary;
    p.rand_table = d_rand_table;
    p.C = n_clusters;
    p.clB = d_ref_cluster;
    p.clA = self ? d_ref_cluster : d_qry_cluster;
    p.n_degenerate = d_n_degenerate;
    p.S = 64.0 * sketchsize64;
    p.inv_S = 1.0 / p.S;
    p.S_pow2 = (sketchsize64 & (sketchsize64 - 1)) == 0;
    p.tol = (double)PPB_MIN_JACCARD_BINS / p.S;
    // OLS constants for every possible series length n (the series is always a prefix of kmers)
    double sx = 0;
    for (int t = 0; t < K; t++) {
        p.x[t] = (double)kmers[t];
        sx += p.x[t];
        const int n = t + 1;
        p.xbar[n] = sx / n;
        double sxx = 0;
        for (int u = 0; u < n; u++) sxx += (p.x[u] - p.xbar[n]) * (p.x[u] - p.xbar[n]);
        p.inv_sxx[n] = n >= 2 ? 1.0 / sxx : 0.0;
        p.inv_n[n] = 1.0 / n;
    }

    TileShape shape;
    if (int rc = tile_shape(dev, K, sketchsize64, &shape)) return rc;
    const int tj = shape.tj;
    const int max_smem = shape.max_smem;
    auto smem_need = [&](int tjv) { return (size_t)ppb::smem_layout(K, tjv, wide).total; };
    p.tj = tj;
    p.wide = wide;

    // L2 eviction-priority knobs, all OFF by default: measured on B200 at N=100k (profiles/), streaming stores,
    // evict_last row-genome loads (with a persisting set-aside) and evict_first column TMA each INCREASED the
    // DRAM read traffic (50 -> 57..208 GB) and changed the kernel time by < 2 %.  Kept as environment knobs.
    p.stream_stores = 0;
    p.a_policy = 0;
    p.b_policy = 0;
    if (const char *e = std::getenv("PPB_STREAM_STORES")) p.stream_stores = atoi(e);
    if (const char *e = std::getenv("PPB_A_POLICY")) p.a_policy = atoi(e);
    if (const char *e = std::getenv("PPB_B_POLICY")) p.b_policy = atoi(e);
    if (const char *e = std::getenv("PPB_DEBUG_SKIP_EPILOGUE")) p.debug_skip_epilogue = atoi(e);

    const TileKey key = make_tile_key(dev, n_ref, n_qry, self, row_begin, row_end, shape);
    TileList tl;
    if (int rc = get_tiles(key, st, &tl)) return rc;
    if (tl.n == 0) return PPB_OK;
    p.tiles = tl.d;
    p.n_tiles = tl.n;

    int sms = 0;
    if (int rc = num_sms(dev, &sms)) return rc;
    const size_t smem = smem_need(tj);
    const bool single = p.n_slices == 1;
    static std::mutex attr_mu;
    {   // function attributes are per device and sticky: set them once per device, not on every launch
        static std::map<int, bool> attr_done;
        std::lock_guard<std::mutex> lk(attr_mu);
        if (!attr_done[dev]) {
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            PPB_CUDA(cudaFuncSetAttribute(ppb::query_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_done[dev] = true;
        }
    }
    // y-table for the fused fit (stream-ordered scratch; skipped when it would be unreasonably large)
    double *d_ytab = nullptr;
    if (out_mode == PPB_OUT_DISTS && !wide) {   // (the fused fit reads uint16 counts: huge sketches take the generic epilogue)
        const size_t entries = (size_t)(d_rand_table ? (size_t)n_clusters * n_clusters : 1) * K * ((size_t)p.S + 1);
        YtabLease *lease = g_ytab_lease;
        if (lease && lease->buf && lease->capacity >= entries) {
            if (!lease->filled) {
                const int threads = 256;
                const unsigned blocks = (unsigned)std::min<size_t>((entries + threads - 1) / threads, (size_t)sms * 8);
                ppb::ytab_kernel<<<blocks, threads, 0, st>>>(p, lease->buf);
                g_launches++;
                PPB_CUDA(cudaGetLastError());
                lease->filled = true;
            }
            p.ytab = lease->buf;
        } else if (entries * sizeof(double) <= ((size_t)64 << 20) && !std::getenv("PPB_NO_YTAB")) {
            PPB_CUDA(cudaMallocAsync(&d_ytab, entries * sizeof(double), st));
            const int threads = 256;
            const unsigned blocks = (unsigned)std::min<size_t>((entries + threads - 1) / threads, (size_t)sms * 8);
            ppb::ytab_kernel<<<blocks, threads, 0, st>>>(p, d_ytab);
            g_launches++;
            PPB_CUDA(cudaGetLastError());
            p.ytab = d_ytab;
        }
    }
    const unsigned grid = (unsigned)std::min<int64_t>(tl.n, (int64_t)ppb::kCtasPerSM * sms);
    if (single)
        ppb::query_kernel<0><<<grid, ppb::kThreads, smem, st>>>(p);
    else if (!wide)
        ppb::query_kernel<1><<<grid, ppb::kThreads, smem, st>>>(p);
    else
        ppb::query_kernel<2><<<grid, ppb::kThreads, smem, st>>>(p);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    if (d_ytab) PPB_CUDA(cudaFreeAsync(d_ytab, st));
    return PPB_OK;
}

int ppb_query_dev(const uint32_t *d_ref_packed, int64_t n_ref, const uint32_t *d_qry_packed, int64_t n_qry,
                  const int32_t *kmers, int32_t K, int32_t sketchsize64, const float *d_rand_table,
                  int32_t n_clusters, const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                  int64_t row_begin, int64_t row_end, int32_t out_mode, void *d_out,
                  const ppb_boundary *boundary, int8_t *d_labels, unsigned long long *d_n_degenerate,
                  void *stream) {
    return query_dev_impl(d_ref_packed, n_ref, d_qry_packed, n_qry, kmers, K, sketchsize64, d_rand_table, n_clusters,
                          d_ref_cluster, d_qry_cluster, row_begin, row_end, out_mode, d_out, boundary, d_labels,
                          d_n_degenerate, stream, nullptr, 0, nullptr);
}

int ppb_query_dev_fused(const uint32_t *d_ref_packed, int64_t n_ref, const uint32_t *d_qry_packed, int64_t n_qry,
                        const int32_t *kmers, int32_t K, int32_t sketchsize64, const float *d_rand_table,
                        int32_t n_clusters, const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                        int64_t row_begin, int64_t row_end, void *d_out, void *const *d_peer_out, int32_t n_peers,
                        void *d_mc_out, unsigned long long *d_n_degenerate, void *stream) {
    if ((!d_peer_out || n_peers < 1) && !d_mc_out) return fail(PPB_ERR_ARG, "ppb_query_dev_fused: no peer buffers");
    return query_dev_impl(d_ref_packed, n_ref, d_qry_packed, n_qry, kmers, K, sketchsize64, d_rand_table, n_clusters,
                          d_ref_cluster, d_qry_cluster, row_begin, row_end, PPB_OUT_DISTS, d_out, nullptr, nullptr,
                          d_n_degenerate, stream, d_peer_out, n_peers, d_mc_out);
}

int ppb_query_edges_dev(const uint32_t *d_ref_packed, int64_t n_ref, const uint32_t *d_qry_packed, int64_t n_qry,
                        const int32_t *kmers, int32_t K, int32_t sketchsize64, const float *d_rand_table,
                        int32_t n_clusters, const uint16_t *d_ref_cluster, const uint16_t *d_qry_cluster,
                        int64_t row_begin, int64_t row_end, const ppb_boundary *boundary, int32_t include_boundary,
                        int64_t *d_edge_rows, int64_t capacity, unsigned long long *d_edge_count, void *d_out,
                        int8_t *d_labels, unsigned long long *d_n_degenerate, void *stream) {
    if (!boundary || !d_edge_rows || !d_edge_count || capacity < 0)
        return fail(PPB_ERR_ARG, "ppb_query_edges_dev: bad argument");
    return query_dev_impl(d_ref_packed, n_ref, d_qry_packed, n_qry, kmers, K, sketchsize64, d_rand_table, n_clusters,
                          d_ref_cluster, d_qry_cluster, row_begin, row_end, PPB_OUT_DISTS, d_out, boundary, d_labels,
                          d_n_degenerate, stream, nullptr, 0, nullptr, include_boundary ? 2 : 1, d_edge_rows, capacity,
                          d_edge_count);
}

// ---------------------------------------------------------------------------------------------------------
// N1 / N2 entry points
// ---------------------------------------------------------------------------------------------------------
extern "C++" {
namespace {
inline unsigned grid_for(int64_t n, int threads, int cap) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, cap));
}
template <typename Pred>
int run_select(Pred pred, int64_t n_rows, ppb::PairMap map, int64_t *d_i, int64_t *d_j, int64_t capacity,
               int64_t *d_count, void *d_scratch, cudaStream_t st) {
    if (!d_count || !d_scratch || capacity < 0 || (capacity > 0 && (!d_i || !d_j)))
        return fail(PPB_ERR_ARG, "edge compaction: bad argument");
    if (n_rows == 0) {
        PPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
        return PPB_OK;
    }
    // the input is read once: mark (bit per row + count per 1024-row unit) -> scan -> emit from the bit mask
    const int64_t units = (n_rows + ppb::kUnitRows - 1) / ppb::kUnitRows;
    uint32_t *bits = (uint32_t *)d_scratch;
    uint32_t *unit_count = bits + units * 32;
    int64_t *unit_off = (int64_t *)(unit_count + ((units + 1) & ~(int64_t)1));
    const unsigned grid = (unsigned)std::min<int64_t>((units + 7) / 8, 148 * 16);
    ppb::select_mark_kernel<Pred><<<grid, 256, 0, st>>>(pred, n_rows, bits, unit_count);
    ppb::unit_scan_kernel<<<1, 1024, 0, st>>>(unit_count, units, unit_off, d_count);
    ppb::select_emit_kernel<ppb::OutPairs><<<grid, 256, 0, st>>>(bits, unit_count, unit_off, n_rows, ppb::OutPairs{map, d_i, d_j}, capacity);
    g_launches += 2;
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}
}  // namespace
}  // extern "C++"

size_t ppb_edges_scratch_bytes(int64_t n_rows) {
    const size_t units = (size_t)((std::max<int64_t>(n_rows, 1) + ppb::kUnitRows - 1) / ppb::kUnitRows);
    return units * 128 + (units + 2) * 4 + (units + 1) * 8;   // bit mask + unit counts + unit offsets
}

int ppb_rows_to_pairs_dev(const int64_t *d_rows, int64_t n, int32_t self, int64_t n_samples_or_num_ref,
                          int64_t int_offset, int64_t *d_i, int64_t *d_j, void *stream) {
    if (n < 0 || (n > 0 && (!d_rows || !d_i || !d_j)) || n_samples_or_num_ref < 1)
        return fail(PPB_ERR_ARG, "ppb_rows_to_pairs_dev: bad argument");
    if (n == 0) return PPB_OK;
    ppb::rows_to_pairs_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        d_rows, n, ppb::PairMap{self, n_samples_or_num_ref, int_offset}, d_i, d_j);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_edges_from_dists_dev(const float *d_dists, int64_t n_rows, int64_t n_samples, int32_t slope, float x_max,
                             float y_max, int64_t *d_i, int64_t *d_j, int64_t capacity, int64_t *d_count,
                             void *d_scratch, void *stream) {
    if (n_rows < 0 || (n_rows > 0 && !d_dists) || slope < 0 || slope > 2 || n_samples < 2)
        return fail(PPB_ERR_ARG, "ppb_edges_from_dists_dev: bad argument");
    ppb::PredDists pred{reinterpret_cast<const float2 *>(d_dists), slope, x_max, y_max};
    return run_select(pred, n_rows, ppb::PairMap{1, n_samples, 0}, d_i, d_j, capacity, d_count, d_scratch,
                      (cudaStream_t)stream);
}

int ppb_edges_from_labels_dev(const void *d_labels, int32_t label_dtype, int64_t n_rows, int32_t within_label,
                              int32_t self, int64_t n_samples_or_num_ref, int64_t int_offset, int64_t *d_i,
                              int64_t *d_j, int64_t capacity, int64_t *d_count, void *d_scratch, void *stream) {
    if (n_rows < 0 || (n_rows > 0 && !d_labels) || n_samples_or_num_ref < 1)
        return fail(PPB_ERR_ARG, "ppb_edges_from_labels_dev: bad argument");
    const ppb::PairMap map{self, n_samples_or_num_ref, int_offset};
    cudaStream_t st = (cudaStream_t)stream;
    switch (label_dtype) {
        case 0: return run_select(ppb::PredLabels<int8_t>{(const int8_t *)d_labels, within_label}, n_rows, map, d_i, d_j, capacity, d_count, d_scratch, st);
        case 1: return run_select(ppb::PredLabels<int32_t>{(const int32_t *)d_labels, within_label}, n_rows, map, d_i, d_j, capacity, d_count, d_scratch, st);
        case 2: return run_select(ppb::PredLabels<float>{(const float *)d_labels, within_label}, n_rows, map, d_i, d_j, capacity, d_count, d_scratch, st);
        default: return fail(PPB_ERR_ARG, "ppb_edges_from_labels_dev: label_dtype must be 0 (int8), 1 (int32) or 2 (float32)");
    }
}

int ppb_long_to_square_dev(const float *d_vec, int64_t stride, int64_t n, float *d_square, void *stream) {
    if (n < 0 || stride < 1 || (n > 0 && !d_square) || (n > 1 && !d_vec))
        return fail(PPB_ERR_ARG, "ppb_long_to_square_dev: bad argument");
    if (n == 0) return PPB_OK;
    {
        const int64_t nt = (n + ppb::kSqTile - 1) / ppb::kSqTile;
        ppb::long_to_square_kernel<<<grid_for(nt * (nt + 1) / 2, 1, 148 * 64), ppb::kSqTile * 8, 0, (cudaStream_t)stream>>>(d_vec, stride, n, d_square);
    }
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_square_to_long_dev(const float *d_square, int64_t n, float *d_vec, void *stream) {
    if (n < 0 || (n > 1 && (!d_square || !d_vec))) return fail(PPB_ERR_ARG, "ppb_square_to_long_dev: bad argument");
    if (n < 2) return PPB_OK;
    ppb::square_to_long_kernel<<<grid_for(n - 1, 1, 148 * 32), 256, 0, (cudaStream_t)stream>>>(d_square, n, d_vec);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_long_to_square_multi_dev(const float *d_rr, int64_t stride_rr, const float *d_qr, int64_t stride_qr,
                                 const float *d_qq, int64_t stride_qq, int64_t n_ref, int64_t n_qry, float *d_square,
                                 void *stream) {
    if (n_ref < 0 || n_qry < 0 || !d_square || stride_rr < 1 || stride_qr < 1 || stride_qq < 1 ||
        (n_ref > 1 && !d_rr) || (n_ref > 0 && n_qry > 0 && !d_qr) || (n_qry > 1 && !d_qq))
        return fail(PPB_ERR_ARG, "ppb_long_to_square_multi_dev: bad argument");
    const int64_t n = n_ref + n_qry;
    if (n == 0) return PPB_OK;
    ppb::long_to_square_multi_kernel<<<grid_for(n * n, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        d_rr, stride_rr, d_qr, stride_qr, d_qq, stride_qq, n_ref, n_qry, d_square);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_assign_threshold_dev(const float *d_dists, int64_t n, int32_t slope, float x_max, float y_max,
                             float *d_out, void *stream) {
    if (n < 0 || (n > 0 && (!d_dists || !d_out)) || slope < 0 || slope > 2)
        return fail(PPB_ERR_ARG, "ppb_assign_threshold_dev: bad argument");
    if (n == 0) return PPB_OK;
    const int threads = 256;
    const int64_t blocks = std::min<int64_t>((n + threads - 1) / threads, 148 * 32);
    ppb::threshold_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(d_dists), n, slope, x_max, y_max, d_out);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// N1 (rest) / N3 entry points: threshold iteration, all tuples, kNN / lowerRank / extend
// ---------------------------------------------------------------------------------------------------------
extern "C++" {
namespace {
// The stream-ordered allocator returns freed memory to the OS at every synchronisation unless told otherwise, and the
// functions below take gigabytes of scratch per call (threshold iteration: ~25 B per row): keep it cached in the
// device's default pool between calls.  ppb_release_workspace() trims the pool.
void keep_scratch_cached() {
    static std::mutex mu;
    static std::map<int, bool> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done[dev] = true;
}
// stream-ordered scratch that frees itself (also on the error paths)
struct Scratch {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) { keep_scratch_cached(); }
    ~Scratch() {
        for (void *q : ptrs) cudaFreeAsync(q, st);
    }
    template <typename T> int get(T **out, size_t count) {
        void *q = nullptr;
        if (cudaMallocAsync(&q, std::max<size_t>(count * sizeof(T), 256), st) != cudaSuccess) {
            cudaGetLastError();
            return fail(PPB_ERR_NOMEM, "cudaMallocAsync failed for " + std::to_string(count * sizeof(T)) + " bytes");
        }
        ptrs.push_back(q);
        *out = (T *)q;
        return PPB_OK;
    }
};
// boundary of one step of threshold_iterate_1D (src/boundary.cpp:171-185): offsets are double, the rest float;
// mixed expressions are evaluated in double and narrowed on assignment, as the reference's are
void iterate_1d_boundary(double offset, int slope, float x0, float y0, float x1, float y1, float *x_max, float *y_max) {
    const float dx = x1 - x0, dy = y1 - y0;
    const float ds = std::sqrt(dx * dx + dy * dy);
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
// bisection over the boundaries is allowed when they move outward: positive, non-decreasing intercepts (csrc/ppb_next.cuh)
ppb::StepSearch make_step_search(const float2 *d_step, const float2 *h_step, int n_off, int slope) {
    ppb::StepSearch S;
    S.step = d_step;
    S.n_off = n_off;
    S.slope = slope;
    S.bisect = n_off >= 2 && !std::getenv("PPB_NO_BISECT");
    S.XM = S.YM = 0.0f;
    for (int o = 0; o < n_off; o++) {
        const float x = h_step[o].x, y = h_step[o].y;
        S.XM = std::max(S.XM, std::fabs(x));
        S.YM = std::max(S.YM, std::fabs(y));
        if (slope == 2 && !(x > 0.0f && y > 0.0f)) S.bisect = 0;            // degenerate / inward boundary: full scan
        if (o > 0) {
            if (slope != 1 && !(x >= h_step[o - 1].x)) S.bisect = 0;
            if (slope != 0 && !(y >= h_step[o - 1].y)) S.bisect = 0;
        }
    }
    return S;
}

// One pass of the least-significant-digit radix sort (8-bit digit at `shift`): count per chunk, scans, stable emit.
// hist: int64 [chunks][256], totals / base: int64 [256].
template <typename Digit, typename Move>
int radix_pass(Digit dg, Move mv, int64_t n, int64_t *hist, int64_t *totals, int64_t *base, cudaStream_t st) {
    const int64_t chunks = (n + ppb::kSelBlockRows - 1) / ppb::kSelBlockRows;
    ppb::bucket_count_kernel<Digit><<<(unsigned)chunks, ppb::kSelThreads, 0, st>>>(dg, n, 256, hist);
    ppb::bucket_scan_chunks_kernel<<<256, 1024, 0, st>>>(hist, chunks, 256, totals);
    ppb::bucket_scan_totals_kernel<<<1, ppb::kBktMax, 0, st>>>(totals, 256, base, nullptr);
    ppb::bucket_emit_kernel<Digit, Move><<<(unsigned)chunks, ppb::kSelThreads, 0, st>>>(dg, mv, n, 256, hist, base, INT64_MAX);
    g_launches += 4;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}
}  // namespace
}  // extern "C++"

int ppb_sort_rows_dev(int64_t *d_rows, int64_t n, int64_t max_row, void *stream) {
    if (n < 0 || max_row < 0 || (n > 0 && !d_rows)) return fail(PPB_ERR_ARG, "ppb_sort_rows_dev: bad argument");
    if (n < 2) return PPB_OK;
    if (n > ((int64_t)1 << 42)) return fail(PPB_ERR_ARG, "ppb_sort_rows_dev: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st);
    unsigned long long *tmp = nullptr;
    int64_t *hist = nullptr, *tb = nullptr;
    const int64_t chunks = (n + ppb::kSelBlockRows - 1) / ppb::kSelBlockRows;
    if (int rc = sc.get(&tmp, (size_t)n)) return rc;
    if (int rc = sc.get(&hist, (size_t)chunks * 256)) return rc;
    if (int rc = sc.get(&tb, 512)) return rc;
    int bits = 1;
    while (bits < 63 && (max_row >> bits)) bits++;
    int passes = (bits + 7) / 8;
    if (passes & 1) passes++;  // an even number of passes leaves the result in the caller's buffer
    unsigned long long *a = (unsigned long long *)d_rows, *b = tmp;
    for (int ps = 0; ps < passes; ps++) {
        if (int rc = radix_pass(ppb::DigitOfU64{a, ps * 8}, ppb::MoveKey64{a, b}, n, hist, tb, tb + 256, st)) return rc;
        std::swap(a, b);
    }
    return PPB_OK;
}

int ppb_generate_all_tuples_dev(int64_t num_ref, int64_t num_queries, int32_t self, int64_t int_offset, int64_t *d_i,
                                int64_t *d_j, void *stream) {
    if (num_ref < 0 || num_queries < 0) return fail(PPB_ERR_ARG, "ppb_generate_all_tuples_dev: bad argument");
    const int64_t total = self ? num_ref * (num_ref - 1) / 2 : num_ref * num_queries;
    if (total == 0) return PPB_OK;
    if (!d_i || !d_j) return fail(PPB_ERR_ARG, "ppb_generate_all_tuples_dev: bad argument");
    ppb::all_tuples_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(num_ref, num_queries, self,
                                                                                           int_offset, total, d_i, d_j);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_threshold_iterate_2d_dev(const float *d_dists, int64_t n_rows, const float *x_max, int32_t n_off, float y_max,
                                 int64_t *d_i, int64_t *d_j, int64_t *d_off, int64_t capacity, int64_t *d_count,
                                 void *stream) {
    if (n_rows < 0 || n_off < 0 || (n_rows > 0 && !d_dists) || (n_off > 0 && !x_max) || !d_count || capacity < 0 ||
        (capacity > 0 && (!d_i || !d_j || !d_off)))
        return fail(PPB_ERR_ARG, "ppb_threshold_iterate_2d_dev: bad argument");
    if (n_off > ppb::kIterMaxOffsets) return fail(PPB_ERR_ARG, "ppb_threshold_iterate_2d_dev: at most 1024 offsets per call");
    for (int o = 1; o < n_off; o++)  // python_bindings.cpp:70-73
        if (x_max[o] < x_max[o - 1]) return fail(PPB_ERR_ARG, "x_max range to thresholdIterate2D must be sorted");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0 || n_off == 0) {
        PPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
        return PPB_OK;
    }
    const int64_t blocks = (n_rows + ppb::kSelBlockRows - 1) / ppb::kSelBlockRows;
    if (blocks > 0x7fffffff) return fail(PPB_ERR_ARG, "ppb_threshold_iterate_2d_dev: too many rows for one call");
    Scratch sc(st);
    float *d_x = nullptr;
    int64_t *cnt = nullptr;
    if (int rc = sc.get(&d_x, (size_t)n_off)) return rc;
    if (n_off <= 64 && !std::getenv("PPB_ITERATE2D_GENERIC")) {
        // what PopPUNK asks for (20-40 steps, refine.py:116-123,190-191): ONE read of the distances.  The classify
        // pass leaves a byte per row (its admitting step), a stable bucket scatter by step places the rows.
        uint8_t *note = nullptr;
        int64_t *hist = nullptr, *tb = nullptr;
        float2 *d_step = nullptr;
        if (int rc = sc.get(&note, (size_t)n_rows)) return rc;
        if (int rc = sc.get(&hist, (size_t)blocks * n_off)) return rc;
        if (int rc = sc.get(&tb, 2 * ppb::kBktMax)) return rc;
        if (int rc = sc.get(&d_step, 64)) return rc;
        float2 h_step[64];
        for (int o = 0; o < n_off; o++) h_step[o] = make_float2(x_max[o], x_max[o] * y_max);  // float product, as line_dist forms it
        PPB_CUDA(cudaMemcpyAsync(d_step, h_step, sizeof(float2) * n_off, cudaMemcpyHostToDevice, st));
        float2 *d_xy = nullptr, h_xy[64];
        if (int rc = sc.get(&d_xy, 64)) return rc;
        for (int o = 0; o < n_off; o++) h_xy[o] = make_float2(x_max[o], y_max);
        PPB_CUDA(cudaMemcpyAsync(d_xy, h_xy, sizeof(float2) * n_off, cudaMemcpyHostToDevice, st));
        PPB_CUDA(cudaStreamSynchronize(st));  // h_step / h_xy are on this frame
        const ppb::Iter2dClass cls{reinterpret_cast<const float2 *>(d_dists), d_step, n_off, y_max, note,
                                   make_step_search(d_xy, h_xy, n_off, 2)};
        const int64_t n_samples = (int64_t)(0.5 * (1.0 + std::sqrt(1.0 + 8.0 * (double)n_rows)));
        const ppb::Iter2dEmit em{n_samples, d_i, d_j, d_off};
        ppb::bucket_count_kernel<ppb::Iter2dClass><<<(unsigned)blocks, ppb::kSelThreads, 0, st>>>(cls, n_rows, n_off, hist);
        ppb::bucket_scan_chunks_kernel<<<(unsigned)n_off, 1024, 0, st>>>(hist, blocks, n_off, tb);
        ppb::bucket_scan_totals_kernel<<<1, ppb::kBktMax, 0, st>>>(tb, n_off, tb + ppb::kBktMax, d_count);
        ppb::bucket_emit_kernel<ppb::Iter2dClass, ppb::Iter2dEmit><<<(unsigned)blocks, ppb::kSelThreads, 0, st>>>(
            cls, em, n_rows, n_off, hist, tb + ppb::kBktMax, capacity);
        g_launches += 4;
        PPB_CUDA(cudaGetLastError());
        return PPB_OK;
    }
    if (int rc = sc.get(&cnt, (size_t)n_off * blocks)) return rc;
    PPB_CUDA(cudaMemcpyAsync(d_x, x_max, sizeof(float) * n_off, cudaMemcpyHostToDevice, st));
    const float2 *d2 = reinterpret_cast<const float2 *>(d_dists);
    ppb::iterate2d_kernel<<<(unsigned)blocks, ppb::kSelThreads, 0, st>>>(d2, n_rows, d_x, n_off, y_max, 0, cnt, blocks,
                                                                         capacity, d_i, d_j, d_off);
    ppb::scan_kernel<<<1, 1024, 0, st>>>(cnt, (int64_t)n_off * blocks, d_count);
    ppb::iterate2d_kernel<<<(unsigned)blocks, ppb::kSelThreads, 0, st>>>(d2, n_rows, d_x, n_off, y_max, 1, cnt, blocks,
                                                                         capacity, d_i, d_j, d_off);
    g_launches += 3;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_threshold_iterate_1d_dev(const float *d_dists, int64_t n_rows, const double *offsets, int32_t n_off, int32_t slope,
                                 float x0, float y0, float x1, float y1, int64_t *d_i, int64_t *d_j, int64_t *d_off,
                                 int64_t capacity, int64_t *d_count, void *stream) {
    if (n_rows < 0 || n_off < 0 || (n_rows > 0 && !d_dists) || (n_off > 0 && !offsets) || !d_count || capacity < 0 ||
        (capacity > 0 && (!d_i || !d_j || !d_off)) || slope < 0 || slope > 2)
        return fail(PPB_ERR_ARG, "ppb_threshold_iterate_1d_dev: bad argument");
    for (int o = 1; o < n_off; o++)  // python_bindings.cpp:56-58
        if (offsets[o] < offsets[o - 1]) return fail(PPB_ERR_ARG, "Offsets to thresholdIterate1D must be sorted");
    cudaStream_t st = (cudaStream_t)stream;
    PPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (n_rows == 0 || n_off == 0) return PPB_OK;
    if (n_rows > ((int64_t)1 << 40)) return fail(PPB_ERR_ARG, "ppb_threshold_iterate_1d_dev: too many rows");
    std::vector<float2> bnd((size_t)n_off);
    for (int o = 0; o < n_off; o++) iterate_1d_boundary(offsets[o], slope, x0, y0, x1, y1, &bnd[o].x, &bnd[o].y);
    Scratch sc(st);
    float2 *d_bnd = nullptr;
    unsigned long long *counters = nullptr;  // [0] non-monotone rows, [1] rows emitted, [2..3] the (key, row) where the walk stops
    if (int rc = sc.get(&d_bnd, (size_t)n_off)) return rc;
    if (int rc = sc.get(&counters, 4)) return rc;
    PPB_CUDA(cudaMemcpyAsync(d_bnd, bnd.data(), sizeof(float2) * n_off, cudaMemcpyHostToDevice, st));
    PPB_CUDA(cudaMemsetAsync(counters, 0, 32, st));
    const float2 *d2 = reinterpret_cast<const float2 *>(d_dists);
    int64_t *hist = nullptr, *tb = nullptr;
    if (int rc = sc.get(&tb, 512)) return rc;

    // ---- fast path: only the rows some offset admits are sorted
    if (n_off < 65535 && !std::getenv("PPB_ITERATE1D_FULL")) {
        uint32_t *key = nullptr, *blk_key = nullptr, *bits = nullptr, *unit_count = nullptr;
        uint16_t *first16 = nullptr;
        int64_t *blk_row = nullptr, *unit_off = nullptr, *m_dev = nullptr;
        const int64_t cblocks = (n_rows + 256 * ppb::kCls1dRows - 1) / (256 * ppb::kCls1dRows);
        const int64_t units = (n_rows + ppb::kUnitRows - 1) / ppb::kUnitRows;
        if (int rc = sc.get(&key, (size_t)n_rows)) return rc;
        if (int rc = sc.get(&first16, (size_t)n_rows)) return rc;
        if (int rc = sc.get(&blk_key, (size_t)cblocks)) return rc;
        if (int rc = sc.get(&blk_row, (size_t)cblocks)) return rc;
        if (int rc = sc.get(&bits, (size_t)units * 32)) return rc;
        if (int rc = sc.get(&unit_count, (size_t)units)) return rc;
        if (int rc = sc.get(&unit_off, (size_t)units)) return rc;
        if (int rc = sc.get(&m_dev, 1)) return rc;
        ppb::iterate1d_classify_kernel<<<(unsigned)cblocks, 256, 0, st>>>(d2, n_rows, slope, d_bnd, n_off, key, first16, counters,
                                                                         blk_key, blk_row,
                                                                         make_step_search(d_bnd, bnd.data(), n_off, slope));
        ppb::iterate1d_cut_kernel<<<1, 1024, 0, st>>>(blk_key, blk_row, cblocks, counters + 2);
        const unsigned sgrid = (unsigned)std::min<int64_t>((units + 7) / 8, 148 * 16);
        ppb::select_mark_kernel<ppb::PredAdmitted><<<sgrid, 256, 0, st>>>(ppb::PredAdmitted{first16, n_off}, n_rows, bits, unit_count);
        ppb::unit_scan_kernel<<<1, 1024, 0, st>>>(unit_count, units, unit_off, m_dev);
        g_launches += 4;
        PPB_CUDA(cudaGetLastError());
        unsigned long long h_irregular = 0;
        int64_t m = 0;
        PPB_CUDA(cudaMemcpyAsync(&h_irregular, counters, 8, cudaMemcpyDeviceToHost, st));
        PPB_CUDA(cudaMemcpyAsync(&m, m_dev, 8, cudaMemcpyDeviceToHost, st));
        PPB_CUDA(cudaStreamSynchronize(st));
        if (h_irregular == 0) {
            if (m == 0) return PPB_OK;  // (d_count is already 0)
            uint32_t *ck = nullptr, *ck2 = nullptr;
            int64_t *cv = nullptr, *cv2 = nullptr, *order = nullptr;
            int32_t *first = nullptr, *block_max = nullptr;
            const int64_t blocks = (m + ppb::kScanBlock - 1) / ppb::kScanBlock;
            const int64_t chunks = (m + ppb::kSelBlockRows - 1) / ppb::kSelBlockRows;
            if (int rc = sc.get(&ck, (size_t)m)) return rc;
            if (int rc = sc.get(&ck2, (size_t)m)) return rc;
            if (int rc = sc.get(&cv, (size_t)m)) return rc;
            if (int rc = sc.get(&cv2, (size_t)m)) return rc;
            if (int rc = sc.get(&order, (size_t)m)) return rc;
            if (int rc = sc.get(&first, (size_t)m)) return rc;
            if (int rc = sc.get(&block_max, (size_t)blocks)) return rc;
            if (int rc = sc.get(&hist, (size_t)chunks * 256)) return rc;
            ppb::select_emit_kernel<ppb::OutKeyed><<<sgrid, 256, 0, st>>>(bits, unit_count, unit_off, n_rows,
                                                                         ppb::OutKeyed{key, first16, ck, cv}, m);
            g_launches++;
            uint32_t *ka = ck, *kb = ck2;
            int64_t *va = cv, *vb = cv2;
            for (int ps = 0; ps < 4; ps++) {  // stable: equal distances keep row order, like the reference's stable sort
                if (int rc = radix_pass(ppb::DigitOfU32{ka, ps * 8}, ppb::MovePair{ka, va, kb, vb}, m, hist, tb, tb + 256, st)) return rc;
                std::swap(ka, kb);
                std::swap(va, vb);
            }
            ppb::iterate1d_unpack_kernel<<<(unsigned)blocks, 1024, 0, st>>>(va, m, order, first, block_max);
            ppb::max_scan_kernel<<<1, 1024, 0, st>>>(block_max, blocks);
            ppb::iterate1d_emit_kernel<<<(unsigned)blocks, ppb::kScanBlock, 0, st>>>(order, first, block_max, m, n_off, capacity, d_i,
                                                                                     d_j, d_off, counters + 1, n_rows, ka, counters + 2);
            g_launches += 3;
            PPB_CUDA(cudaGetLastError());
            PPB_CUDA(cudaMemcpyAsync(d_count, counters + 1, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
            return PPB_OK;
        }
        PPB_CUDA(cudaMemsetAsync(counters, 0, 32, st));  // some row's test flips back with a later offset: the exact walk below
    }

    // ---- general path: every row is ranked (stable radix sort), then the walk as a running maximum — or, if a row's
    // test is not monotone in the offset (float rounding), the reference's sequential walk verbatim
    uint32_t *k_in = nullptr, *k_out = nullptr;
    int64_t *r_in = nullptr, *order = nullptr;
    int32_t *first = nullptr, *block_max = nullptr;
    const int64_t blocks = (n_rows + ppb::kScanBlock - 1) / ppb::kScanBlock;
    if (int rc = sc.get(&k_in, (size_t)n_rows)) return rc;
    if (int rc = sc.get(&k_out, (size_t)n_rows)) return rc;
    if (int rc = sc.get(&r_in, (size_t)n_rows)) return rc;
    if (int rc = sc.get(&order, (size_t)n_rows)) return rc;
    if (int rc = sc.get(&first, (size_t)n_rows)) return rc;
    if (int rc = sc.get(&block_max, (size_t)blocks)) return rc;
    ppb::iterate1d_keys_kernel<<<grid_for(n_rows, 256, 148 * 16), 256, 0, st>>>(d2, n_rows, slope, bnd[0].x, bnd[0].y, k_in, r_in);
    {
        const int64_t chunks = (n_rows + ppb::kSelBlockRows - 1) / ppb::kSelBlockRows;
        if (int rc = sc.get(&hist, (size_t)chunks * 256)) return rc;
        uint32_t *ka = k_in, *kb = k_out;
        int64_t *va = r_in, *vb = order;
        for (int ps = 0; ps < 4; ps++) {
            if (int rc = radix_pass(ppb::DigitOfU32{ka, ps * 8}, ppb::MovePair{ka, va, kb, vb}, n_rows, hist, tb, tb + 256, st))
                return rc;
            std::swap(ka, kb);
            std::swap(va, vb);
        }
        order = va;  // four passes: the sorted order is back in the first pair of buffers
    }
    ppb::iterate1d_first_kernel<<<(unsigned)blocks, ppb::kScanBlock, 0, st>>>(d2, order, n_rows, slope, d_bnd, n_off, first,
                                                                              block_max, counters);
    ppb::max_scan_kernel<<<1, 1024, 0, st>>>(block_max, blocks);
    ppb::iterate1d_emit_kernel<<<(unsigned)blocks, ppb::kScanBlock, 0, st>>>(order, first, block_max, n_rows, n_off, capacity,
                                                                             d_i, d_j, d_off, counters + 1, n_rows, nullptr, nullptr);
    g_launches += 5;
    PPB_CUDA(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    PPB_CUDA(cudaMemcpyAsync(h, counters, 16, cudaMemcpyDeviceToHost, st));
    PPB_CUDA(cudaStreamSynchronize(st));
    if (h[0] != 0) {  // a row's test flips back with a later offset (float rounding): redo as the reference's exact walk
        ppb::iterate1d_sequential_kernel<<<1, 32, 0, st>>>(d2, order, n_rows, slope, d_bnd, n_off, capacity, d_i, d_j, d_off,
                                                           counters + 1);
        g_launches++;
        PPB_CUDA(cudaGetLastError());
    }
    PPB_CUDA(cudaMemcpyAsync(d_count, counters + 1, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return PPB_OK;
}

int ppb_knn_dev(const float *d_mat, int64_t rows, int64_t cols, int32_t knn, int64_t *d_i, int64_t *d_j, float *d_d,
                void *stream) {
    if (rows < 0 || cols < 0 || knn < 1 || knn > ppb::kKnnMax || cols > 0xffffffffLL)
        return fail(PPB_ERR_ARG, "ppb_knn_dev: kNN must be in [1, 2048]");
    if (rows == 0) return PPB_OK;
    if (!d_mat || !d_i || !d_j || !d_d) return fail(PPB_ERR_ARG, "ppb_knn_dev: bad argument");
    int dev = 0, sms = 0;
    PPB_CUDA(cudaGetDevice(&dev));
    if (int rc = num_sms(dev, &sms)) return rc;
    if (knn <= 32 && !std::getenv("PPB_KNN_GENERIC")) {
        ppb::knn_small_kernel<ppb::DenseRowCands><<<(unsigned)std::min<int64_t>(rows, (int64_t)sms * 8), 256, 0,
                                                    (cudaStream_t)stream>>>(ppb::DenseRowCands{d_mat, cols, rows}, rows, knn, d_i,
                                                                            d_j, d_d);
    } else {
        ppb::knn_kernel<ppb::DenseRowCands><<<(unsigned)std::min<int64_t>(rows, (int64_t)sms * 8), ppb::kKnnThreads, 0,
                                              (cudaStream_t)stream>>>(ppb::DenseRowCands{d_mat, cols, rows}, rows, knn, d_i, d_j, d_d);
    }
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_extend_dev(const int64_t *d_sp_i, const int64_t *d_sp_j, const float *d_sp_d, int64_t nnz, const float *d_qq,
                   const float *d_qr, int64_t nr, int64_t nq, int32_t knn, int64_t *d_i, int64_t *d_j, float *d_d,
                   int64_t *d_count, void *stream) {
    if (nnz < 0 || nr < 0 || nq < 0 || knn < 1 || knn > ppb::kKnnMax || !d_count || nr + nq > 0xffffffffLL ||
        (nnz > 0 && (!d_sp_i || !d_sp_j || !d_sp_d)) || (nq > 0 && !d_qq) || (nr > 0 && nq > 0 && !d_qr))
        return fail(PPB_ERR_ARG, "ppb_extend_dev: bad argument (kNN must be in [1, 2048])");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_total = nr + nq;
    if (n_total == 0) {
        PPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
        return PPB_OK;
    }
    if (!d_i || !d_j || !d_d) return fail(PPB_ERR_ARG, "ppb_extend_dev: bad argument");
    Scratch sc(st);
    int64_t *row_start = nullptr, *cnt = nullptr;
    if (int rc = sc.get(&row_start, (size_t)nr + 1)) return rc;
    if (int rc = sc.get(&cnt, (size_t)n_total)) return rc;
    ppb::row_starts_kernel<<<grid_for(nr + 1, 256, 148 * 8), 256, 0, st>>>(d_sp_i, nnz, nr, row_start);
    ppb::ExtendCands c{row_start, d_sp_j, d_sp_d, d_qq, d_qr, nr, nq, cnt};
    ppb::extend_count_kernel<<<grid_for(n_total, 256, 148 * 8), 256, 0, st>>>(c, n_total, knn, cnt);
    ppb::scan_kernel<<<1, 1024, 0, st>>>(cnt, n_total, d_count);
    int dev = 0, sms = 0;
    PPB_CUDA(cudaGetDevice(&dev));
    if (int rc = num_sms(dev, &sms)) return rc;
    if (knn <= 32 && !std::getenv("PPB_KNN_GENERIC"))
        ppb::knn_small_kernel<ppb::ExtendCands><<<(unsigned)std::min<int64_t>(n_total, (int64_t)sms * 8), 256, 0, st>>>(
            c, n_total, knn, d_i, d_j, d_d);
    else
        ppb::knn_kernel<ppb::ExtendCands><<<(unsigned)std::min<int64_t>(n_total, (int64_t)sms * 8), ppb::kKnnThreads, 0, st>>>(
            c, n_total, knn, d_i, d_j, d_d);
    g_launches += 4;
    PPB_CUDA(cudaGetLastError());
    return PPB_OK;
}

int ppb_lower_rank_dev(const int64_t *d_sp_i, const int64_t *d_sp_j, const float *d_sp_d, int64_t nnz, int64_t n_samples,
                       int64_t knn, int32_t reciprocal_only, int32_t count_unique_distances, float epsilon, int64_t *d_i,
                       int64_t *d_j, float *d_d, int64_t *d_count, void *stream) {
    if (nnz < 0 || n_samples < 0 || knn < 0 || !d_count || (nnz > 0 && (!d_sp_i || !d_sp_j || !d_sp_d || !d_i || !d_j || !d_d)))
        return fail(PPB_ERR_ARG, "ppb_lower_rank_dev: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (nnz == 0 || n_samples == 0) return PPB_OK;
    Scratch sc(st);
    int64_t *row_start = nullptr, *st_j = nullptr, *kept = nullptr, *cnt = nullptr;
    float *st_d = nullptr;
    uint8_t *flag = nullptr;
    int32_t *too_long = nullptr;
    if (int rc = sc.get(&row_start, (size_t)n_samples + 1)) return rc;
    if (int rc = sc.get(&st_j, (size_t)nnz)) return rc;
    if (int rc = sc.get(&st_d, (size_t)nnz)) return rc;
    if (int rc = sc.get(&kept, (size_t)n_samples)) return rc;
    if (int rc = sc.get(&cnt, (size_t)n_samples)) return rc;
    if (int rc = sc.get(&too_long, 1)) return rc;
    PPB_CUDA(cudaMemsetAsync(too_long, 0, 4, st));
    ppb::row_starts_kernel<<<grid_for(n_samples + 1, 256, 148 * 8), 256, 0, st>>>(d_sp_i, nnz, n_samples, row_start);
    const unsigned blocks = (unsigned)((n_samples + ppb::kLowerWarps - 1) / ppb::kLowerWarps);
    ppb::lower_rank_keep_kernel<<<blocks, ppb::kLowerWarps * 32, 0, st>>>(row_start, d_sp_j, d_sp_d, n_samples, knn,
                                                                         count_unique_distances, epsilon, st_j, st_d, kept,
                                                                         too_long);
    g_launches += 2;
    if (reciprocal_only) {
        if (int rc = sc.get(&flag, (size_t)nnz)) return rc;
        ppb::lower_rank_reciprocal_kernel<<<grid_for(n_samples, 128, 148 * 16), 128, 0, st>>>(row_start, st_j, kept, n_samples,
                                                                                             flag, cnt);
        g_launches++;
    } else {
        PPB_CUDA(cudaMemcpyAsync(cnt, kept, sizeof(int64_t) * n_samples, cudaMemcpyDeviceToDevice, st));
    }
    ppb::scan_kernel<<<1, 1024, 0, st>>>(cnt, n_samples, d_count);
    ppb::lower_rank_write_kernel<<<grid_for(n_samples, 128, 148 * 16), 128, 0, st>>>(row_start, st_j, st_d, kept, flag, cnt,
                                                                                    n_samples, d_i, d_j, d_d);
    g_launches += 2;
    PPB_CUDA(cudaGetLastError());
    int32_t h_long = 0;
    PPB_CUDA(cudaMemcpyAsync(&h_long, too_long, 4, cudaMemcpyDeviceToHost, st));
    PPB_CUDA(cudaStreamSynchronize(st));
    if (h_long) return fail(PPB_ERR_ARG, "ppb_lower_rank_dev: a sample has more than 1024 sparse neighbours");
    return PPB_OK;
}

int ppb_microbench_dev(int32_t mode, int64_t iters, uint32_t *d_sink, int64_t *lane_ops, void *stream) {
    const bool half_occupancy = mode >= 10;  // 10+m: mode m with ONE 256-thread CTA per SM (2 warps per scheduler)
    if (half_occupancy) mode -= 10;
    if (mode < 0 || mode > 6 || iters < 1 || !d_sink) return fail(PPB_ERR_ARG, "ppb_microbench_dev: bad argument");
    int dev = 0, sms = 0;
    PPB_CUDA(cudaGetDevice(&dev));
    if (int rc = num_sms(dev, &sms)) return rc;
    const unsigned grid = half_occupancy ? sms : sms * 2, threads = 256;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t per_thread_iter = 0;
    switch (mode) {
        case 0: ppb::microbench_kernel<0><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
        case 1: ppb::microbench_kernel<1><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
        case 2: ppb::microbench_kernel<2><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
        case 3: ppb::microbench_kernel<3><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 8; break;
        case 4: ppb::microbench_kernel<4><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
        case 5: ppb::microbench_kernel<5><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
        default: ppb::microbench_kernel<6><<<grid, threads, 0, st>>>(iters, d_sink, 12345u); per_thread_iter = 112; break;
    }
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    // lane-ops of the op being measured (mode 2 counts its LOP3s)
    if (lane_ops) *lane_ops = (int64_t)grid * threads * iters * per_thread_iter;
    return PPB_OK;
}

int ppb_microbench_mix_dev(int32_t rows_per_warp, int32_t warps_per_scheduler, int32_t with_lds, int64_t iters,
                           uint32_t *d_sink, int64_t *lop3_lane_ops, void *stream) {
    const int max_w = rows_per_warp == 8 ? 2 : rows_per_warp == 5 ? 3 : rows_per_warp == 4 ? 4 : 0;  // what the register file holds
    if (warps_per_scheduler < 1 || warps_per_scheduler > max_w || iters < 1 || !d_sink)
        return fail(PPB_ERR_ARG, "ppb_microbench_mix_dev: (rows_per_warp, warps_per_scheduler) must be (8, <=2), (5, <=3) or (4, <=4)");
    int dev = 0, sms = 0;
    PPB_CUDA(cudaGetDevice(&dev));
    if (int rc = num_sms(dev, &sms)) return rc;
    const unsigned threads = 128u * warps_per_scheduler;
    const size_t smem = 4 * ppb::kSliceBytes + 16 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows_per_warp == 8)
        ppb::mixbench_kernel<8, 256><<<sms, threads, smem, st>>>(iters, d_sink, 12345u, with_lds);
    else if (rows_per_warp == 5)
        ppb::mixbench_kernel<5, 384><<<sms, threads, smem, st>>>(iters, d_sink, 12345u, with_lds);
    else
        ppb::mixbench_kernel<4, 512><<<sms, threads, smem, st>>>(iters, d_sink, 12345u, with_lds);
    g_launches++;
    PPB_CUDA(cudaGetLastError());
    if (lop3_lane_ops) *lop3_lane_ops = (int64_t)sms * threads * iters * 4 * rows_per_warp * 14;
    return PPB_OK;
}

// ------------------------------------------------------------------------------------------------
// Host-buffer path: see ppb_host.inl (included below)
// ------------------------------------------------------------------------------------------------
extern "C++" {
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(PPB_ERR_NOMEM, "cudaMalloc failed for " + std::to_string(bytes) + " bytes");
        }
        return PPB_OK;
    }
};
}  // namespace
}  // extern "C++"

int64_t ppb_plan_tiles(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end, int32_t tile_cols,
                       int32_t band_tiles, int32_t *tiles, int64_t max_tiles) {
    const int64_t total_rows = ppb_num_rows(n_ref, n_qry, self);
    if (n_ref < 1 || (!self && n_qry < 1) || row_begin < 0 || row_end > total_rows || row_begin >= row_end || tile_cols < 1 ||
        band_tiles < 1)
        return -1;
    TileShape shape;
    shape.tj = tile_cols, shape.band = band_tiles;
    const TileKey key = make_tile_key(0, n_ref, n_qry, self, row_begin, row_end, shape);   // as ppb_query_dev and the host call do
    std::vector<int2> v;
    plan_tiles(key, &v);
    for (size_t t = 0; t < v.size() && (int64_t)t < max_tiles && tiles; t++) {
        tiles[2 * t] = v[t].x;
        tiles[2 * t + 1] = v[t].y;
    }
    return (int64_t)v.size();
}

int64_t ppb_plan_host_chunks(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end,
                             int64_t cap_rows, int64_t *bounds, int64_t max_chunks) {
    const int64_t total_rows = ppb_num_rows(n_ref, n_qry, self);
    if (n_ref < 0 || (!self && n_qry < 0) || row_begin < 0 || row_end > total_rows || row_begin > row_end || cap_rows < 1)
        return -1;
    std::vector<std::pair<int64_t, int64_t>> chunks;
    plan_chunks(n_ref, n_qry, self, row_begin, row_end, cap_rows, &chunks);
    for (size_t c = 0; c < chunks.size() && (int64_t)c < max_chunks && bounds; c++) {
        bounds[2 * c] = chunks[c].first;
        bounds[2 * c + 1] = chunks[c].second;
    }
    return (int64_t)chunks.size();
}

}  // extern "C"

#include "ppb_host.inl"
