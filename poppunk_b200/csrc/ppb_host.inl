// Host-buffer entry points of libppb.so (included at the end of ppb_api.cu): ppb_query_host / ppb_query_host_multi.
//
// What a PopPUNK process calls (poppunk_b200/sketchlib.py::pp_queryDatabase, replacing pp_sketchlib.queryDatabase at
// PopPUNK/sketchlib.py:528-537, 584-593): NumPy sketch arrays in, ONE NumPy result array out, every visible GPU used.
//
//   caller thread: argument checks, static row shards (one contiguous row range per device, cut on row-tile
//                  boundaries), peer-access setup, one worker thread per device
//   worker g     : phase A  allocate the device workspace (packed sketches, result ring) — rendezvous
//                  phase B  upload 1/G of the reference genomes, pack them INTO EVERY DEVICE'S packed array through
//                           peer-mapped pointers (NVLink), upload + pack this device's own queries — rendezvous
//                  phase C  row-chunked kernel launches on one stream, D2H of finished chunks on a second stream,
//                           straight into the caller's array at the shard's offset
//
// There is no exchange between devices for the host result: each device's rows land in the one caller buffer.
// Results bound for pageable memory (a plain np.empty) go through a pinned staging ring per device and are copied
// out by that device's consumer thread; the destination is first advised MADV_HUGEPAGE so that its first-touch
// faults are 2 MiB each.
#include <sys/mman.h>
#include <unistd.h>

namespace {

struct DeviceGuard {  // host entry points leave the caller's current device as they found it
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Grow-only workspace of one device, reused by successive host calls (cudaMalloc / cudaHostAlloc of multi-GB buffers
// per call would otherwise show up in the end-to-end time).  One host call at a time per device (mu).
struct DevWorkspace {
    std::mutex mu;
    std::map<int, std::pair<void *, size_t>> slots, pinned;
    static int grow(std::pair<void *, size_t> &e, size_t bytes, bool host, void **out) {
        if (e.second < bytes || !e.first) {
            if (e.first) host ? cudaFreeHost(e.first) : cudaFree(e.first);
            e.first = nullptr;
            e.second = 0;
            const size_t want = std::max<size_t>(bytes, 256);
            const cudaError_t rc = host ? cudaHostAlloc(&e.first, want, cudaHostAllocPortable) : cudaMalloc(&e.first, want);
            if (rc != cudaSuccess) {
                cudaGetLastError();
                e.first = nullptr;
                return fail(PPB_ERR_NOMEM, std::string(host ? "cudaHostAlloc" : "cudaMalloc") + " failed for " +
                                               std::to_string(bytes) + " bytes");
            }
            e.second = want;
        }
        *out = e.first;
        return PPB_OK;
    }
    int get(int slot, size_t bytes, void **out) { return grow(slots[slot], bytes, false, out); }
    int get_pinned(int slot, size_t bytes, void **out) { return grow(pinned[slot], bytes, true, out); }
    void release() {
        for (auto &kv : slots)
            if (kv.second.first) cudaFree(kv.second.first);
        slots.clear();
        for (auto &kv : pinned)
            if (kv.second.first) cudaFreeHost(kv.second.first);
        pinned.clear();
    }
};
std::mutex g_ws_mu;
std::map<int, std::unique_ptr<DevWorkspace>> g_ws;
DevWorkspace &workspace(int dev) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto &p = g_ws[dev];
    if (!p) p.reset(new DevWorkspace);
    return *p;
}

constexpr int kHostRing = 8;  // result buffers per device (chunks in flight between kernel and D2H)
constexpr int kUpRing = 4;    // pinned staging buffers of a pageable upload
constexpr size_t kUpPiece = (size_t)32 << 20;
enum { WS_REF_RAW, WS_QRY_RAW, WS_REF, WS_QRY, WS_TAB, WS_RC, WS_QC, WS_DEG, WS_YTAB, WS_TILES, WS_OUT0, WS_LAB0 = WS_OUT0 + kHostRing };
enum { PIN_OUT0 = 0, PIN_UP0 = 16, PIN_TILES = 32 };

// true when the CUDA driver can DMA straight into / out of p (pinned / registered / managed host memory)
bool is_dma_able(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice;
}
// memcpy split over a few threads: one core cannot move 50 GB/s between a staging buffer and fresh pages
void parallel_memcpy(void *dst, const void *src, size_t bytes, int threads) {
    if (threads <= 1 || bytes < ((size_t)4 << 20)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t piece = ((bytes + threads - 1) / threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) {
        const size_t off = (size_t)t * piece;
        if (off >= bytes) break;
        pool.emplace_back([=] { std::memcpy((char *)dst + off, (const char *)src + off, std::min(piece, bytes - off)); });
    }
    std::memcpy(dst, src, std::min(piece, bytes));
    for (auto &th : pool) th.join();
}
// ask for transparent huge pages under a (pageable) destination: first-touch faults become 2 MiB instead of 4 KiB
void advise_hugepages(void *p, size_t bytes) {
    const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, e = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
    if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
}


// ---- host result pool -------------------------------------------------------------------------------------------
// The drop-in allocates the (n_pairs, 2) result itself (PopPUNK owns the returned NumPy array).  Blocks come from
// here: anonymous mappings advised MADV_HUGEPAGE; a block handed back (the array was garbage-collected) is kept — its
// pages are touched, so a staged copy into it no longer faults — and page-locked (cudaHostRegister) on its second
// reuse, after which the result is DMA-ed straight into the array the caller receives (no staging copy).  Measured on the 16-vCPU
// B200 host (tools/host_floor.cu): cudaHostAlloc 2.4 GB/s, cudaHostRegister of touched huge pages 23 GB/s, staged
// memcpy into fresh huge pages 45 GB/s, pinned D2H 54 GB/s — so a first call is fastest through the staging ring
// and a pinned fresh allocation never pays for a single call.
struct HostBlock {
    size_t cap = 0;
    bool in_use = false, pinned = false, touched = false;
    int reuses = 0;
};
std::mutex g_pool_mu;
std::map<void *, HostBlock> g_pool;
size_t pool_limit_bytes() {
    if (const char *e = std::getenv("PPB_HOST_POOL_MAX_GB")) return (size_t)(atof(e) * 1e9);
    const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
    return pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz / 2 : (size_t)64 << 30;
}
void pool_drop(std::map<void *, HostBlock>::iterator it) {
    if (it->second.pinned) {
        cudaHostUnregister(it->first);
        cudaGetLastError();
    }
    munmap(it->first, it->second.cap);
    g_pool.erase(it);
}

struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() {
        if (s) cudaStreamDestroy(s);
    }
};
struct Event {
    cudaEvent_t e = nullptr;
    ~Event() {
        if (e) cudaEventDestroy(e);
    }
};

// Barrier of the per-device workers that also carries failure: arrive(false) makes every participant's arrive()
// return false at this and all later meeting points, so the workers of a failed call leave together.
struct Rendezvous {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0;
    uint64_t gen = 0;
    bool failed = false;
    explicit Rendezvous(int n_) : n(n_) {}
    bool arrive(bool ok) {
        std::unique_lock<std::mutex> lk(mu);
        if (!ok) failed = true;
        if (++waiting == n) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            const uint64_t g = gen;
            cv.wait(lk, [&] { return gen != g; });
        }
        return !failed;
    }
};

struct HostJob {
    const uint64_t *ref, *qry;
    int64_t n_ref, n_qry;
    const int32_t *kmers;
    int32_t K, ss64;
    const float *rand_table;
    int32_t C;
    const uint16_t *ref_cluster, *qry_cluster;
    int64_t row_begin, row_end;
    int32_t out_mode;
    void *out;
    const ppb_boundary *boundary;
    int8_t *labels;
    int self, G;
    bool p2p, staged, trace;
    std::chrono::steady_clock::time_point t0;   // start of the call (PPB_HOST_TRACE: wall-clock stamps of the phases)
    int copy_threads;
    std::vector<int> devs;
    std::vector<int64_t> row_cut;  // G+1 row boundaries: device g computes [row_cut[g], row_cut[g+1])
    std::vector<int64_t> gen_cut;  // G+1 genome boundaries of the packed reference array: device g uploads + packs its part
    std::vector<uint32_t *> d_ref_packed;
    std::vector<cudaEvent_t> packed_ev;
    std::vector<int> rc;
    std::vector<std::string> err;
    std::vector<unsigned long long> deg;
    Rendezvous *rv;
};

// host -> device copy of `bytes`; pageable sources are staged through a small pinned ring by a few memcpy threads
// (the driver's own pageable path is a single-threaded staging loop)
int upload(DevWorkspace &ws, void *d_dst, const void *src, size_t bytes, cudaStream_t st, int threads) {
    if (bytes == 0) return PPB_OK;
    if (is_dma_able(src) || bytes < ((size_t)8 << 20)) {
        PPB_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st));
        return PPB_OK;
    }
    void *ring[kUpRing];
    Event ev[kUpRing];
    for (int b = 0; b < kUpRing; b++) {
        if (int rc = ws.get_pinned(PIN_UP0 + b, kUpPiece, &ring[b])) return rc;
        PPB_CUDA(cudaEventCreateWithFlags(&ev[b].e, cudaEventDisableTiming));
    }
    size_t c = 0;
    for (size_t off = 0; off < bytes; off += kUpPiece, c++) {
        const int b = (int)(c % kUpRing);
        const size_t len = std::min(kUpPiece, bytes - off);
        if (c >= (size_t)kUpRing) PPB_CUDA(cudaEventSynchronize(ev[b].e));
        parallel_memcpy(ring[b], (const char *)src + off, len, threads);
        PPB_CUDA(cudaMemcpyAsync((char *)d_dst + off, ring[b], len, cudaMemcpyHostToDevice, st));
        PPB_CUDA(cudaEventRecord(ev[b].e, st));
    }
    PPB_CUDA(cudaStreamSynchronize(st));  // the ring belongs to the workspace: nobody may reuse it under a copy
    return PPB_OK;
}

// One device's share of a host-buffer call.  Runs on its own thread; meets the other workers at job.rv three times.
int host_worker(HostJob &job, int g) {
    const int dev = job.devs[g];
    Rendezvous &rv = *job.rv;
    const int64_t r_lo = job.row_cut[g], r_hi = job.row_cut[g + 1];
    const int K = job.K, ss64 = job.ss64;
    const int64_t W = (int64_t)ss64 * PPB_BBITS;
    const size_t genome_bytes = (size_t)K * W * 8;

    DevWorkspace &ws = workspace(dev);
    std::unique_lock<std::mutex> ws_lock(ws.mu, std::defer_lock);
    Stream s_compute, s_copy;
    void *d_ref_raw = nullptr, *d_qry_raw = nullptr, *d_ref = nullptr, *d_qry = nullptr, *d_tab = nullptr, *d_rc = nullptr,
         *d_qc = nullptr, *d_deg = nullptr;
    // reference genomes this device uploads and packs: its part of the array (peer-to-peer scatter) or all of it
    const int64_t n_pad = round_up(std::max<int64_t>(job.n_ref, 1), ppb::kPad);
    const int64_t part_lo = job.p2p ? job.gen_cut[g] : 0, part_hi = job.p2p ? job.gen_cut[g + 1] : n_pad;
    const int64_t part_real = std::max<int64_t>(0, std::min(part_hi, job.n_ref) - part_lo);
    // queries this device needs (non-self): the ones its rows belong to
    int64_t q_lo = 0, q_hi = 0;
    if (!job.self && r_hi > r_lo) {
        q_lo = r_lo / job.n_ref;
        q_hi = (r_hi - 1) / job.n_ref + 1;
    }
    const int64_t n_q = q_hi - q_lo;

    // ---- phase A: device, streams, buffers
    auto phase_a = [&]() -> int {
        PPB_CUDA(cudaSetDevice(dev));
        ws_lock.lock();
        PPB_CUDA(cudaStreamCreateWithFlags(&s_compute.s, cudaStreamNonBlocking));
        PPB_CUDA(cudaStreamCreateWithFlags(&s_copy.s, cudaStreamNonBlocking));
        PPB_CUDA(cudaEventCreateWithFlags(&job.packed_ev[g], cudaEventDisableTiming));
        if (int rc = ws.get(WS_REF, ppb_packed_bytes(job.n_ref, K, ss64), &d_ref)) return rc;
        job.d_ref_packed[g] = (uint32_t *)d_ref;
        if (int rc = ws.get(WS_REF_RAW, (size_t)part_real * genome_bytes, &d_ref_raw)) return rc;
        if (n_q > 0) {
            if (int rc = ws.get(WS_QRY_RAW, (size_t)n_q * genome_bytes, &d_qry_raw)) return rc;
            if (int rc = ws.get(WS_QRY, ppb_packed_bytes(n_q, K, ss64), &d_qry)) return rc;
        }
        if (int rc = ws.get(WS_DEG, 8, &d_deg)) return rc;
        return PPB_OK;
    };
    auto stamp = [&](const char *what) {
        if (job.trace)
            std::fprintf(stderr, "[ppb_query_host dev %d] %s at %.1f ms of the call\n", dev, what,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - job.t0).count());
    };
    int rc = phase_a();
    if (rc) job.err[g] = g_err;
    stamp("workspace ready");
    if (!rv.arrive(rc == PPB_OK)) return rc;

    // ---- phase B: upload + pack
    auto phase_b = [&]() -> int {
        cudaStream_t st = s_compute.s;
        if (int rc = upload(ws, d_ref_raw, job.ref + (size_t)part_lo * K * W, (size_t)part_real * genome_bytes, st, job.copy_threads))
            return rc;
        std::vector<uint32_t *> dsts;
        if (job.p2p)
            dsts = job.d_ref_packed;
        else
            dsts.push_back((uint32_t *)d_ref);
        if (int rc = ppb_pack_part_dev((const uint64_t *)d_ref_raw, nullptr, part_lo, part_hi, job.n_ref, K, ss64, dsts.data(),
                                       (int32_t)dsts.size(), st))
            return rc;
        PPB_CUDA(cudaEventRecord(job.packed_ev[g], st));
        if (n_q > 0) {
            if (int rc = upload(ws, d_qry_raw, job.qry + (size_t)q_lo * K * W, (size_t)n_q * genome_bytes, st, job.copy_threads))
                return rc;
            if (int rc = ppb_pack_dev((const uint64_t *)d_qry_raw, n_q, nullptr, n_q, K, ss64, (uint32_t *)d_qry, st)) return rc;
        }
        if (job.rand_table) {
            const size_t tb = (size_t)job.C * job.C * K * sizeof(float);
            if (int rc = ws.get(WS_TAB, tb, &d_tab)) return rc;
            if (int rc = ws.get(WS_RC, (size_t)job.n_ref * 2, &d_rc)) return rc;
            PPB_CUDA(cudaMemcpyAsync(d_tab, job.rand_table, tb, cudaMemcpyHostToDevice, st));
            PPB_CUDA(cudaMemcpyAsync(d_rc, job.ref_cluster, (size_t)job.n_ref * 2, cudaMemcpyHostToDevice, st));
            if (n_q > 0) {
                if (int rc = ws.get(WS_QC, (size_t)n_q * 2, &d_qc)) return rc;
                PPB_CUDA(cudaMemcpyAsync(d_qc, job.qry_cluster + q_lo, (size_t)n_q * 2, cudaMemcpyHostToDevice, st));
            }
        }
        PPB_CUDA(cudaMemsetAsync(d_deg, 0, 8, st));
        return PPB_OK;
    };
    rc = phase_b();
    if (rc) job.err[g] = g_err;
    stamp("inputs uploaded, pack enqueued");
    if (!rv.arrive(rc == PPB_OK)) return rc;

    // ---- phase C: row chunks.  kernel(c) on s_compute overlaps D2H(c-1, c-2, ...) on s_copy.  Chunks end on
    // row-TILE boundaries (kTI genomes of the row side), so no tile is computed by two launches, and they rotate
    // through a ring of up to kHostRing device buffers: the result leaves over PCIe at about the rate the kernel
    // produces it (8 B/pair), so the ring — not a double buffer — is what absorbs the jitter between the two.
    auto phase_c = [&]() -> int {
        if (r_hi <= r_lo) return PPB_OK;
        if (job.p2p)
            for (int h = 0; h < job.G; h++)
                if (h != g) PPB_CUDA(cudaStreamWaitEvent(s_compute.s, job.packed_ev[h], 0));  // every part has landed here
        const int rb = out_row_bytes(job.out_mode, K);
        const int64_t per_row = (job.out ? rb : 0) + (job.labels ? 1 : 0);
        size_t free_b = 0, total_b = 0;
        PPB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        int64_t cap = (int64_t)1 << 26;  // 64 Mi rows = 512 MiB of float2 per device buffer
        if (const char *e = std::getenv("PPB_HOST_CHUNK_ROWS")) cap = std::max<int64_t>(1024, atoll(e));
        while (cap > 1024 && (size_t)(2 * cap * per_row) > free_b / 2) cap >>= 1;
        // rows of this device, in the coordinates of the launch (non-self: relative to its own query range)
        const int64_t shift = job.self ? 0 : q_lo * job.n_ref;
        std::vector<std::pair<int64_t, int64_t>> chunks;
        {   // ramp: the first launches are small (cap/8, cap/4, cap/2), so the first device-to-host copy starts a
            // millisecond after the first launch instead of ten — as long as a launch still covers whole row tiles
            int64_t r = r_lo - shift;
            const int64_t end = r_hi - shift, tile_rows = (int64_t)ppb::kTI * job.n_ref;
            if (!std::getenv("PPB_HOST_CHUNK_ROWS"))
                for (int64_t c = cap / 8; c < cap && r < end && tile_rows <= c; c *= 2) {
                    std::vector<std::pair<int64_t, int64_t>> first;
                    plan_chunks(job.n_ref, n_q, job.self, r, std::min(end, r + c), c, &first);
                    chunks.push_back(first.front());
                    r = first.front().second;
                }
            if (r < end) plan_chunks(job.n_ref, n_q, job.self, r, end, cap, &chunks);
        }
        int64_t max_chunk = 0;
        for (auto &c : chunks) max_chunk = std::max(max_chunk, c.second - c.first);
        int n_buf = (int)std::min<size_t>(chunks.size(), job.staged ? 4 : kHostRing);
        while (n_buf > 2 && (size_t)n_buf * max_chunk * per_row > free_b / 2) n_buf--;
        if (const char *e = std::getenv("PPB_HOST_RING")) n_buf = std::max(1, std::min(atoi(e), kHostRing));
        n_buf = std::max(1, std::min<int>(n_buf, (int)chunks.size()));

        char *out_base = job.out ? (char *)job.out - (size_t)(job.row_begin - shift) * rb : nullptr;  // row r of the launch -> out_base + r*rb
        int8_t *lab_base = job.labels ? job.labels - (job.row_begin - shift) : nullptr;
        void *d_out[kHostRing] = {}, *d_lab[kHostRing] = {};
        Event done_compute[kHostRing], done_copy[kHostRing];
        for (int b = 0; b < n_buf; b++) {
            if (job.out)
                if (int rc = ws.get(WS_OUT0 + b, (size_t)max_chunk * rb, &d_out[b])) return rc;
            if (job.labels)
                if (int rc = ws.get(WS_LAB0 + b, (size_t)max_chunk, &d_lab[b])) return rc;
            PPB_CUDA(cudaEventCreateWithFlags(&done_compute[b].e, cudaEventDisableTiming));
            PPB_CUDA(cudaEventCreateWithFlags(&done_copy[b].e, cudaEventDisableTiming));
        }
        std::vector<cudaEvent_t> tr;  // PPB_HOST_TRACE: (kernel begin, kernel end, copy begin, copy end) per chunk
        cudaEvent_t tr_start = nullptr;
        if (job.trace) {
            tr.resize(chunks.size() * 4);
            for (auto &e : tr) PPB_CUDA(cudaEventCreate(&e));
            PPB_CUDA(cudaEventCreate(&tr_start));
            PPB_CUDA(cudaEventRecord(tr_start, s_compute.s));
        }
        // one y-table for all launches of this call (same k-mers, table and sketch size throughout)
        YtabLease lease;
        if (job.out_mode == PPB_OUT_DISTS) {
            const size_t entries = (size_t)(job.rand_table ? (size_t)job.C * job.C : 1) * K * ((size_t)64 * ss64 + 1);
            if (entries * sizeof(double) <= ((size_t)64 << 20)) {
                void *q = nullptr;
                if (int rc = ws.get(WS_YTAB, entries * sizeof(double), &q)) return rc;
                lease.buf = (double *)q;
                lease.capacity = entries;
            }
        }
        struct LeaseScope {  // the lease is visible to the launches of THIS thread only, whatever path leaves it
            explicit LeaseScope(YtabLease *l) { g_ytab_lease = l; }
            ~LeaseScope() { g_ytab_lease = nullptr; }
        } lease_scope(lease.buf ? &lease : nullptr);

        // tile lists of all launches of this device: planned here, uploaded with ONE copy and lent to the launches of this
        // thread.  (A launch that plans its own list allocates, copies from pageable memory and synchronises its stream;
        // with 100 chunks on each of 8 devices that serialised the devices: 4.5 s instead of 1.0 s for 5e10 labels.)
        TileLease tile_lease;
        {
            TileShape shape;
            if (int rc = tile_shape(dev, K, ss64, &shape)) return rc;
            std::vector<int2> all, v;
            std::vector<TileKey> keys;
            std::vector<std::pair<size_t, size_t>> span;   // (offset, count) in `all`
            for (auto &c : chunks) {
                keys.push_back(make_tile_key(dev, job.n_ref, n_q, job.self, c.first, c.second, shape));
                v.clear();
                plan_tiles(keys.back(), &v);
                span.emplace_back(all.size(), v.size());
                all.insert(all.end(), v.begin(), v.end());
            }
            void *h_tiles = nullptr, *d_tiles = nullptr;
            if (!all.empty()) {
                const size_t bytes = all.size() * sizeof(int2);
                if (int rc = ws.get_pinned(PIN_TILES, bytes, &h_tiles)) return rc;
                if (int rc = ws.get(WS_TILES, bytes, &d_tiles)) return rc;
                std::memcpy(h_tiles, all.data(), bytes);
                PPB_CUDA(cudaMemcpyAsync(d_tiles, h_tiles, bytes, cudaMemcpyHostToDevice, s_compute.s));
            }
            for (size_t c = 0; c < chunks.size(); c++) {
                TileList tl;
                tl.d = span[c].second ? (int2 *)d_tiles + span[c].first : nullptr;
                tl.n = (int64_t)span[c].second;
                tile_lease.lists[keys[c]] = tl;
            }
        }
        struct TileLeaseScope {
            explicit TileLeaseScope(TileLease *l) { g_tile_lease = l; }
            ~TileLeaseScope() { g_tile_lease = nullptr; }
        } tile_lease_scope(&tile_lease);

        // ---- staged mode (pageable destination).  A finished chunk leaves the device in TRANSFERS of <= kXferBytes:
        // D2H into a ring of pinned slots, and a pool of copy threads drains the slots into the caller's pages in
        // 4 MiB pieces (claimed from one shared counter, so no thread idles at a chunk boundary).  First touch of the
        // destination (2 MiB faults after MADV_HUGEPAGE) happens inside those memcpys: the host cores, not PCIe, set
        // the pace of a first call (tools/host_floor.cu: ~40 GB/s with all cores of the 16-vCPU B200 host).
        struct Xfer {
            size_t chunk;
            const char *d_src;   // device address
            char *dst;           // final host address
            size_t len;
            size_t piece0;       // index of its first piece
        };
        constexpr size_t kXferBytes = (size_t)16 << 20, kPieceBytes = (size_t)4 << 20;   // (pinned memory costs ~0.5 s per GiB to get)
        constexpr int kPinRing = 8;
        std::vector<Xfer> xfers;
        std::vector<size_t> chunk_x0(chunks.size() + 1, 0);  // transfers of chunk c: [chunk_x0[c], chunk_x0[c+1])
        size_t n_pieces = 0;
        void *pin[kPinRing] = {};
        int n_pin = 0;
        if (job.staged) {
            for (size_t c = 0; c < chunks.size(); c++) {
                const int b = (int)(c % n_buf);
                const int64_t r0 = chunks[c].first, r1 = chunks[c].second;
                chunk_x0[c] = xfers.size();
                auto add = [&](const char *d_base, char *h_base, size_t bytes) {
                    for (size_t off = 0; off < bytes; off += kXferBytes) {
                        const size_t len = std::min(kXferBytes, bytes - off);
                        xfers.push_back(Xfer{c, d_base + off, h_base + off, len, n_pieces});
                        n_pieces += (len + kPieceBytes - 1) / kPieceBytes;
                    }
                };
                if (job.out) add((const char *)d_out[b], out_base + (size_t)r0 * rb, (size_t)(r1 - r0) * rb);
                if (job.labels) add((const char *)d_lab[b], (char *)(lab_base + r0), (size_t)(r1 - r0));
            }
            chunk_x0[chunks.size()] = xfers.size();
            n_pin = (int)std::min<size_t>(kPinRing, xfers.size());
            for (int q = 0; q < n_pin; q++)
                if (int rc = ws.get_pinned(PIN_OUT0 + q, kXferBytes, &pin[q])) return rc;
        }
        std::vector<cudaEvent_t> landed(xfers.size(), nullptr);  // transfer t has arrived in its pinned slot
        for (auto &e : landed) PPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        struct LandedScope {
            std::vector<cudaEvent_t> &v;
            ~LandedScope() {
                for (auto e : v)
                    if (e) cudaEventDestroy(e);
            }
        } landed_scope{landed};
        std::mutex mu;
        std::condition_variable cv;
        size_t n_enqueued = 0, n_drained = 0;          // transfers whose D2H is enqueued / whose slot is free again (in order)
        std::vector<int> pieces_left(xfers.size(), 0);   // per transfer, under mu
        std::vector<char> drained(xfers.size(), 0);
        for (size_t t = 0; t < xfers.size(); t++) pieces_left[t] = (int)((xfers[t].len + kPieceBytes - 1) / kPieceBytes);
        std::atomic<size_t> next_piece{0};
        bool abort_copy = false;
        std::atomic<bool> copy_failed{false};
        auto copy_worker = [&] {
            cudaSetDevice(dev);
            size_t t = 0;
            for (;;) {
                const size_t p = next_piece.fetch_add(1);
                if (p >= n_pieces) return;
                while (t + 1 < xfers.size() && xfers[t + 1].piece0 <= p) t++;   // pieces are claimed in increasing order
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return n_enqueued > t || abort_copy; });
                    if (abort_copy) return;
                }
                if (cudaEventSynchronize(landed[t]) != cudaSuccess) copy_failed = true;
                const Xfer &x = xfers[t];
                const size_t off = (p - x.piece0) * kPieceBytes, len = std::min(kPieceBytes, x.len - off);
                if (!copy_failed) std::memcpy(x.dst + off, (const char *)pin[t % n_pin] + off, len);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (--pieces_left[t] == 0) {
                        drained[t] = 1;
                        while (n_drained < xfers.size() && drained[n_drained]) n_drained++;
                        cv.notify_all();
                    }
                }
            }
        };
        std::vector<std::thread> copiers;
        struct CopierJoin {  // every exit path below stops and joins the copy threads
            std::vector<std::thread> &v;
            std::mutex &mu;
            std::condition_variable &cv;
            bool &abort_flag;
            bool finished = false;
            ~CopierJoin() {
                if (!finished) {
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        abort_flag = true;
                    }
                    cv.notify_all();
                }
                for (auto &t : v)
                    if (t.joinable()) t.join();
            }
        } joiner{copiers, mu, cv, abort_copy};
        if (job.staged)
            for (int t = 0; t < job.copy_threads; t++) copiers.emplace_back(copy_worker);
        // transfers of chunk c: enqueued AFTER the launch of chunk c+1, so the host thread's waits for free pinned slots
        // never hold back a kernel
        auto enqueue_transfers = [&](size_t c) -> int {
            const int b = (int)(c % n_buf);
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c + 2], s_copy.s));
            for (size_t t = chunk_x0[c]; t < chunk_x0[c + 1]; t++) {
                if (t >= (size_t)n_pin) {  // slot t % n_pin must have been emptied by the copy threads
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return n_drained + n_pin > t; });
                }
                PPB_CUDA(cudaMemcpyAsync(pin[t % n_pin], xfers[t].d_src, xfers[t].len, cudaMemcpyDeviceToHost, s_copy.s));
                PPB_CUDA(cudaEventRecord(landed[t], s_copy.s));
                {
                    std::lock_guard<std::mutex> lk(mu);
                    n_enqueued = t + 1;
                }
                cv.notify_all();
            }
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c + 3], s_copy.s));
            PPB_CUDA(cudaEventRecord(done_copy[b].e, s_copy.s));
            return PPB_OK;
        };
        for (size_t c = 0; c < chunks.size(); c++) {
            const int b = (int)(c % n_buf);
            const int64_t r0 = chunks[c].first, r1 = chunks[c].second;
            if (c >= (size_t)n_buf) PPB_CUDA(cudaStreamWaitEvent(s_compute.s, done_copy[b].e, 0));  // buffer b drained
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c], s_compute.s));
            if (int rc = ppb_query_dev((const uint32_t *)d_ref, job.n_ref, job.self ? nullptr : (const uint32_t *)d_qry, n_q,
                                       job.kmers, K, ss64, (const float *)d_tab, job.C, (const uint16_t *)d_rc,
                                       (const uint16_t *)d_qc, r0, r1, job.out_mode, job.out ? d_out[b] : nullptr,
                                       job.boundary, job.labels ? (int8_t *)d_lab[b] : nullptr,
                                       (unsigned long long *)d_deg, s_compute.s))
                return rc;
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c + 1], s_compute.s));
            PPB_CUDA(cudaEventRecord(done_compute[b].e, s_compute.s));
            if (job.staged) {
                // (with a single device buffer the next launch needs this chunk's copies first)
                if (c > 0 && n_buf > 1)
                    if (int rc = enqueue_transfers(c - 1)) return rc;
                PPB_CUDA(cudaStreamWaitEvent(s_copy.s, done_compute[b].e, 0));
                if (n_buf == 1)
                    if (int rc = enqueue_transfers(c)) return rc;
                continue;
            }
            PPB_CUDA(cudaStreamWaitEvent(s_copy.s, done_compute[b].e, 0));
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c + 2], s_copy.s));
            if (job.out)
                PPB_CUDA(cudaMemcpyAsync(out_base + (size_t)r0 * rb, d_out[b], (size_t)(r1 - r0) * rb, cudaMemcpyDeviceToHost,
                                         s_copy.s));
            if (job.labels)
                PPB_CUDA(cudaMemcpyAsync(lab_base + r0, d_lab[b], (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s_copy.s));
            if (job.trace) PPB_CUDA(cudaEventRecord(tr[4 * c + 3], s_copy.s));
            PPB_CUDA(cudaEventRecord(done_copy[b].e, s_copy.s));
        }
        if (job.staged && n_buf > 1)
            if (int rc = enqueue_transfers(chunks.size() - 1)) return rc;
        unsigned long long deg = 0;
        PPB_CUDA(cudaMemcpyAsync(&deg, d_deg, 8, cudaMemcpyDeviceToHost, s_compute.s));
        PPB_CUDA(cudaStreamSynchronize(s_compute.s));
        PPB_CUDA(cudaStreamSynchronize(s_copy.s));
        job.deg[g] = deg;
        if (job.staged) {
            joiner.finished = true;
            for (auto &t : copiers) t.join();
            if (copy_failed) return fail(PPB_ERR_CUDA, "ppb_query_host: a device-to-host copy failed");
        }
        if (job.trace) {  // when each chunk's kernel and copies ran, relative to the first launch
            double k_sum = 0, c_sum = 0;
            float first_copy = 0, last_copy = 0;
            for (size_t c = 0; c < chunks.size(); c++) {
                float t[4];
                for (int e = 0; e < 4; e++) cudaEventElapsedTime(&t[e], tr_start, tr[4 * c + e]);
                k_sum += t[1] - t[0];
                c_sum += t[3] - t[2];
                if (c == 0) first_copy = t[2];
                last_copy = t[3];
                if (std::getenv("PPB_HOST_TRACE")[0] == '2')
                    std::fprintf(stderr, "[ppb_query_host dev %d] chunk %3zu rows %lld  kernel %8.2f..%8.2f ms  copy %8.2f..%8.2f ms\n",
                                 dev, c, (long long)(chunks[c].second - chunks[c].first), t[0], t[1], t[2], t[3]);
            }
            std::fprintf(stderr,
                         "[ppb_query_host dev %d] rows %lld in %zu chunks (device ring %d%s): kernels %.1f ms, copies %.1f ms "
                         "(first starts at %.1f, last ends at %.1f ms after the first launch)\n",
                         dev, (long long)(r_hi - r_lo), chunks.size(), n_buf, job.staged ? ", staged" : ", direct DMA", k_sum, c_sum,
                         first_copy, last_copy);
            for (auto &e : tr) cudaEventDestroy(e);
            cudaEventDestroy(tr_start);
        }
        return PPB_OK;
    };
    rc = phase_c();
    stamp("row chunks done");
    if (rc) {
        job.err[g] = g_err;
        cudaStreamSynchronize(s_compute.s);  // nothing of this call may still be running on the workspace
        cudaStreamSynchronize(s_copy.s);
        cudaGetLastError();
    }
    // peers may still be reading nothing of ours (their inputs were complete after phase B), but our packed_ev must
    // outlive their cudaStreamWaitEvent calls: meet once more before the events are destroyed by the caller
    rv.arrive(rc == PPB_OK);
    return rc;
}

// Static row shards of a host call: G contiguous ranges of (nearly) equal row count, cut where a row TILE of the row
// side begins (no tile is computed by two devices) — the condensed order is row-major in i, so equal-count shards are
// row bands of growing height and the load is balanced by pair count.
void plan_device_shards(int64_t n_ref, int64_t n_qry, int self, int64_t row_begin, int64_t row_end, int G,
                        std::vector<int64_t> *cut) {
    cut->assign((size_t)G + 1, row_end);
    (*cut)[0] = row_begin;
    const int64_t n_side = self ? n_ref : n_qry;
    const int64_t total_rows = self ? n_ref * (n_ref - 1) / 2 : n_ref * n_qry;
    auto first_row_of = [&](int64_t gi) -> int64_t {
        if (gi >= (self ? n_side - 1 : n_side)) return total_rows;
        return self ? sq2cond(gi, gi + 1, n_ref) : gi * n_ref;
    };
    const int64_t rows = row_end - row_begin;
    for (int g = 1; g < G; g++) {
        const int64_t ideal = row_begin + (int64_t)((__int128)rows * g / G);
        const int64_t gi = self ? row_idx(std::min(ideal, total_rows - 1), n_ref) : ideal / n_ref;
        // tile boundary at or below / above the ideal cut: take the nearer one
        const int64_t t0 = gi / ppb::kTI * ppb::kTI, t1 = t0 + ppb::kTI;
        const int64_t c0 = first_row_of(t0), c1 = first_row_of(t1);
        int64_t c = (ideal - c0 <= c1 - ideal) ? c0 : c1;
        c = std::max(c, (*cut)[g - 1]);
        c = std::min(std::max(c, row_begin), row_end);
        (*cut)[g] = c;
    }
}

}  // namespace

extern "C" {

int ppb_query_host_multi(const uint64_t *ref, int64_t n_ref, const uint64_t *qry, int64_t n_qry, const int32_t *kmers,
                         int32_t K, int32_t sketchsize64, int32_t bbits, const float *rand_table, int32_t n_clusters,
                         const uint16_t *ref_cluster, const uint16_t *qry_cluster, int64_t row_begin, int64_t row_end,
                         int32_t out_mode, void *out, const ppb_boundary *boundary, int8_t *labels,
                         int64_t *n_degenerate, const int32_t *device_ids, int32_t n_devices) {
    const auto t_call = std::chrono::steady_clock::now();
    if (bbits != PPB_BBITS) return fail(PPB_ERR_ARG, "ppb_query_host: bbits must be 14");
    if (!ref || !kmers || K < 1 || K > PPB_MAX_K || sketchsize64 < 1 || n_ref < 0)
        return fail(PPB_ERR_ARG, "ppb_query_host: bad argument");
    const int self = qry == nullptr;
    if (!self && n_qry < 0) return fail(PPB_ERR_ARG, "ppb_query_host: bad argument");
    const int64_t total_rows = ppb_num_rows(n_ref, n_qry, self);
    if (row_begin < 0 || row_end > total_rows || row_begin > row_end)
        return fail(PPB_ERR_ARG, "ppb_query_host: bad row range");
    if (out_mode < PPB_OUT_DISTS || out_mode > PPB_OUT_COUNTS) return fail(PPB_ERR_ARG, "ppb_query_host: bad out_mode");
    if (n_degenerate) *n_degenerate = 0;
    if (row_begin == row_end) return PPB_OK;
    if (!out && !(out_mode == PPB_OUT_DISTS && boundary && labels)) return fail(PPB_ERR_ARG, "ppb_query_host: no output buffer");
    if ((boundary != nullptr) != (labels != nullptr)) return fail(PPB_ERR_ARG, "ppb_query_host: boundary and labels go together");
    if (rand_table) {
        if (n_clusters < 1 || !ref_cluster || (!self && !qry_cluster))
            return fail(PPB_ERR_ARG, "ppb_query_host: random table without cluster ids");
        // the kernel indexes the table with these ids: an inconsistent database must not become an out-of-bounds read
        for (int64_t i = 0; i < n_ref; i++)
            if (ref_cluster[i] >= n_clusters) return fail(PPB_ERR_ARG, "ppb_query_host: reference cluster id out of range");
        for (int64_t i = 0; !self && i < n_qry; i++)
            if (qry_cluster[i] >= n_clusters) return fail(PPB_ERR_ARG, "ppb_query_host: query cluster id out of range");
    }
    const int ndev = ppb_device_count();
    if (ndev <= 0) return fail(PPB_ERR_NO_DEVICE, "ppb_query_host: no CUDA device (this engine has no CPU path)");
    if (n_devices < 1 || n_devices > PPB_MAX_PEERS || !device_ids) return fail(PPB_ERR_ARG, "ppb_query_host: bad device list");
    for (int g = 0; g < n_devices; g++) {
        if (device_ids[g] < 0 || device_ids[g] >= ndev) return fail(PPB_ERR_ARG, "ppb_query_host: bad device id");
        for (int h = 0; h < g; h++)
            if (device_ids[h] == device_ids[g]) return fail(PPB_ERR_ARG, "ppb_query_host: a device is listed twice");
    }
    DeviceGuard guard;
    // this entry reads and writes HOST memory (its copy threads memcpy into the destination); device-resident data
    // goes through ppb_query_dev
    for (const void *p : {(const void *)ref, (const void *)qry, (const void *)out, (const void *)labels})
        if (is_device_ptr(p))
            return fail(PPB_ERR_ARG, "ppb_query_host: a device pointer was passed where host memory is expected (use ppb_query_dev)");
    // one host-buffer call at a time per process (the reference's entry is not re-entrant either); two overlapping
    // multi-device calls could otherwise each hold one device's workspace and wait for the other's
    static std::mutex host_call_mu;
    std::lock_guard<std::mutex> host_call_lock(host_call_mu);

    HostJob job;
    job.ref = ref, job.qry = qry, job.n_ref = n_ref, job.n_qry = n_qry, job.kmers = kmers, job.K = K, job.ss64 = sketchsize64;
    job.rand_table = rand_table, job.C = n_clusters, job.ref_cluster = ref_cluster, job.qry_cluster = qry_cluster;
    job.row_begin = row_begin, job.row_end = row_end, job.out_mode = out_mode, job.out = out, job.boundary = boundary;
    job.labels = labels, job.self = self;
    // small jobs are not worth a second device: at least ~16 Mi rows per device
    int G = n_devices;
    if (const char *e = std::getenv("PPB_MIN_ROWS_PER_DEVICE")) {
        const int64_t m = std::max<int64_t>(1, atoll(e));
        G = (int)std::max<int64_t>(1, std::min<int64_t>(G, (row_end - row_begin) / m));
    } else {
        G = (int)std::max<int64_t>(1, std::min<int64_t>(G, (row_end - row_begin) >> 24));
    }
    // A pageable destination is drained by the host cores (first touch + copy: 30-45 GB/s on the B200 hosts measured,
    // tools/host_floor.cu), which is less than ONE device produces at S = 1024 with 8 B per pair: more devices would only
    // add their start-up cost (context, workspace, pinned rings: 4.5 s instead of 1.8 s for a first call with 8 devices).
    // Use as many devices as it takes to out-produce the host, plus one.
    const bool staged_dest = (out && !is_dma_able(out)) || (labels && !is_dma_able(labels));
    if (staged_dest && G > 1 && !std::getenv("PPB_STAGED_ALL_DEVICES")) {
        const double per_row = (out ? out_row_bytes(out_mode, K) : 0) + (labels ? 1 : 0);
        const double rows_per_s = 15e12 / ((double)K * 2.0 * sketchsize64 * 14.0);   // ~0.8 of the LOP3 pipe of one B200
        const int enough = (int)std::ceil(40e9 / (rows_per_s * per_row)) + 1;
        G = std::max(1, std::min(G, enough));
    }
    job.G = G;
    job.devs.assign(device_ids, device_ids + G);
    plan_device_shards(n_ref, n_qry, self, row_begin, row_end, G, &job.row_cut);
    // peer-to-peer scatter of the packed reference array: every pair of devices must be able to map the other
    job.p2p = G > 1;
    for (int g = 0; g < G && job.p2p; g++)
        for (int h = 0; h < G && job.p2p; h++) {
            int can = 0;
            if (g != h && (cudaDeviceCanAccessPeer(&can, job.devs[g], job.devs[h]) != cudaSuccess || !can)) job.p2p = false;
        }
    if (std::getenv("PPB_NO_P2P")) job.p2p = false;
    if (job.p2p)
        for (int g = 0; g < G; g++) {
            PPB_CUDA(cudaSetDevice(job.devs[g]));
            for (int h = 0; h < G; h++)
                if (g != h) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(job.devs[h], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) job.p2p = false;
                    cudaGetLastError();
                }
        }
    const int64_t n_pad = round_up(std::max<int64_t>(n_ref, 1), ppb::kPad);
    job.gen_cut.assign((size_t)G + 1, n_pad);
    for (int g = 0; g < G; g++) job.gen_cut[g] = std::min<int64_t>(n_pad, round_up((int64_t)((__int128)n_pad * g / G), 4));
    job.staged = staged_dest;
    job.trace = std::getenv("PPB_HOST_TRACE") != nullptr;
    job.t0 = t_call;
    int hw = (int)std::thread::hardware_concurrency();
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = CPU_COUNT(&set);
    }
    // copy threads per device (staged results, pageable uploads): all cores but one per device worker and one spare
    job.copy_threads = std::max(2, std::min(16, (hw - 1 - G) / G));
    if (const char *e = std::getenv("PPB_COPY_THREADS")) job.copy_threads = std::max(1, atoi(e));
    if (job.staged) {
        if (out) advise_hugepages(out, (size_t)(row_end - row_begin) * out_row_bytes(out_mode, K));
        if (labels) advise_hugepages(labels, (size_t)(row_end - row_begin));
    }
    job.d_ref_packed.assign(G, nullptr);
    job.packed_ev.assign(G, nullptr);
    job.rc.assign(G, PPB_OK);
    job.err.assign(G, "");
    job.deg.assign(G, 0);
    Rendezvous rv(G);
    job.rv = &rv;

    if (G == 1) {
        job.rc[0] = host_worker(job, 0);
    } else {
        std::vector<std::thread> workers;
        for (int g = 0; g < G; g++) workers.emplace_back([&job, g] { job.rc[g] = host_worker(job, g); });
        for (auto &t : workers) t.join();
    }
    for (int g = 0; g < G; g++)
        if (job.packed_ev[g]) cudaEventDestroy(job.packed_ev[g]);
    unsigned long long deg = 0;
    for (int g = 0; g < G; g++) {
        if (job.rc[g] != PPB_OK && !job.err[g].empty())
            return fail(job.rc[g], "device " + std::to_string(job.devs[g]) + ": " + job.err[g]);
        deg += job.deg[g];
    }
    for (int g = 0; g < G; g++)
        if (job.rc[g] != PPB_OK) return fail(job.rc[g], "ppb_query_host: a device worker failed");
    if (n_degenerate) *n_degenerate = (int64_t)deg;
    return PPB_OK;
}

int ppb_query_host(const uint64_t *ref, int64_t n_ref, const uint64_t *qry, int64_t n_qry, const int32_t *kmers,
                   int32_t K, int32_t sketchsize64, int32_t bbits, const float *rand_table, int32_t n_clusters,
                   const uint16_t *ref_cluster, const uint16_t *qry_cluster, int64_t row_begin, int64_t row_end,
                   int32_t out_mode, void *out, const ppb_boundary *boundary, int8_t *labels,
                   int64_t *n_degenerate, int32_t device_id) {
    return ppb_query_host_multi(ref, n_ref, qry, n_qry, kmers, K, sketchsize64, bbits, rand_table, n_clusters, ref_cluster,
                                qry_cluster, row_begin, row_end, out_mode, out, boundary, labels, n_degenerate, &device_id, 1);
}

int64_t ppb_plan_device_shards(int64_t n_ref, int64_t n_qry, int32_t self, int64_t row_begin, int64_t row_end,
                               int32_t n_devices, int64_t *cuts) {
    const int64_t total_rows = ppb_num_rows(n_ref, n_qry, self);
    if (n_ref < 0 || (!self && n_qry < 0) || row_begin < 0 || row_end > total_rows || row_begin > row_end || n_devices < 1 || !cuts)
        return -1;
    std::vector<int64_t> cut;
    plan_device_shards(n_ref, n_qry, self, row_begin, row_end, n_devices, &cut);
    for (int g = 0; g <= n_devices; g++) cuts[g] = cut[g];
    return n_devices;
}

void *ppb_host_alloc(size_t bytes) {
    const size_t want = (std::max<size_t>(bytes, 1) + (((size_t)2 << 20) - 1)) & ~(((size_t)2 << 20) - 1);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto best = g_pool.end();
    for (auto it = g_pool.begin(); it != g_pool.end(); ++it)
        if (!it->second.in_use && it->second.cap >= want && it->second.cap <= 2 * want + ((size_t)64 << 20) &&
            (best == g_pool.end() || it->second.cap < best->second.cap))
            best = it;
    if (best != g_pool.end()) {
        HostBlock &b = best->second;
        b.in_use = true;
        b.reuses++;
        // Page-lock a block on its SECOND reuse (PPB_HOST_PIN_AFTER, 0 = never).  Registering 40 GB costs 2-8 s (more with
        // more CUDA contexts); a first reuse is already served well by staging into the block's touched pages (no page
        // faults: the copy threads move ~75 GB/s, above one GPU's PCIe rate), so a process that calls two or three times
        // never pays for the registration and one that keeps calling pays once.
        int pin_after = 2;
        if (const char *e = std::getenv("PPB_HOST_PIN_AFTER")) pin_after = atoi(e);
        if (const char *e = std::getenv("PPB_HOST_PIN")) if (e[0] == '0') pin_after = 0;
        if (!b.pinned && b.touched && pin_after > 0 && b.reuses >= pin_after && ppb_device_count() > 0) {
            if (cudaHostRegister(best->first, b.cap, cudaHostRegisterPortable) == cudaSuccess)
                b.pinned = true;
            else
                cudaGetLastError();  // stays pageable: the staged path handles it
        }
        return best->first;
    }
    // nothing to reuse: drop idle blocks first so the pool never holds more than its limit
    size_t held = 0;
    for (auto &kv : g_pool) held += kv.second.cap;
    for (auto it = g_pool.begin(); it != g_pool.end() && held + want > pool_limit_bytes();) {
        auto cur = it++;
        if (!cur->second.in_use) {
            held -= cur->second.cap;
            pool_drop(cur);
        }
    }
    void *p = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        fail(PPB_ERR_NOMEM, "ppb_host_alloc: mmap failed for " + std::to_string(want) + " bytes");
        return nullptr;
    }
    madvise(p, want, MADV_HUGEPAGE);
    HostBlock b;
    b.cap = want;
    b.in_use = true;
    g_pool[p] = b;
    return p;
}

int ppb_host_free(void *p) {
    if (!p) return PPB_OK;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool.find(p);
    if (it == g_pool.end() || !it->second.in_use) return fail(PPB_ERR_ARG, "ppb_host_free: not a live ppb_host_alloc block");
    it->second.in_use = false;
    it->second.touched = true;
    size_t idle = 0;
    for (auto &kv : g_pool)
        if (!kv.second.in_use) idle += kv.second.cap;
    if (idle > pool_limit_bytes()) pool_drop(it);
    return PPB_OK;
}

int ppb_host_pool_stats(size_t *bytes_held, size_t *bytes_in_use, size_t *bytes_pinned) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t h = 0, u = 0, pn = 0;
    for (auto &kv : g_pool) {
        h += kv.second.cap;
        if (kv.second.in_use) u += kv.second.cap;
        if (kv.second.pinned) pn += kv.second.cap;
    }
    if (bytes_held) *bytes_held = h;
    if (bytes_in_use) *bytes_in_use = u;
    if (bytes_pinned) *bytes_pinned = pn;
    return PPB_OK;
}

int ppb_release_workspace(void) {
    DeviceGuard guard;
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (auto &kv : g_ws) {
        std::lock_guard<std::mutex> lk2(kv.second->mu);
        if (cudaSetDevice(kv.first) == cudaSuccess) kv.second->release();
    }
    g_ws.clear();
    {   // idle result blocks go back to the OS as well (blocks still owned by a live array stay)
        std::lock_guard<std::mutex> lk3(g_pool_mu);
        for (auto it = g_pool.begin(); it != g_pool.end();) {
            auto cur = it++;
            if (!cur->second.in_use) pool_drop(cur);
        }
    }
    if (guard.prev >= 0) {
        cudaMemPool_t pool;  // the stream-ordered scratch cached by the iteration / kNN entry points
        if (cudaSetDevice(guard.prev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, guard.prev) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
    }
    cudaGetLastError();
    return PPB_OK;
}

int ppb_assign_threshold_host(const float *dists, int64_t n, int32_t slope, float x_max, float y_max, float *out,
                              int32_t device_id) {
    if (n < 0 || (n > 0 && (!dists || !out))) return fail(PPB_ERR_ARG, "ppb_assign_threshold_host: bad argument");
    if (n == 0) return PPB_OK;
    int ndev = ppb_device_count();
    if (ndev <= 0) return fail(PPB_ERR_NO_DEVICE, "ppb_assign_threshold_host: no CUDA device (no CPU path)");
    if (device_id < 0 || device_id >= ndev) return fail(PPB_ERR_ARG, "ppb_assign_threshold_host: bad device id");
    DeviceGuard guard;
    PPB_CUDA(cudaSetDevice(device_id));
    DevBuf d_in, d_o;
    if (int rc = d_in.alloc((size_t)n * 8)) return rc;
    if (int rc = d_o.alloc((size_t)n * 4)) return rc;
    PPB_CUDA(cudaMemcpy(d_in.p, dists, (size_t)n * 8, cudaMemcpyHostToDevice));
    if (int rc = ppb_assign_threshold_dev((const float *)d_in.p, n, slope, x_max, y_max, (float *)d_o.p, nullptr)) return rc;
    PPB_CUDA(cudaMemcpy(out, d_o.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return PPB_OK;
}

}  // extern "C"
