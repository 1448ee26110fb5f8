// host_floor — what the HOST side of a B200 box can take: the floor under the end-to-end time of the distance path
// (its result is 8 B/pair that must land in host memory).
//
//   nvcc -O2 -std=c++17 -o /tmp/host_floor tools/host_floor.cu -lpthread && /tmp/host_floor [GiB per device = 4] [n_dev = all]
//
// Prints one JSON object per measurement (stdout):
//   pinned D2H per device alone and with all devices at once (aggregate host ingest), H2D likewise;
//   cudaHostAlloc / cudaHostRegister cost per GiB (with and without transparent huge pages);
//   first-touch (page-fault) rate of fresh pageable memory with T threads, with / without MADV_HUGEPAGE;
//   memcpy pinned -> fresh pageable and -> already-touched pageable with T threads.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void *fresh(size_t bytes, bool huge) {
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        perror("mmap");
        std::exit(1);
    }
    madvise(p, bytes, huge ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
    return p;
}
template <typename F>
static void par(int threads, size_t bytes, F f) {  // f(offset, length) on `threads` threads
    const size_t piece = ((bytes + threads - 1) / threads + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        const size_t off = (size_t)t * piece;
        if (off >= bytes) break;
        pool.emplace_back([=] { f(off, std::min(piece, bytes - off)); });
    }
    for (auto &th : pool) th.join();
}
static std::string slurp(const char *path) {
    FILE *f = std::fopen(path, "r");
    if (!f) return "?";
    char buf[256] = {0};
    size_t n = std::fread(buf, 1, 255, f);
    std::fclose(f);
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}

int main(int argc, char **argv) {
    const double gib_arg = argc > 1 ? atof(argv[1]) : 4.0;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (argc > 2) ndev = std::min(ndev, atoi(argv[2]));
    const bool quick = argc > 3 && std::string(argv[3]) == "quick";  // multi-GPU boxes are charged per GPU: the essentials only
    const size_t bytes = (size_t)(gib_arg * (double)((size_t)1 << 30));
    const double GB = 1e9;
    const int hw = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    CPU_ZERO(&set);
    sched_getaffinity(0, sizeof(set), &set);
    std::printf("{\"what\":\"host\",\"hardware_concurrency\":%d,\"affinity_cpus\":%d,\"n_dev\":%d,\"thp_enabled\":\"%s\",\"thp_defrag\":\"%s\",\"bytes_per_test\":%zu}\n",
                hw, CPU_COUNT(&set), ndev, slurp("/sys/kernel/mm/transparent_hugepage/enabled").c_str(),
                slurp("/sys/kernel/mm/transparent_hugepage/defrag").c_str(), bytes);
    std::fflush(stdout);

    // ---- pinned allocation cost
    std::vector<void *> h_pin(ndev), d_buf(ndev);
    std::vector<cudaStream_t> st(ndev);
    for (int d = 0; d < ndev; d++) {
        CK(cudaSetDevice(d));
        CK(cudaFree(0));
    }
    for (int d = 0; d < ndev; d++) {
        CK(cudaSetDevice(d));
        const double t0 = now();
        CK(cudaHostAlloc(&h_pin[d], bytes, cudaHostAllocPortable));
        const double dt = now() - t0;
        if (d == 0) std::printf("{\"what\":\"cudaHostAlloc\",\"GiB\":%.1f,\"s\":%.3f,\"GBps\":%.2f}\n", gib_arg, dt, bytes / dt / GB);
        CK(cudaMalloc(&d_buf[d], bytes));
        CK(cudaMemset(d_buf[d], 1, bytes));
        CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
        CK(cudaDeviceSynchronize());
    }
    std::fflush(stdout);
    // ---- pinned copies: each device alone, then all at once
    auto copy_rate = [&](int d0, int d1, bool d2h, int reps) {
        const double t0 = now();
        for (int r = 0; r < reps; r++)
            for (int d = d0; d < d1; d++) {
                CK(cudaSetDevice(d));
                if (d2h)
                    CK(cudaMemcpyAsync(h_pin[d], d_buf[d], bytes, cudaMemcpyDeviceToHost, st[d]));
                else
                    CK(cudaMemcpyAsync(d_buf[d], h_pin[d], bytes, cudaMemcpyHostToDevice, st[d]));
            }
        for (int d = d0; d < d1; d++) {
            CK(cudaSetDevice(d));
            CK(cudaStreamSynchronize(st[d]));
        }
        return (double)reps * (d1 - d0) * bytes / (now() - t0) / GB;
    };
    for (int d = 0; d < ndev; d++) {
        copy_rate(d, d + 1, true, 1);
        std::printf("{\"what\":\"pinned_d2h_alone\",\"dev\":%d,\"GBps\":%.2f}\n", d, copy_rate(d, d + 1, true, 3));
    }
    std::printf("{\"what\":\"pinned_h2d_alone\",\"dev\":0,\"GBps\":%.2f}\n", copy_rate(0, 1, false, 3));
    for (int k = 2; k <= ndev; k *= 2)
        std::printf("{\"what\":\"pinned_d2h_concurrent\",\"n_dev\":%d,\"aggregate_GBps\":%.2f}\n", k, copy_rate(0, k, true, 3));
    std::fflush(stdout);

    // ---- fresh pageable memory: first touch, memcpy from pinned, cudaHostRegister, driver-staged D2H
    CK(cudaSetDevice(0));
    for (int huge = 0; huge <= 1; huge++) {
        for (int T : {1, 4, 8, 16, 32, 64}) {
            if (T > 2 * hw || (quick && T != 8 && T != 16 && T != hw)) continue;
            void *p = fresh(bytes, huge);
            double t0 = now();
            par(T, bytes, [&](size_t off, size_t len) {
                for (size_t o = 0; o < len; o += 4096) ((volatile char *)p)[off + o] = 1;
            });
            const double touch = now() - t0;
            t0 = now();
            par(T, bytes, [&](size_t off, size_t len) { std::memcpy((char *)p + off, (char *)h_pin[0] + off, len); });
            const double warm = now() - t0;
            munmap(p, bytes);
            p = fresh(bytes, huge);
            t0 = now();
            par(T, bytes, [&](size_t off, size_t len) { std::memcpy((char *)p + off, (char *)h_pin[0] + off, len); });
            const double cold = now() - t0;
            munmap(p, bytes);
            std::printf("{\"what\":\"pageable\",\"madv_hugepage\":%d,\"threads\":%d,\"first_touch_GBps\":%.2f,\"memcpy_pinned_to_touched_GBps\":%.2f,\"memcpy_pinned_to_fresh_GBps\":%.2f}\n",
                        huge, T, bytes / touch / GB, bytes / warm / GB, bytes / cold / GB);
            std::fflush(stdout);
        }
        if (!quick) {   // cudaHostRegister of fresh / touched memory, then D2H straight into it
            void *p = fresh(bytes, huge);
            double t0 = now();
            cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
            const double reg_fresh = now() - t0;
            double d2h = 0;
            if (e == cudaSuccess) {
                t0 = now();
                CK(cudaMemcpyAsync(p, d_buf[0], bytes, cudaMemcpyDeviceToHost, st[0]));
                CK(cudaStreamSynchronize(st[0]));
                d2h = bytes / (now() - t0) / GB;
                t0 = now();
                CK(cudaHostUnregister(p));
            }
            const double unreg = now() - t0;
            munmap(p, bytes);
            p = fresh(bytes, huge);
            par(std::min(16, hw), bytes, [&](size_t off, size_t len) {
                for (size_t o = 0; o < len; o += 4096) ((volatile char *)p)[off + o] = 1;
            });
            t0 = now();
            e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
            const double reg_touched = now() - t0;
            if (e == cudaSuccess) CK(cudaHostUnregister(p));
            // chunked registration (what a pipelined "register ahead of the copies" scheme would pay): 256 MiB pieces
            munmap(p, bytes);
            p = fresh(bytes, huge);
            t0 = now();
            const size_t piece = (size_t)256 << 20;
            for (size_t off = 0; off < bytes; off += piece) cudaHostRegister((char *)p + off, std::min(piece, bytes - off), cudaHostRegisterPortable);
            const double reg_chunks = now() - t0;
            for (size_t off = 0; off < bytes; off += piece) cudaHostUnregister((char *)p + off);
            cudaGetLastError();
            munmap(p, bytes);
            std::printf("{\"what\":\"cudaHostRegister\",\"madv_hugepage\":%d,\"fresh_GBps\":%.2f,\"touched_GBps\":%.2f,\"fresh_256MiB_pieces_GBps\":%.2f,\"unregister_GBps\":%.2f,\"d2h_into_registered_GBps\":%.2f,\"ok\":%d}\n",
                        huge, bytes / reg_fresh / GB, bytes / reg_touched / GB, bytes / reg_chunks / GB, bytes / unreg / GB, d2h, e == cudaSuccess);
            std::fflush(stdout);
        }
        if (!quick) {   // the driver's own staging: cudaMemcpy D2H into fresh pageable memory
            void *p = fresh(bytes, huge);
            double t0 = now();
            CK(cudaMemcpy(p, d_buf[0], bytes, cudaMemcpyDeviceToHost));
            const double cold = now() - t0;
            t0 = now();
            CK(cudaMemcpy(p, d_buf[0], bytes, cudaMemcpyDeviceToHost));
            const double warm = now() - t0;
            munmap(p, bytes);
            std::printf("{\"what\":\"cudaMemcpy_d2h_pageable\",\"madv_hugepage\":%d,\"fresh_GBps\":%.2f,\"touched_GBps\":%.2f}\n", huge,
                        bytes / cold / GB, bytes / warm / GB);
            std::fflush(stdout);
        }
    }
    // ---- the pipelined staged path as ppb_query_host runs it: D2H into a pinned ring while T threads drain it into fresh pages
    for (int huge = quick ? 1 : 0; huge <= 1; huge++)
        for (int T : {4, 8, 16, 32}) {
            if (T > hw || (quick && T < 16)) continue;
            void *p = fresh(bytes, huge);
            const size_t piece = (size_t)128 << 20;
            const int ring = 4;
            const size_t n_pieces = bytes / piece;
            std::vector<cudaEvent_t> ev(n_pieces);
            for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            const double t0 = now();
            size_t drained = 0;
            for (size_t c = 0; c < n_pieces + ring; c++) {
                if (c >= (size_t)ring) {   // drain piece c - ring (its slot is about to be reused)
                    const size_t q = c - ring;
                    CK(cudaEventSynchronize(ev[q]));
                    char *src = (char *)h_pin[0] + (q % ring) * piece, *dst = (char *)p + q * piece;
                    par(T, piece, [&](size_t off, size_t len) { std::memcpy(dst + off, src + off, len); });
                    drained++;
                }
                if (c < n_pieces) {
                    CK(cudaMemcpyAsync((char *)h_pin[0] + (c % ring) * piece, (char *)d_buf[0] + c * piece, piece, cudaMemcpyDeviceToHost, st[0]));
                    CK(cudaEventRecord(ev[c], st[0]));
                }
            }
            const double dt = now() - t0;
            for (auto &e : ev) cudaEventDestroy(e);
            munmap(p, bytes);
            std::printf("{\"what\":\"staged_pipeline_to_fresh_pageable\",\"madv_hugepage\":%d,\"threads\":%d,\"GBps\":%.2f}\n", huge, T,
                        n_pieces * piece / dt / GB);
            std::fflush(stdout);
        }
    return 0;
}
