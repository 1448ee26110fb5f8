#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels around the distance path (CUDA events, device-resident data):
pack, assign_threshold, edge compaction, boundary iteration, long<->square, kNN.  One JSON line per kernel:
algorithmic bytes (what must be read + written once), time, GB/s and the fraction of the HBM peak.

    python tools/hbm_kernels.py [peak_GBps=6650]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poppunk_b200 import _lib, engine, synth  # noqa: E402

PEAK = float(sys.argv[1]) if len(sys.argv) > 1 else 6650.0
L = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, nbytes, ms, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / PEAK, 3), "peak_GBps": PEAK, "note": note}), flush=True)


kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
# pack: read the canonical array once, write the packed array once
n = 100_000
sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device=dev)
report("pack_kernel (N=100k, S=1024, K=5)", 2 * sk.numel() * 8, timed(lambda: engine.pack(sk)))
del sk
# the (n_pairs, 2) consumers: 400M rows (3.2 GB)
rows = 400_000_000 // 2 * 2
n_samples = 28_285
rows = n_samples * (n_samples - 1) // 2
g = torch.Generator(device=dev)
g.manual_seed(1)
d = torch.rand((rows, 2), device=dev, generator=g) * 0.5
lab = torch.empty(rows, dtype=torch.float32, device=dev)
report(f"threshold_kernel ({rows} rows)", rows * 12,
       timed(lambda: _lib.check(L.ppb_assign_threshold_dev(d.data_ptr(), rows, 2, C.c_float(0.2), C.c_float(0.3), lab.data_ptr(), st))))
oi = torch.empty(rows, dtype=torch.int64, device=dev)
oj = torch.empty(rows, dtype=torch.int64, device=dev)
oo = torch.empty(rows, dtype=torch.int64, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
scratch = torch.empty(L.ppb_edges_scratch_bytes(rows), dtype=torch.uint8, device=dev)
ms = timed(lambda: _lib.check(L.ppb_edges_from_dists_dev(d.data_ptr(), rows, n_samples, 2, C.c_float(0.05), C.c_float(0.08),
                                                        oi.data_ptr(), oj.data_ptr(), rows, cnt.data_ptr(), scratch.data_ptr(), st)))
ne = int(cnt.item())
report(f"edge compaction: mark + scan + emit ({rows} rows -> {ne} edges)", rows * 8 + ne * 16, ms, "one read of the rows; the emit pass reads a bit per row")
xm = np.linspace(0.01, 0.3, 30).astype(np.float32)
ms = timed(lambda: _lib.check(L.ppb_threshold_iterate_2d_dev(d.data_ptr(), rows, xm.ctypes.data, len(xm), C.c_float(0.3), oi.data_ptr(),
                                                            oj.data_ptr(), oo.data_ptr(), rows, cnt.data_ptr(), st)))
ne = int(cnt.item())
report(f"thresholdIterate2D: 30 steps ({rows} rows -> {ne} edges)", rows * 8 + ne * 24, ms, "one read of the rows (classify) + 1 B/row note written and read + stable scatter by step")
offs = np.linspace(0.01, 0.3, 30)
ms = timed(lambda: _lib.check(L.ppb_threshold_iterate_1d_dev(d.data_ptr(), rows, offs.ctypes.data, len(offs), 2, C.c_float(0.0), C.c_float(0.0),
                                                            C.c_float(0.3), C.c_float(0.3), oi.data_ptr(), oj.data_ptr(), oo.data_ptr(), rows,
                                                            cnt.data_ptr(), st)), reps=3)
ne = int(cnt.item())
report(f"thresholdIterate1D: 30 steps ({rows} rows -> {ne} edges)", rows * 8 + ne * 24, ms,
       "algorithmic bytes = one read of the rows + the edges; the work in between: classify (6 B/row written), compaction of the "
       "admitted rows, 4-pass radix sort of their (key, row) pairs, walk")
ne = min(ne, rows)
vals = torch.randint(0, rows, (ne,), device=dev, dtype=torch.int64, generator=g)
work = torch.empty_like(vals)
def sort_rows():
    work.copy_(vals)
    _lib.check(L.ppb_sort_rows_dev(work.data_ptr(), ne, rows - 1, st))
ms = timed(sort_rows, reps=3)
report(f"sort_rows (radix, {ne} int64 keys < {rows}: 4 passes)", ne * 16 + ne * 8 * 3 * 4, ms, "copy + 4 x (count read, emit read + write)")
del vals, work
del oi, oj, oo, lab
# long <-> square
sq = torch.empty((n_samples, n_samples), dtype=torch.float32, device=dev)
report(f"long_to_square_kernel (n={n_samples})", rows * 4 + n_samples * n_samples * 4,
       timed(lambda: _lib.check(L.ppb_long_to_square_dev(d.data_ptr(), 2, n_samples, sq.data_ptr(), st))), "strided column read of the (n,2) array")
vec = torch.empty(rows, dtype=torch.float32, device=dev)
report(f"square_to_long_kernel (n={n_samples})", rows * 8,
       timed(lambda: _lib.check(L.ppb_square_to_long_dev(sq.data_ptr(), n_samples, vec.data_ptr(), st))))
# kNN: 5 passes over each row (4 select + 1 collect), rows stay in L2
k = 10
ki = torch.empty(n_samples * k, dtype=torch.int64, device=dev)
kj = torch.empty(n_samples * k, dtype=torch.int64, device=dev)
kd = torch.empty(n_samples * k, dtype=torch.float32, device=dev)
report(f"knn_small_kernel (n={n_samples}, kNN={k})", n_samples * n_samples * 4 + n_samples * k * 20,
       timed(lambda: _lib.check(L.ppb_knn_dev(sq.data_ptr(), n_samples, n_samples, k, ki.data_ptr(), kj.data_ptr(), kd.data_ptr(), st))),
       "one pass over the matrix (kNN <= 32)")
k = 100
ki = torch.empty(n_samples * k, dtype=torch.int64, device=dev)
kj = torch.empty(n_samples * k, dtype=torch.int64, device=dev)
kd = torch.empty(n_samples * k, dtype=torch.float32, device=dev)
report(f"knn_kernel (n={n_samples}, kNN={k}: radix select)", n_samples * n_samples * 4 + n_samples * k * 20,
       timed(lambda: _lib.check(L.ppb_knn_dev(sq.data_ptr(), n_samples, n_samples, k, ki.data_ptr(), kj.data_ptr(), kd.data_ptr(), st))),
       "the matrix is read from HBM once and 4 more times from L2")
