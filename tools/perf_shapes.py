#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on ONE GPU (not the driver's bench contract — see bench.py).

    python tools/perf_shapes.py [cfg2] [cfg4] [cfg5] [cfg1like]

cfg2      10k genomes, S=1024, k={15,19,23,27,31}, self                      (full config)
cfg4      poppunk_assign shape: 1/8 of the 1M-query x 50k-ref rectangle (what one of 8 GPUs gets),
          fused assign_threshold, int8 labels only (the float2 output of the full job would be 400 GB)
cfg5      1/8 of the row range of 50k genomes at S=16384 (high-resolution sketches), self
cfg1like  29 genomes, sketch size 10000 -> sketchsize64=156, k=13..29 step 3 (shape of the repo's smoke test)
Each line: pairs/s, LOP3-equivalent rate, parity of a row sample against the CPU oracle.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from poppunk_b200 import engine, synth  # noqa: E402
import oracle  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, rows, ms, K, ss64, extra):
    lop3 = rows * K * ss64 * 2 * 14
    print(json.dumps({"config": name, "rows": rows, "ms": ms, "pairs_per_s": rows / (ms * 1e-3),
                      "lop3_per_s": lop3 / (ms * 1e-3), **extra}), flush=True)


def cfg2():
    kmers = np.array([15, 19, 23, 27, 31], dtype=np.int32)
    sk = synth.synth_sketches_torch(10_000, kmers, 16, seed=1, device="cuda")
    packed = engine.pack(sk)
    out = torch.empty((engine.num_rows(10_000), 2), dtype=torch.float32, device="cuda")
    ms = timed(lambda: engine.query(packed, None, kmers, out=out))
    host = sk.cpu().numpy().view(np.uint64)
    exp, _ = oracle.query(host, None, kmers, row_begin=1_000_000, row_end=1_100_000)
    err = float(np.abs(out[1_000_000:1_100_000].cpu().numpy() - exp).max())
    report("cfg2: N=10k S=1024 K=5 self", out.shape[0], ms, 5, 16, {"max_abs_err_vs_oracle": err})


def cfg4():
    kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
    R, Q = 50_000, 125_000
    ref = engine.pack(synth.synth_sketches_torch(R, kmers, 16, seed=2, device="cuda"))
    qsk = synth.synth_sketches_torch(Q, kmers, 16, seed=2, device="cuda")   # same population as the refs
    qry = engine.pack(qsk)
    rows = R * Q
    labels = torch.empty(rows, dtype=torch.int8, device="cuda")
    bnd = (2, 0.02, 0.25, 1.0, 1.0)
    ms = timed(lambda: engine.query(ref, qry, kmers, boundary=bnd, want_out=False, labels=labels), reps=2)
    within = int((labels[:50_000_000] == -1).sum())
    report("cfg4 (1/8): 125k queries x 50k refs, fused assign_threshold -> int8 labels only", rows, ms, 5, 16,
           {"within_boundary_in_first_50M": within, "out_bytes_per_pair": 1})


def cfg5():
    kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
    n, ss64 = 50_000, 256
    sk = synth.synth_sketches_torch(n, kmers, ss64, seed=3, device="cuda", chunk=512)
    packed = engine.pack(sk)
    total = engine.num_rows(n)
    b, e, _ = engine.shard_rows(total, 8, 3)
    out = torch.empty((e - b, 2), dtype=torch.float32, device="cuda")
    ms = timed(lambda: engine.query(packed, None, kmers, row_begin=b, row_end=e, out=out), reps=2)
    host = sk.cpu().numpy().view(np.uint64)
    exp, _ = oracle.query(host, None, kmers, row_begin=b, row_end=b + 20_000)
    err = float(np.abs(out[:20_000].cpu().numpy() - exp).max())
    report("cfg5 (1/8): N=50k S=16384 K=5 self, rows of rank 3 of 8", e - b, ms, 5, ss64,
           {"max_abs_err_vs_oracle": err})


def cfg1like():
    kmers = np.arange(13, 30, 3, dtype=np.int32)
    sk = synth.synth_sketches(29, kmers, 156, seed=4)
    t0 = time.perf_counter()
    out, _, ndeg = engine.query_host(sk, None, kmers)
    dt = time.perf_counter() - t0
    exp, _ = oracle.query(sk, None, kmers)
    report("cfg1-like: 29 genomes, sketchsize64=156, k=13..28 step 3 (host call incl. copies)", out.shape[0], dt * 1e3,
           len(kmers), 156, {"max_abs_err_vs_oracle": float(np.abs(out - exp).max()), "n_degenerate": ndeg})


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1like", "cfg2", "cfg4", "cfg5"]
    for name in which:
        globals()[name]()
        torch.cuda.empty_cache()
