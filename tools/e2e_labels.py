#!/usr/bin/env python
"""End-to-end time of the streamed-labels form of the host call (query-vs-reference, boundary fused, labels only:
engine.query_host(..., boundary, want_out=False)) from one process on g devices, call after call, with the state of
the host pool after each (so a block that failed to page-lock shows):
    python tools/e2e_labels.py [Q per device=125000] [R=50000] [device counts, e.g. 1,2] [calls per count=6] [Q total]
With "Q total" every device count works on the same query set (and so on the same pool block).  One JSON line per call; PPB_HOST_TRACE=1 adds the per-device kernel / copy timeline on stderr."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import _lib, engine, synth  # noqa: E402

q_per_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
L = _lib.load()
counts = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1]
n_calls = int(sys.argv[4]) if len(sys.argv) > 4 else 6
q_total = int(sys.argv[5]) if len(sys.argv) > 5 else 0
kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
table = synth.random_match_table(kmers, 3)
bnd = (2, 0.012, 0.15, 1.0, 1.0)


def pool():
    held, used, pinned = C.c_size_t(), C.c_size_t(), C.c_size_t()
    L.ppb_host_pool_stats(C.byref(held), C.byref(used), C.byref(pinned))
    return {"held_GB": round(held.value / 1e9, 2), "in_use_GB": round(used.value / 1e9, 2), "pinned_GB": round(pinned.value / 1e9, 2)}


for g in counts:
    if not 1 <= g <= L.ppb_device_count():
        continue
    Q = q_total or q_per_dev * g
    pop = dict(n_roots=1, n_lineages=8)
    if not (q_total and g != counts[0]):
        rh = synth.synth_sketches_torch(R, kmers, 16, seed=42, device="cuda:0", **pop).cpu().numpy().view(np.uint64)
        qh = synth.synth_sketches_torch(Q, kmers, 16, seed=43, device="cuda:0", **pop).cpu().numpy().view(np.uint64)
        torch.cuda.empty_cache()
        rcl, qcl = synth.synth_clusters(R, 3), synth.synth_clusters(Q, 3, seed=11)
        L.ppb_release_workspace()
        first = None
    os.environ["PPB_DEVICES"] = str(g)
    for call in range(n_calls if g == counts[0] else max(1, n_calls - 2)):
        t0 = time.perf_counter()
        _, lab, nd = engine.query_host(rh, qh, kmers, table, rcl, qcl, boundary=bnd, want_out=False,
                                       devices=engine.visible_devices(0))
        dt = time.perf_counter() - t0
        head = lab[:1_000_000].astype(np.int64)
        first = head if first is None else first
        print(json.dumps({"call": call, "n_dev": g, "Q": Q, "R": R, "ms": round(dt * 1e3, 1),
                          "Gpairs_per_s": round(Q * R / dt / 1e9, 2), "same_as_first": bool((head == first).all()),
                          "within": int((head == -1).sum()), "pool_during": pool()}), flush=True)
        del lab
