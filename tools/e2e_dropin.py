#!/usr/bin/env python
"""End-to-end time of the drop-in's own call (poppunk_b200.sketchlib.query_arrays = pp_queryDatabase once the sketches
are in memory) on 1..all GPUs of the box, with the destinations a PopPUNK process can end up with:
    python tools/e2e_dropin.py [N=100000] [device counts, e.g. 1,2,8]
  cold      first call of the process: result block fresh from the library pool (huge pages, staged through the ring)
  second    the block comes back from the pool: touched pages, still staged
  third     the block is page-locked on its second reuse (registration inside the timed call)
  warm      pool block already page-locked: direct DMA into the array the caller receives
  np.empty  a caller-provided fresh pageable array per call
One JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import _lib, engine, sketchlib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
L = _lib.load()
counts = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, L.ppb_device_count()]
kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device="cuda:0", n_roots=2).cpu().numpy().view(np.uint64)
torch.cuda.empty_cache()
table = synth.random_match_table(kmers, 3)
cl = synth.synth_clusters(n, 3)
rows = engine.num_rows(n)


def run(tag, g, out=None):
    os.environ["PPB_DEVICES"] = str(g)
    t0 = time.perf_counter()
    res, ndeg = sketchlib.query_arrays(sk, None, kmers, table, cl, None, out=out)
    dt = time.perf_counter() - t0
    print(json.dumps({"what": tag, "n_dev": g, "N": n, "rows": rows, "ms": round(dt * 1e3, 1),
                      "Gpairs_per_s": round(rows / dt / 1e9, 3), "GBps_to_host": round(rows * 8 / dt / 1e9, 1),
                      "n_degenerate": ndeg, "checksum": float(res[:1_000_000].sum())}), flush=True)
    return res


for g in sorted(set(c for c in counts if 1 <= c <= L.ppb_device_count())):
    L.ppb_release_workspace()
    r = run("cold (first call: fresh pool block, staged)", g)
    del r
    r = run("second call (pool block reused: touched pages, staged)", g)
    del r
    r = run("third call (pool block page-locked inside this call)", g)
    del r
    for _ in range(3):
        r = run("warm (page-locked pool block, direct DMA)", g)
        del r
    L.ppb_release_workspace()
    for _ in range(2):
        out = np.empty((rows, 2), dtype=np.float32)
        run("np.empty (fresh pageable array from the caller)", g, out=out)
        del out
