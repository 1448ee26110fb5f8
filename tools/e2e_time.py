#!/usr/bin/env python
"""End-to-end time of the host-buffer call (ppb_query_host, pinned buffers) for several launch plans:
    python tools/e2e_time.py [N] [chunk_rows,ring ...]      e.g.  100000 67108864,8 134217728,2
PPB_HOST_TRACE=1 adds the per-chunk kernel/copy timeline on stderr."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
plans = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(1 << 26, 8)]
kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device="cuda")
sk_host = torch.empty(sk.shape, dtype=torch.int64, pin_memory=True)
sk_host.copy_(sk)
del sk
rows = engine.num_rows(n)
t0 = time.perf_counter()
out_host = torch.empty((rows, 2), dtype=torch.float32, pin_memory=True)
print(f"pinned {rows * 8 / 1e9:.1f} GB result buffer in {time.perf_counter() - t0:.1f} s", flush=True)
ref_np, out_np = sk_host.numpy().view(np.uint64), out_host.numpy()
# plain D2H rate of this box, for reference
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print(f"plain pinned D2H: {4 * (1 << 30) / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
del d, h
for chunk, ring in plans:
    os.environ["PPB_HOST_CHUNK_ROWS"] = str(chunk)
    os.environ["PPB_HOST_RING"] = str(ring)
    trace = os.environ.pop("PPB_HOST_TRACE", None)
    engine.query_host(ref_np, None, kmers, out=out_np)          # warm-up: workspace, tile lists
    ts = []
    for rep in range(3):
        if trace and rep == 2:
            os.environ["PPB_HOST_TRACE"] = trace
        t0 = time.perf_counter()
        engine.query_host(ref_np, None, kmers, out=out_np)
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    print(f"chunk_rows={chunk} ring={ring}: {ms:8.1f} ms  {rows / ms / 1e6:.3f} Gpairs/s  (runs: "
          + ", ".join(f"{t * 1e3:.0f}" for t in ts) + f")  checksum {float(out_np[:1000000].sum()):.6f}", flush=True)
