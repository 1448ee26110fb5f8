#!/usr/bin/env python
"""Time the distance kernel alone (CUDA events) — used to compare build variants: PPB_LIB=<.so> python tools/kernel_time.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000
with_table = len(sys.argv) > 2 and sys.argv[2] == "rand"      # random-match correction on (3 clusters), as in production calls
kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device="cuda")
table = synth.random_match_table(kmers, 3) if with_table else None
packed = engine.pack(sk, clusters=synth.synth_clusters(n, 3) if with_table else None)
out = torch.empty((engine.num_rows(n), 2), dtype=torch.float32, device="cuda")
for _ in range(2):
    engine.query(packed, None, kmers, rand_table=table, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.query(packed, None, kmers, rand_table=table, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(f"{os.environ.get('PPB_LIB', 'default') + (' +random-match table' if with_table else ''):40s} N={n} median {ms:9.3f} ms  {out.shape[0] / ms / 1e6:8.3f} Gpairs/s  "
      f"LOP3 frac of 18.5T: {out.shape[0] * 2240 / (ms * 1e-3) / 18.5e12:.3f}  checksum {float(out[:1000000].sum()):.6f}")
