#!/usr/bin/env python
"""Where does a device-resident bench step spend time beyond the kernel?  Times step variants (CUDA events, 4 steps each)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device="cuda", n_roots=2)
table, cl = synth.random_match_table(kmers, 3), synth.synth_clusters(n, 3)
table_dev, cl_dev = torch.as_tensor(table).cuda(), engine.DeviceClusters.upload(cl, "cuda:0")
out = torch.empty((engine.num_rows(n), 2), dtype=torch.float32, device="cuda")
ndeg = torch.zeros(1, dtype=torch.int64, device="cuda")
packed_np, packed_dev, packed_plain = engine.pack(sk, clusters=cl), engine.pack(sk, clusters=cl_dev), engine.pack(sk)
variants = {
    "query only, device table": lambda: engine.query(packed_dev, None, kmers, rand_table=table_dev, out=out, n_degenerate=ndeg),
    "query only, numpy table": lambda: engine.query(packed_np, None, kmers, rand_table=table, out=out, n_degenerate=ndeg),
    "pack(device clusters) + query(device table)": lambda: engine.query(engine.pack(sk, clusters=cl_dev), None, kmers, rand_table=table_dev, out=out, n_degenerate=ndeg),
    "pack(numpy clusters) + query(device table)": lambda: engine.query(engine.pack(sk, clusters=cl), None, kmers, rand_table=table_dev, out=out, n_degenerate=ndeg),
    "pack(numpy clusters) + query(numpy table)": lambda: engine.query(engine.pack(sk, clusters=cl), None, kmers, rand_table=table, out=out, n_degenerate=ndeg),
    "pack + query, no table": lambda: engine.query(engine.pack(sk), None, kmers, out=out, n_degenerate=ndeg),
}
for name, fn in variants.items():
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(4): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1) / 4:9.2f} ms/step (events)  {(time.perf_counter() - t0) * 250:9.2f} ms/step (wall)", flush=True)
