#!/usr/bin/env python
"""One call each of the N1 consumers at 400 M rows, for an `ncu --metrics gpu__time_duration.sum` launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n1_launches.csv python tools/n1_breakdown.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import _lib
L = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 28_285
rows = n_samples * (n_samples - 1) // 2
g = torch.Generator(device=dev); g.manual_seed(1)
d = torch.rand((rows, 2), device=dev, generator=g) * 0.5
oi, oj, oo = (torch.empty(rows, dtype=torch.int64, device=dev) for _ in range(3))
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
scratch = torch.empty(L.ppb_edges_scratch_bytes(rows), dtype=torch.uint8, device=dev)
for rep in range(2):
    _lib.check(L.ppb_edges_from_dists_dev(d.data_ptr(), rows, n_samples, 2, C.c_float(0.05), C.c_float(0.08), oi.data_ptr(), oj.data_ptr(), rows, cnt.data_ptr(), scratch.data_ptr(), st))
    xm = np.linspace(0.01, 0.3, 30).astype(np.float32)
    _lib.check(L.ppb_threshold_iterate_2d_dev(d.data_ptr(), rows, xm.ctypes.data, 30, C.c_float(0.3), oi.data_ptr(), oj.data_ptr(), oo.data_ptr(), rows, cnt.data_ptr(), st))
    offs = np.linspace(0.01, 0.3, 30)
    _lib.check(L.ppb_threshold_iterate_1d_dev(d.data_ptr(), rows, offs.ctypes.data, 30, 2, C.c_float(0.0), C.c_float(0.0), C.c_float(0.3), C.c_float(0.3), oi.data_ptr(), oj.data_ptr(), oo.data_ptr(), rows, cnt.data_ptr(), st))
    torch.cuda.synchronize()
print("done", int(cnt.item()))
