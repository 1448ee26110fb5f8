#!/usr/bin/env python
"""Integer-pipe micro-roofline and co-issue probes (ppb_microbench_dev): lane-ops/s per mode."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import _lib  # noqa: E402

L = _lib.load()
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
names = {0: "LOP3 only (8 chains x 14)", 1: "POPC only", 2: "14 LOP3 : 1 POPC : 1 IADD", 3: "REDUX",
         4: "56 LOP3 + 16 IMAD", 5: "56 LOP3 + 16 LDS", 6: "56 LOP3 + 16 FFMA"}
for mode in list(range(7)) + [10, 12, 14, 15]:
    ops = C.c_int64(0)
    iters = 20000 if mode % 10 != 3 else 4000
    best = 0.0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.ppb_microbench_dev(mode, iters, sink.data_ptr(), C.byref(ops), torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
    print(f"mode {mode} {names[mode % 10] + (' @ 2 warps/scheduler' if mode >= 10 else ''):52s} {best / 1e12:7.3f} T lane-ops/s (of the counted op)")
