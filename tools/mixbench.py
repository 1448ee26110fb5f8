#!/usr/bin/env python
"""How much of the LOP3 pipe the distance kernel's INSTRUCTION MIX can reach, by warps per scheduler (no barriers, no TMA,
no epilogue warps): python tools/mixbench.py   -> one line per (rows per warp, warps per scheduler, with/without LDS)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poppunk_b200 import _lib  # noqa: E402

L = _lib.load()
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


ops = C.c_int64(0)
peak_t = timed(lambda: _lib.check(L.ppb_microbench_dev(0, 20000, sink.data_ptr(), C.byref(ops), st)))
peak = ops.value / peak_t
print(json.dumps({"what": "LOP3-only peak", "lop3_per_s": peak}), flush=True)
for rows, wmax in ((8, 2), (5, 3), (4, 4)):
    for w in range(1, wmax + 1):
        for lds in (0, 1):
            t = timed(lambda: _lib.check(L.ppb_microbench_mix_dev(rows, w, lds, 4000, sink.data_ptr(), C.byref(ops), st)))
            print(json.dumps({"rows_per_warp": rows, "warps_per_scheduler": w, "with_lds": lds, "lop3_per_s": ops.value / t,
                              "frac_of_lop3_peak": ops.value / t / peak}), flush=True)
