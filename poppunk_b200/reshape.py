"""N2 ("next" row of SURVEY.md section 8f): long <-> square reshapes of the path's output, on the GPU.

Same names and keyword arguments as the pp_sketchlib functions PopPUNK calls (PopPUNK/utils.py:393-405,
network.py:2133-2134, models.py:1217,1357, mandrake.py:165): ``distVec`` is an (n_pairs, 1) float32 column
(PopPUNK passes ``distMat[:, [c]]``), squares are (n, n) float32, symmetric, zero diagonal.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check


def _dev(device_id):
    import torch
    from . import engine
    return engine._require_cuda(f"cuda:{device_id}")


def _col(a):
    a = np.asarray(a, dtype=np.float32)
    return np.ascontiguousarray(a.reshape(-1))


def longToSquare(distVec, num_threads=1, device_id=0):
    import torch
    from . import engine
    del num_threads
    v = _col(distVec)
    n = int(0.5 * (1 + np.sqrt(1 + 8 * v.shape[0])))
    if n * (n - 1) // 2 != v.shape[0]:
        raise RuntimeError("distVec length is not n(n-1)/2")
    dev, L = _dev(device_id), _lib.load()
    with torch.cuda.device(dev):
        d = torch.from_numpy(v).to(dev)
        sq = torch.empty((n, n), dtype=torch.float32, device=dev)
        check(L.ppb_long_to_square_dev(d.data_ptr(), 1, n, sq.data_ptr(), engine._stream_ptr(dev)), "ppb_long_to_square_dev")
        return sq.cpu().numpy()


def squareToLong(distMat, num_threads=1, device_id=0):
    import torch
    from . import engine
    del num_threads
    m = np.ascontiguousarray(distMat, dtype=np.float32)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise RuntimeError("distMat must be square")
    n = m.shape[0]
    dev, L = _dev(device_id), _lib.load()
    with torch.cuda.device(dev):
        d = torch.from_numpy(m).to(dev)
        v = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=dev)
        check(L.ppb_square_to_long_dev(d.data_ptr(), n, v.data_ptr(), engine._stream_ptr(dev)), "ppb_square_to_long_dev")
        return v.cpu().numpy()


def longToSquareMulti(distVec, query_ref_distVec, query_query_distVec, num_threads=1, device_id=0):
    import torch
    from . import engine
    del num_threads
    rr, qr, qq = _col(distVec), _col(query_ref_distVec), _col(query_query_distVec)
    R = int(0.5 * (1 + np.sqrt(1 + 8 * rr.shape[0])))
    Q = int(0.5 * (1 + np.sqrt(1 + 8 * qq.shape[0])))
    if R * (R - 1) // 2 != rr.shape[0] or Q * (Q - 1) // 2 != qq.shape[0] or qr.shape[0] != R * Q:
        raise RuntimeError("inconsistent vector lengths for longToSquareMulti")
    dev, L = _dev(device_id), _lib.load()
    with torch.cuda.device(dev):
        t = [torch.from_numpy(x).to(dev) for x in (rr, qr, qq)]
        sq = torch.empty((R + Q, R + Q), dtype=torch.float32, device=dev)
        check(L.ppb_long_to_square_multi_dev(t[0].data_ptr(), 1, t[1].data_ptr(), 1, t[2].data_ptr(), 1, R, Q,
                                             sq.data_ptr(), engine._stream_ptr(dev)), "ppb_long_to_square_multi_dev")
        return sq.cpu().numpy()
