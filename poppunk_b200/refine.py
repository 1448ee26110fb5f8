"""GPU twin of the one ``poppunk_refine`` function on the hot path: ``assignThreshold``.

Reference: src/boundary.cpp:42-80 (line_dist, assign_threshold), bound at src/python_bindings.cpp:18-25,
79-83 with ``distMat`` as a ``.noconvert()`` Eigen ref — i.e. the array must already be float32 and
C-contiguous; called from PopPUNK/models.py:1085-1089 as ``assignThreshold(X/self.scale, slope, x_max, y_max)``.
The fused form (labels straight from the distance kernel, no (n,2) round trip) is
``poppunk_b200.engine.query(..., boundary=(slope, x_max, y_max, scale_x, scale_y))``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def assignThreshold(distMat, slope, x_max, y_max, num_threads=1, device_id=0):
    """Assign samples based on their relation to a 2D boundary: float32 ``[n]`` in {-1, 0, +1}.

    Same argument order as ``poppunk_refine.assignThreshold``; ``num_threads`` is accepted and unused."""
    del num_threads
    L = _lib.load()
    if not isinstance(distMat, np.ndarray) or distMat.dtype != np.float32 or distMat.ndim != 2 \
            or distMat.shape[1] != 2 or not distMat.flags.c_contiguous:
        # pybind11 .noconvert() raises TypeError for anything that is not float32 C-contiguous (n, 2)
        raise TypeError("assignThreshold(): incompatible function arguments: distMat must be a C-contiguous "
                        "float32 array of shape (n, 2)")
    if slope not in (0, 1, 2):
        return np.zeros(distMat.shape[0], dtype=np.float32)  # boundary.cpp:44-56 leaves boundary_side = 0
    out = np.empty(distMat.shape[0], dtype=np.float32)
    if L.ppb_device_count() <= 0:
        raise RuntimeError("poppunk_b200: no CUDA device visible — this engine has no CPU fallback")
    # python_bindings.cpp:19-23 narrows the double arguments to float
    check(L.ppb_assign_threshold_host(distMat.ctypes.data, distMat.shape[0], int(slope),
                                      C.c_float(x_max), C.c_float(y_max), out.ctypes.data, device_id),
          "ppb_assign_threshold_host")
    return out


# --------------------------------------------------------------------------------------------
# N1 ("next" row of SURVEY.md section 8f): edge lists after the threshold, on the GPU
# --------------------------------------------------------------------------------------------
def _edges_to_tuples(i, j):
    return list(zip(i.tolist(), j.tolist()))


def _run_edges(call, n_rows, device_id):
    """Shared driver: allocate (i, j) for the worst case, run the ordered compaction, trim to the count."""
    import torch
    from . import engine
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    with torch.cuda.device(dev):
        oi = torch.empty(max(n_rows, 1), dtype=torch.int64, device=dev)
        oj = torch.empty(max(n_rows, 1), dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        scratch = torch.empty(L.ppb_edges_scratch_bytes(n_rows), dtype=torch.uint8, device=dev)
        call(L, dev, oi, oj, cnt, scratch, engine._stream_ptr(dev))
        n = int(cnt.item())
    return oi[:n].cpu().numpy(), oj[:n].cpu().numpy()


def edgeThreshold(distMat, slope, x_max, y_max, device_id=0):
    """``poppunk_refine.edgeThreshold`` (src/boundary.cpp:82-95 edge_iterate, bound at
    src/python_bindings.cpp:27-32): list of (i, j) tuples of the rows with ``line_dist <= 0``, in row order."""
    import torch
    if not isinstance(distMat, np.ndarray) or distMat.dtype != np.float32 or distMat.ndim != 2 \
            or distMat.shape[1] != 2 or not distMat.flags.c_contiguous:
        raise TypeError("edgeThreshold(): distMat must be a C-contiguous float32 array of shape (n, 2)")
    n_rows = distMat.shape[0]
    n_samples = int(0.5 * (1 + np.sqrt(1 + 8 * n_rows)))   # rows_to_samples, boundary.cpp:18-20

    def call(L, dev, oi, oj, cnt, scratch, stream):
        d = torch.from_numpy(distMat).to(dev)
        check(L.ppb_edges_from_dists_dev(d.data_ptr(), n_rows, n_samples, int(slope), C.c_float(x_max),
                                         C.c_float(y_max), oi.data_ptr(), oj.data_ptr(), oi.numel(), cnt.data_ptr(),
                                         scratch.data_ptr(), stream), "ppb_edges_from_dists_dev")
    return _edges_to_tuples(*_run_edges(call, n_rows, device_id))


def generateTuples(assignments, within_label, self=True, num_ref=0, int_offset=0, device_id=0):
    """``poppunk_refine.generateTuples`` (src/boundary.cpp:97-123, bound at src/python_bindings.cpp:34-40)."""
    import torch
    a = np.ascontiguousarray(assignments)
    if a.dtype == np.int8:
        code = 0
    elif a.dtype == np.float32:
        code = 2
    else:
        a, code = a.astype(np.int32), 1   # the binding takes std::vector<int>
    n_rows = a.shape[0]
    n_map = int(0.5 * (1 + np.sqrt(1 + 8 * n_rows))) if self else int(num_ref)

    def call(L, dev, oi, oj, cnt, scratch, stream):
        t = torch.from_numpy(a).to(dev)
        check(L.ppb_edges_from_labels_dev(t.data_ptr(), code, n_rows, int(within_label), int(bool(self)), max(n_map, 1),
                                          int(int_offset), oi.data_ptr(), oj.data_ptr(), oi.numel(), cnt.data_ptr(),
                                          scratch.data_ptr(), stream), "ppb_edges_from_labels_dev")
    return _edges_to_tuples(*_run_edges(call, n_rows, device_id))
