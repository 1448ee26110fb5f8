"""GPU twin of the reference's ``poppunk_refine`` extension module (src/python_bindings.cpp:79-136): the consumers of
the (n_pairs, 2) distance array.  Same function names, argument order and return objects, so
``import poppunk_b200.refine as poppunk_refine`` is the swap.

    assignThreshold                      src/boundary.cpp:42-80     (hot path; fused form: engine.query(..., boundary=))
    edgeThreshold / generateTuples /
    generateAllTuples                    src/boundary.cpp:82-149
    thresholdIterate1D / 2D              src/boundary.cpp:151-237
    get_kNN_distances / lowerRank /
    extend                               src/extend.cpp:52-289

``distMat`` arguments must already be float32 and C-contiguous, as the reference's ``.noconvert()`` bindings demand.
Every function runs on the GPU through the C ABI (include/ppb.h); there is no CPU path.  Results are bit-identical to
the reference's own sources compiled into oracle/_ref (tests/golden/refine_ref.npz).
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


def generateAllTuples(num_ref, num_queries=0, self=True, int_offset=0, device_id=0):
    """``poppunk_refine.generateAllTuples`` (src/boundary.cpp:125-149, bound at src/python_bindings.cpp:42-48)."""
    import torch
    from . import engine
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    total = num_ref * (num_ref - 1) // 2 if self else num_ref * num_queries
    with torch.cuda.device(dev):
        oi = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        oj = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        check(L.ppb_generate_all_tuples_dev(int(num_ref), int(num_queries), int(bool(self)), int(int_offset),
                                            oi.data_ptr(), oj.data_ptr(), engine._stream_ptr(dev)),
              "ppb_generate_all_tuples_dev")
    return _edges_to_tuples(oi[:total].cpu().numpy(), oj[:total].cpu().numpy())


def _check_distmat(distMat, what):
    if not isinstance(distMat, np.ndarray) or distMat.dtype != np.float32 or distMat.ndim != 2 \
            or distMat.shape[1] != 2 or not distMat.flags.c_contiguous:
        raise TypeError(f"{what}(): distMat must be a C-contiguous float32 array of shape (n, 2)")


def _run_iterate(call, n_rows, device_id):
    """Allocate (i, j, offset_idx) for n_rows admissions, run, and re-run once with the exact size in the
    (pathological) case that more were counted than fit."""
    import torch
    from . import engine
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    cap = max(n_rows, 1)
    with torch.cuda.device(dev):
        while True:
            oi, oj, oo = (torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(3))
            cnt = torch.zeros(1, dtype=torch.int64, device=dev)
            call(L, dev, oi, oj, oo, cnt, engine._stream_ptr(dev))
            n = int(cnt.item())
            if n <= cap:
                break
            cap = n
    return oi[:n].cpu().tolist(), oj[:n].cpu().tolist(), oo[:n].cpu().tolist()


def thresholdIterate1D(distMat, offsets, slope, x0, y0, x1, y1, num_threads=1, device_id=0):
    """``poppunk_refine.thresholdIterate1D`` (src/boundary.cpp:151-209, bound at src/python_bindings.cpp:50-62):
    move the boundary along the line (x0,y0)->(x1,y1) by the sorted ``offsets``; returns the lists
    ``(i_vec, j_vec, offset_idx)`` of the edges in the order the boundary admits them."""
    import torch
    del num_threads
    _check_distmat(distMat, "thresholdIterate1D")
    off = np.ascontiguousarray(offsets, dtype=np.float64)
    if (np.diff(off) < 0).any():
        raise RuntimeError("Offsets to thresholdIterate1D must be sorted")   # python_bindings.cpp:56-58
    n_rows = distMat.shape[0]

    def call(L, dev, oi, oj, oo, cnt, stream):
        d = torch.from_numpy(distMat).to(dev)
        # python_bindings.cpp:52-55 takes x0..y1 as double and narrows them to float at the call
        check(L.ppb_threshold_iterate_1d_dev(d.data_ptr(), n_rows, off.ctypes.data, off.shape[0], int(slope),
                                             C.c_float(x0), C.c_float(y0), C.c_float(x1), C.c_float(y1),
                                             oi.data_ptr(), oj.data_ptr(), oo.data_ptr(), oi.numel(), cnt.data_ptr(),
                                             stream), "ppb_threshold_iterate_1d_dev")
    return _run_iterate(call, n_rows, device_id)


def thresholdIterate2D(distMat, x_max, y_max, device_id=0):
    """``poppunk_refine.thresholdIterate2D`` (src/boundary.cpp:211-237, bound at src/python_bindings.cpp:64-77)."""
    import torch
    _check_distmat(distMat, "thresholdIterate2D")
    xm = np.ascontiguousarray(x_max, dtype=np.float32)
    if (np.diff(xm) < 0).any():
        raise RuntimeError("x_max range to thresholdIterate2D must be sorted")   # python_bindings.cpp:70-73
    n_rows = distMat.shape[0]
    out = ([], [], [])
    # the kernel takes up to 1024 steps per launch; step o needs boundary o-1, so slices overlap by one step
    for s0 in range(0, max(len(xm), 1), 1023):
        part = np.ascontiguousarray(xm[max(s0 - 1, 0):s0 + 1023])
        skip = 1 if s0 > 0 else 0

        def call(L, dev, oi, oj, oo, cnt, stream, part=part):
            d = torch.from_numpy(distMat).to(dev)
            check(L.ppb_threshold_iterate_2d_dev(d.data_ptr(), n_rows, part.ctypes.data, len(part), C.c_float(y_max),
                                                 oi.data_ptr(), oj.data_ptr(), oo.data_ptr(), oi.numel(),
                                                 cnt.data_ptr(), stream), "ppb_threshold_iterate_2d_dev")
        i, j, o = _run_iterate(call, n_rows, device_id)
        for a, b, c in zip(i, j, o):
            if c >= skip:
                out[0].append(a)
                out[1].append(b)
                out[2].append(c - skip + s0)
    return out


# --------------------------------------------------------------------------------------------
# N3: nearest neighbours for the lineage models (src/extend.cpp)
# --------------------------------------------------------------------------------------------
def _coo_to_device(rr_mat, dev):
    import torch
    i, j, d = rr_mat
    return (torch.as_tensor(np.ascontiguousarray(i, dtype=np.int64)).to(dev),
            torch.as_tensor(np.ascontiguousarray(j, dtype=np.int64)).to(dev),
            torch.as_tensor(np.ascontiguousarray(d, dtype=np.float32)).to(dev))


KNN_MAX = 2048            # kKnnMax in csrc/ppb_refine.cuh: the survivors of a row are sorted in shared memory
LOWER_RANK_MAX_ROW = 1024  # kLowerMaxRow: a sample's sparse row is sorted by one warp in shared memory


def _check_knn(kNN, who):
    """Limits the reference does not have ("Limits"): fail with a clear message, never a wrong result."""
    if kNN < 0 or kNN > KNN_MAX:
        raise ValueError(f"{who}(): kNN = {kNN} is outside this engine's range [0, {KNN_MAX}] "
                         "(PopPUNK's lineage ranks are 1..~100; use the reference's CPU poppunk_refine beyond that)")


def get_kNN_distances(distMat, kNN, dist_col=0, num_threads=1, device_id=0):
    """``poppunk_refine.get_kNN_distances`` (src/extend.cpp:245-289, bound at src/python_bindings.cpp:131-136):
    ``(i_vec, j_vec, dists)`` lists, ``rows * kNN`` long, of each row's nearest columns (ties: lower column first;
    never the row's own index).  ``distMat`` may also be a CUDA tensor (e.g. from ``reshape.longToSquare``)."""
    import torch
    from . import engine
    del dist_col, num_threads     # dist_col is unused by the reference as well
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    if isinstance(distMat, np.ndarray):
        if distMat.dtype != np.float32 or distMat.ndim != 2 or not distMat.flags.c_contiguous:
            raise TypeError("get_kNN_distances(): distMat must be a C-contiguous float32 matrix")  # .noconvert()
        m = torch.from_numpy(distMat).to(dev)
    else:
        m = distMat.to(dev).contiguous()
    rows, cols = m.shape
    kNN = int(kNN)
    if kNN == 0:
        return [], [], []                                       # the reference's loops produce nothing
    _check_knn(kNN, "get_kNN_distances")
    with torch.cuda.device(dev):
        oi = torch.empty(rows * kNN, dtype=torch.int64, device=dev)
        oj = torch.empty(rows * kNN, dtype=torch.int64, device=dev)
        od = torch.empty(rows * kNN, dtype=torch.float32, device=dev)
        check(L.ppb_knn_dev(m.data_ptr(), rows, cols, kNN, oi.data_ptr(), oj.data_ptr(), od.data_ptr(),
                            engine._stream_ptr(dev)), "ppb_knn_dev")
    return oi.cpu().tolist(), oj.cpu().tolist(), od.cpu().tolist()


def lowerRank(rr_mat, n_samples, kNN, reciprocal_only=False, count_unique_distances=False, lineage_resolution=0.0,
              num_threads=1, device_id=0):
    """``poppunk_refine.lowerRank`` (src/extend.cpp:146-243, bound at src/python_bindings.cpp:122-129)."""
    import torch
    from . import engine
    del num_threads
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    with torch.cuda.device(dev):
        si, sj, sd = _coo_to_device(rr_mat, dev)
        nnz = si.numel()
        oi = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
        oj = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
        od = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.ppb_lower_rank_dev(si.data_ptr(), sj.data_ptr(), sd.data_ptr(), nnz, int(n_samples), int(kNN),
                                   int(bool(reciprocal_only)), int(bool(count_unique_distances)),
                                   C.c_float(lineage_resolution), oi.data_ptr(), oj.data_ptr(), od.data_ptr(),
                                   cnt.data_ptr(), engine._stream_ptr(dev)), "ppb_lower_rank_dev")
        n = int(cnt.item())
    return oi[:n].cpu().tolist(), oj[:n].cpu().tolist(), od[:n].cpu().tolist()


def extend(rr_mat, qq_mat, qr_mat, kNN, num_threads=1, device_id=0):
    """``poppunk_refine.extend`` (src/extend.cpp:52-136, bound at src/python_bindings.cpp:115-120): ``qr_mat`` is
    (n_ref, n_query), ``qq_mat`` (n_query, n_query), both float32 C-contiguous."""
    import torch
    from . import engine
    del num_threads
    for name, a in (("qq_mat", qq_mat), ("qr_mat", qr_mat)):
        if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != 2 or not a.flags.c_contiguous:
            raise TypeError(f"extend(): {name} must be a C-contiguous float32 matrix")
    dev = engine._require_cuda(f"cuda:{device_id}")
    L = _lib.load()
    nr, nq = qr_mat.shape
    kNN = int(kNN)
    if kNN == 0:
        return [], [], []
    _check_knn(kNN, "extend")
    with torch.cuda.device(dev):
        si, sj, sd = _coo_to_device(rr_mat, dev)
        qq = torch.from_numpy(qq_mat).to(dev)
        qr = torch.from_numpy(qr_mat).to(dev)
        cap = max((nr + nq) * kNN, 1)
        oi = torch.empty(cap, dtype=torch.int64, device=dev)
        oj = torch.empty(cap, dtype=torch.int64, device=dev)
        od = torch.empty(cap, dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.ppb_extend_dev(si.data_ptr(), sj.data_ptr(), sd.data_ptr(), si.numel(), qq.data_ptr(), qr.data_ptr(),
                               nr, nq, kNN, oi.data_ptr(), oj.data_ptr(), od.data_ptr(), cnt.data_ptr(),
                               engine._stream_ptr(dev)), "ppb_extend_dev")
        n = int(cnt.item())
    return oi[:n].cpu().tolist(), oj[:n].cpu().tolist(), od[:n].cpu().tolist()
