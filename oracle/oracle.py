"""CPU ORACLE — Python face.  TEST INFRASTRUCTURE ONLY (see the header of ppb_oracle.c).

Two independent restatements of the reference algorithm for the distance hot path:

* :func:`query` & co — ctypes bindings of ``libppo.so`` (``ppb_oracle.c``, C + OpenMP), the checker
  used at sizes up to ~10^8 pairs and the CPU baseline timed by ``bench.py``;
* :func:`counts_numpy`, :func:`regress_numpy`, :func:`assign_threshold_numpy` — NumPy restatements
  written a *different way* (un-sliced signature equality instead of bit-sliced popcounts;
  ``numpy.linalg.lstsq`` instead of closed-form OLS) that pin the C oracle in ``tests/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
PARITY STATUS: parity unpinned for (pi, a) values — pp-sketchlib is absent (details in ppb_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DISTS, OUT_JACCARD, OUT_COUNTS = 0, 1, 2
BBITS = 14


class Boundary(C.Structure):
    _fields_ = [("slope", C.c_int32), ("x_max", C.c_float), ("y_max", C.c_float),
                ("scale_x", C.c_float), ("scale_y", C.c_float)]


def _cpu_key() -> str:
    """What -march=native resolved to on the machine that built libppo_native.so: the CPU's feature flags."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(" ".join(sorted(flags.split(":")[-1].split())).encode()).hexdigest()[:12]


def build(native: bool = False) -> str:
    """Compile the C oracle (``make -C oracle``); returns the path of the shared object.  The -march=native build is
    tied to the CPU it was made on (it travels to other boxes with the tree): a sidecar file records the CPU's feature
    flags and a different CPU forces a rebuild instead of risking an illegal instruction."""
    if not native:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
        return os.path.join(_HERE, "libppo.so")
    side = os.path.join(_HERE, "libppo_native.cpu")
    key = _cpu_key()
    try:
        stale = open(side).read().strip() != key
    except OSError:
        stale = True
    subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if stale else []) + ["native"], check=True)
    with open(side, "w") as f:
        f.write(key + "\n")
    return os.path.join(_HERE, "libppo_native.so")


_lib = None
_lib_native = None


def lib(native: bool = False):
    global _lib, _lib_native
    if _lib is not None and not native:
        return _lib
    if _lib_native is not None and native:
        return _lib_native
    path = os.path.join(_HERE, "libppo_native.so" if native else "libppo.so")
    if native or not os.path.exists(path):
        try:
            build(native)       # native: re-checks the CPU the file was built on (cheap when up to date)
        except Exception:
            if not os.path.exists(path):
                raise
    L = C.CDLL(path)
    i64, i32, vp = C.c_int64, C.c_int32, C.c_void_p
    L.ppo_query_host.argtypes = [vp, i64, vp, i64, vp, i32, i32, i32, vp, i32, vp, vp, i64, i64,
                                 i32, vp, vp, vp, vp, i32]
    L.ppo_query_host.restype = C.c_int
    L.ppo_assign_threshold.argtypes = [vp, i64, i32, C.c_float, C.c_float, vp, i32]
    L.ppo_assign_threshold.restype = C.c_int
    for name in ("ppo_square_to_condensed",):
        getattr(L, name).argtypes = [i64, i64, i64]
        getattr(L, name).restype = i64
    L.ppo_calc_row_idx.argtypes = [i64, i64]
    L.ppo_calc_row_idx.restype = i64
    L.ppo_calc_col_idx.argtypes = [i64, i64, i64]
    L.ppo_calc_col_idx.restype = i64
    L.ppo_num_rows.argtypes = [i64, i64, C.c_int]
    L.ppo_num_rows.restype = i64
    L.ppo_max_threads.restype = C.c_int
    L.ppo_edge_iterate.argtypes = [vp, i64, i32, C.c_float, C.c_float, vp, vp]
    L.ppo_edge_iterate.restype = i64
    L.ppo_generate_tuples.argtypes = [vp, i64, i32, i32, i64, i64, vp, vp]
    L.ppo_generate_tuples.restype = i64
    L.ppo_long_to_square.argtypes = [vp, i64, i64, vp]
    L.ppo_square_to_long.argtypes = [vp, i64, vp]
    L.ppo_long_to_square_multi.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, vp]
    L.ppo_regress_rows.argtypes = [vp, i64, vp, i32, C.c_double, vp]
    L.ppo_regress_rows.restype = i64
    f32 = C.c_float
    L.ppo_generate_all_tuples.argtypes = [i64, i64, i32, i64, vp, vp]
    L.ppo_generate_all_tuples.restype = i64
    L.ppo_iterate_1d_boundary.argtypes = [C.c_double, i32, f32, f32, f32, f32, vp, vp]
    L.ppo_iterate_1d_boundary.restype = None
    L.ppo_threshold_iterate_1d.argtypes = [vp, i64, vp, i64, i32, f32, f32, f32, f32, vp, vp, vp]
    L.ppo_threshold_iterate_1d.restype = i64
    L.ppo_threshold_iterate_2d.argtypes = [vp, i64, vp, i64, f32, vp, vp, vp]
    L.ppo_threshold_iterate_2d.restype = i64
    L.ppo_get_knn_distances.argtypes = [vp, i64, i64, i64, vp, vp, vp]
    L.ppo_get_knn_distances.restype = None
    L.ppo_lower_rank.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, f32, vp, vp, vp]
    L.ppo_lower_rank.restype = i64
    L.ppo_extend.argtypes = [vp, vp, vp, i64, vp, vp, i64, i64, i64, vp, vp, vp]
    L.ppo_extend.restype = i64
    L.ppo_tuned_available.restype = C.c_int
    L.ppo_query_host_tuned.argtypes = [vp, i64, vp, i32, i32, i32, vp, i32, vp, i64, i64, vp, vp, i32]
    L.ppo_query_host_tuned.restype = C.c_int
    if native:
        _lib_native = L
    else:
        _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(lib().ppo_max_threads())


def num_rows(n_ref, n_qry=None):
    return n_ref * (n_ref - 1) // 2 if n_qry is None else n_ref * n_qry


def query(ref, qry, kmers, rand_table=None, ref_cluster=None, qry_cluster=None,
          row_begin=0, row_end=None, out_mode=OUT_DISTS, boundary=None, threads=None,
          native=False):
    """Oracle twin of ``ppb_query_host``.  ``ref``/``qry``: uint64 ``[n][K][W]``; qry=None => self.

    Returns ``(out, n_degenerate)`` or ``(out, labels, n_degenerate)`` when ``boundary`` is given
    (tuple ``(slope, x_max, y_max, scale_x, scale_y)``).
    """
    L = lib(native)
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    n_ref, K, W = ref.shape
    if W % BBITS:
        raise ValueError("W must be sketchsize64*14")
    ss64 = W // BBITS
    n_qry = 0
    if qry is not None:
        qry = np.ascontiguousarray(qry, dtype=np.uint64)
        n_qry = qry.shape[0]
        assert qry.shape[1:] == ref.shape[1:]
    kmers = np.ascontiguousarray(kmers, dtype=np.int32)
    assert kmers.shape == (K,)
    total = num_rows(n_ref, None if qry is None else n_qry)
    if row_end is None:
        row_end = total
    rows = row_end - row_begin
    nclus = 0
    if rand_table is not None:
        rand_table = np.ascontiguousarray(rand_table, dtype=np.float32)
        nclus = rand_table.shape[0]
        assert rand_table.shape == (nclus, nclus, K)
        ref_cluster = np.ascontiguousarray(ref_cluster, dtype=np.uint16)
        if qry is not None:
            qry_cluster = np.ascontiguousarray(qry_cluster, dtype=np.uint16)
    if out_mode == OUT_DISTS:
        out = np.empty((rows, 2), dtype=np.float32)
    elif out_mode == OUT_JACCARD:
        out = np.empty((rows, K), dtype=np.float32)
    else:
        out = np.empty((rows, K), dtype=np.uint32)
    labels = None
    bnd = None
    if boundary is not None:
        bnd = Boundary(*boundary)
        labels = np.empty(rows, dtype=np.int8)
    ndeg = C.c_int64(0)
    if threads is None:
        threads = max_threads()
    rc = L.ppo_query_host(_ptr(ref), n_ref, _ptr(qry), n_qry, _ptr(kmers), K, ss64, BBITS,
                          _ptr(rand_table), nclus, _ptr(ref_cluster), _ptr(qry_cluster),
                          row_begin, row_end, out_mode, _ptr(out),
                          C.byref(bnd) if bnd is not None else None, _ptr(labels),
                          C.byref(ndeg), threads)
    if rc != 0:
        raise RuntimeError(f"ppo_query_host failed with code {rc}")
    if boundary is not None:
        return out, labels, int(ndeg.value)
    return out, int(ndeg.value)


def tuned_available() -> bool:
    """The AVX-512 arm (ppb_oracle_tuned.inc) exists only in the -march=native build, on CPUs with VPOPCNTDQ."""
    try:
        return bool(lib(native=True).ppo_tuned_available())
    except Exception:
        return False


def query_tuned(ref, kmers, rand_table=None, ref_cluster=None, row_begin=0, row_end=None, threads=None):
    """The tuned CPU arm of the self-mode (core, accessory) job: same arithmetic as :func:`query`, bit-identical
    results, AVX-512 + cache blocking + ln J table (oracle/ppb_oracle_tuned.inc).  Returns ``(out, n_degenerate)``."""
    L = lib(native=True)
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    n_ref, K, W = ref.shape
    if W % BBITS:
        raise ValueError("W must be sketchsize64*14")
    kmers = np.ascontiguousarray(kmers, dtype=np.int32)
    nclus = 0
    if rand_table is not None:
        rand_table = np.ascontiguousarray(rand_table, dtype=np.float32)
        nclus = rand_table.shape[0]
        ref_cluster = np.ascontiguousarray(ref_cluster, dtype=np.uint16)
    if row_end is None:
        row_end = num_rows(n_ref)
    out = np.empty((row_end - row_begin, 2), dtype=np.float32)
    ndeg = C.c_int64(0)
    rc = L.ppo_query_host_tuned(_ptr(ref), n_ref, _ptr(kmers), K, W // BBITS, BBITS, _ptr(rand_table), nclus,
                                _ptr(ref_cluster) if rand_table is not None else None, row_begin, row_end, _ptr(out),
                                C.byref(ndeg), threads or max_threads())
    if rc != 0:
        raise RuntimeError(f"ppo_query_host_tuned failed with code {rc}" + (" (no AVX-512 VPOPCNTDQ)" if rc == 2 else ""))
    return out, int(ndeg.value)


def assign_threshold(dists, slope, x_max, y_max, threads=1):
    """Oracle twin of ``poppunk_refine.assignThreshold`` (src/boundary.cpp:60-80)."""
    d = np.ascontiguousarray(dists, dtype=np.float32)
    out = np.empty(d.shape[0], dtype=np.float32)
    rc = lib().ppo_assign_threshold(_ptr(d), d.shape[0], slope, x_max, y_max, _ptr(out), threads)
    if rc != 0:
        raise RuntimeError("ppo_assign_threshold failed")
    return out


def regress_rows(jac, kmers, S):
    """(core, acc) float32 [rows][2] from per-k Jaccards float64 [rows][K] (C oracle regression only)."""
    jac = np.ascontiguousarray(jac, dtype=np.float64)
    kmers = np.ascontiguousarray(kmers, dtype=np.int32)
    out = np.empty((jac.shape[0], 2), dtype=np.float32)
    deg = lib().ppo_regress_rows(_ptr(jac), jac.shape[0], _ptr(kmers), jac.shape[1], float(S), _ptr(out))
    return out, int(deg)


def edge_iterate(dists, slope, x_max, y_max):
    """src/boundary.cpp:82-95: (i, j) int64 arrays of the rows with line_dist <= 0, in row order."""
    d = np.ascontiguousarray(dists, dtype=np.float32)
    oi, oj = np.empty(d.shape[0], dtype=np.int64), np.empty(d.shape[0], dtype=np.int64)
    n = lib().ppo_edge_iterate(_ptr(d), d.shape[0], slope, x_max, y_max, _ptr(oi), _ptr(oj))
    return oi[:n], oj[:n]


def generate_tuples(assignments, within_label, self=True, num_ref=0, int_offset=0):
    """src/boundary.cpp:97-123."""
    a = np.ascontiguousarray(assignments, dtype=np.int32)
    oi, oj = np.empty(a.shape[0], dtype=np.int64), np.empty(a.shape[0], dtype=np.int64)
    n = lib().ppo_generate_tuples(_ptr(a), a.shape[0], within_label, int(self), max(num_ref, 1), int_offset, _ptr(oi), _ptr(oj))
    return oi[:n], oj[:n]


def generate_all_tuples(num_ref, num_queries=0, self=True, int_offset=0):
    """src/boundary.cpp:125-149."""
    n = num_ref * (num_ref - 1) // 2 if self else num_ref * num_queries
    oi, oj = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    c = lib().ppo_generate_all_tuples(num_ref, num_queries, int(self), int_offset, _ptr(oi), _ptr(oj))
    return oi[:c], oj[:c]


def iterate_1d_boundary(offset, slope, x0, y0, x1, y1):
    """(x_max, y_max) of one step of threshold_iterate_1D (src/boundary.cpp:171-185)."""
    xm, ym = C.c_float(0), C.c_float(0)
    lib().ppo_iterate_1d_boundary(float(offset), slope, x0, y0, x1, y1, C.byref(xm), C.byref(ym))
    return xm.value, ym.value


def threshold_iterate_1d(dists, offsets, slope, x0, y0, x1, y1):
    """src/boundary.cpp:151-209: (i, j, offset_idx) int64 arrays."""
    d = np.ascontiguousarray(dists, dtype=np.float32)
    off = np.ascontiguousarray(offsets, dtype=np.float64)
    n = d.shape[0]
    oi, oj, oo = (np.empty(n, dtype=np.int64) for _ in range(3))
    c = lib().ppo_threshold_iterate_1d(_ptr(d), n, _ptr(off), off.shape[0], slope, x0, y0, x1, y1, _ptr(oi), _ptr(oj), _ptr(oo))
    return oi[:c], oj[:c], oo[:c]


def threshold_iterate_2d(dists, x_max, y_max):
    """src/boundary.cpp:211-237: (i, j, offset_idx) int64 arrays."""
    d = np.ascontiguousarray(dists, dtype=np.float32)
    xm = np.ascontiguousarray(x_max, dtype=np.float32)
    cap = d.shape[0] * max(1, xm.shape[0])
    oi, oj, oo = (np.empty(cap, dtype=np.int64) for _ in range(3))
    c = lib().ppo_threshold_iterate_2d(_ptr(d), d.shape[0], _ptr(xm), xm.shape[0], y_max, _ptr(oi), _ptr(oj), _ptr(oo))
    return oi[:c], oj[:c], oo[:c]


def get_knn_distances(mat, knn):
    """src/extend.cpp:245-289: (i, j, dist) of length rows*kNN."""
    m = np.ascontiguousarray(mat, dtype=np.float32)
    rows, cols = m.shape
    oi, oj = np.empty(rows * knn, dtype=np.int64), np.empty(rows * knn, dtype=np.int64)
    od = np.empty(rows * knn, dtype=np.float32)
    lib().ppo_get_knn_distances(_ptr(m), rows, cols, knn, _ptr(oi), _ptr(oj), _ptr(od))
    return oi, oj, od


def _coo(i, j, d):
    return (np.ascontiguousarray(i, dtype=np.int64), np.ascontiguousarray(j, dtype=np.int64),
            np.ascontiguousarray(d, dtype=np.float32))


def lower_rank(i, j, d, n_samples, knn, reciprocal_only=False, count_unique_distances=False, epsilon=0.0):
    """src/extend.cpp:146-243."""
    i, j, d = _coo(i, j, d)
    nnz = i.shape[0]
    oi, oj, od = np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.float32)
    c = lib().ppo_lower_rank(_ptr(i), _ptr(j), _ptr(d), nnz, n_samples, knn, int(reciprocal_only),
                             int(count_unique_distances), epsilon, _ptr(oi), _ptr(oj), _ptr(od))
    return oi[:c], oj[:c], od[:c]


def extend(i, j, d, qq, qr, knn):
    """src/extend.cpp:52-136; qr is (n_ref, n_query), qq (n_query, n_query)."""
    i, j, d = _coo(i, j, d)
    qq = np.ascontiguousarray(qq, dtype=np.float32)
    qr = np.ascontiguousarray(qr, dtype=np.float32)
    nr, nq = qr.shape
    cap = (nr + nq) * knn
    oi, oj, od = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.float32)
    c = lib().ppo_extend(_ptr(i), _ptr(j), _ptr(d), i.shape[0], _ptr(qq), _ptr(qr), nr, nq, knn, _ptr(oi), _ptr(oj), _ptr(od))
    return oi[:c], oj[:c], od[:c]


def long_to_square(vec, n):
    v = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1)
    sq = np.empty((n, n), dtype=np.float32)
    lib().ppo_long_to_square(_ptr(v), 1, n, _ptr(sq))
    return sq


def square_to_long(sq):
    sq = np.ascontiguousarray(sq, dtype=np.float32)
    n = sq.shape[0]
    v = np.empty(n * (n - 1) // 2, dtype=np.float32)
    lib().ppo_square_to_long(_ptr(sq), n, _ptr(v))
    return v


def long_to_square_multi(rr, qr, qq, R, Q):
    rr, qr, qq = (np.ascontiguousarray(x, dtype=np.float32).reshape(-1) for x in (rr, qr, qq))
    sq = np.empty((R + Q, R + Q), dtype=np.float32)
    lib().ppo_long_to_square_multi(_ptr(rr), 1, _ptr(qr), 1, _ptr(qq), 1, R, Q, _ptr(sq))
    return sq


def square_to_condensed(i, j, n):
    return int(lib().ppo_square_to_condensed(i, j, n))


def calc_row_idx(k, n):
    return int(lib().ppo_calc_row_idx(k, n))


def calc_col_idx(k, i, n):
    return int(lib().ppo_calc_col_idx(k, i, n))


# --------------------------------------------------------------------------------------------
# NumPy restatements (small cases; a different formulation from the C code on purpose)
# --------------------------------------------------------------------------------------------
def _unslice(words, ss64):
    words = np.ascontiguousarray(words, dtype="<u8")
    lead = words.shape[:-1]
    w = words.reshape(lead + (ss64, BBITS))
    sig = np.zeros(lead + (ss64, 64), dtype=np.uint16)
    for b in range(BBITS):
        plane = np.ascontiguousarray(w[..., b]).view(np.uint8).reshape(lead + (ss64, 8))
        sig |= np.unpackbits(plane, axis=-1, bitorder="little").astype(np.uint16) << np.uint16(b)
    return sig.reshape(lead + (ss64 * 64,))


def pair_rows(n_ref, n_qry=None):
    """(i, j) index arrays in output row order (utils.py:199-226): self -> i<j row-major;
    non-self -> i = query (slow), j = ref (fast)."""
    if n_qry is None:
        i, j = np.triu_indices(n_ref, k=1)
        return i, j
    q, r = np.divmod(np.arange(n_ref * n_qry), n_ref)
    return q, r


def counts_numpy(ref, qry=None):
    """c_k = number of bins whose 14-bit signatures are equal — the b-bit MinHash definition
    (BinDash, citation.py:35-38), computed on UN-sliced signatures."""
    ss64 = ref.shape[-1] // BBITS
    sr = _unslice(ref, ss64)
    i, j = pair_rows(ref.shape[0], None if qry is None else qry.shape[0])
    sa = sr if qry is None else _unslice(qry, ss64)
    out = np.empty((len(i), ref.shape[1]), dtype=np.uint32)
    step = 4096
    for s in range(0, len(i), step):
        out[s:s + step] = (sa[i[s:s + step]] == sr[j[s:s + step]]).sum(axis=-1)
    return out


def regress_numpy(jac, kmers, S):
    """(core, acc) per row from per-k Jaccards via numpy.linalg.lstsq on [1, k] (the design matrix
    of PopPUNK/sketchlib.py:541,652-660), with the < 5/S truncation and the <= 0 clamps."""
    jac = np.asarray(jac, dtype=np.float64)
    kmers = np.asarray(kmers, dtype=np.float64)
    out = np.zeros((jac.shape[0], 2), dtype=np.float32)
    ndeg = 0
    tol = 5.0 / S
    for r in range(jac.shape[0]):
        below = np.nonzero(jac[r] < tol)[0]
        n = below[0] if len(below) else len(kmers)
        if n < 2:
            ndeg += 1
            continue
        X = np.stack([np.ones(n), kmers[:n]], axis=1)
        (alpha, beta), *_ = np.linalg.lstsq(X, np.log(jac[r, :n]), rcond=None)
        out[r, 0] = 1.0 - np.exp(beta) if beta < 0 else 0.0
        out[r, 1] = 1.0 - np.exp(alpha) if alpha < 0 else 0.0
    return out, ndeg


def jaccard_numpy(counts, S, rand_table=None, ref_cluster=None, qry_cluster=None, n_ref=None,
                  n_qry=None):
    counts = np.asarray(counts, dtype=np.float64)
    jobs = counts / S
    if rand_table is None:
        return jobs
    i, j = pair_rows(n_ref, n_qry)
    cq = (ref_cluster if n_qry is None else qry_cluster)[i]
    cr = ref_cluster[j]
    r = rand_table.astype(np.float64)[cr, cq]  # [rows][K]
    return np.maximum(jobs - r, 0.0) / (1.0 - r)


def assign_threshold_numpy(dists, slope, x_max, y_max):
    """float32 restatement of src/boundary.cpp:42-80 (each op rounded to float32, no FMA)."""
    d = np.asarray(dists, dtype=np.float32)
    x0, y0 = d[:, 0], d[:, 1]
    xm, ym = np.float32(x_max), np.float32(y_max)
    if slope == 2:
        if xm == 0 or ym == 0:
            s = np.sqrt((x0 * x0 + y0 * y0).astype(np.float32)).astype(np.float32)
        else:
            s = ((y0 * xm).astype(np.float32) + (x0 * ym).astype(np.float32)).astype(np.float32) - np.float32(xm * ym)
    elif slope == 0:
        s = x0 - xm
    else:
        s = y0 - ym
    return np.sign(s).astype(np.float32)


# --------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own poppunk_refine sources built here (Makefile target `ref`)
# --------------------------------------------------------------------------------------------
REF_LIB = os.path.join(_HERE, "_ref", "libpprefine_ref.so")
_ref = None


def build_ref() -> str:
    """Compile /root/reference/src/{boundary,extend}.cpp (unchanged) into oracle/_ref — needs the reference tree."""
    subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)
    return REF_LIB


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib():
    """ctypes handle of oracle/_ref/libpprefine_ref.so (prebuilt; travels to the GPU box)."""
    global _ref
    if _ref is None:
        L = C.CDLL(REF_LIB)
        i64, vp, f32, ci = C.c_int64, C.c_void_p, C.c_float, C.c_int
        L.ppr_assign_threshold.argtypes = [vp, i64, ci, f32, f32, ci, vp]
        L.ppr_assign_threshold.restype = None
        L.ppr_edge_iterate.argtypes = [vp, i64, ci, f32, f32, vp, vp, i64]
        L.ppr_generate_tuples.argtypes = [vp, i64, ci, ci, ci, ci, vp, vp, i64]
        L.ppr_generate_all_tuples.argtypes = [ci, ci, ci, ci, vp, vp, i64]
        L.ppr_threshold_iterate_1d.argtypes = [vp, i64, vp, i64, ci, f32, f32, f32, f32, ci, vp, vp, vp, i64]
        L.ppr_threshold_iterate_2d.argtypes = [vp, i64, vp, i64, f32, vp, vp, vp, i64]
        L.ppr_get_knn_distances.argtypes = [vp, i64, i64, ci, i64, ci, vp, vp, vp, i64]
        L.ppr_lower_rank.argtypes = [vp, vp, vp, i64, i64, i64, ci, ci, f32, ci, vp, vp, vp, i64]
        L.ppr_extend.argtypes = [vp, vp, vp, i64, vp, vp, i64, i64, i64, ci, vp, vp, vp, i64]
        for name in ("ppr_edge_iterate", "ppr_generate_tuples", "ppr_generate_all_tuples", "ppr_threshold_iterate_1d",
                     "ppr_threshold_iterate_2d", "ppr_get_knn_distances", "ppr_lower_rank", "ppr_extend"):
            getattr(L, name).restype = i64
        _ref = L
    return _ref


class ref:
    """The reference functions themselves (same argument meaning as the oracle functions above)."""

    @staticmethod
    def assign_threshold(dists, slope, x_max, y_max, threads=1):
        d = np.ascontiguousarray(dists, dtype=np.float32)
        out = np.empty(d.shape[0], dtype=np.float32)
        ref_lib().ppr_assign_threshold(_ptr(d), d.shape[0], slope, x_max, y_max, threads, _ptr(out))
        return out

    @staticmethod
    def edge_iterate(dists, slope, x_max, y_max):
        d = np.ascontiguousarray(dists, dtype=np.float32)
        n = d.shape[0]
        oi, oj = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        c = ref_lib().ppr_edge_iterate(_ptr(d), n, slope, x_max, y_max, _ptr(oi), _ptr(oj), n)
        return oi[:c], oj[:c]

    @staticmethod
    def generate_tuples(assignments, within_label, self=True, num_ref=0, int_offset=0):
        a = np.ascontiguousarray(assignments, dtype=np.int32)
        n = a.shape[0]
        oi, oj = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        c = ref_lib().ppr_generate_tuples(_ptr(a), n, within_label, int(self), max(num_ref, 1), int_offset, _ptr(oi), _ptr(oj), n)
        return oi[:c], oj[:c]

    @staticmethod
    def generate_all_tuples(num_ref, num_queries=0, self=True, int_offset=0):
        n = num_ref * (num_ref - 1) // 2 if self else num_ref * num_queries
        oi, oj = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        c = ref_lib().ppr_generate_all_tuples(num_ref, num_queries, int(self), int_offset, _ptr(oi), _ptr(oj), n)
        return oi[:c], oj[:c]

    @staticmethod
    def threshold_iterate_1d(dists, offsets, slope, x0, y0, x1, y1):
        d = np.ascontiguousarray(dists, dtype=np.float32)
        off = np.ascontiguousarray(offsets, dtype=np.float64)
        n = d.shape[0]
        oi, oj, oo = (np.empty(n, dtype=np.int64) for _ in range(3))
        c = ref_lib().ppr_threshold_iterate_1d(_ptr(d), n, _ptr(off), off.shape[0], slope, x0, y0, x1, y1, 1,
                                               _ptr(oi), _ptr(oj), _ptr(oo), n)
        return oi[:c], oj[:c], oo[:c]

    @staticmethod
    def threshold_iterate_2d(dists, x_max, y_max):
        d = np.ascontiguousarray(dists, dtype=np.float32)
        xm = np.ascontiguousarray(x_max, dtype=np.float32)
        cap = d.shape[0] * max(1, xm.shape[0])
        oi, oj, oo = (np.empty(cap, dtype=np.int64) for _ in range(3))
        c = ref_lib().ppr_threshold_iterate_2d(_ptr(d), d.shape[0], _ptr(xm), xm.shape[0], y_max, _ptr(oi), _ptr(oj),
                                               _ptr(oo), cap)
        return oi[:c], oj[:c], oo[:c]

    @staticmethod
    def get_knn_distances(mat, knn):
        m = np.ascontiguousarray(mat, dtype=np.float32)
        rows, cols = m.shape
        cap = rows * knn
        oi, oj, od = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.float32)
        ref_lib().ppr_get_knn_distances(_ptr(m), rows, cols, knn, 0, 1, _ptr(oi), _ptr(oj), _ptr(od), cap)
        return oi, oj, od

    @staticmethod
    def lower_rank(i, j, d, n_samples, knn, reciprocal_only=False, count_unique_distances=False, epsilon=0.0):
        i, j, d = _coo(i, j, d)
        nnz = i.shape[0]
        oi, oj, od = np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.float32)
        c = ref_lib().ppr_lower_rank(_ptr(i), _ptr(j), _ptr(d), nnz, n_samples, knn, int(reciprocal_only),
                                     int(count_unique_distances), epsilon, 1, _ptr(oi), _ptr(oj), _ptr(od), nnz)
        return oi[:c], oj[:c], od[:c]

    @staticmethod
    def extend(i, j, d, qq, qr, knn):
        i, j, d = _coo(i, j, d)
        qq = np.ascontiguousarray(qq, dtype=np.float32)
        qr = np.ascontiguousarray(qr, dtype=np.float32)
        nr, nq = qr.shape
        cap = (nr + nq) * knn
        oi, oj, od = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.float32)
        c = ref_lib().ppr_extend(_ptr(i), _ptr(j), _ptr(d), i.shape[0], _ptr(qq), _ptr(qr), nr, nq, knn, 1,
                                 _ptr(oi), _ptr(oj), _ptr(od), cap)
        return oi[:c], oj[:c], od[:c]
