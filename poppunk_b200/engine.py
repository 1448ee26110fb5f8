"""Host-side engine: torch tensors hold the sketches / results in HBM, the work is done by the
hand-written sm_100a kernels in ``libppb.so`` through the C ABI (``include/ppb.h``).

torch is plumbing here (device memory, streams, ``torch.distributed``); there is no torch compute on
the hot path and no CPU fallback — every entry point raises if the CUDA library or a GPU is missing.

Replaces the native part of ``pp_sketchlib.queryDatabase`` (call sites PopPUNK/sketchlib.py:528-537,
584-593) once the sketches are in memory; ``poppunk_b200.sketchlib.queryDatabase`` is the drop-in wrapper.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import OUT_COUNTS, OUT_DISTS, OUT_JACCARD, BBITS, Boundary, check

__all__ = ["PackedSketches", "pack", "query", "query_host", "assign_threshold", "num_rows", "shard_rows",
           "query_sharded", "query_edges", "FusedExchange", "host_result", "visible_devices", "OUT_DISTS", "OUT_JACCARD", "OUT_COUNTS"]


def _require_cuda(device=None) -> torch.device:
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("poppunk_b200: no CUDA device visible — this engine has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("poppunk_b200: tensors must live on a CUDA device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _stream_ptr(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def num_rows(n_ref: int, n_qry: Optional[int] = None) -> int:
    """Rows of the result: condensed upper triangle (self) or query-major rectangle (utils.py:199-226)."""
    return n_ref * (n_ref - 1) // 2 if n_qry is None else n_ref * n_qry


@dataclass
class PackedSketches:
    """Sketches of ``n`` genomes in the kernel's lane-sliced HBM layout (csrc/ppb_kernels.cuh)."""
    data: torch.Tensor            # uint8 [ppb_packed_bytes]
    n: int
    K: int
    sketchsize64: int
    clusters: Optional[torch.Tensor] = None   # uint16-as-int16 [n] random-match cluster ids
    max_cluster: int = -1                     # largest cluster id (checked against the table size at query time)

    @property
    def device(self):
        return self.data.device


@dataclass
class DeviceClusters:
    """Random-match cluster ids that already live on the device (uint16 bit patterns in an int16 tensor), with their
    maximum known on the host — what a caller that packs repeatedly hands to :func:`pack` instead of a NumPy array."""
    ids: torch.Tensor
    max_id: int

    @staticmethod
    def upload(clusters, device) -> "DeviceClusters":
        cl_np = np.ascontiguousarray(clusters, dtype=np.uint16)
        return DeviceClusters(torch.from_numpy(cl_np.view(np.int16)).to(device), int(cl_np.max()) if cl_np.size else -1)


def as_device_sketches(sketches, device=None) -> torch.Tensor:
    """uint64 ``[n][K][W]`` (numpy or torch, host or device) -> int64-viewed CUDA tensor (bit-identical)."""
    dev = _require_cuda(device)
    if isinstance(sketches, np.ndarray):
        if sketches.dtype != np.uint64:
            raise TypeError("sketch arrays are uint64")
        sketches = torch.from_numpy(np.ascontiguousarray(sketches).view(np.int64))
    if sketches.dtype == torch.uint64:
        sketches = sketches.view(torch.int64)
    if sketches.dtype != torch.int64 or sketches.dim() != 3:
        raise TypeError("sketches must be uint64 [n][K][W]")
    return sketches.to(dev, non_blocking=True).contiguous()


def pack(sketches, idx=None, clusters=None, device=None) -> PackedSketches:
    """Gather genomes ``idx`` (list order = output order, the rList/qList semantics of
    pp_sketchlib.queryDatabase) and re-lay them for the kernel.  One memory-bound pass."""
    L = _lib.load()
    sk = as_device_sketches(sketches, device)
    dev = sk.device
    n_src, K, W = sk.shape
    if W % BBITS:
        raise ValueError("W must be sketchsize64 * 14")
    ss64 = W // BBITS
    idx_t = None
    n = n_src
    if idx is not None:
        idx_t = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(dev)
        n = idx_t.numel()
        if n and (int(idx_t.min()) < 0 or int(idx_t.max()) >= n_src):
            raise IndexError("genome index out of range")
    out = torch.empty(L.ppb_packed_bytes(n, K, ss64), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(L.ppb_pack_dev(sk.data_ptr(), n_src, idx_t.data_ptr() if idx_t is not None else None, n, K, ss64,
                             out.data_ptr(), _stream_ptr(dev)), "ppb_pack_dev")
    cl = None
    if isinstance(clusters, DeviceClusters):      # already resident (list order): nothing to upload, nothing to wait for
        if clusters.ids.numel() != n or clusters.ids.device != dev:
            raise ValueError("one cluster id per packed genome, on the sketches' device")
        return PackedSketches(out, n, K, ss64, clusters.ids, clusters.max_id)
    if clusters is not None:
        cl_np = np.ascontiguousarray(clusters, dtype=np.uint16)
        if idx is not None:
            cl_np = cl_np[np.asarray(idx, dtype=np.int64)]
        if cl_np.shape != (n,):
            raise ValueError("one cluster id per genome")
        cl = torch.from_numpy(cl_np.view(np.int16)).to(dev)
        return PackedSketches(out, n, K, ss64, cl, int(cl_np.max()) if n else -1)
    return PackedSketches(out, n, K, ss64, cl)


def _boundary(boundary) -> Optional[Boundary]:
    if boundary is None:
        return None
    if isinstance(boundary, Boundary):
        return boundary
    slope, x_max, y_max, *scale = boundary
    sx, sy = (scale + [1.0, 1.0])[:2] if scale else (1.0, 1.0)
    return Boundary(int(slope), float(x_max), float(y_max), float(sx), float(sy))


def query(ref: PackedSketches, qry: Optional[PackedSketches], kmers: Sequence[int], rand_table=None,
          row_begin: int = 0, row_end: Optional[int] = None, out_mode: int = OUT_DISTS, boundary=None,
          want_out: bool = True, out: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
          n_degenerate: Optional[torch.Tensor] = None):
    """Launch the distance kernel on the current stream; everything stays on the device.

    Returns ``(out, labels, n_degenerate)``: ``out`` float32 ``[rows][2]`` (DISTS), float32 ``[rows][K]``
    (JACCARD) or int32 ``[rows][K]`` (COUNTS; uint32 bit pattern); ``labels`` int8 ``[rows]`` when a
    ``boundary=(slope, x_max, y_max[, scale_x, scale_y])`` is given; ``n_degenerate`` int64 ``[1]`` device
    counter of rows whose fit had fewer than two usable k (they are (0, 0)).
    """
    L = _lib.load()
    dev = _require_cuda(ref.device)
    self_mode = qry is None
    if not self_mode and (qry.K != ref.K or qry.sketchsize64 != ref.sketchsize64 or qry.device != ref.device):
        raise ValueError("query and reference sketches differ in k-mers, sketch size or device")
    kmers_np = np.ascontiguousarray(kmers, dtype=np.int32)
    K = ref.K
    if kmers_np.shape != (K,):
        raise ValueError("klist length does not match the packed sketches")
    total = num_rows(ref.n, None if self_mode else qry.n)
    if row_end is None:
        row_end = total
    rows = row_end - row_begin
    if rows < 0:
        raise ValueError("bad row range")
    bnd = _boundary(boundary)
    tab = None
    C_ = 0
    if rand_table is not None:
        tab = torch.as_tensor(np.ascontiguousarray(rand_table, dtype=np.float32) if isinstance(rand_table, np.ndarray)
                              else rand_table, dtype=torch.float32).to(dev).contiguous()
        C_ = tab.shape[0]
        if tuple(tab.shape) != (C_, C_, K):
            raise ValueError("random-match table must be [C][C][K]")
        if ref.clusters is None or (not self_mode and qry.clusters is None):
            raise ValueError("random-match table given but sketches carry no cluster ids")
        if ref.max_cluster >= C_ or (not self_mode and qry.max_cluster >= C_):
            raise ValueError("cluster id out of range for the random-match table (the kernel indexes the table with it)")
    if want_out and out is None:
        if out_mode == OUT_DISTS:
            out = torch.empty((rows, 2), dtype=torch.float32, device=dev)
        elif out_mode == OUT_JACCARD:
            out = torch.empty((rows, K), dtype=torch.float32, device=dev)
        else:
            out = torch.empty((rows, K), dtype=torch.int32, device=dev)
    if bnd is not None and labels is None:
        labels = torch.empty(rows, dtype=torch.int8, device=dev)
    if n_degenerate is None:
        n_degenerate = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(L.ppb_query_dev(
            ref.data.data_ptr(), ref.n, None if self_mode else qry.data.data_ptr(), 0 if self_mode else qry.n,
            kmers_np.ctypes.data, K, ref.sketchsize64,
            tab.data_ptr() if tab is not None else None, C_,
            ref.clusters.data_ptr() if (tab is not None) else None,
            qry.clusters.data_ptr() if (tab is not None and not self_mode) else None,
            row_begin, row_end, out_mode, out.data_ptr() if out is not None else None,
            C.byref(bnd) if bnd is not None else None, labels.data_ptr() if bnd is not None else None,
            n_degenerate.data_ptr(), _stream_ptr(dev)), "ppb_query_dev")
    return out, (labels if bnd is not None else None), n_degenerate


def assign_threshold(dists: torch.Tensor, slope: int, x_max: float, y_max: float) -> torch.Tensor:
    """Device twin of ``poppunk_refine.assignThreshold`` (src/boundary.cpp:60-80): float32 [n] in {-1,0,1}."""
    L = _lib.load()
    dev = _require_cuda(dists.device)
    if dists.dtype != torch.float32 or dists.dim() != 2 or dists.shape[1] != 2 or not dists.is_contiguous():
        raise TypeError("distMat must be float32, C-contiguous, shape (n, 2)")  # python_bindings.cpp:82 .noconvert()
    out = torch.empty(dists.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(L.ppb_assign_threshold_dev(dists.data_ptr(), dists.shape[0], int(slope), float(x_max), float(y_max),
                                         out.data_ptr(), _stream_ptr(dev)), "ppb_assign_threshold_dev")
    return out


# --------------------------------------------------------------------------------------------
# host-buffer path (H2D + pack + kernels + D2H inside the library) — what the drop-in wrapper uses
# --------------------------------------------------------------------------------------------
def host_result(shape, dtype) -> np.ndarray:
    """A NumPy array for a result the caller will own, backed by the library's host pool (``ppb_host_alloc``):
    huge-page backed, kept when the array dies, page-locked from its second reuse on, so that repeated calls of a
    process receive their result by direct DMA.  The block returns to the pool when the array (and every view of it) dies."""
    import weakref
    L = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dtype)
    ptr = L.ppb_host_alloc(nbytes)
    if not ptr:
        raise MemoryError(f"ppb_host_alloc({nbytes}) failed: {L.ppb_last_error().decode()}")
    buf = (C.c_byte * nbytes).from_address(ptr)
    weakref.finalize(buf, L.ppb_host_free, ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def visible_devices(first: int = 0):
    """Devices a host-buffer call uses: every visible CUDA device, ``first`` (PopPUNK's ``deviceid``) leading.
    ``PPB_DEVICES`` overrides: ``single`` = only ``first``; ``0,2,3`` = that list; an integer N = the first N."""
    import os
    L = _lib.load()
    n = L.ppb_device_count()
    if n <= 0:
        raise RuntimeError("poppunk_b200: no CUDA device visible — this engine has no CPU fallback")
    if not 0 <= first < n:
        raise RuntimeError(f"poppunk_b200: device id {first} is not one of the {n} visible CUDA devices")
    knob = os.environ.get("PPB_DEVICES", "").strip().lower()
    if knob == "single":
        return [first]
    if "," in knob:
        devs = [int(v) for v in knob.split(",") if v.strip() != ""]
        if any(not 0 <= d < n for d in devs) or len(set(devs)) != len(devs):
            raise RuntimeError(f"PPB_DEVICES={knob!r} does not name distinct visible devices (0..{n - 1})")
        return devs
    order = [first] + [d for d in range(n) if d != first]
    if knob.isdigit() and int(knob) >= 1:
        order = order[:int(knob)]
    return order


def query_host(ref: np.ndarray, qry: Optional[np.ndarray], kmers, rand_table=None, ref_cluster=None,
               qry_cluster=None, row_begin: int = 0, row_end: Optional[int] = None, out_mode: int = OUT_DISTS,
               boundary=None, out: Optional[np.ndarray] = None, want_out: bool = True, device_id: int = 0,
               devices: Optional[Sequence[int]] = None):
    """``ppb_query_host_multi``: NumPy in, NumPy out, on ``devices`` (default: the one ``device_id``).
    Returns ``(out, labels, n_degenerate)``.  Without ``out`` the result array comes from :func:`host_result`."""
    L = _lib.load()
    if L.ppb_device_count() <= 0:
        raise RuntimeError("poppunk_b200: no CUDA device visible — this engine has no CPU fallback")
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    n_ref, K, W = ref.shape
    if W % BBITS:
        raise ValueError("W must be sketchsize64 * 14")
    ss64 = W // BBITS
    n_qry = 0
    if qry is not None:
        qry = np.ascontiguousarray(qry, dtype=np.uint64)
        n_qry = qry.shape[0]
        if qry.shape[1:] != ref.shape[1:]:
            raise ValueError("query and reference sketches differ in k-mers or sketch size")
    kmers_np = np.ascontiguousarray(kmers, dtype=np.int32)
    total = num_rows(n_ref, None if qry is None else n_qry)
    if row_end is None:
        row_end = total
    rows = row_end - row_begin
    C_ = 0
    if rand_table is not None:
        rand_table = np.ascontiguousarray(rand_table, dtype=np.float32)
        C_ = rand_table.shape[0]
        if rand_table.shape != (C_, C_, K):
            raise ValueError("random-match table must be [C][C][K]")
        ref_cluster = np.ascontiguousarray(ref_cluster, dtype=np.uint16)
        if ref_cluster.shape != (n_ref,) or (n_ref and int(ref_cluster.max()) >= C_):
            raise ValueError("reference cluster ids must be one per genome and < C")
        if qry is not None:
            qry_cluster = np.ascontiguousarray(qry_cluster, dtype=np.uint16)
            if qry_cluster.shape != (n_qry,) or (n_qry and int(qry_cluster.max()) >= C_):
                raise ValueError("query cluster ids must be one per genome and < C")
    if want_out and out is None:
        width, dt = (2, np.float32) if out_mode == OUT_DISTS else (K, np.float32 if out_mode == OUT_JACCARD else np.uint32)
        out = host_result((rows, width), dt)
    elif out is not None:
        width, dt = (2, np.float32) if out_mode == OUT_DISTS else (K, np.float32 if out_mode == OUT_JACCARD else np.uint32)
        if not out.flags.c_contiguous or not out.flags.writeable or out.dtype != dt or out.size != rows * width:
            raise ValueError(f"out must be a writeable C-contiguous {np.dtype(dt).name} array of {rows} x {width} elements")
    bnd = _boundary(boundary)
    labels = host_result((rows,), np.int8) if bnd is not None else None   # pool block, like the distances
    ndeg = C.c_int64(0)
    devs = [int(device_id)] if devices is None else [int(d) for d in devices]
    dev_arr = (C.c_int32 * len(devs))(*devs)

    def ptr(a):
        return None if a is None else a.ctypes.data

    check(L.ppb_query_host_multi(ptr(ref), n_ref, ptr(qry), n_qry, ptr(kmers_np), K, ss64, BBITS, ptr(rand_table), C_,
                                 ptr(ref_cluster) if rand_table is not None else None,
                                 ptr(qry_cluster) if (rand_table is not None and qry is not None) else None,
                                 row_begin, row_end, out_mode, ptr(out), C.byref(bnd) if bnd is not None else None,
                                 ptr(labels), C.byref(ndeg), dev_arr, len(devs)), "ppb_query_host_multi")
    return out, labels, int(ndeg.value)


# --------------------------------------------------------------------------------------------
# multi-GPU: static row shards, one all-gather
# --------------------------------------------------------------------------------------------
def shard_rows(total_rows: int, world_size: int, rank: int) -> Tuple[int, int, int]:
    """Equal contiguous row slices (the last ones padded): returns ``(begin, end, slice_len)``.

    Self mode rows are condensed-order, so a slice is a band of whole rows i plus two partial ones — load is
    balanced by pair count.  Non-self rows are query-major, so a slice is a range of queries."""
    slice_len = -(-total_rows // world_size) if total_rows else 0
    b = min(total_rows, rank * slice_len)
    e = min(total_rows, b + slice_len)
    return b, e, slice_len


def query_sharded(ref: PackedSketches, qry: Optional[PackedSketches], kmers, rand_table=None,
                  out_mode: int = OUT_DISTS, group=None, gather: bool = True, _query_fn=None):
    """Every rank holds the (replicated) sketches, computes its static row shard and — if ``gather`` — one
    ``all_gather_into_tensor`` (NCCL over NVLink on GPUs) reassembles the full row-ordered result on every
    rank.  Returns ``(out, n_degenerate_total)``; without ``gather`` ``out`` is the local shard."""
    import torch.distributed as dist
    run = _query_fn or query   # tests substitute a CPU stand-in to exercise the sharding/gather logic
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = num_rows(ref.n, None if qry is None else qry.n)
    b, e, slice_len = shard_rows(total, world, rank)
    width = 2 if out_mode == OUT_DISTS else ref.K
    dtype = torch.int32 if out_mode == OUT_COUNTS else torch.float32
    if world == 1 or not gather:
        out, _, ndeg = run(ref, qry, kmers, rand_table, b, e, out_mode)
        if world > 1:
            dist.all_reduce(ndeg, group=group)
        return out, ndeg
    full = torch.empty((world * slice_len, width), dtype=dtype, device=ref.device)
    mine = full[rank * slice_len:(rank + 1) * slice_len]
    _, _, ndeg = run(ref, qry, kmers, rand_table, b, e, out_mode, out=mine[:e - b])
    dist.all_gather_into_tensor(full, mine, group=group)
    dist.all_reduce(ndeg, group=group)
    return full[:total], ndeg


def query_edges(ref: PackedSketches, qry: Optional[PackedSketches], kmers, boundary, rand_table=None,
                row_begin: int = 0, row_end: Optional[int] = None, include_boundary: bool = False,
                capacity: Optional[int] = None, int_offset: int = 0, sort: bool = True):
    """Distances -> boundary test -> edge list in ONE kernel pass: nothing of size n_pairs is written.

    The fused form of ``queryDatabase`` + ``model.assign`` + ``generateTuples`` (assign.py:502-510, 593-601,
    network.py:1180-1184): returns ``(i, j, n_edges, n_degenerate)`` with int64 device tensors of the
    within-boundary pairs (``line_dist < 0``, or ``<= 0`` with ``include_boundary`` as edge_iterate does),
    in the reference's row order when ``sort`` (the kernel appends unordered).  ``n_edges`` may exceed
    ``capacity`` (default: 1/8 of the rows, at least 1 Mi) — then only ``capacity`` edges were kept."""
    L = _lib.load()
    dev = _require_cuda(ref.device)
    self_mode = qry is None
    kmers_np = np.ascontiguousarray(kmers, dtype=np.int32)
    total = num_rows(ref.n, None if self_mode else qry.n)
    if row_end is None:
        row_end = total
    rows = row_end - row_begin
    if capacity is None:
        capacity = max(1 << 20, rows // 8)
    capacity = max(1, min(capacity, max(rows, 1)))
    bnd = _boundary(boundary)
    tab, C_ = None, 0
    if rand_table is not None:
        tab = torch.as_tensor(rand_table, dtype=torch.float32).to(dev).contiguous()
        C_ = tab.shape[0]
        if ref.clusters is None or (not self_mode and qry.clusters is None):
            raise ValueError("random-match table given but sketches carry no cluster ids")
        if ref.max_cluster >= C_ or (not self_mode and qry.max_cluster >= C_):
            raise ValueError("cluster id out of range for the random-match table")
    edge_rows = torch.empty(capacity, dtype=torch.int64, device=dev)
    n_edges = torch.zeros(1, dtype=torch.int64, device=dev)
    ndeg = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(L.ppb_query_edges_dev(
            ref.data.data_ptr(), ref.n, None if self_mode else qry.data.data_ptr(), 0 if self_mode else qry.n,
            kmers_np.ctypes.data, ref.K, ref.sketchsize64, tab.data_ptr() if tab is not None else None, C_,
            ref.clusters.data_ptr() if tab is not None else None,
            qry.clusters.data_ptr() if (tab is not None and not self_mode) else None,
            row_begin, row_end, C.byref(bnd), int(include_boundary), edge_rows.data_ptr(), capacity,
            n_edges.data_ptr(), None, None, ndeg.data_ptr(), _stream_ptr(dev)), "ppb_query_edges_dev")
        n = int(n_edges.item())
        kept = edge_rows[:min(n, capacity)]
        if sort:                                    # the reference's row order (the kernel appends unordered)
            check(L.ppb_sort_rows_dev(kept.data_ptr(), kept.numel(), max(total - 1, 0), _stream_ptr(dev)),
                  "ppb_sort_rows_dev")
        oi = torch.empty_like(kept)
        oj = torch.empty_like(kept)
        check(L.ppb_rows_to_pairs_dev(kept.data_ptr(), kept.numel(), int(self_mode), ref.n, int(int_offset),
                                      oi.data_ptr(), oj.data_ptr(), _stream_ptr(dev)), "ppb_rows_to_pairs_dev")
    return oi, oj, n, int(ndeg.item())


class FusedExchange:
    """Multi-GPU result exchange fused into the distance kernel (no all-gather).

    Every rank owns a FULL ``[total_rows][2]`` float32 result buffer allocated as torch symmetric memory (peer
    mapped over NVLink).  Each rank's kernel stores its rows, from the epilogue warps, into all G buffers —
    through one NVSwitch multicast store (``multimem.st``) when the fabric supports it, otherwise G coalesced
    peer stores — so when every rank has finished and passed the barrier, every GPU holds the whole
    row-ordered result.  The output rate of a GPU (~53 GB/s at 6.7 Gpairs/s) is far below NVLink's 900 GB/s,
    so the exchange hides completely under the LOP3 stream.
    """

    def __init__(self, total_rows: int, device, group=None, use_multicast: bool = True):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("at most 16 peers")
        self.total = total_rows
        self.full = symm_mem.empty((max(total_rows, 1), 2), dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.full, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.mc_ptr = mc if (use_multicast and mc) else 0

    def run(self, ref: PackedSketches, qry: Optional[PackedSketches], kmers, rand_table=None,
            n_degenerate: Optional[torch.Tensor] = None):
        """Compute this rank's row shard and scatter it to every rank; returns (full_result, n_degenerate).
        The result is complete on every rank after the trailing device-side barrier."""
        L = _lib.load()
        dev = _require_cuda(ref.device)
        self_mode = qry is None
        total = num_rows(ref.n, None if self_mode else qry.n)
        if total != self.total:
            raise ValueError("exchange buffer was sized for a different problem")
        b, e, _ = shard_rows(total, self.world, self.rank)
        kmers_np = np.ascontiguousarray(kmers, dtype=np.int32)
        tab, C_ = None, 0
        if rand_table is not None:
            tab = torch.as_tensor(rand_table, dtype=torch.float32).to(dev).contiguous()
            C_ = tab.shape[0]
            if ref.clusters is None or (not self_mode and qry.clusters is None):
                raise ValueError("random-match table given but sketches carry no cluster ids")
            if ref.max_cluster >= C_ or (not self_mode and qry.max_cluster >= C_):
                raise ValueError("cluster id out of range for the random-match table")
        if n_degenerate is None:
            n_degenerate = torch.zeros(1, dtype=torch.int64, device=dev)
        peers = (C.c_void_p * self.world)(*self.peer_ptrs)
        self.handle.barrier(channel=0)   # nobody is still reading the previous result
        with torch.cuda.device(dev):
            check(L.ppb_query_dev_fused(
                ref.data.data_ptr(), ref.n, None if self_mode else qry.data.data_ptr(), 0 if self_mode else qry.n,
                kmers_np.ctypes.data, ref.K, ref.sketchsize64,
                tab.data_ptr() if tab is not None else None, C_,
                ref.clusters.data_ptr() if tab is not None else None,
                qry.clusters.data_ptr() if (tab is not None and not self_mode) else None,
                b, e, None, peers, self.world, self.mc_ptr if self.mc_ptr else None,
                n_degenerate.data_ptr(), _stream_ptr(dev)), "ppb_query_dev_fused")
        self.handle.barrier(channel=1)   # every rank's stores have landed everywhere
        return self.full[:total], n_degenerate
