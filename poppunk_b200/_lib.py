"""ctypes binding of libppb.so — the C ABI declared in include/ppb.h.  No CPU fallback: if the CUDA
library is missing or fails to load, importing the engine raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPB_LIB") or os.path.join(HERE, "libppb.so")   # PPB_LIB: kernel-variant experiments

OUT_DISTS, OUT_JACCARD, OUT_COUNTS = 0, 1, 2
BBITS = 14
MAX_K = 32

# every symbol include/ppb.h declares (tests check the .so exports each of them)
SYMBOLS = [
    "ppb_version", "ppb_last_error", "ppb_device_count", "ppb_square_to_condensed", "ppb_calc_row_idx",
    "ppb_calc_col_idx", "ppb_num_rows", "ppb_packed_bytes", "ppb_pack_dev", "ppb_query_dev", "ppb_query_dev_fused",
    "ppb_assign_threshold_dev", "ppb_query_host", "ppb_assign_threshold_host", "ppb_microbench_dev",
    "ppb_launch_count", "ppb_release_workspace", "ppb_query_edges_dev", "ppb_rows_to_pairs_dev",
    "ppb_edges_scratch_bytes", "ppb_edges_from_dists_dev", "ppb_edges_from_labels_dev", "ppb_long_to_square_dev",
    "ppb_square_to_long_dev", "ppb_long_to_square_multi_dev", "ppb_plan_host_chunks",
    "ppb_generate_all_tuples_dev", "ppb_threshold_iterate_1d_dev", "ppb_threshold_iterate_2d_dev", "ppb_knn_dev",
    "ppb_lower_rank_dev", "ppb_extend_dev", "ppb_plan_tiles", "ppb_pack_part_dev", "ppb_query_host_multi",
    "ppb_plan_device_shards", "ppb_host_alloc", "ppb_host_free", "ppb_host_pool_stats", "ppb_microbench_mix_dev", "ppb_sort_rows_dev",
]


class Boundary(C.Structure):
    """``ppb_boundary`` (include/ppb.h)."""
    _fields_ = [("slope", C.c_int32), ("x_max", C.c_float), ("y_max", C.c_float),
                ("scale_x", C.c_float), ("scale_y", C.c_float)]


class PpbError(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m poppunk_b200.build` (nvcc, sm_100a). "
            "This engine has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    i64, i32, vp, f32 = C.c_int64, C.c_int32, C.c_void_p, C.c_float
    L.ppb_version.restype = C.c_int
    L.ppb_last_error.restype = C.c_char_p
    L.ppb_device_count.restype = C.c_int
    L.ppb_launch_count.restype = i64
    L.ppb_square_to_condensed.argtypes = [i64, i64, i64]
    L.ppb_square_to_condensed.restype = i64
    L.ppb_calc_row_idx.argtypes = [i64, i64]
    L.ppb_calc_row_idx.restype = i64
    L.ppb_calc_col_idx.argtypes = [i64, i64, i64]
    L.ppb_calc_col_idx.restype = i64
    L.ppb_num_rows.argtypes = [i64, i64, C.c_int]
    L.ppb_num_rows.restype = i64
    L.ppb_packed_bytes.argtypes = [i64, i32, i32]
    L.ppb_packed_bytes.restype = C.c_size_t
    L.ppb_pack_dev.argtypes = [vp, i64, vp, i64, i32, i32, vp, vp]
    L.ppb_pack_dev.restype = C.c_int
    L.ppb_query_dev.argtypes = [vp, i64, vp, i64, vp, i32, i32, vp, i32, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp]
    L.ppb_query_dev.restype = C.c_int
    L.ppb_query_dev_fused.argtypes = [vp, i64, vp, i64, vp, i32, i32, vp, i32, vp, vp, i64, i64, vp, vp, i32, vp, vp, vp]
    L.ppb_query_dev_fused.restype = C.c_int
    L.ppb_assign_threshold_dev.argtypes = [vp, i64, i32, f32, f32, vp, vp]
    L.ppb_assign_threshold_dev.restype = C.c_int
    L.ppb_query_host.argtypes = [vp, i64, vp, i64, vp, i32, i32, i32, vp, i32, vp, vp, i64, i64, i32, vp, vp, vp,
                                 vp, i32]
    L.ppb_query_host.restype = C.c_int
    L.ppb_query_host_multi.argtypes = L.ppb_query_host.argtypes[:-1] + [vp, i32]
    L.ppb_query_host_multi.restype = C.c_int
    L.ppb_plan_device_shards.argtypes = [i64, i64, i32, i64, i64, i32, vp]
    L.ppb_plan_device_shards.restype = i64
    L.ppb_pack_part_dev.argtypes = [vp, vp, i64, i64, i64, i32, i32, vp, i32, vp]
    L.ppb_pack_part_dev.restype = C.c_int
    L.ppb_host_alloc.argtypes = [C.c_size_t]
    L.ppb_host_alloc.restype = vp
    L.ppb_host_free.argtypes = [vp]
    L.ppb_host_free.restype = C.c_int
    L.ppb_host_pool_stats.argtypes = [vp, vp, vp]
    L.ppb_host_pool_stats.restype = C.c_int
    L.ppb_assign_threshold_host.argtypes = [vp, i64, i32, f32, f32, vp, i32]
    L.ppb_assign_threshold_host.restype = C.c_int
    L.ppb_release_workspace.restype = C.c_int
    L.ppb_query_edges_dev.argtypes = [vp, i64, vp, i64, vp, i32, i32, vp, i32, vp, vp, i64, i64, vp, i32, vp, i64, vp,
                                      vp, vp, vp, vp]
    L.ppb_query_edges_dev.restype = C.c_int
    L.ppb_rows_to_pairs_dev.argtypes = [vp, i64, i32, i64, i64, vp, vp, vp]
    L.ppb_rows_to_pairs_dev.restype = C.c_int
    L.ppb_sort_rows_dev.argtypes = [vp, i64, i64, vp]
    L.ppb_sort_rows_dev.restype = C.c_int
    L.ppb_edges_scratch_bytes.argtypes = [i64]
    L.ppb_edges_scratch_bytes.restype = C.c_size_t
    L.ppb_edges_from_dists_dev.argtypes = [vp, i64, i64, i32, f32, f32, vp, vp, i64, vp, vp, vp]
    L.ppb_edges_from_dists_dev.restype = C.c_int
    L.ppb_edges_from_labels_dev.argtypes = [vp, i32, i64, i32, i32, i64, i64, vp, vp, i64, vp, vp, vp]
    L.ppb_edges_from_labels_dev.restype = C.c_int
    L.ppb_long_to_square_dev.argtypes = [vp, i64, i64, vp, vp]
    L.ppb_long_to_square_dev.restype = C.c_int
    L.ppb_square_to_long_dev.argtypes = [vp, i64, vp, vp]
    L.ppb_square_to_long_dev.restype = C.c_int
    L.ppb_long_to_square_multi_dev.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, vp, vp]
    L.ppb_long_to_square_multi_dev.restype = C.c_int
    L.ppb_generate_all_tuples_dev.argtypes = [i64, i64, i32, i64, vp, vp, vp]
    L.ppb_generate_all_tuples_dev.restype = C.c_int
    L.ppb_threshold_iterate_1d_dev.argtypes = [vp, i64, vp, i32, i32, f32, f32, f32, f32, vp, vp, vp, i64, vp, vp]
    L.ppb_threshold_iterate_1d_dev.restype = C.c_int
    L.ppb_threshold_iterate_2d_dev.argtypes = [vp, i64, vp, i32, f32, vp, vp, vp, i64, vp, vp]
    L.ppb_threshold_iterate_2d_dev.restype = C.c_int
    L.ppb_knn_dev.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp]
    L.ppb_knn_dev.restype = C.c_int
    L.ppb_lower_rank_dev.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, f32, vp, vp, vp, vp, vp]
    L.ppb_lower_rank_dev.restype = C.c_int
    L.ppb_extend_dev.argtypes = [vp, vp, vp, i64, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp]
    L.ppb_extend_dev.restype = C.c_int
    L.ppb_plan_tiles.argtypes = [i64, i64, i32, i64, i64, i32, i32, vp, i64]
    L.ppb_plan_tiles.restype = i64
    L.ppb_plan_host_chunks.argtypes = [i64, i64, i32, i64, i64, i64, vp, i64]
    L.ppb_plan_host_chunks.restype = i64
    L.ppb_microbench_dev.argtypes = [i32, i64, vp, vp, vp]
    L.ppb_microbench_dev.restype = C.c_int
    L.ppb_microbench_mix_dev.argtypes = [i32, i32, i32, i64, vp, vp, vp]
    L.ppb_microbench_mix_dev.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str = "libppb"):
    if rc != 0:
        msg = load().ppb_last_error()
        raise PpbError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
