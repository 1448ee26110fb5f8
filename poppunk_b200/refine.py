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
