"""poppunk_b200 — B200-native core/accessory sketch-distance engine behind PopPUNK's ``queryDatabase``.

Only the hot path lives here: ``sketchlib.queryDatabase`` (drop-in wrapper),
``engine`` (torch tensors in HBM -> sm_100a kernels through the C ABI in ``include/ppb.h``),
``refine.assignThreshold`` and ``synth`` (seeded synthetic sketches for tests and bench).
"""
__version__ = "0.1.0"
