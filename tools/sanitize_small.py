#!/usr/bin/env python
"""Small invocations of every kernel, meant to run under compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Each result is also checked against the oracle, so a sanitizer-clean run is also a parity run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from poppunk_b200 import engine, refine, reshape, synth  # noqa: E402
import oracle  # noqa: E402

kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
for n, ss64 in ((200, 16), (70, 40)):            # single-slice and multi-slice (S = 2560) paths
    ref = synth.synth_sketches(n, kmers, ss64, seed=3)
    qry = synth.synth_sketches(37, kmers, ss64, seed=3, sample_seed=1)
    tab, cl, qcl = synth.random_match_table(kmers, 3), synth.synth_clusters(n, 3), synth.synth_clusters(37, 3)
    pr, pq = engine.pack(ref, clusters=cl), engine.pack(qry, clusters=qcl)
    for q_np, q_pk, q_cl in ((None, None, None), (qry, pq, qcl)):
        d, lab, nd = engine.query(pr, q_pk, kmers, rand_table=tab, boundary=(2, 0.02, 0.2, 1.0, 1.0))
        exp, lab_o, nd_o = oracle.query(ref, q_np, kmers, tab, cl, q_cl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
        torch.cuda.synchronize()
        assert np.abs(d.cpu().numpy() - exp).max() <= 1e-6 and int(nd.item()) == nd_o
        c, _, _ = engine.query(pr, q_pk, kmers, out_mode=engine.OUT_COUNTS)
        assert (c.cpu().numpy().view(np.uint32) == oracle.query(ref, q_np, kmers, out_mode=oracle.OUT_COUNTS)[0]).all()
    oi, oj, ne, _ = engine.query_edges(pr, None, kmers, (2, 0.02, 0.2, 1.0, 1.0), rand_table=tab)
    h, _, _ = engine.query_host(ref, None, kmers, tab, cl)
    assert np.abs(h - oracle.query(ref, None, kmers, tab, cl)[0]).max() <= 1e-6
# sketches of more than 65535 bins: the uint32 count tile and the generic (no y-table) epilogue
big = synth.synth_sketches(6, kmers[:2], 1100, seed=3, n_lineages=2, chunk=8)
c, _, _ = engine.query(engine.pack(big), None, kmers[:2], out_mode=engine.OUT_COUNTS)
assert (c.cpu().numpy().view(np.uint32) == oracle.query(big, None, kmers[:2], out_mode=oracle.OUT_COUNTS)[0]).all()
h, _, _ = engine.query_host(big, None, kmers[:2])
assert np.abs(h - oracle.query(big, None, kmers[:2])[0]).max() <= 1e-6
# the host call on every visible device (peer-to-peer scatter of the packed sketches when there are several)
os.environ["PPB_MIN_ROWS_PER_DEVICE"] = "1000"
ref = synth.synth_sketches(300, kmers, 16, seed=3, n_roots=2)
h, _, _ = engine.query_host(ref, None, kmers, devices=engine.visible_devices(0), out=np.empty((300 * 299 // 2, 2), dtype=np.float32))
assert np.abs(h - oracle.query(ref, None, kmers)[0]).max() <= 1e-6
rng = np.random.default_rng(1)
n = 90
d = np.round(rng.random((n * (n - 1) // 2, 2)) * 0.6, 2).astype(np.float32)
lists = lambda t: tuple(np.asarray(a).tolist() for a in t)
assert (refine.assignThreshold(d, 2, 0.3, 0.2) == oracle.assign_threshold(d, 2, 0.3, 0.2)).all()
assert refine.edgeThreshold(d, 2, 0.3, 0.2) == list(zip(*lists(oracle.edge_iterate(d, 2, 0.3, 0.2))))
offs = np.linspace(-0.05, 0.6, 17)
assert refine.thresholdIterate1D(d, offs, 2, 0.02, 0.03, 0.5, 0.45) == lists(oracle.threshold_iterate_1d(d, offs, 2, 0.02, 0.03, 0.5, 0.45))
for slope, o2 in ((2, np.linspace(0.01, 0.5, 12)), (0, offs[:5])):      # bisection over the boundaries; few rows admitted
    assert refine.thresholdIterate1D(d, o2, slope, 0.02, 0.03, 0.5, 0.45) == lists(oracle.threshold_iterate_1d(d, o2, slope, 0.02, 0.03, 0.5, 0.45))
os.environ["PPB_ITERATE1D_FULL"] = "1"
assert refine.thresholdIterate1D(d, offs, 2, 0.02, 0.03, 0.5, 0.45) == lists(oracle.threshold_iterate_1d(d, offs, 2, 0.02, 0.03, 0.5, 0.45))
del os.environ["PPB_ITERATE1D_FULL"]
lab = oracle.assign_threshold(d, 2, 0.3, 0.2).astype(np.int8)
assert refine.generateTuples(lab, -1) == list(zip(*lists(oracle.generate_tuples(lab.astype(np.int32), -1))))
xm = np.linspace(0.05, 0.7, 9).astype(np.float32)
assert refine.thresholdIterate2D(d, xm, 0.4) == lists(oracle.threshold_iterate_2d(d, xm, 0.4))
assert refine.generateAllTuples(n, 0, True, 1) == list(zip(*lists(oracle.generate_all_tuples(n, 0, True, 1))))
sq = reshape.longToSquare(d[:, [1]])
assert (sq == oracle.long_to_square(d[:, 1], n)).all()
knn = refine.get_kNN_distances(np.ascontiguousarray(sq), 7)
assert knn == lists(oracle.get_knn_distances(sq, 7))
assert refine.get_kNN_distances(np.ascontiguousarray(sq), 40) == lists(oracle.get_knn_distances(sq, 40))   # radix-select kernel
assert refine.lowerRank(knn, n, 3, True, True, 0.05) == lists(oracle.lower_rank(*knn, n, 3, True, True, 0.05))
qr = np.round(rng.random((n, 11)), 2).astype(np.float32)
qq = np.round(rng.random((11, 11)), 2).astype(np.float32)
assert refine.extend(knn, qq, qr, 5) == lists(oracle.extend(*knn, qq, qr, 5))
torch.cuda.synchronize()
print("sanitize_small: all kernels ran, all results match the oracle")
