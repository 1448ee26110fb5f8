"""CPU tests that pin the oracle's restatement of poppunk_refine (src/boundary.cpp, src/extend.cpp):

* against tests/golden/refine_ref.npz — outputs of the reference's OWN sources compiled into oracle/_ref
  (tests/golden/make_golden.py::refine_ref), always;
* against oracle/_ref/libpprefine_ref.so itself on fresh random inputs, when that build is present (it is built
  here by __graft_entry__.build(); the prebuilt .so travels to the GPU box, /root/reference does not).
"""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "refine_ref.npz"))


def golden_cases(gold, prefix):
    """{argument-string: [arrays]} of every golden entry `prefix/<args>.<t>`."""
    cases = {}
    for key in gold.files:
        if key.startswith(prefix + "/") or key == prefix + ".0" or key.startswith(prefix + "."):
            name, t = key.rsplit(".", 1)
            cases.setdefault(name[len(prefix):].lstrip("/"), {})[int(t)] = gold[key]
    return {k: [v[t] for t in sorted(v)] for k, v in cases.items()}


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and (x == y).all()


def run_case(impl, gold, family, args):
    """Call `family` of `impl` (the oracle module or oracle.ref) with the arguments a golden key encodes."""
    d, a = gold["dists"], args.split("/") if args else []
    if family == "assign_threshold":
        return [impl.assign_threshold(d, int(a[0]), float(a[1]), float(a[2]))]
    if family == "edge_iterate":
        return impl.edge_iterate(d, int(a[0]), float(a[1]), float(a[2]))
    if family == "generate_tuples":
        return impl.generate_tuples(gold["labels"], -1, bool(int(a[0])), int(a[1]), int(a[2]))
    if family == "generate_all_tuples":
        return impl.generate_all_tuples(int(a[0]), int(a[1]), bool(int(a[2])), int(a[3]))
    if family == "threshold_iterate_1d":
        return impl.threshold_iterate_1d(d, gold["offsets"], int(a[0]), 0.05, 0.05, 0.4, 0.45)
    if family == "threshold_iterate_2d":
        return impl.threshold_iterate_2d(d, gold["x_max_range"], 0.3)
    if family == "knn":
        return impl.get_knn_distances(gold[a[0]], int(a[1]))
    ci, cj, cd = (gold[f"knn/square/39.{t}"].reshape(40, 39)[:, :10].reshape(-1) for t in range(3))
    if family == "lower_rank":
        return impl.lower_rank(ci, cj, cd, 40, int(a[2]), bool(int(a[0])), bool(int(a[1])), 0.05)
    if family == "extend":
        return impl.extend(ci, cj, cd, gold["qq"], gold["qr"], int(a[0]))
    raise KeyError(family)


FAMILIES = ["assign_threshold", "edge_iterate", "generate_tuples", "generate_all_tuples", "threshold_iterate_1d",
            "threshold_iterate_2d", "knn", "lower_rank", "extend"]


@pytest.mark.parametrize("family", FAMILIES)
def test_restatement_matches_reference_golden(oracle, gold, family):
    cases = golden_cases(gold, family)
    assert cases
    for args, expected in cases.items():
        same(run_case(oracle, gold, family, args), expected)


def test_knn_prefix_property(gold):
    """the first 10 neighbours of the k=39 golden are the k=10 result used as the sparse input above"""
    j39 = gold["knn/square/39.1"].reshape(40, 39)
    assert all(r not in j39[r] for r in range(40))            # never the sample itself
    d39 = gold["knn/square/39.2"].reshape(40, 39)
    assert (np.diff(d39, axis=1) >= 0).all()                  # ascending distances


@pytest.mark.parametrize("seed", list(range(12)))
def test_restatement_matches_reference_build(oracle, seed):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference; the driver builds it via __graft_entry__.build())")
    R = oracle.ref
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 90))
    rows = n * (n - 1) // 2
    d = np.round(rng.random((rows, 2)) * 0.6, 2).astype(np.float32)
    for slope in (0, 1, 2):
        xm, ym = float(rng.random() * 0.5), float(rng.random() * 0.5)
        assert (oracle.assign_threshold(d, slope, xm, ym) == R.assign_threshold(d, slope, xm, ym)).all()
        same(oracle.edge_iterate(d, slope, xm, ym), R.edge_iterate(d, slope, xm, ym))
        offs = np.sort(rng.random(int(rng.integers(1, 30))) * 0.8 - 0.1)
        p = [float(v) for v in rng.random(4) * 0.5]
        p[2] += 0.01
        p[3] += 0.01
        same(oracle.threshold_iterate_1d(d, offs, slope, *p), R.threshold_iterate_1d(d, offs, slope, *p))
    xmr = np.sort(rng.random(int(rng.integers(1, 12))).astype(np.float32))
    same(oracle.threshold_iterate_2d(d, xmr, 0.35), R.threshold_iterate_2d(d, xmr, 0.35))
    lab = rng.integers(-1, 2, rows).astype(np.int32)
    same(oracle.generate_tuples(lab, -1, True, 0, 2), R.generate_tuples(lab, -1, True, 0, 2))
    same(oracle.generate_tuples(lab, 1, False, 7, 1), R.generate_tuples(lab, 1, False, 7, 1))
    same(oracle.generate_all_tuples(n, 0, True, 3), R.generate_all_tuples(n, 0, True, 3))
    same(oracle.generate_all_tuples(n, 5, False, 0), R.generate_all_tuples(n, 5, False, 0))
    sq = np.round(rng.random((n, n)), 1).astype(np.float32)
    k = int(rng.integers(1, n))
    same(oracle.get_knn_distances(sq, k), R.get_knn_distances(sq, k))
    ci, cj, cd = R.get_knn_distances(sq, k)
    for rec in (False, True):
        for cu in (False, True):
            kk = int(rng.integers(1, k + 1))
            same(oracle.lower_rank(ci, cj, cd, n, kk, rec, cu, 0.1), R.lower_rank(ci, cj, cd, n, kk, rec, cu, 0.1))
    nq = int(rng.integers(1, 12))
    qr = np.round(rng.random((n, nq)), 1).astype(np.float32)
    qq = np.round(rng.random((nq, nq)), 1).astype(np.float32)
    kk = int(rng.integers(1, min(k, nq + 1) + 1))
    same(oracle.extend(ci, cj, cd, qq, qr, kk), R.extend(ci, cj, cd, qq, qr, kk))
