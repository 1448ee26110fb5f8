"""CPU tests that PIN THE ORACLE (no GPU): against the reference's own fixtures (tests/golden/, made by
tests/golden/make_golden.py from /root/reference) and against an independent NumPy restatement."""
import os

import numpy as np
import pytest

from poppunk_b200 import synth

KMERS = np.array([15, 19, 23, 27, 31], dtype=np.int32)


# ---------------------------------------------------------------- index maps (src/boundary.cpp:18-37)
@pytest.mark.parametrize("n", [2, 3, 4, 7, 64, 129, 1001])
def test_index_maps_roundtrip(oracle, n):
    k = 0
    for i in range(n - 1):
        for j in (i + 1, (i + 1 + n - 1) // 2, n - 1):  # first / middle / last column of the row
            kk = oracle.square_to_condensed(i, j, n)
            assert oracle.calc_row_idx(kk, n) == i
            assert oracle.calc_col_idx(kk, i, n) == j
        assert oracle.square_to_condensed(i, i + 1, n) == k  # rows are contiguous, row-major
        k += n - 1 - i
    assert k == n * (n - 1) // 2


def test_index_maps_large_n(oracle):
    n = 100_000  # BASELINE north-star size: n_pairs > 2^32
    total = n * (n - 1) // 2
    assert total == 4_999_950_000
    for i in (0, 1, 2, 31_337, 49_999, 50_000, 99_997, 99_998):
        for j in (i + 1, n - 1):
            kk = oracle.square_to_condensed(i, j, n)
            assert 0 <= kk < total
            assert oracle.calc_row_idx(kk, n) == i
            assert oracle.calc_col_idx(kk, i, n) == j
    assert oracle.square_to_condensed(n - 2, n - 1, n) == total - 1


def test_row_order_matches_iterDistRows(oracle):
    """utils.py:199-226: self -> for i: for j>i; non-self -> for query: for ref."""
    n = 9
    rows = [(i, j) for i in range(n) for j in range(i + 1, n)]
    for k, (i, j) in enumerate(rows):
        assert oracle.square_to_condensed(i, j, n) == k
    i_arr, j_arr = oracle.pair_rows(n)
    assert list(zip(i_arr.tolist(), j_arr.tolist())) == rows
    q, r = oracle.pair_rows(4, 3)  # 4 refs, 3 queries: row = q*R + r
    assert list(zip(q.tolist(), r.tolist())) == [(qq, rr) for qq in range(3) for rr in range(4)]


# ---------------------------------------------------------------- golden: the reference's real sketch
def test_json_sketch_schema_and_self_distance(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "json_sketch.npz"))
    ss64, bbits = int(g["sketchsize64"]), int(g["bbits"])
    sk = g["sketch"]  # [6 kmers][2184]
    assert bbits == 14 and ss64 == 156 and sk.shape == (6, ss64 * bbits)  # D1: W = sketchsize64*bbits
    kmers = g["kmers"]
    two = np.stack([sk, sk])  # the same genome twice -> every bin agrees at every k
    cnt, _ = oracle.query(two, None, kmers, out_mode=oracle.OUT_COUNTS)
    assert (cnt == 64 * ss64).all()
    d, ndeg = oracle.query(two, None, kmers)
    assert ndeg == 0 and (d == 0).all()  # J = 1 at every k -> (0, 0)
    # sketches of the same genome at different k share no hashes: matches ~ S / 2^14 (b-bit collisions)
    cross = np.stack([sk[[0, 1, 2]], sk[[3, 4, 5]]])
    cnt, _ = oracle.query(cross, None, kmers[:3], out_mode=oracle.OUT_COUNTS)
    assert cnt.max() <= 8
    # the signatures un-slice to uniformly distributed 14-bit values (layout sanity)
    sig = synth.unslice(sk, ss64)
    assert sig.max() < (1 << 14) and sig.max() > (1 << 14) - 64
    assert abs(sig.mean() - (1 << 13)) < 200


# ---------------------------------------------------------------- a4: counts
@pytest.mark.parametrize("ss64", [1, 2, 16, 17])
def test_counts_vs_unsliced_equality(oracle, ss64):
    """Bit-sliced popcount == direct equality count of un-sliced signatures (b-bit MinHash definition)."""
    ref = synth.synth_sketches(23, KMERS, ss64, seed=3, n_lineages=2)
    cnt, _ = oracle.query(ref, None, KMERS, out_mode=oracle.OUT_COUNTS)
    assert (cnt == oracle.counts_numpy(ref)).all()
    qry = synth.synth_sketches(5, KMERS, ss64, seed=4, n_lineages=2)
    cnt, _ = oracle.query(ref, qry, KMERS, out_mode=oracle.OUT_COUNTS)
    assert cnt.shape == (5 * 23, 5)
    assert (cnt == oracle.counts_numpy(ref, qry)).all()


def test_counts_adversarial(oracle):
    ss64, K = 16, 5
    S = 64 * ss64
    rng = np.random.default_rng(0)
    base = rng.integers(0, 1 << 14, size=(K, S), dtype=np.uint16)
    g = [base.copy() for _ in range(6)]
    g[1] = base ^ np.uint16(0x3FFF)          # every bit differs
    g[2] = base.copy(); g[2][:, 0] ^= 1      # first bin, lowest bit
    g[3] = base.copy(); g[3][:, 63] ^= 1 << 13   # last bin of word 0, highest plane
    g[4] = base.copy(); g[4][:, 64] ^= 1 << 7    # first bin of word 1
    g[5] = base.copy(); g[5][:, S - 1] ^= 0x2AAA  # last bin
    sk = synth.bitslice(np.stack(g))
    cnt, _ = oracle.query(sk, None, KMERS, out_mode=oracle.OUT_COUNTS)
    i, j = oracle.pair_rows(6)
    row = {(a, b): r for r, (a, b) in enumerate(zip(i.tolist(), j.tolist()))}
    assert (cnt[row[(0, 1)]] == 0).all()
    for other in (2, 3, 4, 5):
        assert (cnt[row[(0, other)]] == S - 1).all()
    assert (cnt[row[(2, 3)]] == S - 2).all()
    assert (cnt == oracle.counts_numpy(sk)).all()


# ---------------------------------------------------------------- a6: regression
def test_regression_vs_fitKmerCurve_golden(oracle, golden_dir):
    """Golden vectors from the reference's own restatement of the fit (sketchlib.py:635-670)."""
    g = np.load(os.path.join(golden_dir, "fit_kmer_curve.npz"))
    inside = bound = near = 0
    for r in range(len(g["n_k"])):
        n = int(g["n_k"][r])
        out, deg = oracle.regress_rows(g["jaccard"][r:r + 1, :n], g["klist"][r, :n], 1024)
        assert deg == 0
        exp = g["expected"][r]
        if exp.min() > 0.003:
            # unconstrained optimum well inside the bounds: scipy == OLS, the 1e-6 parity bar
            assert np.abs(out[0] - exp).max() < 1e-6
            inside += 1
        elif exp.min() < 1e-5:
            # bound active: scipy's trust-region stops a hair inside it, the library clamps to exactly 0.
            # (the OTHER coordinate is re-fitted by scipy with the bound active — not OLS — so it is not compared)
            assert (out[0][exp < 1e-5] == 0.0).all()
            bound += 1
        else:
            # close to a bound scipy's default ftol/xtol leave ~1e-5 of slack: loose comparison only
            assert np.abs(out[0] - exp).max() < 1e-4
            near += 1
    assert inside > 100 and bound > 20 and near > 20


def test_regression_vs_lstsq_and_truncation(oracle):
    rng = np.random.default_rng(5)
    S = 1024
    jac = rng.uniform(0.02, 1.0, size=(300, 5))
    jac[:, ::-1].sort(axis=1)                      # decreasing in k, as real data
    jac[10:40, 4] = 4.0 / S                        # truncate last
    jac[40:60, 2] = 0.0                            # truncate from the middle: n = 2
    jac[60:70, 1] = 1.0 / S                        # n = 1 -> degenerate
    jac[70:75, 0] = 0.0                            # n = 0 -> degenerate
    jac[75:80] = 1.0                               # identical
    jac[80:90, :] = jac[80:90, ::-1]               # increasing with k -> slope > 0 -> core clamps to 0
    out, deg = oracle.regress_rows(jac, KMERS, S)
    exp, deg_np = oracle.regress_numpy(jac, KMERS, S)
    assert deg == deg_np == 15
    assert np.abs(out - exp).max() < 1e-6
    assert (out[60:75] == 0).all() and (out[75:80] == 0).all()
    assert (out[80:90, 0] == 0).all()
    # 5/S exactly is kept, just below is dropped (J_k < 5/S)
    edge = np.array([[0.9, 0.8, 0.7, 0.6, 5.0 / S], [0.9, 0.8, 0.7, 0.6, np.nextafter(5.0 / S, 0)]])
    o, _ = oracle.regress_rows(edge, KMERS, S)
    e4, _ = oracle.regress_numpy(edge[:1], KMERS, S)
    e3, _ = oracle.regress_numpy(edge[:1, :4], KMERS[:4], S)
    assert np.abs(o[0] - e4[0]).max() < 1e-6 and np.abs(o[1] - e3[0]).max() < 1e-6


def test_full_path_vs_numpy(oracle):
    """counts -> random correction -> regression: C oracle vs the NumPy chain, self and non-self."""
    ss64 = 16
    S = 64 * ss64
    ref = synth.synth_sketches(40, KMERS, ss64, seed=1, n_lineages=3)
    qry = synth.synth_sketches(7, KMERS, ss64, seed=1, n_lineages=3, sample_seed=1)
    table = synth.random_match_table(KMERS, 3)
    rc, qc = synth.synth_clusters(40, 3, seed=1), synth.synth_clusters(7, 3, seed=2)
    for q, qcl in ((None, None), (qry, qc)):
        for tab in (None, table):
            d, ndeg = oracle.query(ref, q, KMERS, tab, rc, qcl)
            cnt = oracle.counts_numpy(ref, q)
            jac = oracle.jaccard_numpy(cnt, S, tab, rc, qcl, 40, None if q is None else 7)
            exp, ndeg_np = oracle.regress_numpy(jac, KMERS, S)
            assert ndeg == ndeg_np
            assert np.abs(d - exp).max() < 1e-6
            j32, _ = oracle.query(ref, q, KMERS, tab, rc, qcl, out_mode=oracle.OUT_JACCARD)
            assert np.abs(j32 - jac).max() < 1e-7
            assert (d[:, 0] > 0).mean() > 0.5     # the synthetic data is not degenerate


def test_recovers_planted_distances(oracle):
    """Sketches built with c_k = round(S (1-a)(1-pi)^k) give back (pi, a) to sketch resolution."""
    ss64 = 256
    S = 64 * ss64
    rng = np.random.default_rng(9)
    pi, a = 0.01, 0.15
    base = rng.integers(0, 1 << 14, size=(5, S), dtype=np.uint16)
    other = base.copy()
    for t, k in enumerate(KMERS):
        keep = int(round(S * (1 - a) * (1 - pi) ** k))
        other[t, keep:] ^= np.uint16(1)           # exactly `keep` bins agree
    sk = synth.bitslice(np.stack([base, other]))
    cnt, _ = oracle.query(sk, None, KMERS, out_mode=oracle.OUT_COUNTS)
    assert cnt[0].tolist() == [int(round(S * (1 - a) * (1 - pi) ** k)) for k in KMERS]
    d, _ = oracle.query(sk, None, KMERS)
    assert abs(d[0, 0] - pi) < 2e-4 and abs(d[0, 1] - a) < 2e-3


def test_row_range_shards(oracle):
    ref = synth.synth_sketches(30, KMERS, 4, seed=6)
    full, nd = oracle.query(ref, None, KMERS)
    parts, nds = [], 0
    for b, e in ((0, 7), (7, 200), (200, 435)):
        p, n = oracle.query(ref, None, KMERS, row_begin=b, row_end=e)
        parts.append(p)
        nds += n
    assert (np.concatenate(parts) == full).all() and nds == nd
    one, _ = oracle.query(ref, None, KMERS, threads=1)
    assert (one == full).all()


# ---------------------------------------------------------------- a7: assign_threshold
def test_assign_threshold_golden_grid(oracle, golden_dir):
    """test/test-refine.py:46-61, exact equality."""
    g = np.load(os.path.join(golden_dir, "refine_grid.npz"))
    for slope in (0, 1, 2):
        res = oracle.assign_threshold(g["dist"], slope, 0.5, 0.5, threads=2)
        assert (res == g["labels"][slope]).all()
        res = oracle.assign_threshold(g["cloud"], slope, 0.5, 0.5, threads=2)
        assert (res == g["cloud_labels"][slope]).all()


def test_assign_threshold_vs_numpy(oracle):
    rng = np.random.default_rng(2)
    d = rng.random((5000, 2)).astype(np.float32)
    d[:50] = 0
    for slope, xm, ym in ((2, 0.3, 0.7), (2, 0.0, 0.5), (2, 0.4, 0.0), (0, 0.25, 0.0), (1, 0.0, 0.6)):
        assert (oracle.assign_threshold(d, slope, xm, ym) == oracle.assign_threshold_numpy(d, slope, xm, ym)).all()
    # fused labels in the query path = assign_threshold(X / scale)
    ref = synth.synth_sketches(25, KMERS, 16, seed=8)
    bnd = (2, 0.02, 0.2, 0.9, 0.8)
    dist, labels, _ = oracle.query(ref, None, KMERS, boundary=bnd)
    scaled = (dist / np.array([0.9, 0.8], dtype=np.float32)).astype(np.float32)
    assert (labels == oracle.assign_threshold(scaled, 2, 0.02, 0.2).astype(np.int8)).all()
    assert set(np.unique(labels)) <= {-1, 0, 1} and (labels == -1).any() and (labels == 1).any()


# ---------------------------------------------------------------- "next" rows N1 / N2 (oracle pinned on CPU)
def test_edges_oracle_vs_reference_loop(oracle, golden_dir):
    """Golden edge lists from the reference test's own Python loop (test/test-refine.py:30-38, 64-82)."""
    g = np.load(os.path.join(golden_dir, "refine_grid.npz"))
    for slope in (0, 1, 2):
        exp = g[f"cloud_edges_{slope}"]
        oi, oj = oracle.generate_tuples(g["cloud_labels"][slope].astype(np.int32), -1)
        assert (np.stack([oi, oj], axis=1) == exp).all()
        ei, ej = oracle.edge_iterate(g["cloud"], slope, 0.5, 0.5)       # <= 0: adds rows labelled 0
        assert len(ei) == int((g["cloud_labels"][slope] <= 0).sum())
        assert set(map(tuple, exp.tolist())) <= set(zip(ei.tolist(), ej.tolist()))
    lab = np.array([-1, 0, -1, 1, -1, -1], dtype=np.int32)             # 2 refs x 3 queries, row = q*R + r
    oi, oj = oracle.generate_tuples(lab, -1, self=False, num_ref=2, int_offset=10)
    assert list(zip(oi.tolist(), oj.tolist())) == [(10, 12), (10, 13), (10, 14), (11, 14)]


def test_long_square_oracle(oracle):
    rng = np.random.default_rng(1)
    n = 31
    v = rng.random(n * (n - 1) // 2).astype(np.float32)
    sq = oracle.long_to_square(v, n)
    i, j = np.triu_indices(n, k=1)
    assert (sq[i, j] == v).all() and (sq == sq.T).all() and (np.diag(sq) == 0).all()
    assert (oracle.square_to_long(sq) == v).all()
    R, Q = 6, 4
    rr, qr, qq = (rng.random(m).astype(np.float32) for m in (15, 24, 6))
    m = oracle.long_to_square_multi(rr, qr, qq, R, Q)
    assert (m[:R, :R] == oracle.long_to_square(rr, R)).all() and (m[R:, R:] == oracle.long_to_square(qq, Q)).all()
    assert (m[R:, :R] == qr.reshape(Q, R)).all() and (m == m.T).all()


# ---------------------------------------------------------------- real genomes (BASELINE config 1 input)
def test_example_set_fixture_oracle_vs_numpy(oracle, golden_dir):
    """The stand-in sketches of the reference's 29 smoke-test assemblies (tests/golden/example_set_sketches.npz,
    tools/make_cfg1_fixture.py): the C oracle and the NumPy restatement agree on counts and on (core, acc), and the
    Jaccard spectrum is that of real genomes (truncation is exercised, nothing is degenerate)."""
    z = np.load(os.path.join(golden_dir, "example_set_sketches.npz"))
    db_k, sk = z["kmers"], z["sketches"]
    assert int(z["sketchsize64"]) == 16 and int(z["bbits"]) == 14 and sk.dtype == np.uint64
    klist = np.arange(13, 30, 4).astype(np.int32)
    ref = np.ascontiguousarray(sk[:, [int(np.where(db_k == k)[0][0]) for k in klist]])
    cnt, _ = oracle.query(ref, None, klist, out_mode=oracle.OUT_COUNTS)
    assert (cnt == oracle.counts_numpy(ref)).all()
    d, ndeg = oracle.query(ref, None, klist)
    exp, ndeg_np = oracle.regress_numpy(cnt.astype(np.float64) / 1024.0, klist, 1024)
    assert ndeg == ndeg_np == 0 and np.abs(d - exp).max() <= 1e-6
    assert (cnt[:, -1] < 5).any() and (cnt[:, 0] >= 5).all()          # some series end early, none is empty
    same = (cnt == 1024).all(axis=1)                                    # the set holds one assembly twice
    assert same.sum() == 1 and (d[same] == 0).all() and 0.02 < np.median(d[:, 0]) < 0.06
