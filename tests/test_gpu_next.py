"""GPU tests of the "next" rows: N1 edge lists, N2 long<->square reshapes."""
import os

import numpy as np
import pytest

from poppunk_b200 import synth

pytestmark = pytest.mark.gpu
KMERS = np.array([15, 19, 23, 27, 31], dtype=np.int32)


@pytest.fixture(scope="module")
def eng():
    import torch
    from poppunk_b200 import engine
    assert torch.cuda.is_available()
    return engine


def test_edges_golden_test_refine(eng, golden_dir):
    """test/test-refine.py:64-82: generateTuples / edgeThreshold vs the reference's own Python loop (golden)."""
    from poppunk_b200 import refine
    g = np.load(os.path.join(golden_dir, "refine_grid.npz"))
    for slope in (0, 1, 2):
        exp = [tuple(x) for x in g[f"cloud_edges_{slope}"].tolist()]
        labels = g["cloud_labels"][slope]
        assert refine.generateTuples([int(x) for x in labels], -1) == exp
        assert refine.generateTuples(labels.astype(np.int8), -1) == exp
        assert refine.generateTuples(labels.astype(np.float32), -1) == exp
        # edge_iterate keeps rows with line_dist <= 0: a superset that adds the rows labelled 0
        got = refine.edgeThreshold(g["cloud"], slope, 0.5, 0.5)
        assert set(exp) <= set(got) and len(got) == int((labels <= 0).sum())
        assert got == sorted(got)                      # row order == lexicographic (i, j) in self mode


def test_edges_vs_oracle_large_and_rect(eng, oracle):
    from poppunk_b200 import refine
    rng = np.random.default_rng(3)
    n = 1500
    d = rng.random((n * (n - 1) // 2, 2)).astype(np.float32)
    for slope, xm, ym in ((2, 0.3, 0.2), (0, 0.05, 0.0), (1, 0.0, 0.9)):
        oi, oj = oracle.edge_iterate(d, slope, xm, ym)
        assert refine.edgeThreshold(d, slope, xm, ym) == list(zip(oi.tolist(), oj.tolist()))
    lab = rng.integers(-1, 2, size=37 * 211).astype(np.int32)
    oi, oj = oracle.generate_tuples(lab, -1, self=False, num_ref=211, int_offset=5)
    assert refine.generateTuples(lab, -1, self=False, num_ref=211, int_offset=5) == list(zip(oi.tolist(), oj.tolist()))
    assert refine.generateTuples(np.zeros(0, dtype=np.int32), -1) == []
    assert refine.generateTuples(np.ones(10, dtype=np.int32), -1) == []


def test_fused_query_edges(eng, oracle):
    """distances -> threshold -> edges in one kernel == oracle distances + assign_threshold + generate_tuples."""
    import torch
    ref, qry = synth.synth_sketches(300, KMERS, 16, seed=1), synth.synth_sketches(90, KMERS, 16, seed=1, sample_seed=1)
    pr, pq = eng.pack(ref), eng.pack(qry)
    bnd = (2, 0.02, 0.2, 1.0, 1.0)
    for q_np, q_pk in ((None, None), (qry, pq)):
        d_o, lab_o, _ = oracle.query(ref, q_np, KMERS, boundary=bnd)
        oi, oj = oracle.generate_tuples(lab_o.astype(np.int32), -1, self=q_np is None, num_ref=300)
        gi, gj, n, _ = eng.query_edges(pr, q_pk, KMERS, bnd)
        torch.cuda.synchronize()
        assert n == len(oi) and n > 100
        assert (gi.cpu().numpy() == oi).all() and (gj.cpu().numpy() == oj).all()
    # row shards concatenate to the same list; <= 0 variant is a superset
    total = eng.num_rows(300)
    parts = [eng.query_edges(pr, None, KMERS, bnd, row_begin=b, row_end=e) for b, e in ((0, 20000), (20000, total))]
    oi, oj = oracle.generate_tuples(oracle.query(ref, None, KMERS, boundary=bnd)[1].astype(np.int32), -1)
    assert (torch.cat([p[0] for p in parts]).cpu().numpy() == oi).all()
    _, _, n_le, _ = eng.query_edges(pr, None, KMERS, bnd, include_boundary=True)
    assert n_le >= len(oi)
    # capacity overflow is reported, not written past
    gi, gj, n, _ = eng.query_edges(pr, None, KMERS, bnd, capacity=10)
    assert n == len(oi) and gi.numel() == 10


def test_long_square_roundtrip(eng, oracle):
    from poppunk_b200 import reshape
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 400):
        v = rng.random(n * (n - 1) // 2).astype(np.float32)
        sq = reshape.longToSquare(distVec=v.reshape(-1, 1), num_threads=2)
        assert (sq == oracle.long_to_square(v, n)).all()
        assert (sq == sq.T).all() and (np.diag(sq) == 0).all()
        assert (reshape.squareToLong(sq, 2) == v).all()
    R, Q = 23, 9
    rr, qr, qq = (rng.random(m).astype(np.float32) for m in (R * (R - 1) // 2, R * Q, Q * (Q - 1) // 2))
    m = reshape.longToSquareMulti(distVec=rr.reshape(-1, 1), query_ref_distVec=qr.reshape(-1, 1),
                                  query_query_distVec=qq.reshape(-1, 1), num_threads=1)
    assert (m == oracle.long_to_square_multi(rr, qr, qq, R, Q)).all()
    assert m[R + 2, 5] == qr[2 * R + 5] and m[5, R + 2] == qr[2 * R + 5]     # row = q*R + r
    # a column of the engine's own (n_pairs, 2) output, as PopPUNK passes it (utils.py:393-396)
    ref = synth.synth_sketches(40, KMERS, 4, seed=2)
    d, _, _ = eng.query_host(ref, None, KMERS)
    core = reshape.longToSquare(distVec=d[:, [0]], num_threads=1)
    i, j = np.triu_indices(40, k=1)
    assert (core[i, j] == d[:, 0]).all()


# ------------------------------------------------------------------------------------------------------------
# N1 (rest) and N3: against the golden vectors of the reference's own compiled sources (tests/golden/refine_ref.npz,
# made from oracle/_ref) and against the oracle restatement on larger seeded inputs
# ------------------------------------------------------------------------------------------------------------
def _lists(arrays):
    return tuple(np.asarray(a).tolist() for a in arrays)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "refine_ref.npz"))


def _gold(gold, name, n):
    return tuple(gold[f"{name}.{t}"].tolist() for t in range(n))


def test_refine_rest_matches_reference_golden(eng, gold):
    from poppunk_b200 import refine
    d = gold["dists"]
    for slope in (0, 1, 2):
        got = refine.thresholdIterate1D(d, gold["offsets"], slope, 0.05, 0.05, 0.4, 0.45)
        assert got == _gold(gold, f"threshold_iterate_1d/{slope}", 3)
    assert refine.thresholdIterate2D(d, gold["x_max_range"], 0.3) == _gold(gold, "threshold_iterate_2d", 3)
    for nr, nq, self_, off in [(17, 0, 1, 0), (17, 0, 1, 4), (5, 7, 0, 0), (5, 7, 0, 3)]:
        exp = _gold(gold, f"generate_all_tuples/{nr}/{nq}/{self_}/{off}", 2)
        assert refine.generateAllTuples(nr, nq, bool(self_), off) == list(zip(*exp))
    for k in (1, 3, 39):
        assert refine.get_kNN_distances(gold["square"], k) == _gold(gold, f"knn/square/{k}", 3)
    assert refine.get_kNN_distances(gold["rect"], 1) == _gold(gold, "knn/rect/1", 3)
    ci, cj, cd = (gold[f"knn/square/39.{t}"].reshape(40, 39)[:, :10].reshape(-1) for t in range(3))
    for rec in (0, 1):
        for cu in (0, 1):
            for k in (1, 2, 5):
                got = refine.lowerRank((ci, cj, cd), 40, k, bool(rec), bool(cu), 0.05)
                assert got == _gold(gold, f"lower_rank/{rec}/{cu}/{k}", 3), (rec, cu, k)
    for k in (1, 3, 8):
        assert refine.extend((ci, cj, cd), gold["qq"], gold["qr"], k) == _gold(gold, f"extend/{k}", 3)


def test_refine_rest_vs_oracle_larger(eng, oracle):
    """sizes that span many CTAs / scan blocks, distances quantised so that ties are everywhere"""
    from poppunk_b200 import refine
    rng = np.random.default_rng(5)
    n = 700
    rows = n * (n - 1) // 2                       # 244,650 rows: 60 compaction blocks, 239 scan blocks
    d = np.round(rng.random((rows, 2)) * 0.6, 3).astype(np.float32)
    offs = np.linspace(-0.05, 0.6, 41)
    for slope in (0, 1, 2):
        exp = _lists(oracle.threshold_iterate_1d(d, offs, slope, 0.02, 0.03, 0.5, 0.45))
        assert refine.thresholdIterate1D(d, offs, slope, 0.02, 0.03, 0.5, 0.45) == exp
        assert len(exp[0]) > 1000
    xm = np.linspace(0.01, 0.7, 37).astype(np.float32)
    assert refine.thresholdIterate2D(d, xm, 0.4) == _lists(oracle.threshold_iterate_2d(d, xm, 0.4))
    assert refine.thresholdIterate1D(d[:0], offs, 2, 0.0, 0.0, 0.5, 0.5) == ([], [], [])
    with pytest.raises(RuntimeError):
        refine.thresholdIterate1D(d, offs[::-1], 2, 0.0, 0.0, 0.5, 0.5)
    exp = oracle.generate_all_tuples(n, 0, True, 2)
    assert refine.generateAllTuples(n, 0, True, 2) == list(zip(exp[0].tolist(), exp[1].tolist()))
    # kNN on a square from the long->square kernel, and on wide rows (radix select over 3000 candidates, k up to 300)
    sq = np.round(rng.random((900, 900)), 2).astype(np.float32)
    for k in (1, 7, 300):
        assert refine.get_kNN_distances(sq, k) == _lists(oracle.get_knn_distances(sq, k))
    wide = np.round(rng.random((50, 3000)), 2).astype(np.float32)
    assert refine.get_kNN_distances(wide, 64) == _lists(oracle.get_knn_distances(wide, 64))
    few = rng.random((5, 4)).astype(np.float32)                     # fewer candidates than kNN: zero padded
    assert refine.get_kNN_distances(few, 6) == _lists(oracle.get_knn_distances(few, 6))
    ci, cj, cd = oracle.get_knn_distances(sq, 20)
    for rec in (False, True):
        for cu in (False, True):
            exp = _lists(oracle.lower_rank(ci, cj, cd, 900, 5, rec, cu, 0.02))
            assert refine.lowerRank((ci, cj, cd), 900, 5, rec, cu, 0.02) == exp
    qr = np.round(rng.random((900, 33)), 2).astype(np.float32)
    qq = np.round(rng.random((33, 33)), 2).astype(np.float32)
    for k in (1, 10, 40):
        assert refine.extend((ci, cj, cd), qq, qr, k) == _lists(oracle.extend(ci, cj, cd, qq, qr, k))


def test_fused_query_edges_without_ytable(eng, oracle, monkeypatch):
    """when the y-table is skipped (huge C*C*K*(S+1); forced here with PPB_NO_YTAB) the epilogue's generic
    branch must append edges too, not silently return none."""
    ref = synth.synth_sketches(400, KMERS, 16, seed=2)
    tab, cl = synth.random_match_table(KMERS, 3), synth.synth_clusters(400, 3)
    pr = eng.pack(ref, clusters=cl)
    bnd = (2, 0.02, 0.2, 1.0, 1.0)
    gi, gj, n, _ = eng.query_edges(pr, None, KMERS, bnd, rand_table=tab)
    monkeypatch.setenv("PPB_NO_YTAB", "1")
    si, sj, n_slow, _ = eng.query_edges(pr, None, KMERS, bnd, rand_table=tab)
    assert n_slow == n > 0 and (si == gi).all() and (sj == gj).all()
    _, lab_o, _ = oracle.query(ref, None, KMERS, tab, cl, boundary=bnd)
    assert n_slow == int((lab_o == -1).sum())


def test_empty_row_range_and_bad_cluster_ids(eng):
    """an empty shard is a no-op (no "no output buffer" error: a rank with nothing to do must still reach
    the collectives), and cluster ids outside the table are refused before they index it."""
    import torch
    ref = synth.synth_sketches(1, KMERS, 16, seed=2)
    out, _, ndeg = eng.query(eng.pack(ref), None, KMERS)               # one genome: zero pairs
    assert out.shape == (0, 2) and int(ndeg.item()) == 0
    ref = synth.synth_sketches(50, KMERS, 16, seed=2)
    pr = eng.pack(ref, clusters=np.full(50, 3, dtype=np.uint16))
    out, _, _ = eng.query(pr, None, KMERS, row_begin=100, row_end=100)
    assert out.shape == (0, 2)
    with pytest.raises(ValueError):
        eng.query(pr, None, KMERS, rand_table=synth.random_match_table(KMERS, 3))
    torch.cuda.synchronize()


def test_knn_zero_and_limits(eng):
    from poppunk_b200 import refine
    m = np.random.default_rng(0).random((20, 20)).astype(np.float32)
    assert refine.get_kNN_distances(m, 0) == ([], [], [])
    with pytest.raises(ValueError):
        refine.get_kNN_distances(m, refine.KNN_MAX + 1)


def test_ordered_stream_primitives_at_size(eng, oracle, monkeypatch):
    """The hand-written ordered primitives (csrc/ppb_sort.cuh) across thousands of chunks: single-pass compaction with
    look-back, the bucket scatter behind thresholdIterate2D (new path == generic path == oracle, with boundaries that
    repeat and rows that sit exactly on them), the radix sort behind thresholdIterate1D and behind the fused edge list."""
    import torch
    from poppunk_b200 import _lib, refine
    rng = np.random.default_rng(11)
    n = 3000
    rows = n * (n - 1) // 2                       # 4,498,500 rows: 1099 chunks of 4096
    d = np.round(rng.random((rows, 2)) * 0.6, 3).astype(np.float32)
    got = refine.edgeThreshold(d, 2, 0.21, 0.17)
    ei, ej = oracle.edge_iterate(d, 2, 0.21, 0.17)
    assert got == list(zip(ei.tolist(), ej.tolist())) and len(got) > 100_000
    lab = oracle.assign_threshold(d, 2, 0.21, 0.17).astype(np.int8)
    ti, tj = oracle.generate_tuples(lab.astype(np.int32), -1)
    assert refine.generateTuples(lab, -1) == list(zip(ti.tolist(), tj.tolist()))
    # 2D: steps that repeat (nothing new is admitted), a y intercept rows sit exactly on (quantised distances)
    xm = np.array([0.0, 0.05, 0.05, 0.1, 0.2, 0.2, 0.3, 0.45, 0.6, 0.9], dtype=np.float32)
    exp = _lists(oracle.threshold_iterate_2d(d, xm, 0.3))
    new = refine.thresholdIterate2D(d, xm, 0.3)
    monkeypatch.setenv("PPB_ITERATE2D_GENERIC", "1")
    old = refine.thresholdIterate2D(d, xm, 0.3)
    monkeypatch.delenv("PPB_ITERATE2D_GENERIC")
    assert new == exp and old == exp and len(exp[0]) > 100_000
    offs = np.linspace(-0.05, 0.5, 30)
    # (short ranges leave most rows never admitted; positive offsets let slope 2 bisect over the boundaries, negative
    #  ones force the full scan; the distances are quantised, so many rows sit exactly ON a boundary -> full scan too)
    for slope, off_set in ((2, offs), (2, np.linspace(0.0, 0.5, 30)), (0, offs[:7]), (1, offs[3:])):
        exp1 = _lists(oracle.threshold_iterate_1d(d, off_set, slope, 0.02, 0.03, 0.5, 0.45))
        assert refine.thresholdIterate1D(d, off_set, slope, 0.02, 0.03, 0.5, 0.45) == exp1      # only admitted rows sorted
        monkeypatch.setenv("PPB_ITERATE1D_FULL", "1")
        assert refine.thresholdIterate1D(d, off_set, slope, 0.02, 0.03, 0.5, 0.45) == exp1      # every row sorted
        monkeypatch.delenv("PPB_ITERATE1D_FULL")
    # kNN <= 32: the one-pass kernel and the radix-select kernel agree (ties everywhere: two decimals)
    sq = np.round(rng.random((700, 5000)), 2).astype(np.float32)
    for k in (1, 5, 32):
        one_pass = refine.get_kNN_distances(sq, k)
        monkeypatch.setenv("PPB_KNN_GENERIC", "1")
        assert refine.get_kNN_distances(sq, k) == one_pass
        monkeypatch.delenv("PPB_KNN_GENERIC")
    assert refine.get_kNN_distances(sq[:300], 5) == _lists(oracle.get_knn_distances(sq[:300], 5))
    # the sort behind engine.query_edges
    L = _lib.load()
    vals = rng.integers(0, 5_000_000_000, size=3_000_001, dtype=np.int64)
    t = torch.from_numpy(vals).cuda()
    _lib.check(L.ppb_sort_rows_dev(t.data_ptr(), t.numel(), 5_000_000_000, torch.cuda.current_stream().cuda_stream))
    assert (t.cpu().numpy() == np.sort(vals)).all()
    small = torch.from_numpy(np.array([5, 3, 3, 0, 9], dtype=np.int64)).cuda()
    _lib.check(L.ppb_sort_rows_dev(small.data_ptr(), 5, 9, torch.cuda.current_stream().cuda_stream))
    assert small.cpu().tolist() == [0, 3, 3, 5, 9]
