"""GPU tests (-m gpu) of the host-buffer entry the drop-in uses: ``ppb_query_host_multi`` (one process, every GPU,
one caller buffer), the host result pool, and the random-match handling of ``pp_queryDatabase`` on databases with and
without a ``/random`` table — all against the CPU oracle.  Bars as in test_gpu_parity.py."""
import ctypes as C
import os

import numpy as np
import pytest

from poppunk_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-6
KMERS = np.array([15, 19, 23, 27, 31], dtype=np.int32)


@pytest.fixture(scope="module")
def eng():
    import torch
    from poppunk_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return engine


def _n_dev():
    from poppunk_b200 import _lib
    return _lib.load().ppb_device_count()


def _job(n_ref=700, n_qry=None, seed=3):
    ref = synth.synth_sketches(n_ref, KMERS, 16, seed=seed, n_roots=2)
    qry = None if n_qry is None else synth.synth_sketches(n_qry, KMERS, 16, seed=seed, sample_seed=1, n_roots=2)
    tab = synth.random_match_table(KMERS, 3)
    rcl = synth.synth_clusters(n_ref, 3)
    qcl = None if n_qry is None else synth.synth_clusters(n_qry, 3, seed=7)
    return ref, qry, tab, rcl, qcl


@pytest.mark.parametrize("n_qry", [None, 333])
def test_host_call_pageable_pinned_pool_outputs_agree(eng, oracle, n_qry):
    """The three kinds of destination the host call can get: a plain np.empty (staged through the pinned ring), a
    page-locked torch buffer (direct DMA), a pool block from host_result()."""
    import torch
    ref, qry, tab, rcl, qcl = _job(700, n_qry)
    exp, _, ndeg_o = oracle.query(ref, qry, KMERS, tab, rcl, qcl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
    assert 0 < ndeg_o < exp.shape[0]                          # related and unrelated pairs in one job
    rows = exp.shape[0]
    pageable = np.full((rows, 2), -1, dtype=np.float32)
    pinned = torch.full((rows, 2), -1, dtype=torch.float32).pin_memory().numpy()
    for out in (pageable, pinned, None):
        got, _, ndeg = eng.query_host(ref, qry, KMERS, tab, rcl, qcl, out=out)
        assert ndeg == ndeg_o and np.abs(got - exp).max() <= TOL
    # a row range: written at offset 0 of the buffer handed in
    b, e = rows // 3, rows // 3 + 1000
    got, _, _ = eng.query_host(ref, qry, KMERS, tab, rcl, qcl, row_begin=b, row_end=e)
    assert got.shape == (1000, 2) and np.abs(got - exp[b:e]).max() <= TOL


def test_host_call_leaves_current_device_alone(eng):
    import torch
    ref, _, _, _, _ = _job(200)
    before = torch.cuda.current_device()
    eng.query_host(ref, None, KMERS, device_id=_n_dev() - 1)
    assert torch.cuda.current_device() == before


def test_host_call_rejects_bad_cluster_ids(eng):
    from poppunk_b200._lib import PpbError
    ref, _, tab, rcl, _ = _job(100)
    rcl = rcl.copy()
    rcl[17] = 3
    with pytest.raises((PpbError, ValueError)):
        eng.query_host(ref, None, KMERS, tab, rcl)


@pytest.mark.parametrize("n_qry,p2p", [(None, True), (None, False), (1500, True), (1500, False)])
def test_multi_device_call_matches_single(eng, oracle, monkeypatch, n_qry, p2p):
    """Every GPU of the box behind ONE call, one caller buffer; byte-identical to the one-device result and within
    tolerance of the oracle (self and query-sharded rectangle; with the peer-to-peer scatter of the packed
    reference array and with every device uploading all of it)."""
    n_dev = _n_dev()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    monkeypatch.setenv("PPB_MIN_ROWS_PER_DEVICE", "1000")
    if not p2p:
        monkeypatch.setenv("PPB_NO_P2P", "1")
    ref, qry, tab, rcl, qcl = _job(1200, n_qry)
    one, lab1, ndeg1 = eng.query_host(ref, qry, KMERS, tab, rcl, qcl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
    out = np.full(one.shape, -1, dtype=np.float32)
    many, labm, ndegm = eng.query_host(ref, qry, KMERS, tab, rcl, qcl, boundary=(2, 0.02, 0.2, 1.0, 1.0), out=out,
                                       devices=list(range(n_dev)))
    assert ndegm == ndeg1 and (many.view(np.uint32) == one.view(np.uint32)).all() and (labm == lab1).all()
    exp, _, _ = oracle.query(ref, qry, KMERS, tab, rcl, qcl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
    assert np.abs(many - exp).max() <= TOL
    cnt, _, _ = eng.query_host(ref, qry, KMERS, out_mode=eng.OUT_COUNTS, devices=list(range(n_dev)))
    cnt_o, _ = oracle.query(ref, qry, KMERS, out_mode=oracle.OUT_COUNTS)
    assert (cnt == cnt_o).all()


def test_more_launches_than_the_tile_list_cache_holds(eng, monkeypatch):
    """A call of several hundred row chunks (cfg4 through one process has 100 per device, 800 per call on 8 GPUs): the
    tile lists of all chunks are planned and uploaded once per call and lent to its launches.  Byte-identical to the
    un-chunked call, call after call, staged and direct, on every device of the box."""
    import torch
    ref, _, tab, rcl, _ = _job(1200)
    bnd = (2, 0.02, 0.2, 1.0, 1.0)
    one, lab1, nd1 = eng.query_host(ref, None, KMERS, tab, rcl, boundary=bnd)
    monkeypatch.setenv("PPB_HOST_CHUNK_ROWS", "2048")          # 352 launches: more than the 256 lists the cache keeps
    monkeypatch.setenv("PPB_MIN_ROWS_PER_DEVICE", "1000")
    devs = list(range(_n_dev()))
    for _ in range(2):
        for out in (np.full(one.shape, -1, dtype=np.float32), torch.full(one.shape, -1.0).pin_memory().numpy()):
            got, lab, nd = eng.query_host(ref, None, KMERS, tab, rcl, boundary=bnd, out=out, devices=devs)
            assert nd == nd1 and (got.view(np.uint32) == one.view(np.uint32)).all() and (lab == lab1).all()


def test_host_pool_block_is_pinned_on_second_reuse(eng, oracle):
    from poppunk_b200 import _lib
    L = _lib.load()
    L.ppb_release_workspace()
    ref, _, tab, rcl, _ = _job(600)
    exp, _ = oracle.query(ref, None, KMERS, tab, rcl)

    def stats():
        h, u, p = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        L.ppb_host_pool_stats(C.byref(h), C.byref(u), C.byref(p))
        return h.value, u.value, p.value

    first, _, _ = eng.query_host(ref, None, KMERS, tab, rcl)
    held, used, pinned = stats()
    assert used >= first.nbytes and pinned == 0                # a first call stages into fresh (huge) pages
    assert np.abs(first - exp).max() <= TOL
    del first
    assert stats()[1] == 0 and stats()[0] == held              # handed back, kept
    second, _, _ = eng.query_host(ref, None, KMERS, tab, rcl)
    assert stats() == (held, held, 0)                          # same block, touched pages, still staged
    assert np.abs(second - exp).max() <= TOL
    del second
    third, _, _ = eng.query_host(ref, None, KMERS, tab, rcl)
    assert stats() == (held, held, held)                       # second reuse: page-locked, direct DMA
    assert np.abs(third - exp).max() <= TOL
    view = third[5:10]
    del third
    assert stats()[1] == held                                  # a view keeps the block alive
    del view
    assert stats()[1] == 0
    L.ppb_release_workspace()
    assert stats() == (0, 0, 0)


def test_database_without_random_table_uses_documented_closed_form(eng, oracle, tmp_path, capfd):
    """docs/query_assignment.rst:110 + docs/sketching.rst:107-118: no /random group -> the closed-form chances from
    the genome lengths, with the reference's message; GPU == oracle given the same table."""
    from poppunk_b200 import sketchlib
    n, nq = 40, 9
    names = [f"g{i:02d}" for i in range(n + nq)]
    sk = synth.synth_sketches(n + nq, KMERS, 16, seed=5)
    rng = np.random.default_rng(0)
    lengths = rng.integers(1_900_000, 2_300_000, size=n + nq).astype(np.float64)
    rp, qp = str(tmp_path / "ref"), str(tmp_path / "qry")
    sketchlib.write_db_npz(rp, names[:n], KMERS, sk[:n], lengths=lengths[:n])
    sketchlib.write_db_npz(qp, names[n:], KMERS, sk[n:], lengths=lengths[n:])
    d = sketchlib.queryDatabase(names[:n], names[:n], rp, rp, KMERS, self=True)
    assert "Could not find random match chances in database, calculating assuming equal base frequencies" in capfd.readouterr().err
    tab, rcl, _ = sketchlib.random_match_fallback(lengths[:n], None, KMERS)
    # 40 distinct lengths -> more than 32 classes are not allowed: classes of log-length
    assert tab.shape == (32, 32, 5)
    exp, _ = oracle.query(sk[:n], None, KMERS, tab, rcl)
    assert np.abs(d - exp).max() <= TOL
    raw, _ = oracle.query(sk[:n], None, KMERS)
    assert np.abs(d - raw).max() > 1e-4                        # the correction is really applied
    # per-genome exactness of the table when lengths repeat (<= 32 distinct values)
    few = np.repeat([2.0e6, 2.2e6, 1.8e6], 20)[:n]
    tab_f, cl_f, _ = sketchlib.random_match_fallback(few, None, KMERS)
    r = -np.expm1(few[:, None] * np.log1p(-2.0 * 4.0 ** (-KMERS.astype(np.float64)[None, :])))   # 1 - (1 - 2 4^-k)^l
    jr = r[3] * r[27] / (r[3] + r[27] - r[3] * r[27])
    assert few[3] != few[27] and np.allclose(tab_f[cl_f[3], cl_f[27]], jr, rtol=1e-5)
    # query-vs-ref: lengths of both sides enter
    d = sketchlib.queryDatabase(names[:n], names[n:], rp, qp, KMERS, self=False)
    tab, rcl, qcl = sketchlib.random_match_fallback(lengths[:n], lengths[n:], KMERS)
    exp, _ = oracle.query(sk[:n], sk[n:], KMERS, tab, rcl, qcl)
    assert d.shape == (n * nq, 2) and np.abs(d - exp).max() <= TOL


def test_queries_outside_the_table_take_the_nearest_centroid(eng, oracle, tmp_path):
    from poppunk_b200 import sketchlib
    n, nq = 30, 6
    names = [f"g{i:02d}" for i in range(n + nq)]
    sk = synth.synth_sketches(n + nq, KMERS, 16, seed=6)
    tab = synth.random_match_table(KMERS, 3)
    cl = synth.synth_clusters(n, 3)
    centroids = np.array([[0.3, 0.2, 0.2, 0.3], [0.25, 0.25, 0.25, 0.25], [0.2, 0.3, 0.3, 0.2]])
    q_cl = np.array([2, 0, 1, 1, 2, 0], dtype=np.uint16)
    rng = np.random.default_rng(1)
    q_bf = centroids[q_cl] + rng.normal(0, 0.004, size=(nq, 4))
    rp, qp = str(tmp_path / "ref"), str(tmp_path / "qry")
    sketchlib.write_db_npz(rp, names[:n], KMERS, sk[:n], tab, cl, random_centroids=centroids,
                           base_freq=centroids[cl])
    sketchlib.write_db_npz(qp, names[n:], KMERS, sk[n:], base_freq=q_bf)
    d = sketchlib.queryDatabase(names[:n], names[n:], rp, qp, KMERS, self=False)
    exp, _ = oracle.query(sk[:n], sk[n:], KMERS, tab, cl, q_cl)
    assert np.abs(d - exp).max() <= TOL


def test_plot_fit_probe(eng, tmp_path):
    """--plot-fit (PopPUNK/sketchlib.py:540-573): two per-k probes of a random pair, fitKmerCurve on both."""
    from poppunk_b200 import sketchlib
    names = [f"g{i:02d}" for i in range(12)]
    sk = synth.synth_sketches(12, KMERS, 16, seed=8)
    tab, cl = synth.random_match_table(KMERS, 3), synth.synth_clusters(12, 3)
    p = str(tmp_path / "db")
    sketchlib.write_db_npz(p, names, KMERS, sk, tab, cl)
    sketchlib.queryDatabase(names, names, p, p, KMERS, self=True, number_plot_fits=2)
    for i in (1, 2):
        lines = open(os.path.join(p, f"db_fit_example_{i}.tsv")).read().splitlines()
        assert lines[3] == "k\traw\tcorrected" and len(lines) == 4 + len(KMERS)
        raw = np.array([float(l.split("\t")[1]) for l in lines[4:]])
        cor = np.array([float(l.split("\t")[2]) for l in lines[4:]])
        assert (cor <= raw + 1e-6).all() and (raw > 0).all()
