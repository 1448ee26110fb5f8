"""GPU parity tests (-m gpu): the CUDA path, called THROUGH THE C ABI (ppb_query_host / ppb_query_dev via
ctypes), against the CPU oracle on the same seeded inputs.  Bars: per-k counts bit-exact; (core, acc) within
1e-6 absolute (the tolerance BASELINE.json's north_star states); labels exact."""
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


def _sk(n, ss64, seed=1, kmers=KMERS, sample_seed=0, **kw):
    return synth.synth_sketches(n, kmers, ss64, seed=seed, sample_seed=sample_seed, **kw)


# ---------------------------------------------------------------- counts: bit-exact
@pytest.mark.parametrize("n,ss64", [(2, 16), (3, 1), (65, 2), (129, 16), (200, 17), (70, 156), (40, 256)])
def test_counts_self_bit_exact(eng, oracle, n, ss64):
    ref = _sk(n, ss64, n_lineages=3)
    got, _, _ = eng.query_host(ref, None, KMERS, out_mode=eng.OUT_COUNTS)
    exp, _ = oracle.query(ref, None, KMERS, out_mode=oracle.OUT_COUNTS)
    assert got.shape == exp.shape and (got == exp).all()


@pytest.mark.parametrize("nr,nq,ss64", [(1, 1, 16), (130, 3, 16), (7, 150, 4), (257, 70, 16), (33, 65, 156)])
def test_counts_rect_bit_exact(eng, oracle, nr, nq, ss64):
    ref, qry = _sk(nr, ss64), _sk(nq, ss64, sample_seed=1)
    got, _, _ = eng.query_host(ref, qry, KMERS, out_mode=eng.OUT_COUNTS)
    exp, _ = oracle.query(ref, qry, KMERS, out_mode=oracle.OUT_COUNTS)
    assert got.shape == (nr * nq, 5) and (got == exp).all()


@pytest.mark.parametrize("K", [1, 2, 8, 9, 16, 17, 32])
def test_counts_many_k(eng, oracle, K):
    kmers = np.arange(7, 7 + 2 * K, 2, dtype=np.int32)
    ref = _sk(150, 16, kmers=kmers)
    got, _, _ = eng.query_host(ref, None, kmers, out_mode=eng.OUT_COUNTS)
    exp, _ = oracle.query(ref, None, kmers, out_mode=oracle.OUT_COUNTS)
    assert (got == exp).all()


def test_counts_adversarial(eng, oracle):
    S = 1024
    rng = np.random.default_rng(0)
    base = rng.integers(0, 1 << 14, size=(5, S), dtype=np.uint16)
    g = [base.copy() for _ in range(8)]
    g[1] = base ^ np.uint16(0x3FFF)
    g[2][:, 0] ^= 1
    g[3][:, 63] ^= 1 << 13
    g[4][:, 64] ^= 1 << 7
    g[5][:, S - 1] ^= 0x2AAA
    g[6][:, 31] ^= 1 << 12      # last bin of a 32-bit half
    g[7][:, 32] ^= 1 << 13      # first bin of the high half, tail plane
    sk = synth.bitslice(np.stack(g))
    got, _, _ = eng.query_host(sk, None, KMERS, out_mode=eng.OUT_COUNTS)
    exp, _ = oracle.query(sk, None, KMERS, out_mode=oracle.OUT_COUNTS)
    assert (got == exp).all()
    assert (got[0] == 0).all() and (got[1:7] == S - 1).all()   # rows (0,1) and (0,2..7)


def test_json_sketch_golden(eng, golden_dir):
    g = np.load(os.path.join(golden_dir, "json_sketch.npz"))
    two = np.stack([g["sketch"], g["sketch"]])
    cnt, _, _ = eng.query_host(two, None, g["kmers"], out_mode=eng.OUT_COUNTS)
    assert (cnt == 64 * int(g["sketchsize64"])).all()
    d, _, ndeg = eng.query_host(two, None, g["kmers"])
    assert ndeg == 0 and (d == 0).all()


# ---------------------------------------------------------------- distances: <= 1e-6
@pytest.mark.parametrize("use_random", [False, True])
@pytest.mark.parametrize("n,ss64", [(300, 16), (131, 156), (64, 256)])
def test_dists_self(eng, oracle, n, ss64, use_random):
    ref = _sk(n, ss64, n_lineages=4)
    tab = synth.random_match_table(KMERS, 3) if use_random else None
    cl = synth.synth_clusters(n, 3) if use_random else None
    got, _, ndeg = eng.query_host(ref, None, KMERS, tab, cl)
    exp, ndeg_o = oracle.query(ref, None, KMERS, tab, cl)
    assert got.dtype == np.float32 and got.flags.c_contiguous and got.shape == (n * (n - 1) // 2, 2)
    assert np.abs(got - exp).max() <= TOL
    assert ndeg == ndeg_o
    assert (got[:, 0] > 0).mean() > 0.5
    jac, _, _ = eng.query_host(ref, None, KMERS, tab, cl, out_mode=eng.OUT_JACCARD)
    jac_o, _ = oracle.query(ref, None, KMERS, tab, cl, out_mode=oracle.OUT_JACCARD)
    assert np.abs(jac - jac_o).max() <= 1e-7


def test_dists_rect_random_and_degenerate(eng, oracle):
    ref = _sk(200, 16)
    qry = np.concatenate([_sk(60, 16, sample_seed=1), _sk(20, 16, seed=99)])   # last 20 queries are unrelated
    tab = synth.random_match_table(KMERS, 3)
    rc, qc = synth.synth_clusters(200, 3), synth.synth_clusters(80, 3, seed=5)
    got, _, ndeg = eng.query_host(ref, qry, KMERS, tab, rc, qc)
    exp, ndeg_o = oracle.query(ref, qry, KMERS, tab, rc, qc)
    assert np.abs(got - exp).max() <= TOL
    assert ndeg == ndeg_o and ndeg >= 20 * 200
    assert (got[60 * 200:] == 0).all()                  # unrelated pairs: fit has < 2 usable k -> (0, 0)


def test_truncation_positions(eng, oracle):
    """Pairs built so the series is cut after 2, 3, 4 k-mers and not at all."""
    S, ss64 = 1024, 16
    rng = np.random.default_rng(3)
    base = rng.integers(0, 1 << 14, size=(5, S), dtype=np.uint16)
    gen = [base]
    for cut in (2, 3, 4, 5, 1, 0):
        o = base.copy()
        for t in range(5):
            keep = int(S * 0.8 * 0.97 ** t) if t < cut else 3        # 3/S < 5/S: dropped
            o[t, keep:] ^= np.uint16(1)
        gen.append(o)
    sk = synth.bitslice(np.stack(gen))
    got, _, ndeg = eng.query_host(sk, None, KMERS)
    exp, ndeg_o = oracle.query(sk, None, KMERS)
    assert np.abs(got - exp).max() <= TOL and ndeg == ndeg_o and ndeg >= 2
    assert len({tuple(r) for r in got[:4].round(5).tolist()}) == 4    # different cuts give different fits


def test_row_range_shards_identical(eng, oracle):
    ref = _sk(333, 16)
    full, _, nd = eng.query_host(ref, None, KMERS)
    total = full.shape[0]
    cuts = [0, 1, 331, 332, 5000, total // 2, total - 1, total]
    parts, nds = [], 0
    for b, e in zip(cuts[:-1], cuts[1:]):
        p, _, n = eng.query_host(ref, None, KMERS, row_begin=b, row_end=e)
        parts.append(p)
        nds += n
    assert (np.concatenate(parts) == full).all() and nds == nd
    qry = _sk(77, 16, sample_seed=1)
    fullr, _, _ = eng.query_host(ref, qry, KMERS)
    pr = [eng.query_host(ref, qry, KMERS, row_begin=b, row_end=e)[0] for b, e in ((0, 400), (400, 20000), (20000, 333 * 77))]
    assert (np.concatenate(pr) == fullr).all()


def test_device_path_gather_and_order(eng, oracle):
    """ppb_pack_dev with an index list = caller-ordered subset (rList/qList semantics) + ppb_query_dev."""
    import torch
    ref = _sk(180, 16)
    order = np.random.default_rng(1).permutation(180)[:97]
    packed = eng.pack(ref, idx=order)
    out, _, ndeg = eng.query(packed, None, KMERS)
    torch.cuda.synchronize()
    exp, ndeg_o = oracle.query(ref[order], None, KMERS)
    assert np.abs(out.cpu().numpy() - exp).max() <= TOL and int(ndeg.item()) == ndeg_o
    q = eng.pack(_sk(50, 16, sample_seed=1))
    outr, _, _ = eng.query(packed, q, KMERS, out_mode=eng.OUT_COUNTS)
    expr, _ = oracle.query(ref[order], _sk(50, 16, sample_seed=1), KMERS, out_mode=oracle.OUT_COUNTS)
    assert (outr.cpu().numpy().view(np.uint32) == expr).all()


def test_self_equals_rect_symmetry(eng):
    """dist(i, j) from the condensed self matrix == dist from the rectangle of the same genomes."""
    ref = _sk(150, 16)
    d_self, _, _ = eng.query_host(ref, None, KMERS)
    d_rect, _, _ = eng.query_host(ref, ref.copy(), KMERS)
    i, j = np.triu_indices(150, k=1)
    assert (d_self == d_rect[i * 150 + j]).all() and (d_self == d_rect[j * 150 + i]).all()


def test_empty_and_tiny(eng):
    one = _sk(1, 16)
    out, _, nd = eng.query_host(one, None, KMERS)
    assert out.shape == (0, 2) and nd == 0
    out, _, _ = eng.query_host(_sk(5, 16), _sk(0, 16), KMERS)
    assert out.shape == (0, 2)


# ---------------------------------------------------------------- a7: assign_threshold
def test_assign_threshold_golden_grid(eng, golden_dir):
    from poppunk_b200 import refine
    g = np.load(os.path.join(golden_dir, "refine_grid.npz"))
    for slope in (0, 1, 2):
        assert (refine.assignThreshold(g["dist"], slope, 0.5, 0.5, 2) == g["labels"][slope]).all()
        assert (refine.assignThreshold(g["cloud"], slope, 0.5, 0.5, 2) == g["cloud_labels"][slope]).all()
    with pytest.raises(TypeError):
        refine.assignThreshold(g["dist"].astype(np.float64), 2, 0.5, 0.5)


def test_assign_threshold_vs_oracle(eng, oracle):
    from poppunk_b200 import refine
    rng = np.random.default_rng(2)
    d = rng.random((100_003, 2)).astype(np.float32)
    d[:50] = 0
    for slope, xm, ym in ((2, 0.3, 0.7), (2, 0.0, 0.5), (2, 0.4, 0.0), (0, 0.25, 0.0), (1, 0.0, 0.6)):
        assert (refine.assignThreshold(d, slope, xm, ym) == oracle.assign_threshold(d, slope, xm, ym, 4)).all()


def test_fused_threshold_labels(eng, oracle):
    ref, qry = _sk(210, 16), _sk(90, 16, sample_seed=1)
    for bnd in ((2, 0.02, 0.2, 0.9, 0.8), (0, 0.015, 0.0, 1.0, 1.0), (1, 0.0, 0.1, 0.5, 0.5), (2, 0.0, 0.2, 1.0, 1.0)):
        d, lab, _ = eng.query_host(ref, qry, KMERS, boundary=bnd)
        d_o, lab_o, _ = oracle.query(ref, qry, KMERS, boundary=bnd)
        # labels are exact wherever the float32 distances are bit-identical (they are compared to a threshold)
        same = (d == d_o).all(axis=1)
        assert same.mean() > 0.99 and (lab[same] == lab_o[same]).all()
        # and they are exactly assign_threshold of the engine's own distances / scale
        scaled = (d / np.array(bnd[3:], dtype=np.float32)).astype(np.float32)
        assert (lab == oracle.assign_threshold(scaled, bnd[0], bnd[1], bnd[2]).astype(np.int8)).all()
    _, lab_only, _ = eng.query_host(ref, qry, KMERS, boundary=(2, 0.02, 0.2, 0.9, 0.8), want_out=False)
    d, lab, _ = eng.query_host(ref, qry, KMERS, boundary=(2, 0.02, 0.2, 0.9, 0.8))
    assert (lab_only == lab).all()


# ---------------------------------------------------------------- the drop-in wrapper
def test_queryDatabase_dropin(eng, oracle, tmp_path):
    from poppunk_b200 import sketchlib
    db_k = np.array([13, 15, 19, 23, 27, 31], dtype=np.int32)
    sk = _sk(60, 16, kmers=db_k)
    names = [f"s{i:03d}" for i in range(60)]
    tab, cl = synth.random_match_table(db_k, 3), synth.synth_clusters(60, 3)
    ref_prefix = str(tmp_path / "refdb")
    sketchlib.write_db_npz(ref_prefix, names[:45], db_k, sk[:45], tab, cl[:45])
    qry_prefix = str(tmp_path / "qrydb")
    sketchlib.write_db_npz(qry_prefix, names[45:], db_k, sk[45:])
    sub = [names[i] for i in (7, 3, 40, 11, 0, 29)]           # caller-ordered subset
    kidx = [1, 2, 3, 4, 5]
    d = sketchlib.queryDatabase(sub, sub, ref_prefix, ref_prefix, KMERS, self=True)
    idx = [7, 3, 40, 11, 0, 29]
    exp, _ = oracle.query(sk[idx][:, kidx], None, KMERS, tab[:, :, kidx], cl[idx])
    assert d.dtype == np.float32 and d.shape == (15, 2) and np.abs(d - exp).max() <= TOL
    qn = names[45:52]
    d = sketchlib.queryDatabase(names[:45], qn, ref_prefix, qry_prefix, KMERS, self=False)
    exp, _ = oracle.query(sk[:45][:, kidx], sk[45:52][:, kidx], KMERS, tab[:, :, kidx], cl[:45],
                          np.zeros(7, dtype=np.uint16))
    assert d.shape == (45 * 7, 2) and np.abs(d - exp).max() <= TOL
    # the native-entry twin, positional as in test/test-update-gpu.py:85-86
    base = ref_prefix + "/refdb"
    j = sketchlib.pp_queryDatabase(base, base, sub, sub, KMERS, False, True, 1, True, 0)
    exp_j, _ = oracle.query(sk[idx][:, kidx], None, KMERS, out_mode=oracle.OUT_JACCARD)
    assert j.shape == (15, 5) and np.abs(j - exp_j).max() <= 1e-7


# ---------------------------------------------------------------- BASELINE config 1: the reference's smoke-test genomes
def test_cfg1_example_set_real_genomes(eng, oracle, golden_dir, tmp_path):
    """The 29 assemblies of the reference's own smoke test (test/example_set.tar.bz2, test/run_test.py:20-21), sketched
    with the stand-in sketcher (tools/standin_sketcher.c: reference schema, not pp-sketchlib's hash values), through the
    drop-in queryDatabase: --create-db style all-vs-all, then the poppunk_assign style query-vs-ref call."""
    from poppunk_b200 import distfiles as utils, sketchlib
    z = np.load(os.path.join(golden_dir, "example_set_sketches.npz"))
    names, db_k, sk = [str(s) for s in z["names"]], z["kmers"], z["sketches"]
    assert len(names) == 29 and names == sorted(names) and sk.shape == (29, len(db_k), 16 * 14)
    prefix = str(tmp_path / "example_db")
    sketchlib.write_db_npz(prefix, names, db_k, sk)
    for klist in (np.arange(13, 30, 4), np.arange(13, 29, 3)):       # PopPUNK defaults; the smoke test's --k-step 3
        kidx = [int(np.where(db_k == k)[0][0]) for k in klist]
        d = sketchlib.queryDatabase(names, names, prefix, prefix, klist, self=True)
        exp, _ = oracle.query(sk[:, kidx], None, klist.astype(np.int32))
        assert d.shape == (406, 2) and d.dtype == np.float32 and np.abs(d - exp).max() <= TOL
        # real-genome sanity: distances of one species, and an assembly against its own truncated copy
        assert 0.0 <= d.min() and np.median(d[:, 0]) < 0.06 and np.median(d[:, 1]) < 0.3
        rows = {pair: r for r, pair in enumerate(utils.iterDistRows(names, names, True))}
        core, acc = d[rows[("12754_4#89_partial", "12754_4#89")]]
        assert core < 0.03 and acc > 0.5                               # same sequence, 98 % of it missing
    # assign-style: 5 of the genomes as queries against the other 24 (names must be disjoint, sketchlib.py:575-580)
    qn, rn = names[::6], [n for i, n in enumerate(names) if i % 6]
    qprefix = str(tmp_path / "query_db")
    sketchlib.write_db_npz(qprefix, qn, db_k, sk[::6])
    klist = np.arange(13, 30, 4)
    kidx = [int(np.where(db_k == k)[0][0]) for k in klist]
    d = sketchlib.queryDatabase(rn, qn, prefix, qprefix, klist, self=False)
    ridx = [names.index(n) for n in rn]
    exp, _ = oracle.query(sk[ridx][:, kidx], sk[::6][:, kidx], klist.astype(np.int32))
    assert d.shape == (len(rn) * len(qn), 2) and np.abs(d - exp).max() <= TOL
    with pytest.raises(SystemExit):
        sketchlib.queryDatabase(names, qn, prefix, qprefix, klist, self=False)


# ---------------------------------------------------------------- BASELINE config 2 at full size
def test_cfg2_full_size_properties(eng, oracle):
    """10k genomes, S=1024, k={15,19,23,27,31}: 49 995 000 rows.  Checked by size-independent properties and
    against the oracle on sampled row ranges (the oracle needs minutes for all of it)."""
    import torch
    n = 10_000
    ref = _sk(n, 16, n_lineages=8)
    tab, cl = synth.random_match_table(KMERS, 3), synth.synth_clusters(n, 3)
    packed = eng.pack(ref, clusters=cl)
    out, _, ndeg = eng.query(packed, None, KMERS, rand_table=tab)
    torch.cuda.synchronize()
    total = n * (n - 1) // 2
    assert out.shape == (total, 2)
    assert bool(torch.isfinite(out).all()) and float(out.min()) >= 0.0 and float(out.max()) < 1.0
    # shards reproduce the full result exactly (checksum of checksums)
    full_sum = out.double().sum(dim=0)
    acc = torch.zeros(2, dtype=torch.float64, device=out.device)
    for r in range(4):
        b, e, _ = eng.shard_rows(total, 4, r)
        part, _, _ = eng.query(packed, None, KMERS, rand_table=tab, row_begin=b, row_end=e)
        assert bool((part == out[b:e]).all())
        acc += part.double().sum(dim=0)
    assert torch.allclose(acc, full_sum, rtol=1e-12)
    # oracle on sampled row ranges, including the first and last rows and a row boundary
    host = out.cpu().numpy()
    rng = np.random.default_rng(0)
    starts = [0, n - 2, total - 4000] + rng.integers(0, total - 4000, size=12).tolist()
    for b in starts:
        exp, _ = oracle.query(ref, None, KMERS, tab, cl, row_begin=int(b), row_end=int(b) + 4000)
        assert np.abs(host[b:b + 4000] - exp).max() <= TOL


# ---------------------------------------------------------------- BASELINE config 3 (the north-star size) on one GPU
def test_cfg3_north_star_size_properties(eng, oracle):
    """100k genomes, S=1024, K=5: 4 999 950 000 rows (> 2^32, 40 GB of float2).  Size-independent properties plus the
    oracle on sampled row ranges: first/last rows, both sides of row 2^32 (64-bit row arithmetic), row-tile and
    genome-row boundaries; a block of the triangle must equal the rectangular query of the same genomes; a shard cut
    inside row tiles must reproduce its slice of the full result."""
    import torch
    n = 100_000
    total = n * (n - 1) // 2
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < total * 8 + (8 << 30):
        pytest.skip("needs ~50 GB of free device memory")
    kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
    sk = synth.synth_sketches_torch(n, kmers, 16, seed=42, device="cuda")
    packed = eng.pack(sk)
    out, _, ndeg = eng.query(packed, None, kmers)
    torch.cuda.synchronize()
    assert out.shape == (total, 2) and total > 2 ** 32
    mn, mx = out.aminmax()
    assert float(mn) >= 0.0 and float(mx) < 1.0 and not bool(torch.isnan(out[::997]).any())
    ref_host = sk.cpu().numpy().view(np.uint64)
    i_mid = oracle.calc_row_idx(2 ** 32, n)
    starts = [0, total - 3000, 2 ** 32 - 1500, oracle.square_to_condensed(i_mid, i_mid + 1, n) - 1500,
              oracle.square_to_condensed(64 * 700, 64 * 700 + 1, n) - 1500, oracle.square_to_condensed(99_935, 99_936, n) - 1500]
    for b in starts:
        exp, _ = oracle.query(ref_host, None, kmers, row_begin=int(b), row_end=int(b) + 3000)
        got = out[b:b + 3000].cpu().numpy()
        assert np.abs(got - exp).max() <= TOL, b
    # triangle block == rectangle of the same genomes: rows (i, j) with i in [70000, 70064), j in [90000, 90200)
    qi, rj = np.arange(70_000, 70_064), np.arange(90_000, 90_200)
    rect, _, _ = eng.query(eng.pack(sk, idx=rj), eng.pack(sk, idx=qi), kmers)     # row = q * 200 + r
    rows = torch.as_tensor([[oracle.square_to_condensed(int(i), int(j), n) for j in rj] for i in qi], device=out.device)
    assert bool((out[rows.reshape(-1)] == rect).all())
    # a shard that starts and ends inside row tiles reproduces its slice bit for bit
    b, e = 2 ** 32 - 12_345_678, 2 ** 32 + 23_456_789
    part, _, _ = eng.query(packed, None, kmers, row_begin=b, row_end=e)
    assert bool((part == out[b:e]).all())
    assert int(ndeg.item()) == 0


# ---------------------------------------------------------------- multi-GPU (needs >= 2 devices; see tests/multigpu_check.py)
def test_multigpu_paths_match_single_gpu(eng):
    """Spawns tests/multigpu_check.py under torchrun on 2 GPUs: NCCL all-gather path, fused peer-store exchange and
    (if the fabric has it) the multimem.st exchange must be byte-identical to the single-GPU result."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "multigpu_check ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


# ---------------------------------------------------------------- host path: launch plan, ring wrap, staged copy-out
@pytest.mark.parametrize("pinned", [False, True])
def test_host_path_many_chunks_ring_wraps(eng, oracle, monkeypatch, pinned):
    """ppb_query_host with a tiny chunk cap: dozens of row-tile-aligned launches through the 8-deep buffer ring,
    result and labels either DMA'd straight into pinned memory or staged and memcpy'd out to pageable memory."""
    import torch
    ref, qry = _sk(333, 16), _sk(150, 16, sample_seed=1)
    bnd = (2, 0.02, 0.2, 1.0, 1.0)
    exp_s, lab_s, nd_s = oracle.query(ref, None, KMERS, boundary=bnd)
    exp_r, lab_r, nd_r = oracle.query(ref, qry, KMERS, boundary=bnd)
    monkeypatch.setenv("PPB_HOST_CHUNK_ROWS", "3000")        # 19 launches (self), 17 (rectangle): the ring wraps twice
    for q, exp, lab_o, nd_o in ((None, exp_s, lab_s, nd_s), (qry, exp_r, lab_r, nd_r)):
        out = None
        if pinned:
            out = torch.empty(exp.shape, dtype=torch.float32, pin_memory=True).numpy()
        got, lab, nd = eng.query_host(ref, q, KMERS, boundary=bnd, out=out)
        assert np.abs(got - exp).max() <= TOL and nd == nd_o
        same = (got == exp).all(axis=1)
        assert (lab[same] == lab_o[same]).all()
        b, e = 1234, exp.shape[0] - 777                          # a shard that starts and ends inside row tiles
        part, _, _ = eng.query_host(ref, q, KMERS, row_begin=b, row_end=e)
        assert (part == got[b:e]).all()


# ---------------------------------------------------------------- sketches of more than 65535 bins (uint32 count tile)
@pytest.mark.parametrize("n,ss64", [(70, 1024), (5, 1500), (4, 15625)])
def test_huge_sketches_counts_and_distances(eng, oracle, n, ss64):
    """PopPUNK accepts --sketch-size up to 10^6 bins (__main__.py:310): sketchsize64 = 15625.  Above 1023 the per-k counts
    no longer fit uint16; the kernel then keeps uint32 counts (and fits with in-place logs)."""
    kmers = np.array([15, 21, 29], dtype=np.int32)
    ref = _sk(n, ss64, kmers=kmers, n_lineages=2, chunk=8)
    cnt, _, _ = eng.query_host(ref, None, kmers, out_mode=eng.OUT_COUNTS)
    cnt_o, _ = oracle.query(ref, None, kmers, out_mode=oracle.OUT_COUNTS)
    assert (cnt == cnt_o).all() and (ss64 < 1500 or int(cnt.max()) > 65535)   # the two big ones really overflow uint16
    tab, cl = synth.random_match_table(kmers, 2), synth.synth_clusters(n, 2)
    d, lab, ndeg = eng.query_host(ref, None, kmers, tab, cl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
    d_o, lab_o, ndeg_o = oracle.query(ref, None, kmers, tab, cl, boundary=(2, 0.02, 0.2, 1.0, 1.0))
    assert ndeg == ndeg_o and np.abs(d - d_o).max() <= TOL
    same = (d == d_o).all(axis=1)
    assert (lab[same] == lab_o[same]).all()
    if n >= 5:
        qry = _sk(3, ss64, kmers=kmers, sample_seed=1, n_lineages=2, chunk=8)
        dq, _, _ = eng.query_host(ref, qry, kmers)
        dq_o, _ = oracle.query(ref, qry, kmers)
        assert np.abs(dq - dq_o).max() <= TOL
