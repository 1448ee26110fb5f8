"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/ppb.h declares, its host-only
entry points agree with the oracle, the host logic (sharding, DB reading, error conventions) behaves like the
reference, and the product path fails loudly without a CUDA device."""
import os
import re
import sys

import numpy as np
import pytest

from poppunk_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KMERS = np.array([15, 19, 23, 27, 31], dtype=np.int32)


@pytest.fixture(scope="module")
def lib():
    from poppunk_b200 import build, _lib
    build.build()
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from poppunk_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ppb.h")).read()
    declared = set(re.findall(r"\b(ppb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ppb_version() == 100


def test_abi_index_maps_match_oracle(lib, oracle):
    for n in (2, 5, 64, 1001, 100_000):
        total = n * (n - 1) // 2
        for k in {0, 1, n - 2, n - 1, total // 3, total // 2, total - 2, total - 1} - {-1}:
            if not 0 <= k < total:
                continue
            i = lib.ppb_calc_row_idx(k, n)
            j = lib.ppb_calc_col_idx(k, i, n)
            assert (i, j) == (oracle.calc_row_idx(k, n), oracle.calc_col_idx(k, i, n))
            assert lib.ppb_square_to_condensed(i, j, n) == k
    assert lib.ppb_num_rows(100_000, 0, 1) == 4_999_950_000
    assert lib.ppb_num_rows(50_000, 1_000_000, 0) == 50_000_000_000
    # packed size: K * ceil(2*ss64/32) slices * n padded to 128 * 1792 B
    assert lib.ppb_packed_bytes(1000, 5, 16) == 5 * 1 * 1024 * 1792
    assert lib.ppb_packed_bytes(10, 6, 156) == 6 * 10 * 128 * 1792


def test_argument_errors_without_gpu(lib):
    from poppunk_b200._lib import OUT_DISTS
    k = KMERS.copy()
    ref = synth.synth_sketches(4, KMERS, 2)
    out = np.empty((6, 2), dtype=np.float32)
    rc = lib.ppb_query_host(ref.ctypes.data, 4, None, 0, k.ctypes.data, 5, 2, 13, None, 0, None, None, 0, 6,
                            OUT_DISTS, out.ctypes.data, None, None, None, 0)
    assert rc == 1 and b"bbits" in lib.ppb_last_error()
    rc = lib.ppb_query_host(ref.ctypes.data, 4, None, 0, k.ctypes.data, 5, 2, 14, None, 0, None, None, 0, 7,
                            OUT_DISTS, out.ctypes.data, None, None, None, 0)
    assert rc == 1 and b"row range" in lib.ppb_last_error()


def test_fails_loudly_without_cuda(lib):
    """No CPU fallback: on a box without a GPU every compute entry raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from poppunk_b200 import engine, refine
    ref = synth.synth_sketches(4, KMERS, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.query_host(ref, None, KMERS)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.pack(ref)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        refine.assignThreshold(np.zeros((3, 2), dtype=np.float32), 2, 0.5, 0.5)


def test_product_never_imports_oracle():
    """The product path must not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poppunk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "libppo" not in txt and "ppo_" not in txt and "import oracle" not in txt, f


def test_shard_rows():
    from poppunk_b200.engine import shard_rows, num_rows
    total = num_rows(100_000)
    for world in (1, 2, 4, 8):
        spans = [shard_rows(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert len({s[2] for s in spans}) == 1 and spans[0][2] * world >= total
    assert shard_rows(5, 8, 7) == (5, 5, 1) and shard_rows(0, 4, 2) == (0, 0, 0)
    assert num_rows(50_000, 1_000_000) == 50_000_000_000


def _plan(lib, n_ref, n_qry, self_mode, b, e, cap):
    n = lib.ppb_plan_host_chunks(n_ref, n_qry, int(self_mode), b, e, cap, None, 0)
    assert n >= 0
    bounds = np.zeros((max(n, 1), 2), dtype=np.int64)
    assert lib.ppb_plan_host_chunks(n_ref, n_qry, int(self_mode), b, e, cap, bounds.ctypes.data, n) == n
    return bounds[:n]


@pytest.mark.parametrize("n_ref,n_qry,self_mode,cap", [(1000, 0, True, 50_000), (1000, 0, True, 700),
                                                       (130, 0, True, 1 << 26), (300, 517, False, 40_000),
                                                       (300, 517, False, 100), (65, 0, True, 64), (2, 0, True, 5)])
def test_host_chunks_cover_rows_and_end_on_row_tiles(lib, n_ref, n_qry, self_mode, cap):
    """ppb_query_host's launch plan: contiguous cover of the shard, <= cap rows each, and every interior cut is
    the first row of a genome that starts a 64-genome row tile whenever such a cut fits under the cap."""
    total = lib.ppb_num_rows(n_ref, n_qry, int(self_mode))
    for b, e in [(0, total), (total // 3, total - total // 5), (7, 8)]:
        if e > total or b >= e:
            continue
        ch = _plan(lib, n_ref, n_qry, self_mode, b, e, cap)
        assert ch[0, 0] == b and ch[-1, 1] == e
        assert (ch[1:, 0] == ch[:-1, 1]).all() and ((ch[:, 1] - ch[:, 0]) > 0).all()
        assert ((ch[:, 1] - ch[:, 0]) <= cap).all()
        n_side = n_ref if self_mode else n_qry
        tile_starts = set()
        for g in range(0, n_side, 64):
            if self_mode:
                tile_starts.add(total if g >= n_ref - 1 else lib.ppb_square_to_condensed(g, g + 1, n_ref))
            else:
                tile_starts.add(g * n_ref)
        rows_per_tile = 64 * n_ref
        for r0, r1 in ch[:-1]:
            assert int(r1) in tile_starts or cap < rows_per_tile
    assert lib.ppb_plan_host_chunks(10, 0, 1, 0, 46, 10, None, 0) == -1   # row_end beyond the triangle


def test_dropin_error_conventions(tmp_path, lib):
    """PopPUNK/sketchlib.py:523-524 (RuntimeError) and :575-580 (message + sys.exit(1))."""
    from poppunk_b200 import sketchlib
    names = [f"g{i}" for i in range(6)]
    sk = synth.synth_sketches(6, KMERS, 2)
    p = str(tmp_path / "db")
    sketchlib.write_db_npz(p, names, KMERS, sk)
    with pytest.raises(RuntimeError, match="Must use same db for self query"):
        sketchlib.queryDatabase(names, names, p, str(tmp_path / "other"), KMERS, self=True)
    with pytest.raises(SystemExit) as e:
        sketchlib.queryDatabase(names, names[2:4], p, p, KMERS, self=False)
    assert e.value.code == 1
    db = sketchlib.read_db(p)
    assert db.names == names and (db.kmers == KMERS).all() and db.sketchsize64 == 2 and db.bbits == 14
    assert (db.index_of(["g4", "g0"]) == [4, 0]).all() and (db.k_index([19, 31]) == [1, 4]).all()
    with pytest.raises(RuntimeError, match="not found"):
        db.index_of(["nope"])
    with pytest.raises(RuntimeError, match="k-mer length 21"):
        db.k_index([21])
    kmers, ss, codon = sketchlib.readDBParams(p)
    assert (kmers == KMERS).all() and ss == 2 and codon is False
    assert sketchlib.getSeqsInDb(p) == names


def _gloo_worker(rank, world, port, tmpdir):
    """world_size-2 CPU check of the N>1 path: static row shards + one all_gather_into_tensor reassemble
    exactly the single-process result.  The per-shard compute is the oracle here (no GPU)."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from poppunk_b200 import engine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ref = synth.synth_sketches(61, KMERS, 4, seed=3)
        qry = synth.synth_sketches(9, KMERS, 4, seed=3, sample_seed=1)

        def oracle_query(r, q, kmers, rand_table, b, e, out_mode, out=None):
            res, ndeg = oracle.query(r, q, kmers, row_begin=b, row_end=e, out_mode=out_mode, threads=1)
            t = torch.from_numpy(res.view(np.int32) if res.dtype == np.uint32 else res)
            if out is not None:
                out.copy_(t)
            return (out if out is not None else t), None, torch.tensor([ndeg], dtype=torch.int64)

        class Host:  # stands in for PackedSketches on a CPU box
            def __init__(self, a):
                self.a, self.n, self.K, self.device = a, a.shape[0], a.shape[1], torch.device("cpu")

        for q in (None, qry):
            full, ndeg = engine.query_sharded(Host(ref), None if q is None else Host(q), KMERS,
                                              _query_fn=lambda r, qq, *a, **k: oracle_query(
                                                  r.a, None if qq is None else qq.a, *a, **k))
            exp, ndeg_o = oracle.query(ref, q, KMERS, threads=1)
            assert full.shape == exp.shape and (full.numpy() == exp).all(), rank
            assert int(ndeg.item()) == ndeg_o
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_allgather_gloo_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_dist_pickle_roundtrip_and_row_order(tmp_path, lib):
    """N4: <prefix>.dists.pkl/.npy as PopPUNK/utils.py:135-197 writes them, and the row <-> pair conventions
    (utils.py:199-261) agree with the C ABI's index maps."""
    import pickle
    from poppunk_b200 import distfiles as utils
    names = [f"s{i}" for i in range(7)]
    X = np.arange(42, dtype=np.float32).reshape(21, 2)
    prefix = str(tmp_path / "db.dists")
    utils.storePickle(names, names, True, X, prefix)
    assert pickle.load(open(prefix + ".pkl", "rb")) == [names, names, True]     # what the reference's reader expects
    r, q, self_, Y = utils.readPickle(prefix, enforce_self=True)
    assert (r, q, self_) == (names, names, True) and Y.dtype == np.float32 and (Y == X).all()
    assert utils.readPickle(prefix, distances=False)[3] is None
    utils.storePickle(names[:3], names[3:], False, None, prefix)
    with pytest.raises(SystemExit):
        utils.readPickle(prefix, enforce_self=True)
    rows = list(utils.listDistInts(names, names, True))
    assert len(rows) == 21
    for k, (j, i) in enumerate(rows):
        assert lib.ppb_square_to_condensed(i, j, 7) == k and lib.ppb_calc_row_idx(k, 7) == i
    assert list(utils.iterDistRows(names, names, True))[:2] == [("s1", "s0"), ("s2", "s0")]
    assert list(utils.listDistInts(names[:2], names[2:5], False)) == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
    with pytest.raises(RuntimeError):
        list(utils.iterDistRows(names, names[:3], True))


def test_reference_arm_under_torchrun_uses_all_cores():
    """`bench.py --impl reference` launched the way the driver launches N>1 runs: rank 0 alone prints ONE JSON line,
    the other rank exits 0, and the CPU arm still uses every core (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--genomes", "2000"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_bench_traffic_capture_plumbing(tmp_path, monkeypatch):
    """bench.py's in-run DRAM-traffic capture: the command it hands to ncu, the CSV it reads back, and that a missing or
    failing profiler yields (None, reason) instead of an exception — with a stand-in for the ncu binary (no GPU here)."""
    import stat
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    fake = tmp_path / "ncu"
    fake.write_text("""#!/bin/bash
# stand-in: records its arguments, writes the log an ncu --csv --metrics run writes
echo "$@" > "$(dirname "$0")/args.txt"
while [ $# -gt 0 ]; do if [ "$1" = "--log-file" ]; then log="$2"; fi; shift; done
cat > "$log" <<'EOF'
==PROF== Connected to process 468 (/usr/bin/python3.12)
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","468","python3.12","127.0.0.1","void query_kernel<0>(QueryParams)","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","byte","202673757952"
"0","468","python3.12","127.0.0.1","void query_kernel<0>(QueryParams)","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Gbyte","40.5"
EOF
""")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    for k in list(os.environ):
        if k.startswith("NV_COMPUTE_PROFILER") or k in ("CUDA_INJECTION64_PATH", "PPB_BENCH_NO_NCU"):
            monkeypatch.delenv(k)
    monkeypatch.setenv("PPB_NCU", str(fake))
    got, src = bench.capture_traffic(100_000, timeout_s=30)
    assert got == 202673757952 + 40_500_000_000 and "in-run ncu capture" in src
    argv = (tmp_path / "args.txt").read_text().split()
    assert "--traffic-child" in argv and argv[argv.index("--genomes") + 1] == "100000"
    assert "regex:query_kernel" in argv and argv[argv.index("-s") + 1] == "1" and argv[argv.index("-c") + 1] == "1"
    assert argv[argv.index("--metrics") + 1] == "dram__bytes_read.sum,dram__bytes_write.sum"
    fake.write_text("#!/bin/bash\necho 'no counters for you' >&2\nexit 1\n")
    got, src = bench.capture_traffic(100_000, timeout_s=30)
    assert got is None and "rc=1" in src
    monkeypatch.setenv("PPB_NCU", str(tmp_path / "absent"))
    assert bench.capture_traffic(100_000)[0] is None
    monkeypatch.setenv("PPB_BENCH_NO_NCU", "1")
    assert bench.capture_traffic(100_000) == (None, "in-run capture disabled (PPB_BENCH_NO_NCU)")


@pytest.mark.parametrize("n_ref,n_qry,self_mode,tile_cols,band", [(300, 0, True, 128, 2), (130, 0, True, 128, 64),
                                                                  (200, 0, True, 32, 3), (257, 150, False, 128, 2),
                                                                  (65, 1, False, 64, 1), (2, 0, True, 128, 4)])
def test_tile_schedule_covers_every_pair_once(lib, n_ref, n_qry, self_mode, tile_cols, band):
    """The launch plan of the distance kernel (ppb_plan_tiles, the list ppb_query_dev uploads): for the whole job and
    for shards that start/end inside row tiles, every (i, j) pair of the shard lies in exactly one listed tile, no tile
    is listed twice, and self mode lists no tile without a j > i."""
    total = lib.ppb_num_rows(n_ref, n_qry, int(self_mode))
    n_rows_side = n_ref if self_mode else n_qry
    for b, e in [(0, total), (total // 3, total - total // 5), (total // 2, total // 2 + 1)]:
        if b >= e:
            continue
        n = lib.ppb_plan_tiles(n_ref, n_qry, int(self_mode), b, e, tile_cols, band, None, 0)
        assert n > 0
        tiles = np.zeros((n, 2), dtype=np.int32)
        assert lib.ppb_plan_tiles(n_ref, n_qry, int(self_mode), b, e, tile_cols, band, tiles.ctypes.data, n) == n
        assert len({(int(t), int(c)) for t, c in tiles}) == n                         # no duplicates
        listed = {(int(t), int(c)) for t, c in tiles}
        rows = np.arange(b, e, max(1, (e - b) // 4000))                                # sample of the shard's rows
        rows = np.unique(np.concatenate([rows, [b, e - 1]]))
        for r in rows:
            if self_mode:
                i = lib.ppb_calc_row_idx(int(r), n_ref)
                j = lib.ppb_calc_col_idx(int(r), i, n_ref)
            else:
                i, j = int(r) // n_ref, int(r) % n_ref
            assert 0 <= i < n_rows_side and (i // 64, j // tile_cols) in listed, (r, i, j)
        if self_mode:                                                                  # every listed tile holds some j > i
            assert all((c + 1) * tile_cols - 1 > t * 64 for t, c in tiles)
        i_lo = lib.ppb_calc_row_idx(b, n_ref) if self_mode else b // n_ref
        i_hi = lib.ppb_calc_row_idx(e - 1, n_ref) if self_mode else (e - 1) // n_ref
        assert tiles[:, 0].min() == i_lo // 64 and tiles[:, 0].max() == i_hi // 64     # no row tile outside the shard
    assert lib.ppb_plan_tiles(n_ref, n_qry, int(self_mode), 0, total + 1, tile_cols, band, None, 0) == -1


@pytest.mark.parametrize("n_ref,n_qry,self_mode,G,cap", [(700, 0, True, 3, 5000), (257, 150, False, 2, 3000),
                                                        (1200, 0, True, 1, 2048), (50, 1000, False, 8, 1024)])
def test_shards_chunks_tiles_compose(lib, n_ref, n_qry, self_mode, G, cap):
    """The three plans a host call stacks — device shards, row chunks of a shard (in the device's own coordinates:
    non-self shards are re-based on the first query the device holds, ppb_host.inl), tile list of a chunk — put every row
    of the job into exactly one launch whose tile list holds the row's tile."""
    import ctypes as C
    total = lib.ppb_num_rows(n_ref, n_qry, int(self_mode))
    cuts = (C.c_int64 * (G + 1))()
    assert lib.ppb_plan_device_shards(n_ref, n_qry, int(self_mode), 0, total, G, cuts) == G
    covered = 0
    for g in range(G):
        r_lo, r_hi = cuts[g], cuts[g + 1]
        if r_hi <= r_lo:
            continue
        q_lo = 0 if self_mode else r_lo // n_ref
        n_q = 0 if self_mode else (r_hi - 1) // n_ref + 1 - q_lo
        shift = q_lo * n_ref
        n = lib.ppb_plan_host_chunks(n_ref, n_q, int(self_mode), r_lo - shift, r_hi - shift, cap, None, 0)
        bounds = np.zeros((n, 2), dtype=np.int64)
        assert lib.ppb_plan_host_chunks(n_ref, n_q, int(self_mode), r_lo - shift, r_hi - shift, cap, bounds.ctypes.data, n) == n
        assert bounds[0, 0] == r_lo - shift and bounds[-1, 1] == r_hi - shift and (bounds[1:, 0] == bounds[:-1, 1]).all()
        for c0, c1 in bounds:
            c0, c1 = int(c0), int(c1)
            assert 0 < c1 - c0 <= max(cap, 64 * n_ref)
            nt = lib.ppb_plan_tiles(n_ref, n_q, int(self_mode), c0, c1, 128, 12, None, 0)
            tiles = np.zeros((nt, 2), dtype=np.int32)
            assert lib.ppb_plan_tiles(n_ref, n_q, int(self_mode), c0, c1, 128, 12, tiles.ctypes.data, nt) == nt
            listed = {(int(t), int(c)) for t, c in tiles}
            assert len(listed) == nt
            for r in np.unique(np.concatenate([np.arange(c0, c1, max(1, (c1 - c0) // 50)), [c1 - 1]])):
                if self_mode:
                    i = lib.ppb_calc_row_idx(int(r), n_ref)
                    j = lib.ppb_calc_col_idx(int(r), i, n_ref)
                else:
                    i, j = int(r) // n_ref, int(r) % n_ref              # i: query index local to the device
                    assert 0 <= i < n_q
                assert (i // 64, j // 128) in listed, (g, c0, c1, r, i, j)
            covered += c1 - c0
    assert covered == total


def test_device_shards_are_balanced_tile_aligned_and_cover(lib):
    """ppb_plan_device_shards: what ppb_query_host_multi gives each device."""
    import ctypes as C
    for n_ref, n_qry, self_ in ((100_000, 0, 1), (50_000, 1_000_000, 0), (1000, 0, 1), (70, 33, 0)):
        total = lib.ppb_num_rows(n_ref, n_qry, self_)
        for G in (1, 2, 3, 8):
            cuts = (C.c_int64 * (G + 1))()
            assert lib.ppb_plan_device_shards(n_ref, n_qry, self_, 0, total, G, cuts) == G
            cuts = list(cuts)
            assert cuts[0] == 0 and cuts[-1] == total and all(a <= b for a, b in zip(cuts, cuts[1:]))
            for c in cuts[1:-1]:      # interior cuts sit where a 64-genome row tile begins (or at the very end)
                if c == total:
                    continue
                if self_:
                    i = lib.ppb_calc_row_idx(c, n_ref)
                    assert i % 64 == 0 and lib.ppb_square_to_condensed(i, i + 1, n_ref) == c
                else:
                    assert c % n_ref == 0 and (c // n_ref) % 64 == 0
            if total > 50_000_000:
                sizes = np.diff(cuts)
                assert sizes.max() / sizes.min() < 1.02
    sub = (C.c_int64 * 3)()
    assert lib.ppb_plan_device_shards(1000, 0, 1, 1234, 400_000, 2, sub) == 2 and sub[0] == 1234 and sub[2] == 400_000
    assert lib.ppb_plan_device_shards(1000, 0, 1, 0, 10**9, 2, sub) == -1


def test_host_pool_without_gpu(lib):
    """ppb_host_alloc / ppb_host_free: blocks are reused; without a CUDA device nothing is page-locked."""
    import ctypes as C
    from poppunk_b200 import engine
    lib.ppb_release_workspace()
    a = engine.host_result((1000, 2), np.float32)
    a[:] = 3.0
    addr = a.ctypes.data
    del a
    b = engine.host_result((999, 2), np.float32)
    assert b.ctypes.data == addr and b.flags.writeable and b.flags.c_contiguous
    h, u, p = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    lib.ppb_host_pool_stats(C.byref(h), C.byref(u), C.byref(p))
    assert h.value == u.value == 2 << 20 and p.value == 0
    assert lib.ppb_host_free(C.c_void_p(12345)) != 0          # not a pool block
    del b
    lib.ppb_release_workspace()
    lib.ppb_host_pool_stats(C.byref(h), C.byref(u), C.byref(p))
    assert h.value == 0
    assert engine.host_result((0, 2), np.float32).shape == (0, 2)


def test_visible_devices_knob(monkeypatch):
    """Which devices a drop-in call uses (engine.visible_devices): all of them with PopPUNK's deviceid leading, or what
    PPB_DEVICES says — with a stand-in for the library's device count (no GPU here)."""
    from poppunk_b200 import _lib, engine

    class FakeLib:
        def __init__(self, n):
            self.n = n

        def ppb_device_count(self):
            return self.n

    monkeypatch.setattr(_lib, "load", lambda: FakeLib(8))
    monkeypatch.delenv("PPB_DEVICES", raising=False)
    assert engine.visible_devices(0) == list(range(8))
    assert engine.visible_devices(3) == [3, 0, 1, 2, 4, 5, 6, 7]
    monkeypatch.setenv("PPB_DEVICES", "single")
    assert engine.visible_devices(5) == [5]
    monkeypatch.setenv("PPB_DEVICES", "4")
    assert engine.visible_devices(6) == [6, 0, 1, 2]
    monkeypatch.setenv("PPB_DEVICES", "0,2, 7")
    assert engine.visible_devices(0) == [0, 2, 7]
    for bad in ("0,0", "1,8", "3,-1"):
        monkeypatch.setenv("PPB_DEVICES", bad)
        with pytest.raises(RuntimeError):
            engine.visible_devices(0)
    monkeypatch.delenv("PPB_DEVICES")
    with pytest.raises(RuntimeError):
        engine.visible_devices(8)                       # PopPUNK's --deviceid beyond the box
    monkeypatch.setattr(_lib, "load", lambda: FakeLib(0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.visible_devices(0)


def test_random_match_fallback_formula():
    """docs/sketching.rst:107-118: r = 1 - (1 - 2 4^-k)^l, J_r = r1 r2 / (r1 + r2 - r1 r2)."""
    from poppunk_b200 import sketchlib
    k = np.array([13, 17, 21], dtype=np.int32)
    lens = np.array([2.0e6, 2.0e6, 3.1e6, 1.2e6])
    tab, rcl, qcl = sketchlib.random_match_fallback(lens, lens[:2], k)
    assert tab.shape == (3, 3, 3) and tab.dtype == np.float32 and qcl.tolist() == [rcl[0], rcl[1]]
    assert rcl[0] == rcl[1] and len({int(c) for c in rcl}) == 3
    r = 1.0 - (1.0 - 2.0 * 4.0 ** (-k.astype(np.float64))) ** 2.0e6
    assert np.allclose(tab[rcl[0], rcl[0]], r * r / (2 * r - r * r), rtol=1e-6)     # the docs' equal-length form
    single, _, _ = sketchlib.random_match_fallback(lens, None, k, use_rc=False)
    r1 = 1.0 - (1.0 - 4.0 ** (-k.astype(np.float64))) ** 2.0e6
    assert np.allclose(single[rcl[0], rcl[0]], r1 * r1 / (2 * r1 - r1 * r1), rtol=1e-6)
    # many distinct lengths: at most 32 classes, every genome within ~its class's spread of the representative
    rng = np.random.default_rng(0)
    many = rng.uniform(1.8e6, 2.4e6, size=5000)
    tab, cl, _ = sketchlib.random_match_fallback(many, None, k)
    assert tab.shape[0] == 32 and np.bincount(cl).min() >= 5000 // 32
    with pytest.raises(RuntimeError):
        sketchlib.random_match_fallback(np.array([2e6, np.nan]), None, k)


def test_fitKmerCurve_matches_the_reference_goldens(golden_dir):
    """tests/golden/fit_kmer_curve.npz = PopPUNK/sketchlib.py:635-670 executed from the reference tree (scipy).
    The drop-in's own fitKmerCurve (used by the --plot-fit probe) solves the same bounded problem in closed form:
    it must land where scipy's trust-region iteration lands (scipy stops ~1e-5 short of an active bound)."""
    from poppunk_b200 import sketchlib
    g = np.load(os.path.join(golden_dir, "fit_kmer_curve.npz"))
    worst_inside = worst = 0.0
    for r in range(len(g["n_k"])):
        n = int(g["n_k"][r])
        got = sketchlib.fitKmerCurve(g["jaccard"][r, :n], g["klist"][r, :n])
        exp = g["expected"][r]
        err = float(np.abs(got - exp).max())
        worst = max(worst, err)
        if exp.min() > 0.003:
            worst_inside = max(worst_inside, err)
    assert worst_inside < 1e-6 and worst < 2e-4
