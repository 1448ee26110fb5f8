"""Generate the committed golden fixtures from the reference tree (run HERE, where /root/reference exists).

    python tests/golden/make_golden.py [/root/reference]

Nothing under tests/ reads /root/reference at run time; only this script does.  It copies no reference
source into the repo: reference *functions* are extracted with ``ast`` and executed in memory, and only
their numeric outputs are saved.

Fixtures written next to this file:
  json_sketch.npz      the reference's only real pp-sketchlib sketch (test/json_sketch.txt): pins the
                       sketch schema W = sketchsize64*bbits, bbits = 14
  refine_grid.npz      test/test-refine.py:46-61 — 10x10 float32 grid, boundary (0.5, 0.5), slopes 0/1/2,
                       labels from the reference's own ``withinBoundary`` restatement (:10-23), and the edge
                       lists its ``iter_tuples`` loop (:30-38) derives from them for a seeded 100-sample cloud
  fit_kmer_curve.npz   PopPUNK/sketchlib.py:635-670 ``fitKmerCurve`` (scipy bounded least squares) run on
                       seeded per-k Jaccard vectors: pins model, clamp and (core, acc) output order
  refine_ref.npz       outputs of the reference's OWN compiled poppunk_refine functions (src/boundary.cpp, src/extend.cpp
                       built into oracle/_ref by `make -C oracle ref`): assign_threshold, edge_iterate, generate_tuples,
                       generate_all_tuples, threshold_iterate_1D/2D, get_kNN_distances, lower_rank, extend on seeded
                       inputs with many tied distances; inputs are stored beside the outputs
"""
import ast
import json
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract_function(path, name, env):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), env)
            return env[name]
    raise KeyError(name)


def json_sketch():
    d = json.load(open(os.path.join(REF, "test", "json_sketch.txt")))
    kmers = sorted(int(k) for k in d if k.isdigit())
    sk = np.stack([np.array(d[str(k)], dtype=np.uint64) for k in kmers])
    np.savez_compressed(os.path.join(HERE, "json_sketch.npz"), kmers=np.array(kmers, dtype=np.int32),
                        sketch=sk, sketchsize64=np.int32(d["sketchsize64"]), bbits=np.int32(d["bbits"]),
                        length=np.int64(d["length"]), bases=np.array(d["bases"], dtype=np.float64))
    print("json_sketch:", kmers, sk.shape, d["sketchsize64"], d["bbits"])


def refine_grid():
    within = extract_function(os.path.join(REF, "test", "test-refine.py"), "withinBoundary", {"np": np})
    x = np.arange(0, 1, 0.1, dtype=np.float32)
    y = np.arange(0, 1, 0.1, dtype=np.float32)
    xv, yv = np.meshgrid(x, y)
    dist = np.hstack((xv.reshape(-1, 1), yv.reshape(-1, 1)))
    labels = np.stack([within(dist, 0.5, 0.5, s) for s in (0, 1, 2)]).astype(np.float32)
    # a seeded random cloud like test-refine.py:64-66 (the reference's is unseeded)
    rng = np.random.default_rng(7)
    cloud = rng.random((4950, 2)).astype(np.float32)
    cloud_labels = np.stack([within(cloud, 0.5, 0.5, s) for s in (0, 1, 2)]).astype(np.float32)
    # edge lists the reference test expects from generateTuples / edgeThreshold (test-refine.py:30-38, 64-82):
    # its own Python loop over the condensed rows of 100 samples
    iter_tuples = extract_function(os.path.join(REF, "test", "test-refine.py"), "iter_tuples", {})
    edges = {f"cloud_edges_{s}": np.array(iter_tuples(cloud_labels[s], 100), dtype=np.int64).reshape(-1, 2)
             for s in (0, 1, 2)}
    np.savez_compressed(os.path.join(HERE, "refine_grid.npz"), dist=dist, labels=labels, cloud=cloud,
                        cloud_labels=cloud_labels, x_max=np.float32(0.5), y_max=np.float32(0.5), **edges)
    print("refine_grid:", dist.shape, labels.shape, [int((l == -1).sum()) for l in labels])


def fit_kmer_curve():
    from scipy import optimize
    fit = extract_function(os.path.join(REF, "PopPUNK", "sketchlib.py"), "fitKmerCurve",
                           {"np": np, "optimize": optimize, "sys": sys})
    rng = np.random.default_rng(11)
    klists = [np.arange(13, 30, 4), np.array([15, 19, 23, 27, 31]), np.arange(13, 30, 3), np.arange(14, 30, 3)]
    rows = []
    for klist in klists:
        jacobian = -np.hstack((np.ones((klist.shape[0], 1)), klist.reshape(-1, 1)))
        for rep in range(60):
            if rep < 40:
                core = rng.uniform(0.0005, 0.05)
                acc = rng.uniform(0.005, 0.5)
            else:  # near-identical genomes: sketch noise pushes slope/intercept to the <= 0 bounds
                core = rng.uniform(0.0, 0.0003)
                acc = rng.uniform(0.0, 0.003)
            S = 1024
            # binomial sketch noise so that some fits hit the <= 0 clamps
            jac = rng.binomial(S, (1 - acc) * (1 - core) ** klist) / S
            if (jac <= 0).any():
                continue
            res = fit(jac, klist, jacobian)
            rows.append((klist, jac, np.asarray(res, dtype=np.float64)))
    K_max = max(len(r[0]) for r in rows)
    kl = np.zeros((len(rows), K_max), dtype=np.int32)
    jc = np.zeros((len(rows), K_max), dtype=np.float64)
    nk = np.zeros(len(rows), dtype=np.int32)
    ex = np.zeros((len(rows), 2), dtype=np.float64)
    for r, (klist, jac, res) in enumerate(rows):
        nk[r] = len(klist)
        kl[r, :nk[r]] = klist
        jc[r, :nk[r]] = jac
        ex[r] = res
    np.savez_compressed(os.path.join(HERE, "fit_kmer_curve.npz"), klist=kl, jaccard=jc, n_k=nk, expected=ex)
    print("fit_kmer_curve:", len(rows), "rows; clamped core:", int((ex[:, 0] == 0).sum()),
          "clamped acc:", int((ex[:, 1] == 0).sum()))


def refine_ref():
    root = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import oracle
    oracle.build_ref()
    R = oracle.ref
    rng = np.random.default_rng(20240)
    out = {}

    def put(name, arrays):
        for t, a in enumerate(arrays):
            out[f"{name}.{t}"] = np.asarray(a)

    n = 60
    rows = n * (n - 1) // 2
    d = (rng.random((rows, 2)) * 0.5).astype(np.float32)
    d[::7] = d[3]            # tied rows
    d[5] = (0.0, 0.0)
    out["dists"] = d
    for slope, xm, ym in [(2, 0.2, 0.3), (0, 0.1, 0.0), (1, 0.0, 0.2), (2, 0.0, 0.3)]:
        put(f"assign_threshold/{slope}/{xm}/{ym}", [R.assign_threshold(d, slope, xm, ym)])
        put(f"edge_iterate/{slope}/{xm}/{ym}", R.edge_iterate(d, slope, xm, ym))
    lab = rng.integers(-1, 2, rows).astype(np.int32)
    out["labels"] = lab
    for self_, nr, off in [(1, 0, 0), (1, 0, 5), (0, 12, 0), (0, 12, 3)]:
        put(f"generate_tuples/{self_}/{nr}/{off}", R.generate_tuples(lab, -1, bool(self_), nr, off))
    for nr, nq, self_, off in [(17, 0, 1, 0), (17, 0, 1, 4), (5, 7, 0, 0), (5, 7, 0, 3)]:
        put(f"generate_all_tuples/{nr}/{nq}/{self_}/{off}", R.generate_all_tuples(nr, nq, bool(self_), off))
    offs = np.linspace(-0.1, 0.5, 23)
    out["offsets"] = offs
    for slope in (0, 1, 2):
        put(f"threshold_iterate_1d/{slope}", R.threshold_iterate_1d(d, offs, slope, 0.05, 0.05, 0.4, 0.45))
    xmr = (np.sort(rng.random(9)) * 0.6).astype(np.float32)
    out["x_max_range"] = xmr
    put("threshold_iterate_2d", R.threshold_iterate_2d(d, xmr, 0.3))
    sq = np.round(rng.random((40, 40)), 1).astype(np.float32)   # one decimal: many ties
    rect = np.round(rng.random((7, 30)), 1).astype(np.float32)
    out["square"], out["rect"] = sq, rect
    for k in (1, 3, 39):
        put(f"knn/square/{k}", R.get_knn_distances(sq, k))
    put("knn/rect/1", R.get_knn_distances(rect, 1))
    ci, cj, cd = R.get_knn_distances(sq, 10)
    for rec in (0, 1):
        for cu in (0, 1):
            for k in (1, 2, 5):
                put(f"lower_rank/{rec}/{cu}/{k}", R.lower_rank(ci, cj, cd, 40, k, bool(rec), bool(cu), 0.05))
    qr = np.round(rng.random((40, 9)), 1).astype(np.float32)
    qq = np.round(rng.random((9, 9)), 1).astype(np.float32)
    out["qr"], out["qq"] = qr, qq
    for k in (1, 3, 8):
        put(f"extend/{k}", R.extend(ci, cj, cd, qq, qr, k))
    np.savez_compressed(os.path.join(HERE, "refine_ref.npz"), **out)
    print("refine_ref:", len(out), "arrays from oracle/_ref/libpprefine_ref.so")


if __name__ == "__main__":
    refine_ref()
    json_sketch()
    refine_grid()
    fit_kmer_curve()
