"""Drop-in for the distance path of ``PopPUNK/sketchlib.py`` — same names, arguments and errors.

    queryDatabase(rNames, qNames, dbPrefix, queryPrefix, klist, self=True, number_plot_fits=0,
                  threads=1, use_gpu=False, deviceid=0) -> np.float32 [n_pairs, 2]

mirrors PopPUNK/sketchlib.py:475-632; ``pp_queryDatabase`` mirrors the native entry it wraps
(``pp_sketchlib.queryDatabase``, positional order pinned by test/test-update-gpu.py:85-86).  Install by
replacing ``dbFuncs['queryDatabase']`` (PopPUNK/utils.py:118-126) or ``PopPUNK.sketchlib.queryDatabase`` —
see INTEGRATION.md.

Differences from the reference, all deliberate:
* the arithmetic always runs on the GPU engine (``use_gpu`` is accepted and ignored; ``threads`` is unused);
  with no CUDA device the call raises — there is no CPU fallback;
* ONE call uses EVERY visible GPU (static row shards, ``ppb_query_host_multi``), ``deviceid`` leading;
  ``PPB_DEVICES=single`` (or ``CUDA_VISIBLE_DEVICES``) restricts it — see :func:`poppunk_b200.engine.visible_devices`;
* a pair whose fit has fewer than two usable k is returned as (0, 0) and counted, with the reference's
  warning text, instead of aborting the process (docs/troubleshooting.rst:176-191);
* sketch databases are read from ``<prefix>/<basename>.h5`` when ``h5py`` is importable (schema:
  PopPUNK/web.py:14-61) and otherwise from ``<prefix>/<basename>.npz`` (same content, documented in
  :func:`write_db_npz`) — h5py is not in this image;
* random-match correction (``random_correct=True`` on every production call, sketchlib.py:533,589): the
  database's ``/random`` table when it has one; otherwise the reference's documented closed form
  (docs/sketching.rst:107-118) from the genome lengths, with the reference's message
  (docs/query_assignment.rst:110) — see :func:`random_match_fallback` for the one approximation made.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from random import sample
from typing import List, Optional, Sequence

import numpy as np

from . import engine
from ._lib import BBITS, OUT_DISTS, OUT_JACCARD

try:  # optional: not present in the build image
    import h5py  # type: ignore
except Exception:  # pragma: no cover
    h5py = None

MAX_FALLBACK_CLUSTERS = 32   # length classes of the closed-form random-match table (C*C*K*(S+1) doubles stay in L2)


# --------------------------------------------------------------------------------------------
# sketch database access (reference readers: PopPUNK/sketchlib.py:109-214)
# --------------------------------------------------------------------------------------------
@dataclass
class SketchDB:
    names: List[str]
    kmers: np.ndarray                 # int32 [Kdb], ascending
    sketchsize64: int
    bbits: int
    sketches: np.ndarray              # uint64 [n][Kdb][W]
    codon_phased: bool = False
    random_table: Optional[np.ndarray] = None      # float32 [C][C][Kdb]
    random_clusters: Optional[np.ndarray] = None   # uint16 [n]; 0xFFFF = sample not in the table
    lengths: Optional[np.ndarray] = None           # float64 [n]  genome length (HDF5 attr 'length', web.py:33-61)
    base_freq: Optional[np.ndarray] = None         # float64 [n][4] (attr 'base_freq')
    random_centroids: Optional[np.ndarray] = None  # float64 [C][4] base-composition centroids of the table's clusters
    use_rc: bool = True

    def index_of(self, names: Sequence[str]) -> np.ndarray:
        lut = {n: i for i, n in enumerate(self.names)}
        try:
            return np.fromiter((lut[n] for n in names), dtype=np.int64, count=len(names))
        except KeyError as e:
            raise RuntimeError(f"Sample {e.args[0]} not found in sketch database") from None

    def k_index(self, klist) -> np.ndarray:
        pos = {int(k): i for i, k in enumerate(self.kmers)}
        try:
            return np.array([pos[int(k)] for k in klist], dtype=np.int64)
        except KeyError as e:
            raise RuntimeError(f"k-mer length {e.args[0]} not found in sketch database") from None


NOT_IN_TABLE = np.uint16(0xFFFF)


def _db_file(prefix: str) -> str:
    base = os.path.join(prefix, os.path.basename(prefix))
    if h5py is not None and os.path.exists(base + ".h5"):
        return base + ".h5"
    if os.path.exists(base + ".npz"):
        return base + ".npz"
    if os.path.exists(base + ".h5"):
        raise RuntimeError(f"{base}.h5 exists but h5py is not importable; convert it to .npz (write_db_npz)")
    raise RuntimeError(f"Cannot find sketch database {base}.h5 / .npz")


def write_db_npz(prefix: str, names, kmers, sketches, random_table=None, random_clusters=None,
                 codon_phased=False, lengths=None, base_freq=None, random_centroids=None, use_rc=True) -> str:
    """Write ``<prefix>/<basename>.npz`` — the .npz mirror of the reference HDF5 schema (web.py:14-61):
    ``names`` [n], ``kmers`` [K], ``sketchsize64``, ``bbits``, ``sketches`` uint64 [n][K][sketchsize64*bbits],
    ``codon_phased`` and, optionally, ``lengths`` [n] / ``base_freq`` [n][4] (the per-sample attrs) and the
    ``/random`` group: the random-match table [C][C][K], per-sample cluster ids, base-composition centroids."""
    os.makedirs(prefix, exist_ok=True)
    sketches = np.ascontiguousarray(sketches, dtype=np.uint64)
    path = os.path.join(prefix, os.path.basename(prefix) + ".npz")
    extra = {}
    if random_table is not None:
        extra = dict(random_table=np.asarray(random_table, dtype=np.float32),
                     random_clusters=np.asarray(random_clusters, dtype=np.uint16))
        if random_centroids is not None:
            extra["random_centroids"] = np.asarray(random_centroids, dtype=np.float64)
    if lengths is not None:
        extra["lengths"] = np.asarray(lengths, dtype=np.float64)
    if base_freq is not None:
        extra["base_freq"] = np.asarray(base_freq, dtype=np.float64)
    np.savez(path, names=np.asarray(list(names)), kmers=np.asarray(kmers, dtype=np.int32),
             sketchsize64=np.int32(sketches.shape[2] // BBITS), bbits=np.int32(BBITS), sketches=sketches,
             codon_phased=np.bool_(codon_phased), use_rc=np.bool_(use_rc), **extra)
    return path


def _text(v) -> str:
    return v.decode() if isinstance(v, (bytes, np.bytes_)) else str(v)


def _read_h5_random(f, names, kmers):
    """The ``/random`` group pp-sketchlib's ``addRandom`` writes (PopPUNK/sketchlib.py:437-473 calls it; 256-322
    copy it between files).  [UPSTREAM-RECALL] object names — pp-sketchlib's source is not in the reference tree:
    ``table_keys`` (sample names) / ``table_values`` (uint16 cluster per sample), ``matches_keys`` (k-mer lengths) /
    ``matches_values`` (one C x C matrix of expected random Jaccards per k, concatenated), ``centroids`` (C x 4
    base frequencies), attrs ``k_min``, ``k_max``, ``use_rc``.  Returns ``(table [C][C][Kdb], clusters [n],
    centroids [C][4] or None, use_rc)`` or ``None`` when the group is absent or does not have this shape."""
    if "random" not in f:
        return None
    rnd = f["random"]
    try:
        keys = [_text(s) for s in rnd["table_keys"][:]]
        vals = np.asarray(rnd["table_values"][:], dtype=np.int64)
        mk = [int(k) for k in np.asarray(rnd["matches_keys"][:]).ravel()]
        mv = np.asarray(rnd["matches_values"][:], dtype=np.float64).ravel()
    except KeyError:
        return None
    if len(mk) == 0 or len(mv) % len(mk):
        return None
    C = int(round((len(mv) // len(mk)) ** 0.5))
    if C < 1 or C * C * len(mk) != len(mv) or (len(vals) and int(vals.max()) >= C):
        return None
    per_k = {k: mv[t * C * C:(t + 1) * C * C].reshape(C, C) for t, k in enumerate(mk)}
    missing = [int(k) for k in kmers if int(k) not in per_k]
    if missing:
        raise RuntimeError(f"random match chances in the database do not cover k = {missing}")
    table = np.stack([per_k[int(k)] for k in kmers], axis=-1).astype(np.float32)
    lut = dict(zip(keys, vals))
    clusters = np.array([lut.get(n, int(NOT_IN_TABLE)) for n in names], dtype=np.uint16)
    centroids = None
    if "centroids" in rnd:
        c = np.asarray(rnd["centroids"][:], dtype=np.float64)
        if c.size == C * 4:
            centroids = c.reshape(C, 4)
    use_rc = bool(rnd.attrs["use_rc"]) if "use_rc" in getattr(rnd, "attrs", {}) else True
    return table, clusters, centroids, use_rc


def read_db(prefix: str, names: Optional[Sequence[str]] = None) -> SketchDB:
    """Read the sketches of ``names`` (default: every sample) — only those: PopPUNK asks for list-ordered subsets
    (assign.py:476-480, 525-526; a 2-sample probe at sketchlib.py:543-550).  The result's sample order is ``names``."""
    path = _db_file(prefix)
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        all_names = [str(s) for s in z["names"]]
        sel = None
        if names is not None:
            lut = {n: i for i, n in enumerate(all_names)}
            try:
                sel = np.fromiter((lut[n] for n in names), dtype=np.int64, count=len(names))
            except KeyError as e:
                raise RuntimeError(f"Sample {e.args[0]} not found in sketch database") from None

        def per_sample(key):
            if key not in z:
                return None
            a = z[key]
            return a if sel is None else a[sel]

        db = SketchDB(all_names if sel is None else list(names), z["kmers"].astype(np.int32), int(z["sketchsize64"]),
                      int(z["bbits"]), per_sample("sketches"), bool(z["codon_phased"]) if "codon_phased" in z else False,
                      z["random_table"] if "random_table" in z else None, per_sample("random_clusters"),
                      per_sample("lengths"), per_sample("base_freq"),
                      z["random_centroids"] if "random_centroids" in z else None,
                      bool(z["use_rc"]) if "use_rc" in z else True)
    else:  # HDF5 written by pp-sketchlib (read only what the distance path needs)
        with h5py.File(path, "r") as f:
            grp = f["sketches"]
            all_names = list(grp.keys()) if names is None else list(names)
            for n in all_names:
                if n not in grp:
                    raise RuntimeError(f"Sample {n} not found in sketch database")
            first = grp[all_names[0]]
            kmers = np.sort(np.asarray(first.attrs["kmers"], dtype=np.int32))
            ss64, bbits = int(first.attrs["sketchsize64"]), int(first.attrs["bbits"])
            sk = np.empty((len(all_names), len(kmers), ss64 * bbits), dtype=np.uint64)
            lengths = np.full(len(all_names), np.nan)
            base_freq = np.full((len(all_names), 4), np.nan)
            kstr = [str(int(k)) for k in kmers]
            for i, n in enumerate(all_names):
                g = grp[n]
                for t, ks in enumerate(kstr):
                    sk[i, t] = g[ks][:]
                if "length" in g.attrs:
                    lengths[i] = float(g.attrs["length"])
                if "base_freq" in g.attrs:
                    bf = np.asarray(g.attrs["base_freq"], dtype=np.float64).ravel()
                    if bf.size == 4:
                        base_freq[i] = bf
            codon = bool(grp.attrs["codon_phased"]) if "codon_phased" in grp.attrs else False
            db = SketchDB(all_names, kmers, ss64, bbits, sk, codon,
                          lengths=None if np.isnan(lengths).any() else lengths,
                          base_freq=None if np.isnan(base_freq).any() else base_freq)
            rnd = _read_h5_random(f, all_names, kmers)
            if rnd is not None:
                db.random_table, db.random_clusters, db.random_centroids, db.use_rc = rnd
            elif "use_rc" in grp.attrs:
                db.use_rc = bool(grp.attrs["use_rc"])
    if db.bbits != BBITS:
        raise RuntimeError(f"bbits = {db.bbits} is not supported (pp-sketchlib writes 14)")
    return db


def getSketchSize(dbPrefix):
    """Sketch size in bins (PopPUNK/sketchlib.py:109-146 returns sketchsize64 and codon_phased)."""
    db = _read_params(dbPrefix)
    return db.sketchsize64, db.codon_phased


def getKmersFromReferenceDatabase(dbPrefix):
    """PopPUNK/sketchlib.py:148-168."""
    return np.asarray(_read_params(dbPrefix).kmers)


def readDBParams(dbPrefix):
    """PopPUNK/sketchlib.py:170-195 -> (kmers, sketch_sizes, codon_phased)."""
    db = _read_params(dbPrefix)
    return np.asarray(db.kmers), db.sketchsize64, db.codon_phased


def _sample_names(prefix: str) -> List[str]:
    path = _db_file(prefix)
    if path.endswith(".npz"):
        return [str(s) for s in np.load(path, allow_pickle=False)["names"]]
    with h5py.File(path, "r") as f:
        return list(f["sketches"].keys())


def _read_params(prefix: str) -> SketchDB:
    """Parameters only: one sample is read, not the whole database."""
    return read_db(prefix, _sample_names(prefix)[:1])


def getSeqsInDb(dbname):
    """PopPUNK/sketchlib.py:197-214."""
    return _sample_names(os.path.dirname(dbname) if dbname.endswith((".h5", ".npz")) else dbname)


# --------------------------------------------------------------------------------------------
# random-match chances (a5)
# --------------------------------------------------------------------------------------------
def random_match_fallback(ref_lengths, qry_lengths, klist, use_rc=True, max_clusters=MAX_FALLBACK_CLUSTERS):
    """Closed-form random-match chances for a database without a ``/random`` table ("Could not find random match
    chances in database, calculating assuming equal base frequencies", docs/query_assignment.rst:110).

    docs/sketching.rst:107-118: ``r = 1 - (1 - 2*4^-k)^l`` (the factor 2 only when both strands are used; the
    docs print the exponent as ``-l``, which would make r negative — the chance that a given k-mer occurs among
    the l k-mers of a random genome is meant), and for a pair ``J_r = r1 r2 / (r1 + r2 - r1 r2)``
    (``r^2 / (2r - r^2)`` for equal lengths).

    The engine looks corrections up per (cluster, cluster, k), so genomes are classed by LENGTH: one class per
    distinct length when there are at most ``max_clusters`` of them (then this is exact per genome), otherwise
    ``max_clusters`` equal-count classes of log-length, each represented by its geometric-mean length (bacterial
    collections: classes ~1 % wide, |dJ_r| < 1e-3 at the smallest k, far less above it).
    Returns ``(table float32 [C][C][K], ref_cluster uint16, qry_cluster uint16 or None)``."""
    ref_lengths = np.asarray(ref_lengths, dtype=np.float64)
    n_ref = len(ref_lengths)
    lens = ref_lengths if qry_lengths is None else np.concatenate([ref_lengths, np.asarray(qry_lengths, dtype=np.float64)])
    if len(lens) == 0 or not np.isfinite(lens).all() or (lens <= 0).any():
        raise RuntimeError("genome lengths are needed for the closed-form random match chances")
    uniq = np.unique(lens)
    if len(uniq) <= max_clusters:
        rep = uniq
        cl = np.searchsorted(uniq, lens)
    else:
        order = np.argsort(lens, kind="stable")
        cl = np.empty(len(lens), dtype=np.int64)
        cl[order] = (np.arange(len(lens)) * max_clusters) // len(lens)
        loglen = np.log(lens)
        rep = np.exp(np.array([loglen[cl == c].mean() for c in range(max_clusters)]))
    k = np.asarray(klist, dtype=np.float64)
    f = 2.0 if use_rc else 1.0
    r = -np.expm1(rep[:, None] * np.log1p(-f * 4.0 ** (-k[None, :])))        # [C][K]: 1 - (1 - f 4^-k)^l
    r1, r2 = r[:, None, :], r[None, :, :]
    den = r1 + r2 - r1 * r2
    table = np.where(den > 0, r1 * r2 / np.where(den > 0, den, 1.0), 0.0).astype(np.float32)
    cl = cl.astype(np.uint16)
    return np.ascontiguousarray(table), cl[:n_ref], (None if qry_lengths is None else cl[n_ref:])


def _nearest_centroid(base_freq, centroids) -> np.ndarray:
    """Cluster of samples the table does not list (queries sketched with calc_random=False, assign.py:296): the
    closest base-composition centroid, as pp-sketchlib's RandomMC does [UPSTREAM-RECALL]."""
    d = ((np.asarray(base_freq)[:, None, :] - np.asarray(centroids)[None, :, :]) ** 2).sum(axis=-1)
    return d.argmin(axis=1).astype(np.uint16)


def _resolve_clusters(db: SketchDB, idx: np.ndarray, table_db: SketchDB) -> np.ndarray:
    """Cluster ids (into ``table_db``'s table) of samples ``idx`` of ``db``."""
    n = len(idx)
    cl = np.full(n, NOT_IN_TABLE, dtype=np.uint16)
    if db is table_db and db.random_clusters is not None:
        cl = np.asarray(db.random_clusters, dtype=np.uint16)[idx].copy()
    unknown = cl == NOT_IN_TABLE
    if unknown.any():
        if table_db.random_centroids is not None and db.base_freq is not None:
            cl[unknown] = _nearest_centroid(db.base_freq[idx][unknown], table_db.random_centroids)
        else:
            cl[unknown] = 0   # no composition data to place them with: the table's first cluster
    C = table_db.random_table.shape[0]
    if n and int(cl.max()) >= C:
        raise RuntimeError("random match cluster id out of range: the database's /random group is inconsistent")
    return cl


def random_match_setup(rdb: SketchDB, ridx, qdb: Optional[SketchDB], qidx, klist):
    """``(table [C][C][K], ref_cluster, qry_cluster)`` for a call with ``random_correct=True`` — the REFERENCE
    database's table (query DBs are built with calc_random=False, assign.py:296), else the closed form."""
    kidx = rdb.k_index(klist)
    if rdb.random_table is not None:
        table = np.ascontiguousarray(np.asarray(rdb.random_table)[:, :, kidx], dtype=np.float32)
        rcl = _resolve_clusters(rdb, ridx, rdb)
        qcl = None if qdb is None else _resolve_clusters(qdb, qidx, rdb)
        return table, rcl, qcl
    sys.stderr.write("Could not find random match chances in database, calculating assuming equal base frequencies\n")
    if rdb.lengths is None or (qdb is not None and qdb.lengths is None):
        sys.stderr.write("poppunk_b200: the database holds no genome lengths either: NO random-match correction applied\n")
        return None, None, None
    return random_match_fallback(rdb.lengths[ridx], None if qdb is None else qdb.lengths[qidx], klist, rdb.use_rc)


# --------------------------------------------------------------------------------------------
# the native entry PopPUNK calls (pp_sketchlib.queryDatabase)
# --------------------------------------------------------------------------------------------
_DEGENERATE_MSG = ("Fitting k-mer gradient failed for {n} pair(s): fewer than two k-mer lengths with Jaccard >= 5/s; "
                   "returned as (0, 0).\nCheck for low quality genomes, or use a wider k-mer range\n")


def query_arrays(ref: np.ndarray, qry: Optional[np.ndarray], klist, table=None, ref_cluster=None, qry_cluster=None,
                 jaccard: bool = False, device_id: int = 0, out: Optional[np.ndarray] = None):
    """What :func:`pp_queryDatabase` does once the sketches are in memory: one host-buffer call on every visible
    GPU; the result is a NumPy array the caller owns (library pool block — see ``engine.host_result``).
    Returns ``(distances, n_degenerate)``."""
    out, _, ndeg = engine.query_host(ref, qry, klist, table, ref_cluster, qry_cluster,
                                     out_mode=OUT_JACCARD if jaccard else OUT_DISTS, out=out,
                                     devices=engine.visible_devices(device_id))
    return out, ndeg


def pp_queryDatabase(ref_db_name, query_db_name, rList, qList, klist, random_correct=True, jaccard=False,
                     num_threads=1, use_gpu=False, device_id=0):
    """Same positional signature as ``pp_sketchlib.queryDatabase`` (test/test-update-gpu.py:85-86).

    ``ref_db_name`` / ``query_db_name`` are ``<prefix>/<basename>`` paths without extension.  When both are the
    same database and ``rList == qList`` the result is the condensed self matrix, otherwise the query-major
    rectangle.  ``jaccard=True`` returns per-k Jaccards ``[n_pairs, K]`` (sketchlib.py:547-566)."""
    del num_threads, use_gpu  # accepted for signature compatibility; the engine is GPU-only
    klist = np.asarray(klist, dtype=np.int32)
    if klist.ndim != 1 or len(klist) < 1 or (np.diff(klist) <= 0).any():
        raise RuntimeError("klist must be ascending k-mer lengths")
    ref_prefix = os.path.dirname(ref_db_name)
    qry_prefix = os.path.dirname(query_db_name)
    same_db = os.path.abspath(ref_db_name) == os.path.abspath(query_db_name)
    rList, qList = list(rList), list(qList)
    self_mode = same_db and rList == qList
    # only the samples this call names are read (list order = row order, utils.py:220-226)
    rdb = read_db(ref_prefix, rList)
    ridx = np.arange(len(rList), dtype=np.int64)
    qdb = qidx = None
    if not self_mode:
        qdb = read_db(ref_prefix if same_db else qry_prefix, qList)
        qidx = np.arange(len(qList), dtype=np.int64)
        if (qdb.sketchsize64, qdb.bbits) != (rdb.sketchsize64, rdb.bbits):
            raise RuntimeError("Query and reference databases have different sketch sizes")
        if same_db:   # one database: its table and cluster ids serve both sides
            qdb.random_table, qdb.random_centroids = rdb.random_table, rdb.random_centroids
    kidx_r = rdb.k_index(klist)
    ref = np.ascontiguousarray(rdb.sketches[:, kidx_r])
    qry = None if self_mode else np.ascontiguousarray(qdb.sketches[:, qdb.k_index(klist)])
    table = rcl = qcl = None
    if random_correct:
        if not self_mode and same_db and rdb.random_table is not None:
            table = np.ascontiguousarray(np.asarray(rdb.random_table)[:, :, kidx_r], dtype=np.float32)
            rcl, qcl = _resolve_clusters(rdb, ridx, rdb), _resolve_clusters(qdb, qidx, qdb)
        else:
            table, rcl, qcl = random_match_setup(rdb, ridx, qdb, qidx, klist)
    out, ndeg = query_arrays(ref, qry, klist, table, rcl, qcl, jaccard=jaccard, device_id=device_id)
    if ndeg:
        sys.stderr.write(_DEGENERATE_MSG.format(n=ndeg))
    return out


def fitKmerCurve(pairwise, klist, jacobian=None):
    """PopPUNK/sketchlib.py:635-670: fit ``log pr = log(1-a) + k log(1-c)`` with both parameters bounded above by
    0, return ``(core, accessory)``.  The reference hands the problem to ``scipy.optimize.least_squares``; a linear
    least-squares problem with two upper bounds has the closed form used here (unconstrained fit; if a parameter
    ends up positive it is fixed at 0 and the other refitted) — same optimum (tests/test_oracle.py)."""
    del jacobian
    k = np.asarray(klist, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = np.log(np.asarray(pairwise, dtype=np.float64))
    if not np.isfinite(y).all():
        sys.stderr.write("Fitting k-mer curve failed: Residuals are not finite in the initial point."
                         "\nWith k-mer match values " +
                         np.array2string(np.asarray(pairwise), precision=4, separator=',', suppress_small=True) +
                         "\nCheck for low quality input genomes\n")
        return np.array([0.0, 0.0])
    kb, yb = k.mean(), y.mean()
    sxx = ((k - kb) ** 2).sum()
    slope = ((k - kb) * (y - yb)).sum() / sxx if sxx > 0 else 0.0
    icpt = yb - slope * kb
    if slope > 0 or icpt > 0:
        # a bound is active: the optimum of this convex problem lies on an edge or the corner of the feasible set
        cands = [(min(0.0, yb), 0.0), (0.0, min(0.0, (k * y).sum() / (k * k).sum())), (0.0, 0.0)]
        icpt, slope = min(cands, key=lambda p: ((y - (p[0] + p[1] * k)) ** 2).sum())
    return np.array([1.0 - np.exp(slope), 1.0 - np.exp(icpt)])


def plot_fit(klist, raw_matching, raw_fit, corrected_matching, corrected_fit, out_prefix, title):
    """PopPUNK/plot.py ``plot_fit``: the example-fit figure of ``--plot-fit``.  Drawn with matplotlib when it is
    importable; the numbers behind the figure are always written to ``<out_prefix>.tsv``."""
    with open(out_prefix + ".tsv", "w") as fh:
        fh.write(f"# {title}\n# raw fit (core, accessory): {raw_fit[0]:.6g}\t{raw_fit[1]:.6g}\n"
                 f"# corrected fit (core, accessory): {corrected_fit[0]:.6g}\t{corrected_fit[1]:.6g}\nk\traw\tcorrected\n")
        for k, r, c in zip(klist, raw_matching, corrected_matching):
            fh.write(f"{int(k)}\t{float(r):.6g}\t{float(c):.6g}\n")
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return
    k_fit = np.linspace(0, int(klist[-1]), num=100)
    fig, ax = plt.subplots()
    ax.set_yscale("log")
    ax.plot(klist, raw_matching, "o", label="Raw matching k-mers")
    ax.plot(k_fit, (1 - raw_fit[1]) * (1 - raw_fit[0]) ** k_fit, label="Fit to raw matches")
    ax.plot(klist, corrected_matching, "x", label="Corrected matching k-mers")
    ax.plot(k_fit, (1 - corrected_fit[1]) * (1 - corrected_fit[0]) ** k_fit, label="Fit to corrected matches")
    ax.set_xlabel("k-mer length")
    ax.set_ylabel("Proportion of matches")
    ax.set_title(title)
    ax.legend(loc="upper right")
    fig.savefig(out_prefix + ".pdf", bbox_inches="tight")
    plt.close(fig)


def queryDatabase(rNames, qNames, dbPrefix, queryPrefix, klist, self=True, number_plot_fits=0,
                  threads=1, use_gpu=False, deviceid=0):
    """Core and accessory distances between query sequences and a sketched database.

    Argument meaning, row order (``PopPUNK.utils.iterDistRows``) and error behaviour follow
    PopPUNK/sketchlib.py:475-632.  Returns float32 ``(n_pairs, 2)``, C-contiguous: column 0 core, column 1
    accessory."""
    ref_db = dbPrefix + "/" + os.path.basename(dbPrefix)
    klist = np.asarray(klist)
    if self:
        if dbPrefix != queryPrefix:
            raise RuntimeError("Must use same db for self query")  # sketchlib.py:523-524
        qNames = rNames
        distMat = pp_queryDatabase(ref_db, ref_db, rNames, rNames, klist, True, False, threads, use_gpu, deviceid)
        # option to plot core/accessory fits: per-k probes of random pairs (sketchlib.py:540-573)
        for plot_idx in range(max(0, number_plot_fits)):
            example = sample(list(rNames), k=2)
            raw = pp_queryDatabase(ref_db, ref_db, [example[0]], [example[1]], klist, False, True, threads, False)
            corrected = pp_queryDatabase(ref_db, ref_db, [example[0]], [example[1]], klist, True, True, threads, False)
            plot_fit(klist, raw[0], fitKmerCurve(raw[0], klist), corrected[0], fitKmerCurve(corrected[0], klist),
                     ref_db + "_fit_example_" + str(plot_idx + 1),
                     "Example fit " + str(plot_idx + 1) + " - " + example[0] + " vs. " + example[1])
    else:
        duplicated = set(rNames).intersection(set(qNames))
        if len(duplicated) > 0:  # sketchlib.py:575-580
            sys.stderr.write("Sample names in query are contained in reference database:\n")
            sys.stderr.write("\n".join(duplicated))
            sys.stderr.write("Unique names are required!\n")
            sys.exit(1)
        query_db = queryPrefix + "/" + os.path.basename(queryPrefix)
        distMat = pp_queryDatabase(ref_db, query_db, rNames, qNames, klist, True, False, threads, use_gpu, deviceid)
        if number_plot_fits > 0:  # sketchlib.py:596-630 (row plot_idx of the example rectangle, as the reference indexes it)
            ref_examples = sample(list(rNames), k=number_plot_fits)
            query_examples = sample(list(qNames), k=number_plot_fits)
            raw = pp_queryDatabase(ref_db, query_db, ref_examples, query_examples, klist, False, True, threads, False)
            corrected = pp_queryDatabase(ref_db, query_db, ref_examples, query_examples, klist, True, True, threads, False)
            for plot_idx in range(number_plot_fits):
                plot_fit(klist, raw[plot_idx], fitKmerCurve(raw[plot_idx], klist), corrected[plot_idx],
                         fitKmerCurve(corrected[plot_idx], klist),
                         os.path.join(os.path.dirname(queryPrefix),
                                      os.path.basename(queryPrefix) + "_fit_example_" + str(plot_idx + 1)),
                         "Example fit " + str(plot_idx + 1) + " - " + ref_examples[plot_idx] + " vs. " +
                         query_examples[plot_idx])
    return distMat
