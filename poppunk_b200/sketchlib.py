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
* a pair whose fit has fewer than two usable k is returned as (0, 0) and counted, with the reference's
  warning text, instead of aborting the process (docs/troubleshooting.rst:176-191);
* sketch databases are read from ``<prefix>/<basename>.h5`` when ``h5py`` is importable (schema:
  PopPUNK/web.py:14-61) and otherwise from ``<prefix>/<basename>.npz`` (same content, documented in
  :func:`write_db_npz`) — h5py is not in this image.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import engine
from ._lib import BBITS, OUT_DISTS, OUT_JACCARD

try:  # optional: not present in the build image
    import h5py  # type: ignore
except Exception:  # pragma: no cover
    h5py = None


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
    random_table: Optional[np.ndarray] = None     # float32 [C][C][Kdb]
    random_clusters: Optional[np.ndarray] = None  # uint16 [n]

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
                 codon_phased=False) -> str:
    """Write ``<prefix>/<basename>.npz`` — the .npz mirror of the reference HDF5 schema (web.py:14-61):
    ``names`` [n], ``kmers`` [K], ``sketchsize64``, ``bbits``, ``sketches`` uint64 [n][K][sketchsize64*bbits],
    ``codon_phased`` and, optionally, the random-match table [C][C][K] + per-sample cluster ids."""
    os.makedirs(prefix, exist_ok=True)
    sketches = np.ascontiguousarray(sketches, dtype=np.uint64)
    path = os.path.join(prefix, os.path.basename(prefix) + ".npz")
    extra = {}
    if random_table is not None:
        extra = dict(random_table=np.asarray(random_table, dtype=np.float32),
                     random_clusters=np.asarray(random_clusters, dtype=np.uint16))
    np.savez(path, names=np.asarray(list(names)), kmers=np.asarray(kmers, dtype=np.int32),
             sketchsize64=np.int32(sketches.shape[2] // BBITS), bbits=np.int32(BBITS), sketches=sketches,
             codon_phased=np.bool_(codon_phased), **extra)
    return path


def read_db(prefix: str, names: Optional[Sequence[str]] = None) -> SketchDB:
    path = _db_file(prefix)
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        db = SketchDB([str(s) for s in z["names"]], z["kmers"].astype(np.int32), int(z["sketchsize64"]),
                      int(z["bbits"]), z["sketches"], bool(z["codon_phased"]) if "codon_phased" in z else False,
                      z["random_table"] if "random_table" in z else None,
                      z["random_clusters"] if "random_clusters" in z else None)
    else:  # HDF5 written by pp-sketchlib (read only what the distance path needs)
        with h5py.File(path, "r") as f:
            grp = f["sketches"]
            all_names = list(grp.keys()) if names is None else list(names)
            first = grp[all_names[0]]
            kmers = np.sort(np.asarray(first.attrs["kmers"], dtype=np.int32))
            ss64, bbits = int(first.attrs["sketchsize64"]), int(first.attrs["bbits"])
            sk = np.empty((len(all_names), len(kmers), ss64 * bbits), dtype=np.uint64)
            for i, n in enumerate(all_names):
                for t, k in enumerate(kmers):
                    sk[i, t] = grp[n][str(int(k))][:]
            codon = bool(grp.attrs["codon_phased"]) if "codon_phased" in grp.attrs else False
            db = SketchDB(all_names, kmers, ss64, bbits, sk, codon)
    if db.bbits != BBITS:
        raise RuntimeError(f"bbits = {db.bbits} is not supported (pp-sketchlib writes 14)")
    return db


def getSketchSize(dbPrefix):
    """Sketch size in bins (PopPUNK/sketchlib.py:109-146 returns sketchsize64 and codon_phased)."""
    db = read_db(dbPrefix)
    return db.sketchsize64, db.codon_phased


def getKmersFromReferenceDatabase(dbPrefix):
    """PopPUNK/sketchlib.py:148-168."""
    return np.asarray(read_db(dbPrefix).kmers)


def readDBParams(dbPrefix):
    """PopPUNK/sketchlib.py:170-195 -> (kmers, sketch_sizes, codon_phased)."""
    db = read_db(dbPrefix)
    return np.asarray(db.kmers), db.sketchsize64, db.codon_phased


def getSeqsInDb(dbname):
    """PopPUNK/sketchlib.py:197-214."""
    return read_db(os.path.dirname(dbname) if dbname.endswith((".h5", ".npz")) else dbname).names


# --------------------------------------------------------------------------------------------
# the native entry PopPUNK calls (pp_sketchlib.queryDatabase)
# --------------------------------------------------------------------------------------------
_DEGENERATE_MSG = ("Fitting k-mer gradient failed for {n} pair(s): fewer than two k-mer lengths with Jaccard >= 5/s; "
                   "returned as (0, 0).\nCheck for low quality genomes, or use a wider k-mer range\n")


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
    rdb = read_db(ref_prefix, None)
    self_mode = (os.path.abspath(ref_db_name) == os.path.abspath(query_db_name)) and list(rList) == list(qList)
    qdb = rdb if os.path.abspath(ref_db_name) == os.path.abspath(query_db_name) else read_db(qry_prefix, None)
    if (qdb.sketchsize64, qdb.bbits) != (rdb.sketchsize64, rdb.bbits):
        raise RuntimeError("Query and reference databases have different sketch sizes")
    ridx, kidx_r = rdb.index_of(rList), rdb.k_index(klist)
    ref = np.ascontiguousarray(rdb.sketches[ridx][:, kidx_r])
    qry = None
    if not self_mode:
        qry = np.ascontiguousarray(qdb.sketches[qdb.index_of(qList)][:, qdb.k_index(klist)])
    table = rcl = qcl = None
    if random_correct:
        # query DBs are built with calc_random=False (assign.py:296): the REF database's table is used
        if rdb.random_table is None:
            # the reference falls back to a closed-form estimate here (docs/query_assignment.rst:110);
            # that formula lives in pp-sketchlib and is not restated: no correction is applied instead
            sys.stderr.write("Could not find random match chances in database, "
                             "no random-match correction applied\n")
        else:
            table = np.ascontiguousarray(rdb.random_table[:, :, kidx_r])
            rcl = rdb.random_clusters[ridx]
            if not self_mode:
                qcl = (qdb.random_clusters[qdb.index_of(qList)] if qdb.random_clusters is not None
                       else np.zeros(len(qList), dtype=np.uint16))
    out, _, ndeg = engine.query_host(ref, qry, klist, table, rcl, qcl,
                                     out_mode=OUT_JACCARD if jaccard else OUT_DISTS, device_id=device_id)
    if ndeg:
        sys.stderr.write(_DEGENERATE_MSG.format(n=ndeg))
    return out


def queryDatabase(rNames, qNames, dbPrefix, queryPrefix, klist, self=True, number_plot_fits=0,
                  threads=1, use_gpu=False, deviceid=0):
    """Core and accessory distances between query sequences and a sketched database.

    Argument meaning, row order (``PopPUNK.utils.iterDistRows``) and error behaviour follow
    PopPUNK/sketchlib.py:475-632.  Returns float32 ``(n_pairs, 2)``, C-contiguous: column 0 core, column 1
    accessory."""
    ref_db = dbPrefix + "/" + os.path.basename(dbPrefix)
    if self:
        if dbPrefix != queryPrefix:
            raise RuntimeError("Must use same db for self query")  # sketchlib.py:523-524
        qNames = rNames
        distMat = pp_queryDatabase(ref_db, ref_db, rNames, rNames, klist, True, False, threads, use_gpu, deviceid)
    else:
        duplicated = set(rNames).intersection(set(qNames))
        if len(duplicated) > 0:  # sketchlib.py:575-580
            sys.stderr.write("Sample names in query are contained in reference database:\n")
            sys.stderr.write("\n".join(duplicated))
            sys.stderr.write("Unique names are required!\n")
            sys.exit(1)
        query_db = queryPrefix + "/" + os.path.basename(queryPrefix)
        distMat = pp_queryDatabase(ref_db, query_db, rNames, qNames, klist, True, False, threads, use_gpu, deviceid)
    if number_plot_fits > 0:
        # the reference re-queries per-k Jaccards and plots them with matplotlib (sketchlib.py:540-573,
        # 596-630); plotting is outside this engine — the per-k probe is `pp_queryDatabase(..., jaccard=True)`.
        sys.stderr.write("poppunk_b200: --plot-fit plots are not produced by this engine\n")
    return distMat
