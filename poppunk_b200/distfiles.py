"""The data formats either side of the distance path: how PopPUNK stores and indexes the
(n_pairs, 2) array ``queryDatabase`` returns.  Host-side file and index conventions only — no arithmetic.

    storePickle / readPickle     <prefix>.pkl holding [rlist, qlist, self] + <prefix>.npy   (PopPUNK/utils.py:135-197)
    iterDistRows / listDistInts  which (ref, query) pair each output row is                 (PopPUNK/utils.py:199-261)

Same names, arguments and error behaviour as the reference, so a caller can switch imports; the row order is the one
the engine writes (include/ppb.h: condensed upper triangle, or query-major rectangle).
"""
from __future__ import annotations

import pickle
import sys
from itertools import combinations, product

import numpy as np


def storePickle(rlist, qlist, self, X, pklName):
    """Write the name lists + self flag to ``pklName.pkl`` and, when ``X`` is an array, the distances to
    ``pklName.npy`` — the pair of files ``poppunk --fit-model`` reads back (PopPUNK/utils.py:135-157)."""
    with open(f"{pklName}.pkl", "wb") as fh:
        pickle.dump([rlist, qlist, self], fh)
    if isinstance(X, np.ndarray):
        np.save(f"{pklName}.npy", X)


def readPickle(pklName, enforce_self=False, distances=True):
    """Inverse of :func:`storePickle`: ``(rlist, qlist, self, X)``; ``X`` is None unless ``distances``.
    ``enforce_self`` rejects anything but a complete all-vs-all set the way the reference does: message on stderr,
    exit status 1 (PopPUNK/utils.py:160-197)."""
    with open(f"{pklName}.pkl", "rb") as fh:
        rlist, qlist, self = pickle.load(fh)
    if enforce_self and not (self and rlist == qlist):
        sys.stderr.write(f"Old distances {pklName}.npy not complete\n")
        sys.exit(1)
    return rlist, qlist, self, (np.load(f"{pklName}.npy") if distances else None)


def _check_self(refSeqs, querySeqs):
    if refSeqs != querySeqs:
        raise RuntimeError("refSeqs must equal querySeqs for db building (self = true)")


def listDistInts(refSeqs, querySeqs, self=True):
    """Index pair of every distance row, in row order (PopPUNK/utils.py:229-261): self -> ``(j, i)`` for every i < j,
    i the slow index; otherwise ``(ref index, query index)`` with the query the slow index."""
    if self:
        _check_self(refSeqs, querySeqs)
        return ((j, i) for i, j in combinations(range(len(refSeqs)), 2))
    return ((r, q) for q, r in product(range(len(querySeqs)), range(len(refSeqs))))


def iterDistRows(refSeqs, querySeqs, self=True):
    """Name pair of every distance row, in row order (PopPUNK/utils.py:199-226)."""
    if self:
        _check_self(refSeqs, querySeqs)
        return ((refSeqs[j], refSeqs[i]) for j, i in listDistInts(refSeqs, querySeqs, True))
    return ((refSeqs[r], querySeqs[q]) for r, q in listDistInts(refSeqs, querySeqs, False))
