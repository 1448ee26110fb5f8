"""The data formats either side of the distance path: how PopPUNK stores and indexes
the (n_pairs, 2) array ``queryDatabase`` returns.  Host-side file and index conventions only — no arithmetic.

    storePickle / readPickle     PopPUNK/utils.py:135-197   <prefix>.dists.pkl ([rlist, qlist, self]) + .npy
    iterDistRows / listDistInts  PopPUNK/utils.py:199-261   which (ref, query) pair each output row is

Same names, arguments and error behaviour as the reference, so a caller can switch imports.
"""
from __future__ import annotations

import pickle
import sys

import numpy as np


def storePickle(rlist, qlist, self, X, pklName):
    """Saves core and accessory distances in a .npy file, names in a .pkl (PopPUNK/utils.py:135-157)."""
    with open(pklName + ".pkl", "wb") as pickle_file:
        pickle.dump([rlist, qlist, self], pickle_file)
    if isinstance(X, np.ndarray):
        np.save(pklName + ".npy", X)


def readPickle(pklName, enforce_self=False, distances=True):
    """Loads what :func:`storePickle` saved (PopPUNK/utils.py:160-197): ``(rlist, qlist, self, X)``."""
    with open(pklName + ".pkl", "rb") as pickle_file:
        rlist, qlist, self = pickle.load(pickle_file)
        if enforce_self and (not self or rlist != qlist):
            sys.stderr.write("Old distances " + pklName + ".npy not complete\n")
            sys.exit(1)
    X = np.load(pklName + ".npy") if distances else None
    return rlist, qlist, self, X


def iterDistRows(refSeqs, querySeqs, self=True):
    """(ref, query) names of every distance row, in row order (PopPUNK/utils.py:199-226)."""
    if self:
        if refSeqs != querySeqs:
            raise RuntimeError("refSeqs must equal querySeqs for db building (self = true)")
        for i, ref in enumerate(refSeqs):
            for j in range(i + 1, len(refSeqs)):
                yield (refSeqs[j], ref)
    else:
        for query in querySeqs:
            for ref in refSeqs:
                yield (ref, query)


def listDistInts(refSeqs, querySeqs, self=True):
    """The same as indices (PopPUNK/utils.py:229-261): self -> (j, i) for i < j; else (ref index, query index)."""
    num_ref, num_query = len(refSeqs), len(querySeqs)
    if self:
        if refSeqs != querySeqs:
            raise RuntimeError("refSeqs must equal querySeqs for db building (self = true)")
        for i in range(num_ref):
            for j in range(i + 1, num_ref):
                yield (j, i)
    else:
        for i in range(num_query):
            for j in range(num_ref):
                yield (j, i)
