"""Seeded synthetic sketch generator (bit-sliced bindash sketches without a sketcher).

The reference obtains sketches from ``pp_sketchlib.constructDatabase`` (PopPUNK/sketchlib.py:410-422),
which is outside the hot path.  For parity tests and bench.py we need
sketch arrays with realistic, non-degenerate Jaccards: pure i.i.d. random signatures give
J ~ 2^-14 < 5/S for every pair, i.e. every regression is truncated away (docs/sketching.rst:161-165).

Model: ``n_roots`` independent ancestor signature tables, ``n_lineages`` lineage tables
derived from them (round-robin) and genomes derived from a lineage.  Pairs under one ancestor are related
(``E[J_k] ~ (1-a)(1-pi)^k``); pairs under different ancestors share only chance matches (J ~ 2^-14 < 5/S), so
their series is truncated away and they come out as degenerate (0, 0) pairs — both paths of the fit are exercised.  At each derivation step and for each k a bin keeps its parent's
signature with probability ``p_k = sqrt((1-a)(1-pi)^k)`` and is redrawn uniformly otherwise, so two
genomes with a common parent have ``E[J_k] ~ (1-a)(1-pi)^k`` — the relation PopPUNK fits
(PopPUNK/sketchlib.py:482).

Layout produced (the reference's HDF5 dataset layout, PopPUNK/web.py:14-61 and
test/json_sketch.txt): ``uint64 [n][K][W]``, ``W = sketchsize64 * 14``, word ``s*14 + b`` = bit ``b``
of the 14-bit signatures of bins ``64 s .. 64 s + 63`` (bin = bit position).
"""
from __future__ import annotations

import numpy as np

BBITS = 14


def bitslice(sig: np.ndarray) -> np.ndarray:
    """uint16 signatures ``[..., S]`` (values < 2**14) -> bit-sliced uint64 ``[..., S//64 * 14]``."""
    sig = np.ascontiguousarray(sig, dtype=np.uint16)
    S = sig.shape[-1]
    if S % 64:
        raise ValueError("number of bins must be a multiple of 64")
    lead = sig.shape[:-1]
    s64 = S // 64
    out = np.empty(lead + (s64, BBITS), dtype=np.uint64)
    blk = sig.reshape(lead + (s64, 64))
    for b in range(BBITS):
        bits = ((blk >> np.uint16(b)) & np.uint16(1)).astype(np.uint8)
        packed = np.packbits(bits, axis=-1, bitorder="little")  # 8 bytes, bin t -> bit t
        out[..., b] = np.ascontiguousarray(packed).view("<u8").reshape(lead + (s64,))
    return out.reshape(lead + (s64 * BBITS,))


def unslice(words: np.ndarray, sketchsize64: int) -> np.ndarray:
    """Inverse of :func:`bitslice`: uint64 ``[..., W]`` -> uint16 signatures ``[..., 64*sketchsize64]``."""
    words = np.ascontiguousarray(words, dtype="<u8")
    lead = words.shape[:-1]
    w = words.reshape(lead + (sketchsize64, BBITS))
    sig = np.zeros(lead + (sketchsize64, 64), dtype=np.uint16)
    for b in range(BBITS):
        plane = np.ascontiguousarray(w[..., b]).view(np.uint8).reshape(lead + (sketchsize64, 8))
        bits = np.unpackbits(plane, axis=-1, bitorder="little").astype(np.uint16)
        sig |= bits << np.uint16(b)
    return sig.reshape(lead + (sketchsize64 * 64,))


def _derive(rng, parent, p_keep):
    """Keep each parent bin with probability p_keep[k] (broadcast over bins), else redraw."""
    keep = rng.random(parent.shape, dtype=np.float32) < p_keep[..., None].astype(np.float32)
    fresh = rng.integers(0, 1 << BBITS, size=parent.shape, dtype=np.uint16)
    return np.where(keep, parent, fresh)


def synth_signatures(n, kmers, sketchsize64, seed=42, n_lineages=8,
                     pi_range=(0.001, 0.02), a_range=(0.01, 0.2), chunk=2048, sample_seed=0, n_roots=1):
    """Generator of ``(start, uint16 [m][K][S])`` signature chunks for ``n`` genomes.

    ``seed`` fixes the population (root + lineages); ``sample_seed`` fixes the genomes drawn from it, so
    a query set (``sample_seed=1``) is related to a reference set (``sample_seed=0``) of the same ``seed``.
    """
    kmers = np.asarray(kmers, dtype=np.float64)
    K = len(kmers)
    S = 64 * sketchsize64
    rng = np.random.default_rng(seed)
    root = rng.integers(0, 1 << BBITS, size=(K, S), dtype=np.uint16)

    def p_keep(m):
        pi = rng.uniform(*pi_range, size=m)
        a = rng.uniform(*a_range, size=m)
        return np.sqrt((1.0 - a)[:, None] * (1.0 - pi)[:, None] ** kmers[None, :])  # [m][K]

    if n_roots <= 1:
        parents = np.broadcast_to(root, (n_lineages, K, S))
    else:   # lineage l descends from ancestor l % n_roots; ancestors are independent
        roots = np.concatenate([root[None], rng.integers(0, 1 << BBITS, size=(n_roots - 1, K, S), dtype=np.uint16)])
        parents = roots[np.arange(n_lineages) % n_roots]
    lin = _derive(rng, parents, p_keep(n_lineages))
    rng = np.random.default_rng([seed, sample_seed])
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        which = rng.integers(0, n_lineages, size=m)
        yield start, _derive(rng, lin[which], p_keep(m))


def synth_sketches(n, kmers, sketchsize64, seed=42, **kw) -> np.ndarray:
    """Bit-sliced synthetic sketch array ``uint64 [n][K][W]`` (host, NumPy)."""
    K = len(kmers)
    out = np.empty((n, K, sketchsize64 * BBITS), dtype=np.uint64)
    for start, sig in synth_signatures(n, kmers, sketchsize64, seed=seed, **kw):
        out[start:start + sig.shape[0]] = bitslice(sig)
    return out


def random_match_table(kmers, n_clusters=3, genome_length=2.1e6, seed=42):
    """A small random-match table shaped like pp-sketchlib's RandomMC output.

    ``r = 1 - (1 - 2*4^-k)^(-l)``... the docs' formula (docs/sketching.rst:107-118) is written with a
    sign slip; the chance that a given k-mer occurs in a random genome of length ``l`` is
    ``r = 1 - (1 - 4^-k)^(2 l)`` (both strands) and ``J_r = r^2 / (2 r - r^2)``.  Clusters get slightly
    different effective lengths so the table is not constant.  Returns float32 ``[C][C][K]`` symmetric.
    """
    kmers = np.asarray(kmers, dtype=np.float64)
    rng = np.random.default_rng(seed)
    lengths = genome_length * rng.uniform(0.8, 1.25, size=n_clusters)
    r = -np.expm1(2.0 * lengths[:, None] * np.log1p(-(4.0 ** (-kmers[None, :]))))  # [C][K]
    r1 = r[:, None, :]
    r2 = r[None, :, :]
    den = r1 + r2 - r1 * r2
    jr = np.where(den > 0, (r1 * r2) / np.where(den > 0, den, 1.0), 0.0)
    return np.ascontiguousarray(jr, dtype=np.float32)


def synth_clusters(n, n_clusters=3, seed=42) -> np.ndarray:
    rng = np.random.default_rng(seed + 1)
    return rng.integers(0, n_clusters, size=n).astype(np.uint16)


def synth_sketches_torch(n, kmers, sketchsize64, seed=42, device="cuda", n_lineages=8,
                         pi_range=(0.001, 0.02), a_range=(0.01, 0.2), chunk=4096, n_roots=1, sample_seed=0):
    """Same population model as :func:`synth_sketches`, generated on ``device`` with torch (fast enough for
    N = 100k).  ``seed`` fixes the population; a non-zero ``sample_seed`` draws OTHER genomes from it (query sets
    related to a reference set of the same ``seed``).  Not bit-identical to the NumPy generator (different RNG) — bench.py uses it and hands the CPU
    baseline a device->host copy of the very same array.  Returns int64 ``[n][K][W]`` (uint64 bit patterns)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    K, S = len(kmers), 64 * sketchsize64
    km = torch.tensor(np.asarray(kmers, dtype=np.float64), device=dev)

    def p_keep(m):
        pi = torch.empty(m, device=dev, dtype=torch.float64).uniform_(*pi_range, generator=g)
        a = torch.empty(m, device=dev, dtype=torch.float64).uniform_(*a_range, generator=g)
        return torch.sqrt((1.0 - a)[:, None] * (1.0 - pi)[:, None] ** km[None, :]).float()

    def derive(parent, p):
        keep = torch.rand(parent.shape, device=dev, generator=g) < p[..., None]
        fresh = torch.randint(0, 1 << BBITS, parent.shape, device=dev, dtype=torch.int16, generator=g)
        return torch.where(keep, parent, fresh)

    root = torch.randint(0, 1 << BBITS, (K, S), device=dev, dtype=torch.int16, generator=g)
    if n_roots <= 1:
        parents = root.expand(n_lineages, K, S)
    else:
        more = torch.randint(0, 1 << BBITS, (n_roots - 1, K, S), device=dev, dtype=torch.int16, generator=g)
        parents = torch.cat([root[None], more])[torch.arange(n_lineages, device=dev) % n_roots]
    lin = derive(parents, p_keep(n_lineages))
    if sample_seed:
        g.manual_seed(1_000_003 * int(sample_seed) + seed)   # the population is fixed; these genomes are new draws from it
    weights = (torch.ones(64, dtype=torch.int64, device=dev) << torch.arange(64, device=dev))  # bin t -> bit t
    out = torch.empty((n, K, sketchsize64 * BBITS), dtype=torch.int64, device=dev)
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        which = torch.randint(0, n_lineages, (m,), device=dev, generator=g)
        sig = derive(lin[which], p_keep(m)).to(torch.int64).view(m, K, sketchsize64, 64)
        dst = out[start:start + m].view(m, K, sketchsize64, BBITS)
        for b in range(BBITS):
            dst[..., b] = (((sig >> b) & 1) * weights).sum(dim=-1)
    return out
