"""The tuned CPU arm of the oracle (oracle/ppb_oracle_tuned.inc: AVX-512, cache-blocked, ln J table) — the second CPU
number bench.py reports — is held to the upstream-shaped restatement bit for bit.  Skipped on hosts without AVX-512
VPOPCNTDQ (the arm then reports itself unavailable and bench.py leaves the key null)."""
import numpy as np
import pytest

from poppunk_b200 import synth

KMERS = np.array([13, 17, 21, 25, 29], dtype=np.int32)


@pytest.fixture(scope="module")
def tuned(oracle):
    if not oracle.tuned_available():
        pytest.skip("no AVX-512 VPOPCNTDQ on this host")
    return oracle


@pytest.mark.parametrize("n,ss64,kmers", [(400, 16, KMERS), (300, 1, KMERS), (130, 17, KMERS), (150, 3, KMERS[:3]),
                                          (65, 8, KMERS), (2, 16, KMERS)])
@pytest.mark.parametrize("with_table", [False, True])
def test_tuned_arm_is_bit_identical_to_the_restatement(tuned, n, ss64, kmers, with_table):
    """Whole jobs and row ranges that start and end inside rows / inside 64 x 8 tiles; sketch sizes that fill one vector
    partly (1, 3), exactly (8, 16) and with a tail (17); related and unrelated pairs (truncated and degenerate fits)."""
    sk = synth.synth_sketches(n, kmers, ss64, seed=5, n_roots=2)
    tab = synth.random_match_table(kmers, 3) if with_table else None
    cl = synth.synth_clusters(n, 3) if with_table else None
    exp, nd = tuned.query(sk, None, kmers, tab, cl, threads=3)
    got, nd_t = tuned.query_tuned(sk, kmers, tab, cl, threads=3)
    assert nd_t == nd and (got.view(np.uint32) == exp.view(np.uint32)).all()
    total = exp.shape[0]
    for b, e in ((total // 3 + 5, total - total // 7), (total // 2, total // 2 + 1), (0, min(total, 77))):
        if b >= e:
            continue
        part, nd_p = tuned.query_tuned(sk, kmers, tab, cl, row_begin=b, row_end=e, threads=2)
        ref, nd_r = tuned.query(sk, None, kmers, tab, cl, row_begin=b, row_end=e, threads=2)
        assert nd_p == nd_r and (part.view(np.uint32) == ref.view(np.uint32)).all()


def test_tuned_arm_rejects_what_it_does_not_do(tuned):
    sk = synth.synth_sketches(10, KMERS, 2, seed=1)
    with pytest.raises(RuntimeError):
        tuned.query_tuned(sk, KMERS, row_begin=0, row_end=46)          # beyond the triangle
    out, nd = tuned.query_tuned(sk, KMERS, row_begin=7, row_end=7)      # empty range
    assert out.shape == (0, 2) and nd == 0
