"""Multi-GPU parity check, run under torchrun on a multi-GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/multigpu_check.py

G ranks: (1) static row shards + one NCCL all_gather_into_tensor, (2) the fused exchange (peer stores, then
multicast if available) must both be byte-identical to the single-GPU result and within 1e-6 of the CPU oracle.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from poppunk_b200 import engine, synth  # noqa: E402

KMERS = np.array([13, 17, 21, 25, 29], dtype=np.int32)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 3001
    ref = synth.synth_sketches(n, KMERS, 16, seed=5)
    tab, cl = synth.random_match_table(KMERS, 3), synth.synth_clusters(n, 3)
    packed = engine.pack(ref, clusters=cl, device=dev)
    single, _, nd1 = engine.query(packed, None, KMERS, rand_table=tab)
    gathered, ndg = engine.query_sharded(packed, None, KMERS, rand_table=tab)
    torch.cuda.synchronize()
    assert gathered.shape == single.shape and bool((gathered == single).all()), "all-gather path differs"
    assert int(ndg.item()) == int(nd1.item())
    results = {"allgather": "ok"}
    for use_mc in (False, True):
        ex = engine.FusedExchange(single.shape[0], dev, use_multicast=use_mc)
        if use_mc and not ex.mc_ptr:
            results["multicast"] = "unsupported on this fabric"
            continue
        for rep in range(2):
            full, nd = ex.run(packed, None, KMERS, rand_table=tab)
            torch.cuda.synchronize()
            dist.barrier()
            assert bool((full == single).all()), f"fused exchange differs (multicast={use_mc}, rep={rep})"
        tot = nd.clone()
        dist.all_reduce(tot)
        assert int(tot.item()) == 2 * int(nd1.item())
        results["multicast" if use_mc else "peer_stores"] = "ok"
    # rectangular mode through the sharded path
    qry = engine.pack(synth.synth_sketches(257, KMERS, 16, seed=5, sample_seed=1), device=dev)
    r1, _, _ = engine.query(packed, qry, KMERS)
    rg, _ = engine.query_sharded(packed, qry, KMERS)
    assert bool((r1 == rg).all())
    if rank == 0:
        import oracle
        exp, _ = oracle.query(ref, None, KMERS, tab, cl, row_end=200_000)
        err = float(np.abs(single[:200_000].cpu().numpy() - exp).max())
        assert err <= 1e-6
        print(f"multigpu_check ok: world={world} {results} max|d-oracle|={err:.1e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
