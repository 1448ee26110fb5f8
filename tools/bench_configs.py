"""The other BASELINE.json configs, at FULL size, as `bench.py --config cfg2|cfg4|cfg5` (1 GPU or N ranks under torchrun).
One JSON line per run (rank 0); the lines recorded on the B200 boxes are committed under profiles/.

cfg2  10k genomes, S=1024, k={15,19,23,27,31}, all-vs-all (BASELINE.json configs[1])
cfg4  poppunk_assign shape (PopPUNK/assign.py:502-510, 593-601 + network.py:1180-1184): 1 M queries x 50 k refs.
      Queries are SCATTERED (rank r generates / holds only its 1/N of them; refs replicated), rows are query-major so a
      rank's output is a contiguous row range.  The float2 result would be 400 GB, so the result forms are
        labels   fused assign_threshold -> int8 label per pair, copied to the host (50 GB in total)
        edges    fused distance -> boundary -> edge list in ONE kernel pass, nothing of size n_pairs written
      and, from ONE process on all GPUs through the host API (ppb_query_host_multi, labels only), the streamed form.
cfg5  50k genomes, S=16384 (sketchsize64=256), K=5, all-vs-all: 16 slices per k accumulate in shared memory; the
      packed sketches (7.2 GB) are 57x the L2.
Every line carries a sampled-row parity check against the CPU oracle.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TOL = 1e-6
LOP3_PEAK = 18.5e12   # LOP3 lane-ops/s of one B200 (bench.py measures it in-run: 148 SM x 64 lanes x 1.965 GHz)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def run(args):
    import torch
    import torch.distributed as dist
    from poppunk_b200 import engine, synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    try:
        oracle.build(native=True)
        native = True
    except Exception:
        native = False
    threads = max(1, len(os.sched_getaffinity(0)))

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # CPU-side waits while rank 0 drives all GPUs from one process

    def host_barrier():
        if host_group is not None:
            dist.barrier(group=host_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup=1):
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    steps = max(1, min(args.steps, 3))
    pop = dict(n_roots=2, n_lineages=8)
    line = None

    if args.config in ("cfg2", "cfg5"):
        if args.config == "cfg2":
            n, ss64, kmers, chunk = 10_000, 16, np.array([15, 19, 23, 27, 31], dtype=np.int32), 4096
        else:
            n, ss64, kmers, chunk = 50_000, 256, np.array([13, 17, 21, 25, 29], dtype=np.int32), 512
        if args.n != 100_000:
            n = args.n
        total = n * (n - 1) // 2
        sk = synth.synth_sketches_torch(n, kmers, ss64, seed=5, device=dev, chunk=chunk, **pop)
        table, cl = synth.random_match_table(kmers, 3), synth.synth_clusters(n, 3)
        table_dev, cl_dev = torch.as_tensor(table).to(dev), engine.DeviceClusters.upload(cl, dev)   # resident inputs
        b, e, _ = engine.shard_rows(total, world, rank)
        packed = engine.pack(sk, clusters=cl_dev)
        out = torch.empty((e - b, 2), dtype=torch.float32, device=dev)
        ndeg = torch.zeros(1, dtype=torch.int64, device=dev)

        def step():
            p = engine.pack(sk, clusters=cl_dev)
            engine.query(p, None, kmers, rand_table=table_dev, row_begin=b, row_end=e, out=out, n_degenerate=ndeg)

        ms = timed(step, steps)
        ndeg.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        engine.query(packed, None, kmers, rand_table=table_dev, row_begin=b, row_end=e, out=out, n_degenerate=ndeg)
        ev1.record()
        torch.cuda.synchronize()
        k_ms = ev0.elapsed_time(ev1)
        if world > 1:
            dist.all_reduce(ndeg)
        # parity: first / last rows of this rank's shard against the oracle
        host = sk.cpu().numpy().view(np.uint64)
        m = min(e - b, 20_000 if ss64 > 16 else 100_000) // 2
        worst = 0.0
        for (r0, r1) in ((b, b + m), (e - m, e)):
            exp, _ = oracle.query(host, None, kmers, table, cl, row_begin=r0, row_end=r1, threads=threads, native=native)
            worst = max(worst, float(np.abs(out[r0 - b:r1 - b].cpu().numpy() - exp).max()))
        worst = max_over_ranks(worst)
        # e2e from ONE process (rank 0) on all GPUs through the drop-in's call
        e2e = None
        del packed, out
        torch.cuda.empty_cache()
        barrier()
        host_barrier()
        if rank == 0:
            from poppunk_b200 import sketchlib
            os.environ["PPB_DEVICES"] = str(world)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                res, nd = sketchlib.query_arrays(host, None, kmers, table, cl, None, device_id=local)
                ts.append(time.perf_counter() - t0)
                del res
            e2e = {"first_call_ms": ts[0] * 1e3, "second_call_ms": ts[1] * 1e3, "pinning_call_ms": ts[2] * 1e3,
                   "steady_ms": min(ts[3:]) * 1e3,
                   "value": total / min(ts[3:]), "unit": "pairs/s", "h2d_bytes_per_step": int(host.nbytes),
                   "d2h_bytes_per_step": total * 8, "api": f"sketchlib.query_arrays, one process, {world} GPU(s)"}
        host_barrier()
        barrier()
        lop3_pair = len(kmers) * ss64 * 2 * 14
        rows_rank = e - b
        line = {"config": args.config, "metric": f"genome-pairs/sec (core+acc dist) at N={n} S={64 * ss64} K={len(kmers)}",
                "workload": f"self all-vs-all, N={n}, S={64 * ss64}, k={kmers.tolist()}, random_correct on, {total} pairs",
                "value": total / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "ms_per_step": ms, "steps": steps,
                "kernel_ms_rank0_shard": k_ms,
                "int_pipe": {"lop3_per_pair": lop3_pair, "achieved_lop3_per_s_per_gpu": rows_rank * lop3_pair / (k_ms * 1e-3),
                             "frac_of_18.5T": rows_rank * lop3_pair / (k_ms * 1e-3) / LOP3_PEAK},
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_pair": 8 + n * len(kmers) * ss64 * 14 * 8 / total,
                             "achieved_GBps_per_gpu": (rows_rank * 8 + n * len(kmers) * ss64 * 14 * 8) / (k_ms * 1e-3) / 1e9,
                             "peak_GBps": 6553.0},
                "parity_max_abs_err_vs_oracle": worst, "parity_ok": bool(worst <= TOL), "n_degenerate": int(ndeg.item()),
                "e2e": e2e}

    elif args.config == "cfg4":
        kmers = np.array([13, 17, 21, 25, 29], dtype=np.int32)
        R, Q = 50_000, 1_000_000
        if args.n != 100_000:          # scaled-down runs for testing: --genomes sets the query count
            Q = args.n
        q_lo, q_hi = Q * rank // world, Q * (rank + 1) // world
        nq = q_hi - q_lo
        # one ancestor: every query is related to the references (a poppunk_assign run against its own species), and the
        # boundary puts ~2 % of the pairs inside (the two-ancestor population of the north-star bench would make half of
        # the pairs degenerate (0, 0), i.e. "within" for any boundary — not what an assign job looks like)
        pop = dict(n_roots=1, n_lineages=8)
        bnd = (2, 0.012, 0.15, 1.0, 1.0)
        table = synth.random_match_table(kmers, 3)
        rsk = synth.synth_sketches_torch(R, kmers, 16, seed=9, device=dev, **pop)           # replicated
        rcl = synth.synth_clusters(R, 3, seed=9)
        qsk = torch.empty((nq, 5, 224), dtype=torch.int64, device=dev)                      # scattered: this rank's only
        blk = 125_000
        for s0 in range(0, nq, blk):   # the same population as the refs (seed), different genomes (block id in the seed)
            m = min(blk, nq - s0)
            qsk[s0:s0 + m] = synth.synth_sketches_torch(m, kmers, 16, seed=9, device=dev, sample_seed=1 + (q_lo + s0) // blk, **pop)
        qcl = synth.synth_clusters(Q, 3, seed=11)[q_lo:q_hi]
        table_dev, qcl_dev = torch.as_tensor(table).to(dev), engine.DeviceClusters.upload(qcl, dev)   # resident inputs
        rows = R * nq
        ref_p = engine.pack(rsk, clusters=rcl)
        labels = torch.empty(rows, dtype=torch.int8, device=dev)
        labels_host = torch.empty(rows, dtype=torch.int8, pin_memory=True)
        ndeg = torch.zeros(1, dtype=torch.int64, device=dev)

        def step_labels():
            qp = engine.pack(qsk, clusters=qcl_dev)
            engine.query(ref_p, qp, kmers, rand_table=table_dev, boundary=bnd, want_out=False, labels=labels, n_degenerate=ndeg)
            labels_host.copy_(labels, non_blocking=True)

        ms_labels = timed(step_labels, steps)
        edges = {}

        def step_edges():
            qp = engine.pack(qsk, clusters=qcl_dev)
            oi, oj, n_e, nd = engine.query_edges(ref_p, qp, kmers, bnd, rand_table=table_dev, capacity=max(1 << 22, rows // 16))
            edges["i"], edges["j"], edges["n"] = oi.cpu(), oj.cpu(), n_e

        ms_edges = timed(step_edges, steps)
        # parity: labels of the first / last rows of this rank's shard; edges == positions of label -1
        rh, qh = rsk.cpu().numpy().view(np.uint64), qsk.cpu().numpy().view(np.uint64)
        torch.cuda.synchronize()
        lab = labels_host.numpy()
        bad = 0
        for (r0, r1) in ((0, 20_000), (rows - 20_000, rows)):
            qa, qb = r0 // R, (r1 - 1) // R + 1
            _, lab_o, _ = oracle.query(rh, qh[qa:qb], kmers, table, rcl, qcl[qa:qb], row_begin=r0 - qa * R,
                                       row_end=r1 - qa * R, boundary=bnd, threads=threads, native=native)
            bad += int((lab[r0:r1] != lab_o).sum())
        n_within = int((labels == -1).sum().item())
        edges_ok = (edges["n"] == n_within)
        tot = torch.tensor([bad, n_within, int(edges_ok), edges["n"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        total = R * Q
        # the streamed host form from ONE process on all GPUs (labels only), if the host can hold 50 GB of labels
        e2e = None
        del labels, labels_host, ref_p
        torch.cuda.empty_cache()
        barrier()
        if world > 1:   # rank 0 needs every rank's queries for the one-process call
            gathered = [torch.empty((Q * (r + 1) // world - Q * r // world, 5, 224), dtype=torch.int64, device=dev) for r in range(world)] if rank == 0 else None
            dist.gather(qsk, gathered, dst=0)
            q_all = torch.cat(gathered).cpu().numpy().view(np.uint64) if rank == 0 else None
            del gathered
        else:
            q_all = qh
        torch.cuda.synchronize()
        host_barrier()
        if rank == 0:
            avail = _mem_available_gb()
            if avail > total / 1e9 * 1.3 + 40:
                os.environ["PPB_DEVICES"] = str(world)
                qcl_all = synth.synth_clusters(Q, 3, seed=11)
                ts, same = [], True
                for _ in range(5):      # first (fresh block, staged) / touched block / page-locking call / steady / steady
                    t0 = time.perf_counter()
                    _, lab_h, nd = engine.query_host(rh, q_all, kmers, table, rcl, qcl_all, boundary=bnd, want_out=False,
                                                     devices=engine.visible_devices(local))
                    ts.append(time.perf_counter() - t0)
                    same = same and bool((lab_h[:20_000] == lab[:20_000]).all())
                    del lab_h           # the block goes back to the pool before the next call asks for one
                e2e = {"calls_ms": [round(t * 1e3, 1) for t in ts], "value": total / min(ts), "unit": "pairs/s",
                       "calls": "first (fresh pool block, staged) / touched block, staged / page-locking call / steady / steady",
                       "h2d_bytes_per_step": int(rh.nbytes + q_all.nbytes), "d2h_bytes_per_step": total,
                       "labels_identical_to_device_run_first_20k": same,
                       "api": f"engine.query_host(..., boundary, want_out=False) = ppb_query_host_multi, one process, {world} GPU(s): "
                              "queries scattered by the library (each device uploads only its own), labels streamed per row chunk"}
            else:
                e2e = {"skipped": f"host has {avail:.0f} GB available; {total / 1e9:.0f} GB of labels + sketches do not fit comfortably"}
        host_barrier()
        barrier()
        line = {"config": "cfg4", "metric": "genome-pairs/sec, query-vs-ref with fused assign_threshold",
                "workload": f"{Q} queries x {R} refs, S=1024, K=5, random_correct on, boundary slope 2 (0.012, 0.15); "
                            f"queries scattered over {world} rank(s) ({nq} on rank 0), refs replicated; {total} pairs",
                "n_gpus": world, "steps": steps,
                "labels": {"ms_per_step": ms_labels, "value": total / (ms_labels * 1e-3), "unit": "pairs/s",
                           "out_bytes_per_pair": 1, "d2h_bytes_per_step_per_rank": rows,
                           "what": "pack queries + fused kernel -> int8 labels + D2H of the labels to pinned host memory, per rank"},
                "edges": {"ms_per_step": ms_edges, "value": total / (ms_edges * 1e-3), "unit": "pairs/s",
                          "n_edges": int(tot[3].item()), "what": "pack queries + fused kernel -> edge list (row order) + D2H of the edges"},
                "int_pipe_frac_of_18.5T_labels": (total / world) * 2240 / (ms_labels * 1e-3) / LOP3_PEAK,
                "parity": {"label_mismatches_vs_oracle_in_sampled_rows": int(tot[0].item()), "rows_sampled_per_rank": 40_000,
                           "edges_equal_within_labels": bool(tot[2].item() == world), "within_boundary_pairs": int(tot[1].item()),
                           "within_boundary_fraction": float(tot[1].item()) / total},
                "parity_ok": bool(tot[0].item() == 0 and tot[2].item() == world), "e2e_streamed_labels": e2e}

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and line and not line.get("parity_ok", True):
        sys.exit(3)


def _mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0
