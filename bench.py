#!/usr/bin/env python
"""bench.py — genome-pairs/sec of the core+accessory sketch-distance path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n GENOMES]

One "step" = one full pass of the hot path over the synthetic workload: pack + distance kernel over all
N(N-1)/2 pairs (self mode, S=1024 bins, K=5 k-mers), result left in HBM in PopPUNK's condensed row order
(N>1 ranks: static row shards + one NCCL all-gather).  Prints ONE JSON line (rank 0).

  value     whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the host-buffer C-ABI call (ppb_query_host): H2D of the sketches from
            pinned host memory, pack, kernels, D2H of the (pairs x 2) float32 result inside the timed region
  roofline  the dominant kernel (query_kernel) against the measured HBM peak (MEASURED_PEAKS.json), using the
            ALGORITHMIC bytes (8 B/pair out + every sketch word read once); plus an "int_pipe" block — the
            kernel is bound by the INT32 logic pipe (14 LOP3 per 32 bins), measured here with a LOP3-only
            micro-kernel, and that fraction is the one that says how good the kernel is
  cpu_baseline  the CPU oracle (a restatement of the pp-sketchlib CPU path; the library itself is absent)
            on all host cores, on a bounded row range of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner out of it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

KMERS = np.array([13, 17, 21, 25, 29], dtype=np.int32)   # PopPUNK defaults: k = 13..29 step 4 (__main__.py:77-79)
SS64 = 16                                                # S = 1024 bins
SEED = 42
METRIC = "genome-pairs/sec (core+acc dist) at N=100k S=1024 K=5"
UNIT = "pairs/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def algorithmic_bytes(n, rows):
    """8 B out per pair + every sketch word read once."""
    return rows * 8 + n * len(KMERS) * SS64 * 14 * 8


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = f"/tmp/ppb_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host cores, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    native = False
    try:  # same image on the GPU box: rebuild for the host CPU when gcc is there, else use the shipped .so
        oracle.build(native=True)
        native = True
    except Exception:
        pass
    return oracle, native


def host_threads():
    """Every core this process may run on.  (torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently
    turn the reference arm of an N>1 launch into a single-thread run: the thread count is passed explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(oracle, native, ref_host, target_s=12.0):
    """Time rows [0, R) of the self job (R sized for ~target_s of CPU work).  Returns (pairs/s, cores, R)."""
    n = ref_host.shape[0]
    total = n * (n - 1) // 2
    threads = host_threads()
    probe = min(total, 200_000 * threads)
    t0 = time.perf_counter()
    oracle.query(ref_host, None, KMERS, row_begin=0, row_end=probe, threads=threads, native=native)
    rate = probe / (time.perf_counter() - t0)
    rows = int(min(total, max(probe, rate * target_s)))
    t0 = time.perf_counter()
    oracle.query(ref_host, None, KMERS, row_begin=0, row_end=rows, threads=threads, native=native)
    dt = time.perf_counter() - t0
    return rows / dt, threads, rows


def host_sketches(n):
    """The workload on the host without a GPU (reference arm on a CPU-only box): NumPy generator."""
    from poppunk_b200 import synth
    return synth.synth_sketches(n, KMERS, SS64, seed=SEED)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (pp-sketchlib itself is
    not in /root/reference and not installable here), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle, native = load_oracle()
    n = args.n
    try:
        import torch
        if torch.cuda.is_available():
            from poppunk_b200 import synth
            ref_host = synth.synth_sketches_torch(n, KMERS, SS64, seed=SEED, device="cuda:0").cpu().numpy().view(np.uint64)
        else:
            ref_host = host_sketches(min(n, 20_000))
    except Exception:
        ref_host = host_sketches(min(n, 20_000))
    n_eff = ref_host.shape[0]
    total = n_eff * (n_eff - 1) // 2
    threads = host_threads()
    # size one step at ~8 s of CPU work
    probe = min(total, 200_000 * threads)
    t0 = time.perf_counter()
    oracle.query(ref_host, None, KMERS, row_begin=0, row_end=probe, threads=threads, native=native)
    rate = probe / (time.perf_counter() - t0)
    rows = int(min(total, max(probe, rate * 8.0)))
    for _ in range(args.warmup):
        oracle.query(ref_host, None, KMERS, row_begin=0, row_end=rows, threads=threads, native=native)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.query(ref_host, None, KMERS, row_begin=0, row_end=rows, threads=threads, native=native)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows / dt
    sample = f"rows [0,{rows}) of the N={n_eff} self job per step ({rows} pairs)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"self all-vs-all, N={n_eff}, S=1024, K=5 (k=13..29 step 4), random_correct off",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the pp-sketchlib path (oracle/ppb_oracle.c, OpenMP, "
                                 + ("-march=native" if native else "-march=x86-64-v3") + "); pp-sketchlib absent"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from poppunk_b200 import _lib, engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    n = args.n
    total = n * (n - 1) // 2
    free_b, _ = torch.cuda.mem_get_info(dev)
    need = total * 8 * (2 if world > 1 else 1) + 3 * n * 5 * SS64 * 14 * 8 + (4 << 30)
    note = ""
    while need > free_b * 0.9 and n > 10_000:   # never expected on a 180 GB B200; keeps the bench alive elsewhere
        n //= 2
        total = n * (n - 1) // 2
        need = total * 8 * (2 if world > 1 else 1) + 3 * n * 5 * SS64 * 14 * 8 + (4 << 30)
        note = f" (reduced from N={args.n}: device memory)"
    if rank == 0:
        log(f"[bench] N={n} pairs={total} world={world}{note}")

    sk = synth.synth_sketches_torch(n, KMERS, SS64, seed=SEED, device=dev)     # identical on every rank
    torch.cuda.synchronize()
    b, e, slice_len = engine.shard_rows(total, world, rank)
    rows_rank = e - b
    ndeg = torch.zeros(1, dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step):
        """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks."""
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.ppb_launch_count()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, L.ppb_launch_count() - n0

    exchange_desc = "none (single GPU)"
    nccl_ms = None
    sampler = None
    if world == 1:
        out = torch.empty((total, 2), dtype=torch.float32, device=dev)

        def step():
            packed = engine.pack(sk)
            engine.query(packed, None, KMERS, out=out, n_degenerate=ndeg)

        sampler = ClockSampler(local)
        ms_step, launches = timed_loop(step)
        clocks = sampler.stop()
        mine = out
    else:
        # (a) NCCL: static row shards + one in-place all_gather_into_tensor per step
        full = torch.empty((world * slice_len, 2), dtype=torch.float32, device=dev)
        mine = full[rank * slice_len:(rank + 1) * slice_len]

        def step_nccl():
            packed = engine.pack(sk)
            engine.query(packed, None, KMERS, row_begin=b, row_end=e, out=mine[:rows_rank], n_degenerate=ndeg)
            dist.all_gather_into_tensor(full, mine)

        nccl_ms, _ = timed_loop(step_nccl)
        del full, mine
        torch.cuda.empty_cache()
        # (b) fused: the kernel's epilogue warps store every row into all ranks' buffers (NVSwitch multicast
        #     when available, else peer stores) — no collective call at all; this is the headline number
        ex = engine.FusedExchange(total, dev)
        exchange_desc = ("fused in-kernel exchange: " + ("multimem.st (NVSwitch multicast)" if ex.mc_ptr else
                         f"{world} coalesced peer stores per row over NVLink") + ", symmetric memory, no all-gather")

        def step():
            packed = engine.pack(sk)
            ex.run(packed, None, KMERS, n_degenerate=ndeg)

        if rank == 0:
            sampler = ClockSampler(local)
        ms_step, launches = timed_loop(step)
        clocks = sampler.stop() if sampler else None
        mine = ex.full[b:e]
    value = total / (ms_step * 1e-3)

    # per-launch duration of the dominant kernel, on the launching stream (same inputs, local shard only)
    packed = engine.pack(sk)
    ev_k0, ev_k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    scratch = mine[:rows_rank]
    for _ in range(args.steps):
        ev_k0.record()
        engine.query(packed, None, KMERS, row_begin=b, row_end=e, out=scratch, n_degenerate=ndeg)
        ev_k1.record()
        torch.cuda.synchronize()
        kernel_ms.append(ev_k0.elapsed_time(ev_k1))
    ndeg.zero_()
    engine.query(packed, None, KMERS, row_begin=b, row_end=e, out=scratch, n_degenerate=ndeg)
    if world > 1:
        dist.all_reduce(ndeg)
    n_degenerate = int(ndeg.item())
    k_ms = float(np.mean(kernel_ms))
    full = None

    # ---- e2e through the host-buffer C-ABI call (pinned host buffers, copies inside the timed region)
    e2e = None
    try:
        sk_host = torch.empty(sk.shape, dtype=torch.int64, pin_memory=True)
        sk_host.copy_(sk)
        out_host = torch.empty((rows_rank, 2), dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()
        del packed, scratch
        if world == 1:
            del out, mine
        torch.cuda.empty_cache()
        ref_np = sk_host.numpy().view(np.uint64)
        out_np = out_host.numpy()

        def e2e_step():
            engine.query_host(ref_np, None, KMERS, row_begin=b, row_end=e, out=out_np, device_id=local)

        e2e_step()  # warm-up (allocations, tile list)
        # Each e2e step is timed on its own (barrier + wall clock on both sides, max over ranks) and the MEDIAN is
        # reported with every run listed: the path runs at the PCIe floor, and on these shared hosts one run in a
        # few picks up a 100-300 ms hiccup on the host side of the link that says nothing about the engine.
        n_e2e = max(3, min(args.steps, 5))
        runs = []
        for _ in range(n_e2e):
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            runs.append(float(dt.item()))
        med = float(np.median(runs))
        e2e = {"value": total / med, "unit": UNIT, "ms_per_step": med * 1e3, "runs_ms": [round(r * 1e3, 1) for r in runs],
               "mean_ms": float(np.mean(runs)) * 1e3,
               "h2d_bytes_per_step": int(ref_np.nbytes), "d2h_bytes_per_step": int(out_np.nbytes),
               "api": "ppb_query_host (poppunk_b200.engine.query_host) on pinned host buffers; median of the listed runs"
                      + ("; each rank copies its own row shard back" if world > 1 else "")}
        checksum = float(out_np[: min(rows_rank, 1 << 20)].sum())
    except Exception as err:  # e.g. the box cannot pin a 40 GB result buffer
        log(f"[bench] e2e leg failed: {err!r}")
        e2e = {"value": None, "unit": UNIT, "error": repr(err)[:200]}
        checksum = None
        ref_np = sk.cpu().numpy().view(np.uint64)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- integer-pipe micro-roofline (the kernel's real bound) measured on this GPU
    int_pipe = None
    try:
        sink = torch.zeros(4, dtype=torch.int32, device=dev)
        import ctypes as C
        rates = {}
        for mode, name in ((0, "lop3"), (2, "lop3_popc_mix"), (1, "popc"), (3, "redux")):
            ops = C.c_int64(0)
            iters = 20000 if mode != 3 else 4000
            for rep in range(2):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                _lib.check(L.ppb_microbench_dev(mode, iters, sink.data_ptr(), C.byref(ops),
                                                torch.cuda.current_stream(dev).cuda_stream))
                s1.record()
                torch.cuda.synchronize()
            rates[name] = ops.value / (s0.elapsed_time(s1) * 1e-3)
        lop3_per_pair = len(KMERS) * SS64 * 2 * 14                      # 2240
        achieved = rows_rank * lop3_per_pair / (k_ms * 1e-3)
        int_pipe = {"bound": "int32 logic pipe (LOP3)", "lop3_per_pair": lop3_per_pair,
                    "achieved_lop3_per_s": achieved, "peak_lop3_per_s": rates["lop3"],
                    "frac": achieved / rates["lop3"], "peak_mix_14lop3_1popc_per_s": rates["lop3_popc_mix"],
                    "frac_of_mix": achieved / rates["lop3_popc_mix"], "popc_per_s": rates["popc"],
                    "redux_lane_ops_per_s": rates["redux"], "how": "LOP3-only micro-kernel on the same GPU, "
                    "same run (ppb_microbench_dev), CUDA events"}
    except Exception as ex:
        log(f"[bench] microbench failed: {ex!r}")

    peaks = measured_peaks()
    peak = peaks["hbm_gbs"] if peaks else 6650.0
    ach = algorithmic_bytes(n, rows_rank) / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("query_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "query_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                "kernel_ms": k_ms, "algorithmic_bytes_per_pair": algorithmic_bytes(n, rows_rank) / rows_rank,
                "note": "integer popcount path: compulsory HBM traffic is ~8 B/pair, so the HBM fraction is low by "
                        "construction; the binding unit is the INT32 logic pipe — see int_pipe"}

    # ---- CPU baseline on the host cores, bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if world == 1:
        try:
            oracle, native = load_oracle()
            v, cores, rows = cpu_sample(oracle, native, ref_np)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"rows [0,{rows}) of the same N={n} self job ({rows} pairs)",
                   "note": "CPU restatement of the pp-sketchlib path (oracle/ppb_oracle.c, OpenMP, "
                           + ("-march=native" if native else "-march=x86-64-v3") + "); pp-sketchlib itself is absent"}
            # and a live parity spot-check of the timed result against the checker
            if checksum is not None:
                exp, _ = oracle.query(ref_np, None, KMERS, row_begin=b, row_end=b + 100_000, native=native)
                err = float(np.abs(out_np[:100_000] - exp).max())
                cpu["parity_max_abs_err_first_100k_rows"] = err
        except Exception as ex:
            log(f"[bench] cpu baseline failed: {ex!r}")

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"self all-vs-all, N={n}, S=1024 (sketchsize64=16, bbits=14), K=5 (k=13..29 step 4), "
                               f"{total} pairs -> condensed (pairs x 2) float32{note}",
                   "parallelism": f"{world} rank(s): replicated sketches, static condensed-row shards",
                   "exchange": exchange_desc,
                   "cache": "inputs (0.9 GB) and output (40 GB) are larger than the 126 MB L2; no flush needed",
                   "step": "pack_kernel + ytab_kernel + query_kernel"},
        "roofline": roofline, "int_pipe": int_pipe, "cpu_baseline": cpu, "e2e": e2e,
        "nccl_allgather": (None if nccl_ms is None else {
            "ms_per_step": nccl_ms, "value": total / (nccl_ms * 1e-3), "unit": UNIT,
            "note": "same step with the exchange done by one NCCL all_gather_into_tensor instead of in-kernel stores"}),
        "gpu_launches": int(launches), "clocks": clocks,
        "n_degenerate": n_degenerate,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--genomes", dest="n", type=int, default=100_000,
                    help="genomes (default: the north-star N=100k); use --genomes under torchrun, whose own parser claims --n*")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
