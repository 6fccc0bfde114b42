#!/usr/bin/env python
"""Benchmark of the Bayesian-BM25 query-time hot path on B200.

Workload (BASELINE.json configs[1]/[2]): MS-MARCO-passage-shaped synthetic corpus,
8.8 M docs, 30 k-term Zipf vocabulary, a 10 k-query batch, top-1000 calibrated
probabilities.  One "step" = one pass of the whole batch through
BayesianBM25Scorer.retrieve (traversal + posterior + top-k [+ all-gather merge]).

    python bench.py --gpus 1 --steps 5 --warmup 3
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU arm: oracle port on all host cores

Prints ONE JSON line (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (top-1000 calibrated probs, 8.8M docs)"
N_DOCS, VOCAB, AVG_LEN, N_QUERIES, TOP_K = 8_800_000, 30_000, 56.0, 10_000, 1000
CORPUS_SEED, QUERY_SEED = 42, 43
ALPHA, BETA, BASE_RATE = 2.0, 0.2, 0.045  # fixed transform constants (index-time estimation is not the timed path)


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # let nvidia-smi finish initialising
                time.sleep(0.05)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per traversal launch from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("tile_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def build_corpus(device, n_docs):
    from bayesian_bm25_b200 import synthetic
    return synthetic.zipf_csc(n_docs, VOCAB, AVG_LEN, CORPUS_SEED, device, k1=1.2, b=0.75, method="lucene")


def algorithmic_bytes(df: np.ndarray, q_terms: np.ndarray, n_queries: int, k: int) -> int:
    """sum over query terms (with multiplicity) of df*8 B + k*20 B per query."""
    return int(df[q_terms].sum()) * 8 + n_queries * k * 20


def cpu_baseline(host_csc, q_terms, q_off, k, sample_q, threads=0):
    from oracle import coracle
    params = coracle.make_params(ALPHA, BETA, BASE_RATE)
    qt = q_terms[: q_off[sample_q]]
    qo = q_off[: sample_q + 1]
    t0 = time.perf_counter()
    _, _, _, used = coracle.retrieve_batch(host_csc, params, qt, qo, k, n_threads=threads)
    dt = time.perf_counter() - t0
    return sample_q / dt, used, dt


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port: bm25s-equivalent scatter-add +
    top-k + the reference's posterior) on all host cores, bounded query sample."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    import torch
    from bayesian_bm25_b200 import synthetic
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    n_docs = args.docs
    csc = build_corpus(dev, n_docs)
    host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}
    del csc
    q_terms, q_off = synthetic.zipf_queries(args.queries, VOCAB, QUERY_SEED)
    from oracle import coracle
    cores = coracle.max_threads()
    sample = min(args.queries, max(256, 24 * cores))
    for _ in range(args.warmup):
        cpu_baseline(host, q_terms, q_off, args.k, min(sample, cores))
    times = []
    for _ in range(args.steps):
        qps, used, dt = cpu_baseline(host, q_terms, q_off, args.k, sample)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n_docs} docs, {VOCAB}-term Zipf vocab, top-{args.k}; each step = {sample} queries of the {args.queries}-query batch"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": used, "kind": "port",
                         "sample": f"{sample} queries/step x {args.steps} steps, oracle/bb25_oracle.c orc_retrieve_batch, {used} threads"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=N_DOCS)
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--k", type=int, default=TOP_K)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--prune-level", type=int, default=3, help="pruning level of the extra 'pruned' pass")
    ap.add_argument("--shard-chunks", type=int, default=1,
                    help="query sub-batches pipelined against the all-gather (N>1); 1 measured best at Q=10k (profiles/r01)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    import __graft_entry__ as entry
    if not os.path.exists(entry.SO):
        entry.build()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from bayesian_bm25_b200 import BayesianBM25Scorer, _lib, sharded, synthetic

    world = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local = _env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- corpus + index (untimed) ---------------------------------------------------
    t_build = time.perf_counter()
    csc = build_corpus(dev, args.docs)
    df_full = (csc["indptr"][1:] - csc["indptr"][:-1]).cpu().numpy()
    nnz_full = int(csc["data"].numel())
    host_csc = None
    if rank == 0 and not args.no_cpu:
        host_csc = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}
    if world > 1:
        csc = sharded.local_shard(csc, rank, world)
    scorer = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=ALPHA, beta=BETA, base_rate=BASE_RATE)
    scorer.index_from_csc(csc)
    del csc
    torch.cuda.empty_cache()
    retr = sharded.ShardedRetriever(scorer, n_chunks=args.shard_chunks)
    q_terms, q_off = synthetic.zipf_queries(args.queries, VOCAB, QUERY_SEED)
    d_terms = torch.from_numpy(q_terms).to(dev)
    d_off = torch.from_numpy(q_off).to(dev)
    t_build = time.perf_counter() - t_build

    def step():
        return retr.retrieve_ids_device(d_terms, d_off, args.k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_step():
        """The call a user makes: host buffers in, host buffers out (H2D + D2H inside)."""
        if world == 1:
            return scorer.retrieve_ids(q_terms, q_off, args.k)
        hp = torch.empty(q_terms.size, dtype=torch.int32, pin_memory=True)
        hp.numpy()[:] = q_terms
        ho = torch.empty(q_off.size, dtype=torch.int64, pin_memory=True)
        ho.numpy()[:] = q_off
        ids, sc, pr = retr.retrieve_ids_device(hp.to(dev, non_blocking=True), ho.to(dev, non_blocking=True), args.k)
        if rank != 0:  # the merged result is identical on every rank; the caller lives on rank 0
            torch.cuda.current_stream().synchronize()
            return None
        h_ids = torch.empty(ids.shape, dtype=ids.dtype, pin_memory=True)
        h_pr = torch.empty(pr.shape, dtype=pr.dtype, pin_memory=True)
        h_ids.copy_(ids, non_blocking=True)
        h_pr.copy_(pr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h_ids.numpy(), h_pr.numpy()

    def measure(level: int, sample_clocks: bool):
        """W warm-up + K timed steps at one pruning level: device-resident timing (CUDA
        events, max over ranks) and the end-to-end host-API timing."""
        scorer.set_pruning(level)
        for _ in range(max(args.warmup, 1)):
            out = step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
        launches0 = _lib.lib().bb25_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {"traverse_ms": 0.0, "traverse_launches": 0, "rerun_queries": 0, "units": 0, "units_skipped": 0,
               "units_maxscore": 0, "routed_queries": 0, "candidate_items": 0}
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step()
            st = retr.stats()
            for key in acc:
                acc[key] += st[key]
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = int(_lib.lib().bb25_launch_count() - launches0)
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        t = torch.tensor([ms, acc["traverse_ms"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for _ in range(3):  # lets torch's pinned-host allocator settle on reusable blocks
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return {"ms": float(t[0]), "trav_ms": float(t[1]), "launches": launches, "clocks": clocks,
                "e2e_s": float(te[0]), "out": out, **{k_: v / args.steps for k_, v in acc.items() if k_ != "traverse_ms"}}

    # headline: exhaustive traversal (every posting of every query term is visited, like the
    # reference); then the same batch with the library's default dynamic pruning (exact)
    ex = measure(0, sample_clocks=True)
    shard_timing = None
    if world > 1:  # one extra, untimed, unpipelined pass that times the phases separately
        retr.profile = True
        step()
        barrier()
        step()
        shard_timing = {k_: v / max(1, retr.timing["calls"]) for k_, v in retr.timing.items() if k_ != "calls"}
        retr.profile = False
    pr = measure(args.prune_level, sample_clocks=False)
    same = all(bool(torch.equal(x, y)) for x, y in zip(ex["out"], pr["out"]))
    out = ex["out"]
    ms, trav_ms_max, launches, clocks, e2e_s = ex["ms"], ex["trav_ms"], ex["launches"], ex["clocks"], ex["e2e_s"]
    trav_launches, reruns = ex["traverse_launches"] * args.steps, ex["rerun_queries"] * args.steps
    h2d = int(q_terms.nbytes + q_off.nbytes)
    d2h = int(args.queries * args.k * (8 + 8))

    if rank == 0:
        qps = args.queries * args.steps / (ms / 1000.0)
        alg_bytes_step = algorithmic_bytes(df_full, q_terms, args.queries, args.k)
        # every rank scans 1/world of each posting list; the roofline line is per GPU
        alg_bytes_gpu = alg_bytes_step / world
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes_gpu * args.steps / (trav_ms_max / 1000.0) / 1e9 if trav_ms_max > 0 else 0.0
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": (f"{args.docs} docs, {VOCAB}-term Zipf vocab (avg len {AVG_LEN:g}), nnz {nnz_full}, "
                             f"{args.queries}-query batch (3-5 terms), top-{args.k}, exhaustive traversal"),
                "parallelism": f"doc-range shards x{world}" + (" + NCCL all-gather merge" if world > 1 else ""),
                "cache": "inputs larger than L2 (CSC index %.2f GB per GPU, 126 MB L2)" % (nnz_full * 8 / world / 1e9),
                "probabilities": "fp64 posterior fused on device", "index_build_s": round(t_build, 1),
                "threshold_reruns_per_step": reruns / args.steps,
                "kernel": os.environ.get("BB25_KERNEL", "block"), "pruning_level": 0,
            },
            "clocks": clocks,
            "sharded_breakdown_ms_per_call": shard_timing,
            "e2e": {"value": args.queries * args.steps / e2e_s, "unit": "queries/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "pruned": {
                "what": "same batch with the library's default dynamic pruning (bb25_index_set_pruning level %d: "
                        "block-max skip + per-block non-essential frequent terms + candidate-driven rare-term queries); results bit-identical "
                        "to the exhaustive pass" % args.prune_level,
                "value": args.queries * args.steps / (pr["ms"] / 1000.0), "unit": "queries/s",
                "ms_per_step": pr["ms"] / args.steps, "kernel_ms_per_step": pr["trav_ms"] / args.steps,
                "e2e_value": args.queries * args.steps / pr["e2e_s"], "results_identical": same,
                "block_docs": 1024, "units_per_step": pr["units"], "units_skipped_per_step": pr["units_skipped"],
                "units_maxscore_per_step": pr["units_maxscore"],
                "queries_routed_to_candidate_path_per_step": pr["routed_queries"],
                "candidate_items_per_step": pr["candidate_items"],
            },
            "roofline": {
                "bound": "hbm", "kernel": "bb25::block_kernel (order-free posting traversal, exhaustive; candidates re-scored in query order by select_kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": ncu_traffic_per_launch(),
                "algorithmic_bytes_per_step_per_gpu": alg_bytes_gpu,
                "kernel_ms_per_step": trav_ms_max / args.steps, "kernel_launches_per_step": trav_launches / args.steps,
                "note": "achieved = sum_q sum_t df(t)*8B (+k*20B) / summed traversal-kernel time (CUDA events in libbb25); "
                        "all warps share a block's index slice through L2, so DRAM traffic is far below the algorithmic "
                        "bytes and the figure can exceed the HBM peak (DESIGN.md 4.1); the exact re-scoring of the "
                        "emitted candidates runs in select_kernel and is part of ms_per_step, not of this kernel time",
            },
        }
        if host_csc is not None:
            from oracle import coracle
            cores = coracle.max_threads()
            sample = args.cpu_sample or min(args.queries, max(256, 24 * cores))
            v, used, dt = cpu_baseline(host_csc, q_terms, q_off, args.k, sample)
            line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": used, "kind": "port",
                                    "sample": f"first {sample} queries of the batch, oracle/bb25_oracle.c, {dt:.1f} s wall"}
            # spot-check the timed output against the oracle on the sampled queries
            o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host_csc, coracle.make_params(ALPHA, BETA, BASE_RATE),
                                                          q_terms[: q_off[8]], q_off[:9], args.k)
            ids8 = out[0][:8].cpu().numpy()
            line["parity_spot_check"] = bool(np.array_equal(ids8, o_ids) and
                                             np.max(np.abs(out[2][:8].cpu().numpy() - o_pr)) < 1e-9)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
