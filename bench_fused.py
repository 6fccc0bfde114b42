"""bench.py --config 4 | 5: the fused-rank configurations of BASELINE.json.

  config 4  hybrid: BM25 posterior + cosine_to_probability through the weighted log_odds_conjunction
            (weights 0.6 / 0.4, alpha 0.5), top-100, 8.8 M documents; a step = one batch of 512 queries,
            each with its own dense cosine row (device-resident: it is the output of the embedding GEMM).
  config 5  MultiFieldScorer(title + body), alpha "auto", equal weights, base_rate "auto", top-10 by fused
            probability with block-max pruning, 50 M documents; a step = one batch of 10 k queries.

One JSON line, same keys as the headline bench (roofline, cpu_baseline, e2e, clocks, gpu_launches).
"""
from __future__ import annotations

import json
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import bench as B

VOCAB = 30_000


def _defaults(args):
    if args.config == 5:
        return args.docs or 50_000_000, args.queries or 10_000, args.k or 10
    return args.docs or 8_800_000, args.queries or 512, args.k or 100


def _host(csc):
    import torch
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}


def _oracle_query(hosts, params, fields, weights, alpha, terms, k, cos_row=None, cos_w=None):
    """The reference's per-query evaluation (CPU oracle): per-field get_probabilities, conjunction, top-k."""
    from oracle import coracle
    cols = [coracle.get_probabilities(hosts[f], params[f], terms) for f in fields]
    w = list(weights)
    if cos_row is not None:
        cols.append(coracle.cosine_to_probability(cos_row.astype(np.float64)))
        w.append(cos_w)
    fused = coracle.log_odds_conjunction(np.stack(cols, axis=-1), alpha=alpha, weights=np.asarray(w))
    return coracle.topk_f64(fused, k)


def _cpu_leg(hosts, params, fields, weights, alpha, qs, k, n_sample, cos=None, cos_w=None):
    from oracle import coracle
    cores = coracle.max_threads()
    n_sample = min(n_sample, len(qs))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        res = list(pool.map(lambda i: _oracle_query(hosts, params, fields, weights, alpha, qs[i], k,
                                                    None if cos is None else cos[i], cos_w), range(n_sample)))
    dt = time.perf_counter() - t0
    return n_sample / dt, cores, dt, res


def run_reference(args):
    """CPU arm of configs 4/5: the oracle's per-query evaluation on all host cores, bounded sample."""
    if B._env_int("RANK", 0) != 0:
        return
    line = run(args, cpu_only=True)
    print(json.dumps(line), flush=True)


def run(args, cpu_only: bool = False):
    import torch
    from bayesian_bm25_b200 import BayesianBM25Scorer, MultiFieldScorer, _lib, fused, hybrid, synthetic
    from oracle import coracle

    n_docs, n_q, k = _defaults(args)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    if B._env_int("WORLD_SIZE", 1) > 1:
        raise SystemExit("configs 4/5 are single-GPU configurations (BASELINE.json); run with --gpus 1")
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    t_build = time.perf_counter()
    terms, off = synthetic.zipf_queries(n_q, VOCAB, B.QUERY_SEED)
    qs = [terms[off[i]:off[i + 1]] for i in range(n_q)]
    if args.config == 5:
        fields, lens, seeds = ["title", "body"], {"title": 8.0, "body": 56.0}, {"title": 142, "body": 42}
        cscs = {f: synthetic.zipf_csc(n_docs, VOCAB, lens[f], seeds[f], dev, k1=1.2, b=0.75, method="lucene", min_len=2)
                for f in fields}
        pseudo = {f: synthetic.zipf_pseudo_queries(n_docs, VOCAB, lens[f], seeds[f], min_len=2) for f in fields}
        mf = MultiFieldScorer(fields, alpha="auto", base_rate="auto", k1=1.2, b=0.75, method="lucene")
        mf.index_from_csc(cscs, pseudo_queries=pseudo)
        scorers = [mf._scorers[f] for f in fields]
        weights, alpha = [0.5, 0.5], "auto"
        metric = "queries/sec (top-%d fused probs, MultiFieldScorer title+body, %.1fM docs, block-max pruned)" % (k, n_docs / 1e6)
        cos = d_cos = None
        cos_w = None
    else:
        fields = ["body"]
        cscs = {"body": synthetic.zipf_csc(n_docs, VOCAB, B.AVG_LEN, B.CORPUS_SEED, dev, k1=1.2, b=0.75, method="lucene")}
        pseudo = synthetic.zipf_pseudo_queries(n_docs, VOCAB, B.AVG_LEN, B.CORPUS_SEED)
        sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", base_rate="auto")
        sc.index_from_csc(cscs["body"], pseudo_queries=pseudo)
        scorers = [sc]
        weights, alpha, cos_w = [0.6], 0.5, 0.4
        metric = "queries/sec (top-%d hybrid BM25+cosine fused probs, %.1fM docs)" % (k, n_docs / 1e6)
        stride = (n_docs + 3) // 4 * 4
        g = torch.Generator(device=dev)
        g.manual_seed(44)
        d_cos = torch.empty((n_q, stride), dtype=torch.float32, device=dev)
        for s in range(0, n_q, 64):  # cos ~ clip(N(0.2, 0.15), -1, 1)
            e = min(n_q, s + 64)
            d_cos[s:e].normal_(0.2, 0.15, generator=g).clamp_(-1.0, 1.0)
        cos = None
    df = {f: (cscs[f]["indptr"][1:] - cscs[f]["indptr"][:-1]).cpu().numpy() for f in fields}
    nnz = {f: int(cscs[f]["data"].numel()) for f in fields}
    params = {}
    for f, s in zip(fields, scorers):
        t = s.transform
        params[f] = (float(t.alpha), float(t.beta), None if t.base_rate is None else float(t.base_rate))
    t_build = time.perf_counter() - t_build
    flat, qoff = fused._flat_queries(qs)
    fq = [(flat, qoff) for _ in fields]
    nf = len(fields)
    n_sig = nf + (1 if d_cos is not None else 0)
    from bayesian_bm25_b200.fusion import _resolve_alpha
    scale = float(n_sig ** _resolve_alpha(alpha, default=0.0 if args.config == 4 else 0.5))

    def step():
        return fused.retrieve_fused_batch_device(scorers, fq, k, scale, weights, d_cos, cos_w or 0.0)

    def e2e_step():
        ids, pr = step()  # term ids go up from the host inside the call; results come back to the host
        return ids.cpu().numpy(), pr.cpu().numpy()

    line = {"metric": metric, "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    # ---- CPU oracle: baseline + parity sample ---------------------------------------------------
    host_ok = True
    try:
        import psutil
        need = sum(nnz.values()) * 8 * 1.3
        host_ok = psutil.virtual_memory().available > need + 16e9
    except Exception:
        pass
    oracle_rows, cpu = None, None
    if not args.no_cpu or cpu_only:
        if host_ok:
            hosts = {f: _host(cscs[f]) for f in fields}
            sample_note = "full corpus"
            lo, hi = 0, n_docs
        else:
            from bayesian_bm25_b200 import index_build
            lo, hi = 0, n_docs // 10
            hosts = {f: _host(index_build.shard_csc(cscs[f], lo, hi)) for f in fields}
            sample_note = "first tenth of the documents (host memory)"
        oparams = {f: coracle.make_params(*params[f]) for f in fields}
        cores = coracle.max_threads()
        n_sample = args.cpu_sample or max(24, 2 * cores)
        cos_host = None if d_cos is None else d_cos[:n_sample, lo:hi].cpu().numpy()
        v, used, dt, oracle_rows = _cpu_leg(hosts, oparams, fields, weights, alpha, qs, k, n_sample, cos_host, cos_w)
        cpu = {"value": v * ((hi - lo) / n_docs), "unit": "queries/s", "cores": used, "kind": "port",
               "sample": f"{n_sample} queries, {sample_note}, oracle get_probabilities + log_odds_conjunction + top-k per query, "
                         f"{used} threads, {dt:.1f} s wall"}
        del hosts
    del cscs
    torch.cuda.empty_cache()
    if cpu_only:
        line.update({"impl": "reference", "value": cpu["value"], "ms_per_step": 1000.0 / cpu["value"] if cpu["value"] else None,
                     "config": {"workload": f"config {args.config}: {n_docs} docs, {VOCAB}-term Zipf vocab, top-{k}"},
                     "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return line

    def measure(level):
        scorers[0].set_pruning(level)
        for _ in range(max(args.warmup, 1)):
            out = step()
        torch.cuda.synchronize()
        launches0 = _lib.lib().bb25_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {}
        ev0.record()
        for _ in range(args.steps):
            out = step()
            for kk, vv in fused.fused_stats(scorers[0]).items():
                acc[kk] = acc.get(kk, 0) + vv
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        launches = int(_lib.lib().bb25_launch_count() - launches0)
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        return {"ms": ms, "launches": launches, "e2e_s": e2e_s, "out": out, **{a: b / args.steps for a, b in acc.items()}}

    sampler = B.ClockSampler(0)
    sampler.start()
    pruned = measure(3)
    clocks = sampler.stop()
    exhaustive = measure(0)
    no_sparse = None
    if args.config == 5 and os.environ.get("BB25_BENCH_AB_SPARSE", "1") != "0":
        # A/B inside the same process: level 3 without the essential-posting evaluation of units
        os.environ["BB25_FUSED_SPARSE"] = "0"
        try:
            no_sparse = measure(3)
        finally:
            del os.environ["BB25_FUSED_SPARSE"]
    same = all(bool(torch.equal(x, y)) for x, y in zip(pruned["out"], exhaustive["out"]))
    head = pruned if args.config == 5 else exhaustive  # config 5 is the pruned configuration; config 4 is exhaustive
    qps = n_q * args.steps / (head["ms"] / 1000.0)
    alg = sum(int(df[f][flat].sum()) * 8 for f in fields) + n_q * k * 16 + (n_q * n_docs * 4 if d_cos is not None else 0)
    hbm_peak, peak_src = B.measured_peak_gbs()
    peaks = B.measure_l2_peak(0)
    kernel_s = exhaustive["traverse_ms"] / 1000.0
    eff = alg / kernel_s / 1e9 if kernel_s > 0 else 0.0
    line.update({
        "value": qps, "ms_per_step": head["ms"] / args.steps,
        "config": {
            "workload": (f"BASELINE config {args.config}: {n_docs} docs, {VOCAB}-term Zipf vocab, fields {fields} nnz {nnz}, "
                         f"{n_q}-query batch (3-5 terms), top-{k} by fused probability"
                         + (", dense cosine row per query (device-resident)" if d_cos is not None else "")),
            "headline_mode": "block-max pruned (level 3)" if args.config == 5 else "exhaustive (level 0)",
            "cache": "inputs larger than L2 (%.1f GB of postings, 126 MB L2)" % (sum(nnz.values()) * 8 / 1e9),
            "transform": {f: dict(zip(("alpha", "beta", "base_rate"), params[f])) for f in fields},
            "weights": weights + ([cos_w] if cos_w else []), "alpha": alpha, "index_build_s": round(t_build, 1),
            "fallback_queries_per_step": head["fallback_queries"], "rerun_queries_per_step": head["rerun_queries"],
            "host_syncs_per_step": head["host_syncs"], "candidates_per_step": head["candidates"],
        },
        "clocks": clocks,
        "e2e": {"value": n_q * args.steps / head["e2e_s"], "unit": "queries/s",
                "h2d_bytes_per_step": int(nf * (flat.nbytes + qoff.nbytes)), "d2h_bytes_per_step": int(n_q * k * 16),
                "call": "retrieve_fused_batch_device: term ids up from the host, (ids, fused) back to the host"
                        + ("; cosine rows stay on the device (they are produced there)" if d_cos is not None else "")},
        "gpu_launches": head["launches"],
        "pruned": {"value": n_q * args.steps / (pruned["ms"] / 1000.0), "ms_per_step": pruned["ms"] / args.steps,
                   "kernel_ms_per_step": pruned["traverse_ms"], "units_per_step": pruned["units"],
                   "units_skipped_per_step": pruned["units_skipped"], "units_abandoned_per_step": pruned["units_abandoned"],
                   "units_no_essential_posting_per_step": pruned["units_no_essential"],
                   "units_by_essential_postings_per_step": pruned["units_sparse"],
                   "documents_evaluated_by_essential_postings_per_step": pruned["sparse_documents"],
                   "results_identical_to_exhaustive": same},
        "exhaustive": {"value": n_q * args.steps / (exhaustive["ms"] / 1000.0), "ms_per_step": exhaustive["ms"] / args.steps,
                       "kernel_ms_per_step": exhaustive["traverse_ms"], "units_per_step": exhaustive["units"]},
        "roofline": {
            "kernel": "bb25::fused_block_kernel (exhaustive pass)", "bound": "l1tex-lsu",
            "kernel_ms_per_step": exhaustive["traverse_ms"],
            "effective": {"value": eff, "unit": "GB/s", "vs_hbm_peak": eff / hbm_peak,
                          "what": "algorithmic bytes (sum_q sum_fields sum_t df*8 B + k*16 B"
                                  + (" + N*4 B of cosines per query" if d_cos is not None else "") + ") / traversal-kernel time"},
            "achieved": eff, "peak": hbm_peak, "unit": "GB/s", "frac": eff / hbm_peak, "traffic": None,
            "algorithmic_bytes_per_step": alg,
            "peaks": {"hbm_copy_gbs": hbm_peak, "hbm_copy_source": peak_src, **peaks},
            "note": "frac here is algorithmic bytes over the HBM copy peak (index slices are shared through L2, so it is an "
                    "effective figure, not a DRAM fraction); the ncu counters of this kernel are under profiles/r02",
        },
    })
    if no_sparse is not None:
        line["pruned_without_essential_evaluation"] = {
            "value": n_q * args.steps / (no_sparse["ms"] / 1000.0), "ms_per_step": no_sparse["ms"] / args.steps,
            "kernel_ms_per_step": no_sparse["traverse_ms"],
            "results_identical": all(bool(torch.equal(x, y)) for x, y in zip(no_sparse["out"], exhaustive["out"]))}
    if args.config == 4:
        # the dense side from embeddings: tcgen05 cosine GEMM (bb25_cosine_gemm) + the same fused retrieval
        try:
            from bayesian_bm25_b200 import dense
            kdim = 768
            del d_cos
            torch.cuda.empty_cache()
            gq = torch.Generator(device=dev)
            gq.manual_seed(45)
            ce = torch.empty((n_docs, kdim), dtype=torch.bfloat16, device=dev)
            for s_ in range(0, n_docs, 1 << 20):
                e_ = min(n_docs, s_ + (1 << 20))
                ce[s_:e_] = torch.nn.functional.normalize(torch.randn((e_ - s_, kdim), device=dev, generator=gq), dim=1).to(torch.bfloat16)
            qe_host = torch.nn.functional.normalize(torch.randn((n_q, kdim), generator=torch.Generator().manual_seed(46)), dim=1).to(torch.bfloat16).pin_memory()
            qe = qe_host.to(dev)
            cos_buf = dense.cosine_scores(qe[:256], ce)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            ev0.record()
            for _ in range(reps):
                for s_ in range(0, n_q, 256):
                    dense.cosine_scores(qe[s_:s_ + 256], ce, out=cos_buf[:min(256, n_q - s_)])
            ev1.record()
            torch.cuda.synchronize()
            gemm_ms = ev0.elapsed_time(ev1) / reps
            flops = 2.0 * n_q * n_docs * kdim
            peaks_json = {}
            pth = os.path.join(B.ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(pth):
                peaks_json = json.load(open(pth))
            tf_peak = float(peaks_json.get("bf16_tflops_sustained", 1400.0))
            want = (qe[:8].float() @ ce[:4096].float().T)
            got = dense.cosine_scores(qe[:8], ce[:4096])[:, :4096]
            gerr = float((got - want).abs().max())

            def hyb_step():
                return hybrid.hybrid_retrieve_batch_embeddings(scorers[0], flat, qoff, qe_host.to(dev, non_blocking=True), ce, k,
                                                               weights=(weights[0], cos_w), alpha=alpha, sub_batch=256)
            hyb_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                hyb_step()
            torch.cuda.synchronize()
            hyb_s = (time.perf_counter() - t0) / args.steps
            line["dense_side"] = {
                "what": "bb25_cosine_gemm: bf16 query_emb[%d,%d] @ corpus_emb[%d,%d]^T on tcgen05 / TMEM / TMA, fp32 cosine rows out; "
                        "random unit embeddings" % (n_q, kdim, n_docs, kdim),
                "gemm_ms_per_batch": gemm_ms, "tflops": flops / (gemm_ms * 1e-3) / 1e12, "peak_tflops": tf_peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks_json else "fallback",
                "frac_of_bf16_peak": flops / (gemm_ms * 1e-3) / 1e12 / tf_peak,
                "hbm_bytes_per_batch": int(-(-n_q // 256) * n_docs * kdim * 2 + n_q * n_docs * 4),
                "hbm_gbs": (-(-n_q // 256) * n_docs * kdim * 2 + n_q * n_docs * 4) / (gemm_ms * 1e-3) / 1e9,
                "max_abs_err_vs_torch_fp32": gerr,
                "hybrid_from_embeddings_qps_e2e": n_q / hyb_s,
                "hybrid_call": "hybrid_retrieve_batch_embeddings: query embeddings + term ids up from the host, GEMM, fused-rank batch retrieval, (ids, fused) back",
            }
        except Exception as e:  # the dense leg is an extra; never lose the bench line to it
            line["dense_side"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if cpu is not None:
        line["cpu_baseline"] = cpu
        ids = head["out"][0][:len(oracle_rows)].cpu().numpy()
        pr = head["out"][1][:len(oracle_rows)].cpu().numpy()
        if host_ok:
            ok_ids = all(np.array_equal(ids[i], oracle_rows[i][0]) or
                         np.max(np.abs(pr[i] - oracle_rows[i][1])) < 1e-12 for i in range(len(oracle_rows)))
            err = float(max(np.max(np.abs(pr[i] - oracle_rows[i][1])) for i in range(len(oracle_rows))))
            line["parity_spot_check"] = bool(ok_ids and err < 1e-9)
            line["parity_spot_check_detail"] = {"queries": len(oracle_rows), "ids_equal": bool(ok_ids), "max_abs_prob_err": err}
        else:
            line["parity_spot_check"] = None
    print(json.dumps(line), flush=True)
    return line
