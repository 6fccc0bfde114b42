#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's
`probability.py`, `fusion.py`, `scorer.py`, `multi_field.py` are imported as
shipped; the third-party `bm25s` they need is absent from the image, so
`oracle/bm25s_equiv.py` is registered under that module name (the
five-call surface).  Everything after the BM25 score is therefore the
reference's own arithmetic.

    python tests/golden/make_golden.py

The outputs are committed; tests never read /root/reference.
"""
from __future__ import annotations

import importlib.metadata as _md
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

_orig_version = _md.version
_md.version = lambda name: "0.12.1" if name == "bayesian-bm25" else _orig_version(name)

from oracle import bm25s_equiv  # noqa: E402

sys.modules["bm25s"] = bm25s_equiv

from bayesian_bm25.fusion import cosine_to_probability, log_odds_conjunction  # noqa: E402
from bayesian_bm25.multi_field import MultiFieldScorer  # noqa: E402
from bayesian_bm25.probability import BayesianProbabilityTransform, logit, sigmoid  # noqa: E402
from bayesian_bm25.scorer import BayesianBM25Scorer, BlockMaxIndex  # noqa: E402
from benchmarks.scalability import generate_synthetic_corpus  # noqa: E402


def probability_fusion():
    rng = np.random.default_rng(2024)
    out = {}
    # --- SURVEY 8c G1..G10 (inputs stored beside outputs) ---
    t = BayesianProbabilityTransform(1.5, 1.0, base_rate=0.01)
    out["g1"] = t.score_to_probability(np.array([.5, 1, 1.5, 2, 3]), np.array([1, 2, 3, 5, 8.]),
                                       np.array([.3, .5, .8, 1, 1.5]))
    t2 = BayesianProbabilityTransform(1.0, 0.0)
    out["g2"] = t2.score_to_probability(np.array([1.0464478, 0.56150854, 1.1230172]),
                                        np.array([5., 3., 7.]), np.array([.5, .5, .5]))
    out["g3"] = BayesianProbabilityTransform.composite_prior(np.array([0, 1, 2, 5, 10, 50.]), 0.8)
    out["g4"] = np.array([BayesianProbabilityTransform(1.5, 2.0, .01).wand_upper_bound(5.0),
                          BayesianProbabilityTransform(1.5, 2.0).wand_upper_bound(5.0)])
    out["g5"] = np.array([BayesianProbabilityTransform.posterior(0.7, 0.5, base_rate=0.01)])
    t6 = BayesianProbabilityTransform(2.020913298430858, 0.2106551229953766, .04501)
    out["g6"] = t6.score_to_probability(np.array([3.1415927, .001, 12.5, 25, 60], dtype=np.float32),
                                        np.array([1, 2, 3, 4, 4.]), np.array([.2, .7, 1, 1.3, 3]))
    out["g7"] = np.array([log_odds_conjunction(np.array([.85, .7, .6]))])
    p8 = np.stack([np.array([.85, .6, .4]), cosine_to_probability(np.array([.92, .35, .7]))], axis=-1)
    out["g8_in"] = p8
    out["g8a"] = log_odds_conjunction(p8, weights=np.array([.6, .4]))
    out["g8b"] = log_odds_conjunction(p8, alpha=.5, weights=np.array([.6, .4]))
    out["g9"] = np.array([
        log_odds_conjunction(np.array([0.0, cosine_to_probability(0.3)]), alpha=.5, weights=np.array([.5, .5])),
        log_odds_conjunction(np.array([0.0, 0.0]), alpha=.5, weights=np.array([.5, .5])),
        logit(1e-10)])
    out["g10"] = np.array([log_odds_conjunction(p8, gating=g)[0] for g in ("relu", "swish", "gelu", "softplus")])

    # --- random sweeps ---
    n = 4096
    s = np.concatenate([rng.uniform(0, 30, n - 6), [0.0, 1e-30, 40.0, 80.0, 700.0, 1e-8]]).astype(np.float32)
    tf = rng.integers(0, 14, n).astype(np.float64)
    r = np.concatenate([rng.uniform(0, 4, n - 4), [0.0, 0.5, 1.0, 25.0]])
    out["sweep_score"] = s
    out["sweep_tf"] = tf
    out["sweep_ratio"] = r
    params = [(1.0, 0.0, None), (2.020913298430858, 0.2106551229953766, 0.04501),
              (0.35, 6.5, 0.001), (5.0, 2.0, 0.5), (0.01, -3.0, 0.3)]
    out["sweep_params"] = np.array([[a, b, -1.0 if br is None else br] for a, b, br in params])
    for i, (a, b, br) in enumerate(params):
        tt = BayesianProbabilityTransform(a, b, base_rate=br)
        out[f"sweep_prob_{i}"] = tt.score_to_probability(s, tf, r)
        out[f"sweep_wand_{i}"] = tt.wand_upper_bound(s.astype(np.float64), 0.9)
        out[f"sweep_like_{i}"] = tt.likelihood(s)
        tt._training_mode = "prior_free"
        out[f"sweep_priorfree_{i}"] = tt.score_to_probability(s, tf, r)
    out["sweep_tf_prior"] = BayesianProbabilityTransform.tf_prior(tf)
    out["sweep_norm_prior"] = BayesianProbabilityTransform.norm_prior(r)
    out["sweep_composite"] = BayesianProbabilityTransform.composite_prior(tf, r)
    lk = rng.uniform(0, 1, n)
    pr = rng.uniform(0, 1, n)
    out["post_l"] = lk
    out["post_p"] = pr
    out["post_nobr"] = BayesianProbabilityTransform.posterior(lk, pr)
    out["post_br"] = BayesianProbabilityTransform.posterior(lk, pr, base_rate=0.02)
    x = np.concatenate([rng.uniform(-50, 50, n - 5), [-800., 800., 0., -36.8, 36.8]])
    out["sig_x"] = x
    out["sig_y"] = sigmoid(x)
    pp = np.concatenate([rng.uniform(0, 1, n - 4), [0., 1., 1e-12, 1 - 1e-12]])
    out["logit_p"] = pp
    out["logit_y"] = logit(pp)
    cs = np.concatenate([rng.uniform(-1, 1, n - 3), [-1., 1., 0.]])
    out["cos_x"] = cs
    out["cos_y"] = cosine_to_probability(cs)

    # log_odds_conjunction sweeps
    for nsig in (1, 2, 3, 5, 9):
        P = rng.uniform(0, 1, (257, nsig))
        P[0, :] = 0.0
        P[1, :] = 1.0
        out[f"loc_in_{nsig}"] = P
        w = rng.uniform(0.1, 1, nsig)
        w /= w.sum()
        out[f"loc_w_{nsig}"] = w
        out[f"loc_unw_{nsig}"] = log_odds_conjunction(P)
        out[f"loc_unw_a0_{nsig}"] = log_odds_conjunction(P, alpha=0.0)
        out[f"loc_unw_auto_{nsig}"] = log_odds_conjunction(P, alpha="auto")
        out[f"loc_w_{nsig}_none"] = log_odds_conjunction(P, weights=w)
        out[f"loc_w_{nsig}_a05"] = log_odds_conjunction(P, alpha=0.5, weights=w)
        for g in ("relu", "swish", "gelu", "softplus"):
            out[f"loc_{g}_{nsig}"] = log_odds_conjunction(P, alpha=0.5, weights=w, gating=g)
            out[f"loc_{g}_b2_{nsig}"] = log_odds_conjunction(P, gating=g, gating_beta=2.0)
        out[f"loc_clip_{nsig}"] = log_odds_conjunction(P, weights=w, max_logit=3.0)
    np.savez_compressed(os.path.join(HERE, "probability_fusion.npz"), **out)
    print("probability_fusion.npz", len(out), "arrays")


def _scorer_case(name, corpus, queries, method, base_rate, br_method, ks, store_csc=True,
                 n_dense=None):
    sc = BayesianBM25Scorer(k1=1.2, b=0.75, method=method, base_rate=base_rate,
                            base_rate_method=br_method)
    sc.index(corpus, show_progress=False)
    bm = sc._bm25
    vocab = bm.vocab_dict
    out = {}
    meta = {
        "name": name, "method": method, "k1": 1.2, "b": 0.75,
        "alpha": float(sc._transform.alpha), "beta": float(sc._transform.beta),
        "base_rate": None if sc._transform.base_rate is None else float(sc._transform.base_rate),
        "base_rate_arg": base_rate, "base_rate_method": br_method,
        "avgdl": float(sc.avgdl), "num_docs": int(sc.num_docs),
        "nnz": int(len(bm.scores["data"])), "ks": list(ks),
        "queries": queries,
    }
    if store_csc:
        out["data"] = bm.scores["data"]
        out["indices"] = bm.scores["indices"]
        out["indptr"] = bm.scores["indptr"]
        out["doc_len"] = bm.scores["doc_len"]
        meta["corpus"] = corpus
    # queries as in-vocabulary term ids (what reaches the engine)
    qt = [[vocab[t] for t in q if t in vocab] for q in queries]
    out["q_terms"] = np.array([t for q in qt for t in q], dtype=np.int32)
    out["q_off"] = np.cumsum([0] + [len(q) for q in qt]).astype(np.int64)
    for k in ks:
        if k > sc.num_docs:
            continue
        ids, probs = sc.retrieve(queries, k=k)
        res = bm.retrieve(queries, k=k)
        out[f"ids_k{k}"] = ids.astype(np.int64)
        out[f"probs_k{k}"] = probs
        out[f"scores_k{k}"] = res.scores
    nd = len(queries) if n_dense is None else min(n_dense, len(queries))
    meta["n_dense"] = nd
    if nd:
        out["dense_probs"] = np.stack([sc.get_probabilities(q) for q in queries[:nd]])
        out["dense_scores"] = np.stack([bm.get_scores(q) for q in queries[:nd]])
    return out, meta


def scorer_cases():
    toy = [
        ["the", "cat", "sat", "on", "the", "mat"],
        ["the", "dog", "chased", "the", "cat"],
        ["a", "quick", "brown", "fox", "jumps", "over", "the", "lazy", "dog"],
        ["hello", "world"],
        ["machine", "learning", "is", "a", "subset", "of", "artificial", "intelligence"],
        ["the", "cat", "and", "the", "dog", "are", "friends"],
    ]
    toy_q = [["cat"], ["dog"], ["machine", "learning"], [], ["xyznonexistent"],
             ["the", "cat", "the"], ["the", "cat", "dog", "fox", "a"], ["zzz", "hello", "qqq"]]
    metas = []
    arrays = {}
    idx = 0
    for method in ("lucene", "robertson", "atire"):
        for br, brm in ((None, "percentile"), ("auto", "percentile"), (0.01, "percentile")):
            o, m = _scorer_case(f"toy_{method}_{br}", toy, toy_q, method, br, brm, ks=(1, 3, 6))
            for key, v in o.items():
                arrays[f"c{idx}_{key}"] = v
            m["prefix"] = f"c{idx}_"
            metas.append(m)
            idx += 1
    corpus, queries = generate_synthetic_corpus(300, 200, 30, np.random.default_rng(7))
    queries = queries + [queries[0] + queries[1] + queries[0], ["term_0", "term_0", "term_1", "nope", "term_2", "term_3", "term_4", "term_5", "term_199"]]
    for method, br, brm in (("lucene", "auto", "percentile"), ("robertson", "auto", "mixture"),
                            ("atire", "auto", "elbow"), ("robertson", None, "percentile")):
        o, m = _scorer_case(f"zipf300_{method}_{brm}", corpus, queries, method, br, brm,
                            ks=(10, 50, 300), n_dense=8)
        for key, v in o.items():
            arrays[f"c{idx}_{key}"] = v
        m["prefix"] = f"c{idx}_"
        metas.append(m)
        idx += 1
    np.savez_compressed(os.path.join(HERE, "scorer_cases.npz"), **arrays)
    with open(os.path.join(HERE, "scorer_cases.json"), "w") as f:
        json.dump(metas, f)
    print("scorer_cases", idx, "cases")


def config1():
    """BASELINE configs[0]: benchmarks/scalability.py on 10k docs, seed 42."""
    corpus, queries = generate_synthetic_corpus(10_000, 10_000, 100, np.random.default_rng(42))
    o, m = _scorer_case("config1", corpus, queries, "lucene", "auto", "percentile", ks=(10,),
                        store_csc=False, n_dense=3)
    # keep the fixture small: drop the dense score rows, keep dense prob rows as
    # sparse (index, value) pairs
    dp = o.pop("dense_probs")
    o.pop("dense_scores")
    nz = [np.nonzero(row)[0] for row in dp]
    o["dense_nz_off"] = np.cumsum([0] + [len(z) for z in nz]).astype(np.int64)
    o["dense_nz_idx"] = np.concatenate(nz).astype(np.int32)
    o["dense_nz_val"] = np.concatenate([row[z] for row, z in zip(dp, nz)])
    o["doc_len_sum"] = np.array([sum(len(d) for d in corpus)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), **o)
    with open(os.path.join(HERE, "config1.json"), "w") as f:
        json.dump(m, f)
    print("config1: nnz", m["nnz"], "alpha", m["alpha"], "beta", m["beta"], "br", m["base_rate"])


def multifield_blockmax():
    rng = np.random.default_rng(11)
    corpus_b, queries = generate_synthetic_corpus(200, 120, 25, rng)
    corpus_t, _ = generate_synthetic_corpus(200, 120, 6, rng)
    docs = [{"title": t, "body": b} for t, b in zip(corpus_t, corpus_b)]
    out = {}
    meta = {"docs": docs, "queries": queries, "cases": []}
    i = 0
    for alpha, fw, br in (("auto", None, None), (0.0, {"title": 0.7, "body": 0.3}, "auto"),
                          (1.0, {"title": 0.25, "body": 0.75}, 0.05)):
        mf = MultiFieldScorer(["title", "body"], field_weights=fw, alpha=alpha, base_rate=br,
                              method="lucene")
        mf.index(docs, show_progress=False)
        out[f"mf{i}_probs"] = np.stack([mf.get_probabilities(q) for q in queries])
        meta["cases"].append({
            "alpha": alpha, "field_weights": fw, "base_rate": br,
            "field_alpha": {f: float(mf._scorers[f]._transform.alpha) for f in mf.fields},
            "field_beta": {f: float(mf._scorers[f]._transform.beta) for f in mf.fields},
            "field_base_rate": {f: mf._scorers[f]._transform.base_rate for f in mf.fields},
        })
        i += 1
    # BlockMaxIndex on a dense matrix (tests/test_bmw.py pattern)
    sm = rng.uniform(0, 5, (7, 1000))
    for bs in (1, 7, 128, 1000, 4096):
        bmi = BlockMaxIndex(block_size=bs)
        bmi.build(sm)
        out[f"bmw_bs{bs}"] = bmi._block_maxes
    out["bmw_matrix"] = sm
    t = BayesianProbabilityTransform(1.3, 1.7, base_rate=0.03)
    bmi = BlockMaxIndex(block_size=128)
    bmi.build(sm)
    out["bmw_bayes_bs128"] = np.array([[bmi.bayesian_block_upper_bound(ti, b, t) for b in range(bmi.n_blocks)]
                                       for ti in range(7)])
    np.savez_compressed(os.path.join(HERE, "multifield_blockmax.npz"), **out)
    with open(os.path.join(HERE, "multifield_blockmax.json"), "w") as f:
        json.dump(meta, f)
    print("multifield_blockmax ok")


def extras():
    """balanced_log_odds_fusion, AttentionLogOddsWeights inference
    (__call__ / compute_upper_bounds / prune), retrieve(explain=True) traces -- outputs of the reference."""
    from bayesian_bm25.debug import FusionDebugger
    from bayesian_bm25.fusion import AttentionLogOddsWeights, balanced_log_odds_fusion
    rng = np.random.default_rng(77)
    out = {}
    meta = {"attention": [], "explain": {}}
    # balanced fusion
    sp = rng.uniform(0, 1, 400)
    sp[:5] = [0.0, 1.0, 1e-12, 0.5, 0.999999]
    de = rng.uniform(-1, 1, 400)
    de[:3] = [-1.0, 1.0, 0.0]
    out["bal_sparse"], out["bal_dense"] = sp, de
    for w in (0.5, 0.3, 0.0, 1.0):
        out[f"bal_w{w}"] = balanced_log_odds_fusion(sp, de, weight=w)
    out["bal_const_sparse"] = balanced_log_odds_fusion(np.full(50, 0.3), de[:50], weight=0.4)
    out["bal_const_dense"] = balanced_log_odds_fusion(sp[:50], np.full(50, 0.2), weight=0.4)
    # attention weights (random "trained" parameters)
    for ci, (n_sig, n_qf, alpha, norm, br) in enumerate(((3, 4, 0.5, False, None), (2, 6, "auto", True, None),
                                                        (4, 3, 0.0, False, 0.1), (3, 5, 1.0, True, 0.02))):
        a = AttentionLogOddsWeights(n_sig, n_qf, alpha=alpha, normalize=norm, seed=3 + ci, base_rate=br)
        out[f"att{ci}_W_init"] = a._W.copy()
        a._W = rng.normal(0, 1.0, (n_sig, n_qf))
        a._b = rng.normal(0, 0.5, n_sig)
        a._W_avg = a._W * 0.9
        a._b_avg = a._b * 1.1
        m = 64
        P = rng.uniform(0, 1, (m, n_sig))
        P[0, :] = 0.0
        P[1, :] = 1.0
        qf1 = rng.normal(0, 1, n_qf)
        qfm = rng.normal(0, 1, (m, n_qf))
        UB = np.minimum(1.0, P + rng.uniform(0, 0.3, (m, n_sig)))
        pre = f"att{ci}_"
        out[pre + "W"], out[pre + "b"], out[pre + "P"], out[pre + "qf1"], out[pre + "qfm"], out[pre + "UB"] = a._W, a._b, P, qf1, qfm, UB
        out[pre + "w1"] = a._compute_weights(qf1)
        out[pre + "wm"] = a._compute_weights(qfm)
        out[pre + "wm_avg"] = a._compute_weights(qfm, use_averaged=True)
        out[pre + "call_single"] = np.array([a(P[5], qf1)])
        out[pre + "call_batch_q1"] = a(P, qf1)
        out[pre + "call_batch_qm"] = a(P, qfm)
        out[pre + "call_batch_avg"] = a(P, qfm, use_averaged=True)
        out[pre + "ub_q1"] = a.compute_upper_bounds(UB, qf1)
        out[pre + "ub_qm"] = a.compute_upper_bounds(UB, qfm)
        thr = float(np.median(out[pre + "ub_qm"]))
        idx, fused = a.prune(P, qfm, thr, upper_bound_probs=UB)
        out[pre + "prune_idx"], out[pre + "prune_fused"] = idx.astype(np.int64), fused
        idx2, fused2 = a.prune(P, qf1, thr)
        out[pre + "prune2_idx"], out[pre + "prune2_fused"] = idx2.astype(np.int64), fused2
        meta["attention"].append({"n_signals": n_sig, "n_query_features": n_qf, "alpha": alpha, "normalize": norm,
                                  "base_rate": br, "seed": 3 + ci, "threshold": thr})
    # trace_bm25 sweep
    s = rng.uniform(0, 25, 300)
    tf = rng.integers(0, 13, 300).astype(np.float64)
    r = rng.uniform(0, 3, 300)
    out["tr_s"], out["tr_tf"], out["tr_r"] = s, tf, r
    for ti, (a_, b_, br_) in enumerate(((1.0, 0.0, None), (2.02, 0.21, 0.045), (0.4, 6.0, 0.5))):
        dbg = FusionDebugger(BayesianProbabilityTransform(a_, b_, base_rate=br_))
        rows = []
        for i in range(300):
            t = dbg.trace_bm25(float(s[i]), float(tf[i]), float(r[i]))
            rows.append([t.likelihood, t.tf_prior, t.norm_prior, t.composite_prior, t.logit_likelihood, t.logit_prior,
                         t.posterior, np.nan if t.logit_base_rate is None else t.logit_base_rate])
        out[f"tr_out{ti}"] = np.array(rows)
    meta["trace_params"] = [[1.0, 0.0, None], [2.02, 0.21, 0.045], [0.4, 6.0, 0.5]]
    # retrieve(explain=True) on a small corpus
    corpus, queries = generate_synthetic_corpus(300, 200, 30, np.random.default_rng(7))
    queries = queries[:12] + [[], ["nope"], ["term_0", "term_0", "term_3"]]
    sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", base_rate="auto")
    sc.index(corpus, show_progress=False)
    res = sc.retrieve(queries, k=10, explain=True)
    out["ex_ids"], out["ex_probs"] = res.doc_ids.astype(np.int64), res.probabilities
    ex = np.full((len(queries), 10, 11), np.nan)
    for qi, row in enumerate(res.explanations):
        for ri, t in enumerate(row):
            if t is not None:
                ex[qi, ri] = [t.raw_score, t.tf, t.doc_len_ratio, t.likelihood, t.tf_prior, t.norm_prior, t.composite_prior,
                              t.logit_likelihood, t.logit_prior, t.logit_base_rate, t.posterior]
    out["ex_traces"] = ex
    meta["explain"] = {"corpus_seed": 7, "queries": queries, "alpha": float(sc._transform.alpha),
                       "beta": float(sc._transform.beta), "base_rate": float(sc._transform.base_rate)}
    np.savez_compressed(os.path.join(HERE, "extras.npz"), **out)
    with open(os.path.join(HERE, "extras.json"), "w") as f:
        json.dump(meta, f)
    print("extras.npz", len(out), "arrays")


if __name__ == "__main__":
    if "--extras-only" in sys.argv:
        extras()
        sys.exit(0)
    probability_fusion()
    scorer_cases()
    config1()
    multifield_blockmax()
    extras()
