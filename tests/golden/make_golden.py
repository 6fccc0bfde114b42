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


if __name__ == "__main__":
    probability_fusion()
    scorer_cases()
    config1()
    multifield_blockmax()
