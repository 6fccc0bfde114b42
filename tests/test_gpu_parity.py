"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
committed golden vectors.  Bars (BASELINE.json north_star): top-k ids bit-exact with
ties by doc id, fp32 scores bit-exact (<= 1e-5 rel allowed), probabilities within
1e-6 absolute -- asserted here at PROB_TOL = 1e-9."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-9

def _pkg():
    import bayesian_bm25_b200 as pkg
    return pkg


def _host(csc):
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)


# ---------------------------------------------------------------------------------
# elementwise probability / fusion surfaces vs vectors produced by the reference
# ---------------------------------------------------------------------------------
def test_probability_transform_vs_reference(golden_pf):
    pkg, g = _pkg(), golden_pf
    T = pkg.BayesianProbabilityTransform
    s, tf, r = g["sweep_score"], g["sweep_tf"], g["sweep_ratio"]
    for i, (a, b, br) in enumerate(g["sweep_params"]):
        t = T(a, b, base_rate=None if br < 0 else float(br))
        np.testing.assert_allclose(t.score_to_probability(s, tf, r), g[f"sweep_prob_{i}"], rtol=0, atol=PROB_TOL)
        np.testing.assert_allclose(t.wand_upper_bound(s.astype(np.float64)), g[f"sweep_wand_{i}"], rtol=0, atol=PROB_TOL)
        np.testing.assert_allclose(t.likelihood(s), g[f"sweep_like_{i}"], rtol=0, atol=PROB_TOL)
        t._training_mode = "prior_free"
        np.testing.assert_allclose(t.score_to_probability(s, tf, r), g[f"sweep_priorfree_{i}"], rtol=0, atol=PROB_TOL)
    np.testing.assert_allclose(T.tf_prior(tf), g["sweep_tf_prior"], atol=1e-15)
    np.testing.assert_allclose(T.norm_prior(r), g["sweep_norm_prior"], atol=1e-15)
    np.testing.assert_allclose(T.composite_prior(tf, r), g["sweep_composite"], atol=1e-15)
    np.testing.assert_allclose(T.posterior(g["post_l"], g["post_p"]), g["post_nobr"], atol=1e-15)
    np.testing.assert_allclose(T.posterior(g["post_l"], g["post_p"], base_rate=0.02), g["post_br"], atol=1e-15)
    np.testing.assert_allclose(pkg.sigmoid(g["sig_x"]), g["sig_y"], atol=1e-15)
    np.testing.assert_allclose(pkg.logit(g["logit_p"]), g["logit_y"], atol=1e-12)
    np.testing.assert_allclose(pkg.cosine_to_probability(g["cos_x"]), g["cos_y"], atol=0)
    # SURVEY 8c known answers + scalar-in/float-out convention (tests/test_wand.py:109-120)
    t = T(1.5, 1.0, base_rate=0.01)
    np.testing.assert_allclose(t.score_to_probability(np.array([.5, 1, 1.5, 2, 3]), np.array([1, 2, 3, 5, 8.]),
                                                      np.array([.3, .5, .8, 1, 1.5])), g["g1"], atol=PROB_TOL)
    assert isinstance(t.wand_upper_bound(5.0), float)
    assert isinstance(t.score_to_probability(1.0, 2.0, 0.5), float)
    assert T(1.5, 2.0, .01).wand_upper_bound(5.0) == pytest.approx(0.8911075788989186, abs=PROB_TOL)
    assert T.posterior(0.7, 0.5, base_rate=0.01) == pytest.approx(0.02302631578947368, abs=1e-15)
    assert T.tf_prior(0) == pytest.approx(0.2) and T.tf_prior(10) == pytest.approx(0.9)
    assert T.norm_prior(0.5) == pytest.approx(0.9) and T.norm_prior(1.0) == pytest.approx(0.3)
    # custom prior_fn runs on the host, Bayes update on the device
    tp = T(1.0, 0.0, prior_fn=lambda s_, tf_, r_: np.full(np.shape(s_), 0.5))
    np.testing.assert_allclose(tp.score_to_probability(s, tf, r), T(1.0, 0.0).likelihood(s), atol=PROB_TOL)


@pytest.mark.parametrize("nsig", [1, 2, 3, 5, 9])
def test_log_odds_conjunction_vs_reference(golden_pf, nsig):
    pkg, g = _pkg(), golden_pf
    loc = pkg.log_odds_conjunction
    P, w = g[f"loc_in_{nsig}"], g[f"loc_w_{nsig}"]
    chk = lambda got, key: np.testing.assert_allclose(got, g[key], rtol=0, atol=PROB_TOL)
    chk(loc(P), f"loc_unw_{nsig}")
    chk(loc(P, alpha=0.0), f"loc_unw_a0_{nsig}")
    chk(loc(P, alpha="auto"), f"loc_unw_auto_{nsig}")
    chk(loc(P, weights=w), f"loc_w_{nsig}_none")
    chk(loc(P, alpha=0.5, weights=w), f"loc_w_{nsig}_a05")
    for gt in ("relu", "swish", "gelu", "softplus"):
        chk(loc(P, alpha=0.5, weights=w, gating=gt), f"loc_{gt}_{nsig}")
        chk(loc(P, gating=gt, gating_beta=2.0), f"loc_{gt}_b2_{nsig}")
    chk(loc(P, weights=w, max_logit=3.0), f"loc_clip_{nsig}")
    if nsig == 3:
        assert loc(np.array([.85, .7, .6])) == pytest.approx(0.8487403513785625, abs=PROB_TOL)
        with pytest.raises(ValueError, match="non-negative"):
            loc(P, weights=[-0.1, 0.6, 0.5])
        with pytest.raises(ValueError, match="sum to 1"):
            loc(P, weights=[0.3, 0.3, 0.3])
        with pytest.raises(ValueError, match="gating"):
            loc(P, gating="tanh")
    if nsig == 2:
        np.testing.assert_allclose(loc(g["g8_in"], weights=[.6, .4]), g["g8a"], atol=PROB_TOL)
        np.testing.assert_allclose(loc(g["g8_in"], alpha=.5, weights=[.6, .4]), g["g8b"], atol=PROB_TOL)
        # BM25-inactive document in a fusion
        assert loc(np.array([0.0, 0.0]), alpha=.5, weights=[.5, .5]) == pytest.approx(7.208823232633578e-15, rel=1e-9)


# ---------------------------------------------------------------------------------
# scorer vs outputs of the reference's scorer.py (string API, full index() path)
# ---------------------------------------------------------------------------------
def test_scorer_cases_vs_reference(golden_scorer):
    pkg = _pkg()
    arrays, metas = golden_scorer
    for m in metas:
        p = m["prefix"]
        sc = pkg.BayesianBM25Scorer(k1=m["k1"], b=m["b"], method=m["method"], base_rate=m["base_rate_arg"],
                                    base_rate_method=m["base_rate_method"])
        sc.index(m["corpus"], show_progress=False)
        assert sc.num_docs == m["num_docs"]
        assert sc.avgdl == m["avgdl"]
        np.testing.assert_array_equal(sc.doc_lengths, arrays[p + "doc_len"])
        assert sc.transform.alpha == m["alpha"], m["name"]
        assert sc.transform.beta == m["beta"], m["name"]
        if m["base_rate"] is None:
            assert sc.base_rate is None
        else:
            assert sc.base_rate == pytest.approx(m["base_rate"], rel=1e-12)
        for k in m["ks"]:
            if f"{p}ids_k{k}" not in arrays:
                with pytest.raises(ValueError):
                    sc.retrieve(m["queries"], k=k)
                continue
            ids, probs = sc.retrieve(m["queries"], k=k)
            assert ids.shape == probs.shape == (len(m["queries"]), k)
            np.testing.assert_array_equal(ids, arrays[f"{p}ids_k{k}"], err_msg=f"{m['name']} k={k}")
            np.testing.assert_allclose(probs, arrays[f"{p}probs_k{k}"], rtol=0, atol=PROB_TOL)
        for i in range(m["n_dense"]):
            dense = sc.get_probabilities(m["queries"][i])
            np.testing.assert_allclose(dense, arrays[p + "dense_probs"][i], rtol=0, atol=PROB_TOL)
            assert np.array_equal(dense == 0.0, arrays[p + "dense_probs"][i] == 0.0)  # exact zeros
            qt = arrays[p + "q_terms"][arrays[p + "q_off"][i]:arrays[p + "q_off"][i + 1]]
            np.testing.assert_array_equal(sc.get_scores_ids(sc._term_ids(m["queries"][i])),
                                          arrays[p + "dense_scores"][i])
            assert len(qt) == len(sc._term_ids(m["queries"][i]))


def test_config1_vs_reference(golden_config1):
    """BASELINE configs[0]: the reference's scalability corpus, 10k docs, 100 queries, k=10."""
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    arrays, meta = golden_config1
    corpus, queries = synthetic.scalability_corpus(10_000, 10_000, 100, np.random.default_rng(42))
    sc = pkg.BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", base_rate="auto")
    sc.index(corpus, show_progress=False)
    assert sc.transform.alpha == meta["alpha"] and sc.transform.beta == meta["beta"]
    assert sc.base_rate == pytest.approx(meta["base_rate"], rel=1e-12)
    ids, probs = sc.retrieve(queries, k=10)
    np.testing.assert_array_equal(ids, arrays["ids_k10"])
    np.testing.assert_allclose(probs, arrays["probs_k10"], rtol=0, atol=PROB_TOL)
    off = arrays["dense_nz_off"]
    for i in range(meta["n_dense"]):
        dense = sc.get_probabilities(queries[i])
        nz = np.nonzero(dense)[0]
        np.testing.assert_array_equal(nz, arrays["dense_nz_idx"][off[i]:off[i + 1]])
        np.testing.assert_allclose(dense[nz], arrays["dense_nz_val"][off[i]:off[i + 1]], rtol=0, atol=PROB_TOL)
    # reference edge cases (tests/test_scorer.py:328-344)
    ids, probs = sc.retrieve([[], ["xyznonexistent"]], k=3)
    assert probs.shape == (2, 3) and np.all(probs == 0.0)
    np.testing.assert_array_equal(ids, [[0, 1, 2], [0, 1, 2]])
    assert np.all(sc.get_probabilities(["xyznonexistent"]) == 0.0)
    res = sc.retrieve(queries[:1], k=3, explain=True)  # scorer.py:538-562; traces checked in test_gpu_extras.py
    assert isinstance(res, pkg.RetrievalResult) and len(res.explanations) == 1 and len(res.explanations[0]) == 3
    n0 = sc.num_docs
    sc.add_documents([["term_1", "brandnewtoken"]], show_progress=False)
    assert sc.num_docs == n0 + 1
    ids, _ = sc.retrieve([["brandnewtoken"]], k=1)
    assert ids[0, 0] == n0


# ---------------------------------------------------------------------------------
# batch retrieve vs the oracle on seeded synthetic corpora, all tile sizes
# ---------------------------------------------------------------------------------
def _queries_with_edge_cases(vocab, seed):
    from bayesian_bm25_b200 import synthetic
    qt, qo = synthetic.zipf_queries(96, vocab, seed)
    qs = [qt[qo[i]:qo[i + 1]].tolist() for i in range(96)]
    rng = np.random.default_rng(seed + 1)
    qs[3] = []                                           # empty query
    qs[5] = [0, 0, 0]                                    # duplicates of the head term
    qs[7] = [0, 1, 2, 3, 4]                              # all-head query (loose threshold seed)
    qs[9] = [vocab - 1]                                  # rarest term only
    qs[11] = rng.integers(0, vocab, 300).tolist()        # long query, > 255 distinct terms
    qs[13] = [5, 9, 5, 2, 9, 9, 700 % vocab]             # repeated terms in mixed order
    qs[95] = []                                          # trailing empty query
    flat = np.array([t for q in qs for t in q], dtype=np.int32)
    off = np.cumsum([0] + [len(q) for q in qs]).astype(np.int64)
    return flat, off


@pytest.mark.parametrize("kernel,tile_docs,prune,tab_sparse", [
    ("block", 8192, 3, "0"), ("block", 8192, 2, "0"), ("block", 8192, 1, "0"), ("block", 8192, 0, "0"),
    ("block", 8192, 3, "1"), ("block", 8192, 0, "2"), ("block", 8192, 3, "2"),
    ("tile", 8192, 1, "0"), ("tile", 16384, 1, "0"), ("tile", 32768, 1, "0")])
def test_retrieve_batch_vs_oracle(kernel, tile_docs, prune, tab_sparse, monkeypatch):
    """Both traversal kernels (warp-private blocks at every pruning level -- exhaustive,
    block-max skip, + frequent-term bound, + candidate-driven queries -- with the block table
    in its dense, mixed and all-bitmap forms, and CTA tiles at every tile size) must give the
    oracle's result bit for bit."""
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    monkeypatch.setenv("BB25_KERNEL", kernel)
    monkeypatch.setenv("BB25_PRUNE", str(prune))
    monkeypatch.setenv("BB25_TILE_DOCS", str(tile_docs))
    monkeypatch.setenv("BB25_TAB_SPARSE", tab_sparse)
    n_docs, vocab = 150_001, (60_000 if tab_sparse == "1" else 4000)  # rare terms must touch few blocks for the mixed form
    csc = synthetic.zipf_csc(n_docs, vocab, 48.0, seed=21, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(method="lucene", alpha=2.1, beta=0.3, base_rate=0.03)
    sc.index_from_csc(csc)
    n_sparse = sc.index_info()["block_table_bitmap_terms"]
    assert (n_sparse == 0) if tab_sparse == "0" else (n_sparse == vocab if tab_sparse == "2" else 0 < n_sparse < vocab)
    flat, off = _queries_with_edge_cases(vocab, seed=22)
    params = coracle.make_params(2.1, 0.3, 0.03)
    for k in (1, 10, 100, 1000, 4096):
        ids, scores, probs = sc.retrieve_ids(flat, off, k, return_scores=True)
        o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
        np.testing.assert_array_equal(ids, o_ids, err_msg=f"k={k} kernel={kernel} tile={tile_docs} prune={prune}")
        np.testing.assert_array_equal(scores.view(np.uint32), o_sc.view(np.uint32))
        np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)
        assert np.all(np.diff(scores.astype(np.float64), axis=1) <= 0)  # sortedness
        st = sc.stats()
        if kernel == "block":
            assert st["units"] > 0
            if prune == 0:
                assert st["units_skipped"] == 0, st
            elif k == 1:
                assert st["units_skipped"] > 0, st  # a top-1 threshold prunes most blocks
            if prune == 2 and k <= 100:
                assert st["units_maxscore"] > 0, st
            if prune == 3 and k <= 100:
                assert st["routed_queries"] > 0 and st["candidate_items"] > 0, st
            if prune < 3:
                assert st["routed_queries"] == 0, st
            if prune < 2:
                assert st["units_maxscore"] == 0, st
    # dense surfaces
    for i in (0, 5, 7, 11, 13):
        q = flat[off[i]:off[i + 1]]
        np.testing.assert_array_equal(sc.get_scores_ids(q).view(np.uint32), coracle.get_scores(host, q).view(np.uint32))
        np.testing.assert_allclose(sc.probabilities_device(q).cpu().numpy(), coracle.get_probabilities(host, params, q),
                                   rtol=0, atol=PROB_TOL)
    # large-k dense path
    ids, scores, probs = sc.retrieve_ids(flat[off[0]:off[2]], off[:3] - off[0], 5000, return_scores=True)
    o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat[off[0]:off[2]], off[:3] - off[0], 5000)
    np.testing.assert_array_equal(ids, o_ids)
    np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)


def test_host_buffer_entry_point_and_pruning_levels():
    """bb25_retrieve_batch_host (the C-ABI call with HOST buffers that INTEGRATION.md binds) and
    bb25_index_set_pruning on one handle: every level returns the oracle's result."""
    pkg = _pkg()
    from bayesian_bm25_b200 import _lib, synthetic
    from oracle import coracle
    n_docs, vocab, k = 90_000, 3000, 50
    csc = synthetic.zipf_csc(n_docs, vocab, 40.0, seed=61, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(alpha=1.9, beta=0.25, base_rate=0.04)
    sc.index_from_csc(csc)
    flat, off = synthetic.zipf_queries(200, vocab, seed=62)
    params = coracle.make_params(1.9, 0.25, 0.04)
    o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
    p = _lib.make_params(1.9, 0.25, 0.04)
    seen = {}
    for level in (0, 1, 2, 3):
        sc.set_pruning(level)
        ids = np.empty((200, k), np.int64); scs = np.empty((200, k), np.float32); prs = np.empty((200, k), np.float64)
        _lib.check(_lib.lib().bb25_retrieve_batch_host(sc._handle, C.byref(p), flat.ctypes.data, off.ctypes.data, 200, k,
                                                       ids.ctypes.data, scs.ctypes.data, prs.ctypes.data))
        np.testing.assert_array_equal(ids, o_ids, err_msg=f"level {level}")
        np.testing.assert_array_equal(scs.view(np.uint32), o_sc.view(np.uint32))
        np.testing.assert_allclose(prs, o_pr, rtol=0, atol=PROB_TOL)
        seen[level] = sc.stats()
    assert seen[0]["units_skipped"] == 0 and seen[0]["units_maxscore"] == 0 and seen[0]["routed_queries"] == 0
    assert seen[1]["units_skipped"] > 0 and seen[1]["units_maxscore"] == 0
    assert seen[2]["units_maxscore"] > 0 and seen[2]["routed_queries"] == 0
    assert seen[3]["routed_queries"] > 0 and seen[3]["units"] < seen[2]["units"]
    # scores output may be omitted (NULL)
    ids = np.empty((200, k), np.int64); prs = np.empty((200, k), np.float64)
    _lib.check(_lib.lib().bb25_retrieve_batch_host(sc._handle, C.byref(p), flat.ctypes.data, off.ctypes.data, 200, k,
                                                   ids.ctypes.data, None, prs.ctypes.data))
    np.testing.assert_array_equal(ids, o_ids)


def test_massive_ties_force_threshold_refinement():
    """Every document holds the same single term with the same value: all N keys tie on
    the score, the candidate rows overflow and the 64-bit key threshold must converge."""
    pkg = _pkg()
    from oracle import coracle
    n = 70_000
    csc = {
        "data": torch.full((n + 3,), 0.5, dtype=torch.float32),
        "indices": torch.cat([torch.arange(n, dtype=torch.int32), torch.tensor([1, 5, 9], dtype=torch.int32)]),
        "indptr": torch.tensor([0, n, n + 3], dtype=torch.int64),
        "doc_len": torch.full((n,), 7, dtype=torch.int32),
        "num_docs": n, "avgdl": 7.0,
    }
    csc["data"][n:] = 0.25
    sc = pkg.BayesianBM25Scorer(alpha=1.0, beta=0.2)
    sc.index_from_csc(csc)
    flat = np.array([0, 0, 1, 1, 0], dtype=np.int32)
    off = np.array([0, 1, 3, 5], dtype=np.int64)
    host = _host(csc)
    params = coracle.make_params(1.0, 0.2, None)
    reruns = 0
    for k in (7, 300, 2000):
        ids, scores, probs = sc.retrieve_ids(flat, off, k, return_scores=True)
        o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
        np.testing.assert_array_equal(ids, o_ids)
        np.testing.assert_array_equal(scores, o_sc)
        np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)
        reruns += sc.stats()["rerun_queries"]
    assert reruns > 0  # at least one candidate row overflowed and was repaired


@pytest.mark.parametrize("relaxed", ["1", "0"])
def test_query_order_sums_survive_order_free_traversal(relaxed, monkeypatch):
    """The traversal sums frequent terms' dense rows in registers in S-then-D order; bm25s adds in
    QUERY order.  Here the two orders give different fp32 sums for most documents (query =
    dense, dense, sparse), and the results must still be the query-order ones bit for bit:
    select_kernel re-scores every surviving candidate exactly."""
    pkg = _pkg()
    from oracle import coracle
    monkeypatch.setenv("BB25_RELAXED", relaxed)
    rng = np.random.default_rng(5)
    n = 8192
    sparse_docs = np.arange(3, n, 9, dtype=np.int32)            # df < N/8: no dense row
    v0 = rng.uniform(0.1, 3.0, n).astype(np.float32)            # df = N: dense row
    v2 = rng.uniform(0.1, 3.0, n).astype(np.float32)
    v1 = rng.uniform(0.1, 3.0, sparse_docs.size).astype(np.float32)
    csc = {
        "data": torch.from_numpy(np.concatenate([v0, v1, v2])),
        "indices": torch.from_numpy(np.concatenate([np.arange(n, dtype=np.int32), sparse_docs, np.arange(n, dtype=np.int32)])),
        "indptr": torch.tensor([0, n, n + sparse_docs.size, 2 * n + sparse_docs.size], dtype=torch.int64),
        "doc_len": torch.from_numpy(rng.integers(5, 60, n).astype(np.int32)),
        "num_docs": n, "avgdl": 30.0,
    }
    in_order = (v0[sparse_docs] + v2[sparse_docs]) + v1         # query [0, 2, 1]
    s_first = (v1 + v0[sparse_docs]) + v2[sparse_docs]          # what an S-then-D sum gives
    assert np.count_nonzero(in_order != s_first) > 50           # the orders really differ
    sc = pkg.BayesianBM25Scorer(alpha=1.3, beta=0.4, base_rate=0.02)
    sc.index_from_csc(csc)
    host = _host(csc)
    params = coracle.make_params(1.3, 0.4, 0.02)
    flat = np.array([0, 2, 1, 1, 0, 2, 0, 1, 2, 2, 2, 1, 0], dtype=np.int32)
    off = np.array([0, 3, 6, 9, 13], dtype=np.int64)
    for level in (0, 3):
        sc.set_pruning(level)
        for k in (5, 200, 1500):
            ids, scores, probs = sc.retrieve_ids(flat, off, k, return_scores=True)
            o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
            np.testing.assert_array_equal(ids, o_ids, err_msg=f"level {level} k {k}")
            np.testing.assert_array_equal(scores.view(np.uint32), o_sc.view(np.uint32))
            np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)
    # the first query's top documents all hold the sparse term: their scores are the query-order sums
    ids, scores, _ = sc.retrieve_ids(flat[:3], off[:2], 50, return_scores=True)
    lookup = dict(zip(sparse_docs.tolist(), in_order.tolist()))
    for d, s_ in zip(ids[0], scores[0]):
        if int(d) in lookup:
            assert np.float32(lookup[int(d)]).view(np.uint32) == np.float32(s_).view(np.uint32)


def test_invalid_inputs_are_rejected():
    pkg = _pkg()
    from bayesian_bm25_b200 import _lib, synthetic
    csc = synthetic.zipf_csc(3000, 200, 20.0, seed=3, device=torch.device("cuda:0"))
    sc = pkg.BayesianBM25Scorer(alpha=1.0, beta=0.0)
    sc.index_from_csc(csc)
    with pytest.raises(ValueError):
        sc.retrieve_ids(np.array([1], np.int32), np.array([0, 1], np.int64), k=3001)
    with pytest.raises(RuntimeError, match="out of range"):
        sc.retrieve_ids(np.array([1, 200], np.int32), np.array([0, 2], np.int64), k=5)
    with pytest.raises(RuntimeError, match="out of range"):
        sc.get_scores_ids([999])
    bad = dict(csc)
    bad["indices"] = csc["indices"].flip(0).contiguous()
    with pytest.raises(RuntimeError, match="invalid CSC"):
        pkg.BayesianBM25Scorer(alpha=1.0, beta=0.0).index_from_csc(bad)
    before = _lib.lib().bb25_launch_count()
    sc.retrieve_ids(np.array([1, 2], np.int32), np.array([0, 2], np.int64), k=5)
    assert _lib.lib().bb25_launch_count() > before


# ---------------------------------------------------------------------------------
# merge, dense top-k, multi-field, block-max
# ---------------------------------------------------------------------------------
def test_merge_topk_vs_oracle():
    from bayesian_bm25_b200 import sharded
    from oracle import coracle
    rng = np.random.default_rng(0)
    for S, Q, k, N in ((8, 33, 1000, 40000), (2, 5, 7, 100), (4, 3, 4096, 20000), (1, 4, 10, 50)):
        scores = np.round(rng.uniform(0, 2, (Q, N)), 2).astype(np.float32)
        scores[:, rng.integers(0, N, N // 2)] = 0.0
        ids = np.empty((S, Q, k), np.int64); sc = np.empty((S, Q, k), np.float32); pr = np.empty((S, Q, k))
        for s in range(S):
            lo, hi = s * N // S, (s + 1) * N // S
            for q in range(Q):
                i, v = coracle.topk_f32(scores[q, lo:hi], k)
                ids[s, q], sc[s, q], pr[s, q] = i + lo, v, v * 0.25
        want = coracle.merge_topk(ids, sc, pr)
        got = sharded.merge_topk_device(torch.from_numpy(ids).cuda(), torch.from_numpy(sc).cuda(), torch.from_numpy(pr).cuda())
        for g_, w_ in zip(got, want):
            np.testing.assert_array_equal(g_.cpu().numpy(), w_)
        # packed form (what the sharded retriever all-gathers)
        packed = torch.stack([sharded.pack_topk_device(torch.from_numpy(ids[s_]).cuda(), torch.from_numpy(sc[s_]).cuda(),
                                                       torch.from_numpy(pr[s_]).cuda()) for s_ in range(S)])
        for g_, w_ in zip(sharded.merge_packed_device(packed), want):
            np.testing.assert_array_equal(g_.cpu().numpy(), w_)


def test_topk_f64_vs_oracle():
    from bayesian_bm25_b200 import _lib
    from oracle import coracle
    rng = np.random.default_rng(1)
    cases = [rng.uniform(0, 1, 100_000), np.full(50_000, 7.2e-15), np.round(rng.uniform(0, 1, 200_000), 2),
             np.concatenate([np.zeros(999), [0.5]]), rng.uniform(0, 1e-300, 5000)]
    for vals in cases:
        for k in (1, 10, 1000, min(8192, len(vals))):
            d = torch.from_numpy(vals).cuda()
            ids = torch.empty(k, dtype=torch.int64, device="cuda")
            out = torch.empty(k, dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib().bb25_topk_f64(0, d.data_ptr(), d.numel(), k, ids.data_ptr(), out.data_ptr(), None))
            w_ids, w_vals = coracle.topk_f64(vals, k)
            np.testing.assert_array_equal(ids.cpu().numpy(), w_ids)
            np.testing.assert_array_equal(out.cpu().numpy(), w_vals)


def test_multifield_vs_reference(golden_mf):
    pkg = _pkg()
    from oracle import coracle
    arrays, meta = golden_mf
    for i, c in enumerate(meta["cases"]):
        mf = pkg.MultiFieldScorer(["title", "body"], field_weights=c["field_weights"], alpha=c["alpha"],
                                  base_rate=c["base_rate"], method="lucene")
        mf.index(meta["docs"], show_progress=False)
        for f in mf.fields:
            assert mf._scorers[f].transform.alpha == c["field_alpha"][f]
            assert mf._scorers[f].transform.beta == c["field_beta"][f]
        want = arrays[f"mf{i}_probs"]
        for qi, q in enumerate(meta["queries"]):
            got = mf.get_probabilities(q)
            np.testing.assert_allclose(got, want[qi], rtol=1e-9, atol=1e-18)
            ids, vals = mf.retrieve(q, k=25)
            w_ids, w_vals = coracle.topk_f64(got, 25)
            np.testing.assert_array_equal(ids, w_ids)
            np.testing.assert_array_equal(vals, w_vals)
    # 1 field == the single scorer (tests/test_multi_field.py:106-129)
    docs = [{"body": d["body"]} for d in meta["docs"]]
    mf = pkg.MultiFieldScorer(["body"], method="lucene")
    mf.index(docs, show_progress=False)
    single = pkg.BayesianBM25Scorer(method="lucene")
    single.index([d["body"] for d in docs], show_progress=False)
    q = meta["queries"][0]
    fused, alone = mf.get_probabilities(q), single.get_probabilities(q)
    nzm = alone > 0
    np.testing.assert_allclose(fused[nzm], alone[nzm], atol=1e-6)


def test_hybrid_fusion_vs_oracle():
    """BASELINE configs[3] semantics: BM25 posterior + cosine_to_probability through weighted /
    unweighted log_odds_conjunction, fused on the device, against the oracle's composition of
    the same reference formulas; top-100 with (value desc, id asc) ties."""
    pkg = _pkg()
    from bayesian_bm25_b200 import hybrid, synthetic
    from oracle import coracle
    n_docs, vocab = 60_000, 2000
    csc = synthetic.zipf_csc(n_docs, vocab, 40.0, seed=31, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(alpha=1.8, beta=0.6, base_rate=0.02)
    sc.index_from_csc(csc)
    params = coracle.make_params(1.8, 0.6, 0.02)
    rng = np.random.default_rng(44)
    for q in ([3, 17, 250], [0, 1], [], [1999, 5, 5]):
        cos = np.clip(rng.normal(0.2, 0.15, n_docs), -1, 1).astype(np.float32)
        p_b = coracle.get_probabilities(host, params, q)
        p_v = coracle.cosine_to_probability(cos.astype(np.float64))
        stacked = np.stack([p_b, p_v], axis=-1)
        for weights, alpha in (((0.6, 0.4), None), ((0.6, 0.4), 0.5), (None, None), (None, "auto"), ((0.5, 0.5), 1.0)):
            want = coracle.log_odds_conjunction(stacked, alpha=alpha, weights=weights)
            got = hybrid.hybrid_probabilities_device(sc, q, torch.from_numpy(cos).cuda(), weights, alpha).cpu().numpy()
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-18)
            # the composed (unfused) public functions give the same numbers
            comp = pkg.log_odds_conjunction(np.stack([sc.probabilities_device(q).cpu().numpy(),
                                                      pkg.cosine_to_probability(cos)], axis=-1), alpha=alpha, weights=weights)
            np.testing.assert_array_equal(got, comp)
            ids, vals = hybrid.hybrid_retrieve(sc, q, torch.from_numpy(cos).cuda(), k=100, weights=weights, alpha=alpha)
            w_ids, w_vals = coracle.topk_f64(got, 100)
            np.testing.assert_array_equal(ids, w_ids)
            np.testing.assert_array_equal(vals, w_vals)


def test_blockmax_vs_reference(golden_mf):
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    arrays, _ = golden_mf
    sm = arrays["bmw_matrix"]
    for bs in (1, 7, 128, 1000, 4096):
        b = pkg.BlockMaxIndex(block_size=bs)
        b.build(sm)
        np.testing.assert_array_equal(b._block_maxes, arrays[f"bmw_bs{bs}"])
        assert b.n_blocks == -(-sm.shape[1] // bs)
    b = pkg.BlockMaxIndex(block_size=128)
    b.build(sm)
    t = pkg.BayesianProbabilityTransform(1.3, 1.7, base_rate=0.03)
    got = np.array([[b.bayesian_block_upper_bound(ti, bi, t) for bi in range(b.n_blocks)] for ti in range(2)])
    np.testing.assert_allclose(got, arrays["bmw_bayes_bs128"][:2], rtol=0, atol=PROB_TOL)
    with pytest.raises(ValueError, match="2D"):
        b.build(np.zeros(5))
    # CSC form + pruning safety: bound >= every document's probability in the block (tests/test_bmw.py:82-113)
    csc = synthetic.zipf_csc(20_000, 300, 30.0, seed=4, device=torch.device("cuda:0"))
    sc = pkg.BayesianBM25Scorer(alpha=1.5, beta=0.5, base_rate=0.05)
    sc.index_from_csc(csc)
    terms = [0, 3, 50, 299]
    b.build_from_scorer(sc, terms)
    want = coracle.blockmax_csc(_host(csc), terms, 128)
    np.testing.assert_array_equal(b._block_maxes.astype(np.float32), want)
    probs = sc.probabilities_device([50]).cpu().numpy()
    bounds = sc.transform.wand_upper_bound(b._block_maxes[2])
    for blk in range(b.n_blocks):
        assert probs[blk * 128:(blk + 1) * 128].max() <= bounds[blk] + 1e-15


# ---------------------------------------------------------------------------------
# BASELINE full size (8.8 M docs): direct oracle comparison on a few queries plus
# size-independent properties (shard union == unsharded, sortedness, dense == top-k)
# ---------------------------------------------------------------------------------
def test_full_size_config2_properties():
    pkg = _pkg()
    from bayesian_bm25_b200 import index_build, sharded, synthetic
    from oracle import coracle
    n_docs, vocab, k = 8_800_000, 30_000, 1000
    dev = torch.device("cuda:0")
    csc = synthetic.zipf_csc(n_docs, vocab, 56.0, seed=42, device=dev)
    nnz = csc["data"].numel()
    assert 3.9e8 < nnz < 4.3e8  # ~408 M postings
    sc = pkg.BayesianBM25Scorer(method="lucene", alpha=2.0, beta=0.2, base_rate=0.045)
    sc.index_from_csc(csc)
    q_terms, q_off = synthetic.zipf_queries(253, vocab, seed=43)
    # plus an all-head-term query, a duplicate-term query and a rare-term query
    extra = [np.array([0, 1, 2, 3], np.int32), np.array([5, 5, 17, 5], np.int32), np.array([29999, 29998], np.int32)]
    q_terms = np.concatenate([q_terms] + extra).astype(np.int32)
    q_off = np.concatenate([q_off, q_off[-1] + np.cumsum([len(e) for e in extra])]).astype(np.int64)
    # the oracle on 64 queries (the first 61 and the three special ones), CPU, seconds
    host = _host(csc)
    params = coracle.make_params(2.0, 0.2, 0.045)
    sel = list(range(61)) + [253, 254, 255]
    o_terms = np.concatenate([q_terms[q_off[i]:q_off[i + 1]] for i in sel]).astype(np.int32)
    o_off = np.concatenate([[0], np.cumsum([q_off[i + 1] - q_off[i] for i in sel])]).astype(np.int64)
    o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, o_terms, o_off, k)
    by_level = {}
    for level in (0, 3):  # exhaustive (the headline mode) and the default pruning level
        sc.set_pruning(level)
        ids, scores, probs = sc.retrieve_ids(q_terms, q_off, k, return_scores=True)
        by_level[level] = (ids, scores, probs)
        assert np.all(np.diff(scores.astype(np.float64), axis=1) <= 0)
        ties = np.diff(scores.astype(np.float64), axis=1) == 0
        assert np.all(np.diff(ids, axis=1)[ties] > 0)  # ties by ascending doc id
        assert np.all((probs >= 0) & (probs <= 1))
        np.testing.assert_array_equal(ids[sel], o_ids)
        np.testing.assert_array_equal(scores[sel].view(np.uint32), o_sc.view(np.uint32))
        np.testing.assert_allclose(probs[sel], o_pr, rtol=0, atol=PROB_TOL)
        assert sc.stats()["host_syncs"] <= 2
    for a_, b_ in zip(by_level[0], by_level[3]):
        np.testing.assert_array_equal(a_, b_)
    ids, scores, probs = by_level[3]
    # dense scores agree with the retrieved ones
    q0 = q_terms[q_off[0]:q_off[1]]
    dense = sc.get_scores_ids(q0)
    np.testing.assert_array_equal(dense[ids[0]], scores[0])
    del host
    nq = 0
    # shard union == unsharded (4 shards scored one after the other on this GPU, merged on device)
    parts = []
    dq, do = torch.from_numpy(q_terms).to(dev), torch.from_numpy(q_off).to(dev)
    for lo, hi in index_build.shard_bounds(n_docs, 4):
        s = pkg.BayesianBM25Scorer(method="lucene", alpha=2.0, beta=0.2, base_rate=0.045)
        s.index_from_csc(index_build.shard_csc(csc, lo, hi))
        parts.append(s.retrieve_ids_device(dq, do, k))
        del s
    m_ids, m_sc, m_pr = sharded.merge_topk_device(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                                  torch.stack([p[2] for p in parts]))
    np.testing.assert_array_equal(m_ids.cpu().numpy(), ids)
    np.testing.assert_array_equal(m_sc.cpu().numpy(), scores)
    np.testing.assert_array_equal(m_pr.cpu().numpy(), probs)


@pytest.mark.parametrize("lookup_div", ["64", "0", "100000"])
def test_essential_posting_evaluation_is_exact(lookup_div, monkeypatch):
    """Pruning levels >= 2 evaluate (block, query) units through their essential postings (MaxScore split per
    1024-document block, non-essential values read from the hot / lookup value rows).  With the lookup rows
    at their default extent, absent (BB25_LOOKUP_DIV=0) and covering every term, and with the evaluation
    switched off, the results are the oracle's bit for bit -- queries with up to 30 terms, duplicate terms,
    rare-only and frequent-only queries included."""
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    monkeypatch.setenv("BB25_LOOKUP_DIV", lookup_div)
    n_docs, vocab, k = 150_000, 6000, 100
    csc = synthetic.zipf_csc(n_docs, vocab, 45.0, seed=71, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(alpha=2.0, beta=0.2, base_rate=0.04)
    sc.index_from_csc(csc)
    terms, off = synthetic.zipf_queries(240, vocab, seed=72)
    qs = [terms[off[i]:off[i + 1]] for i in range(240)]
    rng = np.random.default_rng(73)
    qs += [np.asarray(x, dtype=np.int32) for x in (
        [5999], [5998, 5997, 5996], [0, 1, 2], [0, 0, 3000, 3000], [7, 4000, 7, 4000, 7],
        rng.integers(0, vocab, 10), rng.integers(0, vocab, 20), rng.integers(0, vocab, 30),
        rng.integers(0, 50, 9), rng.integers(2000, vocab, 8), [], [1, 5000])]
    flat = np.concatenate([q for q in qs]).astype(np.int32)
    qoff = np.zeros(len(qs) + 1, dtype=np.int64)
    qoff[1:] = np.cumsum([len(q) for q in qs])
    params = coracle.make_params(2.0, 0.2, 0.04)
    o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, qoff, k)
    seen = {}
    for level, sparse in ((0, "1"), (2, "1"), (2, "0"), (3, "1"), (3, "0")):
        monkeypatch.setenv("BB25_SPARSE", sparse)
        sc.set_pruning(level)
        ids, scores, probs = sc.retrieve_ids(flat, qoff, k, return_scores=True)
        np.testing.assert_array_equal(ids, o_ids, err_msg=f"level {level} sparse {sparse}")
        np.testing.assert_array_equal(scores.view(np.uint32), o_sc.view(np.uint32))
        np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)
        seen[(level, sparse)] = sc.stats()
    assert seen[(0, "1")]["units_sparse"] == 0 and seen[(2, "0")]["units_sparse"] == 0
    assert seen[(2, "1")]["units_sparse"] > 0


def test_pipelined_host_calls_and_concurrent_callers():
    """retrieve(list[list[str]]) on a large batch runs chunked host calls with two in flight (the library's two
    staging slots: one chunk's results travel to the host while the next chunk is traversed), and two threads
    may call retrieve_ids on one scorer at the same time.  Everything equals the single-call result and the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    n_docs, vocab, k = 60_000, 2500, 64
    csc = synthetic.zipf_csc(n_docs, vocab, 40.0, seed=81, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(alpha=1.8, beta=0.3, base_rate=0.05)
    sc.index_from_csc(csc)
    sc.set_vocabulary([f"w{i}" for i in range(vocab)])
    flat, off = synthetic.zipf_queries(3000, vocab, seed=82)
    params = coracle.make_params(1.8, 0.3, 0.05)
    o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
    ids, scs, prs = sc.retrieve_ids(flat, off, k, return_scores=True)
    np.testing.assert_array_equal(ids, o_ids)
    np.testing.assert_array_equal(scs.view(np.uint32), o_sc.view(np.uint32))
    np.testing.assert_allclose(prs, o_pr, rtol=0, atol=PROB_TOL)
    toks = [[f"w{t}" for t in flat[off[i]:off[i + 1]]] + (["not-a-word"] if i % 7 == 0 else []) for i in range(3000)]
    s_ids, s_prs = sc.retrieve(toks, k=k)  # 3 chunks: 1024 + 2 x 988
    np.testing.assert_array_equal(s_ids, o_ids)
    np.testing.assert_allclose(s_prs, o_pr, rtol=0, atol=PROB_TOL)
    halves = [(0, 1500), (1500, 3000)]
    def call(rng_):
        a, b = rng_
        out = []
        for _ in range(3):
            out.append(sc.retrieve_ids(flat[off[a]:off[b]], off[a:b + 1] - off[a], k, return_scores=True))
        return out
    with ThreadPoolExecutor(max_workers=2) as pool:
        res = list(pool.map(call, halves))
    for (a, b), outs in zip(halves, res):
        for i_, s_, p_ in outs:
            np.testing.assert_array_equal(i_, o_ids[a:b])
            np.testing.assert_array_equal(s_.view(np.uint32), o_sc[a:b].view(np.uint32))
            np.testing.assert_allclose(p_, o_pr[a:b], rtol=0, atol=PROB_TOL)


def test_fp16_bound_rows_never_hide_a_winner():
    """The pruned passes read the frequent terms' rows as fp16 UPPER bounds (rounded up, by up to one fp16 ulp =
    2^-10 relative).  Documents whose values are exact in fp16 then compete, in select_kernel's pre-filter, with
    keys that overstate their documents by almost 0.1 %: 600 documents score 1.078125 (exact in fp16), the other
    7 592 score 1.07801 but look like 1.07904 through the fp16 rows.  The true top-1000 is the 600 plus the first
    400 of the others by id -- at every pruning level."""
    pkg = _pkg()
    from oracle import coracle
    n, n_x = 8192, 600
    va = np.full(n, 1.00001, dtype=np.float32)
    vb = np.full(n, 0.0780, dtype=np.float32)
    x_docs = np.arange(n - n_x, n)  # the exactly representable ones come last
    va[x_docs], vb[x_docs] = 1.0, 0.078125
    csc = {
        "data": torch.from_numpy(np.concatenate([va, vb])),
        "indices": torch.cat([torch.arange(n, dtype=torch.int32), torch.arange(n, dtype=torch.int32)]),
        "indptr": torch.tensor([0, n, 2 * n], dtype=torch.int64),
        "doc_len": torch.full((n,), 9, dtype=torch.int32),
        "num_docs": n, "avgdl": 9.0,
    }
    sc = pkg.BayesianBM25Scorer(alpha=1.5, beta=0.3, base_rate=0.05)
    sc.index_from_csc(csc)
    flat = np.array([0, 1, 1, 0, 0, 1, 0], dtype=np.int32)
    off = np.array([0, 2, 4, 7], dtype=np.int64)
    host = _host(csc)
    params = coracle.make_params(1.5, 0.3, 0.05)
    for k in (1000, 700, 601):
        o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, flat, off, k)
        assert set(x_docs.tolist()) <= set(o_ids[0].tolist())
        for level in (0, 1, 2, 3):
            sc.set_pruning(level)
            ids, scores, probs = sc.retrieve_ids(flat, off, k, return_scores=True)
            np.testing.assert_array_equal(ids, o_ids, err_msg=f"k {k} level {level}")
            np.testing.assert_array_equal(scores.view(np.uint32), o_sc.view(np.uint32))
            np.testing.assert_allclose(probs, o_pr, rtol=0, atol=PROB_TOL)
