"""GPU parity of the query-time consumers added around the hot path against vectors
produced by the reference itself (tests/golden/make_golden.py:extras): balanced_log_odds_fusion,
AttentionLogOddsWeights inference (__call__ / compute_upper_bounds / prune), retrieve(explain=True)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-9
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gx():
    return np.load(os.path.join(GOLDEN, "extras.npz")), json.load(open(os.path.join(GOLDEN, "extras.json")))


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)


def test_balanced_log_odds_fusion_vs_reference(gx):
    import bayesian_bm25_b200 as pkg
    g, _ = gx
    sp, de = g["bal_sparse"], g["bal_dense"]
    for w in (0.5, 0.3, 0.0, 1.0):
        np.testing.assert_allclose(pkg.balanced_log_odds_fusion(sp, de, weight=w), g[f"bal_w{w}"], rtol=0, atol=TOL)
    # a zero-variance signal contributes nothing (fusion.py:336-343)
    np.testing.assert_allclose(pkg.balanced_log_odds_fusion(np.full(50, 0.3), de[:50], weight=0.4), g["bal_const_sparse"], atol=TOL)
    np.testing.assert_allclose(pkg.balanced_log_odds_fusion(sp[:50], np.full(50, 0.2), weight=0.4), g["bal_const_dense"], atol=TOL)
    assert isinstance(pkg.balanced_log_odds_fusion(0.7, 0.2), float)


def test_attention_log_odds_weights_inference_vs_reference(gx):
    import bayesian_bm25_b200 as pkg
    g, meta = gx
    for ci, c in enumerate(meta["attention"]):
        a = pkg.AttentionLogOddsWeights(c["n_signals"], c["n_query_features"], alpha=c["alpha"], normalize=c["normalize"],
                                        seed=c["seed"], base_rate=c["base_rate"])
        pre = f"att{ci}_"
        np.testing.assert_array_equal(a.weights_matrix, g[pre + "W_init"])  # same initialisation as the reference
        a.set_parameters(g[pre + "W"], g[pre + "b"], g[pre + "W"] * 0.9, g[pre + "b"] * 1.1)
        P, qf1, qfm, UB = g[pre + "P"], g[pre + "qf1"], g[pre + "qfm"], g[pre + "UB"]
        np.testing.assert_allclose(a._compute_weights(qf1), g[pre + "w1"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(a._compute_weights(qfm), g[pre + "wm"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(a._compute_weights(qfm, use_averaged=True), g[pre + "wm_avg"], rtol=0, atol=1e-12)
        single = a(P[5], qf1)
        assert isinstance(single, float)
        np.testing.assert_allclose(single, g[pre + "call_single"][0], rtol=0, atol=TOL)
        np.testing.assert_allclose(a(P, qf1), g[pre + "call_batch_q1"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a(P, qfm), g[pre + "call_batch_qm"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a(P, qfm, use_averaged=True), g[pre + "call_batch_avg"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a.compute_upper_bounds(UB, qf1), g[pre + "ub_q1"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a.compute_upper_bounds(UB, qfm), g[pre + "ub_qm"], rtol=0, atol=TOL)
        idx, fused = a.prune(P, qfm, c["threshold"], upper_bound_probs=UB)
        np.testing.assert_array_equal(idx, g[pre + "prune_idx"])
        np.testing.assert_allclose(fused, g[pre + "prune_fused"], rtol=0, atol=TOL)
        idx2, fused2 = a.prune(P, qf1, c["threshold"])
        np.testing.assert_array_equal(idx2, g[pre + "prune2_idx"])
        np.testing.assert_allclose(fused2, g[pre + "prune2_fused"], rtol=0, atol=TOL)
        # safety of the bound (Theorem 8.7.1): upper bounds dominate the fused probabilities
        assert np.all(a.compute_upper_bounds(UB, qfm) >= a(P, qfm) - 1e-12) or c["normalize"]
    with pytest.raises(ValueError):
        pkg.AttentionLogOddsWeights(0, 3)
    with pytest.raises(ValueError):
        pkg.AttentionLogOddsWeights(2, 3, base_rate=1.5)
    with pytest.raises(NotImplementedError):
        pkg.AttentionLogOddsWeights(2, 3).fit()


def test_trace_bm25_kernel_vs_reference(gx):
    import ctypes as C
    from bayesian_bm25_b200 import _lib
    g, meta = gx
    dev = torch.device("cuda:0")
    s, tf, r = (torch.from_numpy(g[k]).to(dev) for k in ("tr_s", "tr_tf", "tr_r"))
    for ti, (a_, b_, br_) in enumerate(meta["trace_params"]):
        p = _lib.make_params(a_, b_, br_)
        out = torch.empty((s.numel(), 7), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bb25_trace_bm25(0, C.byref(p), s.data_ptr(), tf.data_ptr(), r.data_ptr(), s.numel(),
                                              out.data_ptr(), None))
        np.testing.assert_allclose(out.cpu().numpy(), g[f"tr_out{ti}"][:, :7], rtol=0, atol=1e-12)


def test_retrieve_explain_vs_reference(gx):
    import bayesian_bm25_b200 as pkg
    from bayesian_bm25_b200 import synthetic
    g, meta = gx
    corpus, _ = synthetic.scalability_corpus(300, 200, 30, np.random.default_rng(meta["explain"]["corpus_seed"]))
    queries = meta["explain"]["queries"]
    sc = pkg.BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", base_rate="auto")
    sc.index(corpus, show_progress=False)
    assert sc.transform.alpha == meta["explain"]["alpha"] and sc.transform.beta == meta["explain"]["beta"]
    res = sc.retrieve(queries, k=10, explain=True)
    assert isinstance(res, pkg.RetrievalResult)
    np.testing.assert_array_equal(res.doc_ids, g["ex_ids"])
    np.testing.assert_allclose(res.probabilities, g["ex_probs"], rtol=0, atol=TOL)
    want = g["ex_traces"]
    assert len(res.explanations) == len(queries)
    for qi, row in enumerate(res.explanations):
        assert len(row) == 10
        for ri, t in enumerate(row):
            if np.isnan(want[qi, ri, 0]):
                assert t is None
                continue
            assert isinstance(t, pkg.BM25SignalTrace)
            got = [t.raw_score, t.tf, t.doc_len_ratio, t.likelihood, t.tf_prior, t.norm_prior, t.composite_prior,
                   t.logit_likelihood, t.logit_prior, t.logit_base_rate, t.posterior]
            np.testing.assert_allclose(got, want[qi, ri], rtol=0, atol=1e-9)
            assert t.alpha == sc.transform.alpha and t.base_rate == sc.transform.base_rate
            # trace == probability (tests/test_scorer.py: trace posterior equals the returned probability)
            assert abs(t.posterior - res.probabilities[qi, ri]) < 1e-9
    # explain=False keeps the tuple contract
    ids, probs = sc.retrieve(queries, k=10)
    np.testing.assert_array_equal(ids, res.doc_ids)


def test_cosine_gemm_tcgen05_vs_torch_fp32():
    """bb25_cosine_gemm (TMA + tcgen05.mma, fp32 accumulation in tensor memory) against a plain PyTorch fp32
    matmul of the same bf16 inputs; tolerance: fp32 summation order over K terms."""
    from bayesian_bm25_b200 import dense
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for n, k, nq in ((5000, 128, 17), (128, 64, 1), (70001, 768, 256), (9000, 768, 300), (257, 192, 16)):
        q = torch.nn.functional.normalize(torch.randn((nq, k), device="cuda", generator=g), dim=1).to(torch.bfloat16)
        c = torch.nn.functional.normalize(torch.randn((n, k), device="cuda", generator=g), dim=1).to(torch.bfloat16)
        got = dense.cosine_scores(q, c)
        assert got.shape == (nq, (n + 3) // 4 * 4)
        want = q.float() @ c.float().T
        torch.testing.assert_close(got[:, :n], want, rtol=0, atol=2e-5)


def test_hybrid_from_embeddings_vs_oracle():
    """Embeddings -> tensor-core cosines -> fused-rank batch retrieval, against the oracle's conjunction of the
    oracle's BM25 posterior and cosine_to_probability of the SAME cosines."""
    import bayesian_bm25_b200 as pkg
    from bayesian_bm25_b200 import dense, hybrid, synthetic
    from oracle import coracle
    n_docs, vocab, kdim = 30_000, 1500, 128
    csc = synthetic.zipf_csc(n_docs, vocab, 40.0, seed=3, device=torch.device("cuda:0"))
    host = {k_: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k_, v in csc.items()}
    sc = pkg.BayesianBM25Scorer(alpha=1.8, beta=0.6, base_rate=0.02)
    sc.index_from_csc(csc)
    params = coracle.make_params(1.8, 0.6, 0.02)
    terms, off = synthetic.zipf_queries(20, vocab, seed=8)
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    qe = torch.nn.functional.normalize(torch.randn((20, kdim), device="cuda", generator=g), dim=1)
    ce = torch.nn.functional.normalize(torch.randn((n_docs, kdim), device="cuda", generator=g), dim=1)
    cos = dense.cosine_scores(qe, ce)[:, :n_docs].cpu().numpy()
    ids, vals = hybrid.hybrid_retrieve_batch_embeddings(sc, terms, off, qe, ce, k=50, weights=(0.6, 0.4), alpha=0.5, sub_batch=8)
    for q in range(20):
        p_b = coracle.get_probabilities(host, params, terms[off[q]:off[q + 1]])
        want = coracle.log_odds_conjunction(np.stack([p_b, coracle.cosine_to_probability(cos[q].astype(np.float64))], axis=-1),
                                            alpha=0.5, weights=(0.6, 0.4))
        w_ids, w_vals = coracle.topk_f64(want, 50)
        np.testing.assert_allclose(vals[q], w_vals, rtol=0, atol=TOL)
        assert np.array_equal(ids[q], w_ids) or np.max(np.abs(vals[q] - w_vals)) < 1e-12
