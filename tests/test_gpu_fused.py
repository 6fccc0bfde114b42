"""GPU parity of the batched fused-rank retrieval (bb25_retrieve_fused_batch: MultiFieldScorer.retrieve_batch,
hybrid_retrieve_batch; BASELINE configs 4 and 5) against the CPU oracle's composition of the reference
formulas -- per-field get_probabilities (scorer.py:564-590) -> log_odds_conjunction (fusion.py:172-280) ->
top-k by (fused desc, doc id asc) -- and against the library's own dense single-query path.
Bars: ids identical, fused probabilities within 1e-9 (north star: 1e-6)."""
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


def _two_field(n_docs, vocab, seed, title_len=8.0, body_len=40.0, base_rate=0.03, alpha="auto", weights=None,
               fields=("title", "body")):
    """MultiFieldScorer over synthetic Zipf fields + the host CSCs and per-field oracle parameters."""
    pkg = _pkg()
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    dev = torch.device("cuda:0")
    lens = {"title": title_len, "body": body_len, "anchor": 5.0}
    cscs, pseudo = {}, {}
    for i, f in enumerate(fields):
        cscs[f] = synthetic.zipf_csc(n_docs, vocab, lens[f], seed + 17 * i, dev, min_len=2)
        pseudo[f] = synthetic.zipf_pseudo_queries(n_docs, vocab, lens[f], seed + 17 * i, min_len=2)
    mf = pkg.MultiFieldScorer(list(fields), field_weights=weights, alpha=alpha, base_rate=base_rate, method="lucene")
    mf.index_from_csc(cscs, pseudo_queries=pseudo)
    hosts = {f: _host(cscs[f]) for f in fields}
    params = {}
    for f in fields:
        t = mf._scorers[f].transform
        params[f] = coracle.make_params(t.alpha, t.beta, t.base_rate)
    return mf, hosts, params


def _oracle_fused(mf, hosts, params, per_field_terms):
    """Fused probability of every document, by the oracle (reference formulas, CPU)."""
    from oracle import coracle
    cols = [coracle.get_probabilities(hosts[f], params[f], np.asarray(t, dtype=np.int32))
            for f, t in zip(mf.fields, per_field_terms)]
    w = np.array([mf.field_weights[f] for f in mf.fields])
    # multi_field.py:170: the scorer resolves alpha itself, None / "auto" -> 0.5, before the conjunction
    alpha = 0.5 if mf._alpha is None or mf._alpha == "auto" else float(mf._alpha)
    return coracle.log_odds_conjunction(np.stack(cols, axis=-1), alpha=alpha, weights=w)


def _check_rows(ids, vals, fused_rows, k):
    """ids/vals [Q,k] against per-query dense fused vectors: canonical (value desc, id asc) top-k."""
    from oracle import coracle
    for q, fused in enumerate(fused_rows):
        w_ids, w_vals = coracle.topk_f64(fused, k)
        np.testing.assert_array_equal(ids[q], w_ids, err_msg=f"query {q}")
        np.testing.assert_array_equal(vals[q], w_vals, err_msg=f"query {q}")


def _queries(n, vocab, seed, extra=()):
    from bayesian_bm25_b200 import synthetic
    terms, off = synthetic.zipf_queries(n, vocab, seed)
    qs = [terms[off[i]:off[i + 1]] for i in range(n)]
    return qs + [np.asarray(e, dtype=np.int32) for e in extra]


def _flat(qs):
    from bayesian_bm25_b200 import fused
    return fused._flat_queries(qs)


def test_multifield_batch_vs_dense_path_and_oracle():
    """60 k documents, two fields, every pruning level, k in {1, 10, 100}: the batch equals the library's
    dense single-query path bit for bit and the oracle within tolerance; edge-case queries included."""
    from oracle import coracle
    vocab = 2000
    mf, hosts, params = _two_field(60_000, vocab, seed=101)
    extra = [[], [5, 5, 5, 17], [1999], [1998, 1997], list(range(40)), [0, 1, 2, 3], [0, 0, 1]]
    qs = _queries(48, vocab, 7, extra)
    fq = [_flat(qs), _flat(qs)]  # the fields share the synthetic vocabulary
    dense = [mf._fused_device([q, q]).cpu().numpy() for q in qs]
    # the dense path itself against the oracle (reference formulas)
    for q, d in zip(qs[:12] + qs[-7:], dense[:12] + dense[-7:]):
        np.testing.assert_allclose(d, _oracle_fused(mf, hosts, params, [q, q]), rtol=1e-9, atol=1e-18)
    for level in (0, 1, 2, 3):
        mf.set_pruning(level)
        for k in (1, 10, 100):
            ids, vals = mf.retrieve_ids_batch(fq, k)
            _check_rows(ids, vals, dense, k)
            st = mf.stats()
            assert st["host_syncs"] >= 1
            if level == 0:
                assert st["units_skipped"] == 0
    # the long query (> 32 terms) and the empty one went through the dense guaranteed path
    assert mf.stats()["fallback_queries"] >= 2
    # string API: same rows
    toks = [[f"t{t}" for t in q] for q in qs]
    for f in mf.fields:
        mf._scorers[f].set_vocabulary([f"t{i}" for i in range(vocab)])
    ids_s, vals_s = mf.retrieve_batch(toks, k=10)
    ids, vals = mf.retrieve_ids_batch(fq, 10)
    np.testing.assert_array_equal(ids_s, ids)
    np.testing.assert_array_equal(vals_s, vals)
    r_ids, r_vals = mf.retrieve(toks[3], k=10)
    np.testing.assert_array_equal(r_ids, ids[3])
    np.testing.assert_array_equal(r_vals, vals[3])
    assert coracle is not None


@pytest.mark.parametrize("fields,weights,alpha", [
    (("body",), None, "auto"),
    (("title", "body", "anchor"), {"title": 0.5, "body": 0.3, "anchor": 0.2}, 0.0),
    (("title", "body"), {"title": 0.7, "body": 0.3}, 1.5),
    (("title", "body"), None, None),
])
def test_multifield_batch_field_counts_and_weights(fields, weights, alpha):
    vocab = 800
    mf, hosts, params = _two_field(30_000, vocab, seed=5, fields=fields, weights=weights, alpha=alpha, base_rate=None)
    qs = _queries(40, vocab, 11, [[], [799, 798]])
    fq = [_flat(qs) for _ in fields]
    dense = [mf._fused_device([q] * len(fields)).cpu().numpy() for q in qs]
    np.testing.assert_allclose(dense[0], _oracle_fused(mf, hosts, params, [qs[0]] * len(fields)), rtol=1e-9, atol=1e-18)
    for level in (0, 1):
        mf.set_pruning(level)
        ids, vals = mf.retrieve_ids_batch(fq, 20)
        _check_rows(ids, vals, dense, 20)


def test_fused_batch_different_queries_per_field_and_tiny_corpus():
    """Per-field vocabularies differ in practice: every field gets its own term lists.  Tiny corpus:
    k close to the number of documents, most queries take the guaranteed path."""
    vocab = 60
    mf, hosts, params = _two_field(300, vocab, seed=9)
    rng = np.random.default_rng(3)
    q_title = [rng.integers(0, vocab, rng.integers(0, 5)).astype(np.int32) for _ in range(30)]
    q_body = [rng.integers(0, vocab, rng.integers(0, 6)).astype(np.int32) for _ in range(30)]
    dense = [mf._fused_device([a, b]).cpu().numpy() for a, b in zip(q_title, q_body)]
    for k in (5, 250, 300):
        ids, vals = mf.retrieve_ids_batch([_flat(q_title), _flat(q_body)], k)
        _check_rows(ids, vals, dense, k)


def test_hybrid_batch_vs_dense_path_and_oracle():
    """BASELINE configs[3]: BM25 posterior + cosine_to_probability, weighted / unweighted conjunction,
    top-100, a batch of queries each with its own cosine row; N not a multiple of 4 (padded rows)."""
    pkg = _pkg()
    from bayesian_bm25_b200 import hybrid, synthetic
    from oracle import coracle
    n_docs, vocab = 60_001, 2000
    csc = synthetic.zipf_csc(n_docs, vocab, 40.0, seed=31, device=torch.device("cuda:0"))
    host = _host(csc)
    sc = pkg.BayesianBM25Scorer(alpha=1.8, beta=0.6, base_rate=0.02)
    sc.index_from_csc(csc)
    params = coracle.make_params(1.8, 0.6, 0.02)
    qs = _queries(20, vocab, 5, [[], [1999, 5, 5], [0, 1]])
    flat, off = _flat(qs)
    rng = np.random.default_rng(44)
    cos = np.clip(rng.normal(0.2, 0.15, (len(qs), n_docs)), -1, 1).astype(np.float32)
    cos[0, :50] = 1.0
    cos[1, 100:150] = -1.0
    d_cos = torch.from_numpy(cos).cuda()
    for weights, alpha in (((0.6, 0.4), None), ((0.6, 0.4), 0.5), (None, None), (None, "auto"), ((0.5, 0.5), 1.0)):
        dense = [hybrid.hybrid_probabilities_device(sc, q, d_cos[i], weights, alpha).cpu().numpy() for i, q in enumerate(qs)]
        p_b = coracle.get_probabilities(host, params, qs[2])
        want = coracle.log_odds_conjunction(np.stack([p_b, coracle.cosine_to_probability(cos[2].astype(np.float64))], axis=-1),
                                            alpha=alpha, weights=weights)
        np.testing.assert_allclose(dense[2], want, rtol=1e-9, atol=1e-18)
        for level in (0, 3):
            sc.set_pruning(level)
            for k in (10, 100):
                ids, vals = hybrid.hybrid_retrieve_batch(sc, flat, off, d_cos, k, weights, alpha)
                _check_rows(ids, vals, dense, k)


def test_config5_shape_1m_docs_vs_oracle():
    """>= 1 M documents, two fields, Q >= 256, k in {10, 100}, every pruning level, against
    the oracle's log_odds_conjunction of get_probabilities: ids identical (ties by doc id), fused <= 1e-9."""
    from oracle import coracle
    vocab, n_docs, nq = 30_000, 1_100_000, 256
    mf, hosts, params = _two_field(n_docs, vocab, seed=42, title_len=8.0, body_len=56.0)
    qs = _queries(nq - 2, vocab, 43, [[0, 1, 2, 3, 4], [7, 7, 29999]])
    fq = [_flat(qs), _flat(qs)]
    want = []
    for q in qs:
        fused = _oracle_fused(mf, hosts, params, [q, q])
        want.append(coracle.topk_f64(fused, 100))
    ref = None
    for level in (0, 1, 2, 3):
        mf.set_pruning(level)
        for k in (10, 100):
            ids, vals = mf.retrieve_ids_batch(fq, k)
            for q in range(len(qs)):
                w_ids, w_vals = want[q][0][:k], want[q][1][:k]
                np.testing.assert_allclose(vals[q], w_vals, rtol=0, atol=PROB_TOL, err_msg=f"query {q}")
                if not np.array_equal(ids[q], w_ids):
                    # libm and CUDA exp/log differ by an ulp or two: ids may only differ where the oracle's
                    # own values are within that distance of each other
                    diff = np.nonzero(ids[q] != w_ids)[0]
                    assert np.all(np.abs(w_vals[diff] - vals[q][diff]) < 1e-12), f"query {q}"
                    assert sorted(ids[q]) == sorted(w_ids) or np.abs(w_vals[-1] - vals[q][-1]) < 1e-12
            if k == 100:
                if ref is None:
                    ref = (ids.copy(), vals.copy())
                np.testing.assert_array_equal(ids, ref[0])  # identical at every level
                np.testing.assert_array_equal(vals, ref[1])
        st = mf.stats()
        assert st["host_syncs"] == 1 and st["fallback_queries"] == 0
        if level >= 1:
            assert st["units_skipped"] > 0


def test_config5_full_size_50m_docs():
    """BASELINE configs[4] at full size: 50 M documents, title + body, top-10 with block-max pruning.
    24 queries: the pruned batch equals the exhaustive batch and the library's dense single-query path
    (itself oracle-checked above and on a document slice here)."""
    from bayesian_bm25_b200 import synthetic
    from oracle import coracle
    free, total = torch.cuda.mem_get_info()
    if total < 150e9:
        pytest.skip("needs a 180 GB device")
    vocab, n_docs, nq = 30_000, 50_000_000, 24
    mf, hosts, params = None, None, None
    pkg = _pkg()
    dev = torch.device("cuda:0")
    cscs, pseudo = {}, {}
    for f, ln, seed in (("title", 8.0, 142), ("body", 56.0, 42)):
        cscs[f] = synthetic.zipf_csc(n_docs, vocab, ln, seed, dev, min_len=2)
        pseudo[f] = synthetic.zipf_pseudo_queries(n_docs, vocab, ln, seed, min_len=2)
    mf = pkg.MultiFieldScorer(["title", "body"], alpha="auto", base_rate="auto", method="lucene")
    mf.index_from_csc(cscs, pseudo_queries=pseudo)
    # a 400 k-document slice for the oracle: same posting values, local ids
    from bayesian_bm25_b200 import index_build
    lo, hi = 20_000_000, 20_400_000
    hosts = {f: _host(index_build.shard_csc(cscs[f], lo, hi)) for f in mf.fields}
    del cscs
    torch.cuda.empty_cache()
    params = {}
    for f in mf.fields:
        t = mf._scorers[f].transform
        params[f] = coracle.make_params(t.alpha, t.beta, t.base_rate)
    qs = _queries(nq, vocab, 43)
    fq = [_flat(qs), _flat(qs)]
    mf.set_pruning(0)
    ids0, vals0 = mf.retrieve_ids_batch(fq, 10)
    mf.set_pruning(3)
    ids3, vals3 = mf.retrieve_ids_batch(fq, 10)
    st = mf.stats()
    np.testing.assert_array_equal(ids0, ids3)
    np.testing.assert_array_equal(vals0, vals3)
    assert st["units_skipped"] > 0 and st["fallback_queries"] == 0
    for q in range(nq):
        fused = mf._fused_device([qs[q], qs[q]])
        d_ids, d_vals = mf._topk_device(fused, 10)
        np.testing.assert_array_equal(ids3[q], d_ids.cpu().numpy())
        np.testing.assert_array_equal(vals3[q], d_vals.cpu().numpy())
        if q < 4:
            want = _oracle_fused(mf, hosts, params, [qs[q], qs[q]])
            np.testing.assert_allclose(fused[lo:hi].cpu().numpy(), want, rtol=1e-9, atol=1e-18)
        del fused


@pytest.mark.parametrize("lookup_div", ["64", "0", "100000"])
def test_essential_posting_evaluation_of_the_fused_key(lookup_div, monkeypatch):
    """Two fields, pruning level >= 2: units are evaluated four queries at a time through their essential
    postings (MaxScore split per block on the fused key; non-essential values from the hot / lookup rows).
    Default lookup rows, none (BB25_LOOKUP_DIV=0), rows for every term, and the evaluation switched off
    (BB25_FUSED_SPARSE=0) all return the exhaustive result, which is checked against the oracle; queries with
    9..30 terms (evaluated one by one), duplicates, rare-only and frequent-only queries included."""
    monkeypatch.setenv("BB25_LOOKUP_DIV", lookup_div)
    vocab = 6000
    mf, hosts, params = _two_field(200_000, vocab, seed=211, title_len=7.0, body_len=45.0)
    rng = np.random.default_rng(5)
    extra = [[5999], [5998, 5997], [0, 1, 2], [0, 0, 3000, 3000], rng.integers(0, vocab, 10), rng.integers(0, vocab, 20),
             rng.integers(0, vocab, 30), rng.integers(0, 40, 9), rng.integers(2500, vocab, 8), [3, 4500]]
    qs = _queries(150, vocab, 9, extra)
    fq = [_flat(qs), _flat(qs)]
    mf.set_pruning(0)
    want = {k: mf.retrieve_ids_batch(fq, k) for k in (10, 100)}
    from oracle import coracle
    for q in (0, 1, 2, 150, 153, 154, 159):  # the exhaustive result against the oracle
        w_ids, w_vals = coracle.topk_f64(_oracle_fused(mf, hosts, params, [qs[q], qs[q]]), 10)
        np.testing.assert_array_equal(want[10][0][q], w_ids, err_msg=f"query {q}")
        np.testing.assert_allclose(want[10][1][q], w_vals, rtol=1e-9, atol=1e-18)
    stats = {}
    for level in (2, 3):
        for sparse in ("1", "0"):
            monkeypatch.setenv("BB25_FUSED_SPARSE", sparse)
            mf.set_pruning(level)
            for k in (10, 100):
                ids, vals = mf.retrieve_ids_batch(fq, k)
                np.testing.assert_array_equal(ids, want[k][0], err_msg=f"level {level} sparse {sparse} k {k}")
                np.testing.assert_array_equal(vals, want[k][1], err_msg=f"level {level} sparse {sparse} k {k}")
            stats[(level, sparse)] = mf.stats()
    assert stats[(3, "1")]["units_sparse"] > 0 and stats[(3, "0")]["units_sparse"] == 0
