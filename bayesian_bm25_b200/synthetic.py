"""Synthetic corpora shaped like the reference's own scalability benchmark.

``scalability_corpus`` reproduces ``benchmarks/scalability.py:34-67``
(``generate_synthetic_corpus``) call for call, so seed 42 gives the very corpus
BASELINE configs[0] is quoted on.  ``zipf_corpus_ids`` is the vectorised,
token-id form used for the 8.8 M-document configurations (same distribution:
Zipf(1) vocabulary, doc length max(5, int(N(avg, 0.3 avg))); not the same random
stream -- SURVEY 8d).  Token draws use a counter-based hash so the corpus is
identical whether it is generated on the CPU or on a GPU, and any rank can
generate it independently.
"""
from __future__ import annotations

import numpy as np
import torch


def scalability_corpus(n_docs: int, vocab_size: int, avg_doc_len: int, rng: np.random.Generator):
    """(corpus_tokens, queries) exactly as the reference's generator draws them."""
    vocab = [f"term_{i}" for i in range(vocab_size)]
    w = 1.0 / np.arange(1, vocab_size + 1)
    w /= w.sum()
    corpus = []
    for _ in range(n_docs):
        n = max(5, int(rng.normal(avg_doc_len, avg_doc_len * 0.3)))
        corpus.append([vocab[i] for i in rng.choice(vocab_size, size=n, p=w)])
    queries = []
    for _ in range(min(100, n_docs // 10)):
        m = rng.integers(3, 6)
        queries.append([vocab[i] for i in rng.choice(vocab_size, size=m, p=w)])
    return corpus, queries


_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    return (z >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(seed: int, idx: torch.Tensor) -> torch.Tensor:
    """splitmix64(seed + idx) -> float64 uniform in [0, 1); integer-only, so
    bit-identical on CPU and CUDA."""
    z = idx.to(torch.int64) * _s64(0x9E3779B97F4A7C15) + _s64(seed * 0xD1342543DE82EF95 + 0x632BE59BD9B4E019)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def zipf_cdf(vocab_size: int) -> np.ndarray:
    w = 1.0 / np.arange(1, vocab_size + 1)
    w /= w.sum()
    c = np.cumsum(w)
    c[-1] = 1.0
    return c


def zipf_doc_lengths(n_docs: int, avg_doc_len: float, seed: int, min_len: int = 5) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.maximum(min_len, rng.normal(avg_doc_len, avg_doc_len * 0.3, n_docs).astype(np.int64))


def zipf_sorted_keys(n_docs: int, vocab_size: int, doc_len: np.ndarray, seed: int, device,
                     chunk: int = 1 << 26) -> torch.Tensor:
    """Sorted ``term * n_docs + doc`` keys of every token of the corpus (the input of
    index_build.csc_from_sorted_keys), generated in chunks on `device`."""
    cdf = torch.from_numpy(zipf_cdf(vocab_size)).to(device)
    offs = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(doc_len, out=offs[1:])
    total = int(offs[-1])
    offs_t = torch.from_numpy(offs).to(device)
    keys = torch.empty(total, dtype=torch.int64, device=device)
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        pos = torch.arange(s, e, device=device, dtype=torch.int64)
        term = torch.searchsorted(cdf, hash_uniform(seed, pos)).clamp_(max=vocab_size - 1)
        doc = torch.searchsorted(offs_t, pos, right=True) - 1
        keys[s:e] = term * n_docs + doc
        del pos, term, doc
    keys, _ = torch.sort(keys)
    return keys


def zipf_queries(n_queries: int, vocab_size: int, seed: int):
    """3-5 Zipf-sampled term ids per query (benchmarks/scalability.py:59-65 shape).
    Returns (q_terms int32[total], q_off int64[Q+1])."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, vocab_size + 1)
    w /= w.sum()
    lens = rng.integers(3, 6, n_queries)
    off = np.zeros(n_queries + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    terms = rng.choice(vocab_size, size=int(off[-1]), p=w).astype(np.int32)
    return terms, off


def zipf_pseudo_queries(n_docs: int, vocab_size: int, avg_doc_len: float, seed: int, min_len: int = 5,
                        n_sample: int = 50, n_tokens: int = 5) -> list[np.ndarray]:
    """The pseudo-queries BayesianBM25Scorer.index() draws (scorer.py:295-311: the first five tokens
    of 50 documents picked by default_rng(42)) for the corpus zipf_csc(n_docs, ..., seed) encodes,
    regenerated from the counter-based token hash without materialising the corpus."""
    dl = zipf_doc_lengths(n_docs, avg_doc_len, seed, min_len)
    offs = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(dl, out=offs[1:])
    docs = np.random.default_rng(42).choice(n_docs, size=min(n_docs, n_sample), replace=False)
    cdf = torch.from_numpy(zipf_cdf(vocab_size))
    out = []
    for d in docs:
        pos = torch.arange(int(offs[d]), int(offs[d]) + min(n_tokens, int(dl[d])), dtype=torch.int64)
        term = torch.searchsorted(cdf, hash_uniform(seed, pos)).clamp_(max=vocab_size - 1)
        out.append(term.numpy().astype(np.int32))
    return out


def zipf_csc(n_docs: int, vocab_size: int, avg_doc_len: float, seed: int, device, k1=1.2, b=0.75,
             method="lucene", min_len: int = 5, n_buckets: int | None = None) -> dict:
    """Synthetic corpus -> CSC tensors on `device` (term ids are Zipf ranks).

    Corpora with more than ~2^31 tokens cannot go through one torch.sort; they are built
    in `n_buckets` term-range pieces (a term's tokens never straddle pieces, the pieces
    concatenate to the same CSC).  n_buckets=None picks 1 or as many as needed."""
    from .index_build import VALID_METHODS, csc_from_pieces, csc_from_sorted_keys, csc_piece

    dl = zipf_doc_lengths(n_docs, avg_doc_len, seed, min_len)
    total = int(dl.sum())
    if n_buckets is None:
        n_buckets = 1 if total < (1 << 30) else int(np.ceil(total / float(1 << 29)))
    if n_buckets <= 1:
        keys = zipf_sorted_keys(n_docs, vocab_size, dl, seed, device)
        return csc_from_sorted_keys(keys, n_docs, vocab_size, torch.from_numpy(dl).to(device), k1, b, method)
    if method not in VALID_METHODS:
        raise ValueError(f"method must be one of {VALID_METHODS}, got {method!r}")
    cdf_np = zipf_cdf(vocab_size)
    # term cut points of (roughly) equal token mass; the head term alone holds ~9 %
    cuts = [0] + [int(np.searchsorted(cdf_np, (i + 1) / n_buckets)) + 1 for i in range(n_buckets - 1)] + [vocab_size]
    cuts = sorted(set(min(max(c, 0), vocab_size) for c in cuts))
    cdf = torch.from_numpy(cdf_np).to(device)
    offs = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(dl, out=offs[1:])
    offs_t = torch.from_numpy(offs).to(device)
    dl_dev = torch.from_numpy(dl).to(device)
    l_avg = total / n_docs
    chunk = 1 << 26
    pieces = []
    for t_lo, t_hi in zip(cuts[:-1], cuts[1:]):
        parts = []
        for s in range(0, total, chunk):
            e = min(total, s + chunk)
            pos = torch.arange(s, e, device=device, dtype=torch.int64)
            term = torch.searchsorted(cdf, hash_uniform(seed, pos)).clamp_(max=vocab_size - 1)
            m = (term >= t_lo) & (term < t_hi)
            pos, term = pos[m], term[m]
            doc = torch.searchsorted(offs_t, pos, right=True) - 1
            parts.append(term * n_docs + doc)
            del pos, term, doc, m
        keys = torch.cat(parts)
        del parts
        keys, _ = torch.sort(keys)
        pieces.append(csc_piece(keys, n_docs, vocab_size, dl_dev, l_avg, k1, b, method))
        del keys
    return csc_from_pieces(pieces, n_docs, vocab_size, dl_dev, l_avg)
