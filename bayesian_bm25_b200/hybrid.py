"""Hybrid sparse + dense fusion (BASELINE configs[3]; pattern of the reference's
``benchmarks/hybrid_beir.py:1708-1765`` and ``README.md:129-140``):

    fused(d) = log_odds_conjunction([P_bm25(d | q), cosine_to_probability(cos(q, d))],
                                    alpha, weights)

evaluated for every document in two fused device passes -- the BM25 signal inside the
traversal kernel's epilogue, the cosine signal as one streaming pass over the fp32
similarities -- and ranked by the dense top-k kernel.  The cosine similarities are an
INPUT (the embedding GEMM that produces them is outside this path).
BM25-inactive documents enter with probability 0 (clamped to 1e-10 inside the
conjunction), exactly as ``hybrid_beir.py:398-411`` + ``fusion.py:243`` do.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, fused
from .fusion import _check_weights, _resolve_alpha


def hybrid_probabilities_device(scorer, term_ids, cosine: torch.Tensor, weights=(0.6, 0.4), alpha=None) -> torch.Tensor:
    """fp64 [N] fused probabilities on the device.  cosine: fp32 CUDA tensor [N]."""
    scorer._require_index("hybrid_probabilities_device()")
    n = scorer.num_docs
    if cosine.numel() != n:
        raise ValueError(f"cosine must have {n} elements, got {cosine.numel()}")
    cosine = cosine.to(device=scorer._device, dtype=torch.float32).contiguous()
    if weights is not None:
        w = _check_weights(weights, 2)
        scale = float(2 ** _resolve_alpha(alpha, default=0.0))
        unw = 0
    else:
        w = np.array([1.0, 1.0])
        scale = float(2 ** _resolve_alpha(alpha, default=0.5))
        unw = 4
    acc = torch.empty(n, dtype=torch.float64, device=scorer._device)
    scorer.fuse_signal_device(term_ids, float(w[0]), 2, scale, 1 | unw, acc)
    _lib.check(_lib.lib().bb25_fuse_cosine_signal(scorer._device.index, cosine.data_ptr(), n, float(w[1]), 2, scale,
                                                  2 | unw, acc.data_ptr(), _lib.stream_ptr()))
    return acc


def topk_device(values: torch.Tensor, k: int):
    """(value desc, index asc) top-k of a dense fp64 CUDA vector (k <= 8192)."""
    k = min(int(k), values.numel())
    ids = torch.empty(k, dtype=torch.int64, device=values.device)
    vals = torch.empty(k, dtype=torch.float64, device=values.device)
    _lib.check(_lib.lib().bb25_topk_f64(values.device.index, values.data_ptr(), values.numel(), k, ids.data_ptr(),
                                        vals.data_ptr(), _lib.stream_ptr()))
    return ids, vals


def hybrid_retrieve(scorer, term_ids, cosine: torch.Tensor, k: int = 100, weights=(0.6, 0.4), alpha=None):
    """Top-k documents by fused probability: (ids int64 [k], fused fp64 [k]) NumPy arrays."""
    ids, vals = topk_device(hybrid_probabilities_device(scorer, term_ids, cosine, weights, alpha), k)
    return ids.cpu().numpy() + scorer._doc_id_offset, vals.cpu().numpy()


def hybrid_retrieve_batch_device(scorer, q_terms, q_off, cosine: torch.Tensor, k: int = 100, weights=(0.6, 0.4),
                                 alpha=None):
    """A batch of hybrid queries: q_terms / q_off (flat in-vocabulary ids + offsets, NumPy), cosine fp32
    CUDA [Q, N] (or [Q, stride], see fused.pad_cosine).  Returns CUDA (ids int64 [Q,k], fused fp64 [Q,k]),
    row q equal to hybrid_retrieve(scorer, query q, cosine[q], k, weights, alpha)."""
    scorer._require_index("hybrid_retrieve_batch()")
    if cosine.shape[1] < scorer.num_docs:
        raise ValueError(f"cosine must have at least {scorer.num_docs} columns, got {cosine.shape[1]}")
    cosine = fused.pad_cosine(cosine.to(device=scorer._device, dtype=torch.float32))
    if weights is not None:
        w = _check_weights(weights, 2)
        scale = float(2 ** _resolve_alpha(alpha, default=0.0))
        return fused.retrieve_fused_batch_device([scorer], [(q_terms, q_off)], k, scale, [float(w[0])], cosine, float(w[1]))
    scale = float(2 ** _resolve_alpha(alpha, default=0.5))
    return fused.retrieve_fused_batch_device([scorer], [(q_terms, q_off)], k, scale, None, cosine, 1.0)


def hybrid_retrieve_batch(scorer, q_terms, q_off, cosine: torch.Tensor, k: int = 100, weights=(0.6, 0.4), alpha=None):
    ids, probs = hybrid_retrieve_batch_device(scorer, q_terms, q_off, cosine, k, weights, alpha)
    return ids.cpu().numpy(), probs.cpu().numpy()


def hybrid_retrieve_batch_embeddings(scorer, q_terms, q_off, query_emb: torch.Tensor, corpus_emb: torch.Tensor,
                                     k: int = 100, weights=(0.6, 0.4), alpha=None, sub_batch: int = 256):
    """The whole hybrid step from embeddings: cosine GEMM on the tensor cores (dense.cosine_scores), then the
    fused-rank batch retrieval, one sub-batch of queries at a time so that the transient [sub_batch, N] cosine
    tile stays small.  Returns NumPy (ids int64 [Q,k], fused fp64 [Q,k])."""
    import numpy as np

    from . import dense
    q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64)
    nq = q_off.size - 1
    ids = np.empty((nq, k), dtype=np.int64)
    probs = np.empty((nq, k), dtype=np.float64)
    cos = None
    for s in range(0, nq, sub_batch):
        e = min(nq, s + sub_batch)
        cos = dense.cosine_scores(query_emb[s:e], corpus_emb, out=cos[:e - s] if cos is not None else None)
        i_, p_ = hybrid_retrieve_batch_device(scorer, q_terms[q_off[s]:q_off[e]], q_off[s:e + 1] - q_off[s], cos[:e - s],
                                              k, weights, alpha)
        ids[s:e], probs[s:e] = i_.cpu().numpy(), p_.cpu().numpy()
    return ids, probs
