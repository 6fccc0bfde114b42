"""Batched top-k by fused probability: the device path behind
``MultiFieldScorer.retrieve_batch`` and ``hybrid.hybrid_retrieve_batch``
(bb25_retrieve_fused_batch, include/bb25.h).

The reference evaluates ``log_odds_conjunction`` over per-signal probability
vectors of every document, one query at a time (``multi_field.py:141-200``,
``benchmarks/hybrid_beir.py:1708-1765``); here a whole batch of queries is
ranked by the fused probability with block-max pruning on the fused key, and
only the (Q, k) result leaves the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_FUSED_K = 1024
MAX_FIELDS = 4


def _flat_queries(per_query_terms):
    """list of per-query term-id sequences -> (flat int32, offsets int64[Q+1])."""
    off = np.zeros(len(per_query_terms) + 1, dtype=np.int64)
    if len(per_query_terms):
        np.cumsum([len(q) for q in per_query_terms], out=off[1:])
    flat = (np.concatenate([np.asarray(q, dtype=np.int32) for q in per_query_terms]).astype(np.int32)
            if off[-1] > 0 else np.zeros(0, dtype=np.int32))
    return flat, off


def retrieve_fused_batch_device(scorers, field_queries, k: int, scale: float, weights=None,
                                cosine: torch.Tensor | None = None, cos_weight: float = 0.0):
    """scorers: indexed BayesianBM25Scorer per BM25 field (same documents);
    field_queries: per field (flat term ids int32, offsets int64[Q+1]) NumPy arrays;
    weights: per-field conjunction weights (None = unweighted mean branch);
    cosine: optional fp32 CUDA tensor [Q, stride] (stride % 4 == 0, >= num_docs), the last signal.
    Returns CUDA tensors (ids int64 [Q,k], fused fp64 [Q,k])."""
    nf = len(scorers)
    if not 1 <= nf <= MAX_FIELDS:
        raise ValueError(f"1..{MAX_FIELDS} BM25 fields are supported, got {nf}")
    for s in scorers:
        s._require_index("retrieve_batch()")
    first = scorers[0]
    dev = first._device
    nq = len(field_queries[0][1]) - 1
    if k > first.num_docs:
        raise ValueError(f"k of {k} is larger than the number of available scores, which is {first.num_docs}")
    if k > MAX_FUSED_K:
        raise ValueError(f"the fused batch path supports k <= {MAX_FUSED_K}; use the per-query retrieve() for larger k")
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    probs = torch.empty((nq, k), dtype=torch.float64, device=dev)
    if nq == 0:
        return ids, probs
    arr = (_lib.FusedField * nf)()
    keep = []
    for i, (s, (flat, off)) in enumerate(zip(scorers, field_queries)):
        flat = np.ascontiguousarray(flat, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        if off.size != nq + 1:
            raise ValueError("every field needs offsets for the same number of queries")
        d_flat = torch.from_numpy(flat if flat.size else np.zeros(1, np.int32)).to(dev)
        d_off = torch.from_numpy(off).to(dev)
        keep += [d_flat, d_off]
        arr[i].index = s._handle
        arr[i].params = s._params()
        arr[i].weight = float(weights[i]) if weights is not None else 1.0
        arr[i].q_terms = d_flat.data_ptr()
        arr[i].q_off = d_off.data_ptr()
        arr[i].term_base = int(off[0])
        arr[i].n_terms_total = int(off[-1] - off[0])
    cos_ptr, cos_stride = None, 0
    if cosine is not None:
        if cosine.dtype != torch.float32 or cosine.dim() != 2 or cosine.shape[0] != nq or not cosine.is_cuda:
            raise ValueError("cosine must be a float32 CUDA tensor of shape [Q, stride]")
        if cosine.stride(1) != 1 or cosine.stride(0) % 4 or cosine.stride(0) < first.num_docs or cosine.data_ptr() % 16:
            raise ValueError("cosine rows must be contiguous, 16-byte aligned, with a row stride that is a multiple of 4 "
                             "and >= num_docs (see pad_cosine)")
        cos_ptr, cos_stride = cosine.data_ptr(), cosine.stride(0)
    _lib.check(_lib.lib().bb25_retrieve_fused_batch(
        nf, arr, cos_ptr, cos_stride, float(cos_weight), int(weights is not None), float(scale), nq, k,
        ids.data_ptr(), probs.data_ptr(), _lib.stream_ptr()))
    del keep
    return ids, probs


def pad_cosine(cosine: torch.Tensor) -> torch.Tensor:
    """[Q, N] fp32 -> the same values with a row stride rounded up to a multiple of 4 floats."""
    q, n = cosine.shape
    stride = (n + 3) // 4 * 4
    if stride == n and cosine.is_contiguous() and cosine.data_ptr() % 16 == 0:
        return cosine
    buf = torch.zeros((q, stride), dtype=torch.float32, device=cosine.device)
    buf[:, :n] = cosine
    return buf


def fused_stats(scorer) -> dict:
    """Counters of the last fused batch whose first field was `scorer`."""
    v = [C.c_int64() for _ in range(7)]
    ms = C.c_double()
    _lib.check(_lib.lib().bb25_fused_stats(scorer._handle, *[C.byref(x) for x in v], C.byref(ms)))
    names = ("units", "units_skipped", "units_abandoned", "candidates", "fallback_queries", "rerun_queries", "host_syncs")
    out = dict(zip(names, (x.value for x in v)))
    out["traverse_ms"] = ms.value
    w = [C.c_int64() for _ in range(3)]
    _lib.check(_lib.lib().bb25_fused_prune_stats(scorer._handle, *[C.byref(x) for x in w]))
    out.update(zip(("units_no_essential", "units_sparse", "sparse_documents"), (x.value for x in w)))
    return out
