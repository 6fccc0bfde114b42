"""Vectorised construction of the bm25s-style CSC score matrix (host-side plumbing).

Replaces, for token-id input, what ``bm25s.BM25.index`` produces for the reference
(``bayesian_bm25/scorer.py:262``; its ``.scores`` dict is read at ``scorer.py:227``)
without the per-document Python loop, so an 8.8 M-document corpus is indexed in
seconds.  The posting values follow bm25s's arithmetic operation by operation
(NumPy-2 promotion rules: the per-document length term is a float64 scalar, so the
term-frequency component is evaluated in float64 and the product with the fp32 idf
is rounded to fp32 once) -- tests/test_host_logic.py checks bit-equality against an
independent per-document restatement for all three variants.

Runs on whatever device the input tensors live on (CUDA in production and in the
benchmark; CPU in the unit tests).  This is index-time code, not the query path.
"""
from __future__ import annotations

import math

import numpy as np
import torch

VALID_METHODS = ("robertson", "lucene", "atire")


def idf_table(method: str, df: np.ndarray, n_docs: int) -> np.ndarray:
    """fp32 idf per term, through Python's math.log like bm25s does (libm, then one
    rounding to fp32).  Terms with df == 0 get 0."""
    out = np.zeros(len(df), dtype=np.float32)
    nz = np.nonzero(df)[0]
    if method == "robertson":
        for t in nz:
            inner = (n_docs - int(df[t]) + 0.5) / (int(df[t]) + 0.5)
            out[t] = math.log(inner if inner >= 1 else 1)
    elif method == "lucene":
        for t in nz:
            out[t] = math.log(1 + (n_docs - int(df[t]) + 0.5) / (int(df[t]) + 0.5))
    elif method == "atire":
        for t in nz:
            out[t] = math.log(n_docs / int(df[t]))
    else:
        raise ValueError(f"method must be one of {VALID_METHODS}, got {method!r}")
    return out


def csc_piece(keys: torch.Tensor, n_docs: int, n_vocab: int, doc_len_dev: torch.Tensor, l_avg: float,
              k1: float, b: float, method: str):
    """Postings of the terms covered by `keys` (SORTED ``term * n_docs + doc``, one entry per
    token occurrence): (values fp32, doc ids int32, df int64[n_vocab]).  Terms never straddle
    pieces, so a piece's document frequencies (hence idf) are complete."""
    dev = keys.device
    uniq, tf = torch.unique_consecutive(keys, return_counts=True)
    del keys
    term = torch.div(uniq, n_docs, rounding_mode="floor")
    doc = uniq - term * n_docs
    del uniq
    df = torch.bincount(term, minlength=n_vocab)
    idf = torch.from_numpy(idf_table(method, df.cpu().numpy(), n_docs)).to(dev)
    tf32 = tf.to(torch.float32)
    del tf
    l_d = doc_len_dev[doc].to(torch.float64)
    # k1 * ((1 - b) + b * l_d / l_avg)   -- same association as bm25s
    x = (l_d * b) / l_avg
    x = (1 - b) + x
    x = x * k1
    tf64 = tf32.to(torch.float64)
    if method == "atire":
        # (tf * (k1 + 1)) is float32 * python-float -> float32 under NEP 50
        num = (tf32 * torch.tensor(k1 + 1, dtype=torch.float32, device=dev)).to(torch.float64)
        tfc = num / (tf64 + x)
    else:
        tfc = tf64 / (x + tf64)
    vals = (idf[term].to(torch.float64) * tfc).to(torch.float32)
    return vals, doc.to(torch.int32), df


def csc_from_pieces(pieces, n_docs: int, n_vocab: int, doc_len: torch.Tensor, l_avg: float) -> dict:
    """Assemble pieces (in ascending term order) into the CSC dict."""
    dev = pieces[0][0].device
    df = pieces[0][2].clone()
    for p in pieces[1:]:
        df += p[2]
    indptr = torch.zeros(n_vocab + 1, dtype=torch.int64, device=dev)
    torch.cumsum(df, 0, out=indptr[1:])
    return {
        "data": torch.cat([p[0] for p in pieces]) if len(pieces) > 1 else pieces[0][0],
        "indices": torch.cat([p[1] for p in pieces]) if len(pieces) > 1 else pieces[0][1],
        "indptr": indptr,
        "doc_len": doc_len.to(torch.int32).to(dev),
        "num_docs": n_docs,
        "n_vocab": n_vocab,
        "avgdl": float(l_avg),
    }


def csc_from_sorted_keys(keys: torch.Tensor, n_docs: int, n_vocab: int, doc_len: torch.Tensor,
                         k1: float, b: float, method: str) -> dict:
    """keys: SORTED int64 tensor of ``term * n_docs + doc`` for every token
    occurrence.  doc_len: int64/int32 [n_docs] token counts."""
    if method not in VALID_METHODS:
        raise ValueError(f"method must be one of {VALID_METHODS}, got {method!r}")
    total_tokens = int(doc_len.sum().item())
    l_avg = total_tokens / n_docs  # == np.array(lens).mean(): exact integer sum, one division
    piece = csc_piece(keys, n_docs, n_vocab, doc_len.to(keys.device), l_avg, k1, b, method)
    return csc_from_pieces([piece], n_docs, n_vocab, doc_len, l_avg)


def build_csc(token_ids: torch.Tensor, doc_offsets: torch.Tensor, n_vocab: int, k1: float = 1.5,
              b: float = 0.75, method: str = "lucene") -> dict:
    """token_ids: flat integer tensor of all documents' token ids, doc_offsets: int64
    [n_docs+1].  Returns CSC tensors on token_ids' device."""
    dev = token_ids.device
    doc_offsets = doc_offsets.to(dev, torch.int64)
    n_docs = doc_offsets.numel() - 1
    doc_len = doc_offsets[1:] - doc_offsets[:-1]
    doc_of_token = torch.repeat_interleave(torch.arange(n_docs, device=dev, dtype=torch.int64), doc_len)
    keys = token_ids.to(torch.int64) * n_docs + doc_of_token
    del doc_of_token
    keys, _ = torch.sort(keys)
    return csc_from_sorted_keys(keys, n_docs, n_vocab, doc_len, k1, b, method)


def shard_csc(csc: dict, lo: int, hi: int) -> dict:
    """Document-range shard [lo, hi) of a CSC: local doc ids, same posting values, the
    GLOBAL avgdl (idf and length normalisation are baked into the values, so shard
    scores are bit-identical to the unsharded index -- SURVEY 8e)."""
    idx = csc["indices"]
    keep = (idx >= lo) & (idx < hi)
    n_vocab = csc["indptr"].numel() - 1
    csum = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=idx.device)
    torch.cumsum(keep.to(torch.int64), 0, out=csum[1:])
    indptr = csum[csc["indptr"]]
    return {
        "data": csc["data"][keep],
        "indices": (idx[keep] - lo).to(torch.int32),
        "indptr": indptr,
        "doc_len": csc["doc_len"][lo:hi],
        "num_docs": hi - lo,
        "n_vocab": n_vocab,
        "avgdl": csc["avgdl"],
        "doc_id_offset": lo + int(csc.get("doc_id_offset", 0)),
    }


def shard_bounds(n_docs: int, n_shards: int) -> list[tuple[int, int]]:
    """Contiguous doc ranges [s*N/S, (s+1)*N/S)."""
    return [(s * n_docs // n_shards, (s + 1) * n_docs // n_shards) for s in range(n_shards)]
