"""MultiFieldScorer on the B200 path (reference: ``bayesian_bm25/multi_field.py``).

One BayesianBM25Scorer (device index) per field.  The weighted log-odds conjunction
is fused into each field's traversal pass (bb25_fuse_bm25_signal), the fused vector is
ranked by the dense top-k kernel -- nothing leaves the device until the final (k,)
result.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, fused
from .fusion import _resolve_alpha
from .scorer import BayesianBM25Scorer

MAX_DEVICE_TOPK = 8192


class MultiFieldScorer:
    """Fuses per-field Bayesian probabilities (multi_field.py:24-236)."""

    def __init__(self, fields: list[str], field_weights: dict[str, float] | None = None,
                 alpha: float | str | None = "auto", base_rate: float | str | None = None,
                 k1: float = 1.2, b: float = 0.75, method: str = "robertson") -> None:
        if not fields:
            raise ValueError("fields must be a non-empty list")
        if len(fields) != len(set(fields)):
            raise ValueError("fields must not contain duplicates")
        self._fields = list(fields)
        self._alpha = alpha
        self._base_rate = base_rate
        self._k1, self._b, self._method = k1, b, method
        if field_weights is None:
            self._field_weights = {f: 1.0 / len(fields) for f in fields}
        else:
            for f in fields:
                if f not in field_weights:
                    raise ValueError(f"field_weights missing key {f!r}")
            total = sum(field_weights[f] for f in fields)
            if abs(total - 1.0) > 1e-6:
                raise ValueError(f"field_weights must sum to 1, got {total}")
            self._field_weights = {f: field_weights[f] for f in fields}
        self._scorers: dict[str, BayesianBM25Scorer] = {}
        self._num_docs = 0

    @property
    def num_docs(self) -> int:
        return self._num_docs

    @property
    def fields(self) -> list[str]:
        return list(self._fields)

    @property
    def field_weights(self) -> dict[str, float]:
        return dict(self._field_weights)

    def _new_scorer(self) -> BayesianBM25Scorer:
        return BayesianBM25Scorer(k1=self._k1, b=self._b, method=self._method, base_rate=self._base_rate)

    def index(self, documents: list[dict[str, list[str]]], show_progress: bool = True) -> None:
        """One index per field (multi_field.py:105-139)."""
        for i, doc in enumerate(documents):
            for f in self._fields:
                if f not in doc:
                    raise ValueError(f"Document {i} missing field {f!r}")
        self._scorers = {}
        for f in self._fields:
            s = self._new_scorer()
            s.index([doc[f] for doc in documents], show_progress=show_progress)
            self._scorers[f] = s
        self._num_docs = len(documents)

    def index_from_csc(self, field_csc: dict[str, dict], pseudo_queries: dict | None = None) -> None:
        """Extension: adopt prebuilt per-field CSCs (same documents in every field)."""
        self._scorers = {}
        for f in self._fields:
            s = self._new_scorer()
            s.index_from_csc(field_csc[f], pseudo_queries=(pseudo_queries or {}).get(f))
            self._scorers[f] = s
        self._num_docs = self._scorers[self._fields[0]].num_docs

    def add_documents(self, new_documents: list[dict[str, list[str]]], show_progress: bool = True) -> None:
        if not self._scorers:
            raise RuntimeError("Call index() before add_documents().")
        for i, doc in enumerate(new_documents):
            for f in self._fields:
                if f not in doc:
                    raise ValueError(f"New document {i} missing field {f!r}")
        for f in self._fields:
            self._scorers[f].add_documents([doc[f] for doc in new_documents], show_progress=show_progress)
        self._num_docs += len(new_documents)

    def _check_weights(self) -> None:
        # the reference passes the weights to log_odds_conjunction at query time, which rejects
        # negative ones (fusion.py:253-255)
        if any(w < 0 for w in self._field_weights.values()):
            raise ValueError("weights must be non-negative")

    def _fused_device(self, per_field_terms) -> torch.Tensor:
        """Fused probability of every document, on the device.  One pass per field: the
        field's traversal kernel turns its BM25 accumulators into the posterior, takes
        the logit and adds w_f * logit into ONE shared accumulator; the last field
        applies n**alpha and the sigmoid (multi_field.py:158-174 without the [N, F]
        matrix)."""
        self._check_weights()
        nf = len(self._fields)
        first = self._scorers[self._fields[0]]
        acc = torch.empty(self._num_docs, dtype=torch.float64, device=first._device)
        scale = float(nf ** _resolve_alpha(self._alpha, default=0.5))
        for j, f in enumerate(self._fields):
            flags = (1 if j == 0 else 0) | (2 if j == nf - 1 else 0)
            self._scorers[f].fuse_signal_device(per_field_terms[j], self._field_weights[f], nf, scale, flags, acc)
        return acc

    def get_probabilities(self, query_tokens: list[str]) -> np.ndarray:
        """Fused probability of every document (multi_field.py:141-174)."""
        if not self._scorers:
            raise RuntimeError("Call index() before get_probabilities().")
        terms = [self._scorers[f]._term_ids(query_tokens) for f in self._fields]
        return self._fused_device(terms).cpu().numpy()

    def _topk_device(self, fused: torch.Tensor, k: int):
        k = min(k, fused.numel())
        if k <= MAX_DEVICE_TOPK:
            ids = torch.empty(k, dtype=torch.int64, device=fused.device)
            vals = torch.empty(k, dtype=torch.float64, device=fused.device)
            _lib.check(_lib.lib().bb25_topk_f64(fused.device.index, fused.data_ptr(), fused.numel(), k,
                                                ids.data_ptr(), vals.data_ptr(), _lib.stream_ptr()))
            return ids, vals
        # very large k: full stable device sort (value desc, id asc)
        order = torch.sort(fused, descending=True, stable=True).indices[:k]
        return order, fused[order]

    def retrieve(self, query_tokens: list[str], k: int = 10):
        """Top-k by fused probability (multi_field.py:176-200); ties resolved by
        ascending doc id (the reference's argsort leaves them unspecified)."""
        if not self._scorers:
            raise RuntimeError("Call index() before get_probabilities().")
        terms = [self._scorers[f]._term_ids(query_tokens) for f in self._fields]
        ids, vals = self._topk_device(self._fused_device(terms), k)
        return ids.cpu().numpy(), vals.cpu().numpy()

    def retrieve_ids(self, per_field_terms, k: int = 10):
        """Extension: query given as one term-id list per field."""
        ids, vals = self._topk_device(self._fused_device(per_field_terms), k)
        return ids.cpu().numpy(), vals.cpu().numpy()

    # ---- batched retrieval (extension): the per-query loop of the reference's consumers, absorbed ----
    def retrieve_ids_batch_device(self, field_queries, k: int = 10):
        """field_queries: per field, in `fields` order, (flat term ids int32, offsets int64[Q+1]).
        Returns CUDA tensors (ids int64 [Q,k], fused fp64 [Q,k]); top-k by fused probability with
        block-max pruning on the fused key (bb25_retrieve_fused_batch)."""
        if not self._scorers:
            raise RuntimeError("Call index() before retrieve_batch().")
        self._check_weights()
        nf = len(self._fields)
        scale = float(nf ** _resolve_alpha(self._alpha, default=0.5))
        weights = [self._field_weights[f] for f in self._fields]
        return fused.retrieve_fused_batch_device([self._scorers[f] for f in self._fields], field_queries, k, scale, weights)

    def retrieve_ids_batch(self, field_queries, k: int = 10):
        ids, probs = self.retrieve_ids_batch_device(field_queries, k)
        return ids.cpu().numpy(), probs.cpu().numpy()

    def retrieve_batch(self, queries: list[list[str]], k: int = 10):
        """Top-k by fused probability for a batch of token-list queries:
        (doc_ids [Q,k] int64, fused probabilities [Q,k] float64); row q equals retrieve(queries[q], k)."""
        if not self._scorers:
            raise RuntimeError("Call index() before retrieve_batch().")
        fq = [self._scorers[f]._term_ids_batch(queries) for f in self._fields]
        return self.retrieve_ids_batch(fq, k)

    def stats(self) -> dict:
        """Counters of the last retrieve_batch (units visited / skipped by the block bound, candidates,
        queries that took the dense guaranteed path, host synchronisations)."""
        return fused.fused_stats(self._scorers[self._fields[0]])

    def set_pruning(self, level: int) -> None:
        """0: exhaustive traversal; >= 1: block-max pruning on the fused key.  Results are identical."""
        self._scorers[self._fields[0]].set_pruning(level)
