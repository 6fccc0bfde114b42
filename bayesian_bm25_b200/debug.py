"""Trace records of ``retrieve(explain=True)`` (reference: ``bayesian_bm25/debug.py:39-64``,
``FusionDebugger.trace_bm25`` :178-216).  The intermediate values are computed on the device
(bb25_match_counts + bb25_trace_bm25); the rest of the reference's FusionDebugger (fusion traces,
formatting) is outside the hot path."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class BM25SignalTrace:
    """One BM25 signal through the probability pipeline (same fields as the reference's dataclass)."""

    raw_score: float
    tf: float
    doc_len_ratio: float
    likelihood: float
    tf_prior: float
    norm_prior: float
    composite_prior: float
    logit_likelihood: float
    logit_prior: float
    logit_base_rate: float | None
    posterior: float
    alpha: float
    beta: float
    base_rate: float | None
