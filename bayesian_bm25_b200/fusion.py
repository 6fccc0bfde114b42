"""cosine_to_probability and log_odds_conjunction of the reference's
``bayesian_bm25/fusion.py`` (:25-45, :103-280), evaluated on the GPU.

Also here: ``balanced_log_odds_fusion`` (:283-333) and the INFERENCE side
of ``AttentionLogOddsWeights`` (:639-828, :1039-1135: ``__call__``, ``compute_upper_bounds``,
``prune``).  Training (``fit`` / ``update``) and the remaining helpers (prob_and/or/not,
LearnableLogOddsWeights, the multi-head wrapper) stay outside the hot path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._device import dev_f64

_SQRT_N_ALPHA = 0.5  # fusion.py:103
_GATINGS = ("none", "relu", "swish", "gelu", "softplus")


def _resolve_alpha(alpha, default: float) -> float:
    """None -> default, "auto" -> 0.5, else float (fusion.py:106-116)."""
    if alpha is None:
        return default
    if isinstance(alpha, str):
        if alpha != "auto":
            raise ValueError(f"alpha must be a float, None, or 'auto', got {alpha!r}")
        return _SQRT_N_ALPHA
    return float(alpha)


def cosine_to_probability(score):
    """(1 + cos) / 2 clamped to [1e-10, 1 - 1e-10] (fusion.py:25-45)."""
    dev = _lib.require_cuda()
    a = np.asarray(score, dtype=np.float64)
    d = dev_f64(np.ascontiguousarray(a).ravel(), f"cuda:{dev}")
    out = torch.empty_like(d)
    _lib.check(_lib.lib().bb25_cosine_to_probability(dev, d.data_ptr(), d.numel(), out.data_ptr(), 1,
                                                     _lib.stream_ptr()))
    res = out.cpu().numpy().reshape(a.shape)
    return float(res) if a.ndim == 0 else res


def _check_weights(weights, n):
    w = np.asarray(weights, dtype=np.float64)
    if np.any(w < 0):
        raise ValueError("weights must be non-negative")
    if abs(float(np.sum(w)) - 1.0) > 1e-6:
        raise ValueError(f"weights must sum to 1, got {float(np.sum(w))}")
    if w.ndim != 1 or w.shape[0] != n:
        raise ValueError(f"weights must have shape ({n},), got {w.shape}")
    return w


def log_odds_conjunction_device(probs: torch.Tensor, alpha=None, weights=None, gating: str = "none",
                                gating_beta: float = 1.0, max_logit: float | None = None) -> torch.Tensor:
    """Device-resident form: probs float64 CUDA tensor (..., n) -> (...)."""
    if gating not in _GATINGS:
        raise ValueError(
            f"gating must be 'none', 'relu', 'swish', 'gelu', or 'softplus', got {gating!r}")
    n = probs.shape[-1]
    p2 = probs.reshape(-1, n).contiguous()
    dev = p2.device.index
    w_t = None
    if weights is not None:
        w = _check_weights(weights, n)
        scale = float(n ** _resolve_alpha(alpha, default=0.0))
        w_t = torch.from_numpy(w).to(p2.device)
    else:
        scale = float(n ** _resolve_alpha(alpha, default=0.5))
    out = torch.empty(p2.shape[0], dtype=torch.float64, device=p2.device)
    _lib.check(_lib.lib().bb25_log_odds_conjunction(
        dev, p2.data_ptr(), p2.shape[0], n, None if w_t is None else w_t.data_ptr(), scale,
        _lib.GATING[gating], float(gating_beta), int(max_logit is not None),
        float(max_logit) if max_logit is not None else 0.0, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(probs.shape[:-1])


def log_odds_conjunction(probs, alpha=None, weights=None, gating: str = "none", gating_beta: float = 1.0,
                         max_logit: float | None = None):
    """sigma(n**alpha * sum_i w_i * logit(P_i)) / sigma(n**alpha * mean logit) with optional
    gating and logit clipping (fusion.py:172-280)."""
    dev = _lib.require_cuda()
    a = np.asarray(probs, dtype=np.float64)
    if a.ndim == 0:
        raise ValueError("probs must have at least one dimension")
    out = log_odds_conjunction_device(dev_f64(a, f"cuda:{dev}"), alpha, weights, gating, gating_beta, max_logit)
    res = out.cpu().numpy()
    return float(res) if res.ndim == 0 else res


def balanced_log_odds_fusion(sparse_probs, dense_similarities, weight: float = 0.5):
    """weight * minmax(logit(cosine_to_probability(dense))) + (1 - weight) * minmax(logit(sparse)),
    min-max over the candidate set (fusion.py:283-333).  Returns fusion scores, not probabilities."""
    dev = _lib.require_cuda()
    sp = np.asarray(sparse_probs, dtype=np.float64)
    de = np.asarray(dense_similarities, dtype=np.float64)
    sp, de = np.broadcast_arrays(sp, de)
    d_sp = dev_f64(np.ascontiguousarray(sp).ravel(), f"cuda:{dev}")
    d_de = dev_f64(np.ascontiguousarray(de).ravel(), f"cuda:{dev}")
    out = torch.empty_like(d_sp)
    _lib.check(_lib.lib().bb25_balanced_fusion(dev, d_sp.data_ptr(), d_de.data_ptr(), d_sp.numel(), float(weight),
                                               out.data_ptr(), _lib.stream_ptr()))
    res = out.cpu().numpy().reshape(sp.shape)
    return float(res) if res.ndim == 0 else res


class AttentionLogOddsWeights:
    """Query-dependent signal weighting via attention -- inference (fusion.py:639-828, 1039-1135).

    Same constructor, initialisation (``default_rng(seed).normal(0, 1/sqrt(n_query_features))``), properties and
    call semantics as the reference; the weight softmax and the weighted log-odds conjunction run on the device.
    ``fit`` / ``update`` are training-time (component 10) and raise; trained parameters are adopted with
    ``set_parameters``."""

    def __init__(self, n_signals: int, n_query_features: int, alpha: float | str = 0.5, normalize: bool = False,
                 seed: int = 0, base_rate: float | None = None) -> None:
        if n_signals < 1:
            raise ValueError(f"n_signals must be >= 1, got {n_signals}")
        if n_query_features < 1:
            raise ValueError(f"n_query_features must be >= 1, got {n_query_features}")
        if base_rate is not None and not (0.0 < base_rate < 1.0):
            raise ValueError(f"base_rate must be in (0, 1), got {base_rate}")
        self._n_signals = n_signals
        self._n_query_features = n_query_features
        self._alpha = _resolve_alpha(alpha, default=0.5)
        self._normalize = normalize
        self._base_rate = base_rate
        self._logit_base_rate = float(np.log(base_rate / (1.0 - base_rate))) if base_rate is not None else None
        rng = np.random.default_rng(seed)
        self._W = rng.normal(0, 1.0 / np.sqrt(n_query_features), size=(n_signals, n_query_features))
        self._b = np.zeros(n_signals, dtype=np.float64)
        self._W_avg = self._W.copy()
        self._b_avg = self._b.copy()

    n_signals = property(lambda self: self._n_signals)
    n_query_features = property(lambda self: self._n_query_features)
    alpha = property(lambda self: self._alpha)
    base_rate = property(lambda self: self._base_rate)
    normalize = property(lambda self: self._normalize)

    @property
    def weights_matrix(self) -> np.ndarray:
        return self._W.copy()

    def set_parameters(self, W, b, W_avg=None, b_avg=None) -> None:
        """Adopt trained parameters (e.g. from the reference's fit())."""
        W = np.asarray(W, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        if W.shape != (self._n_signals, self._n_query_features) or b.shape != (self._n_signals,):
            raise ValueError("W must be (n_signals, n_query_features) and b (n_signals,)")
        self._W, self._b = W.copy(), b.copy()
        self._W_avg = self._W.copy() if W_avg is None else np.asarray(W_avg, dtype=np.float64).copy()
        self._b_avg = self._b.copy() if b_avg is None else np.asarray(b_avg, dtype=np.float64).copy()

    def fit(self, *args, **kwargs):
        raise NotImplementedError("training is outside the B200 hot path; fit with the reference and set_parameters()")

    update = fit

    # ---- device pieces -------------------------------------------------------------------------
    def _weights_device(self, query_features, use_averaged: bool):
        dev = _lib.require_cuda()
        qf = np.atleast_2d(np.asarray(query_features, dtype=np.float64))
        if qf.shape[1] != self._n_query_features:
            raise ValueError(f"query_features must have {self._n_query_features} columns, got {qf.shape[1]}")
        d_qf = dev_f64(np.ascontiguousarray(qf), f"cuda:{dev}")
        d_W = dev_f64(self._W_avg if use_averaged else self._W, f"cuda:{dev}")
        d_b = dev_f64(self._b_avg if use_averaged else self._b, f"cuda:{dev}")
        out = torch.empty((qf.shape[0], self._n_signals), dtype=torch.float64, device=d_qf.device)
        _lib.check(_lib.lib().bb25_attention_weights(dev, d_qf.data_ptr(), d_W.data_ptr(), d_b.data_ptr(), qf.shape[0],
                                                     self._n_query_features, self._n_signals, out.data_ptr(),
                                                     _lib.stream_ptr()))
        return out

    def _compute_weights(self, query_features, use_averaged: bool = False) -> np.ndarray:
        """softmax(W @ features + b) per query (fusion.py:757-772)."""
        w = self._weights_device(query_features, use_averaged).cpu().numpy()
        return w[0] if np.ndim(query_features) == 1 else w

    def fuse_device(self, probs: torch.Tensor, weights: torch.Tensor, normalize: bool) -> torch.Tensor:
        """probs fp64 CUDA [m, n_signals], weights fp64 CUDA [1 or m, n_signals] -> fused fp64 CUDA [m]."""
        m = probs.shape[0]
        if weights.shape[0] not in (1, m):
            raise ValueError("query_features must hold one row, or one row per candidate")
        out = torch.empty(m, dtype=torch.float64, device=probs.device)
        _lib.check(_lib.lib().bb25_attention_fuse(
            probs.device.index, probs.contiguous().data_ptr(), m, self._n_signals, weights.contiguous().data_ptr(),
            weights.shape[0], float(self._n_signals ** self._alpha), int(self._logit_base_rate is not None),
            float(self._logit_base_rate or 0.0), int(bool(normalize)), out.data_ptr(), _lib.stream_ptr()))
        return out

    def __call__(self, probs, query_features, use_averaged: bool = False):
        """Attention-weighted log-odds conjunction (fusion.py:774-828)."""
        dev = _lib.require_cuda()
        p = np.asarray(probs, dtype=np.float64)
        w = self._weights_device(query_features, use_averaged)
        if p.ndim == 1:  # one sample: no candidate set to normalise over
            d_p = dev_f64(p.reshape(1, -1), f"cuda:{dev}")
            return float(self.fuse_device(d_p, w[:1], False).cpu().numpy()[0])
        d_p = dev_f64(np.ascontiguousarray(p.reshape(-1, self._n_signals)), f"cuda:{dev}")
        return np.atleast_1d(self.fuse_device(d_p, w, self._normalize).cpu().numpy())

    def compute_upper_bounds(self, upper_bound_probs, query_features, use_averaged: bool = False) -> np.ndarray:
        """Fused upper bound per candidate from per-signal upper bounds (Theorem 8.7.1; fusion.py:1039-1082)."""
        dev = _lib.require_cuda()
        ub = np.asarray(upper_bound_probs, dtype=np.float64)
        if ub.ndim == 1:
            ub = ub.reshape(1, -1)
        w = self._weights_device(query_features, use_averaged)
        d_ub = dev_f64(np.ascontiguousarray(ub), f"cuda:{dev}")
        return np.atleast_1d(self.fuse_device(d_ub, w, self._normalize).cpu().numpy())

    def prune(self, probs, query_features, threshold: float, upper_bound_probs=None, use_averaged: bool = False):
        """Drop candidates whose fused upper bound is below `threshold` (fusion.py:1084-1135):
        (surviving indices, their fused probabilities)."""
        p = np.asarray(probs, dtype=np.float64)
        qf = np.atleast_2d(np.asarray(query_features, dtype=np.float64))
        if p.ndim == 1:
            p = p.reshape(1, -1)
        if upper_bound_probs is None:
            upper_bound_probs = p
        ub = self.compute_upper_bounds(upper_bound_probs, qf, use_averaged)
        idx = np.where(ub >= threshold)[0]
        if len(idx) == 0:
            return idx, np.array([], dtype=np.float64)
        surv_qf = qf[idx] if qf.shape[0] > 1 else qf
        return idx, np.atleast_1d(self(p[idx], surv_qf, use_averaged))
