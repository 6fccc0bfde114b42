"""cosine_to_probability and log_odds_conjunction of the reference's
``bayesian_bm25/fusion.py`` (:25-45, :103-280), evaluated on the GPU.

The other fusion helpers of the reference (prob_and/or/not, balanced fusion, the
learnable weight classes) are outside the hot path (components 9-10).
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
