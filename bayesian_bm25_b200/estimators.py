"""Host-side estimators for alpha, beta and the corpus base rate.

These run once per ``index()`` on a few thousand pseudo-query scores and stay in
NumPy on the host, as SURVEY 2 (component 7) prescribes: they are fed by the GPU
``get_scores`` and their outputs are constants of the hot path.  Semantics follow
``bayesian_bm25/scorer.py:313-467``; note that the pooled scores are float32, so
``np.median`` / ``np.std`` are taken in float32 exactly as the reference does.
"""
from __future__ import annotations

import numpy as np

VALID_BASE_RATE_METHODS = ("percentile", "mixture", "elbow")
_LO, _HI = 1e-6, 0.5


def _bounded(x: float) -> float:
    return float(min(max(x, _LO), _HI))


def sigmoid_parameters(per_query_scores, user_alpha, user_beta):
    """beta = median(pooled scores), alpha = 1 / std (scorer.py:313-337)."""
    if user_alpha is not None and user_beta is not None:
        return user_alpha, user_beta
    if not per_query_scores:
        return (user_alpha or 1.0, user_beta or 0.0)
    pooled = np.concatenate(per_query_scores)
    spread = float(np.std(pooled))
    alpha_hat = 1.0 / spread if spread > 0 else 1.0
    beta_hat = float(np.median(pooled))
    return (alpha_hat if user_alpha is None else user_alpha,
            beta_hat if user_beta is None else user_beta)


def base_rate_percentile(per_query_scores, n_docs: int) -> float:
    """Mean over pseudo-queries of |{s >= P95(s)}| / N (scorer.py:366-378)."""
    fractions = [int(np.sum(s >= float(np.percentile(s, 95)))) / n_docs for s in per_query_scores]
    return _bounded(float(np.mean(fractions)))


def base_rate_mixture(per_query_scores) -> float:
    """Weight of the higher-mean component of a 2-Gaussian EM fit, 20 iterations,
    initialised by a median split (scorer.py:380-433)."""
    x = np.concatenate(per_query_scores)
    if len(x) < 2:
        return _LO
    med = float(np.median(x))
    low, high = x <= med, x > med
    mu = [float(np.mean(x[low])) if low.any() else med - 1.0,
          float(np.mean(x[high])) if high.any() else med + 1.0]
    var = [max(float(np.var(x[low])) if low.any() else 1.0, 1e-8),
           max(float(np.var(x[high])) if high.any() else 1.0, 1e-8)]
    w_hi = 0.5
    for _ in range(20):
        sd0, sd1 = np.sqrt(var[0]), np.sqrt(var[1])
        ll0 = -0.5 * ((x - mu[0]) / sd0) ** 2 - np.log(sd0)
        ll1 = -0.5 * ((x - mu[1]) / sd1) ** 2 - np.log(sd1)
        a0 = np.log(max(1.0 - w_hi, 1e-10)) + ll0
        a1 = np.log(max(w_hi, 1e-10)) + ll1
        resp = np.exp(a1 - np.logaddexp(a0, a1))
        m1 = float(np.sum(resp))
        m0 = float(np.sum(1.0 - resp))
        if m0 < 1e-8 or m1 < 1e-8:
            break
        mu[0] = float(np.sum((1.0 - resp) * x) / m0)
        mu[1] = float(np.sum(resp * x) / m1)
        var[0] = max(float(np.sum((1.0 - resp) * (x - mu[0]) ** 2) / m0), 1e-8)
        var[1] = max(float(np.sum(resp * (x - mu[1]) ** 2) / m1), 1e-8)
        w_hi = m1 / len(x)
    return _bounded(w_hi if mu[1] >= mu[0] else 1.0 - w_hi)


def base_rate_elbow(per_query_scores) -> float:
    """Fraction of pooled scores above the knee of the sorted-score curve (the point
    farthest from the chord between its ends) (scorer.py:435-467)."""
    y = np.sort(np.concatenate(per_query_scores))[::-1]
    n = len(y)
    if n < 3:
        return _LO
    run = float(n - 1)
    rise = float(y[-1] - y[0])
    chord = np.sqrt(run * run + rise * rise)
    if chord < 1e-12:
        return _LO
    t = np.arange(n, dtype=np.float64)
    dist = np.abs(rise * t - run * (y - y[0])) / chord
    knee = int(np.argmax(dist))
    return _bounded(max(1, knee) / n)


def estimate_base_rate(per_query_scores, n_docs: int, method: str) -> float:
    """Dispatch of scorer.py:339-364."""
    if not per_query_scores:
        return _LO
    if method == "percentile":
        return base_rate_percentile(per_query_scores, n_docs)
    if method == "mixture":
        return base_rate_mixture(per_query_scores)
    if method == "elbow":
        return base_rate_elbow(per_query_scores)
    raise ValueError(f"Unknown base_rate_method: {method!r}")
