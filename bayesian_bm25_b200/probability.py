"""BayesianProbabilityTransform: inference methods of the reference's
``bayesian_bm25/probability.py`` (:20-236), evaluated on the GPU through libbb25.

Same names, argument meaning, return conventions (scalar in -> Python float,
array in -> ndarray) and errors as the reference.  Parameter learning
(``fit`` / ``update``, probability.py:238-666) is out of scope for this package
(component 8) and raises NotImplementedError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._device import elementwise

_EPSILON = 1e-10


def _ret(res, scalar):
    return float(res) if scalar else res


def sigmoid(x):
    """Numerically stable sigmoid (probability.py:29-41)."""
    return _ret(*elementwise("bb25_sigmoid", [x]))


def logit(p):
    """log(p / (1 - p)) after clamping to [1e-10, 1 - 1e-10] (probability.py:44-48)."""
    return _ret(*elementwise("bb25_logit", [p]))


class BayesianProbabilityTransform:
    """Transforms raw BM25 scores into calibrated probabilities (probability.py:51-236)."""

    _VALID_MODES = ("balanced", "prior_aware", "prior_free")

    def __init__(self, alpha: float = 1.0, beta: float = 0.0, base_rate: float | None = None,
                 prior_fn=None) -> None:
        if base_rate is not None and not (0.0 < base_rate < 1.0):
            raise ValueError(f"base_rate must be in (0, 1), got {base_rate}")
        self.alpha = alpha
        self.beta = beta
        self.base_rate = base_rate
        self._prior_fn = prior_fn
        self._training_mode = "balanced"

    # -- kept for drop-in attribute compatibility; no learning happens here --------
    @property
    def averaged_alpha(self) -> float:
        return self.alpha

    @property
    def averaged_beta(self) -> float:
        return self.beta

    def fit(self, *args, **kwargs):
        raise NotImplementedError("parameter learning is outside the B200 hot path; use the reference's fit()")

    def update(self, *args, **kwargs):
        raise NotImplementedError("parameter learning is outside the B200 hot path; use the reference's update()")

    def _params(self) -> _lib.Params:
        return _lib.make_params(self.alpha, self.beta, self.base_rate,
                                prior_free=self._training_mode == "prior_free")

    def likelihood(self, score):
        """sigma(alpha * (score - beta)) (probability.py:106-108)."""
        return _ret(*elementwise("bb25_likelihood", [score], extra_pre=(C.byref(self._params()),)))

    @staticmethod
    def tf_prior(tf):
        """0.2 + 0.7 * min(1, tf / 10) (probability.py:110-115)."""
        return _ret(*elementwise("bb25_tf_prior", [tf]))

    @staticmethod
    def norm_prior(doc_len_ratio):
        """0.3 + 0.6 * (1 - min(1, |r - 0.5| * 2)) (probability.py:117-129)."""
        return _ret(*elementwise("bb25_norm_prior", [doc_len_ratio]))

    @staticmethod
    def composite_prior(tf, doc_len_ratio):
        """clamp(0.7 * P_tf + 0.3 * P_norm, 0.1, 0.9) (probability.py:131-140)."""
        return _ret(*elementwise("bb25_composite_prior", [tf, doc_len_ratio]))

    @staticmethod
    def posterior(likelihood_val, prior, base_rate: float | None = None):
        """Two-step Bayes update (probability.py:142-169)."""
        return _ret(*elementwise(
            "bb25_posterior", [likelihood_val, prior],
            extra_post=(int(base_rate is not None), float(base_rate) if base_rate is not None else 0.0)))

    def score_to_probability(self, score, tf, doc_len_ratio):
        """BM25 score -> calibrated probability (probability.py:171-203)."""
        p = C.byref(self._params())
        if self._training_mode != "prior_free" and self._prior_fn is not None:
            # custom prior: the user's Python callback runs on the host, the Bayes
            # update on the device (probability.py:194-199)
            prior = np.asarray(self._prior_fn(score, tf, doc_len_ratio), dtype=np.float64)
            res, scalar = elementwise("bb25_score_to_probability", [score, tf, doc_len_ratio, prior],
                                      extra_pre=(p,))
            return _ret(res, scalar)
        dev = _lib.require_cuda()
        arrs = [np.asarray(a, dtype=np.float64) for a in (score, tf, doc_len_ratio)]
        scalar = all(a.ndim == 0 for a in arrs)
        bc = np.broadcast_arrays(*arrs)
        shape = bc[0].shape
        import torch

        d = [torch.from_numpy(np.ascontiguousarray(a).ravel()).to(f"cuda:{dev}") for a in bc]
        out = torch.empty_like(d[0])
        _lib.check(_lib.lib().bb25_score_to_probability(
            dev, p, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), None, d[0].numel(),
            out.data_ptr(), _lib.stream_ptr()))
        return _ret(out.cpu().numpy().reshape(shape), scalar)

    def wand_upper_bound(self, bm25_upper_bound, p_max: float = 0.9):
        """posterior(sigma(alpha * (ub - beta)), p_max [, base_rate]) (probability.py:205-236)."""
        return _ret(*elementwise("bb25_wand_upper_bound", [bm25_upper_bound],
                                 extra_pre=(C.byref(self._params()),), extra_post=(float(p_max),)))
