"""bayesian_bm25_b200 -- B200-native query-time hot path of Bayesian BM25.

Drop-in names for the part of cognica-io/bayesian-bm25 this package covers
(SURVEY.md section 8): BayesianBM25Scorer, BlockMaxIndex, MultiFieldScorer,
BayesianProbabilityTransform, cosine_to_probability, log_odds_conjunction.
All compute runs in hand-written CUDA kernels behind libbb25.so
(include/bb25.h); there is no CPU fallback.
"""
from .debug import BM25SignalTrace
from .fusion import (AttentionLogOddsWeights, balanced_log_odds_fusion, cosine_to_probability,
                     log_odds_conjunction)
from .multi_field import MultiFieldScorer
from .probability import BayesianProbabilityTransform, logit, sigmoid
from .scorer import BayesianBM25Scorer, BlockMaxIndex, RetrievalResult

__version__ = "0.1.0"

__all__ = [
    "AttentionLogOddsWeights",
    "BM25SignalTrace",
    "BayesianBM25Scorer",
    "BayesianProbabilityTransform",
    "BlockMaxIndex",
    "MultiFieldScorer",
    "RetrievalResult",
    "balanced_log_odds_fusion",
    "cosine_to_probability",
    "log_odds_conjunction",
    "logit",
    "sigmoid",
]
