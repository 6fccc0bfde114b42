"""ctypes front for oracle/bb25_oracle.c (TEST INFRASTRUCTURE ONLY).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never from the shipped package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbb25_oracle.so")

GATING = {"none": 0, "relu": 1, "swish": 2, "gelu": 3, "softplus": 4}


class Params(C.Structure):
    _fields_ = [
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("has_base_rate", C.c_int),
        ("base_rate", C.c_double),
        ("prior_mode", C.c_int),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bb25_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call(
            [cc, "-O2", "-pthread", "-fPIC", "-shared", "-fno-fast-math",
             "-ffp-contract=off", "-o", _SO, src, "-lm"]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_sigmoid.restype = C.c_double
        _lib.orc_sigmoid.argtypes = [C.c_double]
        _lib.orc_logit.restype = C.c_double
        _lib.orc_logit.argtypes = [C.c_double]
        _lib.orc_posterior.restype = C.c_double
        _lib.orc_posterior.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.orc_composite_prior.restype = C.c_double
        _lib.orc_composite_prior.argtypes = [C.c_double, C.c_double]
        _lib.orc_tf_prior.restype = C.c_double
        _lib.orc_tf_prior.argtypes = [C.c_double]
        _lib.orc_norm_prior.restype = C.c_double
        _lib.orc_norm_prior.argtypes = [C.c_double]
        _lib.orc_retrieve_batch.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def make_params(alpha, beta, base_rate=None, prior_mode=0) -> Params:
    return Params(float(alpha), float(beta), int(base_rate is not None),
                  float(base_rate or 0.0), int(prior_mode))


def score_to_probability(alpha, beta, base_rate, score, tf, ratio, prior_mode=0, prior=None):
    s, tf, r = np.broadcast_arrays(_f64(score), _f64(tf), _f64(ratio))
    s, tf, r = _f64(s).ravel(), _f64(tf).ravel(), _f64(r).ravel()
    pr = _f64(np.broadcast_to(_f64(prior), s.shape)).ravel() if prior is not None else s
    out = np.empty_like(s)
    lib().orc_score_to_probability(
        C.c_double(alpha), C.c_double(beta), C.c_int(base_rate is not None),
        C.c_double(base_rate or 0.0), C.c_int(prior_mode), _p(s, C.c_double),
        _p(tf, C.c_double), _p(r, C.c_double), _p(pr, C.c_double),
        C.c_int64(s.size), _p(out, C.c_double))
    return out


def wand_upper_bound(alpha, beta, base_rate, ub, p_max=0.9):
    u = _f64(ub).ravel()
    out = np.empty_like(u)
    lib().orc_wand_upper_bound(
        C.c_double(alpha), C.c_double(beta), C.c_int(base_rate is not None),
        C.c_double(base_rate or 0.0), C.c_double(p_max), _p(u, C.c_double),
        C.c_int64(u.size), _p(out, C.c_double))
    return out


def cosine_to_probability(c):
    c = _f64(c).ravel()
    out = np.empty_like(c)
    lib().orc_cosine_to_probability(_p(c, C.c_double), C.c_int64(c.size), _p(out, C.c_double))
    return out


def log_odds_conjunction(probs, alpha=None, weights=None, gating="none",
                         gating_beta=1.0, max_logit=None):
    """fusion.py:172-280 with alpha resolution (:106-116, :260, :270)."""
    p = _f64(probs)
    n = p.shape[-1]
    p2 = p.reshape(-1, n)
    if alpha == "auto":
        a = 0.5
    elif alpha is None:
        a = 0.0 if weights is not None else 0.5
    else:
        a = float(alpha)
    scale = float(n ** a)
    out = np.empty(p2.shape[0], dtype=np.float64)
    w = _f64(weights) if weights is not None else None
    lib().orc_log_odds_conjunction(
        _p(p2, C.c_double), C.c_int64(p2.shape[0]), C.c_int(n),
        _p(w, C.c_double) if w is not None else None, C.c_double(scale),
        C.c_int(GATING[gating]), C.c_double(gating_beta),
        C.c_int(max_logit is not None), C.c_double(max_logit or 0.0),
        _p(out, C.c_double))
    return out.reshape(p.shape[:-1])


def _csc(scores):
    return (np.ascontiguousarray(scores["data"], dtype=np.float32),
            np.ascontiguousarray(scores["indices"], dtype=np.int32),
            np.ascontiguousarray(scores["indptr"], dtype=np.int64))


def get_scores(scores: dict, term_ids) -> np.ndarray:
    data, indices, indptr = _csc(scores)
    q = np.ascontiguousarray(term_ids, dtype=np.int32)
    out = np.empty(scores["num_docs"], dtype=np.float32)
    lib().orc_get_scores(_p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64),
                         C.c_int64(scores["num_docs"]), _p(q, C.c_int32), C.c_int(q.size),
                         _p(out, C.c_float))
    return out


def match_counts(scores: dict, term_ids) -> np.ndarray:
    _, indices, indptr = _csc(scores)
    q = np.ascontiguousarray(term_ids, dtype=np.int32)
    out = np.empty(scores["num_docs"], dtype=np.int32)
    lib().orc_match_counts(_p(indices, C.c_int32), _p(indptr, C.c_int64),
                           C.c_int64(scores["num_docs"]), _p(q, C.c_int32), C.c_int(q.size),
                           _p(out, C.c_int32))
    return out


def get_probabilities(scores: dict, params: Params, term_ids) -> np.ndarray:
    data, indices, indptr = _csc(scores)
    dl = np.ascontiguousarray(scores["doc_len"], dtype=np.int32)
    q = np.ascontiguousarray(term_ids, dtype=np.int32)
    out = np.empty(scores["num_docs"], dtype=np.float64)
    lib().orc_get_probabilities(
        _p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64),
        C.c_int64(scores["num_docs"]), _p(dl, C.c_int32), C.c_double(scores["avgdl"]),
        C.byref(params), _p(q, C.c_int32), C.c_int(q.size), _p(out, C.c_double))
    return out


def topk_f32(values, k):
    v = np.ascontiguousarray(values, dtype=np.float32)
    ids = np.empty(k, dtype=np.int64)
    sc = np.empty(k, dtype=np.float32)
    lib().orc_topk_f32(_p(v, C.c_float), C.c_int64(v.size), C.c_int(k),
                       _p(ids, C.c_int64), _p(sc, C.c_float))
    return ids, sc


def topk_f64(values, k):
    v = _f64(values)
    k = min(k, v.size)
    ids = np.empty(k, dtype=np.int64)
    out = np.empty(k, dtype=np.float64)
    lib().orc_topk_f64(_p(v, C.c_double), C.c_int64(v.size), C.c_int(k),
                       _p(ids, C.c_int64), _p(out, C.c_double))
    return ids, out


def retrieve_batch(scores: dict, params: Params, q_terms, q_off, k, n_threads=0):
    """Returns (ids[Q,k] int64, scores[Q,k] f32, probs[Q,k] f64, threads_used)."""
    data, indices, indptr = _csc(scores)
    dl = np.ascontiguousarray(scores["doc_len"], dtype=np.int32)
    qt = np.ascontiguousarray(q_terms, dtype=np.int32)
    qo = np.ascontiguousarray(q_off, dtype=np.int64)
    nq = qo.size - 1
    ids = np.empty((nq, k), dtype=np.int64)
    sc = np.empty((nq, k), dtype=np.float32)
    pr = np.empty((nq, k), dtype=np.float64)
    used = lib().orc_retrieve_batch(
        _p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64),
        C.c_int64(scores["num_docs"]), _p(dl, C.c_int32), C.c_double(scores["avgdl"]),
        C.byref(params), _p(qt, C.c_int32), _p(qo, C.c_int64), C.c_int64(nq), C.c_int(k),
        C.c_int(n_threads), _p(ids, C.c_int64), _p(sc, C.c_float), _p(pr, C.c_double))
    return ids, sc, pr, used


def blockmax_dense(score_matrix, block_size):
    sm = _f64(score_matrix)
    nt, nd = sm.shape
    nb = (nd + block_size - 1) // block_size
    out = np.empty((nt, nb), dtype=np.float64)
    lib().orc_blockmax_dense(_p(sm, C.c_double), C.c_int64(nt), C.c_int64(nd),
                             C.c_int(block_size), _p(out, C.c_double))
    return out


def blockmax_csc(scores: dict, terms, block_size):
    data, indices, indptr = _csc(scores)
    t = np.ascontiguousarray(terms, dtype=np.int32)
    nb = (scores["num_docs"] + block_size - 1) // block_size
    out = np.empty((t.size, nb), dtype=np.float32)
    lib().orc_blockmax_csc(_p(data, C.c_float), _p(indices, C.c_int32), _p(indptr, C.c_int64),
                           C.c_int64(scores["num_docs"]), _p(t, C.c_int32), C.c_int(t.size),
                           C.c_int(block_size), _p(out, C.c_float))
    return out


def merge_topk(ids, scores, probs):
    """[S,Q,k] x3 -> global [Q,k] x3."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    probs = _f64(probs)
    s, q, k = ids.shape
    oi = np.empty((q, k), dtype=np.int64)
    os_ = np.empty((q, k), dtype=np.float32)
    op = np.empty((q, k), dtype=np.float64)
    lib().orc_merge_topk(_p(ids, C.c_int64), _p(scores, C.c_float), _p(probs, C.c_double),
                         C.c_int(s), C.c_int64(q), C.c_int(k), _p(oi, C.c_int64),
                         _p(os_, C.c_float), _p(op, C.c_double))
    return oi, os_, op


def max_threads() -> int:
    return int(lib().orc_max_threads())
