"""BayesianBM25Scorer and BlockMaxIndex on the B200 path.

Mirrors the public surface of the reference's ``bayesian_bm25/scorer.py``
(constructor kwargs, ``index`` / ``retrieve`` / ``get_probabilities`` /
``add_documents``, properties and error behaviour) with the sparse BM25 engine,
the tf counting, the posterior and the top-k running as CUDA kernels behind
libbb25 (include/bb25.h).  Extensions for corpora that do not fit Python lists:
``index_from_ids`` / ``index_from_csc`` / ``retrieve_ids``.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from itertools import chain, repeat

import numpy as np
import torch

from . import _lib, estimators, index_build
from .probability import BayesianProbabilityTransform

_VALID_BASE_RATE_METHODS = estimators.VALID_BASE_RATE_METHODS
MAX_DEVICE_K = 4096      # bb25_retrieve_batch limit; larger k takes the dense path
MAX_DENSE_K = 8192       # bb25_retrieve_one_dense limit
QUERY_CHUNK = 16384      # queries per bb25_retrieve_batch call (bounds the workspace)
PIPELINE_CHUNKS = 3      # large batches: host calls whose copies / token->id mapping overlap the device work
PIPELINE_MIN_CHUNK = 1024


class BlockMaxIndex:
    """Per-(term, document block) maxima for BMW-style bounds (scorer.py:33-142)."""

    def __init__(self, block_size: int = 128) -> None:
        if block_size < 1:
            raise ValueError(f"block_size must be >= 1, got {block_size}")
        self._block_size = block_size
        self._block_maxes: np.ndarray | None = None
        self._n_docs = 0
        self._n_terms = 0

    def build(self, score_matrix) -> None:
        """score_matrix: (n_terms, n_docs) per-term BM25 contributions (scorer.py:55-81)."""
        sm = np.asarray(score_matrix, dtype=np.float64)
        if sm.ndim != 2:
            raise ValueError(f"score_matrix must be 2D (n_terms, n_docs), got {sm.ndim}D")
        self._n_terms, self._n_docs = sm.shape
        dev = _lib.require_cuda()
        nb = (self._n_docs + self._block_size - 1) // self._block_size
        d_sm = torch.from_numpy(np.ascontiguousarray(sm)).to(f"cuda:{dev}")
        out = torch.empty((self._n_terms, nb), dtype=torch.float64, device=d_sm.device)
        if self._n_terms and self._n_docs:
            _lib.check(_lib.lib().bb25_blockmax_dense(dev, d_sm.data_ptr(), self._n_terms, self._n_docs,
                                                      self._block_size, out.data_ptr(), _lib.stream_ptr()))
        self._block_maxes = out.cpu().numpy()

    def build_from_scorer(self, scorer: "BayesianBM25Scorer", term_ids) -> None:
        """Extension: block maxima of `term_ids` straight from the device CSC (an
        absent posting counts as 0.0; posting values are >= 0), for corpora whose
        dense (n_terms, n_docs) matrix cannot be materialised."""
        scorer._require_index("build_from_scorer()")
        terms = torch.as_tensor(np.asarray(term_ids, dtype=np.int32), device=scorer._device)
        nb = (scorer.num_docs + self._block_size - 1) // self._block_size
        out = torch.empty((terms.numel(), nb), dtype=torch.float32, device=scorer._device)
        _lib.check(_lib.lib().bb25_blockmax_csc(scorer._handle, terms.data_ptr(), terms.numel(),
                                                self._block_size, out.data_ptr(), _lib.stream_ptr()))
        self._n_terms, self._n_docs = terms.numel(), scorer.num_docs
        self._block_maxes = out.cpu().numpy().astype(np.float64)

    def block_upper_bound(self, term_idx: int, block_id: int) -> float:
        if self._block_maxes is None:
            raise RuntimeError("Call build() before block_upper_bound().")
        return float(self._block_maxes[term_idx, block_id])

    def bayesian_block_upper_bound(self, term_idx: int, block_id: int,
                                   transform: BayesianProbabilityTransform, p_max: float = 0.9) -> float:
        """transform.wand_upper_bound of the block maximum (scorer.py:101-130)."""
        return float(transform.wand_upper_bound(self.block_upper_bound(term_idx, block_id), p_max))

    @property
    def block_size(self) -> int:
        return self._block_size

    @property
    def n_blocks(self) -> int:
        if self._block_maxes is None:
            raise RuntimeError("Call build() before accessing n_blocks.")
        return self._block_maxes.shape[1]


@dataclass
class RetrievalResult:
    """Return type of ``retrieve(explain=True)`` in the reference (scorer.py:145-163)."""

    doc_ids: np.ndarray
    probabilities: np.ndarray
    explanations: list | None


class BayesianBM25Scorer:
    """BM25 scorer returning Bayesian-calibrated probabilities (scorer.py:166-640)."""

    def __init__(self, k1: float = 1.2, b: float = 0.75, method: str = "robertson",
                 alpha: float | None = None, beta: float | None = None,
                 base_rate: float | str | None = None, base_rate_method: str = "percentile") -> None:
        if base_rate_method not in _VALID_BASE_RATE_METHODS:
            raise ValueError(
                f"base_rate_method must be one of {_VALID_BASE_RATE_METHODS}, got {base_rate_method!r}")
        if method not in index_build.VALID_METHODS:
            raise ValueError(f"method must be one of {index_build.VALID_METHODS}, got {method!r}")
        self._k1, self._b, self._method = k1, b, method
        self._user_alpha, self._user_beta = alpha, beta
        self._user_base_rate = base_rate
        self._base_rate_method = base_rate_method
        self._transform: BayesianProbabilityTransform | None = None
        self._doc_lengths: np.ndarray | None = None
        self._avgdl: float | None = None
        self._corpus_tokens: list[list[str]] | None = None
        self._vocab: dict[str, int] = {}
        self._handle = None
        self._device = None
        self._num_docs = 0
        self._n_vocab = 0
        self._doc_id_offset = 0

    def __del__(self):
        self._release()

    def _release(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.lib().bb25_index_destroy(h)
            except Exception:
                pass

    # ---- properties (scorer.py:224-248) ------------------------------------------
    @property
    def num_docs(self) -> int:
        return int(self._num_docs)

    @property
    def doc_lengths(self) -> np.ndarray:
        if self._doc_lengths is None:
            raise RuntimeError("Call index() before accessing doc_lengths.")
        return self._doc_lengths

    @property
    def avgdl(self) -> float:
        if self._avgdl is None:
            raise RuntimeError("Call index() before accessing avgdl.")
        return self._avgdl

    @property
    def base_rate(self) -> float | None:
        return None if self._transform is None else self._transform.base_rate

    @property
    def transform(self) -> BayesianProbabilityTransform | None:
        return self._transform

    def _require_index(self, what: str) -> None:
        if self._transform is None or self._handle is None:
            raise RuntimeError(f"Call index() before {what}.")

    # ---- indexing -------------------------------------------------------------------
    def index(self, corpus_tokens: list[list[str]], show_progress: bool = True) -> None:
        """Build the index from tokenised documents (scorer.py:250-285)."""
        vocab: dict[str, int] = {}
        flat: list[int] = []
        offs = np.zeros(len(corpus_tokens) + 1, dtype=np.int64)
        for i, doc in enumerate(corpus_tokens):
            for tok in doc:
                j = vocab.get(tok)
                if j is None:
                    j = vocab[tok] = len(vocab)
                flat.append(j)
            offs[i + 1] = len(flat)
        if "" not in vocab:  # bm25s reserves an id for the empty token
            vocab[""] = len(vocab)
        self._corpus_tokens = corpus_tokens
        self._vocab = vocab
        self.index_from_ids(np.asarray(flat, dtype=np.int32), offs, len(vocab))

    def index_from_ids(self, token_ids, doc_offsets, n_vocab: int) -> None:
        """Extension: index a corpus given as a flat token-id array plus document
        offsets (no Python lists / per-document sets)."""
        dev = _lib.require_cuda()
        device = torch.device(f"cuda:{dev}")
        t = torch.as_tensor(np.asarray(token_ids), device=device)
        o = torch.as_tensor(np.asarray(doc_offsets, dtype=np.int64), device=device)
        if o.numel() < 2:
            raise ValueError("cannot index an empty corpus")
        csc = index_build.build_csc(t, o, int(n_vocab), self._k1, self._b, self._method)
        offs = np.asarray(doc_offsets, dtype=np.int64)
        tok = np.asarray(token_ids)
        sample = [tok[offs[i]:offs[i + 1]][:5] for i in self._pseudo_query_docs(len(offs) - 1)]
        self.index_from_csc(csc, pseudo_queries=sample)

    def index_from_csc(self, csc: dict, pseudo_queries=None) -> None:
        """Extension: adopt a prebuilt CSC (dict of tensors as produced by
        index_build / synthetic) and estimate alpha / beta / base rate from
        `pseudo_queries` (lists of term ids; default: none -> user values or 1, 0)."""
        dev = _lib.require_cuda()
        device = torch.device(f"cuda:{dev}")
        self._release()
        data = csc["data"].to(device, torch.float32).contiguous()
        indices = csc["indices"].to(device, torch.int32).contiguous()
        indptr = csc["indptr"].to(device, torch.int64).contiguous()
        doc_len = csc["doc_len"].to(device, torch.int32).contiguous()
        n_docs = int(csc["num_docs"])
        n_vocab = indptr.numel() - 1
        if not (csc["avgdl"] > 0):
            raise ValueError("corpus has no tokens")
        handle = C.c_void_p()
        _lib.check(_lib.lib().bb25_index_create(
            dev, n_docs, n_vocab, data.numel(), data.data_ptr(), indices.data_ptr(), indptr.data_ptr(),
            doc_len.data_ptr(), float(csc["avgdl"]), int(csc.get("doc_id_offset", 0)), C.byref(handle)))
        self._handle, self._device = handle, device
        self._num_docs, self._n_vocab = n_docs, n_vocab
        self._doc_id_offset = int(csc.get("doc_id_offset", 0))
        self._doc_lengths = doc_len.cpu().numpy().astype(np.float64)
        self._avgdl = float(csc["avgdl"])
        self._df = (indptr[1:] - indptr[:-1]).cpu().numpy()

        per_query = []
        for q in (pseudo_queries or []):
            if len(q) == 0:
                continue
            s = self._scores_device(np.asarray(q, dtype=np.int32))
            nz = s[s > 0]
            if nz.numel() > 0:
                per_query.append(nz.cpu().numpy())
        alpha, beta = estimators.sigmoid_parameters(per_query, self._user_alpha, self._user_beta)
        base_rate = None
        if self._user_base_rate == "auto":
            base_rate = estimators.estimate_base_rate(per_query, n_docs, self._base_rate_method)
        elif isinstance(self._user_base_rate, (int, float)):
            base_rate = float(self._user_base_rate)
        self._transform = BayesianProbabilityTransform(alpha=alpha, beta=beta, base_rate=base_rate)

    def set_vocabulary(self, tokens) -> None:
        """Extension: token strings of a prebuilt CSC's term ids (tokens[i] is term i), so that the
        string-level retrieve()/get_probabilities() work on an index adopted with index_from_csc()."""
        self._vocab = {t: i for i, t in enumerate(tokens)}

    @staticmethod
    def _pseudo_query_docs(n: int) -> np.ndarray:
        """Documents whose first five tokens serve as pseudo-queries (scorer.py:295-298)."""
        return np.random.default_rng(42).choice(n, size=min(n, 50), replace=False)

    def add_documents(self, new_corpus_tokens: list[list[str]], show_progress: bool = True) -> None:
        """Append documents and rebuild (scorer.py:469-492)."""
        if self._corpus_tokens is None:
            raise RuntimeError("Call index() before add_documents().")
        self.index(self._corpus_tokens + new_corpus_tokens, show_progress=show_progress)

    # ---- query side ------------------------------------------------------------------
    def _term_ids(self, tokens) -> np.ndarray:
        v = self._vocab
        return np.asarray([v[t] for t in tokens if t in v], dtype=np.int32)

    def _term_ids_batch(self, query_tokens):
        """Token lists -> (flat in-vocabulary ids int32, offsets int64[Q+1]); out-of-vocabulary
        tokens are dropped, order and duplicates kept (what bm25s does with a query)."""
        nq = len(query_tokens)
        lens = np.fromiter(map(len, query_tokens), dtype=np.int64, count=nq)
        total = int(lens.sum())
        ids = np.fromiter(map(self._vocab.get, chain.from_iterable(query_tokens), repeat(-1)), dtype=np.int64,
                          count=total)
        keep = ids >= 0
        off = np.zeros(nq + 1, dtype=np.int64)
        if total:
            ends = np.cumsum(lens)
            kept = np.concatenate(([0], np.cumsum(keep)))
            off[1:] = kept[ends]
        return ids[keep].astype(np.int32), off

    def _params(self) -> _lib.Params:
        t = self._transform
        if t._prior_fn is not None and t._training_mode != "prior_free":
            # probability.py:194-199 evaluates a Python callback per document; the fused kernels
            # cannot, and silently ignoring it would change the numbers
            raise NotImplementedError(
                "the transform has a custom prior_fn: the fused retrieve/get_probabilities kernels evaluate the "
                "composite prior only; use get_scores_ids() + transform.score_to_probability() for a custom prior")
        return _lib.make_params(t.alpha, t.beta, t.base_rate, prior_free=t._training_mode == "prior_free")

    def _scores_device(self, term_ids: np.ndarray) -> torch.Tensor:
        out = torch.empty(self._num_docs, dtype=torch.float32, device=self._device)
        q = np.ascontiguousarray(term_ids, dtype=np.int32)
        _lib.check(_lib.lib().bb25_get_scores(self._handle, q.ctypes.data, q.size, out.data_ptr(),
                                              _lib.stream_ptr()))
        return out

    def get_scores_ids(self, term_ids) -> np.ndarray:
        """fp32 BM25 scores of all documents for in-vocabulary term ids (what
        ``self._bm25.get_scores`` is for the reference, scorer.py:306,583)."""
        self._require_index("get_scores_ids()")
        return self._scores_device(np.asarray(term_ids, dtype=np.int32)).cpu().numpy()

    def probabilities_device(self, term_ids, out: torch.Tensor | None = None, stride: int = 1) -> torch.Tensor:
        """Dense fp64 probabilities on the device; `out` may be a column of a wider
        buffer (element d written at out.data_ptr() + d*stride)."""
        self._require_index("get_probabilities()")
        q = np.ascontiguousarray(term_ids, dtype=np.int32)
        if out is None:
            out = torch.empty(self._num_docs, dtype=torch.float64, device=self._device)
            stride = 1
        p = self._params()
        _lib.check(_lib.lib().bb25_get_probabilities(self._handle, C.byref(p), q.ctypes.data, q.size,
                                                     out.data_ptr(), stride, _lib.stream_ptr()))
        return out

    def fuse_signal_device(self, term_ids, weight: float, n_signals: int, scale: float, flags: int,
                           acc: torch.Tensor) -> torch.Tensor:
        """This index as ONE signal of a log-odds conjunction, fused into the traversal
        pass: acc[d] = [acc[d] +] weight * logit(clamp(P(d | query))) and, on the last
        signal (flags & 2), sigmoid(scale * acc[d]).  flags & 1 = first signal,
        flags & 4 = unweighted-mean branch (see include/bb25.h)."""
        self._require_index("fuse_signal_device()")
        q = np.ascontiguousarray(term_ids, dtype=np.int32)
        p = self._params()
        _lib.check(_lib.lib().bb25_fuse_bm25_signal(self._handle, C.byref(p), q.ctypes.data, q.size, float(weight),
                                                    int(n_signals), float(scale), int(flags), acc.data_ptr(),
                                                    _lib.stream_ptr()))
        return acc

    def get_probabilities(self, query_tokens: list[str]) -> np.ndarray:
        """Probabilities for ALL documents, 0.0 where the score is <= 0 (scorer.py:564-590)."""
        if self._transform is None:
            raise RuntimeError("Call index() before get_probabilities().")
        return self.probabilities_device(self._term_ids(query_tokens)).cpu().numpy()

    def retrieve_ids_device(self, q_terms: torch.Tensor, q_off: torch.Tensor, k: int, host_off=None):
        """Device-resident batch retrieve: q_terms int32 [total], q_off int64 [Q+1]
        CUDA tensors -> (ids int64 [Q,k], scores fp32 [Q,k], probs fp64 [Q,k]) CUDA tensors.
        `host_off` (the same offsets on the host) lets the library enqueue the whole batch
        without reading anything back first."""
        self._require_index("retrieve()")
        nq = q_off.numel() - 1
        ids = torch.empty((nq, k), dtype=torch.int64, device=self._device)
        sc = torch.empty((nq, k), dtype=torch.float32, device=self._device)
        pr = torch.empty((nq, k), dtype=torch.float64, device=self._device)
        p = self._params()
        for s in range(0, nq, QUERY_CHUNK):
            e = min(nq, s + QUERY_CHUNK)
            if host_off is not None:
                _lib.check(_lib.lib().bb25_retrieve_batch_ex(
                    self._handle, C.byref(p), q_terms.data_ptr(), q_off[s:].data_ptr(), e - s, int(host_off[s]),
                    int(host_off[e] - host_off[s]), k, ids[s:].data_ptr(), sc[s:].data_ptr(), pr[s:].data_ptr(),
                    _lib.stream_ptr()))
            else:
                _lib.check(_lib.lib().bb25_retrieve_batch(
                    self._handle, C.byref(p), q_terms.data_ptr(), q_off[s:].data_ptr(), e - s, k,
                    ids[s:].data_ptr(), sc[s:].data_ptr(), pr[s:].data_ptr(), _lib.stream_ptr()))
        return ids, sc, pr

    def retrieve_ids(self, q_terms, q_off, k: int = 10, return_scores: bool = False):
        """Extension: batch retrieve for queries given as in-vocabulary term ids
        (flat int32 array + int64 offsets), host in / host out, through the C ABI's
        host-buffer entry point (bb25_retrieve_batch_host) with page-locked buffers."""
        self._require_index("retrieve()")
        if k > self._num_docs:
            raise ValueError(
                f"k of {k} is larger than the number of available scores, which is {self._num_docs}")
        q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        nq = q_off.size - 1
        if k > MAX_DEVICE_K:
            return self._retrieve_large_k(q_terms, q_off, k, return_scores)
        h_ids = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
        h_pr = torch.empty((nq, k), dtype=torch.float64, pin_memory=True)
        h_sc = torch.empty((nq, k), dtype=torch.float32, pin_memory=True) if return_scores else None
        bounds = self._chunk_bounds(nq)
        if len(bounds) <= 2:
            self._host_call(*self._stage_host(q_terms, q_off), k, h_ids, h_sc, h_pr)
        else:
            # two calls in flight: while one chunk's results travel to the host (the library's second
            # staging slot and copy stream), the next chunk is already being traversed
            with ThreadPoolExecutor(max_workers=2) as pool:
                pending = []
                for s, e in zip(bounds[:-1], bounds[1:]):
                    staged = self._stage_host(q_terms[q_off[s]:q_off[e]], q_off[s:e + 1] - q_off[s])
                    if len(pending) == 2:
                        pending.pop(0).result()
                    pending.append(pool.submit(self._host_call, *staged, k, h_ids[s:e],
                                               h_sc[s:e] if return_scores else None, h_pr[s:e]))
                for f in pending:
                    f.result()
        if return_scores:
            return h_ids.numpy(), h_sc.numpy(), h_pr.numpy()
        return h_ids.numpy(), h_pr.numpy()

    @staticmethod
    def _chunk_bounds(nq: int, first: int = 0) -> list:
        """Query ranges of one host call each: at most QUERY_CHUNK queries per call (bounds the workspace).
        `first` > 0 (the string API, which maps tokens to ids chunk by chunk): a small first chunk, whose
        mapping is the only one that cannot hide behind device work, then two or more equal chunks."""
        if first <= 0 or nq < 2 * first:
            n_chunks = max(1, -(-nq // QUERY_CHUNK))
            step = -(-nq // n_chunks) if nq else 1
            return [0] + [min(nq, (i + 1) * step) for i in range(n_chunks)]
        rest = nq - first
        n_chunks = max(PIPELINE_CHUNKS - 1, -(-rest // QUERY_CHUNK))
        step = max(first, -(-rest // n_chunks))
        bounds = [0, first]
        while bounds[-1] < nq:
            bounds.append(min(nq, bounds[-1] + step))
        return bounds

    @staticmethod
    def _stage_host(q_terms: np.ndarray, q_off: np.ndarray):
        """Page-locked copies of one chunk's queries (torch's caching host allocator recycles
        the blocks from call to call)."""
        hp = torch.empty(max(q_terms.size, 1), dtype=torch.int32, pin_memory=True)
        hp.numpy()[:q_terms.size] = q_terms
        ho = torch.empty(q_off.size, dtype=torch.int64, pin_memory=True)
        ho.numpy()[:] = q_off
        return hp, ho

    def _host_call(self, hp, ho, k, h_ids, h_sc, h_pr):
        """bb25_retrieve_batch_host on page-locked buffers (releases the GIL while it runs)."""
        p = self._params()
        _lib.check(_lib.lib().bb25_retrieve_batch_host(
            self._handle, C.byref(p), hp.data_ptr(), ho.data_ptr(), ho.numel() - 1, k, h_ids.data_ptr(),
            h_sc.data_ptr() if h_sc is not None else None, h_pr.data_ptr()))

    def _retrieve_large_k(self, q_terms, q_off, k, return_scores):
        """k beyond the candidate kernels' limit: the dense guaranteed path of the library, one
        query at a time (dense scores, exact top-k of the whole vector, dense posterior); beyond
        its own limit a full device sort of (score desc, id asc) keys."""
        nq = len(q_off) - 1
        ids = np.empty((nq, k), dtype=np.int64)
        sc = np.empty((nq, k), dtype=np.float32)
        pr = np.empty((nq, k), dtype=np.float64)
        if k <= MAX_DENSE_K:
            d_ids = torch.empty((nq, k), dtype=torch.int64, device=self._device)
            d_sc = torch.empty((nq, k), dtype=torch.float32, device=self._device)
            d_pr = torch.empty((nq, k), dtype=torch.float64, device=self._device)
            p = self._params()
            for i in range(nq):
                t = np.ascontiguousarray(q_terms[q_off[i]:q_off[i + 1]], dtype=np.int32)
                _lib.check(_lib.lib().bb25_retrieve_one_dense(
                    self._handle, C.byref(p), t.ctypes.data, t.size, k, d_ids[i].data_ptr(), d_sc[i].data_ptr(),
                    d_pr[i].data_ptr(), _lib.stream_ptr()))
            ids, sc, pr = d_ids.cpu().numpy(), d_sc.cpu().numpy(), d_pr.cpu().numpy()
            return (ids, sc, pr) if return_scores else (ids, pr)
        inv = (2 ** 31 - 1) - torch.arange(self._num_docs, device=self._device, dtype=torch.int64)
        for i in range(nq):
            t = q_terms[q_off[i]:q_off[i + 1]]
            s = self._scores_device(t)
            p = self.probabilities_device(t)
            key = (s.view(torch.int32).to(torch.int64) << 31) | inv
            top = torch.sort(key, descending=True).indices[:k]
            ids[i] = (top + self._doc_id_offset).cpu().numpy()
            sc[i] = s[top].cpu().numpy()
            pr[i] = p[top].cpu().numpy()
        return (ids, sc, pr) if return_scores else (ids, pr)

    def retrieve(self, query_tokens: list[list[str]], k: int = 10, show_progress: bool = False,
                 explain: bool = False):
        """Top-k documents per query with Bayesian probabilities (scorer.py:494-562).
        Ranking is by fp32 BM25 score (ties: ascending doc id); probabilities are
        attached in rank order.  Returns (doc_ids [Q,k] int64, probabilities [Q,k] float64)."""
        if self._transform is None:
            raise RuntimeError("Call index() before retrieve().")
        if explain:
            return self._retrieve_explained(query_tokens, k)
        nq = len(query_tokens)
        if nq < 2 * PIPELINE_MIN_CHUNK or k > MAX_DEVICE_K or k > self._num_docs:
            flat, off = self._term_ids_batch(query_tokens)
            return self.retrieve_ids(flat, off, k)
        # Large batches: the token -> id mapping of chunk i+1 (Python dict lookups) runs on this thread while
        # worker threads sit in the C calls for chunks i and i-1 (ctypes drops the GIL; the library computes
        # one while the other's results travel to the host).  Only the small first chunk's mapping is exposed.
        bounds = self._chunk_bounds(nq, first=PIPELINE_MIN_CHUNK)
        h_ids = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
        h_pr = torch.empty((nq, k), dtype=torch.float64, pin_memory=True)
        with ThreadPoolExecutor(max_workers=2) as pool:
            pending = []
            for s, e in zip(bounds[:-1], bounds[1:]):
                staged = self._stage_host(*self._term_ids_batch(query_tokens[s:e]))
                if len(pending) == 2:
                    pending.pop(0).result()
                pending.append(pool.submit(self._host_call, *staged, k, h_ids[s:e], None, h_pr[s:e]))
            for f in pending:
                f.result()
        return h_ids.numpy(), h_pr.numpy()

    def _retrieve_explained(self, query_tokens, k: int) -> "RetrievalResult":
        """retrieve(explain=True) (scorer.py:538-562): per returned document the BM25SignalTrace of
        FusionDebugger.trace_bm25 (debug.py:178-216) -- tf and every intermediate of the posterior computed on
        the device for the (Q, k) result set -- or None where the score is <= 0."""
        from .debug import BM25SignalTrace
        flat, off = self._term_ids_batch(query_tokens)
        ids, scores, probs = self.retrieve_ids(flat, off, k, return_scores=True)
        nq = len(query_tokens)
        dev = self._device
        d_terms = torch.from_numpy(flat if flat.size else np.zeros(1, np.int32)).to(dev)
        d_off = torch.from_numpy(off).to(dev)
        d_ids = torch.from_numpy(np.ascontiguousarray(ids)).to(dev)
        d_tf = torch.empty((nq, k), dtype=torch.int32, device=dev)
        lib = _lib.lib()
        _lib.check(lib.bb25_match_counts(self._handle, d_terms.data_ptr(), d_off.data_ptr(), nq, k, d_ids.data_ptr(),
                                         d_tf.data_ptr(), _lib.stream_ptr()))
        tf = d_tf.cpu().numpy().astype(np.float64)
        ratio = self._doc_lengths[ids - self._doc_id_offset] / self._avgdl
        d_sc = torch.from_numpy(scores.astype(np.float64)).to(dev)
        d_tfd = torch.from_numpy(tf).to(dev)
        d_r = torch.from_numpy(np.ascontiguousarray(ratio)).to(dev)
        out = torch.empty((nq * k, 7), dtype=torch.float64, device=dev)
        p = self._params()
        _lib.check(lib.bb25_trace_bm25(dev.index, C.byref(p), d_sc.data_ptr(), d_tfd.data_ptr(), d_r.data_ptr(), nq * k,
                                       out.data_ptr(), _lib.stream_ptr()))
        tr = out.cpu().numpy().reshape(nq, k, 7)
        t = self._transform
        lbr = float(np.log(t.base_rate / (1.0 - t.base_rate))) if t.base_rate is not None else None
        explanations = []
        for q in range(nq):
            row = []
            for r in range(k):
                if scores[q, r] > 0:
                    v = tr[q, r]
                    row.append(BM25SignalTrace(
                        raw_score=float(scores[q, r]), tf=float(tf[q, r]), doc_len_ratio=float(ratio[q, r]),
                        likelihood=float(v[0]), tf_prior=float(v[1]), norm_prior=float(v[2]), composite_prior=float(v[3]),
                        logit_likelihood=float(v[4]), logit_prior=float(v[5]), logit_base_rate=lbr, posterior=float(v[6]),
                        alpha=t.alpha, beta=t.beta, base_rate=t.base_rate))
                else:
                    row.append(None)
            explanations.append(row)
        return RetrievalResult(doc_ids=ids, probabilities=probs, explanations=explanations)

    def index_info(self) -> dict:
        """Sizes of the device index (bb25_index_info / bb25_index_table_info)."""
        self._require_index("index_info()")
        nd, nv, nnz, db = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        td, nt = C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().bb25_index_info(self._handle, C.byref(nd), C.byref(nv), C.byref(nnz), C.byref(td),
                                              C.byref(nt), C.byref(db)))
        bt, tb = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().bb25_index_table_info(self._handle, C.byref(bt), C.byref(tb)))
        return {"n_docs": nd.value, "n_vocab": nv.value, "nnz": nnz.value, "tile_docs": td.value, "n_tiles": nt.value,
                "device_bytes": db.value, "block_table_bitmap_terms": bt.value, "block_table_bytes": tb.value}

    def stats(self) -> dict:
        """Counters of the last batch retrieve (launches, traversal passes, re-runs)."""
        vals = [C.c_int64() for _ in range(4)]
        _lib.check(_lib.lib().bb25_retrieve_stats(self._handle, *[C.byref(v) for v in vals]))
        out = dict(zip(("launches", "passes", "rerun_queries", "candidates"), (v.value for v in vals)))
        ms, nl = C.c_double(), C.c_int64()
        _lib.check(_lib.lib().bb25_retrieve_timing(self._handle, C.byref(ms), C.byref(nl)))
        out["traverse_ms"], out["traverse_launches"] = ms.value, nl.value
        u, sk, ms_ = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().bb25_retrieve_prune_stats(self._handle, C.byref(u), C.byref(sk), C.byref(ms_)))
        out["units"], out["units_skipped"], out["units_maxscore"] = u.value, sk.value, ms_.value
        rq, wi = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().bb25_retrieve_route_stats(self._handle, C.byref(rq), C.byref(wi)))
        out["routed_queries"], out["candidate_items"] = rq.value, wi.value
        us = C.c_int64()
        _lib.check(_lib.lib().bb25_retrieve_sparse_units(self._handle, C.byref(us)))
        out["units_sparse"] = us.value
        sy, bad, dn = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().bb25_retrieve_sync_stats(self._handle, C.byref(sy), C.byref(bad), C.byref(dn)))
        out["host_syncs"], out["repaired_queries"], out["dense_fallback_queries"] = sy.value, bad.value, dn.value
        return out

    def set_pruning(self, level: int) -> None:
        """Dynamic pruning level of batch retrieve: 0 exhaustive, 1 block-max skip,
        2 + per-block non-essential frequent terms (documents matching only them are not evaluated when
        their summed block maxima cannot reach the threshold; counted as "units_maxscore" in stats()),
        3 + candidate-driven evaluation of rare-term queries (default).
        Results are identical at every level."""
        self._require_index("set_pruning()")
        _lib.check(_lib.lib().bb25_index_set_pruning(self._handle, int(level)))
