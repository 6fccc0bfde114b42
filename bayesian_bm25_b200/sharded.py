"""Document-range sharding across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).
Rank s holds the CSC rows of documents [s*N/S, (s+1)*N/S) with local doc ids and a
``doc_id_offset``; posting values carry the GLOBAL idf / length statistics, so a
shard's fp32 scores equal the unsharded index's.  A query batch is scored by
every rank on its shard; the ranks then exchange their per-shard [Q,k]
(id, score, probability) lists and merge them with the same (score desc, id asc) order.

Two things keep the per-rank work shrinking with the number of shards:

* cross-shard thresholds -- between the block groups of a batch the ranks all-gather the
  scores at a few ranks of their running top-k (Q x 4 x 8 bytes per rank) and every shard
  raises its thresholds to the bound on the GLOBAL k-th score that follows
  (bb25_apply_quantiles), instead of hunting its own, looser, local top-k;
* a query-sliced exchange -- rank r merges only the queries [r*Q/S, (r+1)*Q/S).  Over
  symmetric memory (``exchange="sliced"``, NCCL backend) this is ONE kernel per rank
  (bb25_merge_topk_peers): it pulls the S sorted lists of its queries out of the peers'
  memory over NVLink, merges them and stores the merged rows into every rank's result
  buffer -- all-to-all, merge and all-gather in one launch.  Without symmetric memory
  (gloo, or ``BB25_SYMM=0``) the same slicing runs as all_to_all_single + merge +
  all_gather.  ``exchange="allgather"`` is the round-1 scheme (every rank receives and
  merges everything).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .index_build import shard_bounds, shard_csc


def local_shard(csc: dict, rank: int, world_size: int) -> dict:
    lo, hi = shard_bounds(int(csc["num_docs"]), world_size)[rank]
    return shard_csc(csc, lo, hi)


def allgather_topk(ids: torch.Tensor, scores: torch.Tensor, probs: torch.Tensor, group=None):
    """[Q,k] x3 per rank -> [S,Q,k] x3 on every rank (works on NCCL and gloo)."""
    world = dist.get_world_size(group)
    outs = []
    for t in (ids, scores, probs):
        t = t.contiguous()
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)
        outs.append(buf.view((world,) + tuple(t.shape)))
    return tuple(outs)


def merge_topk_device(ids: torch.Tensor, scores: torch.Tensor, probs: torch.Tensor):
    """[S,Q,k] x3 CUDA tensors -> merged [Q,k] x3 (libbb25 merge kernel)."""
    s, q, k = ids.shape
    out_ids = torch.empty((q, k), dtype=torch.int64, device=ids.device)
    out_sc = torch.empty((q, k), dtype=torch.float32, device=ids.device)
    out_pr = torch.empty((q, k), dtype=torch.float64, device=ids.device)
    _lib.check(_lib.lib().bb25_merge_topk(
        ids.device.index, ids.contiguous().data_ptr(), scores.contiguous().data_ptr(),
        probs.contiguous().data_ptr(), s, q, k, out_ids.data_ptr(), out_sc.data_ptr(), out_pr.data_ptr(),
        _lib.stream_ptr()))
    return out_ids, out_sc, out_pr


def pack_topk_device(ids: torch.Tensor, scores: torch.Tensor, probs: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[Q,k] x3 -> [Q,k,2] int64 entries {merge key, probability bits} (global ids < 2^32)."""
    if out is None:
        out = torch.empty(tuple(ids.shape) + (2,), dtype=torch.int64, device=ids.device)
    _lib.check(_lib.lib().bb25_pack_topk(ids.device.index, ids.contiguous().data_ptr(), scores.contiguous().data_ptr(),
                                         probs.contiguous().data_ptr(), ids.numel(), out.data_ptr(), _lib.stream_ptr()))
    return out


def merge_packed_device(packed: torch.Tensor):
    """[S,Q,k,2] int64 -> merged (ids int64, scores fp32, probs fp64) [Q,k]."""
    s, q, k, _ = packed.shape
    out_ids = torch.empty((q, k), dtype=torch.int64, device=packed.device)
    out_sc = torch.empty((q, k), dtype=torch.float32, device=packed.device)
    out_pr = torch.empty((q, k), dtype=torch.float64, device=packed.device)
    _lib.check(_lib.lib().bb25_merge_topk_packed(packed.device.index, packed.contiguous().data_ptr(), s, q, k,
                                                 out_ids.data_ptr(), out_sc.data_ptr(), out_pr.data_ptr(),
                                                 _lib.stream_ptr()))
    return out_ids, out_sc, out_pr


def unpack_topk_device(packed: torch.Tensor):
    """[Q,k,2] int64 packed entries -> (ids int64, scores fp32, probs fp64) [Q,k]."""
    q, k, _ = packed.shape
    out_ids = torch.empty((q, k), dtype=torch.int64, device=packed.device)
    out_sc = torch.empty((q, k), dtype=torch.float32, device=packed.device)
    out_pr = torch.empty((q, k), dtype=torch.float64, device=packed.device)
    _lib.check(_lib.lib().bb25_unpack_topk(packed.device.index, packed.data_ptr(), q * k, out_ids.data_ptr(),
                                           out_sc.data_ptr(), out_pr.data_ptr(), _lib.stream_ptr()))
    return out_ids, out_sc, out_pr


_EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p)


class ShardedRetriever:
    """Wraps a rank-local BayesianBM25Scorer (indexed on this rank's shard).

    ``exchange``: "sliced" (default) or "allgather", see the module docstring.
    ``threshold_exchange``: all-gather score quantiles between block groups (cross-shard thresholds).
    With n_chunks > 1 the batch is processed in sub-batches whose exchange overlaps the next sub-batch's
    traversal (measured slower at Q = 10 k on 8xB200, profiles/r01; kept for larger batches).
    """

    def __init__(self, scorer, group=None, profile: bool = False, n_chunks: int = 1, exchange: str = "sliced",
                 threshold_exchange: bool = True):
        if exchange not in ("sliced", "allgather"):
            raise ValueError(f"exchange must be 'sliced' or 'allgather', got {exchange!r}")
        self.scorer = scorer
        self.group = group
        self.profile = profile
        self.n_chunks = max(1, int(n_chunks))
        self.exchange = exchange
        self.threshold_exchange = bool(threshold_exchange)
        self.timing = {"local_ms": 0.0, "gather_ms": 0.0, "merge_ms": 0.0, "calls": 0}
        self._comm_stream = None
        self._stats: dict = {}
        self._cb = None          # keeps the ctypes callback alive
        self._cb_error = None
        self._quant_bufs: dict = {}
        self._symm: dict = {}    # (Qp, k) -> symmetric buffers and handles
        self._symm_failed = os.environ.get("BB25_SYMM", "1") == "0"
        self._seeded: set = set()
        self.exchange_used = None
        self._install_threshold_exchange()

    # ---- plumbing ----------------------------------------------------------------------------
    def _sharded(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def stats(self) -> dict:
        """Counters of the last retrieve_ids_device call, summed over its sub-batches."""
        return dict(self._stats)

    def _add_stats(self, reset: bool = False):
        st = self.scorer.stats()
        if reset or not self._stats:
            self._stats = st
        else:
            self._stats = {k: self._stats.get(k, 0) + v for k, v in st.items()}

    def _install_threshold_exchange(self):
        """Registers the between-groups callback of libbb25 (include/bb25.h: bb25_exchange_fn)."""
        lib = _lib.lib()
        if not (self._sharded() and self.threshold_exchange) or self.scorer._handle is None:
            return
        world = dist.get_world_size(self.group)
        if world > 32:
            return
        dev = self.scorer._device

        def on_group(user, d_quant, d_thr, n_q, n_levels, k, group_idx, stream):
            try:
                key = (int(n_q), int(n_levels))
                bufs = self._quant_bufs.get(key)
                if bufs is None:
                    bufs = (torch.empty((n_q, n_levels), dtype=torch.int64, device=dev),
                            torch.empty((world * n_q, n_levels), dtype=torch.int64, device=dev))
                    self._quant_bufs[key] = bufs
                mine, everyone = bufs
                _lib.check(lib.bb25_memcpy_device(dev.index, mine.data_ptr(), d_quant, n_q * n_levels * 8, stream))
                # the collective must be ordered with the batch's stream, whatever torch's current stream is
                with torch.cuda.stream(torch.cuda.ExternalStream(stream or 0, device=dev)):
                    dist.all_gather_into_tensor(everyone, mine, group=self.group)
                _lib.check(lib.bb25_apply_quantiles(dev.index, everyone.data_ptr(), world, n_q, k, d_thr, stream))
                return 0
            except Exception as e:  # an exception must not unwind through the C frame
                self._cb_error = e
                return 1

        self._cb = _EXCHANGE_FN(on_group)
        _lib.check(lib.bb25_index_set_threshold_exchange(self.scorer._handle, C.cast(self._cb, C.c_void_p), None, world))

    def close(self):
        if self._cb is not None and self.scorer._handle is not None:
            _lib.lib().bb25_index_set_threshold_exchange(self.scorer._handle, None, None, 0)
        self._cb = None

    def _install_global_seeds(self, k: int) -> None:
        """Once per k: every shard's per-term posting values at the ranks ceil(k/S), ceil(2k/S), ceil(4k/S), k
        are all-gathered and combined into a lower bound of each term's k-th largest value over the WHOLE corpus
        (bb25_apply_quantiles with terms in place of queries); the result replaces the shard's own threshold
        seeds, so batches start with the seeds the unsharded index would use -- tighter thresholds in the first
        block group and the same candidate-path routing decisions as a single GPU."""
        if k in self._seeded or not (self._sharded() and self.threshold_exchange) or self.scorer._handle is None:
            return
        self._seeded.add(k)
        lib = _lib.lib()
        world = dist.get_world_size(self.group)
        if world > 32:
            return
        dev = self.scorer._device
        n_levels, ranks = C.c_int(), (C.c_int * 4)()
        lib.bb25_quantile_ranks(k, world, C.byref(n_levels), ranks)
        j = n_levels.value
        v = self.scorer._n_vocab
        cols = []
        for r in list(ranks)[:j]:
            ptr = C.c_void_p()
            _lib.check(lib.bb25_index_kth_values(self.scorer._handle, int(r), C.byref(ptr), _lib.stream_ptr()))
            t = torch.empty(v, dtype=torch.float32, device=dev)
            _lib.check(lib.bb25_memcpy_device(dev.index, t.data_ptr(), ptr.value, v * 4, _lib.stream_ptr()))
            cols.append(t)
        mine = (torch.stack(cols, dim=1).contiguous().view(torch.int32).to(torch.int64) << 33)  # [V, J] score keys
        everyone = torch.empty((world * v, j), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        bound = torch.zeros(v, dtype=torch.int64, device=dev)
        _lib.check(lib.bb25_apply_quantiles(dev.index, everyone.data_ptr(), world, v, k, bound.data_ptr(), _lib.stream_ptr()))
        seeds = (bound >> 33).to(torch.int32).view(torch.float32)
        seeds = torch.maximum(seeds, cols[-1]).contiguous()  # never below the shard's own k-th value
        _lib.check(lib.bb25_index_set_kth_values(self.scorer._handle, k, seeds.data_ptr(), _lib.stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def _local(self, q_terms, q_off, k, host_off):
        self._install_global_seeds(k)
        try:
            out = self.scorer.retrieve_ids_device(q_terms, q_off, k, host_off=host_off)
        except RuntimeError:
            if self._cb_error is not None:
                err, self._cb_error = self._cb_error, None
                raise err
            raise
        return out

    # ---- exchange variants --------------------------------------------------------------------
    def _exchange_allgather(self, ids, sc, pr):
        packed = pack_topk_device(ids, sc, pr)
        world = dist.get_world_size(self.group)
        buf = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(buf, packed, group=self.group)
        return merge_packed_device(buf.view((world,) + tuple(packed.shape)))

    def _symm_buffers(self, qp: int, k: int, device):
        """Symmetric (peer-mapped) source / result buffers for a padded batch size, made once."""
        key = (qp, k)
        ent = self._symm.get(key)
        if ent is not None:
            return ent
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        src = symm_mem.empty((qp, k, 2), dtype=torch.int64, device=device)
        dst = symm_mem.empty((qp, k, 2), dtype=torch.int64, device=device)
        h_src = symm_mem.rendezvous(src, grp)
        h_dst = symm_mem.rendezvous(dst, grp)
        ent = (src, dst, h_src, h_dst)
        self._symm[key] = ent
        return ent

    def _exchange_sliced(self, ids, sc, pr):
        """Rank r merges the queries [r*Qs, (r+1)*Qs) only; see the module docstring."""
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        q, k = ids.shape
        qs = -(-q // world)
        qp = qs * world
        dev = ids.device
        use_symm = ids.is_cuda and dist.get_backend(self.group) == "nccl" and not self._symm_failed
        if use_symm:
            try:
                src, dst, h_src, h_dst = self._symm_buffers(qp, k, dev)
            except Exception:
                self._symm_failed = True
                use_symm = False
        if use_symm:
            self.exchange_used = "sliced: one merge kernel over symmetric peer memory (NVLink loads + stores)"
            if qp != q:
                src[q:].zero_()
            pack_topk_device(ids, sc, pr, out=src)
            h_src.barrier(channel=0)  # every shard's packed lists are in place
            _lib.check(_lib.lib().bb25_merge_topk_peers(dev.index, h_src.buffer_ptrs_dev, h_dst.buffer_ptrs_dev, world,
                                                        rank * qs, qs, k, _lib.stream_ptr()))
            h_dst.barrier(channel=1)  # every rank's merged rows have landed here; sources may be overwritten again
            return unpack_topk_device(dst[:q])
        # NCCL / gloo: all-to-all of query slices, merge of the own slice, all-gather of the merged slices
        self.exchange_used = "sliced: all_to_all_single + merge + all_gather"
        packed = torch.zeros((qp, k, 2), dtype=torch.int64, device=dev)
        packed[:q] = pack_topk_device(ids, sc, pr)
        recv = torch.empty_like(packed)
        dist.all_to_all_single(recv, packed, group=self.group)  # recv[s*qs:(s+1)*qs] = shard s's lists of my queries
        m_ids, m_sc, m_pr = merge_packed_device(recv.view(world, qs, k, 2))
        mine = pack_topk_device(m_ids, m_sc, m_pr)
        full = torch.empty((qp, k, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return unpack_topk_device(full[:q])

    def _exchange(self, ids, sc, pr):
        if self.exchange == "allgather":
            self.exchange_used = "allgather: every rank receives and merges every list"
            return self._exchange_allgather(ids, sc, pr)
        return self._exchange_sliced(ids, sc, pr)

    # ---- retrieval ----------------------------------------------------------------------------
    def retrieve_ids_device(self, q_terms: torch.Tensor, q_off: torch.Tensor, k: int, host_off=None):
        if not self._sharded():
            out = self._local(q_terms, q_off, k, host_off)
            self._add_stats(reset=True)
            return out
        nq = q_off.numel() - 1
        n_chunks = min(self.n_chunks, max(1, nq // 256))
        if self.profile or n_chunks == 1:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.profile else None
            if ev:
                ev[0].record()
            ids, sc, pr = self._local(q_terms, q_off, k, host_off)
            self._add_stats(reset=True)
            if ev:
                ev[1].record()
            out = self._exchange(ids, sc, pr)
            if ev:
                ev[2].record()
                ev[2].synchronize()
                self.timing["local_ms"] += ev[0].elapsed_time(ev[1])
                self.timing["gather_ms"] += ev[1].elapsed_time(ev[2])  # exchange + merge (one kernel when sliced over peers)
                self.timing["calls"] += 1
            return out
        # pipelined: exchange of sub-batch i on the side stream, traversal of i+1 on the main one
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=q_off.device)
        main = torch.cuda.current_stream()
        bounds = [nq * c // n_chunks for c in range(n_chunks + 1)]
        outs = []
        for c in range(n_chunks):
            lo, hi = bounds[c], bounds[c + 1]
            ids, sc, pr = self._local(q_terms, q_off[lo:hi + 1], k, None if host_off is None else host_off[lo:hi + 1])
            self._add_stats(reset=(c == 0))
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(done)
                for t in (ids, sc, pr):
                    t.record_stream(self._comm_stream)
                outs.append(self._exchange_allgather(ids, sc, pr))
        main.wait_stream(self._comm_stream)
        for o in outs:
            for t in o:
                t.record_stream(main)
        return tuple(torch.cat([o[j] for o in outs], dim=0) for j in range(3))

    def _to_host(self, outs, result: str):
        """Merged device tensors -> page-locked host arrays on the ranks that want them."""
        if result == "rank0" and self._sharded() and dist.get_rank(self.group) != 0:
            torch.cuda.current_stream().synchronize()
            return None
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in outs]
        for h, t in zip(host, outs):
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return tuple(h.numpy() for h in host)

    def retrieve_ids(self, q_terms, q_off, k: int = 10, return_scores: bool = False, result: str = "all"):
        """Host in / host out: every rank passes the same queries.  result="all": every rank gets the merged
        result; "rank0": only rank 0 copies it to the host (the others return None)."""
        q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        dev = self.scorer._device
        hp = torch.empty(max(q_terms.size, 1), dtype=torch.int32, pin_memory=True)
        hp.numpy()[:q_terms.size] = q_terms
        ho = torch.empty(q_off.size, dtype=torch.int64, pin_memory=True)
        ho.numpy()[:] = q_off
        ids, sc, pr = self.retrieve_ids_device(hp.to(dev, non_blocking=True), ho.to(dev, non_blocking=True), k, host_off=q_off)
        return self._to_host([ids, sc, pr] if return_scores else [ids, pr], result)

    def _term_ids_distributed(self, query_tokens):
        """Token strings -> term ids with the work split over the ranks: rank r maps the queries
        [r*Q/S, (r+1)*Q/S) through the vocabulary (Python dict lookups, ~2.5 us per query) and the ranks
        all-gather the ids -- 1/S of the host-side mapping per rank.  Returns (flat ids on the device,
        offsets on the host)."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = self.scorer._device
        nq = len(query_tokens)
        qs = -(-nq // world)
        lo, hi = min(nq, rank * qs), min(nq, (rank + 1) * qs)
        flat, off = self.scorer._term_ids_batch(query_tokens[lo:hi])
        lens = np.zeros(qs, dtype=np.int64)
        lens[:hi - lo] = np.diff(off)
        all_lens = torch.empty(world * qs, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_lens, torch.from_numpy(lens).to(dev), group=self.group)
        all_lens_h = all_lens.cpu().numpy()
        totals = all_lens_h.reshape(world, qs).sum(axis=1)
        tmax = max(1, int(totals.max()))
        mine = torch.zeros(tmax, dtype=torch.int32, device=dev)
        if flat.size:
            mine[:flat.size] = torch.from_numpy(flat).to(dev)
        everyone = torch.empty(world * tmax, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        parts = [everyone[r * tmax:r * tmax + int(totals[r])] for r in range(world)]
        d_flat = torch.cat(parts) if int(totals.sum()) else torch.zeros(1, dtype=torch.int32, device=dev)
        off_h = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(all_lens_h[:nq], out=off_h[1:])
        return d_flat, off_h

    def retrieve(self, query_tokens: list[list[str]], k: int = 10, result: str = "all"):
        """BayesianBM25Scorer.retrieve on the sharded index: (doc_ids [Q,k], probabilities [Q,k]) host arrays
        (on rank 0 only with result="rank0").  Every rank passes the same token lists."""
        if not self._sharded():
            flat, off = self.scorer._term_ids_batch(query_tokens)
            return self.retrieve_ids(flat, off, k)
        d_flat, off_h = self._term_ids_distributed(query_tokens)
        d_off = torch.from_numpy(off_h).to(self.scorer._device)
        ids, _, pr = self.retrieve_ids_device(d_flat, d_off, k, host_off=off_h)
        return self._to_host([ids, pr], result)
