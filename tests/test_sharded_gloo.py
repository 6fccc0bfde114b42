"""world_size-2 gloo test of the sharded path's host logic: shard partition +
all-gather plumbing; per-shard compute and the merge are played by the oracle
(there is no GPU here), and the result must equal the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesian_bm25_b200 import sharded, synthetic
    from oracle import coracle

    csc = synthetic.zipf_csc(5000, 600, 30.0, seed=9, device=torch.device("cpu"))
    q_terms, q_off = synthetic.zipf_queries(12, 600, seed=10)
    k = 40
    params = coracle.make_params(1.3, 0.5, 0.03)
    sh = sharded.local_shard(csc, rank, world)
    hs = {k_: (v.numpy() if isinstance(v, torch.Tensor) else v) for k_, v in sh.items()}
    ids, sc, pr, _ = coracle.retrieve_batch(hs, params, q_terms, q_off, k, n_threads=1)
    ids = ids + sh["doc_id_offset"]
    g_ids, g_sc, g_pr = sharded.allgather_topk(torch.from_numpy(ids), torch.from_numpy(sc), torch.from_numpy(pr))
    assert g_ids.shape == (world, 12, k)
    m_ids, m_sc, m_pr = coracle.merge_topk(g_ids.numpy(), g_sc.numpy(), g_pr.numpy())
    full = {k_: (v.numpy() if isinstance(v, torch.Tensor) else v) for k_, v in csc.items()}
    f_ids, f_sc, f_pr, _ = coracle.retrieve_batch(full, params, q_terms, q_off, k, n_threads=1)
    ok = np.array_equal(m_ids, f_ids) and np.array_equal(m_sc, f_sc) and np.allclose(m_pr, f_pr, rtol=0, atol=1e-15)

    # the query-sliced exchange (all_to_all of query slices, merge of the own slice, all_gather of the merged
    # slices): the host logic of ShardedRetriever._exchange_sliced with the device kernels played by NumPy / the oracle
    def pack(i_, s_, p_, out=None):
        key = (s_.numpy().view(np.uint32).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - i_.numpy().astype(np.uint64))
        t = torch.from_numpy(np.stack([key.view(np.int64), p_.numpy().view(np.int64)], axis=-1))
        if out is not None:
            out.copy_(t)
            return out
        return t

    def unpack(pk):
        a = pk.numpy()
        key = a[..., 0].view(np.uint64)
        ids_ = (np.uint64(0xFFFFFFFF) - (key & np.uint64(0xFFFFFFFF))).astype(np.int64)
        sc_ = (key >> np.uint64(32)).astype(np.uint32).view(np.float32)
        return torch.from_numpy(ids_), torch.from_numpy(sc_.copy()), torch.from_numpy(a[..., 1].view(np.float64).copy())

    def merge_packed(pk):
        parts = [unpack(pk[s_]) for s_ in range(pk.shape[0])]
        m = coracle.merge_topk(np.stack([p[0].numpy() for p in parts]), np.stack([p[1].numpy() for p in parts]),
                               np.stack([p[2].numpy() for p in parts]))
        return tuple(torch.from_numpy(x) for x in m)

    sharded.pack_topk_device, sharded.unpack_topk_device, sharded.merge_packed_device = pack, unpack, merge_packed

    class _Scorer:  # the retriever only needs these attributes for the exchange
        _handle = None
        _device = torch.device("cpu")

    retr = sharded.ShardedRetriever(_Scorer(), exchange="sliced", threshold_exchange=False)
    for nq in (12, 11, 1):  # 11 and 1 exercise the padding of the query slices
        s_ids, s_sc, s_pr = retr._exchange_sliced(torch.from_numpy(ids[:nq].copy()), torch.from_numpy(sc[:nq].copy()),
                                                  torch.from_numpy(pr[:nq].copy()))
        ok = ok and np.array_equal(s_ids.numpy(), f_ids[:nq]) and np.array_equal(s_sc.numpy(), f_sc[:nq]) \
            and np.allclose(s_pr.numpy(), f_pr[:nq], rtol=0, atol=1e-15)
    ok = ok and retr.exchange_used.startswith("sliced: all_to_all_single")

    # token -> id mapping split over the ranks and all-gathered (ShardedRetriever._term_ids_distributed)
    from bayesian_bm25_b200.scorer import BayesianBM25Scorer
    sc_ = BayesianBM25Scorer.__new__(BayesianBM25Scorer)
    sc_._vocab = {f"t{i}": i for i in range(600)}
    sc_._device = torch.device("cpu")
    retr.scorer = sc_
    toks = [[f"t{t}" for t in q_terms[q_off[i]:q_off[i + 1]]] + (["oov"] if i % 3 == 0 else []) for i in range(11)] + [[]]
    d_flat, off_h = retr._term_ids_distributed(toks)
    w_flat, w_off = sc_._term_ids_batch(toks)
    ok = ok and np.array_equal(off_h, w_off) and np.array_equal(d_flat.numpy()[:w_flat.size], w_flat)
    with open(os.path.join(out_dir, f"rank{rank}.ok"), "w") as f:
        f.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_allgather_merge_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
