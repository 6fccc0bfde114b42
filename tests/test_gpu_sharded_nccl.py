"""Two ranks over NCCL (skipped on a single-GPU box): the document-sharded batch retrieve -- cross-shard
threshold exchange between block groups, query-sliced exchange with the peer-memory merge kernel (and its
NCCL fallback, and the round-1 all-gather scheme) -- must equal the unsharded index and the CPU oracle:
ids and fp32 scores bit for bit, probabilities to 1e-9."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

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
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    msg = "ok"
    try:
        from bayesian_bm25_b200 import BayesianBM25Scorer, sharded, synthetic
        from oracle import coracle
        n_docs, vocab, nq, k = 1_000_000, 30_000, 203, 100
        csc = synthetic.zipf_csc(n_docs, vocab, 56.0, seed=42, device=dev)
        q_terms, q_off = synthetic.zipf_queries(nq - 3, vocab, seed=43)
        extra = [np.array([0, 1, 2], np.int32), np.array([7, 7, 29999], np.int32), np.zeros(0, np.int32)]
        q_terms = np.concatenate([q_terms] + extra).astype(np.int32)
        q_off = np.concatenate([q_off, q_off[-1] + np.cumsum([len(e) for e in extra])]).astype(np.int64)
        a, b, br = 2.0, 0.2, 0.04
        full = BayesianBM25Scorer(method="lucene", alpha=a, beta=b, base_rate=br)
        full.index_from_csc(csc)
        want = full.retrieve_ids(q_terms, q_off, k, return_scores=True)
        del full
        sc = BayesianBM25Scorer(method="lucene", alpha=a, beta=b, base_rate=br)
        sc.index_from_csc(sharded.local_shard(csc, rank, world))
        if rank == 0:
            host = {k_: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k_, v in csc.items()}
            o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, coracle.make_params(a, b, br), q_terms, q_off, k)
            assert np.array_equal(want[0], o_ids) and np.array_equal(want[1].view(np.uint32), o_sc.view(np.uint32))
            assert np.max(np.abs(want[2] - o_pr)) < 1e-9
        del csc
        used = []
        for exchange, thr_x, symm in (("sliced", True, "1"), ("sliced", True, "0"), ("sliced", False, "1"),
                                      ("allgather", True, "1"), ("allgather", False, "1")):
            os.environ["BB25_SYMM"] = symm
            retr = sharded.ShardedRetriever(sc, exchange=exchange, threshold_exchange=thr_x)
            for level in (0, 3):
                sc.set_pruning(level)
                got = retr.retrieve_ids(q_terms, q_off, k, return_scores=True)
                assert np.array_equal(got[0], want[0]), (exchange, thr_x, symm, level, "ids")
                assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), (exchange, thr_x, symm, level, "scores")
                assert np.array_equal(got[2], want[2]), (exchange, thr_x, symm, level, "probs")
            used.append(retr.exchange_used)
            retr.close()
        if rank == 0:
            msg = "ok | " + " ; ".join(dict.fromkeys(used))
    except BaseException as e:  # noqa: BLE001 - reported through the file, the parent asserts
        import traceback
        msg = "FAIL " + repr(e) + "\n" + traceback.format_exc()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(msg)
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


def test_two_rank_nccl_sharded_equals_unsharded_and_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        txt = open(tmp_path / f"rank{r}.txt").read()
        assert txt.startswith("ok"), txt
    print(open(tmp_path / "rank0.txt").read())
