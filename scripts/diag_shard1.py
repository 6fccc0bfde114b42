#!/usr/bin/env python
"""Diagnostic: shard 1 of 8 of the bench corpus, query 3050 ([0, 7960, 17]) -- which pruning ingredient loses documents?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from bayesian_bm25_b200 import BayesianBM25Scorer, index_build, synthetic  # noqa: E402
from oracle import coracle  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
csc = bench.build_corpus(dev, bench.N_DOCS)
lo, hi = index_build.shard_bounds(bench.N_DOCS, 8)[1]
shard = index_build.shard_csc(csc, lo, hi)
del csc
q_terms, q_off = synthetic.zipf_queries(bench.N_QUERIES, bench.VOCAB, bench.QUERY_SEED)
host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in shard.items()}
P = (2.0171221734863845, 0.19392475485801697, 0.035683315909090914)
params = coracle.make_params(*P)
one_t = q_terms[q_off[3050]:q_off[3051]].astype(np.int32)
one_o = np.array([0, one_t.size], dtype=np.int64)
o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, params, one_t, one_o, bench.TOP_K)
print("oracle tail", o_ids[0, 995:].tolist(), o_sc[0, 995:].tolist())
dq, do = torch.from_numpy(q_terms).to(dev), torch.from_numpy(q_off).to(dev)
for half in ("1", "0"):
    os.environ["BB25_HALF_ROWS"] = half
    sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=P[0], beta=P[1], base_rate=P[2])
    sc.index_from_csc(shard)
    for level, sparse in ((0, "1"), (1, "1"), (2, "0"), (2, "1"), (3, "0")):
        os.environ["BB25_SPARSE"] = sparse
        sc.set_pruning(level)
        ids, scs, prs = sc.retrieve_ids_device(dq, do, bench.TOP_K, host_off=q_off)
        i1, s1, _ = sc.retrieve_ids(one_t, one_o, bench.TOP_K, return_scores=True)
        ok_b = np.array_equal(ids[3050].cpu().numpy(), o_ids[0]) and np.array_equal(scs[3050].cpu().numpy(), o_sc[0])
        ok_1 = np.array_equal(i1[0], o_ids[0]) and np.array_equal(s1[0], o_sc[0])
        nbad = int((ids.cpu().numpy()[:, :] != ids.cpu().numpy()[:, :]).sum())
        print(f"half {half} level {level} sparse {sparse}: batch {'ok' if ok_b else 'WRONG'} alone {'ok' if ok_1 else 'WRONG'} tail {ids[3050, 996:].tolist()} stats {sc.stats()['units_skipped']} {sc.stats()['units_maxscore']} {sc.stats()['units_sparse']}")
    del sc
