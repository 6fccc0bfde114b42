#!/usr/bin/env python
"""Two batch-retrieve steps of the headline workload at one pruning level (for ncu captures):
    python scripts/prof_step.py --level 0|3 [--docs N --queries Q --k K]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from bayesian_bm25_b200 import BayesianBM25Scorer, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--docs", type=int, default=bench.N_DOCS)
ap.add_argument("--queries", type=int, default=bench.N_QUERIES)
ap.add_argument("--k", type=int, default=bench.TOP_K)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
csc = bench.build_corpus(dev, a.docs)
sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=2.0171221734863845, beta=0.19392475485801697,
                        base_rate=0.035683315909090914)
sc.index_from_csc(csc)
del csc
sc.set_pruning(a.level)
q_terms, q_off = synthetic.zipf_queries(a.queries, bench.VOCAB, bench.QUERY_SEED)
dt, do = torch.from_numpy(q_terms).to(dev), torch.from_numpy(q_off).to(dev)
for _ in range(a.steps):
    sc.retrieve_ids_device(dt, do, a.k, host_off=q_off)
torch.cuda.synchronize()
print("stats", sc.stats())
