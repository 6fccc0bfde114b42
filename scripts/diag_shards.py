#!/usr/bin/env python
"""Diagnostic: the bench corpus cut into S shards scored one after the other on ONE GPU -- per shard the pruned
levels (with / without the essential-posting evaluation) against the exhaustive pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from bayesian_bm25_b200 import BayesianBM25Scorer, index_build, synthetic  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
csc = bench.build_corpus(dev, bench.N_DOCS)
q_terms, q_off = synthetic.zipf_queries(bench.N_QUERIES, bench.VOCAB, bench.QUERY_SEED)
dq, do = torch.from_numpy(q_terms).to(dev), torch.from_numpy(q_off).to(dev)
bad = 0
for si, (lo, hi) in enumerate(index_build.shard_bounds(bench.N_DOCS, S)):
    sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=2.0171221734863845, beta=0.19392475485801697,
                            base_rate=0.035683315909090914)
    sc.index_from_csc(index_build.shard_csc(csc, lo, hi))
    sc.set_pruning(0)
    ref = [t.clone() for t in sc.retrieve_ids_device(dq, do, bench.TOP_K, host_off=q_off)]
    for level, sparse in ((3, "1"), (3, "0"), (2, "1")):
        os.environ["BB25_SPARSE"] = sparse
        sc.set_pruning(level)
        for rep in range(2):
            out = sc.retrieve_ids_device(dq, do, bench.TOP_K, host_off=q_off)
            eq = [bool(torch.equal(a, b)) for a, b in zip(ref, out)]
            if not all(eq):
                bad += 1
                diff_q = torch.nonzero((ref[0] != out[0]).any(dim=1)).flatten().cpu().numpy()
                print(f"shard {si} level {level} sparse {sparse} rep {rep}: MISMATCH {eq}; queries {diff_q[:10]} ({diff_q.size})")
                q = int(diff_q[0]) if diff_q.size else 0
                r = torch.nonzero(ref[0][q] != out[0][q]).flatten().cpu().numpy()
                print("   q", q, "terms", q_terms[q_off[q]:q_off[q + 1]], "first rank", r[:3], "ref", ref[0][q][r[:3]].tolist(), ref[1][q][r[:3]].tolist(),
                      "got", out[0][q][r[:3]].tolist(), out[1][q][r[:3]].tolist())
    print(f"shard {si} done, stats {sc.stats()}")
    del sc
print("MISMATCHES", bad)
