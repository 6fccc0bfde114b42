#!/usr/bin/env python
"""Debug aid: per-call stats of the bench workload, order-free vs query-order traversal."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesian_bm25_b200 import BayesianBM25Scorer, sharded, synthetic

dev = torch.device("cuda:0")
csc = bench.build_corpus(dev, bench.N_DOCS)
sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=bench.ALPHA, beta=bench.BETA, base_rate=bench.BASE_RATE)
sc.index_from_csc(csc)
retr = sharded.ShardedRetriever(sc, n_chunks=1)
q_terms, q_off = synthetic.zipf_queries(bench.N_QUERIES, bench.VOCAB, bench.QUERY_SEED)
d_terms = torch.from_numpy(q_terms).to(dev); d_off = torch.from_numpy(q_off).to(dev)
k = int(sys.argv[1]) if len(sys.argv) > 1 else bench.TOP_K
ref = None
for level in (0, 3):
    sc.set_pruning(level)
    for relaxed in ("0", "1", "1", "1"):
        os.environ["BB25_RELAXED"] = relaxed
        out = retr.retrieve_ids_device(d_terms, d_off, k)
        torch.cuda.synchronize()
        st = retr.stats()
        ids, scores, probs = [o.clone() for o in out]
        if ref is None:
            ref = (ids, scores, probs)
        same = bool(torch.equal(ids, ref[0]) and torch.equal(scores, ref[1]) and torch.equal(probs, ref[2]))
        print(f"level {level} relaxed {relaxed}: reruns {st['rerun_queries']} passes {st.get('passes')} launches {st.get('launches')} "
              f"trav_ms {st['traverse_ms']:.1f} candidates {st.get('candidates')} identical_to_first {same}", flush=True)
