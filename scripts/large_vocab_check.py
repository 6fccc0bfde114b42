#!/usr/bin/env python
"""Large-vocabulary check (run on a B200): 8.8 M docs with a 1 M-term Zipf vocabulary, where a dense
(term, block) table would need 69 GB -- the index must switch rare terms to the bitmap form by
itself, and retrieval must stay bit-identical to the CPU oracle and across pruning levels."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_bm25_b200 import BayesianBM25Scorer, synthetic  # noqa: E402
from oracle import coracle  # noqa: E402

N, V, Q, K = int(os.environ.get("LV_DOCS", 8_800_000)), int(os.environ.get("LV_VOCAB", 1_000_000)), 2000, 100
dev = torch.device("cuda:0")
t0 = time.perf_counter()
csc = synthetic.zipf_csc(N, V, 56.0, 42, dev, k1=1.2, b=0.75, method="lucene")
t_corpus = time.perf_counter() - t0
sc = BayesianBM25Scorer(k1=1.2, b=0.75, method="lucene", alpha=2.0, beta=0.2, base_rate=0.045)
t0 = time.perf_counter()
sc.index_from_csc(csc)
torch.cuda.synchronize()
t_index = time.perf_counter() - t0
info = sc.index_info()
flat, off = synthetic.zipf_queries(Q, V, 43)
out = {}
for level in (0, 3):
    sc.set_pruning(level)
    sc.retrieve_ids(flat, off, K, return_scores=True)  # warm-up (threshold seeds, workspace)
    t0 = time.perf_counter()
    out[level] = sc.retrieve_ids(flat, off, K, return_scores=True)
    out[level] = out[level] + (time.perf_counter() - t0, sc.stats())
same = all(np.array_equal(a, b) for a, b in zip(out[0][:3], out[3][:3]))
host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in csc.items()}
nq = 24
o_ids, o_sc, o_pr, _ = coracle.retrieve_batch(host, coracle.make_params(2.0, 0.2, 0.045), flat[:off[nq]], off[:nq + 1], K)
oracle_ok = bool(np.array_equal(out[3][0][:nq], o_ids) and np.array_equal(out[3][1][:nq].view(np.uint32), o_sc.view(np.uint32))
                 and np.allclose(out[3][2][:nq], o_pr, rtol=0, atol=1e-9))
print(json.dumps({
    "docs": N, "vocab": V, "nnz": int(csc["data"].numel()), "corpus_s": round(t_corpus, 1), "index_s": round(t_index, 1),
    "dense_table_would_be_gb": round(((N + 1023) // 1024) * V * 8 / 1e9, 1),
    "block_table_gb": round(info["block_table_bytes"] / 1e9, 2), "bitmap_terms": info["block_table_bitmap_terms"],
    "device_gb": round(info["device_bytes"] / 1e9, 2),
    "queries": Q, "k": K, "qps_exhaustive_e2e": round(Q / out[0][3]), "qps_level3_e2e": round(Q / out[3][3]),
    "levels_identical": same, "oracle_identical_first_%d" % nq: oracle_ok,
    "level3_stats": {k: out[3][4][k] for k in ("units", "units_skipped", "routed_queries", "rerun_queries")},
}), flush=True)
assert same and oracle_ok and info["block_table_bitmap_terms"] > 0
