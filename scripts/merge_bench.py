#!/usr/bin/env python
"""Device time of the shard-list merge (S=8, Q=10k, k=1000): bitonic merge tree vs radix select + sort."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_bm25_b200 import sharded

S, Q, K = 8, 10_000, 1000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
# lists must be sorted by (score desc, id asc) -- the merge's contract: ascending ids, distinct scores per rank
sc = (1.0 - torch.arange(K, device=dev, dtype=torch.float32) / K).expand(S, Q, K) * torch.rand((S, Q, 1), device=dev, generator=g)
sc = sc.contiguous()
ids = (torch.arange(S, device=dev).view(S, 1, 1) * 1_100_000 + torch.arange(K, device=dev).view(1, 1, K) * 7).expand(S, Q, K).contiguous().to(torch.int64)
pr = sc.double() * 0.5
packed = torch.stack([sharded.pack_topk_device(ids[s], sc[s], pr[s]) for s in range(S)])
ref = None
for mode in ("tree", "radix", "tree", "radix"):
    os.environ["BB25_MERGE"] = mode
    for _ in range(2):
        out = sharded.merge_packed_device(packed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        out = sharded.merge_packed_device(packed)
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = out
    same = all(torch.equal(a, b) for a, b in zip(out, ref))
    print(f"{mode}: {e0.elapsed_time(e1) / 10:.3f} ms per merge, identical {same}", flush=True)
