"""Dense side of hybrid retrieval: cosine similarities of a query batch against every document of the corpus
(``query_emb @ corpus_emb.T``, benchmarks/hybrid_beir.py:1751-1753) on the 5th-generation tensor cores
(bb25_cosine_gemm: TMA-staged bf16 tiles, tcgen05.mma with the fp32 accumulator in tensor memory), written
row-per-query in the layout ``bb25_retrieve_fused_batch`` consumes.  Embeddings are expected L2-normalised
(cosine = dot product), as the reference's benchmarks prepare them."""
from __future__ import annotations

import torch

from . import _lib

MAX_QUERIES_PER_LAUNCH = 256


def cosine_scores(query_emb: torch.Tensor, corpus_emb: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """query_emb [Q, K], corpus_emb [N, K] (CUDA; converted to contiguous bf16 if needed; K a multiple of 64)
    -> fp32 CUDA [Q, stride] with stride = N rounded up to a multiple of 4 (columns >= N are padding)."""
    dev = _lib.require_cuda()
    if query_emb.dim() != 2 or corpus_emb.dim() != 2 or query_emb.shape[1] != corpus_emb.shape[1]:
        raise ValueError("query_emb must be [Q, K] and corpus_emb [N, K]")
    q = query_emb.to(device=f"cuda:{dev}", dtype=torch.bfloat16).contiguous()
    c = corpus_emb.to(device=f"cuda:{dev}", dtype=torch.bfloat16).contiguous()
    nq, k = q.shape
    n = c.shape[0]
    stride = (n + 3) // 4 * 4
    if out is None:
        out = torch.empty((nq, stride), dtype=torch.float32, device=q.device)
        if stride != n:
            out[:, n:].zero_()
    elif out.dtype != torch.float32 or out.shape[0] != nq or out.stride(1) != 1 or out.stride(0) < n:
        raise ValueError("out must be a float32 CUDA tensor [Q, >= N] with contiguous rows")
    for s in range(0, nq, MAX_QUERIES_PER_LAUNCH):
        e = min(nq, s + MAX_QUERIES_PER_LAUNCH)
        _lib.check(_lib.lib().bb25_cosine_gemm(dev, q[s:e].data_ptr(), e - s, c.data_ptr(), n, k, out[s:e].data_ptr(),
                                               out.stride(0), _lib.stream_ptr()))
    return out
