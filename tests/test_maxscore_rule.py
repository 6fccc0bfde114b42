"""CPU check of the RULE behind the essential-posting evaluation (csrc/bb25_traverse.cu: group_units; the fused form in
csrc/bb25_fused.cu): MaxScore -- the partition behind wand_upper_bound (probability.py:205-236) and BlockMaxIndex
(scorer.py:33-142) -- applied per 1024-document block.  A plain NumPy restatement of the split, run against exhaustive
scores: whatever the threshold, every document of a block whose exact fp32 score reaches it has a posting in an ESSENTIAL
slice, and a block whose ladder step 7 (all terms) stays below the threshold holds no such document.  The CUDA path is
checked against the oracle in tests/test_gpu_parity.py; this pins the rule itself."""
import numpy as np

BLOCK = 1024
LADDER = (1, 2, 3, 5, 9, 17, 33)


def _corpus(n_docs, vocab, avg_len, seed):
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, vocab + 1)
    w /= w.sum()
    lens = np.maximum(2, rng.normal(avg_len, 0.3 * avg_len, n_docs).astype(np.int64))
    docs = np.repeat(np.arange(n_docs), lens)
    terms = rng.choice(vocab, size=docs.size, p=w)
    key = np.unique(terms.astype(np.int64) * n_docs + docs, return_counts=True)
    t, d, tf = key[0] // n_docs, key[0] % n_docs, key[1].astype(np.float64)
    df = np.bincount(t, minlength=vocab)
    idf = np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))
    avgdl = lens.mean()
    val = (idf[t] * tf / (tf + 1.2 * (1 - 0.75 + 0.75 * lens[d] / avgdl))).astype(np.float32)
    return t, d, val, df


def _round_up_21(x):
    """block maxima as the table stores them: fp32 rounded UP to 21 significant bits (low 11 bits of the word hold the length)"""
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FF) & ~np.uint64(0x7FF)
    return b.astype(np.uint32).view(np.float32)


def test_essential_slices_hold_every_qualifying_document():
    n_docs, vocab = 20_000, 400
    t, d, val, df = _corpus(n_docs, vocab, 30.0, 11)
    n_blocks = -(-n_docs // BLOCK)
    has_row = df * 64 >= n_docs  # hot or lookup value row
    rng = np.random.default_rng(12)
    w = 1.0 / np.arange(1, vocab + 1)
    w /= w.sum()
    checked_units = sparse_units = skipped_units = 0
    for _ in range(40):
        q = rng.choice(vocab, size=rng.integers(2, 7), p=w)
        # exact scores in query order (fp32, duplicates count twice as in bm25s)
        score = np.zeros(n_docs, dtype=np.float32)
        per_term = []
        for term in q:
            m = t == term
            col = np.zeros(n_docs, dtype=np.float32)
            col[d[m]] = val[m]
            per_term.append(col)
            score = (score + col).astype(np.float32)
        for k in (5, 50, 500):
            thr = np.sort(score)[::-1][k - 1]
            if thr <= 0:
                continue
            for blk in range(n_blocks):
                lo, hi = blk * BLOCK, min(n_docs, (blk + 1) * BLOCK)
                cols = [c[lo:hi] for c in per_term]
                lens = np.array([(c > 0).sum() for c in cols])
                bmax = _round_up_21(np.array([c.max() for c in cols], dtype=np.float32))
                rows = has_row[q]
                qualifying = score[lo:hi] >= thr
                checked_units += 1

                def bound(mask):
                    ub = np.float32(0)
                    for i in np.nonzero(mask)[0]:  # query order
                        ub = np.float32(ub + bmax[i])
                    return ub

                if bound(lens >= 1) < thr:  # ladder step 7: the block-max test
                    assert not qualifying.any()
                    skipped_units += 1
                    continue
                L = next((c for c in LADDER if bound(rows & (lens >= c)) < thr), None)
                if L is None:
                    continue  # the unit takes the pass
                essential = (lens > 0) & (~rows | (lens < L))
                in_essential = np.zeros(hi - lo, dtype=bool)
                for i in np.nonzero(essential)[0]:
                    in_essential |= cols[i] > 0
                assert not (qualifying & ~in_essential).any(), (q, k, blk, L)
                sparse_units += 1
    assert checked_units > 1000 and sparse_units > 100 and skipped_units > 10
