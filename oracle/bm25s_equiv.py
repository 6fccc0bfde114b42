"""NumPy restatement of the slice of `bm25s` the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference imports the
third-party package `bm25s` (pyproject.toml:20,22,28 -- ``bm25s>=0.2.0``, no
lock file) at ``bayesian_bm25/scorer.py:20-26`` and uses exactly five entry
points: ``BM25(k1, b, method)`` (:213), ``.index(corpus_tokens, show_progress)``
(:262), ``.scores["num_docs"]`` (:227), ``.get_scores(tokens)`` (:306, :583)
and ``.retrieve(queries, k, sorted, show_progress)`` -> ``.documents/.scores``
(:525-529).  bm25s is not installed in this image and cannot be, so this
module restates its published algorithm (0.2.x) from memory:

* vocabulary = token -> id plus an extra ``""`` token;
* ``idf``: robertson ``ln(max(1,(N-df+.5)/(df+.5)))``, lucene
  ``ln(1+(N-df+.5)/(df+.5))``, atire ``ln(N/df)``, stored fp32;
* ``tfc``: robertson = lucene ``tf/(k1*(1-b+b*l_d/l_avg)+tf)``, atire
  ``tf*(k1+1)/(tf+k1*(1-b+b*l_d/l_avg))``;
* posting value ``fp32(idf*tfc)`` with whatever intermediate precision NumPy
  gives ``float32_array (op) np.float64_scalar`` (NumPy 2: float64);
* ``scipy.sparse.csc_matrix`` -> ``data/indices/indptr`` (doc ids ascending
  inside a column);
* ``get_scores``: fp32 zeros, then ``np.add.at`` per in-vocabulary query token
  in query order, duplicates included;
* ``retrieve``: per query ``get_scores`` then top-k; ``ValueError`` if k > N.

PARITY UNPINNED for the BM25 values: no reference test asserts one.
Tie order in bm25s's ``argpartition`` is unspecified; here ties are resolved
canonically (score desc, doc id asc), the rule the B200 path implements.
"""
from __future__ import annotations

import math
from collections import Counter, namedtuple

import numpy as np

Results = namedtuple("Results", ["documents", "scores"])


def _idf(method: str, df: int, n: int) -> float:
    if method == "robertson":
        inner = (n - df + 0.5) / (df + 0.5)
        return math.log(inner if inner >= 1 else 1)
    if method == "lucene":
        return math.log(1 + (n - df + 0.5) / (df + 0.5))
    if method == "atire":
        return math.log(n / df)
    raise ValueError(f"unsupported method {method!r}")


def _tfc(method, tf_array, l_d, l_avg, k1, b):
    if method in ("robertson", "lucene"):
        return tf_array / (k1 * ((1 - b) + b * l_d / l_avg) + tf_array)
    if method == "atire":
        return (tf_array * (k1 + 1)) / (tf_array + k1 * (1 - b + b * l_d / l_avg))
    raise ValueError(f"unsupported method {method!r}")


def build_csc(corpus_token_ids, n_vocab, k1=1.5, b=0.75, method="lucene"):
    """Per-document loop, as bm25s builds its score matrix.

    corpus_token_ids: list of 1-D int arrays/lists.  Returns dict with
    data fp32[nnz], indices int32[nnz], indptr int64[n_vocab+1], num_docs,
    doc_len int32[N], avgdl float.
    """
    n_docs = len(corpus_token_ids)
    doc_len = np.array([len(d) for d in corpus_token_ids])
    l_avg = doc_len.mean() if n_docs else np.float64(0.0)

    df = np.zeros(n_vocab, dtype=np.int64)
    counters = []
    for doc in corpus_token_ids:
        c = Counter(int(t) for t in doc)
        counters.append(c)
        for t in c:
            df[t] += 1

    idf = np.zeros(n_vocab, dtype=np.float32)
    for t in range(n_vocab):
        if df[t] > 0:
            idf[t] = _idf(method, int(df[t]), n_docs)

    nnz = int(sum(len(c) for c in counters))
    vals = np.empty(nnz, dtype=np.float32)
    rows = np.empty(nnz, dtype=np.int32)
    cols = np.empty(nnz, dtype=np.int32)
    pos = 0
    for d, c in enumerate(counters):
        voc = np.array(list(c.keys()), dtype=np.int32)
        tf = np.array(list(c.values()), dtype=np.float32)
        tfc = _tfc(method, tf, int(doc_len[d]), l_avg, k1, b)
        sc = idf[voc] * tfc
        m = len(voc)
        vals[pos:pos + m] = sc
        rows[pos:pos + m] = d
        cols[pos:pos + m] = voc
        pos += m

    # csc_matrix((vals,(rows,cols))) : stable counting sort by column, doc ids
    # ascending within a column
    order = np.lexsort((rows, cols))
    data = vals[order]
    indices = rows[order].astype(np.int32)
    indptr = np.zeros(n_vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(cols, minlength=n_vocab), out=indptr[1:])
    return {
        "data": data,
        "indices": indices,
        "indptr": indptr,
        "num_docs": n_docs,
        "doc_len": doc_len.astype(np.int32),
        "avgdl": float(l_avg),
    }


def get_scores_ids(scores: dict, term_ids) -> np.ndarray:
    """bm25s `_compute_relevance_from_scores` (numpy backend)."""
    acc = np.zeros(scores["num_docs"], dtype=np.float32)
    indptr, indices, data = scores["indptr"], scores["indices"], scores["data"]
    for t in term_ids:
        s, e = indptr[t], indptr[t + 1]
        np.add.at(acc, indices[s:e], data[s:e])
    return acc


def topk_canonical(values: np.ndarray, k: int):
    """top-k by (value desc, index asc)."""
    order = np.lexsort((np.arange(len(values)), -values.astype(np.float64)))[:k]
    return order.astype(np.int64), values[order]


def match_counts(scores: dict, term_ids) -> np.ndarray:
    """#distinct query terms whose column holds the doc (scorer.py:592-601)."""
    cnt = np.zeros(scores["num_docs"], dtype=np.int32)
    for t in dict.fromkeys(int(x) for x in term_ids):
        s, e = scores["indptr"][t], scores["indptr"][t + 1]
        cnt[scores["indices"][s:e]] += 1
    return cnt


class BM25:
    """Stand-in exposing the five-call surface, so the reference's scorer.py /
    multi_field.py run unmodified on top of it (golden generation, config-1
    literal CPU timing)."""

    def __init__(self, k1=1.5, b=0.75, delta=0.5, method="lucene", **_):
        if method not in ("robertson", "lucene", "atire"):
            raise ValueError(f"unsupported method {method!r}")
        self.k1, self.b, self.method = k1, b, method
        self.vocab_dict: dict[str, int] = {}
        self.scores: dict = {}

    def index(self, corpus_tokens, show_progress=True, **_):
        vocab: dict[str, int] = {}
        ids = []
        for doc in corpus_tokens:
            row = []
            for tok in doc:
                j = vocab.get(tok)
                if j is None:
                    j = vocab[tok] = len(vocab)
                row.append(j)
            ids.append(row)
        if "" not in vocab:
            vocab[""] = len(vocab)
        self.vocab_dict = vocab
        self.scores = build_csc(ids, len(vocab), self.k1, self.b, self.method)

    def get_tokens_ids(self, tokens):
        return [self.vocab_dict[t] for t in tokens if t in self.vocab_dict]

    def get_scores(self, query_tokens, **_):
        return get_scores_ids(self.scores, self.get_tokens_ids(query_tokens))

    def retrieve(self, query_tokens, k=10, sorted=True, show_progress=False, **_):
        n = self.scores["num_docs"]
        if k > n:
            raise ValueError(
                f"k of {k} is larger than the number of available scores, which is {n}"
            )
        docs = np.empty((len(query_tokens), k), dtype=np.int64)
        scs = np.empty((len(query_tokens), k), dtype=np.float32)
        for i, q in enumerate(query_tokens):
            docs[i], scs[i] = topk_canonical(self.get_scores(q), k)
        return Results(documents=docs, scores=scs)
