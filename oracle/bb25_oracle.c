/*
 * bb25_oracle.c -- CPU restatement of the Bayesian-BM25 query-time hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The shipped package (bayesian_bm25_b200/) never does.
 *
 * Each function names the reference lines it restates (paths relative to the
 * upstream cognica-io/bayesian-bm25 tree, v0.12.1).  The BM25 arithmetic itself
 * lives in the third-party `bm25s` package (pyproject.toml:20 "bm25s>=0.2.0",
 * unpinned, absent from this image): those parts restate its published
 * algorithm and are anchored on the reference's call sites
 * (scorer.py:213,262,306,525-529,583).  PARITY FOR BM25 SCORE VALUES IS
 * THEREFORE UNPINNED; everything after the score (posterior, fusion, bounds)
 * is pinned by tests/golden/ vectors generated from the reference's own
 * probability.py / fusion.py / scorer.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_EPS 1e-10 /* probability.py:20 */

/* ------------------------------------------------------------------ */
/* probability.py                                                      */
/* ------------------------------------------------------------------ */

/* probability.py:24-26 */
static inline double clamp_prob(double p) {
    if (p < ORC_EPS) return ORC_EPS;
    if (p > 1.0 - ORC_EPS) return 1.0 - ORC_EPS;
    return p;
}

/* probability.py:29-41 -- split-form sigmoid */
double orc_sigmoid(double x) {
    if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
    double e = exp(x);
    return e / (1.0 + e);
}

/* probability.py:44-48 */
double orc_logit(double p) {
    p = clamp_prob(p);
    return log(p / (1.0 - p));
}

/* probability.py:110-115 */
double orc_tf_prior(double tf) { return 0.2 + 0.7 * fmin(1.0, tf / 10.0); }

/* probability.py:117-129 */
double orc_norm_prior(double r) {
    return 0.3 + 0.6 * (1.0 - fmin(1.0, fabs(r - 0.5) * 2.0));
}

/* probability.py:131-140 */
double orc_composite_prior(double tf, double r) {
    double v = 0.7 * orc_tf_prior(tf) + 0.3 * orc_norm_prior(r);
    if (v < 0.1) v = 0.1;
    if (v > 0.9) v = 0.9;
    return v;
}

/* probability.py:142-169 -- two-step Bayes update */
double orc_posterior(double l, double p, int has_br, double br) {
    double num = l * p;
    double den = num + (1.0 - l) * (1.0 - p);
    double x = clamp_prob(num / den);
    if (has_br) {
        double nb = x * br;
        double db = nb + (1.0 - x) * (1.0 - br);
        x = clamp_prob(nb / db);
    }
    return x;
}

/* probability.py:106-108 */
double orc_likelihood(double alpha, double beta, double s) {
    return orc_sigmoid(alpha * (s - beta));
}

/*
 * probability.py:171-203.  prior_mode: 0 = composite prior, 1 = prior_free
 * (prior 0.5, :192-193), 2 = explicit prior array (the prior_fn branch
 * :194-199 after the Python callback has been evaluated; clamped here).
 */
void orc_score_to_probability(double alpha, double beta, int has_br, double br,
                              int prior_mode, const double *s, const double *tf,
                              const double *r, const double *prior, int64_t n,
                              double *out) {
    for (int64_t i = 0; i < n; i++) {
        double l = orc_likelihood(alpha, beta, s[i]);
        double p;
        if (prior_mode == 1) p = 0.5;
        else if (prior_mode == 2) p = clamp_prob(prior[i]);
        else p = orc_composite_prior(tf[i], r[i]);
        out[i] = orc_posterior(l, p, has_br, br);
    }
}

/* probability.py:205-236 */
void orc_wand_upper_bound(double alpha, double beta, int has_br, double br,
                          double p_max, const double *ub, int64_t n, double *out) {
    for (int64_t i = 0; i < n; i++)
        out[i] = orc_posterior(orc_likelihood(alpha, beta, ub[i]), p_max, has_br, br);
}

/* ------------------------------------------------------------------ */
/* fusion.py                                                           */
/* ------------------------------------------------------------------ */

/* fusion.py:25-45 */
void orc_cosine_to_probability(const double *c, int64_t n, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = clamp_prob((1.0 + c[i]) / 2.0);
}

/* fusion.py:119-169.  gating: 0 none, 1 relu, 2 swish, 3 gelu, 4 softplus */
static double gate(double x, int gating, double gbeta) {
    switch (gating) {
    case 1: return x > 0.0 ? x : 0.0;
    case 2: return x * orc_sigmoid(gbeta * x);
    case 3: return x * orc_sigmoid(1.702 * x);
    case 4: { /* np.logaddexp(0, b*x) / b */
        double y = gbeta * x;
        double m = y > 0.0 ? y : 0.0;
        return (m + log1p(exp(-fabs(y)))) / gbeta;
    }
    default: return x;
    }
}

/*
 * fusion.py:243-280.  `scale` = n ** effective_alpha, resolved by the caller
 * (fusion.py:260,270).  weights == NULL -> unweighted mean branch (:270-279).
 * Validation of the weights (:253-258) is the caller's job.
 */
void orc_log_odds_conjunction(const double *probs, int64_t m, int n,
                              const double *weights, double scale, int gating,
                              double gbeta, int has_max_logit, double max_logit,
                              double *out) {
    for (int64_t i = 0; i < m; i++) {
        double acc = 0.0;
        for (int j = 0; j < n; j++) {
            double x = gate(orc_logit(probs[i * n + j]), gating, gbeta);
            if (has_max_logit) {
                if (x < -max_logit) x = -max_logit;
                if (x > max_logit) x = max_logit;
            }
            acc += weights ? weights[j] * x : x;
        }
        double l = weights ? scale * acc : (acc / (double)n) * scale;
        out[i] = orc_sigmoid(l);
    }
}

/* ------------------------------------------------------------------ */
/* bm25s scoring (third-party; call sites scorer.py:306,525,583)        */
/* ------------------------------------------------------------------ */

/*
 * bm25s get_scores: zero fp32 accumulator, then for every in-vocabulary query
 * token id, IN QUERY ORDER AND WITH DUPLICATES, acc[indices[j]] += data[j] over
 * the token's CSC column.  fp32 adds, so order matters for >= 3 matches.
 */
void orc_get_scores(const float *data, const int32_t *indices,
                    const int64_t *indptr, int64_t n_docs, const int32_t *q_terms,
                    int n_terms, float *out) {
    memset(out, 0, (size_t)n_docs * sizeof(float));
    for (int i = 0; i < n_terms; i++) {
        int64_t s = indptr[q_terms[i]], e = indptr[q_terms[i] + 1];
        for (int64_t j = s; j < e; j++) {
            out[indices[j]] += data[j]; /* fp32 add (SSE; no excess precision) */
        }
    }
}

/*
 * scorer.py:592-601 `_compute_tf_batch`: len(set(query) & set(doc tokens)) =
 * number of DISTINCT query terms whose column contains the document.
 */
void orc_match_counts(const int32_t *indices, const int64_t *indptr,
                      int64_t n_docs, const int32_t *q_terms, int n_terms,
                      int32_t *out) {
    memset(out, 0, (size_t)n_docs * sizeof(int32_t));
    for (int i = 0; i < n_terms; i++) {
        int dup = 0;
        for (int j = 0; j < i; j++) dup |= (q_terms[j] == q_terms[i]);
        if (dup) continue;
        int64_t s = indptr[q_terms[i]], e = indptr[q_terms[i] + 1];
        for (int64_t j = s; j < e; j++) out[indices[j]] += 1;
    }
}

/* canonical rank key: score descending, doc id ascending */
static inline uint64_t rank_key_f32(float s, int64_t id) {
    uint32_t b;
    memcpy(&b, &s, 4);
    return ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)id);
}

static void heap_sift(uint64_t *h, int k, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < k && h[l] < h[m]) m = l;
        if (r < k && h[r] < h[m]) m = r;
        if (m == i) return;
        uint64_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

static int cmp_u64_desc(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

/*
 * bm25s selection.topk (argpartition + sort desc, via scorer.py:525-529) made
 * deterministic: ties broken by ascending doc id; zero-score documents fill
 * the tail when fewer than k documents match.  Scores must be >= 0.
 */
void orc_topk_f32(const float *scores, int64_t n, int k, int64_t *out_ids,
                  float *out_scores) {
    uint64_t *h = (uint64_t *)malloc((size_t)k * sizeof(uint64_t));
    int filled = 0;
    for (int64_t d = 0; d < n; d++) {
        uint64_t key = rank_key_f32(scores[d], d);
        if (filled < k) {
            h[filled++] = key;
            if (filled == k)
                for (int i = k / 2 - 1; i >= 0; i--) heap_sift(h, k, i);
        } else if (key > h[0]) {
            h[0] = key;
            heap_sift(h, k, 0);
        }
    }
    qsort(h, (size_t)filled, sizeof(uint64_t), cmp_u64_desc);
    for (int i = 0; i < filled; i++) {
        uint32_t b = (uint32_t)(h[i] >> 32);
        memcpy(&out_scores[i], &b, 4);
        out_ids[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(h[i] & 0xFFFFFFFFu));
    }
    free(h);
}

/* top-k of fp64 values (>= 0), value desc then id asc: multi_field.py:199 made
 * deterministic the same way */
typedef struct { double v; int64_t id; } dv_t;
static int cmp_dv(const void *a, const void *b) {
    const dv_t *x = (const dv_t *)a, *y = (const dv_t *)b;
    if (x->v != y->v) return x->v < y->v ? 1 : -1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}
void orc_topk_f64(const double *vals, int64_t n, int k, int64_t *out_ids,
                  double *out_vals) {
    dv_t *a = (dv_t *)malloc((size_t)n * sizeof(dv_t));
    for (int64_t i = 0; i < n; i++) { a[i].v = vals[i]; a[i].id = i; }
    qsort(a, (size_t)n, sizeof(dv_t), cmp_dv);
    for (int i = 0; i < k && i < n; i++) { out_ids[i] = a[i].id; out_vals[i] = a[i].v; }
    free(a);
}

typedef struct {
    double alpha, beta;
    int has_base_rate;
    double base_rate;
    int prior_mode;
} orc_params;

/*
 * scorer.py:603-640 `_scores_to_probabilities` for one row: 0.0 where score
 * <= 0, else score_to_probability(score, tf, doc_len/avgdl).
 */
static double doc_probability(const orc_params *p, float score, int32_t tf,
                              int32_t doc_len, double avgdl) {
    if (!(score > 0.0f)) return 0.0;
    double l = orc_likelihood(p->alpha, p->beta, (double)score);
    double prior = p->prior_mode == 1
                       ? 0.5
                       : orc_composite_prior((double)tf, (double)doc_len / avgdl);
    return orc_posterior(l, prior, p->has_base_rate, p->base_rate);
}

/* scorer.py:564-590 get_probabilities: dense fp64 [N] for one query */
void orc_get_probabilities(const float *data, const int32_t *indices,
                           const int64_t *indptr, int64_t n_docs,
                           const int32_t *doc_len, double avgdl,
                           const orc_params *p, const int32_t *q_terms,
                           int n_terms, double *out) {
    float *acc = (float *)malloc((size_t)n_docs * sizeof(float));
    int32_t *cnt = (int32_t *)malloc((size_t)n_docs * sizeof(int32_t));
    orc_get_scores(data, indices, indptr, n_docs, q_terms, n_terms, acc);
    orc_match_counts(indices, indptr, n_docs, q_terms, n_terms, cnt);
    for (int64_t d = 0; d < n_docs; d++)
        out[d] = doc_probability(p, acc[d], cnt[d], doc_len[d], avgdl);
    free(acc);
    free(cnt);
}

/*
 * scorer.py:494-536 retrieve for a batch: per query get_scores -> top-k by
 * fp32 score -> probabilities attached in rank order (not re-sorted).
 * Queries are independent; a pthread pool pulls queries from a shared counter
 * (`n_threads` <= 0: one per online core).  Returns the thread count used.
 */
typedef struct {
    const float *data; const int32_t *indices; const int64_t *indptr;
    int64_t n_docs; const int32_t *doc_len; double avgdl; const orc_params *p;
    const int32_t *q_terms; const int64_t *q_off; int64_t n_q; int k;
    int64_t *out_ids; float *out_scores; double *out_probs;
    int64_t next; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    float *acc = (float *)malloc((size_t)j->n_docs * sizeof(float));
    int32_t *cnt = (int32_t *)malloc((size_t)j->n_docs * sizeof(int32_t));
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t q = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (q >= j->n_q) break;
        const int32_t *qt = j->q_terms + j->q_off[q];
        int m = (int)(j->q_off[q + 1] - j->q_off[q]);
        int k = j->k;
        orc_get_scores(j->data, j->indices, j->indptr, j->n_docs, qt, m, acc);
        orc_match_counts(j->indices, j->indptr, j->n_docs, qt, m, cnt);
        int64_t *ids = j->out_ids + q * k;
        float *sc = j->out_scores + q * k;
        orc_topk_f32(acc, j->n_docs, k, ids, sc);
        for (int r = 0; r < k; r++)
            j->out_probs[q * k + r] = doc_probability(
                j->p, sc[r], cnt[ids[r]], j->doc_len[ids[r]], j->avgdl);
    }
    free(acc);
    free(cnt);
    return NULL;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int orc_retrieve_batch(const float *data, const int32_t *indices,
                       const int64_t *indptr, int64_t n_docs,
                       const int32_t *doc_len, double avgdl, const orc_params *p,
                       const int32_t *q_terms, const int64_t *q_off, int64_t n_q,
                       int k, int n_threads, int64_t *out_ids, float *out_scores,
                       double *out_probs) {
    if (n_threads <= 0) n_threads = orc_max_threads();
    if (n_threads > n_q) n_threads = n_q > 0 ? (int)n_q : 1;
    batch_job j = {data, indices, indptr, n_docs, doc_len, avgdl, p, q_terms,
                   q_off, n_q, k, out_ids, out_scores, out_probs, 0,
                   PTHREAD_MUTEX_INITIALIZER};
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &j);
    batch_worker(&j);
    for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return n_threads;
}

/* ------------------------------------------------------------------ */
/* BlockMaxIndex (scorer.py:33-142)                                     */
/* ------------------------------------------------------------------ */

/* scorer.py:55-81: block_maxes[t,b] = max(score_matrix[t, b*bs:(b+1)*bs]) */
void orc_blockmax_dense(const double *sm, int64_t n_terms, int64_t n_docs, int bs,
                        double *out) {
    int64_t nb = (n_docs + bs - 1) / bs;
    for (int64_t t = 0; t < n_terms; t++)
        for (int64_t b = 0; b < nb; b++) {
            int64_t s = b * bs, e = s + bs < n_docs ? s + bs : n_docs;
            double m = sm[t * n_docs + s];
            for (int64_t d = s + 1; d < e; d++)
                if (sm[t * n_docs + d] > m) m = sm[t * n_docs + d];
            out[t * nb + b] = m;
        }
}

/* the same table for a list of terms taken from the CSC (absent posting = 0.0;
 * posting values are >= 0) */
void orc_blockmax_csc(const float *data, const int32_t *indices,
                      const int64_t *indptr, int64_t n_docs, const int32_t *terms,
                      int n_terms, int bs, float *out) {
    int64_t nb = (n_docs + bs - 1) / bs;
    memset(out, 0, (size_t)(n_terms * nb) * sizeof(float));
    for (int t = 0; t < n_terms; t++)
        for (int64_t j = indptr[terms[t]]; j < indptr[terms[t] + 1]; j++) {
            int64_t b = indices[j] / bs;
            if (data[j] > out[t * nb + b]) out[t * nb + b] = data[j];
        }
}

/*
 * Multi-shard merge (no reference counterpart): S per-shard lists
 * [S,Q,k] sorted by (score desc, id asc) -> global [Q,k] under the same order.
 */
void orc_merge_topk(const int64_t *ids, const float *scores, const double *probs,
                    int n_shards, int64_t n_q, int k, int64_t *out_ids,
                    float *out_scores, double *out_probs) {
    int *pos = (int *)malloc((size_t)n_shards * sizeof(int));
    for (int64_t q = 0; q < n_q; q++) {
        memset(pos, 0, (size_t)n_shards * sizeof(int));
        for (int r = 0; r < k; r++) {
            int best = -1;
            uint64_t bk = 0;
            for (int s = 0; s < n_shards; s++) {
                if (pos[s] >= k) continue;
                int64_t o = ((int64_t)s * n_q + q) * k + pos[s];
                uint64_t key = rank_key_f32(scores[o], ids[o]);
                if (best < 0 || key > bk) { best = s; bk = key; }
            }
            int64_t o = ((int64_t)best * n_q + q) * k + pos[best]++;
            out_ids[q * k + r] = ids[o];
            out_scores[q * k + r] = scores[o];
            out_probs[q * k + r] = probs[o];
        }
    }
    free(pos);
}
