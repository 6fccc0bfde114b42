/*
 * bb25.h -- C ABI of libbb25.so, the B200 (sm_100a) implementation of the
 * Bayesian-BM25 query-time hot path.
 *
 * The reference (cognica-io/bayesian-bm25 v0.12.1) is pure Python + NumPy and has
 * no FFI of its own; its seam for this path is the five calls it makes into the
 * third-party `bm25s` package plus plain NumPy functions.  Each entry point
 * below names the reference interface (file:line, relative to the upstream
 * tree) it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * would add on the reference side.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    bb25_last_error() returns a thread-local message for the last failure.
 *  - "dev" pointers are CUDA device pointers on the index's device, "host"
 *    pointers are ordinary host memory, "any" may be either (resolved through
 *    unified addressing).  All buffers are caller-owned.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - There is no CPU fallback: without a usable CUDA device every compute call
 *    fails with an error.
 *  - An index handle is immutable after creation; query calls on ONE handle
 *    are serialised internally on the host and share a device workspace, so
 *    concurrent calls on one handle must use the same stream (the single-query
 *    calls are asynchronous: they return once the work is enqueued).
 *  - Doc ids are int64 in the interface; one index (= one shard) holds at most
 *    2^29 documents.  Scores are fp32, probabilities fp64, as in the reference.
 */
#ifndef BB25_H
#define BB25_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB25_VERSION 100

typedef struct bb25_index bb25_index;

/* BayesianProbabilityTransform state used at inference time
 * (bayesian_bm25/probability.py:71-94). */
typedef struct bb25_params {
    double alpha;      /* sigmoid steepness   (probability.py:82) */
    double beta;       /* sigmoid midpoint    (probability.py:83) */
    int has_base_rate; /* base_rate is not None (probability.py:84) */
    double base_rate;
    int prior_mode;    /* 0 composite prior (probability.py:201),
                          1 prior_free -> 0.5  (probability.py:192-193) */
} bb25_params;

/* gating names of fusion.py:156-169 */
enum { BB25_GATE_NONE = 0, BB25_GATE_RELU = 1, BB25_GATE_SWISH = 2, BB25_GATE_GELU = 3, BB25_GATE_SOFTPLUS = 4 };

const char *bb25_last_error(void);
int bb25_version(void);
/* number of CUDA kernels this library has launched since it was loaded */
unsigned long long bb25_launch_count(void);
/* number of CUDA devices visible (0 when there is none / no driver) */
int bb25_device_count(void);
/* Measured ceilings for the roofline lines (no reference counterpart): best-of-`reps` bandwidth of a
 * streaming 128-bit read of a `bytes`-sized device buffer, read `iters` times per launch by every SM.
 * bytes <= ~64 MB stays resident in the 126 MB L2 (L2 -> SM bandwidth); bytes >> 126 MB reads HBM. */
int bb25_measure_read_bandwidth(int device, int64_t bytes, int iters, int reps, double *out_gbs,
                                double *out_ms);

/* ---- index ------------------------------------------------------------- */

/*
 * Upload one shard of a bm25s-style CSC score matrix: replaces the state
 * bm25s.BM25.index() leaves behind for the reference (scorer.py:262, read back
 * through .scores at scorer.py:227) together with the reference's own
 * doc_lengths / avgdl (scorer.py:264-267).
 *   data[nnz]      fp32 posting values (idf*tfc)                      (any)
 *   indices[nnz]   int32 LOCAL doc ids, ascending inside a column      (any)
 *   indptr[V+1]    int64 column starts                                 (any)
 *   doc_len[N]     int32 token counts of the shard's documents         (any)
 *   avgdl          GLOBAL mean document length
 *   doc_id_offset  added to local ids in every returned id (sharding)
 * Builds the per-(term, doc-tile) skip table used by the traversal kernel.
 */
int bb25_index_create(int device, int64_t n_docs, int64_t n_vocab, int64_t nnz,
                      const float *data, const int32_t *indices, const int64_t *indptr,
                      const int32_t *doc_len, double avgdl, int64_t doc_id_offset,
                      bb25_index **out);
void bb25_index_destroy(bb25_index *idx);
int bb25_index_info(const bb25_index *idx, int64_t *n_docs, int64_t *n_vocab, int64_t *nnz,
                    int *tile_docs, int *n_tiles, int64_t *device_bytes);
/* The per-(term, 1024-document block) table behind block-max pruning (BlockMaxIndex semantics,
 * scorer.py:55-99) keeps a dense row per term while that fits a memory budget; for larger
 * vocabularies the terms that touch few blocks are stored as a bitmap plus compact entries.
 * bitmap_terms: how many terms took that form; table_bytes: device bytes of the table.
 * (BB25_TAB_SPARSE=0/1/2 at index creation forces none / wherever smaller / all: test hook.) */
int bb25_index_table_info(const bb25_index *idx, int64_t *bitmap_terms, int64_t *table_bytes);

/* ---- a1/a4: dense per-query outputs -------------------------------------- */

/* bm25s BM25.get_scores as called at scorer.py:306 and :583: fp32 [N], adds in
 * query-term order, duplicates included.  q_terms: host, in-vocabulary ids. */
int bb25_get_scores(bb25_index *idx, const int32_t *q_terms, int n_terms,
                    float *out_scores /*dev [N]*/, void *stream);

/* BayesianBM25Scorer.get_probabilities (scorer.py:564-590): fp64, exactly 0.0
 * where the score is <= 0.  Element d is written at out_probs[d*out_stride]
 * (out_stride >= 1) so several fields can be column-stacked as
 * multi_field.py:158-161 does. */
int bb25_get_probabilities(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                           int n_terms, double *out_probs /*dev*/, int64_t out_stride,
                           void *stream);

/* ---- a2/a3/a5/a6: batched top-k retrieval ------------------------------- */

/*
 * BayesianBM25Scorer.retrieve (scorer.py:494-536) for a batch of queries:
 * bm25s retrieve (scores + top-k by fp32 score, scorer.py:525-529) followed by
 * _scores_to_probabilities (scorer.py:603-640) with tf = number of distinct
 * query terms matched (scorer.py:592-601).  Rank order is (score desc, doc id
 * asc); when fewer than k documents match, the lowest-id zero-score documents
 * fill the tail with probability 0.0.
 *   q_terms[q_off[Q]]  int32 in-vocabulary term ids, query order, duplicates kept (dev)
 *   q_off[Q+1]         int64 offsets                                            (dev)
 *   out_ids[Q*k] int64, out_scores[Q*k] fp32 (may be NULL), out_probs[Q*k] fp64 (dev)
 * Requires 1 <= k <= min(n_docs, 4096).  Synchronises `stream` before returning.
 * Scores are bm25s's fp32 sums bit for bit (terms added in query order): the
 * traversal may sum in another order to find candidates, but every candidate is
 * re-scored in query order before it is ranked.
 */
int bb25_retrieve_batch(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                        const int64_t *q_off, int64_t n_queries, int k, int64_t *out_ids,
                        float *out_scores, double *out_probs, void *stream);

/* The same call when the caller knows the batch's extent on the host (q_off[0] and
 * q_off[Q] - q_off[0]): nothing is read back before the work is enqueued, and the whole batch --
 * threshold repairs included -- costs ONE stream synchronisation, at the end (the report that
 * carries the input-validation flag and the statistics).  bb25_retrieve_batch reads the two
 * offsets from the device first (one more synchronisation). */
int bb25_retrieve_batch_ex(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                           const int64_t *q_off, int64_t n_queries, int64_t term_base, int64_t n_terms_total,
                           int k, int64_t *out_ids, float *out_scores, double *out_probs, void *stream);

/* Same call with HOST buffers: copies the queries in, runs the batch, copies
 * the results out (the end-to-end path the Python wrapper uses).  The device staging area and
 * the stream are owned by the handle and reused across calls; with page-locked caller buffers
 * the copies are asynchronous and the call synchronises twice (report, results). */
int bb25_retrieve_batch_host(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                             const int64_t *q_off, int64_t n_queries, int k, int64_t *out_ids,
                             float *out_scores, double *out_probs);

/* One query through the dense guaranteed path: fp32 scores of every document (bm25s get_scores,
 * scorer.py:583), exact top-k of the whole vector by (score desc, doc id asc), probabilities from
 * the dense posterior pass (scorer.py:603-640).  O(N); 1 <= k <= min(n_docs, 8192).  This is also
 * what bb25_retrieve_batch falls back to for a query whose threshold refinement does not settle.
 * q_terms: host. */
int bb25_retrieve_one_dense(bb25_index *idx, const bb25_params *params, const int32_t *q_terms, int n_terms, int k,
                            int64_t *out_ids /*dev [k]*/, float *out_scores /*dev [k] or NULL*/,
                            double *out_probs /*dev [k]*/, void *stream);

/* host synchronisations inside the last batch, queries that needed the host-driven repair after the
 * enqueued repair rounds, queries that took the dense guaranteed path */
int bb25_retrieve_sync_stats(const bb25_index *idx, int64_t *host_syncs, int64_t *repaired_queries,
                             int64_t *dense_fallback_queries);

/* Sharded retrieval, cross-shard thresholds.  When `fn` is set, the batch publishes -- per query,
 * after every block group but the last -- the scores at a few ranks of the shard's running top-k
 * (bb25_quantile_ranks: ceil(k/S), ceil(2k/S), ceil(4k/S), k) into d_quant [n_queries][n_levels] (uint64,
 * fp32 score bits << 33, 0 = not known) and calls `fn` on the host, with the group's work already enqueued on
 * `stream`.  `fn` all-gathers d_quant across the ranks and hands the result to bb25_apply_quantiles, which
 * raises d_thr to the largest score for which the shards' published counts add up to k (a lower bound of the
 * GLOBAL k-th score), so every shard prunes and emits against the global bound.  Everything is stream-ordered;
 * `fn` must not synchronise.  Return non-zero to fail the batch. */
typedef int (*bb25_exchange_fn)(void *user, void *d_quant, void *d_thr, int64_t n_queries, int n_levels, int k,
                                int group, void *stream);
int bb25_index_set_threshold_exchange(bb25_index *idx, bb25_exchange_fn fn, void *user, int n_shards);
void bb25_quantile_ranks(int k, int n_shards, int *n_levels, int *ranks4);
/* all_quantiles: dev uint64 [n_shards][n_queries][n_levels], rank-major as an all-gather leaves it */
int bb25_apply_quantiles(int device, const void *all_quantiles, int n_shards, int64_t n_queries, int k, void *d_thr,
                         void *stream);

/* Threshold seeds.  bb25_index_kth_values: per term, the k-th largest posting value of THIS index (0 where the
 * term has fewer than k postings) -- max over a query's terms bounds the query's k-th best score from below;
 * the pointer (dev fp32 [n_vocab]) stays valid while the handle lives (a bounded cache keyed by k).
 * bb25_index_set_kth_values replaces the cached values for k: a sharded deployment installs the bound that
 * follows from ALL shards' values (k-th of the union >= any valid combination of per-shard ranks), so that a
 * shard starts a batch with the global seed instead of its own, looser one.  Any lower bound is valid. */
int bb25_index_kth_values(bb25_index *idx, int k, const float **out_dev, void *stream);
int bb25_index_set_kth_values(bb25_index *idx, int k, const float *values_dev, void *stream);

/* Exchange + merge in ONE kernel over peer memory (NVLink): this rank merges the queries
 * [q_begin, q_begin + n_queries).  src_tab_dev / dst_tab_dev: DEVICE arrays of n_shards pointers (symmetric
 * memory: entry s = rank s's buffer mapped into this process); every source buffer holds that shard's packed
 * [Q][k][2] lists (bb25_pack_topk), every destination buffer receives the merged packed rows of these queries.
 * The caller separates it from the producers / consumers of the buffers with device-side barriers. */
int bb25_merge_topk_peers(int device, const void *src_tab_dev, const void *dst_tab_dev, int n_shards, int64_t q_begin,
                          int64_t n_queries, int k, void *stream);
int bb25_unpack_topk(int device, const int64_t *packed /*dev [n][2]*/, int64_t n, int64_t *out_ids, float *out_scores,
                     double *out_probs, void *stream);
int bb25_memcpy_device(int device, void *dst, const void *src, int64_t bytes, void *stream);

/* statistics of the last bb25_retrieve_batch on this handle: kernel launches,
 * traversal passes (1 per tile group + re-runs), re-run (query,group) units,
 * candidates emitted. */
int bb25_retrieve_stats(const bb25_index *idx, int64_t *launches, int64_t *passes,
                        int64_t *rerun_queries, int64_t *candidates);

/* Dynamic pruning (north star item 4).  The batch traversal works on 1024-document
 * blocks with per-(block, term) maxima (BlockMaxIndex semantics, scorer.py:55-99).
 *   level 0: exhaustive traversal;
 *   level 1: a (block, query) unit whose summed block maxima stay below the query's
 *            current top-k threshold cannot contribute and is skipped;
 *   level 2: additionally, in the remaining units, when the frequent terms' block
 *            maxima alone cannot reach the threshold, documents matching only those terms
 *            are not evaluated: the frequent terms' values are added only where another
 *            query term has a posting (MaxScore's non-essential terms, per block);
 *   level 3 (default): additionally, a query whose threshold exceeds the summed GLOBAL
 *            maxima of its frequent terms leaves the block traversal altogether: only the
 *            postings of its remaining (rare) terms are walked and each such document is
 *            scored completely by one thread (bb25_retrieve_route_stats: how many queries
 *            went that way, and their 1024-posting work items).
 * Results are bit-identical at every level.  bb25_retrieve_prune_stats: units visited /
 * skipped / evaluated with the level-2 restriction in the last batch. */
int bb25_index_set_pruning(bb25_index *idx, int level);
int bb25_retrieve_prune_stats(const bb25_index *idx, int64_t *units, int64_t *units_skipped,
                              int64_t *units_maxscore);
int bb25_retrieve_route_stats(const bb25_index *idx, int64_t *routed_queries, int64_t *work_items);
/* Levels >= 2 also evaluate (block, query) units through their ESSENTIAL postings (MaxScore's split per
 * 1024-document block: the terms with a value row and many postings in the block whose summed block
 * maxima stay below the threshold are non-essential; the <= 32 documents of the other slices are
 * evaluated one by one, non-essential values read from the rows).  Units handled that way in the last batch: */
int bb25_retrieve_sparse_units(const bb25_index *idx, int64_t *units_sparse);

/* Device time of the traversal kernel in the last bb25_retrieve_batch on this handle,
 * measured with CUDA events on the call's stream around every traversal launch
 * (sum over launches, and their number).  Used for the roofline figure. */
int bb25_retrieve_timing(const bb25_index *idx, double *traverse_ms, int64_t *traverse_launches);

/* K7 (no reference counterpart): merge S per-shard [Q,k] lists (global ids,
 * each sorted by (score desc, id asc)) into the global [Q,k].  All dev. */
int bb25_merge_topk(int device, const int64_t *ids, const float *scores, const double *probs,
                    int n_shards, int64_t n_queries, int k, int64_t *out_ids, float *out_scores,
                    double *out_probs, void *stream);

/* The same merge for lists packed as 16-byte entries {(fp32 score bits << 32) | (2^32-1 - id),
 * fp64 probability bits}: ONE all-gather of [Q,k,2] int64 per rank instead of three tensors.
 * bb25_pack_topk builds the entries from a shard's (ids, scores, probs). */
int bb25_pack_topk(int device, const int64_t *ids, const float *scores, const double *probs, int64_t n,
                   int64_t *out_packed /*dev [n][2]*/, void *stream);
int bb25_merge_topk_packed(int device, const int64_t *packed /*dev [S][Q][k][2]*/, int n_shards,
                           int64_t n_queries, int k, int64_t *out_ids, float *out_scores,
                           double *out_probs, void *stream);

/* top-k of a dense fp64 vector (values >= 0), (value desc, index asc):
 * MultiFieldScorer.retrieve's argsort (multi_field.py:199) made deterministic. */
int bb25_topk_f64(int device, const double *vals /*dev [n]*/, int64_t n, int k,
                  int64_t *out_ids /*dev [k]*/, double *out_vals /*dev [k]*/, void *stream);

/* ---- a7-a11: BayesianProbabilityTransform, elementwise, all dev fp64 ------ */

int bb25_sigmoid(int device, const double *x, int64_t n, double *out, void *stream);              /* probability.py:29-41 */
int bb25_logit(int device, const double *p, int64_t n, double *out, void *stream);                /* probability.py:44-48 */
int bb25_likelihood(int device, const bb25_params *p, const double *score, int64_t n, double *out, void *stream); /* :106-108 */
int bb25_tf_prior(int device, const double *tf, int64_t n, double *out, void *stream);            /* :110-115 */
int bb25_norm_prior(int device, const double *ratio, int64_t n, double *out, void *stream);       /* :117-129 */
int bb25_composite_prior(int device, const double *tf, const double *ratio, int64_t n, double *out, void *stream); /* :131-140 */
int bb25_posterior(int device, const double *lik, const double *prior, int has_base_rate,
                   double base_rate, int64_t n, double *out, void *stream);                       /* :142-169 */
/* score_to_probability (:171-203).  prior == NULL: params->prior_mode decides;
 * prior != NULL: explicit per-element prior (the prior_fn branch :194-199,
 * evaluated by the caller), clamped to [1e-10, 1-1e-10]. */
int bb25_score_to_probability(int device, const bb25_params *p, const double *score,
                              const double *tf, const double *ratio, const double *prior,
                              int64_t n, double *out, void *stream);
int bb25_wand_upper_bound(int device, const bb25_params *p, const double *bm25_ub, double p_max,
                          int64_t n, double *out, void *stream);                                  /* :205-236 */

/* ---- a13/a14: fusion ------------------------------------------------------- */

/* cosine_to_probability (fusion.py:25-45); element i written at out[i*out_stride] */
int bb25_cosine_to_probability(int device, const double *cos, int64_t n, double *out,
                               int64_t out_stride, void *stream);
/* log_odds_conjunction (fusion.py:172-280) over rows of probs[m][n].
 * weights: dev [n] or NULL (unweighted mean branch).  scale = n ** alpha with
 * alpha already resolved by the caller (fusion.py:106-116,260,270).
 * Weight validation (fusion.py:253-258) is the caller's. */
int bb25_log_odds_conjunction(int device, const double *probs, int64_t m, int n,
                              const double *weights, double scale, int gating,
                              double gating_beta, int has_max_logit, double max_logit,
                              double *out, void *stream);

/*
 * Fused form of the same conjunction for dense, per-document signals (ungated, no
 * max_logit): one accumulator acc[N] (dev fp64) is updated in place, one call per
 * signal in signal order, so neither the per-signal probability vectors nor the
 * column-stacked [N, n] matrix of multi_field.py:158-161 are ever materialised:
 *     acc = [acc +] w * logit(clamp(p_signal))            flags & 1: first signal
 *     acc = sigmoid(scale * acc)  (or sigmoid(acc / n * scale) when flags & 4) on the
 *                                                         flags & 2: last signal
 * bb25_fuse_bm25_signal evaluates the signal inside the traversal kernel's epilogue
 * (BM25 accumulate -> posterior -> logit -> weighted add in one pass over the index);
 * p = 0.0 for documents the query does not match (scorer.py:618, fusion.py:243).
 * bb25_fuse_cosine_signal takes fp32 cosine similarities (cosine_to_probability,
 * fusion.py:43-45), bb25_fuse_prob_signal ready-made fp64 probabilities.
 */
int bb25_fuse_bm25_signal(bb25_index *idx, const bb25_params *params, const int32_t *q_terms /*host*/,
                          int n_terms, double weight, int n_signals, double scale, int flags,
                          double *acc /*dev [N]*/, void *stream);
int bb25_fuse_cosine_signal(int device, const float *cosine /*dev*/, int64_t n, double weight, int n_signals,
                            double scale, int flags, double *acc, void *stream);
int bb25_fuse_prob_signal(int device, const double *probs /*dev*/, int64_t n, double weight, int n_signals,
                          double scale, int flags, double *acc, void *stream);

/* ---- a15 + configs 4/5: batched top-k by FUSED probability ------------------------- */

/* One BM25 field (= one index handle) of a fused query batch.  The queries of the batch are given
 * per field because every field has its own vocabulary (multi_field.py:105-139). */
typedef struct bb25_fused_field {
    bb25_index *index;
    bb25_params params;      /* the field's transform (alpha must be > 0) */
    double weight;           /* conjunction weight of the field; ignored when `weighted` is 0 */
    const int32_t *q_terms;  /* dev: in-vocabulary term ids of the field, all queries back to back */
    const int64_t *q_off;    /* dev [Q+1] */
    int64_t term_base;       /* host copies of q_off[0] and q_off[Q] - q_off[0] */
    int64_t n_terms_total;
} bb25_fused_field;

/*
 * MultiFieldScorer.retrieve (multi_field.py:176-200) and the hybrid BM25 + vector pattern
 * (benchmarks/hybrid_beir.py:1708-1765, README.md:129-140) for a BATCH of queries: top-k documents by
 *     fused = log_odds_conjunction([p_field_0, ..., p_field_F-1 (, cosine_to_probability(cos))], alpha, weights)
 * (fusion.py:172-280, gating "none"), ranked (fused desc, doc id asc).  p_field_i is the field's
 * posterior (scorer.py:564-590), 0.0 where the field's BM25 score is <= 0.
 *   n_fields 1..4; all fields index the same documents.
 *   cosine: dev fp32 [Q][cos_stride] or NULL; rows 16-byte aligned, cos_stride % 4 == 0, >= n_docs.
 *           The dense signal is the LAST signal of the conjunction.
 *   weighted != 0: weights (fields[i].weight..., cos_weight), validated by the caller as fusion.py:253-258;
 *   weighted == 0: unweighted mean branch (fusion.py:270-279).
 *   scale = n_signals ** alpha, alpha resolved by the caller (fusion.py:106-116).
 *   out_ids int64 [Q][k], out_probs fp64 [Q][k] (dev).  1 <= k <= min(n_docs, 1024).
 * The traversal prunes (block, query) units with the per-(term, block) maxima pushed through each
 * field's probability bound and the monotone conjunction (BlockMaxIndex.bayesian_block_upper_bound,
 * scorer.py:101-130; wand_upper_bound, probability.py:205-236) unless the first field's index is at
 * pruning level 0; every candidate is evaluated exactly before it is ranked, so results do not depend
 * on the level.  One stream synchronisation per batch (more only when queries take the dense path).
 */
int bb25_retrieve_fused_batch(int n_fields, const bb25_fused_field *fields, const float *cosine, int64_t cos_stride,
                              double cos_weight, int weighted, double scale, int64_t n_queries, int k,
                              int64_t *out_ids, double *out_probs, void *stream);
/* statistics of the last bb25_retrieve_fused_batch whose first field was `idx`: (block, query) units
 * handed out, skipped by the block-max bound, abandoned between fields (no document's running bound could
 * still reach the threshold), candidates evaluated exactly, queries on the dense guaranteed path,
 * re-run (query, group) pairs, host synchronisations, summed traversal-kernel time (CUDA events) */
int bb25_fused_stats(const bb25_index *idx, int64_t *units, int64_t *units_skipped, int64_t *units_abandoned,
                     int64_t *candidates, int64_t *fallback_queries, int64_t *rerun_queries, int64_t *host_syncs,
                     double *traverse_ms);
/* essential-posting evaluation (two fields, no dense signal, pruning level >= 2: MaxScore's split --
 * the partition behind wand_upper_bound, probability.py:205-236 -- per 1024-document block): units
 * skipped because no essential slice has a posting in the block, units evaluated through their essential
 * postings instead of a pass over the block, documents evaluated that way */
int bb25_fused_prune_stats(const bb25_index *idx, int64_t *units_no_essential, int64_t *units_sparse,
                           int64_t *sparse_documents);

/* ---- query-time consumers of the probabilities ------------------- */

/* tf of returned documents: out_tf[q*k + r] = number of DISTINCT terms of query q whose posting list holds
 * document ids[q*k + r] (global id); scorer.py:592-601 as used by retrieve(explain=True), scorer.py:546-552.
 * q_terms / q_off / ids / out_tf: dev. */
int bb25_match_counts(bb25_index *idx, const int32_t *q_terms, const int64_t *q_off, int64_t n_queries, int k,
                      const int64_t *ids, int32_t *out_tf, void *stream);
/* FusionDebugger.trace_bm25 (debug.py:178-216) for n (score, tf, doc_len_ratio) triples: out[i*7 ..] =
 * likelihood, tf_prior, norm_prior, composite_prior, logit(likelihood), logit(composite_prior), posterior. */
int bb25_trace_bm25(int device, const bb25_params *p, const double *score, const double *tf, const double *ratio,
                    int64_t n, double *out /*dev [n][7]*/, void *stream);
/* AttentionLogOddsWeights._compute_weights (fusion.py:757-772): softmax(query_features @ W^T + b) per row.
 * query_features [m][n_features], W [n_signals][n_features], b [n_signals], out [m][n_signals]; all dev fp64. */
int bb25_attention_weights(int device, const double *query_features, const double *W, const double *b, int64_t m,
                           int n_features, int n_signals, double *out, void *stream);
/* AttentionLogOddsWeights.__call__ / compute_upper_bounds (fusion.py:774-828, 1039-1082) over m candidates:
 * sigmoid(scale * sum_i w_i * x_i [+ logit_base_rate]), x = logit(clamp(p)), optionally min-max normalised per
 * signal over the m candidates (normalize != 0).  weights: [1][n] (one query) or [m][n] (one row per candidate). */
int bb25_attention_fuse(int device, const double *probs /*dev [m][n]*/, int64_t m, int n_signals, const double *weights,
                        int64_t n_weight_rows, double scale, int has_base_rate, double logit_base_rate, int normalize,
                        double *out /*dev [m]*/, void *stream);
/* balanced_log_odds_fusion (fusion.py:283-333): weight * minmax(logit(cosine_to_probability(dense))) +
 * (1 - weight) * minmax(logit(clamp(sparse))), min-max over the n candidates. */
int bb25_balanced_fusion(int device, const double *sparse_probs, const double *dense_similarities, int64_t n, double weight,
                         double *out, void *stream);

/* Dense side of hybrid retrieval (benchmarks/hybrid_beir.py:1751-1753: query_emb @ corpus_emb.T): cosine
 * similarities of up to 256 queries against every document, on the tensor cores (TMA + tcgen05.mma, fp32
 * accumulation in tensor memory).  query_emb bf16 [n_queries][k], corpus_emb bf16 [n_docs][k] (dev, 16-byte
 * aligned, rows L2-normalised by the caller, k a multiple of 64); out fp32 [n_queries][out_stride] (dev) --
 * the cosine-row layout of bb25_retrieve_fused_batch. */
int bb25_cosine_gemm(int device, const void *query_emb, int n_queries, const void *corpus_emb, int64_t n_docs, int k,
                     float *out, int64_t out_stride, void *stream);

/* ---- a12: BlockMaxIndex ---------------------------------------------------- */

/* BlockMaxIndex.build (scorer.py:55-81) on a dense [n_terms][n_docs] fp64
 * matrix -> [n_terms][ceil(n_docs/block_size)] fp64.  All dev. */
int bb25_blockmax_dense(int device, const double *score_matrix, int64_t n_terms, int64_t n_docs,
                        int block_size, double *out, void *stream);
/* The same table for `n_terms` columns of the index (absent posting = 0.0). */
int bb25_blockmax_csc(bb25_index *idx, const int32_t *terms /*dev*/, int n_terms, int block_size,
                      float *out /*dev [n_terms][n_blocks]*/, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BB25_H */
