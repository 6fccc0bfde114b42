// Batched top-k by FUSED probability (BASELINE configs 4 and 5): the rank key is
//     fused(d) = sigma(n^alpha * sum_i w_i * logit(clamp(p_i(d))))          fusion.py:243-268
// over F BM25 fields (MultiFieldScorer, multi_field.py:141-200; p_i = the field's Bayesian posterior,
// 0.0 where the field's BM25 score is <= 0, scorer.py:618) and, optionally, one dense signal
// p = cosine_to_probability(cos(q, d)) (fusion.py:25-45; the hybrid pattern of
// benchmarks/hybrid_beir.py:1708-1765).  Ranking is (fused desc, doc id asc).
//
// How the batch is evaluated
//  * Traversal (fused_block_kernel), warp-private 1024-document blocks like block_kernel: per field the
//    terms' posting slices are scattered into the warp's fp32 accumulators (dense value rows of head
//    terms are added in registers), and a fold pass turns the field's score s_i into a cheap fp32 UPPER
//    BOUND of w_i * (logit(p_i) + 23.03):
//        logit(posterior) = alpha*(s - beta) + logit(prior) + logit(base_rate)      (Bayes in odds form)
//                         <= alpha*s + c_iq,   prior <= composite_prior(tf = #distinct query terms, P_norm = 0.9)
//    i.e. the bound wand_upper_bound (probability.py:205-236) gives, with the tf the query allows instead
//    of p_max = 0.9.  The per-document sum V(d) over the fields (+ the dense signal's exact-to-1e-3 term)
//    is compared with the query's threshold; survivors are emitted as candidates.
//  * Block-max pruning (BlockMaxIndex.bayesian_block_upper_bound, scorer.py:101-130, applied to the fused
//    key as SURVEY 7 describes): the same bound evaluated at the per-(term, block) maxima bounds every
//    document of the block because the conjunction is monotone in every signal; a unit below the
//    threshold is skipped.
//  * Essential-posting evaluation (two fields, no dense signal, pruning level >= 2; fused_group_kernel, a
//    light kernel that runs before fused_block_kernel): MaxScore per block on the fused key -- the entries
//    whose bound alone cannot reach the threshold are non-essential, and a unit whose other (essential)
//    slices hold at most 32 postings is evaluated document-at-a-time, non-essential values read from the
//    terms' value rows (hot or lookup).  The pass kernel visits only the units left in the mask.
//  * Selection (fused_select_kernel): every candidate is evaluated EXACTLY -- per field bm25s's fp32 sum
//    in query order, tf, the fp64 posterior; then the conjunction in signal order, the same code path
//    (fuse_step) as the dense single-query kernels -- and the best k by (fused, doc id) are kept.  The k-th
//    fused value is mapped back to a threshold for the traversal by inverting the final sigmoid with a
//    step-down that keeps ties of the rounded fp64 sigmoid on the safe side.
//  * Thresholds are seeded from a strided sample of the query's rarest posting list, then tightened after
//    each block group; candidate-row overflows are repaired by enqueued re-runs (no host round trip).
//    Queries the machinery cannot decide (fewer than k matching documents, > 32 terms in a field,
//    rows that keep overflowing, a threshold a skipped unit could still beat) take the dense
//    guaranteed path: bb25_fuse_*_signal over all documents + bb25_topk_f64.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "bb25_internal.cuh"
#include "bb25_device.cuh"

namespace bb25 {

constexpr int kMaxFields = 4;
constexpr int kFusedMaxK = 1024;
constexpr int kFusedRowCap = 28672;  // candidate row per query, consumed in chunks of (sort slots - k)
constexpr int kFusedSortCap = 8192;
constexpr int kSeedMax = 512;
#ifndef BB25_PRED_MAX
#define BB25_PRED_MAX 96
#endif
constexpr int kPredMaxPostings = BB25_PRED_MAX;  // frequent-term restriction only for units with at most this many S postings
constexpr int FQC = 8;     // queries per warp work item
constexpr int FWARPS = 8;  // warps per CTA
constexpr double kNeg = 23.02585092984;  // -logit(1e-10), probability.py:20,44-48
constexpr float kNegF = 23.0258522f;     // fp32, rounded up

struct FField {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    BlockTable tab;
    const float *dense_vals;
    const float *lookup_vals;  // value rows of the mid-frequency terms (bb25_index::lookup_vals) or NULL
    int n_hot;                 // row slots below this are rows of dense_vals, the others rows of lookup_vals
    int64_t dense_stride;
    const int32_t *doc_len;
    double avgdl;
    bb25_params params;
    double cw;  // conjunction weight of the signal (w_i; 1 for the unweighted mean)
    float ka;   // fp32, rounded up: cw * alpha
    // the batch's queries for this field (sanitised workspace copies)
    const int32_t *q_terms;
    const uint8_t *q_nocount;
    const int64_t *q_off;
    const longlong2 *qt_info;
    float *q_o;    // [n_q] per-query constant of the bound (fp32, rounded up)
    float *q_ptf;  // [n_q] tf prior at tf = #distinct query terms (probability.py:110-115), rounded up
    // per-document refinement of the bound on the emit path: o_doc = cwf * max(ef, basef + logit(prior(doc)))
    float cwf, basef, ef, inv_avgdl;
    float ucap;  // fp32, rounded up: the field's largest shifted contribution, cw * (2*23.03 + min(0, logit(base_rate)))
};

struct FusedCommon {
    int n_fields, has_cos, unweighted, n_signals;
    double scale, cos_w, c_total;
    const float *cosine;
    int64_t cos_stride;
    float kc;        // fp32, rounded up: cos_w
    float cos_ub;    // fp32, rounded up: cos_w * 2 * 23.03 (a unit's dense-signal bound)
    int64_t n_docs, doc_id_offset;
    int k, kpad, cap;
    float *thr;  // [n_q] shifted threshold with the safety margins applied; 0 = none yet
    unsigned int *cand_cnt;
    unsigned long long *cand_key;  // [n_q][cap]: (fp32 bound bits << 32) | local doc id
    unsigned long long *best_f;    // [n_q][kpad] fused fp64 bits, sorted
    uint32_t *best_id;             // [n_q][kpad]
    unsigned int *best_n;
    uint8_t *flags;  // 1 = the dense guaranteed path decides this query
};

struct FusedBlockArgs {
    FusedCommon c;
    FField f[kMaxFields];
    const int32_t *q_list;
    int n_q;
    const unsigned int *n_q_ptr;
    int blk_begin, blk_end;
    int prune;
    int pair_mode;  // two fields: both in one pass (A/B accumulators) instead of one fold pass per field
    int sparse_mode;  // pair mode: essential-posting evaluation of units (pruning level >= 2)
    uint8_t *unit_mask;  // sparse mode: per work item (block, 8-query chunk) the query slots that need the pass, written by
                         // fused_group_kernel and read by fused_block_kernel
    unsigned long long *work_counter;
    unsigned long long *stats;  // [0] units handed out, [1] skipped by the block bound, [2] abandoned between fields,
                                // [3] evaluated under the frequent-term restriction, [4] skipped: no essential posting,
                                // [5] evaluated through their essential postings, [6] documents evaluated that way
};

struct FTermEnt {
    long long start;
    int len;
    float bmax;
    int dslot;    // hot (dense) row: the passes add it in registers
    int rslot;    // any value row (hot or lookup): document-at-a-time evaluation reads it
    uint32_t pk;  // the table entry's packed (block maximum | length) word
};

__device__ __forceinline__ FTermEnt fused_load_entry(const FField &ff, int blk, long long pos, bool active) {
    FTermEnt e;
    e.start = 0;
    e.len = 0;
    e.bmax = 0.f;
    e.dslot = -1;
    e.rslot = -1;
    e.pk = 0u;
    if (active) {
        const longlong2 info = ff.qt_info[2 * pos];      // (indptr[t], hot slot | any-row slot << 32)
        const longlong2 trow = ff.qt_info[2 * pos + 1];  // the term's block-table row
        const uint2 ent = tab_lookup(ff.tab, trow, blk);
        e.pk = ent.y;
        e.len = (int)(ent.y & kBlkLenMask);
        e.bmax = __uint_as_float(ent.y & ~kBlkLenMask);
        e.start = info.x + (long long)ent.x;
        e.dslot = e.len > 0 ? (int)info.y : -1;
        e.rslot = e.len > 0 ? (int)(info.y >> 32) : -1;
    }
    return e;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}

// fp32 upper bound of cos_w * (logit(cosine_to_probability(c)) + 23.03), fusion.py:43-45 + probability.py:44-48
__device__ __forceinline__ float cos_bound(float c, float kc) {
    c = fminf(fmaxf(c, -1.f), 1.f);
    float t = __logf(__fdividef(1.f + c, 1.f - c));  // c = 1 -> +inf, c = -1 -> -inf
    t = fminf(fmaxf(t, -kNegF), kNegF);
    return __fmul_rn(kc, __fadd_rn(t, kNegF + 2e-3f));  // 2e-3: the fast log / divide
}

// Rare path of the traversal: a quad whose bound reaches the threshold.  Before a document becomes a
// candidate its bound is tightened with what only this document knows: the length prior of every field
// it matched (probability.py:117-129; the traversal bound assumes the best case P_norm = 0.9).  `mask4`
// holds, per document, the fields (traversal order) that matched.  The refined value is still an upper
// bound (tf stays at the query's number of distinct terms).
__device__ __forceinline__ void emit_one_fused(const FusedBlockArgs &a, float v, uint32_t m, uint32_t doc, float thr, int q,
                                               const uint4 *sfd_q) {
    if (!(v > 0.f && v >= thr) || (int64_t)doc >= a.c.n_docs) return;
    for (int i = 0; i < a.c.n_fields; i++) {
        if (!((m >> i) & 1u)) continue;
        const FField &ff = a.f[i];
        if (ff.params.prior_mode != 0) continue;  // prior_free: the bound already uses the exact prior
        const float o_max = __uint_as_float(sfd_q[i].z);
        const float ptf = __uint_as_float(sfd_q[i].w);
        const float r = __fmul_rn((float)ff.doc_len[doc], ff.inv_avgdl);
        const float pn = 0.3f + 0.6f * (1.f - fminf(1.f, fabsf(r - 0.5f) * 2.f)) + 2e-6f;
        const float pr = fminf(0.9f, fmaxf(0.1f, 0.7f * ptf + 0.3f * pn));
        const float lp = __logf(__fdividef(pr, 1.f - pr)) + 2e-5f;
        const float o_doc = __fmul_rn(ff.cwf, fmaxf(ff.ef, ff.basef + lp)) * 1.000002f + 1e-6f;
        if (o_doc < o_max) v -= (o_max - o_doc);
    }
    if (v >= thr) {
        const unsigned int pos = atomicAdd(a.c.cand_cnt + q, 1u);
        if (pos < (unsigned)a.c.cap)
            a.c.cand_key[(size_t)q * (size_t)a.c.cap + pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)doc;
    }
}
__device__ __noinline__ void emit_quad_fused(const FusedBlockArgs &a, float4 u, uint32_t mask4, uint32_t first_id, float thr,
                                             int q, const uint4 *sfd_q) {
    const float av[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int c = 0; c < 4; c++) emit_one_fused(a, av[c], (mask4 >> (8 * c)) & 15u, first_id + c, thr, q, sfd_q);
}

// One field's fold pass over the warp's block: s = S-term sums (accumulators) + D-term rows (registers);
// u = s > 0 ? ka*s + o : 0 is added to the running bound, which lives in shared memory between fields.  The
// last active field adds the dense signal's term, tests against the threshold and emits.
// The emit path wants to know WHICH fields a document matched (per-document prior refinement).  With two
// fields that is free: the stored bound is > 0 exactly when the first field matched (o > 0).  With more
// (MASKED) the set rides in the low 4 mantissa bits of the stored bound, which is rounded UP to a multiple
// of 16 ulps first.  Returns the lane's maximum of the running bound (used to abandon a unit between fields).
template <bool HAS_COS, bool MASKED>
__device__ __forceinline__ float fold_pass(const FusedBlockArgs &a, const FField &ff, int fbit, float4 *B4, uint4 *U4,
                                           bool has_s, unsigned dmask, int dslot, float ka, float o, bool is_first,
                                           bool is_last, int doc_base, int lane, float thr, int q, const uint4 *sfd_q) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *dbase = ff.dense_vals + doc_base + lane * 4;
    float lane_max = 0.f;
#pragma unroll 1
    for (int h = 0; h < kBlockDocs / 256; h++) {
        float4 v[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = h * 64 + j * 32 + lane;
            v[j] = has_s ? B4[w] : zero4;
            if (has_s) B4[w] = zero4;
        }
        for (unsigned mm = dmask; mm; mm &= mm - 1) {
            const int slot = __shfl_sync(0xFFFFFFFFu, dslot, __ffs(mm) - 1);
            const float4 *rp = reinterpret_cast<const float4 *>(dbase + (size_t)slot * (size_t)ff.dense_stride);
            float4 r[2];
#pragma unroll
            for (int j = 0; j < 2; j++) r[j] = ld_row_f4(rp + h * 64 + j * 32);
#pragma unroll
            for (int j = 0; j < 2; j++) add_f4(v[j], r[j]);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = h * 64 + j * 32 + lane;
            const float sv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint4 pu = make_uint4(0u, 0u, 0u, 0u);
            if (!is_first) pu = U4[w];
            const uint32_t pb[4] = {pu.x, pu.y, pu.z, pu.w};
            float un[4];
            uint32_t mk[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const bool hit = sv[c] > 0.f;
                const float uc = hit ? __fmaf_rn(ka, sv[c], o) : 0.f;
                if (MASKED) {
                    un[c] = __fadd_rn(__uint_as_float(pb[c] & ~15u), uc);
                    mk[c] = (pb[c] & 15u) | ((hit ? 1u : 0u) << fbit);
                } else {
                    un[c] = __fadd_rn(__uint_as_float(pb[c]), uc);
                    mk[c] = 0u;
                }
            }
            if (!is_last) {
                uint4 st;
                if (MASKED) {
                    st.x = ((__float_as_uint(un[0]) + 15u) & ~15u) | mk[0];
                    st.y = ((__float_as_uint(un[1]) + 15u) & ~15u) | mk[1];
                    st.z = ((__float_as_uint(un[2]) + 15u) & ~15u) | mk[2];
                    st.w = ((__float_as_uint(un[3]) + 15u) & ~15u) | mk[3];
                } else {
                    st = make_uint4(__float_as_uint(un[0]), __float_as_uint(un[1]), __float_as_uint(un[2]), __float_as_uint(un[3]));
                }
                U4[w] = st;
                lane_max = fmaxf(lane_max, fmaxf(fmaxf(un[0], un[1]), fmaxf(un[2], un[3])));
                continue;
            }
            const uint32_t first_id = (uint32_t)(doc_base + (h * 64 + j * 32 + lane) * 4);
            float4 u = make_float4(un[0], un[1], un[2], un[3]);
            if (HAS_COS) {
                // rows are padded to a multiple of 4 floats; documents beyond n_docs are never emitted
                if ((int64_t)first_id < a.c.n_docs) {
                    const float4 c4 = ld_nc_f4(reinterpret_cast<const float4 *>(a.c.cosine + (size_t)q * (size_t)a.c.cos_stride + first_id));
                    u.x = __fadd_rn(u.x, cos_bound(c4.x, a.c.kc));
                    u.y = __fadd_rn(u.y, cos_bound(c4.y, a.c.kc));
                    u.z = __fadd_rn(u.z, cos_bound(c4.z, a.c.kc));
                    u.w = __fadd_rn(u.w, cos_bound(c4.w, a.c.kc));
                }
            }
            const float mx = fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w));
            if (mx > 0.f && mx >= thr) {
                uint32_t m4;
                if (MASKED) {
                    m4 = mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
                } else {
                    // at most two fields: an earlier field matched iff the stored bound is positive
                    m4 = 0u;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t prev_hit = (!is_first && __uint_as_float(pb[c]) > 0.f) ? ((1u << fbit) - 1u) : 0u;
                        m4 |= (prev_hit | ((sv[c] > 0.f ? 1u : 0u) << fbit)) << (8 * c);
                    }
                }
                emit_quad_fused(a, u, m4, first_id, thr, q, sfd_q);
            }
        }
    }
    return lane_max;
}

// Two fields in ONE pass (the MultiFieldScorer title + body shape): field 0's S-term sums sit in accumulator A,
// field 1's in B; per quad both are read, the D rows of both fields are added in registers, the two bounds
// are formed and summed, tested, emitted.  Half the shared-memory traffic of two fold passes, no running
// bound to store -- and the frequent-term restriction of block_kernel's level 2 carries over: when the D terms
// alone (summed block maxima, both fields) cannot reach the threshold, a quad without any S contribution
// cannot qualify and is skipped before its rows are loaded (PRED).
template <bool HAS_COS, bool PRED>
__device__ __forceinline__ void pair_pass(const FusedBlockArgs &a, float4 *A4, float4 *B4, bool has_s0, bool has_s1,
                                          unsigned dmask0, unsigned dmask1, int dslot0, int dslot1, float o0, float o1,
                                          int doc_base, int lane, float thr, int q, const uint4 *sfd_q) {
    const FField &f0 = a.f[0];
    const FField &f1 = a.f[1];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *dbase0 = f0.dense_vals + doc_base + lane * 4;
    const float *dbase1 = f1.dense_vals + doc_base + lane * 4;
    const float ka0 = f0.ka, ka1 = f1.ka;
#pragma unroll 1
    for (int h = 0; h < kBlockDocs / 256; h++) {
        float4 va[2], vb[2];
        bool live[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = h * 64 + j * 32 + lane;
            va[j] = has_s0 ? A4[w] : zero4;
            vb[j] = has_s1 ? B4[w] : zero4;
            live[j] = true;
            if (PRED) {
                const float ma = fmaxf(fmaxf(va[j].x, va[j].y), fmaxf(va[j].z, va[j].w));
                const float mb = fmaxf(fmaxf(vb[j].x, vb[j].y), fmaxf(vb[j].z, vb[j].w));
                live[j] = fmaxf(ma, mb) > 0.f;
            }
            if (live[j]) {  // an untouched quad holds zeros already
                if (has_s0) A4[w] = zero4;
                if (has_s1) B4[w] = zero4;
            }
        }
        for (unsigned mm = dmask0; mm; mm &= mm - 1) {
            const int slot = __shfl_sync(0xFFFFFFFFu, dslot0, __ffs(mm) - 1);
            const float4 *rp = reinterpret_cast<const float4 *>(dbase0 + (size_t)slot * (size_t)f0.dense_stride);
            float4 r[2];
#pragma unroll
            for (int j = 0; j < 2; j++) r[j] = live[j] ? ld_row_f4(rp + h * 64 + j * 32) : zero4;
#pragma unroll
            for (int j = 0; j < 2; j++) add_f4(va[j], r[j]);
        }
        for (unsigned mm = dmask1; mm; mm &= mm - 1) {
            const int slot = __shfl_sync(0xFFFFFFFFu, dslot1, __ffs(mm) - 1);
            const float4 *rp = reinterpret_cast<const float4 *>(dbase1 + (size_t)slot * (size_t)f1.dense_stride);
            float4 r[2];
#pragma unroll
            for (int j = 0; j < 2; j++) r[j] = live[j] ? ld_row_f4(rp + h * 64 + j * 32) : zero4;
#pragma unroll
            for (int j = 0; j < 2; j++) add_f4(vb[j], r[j]);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (PRED && !live[j]) continue;
            const float sa[4] = {va[j].x, va[j].y, va[j].z, va[j].w};
            const float sb[4] = {vb[j].x, vb[j].y, vb[j].z, vb[j].w};
            float un[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float ua = sa[c] > 0.f ? __fmaf_rn(ka0, sa[c], o0) : 0.f;
                const float ub = sb[c] > 0.f ? __fmaf_rn(ka1, sb[c], o1) : 0.f;
                un[c] = __fadd_rn(ua, ub);
            }
            const uint32_t first_id = (uint32_t)(doc_base + (h * 64 + j * 32 + lane) * 4);
            float4 u = make_float4(un[0], un[1], un[2], un[3]);
            if (HAS_COS) {
                if ((int64_t)first_id < a.c.n_docs) {
                    const float4 c4 = ld_nc_f4(reinterpret_cast<const float4 *>(a.c.cosine + (size_t)q * (size_t)a.c.cos_stride + first_id));
                    u.x = __fadd_rn(u.x, cos_bound(c4.x, a.c.kc));
                    u.y = __fadd_rn(u.y, cos_bound(c4.y, a.c.kc));
                    u.z = __fadd_rn(u.z, cos_bound(c4.z, a.c.kc));
                    u.w = __fadd_rn(u.w, cos_bound(c4.w, a.c.kc));
                }
            }
            const float mx = fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w));
            if (mx > 0.f && mx >= thr) {
                uint32_t m4 = 0u;
#pragma unroll
                for (int c = 0; c < 4; c++) m4 |= ((sa[c] > 0.f ? 1u : 0u) | (sb[c] > 0.f ? 2u : 0u)) << (8 * c);
                emit_quad_fused(a, u, m4, first_id, thr, q, sfd_q);
            }
        }
    }
}

// A unit in which no field has a posting still holds 1024 documents with a dense-signal value each.
template <bool HAS_COS>
__device__ __forceinline__ void cos_only_pass(const FusedBlockArgs &a, int doc_base, int lane, float thr, int q,
                                              const uint4 *sfd_q) {
    if (!HAS_COS) return;
#pragma unroll 2
    for (int h = 0; h < kBlockDocs / 128; h++) {
        const uint32_t first_id = (uint32_t)(doc_base + (h * 32 + lane) * 4);
        if ((int64_t)first_id >= a.c.n_docs) continue;
        const float4 c4 = ld_nc_f4(reinterpret_cast<const float4 *>(a.c.cosine + (size_t)q * (size_t)a.c.cos_stride + first_id));
        float4 u;
        u.x = cos_bound(c4.x, a.c.kc);
        u.y = cos_bound(c4.y, a.c.kc);
        u.z = cos_bound(c4.z, a.c.kc);
        u.w = cos_bound(c4.w, a.c.kc);
        const float mx = fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w));
        if (mx > 0.f && mx >= thr) emit_quad_fused(a, u, 0u, first_id, thr, q, sfd_q);
    }
}

// ---------------------------------------------------------------------------------
// Essential-posting evaluation of a unit (two fields, no dense signal, pruning level >= 2): MaxScore
// (the partition behind wand_upper_bound, probability.py:205-236) at the granularity of one 1024-document
// block, on the fused key.
//
// The unit's (field, term) entries are split: NON-ESSENTIAL = the entries that have a value row (hot or
// lookup) and at least L postings in this block, for the smallest L of a fixed ladder whose bound -- the
// unit bound of the block-max test restricted to that subset: per field ka * (sum of block maxima) + o,
// capped -- stays below the threshold.  A document of the block that matches non-essential entries only
// cannot qualify, so every qualifying document has a posting in one of the ESSENTIAL slices (row-less
// terms, and row terms with fewer than L postings here).  When those slices hold at most 32 postings the
// unit is evaluated document-at-a-time in ONE round: lane = essential posting (doc id and value straight
// from the slice), postings of the same document are combined with a warp match, each non-essential
// term's value is one 4-byte load from its row -- no accumulators, no scatter, no pass over 1024
// documents.  Zipf queries carry a rare term more often than not, and a rare term has a handful of
// postings per block: this prunes where the block-max test alone cannot (nearly every block holds every
// frequent term near its maximum).  The ladder's last step is the full entry set: the block-max skip test
// of the traversal, folded in.
// ---------------------------------------------------------------------------------
constexpr int kSparseMax = 32;  // essential postings evaluated in one round; longer lists take the pass

__device__ __forceinline__ float row_value(const FField &ff, int rslot, uint32_t doc) {
    return rslot < ff.n_hot ? ff.dense_vals[(size_t)rslot * (size_t)ff.dense_stride + doc]
                            : ff.lookup_vals[(size_t)(rslot - ff.n_hot) * (size_t)ff.dense_stride + doc];  // absent: -0.0f
}

// One round of document-at-a-time evaluation for the queries of `batch` (a mask of 8-lane groups whose
// essential postings add up to at most 32): lane = essential posting.
__device__ __forceinline__ void sparse_round(const FusedBlockArgs &a, const FTermEnt &e0, const FTermEnt &e1, bool sparse_q, int L,
                                             float thr_l, float o0, float o1, unsigned batch, int mmax, int s0, const uint4 *sfd,
                                             const uint2 *sq, int doc_base, int lane) {
    int o = -1, fld = 0, qid = 0, fill = 0;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const FTermEnt &e = i == 0 ? e0 : e1;
        const bool ess = sparse_q && e.len > 0 && (e.rslot < 0 || e.len < L);
        for (unsigned mm = __ballot_sync(0xFFFFFFFFu, ess) & batch; mm; mm &= mm - 1) {
            const int src = __ffs(mm) - 1;
            const int len = __shfl_sync(0xFFFFFFFFu, e.len, src);
            const long long s = shfl_ll(e.start, src);
            const int r = lane - fill;
            if (r >= 0 && r < len) {
                o = ld_nc_s32(a.f[i].indices + s + r) - doc_base;
                v = ld_nc_f32(a.f[i].data + s + r);
                fld = i;
                qid = src >> 3;
            }
            fill += len;
        }
    }
    const bool valid = o >= 0;
    // postings of one (query, document) -- several essential slices may hold it -- are summed per field by the first lane
    const unsigned g = __match_any_sync(0xFFFFFFFFu, valid ? ((qid << 10) | o) : (0x10000 | lane));
    const int rounds = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)__popc(g));
    float sa = 0.f, sb = 0.f;
    unsigned rem = g;
    for (int r = 0; r < rounds; r++) {
        const int src = rem ? __ffs(rem) - 1 : lane;
        const float vv = __shfl_sync(0xFFFFFFFFu, v, src);
        const int ff = __shfl_sync(0xFFFFFFFFu, fld, src);
        if (rem) {
            if (ff) sb = __fadd_rn(sb, vv);
            else sa = __fadd_rn(sa, vv);
            rem &= rem - 1;
        }
    }
    const bool leader = valid && (__ffs(g) - 1 == lane);
    const uint32_t doc = (uint32_t)(doc_base + (valid ? o : 0));
    const int ql = qid << 3;  // first lane of the posting's query
    const int Lq = __shfl_sync(0xFFFFFFFFu, L, ql);
    const float thrq = __shfl_sync(0xFFFFFFFFu, thr_l, ql);
    const float o0q = __shfl_sync(0xFFFFFFFFu, o0, ql), o1q = __shfl_sync(0xFFFFFFFFu, o1, ql);
    // the non-essential terms of the posting's query: one load from the term's value row each
    for (int t = 0; t < mmax; t++) {
        const int src = ql | t;
        const int rs0 = __shfl_sync(0xFFFFFFFFu, e0.rslot, src), l0 = __shfl_sync(0xFFFFFFFFu, e0.len, src);
        const int rs1 = __shfl_sync(0xFFFFFFFFu, e1.rslot, src), l1 = __shfl_sync(0xFFFFFFFFu, e1.len, src);
        if (leader && rs0 >= 0 && l0 >= Lq) sa = __fadd_rn(sa, row_value(a.f[0], rs0, doc));
        if (leader && rs1 >= 0 && l1 >= Lq) sb = __fadd_rn(sb, row_value(a.f[1], rs1, doc));
    }
    if (leader) {
        const float ua = sa > 0.f ? __fmaf_rn(a.f[0].ka, sa, o0q) : 0.f;
        const float ub = sb > 0.f ? __fmaf_rn(a.f[1].ka, sb, o1q) : 0.f;
        emit_one_fused(a, __fadd_rn(ua, ub), (sa > 0.f ? 1u : 0u) | (sb > 0.f ? 2u : 0u), doc, thrq, (int)sq[s0 + qid].x,
                       sfd + (s0 + qid) * 2);
    }
}

// The chunk's queries, FOUR AT A TIME (8 lanes per query: lane = term position of both fields for the table
// entries, = ladder step for the split): block-max skip, essential split, document-at-a-time evaluation.
// Returns the mask of query slots that need the pass (or cannot be handled here: more than 8 terms in a
// field, no threshold yet); the caller runs those one by one.
__device__ __forceinline__ unsigned group_units(const FusedBlockArgs &a, const uint4 *sfd, const uint2 *sq, int nslots, int blk,
                                                int doc_base, int lane, unsigned &skipped, unsigned &sparse_units,
                                                unsigned &sparse_docs, unsigned &hist_a, unsigned &hist_b) {
    const int qs = lane >> 3, tl = lane & 7, sh = qs * 8;
    // ladder: row entries with >= 1, 2, 3, 5, 9, 17, 33 postings; step 7: every entry (the block-max test)
    const int mycut = (tl == 0 || tl == 7) ? 1 : 1 + (1 << (tl - 1));
    unsigned serial = 0u;
    for (int s0 = 0; s0 < nslots; s0 += 4) {
        const int sl = s0 + qs;
        const bool qact = sl < nslots;
        uint4 d0 = make_uint4(0u, 0u, 0u, 0u), d1 = d0;
        float thr_l = 0.f;
        if (qact) {
            d0 = sfd[sl * 2];
            d1 = sfd[sl * 2 + 1];
            thr_l = __uint_as_float(sq[sl].y);
        }
        const int m0 = (int)d0.y, m1 = (int)d1.y;
        const bool ser = qact && (m0 > 8 || m1 > 8 || !(thr_l > 0.f));
        const bool grp = qact && !ser;
        const FTermEnt e0 = fused_load_entry(a.f[0], blk, (long long)d0.x + tl, grp && tl < m0);
        const FTermEnt e1 = fused_load_entry(a.f[1], blk, (long long)d1.x + tl, grp && tl < m1);
        const unsigned pb = __ballot_sync(0xFFFFFFFFu, e0.len > 0 || e1.len > 0);
        const unsigned rb0 = __ballot_sync(0xFFFFFFFFu, e0.rslot >= 0), rb1 = __ballot_sync(0xFFFFFFFFu, e1.rslot >= 0);
        const unsigned rm0 = tl == 7 ? 0xFFu : (rb0 >> sh) & 0xFFu;
        const unsigned rm1 = tl == 7 ? 0xFFu : (rb1 >> sh) & 0xFFu;
        const int mmax = min(8, (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(grp ? max(m0, m1) : 0)));
        float b0 = 0.f, b1 = 0.f;
        bool any0 = false, any1 = false;
        for (int t = 0; t < mmax; t++) {
            const int src = (lane & 24) | t;
            const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, e0.pk, src), w1 = __shfl_sync(0xFFFFFFFFu, e1.pk, src);
            if ((int)(w0 & kBlkLenMask) >= mycut && ((rm0 >> t) & 1u)) {
                b0 = __fadd_rn(b0, __uint_as_float(w0 & ~kBlkLenMask));
                any0 = true;
            }
            if ((int)(w1 & kBlkLenMask) >= mycut && ((rm1 >> t) & 1u)) {
                b1 = __fadd_rn(b1, __uint_as_float(w1 & ~kBlkLenMask));
                any1 = true;
            }
        }
        const float o0 = __uint_as_float(d0.z), o1 = __uint_as_float(d1.z);
        // the margins of the block-max test
        const float f0 = any0 ? fminf(__fmul_rn(__fmaf_rn(a.f[0].ka, __fmul_rn(b0, 1.000004f), o0), 1.000002f), a.f[0].ucap) : 0.f;
        const float f1 = any1 ? fminf(__fmul_rn(__fmaf_rn(a.f[1].ka, __fmul_rn(b1, 1.000004f), o1), 1.000002f), a.f[1].ucap) : 0.f;
        const bool below = grp && __fmul_rn(__fadd_rn(f0, f1), 1.000002f) < thr_l;
        const unsigned bm = (__ballot_sync(0xFFFFFFFFu, below) >> sh) & 0xFFu;
        int L = 1;
        if (bm & 0x7Fu) {
            const int ci = __ffs(bm & 0x7Fu) - 1;
            L = ci == 0 ? 1 : 1 + (1 << (ci - 1));
        }
        int n_e = ((e0.len > 0 && (e0.rslot < 0 || e0.len < L)) ? e0.len : 0) + ((e1.len > 0 && (e1.rslot < 0 || e1.len < L)) ? e1.len : 0);
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 1);
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 2);
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 4);
        // 0 nothing to do, 1 skipped by the bound, 2 essential postings, 3 the pass
        int cls;
        if (!qact || (grp && !((pb >> sh) & 0xFFu))) cls = 0;
        else if (ser) cls = 3;
        else if (bm & 0x80u) cls = 1;
        else if (!(bm & 0x7Fu) || n_e > kSparseMax) cls = 3;
        else cls = n_e > 0 ? 2 : 1;
        skipped += (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, cls == 1 && tl == 0));
#ifdef BB25_NE_HIST  // tuning builds: how long are the essential lists of the units that take the pass?
        hist_a += (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, cls == 3 && (bm & 0x7Fu) && n_e > 32 && n_e <= 64 && tl == 0));
        hist_b += (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, cls == 3 && (bm & 0x7Fu) && n_e > 64 && n_e <= 128 && tl == 0));
#endif
        const unsigned sb3 = __ballot_sync(0xFFFFFFFFu, cls == 3 && tl == 0);
#pragma unroll
        for (int gq = 0; gq < 4; gq++)
            if ((sb3 >> (gq * 8)) & 1u) serial |= 1u << (s0 + gq);
        unsigned pend = __ballot_sync(0xFFFFFFFFu, cls == 2 && tl == 0);
        while (pend) {
            unsigned batch = 0u;
            int tot = 0;
            for (unsigned pm = pend; pm; pm &= pm - 1) {
                const int gl = __ffs(pm) - 1;
                const int ne = __shfl_sync(0xFFFFFFFFu, n_e, gl);
                if (tot + ne > kSparseMax) break;
                tot += ne;
                batch |= 0xFFu << gl;
            }
            pend &= ~batch;
            sparse_round(a, e0, e1, cls == 2, L, thr_l, o0, o1, batch, mmax, s0, sfd, sq, doc_base, lane);
            sparse_units += (unsigned)__popc(batch) >> 3;
            sparse_docs += (unsigned)tot;
        }
    }
    return serial;
}

// The light half of the two-field traversal at pruning level >= 2: no accumulators, few registers, twice the resident
// warps of fused_block_kernel -- which matters because a unit here is a short chain of dependent loads (query
// descriptors -> term records -> table entries -> essential postings -> value rows).  Per work item (block, 8-query
// chunk) it runs group_units and leaves the query slots that need the pass in the unit mask.
constexpr int GWARPS = 8;
#ifndef BB25_GROUP_CTAS
#define BB25_GROUP_CTAS 6
#endif
__global__ void __launch_bounds__(GWARPS * 32, BB25_GROUP_CTAS) fused_group_kernel(const __grid_constant__ FusedBlockArgs a) {
    __shared__ __align__(16) unsigned char gsm[GWARPS * (FQC * 2 * 16 + FQC * 8)];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint4 *sfd = reinterpret_cast<uint4 *>(gsm + (size_t)warp * (FQC * 2 * 16 + FQC * 8));  // [FQC][2]: t0, m, o bits, P_tf bits
    uint2 *sq = reinterpret_cast<uint2 *>(sfd + FQC * 2);                                   // [FQC]: q, threshold bits
    const int n_q = a.n_q_ptr ? (int)*a.n_q_ptr : a.n_q;
    const int n_chunks = (n_q + FQC - 1) / FQC;
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&a.stats[0], (unsigned long long)(a.blk_end - a.blk_begin) * (unsigned long long)n_q);
    unsigned int skipped = 0u, sparse_units = 0u, sparse_docs = 0u, hist_a = 0u, hist_b = 0u;
    for (;;) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(a.work_counter, 1ull);
        item = shfl_ll(item, 0);
        if (item >= n_items) break;
        const int blk = a.blk_begin + (int)(item / n_chunks);
        const int slot0 = (int)(item % n_chunks) * FQC;
        const int nslots = min(FQC, n_q - slot0);
        __syncwarp();
        if (lane < nslots * 2) {
            const int slot = lane >> 1, i = lane & 1;
            const int my_q = a.q_list ? a.q_list[slot0 + slot] : slot0 + slot;
            const FField &ff = a.f[i];
            const long long t0 = ff.q_off[my_q];
            long long m = (long long)ff.q_off[my_q + 1] - t0;
            m = m < 0 ? 0 : (m > 32 ? 32 : m);
            sfd[slot * 2 + i] = make_uint4((unsigned)t0, (unsigned)m, __float_as_uint(ff.q_o[my_q]), __float_as_uint(ff.q_ptf[my_q]));
            if (i == 0) sq[slot] = make_uint2((unsigned)my_q, __float_as_uint(a.c.thr[my_q]));
        }
        __syncwarp();
        const unsigned serial = group_units(a, sfd, sq, nslots, blk, blk * kBlockDocs, lane, skipped, sparse_units, sparse_docs, hist_a, hist_b);
        if (lane == 0) a.unit_mask[item] = (uint8_t)serial;
    }
    if (lane == 0) {
        if (skipped) atomicAdd(&a.stats[1], (unsigned long long)skipped);
        if (hist_a) atomicAdd(&a.stats[4], (unsigned long long)hist_a);
        if (hist_b) atomicAdd(&a.stats[2], (unsigned long long)hist_b);
        if (sparse_units) atomicAdd(&a.stats[5], (unsigned long long)sparse_units);
        if (sparse_docs) atomicAdd(&a.stats[6], (unsigned long long)sparse_docs);
    }
}

// In FusedBlockArgs the fields are in TRAVERSAL order (cheapest index first): after each field but the last
// the unit is abandoned when no document's running bound plus the remaining fields' block bounds can reach
// the threshold -- the per-document version of the block-max test, far tighter because it uses the scores
// the documents of the block actually have in the fields seen so far instead of the sum of per-term maxima.
template <int F, bool HAS_COS, bool PAIR>
__global__ void __launch_bounds__(FWARPS * 32, F >= 2 ? 3 : 5) fused_block_kernel(const __grid_constant__ FusedBlockArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int kAccBytes = (F >= 2 ? 2 : 1) * kBlockDocs * 4;
    constexpr int kWarpBytes = kAccBytes + FQC * 8 + FQC * F * 16;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem + (size_t)warp * kWarpBytes;
    float *B = reinterpret_cast<float *>(wbase);
    float4 *B4 = reinterpret_cast<float4 *>(B);
    uint4 *U4 = reinterpret_cast<uint4 *>(wbase + (F >= 2 ? kBlockDocs * 4 : 0));  // unused when F == 1
    uint4 *sfd = reinterpret_cast<uint4 *>(wbase + kAccBytes);                    // [FQC][F]: t0, m, o bits, P_tf bits
    uint2 *sq = reinterpret_cast<uint2 *>(wbase + kAccBytes + FQC * F * 16);      // [FQC]: q, threshold bits
    for (int i = lane; i < kBlockDocs / 4; i += 32) B4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (PAIR)
        for (int i = lane; i < kBlockDocs / 4; i += 32) U4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    const int n_q = a.n_q_ptr ? (int)*a.n_q_ptr : a.n_q;
    const int n_chunks = (n_q + FQC - 1) / FQC;
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    // two fields at pruning level >= 2: fused_group_kernel has already skipped / evaluated most units; what is left
    // for the pass is in the unit mask
    const bool masked = PAIR && F == 2 && !HAS_COS && a.unit_mask != nullptr;
    if (blockIdx.x == 0 && threadIdx.x == 0 && !masked)
        atomicAdd(&a.stats[0], (unsigned long long)(a.blk_end - a.blk_begin) * (unsigned long long)n_q);
    unsigned int skipped = 0u, abandoned = 0u, restricted = 0u;

    for (;;) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(a.work_counter, 1ull);
        item = shfl_ll(item, 0);
        if (item >= n_items) break;
        unsigned serial = 0xFFFFFFFFu;
        if (masked) {
            serial = (unsigned)a.unit_mask[item];
            if (!serial) continue;
        }
        const int blk = a.blk_begin + (int)(item / n_chunks);
        const int slot0 = (int)(item % n_chunks) * FQC;
        const int nslots = min(FQC, n_q - slot0);
        const int doc_base = blk * kBlockDocs;

        __syncwarp();
        if (lane < nslots * F) {
            const int slot = lane / F, i = lane % F;
            const int my_q = a.q_list ? a.q_list[slot0 + slot] : slot0 + slot;
            const FField &ff = a.f[i];
            const long long t0 = ff.q_off[my_q];
            long long m = (long long)ff.q_off[my_q + 1] - t0;
            m = m < 0 ? 0 : (m > 32 ? 32 : m);  // longer queries are flagged for the dense path by fused_prep_kernel
            sfd[slot * F + i] = make_uint4((unsigned)t0, (unsigned)m, __float_as_uint(ff.q_o[my_q]), __float_as_uint(ff.q_ptf[my_q]));
            if (i == 0) sq[slot] = make_uint2((unsigned)my_q, __float_as_uint(a.c.thr[my_q]));
        }
        __syncwarp();

        for (int sidx = 0; sidx < nslots; sidx++) {
            if (!((serial >> sidx) & 1u)) continue;
            const int q = (int)sq[sidx].x;
            const float thr = __uint_as_float(sq[sidx].y);
            const uint4 *sfd_q = sfd + sidx * F;
            FTermEnt e[F];
            unsigned pres[F];
            float ofs[F], fub[F];
            bool any = false;
            int last = -1;
#pragma unroll
            for (int i = 0; i < F; i++) {
                const uint4 d = sfd_q[i];
                e[i] = fused_load_entry(a.f[i], blk, (long long)d.x + lane, lane < (int)d.y);
                ofs[i] = __uint_as_float(d.z);
                fub[i] = 0.f;
                pres[i] = __ballot_sync(0xFFFFFFFFu, e[i].len > 0);
                if (pres[i]) {
                    any = true;
                    last = i;
                    if (a.prune) {
                        // the field's bound at the block maxima: no document of the block can exceed it
                        const float bsum = __fmul_rn(warp_sum_f(e[i].bmax), 1.000004f);
                        fub[i] = fminf(__fmul_rn(__fmaf_rn(a.f[i].ka, bsum, ofs[i]), 1.000002f), a.f[i].ucap);
                    }
                }
            }
            if (!any && !HAS_COS) continue;  // without a dense signal such documents tie at the all-unmatched value
            // rest[i] = what the fields after i (and the dense signal) can still add to any document of the block
            float rest[F];
            {
                float r = HAS_COS ? a.c.cos_ub : 0.f;
#pragma unroll
                for (int i = F - 1; i >= 0; i--) {
                    rest[i] = r;
                    r = __fadd_rn(r, fub[i]);
                }
                if (a.prune && __fmul_rn(r, 1.000002f) < thr) {
                    skipped++;
                    continue;
                }
            }
            if (!any) {
                cos_only_pass<HAS_COS>(a, doc_base, lane, thr, q, sfd_q);
                continue;
            }
            if (PAIR) {
                // ---- two fields, one pass ----
                unsigned dm[2], sm[2];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    dm[i] = __ballot_sync(0xFFFFFFFFu, e[i].dslot >= 0);
                    sm[i] = pres[i] & ~dm[i];
                }
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    float *acc = i == 0 ? reinterpret_cast<float *>(U4) : B;
                    bool fresh = true;
                    for (unsigned mm = sm[i]; mm; mm &= mm - 1) {
                        const int t = __ffs(mm) - 1;
                        const int len = __shfl_sync(0xFFFFFFFFu, e[i].len, t);
                        const long long s = shfl_ll(e[i].start, t);
                        if (fresh) scatter_block<true>(a.f[i].data, a.f[i].indices, s, len, acc, doc_base, lane);
                        else scatter_block<false>(a.f[i].data, a.f[i].indices, s, len, acc, doc_base, lane);
                        fresh = false;
                        __syncwarp();
                    }
                }
                // documents matching frequent (D) terms only: bounded by the D terms' block maxima in both fields
                bool pred = false;
                if (a.prune >= 2 && thr > 0.f && !HAS_COS) {
                    float dub = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const float ds = __fmul_rn(warp_sum_f(e[i].dslot >= 0 ? e[i].bmax : 0.f), 1.000004f);
                        if (dm[i]) dub = __fadd_rn(dub, fminf(__fmul_rn(__fmaf_rn(a.f[i].ka, ds, ofs[i]), 1.000002f), a.f[i].ucap));
                    }
                    pred = __fmul_rn(dub, 1.000002f) < thr;
                    if (pred && !(sm[0] | sm[1])) {
                        restricted++;
                        continue;  // no S posting in the block at all: nothing here can qualify
                    }
                    if (pred) {
                        // worth it only while the S postings are sparse (most quads untouched); measured: with
                        // dense S hits the predicated row loads cost more than they save
                        const int n_s = warp_sum((((sm[0] >> lane) & 1u) ? e[0].len : 0) + (((sm[1] >> lane) & 1u) ? e[1].len : 0));
                        pred = n_s <= kPredMaxPostings;
                    }
                    if (pred) restricted++;
                }
                float4 *A4 = reinterpret_cast<float4 *>(U4);
                if (pred)
                    pair_pass<HAS_COS, true>(a, A4, B4, sm[0] != 0u, sm[1] != 0u, dm[0], dm[1], e[0].dslot, e[1].dslot, ofs[0], ofs[1],
                                             doc_base, lane, thr, q, sfd_q);
                else
                    pair_pass<HAS_COS, false>(a, A4, B4, sm[0] != 0u, sm[1] != 0u, dm[0], dm[1], e[0].dslot, e[1].dslot, ofs[0], ofs[1],
                                              doc_base, lane, thr, q, sfd_q);
                __syncwarp();
                continue;
            }
            bool started = false;
#pragma unroll
            for (int i = 0; i < (PAIR ? 0 : F); i++) {
                if (!pres[i]) continue;
                const FField &ff = a.f[i];
                const unsigned dmask = __ballot_sync(0xFFFFFFFFu, e[i].dslot >= 0);
                const unsigned smask = pres[i] & ~dmask;
                bool fresh = true;
                for (unsigned mm = smask; mm; mm &= mm - 1) {
                    const int t = __ffs(mm) - 1;
                    const int len = __shfl_sync(0xFFFFFFFFu, e[i].len, t);
                    const long long s = shfl_ll(e[i].start, t);
                    if (fresh) scatter_block<true>(ff.data, ff.indices, s, len, B, doc_base, lane);
                    else scatter_block<false>(ff.data, ff.indices, s, len, B, doc_base, lane);
                    fresh = false;
                    __syncwarp();
                }
                float lm = fold_pass<HAS_COS, (F >= 3)>(a, ff, i, B4, U4, smask != 0u, dmask, e[i].dslot, ff.ka, ofs[i], !started,
                                                        i == last, doc_base, lane, thr, q, sfd_q);
                started = true;
                __syncwarp();
                if (F >= 2 && a.prune && i != last) {
                    // no document's running bound plus what the remaining fields (and the dense signal) can add
                    // to ANY document of the block reaches the threshold: the unit is abandoned.  Margins: 16
                    // ulps of rounding per stored field, the summation order of the rest.
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xFFFFFFFFu, lm, d));
                    if (__fmul_rn(__fadd_rn(lm, rest[i]), 1.00002f) + 1e-5f < thr) {
                        abandoned++;
                        break;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        if (skipped) atomicAdd(&a.stats[1], (unsigned long long)skipped);
        if (abandoned) atomicAdd(&a.stats[2], (unsigned long long)abandoned);
        if (restricted) atomicAdd(&a.stats[3], (unsigned long long)restricted);
    }
}

// ---------------------------------------------------------------------------------
// per-query preparation of the fused batch
// ---------------------------------------------------------------------------------
struct FusedPrepArgs {
    FusedCommon c;
    FField f[kMaxFields];
    int64_t n_q;
};

__device__ inline double field_bound_constant(const bb25_params &p, int n_distinct) {
    // upper bound of logit(posterior) - alpha*s, shifted by +23.03 (see the header of this file)
    double lp = 0.0;  // prior_free: prior 0.5 (probability.py:192-193)
    if (p.prior_mode == 0) {
        // composite_prior is monotone in tf and in P_norm <= 0.9 (probability.py:110-140); tf <= #distinct query terms
        const double ptf = 0.2 + 0.7 * fmin(1.0, (double)n_distinct / 10.0);
        double pr = 0.7 * ptf + 0.3 * 0.9;
        pr = pr < 0.1 ? 0.1 : (pr > 0.9 ? 0.9 : pr);
        lp = log(pr / (1.0 - pr));
    }
    const double lbr = p.has_base_rate ? log(p.base_rate / (1.0 - p.base_rate)) : 0.0;
    const double cc = -p.alpha * p.beta + lp + lbr;
    const double e = lbr > 0.0 ? lbr : 0.0;
    const double o = cc + kNeg;
    return o > e ? o : e;
}

__global__ void fused_prep_kernel(const FusedPrepArgs a) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_q) return;
    uint8_t bad = 0;
    for (int i = 0; i < a.c.n_fields; i++) {
        const FField &ff = a.f[i];
        const long long t0 = ff.q_off[q];
        const long long m = (long long)ff.q_off[q + 1] - t0;
        int nd = 0;
        for (long long j = 0; j < m; j++) nd += ff.q_nocount[t0 + j] ? 0 : 1;
        if (m > 32) bad = 1;
        const double o = ff.cw * field_bound_constant(ff.params, nd);
        ff.q_o[q] = __double2float_ru(o * (1.0 + 1e-6) + 1e-6);
        ff.q_ptf[q] = __double2float_ru((0.2 + 0.7 * fmin(1.0, (double)nd / 10.0)) * (1.0 + 1e-7));
    }
    a.c.thr[q] = 0.f;
    a.c.cand_cnt[q] = 0u;
    a.c.best_n[q] = 0u;
    a.c.flags[q] = bad;
}

// threshold seeds: a strided sample of the posting list with the smallest document frequency >= k
// becomes the query's first candidate set (evaluated exactly, used for the threshold only)
__global__ void __launch_bounds__(256) fused_seed_kernel(const FusedPrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= a.n_q) return;
    long long best_df = 0x7FFFFFFFFFFFFFFFll;
    long long best_base = 0;
    int best_f = -1;
    for (int i = 0; i < a.c.n_fields; i++) {
        const FField &ff = a.f[i];
        const long long t0 = ff.q_off[q];
        long long m = (long long)ff.q_off[q + 1] - t0;
        m = m < 0 ? 0 : (m > 32 ? 32 : m);
        if (lane < m) {
            const int t = ff.q_terms[t0 + lane];
            const long long b = ff.indptr[t], df = ff.indptr[t + 1] - b;
            if (df >= a.c.k && df < best_df) {
                best_df = df;
                best_base = b;
                best_f = i;
            }
        }
    }
    // warp arg-min by (df, field, list start): every lane ends up with the same list
    for (int d = 16; d > 0; d >>= 1) {
        const long long odf = shfl_ll(best_df, lane ^ d);
        const long long ob = shfl_ll(best_base, lane ^ d);
        const int of = __shfl_xor_sync(0xFFFFFFFFu, best_f, d);
        const bool take = of >= 0 && (best_f < 0 || odf < best_df ||
                                      (odf == best_df && (of < best_f || (of == best_f && ob < best_base))));
        if (take) {
            best_df = odf;
            best_base = ob;
            best_f = of;
        }
    }
    if (best_f < 0) return;  // no list holds k documents: no seed
    const int n_seed = (int)(best_df < kSeedMax ? best_df : kSeedMax);
    const long long stride = best_df / n_seed;
    unsigned long long *row = a.c.cand_key + (size_t)q * (size_t)a.c.cap;
    const int32_t *ind = nullptr;
    for (int i = 0; i < a.c.n_fields; i++)
        if (i == best_f) ind = a.f[i].indices;
    for (int s = lane; s < n_seed; s += 32) row[s] = (unsigned long long)(uint32_t)ind[best_base + (long long)s * stride];
    if (lane == 0) a.c.cand_cnt[q] = (unsigned int)n_seed;
}

// ---------------------------------------------------------------------------------
// exact evaluation + selection
// ---------------------------------------------------------------------------------
enum { FSEL_KEEP = 0, FSEL_FINAL = 1, FSEL_SEED = 2 };

struct FusedSelectArgs {
    FusedCommon c;
    FField f[kMaxFields];
    const int32_t *q_list;
    int n_list;
    const unsigned int *n_list_ptr;
    int mode;
    unsigned int *n_over;
    int32_t *over_list;
    int64_t *out_ids;
    double *out_probs;
    unsigned long long *n_cand_total;
};

__device__ __forceinline__ double fused_final(double acc, const FusedCommon &c) {  // fusion.py:260-279, as fuse_step
    return d_sigmoid(c.unweighted ? (acc / (double)c.n_signals) * c.scale : c.scale * acc);
}

// Largest safe traversal threshold for "k documents with fused >= fk exist": documents whose exact
// conjunction sum lies below the returned (shifted) value have a strictly smaller fused probability.
// The final sigmoid rounds, so distinct sums can share one fp64 value; the step-down leaves that plateau.
__device__ inline double fused_theta_real(unsigned long long fk_bits, const FusedCommon &c) {
    const double fk = __longlong_as_double((long long)fk_bits);
    const double l = d_logit(fk);
    double L = c.unweighted ? (l / c.scale) * (double)c.n_signals : l / c.scale;
    double delta = 1e-9 * fmax(1.0, fabs(L));
    bool ok = false;
    for (int it = 0; it < 80; it++) {
        const unsigned long long fb = (unsigned long long)__double_as_longlong(fused_final(L, c));
        if (fb + 4ull <= fk_bits) {
            ok = true;
            break;
        }
        L -= delta;
        delta *= 2.0;
    }
    if (!ok) return 0.0;
    return L + kNeg * c.c_total;
}
__device__ inline float fused_theta(unsigned long long fk_bits, const FusedCommon &c) {
    const double th = fused_theta_real(fk_bits, c) * (1.0 - 2e-5) - 1e-4;
    return th > 0.0 ? __double2float_rd(th) : 0.f;
}

// order: fused desc, then doc id asc
__device__ __forceinline__ bool before(unsigned long long fa, uint32_t ia, unsigned long long fb, uint32_t ib) {
    return fa > fb || (fa == fb && ia < ib);
}

template <int NT>
__device__ __forceinline__ void fused_select_one(const FusedSelectArgs &a, unsigned char *smem, const int q) {
    unsigned long long *ef = reinterpret_cast<unsigned long long *>(smem);
    uint32_t *eid = reinterpret_cast<uint32_t *>(ef + kFusedSortCap);
    unsigned int *st = reinterpret_cast<unsigned int *>(eid + kFusedSortCap);
    const int tid = threadIdx.x;
    const int k = a.c.k, cap = a.c.cap, kpad = a.c.kpad;
    const unsigned int n_raw = a.c.cand_cnt[q];
    const int n = (int)min(n_raw, (unsigned)cap);
    const bool overflow = n_raw > (unsigned)cap;
    const int nb = a.mode == FSEL_SEED ? 0 : (int)a.c.best_n[q];
    const unsigned long long *row = a.c.cand_key + (size_t)q * (size_t)cap;
    if (n == 0 && a.mode == FSEL_KEEP) return;  // nothing new, nothing to write
    for (int i = tid; i < nb; i += NT) {
        ef[i] = a.c.best_f[(size_t)q * kpad + i];
        eid[i] = a.c.best_id[(size_t)q * kpad + i];
    }
    __syncthreads();

    // ---- exact evaluation of the new candidates: groups of g lanes, lane j looks term j up ----
    int mmax = 1;
    int m_i[kMaxFields];
    long long t0_i[kMaxFields];
#pragma unroll
    for (int i = 0; i < kMaxFields; i++) {
        m_i[i] = 0;
        t0_i[i] = 0;
        if (i < a.c.n_fields) {
            t0_i[i] = a.f[i].q_off[q];
            long long m = (long long)a.f[i].q_off[q + 1] - t0_i[i];
            m_i[i] = (int)(m < 0 ? 0 : (m > 32 ? 32 : m));
            mmax = max(mmax, m_i[i]);
        }
    }
    int g = 1;
    while (g < mmax) g <<= 1;
    const int lane = tid & 31;
    const int sub = lane / g, j = lane % g;
    const int groups = (NT / 32) * (32 / g);
    const int gid = (tid >> 5) * (32 / g) + sub;
    const unsigned gmask = g == 32 ? 0xFFFFFFFFu : ((1u << g) - 1u);
    // this lane's term in every field
    long long ip[kMaxFields];
    longlong2 trow[kMaxFields];
    int slot[kMaxFields];
    bool isdup[kMaxFields];
#pragma unroll
    for (int i = 0; i < kMaxFields; i++) {
        ip[i] = 0;
        trow[i] = make_longlong2(0, -1);
        slot[i] = -1;
        isdup[i] = false;
        if (i < a.c.n_fields && j < m_i[i]) {
            const longlong2 info = a.f[i].qt_info[2 * (t0_i[i] + j)];
            ip[i] = info.x;
            slot[i] = (int)(info.y >> 32);  // any value row: hot or lookup
            trow[i] = a.f[i].qt_info[2 * (t0_i[i] + j) + 1];
            isdup[i] = a.f[i].q_nocount[t0_i[i] + j] != 0;
        }
    }
    // The row is consumed in chunks of at most (sort slots - k) candidates: evaluate a chunk, merge it with the
    // k best so far (one bitonic sort of the survivors), keep the k best -- the sets are disjoint, so nothing
    // below the running k-th value can matter later.
    const int chunk = kFusedSortCap - kpad;
    int n_cur = nb;  // entries kept so far, sorted
    for (int c0 = 0; c0 == 0 || c0 < n; c0 += chunk) {
        const int nc = min(chunk, n - c0);
        const unsigned long long prev_kth = n_cur >= k ? ef[k - 1] : 0ull;
        __syncthreads();
        if (tid == 0) st[0] = (unsigned int)n_cur;
        __syncthreads();
        for (int base = 0; base < nc; base += groups) {
            const bool act = base + gid < nc;
            const uint32_t doc = act ? (uint32_t)(row[c0 + base + gid] & 0xFFFFFFFFull) : 0u;
            double acc = 0.0;
            int sig = 0;
    #pragma unroll
            for (int i = 0; i < kMaxFields; i++) {
                if (i >= a.c.n_fields) continue;
                const FField &ff = a.f[i];
                float val = 0.f;
                bool present = false;
                if (act && j < m_i[i]) {
                    if (slot[i] >= 0) {
                        val = row_value(ff, slot[i], doc);
                        present = __float_as_uint(val) != 0x80000000u;
                    } else {
                        const uint2 ent = tab_lookup(ff.tab, trow[i], (int)(doc / (uint32_t)kBlockDocs));
                        const int len = (int)(ent.y & kBlkLenMask);
                        if (len) {
                            const long long s0 = ip[i] + (long long)ent.x;
                            const int pos = slice_find(ff.indices + s0, len, doc, doc & ~(uint32_t)(kBlockDocs - 1));
                            if (pos >= 0) {
                                val = ff.data[s0 + pos];
                                present = true;
                            }
                        }
                    }
                }
                // bm25s's fp32 sum: the field's terms in query order (absent terms add +-0.0f)
                float sc = 0.f;
                for (int u = 0; u < m_i[i]; u++) sc = __fadd_rn(sc, __shfl_sync(0xFFFFFFFFu, val, sub * g + u));
                const unsigned pb = __ballot_sync(0xFFFFFFFFu, present && !isdup[i]);
                const int tfc = __popc((pb >> (sub * g)) & gmask);  // scorer.py:592-601
                if (act && j == 0) {
                    const double p = d_doc_probability(ff.params, sc, tfc, ff.doc_len[doc], ff.avgdl);
                    const FuseSpec fs{ff.cw, a.c.scale, a.c.n_signals,
                                      (sig == 0 ? 1 : 0) | (sig == a.c.n_signals - 1 ? 2 : 0) | (a.c.unweighted ? 4 : 0)};
                    acc = fuse_step(acc, d_logit(p), fs);
                }
                sig++;
            }
            if (a.c.has_cos && act && j == 0) {
                const double pc = clamp_prob((1.0 + (double)a.c.cosine[(size_t)q * (size_t)a.c.cos_stride + doc]) / 2.0);
                const FuseSpec fs{a.c.cos_w, a.c.scale, a.c.n_signals, (sig == 0 ? 1 : 0) | 2 | (a.c.unweighted ? 4 : 0)};
                acc = fuse_step(acc, d_logit(pc), fs);
            }
            if (act && j == 0) {
                const unsigned long long fb = (unsigned long long)__double_as_longlong(acc);
                if (fb >= prev_kth) {
                    const unsigned int pos = atomicAdd(&st[0], 1u);
                    ef[pos] = fb;
                    eid[pos] = doc;
                }
            }
        }
        __syncthreads();
        const int n_new = (int)st[0];
        if (n_new > n_cur) {
            const int n_tot = n_new;
            int P = 2;
            while (P < n_tot) P <<= 1;
            for (int i = n_tot + tid; i < P; i += NT) {
                ef[i] = 0ull;
                eid[i] = 0xFFFFFFFFu;
            }
            __syncthreads();
            for (int size = 2; size <= P; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = tid; t < (P >> 1); t += NT) {
                        const int lo = 2 * t - (t & (stride - 1));
                        const int hi = lo + stride;
                        const bool desc = (lo & size) == 0;
                        const unsigned long long fx = ef[lo], fy = ef[hi];
                        const uint32_t ix = eid[lo], iy = eid[hi];
                        if (before(fy, iy, fx, ix) == desc) {
                            ef[lo] = fy;
                            ef[hi] = fx;
                            eid[lo] = iy;
                            eid[hi] = ix;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        n_cur = min(n_new, k);
        __syncthreads();
    }
    const int n_tot = n_cur;
    const bool have_k = n_tot >= k;
    if (overflow) {
        // the row holds an arbitrary subset: use it for a tighter threshold only, evaluate the group again
        if (tid == 0) {
            if (have_k) {
                const float th = fused_theta(ef[k - 1], a.c);
                if (th > a.c.thr[q]) a.c.thr[q] = th;
            }
            a.c.cand_cnt[q] = 0u;
            a.over_list[atomicAdd(a.n_over, 1u)] = q;
        }
        return;
    }
    if (a.mode == FSEL_SEED) {
        if (tid == 0) {
            if (have_k) a.c.thr[q] = fused_theta(ef[k - 1], a.c);
            a.c.cand_cnt[q] = 0u;
        }
        return;
    }
    const int nk = min(n_tot, k);
    for (int r = tid; r < nk; r += NT) {
        a.c.best_f[(size_t)q * kpad + r] = ef[r];
        a.c.best_id[(size_t)q * kpad + r] = eid[r];
    }
    if (tid == 0) {
        a.c.best_n[q] = (unsigned int)nk;
        a.c.cand_cnt[q] = 0u;
        if (have_k) {
            const float th = fused_theta(ef[k - 1], a.c);
            if (th > a.c.thr[q]) a.c.thr[q] = th;
        }
        if (a.n_cand_total) atomicAdd(a.n_cand_total, (unsigned long long)n);
    }
    if (a.mode != FSEL_FINAL) return;
    for (int r = tid; r < nk; r += NT) {
        const size_t o = (size_t)q * (size_t)k + r;
        a.out_ids[o] = (int64_t)eid[r] + a.c.doc_id_offset;
        a.out_probs[o] = __longlong_as_double((long long)ef[r]);
    }
    if (tid == 0) {
        // Fewer than k candidates: the tail is made of documents no field matches -- the dense guaranteed
        // path fills it.  Without a dense signal the traversal never looks at such documents; they all
        // share one fused value (every signal at logit(1e-10)), so the k-th value must beat it strictly.
        bool fb = !have_k;
        if (!fb && !a.c.has_cos) {
            double acc = 0.0;
            for (int s_ = 0; s_ < a.c.n_signals; s_++) {
                const FuseSpec fs{a.f[s_ < kMaxFields ? s_ : 0].cw, a.c.scale, a.c.n_signals,
                                  (s_ == 0 ? 1 : 0) | (s_ == a.c.n_signals - 1 ? 2 : 0) | (a.c.unweighted ? 4 : 0)};
                acc = fuse_step(acc, d_logit(0.0), fs);
            }
            fb = !(ef[k - 1] > (unsigned long long)__double_as_longlong(acc));
        }
        if (fb) a.c.flags[q] = 1;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) fused_select_kernel(const __grid_constant__ FusedSelectArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int n_list = a.n_list_ptr ? *a.n_list_ptr : (unsigned int)a.n_list;
    for (unsigned int b = blockIdx.x; b < n_list; b += gridDim.x) {
        fused_select_one<NT>(a, smem, a.q_list ? a.q_list[b] : (int)b);
        __syncthreads();
    }
}

__global__ void fused_mark_bad_kernel(const int32_t *__restrict__ list, const unsigned int *__restrict__ n_list,
                                      uint8_t *__restrict__ flags) {
    const unsigned int n = *n_list;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flags[list[i]] = 1;
}

struct FusedReport {
    unsigned long long n_cand, units, skipped, abandoned, restricted, ne_skipped, sparse_units, sparse_docs;
    unsigned int n_fallback, reruns;
    int err;
    int pad;
};
__global__ void fused_report_kernel(const uint8_t *__restrict__ flags, int64_t n_q, int32_t *__restrict__ fb_list,
                                    const unsigned long long *__restrict__ n_cand, const unsigned long long *__restrict__ stats,
                                    const int *__restrict__ err, const unsigned int *__restrict__ round_cnt, int n_round_cnt,
                                    FusedReport *out) {
    __shared__ unsigned int s_n;
    if (threadIdx.x == 0) s_n = 0u;
    __syncthreads();
    for (int64_t q = threadIdx.x; q < n_q; q += blockDim.x)
        if (flags[q]) fb_list[atomicAdd(&s_n, 1u)] = (int32_t)q;
    __syncthreads();
    if (threadIdx.x == 0) {
        FusedReport r;
        r.n_cand = *n_cand;
        r.units = stats[0];
        r.skipped = stats[1];
        r.abandoned = stats[2];
        r.restricted = stats[3];
        r.ne_skipped = stats[4];
        r.sparse_units = stats[5];
        r.sparse_docs = stats[6];
        r.n_fallback = s_n;
        unsigned int re = 0;
        for (int i = 0; i < n_round_cnt; i++) re += round_cnt[i];
        r.reruns = re;
        r.err = *err;
        r.pad = 0;
        *out = r;
    }
}
__global__ void add_offset_kernel(int64_t *ids, int n, int64_t off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] += off;
}

template <int F, bool HAS_COS, bool PAIR>
static int launch_fused_block_inst(const bb25_index *idx, const FusedBlockArgs &a, cudaStream_t st) {
    constexpr int kAccBytes = (F >= 2 ? 2 : 1) * kBlockDocs * 4;
    constexpr int kWarpBytes = kAccBytes + FQC * 8 + FQC * F * 16;
    const size_t smem = (size_t)FWARPS * kWarpBytes;
    BB25_CUDA(cudaFuncSetAttribute(fused_block_kernel<F, HAS_COS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_chunks = (a.n_q + FQC - 1) / FQC;
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    if (n_items <= 0) return 0;
    long long grid = (long long)idx->sm_count * (F >= 2 ? 3 : 5);
    const long long need = (n_items + FWARPS - 1) / FWARPS;
    if (grid > need) grid = need;
    BB25_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    fused_block_kernel<F, HAS_COS, PAIR><<<(unsigned)grid, FWARPS * 32, smem, st>>>(a);
    BB25_LAUNCH_CHECK();
    return 0;
}
static int launch_fused_block(const bb25_index *idx, const FusedBlockArgs &a, cudaStream_t st) {
    const bool hc = a.c.has_cos != 0;
    if (a.unit_mask) {
        // the light kernel first: skips / evaluates through essential postings, leaves the pass units in the mask
        const int n_chunks = (a.n_q + FQC - 1) / FQC;
        const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
        if (n_items <= 0) return 0;
        long long grid = (long long)idx->sm_count * BB25_GROUP_CTAS;
        const long long need = (n_items + GWARPS - 1) / GWARPS;
        if (grid > need) grid = need;
        BB25_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
        fused_group_kernel<<<(unsigned)grid, GWARPS * 32, 0, st>>>(a);
        BB25_LAUNCH_CHECK();
    }
    switch (a.c.n_fields) {
    case 1: return hc ? launch_fused_block_inst<1, true, false>(idx, a, st) : launch_fused_block_inst<1, false, false>(idx, a, st);
    case 2:
        if (a.pair_mode) return hc ? launch_fused_block_inst<2, true, true>(idx, a, st) : launch_fused_block_inst<2, false, true>(idx, a, st);
        return hc ? launch_fused_block_inst<2, true, false>(idx, a, st) : launch_fused_block_inst<2, false, false>(idx, a, st);
    case 3: return hc ? launch_fused_block_inst<3, true, false>(idx, a, st) : launch_fused_block_inst<3, false, false>(idx, a, st);
    case 4: return hc ? launch_fused_block_inst<4, true, false>(idx, a, st) : launch_fused_block_inst<4, false, false>(idx, a, st);
    default: set_error("unsupported number of fields %d", a.c.n_fields); return 1;
    }
}

}  // namespace bb25

using namespace bb25;

extern "C" {

int bb25_retrieve_fused_batch(int n_fields, const bb25_fused_field *fields, const float *cosine, int64_t cos_stride,
                              double cos_weight, int weighted, double scale, int64_t n_queries, int k,
                              int64_t *out_ids, double *out_probs, void *stream) {
    if (n_fields < 1 || n_fields > kMaxFields || !fields) { set_error("need 1..%d BM25 fields", kMaxFields); return 1; }
    if (n_queries < 0 || !out_ids || !out_probs) { set_error("bad arguments"); return 1; }
    if (!(scale > 0.0)) { set_error("scale must be > 0"); return 1; }
    bb25_index *idx0 = fields[0].index;
    if (!idx0) { set_error("index is NULL"); return 1; }
    const int has_cos = cosine != nullptr;
    for (int i = 0; i < n_fields; i++) {
        const bb25_fused_field &f = fields[i];
        if (!f.index || f.index->device != idx0->device || f.index->n_docs != idx0->n_docs ||
            f.index->doc_id_offset != idx0->doc_id_offset) {
            set_error("every field must index the same documents on the same device");
            return 1;
        }
        if (f.params.has_base_rate && !(f.params.base_rate > 0.0 && f.params.base_rate < 1.0)) {
            set_error("base_rate must be in (0, 1)");
            return 1;
        }
        if (f.params.prior_mode != 0 && f.params.prior_mode != 1) { set_error("prior_mode must be 0 or 1"); return 1; }
        if (!(f.params.alpha > 0.0)) { set_error("the fused batch path needs alpha > 0 in every field (the bound is monotone in the score)"); return 1; }
        if (weighted && !(f.weight >= 0.0)) { set_error("weights must be >= 0 (fusion.py:253-255)"); return 1; }
        if (!f.q_off || f.n_terms_total < 0 || f.term_base < 0 || (f.n_terms_total > 0 && !f.q_terms)) { set_error("bad query arrays"); return 1; }
    }
    if (has_cos && (cos_stride < idx0->n_docs || (cos_stride & 3) || ((uintptr_t)cosine & 15))) {
        set_error("cosine rows must be 16-byte aligned with a stride that is a multiple of 4 and >= n_docs");
        return 1;
    }
    if (has_cos && weighted && !(cos_weight >= 0.0)) { set_error("weights must be >= 0 (fusion.py:253-255)"); return 1; }
    if (k < 1 || k > kFusedMaxK || (int64_t)k > idx0->n_docs) {
        set_error("k must satisfy 1 <= k <= min(n_docs, %d), got %d", kFusedMaxK, k);
        return 1;
    }
    if (n_queries == 0) return 0;
    if (n_queries > 0x7FFFFFF0ll) { set_error("too many queries in one batch"); return 1; }
    DeviceGuard g(idx0->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_q = n_queries;
    const int n_signals = n_fields + has_cos;
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    const int cap = kFusedRowCap;
    int n_rounds = 3;
    if (const char *e = getenv("BB25_REPAIR_ROUNDS")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 6) n_rounds = v;
    }

    int sparse_mode = 1;
    if (const char *e = getenv("BB25_FUSED_SPARSE")) sparse_mode = atoi(e) != 0 ? 1 : 0;
    if (n_fields == 2 && !has_cos && idx0->prune >= 2 && sparse_mode) {
        // value rows of the mid-frequency terms for the essential-posting evaluation (built once per index)
        for (int i = 0; i < n_fields; i++) {
            std::lock_guard<std::mutex> lock(fields[i].index->mu);
            if (ensure_lookup_rows(fields[i].index, st)) return 1;
        }
    }

    std::vector<int32_t> fb_list;
    std::vector<std::vector<int32_t>> h_terms((size_t)n_fields);
    std::vector<std::vector<int64_t>> h_qo((size_t)n_fields);
    {
        std::lock_guard<std::mutex> lock(idx0->mu);
        ws_acquire(idx0, st);
        // ---- workspace -------------------------------------------------------------------
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes); return o; };
        size_t o_terms[kMaxFields], o_nc[kMaxFields], o_qo[kMaxFields], o_info[kMaxFields], o_qcst[kMaxFields], o_ptf[kMaxFields];
        for (int i = 0; i < n_fields; i++) {
            const size_t nt = (size_t)std::max<int64_t>(fields[i].n_terms_total, 1);
            o_terms[i] = take(sizeof(int32_t) * nt);
            o_nc[i] = take(nt);
            o_qo[i] = take(sizeof(int64_t) * (size_t)(n_q + 1));
            o_info[i] = take(2 * sizeof(longlong2) * nt);
            o_qcst[i] = take(sizeof(float) * (size_t)n_q);
            o_ptf[i] = take(sizeof(float) * (size_t)n_q);
        }
        const size_t o_thr = take(sizeof(float) * (size_t)n_q);
        const size_t o_cnt = take(sizeof(unsigned int) * (size_t)n_q);
        const size_t o_bn = take(sizeof(unsigned int) * (size_t)n_q);
        const size_t o_flags = take((size_t)n_q);
        const size_t o_la = take(sizeof(int32_t) * (size_t)n_q);
        const size_t o_lb = take(sizeof(int32_t) * (size_t)n_q);
        const size_t o_fbl = take(sizeof(int32_t) * (size_t)n_q);
        const size_t o_ctr = take(1024);
        const size_t o_bf = take(sizeof(unsigned long long) * (size_t)n_q * (size_t)kpad);
        const size_t o_bi = take(sizeof(uint32_t) * (size_t)n_q * (size_t)kpad);
        const size_t o_key = take(sizeof(unsigned long long) * (size_t)n_q * (size_t)cap);
        const bool use_mask = n_fields == 2 && !has_cos && idx0->prune >= 2 && sparse_mode;
        const size_t o_mask = take(use_mask ? (size_t)idx0->n_blocks * (size_t)((n_q + FQC - 1) / FQC) : 1);
        if (ensure_workspace(idx0, off)) return 1;
        unsigned char *ws = (unsigned char *)idx0->ws;
        unsigned long long *d_work = (unsigned long long *)(ws + o_ctr);
        int *d_err = (int *)(ws + o_ctr + 16);
        unsigned long long *d_ncand = (unsigned long long *)(ws + o_ctr + 24);
        unsigned long long *d_stats = (unsigned long long *)(ws + o_ctr + 32);  // [8]
        unsigned int *d_round = (unsigned int *)(ws + o_ctr + 128);             // [stage][round]
        const int kRoundStride = 8;
        FusedReport *d_report = (FusedReport *)(ws + o_ctr + 512);
        int32_t *d_list[2] = {(int32_t *)(ws + o_la), (int32_t *)(ws + o_lb)};
        int32_t *d_fbl = (int32_t *)(ws + o_fbl);
        BB25_CUDA(cudaMemsetAsync(ws + o_ctr, 0, 1024, st));

        // ---- arguments -------------------------------------------------------------------
        FusedCommon c{};
        c.n_fields = n_fields;
        c.has_cos = has_cos;
        c.unweighted = weighted ? 0 : 1;
        c.n_signals = n_signals;
        c.scale = scale;
        c.cos_w = has_cos ? (weighted ? cos_weight : 1.0) : 0.0;
        c.cosine = cosine;
        c.cos_stride = cos_stride;
        c.n_docs = idx0->n_docs;
        c.doc_id_offset = idx0->doc_id_offset;
        c.k = k;
        c.kpad = kpad;
        c.cap = cap;
        c.thr = (float *)(ws + o_thr);
        c.cand_cnt = (unsigned int *)(ws + o_cnt);
        c.cand_key = (unsigned long long *)(ws + o_key);
        c.best_f = (unsigned long long *)(ws + o_bf);
        c.best_id = (uint32_t *)(ws + o_bi);
        c.best_n = (unsigned int *)(ws + o_bn);
        c.flags = ws + o_flags;
        double c_total = c.cos_w;
        FField ff[kMaxFields] = {};
        for (int i = 0; i < n_fields; i++) {
            const bb25_index *ix = fields[i].index;
            FField &f = ff[i];
            f.data = ix->data;
            f.indices = ix->indices;
            f.indptr = ix->indptr;
            f.tab = BlockTable{ix->tab_ent, ix->tab_bits, ix->tab_row};
            f.dense_vals = ix->dense_vals;
            f.lookup_vals = ix->lookup_vals;
            f.n_hot = ix->dense_vals ? ix->n_dense : 0;
            f.dense_stride = ix->dense_stride;
            f.doc_len = ix->doc_len;
            f.avgdl = ix->avgdl;
            f.params = fields[i].params;
            f.cw = weighted ? fields[i].weight : 1.0;
            f.ka = (float)(f.cw * f.params.alpha);
            if ((double)f.ka < f.cw * f.params.alpha) f.ka = nextafterf(f.ka, 3.4e38f);
            f.ka = nextafterf(f.ka, 3.4e38f);
            f.q_terms = (const int32_t *)(ws + o_terms[i]);
            f.q_nocount = ws + o_nc[i];
            f.q_off = (const int64_t *)(ws + o_qo[i]);
            f.qt_info = (const longlong2 *)(ws + o_info[i]);
            f.q_o = (float *)(ws + o_qcst[i]);
            f.q_ptf = (float *)(ws + o_ptf[i]);
            {
                // constants of the per-document refinement (emit path), all rounded up
                const bb25_params &pp = f.params;
                const double lbr = pp.has_base_rate ? log(pp.base_rate / (1.0 - pp.base_rate)) : 0.0;
                auto up = [](double x) { float y = (float)x; if ((double)y < x) y = nextafterf(y, 3.4e38f); return nextafterf(y, 3.4e38f); };
                f.cwf = up(f.cw);
                f.basef = up(-pp.alpha * pp.beta + lbr + kNeg);
                f.ef = up(lbr > 0.0 ? lbr : 0.0);
                f.inv_avgdl = (float)(1.0 / ix->avgdl);
                // posterior clamps at 1 - 1e-10 BEFORE the base-rate step (probability.py:163-168)
                f.ucap = up(f.cw * (2.0 * kNeg + (lbr < 0.0 ? lbr : 0.0)) * (1.0 + 1e-6) + 1e-5);
            }
            c_total += f.cw;
            if (prep_queries_launch(ix, fields[i].q_terms, fields[i].q_off, n_q, fields[i].term_base, fields[i].n_terms_total,
                                    (int32_t *)(ws + o_terms[i]), ws + o_nc[i], (int64_t *)(ws + o_qo[i]),
                                    (longlong2 *)(ws + o_info[i]), d_err, st))
                return 1;
        }
        c.c_total = c_total;
        c.kc = nextafterf((float)c.cos_w, 3.4e38f);
        c.cos_ub = has_cos ? nextafterf((float)(c.cos_w * 2.0 * 23.02586 + 4e-3), 3.4e38f) : 0.f;

        FusedPrepArgs pa{};
        pa.c = c;
        for (int i = 0; i < kMaxFields; i++) pa.f[i] = ff[i];
        pa.n_q = n_q;
        fused_prep_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(pa);
        BB25_LAUNCH_CHECK();

        FusedSelectArgs sa{};
        sa.c = c;
        for (int i = 0; i < kMaxFields; i++) sa.f[i] = ff[i];
        sa.out_ids = out_ids;
        sa.out_probs = out_probs;
        sa.n_cand_total = d_ncand;
        const size_t sel_smem = (size_t)kFusedSortCap * 12 + 64;
        BB25_CUDA(cudaFuncSetAttribute(fused_select_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
        const unsigned repair_grid = (unsigned)std::min<int64_t>(n_q, (int64_t)idx0->sm_count * 2);

        // ---- threshold seeds ---------------------------------------------------------------
        int use_seed = 1;
        if (const char *e = getenv("BB25_FUSED_SEED")) use_seed = atoi(e);
        if (use_seed) {
            fused_seed_kernel<<<(unsigned)((n_q + 7) / 8), 256, 0, st>>>(pa);
            BB25_LAUNCH_CHECK();
            sa.q_list = nullptr;
            sa.n_list = (int)n_q;
            sa.n_list_ptr = nullptr;
            sa.mode = FSEL_SEED;
            sa.n_over = d_round;  // not reached: seeds never overflow the row
            sa.over_list = d_list[0];
            fused_select_kernel<512><<<(unsigned)n_q, 512, sel_smem, st>>>(sa);
            BB25_LAUNCH_CHECK();
        }

        // ---- block groups --------------------------------------------------------------------
        const int T = idx0->n_blocks;
        int bounds[4] = {0, T, T, T};
        int ng = 1;
        if (T >= 64) {
            bounds[1] = std::max(1, T / 64);
            bounds[2] = std::max(bounds[1] + 1, T / 8);
            bounds[3] = T;
            ng = 3;
        } else if (T >= 8) {
            bounds[1] = std::max(1, T / 8);
            bounds[2] = T;
            ng = 2;
        }
        FusedBlockArgs ba{};
        ba.c = c;
        {
            int order[kMaxFields] = {0, 1, 2, 3};
            std::sort(order, order + n_fields, [&](int x, int y) {
                return fields[x].index->nnz != fields[y].index->nnz ? fields[x].index->nnz < fields[y].index->nnz : x < y;
            });
            for (int i = 0; i < n_fields; i++) ba.f[i] = ff[order[i]];
        }
        ba.prune = idx0->prune;
        ba.pair_mode = n_fields == 2 ? 1 : 0;
        if (const char *e = getenv("BB25_FUSED_PAIR")) ba.pair_mode = (n_fields == 2 && atoi(e) != 0) ? 1 : 0;
        ba.sparse_mode = sparse_mode;
        ba.unit_mask = (use_mask && ba.pair_mode) ? ws + o_mask : nullptr;
        ba.work_counter = d_work;
        ba.stats = d_stats;
        float trav_ms = 0.f;
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
        for (int gi = 0; gi < ng; gi++) {
            unsigned int *rc = d_round + (size_t)(gi + 1) * kRoundStride;
            for (int r = 0; r <= n_rounds; r++) {
                const int32_t *list = r == 0 ? nullptr : d_list[(r - 1) & 1];
                const unsigned int *n_list = r == 0 ? nullptr : rc + r;
                ba.q_list = list;
                ba.n_q = (int)n_q;
                ba.n_q_ptr = n_list;
                ba.blk_begin = bounds[gi];
                ba.blk_end = bounds[gi + 1];
                cudaEvent_t a0 = nullptr, a1 = nullptr;
                if (evs.size() < 16 && cudaEventCreate(&a0) == cudaSuccess && cudaEventCreate(&a1) == cudaSuccess) cudaEventRecord(a0, st);
                if (launch_fused_block(idx0, ba, st)) return 1;
                if (a0 && a1) {
                    cudaEventRecord(a1, st);
                    evs.emplace_back(a0, a1);
                }
                sa.q_list = list;
                sa.n_list = (int)n_q;
                sa.n_list_ptr = n_list;
                sa.mode = gi == ng - 1 ? FSEL_FINAL : FSEL_KEEP;
                sa.over_list = d_list[r & 1];
                sa.n_over = rc + r + 1;
                fused_select_kernel<512><<<r == 0 ? (unsigned)n_q : repair_grid, 512, sel_smem, st>>>(sa);
                BB25_LAUNCH_CHECK();
            }
            fused_mark_bad_kernel<<<4, 256, 0, st>>>(d_list[n_rounds & 1], rc + n_rounds + 1, c.flags);
            BB25_LAUNCH_CHECK();
        }
        fused_report_kernel<<<1, 1024, 0, st>>>(c.flags, n_q, d_fbl, d_ncand, d_stats, d_err, d_round, (ng + 1) * kRoundStride,
                                                d_report);
        BB25_LAUNCH_CHECK();
        FusedReport *h_rep = (FusedReport *)idx0->pinned;
        BB25_CUDA(cudaMemcpyAsync(h_rep, d_report, sizeof(FusedReport), cudaMemcpyDeviceToHost, st));
        BB25_CUDA(cudaStreamSynchronize(st));
        idx0->fz_syncs = 1;
        for (auto &pr : evs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) trav_ms += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        if (h_rep->err) {
            set_error("invalid query batch (flags=%d: 1 term id out of range, 2 q_off not monotone / outside the batch)", h_rep->err);
            return 1;
        }
        idx0->fz_units = (int64_t)h_rep->units;
        idx0->fz_skipped = (int64_t)h_rep->skipped;
        idx0->fz_abandoned = (int64_t)h_rep->abandoned + (int64_t)h_rep->restricted;
        idx0->fz_ne_skipped = (int64_t)h_rep->ne_skipped;
        idx0->fz_sparse_units = (int64_t)h_rep->sparse_units;
        idx0->fz_sparse_docs = (int64_t)h_rep->sparse_docs;
        idx0->fz_candidates = (int64_t)h_rep->n_cand;
        idx0->fz_fallback = (int64_t)h_rep->n_fallback;
        idx0->fz_reruns = (int64_t)h_rep->reruns;
        idx0->fz_traverse_ms = (double)trav_ms;
        const unsigned int n_fb = h_rep->n_fallback;
        if (n_fb > 0) {
            // the flagged queries' terms, back on the host for the dense path
            fb_list.resize(n_fb);
            BB25_CUDA(cudaMemcpyAsync(fb_list.data(), d_fbl, sizeof(int32_t) * n_fb, cudaMemcpyDeviceToHost, st));
            for (int i = 0; i < n_fields; i++) {
                h_terms[i].resize((size_t)std::max<int64_t>(fields[i].n_terms_total, 1));
                h_qo[i].resize((size_t)n_q + 1);
                if (fields[i].n_terms_total > 0)
                    BB25_CUDA(cudaMemcpyAsync(h_terms[i].data(), ws + o_terms[i], sizeof(int32_t) * (size_t)fields[i].n_terms_total,
                                              cudaMemcpyDeviceToHost, st));
                BB25_CUDA(cudaMemcpyAsync(h_qo[i].data(), ws + o_qo[i], sizeof(int64_t) * (size_t)(n_q + 1), cudaMemcpyDeviceToHost, st));
            }
            BB25_CUDA(cudaStreamSynchronize(st));
            idx0->fz_syncs++;
        }
        ws_release(idx0, st);
    }
    // ---- dense guaranteed path for the flagged queries (takes each index's own lock) ----------
    if (!fb_list.empty()) {
        double *acc = nullptr;
        int64_t *t_ids = nullptr;
        BB25_CUDA(cudaMalloc(&acc, sizeof(double) * (size_t)idx0->n_docs));
        int rc = 0;
        for (size_t u = 0; u < fb_list.size() && !rc; u++) {
            const int q = fb_list[u];
            for (int i = 0; i < n_fields && !rc; i++) {
                const int64_t t0 = h_qo[i][q];
                const int m = (int)std::max<int64_t>(0, h_qo[i][q + 1] - t0);
                const int flags = (i == 0 ? 1 : 0) | (i == n_signals - 1 ? 2 : 0) | (weighted ? 0 : 4);
                rc = bb25_fuse_bm25_signal(fields[i].index, &fields[i].params, h_terms[i].data() + t0, m,
                                           weighted ? fields[i].weight : 1.0, n_signals, scale, flags, acc, st);
            }
            if (!rc && has_cos)
                rc = bb25_fuse_cosine_signal(idx0->device, cosine + (size_t)q * (size_t)cos_stride, idx0->n_docs,
                                             weighted ? cos_weight : 1.0, n_signals, scale, 2 | (weighted ? 0 : 4), acc, st);
            if (!rc) rc = bb25_topk_f64(idx0->device, acc, idx0->n_docs, k, out_ids + (size_t)q * k, out_probs + (size_t)q * k, st);
            if (!rc && idx0->doc_id_offset) {
                add_offset_kernel<<<(k + 255) / 256, 256, 0, st>>>(out_ids + (size_t)q * k, k, idx0->doc_id_offset);
                count_launch();
            }
        }
        cudaStreamSynchronize(st);
        idx0->fz_syncs++;
        cudaFree(acc);
        cudaFree(t_ids);
        if (rc) return 1;
    }
    return 0;
}

int bb25_fused_stats(const bb25_index *idx, int64_t *units, int64_t *units_skipped, int64_t *units_abandoned,
                     int64_t *candidates, int64_t *fallback_queries, int64_t *rerun_queries, int64_t *host_syncs,
                     double *traverse_ms) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (units) *units = idx->fz_units;
    if (units_skipped) *units_skipped = idx->fz_skipped;
    if (units_abandoned) *units_abandoned = idx->fz_abandoned;
    if (candidates) *candidates = idx->fz_candidates;
    if (fallback_queries) *fallback_queries = idx->fz_fallback;
    if (rerun_queries) *rerun_queries = idx->fz_reruns;
    if (host_syncs) *host_syncs = idx->fz_syncs;
    if (traverse_ms) *traverse_ms = idx->fz_traverse_ms;
    return 0;
}

int bb25_fused_prune_stats(const bb25_index *idx, int64_t *units_no_essential, int64_t *units_sparse, int64_t *sparse_documents) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (units_no_essential) *units_no_essential = idx->fz_ne_skipped;
    if (units_sparse) *units_sparse = idx->fz_sparse_units;
    if (sparse_documents) *sparse_documents = idx->fz_sparse_docs;
    return 0;
}

}  // extern "C"
