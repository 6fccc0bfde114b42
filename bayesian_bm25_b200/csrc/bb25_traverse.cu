// Posting-list traversal (K1/K2), fused posterior epilogue (K3), threshold-driven
// candidate emission and per-query selection (K4).
//
// Layout of the work: the document space of a shard is cut into tiles of
// `tile_docs` documents.  One CTA owns one tile at a time: fp32 score accumulators
// and 8-bit matched-term counters for the tile live in shared memory, the CTA walks
// the (query, term) list of a chunk of queries, and for every term streams the slice
// of the posting list that falls inside the tile (located through the per-(term,tile)
// skip table) with 128-bit loads and adds it into the accumulators.  Terms of one
// query are applied in query order with a block barrier between them, so every
// document's fp32 sum is formed in exactly the order bm25s forms it (bit-exact).
// Work items are handed out tile-major through a global counter, so all CTAs work on
// the same one or two tiles for different queries and the tile's slice of the index
// is served from L2 after its first touch.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cuda_fp16.h>

#include "bb25_internal.cuh"
#include "bb25_device.cuh"

namespace bb25 {

constexpr int QB = 8;      // queries per work item
constexpr int MAXT = 128;  // (query, term) occurrences staged per metadata batch
constexpr int kMaxK = 4096;

enum { MODE_RETRIEVE = 0, MODE_SCORES = 1, MODE_PROBS = 2, MODE_FUSED = 3 };

struct TileArgs {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    const uint32_t *tile_off;
    const int32_t *doc_len;
    double avgdl;
    int n_tiles;
    int64_t n_docs;
    const int32_t *q_terms;    // sanitised copy, indexed by absolute position - term_base
    const uint8_t *q_nocount;  // 1 = duplicate occurrence: add the score, do not count
    const int64_t *q_off;
    int64_t term_base;
    const int32_t *q_list;  // NULL = queries [0, n_q)
    int n_q;
    const unsigned int *n_q_ptr;  // when set, the number of queries is read on the device (no host round trip)
    int tile_begin, tile_end;
    const unsigned long long *thr;
    unsigned int *cand_cnt;
    unsigned long long *cand_key;
    int cap;
    float *out_scores;
    double *out_probs;
    int64_t out_stride;
    bb25_params params;
    FuseSpec fuse;  // MODE_FUSED only
    unsigned long long *work_counter;
};

struct TileMeta {
    long long m_start[MAXT];
    unsigned long long s_thr[QB];
    long long s_t0[QB];
    long long item;
    uint32_t m_len[MAXT];
    int s_q[QB];
    int s_m[QB];
    int s_pre[QB + 1];
    uint8_t m_flags[MAXT];
    uint8_t m_slot[MAXT];
};

template <bool COUNT>
__device__ __forceinline__ void rmw1(float *acc, uint8_t *cnt, int off, float v) {
    acc[off] = __fadd_rn(acc[off], v);
    if (COUNT) {
        unsigned c = cnt[off];
        cnt[off] = (uint8_t)(c < 255u ? c + 1u : 255u);
    }
}

// Add postings [s, s+len) of one term into the tile accumulators.  Doc ids inside a
// posting list are distinct, so no two threads touch the same accumulator.  Every
// warp-wide load / accumulator update covers 32 CONSECUTIVE postings (128-byte aligned
// in the main loop): doc ids of a dense posting list are (nearly) consecutive, so the
// shared-memory read-modify-writes of a warp fall into distinct banks (a 128-bit load
// per lane would put lanes 4 documents apart: 4-way bank conflicts, measured in round 1).
template <int NT, bool COUNT, int U>
__device__ __forceinline__ void scatter_slice_s32(const float *__restrict__ data,
                                                  const int32_t *__restrict__ indices, long long s,
                                                  uint32_t len, float *acc, uint8_t *cnt, int doc_base,
                                                  int tid) {
    const long long e = s + (long long)len;
    long long a0 = (s + 31) & ~31ll;  // first 128-byte aligned element
    if (a0 > e) a0 = e;
    if (tid < (int)(a0 - s)) {
        long long j = s + tid;
        rmw1<COUNT>(acc, cnt, indices[j] - doc_base, data[j]);
    }
    const int n = (int)(e - a0);
    const int32_t *ip = indices + a0;
    const float *dp = data + a0;
    int j = tid;
    for (; j + (U - 1) * NT < n; j += U * NT) {
        int d[U];
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            d[u] = ld_nc_s32(ip + j + u * NT);
            v[u] = ld_nc_f32(dp + j + u * NT);
        }
#pragma unroll
        for (int u = 0; u < U; u++) rmw1<COUNT>(acc, cnt, d[u] - doc_base, v[u]);
    }
    for (; j < n; j += NT) rmw1<COUNT>(acc, cnt, ld_nc_s32(ip + j) - doc_base, ld_nc_f32(dp + j));
}

// Retrieve mode keeps no matched-term counters in the tile (the select kernel recovers tf
// for the k winners); the dense-output modes count matched terms per document.
// resident CTAs per SM the kernel is built for: 5 B/doc of shared memory with counters,
// 4 B/doc without
template <int D, int MODE>
constexpr int ctas_per_sm() {
    return D > 16384 ? 1 : (D > 8192 ? 1 : 2) * ((MODE == MODE_RETRIEVE) ? 3 : 2);
}

template <int D, int NT, int MODE>
__global__ void __launch_bounds__(NT, ctas_per_sm<D, MODE>()) tile_kernel(const __grid_constant__ TileArgs a) {
    constexpr bool HAS_CNT = (MODE != MODE_RETRIEVE);
    extern __shared__ __align__(16) unsigned char smem[];
    float *acc = reinterpret_cast<float *>(smem);
    uint8_t *cnt = smem + (size_t)D * 4;
    TileMeta *meta = reinterpret_cast<TileMeta *>(smem + (size_t)D * (HAS_CNT ? 5 : 4));
    const int tid = threadIdx.x;

    for (int i = tid; i < D / 4; i += NT) reinterpret_cast<float4 *>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HAS_CNT)
        for (int i = tid; i < D / 16; i += NT) reinterpret_cast<uint4 *>(cnt)[i] = make_uint4(0, 0, 0, 0);

    const int n_q = a.n_q_ptr ? (int)*a.n_q_ptr : a.n_q;
    const int n_chunks = (n_q + QB - 1) / QB;
    const long long n_items = (long long)(a.tile_end - a.tile_begin) * n_chunks;

    for (;;) {
        __syncthreads();  // previous item fully consumed (also covers the zero-fill above)
        if (tid == 0) meta->item = (long long)atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const long long item = meta->item;
        if (item >= n_items) break;
        const int tile = a.tile_begin + (int)(item / n_chunks);
        const int slot0 = (int)(item % n_chunks) * QB;
        const int nslots = min(QB, n_q - slot0);
        const int doc_base = tile * D;

        if (tid < nslots) {
            const int q = a.q_list ? a.q_list[slot0 + tid] : slot0 + tid;
            const long long t0 = a.q_off[q];
            meta->s_q[tid] = q;
            meta->s_t0[tid] = t0;
            meta->s_m[tid] = (int)max(0ll, (long long)a.q_off[q + 1] - t0);
            if (MODE == MODE_RETRIEVE) meta->s_thr[tid] = a.thr[q];
        }
        __syncthreads();
        if (tid == 0) {
            int p = 0;
            for (int i = 0; i < nslots; i++) {
                meta->s_pre[i] = p;
                p += meta->s_m[i];
            }
            meta->s_pre[nslots] = p;
        }
        __syncthreads();
        const int total = meta->s_pre[nslots];

        for (int b0 = 0; b0 < total; b0 += MAXT) {
            const int nb = min(MAXT, total - b0);
            if (tid < nb) {
                const int j = b0 + tid;
                int slot = 0;
                while (j >= meta->s_pre[slot + 1]) slot++;
                const long long pos = meta->s_t0[slot] + (j - meta->s_pre[slot]) - a.term_base;
                const int t = a.q_terms[pos];
                const uint32_t *to = a.tile_off + (size_t)t * (size_t)(a.n_tiles + 1) + tile;
                const uint32_t lo = to[0], hi = to[1];
                meta->m_start[tid] = a.indptr[t] + (long long)lo;
                meta->m_len[tid] = hi - lo;
                meta->m_flags[tid] = (uint8_t)((a.q_nocount[pos] ? 1 : 0) | ((j == meta->s_pre[slot + 1] - 1) ? 2 : 0));
                meta->m_slot[tid] = (uint8_t)slot;
            }
            __syncthreads();
            for (int i = 0; i < nb; i++) {
                const long long s = meta->m_start[i];
                const uint32_t len = meta->m_len[i];
                const int flags = meta->m_flags[i];
                if (len) {
                    const bool count = HAS_CNT && !(flags & 1);
                    if (count) scatter_slice_s32<NT, true, 4>(a.data, a.indices, s, len, acc, cnt, doc_base, tid);
                    else scatter_slice_s32<NT, false, 4>(a.data, a.indices, s, len, acc, cnt, doc_base, tid);
                }
                __syncthreads();
                if (flags & 2) {
                    // ---- last term of this query: fused epilogue over the tile ----
                    const int slot = meta->m_slot[i];
                    if (MODE == MODE_RETRIEVE) {
                        const int q = meta->s_q[slot];
                        const unsigned long long thr = meta->s_thr[slot];
                        const uint32_t thr_score = (uint32_t)(thr >> 33);
                        unsigned int *ccnt = a.cand_cnt + q;
                        unsigned long long *crow = a.cand_key + (size_t)q * (size_t)a.cap;
                        float4 *acc4 = reinterpret_cast<float4 *>(acc);
                        uint32_t *cnt4 = reinterpret_cast<uint32_t *>(cnt);
                        for (int w = tid; w < D / 4; w += NT) {
                            uint32_t c4 = 0;
                            float4 v;
                            if (HAS_CNT) {
                                c4 = cnt4[w];
                                if (c4 == 0) continue;  // untouched quad
                                v = acc4[w];
                                cnt4[w] = 0;
                            } else {
                                v = acc4[w];
                                if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
                            }
                            acc4[w] = make_float4(0.f, 0.f, 0.f, 0.f);
                            const float av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const uint32_t bits = __float_as_uint(av[c]);
                                if (bits != 0u && bits >= thr_score) {
                                    const unsigned long long key =
                                        make_key(bits, (uint32_t)(doc_base + w * 4 + c), (c4 >> (8 * c)) & 255u);
                                    if (key >= thr) {
                                        const unsigned int pos = atomicAdd(ccnt, 1u);
                                        if (pos < (unsigned)a.cap) crow[pos] = key;
                                    }
                                }
                            }
                        }
                    } else {
                        for (int w = tid; w < D; w += NT) {
                            const long long d = (long long)doc_base + w;
                            const float sc = acc[w];
                            const int c = cnt[w];
                            acc[w] = 0.f;
                            cnt[w] = 0;
                            if (d < a.n_docs) {
                                if (MODE == MODE_SCORES) {
                                    a.out_scores[d] = sc;
                                } else if (MODE == MODE_PROBS) {
                                    a.out_probs[d * a.out_stride] = d_doc_probability(a.params, sc, c, a.doc_len[d], a.avgdl);
                                } else {
                                    // MODE_FUSED: this index is one signal of a log-odds conjunction
                                    // (fusion.py:243-268): acc = [acc +] w * logit(clamp(p)); the last
                                    // signal turns the sum into sigma(scale * acc)
                                    const double p = d_doc_probability(a.params, sc, c, a.doc_len[d], a.avgdl);
                                    a.out_probs[d] = fuse_step(a.out_probs[d], d_logit(p), a.fuse);
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// per-query preparation: sanitised term copy, duplicate flags, threshold seed
// ---------------------------------------------------------------------------------
__global__ void prep_queries_kernel(const int32_t *__restrict__ q_terms, const int64_t *__restrict__ q_off,
                                    int64_t n_q, int64_t term_base, int64_t n_terms_total, int64_t n_vocab,
                                    const float *__restrict__ kth, int32_t *__restrict__ qt_ws,
                                    uint8_t *__restrict__ nocount, int64_t *__restrict__ qo_ws,
                                    unsigned long long *__restrict__ thr,
                                    unsigned int *__restrict__ cand_cnt, unsigned int *__restrict__ n_prev,
                                    int *err, const int64_t *__restrict__ indptr = nullptr,
                                    const int32_t *__restrict__ dense_slot = nullptr,
                                    const longlong2 *__restrict__ tab_row = nullptr,
                                    longlong2 *__restrict__ qt_info = nullptr,
                                    const int32_t *__restrict__ row_slot = nullptr) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    // Offsets are rebased to the batch's first term and clamped into [0, n_terms_total]: whatever
    // q_off holds, no later kernel can index outside the n_terms_total-entry workspaces.  A query
    // whose range was clamped or runs backwards is flagged (the call fails) and treated as empty.
    const int64_t r0 = q_off[q] - term_base, r1 = q_off[q + 1] - term_base;
    const int64_t t0 = r0 < 0 ? 0 : (r0 > n_terms_total ? n_terms_total : r0);
    int64_t t1 = r1 < 0 ? 0 : (r1 > n_terms_total ? n_terms_total : r1);
    if (t0 != r0 || t1 != r1 || t1 < t0) atomicOr(err, 2);
    qo_ws[q] = t0;
    if (q == n_q - 1) qo_ws[n_q] = t1;
    if (t1 < t0) t1 = t0;
    uint32_t best = 0;
    for (int64_t i = t0; i < t1; i++) {
        int32_t t = q_terms[term_base + i];
        if (t < 0 || (int64_t)t >= n_vocab) {
            atomicOr(err, 1);
            t = 0;
        }
        bool dup = false;
        for (int64_t j = t0; j < i && !dup; j++) dup = (q_terms[term_base + j] == t);
        qt_ws[i] = t;
        nocount[i] = dup ? 1 : 0;
        // per query term: posting-list start and dense-row slot, so that the traversal reads them
        // alongside the term id instead of after it
        if (qt_info) {
            // .y: low half = hot (dense) slot, high half = slot of any value row (hot or lookup), -1 = none
            const int hot = dense_slot ? dense_slot[t] : -1;
            const int any = row_slot ? row_slot[t] : hot;
            qt_info[2 * i] = make_longlong2(indptr[t], (long long)(((unsigned long long)(uint32_t)any << 32) | (unsigned long long)(uint32_t)hot));
            qt_info[2 * i + 1] = tab_row[t];
        }
        if (kth) {
            uint32_t b = __float_as_uint(kth[t]);
            best = b > best ? b : best;
        }
    }
    if (thr) {
        thr[q] = (unsigned long long)best << 33;
        cand_cnt[q] = 0;
        n_prev[q] = 0;
    }
}

// The same preparation for other translation units (bb25_fused.cu): sanitised term copy,
// duplicate flags, rebased offsets and the per-term (indptr, dense slot, table row) records of one index.
int prep_queries_launch(const bb25_index *idx, const int32_t *q_terms, const int64_t *q_off, int64_t n_q,
                        int64_t term_base, int64_t n_terms_total, int32_t *qt_ws, uint8_t *nocount, int64_t *qo_ws,
                        longlong2 *qt_info, int *err, cudaStream_t st) {
    prep_queries_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(
        q_terms, q_off, n_q, term_base, n_terms_total, idx->n_vocab, nullptr, qt_ws, nocount, qo_ws, nullptr, nullptr,
        nullptr, err, idx->indptr, idx->dense_vals ? idx->dense_slot : nullptr, idx->tab_row, qt_info,
        idx->lookup_vals ? idx->row_slot : nullptr);
    BB25_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------
// per-query selection: sort the candidate keys (descending), then either tighten the
// threshold (overflow / intermediate group) or write the final top-k.
// ---------------------------------------------------------------------------------
struct SelectArgs {
    const int32_t *q_list;
    // sharded retrieval: scores at a few ranks of the running top-k, published between block groups
    unsigned long long *quant;  // [n_q][n_quant] or NULL; entry n_quant-1 (rank k = the threshold) is filled separately
    int n_quant;
    int qrank[4];
    int n_list;                      // number of queries to process (host-side count) ...
    const unsigned int *n_list_ptr;  // ... or, when set, read on the device
    unsigned int *cand_cnt;
    unsigned int *n_prev;
    unsigned long long *cand_key;
    unsigned long long *thr;
    int cap, k, final_pass;
    unsigned int *n_over;
    int32_t *over_list;
    const int32_t *doc_len;
    double avgdl;
    int64_t n_docs, doc_id_offset;
    bb25_params params;
    int64_t *out_ids;
    float *out_scores;
    double *out_probs;
    unsigned long long *n_cand_total;
    // tf_search != 0: keys carry no matched-term count; recover it for the winners by
    // searching each distinct query term's posting list inside the document's tile
    int tf_search;
    // rescore != 0: keys [n_prev, n) carry order-free sums (block_kernel's relaxed mode); replace
    // them by the exact query-order score and drop the ones that miss the threshold
    int rescore;
    // when at most sort_cap (<= kpad) valid keys are left, one bitonic sort replaces the 8-pass radix
    // select and the separate sort of the winners
    int sort_cap;
    // pre-filter of the re-scoring: a new key below prefilter * (k-th largest key as it stands) cannot be among
    // the best k; 1 - 3e-5 for fp32 order-free sums, 1 - 1.1e-3 when the sums come from fp16-bound rows (rounded up: <= 2^-10)
    float prefilter;
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    BlockTable tab;
    int64_t n_vocab;
    const int32_t *dense_slot;
    const float *dense_vals;
    int64_t dense_stride;
    const int32_t *q_terms;
    const uint8_t *q_nocount;
    const int64_t *q_off;
    int64_t term_base;
};

// scorer.py:592-601 for one (query, document): number of distinct query terms whose
// posting list contains the document.  Head terms answer from their dense value row
// (absent documents hold -0.0f, so presence is one load even when the value is +0.0f);
// other terms by binary search inside the document's 1024-doc block slice.
__device__ inline int count_matched_terms(const SelectArgs &a, int q, uint32_t doc) {
    const long long t0 = a.q_off[q] - a.term_base, t1 = a.q_off[q + 1] - a.term_base;
    const int blk = (int)(doc / (uint32_t)kBlockDocs);
    int c = 0;
    for (long long i = t0; i < t1; i++) {
        if (a.q_nocount[i]) continue;  // duplicate occurrence of an earlier term
        const int t = a.q_terms[i];
        const int slot = a.dense_slot[t];
        if (slot >= 0) {
            c += (__float_as_uint(a.dense_vals[(size_t)slot * (size_t)a.dense_stride + doc]) != 0x80000000u);
            continue;
        }
        const uint2 ent = tab_lookup(a.tab, a.tab.row[t], blk);
        const int len = (int)(ent.y & kBlkLenMask);
        if (len == 0) continue;
        if (slice_find(a.indices + a.indptr[t] + (long long)ent.x, len, doc, doc & ~(uint32_t)(kBlockDocs - 1)) >= 0) c++;
    }
    return c;
}

// bm25s's score of one (query, document): the values of the query's terms at the document,
// added in query order in fp32 (the order block_kernel's relaxed mode does not keep).  Warp
// cooperative: lane i looks term i up -- dense value row, or binary search inside the
// document's block slice -- and the sum runs over the lanes in order.  Absent terms add
// +-0.0f, which leaves a non-negative sum unchanged.
__device__ inline float exact_score_warp(const SelectArgs &a, int q, uint32_t doc, int lane) {
    const long long t0 = a.q_off[q] - a.term_base, t1 = a.q_off[q + 1] - a.term_base;
    const int blk = (int)(doc / (uint32_t)kBlockDocs);
    float s = 0.f;
    for (long long b0 = t0; b0 < t1; b0 += 32) {
        const int nb = (int)min(32ll, t1 - b0);
        float val = 0.f;
        if (lane < nb) {
            const int t = a.q_terms[b0 + lane];
            const int slot = a.dense_slot ? a.dense_slot[t] : -1;
            if (slot >= 0) {
                val = a.dense_vals[(size_t)slot * (size_t)a.dense_stride + doc];
            } else {
                const uint2 ent = tab_lookup(a.tab, a.tab.row[t], blk);
                const int len = (int)(ent.y & kBlkLenMask);
                if (len) {
                    const long long s0 = a.indptr[t] + (long long)ent.x;
                    const int pos = slice_find(a.indices + s0, len, doc, doc & ~(uint32_t)(kBlockDocs - 1));
                    if (pos >= 0) val = a.data[s0 + pos];
                }
            }
        }
        for (int i = 0; i < nb; i++) s = __fadd_rn(s, __shfl_sync(0xFFFFFFFFu, val, i));
    }
    return s;
}

template <int NT>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long *keys, int P, int tid) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (P >> 1); t += NT) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = keys[lo], y = keys[hi];
                if ((x < y) == desc) {
                    keys[lo] = y;
                    keys[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}

// Shared-memory plan of select_kernel: keys u64[cap] | top u64[kpad] | hist u32[256] |
// st u32[4] | flag u8[kpad] | list u16[cap]   (kpad = k rounded up to a power of two; top holds
// max(kpad, sort_cap) keys)
template <int NT>
__device__ __forceinline__ void select_one(const SelectArgs &a, unsigned char *smem, const int q) {
    const int k = a.k;
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *top = keys + a.cap;  // max(kpad, sort_cap) entries
    unsigned int *hist = reinterpret_cast<unsigned int *>(top + max(kpad, a.sort_cap));
    unsigned int *st = hist + 256;
    uint8_t *flag = reinterpret_cast<uint8_t *>(st + 4);
    const int tid = threadIdx.x;
    const unsigned int n_raw = a.cand_cnt[q];
    const int n = (int)min(n_raw, (unsigned)a.cap);
    unsigned long long *row = a.cand_key + (size_t)q * (size_t)a.cap;
    const bool overflow = n_raw > (unsigned)a.cap;

    if (!a.rescore && !overflow && !a.final_pass && n <= k) {
        // nothing to drop and no tighter bound to learn yet
        if (tid == 0) a.n_prev[q] = (unsigned int)n;
        return;
    }
    for (int i = tid; i < n; i += NT) keys[i] = row[i];
    if (tid == 0) st[3] = 0u;
    __syncthreads();
    int n_valid = n;
    if (a.rescore) {
        // the keys added since the last selection carry order-free sums: one warp per key
        // recomputes the exact score; keys that miss the threshold become 0 (= empty)
        const unsigned long long thr64 = a.thr[q];
        const int lane = tid & 31;
        const long long t0 = a.q_off[q] - a.term_base;
        const int m = (int)(a.q_off[q + 1] - a.term_base - t0);
        const int n0 = (int)a.n_prev[q];
        // Keys that cannot be among the best k need no exact score: with T the k-th largest
        // key as it stands (order-free sums are within 4e-6 relative of the exact ones for
        // m <= 32), a new key below T * (1 - 3e-5) is strictly below k documents' exact
        // scores.  Longer queries were summed in query order already (margin unused).
        uint32_t cut_bits = 0u;
        if (n >= k && m <= 32) {
            const unsigned long long ta = block_kth_largest<NT>(keys, n, k, hist, st, tid, 3);  // top 24 bits: T rounded down
            cut_bits = __float_as_uint(__fmul_rn(__uint_as_float(key_score_bits(ta)), a.prefilter));
            __syncthreads();
        }
        uint16_t *list = reinterpret_cast<uint16_t *>(flag + kpad);
        if (tid == 0) st[2] = 0u;
        __syncthreads();
        for (int i = n0 + tid; i < n; i += NT) {
            if (key_score_bits(keys[i]) >= cut_bits) list[atomicAdd(&st[2], 1u)] = (uint16_t)i;
            else keys[i] = 0ull;
        }
        __syncthreads();
        const int n_list = (int)st[2];
        if (m <= 32) {
            // a group of g = 2^ceil(log2 m) lanes per key, lane j of the group looks term j up
            int g = 1;
            while (g < m) g <<= 1;
            const int sub = lane / g, j = lane % g;
            const int groups = (NT / 32) * (32 / g);
            const int gid = (tid >> 5) * (32 / g) + sub;
            int slot = -1;
            long long ip = 0;
            longlong2 trow = make_longlong2(0, -1);
            bool is_dup = false;  // a repeated occurrence of an earlier term does not count twice
            if (j < m) {
                const int t = a.q_terms[t0 + j];
                is_dup = a.q_nocount[t0 + j] != 0;
                slot = a.dense_slot ? a.dense_slot[t] : -1;
                if (slot < 0) {
                    ip = a.indptr[t];
                    trow = a.tab.row[t];
                }
            }
            for (int base = 0; base < n_list; base += groups) {
                const bool act = base + gid < n_list;
                const int i = act ? (int)list[base + gid] : 0;
                float val = 0.f;
                uint32_t id = 0;
                bool present = false;
                if (act) {
                    id = key_local_id(keys[i]);
                    if (j < m) {
                        if (slot >= 0) {
                            val = a.dense_vals[(size_t)slot * (size_t)a.dense_stride + id];
                            present = __float_as_uint(val) != 0x80000000u;
                        } else {
                            const uint2 ent = tab_lookup(a.tab, trow, (int)(id / (uint32_t)kBlockDocs));
                            const int len = (int)(ent.y & kBlkLenMask);
                            if (len) {
                                const long long s0 = ip + (long long)ent.x;
                                const int pos = slice_find(a.indices + s0, len, id, id & ~(uint32_t)(kBlockDocs - 1));
                                if (pos >= 0) {
                                    val = a.data[s0 + pos];
                                    present = true;
                                }
                            }
                        }
                    }
                }
                float sc = 0.f;
                for (int u = 0; u < m; u++) sc = __fadd_rn(sc, __shfl_sync(0xFFFFFFFFu, val, sub * g + u));
                // matched distinct terms (scorer.py:592-601) ride in the key's low bits so that the
                // output stage need not search again; 0 = not recorded (15 or more: search)
                const unsigned pres = __ballot_sync(0xFFFFFFFFu, present && !is_dup);
                const unsigned tfc = (unsigned)__popc(g == 32 ? pres : ((pres >> (sub * g)) & ((1u << g) - 1u)));
                if (act && j == 0) {
                    const uint32_t bits = __float_as_uint(sc);
                    unsigned long long key = make_key(bits, id, tfc < 15u ? tfc : 0u);
                    if (bits == 0u || key < thr64) key = 0ull;
                    keys[i] = key;
                }
            }
        } else {
            for (int c = tid >> 5; c < n_list; c += NT / 32) {
                const int i = (int)list[c];
                const uint32_t id = key_local_id(keys[i]);
                const uint32_t bits = __float_as_uint(exact_score_warp(a, q, id, lane));
                unsigned long long key = make_key(bits, id, 0u);
                if (bits == 0u || key < thr64) key = 0ull;
                if (lane == 0) keys[i] = key;
            }
        }
        __syncthreads();
        unsigned int c = 0;
        for (int i = tid; i < n; i += NT) c += keys[i] != 0ull;
        if (c) atomicAdd(&st[3], c);
        __syncthreads();
        n_valid = (int)st[3];
        if (!overflow && !a.final_pass && n_valid <= k) {
            // keep the exact keys, compacted; no tighter bound to learn yet
            if (tid == 0) st[2] = 0u;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const unsigned long long key = keys[i];
                if (key) row[atomicAdd(&st[2], 1u)] = key;
            }
            if (tid == 0) {
                a.cand_cnt[q] = (unsigned int)n_valid;
                a.n_prev[q] = (unsigned int)n_valid;
            }
            return;
        }
    }
    unsigned long long kth = 0ull;
    bool sorted = false;  // top[0 .. n_valid) holds every valid key in descending order
    if (n_valid >= k && n_valid <= a.sort_cap) {
        int P = 2;
        while (P < n_valid) P <<= 1;
        if (tid == 0) st[2] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const unsigned long long key = keys[i];
            if (key != 0ull) top[atomicAdd(&st[2], 1u)] = key;
        }
        for (int i = n_valid + tid; i < P; i += NT) top[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc<NT>(top, P, tid);
        kth = top[k - 1];
        sorted = true;
    } else if (n_valid >= k) {
        kth = block_kth_largest<NT>(keys, n, k, hist, st, tid);
    }

    if (overflow) {
        // more candidates than the row holds: the k-th best of the stored ones is a
        // valid bound, strictly tighter when the keys are exact; the host redoes this group
        // for this query in query order (exact keys against the full 64-bit threshold), which
        // always converges
        if (tid == 0) {
            const unsigned long long t = kth & ~15ull;
            if (t > a.thr[q]) a.thr[q] = t;
            a.cand_cnt[q] = a.n_prev[q];
            const unsigned int pos = atomicAdd(a.n_over, 1u);
            a.over_list[pos] = q;
        }
        return;
    }
    if (!a.final_pass) {
        // keep the best k (order irrelevant), raise the threshold to the k-th key
        if (sorted) {
            for (int i = tid; i < k; i += NT) row[i] = top[i];
            // "this shard holds qrank[j] documents with at least this score" (score bits only: doc ids are shard-local)
            if (a.quant && tid < a.n_quant - 1) a.quant[(size_t)q * a.n_quant + tid] = top[a.qrank[tid] - 1] & ~((1ull << 33) - 1ull);
        } else {
            if (tid == 0) st[2] = 0u;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const unsigned long long key = keys[i];
                if (key != 0ull && key >= kth) row[atomicAdd(&st[2], 1u)] = key;
            }
        }
        if (tid == 0) {
            a.cand_cnt[q] = (unsigned int)k;
            a.n_prev[q] = (unsigned int)k;
            const unsigned long long t = kth & ~15ull;
            if (t > a.thr[q]) a.thr[q] = t;
        }
        return;
    }
    // final pass: the best min(n, k) keys, sorted
    const int n_pos = min(n_valid, k);
    if (!sorted) {
        if (tid == 0) st[2] = 0u;
        for (int i = tid; i < kpad; i += NT) top[i] = 0ull;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const unsigned long long key = keys[i];
            if (key != 0ull && key >= kth) top[atomicAdd(&st[2], 1u)] = key;
        }
        __syncthreads();
        int P = 2;
        while (P < n_pos) P <<= 1;
        bitonic_sort_desc<NT>(top, P, tid);
    }
    if (tid == 0 && a.n_cand_total) atomicAdd(a.n_cand_total, (unsigned long long)n);
    for (int r = tid; r < n_pos; r += NT) {
        const unsigned long long key = top[r];
        const uint32_t id = key_local_id(key);
        const float sc = __uint_as_float(key_score_bits(key));
        const size_t o = (size_t)q * (size_t)k + r;
        a.out_ids[o] = (int64_t)id + a.doc_id_offset;
        if (a.out_scores) a.out_scores[o] = sc;
        const int tf = (a.tf_search && key_tf(key) == 0u) ? count_matched_terms(a, q, id) : (int)key_tf(key);
        a.out_probs[o] = d_doc_probability(a.params, sc, tf, a.doc_len[id], a.avgdl);
    }
    if (n_pos < k) {
        // fewer than k matching documents: bm25s fills the tail with zero-score
        // documents; canonically the lowest doc ids not already listed.  At most
        // n_pos < k ids are taken, so the fill ids all lie in [0, k).
        for (int i = tid; i < k; i += NT) flag[i] = 0;
        __syncthreads();
        for (int r = tid; r < n_pos; r += NT) {
            const uint32_t id = key_local_id(top[r]);
            if (id < (uint32_t)k) flag[id] = 1;
        }
        __syncthreads();
        if (tid < 32) {
            int r = n_pos;
            for (int base = 0; base < k && r < k; base += 32) {
                const int id = base + tid;
                const bool free_id = id < k && (int64_t)id < a.n_docs && !flag[id];
                const unsigned m = __ballot_sync(0xFFFFFFFFu, free_id);
                const int my = r + __popc(m & ((1u << tid) - 1u));
                if (free_id && my < k) {
                    const size_t o = (size_t)q * (size_t)k + my;
                    a.out_ids[o] = (int64_t)id + a.doc_id_offset;
                    if (a.out_scores) a.out_scores[o] = 0.0f;
                    a.out_probs[o] = 0.0;
                }
                r += __popc(m);
            }
        }
    }
}

// One CTA per query of the list, grid-strided: the list length may live on the device (repair
// rounds are launched without knowing how many queries overflowed), in which case the grid is a
// small fixed size and CTAs without work leave at once.
template <int NT>
__global__ void __launch_bounds__(NT) select_kernel(const __grid_constant__ SelectArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int n_list = a.n_list_ptr ? *a.n_list_ptr : (unsigned int)a.n_list;
    for (unsigned int b = blockIdx.x; b < n_list; b += gridDim.x) {
        select_one<NT>(a, smem, a.q_list ? a.q_list[b] : (int)b);
        __syncthreads();
    }
}

// =================================================================================
// Warp-private traversal (retrieve mode).  A BLOCK = 1024 consecutive documents whose
// fp32 accumulators (4 KB) belong to ONE warp; a warp pulls (block, 8-query chunk)
// items from a global counter and, for every query of the chunk,
//   1. reads the block-table entries of the query's terms (offset, length, block max),
//   2. forms the block's score upper bound in query order and SKIPS the unit when the
//      bound is below the query's threshold (block-max pruning; exact because fp32
//      addition of non-negative values is monotone), or when no term has a posting
//      in the block,
//   3. otherwise evaluates the unit, in one of two instantiations:
//      EXACT = false (first evaluation of every block group).  Sums are formed in ANY
//        order: terms without a dense value row are scattered into the accumulators, then
//        one pass adds the dense rows of the frequent terms to them IN REGISTERS (FADD2),
//        tests the result against 0.99999 x threshold and drops it -- the frequent terms
//        never touch shared memory.  Any order of <= 32 non-negative fp32 additions is
//        within 4e-6 relative of the query-order sum, so no qualifying document is missed;
//        select_kernel then replaces each emitted key's sum by the exact query-order score
//        (bm25s's arithmetic) before anything is ranked or a threshold is raised.
//      EXACT = true (threshold repair after a candidate-row overflow, and indexes without
//        dense rows).  Posting slices / dense rows are added in query order (a __syncwarp
//        between terms is all the ordering needs: one warp owns every document of the
//        block), the 1024 accumulators are scanned, keys >= the 64-bit threshold emitted.
// No block-level barrier anywhere.  Items are block-major, so the whole chip works on a
// handful of adjacent blocks whose index slices and value rows sit in L2 / L1.
// =================================================================================
#ifndef BB25_QC
#define BB25_QC 8
#endif
#ifndef BB25_ENT_PREFETCH
#define BB25_ENT_PREFETCH 0
#endif
constexpr int QC = BB25_QC;  // queries per warp work item
#ifndef BB25_BK_WARPS
#define BB25_BK_WARPS 8
#endif
constexpr int BK_WARPS = BB25_BK_WARPS;  // warps per CTA (4 KB of accumulators per warp)

struct BlockArgs {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    BlockTable tab;
    int64_t n_vocab;
    const int32_t *q_terms;  // sanitised copy, indexed by absolute position - term_base
    const longlong2 *qt_info;  // per query term, two records: (indptr[t], dense slot or -1), block-table row of t
    const int64_t *q_off;
    int64_t term_base;
    const int32_t *q_list;
    int n_q;
    const unsigned int *n_q_ptr;  // when set, the number of queries is read on the device
    int blk_begin, blk_end;
    const unsigned long long *thr;
    unsigned int *cand_cnt;
    unsigned long long *cand_key;
    int cap;
    int prune;  // 0 exhaustive, 1 block-max skip, 2 + skip of frequent-term-only documents by their bound
    const int32_t *dense_slot;
    const float *dense_vals;
    const __half *dense_h;  // fp16 upper-bound copy of the dense rows (order-free pass only) or NULL
    int64_t dense_stride;
    const float *lookup_vals;  // value rows of the mid-frequency terms (bb25_index::lookup_vals) or NULL
    int n_hot;                 // row slots below this are rows of dense_vals, the others rows of lookup_vals
    int sparse_mode;           // pruning level >= 2: units evaluated through their essential postings (group_units)
    uint8_t *unit_mask;        // sparse mode: per work item (block, QC-query chunk) the query slots that need the pass,
                               // written by group_kernel and read by block_kernel
    unsigned long long *work_counter;
    unsigned long long *stats;  // [0] (block, query) units handed out, [1] units pruned by the block-max bound, [2] units under
                                // the level-2 restriction, [3] units evaluated through their essential postings
};

// per-warp shared memory: 1024 fp32 accumulators

// postings [s, s+len) of one term, len <= 1024, into the warp's block accumulators
template <bool FRESH = false>
__device__ __forceinline__ void scatter_warp(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                             long long s, int len, float *acc, int doc_base, int lane) {
    int head = (int)((32 - (s & 31)) & 31);  // elements before the first 128-byte boundary
    if (head > len) head = len;
    if (lane < head) {
        const long long j = s + lane;
        const int o = ld_nc_s32(indices + j) - doc_base;
        acc_add<FRESH>(acc, o, ld_nc_f32(data + j));
    }
    const int n = len - head;
    const int32_t *ip = indices + s + head;
    const float *dp = data + s + head;
    int j = lane;
    for (; j + 96 < n; j += 128) {
        int d[4];
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            d[u] = ld_nc_s32(ip + j + 32 * u);
            v[u] = ld_nc_f32(dp + j + 32 * u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc_add<FRESH>(acc, d[u] - doc_base, v[u]);
    }
    for (; j < n; j += 32) {
        const int o = ld_nc_s32(ip + j) - doc_base;
        acc_add<FRESH>(acc, o, ld_nc_f32(dp + j));
    }
}

// A head term whose block slice is (nearly) full is cheaper to add from its dense value
// row: 128-bit row loads and 128-bit shared read-modify-writes, 4 documents per lane per
// step, instead of walking ~1000 (doc id, value) postings.  Absent documents hold -0.0f,
// and x + (-0.0f) == x, so the sums stay bit-exact.
#ifndef BB25_DENSE_MIN
#define BB25_DENSE_MIN 320
#endif
constexpr int kDenseAddMinLen = BB25_DENSE_MIN;
template <bool FRESH = false>
__device__ __forceinline__ void dense_add_warp(const float *__restrict__ row, float4 *acc4, int lane) {
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
#pragma unroll 2
    for (int i = 0; i < kBlockDocs / 128; i++) {
        const int w = i * 32 + lane;
        const float4 r = ld_nc_f4(r4 + w);
        float4 v = FRESH ? make_float4(0.f, 0.f, 0.f, 0.f) : acc4[w];  // FRESH: accumulators are all zero
        v.x = __fadd_rn(v.x, r.x);
        v.y = __fadd_rn(v.y, r.y);
        v.z = __fadd_rn(v.z, r.z);
        v.w = __fadd_rn(v.w, r.w);
        acc4[w] = v;
    }
}

struct TermEnt {
    long long start;
    int len;
    float bmax;
    int dslot;  // dense value row of the term (-1: none, or no posting in this block)
};
// the same with the slot of ANY value row (hot or lookup) and the packed table word, for group_units
struct TermEntX {
    long long start;
    int len;
    int rslot;
    uint32_t pk;  // block maximum (upper 21 bits, rounded up) | slice length
};
template <bool SPARSE_TAB>
__device__ __forceinline__ TermEntX load_term_entry_x(const BlockArgs &a, int blk, long long pos, bool active) {
    TermEntX e;
    e.start = 0;
    e.len = 0;
    e.rslot = -1;
    e.pk = 0u;
    if (active) {
        const longlong2 info = a.qt_info[2 * pos];       // (indptr[t], hot slot | any-row slot << 32)
        const longlong2 trow = a.qt_info[2 * pos + 1];   // the term's block-table row
        const uint2 ent = SPARSE_TAB ? tab_lookup(a.tab, trow, blk) : a.tab.ent[trow.x + blk];
        e.pk = ent.y;
        e.len = (int)(ent.y & kBlkLenMask);
        e.start = info.x + (long long)ent.x;
        e.rslot = e.len > 0 ? (int)(info.y >> 32) : -1;
    }
    return e;
}

template <bool SPARSE_TAB>
__device__ __forceinline__ TermEnt load_term_entry(const BlockArgs &a, int blk, long long pos, bool active) {
    TermEnt e;
    e.start = 0;
    e.len = 0;
    e.bmax = 0.f;
    e.dslot = -1;
    if (active) {
        const longlong2 info = a.qt_info[2 * pos];       // (indptr[t], dense slot)
        const longlong2 trow = a.qt_info[2 * pos + 1];   // the term's block-table row
        // SPARSE_TAB = false: the index keeps a dense row for every term (no bitmap branch)
        const uint2 ent = SPARSE_TAB ? tab_lookup(a.tab, trow, blk) : a.tab.ent[trow.x + blk];
        e.len = (int)(ent.y & kBlkLenMask);
        e.bmax = __uint_as_float(ent.y & ~kBlkLenMask);
        e.start = info.x + (long long)ent.x;
        e.dslot = e.len > 0 ? (int)info.y : -1;
    }
    return e;
}

// emit one accumulator if it can enter the query's top-k
__device__ __forceinline__ void emit_if_candidate(float v, uint32_t local_id, uint32_t thr_score,
                                                  unsigned long long thr, unsigned int *ccnt,
                                                  unsigned long long *crow, int cap) {
    const uint32_t bits = __float_as_uint(v);
    if (bits != 0u && bits >= thr_score) {
        const unsigned long long key = make_key(bits, local_id, 0u);
        if (key >= thr) {
            const unsigned int pos = atomicAdd(ccnt, 1u);
            if (pos < (unsigned)cap) crow[pos] = key;
        }
    }
}

#ifndef BB25_PASS_CHUNKS
#define BB25_PASS_CHUNKS 2
#endif
#ifndef BB25_PASS_UNROLL
#define BB25_PASS_UNROLL 1
#endif
constexpr int kPassUnroll = BB25_PASS_UNROLL;
constexpr int kPassChunks = BB25_PASS_CHUNKS;  // 128-document chunks in flight per round of the order-free pass

// rare path of the order-free pass, kept out of line so it does not cost registers: keys carry
// the any-order sum; select_kernel replaces it by the exact score before anything is ranked
__device__ __noinline__ void emit_quad_relaxed(float4 v, uint32_t first_id, uint32_t thr_rel, unsigned int *cand_cnt,
                                               unsigned long long *cand_key, int q, int cap) {
    unsigned int *ccnt = cand_cnt + q;
    unsigned long long *crow = cand_key + (size_t)q * (size_t)cap;
    const float av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t bits = __float_as_uint(av[c]);
        if (bits != 0u && bits >= thr_rel) {
            const unsigned int pos = atomicAdd(ccnt, 1u);
            if (pos < (unsigned)cap) crow[pos] = make_key(bits, first_id + c, 0u);
        }
    }
}

// One quad (4 consecutive documents) of a D term's row.  HALF: the row of fp16 UPPER BOUNDS (values rounded
// up at index creation, absent = 0) -- half the bytes and L1 wavefronts per quad; the sums formed from it are
// upper bounds of the fp32 sums, which is all the order-free pass needs (every candidate it emits is
// re-scored exactly).
template <bool HALF>
__device__ __forceinline__ float4 ld_row_quad(const unsigned char *row, int quad) {
    if (!HALF) return ld_row_f4(reinterpret_cast<const float4 *>(row) + quad);
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(reinterpret_cast<const uint2 *>(row) + quad));
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

struct PassArgs {
    float4 *acc4;          // the warp's accumulators, already offset by the lane
    const unsigned char *row_a;  // first / second D term's value row at (block, lane): fp32 or fp16-bound quads
    const unsigned char *row_b;
    unsigned rest;         // third and later D terms (query positions)
    int dslot;             // lane i: dense slot of term i
    uint32_t thr_rel;      // relaxed threshold, >= 1
    uint32_t first_id;     // local id of the lane's first document in chunk 0
    int q;
};

// The order-free pass over the block: per 128-document chunk, lane L owns documents 4L..4L+3.
//   ND     D terms whose rows are added from registers (2 = two or more, the others via `rest`)
//   HAS_S  the accumulators hold S-term sums: read them and leave zeros behind
//   PRED   only quads with an S contribution are completed (D-only documents cannot qualify)
template <int ND, bool HAS_S, bool PRED, bool HALF>
__device__ __forceinline__ void order_free_pass(const BlockArgs &a, const PassArgs &pa, const unsigned char *dbase) {
    const size_t row_bytes = (size_t)a.dense_stride * (HALF ? 2 : 4);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll kPassUnroll
    for (int h = 0; h < kBlockDocs / (128 * kPassChunks); h++) {
        const int w0 = h * (32 * kPassChunks);
        float4 v[kPassChunks], ra[kPassChunks], rb[kPassChunks];
#pragma unroll
        for (int j = 0; j < kPassChunks; j++) {
            const int w = w0 + j * 32;
            v[j] = HAS_S ? pa.acc4[w] : zero4;
            if (PRED) {
                const bool nz = fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)) > 0.f;
                ra[j] = zero4;
                rb[j] = zero4;
                if (nz) {
                    if (ND >= 1) ra[j] = ld_row_quad<HALF>(pa.row_a, w);
                    if (ND >= 2) rb[j] = ld_row_quad<HALF>(pa.row_b, w);
                }
            } else {
                if (ND >= 1) ra[j] = ld_row_quad<HALF>(pa.row_a, w);
                if (ND >= 2) rb[j] = ld_row_quad<HALF>(pa.row_b, w);
            }
            if (HAS_S) pa.acc4[w] = zero4;
        }
#pragma unroll
        for (int j = 0; j < kPassChunks; j++) {
            if (ND >= 1) add_f4(v[j], ra[j]);
            if (ND >= 2) add_f4(v[j], rb[j]);
        }
        if (ND >= 2) {
            for (unsigned mm = pa.rest; mm; mm &= mm - 1) {  // third and later D terms
                const int slot = __shfl_sync(0xFFFFFFFFu, pa.dslot, __ffs(mm) - 1);
                const unsigned char *rp = dbase + (size_t)slot * row_bytes;
#pragma unroll
                for (int j = 0; j < kPassChunks; j++) {
                    if (!PRED || fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)) > 0.f)
                        add_f4(v[j], ld_row_quad<HALF>(rp, w0 + j * 32));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kPassChunks; j++) {
            const float mx = fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w));
            if (__float_as_uint(mx) >= pa.thr_rel)
                emit_quad_relaxed(v[j], pa.first_id + (uint32_t)((w0 + j * 32) * 4), pa.thr_rel, a.cand_cnt, a.cand_key, pa.q, a.cap);
        }
    }
}

// ---------------------------------------------------------------------------------
// Essential-posting evaluation of units (pruning level >= 2, order-free kernel): MaxScore -- the partition
// behind wand_upper_bound (probability.py:205-236) and BlockMaxIndex (scorer.py:33-142) -- at the
// granularity of one 1024-document block, FOUR queries of the chunk at a time (8 lanes per query).
//
// A unit's terms are split: NON-ESSENTIAL = the terms that have a value row (hot or lookup) and at least L
// postings in this block, for the smallest L of a fixed ladder whose summed block maxima (query order,
// exactly the block-max test restricted to the subset) stay below the query's threshold.  A document that
// matches non-essential terms only cannot reach the threshold, so every qualifying document has a posting
// in an ESSENTIAL slice (row-less terms; row terms with fewer than L postings here).  When the unit's
// essential slices hold at most 32 postings it is evaluated document-at-a-time in one round: lane =
// essential posting (doc id and value straight from the slice), postings of one document are combined
// with a warp match, every non-essential term's value is ONE 4-byte load from its row.  No accumulators,
// no scatter, no pass over 1024 documents.  The candidates carry any-order sums; select_kernel re-scores
// them exactly, as it does for the order-free pass.  The ladder's last step is the full term set = the
// block-max skip test.  Returns the query slots that need the pass (or cannot be handled here).
// ---------------------------------------------------------------------------------
constexpr int kSparseMax = 32;

__device__ __forceinline__ float row_value(const BlockArgs &a, int rslot, uint32_t doc) {
    return rslot < a.n_hot ? a.dense_vals[(size_t)rslot * (size_t)a.dense_stride + doc]
                           : a.lookup_vals[(size_t)(rslot - a.n_hot) * (size_t)a.dense_stride + doc];  // absent: -0.0f
}

template <bool SPARSE_TAB>
__device__ __forceinline__ unsigned group_units(const BlockArgs &a, const uint4 *sdesc, int nslots, int blk, int doc_base,
                                                int lane, unsigned int *scnt) {
    const int qs = lane >> 3, tl = lane & 7, sh = qs * 8;
    // ladder: row terms with >= 1, 2, 3, 5, 9, 17, 33 postings here; step 7: every term (the block-max test)
    const int mycut = (tl == 0 || tl == 7) ? 1 : 1 + (1 << (tl - 1));
    unsigned serial = 0u;
    for (int s0 = 0; s0 < nslots; s0 += 4) {
        const int sl = s0 + qs;
        const bool qact = sl < nslots;
        const uint4 d = qact ? sdesc[sl] : make_uint4(0u, 0u, 0u, 0u);  // q, term count, first term position, threshold score bits
        const int m = (int)d.y;
        const uint32_t thr_score = d.w;
        const bool ser = qact && (m > 8 || thr_score == 0u);
        const bool grp = qact && !ser && m > 0;
        const TermEntX e = load_term_entry_x<SPARSE_TAB>(a, blk, (long long)(int)d.z + tl, grp && tl < m);
        const unsigned pb = __ballot_sync(0xFFFFFFFFu, e.len > 0);
        const unsigned rb = __ballot_sync(0xFFFFFFFFu, e.rslot >= 0);
        const unsigned rm = tl == 7 ? 0xFFu : (rb >> sh) & 0xFFu;
        const int mmax = min(8, (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(grp ? m : 0)));
        float ub = 0.f;
        for (int t = 0; t < mmax; t++) {
            const uint32_t w = __shfl_sync(0xFFFFFFFFu, e.pk, (lane & 24) | t);
            if ((int)(w & kBlkLenMask) >= mycut && ((rm >> t) & 1u)) ub = __fadd_rn(ub, __uint_as_float(w & ~kBlkLenMask));
        }
        const bool below = grp && __float_as_uint(ub) < thr_score;
        const unsigned bm = (__ballot_sync(0xFFFFFFFFu, below) >> sh) & 0xFFu;
        int L = 1;
        if (bm & 0x7Fu) {
            const int ci = __ffs(bm & 0x7Fu) - 1;
            L = ci == 0 ? 1 : 1 + (1 << (ci - 1));
        }
        const bool ess = e.len > 0 && (e.rslot < 0 || e.len < L);
        int n_e = ess ? e.len : 0;
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 1);
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 2);
        n_e += __shfl_xor_sync(0xFFFFFFFFu, n_e, 4);
        // 0 nothing to do, 1 skipped by the bound, 2 essential postings, 3 the pass
        int cls;
        if (!qact || (!ser && !((pb >> sh) & 0xFFu))) cls = 0;
        else if (ser) cls = 3;
        else if (bm & 0x80u) cls = 1;
        else if (!(bm & 0x7Fu) || n_e > kSparseMax) cls = 3;
        else cls = n_e > 0 ? 2 : 1;
        const unsigned n_skip = (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, cls == 1 && tl == 0));
        const unsigned sb3 = __ballot_sync(0xFFFFFFFFu, cls == 3 && tl == 0);
#pragma unroll
        for (int gq = 0; gq < 4; gq++)
            if ((sb3 >> (gq * 8)) & 1u) serial |= 1u << (s0 + gq);
        unsigned pend = __ballot_sync(0xFFFFFFFFu, cls == 2 && tl == 0);
        if (lane == 0) {
            scnt[0] += n_skip;
            scnt[2] += (unsigned)__popc(pend);
        }
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
            // ---- one round: lane = essential posting of one of the batch's queries ----
            int o = -1, qid = 0, fill = 0;
            float v = 0.f;
            for (unsigned mm = __ballot_sync(0xFFFFFFFFu, cls == 2 && ess) & batch; mm; mm &= mm - 1) {
                const int src = __ffs(mm) - 1;
                const int len = __shfl_sync(0xFFFFFFFFu, e.len, src);
                const long long s = shfl_ll(e.start, src);
                const int r = lane - fill;
                if (r >= 0 && r < len) {
                    o = ld_nc_s32(a.indices + s + r) - doc_base;
                    v = ld_nc_f32(a.data + s + r);
                    qid = src >> 3;
                }
                fill += len;
            }
            const bool valid = o >= 0;
            // postings of one (query, document) -- several essential slices may hold it -- are summed by the first lane
            const unsigned g = __match_any_sync(0xFFFFFFFFu, valid ? ((qid << 10) | o) : (0x10000 | lane));
            const int rounds = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)__popc(g));
            float sum = 0.f;
            unsigned rem = g;
            for (int r = 0; r < rounds; r++) {
                const float vv = __shfl_sync(0xFFFFFFFFu, v, rem ? __ffs(rem) - 1 : lane);
                if (rem) {
                    sum = __fadd_rn(sum, vv);
                    rem &= rem - 1;
                }
            }
            const bool leader = valid && (__ffs(g) - 1 == lane);
            const uint32_t doc = (uint32_t)(doc_base + (valid ? o : 0));
            const int ql = qid << 3;  // first lane of the posting's query
            const int Lq = __shfl_sync(0xFFFFFFFFu, L, ql);
            const uint32_t thrq = __shfl_sync(0xFFFFFFFFu, thr_score, ql);
            const int qq = __shfl_sync(0xFFFFFFFFu, (int)d.x, ql);
            for (int t = 0; t < mmax; t++) {
                const int rs = __shfl_sync(0xFFFFFFFFu, e.rslot, ql | t), ln = __shfl_sync(0xFFFFFFFFu, e.len, ql | t);
                if (leader && rs >= 0 && ln >= Lq) sum = __fadd_rn(sum, row_value(a, rs, doc));
            }
            if (leader) {
                const uint32_t thr_rel = __float_as_uint(__fmul_rn(__uint_as_float(thrq), 0.99999f));
                const uint32_t bits = __float_as_uint(sum);
                if (sum > 0.f && bits >= thr_rel) {
                    const unsigned int pos = atomicAdd(a.cand_cnt + qq, 1u);
                    if (pos < (unsigned)a.cap) a.cand_key[(size_t)qq * (size_t)a.cap + pos] = make_key(bits, doc, 0u);
                }
            }
        }
    }
    return serial;
}

// Resident CTAs per SM (= register cap) of the order-free kernel, measured on config 2: the exhaustive
// pass likes 5 x 8 warps at 48 registers (61.0 vs 62.5 ms per step), the pruned passes, whose units are
// fewer and heavier, 4 x 8 warps at 64 registers (39.7 vs 40.7 ms).  The query-order kernel runs 6.
#ifndef BB25_BLOCK_CTAS
#define BB25_BLOCK_CTAS 5
#endif
#ifndef BB25_BLOCK_CTAS_PRUNED
#define BB25_BLOCK_CTAS_PRUNED 4
#endif
template <int WARPS, bool EXACT, bool SPARSE_TAB, int CTAS, bool HALF, bool GROUP = false>
__global__ void __launch_bounds__(WARPS * 32, CTAS) block_kernel(const __grid_constant__ BlockArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *acc = reinterpret_cast<float *>(smem + (size_t)warp * (kBlockDocs * 4));
    float4 *acc4 = reinterpret_cast<float4 *>(acc);
    // per warp: QC query descriptors (q, term count, first term position, threshold score bits) and one
    // slot of counters ([0] units pruned by the block-max bound, [1] units under the level-2 restriction)
    uint4 *sdesc = reinterpret_cast<uint4 *>(smem + (size_t)WARPS * (kBlockDocs * 4)) + warp * (QC + 1);
    unsigned int *scnt = reinterpret_cast<unsigned int *>(sdesc + QC);
    for (int i = lane; i < kBlockDocs / 4; i += 32) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 4) scnt[lane] = 0u;
    __syncwarp();

    const int n_q = a.n_q_ptr ? (int)*a.n_q_ptr : a.n_q;
    const int n_chunks = (n_q + QC - 1) / QC;
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    if (a.stats && blockIdx.x == 0 && threadIdx.x == 0 && !(GROUP && !EXACT && a.unit_mask))
        atomicAdd(&a.stats[0], (unsigned long long)(a.blk_end - a.blk_begin) * (unsigned long long)n_q);

    for (;;) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(a.work_counter, 1ull);
        item = shfl_ll(item, 0);
        if (item >= n_items) break;
        // pruning level >= 2: group_kernel has already skipped / evaluated most units through their essential
        // postings; the query slots left for the pass are in the unit mask
        unsigned serial = 0xFFFFFFFFu;
        if (GROUP && !EXACT && a.unit_mask) {
            serial = (unsigned)a.unit_mask[item];
            if (!serial) continue;
        }
        const int blk = a.blk_begin + (int)(item / n_chunks);
        const int slot0 = (int)(item % n_chunks) * QC;
        const int nslots = min(QC, n_q - slot0);
        const int doc_base = blk * kBlockDocs;

        // the descriptions of the chunk's queries are fetched by lanes 0..nslots-1 in one round and
        // parked in shared memory (registers are what limits the kernel's occupancy); the full 64-bit
        // threshold is re-read by the query-order path, which alone needs it
        __syncwarp();
        if (lane < nslots) {
            const int my_q = a.q_list ? a.q_list[slot0 + lane] : slot0 + lane;
            const long long t0 = a.q_off[my_q];
            const int my_m = (int)max(0ll, (long long)a.q_off[my_q + 1] - t0);
            sdesc[lane] = make_uint4((unsigned)my_q, (unsigned)my_m, (unsigned)(int)(t0 - a.term_base),
                                     (uint32_t)(a.thr[my_q] >> 33));
        }
        __syncwarp();

#if BB25_ENT_PREFETCH
        // the next query's table entries are fetched while this one is evaluated: two dependent loads
        // (term record, block-table entry) leave the per-unit latency chain
        TermEnt e_pf = load_term_entry<SPARSE_TAB>(a, blk, (int)sdesc[0].z + lane, lane < (int)sdesc[0].y && sdesc[0].y <= 32u);
#endif
        for (int sidx = 0; sidx < nslots; sidx++) {
#if BB25_ENT_PREFETCH
            const TermEnt e_cur = e_pf;
            if (sidx + 1 < nslots) {
                const uint4 dn = sdesc[sidx + 1];
                e_pf = load_term_entry<SPARSE_TAB>(a, blk, (int)dn.z + lane, lane < (int)dn.y && dn.y <= 32u);
            }
#endif
            if (GROUP && !((serial >> sidx) & 1u)) continue;
            const uint4 desc = sdesc[sidx];
            const int m = (int)desc.y;
            if (m == 0) continue;
            const int q = (int)desc.x;
            const int t0 = (int)desc.z;
            const uint32_t thr_score = desc.w;

            if (m <= 32) {
#if BB25_ENT_PREFETCH
                const TermEnt e = e_cur;
#else
                const TermEnt e = load_term_entry<SPARSE_TAB>(a, blk, t0 + lane, lane < m);
#endif
                if (__ballot_sync(0xFFFFFFFFu, e.len > 0) == 0u) continue;  // no posting of any term in this block
                if (a.prune) {
                    float ub = 0.f;
                    for (int i = 0; i < m; i++) ub = __fadd_rn(ub, __shfl_sync(0xFFFFFFFFu, e.bmax, i));
                    if (__float_as_uint(ub) < thr_score) {
                        if (lane == 0) scnt[0]++;
                        continue;
                    }
                }
                const int dslot = e.dslot;
                if (!EXACT) {
                    // ---- order-free evaluation of the unit ---------------------------------------
                    // The candidates this kernel emits are re-scored exactly (query order) by
                    // select_kernel, so here a document's sum may be formed in ANY order as long as
                    // no document whose exact score reaches the threshold is missed: m <= 32 fp32
                    // additions of non-negative values differ between orders by < 2*31*2^-24
                    // relative, hence "sum >= 0.99999 * threshold" keeps them all.  That freedom
                    // lets the frequent terms (D, with a dense value row) stay out of shared
                    // memory: the other terms (S) are scattered into the warp's accumulators, and
                    // one pass adds the D rows to them in registers, tests and drops the result.
                    const unsigned present = __ballot_sync(0xFFFFFFFFu, e.len > 0);
                    const unsigned dmask = __ballot_sync(0xFFFFFFFFu, dslot >= 0);
                    const unsigned smask = present & ~dmask;
                    bool first = true;
                    for (unsigned mm = smask; mm; mm &= mm - 1) {
                        const int i = __ffs(mm) - 1;
                        const int len = __shfl_sync(0xFFFFFFFFu, e.len, i);
                        const long long s = shfl_ll(e.start, i);
                        if (first) scatter_block<true>(a.data, a.indices, s, len, acc, doc_base, lane);
                        else scatter_block<false>(a.data, a.indices, s, len, acc, doc_base, lane);
                        first = false;
                        __syncwarp();
                    }
                    // documents matching D terms only cannot qualify when the D terms' block maxima,
                    // summed in query order, stay below the threshold (pruning level >= 2): then
                    // only quads holding an S contribution are completed
                    bool dense_all = dmask != 0u;
                    if (dense_all && a.prune >= 2 && thr_score != 0u) {
                        float dub = 0.f;
                        for (unsigned mm = dmask; mm; mm &= mm - 1)
                            dub = __fadd_rn(dub, __shfl_sync(0xFFFFFFFFu, e.bmax, __ffs(mm) - 1));
                        if (__float_as_uint(dub) < thr_score) {
                            dense_all = false;
                            if (lane == 0) scnt[1]++;
                        }
                    }
                    if (!smask && !dense_all) continue;
                    const uint32_t thr_rel =
                        thr_score ? __float_as_uint(__fmul_rn(__uint_as_float(thr_score), 0.99999f)) : 1u;
                    const int n_d = __popc(dmask);
                    PassArgs pa;
                    pa.acc4 = acc4 + lane;
                    pa.thr_rel = thr_rel;
                    pa.first_id = (uint32_t)(doc_base + lane * 4);
                    pa.q = q;
                    pa.rest = 0u;
                    pa.dslot = dslot;
                    const size_t row_bytes = (size_t)a.dense_stride * (HALF ? 2 : 4);
                    const unsigned char *dbase = HALF ? reinterpret_cast<const unsigned char *>(a.dense_h + doc_base + lane * 4)
                                                      : reinterpret_cast<const unsigned char *>(a.dense_vals + doc_base + lane * 4);
                    if (n_d >= 1) pa.row_a = dbase + (size_t)__shfl_sync(0xFFFFFFFFu, dslot, __ffs(dmask) - 1) * row_bytes;
                    if (n_d >= 2) {
                        const unsigned d2 = dmask & (dmask - 1);
                        pa.row_b = dbase + (size_t)__shfl_sync(0xFFFFFFFFu, dslot, __ffs(d2) - 1) * row_bytes;
                        pa.rest = d2 & (d2 - 1);
                    }
                    if (!smask) {
                        if (n_d == 1) order_free_pass<1, false, false, HALF>(a, pa, dbase);
                        else order_free_pass<2, false, false, HALF>(a, pa, dbase);
                    } else if (n_d == 0) {
                        order_free_pass<0, true, false, HALF>(a, pa, dbase);
                    } else if (dense_all) {
                        if (n_d == 1) order_free_pass<1, true, false, HALF>(a, pa, dbase);
                        else order_free_pass<2, true, false, HALF>(a, pa, dbase);
                    } else {
                        if (n_d == 1) order_free_pass<1, true, true, HALF>(a, pa, dbase);
                        else order_free_pass<2, true, true, HALF>(a, pa, dbase);
                    }
                    __syncwarp();
                    continue;
                }
                bool fresh = true;  // accumulators all zero until the first term with postings here
                for (int i = 0; EXACT && i < m; i++) {
                    const int len = __shfl_sync(0xFFFFFFFFu, e.len, i);
                    const long long s = shfl_ll(e.start, i);
                    const int slot = __shfl_sync(0xFFFFFFFFu, dslot, i);
                    if (!len) continue;
                    if (slot >= 0 && len >= kDenseAddMinLen) {
                        const float *row_v = a.dense_vals + (size_t)slot * (size_t)a.dense_stride + doc_base;
                        if (fresh) dense_add_warp<true>(row_v, acc4, lane);
                        else dense_add_warp<false>(row_v, acc4, lane);
                    } else {
                        if (fresh) scatter_warp<true>(a.data, a.indices, s, len, acc, doc_base, lane);
                        else scatter_warp<false>(a.data, a.indices, s, len, acc, doc_base, lane);
                    }
                    fresh = false;
                    __syncwarp();
                }
            } else {
                // long query: bound first (one pass over the entries), then the adds
                float ub = 0.f;
                unsigned any = 0u;
                for (int b0 = 0; b0 < m; b0 += 32) {
                    const int nb = min(32, m - b0);
                    const TermEnt e = load_term_entry<SPARSE_TAB>(a, blk, t0 + b0 + lane, lane < nb);
                    any |= __ballot_sync(0xFFFFFFFFu, e.len > 0);
                    for (int i = 0; i < nb; i++) ub = __fadd_rn(ub, __shfl_sync(0xFFFFFFFFu, e.bmax, i));
                }
                if (any == 0u) continue;
                if (a.prune && __float_as_uint(ub) < thr_score) {
                    if (lane == 0) scnt[0]++;
                    continue;
                }
                for (int b0 = 0; b0 < m; b0 += 32) {
                    const int nb = min(32, m - b0);
                    const TermEnt e = load_term_entry<SPARSE_TAB>(a, blk, t0 + b0 + lane, lane < nb);
                    for (int i = 0; i < nb; i++) {
                        const int len = __shfl_sync(0xFFFFFFFFu, e.len, i);
                        const long long s = shfl_ll(e.start, i);
                        if (len) scatter_warp(a.data, a.indices, s, len, acc, doc_base, lane);
                        __syncwarp();
                    }
                }
                if (!EXACT) {
                    // every term went through the accumulators (any order is fine here too)
                    PassArgs pa;
                    pa.acc4 = acc4 + lane;
                    pa.thr_rel = thr_score ? __float_as_uint(__fmul_rn(__uint_as_float(thr_score), 0.99999f)) : 1u;
                    pa.first_id = (uint32_t)(doc_base + lane * 4);
                    pa.q = q;
                    pa.rest = 0u;
                    pa.dslot = -1;
                    order_free_pass<0, true, false, HALF>(a, pa, nullptr);
                    __syncwarp();
                    continue;
                }
            }
            if (!EXACT) continue;

            // fused epilogue over the warp's 1024 accumulators
            const unsigned long long thr = a.thr[q];
            unsigned int *ccnt = a.cand_cnt + q;
            unsigned long long *crow = a.cand_key + (size_t)q * (size_t)a.cap;
            for (int w = lane; w < kBlockDocs / 4; w += 32) {
                const float4 v = acc4[w];
                // sums are >= +0: the quad's maximum decides both "untouched" and "can any qualify"
                const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                if (mx == 0.f) continue;
                acc4[w] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (__float_as_uint(mx) < thr_score) continue;
                const float av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t bits = __float_as_uint(av[c]);
                    if (bits != 0u && bits >= thr_score) {
                        const unsigned long long key = make_key(bits, (uint32_t)(doc_base + w * 4 + c), 0u);
                        if (key >= thr) {
                            const unsigned int pos = atomicAdd(ccnt, 1u);
                            if (pos < (unsigned)a.cap) crow[pos] = key;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && a.stats) {
        atomicAdd(&a.stats[1], (unsigned long long)scnt[0]);
        atomicAdd(&a.stats[2], (unsigned long long)scnt[1]);
    }
}

// The light half of the traversal at pruning level >= 2: no accumulators, few registers, more resident warps than
// block_kernel -- a unit here is a short chain of dependent loads (query descriptor -> term records -> table entries ->
// essential postings -> value rows), and warps are what hides it.  Per work item (block, QC-query chunk) it runs
// group_units and leaves the query slots that need the pass in the unit mask.
#ifndef BB25_GROUP_CTAS
#define BB25_GROUP_CTAS 6
#endif
template <bool SPARSE_TAB>
__global__ void __launch_bounds__(BK_WARPS * 32, BB25_GROUP_CTAS) group_kernel(const __grid_constant__ BlockArgs a) {
    __shared__ __align__(16) uint4 gsm[BK_WARPS * (QC + 1)];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint4 *sdesc = gsm + warp * (QC + 1);
    unsigned int *scnt = reinterpret_cast<unsigned int *>(sdesc + QC);
    if (lane < 4) scnt[lane] = 0u;
    __syncwarp();
    const int n_q = a.n_q_ptr ? (int)*a.n_q_ptr : a.n_q;
    const int n_chunks = (n_q + QC - 1) / QC;
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    if (a.stats && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&a.stats[0], (unsigned long long)(a.blk_end - a.blk_begin) * (unsigned long long)n_q);
    for (;;) {
        long long item = 0;
        if (lane == 0) item = (long long)atomicAdd(a.work_counter, 1ull);
        item = shfl_ll(item, 0);
        if (item >= n_items) break;
        const int blk = a.blk_begin + (int)(item / n_chunks);
        const int slot0 = (int)(item % n_chunks) * QC;
        const int nslots = min(QC, n_q - slot0);
        __syncwarp();
        if (lane < nslots) {
            const int my_q = a.q_list ? a.q_list[slot0 + lane] : slot0 + lane;
            const long long t0 = a.q_off[my_q];
            const int my_m = (int)max(0ll, (long long)a.q_off[my_q + 1] - t0);
            sdesc[lane] = make_uint4((unsigned)my_q, (unsigned)my_m, (unsigned)(int)(t0 - a.term_base),
                                     (uint32_t)(a.thr[my_q] >> 33));
        }
        __syncwarp();
        const unsigned serial = group_units<SPARSE_TAB>(a, sdesc, nslots, blk, blk * kBlockDocs, lane, scnt);
        if (lane == 0) a.unit_mask[item] = (uint8_t)serial;
    }
    __syncwarp();
    if (lane == 0 && a.stats) {
        if (scnt[0]) atomicAdd(&a.stats[1], (unsigned long long)scnt[0]);
        if (scnt[2]) atomicAdd(&a.stats[3], (unsigned long long)scnt[2]);
    }
}

static int launch_block(const bb25_index *idx, const BlockArgs &a, bool exact, cudaStream_t st) {
    const int n_chunks = (a.n_q + QC - 1) / QC;  // a.n_q: host-side upper bound of the query count
    const long long n_items = (long long)(a.blk_end - a.blk_begin) * n_chunks;
    if (n_items <= 0) return 0;
    BB25_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    const size_t smem = (size_t)BK_WARPS * (kBlockDocs * 4 + (QC + 1) * sizeof(uint4));
    const bool pruned_cfg = a.prune != 0;
    int per_sm = exact ? 6 : (pruned_cfg ? BB25_BLOCK_CTAS_PRUNED : BB25_BLOCK_CTAS);
    if (const char *e = getenv("BB25_CTAS_PER_SM")) {
        const int v = atoi(e);
        if (v >= 1 && v <= per_sm) per_sm = v;
    }
    long long grid = (long long)idx->sm_count * per_sm;
    const long long need = (n_items + BK_WARPS - 1) / BK_WARPS;
    if (grid > need) grid = need;
    const bool sparse_tab = idx->tab_sparse_terms > 0;
    if (!exact && pruned_cfg && a.sparse_mode && a.unit_mask) {
        long long ggrid = (long long)idx->sm_count * BB25_GROUP_CTAS;
        if (ggrid > need) ggrid = need;
        if (sparse_tab) group_kernel<true><<<(unsigned)ggrid, BK_WARPS * 32, 0, st>>>(a);
        else group_kernel<false><<<(unsigned)ggrid, BK_WARPS * 32, 0, st>>>(a);
        BB25_LAUNCH_CHECK();
        BB25_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    }
#define BB25_LAUNCH_BLOCK_G(EX, SP, CT, HF, GR)                                                                       \
    do {                                                                                                              \
        BB25_CUDA(cudaFuncSetAttribute(block_kernel<BK_WARPS, EX, SP, CT, HF, GR>,                                    \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
        block_kernel<BK_WARPS, EX, SP, CT, HF, GR><<<(unsigned)grid, BK_WARPS * 32, smem, st>>>(a);                   \
    } while (0)
#define BB25_LAUNCH_BLOCK(EX, SP, CT, HF) BB25_LAUNCH_BLOCK_G(EX, SP, CT, HF, false)
    // fp16 bound rows pay only where units are skipped or restricted (+4 % pruned); the exhaustive pass is issue-bound and
    // the conversions cost it 0.3 % (8.8 M documents) to 4 % (shard-sized corpora): fp32 rows there (profiles/r02/half_rows_ab.txt)
    const bool half_rows = !exact && a.dense_h != nullptr && a.prune != 0;
    if (exact) {
        if (sparse_tab) BB25_LAUNCH_BLOCK(true, true, 6, false);
        else BB25_LAUNCH_BLOCK(true, false, 6, false);
    } else if (pruned_cfg && a.sparse_mode && a.unit_mask) {
        if (sparse_tab) { if (half_rows) BB25_LAUNCH_BLOCK_G(false, true, BB25_BLOCK_CTAS_PRUNED, true, true); else BB25_LAUNCH_BLOCK_G(false, true, BB25_BLOCK_CTAS_PRUNED, false, true); }
        else { if (half_rows) BB25_LAUNCH_BLOCK_G(false, false, BB25_BLOCK_CTAS_PRUNED, true, true); else BB25_LAUNCH_BLOCK_G(false, false, BB25_BLOCK_CTAS_PRUNED, false, true); }
    } else if (pruned_cfg) {
        if (sparse_tab) { if (half_rows) BB25_LAUNCH_BLOCK(false, true, BB25_BLOCK_CTAS_PRUNED, true); else BB25_LAUNCH_BLOCK(false, true, BB25_BLOCK_CTAS_PRUNED, false); }
        else { if (half_rows) BB25_LAUNCH_BLOCK(false, false, BB25_BLOCK_CTAS_PRUNED, true); else BB25_LAUNCH_BLOCK(false, false, BB25_BLOCK_CTAS_PRUNED, false); }
    } else {
        if (sparse_tab) { if (half_rows) BB25_LAUNCH_BLOCK(false, true, BB25_BLOCK_CTAS, true); else BB25_LAUNCH_BLOCK(false, true, BB25_BLOCK_CTAS, false); }
        else { if (half_rows) BB25_LAUNCH_BLOCK(false, false, BB25_BLOCK_CTAS, true); else BB25_LAUNCH_BLOCK(false, false, BB25_BLOCK_CTAS, false); }
    }
#undef BB25_LAUNCH_BLOCK
#undef BB25_LAUNCH_BLOCK_G
    BB25_LAUNCH_CHECK();
    return 0;
}


// =================================================================================
// Candidate-driven evaluation (pruning level 3, query-level MaxScore).  For a query
// whose threshold seed exceeds the summed GLOBAL maxima of its most frequent terms,
// those terms are non-essential for the whole corpus: only documents holding one of the
// remaining (essential, rare) terms can reach the top-k.  Such a query is routed away
// from the block traversal: one thread per posting of an essential term evaluates that
// document completely -- every query term looked up in query order (dense value row, or
// block-table entry + binary search of the <= 1024-posting slice), the fp32 sum formed
// in exactly bm25s's order -- and emits it if it passes the threshold.  A document
// reachable through several essential terms is evaluated by the first of them only.
// Work is proportional to the essential terms' document frequencies, not to N.
// =================================================================================
constexpr int kCandChunk = 1024;  // postings per work item

struct RouteArgs {
    const int32_t *q_terms;  // sanitised copy
    const int64_t *q_off;
    int64_t term_base;
    int64_t n_q;
    const unsigned long long *thr;
    const float *gmax;  // per-term global maximum posting value
    const int64_t *indptr;
    long long route_max;  // route only if the essential terms' summed df is <= this
    uint32_t *ne_mask;    // [n_q] bit i = i-th query term is non-essential
    int32_t *list_a, *list_b;
    unsigned int *n_a, *n_b;
    uint2 *items;  // {query, position << 24 | chunk}
    unsigned int *n_items;
    unsigned int items_cap;
};

__device__ inline unsigned int count_items(const RouteArgs &a, long long t0, int m, uint32_t ne) {
    unsigned int n = 0;
    for (int i = 0; i < m; i++)
        if (!((ne >> i) & 1u)) {
            const int t = a.q_terms[t0 + i];
            n += (unsigned int)((a.indptr[t + 1] - a.indptr[t] + kCandChunk - 1) / kCandChunk);
        }
    return n;
}
__device__ inline void write_items(const RouteArgs &a, int q, long long t0, int m, uint32_t ne, unsigned int base) {
    for (int i = 0; i < m; i++)
        if (!((ne >> i) & 1u)) {
            const int t = a.q_terms[t0 + i];
            const unsigned int nch = (unsigned int)((a.indptr[t + 1] - a.indptr[t] + kCandChunk - 1) / kCandChunk);
            for (unsigned int c = 0; c < nch; c++) a.items[base++] = make_uint2((unsigned int)q, ((unsigned int)i << 24) | c);
        }
}

// one thread per query: essential / non-essential split against the GLOBAL term maxima,
// routing decision, work items of the routed queries
__global__ void route_kernel(const RouteArgs a) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_q) return;
    const long long t0 = a.q_off[q] - a.term_base;
    const int m = (int)max(0ll, (long long)(a.q_off[q + 1] - a.q_off[q]));
    const float thr_val = __uint_as_float((uint32_t)(a.thr[q] >> 33));
    uint32_t ne = 0u;
    bool routed = false;
    if (m >= 1 && m <= 24 && thr_val > 0.f) {
        // greedy: take terms in ascending order of their maximum while the running sum,
        // with a 1e-5 relative margin for the summation order, stays below the threshold
        float run = 0.f;
        for (;;) {
            int best = -1;
            float bv = 0.f;
            for (int i = 0; i < m; i++)
                if (!((ne >> i) & 1u)) {
                    const float g = a.gmax[a.q_terms[t0 + i]];
                    if (best < 0 || g < bv) { best = i; bv = g; }
                }
            if (best < 0) break;
            const float cand = __fadd_rn(run, bv);
            if (!(__fmul_rn(cand, 1.00001f) < thr_val)) break;
            run = cand;
            ne |= 1u << best;
        }
        if (ne != 0u) {
            long long w = 0;
            for (int i = 0; i < m; i++)
                if (!((ne >> i) & 1u)) {
                    const int t = a.q_terms[t0 + i];
                    w += a.indptr[t + 1] - a.indptr[t];
                }
            if (w <= a.route_max) {
                const unsigned int cnt = count_items(a, t0, m, ne);
                const unsigned int base = atomicAdd(a.n_items, cnt);
                if (base + cnt <= a.items_cap) {
                    write_items(a, (int)q, t0, m, ne, base);
                    routed = true;
                } else {
                    atomicSub(a.n_items, cnt);  // does not fit: leave the query to the block traversal
                }
            }
        }
    }
    a.ne_mask[q] = routed ? ne : 0u;
    if (routed) a.list_a[atomicAdd(a.n_a, 1u)] = (int)q;
    else a.list_b[atomicAdd(a.n_b, 1u)] = (int)q;
}

// work items again for routed queries whose candidate row overflowed
__global__ void rebuild_items_kernel(const RouteArgs a, const int32_t *list, const unsigned int *n_list) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int)*n_list) return;
    const int q = list[idx];
    const long long t0 = a.q_off[q] - a.term_base;
    const int m = (int)(a.q_off[q + 1] - a.q_off[q]);
    const uint32_t ne = a.ne_mask[q];
    const unsigned int cnt = count_items(a, t0, m, ne);
    const unsigned int base = atomicAdd(a.n_items, cnt);
    if (base + cnt <= a.items_cap) write_items(a, q, t0, m, ne, base);
}

struct CandArgs {
    const float *data;
    const int32_t *indices;
    const int64_t *indptr;
    BlockTable tab;
    int64_t n_vocab;
    const int32_t *dense_slot;
    const float *dense_vals;
    int64_t dense_stride;
    const int32_t *q_terms;
    const int64_t *q_off;
    int64_t term_base;
    const uint32_t *ne_mask;
    const uint2 *items;
    const unsigned int *n_items;  // device-side number of work items
    const unsigned long long *thr;
    unsigned int *cand_cnt;
    unsigned long long *cand_key;
    int cap;
    const float *gmax;
};

__global__ void __launch_bounds__(256) cand_kernel(const __grid_constant__ CandArgs a) {
    __shared__ int s_term[24];
    __shared__ int s_slot[24];
    __shared__ long long s_base[24];
    __shared__ longlong2 s_row[24];
    __shared__ float s_rest[25];  // s_rest[i] = sum of the global maxima of terms at positions >= i, except `pos`
    const unsigned int n_items = *a.n_items;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        __syncthreads();  // the previous item's shared descriptors are no longer read
        const uint2 item = a.items[it];
        const int q = (int)item.x;
        const int pos = (int)(item.y >> 24);
        const unsigned int chunk = item.y & 0xFFFFFFu;
        const long long t0 = a.q_off[q] - a.term_base;
        const int m = (int)(a.q_off[q + 1] - a.q_off[q]);
        if (threadIdx.x < m) {
            const int t = a.q_terms[t0 + threadIdx.x];
            s_term[threadIdx.x] = t;
            s_slot[threadIdx.x] = a.dense_slot[t];
            s_base[threadIdx.x] = a.indptr[t];
            s_row[threadIdx.x] = a.tab.row[t];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float r = 0.f;
            s_rest[m] = 0.f;
            for (int i = m - 1; i >= 0; i--) {
                if (i != pos) r = __fadd_rn(r, a.gmax[s_term[i]]);
                s_rest[i] = r;
            }
        }
        __syncthreads();
        const uint32_t ne = a.ne_mask[q];
        const unsigned long long thr = a.thr[q];
        const uint32_t thr_score = (uint32_t)(thr >> 33);
        const float thr_val = __uint_as_float(thr_score);
        const long long ebase = s_base[pos];
        const long long df = a.indptr[s_term[pos] + 1] - ebase;
        unsigned int *ccnt = a.cand_cnt + q;
        unsigned long long *crow = a.cand_key + (size_t)q * (size_t)a.cap;
#pragma unroll 1
        for (int u = 0; u < kCandChunk / 256; u++) {
            const long long j = (long long)chunk * kCandChunk + u * 256 + threadIdx.x;
            if (j >= df) break;
            const uint32_t d = (uint32_t)ld_nc_s32(a.indices + ebase + j);
            const float ve = ld_nc_f32(a.data + ebase + j);
            // MaxScore bound: own value + the other terms' global maxima (1e-5 relative margin
            // for the summation order); tightened term by term as actual values replace maxima
            if (__fmul_rn(__fadd_rn(ve, s_rest[0]), 1.00001f) < thr_val) continue;
            float acc = 0.f;
            bool dup = false;
            for (int i = 0; i < m; i++) {
                if (i == pos) {
                    acc = __fadd_rn(acc, ve);
                    continue;
                }
                {
                    const float rest = i < pos ? __fadd_rn(s_rest[i], ve) : s_rest[i];
                    if (__fmul_rn(__fadd_rn(acc, rest), 1.00001f) < thr_val) {
                        dup = true;  // cannot reach the threshold any more
                        break;
                    }
                }
                float val = 0.f;
                bool present = false;
                const int slot = s_slot[i];
                if (slot >= 0) {
                    val = a.dense_vals[(size_t)slot * (size_t)a.dense_stride + d];
                    present = __float_as_uint(val) != 0x80000000u;
                } else {
                    const uint2 ent = tab_lookup(a.tab, s_row[i], (int)(d >> 10));
                    const int len = (int)(ent.y & kBlkLenMask);
                    if (len) {
                        // plain bisection: the threads of a warp search different terms' slices, and the uniform
                        // trip count beats the data-dependent probing of slice_find here (measured)
                        long long lo = s_base[i] + (long long)ent.x;
                        const long long end = lo + len;
                        long long hi = end;
                        while (lo < hi) {
                            const long long mid = (lo + hi) >> 1;
                            if ((uint32_t)a.indices[mid] < d) lo = mid + 1;
                            else hi = mid;
                        }
                        if (lo < end && (uint32_t)a.indices[lo] == d) {
                            present = true;
                            val = a.data[lo];
                        }
                    }
                }
                if (present) {
                    if (i < pos && !((ne >> i) & 1u)) {  // an earlier essential term owns this document
                        dup = true;
                        break;
                    }
                    acc = __fadd_rn(acc, val);
                }
            }
            if (!dup) emit_if_candidate(acc, d, thr_score, thr, ccnt, crow, a.cap);
        }
    }
}

// traversal kernel family used by retrieve: BB25_KERNEL=tile selects the CTA-tile kernel
static bool use_block_kernel() {
    const char *e = getenv("BB25_KERNEL");
    return !(e && e[0] == 't');
}

__global__ void fill_strided_f64_kernel(double *out, int64_t n, int64_t stride, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i * stride] = v;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <int DV, int NTV, int MODE>
static int launch_tile_inst(const bb25_index *idx, const TileArgs &a, long long n_items, cudaStream_t st) {
    constexpr bool has_cnt = (MODE != MODE_RETRIEVE);
    const size_t smem = (size_t)DV * (has_cnt ? 5 : 4) + sizeof(TileMeta);
    BB25_CUDA(cudaFuncSetAttribute(tile_kernel<DV, NTV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (long long)idx->sm_count * ctas_per_sm<DV, MODE>();
    if (grid > n_items) grid = n_items;
    tile_kernel<DV, NTV, MODE><<<(unsigned)grid, NTV, smem, st>>>(a);
    BB25_LAUNCH_CHECK();
    return 0;
}

template <int MODE>
static int launch_tile(const bb25_index *idx, const TileArgs &a, cudaStream_t st) {
    const int n_chunks = (a.n_q + QB - 1) / QB;
    const long long n_items = (long long)(a.tile_end - a.tile_begin) * n_chunks;
    if (n_items <= 0) return 0;
    BB25_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), st));
    switch (idx->tile_docs) {
    case 8192: return launch_tile_inst<8192, 256, MODE>(idx, a, n_items, st);
    case 16384: return launch_tile_inst<16384, 512, MODE>(idx, a, n_items, st);
    case 32768: return launch_tile_inst<32768, 1024, MODE>(idx, a, n_items, st);
    default: set_error("unsupported tile size %d", idx->tile_docs); return 1;
    }
}

static void base_args(const bb25_index *idx, TileArgs &a) {
    a.data = idx->data;
    a.indices = idx->indices;
    a.indptr = idx->indptr;
    a.tile_off = idx->tile_off;
    a.doc_len = idx->doc_len;
    a.avgdl = idx->avgdl;
    a.n_tiles = idx->n_tiles;
    a.n_docs = idx->n_docs;
}


// dense single-query outputs (get_scores / get_probabilities)
__global__ void fuse_const_kernel(double *acc, int64_t n, double p, FuseSpec f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = fuse_step(acc[i], d_logit(p), f);
}

// a short query passed by value: no host->device copy, no stream synchronisation
struct InlineQuery {
    int n;
    int32_t t[64];
};
__global__ void stage_query_kernel(const InlineQuery iq, int32_t *d_src, int64_t *d_qoff) {
    if (threadIdx.x < iq.n) d_src[threadIdx.x] = iq.t[threadIdx.x];
    if (threadIdx.x == 0) {
        d_qoff[0] = 0;
        d_qoff[1] = iq.n;
    }
}
// the same for a query that already sits on the device (sanitised workspace copy of a batch)
__global__ void stage_query_dev_kernel(const int32_t *__restrict__ src, int n, int32_t *d_src, int64_t *d_qoff) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) d_src[i] = src[i];
    if (threadIdx.x == 0) {
        d_qoff[0] = 0;
        d_qoff[1] = n;
    }
}

// workspace plan of the dense single-query passes
struct DenseWs {
    int32_t *d_terms;
    uint8_t *d_nc;
    int64_t *d_qoff, *d_qo;
    int32_t *d_src;
    unsigned long long *d_ctr;
    int *d_err;
};
static int dense_workspace(bb25_index *idx, int n_terms, size_t extra, DenseWs &w, unsigned char **extra_ptr) {
    const size_t nt = (size_t)(n_terms > 0 ? n_terms : 1);
    const size_t o_terms = 0;
    const size_t o_nc = align_up(o_terms + sizeof(int32_t) * nt);
    const size_t o_qoff = align_up(o_nc + nt);
    const size_t o_qo = align_up(o_qoff + 2 * sizeof(int64_t));
    const size_t o_src = align_up(o_qo + 2 * sizeof(int64_t));
    const size_t o_ctr = align_up(o_src + sizeof(int32_t) * nt);
    const size_t o_err = align_up(o_ctr + sizeof(unsigned long long));
    const size_t o_extra = align_up(o_err + sizeof(int));
    if (ensure_workspace(idx, o_extra + extra)) return 1;
    unsigned char *ws = (unsigned char *)idx->ws;
    w.d_terms = (int32_t *)(ws + o_terms);
    w.d_nc = ws + o_nc;
    w.d_qoff = (int64_t *)(ws + o_qoff);
    w.d_qo = (int64_t *)(ws + o_qo);
    w.d_src = (int32_t *)(ws + o_src);
    w.d_ctr = (unsigned long long *)(ws + o_ctr);
    w.d_err = (int *)(ws + o_err);
    if (extra_ptr) *extra_ptr = ws + o_extra;
    return 0;
}

// the traversal pass of a dense single-query output; the query's terms are staged at w.d_src
static int run_dense_staged(bb25_index *idx, int mode, const bb25_params *params, const DenseWs &w, int n_terms,
                            float *out_scores, double *out_probs, int64_t out_stride, cudaStream_t st,
                            const FuseSpec *fuse) {
    if (ensure_tile_table(idx, st)) return 1;
    BB25_CUDA(cudaMemsetAsync(w.d_err, 0, sizeof(int), st));
    prep_queries_kernel<<<1, 32, 0, st>>>(w.d_src, w.d_qoff, 1, 0, n_terms, idx->n_vocab, nullptr, w.d_terms, w.d_nc, w.d_qo,
                                          nullptr, nullptr, nullptr, w.d_err);
    BB25_LAUNCH_CHECK();
    TileArgs a{};
    base_args(idx, a);
    a.q_terms = w.d_terms;
    a.q_nocount = w.d_nc;
    a.q_off = w.d_qo;
    a.term_base = 0;
    a.q_list = nullptr;
    a.n_q = 1;
    a.tile_begin = 0;
    a.tile_end = idx->n_tiles;
    a.out_scores = out_scores;
    a.out_probs = out_probs;
    a.out_stride = out_stride;
    if (params) a.params = *params;
    a.work_counter = w.d_ctr;
    if (fuse) a.fuse = *fuse;
    if (mode == MODE_SCORES) return launch_tile<MODE_SCORES>(idx, a, st);
    if (mode == MODE_FUSED) return launch_tile<MODE_FUSED>(idx, a, st);
    return launch_tile<MODE_PROBS>(idx, a, st);
}

static int run_dense(bb25_index *idx, int mode, const bb25_params *params, const int32_t *q_terms_host,
                     int n_terms, float *out_scores, double *out_probs, int64_t out_stride,
                     cudaStream_t st, const FuseSpec *fuse = nullptr) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (n_terms < 0 || (n_terms > 0 && !q_terms_host)) { set_error("bad query"); return 1; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    for (int i = 0; i < n_terms; i++)
        if (q_terms_host[i] < 0 || q_terms_host[i] >= idx->n_vocab) {
            set_error("query term id %d out of range [0, %lld)", q_terms_host[i], (long long)idx->n_vocab);
            return 1;
        }
    if (n_terms == 0) {
        if (mode == MODE_SCORES) {
            BB25_CUDA(cudaMemsetAsync(out_scores, 0, sizeof(float) * (size_t)idx->n_docs, st));
        } else if (mode == MODE_FUSED) {
            // no query term in the vocabulary: every document enters with probability 0 -> logit(1e-10)
            fuse_const_kernel<<<(unsigned)((idx->n_docs + 255) / 256), 256, 0, st>>>(out_probs, idx->n_docs,
                                                                                    0.0, *fuse);
            BB25_LAUNCH_CHECK();
        } else {
            fill_strided_f64_kernel<<<(unsigned)((idx->n_docs + 255) / 256), 256, 0, st>>>(out_probs, idx->n_docs, out_stride, 0.0);
            BB25_LAUNCH_CHECK();
        }
        return 0;
    }
    DenseWs w;
    ws_acquire(idx, st);
    if (dense_workspace(idx, n_terms, 0, w, nullptr)) return 1;
    if (n_terms <= 64) {
        InlineQuery iq;
        iq.n = n_terms;
        for (int i = 0; i < n_terms; i++) iq.t[i] = q_terms_host[i];
        stage_query_kernel<<<1, 64, 0, st>>>(iq, w.d_src, w.d_qoff);
        BB25_LAUNCH_CHECK();
    } else {
        int64_t hq[2] = {0, n_terms};
        BB25_CUDA(cudaMemcpyAsync(w.d_src, q_terms_host, sizeof(int32_t) * (size_t)n_terms, cudaMemcpyHostToDevice, st));
        BB25_CUDA(cudaMemcpyAsync(w.d_qoff, hq, sizeof(hq), cudaMemcpyHostToDevice, st));
        BB25_CUDA(cudaStreamSynchronize(st));  // hq / q_terms_host are pageable stack/user memory
    }
    const int rc = run_dense_staged(idx, mode, params, w, n_terms, out_scores, out_probs, out_stride, st, fuse);
    ws_release(idx, st);
    return rc;
}

static int check_params(const bb25_params *p) {
    if (!p) { set_error("params is NULL"); return 1; }
    if (p->has_base_rate && !(p->base_rate > 0.0 && p->base_rate < 1.0)) {
        set_error("base_rate must be in (0, 1), got %g", p->base_rate);
        return 1;
    }
    if (p->prior_mode != 0 && p->prior_mode != 1) { set_error("prior_mode must be 0 or 1"); return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------
// Guaranteed path for one query: dense fp32 scores of every document, exact top-k of the
// whole vector by (score desc, doc id asc), probabilities of the k winners from the dense
// posterior pass.  O(N) per query, no candidate rows, no thresholds -- used for queries whose
// threshold refinement does not settle (adversarial tie structures) and for k beyond the
// candidate kernel's limit.
// ---------------------------------------------------------------------------------
__global__ void widen_scores_kernel(const float *__restrict__ s, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)s[i];  // exact: ranking by the fp64 value is ranking by the fp32 score
}
__global__ void gather_dense_row_kernel(const int64_t *__restrict__ ids, const double *__restrict__ vals,
                                        const double *__restrict__ probs, int k, int64_t doc_id_offset,
                                        int64_t *__restrict__ out_ids, float *__restrict__ out_scores,
                                        double *__restrict__ out_probs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    const int64_t d = ids[r];
    out_ids[r] = d + doc_id_offset;
    if (out_scores) out_scores[r] = (float)vals[r];
    out_probs[r] = probs[d];
}

constexpr int kMaxDenseK = 8192;

// caller holds idx->mu and the device guard; the query's terms are staged in w.d_src already
static int dense_topk_staged(bb25_index *idx, const bb25_params *params, const DenseWs &w, unsigned char *scratch,
                             int n_terms, int k, int64_t *out_ids, float *out_scores, double *out_probs,
                             cudaStream_t st) {
    const size_t n = (size_t)idx->n_docs;
    float *d_sc = (float *)scratch;
    double *d_wide = (double *)(scratch + align_up(n * 4));
    double *d_pr = (double *)(scratch + align_up(n * 4) + align_up(n * 8));
    int64_t *d_ids = (int64_t *)(scratch + align_up(n * 4) + 2 * align_up(n * 8));
    double *d_vals = (double *)(scratch + align_up(n * 4) + 2 * align_up(n * 8) + align_up((size_t)k * 8));
    if (n_terms == 0) {
        BB25_CUDA(cudaMemsetAsync(d_wide, 0, n * 8, st));
        BB25_CUDA(cudaMemsetAsync(d_pr, 0, n * 8, st));
    } else {
        if (run_dense_staged(idx, MODE_SCORES, nullptr, w, n_terms, d_sc, nullptr, 1, st, nullptr)) return 1;
        widen_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_sc, (int64_t)n, d_wide);
        BB25_LAUNCH_CHECK();
        if (run_dense_staged(idx, MODE_PROBS, params, w, n_terms, nullptr, d_pr, 1, st, nullptr)) return 1;
    }
    if (bb25_topk_f64(idx->device, d_wide, (int64_t)n, k, d_ids, d_vals, st)) return 1;
    gather_dense_row_kernel<<<(k + 255) / 256, 256, 0, st>>>(d_ids, d_vals, d_pr, k, idx->doc_id_offset, out_ids, out_scores,
                                                             out_probs);
    BB25_LAUNCH_CHECK();
    return 0;
}
static size_t dense_topk_scratch_bytes(const bb25_index *idx, int k) {
    const size_t n = (size_t)idx->n_docs;
    return align_up(n * 4) + 2 * align_up(n * 8) + 2 * align_up((size_t)k * 8);
}

// ---------------------------------------------------------------------------------
// small device-side bookkeeping of the batch pipeline
// ---------------------------------------------------------------------------------
// queries still overflowing after the last repair round of a stage are set aside
__global__ void mark_bad_kernel(const int32_t *__restrict__ list, const unsigned int *__restrict__ n_list,
                                uint8_t *__restrict__ bad) {
    const unsigned int n = *n_list;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) bad[list[i]] = 1;
}
__global__ void compact_bad_kernel(const uint8_t *__restrict__ bad, int64_t n_q, int32_t *__restrict__ list,
                                   unsigned int *__restrict__ n_list, unsigned int *__restrict__ cand_cnt,
                                   unsigned int *__restrict__ n_prev) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q || !bad[q]) return;
    list[atomicAdd(n_list, 1u)] = (int32_t)q;
    cand_cnt[q] = 0;  // everything collected so far is collected again by the repair traversal
    n_prev[q] = 0;
}

// what the host reads back, once, at the end of a batch
struct BatchReport {
    unsigned long long n_cand, units, units_skipped, units_maxscore, units_sparse;
    unsigned int route[4];  // queries on the candidate path, on the block path, candidate work items (first round), -
    unsigned int n_bad;
    int err;
    unsigned int reruns;
    unsigned int pad;
};
__global__ void report_kernel(const unsigned long long *__restrict__ n_cand, const unsigned long long *__restrict__ stats,
                              const unsigned int *__restrict__ route, const unsigned int *__restrict__ n_bad,
                              const int *__restrict__ err, const unsigned int *__restrict__ round_cnt, int n_round_cnt,
                              BatchReport *out) {
    BatchReport r;
    r.n_cand = *n_cand;
    r.units = stats[0];
    r.units_skipped = stats[1];
    r.units_maxscore = stats[2];
    r.units_sparse = stats[3];
    r.route[0] = route[0];
    r.route[1] = route[1];
    r.route[2] = route[3];  // first-round work items (route[2] is reused by the repair rounds)
    r.route[3] = 0;
    r.n_bad = *n_bad;
    r.err = *err;
    unsigned int re = 0;
    for (int i = 0; i < n_round_cnt; i++) re += round_cnt[i];
    r.reruns = re;
    r.pad = 0;
    *out = r;
}
__global__ void copy_u32_kernel(const unsigned int *src, unsigned int *dst) { *dst = *src; }
// rank-k entry of the published quantiles = the shard's current threshold (score bits)
__global__ void publish_thr_kernel(const unsigned long long *__restrict__ thr, int64_t n_q, int J,
                                   unsigned long long *__restrict__ quant) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_q) quant[q * J + (J - 1)] = thr[q] & ~((1ull << 33) - 1ull);
}

constexpr int kMaxRepairRounds = 6;

static int retrieve_device(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                           const int64_t *q_off, int64_t n_q, int64_t term_base, int64_t n_terms_total,
                           int k, int64_t *out_ids, float *out_scores, double *out_probs,
                           cudaStream_t st) {
    // caller holds idx->mu and the device guard.  The whole batch is enqueued without a host
    // round trip: list lengths and overflow counts stay on the device, every stage is followed by
    // a fixed number of repair rounds (kernels that find an empty list leave at once), and ONE
    // synchronisation at the end reads the report.
    int sparse_mode = idx->prune >= 2 && use_block_kernel() ? 1 : 0;
    if (const char *e = getenv("BB25_SPARSE")) sparse_mode = sparse_mode && atoi(e) != 0;
    if (sparse_mode) {
        // value rows of the mid-frequency terms for the essential-posting evaluation (built once per index)
        if (ensure_lookup_rows(idx, st)) return 1;
        if (!idx->lookup_vals && !idx->dense_vals) sparse_mode = 0;
    }
    int cap = 1024;
    while (cap < 8 * k && cap < 16384) cap <<= 1;
    int n_rounds = 3;  // repair rounds enqueued after every stage
    if (const char *e = getenv("BB25_REPAIR_ROUNDS")) {
        const int v = atoi(e);
        if (v >= 0 && v <= kMaxRepairRounds) n_rounds = v;
    }
    const size_t nt = (size_t)(n_terms_total > 0 ? n_terms_total : 1);
    const size_t o_terms = 0;
    const size_t o_nc = align_up(o_terms + sizeof(int32_t) * nt);
    const size_t o_qo = align_up(o_nc + nt);
    const size_t o_thr = align_up(o_qo + sizeof(int64_t) * (size_t)(n_q + 1));
    const size_t o_cnt = align_up(o_thr + sizeof(unsigned long long) * (size_t)n_q);
    const size_t o_prev = align_up(o_cnt + sizeof(unsigned int) * (size_t)n_q);
    const size_t o_la = align_up(o_prev + sizeof(unsigned int) * (size_t)n_q);
    const size_t o_lb = align_up(o_la + sizeof(int32_t) * (size_t)n_q);
    const size_t o_ctr = align_up(o_lb + sizeof(int32_t) * (size_t)n_q);
    const size_t kCtrBytes = 1024;
    const size_t o_ne = align_up(o_ctr + kCtrBytes);
    const size_t o_lista = align_up(o_ne + sizeof(uint32_t) * (size_t)n_q);
    const size_t o_listb = align_up(o_lista + sizeof(int32_t) * (size_t)n_q);
    const size_t o_bad = align_up(o_listb + sizeof(int32_t) * (size_t)n_q);
    const size_t o_badlist = align_up(o_bad + (size_t)n_q);
    // candidate-path work items: <= route_max / chunk + 24 per query
    int route_div = 32;
    if (const char *e = getenv("BB25_ROUTE_DIV")) {
        const int v = atoi(e);
        if (v >= 2 && v <= 4096) route_div = v;
    }
    const long long route_max = (long long)idx->n_docs / route_div;
    const bool use_cand = use_block_kernel() && idx->prune >= 3 && idx->dense_slot != nullptr && n_q > 0;
    const size_t items_cap = use_cand ? (size_t)n_q * (size_t)(route_max / kCandChunk + 25) : 1;
    const size_t o_quant = align_up(o_badlist + sizeof(int32_t) * (size_t)n_q);
    const size_t o_info = align_up(o_quant + sizeof(unsigned long long) * 4 * (size_t)n_q);
    const size_t o_items = align_up(o_info + 2 * sizeof(longlong2) * nt);
    const size_t o_mask = align_up(o_items + sizeof(uint2) * items_cap);
    const size_t mask_bytes = sparse_mode ? (size_t)idx->n_blocks * (size_t)((n_q + QC - 1) / QC) : 1;
    const size_t o_key = align_up(o_mask + mask_bytes);
    const size_t total = o_key + sizeof(unsigned long long) * (size_t)n_q * (size_t)cap;
    if (ensure_workspace(idx, total)) return 1;
    unsigned char *ws = (unsigned char *)idx->ws;
    int32_t *d_terms = (int32_t *)(ws + o_terms);
    uint8_t *d_nc = ws + o_nc;
    int64_t *d_qo = (int64_t *)(ws + o_qo);
    unsigned long long *d_thr = (unsigned long long *)(ws + o_thr);
    unsigned int *d_cnt = (unsigned int *)(ws + o_cnt);
    unsigned int *d_prev = (unsigned int *)(ws + o_prev);
    int32_t *d_list[2] = {(int32_t *)(ws + o_la), (int32_t *)(ws + o_lb)};
    // counter block (zeroed once per batch)
    unsigned long long *d_work = (unsigned long long *)(ws + o_ctr);
    int *d_err = (int *)(ws + o_ctr + 16);
    unsigned long long *d_ncand = (unsigned long long *)(ws + o_ctr + 24);
    unsigned long long *d_stats = (unsigned long long *)(ws + o_ctr + 32);  // [4]
    unsigned int *d_route = (unsigned int *)(ws + o_ctr + 64);  // [0] n_a, [1] n_b, [2] n_items, [3] n_items of the first round
    unsigned int *d_nbad = (unsigned int *)(ws + o_ctr + 80);
    unsigned int *d_round = (unsigned int *)(ws + o_ctr + 128);  // overflow counts: [stage][round], stage 0 = candidate path
    const int kRoundStride = kMaxRepairRounds + 2;
    BatchReport *d_report = (BatchReport *)(ws + o_ctr + 512);
    uint8_t *d_bad = ws + o_bad;
    int32_t *d_badlist = (int32_t *)(ws + o_badlist);
    unsigned long long *d_quant = (unsigned long long *)(ws + o_quant);
    int n_quant = 0, qranks[4] = {k, k, k, k};
    if (idx->exchange_cb) bb25_quantile_ranks(k, idx->exchange_shards, &n_quant, qranks);
    unsigned long long *d_keys = (unsigned long long *)(ws + o_key);

    const float *kth = nullptr;
    if (get_kth_values(idx, k, st, &kth)) return 1;
    idx->st_launches = idx->st_passes = idx->st_reruns = idx->st_candidates = 0;
    idx->st_syncs = 0;
    idx->ev_used = 0;
    idx->st_traverse_ms = 0.0;
    idx->st_traverse_launches = 0;
    int64_t launches0 = (int64_t)bb25_launch_count();

    BB25_CUDA(cudaMemsetAsync(ws + o_ctr, 0, kCtrBytes, st));
    BB25_CUDA(cudaMemsetAsync(d_bad, 0, (size_t)n_q, st));
    prep_queries_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(q_terms, q_off, n_q, term_base, (int64_t)n_terms_total,
                                                                      idx->n_vocab, kth, d_terms, d_nc, d_qo, d_thr, d_cnt,
                                                                      d_prev, d_err, idx->indptr,
                                                                      idx->dense_vals ? idx->dense_slot : nullptr,
                                                                      idx->tab_row, (longlong2 *)(ws + o_info),
                                                                      idx->lookup_vals ? idx->row_slot : nullptr);
    BB25_LAUNCH_CHECK();

    // tile groups: a small first group makes a loose threshold seed cheap to repair,
    // later groups run with the exact k-th key of everything seen so far
    const bool blockk = use_block_kernel();
    int bounds[4];
    int ng = 0;
    const int T = blockk ? idx->n_blocks : idx->n_tiles;
    bounds[0] = 0;
    int want_groups = 3;
    if (const char *e = getenv("BB25_GROUPS")) want_groups = atoi(e);  // tuning: 1, 2 or 3 block groups
    if (T >= 16 && want_groups >= 3) {
        bounds[1] = std::max(1, T / 16);
        bounds[2] = std::max(bounds[1] + 1, T / 4);
        bounds[3] = T;
        ng = 3;
    } else if (T >= 16 && want_groups == 2) {
        bounds[1] = std::max(1, T / 8);
        bounds[2] = T;
        ng = 2;
    } else {
        bounds[1] = T;
        ng = 1;
    }

    if (!use_block_kernel() && ensure_tile_table(idx, st)) return 1;
    TileArgs ta{};
    base_args(idx, ta);
    ta.q_terms = d_terms;
    ta.q_nocount = d_nc;
    ta.q_off = d_qo;
    ta.term_base = 0;
    ta.thr = d_thr;
    ta.cand_cnt = d_cnt;
    ta.cand_key = d_keys;
    ta.cap = cap;
    ta.params = *params;
    ta.work_counter = d_work;

    BlockArgs ba{};
    ba.data = idx->data;
    ba.indices = idx->indices;
    ba.indptr = idx->indptr;
    ba.tab = BlockTable{idx->tab_ent, idx->tab_bits, idx->tab_row};
    ba.n_vocab = idx->n_vocab;
    ba.q_terms = d_terms;
    ba.qt_info = (const longlong2 *)(ws + o_info);
    ba.q_off = d_qo;
    ba.term_base = 0;
    ba.thr = d_thr;
    ba.cand_cnt = d_cnt;
    ba.cand_key = d_keys;
    ba.cap = cap;
    ba.prune = idx->prune;
    // first evaluation of a group: order-free sums, candidates re-scored exactly by select_kernel;
    // threshold repairs after an overflow: query order, exact keys against the 64-bit threshold
    bool relaxed = idx->dense_vals != nullptr;
    if (const char *e = getenv("BB25_RELAXED")) relaxed = relaxed && atoi(e) != 0;
    ba.dense_slot = idx->dense_slot;
    ba.dense_vals = idx->dense_vals;
    ba.dense_h = idx->dense_h;
    if (const char *e = getenv("BB25_HALF_ROWS")) {
        if (atoi(e) == 0) ba.dense_h = nullptr;
    }
    ba.dense_stride = idx->dense_stride;
    ba.lookup_vals = idx->lookup_vals;
    ba.n_hot = idx->dense_vals ? idx->n_dense : 0;
    ba.sparse_mode = sparse_mode;
    ba.unit_mask = sparse_mode ? ws + o_mask : nullptr;
    ba.work_counter = d_work;
    ba.stats = d_stats;

    SelectArgs sa{};
    sa.cand_cnt = d_cnt;
    sa.n_prev = d_prev;
    sa.cand_key = d_keys;
    sa.thr = d_thr;
    sa.cap = cap;
    sa.k = k;
    sa.doc_len = idx->doc_len;
    sa.avgdl = idx->avgdl;
    sa.n_docs = idx->n_docs;
    sa.doc_id_offset = idx->doc_id_offset;
    sa.params = *params;
    sa.out_ids = out_ids;
    sa.out_scores = out_scores;
    sa.out_probs = out_probs;
    sa.n_cand_total = d_ncand;
    sa.tf_search = 1;  // candidate keys carry no matched-term count
    sa.data = idx->data;
    sa.rescore = 0;
    sa.indices = idx->indices;
    sa.indptr = idx->indptr;
    sa.tab = BlockTable{idx->tab_ent, idx->tab_bits, idx->tab_row};
    sa.n_vocab = idx->n_vocab;
    sa.dense_slot = idx->dense_slot;
    sa.dense_vals = idx->dense_vals;
    sa.dense_stride = idx->dense_stride;
    sa.q_terms = d_terms;
    sa.q_nocount = d_nc;
    sa.q_off = d_qo;
    sa.term_base = 0;
    sa.quant = idx->exchange_cb ? d_quant : nullptr;
    sa.n_quant = n_quant;
    for (int j = 0; j < 4; j++) sa.qrank[j] = qranks[j];

    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    // at most kpad valid keys left (the usual case after the re-scoring pre-filter): one bitonic sort
    // replaces the 8-pass radix select plus the sort of the winners; more keys: radix select first
    sa.sort_cap = kpad;
    // order-free sums formed from fp16-bound rows may exceed the exact score by up to 2^-11 relative per addend
    // fp16 bound rows (pruned passes only, see launch_block) round every value UP by as much as one fp16 ulp = 2^-10
    // relative, so the k-th largest key as it stands may overstate the true k-th score by that much: a key may be
    // dropped unseen only below (1 - 1.1e-3) of it
    sa.prefilter = (ba.dense_h && ba.prune != 0) ? 0.9989f : 0.99997f;
    const size_t sel_smem = (size_t)cap * 10 + (size_t)kpad * 9 + 260 * 4;  // keys, top, hist+st, flag, rescore list
    BB25_CUDA(cudaFuncSetAttribute(select_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    // grid of a select launch whose list length is only known on the device (repair rounds)
    const unsigned repair_grid = (unsigned)std::min<int64_t>(n_q, (int64_t)idx->sm_count * 2);

    auto timed_begin = [&]() -> int {
        const int pair = idx->ev_used < bb25_index::kMaxEv ? idx->ev_used : -1;
        if (pair >= 0) {
            while (idx->n_ev < 2 * (pair + 1)) {
                if (cudaEventCreate(&idx->ev[idx->n_ev]) != cudaSuccess) return -1;
                idx->n_ev++;
            }
            cudaEventRecord(idx->ev[2 * pair], st);
        }
        return pair;
    };
    auto timed_end = [&](int pair) {
        if (pair >= 0) {
            cudaEventRecord(idx->ev[2 * pair + 1], st);
            idx->ev_used++;
        }
    };

    idx->st_routed = 0;
    idx->st_cand_items = 0;

    // ---- candidate-driven queries (pruning level 3) ---------------------------------
    const int32_t *blk_list = nullptr;        // queries left to the block traversal (nullptr = all)
    const unsigned int *blk_n_ptr = nullptr;  // their number, on the device
    RouteArgs ra{};
    CandArgs ca{};
    if (use_cand) {
        const float *gmax = nullptr;
        if (get_kth_values(idx, 1, st, &gmax)) return 1;
        ra.q_terms = d_terms;
        ra.q_off = d_qo;
        ra.term_base = 0;
        ra.n_q = n_q;
        ra.thr = d_thr;
        ra.gmax = gmax;
        ra.indptr = idx->indptr;
        ra.route_max = route_max;
        ra.ne_mask = (uint32_t *)(ws + o_ne);
        ra.list_a = (int32_t *)(ws + o_lista);
        ra.list_b = (int32_t *)(ws + o_listb);
        ra.n_a = d_route;
        ra.n_b = d_route + 1;
        ra.items = (uint2 *)(ws + o_items);
        ra.n_items = d_route + 2;
        ra.items_cap = (unsigned int)std::min<size_t>(items_cap, 0xFFFFFFF0u);
        route_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(ra);
        BB25_LAUNCH_CHECK();
        copy_u32_kernel<<<1, 1, 0, st>>>(d_route + 2, d_route + 3);
        BB25_LAUNCH_CHECK();
        blk_list = ra.list_b;
        blk_n_ptr = d_route + 1;
        ca.data = idx->data;
        ca.indices = idx->indices;
        ca.indptr = idx->indptr;
        ca.tab = BlockTable{idx->tab_ent, idx->tab_bits, idx->tab_row};
        ca.n_vocab = idx->n_vocab;
        ca.dense_slot = idx->dense_slot;
        ca.dense_vals = idx->dense_vals;
        ca.dense_stride = idx->dense_stride;
        ca.q_terms = d_terms;
        ca.q_off = d_qo;
        ca.term_base = 0;
        ca.ne_mask = ra.ne_mask;
        ca.items = ra.items;
        ca.n_items = d_route + 2;
        ca.thr = d_thr;
        ca.cand_cnt = d_cnt;
        ca.cand_key = d_keys;
        ca.cap = cap;
        ca.gmax = gmax;
        const unsigned cand_grid = (unsigned)idx->sm_count * 8;
        unsigned int *rc = d_round;  // stage 0
        for (int r = 0; r <= n_rounds; r++) {
            // round 0: every routed query; round r: the queries whose row overflowed in round r-1
            const int32_t *list = r == 0 ? ra.list_a : d_list[(r - 1) & 1];
            const unsigned int *n_list = r == 0 ? d_route : rc + r;
            if (r > 0) {
                BB25_CUDA(cudaMemsetAsync(ra.n_items, 0, sizeof(unsigned int), st));
                rebuild_items_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(ra, list, n_list);
                BB25_LAUNCH_CHECK();
            }
            const int pair = timed_begin();
            cand_kernel<<<cand_grid, 256, 0, st>>>(ca);
            BB25_LAUNCH_CHECK();
            timed_end(pair);
            idx->st_passes++;
            sa.q_list = list;
            sa.n_list = 0;
            sa.n_list_ptr = n_list;
            sa.final_pass = 1;
            sa.rescore = 0;
            sa.over_list = d_list[r & 1];
            sa.n_over = rc + r + 1;
            select_kernel<512><<<r == 0 ? (unsigned)n_q : repair_grid, 512, sel_smem, st>>>(sa);
            BB25_LAUNCH_CHECK();
        }
        mark_bad_kernel<<<4, 256, 0, st>>>(d_list[n_rounds & 1], rc + n_rounds + 1, d_bad);
        BB25_LAUNCH_CHECK();
    }

    // ---- block / tile traversal of the remaining queries, group by group -----------
    for (int gi = 0; gi < ng; gi++) {
        unsigned int *rc = d_round + (size_t)(gi + 1) * kRoundStride;
        if (idx->exchange_cb && gi + 1 < ng)
            BB25_CUDA(cudaMemsetAsync(d_quant, 0, sizeof(unsigned long long) * (size_t)n_quant * (size_t)n_q, st));
        for (int r = 0; r <= n_rounds; r++) {
            const int32_t *list = r == 0 ? blk_list : d_list[(r - 1) & 1];
            const unsigned int *n_list = r == 0 ? blk_n_ptr : rc + r;
            const bool first = r == 0;
            const int pair = timed_begin();
            if (blockk) {
                ba.q_list = list;
                ba.n_q = (int)n_q;
                ba.n_q_ptr = n_list;
                ba.blk_begin = bounds[gi];
                ba.blk_end = bounds[gi + 1];
                if (launch_block(idx, ba, !(relaxed && first), st)) return 1;
            } else {
                ta.q_list = list;
                ta.n_q = (int)n_q;
                ta.n_q_ptr = n_list;
                ta.tile_begin = bounds[gi];
                ta.tile_end = bounds[gi + 1];
                if (launch_tile<MODE_RETRIEVE>(idx, ta, st)) return 1;
            }
            timed_end(pair);
            idx->st_passes++;
            sa.q_list = list;
            sa.n_list = (int)n_q;
            sa.n_list_ptr = n_list;
            sa.rescore = (blockk && relaxed && first) ? 1 : 0;
            sa.final_pass = (gi == ng - 1) ? 1 : 0;
            sa.over_list = d_list[r & 1];
            sa.n_over = rc + r + 1;
            select_kernel<512><<<first ? (unsigned)n_q : repair_grid, 512, sel_smem, st>>>(sa);
            BB25_LAUNCH_CHECK();
        }
        mark_bad_kernel<<<4, 256, 0, st>>>(d_list[n_rounds & 1], rc + n_rounds + 1, d_bad);
        BB25_LAUNCH_CHECK();
        if (gi + 1 < ng && idx->exchange_cb) {
            // sharded retrieval: the ranks agree on tighter thresholds between block groups
            publish_thr_kernel<<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(d_thr, n_q, n_quant, d_quant);
            BB25_LAUNCH_CHECK();
            if (idx->exchange_cb(idx->exchange_user, (void *)d_quant, (void *)d_thr, n_q, n_quant, k, gi, (void *)st)) {
                set_error("threshold exchange callback failed");
                return 1;
            }
        }
    }
    compact_bad_kernel<<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(d_bad, n_q, d_badlist, d_nbad, d_cnt, d_prev);
    BB25_LAUNCH_CHECK();
    report_kernel<<<1, 1, 0, st>>>(d_ncand, d_stats, d_route, d_nbad, d_err, d_round, (ng + 1) * kRoundStride, d_report);
    BB25_LAUNCH_CHECK();
    BatchReport *h_rep = (BatchReport *)idx->pinned;
    BB25_CUDA(cudaMemcpyAsync(h_rep, d_report, sizeof(BatchReport), cudaMemcpyDeviceToHost, st));
    BB25_CUDA(cudaStreamSynchronize(st));
    idx->st_syncs++;
    if (h_rep->err) {
        set_error("invalid query batch (flags=%d: 1 term id out of range, 2 q_off not monotone / outside the batch)", h_rep->err);
        return 1;
    }
    idx->st_candidates = (int64_t)h_rep->n_cand;
    idx->st_units = (int64_t)h_rep->units;
    idx->st_units_skipped = (int64_t)h_rep->units_skipped;
    idx->st_units_maxscore = (int64_t)h_rep->units_maxscore;
    idx->st_units_sparse = (int64_t)h_rep->units_sparse;
    idx->st_routed = (int64_t)h_rep->route[0];
    idx->st_cand_items = (int64_t)h_rep->route[2];
    idx->st_reruns = (int64_t)h_rep->reruns;

    // ---- rare: queries whose rows still overflowed after the enqueued repair rounds ----------
    // They are evaluated again over ALL blocks against their (by now tight) thresholds in query
    // order, host-driven; whatever is left after a few such passes takes the dense guaranteed path.
    unsigned int n_bad = h_rep->n_bad;
    idx->st_bad = (int64_t)n_bad;
    if (n_bad > 0) {
        const int32_t *list = d_badlist;
        unsigned int *d_cntr = d_round;  // counters are free again
        unsigned int *h_n = (unsigned int *)((unsigned char *)idx->pinned + 512);
        int flip = 0;
        for (int iter = 0; n_bad > 0 && iter < 6; iter++) {
            BB25_CUDA(cudaMemsetAsync(d_cntr, 0, sizeof(unsigned int), st));
            if (blockk) {
                ba.q_list = list;
                ba.n_q = (int)n_bad;
                ba.n_q_ptr = nullptr;
                ba.blk_begin = 0;
                ba.blk_end = T;
                if (launch_block(idx, ba, true, st)) return 1;
            } else {
                ta.q_list = list;
                ta.n_q = (int)n_bad;
                ta.n_q_ptr = nullptr;
                ta.tile_begin = 0;
                ta.tile_end = T;
                if (launch_tile<MODE_RETRIEVE>(idx, ta, st)) return 1;
            }
            idx->st_passes++;
            sa.q_list = list;
            sa.n_list = (int)n_bad;
            sa.n_list_ptr = nullptr;
            sa.rescore = 0;
            sa.final_pass = 1;
            sa.over_list = d_list[flip];
            sa.n_over = d_cntr;
            select_kernel<512><<<n_bad, 512, sel_smem, st>>>(sa);
            BB25_LAUNCH_CHECK();
            BB25_CUDA(cudaMemcpyAsync(h_n, d_cntr, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            BB25_CUDA(cudaStreamSynchronize(st));
            idx->st_syncs++;
            n_bad = *h_n;
            list = d_list[flip];
            flip ^= 1;
        }
        if (n_bad > 0) {
            // dense guaranteed path, one query at a time (terms come from the sanitised device copy)
            std::vector<int32_t> h_list(n_bad);
            BB25_CUDA(cudaMemcpyAsync(h_list.data(), list, sizeof(int32_t) * n_bad, cudaMemcpyDeviceToHost, st));
            std::vector<int64_t> h_qo((size_t)n_q + 1);
            BB25_CUDA(cudaMemcpyAsync(h_qo.data(), d_qo, sizeof(int64_t) * (size_t)(n_q + 1), cudaMemcpyDeviceToHost, st));
            BB25_CUDA(cudaStreamSynchronize(st));
            idx->st_syncs++;
            if (k > kMaxDenseK) { set_error("threshold refinement did not converge"); return 1; }
            // the dense pass needs its own scratch and its own small workspace: the batch workspace
            // (sanitised terms) must stay alive, so both come from a temporary allocation
            unsigned char *tmp = nullptr;
            int max_m = 1;
            for (unsigned int i = 0; i < n_bad; i++) max_m = std::max<int>(max_m, (int)(h_qo[h_list[i] + 1] - h_qo[h_list[i]]));
            const size_t small = align_up(sizeof(int32_t) * (size_t)max_m) * 2 + align_up((size_t)max_m) + 5 * 256;
            BB25_CUDA(cudaMalloc(&tmp, small + dense_topk_scratch_bytes(idx, k)));
            DenseWs w;
            unsigned char *p = tmp;
            w.d_terms = (int32_t *)p; p += align_up(sizeof(int32_t) * (size_t)max_m);
            w.d_src = (int32_t *)p; p += align_up(sizeof(int32_t) * (size_t)max_m);
            w.d_nc = p; p += align_up((size_t)max_m);
            w.d_qoff = (int64_t *)p; p += 256;
            w.d_qo = (int64_t *)p; p += 256;
            w.d_ctr = (unsigned long long *)p; p += 256;
            w.d_err = (int *)p; p += 256;
            p += 256;
            int rc = 0;
            for (unsigned int i = 0; i < n_bad && !rc; i++) {
                const int q = h_list[i];
                const int m = (int)std::max<int64_t>(0, h_qo[q + 1] - h_qo[q]);
                stage_query_dev_kernel<<<1, 128, 0, st>>>(d_terms + h_qo[q], m, w.d_src, w.d_qoff);
                count_launch();
                rc = dense_topk_staged(idx, params, w, p, m, k, out_ids + (size_t)q * k,
                                       out_scores ? out_scores + (size_t)q * k : nullptr, out_probs + (size_t)q * k, st);
            }
            cudaStreamSynchronize(st);
            idx->st_syncs++;
            cudaFree(tmp);
            if (rc) return 1;
            idx->st_dense_fallback = (int64_t)n_bad;
        }
    }
    for (int i = 0; i < idx->ev_used; i++) {
        float ms = 0.f;
        BB25_CUDA(cudaEventElapsedTime(&ms, idx->ev[2 * i], idx->ev[2 * i + 1]));
        idx->st_traverse_ms += (double)ms;
    }
    idx->st_traverse_launches = idx->ev_used;
    idx->st_launches = (int64_t)bb25_launch_count() - launches0;
    return 0;
}

}  // namespace bb25

using namespace bb25;

extern "C" {

int bb25_get_scores(bb25_index *idx, const int32_t *q_terms, int n_terms, float *out_scores, void *stream) {
    if (!out_scores) { set_error("out_scores is NULL"); return 1; }
    return run_dense(idx, MODE_SCORES, nullptr, q_terms, n_terms, out_scores, nullptr, 1, (cudaStream_t)stream);
}

int bb25_get_probabilities(bb25_index *idx, const bb25_params *params, const int32_t *q_terms, int n_terms,
                           double *out_probs, int64_t out_stride, void *stream) {
    if (!out_probs || out_stride < 1) { set_error("bad output arguments"); return 1; }
    if (check_params(params)) return 1;
    return run_dense(idx, MODE_PROBS, params, q_terms, n_terms, nullptr, out_probs, out_stride, (cudaStream_t)stream);
}

int bb25_fuse_bm25_signal(bb25_index *idx, const bb25_params *params, const int32_t *q_terms, int n_terms,
                          double weight, int n_signals, double scale, int flags, double *acc, void *stream) {
    if (!acc || n_signals < 1 || (flags & ~7)) { set_error("bad fuse arguments"); return 1; }
    if (check_params(params)) return 1;
    FuseSpec f{weight, scale, n_signals, flags};
    return run_dense(idx, MODE_FUSED, params, q_terms, n_terms, nullptr, acc, 1, (cudaStream_t)stream, &f);
}

static int retrieve_checked(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                            const int64_t *q_off, int64_t n_queries, int64_t term_base, int64_t n_terms_total,
                            int k, int64_t *out_ids, float *out_scores, double *out_probs, cudaStream_t st,
                            bool read_ends, cudaEvent_t done_ev = nullptr) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (check_params(params)) return 1;
    if (n_queries < 0 || !q_off || !out_ids || !out_probs) { set_error("bad arguments"); return 1; }
    if (k < 1 || k > kMaxK || (int64_t)k > idx->n_docs) {
        set_error("k must satisfy 1 <= k <= min(n_docs, %d), got k=%d with n_docs=%lld", kMaxK, k, (long long)idx->n_docs);
        return 1;
    }
    if (n_queries == 0) return 0;
    if (n_queries > 0x7FFFFFF0ll) { set_error("too many queries in one batch"); return 1; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    int extra_syncs = 0;
    if (read_ends) {
        int64_t ends[2];
        BB25_CUDA(cudaMemcpyAsync(&ends[0], q_off, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        BB25_CUDA(cudaMemcpyAsync(&ends[1], q_off + n_queries, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        BB25_CUDA(cudaStreamSynchronize(st));
        extra_syncs = 1;
        term_base = ends[0];
        n_terms_total = ends[1] - ends[0];
    }
    if (n_terms_total < 0 || term_base < 0) { set_error("bad q_off"); return 1; }
    if (n_terms_total > 0 && !q_terms) { set_error("q_terms is NULL"); return 1; }
    ws_acquire(idx, st);
    const int rc = retrieve_device(idx, params, q_terms, q_off, n_queries, term_base, n_terms_total, k, out_ids,
                                   out_scores, out_probs, st);
    ws_release(idx, st);
    if (done_ev) cudaEventRecord(done_ev, st);  // still under the index lock: nothing of another call is ahead of it
    idx->st_syncs += extra_syncs;
    return rc;
}

int bb25_retrieve_batch(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                        const int64_t *q_off, int64_t n_queries, int k, int64_t *out_ids,
                        float *out_scores, double *out_probs, void *stream) {
    return retrieve_checked(idx, params, q_terms, q_off, n_queries, 0, 0, k, out_ids, out_scores, out_probs,
                            (cudaStream_t)stream, true);
}

int bb25_retrieve_batch_ex(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                           const int64_t *q_off, int64_t n_queries, int64_t term_base, int64_t n_terms_total,
                           int k, int64_t *out_ids, float *out_scores, double *out_probs, void *stream) {
    return retrieve_checked(idx, params, q_terms, q_off, n_queries, term_base, n_terms_total, k, out_ids, out_scores,
                            out_probs, (cudaStream_t)stream, false);
}

int bb25_retrieve_batch_host(bb25_index *idx, const bb25_params *params, const int32_t *q_terms,
                             const int64_t *q_off, int64_t n_queries, int k, int64_t *out_ids,
                             float *out_scores, double *out_probs) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (n_queries < 0 || !q_off || !out_ids || !out_probs) { set_error("bad arguments"); return 1; }
    if (n_queries == 0) return 0;
    if (k < 1) { set_error("k must be >= 1"); return 1; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    const int64_t nt = q_off[n_queries] - q_off[0];
    if (nt < 0) { set_error("bad q_off"); return 1; }
    const size_t nk = (size_t)n_queries * (size_t)k;
    // Device staging slots and streams owned by the handle, grown on demand and kept (no per-call
    // allocation); the copies are asynchronous when the caller's buffers are page-locked.  There are TWO
    // slots: the results of this call go to the host on the slot's own copy stream after the index lock
    // has been released, so a second thread's call computes (into the other slot) while they travel.
    const size_t o_terms = 0;
    const size_t o_off = align_up(o_terms + sizeof(int32_t) * (size_t)(nt > 0 ? nt : 1));
    const size_t o_ids = align_up(o_off + sizeof(int64_t) * (size_t)(n_queries + 1));
    const size_t o_pr = align_up(o_ids + sizeof(int64_t) * nk);
    const size_t o_sc = align_up(o_pr + sizeof(double) * nk);
    const size_t total = align_up(o_sc + sizeof(float) * nk);
    int slot = 0;
    {
        std::lock_guard<std::mutex> lock(idx->mu);
        if (!idx->hs_stream) BB25_CUDA(cudaStreamCreateWithFlags(&idx->hs_stream, cudaStreamNonBlocking));
        slot = idx->hs_next;
        idx->hs_next ^= 1;
    }
    std::lock_guard<std::mutex> slot_lock(idx->hs_mu[slot]);
    if (!idx->hs_copy[slot]) BB25_CUDA(cudaStreamCreateWithFlags(&idx->hs_copy[slot], cudaStreamNonBlocking));
    if (!idx->hs_ev[slot]) BB25_CUDA(cudaEventCreateWithFlags(&idx->hs_ev[slot], cudaEventDisableTiming));
    if (total > idx->hs_bytes[slot]) {
        if (idx->hs_dev[slot]) {
            BB25_CUDA(cudaFree(idx->hs_dev[slot]));  // nothing of this slot is in flight: its previous call has returned
            idx->hs_dev[slot] = nullptr;
            idx->hs_bytes[slot] = 0;
        }
        const size_t want = total + (total >> 3);
        BB25_CUDA(cudaMalloc(&idx->hs_dev[slot], want));
        idx->hs_bytes[slot] = want;
    }
    cudaStream_t st = idx->hs_stream, cp = idx->hs_copy[slot];
    unsigned char *base = (unsigned char *)idx->hs_dev[slot];
    int32_t *d_terms = (int32_t *)(base + o_terms);
    int64_t *d_off = (int64_t *)(base + o_off);
    int64_t *d_ids = (int64_t *)(base + o_ids);
    double *d_pr = (double *)(base + o_pr);
    float *d_sc = (float *)(base + o_sc);
    // the queries go up on the copy stream too (the compute stream may still be busy with another call)
    if (nt > 0) BB25_CUDA(cudaMemcpyAsync(d_terms, q_terms + q_off[0], sizeof(int32_t) * (size_t)nt, cudaMemcpyHostToDevice, cp));
    BB25_CUDA(cudaMemcpyAsync(d_off, q_off, sizeof(int64_t) * (size_t)(n_queries + 1), cudaMemcpyHostToDevice, cp));
    BB25_CUDA(cudaStreamSynchronize(cp));
    // d_terms holds positions [q_off[0], q_off[Q]): the batch's first term is element 0 of d_terms.  retrieve_checked
    // takes the index lock and returns with the compute stream drained (it reads the batch report).
    if (retrieve_checked(idx, params, d_terms - q_off[0], d_off, n_queries, q_off[0], nt, k, d_ids, d_sc, d_pr, st, false,
                         idx->hs_ev[slot]))
        return 1;
    BB25_CUDA(cudaStreamWaitEvent(cp, idx->hs_ev[slot], 0));  // whatever the batch still had on the compute stream
    BB25_CUDA(cudaMemcpyAsync(out_ids, d_ids, sizeof(int64_t) * nk, cudaMemcpyDeviceToHost, cp));
    if (out_scores) BB25_CUDA(cudaMemcpyAsync(out_scores, d_sc, sizeof(float) * nk, cudaMemcpyDeviceToHost, cp));
    BB25_CUDA(cudaMemcpyAsync(out_probs, d_pr, sizeof(double) * nk, cudaMemcpyDeviceToHost, cp));
    BB25_CUDA(cudaStreamSynchronize(cp));
    {
        std::lock_guard<std::mutex> lock(idx->mu);
        idx->st_syncs++;
    }
    return 0;
}

int bb25_retrieve_one_dense(bb25_index *idx, const bb25_params *params, const int32_t *q_terms, int n_terms, int k,
                            int64_t *out_ids, float *out_scores, double *out_probs, void *stream) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (check_params(params)) return 1;
    if (n_terms < 0 || (n_terms > 0 && !q_terms) || !out_ids || !out_probs) { set_error("bad arguments"); return 1; }
    if (k < 1 || k > kMaxDenseK || (int64_t)k > idx->n_docs) {
        set_error("k must satisfy 1 <= k <= min(n_docs, %d), got %d", kMaxDenseK, k);
        return 1;
    }
    for (int i = 0; i < n_terms; i++)
        if (q_terms[i] < 0 || q_terms[i] >= idx->n_vocab) {
            set_error("query term id %d out of range [0, %lld)", q_terms[i], (long long)idx->n_vocab);
            return 1;
        }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    cudaStream_t st = (cudaStream_t)stream;
    DenseWs w;
    unsigned char *scratch = nullptr;
    ws_acquire(idx, st);
    if (dense_workspace(idx, n_terms, dense_topk_scratch_bytes(idx, k), w, &scratch)) return 1;
    if (n_terms > 0) {
        int64_t hq[2] = {0, n_terms};
        BB25_CUDA(cudaMemcpyAsync(w.d_src, q_terms, sizeof(int32_t) * (size_t)n_terms, cudaMemcpyHostToDevice, st));
        BB25_CUDA(cudaMemcpyAsync(w.d_qoff, hq, sizeof(hq), cudaMemcpyHostToDevice, st));
        BB25_CUDA(cudaStreamSynchronize(st));
    }
    const int rc = dense_topk_staged(idx, params, w, scratch, n_terms, k, out_ids, out_scores, out_probs, st);
    ws_release(idx, st);
    return rc;
}

int bb25_index_set_threshold_exchange(bb25_index *idx, bb25_exchange_fn fn, void *user, int n_shards) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (fn && (n_shards < 2 || n_shards > 32)) { set_error("n_shards must be in [2, 32]"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->exchange_cb = fn;
    idx->exchange_user = user;
    idx->exchange_shards = fn ? n_shards : 0;
    return 0;
}

int bb25_retrieve_sync_stats(const bb25_index *idx, int64_t *host_syncs, int64_t *repaired_queries,
                             int64_t *dense_fallback_queries) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (host_syncs) *host_syncs = idx->st_syncs;
    if (repaired_queries) *repaired_queries = idx->st_bad;
    if (dense_fallback_queries) *dense_fallback_queries = idx->st_dense_fallback;
    return 0;
}

int bb25_retrieve_stats(const bb25_index *idx, int64_t *launches, int64_t *passes, int64_t *rerun_queries,
                        int64_t *candidates) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (launches) *launches = idx->st_launches;
    if (passes) *passes = idx->st_passes;
    if (rerun_queries) *rerun_queries = idx->st_reruns;
    if (candidates) *candidates = idx->st_candidates;
    return 0;
}

int bb25_retrieve_prune_stats(const bb25_index *idx, int64_t *units, int64_t *units_skipped,
                              int64_t *units_maxscore) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (units) *units = idx->st_units;
    if (units_skipped) *units_skipped = idx->st_units_skipped;
    if (units_maxscore) *units_maxscore = idx->st_units_maxscore;
    return 0;
}

int bb25_retrieve_sparse_units(const bb25_index *idx, int64_t *units_sparse) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (units_sparse) *units_sparse = idx->st_units_sparse;
    return 0;
}

int bb25_retrieve_route_stats(const bb25_index *idx, int64_t *routed_queries, int64_t *work_items) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (routed_queries) *routed_queries = idx->st_routed;
    if (work_items) *work_items = idx->st_cand_items;
    return 0;
}

int bb25_index_set_pruning(bb25_index *idx, int enable) {
    if (!idx) { set_error("index is NULL"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->prune = enable < 0 ? 0 : (enable > 3 ? 3 : enable);
    return 0;
}

int bb25_retrieve_timing(const bb25_index *idx, double *traverse_ms, int64_t *traverse_launches) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (traverse_ms) *traverse_ms = idx->st_traverse_ms;
    if (traverse_launches) *traverse_launches = idx->st_traverse_launches;
    return 0;
}

}  // extern "C"
