// Internal declarations shared by the libbb25.so translation units.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>

#include "../../include/bb25.h"

namespace bb25 {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define BB25_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            bb25::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                            __FILE__, __LINE__);                                     \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define BB25_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        bb25::count_launch();                                                        \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) {                                                     \
            bb25::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                     \
            return 1;                                                                \
        }                                                                            \
    } while (0)

// RAII device switch
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; }
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr double kEps = 1e-10;  // probability.py:20
constexpr int kBlockDocs = 1024;  // documents per warp-private block
constexpr uint32_t kBlkLenMask = 0x7FFu;
constexpr int kMaxDenseTerms = 64;

// ---- probability.py / fusion.py scalar math, fp64 -------------------------------
__host__ __device__ inline double clamp_prob(double p) {  // probability.py:24-26
    return p < kEps ? kEps : (p > 1.0 - kEps ? 1.0 - kEps : p);
}
__device__ inline double d_sigmoid(double x) {  // probability.py:29-41 (split form)
    if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
    double e = exp(x);
    return e / (1.0 + e);
}
__device__ inline double d_logit(double p) {  // probability.py:44-48
    p = clamp_prob(p);
    return log(p / (1.0 - p));
}
__device__ inline double d_tf_prior(double tf) {  // probability.py:110-115
    return 0.2 + 0.7 * fmin(1.0, tf / 10.0);
}
__device__ inline double d_norm_prior(double r) {  // probability.py:117-129
    return 0.3 + 0.6 * (1.0 - fmin(1.0, fabs(r - 0.5) * 2.0));
}
__device__ inline double d_composite_prior(double tf, double r) {  // probability.py:131-140
    double v = 0.7 * d_tf_prior(tf) + 0.3 * d_norm_prior(r);
    return v < 0.1 ? 0.1 : (v > 0.9 ? 0.9 : v);
}
__device__ inline double d_posterior(double l, double p, int has_br, double br) {  // :142-169
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
// scorer.py:618-638 for one document: 0.0 unless score > 0
__device__ inline double d_doc_probability(const bb25_params &p, float score, int tf, int doc_len,
                                           double avgdl) {
    if (!(score > 0.0f)) return 0.0;
    double l = d_sigmoid(p.alpha * ((double)score - p.beta));
    double prior = p.prior_mode == 1 ? 0.5 : d_composite_prior((double)tf, (double)doc_len / avgdl);
    return d_posterior(l, prior, p.has_base_rate, p.base_rate);
}

// ---- one signal of a log-odds conjunction, accumulated in place (fusion.py:243-280) ----
// flags: 1 = first signal (ignore the previous contents), 2 = last signal (apply the
// n**alpha scale and the sigmoid), 4 = unweighted mean branch (fusion.py:270-279)
struct FuseSpec {
    double weight;
    double scale;  // n ** alpha, resolved by the caller
    int n_signals;
    int flags;
};
__device__ inline double fuse_step(double prev, double x, const FuseSpec &f) {
    const double term = (f.flags & 4) ? x : f.weight * x;
    const double acc = (f.flags & 1) ? (0.0 + term) : (prev + term);  // left-to-right, as NumPy sums the last axis
    if (f.flags & 2) {
        const double l = (f.flags & 4) ? (acc / (double)f.n_signals) * f.scale : f.scale * acc;
        return d_sigmoid(l);
    }
    return acc;
}

// ---- candidate key: score desc, then local doc id asc ----------------------------
// [63:33] fp32 score bits (scores are >= 0 so bit 31 is clear)
// [32:4]  0x1FFFFFFF - local doc id   (n_docs <= 2^29 per index)
// [3:0]   matched-term count, saturated at 15 (the tf prior saturates at 10)
constexpr uint32_t kIdMask = 0x1FFFFFFFu;
__host__ __device__ inline unsigned long long make_key(uint32_t score_bits, uint32_t local_id,
                                                       uint32_t tf) {
    return ((unsigned long long)score_bits << 33) |
           ((unsigned long long)(kIdMask - local_id) << 4) | (unsigned long long)(tf > 15u ? 15u : tf);
}
__host__ __device__ inline uint32_t key_score_bits(unsigned long long k) { return (uint32_t)(k >> 33); }
__host__ __device__ inline uint32_t key_local_id(unsigned long long k) {
    return kIdMask - (uint32_t)((k >> 4) & kIdMask);
}
__host__ __device__ inline uint32_t key_tf(unsigned long long k) { return (uint32_t)(k & 15ull); }

// k-th largest of n distinct 64-bit keys held in shared memory (n >= k >= 1): MSB-first
// radix select, 8 passes of 8 bits, one 256-bin histogram per pass.  All threads of the
// block must call it; hist = 256 words, st = 2 words of shared scratch.  With passes < 8
// only the top 8*passes bits of the k-th largest key are resolved (the rest are 0).
template <int NT>
__device__ unsigned long long block_kth_largest(const unsigned long long *keys, int n, int k, unsigned int *hist,
                                                unsigned int *st, int tid, int passes = 8) {
    unsigned long long prefix = 0ull, mask = 0ull;
    unsigned int rem = (unsigned int)k;
    for (int pass = 0; pass < passes; pass++) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += NT) hist[i] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const unsigned long long key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // lane L owns bins 255-8L .. 248-8L (descending); scan the lane sums, then the
            // owning lane walks its 8 bins
            unsigned int loc[8];
            unsigned int sum = 0u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                loc[j] = hist[255 - (8 * tid + j)];
                sum += loc[j];
            }
            unsigned int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (tid >= d) incl += y;
            }
            const unsigned int excl = incl - sum;
            if (excl < rem && rem <= incl) {
                unsigned int c = excl;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (c + loc[j] >= rem) {
                        st[0] = (unsigned int)(255 - (8 * tid + j));
                        st[1] = rem - c;
                        break;
                    }
                    c += loc[j];
                }
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)st[0] << shift;
        mask |= 255ull << shift;
        rem = st[1];
        // the next pass (or the caller) synchronises before st / hist are written again
    }
    return prefix;
}

}  // namespace bb25

// Block table: for every (term t, 1024-document block b) that holds postings of t, one entry
//   .x = offset of the term's first posting in the block, relative to indptr[t]
//   .y = (fp32 bits of the block maximum, rounded UP to a multiple of 2^11) | posting count (<= 1024)
// (skip pointer and BlockMaxIndex bound in one 8-byte load; empty pairs read {0xFFFFFFFF, 0}).
// Term-major.  row[t] = (entry offset, bitmap offset):
//   bitmap offset < 0   dense row: ent[entry offset + b], one load;
//   otherwise           the term touches few blocks: bits[bitmap offset + b/32] = (32-block
//                       bitmap word, number of the term's entries before this word) and the
//                       entry, if the bit is set, sits at ent[entry offset + prefix + rank in word].
// Dense rows cost n_blocks x 8 B per term whatever its df, so they are kept only while the whole
// table fits a budget (every BASELINE configuration); beyond it (million-term vocabularies) the
// rare terms take the bitmap form: 0.25 B per block plus 8 B per block actually touched.
struct BlockTable {
    const uint2 *ent;
    const uint2 *bits;
    const longlong2 *row;
};

__device__ __forceinline__ uint2 tab_lookup(const BlockTable &tb, longlong2 r, int blk) {
    if (r.y < 0) return tb.ent[r.x + blk];
    const uint2 wb = tb.bits[r.y + (blk >> 5)];
    const unsigned bit = 1u << (blk & 31);
    if (!(wb.x & bit)) return make_uint2(0xFFFFFFFFu, 0u);
    return tb.ent[r.x + wb.y + __popc(wb.x & (bit - 1u))];
}

struct bb25_index {
    int device = 0;
    int sm_count = 0;
    int64_t n_docs = 0, n_vocab = 0, nnz = 0, doc_id_offset = 0;
    double avgdl = 0.0;
    float *data = nullptr;
    int32_t *indices = nullptr;
    int64_t *indptr = nullptr;
    int32_t *doc_len = nullptr;
    int tile_docs = 0;  // docs per traversal tile (shared-memory accumulator span)
    int n_tiles = 0;
    uint32_t *tile_off = nullptr;  // [n_vocab][n_tiles+1] offsets relative to indptr[t]
    // block table for the warp-private traversal (struct BlockTable below), term-major
    uint2 *tab_ent = nullptr;
    uint2 *tab_bits = nullptr;
    longlong2 *tab_row = nullptr;
    int64_t tab_sparse_terms = 0;  // terms kept as bitmap + compact entries
    size_t tab_bytes = 0;
    int n_blocks = 0;
    int prune = 3;  // 0 exhaustive, 1 block-max skip, 2 + frequent-term bound per block, 3 + candidate-driven queries
    // dense value rows for the most frequent terms (df >= n_docs/8, at most kMaxDenseTerms):
    // dense_vals[slot][doc] = posting value, or -0.0f where absent; O(1) lookup of a head term's
    // contribution to one document (exact re-scoring, candidate path, tf recovery), and the rows the
    // order-free traversal pass streams 128 bits per lane
    int32_t *dense_slot = nullptr;  // [n_vocab] slot or -1
    float *dense_vals = nullptr;    // [n_dense][dense_stride]
    __half *dense_h = nullptr;      // the same rows as fp16 UPPER BOUNDS (rounded up, absent = 0): order-free pass only
    int n_dense = 0;
    int64_t dense_stride = 0;       // n_blocks * kBlockDocs
    // LOOKUP rows (built on the first fused batch that can use them, ensure_lookup_rows): the same fp32 value
    // rows for the mid-frequency terms (df >= n_docs / BB25_LOOKUP_DIV, within a memory budget).  No traversal
    // pass streams them; they make "value of term t in document d" one 4-byte load where a document is
    // evaluated on its own (essential-posting evaluation of the fused batch, exact evaluation of candidates).
    // row_slot[t]: the hot slot (< n_dense), n_dense + lookup row, or -1.
    float *lookup_vals = nullptr;   // [n_lookup][dense_stride]
    int32_t *row_slot = nullptr;    // [n_vocab]
    int n_lookup = 0;
    bool lookup_tried = false;
    std::map<int, float *> kth_cache;  // k -> fp32[n_vocab] k-th largest posting value per term
    // grow-only device workspace shared by query calls (serialised by mu)
    std::mutex mu;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *pinned = nullptr;  // small pinned host scratch
    size_t device_bytes = 0;
    // stats of the last retrieve_batch
    int64_t st_launches = 0, st_passes = 0, st_reruns = 0, st_candidates = 0;
    int64_t st_routed = 0, st_cand_items = 0;  // queries evaluated candidate-by-candidate / their work items
    int64_t st_units = 0, st_units_skipped = 0, st_units_maxscore = 0, st_units_sparse = 0;  // (block, query) units visited / pruned / evaluated with the level-2 restriction
    // CUDA-event pairs around the traversal launches of the last retrieve_batch
    static constexpr int kMaxEv = 256;
    cudaEvent_t ev[2 * kMaxEv] = {};
    int n_ev = 0;       // events created so far
    int ev_used = 0;    // pairs recorded in the last call
    double st_traverse_ms = 0.0;
    int64_t st_traverse_launches = 0;
    int64_t st_syncs = 0;           // host synchronisations inside the last retrieve_batch
    int64_t st_bad = 0;             // queries that needed the host-driven repair / ...
    int64_t st_dense_fallback = 0;  // ... the dense guaranteed path
    // stats of the last bb25_retrieve_fused_batch whose first field is this index
    int64_t fz_units = 0, fz_skipped = 0, fz_abandoned = 0, fz_candidates = 0, fz_fallback = 0, fz_reruns = 0, fz_syncs = 0;
    int64_t fz_ne_skipped = 0, fz_sparse_units = 0, fz_sparse_docs = 0;
    double fz_traverse_ms = 0.0;
    // device + stream of the host-buffer entry points (grow-only, reused across calls)
    // two staging slots: while one call's results travel to the host on the slot's copy stream, the next
    // call (another thread) already computes into the other slot
    void *hs_dev[2] = {nullptr, nullptr};
    size_t hs_bytes[2] = {0, 0};
    cudaStream_t hs_copy[2] = {nullptr, nullptr};
    cudaEvent_t hs_ev[2] = {nullptr, nullptr};
    std::mutex hs_mu[2];  // a slot serves one call at a time
    int hs_next = 0;
    cudaStream_t hs_stream = nullptr;
    // last use of the shared workspace: calls on another stream wait for it before touching the workspace
    cudaEvent_t ws_ev = nullptr;
    // sharded retrieval: called between block groups so that the ranks can agree on tighter thresholds
    bb25_exchange_fn exchange_cb = nullptr;
    void *exchange_user = nullptr;
    int exchange_shards = 0;
};

namespace bb25 {
// stream-order the shared workspace across streams: acquire before the first use in a call, release after the last
inline void ws_acquire(bb25_index *idx, cudaStream_t st) {
    if (idx->ws_ev) cudaStreamWaitEvent(st, idx->ws_ev, 0);
}
inline void ws_release(bb25_index *idx, cudaStream_t st) {
    if (!idx->ws_ev && cudaEventCreateWithFlags(&idx->ws_ev, cudaEventDisableTiming) != cudaSuccess) {
        idx->ws_ev = nullptr;
        return;
    }
    cudaEventRecord(idx->ws_ev, st);
}
int ensure_workspace(bb25_index *idx, size_t bytes);
int prep_queries_launch(const bb25_index *idx, const int32_t *q_terms, const int64_t *q_off, int64_t n_q,
                        int64_t term_base, int64_t n_terms_total, int32_t *qt_ws, uint8_t *nocount, int64_t *qo_ws,
                        longlong2 *qt_info, int *err, cudaStream_t st);
int get_kth_values(bb25_index *idx, int k, cudaStream_t st, const float **out);
int ensure_tile_table(bb25_index *idx, cudaStream_t st);
int ensure_lookup_rows(bb25_index *idx, cudaStream_t st);
}  // namespace bb25
