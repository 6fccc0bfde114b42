// Query-time consumers of the hot path's probabilities, device-resident:
//  * bb25_match_counts + bb25_trace_bm25: what retrieve(explain=True) records per returned document
//    (scorer.py:538-562, debug.py:178-216) -- tf, every intermediate of the posterior;
//  * bb25_attention_weights + bb25_attention_fuse: AttentionLogOddsWeights inference
//    (fusion.py:774-828 __call__, :1039-1082 compute_upper_bounds): softmax(W q + b) per query, per-column
//    min-max normalisation of the logits (optional), scale * sum_i w_i logit(p_i) (+ logit base rate), sigmoid;
//  * bb25_balanced_fusion: balanced_log_odds_fusion (fusion.py:283-343): both signals' logits min-max
//    normalised over the candidate set, then weight * dense + (1 - weight) * sparse.
#include "bb25_internal.cuh"
#include "bb25_device.cuh"

namespace bb25 {

// ---- tf of returned documents -------------------------------------------------------------
__global__ void __launch_bounds__(256) match_counts_kernel(const int32_t *__restrict__ q_terms,
                                                           const int64_t *__restrict__ q_off, int64_t n_q, int k,
                                                           const int64_t *__restrict__ ids, int64_t doc_id_offset,
                                                           int64_t n_docs, int64_t n_vocab,
                                                           const int32_t *__restrict__ indices,
                                                           const int64_t *__restrict__ indptr, BlockTable tab,
                                                           const int32_t *__restrict__ dense_slot,
                                                           const float *__restrict__ dense_vals, int64_t dense_stride,
                                                           int32_t *__restrict__ out_tf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q * k) return;
    const int64_t q = i / k;
    const int64_t d = ids[i] - doc_id_offset;
    int c = 0;
    if (d >= 0 && d < n_docs) {
        const uint32_t doc = (uint32_t)d;
        const int blk = (int)(doc / (uint32_t)kBlockDocs);
        const int64_t t0 = q_off[q], t1 = q_off[q + 1];
        for (int64_t a = t0; a < t1; a++) {
            const int t = q_terms[a];
            if (t < 0 || (int64_t)t >= n_vocab) continue;
            bool dup = false;  // distinct query terms only (scorer.py:592-601: a set intersection)
            for (int64_t b = t0; b < a && !dup; b++) dup = q_terms[b] == t;
            if (dup) continue;
            const int slot = dense_slot ? dense_slot[t] : -1;
            if (slot >= 0 && dense_vals) {
                c += (__float_as_uint(dense_vals[(size_t)slot * (size_t)dense_stride + doc]) != 0x80000000u);
                continue;
            }
            const uint2 ent = tab_lookup(tab, tab.row[t], blk);
            const int len = (int)(ent.y & kBlkLenMask);
            if (len == 0) continue;
            if (slice_find(indices + indptr[t] + (long long)ent.x, len, doc, doc & ~(uint32_t)(kBlockDocs - 1)) >= 0) c++;
        }
    }
    out_tf[i] = c;
}

// ---- debug.py:178-216 for n (score, tf, ratio) triples; out[i] = 7 doubles ------------------
__global__ void __launch_bounds__(256) trace_bm25_kernel(bb25_params p, const double *__restrict__ score,
                                                         const double *__restrict__ tf, const double *__restrict__ ratio,
                                                         int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double l = d_sigmoid(p.alpha * (score[i] - p.beta));
    const double ptf = d_tf_prior(tf[i]);
    const double pn = d_norm_prior(ratio[i]);
    const double pc = d_composite_prior(tf[i], ratio[i]);
    double *o = out + 7 * i;
    o[0] = l;
    o[1] = ptf;
    o[2] = pn;
    o[3] = pc;
    o[4] = d_logit(l);
    o[5] = d_logit(pc);
    o[6] = d_posterior(l, pc, p.has_base_rate, p.base_rate);
}

// ---- attention weights: softmax(qf @ W^T + b), one thread per query row ------------------------
__global__ void __launch_bounds__(128) attention_weights_kernel(const double *__restrict__ qf, const double *__restrict__ W,
                                                                const double *__restrict__ b, int64_t m, int nf, int n,
                                                                double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double *o = out + r * n;
    double mx = -1.0e308;
    for (int s = 0; s < n; s++) {
        double z = 0.0;
        for (int f = 0; f < nf; f++) z += qf[r * nf + f] * W[(int64_t)s * nf + f];
        z += b[s];
        o[s] = z;
        mx = z > mx ? z : mx;
    }
    double sum = 0.0;
    for (int s = 0; s < n; s++) {
        const double e = exp(o[s] - mx);
        o[s] = e;
        sum += e;
    }
    for (int s = 0; s < n; s++) o[s] = o[s] / sum;
}

// ---- column-wise / whole-array min and max of logits -------------------------------------------
// mode 0: x = logit(clamp(p[i*n + c])) over rows i, one result pair per column c (grid.y = column)
// mode 1: x = logit(cosine_to_probability(v[i]))  (n == 1)
__global__ void __launch_bounds__(256) logit_minmax_kernel(const double *__restrict__ v, int64_t m, int n, int mode,
                                                           double *__restrict__ part /*[cols][grid.x][2]*/) {
    __shared__ double s_lo[256], s_hi[256];
    const int c = blockIdx.y;
    double lo = 1.0e308, hi = -1.0e308;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double raw = v[i * n + c];
        const double x = d_logit(mode == 1 ? clamp_prob((1.0 + raw) / 2.0) : raw);
        lo = x < lo ? x : lo;
        hi = x > hi ? x : hi;
    }
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            s_lo[threadIdx.x] = fmin(s_lo[threadIdx.x], s_lo[threadIdx.x + d]);
            s_hi[threadIdx.x] = fmax(s_hi[threadIdx.x], s_hi[threadIdx.x + d]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[((int64_t)c * gridDim.x + blockIdx.x) * 2] = s_lo[0];
        part[((int64_t)c * gridDim.x + blockIdx.x) * 2 + 1] = s_hi[0];
    }
}
__global__ void minmax_final_kernel(const double *__restrict__ part, int n_part, double *__restrict__ out /*[cols][2]*/) {
    const int c = blockIdx.x;
    if (threadIdx.x != 0) return;
    double lo = 1.0e308, hi = -1.0e308;
    for (int i = 0; i < n_part; i++) {
        lo = fmin(lo, part[((int64_t)c * n_part + i) * 2]);
        hi = fmax(hi, part[((int64_t)c * n_part + i) * 2 + 1]);
    }
    out[2 * c] = lo;
    out[2 * c + 1] = hi;
}
__device__ inline double d_minmax_norm(double x, double lo, double hi) {  // fusion.py:336-343
    return (hi - lo < 1e-12) ? 0.0 : (x - lo) / (hi - lo);
}

// sigma(scale * sum_i w[r or 0][i] * x_i (+ logit_br)); x = logit(clamp(p)), optionally min-max normalised per column
__global__ void __launch_bounds__(256) attention_fuse_kernel(const double *__restrict__ probs, int64_t m, int n,
                                                             const double *__restrict__ w, int64_t mw, double scale,
                                                             int has_br, double logit_br,
                                                             const double *__restrict__ colmm /*[n][2] or NULL*/,
                                                             double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double *wr = w + (mw > 1 ? r * n : 0);
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        double x = d_logit(probs[r * n + i]);
        if (colmm) x = d_minmax_norm(x, colmm[2 * i], colmm[2 * i + 1]);
        acc += wr[i] * x;
    }
    double l = scale * acc;
    if (has_br) l = l + logit_br;
    out[r] = d_sigmoid(l);
}

__global__ void __launch_bounds__(256) balanced_fusion_kernel(const double *__restrict__ sparse, const double *__restrict__ dense,
                                                              int64_t n, double weight, const double *__restrict__ mm /*[2][2]*/,
                                                              double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xs = d_minmax_norm(d_logit(sparse[i]), mm[0], mm[1]);
    const double xd = d_minmax_norm(d_logit(clamp_prob((1.0 + dense[i]) / 2.0)), mm[2], mm[3]);
    out[i] = weight * xd + (1.0 - weight) * xs;
}

static int column_minmax(const double *v, int64_t m, int n, int mode, double *out_dev, cudaStream_t st) {
    const int gx = (int)((m + 255) / 256 < 592 ? (m + 255) / 256 : 592);
    double *part = nullptr;
    BB25_CUDA(cudaMallocAsync(&part, sizeof(double) * 2 * (size_t)gx * (size_t)n, st));
    logit_minmax_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, st>>>(v, m, n, mode, part);
    minmax_final_kernel<<<n, 32, 0, st>>>(part, gx, out_dev);
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(part, st);
    if (e != cudaSuccess) { set_error("min/max kernels failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

}  // namespace bb25

using namespace bb25;

extern "C" {

int bb25_match_counts(bb25_index *idx, const int32_t *q_terms, const int64_t *q_off, int64_t n_queries, int k,
                      const int64_t *ids, int32_t *out_tf, void *stream) {
    if (!idx || !q_off || !ids || !out_tf || n_queries < 0 || k < 1) { set_error("bad arguments"); return 1; }
    if (n_queries == 0) return 0;
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    const int64_t n = n_queries * k;
    match_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        q_terms, q_off, n_queries, k, ids, idx->doc_id_offset, idx->n_docs, idx->n_vocab, idx->indices, idx->indptr,
        BlockTable{idx->tab_ent, idx->tab_bits, idx->tab_row}, idx->dense_vals ? idx->dense_slot : nullptr, idx->dense_vals,
        idx->dense_stride, out_tf);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_trace_bm25(int device, const bb25_params *p, const double *score, const double *tf, const double *ratio, int64_t n,
                    double *out, void *stream) {
    if (!p || n < 0 || (n > 0 && (!score || !tf || !ratio || !out))) { set_error("bad arguments"); return 1; }
    if (n == 0) return 0;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    trace_bm25_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, score, tf, ratio, n, out);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_attention_weights(int device, const double *query_features, const double *W, const double *b, int64_t m,
                           int n_features, int n_signals, double *out, void *stream) {
    if (m < 0 || n_features < 1 || n_signals < 1 || (m > 0 && (!query_features || !W || !b || !out))) { set_error("bad arguments"); return 1; }
    if (m == 0) return 0;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    attention_weights_kernel<<<(unsigned)((m + 127) / 128), 128, 0, (cudaStream_t)stream>>>(query_features, W, b, m, n_features,
                                                                                         n_signals, out);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_attention_fuse(int device, const double *probs, int64_t m, int n_signals, const double *weights, int64_t n_weight_rows,
                        double scale, int has_base_rate, double logit_base_rate, int normalize, double *out, void *stream) {
    if (m < 0 || n_signals < 1 || (n_weight_rows != 1 && n_weight_rows != m) || (m > 0 && (!probs || !weights || !out))) {
        set_error("bad arguments (weights: one row, or one row per candidate)");
        return 1;
    }
    if (m == 0) return 0;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    double *mm = nullptr;
    if (normalize) {
        BB25_CUDA(cudaMallocAsync(&mm, sizeof(double) * 2 * (size_t)n_signals, st));
        if (column_minmax(probs, m, n_signals, 0, mm, st)) { cudaFreeAsync(mm, st); return 1; }
    }
    attention_fuse_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(probs, m, n_signals, weights, n_weight_rows, scale,
                                                                     has_base_rate, logit_base_rate, mm, out);
    count_launch();
    const cudaError_t e = cudaGetLastError();
    if (mm) cudaFreeAsync(mm, st);
    if (e != cudaSuccess) { set_error("attention_fuse_kernel failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

int bb25_balanced_fusion(int device, const double *sparse_probs, const double *dense_similarities, int64_t n, double weight,
                         double *out, void *stream) {
    if (n < 0 || (n > 0 && (!sparse_probs || !dense_similarities || !out))) { set_error("bad arguments"); return 1; }
    if (n == 0) return 0;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    double *mm = nullptr;
    BB25_CUDA(cudaMallocAsync(&mm, sizeof(double) * 4, st));
    int rc = column_minmax(sparse_probs, n, 1, 0, mm, st);
    if (!rc) rc = column_minmax(dense_similarities, n, 1, 1, mm + 2, st);
    if (!rc) {
        balanced_fusion_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sparse_probs, dense_similarities, n, weight, mm, out);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("balanced_fusion_kernel failed"); rc = 1; }
    }
    cudaFreeAsync(mm, st);
    return rc;
}

}  // extern "C"
