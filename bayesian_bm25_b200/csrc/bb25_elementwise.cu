// Elementwise probability / fusion surfaces (a7-a11, a13, a14), BlockMaxIndex
// builders (a12), the multi-shard merge (K7) and the dense fp64 top-k (a15).
#include <stdlib.h>

#include "bb25_internal.cuh"

namespace bb25 {

enum {
    OP_SIGMOID, OP_LOGIT, OP_LIKELIHOOD, OP_TF_PRIOR, OP_NORM_PRIOR, OP_COMPOSITE,
    OP_POSTERIOR, OP_S2P, OP_S2P_PRIOR, OP_WAND, OP_COSINE
};

struct EwArgs {
    const double *a, *b, *c, *d;
    double *out;
    int64_t n, out_stride;
    bb25_params p;
    double x0;
    int flag;
};

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const __grid_constant__ EwArgs g) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += stride) {
        double r;
        if (OP == OP_SIGMOID) r = d_sigmoid(g.a[i]);
        else if (OP == OP_LOGIT) r = d_logit(g.a[i]);
        else if (OP == OP_LIKELIHOOD) r = d_sigmoid(g.p.alpha * (g.a[i] - g.p.beta));
        else if (OP == OP_TF_PRIOR) r = d_tf_prior(g.a[i]);
        else if (OP == OP_NORM_PRIOR) r = d_norm_prior(g.a[i]);
        else if (OP == OP_COMPOSITE) r = d_composite_prior(g.a[i], g.b[i]);
        else if (OP == OP_POSTERIOR) r = d_posterior(g.a[i], g.b[i], g.flag, g.x0);
        else if (OP == OP_S2P) {
            const double l = d_sigmoid(g.p.alpha * (g.a[i] - g.p.beta));
            const double pr = g.p.prior_mode == 1 ? 0.5 : d_composite_prior(g.b[i], g.c[i]);
            r = d_posterior(l, pr, g.p.has_base_rate, g.p.base_rate);
        } else if (OP == OP_S2P_PRIOR) {
            const double l = d_sigmoid(g.p.alpha * (g.a[i] - g.p.beta));
            r = d_posterior(l, clamp_prob(g.d[i]), g.p.has_base_rate, g.p.base_rate);
        } else if (OP == OP_WAND) {
            const double l = d_sigmoid(g.p.alpha * (g.a[i] - g.p.beta));
            r = d_posterior(l, g.x0, g.p.has_base_rate, g.p.base_rate);
        } else {  // OP_COSINE, fusion.py:43-45
            r = clamp_prob((1.0 + g.a[i]) / 2.0);
        }
        g.out[i * g.out_stride] = r;
    }
}

template <int OP>
static int run_ew(int device, EwArgs g, void *stream) {
    if (g.n < 0 || !g.out || (g.n > 0 && !g.a)) { set_error("bad elementwise arguments"); return 1; }
    if (g.n == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    int64_t blocks = (g.n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ew_kernel<OP><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g);
    BB25_LAUNCH_CHECK();
    return 0;
}

// one cosine / probability signal of a log-odds conjunction, accumulated in place
__global__ void __launch_bounds__(256) fuse_signal_kernel(const float *__restrict__ cosv,
                                                          const double *__restrict__ probs, int64_t n,
                                                          FuseSpec f, double *__restrict__ acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double p = cosv ? clamp_prob((1.0 + (double)cosv[i]) / 2.0) : probs[i];  // fusion.py:43-45
        acc[i] = fuse_step(acc[i], d_logit(p), f);
    }
}

// fusion.py:119-169
__device__ inline double d_gate(double x, int gating, double gb) {
    switch (gating) {
    case BB25_GATE_RELU: return x > 0.0 ? x : 0.0;
    case BB25_GATE_SWISH: return x * d_sigmoid(gb * x);
    case BB25_GATE_GELU: return x * d_sigmoid(1.702 * x);
    case BB25_GATE_SOFTPLUS: {  // np.logaddexp(0, gb*x) / gb
        const double y = gb * x;
        return ((y > 0.0 ? y : 0.0) + log1p(exp(-fabs(y)))) / gb;
    }
    default: return x;
    }
}

// fusion.py:243-280: one thread per row, signals summed left to right
__global__ void __launch_bounds__(256) loc_kernel(const double *__restrict__ probs, int64_t m, int n,
                                                  const double *__restrict__ w, double scale, int gating,
                                                  double gb, int has_ml, double ml, double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        double acc = 0.0;
        for (int j = 0; j < n; j++) {
            double x = d_gate(d_logit(probs[i * n + j]), gating, gb);
            if (has_ml) x = x < -ml ? -ml : (x > ml ? ml : x);
            acc += w ? w[j] * x : x;
        }
        const double l = w ? scale * acc : (acc / (double)n) * scale;
        out[i] = d_sigmoid(l);
    }
}

// scorer.py:55-81: one thread per (term, block)
__global__ void __launch_bounds__(256) blockmax_dense_kernel(const double *__restrict__ sm, int64_t n_terms,
                                                             int64_t n_docs, int bs, int64_t nb,
                                                             double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_terms * nb; i += stride) {
        const int64_t t = i / nb, b = i % nb;
        const int64_t s = b * bs, e = s + bs < n_docs ? s + bs : n_docs;
        double mx = sm[t * n_docs + s];
        for (int64_t d = s + 1; d < e; d++) {
            const double v = sm[t * n_docs + d];
            mx = v > mx ? v : mx;
        }
        out[i] = mx;
    }
}

// block maxima of CSC columns; values are >= 0 so the fp32 bit pattern orders like
// an unsigned integer and atomicMax applies
__global__ void __launch_bounds__(256) blockmax_csc_kernel(const float *__restrict__ data,
                                                           const int32_t *__restrict__ indices,
                                                           const int64_t *__restrict__ indptr,
                                                           const int32_t *__restrict__ terms, int n_terms,
                                                           int64_t n_vocab, int bs, int64_t nb,
                                                           unsigned int *__restrict__ out) {
    for (int t = blockIdx.y; t < n_terms; t += gridDim.y) {
        const int32_t term = terms[t];
        if (term < 0 || term >= n_vocab) continue;
        const int64_t s = indptr[term], e = indptr[term + 1];
        for (int64_t j = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < e;
             j += (int64_t)gridDim.x * blockDim.x)
            atomicMax(&out[(int64_t)t * nb + indices[j] / bs], __float_as_uint(data[j]));
    }
}

// ---- K7: merge S sorted [Q,k] lists --------------------------------------------------
// One CTA per query: the S*k (score, global id) keys are loaded into shared memory, the
// k-th largest is found by radix select, the k winners are compacted with their source
// positions and sorted (k elements instead of S*k).
// PACKED: the lists arrive as 16-byte entries {merge key, fp64 probability bits} (one
// all-gather instead of three, 16 instead of 20 bytes per entry); `ids` then points to them.
template <int NT, bool PACKED>
__global__ void __launch_bounds__(NT) merge_kernel(const int64_t *__restrict__ ids,
                                                   const float *__restrict__ scores,
                                                   const double *__restrict__ probs, int S, int64_t Q,
                                                   int k, int kpad, int64_t *__restrict__ out_ids,
                                                   float *__restrict__ out_scores,
                                                   double *__restrict__ out_probs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = S * k;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *top = keys + n;
    unsigned int *tsrc = reinterpret_cast<unsigned int *>(top + kpad);
    unsigned int *hist = tsrc + kpad;
    unsigned int *st = hist + 256;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += NT) {
        const int s = i / k, r = i % k;
        const int64_t o = ((int64_t)s * Q + q) * k + r;
        // (score desc, global id asc); ids < 2^32
        keys[i] = PACKED ? (unsigned long long)ids[2 * o]
                         : (((unsigned long long)__float_as_uint(scores[o]) << 32) |
                            (unsigned long long)(0xFFFFFFFFu - (uint32_t)ids[o]));
    }
    for (int i = tid; i < kpad; i += NT) {
        top[i] = 0ull;
        tsrc[i] = 0xFFFFFFFFu;
    }
    if (tid == 0) st[2] = 0u;
    __syncthreads();
    const unsigned long long kth = block_kth_largest<NT>(keys, n, k, hist, st, tid);
    for (int i = tid; i < n; i += NT) {
        const unsigned long long key = keys[i];
        if (key >= kth) {
            const unsigned int pos = atomicAdd(&st[2], 1u);
            if (pos < (unsigned)kpad) {
                top[pos] = key;
                tsrc[pos] = (unsigned int)i;
            }
        }
    }
    __syncthreads();
    for (int size = 2; size <= kpad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (kpad >> 1); t += NT) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = top[lo], y = top[hi];
                // padding (src == ~0) loses against any real entry with an equal key
                const bool less = x < y || (x == y && tsrc[lo] > tsrc[hi]);
                if (less == desc) {
                    top[lo] = y;
                    top[hi] = x;
                    const unsigned int sx = tsrc[lo];
                    tsrc[lo] = tsrc[hi];
                    tsrc[hi] = sx;
                }
            }
            __syncthreads();
        }
    for (int r = tid; r < k; r += NT) {
        const unsigned int i = tsrc[r];
        const int s = i / k, rr = i % k;
        const int64_t o = ((int64_t)s * Q + q) * k + rr;
        if (PACKED) {
            const unsigned long long key = top[r];
            out_ids[q * k + r] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
            if (out_scores) out_scores[q * k + r] = __uint_as_float((uint32_t)(key >> 32));
            out_probs[q * k + r] = __longlong_as_double((long long)ids[2 * o + 1]);
        } else {
            out_ids[q * k + r] = ids[o];
            if (out_scores) out_scores[q * k + r] = scores[o];
            out_probs[q * k + r] = probs[o];
        }
    }
}

// The same merge exploiting that every input list is sorted (the C ABI's contract): a tree of
// bitonic merges.  Lists sit in shared memory padded to kpad entries; for a pair (A, B), element
// i of A is replaced by max(A[i], B[kpad-1-i]) -- the kpad largest of the union, as a bitonic
// sequence -- and log2(kpad) compare-exchange stages sort it; log2(S) levels.  No radix select,
// no atomics: 11 barriers per level instead of 24 + 55 per merge.
template <int NT, bool PACKED>
__global__ void __launch_bounds__(NT) merge_sorted_kernel(const int64_t *__restrict__ ids,
                                                          const float *__restrict__ scores,
                                                          const double *__restrict__ probs, int S, int spad,
                                                          int64_t Q, int k, int kpad, int64_t *__restrict__ out_ids,
                                                          float *__restrict__ out_scores,
                                                          double *__restrict__ out_probs) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    unsigned int *src = reinterpret_cast<unsigned int *>(keys + (size_t)spad * kpad);
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < spad * kpad; i += NT) {
        const int s = i / kpad, r = i % kpad;
        unsigned long long key = 0ull;
        unsigned int from = 0xFFFFFFFFu;
        if (s < S && r < k) {
            const int64_t o = ((int64_t)s * Q + q) * k + r;
            key = PACKED ? (unsigned long long)ids[2 * o]
                         : (((unsigned long long)__float_as_uint(scores[o]) << 32) |
                            (unsigned long long)(0xFFFFFFFFu - (uint32_t)ids[o]));
            from = (unsigned int)(s * k + r);
        }
        keys[i] = key;
        src[i] = from;
    }
    __syncthreads();
    for (int step = 1; step < spad; step <<= 1) {
        const int pairs = spad / (2 * step);
        for (int i = tid; i < pairs * kpad; i += NT) {
            const int pr = i / kpad, e = i % kpad;
            const int ia = (2 * pr * step) * kpad + e;
            const int ib = ((2 * pr + 1) * step) * kpad + (kpad - 1 - e);
            if (keys[ib] > keys[ia]) {
                keys[ia] = keys[ib];
                src[ia] = src[ib];
            }
        }
        __syncthreads();
        for (int stride = kpad >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < pairs * (kpad >> 1); i += NT) {
                const int pr = i / (kpad >> 1), t = i % (kpad >> 1);
                const int lo = (2 * pr * step) * kpad + 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const unsigned long long x = keys[lo], y = keys[hi];
                if (x < y) {
                    keys[lo] = y;
                    keys[hi] = x;
                    const unsigned int sx = src[lo];
                    src[lo] = src[hi];
                    src[hi] = sx;
                }
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < k; r += NT) {
        const unsigned int i = src[r];
        const int s = i / k, rr = i % k;
        const int64_t o = ((int64_t)s * Q + q) * k + rr;
        if (PACKED) {
            const unsigned long long key = keys[r];
            out_ids[q * k + r] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
            if (out_scores) out_scores[q * k + r] = __uint_as_float((uint32_t)(key >> 32));
            out_probs[q * k + r] = __longlong_as_double((long long)ids[2 * o + 1]);
        } else {
            out_ids[q * k + r] = ids[o];
            if (out_scores) out_scores[q * k + r] = scores[o];
            out_probs[q * k + r] = probs[o];
        }
    }
}

// launches the merge: bitonic merge tree when the padded lists fit in shared memory
template <bool PACKED>
static int launch_merge(const int64_t *ids, const float *scores, const double *probs, int n_shards, int64_t n_queries,
                        int k, int64_t *out_ids, float *out_scores, double *out_probs, cudaStream_t st) {
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    int spad = 1;
    while (spad < n_shards) spad <<= 1;
    const size_t smem_tree = (size_t)spad * kpad * 12;
    const char *force = getenv("BB25_MERGE");  // "radix": the select-and-sort kernel (A/B only)
    if (smem_tree <= 200 * 1024 && !(force && force[0] == 'r')) {
        BB25_CUDA(cudaFuncSetAttribute(merge_sorted_kernel<512, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tree));
        merge_sorted_kernel<512, PACKED><<<(unsigned)n_queries, 512, smem_tree, st>>>(ids, scores, probs, n_shards, spad, n_queries,
                                                                                   k, kpad, out_ids, out_scores, out_probs);
    } else {
        const size_t smem = (size_t)n_shards * k * 8 + (size_t)kpad * 12 + 260 * 4;
        BB25_CUDA(cudaFuncSetAttribute(merge_kernel<512, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_kernel<512, PACKED><<<(unsigned)n_queries, 512, smem, st>>>(ids, scores, probs, n_shards, n_queries, k, kpad,
                                                                       out_ids, out_scores, out_probs);
    }
    BB25_LAUNCH_CHECK();
    return 0;
}

// (ids, scores, probs) -> 16-byte entries {(score bits << 32) | (2^32-1 - id), probability bits}
__global__ void __launch_bounds__(256) pack_topk_kernel(const int64_t *__restrict__ ids, const float *__restrict__ scores,
                                                        const double *__restrict__ probs, int64_t n,
                                                        int64_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[2 * i] = (int64_t)(((unsigned long long)__float_as_uint(scores[i]) << 32) |
                               (unsigned long long)(0xFFFFFFFFu - (uint32_t)ids[i]));
        out[2 * i + 1] = __double_as_longlong(probs[i]);
    }
}

// ---- dense fp64 top-k: MSB-first radix select on (value bits desc, index asc) ---------
// state: [0] prefix (value bits), [1] mask, [2] remaining, [3] idx prefix, [4] idx mask,
//        [5] remaining among ties, [6] output cursor
__global__ void topk_hist_kernel(const double *__restrict__ v, int64_t n, const unsigned long long *state,
                                 int shift, int bits, int phase, unsigned int *hist) {
    __shared__ unsigned int sh[2048];
    if (state[8]) return;                // the seeded fast path already produced the result
    if (phase == 1 && state[7]) return;  // every tie on the threshold value is taken: nothing to select
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned long long prefix = state[0], mask = state[1];
    const unsigned long long iprefix = state[3], imask = state[4];
    const unsigned int dm = (1u << bits) - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // block-uniform trip count: the warp-wide votes below need converged warps.  Runs of
    // equal values are common here (every unmatched document carries the same fused
    // probability).
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        const int64_t i = base + threadIdx.x;
        bool in = false;
        unsigned int digit = 0u;
        if (i < n) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(v[i]);
            if (phase == 0) {
                in = (b & mask) == prefix;
                digit = (unsigned int)(b >> shift) & dm;
            } else {
                // ties on the threshold value: select the smallest indices (digit of ~index)
                const unsigned long long inv = 0xFFFFFFFFFFFFFFFFull - (unsigned long long)i;
                in = b == prefix && (inv & imask) == iprefix;
                digit = (unsigned int)(inv >> shift) & dm;
            }
        }
        // cheap aggregation of the all-equal case (one shuffle + two votes); mixed warps
        // fall back to per-lane atomics
        const unsigned int act = __ballot_sync(0xFFFFFFFFu, in);
        if (act) {
            const unsigned int d0 = __shfl_sync(0xFFFFFFFFu, digit, __ffs(act) - 1);
            if (__all_sync(0xFFFFFFFFu, !in || digit == d0)) {
                if ((int)(__ffs(act) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[d0], (unsigned int)__popc(act));
            } else if (in) {
                atomicAdd(&sh[digit], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void topk_scan_kernel(unsigned long long *state, int shift, int bits, int phase, unsigned int *hist) {
    if (state[8]) return;
    if (phase == 1 && state[7]) return;
    if (threadIdx.x == 0) {
        const int nb = 1 << bits;
        unsigned long long rem = phase == 0 ? state[2] : state[5];
        unsigned long long cum = 0;
        int d = nb - 1;
        for (; d > 0; d--) {
            if (cum + hist[d] >= rem) break;
            cum += hist[d];
        }
        rem -= cum;
        const unsigned long long dm = (unsigned long long)(nb - 1);
        if (phase == 0) {
            state[0] |= (unsigned long long)d << shift;
            state[1] |= dm << shift;
            state[2] = rem;
            // after the last value pass hist[d] = number of elements equal to the threshold
            if (shift == 0) state[7] = (rem == (unsigned long long)hist[d]) ? 1ull : 0ull;
        } else {
            state[3] |= (unsigned long long)d << shift;
            state[4] |= dm << shift;
            state[5] = rem;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
}

// state words: [0..7] radix select (see above), [8] done flag, [9] seed threshold, [10] seeded-collect counter
__global__ void topk_init_kernel(unsigned long long *state, unsigned int *hist, int k) {
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x < 16) state[threadIdx.x] = threadIdx.x == 2 ? (unsigned long long)k : 0ull;
}

// ---- seeded fast path for k <= number of strips -----------------------------------------
// The maxima of G disjoint strips are G distinct elements, so the k-th largest strip maximum
// is a valid lower bound of the k-th largest element.  Two scans (strip maxima, collect >=
// seed) and one small sort replace the 12 radix passes whenever the collected set fits.
constexpr int kTopkStrips = 148 * 8;
constexpr int kTopkCollectCap = 8192;

__global__ void __launch_bounds__(256) topk_strip_max_kernel(const double *__restrict__ v, int64_t n,
                                                             unsigned long long *__restrict__ strip_max) {
    __shared__ unsigned long long sm[8];
    const int64_t len = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * len, hi = lo + len < n ? lo + len : n;
    unsigned long long m = 0ull;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v[i]);
        m = b > m ? b : m;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, m, d);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = sm[w] > m ? sm[w] : m;
        strip_max[blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(1024) topk_seed_kernel(const unsigned long long *__restrict__ strip_max, int n_strips,
                                                         int k, unsigned long long *state) {
    __shared__ unsigned long long s[2048];
    for (int i = threadIdx.x; i < 2048; i += 1024) s[i] = i < n_strips ? strip_max[i] : 0ull;
    __syncthreads();
    for (int size = 2; size <= 2048; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int t = threadIdx.x;
            const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
            const bool desc = (lo & size) == 0;
            const unsigned long long x = s[lo], y = s[hi];
            if ((x < y) == desc) { s[lo] = y; s[hi] = x; }
            __syncthreads();
        }
    if (threadIdx.x == 0) state[9] = s[k - 1];
}

__global__ void __launch_bounds__(256) topk_collect_seeded_kernel(const double *__restrict__ v, int64_t n,
                                                                  unsigned long long *state,
                                                                  unsigned long long *cb, unsigned long long *ci) {
    const unsigned long long seed = state[9];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v[i]);
        if (b >= seed) {
            const unsigned long long pos = atomicAdd(&state[10], 1ull);
            if (pos < (unsigned long long)kTopkCollectCap) {
                cb[pos] = b;
                ci[pos] = 0xFFFFFFFFFFFFFFFFull - (unsigned long long)i;
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) topk_seeded_final_kernel(unsigned long long *state, const unsigned long long *cb,
                                                               const unsigned long long *ci, int k,
                                                               int64_t *out_ids, double *out_vals) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned long long cnt = state[10];
    if (cnt > (unsigned long long)kTopkCollectCap || cnt < (unsigned long long)k) return;  // radix path takes over
    int P = 2;
    while (P < (int)cnt) P <<= 1;
    unsigned long long *kb = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *ki = kb + P;
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += NT) {
        kb[i] = i < (int)cnt ? cb[i] : 0ull;
        ki[i] = i < (int)cnt ? ci[i] : 0ull;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (P >> 1); t += NT) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long xb = kb[lo], yb = kb[hi], xi = ki[lo], yi = ki[hi];
                const bool less = xb < yb || (xb == yb && xi < yi);
                if (less == desc) {
                    kb[lo] = yb; kb[hi] = xb;
                    ki[lo] = yi; ki[hi] = xi;
                }
            }
            __syncthreads();
        }
    for (int r = tid; r < k; r += NT) {
        out_ids[r] = (int64_t)(0xFFFFFFFFFFFFFFFFull - ki[r]);
        out_vals[r] = __longlong_as_double((long long)kb[r]);
    }
    if (tid == 0) state[8] = 1ull;
}

__global__ void topk_begin_ties_kernel(unsigned long long *state) {
    if (state[8]) return;
    // after the value phase state[2] = how many elements equal to the threshold are needed
    state[3] = 0;  // with state[7] set this already means "every tie qualifies"
    state[4] = 0;
    state[5] = state[2];
    state[6] = 0;
}

// gather the k winners (unordered) as (value bits, ~index) pairs
__global__ void topk_collect_kernel(const double *__restrict__ v, int64_t n, unsigned long long *state, int k,
                                    unsigned long long *cand_bits, unsigned long long *cand_inv) {
    if (state[8]) return;
    const unsigned long long thr = state[0];
    const unsigned long long ithr = state[3];  // smallest admissible ~index among ties
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v[i]);
        const unsigned long long inv = 0xFFFFFFFFFFFFFFFFull - (unsigned long long)i;
        if (b > thr || (b == thr && inv >= ithr)) {
            const unsigned long long pos = atomicAdd(&state[6], 1ull);
            if (pos < (unsigned long long)k) {
                cand_bits[pos] = b;
                cand_inv[pos] = inv;
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) topk_sort_kernel(const unsigned long long *state,
                                                       const unsigned long long *cand_bits,
                                                       const unsigned long long *cand_inv, int k, int P,
                                                       int64_t *out_ids, double *out_vals) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (state[8]) return;
    unsigned long long *kb = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *ki = kb + P;
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += NT) {
        kb[i] = i < k ? cand_bits[i] : 0ull;
        ki[i] = i < k ? cand_inv[i] : 0ull;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (P >> 1); t += NT) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long xb = kb[lo], yb = kb[hi], xi = ki[lo], yi = ki[hi];
                const bool less = xb < yb || (xb == yb && xi < yi);
                if (less == desc) {
                    kb[lo] = yb; kb[hi] = xb;
                    ki[lo] = yi; ki[hi] = xi;
                }
            }
            __syncthreads();
        }
    for (int r = tid; r < k; r += NT) {
        out_ids[r] = (int64_t)(0xFFFFFFFFFFFFFFFFull - ki[r]);
        out_vals[r] = __longlong_as_double((long long)kb[r]);
    }
}

}  // namespace bb25

using namespace bb25;

extern "C" {

#define EW_BEGIN EwArgs g{}; g.out = out; g.n = n; g.out_stride = 1;

int bb25_sigmoid(int device, const double *x, int64_t n, double *out, void *stream) {
    EW_BEGIN g.a = x;
    return run_ew<OP_SIGMOID>(device, g, stream);
}
int bb25_logit(int device, const double *p, int64_t n, double *out, void *stream) {
    EW_BEGIN g.a = p;
    return run_ew<OP_LOGIT>(device, g, stream);
}
int bb25_likelihood(int device, const bb25_params *p, const double *score, int64_t n, double *out, void *stream) {
    if (!p) { set_error("params is NULL"); return 1; }
    EW_BEGIN g.a = score; g.p = *p;
    return run_ew<OP_LIKELIHOOD>(device, g, stream);
}
int bb25_tf_prior(int device, const double *tf, int64_t n, double *out, void *stream) {
    EW_BEGIN g.a = tf;
    return run_ew<OP_TF_PRIOR>(device, g, stream);
}
int bb25_norm_prior(int device, const double *ratio, int64_t n, double *out, void *stream) {
    EW_BEGIN g.a = ratio;
    return run_ew<OP_NORM_PRIOR>(device, g, stream);
}
int bb25_composite_prior(int device, const double *tf, const double *ratio, int64_t n, double *out, void *stream) {
    if (n > 0 && !ratio) { set_error("ratio is NULL"); return 1; }
    EW_BEGIN g.a = tf; g.b = ratio;
    return run_ew<OP_COMPOSITE>(device, g, stream);
}
int bb25_posterior(int device, const double *lik, const double *prior, int has_base_rate, double base_rate,
                   int64_t n, double *out, void *stream) {
    if (n > 0 && !prior) { set_error("prior is NULL"); return 1; }
    EW_BEGIN g.a = lik; g.b = prior; g.flag = has_base_rate; g.x0 = base_rate;
    return run_ew<OP_POSTERIOR>(device, g, stream);
}
int bb25_score_to_probability(int device, const bb25_params *p, const double *score, const double *tf,
                              const double *ratio, const double *prior, int64_t n, double *out,
                              void *stream) {
    if (!p) { set_error("params is NULL"); return 1; }
    EW_BEGIN g.a = score; g.b = tf; g.c = ratio; g.d = prior; g.p = *p;
    if (prior) return run_ew<OP_S2P_PRIOR>(device, g, stream);
    if (p->prior_mode != 1 && n > 0 && (!tf || !ratio)) { set_error("tf / ratio is NULL"); return 1; }
    return run_ew<OP_S2P>(device, g, stream);
}
int bb25_wand_upper_bound(int device, const bb25_params *p, const double *bm25_ub, double p_max, int64_t n,
                          double *out, void *stream) {
    if (!p) { set_error("params is NULL"); return 1; }
    EW_BEGIN g.a = bm25_ub; g.p = *p; g.x0 = p_max;
    return run_ew<OP_WAND>(device, g, stream);
}
int bb25_cosine_to_probability(int device, const double *cosv, int64_t n, double *out, int64_t out_stride,
                               void *stream) {
    if (out_stride < 1) { set_error("out_stride must be >= 1"); return 1; }
    EW_BEGIN g.a = cosv; g.out_stride = out_stride;
    return run_ew<OP_COSINE>(device, g, stream);
}

int bb25_log_odds_conjunction(int device, const double *probs, int64_t m, int n, const double *weights,
                              double scale, int gating, double gating_beta, int has_max_logit,
                              double max_logit, double *out, void *stream) {
    if (m < 0 || n < 1 || !out || (m > 0 && !probs)) { set_error("bad log_odds_conjunction arguments"); return 1; }
    if (gating < BB25_GATE_NONE || gating > BB25_GATE_SOFTPLUS) { set_error("unknown gating %d", gating); return 1; }
    if (m == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    int64_t blocks = (m + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    loc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(probs, m, n, weights, scale, gating, gating_beta,
                                                                  has_max_logit, max_logit, out);
    BB25_LAUNCH_CHECK();
    return 0;
}

static int run_fuse_signal(int device, const float *cosv, const double *probs, int64_t n, double weight,
                           int n_signals, double scale, int flags, double *acc, void *stream) {
    if (n < 0 || !acc || n_signals < 1 || (flags & ~7) || (n > 0 && !cosv && !probs)) { set_error("bad fuse arguments"); return 1; }
    if (n == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    FuseSpec f{weight, scale, n_signals, flags};
    fuse_signal_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(cosv, probs, n, f, acc);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_fuse_cosine_signal(int device, const float *cosv, int64_t n, double weight, int n_signals, double scale,
                            int flags, double *acc, void *stream) {
    return run_fuse_signal(device, cosv, nullptr, n, weight, n_signals, scale, flags, acc, stream);
}
int bb25_fuse_prob_signal(int device, const double *probs, int64_t n, double weight, int n_signals, double scale,
                          int flags, double *acc, void *stream) {
    return run_fuse_signal(device, nullptr, probs, n, weight, n_signals, scale, flags, acc, stream);
}

int bb25_blockmax_dense(int device, const double *score_matrix, int64_t n_terms, int64_t n_docs,
                        int block_size, double *out, void *stream) {
    if (block_size < 1 || n_terms < 0 || n_docs < 1 || !score_matrix || !out) { set_error("bad blockmax arguments"); return 1; }
    if (n_terms == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    const int64_t nb = (n_docs + block_size - 1) / block_size;
    int64_t blocks = (n_terms * nb + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    blockmax_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(score_matrix, n_terms, n_docs, block_size, nb, out);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_blockmax_csc(bb25_index *idx, const int32_t *terms, int n_terms, int block_size, float *out,
                      void *stream) {
    if (!idx || block_size < 1 || n_terms < 0 || !out || (n_terms > 0 && !terms)) { set_error("bad blockmax arguments"); return 1; }
    if (n_terms == 0) return 0;
    DeviceGuard dg(idx->device);
    if (!dg.ok) { set_error("cannot select device"); return 1; }
    const int64_t nb = (idx->n_docs + block_size - 1) / block_size;
    cudaStream_t st = (cudaStream_t)stream;
    BB25_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_terms * (size_t)nb, st));
    dim3 grid(64, (unsigned)(n_terms < 1024 ? n_terms : 1024));
    blockmax_csc_kernel<<<grid, 256, 0, st>>>(idx->data, idx->indices, idx->indptr, terms, n_terms, idx->n_vocab,
                                              block_size, nb, reinterpret_cast<unsigned int *>(out));
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_merge_topk(int device, const int64_t *ids, const float *scores, const double *probs, int n_shards,
                    int64_t n_queries, int k, int64_t *out_ids, float *out_scores, double *out_probs,
                    void *stream) {
    if (n_shards < 1 || n_queries < 0 || k < 1 || !ids || !scores || !probs || !out_ids || !out_probs) {
        set_error("bad merge arguments");
        return 1;
    }
    if ((int64_t)n_shards * k > 16384) { set_error("n_shards * k must be <= 16384"); return 1; }
    if (n_queries == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    return launch_merge<false>(ids, scores, probs, n_shards, n_queries, k, out_ids, out_scores, out_probs, (cudaStream_t)stream);
}

int bb25_pack_topk(int device, const int64_t *ids, const float *scores, const double *probs, int64_t n,
                   int64_t *out_packed, void *stream) {
    if (n < 0 || !out_packed || (n > 0 && (!ids || !scores || !probs))) { set_error("bad pack arguments"); return 1; }
    if (n == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_topk_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ids, scores, probs, n, out_packed);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_merge_topk_packed(int device, const int64_t *packed, int n_shards, int64_t n_queries, int k,
                           int64_t *out_ids, float *out_scores, double *out_probs, void *stream) {
    if (n_shards < 1 || n_queries < 0 || k < 1 || !packed || !out_ids || !out_probs) { set_error("bad merge arguments"); return 1; }
    if ((int64_t)n_shards * k > 16384) { set_error("n_shards * k must be <= 16384"); return 1; }
    if (n_queries == 0) return 0;
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    return launch_merge<true>(packed, nullptr, nullptr, n_shards, n_queries, k, out_ids, out_scores, out_probs, (cudaStream_t)stream);
}

int bb25_topk_f64(int device, const double *vals, int64_t n, int k, int64_t *out_ids, double *out_vals,
                  void *stream) {
    if (!vals || !out_ids || !out_vals || n < 1 || k < 1 || (int64_t)k > n || k > 8192) {
        set_error("bad topk arguments (need 1 <= k <= min(n, 8192))");
        return 1;
    }
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *ws = nullptr;
    const size_t o_hist = 128, o_cb = o_hist + 2048 * 4, o_ci = o_cb + (size_t)k * 8;
    const size_t o_sm = o_ci + (size_t)k * 8, o_cb2 = o_sm + (size_t)kTopkStrips * 8;
    const size_t o_ci2 = o_cb2 + (size_t)kTopkCollectCap * 8;
    BB25_CUDA(cudaMallocAsync(&ws, o_ci2 + (size_t)kTopkCollectCap * 8, st));
    unsigned long long *state = (unsigned long long *)ws;
    unsigned int *hist = (unsigned int *)(ws + o_hist);
    unsigned long long *cb = (unsigned long long *)(ws + o_cb), *ci = (unsigned long long *)(ws + o_ci);
    unsigned long long *strip_max = (unsigned long long *)(ws + o_sm);
    unsigned long long *cb2 = (unsigned long long *)(ws + o_cb2), *ci2 = (unsigned long long *)(ws + o_ci2);
    int rc = 1;
    do {
        topk_init_kernel<<<1, 256, 0, st>>>(state, hist, k);
        count_launch();
        int64_t blocks = (n + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (k <= kTopkStrips && n >= (int64_t)kTopkStrips * 64) {
            // seeded fast path; on overflow (massive ties) it leaves state[8] == 0 and the radix
            // passes below do the work, otherwise they return at once
            topk_strip_max_kernel<<<kTopkStrips, 256, 0, st>>>(vals, n, strip_max);
            topk_seed_kernel<<<1, 1024, 0, st>>>(strip_max, kTopkStrips, k, state);
            topk_collect_seeded_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, n, state, cb2, ci2);
            if (cudaFuncSetAttribute(topk_seeded_final_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTopkCollectCap * 16) != cudaSuccess) break;
            topk_seeded_final_kernel<512><<<1, 512, (size_t)kTopkCollectCap * 16, st>>>(state, cb2, ci2, k, out_ids, out_vals);
            count_launch(4);
        }
        // value phase: 64 bits as 11,11,11,11,11,9
        const int vshift[6] = {53, 42, 31, 20, 9, 0}, vbits[6] = {11, 11, 11, 11, 11, 9};
        for (int p = 0; p < 6; p++) {
            topk_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, n, state, vshift[p], vbits[p], 0, hist);
            topk_scan_kernel<<<1, 256, 0, st>>>(state, vshift[p], vbits[p], 0, hist);
            count_launch(2);
        }
        topk_begin_ties_kernel<<<1, 1, 0, st>>>(state);
        count_launch();
        // tie phase on ~index (64 bits, same digit plan)
        for (int p = 0; p < 6; p++) {
            topk_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, n, state, vshift[p], vbits[p], 1, hist);
            topk_scan_kernel<<<1, 256, 0, st>>>(state, vshift[p], vbits[p], 1, hist);
            count_launch(2);
        }
        topk_collect_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, n, state, k, cb, ci);
        count_launch();
        int P = 2;
        while (P < k) P <<= 1;
        if (cudaFuncSetAttribute(topk_sort_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16) != cudaSuccess) break;
        topk_sort_kernel<512><<<1, 512, (size_t)P * 16, st>>>(state, cb, ci, k, P, out_ids, out_vals);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) break;
        rc = 0;
    } while (0);
    if (rc) set_error("CUDA failure in bb25_topk_f64: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFreeAsync(ws, st);
    return rc;
}

}  // extern "C"
