// Multi-GPU exchange step of document-range sharding, B200-first:
//  * merge_peers_kernel: ONE kernel that is both the collective and the merge.  Rank r owns the queries
//    [q_begin, q_begin + n_q); for each of them it pulls the S per-shard sorted lists straight out of the
//    peers' memory (NVLink loads through the symmetric-memory pointer table -- no staging all-to-all), merges
//    them with the bitonic merge tree of merge_sorted_kernel, and stores the merged list into EVERY rank's
//    result buffer (NVLink stores): all-to-all + merge + all-gather in one launch, transfers overlapped with
//    the merges of other queries.  1/S of the merge work per rank instead of all of it on every rank.
//  * apply_quantiles_kernel: cross-shard thresholds.  Between block groups every shard publishes, per query,
//    the score at a few ranks of its running top-k (rank r: "this shard holds r documents with at least this
//    score"); any value v for which the shards' published counts add up to k bounds the GLOBAL k-th score from
//    below, so every shard can prune and emit against it instead of its own, looser, local k-th.
#include "bb25_internal.cuh"

namespace bb25 {

template <int NT>
__global__ void __launch_bounds__(NT) merge_peers_kernel(const int64_t *const *__restrict__ src_tab,
                                                         int64_t *const *__restrict__ dst_tab, int S, int spad,
                                                         int64_t q_begin, int k, int kpad) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    unsigned int *src = reinterpret_cast<unsigned int *>(keys + (size_t)spad * kpad);
    const int64_t q = q_begin + blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < spad * kpad; i += NT) {
        const int s = i / kpad, r = i % kpad;
        unsigned long long key = 0ull;
        unsigned int from = 0xFFFFFFFFu;
        if (s < S && r < k) {
            key = (unsigned long long)src_tab[s][2 * (q * k + r)];  // peer load
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
        long long key = 0, pb = 0;
        if (i != 0xFFFFFFFFu) {
            const int s = (int)(i / (unsigned)k), rr = (int)(i % (unsigned)k);
            key = (long long)keys[r];
            pb = src_tab[s][2 * (q * k + rr) + 1];
        }
        const int64_t o = 2 * (q * k + r);
        for (int d = 0; d < S; d++) {  // the merged row goes to every rank's result buffer
            longlong2 *p = reinterpret_cast<longlong2 *>(dst_tab[d] + o);
            *p = make_longlong2(key, pb);
        }
    }
}

__global__ void __launch_bounds__(256) unpack_topk_kernel(const int64_t *__restrict__ packed, int64_t n,
                                                          int64_t *__restrict__ ids, float *__restrict__ scores,
                                                          double *__restrict__ probs) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = (unsigned long long)packed[2 * i];
        ids[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        if (scores) scores[i] = __uint_as_float((uint32_t)(key >> 32));
        probs[i] = __longlong_as_double(packed[2 * i + 1]);
    }
}

struct QuantRanks {
    int r[4];
};

// one warp per query; lane e = (shard e / J, level e % J) holds one published score
__global__ void __launch_bounds__(256) apply_quantiles_kernel(const unsigned long long *__restrict__ all, int S, int64_t Q,
                                                              int J, QuantRanks ranks, int k,
                                                              unsigned long long *__restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    unsigned long long v = 0ull;
    if (lane < S * J) v = all[((int64_t)(lane / J) * Q + q) * J + (lane % J)];
    // how many documents do the shards hold, between them, with a score >= v ?
    long long total = 0;
    for (int s = 0; s < S; s++) {
        int best = 0;
        for (int j = 0; j < J; j++) {
            const unsigned long long o = __shfl_sync(0xFFFFFFFFu, v, s * J + j);
            if (o != 0ull && o >= v && ranks.r[j] > best) best = ranks.r[j];
        }
        total += best;
    }
    unsigned long long t = (v != 0ull && total >= (long long)k) ? v : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, t, d);
        t = o > t ? o : t;
    }
    if (lane == 0 && t > thr[q]) thr[q] = t;
}

}  // namespace bb25

using namespace bb25;

extern "C" {

void bb25_quantile_ranks(int k, int n_shards, int *n_levels, int *ranks4) {
    // levels: ceil(k/S), ceil(2k/S), ceil(4k/S), k (fewer when S is large: S * J <= 32)
    int J = n_shards <= 8 ? 4 : (n_shards <= 16 ? 2 : 1);
    for (int j = 0; j < 4; j++) ranks4[j] = k;
    for (int j = 0; j + 1 < J; j++) {
        long long r = ((long long)k * (1ll << j) + n_shards - 1) / n_shards;
        ranks4[j] = (int)(r < 1 ? 1 : (r > k ? k : r));
    }
    *n_levels = J;
}

int bb25_apply_quantiles(int device, const void *all_quantiles, int n_shards, int64_t n_queries, int k, void *d_thr,
                         void *stream) {
    if (!all_quantiles || !d_thr || n_shards < 1 || n_shards > 32 || n_queries < 0 || k < 1) { set_error("bad arguments"); return 1; }
    if (n_queries == 0) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    QuantRanks rk;
    int J = 0;
    bb25_quantile_ranks(k, n_shards, &J, rk.r);
    apply_quantiles_kernel<<<(unsigned)((n_queries + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)all_quantiles, n_shards, n_queries, J, rk, k, (unsigned long long *)d_thr);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_merge_topk_peers(int device, const void *src_tab_dev, const void *dst_tab_dev, int n_shards, int64_t q_begin,
                          int64_t n_queries, int k, void *stream) {
    if (!src_tab_dev || !dst_tab_dev || n_shards < 1 || q_begin < 0 || n_queries < 0 || k < 1) { set_error("bad arguments"); return 1; }
    if (n_queries == 0) return 0;
    int kpad = 2;
    while (kpad < k) kpad <<= 1;
    int spad = 1;
    while (spad < n_shards) spad <<= 1;
    const size_t smem = (size_t)spad * kpad * 12;
    if (smem > 200 * 1024) { set_error("n_shards * k too large for the peer merge (%d x %d)", n_shards, k); return 1; }
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    BB25_CUDA(cudaFuncSetAttribute(merge_peers_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_peers_kernel<512><<<(unsigned)n_queries, 512, smem, (cudaStream_t)stream>>>(
        (const int64_t *const *)src_tab_dev, (int64_t *const *)dst_tab_dev, n_shards, spad, q_begin, k, kpad);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_unpack_topk(int device, const int64_t *packed, int64_t n, int64_t *out_ids, float *out_scores, double *out_probs,
                     void *stream) {
    if (n < 0 || (n > 0 && (!packed || !out_ids || !out_probs))) { set_error("bad unpack arguments"); return 1; }
    if (n == 0) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    unpack_topk_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(packed, n, out_ids, out_scores, out_probs);
    BB25_LAUNCH_CHECK();
    return 0;
}

int bb25_memcpy_device(int device, void *dst, const void *src, int64_t bytes, void *stream) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) { set_error("bad copy arguments"); return 1; }
    if (bytes == 0) return 0;
    DeviceGuard dg(device);
    if (!dg.ok) { set_error("cannot select CUDA device %d", device); return 1; }
    BB25_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

}  // extern "C"
