// Index upload, per-(term, doc-tile) skip table, per-term k-th largest posting
// value (threshold seeds), workspace management, error plumbing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <utility>
#include <vector>

#include "bb25_internal.cuh"

namespace bb25 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

int ensure_workspace(bb25_index *idx, size_t bytes) {
    if (bytes <= idx->ws_bytes) return 0;
    if (idx->ws) {
        BB25_CUDA(cudaDeviceSynchronize());
        BB25_CUDA(cudaFree(idx->ws));
        idx->ws = nullptr;
        idx->ws_bytes = 0;
    }
    size_t want = bytes + (bytes >> 3);
    BB25_CUDA(cudaMalloc(&idx->ws, want));
    idx->ws_bytes = want;
    return 0;
}

// ---------------------------------------------------------------------------------
// tile_off[t][b] = first posting of term t whose doc id >= b*tile_docs, relative to
// indptr[t]; b in [0, n_tiles].  One thread per (t, b): binary search in the column.
// ---------------------------------------------------------------------------------
__global__ void build_tile_table_kernel(const int32_t *__restrict__ indices,
                                        const int64_t *__restrict__ indptr, int64_t n_vocab,
                                        int n_tiles, int tile_docs, uint32_t *__restrict__ tile_off) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_vocab * (int64_t)(n_tiles + 1);
    if (gid >= total) return;
    int64_t t = gid / (n_tiles + 1);
    int b = (int)(gid % (n_tiles + 1));
    int64_t s = indptr[t], e = indptr[t + 1];
    int64_t target = (int64_t)b * tile_docs;
    int64_t lo = s, hi = e;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)indices[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    tile_off[gid] = (uint32_t)(lo - s);
}

// ---------------------------------------------------------------------------------
// Block table (struct BlockTable in bb25_internal.cuh), term-major.
//   count_term_blocks_kernel   blocks each term touches (-> dense or bitmap form, host decides)
//   set_block_bits_kernel      bitmap words of the bitmap-form terms
//   scan_block_bits_kernel     per word: number of the term's entries before it
//   build_block_table_kernel   pass 0: first-posting offset (min) and block maximum (max of the
//                              rounded-up bit pattern; values are >= 0 so bit patterns order like
//                              the floats); pass 1 (after pass 0 finished): posting count into the
//                              low 11 bits.  Entries are pre-filled with {0xFFFFFFFF, 0}.
// One CTA walks one term's posting list at a time.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_term_blocks_kernel(const int32_t *__restrict__ indices,
                                                                const int64_t *__restrict__ indptr, int64_t n_vocab,
                                                                int32_t *__restrict__ n_touched) {
    __shared__ int s_cnt;
    for (int64_t t = blockIdx.x; t < n_vocab; t += gridDim.x) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const int64_t s = indptr[t], e = indptr[t + 1];
        int c = 0;
        for (int64_t j = s + threadIdx.x; j < e; j += blockDim.x)  // doc ids ascend: count block changes
            c += (j == s) || (indices[j] / kBlockDocs != indices[j - 1] / kBlockDocs);
        if (c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (threadIdx.x == 0) n_touched[t] = s_cnt;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) set_block_bits_kernel(const int32_t *__restrict__ indices,
                                                             const int64_t *__restrict__ indptr, int64_t n_vocab,
                                                             const longlong2 *__restrict__ row, uint2 *__restrict__ bits) {
    for (int64_t t = blockIdx.x; t < n_vocab; t += gridDim.x) {
        const longlong2 r = row[t];
        if (r.y < 0) continue;
        const int64_t s = indptr[t], e = indptr[t + 1];
        for (int64_t j = s + threadIdx.x; j < e; j += blockDim.x) {
            const int b = indices[j] / kBlockDocs;
            atomicOr(&bits[r.y + (b >> 5)].x, 1u << (b & 31));
        }
    }
}

__global__ void scan_block_bits_kernel(int64_t n_vocab, int n_words, const longlong2 *__restrict__ row,
                                       uint2 *__restrict__ bits) {
    // one warp per bitmap-form term
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_vocab) return;
    const longlong2 r = row[t];
    if (r.y < 0) return;
    unsigned int run = 0;
    for (int w0 = 0; w0 < n_words; w0 += 32) {
        const int w = w0 + lane;
        const unsigned int c = w < n_words ? (unsigned int)__popc(bits[r.y + w].x) : 0u;
        unsigned int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += y;
        }
        if (w < n_words) bits[r.y + w].y = run + incl - c;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
}

__device__ __forceinline__ uint2 *tab_slot(uint2 *ent, const uint2 *bits, longlong2 r, int b) {
    if (r.y < 0) return ent + r.x + b;
    const uint2 wb = bits[r.y + (b >> 5)];
    return ent + r.x + wb.y + __popc(wb.x & ((1u << (b & 31)) - 1u));
}

__global__ void __launch_bounds__(256) build_block_table_kernel(const float *__restrict__ data,
                                                                const int32_t *__restrict__ indices,
                                                                const int64_t *__restrict__ indptr,
                                                                int64_t n_vocab, int pass,
                                                                const longlong2 *__restrict__ row,
                                                                const uint2 *__restrict__ bits, uint2 *__restrict__ ent_base) {
    for (int64_t t = blockIdx.x; t < n_vocab; t += gridDim.x) {
        const int64_t s = indptr[t], e = indptr[t + 1];
        const longlong2 r = row[t];
        for (int64_t j = s + threadIdx.x; j < e; j += blockDim.x) {
            uint2 *ent = tab_slot(ent_base, bits, r, indices[j] / kBlockDocs);
            if (pass == 0) {
                atomicMin(&ent->x, (unsigned int)(j - s));
                unsigned int v = __float_as_uint(data[j]);
                v = (v + kBlkLenMask) & ~kBlkLenMask;  // round the bound up, never down
                atomicMax(&ent->y, v);
            } else {
                atomicAdd(&ent->y, 1u);
            }
        }
    }
}
// dense_vals[slot][doc] = value of the slot's term in doc (rows pre-filled with -0.0f = absent)
__global__ void __launch_bounds__(256) build_dense_rows_kernel(const float *__restrict__ data,
                                                               const int32_t *__restrict__ indices,
                                                               const int64_t *__restrict__ indptr,
                                                               const int32_t *__restrict__ terms, int64_t stride,
                                                               float *__restrict__ dense) {
    const int32_t t = terms[blockIdx.y];
    const int64_t s = indptr[t], e = indptr[t + 1];
    float *row = dense + (int64_t)blockIdx.y * stride;
    for (int64_t j = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < e; j += (int64_t)gridDim.x * blockDim.x)
        row[indices[j]] = data[j];
}
// fp16 upper-bound copy of the dense value rows; flag set when a value does not fit fp16
__global__ void half_rows_kernel(const float *__restrict__ v, size_t n, __half *__restrict__ out, int *flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = v[i];
    if (__float_as_uint(x) == 0x80000000u || x == 0.f) {
        out[i] = __float2half_rn(0.f);
        return;
    }
    if (x > 60000.f) atomicOr(flag, 1);
    out[i] = __float2half_ru(x);
}
__global__ void fill_u32_kernel(unsigned int *p, size_t n, unsigned int v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void init_block_table_kernel(uint2 *tab, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tab[i] = make_uint2(0xFFFFFFFFu, 0u);
}

// Input validation on device: column starts monotone, doc ids in range and strictly
// ascending inside a column, posting values >= 0 and not NaN.  flags[0] |= bit.
__global__ void validate_csc_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                    const int64_t *__restrict__ indptr, int64_t n_vocab, int64_t n_docs,
                                    int64_t nnz, int *flags) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    for (int64_t t = gid; t < n_vocab; t += stride) {
        int64_t s = indptr[t], e = indptr[t + 1];
        if (s > e || s < 0 || e > nnz) { bad |= 1; continue; }
    }
    if (gid == 0 && (indptr[0] != 0 || indptr[n_vocab] != nnz)) bad |= 1;
    for (int64_t j = gid; j < nnz; j += stride) {
        int32_t d = indices[j];
        if (d < 0 || d >= n_docs) bad |= 2;
        float v = data[j];
        if (!(v >= 0.0f)) bad |= 4;
    }
    if (bad) atomicOr(flags, bad);
}
// strictly ascending doc ids inside every column: short columns one thread each ...
__global__ void validate_sorted_kernel(const int32_t *__restrict__ indices,
                                       const int64_t *__restrict__ indptr, int64_t n_vocab, int *flags) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_vocab; t += stride) {
        int64_t s = indptr[t], e = indptr[t + 1];
        if (e - s <= 64)
            for (int64_t j = s + 1; j < e; j++)
                if (indices[j] <= indices[j - 1]) bad |= 8;
    }
    if (bad) atomicOr(flags, bad);
}
__global__ void validate_sorted_long_kernel(const int32_t *__restrict__ indices,
                                            const int64_t *__restrict__ indptr, int64_t n_vocab,
                                            int *flags) {
    // ... long columns one block each
    for (int64_t t = blockIdx.x; t < n_vocab; t += gridDim.x) {
        int64_t s = indptr[t], e = indptr[t + 1];
        if (e - s <= 64) continue;
        int bad = 0;
        for (int64_t j = s + 1 + threadIdx.x; j < e; j += blockDim.x)
            if (indices[j] <= indices[j - 1]) bad = 8;
        if (bad) atomicOr(flags, bad);
    }
}

// ---------------------------------------------------------------------------------
// kth[t] = k-th largest posting value of term t (0 when df < k): a valid lower bound
// on the k-th best score of any query containing t, because posting values are >= 0
// and fp32 addition of non-negative terms is monotone.  MSB-first radix select on the
// fp32 bit pattern, one block per term, 4 passes of 8 bits.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kth_value_kernel(const float *__restrict__ data,
                                                        const int64_t *__restrict__ indptr,
                                                        int64_t n_vocab, int k, float *__restrict__ kth) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_mask, s_remaining;
    for (int64_t t = blockIdx.x; t < n_vocab; t += gridDim.x) {
        int64_t s = indptr[t], e = indptr[t + 1];
        if (e - s < k) {
            if (threadIdx.x == 0) kth[t] = 0.0f;
            continue;
        }
        if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_remaining = (unsigned)k; }
        __syncthreads();
        for (int pass = 0; pass < 4; pass++) {
            int shift = 24 - 8 * pass;
            hist[threadIdx.x] = 0;
            __syncthreads();
            unsigned prefix = s_prefix, mask = s_mask;
            // block-uniform trip count so the warp-wide match below is convergent
            for (int64_t base = s; base < e; base += blockDim.x) {
                int64_t j = base + threadIdx.x;
                unsigned v = j < e ? __float_as_uint(data[j]) : 0u;
                bool in = j < e && (v & mask) == prefix;
                unsigned digit = (v >> shift) & 255u;
                // warp-aggregate equal digits (values of one term cluster heavily)
                unsigned key = in ? digit : 0xFFFFFFFFu;
                unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
                if (in && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31))
                    atomicAdd(&hist[digit], (unsigned)__popc(peers));
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned rem = s_remaining, cum = 0;
                int d = 255;
                for (; d > 0; d--) {
                    if (cum + hist[d] >= rem) break;
                    cum += hist[d];
                }
                s_remaining = rem - cum;
                s_prefix = prefix | ((unsigned)d << shift);
                s_mask = mask | (255u << shift);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) kth[t] = __uint_as_float(s_prefix);
        __syncthreads();
    }
}

// caller holds idx->mu
int ensure_tile_table(bb25_index *idx, cudaStream_t st) {
    if (idx->tile_off) return 0;
    const size_t tt = (size_t)idx->n_vocab * (size_t)(idx->n_tiles + 1);
    BB25_CUDA(cudaMalloc(&idx->tile_off, tt * sizeof(uint32_t)));
    idx->device_bytes += tt * sizeof(uint32_t);
    const int64_t blocks = (int64_t)((tt + 255) / 256);
    build_tile_table_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx->indices, idx->indptr, idx->n_vocab, idx->n_tiles,
                                                             idx->tile_docs, idx->tile_off);
    BB25_LAUNCH_CHECK();
    return 0;
}

// Lookup rows (see bb25_index::lookup_vals): built once, on the first call that can use them.  Terms with
// df >= n_docs / DIV that have no hot row, most frequent first, as many as fit the budget: at most
// BB25_LOOKUP_MAX_GB (default: 45 % of the free device memory, leaving 6 GB).  BB25_LOOKUP_DIV=0: none.
// Failing to allocate is not an error: row_slot then equals dense_slot and callers take their other paths.
int ensure_lookup_rows(bb25_index *idx, cudaStream_t st) {
    if (idx->lookup_tried) return 0;
    idx->lookup_tried = true;
    const int64_t n_vocab = idx->n_vocab;
    std::vector<int32_t> h_slot((size_t)n_vocab, -1);
    if (idx->dense_slot)
        BB25_CUDA(cudaMemcpy(h_slot.data(), idx->dense_slot, (size_t)n_vocab * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (!idx->dense_vals) std::fill(h_slot.begin(), h_slot.end(), -1);
    int div = 64;
    if (const char *e = getenv("BB25_LOOKUP_DIV")) div = atoi(e);
    std::vector<int32_t> h_terms;
    if (div > 0 && idx->dense_stride > 0) {
        std::vector<int64_t> h_indptr((size_t)n_vocab + 1);
        BB25_CUDA(cudaMemcpy(h_indptr.data(), idx->indptr, h_indptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        std::vector<std::pair<int64_t, int32_t>> cand;
        for (int64_t t = 0; t < n_vocab; t++) {
            const int64_t df = h_indptr[t + 1] - h_indptr[t];
            if (df > 0 && h_slot[t] < 0 && df * (int64_t)div >= idx->n_docs) cand.emplace_back(-df, (int32_t)t);
        }
        std::sort(cand.begin(), cand.end());
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        double budget = 0.45 * (double)free_b;
        if ((double)free_b - budget < 6e9) budget = (double)free_b - 6e9;
        if (const char *e = getenv("BB25_LOOKUP_MAX_GB")) budget = std::min(budget, atof(e) * 1e9);
        const double row_b = (double)idx->dense_stride * sizeof(float);
        size_t n_rows = budget > 0 ? (size_t)(budget / row_b) : 0;
        n_rows = std::min(n_rows, std::min(cand.size(), (size_t)32768));
        if (n_rows > 0) {
            const size_t nb = n_rows * (size_t)idx->dense_stride * sizeof(float);
            if (cudaMalloc(&idx->lookup_vals, nb) != cudaSuccess) {
                cudaGetLastError();
                idx->lookup_vals = nullptr;
                n_rows = 0;
            } else {
                for (size_t i = 0; i < n_rows; i++) {
                    h_slot[cand[i].second] = idx->n_dense + (int32_t)i;
                    h_terms.push_back(cand[i].second);
                }
                fill_u32_kernel<<<(unsigned)((nb / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<unsigned int *>(idx->lookup_vals), nb / 4,
                                                                                 0x80000000u);
                int32_t *d_terms = nullptr;
                BB25_CUDA(cudaMalloc(&d_terms, h_terms.size() * sizeof(int32_t)));
                BB25_CUDA(cudaMemcpyAsync(d_terms, h_terms.data(), h_terms.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                dim3 grid(32, (unsigned)n_rows);
                build_dense_rows_kernel<<<grid, 256, 0, st>>>(idx->data, idx->indices, idx->indptr, d_terms, idx->dense_stride,
                                                             idx->lookup_vals);
                count_launch(2);
                BB25_CUDA(cudaGetLastError());
                BB25_CUDA(cudaStreamSynchronize(st));
                cudaFree(d_terms);
                idx->device_bytes += nb;
            }
        }
        idx->n_lookup = (int)n_rows;
    }
    BB25_CUDA(cudaMalloc(&idx->row_slot, (size_t)std::max<int64_t>(n_vocab, 1) * sizeof(int32_t)));
    BB25_CUDA(cudaMemcpyAsync(idx->row_slot, h_slot.data(), (size_t)n_vocab * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    BB25_CUDA(cudaStreamSynchronize(st));
    idx->device_bytes += (size_t)n_vocab * sizeof(int32_t);
    return 0;
}

int get_kth_values(bb25_index *idx, int k, cudaStream_t st, const float **out) {
    auto it = idx->kth_cache.find(k);
    if (it != idx->kth_cache.end()) {
        *out = it->second;
        return 0;
    }
    float *buf = nullptr;
    if (idx->kth_cache.size() >= 8) {
        // bounded cache: drop the entry for the largest k other than 1 (k = 1 holds the global term maxima)
        auto victim = idx->kth_cache.end();
        for (auto jt = idx->kth_cache.begin(); jt != idx->kth_cache.end(); ++jt)
            if (jt->first != 1) victim = jt;
        if (victim != idx->kth_cache.end()) {
            BB25_CUDA(cudaStreamSynchronize(st));
            cudaFree(victim->second);
            idx->kth_cache.erase(victim);
            idx->device_bytes -= sizeof(float) * (size_t)idx->n_vocab;
        }
    }
    BB25_CUDA(cudaMalloc(&buf, sizeof(float) * (size_t)idx->n_vocab));
    idx->device_bytes += sizeof(float) * (size_t)idx->n_vocab;
    int grid = (int)(idx->n_vocab < (int64_t)idx->sm_count * 8 ? idx->n_vocab : (int64_t)idx->sm_count * 8);
    if (grid < 1) grid = 1;
    kth_value_kernel<<<grid, 256, 0, st>>>(idx->data, idx->indptr, idx->n_vocab, k, buf);
    BB25_LAUNCH_CHECK();
    idx->kth_cache[k] = buf;
    *out = buf;
    return 0;
}

}  // namespace bb25

using namespace bb25;

extern "C" {

const char *bb25_last_error(void) { return bb25::g_err; }
int bb25_version(void) { return BB25_VERSION; }
unsigned long long bb25_launch_count(void) { return bb25::g_launches.load(); }
int bb25_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int pick_tile_docs() {
    const char *e = getenv("BB25_TILE_DOCS");
    if (e) {
        int v = atoi(e);
        if (v == 8192 || v == 16384 || v == 32768) return v;
    }
    return 8192;  // 6 resident CTAs x 256 threads per SM in retrieve mode: best measured (profiles/r01)
}

int bb25_index_create(int device, int64_t n_docs, int64_t n_vocab, int64_t nnz, const float *data,
                      const int32_t *indices, const int64_t *indptr, const int32_t *doc_len,
                      double avgdl, int64_t doc_id_offset, bb25_index **out) {
    if (!out) { set_error("out is NULL"); return 1; }
    *out = nullptr;
    if (n_docs < 1 || n_docs > (int64_t)kIdMask + 1) {
        set_error("n_docs must be in [1, 2^29], got %lld", (long long)n_docs);
        return 1;
    }
    if (n_vocab < 1 || nnz < 0 || !indptr || !doc_len || (nnz > 0 && (!data || !indices))) {
        set_error("bad index arguments");
        return 1;
    }
    if (!(avgdl > 0.0)) { set_error("avgdl must be > 0"); return 1; }
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select CUDA device %d", device); return 1; }

    bb25_index *idx = new bb25_index();
    idx->device = device;
    idx->n_docs = n_docs;
    idx->n_vocab = n_vocab;
    idx->nnz = nnz;
    idx->avgdl = avgdl;
    idx->doc_id_offset = doc_id_offset;
    auto fail = [&]() { bb25_index_destroy(idx); return 1; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return fail(); }
    idx->sm_count = prop.multiProcessorCount;

    size_t nn = (size_t)(nnz > 0 ? nnz : 1);
#define TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return fail(); } } while (0)
    TRY(cudaMalloc(&idx->data, nn * sizeof(float)));
    TRY(cudaMalloc(&idx->indices, nn * sizeof(int32_t)));
    TRY(cudaMalloc(&idx->indptr, (size_t)(n_vocab + 1) * sizeof(int64_t)));
    TRY(cudaMalloc(&idx->doc_len, (size_t)n_docs * sizeof(int32_t)));
    idx->device_bytes = nn * 8 + (size_t)(n_vocab + 1) * 8 + (size_t)n_docs * 4;
    if (nnz > 0) {
        TRY(cudaMemcpy(idx->data, data, (size_t)nnz * sizeof(float), cudaMemcpyDefault));
        TRY(cudaMemcpy(idx->indices, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyDefault));
    }
    TRY(cudaMemcpy(idx->indptr, indptr, (size_t)(n_vocab + 1) * sizeof(int64_t), cudaMemcpyDefault));
    TRY(cudaMemcpy(idx->doc_len, doc_len, (size_t)n_docs * sizeof(int32_t), cudaMemcpyDefault));
    TRY(cudaMallocHost(&idx->pinned, 4096));

    // validate
    int *flags = nullptr;
    TRY(cudaMalloc(&flags, sizeof(int)));
    TRY(cudaMemset(flags, 0, sizeof(int)));
    validate_csc_kernel<<<idx->sm_count * 4, 256>>>(idx->data, idx->indices, idx->indptr, n_vocab, n_docs, nnz, flags);
    count_launch();
    int hflags = 0;
    TRY(cudaMemcpy(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (hflags == 0) {
        validate_sorted_kernel<<<idx->sm_count * 4, 256>>>(idx->indices, idx->indptr, n_vocab, flags);
        validate_sorted_long_kernel<<<idx->sm_count * 4, 256>>>(idx->indices, idx->indptr, n_vocab, flags);
        count_launch(2);
        TRY(cudaMemcpy(&hflags, flags, sizeof(int), cudaMemcpyDeviceToHost));
    }
    cudaFree(flags);
    if (hflags) {
        set_error("invalid CSC input (flags=%d: 1 indptr, 2 doc id out of range, 4 negative/NaN value, 8 doc ids not ascending)", hflags);
        return fail();
    }

    // the per-(term, tile) offsets of the dense-output kernel are built on first use (ensure_tile_table):
    // V x (N/8192 + 1) words -- 4 GB for a million-term vocabulary -- that batch retrieval never touches
    idx->tile_docs = pick_tile_docs();
    idx->n_tiles = (int)((n_docs + idx->tile_docs - 1) / idx->tile_docs);
    {
        idx->n_blocks = (int)((n_docs + kBlockDocs - 1) / kBlockDocs);
        // dense rows for every term while the table fits the budget; otherwise (or with
        // BB25_TAB_SPARSE=1: wherever smaller, =2: everywhere -- test hooks) the bitmap form
        const int n_words = (idx->n_blocks + 31) / 32;
        const int grid = (int)(n_vocab < (int64_t)idx->sm_count * 16 ? n_vocab : (int64_t)idx->sm_count * 16);
        int32_t *d_touched = nullptr;
        TRY(cudaMalloc(&d_touched, (size_t)n_vocab * sizeof(int32_t)));
        count_term_blocks_kernel<<<grid, 256>>>(idx->indices, idx->indptr, n_vocab, d_touched);
        count_launch();
        std::vector<int32_t> h_touched((size_t)n_vocab);
        TRY(cudaMemcpy(h_touched.data(), d_touched, (size_t)n_vocab * sizeof(int32_t), cudaMemcpyDeviceToHost));
        cudaFree(d_touched);
        size_t free_b = 0, total_b = 0;
        TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t dense_bytes = (size_t)idx->n_blocks * (size_t)n_vocab * sizeof(uint2);
        int sparse_mode = dense_bytes > std::min<size_t>(total_b / 8, free_b / 4) ? 1 : 0;
        if (const char *e = getenv("BB25_TAB_SPARSE")) sparse_mode = atoi(e);
        std::vector<longlong2> h_row((size_t)n_vocab);
        size_t n_ent = 0, n_bits = 0;
        idx->tab_sparse_terms = 0;
        for (int64_t t = 0; t < n_vocab; t++) {
            const size_t sparse_cost = (size_t)n_words + (size_t)h_touched[t];
            const bool sparse = sparse_mode >= 2 || (sparse_mode == 1 && sparse_cost * 2 < (size_t)idx->n_blocks);
            h_row[t].x = (long long)n_ent;
            if (sparse) {
                h_row[t].y = (long long)n_bits;
                n_bits += (size_t)n_words;
                n_ent += (size_t)h_touched[t];
                idx->tab_sparse_terms++;
            } else {
                h_row[t].y = -1;
                n_ent += (size_t)idx->n_blocks;
            }
        }
        TRY(cudaMalloc(&idx->tab_row, (size_t)n_vocab * sizeof(longlong2)));
        TRY(cudaMemcpy(idx->tab_row, h_row.data(), (size_t)n_vocab * sizeof(longlong2), cudaMemcpyHostToDevice));
        TRY(cudaMalloc(&idx->tab_ent, std::max<size_t>(n_ent, 1) * sizeof(uint2)));
        TRY(cudaMalloc(&idx->tab_bits, std::max<size_t>(n_bits, 1) * sizeof(uint2)));
        idx->tab_bytes = (size_t)n_vocab * sizeof(longlong2) + (n_ent + n_bits) * sizeof(uint2);
        idx->device_bytes += idx->tab_bytes;
        if (n_ent) init_block_table_kernel<<<(unsigned)((n_ent + 255) / 256), 256>>>(idx->tab_ent, n_ent);
        TRY(cudaMemset(idx->tab_bits, 0, std::max<size_t>(n_bits, 1) * sizeof(uint2)));
        if (n_bits) {
            set_block_bits_kernel<<<grid, 256>>>(idx->indices, idx->indptr, n_vocab, idx->tab_row, idx->tab_bits);
            scan_block_bits_kernel<<<(unsigned)((n_vocab * 32 + 255) / 256), 256>>>(n_vocab, n_words, idx->tab_row, idx->tab_bits);
            count_launch(2);
        }
        build_block_table_kernel<<<grid, 256>>>(idx->data, idx->indices, idx->indptr, n_vocab, 0, idx->tab_row, idx->tab_bits, idx->tab_ent);
        build_block_table_kernel<<<grid, 256>>>(idx->data, idx->indices, idx->indptr, n_vocab, 1, idx->tab_row, idx->tab_bits, idx->tab_ent);
        count_launch(3);
        TRY(cudaGetLastError());
        TRY(cudaDeviceSynchronize());
        if (const char *e = getenv("BB25_PRUNE")) {
            const int v = atoi(e);
            idx->prune = v < 0 ? 0 : (v > 3 ? 3 : v);
        }
    }
    {
        // dense value rows for the head terms
        std::vector<int64_t> h_indptr((size_t)n_vocab + 1);
        TRY(cudaMemcpy(h_indptr.data(), idx->indptr, h_indptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        std::vector<std::pair<int64_t, int32_t>> cand;
        for (int64_t t = 0; t < n_vocab; t++) {
            const int64_t df = h_indptr[t + 1] - h_indptr[t];
            if (df > 0 && df * 8 >= n_docs) cand.emplace_back(-df, (int32_t)t);
        }
        std::sort(cand.begin(), cand.end());
        if ((int)cand.size() > kMaxDenseTerms) cand.resize(kMaxDenseTerms);
        std::vector<int32_t> h_slot((size_t)n_vocab, -1);
        std::vector<int32_t> h_terms;
        for (size_t i = 0; i < cand.size(); i++) {
            h_slot[cand[i].second] = (int32_t)i;
            h_terms.push_back(cand[i].second);
        }
        idx->n_dense = (int)cand.size();
        idx->dense_stride = (int64_t)idx->n_blocks * kBlockDocs;
        TRY(cudaMalloc(&idx->dense_slot, (size_t)n_vocab * sizeof(int32_t)));
        TRY(cudaMemcpy(idx->dense_slot, h_slot.data(), (size_t)n_vocab * sizeof(int32_t), cudaMemcpyHostToDevice));
        idx->device_bytes += (size_t)n_vocab * sizeof(int32_t);
        if (idx->n_dense > 0) {
            const size_t nb = (size_t)idx->n_dense * (size_t)idx->dense_stride * sizeof(float);
            TRY(cudaMalloc(&idx->dense_vals, nb));
            // absent documents hold -0.0f: x + (-0.0f) == x exactly, and the sign bit keeps
            // "absent" distinguishable from a posting whose value is +0.0f
            fill_u32_kernel<<<(unsigned)((nb / 4 + 255) / 256), 256>>>(reinterpret_cast<unsigned int *>(idx->dense_vals), nb / 4, 0x80000000u);
            count_launch();
            idx->device_bytes += nb;
            int32_t *d_terms = nullptr;
            TRY(cudaMalloc(&d_terms, h_terms.size() * sizeof(int32_t)));
            TRY(cudaMemcpy(d_terms, h_terms.data(), h_terms.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            dim3 grid(128, (unsigned)idx->n_dense);
            build_dense_rows_kernel<<<grid, 256>>>(idx->data, idx->indices, idx->indptr, d_terms, idx->dense_stride,
                                                   idx->dense_vals);
            count_launch();
            TRY(cudaGetLastError());
            TRY(cudaDeviceSynchronize());
            cudaFree(d_terms);
            // fp16 upper-bound copy for the order-free pass (BB25_HALF_ROWS=0 at index creation: none)
            const char *eh = getenv("BB25_HALF_ROWS");
            if (!(eh && atoi(eh) == 0)) {
                const size_t ne = (size_t)idx->n_dense * (size_t)idx->dense_stride;
                int *hflag = nullptr;
                TRY(cudaMalloc(&hflag, sizeof(int)));
                TRY(cudaMemset(hflag, 0, sizeof(int)));
                TRY(cudaMalloc(&idx->dense_h, ne * sizeof(__half)));
                half_rows_kernel<<<(unsigned)((ne + 255) / 256), 256>>>(idx->dense_vals, ne, idx->dense_h, hflag);
                count_launch();
                int hf = 0;
                TRY(cudaMemcpy(&hf, hflag, sizeof(int), cudaMemcpyDeviceToHost));
                cudaFree(hflag);
                if (hf) {  // values beyond the fp16 range: keep the fp32 rows only
                    cudaFree(idx->dense_h);
                    idx->dense_h = nullptr;
                } else {
                    idx->device_bytes += ne * sizeof(__half);
                }
            }
        }
    }
#undef TRY
    *out = idx;
    return 0;
}

void bb25_index_destroy(bb25_index *idx) {
    if (!idx) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(idx->device);
    cudaFree(idx->data);
    cudaFree(idx->indices);
    cudaFree(idx->indptr);
    cudaFree(idx->doc_len);
    cudaFree(idx->tile_off);
    cudaFree(idx->tab_ent);
    cudaFree(idx->tab_bits);
    cudaFree(idx->tab_row);
    cudaFree(idx->dense_slot);
    cudaFree(idx->dense_vals);
    cudaFree(idx->dense_h);
    cudaFree(idx->lookup_vals);
    cudaFree(idx->row_slot);
    for (auto &kv : idx->kth_cache) cudaFree(kv.second);
    if (idx->ws) cudaFree(idx->ws);
    for (int i = 0; i < 2; i++) {
        if (idx->hs_dev[i]) cudaFree(idx->hs_dev[i]);
        if (idx->hs_copy[i]) cudaStreamDestroy(idx->hs_copy[i]);
        if (idx->hs_ev[i]) cudaEventDestroy(idx->hs_ev[i]);
    }
    if (idx->hs_stream) cudaStreamDestroy(idx->hs_stream);
    if (idx->ws_ev) cudaEventDestroy(idx->ws_ev);
    if (idx->pinned) cudaFreeHost(idx->pinned);
    for (int i = 0; i < idx->n_ev; i++) cudaEventDestroy(idx->ev[i]);
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    delete idx;
}

int bb25_index_kth_values(bb25_index *idx, int k, const float **out_dev, void *stream) {
    if (!idx || !out_dev || k < 1) { set_error("bad arguments"); return 1; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    return get_kth_values(idx, k, (cudaStream_t)stream, out_dev);
}

int bb25_index_set_kth_values(bb25_index *idx, int k, const float *values_dev, void *stream) {
    if (!idx || !values_dev || k < 1) { set_error("bad arguments"); return 1; }
    DeviceGuard g(idx->device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    std::lock_guard<std::mutex> lock(idx->mu);
    const float *cur = nullptr;
    if (get_kth_values(idx, k, (cudaStream_t)stream, &cur)) return 1;
    BB25_CUDA(cudaMemcpyAsync(const_cast<float *>(cur), values_dev, sizeof(float) * (size_t)idx->n_vocab, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
    return 0;
}

int bb25_index_info(const bb25_index *idx, int64_t *n_docs, int64_t *n_vocab, int64_t *nnz,
                    int *tile_docs, int *n_tiles, int64_t *device_bytes) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (n_docs) *n_docs = idx->n_docs;
    if (n_vocab) *n_vocab = idx->n_vocab;
    if (nnz) *nnz = idx->nnz;
    if (tile_docs) *tile_docs = idx->tile_docs;
    if (n_tiles) *n_tiles = idx->n_tiles;
    if (device_bytes) *device_bytes = (int64_t)(idx->device_bytes + idx->ws_bytes);
    return 0;
}

int bb25_index_table_info(const bb25_index *idx, int64_t *bitmap_terms, int64_t *table_bytes) {
    if (!idx) { set_error("index is NULL"); return 1; }
    if (bitmap_terms) *bitmap_terms = idx->tab_sparse_terms;
    if (table_bytes) *table_bytes = (int64_t)idx->tab_bytes;
    return 0;
}

}  // extern "C"
