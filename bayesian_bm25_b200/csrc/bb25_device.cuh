// Device helpers shared by the traversal translation units (bb25_traverse.cu, bb25_fused.cu).
#pragma once

#include "bb25_internal.cuh"

namespace bb25 {

__device__ __forceinline__ float4 ld_nc_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// Dense value rows are shared by every warp of the SM working on the same block (items are
// block-major), so unlike the posting streams they are worth keeping in L1.
__device__ __forceinline__ float4 ld_row_f4(const float4 *p) {
#if defined(BB25_ROW_NOALLOC)
    return ld_nc_f4(p);
#else
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#endif
}

__device__ __forceinline__ int ld_nc_s32(const int32_t *p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_nc_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    int lo = __shfl_sync(0xFFFFFFFFu, (int)(v & 0xFFFFFFFFll), src);
    int hi = __shfl_sync(0xFFFFFFFFu, (int)(v >> 32), src);
    return ((long long)hi << 32) | (unsigned int)lo;
}

// acc[o] (+)= v.  FRESH: the accumulators are known to be all zero (first term of the
// unit), so the read half of the read-modify-write is dropped; v + 0.0f keeps -0.0f out.
template <bool FRESH>
__device__ __forceinline__ void acc_add(float *acc, int o, float v) {
    acc[o] = FRESH ? __fadd_rn(v, 0.0f) : __fadd_rn(acc[o], v);
}

// postings [s, s+len) of one term into the warp's block accumulators, up to 64 postings per
// round of loads (no alignment peel: a short slice must not cost two dependent rounds)
template <bool FRESH>
__device__ __forceinline__ void scatter_block(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                              long long s, int len, float *acc, int doc_base, int lane) {
    const int32_t *ip = indices + s;
    const float *dp = data + s;
    for (int j0 = 0; j0 < len; j0 += 64) {
        int d[2];
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int j = j0 + 32 * u + lane;
            if (j < len) {
                d[u] = ld_nc_s32(ip + j);
                v[u] = ld_nc_f32(dp + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++)
            if (j0 + 32 * u + lane < len) acc_add<FRESH>(acc, d[u] - doc_base, v[u]);
    }
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

// fp32 x 2 add (Blackwell FADD2): the same round-to-nearest result per element as two FADDs
__device__ __forceinline__ void add_f4(float4 &v, const float4 &r) {
    unsigned long long a0, a1, b0, b1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(v.z), "f"(v.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b0) : "f"(r.x), "f"(r.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b1) : "f"(r.z), "f"(r.w));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a0) : "l"(b0));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a1) : "l"(b1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.z), "=f"(v.w) : "l"(a1));
}

// Position of document `doc` in one term's slice of a 1024-document block, or -1.  ids[0 .. len) are the
// slice's doc ids (ascending, all inside [doc_base, doc_base + 1024)).  The slice is probed where a uniform
// spread would put the document, bracketed by doubling steps and bisected: two or three dependent loads,
// mostly from one cache line, instead of log2(len) ~ 10 for a plain binary search.
__host__ __device__ __forceinline__ int slice_find(const int32_t *__restrict__ ids, int len, uint32_t doc, uint32_t doc_base) {
    if (len <= 0) return -1;
    int g = (int)(((unsigned long long)(doc - doc_base) * (unsigned long long)len) >> 10);
    g = g >= len ? len - 1 : g;
    const uint32_t v = (uint32_t)ids[g];
    if (v == doc) return g;
    int lo, hi;  // the posting, if there is one, lies in [lo, hi)
    if (v < doc) {
        lo = g + 1;
        hi = len;
        int step = 1;
        while (lo < hi) {
            int p = lo + step - 1;
            if (p >= hi) p = hi - 1;
            const uint32_t w = (uint32_t)ids[p];
            if (w == doc) return p;
            if (w > doc) {
                hi = p;
                break;
            }
            lo = p + 1;
            step <<= 1;
        }
    } else {
        lo = 0;
        hi = g;
        int step = 1;
        while (lo < hi) {
            int p = hi - step;
            if (p < lo) p = lo;
            const uint32_t w = (uint32_t)ids[p];
            if (w == doc) return p;
            if (w < doc) {
                lo = p + 1;
                break;
            }
            hi = p;
            step <<= 1;
        }
    }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)ids[mid] < doc) lo = mid + 1;
        else hi = mid;
    }
    return (lo < len && (uint32_t)ids[lo] == doc) ? lo : -1;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace bb25
