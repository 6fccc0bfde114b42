// Dense side of hybrid retrieval: cos[q][d] = <query_emb[q], corpus_emb[d]> for a batch of
// queries against every document -- the `query_emb @ corpus_emb.T` of benchmarks/hybrid_beir.py:1751-1753 --
// as a hand-written Blackwell GEMM: TMA (cp.async.bulk.tensor, 128-byte swizzle) stages bf16 tiles in shared
// memory, one elected thread issues tcgen05.mma (UMMA 128 x N x 16, cta_group::1) with the fp32 accumulator in
// tensor memory, four epilogue warps read it back with tcgen05.ld and store fp32 cosines row-per-query, the
// layout bb25_retrieve_fused_batch consumes.  This IS a dense contraction, the one place on the path where
// tensor cores belong.
//
//   M (UMMA rows, TMEM lanes) = 128 documents of a tile          A operand: corpus tile  [128][64] bf16, K-major
//   N (UMMA columns)          = the query sub-batch, <= 256      B operand: query tile   [N][64]   bf16, K-major
//   K                         = embedding width, 64 per stage (one 128-byte swizzle atom), 4 x UMMA_K = 16
//
// Warp roles (192 threads, one CTA per SM, persistent over document tiles):
//   warp 0   TMA producer: per k-block one corpus box and one query box into a 4-stage ring (full/empty mbarriers)
//   warp 1   TMEM allocation; one lane issues the MMAs, tcgen05.commit releases ring slots / publishes accumulators
//   warps 2-5 epilogue: TMEM lane quadrant (warp % 4) -> registers -> coalesced global stores; the accumulator
//            is double-buffered (2 x N columns), so a tile's epilogue overlaps the next tile's MMAs.
#include <cuda.h>
#include <cuda_bf16.h>

#include "bb25_internal.cuh"

namespace bb25 {

constexpr int GM = 128;      // documents per tile
constexpr int GK = 64;       // K elements per pipeline stage
constexpr int GUK = 16;      // K per tcgen05.mma for 16-bit inputs
constexpr int GSTAGES = 4;
constexpr int GMAXN = 256;   // queries per launch
constexpr int GTHREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle, rows of 64 bf16 (= one swizzle atom wide):
//   start address >> 4 | SBO = 1024 B (8 rows x 128 B between core-matrix groups) | version 1 | layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(const void *tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);  // bits [0,14)
    d |= (uint64_t)(1024u >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= 1ull << 46;                                    // descriptor version (sm_100)
    d |= 2ull << 61;                                    // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmArgs {
    float *out;          // [n_queries][out_stride]
    int64_t out_stride;
    int64_t n_docs;
    int n_queries;       // valid query rows (<= n_cols)
    int n_cols;          // UMMA N: n_queries rounded up to a multiple of 16
    int k_blocks;        // K / 64
    int n_tiles;         // ceil(n_docs / 128)
};

__global__ void __launch_bounds__(GTHREADS, 1)
cosine_gemm_kernel(const __grid_constant__ CUtensorMap map_corpus, const __grid_constant__ CUtensorMap map_query,
                   const __grid_constant__ GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment of every operand tile (the 128-byte swizzle pattern repeats every 8 rows = 1024 bytes)
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = GM * GK * 2;            // 16 KB
    const int b_bytes = a.n_cols * GK * 2;      // <= 32 KB
    const int stage_bytes = a_bytes + GMAXN * GK * 2;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)GSTAGES * stage_bytes);
    uint64_t *full = bars, *empty = bars + GSTAGES, *tfull = bars + 2 * GSTAGES, *tempty = bars + 2 * GSTAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * GSTAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < GSTAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 accumulators x up to 256 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < a.k_blocks; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char *sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full[stage], (uint32_t)(a_bytes + b_bytes));
                    tma_load_2d(sa, &map_corpus, &full[stage], kb * GK, tile * GM);
                    tma_load_2d(sa + a_bytes, &map_query, &full[stage], kb * GK, 0);
                    if (++stage == GSTAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc(GM, a.n_cols);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * GMAXN);
                for (int kb = 0; kb < a.k_blocks; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const unsigned char *sa = smem + (size_t)stage * stage_bytes;
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + a_bytes);
#pragma unroll
                    for (int k = 0; k < GK / GUK; k++) {
                        // advancing K inside the swizzle atom: + k * 16 elements * 2 bytes, in 16-byte units
                        const uint64_t adv = (uint64_t)((k * GUK * 2) >> 4);
                        tc_mma_bf16(tmem_d, adesc + adv, bdesc + adv, idesc, (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(&empty[stage]);  // the ring slot is free once these MMAs have read it
                    if (++stage == GSTAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(&tfull[acc]);  // accumulator complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int quad = warp & 3;  // the TMEM lane quadrant this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int64_t doc = (int64_t)tile * GM + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * GMAXN);
            for (int c0 = 0; c0 < a.n_cols; c0 += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (doc < a.n_docs) {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (c0 + j < a.n_queries) a.out[(int64_t)(c0 + j) * a.out_stride + doc] = __uint_as_float(r[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 row-major [rows][k]: box = [box_rows][64], 128-byte swizzle, rows beyond the matrix read as zeros
static int make_map(CUtensorMap *map, const void *base, int64_t rows, int k, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return 1; }
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)GK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return 1; }
    return 0;
}

}  // namespace bb25

using namespace bb25;

extern "C" int bb25_cosine_gemm(int device, const void *query_emb, int n_queries, const void *corpus_emb, int64_t n_docs,
                                int k, float *out, int64_t out_stride, void *stream) {
    if (!query_emb || !corpus_emb || !out || n_queries < 1 || n_queries > GMAXN || n_docs < 1 || out_stride < n_docs) {
        set_error("bad arguments (1 <= n_queries <= %d per call, out_stride >= n_docs)", GMAXN);
        return 1;
    }
    if (k < GK || (k % GK) != 0) { set_error("the embedding width must be a multiple of %d, got %d", GK, k); return 1; }
    if (((uintptr_t)query_emb & 15) || ((uintptr_t)corpus_emb & 15)) { set_error("embedding matrices must be 16-byte aligned"); return 1; }
    if (bb25_device_count() < 1) { set_error("no CUDA device available (libbb25 has no CPU fallback)"); return 1; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cannot select device"); return 1; }
    cudaDeviceProp prop;
    BB25_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("bb25_cosine_gemm needs an sm_100 device (tcgen05 / TMEM)"); return 1; }
    GemmArgs a{};
    a.out = out;
    a.out_stride = out_stride;
    a.n_docs = n_docs;
    a.n_queries = n_queries;
    a.n_cols = (n_queries + 15) & ~15;
    a.k_blocks = k / GK;
    a.n_tiles = (int)((n_docs + GM - 1) / GM);
    CUtensorMap mc, mq;
    if (make_map(&mc, corpus_emb, n_docs, k, GM)) return 1;
    // the query box always spans n_cols rows; rows beyond n_queries are out of bounds and read as zeros
    if (make_map(&mq, query_emb, n_queries, k, a.n_cols)) return 1;
    const size_t smem = (size_t)GSTAGES * (GM * GK * 2 + GMAXN * GK * 2) + 256 + 1024;
    BB25_CUDA(cudaFuncSetAttribute(cosine_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.n_tiles < prop.multiProcessorCount ? a.n_tiles : prop.multiProcessorCount;
    cosine_gemm_kernel<<<grid, GTHREADS, smem, (cudaStream_t)stream>>>(mc, mq, a);
    BB25_LAUNCH_CHECK();
    return 0;
}
