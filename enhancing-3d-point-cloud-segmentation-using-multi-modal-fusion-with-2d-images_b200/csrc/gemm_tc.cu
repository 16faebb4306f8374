// Tensor-core contraction for KPConv on sm_100a: tcgen05.mma (kind::f16, bf16 operands, fp32
// accumulators in TMEM) fed by TMA (cp.async.bulk.tensor, 128B-swizzled tiles) through a 3-stage
// mbarrier pipeline, warp-specialised:
//     warp 0      TMA producer (one elected lane)
//     warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//     warps 2..5  epilogue: tcgen05.ld 32x32b -> registers -> global (store or fp32 atomics)
//
// fp32-grade accuracy from bf16 tensor cores: every fp32 operand v is carried as the pair
// hi = bf16(v), lo = bf16(v - hi); the product is accumulated as hi*hi + lo*hi + hi*lo in the
// same fp32 TMEM accumulator ("bf16x3", relative error ~2^-16, inside the 1e-4 parity budget).
// terms = 1 runs the plain bf16 product.
//
// Operand layouts (all row-major in global memory, bf16):
//   K-major  : X[rows, K]  K contiguous   -> one TMA box {64 k, R rows}, canonical SW128 K-major
//   MN-major : X[K, cols]  MN contiguous  -> one TMA box {64 mn, 64 k} per 64 columns, canonical
//                                            SW128 MN-major (LBO = 8 KB between column chunks)
// so the forward product (A K-major, W MN-major), dX-side product (both K-major) and the dW product
// (both MN-major, split over K with atomics) all run through this one kernel without transposes.
#include <cuda.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 3;
constexpr int TILE_BYTES = 16384;              // 128 rows x 128 B
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi | A_lo | B_hi | B_lo
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 128;

struct GemmParams {
    float* D;
    int M, ldd, n_valid, bn, kb_total, kb_per_split, terms, a_mn, b_mn, atomic_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int mn_major) {
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
    // version=1 [46,48), layout_type SWIZZLE_128B=2 [61,64).  SBO = 1024 B (8 rows x 128 B);
    // LBO = 8192 B between 64-element MN chunks (MN-major), unused (1) for swizzled K-major.
    uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((mn_major ? 512u : 1u) << 16);
    uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               GemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SW128 tiles need 1 KB alignment
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[STAGES]);
    const uint32_t bar_tmem = smem_u32(&bars[2 * STAGES]);

    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * p.bn;
    const int kb0 = blockIdx.z * p.kb_per_split;
    const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_smem)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t a_bytes = TILE_BYTES;
            const uint32_t b_bytes = (uint32_t)p.bn * 128u;
            const uint32_t stage_tx = (a_bytes + b_bytes) * (p.terms == 3 ? 2u : 1u);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, stage_tx);
                const uint32_t st = smem_base + s * STAGE_BYTES;
                const int k0 = kb * BK;
                for (int half = 0; half < (p.terms == 3 ? 2 : 1); half++) {
                    const CUtensorMap* ma = half ? &tm_a_lo : &tm_a_hi;
                    const CUtensorMap* mb = half ? &tm_b_lo : &tm_b_hi;
                    const uint32_t sa = st + half * TILE_BYTES;
                    const uint32_t sb = st + (2 + half) * TILE_BYTES;
                    if (!p.a_mn) {
                        tma_load_2d(sa, ma, full, k0, m0);
                    } else {
                        tma_load_2d(sa, ma, full, m0, k0);
                        tma_load_2d(sa + 8192, ma, full, m0 + 64, k0);
                    }
                    if (!p.b_mn) {
                        tma_load_2d(sb, mb, full, k0, n0);
                    } else {
                        for (int c = 0; c < p.bn / 64; c++) tma_load_2d(sb + 8192 * c, mb, full, n0 + 64 * c, k0);
                    }
                }
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // InstrDescriptor: c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13), a_major 15, b_major 16,
            // N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) |
                                   ((uint32_t)p.b_mn << 16) | ((uint32_t)(p.bn >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            const uint32_t a_kstep = p.a_mn ? 2048u : 32u;  // 16 k-elements: 2 SBO groups vs 32 B
            const uint32_t b_kstep = p.b_mn ? 2048u : 32u;
            int s = 0;
            uint32_t ph = 0;
            uint32_t acc = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_base + s * STAGE_BYTES;
                for (int term = 0; term < p.terms; term++) {
                    const uint32_t sa = st + (term == 1 ? TILE_BYTES : 0);
                    const uint32_t sb = st + (term == 2 ? 3 : 2) * TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        umma_bf16(tmem_base, make_desc(sa + k * a_kstep, p.a_mn),
                                  make_desc(sb + k * b_kstep, p.b_mn), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            umma_commit(bar_tmem);  // accumulator complete
        }
    } else {
        // ===================== epilogue =====================
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;  // TMEM lanes [32q, 32q+32) are only visible to warps with id%4 == q
        const int row = m0 + quarter * 32 + lane;
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            if (row < p.M) {
                float* dst = p.D + (size_t)row * p.ldd + n0 + c0;
                const int ncols = min(32, p.n_valid - (n0 + c0));
                if (p.atomic_out) {
                    for (int j = 0; j < ncols; j++) atomicAdd(dst + j, __uint_as_float(v[j]));
                } else if (ncols == 32 && ((((size_t)dst) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *(float4*)(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else {
                    for (int j = 0; j < ncols; j++) dst[j] = __uint_as_float(v[j]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D bf16 tensor map: inner (contiguous) extent d0, outer extent d1, row pitch ld elements.
int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box0, uint32_t box1) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled entry point unavailable");
        return MVK_ERR_CUDA;
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return MVK_ERR_CUDA;
    }
    return MVK_OK;
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" int mvk_gemm_bf16x3(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                               const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                               int n_valid, int terms, int split_k, mvk_stream_t stream) {
    if (!a_hi || !b_hi || !D || M < 1 || N < 64 || K < 1 || (N % 64) != 0 || (lda % 8) != 0 || (ldb % 8) != 0 ||
        n_valid < 1 || n_valid > N || ldd < n_valid || (terms != 1 && terms != 3))
        return MVK_ERR_INVALID_ARG;
    if (terms == 3 && (!a_lo || !b_lo)) return MVK_ERR_INVALID_ARG;
    if ((((size_t)a_hi | (size_t)b_hi | (size_t)(a_lo ? a_lo : a_hi) | (size_t)(b_lo ? b_lo : b_hi)) & 15) != 0)
        return MVK_ERR_INVALID_ARG;
    const int bn = (N % 128 == 0) ? 128 : 64;
    GemmParams p;
    p.D = D;
    p.M = M;
    p.ldd = ldd;
    p.n_valid = n_valid;
    p.bn = bn;
    p.kb_total = (K + BK - 1) / BK;
    if (split_k < 1) split_k = 1;
    if (split_k > p.kb_total) split_k = p.kb_total;
    p.kb_per_split = (p.kb_total + split_k - 1) / split_k;
    int splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.terms = terms;
    p.a_mn = a_mn_major ? 1 : 0;
    p.b_mn = b_mn_major ? 1 : 0;
    p.atomic_out = splits > 1 ? 1 : 0;

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    const void* al = a_lo ? a_lo : a_hi;
    const void* bl = b_lo ? b_lo : b_hi;
    if (!p.a_mn) {
        if ((rc = make_map(&ma_hi, a_hi, K, M, lda, 64, 128))) return rc;
        if ((rc = make_map(&ma_lo, al, K, M, lda, 64, 128))) return rc;
    } else {
        if ((rc = make_map(&ma_hi, a_hi, M, K, lda, 64, 64))) return rc;
        if ((rc = make_map(&ma_lo, al, M, K, lda, 64, 64))) return rc;
    }
    if (!p.b_mn) {
        if ((rc = make_map(&mb_hi, b_hi, K, N, ldb, 64, bn))) return rc;
        if ((rc = make_map(&mb_lo, bl, K, N, ldb, 64, bn))) return rc;
    } else {
        if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, 64, 64))) return rc;
        if ((rc = make_map(&mb_lo, bl, N, K, ldb, 64, 64))) return rc;
    }
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    MVK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((M + BM - 1) / BM, N / bn, splits);
    gemm_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    MVK_LAUNCHED("gemm_tc_kernel");
    return MVK_OK;
}
