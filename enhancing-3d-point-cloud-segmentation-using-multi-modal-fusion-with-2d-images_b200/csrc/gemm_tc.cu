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
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvk {
namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 8;
constexpr int A_TILE_BYTES = 16384;            // 128 rows x 128 B (one of hi / lo)
constexpr int MAX_THREADS = 320;               // 2 + up to 8 epilogue warps
constexpr int STAGING_PER_WARP = 2 * 4096;     // 2 buffers x [32 rows x 128 B] per epilogue warp
constexpr int COLSUM_BYTES = 4 * 2 * 256 * 4;  // per epilogue warp: {sum, sum of squares} x 256 columns
constexpr int SMEM_LIMIT = 232448;             // 227 KB opt-in maximum per CTA

struct GemmParams {
    float* D;
    int M, N, ldd, bn, kb_total, kb_per_split, splits, terms, a_mn, b_mn, atomic_out, stages, tma_store;
    int m_tiles, n_tiles, num_tiles, tmem_cols, stage_bytes, ew, d_swizzle;
    double* col_stats;  // optional [2 * N]: column sums / sums of squares of D (batch-norm statistics), fp64 atomics
};

// Stores one 16-row x 64-column fragment (see tmem_ld_16x256b_x8); static register indices only.
__device__ __forceinline__ void frag_store(const uint32_t (&v)[32], const GemmParams& p, int r_lo, int col0, int lane) {
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int r = r_lo + rr * 8;
        if (r < p.M) {
            float* drow = p.D + (size_t)r * p.ldd + col0 + 2 * (lane & 3);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (col0 + 8 * i + 2 * (lane & 3) + 1 < p.N) {
                    const float a = __uint_as_float(v[4 * i + 2 * rr]);
                    const float b = __uint_as_float(v[4 * i + 2 * rr + 1]);
                    if (p.atomic_out) red_add_v2(drow + 8 * i, a, b);
                    else *(float2*)(drow + 8 * i) = make_float2(a, b);
                }
            }
        }
    }
}

// Persistent, warp-specialised: every CTA (one per SM) walks the tile list t = blockIdx.x,
// blockIdx.x + gridDim.x, ...; tile -> (split, m, n) with n fastest so that the CTAs working on
// one row block share its A tile through L2.  The smem ring and the two TMEM accumulator buffers
// run across tile boundaries, so the loads / MMAs of tile i+1 overlap the epilogue of tile i.
__global__ void __launch_bounds__(MAX_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const __grid_constant__ CUtensorMap tm_d, GemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 4];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SW128 tiles need 1 KB alignment
    const uint32_t staging = smem_base + (uint32_t)p.stages * (uint32_t)p.stage_bytes;
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[MAX_STAGES]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * MAX_STAGES]);       // [2] accumulator ready
    const uint32_t bar_tempty = smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2] accumulator drained
    const int nstages = p.stages;
    const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_d) : "memory");
        for (int s = 0; s < nstages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, (uint32_t)p.ew);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_smem)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    pdl_enter();  // barriers, TMEM and tensor-map prefetch are set up while the previous kernel drains

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t stage_tx = (A_TILE_BYTES + b_tile_bytes) * (p.terms == 3 ? 2u : 1u);
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int nt = t % p.n_tiles, rest = t / p.n_tiles;
                const int mt = rest % p.m_tiles, sp = rest / p.m_tiles;
                const int m0 = mt * BM, n0 = nt * p.bn;
                const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    const uint32_t full = bar_full + 8 * s;
                    mbar_expect_tx(full, stage_tx);
                    const uint32_t st = smem_base + s * p.stage_bytes;
                    const int k0 = kb * BK;
                    for (int half = 0; half < (p.terms == 3 ? 2 : 1); half++) {
                        const CUtensorMap* ma = half ? &tm_a_lo : &tm_a_hi;
                        const CUtensorMap* mb = half ? &tm_b_lo : &tm_b_hi;
                        const uint32_t sa = st + half * A_TILE_BYTES;
                        const uint32_t sb = st + 2 * A_TILE_BYTES + half * b_tile_bytes;
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
                    if (++s == nstages) {
                        s = 0;
                        ph ^= 1u;
                    }
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
            int it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
                const int sp = (t / p.n_tiles) / p.m_tiles;
                const int kb0 = sp * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                const int buf = it & 1;
                const uint32_t tph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(bar_tempty + 8 * buf, tph ^ 1u);  // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.bn);
                uint32_t acc = 0;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(bar_full + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_base + s * p.stage_bytes;
                    for (int term = 0; term < p.terms; term++) {
                        const uint32_t sa = st + (term == 1 ? A_TILE_BYTES : 0);
                        const uint32_t sb = st + 2 * A_TILE_BYTES + (term == 2 ? b_tile_bytes : 0);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++) {
                            umma_bf16(tmem_d, make_desc(sa + k * a_kstep, p.a_mn),
                                      make_desc(sb + k * b_kstep, p.b_mn), idesc, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
                    if (++s == nstages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                umma_commit(bar_tfull + 8 * buf);  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue =====================
        // TMEM lanes [32q, 32q+32) are only visible to warps with id%4 == q.  With 8 epilogue warps
        // (output-bound shapes) two warps share a lane quarter and take alternate 32-column chunks.
        const int quarter = warp & 3;
        const int nhalf = p.ew >> 2, half = (warp - 2) >> 2;
        // fused batch-norm statistics: warp-private column partials [2][256] behind the staging tiles
        float* colsum = (float*)(smem_dyn + (staging - smem_u32(smem_dyn)) + p.ew * STAGING_PER_WARP) + quarter * 512;
        const bool stats_on = p.col_stats != nullptr;
        if (stats_on)
            for (int c = lane; c < 512; c += 32) colsum[c] = 0.f;
        int it = 0;
        int sbuf = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, it++) {
            const int nt = t % p.n_tiles, mt = (t / p.n_tiles) % p.m_tiles;
            const int m0 = mt * BM, n0 = nt * p.bn;
            const int buf = it & 1;
            const uint32_t tph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(bar_tfull + 8 * buf, tph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + quarter * 32 + lane;
            const bool rows_live = (m0 + quarter * 32) < p.M;
            if (p.tma_store == 2) {
                // fragment-layout epilogue: TMEM -> registers -> sector-exact 8-byte global stores
                // (or vector reductions for split-K); no shared-memory staging, no proxy fences.
                if (rows_live) {
                    const int r_lo = m0 + quarter * 32 + (lane >> 2);
                    for (int c0 = 0; c0 < p.bn; c0 += 64) {
                        if (n0 + c0 >= p.N) break;
                        uint32_t va[32], vb[32];
                        const uint32_t tcol = (uint32_t)(buf * p.bn + c0);
                        tmem_ld_16x256b_x8(tmem_base + ((uint32_t)(quarter * 32) << 16) + tcol, va);
                        tmem_ld_16x256b_x8(tmem_base + ((uint32_t)(quarter * 32 + 16) << 16) + tcol, vb);
                        tmem_wait_ld();
                        frag_store(va, p, r_lo, n0 + c0, lane);
                        frag_store(vb, p, r_lo + 16, n0 + c0, lane);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                continue;
            }
            for (int c0 = half * 32; c0 < p.bn; c0 += 32 * nhalf) {
                if (n0 + c0 >= p.N) break;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.bn + c0), v);
                if (!rows_live && !p.tma_store) continue;
                if (p.tma_store) {
                    // registers -> 128B-swizzled staging tile [128 rows x 32 cols] shared by the four
                    // epilogue warps -> ONE TMA store per chunk (clips at M, N)
                    const uint32_t tile = staging + (uint32_t)sbuf * 16384u;
                    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const uint32_t dst = tile + (uint32_t)(quarter * 32 + lane) * 128u;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t addr = dst + (uint32_t)((p.d_swizzle ? (j ^ (lane & 7)) : j) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * j]),
                                     "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (warp == 2 && lane == 0) {
                        tma_store_2d(&tm_d, tile, n0 + c0, m0, p.atomic_out);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (stats_on) {
                        // lane = column of the chunk: sum this warp's 32 rows from the staging tile (for a fixed
                        // row the 32 lanes read one swizzled 128-byte line: conflict-free); rows past M are zero
                        float s0 = 0.f, s1 = 0.f;
                        const uint32_t col_off = (uint32_t)((lane & 3) << 2);
#pragma unroll 8
                        for (int r = 0; r < 32; r++) {
                            const int rr = quarter * 32 + r;
                            const uint32_t addr = tile + (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7))) << 4) + col_off;
                            float x;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
                            s0 += x;
                            s1 = fmaf(x, x, s1);
                        }
                        colsum[c0 + lane] += s0;
                        colsum[256 + c0 + lane] += s1;
                    }
                    sbuf ^= 1;
                } else if (row < p.M) {
                    float* dst = p.D + (size_t)row * p.ldd + n0 + c0;
                    const int ncols = min(32, p.N - (n0 + c0));
#pragma unroll
                    for (int j = 0; j < 32; j++) {  // static indices keep v[] in registers
                        if (j < ncols) {
                            if (p.atomic_out) atomicAdd(dst + j, __uint_as_float(v[j]));
                            else dst[j] = __uint_as_float(v[j]);
                        }
                    }
                }
            }
            // all tcgen05.ld of this accumulator have completed (wait::ld): hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (stats_on) {
            // four warp-private partial vectors -> one fp64 atomic per column, statistic and CTA
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float* all = colsum - quarter * 512;
            for (int c = (warp - 2) * 32 + lane; c < p.N; c += 128) {
                const float t0 = all[c] + all[512 + c] + all[1024 + c] + all[1536 + c];
                const float t1 = all[256 + c] + all[768 + c] + all[1280 + c] + all[1792 + c];
                atomicAdd(p.col_stats + c, (double)t0);
                atomicAdd(p.col_stats + p.N + c, (double)t1);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// Widest N tile whose padding waste stays small: wide tiles re-read the A operand least.
int choose_bn(int N, bool b_mn, int max_bn) {
    const int cands[4] = {256, 128, 64, 32};
    int min_pad = 1 << 30;
    for (int i = 0; i < 4; i++) {
        if ((b_mn && cands[i] < 64) || cands[i] > max_bn) continue;
        int pad = (N + cands[i] - 1) / cands[i] * cands[i];
        if (pad < min_pad) min_pad = pad;
    }
    for (int i = 0; i < 4; i++) {
        if ((b_mn && cands[i] < 64) || cands[i] > max_bn) continue;
        int pad = (N + cands[i] - 1) / cands[i] * cands[i];
        if ((long long)pad * 100 <= (long long)min_pad * 115) return cands[i];
    }
    return 64;
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" int mvk_col_stats(const float* y, int rows, int cols, int ld, double* stats, mvk_stream_t stream);

static int gemm_impl(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                     const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                     int n_valid, int terms, int split_k, double* col_stats, mvk_stream_t stream) {
    if (!a_hi || !b_hi || !D || M < 1 || N < 1 || K < 1 || (lda % 8) != 0 || (ldb % 8) != 0 ||
        n_valid < 1 || n_valid > N || ldd < n_valid || (terms != 1 && terms != 3))
        return MVK_ERR_INVALID_ARG;
    if (terms == 3 && (!a_lo || !b_lo)) return MVK_ERR_INVALID_ARG;
    if ((((size_t)a_hi | (size_t)b_hi | (size_t)(a_lo ? a_lo : a_hi) | (size_t)(b_lo ? b_lo : b_hi)) & 15) != 0)
        return MVK_ERR_INVALID_ARG;
    GemmParams p;
    p.D = D;
    p.M = M;
    p.N = n_valid;  // columns past n_valid are neither loaded (zero-filled) nor stored
    p.ldd = ldd;
    p.a_mn = a_mn_major ? 1 : 0;
    p.b_mn = b_mn_major ? 1 : 0;
    p.kb_total = (K + BK - 1) / BK;
    // output-bound shapes (short reductions): 8 epilogue warps, tiles at most 128 wide
    p.ew = 4;
    p.bn = choose_bn(n_valid, p.b_mn != 0, 256);
    p.kb_total = (K + BK - 1) / BK;
    bool auto_split = false;
    if (split_k == 0) {
        // auto: few output tiles but a long reduction (the deep, small layers) -> spread K over the SMs;
        // the library zeroes D itself in that case
        const int tiles = ((M + BM - 1) / BM) * ((n_valid + p.bn - 1) / p.bn);
        split_k = 1;
        if (tiles * 2 <= num_sms() && p.kb_total >= 8) {
            split_k = 2 * num_sms() / tiles;  // two work items per persistent CTA
            if (split_k > p.kb_total / 4) split_k = p.kb_total / 4;
            if (split_k < 1) split_k = 1;
        }
        auto_split = split_k > 1;
    }
    if (split_k < 1) split_k = 1;
    if (split_k > p.kb_total) split_k = p.kb_total;
    p.kb_per_split = (p.kb_total + split_k - 1) / split_k;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.terms = terms;
    p.atomic_out = p.splits > 1 ? 1 : 0;
    p.m_tiles = (M + BM - 1) / BM;
    p.n_tiles = (n_valid + p.bn - 1) / p.bn;
    p.num_tiles = p.m_tiles * p.n_tiles * p.splits;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.bn) p.tmem_cols <<= 1;
    p.stage_bytes = 2 * A_TILE_BYTES + 2 * p.bn * 128;
    int staging_bytes = p.ew * STAGING_PER_WARP;
    p.stages = (SMEM_LIMIT - 2048 - staging_bytes) / p.stage_bytes;  // 1 KB alignment slack + static smem
    bool stats_fit = false;
    if (col_stats) {  // the column partials need 8 KB more: only if that does not cost a pipeline stage
        const int with = (SMEM_LIMIT - 2048 - staging_bytes - COLSUM_BYTES) / p.stage_bytes;
        stats_fit = with >= (p.stages < MAX_STAGES ? p.stages : MAX_STAGES);
        if (stats_fit) staging_bytes += COLSUM_BYTES;
    }
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.tma_store = ((ldd % 4) == 0 && (((size_t)D) & 15) == 0) ? 1 : 0;
    // output row pitch not a multiple of 16 bytes: fragment-layout epilogue (8-byte sector-exact stores)
    if (!p.tma_store && (n_valid % 2) == 0 && (ldd % 2) == 0 && (((size_t)D) & 7) == 0) p.tma_store = 2;

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, md;
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
        if ((rc = make_map(&mb_hi, b_hi, K, N, ldb, 64, p.bn))) return rc;
        if ((rc = make_map(&mb_lo, bl, K, N, ldb, 64, p.bn))) return rc;
    } else {
        if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, 64, 64))) return rc;
        if ((rc = make_map(&mb_lo, bl, N, K, ldb, 64, 64))) return rc;
    }
    // statistics ride in the TMA-store epilogue of unsplit, single-column-tile problems; otherwise a
    // separate pass over D below
    p.col_stats = (col_stats && stats_fit && p.tma_store == 1 && p.splits == 1 && p.n_tiles == 1 && n_valid <= 256)
                      ? col_stats : nullptr;
    if (p.tma_store == 1) {
        p.d_swizzle = 1;
        if ((rc = make_map(&md, D, n_valid, M, ldd, 32, 128, 4, p.d_swizzle != 0))) return rc;
    } else {
        md = ma_hi;  // unused
        p.d_swizzle = 0;
    }
    if (auto_split && p.splits > 1)
        MVK_CUDA(cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)n_valid * 4, (size_t)M, (cudaStream_t)stream));
    const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + staging_bytes;
    static thread_local int attr_dev = -1;  // the opt-in shared-memory limit is per device: set it once
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (attr_dev != cur_dev) {
        MVK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024));
        attr_dev = cur_dev;
    }
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    launch_pdl(gemm_tc_kernel, dim3(grid), dim3(64 + 32 * p.ew), smem, (cudaStream_t)stream, 1, ma_hi, ma_lo, mb_hi, mb_lo,
               md, p);
    MVK_LAUNCHED("gemm_tc_kernel");
    if (col_stats && !p.col_stats) return mvk_col_stats(D, M, n_valid, ldd, col_stats, stream);
    return MVK_OK;
}

extern "C" int mvk_gemm_bf16x3(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                               const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                               int n_valid, int terms, int split_k, mvk_stream_t stream) {
    return gemm_impl(a_hi, a_lo, a_mn_major, lda, b_hi, b_lo, b_mn_major, ldb, M, N, K, D, ldd, n_valid, terms, split_k,
                     nullptr, stream);
}

extern "C" int mvk_gemm_bf16x3_stats(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                                     const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                                     int n_valid, int terms, int split_k, double* col_stats, mvk_stream_t stream) {
    if (!col_stats) return MVK_ERR_INVALID_ARG;
    return gemm_impl(a_hi, a_lo, a_mn_major, lda, b_hi, b_lo, b_mn_major, ldb, M, N, K, D, ldd, n_valid, terms, split_k,
                     col_stats, stream);
}
