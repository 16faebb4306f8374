// KPConv forward as ONE kernel: neighbour gather + kernel-point influence (stage A) feeding the tcgen05
// contraction with W through shared memory -- the weighted operand [N, 15*Cin] never goes to HBM
// ("+ 4*15*Cin written if the weighted tensor is staged to HBM; 0 if fused").
// Reference semantics: KPConv-PyTorch/models/blocks.py:277-374 (rigid, 'linear' influence, 'sum' aggregation).
//
// One CTA = a tile of 128 query points = one UMMA M tile.  17 warps:
//   warps 0..15  gather warps, 8 points each.  Per tile they cache their points' neighbourhood geometry
//                {r = s_j - q_i, byte offset of row j of x} in shared memory, then walk the K dimension in blocks of
//                64 columns (= 64/Cin kernel points): influences of the block's kernel points from the cached
//                geometry, non-zero entries ballot-compacted into short lists, 128-bit gathers of the support
//                feature rows (L2), fp32 accumulation in registers, bf16 hi/lo split, and 8-byte stores straight into
//                the 128B-swizzled K-major UMMA tiles A_hi / A_lo of the current stage (2 stages).  Two points are
//                in flight per warp: sub-groups of 8 / 16 lanes each own one (point, kernel point) list.
//   warp 16      one lane: TMA loads of the W hi/lo k-blocks (MN-major boxes, L2-resident) and the
//                tcgen05.mma issue: 3 terms (hi*hi + lo*hi + hi*lo) x 4 K-steps per k-block into the TMEM accumulator
//                (double-buffered across tiles), tcgen05.commit onto the stage / accumulator mbarriers.
//   warps 0..3   also the epilogue of a tile: tcgen05.ld -> swizzled staging tile -> ONE TMA store per 32 columns.
// Training keeps the weighted operand for dW = A^T dOut: the gather warps then also stream the hi/lo values to
// HBM once (save_a), but nothing reads them back in the forward pass.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvk {
namespace {
using namespace tc;

constexpr int FT_M = 128;                 // points per tile
constexpr int FT_GW = 16;                 // gather warps
constexpr int FT_PPW = FT_M / FT_GW;      // points per gather warp
constexpr int FT_THREADS = (FT_GW + 1) * 32;
constexpr int FT_HCAP = 48;               // neighbour slots cached per point
constexpr int FT_KF = 16;                 // kernel-point slots
constexpr int FT_ASTAGES = 2;
constexpr int FT_A_TILE = 16384;          // 128 rows x 128 B (hi or lo)
constexpr int FT_GEOM_BYTES = FT_M * FT_HCAP * 16;
constexpr int FT_LIST_BYTES = FT_GW * 4 * FT_HCAP * 4;  // per warp: 4 lists x 48 packed entries
constexpr int FT_NL = 8;                 // list entries gathered per trip
constexpr int FT_STAGING = 16384;         // epilogue tile [128 rows x 32 fp32]

struct FusedArgs {
    const float* q;
    const float* s;
    const void* inds;
    const float* x;
    const float* kp;
    int nq, ns, h, K, ld;
    float extent;
    __nv_bfloat16* a_hi;  // optional [nq, ld] copy of the weighted operand (training)
    __nv_bfloat16* a_lo;
    int cout, bn, w_stages, w_stage_bytes, tiles;  // bn = MMA N = max(64, cout): MN-major SW128 boxes are 64 columns wide
};

__device__ __forceinline__ float sqrt_approx_f(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float influence_f(float rx, float ry, float rz, float r2, float4 c, float inv_ext) {
    const float d2 = fmaf(rx, c.x, fmaf(ry, c.y, fmaf(rz, c.z, c.w))) + r2;
    return fmaf(-sqrt_approx_f(fmaxf(d2, 0.f)), inv_ext, 1.f);
}
// list entry: neighbour slot in the top 8 bits, weight in (0, 1] as 24-bit fixed point (6e-8 resolution)
__device__ __forceinline__ unsigned int pack_entry(int slot, float w) {
    unsigned int q = (unsigned int)fminf(w * 16777216.f, 16777215.f);
    return ((unsigned int)slot << 24) | q;
}

template <typename IdxT, int CIN>
__global__ void __launch_bounds__(FT_THREADS, 1)
kp_fused_fwd(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
             const __grid_constant__ CUtensorMap tm_d, FusedArgs a) {
    constexpr int KPB = CIN <= 64 ? 64 / CIN : 1;         // kernel points per 64-column k-block
    constexpr int G = CIN == 32 ? 8 : 16;                 // lanes per sub-group (4 channels per lane)
    constexpr int NSG = 32 / G;                           // sub-groups = (point of the pair) x (kernel point of the block)
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bars[2 * FT_ASTAGES + 4 + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float4 s_kc[FT_KF];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* base_g = smem_dyn + (base - smem_u32(smem_dyn));
    // layout: [A stages: 2 x (hi 16K + lo 16K)] [W stages] [staging 16K] [geometry 96K] [lists 12K]
    const uint32_t a_smem = base;
    const uint32_t w_smem = a_smem + FT_ASTAGES * 2 * FT_A_TILE;
    const uint32_t stg_smem = w_smem + (uint32_t)a.w_stages * (uint32_t)a.w_stage_bytes;
    const uint32_t geom_off = (stg_smem - base) + FT_STAGING;
    float4* geom = (float4*)(base_g + geom_off);
    unsigned int* lists = (unsigned int*)(base_g + geom_off + FT_GEOM_BYTES);
    const uint32_t bar_afull = smem_u32(&bars[0]);                    // [2] gather warps -> MMA
    const uint32_t bar_aempty = smem_u32(&bars[FT_ASTAGES]);          // [2] MMA -> gather warps
    const uint32_t bar_wfull = smem_u32(&bars[2 * FT_ASTAGES]);       // [2] TMA -> MMA
    const uint32_t bar_wempty = smem_u32(&bars[2 * FT_ASTAGES + 2]);  // [2] MMA -> TMA
    const uint32_t bar_tfull = smem_u32(&bars[2 * FT_ASTAGES + 4]);   // [2] accumulator ready
    const uint32_t bar_tempty = smem_u32(&bars[2 * FT_ASTAGES + 6]);  // [2] accumulator drained
    const int tmem_cols = 2 * a.bn;                  // two accumulator buffers: 128 or 256 columns (powers of two)
    const int nkb = (a.K * CIN + 63) / 64;           // k-blocks of 64 columns over the K*Cin reduction

    // kernel points are a constant of the layer (requires_grad = False, never written on the device): reading them
    // before pdl_enter() is safe.  c[k] = (-2 kx, -2 ky, -2 kz, |kp|^2); unused slots can never have influence
    if (threadIdx.x < FT_KF) {
        const int k = threadIdx.x;
        float4 c = make_float4(0.f, 0.f, 0.f, 1e30f);
        if (k < a.K) {
            const float x = a.kp[3 * k], y = a.kp[3 * k + 1], z = a.kp[3 * k + 2];
            c = make_float4(-2.f * x, -2.f * y, -2.f * z, x * x + y * y + z * z);
        }
        s_kc[k] = c;
    }
    if (warp == FT_GW && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_d) : "memory");
        for (int s = 0; s < FT_ASTAGES; s++) {
            mbar_init(bar_afull + 8 * s, FT_GW);
            mbar_init(bar_aempty + 8 * s, 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == FT_GW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_smem)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    pdl_enter();

    if (warp == FT_GW) {
        // ===================== W loads + MMA issue (one lane) =====================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) |
                                   ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(FT_M >> 4) << 24);
            const uint32_t w_half = (uint32_t)a.bn * 128u;  // one of hi / lo: 64 k-rows x bn columns x 2 B
            uint32_t kbi = 0;                                  // k-blocks issued so far (A stage ring)
            uint32_t wi = 0;                                   // W stage ring
            int it = 0;
            for (int t = blockIdx.x; t < a.tiles; t += gridDim.x, it++) {
                const int buf = it & 1;
                const uint32_t tph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(bar_tempty + 8 * buf, tph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * a.bn);
                for (int kb = 0; kb < nkb; kb++, kbi++, wi++) {
                    const uint32_t ws = wi % (uint32_t)a.w_stages, wph = (wi / (uint32_t)a.w_stages) & 1u;
                    mbar_wait(bar_wempty + 8 * ws, wph ^ 1u);
                    const uint32_t wfull = bar_wfull + 8 * ws;
                    const uint32_t sw = w_smem + ws * (uint32_t)a.w_stage_bytes;
                    mbar_expect_tx(wfull, 2u * w_half);
                    for (int c = 0; c < a.bn / 64; c++) {
                        tma_load_2d(sw + 8192 * c, &tm_w_hi, wfull, 64 * c, kb * 64);
                        tma_load_2d(sw + w_half + 8192 * c, &tm_w_lo, wfull, 64 * c, kb * 64);
                    }
                    const uint32_t s = kbi & 1u, ph = (kbi >> 1) & 1u;
                    mbar_wait(bar_afull + 8 * s, ph);
                    mbar_wait(wfull, wph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = a_smem + s * 2 * FT_A_TILE;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint32_t pa = sa + (term == 1 ? FT_A_TILE : 0);
                        const uint32_t pb = sw + (term == 2 ? w_half : 0);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma_bf16(tmem_d, make_desc(pa + k * 32, 0), make_desc(pb + k * 2048, 1), idesc,
                                      (kb > 0 || term > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_aempty + 8 * s);
                    umma_commit(bar_wempty + 8 * ws);
                }
                umma_commit(bar_tfull + 8 * buf);
            }
        }
    } else {
        // ===================== gather warps =====================
        const float inv_ext = 1.f / a.extent;
        const int H = a.h, ns = a.ns;
        const unsigned int row_bytes = CIN * 4u;
        const int sg = lane / G, lg = lane % G;          // sub-group, lane in sub-group
        const int pt_of_sg = sg / KPB, kp_of_sg = sg % KPB;  // which point of the pair / which kernel point of the block
        float4* my_geom = geom + (size_t)warp * FT_PPW * FT_HCAP;
        unsigned int* my_lists = lists + (size_t)warp * 4 * FT_HCAP;
        const unsigned int lt_mask = (1u << lane) - 1u;
        uint32_t kbi = 0;
        int it = 0;
        for (int t = blockIdx.x; t < a.tiles; t += gridDim.x, it++) {
            const int m0 = t * FT_M;
            // ---- cache the neighbourhood geometry of this warp's 8 points
            for (int p = 0; p < FT_PPW; p++) {
                const int i = m0 + warp * FT_PPW + p;
                float qx = 0.f, qy = 0.f, qz = 0.f;
                if (i < a.nq) {
                    qx = a.q[3 * (size_t)i];
                    qy = a.q[3 * (size_t)i + 1];
                    qz = a.q[3 * (size_t)i + 2];
                }
                for (int h = lane; h < FT_HCAP; h += 32) {
                    int j = ns;
                    if (i < a.nq && h < H) j = (int)((const IdxT*)a.inds)[(size_t)i * H + h];
                    float4 g = make_float4(1e6f, 1e6f, 1e6f, 0.f);  // shadow: no kernel point within reach
                    if ((unsigned int)j < (unsigned int)ns) {
                        g.x = __ldg(a.s + 3 * (size_t)j) - qx;
                        g.y = __ldg(a.s + 3 * (size_t)j + 1) - qy;
                        g.z = __ldg(a.s + 3 * (size_t)j + 2) - qz;
                        g.w = __uint_as_float((unsigned int)j * row_bytes);
                    }
                    my_geom[p * FT_HCAP + h] = g;
                }
            }
            __syncwarp();
            // ---- K loop
            for (int kb = 0; kb < nkb; kb++, kbi++) {
                const uint32_t s = kbi & 1u, ph = (kbi >> 1) & 1u;
                mbar_wait(bar_aempty + 8 * s, ph ^ 1u);
                unsigned char* a_hi_t = base_g + (a_smem - base) + s * 2 * FT_A_TILE;
                unsigned char* a_lo_t = a_hi_t + FT_A_TILE;
                const int k_first = CIN <= 64 ? kb * KPB : kb / (CIN / 64);  // first kernel point of the block
                const int ch_off = CIN <= 64 ? 0 : (kb % (CIN / 64)) * 64;  // channel offset inside the kernel point
                for (int pp = 0; pp < FT_PPW; pp += 2) {
                    // phase 1: influences of the block's kernel points for points pp (A) and pp + 1 (B)
                    int cnt[4] = {0, 0, 0, 0};  // list sizes: index = point * KPB + kernel point
#pragma unroll
                    for (int pass = 0; pass < 3; pass++) {
                        // pass 0: A slots 0..31, pass 1: B slots 0..31, pass 2: A slots 32..47 | B slots 32..47
                        const int pt = pass == 2 ? (lane >> 4) : pass;
                        const int slot = pass == 2 ? 32 + (lane & 15) : lane;
                        const float4 g = my_geom[(pp + pt) * FT_HCAP + slot];
                        const float r2 = fmaf(g.x, g.x, fmaf(g.y, g.y, g.z * g.z));
#pragma unroll
                        for (int kk = 0; kk < KPB; kk++) {
                            const float w = influence_f(g.x, g.y, g.z, r2, s_kc[k_first + kk], inv_ext);
                            const bool nz = w > 0.f;
                            const unsigned int m = __ballot_sync(0xffffffffu, nz);
                            if (pass < 2) {
                                const int li = pass * KPB + kk;
                                if (nz) my_lists[li * FT_HCAP + cnt[li] + __popc(m & lt_mask)] = pack_entry(slot, w);
                                cnt[li] += __popc(m);
                            } else {
                                const unsigned int mA = m & 0xffffu, mB = m >> 16;
                                const int liA = kk, liB = KPB + kk;
                                if (nz) {
                                    if (lane < 16) my_lists[liA * FT_HCAP + cnt[liA] + __popc(mA & lt_mask)] = pack_entry(slot, w);
                                    else my_lists[liB * FT_HCAP + cnt[liB] + __popc(mB & (lt_mask >> 16))] = pack_entry(slot, w);
                                }
                                cnt[liA] += __popc(mA);
                                cnt[liB] += __popc(mB);
                            }
                        }
                    }
                    __syncwarp();
                    // phase 2: sub-group sg walks the list of (point pt_of_sg, kernel point kp_of_sg)
                    int n = cnt[0];
#pragma unroll
                    for (int u = 1; u < NSG; u++) n = (sg == u) ? cnt[u] : n;
                    const unsigned int* lk = my_lists + sg * FT_HCAP;
                    const float4* gk = my_geom + (pp + pt_of_sg) * FT_HCAP;
                    const char* xb = (const char*)(a.x + ch_off + lg * 4);
                    // up to FT_NL gathers in flight per lane (a list holds ~4.5 entries on average): all loads are issued
                    // before the first FMA, so one L2 round trip covers the whole list
                    float w8[FT_NL];
                    float4 x8[FT_NL];
#pragma unroll
                    for (int u = 0; u < FT_NL; u++) {
                        w8[u] = 0.f;
                        x8[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (u < n) {
                            const unsigned int e = lk[u];
                            w8[u] = (float)(e & 0xffffffu) * (1.f / 16777216.f);
                            x8[u] = __ldg((const float4*)(xb + __float_as_uint(gk[e >> 24].w)));
                        }
                    }
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < FT_NL; u++) {
                        acc.x = fmaf(w8[u], x8[u].x, acc.x); acc.y = fmaf(w8[u], x8[u].y, acc.y);
                        acc.z = fmaf(w8[u], x8[u].z, acc.z); acc.w = fmaf(w8[u], x8[u].w, acc.w);
                    }
#pragma unroll 1
                    for (int u = FT_NL; u < n; u++) {  // long lists (rare)
                        const unsigned int e = lk[u];
                        const float w = (float)(e & 0xffffffu) * (1.f / 16777216.f);
                        const float4 xv = __ldg((const float4*)(xb + __float_as_uint(gk[e >> 24].w)));
                        acc.x = fmaf(w, xv.x, acc.x); acc.y = fmaf(w, xv.y, acc.y);
                        acc.z = fmaf(w, xv.z, acc.z); acc.w = fmaf(w, xv.w, acc.w);
                    }
                    // bf16 hi / lo, 8 bytes each, into the swizzled K-major tile: row = point, column = position in the k-block
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
                    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                    const __nv_bfloat162 l0 = __floats2bfloat162_rn(acc.x - f0.x, acc.y - f0.y);
                    const __nv_bfloat162 l1 = __floats2bfloat162_rn(acc.z - f1.x, acc.w - f1.y);
                    uint2 ph2, pl2;
                    ph2.x = *(const unsigned int*)&h0; ph2.y = *(const unsigned int*)&h1;
                    pl2.x = *(const unsigned int*)&l0; pl2.y = *(const unsigned int*)&l1;
                    const int r = warp * FT_PPW + pp + pt_of_sg;            // row of the tile
                    const int col = (CIN <= 64 ? kp_of_sg * CIN : 0) + lg * 4;  // column inside the 64-wide k-block
                    const unsigned int off = (unsigned int)r * 128u + ((((unsigned int)col >> 3) ^ ((unsigned int)r & 7u)) << 4) +
                                             (((unsigned int)col & 7u) << 1);
                    *(uint2*)(a_hi_t + off) = ph2;
                    *(uint2*)(a_lo_t + off) = pl2;
                    if (a.a_hi) {
                        const int i = m0 + r;
                        const int kcol = (k_first + kp_of_sg) * CIN + ch_off + lg * 4;
                        if (i < a.nq && k_first + kp_of_sg < a.K) {
                            *(uint2*)(a.a_hi + (size_t)i * a.ld + kcol) = ph2;
                            *(uint2*)(a.a_lo + (size_t)i * a.ld + kcol) = pl2;
                        }
                    }
                    __syncwarp();
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * s);
            }
            // ---- epilogue of the tile (warps 0..3: TMEM lane quarter = warp)
            if (warp < 4) {
                const int buf = it & 1;
                const uint32_t tph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(bar_tfull + 8 * buf, tph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c0 = 0; c0 < a.cout; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * a.bn + c0), v);
                    if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const uint32_t dst = stg_smem + (uint32_t)(warp * 32 + lane) * 128u;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t addr = dst + (uint32_t)((j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * j]),
                                     "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (warp == 0 && lane == 0) {
                        tma_store_2d(&tm_d, stg_smem, c0, m0, 0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            }
        }
        if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == FT_GW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" int mvk_kpconv_fused_supported(int cin, int cout, int num_kp, int h, int influence, int aggregation) {
    return (influence == 1 && aggregation == 0 && num_kp >= 1 && num_kp <= 15 && (cin == 32 || cin == 64 || cin == 128) &&
            (cout == 32 || cout == 64 || cout == 128) && h >= 1 && h <= FT_HCAP) ? 1 : 0;
}

extern "C" int mvk_kpconv_fused(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                                int idx_is_i64, int h, const float* x, int cin, const float* kernel_points, int num_kp,
                                float kp_extent, int cout, const void* w_hi, const void* w_lo, int ldw, float* out,
                                void* a_hi, void* a_lo, int ld, mvk_stream_t stream) {
    if (!q_pts || !s_pts || !neighb_inds || !x || !kernel_points || !w_hi || !w_lo || !out || nq < 0 || ns < 0 ||
        !(kp_extent > 0.f) || (ldw % 8) != 0 || ldw < cout || ((a_hi == nullptr) != (a_lo == nullptr)) ||
        (a_hi && (ld < num_kp * cin || (ld % 4) != 0)))
        return MVK_ERR_INVALID_ARG;
    if (!mvk_kpconv_fused_supported(cin, cout, num_kp, h, 1, 0)) return MVK_ERR_UNSUPPORTED;
    if ((size_t)(ns + 1) * cin * 4 >= 0xffffffffull) return MVK_ERR_RANGE;
    if (((((size_t)w_hi) | ((size_t)w_lo) | ((size_t)out)) & 15) != 0) return MVK_ERR_INVALID_ARG;
    if (nq == 0) return MVK_OK;
    FusedArgs a;
    a.q = q_pts; a.s = s_pts; a.inds = neighb_inds; a.x = x; a.kp = kernel_points;
    a.nq = nq; a.ns = ns; a.h = h; a.K = num_kp; a.ld = ld; a.extent = kp_extent;
    a.a_hi = (__nv_bfloat16*)a_hi; a.a_lo = (__nv_bfloat16*)a_lo;
    a.cout = cout;
    a.bn = cout < 64 ? 64 : cout;
    a.w_stages = a.bn <= 64 ? 2 : 1;
    a.w_stage_bytes = 2 * a.bn * 128;
    a.tiles = (nq + FT_M - 1) / FT_M;
    CUtensorMap mw_hi, mw_lo, md;
    int rc;
    const int kd = num_kp * cin;
    // W [kd rows (K), cout cols] row-major bf16, MN-major operand: boxes of 64 columns x 64 k-rows; rows past kd zero-fill
    // (a 64-wide box over a 32-column matrix: the columns past cout arrive as zeros)
    if ((rc = make_map(&mw_hi, w_hi, (uint64_t)cout, (uint64_t)kd, (uint64_t)ldw, 64, 64))) return rc;
    if ((rc = make_map(&mw_lo, w_lo, (uint64_t)cout, (uint64_t)kd, (uint64_t)ldw, 64, 64))) return rc;
    if ((rc = make_map(&md, out, (uint64_t)cout, (uint64_t)nq, (uint64_t)cout, 32, 128, 4, true))) return rc;
    const size_t smem = 1024 + (size_t)FT_ASTAGES * 2 * FT_A_TILE + (size_t)a.w_stages * a.w_stage_bytes + FT_STAGING +
                        FT_GEOM_BYTES + FT_LIST_BYTES;
    int grid = a.tiles < num_sms() ? a.tiles : num_sms();
#define LAUNCH_FUSED(IDX, C)                                                                                         \
    do {                                                                                                             \
        auto kern = kp_fused_fwd<IDX, C>;                                                                            \
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        launch_pdl(kern, dim3(grid), dim3(FT_THREADS), smem, (cudaStream_t)stream, 1, mw_hi, mw_lo, md, a);          \
    } while (0)
    if (idx_is_i64) {
        if (cin == 32) LAUNCH_FUSED(long long, 32);
        else if (cin == 64) LAUNCH_FUSED(long long, 64);
        else LAUNCH_FUSED(long long, 128);
    } else {
        if (cin == 32) LAUNCH_FUSED(int, 32);
        else if (cin == 64) LAUNCH_FUSED(int, 64);
        else LAUNCH_FUSED(int, 128);
    }
#undef LAUNCH_FUSED
    MVK_LAUNCHED("kp_fused_fwd");
    return MVK_OK;
}
