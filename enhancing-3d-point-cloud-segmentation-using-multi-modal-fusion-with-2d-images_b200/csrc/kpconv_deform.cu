// Deformable KPConv, stage A (forward and backward).
// Reference: KPConv-PyTorch/models/blocks.py:243-270 (offsets / modulations), :286-325 (per-point
// kernel points, min_d2, in-range compaction), :329-367 (influences, modulated weighting).
//
//   weighted[i, k*cin + c] = mod_ik * sum_h w_ihk * x[j_ih, c]
//   w_ihk    = influence(|| (s_j - q_i) - kp_ik ||)          kp_ik = kernel_points[k] + offsets[i, k]
//   min_d2[i, k] = min_h || (s_j - q_i) - kp_ik ||^2          (over ALL neighbour slots, shadows at 1e6)
//
// The reference drops neighbours that are out of range of every kernel point before computing the
// influences (an optimisation: with the 'linear' influence their weights are zero anyway); the same
// mask is applied here so 'constant' / 'gaussian' influences agree too.
//
// One warp per query point; lanes = neighbours while the influences are computed (non-zero entries
// are ballot-compacted into per-kernel-point lists in shared memory), lanes = channels afterwards.
// Deformable layers sit on the deep, small levels of the networks, so these kernels favour
// generality (any cin, K <= 16, all influence modes) over the sub-group tricks of the rigid fast path.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr int KD = 16;  // kernel-point slots

struct DArgs {
    const float* q;
    const float* s;
    const void* inds;
    const float* x;
    const float* kp;   // [nq, K, 3] deformed kernel points
    const float* mod;  // [nq, K] or NULL
    int nq, ns, h, cin, K, ld;
    float extent;
    int influence, aggregation;
};

template <typename IdxT>
__device__ __forceinline__ int ld_index(const void* inds, size_t off) {
    return (int)((const IdxT*)inds)[off];
}

// influence and d(influence)/d(d2) * 2 ... returned as the factor f such that  dw/dkp = f * (r - kp)
__device__ __forceinline__ float influence_and_grad(float d2, float extent, int mode, float* f) {
    if (mode == 1) {  // linear: w = max(0, 1 - d / extent); dw/dkp = (r - kp) / (d * extent)
        const float d = __fsqrt_rn(d2);
        const float w = fmaxf(0.f, 1.f - __fdiv_rn(d, extent));
        *f = (w > 0.f && d > 0.f) ? 1.f / (d * extent) : 0.f;
        return w;
    }
    if (mode == 0) {
        *f = 0.f;
        return 1.f;
    }
    const float sigma = extent * 0.3f;
    const float den = 2.f * sigma * sigma + 1e-9f;
    const float w = __expf(-d2 / den);
    *f = 2.f * w / den;  // w = exp(-d2 / den), d(d2)/dkp = -2 (r - kp)
    return w;
}

// Phase 1 for one point: lists of non-zero {support row, weight[, gradient factor * (r - kp)]} per kernel
// point, running minimum of d2 per kernel point.  Returns nothing; lists / counts live in shared memory.
template <typename IdxT, bool GRAD>
__device__ __forceinline__ void deform_lists(const DArgs& a, int i, int lane, const float* kc /*[KD][4]*/, int* jl,
                                             float* wl, float4* gl, int* cnt, int hcap, float (&mind)[KD],
                                             int (&minh)[KD]) {
    const float qx = a.q[3 * i], qy = a.q[3 * i + 1], qz = a.q[3 * i + 2];
    const float ext2 = a.extent * a.extent;
    int count[KD];
#pragma unroll
    for (int k = 0; k < KD; k++) {
        count[k] = 0;
        mind[k] = 3.4e38f;
        minh[k] = 0x7fffffff;
    }
    const unsigned int lt_mask = (1u << lane) - 1u;
    for (int h0 = 0; h0 < a.h; h0 += 32) {
        const int h = h0 + lane;
        const bool slot = h < a.h;
        int j = a.ns;
        if (slot) j = ld_index<IdxT>(a.inds, (size_t)i * a.h + h);
        const bool real = slot && j >= 0 && j < a.ns;
        // shadow neighbours sit at (1e6, 1e6, 1e6) like the reference's padded support row
        float rx = 1e6f - qx, ry = 1e6f - qy, rz = 1e6f - qz;
        if (real) {
            rx = a.s[3 * (size_t)j] - qx;
            ry = a.s[3 * (size_t)j + 1] - qy;
            rz = a.s[3 * (size_t)j + 2] - qz;
        }
        float d2k[KD];
        bool in_range = false;
        int kmin = 0;
        float dmin = 3.4e38f;
#pragma unroll
        for (int k = 0; k < KD; k++) {
            const float dx = rx - kc[4 * k], dy = ry - kc[4 * k + 1], dz = rz - kc[4 * k + 2];
            d2k[k] = dx * dx + dy * dy + dz * dz;
            if (k < a.K) {
                in_range = in_range || (d2k[k] < ext2);
                if (d2k[k] < dmin) {
                    dmin = d2k[k];
                    kmin = k;
                }
                if (slot && (d2k[k] < mind[k])) {  // first minimum along the row, like torch.min
                    mind[k] = d2k[k];
                    minh[k] = h;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KD; k++) {
            float f = 0.f;
            float w = influence_and_grad(d2k[k], a.extent, a.influence, &f);
            if (a.aggregation == 1 && k != kmin) w = 0.f;
            const bool nz = real && in_range && (k < a.K) && (w > 0.f);
            const unsigned int m = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                const int pos = count[k] + __popc(m & lt_mask);
                jl[k * hcap + pos] = j;
                wl[k * hcap + pos] = w;
                if (GRAD) gl[k * hcap + pos] = make_float4(f * (rx - kc[4 * k]), f * (ry - kc[4 * k + 1]),
                                                           f * (rz - kc[4 * k + 2]), 0.f);
            }
            count[k] += __popc(m);
        }
    }
#pragma unroll
    for (int k = 0; k < KD; k++)
        if (lane == 0) cnt[k] = count[k];
    // warp-wide (min d2, first slot) per kernel point
#pragma unroll
    for (int k = 0; k < KD; k++) {
        unsigned long long key = ((unsigned long long)__float_as_uint(mind[k]) << 32) | (unsigned int)minh[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        mind[k] = __uint_as_float((unsigned int)(key >> 32));
        minh[k] = (int)(key & 0xffffffffu);
    }
    __syncwarp();
}

__device__ __forceinline__ void put(float* of, __nv_bfloat16* ohi, __nv_bfloat16* olo, size_t off, float v) {
    if (of) of[off] = v;
    if (ohi) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        ohi[off] = hi;
        olo[off] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

__device__ __forceinline__ void load_kc(const DArgs& a, int i, int lane, float* kc) {
    __syncwarp();
    if (lane < KD) {
        float x = 1e30f, y = 1e30f, z = 1e30f;  // unused slots: infinitely far
        if (lane < a.K) {
            const float* p = a.kp + ((size_t)i * a.K + lane) * 3;
            x = p[0]; y = p[1]; z = p[2];
        }
        kc[4 * lane] = x; kc[4 * lane + 1] = y; kc[4 * lane + 2] = z; kc[4 * lane + 3] = 0.f;
    }
    __syncwarp();
}

template <typename IdxT>
__global__ void __launch_bounds__(128)
kpd_fwd(DArgs a, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
        float* __restrict__ min_d2, int* __restrict__ argmin, int hcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)KD * 16 + (size_t)KD * hcap * 8 + KD * 4;
    unsigned char* wbase = smem_raw + wib * per_warp;
    float* kc = (float*)wbase;
    int* jl = (int*)(wbase + KD * 16);
    float* wl = (float*)(wbase + KD * 16 + (size_t)KD * hcap * 4);
    int* cnt = (int*)(wbase + KD * 16 + (size_t)KD * hcap * 8);
    const int kd = a.K * a.cin;
    for (int i = blockIdx.x * wpb + wib; i < a.nq; i += gridDim.x * wpb) {
        load_kc(a, i, lane, kc);
        float mind[KD];
        int minh[KD];
        deform_lists<IdxT, false>(a, i, lane, kc, jl, wl, nullptr, cnt, hcap, mind, minh);
#pragma unroll
        for (int k = 0; k < KD; k++)
            if (lane == k && k < a.K) {
                if (min_d2) min_d2[(size_t)i * a.K + k] = mind[k];
                if (argmin) argmin[(size_t)i * a.K + k] = minh[k];
            }
        const size_t row = (size_t)i * a.ld;
        for (int cb = 0; cb < a.cin; cb += 32) {
            const int c = cb + lane;
            if (c < a.cin) {
                for (int k = 0; k < a.K; k++) {
                    const int n = cnt[k];
                    const int* jk = jl + k * hcap;
                    const float* wk = wl + k * hcap;
                    float acc = 0.f;
                    for (int t = 0; t < n; t++) acc = fmaf(wk[t], __ldg(a.x + (size_t)jk[t] * a.cin + c), acc);
                    if (a.mod) acc *= a.mod[(size_t)i * a.K + k];
                    put(out_f32, out_hi, out_lo, row + (size_t)k * a.cin + c, acc);
                }
            }
        }
        for (int c = kd + lane; c < a.ld; c += 32) put(out_f32, out_hi, out_lo, row + c, 0.f);
        __syncwarp();
    }
}

// Backward: grad_x (atomics), grad_kp [nq, K, 3], grad_mod [nq, K].
template <typename IdxT>
__global__ void __launch_bounds__(128)
kpd_bwd(DArgs a, const float* __restrict__ gw, const float* __restrict__ g_min, const int* __restrict__ argmin,
        float* __restrict__ gx, float* __restrict__ gkp, float* __restrict__ gmod, int hcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const size_t per_warp = (size_t)KD * 16 + (size_t)KD * hcap * 24 + KD * 4;
    unsigned char* wbase = smem_raw + wib * per_warp;
    float* kc = (float*)wbase;
    float4* gl = (float4*)(wbase + KD * 16);
    int* jl = (int*)(wbase + KD * 16 + (size_t)KD * hcap * 16);
    float* wl = (float*)(wbase + KD * 16 + (size_t)KD * hcap * 20);
    int* cnt = (int*)(wbase + KD * 16 + (size_t)KD * hcap * 24);
    for (int i = blockIdx.x * wpb + wib; i < a.nq; i += gridDim.x * wpb) {
        load_kc(a, i, lane, kc);
        float mind[KD];
        int minh[KD];
        deform_lists<IdxT, true>(a, i, lane, kc, jl, wl, gl, cnt, hcap, mind, minh);
        const size_t row = (size_t)i * a.ld;
        for (int k = 0; k < a.K; k++) {
            const int n = cnt[k];
            const float mk = a.mod ? a.mod[(size_t)i * a.K + k] : 1.f;
            float dkx = 0.f, dky = 0.f, dkz = 0.f, dm = 0.f;
            for (int t = 0; t < n; t++) {
                const int j = jl[k * hcap + t];
                const float w = wl[k * hcap + t];
                float s = 0.f;
                for (int c = lane; c < a.cin; c += 32) {
                    const float g = gw[row + (size_t)k * a.cin + c];
                    s = fmaf(__ldg(a.x + (size_t)j * a.cin + c), g, s);
                    if (gx) atomicAdd(gx + (size_t)j * a.cin + c, w * mk * g);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float4 gr = gl[k * hcap + t];
                dkx = fmaf(s, gr.x, dkx);
                dky = fmaf(s, gr.y, dky);
                dkz = fmaf(s, gr.z, dkz);
                dm = fmaf(s, w, dm);
            }
            if (lane == 0) {
                dkx *= mk; dky *= mk; dkz *= mk;
                if (g_min && argmin) {
                    // d(min_d2)/d(kp) = -2 (r_h* - kp): r of the arg-min slot (shadow slots sit at 1e6)
                    const int hs = argmin[(size_t)i * a.K + k];
                    if (hs >= 0 && hs < a.h) {
                        const int j = ld_index<IdxT>(a.inds, (size_t)i * a.h + hs);
                        float rx = 1e6f - a.q[3 * i], ry = 1e6f - a.q[3 * i + 1], rz = 1e6f - a.q[3 * i + 2];
                        if (j >= 0 && j < a.ns) {
                            rx = a.s[3 * (size_t)j] - a.q[3 * i];
                            ry = a.s[3 * (size_t)j + 1] - a.q[3 * i + 1];
                            rz = a.s[3 * (size_t)j + 2] - a.q[3 * i + 2];
                        }
                        const float gm = g_min[(size_t)i * a.K + k];
                        dkx += gm * -2.f * (rx - kc[4 * k]);
                        dky += gm * -2.f * (ry - kc[4 * k + 1]);
                        dkz += gm * -2.f * (rz - kc[4 * k + 2]);
                    }
                }
                if (gkp) {
                    float* o = gkp + ((size_t)i * a.K + k) * 3;
                    o[0] = dkx; o[1] = dky; o[2] = dkz;
                }
                if (gmod) gmod[(size_t)i * a.K + k] = dm;
            }
        }
        __syncwarp();
    }
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

int mvk_kpconv_deform_weighted(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                               int idx_is_i64, int h, const float* x, int cin, const float* deformed_kp,
                               const float* modulations, int num_kp, float kp_extent, int influence, int aggregation,
                               int ld, float* out_f32, void* out_hi, void* out_lo, float* min_d2, int* argmin,
                               mvk_stream_t stream) {
    if (nq < 0 || ns < 0 || h < 1 || cin < 1 || num_kp < 1 || ld < num_kp * cin || !deformed_kp ||
        (!out_f32 && !(out_hi && out_lo)) || !(kp_extent > 0.f))
        return MVK_ERR_INVALID_ARG;
    if (num_kp > KD || influence < 0 || influence > 2 || aggregation < 0 || aggregation > 1) return MVK_ERR_UNSUPPORTED;
    if (nq == 0) return MVK_OK;
    DArgs a{q_pts, s_pts, neighb_inds, x, deformed_kp, modulations, nq, ns, h, cin, num_kp, ld, kp_extent,
            influence, aggregation};
    const int hcap = (h + 3) & ~3;
    const size_t per_warp = (size_t)KD * 16 + (size_t)KD * hcap * 8 + KD * 4;
    // wide neighbourhoods (deform_radius = 6 makes rows 5-6x wider than the rigid ones): fewer warps per CTA
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
    if (per_warp * wpb > 200 * 1024) return MVK_ERR_RANGE;
    const size_t smem = per_warp * wpb;
    int blocks = (nq + wpb - 1) / wpb;
    const int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_is_i64) {
        auto kern = kpd_fwd<long long>;
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, wpb * 32, smem, st>>>(a, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, min_d2, argmin, hcap);
    } else {
        auto kern = kpd_fwd<int>;
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, wpb * 32, smem, st>>>(a, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, min_d2, argmin, hcap);
    }
    MVK_LAUNCHED("kpd_fwd");
    return MVK_OK;
}

int mvk_kpconv_deform_weighted_bwd(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                                   int idx_is_i64, int h, const float* x, int cin, const float* deformed_kp,
                                   const float* modulations, int num_kp, float kp_extent, int influence,
                                   int aggregation, const float* grad_weighted, int ld, const float* grad_min_d2,
                                   const int* argmin, float* grad_x, float* grad_kp, float* grad_mod,
                                   mvk_stream_t stream) {
    if (nq < 0 || ns < 0 || h < 1 || cin < 1 || num_kp < 1 || ld < num_kp * cin || !deformed_kp || !grad_weighted ||
        !x || !(kp_extent > 0.f) || (grad_min_d2 && !argmin))
        return MVK_ERR_INVALID_ARG;
    if (num_kp > KD || influence < 0 || influence > 2 || aggregation < 0 || aggregation > 1) return MVK_ERR_UNSUPPORTED;
    if (nq == 0) return MVK_OK;
    DArgs a{q_pts, s_pts, neighb_inds, x, deformed_kp, modulations, nq, ns, h, cin, num_kp, ld, kp_extent,
            influence, aggregation};
    const int hcap = (h + 3) & ~3;
    const size_t per_warp = (size_t)KD * 16 + (size_t)KD * hcap * 24 + KD * 4;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;  // rows up to ~530 wide with one warp per CTA
    if (per_warp * wpb > 200 * 1024) return MVK_ERR_RANGE;
    const size_t smem = per_warp * wpb;
    int blocks = (nq + wpb - 1) / wpb;
    const int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_is_i64) {
        auto kern = kpd_bwd<long long>;
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, wpb * 32, smem, st>>>(a, grad_weighted, grad_min_d2, argmin, grad_x, grad_kp, grad_mod, hcap);
    } else {
        auto kern = kpd_bwd<int>;
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, wpb * 32, smem, st>>>(a, grad_weighted, grad_min_d2, argmin, grad_x, grad_kp, grad_mod, hcap);
    }
    MVK_LAUNCHED("kpd_bwd");
    return MVK_OK;
}

}  // extern "C"
