// KPConv stage A: neighbour-feature gather + kernel-point influence (forward and backward),
// the bf16 hi/lo splitter feeding the tensor-core contraction, and the gather pools.
// Reference semantics: KPConv-PyTorch/models/blocks.py:277-363 (stage A), :79-110 (pools).
//
// One warp owns one query point.
//   phase 1  lanes = neighbours: relative position, the K kernel-point distances and influences.
//            The influence matrix w[h][k] is SPARSE (a neighbour lies within KP_extent of ~2 of the
//            15 kernel points), so non-zero entries are ballot-compacted into per-kernel-point lists
//            {support row, weight} in shared memory (deterministic order: by neighbour slot).
//   phase 2  lanes = channels: for every kernel point walk its list, gather the support feature
//            rows with coalesced vector loads (x is L2-resident at these sizes) and accumulate in
//            registers; rows of the weighted matrix [nq, K*cin] are written once, coalesced.
// Shadow neighbours (index >= ns) contribute exactly zero in the reference (zero feature row) and
// are skipped.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr int KP_MAX = 32;

struct KpArgs {
    const float* q;
    const float* s;
    const void* inds;
    const float* x;
    const float* kp;
    int nq, ns, h, cin, K, ld;
    float extent;
    int influence, aggregation;
    int width;  // columns of a row this call owns (K*cin values + zero padding up to width); == ld unless the row is shared
};

template <typename IdxT>
__device__ __forceinline__ int load_idx(const void* inds, size_t off) {
    return (int)((const IdxT*)inds)[off];
}

// Influence of kernel point at squared distance d2 (blocks.py:329-346).
__device__ __forceinline__ float influence_w(float d2, float extent, int mode) {
    if (mode == 1) {
        // zero beyond the extent: skip the IEEE sqrt/div there (guard band keeps boundary cases exact)
        if (d2 > extent * extent * 1.0001f) return 0.f;
        return fmaxf(0.f, 1.f - __fdiv_rn(__fsqrt_rn(d2), extent));
    }
    if (mode == 0) return 1.f;
    float sigma = extent * 0.3f;
    return __expf(-d2 / (2.f * sigma * sigma + 1e-9f));
}

// Phase 1 shared by forward and backward: fills per-kernel-point lists.
//   jl/wl : [K][hcap]   cnt : [K]
template <typename IdxT>
__device__ __forceinline__ void build_lists(const KpArgs& a, int i, int lane, const float* s_kp,
                                            int* jl, float* wl, int* cnt, int hcap) {
    for (int k = lane; k < a.K; k += 32) cnt[k] = 0;
    __syncwarp();
    const float qx = a.q[3 * i], qy = a.q[3 * i + 1], qz = a.q[3 * i + 2];
    for (int h0 = 0; h0 < a.h; h0 += 32) {
        int h = h0 + lane;
        int j = a.ns;
        if (h < a.h) j = load_idx<IdxT>(a.inds, (size_t)i * a.h + h);
        bool real = (j >= 0 && j < a.ns);
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (real) {
            rx = a.s[3 * j] - qx;
            ry = a.s[3 * j + 1] - qy;
            rz = a.s[3 * j + 2] - qz;
        }
        int kmin = 0;
        if (a.aggregation == 1) {  // 'closest': only the nearest kernel point keeps its influence
            float best = 3.4e38f;
            for (int k = 0; k < a.K; k++) {
                float dx = rx - s_kp[3 * k], dy = ry - s_kp[3 * k + 1], dz = rz - s_kp[3 * k + 2];
                float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) {
                    best = d2;
                    kmin = k;
                }
            }
        }
        for (int k = 0; k < a.K; k++) {
            float dx = rx - s_kp[3 * k], dy = ry - s_kp[3 * k + 1], dz = rz - s_kp[3 * k + 2];
            float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float w = influence_w(d2, a.extent, a.influence);
            if (a.aggregation == 1 && k != kmin) w = 0.f;
            bool nz = real && (w > 0.f);
            int base = cnt[k];
            unsigned int m = __ballot_sync(0xffffffffu, nz);
            if (nz) {
                int pos = base + __popc(m & ((1u << lane) - 1));
                jl[k * hcap + pos] = j;
                wl[k * hcap + pos] = w;
            }
            if (lane == 0) cnt[k] = base + __popc(m);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void store_out(float* of, __nv_bfloat16* ohi, __nv_bfloat16* olo,
                                          size_t off, float v) {
    if (of) of[off] = v;
    if (ohi) {
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        ohi[off] = hi;
        olo[off] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

template <typename IdxT, int V>
__global__ void __launch_bounds__(256)
kp_weighted_fwd(KpArgs a, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                __nv_bfloat16* __restrict__ out_lo, int hcap) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float* s_kp = (float*)smem_raw;  // [KP_MAX*3]
    size_t per_warp = (size_t)a.K * hcap * 8 + KP_MAX * 4;
    unsigned char* wbase = smem_raw + KP_MAX * 3 * 4 + wib * per_warp;
    int* jl = (int*)wbase;
    float* wl = (float*)(wbase + (size_t)a.K * hcap * 4);
    int* cnt = (int*)(wbase + (size_t)a.K * hcap * 8);
    for (int t = threadIdx.x; t < a.K * 3; t += blockDim.x) s_kp[t] = a.kp[t];
    __syncthreads();
    const int kd = a.K * a.cin;

    for (int i = blockIdx.x * wpb + wib; i < a.nq; i += gridDim.x * wpb) {
        build_lists<IdxT>(a, i, lane, s_kp, jl, wl, cnt, hcap);
        const size_t row = (size_t)i * a.ld;
        for (int cb = 0; cb < a.cin; cb += 32 * V) {
            const int c = cb + lane * V;
            const bool cok = c < a.cin;  // cin % V == 0 guaranteed by the launcher
            for (int k = 0; k < a.K; k++) {
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = 0.f;
                const int n = cnt[k];
                const int* jk = jl + k * hcap;
                const float* wk = wl + k * hcap;
                if (cok) {
                    int t = 0;
                    for (; t + 4 <= n; t += 4) {
                        float xv[4][V];
                        float w4[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float* xp = a.x + (size_t)jk[t + u] * a.cin + c;
                            w4[u] = wk[t + u];
                            if (V == 4) {
                                float4 t4 = __ldg((const float4*)xp);
                                xv[u][0] = t4.x; xv[u][1 % V] = t4.y; xv[u][2 % V] = t4.z; xv[u][3 % V] = t4.w;
                            } else if (V == 2) {
                                float2 t2 = __ldg((const float2*)xp);
                                xv[u][0] = t2.x; xv[u][1 % V] = t2.y;
                            } else {
                                xv[u][0] = __ldg(xp);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++)
#pragma unroll
                            for (int v = 0; v < V; v++) acc[v] = fmaf(w4[u], xv[u][v], acc[v]);
                    }
                    for (; t < n; t++) {
                        const float* xp = a.x + (size_t)jk[t] * a.cin + c;
                        float w = wk[t];
#pragma unroll
                        for (int v = 0; v < V; v++) acc[v] = fmaf(w, __ldg(xp + v), acc[v]);
                    }
                    const size_t off = row + (size_t)k * a.cin + c;
                    if (out_f32) {
                        if (V == 4) *(float4*)(out_f32 + off) = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
                        else if (V == 2) *(float2*)(out_f32 + off) = make_float2(acc[0], acc[1 % V]);
                        else out_f32[off] = acc[0];
                    }
                    if (out_hi) {
                        __nv_bfloat16 hi[V], lo[V];
#pragma unroll
                        for (int v = 0; v < V; v++) {
                            hi[v] = __float2bfloat16_rn(acc[v]);
                            lo[v] = __float2bfloat16_rn(acc[v] - __bfloat162float(hi[v]));
                        }
                        if (V == 4) {
                            *(uint2*)(out_hi + off) = *(uint2*)hi;
                            *(uint2*)(out_lo + off) = *(uint2*)lo;
                        } else if (V == 2) {
                            *(unsigned int*)(out_hi + off) = *(unsigned int*)hi;
                            *(unsigned int*)(out_lo + off) = *(unsigned int*)lo;
                        } else {
                            out_hi[off] = hi[0];
                            out_lo[off] = lo[0];
                        }
                    }
                }
            }
        }
        for (int c = kd + lane; c < a.width; c += 32) store_out(out_f32, out_hi, out_lo, row + c, 0.f);
        __syncwarp();
    }
}

// grad_x[j, c] += sum_k w_ihk * gw[i, k*cin + c].  Per neighbour: one vector atomic per channel group.
template <typename IdxT, int V>
__global__ void __launch_bounds__(256)
kp_weighted_bwd(KpArgs a, const float* __restrict__ gw, float* __restrict__ gx, int hcap) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float* s_kp = (float*)smem_raw;
    // per warp: lists (K*hcap*8 + KP_MAX*4) + gradient tile [K][32*V] floats
    size_t per_warp = (size_t)a.K * hcap * 8 + KP_MAX * 4 + (size_t)a.K * 32 * V * 4;
    unsigned char* wbase = smem_raw + KP_MAX * 3 * 4 + wib * per_warp;
    int* jl = (int*)wbase;
    float* wl = (float*)(wbase + (size_t)a.K * hcap * 4);
    int* cnt = (int*)(wbase + (size_t)a.K * hcap * 8);
    float* gt = (float*)(wbase + (size_t)a.K * hcap * 8 + KP_MAX * 4);
    for (int t = threadIdx.x; t < a.K * 3; t += blockDim.x) s_kp[t] = a.kp[t];
    __syncthreads();

    for (int i = blockIdx.x * wpb + wib; i < a.nq; i += gridDim.x * wpb) {
        build_lists<IdxT>(a, i, lane, s_kp, jl, wl, cnt, hcap);
        const size_t row = (size_t)i * a.ld;
        for (int cb = 0; cb < a.cin; cb += 32 * V) {
            const int c = cb + lane * V;
            const bool cok = c < a.cin;
            // this lane's gradient columns for every kernel point (own column: no sync needed)
            for (int k = 0; k < a.K; k++) {
#pragma unroll
                for (int v = 0; v < V; v++)
                    gt[(k * 32 + lane) * V + v] = cok ? gw[row + (size_t)k * a.cin + c + v] : 0.f;
            }
            // Walk neighbours in slot order; a neighbour appears in list k at a monotonically
            // increasing cursor (lists are ordered by slot), so keep one cursor per kernel point.
            // Cursors live in registers of lane k (K <= 32) and are broadcast by shuffle.
            int cursor = 0;
            for (int h = 0; h < a.h; h++) {
                int j = load_idx<IdxT>(a.inds, (size_t)i * a.h + h);
                if (j < 0 || j >= a.ns) continue;  // warp-uniform
                // which kernel points list this neighbour next?
                bool mine = false;
                if (lane < a.K && cursor < cnt[lane]) mine = (jl[lane * hcap + cursor] == j);
                unsigned int m = __ballot_sync(0xffffffffu, mine);
                if (m == 0) continue;
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = 0.f;
                unsigned int mm = m;
                while (mm) {
                    int k = __ffs(mm) - 1;
                    mm &= mm - 1;
                    int cur = __shfl_sync(0xffffffffu, cursor, k);
                    float w = wl[k * hcap + cur];
#pragma unroll
                    for (int v = 0; v < V; v++) acc[v] = fmaf(w, gt[(k * 32 + lane) * V + v], acc[v]);
                }
                if (mine) cursor++;
                if (cok) {
                    float* dst = gx + (size_t)j * a.cin + c;
                    if (V == 4) atomicAdd((float4*)dst, make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]));
                    else if (V == 2) atomicAdd((float2*)dst, make_float2(acc[0], acc[1 % V]));
                    else atomicAdd(dst, acc[0]);
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, int rows, int cols, int src_ld,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int rows_pad, int ld) {
    pdl_enter();
    size_t total = (size_t)rows_pad * ld;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(t / ld), c = (int)(t % ld);
        float v = (r < rows && c < cols) ? src[(size_t)r * src_ld + c] : 0.f;
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[t] = h;
        if (lo) lo[t] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// dense [rows, 4*cv] case: one float4 in, two 8-byte stores out per thread and trip
__global__ void __launch_bounds__(256)
split_bf16_vec4(const float* __restrict__ src, int rows, int cv, int src_ld, __nv_bfloat16* __restrict__ hi,
                __nv_bfloat16* __restrict__ lo) {
    pdl_enter();
    const size_t total = (size_t)rows * cv;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(t / cv), c = (int)(t % cv) * 4;
        const float4 v = *(const float4*)(src + (size_t)r * src_ld + c);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
        __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        uint2 ph, pl;
        ph.x = *(unsigned int*)&h0; ph.y = *(unsigned int*)&h1;
        pl.x = *(unsigned int*)&l0; pl.y = *(unsigned int*)&l1;
        *(uint2*)(hi + t * 4) = ph;
        *(uint2*)(lo + t * 4) = pl;
    }
}

// Every weight matrix of a model in one launch (the parameters change once per optimiser step: one split
// per step instead of one tiny launch per layer).  CTA b serves 4096 destination elements of the tensor
// whose chunk range contains b.
__global__ void __launch_bounds__(256)
split_bf16_multi(const mvk_split_desc* __restrict__ table, int n) {
    pdl_enter();
    __shared__ int s_i;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].first_chunk <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_i = lo;
    }
    __syncthreads();
    const mvk_split_desc d = table[s_i];
    const size_t total = (size_t)d.rows_pad * d.dst_ld;  // dst_ld is even (16-byte row pitch)
    const size_t base = (size_t)((int)blockIdx.x - d.first_chunk) * 4096;
    const size_t end = base + 4096 < total ? base + 4096 : total;
    __nv_bfloat16* hi = (__nv_bfloat16*)d.hi;
    __nv_bfloat16* lo = (__nv_bfloat16*)d.lo;
    for (size_t t = base + 2 * threadIdx.x; t < end; t += 512) {
        const int r = (int)(t / d.dst_ld), c = (int)(t % d.dst_ld);
        const float* row = d.src + (size_t)r * d.src_ld;
        const float v0 = (r < d.rows && c < d.cols) ? row[c] : 0.f;
        const float v1 = (r < d.rows && c + 1 < d.cols) ? row[c + 1] : 0.f;
        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        const float2 f = __bfloat1622float2(h);
        *(__nv_bfloat162*)(hi + t) = h;
        *(__nv_bfloat162*)(lo + t) = __floats2bfloat162_rn(v0 - f.x, v1 - f.y);
    }
}

// Decoder entry of KPFCNN (architectures.py:300-306): x = cat([closest_pool(x_coarse, up_inds), skip], dim=1)
// feeding a unary block.  The concatenated row only ever exists as the bf16 hi/lo operand of that block's
// Linear: row r = [x_coarse[inds[r, 0]] (zero for the shadow index), skip[r]].  One float4 per thread and trip.
template <typename IdxT>
__global__ void __launch_bounds__(256)
upcat_split_kernel(const float* __restrict__ xc, int ns, int c1, const IdxT* __restrict__ inds, int h,
                   const float* __restrict__ skip, int lds, int c2, int nq, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo, int ldh) {
    pdl_enter();
    const int cv = (c1 + c2) / 4, cv1 = c1 / 4;
    const size_t total = (size_t)nq * cv;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(t / cv), c = (int)(t % cv);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < cv1) {
            const long long j = (long long)inds[(size_t)r * h];
            if (j >= 0 && j < ns) v = *(const float4*)(xc + (size_t)j * c1 + 4 * c);
        } else {
            v = *(const float4*)(skip + (size_t)r * lds + 4 * (c - cv1));
        }
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
        __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        uint2 ph, pl;
        ph.x = *(unsigned int*)&h0; ph.y = *(unsigned int*)&h1;
        pl.x = *(unsigned int*)&l0; pl.y = *(unsigned int*)&l1;
        *(uint2*)(hi + (size_t)r * ldh + 4 * c) = ph;
        *(uint2*)(lo + (size_t)r * ldh + 4 * c) = pl;
    }
}

// mode 0: max over neighbours with a ZERO shadow row (blocks.py:93-110); mode 1: first column.
template <typename IdxT>
__global__ void __launch_bounds__(256)
pool_fwd(const float* __restrict__ x, int ns, int c, const void* __restrict__ inds, int nq, int h,
         int mode, float* __restrict__ out, int* __restrict__ arg) {
    pdl_enter();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nq; i += gridDim.x * wpb) {
        for (int ch = lane; ch < c; ch += 32) {
            if (mode == 1) {
                int j = load_idx<IdxT>(inds, (size_t)i * h);
                out[(size_t)i * c + ch] = (j >= 0 && j < ns) ? x[(size_t)j * c + ch] : 0.f;
            } else {
                float best = 0.f;
                int bj = ns;
                bool first = true;
                for (int t = 0; t < h; t++) {
                    int j = load_idx<IdxT>(inds, (size_t)i * h + t);
                    bool real = j >= 0 && j < ns;
                    float v = real ? x[(size_t)j * c + ch] : 0.f;
                    if (first || v > best) {
                        best = v;
                        bj = real ? j : ns;
                        first = false;
                    }
                }
                out[(size_t)i * c + ch] = best;
                if (arg) arg[(size_t)i * c + ch] = bj;
            }
        }
    }
}

// float4 flavour (c % 4 == 0): one 128-bit load per lane and neighbour, indices read once per row.
template <typename IdxT>
__global__ void __launch_bounds__(256)
pool_fwd_vec4(const float* __restrict__ x, int ns, int c, const void* __restrict__ inds, int nq, int h,
              int mode, float* __restrict__ out, int* __restrict__ arg) {
    pdl_enter();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int ncg = (c + 127) / 128;  // work item = (query, group of 128 channels): deep layers have few queries
    const long long items = (long long)nq * ncg;
    for (long long it = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); it < items; it += (long long)gridDim.x * wpb) {
        const int i = (int)(it / ncg);
        {
            const int ch = (int)(it % ncg) * 128 + lane * 4;
            if (ch >= c) continue;
            float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
            int4 bj = make_int4(ns, ns, ns, ns);
            const int hh = mode == 1 ? 1 : h;
            for (int t = 0; t < hh; t++) {
                const int j = load_idx<IdxT>(inds, (size_t)i * h + t);
                const bool real = j >= 0 && j < ns;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (real) v = __ldg((const float4*)(x + (size_t)j * c + ch));
                const int jj = real ? j : ns;
                if (t == 0 || v.x > best.x) { best.x = v.x; bj.x = jj; }
                if (t == 0 || v.y > best.y) { best.y = v.y; bj.y = jj; }
                if (t == 0 || v.z > best.z) { best.z = v.z; bj.z = jj; }
                if (t == 0 || v.w > best.w) { best.w = v.w; bj.w = jj; }
            }
            *(float4*)(out + (size_t)i * c + ch) = best;
            if (arg && mode == 0) *(int4*)(arg + (size_t)i * c + ch) = bj;
        }
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
pool_bwd(const float* __restrict__ go, int ldg, int nq, int c, const int* __restrict__ arg,
         const void* __restrict__ inds, int h, int mode, int ns, float* __restrict__ gx) {
    pdl_enter();
    size_t total = (size_t)nq * c;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(t / c), ch = (int)(t % c);
        int j = mode == 1 ? load_idx<IdxT>(inds, (size_t)i * h) : arg[t];
        if (j >= 0 && j < ns) atomicAdd(&gx[(size_t)j * c + ch], go[(size_t)i * ldg + ch]);
    }
}

// closest_pool backward (nearest upsampling of the decoder): the whole row goes to ONE support row, so a thread
// moves four channels with one 128-bit vector reduction (4x fewer L2 atomics than the element-wise kernel above).
template <typename IdxT>
__global__ void __launch_bounds__(256)
pool_bwd_closest_vec4(const float* __restrict__ go, int ldg, int nq, int c, const void* __restrict__ inds, int h, int ns,
                      float* __restrict__ gx) {
    pdl_enter();
    const int cv = c >> 2;
    const size_t total = (size_t)nq * cv;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / cv), ch = (int)(t % cv) * 4;
        const int j = load_idx<IdxT>(inds, (size_t)i * h);
        if (j >= 0 && j < ns) {
            const float4 v = *(const float4*)(go + (size_t)i * ldg + ch);
            atomicAdd((float4*)(gx + (size_t)j * c + ch), v);
        }
    }
}

// =================================================================================================
// Fast paths (KP_influence = 'linear', aggregation = 'sum', K <= 16): the configuration every
// reference script uses (utils/config.py, train_ScanNet_*.py).  Everything else keeps the generic
// kernels above.
//
// Distances use the expansion d2 = |r|^2 - 2 r.kp + |kp|^2 with per-kernel-point constants staged
// in shared memory (one LDS.128 per kernel point), sqrt.approx and a multiply by 1/extent: the
// influence is accurate to a few 1e-7 absolute, far inside the 1e-4 feature tolerance (only the
// neighbour / voxel INDEX kernels need bit-exact arithmetic).
// =================================================================================================
constexpr int KF = 16;  // kernel-point slots of the fast paths

__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// c[k] = (-2 kx, -2 ky, -2 kz, |kp|^2); unused slots get a huge constant -> influence < 0.
__device__ __forceinline__ void stage_kp_constants(const KpArgs& a, float4* s_kc) {
    if (threadIdx.x < KF) {
        int k = threadIdx.x;
        float4 c = make_float4(0.f, 0.f, 0.f, 1e30f);
        if (k < a.K) {
            float x = a.kp[3 * k], y = a.kp[3 * k + 1], z = a.kp[3 * k + 2];
            c = make_float4(-2.f * x, -2.f * y, -2.f * z, x * x + y * y + z * z);
        }
        s_kc[k] = c;
    }
    __syncthreads();
}

__device__ __forceinline__ float influence_fast(float rx, float ry, float rz, float r2, float4 c, float inv_ext) {
    float d2 = fmaf(rx, c.x, fmaf(ry, c.y, fmaf(rz, c.z, c.w))) + r2;
    return fmaf(-sqrt_approx(fmaxf(d2, 0.f)), inv_ext, 1.f);
}

__device__ __forceinline__ void store4(float* of, __nv_bfloat16* ohi, __nv_bfloat16* olo, size_t off, float4 v) {
    if (of) {
        *(float4*)(of + off) = v;
    } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
        __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        uint2 ph, pl;
        ph.x = *(unsigned int*)&h0; ph.y = *(unsigned int*)&h1;
        pl.x = *(unsigned int*)&l0; pl.y = *(unsigned int*)&l1;
        *(uint2*)(ohi + off) = ph;
        *(uint2*)(olo + off) = pl;
    }
}

// ---- forward, cin in {32, 64, 128 m} ---------------------------------------------------------------
// One warp per query point.
//   phase 1  lanes = neighbours: influences of the 16 kernel-point slots; non-zero entries are
//            ballot-compacted into per-kernel-point lists {byte offset of the support row, weight}
//            in shared memory (list counters live in registers).
//   phase 2  the warp splits into SUB = 32/G sub-groups of G lanes (G*4 = channels per pass, one
//            float4 per lane); in round r sub-group g walks the list of kernel point r*SUB + g,
//            four entries per trip (four independent 128-bit gathers in flight), accumulating in
//            registers; the [K, cin] row of the weighted matrix leaves as 8-byte bf16 hi / lo
//            stores (or float4).
template <typename IdxT, int G, int HCAP>
__global__ void __launch_bounds__(256)
kp_fwd_fast(KpArgs a, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
            __nv_bfloat16* __restrict__ out_lo) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SUB = 32 / G;
    constexpr int CW = G * 4;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float4* s_kc = (float4*)smem_raw;                                        // [KF]
    int2* lists = (int2*)(smem_raw + KF * 16) + (size_t)wib * KF * HCAP;     // [KF][HCAP]
    // finite weights everywhere: the 4-wide trips of phase 2 read (and ignore) entries past a list's end
    for (int t = lane; t < KF * HCAP; t += 32) lists[t] = make_int2(0, 0);
    stage_kp_constants(a, s_kc);
    const float inv_ext = 1.f / a.extent;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int g = lane / G, lg = lane % G;
    const int cin = a.cin, H = a.h, ns = a.ns, nq = a.nq;
    const unsigned int row_bytes = (unsigned int)cin * 4u;
    const int kd = a.K * cin;
    const int2* lists_g = lists + g * HCAP;
    const int stride = gridDim.x * wpb;

    int i = blockIdx.x * wpb + wib;
    int j_next = ns;
    if (i < nq && lane < H) j_next = load_idx<IdxT>(a.inds, (size_t)i * H + lane);
    for (; i < nq; i += stride) {
        // ---------------- phase 1 ----------------
        int cnt[KF];
#pragma unroll
        for (int k = 0; k < KF; k++) cnt[k] = 0;
        const float qx = a.q[3 * i], qy = a.q[3 * i + 1], qz = a.q[3 * i + 2];
        int j = j_next;
        {   // prefetch the first index chunk of this warp's next point
            const int in = i + stride;
            j_next = ns;
            if (in < nq && lane < H) j_next = load_idx<IdxT>(a.inds, (size_t)in * H + lane);
        }
        for (int h0 = 0; h0 < H; h0 += 32) {
            if (h0 > 0) {
                const int h = h0 + lane;
                j = ns;
                if (h < H) j = load_idx<IdxT>(a.inds, (size_t)i * H + h);
            }
            const bool real = (unsigned int)j < (unsigned int)ns;
            if (__ballot_sync(0xffffffffu, real) == 0u) continue;
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if (real) {
                rx = __ldg(a.s + 3 * (size_t)j) - qx;
                ry = __ldg(a.s + 3 * (size_t)j + 1) - qy;
                rz = __ldg(a.s + 3 * (size_t)j + 2) - qz;
            }
            const float r2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
            const int2 ent = make_int2((int)((unsigned int)j * row_bytes), 0);
#pragma unroll
            for (int k = 0; k < KF; k++) {
                const float w = influence_fast(rx, ry, rz, r2, s_kc[k], inv_ext);
                const bool nz = real && (w > 0.f);
                const unsigned int m = __ballot_sync(0xffffffffu, nz);
                if (nz) lists[k * HCAP + cnt[k] + __popc(m & lt_mask)] = make_int2(ent.x, __float_as_int(w));
                cnt[k] += __popc(m);
            }
        }
        __syncwarp();

        // ---------------- phase 2 ----------------
        const size_t row = (size_t)i * a.ld;
        for (int cb = 0; cb < cin; cb += CW) {
            const char* xb = (const char*)(a.x + cb + lg * 4);
#pragma unroll
            for (int r = 0; r < KF / SUB; r++) {
                int n = cnt[r * SUB];
#pragma unroll
                for (int u = 1; u < SUB; u++) n = (g == u) ? cnt[r * SUB + u] : n;
                const int2* lk = lists_g + r * SUB * HCAP;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
                for (int t = 0; t < n; t += 4) {
                    const int4 e01 = *(const int4*)(lk + t);
                    const int4 e23 = *(const int4*)(lk + t + 2);
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 x0 = __ldg((const float4*)(xb + (unsigned int)e01.x));
                    float4 x1 = z, x2 = z, x3 = z;
                    if (t + 1 < n) x1 = __ldg((const float4*)(xb + (unsigned int)e01.z));
                    if (t + 2 < n) x2 = __ldg((const float4*)(xb + (unsigned int)e23.x));
                    if (t + 3 < n) x3 = __ldg((const float4*)(xb + (unsigned int)e23.z));
                    const float w0 = __int_as_float(e01.y), w1 = __int_as_float(e01.w);
                    const float w2 = __int_as_float(e23.y), w3 = __int_as_float(e23.w);
                    acc.x = fmaf(w0, x0.x, acc.x); acc.y = fmaf(w0, x0.y, acc.y);
                    acc.z = fmaf(w0, x0.z, acc.z); acc.w = fmaf(w0, x0.w, acc.w);
                    acc.x = fmaf(w1, x1.x, acc.x); acc.y = fmaf(w1, x1.y, acc.y);
                    acc.z = fmaf(w1, x1.z, acc.z); acc.w = fmaf(w1, x1.w, acc.w);
                    acc.x = fmaf(w2, x2.x, acc.x); acc.y = fmaf(w2, x2.y, acc.y);
                    acc.z = fmaf(w2, x2.z, acc.z); acc.w = fmaf(w2, x2.w, acc.w);
                    acc.x = fmaf(w3, x3.x, acc.x); acc.y = fmaf(w3, x3.y, acc.y);
                    acc.z = fmaf(w3, x3.z, acc.z); acc.w = fmaf(w3, x3.w, acc.w);
                }
                const int k = r * SUB + g;
                if (k < a.K) store4(out_f32, out_hi, out_lo, row + (unsigned int)(k * cin + cb + lg * 4), acc);
            }
        }
        for (int c = kd + lane; c < a.width; c += 32) store_out(out_f32, out_hi, out_lo, row + c, 0.f);
        __syncwarp();
    }
}

// ---- forward, cin <= 4 (first layer of every network: 2, 4 input features) -----------------------
// One THREAD per query point, dense over the kernel points (cin FMAs per (neighbour, kernel point)
// are cheaper than any sparsity bookkeeping); K*CIN accumulators in registers.
template <typename IdxT, int CIN>
__global__ void __launch_bounds__(128)
kp_fwd_tiny(KpArgs a, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
            __nv_bfloat16* __restrict__ out_lo) {
    pdl_enter();
    __shared__ float4 s_kc[KF];
    stage_kp_constants(a, s_kc);
    const float inv_ext = 1.f / a.extent;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nq) return;
    float acc[KF][CIN];
#pragma unroll
    for (int k = 0; k < KF; k++)
#pragma unroll
        for (int c = 0; c < CIN; c++) acc[k][c] = 0.f;
    const float qx = a.q[3 * i], qy = a.q[3 * i + 1], qz = a.q[3 * i + 2];
    for (int h = 0; h < a.h; h++) {
        const int j = load_idx<IdxT>(a.inds, (size_t)i * a.h + h);
        if ((unsigned int)j >= (unsigned int)a.ns) continue;
        const float rx = __ldg(a.s + 3 * (size_t)j) - qx;
        const float ry = __ldg(a.s + 3 * (size_t)j + 1) - qy;
        const float rz = __ldg(a.s + 3 * (size_t)j + 2) - qz;
        float xv[CIN];
#pragma unroll
        for (int c = 0; c < CIN; c++) xv[c] = __ldg(a.x + (size_t)j * CIN + c);
        const float r2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
#pragma unroll
        for (int k = 0; k < KF; k++) {
            const float w = fmaxf(influence_fast(rx, ry, rz, r2, s_kc[k], inv_ext), 0.f);
#pragma unroll
            for (int c = 0; c < CIN; c++) acc[k][c] = fmaf(w, xv[c], acc[k][c]);
        }
    }
    const size_t row = (size_t)i * a.ld;
    if (out_hi && a.width == a.ld && (a.ld % 8) == 0 && a.ld <= KF * 4) {
        // the whole bf16 row (zero padding included) as ld/8 + ld/8 128-bit stores
#pragma unroll
        for (int v8 = 0; v8 < 8; v8++) {
            if (v8 * 8 >= a.ld) break;
            unsigned int ph[4], pl[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float f[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = v8 * 8 + u * 2 + e;  // compile-time
                    const int k = col / CIN, c = col % CIN;
                    f[e] = (k < KF) ? ((k < a.K) ? acc[k < KF ? k : 0][c] : 0.f) : 0.f;
                }
                __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
                float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 l = __floats2bfloat162_rn(f[0] - hf.x, f[1] - hf.y);
                ph[u] = *(unsigned int*)&h;
                pl[u] = *(unsigned int*)&l;
            }
            *(uint4*)(out_hi + row + v8 * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *(uint4*)(out_lo + row + v8 * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < KF; k++)
#pragma unroll
        for (int c = 0; c < CIN; c++)
            if (k < a.K) store_out(out_f32, out_hi, out_lo, row + k * CIN + c, acc[k][c]);
    for (int c = a.K * CIN; c < a.width; c++) store_out(out_f32, out_hi, out_lo, row + c, 0.f);
}

// ---- backward w.r.t. x, cin in {32, 64, 128 m} ---------------------------------------------------------
// One warp per query point.
//   phase 1  lanes = neighbours; every lane appends its own non-zero {kernel point, weight} entries
//            to a per-neighbour list (slot-major in shared memory: no cross-lane traffic at all).
//            Meanwhile the [K, cin] gradient row of the weighted matrix streams into shared memory
//            with cp.async.
//   phase 2  sub-group g combines the entries of neighbours g, g+SUB, ... and issues ONE 128-bit
//            vector atomic per lane and neighbour into grad_x.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename IdxT, int G, int HCAP>
__global__ void __launch_bounds__(256)
kp_bwd_fast(KpArgs a, const float* __restrict__ gw, float* __restrict__ gx) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SUB = 32 / G;
    constexpr int CW = G * 4;  // channels per pass
    constexpr size_t PER_WARP = (size_t)KF * HCAP * 8 + (size_t)HCAP * 8 + (size_t)KF * CW * 4;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float4* s_kc = (float4*)smem_raw;
    unsigned char* wbase = smem_raw + KF * 16 + wib * PER_WARP;
    int2* ent = (int2*)wbase;                                       // [KF slots][HCAP] {tile byte offset, weight}
    int2* meta = (int2*)(wbase + (size_t)KF * HCAP * 8);            // [HCAP] {entries, byte offset of the grad_x row}
    float* tile = (float*)(wbase + (size_t)KF * HCAP * 8 + (size_t)HCAP * 8);  // [KF][CW]
    for (int t = lane; t < KF * HCAP; t += 32) ent[t] = make_int2(0, 0);
    stage_kp_constants(a, s_kc);
    const float inv_ext = 1.f / a.extent;
    const int g = lane / G, lg = lane % G;
    const int cin = a.cin, H = a.h, ns = a.ns, nq = a.nq;
    const unsigned int row_bytes = (unsigned int)cin * 4u;
    const int kd = a.K * cin;
    const int stride = gridDim.x * wpb;
    const char* tile_l = (const char*)tile + lg * 16;

    int i = blockIdx.x * wpb + wib;
    int j_next = ns;
    if (i < nq && lane < H) j_next = load_idx<IdxT>(a.inds, (size_t)i * H + lane);
    for (; i < nq; i += stride) {
        const size_t row = (size_t)i * a.ld;
        // gradient rows of pass 0 -> shared memory, asynchronously (consumed after phase 1)
        for (int t = lane * 4; t < KF * CW; t += 128) {
            const int k = t / CW, c = t % CW;
            if (k < a.K) cp_async16(tile + t, gw + row + (unsigned int)(k * cin + c));
        }
        const float qx = a.q[3 * i], qy = a.q[3 * i + 1], qz = a.q[3 * i + 2];
        int j = j_next;
        {
            const int in = i + stride;
            j_next = ns;
            if (in < nq && lane < H) j_next = load_idx<IdxT>(a.inds, (size_t)in * H + lane);
        }
        for (int h0 = 0; h0 < H; h0 += 32) {
            const int h = h0 + lane;
            if (h0 > 0) {
                j = ns;
                if (h < H) j = load_idx<IdxT>(a.inds, (size_t)i * H + h);
            }
            const bool real = (unsigned int)j < (unsigned int)ns;
            int n = 0;
            if (__ballot_sync(0xffffffffu, real) != 0u) {
                float rx = 0.f, ry = 0.f, rz = 0.f;
                if (real) {
                    rx = __ldg(a.s + 3 * (size_t)j) - qx;
                    ry = __ldg(a.s + 3 * (size_t)j + 1) - qy;
                    rz = __ldg(a.s + 3 * (size_t)j + 2) - qz;
                }
                const float r2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
#pragma unroll
                for (int k = 0; k < KF; k++) {
                    const float w = influence_fast(rx, ry, rz, r2, s_kc[k], inv_ext);
                    if (real && w > 0.f) {
                        ent[n * HCAP + h] = make_int2(k * CW * 4, __float_as_int(w));
                        n++;
                    }
                }
            }
            if (h < HCAP) meta[h] = make_int2(n, (int)((unsigned int)j * row_bytes));
        }
        cp_async_wait_all();
        __syncwarp();
        for (int cb = 0; cb < cin; cb += CW) {
            if (cb > 0) {
                __syncwarp();
                for (int t = lane * 4; t < KF * CW; t += 128) {
                    const int k = t / CW, c = t % CW;
                    if (k < a.K) cp_async16(tile + t, gw + row + (unsigned int)(k * cin + cb + c));
                }
                cp_async_wait_all();
                __syncwarp();
            }
            char* gxb = (char*)(gx + cb + lg * 4);
            for (int h = g; h < H; h += SUB) {
                const int2 m = meta[h];
                const int n = m.x;
                if (n == 0) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
                for (int t = 0; t < n; t += 2) {
                    const int2 e0 = ent[t * HCAP + h];
                    const int2 e1 = ent[(t + 1) * HCAP + h];
                    const float w0 = __int_as_float(e0.y);
                    const float w1 = (t + 1 < n) ? __int_as_float(e1.y) : 0.f;
                    const float4 v0 = *(const float4*)(tile_l + e0.x);
                    const float4 v1 = *(const float4*)(tile_l + e1.x);
                    acc.x = fmaf(w0, v0.x, acc.x); acc.y = fmaf(w0, v0.y, acc.y);
                    acc.z = fmaf(w0, v0.z, acc.z); acc.w = fmaf(w0, v0.w, acc.w);
                    acc.x = fmaf(w1, v1.x, acc.x); acc.y = fmaf(w1, v1.y, acc.y);
                    acc.z = fmaf(w1, v1.z, acc.z); acc.w = fmaf(w1, v1.w, acc.w);
                }
                atomicAdd((float4*)(gxb + (unsigned int)m.y), acc);
            }
        }
        __syncwarp();
    }
    (void)kd;
}

// y[r, :] += z[inds[r, 0], :]  (rows whose first index is the shadow index keep y): the upsampled half of the decoder
// Linear, contracted at the COARSE level and gathered here -- cat([up(x), skip]) W^T = up(x W_up^T) + skip W_skip^T.
template <typename IdxT>
__global__ void __launch_bounds__(256)
gather_add_rows_kernel(float* __restrict__ y, int ldy, int nq, int cv, const float* __restrict__ z, int ns,
                       const void* __restrict__ inds, int h) {
    pdl_enter();
    const size_t total = (size_t)nq * cv;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(t / cv), c = (int)(t % cv) * 4;
        const int j = load_idx<IdxT>(inds, (size_t)r * h);
        if (j < 0 || j >= ns) continue;
        float4 a = *(const float4*)(y + (size_t)r * ldy + c);
        const float4 b = __ldg((const float4*)(z + (size_t)j * (cv * 4) + c));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        *(float4*)(y + (size_t)r * ldy + c) = a;
    }
}

bool fast_ok(int cin, int num_kp, int influence, int aggregation, int ld) {
    return influence == 1 && aggregation == 0 && num_kp <= KF && (cin == 32 || cin == 64 || cin % 128 == 0) &&
           (ld % 4 == 0);
}
bool tiny_ok(int cin, int num_kp, int influence, int aggregation) {
    return influence == 1 && aggregation == 0 && num_kp <= KF && cin <= 4;
}
int fast_g(int cin) { return cin == 32 ? 8 : (cin == 64 ? 16 : 32); }

// channels per lane of the generic kernels: any width that keeps the 16 / 8-byte row alignment of x
int pick_v(int cin) { return (cin % 4 == 0) ? 4 : ((cin % 2 == 0) ? 2 : 1); }

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

int mvk_kpconv_weighted_part(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                             int idx_is_i64, int h, const float* x, int cin, const float* kernel_points, int num_kp,
                             float kp_extent, int influence, int aggregation, int ld, int width, float* out_f32,
                             void* out_hi, void* out_lo, mvk_stream_t stream);

int mvk_kpconv_weighted(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                        int idx_is_i64, int h, const float* x, int cin, const float* kernel_points,
                        int num_kp, float kp_extent, int influence, int aggregation, int ld,
                        float* out_f32, void* out_hi, void* out_lo, mvk_stream_t stream) {
    return mvk_kpconv_weighted_part(q_pts, nq, s_pts, ns, neighb_inds, idx_is_i64, h, x, cin, kernel_points, num_kp,
                                    kp_extent, influence, aggregation, ld, ld, out_f32, out_hi, out_lo, stream);
}

int mvk_kpconv_weighted_part(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                             int idx_is_i64, int h, const float* x, int cin, const float* kernel_points, int num_kp,
                             float kp_extent, int influence, int aggregation, int ld, int width, float* out_f32,
                             void* out_hi, void* out_lo, mvk_stream_t stream) {
    if (nq < 0 || ns < 0 || h < 1 || cin < 1 || num_kp < 1 || num_kp > KP_MAX || ld < num_kp * cin || width < num_kp * cin ||
        width > ld || (!out_f32 && !(out_hi && out_lo)) || !(kp_extent > 0.f))
        return MVK_ERR_INVALID_ARG;
    if (influence < 0 || influence > 2 || aggregation < 0 || aggregation > 1) return MVK_ERR_UNSUPPORTED;
    if (nq == 0) return MVK_OK;
    KpArgs a{q_pts, s_pts, neighb_inds, x, kernel_points, nq, ns, h, cin, num_kp, ld, kp_extent,
             influence, aggregation, width};
    cudaStream_t st = (cudaStream_t)stream;
    if (tiny_ok(cin, num_kp, influence, aggregation)) {
        const int threads = 128, blocks_t = (nq + threads - 1) / threads;
#define LAUNCH_TINY(IDX, C)                                                                          \
    launch_pdl(kp_fwd_tiny<IDX, C>, dim3(blocks_t), dim3(threads), 0, st, 1, a, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo)
#define LAUNCH_TINY_C(IDX)                                                                           \
    do {                                                                                             \
        if (cin == 1) LAUNCH_TINY(IDX, 1);                                                           \
        else if (cin == 2) LAUNCH_TINY(IDX, 2);                                                      \
        else if (cin == 3) LAUNCH_TINY(IDX, 3);                                                      \
        else LAUNCH_TINY(IDX, 4);                                                                    \
    } while (0)
        if (idx_is_i64) LAUNCH_TINY_C(long long);
        else LAUNCH_TINY_C(int);
#undef LAUNCH_TINY_C
#undef LAUNCH_TINY
        MVK_LAUNCHED("kp_fwd_tiny");
        return MVK_OK;
    }
    if (fast_ok(cin, num_kp, influence, aggregation, ld) && h <= 64 && (size_t)(ns + 1) * cin * 4 < 0xffffffffull &&
        (size_t)ld < 0x7fffffffull) {
        const int HC = h <= 48 ? 48 : 64;
        const size_t pw = (size_t)KF * HC * 8;
        int wpb_f = (int)((44 * 1024) / pw);
        wpb_f = wpb_f > 8 ? 8 : wpb_f;
        const size_t smem_f = KF * 16 + pw * wpb_f;
        int blocks_f = (nq + wpb_f - 1) / wpb_f;
        const int maxb_f = num_sms() * 32;
        if (blocks_f > maxb_f) blocks_f = maxb_f;
        const int G = fast_g(cin);
#define LAUNCH_FAST(IDX, GG, HH)                                                                     \
    do {                                                                                             \
        auto kern = kp_fwd_fast<IDX, GG, HH>;                                                        \
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f)); \
        launch_pdl(kern, dim3(blocks_f), dim3(wpb_f * 32), smem_f, st, 1, a, out_f32, (__nv_bfloat16*)out_hi,               \
                                                   (__nv_bfloat16*)out_lo);                          \
    } while (0)
#define LAUNCH_FAST_G(IDX, HH)                                                                       \
    do {                                                                                             \
        if (G == 8) LAUNCH_FAST(IDX, 8, HH);                                                         \
        else if (G == 16) LAUNCH_FAST(IDX, 16, HH);                                                  \
        else LAUNCH_FAST(IDX, 32, HH);                                                               \
    } while (0)
        if (idx_is_i64) {
            if (HC == 48) LAUNCH_FAST_G(long long, 48);
            else LAUNCH_FAST_G(long long, 64);
        } else {
            if (HC == 48) LAUNCH_FAST_G(int, 48);
            else LAUNCH_FAST_G(int, 64);
        }
#undef LAUNCH_FAST_G
#undef LAUNCH_FAST
        MVK_LAUNCHED("kp_fwd_fast");
        return MVK_OK;
    }
    int hcap = (h + 31) / 32 * 32;
    size_t per_warp = (size_t)num_kp * hcap * 8 + KP_MAX * 4;
    if (per_warp + 512 > 200 * 1024) return MVK_ERR_RANGE;
    int wpb = (int)((96 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
    size_t smem = KP_MAX * 3 * 4 + per_warp * wpb;
    int blocks = (nq + wpb - 1) / wpb;
    int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    int V = pick_v(cin);
    // vector stores need 16B/8B aligned rows
    if (V == 4 && (ld % 4 != 0)) V = 1;
    if (V == 2 && (ld % 2 != 0)) V = 1;
#define LAUNCH_FWD(IDX, VV)                                                                          \
    do {                                                                                             \
        auto kern = kp_weighted_fwd<IDX, VV>;                                                        \
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        launch_pdl(kern, dim3(blocks), dim3(wpb * 32), smem, st, 1, a, out_f32, (__nv_bfloat16*)out_hi,                     \
                                             (__nv_bfloat16*)out_lo, hcap);                          \
    } while (0)
    if (idx_is_i64) {
        if (V == 4) LAUNCH_FWD(long long, 4);
        else if (V == 2) LAUNCH_FWD(long long, 2);
        else LAUNCH_FWD(long long, 1);
    } else {
        if (V == 4) LAUNCH_FWD(int, 4);
        else if (V == 2) LAUNCH_FWD(int, 2);
        else LAUNCH_FWD(int, 1);
    }
#undef LAUNCH_FWD
    MVK_LAUNCHED("kp_weighted_fwd");
    return MVK_OK;
}

int mvk_kpconv_weighted_bwd(const float* q_pts, int nq, const float* s_pts, int ns,
                            const void* neighb_inds, int idx_is_i64, int h, int cin,
                            const float* kernel_points, int num_kp, float kp_extent, int influence,
                            int aggregation, const float* grad_weighted, int ld, float* grad_x,
                            mvk_stream_t stream) {
    if (nq < 0 || ns < 0 || h < 1 || cin < 1 || num_kp < 1 || num_kp > KP_MAX || ld < num_kp * cin ||
        !grad_weighted || !grad_x || !(kp_extent > 0.f))
        return MVK_ERR_INVALID_ARG;
    if (influence < 0 || influence > 2 || aggregation < 0 || aggregation > 1) return MVK_ERR_UNSUPPORTED;
    if (nq == 0) return MVK_OK;
    KpArgs a{q_pts, s_pts, neighb_inds, nullptr, kernel_points, nq, ns, h, cin, num_kp, ld, kp_extent,
             influence, aggregation, ld};
    if (fast_ok(cin, num_kp, influence, aggregation, ld) && h <= 64 && (size_t)(ns + 1) * cin * 4 < 0xffffffffull &&
        (size_t)ld < 0x7fffffffull) {
        const int G = fast_g(cin);
        const int HC = h <= 48 ? 48 : 64;
        const size_t pw = (size_t)KF * HC * 8 + (size_t)HC * 8 + (size_t)KF * G * 4 * 4;
        int wpb_f = (int)((44 * 1024) / pw);
        wpb_f = wpb_f < 1 ? 1 : (wpb_f > 8 ? 8 : wpb_f);
        const size_t smem_f = KF * 16 + pw * wpb_f;
        int blocks_f = (nq + wpb_f - 1) / wpb_f;
        const int maxb_f = num_sms() * 32;
        if (blocks_f > maxb_f) blocks_f = maxb_f;
        cudaStream_t st_f = (cudaStream_t)stream;
#define LAUNCH_FASTB(IDX, GG, HH)                                                                    \
    do {                                                                                             \
        auto kern = kp_bwd_fast<IDX, GG, HH>;                                                        \
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f)); \
        launch_pdl(kern, dim3(blocks_f), dim3(wpb_f * 32), smem_f, st_f, 1, a, grad_weighted, grad_x);                      \
    } while (0)
#define LAUNCH_FASTB_G(IDX, HH)                                                                      \
    do {                                                                                             \
        if (G == 8) LAUNCH_FASTB(IDX, 8, HH);                                                        \
        else if (G == 16) LAUNCH_FASTB(IDX, 16, HH);                                                 \
        else LAUNCH_FASTB(IDX, 32, HH);                                                              \
    } while (0)
        if (idx_is_i64) {
            if (HC == 48) LAUNCH_FASTB_G(long long, 48);
            else LAUNCH_FASTB_G(long long, 64);
        } else {
            if (HC == 48) LAUNCH_FASTB_G(int, 48);
            else LAUNCH_FASTB_G(int, 64);
        }
#undef LAUNCH_FASTB_G
#undef LAUNCH_FASTB
        MVK_LAUNCHED("kp_bwd_fast");
        return MVK_OK;
    }
    int V = pick_v(cin);
    int hcap = (h + 31) / 32 * 32;
    size_t per_warp = (size_t)num_kp * hcap * 8 + KP_MAX * 4 + (size_t)num_kp * 32 * V * 4;
    if (per_warp + 512 > 200 * 1024) return MVK_ERR_RANGE;
    int wpb = (int)((96 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
    size_t smem = KP_MAX * 3 * 4 + per_warp * wpb;
    int blocks = (nq + wpb - 1) / wpb;
    int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_BWD(IDX, VV)                                                                          \
    do {                                                                                             \
        auto kern = kp_weighted_bwd<IDX, VV>;                                                        \
        MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        launch_pdl(kern, dim3(blocks), dim3(wpb * 32), smem, st, 1, a, grad_weighted, grad_x, hcap);                        \
    } while (0)
    if (idx_is_i64) {
        if (V == 4) LAUNCH_BWD(long long, 4);
        else if (V == 2) LAUNCH_BWD(long long, 2);
        else LAUNCH_BWD(long long, 1);
    } else {
        if (V == 4) LAUNCH_BWD(int, 4);
        else if (V == 2) LAUNCH_BWD(int, 2);
        else LAUNCH_BWD(int, 1);
    }
#undef LAUNCH_BWD
    MVK_LAUNCHED("kp_weighted_bwd");
    return MVK_OK;
}

int mvk_split_bf16(const float* src, int rows, int cols, int src_ld, void* hi, void* lo, int rows_pad,
                   int ld, mvk_stream_t stream) {
    if (!src || !hi || rows < 0 || cols < 0 || rows_pad < rows || ld < cols || src_ld < cols)
        return MVK_ERR_INVALID_ARG;
    size_t total = (size_t)rows_pad * ld;
    if (total == 0) return MVK_OK;
    if (lo && rows_pad == rows && ld == cols && cols % 4 == 0 && src_ld % 4 == 0 && (((size_t)src) & 15) == 0 &&
        (((size_t)hi | (size_t)lo) & 7) == 0) {
        const size_t nv = (size_t)rows * (cols / 4);
        size_t nb = (nv + 255) / 256, mb = (size_t)num_sms() * 16;
        launch_pdl(split_bf16_vec4, dim3((int)(nb < mb ? nb : mb)), dim3(256), 0, (cudaStream_t)stream, 1, 
            src, rows, cols / 4, src_ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
        MVK_LAUNCHED("split_bf16_vec4");
        return MVK_OK;
    }
    int blocks = (int)((total + 255) / 256);
    int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    launch_pdl(split_bf16_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, src, rows, cols, src_ld,
                                                                (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                rows_pad, ld);
    MVK_LAUNCHED("split_bf16");
    return MVK_OK;
}

int mvk_split_bf16_multi(const mvk_split_desc* table_dev, int n_tensors, int total_chunks, mvk_stream_t stream) {
    if (!table_dev || n_tensors < 1 || total_chunks < 1) return MVK_ERR_INVALID_ARG;
    launch_pdl(split_bf16_multi, dim3(total_chunks), dim3(256), 0, (cudaStream_t)stream, 1, table_dev, n_tensors);
    MVK_LAUNCHED("split_bf16_multi");
    return MVK_OK;
}

int mvk_upsample_concat_split(const float* x_coarse, int ns, int c1, const void* inds, int idx_is_i64, int nq, int h,
                              const float* skip, int lds, int c2, void* hi, void* lo, int ldh, mvk_stream_t stream) {
    if (!x_coarse || !inds || !skip || !hi || !lo || ns < 0 || nq < 0 || h < 1 || c1 < 4 || c2 < 4 || (c1 % 4) != 0 ||
        (c2 % 4) != 0 || lds < c2 || (lds % 4) != 0 || ldh < c1 + c2 || (ldh % 4) != 0 ||
        ((((size_t)x_coarse) | ((size_t)skip)) & 15) != 0 || ((((size_t)hi) | ((size_t)lo)) & 7) != 0)
        return MVK_ERR_INVALID_ARG;
    if (nq == 0) return MVK_OK;
    const size_t nv = (size_t)nq * ((c1 + c2) / 4);
    size_t nb = (nv + 255) / 256, mb = (size_t)num_sms() * 16;
    const dim3 grid((unsigned)(nb < mb ? nb : mb));
    if (idx_is_i64)
        launch_pdl(upcat_split_kernel<long long>, grid, dim3(256), 0, (cudaStream_t)stream, 1, x_coarse, ns, c1,
                   (const long long*)inds, h, skip, lds, c2, nq, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ldh);
    else
        launch_pdl(upcat_split_kernel<int>, grid, dim3(256), 0, (cudaStream_t)stream, 1, x_coarse, ns, c1,
                   (const int*)inds, h, skip, lds, c2, nq, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ldh);
    MVK_LAUNCHED("upcat_split_kernel");
    return MVK_OK;
}

int mvk_gather_add_rows(float* y, int ldy, int nq, int c, const float* z, int ns, const void* inds, int idx_is_i64, int h,
                        mvk_stream_t stream) {
    if (!y || !z || !inds || nq < 0 || ns < 0 || c < 4 || (c % 4) != 0 || ldy < c || (ldy % 4) != 0 || h < 1 ||
        ((((size_t)y) | ((size_t)z)) & 15) != 0)
        return MVK_ERR_INVALID_ARG;
    if (nq == 0) return MVK_OK;
    const size_t nv = (size_t)nq * (c / 4);
    size_t nb = (nv + 255) / 256, mb = (size_t)num_sms() * 16;
    const dim3 grid((unsigned)(nb < mb ? nb : mb));
    if (idx_is_i64)
        launch_pdl(gather_add_rows_kernel<long long>, grid, dim3(256), 0, (cudaStream_t)stream, 1, y, ldy, nq, c / 4, z, ns, inds, h);
    else
        launch_pdl(gather_add_rows_kernel<int>, grid, dim3(256), 0, (cudaStream_t)stream, 1, y, ldy, nq, c / 4, z, ns, inds, h);
    MVK_LAUNCHED("gather_add_rows_kernel");
    return MVK_OK;
}

int mvk_pool(const float* x, int ns, int c, const void* inds, int idx_is_i64, int nq, int h, int mode,
             float* out, int* arg_out, mvk_stream_t stream) {
    if (!x || !inds || !out || c < 1 || h < 1 || nq < 0 || mode < 0 || mode > 1) return MVK_ERR_INVALID_ARG;
    if (nq == 0) return MVK_OK;
    int blocks = (nq + 7) / 8;
    int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    if (c % 4 == 0 && (((size_t)x | (size_t)out | (size_t)arg_out) & 15) == 0) {
        const long long items = (long long)nq * ((c + 127) / 128);
        blocks = (int)((items + 7) / 8 < (long long)maxb ? (items + 7) / 8 : maxb);
        if (idx_is_i64)
            launch_pdl(pool_fwd_vec4<long long>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, x, ns, c, inds, nq, h, mode, out, arg_out);
        else
            launch_pdl(pool_fwd_vec4<int>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, x, ns, c, inds, nq, h, mode, out, arg_out);
    } else if (idx_is_i64)
        launch_pdl(pool_fwd<long long>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, x, ns, c, inds, nq, h, mode, out, arg_out);
    else
        launch_pdl(pool_fwd<int>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, x, ns, c, inds, nq, h, mode, out, arg_out);
    MVK_LAUNCHED("pool_fwd");
    return MVK_OK;
}

int mvk_pool_bwd(const float* grad_out, int ldg, int nq, int c, const int* arg, const void* inds,
                 int idx_is_i64, int h, int mode, int ns, float* grad_x, mvk_stream_t stream) {
    if (!grad_out || !grad_x || c < 1 || ldg < c || nq < 0 || (mode == 0 && !arg) || (mode == 1 && !inds))
        return MVK_ERR_INVALID_ARG;
    if (nq == 0) return MVK_OK;
    size_t total = (size_t)nq * c;
    int blocks = (int)((total + 255) / 256);
    int maxb = num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    if (mode == 1 && (c % 4) == 0 && (ldg % 4) == 0 && ((((size_t)grad_out) | ((size_t)grad_x)) & 15) == 0) {
        int vb = (int)((total / 4 + 255) / 256);
        if (vb > maxb) vb = maxb;
        if (idx_is_i64)
            launch_pdl(pool_bwd_closest_vec4<long long>, dim3(vb), dim3(256), 0, (cudaStream_t)stream, 1, grad_out, ldg, nq, c, inds, h, ns, grad_x);
        else
            launch_pdl(pool_bwd_closest_vec4<int>, dim3(vb), dim3(256), 0, (cudaStream_t)stream, 1, grad_out, ldg, nq, c, inds, h, ns, grad_x);
        MVK_LAUNCHED("pool_bwd_closest_vec4");
        return MVK_OK;
    }
    if (idx_is_i64)
        launch_pdl(pool_bwd<long long>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, grad_out, ldg, nq, c, arg, inds, h, mode, ns, grad_x);
    else
        launch_pdl(pool_bwd<int>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, 1, grad_out, ldg, nq, c, arg, inds, h, mode, ns, grad_x);
    MVK_LAUNCHED("pool_bwd");
    return MVK_OK;
}

}  // extern "C"
