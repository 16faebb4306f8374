// Point-wise blocks either side of KPConv: batch-norm statistics, folded scale/shift + LeakyReLU,
// their backward, and the residual add + LeakyReLU.  Reference: KPConv-PyTorch/models/blocks.py
// :430-466 (BatchNormBlock), :469-504 (UnaryBlock), :637-649 (ResnetBottleneckBlock tail).
// The Linear of the unary block runs on the tcgen05 contraction (gemm_tc.cu); everything here is
// HBM-bound streaming over [rows, cols] fp32 matrices:
//   thread (tx, ty): tx = column vector (VEC floats), ty = row inside the CTA's row slab;
//   column sums are reduced through shared memory, then across the CTAs of a thread-block cluster
//   through distributed shared memory, and leave as one fp64 atomic per column and CLUSTER
//   (ATOMG.F64 is slow on sm_100: ~75 k of them cost 8 us, see scripts/stats_bench.py).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace mvk {
namespace {

namespace cg = cooperative_groups;

constexpr int TB = 256;
constexpr int CLUSTER = 8;  // CTAs (row slabs) per cluster in the column reductions

struct Map2D {
    int cv;    // vector columns per row (cols / VEC)
    int cpb;   // vector columns handled per pass by one CTA row of threads
    int rpi;   // rows per CTA iteration
};

__host__ __device__ inline Map2D make_map2d(int cols, int vec) {
    Map2D m;
    m.cv = cols / vec;
    m.cpb = m.cv < 32 ? m.cv : 32;  // one CTA covers at most 32 vector columns; blockIdx.y picks the group
    m.rpi = TB / m.cpb;
    return m;
}

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    typedef float4 T;
};
template <>
struct VecT<1> {
    typedef float T;
};

template <int VEC>
__device__ __forceinline__ void loadv(const float* p, float (&v)[VEC]) {
    if (VEC == 4) {
        float4 t = *(const float4*)p;
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = *p;
    }
}
template <int VEC>
__device__ __forceinline__ void storev(float* p, const float (&v)[VEC]) {
    if (VEC == 4) *(float4*)p = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else *p = v[0];
}
template <int VEC>
__device__ __forceinline__ void store_hilo(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        __nv_bfloat16 h = __float2bfloat16_rn(v[e]);
        hi[e] = h;
        lo[e] = __float2bfloat16_rn(v[e] - __bfloat162float(h));
    }
}
template <>
__device__ __forceinline__ void store_hilo<4>(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[4]) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    __nv_bfloat162 l0 = __floats2bfloat162_rn(v[0] - f0.x, v[1] - f0.y);
    __nv_bfloat162 l1 = __floats2bfloat162_rn(v[2] - f1.x, v[3] - f1.y);
    uint2 ph, pl;
    ph.x = *(unsigned int*)&h0; ph.y = *(unsigned int*)&h1;
    pl.x = *(unsigned int*)&l0; pl.y = *(unsigned int*)&l1;
    *(uint2*)hi = ph;
    *(uint2*)lo = pl;
}

// Reduces NV per-thread partial column vectors over the CTA's `rpi` thread rows, then over the CTAs of
// the cluster (which all share blockIdx.y), and adds them to out[q * cols + column] (one fp64 atomic per
// column and cluster, issued by cluster rank 0).  red: shared [NV][TB * VEC]; csum: shared [NV * 32 * VEC].
template <int VEC, int NV>
__device__ __forceinline__ void reduce_columns(float (&part)[NV][VEC], float* red, float* csum, const Map2D& m, int tx,
                                               int ty, bool active, int c0, int cols, double* out) {
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; q++)
#pragma unroll
        for (int e = 0; e < VEC; e++) red[(q * TB + ty * m.cpb + tx) * VEC + e] = active ? part[q][e] : 0.f;
    __syncthreads();
    // value v = (q, tx, e): every thread sums one value over the rpi thread rows (conflict-free)
    const int per_q = m.cpb * VEC, nval = NV * per_q;
    const int v = threadIdx.x, q = v / per_q, rem = v % per_q;
    if (v < nval) {
        float s0 = 0.f;
        for (int r = 0; r < m.rpi; r++) s0 += red[q * TB * VEC + r * per_q + rem];
        csum[v] = s0;
    }
    cluster.sync();
    if (cluster.block_rank() == 0 && v < nval && c0 + rem / VEC < m.cv) {
        double tot = 0.0;
        const unsigned int nb = cluster.num_blocks();
        for (unsigned int b = 0; b < nb; b++) tot += (double)cluster.map_shared_rank(csum, b)[v];
        atomicAdd(&out[(size_t)q * cols + (size_t)c0 * VEC + rem], tot);
    }
    cluster.sync();  // csum stays alive until rank 0 has read it
}

struct BnFin {
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* running_mean;
    float* running_var;
    float* scale;
    float* shift;
    float* mean_out;
    float* invstd_out;
    unsigned int* ticket;  // zero-initialised by the caller; NULL = no fused finalize
    long long* num_batches_tracked;  // optional: += 1 (torch.nn.BatchNorm1d bookkeeping)
};

// Per-column batch-norm bookkeeping with the inputs already in registers (the caller issues the loads of
// several columns before the first use: the last CTA's chain of L2 round trips is the tail of the kernel).
struct BnColIn {
    double s0, s1;
    float rm, rv, g, b;
};
__device__ __forceinline__ BnColIn bn_load_column(const double* stats, int cols, int c, const BnFin& f) {
    BnColIn in;
    in.s0 = __ldcg(stats + c);
    in.s1 = __ldcg(stats + cols + c);
    in.rm = f.running_mean ? f.running_mean[c] : 0.f;
    in.rv = f.running_mean ? f.running_var[c] : 0.f;
    in.g = f.gamma ? f.gamma[c] : 1.f;
    in.b = f.beta ? f.beta[c] : 0.f;
    return in;
}
__device__ __forceinline__ void bn_finalize_column(const BnColIn& in, int rows, int c, const BnFin& f) {
    const double mu = in.s0 / rows;
    double var = in.s1 / rows - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mu;
    const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
    if (f.running_mean) {
        const double unbiased = rows > 1 ? var * ((double)rows / (double)(rows - 1)) : var;
        f.running_mean[c] = (1.f - f.momentum) * in.rm + f.momentum * mean;
        f.running_var[c] = (1.f - f.momentum) * in.rv + f.momentum * (float)unbiased;
    }
    f.scale[c] = in.g * invstd;
    f.shift[c] = in.b - mean * in.g * invstd;
    if (f.mean_out) f.mean_out[c] = mean;
    if (f.invstd_out) f.invstd_out[c] = invstd;
}

// stats[c] += sum_r y[r,c] ; stats[cols + c] += sum_r y[r,c]^2
template <int VEC>
__global__ void __launch_bounds__(TB)
col_stats_kernel(const float* __restrict__ y, int rows, int cols, int ld, double* __restrict__ stats, int rows_per_cta,
                 BnFin fin) {
    pdl_enter();
    __shared__ float red[2 * TB * VEC];
    __shared__ float csum[2 * 32 * VEC];
    __shared__ unsigned int s_last;
    const Map2D m = make_map2d(cols, VEC);
    const int tx = threadIdx.x % m.cpb, ty = threadIdx.x / m.cpb;
    const bool active = ty < m.rpi;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int c0 = blockIdx.y * m.cpb; c0 < m.cv; c0 += m.cpb * gridDim.y) {
        float part[2][VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) part[0][e] = part[1][e] = 0.f;
        const bool cok = active && (c0 + tx < m.cv);
        if (cok) {
            const float* yc = y + (size_t)(c0 + tx) * VEC;
            int r = r0 + ty;
            for (; r + 7 * m.rpi < r1; r += 8 * m.rpi) {  // eight independent loads in flight
                float v[8][VEC];
#pragma unroll
                for (int u = 0; u < 8; u++) loadv<VEC>(yc + (size_t)(r + u * m.rpi) * ld, v[u]);
#pragma unroll
                for (int u = 0; u < 8; u++)
#pragma unroll
                    for (int e = 0; e < VEC; e++) {
                        part[0][e] += v[u][e];
                        part[1][e] = fmaf(v[u][e], v[u][e], part[1][e]);
                    }
            }
            for (; r < r1; r += m.rpi) {
                float v[VEC];
                loadv<VEC>(yc + (size_t)r * ld, v);
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    part[0][e] += v[e];
                    part[1][e] = fmaf(v[e], v[e], part[1][e]);
                }
            }
        }
        reduce_columns<VEC, 2>(part, red, csum, m, tx, ty, cok, c0, cols, stats);
    }
    if (fin.ticket) {
        // the CTA that retires last turns the column sums into scale / shift / running statistics
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(fin.ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (threadIdx.x == 0 && fin.num_batches_tracked) *fin.num_batches_tracked += 1;
            for (int cb = threadIdx.x; cb < cols; cb += 4 * TB) {
                BnColIn in[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (cb + u * TB < cols) in[u] = bn_load_column(stats, cols, cb + u * TB, fin);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (cb + u * TB < cols) bn_finalize_column(in[u], rows, cb + u * TB, fin);
            }
        }
    }
}

// Batch-norm bookkeeping on [cols] vectors (one tiny CTA).
//   training: mean / biased var from stats -> scale = gamma * invstd, shift = beta - mean * scale,
//             running stats updated with the unbiased variance (torch.nn.BatchNorm1d semantics).
//   eval    : scale / shift from the running statistics.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int rows, int cols, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, long long* __restrict__ num_batches_tracked) {
    if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
        float mean, invstd;
        if (training) {
            const double mu = stats[c] / rows;
            double var = stats[cols + c] / rows - mu * mu;
            if (var < 0.0) var = 0.0;
            mean = (float)mu;
            invstd = (float)(1.0 / sqrt(var + (double)eps));
            if (running_mean) {
                const double unbiased = rows > 1 ? var * ((double)rows / (double)(rows - 1)) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        } else {
            mean = running_mean[c];
            invstd = 1.f / sqrtf(running_var[c] + eps);
        }
        const float g = gamma ? gamma[c] : 1.f;
        const float b = beta ? beta[c] : 0.f;
        scale[c] = g * invstd;
        shift[c] = b - mean * g * invstd;
        if (mean_out) mean_out[c] = mean;
        if (invstd_out) invstd_out[c] = invstd;
    }
}

// out = leaky(y * scale + shift [+ residual]) ; optional bf16 hi/lo copy for a following contraction.
// Thread (tx, ty) owns one column vector for the whole kernel (its scale / shift live in registers) and
// walks the CTA's row slab four rows at a time (up to eight independent 16-byte loads in flight).
template <int VEC>
__global__ void __launch_bounds__(TB)
scale_shift_act_kernel(const float* __restrict__ y, int rows, int cols, int ld, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ residual, int ldr, float slope,
                       float* __restrict__ out, int ldo, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                       int ldh, int rows_per_cta) {
    pdl_enter();
    const Map2D m = make_map2d(cols, VEC);
    const int tx = threadIdx.x % m.cpb, ty = threadIdx.x / m.cpb;
    const int cvi = blockIdx.y * m.cpb + tx;
    if (ty >= m.rpi || cvi >= m.cv) return;
    const int c = cvi * VEC;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        sc[e] = scale ? scale[c + e] : 1.f;
        sh[e] = shift ? shift[c + e] : 0.f;
    }
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int rb = r0 + ty; rb < r1; rb += 4 * m.rpi) {
        float v[4][VEC], rs[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int r = rb + u * m.rpi;
#pragma unroll
            for (int e = 0; e < VEC; e++) v[u][e] = rs[u][e] = 0.f;
            if (r < r1) {
                loadv<VEC>(y + (size_t)r * ld + c, v[u]);
                if (residual) loadv<VEC>(residual + (size_t)r * ldr + c, rs[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int r = rb + u * m.rpi;
            if (r >= r1) break;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const float t = fmaf(v[u][e], sc[e], sh[e]) + rs[u][e];
                v[u][e] = t > 0.f ? t : t * slope;
            }
            if (out) storev<VEC>(out + (size_t)r * ldo + c, v[u]);
            if (hi) store_hilo<VEC>(hi + (size_t)r * ldh + c, lo + (size_t)r * ldh + c, v[u]);
        }
    }
}

// d = dz * leaky'(y * scale + shift [+ residual]);  sums[c] += d ; sums[cols + c] += d * xhat
// with xhat = (y - mean) * invstd (only when mean != NULL).
template <int VEC>
__global__ void __launch_bounds__(TB)
act_bwd_reduce_kernel(const float* __restrict__ dz, int lddz, const float* __restrict__ y, int rows, int cols, int ld,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ residual, int ldr, const float* __restrict__ mean,
                      const float* __restrict__ invstd, float slope, double* __restrict__ sums, int rows_per_cta) {
    pdl_enter();
    __shared__ float red[2 * TB * VEC];
    __shared__ float csum[2 * 32 * VEC];
    const Map2D m = make_map2d(cols, VEC);
    const int tx = threadIdx.x % m.cpb, ty = threadIdx.x / m.cpb;
    const bool active = ty < m.rpi;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int c0 = blockIdx.y * m.cpb; c0 < m.cv; c0 += m.cpb * gridDim.y) {
        float part[2][VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) part[0][e] = part[1][e] = 0.f;
        const bool cok = active && (c0 + tx < m.cv);
        if (cok) {
            const int c = (c0 + tx) * VEC;
            float sc[VEC], sh[VEC], mu[VEC], is[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                sc[e] = scale ? scale[c + e] : 1.f;
                sh[e] = shift ? shift[c + e] : 0.f;
                mu[e] = mean ? mean[c + e] : 0.f;
                is[e] = invstd ? invstd[c + e] : 0.f;
            }
            for (int rb = r0 + ty; rb < r1; rb += 2 * m.rpi) {
                float v[2][VEC], g[2][VEC], rs[2][VEC];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int r = rb + u * m.rpi;
#pragma unroll
                    for (int e = 0; e < VEC; e++) v[u][e] = g[u][e] = rs[u][e] = 0.f;
                    if (r < r1) {
                        loadv<VEC>(y + (size_t)r * ld + c, v[u]);
                        loadv<VEC>(dz + (size_t)r * lddz + c, g[u]);
                        if (residual) loadv<VEC>(residual + (size_t)r * ldr + c, rs[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; u++)
#pragma unroll
                    for (int e = 0; e < VEC; e++) {
                        const float pre = fmaf(v[u][e], sc[e], sh[e]) + rs[u][e];
                        const float d = pre > 0.f ? g[u][e] : g[u][e] * slope;  // g = 0 for rows past the slab
                        part[0][e] += d;
                        part[1][e] = fmaf(d, (v[u][e] - mu[e]) * is[e], part[1][e]);
                    }
            }
        }
        reduce_columns<VEC, 2>(part, red, csum, m, tx, ty, cok, c0, cols, sums);
    }
}

// dy = scale * (d - sum_d / rows - xhat * sum_dxhat / rows)   (batch norm, training)
// dy = scale * d                                              (eval / no batch norm: scale may be NULL = 1)
// d_res = d (gradient of the residual input), optional.
// Same thread mapping as scale_shift_act_kernel: the six per-column constants stay in registers.
template <int VEC>
__global__ void __launch_bounds__(TB)
act_bwd_apply_kernel(const float* __restrict__ dz, int lddz, const float* __restrict__ y, int rows, int cols, int ld,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ residual, int ldr, const float* __restrict__ mean,
                     const float* __restrict__ invstd, float slope, const double* __restrict__ sums, int batch_stats,
                     float* __restrict__ dy, int lddy, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                     int ldh, float* __restrict__ dres, int lddres, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int rows_per_cta) {
    pdl_enter();
    if (blockIdx.x == 0 && blockIdx.y == 0 && sums) {
        for (int c = threadIdx.x; c < cols; c += TB) {
            if (dbeta) dbeta[c] = (float)sums[c];
            if (dgamma) dgamma[c] = (float)sums[cols + c];
        }
    }
    const Map2D m = make_map2d(cols, VEC);
    const int tx = threadIdx.x % m.cpb, ty = threadIdx.x / m.cpb;
    const int cvi = blockIdx.y * m.cpb + tx;
    if (ty >= m.rpi || cvi >= m.cv) return;
    const int c = cvi * VEC;
    const float inv_rows = 1.f / (float)rows;
    float sc[VEC], sh[VEC], mu[VEC], is[VEC], s0[VEC], s1[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        sc[e] = scale ? scale[c + e] : 1.f;
        sh[e] = shift ? shift[c + e] : 0.f;
        mu[e] = batch_stats ? mean[c + e] : 0.f;
        is[e] = batch_stats ? invstd[c + e] : 0.f;
        s0[e] = batch_stats ? (float)sums[c + e] * inv_rows : 0.f;
        s1[e] = batch_stats ? (float)sums[cols + c + e] * inv_rows : 0.f;
    }
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int rb = r0 + ty; rb < r1; rb += 2 * m.rpi) {
        float v[2][VEC], g[2][VEC], rs[2][VEC];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = rb + u * m.rpi;
#pragma unroll
            for (int e = 0; e < VEC; e++) v[u][e] = g[u][e] = rs[u][e] = 0.f;
            if (r < r1) {
                loadv<VEC>(y + (size_t)r * ld + c, v[u]);
                loadv<VEC>(dz + (size_t)r * lddz + c, g[u]);
                if (residual) loadv<VEC>(residual + (size_t)r * ldr + c, rs[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = rb + u * m.rpi;
            if (r >= r1) break;
            float o[VEC], d[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const float pre = fmaf(v[u][e], sc[e], sh[e]) + rs[u][e];
                d[e] = pre > 0.f ? g[u][e] : g[u][e] * slope;
                // batch_stats == 0: s0 = s1 = 0, so this is sc * d
                o[e] = sc[e] * (d[e] - s0[e] - (v[u][e] - mu[e]) * is[e] * s1[e]);
            }
            if (dy) storev<VEC>(dy + (size_t)r * lddy + c, o);
            if (hi) store_hilo<VEC>(hi + (size_t)r * ldh + c, lo + (size_t)r * ldh + c, o);
            if (dres) storev<VEC>(dres + (size_t)r * lddres + c, d);
        }
    }
}

// dgamma[c] = sums[cols + c] ; dbeta[c] = sums[c]   (fp64 -> fp32)
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int cols, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
        if (dbeta) dbeta[c] = (float)sums[c];
        if (dgamma) dgamma[c] = (float)sums[cols + c];
    }
}

inline int vec_for(int cols, int a, int b, int c, int d) {
    return (cols % 4 == 0 && a % 4 == 0 && b % 4 == 0 && c % 4 == 0 && d % 4 == 0) ? 4 : 1;
}
inline int slab_rows(int rows, int cols, int vec, dim3* grid, int* cluster) {
    Map2D m = make_map2d(cols, vec);
    const int ncg = (m.cv + m.cpb - 1) / m.cpb;  // column groups (blockIdx.y)
    int target = num_sms() * 2 / ncg;            // ~2 fat CTAs per SM
    if (target < 1) target = 1;
    int rpc = (rows + target - 1) / target;
    rpc = (rpc + m.rpi - 1) / m.rpi * m.rpi;
    if (rpc < m.rpi * 4) rpc = m.rpi * 4;
    int gx = (rows + rpc - 1) / rpc;
    *cluster = gx >= CLUSTER ? CLUSTER : 1;
    gx = (gx + *cluster - 1) / *cluster * *cluster;  // CTAs past the last slab see an empty row range
    *grid = dim3(gx, ncg, 1);
    return rpc;
}
// Row slabs for the streaming kernels with the (tx, ty) mapping: ~8 CTAs per SM, slabs a multiple of the
// rows one CTA covers per trip.
inline int stream_rows(int rows, int cols, int vec, int unroll, dim3* grid) {
    Map2D m = make_map2d(cols, vec);
    const int ncg = (m.cv + m.cpb - 1) / m.cpb;
    int target = num_sms() * 8 / ncg;
    if (target < 1) target = 1;
    const int trip = m.rpi * unroll;
    int rpc = (rows + target - 1) / target;
    rpc = (rpc + trip - 1) / trip * trip;
    if (rpc < trip) rpc = trip;
    *grid = dim3((rows + rpc - 1) / rpc, ncg, 1);
    return rpc;
}
inline int ew_grid(size_t total) {
    size_t b = (total + TB - 1) / TB;
    size_t maxb = (size_t)num_sms() * 16;
    return (int)(b < maxb ? b : maxb);
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

int mvk_col_stats(const float* y, int rows, int cols, int ld, double* stats, mvk_stream_t stream) {
    if (!y || !stats || rows < 0 || cols < 1 || ld < cols) return MVK_ERR_INVALID_ARG;
    if (rows == 0) return MVK_OK;
    const int vec = vec_for(cols, ld, 4, 4, 4);
    dim3 grid;
    int cl;
    const int rpc = slab_rows(rows, cols, vec, &grid, &cl);
    BnFin fin = {};
    if (vec == 4) launch_pdl(col_stats_kernel<4>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, y, rows, cols, ld, stats, rpc, fin);
    else launch_pdl(col_stats_kernel<1>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, y, rows, cols, ld, stats, rpc, fin);
    MVK_LAUNCHED("col_stats");
    return MVK_OK;
}

int mvk_bn_batch_stats(const float* y, int rows, int cols, int ld, double* stats, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* mean_out, float* invstd_out, long long* num_batches_tracked,
                       mvk_stream_t stream) {
    if (!y || !stats || rows < 1 || cols < 1 || ld < cols || !scale || !shift) return MVK_ERR_INVALID_ARG;
    const int vec = vec_for(cols, ld, 4, 4, 4);
    dim3 grid;
    int cl;
    const int rpc = slab_rows(rows, cols, vec, &grid, &cl);
    BnFin fin = {gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean_out, invstd_out,
                 (unsigned int*)(stats + 2 * (size_t)cols), num_batches_tracked};
    if (vec == 4) launch_pdl(col_stats_kernel<4>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, y, rows, cols, ld, stats, rpc, fin);
    else launch_pdl(col_stats_kernel<1>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, y, rows, cols, ld, stats, rpc, fin);
    MVK_LAUNCHED("col_stats+finalize");
    return MVK_OK;
}

int mvk_bn_finalize(const double* stats, int rows, int cols, const float* gamma, const float* beta, float eps,
                    float momentum, int training, float* running_mean, float* running_var, float* scale,
                    float* shift, float* mean_out, float* invstd_out, long long* num_batches_tracked,
                    mvk_stream_t stream) {
    if (cols < 1 || !scale || !shift || (training && !stats) || (!training && (!running_mean || !running_var)))
        return MVK_ERR_INVALID_ARG;
    bn_finalize_kernel<<<(cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        stats, rows, cols, gamma, beta, eps, momentum, training, running_mean, running_var, scale, shift, mean_out,
        invstd_out, training ? num_batches_tracked : nullptr);
    MVK_LAUNCHED("bn_finalize");
    return MVK_OK;
}

int mvk_scale_shift_act(const float* y, int rows, int cols, int ld, const float* scale, const float* shift,
                        const float* residual, int ldr, float slope, float* out, int ldo, void* out_hi,
                        void* out_lo, int ldh, mvk_stream_t stream) {
    if (!y || rows < 0 || cols < 1 || ld < cols || (!out && !out_hi) || (out_hi && !out_lo) || (scale && !shift))
        return MVK_ERR_INVALID_ARG;
    if (rows == 0) return MVK_OK;
    const int vec = vec_for(cols, ld, residual ? ldr : 4, out ? ldo : 4, out_hi ? ldh : 4);
    dim3 grid;
    const int rpc = stream_rows(rows, cols, vec, 4, &grid);
    if (vec == 4)
        launch_pdl(scale_shift_act_kernel<4>, grid, dim3(TB), 0, (cudaStream_t)stream, 1, y, rows, cols, ld, scale, shift,
                   residual, ldr, slope, out, ldo, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ldh, rpc);
    else
        launch_pdl(scale_shift_act_kernel<1>, grid, dim3(TB), 0, (cudaStream_t)stream, 1, y, rows, cols, ld, scale, shift,
                   residual, ldr, slope, out, ldo, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ldh, rpc);
    MVK_LAUNCHED("scale_shift_act");
    return MVK_OK;
}

int mvk_act_bwd_reduce(const float* dz, int lddz, const float* y, int rows, int cols, int ld, const float* scale,
                       const float* shift, const float* residual, int ldr, const float* mean, const float* invstd,
                       float slope, double* sums, mvk_stream_t stream) {
    if (!dz || !y || !sums || rows < 0 || cols < 1 || ld < cols || lddz < cols || (scale && !shift))
        return MVK_ERR_INVALID_ARG;
    if (rows == 0) return MVK_OK;
    const int vec = vec_for(cols, ld, lddz, residual ? ldr : 4, 4);
    dim3 grid;
    int cl;
    const int rpc = slab_rows(rows, cols, vec, &grid, &cl);
    if (vec == 4)
        launch_pdl(act_bwd_reduce_kernel<4>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, dz, lddz, y, rows, cols, ld, scale,
                         shift, residual, ldr, mean, invstd, slope, sums, rpc);
    else
        launch_pdl(act_bwd_reduce_kernel<1>, grid, dim3(TB), 0, (cudaStream_t)stream, cl, dz, lddz, y, rows, cols, ld, scale,
                         shift, residual, ldr, mean, invstd, slope, sums, rpc);
    MVK_LAUNCHED("act_bwd_reduce");
    return MVK_OK;
}

int mvk_act_bwd_apply(const float* dz, int lddz, const float* y, int rows, int cols, int ld, const float* scale,
                      const float* shift, const float* residual, int ldr, const float* mean, const float* invstd,
                      float slope, const double* sums, int batch_stats, float* dy, int lddy, void* dy_hi,
                      void* dy_lo, int ldh, float* dres, int lddres, float* dgamma, float* dbeta,
                      mvk_stream_t stream) {
    if (!dz || !y || rows < 0 || cols < 1 || ld < cols || lddz < cols || (scale && !shift) ||
        (batch_stats && (!sums || !mean || !invstd || !scale)) || (dy_hi && !dy_lo))
        return MVK_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows > 0 && (dy || dy_hi || dres)) {
        int vec = vec_for(cols, ld, lddz, residual ? ldr : 4, dy ? lddy : 4);
        if (vec == 4 && ((dy_hi && ldh % 4 != 0) || (dres && lddres % 4 != 0))) vec = 1;
        dim3 grid;
        const int rpc = stream_rows(rows, cols, vec, 2, &grid);
        if (vec == 4)
            launch_pdl(act_bwd_apply_kernel<4>, grid, dim3(TB), 0, st, 1, dz, lddz, y, rows, cols, ld, scale, shift, residual,
                       ldr, mean, invstd, slope, sums, batch_stats, dy, lddy, (__nv_bfloat16*)dy_hi,
                       (__nv_bfloat16*)dy_lo, ldh, dres, lddres, dgamma, dbeta, rpc);
        else
            launch_pdl(act_bwd_apply_kernel<1>, grid, dim3(TB), 0, st, 1, dz, lddz, y, rows, cols, ld, scale, shift, residual,
                       ldr, mean, invstd, slope, sums, batch_stats, dy, lddy, (__nv_bfloat16*)dy_hi,
                       (__nv_bfloat16*)dy_lo, ldh, dres, lddres, dgamma, dbeta, rpc);
        MVK_LAUNCHED("act_bwd_apply");
    } else if ((dgamma || dbeta) && sums) {
        bn_param_grads_kernel<<<(cols + 255) / 256, 256, 0, st>>>(sums, cols, dgamma, dbeta);
        MVK_LAUNCHED("bn_param_grads");
    }
    return MVK_OK;
}

}  // extern "C"
