// 2D -> 3D lifting: depth unprojection, point-to-pixel k-NN, group_points, FeatureAggregation.
// Reference: KPConv-PyTorch/datasets/ScanNet_sphere_color.py:66-72, 409-452;
//            mvpnet/ops/group_points.py:5-31 (+ ops/cuda/group_points_kernel.cu:25-144);
//            mvpnet/models/mvpnet_3d.py:40-64 (+ common/nn/modules/mlp.py:38-75).
#include <stdlib.h>

#include "common.cuh"

namespace mvk {
namespace {

// ------------------------------------------------------------------------------------------------
// depth2xyz + pose, fp64 like numpy (int64 pixel grid x fp32 K^-1 promotes to float64).
__global__ void __launch_bounds__(256)
unproject_kernel(const double* __restrict__ kinv, const float* __restrict__ depth,
                 const float* __restrict__ pose, int nv, int h, int w, double* __restrict__ xyz64,
                 float* __restrict__ xyz32, unsigned char* __restrict__ mask) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int hw = h * w;
    if (t >= nv * hw) return;
    int view = t / hw, pix = t % hw;
    double u = (double)(pix % w), v = (double)(pix / w);
    double d = (double)depth[t];
    // xyz_cam = (Kinv . [u, v, 1]) * depth          (ScanNet_sphere_color.py:66-72)
    double c[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double acc = __dmul_rn(kinv[3 * r], u);
        acc = __dadd_rn(acc, __dmul_rn(kinv[3 * r + 1], v));
        acc = __dadd_rn(acc, kinv[3 * r + 2]);
        c[r] = __dmul_rn(acc, d);
    }
    mask[t] = c[2] > 0.0 ? 1 : 0;  // :415
    // xyz_world = xyz_cam . R^T + t                 (:417)
    const float* P = pose + 16 * view;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double acc = __dmul_rn(c[0], (double)P[4 * r]);
        acc = __dadd_rn(acc, __dmul_rn(c[1], (double)P[4 * r + 1]));
        acc = __dadd_rn(acc, __dmul_rn(c[2], (double)P[4 * r + 2]));
        acc = __dadd_rn(acc, (double)P[4 * r + 3]);
        xyz64[3 * (size_t)t + r] = acc;
        xyz32[3 * (size_t)t + r] = (float)acc;
    }
}

// ------------------------------------------------------------------------------------------------
// k-NN of queries among the valid pixels, brute force in fp64 with keys tiled through shared memory.
constexpr int KNN_K = 8;
constexpr int KNN_TILE = 1024;
constexpr int KNN_THREADS = 256;

struct KnnWork {
    int* flag;      // [npix]
    int* pos;       // [npix]
    int* nvalid;    // [1]
    double* kx;     // [npix] compacted keys (SoA)
    double* ky;
    double* kz;
    int* kid;       // [npix] flat pixel id of the compacted key
    double* pd;     // [nq * splits * k] partial distances
    int* pid;       // [nq * splits * k]
    int* scan_tmp;
};

constexpr int KNN_MAX_SPLITS = 64;

KnnWork knn_carve(Arena& a, int npix, int nq) {
    KnnWork w;
    int n1 = npix > 0 ? npix : 1, q1 = nq > 0 ? nq : 1;
    w.flag = a.take<int>(n1);
    w.pos = a.take<int>(n1);
    w.nvalid = a.take<int>(1);
    w.kx = a.take<double>(n1);
    w.ky = a.take<double>(n1);
    w.kz = a.take<double>(n1);
    w.kid = a.take<int>(n1);
    w.pd = a.take<double>((size_t)q1 * KNN_MAX_SPLITS * KNN_K);
    w.pid = a.take<int>((size_t)q1 * KNN_MAX_SPLITS * KNN_K);
    w.scan_tmp = a.take<int>(scan_tmp_ints(n1));
    return w;
}

__global__ void __launch_bounds__(256) knn_flags(const unsigned char* __restrict__ mask, int npix, KnnWork w) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < npix) w.flag[t] = mask[t] ? 1 : 0;
}
__global__ void __launch_bounds__(256)
knn_compact(const double* __restrict__ xyz, int npix, KnnWork w) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < npix && w.flag[t]) {
        int p = w.pos[t];
        w.kx[p] = xyz[3 * (size_t)t];
        w.ky[p] = xyz[3 * (size_t)t + 1];
        w.kz[p] = xyz[3 * (size_t)t + 2];
        w.kid[p] = t;
    }
}

__device__ __forceinline__ bool knn_less(double d, int id, double d2, int id2) {
    return d < d2 || (d == d2 && id < id2);
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_partial(const float* __restrict__ q, int nq, int k, int splits, KnnWork w) {
    __shared__ double sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    __shared__ int sid[KNN_TILE];
    const int nkeys = *w.nvalid;
    const int per = (nkeys + splits - 1) / splits;
    const int k0 = blockIdx.y * per, k1 = min(nkeys, k0 + per);
    const int qi = blockIdx.x * KNN_THREADS + threadIdx.x;
    double qx = 0, qy = 0, qz = 0;
    if (qi < nq) {
        qx = (double)q[3 * qi];
        qy = (double)q[3 * qi + 1];
        qz = (double)q[3 * qi + 2];
    }
    double bd[KNN_K];
    int bi[KNN_K];
#pragma unroll
    for (int j = 0; j < KNN_K; j++) {
        bd[j] = 1.0e300;
        bi[j] = 0x7fffffff;
    }
    for (int base = k0; base < k1; base += KNN_TILE) {
        int n = min(KNN_TILE, k1 - base);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += KNN_THREADS) {
            sx[t] = w.kx[base + t];
            sy[t] = w.ky[base + t];
            sz[t] = w.kz[base + t];
            sid[t] = w.kid[base + t];
        }
        __syncthreads();
        if (qi < nq) {
            for (int t = 0; t < n; t++) {
                double dx = __dsub_rn(qx, sx[t]), dy = __dsub_rn(qy, sy[t]), dz = __dsub_rn(qz, sz[t]);
                double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                int id = sid[t];
                if (knn_less(d, id, bd[KNN_K - 1], bi[KNN_K - 1])) {
                    // insertion into the sorted top list (fully unrolled: stays in registers)
                    bd[KNN_K - 1] = d;
                    bi[KNN_K - 1] = id;
#pragma unroll
                    for (int j = KNN_K - 1; j > 0; j--) {
                        if (knn_less(bd[j], bi[j], bd[j - 1], bi[j - 1])) {
                            double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                            int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (qi < nq) {
        size_t o = ((size_t)qi * splits + blockIdx.y) * KNN_K;
#pragma unroll
        for (int j = 0; j < KNN_K; j++) {
            w.pd[o + j] = bd[j];
            w.pid[o + j] = bi[j];
        }
    }
    (void)k;
}

__global__ void __launch_bounds__(256)
knn_merge(int nq, int k, int splits, long long* __restrict__ out, KnnWork w) {
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    double bd[KNN_K];
    int bi[KNN_K];
#pragma unroll
    for (int j = 0; j < KNN_K; j++) {
        bd[j] = 1.0e300;
        bi[j] = 0x7fffffff;
    }
    for (int s = 0; s < splits; s++) {
        size_t o = ((size_t)qi * splits + s) * KNN_K;
        for (int e = 0; e < KNN_K; e++) {
            double d = w.pd[o + e];
            int id = w.pid[o + e];
            if (knn_less(d, id, bd[KNN_K - 1], bi[KNN_K - 1])) {
                bd[KNN_K - 1] = d;
                bi[KNN_K - 1] = id;
#pragma unroll
                for (int j = KNN_K - 1; j > 0; j--) {
                    if (knn_less(bd[j], bi[j], bd[j - 1], bi[j - 1])) {
                        double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                        int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KNN_K; j++)
        if (j < k) out[(size_t)qi * k + j] = (bi[j] == 0x7fffffff) ? -1ll : (long long)bi[j];
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
group_points_fwd(const float* __restrict__ pts, int b, int c, int n1, const long long* __restrict__ index,
                 int n2, int k, float* __restrict__ out) {
    size_t total = (size_t)b * c * n2 * k;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t nk = t % ((size_t)n2 * k);
        size_t bc = t / ((size_t)n2 * k);
        int bi = (int)(bc / c);
        long long idx = index[(size_t)bi * n2 * k + nk];
        out[t] = (idx >= 0 && idx < n1) ? pts[bc * n1 + idx] : 0.f;
    }
}
__global__ void __launch_bounds__(256)
group_points_bwd(const float* __restrict__ go, int b, int c, int n1, const long long* __restrict__ index,
                 int n2, int k, float* __restrict__ gp) {
    size_t total = (size_t)b * c * n2 * k;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t nk = t % ((size_t)n2 * k);
        size_t bc = t / ((size_t)n2 * k);
        int bi = (int)(bc / c);
        long long idx = index[(size_t)bi * n2 * k + nk];
        if (idx >= 0 && idx < n1) atomicAdd(&gp[bc * n1 + idx], go[t]);
    }
}

// ------------------------------------------------------------------------------------------------
// FeatureAggregation building blocks.
__global__ void __launch_bounds__(256)
fa_gather_kernel(const float* __restrict__ feat, long long chan_stride, long long pix_stride, int c,
                 const float* __restrict__ xyz32, const long long* __restrict__ knn, int np, int k,
                 const float* __restrict__ tgt, float* __restrict__ X, int ldx) {
    // one warp per (point, neighbour) row
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int rows = np * k;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        long long pix = knn[r];
        float* xr = X + (size_t)r * ldx;
        for (int ch = lane; ch < c; ch += 32) xr[ch] = feat[ch * chan_stride + pix * pix_stride];
        if (lane == 0) {
            int p = r / k;
            float dx = xyz32[3 * pix] - tgt[3 * p], dy = xyz32[3 * pix + 1] - tgt[3 * p + 1],
                  dz = xyz32[3 * pix + 2] - tgt[3 * p + 2];
            xr[c] = dx;
            xr[c + 1] = dy;
            xr[c + 2] = dz;
            xr[c + 3] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        }
    }
}

constexpr int FA_ROWS = 4;  // rows per warp iteration

// Y[r, o] = sum_c act(X[r, c]) W[o, c]; per-channel sum / sumsq of Y accumulated in fp64.
__global__ void __launch_bounds__(256)
fa_layer_kernel(const float* __restrict__ X, int rows, int cin, int ldx, const float* __restrict__ in_scale,
                const float* __restrict__ in_shift, const float* __restrict__ W, int cout,
                float* __restrict__ Y, double* __restrict__ stats) {
    extern __shared__ __align__(16) float fa_smem[];
    float* Wt = fa_smem;                       // [cin][cout]  (transposed: conflict-free lane reads)
    float* xs = fa_smem + (size_t)cin * cout;  // [warps][FA_ROWS][cin]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int t = threadIdx.x; t < cin * cout; t += blockDim.x) {
        int o = t / cin, cc = t % cin;
        Wt[cc * cout + o] = W[t];
    }
    __syncthreads();
    float* xw = xs + (size_t)wib * FA_ROWS * cin;
    const int och = (cout + 31) / 32;  // output channels per lane (<= 8 supported)
    double ssum[8], ssq[8];
#pragma unroll
    for (int j = 0; j < 8; j++) ssum[j] = ssq[j] = 0.0;

    for (int r0 = (blockIdx.x * wpb + wib) * FA_ROWS; r0 < rows; r0 += gridDim.x * wpb * FA_ROWS) {
        __syncwarp();
        for (int rr = 0; rr < FA_ROWS; rr++) {
            int r = r0 + rr;
            for (int cc = lane; cc < cin; cc += 32) {
                float v = 0.f;
                if (r < rows) {
                    v = X[(size_t)r * ldx + cc];
                    if (in_scale) v = fmaxf(0.f, fmaf(v, in_scale[cc], in_shift[cc]));
                }
                xw[rr * cin + cc] = v;
            }
        }
        __syncwarp();
        for (int j = 0; j < och; j++) {
            int o = lane + 32 * j;
            if (o >= cout) break;
            float acc[FA_ROWS];
#pragma unroll
            for (int rr = 0; rr < FA_ROWS; rr++) acc[rr] = 0.f;
            for (int cc = 0; cc < cin; cc++) {
                float wv = Wt[cc * cout + o];
#pragma unroll
                for (int rr = 0; rr < FA_ROWS; rr++) acc[rr] = fmaf(xw[rr * cin + cc], wv, acc[rr]);
            }
#pragma unroll
            for (int rr = 0; rr < FA_ROWS; rr++) {
                int r = r0 + rr;
                if (r < rows) {
                    Y[(size_t)r * cout + o] = acc[rr];
                    ssum[j] += (double)acc[rr];
                    ssq[j] += (double)acc[rr] * (double)acc[rr];
                }
            }
        }
    }
    if (stats) {
        for (int j = 0; j < och; j++) {
            int o = lane + 32 * j;
            if (o < cout) {
                atomicAdd(&stats[o], ssum[j]);
                atomicAdd(&stats[cout + o], ssq[j]);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
fa_reduce_kernel(const float* __restrict__ Y, int np, int k, int cout, const float* __restrict__ scale,
                 const float* __restrict__ shift, int reduction, float* __restrict__ out) {
    size_t total = (size_t)np * cout;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int p = (int)(t % np), o = (int)(t / np);
        float sc = scale[o], sh = shift[o];
        float acc = reduction == 0 ? 0.f : -3.4e38f;
        for (int kk = 0; kk < k; kk++) {
            float v = fmaxf(0.f, fmaf(Y[((size_t)p * k + kk) * cout + o], sc, sh));
            acc = reduction == 0 ? acc + v : fmaxf(acc, v);
        }
        out[t] = acc;
    }
}

int grid_for(size_t total, int threads) {
    size_t b = (total + threads - 1) / threads;
    size_t maxb = (size_t)num_sms() * 32;
    return (int)(b < 1 ? 1 : (b > maxb ? maxb : b));
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

int mvk_unproject_views(const double* kinv, const float* depth, const float* pose, int nv, int h, int w,
                        double* xyz64, float* xyz32, unsigned char* mask, mvk_stream_t stream) {
    if (!kinv || !depth || !pose || !xyz64 || !xyz32 || !mask || nv < 1 || h < 1 || w < 1) return MVK_ERR_INVALID_ARG;
    int total = nv * h * w;
    unproject_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(kinv, depth, pose, nv, h, w, xyz64,
                                                                           xyz32, mask);
    MVK_LAUNCHED("unproject_kernel");
    return MVK_OK;
}

size_t mvk_knn_workspace_bytes(int npix, int nq) {
    Arena a(nullptr, 0);
    knn_carve(a, npix, nq);
    return a.off + 256;
}

int mvk_knn_pixels(const double* xyz64, const unsigned char* mask, int npix, const float* queries, int nq, int k,
                   void* ws, size_t ws_bytes, long long* out, mvk_stream_t stream) {
    if (!xyz64 || !mask || !queries || !out || npix < 1 || nq < 0 || k < 1 || k > KNN_K) return MVK_ERR_INVALID_ARG;
    if (!ws || ws_bytes < mvk_knn_workspace_bytes(npix, nq)) return MVK_ERR_WORKSPACE;
    if (nq == 0) return MVK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    KnnWork w = knn_carve(a, npix, nq);
    knn_flags<<<(npix + 255) / 256, 256, 0, st>>>(mask, npix, w);
    MVK_LAUNCHED("knn_flags");
    int rc = exclusive_scan_i32(w.flag, w.pos, npix, w.nvalid, w.scan_tmp, st);
    if (rc) return rc;
    knn_compact<<<(npix + 255) / 256, 256, 0, st>>>(xyz64, npix, w);
    MVK_LAUNCHED("knn_compact");
    int qblocks = (nq + KNN_THREADS - 1) / KNN_THREADS;
    int splits = (2 * num_sms() + qblocks - 1) / qblocks;
    int max_splits_by_keys = (npix + KNN_TILE - 1) / KNN_TILE;
    if (splits > max_splits_by_keys) splits = max_splits_by_keys;
    if (splits > KNN_MAX_SPLITS) splits = KNN_MAX_SPLITS;
    if (splits < 1) splits = 1;
    knn_partial<<<dim3(qblocks, splits), KNN_THREADS, 0, st>>>(queries, nq, k, splits, w);
    MVK_LAUNCHED("knn_partial");
    knn_merge<<<(nq + 255) / 256, 256, 0, st>>>(nq, k, splits, out, w);
    MVK_LAUNCHED("knn_merge");
    return MVK_OK;
}

int mvk_group_points(const float* points, int b, int c, int n1, const long long* index, int n2, int k,
                     float* out, mvk_stream_t stream) {
    if (!points || !index || !out || b < 1 || c < 1 || n1 < 1 || n2 < 0 || k < 1) return MVK_ERR_INVALID_ARG;
    size_t total = (size_t)b * c * n2 * k;
    if (total == 0) return MVK_OK;
    group_points_fwd<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(points, b, c, n1, index, n2, k, out);
    MVK_LAUNCHED("group_points_fwd");
    return MVK_OK;
}

int mvk_group_points_bwd(const float* grad_out, int b, int c, int n1, const long long* index, int n2, int k,
                         float* grad_points, mvk_stream_t stream) {
    if (!grad_out || !index || !grad_points || b < 1 || c < 1 || n1 < 1 || n2 < 0 || k < 1) return MVK_ERR_INVALID_ARG;
    size_t total = (size_t)b * c * n2 * k;
    if (total == 0) return MVK_OK;
    group_points_bwd<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, b, c, n1, index, n2, k,
                                                                             grad_points);
    MVK_LAUNCHED("group_points_bwd");
    return MVK_OK;
}

int mvk_fa_gather(const float* feat2d, long long chan_stride, long long pix_stride, int c, const float* xyz32,
                  const long long* knn, int np, int k, const float* tgt_xyz, float* X, int ldx,
                  mvk_stream_t stream) {
    if (!feat2d || !xyz32 || !knn || !tgt_xyz || !X || c < 1 || np < 0 || k < 1 || ldx < c + 4) return MVK_ERR_INVALID_ARG;
    if (np == 0) return MVK_OK;
    int rows = np * k;
    fa_gather_kernel<<<grid_for((size_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        feat2d, chan_stride, pix_stride, c, xyz32, knn, np, k, tgt_xyz, X, ldx);
    MVK_LAUNCHED("fa_gather_kernel");
    return MVK_OK;
}

int mvk_fa_layer(const float* X, int rows, int cin, int ldx, const float* in_scale, const float* in_shift,
                 const float* W, int cout, float* Y, double* stats, mvk_stream_t stream) {
    if (!X || !W || !Y || rows < 0 || cin < 1 || cout < 1 || cout > 256 || ldx < cin) return MVK_ERR_INVALID_ARG;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return MVK_ERR_INVALID_ARG;
    if (rows == 0) return MVK_OK;
    size_t smem = ((size_t)cin * cout + 8 * FA_ROWS * cin) * sizeof(float);
    if (smem > 200 * 1024) return MVK_ERR_RANGE;
    MVK_CUDA(cudaFuncSetAttribute(fa_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (rows + 8 * FA_ROWS - 1) / (8 * FA_ROWS);
    int maxb = num_sms() * 4;
    if (blocks > maxb) blocks = maxb;
    fa_layer_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(X, rows, cin, ldx, in_scale, in_shift, W, cout, Y,
                                                                 stats);
    MVK_LAUNCHED("fa_layer_kernel");
    return MVK_OK;
}

int mvk_fa_reduce(const float* Y, int np, int k, int cout, const float* scale, const float* shift, int reduction,
                  float* out, mvk_stream_t stream) {
    if (!Y || !scale || !shift || !out || np < 0 || k < 1 || cout < 1 || reduction < 0 || reduction > 1)
        return MVK_ERR_INVALID_ARG;
    if (np == 0) return MVK_OK;
    fa_reduce_kernel<<<grid_for((size_t)np * cout, 256), 256, 0, (cudaStream_t)stream>>>(Y, np, k, cout, scale, shift,
                                                                                         reduction, out);
    MVK_LAUNCHED("fa_reduce_kernel");
    return MVK_OK;
}

}  // extern "C"
