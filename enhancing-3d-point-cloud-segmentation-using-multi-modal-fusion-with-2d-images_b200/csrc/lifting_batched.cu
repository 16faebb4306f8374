// Batched 2D -> 3D lifting: every sphere of a stacked batch in ONE launch set (no per-sphere loop).
//
// Reference: the fusion nets loop over the batch elements in Python (KPConv-PyTorch/models/
// architectures_sphere.py:246-279: one group_points pair per sphere) on indices that the dataset computed per
// sphere on the CPU (datasets/ScanNet_sphere_color.py:409-452: numpy unprojection of every view, sklearn
// ball-tree 3-NN of the sphere points among the valid pixels, fp64).
//
//   mvk_unproject_views_batched   all nb*nv views, one intrinsics matrix per view
//   mvk_knn_pixels_batched        k nearest VALID pixels of every sphere point among the pixels of ITS OWN
//                                 sphere's views: per-sphere uniform grid over the unprojected pixels (counting
//                                 sort by cell), ring search with an exact stopping bound, warp-cooperative
//                                 exhaustive search for the queries the rings do not resolve.  Distances in fp64
//                                 with the arithmetic of the single-sphere kernel (lifting.cu), ties by lower
//                                 pixel id: bit-identical results.
//   mvk_fa_gather_views           first-layer rows of FeatureAggregation straight from the 2D network's output
//                                 tensor [views, C, h, w] in any memory format (NCHW or channels-last strides)
#include <float.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr int KB_RMAX = 4;        // rings searched on the grid before a query goes to the exhaustive pass
constexpr int KB_SLOTS = 8;       // k <= 8 like the single-sphere kernel

__global__ void __launch_bounds__(256)
unproject_batched_kernel(const double* __restrict__ kinv, const float* __restrict__ depth, const float* __restrict__ pose,
                         int nv, int h, int w, double* __restrict__ xyz64, float* __restrict__ xyz32,
                         unsigned char* __restrict__ mask) {
    const int hw = h * w;
    const size_t total = (size_t)nv * hw;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int view = (int)(t / hw), pix = (int)(t % hw);
    const double u = (double)(pix % w), v = (double)(pix / w);
    const double d = (double)depth[t];
    const double* Ki = kinv + 9 * (size_t)view;
    // xyz_cam = (Kinv . [u, v, 1]) * depth          (ScanNet_sphere_color.py:66-72; same op order as lifting.cu)
    double c[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double acc = __dmul_rn(Ki[3 * r], u);
        acc = __dadd_rn(acc, __dmul_rn(Ki[3 * r + 1], v));
        acc = __dadd_rn(acc, Ki[3 * r + 2]);
        c[r] = __dmul_rn(acc, d);
    }
    mask[t] = c[2] > 0.0 ? 1 : 0;
    const float* P = pose + 16 * (size_t)view;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double acc = __dmul_rn(c[0], (double)P[4 * r]);
        acc = __dadd_rn(acc, __dmul_rn(c[1], (double)P[4 * r + 1]));
        acc = __dadd_rn(acc, __dmul_rn(c[2], (double)P[4 * r + 2]));
        acc = __dadd_rn(acc, (double)P[4 * r + 3]);
        xyz64[3 * t + r] = acc;
        xyz32[3 * t + r] = (float)acc;
    }
}

// ------------------------------------------------------------------------------------------------
struct KbWork {
    int* qstart;      // [nb + 1] exclusive prefix of the query lengths
    int* bbox;        // [nb][6] ordered-int encoded min xyz / max xyz of the valid pixels
    float* params;    // [nb][8] origin xyz, cell, inv_cell, (pad)
    int* count;       // [nb * D^3]
    int* start;       // [nb * D^3 + 1]
    int* cursor;      // [nb * D^3]
    int* keycell;     // [nb * npix] cell id of every pixel (-1 = invalid)
    double* kx;       // [nb * npix] keys sorted by cell
    double* ky;
    double* kz;
    int* kid;         // [nb * npix] local flat pixel id of the sorted key
    int* far_idx;     // [nq] unresolved queries, segment of element b starts at qstart[b]
    int* far_count;   // [nb]
    int* scan_tmp;
};

KbWork kb_carve(Arena& a, int nb, int npix, int nq, int D) {
    KbWork w;
    const size_t cells = (size_t)nb * D * D * D;
    const size_t keys = (size_t)nb * npix;
    w.qstart = a.take<int>(nb + 1);
    w.bbox = a.take<int>((size_t)nb * 6);
    w.params = a.take<float>((size_t)nb * 8);
    w.count = a.take<int>(cells);
    w.start = a.take<int>(cells + 1);
    w.cursor = a.take<int>(cells);
    w.keycell = a.take<int>(keys);
    w.kx = a.take<double>(keys);
    w.ky = a.take<double>(keys);
    w.kz = a.take<double>(keys);
    w.kid = a.take<int>(keys);
    w.far_idx = a.take<int>(nq > 0 ? nq : 1);
    w.far_count = a.take<int>(nb);
    w.scan_tmp = a.take<int>(scan_tmp_ints((int)cells + 1));
    return w;
}

// monotone float <-> int mapping for atomicMin / atomicMax
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void kb_init(KbWork w, const int* __restrict__ qlen, int nb) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int s = 0;
        for (int b = 0; b < nb; b++) {
            w.qstart[b] = s;
            s += qlen[b];
        }
        w.qstart[nb] = s;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nb * 6; t += gridDim.x * blockDim.x)
        w.bbox[t] = (t % 6) < 3 ? 0x7fffffff : (int)0x80000000;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nb; t += gridDim.x * blockDim.x) w.far_count[t] = 0;
}

__global__ void __launch_bounds__(256)
kb_bbox(const float* __restrict__ xyz32, const unsigned char* __restrict__ mask, int nb, int npix, KbWork w) {
    // a warp never straddles two elements when npix % 32 == 0; otherwise lanes vote per element below
    const size_t total = (size_t)nb * npix;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = t < total && mask[t];
    const int b = t < total ? (int)(t / npix) : -1;
    float x = 0, y = 0, z = 0;
    if (ok) {
        x = xyz32[3 * t];
        y = xyz32[3 * t + 1];
        z = xyz32[3 * t + 2];
    }
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0);
    if (uniform) {
        float mnx = ok ? x : FLT_MAX, mny = ok ? y : FLT_MAX, mnz = ok ? z : FLT_MAX;
        float mxx = ok ? x : -FLT_MAX, mxy = ok ? y : -FLT_MAX, mxz = ok ? z : -FLT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
        }
        if ((threadIdx.x & 31) == 0 && b0 >= 0 && mnx <= mxx) {
            int* bb = w.bbox + 6 * b0;
            atomicMin(bb + 0, f2ord(mnx)); atomicMin(bb + 1, f2ord(mny)); atomicMin(bb + 2, f2ord(mnz));
            atomicMax(bb + 3, f2ord(mxx)); atomicMax(bb + 4, f2ord(mxy)); atomicMax(bb + 5, f2ord(mxz));
        }
    } else if (ok) {
        int* bb = w.bbox + 6 * b;
        atomicMin(bb + 0, f2ord(x)); atomicMin(bb + 1, f2ord(y)); atomicMin(bb + 2, f2ord(z));
        atomicMax(bb + 3, f2ord(x)); atomicMax(bb + 4, f2ord(y)); atomicMax(bb + 5, f2ord(z));
    }
}

__global__ void kb_params(KbWork w, int nb, int D, float cell_min) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int* bb = w.bbox + 6 * b;
    float* p = w.params + 8 * b;
    if (bb[0] == 0x7fffffff) {  // no valid pixel in this element
        p[0] = p[1] = p[2] = 0.f;
        p[3] = 1.f;
        p[4] = 1.f;
        return;
    }
    const float mnx = ord2f(bb[0]), mny = ord2f(bb[1]), mnz = ord2f(bb[2]);
    const float ext = fmaxf(fmaxf(ord2f(bb[3]) - mnx, ord2f(bb[4]) - mny), ord2f(bb[5]) - mnz);
    const float cell = fmaxf(cell_min, ext / (float)(D - 2));
    p[0] = mnx - 0.5f * cell;
    p[1] = mny - 0.5f * cell;
    p[2] = mnz - 0.5f * cell;
    p[3] = cell;
    p[4] = 1.f / cell;
}

__device__ __forceinline__ int cell_of(float x, float o, float inv, int D) {
    const int c = (int)floorf((x - o) * inv);
    return c < 0 ? 0 : (c > D - 1 ? D - 1 : c);
}

__global__ void __launch_bounds__(256)
kb_count(const float* __restrict__ xyz32, const unsigned char* __restrict__ mask, int nb, int npix, int D, KbWork w) {
    const size_t total = (size_t)nb * npix;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int cid = -1;
    if (mask[t]) {
        const int b = (int)(t / npix);
        const float* p = w.params + 8 * b;
        const int cx = cell_of(xyz32[3 * t], p[0], p[4], D), cy = cell_of(xyz32[3 * t + 1], p[1], p[4], D),
                  cz = cell_of(xyz32[3 * t + 2], p[2], p[4], D);
        cid = ((b * D + cz) * D + cy) * D + cx;
        atomicAdd(w.count + cid, 1);
    }
    w.keycell[t] = cid;
}

__global__ void __launch_bounds__(256)
kb_scatter(const double* __restrict__ xyz64, int nb, int npix, KbWork w) {
    const size_t total = (size_t)nb * npix;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cid = w.keycell[t];
    if (cid < 0) return;
    const int pos = w.start[cid] + atomicAdd(w.cursor + cid, 1);
    w.kx[pos] = xyz64[3 * t];
    w.ky[pos] = xyz64[3 * t + 1];
    w.kz[pos] = xyz64[3 * t + 2];
    w.kid[pos] = (int)(t % npix);
}

__device__ __forceinline__ bool kb_less(double d, int id, double d2, int id2) { return d < d2 || (d == d2 && id < id2); }

template <int S>
__device__ __forceinline__ void kb_insert(double (&bd)[S], int (&bi)[S], double d, int id) {
    if (kb_less(d, id, bd[S - 1], bi[S - 1])) {
        bd[S - 1] = d;
        bi[S - 1] = id;
#pragma unroll
        for (int j = S - 1; j > 0; j--) {
            if (kb_less(bd[j], bi[j], bd[j - 1], bi[j - 1])) {
                double td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
            }
        }
    }
}

__device__ __forceinline__ double kb_dist(double qx, double qy, double qz, double x, double y, double z) {
    const double dx = __dsub_rn(qx, x), dy = __dsub_rn(qy, y), dz = __dsub_rn(qz, z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// One thread per query: rings of cells around the query's cell until the k-th best distance is provably final.
template <int S>
__global__ void __launch_bounds__(128)
kb_query(const float* __restrict__ q, int nq, int nb, int npix, int D, int k, int global_ids, long long* __restrict__ out,
         KbWork w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int b = batch_of(w.qstart, nb, i);
    const float* p = w.params + 8 * b;
    const float ox = p[0], oy = p[1], oz = p[2], cell = p[3], inv = p[4];
    const float qxf = q[3 * i], qyf = q[3 * i + 1], qzf = q[3 * i + 2];
    const double qx = (double)qxf, qy = (double)qyf, qz = (double)qzf;
    const int hx = cell_of(qxf, ox, inv, D), hy = cell_of(qyf, oy, inv, D), hz = cell_of(qzf, oz, inv, D);
    double bd[S];
    int bi[S];
#pragma unroll
    for (int j = 0; j < S; j++) {
        bd[j] = 1.0e300;
        bi[j] = 0x7fffffff;
    }
    const int* start = w.start + (size_t)b * D * D * D;
    bool done = false;
    for (int r = 0; r <= KB_RMAX && !done; r++) {
        const int z0 = max(hz - r, 0), z1 = min(hz + r, D - 1);
        const int y0 = max(hy - r, 0), y1 = min(hy + r, D - 1);
        const int x0 = max(hx - r, 0), x1 = min(hx + r, D - 1);
        for (int cz = z0; cz <= z1; cz++) {
            const bool zface = (cz == hz - r) || (cz == hz + r);
            for (int cy = y0; cy <= y1; cy++) {
                const bool yface = zface || (cy == hy - r) || (cy == hy + r);
                // cells of the shell only: the whole x-run on a z / y face, else just the two x ends
                const int* row = start + ((size_t)cz * D + cy) * D;
                if (yface) {
                    const int s0 = row[x0], s1 = row[x1 + 1];  // contiguous run of cells = contiguous keys
                    for (int t = s0; t < s1; t++) kb_insert<S>(bd, bi, kb_dist(qx, qy, qz, w.kx[t], w.ky[t], w.kz[t]), w.kid[t]);
                } else {
                    if (hx - r >= 0) {
                        const int s0 = row[hx - r], s1 = row[hx - r + 1];
                        for (int t = s0; t < s1; t++) kb_insert<S>(bd, bi, kb_dist(qx, qy, qz, w.kx[t], w.ky[t], w.kz[t]), w.kid[t]);
                    }
                    if (hx + r <= D - 1 && r > 0) {
                        const int s0 = row[hx + r], s1 = row[hx + r + 1];
                        for (int t = s0; t < s1; t++) kb_insert<S>(bd, bi, kb_dist(qx, qy, qz, w.kx[t], w.ky[t], w.kz[t]), w.kid[t]);
                    }
                }
            }
        }
        // every key not visited yet lies outside the box of cells [h - r, h + r]: its distance is at least the
        // distance from the query to the nearest face of that box that still has cells behind it
        float bound = FLT_MAX;
        bool open = false;
        if (hx - r > 0) { bound = fminf(bound, qxf - (ox + (float)(hx - r) * cell)); open = true; }
        if (hx + r < D - 1) { bound = fminf(bound, (ox + (float)(hx + r + 1) * cell) - qxf); open = true; }
        if (hy - r > 0) { bound = fminf(bound, qyf - (oy + (float)(hy - r) * cell)); open = true; }
        if (hy + r < D - 1) { bound = fminf(bound, (oy + (float)(hy + r + 1) * cell) - qyf); open = true; }
        if (hz - r > 0) { bound = fminf(bound, qzf - (oz + (float)(hz - r) * cell)); open = true; }
        if (hz + r < D - 1) { bound = fminf(bound, (oz + (float)(hz + r + 1) * cell) - qzf); open = true; }
        if (!open) {
            done = true;  // the box covers the whole grid: every key was visited
        } else {
            // fp32 cell arithmetic vs fp64 key positions: keep a margin far above their rounding (1e-3 cell)
            const double bsafe = (double)bound - 1.0e-3 * (double)cell;
            if (bsafe > 0.0 && bd[k - 1] < bsafe * bsafe) done = true;
        }
    }
    if (!done) {
        const int pos = atomicAdd(w.far_count + b, 1);
        w.far_idx[w.qstart[b] + pos] = i;
        return;
    }
    const long long off = global_ids ? (long long)b * npix : 0ll;
#pragma unroll
    for (int j = 0; j < S; j++)
        if (j < k) out[(size_t)i * k + j] = (bi[j] == 0x7fffffff) ? -1ll : off + (long long)bi[j];
}

// Exhaustive pass for the unresolved queries: one warp per query, lanes stride over the element's keys, the 32
// per-lane top lists are merged by k rounds of a warp-wide lexicographic arg-min.  (Measured alternative: growing
// the ring search warp-cooperatively up to 16 rings before scanning -- 19 ms instead of 4 ms on the bench batch:
// the shell walk is a chain of dependent, scattered cell reads, the scan streams the sorted keys.)
template <int S>
__global__ void __launch_bounds__(256)
kb_far(const float* __restrict__ q, int nq, int nb, int npix, int D, int k, int global_ids, long long* __restrict__ out,
       KbWork w) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int pos = warp; pos < nq; pos += nwarps) {
        const int b = batch_of(w.qstart, nb, pos);
        if (pos - w.qstart[b] >= w.far_count[b]) continue;  // warp-uniform
        const int i = w.far_idx[pos];
        const double qx = (double)q[3 * i], qy = (double)q[3 * i + 1], qz = (double)q[3 * i + 2];
        const size_t cells = (size_t)D * D * D;
        const int k0 = w.start[(size_t)b * cells], k1 = w.start[(size_t)(b + 1) * cells];
        double bd[S];
        int bi[S];
#pragma unroll
        for (int j = 0; j < S; j++) {
            bd[j] = 1.0e300;
            bi[j] = 0x7fffffff;
        }
        for (int t = k0 + lane; t < k1; t += 32) kb_insert<S>(bd, bi, kb_dist(qx, qy, qz, w.kx[t], w.ky[t], w.kz[t]), w.kid[t]);
        const long long off = global_ids ? (long long)b * npix : 0ll;
        for (int j = 0; j < k; j++) {
            // warp arg-min of the lanes' current heads (bd[0], bi[0])
            double d = bd[0];
            int id = bi[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double d2 = __shfl_xor_sync(0xffffffffu, d, o);
                const int id2 = __shfl_xor_sync(0xffffffffu, id, o);
                if (kb_less(d2, id2, d, id)) {
                    d = d2;
                    id = id2;
                }
            }
            if (lane == 0) out[(size_t)i * k + j] = (id == 0x7fffffff) ? -1ll : off + (long long)id;
            if (bd[0] == d && bi[0] == id && id != 0x7fffffff) {  // the winning lane pops its head (ids are unique)
#pragma unroll
                for (int s = 0; s < S - 1; s++) {
                    bd[s] = bd[s + 1];
                    bi[s] = bi[s + 1];
                }
                bd[S - 1] = 1.0e300;
                bi[S - 1] = 0x7fffffff;
            }
        }
    }
}

// First-layer rows of FeatureAggregation, [np*k, c + 4] = [feature, diff xyz, |diff|^2] (mvpnet_3d.py:53-57),
// gathered from the 2D network's output [views, c, h*w] through global pixel ids g = view * hw + pix.
__global__ void __launch_bounds__(256)
fa_gather_views_kernel(const float* __restrict__ feat, long long view_stride, long long chan_stride, long long pix_stride,
                       int c, int hw, const float* __restrict__ xyz32, const long long* __restrict__ knn, int np, int k,
                       const float* __restrict__ tgt, float* __restrict__ X, int ldx) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int rows = np * k;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const long long g = knn[r];
        float* xr = X + (size_t)r * ldx;
        if (g < 0) {  // fewer than k valid pixels in this sphere: zero row
            for (int ch = lane; ch < c + 4; ch += 32) xr[ch] = 0.f;
            continue;
        }
        const long long v = g / hw, pix = g % hw;
        const float* f = feat + v * view_stride + pix * pix_stride;
        for (int ch = lane; ch < c; ch += 32) xr[ch] = __ldg(f + ch * chan_stride);
        if (lane == 0) {
            const int p = r / k;
            const float dx = xyz32[3 * g] - tgt[3 * p], dy = xyz32[3 * g + 1] - tgt[3 * p + 1],
                        dz = xyz32[3 * g + 2] - tgt[3 * p + 2];
            xr[c] = dx;
            xr[c + 1] = dy;
            xr[c + 2] = dz;
            xr[c + 3] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        }
    }
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

int mvk_unproject_views_batched(const double* kinv, const float* depth, const float* pose, int nviews, int h, int w,
                                double* xyz64, float* xyz32, unsigned char* mask, mvk_stream_t stream) {
    if (!kinv || !depth || !pose || !xyz64 || !xyz32 || !mask || nviews < 1 || h < 1 || w < 1) return MVK_ERR_INVALID_ARG;
    const size_t total = (size_t)nviews * h * w;
    unproject_batched_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kinv, depth, pose, nviews, h,
                                                                                             w, xyz64, xyz32, mask);
    MVK_LAUNCHED("unproject_batched_kernel");
    return MVK_OK;
}

size_t mvk_knn_batched_workspace_bytes(int nb, int npix, int nq, int grid_dim) {
    Arena a(nullptr, 0);
    kb_carve(a, nb > 0 ? nb : 1, npix > 0 ? npix : 1, nq, grid_dim);
    return a.off + 256;
}

int mvk_knn_pixels_batched(const double* xyz64, const float* xyz32, const unsigned char* mask, int nb, int npix,
                           const float* queries, const int* q_lengths, int nq, int k, int grid_dim, float cell_min,
                           int global_ids, void* ws, size_t ws_bytes, long long* out, int* far_counts,
                           mvk_stream_t stream) {
    if (!xyz64 || !xyz32 || !mask || !queries || !q_lengths || !out || nb < 1 || npix < 1 || nq < 0 || k < 1 ||
        k > KB_SLOTS || grid_dim < 4 || grid_dim > 256 || !(cell_min > 0.f))
        return MVK_ERR_INVALID_ARG;
    if ((size_t)nb * npix > 0x7fffffffull || (size_t)nb * grid_dim * grid_dim * grid_dim > 0x7ffffff0ull) return MVK_ERR_RANGE;
    if (!ws || ws_bytes < mvk_knn_batched_workspace_bytes(nb, npix, nq, grid_dim)) return MVK_ERR_WORKSPACE;
    if (nq == 0) return MVK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = grid_dim;
    Arena a(ws, ws_bytes);
    KbWork w = kb_carve(a, nb, npix, nq, D);
    const size_t cells = (size_t)nb * D * D * D, keys = (size_t)nb * npix;
    MVK_CUDA(cudaMemsetAsync(w.count, 0, cells * sizeof(int), st));
    MVK_CUDA(cudaMemsetAsync(w.cursor, 0, cells * sizeof(int), st));
    kb_init<<<8, 256, 0, st>>>(w, q_lengths, nb);
    MVK_LAUNCHED("kb_init");
    kb_bbox<<<(unsigned)((keys + 255) / 256), 256, 0, st>>>(xyz32, mask, nb, npix, w);
    MVK_LAUNCHED("kb_bbox");
    kb_params<<<(nb + 63) / 64, 64, 0, st>>>(w, nb, D, cell_min);
    MVK_LAUNCHED("kb_params");
    kb_count<<<(unsigned)((keys + 255) / 256), 256, 0, st>>>(xyz32, mask, nb, npix, D, w);
    MVK_LAUNCHED("kb_count");
    int rc = exclusive_scan_i32(w.count, w.start, (int)cells, w.start + cells, w.scan_tmp, st);
    if (rc) return rc;
    kb_scatter<<<(unsigned)((keys + 255) / 256), 256, 0, st>>>(xyz64, nb, npix, w);
    MVK_LAUNCHED("kb_scatter");
    if (k <= 4) kb_query<4><<<(nq + 127) / 128, 128, 0, st>>>(queries, nq, nb, npix, D, k, global_ids, out, w);
    else kb_query<8><<<(nq + 127) / 128, 128, 0, st>>>(queries, nq, nb, npix, D, k, global_ids, out, w);
    MVK_LAUNCHED("kb_query");
    const int fblocks = num_sms() * 4;
    if (k <= 4) kb_far<4><<<fblocks, 256, 0, st>>>(queries, nq, nb, npix, D, k, global_ids, out, w);
    else kb_far<8><<<fblocks, 256, 0, st>>>(queries, nq, nb, npix, D, k, global_ids, out, w);
    MVK_LAUNCHED("kb_far");
    if (far_counts) MVK_CUDA(cudaMemcpyAsync(far_counts, w.far_count, nb * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return MVK_OK;
}

int mvk_fa_gather_views(const float* feat, long long view_stride, long long chan_stride, long long pix_stride, int c,
                        int hw, const float* xyz32, const long long* knn_global, int np, int k, const float* tgt_xyz,
                        float* X, int ldx, mvk_stream_t stream) {
    if (!feat || !xyz32 || !knn_global || !tgt_xyz || !X || c < 1 || hw < 1 || np < 0 || k < 1 || ldx < c + 4)
        return MVK_ERR_INVALID_ARG;
    if (np == 0) return MVK_OK;
    const size_t rows = (size_t)np * k;
    size_t blocks = (rows + 7) / 8, maxb = (size_t)num_sms() * 32;
    fa_gather_views_kernel<<<(unsigned)(blocks < maxb ? blocks : maxb), 256, 0, (cudaStream_t)stream>>>(
        feat, view_stride, chan_stride, pix_stride, c, hw, xyz32, knn_global, np, k, tgt_xyz, X, ldx);
    MVK_LAUNCHED("fa_gather_views_kernel");
    return MVK_OK;
}

}  // extern "C"
