// Radius neighbours on a hashed cell grid in HBM (replaces nanoflann, see include/mvk.h).
//
// Build (supports):   cell = floor(p / c) per axis (c = r*(1+1e-5), fp64), key = (batch, cz, cy, cx)
//                     -> open-addressing table in HBM (atomicCAS), per-cell count -> exclusive scan
//                     -> supports scattered into cell-contiguous float4 {x, y, z, index}.
// Query (one warp per query): lanes 0..26 probe the 27 surrounding cells, every lane walks its own
//                     cell's run, hits (fp32 d2 < r2, no FMA contraction: bit-identical to the
//                     reference metric) are ballot-compacted into a per-warp shared-memory list and
//                     rank-sorted by (d2, index) -- the order of the reference's batch_ordered_neighbors
//                     and of nanoflann up to exact-d2 ties.
#include <stdlib.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr int CELL_BIAS = 1 << 17;
constexpr int CELL_MAX = (1 << 18) - 2;
constexpr int CAND_CAP = 256;  // flattened candidate list per query (falls back to the per-cell walk beyond)

struct Grid {
    int* q_starts;              // [nb+1]
    int* s_starts;              // [nb+1]
    unsigned long long* keys;   // [cap]
    int* cnt;                   // [cap]
    int* start;                 // [cap]
    int* slot;                  // [ns]
    int* rank;                  // [ns]
    float4* sorted;             // [ns]
    int* scan_tmp;
    int* err;                   // [1]
    int cap;
};

int table_cap(int ns) {
    int cap = 64;
    while (cap < 2 * ns) cap <<= 1;
    return cap;
}

Grid carve(Arena& a, int nq, int ns, int nb) {
    Grid g;
    g.cap = table_cap(ns);
    g.q_starts = a.take<int>(nb + 1);
    g.s_starts = a.take<int>(nb + 1);
    g.keys = a.take<unsigned long long>(g.cap);
    g.cnt = a.take<int>(g.cap);
    g.start = a.take<int>(g.cap);
    g.slot = a.take<int>(ns > 0 ? ns : 1);
    g.rank = a.take<int>(ns > 0 ? ns : 1);
    g.sorted = a.take<float4>(ns > 0 ? ns : 1);
    g.scan_tmp = a.take<int>(scan_tmp_ints(g.cap));
    g.err = a.take<int>(1);
    return g;
}

__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}

__device__ __forceinline__ unsigned long long pack_key(int b, int cx, int cy, int cz) {
    return ((unsigned long long)b << 54) | ((unsigned long long)cz << 36) |
           ((unsigned long long)cy << 18) | (unsigned long long)cx;
}

__device__ __forceinline__ int cell_coord(float v, double inv_cell) {
    return (int)floor((double)v * inv_cell) + CELL_BIAS;
}

__global__ void k_starts(const int* __restrict__ ql, const int* __restrict__ sl, int nb,
                         int* __restrict__ qs, int* __restrict__ ss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int a = 0, b = 0;
        for (int i = 0; i < nb; i++) {
            qs[i] = a;
            ss[i] = b;
            a += ql[i];
            b += sl[i];
        }
        qs[nb] = a;
        ss[nb] = b;
    }
}

__global__ void __launch_bounds__(256) k_insert(const float* __restrict__ s, int ns,
                                                const int* __restrict__ s_starts, int nb,
                                                double inv_cell, Grid g) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns) return;
    int b = batch_of(s_starts, nb, j);
    if (j >= s_starts[nb]) {  // support beyond the declared batch lengths: never a neighbour
        g.slot[j] = -1;
        return;
    }
    int cx = cell_coord(s[3 * j], inv_cell), cy = cell_coord(s[3 * j + 1], inv_cell),
        cz = cell_coord(s[3 * j + 2], inv_cell);
    if (cx < 1 || cy < 1 || cz < 1 || cx > CELL_MAX || cy > CELL_MAX || cz > CELL_MAX) {
        atomicExch(g.err, 1);
        g.slot[j] = -1;
        return;
    }
    unsigned long long key = pack_key(b, cx, cy, cz);
    unsigned int mask = g.cap - 1;
    unsigned int h = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&g.keys[h], EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) break;
        h = (h + 1) & mask;
    }
    g.slot[j] = (int)h;
    g.rank[j] = atomicAdd(&g.cnt[h], 1);
}

__global__ void __launch_bounds__(256) k_scatter(const float* __restrict__ s, int ns, Grid g) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ns) return;
    int sl = g.slot[j];
    if (sl < 0) return;
    g.sorted[g.start[sl] + g.rank[j]] =
        make_float4(s[3 * j], s[3 * j + 1], s[3 * j + 2], __int_as_float(j));
}

template <bool WRITE, typename OutT>
__global__ void __launch_bounds__(256)
k_query(const float* __restrict__ q, int nq, int nb, float r2, double inv_cell, Grid g,
        int* __restrict__ counts, int* __restrict__ max_count, int list_cap, int width,
        OutT* __restrict__ out, int ns, int cand_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    // hit list of this warp: one 64-bit key per hit = (d2 bits << 32) | support index.  d2 >= 0, so the
    // unsigned order of the keys IS the reference order (d2 ascending, index ascending on exact ties)
    unsigned long long* l_key = (unsigned long long*)smem_raw + (size_t)wib * list_cap;
    // candidate slots of this warp (WRITE mode only): positions in g.sorted of every point of the 27 cells
    int* own = (int*)((unsigned long long*)smem_raw + (size_t)wpb * list_cap) + (size_t)wib * CAND_CAP;
    const unsigned int mask = g.cap - 1;
    const int nq_valid = min(nq, g.q_starts[nb]);
    if (*g.err != 0) {  // a support fell outside the indexable cell range: report, do nothing
        if (max_count && blockIdx.x == 0 && threadIdx.x == 0) *max_count = -1;
        return;
    }

    for (int i = blockIdx.x * wpb + wib; i < nq; i += gridDim.x * wpb) {
        int nfound = 0;
        if (i < nq_valid) {
            const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            const int b = batch_of(g.q_starts, nb, i);
            int cstart = 0, ccnt = 0;
            if (lane < 27) {
                int cx = cell_coord(qx, inv_cell) + (lane % 3) - 1;
                int cy = cell_coord(qy, inv_cell) + ((lane / 3) % 3) - 1;
                int cz = cell_coord(qz, inv_cell) + (lane / 9) - 1;
                if (cx >= 1 && cy >= 1 && cz >= 1 && cx <= CELL_MAX && cy <= CELL_MAX &&
                    cz <= CELL_MAX) {
                    unsigned long long key = pack_key(b, cx, cy, cz);
                    unsigned int h = hash_key(key) & mask;
                    while (true) {
                        unsigned long long k = g.keys[h];
                        if (k == key) {
                            cstart = g.start[h];
                            ccnt = g.cnt[h];
                            break;
                        }
                        if (k == EMPTY_KEY) break;
                        h = (h + 1) & mask;
                    }
                }
            }
            // The 27 cells hold very different numbers of points (most are empty on surface data), so
            // walking them lane-per-cell leaves most lanes idle: flatten the candidates first (prefix sum of
            // the cell populations, candidate positions staged in shared memory), then test 32 candidates
            // per trip.  The order of the hits does not matter: the rows are rank-sorted afterwards.
            int incl = ccnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            bool flat = WRITE && total <= cand_cap;
            if (flat) {
                const int pre = incl - ccnt;
                for (int t = 0; t < ccnt; t++) own[pre + t] = cstart + t;
                __syncwarp();
                for (int base = 0; base < total; base += 32) {
                    const int cpos = base + lane;
                    bool hit = false;
                    float d2 = 0.f;
                    int idx = 0;
                    if (cpos < total) {
                        const float4 c = g.sorted[own[cpos]];
                        const float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
                        d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                        idx = __float_as_int(c.w);
                        hit = d2 < r2;
                    }
                    const unsigned int m = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = nfound + __popc(m & ((1u << lane) - 1));
                        if (pos < list_cap)
                            l_key[pos] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;
                    }
                    nfound += __popc(m);
                }
                __syncwarp();
            }
            int maxc = flat ? 0 : ccnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
            for (int it = 0; it < maxc; it++) {
                bool hit = false;
                float d2 = 0.f;
                int idx = 0;
                if (it < ccnt) {
                    float4 c = g.sorted[cstart + it];
                    float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
                    d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    idx = __float_as_int(c.w);
                    hit = d2 < r2;
                }
                unsigned int m = __ballot_sync(0xffffffffu, hit);
                if (WRITE && hit) {
                    int pos = nfound + __popc(m & ((1u << lane) - 1));
                    if (pos < list_cap)
                        l_key[pos] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;
                }
                nfound += __popc(m);
            }
        }
        if (!WRITE) {
            if (lane == 0) {
                counts[i] = nfound;
                if (nfound > 0) atomicMax(max_count, nfound);
            }
        } else {
            if (max_count && lane == 0) {  // single-pass protocol: the caller checks max_count <= list_cap
                if (counts) counts[i] = nfound;
                if (nfound > 0) atomicMax(max_count, nfound);
            }
            __syncwarp();
            int n = min(nfound, list_cap);
            OutT* row = out + (size_t)i * width;
            if ((n & 1) && lane == 0 && n < list_cap) l_key[n] = 0xFFFFFFFFFFFFFFFFull;  // pad to an even count
            __syncwarp();
            const int n2 = (n + 1) & ~1;
            for (int e = lane; e < n; e += 32) {
                const unsigned long long my = l_key[e];
                int rank = 0;
                for (int f = 0; f < n2; f += 2) {  // two keys per 128-bit shared load (warp-wide broadcast)
                    const ulonglong2 kk = *(const ulonglong2*)(l_key + f);
                    rank += (kk.x < my ? 1 : 0) + (kk.y < my ? 1 : 0);
                }
                if (rank < width) row[rank] = (OutT)(unsigned int)(my & 0xFFFFFFFFull);
            }
            for (int p = n + lane; p < width; p += 32) row[p] = (OutT)ns;
            __syncwarp();
        }
    }
}

int flat_cap() { return CAND_CAP; }

int build(const float* s, int ns, const int* ql, const int* sl, int nb, float radius, Grid& g,
          cudaStream_t st) {
    double inv_cell = 1.0 / ((double)radius * 1.00001);
    k_starts<<<1, 32, 0, st>>>(ql, sl, nb, g.q_starts, g.s_starts);
    MVK_LAUNCHED("k_starts");
    MVK_CUDA(cudaMemsetAsync(g.keys, 0xFF, sizeof(unsigned long long) * g.cap, st));
    MVK_CUDA(cudaMemsetAsync(g.cnt, 0, sizeof(int) * g.cap, st));
    MVK_CUDA(cudaMemsetAsync(g.err, 0, sizeof(int), st));
    if (ns > 0) {
        k_insert<<<(ns + 255) / 256, 256, 0, st>>>(s, ns, g.s_starts, nb, inv_cell, g);
        MVK_LAUNCHED("k_insert");
    }
    int rc = exclusive_scan_i32(g.cnt, g.start, g.cap, nullptr, g.scan_tmp, st);
    if (rc) return rc;
    if (ns > 0) {
        k_scatter<<<(ns + 255) / 256, 256, 0, st>>>(s, ns, g);
        MVK_LAUNCHED("k_scatter");
    }
    return MVK_OK;
}

template <typename OutT>
int fill(const float* q, int nq, const float* s, int ns, int nb, float radius, void* ws,
         size_t ws_bytes, int max_count, int width, OutT* out, cudaStream_t st) {
    if (nq < 0 || ns < 0 || nb < 1 || nb > 1023 || width < 1 || !out) return MVK_ERR_INVALID_ARG;
    if (ws_bytes < mvk_neighbors_workspace_bytes(nq, ns, nb)) return MVK_ERR_WORKSPACE;
    if (nq == 0) return MVK_OK;
    Arena a(ws, ws_bytes);
    Grid g = carve(a, nq, ns, nb);
    int list_cap = max_count < 32 ? 32 : (max_count + 31) / 32 * 32;
    size_t per_warp = (size_t)list_cap * 8 + CAND_CAP * 4;
    if (per_warp > 200 * 1024) return MVK_ERR_RANGE;
    int wpb = (int)((64 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
    size_t smem = per_warp * wpb;
    auto kern = k_query<true, OutT>;
    MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (nq + wpb - 1) / wpb;
    int max_blocks = num_sms() * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    float r2 = radius * radius;
    double inv_cell = 1.0 / ((double)radius * 1.00001);
    kern<<<blocks, wpb * 32, smem, st>>>(q, nq, nb, r2, inv_cell, g, nullptr, nullptr, list_cap,
                                         width, out, ns, flat_cap());
    MVK_LAUNCHED("k_query<fill>");
    return MVK_OK;
}

template <typename OutT>
int query_capped(const float* q, int nq, const float* s, int ns, const int* ql, const int* sl, int nb, float radius,
                 void* ws, size_t ws_bytes, int width, int list_cap, OutT* out, int* counts, int* max_count,
                 int reuse_grid, cudaStream_t st) {
    if (nq < 0 || ns < 0 || nb < 1 || nb > 1023 || !(radius > 0.f) || width < 1 || list_cap < width || !out ||
        !max_count)
        return MVK_ERR_INVALID_ARG;
    if (ws_bytes < mvk_neighbors_workspace_bytes(nq, ns, nb) || !ws) return MVK_ERR_WORKSPACE;
    Arena a(ws, ws_bytes);
    Grid g = carve(a, nq, ns, nb);
    if (reuse_grid) {
        // the workspace still holds the cell grid of these supports at this radius (previous call): only
        // the batch offsets of the new queries are needed
        k_starts<<<1, 32, 0, st>>>(ql, sl, nb, g.q_starts, g.s_starts);
        MVK_LAUNCHED("k_starts");
    } else {
        int rc = build(s, ns, ql, sl, nb, radius, g, st);
        if (rc) return rc;
    }
    MVK_CUDA(cudaMemsetAsync(max_count, 0, sizeof(int), st));
    if (nq == 0) return MVK_OK;
    list_cap = (list_cap + 31) / 32 * 32;
    size_t per_warp = (size_t)list_cap * 8 + CAND_CAP * 4;
    if (per_warp > 200 * 1024) return MVK_ERR_RANGE;
    int wpb = (int)((64 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 8 ? 8 : wpb);
    size_t smem = per_warp * wpb;
    auto kern = k_query<true, OutT>;
    MVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (nq + wpb - 1) / wpb;
    int max_blocks = num_sms() * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    float r2 = radius * radius;
    double inv_cell = 1.0 / ((double)radius * 1.00001);
    kern<<<blocks, wpb * 32, smem, st>>>(q, nq, nb, r2, inv_cell, g, counts, max_count, list_cap, width, out, ns,
                                         flat_cap());
    MVK_LAUNCHED("k_query<capped>");
    return MVK_OK;
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

size_t mvk_neighbors_workspace_bytes(int nq, int ns, int nb) {
    Arena a(nullptr, 0);
    carve(a, nq, ns, nb < 1 ? 1 : nb);
    return a.off + 256;
}

int mvk_neighbors_count(const float* queries, int nq, const float* supports, int ns,
                        const int* q_lengths, const int* s_lengths, int nb, float radius, void* ws,
                        size_t ws_bytes, int* counts, int* max_count, mvk_stream_t stream) {
    if (nq < 0 || ns < 0 || nb < 1 || nb > 1023 || !(radius > 0.f) || !counts || !max_count)
        return MVK_ERR_INVALID_ARG;
    if (ws_bytes < mvk_neighbors_workspace_bytes(nq, ns, nb) || !ws) return MVK_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    Grid g = carve(a, nq, ns, nb);
    int rc = build(supports, ns, q_lengths, s_lengths, nb, radius, g, st);
    if (rc) return rc;
    MVK_CUDA(cudaMemsetAsync(max_count, 0, sizeof(int), st));
    if (nq == 0) return MVK_OK;
    int wpb = 8;
    int blocks = (nq + wpb - 1) / wpb;
    int max_blocks = num_sms() * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    float r2 = radius * radius;  // fp32 product like neighbors.cpp:226
    double inv_cell = 1.0 / ((double)radius * 1.00001);
    k_query<false, int><<<blocks, wpb * 32, 0, st>>>(queries, nq, nb, r2, inv_cell, g, counts,
                                                     max_count, 0, 0, nullptr, ns, 0);
    MVK_LAUNCHED("k_query<count>");
    return MVK_OK;
}

int mvk_neighbors_fill(const float* queries, int nq, const float* supports, int ns,
                       const int* q_lengths, const int* s_lengths, int nb, float radius, void* ws,
                       size_t ws_bytes, int max_count, int width, int* out, mvk_stream_t stream) {
    (void)q_lengths;
    (void)s_lengths;
    return fill<int>(queries, nq, supports, ns, nb, radius, ws, ws_bytes, max_count, width, out,
                     (cudaStream_t)stream);
}

int mvk_neighbors_fill_i64(const float* queries, int nq, const float* supports, int ns,
                           const int* q_lengths, const int* s_lengths, int nb, float radius,
                           void* ws, size_t ws_bytes, int max_count, int width, long long* out,
                           mvk_stream_t stream) {
    (void)q_lengths;
    (void)s_lengths;
    return fill<long long>(queries, nq, supports, ns, nb, radius, ws, ws_bytes, max_count, width,
                           out, (cudaStream_t)stream);
}

int mvk_neighbors_query_capped(const float* queries, int nq, const float* supports, int ns, const int* q_lengths,
                               const int* s_lengths, int nb, float radius, void* ws, size_t ws_bytes, int width,
                               int list_cap, void* out, int out_is_i64, int* counts, int* max_count,
                               int reuse_grid, mvk_stream_t stream) {
    if (out_is_i64)
        return query_capped<long long>(queries, nq, supports, ns, q_lengths, s_lengths, nb, radius, ws, ws_bytes, width,
                                       list_cap, (long long*)out, counts, max_count, reuse_grid, (cudaStream_t)stream);
    return query_capped<int>(queries, nq, supports, ns, q_lengths, s_lengths, nb, radius, ws, ws_bytes, width, list_cap,
                             (int*)out, counts, max_count, reuse_grid, (cudaStream_t)stream);
}

int mvk_batch_neighbors_host(const float* qh, int nq, const float* sh, int ns, const int* qlh,
                             const int* slh, int nb, float radius, int** out_host, int* width) {
    if (!out_host || !width) return MVK_ERR_INVALID_ARG;
    *out_host = nullptr;
    *width = 0;
    if (nq <= 0) return MVK_ERR_EMPTY;
    float *dq = nullptr, *ds = nullptr;
    int *dql = nullptr, *dsl = nullptr, *dcounts = nullptr, *dmax = nullptr, *dout = nullptr;
    void* ws = nullptr;
    size_t wsb = mvk_neighbors_workspace_bytes(nq, ns, nb);
    int rc = MVK_OK, hmax = 0;
    cudaError_t e;
#define HCHK(x)                                  \
    if ((e = (x)) != cudaSuccess) {              \
        rc = cuda_fail(e, #x);                   \
        goto done;                               \
    }
    HCHK(cudaMalloc(&dq, sizeof(float) * 3 * (size_t)nq));
    HCHK(cudaMalloc(&ds, sizeof(float) * 3 * (size_t)(ns > 0 ? ns : 1)));
    HCHK(cudaMalloc(&dql, sizeof(int) * nb));
    HCHK(cudaMalloc(&dsl, sizeof(int) * nb));
    HCHK(cudaMalloc(&dcounts, sizeof(int) * (size_t)nq));
    HCHK(cudaMalloc(&dmax, sizeof(int)));
    HCHK(cudaMalloc(&ws, wsb));
    HCHK(cudaMemcpyAsync(dq, qh, sizeof(float) * 3 * (size_t)nq, cudaMemcpyHostToDevice, 0));
    HCHK(cudaMemcpyAsync(ds, sh, sizeof(float) * 3 * (size_t)ns, cudaMemcpyHostToDevice, 0));
    HCHK(cudaMemcpyAsync(dql, qlh, sizeof(int) * nb, cudaMemcpyHostToDevice, 0));
    HCHK(cudaMemcpyAsync(dsl, slh, sizeof(int) * nb, cudaMemcpyHostToDevice, 0));
    rc = mvk_neighbors_count(dq, nq, ds, ns, dql, dsl, nb, radius, ws, wsb, dcounts, dmax, 0);
    if (rc) goto done;
    HCHK(cudaMemcpy(&hmax, dmax, sizeof(int), cudaMemcpyDeviceToHost));
    if (hmax < 1) {
        rc = MVK_ERR_EMPTY;
        goto done;
    }
    HCHK(cudaMalloc(&dout, sizeof(int) * (size_t)nq * hmax));
    rc = mvk_neighbors_fill(dq, nq, ds, ns, dql, dsl, nb, radius, ws, wsb, hmax, hmax, dout, 0);
    if (rc) goto done;
    *out_host = (int*)malloc(sizeof(int) * (size_t)nq * hmax);
    HCHK(cudaMemcpy(*out_host, dout, sizeof(int) * (size_t)nq * hmax, cudaMemcpyDeviceToHost));
    *width = hmax;
done:
#undef HCHK
    cudaFree(dq);
    cudaFree(ds);
    cudaFree(dql);
    cudaFree(dsl);
    cudaFree(dcounts);
    cudaFree(dmax);
    cudaFree(dout);
    cudaFree(ws);
    if (rc && *out_host) {
        free(*out_host);
        *out_host = nullptr;
    }
    return rc;
}

}  // extern "C"
