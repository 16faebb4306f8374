// Grid (voxel barycentre) subsampling on a hashed voxel grid in HBM -- bit-exact with the
// reference C++ (grid_subsampling.cpp:5-210) INCLUDING its output order, which is the iteration
// order of libstdc++'s std::unordered_map<size_t, SampledData>.
//
// Pipeline (all batch elements in the same launches):
//   1. per-element min / max corner (order independent)            ss_minmax
//   2. origin = floor(min * (1/dl)) * dl, NX, NY  (fp32, no FMA)    ss_params
//   3. voxel key per point (true fp32 divisions), hash insert,
//      per-voxel count + first point index                          ss_insert
//   4. scan -> per-voxel segments; points grouped per voxel         ss_fill
//   5. first-appearance rank of every voxel (scan of "is first")    ss_voxel
//   6. per voxel: sort its few point indices, then SEQUENTIAL fp32
//      sums in original point order (grid_subsampling.h:74-79),
//      barycentre = sum * (float)(1.0 / count)                      ss_reduce  (+ ss_labels)
//   7. emulate the unordered_map node list: for every bucket count
//      B of libstdc++'s growth schedule the list is re-ordered by
//      (first occupancy time of the bucket DESC, insertion time DESC) ss_order (one CTA / element)
//   8. max_p truncation, compaction                                 ss_lengths, ss_emit
#include <stdlib.h>

#include "common.cuh"

namespace mvk {
namespace {

constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long KEY_LIMIT = 1ull << 54;
constexpr int MAX_LABELS = 64;  // distinct labels per voxel handled by the majority vote

// libstdc++ (GCC 13) _Prime_rehash_policy bucket counts, max_load_factor 1, growth factor 2.
// Pinned against the live std::unordered_map in tests/test_oracle.py (oracle/stl_probe.cpp).
__constant__ unsigned int c_sched[28] = {
    13u,       29u,       59u,        127u,       257u,       541u,       1109u,
    2357u,     5087u,     10273u,     20753u,     42043u,     85229u,     172933u,
    351061u,   712697u,   1447153u,   2938679u,   5967347u,   12117689u,  24607243u,
    49969847u, 101473717u, 206062531u, 418451333u, 849749479u, 1725587117u, 3504151727u};

struct ElemParam {
    float ox, oy, oz;
    unsigned long long nx, nxny;
};

struct Work {
    int* starts;                // [nb+1] point starts
    int* mm;                    // [nb*6] ordered-int min xyz, max xyz
    ElemParam* ep;              // [nb]
    unsigned long long* tkeys;  // [cap]
    int* tcnt;                  // [cap]
    int* tfirst;                // [cap]
    int* tstart;                // [cap]
    int* pslot;                 // [n]
    int* prank;                 // [n]
    int* seg;                   // [n]
    int* isfirst;               // [n]
    int* fa;                    // [n]
    int* vslot;                 // [n]
    unsigned long long* vkey;   // [n]
    int* vstart;                // [nb+1]
    int* total;                 // [1]  number of voxels over all elements
    float* vsum;                // [n*3]
    float* vfeat;               // [n*fdim]
    int* vlab;                  // [n*ldim]
    int* lstA;                  // [n]
    int* lstB;                  // [n]
    int* order;                 // [n]
    int* pbk;                   // [n]
    int* bfirst;                // [tabsz]
    int* bcnt;                  // [tabsz]
    int* bcur;                  // [tabsz]
    int* ostart;                // [nb+1]
    int* scan_tmp;
    int* err;                   // [1]
    int cap;
};

int table_cap(int n) {
    int cap = 64;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}
size_t tab_size(int n, int nb) { return (size_t)n * 9 / 4 + 64 * (size_t)nb + 64; }

Work carve(Arena& a, int n, int nb, int fdim, int ldim) {
    Work w;
    int n1 = n > 0 ? n : 1;
    w.cap = table_cap(n);
    w.starts = a.take<int>(nb + 1);
    w.mm = a.take<int>(nb * 6);
    w.ep = a.take<ElemParam>(nb);
    w.tkeys = a.take<unsigned long long>(w.cap);
    w.tcnt = a.take<int>(w.cap);
    w.tfirst = a.take<int>(w.cap);
    w.tstart = a.take<int>(w.cap);
    w.pslot = a.take<int>(n1);
    w.prank = a.take<int>(n1);
    w.seg = a.take<int>(n1);
    w.isfirst = a.take<int>(n1);
    w.fa = a.take<int>(n1);
    w.vslot = a.take<int>(n1);
    w.vkey = a.take<unsigned long long>(n1);
    w.vstart = a.take<int>(nb + 1);
    w.total = a.take<int>(1);
    w.vsum = a.take<float>((size_t)n1 * 3);
    w.vfeat = a.take<float>((size_t)n1 * (fdim > 0 ? fdim : 1));
    w.vlab = a.take<int>((size_t)n1 * (ldim > 0 ? ldim : 1));
    w.lstA = a.take<int>(n1);
    w.lstB = a.take<int>(n1);
    w.order = a.take<int>(n1);
    w.pbk = a.take<int>(n1);
    size_t ts = tab_size(n, nb);
    w.bfirst = a.take<int>(ts);
    w.bcnt = a.take<int>(ts);
    w.bcur = a.take<int>(ts);
    w.ostart = a.take<int>(nb + 1);
    size_t st = scan_tmp_ints(w.cap > n1 ? w.cap : n1);
    w.scan_tmp = a.take<int>(st);
    w.err = a.take<int>(1);
    return w;
}

__device__ __forceinline__ int f2ord(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}

__global__ void ss_starts(const int* __restrict__ len, int nb, int n, Work w) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int a = 0;
        for (int i = 0; i < nb; i++) {
            w.starts[i] = a;
            a += len[i];
            if (a > n) a = n;
        }
        w.starts[nb] = a;
        for (int i = 0; i < nb; i++) {
            for (int k = 0; k < 3; k++) {
                w.mm[i * 6 + k] = 0x7fffffff;
                w.mm[i * 6 + 3 + k] = (int)0x80000000;
            }
        }
    }
}

__global__ void __launch_bounds__(256) ss_minmax(const float* __restrict__ p, int n, int nb, Work w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    bool in = i < w.starts[nb];
    int b = in ? batch_of(w.starts, nb, i) : -1;
    int v[6];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int o = in ? f2ord(p[3 * i + k]) : 0;
        v[k] = in ? o : 0x7fffffff;
        v[3 + k] = in ? o : (int)0x80000000;
    }
    int b0 = __shfl_sync(0xffffffffu, b, 0);
    bool uniform = __all_sync(0xffffffffu, b == b0 || b < 0) && b0 >= 0;
    if (uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                v[k] = min(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
                v[3 + k] = max(v[3 + k], __shfl_xor_sync(0xffffffffu, v[3 + k], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicMin(&w.mm[b0 * 6 + k], v[k]);
                atomicMax(&w.mm[b0 * 6 + 3 + k], v[3 + k]);
            }
        }
    } else if (in) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&w.mm[b * 6 + k], v[k]);
            atomicMax(&w.mm[b * 6 + 3 + k], v[3 + k]);
        }
    }
}

// grid_subsampling.cpp:27-31 in fp32 without contraction.
__global__ void ss_params(int nb, float dl, Work w) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    ElemParam e;
    e.ox = e.oy = e.oz = 0.f;
    e.nx = e.nxny = 1;
    if (w.starts[b + 1] > w.starts[b]) {
        float inv = __fdiv_rn(1.0f, dl);
        float mnx = ord2f(w.mm[b * 6]), mny = ord2f(w.mm[b * 6 + 1]), mnz = ord2f(w.mm[b * 6 + 2]);
        float mxx = ord2f(w.mm[b * 6 + 3]), mxy = ord2f(w.mm[b * 6 + 4]);
        e.ox = __fmul_rn(floorf(__fmul_rn(mnx, inv)), dl);
        e.oy = __fmul_rn(floorf(__fmul_rn(mny, inv)), dl);
        e.oz = __fmul_rn(floorf(__fmul_rn(mnz, inv)), dl);
        unsigned long long nx = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxx, e.ox), dl)) + 1ull;
        unsigned long long ny = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(mxy, e.oy), dl)) + 1ull;
        e.nx = nx;
        e.nxny = nx * ny;
    }
    w.ep[b] = e;
}

__global__ void __launch_bounds__(256) ss_insert(const float* __restrict__ p, int n, int nb, float dl,
                                                 Work w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i >= w.starts[nb]) {
        w.pslot[i] = -1;
        return;
    }
    int b = batch_of(w.starts, nb, i);
    ElemParam e = w.ep[b];
    // grid_subsampling.cpp:53-56
    unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * i], e.ox), dl));
    unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * i + 1], e.oy), dl));
    unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * i + 2], e.oz), dl));
    unsigned long long vk = ix + e.nx * iy + e.nxny * iz;
    if (vk >= KEY_LIMIT) {
        atomicExch(w.err, 1);
        w.pslot[i] = -1;
        return;
    }
    unsigned long long key = ((unsigned long long)b << 54) | vk;
    unsigned int mask = w.cap - 1;
    unsigned int h = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&w.tkeys[h], EMPTY_KEY, key);
        if (prev == EMPTY_KEY || prev == key) break;
        h = (h + 1) & mask;
    }
    w.pslot[i] = (int)h;
    w.prank[i] = atomicAdd(&w.tcnt[h], 1);
    atomicMin(&w.tfirst[h], i);
}

__global__ void __launch_bounds__(256) ss_fill(int n, Work w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = w.pslot[i];
    if (s < 0) {
        w.isfirst[i] = 0;
        return;
    }
    w.seg[w.tstart[s] + w.prank[i]] = i;
    w.isfirst[i] = (w.tfirst[s] == i) ? 1 : 0;
}

__global__ void __launch_bounds__(256) ss_voxel(int n, int nb, Work w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && w.isfirst[i]) {
        int v = w.fa[i];
        int s = w.pslot[i];
        w.vslot[v] = s;
        w.vkey[v] = w.tkeys[s] & (KEY_LIMIT - 1);
    }
    if (i <= nb) {
        int st = w.starts[i < nb ? i : nb];
        w.vstart[i] = (i < nb && st < n) ? w.fa[st] : *w.total;
    }
}

// One thread per voxel: order its point indices, then accumulate sequentially in fp32.
__global__ void __launch_bounds__(128) ss_reduce(const float* __restrict__ p,
                                                 const float* __restrict__ f, int fdim, Work w) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= *w.total) return;
    int s = w.vslot[v];
    int* seg = w.seg + w.tstart[s];
    int c = w.tcnt[s];
    for (int a = 1; a < c; a++) {  // insertion sort: segments hold a handful of points
        int x = seg[a];
        int bq = a - 1;
        while (bq >= 0 && seg[bq] > x) {
            seg[bq + 1] = seg[bq];
            bq--;
        }
        seg[bq + 1] = x;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int a = 0; a < c; a++) {
        int i = seg[a];
        sx = __fadd_rn(sx, p[3 * i]);
        sy = __fadd_rn(sy, p[3 * i + 1]);
        sz = __fadd_rn(sz, p[3 * i + 2]);
    }
    float wgt = (float)(1.0 / (double)c);  // grid_subsampling.cpp:87
    w.vsum[3 * (size_t)v] = __fmul_rn(sx, wgt);
    w.vsum[3 * (size_t)v + 1] = __fmul_rn(sy, wgt);
    w.vsum[3 * (size_t)v + 2] = __fmul_rn(sz, wgt);
    if (fdim > 0) {
        float fc = (float)c;  // :90
        for (int k = 0; k < fdim; k++) {
            float acc = 0.f;
            for (int a = 0; a < c; a++) acc = __fadd_rn(acc, f[(size_t)seg[a] * fdim + k]);
            w.vfeat[(size_t)v * fdim + k] = __fdiv_rn(acc, fc);
        }
    }
}

// Iteration order of std::unordered_map<int,int> holding `n` distinct labels inserted in the given
// order (direct simulation of the node list, n <= MAX_LABELS).
__device__ void small_stl_order(const int* lab, int n, int* order) {
    short nxt[MAX_LABELS + 1];
    short before[127];
    const int SENT = MAX_LABELS;
    nxt[SENT] = -1;
    unsigned int B = 1, next_resize = 0;
    int si = 0;
    before[0] = -1;
    for (int i = 0; i < n; i++) {
        if ((unsigned int)(i + 1) > next_resize) {
            unsigned int nB = c_sched[si++];
            for (unsigned int k = 0; k < nB; k++) before[k] = -1;
            int pp = nxt[SENT];
            nxt[SENT] = -1;
            unsigned int bb = 0;
            while (pp >= 0) {
                int nx = nxt[pp];
                unsigned int bk = (unsigned int)(((unsigned long long)(long long)lab[pp]) % nB);
                if (before[bk] < 0) {
                    nxt[pp] = nxt[SENT];
                    nxt[SENT] = (short)pp;
                    before[bk] = (short)SENT;
                    if (nxt[pp] >= 0) before[bb] = (short)pp;
                    bb = bk;
                } else {
                    nxt[pp] = nxt[before[bk]];
                    nxt[before[bk]] = (short)pp;
                }
                pp = nx;
            }
            B = nB;
            next_resize = B;
        }
        unsigned int bk = (unsigned int)(((unsigned long long)(long long)lab[i]) % B);
        if (before[bk] >= 0) {
            nxt[i] = nxt[before[bk]];
            nxt[before[bk]] = (short)i;
        } else {
            nxt[i] = nxt[SENT];
            nxt[SENT] = (short)i;
            if (nxt[i] >= 0)
                before[(unsigned int)(((unsigned long long)(long long)lab[nxt[i]]) % B)] = (short)i;
            before[bk] = (short)SENT;
        }
    }
    int k = 0;
    for (int pp = nxt[SENT]; pp >= 0; pp = nxt[pp]) order[k++] = pp;
}

// Majority label per voxel and label column: first maximum in unordered_map<int,int> iteration
// order (grid_subsampling.cpp:97-102).  Runs after ss_reduce (segments already index-sorted).
__global__ void __launch_bounds__(64) ss_labels(const int* __restrict__ cls, int ldim, Work w) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= *w.total) return;
    int s = w.vslot[v];
    const int* seg = w.seg + w.tstart[s];
    int c = w.tcnt[s];
    int lab[MAX_LABELS], cnt[MAX_LABELS], ord[MAX_LABELS];
    for (int k = 0; k < ldim; k++) {
        int nl = 0;
        for (int a = 0; a < c; a++) {
            int l = cls[(size_t)seg[a] * ldim + k];
            int j = 0;
            for (; j < nl; j++)
                if (lab[j] == l) break;
            if (j == nl) {
                if (nl == MAX_LABELS) {
                    atomicExch(w.err, 2);
                    break;
                }
                lab[nl] = l;
                cnt[nl] = 0;
                nl++;
            }
            cnt[j]++;
        }
        small_stl_order(lab, nl, ord);
        int best = ord[0];
        for (int t = 1; t < nl; t++)
            if (cnt[best] < cnt[ord[t]]) best = ord[t];
        w.vlab[(size_t)v * ldim + k] = lab[best];
    }
}

// One CTA per batch element: emission order of its voxels (local first-appearance ranks).
__global__ void __launch_bounds__(1024) ss_order(Work w) {
    __shared__ int s_total;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, T = blockDim.x;
    const int v0 = w.vstart[b];
    const int M = w.vstart[b + 1] - v0;
    if (M <= 0) return;
    const unsigned long long* key = w.vkey + v0;
    int* cur = w.lstA + v0;
    int* nxt = w.lstB + v0;
    int* pbk = w.pbk + v0;
    const size_t toff = (size_t)v0 * 9 / 4 + 64 * (size_t)b;
    int* bfirst = w.bfirst + toff;
    int* bcnt = w.bcnt + toff;
    int* bcur = w.bcur + toff;
    int n_done = 0;
    for (int si = 0; si < 28; si++) {
        const unsigned int B = c_sched[si];
        const int n = (unsigned int)M < B ? M : (int)B;
        for (unsigned int t = tid; t < B; t += T) {  // B <= 2.25 M + 64 (table slice size)
            bfirst[t] = 0x7fffffff;
            bcnt[t] = 0;
        }
        __syncthreads();
        for (int pp = tid; pp < n; pp += T) {
            int v = pp < n_done ? cur[pp] : pp;
            unsigned long long k = key[v];
            int bk = (k >> 32) ? (int)(k % B) : (int)((unsigned int)k % B);
            pbk[pp] = bk;
            atomicMin(&bfirst[bk], pp);
            atomicAdd(&bcnt[bk], 1);
        }
        __syncthreads();
        // group starts: buckets ordered by first-occupancy position DESCENDING
        {
            int chunk = (n + T - 1) / T;
            int c0 = tid * chunk, c1 = min(n, c0 + chunk);
            int local = 0;
            for (int pp = c0; pp < c1; pp++) {
                int bk = pbk[pp];
                if (bfirst[bk] == pp) local += bcnt[bk];
            }
            int pre = block_exclusive_scan(local, &s_total);
            int total = s_total;
            int run = pre;
            for (int pp = c0; pp < c1; pp++) {
                int bk = pbk[pp];
                if (bfirst[bk] == pp) {
                    run += bcnt[bk];
                    bcur[bk] = total - run;
                }
            }
        }
        __syncthreads();
        for (int pp = tid; pp < n; pp += T) {
            int slot = atomicAdd(&bcur[pbk[pp]], 1);
            nxt[slot] = pp;
        }
        __syncthreads();
        // inside a bucket group: insertion time DESCENDING; then positions -> voxel ranks
        for (int pp = tid; pp < n; pp += T) {
            int bk = pbk[pp];
            if (bfirst[bk] != pp) continue;
            int g1 = bcur[bk], c = bcnt[bk], g0 = g1 - c;
            for (int a = g0 + 1; a < g1; a++) {
                int x = nxt[a];
                int q = a - 1;
                while (q >= g0 && nxt[q] < x) {
                    nxt[q + 1] = nxt[q];
                    q--;
                }
                nxt[q + 1] = x;
            }
            for (int a = g0; a < g1; a++) {
                int q = nxt[a];
                nxt[a] = q < n_done ? cur[q] : q;
            }
        }
        __syncthreads();
        int* tmp = cur;
        cur = nxt;
        nxt = tmp;
        n_done = n;
        if (n == M) break;
    }
    int* order = w.order + v0;
    for (int k = tid; k < M; k += T) order[k] = cur[k];
}

__global__ void ss_lengths(int nb, int max_p, int* __restrict__ out_len, int* __restrict__ out_total,
                           Work w) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int a = 0;
        for (int b = 0; b < nb; b++) {
            int m = w.vstart[b + 1] - w.vstart[b];
            if (m > max_p) m = max_p;  // grid_subsampling.cpp:181-204
            w.ostart[b] = a;
            out_len[b] = m;
            a += m;
        }
        w.ostart[nb] = a;
        *out_total = (*w.err) ? -(*w.err) : a;
    }
}

__global__ void __launch_bounds__(256) ss_emit(int nb, int fdim, int ldim, float* __restrict__ op,
                                               float* __restrict__ of, int* __restrict__ ol, Work w) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *w.total) return;
    int b = batch_of(w.vstart, nb, t);
    int k = t - w.vstart[b];
    if (k >= w.ostart[b + 1] - w.ostart[b]) return;
    int v = w.vstart[b] + w.order[t];
    size_t o = (size_t)w.ostart[b] + k;
    op[3 * o] = w.vsum[3 * (size_t)v];
    op[3 * o + 1] = w.vsum[3 * (size_t)v + 1];
    op[3 * o + 2] = w.vsum[3 * (size_t)v + 2];
    for (int j = 0; j < fdim; j++) of[o * fdim + j] = w.vfeat[(size_t)v * fdim + j];
    for (int j = 0; j < ldim; j++) ol[o * ldim + j] = w.vlab[(size_t)v * ldim + j];
}

__global__ void __launch_bounds__(256) ss_rotate(const float* __restrict__ p, int n,
                                                 const int* __restrict__ len, int nb,
                                                 const float* __restrict__ rot, int transpose,
                                                 float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = 0, a = 0;
    while (b < nb - 1 && i >= a + len[b]) {
        a += len[b];
        b++;
    }
    const float* R = rot + 9 * b;
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float r0 = transpose ? R[c * 3 + 0] : R[0 * 3 + c];
        float r1 = transpose ? R[c * 3 + 1] : R[1 * 3 + c];
        float r2 = transpose ? R[c * 3 + 2] : R[2 * 3 + c];
        o[c] = __fadd_rn(__fadd_rn(__fmul_rn(x, r0), __fmul_rn(y, r1)), __fmul_rn(z, r2));
    }
    out[3 * i] = o[0];
    out[3 * i + 1] = o[1];
    out[3 * i + 2] = o[2];
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" {

size_t mvk_subsample_workspace_bytes(int n, int nb, int fdim, int ldim) {
    Arena a(nullptr, 0);
    carve(a, n, nb < 1 ? 1 : nb, fdim, ldim);
    return a.off + 256;
}

int mvk_grid_subsample(const float* points, int n, const float* features, int fdim,
                       const int* labels, int ldim, const int* lengths, int nb, float sampleDl,
                       int max_p, void* ws, size_t ws_bytes, float* out_points,
                       float* out_features, int* out_labels, int* out_lengths, int* out_total,
                       mvk_stream_t stream) {
    if (n < 0 || nb < 1 || nb > 1023 || !(sampleDl > 0.f) || !out_points || !out_lengths || !out_total)
        return MVK_ERR_INVALID_ARG;
    if (!features) fdim = 0;
    if (!labels) ldim = 0;
    if ((fdim > 0 && !out_features) || (ldim > 0 && !out_labels)) return MVK_ERR_INVALID_ARG;
    if (!ws || ws_bytes < mvk_subsample_workspace_bytes(n, nb, fdim, ldim)) return MVK_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    Work w = carve(a, n, nb, fdim, ldim);
    if (max_p < 1) max_p = n;  // grid_subsampling.cpp:134-135

    ss_starts<<<1, 32, 0, st>>>(lengths, nb, n, w);
    MVK_LAUNCHED("ss_starts");
    MVK_CUDA(cudaMemsetAsync(w.tkeys, 0xFF, sizeof(unsigned long long) * w.cap, st));
    MVK_CUDA(cudaMemsetAsync(w.tcnt, 0, sizeof(int) * w.cap, st));
    MVK_CUDA(cudaMemsetAsync(w.tfirst, 0x7F, sizeof(int) * w.cap, st));
    MVK_CUDA(cudaMemsetAsync(w.err, 0, sizeof(int), st));
    MVK_CUDA(cudaMemsetAsync(w.total, 0, sizeof(int), st));
    int nblk = (n + 255) / 256;
    if (n > 0) {
        ss_minmax<<<nblk, 256, 0, st>>>(points, n, nb, w);
        MVK_LAUNCHED("ss_minmax");
    }
    ss_params<<<(nb + 63) / 64, 64, 0, st>>>(nb, sampleDl, w);
    MVK_LAUNCHED("ss_params");
    if (n > 0) {
        ss_insert<<<nblk, 256, 0, st>>>(points, n, nb, sampleDl, w);
        MVK_LAUNCHED("ss_insert");
        int rc = exclusive_scan_i32(w.tcnt, w.tstart, w.cap, nullptr, w.scan_tmp, st);
        if (rc) return rc;
        ss_fill<<<nblk, 256, 0, st>>>(n, w);
        MVK_LAUNCHED("ss_fill");
        rc = exclusive_scan_i32(w.isfirst, w.fa, n, w.total, w.scan_tmp, st);
        if (rc) return rc;
    }
    {
        int m = n > nb + 1 ? n : nb + 1;
        ss_voxel<<<(m + 255) / 256, 256, 0, st>>>(n, nb, w);
        MVK_LAUNCHED("ss_voxel");
    }
    if (n > 0) {
        ss_reduce<<<(n + 127) / 128, 128, 0, st>>>(points, features, fdim, w);
        MVK_LAUNCHED("ss_reduce");
        if (ldim > 0) {
            ss_labels<<<(n + 63) / 64, 64, 0, st>>>(labels, ldim, w);
            MVK_LAUNCHED("ss_labels");
        }
        ss_order<<<nb, 1024, 0, st>>>(w);
        MVK_LAUNCHED("ss_order");
    }
    ss_lengths<<<1, 32, 0, st>>>(nb, max_p, out_lengths, out_total, w);
    MVK_LAUNCHED("ss_lengths");
    if (n > 0) {
        ss_emit<<<nblk, 256, 0, st>>>(nb, fdim, ldim, out_points, out_features, out_labels, w);
        MVK_LAUNCHED("ss_emit");
    }
    return MVK_OK;
}

int mvk_rotate_batch(const float* points, int n, const int* lengths, int nb, const float* rot,
                     int transpose, float* out, mvk_stream_t stream) {
    if (n < 0 || nb < 1 || !points || !lengths || !rot || !out) return MVK_ERR_INVALID_ARG;
    if (n == 0) return MVK_OK;
    ss_rotate<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, n, lengths, nb, rot, transpose, out);
    MVK_LAUNCHED("ss_rotate");
    return MVK_OK;
}

int mvk_grid_subsample_host(const float* ph, int n, const float* fh, int fdim, const int* lh,
                            int ldim, const int* lenh, int nb, float dl, int max_p, float** oph,
                            float** ofh, int** olh, int* out_lengths_host, int* out_total) {
    if (!oph || !out_lengths_host || !out_total || n <= 0) return n <= 0 ? MVK_ERR_EMPTY : MVK_ERR_INVALID_ARG;
    if (!fh) fdim = 0;
    if (!lh) ldim = 0;
    float *dp = nullptr, *df = nullptr, *dop = nullptr, *dof = nullptr;
    int *dl_ = nullptr, *dlen = nullptr, *dol = nullptr, *dolen = nullptr, *dtot = nullptr;
    void* ws = nullptr;
    size_t wsb = mvk_subsample_workspace_bytes(n, nb, fdim, ldim);
    int rc = MVK_OK, m = 0;
    cudaError_t e;
    *oph = nullptr;
    if (ofh) *ofh = nullptr;
    if (olh) *olh = nullptr;
#define HCHK(x)                     \
    if ((e = (x)) != cudaSuccess) { \
        rc = cuda_fail(e, #x);      \
        goto done;                  \
    }
    HCHK(cudaMalloc(&dp, sizeof(float) * 3 * (size_t)n));
    HCHK(cudaMalloc(&dop, sizeof(float) * 3 * (size_t)n));
    HCHK(cudaMalloc(&dlen, sizeof(int) * nb));
    HCHK(cudaMalloc(&dolen, sizeof(int) * nb));
    HCHK(cudaMalloc(&dtot, sizeof(int)));
    HCHK(cudaMalloc(&ws, wsb));
    HCHK(cudaMemcpyAsync(dp, ph, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, 0));
    HCHK(cudaMemcpyAsync(dlen, lenh, sizeof(int) * nb, cudaMemcpyHostToDevice, 0));
    if (fdim > 0) {
        HCHK(cudaMalloc(&df, sizeof(float) * (size_t)n * fdim));
        HCHK(cudaMalloc(&dof, sizeof(float) * (size_t)n * fdim));
        HCHK(cudaMemcpyAsync(df, fh, sizeof(float) * (size_t)n * fdim, cudaMemcpyHostToDevice, 0));
    }
    if (ldim > 0) {
        HCHK(cudaMalloc(&dl_, sizeof(int) * (size_t)n * ldim));
        HCHK(cudaMalloc(&dol, sizeof(int) * (size_t)n * ldim));
        HCHK(cudaMemcpyAsync(dl_, lh, sizeof(int) * (size_t)n * ldim, cudaMemcpyHostToDevice, 0));
    }
    rc = mvk_grid_subsample(dp, n, df, fdim, dl_, ldim, dlen, nb, dl, max_p, ws, wsb, dop, dof, dol,
                            dolen, dtot, 0);
    if (rc) goto done;
    HCHK(cudaMemcpy(&m, dtot, sizeof(int), cudaMemcpyDeviceToHost));
    if (m < 0) {
        rc = MVK_ERR_RANGE;
        goto done;
    }
    if (m < 1) {
        rc = MVK_ERR_EMPTY;  // cpp_subsampling/wrapper.cpp:266-270
        goto done;
    }
    *out_total = m;
    HCHK(cudaMemcpy(out_lengths_host, dolen, sizeof(int) * nb, cudaMemcpyDeviceToHost));
    *oph = (float*)malloc(sizeof(float) * 3 * (size_t)m);
    HCHK(cudaMemcpy(*oph, dop, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToHost));
    if (fdim > 0 && ofh) {
        *ofh = (float*)malloc(sizeof(float) * (size_t)m * fdim);
        HCHK(cudaMemcpy(*ofh, dof, sizeof(float) * (size_t)m * fdim, cudaMemcpyDeviceToHost));
    }
    if (ldim > 0 && olh) {
        *olh = (int*)malloc(sizeof(int) * (size_t)m * ldim);
        HCHK(cudaMemcpy(*olh, dol, sizeof(int) * (size_t)m * ldim, cudaMemcpyDeviceToHost));
    }
done:
#undef HCHK
    cudaFree(dp);
    cudaFree(df);
    cudaFree(dop);
    cudaFree(dof);
    cudaFree(dl_);
    cudaFree(dlen);
    cudaFree(dol);
    cudaFree(dolen);
    cudaFree(dtot);
    cudaFree(ws);
    return rc;
}

}  // extern "C"
