/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported, linked or executed by the product.
 *
 * Plain-C restatement of the reference's geometric pre-processing path:
 *
 *   orc_batch_neighbors       <- KPConv-PyTorch/cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332
 *                                (batch_nanoflann_neighbors) and :125-208 (batch_ordered_neighbors);
 *                                distance metric nanoflann.hpp:432-440, strict '<' :249-252, :1361.
 *   orc_grid_subsample_batch  <- KPConv-PyTorch/cpp_wrappers/cpp_subsampling/grid_subsampling/
 *                                grid_subsampling.cpp:5-105 (one cloud) and :109-210 (batch),
 *                                accumulator grid_subsampling.h:10-80, min/max cloud.cpp:27-66.
 *   orc_stl_order             <- iteration order of libstdc++ std::unordered_map<size_t,...>
 *                                (GCC 13 bits/hashtable.h _M_insert_bucket_begin / _M_rehash_aux,
 *                                bits/hashtable_policy.h _Prime_rehash_policy::_M_need_rehash).
 *                                Third-party behaviour (toolchain-defined): pinned in
 *                                tests/test_oracle.py against the live std::unordered_map through
 *                                oracle/stl_probe.cpp, and against oracle/_ref (the compiled
 *                                reference) end to end.
 *
 * All floating point is fp32 with the reference's operation order; compile WITHOUT
 * -ffast-math / -march=native and with -ffp-contract=off (see oracle/Makefile) so that no FMA
 * contraction changes d2 or voxel-index bits.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* radius neighbours                                                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    float d2;
    int idx;
} nd_t;

static int nd_cmp(const void* a, const void* b) {
    const nd_t* x = (const nd_t*)a;
    const nd_t* y = (const nd_t*)b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* Brute force over the supports of the query's own batch element.
 * d2 = ((dx*dx) + dy*dy) + dz*dz in fp32 (nanoflann.hpp:432-440; cloud.h:71-74 for the
 * ordered variant gives the same expression), kept when d2 < r2 with r2 = radius*radius in fp32
 * (neighbors.cpp:226).  Rows are ordered by (d2 ascending, support index ascending): exactly
 * what batch_ordered_neighbors emits (stable upper_bound insertion, neighbors.cpp:176-181), and
 * what nanoflann emits up to permutations inside exact-d2 ties (std::sort, unstable).
 * Output: [nq, max_count] int32 of STACKED support indices, padded with ns (neighbors.cpp:322-324).
 * Returns max_count; *out is malloc'ed.  counts (optional, [nq]) receives per-row counts. */
int orc_batch_neighbors(const float* q, int nq, const float* s, int ns, const int* qb,
                        const int* sb, int nb, float radius, int** out, int* counts) {
    float r2 = radius * radius;
    nd_t** rows = (nd_t**)calloc((size_t)(nq > 0 ? nq : 1), sizeof(nd_t*));
    int* cnt = (int*)calloc((size_t)(nq > 0 ? nq : 1), sizeof(int));
    int max_count = 0;
    int q0 = 0, s0 = 0;
    for (int b = 0; b < nb; b++) {
        int nqb = qb[b], nsb = sb[b];
        nd_t* tmp = (nd_t*)malloc(sizeof(nd_t) * (size_t)(nsb > 0 ? nsb : 1));
        for (int i = q0; i < q0 + nqb && i < nq; i++) {
            float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            int c = 0;
            for (int j = s0; j < s0 + nsb && j < ns; j++) {
                float dx = qx - s[3 * j], dy = qy - s[3 * j + 1], dz = qz - s[3 * j + 2];
                float d2 = dx * dx;
                d2 = d2 + dy * dy;
                d2 = d2 + dz * dz;
                if (d2 < r2) {
                    tmp[c].d2 = d2;
                    tmp[c].idx = j;
                    c++;
                }
            }
            qsort(tmp, (size_t)c, sizeof(nd_t), nd_cmp);
            rows[i] = (nd_t*)malloc(sizeof(nd_t) * (size_t)(c > 0 ? c : 1));
            memcpy(rows[i], tmp, sizeof(nd_t) * (size_t)c);
            cnt[i] = c;
            if (c > max_count) max_count = c;
        }
        free(tmp);
        q0 += nqb;
        s0 += nsb;
    }
    int* o = (int*)malloc(sizeof(int) * (size_t)(nq > 0 ? nq : 1) * (size_t)(max_count > 0 ? max_count : 1));
    for (int i = 0; i < nq; i++) {
        for (int j = 0; j < max_count; j++) o[(size_t)i * max_count + j] = (rows[i] && j < cnt[i]) ? rows[i][j].idx : ns;
        if (counts) counts[i] = cnt[i];
        free(rows[i]);
    }
    free(rows);
    free(cnt);
    *out = o;
    return max_count;
}

/* ------------------------------------------------------------------------------------------ */
/* libstdc++ unordered_map iteration order                                                    */
/* ------------------------------------------------------------------------------------------ */

/* Bucket-count schedule of _Prime_rehash_policy with max_load_factor 1 and growth factor 2:
 * 1 -> 13 on the first insert, then "smallest prime in libstdc++'s table >= 2*B" when element
 * number B+1 arrives.  Filled by orc_set_schedule() from oracle/stl_probe.cpp (live libstdc++);
 * the default below is the GCC 13.3 table prefix observed in this image. */
static uint64_t g_sched[64] = {13ull,        29ull,        59ull,         127ull,       257ull,
                               541ull,       1109ull,      2357ull,       5087ull,      10273ull,
                               20753ull,     42043ull,     85229ull,      172933ull,    351061ull,
                               712697ull,    1447153ull,   2938679ull,    5967347ull,   12117689ull,
                               24607243ull,  49969847ull,  101473717ull,  206062531ull, 418451333ull,
                               849749479ull, 1725587117ull, 3504151727ull};
static int g_nsched = 28;

void orc_set_schedule(const uint64_t* s, int n) {
    if (n > 64) n = 64;
    memcpy(g_sched, s, sizeof(uint64_t) * (size_t)n);
    g_nsched = n;
}

/* Direct simulation of the node list.  keys[0..n) are DISTINCT hash values (identity hash of
 * the size_t key) in insertion order.  order[0..n) receives the insertion ranks in iteration
 * order (begin() first).
 *   insert into empty bucket  -> node becomes the global list head;
 *   insert into used bucket   -> node goes right after that bucket's "before" node, i.e. to the
 *                                head of the bucket's group;
 *   rehash                    -> re-insert all nodes in current list order with the same rule. */
void orc_stl_order(const uint64_t* keys, int n, int* order) {
    if (n <= 0) return;
    /* node i = i-th inserted key; node n = the before_begin sentinel */
    int* next = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    int SENT = n;
    next[SENT] = -1;
    uint64_t B = 1;
    int si = 0;
    int* before = (int*)malloc(sizeof(int) * 1);
    before[0] = -1;
    uint64_t next_resize = 0;
    for (int i = 0; i < n; i++) {
        if ((uint64_t)i + 1 > next_resize) {
            /* grow: table has i elements, one is being inserted */
            uint64_t nB = g_sched[si < g_nsched ? si : g_nsched - 1];
            si++;
            free(before);
            before = (int*)malloc(sizeof(int) * (size_t)nB);
            for (uint64_t k = 0; k < nB; k++) before[k] = -1;
            int p = next[SENT];
            next[SENT] = -1;
            uint64_t bbegin_bkt = 0;
            while (p >= 0) {
                int nx = next[p];
                uint64_t bkt = keys[p] % nB;
                if (before[bkt] < 0) {
                    next[p] = next[SENT];
                    next[SENT] = p;
                    before[bkt] = SENT;
                    if (next[p] >= 0) before[bbegin_bkt] = p;
                    bbegin_bkt = bkt;
                } else {
                    next[p] = next[before[bkt]];
                    next[before[bkt]] = p;
                }
                p = nx;
            }
            B = nB;
            next_resize = B; /* floor(B * max_load_factor), max_load_factor = 1 */
        }
        uint64_t bkt = keys[i] % B;
        if (before[bkt] >= 0) {
            next[i] = next[before[bkt]];
            next[before[bkt]] = i;
        } else {
            next[i] = next[SENT];
            next[SENT] = i;
            if (next[i] >= 0) before[keys[next[i]] % B] = i;
            before[bkt] = SENT;
        }
    }
    int k = 0;
    for (int p = next[SENT]; p >= 0; p = next[p]) order[k++] = p;
    free(next);
    free(before);
}

/* ------------------------------------------------------------------------------------------ */
/* grid subsampling                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t* keys; /* open addressing, UINT64_MAX = empty */
    int* vals;
    size_t cap;
} kmap_t;

static void kmap_init(kmap_t* m, size_t n) {
    size_t cap = 16;
    while (cap < 2 * n + 8) cap <<= 1;
    m->cap = cap;
    m->keys = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    m->vals = (int*)malloc(sizeof(int) * cap);
    for (size_t i = 0; i < cap; i++) m->keys[i] = UINT64_MAX;
}
static void kmap_free(kmap_t* m) {
    free(m->keys);
    free(m->vals);
}
static int* kmap_slot(kmap_t* m, uint64_t key, int* is_new) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    size_t i = (size_t)(h >> 20) & (m->cap - 1);
    while (m->keys[i] != UINT64_MAX && m->keys[i] != key) i = (i + 1) & (m->cap - 1);
    *is_new = (m->keys[i] == UINT64_MAX);
    m->keys[i] = key;
    return &m->vals[i];
}

typedef struct {
    int label, count;
} lc_t;
typedef struct {
    lc_t* e;
    int n, cap;
} hist_t;

/* One cloud.  Returns M; outputs (caller-allocated, capacity n) are filled in the reference's
 * emission order (= unordered_map iteration order). */
static int subsample_one(const float* p, int n, const float* f, int fdim, const int* c, int ldim,
                         float dl, float* op, float* of, int* oc) {
    if (n <= 0) return 0;
    /* cloud.cpp:27-66 */
    float mnx = p[0], mny = p[1], mnz = p[2], mxx = p[0], mxy = p[1];
    for (int i = 0; i < n; i++) {
        float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        if (x < mnx) mnx = x;
        if (y < mny) mny = y;
        if (z < mnz) mnz = z;
        if (x > mxx) mxx = x;
        if (y > mxy) mxy = y;
    }
    /* grid_subsampling.cpp:27  origin = floor(min * (1/dl)) * dl */
    float inv = 1 / dl;
    float ox = floorf(mnx * inv) * dl, oy = floorf(mny * inv) * dl, oz = floorf(mnz * inv) * dl;
    /* :30-31 */
    size_t NX = (size_t)floorf((mxx - ox) / dl) + 1;
    size_t NY = (size_t)floorf((mxy - oy) / dl) + 1;

    kmap_t map;
    kmap_init(&map, (size_t)n);
    uint64_t* vkey = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    int* cnt = (int*)calloc((size_t)n, sizeof(int));
    float* sum = (float*)calloc((size_t)n * 3, sizeof(float));
    float* fsum = fdim ? (float*)calloc((size_t)n * fdim, sizeof(float)) : NULL;
    hist_t* hist = ldim ? (hist_t*)calloc((size_t)n * ldim, sizeof(hist_t)) : NULL;
    int M = 0;
    for (int i = 0; i < n; i++) {
        /* :53-56  true fp32 divisions */
        size_t iX = (size_t)floorf((p[3 * i] - ox) / dl);
        size_t iY = (size_t)floorf((p[3 * i + 1] - oy) / dl);
        size_t iZ = (size_t)floorf((p[3 * i + 2] - oz) / dl);
        uint64_t key = (uint64_t)(iX + NX * iY + NX * NY * iZ);
        int is_new;
        int* slot = kmap_slot(&map, key, &is_new);
        if (is_new) {
            *slot = M;
            vkey[M] = key;
            M++;
        }
        int v = *slot;
        /* grid_subsampling.h:74-79: sequential fp32 accumulation in point order */
        cnt[v] += 1;
        sum[3 * v] += p[3 * i];
        sum[3 * v + 1] += p[3 * i + 1];
        sum[3 * v + 2] += p[3 * i + 2];
        for (int k = 0; k < fdim; k++) fsum[(size_t)v * fdim + k] += f[(size_t)i * fdim + k];
        for (int k = 0; k < ldim; k++) {
            hist_t* h = &hist[(size_t)v * ldim + k];
            int lab = c[(size_t)i * ldim + k], j;
            for (j = 0; j < h->n; j++)
                if (h->e[j].label == lab) break;
            if (j == h->n) {
                if (h->n == h->cap) {
                    h->cap = h->cap ? 2 * h->cap : 4;
                    h->e = (lc_t*)realloc(h->e, sizeof(lc_t) * (size_t)h->cap);
                }
                h->e[j].label = lab;
                h->e[j].count = 0;
                h->n++;
            }
            h->e[j].count++;
        }
    }
    int* order = (int*)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
    orc_stl_order(vkey, M, order);
    for (int k = 0; k < M; k++) {
        int v = order[k];
        /* :87  point * (1.0 / count): double reciprocal, converted to float, fp32 multiply */
        float w = (float)(1.0 / cnt[v]);
        op[3 * k] = sum[3 * v] * w;
        op[3 * k + 1] = sum[3 * v + 1] * w;
        op[3 * k + 2] = sum[3 * v + 2] * w;
        /* :88-96  f / (float)count */
        float fc = (float)cnt[v];
        for (int j = 0; j < fdim; j++) of[(size_t)k * fdim + j] = fsum[(size_t)v * fdim + j] / fc;
        /* :97-102 first maximum in unordered_map<int,int> iteration order; hash(int) is the
         * value converted to size_t */
        for (int j = 0; j < ldim; j++) {
            hist_t* h = &hist[(size_t)v * ldim + j];
            uint64_t* lk = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)h->n);
            int* lo = (int*)malloc(sizeof(int) * (size_t)h->n);
            for (int t = 0; t < h->n; t++) lk[t] = (uint64_t)(size_t)h->e[t].label;
            orc_stl_order(lk, h->n, lo);
            int best = lo[0];
            for (int t = 1; t < h->n; t++)
                if (h->e[best].count < h->e[lo[t]].count) best = lo[t];
            oc[(size_t)k * ldim + j] = h->e[best].label;
            free(lk);
            free(lo);
        }
    }
    if (hist) {
        for (size_t t = 0; t < (size_t)n * ldim; t++) free(hist[t].e);
        free(hist);
    }
    free(order);
    free(vkey);
    free(cnt);
    free(sum);
    free(fsum);
    kmap_free(&map);
    return M;
}

/* grid_subsampling.cpp:109-210.  max_p < 1 means "no limit" (:134-135); an element with more
 * than max_p voxels keeps the first max_p in emission order (:181-204).
 * Outputs are malloc'ed; returns the total number of subsampled points. */
int orc_grid_subsample_batch(const float* pts, int n, const float* feats, int fdim, const int* cls,
                             int ldim, const int* batches, int nb, float dl, int max_p,
                             float** o_pts, float** o_feats, int** o_cls, int** o_batches) {
    if (max_p < 1) max_p = n;
    size_t cap = (size_t)(n > 0 ? n : 1);
    float* op = (float*)malloc(sizeof(float) * 3 * cap);
    float* of = (float*)malloc(sizeof(float) * (size_t)(fdim > 0 ? fdim : 1) * cap);
    int* oc = (int*)malloc(sizeof(int) * (size_t)(ldim > 0 ? ldim : 1) * cap);
    int* ob = (int*)malloc(sizeof(int) * (size_t)(nb > 0 ? nb : 1));
    float* tp = (float*)malloc(sizeof(float) * 3 * cap);
    float* tf = (float*)malloc(sizeof(float) * (size_t)(fdim > 0 ? fdim : 1) * cap);
    int* tc = (int*)malloc(sizeof(int) * (size_t)(ldim > 0 ? ldim : 1) * cap);
    int s0 = 0, M = 0;
    for (int b = 0; b < nb; b++) {
        int nbp = batches[b];
        int m = subsample_one(pts + 3 * (size_t)s0, nbp, feats ? feats + (size_t)s0 * fdim : NULL,
                              feats ? fdim : 0, cls ? cls + (size_t)s0 * ldim : NULL,
                              cls ? ldim : 0, dl, tp, tf, tc);
        if (m > max_p) m = max_p;
        memcpy(op + 3 * (size_t)M, tp, sizeof(float) * 3 * (size_t)m);
        if (feats) memcpy(of + (size_t)M * fdim, tf, sizeof(float) * (size_t)fdim * m);
        if (cls) memcpy(oc + (size_t)M * ldim, tc, sizeof(int) * (size_t)ldim * m);
        ob[b] = m;
        M += m;
        s0 += nbp;
    }
    free(tp);
    free(tf);
    free(tc);
    *o_pts = op;
    *o_feats = of;
    *o_cls = oc;
    *o_batches = ob;
    return M;
}

void orc_free(void* p) { free(p); }
