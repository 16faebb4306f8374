// TEST INFRASTRUCTURE ONLY -- never imported by the product package.
//
// extern "C" shim around the UNMODIFIED reference C++ core, compiled in place from
// /root/reference (see oracle/Makefile; outputs go to oracle/_ref/, which is git-ignored).
// It reproduces the dtype/shape contract of the reference CPython glue, which no longer
// compiles against NumPy 2:
//   KPConv-PyTorch/cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238   (batch_query)
//   KPConv-PyTorch/cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333 (subsample_batch)
//   KPConv-PyTorch/cpp_wrappers/cpp_subsampling/wrapper.cpp:338-566 (subsample)
// i.e. float32 (N,3) points, int32 batch lengths, copy-in to std::vector, fresh int32 /
// float32 outputs.  No reference source is copied here: only the public function
// prototypes declared by the reference headers are called.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

namespace {
template <typename T>
T* dup(const std::vector<T>& v) {
    T* p = (T*)std::malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (v.size()) std::memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}
std::vector<PointXYZ> as_cloud(const float* p, int n) {
    // PointXYZ is three packed floats (cloud.h:40-105), the glue relies on the same cast
    // (cpp_neighbors/wrapper.cpp:188-189).
    return std::vector<PointXYZ>((const PointXYZ*)p, (const PointXYZ*)p + n);
}
}  // namespace

extern "C" {

void ref_free(void* p) { std::free(p); }

// which: 0 = batch_nanoflann_neighbors (what wrapper.cpp:198 calls),
//        1 = batch_ordered_neighbors   (the commented-out alternative, wrapper.cpp:197)
// Returns max_count (row width); *out is malloc'ed [nq, max_count] int32 (free with ref_free).
int ref_batch_neighbors(const float* q, int nq, const float* s, int ns, const int* qb,
                        const int* sb, int nb, float radius, int which, int** out) {
    std::vector<PointXYZ> queries = as_cloud(q, nq), supports = as_cloud(s, ns);
    std::vector<int> q_batches(qb, qb + nb), s_batches(sb, sb + nb), idx;
    if (which == 0)
        batch_nanoflann_neighbors(queries, supports, q_batches, s_batches, idx, radius);
    else
        batch_ordered_neighbors(queries, supports, q_batches, s_batches, idx, radius);
    *out = dup(idx);
    return nq > 0 ? (int)(idx.size() / (size_t)nq) : 0;
}

// subsample_batch contract.  features / classes may be NULL (fdim / ldim = 0).
// Returns the number of subsampled points M; outputs are malloc'ed.
int ref_grid_subsample_batch(const float* pts, int n, const float* feats, int fdim,
                             const int* cls, int ldim, const int* batches, int nb, float dl,
                             int max_p, float** o_pts, float** o_feats, int** o_cls,
                             int** o_batches) {
    std::vector<PointXYZ> op = as_cloud(pts, n), sp;
    std::vector<float> of, sf;
    std::vector<int> oc, sc, ob(batches, batches + nb), sb;
    if (feats && fdim > 0) of.assign(feats, feats + (size_t)n * fdim);
    if (cls && ldim > 0) oc.assign(cls, cls + (size_t)n * ldim);
    batch_grid_subsampling(op, sp, of, sf, oc, sc, ob, sb, dl, max_p);
    *o_pts = (float*)dup(sp);
    *o_feats = dup(sf);
    *o_cls = dup(sc);
    *o_batches = dup(sb);
    return (int)sp.size();
}

// subsample contract (single cloud).
int ref_grid_subsample(const float* pts, int n, const float* feats, int fdim, const int* cls,
                       int ldim, float dl, float** o_pts, float** o_feats, int** o_cls) {
    std::vector<PointXYZ> op = as_cloud(pts, n), sp;
    std::vector<float> of, sf;
    std::vector<int> oc, sc;
    if (feats && fdim > 0) of.assign(feats, feats + (size_t)n * fdim);
    if (cls && ldim > 0) oc.assign(cls, cls + (size_t)n * ldim);
    grid_subsampling(op, sp, of, sf, oc, sc, dl, 0);
    *o_pts = (float*)dup(sp);
    *o_feats = dup(sf);
    *o_cls = dup(sc);
    return (int)sp.size();
}

}  // extern "C"
