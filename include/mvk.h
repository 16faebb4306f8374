/* mvk.h -- C ABI of the B200-native MV-KPConv hot path (libmvk.so, sm_100a only).
 *
 * Conventions (all entry points):
 *   - plain C: raw pointers + sizes, no torch / C++ types in any signature;
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); the library
 *     never synchronises except inside the `*_host` convenience entry points;
 *   - scratch memory is caller-provided: ask `*_workspace_bytes`, pass `ws`/`ws_bytes`;
 *   - return value: 0 (MVK_OK) or a negative MVK_ERR_* code; nothing is thrown across the ABI.
 *     `mvk_error_string` names a code, `mvk_last_cuda_error` returns the last CUDA runtime error
 *     string seen by this thread.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).
 */
#ifndef MVK_H_
#define MVK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVK_OK 0
#define MVK_ERR_INVALID_ARG (-1)   /* bad shape / null pointer / unsupported parameter            */
#define MVK_ERR_WORKSPACE (-2)     /* ws_bytes smaller than *_workspace_bytes                      */
#define MVK_ERR_CUDA (-3)          /* a CUDA runtime / driver call failed (see mvk_last_cuda_error)*/
#define MVK_ERR_RANGE (-4)         /* coordinates / sizes outside what the hashed grid can index   */
#define MVK_ERR_UNSUPPORTED (-5)   /* mode of the reference operator not implemented here          */
#define MVK_ERR_EMPTY (-6)         /* empty result (the reference raises RuntimeError("Error"))    */

typedef void* mvk_stream_t; /* cudaStream_t */

const char* mvk_error_string(int code);
const char* mvk_last_cuda_error(void);
int mvk_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long mvk_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Radius neighbours.
 * Replaces cpp_wrappers.cpp_neighbors.radius_neighbors.batch_query
 *   KPConv-PyTorch/cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238  (glue, arg contract)
 *   KPConv-PyTorch/cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332 (batch_nanoflann_neighbors)
 * called through datasets/common.py:185-196 (batch_neighbors).
 *
 * queries [nq,3] f32, supports [ns,3] f32, q_lengths/s_lengths [nb] i32 (stacked batch elements),
 * radius f32.  A support j is a neighbour of query i iff both belong to the same batch element and
 * ((dx*dx)+dy*dy)+dz*dz < radius*radius evaluated in fp32 WITHOUT fma contraction.  Rows are sorted
 * by (d2, support index) ascending and padded with ns; indices are stacked (global) support indices.
 *
 * Two-phase protocol (the row width is data dependent):
 *   1. mvk_neighbors_count : builds the hashed cell grid in `ws`, writes counts[nq] and
 *                            *max_count (device int, must be zero-initialised by the callee: it is).
 *   2. caller reads max_count (the ONLY host sync), allocates out[nq, width], width >= 1.
 *   3. mvk_neighbors_fill  : re-uses the grid left in `ws` by phase 1 (same ws, same inputs);
 *      max_count is the value read in step 2 (sizes the per-warp sort buffers).
 *      width may be smaller than max_count: rows then keep their `width` NEAREST neighbours
 *      (== the reference's big_neighborhood_filter crop, datasets/common.py:411-421).
 * ---------------------------------------------------------------------------------------------- */
size_t mvk_neighbors_workspace_bytes(int nq, int ns, int nb);
int mvk_neighbors_count(const float* queries, int nq, const float* supports, int ns,
                        const int* q_lengths, const int* s_lengths, int nb, float radius, void* ws,
                        size_t ws_bytes, int* counts, int* max_count, mvk_stream_t stream);
int mvk_neighbors_fill(const float* queries, int nq, const float* supports, int ns,
                       const int* q_lengths, const int* s_lengths, int nb, float radius, void* ws,
                       size_t ws_bytes, int max_count, int width, int* out, mvk_stream_t stream);
/* Same with 64-bit output (the dtype the reference model consumes, datasets/common.py:874-876). */
int mvk_neighbors_fill_i64(const float* queries, int nq, const float* supports, int ns,
                           const int* q_lengths, const int* s_lengths, int nb, float radius,
                           void* ws, size_t ws_bytes, int max_count, int width, long long* out,
                           mvk_stream_t stream);
/* Single-pass variant for callers that crop the rows anyway (the pyramid builder: big_neighborhood_filter,
 * datasets/common.py:411-421): builds the grid and writes out[nq, width] = the `width` nearest
 * neighbours per row in one query pass, collecting up to list_cap hits per query in shared memory.
 * *max_count (device int) receives the true maximum hit count: if it exceeds list_cap the rows of
 * such queries may be wrong and the caller must fall back to the two-phase protocol; if it is
 * smaller than width the caller crops the columns (reference width = min(max_count, limit)).
 * counts may be NULL.  reuse_grid != 0 skips the grid build: `ws` must still hold the grid that the
 * previous call built for the SAME supports, s_lengths and radius (the pyramid queries one support
 * cloud three times per level: up-sampling of the finer level, convolution, pooling). */
int mvk_neighbors_query_capped(const float* queries, int nq, const float* supports, int ns, const int* q_lengths,
                               const int* s_lengths, int nb, float radius, void* ws, size_t ws_bytes, int width,
                               int list_cap, void* out, int out_is_i64, int* counts, int* max_count,
                               int reuse_grid, mvk_stream_t stream);
/* Host-buffer convenience entry point == the reference call: all pointers are HOST pointers,
 * *out_host is malloc'ed [nq, *width] int32 (free with mvk_free_host).  Copies H2D, runs both
 * phases on the default stream, copies D2H.  Returns MVK_ERR_EMPTY when nq * max_count == 0
 * (wrapper.cpp:201-205). */
int mvk_batch_neighbors_host(const float* queries_host, int nq, const float* supports_host, int ns,
                             const int* q_lengths_host, const int* s_lengths_host, int nb,
                             float radius, int** out_host, int* width);
void mvk_free_host(void* p);

/* ------------------------------------------------------------------------------------------------
 * Grid subsampling (voxel barycentres).
 * Replaces cpp_wrappers.cpp_subsampling.grid_subsampling.subsample / subsample_batch
 *   KPConv-PyTorch/cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333, 338-566 (glue)
 *   KPConv-PyTorch/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-210
 * called through datasets/common.py:44-182.
 *
 * points [n,3] f32, features [n,fdim] f32 or NULL, labels [n,ldim] i32 or NULL (ldim <= 1 in batch
 * mode, like the reference's usable range), lengths [nb] i32, sampleDl f32, max_p (<1 = no limit).
 * Outputs are sized for the worst case (n rows); out_lengths[nb] receives the per-element counts
 * (after max_p), *out_total (device int) their sum.  Output order is the reference's emission
 * order (libstdc++ unordered_map iteration order), bit-exact fp32 barycentres.
 * ---------------------------------------------------------------------------------------------- */
size_t mvk_subsample_workspace_bytes(int n, int nb, int fdim, int ldim);
int mvk_grid_subsample(const float* points, int n, const float* features, int fdim,
                       const int* labels, int ldim, const int* lengths, int nb, float sampleDl,
                       int max_p, void* ws, size_t ws_bytes, float* out_points,
                       float* out_features, int* out_labels, int* out_lengths, int* out_total,
                       mvk_stream_t stream);
/* Per-element rotation used by batch_grid_subsampling(random_grid_orient=True),
 * datasets/common.py:114-119 and :131-135:  out[i, c] = (p[i,0]*R[0,c] + p[i,1]*R[1,c]) + p[i,2]*R[2,c]
 * in fp32 without contraction (numpy's multiply-then-sum order); transpose != 0 applies R^T.
 * rot [nb, 9] f32 row-major.  In-place allowed (out == points). */
int mvk_rotate_batch(const float* points, int n, const int* lengths, int nb, const float* rot,
                     int transpose, float* out, mvk_stream_t stream);
/* Host-buffer convenience entry point (outputs malloc'ed; free with mvk_free_host). */
int mvk_grid_subsample_host(const float* points_host, int n, const float* features_host, int fdim,
                            const int* labels_host, int ldim, const int* lengths_host, int nb,
                            float sampleDl, int max_p, float** out_points_host,
                            float** out_features_host, int** out_labels_host,
                            int* out_lengths_host, int* out_total);

/* ------------------------------------------------------------------------------------------------
 * KPConv (rigid) forward / backward.
 * Replaces the ATen op chain of KPConv.forward, KPConv-PyTorch/models/blocks.py:277-374, and its
 * autograd.  out[i,:] = sum_k ( sum_h w_ihk x[j_ih,:] ) W[k]   with shadow index j == ns.
 *
 * Stage A  (gather + kernel-point influence, HBM/L2 bound):
 *      weighted[i, k*cin + c] = sum_h w_ihk * x[j_ih, c]
 *   written either as fp32 [nq, ld] (contraction "fp32") or as a bf16 hi/lo pair [nq, ld] each
 *   (hi = bf16(v), lo = bf16(v - hi)) for the tensor-core contraction.  ld >= 15*cin, pad zeroed.
 * Stage B  (contraction [nq, 15*cin] x [15*cin, cout]): mvk_gemm_* below.
 *
 * influence: 0 constant, 1 linear, 2 gaussian (blocks.py:329-346); aggregation: 0 sum, 1 closest
 * (:349-354).  idx_is_i64: neighb_inds dtype (the reference passes int64).
 * ---------------------------------------------------------------------------------------------- */
int mvk_kpconv_weighted(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                        int idx_is_i64, int h, const float* x, int cin, const float* kernel_points,
                        int num_kp, float kp_extent, int influence, int aggregation, int ld,
                        float* out_f32, void* out_hi_bf16, void* out_lo_bf16, mvk_stream_t stream);
/* Backward of stage A w.r.t. x:  grad_x[j, c] += sum_k w_ihk * grad_weighted[i, k*cin + c]
 * (grad_x [ns, cin] must be zero-initialised by the caller; fp32 atomics). */
/* mvk_kpconv_weighted for a row SHARED by several calls (K-concatenated operand of a layer whose input channels are
 * handled in parts, e.g. 66 = 64 on the fast path + 2 on the small-Cin path): this call writes columns [0, num_kp*cin) of
 * the rows starting at out_* (row pitch ld) and zero-fills up to `width`; nothing beyond `width` is touched. */
int mvk_kpconv_weighted_part(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                             int idx_is_i64, int h, const float* x, int cin, const float* kernel_points, int num_kp,
                             float kp_extent, int influence, int aggregation, int ld, int width, float* out_f32,
                             void* out_hi, void* out_lo, mvk_stream_t stream);
int mvk_kpconv_weighted_bwd(const float* q_pts, int nq, const float* s_pts, int ns,
                            const void* neighb_inds, int idx_is_i64, int h, int cin,
                            const float* kernel_points, int num_kp, float kp_extent, int influence,
                            int aggregation, const float* grad_weighted, int ld, float* grad_x,
                            mvk_stream_t stream);

/* Deformable KPConv, stage A (blocks.py:243-367): per-point kernel points deformed_kp [nq, K, 3]
 * (= kernel_points + offsets * KP_extent) and optional modulations [nq, K] (NULL = none), K <= 16.
 *      weighted[i, k*cin + c] = mod_ik * sum_h w_ihk * x[j_ih, c]
 *      min_d2[i, k] = min_h ||(s_j - q_i) - kp_ik||^2   (all slots, shadows at 1e6; argmin = its slot)
 * Neighbours out of range of every kernel point are dropped like the reference's compaction
 * (:300-325).  Backward: grad_x (+=, caller zeroes; may be NULL), grad_kp [nq, K, 3] and grad_mod
 * [nq, K] (written; may be NULL); grad_min_d2 (may be NULL) is routed to grad_kp through argmin. */
/* Deformable stage A supports num_kp <= 16 kernel points (one lane per kernel point holds its min_d2 / argmin; the
 * reference scripts use 15) and rows up to ~530 neighbours wide (shared-memory lists, warps per CTA adapt). */
int mvk_kpconv_deform_weighted(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                               int idx_is_i64, int h, const float* x, int cin, const float* deformed_kp,
                               const float* modulations, int num_kp, float kp_extent, int influence, int aggregation,
                               int ld, float* out_f32, void* out_hi_bf16, void* out_lo_bf16, float* min_d2,
                               int* argmin, mvk_stream_t stream);
int mvk_kpconv_deform_weighted_bwd(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                                   int idx_is_i64, int h, const float* x, int cin, const float* deformed_kp,
                                   const float* modulations, int num_kp, float kp_extent, int influence,
                                   int aggregation, const float* grad_weighted, int ld, const float* grad_min_d2,
                                   const int* argmin, float* grad_x, float* grad_kp, float* grad_mod,
                                   mvk_stream_t stream);

/* fp32 -> bf16 hi/lo split of a [rows, cols] matrix into zero-padded [rows_pad, ld] buffers. */
int mvk_split_bf16(const float* src, int rows, int cols, int src_ld, void* hi, void* lo, int rows_pad,
                   int ld, mvk_stream_t stream);

/* The same split for MANY matrices in one launch (all weight matrices of a model, once per optimiser
 * step).  table_dev: device array of n_tensors descriptors ordered by first_chunk; tensor i owns the chunk
 * range [first_chunk_i, first_chunk_i + ceil(rows_pad * dst_ld / 4096)) of the grid; dst_ld must be even and
 * hi / lo 4-byte aligned. */
typedef struct mvk_split_desc {
    const float* src;
    void* hi;
    void* lo;
    int rows, cols, src_ld, rows_pad, dst_ld, first_chunk;
} mvk_split_desc;
int mvk_split_bf16_multi(const mvk_split_desc* table_dev, int n_tensors, int total_chunks, mvk_stream_t stream);

/* Tensor-core contraction (tcgen05 + TMA, fp32 accumulation in TMEM):
 *      D[m, n] (+)= sum_k A(m,k) * B(k,n)          for m < M, n < n_valid
 * Operands are bf16 hi/lo pairs; terms = 3 evaluates hi*hi + lo*hi + hi*lo (fp32-grade, ~2^-16
 * relative), terms = 1 evaluates hi*hi only (plain bf16).
 *   a_mn_major = 0: A stored [M, lda] (K contiguous);   1: A stored [K, lda] (M contiguous)
 *   b_mn_major = 0: B stored [N, ldb] (K contiguous);   1: B stored [K, ldb] (N contiguous)
 * M, N, K are the true extents: tiles that reach past them are zero-filled by TMA (out-of-bounds
 * fill), so no operand padding is needed; lda/ldb must be multiples of 8 elements (16-byte row
 * pitch) and the base pointers 16-byte aligned.  n_valid <= N columns of D are written.
 * split_k > 1 partitions K over CTAs and accumulates into D with fp32 reductions (D must be zeroed
 * by the caller); split_k = 0 lets the library decide (it zeroes D itself when it splits). */
int mvk_gemm_bf16x3(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                    const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                    int n_valid, int terms, int split_k, mvk_stream_t stream);
/* Same contraction, additionally accumulating the column sums and sums of squares of D into
 * col_stats[0:n_valid] / col_stats[n_valid:2 n_valid] (fp64, zero-initialised by the caller) -- the
 * batch statistics of a following batch norm.  Unsplit single-column-tile problems fold this into
 * the epilogue (no extra pass over D); other shapes run mvk_col_stats afterwards. */
int mvk_gemm_bf16x3_stats(const void* a_hi, const void* a_lo, int a_mn_major, int lda, const void* b_hi,
                          const void* b_lo, int b_mn_major, int ldb, int M, int N, int K, float* D, int ldd,
                          int n_valid, int terms, int split_k, double* col_stats, mvk_stream_t stream);
/* Strict fp32 SIMT contraction with arbitrary strides:
 *      D[m, n] (+)= sum_k A[m*a_rs + k*a_cs] * B[k*b_rs + n*b_cs]                                */
int mvk_gemm_f32(const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs,
                 long long b_cs, int M, int N, int K, float* D, int ldd, int split_k,
                 mvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Point-wise blocks either side of KPConv.
 * Replaces the ATen chains of BatchNormBlock / UnaryBlock / the ResnetBottleneckBlock tail,
 * KPConv-PyTorch/models/blocks.py:430-466, :469-504, :637-649:
 *      UnaryBlock      z = leaky( bn( x W^T ) )          (Linear without bias on mvk_gemm_bf16x3)
 *      block tail      z = leaky( bn(y) + shortcut )
 * All matrices are [rows, cols] fp32 row-major with explicit leading dimensions.
 *
 * mvk_col_stats       stats[c] += sum_r y[r,c], stats[cols+c] += sum_r y[r,c]^2  (fp64, caller zeroes)
 * mvk_bn_finalize     training: batch mean / biased variance -> scale = gamma*invstd,
 *                     shift = beta - mean*scale, running stats updated like torch.nn.BatchNorm1d
 *                     (momentum, unbiased variance); eval: scale/shift from the running stats.
 * mvk_scale_shift_act out = leaky(y*scale + shift [+ residual]); scale NULL = 1 (bias-only block);
 *                     optional bf16 hi/lo copy (the operand format of mvk_gemm_bf16x3).
 * Column reductions (mvk_col_stats, mvk_bn_batch_stats, mvk_act_bwd_reduce) run as thread-block clusters:
 * the CTAs of a cluster fold their partial sums through distributed shared memory and issue one fp64
 * atomic per column and cluster.
 * mvk_act_bwd_reduce  d = dz * leaky'(pre); sums[c] += d, sums[cols+c] += d * xhat   (fp64)
 * mvk_act_bwd_apply   dy = scale*(d - sum_d/rows - xhat*sum_dxhat/rows)  (batch_stats != 0)
 *                     dy = scale*d                                         (otherwise)
 *                     written as fp32 and/or bf16 hi/lo; dres = d (gradient of the residual);
 *                     dgamma = sums[cols:], dbeta = sums[:cols].
 * ---------------------------------------------------------------------------------------------- */
int mvk_col_stats(const float* y, int rows, int cols, int ld, double* stats, mvk_stream_t stream);
/* mvk_col_stats + mvk_bn_finalize(training) in ONE launch: the CTA that retires last converts the
 * sums.  stats must hold 2*cols + 1 zero-initialised doubles (the extra slot is the retirement ticket).
 * num_batches_tracked (device int64, may be NULL) is incremented like torch.nn.BatchNorm1d does. */
int mvk_bn_batch_stats(const float* y, int rows, int cols, int ld, double* stats, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* mean_out, float* invstd_out, long long* num_batches_tracked,
                       mvk_stream_t stream);
int mvk_bn_finalize(const double* stats, int rows, int cols, const float* gamma, const float* beta, float eps,
                    float momentum, int training, float* running_mean, float* running_var, float* scale,
                    float* shift, float* mean_out, float* invstd_out, long long* num_batches_tracked /* may be NULL */,
                    mvk_stream_t stream);
int mvk_scale_shift_act(const float* y, int rows, int cols, int ld, const float* scale, const float* shift,
                        const float* residual, int ldr, float slope, float* out, int ldo, void* out_hi_bf16,
                        void* out_lo_bf16, int ldh, mvk_stream_t stream);
int mvk_act_bwd_reduce(const float* dz, int lddz, const float* y, int rows, int cols, int ld, const float* scale,
                       const float* shift, const float* residual, int ldr, const float* mean, const float* invstd,
                       float slope, double* sums, mvk_stream_t stream);
int mvk_act_bwd_apply(const float* dz, int lddz, const float* y, int rows, int cols, int ld, const float* scale,
                      const float* shift, const float* residual, int ldr, const float* mean, const float* invstd,
                      float slope, const double* sums, int batch_stats, float* dy, int lddy, void* dy_hi_bf16,
                      void* dy_lo_bf16, int ldh, float* dres, int lddres, float* dgamma, float* dbeta,
                      mvk_stream_t stream);

/* Segmentation loss of the KPFCNN head (architectures.py:176-181, 352-373: CrossEntropyLoss(ignore_index=-1)
 * on [rows, classes] logits, mean over the valid rows).  Forward: lse [rows] receives the per-row
 * log-sum-exp (kept for the backward), loss_acc (1 double) and count (2 uints: valid rows, retirement
 * ticket) must be zero-initialised, loss_out (1 float, device) receives the mean (NaN if no row is valid).
 * Backward: grad [rows, ldg] = (softmax - onehot) * upstream[0] / valid, zero rows for ignored labels;
 * upstream is a DEVICE scalar (no host read-back). */
int mvk_softmax_xent(const float* logits, int ld, const long long* labels, int rows, int classes,
                     long long ignore_index, float* lse, double* loss_acc, unsigned int* count, float* loss_out,
                     mvk_stream_t stream);
int mvk_softmax_xent_bwd(const float* logits, int ld, const long long* labels, int rows, int classes,
                         long long ignore_index, const float* lse, const unsigned int* count, const float* upstream,
                         float* grad, int ldg, mvk_stream_t stream);

/* Decoder entry (architectures.py:300-306): cat([closest_pool(x_coarse, inds), skip], dim=1) emitted directly as the
 * bf16 hi/lo operand [nq, ldh] of the following unary block's Linear (the fp32 concatenation is never
 * materialised).  inds [nq, h]: only column 0 is used; index ns (shadow) gathers zeros.  c1, c2, lds, ldh
 * multiples of 4. */
/* y[r, 0:c] += z[inds[r, 0], 0:c] for r < nq (y row pitch ldy; rows whose first index is outside [0, ns) are left
 * alone).  The gathered half of the decoder step contracted at the coarse level:
 * cat([closest_pool(x, up), skip], 1) W^T = closest_pool(x W_up^T, up) + skip W_skip^T  (architectures.py:300-306). */
int mvk_gather_add_rows(float* y, int ldy, int nq, int c, const float* z, int ns, const void* inds, int idx_is_i64,
                        int h, mvk_stream_t stream);
int mvk_upsample_concat_split(const float* x_coarse, int ns, int c1, const void* inds, int idx_is_i64, int nq, int h,
                              const float* skip, int lds, int c2, void* hi_bf16, void* lo_bf16, int ldh,
                              mvk_stream_t stream);

/* Gather pools on the neighbour matrices (blocks.py:79-110): mode 0 = max_pool (zero-padded
 * shadow row!), 1 = closest_pool (first column).  arg_out [nq, c] i32 (max_pool only) records the
 * winning support row for the backward.  Backward: grad_x[arg] += grad_out. */
int mvk_pool(const float* x, int ns, int c, const void* inds, int idx_is_i64, int nq, int h, int mode,
             float* out, int* arg_out, mvk_stream_t stream);
int mvk_pool_bwd(const float* grad_out, int ldg /* row pitch of grad_out, >= c */, int nq, int c, const int* arg,
                 const void* inds, int idx_is_i64, int h, int mode, int ns, float* grad_x, mvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 2D -> 3D lifting.
 *   mvk_unproject_views : KPConv-PyTorch/datasets/ScanNet_sphere_color.py:66-72 (depth2xyz) and
 *                         :409-417 (valid mask, camera->world), fp64 arithmetic like numpy.
 *   mvk_knn_pixels      : :442-452 (sklearn NearestNeighbors(k, 'ball_tree') on the valid pixels,
 *                         remapped to flat pixel ids view*h*w + pix), fp64 distances.
 *   mvk_group_points    : mvpnet/ops/group_points.py:5-31, mvpnet/ops/cuda/group_points_kernel.cu:25-144.
 *   mvk_feature_aggregation_* : mvpnet/models/mvpnet_3d.py:40-64 (+ common/nn/modules/mlp.py:38-75).
 * ---------------------------------------------------------------------------------------------- */
/* kinv [9] f64 row-major (= inv(cam_matrix[:3,:3]) computed by the host mirror exactly like the
 * reference, in fp32 then widened), depth [nv,h,w] f32, pose [nv,16] f32 row-major 4x4.
 * Outputs: xyz64 [nv*h*w,3] f64 (what the reference feeds to the kNN), xyz32 [nv*h*w,3] f32
 * (what the reference stacks into the batch), mask [nv*h*w] u8. */
int mvk_unproject_views(const double* kinv, const float* depth, const float* pose, int nv, int h,
                        int w, double* xyz64, float* xyz32, unsigned char* mask, mvk_stream_t stream);
size_t mvk_knn_workspace_bytes(int npix, int nq);
/* keys: xyz64 [npix,3] + mask [npix]; queries [nq,3] f32; out [nq,k] i64 flat pixel ids sorted by
 * (distance, pixel id); k <= 8. */
int mvk_knn_pixels(const double* xyz64, const unsigned char* mask, int npix, const float* queries,
                   int nq, int k, void* ws, size_t ws_bytes, long long* out, mvk_stream_t stream);
/* points [b,c,n1] f32, index [b,n2,k] i64 -> out [b,c,n2,k]. */
int mvk_group_points(const float* points, int b, int c, int n1, const long long* index, int n2, int k,
                     float* out, mvk_stream_t stream);
int mvk_group_points_bwd(const float* grad_out, int b, int c, int n1, const long long* index, int n2,
                         int k, float* grad_points, mvk_stream_t stream);
/* One SharedMLP layer over rows:  Y[r, o] = sum_c act(X[r, c]) * W[o, c]   where act applies the
 * PREVIOUS layer's folded batch-norm + ReLU (scale/shift [cin], NULL = identity), and accumulates
 * per-channel sum / sum-of-squares of Y into stats[2*cout] (f64, zero-initialised) for the
 * batch statistics of THIS layer.  rows = np*k. */
int mvk_fa_layer(const float* X, int rows, int cin, int ldx, const float* in_scale,
                 const float* in_shift, const float* W, int cout, float* Y, double* stats,
                 mvk_stream_t stream);
/* Builds the first-layer input rows [np*k, cin+4] = [feature, diff_xyz, dist] (mvpnet_3d.py:53-57)
 * directly from the 2D feature map: feat2d either channel-major [c, npix] (chan_stride = npix,
 * pix_stride = 1: the reference layout) or pixel-major; src xyz taken from xyz32 [npix,3]. */
int mvk_fa_gather(const float* feat2d, long long chan_stride, long long pix_stride, int c,
                  const float* xyz32, const long long* knn, int np, int k, const float* tgt_xyz,
                  float* X, int ldx, mvk_stream_t stream);
/* Final: out[o, p] = reduce_k relu(bn(Y[p*k + kk, o]))   (reduction 0 = sum, 1 = max), written
 * channel-major [cout, np] like the reference output (b=1). */
int mvk_fa_reduce(const float* Y, int np, int k, int cout, const float* scale, const float* shift,
                  int reduction, float* out, mvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Batched lifting: all spheres of a stacked batch in one launch set.  Replaces the per-sphere Python loop of
 * the fusion nets (KPConv-PyTorch/models/architectures_sphere.py:246-279, same loop in
 * architectures_sphere_middle_fusion.py and architectures_sphere_late_fusion.py) and the per-sphere numpy /
 * sklearn work of get_rgbd_data (datasets/ScanNet_sphere_color.py:409-452).
 * ---------------------------------------------------------------------------------------------- */
/* Like mvk_unproject_views with ONE inverse intrinsics matrix PER VIEW: kinv [nviews, 9] f64. */
int mvk_unproject_views_batched(const double* kinv, const float* depth, const float* pose, int nviews, int h, int w,
                                double* xyz64, float* xyz32, unsigned char* mask, mvk_stream_t stream);
size_t mvk_knn_batched_workspace_bytes(int nb, int npix_per_element, int nq, int grid_dim);
/* k nearest valid pixels of every query among the pixels of ITS batch element: xyz64 / xyz32 / mask are
 * [nb * npix_per_element, ...] (element b owns pixels [b*npix, (b+1)*npix)), queries [nq,3] f32 stacked with
 * q_lengths [nb] i32 (device).  Per-element uniform grid of grid_dim^3 cells (cell edge >= cell_min) + exact
 * ring search + exhaustive pass for unresolved queries; fp64 distances, ties by lower pixel id: the same result
 * as mvk_knn_pixels run per element.  out [nq,k] i64: pixel ids local to the element (view*h*w + pix, the
 * reference's knn_list[i]) or, with global_ids != 0, offset by b*npix (what mvk_fa_gather_views consumes);
 * -1 where the element has fewer than k valid pixels.  far_counts [nb] i32 (device, may be NULL): queries per
 * element that needed the exhaustive pass. */
int mvk_knn_pixels_batched(const double* xyz64, const float* xyz32, const unsigned char* mask, int nb,
                           int npix_per_element, const float* queries, const int* q_lengths, int nq, int k,
                           int grid_dim, float cell_min, int global_ids, void* ws, size_t ws_bytes, long long* out,
                           int* far_counts, mvk_stream_t stream);
/* mvk_fa_gather for a whole batch: feat = the 2D network's output [views, c, h*w] addressed through
 * (view_stride, chan_stride, pix_stride) in elements (NCHW or channels-last), knn_global [np,k] global pixel
 * ids (view * hw + pix), xyz32 [views*hw, 3], tgt_xyz [np,3] the stacked sphere points. */
int mvk_fa_gather_views(const float* feat, long long view_stride, long long chan_stride, long long pix_stride, int c,
                        int hw, const float* xyz32, const long long* knn_global, int np, int k, const float* tgt_xyz,
                        float* X, int ldx, mvk_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused KPConv forward (blocks.py:277-374, rigid / 'linear' / 'sum'): stage A + contraction in ONE kernel, the
 * weighted operand [nq, K*cin] staying in shared memory ("0 if fused").
 *   out [nq, cout] f32 = sum_k (sum_h w_ihk x[j_ih]) @ W[k]   with W given as its bf16 hi/lo pair [K*cin, ldw]
 *   (rows k*cin + c, row pitch ldw elements; what mvk_split_bf16 writes), bf16x3 arithmetic on tcgen05.
 *   a_hi / a_lo (both or neither; [nq, ld] bf16): when given, the weighted operand is ALSO written out once for
 *   the backward pass (dW = A^T dOut); the forward never reads it back.
 * Supported: cin in {32, 64, 128}, cout in {32, 64, 128}, num_kp <= 15, h <= 48 (mvk_kpconv_fused_supported);
 * everything else returns MVK_ERR_UNSUPPORTED and the caller runs mvk_kpconv_weighted + mvk_gemm_bf16x3.
 * ---------------------------------------------------------------------------------------------- */
int mvk_kpconv_fused_supported(int cin, int cout, int num_kp, int h, int influence, int aggregation);
int mvk_kpconv_fused(const float* q_pts, int nq, const float* s_pts, int ns, const void* neighb_inds,
                     int idx_is_i64, int h, const float* x, int cin, const float* kernel_points, int num_kp,
                     float kp_extent, int cout, const void* w_hi, const void* w_lo, int ldw, float* out,
                     void* a_hi, void* a_lo, int ld, mvk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVK_H_ */
