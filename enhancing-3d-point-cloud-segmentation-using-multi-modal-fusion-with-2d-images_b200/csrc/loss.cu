// Segmentation loss of the KPFCNN head: softmax cross-entropy over [rows, classes] logits with an
// ignore label, mean over the valid rows.  Reference: KPConv-PyTorch/models/architectures.py:176-181,
// 352-373 (torch.nn.CrossEntropyLoss(ignore_index=-1) on the transposed logits).  Forward and backward
// are one streaming pass each (20 classes x 267 k points = 21 MB): one thread per row with an online
// softmax; the per-row log-sum-exp is kept for the backward pass.
#include "common.cuh"

namespace mvk {
namespace {

constexpr int TB = 256;

// loss_acc[0] += sum of per-row losses (fp64), count[0] += valid rows, count[1] = retirement ticket.
// The CTA that retires last writes loss_out = sum / max(count, 1) (NaN when no row is valid, like torch).
__global__ void __launch_bounds__(TB)
xent_fwd_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ labels, int rows, int classes,
                long long ignore_index, float* __restrict__ lse, double* __restrict__ loss_acc,
                unsigned int* __restrict__ count, float* __restrict__ loss_out) {
    pdl_enter();
    __shared__ double s_sum[TB / 32];
    __shared__ unsigned int s_cnt[TB / 32];
    double my = 0.0;
    unsigned int valid = 0;
    for (int r = blockIdx.x * TB + threadIdx.x; r < rows; r += gridDim.x * TB) {
        const float* x = logits + (size_t)r * ld;
        const long long lab = labels[r];
        float m = -INFINITY, s = 0.f, picked = 0.f;
        for (int c = 0; c < classes; c++) {
            const float v = x[c];
            if (v > m) {
                s = s * __expf(m - v);
                m = v;
            }
            s += __expf(v - m);
            if (c == lab) picked = v;
        }
        const float l = m + __logf(s);
        lse[r] = l;
        if (lab != ignore_index && lab >= 0 && lab < classes) {
            my += (double)(l - picked);
            valid++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my += __shfl_xor_sync(0xffffffffu, my, o);
        valid += __shfl_xor_sync(0xffffffffu, valid, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_sum[threadIdx.x >> 5] = my;
        s_cnt[threadIdx.x >> 5] = valid;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        unsigned int n = 0;
        for (int w = 0; w < TB / 32; w++) {
            t += s_sum[w];
            n += s_cnt[w];
        }
        if (n) {
            atomicAdd(loss_acc, t);
            atomicAdd(&count[0], n);
        }
        __threadfence();
        if (atomicAdd(&count[1], 1u) == gridDim.x - 1) {
            __threadfence();
            const double tot = __ldcg(loss_acc);
            const unsigned int nv = __ldcg(&count[0]);
            *loss_out = nv ? (float)(tot / (double)nv) : __int_as_float(0x7fc00000);
        }
    }
}

// grad[r, c] = (softmax(x)[r, c] - [c == label]) * upstream / valid   (0 for ignored rows)
__global__ void __launch_bounds__(TB)
xent_bwd_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ labels, int rows, int classes,
                long long ignore_index, const float* __restrict__ lse, const unsigned int* __restrict__ count,
                const float* __restrict__ upstream, float* __restrict__ grad, int ldg) {
    pdl_enter();
    const size_t total = (size_t)rows * classes;
    const unsigned int nv = count[0];
    const float scale = nv ? upstream[0] / (float)nv : 0.f;
    for (size_t t = (size_t)blockIdx.x * TB + threadIdx.x; t < total; t += (size_t)gridDim.x * TB) {
        const int r = (int)(t / classes), c = (int)(t % classes);
        const long long lab = labels[r];
        float g = 0.f;
        if (lab != ignore_index && lab >= 0 && lab < classes)
            g = (__expf(logits[(size_t)r * ld + c] - lse[r]) - (c == lab ? 1.f : 0.f)) * scale;
        grad[(size_t)r * ldg + c] = g;
    }
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" int mvk_softmax_xent(const float* logits, int ld, const long long* labels, int rows, int classes,
                                long long ignore_index, float* lse, double* loss_acc, unsigned int* count,
                                float* loss_out, mvk_stream_t stream) {
    if (!logits || !labels || !lse || !loss_acc || !count || !loss_out || rows < 1 || classes < 1 || ld < classes)
        return MVK_ERR_INVALID_ARG;
    int grid = (rows + TB - 1) / TB;
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    launch_pdl(xent_fwd_kernel, dim3(grid), dim3(TB), 0, (cudaStream_t)stream, 1, logits, ld, labels, rows, classes,
               ignore_index, lse, loss_acc, count, loss_out);
    MVK_LAUNCHED("xent_fwd_kernel");
    return MVK_OK;
}

extern "C" int mvk_softmax_xent_bwd(const float* logits, int ld, const long long* labels, int rows, int classes,
                                    long long ignore_index, const float* lse, const unsigned int* count,
                                    const float* upstream, float* grad, int ldg, mvk_stream_t stream) {
    if (!logits || !labels || !lse || !count || !upstream || !grad || rows < 1 || classes < 1 || ld < classes ||
        ldg < classes)
        return MVK_ERR_INVALID_ARG;
    size_t blocks = ((size_t)rows * classes + TB - 1) / TB;
    const size_t maxb = (size_t)num_sms() * 16;
    launch_pdl(xent_bwd_kernel, dim3((unsigned)(blocks < maxb ? blocks : maxb)), dim3(TB), 0, (cudaStream_t)stream, 1,
               logits, ld, labels, rows, classes, ignore_index, lse, count, upstream, grad, ldg);
    MVK_LAUNCHED("xent_bwd_kernel");
    return MVK_OK;
}
