// Strict fp32 contraction on the CUDA cores (FFMA, fp32 accumulate): the bit-for-bit "fp32"
// precision mode of KPConv's [nq, K*cin] x [K*cin, cout] product and of its two backward
// products, with arbitrary operand strides so that no transposed copies are needed.
//   D[m, n] (+)= sum_k A[m*a_rs + k*a_cs] * B[k*b_rs + n*b_cs]
// 64x64 output tile per CTA, 16-deep k slices staged through shared memory, 4x4 micro-tile per
// thread.  split_k > 1 spreads the k range over blockIdx.z and accumulates with fp32 atomics.
#include "common.cuh"

namespace mvk {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long a_rs, long long a_cs,
                const float* __restrict__ B, long long b_rs, long long b_cs, int M, int N, int K,
                float* __restrict__ D, int ldd, int k_per_split, int atomic_out) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    // loader mapping: pick the fast-varying thread index along the contiguous operand dimension
    const bool a_k_contig = (a_cs == 1);
    const bool b_n_contig = (b_cs == 1);
    for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
        for (int r = 0; r < (TM * TK) / 256; r++) {
            int e = tid + r * 256;
            int mm, kk;
            if (a_k_contig) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
            int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < kend) ? A[(long long)gm * a_rs + (long long)gk * a_cs] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < (TN * TK) / 256; r++) {
            int e = tid + r * 256;
            int nn, kk;
            if (b_n_contig) { nn = e % TN; kk = e / TN; } else { kk = e % TK; nn = e / TK; }
            int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < kend) ? B[(long long)gk * b_rs + (long long)gn * b_cs] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk++) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float* dst = D + (size_t)gm * ldd + gn;
            if (atomic_out) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
        }
    }
}

}  // namespace
}  // namespace mvk

using namespace mvk;

extern "C" int mvk_gemm_f32(const float* A, long long a_rs, long long a_cs, const float* B,
                            long long b_rs, long long b_cs, int M, int N, int K, float* D, int ldd,
                            int split_k, mvk_stream_t stream) {
    if (!A || !B || !D || M < 0 || N < 0 || K < 0 || ldd < N) return MVK_ERR_INVALID_ARG;
    if (M == 0 || N == 0) return MVK_OK;
    if (split_k < 1) split_k = 1;
    int kps = (K + split_k - 1) / split_k;
    kps = (kps + TK - 1) / TK * TK;
    if (kps < TK) kps = TK;
    int splits = K > 0 ? (K + kps - 1) / kps : 1;
    dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, splits);
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K, D,
                                                            ldd, kps, split_k > 1 ? 1 : 0);
    MVK_LAUNCHED("gemm_f32");
    return MVK_OK;
}
