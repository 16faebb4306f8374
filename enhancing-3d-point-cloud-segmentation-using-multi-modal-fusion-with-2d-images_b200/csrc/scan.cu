// Exclusive prefix sum of int32 (three-kernel reduce / scan-of-sums / downsweep).
#include <stdlib.h>

#include "common.cuh"

namespace mvk {

thread_local char g_last_cuda_error[256] = "";
unsigned long long g_launches = 0;

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 ints per block

// pass 1: per-tile sums.  pass 3: per-tile scan + offset.
template <bool WRITE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int* __restrict__ in,
                                                           int* __restrict__ out, int n,
                                                           int* __restrict__ tile_sums) {
    __shared__ int total;
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int ex = block_exclusive_scan(s, &total);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
    } else {
        int off = tile_sums[blockIdx.x] + ex;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = off;
            off += v[k];
        }
    }
}

// pass 2: single block scans the tile sums in place (exclusive), writes grand total.
__global__ void __launch_bounds__(1024) scan_sums(int* __restrict__ sums, int m,
                                                 int* __restrict__ total_out) {
    __shared__ int total;
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < m ? sums[i] : 0;
        int ex = block_exclusive_scan(v, &total);
        int c = carry_s;
        if (i < m) sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

size_t scan_tmp_ints(int n) { return (size_t)(n / SCAN_TILE + 2); }

int exclusive_scan_i32(const int* in, int* out, int n, int* total_out, int* tmp,
                       cudaStream_t stream) {
    if (n <= 0) {
        if (total_out) MVK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int), stream));
        return MVK_OK;
    }
    int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tiles<false><<<tiles, SCAN_THREADS, 0, stream>>>(in, nullptr, n, tmp);
    MVK_LAUNCHED("scan_tiles<0>");
    scan_sums<<<1, 1024, 0, stream>>>(tmp, tiles, total_out);
    MVK_LAUNCHED("scan_sums");
    scan_tiles<true><<<tiles, SCAN_THREADS, 0, stream>>>(in, out, n, tmp);
    MVK_LAUNCHED("scan_tiles<1>");
    return MVK_OK;
}

}  // namespace mvk

extern "C" {
const char* mvk_error_string(int code) {
    switch (code) {
        case MVK_OK: return "ok";
        case MVK_ERR_INVALID_ARG: return "invalid argument";
        case MVK_ERR_WORKSPACE: return "workspace too small";
        case MVK_ERR_CUDA: return "CUDA error";
        case MVK_ERR_RANGE: return "coordinates outside the indexable grid range";
        case MVK_ERR_UNSUPPORTED: return "unsupported mode";
        case MVK_ERR_EMPTY: return "Error";  // the reference's message for an empty result
        default: return "unknown error";
    }
}
const char* mvk_last_cuda_error(void) { return mvk::g_last_cuda_error; }
int mvk_version(void) { return 100; }
unsigned long long mvk_launch_count(void) { return mvk::g_launches; }
void mvk_free_host(void* p) { free(p); }
}
