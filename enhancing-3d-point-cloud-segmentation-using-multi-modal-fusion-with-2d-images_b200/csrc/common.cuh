// Shared helpers for libmvk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mvk.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmvk is written for sm_100a (B200) only"
#endif

namespace mvk {

extern thread_local char g_last_cuda_error[256];
extern unsigned long long g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
    return MVK_ERR_CUDA;
}

#define MVK_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return ::mvk::cuda_fail(e__, #call);     \
    } while (0)

// Count + check a kernel launch (no sync).
#define MVK_LAUNCHED(name)                                               \
    do {                                                                 \
        __atomic_add_fetch(&::mvk::g_launches, 1ull, __ATOMIC_RELAXED);  \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return ::mvk::cuda_fail(e__, name);      \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Programmatic dependent launch.  The step is ~760 short launches back to back; a kernel launched through
// launch_pdl may be scheduled while its predecessor in the stream is still draining, runs its prologue,
// and blocks in pdl_enter() until the predecessor has completed and its writes are visible.  pdl_enter()
// must therefore precede the first global-memory access of EVERY thread of a kernel launched this way
// (it is a no-op under a plain launch).  It releases the next launch only after its own wait, so at most
// one generation of CTAs is ever parked on the SMs.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MVK_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
// cluster > 1: (cluster, 1, 1) thread-block clusters.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Bump allocator over the caller's workspace.
struct Arena {
    char* base;
    size_t off, cap;
    Arena(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = (T*)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Exclusive scan of int32 (scan.cu).  tmp must hold scan_tmp_ints(n) ints.
size_t scan_tmp_ints(int n);
int exclusive_scan_i32(const int* in, int* out, int n, int* total_out /*device, may be null*/,
                       int* tmp, cudaStream_t stream);

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// *total (shared or global) receives the block sum.  Ends with a __syncthreads().
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = blockDim.x >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
        int sinc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, sinc, o);
            if (lane >= o) sinc += t;
        }
        warp_sums[lane] = sinc - s;  // exclusive warp offsets
        if (lane == 31) *total = sinc;
    }
    __syncthreads();
    int r = warp_sums[wid] + inc - v;
    __syncthreads();
    return r;
}

// Batch element of stacked point i given lengths[nb] (nb small): linear walk.
__device__ __forceinline__ int batch_of(const int* __restrict__ starts, int nb, int i) {
    // starts[b] = exclusive prefix, starts[nb] = total
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (starts[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

}  // namespace mvk
