// Shared helpers for the ddp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ddp_b200.h"

#define DDP_LAUNCH_CHECK()                                  \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

static inline int ddp_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

__device__ __forceinline__ int ddp_find_segment(const int32_t *__restrict__ ptr, int n_seg, int i) {
    // largest b with ptr[b] <= i  (ptr ascending, ptr[n_seg] = total)
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float ddp_sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    // torch_cluster kernels: dist += (x_d - y_d) * (x_d - y_d) for d = 0,1,2 with nvcc's default FMA contraction
    float t = ax - bx;
    float d = __fmaf_rn(t, t, 0.f);
    t = ay - by;
    d = __fmaf_rn(t, t, d);
    t = az - bz;
    d = __fmaf_rn(t, t, d);
    return d;
}
