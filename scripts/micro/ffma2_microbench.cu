// FFMA2 (fma.rn.f32x2) latency / issue rate on sm_100a against scalar FFMA, one warp per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/ffma2_microbench scripts/micro/ffma2_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2f(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ffma2(u64 &d, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ void ffma1(float &d, float a, float b) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b)); }

template <int CHAINS, int MODE>   // MODE 0: FFMA, 1: FFMA2 pair x pair, 2: FFMA2 broadcast a
__global__ void k(float *out, long long *cyc, float a0, float b0) {
    u64 acc2[CHAINS];
    float acc1[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { acc2[i] = pack2f(i, i + 1); acc1[i] = i; }
    const u64 a2 = MODE == 2 ? pack2f(a0, a0) : pack2f(a0, a0 + 1.f), b2 = pack2f(b0, b0 + 2.f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (MODE == 0) ffma1(acc1[i], a0, b0);
                else ffma2(acc2[i], a2, b2);
            }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += acc1[i] + (float)(acc2[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS, int MODE>
void run(const char *name, int threads, float *out, long long *cyc) {
    k<CHAINS, MODE><<<1, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = 256.0 * 8 * CHAINS;
    printf("%-28s chains=%2d warps/SMSP=%d  %7.2f cycles per instruction per warp  (%5.2f warp-instr/cycle/SMSP)\n", name, CHAINS, threads / 128, c / n,
           n * (threads / 128) / c);
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    for (int threads : {128, 256, 512}) {
        run<1, 0>("FFMA dependent", threads, out, cyc);
        run<1, 1>("FFMA2 dependent", threads, out, cyc);
        run<1, 2>("FFMA2 bcast dependent", threads, out, cyc);
        run<5, 1>("FFMA2", threads, out, cyc);
        run<5, 2>("FFMA2 bcast", threads, out, cyc);
        run<8, 0>("FFMA", threads, out, cyc);
        run<8, 1>("FFMA2", threads, out, cyc);
        run<8, 2>("FFMA2 bcast", threads, out, cyc);
        run<16, 0>("FFMA", threads, out, cyc);
        run<16, 1>("FFMA2", threads, out, cyc);
    }
    return 0;
}
