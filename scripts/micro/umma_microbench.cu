// Microbenchmark of tcgen05.mma issue/throughput on sm_100a (developer tool, not part of the product path):
// cycles per MMA (M=128, K=16, bf16) as a function of N, with one or two alternating accumulators, A operand from
// shared memory (SS) or tensor memory (TS); plus a numerical check that the TS layout assumption (lane = row,
// column c = packed bf16 pair k=2c,2c+1) reproduces the SS result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench umma_microbench.cu && ./umma_microbench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr uint32_t DESC_HI = 8u | (1u << 14);
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// mode bit0: alternate between two accumulators; bit1: A from TMEM
__global__ void __launch_bounds__(128, 1) bench_kernel(int n, int mode, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *a = smem;                       // [2 k-chunks][128 rows][16 B] x 4 K-steps = 16 KB
    uint8_t *b = smem + 16384;               // [8 chunks][256 rows][16 B] = 32 KB
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(n);
        const uint32_t a_lo = ((smem_u32(a) >> 4) & 0x3FFFu) | (128u << 16);
        const uint32_t b_lo = ((smem_u32(b) >> 4) & 0x3FFFu) | ((uint32_t)n << 16);
        const bool alt = mode & 1, ts = mode & 2;
        // accumulators at columns 0 and 256 (n <= 240 when ts: A lives at columns 496..511)
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tb + ((alt && (i & 1)) ? 256u : 0u);
            const int k = i & 3;
            if (ts) mma_ts(d, tb + 480u + (uint32_t)k * 8u, b_lo + (uint32_t)k * 2u * n, idesc, i > 1);
            else mma_ss(d, a_lo + (uint32_t)k * 256u, b_lo + (uint32_t)k * 2u * n, idesc, i > 1);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// Numerical check SS vs TS: A[r][k] = ((r * 7 + k * 3) % 13 - 6) / 8, B[n][k] = ((n * 5 + k) % 11 - 5) / 4, K = 64, N = 64.
__global__ void __launch_bounds__(128, 1) check_kernel(float *out_ss, float *out_ts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __nv_bfloat16 *a = reinterpret_cast<__nv_bfloat16 *>(smem);            // core-matrix layout, K = 64: 8 chunks x 128 rows x 8
    __nv_bfloat16 *b = reinterpret_cast<__nv_bfloat16 *>(smem + 16384);    // 8 chunks x 64 rows x 8
    const int N = 64, K = 64;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        a[((k / 8) * 128 + r) * 8 + (k % 8)] = __float2bfloat16(((r * 7 + k * 3) % 13 - 6) / 8.f);
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        b[((k / 8) * N + n) * 8 + (k % 8)] = __float2bfloat16(((n * 5 + k) % 11 - 5) / 4.f);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const int r = threadIdx.x;
    const uint32_t lane_base = (uint32_t)((r / 32) * 32) << 16;
    // A into TMEM columns 256..287: column c holds (k = 2c, 2c + 1) of this thread's row
    {
        uint32_t v[32];
        for (int c = 0; c < 32; ++c) {
            __nv_bfloat162 t = __floats2bfloat162_rn(((r * 7 + (2 * c) * 3) % 13 - 6) / 8.f, ((r * 7 + (2 * c + 1) * 3) % 13 - 6) / 8.f);
            v[c] = *reinterpret_cast<uint32_t *>(&t);
        }
        for (int c8 = 0; c8 < 4; ++c8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + lane_base + 256u + c8 * 8u),
                         "r"(v[c8 * 8]), "r"(v[c8 * 8 + 1]), "r"(v[c8 * 8 + 2]), "r"(v[c8 * 8 + 3]), "r"(v[c8 * 8 + 4]), "r"(v[c8 * 8 + 5]),
                         "r"(v[c8 * 8 + 6]), "r"(v[c8 * 8 + 7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(N);
        const uint32_t a_lo = ((smem_u32(a) >> 4) & 0x3FFFu) | (128u << 16);
        const uint32_t b_lo = ((smem_u32(b) >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
        for (int k = 0; k < K / 16; ++k) {
            mma_ss(tb + 0u, a_lo + (uint32_t)k * 256u, b_lo + (uint32_t)k * 2u * N, idesc, k > 0);
            mma_ts(tb + 64u, tb + 256u + (uint32_t)k * 8u, b_lo + (uint32_t)k * 2u * N, idesc, k > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
        uint32_t x, y;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(x) : "r"(tb + lane_base + (uint32_t)c) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(y) : "r"(tb + lane_base + 64u + (uint32_t)c) : "memory");
        out_ss[r * N + c] = __uint_as_float(x);
        out_ts[r * N + c] = __uint_as_float(y);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// ---- 128-byte-swizzled K-major operands (bit 0 of sw: A, bit 1: B) instead of the no-swizzle core-matrix layout ----
// descriptor high word: SBO (8 rows x 128 B = 1024 B) >> 4 | version 1 << 14 | layout SWIZZLE_128B (2) << 29
constexpr uint32_t DESC_HI_SW128 = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ void mma_ss2(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(a_hi), "r"(b_hi) : "memory");
}
__global__ void __launch_bounds__(128, 1) bench_sw_kernel(int n, int sw, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *a = smem;                       // 16 KB: 128 rows x 128 B (K = 64)
    uint8_t *b = smem + 16384;               // 32 KB: 256 rows x 128 B
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(n);
        const bool swa = sw & 1, swb = sw & 2;
        const uint32_t a_lo = ((smem_u32(a) >> 4) & 0x3FFFu) | (swa ? (1u << 16) : (128u << 16));
        const uint32_t b_lo = ((smem_u32(b) >> 4) & 0x3FFFu) | (swb ? (1u << 16) : ((uint32_t)n << 16));
        const uint32_t a_hi = swa ? DESC_HI_SW128 : DESC_HI, b_hi = swb ? DESC_HI_SW128 : DESC_HI;
        const uint32_t a_step = swa ? 2u : 256u, b_step = swb ? 2u : 2u * (uint32_t)n;     // one K = 16 step, in 16-byte units
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tb + ((i & 1) ? 256u : 0u);
            const int k = i & 3;
            mma_ss2(d, a_lo + (uint32_t)k * a_step, a_hi, b_lo + (uint32_t)k * b_step, b_hi, idesc, i > 1);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// Numerical check of the swizzled layout: element (row r, k) at byte r * 128 + ((k / 8) ^ (r % 8)) * 16 + (k % 8) * 2
__global__ void __launch_bounds__(128, 1) check_sw_kernel(float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __nv_bfloat16 *a = reinterpret_cast<__nv_bfloat16 *>(smem);
    __nv_bfloat16 *b = reinterpret_cast<__nv_bfloat16 *>(smem + 16384);
    const int N = 64, K = 64;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        a[r * 64 + (((k / 8) ^ (r % 8)) * 8) + (k % 8)] = __float2bfloat16(((r * 7 + k * 3) % 13 - 6) / 8.f);
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        b[n * 64 + (((k / 8) ^ (n % 8)) * 8) + (k % 8)] = __float2bfloat16(((n * 5 + k) % 11 - 5) / 4.f);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const int r = threadIdx.x;
    const uint32_t lane_base = (uint32_t)((r / 32) * 32) << 16;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(N);
        const uint32_t a_lo = ((smem_u32(a) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = ((smem_u32(b) >> 4) & 0x3FFFu) | (1u << 16);
        for (int k = 0; k < K / 16; ++k) mma_ss2(tb, a_lo + (uint32_t)k * 2u, DESC_HI_SW128, b_lo + (uint32_t)k * 2u, DESC_HI_SW128, idesc, k > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; ++c) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tb + lane_base + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[r * N + c] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
    long long *out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    const int ns[] = {64, 96, 112, 128, 160, 192, 208, 240, 256};
    const int iters = 2000;
    printf("cycles per MMA (M=128,K=16,bf16), issue / complete; grid 148\n");
    for (int mode = 0; mode < 4; ++mode) {
        for (int n : ns) {
            if ((mode & 2) && n > 224) continue;
            bench_kernel<<<148, 128, 49152>>>(n, mode, iters, out);
            bench_kernel<<<148, 128, 49152>>>(n, mode, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("mode %s %s N=%3d : issue %.1f  complete %.1f  (floor %d)  %s\n", (mode & 2) ? "TS" : "SS", (mode & 1) ? "alt2" : "same",
                   n, (double)h[0] / iters, (double)h[1] / iters, n / 2, cudaGetErrorString(e));
        }
    }
    cudaFuncSetAttribute(bench_sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    cudaFuncSetAttribute(check_sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    for (int sw = 0; sw < 4; ++sw) {
        for (int n : ns) {
            bench_sw_kernel<<<148, 128, 49152>>>(n, sw, iters, out);
            bench_sw_kernel<<<148, 128, 49152>>>(n, sw, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("SS alt2 swizzle128 A=%d B=%d N=%3d : issue %.1f  complete %.1f  (floor %d)  %s\n", sw & 1, (sw >> 1) & 1, n, (double)h[0] / iters,
                   (double)h[1] / iters, n / 2, cudaGetErrorString(e));
        }
    }
    {
        float *o;
        cudaMalloc(&o, 128 * 64 * 4);
        check_sw_kernel<<<1, 128, 49152>>>(o);
        cudaError_t e = cudaDeviceSynchronize();
        static float ho[128 * 64];
        cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
        double err = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (((r * 7 + k * 3) % 13 - 6) / 8.0) * (((n * 5 + k) % 11 - 5) / 4.0);
                err = fmax(err, fabs(ho[r * 64 + n] - ref));
            }
        printf("check swizzle128 (%s): max err %.3g\n", cudaGetErrorString(e), err);
    }
    float *ss, *ts;
    cudaMalloc(&ss, 128 * 64 * 4);
    cudaMalloc(&ts, 128 * 64 * 4);
    check_kernel<<<1, 128, 49152>>>(ss, ts);
    cudaError_t e = cudaDeviceSynchronize();
    static float hss[128 * 64], hts[128 * 64];
    cudaMemcpy(hss, ss, sizeof(hss), cudaMemcpyDeviceToHost);
    cudaMemcpy(hts, ts, sizeof(hts), cudaMemcpyDeviceToHost);
    double max_ref = 0, err_ss = 0, err_ts = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (((r * 7 + k * 3) % 13 - 6) / 8.0) * (((n * 5 + k) % 11 - 5) / 4.0);
            max_ref = fmax(max_ref, fabs(ref));
            err_ss = fmax(err_ss, fabs(hss[r * 64 + n] - ref));
            err_ts = fmax(err_ts, fabs(hts[r * 64 + n] - ref));
        }
    printf("check (%s): max|ref| %.3f  max err SS %.3g  TS %.3g\n", cudaGetErrorString(e), max_ref, err_ss, err_ts);
    return 0;
}
