// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of the shared-memory
// operand layout (no-swizzle K-major with arbitrary row pitch vs 128-byte swizzle) and of how often the A / B
// descriptors change.  Operands are whatever is in shared memory (zeros): timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "../c2-matching_b200/csrc/c2m_common.cuh"

using namespace c2m;

struct Cfg {
    int N;            // MMA N
    int nA, nB;       // distinct A / B start addresses rotated through (1 = same operand every MMA)
    int a_step, b_step;   // bytes between rotated operands
    int a_sbo, a_lbo, b_sbo, b_lbo;
    int layout;       // 0 = no swizzle, 2 = SWIZZLE_128B
    int nacc;         // accumulators rotated through
    int iters;
    int N2;           // if > 0: every second MMA uses N2 instead of N (and A offset a_alt), like the conv's x_lo MMA
    int a_alt;
};

template <int R>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(Cfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 180 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        const uint32_t hi_a = (uint32_t)(c.a_sbo >> 4) | (1u << 14) | ((uint32_t)c.layout << 29);
        const uint32_t hi_b = (uint32_t)(c.b_sbo >> 4) | (1u << 14) | ((uint32_t)c.layout << 29);
        const uint32_t idesc = umma_idesc_f16(128, c.N, 0), idesc2 = umma_idesc_f16(128, c.N2 > 0 ? c.N2 : c.N, 0);
        // the R operand pairs of one round, precomputed: the timed loop is R (or 2R) MMAs, fully unrolled
        uint64_t da[R], da2[R], db[R];
        uint32_t dd[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t aa = a0 + (r % c.nA) * c.a_step, bb = b0 + (r % c.nB) * c.b_step;
            da[r] = ((uint64_t)hi_a << 32) | ((aa & 0x3FFFF) >> 4) | ((uint32_t)(c.a_lbo >> 4) << 16);
            da2[r] = ((uint64_t)hi_a << 32) | (((aa + c.a_alt) & 0x3FFFF) >> 4) | ((uint32_t)(c.a_lbo >> 4) << 16);
            db[r] = ((uint64_t)hi_b << 32) | ((bb & 0x3FFFF) >> 4) | ((uint32_t)(c.b_lbo >> 4) << 16);
            dd[r] = tmem_base + (r % c.nacc) * 128;
        }
        const bool pair = c.N2 > 0;
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {           // rep 0 = warm-up
            __syncwarp();
            t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < c.iters; i += R) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        umma_f16(dd[r], da[r], db[r], idesc, 1);
                        if (pair) umma_f16(dd[r], da2[r], db[r], idesc2, 1);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, rep & 1);
            tc_fence_after();
            t1 = clock64();
        }
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static double run(const char *name, Cfg c) {
    static long long *d_out = nullptr;
    int sms = 148;
    if (!d_out) cudaMalloc(&d_out, sizeof(long long) * 256);
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(umma_bench_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_bench_kernel<12><<<sms, 128, smem>>>(c, d_out);      // 12 = lcm of the rotation lengths used (1, 3, 4)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-64s ERROR %s\n", name, cudaGetErrorString(e));
        exit(1);
    }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const int per = c.N2 > 0 ? 2 : 1;
    const double cyc = (double)h[sms / 2] / (c.iters * per);
    const double math = c.N2 > 0 ? (c.N + c.N2) / 2.0 / 2.0 : c.N / 2.0;      // cycles of math per MMA at 8192 flop/clk/SM
    printf("%-64s %7.1f cycles/MMA  (math %5.1f)  %.0f%% of peak\n", name, cyc, math, 100.0 * math / cyc);
    return cyc;
}

int main() {
    const int IT = 4092;        // multiple of 12
    // no-swizzle K-major: core matrix = 8 rows x 16 B contiguous (128 B); SBO = pitch between 8-row groups along M/N,
    // LBO = pitch between the two 8-element K halves of a K = 16 step.
    for (int N : {256, 128, 64}) {
        char nm[128];
        Cfg c{N, 4, 4, 4096, 8192, 128, 2048, 128, (N * 16), 0, 2, IT, 0, 0};
        snprintf(nm, sizeof nm, "noswz N=%d A,B rotate(4) SBO128 aligned", N);
        run(nm, c);
        c.nA = 1; c.nB = 1;
        snprintf(nm, sizeof nm, "noswz N=%d A,B fixed", N);
        run(nm, c);
        c.nA = 4; c.nB = 1;
        snprintf(nm, sizeof nm, "noswz N=%d A rotate, B fixed", N);
        run(nm, c);
        c.nA = 1; c.nB = 4;
        snprintf(nm, sizeof nm, "noswz N=%d A fixed, B rotate", N);
        run(nm, c);
        c.nA = 4; c.nB = 4; c.nacc = 1;
        snprintf(nm, sizeof nm, "noswz N=%d rotate, ONE accumulator", N);
        run(nm, c);
    }
    {   // the conv kernel's A geometry: row pitch 160 B, octet pitch 2880 B, tap shifts of 16 B
        Cfg c{128, 3, 4, 16, 128 * 16 * 2 * 4, 160, 2880, 128, 128 * 16, 0, 2, IT, 0, 0};
        run("conv-like N=128: A pitch 160/2880, 16-B tap shifts, B rotate", c);
        c.N2 = 64; c.a_alt = 11520;
        run("conv-like pair N=128 + N=64 (x_lo at +11520), per MMA", c);
        c.N2 = 0; c.N = 64;
        run("conv-like N=64 only", c);
        Cfg d{128, 3, 4, 256, 128 * 16 * 2 * 4, 128, 2816, 128, 128 * 16, 0, 2, IT, 0, 0};
        run("same but A pitch 128/2816 and 256-B shifts (aligned)", d);
    }
    // 128-byte swizzle, K-major: rows of 128 B, 8-row atoms of 1024 B (SBO = 1024), a K = 16 step = +32 B inside the row
    for (int N : {256, 128, 64}) {
        char nm[128];
        Cfg c{N, 4, 4, 32, 32, 1024, 16, 1024, 16, 2, 2, IT, 0, 0};
        snprintf(nm, sizeof nm, "SW128 N=%d A,B advance 32 B along K (4 steps)", N);
        run(nm, c);
        c.a_step = 128; c.nA = 3;     // pixel-shifted window: +128 B = next row (tap shift)
        snprintf(nm, sizeof nm, "SW128 N=%d A shifted by whole rows (+128 B x3), B along K", N);
        run(nm, c);
    }
    {
        Cfg c{128, 4, 4, 32, 32, 1024, 16, 1024, 16, 2, 2, IT, 64, 16384};
        run("SW128 pair N=128 + N=64 (other A), per MMA", c);
        Cfg d{128, 4, 4, 32, 32, 2048, 16, 1024, 16, 2, 2, IT, 64, 32768};
        run("SW128 pair, A with SBO 2048 (16-px pitch rows)", d);
    }
    return 0;
}
