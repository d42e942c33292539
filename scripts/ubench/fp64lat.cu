// Micro-benchmarks behind the banded sweep's per-step cost (one warp per scheduler):
// FP64 dependent latency, FP64 issue rate of one warp, LDS.64 latency, and the forward step itself.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double lds(unsigned a, int off) {
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + off)); return v;
}
template <int OFF> __device__ __forceinline__ double ldso(unsigned a) {
    double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v;
}
extern __shared__ double sh[];
__global__ void k(int mode, int iters, double *out, long long *cyc, double seed) {
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 1.0 / (1.0 + i + seed);
    __syncthreads();
    unsigned a = (unsigned)__cvta_generic_to_shared(sh) + lane * 8;
    double y1 = seed, y2 = seed * 2, y3 = seed * 3, y4 = seed * 4, y5 = seed * 5, y6 = seed * 6, y7 = seed * 7;
    double acc = 0;
    long long t0 = clock64();
    if (mode == 0) {            // dependent DFMA chain
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 16; ++q) y1 = __fma_rn(y1, y2, y3);
        }
    } else if (mode == 1) {     // 8 independent DFMA chains
        double z0 = seed, z1 = seed + 1, z2 = seed + 2, z3 = seed + 3, z4 = seed + 4, z5 = seed + 5, z6 = seed + 6, z7 = seed + 7;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                z0 = __fma_rn(z0, y2, y3); z1 = __fma_rn(z1, y2, y3); z2 = __fma_rn(z2, y2, y3); z3 = __fma_rn(z3, y2, y3);
                z4 = __fma_rn(z4, y2, y3); z5 = __fma_rn(z5, y2, y3); z6 = __fma_rn(z6, y2, y3); z7 = __fma_rn(z7, y2, y3);
            }
        }
        y1 = z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7;
    } else if (mode == 2) {     // dependent LDS.64 chain (pointer chase through address arithmetic)
        unsigned p = a;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 16; ++q) { double v = lds(p, 0); p = a + (((unsigned)__double2loint(v)) & 0x700u); acc += v; }
        }
        y1 = acc + p;
    } else if (mode == 3) {     // forward step, rotated rows, from shared memory (4 steps per iteration)
        double l0 = ldso<0>(a), l1 = ldso<256>(a), l2 = ldso<512>(a), l3 = ldso<768>(a), l4 = ldso<1024>(a), l5 = ldso<1280>(a), l6 = ldso<1536>(a), l7 = ldso<1792>(a);
        double rc = seed;
#define STEP(ON)  { double pa = l3 * y3; l3 = ldso<ON + 768>(a); double pb = l2 * y2; l2 = ldso<ON + 512>(a); \
        pa = __fma_rn(l5, y5, pa); l5 = ldso<ON + 1280>(a); pb = __fma_rn(l4, y4, pb); l4 = ldso<ON + 1024>(a); \
        pa = __fma_rn(l7, y7, pa); l7 = ldso<ON + 1792>(a); pb = __fma_rn(l6, y6, pb); l6 = ldso<ON + 1536>(a); \
        const double t = __fma_rn(-l1, y1, rc - (pa + pb)); l1 = ldso<ON + 256>(a); const double y = t * l0; l0 = ldso<ON>(a); \
        y7 = y6; y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y; acc += rc * rc; }
        for (int i = 0; i < iters; ++i) { STEP(2048) STEP(4096) STEP(6144) STEP(0) }
        y1 += acc;
    } else if (mode == 4) {     // forward step arithmetic only, rows constant in registers
        double l0 = ldso<0>(a), l1 = ldso<256>(a), l2 = ldso<512>(a), l3 = ldso<768>(a), l4 = ldso<1024>(a), l5 = ldso<1280>(a), l6 = ldso<1536>(a), l7 = ldso<1792>(a);
        double rc = seed;
#define STEPN  { double pa = l3 * y3; double pb = l2 * y2; pa = __fma_rn(l5, y5, pa); pb = __fma_rn(l4, y4, pb); \
        pa = __fma_rn(l7, y7, pa); pb = __fma_rn(l6, y6, pb); const double t = __fma_rn(-l1, y1, rc - (pa + pb)); const double y = t * l0; \
        y7 = y6; y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y; acc += rc * rc; }
        for (int i = 0; i < iters; ++i) { STEPN STEPN STEPN STEPN }
        y1 += acc;
    } else if (mode == 5) {     // independent LDS.64 stream: 8 loads then one add each
        for (int i = 0; i < iters; ++i) {
            double v0 = ldso<0>(a), v1 = ldso<256>(a), v2 = ldso<512>(a), v3 = ldso<768>(a), v4 = ldso<1024>(a), v5 = ldso<1280>(a), v6 = ldso<1536>(a), v7 = ldso<1792>(a);
            acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
        }
        y1 = acc;
    } else if (mode == 6) {     // dependent DADD chain
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 16; ++q) y1 = y1 + y2;
        }
    } else if (mode == 7) {     // dependent FFMA chain (fp32) for comparison
        float f = (float)seed, g = (float)y2, h = (float)y3;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 16; ++q) f = __fmaf_rn(f, g, h);
        }
        y1 = f;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = y1 + y2 + y3 + y4 + y5 + y6 + y7;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 32 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const char *names[] = {"dep DFMA chain (per op)", "8 indep DFMA (per op)", "dep LDS.64+addr (per load)", "fwd step rotated LDS (per step)",
                           "fwd step no LDS (per step)", "8 indep LDS.64 + 7 DADD (per group)", "dep DADD chain (per op)", "dep FFMA chain (per op)"};
    const int per[] = {16, 16, 16, 4, 4, 1, 16, 16};
    const int iters = 2000;
    for (int threads : {32, 128, 384}) {
        for (int mode = 0; mode < 8; ++mode) {
            k<<<148, threads, 32768>>>(mode, iters, out, cyc, 1.0);
            k<<<148, threads, 32768>>>(mode, iters, out, cyc, 1.0);
            cudaDeviceSynchronize();
            long long h[32]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("threads/CTA %3d  %-40s %.1f cycles\n", threads, names[mode], (double)h[0] / iters / per[mode]);
        }
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
