// How fast can ONE warp stream HBM through a ring of TMA bulk copies?  (banded sweep, small sizes)
// mode 0: producer warp (warp 1) waits for "released" and issues; consumer (warp 0) waits "full", releases
// mode 1: the consumer lane 0 re-issues the copy itself right after the release
// mode 2: like 0 but the consumer also reads the block (8 LDS.64 per step, 4 steps per 8 KB)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void issue(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
extern __shared__ __align__(128) unsigned char smem[];
__global__ void k(const double *src, size_t per_cta_bytes, int nblk, int depth, unsigned bytes, int mode, double *out) {
    __shared__ unsigned long long full[16], empty[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&empty[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char *base = (const char *)src + (size_t)blockIdx.x * per_cta_bytes;
    const unsigned data = sa(smem), f0 = sa(full), e0 = sa(empty);
    double acc = 0;
    if (mode == 1) {
        if (warp == 0) {
            if (lane == 0) for (int u = 0; u < depth && u < nblk; ++u) issue(data + u * bytes, base + (size_t)u * bytes, bytes, f0 + u * 8);
            for (int u = 0; u < nblk; ++u) {
                const int buf = u % depth;
                mwait(f0 + buf * 8, (u / depth) & 1);
                __syncwarp();
                if (lane == 0 && u + depth < nblk) issue(data + buf * bytes, base + (size_t)(u + depth) * bytes, bytes, f0 + buf * 8);
            }
        }
    } else {
        if (warp == 1) {
            if (lane == 0) for (int u = 0; u < nblk; ++u) {
                const int buf = u % depth;
                mwait(e0 + buf * 8, ((u / depth) & 1) ^ 1);
                issue(data + buf * bytes, base + (size_t)u * bytes, bytes, f0 + buf * 8);
            }
        } else if (warp == 0) {
            for (int u = 0; u < nblk; ++u) {
                const int buf = u % depth;
                mwait(f0 + buf * 8, (u / depth) & 1);
                if (mode == 2) {
                    const unsigned a = data + buf * bytes + lane * 8;
                    for (unsigned o = 0; o < bytes; o += 256) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + o)); acc += v; }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(e0 + buf * 8) : "memory");
            }
        }
    }
    if (acc == 1.2345) out[0] = acc;
}
int main() {
    const size_t total = (size_t)3 << 30;
    double *src, *out; cudaMalloc(&src, total); cudaMalloc(&out, 8); cudaMemset(src, 0, total);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    printf("mode ctas depth blockKB | us  GB/s  ns/block/warp  implied latency us (depth*ns/block)\n");
    for (int mode : {0, 1, 2}) for (int ctas : {148, 296, 444}) for (int cfg = 0; cfg < 6; ++cfg) {
        const int depths[] = {8, 4, 2, 4, 16, 32}; const unsigned kbs[] = {8, 8, 8, 16, 4, 2};
        const int depth = depths[cfg]; const unsigned bytes = kbs[cfg] * 1024;
        if (depth > 16) continue;
        const size_t per_cta = (size_t)4 << 20;      // 4 MB per CTA
        const int nblk = (int)(per_cta / bytes);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            k<<<ctas, 64, 65536>>>(src, per_cta, nblk, depth, bytes, mode, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double us = ms * 1e3, gbs = (double)ctas * per_cta / (ms * 1e-3) / 1e9, nsblk = us * 1e3 / nblk;
        printf("%d %4d %3d %3u | %8.1f %7.0f %8.0f %6.2f\n", mode, ctas, depth, kbs[cfg], us, gbs, nsblk, depth * nsblk / 1e3);
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
