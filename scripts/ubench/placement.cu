// Where do the CTAs of a 3-per-SM persistent grid land, and which hardware warp slots do their warps get?
// (444 CTAs x 256 threads, 74 KB dynamic shared memory each -- the launch shape of the solve kernels.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *smid, int *warp0, int *warp4) {
    extern __shared__ char sm[];
    unsigned s, w;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    if (threadIdx.x == 0) { smid[blockIdx.x] = (int)s; warp0[blockIdx.x] = (int)w; sm[0] = 1; }
    if (threadIdx.x == 128) warp4[blockIdx.x] = (int)w;
    // keep every CTA resident until all have started (a persistent grid)
    __nanosleep(20000);
}
int main() {
    const int G = 444;
    int *d; cudaMalloc(&d, 3 * G * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 74112);
    k<<<G, 256, 74112>>>(d, d + G, d + 2 * G);
    int h[3 * G]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int same_mod = 0, slot_by_div = 0, w0mod[4] = {0, 0, 0, 0};
    for (int b = 0; b < G; ++b) { w0mod[h[G + b] & 3]++; }
    for (int b = 0; b + 148 < G; ++b) same_mod += h[b] == h[b + 148];
    // CTAs that share an SM: are their blockIdx / 148 all different?
    int per_sm[256][3], cnt[256] = {0};
    for (int b = 0; b < G; ++b) if (cnt[h[b]] < 3) per_sm[h[b]][cnt[h[b]]++] = b;
    int distinct_div = 0, sms = 0;
    for (int s = 0; s < 256; ++s) if (cnt[s] == 3) {
        ++sms; int a = per_sm[s][0] / 148, b2 = per_sm[s][1] / 148, c = per_sm[s][2] / 148;
        distinct_div += (a != b2 && a != c && b2 != c);
    }
    printf("CTAs b and b+148 on the same SM: %d of %d\n", same_mod, G - 148);
    printf("SMs with 3 CTAs: %d; of those with three different blockIdx/148: %d\n", sms, distinct_div);
    printf("warp 0 hardware slot mod 4 histogram: %d %d %d %d\n", w0mod[0], w0mod[1], w0mod[2], w0mod[3]);
    printf("first SMs: "); for (int s = 0; s < 6; ++s) if (cnt[s] == 3) printf("sm%d:{%d,%d,%d} ", s, per_sm[s][0], per_sm[s][1], per_sm[s][2]); printf("\n");
    printf("CTA 0..7: smid/warp0/warp4 "); for (int b = 0; b < 8; ++b) printf("%d/%d/%d ", h[b], h[G + b], h[2 * G + b]); printf("\n");
    for (int b = 148; b < 152; ++b) printf("CTA %d: %d/%d/%d ", b, h[b], h[G + b], h[2 * G + b]); printf("\n");
    for (int b = 296; b < 300; ++b) printf("CTA %d: %d/%d/%d ", b, h[b], h[G + b], h[2 * G + b]); printf("\n");
    return 0;
}
