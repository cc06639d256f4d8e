// Where do the warps of 4-warp blocks land?  (smid, hardware warp slot) of every warp of a config-2-shaped launch:
// 1024 blocks x 128 threads, ~19.8 KB of dynamic shared memory, 7 blocks per SM.  Development aid.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 7) probe(int *out, long long spin)
{
    extern __shared__ unsigned char sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2] = (int)smid; out[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 2 + 1] = (int)warpid; }
    long long t0 = clock64();
    while (clock64() - t0 < spin) { sm[threadIdx.x] += 1; }
}
int main()
{
    const int B = 1024;
    int *d; cudaMalloc(&d, B * 4 * 2 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
    probe<<<B, 128, 19792>>>(d, 2000000);
    cudaDeviceSynchronize();
    static int h[B * 4 * 2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    // per SM: which blocks, and the warp slots of their warps
    for (int s = 0; s < 3; ++s) {
        printf("SM %d:", s);
        for (int b = 0; b < B; ++b) if (h[b * 8] == s) printf("  b%d[%d %d %d %d]", b, h[b * 8 + 1], h[b * 8 + 3], h[b * 8 + 5], h[b * 8 + 7]);
        printf("\n");
    }
    // histogram: (warp index in block) x (slot % 4)
    int hist[4][4] = {};
    for (int b = 0; b < B; ++b) for (int w = 0; w < 4; ++w) hist[w][h[(b * 4 + w) * 2 + 1] & 3]++;
    for (int w = 0; w < 4; ++w) printf("warp %d: slot%%4 histogram %d %d %d %d\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
    // how many distinct SMs does block b share with b+148?
    int same = 0; for (int b = 0; b + 148 < B; ++b) same += h[b * 8] == h[(b + 148) * 8];
    printf("blocks b and b+148 on the same SM: %d of %d\n", same, B - 148);
    int same2 = 0; for (int b = 0; b + 1 < B; ++b) same2 += h[b * 8] == h[(b + 1) * 8];
    printf("blocks b and b+1 on the same SM: %d of %d\n", same2, B - 1);
    return 0;
}
