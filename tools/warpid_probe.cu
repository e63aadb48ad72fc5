// warpid_probe.cu -- which hardware warp slots (%warpid; slot % 4 = sub-partition) do the warps of two co-resident 10-warp CTAs get?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/warpid_probe tools/warpid_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 2) probe(int *out)
{
    extern __shared__ char sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if ((threadIdx.x & 31) == 0) {
        int *o = out + (blockIdx.x * 10 + (threadIdx.x >> 5)) * 2;
        o[0] = (int)smid; o[1] = (int)wid;
    }
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) { sm[threadIdx.x] = (char)t0; }      // keep every CTA resident while the others start
}
int main()
{
    const int n = 289;
    int *d, *h = new int[n * 20];
    cudaMalloc(&d, n * 20 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    probe<<<n, 320, 112 * 1024>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
    cudaMemcpy(h, d, n * 20 * sizeof(int), cudaMemcpyDeviceToHost);
    for (int sm = 0; sm < 3; ++sm)
        for (int b = 0; b < n; ++b)
            if (h[b * 20] == sm) {
                printf("sm %d block %3d slots:", sm, b);
                for (int w = 0; w < 10; ++w) printf(" %2d", h[(b * 10 + w) * 2 + 1]);
                printf("   slot%%4:");
                for (int w = 0; w < 10; ++w) printf(" %d", h[(b * 10 + w) * 2 + 1] & 3);
                printf("\n");
            }
    return 0;
}
