// Development aid: how many thread-block clusters of each size are co-resident on this GPU for a kernel that takes a
// whole SM (1 CTA per SM by shared memory)?   nvcc -arch=sm_100a -o /tmp/probe_clusters tools/probe_clusters.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: %d SMs, %zu KB smem/block optin\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin / 1024);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int threads : {1024, 768, 512}) {
        for (int cl = 1; cl <= 16; cl++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); n = -1; }
            printf("threads %4d cluster %2d: %3d clusters = %3d SMs\n", threads, cl, n, n * cl);
        }
    }
    return 0;
}
