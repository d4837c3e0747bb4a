// Developer probe: how many clusters of size 2 / 4 / 8 with ~225 KiB of shared memory can be resident at once?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p && threadIdx.x == 9999) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 148, 1, 1);
        cfg.blockDim = dim3(352, 1, 1);
        cfg.dynamicSmemBytes = 225 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
