#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) *p = 1; }
int main() {
    for (int cs : {2, 4, 8}) {
        for (int smem : {206 * 1024, 100 * 1024}) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148 * 2); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("cluster %d smem %d KB: max active clusters %d (%s) -> %d SMs\n", cs, smem / 1024, n, cudaGetErrorString(e), n * cs);
        }
    }
    return 0;
}
