// How many clusters of a 1-CTA/SM kernel (640 threads, ~220 KiB dynamic shared memory) can be co-resident
// for cluster sizes 1, 2, 4, 8?  (cudaOccupancyMaxActiveClusters; decides whether wider TMA multicast pays)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/cluster_probe tools/probe/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(640, 1) probe_kernel(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    const int smem = 220 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        printf("cluster %2d: max active clusters %d -> %d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorName(e));
    }
    return 0;
}
