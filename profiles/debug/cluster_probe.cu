// How many thread-block clusters of a given size does a B200 run at once?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_probe profiles/debug/cluster_probe.cu && /tmp/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(736, 1) probe_kernel(float* out) {
    extern __shared__ float smem[];
    if (out) out[blockIdx.x] = smem[threadIdx.x];
}

int main() {
    const int smem_sizes[] = {200 * 1024, 100 * 1024, 48 * 1024};
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int smem : smem_sizes) {
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int size : {1, 2, 4, 6, 8, 16}) {
            cudaLaunchConfig_t config = {};
            config.gridDim = dim3(size * 64);
            config.blockDim = dim3(736);
            config.dynamicSmemBytes = smem;
            cudaLaunchAttribute attribute;
            attribute.id = cudaLaunchAttributeClusterDimension;
            attribute.val.clusterDim.x = size;
            attribute.val.clusterDim.y = 1;
            attribute.val.clusterDim.z = 1;
            config.attrs = &attribute;
            config.numAttrs = 1;
            int clusters = -1;
            cudaError_t status = cudaOccupancyMaxActiveClusters(&clusters, probe_kernel, &config);
            printf("smem %3d KB, 736 threads, cluster of %2d: %3d clusters at once (%3d SMs)%s\n", smem / 1024,
                   size, clusters, clusters * size, status == cudaSuccess ? "" : cudaGetErrorString(status));
        }
    }
    return 0;
}
