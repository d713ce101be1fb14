// Feasibility probe: can a 16-CTA (non-portable) cluster with ~223 KB of dynamic shared memory per CTA be
// scheduled on this B200, and how many such clusters run at once?  Also times a cluster barrier + DSMEM pull.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void probe_kernel(float *out, int steps, int smem_floats) {
    extern __shared__ float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank(), nblk = cluster.num_blocks();
    float *own = sm;                      // 2 x 256 floats double buffer
    float *full = sm + 512;               // nblk*256
    float acc = 0.f;
    for (int i = threadIdx.x; i < smem_floats; i += blockDim.x) sm[i] = 0.f;
    cluster.sync();
    for (int t = 0; t < steps; ++t) {
        own[(t & 1) * 256 + threadIdx.x] = (float)(t + rank);
        cluster.sync();
        for (int i = threadIdx.x; i < nblk * 64; i += blockDim.x) {
            const int r = i / 64, j = i % 64;
            const float4 *src = reinterpret_cast<const float4 *>(cluster.map_shared_rank(own + (t & 1) * 256, r));
            reinterpret_cast<float4 *>(full)[r * 64 + j] = src[j];
        }
        __syncthreads();
        acc += full[(threadIdx.x * 7) % (nblk * 256)];
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

int main() {
    for (int csize : {8, 16}) {
        for (size_t smem : {(size_t)100 * 1024, (size_t)200 * 1024, (size_t)223 * 1024, (size_t)227 * 1024}) {
            cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize * 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            int nclusters = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe_kernel, &cfg);
            printf("cluster %2d smem %3zu KB: maxActiveClusters=%d (%s)\n", csize, smem / 1024, nclusters, cudaGetErrorString(e));
            if (e == cudaSuccess && nclusters > 0) {
                float *out; cudaMalloc(&out, 4096);
                cfg.gridDim = dim3(csize * (nclusters < 8 ? nclusters : 8));
                int steps = 400, sf = (int)(smem / 4);
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                cudaLaunchKernelEx(&cfg, probe_kernel, out, 4, sf); cudaDeviceSynchronize();
                cudaEventRecord(a);
                e = cudaLaunchKernelEx(&cfg, probe_kernel, out, steps, sf);
                cudaEventRecord(b); cudaError_t e2 = cudaDeviceSynchronize();
                float ms = 0; cudaEventElapsedTime(&ms, a, b);
                printf("   launch %s / %s: %d steps (barrier + 16 KB DSMEM pull) %.3f ms -> %.2f us per step\n",
                       cudaGetErrorString(e), cudaGetErrorString(e2), steps, ms, 1e3 * ms / steps);
                cudaFree(out);
            }
            cudaGetLastError();
        }
    }
    return 0;
}
