// Cycles per tcgen05.mma (kind::f16, bf16, M128, K16) issued back to back by one thread:
//   SS (A and B from shared memory) with N = 64, 128, 256 and TS (A from tensor memory) with N = 128, 256.
// Operand contents do not matter (zeros); one CTA per SM on a few SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_pace_probe umma_pace_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int mode, int N, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned long long bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a = desc_k_sw128(base), b = desc_k_sw128(base + 32 * 1024);
        uint64_t ak[4], bk[4];
        for (int k = 0; k < 4; ++k) { ak[k] = a + 2 * k; bk[k] = b + 2 * k; }     // 32 B further per k step
        long long t0 = clock64();
        if (mode == 0) {
            for (int i = 0; i < iters; i += 8) {                     // precomputed descriptors, 8 MMAs per trip
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_ss(tm, ak[j & 3], bk[j & 3], idesc);
            }
        } else {
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_ts(tm, tm + 256 + (j & 3) * 8, bk[j & 3], idesc);   // A: 8 columns per k step
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    const int smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    struct { int mode, N; const char *name; } cases[] = {{0, 64, "SS N=64"}, {0, 128, "SS N=128"}, {0, 256, "SS N=256"},
                                                          {1, 128, "TS N=128"}, {1, 256, "TS N=256"}};
    for (auto &c : cases) {
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<148, 128, smem>>>(c.mode, c.N, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        printf("%-10s %.1f cycles per MMA (M128 K16 bf16), all 148 SMs busy\n", c.name, (double)out[0] / iters);
    }
    return 0;
}
