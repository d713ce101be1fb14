// Issue pace of the FP32 pipe on sm_100a: scalar FFMA / FADD / FMUL against the packed f32x2 forms, alone and
// mixed with integer (ALU pipe) and shared-memory (LSU) instructions.  Decides how the FFT butterflies of the
// spectral-loss kernel are written.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pace_probe fp32_pace_probe.cu
// Output: one line per variant: lane-ops per clock per SM (FMA = 1 lane-op), at the clock nvml would report
// (computed from the measured time and the clock64 span of the kernel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fadd(float a, float b) {
    float r;
    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fmul(float a, float b) {
    float r;
    asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

constexpr int CH = 8;          // independent chains per thread

// (the clock64 span is CTA 0's; with 8 CTAs per SM the CTAs run in waves, so compare the ms column)
// MODE 0 ffma  1 ffma2  2 fadd  3 fadd2  4 fmul  5 fmul2  6 ffma2 + 1 iadd/lop per ffma2  7 ffma2 + 1 LDS.128 per 4
//      8 ffma + 1 int op per ffma   9 ffma2 + 2 int ops per ffma2   10 ffma2 + 1 LDS.64 + 1 STS.64 per 4
//      11 fadd2 + fmul2 + ffma2 round robin (butterfly-like mix)
template <int MODE>
__global__ void __launch_bounds__(256) pace(int iters, float seed, float *out, long long *clk) {
    __shared__ __align__(16) float sm[256 * 4 + 64];
    float a[CH], b = seed, c = seed * 0.5f;
    uint64_t A[CH], B2 = pk2(seed, seed), C2 = pk2(c, c);
    int x[CH];
    uint64_t Bv[CH], Cv[CH];
    float bv[CH], cv[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = seed + i; A[i] = pk2(a[i], a[i] + 1.f); x[i] = threadIdx.x + i;
        bv[i] = seed * (1.f + 0.001f * i); cv[i] = seed * (0.5f - 0.001f * i);
        Bv[i] = pk2(bv[i], bv[i] * 1.001f); Cv[i] = pk2(cv[i], cv[i] * 1.001f);
    }
    sm[threadIdx.x * 4] = seed; sm[threadIdx.x * 4 + 1] = seed; sm[threadIdx.x * 4 + 2] = seed; sm[threadIdx.x * 4 + 3] = seed;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) a[i] = ffma(a[i], b, c);
                if (MODE == 1) A[i] = fma2(A[i], B2, C2);
                if (MODE == 2) a[i] = fadd(a[i], b);
                if (MODE == 3) A[i] = add2(A[i], B2);
                if (MODE == 4) a[i] = fmul(a[i], b);
                if (MODE == 5) A[i] = mul2(A[i], B2);
                if (MODE == 6) { A[i] = fma2(A[i], B2, C2); x[i] = (x[i] ^ it) + u; }
                if (MODE == 7) {
                    A[i] = fma2(A[i], B2, C2);
                    if ((i & 3) == 0) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                     : "r"(sbase + (((threadIdx.x + i + u) & 255) << 4)));
                        x[i] += __float_as_int(v.x) + __float_as_int(v.w) + __float_as_int(v.y) + __float_as_int(v.z);
                    }
                }
                if (MODE == 8) { a[i] = ffma(a[i], b, c); x[i] = (x[i] ^ it) + u; }
                if (MODE == 9) { A[i] = fma2(A[i], B2, C2); x[i] = (x[i] ^ it) + u; x[i] = (x[i] & 0xffff) | (it << 16); }
                if (MODE == 10) {
                    A[i] = fma2(A[i], B2, C2);
                    if ((i & 3) == 0) {
                        float2 v;
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y)
                                     : "r"(sbase + (((threadIdx.x + i + u) & 255) << 4)));
                        x[i] += __float_as_int(v.x);
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sbase + (threadIdx.x << 4) + 8), "f"(v.x), "f"(v.y));
                    }
                }
                if (MODE == 12) A[i] = fma2(A[i], Bv[i], Cv[i]);          // three distinct register pairs per instruction
                if (MODE == 13) a[i] = ffma(a[i], bv[i], cv[i]);          // three distinct registers per instruction
                if (MODE == 14) { A[i] = fma2(Bv[i], Cv[i], A[i]); Cv[i] = add2(Cv[i], A[i]); }   // Reinsch-like pair
                if (MODE == 11) {
                    if (i % 3 == 0) A[i] = add2(A[i], B2);
                    else if (i % 3 == 1) A[i] = mul2(A[i], B2);
                    else A[i] = fma2(A[i], B2, C2);
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i]));
        s += a[i] + lo + hi + (float)x[i] + bv[i] + cv[i];
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(Cv[i]));
        s += lo + hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(Bv[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int lanes_per_instr, int ctas_per_sm, float *out, long long *clk) {
    const int iters = 4096, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pace<MODE><<<grid, 256>>>(iters, 1.0001f, out, clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pace<MODE><<<grid, 256>>>(iters, 1.0001f, out, clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[8];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    // FP instructions per thread: iters * 4 * CH ; per CTA x256 ; per SM x ctas_per_sm
    const double fp_instr_thread = (double)iters * 4 * CH;
    const double warp_instr_sm = fp_instr_thread * 8 * ctas_per_sm;            // 8 warps per CTA
    const double cyc = (double)h[0];                                            // CTA 0's span (all CTAs resident together)
    printf("%-44s ctas/SM %d  %8.3f ms  %9.0f cyc  fp warp-instr/clk/SM %.3f  lane-ops/clk/SM %.1f  (clk %.0f MHz)\n", name,
           ctas_per_sm, ms, cyc, warp_instr_sm / cyc, warp_instr_sm * 32 * lanes_per_instr / cyc, cyc / (ms * 1e3));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&clk, 148 * 8 * sizeof(long long));
    for (int c : {8}) {
        run<0>("ffma (scalar, 3 registers)", 1, c, out, clk);
        run<1>("ffma2 (fma.rn.f32x2)", 2, c, out, clk);
        run<2>("fadd (scalar)", 1, c, out, clk);
        run<3>("fadd2", 2, c, out, clk);
        run<4>("fmul (scalar)", 1, c, out, clk);
        run<5>("fmul2", 2, c, out, clk);
        run<11>("fadd2 / fmul2 / ffma2 mix", 2, c, out, clk);
        run<12>("ffma2, three distinct register pairs", 2, c, out, clk);
        run<13>("ffma, three distinct registers", 1, c, out, clk);
        run<14>("ffma2 + fadd2 (Reinsch step, distinct pairs)", 2, c, out, clk);
        run<8>("ffma + 2 int ops (xor, add) per ffma", 1, c, out, clk);
        run<6>("ffma2 + 2 int ops (xor, add) per ffma2", 2, c, out, clk);
        run<9>("ffma2 + 4 int ops per ffma2", 2, c, out, clk);
        run<7>("ffma2 + 1 LDS.128 per 4 ffma2", 2, c, out, clk);
        run<10>("ffma2 + LDS.64 + STS.64 per 4 ffma2", 2, c, out, clk);
    }
    return 0;
}
