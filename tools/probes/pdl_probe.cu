// Dependent-launch gap between graph nodes, with and without programmatic dependent launch (PDL).
// A chain of 30 small kernels (each reads the previous one's output) is captured into a CUDA graph by stream
// capture; the PDL variant launches every kernel with cudaLaunchAttributeProgrammaticStreamSerialization and the
// kernels call griddepcontrol.launch_dependents / griddepcontrol.wait.  Prints microseconds per replay.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void step_kernel(const float *__restrict__ in, float *__restrict__ out, int n, int work) {
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = in[i];
        for (int k = 0; k < work; ++k) v = fmaf(v, 1.0001f, 0.5f);
        out[i] = v;
    }
}

template <bool PDL>
float run(int chain, int grid, int n, int work, float *a, float *b, cudaStream_t st) {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < chain; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = PDL ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const float *in = (i & 1) ? b : a;
        float *out = (i & 1) ? a : b;
        cudaError_t e = cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n, work);
        if (e != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(e));
    }
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (e != cudaSuccess) { printf("capture error %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { printf("instantiate error %s\n", cudaGetErrorString(e)); return -1; }
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(exec, st);
    cudaStreamSynchronize(st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 50; ++i) cudaGraphLaunch(exec, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / 50;
}

int main() {
    const int n = 1 << 20;
    float *a, *b;
    cudaMalloc(&a, n * sizeof(float));
    cudaMalloc(&b, n * sizeof(float));
    cudaMemset(a, 0, n * sizeof(float));
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int work : {16, 256}) {
        for (int grid : {148, 592, 4096}) {
            const float t0 = run<false>(30, grid, n, work, a, b, st);
            const float t1 = run<true>(30, grid, n, work, a, b, st);
            printf("work %4d grid %5d: 30-kernel chain %8.1f us plain, %8.1f us PDL  (%.2f us per edge saved)\n", work, grid, t0,
                   t1, (t0 - t1) / 30);
        }
    }
    float h;
    cudaMemcpy(&h, a, sizeof(float), cudaMemcpyDeviceToHost);
    printf("check %f, last error %s\n", h, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
