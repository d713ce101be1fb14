// LayerNorm + LeakyReLU of the control net's MLP blocks in one pass ("fuse
// MLP+LayerNorm+LeakyReLU"; ddsp/core.py:122-129 builds Linear -> LayerNorm -> LeakyReLU three times per MLP).
//
// Reference path replaced: torch's layer_norm forward, LeakyReLU forward/backward element-wise kernels and
// the layer-norm backward pair, of which the gamma/beta reduction alone (GammaBetaBackwardCUDAKernel) takes
// 2.3 ms of a batch-64 training step on a B200 (profiles/r01_model_step_profile.txt).
//
// One warp per row, the row (N = 128 V floats, V <= 4) lives in registers: forward is one read and one write
// of the activations, backward one read of x and dy and one write of dx; the column sums for gamma/beta are
// kept per lane across the rows a warp walks, reduced over the CTA's warps in shared memory and over CTAs by
// a second small launch in a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace {

constexpr int kLnWarps = 8;

template <int V>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_lrelu_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                    float *__restrict__ y, float *__restrict__ stats, int64_t rows, float eps, float slope) {
    constexpr int N = 128 * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 g[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        g[i] = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
        b[i] = __ldg(reinterpret_cast<const float4 *>(beta) + i * 32 + lane);
    }
    for (int64_t r = (int64_t)blockIdx.x * kLnWarps + warp; r < rows; r += (int64_t)gridDim.x * kLnWarps) {
        float4 v[V];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = __ldg(reinterpret_cast<const float4 *>(x + r * N) + i * 32 + lane);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = ddsp_warp_sum(s) * (1.f / N);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(ddsp_warp_sum(q) * (1.f / N) + eps);
        if (lane == 0 && stats) {
            stats[2 * r] = mean;
            stats[2 * r + 1] = rstd;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4 o;
            o.x = fmaf(v[i].x * rstd, g[i].x, b[i].x);
            o.y = fmaf(v[i].y * rstd, g[i].y, b[i].y);
            o.z = fmaf(v[i].z * rstd, g[i].z, b[i].z);
            o.w = fmaf(v[i].w * rstd, g[i].w, b[i].w);
            o.x = o.x > 0.f ? o.x : o.x * slope;
            o.y = o.y > 0.f ? o.y : o.y * slope;
            o.z = o.z > 0.f ? o.z : o.z * slope;
            o.w = o.w > 0.f ? o.w : o.w * slope;
            reinterpret_cast<float4 *>(y + r * N)[i * 32 + lane] = o;
        }
    }
}

template <int V>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_lrelu_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ gamma,
                    const float *__restrict__ beta, const float *__restrict__ stats, float *__restrict__ dx,
                    float *__restrict__ partial, int64_t rows, float slope) {
    constexpr int N = 128 * V;
    __shared__ float red[kLnWarps][2][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 g[V], b[V], dg[V], db[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        g[i] = __ldg(reinterpret_cast<const float4 *>(gamma) + i * 32 + lane);
        b[i] = __ldg(reinterpret_cast<const float4 *>(beta) + i * 32 + lane);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t r = (int64_t)blockIdx.x * kLnWarps + warp; r < rows; r += (int64_t)gridDim.x * kLnWarps) {
        const float mean = __ldg(stats + 2 * r), rstd = __ldg(stats + 2 * r + 1);
        float4 xh[V], gz[V];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + r * N) + i * 32 + lane);
            const float4 d = __ldg(reinterpret_cast<const float4 *>(dy + r * N) + i * 32 + lane);
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            float4 dz;
            dz.x = fmaf(xh[i].x, g[i].x, b[i].x) > 0.f ? d.x : d.x * slope;
            dz.y = fmaf(xh[i].y, g[i].y, b[i].y) > 0.f ? d.y : d.y * slope;
            dz.z = fmaf(xh[i].z, g[i].z, b[i].z) > 0.f ? d.z : d.z * slope;
            dz.w = fmaf(xh[i].w, g[i].w, b[i].w) > 0.f ? d.w : d.w * slope;
            dg[i].x = fmaf(dz.x, xh[i].x, dg[i].x); dg[i].y = fmaf(dz.y, xh[i].y, dg[i].y);
            dg[i].z = fmaf(dz.z, xh[i].z, dg[i].z); dg[i].w = fmaf(dz.w, xh[i].w, dg[i].w);
            db[i].x += dz.x; db[i].y += dz.y; db[i].z += dz.z; db[i].w += dz.w;
            gz[i] = make_float4(dz.x * g[i].x, dz.y * g[i].y, dz.z * g[i].z, dz.w * g[i].w);
            s1 += (gz[i].x + gz[i].y) + (gz[i].z + gz[i].w);
            s2 += (gz[i].x * xh[i].x + gz[i].y * xh[i].y) + (gz[i].z * xh[i].z + gz[i].w * xh[i].w);
        }
        s1 = ddsp_warp_sum(s1) * (1.f / N);
        s2 = ddsp_warp_sum(s2) * (1.f / N);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4 o;
            o.x = rstd * (gz[i].x - s1 - xh[i].x * s2);
            o.y = rstd * (gz[i].y - s1 - xh[i].y * s2);
            o.z = rstd * (gz[i].z - s1 - xh[i].z * s2);
            o.w = rstd * (gz[i].w - s1 - xh[i].w * s2);
            reinterpret_cast<float4 *>(dx + r * N)[i * 32 + lane] = o;
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        reinterpret_cast<float4 *>(red[warp][0])[i * 32 + lane] = dg[i];
        reinterpret_cast<float4 *>(red[warp][1])[i * 32 + lane] = db[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * N; c += kLnWarps * 32) {
        const int which = c / N, col = c - which * N;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) s += red[w][which][col];
        partial[(size_t)blockIdx.x * 2 * N + c] = s;
    }
}

// d_gamma, d_beta = sum over the CTAs' partials: 8 lanes per column, fixed stride and combination order
__global__ void ln_param_grad_kernel(const float *__restrict__ partial, float *__restrict__ d_gamma,
                                     float *__restrict__ d_beta, int N, int slots) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    float s = 0.f;
    if (c < 2 * N)
        for (int k = sub; k < slots; k += 8) s += __ldg(partial + (size_t)k * 2 * N + c);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (c < 2 * N && sub == 0) {
        if (c < N) d_gamma[c] = s;
        else d_beta[c - N] = s;
    }
}

int ln_grid(int64_t rows) {
    const int64_t want = (rows + kLnWarps - 1) / kLnWarps;
    const int64_t cap = 2 * DDSP_SM_COUNT;
    return (int)(want < cap ? want : cap);
}

}  // namespace

// CTAs (= partial slots) ddsp_b200_ln_lrelu_bwd uses for `rows` rows: partial holds slots * 2 * N floats
extern "C" int ddsp_b200_ln_lrelu_slots(int64_t rows) { return ln_grid(rows); }

// y = leaky_relu(layer_norm(x)), x, y [rows][N]; stats [rows][2] = mean, rstd (NULL for inference).
// N must be 128, 256, 384 or 512.
extern "C" int ddsp_b200_ln_lrelu_fwd(const float *x, const float *gamma, const float *beta, float *y, float *stats,
                                      int64_t rows, int N, float eps, float slope, void *stream) {
    if (rows == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(x && gamma && beta && y && rows > 0);
    if (N % 128 != 0 || N < 128 || N > 512) return DDSP_B200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ln_grid(rows);
    switch (N / 128) {
        case 1: ln_lrelu_fwd_kernel<1><<<grid, kLnWarps * 32, 0, st>>>(x, gamma, beta, y, stats, rows, eps, slope); break;
        case 2: ln_lrelu_fwd_kernel<2><<<grid, kLnWarps * 32, 0, st>>>(x, gamma, beta, y, stats, rows, eps, slope); break;
        case 3: ln_lrelu_fwd_kernel<3><<<grid, kLnWarps * 32, 0, st>>>(x, gamma, beta, y, stats, rows, eps, slope); break;
        default: ln_lrelu_fwd_kernel<4><<<grid, kLnWarps * 32, 0, st>>>(x, gamma, beta, y, stats, rows, eps, slope); break;
    }
    return ddsp_launch_status();
}

// dx [rows][N], d_gamma, d_beta [N] from dy, x and the forward's stats; partial: ln_lrelu_slots(rows)*2*N floats
extern "C" int ddsp_b200_ln_lrelu_bwd(const float *dy, const float *x, const float *gamma, const float *beta,
                                      const float *stats, float *dx, float *d_gamma, float *d_beta, float *partial,
                                      int64_t rows, int N, float slope, void *stream) {
    DDSP_REQUIRE(dy && x && gamma && beta && stats && dx && d_gamma && d_beta && partial && rows > 0);
    if (N % 128 != 0 || N < 128 || N > 512) return DDSP_B200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ln_grid(rows);
    switch (N / 128) {
        case 1: ln_lrelu_bwd_kernel<1><<<grid, kLnWarps * 32, 0, st>>>(dy, x, gamma, beta, stats, dx, partial, rows, slope); break;
        case 2: ln_lrelu_bwd_kernel<2><<<grid, kLnWarps * 32, 0, st>>>(dy, x, gamma, beta, stats, dx, partial, rows, slope); break;
        case 3: ln_lrelu_bwd_kernel<3><<<grid, kLnWarps * 32, 0, st>>>(dy, x, gamma, beta, stats, dx, partial, rows, slope); break;
        default: ln_lrelu_bwd_kernel<4><<<grid, kLnWarps * 32, 0, st>>>(dy, x, gamma, beta, stats, dx, partial, rows, slope); break;
    }
    int s = ddsp_launch_status();
    if (s) return s;
    ln_param_grad_kernel<<<(2 * N * 8 + 127) / 128, 128, 0, st>>>(partial, d_gamma, d_beta, N, grid);
    return ddsp_launch_status();
}
