// Frame-rate control maps (SURVEY 8a rows a1, a2, a3 and their backward).
//
// Reference path replaced:
//   ddsp/core.py:77-78               scale_function          2*sigmoid(x)^ln10 + 1e-7
//   ddsp/core.py:70-74               remove_above_nyquist    amp * ((f0*k < sr/2).float() + 1e-4)
//   ddsp/models/modules.py:44-67     HarmonicSynth.get_controls  (both of the above + /= sum)
// ~15 eager elementwise/reduction launches at frame rate become one launch; one warp owns one
// (voice, frame) row so the normalising sum is a shuffle reduction.
#include "common.cuh"

namespace {

__device__ __forceinline__ float scale_fn(float x) { return ddsp_scale_fn(x); }
__device__ __forceinline__ float scale_grad(float x) { return ddsp_scale_grad(x); }
__device__ __forceinline__ float nyquist_mask(float f0, int k1, float nyq) { return ddsp_nyquist_mask(f0, k1, nyq); }

__global__ void scale_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        y[i] = scale_fn(x[i]);
}
__global__ void scale_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                 float *__restrict__ dx, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * scale_grad(x[i]);
}
__global__ void nyquist_kernel(const float *__restrict__ amp, const float *__restrict__ f0,
                               float *__restrict__ out, int64_t rows, int H, float nyq) {
    const int64_t n = rows * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / H;
        const int k = (int)(i - r * H);
        out[i] = amp[i] * nyquist_mask(f0[r], k + 1, nyq);
    }
}

constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32)
controls_fwd_kernel(const float *__restrict__ amp_raw, const float *__restrict__ dist_raw,
                    const float *__restrict__ f0, float *__restrict__ amps,
                    float *__restrict__ dist, float *__restrict__ weights, int64_t rows, int H,
                    float nyq) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float f = f0[row];
    const float *dr = dist_raw + row * H;
    float *dn = dist + row * H;
    // unnormalised values stay in registers for rows up to kKeepF * 32 harmonics (longer rows park them in `dist`)
    constexpr int kKeepF = 8;
    float vk[kKeepF];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kKeepF; ++i) {
        const int k = lane + 32 * i;
        vk[i] = k < H ? scale_fn(dr[k]) * nyquist_mask(f, k + 1, nyq) : 0.f;
        sum += vk[i];
    }
    for (int k = lane + 32 * kKeepF; k < H; k += 32) {
        const float v = scale_fn(dr[k]) * nyquist_mask(f, k + 1, nyq);
        dn[k] = v;
        sum += v;
    }
    sum = ddsp_warp_sum(sum);
    const float amp = scale_fn(amp_raw[row]);
#pragma unroll
    for (int i = 0; i < kKeepF; ++i) {
        const int k = lane + 32 * i;
        if (k < H) {
            const float n = vk[i] / sum;
            dn[k] = n;
            if (weights) weights[row * H + k] = n * amp;            // modules.py:73: distribution *= amplitudes
        }
    }
    for (int k = lane + 32 * kKeepF; k < H; k += 32) {
        const float n = dn[k] / sum;                            // same thread re-reads its own write
        dn[k] = n;
        if (weights) weights[row * H + k] = n * amp;
    }
    if (lane == 0) amps[row] = amp;
}

__global__ void __launch_bounds__(kRowWarps * 32)
controls_bwd_kernel(const float *__restrict__ amp_raw, const float *__restrict__ dist_raw,
                    const float *__restrict__ f0, const float *__restrict__ d_amps,
                    const float *__restrict__ d_dist, const float *__restrict__ d_weights,
                    float *__restrict__ d_amp_raw, float *__restrict__ d_dist_raw, int64_t rows, int H,
                    float nyq) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float f = f0[row];
    const float *dr = dist_raw + row * H;
    const float *gd = d_dist ? d_dist + row * H : nullptr;
    const float *gw = d_weights ? d_weights + row * H : nullptr;
    float *out = d_dist_raw + row * H;
    float amp, amp_grad;
    ddsp_scale_fn_grad(amp_raw[row], &amp, &amp_grad);
    // n_k = v_k / S;  total gradient on n_k:  g_k = d_dist_k + d_weights_k * amp
    //   dv_k = (g_k - sum_j g_j n_j) / S ;   d amp = d_amps + sum_k d_weights_k n_k
    // value and derivative of scale_function are evaluated together, once per element, and kept in registers for
    // the second pass (rows up to kKeep * 32 harmonics; longer rows re-evaluate)
    constexpr int kKeep = 8;
    float gk[kKeep], dk[kKeep];
    float sum = 0.f, dot = 0.f, wdot = 0.f;
#pragma unroll
    for (int i = 0; i < kKeep; ++i) {
        const int k = lane + 32 * i;
        gk[i] = 0.f;
        dk[i] = 0.f;
        if (k < H) {
            float fn, gr;
            ddsp_scale_fn_grad(dr[k], &fn, &gr);
            const float mask = nyquist_mask(f, k + 1, nyq);
            const float v = fn * mask;
            const float g = (gd ? gd[k] : 0.f) + (gw ? gw[k] * amp : 0.f);
            sum += v;
            dot = fmaf(g, v, dot);
            if (gw) wdot = fmaf(gw[k], v, wdot);
            gk[i] = g;
            dk[i] = mask * gr;
        }
    }
    for (int k = lane + 32 * kKeep; k < H; k += 32) {
        const float v = scale_fn(dr[k]) * nyquist_mask(f, k + 1, nyq);
        const float g = (gd ? gd[k] : 0.f) + (gw ? gw[k] * amp : 0.f);
        sum += v;
        dot = fmaf(g, v, dot);
        if (gw) wdot = fmaf(gw[k], v, wdot);
    }
    sum = ddsp_warp_sum(sum);
    const float inv = 1.f / sum;
    dot = ddsp_warp_sum(dot) * inv;                   // sum_j g_j n_j
    wdot = ddsp_warp_sum(wdot) * inv;                 // sum_k d_weights_k n_k
#pragma unroll
    for (int i = 0; i < kKeep; ++i) {
        const int k = lane + 32 * i;
        if (k < H) out[k] = (gk[i] - dot) * inv * dk[i];
    }
    for (int k = lane + 32 * kKeep; k < H; k += 32) {
        const float g = (gd ? gd[k] : 0.f) + (gw ? gw[k] * amp : 0.f);
        out[k] = (g - dot) * inv * nyquist_mask(f, k + 1, nyq) * scale_grad(dr[k]);
    }
    if (lane == 0) d_amp_raw[row] = ((d_amps ? d_amps[row] : 0.f) + wdot) * amp_grad;
}

inline int ew_blocks(int64_t n) {
    int64_t b = ddsp_ceil_div(n, 256);
    const int64_t cap = (int64_t)DDSP_SM_COUNT * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int ddsp_b200_scale_function_fwd(const float *x, float *y, int64_t n, void *stream) {
    if (n == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(x && y && n > 0);
    scale_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_scale_function_bwd(const float *x, const float *dy, float *dx, int64_t n,
                                            void *stream) {
    if (n == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(x && dy && dx && n > 0);
    scale_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, n);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_remove_above_nyquist(const float *amp, const float *f0, float *out,
                                              int64_t rows, int H, float sample_rate,
                                              void *stream) {
    if (rows == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(amp && f0 && out && rows > 0 && H > 0);
    nyquist_kernel<<<ew_blocks(rows * H), 256, 0, (cudaStream_t)stream>>>(amp, f0, out, rows, H,
                                                                           sample_rate * 0.5f);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_harmonic_controls_fwd(const float *amp_raw, const float *dist_raw,
                                               const float *f0, float *amps, float *dist,
                                               float *weights, int64_t rows, int H, float sample_rate,
                                               void *stream) {
    if (rows == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(amp_raw && dist_raw && f0 && amps && dist && rows > 0 && H > 0);
    controls_fwd_kernel<<<(unsigned)ddsp_ceil_div(rows, kRowWarps), kRowWarps * 32, 0,
                          (cudaStream_t)stream>>>(amp_raw, dist_raw, f0, amps, dist, weights, rows,
                                                  H, sample_rate * 0.5f);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_harmonic_controls_bwd(const float *amp_raw, const float *dist_raw,
                                               const float *f0, const float *d_amps,
                                               const float *d_dist, const float *d_weights,
                                               float *d_amp_raw, float *d_dist_raw, int64_t rows, int H,
                                               float sample_rate, void *stream) {
    if (rows == 0) return DDSP_B200_OK;
    DDSP_REQUIRE(amp_raw && dist_raw && f0 && d_amp_raw && d_dist_raw);
    DDSP_REQUIRE(rows > 0 && H > 0);
    controls_bwd_kernel<<<(unsigned)ddsp_ceil_div(rows, kRowWarps), kRowWarps * 32, 0,
                          (cudaStream_t)stream>>>(amp_raw, dist_raw, f0, d_amps, d_dist, d_weights,
                                                  d_amp_raw, d_dist_raw, rows, H, sample_rate * 0.5f);
    return ddsp_launch_status();
}
