// K2: filtered-noise branch (SURVEY 8a rows a7, a8, a9-per-frame and their backward).
//
// Reference path replaced:
//   ddsp/core.py:144-166            amp_to_impulse_response  irfft(128) -> roll -> hann -> pad -> roll
//   ddsp/models/modules.py:116-128  FilteredNoise.forward    IR, uniform noise, fft_convolve, reshape
//   ddsp/core.py:169-176            fft_convolve             rfft/irfft of size 2*block per frame
//
// What the maths is (oracle/closed_form.py: fir_taps, filtered_noise).  With F = 2(NB-1), half = F/2
//   h[n]   = (m_0 + (-1)^n m_{NB-1} + 2 sum_{k=1}^{NB-2} m_k cos(2 pi k n / F)) / F     (irfft, even)
//   IR[d]             = h[d]      * w[d+half]   d = 0..half-1        (w = periodic Hann(F))
//   IR[bs-half+e]     = h[half-e] * w[e]        e = 1..half-1
// and fft_convolve is the causal convolution truncated to the frame: only 2*half-1 = 127 taps are
// non-zero whatever the block size, so the per-frame FFTs (3 of size 2*bs) are replaced by a direct
// FIR from shared memory; the IR never goes to HBM.  Algorithmic HBM bytes: 4*(NB + 2*bs) per frame.
#include "common.cuh"

namespace {

constexpr int kNoiseThreads = 128;

// h[0..half] from mags[0..NB) for one row, cooperatively by `nthr` threads (tid in [0,nthr)).
// Uses cos(2 pi k (half-n)/F) = (-1)^k cos(2 pi k n/F): thread n produces h[n] and h[half-n].
// `ct` is a shared cos table ct[i] = cos(2 pi i / F), i in [0,F).
__device__ __forceinline__ void ir_design_row(const float *m, const float *ct, float *h, int NB,
                                              int tid, int nthr) {
    const int half = NB - 1, F = 2 * half;
    const float invF = 1.f / (float)F;
    for (int n = tid; 2 * n <= half; n += nthr) {
        float ev = 0.f, od = 0.f;
        int idx = 0;                                  // (k*n) mod F
        for (int k = 0; k < NB; ++k) {
            const float ck = (k == 0 || k == half) ? 1.f : 2.f;
            const float t = ck * m[k] * ct[idx];
            if (k & 1) od += t; else ev += t;
            idx += n;
            if (idx >= F) idx -= F;
        }
        h[n] = (ev + od) * invF;
        h[half - n] = (ev - od) * invF;               // same value twice when 2n == half
    }
}

// adjoint of ir_design_row: dm[k] = c_k/F * sum_{n=0}^{half} dh[n] cos(2 pi k n / F)
__device__ __forceinline__ void ir_design_row_bwd(const float *dh, const float *ct, float *dm, int NB,
                                                  int tid, int nthr) {
    const int half = NB - 1, F = 2 * half;
    const float invF = 1.f / (float)F;
    for (int k = tid; k < NB; k += nthr) {
        float acc = 0.f;
        int idx = 0;
        for (int n = 0; n <= half; ++n) {
            acc = fmaf(dh[n], ct[idx], acc);
            idx += k;
            if (idx >= F) idx -= F;
        }
        dm[k] = acc * invF * ((k == 0 || k == half) ? 1.f : 2.f);
    }
}

__device__ __forceinline__ float hann_periodic(int i, int F) {
    return 0.5f - 0.5f * cospif(2.f * (float)i / (float)F);
}

// ---------------------------------------------------------------------------------------------
// a7 standalone: amp[rows,NB] -> ir[rows,target], any target (pad or crop).
//   ir[i] = g'[(i+half) mod target],  g'[j] = (j < min(F,target)) ? h[(j-half) mod F] * w[j] : 0
// ---------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kNoiseThreads)
amp_to_ir_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t rows, int NB,
                 int target) {
    extern __shared__ __align__(16) float smem[];
    const int half = NB - 1, F = 2 * half;
    float *ct = smem;                // [F]
    float *m = ct + F;               // [NB]   mags (fwd) / d_amp staging (bwd)
    float *h = m + NB;               // [half+1]
    const int tid = threadIdx.x;
    for (int i = tid; i < F; i += kNoiseThreads) ct[i] = cospif(2.f * (float)i / (float)F);
    const int lim = min(F, target);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        __syncthreads();
        if (!BWD) {
            for (int k = tid; k < NB; k += kNoiseThreads) m[k] = in[row * NB + k];
            __syncthreads();
            ir_design_row(m, ct, h, NB, tid, kNoiseThreads);
            __syncthreads();
            for (int i = tid; i < target; i += kNoiseThreads) {
                int j = i + half;
                j = j % target;
                float v = 0.f;
                if (j < lim) {
                    int n = j - half;
                    if (n < 0) n += F;
                    n = min(n, F - n);
                    v = h[n] * hann_periodic(j, F);
                }
                out[row * target + i] = v;
            }
        } else {
            for (int n = tid; n <= half; n += kNoiseThreads) h[n] = 0.f;
            __syncthreads();
            for (int i = tid; i < target; i += kNoiseThreads) {
                int j = (i + half) % target;
                if (j < lim) {
                    int n = j - half;
                    if (n < 0) n += F;
                    n = min(n, F - n);
                    atomicAdd(&h[n], in[row * target + i] * hann_periodic(j, F));  // <= 2 adds per slot
                }
            }
            __syncthreads();
            ir_design_row_bwd(h, ct, m, NB, tid, kNoiseThreads);
            __syncthreads();
            for (int k = tid; k < NB; k += kNoiseThreads) out[row * NB + k] = m[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused FilteredNoise.forward.  CTA processes RPC rows (frames) at once; QT = bs/4 threads per row,
// each thread owns 4 consecutive outputs and slides a register window over the noise.
// Shared layout per row: m[NBp] | h[HP] | taps[F] (causal | far) | xs[half + bs] | ys[bs]
// ---------------------------------------------------------------------------------------------
struct NoiseLayout {
    int NBp, HP, F, half, bs, stride;
    __host__ __device__ NoiseLayout(int NB, int bs_) {
        half = NB - 1; F = 2 * half; bs = bs_;
        NBp = (NB + 3) & ~3; HP = (half + 1 + 3) & ~3;
        stride = NBp + HP + F + (half + bs) + bs;
        stride = (stride + 3) & ~3;
    }
    __host__ __device__ int m() const { return 0; }
    __host__ __device__ int h() const { return NBp; }
    __host__ __device__ int taps() const { return NBp + HP; }
    __host__ __device__ int xs() const { return NBp + HP + F; }
    __host__ __device__ int ys() const { return NBp + HP + F + half + bs; }
};

__global__ void __launch_bounds__(kNoiseThreads)
filtered_noise_fwd_kernel(const float *__restrict__ mags, const float *__restrict__ noise,
                          const float *__restrict__ add, float *__restrict__ out, int64_t rows, int NB,
                          int bs, int RPC, int apply_scale, float bias) {
    extern __shared__ __align__(16) float smem[];
    const NoiseLayout L(NB, bs);
    const int half = L.half, F = L.F;
    float *ct = smem;                                   // [F] cos table, then per-row areas
    float *rowbase = smem + ((F + 3) & ~3);
    const int tid = threadIdx.x;
    for (int i = tid; i < F; i += kNoiseThreads) ct[i] = cospif(2.f * (float)i / (float)F);

    const int QT = bs >> 2;                             // threads per row in the FIR phase
    for (int64_t r0 = (int64_t)blockIdx.x * RPC; r0 < rows; r0 += (int64_t)gridDim.x * RPC) {
        const int nr = (int)min((int64_t)RPC, rows - r0);
        __syncthreads();
        // ---- stage mags and noise (coalesced over the CTA's rows)
        for (int i = tid; i < nr * NB; i += kNoiseThreads) {
            const int r = i / NB, k = i - r * NB;
            const float m = __ldg(mags + (r0 + r) * NB + k);
            // modules.py:111-114 folded in: magnitudes = scale_function(raw + initial_bias)
            rowbase[r * L.stride + L.m() + k] = apply_scale ? ddsp_scale_fn(m + bias) : m;
        }
        for (int i = tid; i < nr * (half + bs); i += kNoiseThreads) {
            const int r = i / (half + bs), j = i - r * (half + bs);
            rowbase[r * L.stride + L.xs() + j] = j < half ? 0.f : __ldg(noise + (r0 + r) * bs + (j - half));
        }
        __syncthreads();
        // ---- IR design: split the CTA's threads over the rows
        {
            const int per = kNoiseThreads / nr;
            const int r = tid / per, t = tid - r * per;
            if (r < nr) {
                float *rb = rowbase + r * L.stride;
                ir_design_row(rb + L.m(), ct, rb + L.h(), NB, t, per);
            }
        }
        __syncthreads();
        for (int i = tid; i < nr * F; i += kNoiseThreads) {
            const int r = i / F, d = i - r * F;
            float *rb = rowbase + r * L.stride;
            const float *h = rb + L.h();
            // taps[0..half) causal, taps[half..F) far (far[0] = 0 because w[0] = 0)
            rb[L.taps() + d] = d < half ? h[d] * hann_periodic(d + half, F)
                                        : h[F - d] * hann_periodic(d - half, F);
        }
        __syncthreads();
        // ---- main FIR: y[i..i+3] = sum_d causal[d] * x[i-d]
        for (int w = tid; w < nr * QT; w += kNoiseThreads) {
            const int r = w / QT, i = (w - r * QT) * 4;
            const float *rb = rowbase + r * L.stride;
            const float4 *c4 = reinterpret_cast<const float4 *>(rb + L.taps());
            const float *xs = rb + L.xs() + half + i;        // xs[-d] = x[i-d], zero padded below 0
            float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
            float4 hi = *reinterpret_cast<const float4 *>(xs);   // x[i..i+3]
            for (int d = 0; d < half; d += 4) {
                const float4 lo = *reinterpret_cast<const float4 *>(xs - d - 4);   // x[i-d-4..i-d-1]
                const float4 c = c4[d >> 2];
                // tap d   : x[i-d+r]      = hi.{x,y,z,w}
                y0 = fmaf(c.x, hi.x, y0); y1 = fmaf(c.x, hi.y, y1); y2 = fmaf(c.x, hi.z, y2); y3 = fmaf(c.x, hi.w, y3);
                // tap d+1 : x[i-d-1+r]    = lo.w, hi.x, hi.y, hi.z
                y0 = fmaf(c.y, lo.w, y0); y1 = fmaf(c.y, hi.x, y1); y2 = fmaf(c.y, hi.y, y2); y3 = fmaf(c.y, hi.z, y3);
                // tap d+2
                y0 = fmaf(c.z, lo.z, y0); y1 = fmaf(c.z, lo.w, y1); y2 = fmaf(c.z, hi.x, y2); y3 = fmaf(c.z, hi.y, y3);
                // tap d+3
                y0 = fmaf(c.w, lo.y, y0); y1 = fmaf(c.w, lo.z, y1); y2 = fmaf(c.w, lo.w, y2); y3 = fmaf(c.w, hi.x, y3);
                hi = lo;
            }
            *reinterpret_cast<float4 *>(const_cast<float *>(rb) + L.ys() + i) = make_float4(y0, y1, y2, y3);
        }
        __syncthreads();
        // ---- far taps: y[bs-half+t] += sum_{e=1}^{t} far[e] * x[t-e],  t = 1..half-1
        for (int w = tid; w < nr * half; w += kNoiseThreads) {
            const int r = w / half, t = w - r * half;
            float *rb = rowbase + r * L.stride;
            const float *far = rb + L.taps() + half;
            const float *x = rb + L.xs() + half;
            float acc = 0.f;
            for (int e = 1; e <= t; ++e) acc = fmaf(far[e], x[t - e], acc);
            rb[L.ys() + bs - half + t] += acc;
        }
        __syncthreads();
        for (int i = tid; i < nr * bs; i += kNoiseThreads) {
            const int r = i / bs, j = i - r * bs;
            float y = rowbase[r * L.stride + L.ys() + j];
            if (add) y += __ldg(add + (r0 + r) * bs + j);       // decoder.py:121: harmonic + noise
            out[(r0 + r) * bs + j] = y;
        }
    }
}

// Backward: d_mags from g (grad of the frame's output) and the same noise.
//   d causal[d] = sum_{i>=d} g[i] x[i-d];  d far[e] = sum_{t>=e} g[bs-half+t] x[t-e]
// ys area is reused for g; taps area receives d taps; h area receives dh.
__global__ void __launch_bounds__(kNoiseThreads)
filtered_noise_bwd_kernel(const float *__restrict__ g_out, const float *__restrict__ noise,
                          const float *__restrict__ mags_raw, float *__restrict__ d_mags, int64_t rows,
                          int NB, int bs, int RPC, int apply_scale, float bias) {
    extern __shared__ __align__(16) float smem[];
    const NoiseLayout L(NB, bs);
    const int half = L.half, F = L.F;
    float *ct = smem;
    float *rowbase = smem + ((F + 3) & ~3);
    const int tid = threadIdx.x;
    for (int i = tid; i < F; i += kNoiseThreads) ct[i] = cospif(2.f * (float)i / (float)F);

    for (int64_t r0 = (int64_t)blockIdx.x * RPC; r0 < rows; r0 += (int64_t)gridDim.x * RPC) {
        const int nr = (int)min((int64_t)RPC, rows - r0);
        __syncthreads();
        for (int i = tid; i < nr * (half + bs); i += kNoiseThreads) {
            const int r = i / (half + bs), j = i - r * (half + bs);
            rowbase[r * L.stride + L.xs() + j] = j < half ? 0.f : __ldg(noise + (r0 + r) * bs + (j - half));
        }
        for (int i = tid; i < nr * bs; i += kNoiseThreads) {
            const int r = i / bs, j = i - r * bs;
            rowbase[r * L.stride + L.ys() + j] = __ldg(g_out + (r0 + r) * bs + j);
        }
        __syncthreads();
        // d taps: one thread per (row, tap); causal taps walk the whole frame, far taps a triangle
        for (int w = tid; w < nr * F; w += kNoiseThreads) {
            const int r = w / F, d = w - r * F;
            float *rb = rowbase + r * L.stride;
            const float *g = rb + L.ys();
            const float *x = rb + L.xs() + half;             // x[-1..-half] are zeros
            float a0 = 0.f, a1 = 0.f;
            if (d < half) {
                int i = 0;
                for (; i + 1 < bs; i += 2) {
                    a0 = fmaf(g[i], x[i - d], a0);
                    a1 = fmaf(g[i + 1], x[i + 1 - d], a1);
                }
                if (i < bs) a0 = fmaf(g[i], x[i - d], a0);
            } else {
                const int e = d - half;
                if (e >= 1)
                    for (int t = e; t < half; ++t) a0 = fmaf(g[bs - half + t], x[t - e], a0);
            }
            rb[L.taps() + d] = a0 + a1;
        }
        __syncthreads();
        // dh[n] = d causal[n] w[n+half] (n < half)  +  d far[half-n] w[half-n] (1 <= half-n < half)
        for (int w = tid; w < nr * (half + 1); w += kNoiseThreads) {
            const int r = w / (half + 1), n = w - r * (half + 1);
            float *rb = rowbase + r * L.stride;
            const float *dt = rb + L.taps();
            float v = 0.f;
            if (n < half) v = dt[n] * hann_periodic(n + half, F);
            const int e = half - n;
            if (e >= 1 && e < half) v = fmaf(dt[half + e], hann_periodic(e, F), v);
            rb[L.h() + n] = v;
        }
        __syncthreads();
        {
            const int per = kNoiseThreads / nr;
            const int r = tid / per, t = tid - r * per;
            if (r < nr) {
                float *rb = rowbase + r * L.stride;
                ir_design_row_bwd(rb + L.h(), ct, rb + L.m(), NB, t, per);
            }
        }
        __syncthreads();
        for (int i = tid; i < nr * NB; i += kNoiseThreads) {
            const int r = i / NB, k = i - r * NB;
            float d = rowbase[r * L.stride + L.m() + k];
            if (apply_scale) d *= ddsp_scale_grad(__ldg(mags_raw + (r0 + r) * NB + k) + bias);
            d_mags[(r0 + r) * NB + k] = d;
        }
    }
}

inline int noise_rows_per_cta(int bs) {
    int rpc = (kNoiseThreads * 4) / bs;
    if (rpc < 1) rpc = 1;
    if (rpc > 8) rpc = 8;
    return rpc;
}

}  // namespace

extern "C" int ddsp_b200_amp_to_ir_fwd(const float *amp, float *ir, int64_t rows, int NB, int target,
                                       void *stream) {
    DDSP_REQUIRE(amp && ir && rows >= 0 && NB >= 2 && target >= 1);
    if (rows == 0) return DDSP_B200_OK;
    const size_t smem = (size_t)(2 * (NB - 1) + NB + NB) * sizeof(float);
    if (smem > 48 * 1024) return DDSP_B200_EUNSUPPORTED;
    const int grid = (int)(rows < DDSP_SM_COUNT * 16 ? rows : DDSP_SM_COUNT * 16);
    amp_to_ir_kernel<false><<<grid, kNoiseThreads, smem, (cudaStream_t)stream>>>(amp, ir, rows, NB,
                                                                                 target);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_amp_to_ir_bwd(const float *d_ir, float *d_amp, int64_t rows, int NB,
                                       int target, void *stream) {
    DDSP_REQUIRE(d_ir && d_amp && rows >= 0 && NB >= 2 && target >= 1);
    if (rows == 0) return DDSP_B200_OK;
    const size_t smem = (size_t)(2 * (NB - 1) + NB + NB) * sizeof(float);
    if (smem > 48 * 1024) return DDSP_B200_EUNSUPPORTED;
    const int grid = (int)(rows < DDSP_SM_COUNT * 16 ? rows : DDSP_SM_COUNT * 16);
    amp_to_ir_kernel<true><<<grid, kNoiseThreads, smem, (cudaStream_t)stream>>>(d_ir, d_amp, rows,
                                                                                NB, target);
    return ddsp_launch_status();
}

static int noise_launch_cfg(int64_t rows, int NB, int bs, int *rpc, size_t *smem, int *grid) {
    if (NB < 3 || (NB - 1) % 4 != 0) return DDSP_B200_EUNSUPPORTED;     // half % 4 == 0 for LDS.128 taps
    if (bs % 4 != 0 || bs < 2 * (NB - 1)) return DDSP_B200_EUNSUPPORTED;  // see header
    *rpc = noise_rows_per_cta(bs);
    const NoiseLayout L(NB, bs);
    *smem = ((size_t)((L.F + 3) & ~3) + (size_t)*rpc * L.stride) * sizeof(float);
    if (*smem > 200 * 1024) return DDSP_B200_EUNSUPPORTED;
    const int64_t ctas = ddsp_ceil_div(rows, *rpc);
    *grid = (int)(ctas < DDSP_SM_COUNT * 8 ? ctas : DDSP_SM_COUNT * 8);
    return DDSP_B200_OK;
}

extern "C" int ddsp_b200_filtered_noise_fwd(const float *mags, const float *noise, const float *add,
                                            float *out, int64_t rows, int NB, int block_size,
                                            int apply_scale, float bias, void *stream) {
    DDSP_REQUIRE(mags && noise && out && rows >= 0);
    if (rows == 0) return DDSP_B200_OK;
    int rpc, grid;
    size_t smem;
    int s = noise_launch_cfg(rows, NB, block_size, &rpc, &smem, &grid);
    if (s) return s;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(filtered_noise_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    filtered_noise_fwd_kernel<<<grid, kNoiseThreads, smem, (cudaStream_t)stream>>>(
        mags, noise, add, out, rows, NB, block_size, rpc, apply_scale, bias);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_filtered_noise_bwd(const float *g_out, const float *noise,
                                            const float *mags_raw, float *d_mags, int64_t rows, int NB,
                                            int block_size, int apply_scale, float bias, void *stream) {
    DDSP_REQUIRE(g_out && noise && d_mags && rows >= 0 && (!apply_scale || mags_raw));
    if (rows == 0) return DDSP_B200_OK;
    int rpc, grid;
    size_t smem;
    int s = noise_launch_cfg(rows, NB, block_size, &rpc, &smem, &grid);
    if (s) return s;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(filtered_noise_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    filtered_noise_bwd_kernel<<<grid, kNoiseThreads, smem, (cudaStream_t)stream>>>(
        g_out, noise, mags_raw, d_mags, rows, NB, block_size, rpc, apply_scale, bias);
    return ddsp_launch_status();
}
