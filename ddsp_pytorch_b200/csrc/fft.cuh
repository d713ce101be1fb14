// CTA-cooperative shared-memory FFT (Stockham autosort, radix 4 with one leading radix-2 pass
// when log2(n) is odd).  Natural order in, natural order out, ping-pong between two shared
// buffers.  Used by the STFT / spectral-loss kernels (K4) and by the row / column passes of the
// long convolution (K3).  Twiddles come from a caller-owned table tw[m] = exp(-2 pi i m / n_tab)
// (ddsp_b200_twiddle_table), read through the read-only path with stride n_tab / n.
#pragma once
#include "common.cuh"

// Padding: one float2 every 16 so that the stride-4^p writes of the early passes spread over banks.
__host__ __device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int fpad_size(int n) { return (fpad(n) + 1) & ~1; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <bool INV>
__device__ __forceinline__ float2 tw_load(const float2 *__restrict__ tw, int idx) {
    float2 w = __ldg(tw + idx);
    if (INV) w.y = -w.y;
    return w;
}

// Radix-4 Stockham butterfly jj in [0, n/4) of the pass with sub-transform size Ns (1, 4, 16, ...
// or 2, 8, 32 ... after a leading radix-2 pass).
template <bool INV>
__device__ __forceinline__ void fft_r4(const float2 *in, float2 *out, int n, int Ns, int jj,
                                       const float2 *__restrict__ tw, int tws) {
    const int q = n >> 2;
    const int k = jj & (Ns - 1);
    float2 v0 = in[fpad(jj)], v1 = in[fpad(jj + q)], v2 = in[fpad(jj + 2 * q)], v3 = in[fpad(jj + 3 * q)];
    if (Ns > 1) {
        const int ti = k * (q / Ns) * tws;            // k * n/(4 Ns) table steps
        v1 = cmul(v1, tw_load<INV>(tw, ti));
        v2 = cmul(v2, tw_load<INV>(tw, 2 * ti));
        v3 = cmul(v3, tw_load<INV>(tw, 3 * ti));
    }
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = csub(v1, v3);
    // forward: -i * a3 ; inverse: +i * a3
    const float2 r = INV ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);
    const int j0 = ((jj - k) << 2) + k;
    out[fpad(j0)] = cadd(a0, a2);
    out[fpad(j0 + Ns)] = cadd(a1, r);
    out[fpad(j0 + 2 * Ns)] = csub(a0, a2);
    out[fpad(j0 + 3 * Ns)] = csub(a1, r);
}

__device__ __forceinline__ void fft_r2_first(const float2 *in, float2 *out, int n, int jj) {
    const float2 v0 = in[fpad(jj)], v1 = in[fpad(jj + (n >> 1))];
    out[fpad(2 * jj)] = cadd(v0, v1);
    out[fpad(2 * jj + 1)] = csub(v0, v1);
}

// `batch` independent n-point FFTs; FFT b lives at a + b*pitch (padded indexing), scratch at
// bb + b*pitch.  All nthr threads of the CTA call this after a __syncthreads() that made the
// input visible; it ends with a __syncthreads().  Returns the buffer holding the result.
// INV = unnormalised inverse (e^{+i...}).
template <bool INV>
__device__ __forceinline__ float2 *cta_fft(float2 *a, float2 *bb, int pitch, int batch, int n, int lg,
                                           const float2 *__restrict__ tw, int tws, int tid, int nthr) {
    float2 *src = a, *dst = bb;
    int Ns = 1;
    if (lg & 1) {
        const int half = n >> 1, tot = batch * half;
        for (int w = tid; w < tot; w += nthr) {
            const int b = w >> (lg - 1), jj = w & (half - 1);
            fft_r2_first(src + b * pitch, dst + b * pitch, n, jj);
        }
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
        Ns = 2;
    }
    const int q = n >> 2, tot = batch * q;
    while (Ns < n) {
        for (int w = tid; w < tot; w += nthr) {
            const int b = w >> (lg - 2), jj = w & (q - 1);
            fft_r4<INV>(src + b * pitch, dst + b * pitch, n, Ns, jj, tw, tws);
        }
        __syncthreads();
        float2 *t = src; src = dst; dst = t;
        Ns <<= 2;
    }
    return src;
}

static inline int ddsp_ilog2(int64_t n) {
    int l = 0;
    while (((int64_t)1 << l) < n) ++l;
    return l;
}
