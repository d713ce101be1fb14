// Register-tiled FFT for the STFT / spectral-loss kernels (sizes 64 .. 4096, compile time).
//
// A transform of N = 2^LG points is done by T = N/16 threads that each keep 16 complex values in
// registers.  Stockham autosort in two or three stages (radix 16, then 16 / N/16, then N/256):
// a stage is "load 16 values (strided) -> twiddle -> radix-R butterflies in registers -> store at the
// autosort positions" with ONE shared-memory round trip between stages, against one per radix-4
// pass in the generic cta_fft (fft.cuh).  Stage 0 takes its inputs from the caller (global memory,
// already windowed) and needs no twiddles.  Everything is a pure per-thread function; the kernels
// place the __syncthreads() between phases.  The same functions compile for the host so that
// tests/host/regfft_host_test.cu can check them without a GPU.
#pragma once
#include <cuda_runtime.h>

#ifndef DDSP_HD
#define DDSP_HD __host__ __device__ __forceinline__
#endif

namespace regfft {

DDSP_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DDSP_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DDSP_HD float2 c_mul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV> DDSP_HD float2 c_rot(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

DDSP_HD int pad16(int i) { return i + (i >> 4); }                  // one float2 of padding every 16

// ---- radix-R DFTs in registers, natural order in and out, stride S inside v ---------------------
template <bool INV> DDSP_HD void dft2(float2 &a, float2 &b) {
    const float2 t = c_sub(a, b);
    a = c_add(a, b);
    b = t;
}

template <bool INV> DDSP_HD void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
    const float2 a0 = c_add(v0, v2), a1 = c_sub(v0, v2), a2 = c_add(v1, v3), a3 = c_rot<INV>(c_sub(v1, v3));
    v0 = c_add(a0, a2);
    v1 = c_add(a1, a3);
    v2 = c_sub(a0, a2);
    v3 = c_sub(a1, a3);
}

template <bool INV> DDSP_HD void dft8(float2 *v) {      // v[0..7]
    // even / odd halves
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    // W8^1 = (1 -+ i)/sqrt2, W8^2 = -+ i, W8^3 = (-1 -+ i)/sqrt2   (upper sign forward)
    const float2 t1 = INV ? make_float2(h * (o1.x - o1.y), h * (o1.x + o1.y))
                          : make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    const float2 t2 = c_rot<INV>(o2);
    const float2 t3 = INV ? make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y))
                          : make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = c_add(e0, o0);  v[4] = c_sub(e0, o0);
    v[1] = c_add(e1, t1);  v[5] = c_sub(e1, t1);
    v[2] = c_add(e2, t2);  v[6] = c_sub(e2, t2);
    v[3] = c_add(e3, t3);  v[7] = c_sub(e3, t3);
}

template <bool INV> DDSP_HD void dft16(float2 *v) {     // v[0..15]
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft8<INV>(e);
    dft8<INV>(o);
    // W16^k, k = 1..7: cos/sin of k*pi/8
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    const float cw[8] = {1.f, c1, h, s1, 0.f, -s1, -h, -c1};
    const float sw[8] = {0.f, s1, h, c1, 1.f, c1, h, s1};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // forward: W = cw - i sw ; inverse: W = cw + i sw
        const float wi = INV ? sw[k] : -sw[k];
        const float2 t = make_float2(o[k].x * cw[k] - o[k].y * wi, o[k].x * wi + o[k].y * cw[k]);
        v[k] = c_add(e[k], t);
        v[k + 8] = c_sub(e[k], t);
    }
}

template <int R, bool INV> DDSP_HD void dft_r(float2 *v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else if (R == 16) dft16<INV>(v);
}

// ---- plan ----------------------------------------------------------------------------------------
template <int LG> struct Plan {
    static_assert(LG >= 6 && LG <= 12, "register FFT covers 64..4096 points");
    static constexpr int N = 1 << LG;
    static constexpr int T = N / 16;                          // threads per transform
    static constexpr int STAGES = LG <= 8 ? 2 : 3;
    static constexpr int R0 = 16;
    static constexpr int R1 = LG <= 8 ? N / 16 : 16;
    static constexpr int R2 = LG <= 8 ? 1 : N / 256;
    static constexpr int PITCH = (N + (N >> 4) + 1) & ~1;     // float2 per transform buffer (padded)
};
template <int LG, int S> struct Stage {
    static constexpr int R = S == 0 ? Plan<LG>::R0 : (S == 1 ? Plan<LG>::R1 : Plan<LG>::R2);
    static constexpr int NS = S == 0 ? 1 : (S == 1 ? 16 : 256);
};

// Per-size twiddle table ("stage table"), built by ddsp_b200_stft_stage_twiddles:
//   stage 1: (R1-1) x 16  entries  exp(-2 pi i r k / (16 R1)),   [r-1][k]
//   stage 2: (R2-1) x 256 entries  exp(-2 pi i r k / (256 R2)),  [r-1][k]   (3-stage sizes only)
// A warp's lanes hold consecutive k, so every table read is one or two wavefronts (the generic
// exp(-2 pi i m / N) table read with stride r*k*... touched 16 cache lines per read).
template <int LG> DDSP_HD constexpr int stage_table_offset2() { return (Plan<LG>::R1 - 1) * 16; }
template <int LG> DDSP_HD constexpr int stage_table_size() {
    return (Plan<LG>::R1 - 1) * 16 + (Plan<LG>::STAGES == 3 ? (Plan<LG>::R2 - 1) * 256 : 0);
}

// Stage S of the transform, for thread t (0 <= t < T) of the group working on `buf`.
// On entry x[m*R + r] = stage input (j_m + r*N/R), j_m = t + m*T, m < 16/R.
// Applies the stage's twiddles and butterflies and stores to the autosort positions of buf.
// tw = the stage table of this size.
struct SmemStore {           // default sink of a stage: the transform's padded shared buffer
    float2 *buf;
    DDSP_HD void operator()(int idx, float2 v) const { buf[pad16(idx)] = v; }
};

// twiddle + butterflies of stage S, results left in x: x[m*R + r] is output (j_m - k)*R + k + r*NS
template <int LG, int S, bool INV>
DDSP_HD void stage_compute_regs(float2 (&x)[16], int t, const float2 *tw) {
    using P = Plan<LG>;
    constexpr int R = Stage<LG, S>::R;
    constexpr int NS = Stage<LG, S>::NS;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int j = t + m * P::T;
        const int k = j & (NS - 1);
        if (S > 0) {
            // stage table laid out [r-1][k]: lanes (consecutive k) read consecutive entries
            const float2 *tab = tw + (S == 1 ? 0 : stage_table_offset2<LG>());
#pragma unroll
            for (int r = 1; r < R; ++r) {
#ifdef __CUDA_ARCH__
                float2 w = __ldg(tab + (r - 1) * NS + k);
#else
                float2 w = tab[(r - 1) * NS + k];
#endif
                if (INV) w.y = -w.y;
                x[m * R + r] = c_mul(x[m * R + r], w);
            }
        }
        dft_r<R, INV>(&x[m * R]);
    }
}

// output index of register slot m*R + r of stage S for thread t
template <int LG, int S>
DDSP_HD int stage_out_index(int t, int m, int r) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int NS = Stage<LG, S>::NS;
    const int j = t + m * Plan<LG>::T;
    const int k = j & (NS - 1);
    return (j - k) * R + k + r * NS;
}

template <int LG, int S, bool INV, typename Store>
DDSP_HD void stage_compute_sink(float2 (&x)[16], int t, const float2 *tw, const Store &store) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
    stage_compute_regs<LG, S, INV>(x, t, tw);
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int r = 0; r < R; ++r) store(stage_out_index<LG, S>(t, m, r), x[m * R + r]);
}

// Padded shared-memory addresses as "base(t, m) + r * constant": the strides between a thread's R accesses of one
// butterfly are multiples of 16 elements (or the accesses stay inside one run of 16), so the padding term splits
// off exactly and the compiler addresses all R accesses from one register with immediate offsets.
//   loads  (stage S >= 1): pad16(j + r*N/R)                  = pad16(j)            + r * (N/R) * 17/16
//   stores stage 0       : pad16(16 j + r)                   = 17 j                + r
//          stage 1       : pad16((j-k) R + k + 16 r), k<16   = pad16((j-k) R) + k  + r * 17
//          stage 2       : pad16((j-k) R + k + 256 r), k<256 = pad16((j-k) R + k)  + r * 272
template <int LG, int S> DDSP_HD int in_base_padded(int t, int m) { return pad16(t + m * Plan<LG>::T); }
template <int LG, int S> DDSP_HD constexpr int in_stride_padded() {
    return (Plan<LG>::N / Stage<LG, S>::R) / 16 * 17;
}
template <int LG, int S> DDSP_HD int out_base_padded(int t, int m) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int NS = Stage<LG, S>::NS;
    const int j = t + m * Plan<LG>::T;
    const int k = j & (NS - 1);
    return S == 0 ? 17 * j : pad16((j - k) * R + k);     // stage 1: (j-k) R is a multiple of 16, k < 16
}
template <int LG, int S> DDSP_HD constexpr int out_stride_padded() { return S == 0 ? 1 : (S == 1 ? 17 : 272); }
static_assert(Plan<6>::N / Stage<6, 1>::R % 16 == 0 && Plan<9>::N / Stage<9, 2>::R % 16 == 0, "strides are multiples of 16");

template <int LG, int S, bool INV>
DDSP_HD void stage_compute_store(float2 (&x)[16], float2 *buf, int t, const float2 *tw) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
    stage_compute_regs<LG, S, INV>(x, t, tw);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 *dst = buf + out_base_padded<LG, S>(t, m);
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * out_stride_padded<LG, S>()] = x[m * R + r];
    }
}

// Load the inputs of stage S (S >= 1) from buf into registers.
template <int LG, int S>
DDSP_HD void stage_load(float2 (&x)[16], const float2 *buf, int t) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float2 *src = buf + in_base_padded<LG, S>(t, m);
#pragma unroll
        for (int r = 0; r < R; ++r) x[m * R + r] = src[r * in_stride_padded<LG, S>()];
    }
}

}  // namespace regfft
