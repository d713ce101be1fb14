// Two-transform ("lane-packed") register FFT for the fused spectral loss (mss_fused.cu).
//
// Same plan as regfft.cuh (T = N/16 threads per transform, 16 points per thread, Stockham autosort in 2-3
// stages), but every value is a PAIR: lane A and lane B are the same point of two independent transforms
// (two STFT frames), held in one 64-bit register pair and processed by the packed sm_100 instructions
// FADD2 / FMUL2 / FFMA2.  The FP32 lane rate of the packed forms equals the scalar rate (probe:
// profiles/r02_fp32_pace_probe.txt); what packing halves is everything else: issue slots, index maths,
// twiddle loads (one scalar twiddle serves both lanes through the .F32 broadcast operand), and the
// shared-memory instructions (one LDS.128 / STS.128 per complex pair).
//
// The functions are per-thread and pure; the kernels place the barriers.  Host build (no __CUDA_ARCH__)
// uses a two-float struct so tests/host/pfft_host_test.cu can check the maths without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "regfft.cuh"

namespace pfft {

using regfft::pad16;
using regfft::Plan;
using regfft::Stage;

// ---- the packed value ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
struct V { uint64_t v; };
DDSP_HD V mk(float a, float b) { V r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
DDSP_HD void get(V x, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
DDSP_HD V operator+(V a, V b) { V r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DDSP_HD V operator-(V a, V b) { V r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DDSP_HD V operator*(V a, V b) { V r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DDSP_HD V fma(V a, V b, V c) { V r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct V { float a, b; };
DDSP_HD V mk(float a, float b) { V r; r.a = a; r.b = b; return r; }
DDSP_HD void get(V x, float &a, float &b) { a = x.a; b = x.b; }
DDSP_HD V operator+(V a, V b) { return mk(a.a + b.a, a.b + b.b); }
DDSP_HD V operator-(V a, V b) { return mk(a.a - b.a, a.b - b.b); }
DDSP_HD V operator*(V a, V b) { return mk(a.a * b.a, a.b * b.b); }
DDSP_HD V fma(V a, V b, V c) { return mk(fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)); }
#endif
DDSP_HD V bc(float s) { return mk(s, s); }                       // ptxas: scalar .F32 operand, no MOV
DDSP_HD V neg(V x) { float a, b; get(x, a, b); return mk(-a, -b); }   // ptxas: operand negation, no instruction
DDSP_HD V fmas(V a, float s, V c) { return fma(a, bc(s), c); }    // a*s + c
DDSP_HD V fnmas(V a, float s, V c) { return fma(a, bc(-s), c); }  // -a*s + c
DDSP_HD V muls(V a, float s) { return a * bc(s); }

struct C { V re, im; };                                           // one complex point of both transforms
DDSP_HD C cadd(C a, C b) { C r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
DDSP_HD C csub(C a, C b) { C r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
// a * (wr + i wi)
DDSP_HD C cmuls(C a, float wr, float wi) {
    C r;
    r.re = fnmas(a.im, wi, muls(a.re, wr));
    r.im = fmas(a.re, wi, muls(a.im, wr));
    return r;
}

// ---- radix-R DFTs, natural order in and out.  INV selects the conjugate kernel. -------------------
// a + (-i) b (forward) / a + (+i) b (inverse), and the matching a - ...
template <bool INV> DDSP_HD C add_rot(C a, C b) {
    C r;
    if (INV) { r.re = a.re - b.im; r.im = a.im + b.re; }
    else { r.re = a.re + b.im; r.im = a.im - b.re; }
    return r;
}
template <bool INV> DDSP_HD C sub_rot(C a, C b) { return add_rot<!INV>(a, b); }

template <bool INV> DDSP_HD void dft2(C &a, C &b) {
    const C t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <bool INV> DDSP_HD void dft4(C &v0, C &v1, C &v2, C &v3) {
    const C a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = csub(v1, v3);
    v0 = cadd(a0, a2);
    v2 = csub(a0, a2);
    v1 = add_rot<INV>(a1, a3);
    v3 = sub_rot<INV>(a1, a3);
}

// e +/- o * W with W = (h, -+h): the common factor h is folded into the butterfly's FMAs.
//   forward: o*W = h((o.re + o.im), (o.im - o.re));  inverse: h((o.re - o.im), (o.im + o.re))
template <bool INV> DDSP_HD void bfly_w8_1(C e, C o, C &p, C &m) {
    const float h = 0.70710678118654752440f;
    const V s = INV ? o.re - o.im : o.re + o.im;
    const V d = INV ? o.im + o.re : o.im - o.re;
    p.re = fmas(s, h, e.re);  p.im = fmas(d, h, e.im);
    m.re = fnmas(s, h, e.re); m.im = fnmas(d, h, e.im);
}
// W = (-h, -+h):  forward: o*W = h((o.im - o.re), -(o.re + o.im));  inverse: h(-(o.re + o.im), (o.re - o.im))
template <bool INV> DDSP_HD void bfly_w8_3(C e, C o, C &p, C &m) {
    const float h = 0.70710678118654752440f;
    const V s = INV ? o.re + o.im : o.im - o.re;       // +h * s on re (inverse: -h)
    const V d = INV ? o.re - o.im : o.re + o.im;       // forward: -h * d on im; inverse: +h * d
    if (INV) {
        p.re = fnmas(s, h, e.re); p.im = fmas(d, h, e.im);
        m.re = fmas(s, h, e.re);  m.im = fnmas(d, h, e.im);
    } else {
        p.re = fmas(s, h, e.re);  p.im = fnmas(d, h, e.im);
        m.re = fnmas(s, h, e.re); m.im = fmas(d, h, e.im);
    }
}
// generic: p = e + o*(wr + i wi), m = e - o*(wr + i wi), four FMAs per output pair
DDSP_HD void bfly_w(C e, C o, float wr, float wi, C &p, C &m) {
    p.re = fnmas(o.im, wi, fmas(o.re, wr, e.re));
    p.im = fmas(o.re, wi, fmas(o.im, wr, e.im));
    m.re = fmas(o.im, wi, fnmas(o.re, wr, e.re));
    m.im = fnmas(o.re, wi, fnmas(o.im, wr, e.im));
}

template <bool INV> DDSP_HD void dft8(C *v) {
    C e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    C o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    v[0] = cadd(e0, o0);               v[4] = csub(e0, o0);
    bfly_w8_1<INV>(e1, o1, v[1], v[5]);
    v[2] = add_rot<INV>(e2, o2);       v[6] = sub_rot<INV>(e2, o2);
    bfly_w8_3<INV>(e3, o3, v[3], v[7]);
}

template <bool INV> DDSP_HD void dft16(C *v) {
    C e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft8<INV>(e);
    dft8<INV>(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float sg = INV ? 1.f : -1.f;                 // W16^k = cos - i sin (forward)
    v[0] = cadd(e[0], o[0]);                     v[8] = csub(e[0], o[0]);
    bfly_w(e[1], o[1], c1, sg * s1, v[1], v[9]);
    bfly_w8_1<INV>(e[2], o[2], v[2], v[10]);
    bfly_w(e[3], o[3], s1, sg * c1, v[3], v[11]);
    v[4] = add_rot<INV>(e[4], o[4]);             v[12] = sub_rot<INV>(e[4], o[4]);
    bfly_w(e[5], o[5], -s1, sg * c1, v[5], v[13]);
    bfly_w8_3<INV>(e[6], o[6], v[6], v[14]);
    bfly_w(e[7], o[7], -c1, sg * s1, v[7], v[15]);
}

template <int R, bool INV> DDSP_HD void dft_r(C *v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else if (R == 16) dft16<INV>(v);
}

// ---- shared-memory element: 16 bytes = (reA, reB, imA, imB) ------------------------------------------
#ifdef __CUDA_ARCH__
typedef ulonglong2 E;
DDSP_HD E to_e(C c) { return make_ulonglong2(c.re.v, c.im.v); }
DDSP_HD C from_e(E e) { C c; c.re.v = e.x; c.im.v = e.y; return c; }
#else
struct E { V re, im; };
DDSP_HD E to_e(C c) { E e; e.re = c.re; e.im = c.im; return e; }
DDSP_HD C from_e(E e) { C c; c.re = e.re; c.im = e.im; return c; }
#endif

// ---- stages (index conventions of regfft.cuh) ------------------------------------------------------
// twiddles + butterflies of stage S in registers; tw = the size's stage table (float2 (cos, -sin), [r-1][k])
template <int LG, int S, bool INV>
DDSP_HD void stage_compute_regs(C (&x)[16], int t, const float2 *__restrict__ tw) {
    using P = Plan<LG>;
    constexpr int R = Stage<LG, S>::R;
    constexpr int NS = Stage<LG, S>::NS;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int j = t + m * P::T;
        const int k = j & (NS - 1);
        if (S > 0) {
            const float2 *tab = tw + (S == 1 ? 0 : regfft::stage_table_offset2<LG>());
#pragma unroll
            for (int r = 1; r < R; ++r) {
#ifdef __CUDA_ARCH__
                const float2 w = __ldg(tab + (r - 1) * NS + k);
#else
                const float2 w = tab[(r - 1) * NS + k];
#endif
                x[m * R + r] = cmuls(x[m * R + r], w.x, INV ? -w.y : w.y);
            }
        }
        dft_r<R, INV>(&x[m * R]);
    }
}

template <int LG, int S, bool INV>
DDSP_HD void stage_compute_store(C (&x)[16], E *buf, int t, const float2 *__restrict__ tw) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
    stage_compute_regs<LG, S, INV>(x, t, tw);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        E *dst = buf + regfft::out_base_padded<LG, S>(t, m);          // base + r * constant (regfft.cuh)
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * regfft::out_stride_padded<LG, S>()] = to_e(x[m * R + r]);
    }
}

template <int LG, int S>
DDSP_HD void stage_load(C (&x)[16], const E *buf, int t) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const E *src = buf + regfft::in_base_padded<LG, S>(t, m);
#pragma unroll
        for (int r = 0; r < R; ++r) x[m * R + r] = from_e(src[r * regfft::in_stride_padded<LG, S>()]);
    }
}

// After the LAST stage, register slot m*R + r of thread t holds output index t + T*q, q = m + M*r
// (R, M of the last stage).  slot_of_q inverts that.
template <int LG> DDSP_HD constexpr int last_R() { return Plan<LG>::STAGES == 3 ? Plan<LG>::R2 : Plan<LG>::R1; }
template <int LG> DDSP_HD constexpr int slot_of_q(int q) {
    return (q % (16 / last_R<LG>())) * last_R<LG>() + q / (16 / last_R<LG>());
}

}  // namespace pfft

// =====================================================================================================
// One transform per thread with (re, im) in the two lanes ("zfft"): the inverse transform of the fused loss.
// ptxas folds the half swap (.LO_HI), the negation of one half (.NP / .PN) and scalar broadcasts into the packed
// operand, so a complex multiply is FMUL2 + FFMA2 and a multiplication by +-i rides in the butterfly's FFMA2:
// half the instructions of the scalar float2 code in regfft.cuh for the same FP32 lane work.
// =====================================================================================================
namespace zfft {

using pfft::V;
using pfft::mk;
using pfft::get;
using pfft::bc;
using pfft::fma;
using regfft::Plan;
using regfft::Stage;

DDSP_HD V swp(V a) { float x, y; get(a, x, y); return mk(y, x); }
DDSP_HD V from_f2(float2 a) { return mk(a.x, a.y); }
DDSP_HD float2 to_f2(V a) { float2 r; get(a, r.x, r.y); return r; }
// a * (wr + i wi)
DDSP_HD V zmul(V a, float wr, float wi) { return fma(swp(a), mk(-wi, wi), a * bc(wr)); }
DDSP_HD V add_i(V a, V b) { return fma(swp(b), mk(-1.f, 1.f), a); }     // a + i b
DDSP_HD V sub_i(V a, V b) { return fma(swp(b), mk(1.f, -1.f), a); }     // a - i b
// a + W4 b with W4 = -i (forward) / +i (inverse)
template <bool INV> DDSP_HD V add_rot(V a, V b) { return INV ? add_i(a, b) : sub_i(a, b); }
template <bool INV> DDSP_HD V sub_rot(V a, V b) { return INV ? sub_i(a, b) : add_i(a, b); }

template <bool INV> DDSP_HD void dft2(V &a, V &b) {
    const V t = a - b;
    a = a + b;
    b = t;
}
template <bool INV> DDSP_HD void dft4(V &v0, V &v1, V &v2, V &v3) {
    const V a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = v1 - v3;
    v0 = a0 + a2;
    v2 = a0 - a2;
    v1 = add_rot<INV>(a1, a3);
    v3 = sub_rot<INV>(a1, a3);
}
// p = e + o W, m = e - o W for W = h (1 -+ i)   (forward: 1 - i)
template <bool INV> DDSP_HD void bfly_w8_1(V e, V o, V &p, V &m) {
    const float h = 0.70710678118654752440f;
    const V s = INV ? add_i(o, o) : sub_i(o, o);        // o (1 +- i)
    p = fma(s, bc(h), e);
    m = fma(s, bc(-h), e);
}
// W = h (-1 -+ i) = -h (1 +- i)   (forward: -1 - i)
template <bool INV> DDSP_HD void bfly_w8_3(V e, V o, V &p, V &m) {
    const float h = 0.70710678118654752440f;
    const V s = INV ? sub_i(o, o) : add_i(o, o);        // o (1 -+ i)
    p = fma(s, bc(-h), e);
    m = fma(s, bc(h), e);
}
DDSP_HD void bfly_w(V e, V o, float wr, float wi, V &p, V &m) {
    p = fma(swp(o), mk(-wi, wi), fma(o, bc(wr), e));
    m = fma(swp(o), mk(wi, -wi), fma(o, bc(-wr), e));
}
template <bool INV> DDSP_HD void dft8(V *v) {
    V e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    V o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    v[0] = e0 + o0;                    v[4] = e0 - o0;
    bfly_w8_1<INV>(e1, o1, v[1], v[5]);
    v[2] = add_rot<INV>(e2, o2);       v[6] = sub_rot<INV>(e2, o2);
    bfly_w8_3<INV>(e3, o3, v[3], v[7]);
}
template <bool INV> DDSP_HD void dft16(V *v) {
    V e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    dft8<INV>(e);
    dft8<INV>(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float sg = INV ? 1.f : -1.f;
    v[0] = e[0] + o[0];                          v[8] = e[0] - o[0];
    bfly_w(e[1], o[1], c1, sg * s1, v[1], v[9]);
    bfly_w8_1<INV>(e[2], o[2], v[2], v[10]);
    bfly_w(e[3], o[3], s1, sg * c1, v[3], v[11]);
    v[4] = add_rot<INV>(e[4], o[4]);             v[12] = sub_rot<INV>(e[4], o[4]);
    bfly_w(e[5], o[5], -s1, sg * c1, v[5], v[13]);
    bfly_w8_3<INV>(e[6], o[6], v[6], v[14]);
    bfly_w(e[7], o[7], -c1, sg * s1, v[7], v[15]);
}
template <int R, bool INV> DDSP_HD void dft_r(V *v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else if (R == 16) dft16<INV>(v);
}

template <int LG, int S, bool INV>
DDSP_HD void stage_compute_regs(V (&x)[16], int t, const float2 *__restrict__ tw) {
    using P = Plan<LG>;
    constexpr int R = Stage<LG, S>::R;
    constexpr int NS = Stage<LG, S>::NS;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int j = t + m * P::T;
        const int k = j & (NS - 1);
        if (S > 0) {
            const float2 *tab = tw + (S == 1 ? 0 : regfft::stage_table_offset2<LG>());
#pragma unroll
            for (int r = 1; r < R; ++r) {
#ifdef __CUDA_ARCH__
                const float2 w = __ldg(tab + (r - 1) * NS + k);
#else
                const float2 w = tab[(r - 1) * NS + k];
#endif
                x[m * R + r] = zmul(x[m * R + r], w.x, INV ? -w.y : w.y);
            }
        }
        dft_r<R, INV>(&x[m * R]);
    }
}
template <int LG, int S, bool INV>
DDSP_HD void stage_compute_store(V (&x)[16], float2 *buf, int t, const float2 *__restrict__ tw) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
    stage_compute_regs<LG, S, INV>(x, t, tw);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 *dst = buf + regfft::out_base_padded<LG, S>(t, m);
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * regfft::out_stride_padded<LG, S>()] = to_f2(x[m * R + r]);
    }
}
template <int LG, int S>
DDSP_HD void stage_load(V (&x)[16], const float2 *buf, int t) {
    constexpr int R = Stage<LG, S>::R;
    constexpr int M = 16 / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float2 *src = buf + regfft::in_base_padded<LG, S>(t, m);
#pragma unroll
        for (int r = 0; r < R; ++r) x[m * R + r] = from_f2(src[r * regfft::in_stride_padded<LG, S>()]);
    }
}

}  // namespace zfft
