// Host-side check of the lane-packed register FFT (ddsp_pytorch_b200/csrc/pfft.cuh): two independent
// transforms ride in the two lanes of every value.  The per-thread stage functions are run thread by
// thread, phase by phase (what the barriers separate on the GPU) and both lanes are compared with a
// float64 DFT.  The last stage is left in registers and read through slot_of_q, the way the fused loss
// kernel consumes it (thread t, slot q  <->  output index t + q*T).  No GPU needed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pfft.cuh"

using namespace pfft;

template <int LG, bool INV>
double check() {
    using P = Plan<LG>;
    const int N = P::N, T = P::T;
    std::vector<float2> tw(regfft::stage_table_size<LG>() + 1);
    for (int r = 1; r < P::R1; ++r)
        for (int k = 0; k < 16; ++k) {
            const double a = -2 * M_PI * r * k / (16.0 * P::R1);
            tw[(r - 1) * 16 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    if (P::STAGES == 3)
        for (int r = 1; r < P::R2; ++r)
            for (int k = 0; k < 256; ++k) {
                const double a = -2 * M_PI * r * k / (256.0 * P::R2);
                tw[regfft::stage_table_offset2<LG>() + (r - 1) * 256 + k] = make_float2((float)cos(a), (float)sin(a));
            }
    std::vector<float2> inA(N), inB(N);
    std::vector<E> buf(P::PITCH);
    srand(LG * 2 + INV);
    for (auto &v : inA) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    for (auto &v : inB) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    struct Regs { C v[16]; };
    std::vector<Regs> x(T);
    for (int t = 0; t < T; ++t)
        for (int r = 0; r < 16; ++r) {
            x[t].v[r].re = mk(inA[t + r * T].x, inB[t + r * T].x);
            x[t].v[r].im = mk(inA[t + r * T].y, inB[t + r * T].y);
        }
    for (int t = 0; t < T; ++t) stage_compute_store<LG, 0, INV>(x[t].v, buf.data(), t, tw.data());
    for (int t = 0; t < T; ++t) stage_load<LG, 1>(x[t].v, buf.data(), t);
    if (P::STAGES == 3) {
        for (int t = 0; t < T; ++t) stage_compute_store<LG, 1, INV>(x[t].v, buf.data(), t, tw.data());
        for (int t = 0; t < T; ++t) stage_load<LG, 2>(x[t].v, buf.data(), t);
        for (int t = 0; t < T; ++t) stage_compute_regs<LG, 2, INV>(x[t].v, t, tw.data());
    } else {
        for (int t = 0; t < T; ++t) stage_compute_regs<LG, 1, INV>(x[t].v, t, tw.data());
    }
    double worst = 0, scale = 0;
    for (int k = 0; k < N; ++k) {
        const int t = k % T, q = k / T;
        const C got = x[t].v[slot_of_q<LG>(q)];
        float gra, grb, gia, gib;
        get(got.re, gra, grb);
        get(got.im, gia, gib);
        for (int lane = 0; lane < 2; ++lane) {
            const std::vector<float2> &in = lane ? inB : inA;
            double re = 0, im = 0;
            for (int n = 0; n < N; ++n) {
                const double a = (INV ? 2 : -2) * M_PI * (double)((long long)k * n % N) / N;
                re += in[n].x * cos(a) - in[n].y * sin(a);
                im += in[n].x * sin(a) + in[n].y * cos(a);
            }
            const double gr = lane ? grb : gra, gi = lane ? gib : gia;
            worst = fmax(worst, fmax(fabs(gr - re), fabs(gi - im)));
            scale = fmax(scale, fmax(fabs(re), fabs(im)));
        }
    }
    printf("N=%4d %s  max abs err %.3e  (max |X| %.2f)\n", N, INV ? "inverse" : "forward", worst, scale);
    return worst / scale;
}

// (re, im)-lane transform (zfft): one transform per "thread", same stage walk
template <int LG, bool INV>
double check_z() {
    using P = Plan<LG>;
    const int N = P::N, T = P::T;
    std::vector<float2> tw(regfft::stage_table_size<LG>() + 1);
    for (int r = 1; r < P::R1; ++r)
        for (int k = 0; k < 16; ++k) {
            const double a = -2 * M_PI * r * k / (16.0 * P::R1);
            tw[(r - 1) * 16 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    if (P::STAGES == 3)
        for (int r = 1; r < P::R2; ++r)
            for (int k = 0; k < 256; ++k) {
                const double a = -2 * M_PI * r * k / (256.0 * P::R2);
                tw[regfft::stage_table_offset2<LG>() + (r - 1) * 256 + k] = make_float2((float)cos(a), (float)sin(a));
            }
    std::vector<float2> in(N), buf(P::PITCH);
    srand(100 + LG * 2 + INV);
    for (auto &v : in) v = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    struct Regs { V v[16]; };
    std::vector<Regs> x(T);
    for (int t = 0; t < T; ++t)
        for (int r = 0; r < 16; ++r) x[t].v[r] = zfft::from_f2(in[t + r * T]);
    for (int t = 0; t < T; ++t) zfft::stage_compute_store<LG, 0, INV>(x[t].v, buf.data(), t, tw.data());
    for (int t = 0; t < T; ++t) zfft::stage_load<LG, 1>(x[t].v, buf.data(), t);
    if (P::STAGES == 3) {
        for (int t = 0; t < T; ++t) zfft::stage_compute_store<LG, 1, INV>(x[t].v, buf.data(), t, tw.data());
        for (int t = 0; t < T; ++t) zfft::stage_load<LG, 2>(x[t].v, buf.data(), t);
        for (int t = 0; t < T; ++t) zfft::stage_compute_regs<LG, 2, INV>(x[t].v, t, tw.data());
    } else {
        for (int t = 0; t < T; ++t) zfft::stage_compute_regs<LG, 1, INV>(x[t].v, t, tw.data());
    }
    double worst = 0, scale = 0;
    for (int k = 0; k < N; ++k) {
        const float2 got = zfft::to_f2(x[k % T].v[slot_of_q<LG>(k / T)]);
        double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const double a = (INV ? 2 : -2) * M_PI * (double)((long long)k * n % N) / N;
            re += in[n].x * cos(a) - in[n].y * sin(a);
            im += in[n].x * sin(a) + in[n].y * cos(a);
        }
        worst = fmax(worst, fmax(fabs(got.x - re), fabs(got.y - im)));
        scale = fmax(scale, fmax(fabs(re), fabs(im)));
    }
    printf("N=%4d %s (re,im lanes)  max abs err %.3e  (max |X| %.2f)\n", N, INV ? "inverse" : "forward", worst, scale);
    return worst / scale;
}

int main() {
    double w = 0;
    w = fmax(w, check<6, false>());  w = fmax(w, check<6, true>());
    w = fmax(w, check<7, false>());  w = fmax(w, check<7, true>());
    w = fmax(w, check<8, false>());  w = fmax(w, check<8, true>());
    w = fmax(w, check<9, false>());  w = fmax(w, check<9, true>());
    w = fmax(w, check<10, false>()); w = fmax(w, check<10, true>());
    w = fmax(w, check<11, false>()); w = fmax(w, check<11, true>());
    w = fmax(w, check<12, false>()); w = fmax(w, check<12, true>());
    w = fmax(w, check_z<6, true>());  w = fmax(w, check_z<7, true>());  w = fmax(w, check_z<8, false>());
    w = fmax(w, check_z<8, true>());  w = fmax(w, check_z<9, true>());  w = fmax(w, check_z<10, true>());
    w = fmax(w, check_z<11, true>()); w = fmax(w, check_z<12, true>()); w = fmax(w, check_z<12, false>());
    printf("worst relative error %.3e\n", w);
    return w < 2e-6 ? 0 : 1;
}
