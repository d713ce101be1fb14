// K4 / K4L: STFT magnitudes and the fused multi-scale spectral loss (SURVEY 8a rows a11, a12 and
// their backward).
//
// Reference path replaced:
//   ddsp/core.py:27-41     multiscale_fft: per scale torch.stft(center=True reflect pad, periodic hann,
//                          normalized) .abs()  -> reflection_pad1d + cuFFT R2C + abs per scale
//   train.py:70-76,92-103  multiscale_spec_loss: ~40 elementwise/reduction launches over 12.4 M bins
//   autograd backward of both (sgn, irfft, fold of the overlapping frames, reflection_pad backward)
//
// Design.  One CTA owns a tile of FT*hop consecutive positions of the PADDED signal
// of one voice and computes every frame that overlaps the tile (ov = ceil(s/hop)-1 extra frames at
// the left), so the overlap-add of the gradient is an exclusive, ordered gather in shared memory:
// no atomics, bit-reproducible.  Two real sequences ride in one complex FFT:
//   loss     : frame of rec (re) + frame of target (im)          -> one forward FFT per frame
//   gradient : Hermitian-extended gradient spectra of two frames  -> one inverse FFT per frame pair
//   mags     : two frames of the same signal                      -> one forward FFT per frame pair
// Gradient of the samples in the reflect padding goes to a small per-voice edge buffer and is folded
// back by ddsp_b200_stft_fold_edges (again a gather).
#include "fft.cuh"
#include "regfft.cuh"

namespace {

constexpr int kStftThreads = 256;

__device__ __forceinline__ int64_t reflect_index(int64_t m, int64_t N) {
    if (m < 0) m = -m;
    if (m >= N) m = 2 * (N - 1) - m;
    return m;
}

struct TileGeom {
    int s, lg, hop, frames, FT, ov, NF;       // NF = frame slots per batch (even)
    int64_t N;
};

__device__ __forceinline__ float2 untangle_re(float2 zk, float2 zm) {   // (Z[k] + conj(Z[s-k]))/2
    return make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
}
__device__ __forceinline__ float2 untangle_im(float2 zk, float2 zm) {   // (Z[k] - conj(Z[s-k]))/(2i)
    return make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
}

// Ordered gather overlap-add of the frames of one batch into the tile's owned positions.
//   frame slot q (frame fb+q) has its real gradient sequence in buf + (q>>1)*pair_pitch, component
//   (q&1 ? .y : .x).
__device__ __forceinline__ void gather_ola(float *ola, const float2 *buf, int pair_pitch,
                                           const float *__restrict__ window, const TileGeom &g,
                                           int64_t P0, int fb, int fe, int tid) {
    const int owned = g.FT * g.hop;
    for (int t = tid; t < owned; t += kStftThreads) {
        const int64_t i = P0 + t;
        int flo = (int)((i - g.s + g.hop) / g.hop);            // ceil((i-s+1)/hop) for i-s+1 >= 0
        if (i - g.s + 1 <= 0) flo = 0;
        int fhi = (int)(i / g.hop);
        flo = max(flo, fb);
        fhi = min(fhi, min(fb + g.NF, fe) - 1);
        float acc = ola[t];
        for (int f = flo; f <= fhi; ++f) {
            const int q = f - fb;
            const int n = (int)(i - (int64_t)f * g.hop);
            const float2 v = buf[(q >> 1) * pair_pitch + fpad(n)];
            acc = fmaf(__ldg(window + n), (q & 1) ? v.y : v.x, acc);
        }
        ola[t] = acc;
    }
}

// Owned positions -> d_sig (interior) or edge buffer (reflect padding).
__device__ __forceinline__ void store_owned(const float *ola, float *__restrict__ d_sig,
                                            float *__restrict__ edge, const TileGeom &g, int64_t P0,
                                            int accumulate, int tid) {
    const int owned = g.FT * g.hop;
    const int hs = g.s >> 1;
    for (int t = tid; t < owned; t += kStftThreads) {
        const int64_t i = P0 + t;
        if (i >= g.N + g.s) break;
        const float v = ola[t];
        if (i < hs) edge[i] = v;
        else if (i < g.N + hs) {
            float *p = d_sig + (i - hs);
            *p = accumulate ? *p + v : v;
        } else edge[hs + (i - g.N - hs)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused loss (+ gradient w.r.t. rec) for one scale.
// ---------------------------------------------------------------------------------------------
template <bool GRAD>
__global__ void __launch_bounds__(kStftThreads)
mss_scale_kernel(const float *__restrict__ target, const float *__restrict__ rec,
                 const float *__restrict__ window, const float2 *__restrict__ tw, int tws,
                 float *__restrict__ partial, float *__restrict__ d_rec, float *__restrict__ edge,
                 TileGeom g, int accumulate, float inv_cnt) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = fpad_size(g.s);
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + (size_t)g.NF * pitch;
    float *ola = reinterpret_cast<float *>(bufB + (size_t)g.NF * pitch);
    __shared__ float red[2][kStftThreads / 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * g.FT;
    const int64_t P0 = (int64_t)f0 * g.hop;
    const int fs = GRAD ? max(0, f0 - g.ov) : min(f0, g.frames);
    const int fe = min(f0 + g.FT, g.frames);
    const float *xr = rec + (size_t)b * g.N;
    const float *xt = target + (size_t)b * g.N;
    const float rs = rsqrtf((float)g.s);
    const int hs = g.s >> 1;

    if (GRAD)
        for (int t = tid; t < g.FT * g.hop; t += kStftThreads) ola[t] = 0.f;
    float lin = 0.f, lg_ = 0.f;

    for (int fb = fs; fb < fe; fb += g.NF) {
        __syncthreads();
        // 1. windowed frames: re = rec, im = target
        for (int idx = tid; idx < g.NF * g.s; idx += kStftThreads) {
            const int q = idx >> g.lg, n = idx & (g.s - 1);
            const int f = fb + q;
            float2 v = make_float2(0.f, 0.f);
            if (f < fe) {
                const int64_t m = reflect_index((int64_t)f * g.hop + n - hs, g.N);
                const float w = __ldg(window + n);
                v = make_float2(__ldg(xr + m) * w, __ldg(xt + m) * w);
            }
            bufA[q * pitch + fpad(n)] = v;
        }
        __syncthreads();
        float2 *Z = cta_fft<false>(bufA, bufB, pitch, g.NF, g.s, g.lg, tw, tws, tid, kStftThreads);
        float2 *other = (Z == bufA) ? bufB : bufA;
        // 2. per bin: magnitudes, loss terms, gradient spectrum; pairs of frames share a buffer
        const int bins = hs + 1;
        for (int idx = tid; idx < (g.NF >> 1) * bins; idx += kStftThreads) {
            const int pr = idx / bins, k = idx - pr * bins;
            const int km = (g.s - k) & (g.s - 1);
            float2 u[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = 2 * pr + e, f = fb + q;
                u[e] = make_float2(0.f, 0.f);
                if (f < fe) {
                    const float2 zk = Z[q * pitch + fpad(k)], zm = Z[q * pitch + fpad(km)];
                    const float2 Y = untangle_re(zk, zm);     // rec spectrum
                    const float2 X = untangle_im(zk, zm);     // target spectrum
                    const float ay = sqrtf(fmaf(Y.x, Y.x, Y.y * Y.y));
                    const float ax = sqrtf(fmaf(X.x, X.x, X.y * X.y));
                    const float sy = ay * rs, sx = ax * rs;
                    const float dl = logf(sy + 1e-7f) - logf(sx + 1e-7f);
                    if (f >= f0) {                              // loss counted by the owning tile only
                        lin += fabsf(sx - sy);
                        lg_ += fabsf(dl);
                    }
                    if (GRAD && ay > 0.f) {
                        const float d = sy - sx;
                        const float sg = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
                        const float sl = (dl > 0.f) ? 1.f : (dl < 0.f ? -1.f : 0.f);
                        // dL/dSy * (1/sqrt(s)) / |Y|
                        const float c = (sg + sl / (sy + 1e-7f)) * inv_cnt * rs / ay;
                        u[e] = make_float2(c * Y.x, c * Y.y);
                    }
                }
            }
            if (GRAD) {
                float2 *dst = Z + (2 * pr) * pitch;
                if (k == 0 || k == hs) {
                    dst[fpad(k)] = make_float2(u[0].x, u[1].x);
                } else {
                    // Zi[k] = (Ua + i Ub)/2 ; Zi[s-k] = (conj Ua + i conj Ub)/2
                    dst[fpad(k)] = make_float2(0.5f * (u[0].x - u[1].y), 0.5f * (u[0].y + u[1].x));
                    dst[fpad(km)] = make_float2(0.5f * (u[0].x + u[1].y), 0.5f * (-u[0].y + u[1].x));
                }
            }
        }
        if (GRAD) {
            __syncthreads();
            float2 *R = cta_fft<true>(Z, other, 2 * pitch, g.NF >> 1, g.s, g.lg, tw, tws, tid,
                                      kStftThreads);
            gather_ola(ola, R, 2 * pitch, window, g, P0, fb, fe, tid);
        }
    }
    __syncthreads();
    if (GRAD) store_owned(ola, d_rec + (size_t)b * g.N, edge + (size_t)b * g.s, g, P0, accumulate, tid);

    lin = ddsp_warp_sum(lin);
    lg_ = ddsp_warp_sum(lg_);
    if ((tid & 31) == 0) { red[0][tid >> 5] = lin; red[1][tid >> 5] = lg_; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, c = 0.f;
        for (int i = 0; i < kStftThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
        float *p = partial + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
        p[0] = a;
        p[1] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused loss, register-tiled FFT version (n_fft = 64 .. 4096; regfft.cuh).  Same tiling and maths as
// mss_scale_kernel above; what changes is the transform: T = n_fft/16 threads per frame, 16 values
// per thread in registers, 2 or 3 shared-memory round trips per transform instead of log4(n_fft).
// A batch is NF = 2G frame slots (G = transforms the CTA runs at once): two forward rounds, then ONE
// round of G inverse transforms, each carrying the gradient spectra of slots p (re) and p+G (im).
// ---------------------------------------------------------------------------------------------
template <int LG>
struct RegCfg {
    using P = regfft::Plan<LG>;
    static constexpr int G = (kStftThreads / P::T) > 32 ? 32 : (kStftThreads / P::T);
    static constexpr int NF = 2 * G;
};

// one full transform pass over the CTA's two slot groups (forward) or one (inverse) is written out
// phase by phase in the kernel; this helper only hides the 2- vs 3-stage difference.
template <int LG, bool INV>
__device__ __forceinline__ void reg_stages_after0(float2 (&xa)[16], float2 (&xb)[16], float2 *bufa,
                                                  float2 *bufb, bool act_a, bool act_b, int t,
                                                  const float2 *__restrict__ tw) {
    using namespace regfft;
    __syncthreads();
    if (act_a) stage_load<LG, 1>(xa, bufa, t);
    if (act_b) stage_load<LG, 1>(xb, bufb, t);
    __syncthreads();
    if (act_a) stage_compute_store<LG, 1, INV>(xa, bufa, t, tw);
    if (act_b) stage_compute_store<LG, 1, INV>(xb, bufb, t, tw);
    if (Plan<LG>::STAGES == 3) {
        __syncthreads();
        if (act_a) stage_load<LG, 2>(xa, bufa, t);
        if (act_b) stage_load<LG, 2>(xb, bufb, t);
        __syncthreads();
        if (act_a) stage_compute_store<LG, 2, INV>(xa, bufa, t, tw);
        if (act_b) stage_compute_store<LG, 2, INV>(xb, bufb, t, tw);
    }
    __syncthreads();
}

// MODE: what the per-bin phase does with the spectra
//   kLoss / kLossGrad : rec (re) + target (im) per frame; loss terms (+ gradient spectra)
//   kMagFwd           : one signal; write |STFT| to mag_io[b][k][frame]               (multiscale_fft)
//   kMagBwd           : one signal; gradient spectra from d_mag = mag_io[b][k][frame]  (its backward)
enum { kLoss = 0, kLossGrad = 1, kMagFwd = 2, kMagBwd = 3 };

template <int LG, int MODE>
__global__ void __launch_bounds__(kStftThreads, 2)
mss_scale_reg_kernel(const float *__restrict__ target, const float *__restrict__ rec,
                     const float *__restrict__ window, const float2 *__restrict__ tw,
                     float *__restrict__ partial, float *__restrict__ d_rec, float *__restrict__ edge,
                     float *__restrict__ mag_io, TileGeom g, int accumulate, float inv_cnt) {
    using namespace regfft;
    using P = Plan<LG>;
    using C = RegCfg<LG>;
    constexpr bool GRAD = MODE == kLossGrad || MODE == kMagBwd;
    constexpr bool MAG = MODE == kMagFwd || MODE == kMagBwd;
    constexpr int N = P::N, T = P::T, G = C::G, NF = C::NF, PITCH = P::PITCH;
    constexpr int HS = N / 2, BINS = HS + 1;
    extern __shared__ __align__(16) float smem[];
    float2 *buf = reinterpret_cast<float2 *>(smem);                       // [NF][PITCH]
    float *ola = reinterpret_cast<float *>(buf + (size_t)NF * PITCH);     // [FT*hop]
    __shared__ float red[2][kStftThreads / 32];

    const int tid = threadIdx.x;
    const int grp = tid / T, t = tid - grp * T;
    const bool act = grp < G;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * g.FT;
    const int64_t P0 = (int64_t)f0 * g.hop;
    const int owned = g.FT * g.hop;
    const int fs = GRAD ? max(0, f0 - g.ov) : min(f0, g.frames);
    const int fe = min(f0 + g.FT, g.frames);
    const float *xr = rec + (size_t)b * g.N;
    const float *xt = MAG ? xr : target + (size_t)b * g.N;
    const float rs = rsqrtf((float)N);
    const int Ni = (int)g.N;
    float *mg = MAG ? mag_io + (size_t)b * BINS * g.frames : nullptr;

    if (GRAD)
        for (int i = tid; i < owned; i += kStftThreads) ola[i] = 0.f;
    float lin = 0.f, lgs = 0.f;

    for (int fb = fs; fb < fe; fb += NF) {
        __syncthreads();                       // previous batch fully consumed
        float2 xa[16], xb[16];
        float2 *bufa = buf + (size_t)grp * PITCH, *bufb = buf + (size_t)(grp + G) * PITCH;
        // ---- forward: slot grp (round 0) and slot grp+G (round 1); stage 0 straight from global
        if (act) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int f = fb + grp + e * G;
                float2 (&x)[16] = e ? xb : xa;
                if (f < fe) {
                    const int base = f * g.hop - HS + t;
                    if (f * g.hop - HS >= 0 && f * g.hop - HS + N <= Ni) {      // interior frame: no reflection
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const float w = __ldg(window + t + r * T);
                            x[r] = make_float2(__ldg(xr + base + r * T) * w, MAG ? 0.f : __ldg(xt + base + r * T) * w);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            int m = base + r * T;
                            m = m < 0 ? -m : m;
                            m = m >= Ni ? 2 * (Ni - 1) - m : m;
                            const float w = __ldg(window + t + r * T);
                            x[r] = make_float2(__ldg(xr + m) * w, MAG ? 0.f : __ldg(xt + m) * w);
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) x[r] = make_float2(0.f, 0.f);
                }
                stage_compute_store<LG, 0, false>(x, e ? bufb : bufa, t, tw);
            }
        }
        reg_stages_after0<LG, false>(xa, xb, bufa, bufb, act, act, t, tw);

        if (MODE == kMagFwd) {
            // ---- |STFT| out; lanes take consecutive frame slots so the (B, bins, frames) store coalesces
            for (int idx = tid; idx < BINS * NF; idx += kStftThreads) {
                const int q = idx & (NF - 1), k = idx / NF;
                const int f = fb + q;
                if (f < fe) {
                    const float2 *Z = buf + (size_t)q * PITCH;
                    const float2 Y = untangle_re(Z[pad16(k)], Z[pad16((N - k) & (N - 1))]);
                    mg[(size_t)k * g.frames + f] = sqrtf(fmaf(Y.x, Y.x, Y.y * Y.y)) * rs;
                }
            }
            continue;
        }
        if (MODE == kMagBwd) {
            // ---- pass 1 (coalesced read of d_mag): U[q][k] = d_mag * X/|X| / sqrt(N), in place at index k
            for (int idx = tid; idx < BINS * NF; idx += kStftThreads) {
                const int q = idx & (NF - 1), k = idx / NF;
                const int f = fb + q;
                float2 *Z = buf + (size_t)q * PITCH;
                const float2 Y = untangle_re(Z[pad16(k)], Z[pad16((N - k) & (N - 1))]);
                const float yy = fmaf(Y.x, Y.x, Y.y * Y.y);
                float c = 0.f;
                if (f < fe && yy > 0.f) c = __ldg(mg + (size_t)k * g.frames + f) * rs * rsqrtf(yy);
                Z[pad16(k)] = make_float2(c * Y.x, c * Y.y);
            }
            __syncthreads();
            // ---- pass 2: Hermitian packing of the pair (slot p, slot p+G)
            for (int idx = tid; idx < G * BINS; idx += kStftThreads) {
                const int p = idx / BINS, k = idx - p * BINS;
                const int km = (N - k) & (N - 1);
                const float2 ua = buf[(size_t)p * PITCH + pad16(k)], ub = buf[(size_t)(p + G) * PITCH + pad16(k)];
                float2 *dst = buf + (size_t)p * PITCH;
                const bool edge_bin = (k == 0) | (k == HS);
                dst[pad16(k)] = edge_bin ? make_float2(ua.x, ub.x)
                                         : make_float2(0.5f * (ua.x - ub.y), 0.5f * (ua.y + ub.x));
                if (!edge_bin) dst[pad16(km)] = make_float2(0.5f * (ua.x + ub.y), 0.5f * (-ua.y + ub.x));
            }
        }
        if (!MAG)
        // ---- per bin: magnitudes, loss, gradient spectra of the pair (slot p, slot p+G)
        for (int idx = tid; idx < G * BINS; idx += kStftThreads) {
            const int p = idx / BINS, k = idx - p * BINS;
            const int km = (N - k) & (N - 1);
            float2 u[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int f = fb + p + e * G;
                const float2 *Z = buf + (size_t)(p + e * G) * PITCH;
                const float2 zk = Z[pad16(k)], zm = Z[pad16(km)];
                const float2 Y = untangle_re(zk, zm);         // rec spectrum
                const float2 X = untangle_im(zk, zm);         // target spectrum
                const float yy = fmaf(Y.x, Y.x, Y.y * Y.y), xx = fmaf(X.x, X.x, X.y * X.y);
                const float ry = rsqrtf(fmaxf(yy, 1e-37f));   // 1/|Y| (|Y| = 0 -> gradient 0 below)
                const float sy = sqrtf(yy) * rs, sx = sqrtf(xx) * rs;
                const float d = sy - sx;
                const float iy = __fdividef(1.0f, sy + 1e-7f);
                // loss is counted by the owning tile only, and only for frames that exist
                const float own = (f >= f0 && f < fe) ? 1.f : 0.f;
                lin = fmaf(own, fabsf(d), lin);
                lgs = fmaf(own, fabsf(__logf((sx + 1e-7f) * iy)), lgs);
                // log is monotonic: sign(log(sy+eps) - log(sx+eps)) == sign(sy - sx)
                const float sg = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
                float c = fmaf(sg, iy, sg) * (inv_cnt * rs) * ry;
                c = (f < fe && yy > 0.f) ? c : 0.f;
                u[e] = make_float2(c * Y.x, c * Y.y);
            }
            if (GRAD) {
                float2 *dst = buf + (size_t)p * PITCH;
                // Zi[k] = (Ua + i Ub)/2 ; Zi[N-k] = (conj Ua + i conj Ub)/2 ; real-only at k = 0, N/2
                const bool edge_bin = (k == 0) | (k == HS);
                const float2 a = edge_bin ? make_float2(u[0].x, u[1].x)
                                          : make_float2(0.5f * (u[0].x - u[1].y), 0.5f * (u[0].y + u[1].x));
                dst[pad16(k)] = a;
                if (!edge_bin)
                    dst[pad16(km)] = make_float2(0.5f * (u[0].x + u[1].y), 0.5f * (-u[0].y + u[1].x));
            }
        }
        if (GRAD) {
            // ---- inverse: one transform per pair, in place in slot p
            __syncthreads();
            if (act) {
#pragma unroll
                for (int r = 0; r < 16; ++r) xa[r] = bufa[pad16(t + r * T)];
            }
            __syncthreads();
            if (act) stage_compute_store<LG, 0, true>(xa, bufa, t, tw);
            reg_stages_after0<LG, true>(xa, xb, bufa, bufb, act, false, t, tw);
            // ---- ordered overlap-add.  Frames f and f + (ov+1) never overlap, so the batch is added in
            // ov+1 phases (one residue class of frames per phase, all its frames in parallel); every owned
            // position receives its <= ov+1 contributions in class order: deterministic, no atomics.
            const int ncls = g.ov + 1;
            const bool vec4 = (g.hop & 3) == 0;            // then every frame start is 16-byte aligned in ola
            for (int c = 0; c < ncls; ++c) {
                // frames fb + c, fb + c + ncls, ... of this batch
                const int nfr = (NF - c + ncls - 1) / ncls;
                if (vec4) {
                    for (int idx = tid; idx < nfr * (N / 4); idx += kStftThreads) {
                        const int q = c + (idx >> (LG - 2)) * ncls, n = (idx & (N / 4 - 1)) * 4;
                        const int f = fb + q;
                        const int i = f * g.hop + n - (int)P0;
                        if (f < fe && i >= 0 && i < owned) {       // owned % 4 == 0: all four or none
                            const float2 *src = buf + (size_t)(q >= G ? q - G : q) * PITCH + pad16(n);
                            // (the padded index may be odd: 8-byte loads only)
                            const float2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
                            const float4 w = __ldg(reinterpret_cast<const float4 *>(window + n));
                            float4 o = *reinterpret_cast<float4 *>(ola + i);
                            const bool im = q >= G;
                            o.x = fmaf(w.x, im ? v0.y : v0.x, o.x);
                            o.y = fmaf(w.y, im ? v1.y : v1.x, o.y);
                            o.z = fmaf(w.z, im ? v2.y : v2.x, o.z);
                            o.w = fmaf(w.w, im ? v3.y : v3.x, o.w);
                            *reinterpret_cast<float4 *>(ola + i) = o;
                        }
                    }
                } else {
                    for (int idx = tid; idx < nfr * N; idx += kStftThreads) {
                        const int q = c + (idx >> LG) * ncls, n = idx & (N - 1);
                        const int f = fb + q;
                        const int i = f * g.hop + n - (int)P0;
                        if (f < fe && i >= 0 && i < owned) {
                            const float2 v = buf[(size_t)(q >= G ? q - G : q) * PITCH + pad16(n)];
                            ola[i] = fmaf(__ldg(window + n), q >= G ? v.y : v.x, ola[i]);
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (GRAD) store_owned(ola, d_rec + (size_t)b * g.N, edge + (size_t)b * g.s, g, P0, accumulate, tid);

    lin = ddsp_warp_sum(lin);
    lgs = ddsp_warp_sum(lgs);
    if ((tid & 31) == 0) { red[0][tid >> 5] = lin; red[1][tid >> 5] = lgs; }
    __syncthreads();
    if (tid == 0 && !MAG) {
        float a = 0.f, c = 0.f;
        for (int i = 0; i < kStftThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
        float *pp = partial + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
        pp[0] = a;
        pp[1] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// |STFT| forward (API path of multiscale_fft): two frames of the signal per complex FFT.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStftThreads)
stft_mag_fwd_kernel(const float *__restrict__ signal, const float *__restrict__ window,
                    const float2 *__restrict__ tw, int tws, float *__restrict__ mag, TileGeom g) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = fpad_size(g.s);
    const int NP = g.NF >> 1;                           // FFTs per batch
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + (size_t)NP * pitch;
    const int tid = threadIdx.x, b = blockIdx.y;
    const int f0 = blockIdx.x * g.FT;
    const int fe = min(f0 + g.FT, g.frames);
    const float *x = signal + (size_t)b * g.N;
    const float rs = rsqrtf((float)g.s);
    const int hs = g.s >> 1, bins = hs + 1;
    float *mg = mag + (size_t)b * bins * g.frames;

    for (int fb = f0; fb < fe; fb += g.NF) {
        __syncthreads();
        for (int idx = tid; idx < NP * g.s; idx += kStftThreads) {
            const int p = idx >> g.lg, n = idx & (g.s - 1);
            const int fa = fb + 2 * p;
            float2 v = make_float2(0.f, 0.f);
            const float w = __ldg(window + n);
            if (fa < fe) v.x = __ldg(x + reflect_index((int64_t)fa * g.hop + n - hs, g.N)) * w;
            if (fa + 1 < fe) v.y = __ldg(x + reflect_index((int64_t)(fa + 1) * g.hop + n - hs, g.N)) * w;
            bufA[p * pitch + fpad(n)] = v;
        }
        __syncthreads();
        const float2 *Z = cta_fft<false>(bufA, bufB, pitch, NP, g.s, g.lg, tw, tws, tid, kStftThreads);
        for (int idx = tid; idx < NP * bins; idx += kStftThreads) {
            const int p = idx / bins, k = idx - p * bins;
            const int fa = fb + 2 * p;
            if (fa >= fe) continue;
            const float2 zk = Z[p * pitch + fpad(k)], zm = Z[p * pitch + fpad((g.s - k) & (g.s - 1))];
            const float2 A = untangle_re(zk, zm), C = untangle_im(zk, zm);
            mg[(size_t)k * g.frames + fa] = sqrtf(fmaf(A.x, A.x, A.y * A.y)) * rs;
            if (fa + 1 < fe) mg[(size_t)k * g.frames + fa + 1] = sqrtf(fmaf(C.x, C.x, C.y * C.y)) * rs;
        }
    }
}

// |STFT| backward: d_signal from d_mag (recomputes the spectra).
__global__ void __launch_bounds__(kStftThreads)
stft_mag_bwd_kernel(const float *__restrict__ signal, const float *__restrict__ d_mag,
                    const float *__restrict__ window, const float2 *__restrict__ tw, int tws,
                    float *__restrict__ d_sig, float *__restrict__ edge, TileGeom g, int accumulate) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = fpad_size(g.s);
    const int NP = g.NF >> 1;
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + (size_t)NP * pitch;
    float *ola = reinterpret_cast<float *>(bufB + (size_t)NP * pitch);
    const int tid = threadIdx.x, b = blockIdx.y;
    const int f0 = blockIdx.x * g.FT;
    const int64_t P0 = (int64_t)f0 * g.hop;
    const int fs = max(0, f0 - g.ov);
    const int fe = min(f0 + g.FT, g.frames);
    const float *x = signal + (size_t)b * g.N;
    const float rs = rsqrtf((float)g.s);
    const int hs = g.s >> 1, bins = hs + 1;
    const float *gm = d_mag + (size_t)b * bins * g.frames;

    for (int t = tid; t < g.FT * g.hop; t += kStftThreads) ola[t] = 0.f;
    for (int fb = fs; fb < fe; fb += g.NF) {
        __syncthreads();
        for (int idx = tid; idx < NP * g.s; idx += kStftThreads) {
            const int p = idx >> g.lg, n = idx & (g.s - 1);
            const int fa = fb + 2 * p;
            float2 v = make_float2(0.f, 0.f);
            const float w = __ldg(window + n);
            if (fa < fe) v.x = __ldg(x + reflect_index((int64_t)fa * g.hop + n - hs, g.N)) * w;
            if (fa + 1 < fe) v.y = __ldg(x + reflect_index((int64_t)(fa + 1) * g.hop + n - hs, g.N)) * w;
            bufA[p * pitch + fpad(n)] = v;
        }
        __syncthreads();
        float2 *Z = cta_fft<false>(bufA, bufB, pitch, NP, g.s, g.lg, tw, tws, tid, kStftThreads);
        float2 *other = (Z == bufA) ? bufB : bufA;
        for (int idx = tid; idx < NP * bins; idx += kStftThreads) {
            const int p = idx / bins, k = idx - p * bins;
            const int km = (g.s - k) & (g.s - 1);
            const int fa = fb + 2 * p;
            float2 *zb = Z + p * pitch;
            const float2 zk = zb[fpad(k)], zm = zb[fpad(km)];
            const float2 A = untangle_re(zk, zm), C = untangle_im(zk, zm);
            float2 ua = make_float2(0.f, 0.f), ub = ua;
            if (fa < fe) {
                const float a = sqrtf(fmaf(A.x, A.x, A.y * A.y));
                if (a > 0.f) {
                    const float c = __ldg(gm + (size_t)k * g.frames + fa) * rs / a;
                    ua = make_float2(c * A.x, c * A.y);
                }
            }
            if (fa + 1 < fe) {
                const float a = sqrtf(fmaf(C.x, C.x, C.y * C.y));
                if (a > 0.f) {
                    const float c = __ldg(gm + (size_t)k * g.frames + fa + 1) * rs / a;
                    ub = make_float2(c * C.x, c * C.y);
                }
            }
            if (k == 0 || k == hs) {
                zb[fpad(k)] = make_float2(ua.x, ub.x);
            } else {
                zb[fpad(k)] = make_float2(0.5f * (ua.x - ub.y), 0.5f * (ua.y + ub.x));
                zb[fpad(km)] = make_float2(0.5f * (ua.x + ub.y), 0.5f * (-ua.y + ub.x));
            }
        }
        __syncthreads();
        float2 *R = cta_fft<true>(Z, other, pitch, NP, g.s, g.lg, tw, tws, tid, kStftThreads);
        gather_ola(ola, R, pitch, window, g, P0, fb, fe, tid);
    }
    __syncthreads();
    store_owned(ola, d_sig + (size_t)b * g.N, edge + (size_t)b * g.s, g, P0, accumulate, tid);
}

// d_sig[b,m] += edge contributions of every scale, gathered per target sample (deterministic).
struct FoldArgs {
    int n_scales;
    int s[8];
    int64_t off[8];          // float offset of scale i's edge block: layout [scale][B][s]
};

__global__ void stft_fold_kernel(const float *__restrict__ edge, float *__restrict__ d_sig, int B,
                                 int64_t N, FoldArgs fa, int hs_max) {
    // only samples 1..hs_max and N-1-hs_max..N-2 mirror into the padding; when the two ranges would
    // overlap (short signals) every sample is visited instead
    const int b = blockIdx.y;
    const bool full = 2 * (int64_t)hs_max + 2 >= N;
    const int64_t count = full ? N : 2 * (int64_t)hs_max;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = full ? i : (i < hs_max ? 1 + i : N - 1 - hs_max + (i - hs_max));
        float acc = 0.f;
        for (int k = 0; k < fa.n_scales; ++k) {
            const int hs = fa.s[k] >> 1;
            const float *e = edge + fa.off[k] + (size_t)b * fa.s[k];
            if (m >= 1 && m <= hs) acc += e[hs - m];                         // left pad: i = hs - m
            if (m <= N - 2 && m >= N - 1 - hs) acc += e[hs + (N - 2 - m)];   // right pad
        }
        if (acc != 0.f) d_sig[(size_t)b * N + m] += acc;
    }
}

// d_rec[b,m] = sum over scales of the per-scale gradients (fixed order) + the folded edges.  Lets the
// per-scale launches run concurrently (own output buffers) and still be deterministic.
__global__ void mss_combine_kernel(const float *__restrict__ per_scale, const float *__restrict__ edge,
                                   float *__restrict__ d_sig, int B, int64_t N, FoldArgs fa, int hs_max) {
    const int b = blockIdx.y;
    const size_t plane = (size_t)B * N;
    const bool vec = (N & 3) == 0;
    const int64_t count = vec ? N / 4 : N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m0 = vec ? 4 * i : i;
        const int width = vec ? 4 : 1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            for (int k = 0; k < fa.n_scales; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(per_scale + k * plane + (size_t)b * N + m0));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            }
        } else {
            for (int k = 0; k < fa.n_scales; ++k) acc[0] += per_scale[k * plane + (size_t)b * N + m0];
        }
        // only samples 1..hs_max and N-1-hs_max..N-2 mirror into the reflect padding
        if (m0 <= hs_max || m0 + width >= N - 1 - hs_max) {
            for (int e = 0; e < width; ++e) {
                const int64_t m = m0 + e;
                for (int k = 0; k < fa.n_scales; ++k) {
                    const int hs = fa.s[k] >> 1;
                    const float *ed = edge + fa.off[k] + (size_t)b * fa.s[k];
                    if (m >= 1 && m <= hs) acc[e] += ed[hs - m];
                    if (m <= N - 2 && m >= N - 1 - hs) acc[e] += ed[hs + (N - 2 - m)];
                }
            }
        }
        if (vec) *reinterpret_cast<float4 *>(d_sig + (size_t)b * N + m0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else d_sig[(size_t)b * N + m0] = acc[0];
    }
}

struct FinArgs {
    int n_scales;
    int64_t off[8];          // pair offset of scale i's partials
    int64_t cnt[8];          // pairs for scale i
    float inv[8];            // 1 / (B * bins * frames)
};

__global__ void __launch_bounds__(1024)
mss_finalize_kernel(const float *__restrict__ partial, float *__restrict__ loss, FinArgs fa) {
    // one pass: every thread adds its share of every scale's partials (already weighted by the scale's
    // 1/count), then a single block reduction in double; fixed order -> deterministic
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = 0; i < fa.n_scales; ++i) {
        const float *p = partial + 2 * fa.off[i];
        double a = 0.0;
        for (int64_t j = threadIdx.x; j < fa.cnt[i]; j += blockDim.x) a += (double)p[2 * j] + (double)p[2 * j + 1];
        acc += a * (double)fa.inv[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) loss[0] = (float)v;
    }
}

// ---- host-side geometry ---------------------------------------------------------------------
struct HostGeom {
    TileGeom g;
    int tiles;
    size_t smem_loss, smem_mag_fwd, smem_mag_bwd;
};

bool reg_path(int n_fft) { return n_fft >= 64 && n_fft <= 4096; }

// regfft.cuh stage table: stage 1 [r-1][k] (k < 16), then stage 2 [r-1][k] (k < 256)
__global__ void stage_twiddle_kernel(float2 *__restrict__ tab, int n_fft, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r1 = n_fft <= 256 ? n_fft / 16 : 16, r2 = n_fft <= 256 ? 1 : n_fft / 256;
    const int n1 = (r1 - 1) * 16;
    int r, k, denom;
    if (i < n1) { r = i / 16 + 1; k = i % 16; denom = 16 * r1; }
    else { r = (i - n1) / 256 + 1; k = (i - n1) % 256; denom = 256 * r2; }
    double sn, cs;
    sincospi(2.0 * (double)((long long)r * k % denom) / (double)denom, &sn, &cs);
    tab[i] = make_float2((float)cs, (float)(-sn));
}

// geometry of the register-FFT loss kernel: NF = 2G frame slots per batch, tiles of TF = FT + ov frames
int make_geom_reg(int B, int64_t N, int n_fft, int hop, HostGeom *out) {
    if (n_fft & (n_fft - 1)) return DDSP_B200_EUNSUPPORTED;
    if (hop < 1 || hop > n_fft) return DDSP_B200_EUNSUPPORTED;
    if (N <= n_fft / 2 || N >= (1ll << 30)) return DDSP_B200_EUNSUPPORTED;
    TileGeom g;
    g.s = n_fft; g.lg = ddsp_ilog2(n_fft); g.hop = hop; g.N = N;
    g.frames = 1 + (int)(N / hop);
    g.ov = (n_fft + hop - 1) / hop - 1;
    const int T = n_fft / 16;
    int G = kStftThreads / T;
    if (G > 32) G = 32;
    const int nf = 2 * G;
    g.NF = nf;
    const size_t pitch = ((size_t)n_fft + (n_fft >> 4) + 1) & ~(size_t)1;
    auto total = [&](int tf_) {
        return (size_t)nf * pitch * sizeof(float2) + (size_t)(tf_ - g.ov) * hop * sizeof(float);
    };
    // frames per tile (a whole number of batches): as many as keep two CTAs per SM (<= 113 KB each),
    // up to 64 (228 KB per SM, 1 KB reserved per CTA, a few static bytes) -- the ov frames a tile recomputes are then a small fraction
    int tf = nf;
    while (tf - g.ov < 1) tf += nf;
    while (tf + nf <= 64 && total(tf + nf) <= 110 * 1024) tf += nf;
    // small batches (strong scaling leaves 8 voices per GPU): prefer more, smaller tiles until there are
    // about two CTAs per SM, at the price of recomputing the ov overlap frames more often
    auto tiles_of = [&](int tf_) { return ddsp_ceil_div(N + n_fft, (int64_t)(tf_ - g.ov) * hop); };
    while (tiles_of(tf) * B < 2 * DDSP_SM_COUNT && tf - nf - g.ov >= 1 && 2 * (tf - nf - g.ov) >= g.ov) tf -= nf;
    if (total(tf) > 220 * 1024) return DDSP_B200_EUNSUPPORTED;
    g.FT = tf - g.ov;
    out->g = g;
    out->smem_loss = total(tf);
    out->smem_mag_fwd = out->smem_mag_bwd = 0;
    out->tiles = (int)ddsp_ceil_div(N + n_fft, (int64_t)g.FT * hop);
    return DDSP_B200_OK;
}

int make_geom(int64_t N, int n_fft, int hop, bool with_overlap, HostGeom *out) {
    if (n_fft < 8 || (n_fft & (n_fft - 1)) || n_fft > 8192) return DDSP_B200_EUNSUPPORTED;
    if (hop < 1 || hop > n_fft) return DDSP_B200_EUNSUPPORTED;
    if (N <= n_fft / 2) return DDSP_B200_EUNSUPPORTED;            // reflect padding needs pad < N
    TileGeom g;
    g.s = n_fft; g.lg = ddsp_ilog2(n_fft); g.hop = hop; g.N = N;
    g.frames = 1 + (int)(N / hop);
    g.ov = (n_fft + hop - 1) / hop - 1;
    int nf = 2048 / n_fft;                                         // frame slots per batch
    if (nf < 2) nf = 2;
    if (nf > 16) nf = 16;
    g.NF = nf;
    // tile = a whole number of batches of frames, minus the overlap frames it recomputes
    int ft = nf * ((16 + nf - 1) / nf);
    if (with_overlap) {
        ft -= g.ov;
        while (ft < 1) ft += nf;
    }
    const size_t pitch = fpad_size(n_fft);
    auto total = [&](int ft_) {
        return 2 * (size_t)nf * pitch * sizeof(float2) + (size_t)ft_ * hop * sizeof(float);
    };
    while (total(ft) > 200 * 1024 && ft > nf) ft -= nf;
    if (total(ft) > 220 * 1024) return DDSP_B200_EUNSUPPORTED;
    g.FT = ft;
    out->g = g;
    out->smem_loss = total(ft);
    out->smem_mag_fwd = 2 * (size_t)(nf / 2) * pitch * sizeof(float2);
    out->smem_mag_bwd = out->smem_mag_fwd + (size_t)ft * hop * sizeof(float);
    // gradient tiles must cover every padded position; magnitude tiles only the frames
    const int64_t span = (int64_t)ft * hop;
    out->tiles = with_overlap ? (int)ddsp_ceil_div(N + n_fft, span) : (int)ddsp_ceil_div(g.frames, ft);
    return DDSP_B200_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int loss_geom(int B, int64_t N, int n_fft, int hop, HostGeom *hg) {
    return reg_path(n_fft) ? make_geom_reg(B, N, n_fft, hop, hg) : make_geom(N, n_fft, hop, true, hg);
}

template <int LG, int MODE>
int launch_reg_mode(const float *target, const float *rec, const float *window, const float2 *tw,
                    float *partial, float *d_rec, float *edge, float *mag_io, const HostGeom &hg, int B,
                    int accumulate, float inv_cnt, cudaStream_t st) {
    int s;
    if ((s = set_smem(mss_scale_reg_kernel<LG, MODE>, hg.smem_loss))) return s;
    mss_scale_reg_kernel<LG, MODE><<<dim3(hg.tiles, B), kStftThreads, hg.smem_loss, st>>>(
        target, rec, window, tw, partial, d_rec, edge, mag_io, hg.g, accumulate, inv_cnt);
    return 0;
}

template <int LG>
int launch_reg(int mode, const float *target, const float *rec, const float *window, const float2 *tw,
               float *partial, float *d_rec, float *edge, float *mag_io, const HostGeom &hg, int B,
               int accumulate, float inv_cnt, cudaStream_t st) {
    switch (mode) {
        case kLoss: return launch_reg_mode<LG, kLoss>(target, rec, window, tw, partial, d_rec, edge, mag_io, hg, B, accumulate, inv_cnt, st);
        case kLossGrad: return launch_reg_mode<LG, kLossGrad>(target, rec, window, tw, partial, d_rec, edge, mag_io, hg, B, accumulate, inv_cnt, st);
        case kMagFwd: return launch_reg_mode<LG, kMagFwd>(target, rec, window, tw, partial, d_rec, edge, mag_io, hg, B, accumulate, inv_cnt, st);
        default: return launch_reg_mode<LG, kMagBwd>(target, rec, window, tw, partial, d_rec, edge, mag_io, hg, B, accumulate, inv_cnt, st);
    }
}

int dispatch_reg(int mode, const float *target, const float *rec, const float *window, const float2 *stw,
                 float *partial, float *d_rec, float *edge, float *mag_io, const HostGeom &hg, int B,
                 int accumulate, float inv_cnt, cudaStream_t st) {
    int s = DDSP_B200_EUNSUPPORTED;
#define DDSP_REG_CASE(LG)                                                                            \
    case LG:                                                                                         \
        s = launch_reg<LG>(mode, target, rec, window, stw, partial, d_rec, edge, mag_io, hg, B,      \
                           accumulate, inv_cnt, st);                                                 \
        break;
    switch (hg.g.lg) {
        DDSP_REG_CASE(6) DDSP_REG_CASE(7) DDSP_REG_CASE(8) DDSP_REG_CASE(9) DDSP_REG_CASE(10)
        DDSP_REG_CASE(11) DDSP_REG_CASE(12)
        default: break;
    }
#undef DDSP_REG_CASE
    return s ? s : ddsp_launch_status();
}

}  // namespace

extern "C" int64_t ddsp_b200_fft_stage_twiddles_size(int n_fft) {
    if (!reg_path(n_fft) || (n_fft & (n_fft - 1))) return 0;
    const int r1 = n_fft <= 256 ? n_fft / 16 : 16, r2 = n_fft <= 256 ? 1 : n_fft / 256;
    return (int64_t)(r1 - 1) * 16 + (n_fft > 256 ? (int64_t)(r2 - 1) * 256 : 0);
}

extern "C" int ddsp_b200_fft_stage_twiddles(float *table, int n_fft, void *stream) {
    DDSP_REQUIRE(table && reg_path(n_fft) && (n_fft & (n_fft - 1)) == 0);
    const int total = (int)ddsp_b200_fft_stage_twiddles_size(n_fft);
    stage_twiddle_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float2 *>(table), n_fft, total);
    return ddsp_launch_status();
}

extern "C" int64_t ddsp_b200_mss_tiles(int B, int64_t N, int n_fft, int hop) {
    HostGeom hg;
    if (B < 1 || loss_geom(B, N, n_fft, hop, &hg)) return -1;
    return hg.tiles;
}

extern "C" int ddsp_b200_mss_scale(const float *target, const float *rec, const float *window,
                                   const float *twiddle, int n_tab, const float *stage_twiddle,
                                   float *partial, float *d_rec, float *edge, int B, int64_t N,
                                   int n_fft, int hop, int accumulate, void *stream) {
    DDSP_REQUIRE(target && rec && window && twiddle && partial && B > 0 && B <= 65535);
    DDSP_REQUIRE(!d_rec || edge);
    DDSP_REQUIRE(n_tab >= n_fft && n_tab % n_fft == 0);
    HostGeom hg;
    int s = loss_geom(B, N, n_fft, hop, &hg);
    if (s) return s;
    const float inv_cnt = 1.0f / ((float)B * (float)(n_fft / 2 + 1) * (float)hg.g.frames);
    dim3 grid(hg.tiles, B);
    const float2 *tw = reinterpret_cast<const float2 *>(twiddle);
    cudaStream_t st = (cudaStream_t)stream;
    if (reg_path(n_fft)) {
        if (!stage_twiddle) return DDSP_B200_EINVAL;
        return dispatch_reg(d_rec ? kLossGrad : kLoss, target, rec, window,
                            reinterpret_cast<const float2 *>(stage_twiddle), partial, d_rec, edge, nullptr, hg, B,
                            accumulate, inv_cnt, st);
    }
    if (d_rec) {
        if ((s = set_smem(mss_scale_kernel<true>, hg.smem_loss))) return s;
        mss_scale_kernel<true><<<grid, kStftThreads, hg.smem_loss, st>>>(
            target, rec, window, tw, n_tab / n_fft, partial, d_rec, edge, hg.g, accumulate, inv_cnt);
    } else {
        if ((s = set_smem(mss_scale_kernel<false>, hg.smem_loss))) return s;
        mss_scale_kernel<false><<<grid, kStftThreads, hg.smem_loss, st>>>(
            target, rec, window, tw, n_tab / n_fft, partial, nullptr, nullptr, hg.g, accumulate, inv_cnt);
    }
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_mss_finish(const float *partial, const float *edge, const float *d_rec_scales,
                                    float *d_rec, float *loss, int B, int64_t N, const int *scales,
                                    const int *hops, int n_scales, void *stream) {
    DDSP_REQUIRE(partial && loss && scales && hops && n_scales > 0 && n_scales <= 8 && B > 0);
    DDSP_REQUIRE(!d_rec || edge);
    FinArgs fin;
    FoldArgs fold;
    fin.n_scales = fold.n_scales = n_scales;
    int64_t poff = 0, eoff = 0;
    int hs_max = 0;
    for (int i = 0; i < n_scales; ++i) {
        HostGeom hg;
        int s = loss_geom(B, N, scales[i], hops[i], &hg);
        if (s) return s;
        fin.off[i] = poff;
        fin.cnt[i] = (int64_t)hg.tiles * B;
        fin.inv[i] = 1.0f / ((float)B * (float)(scales[i] / 2 + 1) * (float)hg.g.frames);
        poff += fin.cnt[i];
        fold.s[i] = scales[i];
        fold.off[i] = eoff;
        eoff += (int64_t)B * scales[i];
        hs_max = scales[i] / 2 > hs_max ? scales[i] / 2 : hs_max;
    }
    cudaStream_t st = (cudaStream_t)stream;
    mss_finalize_kernel<<<1, 1024, 0, st>>>(partial, loss, fin);
    int s = ddsp_launch_status();
    if (s || !d_rec) return s;
    if (d_rec_scales) {
        // per-scale gradient buffers [n_scales][B][N] (scales launched concurrently): sum + fold
        int gx = (int)ddsp_ceil_div((N & 3) == 0 ? N / 4 : N, 256);
        if (gx > 512) gx = 512;
        mss_combine_kernel<<<dim3(gx, B), 256, 0, st>>>(d_rec_scales, edge, d_rec, B, N, fold, hs_max);
        return ddsp_launch_status();
    }
    int gx = (int)ddsp_ceil_div(2 * (int64_t)hs_max + 2 >= N ? N : 2 * (int64_t)hs_max, 256);
    if (gx > 1024) gx = 1024;
    stft_fold_kernel<<<dim3(gx, B), 256, 0, st>>>(edge, d_rec, B, N, fold, hs_max);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_stft_mag_fwd(const float *signal, const float *window, const float *twiddle,
                                      int n_tab, const float *stage_twiddle, float *mag, int B, int64_t N,
                                      int n_fft, int hop, void *stream) {
    DDSP_REQUIRE(signal && window && twiddle && mag && B > 0 && B <= 65535);
    DDSP_REQUIRE(n_tab >= n_fft && n_tab % n_fft == 0);
    HostGeom hg;
    if (reg_path(n_fft) && stage_twiddle) {
        int s = loss_geom(B, N, n_fft, hop, &hg);
        if (s) return s;
        return dispatch_reg(kMagFwd, nullptr, signal, window, reinterpret_cast<const float2 *>(stage_twiddle),
                            nullptr, nullptr, nullptr, mag, hg, B, 0, 0.f, (cudaStream_t)stream);
    }
    int s = make_geom(N, n_fft, hop, false, &hg);
    if (s) return s;
    if ((s = set_smem(stft_mag_fwd_kernel, hg.smem_mag_fwd))) return s;
    stft_mag_fwd_kernel<<<dim3(hg.tiles, B), kStftThreads, hg.smem_mag_fwd, (cudaStream_t)stream>>>(
        signal, window, reinterpret_cast<const float2 *>(twiddle), n_tab / n_fft, mag, hg.g);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_stft_mag_bwd(const float *signal, const float *d_mag, const float *window,
                                      const float *twiddle, int n_tab, const float *stage_twiddle,
                                      float *d_signal, float *edge, int B, int64_t N, int n_fft, int hop,
                                      int accumulate, void *stream) {
    DDSP_REQUIRE(signal && d_mag && window && twiddle && d_signal && edge && B > 0 && B <= 65535);
    DDSP_REQUIRE(n_tab >= n_fft && n_tab % n_fft == 0);
    HostGeom hg;
    if (reg_path(n_fft) && stage_twiddle) {
        int s = loss_geom(B, N, n_fft, hop, &hg);
        if (s) return s;
        return dispatch_reg(kMagBwd, nullptr, signal, window, reinterpret_cast<const float2 *>(stage_twiddle),
                            nullptr, d_signal, edge, const_cast<float *>(d_mag), hg, B, accumulate, 0.f,
                            (cudaStream_t)stream);
    }
    int s = make_geom(N, n_fft, hop, true, &hg);
    if (s) return s;
    if ((s = set_smem(stft_mag_bwd_kernel, hg.smem_mag_bwd))) return s;
    stft_mag_bwd_kernel<<<dim3(hg.tiles, B), kStftThreads, hg.smem_mag_bwd, (cudaStream_t)stream>>>(
        signal, d_mag, window, reinterpret_cast<const float2 *>(twiddle), n_tab / n_fft, d_signal, edge,
        hg.g, accumulate);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_stft_fold_edges(const float *edge, float *d_signal, int B, int64_t N,
                                         const int *scales, int n_scales, void *stream) {
    DDSP_REQUIRE(edge && d_signal && scales && n_scales > 0 && n_scales <= 8 && B > 0 && B <= 65535);
    FoldArgs fold;
    fold.n_scales = n_scales;
    int64_t eoff = 0;
    for (int i = 0; i < n_scales; ++i) {
        fold.s[i] = scales[i];
        fold.off[i] = eoff;
        eoff += (int64_t)B * scales[i];
    }
    int hs_max = 0;
    for (int i = 0; i < n_scales; ++i) hs_max = scales[i] / 2 > hs_max ? scales[i] / 2 : hs_max;
    int gx = (int)ddsp_ceil_div(2 * (int64_t)hs_max + 2 >= N ? N : 2 * (int64_t)hs_max, 256);
    if (gx > 1024) gx = 1024;
    stft_fold_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(edge, d_signal, B, N, fold, hs_max);
    return ddsp_launch_status();
}
