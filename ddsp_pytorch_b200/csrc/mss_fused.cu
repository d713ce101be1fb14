// K4L, second generation: the multi-scale spectral loss and its gradient w.r.t. the reconstruction, ALL
// scales in ONE launch (SURVEY 8a rows a11 + a12 + their backward).
//
// Reference path replaced (same as stft.cu):
//   ddsp/core.py:27-41     multiscale_fft (torch.stft per scale, reflect pad, periodic hann, normalized, abs)
//   train.py:70-76,92-103  multiscale_spec_loss (lin + log L1 per scale)
//   and the autograd backward of both.
//
// Design.  Valid for hop = n_fft/4 (overlap 0.75, the reference's only setting), n_fft = 64..4096.
//   * Work item = (scale, voice, tile of FT consecutive FRAMES); one CTA per item, items ordered by scale (largest
//     first) so the hardware block scheduler balances them over the SMs.  Every frame is transformed exactly once
//     (the first generation recomputed the 3 overlap frames of every tile: +33 % at 4096).
//   * A group of T = n_fft/16 threads owns TWO adjacent frames at a time, one per lane of the packed f32x2
//     registers (pfft.cuh): rec + i*target of frame A in lane 0, of frame B in lane 1.  Frame B's samples are frame
//     A's shifted by one hop = 4 register slots, so 20 loads per signal serve both frames.
//   * After the last forward stage the spectra stay in registers.  Bin k needs Z[k] and Z[n_fft-k]; the mirror
//     lives in thread T-t, so the upper half of each thread's bins goes through shared memory once (8 of 16 slots),
//     the loss terms and gradient spectra are computed in registers, the mirrored half of the gradient spectrum
//     goes back the same way, and the inverse transform (one complex FFT carries the real gradients of BOTH
//     frames: U_A + i U_B) starts from registers.
//   * Overlap-add is a gather: the batch's gradient frames are parked in shared memory, every padded-signal position
//     sums its <= 4 frames in frame order (+ the carry of the previous batch), finished positions go to the scale's
//     gradient plane P, the last 3 hops are carried.  The carry of a tile's last batch goes to a small halo buffer
//     that mss_combine2_kernel adds to the head of the next tile: no atomics, bit-reproducible.
//   * mss_combine2_kernel sums the scales in order and folds the reflect padding back.
#include "common.cuh"
#include "pfft.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxScales = 8;
constexpr int kWorkBytes = kThreads * 16 * 17;           // 256 threads x 16 points x 16 B, padded 17/16

struct ScaleDesc {
    int lg, hop, frames, ft_log;       // frames per tile = 1 << ft_log
    int tiles;                          // tiles per voice
    int item0;                          // first work item of this scale
    int last_end;                       // first frame slot not finished by the last tile's batches
    int rowlen;                         // floats per voice in P (multiple of 4)
    long long p_off;                    // float offset of P      [B][rowlen]
    long long h_off;                    // float offset of halos  [B][tiles][3*hop]
    long long part_off;                 // pair offset of the loss partials [tiles][B]
    const float *window;
    const float2 *tw;
    float inv_cnt;
    int pad_;
};

struct FusedArgs {
    int n_scales, B, n_items, pad_;
    long long N;
    ScaleDesc sc[kMaxScales];
};

template <int T> __device__ __forceinline__ void gsync(int grp) {
    if (T <= 32) __syncwarp();
    else if (T == kThreads) __syncthreads();
    else if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
    else if (grp == 1) asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory");
    else if (grp == 2) asm volatile("bar.sync 3, %0;" ::"n"(T) : "memory");
    else asm volatile("bar.sync 4, %0;" ::"n"(T) : "memory");
}

// sign(d) in {-1, 0, +1}
__device__ __forceinline__ float sgn3(float d) { return (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f); }

// Loss terms and gradient spectra of bin k = t + Q*T (Q < 8; Q = 8 is k = n_fft/2, thread 0 only) for the two
// frames in the lanes.  x = the thread's spectra after the last forward stage, gbuf = the parked mirror bins.
// Q is a template parameter so that x[] and zi[] are indexed statically and stay in registers.
template <int LG, bool GRAD, int Q>
__device__ __forceinline__ void bin_math(const pfft::C (&x)[16], float2 (&zi)[16], const pfft::E *gbuf, float2 *ex2,
                                         int t, float mA, float mB, float rs2, float kc, float &lin, float &lgs) {
    using namespace pfft;
    constexpr int T = Plan<LG>::T;
    constexpr int q = Q;
    const C zk = x[slot_of_q<LG>(q)];
    C zm = from_e(gbuf[q < 8 ? (8 - q) * T - t : 0]);     // Z[n_fft - k]
    const bool self = (q == 0 && t == 0) || q == 8;       // k = 0 and k = n_fft/2 mirror themselves
    if (q == 0 && t == 0) zm = zk;
    // doubled spectra: rec Y2 = Z[k] + conj Z[-k], target X2 = -i (Z[k] - conj Z[-k])
    const V yr = zk.re + zm.re, yi = zk.im - zm.im;
    const V xre = zk.im + zm.im, xim = zm.re - zk.re;
    const V yy = fma(yr, yr, yi * yi), xx = fma(xre, xre, xim * xim);
    float yyv[2], xxv[2], yrv[2], yiv[2], c[2];
    get(yy, yyv[0], yyv[1]);
    get(xx, xxv[0], xxv[1]);
    get(yr, yrv[0], yrv[1]);
    get(yi, yiv[0], yiv[1]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float own = e ? mB : mA;
        const float ry = rsqrtf(fmaxf(yyv[e], 1e-37f));                  // 1 / |Y2|
        const float sy = yyv[e] * ry * rs2;                              // |Y| / sqrt(n_fft)
        const float sx = xxv[e] * rsqrtf(fmaxf(xxv[e], 1e-37f)) * rs2;
        const float d = sy - sx;
        const float iy = __fdividef(1.0f, sy + 1e-7f);
        lin = fmaf(own, fabsf(d), lin);
        lgs = fmaf(own, fabsf(__log2f((sx + 1e-7f) * iy)), lgs);         // * ln 2 at the end
        // log is monotonic: sign(log(sy+eps) - log(sx+eps)) == sign(sy - sx)
        const float sg = sgn3(d);
        float cc = fmaf(sg, iy, sg) * kc * ry;                           // dL/dY2 = cc * Y2
        cc = (own != 0.f && yyv[e] > 0.f) ? cc : 0.f;
        c[e] = self ? cc : 0.5f * cc;                                    // (U_A + i U_B) / 2 for interior bins
    }
    if (GRAD) {
        const float pa = c[0] * yrv[0], qa = c[0] * yiv[0];              // U_A
        const float2 own_bin = self ? make_float2(pa, c[1] * yrv[1])     // real bins: (U_A, U_B)
                                    : make_float2(fmaf(-c[1], yiv[1], pa), fmaf(c[1], yrv[1], qa));
        zi[q] = own_bin;
        if (q == 8) ex2[0] = own_bin;                                    // keeps the read of the upper slots uniform
        else if (!self) ex2[(8 - q) * T - t] = make_float2(fmaf(c[1], yiv[1], pa), fmaf(c[1], yrv[1], -qa));
    }
}

template <int LG, bool GRAD>
__device__ __forceinline__ void tile_body(const ScaleDesc &sc, int b, int tile, int B, int Ni,
                                          const float *__restrict__ target, const float *__restrict__ rec,
                                          float *__restrict__ ws, float *__restrict__ partial,
                                          unsigned char *smem) {
    using namespace pfft;
    using P = Plan<LG>;
    constexpr int N = P::N, T = P::T, G = kThreads / T, NFB = 2 * G, HOP = N / 4, HS = N / 2;
    constexpr int LH = LG - 2;                                    // log2(hop)
    constexpr int GBYTES = (N + N / 16) * 16;                     // bytes of one group's work buffer
    constexpr int PLANE_OFF = (N + N / 16) * 8;                   // gradient frames parked here (after the inverse buffer)
    constexpr int EX2_OFF = 13 * N;                               // mirrored gradient bins
    static_assert(kThreads % T == 0 && G * GBYTES == kWorkBytes, "work buffer layout");
    float *carry = reinterpret_cast<float *>(smem + kWorkBytes);  // [2][3*HOP]
    __shared__ float red[2][kThreads / 32];

    const int tid = threadIdx.x;
    const int grp = tid / T, t = tid - grp * T;
    unsigned char *gbase = smem + (size_t)grp * GBYTES;
    E *gbuf = reinterpret_cast<E *>(gbase);
    float2 *ibuf = reinterpret_cast<float2 *>(gbase);
    float2 *ex2 = reinterpret_cast<float2 *>(gbase + EX2_OFF);
    float *plane = reinterpret_cast<float *>(gbase + PLANE_OFF);

    const float *xr = rec + (size_t)b * Ni;
    const float *xt = target + (size_t)b * Ni;
    const float *__restrict__ window = sc.window;
    const float2 *__restrict__ tw = sc.tw;
    const int FT = 1 << sc.ft_log;
    const int f0 = tile << sc.ft_log;
    const int f1 = min(f0 + FT, sc.frames);
    const float rs = rsqrtf((float)N);
    const float rs2 = 0.5f * rs;                                  // spectra are kept doubled (no 1/2 in the untangle)
    const float kc = sc.inv_cnt * rs;
    float *Pb = ws + sc.p_off + (size_t)b * sc.rowlen;
    float *Hb = ws + sc.h_off + ((size_t)b * sc.tiles + tile) * (3 * HOP);

    if (GRAD)
        for (int i = tid; i < 3 * HOP; i += kThreads) carry[i] = 0.f;
    int cb = 0;
    float lin = 0.f, lgs = 0.f;

    for (int fb = f0; fb < f1; fb += NFB) {
        const int fA = fb + 2 * grp;
        const float mA = fA < f1 ? 1.f : 0.f, mB = fA + 1 < f1 ? 1.f : 0.f;
        C x[16];
        {
            // ---- windowed frames A (lane 0) and B = A + 1 (lane 1): sample slot q of B is slot q + 4 of A
            float r[20], g[20];
            const int start = fA * HOP - HS;
            if (start >= 0 && start + HOP + N <= Ni) {
#pragma unroll
                for (int q = 0; q < 20; ++q) {
                    r[q] = __ldg(xr + start + t + q * T);
                    g[q] = __ldg(xt + start + t + q * T);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 20; ++q) {
                    int m = start + t + q * T;
                    m = m < 0 ? -m : m;
                    m = m >= Ni ? 2 * (Ni - 1) - m : m;
                    m = min(max(m, 0), Ni - 1);                  // only frames that do not exist reach this clamp
                    r[q] = __ldg(xr + m);
                    g[q] = __ldg(xt + m);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float w = __ldg(window + t + q * T);
                const float wa = w * mA, wb = w * mB;
                x[q].re = mk(r[q] * wa, r[q + 4] * wb);
                x[q].im = mk(g[q] * wa, g[q + 4] * wb);
            }
        }
        // ---- forward transform of both frames; the last stage stays in registers
        stage_compute_store<LG, 0, false>(x, gbuf, t, tw);
        gsync<T>(grp);
        stage_load<LG, 1>(x, gbuf, t);
        if (P::STAGES == 3) {
            gsync<T>(grp);
            stage_compute_store<LG, 1, false>(x, gbuf, t, tw);
            gsync<T>(grp);
            stage_load<LG, 2>(x, gbuf, t);
            stage_compute_regs<LG, 2, false>(x, t, tw);
        } else {
            stage_compute_regs<LG, 1, false>(x, t, tw);
        }
        gsync<T>(grp);                                            // every load of the last stage is done
        // ---- mirror exchange: slot q >= 8 holds bin t + qT >= n_fft/2; park it at index (bin - n_fft/2)
#pragma unroll
        for (int q = 8; q < 16; ++q) gbuf[(q - 8) * T + t] = to_e(x[slot_of_q<LG>(q)]);
        gsync<T>(grp);
        // ---- per bin k = t + qT (q < 8; thread 0 also takes k = n_fft/2): loss terms and gradient spectra
        float2 zi[16];
        bin_math<LG, GRAD, 0>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 1>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 2>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 3>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 4>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 5>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 6>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        bin_math<LG, GRAD, 7>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        if (t == 0) bin_math<LG, GRAD, 8>(x, zi, gbuf, ex2, t, mA, mB, rs2, kc, lin, lgs);
        if (GRAD) {
            gsync<T>(grp);
#pragma unroll
            for (int q = 8; q < 16; ++q) zi[q] = ex2[(q - 8) * T + t];
            // ---- inverse transform of U_A + i U_B (scalar registers: one transform per thread here)
            regfft::stage_compute_store<LG, 0, true>(zi, ibuf, t, tw);
            gsync<T>(grp);
            regfft::stage_load<LG, 1>(zi, ibuf, t);
            if (P::STAGES == 3) {
                gsync<T>(grp);
                regfft::stage_compute_store<LG, 1, true>(zi, ibuf, t, tw);
                gsync<T>(grp);
                regfft::stage_load<LG, 2>(zi, ibuf, t);
                regfft::stage_compute_regs<LG, 2, true>(zi, t, tw);
            } else {
                regfft::stage_compute_regs<LG, 1, true>(zi, t, tw);
            }
            // ---- park the two real gradient frames (unwindowed) for the gather
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float2 v = zi[slot_of_q<LG>(q)];
                plane[t + q * T] = v.x;
                plane[N + t + q * T] = v.y;
            }
            __syncthreads();
            // ---- ordered gather overlap-add over the batch's (NFB + 3) hops, four positions per thread
            const float *cold = carry + cb * (3 * HOP);
            float *cnew = carry + (cb ^ 1) * (3 * HOP);
            for (int v = tid; v < (NFB + 3) * (HOP / 4); v += kThreads) {
                const int rel = v * 4, j = rel >> LH, n0 = rel & (HOP - 1);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < 3) acc = *reinterpret_cast<const float4 *>(cold + j * HOP + n0);
#pragma unroll
                for (int dq = 3; dq >= 0; --dq) {                 // frames j-3 .. j, oldest first
                    const int qf = j - dq;
                    if (qf >= 0 && qf < NFB) {
                        const int n = n0 + dq * HOP;
                        const float *pl = reinterpret_cast<const float *>(smem + (size_t)(qf >> 1) * GBYTES + PLANE_OFF) +
                                          (qf & 1) * N;
                        const float4 xv = *reinterpret_cast<const float4 *>(pl + n);
                        const float4 wv = __ldg(reinterpret_cast<const float4 *>(window + n));
                        acc.x = fmaf(wv.x, xv.x, acc.x);
                        acc.y = fmaf(wv.y, xv.y, acc.y);
                        acc.z = fmaf(wv.z, xv.z, acc.z);
                        acc.w = fmaf(wv.w, xv.w, acc.w);
                    }
                }
                if (j < NFB) *reinterpret_cast<float4 *>(Pb + (size_t)fb * HOP + rel) = acc;
                else *reinterpret_cast<float4 *>(cnew + (j - NFB) * HOP + n0) = acc;
            }
            cb ^= 1;
            __syncthreads();
        } else {
            gsync<T>(grp);                                        // exchange area is reused by the next batch
        }
    }
    if (GRAD) {
        // tail of the tile: belongs to the head of the next tile (or to the end of the signal): halo buffer
        const float *cold = carry + cb * (3 * HOP);
        for (int i = tid; i < 3 * HOP / 4; i += kThreads)
            reinterpret_cast<float4 *>(Hb)[i] = reinterpret_cast<const float4 *>(cold)[i];
    }
    lin = ddsp_warp_sum(lin);
    lgs = ddsp_warp_sum(lgs * 0.69314718055994530942f);
    if ((tid & 31) == 0) { red[0][tid >> 5] = lin; red[1][tid >> 5] = lgs; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; c += red[1][i]; }
        float *pp = partial + 2 * (sc.part_off + (size_t)tile * B + b);
        pp[0] = a;
        pp[1] = c;
    }
}

template <bool GRAD>
__global__ void __launch_bounds__(kThreads, 2)
mss_fused_kernel(const float *__restrict__ target, const float *__restrict__ rec, float *__restrict__ ws,
                 float *__restrict__ partial, const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int item = blockIdx.x;
    int si = 0;
#pragma unroll
    for (int i = 1; i < kMaxScales; ++i)
        if (i < a.n_scales && item >= a.sc[i].item0) si = i;
    const ScaleDesc &sc = a.sc[si];
    const int local = item - sc.item0;
    const int tile = local / a.B, b = local - tile * a.B;
    const int Ni = (int)a.N;
    switch (sc.lg) {
        case 6: tile_body<6, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        case 7: tile_body<7, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        case 8: tile_body<8, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        case 9: tile_body<9, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        case 10: tile_body<10, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        case 11: tile_body<11, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
        default: tile_body<12, GRAD>(sc, b, tile, a.B, Ni, target, rec, ws, partial, smem); break;
    }
}

// value of scale s' padded gradient at padded position i (0 outside what the tiles produced)
__device__ __forceinline__ float padded_grad(const ScaleDesc &sc, const float *__restrict__ ws, int b, int i) {
    const int lh = sc.lg - 2;
    const int fi = i >> lh, n0 = i & (sc.hop - 1);
    const int tl = fi >> sc.ft_log, within = fi & ((1 << sc.ft_log) - 1);
    float v = 0.f;
    if (fi < sc.last_end) v = ws[sc.p_off + (size_t)b * sc.rowlen + i];
    if (tl >= 1 && tl < sc.tiles && within < 3)
        v += ws[sc.h_off + ((size_t)b * sc.tiles + (tl - 1)) * (3 * sc.hop) + within * sc.hop + n0];
    if (fi >= sc.last_end && fi < sc.last_end + 3)
        v += ws[sc.h_off + ((size_t)b * sc.tiles + (sc.tiles - 1)) * (3 * sc.hop) + (fi - sc.last_end) * sc.hop + n0];
    return v;
}

// d_rec[b, m] = sum over scales (fixed order) of the padded gradient at m + n_fft/2, plus the two reflections
__global__ void __launch_bounds__(256)
mss_combine2_kernel(const float *__restrict__ ws, float *__restrict__ d_rec, const __grid_constant__ FusedArgs a) {
    const int b = blockIdx.y;
    const int N = (int)a.N;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < N; m += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < a.n_scales; ++k) {
            const ScaleDesc &sc = a.sc[k];
            const int hs = 2 * sc.hop;
            acc += padded_grad(sc, ws, b, m + hs);
            if (m >= 1 && m <= hs) acc += padded_grad(sc, ws, b, hs - m);                    // left reflect pad
            if (m <= N - 2 && m >= N - 1 - hs) acc += padded_grad(sc, ws, b, 2 * (N - 1) - m + hs);   // right
        }
        d_rec[(size_t)b * N + m] = acc;
    }
}

struct FinArgs2 {
    int n_scales;
    long long off[kMaxScales], cnt[kMaxScales];
    float inv[kMaxScales];
};

__global__ void __launch_bounds__(1024)
mss_finalize2_kernel(const float *__restrict__ partial, float *__restrict__ loss, const __grid_constant__ FinArgs2 fa) {
    // every thread adds its share of every scale's partials (weighted by the scale's 1/count), then one block
    // reduction in double; fixed order -> deterministic
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = 0; i < fa.n_scales; ++i) {
        const float *p = partial + 2 * fa.off[i];
        double s = 0.0;
        for (long long j = threadIdx.x; j < fa.cnt[i]; j += blockDim.x) s += (double)p[2 * j] + (double)p[2 * j + 1];
        acc += s * (double)fa.inv[i];
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

// ---- host-side plan ------------------------------------------------------------------------------------
struct Plan2 {
    FusedArgs args;
    FinArgs2 fin;
    long long ws_floats, partial_pairs;
    size_t smem;
};

int make_plan(int B, int64_t N, const int *scales, int n_scales, Plan2 *out) {
    if (B < 1 || n_scales < 1 || n_scales > kMaxScales || N >= (1ll << 30)) return DDSP_B200_EINVAL;
    FusedArgs &a = out->args;
    a.n_scales = n_scales;
    a.B = B;
    a.N = N;
    out->fin.n_scales = n_scales;
    long long off = 0, pairs = 0;
    int item = 0, max_hop = 0;
    // about 8 tiles per resident CTA slot in total, shared evenly by the scales (their work is about equal)
    const long long want_tiles = ddsp_ceil_div(8ll * 2 * DDSP_SM_COUNT, (long long)n_scales * B);
    for (int i = 0; i < n_scales; ++i) {
        const int s = scales[i];
        if (s < 64 || s > 4096 || (s & (s - 1))) return DDSP_B200_EUNSUPPORTED;
        if (N <= s / 2) return DDSP_B200_EUNSUPPORTED;               // reflect padding needs pad < N
        ScaleDesc &d = a.sc[i];
        d.lg = 0;
        while ((1 << d.lg) < s) ++d.lg;
        d.hop = s / 4;
        d.frames = 1 + (int)(N / d.hop);
        const int nfb = 2 * (kThreads / (s / 16));
        int ft_log = 2;                                              // halo logic needs >= 3 frames per tile
        while ((1 << ft_log) < nfb) ++ft_log;
        while ((1 << (ft_log + 1)) <= d.frames && ddsp_ceil_div(d.frames, 1 << ft_log) > want_tiles) ++ft_log;
        d.ft_log = ft_log;
        d.tiles = (int)ddsp_ceil_div(d.frames, 1 << ft_log);
        d.item0 = item;
        item += d.tiles * B;
        const int last_f0 = (d.tiles - 1) << ft_log;
        d.last_end = last_f0 + (int)ddsp_ceil_div(d.frames - last_f0, nfb) * nfb;
        long long rl = (long long)d.last_end * d.hop;
        if (rl < N + s) rl = N + s;
        d.rowlen = (int)((rl + 3) & ~3ll);
        d.p_off = off;
        off += (long long)B * d.rowlen;
        d.h_off = off;
        off += (long long)B * d.tiles * 3 * d.hop;
        d.part_off = pairs;
        pairs += (long long)d.tiles * B;
        d.inv_cnt = 1.0f / ((float)B * (float)(s / 2 + 1) * (float)d.frames);
        d.window = nullptr;
        d.tw = nullptr;
        d.pad_ = 0;
        out->fin.off[i] = d.part_off;
        out->fin.cnt[i] = (long long)d.tiles * B;
        out->fin.inv[i] = d.inv_cnt;
        if (d.hop > max_hop) max_hop = d.hop;
    }
    a.n_items = item;
    a.pad_ = 0;
    out->ws_floats = off;
    out->partial_pairs = pairs;
    out->smem = (size_t)kWorkBytes + 2 * 3 * (size_t)max_hop * sizeof(float);
    return DDSP_B200_OK;
}

}  // namespace

extern "C" int ddsp_b200_mss_fused_supported(const int *scales, const int *hops, int n_scales) {
    if (!scales || !hops || n_scales < 1 || n_scales > kMaxScales) return 0;
    for (int i = 0; i < n_scales; ++i)
        if (scales[i] < 64 || scales[i] > 4096 || (scales[i] & (scales[i] - 1)) || hops[i] * 4 != scales[i]) return 0;
    return 1;
}

extern "C" int ddsp_b200_mss_fused_sizes(int B, int64_t N, const int *scales, int n_scales, int64_t *workspace_floats,
                                         int64_t *partial_floats) {
    DDSP_REQUIRE(scales && workspace_floats && partial_floats);
    Plan2 p;
    int s = make_plan(B, N, scales, n_scales, &p);
    if (s) return s;
    *workspace_floats = p.ws_floats;
    *partial_floats = 2 * p.partial_pairs;
    return DDSP_B200_OK;
}

extern "C" int ddsp_b200_mss_fused(const float *target, const float *rec, const float *windows,
                                   const float *const *stage_twiddles, float *workspace, float *partial,
                                   float *d_rec, float *loss, int B, int64_t N, const int *scales, int n_scales,
                                   void *stream) {
    DDSP_REQUIRE(target && rec && windows && stage_twiddles && partial && loss && scales);
    DDSP_REQUIRE(!d_rec || workspace);
    DDSP_REQUIRE(B <= 65535);
    Plan2 p;
    int s = make_plan(B, N, scales, n_scales, &p);
    if (s) return s;
    int64_t woff = 0;
    for (int i = 0; i < n_scales; ++i) {
        DDSP_REQUIRE(stage_twiddles[i]);
        p.args.sc[i].window = windows + woff;
        p.args.sc[i].tw = reinterpret_cast<const float2 *>(stage_twiddles[i]);
        woff += scales[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (d_rec) {
        if ((e = cudaFuncSetAttribute(mss_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem)) != cudaSuccess)
            return (int)e;
        mss_fused_kernel<true><<<p.args.n_items, kThreads, p.smem, st>>>(target, rec, workspace, partial, p.args);
    } else {
        if ((e = cudaFuncSetAttribute(mss_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem)) != cudaSuccess)
            return (int)e;
        mss_fused_kernel<false><<<p.args.n_items, kThreads, p.smem, st>>>(target, rec, workspace, partial, p.args);
    }
    if ((s = ddsp_launch_status())) return s;
    mss_finalize2_kernel<<<1, 1024, 0, st>>>(partial, loss, p.fin);
    if ((s = ddsp_launch_status()) || !d_rec) return s;
    int gx = (int)ddsp_ceil_div(N, 256);
    if (gx > 64) gx = 64;
    mss_combine2_kernel<<<dim3(gx, B), 256, 0, st>>>(workspace, d_rec, p.args);
    return ddsp_launch_status();
}
