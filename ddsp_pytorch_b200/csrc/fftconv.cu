// K3: long FFT convolution (SURVEY 8a rows a9 generic, a10 reverb, and their backward).
//
// Reference path replaced:
//   ddsp/core.py:169-176          fft_convolve: pad to 2N, rfft x2, complex multiply, irfft, slice
//   ddsp/models/modules.py:21-35  Reverb.build_impulse / forward (IR padded to the signal, fft_convolve)
// i.e. three cuFFT launches of length 2N = 128000 / 384000 (non power of two) plus pads and slices.
//
// fft_convolve is a causal linear convolution truncated to the signal length, so any
// transform length n >= N + L - 1 gives the same result.  We use n = n1*n2 = 2^17 / 2^18 and the
// four-step decomposition with both factors' sub-FFTs in shared memory:
//   i = i1*n2 + i2  (time),  k = k1 + n1*k2  (frequency)
//   A: column FFTs over i1 (n1 points, CW adjacent columns per CTA) * W_n^(i2*k1)   -> work[k1][i2]
//   B: row FFT over i2 (n2 points) -> spectrum row k1; multiply by the filter spectrum (same
//      layout, never transposed); inverse row FFT; * W_n^(-i2*k1)                    (one kernel)
//   C: inverse column FFTs over k1, scale 1/n, write the first N samples
// Two real voices ride in one complex transform (re = voice 2p, im = voice 2p+1): a real filter acts
// on both parts independently, so no Hermitian untangling is ever needed.  For the gradient of the
// filter, sum_pairs G_p * conj(X_p) has the wanted cross-correlation in its real part.
#include <atomic>

#include "fft.cuh"

namespace {

constexpr int kCW = 8;            // adjacent columns per CTA in the column passes (64 B runs)
constexpr int kColThreads = 256;
constexpr int kRowThreads = 128;

__global__ void twiddle_kernel(float2 *__restrict__ tab, int n) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)n, &s, &c);
    tab[m] = make_float2((float)c, (float)(-s));
}

__device__ __forceinline__ int col_pitch(int n1) { return fpad_size(n1) + 1; }   // odd: columns land on distinct banks

// A: x rows -> work[slot][k1][i2]
__global__ void __launch_bounds__(kColThreads)
cols_fwd_kernel(const float *__restrict__ x, int64_t rows, int64_t len, int pair,
                float2 *__restrict__ work, const float2 *__restrict__ tw, int n1, int lg1, int n2) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = col_pitch(n1);
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + kCW * pitch;
    const int tid = threadIdx.x;
    const int64_t p = blockIdx.y;
    const int c0 = blockIdx.x * kCW;
    const int64_t rre = pair ? 2 * p : p, rim = pair ? 2 * p + 1 : rows;
    const float *xr = x + rre * len;
    const float *xi = rim < rows ? x + rim * len : nullptr;
    for (int idx = tid; idx < kCW * n1; idx += kColThreads) {
        const int i1 = idx / kCW, c = idx - i1 * kCW;
        const int64_t pos = (int64_t)i1 * n2 + c0 + c;
        float2 v = make_float2(0.f, 0.f);
        if (pos < len) {
            v.x = __ldg(xr + pos);
            if (xi) v.y = __ldg(xi + pos);
        }
        bufA[c * pitch + fpad(i1)] = v;
    }
    __syncthreads();
    const float2 *Z = cta_fft<false>(bufA, bufB, pitch, kCW, n1, lg1, tw, n2, tid, kColThreads);
    float2 *w = work + (size_t)p * n1 * n2;
    for (int idx = tid; idx < kCW * n1; idx += kColThreads) {
        const int k1 = idx / kCW, c = idx - k1 * kCW;
        const float2 t = __ldg(tw + (size_t)(c0 + c) * k1);
        w[(size_t)k1 * n2 + c0 + c] = cmul(Z[c * pitch + fpad(k1)], t);
    }
}

// C: work[slot][k1][i2] -> real rows (first len samples), scaled by 1/n
__global__ void __launch_bounds__(kColThreads)
cols_inv_kernel(const float2 *__restrict__ work, float *__restrict__ out, int64_t rows, int64_t len,
                int pair, const float2 *__restrict__ tw, int n1, int lg1, int n2) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = col_pitch(n1);
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + kCW * pitch;
    const int tid = threadIdx.x;
    const int64_t p = blockIdx.y;
    const int c0 = blockIdx.x * kCW;
    const float2 *w = work + (size_t)p * n1 * n2;
    for (int idx = tid; idx < kCW * n1; idx += kColThreads) {
        const int k1 = idx / kCW, c = idx - k1 * kCW;
        bufA[c * pitch + fpad(k1)] = w[(size_t)k1 * n2 + c0 + c];
    }
    __syncthreads();
    const float2 *Y = cta_fft<true>(bufA, bufB, pitch, kCW, n1, lg1, tw, n2, tid, kColThreads);
    const float sc = 1.0f / ((float)n1 * (float)n2);
    const int64_t rre = pair ? 2 * p : p, rim = pair ? 2 * p + 1 : rows;
    float *yr = out + rre * len;
    float *yi = rim < rows ? out + rim * len : nullptr;
    for (int idx = tid; idx < kCW * n1; idx += kColThreads) {
        const int i1 = idx / kCW, c = idx - i1 * kCW;
        const int64_t pos = (int64_t)i1 * n2 + c0 + c;
        if (pos < len) {
            const float2 v = Y[c * pitch + fpad(i1)];
            yr[pos] = v.x * sc;
            if (yi) yi[pos] = v.y * sc;
        }
    }
}

// B, three flavours on one row (k1) of one slot:
//   MODE 0 spectrum : row FFT, write
//   MODE 1 filter   : row FFT, * H (or conj H), inverse row FFT, * W_n^(-i2 k1), write in place
template <int MODE>
__global__ void __launch_bounds__(kRowThreads)
rows_kernel(float2 *__restrict__ work, const float2 *__restrict__ hspec, int64_t h_slot_stride,
            int conj_h, const float2 *__restrict__ tw, int n1, int n2, int lg2) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = fpad_size(n2);
    float2 *bufA = reinterpret_cast<float2 *>(smem);
    float2 *bufB = bufA + pitch;
    const int tid = threadIdx.x;
    const int k1 = blockIdx.x;
    const int64_t p = blockIdx.y;
    float2 *row = work + ((size_t)p * n1 + k1) * n2;
    for (int i = tid; i < n2; i += kRowThreads) bufA[fpad(i)] = row[i];
    __syncthreads();
    float2 *S = cta_fft<false>(bufA, bufB, pitch, 1, n2, lg2, tw, n1, tid, kRowThreads);
    if (MODE == 0) {
        for (int i = tid; i < n2; i += kRowThreads) row[i] = S[fpad(i)];
        return;
    }
    float2 *other = (S == bufA) ? bufB : bufA;
    const float2 *h = hspec + (size_t)p * h_slot_stride + (size_t)k1 * n2;
    for (int i = tid; i < n2; i += kRowThreads) {
        const float2 hv = __ldg(h + i);
        S[fpad(i)] = conj_h ? cmulc(S[fpad(i)], hv) : cmul(S[fpad(i)], hv);
    }
    __syncthreads();
    const float2 *Y = cta_fft<true>(S, other, pitch, 1, n2, lg2, tw, n1, tid, kRowThreads);
    for (int i = tid; i < n2; i += kRowThreads) {
        float2 t = __ldg(tw + (size_t)i * k1);
        t.y = -t.y;
        row[i] = cmul(Y[fpad(i)], t);
    }
}

// Correlation rows: out[k1] = IFFT_row( sum_slots FFT_row(G) * conj(FFT_row(X)) ) * W_n^(-i2 k1).
// reduce != 0: one output slot (sum over all slots); else one output slot per input slot.
__global__ void __launch_bounds__(kRowThreads)
rows_corr_kernel(const float2 *__restrict__ work_g, const float2 *__restrict__ work_x, int64_t slots,
                 int reduce, float2 *__restrict__ out, const float2 *__restrict__ tw, int n1, int n2,
                 int lg2) {
    extern __shared__ __align__(16) float smem[];
    const int pitch = fpad_size(n2);
    float2 *bufA = reinterpret_cast<float2 *>(smem);          // [2][pitch]
    float2 *bufB = bufA + 2 * pitch;                          // [2][pitch]
    float2 *acc = bufB + 2 * pitch;                           // [pitch]
    float2 *scr = acc + pitch;                                // [pitch]
    const int tid = threadIdx.x;
    const int k1 = blockIdx.x;
    const int64_t p0 = reduce ? 0 : blockIdx.y, p1 = reduce ? slots : p0 + 1;
    for (int i = tid; i < n2; i += kRowThreads) acc[fpad(i)] = make_float2(0.f, 0.f);
    for (int64_t p = p0; p < p1; ++p) {
        const float2 *rg = work_g + ((size_t)p * n1 + k1) * n2;
        const float2 *rx = work_x + ((size_t)p * n1 + k1) * n2;
        __syncthreads();
        for (int i = tid; i < n2; i += kRowThreads) {
            bufA[fpad(i)] = rg[i];
            bufA[pitch + fpad(i)] = rx[i];
        }
        __syncthreads();
        const float2 *S = cta_fft<false>(bufA, bufB, pitch, 2, n2, lg2, tw, n1, tid, kRowThreads);
        for (int i = tid; i < n2; i += kRowThreads) {          // same thread owns acc[i] throughout
            const float2 v = cmulc(S[fpad(i)], S[pitch + fpad(i)]);
            acc[fpad(i)] = cadd(acc[fpad(i)], v);
        }
    }
    __syncthreads();
    const float2 *Y = cta_fft<true>(acc, scr, pitch, 1, n2, lg2, tw, n1, tid, kRowThreads);
    float2 *row = out + ((size_t)(reduce ? 0 : p0) * n1 + k1) * n2;
    for (int i = tid; i < n2; i += kRowThreads) {
        float2 t = __ldg(tw + (size_t)i * k1);
        t.y = -t.y;
        row[i] = cmul(Y[fpad(i)], t);
    }
}

// ---- Reverb impulse (modules.py:21-26) -------------------------------------------------------
__device__ __forceinline__ float softplusf(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void reverb_impulse_fwd_kernel(const float *__restrict__ noise, const float *__restrict__ decay,
                                          const float *__restrict__ wet, const float *__restrict__ t,
                                          float *__restrict__ impulse, int L) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const float sp = softplusf(-decay[0]);
    const float sg = sigmoidf_(wet[0]);
    impulse[l] = l == 0 ? 1.f : noise[l] * expf(-sp * t[l] * 500.f) * sg;
}

__global__ void __launch_bounds__(1024)
reverb_impulse_bwd_kernel(const float *__restrict__ d_imp, int Lvalid, const float *__restrict__ noise,
                          const float *__restrict__ decay, const float *__restrict__ wet,
                          const float *__restrict__ t, float *__restrict__ d_noise,
                          float *__restrict__ d_decay, float *__restrict__ d_wet, int L) {
    __shared__ double r0[1024], r1[1024];
    const int tid = threadIdx.x;
    const float dec = decay[0];
    const float sp = softplusf(-dec);
    const float sg = sigmoidf_(wet[0]);
    const float sneg = sigmoidf_(-dec);                 // -d softplus(-decay)/d decay
    double aw = 0.0, ad = 0.0;
    for (int l = tid; l < L; l += 1024) {
        float dn = 0.f;
        if (l >= 1 && l < Lvalid) {
            const float g = d_imp[l];
            const float env = expf(-sp * t[l] * 500.f);
            dn = g * env * sg;
            const float base = g * noise[l] * env * sg;   // g * impulse[l]
            aw += (double)(base * (1.f - sg));
            ad += (double)(base * t[l] * 500.f * sneg);
        }
        d_noise[l] = dn;
    }
    r0[tid] = aw;
    r1[tid] = ad;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) { r0[tid] += r0[tid + o]; r1[tid] += r1[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) { d_wet[0] = (float)r0[0]; d_decay[0] = (float)r1[0]; }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

bool plan_ok(int n1, int n2) {
    auto p2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    return p2(n1) && p2(n2) && n1 >= 4 && n1 <= 1024 && n2 >= kCW && n2 <= 4096;
}

size_t col_smem(int n1) { return 2 * (size_t)kCW * (fpad_size(n1) + 1) * sizeof(float2); }

}  // namespace

extern "C" int ddsp_b200_twiddle_table(float *table, int n, void *stream) {
    DDSP_REQUIRE(table && n > 0 && (n & (n - 1)) == 0);
    twiddle_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2 *>(table), n);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_conv_plan(int64_t min_len, int *n1, int *n2) {
    DDSP_REQUIRE(n1 && n2 && min_len >= 1);
    int lg = ddsp_ilog2(min_len);
    if (lg < 5) lg = 5;
    const int lg1 = lg / 2, lg2 = lg - lg1;
    *n1 = 1 << lg1;
    *n2 = 1 << lg2;
    return plan_ok(*n1, *n2) ? DDSP_B200_OK : DDSP_B200_EUNSUPPORTED;
}

extern "C" int ddsp_b200_fft4_cols_fwd(const float *x, int64_t rows, int64_t len, int pair, float *work,
                                       const float *twiddle, int n1, int n2, void *stream) {
    DDSP_REQUIRE(x && work && twiddle && rows > 0 && len > 0 && plan_ok(n1, n2));
    DDSP_REQUIRE(len <= (int64_t)n1 * n2);
    const int64_t slots = pair ? (rows + 1) / 2 : rows;
    DDSP_REQUIRE(slots <= 65535);
    const size_t smem = col_smem(n1);
    int s = set_smem(cols_fwd_kernel, smem);
    if (s) return s;
    cols_fwd_kernel<<<dim3(n2 / kCW, (unsigned)slots), kColThreads, smem, (cudaStream_t)stream>>>(
        x, rows, len, pair, reinterpret_cast<float2 *>(work), reinterpret_cast<const float2 *>(twiddle),
        n1, ddsp_ilog2(n1), n2);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_fft4_cols_inv(const float *work, float *out, int64_t rows, int64_t len, int pair,
                                       const float *twiddle, int n1, int n2, void *stream) {
    DDSP_REQUIRE(work && out && twiddle && rows > 0 && len > 0 && plan_ok(n1, n2));
    DDSP_REQUIRE(len <= (int64_t)n1 * n2);
    const int64_t slots = pair ? (rows + 1) / 2 : rows;
    DDSP_REQUIRE(slots <= 65535);
    const size_t smem = col_smem(n1);
    int s = set_smem(cols_inv_kernel, smem);
    if (s) return s;
    cols_inv_kernel<<<dim3(n2 / kCW, (unsigned)slots), kColThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(work), out, rows, len, pair,
        reinterpret_cast<const float2 *>(twiddle), n1, ddsp_ilog2(n1), n2);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_fft4_rows_spectrum(float *work, int64_t slots, const float *twiddle, int n1,
                                            int n2, void *stream) {
    DDSP_REQUIRE(work && twiddle && slots > 0 && slots <= 65535 && plan_ok(n1, n2));
    const size_t smem = 2 * (size_t)fpad_size(n2) * sizeof(float2);
    int s = set_smem(rows_kernel<0>, smem);
    if (s) return s;
    rows_kernel<0><<<dim3(n1, (unsigned)slots), kRowThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<float2 *>(work), nullptr, 0, 0, reinterpret_cast<const float2 *>(twiddle), n1, n2,
        ddsp_ilog2(n2));
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_fft4_rows_filter(float *work, int64_t slots, const float *hspec,
                                          int64_t h_slot_stride, int conj_h, const float *twiddle, int n1,
                                          int n2, void *stream) {
    DDSP_REQUIRE(work && hspec && twiddle && slots > 0 && slots <= 65535 && plan_ok(n1, n2));
    const size_t smem = 2 * (size_t)fpad_size(n2) * sizeof(float2);
    int s = set_smem(rows_kernel<1>, smem);
    if (s) return s;
    rows_kernel<1><<<dim3(n1, (unsigned)slots), kRowThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<float2 *>(work), reinterpret_cast<const float2 *>(hspec), h_slot_stride, conj_h,
        reinterpret_cast<const float2 *>(twiddle), n1, n2, ddsp_ilog2(n2));
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_fft4_rows_correlate(const float *work_g, const float *work_x, int64_t slots,
                                             int reduce, float *out, const float *twiddle, int n1, int n2,
                                             void *stream) {
    DDSP_REQUIRE(work_g && work_x && out && twiddle && slots > 0 && slots <= 65535 && plan_ok(n1, n2));
    const size_t smem = 6 * (size_t)fpad_size(n2) * sizeof(float2);
    int s = set_smem(rows_corr_kernel, smem);
    if (s) return s;
    rows_corr_kernel<<<dim3(n1, reduce ? 1u : (unsigned)slots), kRowThreads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(work_g), reinterpret_cast<const float2 *>(work_x), slots, reduce,
        reinterpret_cast<float2 *>(out), reinterpret_cast<const float2 *>(twiddle), n1, n2, ddsp_ilog2(n2));
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_reverb_impulse_fwd(const float *noise, const float *decay, const float *wet,
                                            const float *t, float *impulse, int L, void *stream) {
    DDSP_REQUIRE(noise && decay && wet && t && impulse && L > 0);
    reverb_impulse_fwd_kernel<<<(L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(noise, decay, wet, t,
                                                                                 impulse, L);
    return ddsp_launch_status();
}

extern "C" int ddsp_b200_reverb_impulse_bwd(const float *d_impulse, int Lvalid, const float *noise,
                                            const float *decay, const float *wet, const float *t,
                                            float *d_noise, float *d_decay, float *d_wet, int L,
                                            void *stream) {
    DDSP_REQUIRE(d_impulse && noise && decay && wet && t && d_noise && d_decay && d_wet && L > 0);
    DDSP_REQUIRE(Lvalid >= 0 && Lvalid <= L);
    reverb_impulse_bwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_impulse, Lvalid, noise, decay, wet, t,
                                                                    d_noise, d_decay, d_wet, L);
    return ddsp_launch_status();
}

static std::atomic<unsigned long long> g_launches{0};
void ddsp_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t ddsp_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int ddsp_b200_abi_version(void) { return 1; }

extern "C" const char *ddsp_b200_strerror(int status) {
    if (status == DDSP_B200_OK) return "ok";
    if (status == DDSP_B200_EINVAL) return "invalid argument (null pointer or bad size)";
    if (status == DDSP_B200_EUNSUPPORTED) return "shape not supported by the sm_100a kernels";
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown ddsp_b200 status";
}
