// Persistent GRU recurrence on thread-block clusters (control-net decoder).
//
// Reference path replaced: ddsp/core.py:132-133 `nn.GRU(...)` as decoder.py:40,59,65 calls it, i.e. the
// cuDNN RNN: on B200 it runs one 64x512 x 512x1536 SGEMM launch plus one element-wise launch PER TIME
// STEP (400 + 400 launches of ~19 us and ~3 us forward, the same again backward; measured 16 ms of a
// 23 ms training step at batch 64, profiles/r01_model_step_b64.json).
//
// Design.  The input projections gi = x W_ih^T + b_ih for all time steps stay one big library GEMM.  The
// recurrence is ONE launch: a cluster of 16 CTAs (non-portable size; 7 such clusters are resident on a
// B200) keeps the whole 1536x512 fp32 W_hh in its distributed shared memory (96 rows = 32 hidden units x
// 3 gates = 192 KB per CTA, XOR-swizzled instead of padded) for all T steps and owns up to 10 voices.
// Per step each CTA computes its 96 x V gate pre-activations (thread = (unit, eighth of k), all-reduce
// over k with shuffles), applies the gates, publishes its 32 new hidden values per voice in a double
// buffer, and after ONE cluster barrier every CTA pulls the other 15 slices through DSMEM.
// Backward walks time in reverse with the same resident W_hh: per step each CTA turns dh of its units
// into the gate gradients, multiplies them with its 96 rows (partial W_hh^T d for all 512 inputs), and
// the 16 partials are reduce-scattered through DSMEM (two cluster barriers).  dW_hh, dW_ih, dx are
// library GEMMs over the stored gate gradients afterwards.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kH = 512;          // hidden size this kernel is built for (16 CTAs x 32 units)
constexpr int kC = 16;           // CTAs per cluster
constexpr int kU = 32;           // hidden units per CTA
constexpr int kV = 10;           // voices per cluster pass
constexpr int kGruThreads = 256;
constexpr int kRows = 3 * kU;    // W_hh rows resident per CTA

// XOR swizzle of the float4 index inside a 512-float row: chunk = q/16 (eighth of k), j = q%16.
// Lanes that differ in `chunk` (forward) or in `j` (backward) hit distinct bank groups.
__device__ __forceinline__ int swz(int q) { return (q & ~15) | ((q ^ (q >> 4)) & 15); }

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

struct GruSmemFwd {
    float w[kRows * kH];         // [gate*32 + u][512] swizzled
    float hfull[kV * kH];        // [v][512] swizzled
    float hown[2 * kV * kU];     // [buf][v][u]
};
struct GruSmemBwd {
    float w[kRows * kH];
    float partial[kV * kH];      // [v][512]: this CTA's rows' contribution to W_hh^T d
    float down[kRows * 24];      // [row][12 (>= kV)][2] gate gradients of the own rows, each stored twice (FFMA2 operand)
    float dhn[kV * kU];          // recurrent part of dh for the own units
};

__device__ __forceinline__ void load_w_slice(float *w, const float *__restrict__ w_hh, int rank, int tid) {
    // rows g*H + 32*rank + u  ->  w[(g*32+u)][swizzled]
    for (int i = tid; i < kRows * (kH / 4); i += kGruThreads) {
        const int row = i / (kH / 4), q = i - row * (kH / 4);
        const int g = row / kU, u = row - g * kU;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(w_hh + ((size_t)g * kH + rank * kU + u) * kH) + q);
        reinterpret_cast<float4 *>(w + (size_t)row * kH)[swz(q)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGruThreads, 1)
gru_fwd_kernel(const float *__restrict__ gi, const float *__restrict__ w_hh, const float *__restrict__ b_hh,
               const float *__restrict__ h0, float *__restrict__ y, float *__restrict__ gates, int B, int T,
               int vpc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruSmemFwd &s = *reinterpret_cast<GruSmemFwd *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank();
    const int cid = blockIdx.x / kC, ncl = gridDim.x / kC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u = warp * 4 + (lane >> 3), kq = lane & 7;
    const int col = rank * kU + u;                                  // global hidden unit of this thread

    load_w_slice(s.w, w_hh, rank, tid);
    const float br = __ldg(b_hh + col), bz = __ldg(b_hh + kH + col), bn = __ldg(b_hh + 2 * kH + col);

    for (int b0 = cid * vpc; b0 < B; b0 += ncl * vpc) {     // vpc <= kV voices per cluster pass
        const int nv = min(vpc, B - b0);
        __syncthreads();
        for (int i = tid; i < kV * (kH / 4); i += kGruThreads) {
            const int v = i / (kH / 4), q = i - v * (kH / 4);
            float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h0 && v < nv) hv = __ldg(reinterpret_cast<const float4 *>(h0 + (size_t)(b0 + v) * kH) + q);
            reinterpret_cast<float4 *>(s.hfull + (size_t)v * kH)[swz(q)] = hv;
        }
        __syncthreads();
        for (int t = 0; t < T; ++t) {
            // prefetch the input projections of the (up to two) voices this lane finishes: kq and kq + 8
            float gir[2], giz[2], gin[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int vi = kq + 8 * e;
                gir[e] = giz[e] = gin[e] = 0.f;
                if (vi < nv) {
                    const float *p = gi + ((size_t)(b0 + vi) * T + t) * 3 * kH + col;
                    gir[e] = __ldg(p);
                    giz[e] = __ldg(p + kH);
                    gin[e] = __ldg(p + 2 * kH);
                }
            }
            // ---- gh[g][v] partial over this thread's eighth of k, packed FP32: each accumulator is a pair
            //      (sum over even k, sum over odd k), one FFMA2 per W pair x h pair
            uint64_t acc2[3][kV];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int v = 0; v < kV; ++v) acc2[g][v] = 0ull;
            const ulonglong2 *w0 = reinterpret_cast<const ulonglong2 *>(s.w + (size_t)(0 * kU + u) * kH);
            const ulonglong2 *w1 = reinterpret_cast<const ulonglong2 *>(s.w + (size_t)(1 * kU + u) * kH);
            const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(s.w + (size_t)(2 * kU + u) * kH);
            const ulonglong2 *hf = reinterpret_cast<const ulonglong2 *>(s.hfull);
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
                const int q = kq * 16 + (j ^ kq);                  // == swz(kq*16 + j)
                const ulonglong2 a = w0[q], bq = w1[q], c = w2[q];
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    const ulonglong2 h = hf[v * (kH / 4) + q];
                    acc2[0][v] = fma2(a.y, h.y, fma2(a.x, h.x, acc2[0][v]));
                    acc2[1][v] = fma2(bq.y, h.y, fma2(bq.x, h.x, acc2[1][v]));
                    acc2[2][v] = fma2(c.y, h.y, fma2(c.x, h.x, acc2[2][v]));
                }
            }
            float acc[3][kV];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    float lo, hi;
                    unpk2(acc2[g][v], lo, hi);
                    acc[g][v] = lo + hi;
                }
            // ---- all-reduce over the 8 k-eighths (lane bits 0..2)
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    float x = acc[g][v];
                    x += __shfl_xor_sync(0xffffffffu, x, 1);
                    x += __shfl_xor_sync(0xffffffffu, x, 2);
                    x += __shfl_xor_sync(0xffffffffu, x, 4);
                    acc[g][v] = x;
                }
            // ---- gates for voices kq and kq+8 of unit u
            const int colq = col >> 2;                              // float4 index of the own column in hfull
            float rr[2], zz[2], nn[2], gg[2], hh[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int vi = kq + 8 * e;
                if (vi < nv) {
                    float ghr = 0.f, ghz = 0.f, ghn = 0.f;
#pragma unroll
                    for (int v = 0; v < kV; ++v)
                        if (v == vi) { ghr = acc[0][v]; ghz = acc[1][v]; ghn = acc[2][v]; }
                    ghr += br; ghz += bz; ghn += bn;
                    const float hprev = s.hfull[(size_t)vi * kH + swz(colq) * 4 + (col & 3)];
                    const float r = sigmoid_acc(gir[e] + ghr);
                    const float z = sigmoid_acc(giz[e] + ghz);
                    const float n = tanhf(fmaf(r, ghn, gin[e]));
                    const float hnew = fmaf(z, hprev - n, n);        // (1-z) n + z h
                    s.hown[((t & 1) * kV + vi) * kU + u] = hnew;
                    rr[e] = r; zz[e] = z; nn[e] = n; gg[e] = ghn; hh[e] = hnew;
                }
            }
            cluster.barrier_arrive();                               // own slice of h_t is published ...
#pragma unroll
            for (int e = 0; e < 2; ++e) {                           // ... the global stores ride under the barrier
                const int vi = kq + 8 * e;
                if (vi < nv) {
                    const size_t o = (size_t)(b0 + vi) * T + t;
                    y[o * kH + col] = hh[e];
                    if (gates) {
                        float *gp = gates + o * 4 * kH + col;
                        gp[0] = rr[e]; gp[kH] = zz[e]; gp[2 * kH] = nn[e]; gp[3 * kH] = gg[e];
                    }
                }
            }
            cluster.barrier_wait();                                 // every CTA's slice is visible
            // ---- pull the 16 slices into the local full h (DSMEM), swizzled
            for (int i = tid; i < kV * (kH / 4); i += kGruThreads) {
                const int v = i / (kH / 4), q = i - v * (kH / 4);
                const int src = q >> 3, qq = q & 7;                 // owner CTA, float4 inside its 32 units
                const float *remote = cluster.map_shared_rank(s.hown, src);
                const float4 hv = *reinterpret_cast<const float4 *>(remote + ((t & 1) * kV + v) * kU + qq * 4);
                reinterpret_cast<float4 *>(s.hfull + (size_t)v * kH)[swz(q)] = hv;
            }
            __syncthreads();
        }
    }
    cluster.sync();        // nobody leaves while a neighbour may still read its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward: dgi, dgh (B,T,3H), dh0 (B,H) from dy (B,T,H) (+ dhT (B,H) or NULL)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGruThreads, 1)
gru_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ dhT, const float *__restrict__ w_hh,
               const float *__restrict__ y, const float *__restrict__ h0, const float *__restrict__ gates,
               float *__restrict__ dgi, float *__restrict__ dgh, float *__restrict__ dh0, int B, int T,
               int vpc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruSmemBwd &s = *reinterpret_cast<GruSmemBwd *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank();
    const int cid = blockIdx.x / kC, ncl = gridDim.x / kC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    load_w_slice(s.w, w_hh, rank, tid);

    for (int b0 = cid * vpc; b0 < B; b0 += ncl * vpc) {     // vpc <= kV voices per cluster pass
        const int nv = min(vpc, B - b0);
        __syncthreads();
        for (int i = tid; i < kV * kU; i += kGruThreads) {
            const int v = i / kU, u = i - v * kU;
            s.dhn[i] = (dhT && v < nv) ? __ldg(dhT + (size_t)(b0 + v) * kH + rank * kU + u) : 0.f;
        }
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
            // ---- gate gradients of the own units: thread = (voice, unit)
            for (int i = tid; i < kV * kU; i += kGruThreads) {
                const int v = i / kU, u = i - v * kU;
                const int col = rank * kU + u;
                float dar = 0.f, daz = 0.f, dan = 0.f, dghn = 0.f, direct = 0.f;
                if (v < nv) {
                    const size_t o = (size_t)(b0 + v) * T + t;
                    const float dh = __ldg(dy + o * kH + col) + s.dhn[i];
                    const float *gp = gates + o * 4 * kH + col;
                    const float r = __ldg(gp), z = __ldg(gp + kH), n = __ldg(gp + 2 * kH), ghn = __ldg(gp + 3 * kH);
                    const float hprev = t > 0 ? __ldg(y + (o - 1) * kH + col)
                                              : (h0 ? __ldg(h0 + (size_t)(b0 + v) * kH + col) : 0.f);
                    const float dn = dh * (1.f - z);
                    const float dz = dh * (hprev - n);
                    dan = dn * (1.f - n * n);
                    dar = dan * ghn * r * (1.f - r);
                    daz = dz * z * (1.f - z);
                    dghn = dan * r;
                    direct = dh * z;
                    float *a = dgi + o * 3 * kH + col;
                    a[0] = dar; a[kH] = daz; a[2 * kH] = dan;
                    float *c = dgh + o * 3 * kH + col;
                    c[0] = dar; c[kH] = daz; c[2 * kH] = dghn;
                }
                reinterpret_cast<float2 *>(s.down)[(0 * kU + u) * 12 + v] = make_float2(dar, dar);
                reinterpret_cast<float2 *>(s.down)[(1 * kU + u) * 12 + v] = make_float2(daz, daz);
                reinterpret_cast<float2 *>(s.down)[(2 * kU + u) * 12 + v] = make_float2(dghn, dghn);
                s.dhn[i] = direct;                       // the recurrent part is added after the reduce-scatter
            }
            __syncthreads();
            // ---- partial[v][k] = sum over own 96 rows of d[row][v] * W[row][k]; thread = (k-quad, row half)
            {
                const int half = lane >> 4;
                const int q = warp * 16 + (lane & 15);               // logical float4 index of k
                const int pq = swz(q);
                uint64_t acc2[kV][2];                                // (k, k+1) and (k+2, k+3) of the quad
#pragma unroll
                for (int v = 0; v < kV; ++v) acc2[v][0] = acc2[v][1] = 0ull;
                const int r0 = half * (kRows / 2);
#pragma unroll 2
                for (int rr = 0; rr < kRows / 2; ++rr) {
                    const int row = r0 + rr;
                    const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(s.w + (size_t)row * kH)[pq];
                    const ulonglong2 *dp = reinterpret_cast<const ulonglong2 *>(s.down + row * 24);
#pragma unroll
                    for (int vp = 0; vp < kV / 2; ++vp) {
                        const ulonglong2 d = dp[vp];                 // (d_v, d_v), (d_v+1, d_v+1)
                        acc2[2 * vp][0] = fma2(d.x, w.x, acc2[2 * vp][0]);
                        acc2[2 * vp][1] = fma2(d.x, w.y, acc2[2 * vp][1]);
                        acc2[2 * vp + 1][0] = fma2(d.y, w.x, acc2[2 * vp + 1][0]);
                        acc2[2 * vp + 1][1] = fma2(d.y, w.y, acc2[2 * vp + 1][1]);
                    }
                }
                float acc[kV][4];
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    unpk2(acc2[v][0], acc[v][0], acc[v][1]);
                    unpk2(acc2[v][1], acc[v][2], acc[v][3]);
                }
#pragma unroll
                for (int v = 0; v < kV; ++v)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[v][e] += __shfl_xor_sync(0xffffffffu, acc[v][e], 16);
                if (half == 0) {
#pragma unroll
                    for (int v = 0; v < kV; ++v)
                        reinterpret_cast<float4 *>(s.partial + (size_t)v * kH)[q] =
                            make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
                }
            }
            cluster.sync();                                          // all 16 partials are complete
            // ---- reduce-scatter: own 32 units of every voice summed over the 16 CTAs (fixed order)
            for (int i = tid; i < kV * (kU / 4); i += kGruThreads) {
                const int v = i / (kU / 4), uq = i - v * (kU / 4);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int src = 0; src < kC; ++src) {
                    const float *remote = cluster.map_shared_rank(s.partial, src);
                    const float4 p = *reinterpret_cast<const float4 *>(remote + (size_t)v * kH + rank * kU + uq * 4);
                    sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
                }
                float *d = s.dhn + v * kU + uq * 4;
                d[0] += sum.x; d[1] += sum.y; d[2] += sum.z; d[3] += sum.w;
            }
            cluster.sync();                                          // partial buffers may be overwritten again
        }
        for (int i = tid; i < kV * kU; i += kGruThreads) {
            const int v = i / kU, u = i - v * kU;
            if (v < nv) dh0[(size_t)(b0 + v) * kH + rank * kU + u] = s.dhn[i];
        }
    }
    cluster.sync();
}

template <typename K>
int launch_cluster(K kernel, size_t smem, int clusters, cudaStream_t st, void **args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kC);
    cfg.blockDim = dim3(kGruThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kC;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelExC(&cfg, (const void *)kernel, args);
    return e == cudaSuccess ? 0 : (int)e;
}

template <typename K>
int max_clusters(K kernel, size_t smem) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kC * 8);
    cfg.blockDim = dim3(kGruThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kC;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// clusters to launch and voices per cluster pass: small batches are spread over the resident clusters
// (fewer voices each, same number of time steps), large ones take kV voices per pass
void pick_clusters(int B, int resident, int *clusters, int *vpc) {
    *clusters = B < resident ? B : resident;
    int v = (B + *clusters - 1) / *clusters;
    *vpc = v < kV ? v : kV;
}

}  // namespace

// Number of 16-CTA clusters of the recurrence kernels that can be resident at once (0 = not supported here)
extern "C" int ddsp_b200_gru_resident_clusters(void) {
    static int cached = -1;
    if (cached < 0) {
        const int a = max_clusters(gru_fwd_kernel, sizeof(GruSmemFwd));
        const int b = max_clusters(gru_bwd_kernel, sizeof(GruSmemBwd));
        cached = a < b ? a : b;
    }
    return cached;
}

extern "C" int ddsp_b200_gru_fwd(const float *gi, const float *w_hh, const float *b_hh, const float *h0,
                                 float *y, float *gates, int B, int T, int H, void *stream) {
    DDSP_REQUIRE(gi && w_hh && b_hh && y && B > 0 && T > 0);
    if (H != kH) return DDSP_B200_EUNSUPPORTED;
    const int resident = ddsp_b200_gru_resident_clusters();
    if (resident < 1) return DDSP_B200_EUNSUPPORTED;
    int clusters, vpc;
    pick_clusters(B, resident, &clusters, &vpc);
    void *args[] = {(void *)&gi, (void *)&w_hh, (void *)&b_hh, (void *)&h0, (void *)&y, (void *)&gates, (void *)&B, (void *)&T,
                    (void *)&vpc};
    int s = launch_cluster(gru_fwd_kernel, sizeof(GruSmemFwd), clusters, (cudaStream_t)stream, args);
    return s ? s : ddsp_launch_status();
}

extern "C" int ddsp_b200_gru_bwd(const float *dy, const float *dhT, const float *w_hh, const float *y,
                                 const float *h0, const float *gates, float *dgi, float *dgh, float *dh0,
                                 int B, int T, int H, void *stream) {
    DDSP_REQUIRE(dy && w_hh && y && gates && dgi && dgh && dh0 && B > 0 && T > 0);
    if (H != kH) return DDSP_B200_EUNSUPPORTED;
    const int resident = ddsp_b200_gru_resident_clusters();
    if (resident < 1) return DDSP_B200_EUNSUPPORTED;
    int clusters, vpc;
    pick_clusters(B, resident, &clusters, &vpc);
    void *args[] = {(void *)&dy, (void *)&dhT, (void *)&w_hh, (void *)&y, (void *)&h0, (void *)&gates,
                    (void *)&dgi, (void *)&dgh, (void *)&dh0, (void *)&B, (void *)&T, (void *)&vpc};
    int s = launch_cluster(gru_bwd_kernel, sizeof(GruSmemBwd), clusters, (cudaStream_t)stream, args);
    return s ? s : ddsp_launch_status();
}
