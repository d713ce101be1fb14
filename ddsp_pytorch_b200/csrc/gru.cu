// Persistent GRU recurrence on thread-block clusters (control-net decoder).
//
// Reference path replaced: ddsp/core.py:132-133 `nn.GRU(...)` as decoder.py:40,59,65 calls it, i.e. the
// cuDNN RNN: on B200 it runs one 64x512 x 512x1536 SGEMM launch plus one element-wise launch PER TIME
// STEP (400 + 400 launches of ~19 us and ~3 us forward, the same again backward; measured 16 ms of a
// 23 ms training step at batch 64, profiles/r01_model_step_b64.json).
//
// Design.  The input projections gi = x W_ih^T + b_ih for all time steps stay one big GEMM (csrc/gemm3x.cu).
// The recurrence is ONE launch: a cluster of 16 CTAs (non-portable size; 7 such clusters are resident on a
// B200) keeps the whole 1536x512 fp32 W_hh on chip for all T steps -- 96 rows = 32 hidden units x 3 gates per
// CTA, part in shared memory, part in registers -- and owns up to 10 voices.
// Per step each CTA computes its 96 x V gate pre-activations with packed FP32 FMAs, applies the gates and
// PUSHES its 32 new hidden values per voice to all 16 CTAs with bulk DSMEM copies that signal the receivers'
// mbarriers (no cluster barrier on the critical path; h double buffered).
// Backward walks time in reverse with the same resident W_hh: per step each CTA turns dh of its units into the
// gate gradients, multiplies them with its 96 rows (partial W_hh^T d for all 512 inputs) and the 16 partials
// are reduce-scattered with the same push protocol.  dW_hh, dW_ih, dx are GEMMs over the stored gate
// gradients afterwards.
// The kernels are specialised on the voices a cluster holds (1, 2, 3, 5, 10): fewer voices leave registers for
// a larger share of W_hh.
// Measured (B200, T = 400, hidden 512): forward 4.6 us per step, backward 4.9 us per step at 64 voices,
// 1.9 us forward at one voice, against 11.8 / 15.8 us for cuDNN's per-step SGEMM + element-wise launches
// (profiles/r01_gru.json).
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

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// forward
//
// Thread = (pair of hidden units up = tid/16, k slice ks = tid%16): 6 rows of W_hh x all voices, k quads
// ks + 16 j.  The j = 0 quad of its rows lives in registers for the whole launch (that is what makes room for
// a double-buffered h in shared memory), the other seven stream from shared memory.  Accumulators are packed
// pairs (even k, odd k) so the inner loop is FFMA2 only.  The 16 k slices are combined by a recursive-halving
// shuffle reduction (60 values -> 4 per lane), gates are applied with thread = (voice, unit).
//
// Exchange of h_t: no cluster barrier.  Every CTA writes its 32 x V slice into a staging buffer and 16 threads
// push it with one bulk DSMEM copy each (cp.async.bulk.shared::cluster, 1280 B) into the SAME slot of every
// CTA's next-h buffer; each copy signals the destination's mbarrier (complete_tx), and a CTA starts step t+1
// as soon as its own barrier has seen all 16 slices.  h is double buffered: a slice of h_{t+1} can only be
// sent by a CTA that has received all of h_t, i.e. after every CTA finished reading h_{t-1} (same buffer).
// ---------------------------------------------------------------------------------------------
constexpr int kWq = kH / 4 - 16;                 // W quads per row kept in shared memory (112 of 128)
constexpr int kSliceBytes = kV * kU * 4;          // one CTA's slice of h for all voices (1280 B)

struct GruSmemFwd {
    float w[kRows * kWq * 4];                     // [g*32 + u][quads 16..127]
    float hbuf[2][kC][kV][kU];                    // h, slice-major: [buffer][owner CTA][voice][unit]
    float stage[2][kV][kU];                       // own slice of the step being published
    float gh[3][kV][kU];                          // reduced W_hh h of the own units
    unsigned long long bar[2];                    // one mbarrier per h buffer
};

__device__ __forceinline__ uint32_t gru_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gru_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void gru_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();               // a lost slice traps instead of hanging the GPU
    } while (!ok);
}
// push `bytes` of own shared memory into CTA `dst_rank`'s shared memory at the same offsets, signalling its barrier
__device__ __forceinline__ void gru_push(uint32_t dst_local, uint32_t src, uint32_t bar_local, uint32_t dst_rank,
                                         uint32_t bytes) {
    const uint32_t dst = gru_mapa(dst_local, dst_rank), bar = gru_mapa(bar_local, dst_rank);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// NV = voices the matrix product is computed for (>= vpc); shared-memory layouts keep kV slots
template <int NV>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_fwd_kernel(const float *__restrict__ gi, const float *__restrict__ w_hh, const float *__restrict__ b_hh,
               const float *__restrict__ h0, float *__restrict__ y, float *__restrict__ gates, int B, int T,
               int vpc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruSmemFwd &s = *reinterpret_cast<GruSmemFwd *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank();
    const int cid = blockIdx.x / kC, ncl = gridDim.x / kC;
    const int tid = threadIdx.x;
    const int up = tid >> 4, ks = tid & 15;                        // unit pair, k slice

    // ---- W_hh slice: quad ks of the six own rows into registers, quads 16.. into shared memory
    // few voices leave registers free: then four of the eight k quads of W_hh stay in registers instead of one
    // and the per-step stream from shared memory (172 KB, 0.7 us) halves
    constexpr int JR = NV <= 3 ? 4 : (NV <= 5 ? 2 : 1);
    ulonglong2 wreg[JR][6];
#pragma unroll
    for (int jr = 0; jr < JR; ++jr)
#pragma unroll
        for (int r6 = 0; r6 < 6; ++r6) {
            const int g = r6 >> 1, u = 2 * up + (r6 & 1);
            wreg[jr][r6] = __ldg(reinterpret_cast<const ulonglong2 *>(w_hh + ((size_t)g * kH + rank * kU + u) * kH) + ks + 16 * jr);
        }
    for (int i = tid; i < kRows * kWq; i += kGruThreads) {
        const int row = i / kWq, q = i - row * kWq;
        const int g = row / kU, u = row - g * kU;
        reinterpret_cast<float4 *>(s.w)[i] =
            __ldg(reinterpret_cast<const float4 *>(w_hh + ((size_t)g * kH + rank * kU + u) * kH) + 16 + q);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gru_smem_u32(&s.bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gru_smem_u32(&s.bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // gate-phase items of this thread: (voice, unit) = (tid/32, tid%32) and, for tid < 64, (8 + tid/32, tid%32)
    const int gu = tid & 31, gcol = rank * kU + gu;
    const float br = __ldg(b_hh + gcol), bz = __ldg(b_hh + kH + gcol), bn = __ldg(b_hh + 2 * kH + gcol);

    int it = 0;                                                     // steps done by this cluster (buffer / parity clock)
    for (int b0 = cid * vpc; b0 < B; b0 += ncl * vpc) {             // vpc <= kV voices per cluster pass
        const int nv = min(vpc, B - b0);
        cluster.sync();                                             // previous pass fully drained everywhere
        {   // h_{-1} = h0 into the buffer step `it` reads; staging zeroed (unused voices stay zero)
            float *hb = &s.hbuf[(it + 1) & 1][0][0][0];
            for (int i = tid; i < kC * kV * kU; i += kGruThreads) {
                const int u = i % kU, v = (i / kU) % kV, r = i / (kU * kV);
                hb[i] = (h0 && v < nv) ? __ldg(h0 + (size_t)(b0 + v) * kH + r * kU + u) : 0.f;
            }
            for (int i = tid; i < 2 * kV * kU; i += kGruThreads) (&s.stage[0][0][0])[i] = 0.f;
        }
        __syncthreads();
        for (int t = 0; t < T; ++t, ++it) {
            const int rb = (it + 1) & 1, wb = it & 1;               // h_{t-1} buffer, h_t buffer
            // prefetch the input projections of this thread's gate items
            float gir[2], giz[2], gin[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = (tid >> 5) + 8 * e;
                gir[e] = giz[e] = gin[e] = 0.f;
                if ((e == 0 || tid < 64) && v < nv) {
                    const float *p = gi + ((size_t)(b0 + v) * T + t) * 3 * kH + gcol;
                    gir[e] = __ldg(p);
                    giz[e] = __ldg(p + kH);
                    gin[e] = __ldg(p + 2 * kH);
                }
            }
            // ---- partial W_hh h over this thread's k slice
            uint64_t acc2[6][NV];
            const ulonglong2 *hq = reinterpret_cast<const ulonglong2 *>(&s.hbuf[rb][0][0][0]);
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int r6 = 0; r6 < 6; ++r6) acc2[r6][v] = 0ull;
#pragma unroll
            for (int jr = 0; jr < JR; ++jr) {
                const int q = ks + 16 * jr;                         // quad q: owner q/8, quad q%8 inside its 32 units
                const int hoff = (q >> 3) * kV * 8 + (q & 7);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const ulonglong2 h = hq[hoff + v * 8];
#pragma unroll
                    for (int r6 = 0; r6 < 6; ++r6)
                        acc2[r6][v] = fma2(wreg[jr][r6].y, h.y, fma2(wreg[jr][r6].x, h.x, acc2[r6][v]));
                }
            }
            const ulonglong2 *wq = reinterpret_cast<const ulonglong2 *>(s.w);
#pragma unroll
            for (int j = JR; j < 8; ++j) {
                const int q = ks + 16 * j;
                ulonglong2 w[6];
#pragma unroll
                for (int r6 = 0; r6 < 6; ++r6) w[r6] = wq[((r6 >> 1) * kU + 2 * up + (r6 & 1)) * kWq + (q - 16)];
                const int hoff = (q >> 3) * kV * 8 + (q & 7);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const ulonglong2 h = hq[hoff + v * 8];
#pragma unroll
                    for (int r6 = 0; r6 < 6; ++r6) acc2[r6][v] = fma2(w[r6].y, h.y, fma2(w[r6].x, h.x, acc2[r6][v]));
                }
            }
            // ---- combine the 16 k slices: recursive halving over P >= 6 NV entries, lane ks ends with entries
            //      (P/16) ks .. (P/16) ks + P/16 - 1
            constexpr int P = 6 * NV > 32 ? 64 : (6 * NV > 16 ? 32 : 16);
            float a[P];
#pragma unroll
            for (int i = 0; i < P; ++i) a[i] = 0.f;
#pragma unroll
            for (int r6 = 0; r6 < 6; ++r6)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float lo, hi;
                    unpk2(acc2[r6][v], lo, hi);
                    a[r6 * NV + v] = lo + hi;
                }
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                const int n = (P / 2) >> st, m = 8 >> st;
                const bool upper = (ks & m) != 0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float keep = upper ? a[i + n] : a[i];
                    const float send = upper ? a[i] : a[i + n];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            }
#pragma unroll
            for (int i = 0; i < P / 16; ++i) {
                const int e = (P / 16) * ks + i;                    // = r6 * NV + v
                if (e < 6 * NV) {
                    const int r6 = e / NV, v = e - r6 * NV;
                    s.gh[r6 >> 1][v][2 * up + (r6 & 1)] = a[i];
                }
            }
            __syncthreads();
            // ---- gates: thread = (voice, unit)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = (tid >> 5) + 8 * e;
                if ((e == 0 || tid < 64) && v < nv) {
                    const float ghr = s.gh[0][v][gu] + br, ghz = s.gh[1][v][gu] + bz, ghn = s.gh[2][v][gu] + bn;
                    const float hprev = s.hbuf[rb][rank][v][gu];
                    const float r = sigmoid_acc(gir[e] + ghr);
                    const float z = sigmoid_acc(giz[e] + ghz);
                    const float n = tanhf(fmaf(r, ghn, gin[e]));
                    const float hnew = fmaf(z, hprev - n, n);        // (1-z) n + z h
                    s.stage[wb][v][gu] = hnew;
                    const size_t o = (size_t)(b0 + v) * T + t;
                    y[o * kH + gcol] = hnew;
                    if (gates) {
                        float *gp = gates + o * 4 * kH + gcol;
                        gp[0] = r; gp[kH] = z; gp[2 * kH] = n; gp[3 * kH] = ghn;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging visible to the bulk-copy engine
            __syncthreads();
            // ---- publish: 16 bulk copies of the own slice, one per destination CTA; arm the own barrier
            const uint32_t bar = gru_smem_u32(&s.bar[wb]);
            if (tid < kC)
                gru_push(gru_smem_u32(&s.hbuf[wb][rank][0][0]), gru_smem_u32(&s.stage[wb][0][0]), bar, tid, kSliceBytes);
            if (tid == 32)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kC * kSliceBytes)
                             : "memory");
            gru_bar_wait(bar, (it >> 1) & 1);                        // all 16 slices of h_t have landed here
        }
    }
    cluster.sync();        // nobody leaves while a neighbour may still write into its shared memory
}

// ---------------------------------------------------------------------------------------------
// backward: dgi, dgh (B,T,3H), dh0 (B,H) from dy (B,T,H) (+ dhT (B,H) or NULL)
//
// Per step (time reversed): gate gradients of the own 32 units (thread = (voice, unit); their global inputs
// are prefetched one step ahead), then partial[v][k] = sum over the own 96 rows of d[row][v] W_hh[row][k]
// (thread = (k quad, half of the rows); 16 of its 48 rows of W_hh live in registers, 32 in shared memory),
// stored slice-major by destination CTA.  Reduce-scatter without a cluster barrier: 16 bulk DSMEM copies push
// slice d to CTA d's receive buffer and signal its mbarrier; each CTA then adds the 16 slices it received.
// Receive and send buffers are double buffered (same argument as in the forward).
// ---------------------------------------------------------------------------------------------
constexpr int kRegRows = 16;                      // rows per thread (of its 48) whose W_hh quad stays in registers
constexpr int kSmemRows = 2 * (kRows / 2 - kRegRows);

struct GruSmemBwd {
    float w[kSmemRows * kH];                      // [half*32 + (rr - 16)][512]
    float partial[2][kC][kV][kU];                 // [buffer][destination CTA][voice][unit]
    float recv[2][kC][kV][kU];                    // [buffer][source CTA][voice][unit]
    float down[kRows * 12];                       // [row][12 (>= kV)] gate gradients of the own rows
    float dhn[kV * kU];                           // recurrent part of dh for the own units
    unsigned long long bar[2];
};

// the gate gradients of one row for the voices the product covers: 1..3 LDS.128 (warp-uniform address)
template <int NV>
__device__ __forceinline__ void load_down(const float *p, float (&dv)[12]) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    dv[0] = a.x; dv[1] = a.y; dv[2] = a.z; dv[3] = a.w;
    if (NV > 4) {
        const float4 b = *reinterpret_cast<const float4 *>(p + 4);
        dv[4] = b.x; dv[5] = b.y; dv[6] = b.z; dv[7] = b.w;
    }
    if (NV > 8) {
        const float4 c = *reinterpret_cast<const float4 *>(p + 8);
        dv[8] = c.x; dv[9] = c.y; dv[10] = c.z; dv[11] = c.w;
    }
}

template <int NV>
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
    const int half = lane >> 4, q = warp * 16 + (lane & 15);        // row half, float4 index of k
    const int r0 = half * (kRows / 2);

    // ---- W_hh slice (slice row = g*32 + u  <->  global row g*H + 32*rank + u).  KR of a thread's 48 rows stay
    //      in registers: 16 with many voices, 40 when few voices leave the registers free (the shared-memory
    //      buffer is sized for the first case)
    constexpr int KR = NV <= 3 ? 40 : (NV <= 5 ? 32 : kRegRows);
    constexpr int SR = kRows / 2 - KR;                               // rows per half streamed from shared memory
    ulonglong2 wreg[KR];
#pragma unroll
    for (int rr = 0; rr < KR; ++rr) {
        const int row = r0 + rr, g = row / kU, u = row - g * kU;
        wreg[rr] = __ldg(reinterpret_cast<const ulonglong2 *>(w_hh + ((size_t)g * kH + rank * kU + u) * kH) + q);
    }
    for (int i = tid; i < 2 * SR * (kH / 4); i += kGruThreads) {
        const int srow = i / (kH / 4), qq = i - srow * (kH / 4);
        const int row = (srow / SR) * (kRows / 2) + KR + (srow % SR);
        const int g = row / kU, u = row - g * kU;
        reinterpret_cast<float4 *>(s.w)[i] =
            __ldg(reinterpret_cast<const float4 *>(w_hh + ((size_t)g * kH + rank * kU + u) * kH) + qq);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gru_smem_u32(&s.bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gru_smem_u32(&s.bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int gu = tid & 31, gcol = rank * kU + gu;                  // gate items: (tid/32, gu) and, tid < 64, (8 + tid/32, gu)

    int it = 0;
    for (int b0 = cid * vpc; b0 < B; b0 += ncl * vpc) {             // vpc <= kV voices per cluster pass
        const int nv = min(vpc, B - b0);
        cluster.sync();
        for (int i = tid; i < kV * kU; i += kGruThreads) {
            const int v = i / kU, u = i - v * kU;
            s.dhn[i] = (dhT && v < nv) ? __ldg(dhT + (size_t)(b0 + v) * kH + rank * kU + u) : 0.f;
        }
        for (int i = tid; i < 2 * kC * kV * kU; i += kGruThreads) (&s.partial[0][0][0][0])[i] = 0.f;   // slots >= NV stay zero
        // inputs of the gate phase of step T-1
        float p_dy[2], p_r[2], p_z[2], p_n[2], p_g[2], p_h[2];
        auto prefetch = [&](int t) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = (tid >> 5) + 8 * e;
                p_dy[e] = p_r[e] = p_z[e] = p_n[e] = p_g[e] = p_h[e] = 0.f;
                if ((e == 0 || tid < 64) && v < nv) {
                    const size_t o = (size_t)(b0 + v) * T + t;
                    p_dy[e] = __ldg(dy + o * kH + gcol);
                    const float *gp = gates + o * 4 * kH + gcol;
                    p_r[e] = __ldg(gp); p_z[e] = __ldg(gp + kH); p_n[e] = __ldg(gp + 2 * kH); p_g[e] = __ldg(gp + 3 * kH);
                    p_h[e] = t > 0 ? __ldg(y + (o - 1) * kH + gcol) : (h0 ? __ldg(h0 + (size_t)(b0 + v) * kH + gcol) : 0.f);
                }
            }
        };
        prefetch(T - 1);
        __syncthreads();
        for (int t = T - 1; t >= 0; --t, ++it) {
            const int pb = it & 1;
            // ---- gate gradients of the own units: thread = (voice, unit)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int v = (tid >> 5) + 8 * e;
                if (e == 0 || tid < 64) {
                    const int i = v * kU + gu;
                    float dar = 0.f, daz = 0.f, dghn = 0.f, direct = 0.f;
                    if (v < nv) {
                        const float r = p_r[e], z = p_z[e], n = p_n[e];
                        const float dh = p_dy[e] + s.dhn[i];
                        const float dn = dh * (1.f - z);
                        const float dz = dh * (p_h[e] - n);
                        const float dan = dn * (1.f - n * n);
                        dar = dan * p_g[e] * r * (1.f - r);
                        daz = dz * z * (1.f - z);
                        dghn = dan * r;
                        direct = dh * z;
                        const size_t o = (size_t)(b0 + v) * T + t;
                        float *a = dgi + o * 3 * kH + gcol;
                        a[0] = dar; a[kH] = daz; a[2 * kH] = dan;
                        float *c = dgh + o * 3 * kH + gcol;
                        c[0] = dar; c[kH] = daz; c[2 * kH] = dghn;
                    }
                    s.down[(0 * kU + gu) * 12 + v] = dar;
                    s.down[(1 * kU + gu) * 12 + v] = daz;
                    s.down[(2 * kU + gu) * 12 + v] = dghn;
                    s.dhn[i] = direct;                   // the recurrent part is added after the reduce-scatter
                }
            }
            __syncthreads();
            if (t > 0) prefetch(t - 1);                  // rides under the matrix product below
            // ---- partial[v][k] over the own row half
            {
                constexpr int NVP = (NV + 1) / 2;                    // voice pairs the product is computed for
                uint64_t acc2[2 * NVP][2];                           // (k, k+1) and (k+2, k+3) of the quad
#pragma unroll
                for (int v = 0; v < 2 * NVP; ++v) acc2[v][0] = acc2[v][1] = 0ull;
#pragma unroll
                for (int rr = 0; rr < KR; ++rr) {
                    float dv[12];
                    load_down<NV>(s.down + (r0 + rr) * 12, dv);
#pragma unroll
                    for (int v = 0; v < 2 * NVP; ++v) {
                        const uint64_t d = pk2(dv[v], dv[v]);        // duplicated in registers (ALU pipe), not in shared memory
                        acc2[v][0] = fma2(d, wreg[rr].x, acc2[v][0]);
                        acc2[v][1] = fma2(d, wreg[rr].y, acc2[v][1]);
                    }
                }
                const ulonglong2 *wq = reinterpret_cast<const ulonglong2 *>(s.w) + (size_t)half * SR * (kH / 4) + q;
#pragma unroll 8
                for (int rr = KR; rr < kRows / 2; ++rr) {
                    const ulonglong2 w = wq[(size_t)(rr - KR) * (kH / 4)];
                    float dv[12];
                    load_down<NV>(s.down + (r0 + rr) * 12, dv);
#pragma unroll
                    for (int v = 0; v < 2 * NVP; ++v) {
                        const uint64_t d = pk2(dv[v], dv[v]);
                        acc2[v][0] = fma2(d, w.x, acc2[v][0]);
                        acc2[v][1] = fma2(d, w.y, acc2[v][1]);
                    }
                }
                float acc[2 * NVP][4];
#pragma unroll
                for (int v = 0; v < 2 * NVP; ++v) {
                    unpk2(acc2[v][0], acc[v][0], acc[v][1]);
                    unpk2(acc2[v][1], acc[v][2], acc[v][3]);
                }
#pragma unroll
                for (int v = 0; v < 2 * NVP; ++v)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[v][e] += __shfl_xor_sync(0xffffffffu, acc[v][e], 16);
                if (half == 0) {
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        *reinterpret_cast<float4 *>(&s.partial[pb][q >> 3][v][(q & 7) * 4]) =
                            make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            // ---- reduce-scatter: slice d goes to CTA d's receive slot `rank`
            const uint32_t bar = gru_smem_u32(&s.bar[pb]);
            if (tid < kC)
                gru_push(gru_smem_u32(&s.recv[pb][rank][0][0]), gru_smem_u32(&s.partial[pb][tid][0][0]), bar, tid, kSliceBytes);
            if (tid == 32)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kC * kSliceBytes)
                             : "memory");
            gru_bar_wait(bar, (it >> 1) & 1);
            for (int i = tid; i < kV * (kU / 4); i += kGruThreads) {
                const int v = i / (kU / 4), uq = i - v * (kU / 4);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int src = 0; src < kC; ++src) {                 // fixed order: deterministic
                    const float4 pv = *reinterpret_cast<const float4 *>(&s.recv[pb][src][v][uq * 4]);
                    sum.x += pv.x; sum.y += pv.y; sum.z += pv.z; sum.w += pv.w;
                }
                float *d = s.dhn + v * kU + uq * 4;
                d[0] += sum.x; d[1] += sum.y; d[2] += sum.z; d[3] += sum.w;
            }
            __syncthreads();
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
        const int a = max_clusters(gru_fwd_kernel<kV>, sizeof(GruSmemFwd));
        const int b = max_clusters(gru_bwd_kernel<kV>, sizeof(GruSmemBwd));
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
    // the kernel variant computes the product for the smallest supported voice count >= vpc
    cudaStream_t st = (cudaStream_t)stream;
    int s;
    if (vpc <= 1) s = launch_cluster(gru_fwd_kernel<1>, sizeof(GruSmemFwd), clusters, st, args);
    else if (vpc <= 2) s = launch_cluster(gru_fwd_kernel<2>, sizeof(GruSmemFwd), clusters, st, args);
    else if (vpc <= 3) s = launch_cluster(gru_fwd_kernel<3>, sizeof(GruSmemFwd), clusters, st, args);
    else if (vpc <= 5) s = launch_cluster(gru_fwd_kernel<5>, sizeof(GruSmemFwd), clusters, st, args);
    else s = launch_cluster(gru_fwd_kernel<kV>, sizeof(GruSmemFwd), clusters, st, args);
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
    cudaStream_t st = (cudaStream_t)stream;
    int s;
    if (vpc <= 1) s = launch_cluster(gru_bwd_kernel<1>, sizeof(GruSmemBwd), clusters, st, args);
    else if (vpc <= 2) s = launch_cluster(gru_bwd_kernel<2>, sizeof(GruSmemBwd), clusters, st, args);
    else if (vpc <= 3) s = launch_cluster(gru_bwd_kernel<3>, sizeof(GruSmemBwd), clusters, st, args);
    else if (vpc <= 5) s = launch_cluster(gru_bwd_kernel<5>, sizeof(GruSmemBwd), clusters, st, args);
    else s = launch_cluster(gru_bwd_kernel<kV>, sizeof(GruSmemBwd), clusters, st, args);
    return s ? s : ddsp_launch_status();
}
