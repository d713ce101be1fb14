// FP32-accurate GEMM on the 5th-generation tensor cores (the control net's nn.Linear /
// GRU input projections, ddsp/core.py:122-133, decoder.py:40-68,86-87).
//
// Reference path replaced: cuBLAS SIMT SGEMM (torch's default for float32 nn.Linear with TF32 off, the
// reference's setting): ~50 TFLOP/s on a B200, half of a training step at batch 64
// (profiles/r01_model_step_profile.txt).
//
// Method (bf16 x 3 split, 6 products): every fp32 operand x is split once into three bf16 parts
// x = b0 + b1 + b2 (b0 = bf16(x), b1 = bf16(x - b0), b2 = bf16(x - b0 - b1): 3 x 8 mantissa bits, exact), and
// C = A0 B0 + (A0 B1 + A1 B0) + (A1 B1 + A0 B2 + A2 B0) is accumulated in fp32 in tensor memory; the dropped
// products are below 2^-24 relative.  Six bf16 MMAs cost half of three TF32 MMAs on this tensor core (the
// first version of this kernel, "3xTF32", measured 128 cycles per M128 N128 K8 TF32 instruction).
//
// Kernel: persistent, one CTA per SM walking 128 x 128 output tiles (x K splits), operands K-major.
// Warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, 2-stage mbarrier ring of 96 KB stages: three
// parts of A and of B, 64 values of K), warp 1 = tcgen05.mma issuer (one thread, kind::f16, M128 N128 K16,
// 24 MMAs per stage, two ping-pong pairs of 128-column TMEM accumulators), warps 2-5 = epilogue (tcgen05.ld
// 32x32b of every finished 64-deep K chunk, accumulated in registers; + bias and global stores at tile end).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;          // CTA tile; kBK bf16 = 128 B = one swizzle row
constexpr int kStages = 2;
constexpr int kChunk = 1;                                // k blocks accumulated in TMEM before promotion to registers
constexpr int kTileBytes = kBM * kBK * 2;              // 16 KB per operand part
constexpr int kStageBytes = 6 * kTileBytes;            // A0, A1, A2, B0, B1, B2
constexpr int kGemmThreads = 192;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr uint32_t kSpinLimit = 1u << 26;              // a lost barrier traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > kSpinLimit) __trap();
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// MN-major operand tile (the matrix is stored K rows x MN columns), 128-byte swizzle: one 128 B row = 64
// consecutive MN elements of one k; 8 consecutive k = 1024 B (SBO); the next 64 MN elements kBK rows later (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((kBK * 128) >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// descriptors are given as (low word, high word): the low word carries the 16-byte start address, so stepping
// through a stage is one 32-bit add per operand and the issuing thread stays far below the 64-cycle MMA pace
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

struct GemmParams {
    float *c;                  // output, or split-K partials [splits][M][N] (then ldc == N)
    const float *bias;         // added in the epilogue when splits == 1
    int M, N, ldc;
    int a_lo_row, b_lo_row;    // row distance between the three parts inside the operand's tensor map
    int a_mn, b_mn;            // operand stored K rows x MN columns (MN-major) instead of MN rows x K columns
    int k_blocks, k_blocks_per_split, splits;
    int m_tiles, n_tiles;
};

// tile t -> (split z, m tile, n tile); consecutive tiles share the A rows (L2 reuse)
__device__ __forceinline__ void tile_coords(const GemmParams &p, int t, int &z, int &m0, int &n0, int &kb0, int &nkb) {
    const int per_split = p.m_tiles * p.n_tiles;
    z = t / per_split;
    const int r = t - z * per_split;
    m0 = (r / p.n_tiles) * kBM;
    n0 = (r % p.n_tiles) * kBN;
    kb0 = z * p.k_blocks_per_split;
    nkb = min(p.k_blocks, kb0 + p.k_blocks_per_split) - kb0;
}

// Persistent: CTA b works on tiles b, b + grid, ...  The tensor core's fp32 accumulation truncates, so the
// error of a long K loop grows linearly with K; therefore K is accumulated in TMEM only over chunks of
// kChunk k blocks (64 values of K: 4 accumulations of the main term) and the epilogue warps add every
// finished chunk into fp32 REGISTERS
// (round-to-nearest) while the next chunk is being multiplied into the other TMEM buffer.  The two small
// cross terms go to their own accumulator so that their rounding happens at their own (2^-11) scale.
// TMEM columns: [main0 | small0 | main1 | small1], 128 each.
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kStages * kStageBytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStages;
    const uint32_t tfull0 = bars + 16 * kStages, tempty0 = tfull0 + 16;     // 2 TMEM buffers each
    const uint32_t tmem_slot = tempty0 + 16;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.m_tiles * p.n_tiles * p.splits;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);
            mbar_init(tempty0 + 8 * b, 4);            // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;                                                      // k blocks issued so far (ring position)
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int z, m0, n0, kb0, nkb;
                tile_coords(p, t, z, m0, n0, kb0, nkb);
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % kStages, round = it / kStages;
                    mbar_wait(empty0 + 8 * s, (round & 1) ^ 1);
                    const uint32_t st = base + s * kStageBytes, fb = full0 + 8 * s;
                    mbar_expect_tx(fb, kStageBytes);
                    const int k = (kb0 + i) * kBK;
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const uint32_t ta = st + part * kTileBytes, tb = st + (3 + part) * kTileBytes;
                        if (!p.a_mn) {
                            tma_load_2d(ta, &map_a, fb, k, part * p.a_lo_row + m0);
                        } else {                                   // two boxes of 64 MN x 64 k
                            tma_load_2d(ta, &map_a, fb, m0, part * p.a_lo_row + k);
                            tma_load_2d(ta + kTileBytes / 2, &map_a, fb, m0 + 64, part * p.a_lo_row + k);
                        }
                        if (!p.b_mn) {
                            tma_load_2d(tb, &map_b, fb, k, part * p.b_lo_row + n0);
                        } else {
                            tma_load_2d(tb, &map_b, fb, n0, part * p.b_lo_row + k);
                            tma_load_2d(tb + kTileBytes / 2, &map_b, fb, n0 + 64, part * p.b_lo_row + k);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                                   ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            // descriptor words of the six operand parts at stage 0, k step 0; a k step (16 values of K) is 32 B
            // inside the 128 B row (K-major) or 16 rows of 128 B (MN-major) further, a stage kStageBytes
            uint32_t a_lo[3], b_lo[3], a_hi = 0, b_hi = 0;
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const uint32_t ta = base + part * kTileBytes, tb = base + (3 + part) * kTileBytes;
                const uint64_t da = p.a_mn ? umma_desc_mn_sw128(ta) : umma_desc_k_sw128(ta);
                const uint64_t db = p.b_mn ? umma_desc_mn_sw128(tb) : umma_desc_k_sw128(tb);
                a_lo[part] = (uint32_t)da; a_hi = (uint32_t)(da >> 32);
                b_lo[part] = (uint32_t)db; b_hi = (uint32_t)(db >> 32);
            }
            const uint32_t a_inc = p.a_mn ? (2048u >> 4) : (32u >> 4), b_inc = p.b_mn ? (2048u >> 4) : (32u >> 4);
            int it = 0, chunk = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int z, m0, n0, kb0, nkb;
                tile_coords(p, t, z, m0, n0, kb0, nkb);
                for (int i0 = 0; i0 < nkb; i0 += kChunk, ++chunk) {
                    const int buf = chunk & 1, use = chunk >> 1;
                    mbar_wait(tempty0 + 8 * buf, (use & 1) ^ 1);             // epilogue has drained this buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_main = tmem_base + buf * 2 * kBN, d_small = d_main + kBN;
                    const int i1 = min(nkb, i0 + kChunk);
                    for (int i = i0; i < i1; ++i, ++it) {
                        const int s = it % kStages, round = it / kStages;
                        mbar_wait(full0 + 8 * s, round & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        uint32_t ao = s * (kStageBytes >> 4), bo = ao;
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k, ao += a_inc, bo += b_inc) {
                            const uint32_t a0 = a_lo[0] + ao, a1 = a_lo[1] + ao, a2 = a_lo[2] + ao;
                            const uint32_t b0 = b_lo[0] + bo, b1 = b_lo[1] + bo, b2 = b_lo[2] + bo;
                            const uint32_t acc = (i != i0 || k != 0) ? 1u : 0u;
                            umma_bf16_w(d_main, a0, a_hi, b0, b_hi, idesc, acc);
                            umma_bf16_w(d_small, a1, a_hi, b1, b_hi, idesc, acc);   // 2^-16 terms first
                            umma_bf16_w(d_small, a0, a_hi, b2, b_hi, idesc, 1u);
                            umma_bf16_w(d_small, a2, a_hi, b0, b_hi, idesc, 1u);
                            umma_bf16_w(d_small, a0, a_hi, b1, b_hi, idesc, 1u);    // 2^-8 terms
                            umma_bf16_w(d_small, a1, a_hi, b0, b_hi, idesc, 1u);
                        }
                        umma_commit(empty0 + 8 * s);                               // stage free when these MMAs retire
                    }
                    umma_commit(tfull0 + 8 * buf);                                 // chunk complete in TMEM
                }
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (its hardware quarter) = 32 output rows =====
        const int quarter = warp & 3;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
        int chunk = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            int z, m0, n0, kb0, nkb;
            tile_coords(p, t, z, m0, n0, kb0, nkb);
            float acc[kBN];
#pragma unroll
            for (int j = 0; j < kBN; ++j) acc[j] = 0.f;
            for (int i0 = 0; i0 < nkb; i0 += kChunk, ++chunk) {
                const int buf = chunk & 1, use = chunk >> 1;
                mbar_wait(tfull0 + 8 * buf, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_main = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 2 * kBN;
#pragma unroll
                for (int cc = 0; cc < kBN / 32; ++cc) {
                    uint32_t u[32];
                    tmem_ld32(t_main + kBN + cc * 32, u);                     // cross terms first (small)
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(u[j]);
                    tmem_ld32(t_main + cc * 32, u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(u[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            }
            const int row = m0 + quarter * 32 + lane;
            if (row < p.M) {
                float *crow = p.c + (size_t)z * p.M * p.ldc + (size_t)row * p.ldc;
#pragma unroll
                for (int cc = 0; cc < kBN / 32; ++cc) {
                    const int col0 = n0 + cc * 32;
                    if (col0 >= p.N) break;
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) acc[cc * 32 + j] += __ldg(p.bias + col0 + j);
                    }
                    if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(crow + col0 + j) =
                                make_float4(acc[cc * 32 + j], acc[cc * 32 + j + 1], acc[cc * 32 + j + 2], acc[cc * 32 + j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) crow[col0 + j] = acc[cc * 32 + j];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- split-K reduction: c[m][n] = bias[n] + sum_s partial[s][m][n]  (fixed order: deterministic)
__global__ void gemm3x_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ bias,
                                     float *__restrict__ c, int M, int N, int ldc, int splits) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float acc = bias ? __ldg(bias + n) : 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldg(partial + (size_t)s * M * N + i);
    c[(size_t)m * ldc + n] = acc;
}

// ---- operand split: x = b0 + b1 + b2 in bf16; part p of row r at out[(p * part_rows + r)][c]; columns
// cols..out_ld zeroed.  transpose == 0: logical operand = x (rows x cols, row pitch ld); 1: operand = x^T.
__device__ __forceinline__ void split_bf16x3(float x, __nv_bfloat16 &b0, __nv_bfloat16 &b1, __nv_bfloat16 &b2) {
    b0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b0);          // exact
    b1 = __float2bfloat16_rn(r1);
    b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));  // exact residual, <= 8 significant bits
}
__device__ __forceinline__ void store_parts(__nv_bfloat16 *out, size_t idx, size_t part_stride, float x) {
    __nv_bfloat16 b0, b1, b2;
    split_bf16x3(x, b0, b1, b2);
    out[idx] = b0;
    out[idx + part_stride] = b1;
    out[idx + 2 * part_stride] = b2;
}

// Non-transposed split, vectorised: a block owns 64 consecutive operand rows; thread = (column group of 8, row
// lane); 2 x 16 B loads and 3 x 16 B stores per item.  COLSUM additionally produces the column sums of x (the
// bias gradient of a Linear layer whose dy is being split) as per-block partials, reduced by a second launch in
// a fixed order (deterministic).
constexpr int kSplitRows = 64;

template <bool COLSUM>
__global__ void __launch_bounds__(256)
split3x_kernel(const float *__restrict__ x, int rows, int cols, int64_t ld, __nv_bfloat16 *__restrict__ out,
               int64_t part_rows, int out_ld, float *__restrict__ partial, int cgroups, int lanes, int rpb) {
    extern __shared__ float red[];                                  // [lanes][cgroups * 8] when COLSUM
    const int cl = threadIdx.x % cgroups, rl = threadIdx.x / cgroups;
    const int c0 = (blockIdx.y * cgroups + cl) * 8;                 // grid.y > 1 only beyond 2048 columns
    const bool live = c0 < out_ld;
    const int rbeg = blockIdx.x * rpb;                              // rpb rows per block (<= kSplitRows)
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && c0 + 8 <= cols;
    const size_t pstride = (size_t)part_rows * out_ld;
    float sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0.f;
    for (int r = rbeg + rl; live && r < rbeg + rpb && r < part_rows; r += lanes) {
        float v[8];
        if (r < rows && vec) {
            const float4 lo = __ldg(reinterpret_cast<const float4 *>(x + (size_t)r * ld + c0));
            const float4 hi = __ldg(reinterpret_cast<const float4 *>(x + (size_t)r * ld + c0 + 4));
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (r < rows && c0 + j < cols) ? __ldg(x + (size_t)r * ld + c0 + j) : 0.f;
        }
        __align__(16) __nv_bfloat16 b0[8], b1[8], b2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            split_bf16x3(v[j], b0[j], b1[j], b2[j]);
            if (COLSUM) sum[j] += v[j];
        }
        const size_t o = (size_t)r * out_ld + c0;
        *reinterpret_cast<uint4 *>(out + o) = *reinterpret_cast<const uint4 *>(b0);
        *reinterpret_cast<uint4 *>(out + o + pstride) = *reinterpret_cast<const uint4 *>(b1);
        *reinterpret_cast<uint4 *>(out + o + 2 * pstride) = *reinterpret_cast<const uint4 *>(b2);
    }
    if (COLSUM) {
        const int w = cgroups * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) red[rl * w + cl * 8 + j] = sum[j];
        __syncthreads();
        for (int c = threadIdx.x; c < w; c += blockDim.x) {
            const int cglob = blockIdx.y * w + c;
            if (cglob < out_ld) {
                float t = 0.f;
                for (int l = 0; l < lanes; ++l) t += red[l * w + c];
                partial[(size_t)blockIdx.x * out_ld + cglob] = t;
            }
        }
    }
}

// 8 lanes per column walk the per-block partials with a fixed stride and are combined in a fixed order
__global__ void colsum_finalize_kernel(const float *__restrict__ partial, float *__restrict__ colsum, int cols, int out_ld,
                                       int blocks) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    float t = 0.f;
    if (c < cols)
        for (int b = sub; b < blocks; b += 8) t += __ldg(partial + (size_t)b * out_ld + c);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    if (c < cols && sub == 0) colsum[c] = t;
}

// operand rows = x's columns, operand columns (K) = x's rows
__global__ void split3x_transpose_kernel(const float *__restrict__ x, int rows, int cols, int64_t ld,
                                         __nv_bfloat16 *__restrict__ out, int64_t part_rows, int out_ld) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;          // tile of x
    const int tx = threadIdx.x, ty = threadIdx.y;                   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < rows && c < cols) ? __ldg(x + (size_t)r * ld + c) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int orow = c0 + j, ocol = r0 + tx;                    // operand row = x column
        if (orow < cols && ocol < out_ld)
            store_parts(out, (size_t)orow * out_ld + ocol, (size_t)part_rows * out_ld, tile[tx][j]);
    }
}

// both operands of one matrix in one pass: out = split(x), out_t = split(x^T)
__global__ void split3x_both_kernel(const float *__restrict__ x, int rows, int cols, int64_t ld, __nv_bfloat16 *__restrict__ out,
                                    int64_t part_rows, int out_ld, __nv_bfloat16 *__restrict__ out_t, int64_t part_rows_t,
                                    int out_ld_t) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        const float v = (r < rows && c < cols) ? __ldg(x + (size_t)r * ld + c) : 0.f;
        tile[j][tx] = v;
        if (r < rows && c < out_ld) store_parts(out, (size_t)r * out_ld + c, (size_t)part_rows * out_ld, v);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int orow = c0 + j, ocol = r0 + tx;
        if (orow < cols && ocol < out_ld_t)
            store_parts(out_t, (size_t)orow * out_ld_t + ocol, (size_t)part_rows_t * out_ld_t, tile[tx][j]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// operand [total_rows][ld] bf16, box = 64 columns (128 B) x box_rows rows, 128-byte swizzle
int make_operand_map(CUtensorMap *map, const void *ptr, int64_t total_rows, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return DDSP_B200_EUNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)total_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : DDSP_B200_EINVAL;
}

}  // namespace

// Padded K extent (columns) of a split operand whose logical K is `k`
extern "C" int64_t ddsp_b200_gemm3x_ld(int64_t k) { return (k + kBK - 1) / kBK * kBK; }

// Number of K splits ddsp_b200_gemm3x uses for this shape (workspace = splits * M * N floats when > 1)
extern "C" int ddsp_b200_gemm3x_splits(int M, int N, int K) {
    const int tiles = ((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
    const int kblocks = (int)(ddsp_b200_gemm3x_ld(K) / kBK);
    // only skinny outputs (weight gradients) are split; splitting K to even out a ragged last wave (200 tiles
    // on 148 SMs) was measured and lost to the extra reduction pass
    if (tiles >= 96 || kblocks < 8) return 1;
    int s = (DDSP_SM_COUNT + tiles - 1) / tiles;
    if (s > kblocks / 2) s = kblocks / 2;
    return s < 1 ? 1 : s;
}

// launch geometry of the non-transposed split: column groups and row lanes per block, rows per block (small
// matrices get fewer rows per block so that the grid still covers the SMs), blocks along the rows
struct SplitGeom { int out_ld, cgroups, lanes, rpb, row_blocks, col_blocks; };
static SplitGeom split_geometry(int64_t part_rows, int64_t cols) {
    SplitGeom g;
    g.out_ld = (int)ddsp_b200_gemm3x_ld(cols);
    g.cgroups = g.out_ld / 8 < 256 ? g.out_ld / 8 : 256;
    g.lanes = 256 / g.cgroups;
    int rpb = (int)((part_rows + 2 * DDSP_SM_COUNT - 1) / (2 * DDSP_SM_COUNT));
    rpb = (rpb + g.lanes - 1) / g.lanes * g.lanes;
    g.rpb = rpb < g.lanes ? g.lanes : (rpb > kSplitRows ? kSplitRows : rpb);
    g.row_blocks = (int)((part_rows + g.rpb - 1) / g.rpb);
    g.col_blocks = (g.out_ld / 8 + g.cgroups - 1) / g.cgroups;
    return g;
}

static int launch_split(const float *x, int64_t rows, int64_t cols, int64_t ld, __nv_bfloat16 *o, int64_t part_rows,
                        float *colsum, float *partial, cudaStream_t st) {
    const SplitGeom g = split_geometry(part_rows, cols);
    dim3 blocks((unsigned)g.row_blocks, (unsigned)g.col_blocks);
    if (colsum) {
        split3x_kernel<true><<<blocks, g.cgroups * g.lanes, (size_t)g.lanes * g.cgroups * 8 * 4, st>>>(
            x, (int)rows, (int)cols, ld, o, part_rows, g.out_ld, partial, g.cgroups, g.lanes, g.rpb);
        int s = ddsp_launch_status();
        if (s) return s;
        colsum_finalize_kernel<<<(unsigned)((cols * 8 + 127) / 128), 128, 0, st>>>(partial, colsum, (int)cols, g.out_ld, g.row_blocks);
    } else {
        split3x_kernel<false><<<blocks, g.cgroups * g.lanes, 0, st>>>(x, (int)rows, (int)cols, ld, o, part_rows, g.out_ld,
                                                                     nullptr, g.cgroups, g.lanes, g.rpb);
    }
    return ddsp_launch_status();
}

// Scratch floats ddsp_b200_gemm3x_split_colsum needs for a part_rows x cols operand
extern "C" int64_t ddsp_b200_gemm3x_colsum_scratch(int64_t part_rows, int64_t cols) {
    const SplitGeom g = split_geometry(part_rows, cols);
    return (int64_t)g.row_blocks * g.out_ld;
}

// gemm3x_split (non-transposed) that also returns colsum[c] = sum over rows of x[r][c] (a Linear layer's bias
// gradient when x = dy); partial: ddsp_b200_gemm3x_colsum_scratch(part_rows, cols) floats
extern "C" int ddsp_b200_gemm3x_split_colsum(const float *x, int64_t rows, int64_t cols, int64_t ld, void *out,
                                             int64_t part_rows, float *colsum, float *partial, void *stream) {
    DDSP_REQUIRE(x && out && colsum && partial && rows > 0 && cols > 0 && ld >= cols && part_rows >= rows);
    return launch_split(x, rows, cols, ld, static_cast<__nv_bfloat16 *>(out), part_rows, colsum, partial, (cudaStream_t)stream);
}

// out: [3 * part_rows][out_ld] bf16 (uint16_t bits) with out_ld = ddsp_b200_gemm3x_ld(K), part_rows >= operand
// rows.  transpose = 0: operand (rows x cols) = x; 1: operand (cols x rows) = x^T.  x has row pitch ld.
extern "C" int ddsp_b200_gemm3x_split(const float *x, int64_t rows, int64_t cols, int64_t ld, int transpose, void *out,
                                      int64_t part_rows, void *stream) {
    DDSP_REQUIRE(x && out && rows > 0 && cols > 0 && ld >= cols);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(out);
    if (!transpose) {
        DDSP_REQUIRE(part_rows >= rows);
        return launch_split(x, rows, cols, ld, o, part_rows, nullptr, nullptr, st);
    } else {
        DDSP_REQUIRE(part_rows >= cols);
        const int out_ld = (int)ddsp_b200_gemm3x_ld(rows);
        dim3 grid((unsigned)((cols + 31) / 32), (unsigned)(out_ld / 32));
        split3x_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(x, (int)rows, (int)cols, ld, o, part_rows, out_ld);
    }
    return ddsp_launch_status();
}

// out = split operand of x (parts part_rows apart), out_t = split operand of x^T (parts part_rows_t apart)
extern "C" int ddsp_b200_gemm3x_split_both(const float *x, int64_t rows, int64_t cols, int64_t ld, void *out,
                                           int64_t part_rows, void *out_t, int64_t part_rows_t, void *stream) {
    DDSP_REQUIRE(x && out && out_t && rows > 0 && cols > 0 && ld >= cols && part_rows >= rows && part_rows_t >= cols);
    const int out_ld = (int)ddsp_b200_gemm3x_ld(cols), out_ld_t = (int)ddsp_b200_gemm3x_ld(rows);
    dim3 grid((unsigned)(out_ld / 32), (unsigned)(out_ld_t / 32));
    split3x_both_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        x, (int)rows, (int)cols, ld, static_cast<__nv_bfloat16 *>(out), part_rows, out_ld,
        static_cast<__nv_bfloat16 *>(out_t), part_rows_t, out_ld_t);
    return ddsp_launch_status();
}

// C[M][N] (row pitch ldc) = A B^T + bias, A logically M x K, B logically N x K.  Each operand is a split
// matrix (ddsp_b200_gemm3x_split, parts x_part_rows rows apart, row pitch x_ld elements) stored either
//   K-major  (x_mn = 0): rows = M (or N), columns = K, x_ld = gemm3x_ld(K), or
//   MN-major (x_mn = 1): rows = K, columns = M (or N), x_ld = gemm3x_ld(M or N); x_part_rows must then be a
//   multiple of 64 with the rows K..x_part_rows zero (gemm3x_split does that when given such a part_rows),
// so that a matrix split once serves as operand of y = x W^T, dx = dy W and dW = dy^T x without transposes.
// workspace: splits * M * N floats or NULL.
extern "C" int ddsp_b200_gemm3x(const void *a, int64_t a_lo_row, int64_t a_ld, int a_mn, const void *b, int64_t b_lo_row,
                                int64_t b_ld, int b_mn, const float *bias, float *c, int64_t ldc, int M, int N, int K,
                                float *workspace, void *stream) {
    DDSP_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0 && ldc >= N);
    DDSP_REQUIRE(a_mn ? (a_lo_row >= K && a_lo_row % 64 == 0 && a_ld >= M) : (a_lo_row >= M && a_ld >= K));
    DDSP_REQUIRE(b_mn ? (b_lo_row >= K && b_lo_row % 64 == 0 && b_ld >= N) : (b_lo_row >= N && b_ld >= K));
    DDSP_REQUIRE(a_ld % 8 == 0 && b_ld % 8 == 0);
    const int kblocks = (K + kBK - 1) / kBK;
    int splits = ddsp_b200_gemm3x_splits(M, N, K);
    if (splits > 1 && !workspace) splits = 1;
    alignas(64) CUtensorMap map_a, map_b;
    int s = make_operand_map(&map_a, a, 3 * a_lo_row, a_ld, a_mn ? 64 : kBM);
    if (s) return s;
    s = make_operand_map(&map_b, b, 3 * b_lo_row, b_ld, b_mn ? 64 : kBN);
    if (s) return s;
    GemmParams p;
    p.c = splits > 1 ? workspace : c;
    p.bias = splits > 1 ? nullptr : bias;
    p.M = M;
    p.N = N;
    p.ldc = splits > 1 ? N : (int)ldc;
    p.a_lo_row = (int)a_lo_row;
    p.b_lo_row = (int)b_lo_row;
    p.a_mn = a_mn;
    p.b_mn = b_mn;
    p.k_blocks = kblocks;
    p.k_blocks_per_split = (kblocks + splits - 1) / splits;
    p.splits = splits;
    p.m_tiles = (M + kBM - 1) / kBM;
    p.n_tiles = (N + kBN - 1) / kBN;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e != cudaSuccess) return (int)e;
    const int tiles = p.m_tiles * p.n_tiles * splits;
    gemm3x_kernel<<<tiles < DDSP_SM_COUNT ? tiles : DDSP_SM_COUNT, kGemmThreads, kGemmSmem, st>>>(map_a, map_b, p);
    s = ddsp_launch_status();
    if (s || splits == 1) return s;
    const size_t total = (size_t)M * N;
    gemm3x_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(workspace, bias, c, M, N, (int)ldc, splits);
    return ddsp_launch_status();
}
