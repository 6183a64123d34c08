// Dense operator application on the 5th-generation tensor cores (Matrix.apply,
// dense_cublas.py:732-776 -- the reference calls cublasSgemm here):
//
//   transp == 0:  Y[v,i] = alpha * sum_j X[v,j]*A[i,j] + beta*Y[v,i]     A (M,N) row-major fp32
//   transp != 0:  Y[v,j] = alpha * sum_i X[v,i]*A[i,j] + beta*Y[v,j]
//
// fp32 accuracy from TF32 tensor cores by the 3xTF32 split: every operand is
// written as hi + lo with hi = the tf32 truncation the tensor core applies to an
// fp32 word and lo = a - hi (exact in fp32), and  X.A ~= Xhi.Ahi + Xhi.Alo + Xlo.Ahi.
// The lo parts are materialised in global memory (the data matrix's once per
// matrix version, the block's per call), so the kernel is a pure TMA -> shared
// memory -> tcgen05.mma -> TMEM pipeline:
//
//   warp 0   : TMA producer (one elected lane): 4 swizzled tiles per stage
//              (Xhi, Xlo: 128 vectors x 32 K;  Ahi, Alo: 128 outputs x 32 K)
//   warp 1   : tcgen05.mma issuer (one elected lane), 12 MMAs 128x128x8 per stage,
//              fp32 accumulators in TMEM (2 x 128 columns, double buffered)
//   warps 2-5: epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers
//              -> alpha/beta -> global
// The grid is persistent (one CTA per SM).  Work units are (vector block, output
// tile, K split); the split factor is chosen so the units fill whole waves of
// SMs, partial tiles go to fixed workspace slots and a second kernel adds them
// in a fixed order -> deterministic.
//
// UMMA shared-memory descriptors follow the canonical SWIZZLE_128B layouts
// (K-major: SWIZZLE_128B, 8-row x 128-byte atoms, SBO = 1024 B;  MN-major tf32:
// SWIZZLE_128B_BASE32B, 32-element x 4-K atoms, LBO = bytes between 32-wide MN
// blocks, SBO = 512 B), which is exactly what a TMA box with
// CU_TENSOR_MAP_SWIZZLE_128B (resp. ..._128B_ATOM_32B) and a 128-byte inner
// extent writes.  Out-of-bounds parts of a box are zero-filled by TMA, so ragged
// M, N, K and fewer than 128 vectors need no special code in the main loop.
#include <cuda.h>
#include "common.cuh"
#include "tma.cuh"

namespace rl {

constexpr int TC_BM = 128;        // vectors per tile  (UMMA M)
constexpr int TC_BN = 128;        // outputs per tile  (UMMA N)
constexpr int TC_BK = 32;         // fp32 elements of K per stage = one 128-byte swizzle span
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;             // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;            // Xhi, Xlo, Ahi, Alo
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 + 256;

struct TcParams {
    int64_t nout, kred, ldy, ld_ws;
    float* y;
    float* ws;            // partials [split][k_pad][ld_ws] when splits > 1
    int k, tiles, vblocks, splits, kblocks, kb_per_split, transp;
    float alpha, beta;
};

// mbarrier / TMA wrappers: tma.cuh
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
// layout: 2 = SWIZZLE_128B (16-byte chunks XOR row%8; K-major operands),
//         1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4) -- the only layout the
//             tensor core accepts for MN-major 32-bit (tf32) operands.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout = 2) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

// INSPLIT (EXPERIMENTAL, knob 8, unmeasured): only the raw fp32 tiles of X and A are loaded; four extra
// warps compute the low parts (a - tf32_trunc(a), elementwise, hence layout-agnostic) from shared memory
// into the Xlo / Alo slots of the stage and hand the stage to the MMA issuer through ready[].  Halves
// the HBM traffic (no materialised lo copy of the data matrix) and the L2->SM traffic of the kernel.
template <bool INSPLIT>
__global__ void __launch_bounds__(INSPLIT ? TC_THREADS + 128 : TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
               const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
               const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full = bars;                         // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;            // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;        // [2]
    uint64_t* tempty = bars + 2 * TC_STAGES + 2;   // [2]
    uint64_t* ready = bars + 2 * TC_STAGES + 4;    // [TC_STAGES] (INSPLIT: lo tiles written)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.vblocks * p.tiles * p.splits;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&ready[s], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int split = u % p.splits, tile = (u / p.splits) % p.tiles, vb = u / (p.splits * p.tiles);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * TC_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], INSPLIT ? 2 * TC_TILE_BYTES : TC_STAGE_BYTES);
                    tma_load_2d(st, &tm_xhi, &full[stage], kb * TC_BK, vb * TC_BM);
                    if (!INSPLIT) tma_load_2d(st + TC_TILE_BYTES, &tm_xlo, &full[stage], kb * TC_BK, vb * TC_BM);
                    if (!p.transp) {
                        tma_load_2d(st + 2 * TC_TILE_BYTES, &tm_ahi, &full[stage], kb * TC_BK, tile * TC_BN);
                        if (!INSPLIT) tma_load_2d(st + 3 * TC_TILE_BYTES, &tm_alo, &full[stage], kb * TC_BK, tile * TC_BN);
                    } else {
#pragma unroll
                        for (int c = 0; c < TC_BN / 32; ++c) {
                            tma_load_2d(st + 2 * TC_TILE_BYTES + c * 4096, &tm_ahi, &full[stage], tile * TC_BN + c * 32, kb * TC_BK);
                            if (!INSPLIT) tma_load_2d(st + 3 * TC_TILE_BYTES + c * 4096, &tm_alo, &full[stage], tile * TC_BN + c * 32, kb * TC_BK);
                        }
                    }
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.transp ? 1u : 0u) << 16) |
                                   ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
                const int split = u % p.splits;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * TC_BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(INSPLIT ? &ready[stage] : &full[stage], phase);
                    tc_fence_after();
                    const uint32_t sx_hi = smem_u32(smem + stage * TC_STAGE_BYTES);
                    const uint32_t sx_lo = sx_hi + TC_TILE_BYTES;
                    const uint32_t sa_hi = sx_hi + 2 * TC_TILE_BYTES;
                    const uint32_t sa_lo = sx_hi + 3 * TC_TILE_BYTES;
#pragma unroll
                    for (int s = 0; s < TC_BK / 8; ++s) {
                        const uint64_t xh = umma_desc(sx_hi + s * 32, 16, 1024);
                        const uint64_t xl = umma_desc(sx_lo + s * 32, 16, 1024);
                        uint64_t ah, al;
                        if (!p.transp) {
                            ah = umma_desc(sa_hi + s * 32, 16, 1024);
                            al = umma_desc(sa_lo + s * 32, 16, 1024);
                        } else {
                            // MN-major: 32-float x 4-K atoms (512 B); MN blocks 4096 B apart
                            ah = umma_desc(sa_hi + s * 1024, 4096, 512, 1);
                            al = umma_desc(sa_lo + s * 1024, 4096, 512, 1);
                        }
                        // small terms first, then the leading product
                        tc_mma_tf32(d_tmem, xl, ah, idesc, (kb > kb0 || s > 0) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, xh, al, idesc, 1u);
                        tc_mma_tf32(d_tmem, xh, ah, idesc, 1u);
                    }
                    tc_commit(&empty[stage]);            // frees the smem stage when the MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[as]);                   // accumulator complete
            }
        }
    } else if (warp >= 6) {
        // ===== INSPLIT only: low-part warps (6..9) =====
        // same (unit, k-block) walk as the producer; each stage: wait for the raw tiles, write
        // lo = a - tf32_trunc(a) of the X and A tiles into the slots behind them (same offsets, so the
        // swizzled layout carries over), make the writes visible to the async proxy, hand over
        uint8_t* smem_s = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
        const int tid = threadIdx.x - TC_THREADS;
        int stage = 0; uint32_t phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int split = u % p.splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                uint8_t* st = smem_s + stage * TC_STAGE_BYTES;
#pragma unroll
                for (int tsel = 0; tsel < 2; ++tsel) {
                    const float4* src = reinterpret_cast<const float4*>(st + 2 * tsel * TC_TILE_BYTES);
                    float4* dst = reinterpret_cast<float4*>(st + (2 * tsel + 1) * TC_TILE_BYTES);
#pragma unroll
                    for (int i = 0; i < TC_TILE_BYTES / 16 / 128; ++i) {
                        const float4 a = src[tid + i * 128];
                        float4 o;
                        o.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u);
                        o.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
                        o.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u);
                        o.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
                        dst[tid + i * 128] = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps (2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        int it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
            const int split = u % p.splits, tile = (u / p.splits) % p.tiles, vb = u / (p.splits * p.tiles);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int v = vb * TC_BM + q * 32 + lane;
            const bool vok = v < p.k;
            float* row = p.splits > 1 ? p.ws + ((int64_t)split * p.vblocks * TC_BM + v) * p.ld_ws
                                      : p.y + (int64_t)v * p.ldy;
            const bool vec_ok = p.splits > 1 || ((p.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TC_BN + c * 32), r);
                const int64_t o0 = (int64_t)tile * TC_BN + c * 32;
                if (!vok || o0 >= p.nout) continue;
                if (p.splits > 1) {
                    // partial sums: plain store, padded row => always in bounds and aligned
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(row + o0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                } else if (vec_ok && o0 + 32 <= p.nout) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(p.alpha * __uint_as_float(r[j]), p.alpha * __uint_as_float(r[j + 1]),
                                               p.alpha * __uint_as_float(r[j + 2]), p.alpha * __uint_as_float(r[j + 3]));
                        if (p.beta != 0.f) {
                            float4 old = *reinterpret_cast<const float4*>(row + o0 + j);
                            o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                        }
                        *reinterpret_cast<float4*>(row + o0 + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (o0 + j < p.nout) {
                            float o = p.alpha * __uint_as_float(r[j]);
                            if (p.beta != 0.f) o += p.beta * row[o0 + j];
                            row[o0 + j] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// Y = alpha * sum_s ws[s] + beta * Y   (fixed summation order)
__global__ void __launch_bounds__(256)
gemm_tc_reduce_kernel(const float* __restrict__ ws, int64_t ld_ws, int64_t slab, int splits, float* __restrict__ y,
                      int64_t ldy, int k, int64_t nout, float alpha, float beta) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (o >= nout || v >= k) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(int64_t)s * slab + (int64_t)v * ld_ws + o];
    float* dst = y + (int64_t)v * ldy + o;
    *dst = beta != 0.f ? alpha * acc + beta * *dst : alpha * acc;
}

// lo = a - tf32_trunc(a): the part of an fp32 word the tensor core drops
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst,
                  int64_t rows, int64_t cols) {
    const int64_t c4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        if (c4 + 4 <= cols) {
            float4 a = *reinterpret_cast<const float4*>(src + r * ld_src + c4);
            float4 o;
            o.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u);
            o.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
            o.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u);
            o.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(dst + r * ld_dst + c4) = o;
        } else {
            for (int64_t c = c4; c < cols; ++c) {
                float a = src[r * ld_src + c];
                dst[r * ld_dst + c] = a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------
struct TcPlan { int tiles, vblocks, splits, kblocks, kb_per_split; int64_t ld_ws; size_t ws_bytes, xlo_bytes; int64_t ld_xlo; };

static TcPlan tc_plan(int64_t M, int64_t N, int64_t k, int transp) {
    TcPlan pl;
    const int64_t nout = transp ? N : M, kred = transp ? M : N;
    pl.tiles = (int)((nout + TC_BN - 1) / TC_BN);
    pl.vblocks = (int)((k + TC_BM - 1) / TC_BM);
    pl.kblocks = (int)((kred + TC_BK - 1) / TC_BK);
    const int sms = sm_count();
    const int64_t base = (int64_t)pl.tiles * pl.vblocks;
    int best = 1; double best_eff = 0.0;
    const int max_split = pl.kblocks / 8 > 0 ? (pl.kblocks / 8 < 32 ? pl.kblocks / 8 : 32) : 1;   // >= 8 k-blocks per unit
    for (int s = 1; s <= max_split; ++s) {
        const int64_t units = base * s;
        const double eff = (double)units / (double)(((units + sms - 1) / sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
        if (eff >= 0.93) { best = s; break; }
    }
    pl.splits = best;
    pl.kb_per_split = (pl.kblocks + pl.splits - 1) / pl.splits;
    pl.splits = (pl.kblocks + pl.kb_per_split - 1) / pl.kb_per_split;
    pl.ld_ws = (int64_t)pl.tiles * TC_BN;
    pl.ld_xlo = (kred + 31) / 32 * 32;
    pl.xlo_bytes = ((size_t)k * pl.ld_xlo * 4 + 255) & ~size_t(255);
    pl.ws_bytes = pl.xlo_bytes + (pl.splits > 1 ? (size_t)pl.splits * pl.vblocks * TC_BM * pl.ld_ws * 4 : 0);
    return pl;
}

bool gemm_tc_supported(const void* a, int64_t lda, const void* x, int64_t ldx) {
    return tma_encode_fn() != nullptr && host_aligned16(a) && host_aligned16(x) && (lda % 4 == 0) && (ldx % 4 == 0);
}

int gemm_tc(const float* a_hi, const float* a_lo_in, int64_t lda, int64_t M, int64_t N, const float* x, int64_t ldx,
            float* y, int64_t ldy, int64_t k, int transp, double alpha, double beta, void* ws, size_t ws_bytes,
            cudaStream_t st) {
    TcPlan pl = tc_plan(M, N, k, transp);
    if (ws_bytes < pl.ws_bytes) return RL_E_WORKSPACE;
    const float* a_lo = a_lo_in;
    const int64_t nout = transp ? N : M, kred = transp ? M : N;
    float* x_lo = reinterpret_cast<float*>(ws);
    // low parts of the 3xTF32 split computed in shared memory by four extra warps (default since r2:
    // bit-identical results, half the HBM traffic, no 1.9 GB lo copy of the data matrix; measured
    // 0.659 vs 0.681 ms at config 2) unless a materialised lo copy is handed in (knob -1 / a_lo != NULL)
    const bool insplit = a_lo_in == nullptr || g_knob[KNOB_GEMM_INSPLIT] == 1;
    if (!insplit) {
        dim3 g((unsigned)((kred / 4 + 256) / 256), (unsigned)(k < 65535 ? k : 65535));
        split_tf32_kernel<<<g, 256, 0, st>>>(x, ldx, x_lo, pl.ld_xlo, k, kred);
        int rc = check_launch();
        if (rc) return rc;
    }
    CUtensorMap mxh, mxl, mah, mal;
    if (insplit) { x_lo = const_cast<float*>(x); a_lo = a_hi; }     // the lo maps are encoded but never used
    int rc = make_map(&mxh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, kred, k, ldx, TC_BK, TC_BM);
    if (!rc) rc = make_map(&mxl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x_lo, kred, k, insplit ? ldx : pl.ld_xlo, TC_BK, TC_BM);
    if (!transp) {
        if (!rc) rc = make_map(&mah, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_hi, N, M, lda, TC_BK, TC_BN);
        if (!rc) rc = make_map(&mal, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_lo, N, M, lda, TC_BK, TC_BN);
    } else {
        if (!rc) rc = make_map(&mah, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_hi, N, M, lda, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (!rc) rc = make_map(&mal, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_lo, N, M, lda, 32, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    }
    if (rc) return rc;
    TcParams p;
    p.nout = nout; p.kred = kred; p.ldy = ldy; p.ld_ws = pl.ld_ws; p.y = y;
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.xlo_bytes);
    p.k = (int)k; p.tiles = pl.tiles; p.vblocks = pl.vblocks; p.splits = pl.splits; p.kblocks = pl.kblocks;
    p.kb_per_split = pl.kb_per_split; p.transp = transp ? 1 : 0; p.alpha = (float)alpha; p.beta = (float)beta;
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        RL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        configured = true;
    }
    const int units = pl.tiles * pl.vblocks * pl.splits;
    const int grid = units < sm_count() ? units : sm_count();
    if (insplit) gemm_tc_kernel<true><<<grid, TC_THREADS + 128, TC_SMEM, st>>>(mxh, mxl, mah, mal, p);
    else gemm_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, st>>>(mxh, mxl, mah, mal, p);
    rc = check_launch();
    if (rc) return rc;
    if (pl.splits > 1) {
        dim3 g((unsigned)((nout + 255) / 256), (unsigned)k);
        gemm_tc_reduce_kernel<<<g, 256, 0, st>>>(p.ws, pl.ld_ws, (int64_t)pl.vblocks * TC_BM * pl.ld_ws, pl.splits, y,
                                                 ldy, (int)k, nout, (float)alpha, (float)beta);
        rc = check_launch();
    }
    return rc;
}

size_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t k, int transp) { return tc_plan(M, N, k, transp).ws_bytes; }

int split_tf32(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
               cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 g((unsigned)((cols / 4 + 256) / 256), (unsigned)(rows < 65535 ? rows : 65535));
    split_tf32_kernel<<<g, 256, 0, st>>>(src, ld_src, dst, ld_dst, rows, cols);
    return check_launch();
}

}  // namespace rl
