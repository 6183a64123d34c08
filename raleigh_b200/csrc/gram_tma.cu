// TMA-fed variant of the fp64 Gram kernel (gram.cu).
//
// gram_dmma_kernel is bound by the number of load REQUESTS it can keep in flight
// (profiles/r1c_kernel_tuning.md: going from 128-bit to 256-bit fragment loads took it
// from 41 % to 64 % of HBM peak, more warps did nothing; its 248 registers leave no room
// for a second set of fragments in flight).  Here a producer lane moves 16-row x
// (8*NI | 8*NJ)-vector boxes with cp.async.bulk.tensor into a shared-memory ring, so
// memory latency is hidden by the ring instead of by registers, and four consumer warps
// read their DMMA fragments with conflict-free 128-bit shared loads (128B-swizzled rows:
// lane (g, c) needs bytes [32c, 32c+32) of vector row g, 16-byte chunks 2c^g and (2c+1)^g).
// Out-of-range rows / vectors are zero-filled by TMA, so there is no tail code.
//
// Stage layout: 4 warps x KSUB sub-steps x { O box (NI KB) | S box (NJ KB) }, where a box is
// 16 rows of 8*NI (8*NJ) vectors; KSUB is chosen so that a warp consumes ~8 KB per stage.
#include "common.cuh"
#include "tma.cuh"

namespace rl {

constexpr int GT_THREADS = 160;
#ifndef RL_GRAM_TMA_WAVES
#define RL_GRAM_TMA_WAVES 4
#endif                // 4 consumer warps + 1 producer warp

__host__ __device__ constexpr int gt_ksub(int ni, int nj) { return (8 / (ni + nj)) > 0 ? 8 / (ni + nj) : 1; }
__host__ __device__ constexpr int gt_stage_bytes(int ni, int nj) { return 4 * gt_ksub(ni, nj) * (ni + nj) * 1024; }
__host__ __device__ constexpr int gt_smem(int ni, int nj, int stages) { return stages * gt_stage_bytes(ni, nj) + 1024 + 256; }

__device__ __forceinline__ void gt_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NI, int NJ, bool SAME, int STAGES, int MINB>
__global__ void __launch_bounds__(GT_THREADS, MINB)
gram_tma_kernel(const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_s, int m, int k,
                int64_t n, int64_t rows_per_cta, int interleave, double* __restrict__ part) {
    constexpr int KSUB = gt_ksub(NI, NJ);
    constexpr int OBOX = NI * 1024, SBOX = NJ * 1024, SUB = OBOX + SBOX;
    constexpr int STAGE = gt_stage_bytes(NI, NJ);
    constexpr int ROWS = 64 * KSUB;                      // rows per stage
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the __shared__ symbol keeps the address space (LDS/STS, not generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // multi-tile products come as a one-dimensional grid, the tiles of a row chunk adjacent (see gram.cu)
    int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, nchunks = gridDim.x;
    if (gridDim.y == 1 && gridDim.z == 1) {
        const int tj = (m + 8 * NJ - 1) / (8 * NJ), ti = (k + 8 * NI - 1) / (8 * NI);
        const int t = bx % (tj * ti);
        bx /= tj * ti;
        nchunks /= tj * ti;
        by = t % tj;
        bz = t / tj;
    }
    const int i0 = bz * (8 * NI), j0 = by * (8 * NJ);
    // contiguous: CTA c owns rows [c * rows_per_cta, (c+1) * rows_per_cta); interleaved: CTA c owns
    // stages c, c + grid, c + 2 grid, ... so that the CTAs running at any moment stream one
    // contiguous region of every vector (whole DRAM pages) instead of grid-many 512-byte pieces
    const int64_t r_begin = interleave ? (int64_t)bx * ROWS : (int64_t)bx * rows_per_cta;
    const int64_t r_step = interleave ? (int64_t)nchunks * ROWS : (int64_t)ROWS;
    const int64_t r_end = interleave ? n : (r_begin + rows_per_cta < n ? r_begin + rows_per_cta : n);
    const int nstages = r_begin < r_end ? (int)((r_end - r_begin + r_step - 1) / r_step) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < nstages; ++it) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * STAGE;
                mbar_expect_tx(&full[stage], 4 * KSUB * (SAME ? OBOX : SUB));
                const int64_t r0 = r_begin + (int64_t)it * r_step;
#pragma unroll
                for (int q = 0; q < 4 * KSUB; ++q) {
                    // sub-step q = warp * KSUB + u covers rows r0 + 16 q
                    tma_load_2d(st + q * SUB, &tm_o, &full[stage], (int)(r0 + 16 * q), i0);
                    if (!SAME) tma_load_2d(st + q * SUB + OBOX, &tm_s, &full[stage], (int)(r0 + 16 * q), j0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---- consumers: warp w owns sub-steps w*KSUB .. w*KSUB+KSUB-1 of every stage ----------
    const int g = lane >> 2, c = lane & 3;
    const uint32_t ch0 = (uint32_t)((2 * c) ^ g) * 16, ch1 = (uint32_t)((2 * c + 1) ^ g) * 16;
    double acc[NI][NJ][2];
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < nstages; ++it) {
        mbar_wait(&full[stage], phase);
#pragma unroll
        for (int u = 0; u < KSUB; ++u) {
            const uint8_t* bo = smem + stage * STAGE + (warp * KSUB + u) * SUB;
            const uint8_t* bs = bo + OBOX;
            double fa[NI][4], fb[NJ][4];
#pragma unroll
            for (int t = 0; t < NI; ++t) {
                const uint8_t* row = bo + (8 * t + g) * 128;
                const double2 x0 = *reinterpret_cast<const double2*>(row + ch0);
                const double2 x1 = *reinterpret_cast<const double2*>(row + ch1);
                fa[t][0] = x0.x; fa[t][1] = x0.y; fa[t][2] = x1.x; fa[t][3] = x1.y;
            }
            if (SAME) {
#pragma unroll
                for (int t = 0; t < NJ; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) fb[t][e] = fa[t < NI ? t : 0][e];
            } else {
#pragma unroll
                for (int t = 0; t < NJ; ++t) {
                    const uint8_t* row = bs + (8 * t + g) * 128;
                    const double2 x0 = *reinterpret_cast<const double2*>(row + ch0);
                    const double2 x1 = *reinterpret_cast<const double2*>(row + ch1);
                    fb[t][0] = x0.x; fb[t][1] = x0.y; fb[t][2] = x1.x; fb[t][3] = x1.y;
                }
            }
            if (u == KSUB - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);   // fragments are in registers: release the slot
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) gt_dmma(acc[a][b][0], acc[a][b][1], fa[a][s], fb[b][s]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // cross-warp reduction through the (now idle) ring: every TMA write has been consumed
    asm volatile("bar.sync 1, 128;" ::: "memory");
    constexpr int TILE = NI * NJ * 64;
    double* red = reinterpret_cast<double*>(smem);            // 4 warps x TILE doubles <= 32 KB
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b) {
            red[warp * TILE + (a * NJ + b) * 64 + g * 8 + 2 * c] = acc[a][b][0];
            red[warp * TILE + (a * NJ + b) * 64 + g * 8 + 2 * c + 1] = acc[a][b][1];
        }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    double* out = part + (int64_t)bx * k * m;
    for (int e = threadIdx.x; e < TILE; e += 128) {
        const double v = (red[e] + red[TILE + e]) + (red[2 * TILE + e] + red[3 * TILE + e]);
        const int blk = e >> 6, a = blk / NJ, b = blk % NJ;
        const int i = i0 + 8 * a + ((e & 63) >> 3), j = j0 + 8 * b + (e & 7);
        if (i < k && j < m) out[(int64_t)i * m + j] = v;
    }
}

struct GramTmaPlan { int ni, nj, tiles_i, tiles_j, chunks, stages, minb, interleave; int64_t rows_per_cta; };

// mode: 1 = deep ring, one CTA per SM; 2 = shallower ring, two CTAs per SM (two consumer warps
// per scheduler, so one warp's shared-memory reads overlap the other's DMMAs)
static GramTmaPlan gram_tma_plan(int64_t m, int64_t k, int64_t n, int mode, bool same = false) {
    GramTmaPlan p;
    auto frag = [](int64_t v) { return v <= 8 ? 1 : v <= 16 ? 2 : 4; };
    p.ni = frag(k); p.nj = frag(m);
    p.tiles_i = (int)((k + 8 * p.ni - 1) / (8 * p.ni));
    p.tiles_j = (int)((m + 8 * p.nj - 1) / (8 * p.nj));
    p.interleave = (mode & 4) ? 1 : 0;
    mode &= 3;
    p.minb = mode == 2 ? 2 : 1;
    p.stages = mode == 2 ? 3 : 5;
    const int rows = 64 * gt_ksub(p.ni, p.nj);
    const int64_t tiles = (int64_t)p.tiles_i * p.tiles_j;
    // CTAs per SM slot: with one CTA per slot (static partition) ncu shows SMs busy 45..100 % of the
    // kernel (some SMs are served faster by the memory system); several CTAs per slot let the block
    // scheduler even that out at the price of one partial tile + ring ramp-up per CTA
    // measured (profiles/r1e_gram_spmm.md): 4 CTAs per slot pay off only for the DMMA-heavy tiles
    // (32x32, 32x16) of X^T Y; X^T X and the small tiles are fastest with the static partition
    const int waves = g_knob[KNOB_GRAM_WAVES] > 0 ? g_knob[KNOB_GRAM_WAVES]
                                                  : ((p.ni + p.nj >= 6 && !same) ? RL_GRAM_TMA_WAVES : 1);
    int64_t want = ((int64_t)sm_count() * 2 * waves + tiles - 1) / tiles;
    const int64_t maxc = (n + 8 * rows - 1) / (8 * rows);                // at least 8 stages per CTA
    if (want > maxc) want = maxc;
    if (want > (int64_t)sm_count() * 2 * 32) want = (int64_t)sm_count() * 2 * 32;
    if (want < 1) want = 1;
    int64_t rpc = (n + want - 1) / want;
    rpc = (rpc + rows - 1) / rows * rows;
    p.rows_per_cta = rpc;
    p.chunks = (int)((n + rpc - 1) / rpc);
    return p;
}


// ---- persistent variant with dynamic row-chunk scheduling (mode 3) -----------------------------
// ncu (profiles/r1e_gram_spmm.md): with one CTA per slot and equal row ranges the SMs are busy
// between 45 % and 100 % of the kernel -- the memory system does not serve them equally -- and
// several CTAs per slot (block scheduler) pay a ring ramp-up each.  Here 2 CTAs per SM stay
// resident, the producer lane pulls chunks of `chunk_stages` stages from an atomic counter and
// keeps the ring full ACROSS chunk boundaries; the chunk id travels with each stage, and when a
// chunk's last stage has been consumed the four consumer warps fold their tiles in fixed order
// (w0 + w1 + w2 + w3 through an 8 KB buffer) into that chunk's partial slot.  Which CTA computes
// a chunk varies from run to run, its value does not: the result stays bit-reproducible.
constexpr int GD_STAGES = 3;
__host__ __device__ constexpr int gd_smem(int ni, int nj) {
    return GD_STAGES * gt_stage_bytes(ni, nj) + ni * nj * 64 * 8 + 2 * GD_STAGES * 8 + 2 * GD_STAGES * 4 + 1024 + 64;
}

template <int NI, int NJ, bool SAME>
__global__ void __launch_bounds__(GT_THREADS, 2)
gram_tma_dyn_kernel(const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_s, int m, int k,
                    int64_t n, int chunk_stages, int nchunks, int* __restrict__ counter, double* __restrict__ part) {
    constexpr int KSUB = gt_ksub(NI, NJ);
    constexpr int OBOX = NI * 1024, SBOX = NJ * 1024, SUB = OBOX + SBOX;
    constexpr int STAGE = gt_stage_bytes(NI, NJ);
    constexpr int ROWS = 64 * KSUB;
    constexpr int TILE = NI * NJ * 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    double* red = reinterpret_cast<double*>(smem + GD_STAGES * STAGE);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GD_STAGES * STAGE + TILE * 8);
    uint64_t* empty = full + GD_STAGES;
    volatile int* meta = reinterpret_cast<volatile int*>(empty + GD_STAGES);     // {chunk id, last stage of chunk} per stage
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (;;) {
                const int c = atomicAdd(counter, 1);
                if (c >= nchunks) break;
                const int64_t rc0 = (int64_t)c * chunk_stages * ROWS;
                int64_t left = (n - rc0 + ROWS - 1) / ROWS;
                const int ns = left < chunk_stages ? (int)left : chunk_stages;
                for (int s = 0; s < ns; ++s) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    meta[2 * stage] = c;
                    meta[2 * stage + 1] = (s == ns - 1);
                    uint8_t* st = smem + stage * STAGE;
                    mbar_expect_tx(&full[stage], 4 * KSUB * (SAME ? OBOX : SUB));     // release: meta is visible
                    const int64_t r0 = rc0 + (int64_t)s * ROWS;
#pragma unroll
                    for (int q = 0; q < 4 * KSUB; ++q) {
                        tma_load_2d(st + q * SUB, &tm_o, &full[stage], (int)(r0 + 16 * q), 0);
                        if (!SAME) tma_load_2d(st + q * SUB + OBOX, &tm_s, &full[stage], (int)(r0 + 16 * q), 0);
                    }
                    if (++stage == GD_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // no more work: a stage that carries no data tells the consumers to stop
            mbar_wait(&empty[stage], phase ^ 1);
            meta[2 * stage] = -1;
            mbar_arrive(&full[stage]);
        }
        return;
    }

    const int g = lane >> 2, c4 = lane & 3;
    const uint32_t ch0 = (uint32_t)((2 * c4) ^ g) * 16, ch1 = (uint32_t)((2 * c4 + 1) ^ g) * 16;
    double acc[NI][NJ][2];
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    int stage = 0; uint32_t phase = 0;
    for (;;) {
        mbar_wait(&full[stage], phase);
        const int ck = meta[2 * stage];
        if (ck < 0) break;
        const int last = meta[2 * stage + 1];
#pragma unroll
        for (int u = 0; u < KSUB; ++u) {
            const uint8_t* bo = smem + stage * STAGE + (warp * KSUB + u) * SUB;
            const uint8_t* bs = bo + OBOX;
            double fa[NI][4], fb[NJ][4];
#pragma unroll
            for (int t = 0; t < NI; ++t) {
                const uint8_t* row = bo + (8 * t + g) * 128;
                const double2 x0 = *reinterpret_cast<const double2*>(row + ch0);
                const double2 x1 = *reinterpret_cast<const double2*>(row + ch1);
                fa[t][0] = x0.x; fa[t][1] = x0.y; fa[t][2] = x1.x; fa[t][3] = x1.y;
            }
            if (SAME) {
#pragma unroll
                for (int t = 0; t < NJ; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) fb[t][e] = fa[t < NI ? t : 0][e];
            } else {
#pragma unroll
                for (int t = 0; t < NJ; ++t) {
                    const uint8_t* row = bs + (8 * t + g) * 128;
                    const double2 x0 = *reinterpret_cast<const double2*>(row + ch0);
                    const double2 x1 = *reinterpret_cast<const double2*>(row + ch1);
                    fb[t][0] = x0.x; fb[t][1] = x0.y; fb[t][2] = x1.x; fb[t][3] = x1.y;
                }
            }
            if (u == KSUB - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) gt_dmma(acc[a][b][0], acc[a][b][1], fa[a][s], fb[b][s]);
        }
        if (++stage == GD_STAGES) { stage = 0; phase ^= 1; }
        if (last) {
            // fold the four warp tiles in fixed order: w0 writes, w1..w3 add in turn
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (warp == w) {
#pragma unroll
                    for (int a = 0; a < NI; ++a)
#pragma unroll
                        for (int b = 0; b < NJ; ++b) {
                            double* p2 = red + (a * NJ + b) * 64 + g * 8 + 2 * c4;
                            if (w == 0) { p2[0] = acc[a][b][0]; p2[1] = acc[a][b][1]; }
                            else { p2[0] += acc[a][b][0]; p2[1] += acc[a][b][1]; }
                        }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            double* out = part + (int64_t)ck * k * m;
            for (int e = threadIdx.x; e < TILE; e += 128) {
                const int blk = e >> 6, a = blk / NJ, b = blk % NJ;
                const int i = 8 * a + ((e & 63) >> 3), j = 8 * b + (e & 7);
                if (i < k && j < m) out[(int64_t)i * m + j] = red[e];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int a = 0; a < NI; ++a)
#pragma unroll
                for (int b = 0; b < NJ; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        }
    }
}

struct GramDynPlan { int ni, nj, chunk_stages, nchunks, grid; };

static GramDynPlan gram_dyn_plan(int64_t m, int64_t k, int64_t n) {
    GramDynPlan p;
    auto frag = [](int64_t v) { return v <= 8 ? 1 : v <= 16 ? 2 : 4; };
    p.ni = frag(k); p.nj = frag(m);
    const int rows = 64 * gt_ksub(p.ni, p.nj);
    const int64_t total = (n + rows - 1) / rows;                      // stages in all
    p.grid = sm_count() * 2;
    if ((int64_t)p.grid > total / 4) p.grid = (int)(total / 4 > 0 ? total / 4 : 1);
    const int per_cta = g_knob[KNOB_GRAM_WAVES] > 0 ? g_knob[KNOB_GRAM_WAVES] : 8;    // chunks per CTA on average
    int64_t cs = total / ((int64_t)p.grid * per_cta);
    if (cs < 4) cs = 4;
    p.chunk_stages = (int)cs;
    p.nchunks = (int)((total + cs - 1) / cs);
    return p;
}

template <int NI, int NJ, bool SAME>
static int gram_dyn_launch(const GramDynPlan& p, const CUtensorMap& mo, const CUtensorMap& ms, int m, int k, int64_t n,
                           int* counter, double* part, cudaStream_t st) {
    constexpr int SMEM = gd_smem(NI, NJ);
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(gram_tma_dyn_kernel<NI, NJ, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    gram_tma_dyn_kernel<NI, NJ, SAME><<<p.grid, GT_THREADS, SMEM, st>>>(mo, ms, m, k, n, p.chunk_stages, p.nchunks, counter, part);
    return check_launch();
}

static int gram_tma_dyn(const double* S, int64_t lds, int64_t m, const double* O, int64_t ldo, int64_t k, int64_t n,
                        double* part, int* chunks_out, cudaStream_t st) {
    const GramDynPlan p = gram_dyn_plan(m, k, n);
    const bool same = S == O && lds == ldo && m == k;
    CUtensorMap mo, ms;
    int rc = make_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, O, n, k, ldo, 16, 8 * p.ni);
    if (!rc) rc = make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, S, n, m, lds, 16, 8 * p.nj);
    if (rc) return rc;
    // the chunk counter lives behind the partial tiles (gram_tma_ws_bytes reserves it)
    int* counter = reinterpret_cast<int*>(part + (size_t)p.nchunks * k * m);
    RL_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    *chunks_out = p.nchunks;
    const int im = (int)m, ik = (int)k;
#define RL_GD(NI_, NJ_)                                                                                   \
    if (p.ni == NI_ && p.nj == NJ_) {                                                                     \
        if constexpr (NI_ == NJ_) {                                                                       \
            if (same) return gram_dyn_launch<NI_, NJ_, true>(p, mo, ms, im, ik, n, counter, part, st);    \
        }                                                                                                 \
        return gram_dyn_launch<NI_, NJ_, false>(p, mo, ms, im, ik, n, counter, part, st);                 \
    }
    RL_GD(4, 4) RL_GD(4, 2) RL_GD(4, 1) RL_GD(2, 4) RL_GD(2, 2) RL_GD(2, 1) RL_GD(1, 4) RL_GD(1, 2) RL_GD(1, 1)
#undef RL_GD
    return RL_E_ARG;
}

bool gram_tma_ok(const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k, int64_t n) {
    return n >= 8192 && n < INT32_MAX && tma_encode_fn() != nullptr && host_aligned16(s) && host_aligned16(o) &&
           (lds % 2 == 0) && (ldo % 2 == 0);
}

size_t gram_tma_ws_bytes(int64_t m, int64_t k, int64_t n) {
    // for the wave knob in force now (callers size the workspace right before the call; a stale,
    // smaller workspace is refused with RL_E_WORKSPACE, never overrun)
    const int c1 = gram_tma_plan(m, k, n, 1).chunks, c2 = gram_tma_plan(m, k, n, 2).chunks;
    const int c3 = (m <= 32 && k <= 32) ? gram_dyn_plan(m, k, n).nchunks : 0;
    const int c = c1 > c2 ? (c1 > c3 ? c1 : c3) : (c2 > c3 ? c2 : c3);
    return (size_t)c * k * m * sizeof(double) + 256;                  // + the chunk counter of mode 3
}

template <int NI, int NJ, bool SAME, int STAGES, int MINB>
static int gram_tma_launch(const GramTmaPlan& p, const CUtensorMap& mo, const CUtensorMap& ms, int m, int k, int64_t n,
                           double* part, cudaStream_t st) {
    constexpr int SMEM = gt_smem(NI, NJ, STAGES);
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(gram_tma_kernel<NI, NJ, SAME, STAGES, MINB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    dim3 grid((unsigned)p.chunks, (unsigned)p.tiles_j, (unsigned)p.tiles_i);
    if (p.tiles_i * p.tiles_j > 1 && !g_knob[KNOB_GRAM_CHUNK_MAJOR])
        grid = dim3((unsigned)(p.chunks * p.tiles_j * p.tiles_i), 1, 1);
    gram_tma_kernel<NI, NJ, SAME, STAGES, MINB><<<grid, GT_THREADS, SMEM, st>>>(mo, ms, m, k, n, p.rows_per_cta, p.interleave, part);
    return check_launch();
}

template <int NI, int NJ>
static int gram_tma_dispatch(const GramTmaPlan& p, bool same, const CUtensorMap& mo, const CUtensorMap& ms, int m,
                             int k, int64_t n, double* part, cudaStream_t st) {
    if (p.minb == 2) {
        if constexpr (NI == NJ) {
            if (same) return gram_tma_launch<NI, NJ, true, 3, 2>(p, mo, ms, m, k, n, part, st);
        }
        return gram_tma_launch<NI, NJ, false, 3, 2>(p, mo, ms, m, k, n, part, st);
    }
    if constexpr (NI == NJ) {
        if (same) return gram_tma_launch<NI, NJ, true, 5, 1>(p, mo, ms, m, k, n, part, st);
    }
    return gram_tma_launch<NI, NJ, false, 5, 1>(p, mo, ms, m, k, n, part, st);
}

int gram_tma(const double* S, int64_t lds, int64_t m, const double* O, int64_t ldo, int64_t k, int64_t n, double* part,
             int* chunks_out, int mode, cudaStream_t st) {
    if ((mode & 3) == 3 && m <= 32 && k <= 32) return gram_tma_dyn(S, lds, m, O, ldo, k, n, part, chunks_out, st);
    const bool same = S == O && lds == ldo && m == k && m <= 32;
    const GramTmaPlan p = gram_tma_plan(m, k, n, mode, same);
    CUtensorMap mo, ms;
    int rc = make_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, O, n, k, ldo, 16, 8 * p.ni);
    if (!rc) rc = make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, S, n, m, lds, 16, 8 * p.nj);
    if (rc) return rc;
    *chunks_out = p.chunks;
    const int im = (int)m, ik = (int)k;
#define RL_GT(NI_, NJ_) if (p.ni == NI_ && p.nj == NJ_) return gram_tma_dispatch<NI_, NJ_>(p, same, mo, ms, im, ik, n, part, st)
    RL_GT(4, 4); RL_GT(4, 2); RL_GT(4, 1); RL_GT(2, 4); RL_GT(2, 2); RL_GT(2, 1); RL_GT(1, 4); RL_GT(1, 2); RL_GT(1, 1);
#undef RL_GT
    return RL_E_ARG;
}

}  // namespace rl
