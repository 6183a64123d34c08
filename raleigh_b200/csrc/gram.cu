// Tall-skinny Gram product  G[i,j] = sum_r O[i,r] * S[j,r]   (Vectors.dot,
// dense_cublas.py:245-269).  O is (k, n), S is (m, n), both vector-major, so the
// reduction index r is the contiguous one for both operands.
//
// Algorithmic traffic: (m + k) * n * w bytes in, k*m*w out  -> HBM-bound for
// m,k <= ~32 in fp64 (AI = m*k/(4(m+k)) flop/B), FP64-pipe-bound beyond.
//
// fp64 path: warp-level DMMA (mma.sync.m8n8k4.f64).  A warp owns an
// (8*NI) x (8*NJ) tile of G and a contiguous range of rows; per step of 16 rows
// each lane issues NI+NJ pairs of 128-bit loads straight from global memory
// (the fragment layout of m8n8k4 lets lane (g, c) read 4 consecutive rows
// 4c..4c+3 of vector g -- a full 128-byte line per vector per warp) and then
// NI*NJ*4 DMMAs.  No shared-memory staging is needed: every loaded element is
// used NI (or NJ) times from registers.
//
// fp32 path (and fp32-data/fp64-accumulate used by svd()): SIMT, lane = row,
// TI x TJ register tile, warp-shuffle reduction at the end.
//
// Determinism: each (chunk, tile) CTA writes its partial tile into a fixed slot
// of the workspace; a second kernel adds the chunk partials in a fixed order
// (strided per lane + xor-shuffle tree).  No floating-point atomics anywhere.
#include "common.cuh"

namespace rl {

constexpr int GRAM_WARPS = 4;
#ifndef RL_GRAM_TMA_DEFAULT
#define RL_GRAM_TMA_DEFAULT 2   // TMA ring, two CTAs per SM: +7..12 points of HBM peak on single-tile shapes (r1e sweep)
#endif
constexpr int GRAM_THREADS = GRAM_WARPS * 32;

struct GramPlan {
    int warps;           // warps per CTA of the DMMA kernel
    int ni, nj, js;      // 8-wide fragments per WARP tile along k (other) and m (self); warps side by side along m
    int tiles_i, tiles_j;
    int chunks;          // row chunks (CTAs along the reduction)
    int64_t rows_per_cta;
};

// A CTA (4 warps) owns an (8*NI) x (8*NJ*JS) tile of G; its 4/JS row-warps
// interleave 16-row steps so the CTA reads 4/JS*128 contiguous bytes of every
// vector per step.  Measured on B200 (profiles/r1c_kernel_tuning.md): the kernel is
// bound by the number of load requests, not by occupancy -- one warp per 32x32
// tile (NJ = 4, 248 registers, 8 warps/SM) with one 256-bit load per fragment
// beats two warps side by side (JS = 2, 16 warps/SM) 64% to 49% of HBM peak.
static int g_gram_prefetch = 0;  // L2 prefetch distance in steps (0 = off; measured: any distance 1..8 costs 35 %)
static int g_gram_warps8 = 0;
static int g_gram_minb = 0;      // debug: register cap of the 32x32 variant via min CTAs/SM (3 -> 168 regs, 4 -> 128)
static int g_gram_variant = 0;   // debug: 1 = two warps side by side per 32-wide tile (NJ = 2, JS = 2)

// gram_tma.cu
bool gram_tma_ok(const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k, int64_t n);
size_t gram_tma_ws_bytes(int64_t m, int64_t k, int64_t n);
int gram_tma(const double* S, int64_t lds, int64_t m, const double* O, int64_t ldo, int64_t k, int64_t n, double* part,
             int* chunks_out, int mode, cudaStream_t st);

static GramPlan gram_plan(int64_t m, int64_t k, int64_t n) {
    GramPlan p;
    auto frag = [](int64_t v) { return v <= 8 ? 1 : v <= 16 ? 2 : 4; };
    p.ni = frag(k);
    if (m > 16 && g_gram_variant == 1) { p.nj = 2; p.js = 2; } else { p.nj = frag(m); p.js = 1; }
    // 8 interleaved warps on the big tile: the CTA then reads 1 KB (instead of 512 B) of every
    // vector per step -- DRAM-page locality; same 8 warps per SM either way (248 registers)
    p.warps = (p.ni == 4 && p.nj == 4 && g_gram_warps8) ? 8 : GRAM_WARPS;
    p.tiles_i = (int)((k + 8 * p.ni - 1) / (8 * p.ni));
    p.tiles_j = (int)((m + 8 * p.nj * p.js - 1) / (8 * p.nj * p.js));
    int64_t tiles = (int64_t)p.tiles_i * p.tiles_j;
    const int64_t step = 16 * (p.warps / p.js);             // rows a CTA consumes per step
    // enough CTAs for ~6 per SM, but at least 4 steps per CTA
    int64_t want = ((int64_t)sm_count() * 6 + tiles - 1) / tiles;
    int64_t maxc = (n + 4 * step - 1) / (4 * step);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    int64_t rpc = (n + want - 1) / want;
    rpc = (rpc + step - 1) / step * step;
    p.rows_per_cta = rpc;
    p.chunks = (int)((n + rpc - 1) / rpc);
    return p;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Loads 4 consecutive rows r..r+3 of one vector for the current lane.
// ALIGNED: two 128-bit loads, caller guarantees r+3 < n and 16-byte alignment.
template <bool ALIGNED>
__device__ __forceinline__ void load4(const double* __restrict__ p, int64_t r, int64_t n, bool active,
                                      double (&v)[4]) {
    if (ALIGNED) {
        if (active) {
            // one 256-bit load (sm_100): the lane's whole 32-byte sector in a single request
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                         : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p + r));
        } else {
            v[0] = v[1] = v[2] = v[3] = 0.0;
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = (active && r + t < n) ? __ldg(p + r + t) : 0.0;
    }
}

template <int NI, int NJ, int JS, bool SAME, int MINB = 1, int DW = GRAM_WARPS>
__global__ void __launch_bounds__(DW * 32, MINB)
gram_dmma_kernel(const double* __restrict__ S, int64_t lds, int m, const double* __restrict__ O, int64_t ldo,
                 int k, int64_t n, int64_t rows_per_cta, int fast, double* __restrict__ part, int g_flags) {
    const int g_pf = g_flags & 255;
    const bool interleave = (g_flags >> 8) & 1;   // A/B: CTA c owns steps c, c + grid, ... (see gram_tma.cu)
    constexpr int RW = DW / JS;                      // warps interleaving row steps
    constexpr int TILE = JS * NI * NJ * 64;
    extern __shared__ double red_raw[];
    double (*red)[TILE] = reinterpret_cast<double (*)[TILE]>(red_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int jw = warp % JS, rw = warp / JS;
    const int g = lane >> 2, c = lane & 3;
    // multi-tile products: one-dimensional grid with the TILES of a row chunk adjacent in launch order, so that the
    // CTAs that read the same rows of S and O run at the same time and share them through L2 (chunk-major order
    // streams every vector from HBM once per tile that uses it: 4.3x the algorithmic traffic at m = k = 120)
    int bz = blockIdx.z, by = blockIdx.y, bx = blockIdx.x, nchunks = gridDim.x;
    if (gridDim.y == 1 && gridDim.z == 1) {
        const int tj = (m + 8 * NJ * JS - 1) / (8 * NJ * JS), ti = (k + 8 * NI - 1) / (8 * NI);
        const int t = bx % (tj * ti);
        bx /= tj * ti;
        nchunks /= tj * ti;
        by = t % tj;
        bz = t / tj;
    }
    const int i0 = bz * (8 * NI), j0 = by * (8 * NJ * JS) + jw * (8 * NJ);
    const int chunk = bx;

    const double* po[NI];
    const double* ps[NJ];
    bool ai[NI], aj[NJ];
#pragma unroll
    for (int t = 0; t < NI; ++t) { int i = i0 + 8 * t + g; ai[t] = i < k; po[t] = O + (int64_t)(ai[t] ? i : 0) * ldo; }
#pragma unroll
    for (int t = 0; t < NJ; ++t) { int j = j0 + 8 * t + g; aj[t] = j < m; ps[t] = S + (int64_t)(aj[t] ? j : 0) * lds; }

    double acc[NI][NJ][2];
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int64_t c_begin = interleave ? (int64_t)chunk * (16 * RW) : (int64_t)chunk * rows_per_cta;
    const int64_t c_end = interleave ? n : (c_begin + rows_per_cta < n ? c_begin + rows_per_cta : n);
    const int64_t r_step = interleave ? (int64_t)nchunks * (16 * RW) : (int64_t)(16 * RW);
    int64_t r = c_begin + rw * 16;
    // SAME: X.dot(X) with a single tile -- both operands are the same fragments, load once
    if (fast) {
        // full 16-row steps, one 256-bit load per fragment
        for (; r + 16 <= c_end; r += r_step) {
            double fa[NI][4], fb[NJ][4];
            if (g_pf > 0 && r + (int64_t)g_pf * r_step + 16 <= c_end) {
                // pull the fragments of a later step into L2 while this one computes
#pragma unroll
                for (int t = 0; t < NI; ++t)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(po[t] + r + (int64_t)g_pf * r_step + 4 * c));
                if (!SAME) {
#pragma unroll
                    for (int t = 0; t < NJ; ++t)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ps[t] + r + (int64_t)g_pf * r_step + 4 * c));
                }
            }
#pragma unroll
            for (int t = 0; t < NI; ++t) load4<true>(po[t], r + 4 * c, n, ai[t], fa[t]);
            if (SAME) {
#pragma unroll
                for (int t = 0; t < NJ; ++t)
#pragma unroll
                    for (int e = 0; e < 4; ++e) fb[t][e] = fa[t < NI ? t : 0][e];
            } else {
#pragma unroll
                for (int t = 0; t < NJ; ++t) load4<true>(ps[t], r + 4 * c, n, aj[t], fb[t]);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int a = 0; a < NI; ++a)
#pragma unroll
                    for (int b = 0; b < NJ; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a][s], fb[b][s]);
        }
    }
    for (; r < c_end; r += r_step) {   // ragged / unaligned steps
        double fa[NI][4], fb[NJ][4];
#pragma unroll
        for (int t = 0; t < NI; ++t) load4<false>(po[t], r + 4 * c, c_end, ai[t], fa[t]);
#pragma unroll
        for (int t = 0; t < NJ; ++t) load4<false>(ps[t], r + 4 * c, c_end, aj[t], fb[t]);
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int a = 0; a < NI; ++a)
#pragma unroll
                for (int b = 0; b < NJ; ++b) dmma(acc[a][b][0], acc[a][b][1], fa[a][s], fb[b][s]);
    }

    // C fragment: lane (g, c) holds rows g, cols 2c, 2c+1 of each 8x8 block
#pragma unroll
    for (int a = 0; a < NI; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b) {
            red[rw][((jw * NI + a) * NJ + b) * 64 + g * 8 + 2 * c] = acc[a][b][0];
            red[rw][((jw * NI + a) * NJ + b) * 64 + g * 8 + 2 * c + 1] = acc[a][b][1];
        }
    __syncthreads();
    double* out = part + (int64_t)chunk * k * m;
    const int jbase = by * (8 * NJ * JS);
    for (int e = threadIdx.x; e < TILE; e += DW * 32) {
        double v = red[0][e];
#pragma unroll
        for (int w = 1; w < RW; ++w) v += red[w][e];
        int blk = e >> 6;
        int b = blk % NJ, a = (blk / NJ) % NI, jj = blk / (NJ * NI);
        int i = i0 + 8 * a + ((e & 63) >> 3), j = jbase + jj * (8 * NJ) + 8 * b + (e & 7);
        if (i < k && j < m) out[(int64_t)i * m + j] = v;
    }
}

// ---- SIMT path (fp32, or fp32 data with fp64 accumulation) ---------------------
// lane = row; the warp owns a TI x TJ tile of G; per step every lane reads VR
// consecutive rows of TI + TJ vectors with one 128-bit load each.
template <typename T, typename TA, int TI, int TJ>
__global__ void __launch_bounds__(GRAM_THREADS)
gram_simt_kernel(const T* __restrict__ S, int64_t lds, int m, const T* __restrict__ O, int64_t ldo, int k,
                 int64_t n, int64_t rows_per_warp, int fast, TA* __restrict__ part) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    __shared__ TA red[GRAM_WARPS][TI * TJ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.z * TI, j0 = blockIdx.y * TJ;
    const int chunk = blockIdx.x;
    TA acc[TI][TJ];
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
        for (int b = 0; b < TJ; ++b) acc[a][b] = TA(0);

    int64_t r_begin = ((int64_t)chunk * GRAM_WARPS + warp) * rows_per_warp;
    int64_t r_end = r_begin + rows_per_warp < n ? r_begin + rows_per_warp : n;
    int64_t r = r_begin;
    if (fast) {
        for (; r + 32 * V <= r_end; r += 32 * V) {
            VT fa[TI], fb[TJ];
#pragma unroll
            for (int a = 0; a < TI; ++a) {
                int i = i0 + a;
                fa[a] = ldg_stream(reinterpret_cast<const VT*>(O + (int64_t)(i < k ? i : 0) * ldo + r) + lane);
            }
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                int j = j0 + b;
                fb[b] = ldg_stream(reinterpret_cast<const VT*>(S + (int64_t)(j < m ? j : 0) * lds + r) + lane);
            }
#pragma unroll
            for (int a = 0; a < TI; ++a)
#pragma unroll
                for (int b = 0; b < TJ; ++b) {
                    const T* ea = reinterpret_cast<const T*>(&fa[a]);
                    const T* eb = reinterpret_cast<const T*>(&fb[b]);
#pragma unroll
                    for (int t = 0; t < V; ++t) acc[a][b] = fma((TA)ea[t], (TA)eb[t], acc[a][b]);
                }
        }
    }
    for (int64_t rr = r + lane; rr < r_end; rr += 32) {
        T fa[TI], fb[TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a) { int i = i0 + a; fa[a] = i < k ? __ldg(O + (int64_t)i * ldo + rr) : T(0); }
#pragma unroll
        for (int b = 0; b < TJ; ++b) { int j = j0 + b; fb[b] = j < m ? __ldg(S + (int64_t)j * lds + rr) : T(0); }
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) acc[a][b] = fma((TA)fa[a], (TA)fb[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
        for (int b = 0; b < TJ; ++b) {
            TA v = warp_sum(acc[a][b]);
            if (lane == 0) red[warp][a * TJ + b] = v;
        }
    __syncthreads();
    TA* out = part + (int64_t)chunk * k * m;
    for (int e = threadIdx.x; e < TI * TJ; e += GRAM_THREADS) {
        TA v = red[0][e];
#pragma unroll
        for (int w = 1; w < GRAM_WARPS; ++w) v += red[w][e];
        int i = i0 + e / TJ, j = j0 + e % TJ;
        if (i < k && j < m) out[(int64_t)i * m + j] = v;
    }
}

// ---- second phase: fixed-order sum over the chunk partials ------------------------
// A block owns 32 consecutive entries of G; warp w adds chunks w, w+32, ... (coalesced 256-byte
// rows, 8 independent loads in flight per lane), then warp 0 adds the 32 warp sums in order.
// The order depends only on (chunks, km): same bits on every run and every rank.
constexpr int GRED_U = 8;
template <typename TA, typename TO>
__global__ void __launch_bounds__(1024) gram_reduce_kernel(const TA* __restrict__ part, int64_t km, int chunks,
                                                           TO* __restrict__ g) {
    __shared__ TA red[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane;
    TA acc[GRED_U];
#pragma unroll
    for (int u = 0; u < GRED_U; ++u) acc[u] = TA(0);
    if (e < km) {
        for (int c0 = warp; c0 < chunks; c0 += 32 * GRED_U) {
#pragma unroll
            for (int u = 0; u < GRED_U; ++u) {
                const int c = c0 + 32 * u;
                if (c < chunks) acc[u] += part[(int64_t)c * km + e];
            }
        }
    }
    red[warp][lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (warp == 0 && e < km) {
        TA s = red[0][lane];
#pragma unroll
        for (int w = 1; w < 32; ++w) s += red[w][lane];
        g[e] = (TO)s;
    }
}

static inline unsigned gram_reduce_blocks(int64_t km) { return (unsigned)((km + 31) / 32); }

template <int NI>
static int launch_dmma_nj(const GramPlan& p, const double* S, int64_t lds, int m, const double* O, int64_t ldo,
                          int k, int64_t n, int fast, double* part, cudaStream_t st) {
    dim3 grid((unsigned)p.chunks, (unsigned)p.tiles_j, (unsigned)p.tiles_i);
    if (p.tiles_i * p.tiles_j > 1 && !g_knob[KNOB_GRAM_CHUNK_MAJOR])
        grid = dim3((unsigned)(p.chunks * p.tiles_j * p.tiles_i), 1, 1);
    const bool same = S == O && lds == ldo && m == k && p.tiles_i == 1 && p.tiles_j == 1 && p.ni == p.nj && p.js == 1;
#define RL_GRAM_LAUNCH(NJ_, JS_, SAME_) \
    gram_dmma_kernel<NI, NJ_, JS_, SAME_><<<grid, GRAM_THREADS, (GRAM_WARPS / JS_) * JS_ * NI * NJ_ * 64 * sizeof(double), st>>>(S, lds, m, O, ldo, k, n, p.rows_per_cta, fast, part, g_gram_prefetch | ((g_knob[KNOB_GRAM_INTERLEAVE] & 1) << 8))
    if (same) {
        RL_GRAM_LAUNCH(NI, 1, true);
    } else if (p.nj == 4 && NI == 4 && p.warps == 8) {
        constexpr size_t smem8 = 8 * NI * 4 * 64 * sizeof(double);          // 64 KB of partial tiles
        static bool configured = false;
        if (!configured) {
            RL_CUDA(cudaFuncSetAttribute(gram_dmma_kernel<NI, 4, 1, false, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
            configured = true;
        }
        gram_dmma_kernel<NI, 4, 1, false, 1, 8><<<grid, 256, smem8, st>>>(S, lds, m, O, ldo, k, n, p.rows_per_cta, fast, part, g_gram_prefetch | ((g_knob[KNOB_GRAM_INTERLEAVE] & 1) << 8));
    } else if (p.nj == 4 && NI == 4 && g_gram_minb == 3) {
        gram_dmma_kernel<NI, 4, 1, false, 3><<<grid, GRAM_THREADS, 4 * NI * 4 * 64 * sizeof(double), st>>>(S, lds, m, O, ldo, k, n, p.rows_per_cta, fast, part, g_gram_prefetch | ((g_knob[KNOB_GRAM_INTERLEAVE] & 1) << 8));
    } else if (p.nj == 4) {
        RL_GRAM_LAUNCH(4, 1, false);
    } else if (p.js == 2) {
        RL_GRAM_LAUNCH(2, 2, false);
    } else if (p.nj == 1) {
        RL_GRAM_LAUNCH(1, 1, false);
    } else {
        RL_GRAM_LAUNCH(2, 1, false);
    }
#undef RL_GRAM_LAUNCH
    return check_launch();
}

// simt plan: 8x8 tiles
static GramPlan gram_plan_simt(int64_t m, int64_t k, int64_t n, int vec) {
    GramPlan p;
    p.ni = p.nj = p.js = 1;
    p.tiles_i = (int)((k + 7) / 8);
    p.tiles_j = (int)((m + 7) / 8);
    int64_t tiles = (int64_t)p.tiles_i * p.tiles_j;
    int64_t step = 32 * vec;
    int64_t want = ((int64_t)sm_count() * 8 + tiles - 1) / tiles;
    int64_t maxc = (n + GRAM_WARPS * step * 2 - 1) / (GRAM_WARPS * step * 2);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    int64_t rows_per_cta = (n + want - 1) / want;
    int64_t rpw = (rows_per_cta + GRAM_WARPS - 1) / GRAM_WARPS;
    rpw = (rpw + step - 1) / step * step;
    p.rows_per_cta = rpw;            // rows per WARP for the SIMT kernel
    p.chunks = (int)((n + rpw * GRAM_WARPS - 1) / (rpw * GRAM_WARPS));
    return p;
}

}  // namespace rl

using namespace rl;

extern "C" {

// gram_mode: 0 = dtype default (fp64 -> DMMA, fp32 -> SIMT fp32 accumulate),
//            1 = force SIMT, 2 = fp32 data with fp64 accumulation and fp64 output
static int g_gram_force_simt = 0;
void rl_debug_set_gram_simt(int on) { g_gram_force_simt = on & 1; g_gram_variant = (on >> 1) & 1; g_gram_prefetch = (on >> 8) & 255; g_gram_minb = (on >> 16) & 7; g_gram_warps8 = (on >> 20) & 1; }

static size_t gram_ws_bytes_impl(int dtype, int64_t m, int64_t k, int64_t n, int acc64) {
    if (m <= 0 || k <= 0 || n <= 0) return 0;
    if (dtype == RL_F64 && !g_gram_force_simt) {
        GramPlan p = gram_plan(m, k, n);
        size_t a = (size_t)p.chunks * k * m * sizeof(double), b = n >= 8192 ? gram_tma_ws_bytes(m, k, n) : 0;
        return a > b ? a : b;
    }
    GramPlan p = gram_plan_simt(m, k, n, dtype == RL_F32 ? 4 : 2);
    return (size_t)p.chunks * k * m * ((dtype == RL_F64 || acc64) ? 8 : 4);
}

size_t rl_gram_ws_bytes(int dtype, int64_t m, int64_t k, int64_t n) {
    return gram_ws_bytes_impl(dtype, m, k, n, 0);
}

static int gram_impl(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k,
                     int64_t n, void* g, void* ws, size_t ws_bytes, int acc64, cudaStream_t st) {
    if (m < 0 || k < 0 || n < 0 || m > INT32_MAX || k > INT32_MAX) return RL_E_ARG;
    if (m == 0 || k == 0) return 0;
    size_t esz = dtype == RL_F32 ? (acc64 ? 8 : 4) : 8;
    if (dtype != RL_F32 && dtype != RL_F64) return RL_E_DTYPE;
    if (n == 0) return (int)cudaMemsetAsync(g, 0, (size_t)k * m * esz, st);
    if (ws_bytes < gram_ws_bytes_impl(dtype, m, k, n, acc64)) return RL_E_WORKSPACE;
    int64_t km = k * m;
    const double wbytes = dtype == RL_F32 ? 4.0 : 8.0;
    Span span(PK_GRAM, st, ((s == o ? 1.0 : 2.0) * 0 + 1.0 * (s == o ? m : m + k)) * n * wbytes + km * esz,
              2.0 * n * m * k);
    int rc;
    int chunks;
    if (dtype == RL_F64 && !g_gram_force_simt) {
        // TMA-fed ring variant (gram_tma.cu): KNOB_GRAM_TMA 0 = default policy, -1 = off, 1/2 = forced mode
        const int tma_knob = g_knob[KNOB_GRAM_TMA];
        const int tma_mode = tma_knob > 0 ? tma_knob : (tma_knob == 0 ? RL_GRAM_TMA_DEFAULT : 0);
        // multi-tile products (m or k > 32) are DMMA-bound; from 64 vectors on the deep TMA ring (mode 1) feeds the
        // tensor pipe better than register fragments do (r2z, tools/time_gram120.py: 120 x 120 17.1 -> 20.1 TFLOP/s,
        // 96: 18.5 -> 22.1, 64: 18.2 -> 21.7; 48: no difference)
        const bool multi = m > 32 || k > 32;
        const bool multi_tma = multi && tma_knob == 0 && (m >= 64 || k >= 64) && m > 16 && k > 16;
        if (tma_mode > 0 && (tma_knob > 0 || !multi || multi_tma) && gram_tma_ok(s, lds, m, o, ldo, k, n)) {
            rc = gram_tma((const double*)s, lds, m, (const double*)o, ldo, k, n, (double*)ws, &chunks,
                          multi_tma ? 1 : tma_mode, st);
            if (rc) return rc;
            gram_reduce_kernel<double, double><<<gram_reduce_blocks(km), 1024, 0, st>>>((const double*)ws, km, chunks, (double*)g);
            return check_launch();
        }
        GramPlan p = gram_plan(m, k, n);
        int fast = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(o)) & 31) == 0 &&
                   (lds % 4 == 0) && (ldo % 4 == 0);
        const double* S = (const double*)s;
        const double* O = (const double*)o;
        switch (p.ni) {
            case 1: rc = launch_dmma_nj<1>(p, S, lds, (int)m, O, ldo, (int)k, n, fast, (double*)ws, st); break;
            case 2: rc = launch_dmma_nj<2>(p, S, lds, (int)m, O, ldo, (int)k, n, fast, (double*)ws, st); break;
            default: rc = launch_dmma_nj<4>(p, S, lds, (int)m, O, ldo, (int)k, n, fast, (double*)ws, st); break;
        }
        if (rc) return rc;
        chunks = p.chunks;
        gram_reduce_kernel<double, double><<<gram_reduce_blocks(km), 1024, 0, st>>>((const double*)ws, km, chunks, (double*)g);
        return check_launch();
    }
    GramPlan p = gram_plan_simt(m, k, n, dtype == RL_F32 ? 4 : 2);
    dim3 grid((unsigned)p.chunks, (unsigned)p.tiles_j, (unsigned)p.tiles_i);
    chunks = p.chunks;
    unsigned rblocks = gram_reduce_blocks(km);
    if (dtype == RL_F64) {
        int fast = host_aligned16(s) && host_aligned16(o) && (lds % 2 == 0) && (ldo % 2 == 0);
        gram_simt_kernel<double, double, 8, 8><<<grid, GRAM_THREADS, 0, st>>>((const double*)s, lds, (int)m, (const double*)o, ldo, (int)k, n, p.rows_per_cta, fast, (double*)ws);
        rc = check_launch(); if (rc) return rc;
        gram_reduce_kernel<double, double><<<rblocks, 1024, 0, st>>>((const double*)ws, km, chunks, (double*)g);
    } else if (acc64) {
        int fast = host_aligned16(s) && host_aligned16(o) && (lds % 4 == 0) && (ldo % 4 == 0);
        gram_simt_kernel<float, double, 8, 8><<<grid, GRAM_THREADS, 0, st>>>((const float*)s, lds, (int)m, (const float*)o, ldo, (int)k, n, p.rows_per_cta, fast, (double*)ws);
        rc = check_launch(); if (rc) return rc;
        gram_reduce_kernel<double, double><<<rblocks, 1024, 0, st>>>((const double*)ws, km, chunks, (double*)g);
    } else {
        int fast = host_aligned16(s) && host_aligned16(o) && (lds % 4 == 0) && (ldo % 4 == 0);
        gram_simt_kernel<float, float, 8, 8><<<grid, GRAM_THREADS, 0, st>>>((const float*)s, lds, (int)m, (const float*)o, ldo, (int)k, n, p.rows_per_cta, fast, (float*)ws);
        rc = check_launch(); if (rc) return rc;
        gram_reduce_kernel<float, float><<<rblocks, 1024, 0, st>>>((const float*)ws, km, chunks, (float*)g);
    }
    return check_launch();
}

int rl_gram(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k, int64_t n,
            void* g, void* ws, size_t ws_bytes, void* stream) {
    return gram_impl(dtype, s, lds, m, o, ldo, k, n, g, ws, ws_bytes, 0, as_stream(stream));
}

// fp32 (or fp64) data, fp64 accumulation, fp64 (k, m) result: the Gram used by
// the on-device SVD / orthonormalisation, where squaring the condition number
// in fp32 would lose the small singular values.
size_t rl_gram_acc64_ws_bytes(int dtype, int64_t m, int64_t k, int64_t n) {
    return gram_ws_bytes_impl(dtype, m, k, n, 1);
}
int rl_gram_acc64(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k,
                  int64_t n, double* g, void* ws, size_t ws_bytes, void* stream) {
    return gram_impl(dtype, s, lds, m, o, ldo, k, n, g, ws, ws_bytes, 1, as_stream(stream));
}

int rl_gram_h(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k, int64_t n,
              void* g_h, void* stream) {
    if (m < 0 || k < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || k == 0) return 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    size_t bytes = (size_t)k * m * w;
    void *pinned = nullptr, *dev = nullptr, *ws = nullptr;
    int rc = staging_acquire(bytes, &pinned, &dev);
    if (rc) return rc;
    size_t wsb = rl_gram_ws_bytes(dtype, m, k, n);
    if (wsb) { rc = scratch_acquire(wsb, &ws); if (rc) return rc; }
    rc = rl_gram(dtype, s, lds, m, o, ldo, k, n, dev, ws, wsb, stream);
    if (rc) return rc;
    RL_CUDA(cudaMemcpyAsync(pinned, dev, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    RL_CUDA(cudaStreamSynchronize(as_stream(stream)));
    memcpy(g_h, pinned, bytes);
    return 0;
}

}  // extern "C"
