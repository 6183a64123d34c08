// EXPERIMENTAL (knob 6, off by default, not measured yet -- written at the end of round 1 when the
// GPU budget was spent; profiles/r1e_gram_spmm.md has the evidence it responds to).
//
// Band-window block SpMM: the north-star "TMA staging of X tiles" for banded / stencil operators.
// ncu on the staged-CSR kernel: no unit saturated, L1/TEX busiest (62 %), the gathers wait on
// L2 / first-touch DRAM latency; 5 of the 7 gathers of a 3-D 7-point stencil (x+-1, y+-1, centre)
// fall within one 128-row tile of the row itself.
//
// Here a persistent CTA walks a contiguous range of 128-row tiles.  For the group of VG vectors it
// works on it keeps the X rows of tiles t-1, t, t+1 in shared memory (4 slots of VG x 128 values,
// slot = tile mod 4, filled two tiles ahead by one TMA box each: bulk, asynchronous, no registers,
// no tags), so every entry whose COLUMN lies in tiles t-1..t+1 is a fixed-latency LDS, conflict-free
// because lane = row; only entries outside the window (the +-N^2 neighbours) are gathered from
// global memory as before.  Classification is arithmetic (column >> 7 against the tile index): no
// set-up pass, any matrix is handled, a matrix without band structure simply takes the far path.
// Vector groups are spread over blockIdx.y; the CSR entries of a 32-row run are staged per group.
#include "common.cuh"
#include "tma.cuh"

namespace rl {

constexpr int SW_T = 128;            // rows per tile = 4 warps x 32 lanes
constexpr int SW_LOGT = 7;
constexpr int SW_SLOTS = 4;
constexpr int SW_WARPS = 4;

template <typename T, int VG>
__global__ void __launch_bounds__(SW_WARPS * 32, 4)
spmm_win_kernel(const __grid_constant__ CUtensorMap tmx, int64_t nrows, const int64_t* __restrict__ indptr,
                const int32_t* __restrict__ indices, const T* __restrict__ values, const T* __restrict__ X, int64_t ldx,
                T* __restrict__ Y, int64_t ldy, int m, int cap, int tiles_per_cta, int ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    T* win = reinterpret_cast<T*>(smem);                                       // [SW_SLOTS][VG][SW_T]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SW_SLOTS * VG * SW_T * sizeof(T));
    T* sval_all = reinterpret_cast<T*>(full + SW_SLOTS);
    int32_t* scol_all = reinterpret_cast<int32_t*>(sval_all + (size_t)SW_WARPS * cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* sval = sval_all + (size_t)warp * cap;
    int32_t* scol = scol_all + (size_t)warp * cap;

    const int v0 = blockIdx.y * VG;
    const int nv = m - v0 < VG ? m - v0 : VG;
    const int t0 = blockIdx.x * tiles_per_cta;
    const int t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
    if (t0 >= t1) return;
    const int u_first = t0 > 0 ? t0 - 1 : 0;                  // first tile this CTA ever loads
    const int u_last = t1 < ntiles ? t1 : ntiles - 1;         // last one (the "next" of its last tile)
    constexpr uint32_t TILE_BYTES = VG * SW_T * sizeof(T);

    if (threadIdx.x == 0) {
        for (int s = 0; s < SW_SLOTS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tile u -> slot u & 3; its (u - u_first) / 4-th use of that slot gives the phase parity
    auto issue = [&](int u) {
        mbar_expect_tx(&full[u & 3], TILE_BYTES);
        tma_load_2d(win + (size_t)(u & 3) * VG * SW_T, &tmx, &full[u & 3], u * SW_T, v0);
    };
    auto wait_tile = [&](int u) { mbar_wait(&full[u & 3], (uint32_t)(((u - u_first) >> 2) & 1)); };
    if (threadIdx.x == 0)
        for (int u = u_first; u <= u_last && u <= t0 + 1; ++u) issue(u);

    const T* xb[VG];
#pragma unroll
    for (int g = 0; g < VG; ++g) xb[g] = X + (int64_t)(v0 + (g < nv ? g : 0)) * ldx;

    for (int t = t0; t < t1; ++t) {
        // two tiles ahead: slot (t+2)&3 == (t-2)&3 was last read while computing tile t-1, and every
        // warp has passed the barrier that closes that iteration
        if (threadIdx.x == 0 && t + 2 <= u_last) issue(t + 2);
        if (t == t0) {
            for (int u = u_first; u <= t0; ++u) wait_tile(u);
        }
        if (t + 1 <= u_last) wait_tile(t + 1);

        const int64_t row0 = (int64_t)t * SW_T + warp * 32;
        if (row0 < nrows) {
            const int64_t r = row0 + lane;
            const bool live = r < nrows;
            const int64_t p0 = live ? __ldg(indptr + r) : 0;
            const int64_t p1 = live ? __ldg(indptr + r + 1) : 0;
            const int64_t base = __shfl_sync(0xffffffffu, p0, 0);
            const int64_t last_row = (row0 + 32 <= nrows ? row0 + 32 : nrows);
            const int64_t cnt = __ldg(indptr + last_row) - base;
            const bool staged = cnt <= cap;
            if (staged) {
                for (int64_t e = lane; e < cnt; e += 32) {
                    scol[e] = ldg_stream(indices + base + e);
                    sval[e] = ldg_stream(values + base + e);
                }
                __syncwarp();
            }
            T acc[VG];
#pragma unroll
            for (int g = 0; g < VG; ++g) acc[g] = T(0);
            // one entry: window hit -> VG shared loads, else VG global gathers
            auto fetch = [&](int c, T (&x)[VG]) {
                const int ct = c >> SW_LOGT;
                if ((unsigned)(ct - t + 1) <= 2u) {
                    const T* w = win + (size_t)(ct & 3) * VG * SW_T + (c & (SW_T - 1));
#pragma unroll
                    for (int g = 0; g < VG; ++g) x[g] = w[g * SW_T];
                } else {
                    const int64_t off = (int64_t)c * (int64_t)sizeof(T);
#pragma unroll
                    for (int g = 0; g < VG; ++g)
                        x[g] = __ldg(reinterpret_cast<const T*>(reinterpret_cast<const char*>(xb[g]) + off));
                }
            };
            if (staged) {
                const int q0 = (int)(p0 - base), q1 = (int)(p1 - base);
                int p = q0;
#pragma unroll 1
                for (; p + 2 <= q1; p += 2) {
                    const int c0 = scol[p], c1 = scol[p + 1];
                    const T a0 = sval[p], a1 = sval[p + 1];
                    T x0[VG], x1[VG];
                    fetch(c0, x0);
                    fetch(c1, x1);
#pragma unroll
                    for (int g = 0; g < VG; ++g) { acc[g] = fma(a0, x0[g], acc[g]); acc[g] = fma(a1, x1[g], acc[g]); }
                }
                if (p < q1) {
                    T x0[VG];
                    fetch(scol[p], x0);
                    const T a0 = sval[p];
#pragma unroll
                    for (int g = 0; g < VG; ++g) acc[g] = fma(a0, x0[g], acc[g]);
                }
            } else {
                for (int64_t p = p0; p < p1; ++p) {
                    T x0[VG];
                    fetch(__ldg(indices + p), x0);
                    const T a0 = __ldg(values + p);
#pragma unroll
                    for (int g = 0; g < VG; ++g) acc[g] = fma(a0, x0[g], acc[g]);
                }
            }
            if (live) {
#pragma unroll
                for (int g = 0; g < VG; ++g)
                    if (g < nv) Y[(int64_t)(v0 + g) * ldy + r] = acc[g];
            }
        }
        __syncthreads();        // tile t is done by every warp: its oldest slot may be refilled
    }
}

bool spmm_win_ok(int dtype, const void* x, int64_t ldx, int64_t nrows, int64_t m) {
    const int64_t w = dtype == RL_F32 ? 4 : 8;
    return tma_encode_fn() != nullptr && host_aligned16(x) && (ldx * w) % 16 == 0 && nrows >= 4 * SW_T &&
           nrows < INT32_MAX - 4 * SW_T && m >= 1;
}

template <typename T>
static int spmm_win_impl(int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const void* values,
                         const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, cudaStream_t st) {
    constexpr int VG = 8;
    int64_t avg = nrows > 0 ? (nnz * 32 + nrows - 1) / nrows : 0;
    int64_t want = (avg * 5 / 4 + 63) / 64 * 64;
    const int cap = (int)(want < 256 ? 256 : want > 2048 ? 2048 : want);
    const size_t smem = (size_t)SW_SLOTS * VG * SW_T * sizeof(T) + SW_SLOTS * 8 + (size_t)SW_WARPS * cap * (sizeof(T) + 4) + 128;
    auto kern = spmm_win_kernel<T, VG>;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    CUtensorMap tmx;
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    int rc = make_map(&tmx, dt, (int)sizeof(T), x, nrows, m, ldx, SW_T, VG, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    const int ntiles = (int)((nrows + SW_T - 1) / SW_T);
    const int groups = (int)((m + VG - 1) / VG);
    // ~4 resident CTAs per SM and a few chunks per slot so that the block scheduler evens out SM speeds
    const int per_slot = g_knob[KNOB_SPMM_WIN_CHUNKS] > 0 ? g_knob[KNOB_SPMM_WIN_CHUNKS] : 4;
    int64_t ctas = ((int64_t)sm_count() * 4 * per_slot + groups - 1) / groups;
    if (ctas > ntiles / 4) ctas = ntiles / 4 > 0 ? ntiles / 4 : 1;
    const int tiles_per_cta = (int)((ntiles + ctas - 1) / ctas);
    const unsigned gx = (unsigned)((ntiles + tiles_per_cta - 1) / tiles_per_cta);
    kern<<<dim3(gx, (unsigned)groups), SW_WARPS * 32, smem, st>>>(tmx, nrows, indptr, indices, (const T*)values, (const T*)x, ldx,
                                                                  (T*)y, ldy, (int)m, cap, tiles_per_cta, ntiles);
    return check_launch();
}

int spmm_win(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const void* values,
             const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, cudaStream_t st) {
    if (dtype == RL_F32) return spmm_win_impl<float>(nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, st);
    if (dtype == RL_F64) return spmm_win_impl<double>(nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, st);
    return RL_E_DTYPE;
}

}  // namespace rl
