// CSR block SpMM  Y[v, r] = sum_{p in row r} val[p] * X[v, col[p]]
// (SparseSymmetricMatrix.apply, sparse_mkl.py:42-48 -> mkl_?csrmm, mkl_wrap.py:274-276).
// The device holds the full symmetric matrix (both triangles) as 0-based CSR.
//
// Algorithmic traffic: nnz*(w+4) + (nrows+1)*8 + 2*nrows*m*w bytes, 2*nnz*m flops.
//
// Mapping for VECTOR-MAJOR block vectors: lane = row, so for stencil / banded
// matrices the 32 lanes of a warp gather 32 neighbouring components of one
// vector -- a coalesced 128/256-byte request -- and the Y store is coalesced by
// construction.  A warp stages the (col, val) entries of its 32 rows into shared
// memory ONCE with coalesced loads and re-reads them from there for every group
// of VG vectors, so the matrix is streamed from HBM exactly once per SpMM no
// matter how many vectors the block has.  The X gathers go through L1/L2: the
// reuse distance of a stencil (one grid plane) is L2-resident; what they wait for
// is lines still in flight from DRAM, hence the L2 prefetch below.
// Measurements and ncu readings behind every choice here: profiles/r1e_gram_spmm.md.
#include "common.cuh"

namespace rl {

#ifndef RL_SPMM_PF_DEFAULT
#define RL_SPMM_PF_DEFAULT 1      // bit 0 measured: 0.381 -> 0.330 ms on the 128^3 Laplacian (m = 32)
#endif
// Warps (32-row runs) per CTA: 4.  8- and 16-warp CTAs were slower with consecutive runs (r1c) and with
// footprint-clustered runs (r1e: 0.3625 / 0.381 / 0.479 ms for 4 / 8 / 16).
static int g_spmm_warps = 4;

// Column c of the operator: owned columns come from the local block X (vector-major),
// halo columns (c >= ncols_local, row-sharded operator) from the exchanged halo buffer H,
// which is row-interleaved: H[(c - ncols_local) * m + v].
//
// Gather addressing: xb[g] = X + (v0 + g) * ldx is hoisted out of the entry loop, so a gather is
// one IMAD.WIDE (xb[g] + 8c) + one LDG.  (ncu, r1e: with the address of every (vector, column)
// pair rebuilt from scratch and both the owned and the halo load issued under predicates, the
// kernel was ISSUE-bound -- 58 % issue slots, 34 % L2, 38 % DRAM -- at ~200 instructions per
// two entries x 8 vectors instead of ~55.)
template <typename T, int VG, bool HALO>
struct Gather {
    const T* xb[VG];
    const T* hb;
    int ncl, m;
    // pointers of vectors 0..VG-1 (surplus ones alias vector 0: their results are dropped)
    __device__ __forceinline__ void init(const T* __restrict__ X, int64_t ldx, const T* __restrict__ H, int m_,
                                         int ncols_local, int v0, int nv) {
#pragma unroll
        for (int g = 0; g < VG; ++g) xb[g] = X + (int64_t)(v0 + (g < nv ? g : 0)) * ldx;
        hb = HALO ? H + v0 : nullptr;
        ncl = ncols_local;
        m = m_;
    }
    // next group of VG vectors (the last, ragged group re-points its surplus entries)
    __device__ __forceinline__ void advance(int64_t ldx, int nv_next) {
#pragma unroll
        for (int g = 0; g < VG; ++g) xb[g] += (g < nv_next ? (int64_t)VG * ldx : (int64_t)(VG - g) * ldx);
        if (HALO) hb += VG;
    }
    __device__ __forceinline__ void load(int c, T (&x)[VG]) const {
        if (HALO && c >= ncl) {
            const T* h = hb + (int64_t)(c - ncl) * m;     // v0 .. v0+VG-1 are contiguous here
#pragma unroll
            for (int g = 0; g < VG; ++g) x[g] = __ldg(h + g);
        } else {
            // one 64-bit byte offset per entry, one 64-bit add per gather; plain __ldg so that the
            // compiler keeps hoisting all 2*VG loads above the FMAs (inline-asm loads were sunk next
            // to their uses: serialised gathers, 0.40 -> 0.58 ms on the 128^3 Laplacian)
            const int64_t off = (int64_t)c * (int64_t)sizeof(T);
#pragma unroll
            for (int g = 0; g < VG; ++g)
                x[g] = __ldg(reinterpret_cast<const T*>(reinterpret_cast<const char*>(xb[g]) + off));
        }
    }
};

// WPS = resident warps per SM the register allocation is held to (24 -> ~80 registers, 16 -> ~127):
// with fewer registers ptxas sinks the gathers next to their FMAs and serialises them.
template <typename T, int VG, int SPMM_WARPS, bool HALO, int WPS, int NB>
__global__ void __launch_bounds__(SPMM_WARPS * 32, WPS / SPMM_WARPS)
spmm_kernel(int64_t nrows, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
            const T* __restrict__ values, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y, int64_t ldy,
            int m, int cap, int ncols_local, const T* __restrict__ H, const int32_t* __restrict__ run_order,
            int pf_mode, int pf_dist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* sval = reinterpret_cast<T*>(smem_raw) + (size_t)warp * cap;
    int32_t* scol = reinterpret_cast<int32_t*>(reinterpret_cast<T*>(smem_raw) + (size_t)SPMM_WARPS * cap) + (size_t)warp * cap;

    // slot -> 32-row run: identity, or the footprint-clustered order built at set-up
    // (rl_spmm_cluster_runs): the SPMM_WARPS runs of a CTA then gather from overlapping
    // column segments, so a line fetched for one warp is an L1 hit for the others
    const int64_t slot = (int64_t)blockIdx.x * SPMM_WARPS + warp;
    if (slot * 32 >= nrows) return;
    const int64_t row0 = (run_order ? (int64_t)__ldg(run_order + slot) : slot) * 32;
    const int64_t r = row0 + lane;
    const bool live = r < nrows;
    const int64_t p0 = live ? __ldg(indptr + r) : 0;
    const int64_t p1 = live ? __ldg(indptr + r + 1) : 0;
    const int64_t base = __shfl_sync(0xffffffffu, p0, 0);
    const int64_t last_row = (row0 + 32 <= nrows ? row0 + 32 : nrows);
    const int64_t end = __ldg(indptr + last_row);
    const int64_t cnt = end - base;
    const bool staged = cnt <= cap;
    if (staged) {
        for (int64_t e = lane; e < cnt; e += 32) {
            scol[e] = ldg_stream(indices + base + e);      // read once: keep L1 for the X gathers
            sval[e] = ldg_stream(values + base + e);
        }
        __syncwarp();
    }
    const int q0 = (int)(p0 - base), q1 = (int)(p1 - base);   // valid when staged

    // L2 prefetch (ncu, r1e: half of the gathers' L2 lookups missed although DRAM reads X once --
    // they waited for lines still in flight from DRAM).  Rows are processed in increasing order, so
    //  bit 0: the LARGEST column of a row (the +N^2 neighbour of a stencil) is usually a first touch:
    //         pull those lines of every vector into L2 now, one warp-wide request per vector;
    //  bit 1: the CSR entries of the run that will start when this one retires (pf_dist slots ahead)
    //         are a cold DRAM read in front of two dependent round trips: pull them in as well;
    //  bit 2: also the same X lines shifted by pf_dist runs (banded matrices: column ~ row + const).
    if (pf_mode && staged && live && q1 > q0) {
        const int cl = scol[q1 - 1];
        if ((pf_mode & 1) && (!HALO || cl < ncols_local)) {
            const T* px = X + cl;
            for (int v = 0; v < m; ++v) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + (int64_t)v * ldx));
        }
        if (pf_mode & 4) {
            const int64_t cf = (int64_t)cl + (int64_t)pf_dist * 32;
            if (cf < (HALO ? (int64_t)ncols_local : nrows)) {
                const T* px = X + cf;
                for (int v = 0; v < m; ++v) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + (int64_t)v * ldx));
            }
        }
    }
    if ((pf_mode & 2) && run_order == nullptr) {
        const int64_t fr = (slot + pf_dist) * 32;
        if (fr + 32 <= nrows) {
            const int64_t f0 = __ldg(indptr + fr), f1 = __ldg(indptr + fr + 32);
            for (int64_t e = f0 + lane * 32; e < f1; e += 1024) asm volatile("prefetch.global.L2 [%0];" ::"l"(indices + e));
            for (int64_t e = f0 + lane * 16; e < f1; e += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(values + e));
        }
    }

    Gather<T, VG, HALO> ga;                         // loop-carried pointers: one per vector of the group
    ga.init(X, ldx, H, m, ncols_local, 0, m < VG ? m : VG);
    for (int v0 = 0; v0 < m; v0 += VG) {
        T acc[VG];
#pragma unroll
        for (int g = 0; g < VG; ++g) acc[g] = T(0);
        const int nv = m - v0 < VG ? m - v0 : VG;
        if (HALO && nv < VG) {
            // the halo buffer has only m values per column: stay scalar for the ragged last group
            for (int64_t p = staged ? q0 : p0; p < (staged ? q1 : p1); ++p) {
                const int c0 = staged ? scol[p] : __ldg(indices + p);
                const T a0 = staged ? sval[p] : __ldg(values + p);
#pragma unroll
                for (int g = 0; g < VG; ++g) {
                    if (g < nv) {
                        const T xv = c0 < ncols_local ? __ldg(ga.xb[g] + c0)
                                                      : __ldg(H + (int64_t)(c0 - ncols_local) * m + v0 + g);
                        acc[g] = fma(a0, xv, acc[g]);
                    }
                }
            }
        } else if (staged) {
            int p = q0;
            if (NB == 4) {
                // four entries per step: 4 * VG gathers in flight per lane
#pragma unroll 1
                for (; p + 4 <= q1; p += 4) {
                    T x0[VG], x1[VG], x2[VG], x3[VG];
                    const int c0 = scol[p], c1 = scol[p + 1], c2 = scol[p + 2], c3 = scol[p + 3];
                    ga.load(c0, x0);
                    ga.load(c1, x1);
                    ga.load(c2, x2);
                    ga.load(c3, x3);
                    const T a0 = sval[p], a1 = sval[p + 1], a2 = sval[p + 2], a3 = sval[p + 3];
#pragma unroll
                    for (int g = 0; g < VG; ++g) {
                        acc[g] = fma(a0, x0[g], acc[g]); acc[g] = fma(a1, x1[g], acc[g]);
                        acc[g] = fma(a2, x2[g], acc[g]); acc[g] = fma(a3, x3[g], acc[g]);
                    }
                }
            }
#pragma unroll 1
            for (; p + 2 <= q1; p += 2) {
                const int c0 = scol[p], c1 = scol[p + 1];
                const T a0 = sval[p], a1 = sval[p + 1];
                T x0[VG], x1[VG];
                ga.load(c0, x0);
                ga.load(c1, x1);
#pragma unroll
                for (int g = 0; g < VG; ++g) { acc[g] = fma(a0, x0[g], acc[g]); acc[g] = fma(a1, x1[g], acc[g]); }
            }
            if (p < q1) {
                const int c0 = scol[p];
                const T a0 = sval[p];
                T x0[VG];
                ga.load(c0, x0);
#pragma unroll
                for (int g = 0; g < VG; ++g) acc[g] = fma(a0, x0[g], acc[g]);
            }
        } else {
            for (int64_t p = p0; p < p1; ++p) {
                const int c0 = __ldg(indices + p);
                const T a0 = __ldg(values + p);
                T x0[VG];
                ga.load(c0, x0);
#pragma unroll
                for (int g = 0; g < VG; ++g) acc[g] = fma(a0, x0[g], acc[g]);
            }
        }
        if (live) {
#pragma unroll
            for (int g = 0; g < VG; ++g)
                if (g < nv) Y[(int64_t)(v0 + g) * ldy + r] = acc[g];
        }
        const int left = m - v0 - VG;
        ga.advance(ldx, left < VG ? (left > 0 ? left : 0) : VG);
    }
}

template <typename T, int W, bool HALO, int WPS, int NB>
static int spmm_launch(int64_t nrows, const int64_t* indptr, const int32_t* indices, const T* values, const T* x,
                       int64_t ldx, T* y, int64_t ldy, int m, int cap, int ncols_local, const T* halo,
                       const int32_t* run_order, cudaStream_t st) {
    constexpr int VG = 8;
    size_t smem = (size_t)W * cap * (sizeof(T) + 4);
    auto kern = spmm_kernel<T, VG, W, HALO, WPS, NB>;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // A/B knob: shared-memory carve-out in percent (the rest of the 256 KB is L1 for the X gathers)
    static int carveout = 0;
    if (g_knob[KNOB_SPMM_CARVEOUT] != carveout) {
        carveout = g_knob[KNOB_SPMM_CARVEOUT];
        RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     carveout > 0 ? carveout : cudaSharedmemCarveoutDefault));
    }
    int64_t blocks = (nrows + W * 32 - 1) / (W * 32);
    kern<<<(unsigned)blocks, W * 32, smem, st>>>(nrows, indptr, indices, values, x, ldx, y, ldy, m, cap, ncols_local,
                                                 halo, run_order,
                                                 g_knob[KNOB_SPMM_PREFETCH] > 0 ? g_knob[KNOB_SPMM_PREFETCH] : (g_knob[KNOB_SPMM_PREFETCH] < 0 ? 0 : RL_SPMM_PF_DEFAULT),
                                                 sm_count() * WPS);
    return check_launch();
}

template <typename T>
static int spmm_impl(int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices, const void* values,
                     const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int ncols_local, const void* halo,
                     const int32_t* run_order, int warps, cudaStream_t st) {
    // shared-memory capacity per warp: the average entries of 32 rows + 25 %, in [256, 4096];
    // warps whose segment is longer read the matrix from global memory instead
    int64_t avg = nrows > 0 ? (nnz * 32 + nrows - 1) / nrows : 0;
    int64_t want = (avg * 5 / 4 + 63) / 64 * 64;            // 25 % head room over the average 32-row segment
    int cap = (int)(want < 256 ? 256 : want > 4096 ? 4096 : want);

    if (warps <= 0) warps = g_spmm_warps;
    // big CTAs only when the staged entries of all their warps leave most of L1 free
    if (warps >= 8 && (size_t)warps * cap * (sizeof(T) + 4) > 96 * 1024) warps = 4;
    // long rows (>= 16 entries on average): four entries x 8 vectors in flight per lane and the
    // 127-register budget win (55 nnz/row: 0.24 -> 0.16 ms); stencils keep 24 warps per SM
    const bool fat = g_knob[KNOB_SPMM_WPS] == 16 || (g_knob[KNOB_SPMM_WPS] == 0 && nnz >= 16 * nrows);
#define RL_SPMM_ARGS nrows, indptr, indices, (const T*)values, (const T*)x, ldx, (T*)y, ldy, (int)m, cap, ncols_local
#define RL_SPMM_W(W_) do { \
        if (halo) return spmm_launch<T, W_, true, 24, 2>(RL_SPMM_ARGS, (const T*)halo, run_order, st); \
        if (fat) return spmm_launch<T, W_, false, 16, 4>(RL_SPMM_ARGS, nullptr, run_order, st); \
        return spmm_launch<T, W_, false, 24, 2>(RL_SPMM_ARGS, nullptr, run_order, st); } while (0)
    if (warps >= 16) RL_SPMM_W(16);
    if (warps >= 8) RL_SPMM_W(8);
    RL_SPMM_W(4);
#undef RL_SPMM_W
#undef RL_SPMM_ARGS
}

// ---- SELL-32 variant ---------------------------------------------------------------
// Sliced ELLPACK with 32-row slices: entry j of row 32*s + l sits at
// slice_ptr[s] + 32*j + l, so the (col, val) reads of a warp are perfectly
// coalesced without any staging, the matrix is streamed exactly once per pass
// and nothing but registers is used -> high occupancy.  A thread keeps up to MT
// accumulators (one per vector), so for m <= 32 a single pass over the matrix
// serves the whole block.  Padding entries carry val = 0 and a valid column.
template <typename T, int MT, bool HALO>
__global__ void __launch_bounds__(128)
sell_spmm_kernel(int64_t nrows, int64_t nslices, const int64_t* __restrict__ slice_ptr,
                 const int32_t* __restrict__ cols, const T* __restrict__ vals, const T* __restrict__ X, int64_t ldx,
                 T* __restrict__ Y, int64_t ldy, int v0, int nv, int m, int ncols_local, const T* __restrict__ H) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nslices) return;
    const int64_t r = s * 32 + lane;
    const int64_t b0 = __ldg(slice_ptr + s), b1 = __ldg(slice_ptr + s + 1);
    T acc[MT];
#pragma unroll
    for (int g = 0; g < MT; ++g) acc[g] = T(0);
    int64_t p = b0 + lane;
    Gather<T, MT, HALO> ga;
    ga.init(X, ldx, H, m, ncols_local, v0, nv);
    if (!HALO || nv == MT) {
        for (; p + 32 < b1; p += 64) {
            const int c0 = ldg_stream(cols + p), c1 = ldg_stream(cols + p + 32);
            const T a0 = ldg_stream(vals + p), a1 = ldg_stream(vals + p + 32);
            T x0[MT], x1[MT];
            ga.load(c0, x0);
            ga.load(c1, x1);
#pragma unroll
            for (int g = 0; g < MT; ++g) { acc[g] = fma(a0, x0[g], acc[g]); acc[g] = fma(a1, x1[g], acc[g]); }
        }
        if (p < b1) {
            const int c0 = ldg_stream(cols + p);
            const T a0 = ldg_stream(vals + p);
            T x0[MT];
            ga.load(c0, x0);
#pragma unroll
            for (int g = 0; g < MT; ++g) acc[g] = fma(a0, x0[g], acc[g]);
        }
    } else {
        // ragged last group of a sharded operator: the halo buffer holds only m values per column
        for (; p < b1; p += 32) {
            const int c0 = __ldg(cols + p);
            const T a0 = __ldg(vals + p);
#pragma unroll
            for (int g = 0; g < MT; ++g) {
                if (g < nv) {
                    const T xv = c0 < ncols_local ? __ldg(ga.xb[g] + c0) : __ldg(H + (int64_t)(c0 - ncols_local) * m + v0 + g);
                    acc[g] = fma(a0, xv, acc[g]);
                }
            }
        }
    }
    if (r < nrows) {
#pragma unroll
        for (int g = 0; g < MT; ++g)
            if (g < nv) Y[(int64_t)(v0 + g) * ldy + r] = acc[g];
    }
}

template <typename T, bool HALO>
static int sell_impl_h(int64_t nrows, int64_t nslices, const int64_t* slice_ptr, const int32_t* cols, const void* vals,
                       const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int ncols_local, const void* halo,
                       cudaStream_t st) {
    const unsigned blocks = (unsigned)((nslices + 3) / 4);
    for (int64_t v0 = 0; v0 < m;) {
        const int64_t left = m - v0;
        int rc;
        if (left > 16) {
            const int nv = left < 32 ? (int)left : 32;
            sell_spmm_kernel<T, 32, HALO><<<blocks, 128, 0, st>>>(nrows, nslices, slice_ptr, cols, (const T*)vals, (const T*)x, ldx, (T*)y, ldy, (int)v0, nv, (int)m, ncols_local, (const T*)halo);
            v0 += nv;
        } else if (left > 8) {
            sell_spmm_kernel<T, 16, HALO><<<blocks, 128, 0, st>>>(nrows, nslices, slice_ptr, cols, (const T*)vals, (const T*)x, ldx, (T*)y, ldy, (int)v0, (int)left, (int)m, ncols_local, (const T*)halo);
            v0 += left;
        } else {
            sell_spmm_kernel<T, 8, HALO><<<blocks, 128, 0, st>>>(nrows, nslices, slice_ptr, cols, (const T*)vals, (const T*)x, ldx, (T*)y, ldy, (int)v0, (int)left, (int)m, ncols_local, (const T*)halo);
            v0 += left;
        }
        rc = check_launch();
        if (rc) return rc;
    }
    return 0;
}

template <typename T>
static int sell_impl(int64_t nrows, int64_t nslices, const int64_t* slice_ptr, const int32_t* cols, const void* vals,
                     const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, int ncols_local, const void* halo,
                     cudaStream_t st) {
    return halo ? sell_impl_h<T, true>(nrows, nslices, slice_ptr, cols, vals, x, ldx, y, ldy, m, ncols_local, halo, st)
                : sell_impl_h<T, false>(nrows, nslices, slice_ptr, cols, vals, x, ldx, y, ldy, m, ncols_local, nullptr, st);
}

// Halo packing for the row-sharded operator: out[t*m + v] = X[v, idx[t]]
// (row-interleaved, so that the rows requested by one peer are contiguous).
template <typename T>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const T* __restrict__ X, int64_t ldx, int m, const int64_t* __restrict__ idx, int64_t count,
                 T* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count * m) return;
    const int64_t t = e / m;
    const int v = (int)(e - t * m);
    out[e] = __ldg(X + (int64_t)v * ldx + __ldg(idx + t));
}

}  // namespace rl

using namespace rl;

extern "C" {

void rl_debug_set_spmm_warps(int warps) { g_spmm_warps = warps; }

int rl_pack_rows(int dtype, const void* x, int64_t ldx, int64_t m, const int64_t* idx, int64_t count, void* out,
                 void* stream) {
    if (m < 0 || count < 0 || m > INT32_MAX) return RL_E_ARG;
    if (m == 0 || count == 0) return 0;
    const unsigned blocks = (unsigned)((count * m + 255) / 256);
    if (dtype == RL_F32) pack_rows_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>((const float*)x, ldx, (int)m, idx, count, (float*)out);
    else if (dtype == RL_F64) pack_rows_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>((const double*)x, ldx, (int)m, idx, count, (double*)out);
    else return RL_E_DTYPE;
    return check_launch();
}

int rl_sell_spmm_halo(int dtype, int64_t nrows, int64_t nnz, int64_t nslices, const int64_t* slice_ptr,
                      const int32_t* cols, const void* vals, const void* x, int64_t ldx, void* y, int64_t ldy,
                      int64_t m, int64_t ncols_local, const void* halo, void* stream);

int rl_sell_spmm(int dtype, int64_t nrows, int64_t nnz, int64_t nslices, const int64_t* slice_ptr,
                 const int32_t* cols, const void* vals, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m,
                 void* stream) {
    return rl_sell_spmm_halo(dtype, nrows, nnz, nslices, slice_ptr, cols, vals, x, ldx, y, ldy, m, 0, nullptr, stream);
}

int rl_sell_spmm_halo(int dtype, int64_t nrows, int64_t nnz, int64_t nslices, const int64_t* slice_ptr,
                      const int32_t* cols, const void* vals, const void* x, int64_t ldx, void* y, int64_t ldy,
                      int64_t m, int64_t ncols_local, const void* halo, void* stream) {
    if (nrows < 0 || m < 0 || nnz < 0 || nslices < 0 || m > INT32_MAX || ncols_local > INT32_MAX) return RL_E_ARG;
    if (nrows == 0 || m == 0) return 0;
    if (x == y) return RL_E_ALIAS;
    const double w = dtype == RL_F32 ? 4.0 : 8.0;
    // algorithmic traffic is that of the unpadded matrix
    Span span(PK_SPMM, as_stream(stream), nnz * (w + 4.0) + (nrows + 1) * 8.0 + 2.0 * nrows * m * w,
              2.0 * nnz * m);
    if (dtype == RL_F32) return sell_impl<float>(nrows, nslices, slice_ptr, cols, vals, x, ldx, y, ldy, m, (int)ncols_local, halo, as_stream(stream));
    if (dtype == RL_F64) return sell_impl<double>(nrows, nslices, slice_ptr, cols, vals, x, ldx, y, ldy, m, (int)ncols_local, halo, as_stream(stream));
    return RL_E_DTYPE;
}

int rl_csr_spmm_ex(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                   const void* values, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m,
                   int64_t ncols_local, const void* halo, const int32_t* run_order, int warps, void* stream) {
    if (nrows < 0 || m < 0 || nnz < 0 || m > INT32_MAX || ncols_local > INT32_MAX) return RL_E_ARG;
    if (nrows == 0 || m == 0) return 0;
    if (x == y) return RL_E_ALIAS;
    const double w = dtype == RL_F32 ? 4.0 : 8.0;
    Span span(PK_SPMM, as_stream(stream), nnz * (w + 4.0) + (nrows + 1) * 8.0 + 2.0 * nrows * m * w,
              2.0 * nnz * m);
    if (dtype == RL_F32) return spmm_impl<float>(nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, (int)ncols_local, halo, run_order, warps, as_stream(stream));
    if (dtype == RL_F64) return spmm_impl<double>(nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, (int)ncols_local, halo, run_order, warps, as_stream(stream));
    return RL_E_DTYPE;
}

int rl_csr_spmm(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                const void* values, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m, void* stream) {
    return rl_csr_spmm_ex(dtype, nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, 0, nullptr, nullptr, 0, stream);
}

int rl_csr_spmm_halo(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                     const void* values, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m,
                     int64_t ncols_local, const void* halo, void* stream) {
    return rl_csr_spmm_ex(dtype, nrows, nnz, indptr, indices, values, x, ldx, y, ldy, m, ncols_local, halo, nullptr, 0, stream);
}

}  // extern "C"
