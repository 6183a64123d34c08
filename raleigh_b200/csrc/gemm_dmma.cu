// Dense operator application in fp64 on the FP64 tensor pipe (Matrix.apply, dense_cublas.py:732-776 ->
// cublasDgemm in the reference):
//   transp == 0:  Y[v,o] = alpha * sum_r X[v,r] * A[o,r] + beta * Y[v,o]     (A is (M,N) row-major, o < M, r < N)
//   transp != 0:  Y[v,o] = alpha * sum_r X[v,r] * A[r,o] + beta * Y[v,o]     (o < N, r < M)
// mma.sync m8n8k4 (DMMA): D (8 vectors x 8 outputs) += Xfrag (8 x 4, row) * Bfrag (4 x 8, col).  A lane
// (g = lane / 4, c = lane % 4) holds X[v0+g][r0+c], B(o0+g, r0+c) and gets D[g][2c], D[g][2c+1] -- two neighbouring
// outputs of one vector, stored as one 16-byte word.
//
// CTA tile: 32 / 64 / 128 vectors (the smallest that holds k) x 64 outputs, reduction in slices of 16 through a
// 4-stage cp.async ring; 8 warps, each 32 vectors x 8 / 16 / 32 outputs.  Global reads: 8 lanes fetch the 128
// contiguous bytes of a row slice (r-contiguous operands), a warp the 512 bytes of 64 outputs (o-contiguous matrix).  Shared-memory layouts make those 64-bit fragment loads conflict-free (a half-warp must hit
// 16 distinct 8-byte banks):
//   vectors, and the matrix when r is its contiguous index (transp == 0):  [r / 4][row][r % 4]  -> offset 4 g + c
//   the matrix when o is its contiguous index (transp != 0):               [o / 4][r][o % 4]    -> offset 4 c + g % 4
// Both are filled with 32-byte pieces of a global row (two 16-byte cp.async each, zero-filled outside the matrix).
// Measured (B200, 12 000 x 39 375 fp64, k = 128): see DESIGN.md section 4; FMA-pipe kernel (gemm_simt.cu) for comparison.
#include "common.cuh"

namespace rl {

constexpr int GD_BO = 64, GD_STAGES = 4, GD_THREADS = 256;
// shared-memory strides (doubles) with one padding row: the cp.async writes of a warp -- 8 lanes per 128-byte piece
// of a row, i.e. 4 different r-groups (resp. 16 o-groups) at once -- then spread over all banks
__host__ __device__ constexpr int gd_kmajor_stride(int rows) { return (rows + 1) * 4; }          // [r / 4][rows + 1][4]
__host__ __device__ constexpr int gd_nmajor_stride(int br) { return (br + 1) * 4; }            // [o / 4][br + 1][4]
__host__ __device__ constexpr int gd_xs(int bv, int br) { return (br / 4) * gd_kmajor_stride(bv); }
__host__ __device__ constexpr int gd_bs(int br) {
    return (br / 4) * gd_kmajor_stride(GD_BO) > (GD_BO / 4) * gd_nmajor_stride(br) ? (br / 4) * gd_kmajor_stride(GD_BO)
                                                                                  : (GD_BO / 4) * gd_nmajor_stride(br);
}
__host__ __device__ constexpr int gd_stage(int bv, int br) { return gd_xs(bv, br) + gd_bs(br); }
__host__ __device__ constexpr int gd_stages(int br) { return br >= 32 ? 3 : GD_STAGES; }      // ring depth

__device__ __forceinline__ void gd_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 16 bytes global -> shared; `bytes` (0, 8 or 16) are read, the rest is zero-filled
__device__ __forceinline__ void gd_cp16(void* smem, const void* gmem, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

// WVN = warps along the vectors (4, 2, 1): CTA tile 32 WVN vectors x 64 outputs; a warp owns 32 vectors x
// 64 / (8 / WVN) outputs = 4 x TO DMMA tiles.  BR = reduction slice per stage: 16, or 32 for the small-k variants,
// which are HBM-bound and want 256 contiguous bytes per row and request.
template <bool TRANSP, int WVN, int BR>
__global__ void __launch_bounds__(GD_THREADS, 2)
gemm_dmma_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ X, int64_t ldx,
                 double* __restrict__ Y, int64_t ldy, int64_t nvec, int64_t nout, int64_t nred, double alpha,
                 double beta) {
    constexpr int BV = 32 * WVN;
    constexpr int WON = 8 / WVN;                    // warps along the outputs
    constexpr int TO = GD_BO / (8 * WON);           // 8-wide output tiles per warp: 4, 2, 1
    constexpr int XSTR = gd_kmajor_stride(BV), BSTR = gd_kmajor_stride(GD_BO);
    constexpr int STAGE = gd_stage(BV, BR);
    constexpr int NST = gd_stages(BR);
    constexpr int NSTR = gd_nmajor_stride(BR);
    extern __shared__ __align__(16) double gd_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int wv = warp / WON, wo = warp % WON;
    const int64_t v0 = (int64_t)blockIdx.y * BV, o0 = (int64_t)blockIdx.x * GD_BO;

    double acc[4][TO][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < TO; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int nslices = (int)((nred + BR - 1) / BR);

    auto load_stage = [&](int slice, int stage) {
        double* xs = gd_smem + (size_t)stage * STAGE;
        double* bs = xs + gd_xs(BV, BR);
        const int64_t r0 = (int64_t)slice * BR;
        // r-contiguous operands: BR / 2 lanes fetch the 8 BR contiguous bytes of one row, RPP rows per pass
        {
            constexpr int CPR = BR / 2, RPP = GD_THREADS / CPR;   // 16-byte chunks per row, rows per pass
            const int ch = threadIdx.x % CPR;                     // 16-byte chunk of the row slice
            const int64_t gr = r0 + ch * 2;
            const int bytes_r = gr < nred ? (gr + 2 <= nred ? 16 : 8) : 0;
            const int off = (ch >> 1) * XSTR + (ch & 1) * 2;
#pragma unroll
            for (int pass = 0; pass < BV / RPP; ++pass) {
                const int row = threadIdx.x / CPR + pass * RPP;
                const int64_t gv = v0 + row;
                const int bytes = gv < nvec ? bytes_r : 0;
                gd_cp16(xs + off + row * 4, X + (bytes ? gv * ldx + gr : 0), bytes);
            }
            if (!TRANSP) {
                const int offb = (ch >> 1) * BSTR + (ch & 1) * 2;
#pragma unroll
                for (int pass = 0; pass < GD_BO / RPP; ++pass) {
                    const int row = threadIdx.x / CPR + pass * RPP;
                    const int64_t go = o0 + row;
                    const int bytes = go < nout ? bytes_r : 0;
                    gd_cp16(bs + offb + row * 4, A + (bytes ? go * lda + gr : 0), bytes);
                }
            }
        }
        if (TRANSP) {
            // o-contiguous matrix: a warp fetches the 512 bytes (64 outputs) of one row r, 8 rows per pass
            const int ch = lane;                                  // 16-byte chunk: outputs 2 ch, 2 ch + 1
            const int64_t go = o0 + ch * 2;
            const int bytes_o = go < nout ? (go + 2 <= nout ? 16 : 8) : 0;
            const int offb = (ch >> 1) * NSTR + (ch & 1) * 2;
#pragma unroll
            for (int pass = 0; pass < BR / 8; ++pass) {
                const int r = warp + pass * 8;
                const int64_t gr = r0 + r;
                const int bytes = gr < nred ? bytes_o : 0;
                gd_cp16(bs + offb + r * 4, A + (bytes ? gr * lda + go : 0), bytes);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nslices) load_stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int slice = 0; slice < nslices; ++slice) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
        __syncthreads();                                   // slice `slice` landed; everybody is done with slice - 1
        {
            const int nxt = slice + NST - 1;
            if (nxt < nslices) load_stage(nxt, nxt % NST);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const double* xs = gd_smem + (size_t)(slice % NST) * STAGE;
        const double* bs = xs + gd_xs(BV, BR);
#pragma unroll
        for (int rg = 0; rg < BR / 4; ++rg) {
            double xa[4], bb[TO];
#pragma unroll
            for (int t = 0; t < 4; ++t) xa[t] = xs[rg * XSTR + (wv * 32 + t * 8 + g) * 4 + c];
#pragma unroll
            for (int u = 0; u < TO; ++u) {
                const int o = wo * (8 * TO) + u * 8 + g;
                bb[u] = TRANSP ? bs[(o >> 2) * NSTR + (rg * 4 + c) * 4 + (o & 3)]
                               : bs[rg * BSTR + o * 4 + c];
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < TO; ++u) gd_dmma(acc[t][u][0], acc[t][u][1], xa[t], bb[u]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    // epilogue: lane (g, c) holds outputs 2c, 2c + 1 of vector g in every 8 x 8 tile
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t gv = v0 + wv * 32 + t * 8 + g;
        if (gv >= nvec) continue;
#pragma unroll
        for (int u = 0; u < TO; ++u) {
            const int64_t go = o0 + wo * (8 * TO) + u * 8 + 2 * c;
            if (go >= nout) continue;
            double* p = Y + gv * ldy + go;
            double r0v = alpha * acc[t][u][0], r1v = alpha * acc[t][u][1];
            if (go + 1 < nout) {
                if (beta != 0.0) { const double2 old = *reinterpret_cast<const double2*>(p); r0v = fma(beta, old.x, r0v); r1v = fma(beta, old.y, r1v); }
                *reinterpret_cast<double2*>(p) = make_double2(r0v, r1v);
            } else {
                if (beta != 0.0) r0v = fma(beta, *p, r0v);
                *p = r0v;
            }
        }
    }
}

bool gemm_dmma_supported(const void* a, int64_t lda, const void* x, int64_t ldx, const void* y, int64_t ldy) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al(a) && al(x) && al(y) && lda % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0;
}

template <bool TRANSP, int WVN, int BR>
static int gd_launch(const double* a, int64_t lda, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t k,
                     int64_t nout, int64_t nred, double alpha, double beta, cudaStream_t st) {
    constexpr size_t smem = (size_t)gd_stages(BR) * gd_stage(32 * WVN, BR) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel<TRANSP, WVN, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)((nout + GD_BO - 1) / GD_BO), (unsigned)((k + 32 * WVN - 1) / (32 * WVN)));
    gemm_dmma_kernel<TRANSP, WVN, BR><<<grid, GD_THREADS, smem, st>>>(a, lda, x, ldx, y, ldy, k, nout, nred, alpha, beta);
    ++g_launches;
    return check_launch();
}

int gemm_dmma(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
              int64_t k, int transp, double alpha, double beta, cudaStream_t st) {
    const int64_t nout = transp ? N : M, nred = transp ? M : N;
    const double* A = (const double*)a;
    const double* X = (const double*)x;
    double* Y = (double*)y;
    // vectors per CTA tile: the smallest of 32 / 64 / 128 that holds them all (else 128): no DMMA spent on padding
    const int wvn = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
#define RL_GD(T_, W_) return gd_launch<T_, W_, (W_ == 4 ? 16 : 32)>(A, lda, X, ldx, Y, ldy, k, nout, nred, alpha, beta, st)
    if (transp) { if (wvn == 1) RL_GD(true, 1); if (wvn == 2) RL_GD(true, 2); RL_GD(true, 4); }
    if (wvn == 1) RL_GD(false, 1);
    if (wvn == 2) RL_GD(false, 2);
    RL_GD(false, 4);
#undef RL_GD
}

}  // namespace rl
