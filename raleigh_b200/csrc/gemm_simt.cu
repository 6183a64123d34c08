// Dense operator application, FMA-pipe version (Matrix.apply, dense_cublas.py:732-776):
//   transp == 0:  Y[v,i] = alpha * sum_j X[v,j]*A[i,j] + beta*Y[v,i]     (A is (M,N) row-major)
//   transp != 0:  Y[v,j] = alpha * sum_i X[v,i]*A[i,j] + beta*Y[v,j]
// This is the general-shape / fp64 path and the fallback for shapes the
// tensor-core kernel (gemm_tc.cu) does not take.  64x64x16 tiles, 4x4 per thread.
#include "common.cuh"

namespace rl {

constexpr int GS_BV = 64, GS_BO = 64, GS_BK = 16, GS_THREADS = 256;

template <typename T, bool TRANSP>
__global__ void __launch_bounds__(GS_THREADS)
gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ X, int64_t ldx, T* __restrict__ Y,
                 int64_t ldy, int64_t nvec, int64_t nout, int64_t nred, T alpha, T beta) {
    __shared__ T Xs[GS_BK][GS_BV + 4];
    __shared__ T Bs[GS_BK][GS_BO + 4];
    const int64_t v0 = (int64_t)blockIdx.y * GS_BV, o0 = (int64_t)blockIdx.x * GS_BO;
    const int tv = threadIdx.x / 16, to = threadIdx.x % 16;   // 16 x 16 threads, 4x4 each
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = T(0);

    for (int64_t t0 = 0; t0 < nred; t0 += GS_BK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int idx = threadIdx.x + u * GS_THREADS;
            int v = idx / GS_BK, tt = idx % GS_BK;
            int64_t gv = v0 + v, gt = t0 + tt;
            Xs[tt][v] = (gv < nvec && gt < nred) ? __ldg(X + gv * ldx + gt) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int idx = threadIdx.x + u * GS_THREADS;
            if (TRANSP) {
                int tt = idx / GS_BO, o = idx % GS_BO;
                int64_t gt = t0 + tt, go = o0 + o;
                Bs[tt][o] = (gt < nred && go < nout) ? __ldg(A + gt * lda + go) : T(0);
            } else {
                int o = idx / GS_BK, tt = idx % GS_BK;
                int64_t gt = t0 + tt, go = o0 + o;
                Bs[tt][o] = (gt < nred && go < nout) ? __ldg(A + go * lda + gt) : T(0);
            }
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < GS_BK; ++tt) {
            T a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { a[e] = Xs[tt][tv * 4 + e]; b[e] = Bs[tt][to * 4 + e]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        int64_t gv = v0 + tv * 4 + x;
        if (gv >= nvec) continue;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            int64_t go = o0 + to * 4 + y;
            if (go >= nout) continue;
            T* p = Y + gv * ldy + go;
            *p = beta != T(0) ? fma(alpha, acc[x][y], beta * *p) : alpha * acc[x][y];
        }
    }
}

template <typename T>
int gemm_simt(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
              int64_t k, int transp, double alpha, double beta, cudaStream_t st) {
    int64_t nout = transp ? N : M, nred = transp ? M : N;
    dim3 grid((unsigned)((nout + GS_BO - 1) / GS_BO), (unsigned)((k + GS_BV - 1) / GS_BV));
    if (transp)
        gemm_simt_kernel<T, true><<<grid, GS_THREADS, 0, st>>>((const T*)a, lda, (const T*)x, ldx, (T*)y, ldy, k, nout, nred, (T)alpha, (T)beta);
    else
        gemm_simt_kernel<T, false><<<grid, GS_THREADS, 0, st>>>((const T*)a, lda, (const T*)x, ldx, (T*)y, ldy, k, nout, nred, (T)alpha, (T)beta);
    return check_launch();
}

template int gemm_simt<float>(const void*, int64_t, int64_t, int64_t, const void*, int64_t, void*, int64_t, int64_t, int, double, double, cudaStream_t);
template int gemm_simt<double>(const void*, int64_t, int64_t, int64_t, const void*, int64_t, void*, int64_t, int64_t, int, double, double, cudaStream_t);

}  // namespace rl
