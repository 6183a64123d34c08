// Dense operator application for a HANDFUL of vectors (k <= 8): Matrix.apply with the mean-shift vectors of
// the PCA operator (partial_svd.py:256-277: `ones`, `aves`, k = 1) and any other skinny block.  With so few
// vectors the product is a matrix-vector sweep, not a contraction: every element of A is used k times, the
// kernel is HBM-bound at M*N*w bytes (config 2: 1.89 GB, 0.29 ms at the measured 6.5 TB/s).  The tiled FMA
// kernel of gemm_simt.cu ran this at 0.11 of HBM bandwidth (2.64 ms, r1 verdict).
//
//   transp == 0:  Y[v,i] = alpha sum_j X[v,j] A[i,j] + beta Y[v,i]   one warp per group of 4 rows of A: the row is
//                 streamed with 128-bit loads, the k vectors come from L2 once per 4 rows, shuffle reduction
//   transp != 0:  Y[v,j] = alpha sum_i X[v,i] A[i,j] + beta Y[v,j]   CTA = 256 x 128-bit columns x a chunk of rows,
//                 coefficients broadcast from shared memory, partial sums per row chunk in a workspace, reduced
//                 in a fixed order by a second kernel (deterministic, no atomics)
#include "common.cuh"

namespace rl {

constexpr int SK_MAXK = 8;
constexpr int SK_ROWS = 4;        // rows of A per warp (transp == 0)
constexpr int SK_CHUNK = 64;      // rows of A per CTA (transp != 0)

template <typename T> struct Acc { using type = T; };

template <typename T, int K>
__global__ void __launch_bounds__(256)
skinny_rows_kernel(const T* __restrict__ A, int64_t lda, int64_t M, int64_t N, const T* __restrict__ X, int64_t ldx,
                   T* __restrict__ Y, int64_t ldy, T alpha, T beta, int vec) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = warp * SK_ROWS;
    if (r0 >= M) return;
    T acc[SK_ROWS][K];
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
        for (int v = 0; v < K; ++v) acc[r][v] = T(0);
    const int rows = (int)(M - r0 < SK_ROWS ? M - r0 : SK_ROWS);
    if (vec) {
        const int64_t nv = N / V;
        for (int64_t c = lane; c < nv; c += 32) {
            VT xv[K];
#pragma unroll
            for (int v = 0; v < K; ++v) xv[v] = __ldg(reinterpret_cast<const VT*>(X + (int64_t)v * ldx) + c);
#pragma unroll
            for (int r = 0; r < SK_ROWS; ++r) {
                if (r < rows) {
                    const VT a = ldg_stream(reinterpret_cast<const VT*>(A + (r0 + r) * lda) + c);
                    const T* ae = reinterpret_cast<const T*>(&a);
#pragma unroll
                    for (int v = 0; v < K; ++v) {
                        const T* xe = reinterpret_cast<const T*>(&xv[v]);
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[r][v] = fma(ae[e], xe[e], acc[r][v]);
                    }
                }
            }
        }
        for (int64_t j = nv * V + lane; j < N; j += 32)
#pragma unroll
            for (int r = 0; r < SK_ROWS; ++r)
                if (r < rows)
#pragma unroll
                    for (int v = 0; v < K; ++v) acc[r][v] = fma(A[(r0 + r) * lda + j], X[(int64_t)v * ldx + j], acc[r][v]);
    } else {
        for (int64_t j = lane; j < N; j += 32)
#pragma unroll
            for (int r = 0; r < SK_ROWS; ++r)
                if (r < rows)
#pragma unroll
                    for (int v = 0; v < K; ++v) acc[r][v] = fma(A[(r0 + r) * lda + j], X[(int64_t)v * ldx + j], acc[r][v]);
    }
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
        for (int v = 0; v < K; ++v) {
            T s = warp_sum(acc[r][v]);
            if (lane == 0 && r < rows) {
                T* y = Y + (int64_t)v * ldy + r0 + r;
                *y = (beta == T(0) ? T(0) : beta * *y) + alpha * s;
            }
        }
}

template <typename T, int K>
__global__ void __launch_bounds__(256)
skinny_cols_kernel(const T* __restrict__ A, int64_t lda, int64_t M, int64_t N, const T* __restrict__ X, int64_t ldx,
                   T* __restrict__ part, int vec) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    __shared__ T xs[K][SK_CHUNK];
    const int64_t i0 = (int64_t)blockIdx.y * SK_CHUNK;
    const int rows = (int)(M - i0 < SK_CHUNK ? M - i0 : SK_CHUNK);
    for (int e = threadIdx.x; e < K * SK_CHUNK; e += blockDim.x) {
        const int v = e / SK_CHUNK, i = e - v * SK_CHUNK;
        xs[v][i] = i < rows ? X[(int64_t)v * ldx + i0 + i] : T(0);
    }
    __syncthreads();
    T* out = part + (int64_t)blockIdx.y * K * N;
    if (vec) {
        const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // 128-bit column group
        if (c * V >= N) return;
        if ((c + 1) * V <= N) {
            T acc[K][V];
#pragma unroll
            for (int v = 0; v < K; ++v)
#pragma unroll
                for (int e = 0; e < V; ++e) acc[v][e] = T(0);
#pragma unroll 4
            for (int i = 0; i < rows; ++i) {
                const VT a = ldg_stream(reinterpret_cast<const VT*>(A + (i0 + i) * lda) + c);
                const T* ae = reinterpret_cast<const T*>(&a);
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    const T x = xs[v][i];
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[v][e] = fma(x, ae[e], acc[v][e]);
                }
            }
#pragma unroll
            for (int v = 0; v < K; ++v)
#pragma unroll
                for (int e = 0; e < V; ++e) out[(int64_t)v * N + c * V + e] = acc[v][e];
            return;
        }
        // ragged last group: scalar
        for (int64_t j = c * V; j < N; ++j)
            for (int v = 0; v < K; ++v) {
                T s = T(0);
                for (int i = 0; i < rows; ++i) s = fma(xs[v][i], A[(i0 + i) * lda + j], s);
                out[(int64_t)v * N + j] = s;
            }
    } else {
        const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= N) return;
        for (int v = 0; v < K; ++v) {
            T s = T(0);
            for (int i = 0; i < rows; ++i) s = fma(xs[v][i], A[(i0 + i) * lda + j], s);
            out[(int64_t)v * N + j] = s;
        }
    }
}

template <typename T>
__global__ void skinny_reduce_kernel(const T* __restrict__ part, int chunks, int k, int64_t N, T* __restrict__ Y,
                                     int64_t ldy, T alpha, T beta) {
    const int v = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (int64_t)gridDim.x * blockDim.x) {
        T s = T(0);
        for (int c = 0; c < chunks; ++c) s += part[((int64_t)c * k + v) * N + j];
        T* y = Y + (int64_t)v * ldy + j;
        *y = (beta == T(0) ? T(0) : beta * *y) + alpha * s;
    }
}

template <typename T, int K>
static int skinny_launch(const T* A, int64_t lda, int64_t M, int64_t N, const T* X, int64_t ldx, T* Y, int64_t ldy,
                         int transp, T alpha, T beta, cudaStream_t st) {
    constexpr int V = Vec128<T>::N;
    const int vec = host_aligned16(A) && host_aligned16(X) && lda % V == 0 && ldx % V == 0;
    if (!transp) {
        const int64_t warps = (M + SK_ROWS - 1) / SK_ROWS;
        const int64_t blocks = (warps * 32 + 255) / 256;
        skinny_rows_kernel<T, K><<<(unsigned)blocks, 256, 0, st>>>(A, lda, M, N, X, ldx, Y, ldy, alpha, beta, vec);
        return check_launch();
    }
    const int64_t chunks = (M + SK_CHUNK - 1) / SK_CHUNK;
    void* ws = nullptr;
    int rc = scratch_acquire((size_t)chunks * K * N * sizeof(T), &ws);
    if (rc) return rc;
    const int vecA = host_aligned16(A) && lda % V == 0;
    const int64_t groups = vecA ? (N + V - 1) / V : N;
    dim3 grid((unsigned)((groups + 255) / 256), (unsigned)chunks);
    skinny_cols_kernel<T, K><<<grid, 256, 0, st>>>(A, lda, M, N, X, ldx, (T*)ws, vecA);
    rc = check_launch();
    if (rc) return rc;
    int64_t gx = (N + 255) / 256; if (gx > 2048) gx = 2048;
    skinny_reduce_kernel<T><<<dim3((unsigned)gx, (unsigned)K), 256, 0, st>>>((const T*)ws, (int)chunks, K, N, Y, ldy, alpha, beta);
    return check_launch();
}

template <typename T>
int gemm_skinny(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
                int64_t k, int transp, double alpha, double beta, cudaStream_t st) {
    const T* A = (const T*)a; const T* X = (const T*)x; T* Y = (T*)y;
    switch (k) {
        case 1: return skinny_launch<T, 1>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 2: return skinny_launch<T, 2>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 3: return skinny_launch<T, 3>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 4: return skinny_launch<T, 4>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 5: return skinny_launch<T, 5>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 6: return skinny_launch<T, 6>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 7: return skinny_launch<T, 7>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        case 8: return skinny_launch<T, 8>(A, lda, M, N, X, ldx, Y, ldy, transp, (T)alpha, (T)beta, st);
        default: return RL_E_ARG;
    }
}

template int gemm_skinny<float>(const void*, int64_t, int64_t, int64_t, const void*, int64_t, void*, int64_t, int64_t,
                                int, double, double, cudaStream_t);
template int gemm_skinny<double>(const void*, int64_t, int64_t, int64_t, const void*, int64_t, void*, int64_t, int64_t,
                                 int, double, double, cudaStream_t);

}  // namespace rl
