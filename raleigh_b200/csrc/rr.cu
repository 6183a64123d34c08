// Device-resident small dense algebra of the block Jacobi-CG driver (raleigh_b200/jcg.py):
// everything the reference does on the host between two block-vector operations
// (raleigh/core/solver.py:854-873, 1331-1347, 1393-1473, 1589-1607) runs here on fp64
// matrices of order <= 2 x block size that never leave device memory.
//
//   rl_gram_dev / rl_dots_dev     Gram products straight into a device small matrix (fp64)
//   rl_update_dev                 block update with device-resident fp64 coefficients
//   rl_residual_dev               W = AX - X diag(lmd)             (solver.py:946-952)
//   rl_scale_rsqrt_dev            Y_i /= sqrt(|s2_i|)              (solver.py:1377-1378)
//   rl_small_gemm / copy / mirror / transpose
//   rl_rr_ritz_check              Rayleigh quotients + restart indicators (solver.py:859-873)
//   rl_rr_conjugation             conjugation coefficients Beta    (solver.py:1331-1347)
//   rl_rr_piv_chol                pivoted Cholesky with the reference's drop rule (:1749-1845)
//   rl_small_trsm                 triangular solves with many right-hand sides (:1685-1688, :1592)
//   rl_rr_estimates / rl_rr_select  change estimates and coefficient blocks (:1475-1493, :1593-1607)
// The symmetric eigensolver is in jacobi.cu.  All matrices are row-major fp64 with a leading
// dimension; everything is enqueued on the caller's stream, nothing synchronises.
#include "common.cuh"

namespace rl {
bool gemm_dmma_supported(const void* a, int64_t lda, const void* x, int64_t ldx, const void* y, int64_t ldy);
int gemm_dmma(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
              int64_t k, int transp, double alpha, double beta, cudaStream_t st);


// tcgen05 / TMEM / TMA GEMM of gemm_tc.cu (3xTF32, fp32 result)
bool gemm_tc_supported(const void* a, int64_t lda, const void* x, int64_t ldx);
int gemm_tc(const float* a_hi, const float* a_lo, int64_t lda, int64_t M, int64_t N, const float* x, int64_t ldx,
            float* y, int64_t ldy, int64_t k, int transp, double alpha, double beta, void* ws, size_t ws_bytes,
            cudaStream_t st);
size_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t k, int transp);

// fp32 blocks with at least this many vectors on both sides go to the tensor cores: the Gram product and the
// block update are then real contractions (32 flop/B at m = k = 128), not streaming kernels
constexpr int64_t TC_BLOCK_MIN = 48;

// ------------------------------------------------------------------ small helpers
__global__ void small_copy_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                  int64_t ldd, int rows, int cols) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)r * ldd + c] = src[(int64_t)r * lds + c];
}

// dst (cols x rows) = src^T
__global__ void small_transpose_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                       int64_t ldd, int rows, int cols) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * lds + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) dst[(int64_t)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// G[nx + j][i] = G[i][nx + j]  (lower-left block from the upper-right one)
__global__ void small_mirror_kernel(double* __restrict__ g, int64_t ld, int nx, int ny) {
    for (int j = blockIdx.y; j < ny; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x)
            g[(int64_t)(nx + j) * ld + i] = g[(int64_t)i * ld + nx + j];
}

template <typename T>
__global__ void small_cast_kernel(const double* __restrict__ src, int64_t lds, T* __restrict__ dst, int rows,
                                  int cols) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)r * cols + c] = (T)src[(int64_t)r * lds + c];
}

template <typename T>
__global__ void small_widen_kernel(const T* __restrict__ src, double* __restrict__ dst, int count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        dst[i] = (double)src[i];
}

static inline dim3 small_grid(int rows, int cols) {
    int gx = (cols + 127) / 128; if (gx < 1) gx = 1;
    int gy = rows < 1 ? 1 : (rows > 65535 ? 65535 : rows);
    return dim3((unsigned)gx, (unsigned)gy);
}

// ------------------------------------------------------------------ C = alpha op(A) op(B) + beta C
// Plain shared-memory tiled fp64 GEMM (64 x 64 tile, 4 x 4 per thread).  The matrices here are of
// order <= 2 x block size (plus the locked-vector Gram matrix, <= a few thousand): launch latency,
// not throughput, is what matters.
constexpr int SG_T = 64, SG_K = 16;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
small_gemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t lda,
                  const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C, int64_t ldc) {
    __shared__ double As[SG_K][SG_T + 1];
    __shared__ double Bs[SG_K][SG_T + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < K; k0 += SG_K) {
        for (int e = threadIdx.x; e < SG_K * SG_T; e += 256) {
            int kk, mm;
            if (TA) { mm = e % SG_T; kk = e / SG_T; } else { kk = e % SG_K; mm = e / SG_K; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < M && gk < K) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
            As[kk][mm] = v;
        }
        for (int e = threadIdx.x; e < SG_K * SG_T; e += 256) {
            int kk, nn;
            if (TB) { kk = e % SG_K; nn = e / SG_K; } else { nn = e % SG_T; kk = e / SG_T; }
            const int gn = n0 + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < N && gk < K) v = TB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx + 16 * j;
            if (gn >= N) continue;
            double* c = C + (int64_t)gm * ldc + gn;
            *c = (beta == 0.0 ? 0.0 : beta * *c) + alpha * acc[i][j];
        }
    }
}

// ------------------------------------------------------------------ streaming block kernels
// W[j] = AX[j] - lmd[j] * X[j]
template <typename T>
__global__ void __launch_bounds__(256)
residual_kernel(T* __restrict__ w, int64_t ldw, const T* __restrict__ ax, int64_t ldax, const T* __restrict__ x,
                int64_t ldx, int64_t n, const double* __restrict__ lmd, int vec) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    const int j = blockIdx.y;
    const T a = (T)lmd[j];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const int64_t nv = n / V;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
            VT p = ldg_stream(reinterpret_cast<const VT*>(ax + (int64_t)j * ldax) + i);
            VT q = ldg_stream(reinterpret_cast<const VT*>(x + (int64_t)j * ldx) + i);
            T* pe = reinterpret_cast<T*>(&p);
            const T* qe = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int t = 0; t < V; ++t) pe[t] = pe[t] - a * qe[t];
            reinterpret_cast<VT*>(w + (int64_t)j * ldw)[i] = p;
        }
        for (int64_t r = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
            w[(int64_t)j * ldw + r] = ax[(int64_t)j * ldax + r] - a * x[(int64_t)j * ldx + r];
    } else {
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
            w[(int64_t)j * ldw + r] = ax[(int64_t)j * ldax + r] - a * x[(int64_t)j * ldx + r];
    }
}

// Y[j] /= sqrt(|s2[j]|) unless that is zero (Vectors.scale semantics, dense_numpy.py:44-52)
template <typename T>
__global__ void __launch_bounds__(256)
scale_rsqrt_kernel(T* __restrict__ y, int64_t ldy, int64_t n, const double* __restrict__ s2, int vec) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    const int j = blockIdx.y;
    const double s = sqrt(fabs(s2[j]));
    if (s == 0.0) return;
    const T d = (T)s;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const int64_t nv = n / V;
        VT* row = reinterpret_cast<VT*>(y + (int64_t)j * ldy);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
            VT p = row[i];
            T* pe = reinterpret_cast<T*>(&p);
#pragma unroll
            for (int t = 0; t < V; ++t) pe[t] = pe[t] / d;
            row[i] = p;
        }
        for (int64_t r = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
            y[(int64_t)j * ldy + r] /= d;
    } else {
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
            y[(int64_t)j * ldy + r] /= d;
    }
}

static inline dim3 stream_grid(int64_t m, int64_t n, int per_thread) {
    int64_t gx = (n + 256 * (int64_t)per_thread - 1) / (256 * (int64_t)per_thread);
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    return dim3((unsigned)gx, (unsigned)m);
}

// ------------------------------------------------------------------ Ritz check (solver.py:859-873)
// lmd[i] = XAX[i][i] / XBX[i][i]; stats[0] = max|lmd - lmdx| / max|lmdx|; stats[1] = max|XBX - I|
__global__ void __launch_bounds__(1024)
ritz_check_kernel(const double* __restrict__ xax, const double* __restrict__ xbx, int64_t ld, int nx,
                  const double* __restrict__ lmdx, double* __restrict__ lmd, double* __restrict__ stats) {
    __shared__ double red[3][32];
    double dmax = 0.0, lmax = 0.0, omax = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const double v = xax[(int64_t)i * ld + i] / xbx[(int64_t)i * ld + i];
        lmd[i] = v;
        const double d = fabs(v - lmdx[i]);
        // NaNs must surface (the host restarts on them): max() with a NaN-propagating compare
        dmax = (d > dmax || d != d) ? d : dmax;
        lmax = fmax(lmax, fabs(lmdx[i]));
    }
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {
        const int i = e / nx, j = e - i * nx;
        const double d = fabs(xbx[(int64_t)i * ld + j] - (i == j ? 1.0 : 0.0));
        omax = (d > omax || d != d) ? d : omax;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, dmax, o), b = __shfl_xor_sync(0xffffffffu, lmax, o),
                     c = __shfl_xor_sync(0xffffffffu, omax, o);
        dmax = (a > dmax || a != a) ? a : dmax;
        lmax = fmax(lmax, b);
        omax = (c > omax || c != c) ? c : omax;
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dmax; red[1][threadIdx.x >> 5] = lmax; red[2][threadIdx.x >> 5] = omax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            const double a = red[0][w], b = red[1][w], c = red[2][w];
            dmax = (a > dmax || a != a) ? a : dmax;
            lmax = fmax(lmax, b);
            omax = (c > omax || c != c) ? c : omax;
        }
        stats[0] = nx > 0 ? dmax / lmax : 0.0;
        stats[1] = omax;
    }
}

// ------------------------------------------------------------------ conjugation (solver.py:1331-1347)
__global__ void conjugation_kernel(const double* __restrict__ zay, const double* __restrict__ zby,
                                   double* __restrict__ beta, int64_t ld, int nz, int ny,
                                   const double* __restrict__ lmd, const double* __restrict__ lmdz,
                                   const double* __restrict__ sy2, const double* __restrict__ sz2) {
    for (int z = blockIdx.y; z < nz; z += gridDim.y)
        for (int y = blockIdx.x * blockDim.x + threadIdx.x; y < ny; y += gridDim.x * blockDim.x) {
            const double num = zay[(int64_t)z * ld + y] - zby[(int64_t)z * ld + y] * lmd[y];
            const double den = lmdz[z] - lmd[y];
            const double s = sqrt(fabs(sy2[y])) / sqrt(fabs(sz2[z]));
            double b = 0.0;
            if (!(fabs(num) >= 100.0 * s * fabs(den))) b = num / den;
            beta[(int64_t)z * ld + y] = b;
        }
}

// ------------------------------------------------------------------ pivoted Cholesky (solver.py:1749-1845)
// One CTA of 1024 threads; the matrix stays in global memory (L1/L2 resident: <= 2 MB).
// Right-looking: after column i the whole trailing square is updated, so the pivot search reads
// the current diagonal directly (the reference keeps delayed updates in blocks of 64 and corrects
// the diagonal on the fly: same numbers up to rounding).
struct CholState {
    int done, dropped, drop_case, last_check, status, l;
};

constexpr int CH_THREADS = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];     // same order in every thread
    return s;
}

// y = U^-T x (forward) or x = U^-1 y (backward) on the leading p x p block, vector in shared memory `v`
// (overwritten by the solution).  Blocks of 32: the diagonal block is solved by warp 0 with shuffles from a
// shared-memory copy of the tile, the coupling to the other blocks is a coalesced matrix-vector update.
__device__ void tri_solve_vec(const double* __restrict__ U, int64_t ld, int p, double* v, bool forward,
                              double (*tile)[33], double* part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = (p + 31) / 32;
    for (int step = 0; step < nb; ++step) {
        const int J = forward ? step : nb - 1 - step;
        const int j0 = J * 32;
        {   // diagonal tile, identity-padded
            const int r = j0 + warp, c = j0 + lane;
            double t = (warp == lane) ? 1.0 : 0.0;
            if (r < p && c < p) t = U[(int64_t)r * ld + c];
            tile[warp][lane] = t;
        }
        if (!forward) {
            // rows of block J dotted with the already solved tail x[j0+32 ..)
            const int r = j0 + warp;
            double acc = 0.0;
            if (r < p)
                for (int c = j0 + 32 + lane; c < p; c += 32) acc = fma(U[(int64_t)r * ld + c], v[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) part[warp] = acc;
        }
        __syncthreads();
        if (warp == 0) {
            const int g = j0 + lane;
            double xv = g < p ? v[g] : 0.0;
            // one division per lane instead of one per elimination step on the critical path
            const double rinv = 1.0 / tile[lane][lane];
            if (forward) {
#pragma unroll 8
                for (int t = 0; t < 32; ++t) {
                    const double yt = __shfl_sync(0xffffffffu, xv, t) * __shfl_sync(0xffffffffu, rinv, t);
                    if (lane == t) xv = yt;
                    else if (lane > t) xv = fma(-tile[t][lane], yt, xv);     // U[t][s], s > t
                }
            } else {
                xv -= part[lane];
#pragma unroll 8
                for (int t = 31; t >= 0; --t) {
                    const double xt = __shfl_sync(0xffffffffu, xv, t) * __shfl_sync(0xffffffffu, rinv, t);
                    if (lane == t) xv = xt;
                    else if (lane < t) xv = fma(-tile[lane][t], xt, xv);     // U[s][t], s < t
                }
            }
            if (g < p) v[g] = xv;
        }
        __syncthreads();
        if (forward) {
            // x[i] -= sum_t U[j0+t][i] * y[j0+t]  for the rows below this block (coalesced over i)
            const int tmax = min(32, p - j0);
            for (int i = j0 + 32 + tid; i < p; i += blockDim.x) {
                double acc = v[i];
                for (int t = 0; t < tmax; ++t) acc = fma(-U[(int64_t)(j0 + t) * ld + i], v[j0 + t], acc);
                v[i] = acc;
            }
        }
        __syncthreads();
    }
}

// lmin/lmax estimate of the leading p x p block (solver.py:1828-1845); uniform result in every thread
__device__ double cond_inverse(const double* __restrict__ U, const double* __restrict__ A0, int64_t ld, int p,
                               const int* __restrict__ ind, double* vec, double (*tile)[33], double* part,
                               double* red) {
    const int tid = threadIdx.x;
    // lmax: 1-norm of the permuted leading block of the matrix being factorised
    double cmax = 0.0;
    for (int c = tid; c < p; c += blockDim.x) {
        const double* row = A0 + (int64_t)ind[c] * ld;          // symmetric: column c == row c
        double s = 0.0;
#pragma unroll 8
        for (int r = 0; r < p; ++r) s += fabs(__ldg(row + ind[r]));
        cmax = fmax(cmax, s);
    }
    for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = cmax;
    __syncthreads();
    double lmax = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) lmax = fmax(lmax, red[w]);
    // lmin: three steps of inverse iteration from the vector of ones
    for (int i = tid; i < p; i += blockDim.x) vec[i] = 1.0;
    __syncthreads();
    double s = (double)p, rq = 0.0;
    for (int it = 0; it < 3; ++it) {
        tri_solve_vec(U, ld, p, vec, true, tile, part);
        double t = 0.0;
        for (int i = tid; i < p; i += blockDim.x) t += vec[i] * vec[i];
        t = block_sum(t, red);
        rq = s / t;
        tri_solve_vec(U, ld, p, vec, false, tile, part);
        double s2 = 0.0;
        for (int i = tid; i < p; i += blockDim.x) s2 += vec[i] * vec[i];
        s = block_sum(s2, red);
    }
    return rq / lmax;
}

__device__ void zero_rows(double* A, int64_t ld, int n, int from) {
    for (int r = from; r < n; ++r)
        for (int c = threadIdx.x; c < n; c += blockDim.x) A[(int64_t)r * ld + c] = 0.0;
}

__global__ void __launch_bounds__(CH_THREADS)
piv_chol_kernel(double* __restrict__ A, double* __restrict__ A0, int64_t ld, int n, int k, double eps,
                int* __restrict__ ind, int* __restrict__ info) {
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ double tile[32][33];
    __shared__ double part[32];
    __shared__ int s_j;
    extern __shared__ double vec[];              // n doubles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 31, ty = tid >> 5;
    for (int i = tid; i < n; i += blockDim.x) ind[i] = i;
    for (int r = ty; r < n; r += 32)
        for (int c = tx; c < n; c += 32) A0[(int64_t)r * ld + c] = A[(int64_t)r * ld + c];
    __syncthreads();
    int dropped = 0, drop_case = 0, last_check = -1, status = 0, l = k;
    const int blk = 64;
    for (int i = 0; i < n; ++i) {
        if (i >= k) {
            // first maximum of the current diagonal over [i, n)
            double best = -1.0e308; int bj = n;
            for (int j = i + tid; j < n; j += blockDim.x) {
                const double d = A[(int64_t)j * ld + j];
                if (d > best) { best = d; bj = j; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
            }
            if (lane == 0) { red[warp] = best; redi[warp] = bj; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < 32; ++w)
                    if (red[w] > best || (red[w] == best && redi[w] < bj)) { best = red[w]; bj = redi[w]; }
                s_j = bj;
            }
            __syncthreads();
            const int j = s_j;
            if (j != i && j < n) {
                for (int c = tid; c < n; c += blockDim.x) {
                    const double a = A[(int64_t)i * ld + c];
                    A[(int64_t)i * ld + c] = A[(int64_t)j * ld + c];
                    A[(int64_t)j * ld + c] = a;
                }
                __syncthreads();
                for (int r = tid; r < n; r += blockDim.x) {
                    const double a = A[(int64_t)r * ld + i];
                    A[(int64_t)r * ld + i] = A[(int64_t)r * ld + j];
                    A[(int64_t)r * ld + j] = a;
                }
                if (tid == 0) { const int t = ind[i]; ind[i] = ind[j]; ind[j] = t; }
                __syncthreads();
            }
        }
        const double piv = A[(int64_t)i * ld + i];
        if ((i >= k && piv <= eps) || !(piv > 0.0)) {
            if (!(i >= k && piv <= eps)) status = 1;      // leading block not positive definite / NaN
            __syncthreads();
            zero_rows(A, ld, n, i);
            drop_case = status ? 2 : 1;
            dropped = n - i;
            break;
        }
        const double r = sqrt(piv);
        __syncthreads();                                   // everybody has read the pivot
        for (int c = i + 1 + tid; c < n; c += blockDim.x) {
            A[(int64_t)i * ld + c] /= r;
            A[(int64_t)c * ld + i] = 0.0;
        }
        if (tid == 0) A[(int64_t)i * ld + i] = r;
        __syncthreads();
        // row i staged in shared memory: the update's loads must not be ordered against its own stores
        for (int c = i + 1 + tid; c < n; c += blockDim.x) vec[c] = A[(int64_t)i * ld + c];
        __syncthreads();
        for (int rr = i + 1 + ty; rr < n; rr += 32) {
            const double f = vec[rr];
            double* row = A + (int64_t)rr * ld;
#pragma unroll 4
            for (int cc = i + 1 + tx; cc < n; cc += 32) row[cc] = fma(-f, vec[cc], row[cc]);
        }
        __syncthreads();
        if (i >= k && (i - l == blk - 1 || i == n - 1)) {
            last_check = i;
            const double ratio = cond_inverse(A, A0, ld, i + 1, ind, vec, tile, part, red);
            if (ratio <= eps) {
                __syncthreads();
                zero_rows(A, ld, n, i);
                drop_case = 2;
                dropped = n - i;
                break;
            }
            if (i - l == blk - 1) l += blk;
        }
    }
    __syncthreads();
    if (last_check < n - 1 && drop_case == 1) {
        // a pivot fell below eps: bisection for the largest well-conditioned leading block
        int i = last_check, j = n - dropped - 1;
        while (i < j) {
            const int mid = i + (j - i + 1) / 2;
            const double ratio = cond_inverse(A, A0, ld, mid + 1, ind, vec, tile, part, red);
            if (ratio <= eps) {
                if (j > mid) { j = mid; continue; }
                __syncthreads();
                zero_rows(A, ld, n, j);
                dropped = n - j;
                break;
            }
            i = mid;
        }
    }
    __syncthreads();
    if (tid == 0) { info[0] = dropped; info[1] = status; info[2] = drop_case; info[3] = last_check; }
}

// ---- the same factorisation with the working tile in SHARED memory --------------------------------
// The global-memory kernel above pays an L2 round trip per dependent step (measured: 2.8 ms at n = 256).
// Split as the reference itself does (solver.py:1756-1762): (1) unpivoted Cholesky of the leading k x k
// block in shared memory, (2) U12 = U11^-T A12 by the multi-CTA triangular solve, (3) A22 -= U12^T U12 by
// the tiled GEMM, (4) pivoted factorisation of the (n-k) x (n-k) trailing block in shared memory, columns
// of U12 swapped in global memory; the tile is flushed to global memory before every condition estimate.
constexpr int CH_SMEM_MAX = 150;       // 150 x 151 doubles = 181 KB

__global__ void __launch_bounds__(CH_THREADS)
chol_lead_smem_kernel(double* __restrict__ A, double* __restrict__ A0, int64_t ld, int n, int k,
                      int* __restrict__ ind, int* __restrict__ info) {
    extern __shared__ double T[];                 // k x (k + 1)
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int lt = k + 1;
    for (int i = tid; i < n; i += blockDim.x) ind[i] = i;
    for (int r = ty; r < n; r += 32)
        for (int c = tx; c < n; c += 32) {
            const double v = A[(int64_t)r * ld + c];
            A0[(int64_t)r * ld + c] = v;
            if (r < k && c < k) T[r * lt + c] = v;
        }
    __syncthreads();
    int status = 0, bad = k;
    for (int i = 0; i < k; ++i) {
        // two barriers per column: the pivot is read by everybody BEFORE the barrier that precedes its
        // overwrite (the diagonal entry is rewritten together with the trailing update, which never reads it)
        const double piv = T[i * lt + i];
        if (!(piv > 0.0)) { status = 1; bad = i; break; }
        const double rs = rsqrt(piv);
        for (int c = i + 1 + tid; c < k; c += blockDim.x) T[i * lt + c] *= rs;
        __syncthreads();
        if (tid == 0) T[i * lt + i] = piv * rs;
        for (int rr = i + 1 + ty; rr < k; rr += 32) {
            const double f = T[i * lt + rr];
            for (int cc = i + 1 + tx; cc < k; cc += 32)
                if (cc >= rr) T[rr * lt + cc] = fma(-f, T[i * lt + cc], T[rr * lt + cc]);
        }
        __syncthreads();
    }
    for (int r = ty; r < k; r += 32)
        for (int c = tx; c < k; c += 32)
            A[(int64_t)r * ld + c] = (c >= r && r < bad) ? T[r * lt + c] : 0.0;
    // the block below U11 is zero in the factor (solver.py:1761)
    for (int r = k + ty; r < n; r += 32)
        for (int c = tx; c < k; c += 32) A[(int64_t)r * ld + c] = 0.0;
    if (tid == 0) { info[0] = status ? n - bad : 0; info[1] = status; info[2] = status ? 2 : 0; info[3] = -1; }
}

__global__ void __launch_bounds__(CH_THREADS)
chol_tail_smem_kernel(double* __restrict__ A, const double* __restrict__ A0, int64_t ld, int n, int k, double eps,
                      int* __restrict__ ind, int* __restrict__ info, int no_estimates) {
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ double tile[32][33];
    __shared__ double part[32];
    __shared__ int s_j;
    extern __shared__ double dyn[];               // vec (n) | T (ny x (ny + 1))
    double* vec = dyn;
    const int ny = n - k, lt = ny + 1;
    double* T = dyn + ((n + 1) & ~1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tx = lane, ty = warp;
    if (info[1] != 0) return;                     // the leading block was not positive definite
    for (int r = ty; r < ny; r += 32)
        for (int c = tx; c < ny; c += 32) T[r * lt + c] = A[(int64_t)(k + r) * ld + k + c];
    __syncthreads();
    int dropped = 0, drop_case = 0, last_check = -1, l = k;
    const int blk = 64;
    // rows [0, upto) of the tile -> global factor (upper part, zeros below the diagonal); always from row 0:
    // later pivots swap columns of the rows factored earlier
    auto flush = [&](int upto) {
        for (int r = ty; r < upto; r += 32)
            for (int c = tx; c < ny; c += 32)
                A[(int64_t)(k + r) * ld + k + c] = c >= r ? T[r * lt + c] : 0.0;
        __syncthreads();
    };
    for (int ii = 0; ii < ny; ++ii) {
        const int i = k + ii;
        // first maximum of the diagonal: warp partials -> shared memory -> every warp folds them again
        // (no broadcast barrier, no serial loop)
        double best = -1.0e308; int bj = ny;
        for (int j = ii + tid; j < ny; j += blockDim.x) {
            const double d = T[j * lt + j];
            if (d > best) { best = d; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0) { red[warp] = best; redi[warp] = bj; }
        __syncthreads();
        best = red[lane]; bj = redi[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        const int j = bj;
        if (j != ii && j < ny) {
            // symmetric swap ii <-> j in one pass: rows and columns away from the 2 x 2 corner, the corner by
            // one thread; the same two columns of U12 in global memory
            for (int c = tid; c < ny; c += blockDim.x) {
                if (c == ii || c == j) continue;
                const double a = T[ii * lt + c];
                T[ii * lt + c] = T[j * lt + c];
                T[j * lt + c] = a;
                const double b = T[c * lt + ii];
                T[c * lt + ii] = T[c * lt + j];
                T[c * lt + j] = b;
            }
            if (tid == 0) {
                const double dii = T[ii * lt + ii], djj = T[j * lt + j], dij = T[ii * lt + j], dji = T[j * lt + ii];
                T[ii * lt + ii] = djj; T[j * lt + j] = dii; T[ii * lt + j] = dji; T[j * lt + ii] = dij;
                const int t = ind[i]; ind[i] = ind[k + j]; ind[k + j] = t;
            }
            for (int r = tid; r < k; r += blockDim.x) {
                const double a = A[(int64_t)r * ld + k + ii];
                A[(int64_t)r * ld + k + ii] = A[(int64_t)r * ld + k + j];
                A[(int64_t)r * ld + k + j] = a;
            }
        }
        __syncthreads();                  // also fences red/redi against the next pivot search
        const double piv = T[ii * lt + ii];
        if (piv <= eps || !(piv > 0.0)) {
            __syncthreads();
            for (int r = ii + ty; r < ny; r += 32)
                for (int c = tx; c < ny; c += 32) T[r * lt + c] = 0.0;
            __syncthreads();
            drop_case = 1;
            dropped = n - i;
            break;
        }
        const double rs = rsqrt(piv);
        for (int c = ii + 1 + tid; c < ny; c += blockDim.x) {
            T[ii * lt + c] *= rs;
            T[c * lt + ii] = 0.0;
        }
        __syncthreads();
        if (tid == 0) T[ii * lt + ii] = piv * rs;
        for (int rr = ii + 1 + ty; rr < ny; rr += 32) {
            const double f = T[ii * lt + rr];
            for (int cc = ii + 1 + tx; cc < ny; cc += 32) T[rr * lt + cc] = fma(-f, T[ii * lt + cc], T[rr * lt + cc]);
        }
        __syncthreads();
        if (!no_estimates && (i - l == blk - 1 || i == n - 1)) {
            last_check = i;
            flush(ii + 1);
            const double ratio = cond_inverse(A, A0, ld, i + 1, ind, vec, tile, part, red);
            if (ratio <= eps) {
                __syncthreads();
                for (int r = ii + ty; r < ny; r += 32)
                    for (int c = tx; c < ny; c += 32) T[r * lt + c] = 0.0;
                __syncthreads();
                drop_case = 2;
                dropped = n - i;
                break;
            }
            if (i - l == blk - 1) l += blk;
        }
    }
    __syncthreads();
    flush(ny);
    if (last_check < n - 1 && drop_case == 1) {
        int i = last_check, j = n - dropped - 1;
        while (i < j) {
            const int mid = i + (j - i + 1) / 2;
            const double ratio = cond_inverse(A, A0, ld, mid + 1, ind, vec, tile, part, red);
            if (ratio <= eps) {
                if (j > mid) { j = mid; continue; }
                __syncthreads();
                zero_rows(A, ld, n, j);
                dropped = n - j;
                break;
            }
            i = mid;
        }
    }
    __syncthreads();
    if (tid == 0) { info[0] = dropped; info[1] = 0; info[2] = drop_case; info[3] = last_check; }
}

// unpivoted Cholesky of one nb x nb diagonal block (nb <= 128) in shared memory, upper factor written back with
// zeros below the diagonal; info[0] = 1 + global index of the first non-positive pivot (sticky)
__global__ void __launch_bounds__(CH_THREADS)
potrf_diag_kernel(double* __restrict__ A, int64_t ld, int j0, int nb, int* __restrict__ info) {
    extern __shared__ double T[];                 // nb x (nb + 1)
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int lt = nb + 1;
    if (info[0] != 0) return;
    double* D = A + (int64_t)j0 * ld + j0;
    for (int r = ty; r < nb; r += 32)
        for (int c = tx; c < nb; c += 32) T[r * lt + c] = D[(int64_t)r * ld + c];
    __syncthreads();
    int bad = -1;
    for (int i = 0; i < nb; ++i) {
        const double piv = T[i * lt + i];
        if (!(piv > 0.0)) { bad = i; break; }
        const double rs = rsqrt(piv);
        __syncthreads();
        for (int c = i + 1 + tid; c < nb; c += blockDim.x) T[i * lt + c] *= rs;
        if (tid == 0) T[i * lt + i] = piv * rs;
        __syncthreads();
        for (int rr = i + 1 + ty; rr < nb; rr += 32) {
            const double f = T[i * lt + rr];
            for (int cc = i + 1 + tx; cc < nb; cc += 32)
                if (cc >= rr) T[rr * lt + cc] = fma(-f, T[i * lt + cc], T[rr * lt + cc]);
        }
        __syncthreads();
    }
    for (int r = ty; r < nb; r += 32)
        for (int c = tx; c < nb; c += 32) D[(int64_t)r * ld + c] = c >= r ? T[r * lt + c] : 0.0;
    if (bad >= 0 && tid == 0) info[0] = 1 + j0 + bad;
}

__global__ void zero_lower_kernel(double* __restrict__ A, int64_t ld, int n) {
    for (int r = blockIdx.y; r < n; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < r; c += gridDim.x * blockDim.x)
            A[(int64_t)r * ld + c] = 0.0;
}

// ------------------------------------------------------------------ triangular solves, many right-hand sides
// mode 0: solve U^T X = B (forward substitution with L = U^T);  mode 1: solve U X = B (backward).
// U is n x n upper triangular (ldu), B is n x r (ldb) and is overwritten by X.  One CTA owns 32
// right-hand sides and walks the block rows; the 32 x 32 diagonal systems are solved in registers
// by one warp (one column per lane), the coupling is a small tiled product from shared memory.
__global__ void __launch_bounds__(256)
small_trsm_kernel(const double* __restrict__ U, int64_t ldu, int n, double* __restrict__ B, int64_t ldb, int r,
                  int mode) {
    __shared__ double Lt[32][33];      // Lt[a][b]: coefficient of unknown (J*32 + b) in equation (I*32 + a)
    __shared__ double Xt[32][33];
    const int tid = threadIdx.x, col = tid & 31, rg = tid >> 5;      // rows rg*4 .. rg*4+3
    const int c0 = blockIdx.x * 32;
    const int nb = (n + 31) / 32;
    const bool colok = c0 + col < r;
    for (int step = 0; step < nb; ++step) {
        const int I = mode == 0 ? step : nb - 1 - step;
        double acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int row = I * 32 + rg * 4 + q;
            acc[q] = (row < n && colok) ? B[(int64_t)row * ldb + c0 + col] : 0.0;
        }
        for (int s2 = 0; s2 < step; ++s2) {
            const int J = mode == 0 ? s2 : nb - 1 - s2;
            __syncthreads();
            for (int e = tid; e < 1024; e += 256) {
                // the fastest index follows the contiguous direction of U (rows)
                const int a = mode == 0 ? (e & 31) : (e >> 5), b = mode == 0 ? (e >> 5) : (e & 31);
                const int ri = I * 32 + a, cj = J * 32 + b;
                double v = 0.0;
                if (ri < n && cj < n) v = mode == 0 ? U[(int64_t)cj * ldu + ri] : U[(int64_t)ri * ldu + cj];
                Lt[a][b] = v;
                const int xr = J * 32 + (e >> 5), xc = c0 + (e & 31);
                Xt[e >> 5][e & 31] = (xr < n && xc < r) ? B[(int64_t)xr * ldb + xc] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const double xv = Xt[b][col];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fma(-Lt[rg * 4 + q][b], xv, acc[q]);
            }
        }
        __syncthreads();
        // diagonal block (identity padded) and the right-hand side tile
        for (int e = tid; e < 1024; e += 256) {
            const int a = mode == 0 ? (e & 31) : (e >> 5), b = mode == 0 ? (e >> 5) : (e & 31);
            const int ri = I * 32 + a, cj = I * 32 + b;
            double v = (a == b) ? 1.0 : 0.0;
            if (ri < n && cj < n) v = mode == 0 ? U[(int64_t)cj * ldu + ri] : U[(int64_t)ri * ldu + cj];
            Lt[a][b] = v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Xt[rg * 4 + q][col] = acc[q];
        __syncthreads();
        if (rg == 0) {
            double a[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) a[t] = Xt[t][col];
            if (mode == 0) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const double x = a[t] / Lt[t][t];
                    a[t] = x;
#pragma unroll
                    for (int s = t + 1; s < 32; ++s) a[s] = fma(-Lt[s][t], x, a[s]);
                }
            } else {
#pragma unroll
                for (int t = 31; t >= 0; --t) {
                    const double x = a[t] / Lt[t][t];
                    a[t] = x;
#pragma unroll
                    for (int s = 0; s < t; ++s) a[s] = fma(-Lt[s][t], x, a[s]);
                }
            }
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int row = I * 32 + t;
                if (row < n && colok) B[(int64_t)row * ldb + c0 + col] = a[t];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ change estimates (solver.py:1475-1493)
// Column sel(i) of Q: i < leftX -> i, else nxy - rightX + (i - leftX).  QYX = Q[nx:, sel].
// dX[i] = ||QYX[:, i]||, dlmd[i] = sum_y (w[leftX + y] - w[sel(i)]) QYX[y, i]^2
__global__ void rr_estimates_kernel(const double* __restrict__ Q, int64_t ldq, const double* __restrict__ w,
                                    int nx, int ny, int leftX, int rightX, double* __restrict__ dX,
                                    double* __restrict__ dlmd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const int nxy = nx + ny;
    const int c = i < leftX ? i : nxy - rightX + (i - leftX);
    const double lx = w[c];
    double s = 0.0, d = 0.0;
    for (int y = 0; y < ny; ++y) {
        const double q = Q[(int64_t)(nx + y) * ldq + c];
        s = fma(q, q, s);
        d = fma((w[leftX + y] - lx) * q, q, d);
    }
    dX[i] = sqrt(s);
    dlmd[i] = d;
}

// CX[:, c] = Q[:, selnew(c)], lmdx[c] = w[selnew(c)];  CZ[:, c] = Q[:, leftXn + c], lmdz[c] = w[leftXn + c]
__global__ void rr_select_kernel(const double* __restrict__ Q, int64_t ldq, const double* __restrict__ w, int nxy,
                                 int leftXn, int rightXn, double* __restrict__ CX, int64_t ldcx,
                                 double* __restrict__ CZ, int64_t ldcz, double* __restrict__ lmdx,
                                 double* __restrict__ lmdz) {
    const int nxn = leftXn + rightXn, nz = nxy - nxn;
    for (int r = blockIdx.y; r < nxy; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += gridDim.x * blockDim.x) {
            if (c < nxn) {
                const int src = c < leftXn ? c : nxy - rightXn + (c - leftXn);
                CX[(int64_t)r * ldcx + c] = Q[(int64_t)r * ldq + src];
                if (r == 0) lmdx[c] = w[src];
            } else {
                const int z = c - nxn;
                if (z < nz) {
                    CZ[(int64_t)r * ldcz + z] = Q[(int64_t)r * ldq + leftXn + z];
                    if (r == 0) lmdz[z] = w[leftXn + z];
                }
            }
        }
}

// ------------------------------------------------------------------ small matrix <-> block of vectors
// dst[r*ldd + c] = (T) src[r][c]  (or src[c][r] when trans): fp64 coefficients -> typed block
template <typename T>
__global__ void small_to_block_kernel(const double* __restrict__ src, int64_t lds, int rows, int cols, int trans,
                                      T* __restrict__ dst, int64_t ldd) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)r * ldd + c] = (T)(trans ? src[(int64_t)c * lds + r] : src[(int64_t)r * lds + c]);
}

template <typename T>
__global__ void block_to_small_kernel(const T* __restrict__ src, int64_t lds, int rows, int cols,
                                      double* __restrict__ dst, int64_t ldd) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)r * ldd + c] = (double)src[(int64_t)r * lds + c];
}

// ------------------------------------------------------------------ partial SVD post-processing (partial_svd.py:163-235)
// Gershgorin certificate for the diagonally scaled Gram matrix S = D^-1/2 G D^-1/2:
// out[0] = max_i sum_{j != i} |S_ij|, out[1] = min_i G_ii, out[2] = max_i G_ii
__global__ void __launch_bounds__(1024)
psvd_gershgorin_kernel(const double* __restrict__ G, int64_t ld, int n, double* __restrict__ out) {
    __shared__ double r0[32], r1[32], r2[32];
    double rmax = 0.0, dmin = 1.0e308, dmax = -1.0e308;
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
        const double di = G[(int64_t)i * ld + i];
        double acc = 0.0;
        for (int j = threadIdx.x & 31; j < n; j += 32)
            if (j != i) acc += fabs(G[(int64_t)i * ld + j]) / sqrt(fabs(di * G[(int64_t)j * ld + j]));
        acc = warp_sum(acc);
        rmax = (acc > rmax || acc != acc) ? acc : rmax;
        dmin = fmin(dmin, di);
        dmax = fmax(dmax, di);
    }
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = rmax; r1[threadIdx.x >> 5] = dmin; r2[threadIdx.x >> 5] = dmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            rmax = (r0[w] > rmax || r0[w] != r0[w]) ? r0[w] : rmax;
            dmin = fmin(dmin, r1[w]); dmax = fmax(dmax, r2[w]);
        }
        out[0] = rmax; out[1] = dmin; out[2] = dmax;
    }
}

// S = D^-1/2 G D^-1/2
__global__ void psvd_scale_kernel(const double* __restrict__ G, int64_t ld, int n, double* __restrict__ S, int64_t lds) {
    for (int r = blockIdx.y; r < n; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
            S[(int64_t)r * lds + c] = G[(int64_t)r * ld + c] / sqrt(fabs(G[(int64_t)r * ld + r] * G[(int64_t)c * ld + c]));
}

// dst[r][c] = src[r][c] * s[c]
__global__ void small_scale_cols_kernel(const double* __restrict__ src, int64_t lds, int rows, int cols,
                                        const double* __restrict__ sc, double* __restrict__ dst, int64_t ldd) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x)
            dst[(int64_t)r * ldd + c] = src[(int64_t)r * lds + c] * sc[c];
}

__global__ void small_identity_kernel(double* __restrict__ a, int64_t ld, int n) {
    for (int r = blockIdx.y; r < n; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
            a[(int64_t)r * ld + c] = r == c ? 1.0 : 0.0;
}

// out[0] = sum_ij g_ii uinv_ij^2 (squared Frobenius norm of D^1/2 U^-1); one CTA, fixed summation order
__global__ void __launch_bounds__(1024)
psvd_invbound_kernel(const double* __restrict__ uinv, int64_t ldu, const double* __restrict__ g, int64_t ldg, int n,
                     double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int r = threadIdx.x >> 5; r < n; r += 32) {
        const double d = g[(int64_t)r * ldg + r];
        double s = 0.0;
        for (int c = threadIdx.x & 31; c < n; c += 32) { const double v = uinv[(int64_t)r * ldu + c]; s = fma(v, v, s); }
        acc = fma(d, s, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < 32; ++w) t += red[w]; out[0] = t; }
}

// eigenpairs ascending (w, Q) -> descending singular values and the two coefficient matrices:
// q[:, c] = Q[:, n-1-c];  cs[:, c] = q[:, c] / sigma[c] (0 where sigma == 0);  sigma[c] = sqrt(max(w[n-1-c], 0))
__global__ void psvd_coeffs_kernel(const double* __restrict__ Q, int64_t ldq, const double* __restrict__ w, int n,
                                   double* __restrict__ q, double* __restrict__ cs, int64_t ldo,
                                   double* __restrict__ sigma) {
    for (int r = blockIdx.y; r < n; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
            const double lam = w[n - 1 - c];
            const double sg = lam > 0.0 ? sqrt(lam) : 0.0;
            const double v = Q[(int64_t)r * ldq + (n - 1 - c)];
            q[(int64_t)r * ldo + c] = v;
            cs[(int64_t)r * ldo + c] = sg > 0.0 ? v / sg : 0.0;
            if (r == 0) sigma[c] = sg;
        }
}

}  // namespace rl

using namespace rl;

extern "C" {

int rl_small_copy(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols,
                  void* stream) {
    if (rows < 0 || cols < 0) return RL_E_ARG;
    if (rows == 0 || cols == 0) return 0;
    small_copy_kernel<<<small_grid((int)rows, (int)cols), 128, 0, as_stream(stream)>>>(src, lds, dst, ldd, (int)rows, (int)cols);
    return check_launch();
}

int rl_small_transpose(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols,
                       void* stream) {
    if (rows < 0 || cols < 0) return RL_E_ARG;
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    small_transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, lds, dst, ldd, (int)rows, (int)cols);
    return check_launch();
}

int rl_small_mirror(double* g, int64_t ld, int64_t nx, int64_t ny, void* stream) {
    if (nx < 0 || ny < 0) return RL_E_ARG;
    if (nx == 0 || ny == 0) return 0;
    small_mirror_kernel<<<small_grid((int)ny, (int)nx), 128, 0, as_stream(stream)>>>(g, ld, (int)nx, (int)ny);
    return check_launch();
}

int rl_small_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                  int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream) {
    if (M < 0 || N < 0 || K < 0) return RL_E_ARG;
    if (M == 0 || N == 0) return 0;
    dim3 grid((unsigned)((N + SG_T - 1) / SG_T), (unsigned)((M + SG_T - 1) / SG_T));
    cudaStream_t st = as_stream(stream);
    Span span(PK_SMALL, st, (1.0 * M * K + 1.0 * K * N + 2.0 * M * N) * 8, 2.0 * M * N * K);
    // big products (the locked-vector projections late in a config-2 solve: 1000 x 1000 x 128, 0.22 ms on the FMA
    // kernel) go to the FP64 tensor-pipe GEMM of the dense apply: C[v,o] = sum_r A[v,r] op(B)[r,o] is its
    // "X times matrix" form with X = A
    if (!transa && K > 0 && (double)M * N * K >= 5.0e7 && C != A && C != B && g_knob[KNOB_GEMM_DMMA] >= 0 &&
        gemm_dmma_supported(B, ldb, A, lda, C, ldc))
        return transb ? gemm_dmma(B, ldb, N, K, A, lda, C, ldc, M, 0, alpha, beta, st)
                      : gemm_dmma(B, ldb, K, N, A, lda, C, ldc, M, 1, alpha, beta, st);
    if (transa && transb) small_gemm_kernel<true, true><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (transa) small_gemm_kernel<true, false><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (transb) small_gemm_kernel<false, true><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
    else small_gemm_kernel<false, false><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc);
    return check_launch();
}

int rl_small_trsm(int mode, const double* U, int64_t ldu, int64_t n, double* B, int64_t ldb, int64_t r,
                  void* stream) {
    if (n < 0 || r < 0 || (mode != 0 && mode != 1)) return RL_E_ARG;
    if (n == 0 || r == 0) return 0;
    Span span(PK_SMALL, as_stream(stream), (0.5 * n * n + 2.0 * n * r) * 8, 1.0 * n * n * r);
    small_trsm_kernel<<<(unsigned)((r + 31) / 32), 256, 0, as_stream(stream)>>>(U, ldu, (int)n, B, ldb, (int)r, mode);
    return check_launch();
}

/* Gram product into a device-resident fp64 matrix: out[i*ldout + j] = <o_i, s_j>, i < k, j < m
 * (Vectors.dot, dense_cublas.py:245-269, without the D2H copy).  fp32 data accumulate in fp64. */
int rl_gram_dev(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo, int64_t k,
                int64_t n, double* out, int64_t ldout, void* stream) {
    if (m < 0 || k < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || k == 0) return 0;
    if (dtype == RL_F32 && m >= TC_BLOCK_MIN && k >= TC_BLOCK_MIN && n >= 1024 && g_knob[KNOB_BLOCK_TC] >= 0 &&
        gemm_tc_supported(s, lds, o, ldo)) {
        // G (k x m) = O (k x n) . S^T: the dense-apply kernel with the block S as the data matrix, n as the
        // reduction dimension (split over the SMs, partial tiles reduced in a fixed order)
        const size_t wsb = (gemm_tc_ws_bytes(m, n, k, 0) + 255) & ~size_t(255);
        const int64_t ldt = (m + 3) / 4 * 4;
        void* base = nullptr;
        int rc = scratch_acquire(wsb + (size_t)k * ldt * sizeof(float), &base);
        if (rc) return rc;
        float* tmp = (float*)((char*)base + wsb);
        {
            Span span(PK_GRAM, as_stream(stream), (double)(s == o ? m : m + k) * n * 4.0 + (double)k * m * 8.0, 2.0 * n * m * k);
            rc = gemm_tc((const float*)s, nullptr, lds, m, n, (const float*)o, ldo, tmp, ldt, k, 0, 1.0, 0.0, base, wsb,
                         as_stream(stream));
        }
        if (rc) return rc;
        return rl_block_to_small(RL_F32, tmp, ldt, k, m, out, ldout, stream);
    }
    const size_t wsb = (rl_gram_acc64_ws_bytes(dtype, m, k, n) + 255) & ~size_t(255);
    void* base = nullptr;
    int rc = scratch_acquire(wsb + (size_t)k * m * sizeof(double), &base);
    if (rc) return rc;
    double* tmp = (double*)((char*)base + wsb);
    rc = rl_gram_acc64(dtype, s, lds, m, o, ldo, k, n, tmp, base, wsb, stream);
    if (rc) return rc;
    return rl_small_copy(tmp, m, out, ldout, k, m, stream);
}

/* Vectors.dots into a device fp64 vector: out[i] = <o_i, s_i> (dense_cublas.py:222-243) */
int rl_dots_dev(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n,
                double* out, void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    if (m == 0) return 0;
    if (dtype == RL_F64) {
        const size_t wsb = rl_dots_ws_bytes(dtype, m, n);
        void* ws = nullptr;
        if (wsb) { int rc = scratch_acquire(wsb, &ws); if (rc) return rc; }
        return rl_dots(dtype, s, lds, o, ldo, m, n, out, ws, wsb, stream);
    }
    if (dtype != RL_F32) return RL_E_DTYPE;
    const size_t wsb = (rl_dots_ws_bytes(dtype, m, n) + 255) & ~size_t(255);
    void* base = nullptr;
    int rc = scratch_acquire(wsb + (size_t)m * sizeof(float), &base);
    if (rc) return rc;
    float* tmp = (float*)((char*)base + wsb);
    rc = rl_dots(dtype, s, lds, o, ldo, m, n, tmp, base, wsb, stream);
    if (rc) return rc;
    small_widen_kernel<float><<<(unsigned)((m + 255) / 256), 256, 0, as_stream(stream)>>>(tmp, out, (int)m);
    return check_launch();
}

/* Block update with device-resident fp64 coefficients (Vectors.multiply / add(other, s, q),
 * dense_cublas.py:271-342, without the H2D copy of q): Out = beta Out + alpha q^T X, q (k x m, ldq). */
int rl_update_dev(int dtype, void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx, int64_t k,
                  const double* q, int64_t ldq, double alpha, double beta, int64_t n, void* stream) {
    if (m < 0 || k < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    if (dtype == RL_F64 || k == 0) return rl_update(dtype, out, ldo, m, x, ldx, k, q, ldq, 1, alpha, beta, n, stream);
    if (dtype != RL_F32) return RL_E_DTYPE;
    if (out == x) return RL_E_ALIAS;
    if (m >= TC_BLOCK_MIN && k >= TC_BLOCK_MIN && n >= 1024 && g_knob[KNOB_BLOCK_TC] >= 0 && (ldo % 4 == 0) &&
        host_aligned16(out)) {
        // Out (m x n) = beta Out + alpha q^T (m x k) . X (k x n): dense apply, transposed form, X as the data matrix
        const int64_t ldq32 = (k + 3) / 4 * 4;
        const size_t wsb = (gemm_tc_ws_bytes(k, n, m, 1) + 255) & ~size_t(255);
        void* base = nullptr;
        int rc = scratch_acquire(wsb + (size_t)m * ldq32 * sizeof(float), &base);
        if (rc) return rc;
        float* qt = (float*)((char*)base + wsb);
        if (gemm_tc_supported(x, ldx, qt, ldq32)) {
            rc = rl_small_to_block(RL_F32, q, ldq, m, k, 1, qt, ldq32, stream);
            if (rc) return rc;
            Span span(PK_UPDATE, as_stream(stream), (1.0 * k + (beta != 0.0 ? 2.0 : 1.0) * m) * n * 4.0, 2.0 * n * k * m);
            return gemm_tc((const float*)x, nullptr, ldx, k, n, qt, ldq32, (float*)out, ldo, m, 1, alpha, beta, base, wsb,
                           as_stream(stream));
        }
    }
    void *pinned = nullptr, *dev = nullptr;
    int rc = staging_acquire((size_t)k * m * sizeof(float), &pinned, &dev);       // device half of the ring only
    if (rc) return rc;
    small_cast_kernel<float><<<small_grid((int)k, (int)m), 128, 0, as_stream(stream)>>>(q, ldq, (float*)dev, (int)k, (int)m);
    rc = check_launch();
    if (rc) return rc;
    return rl_update(dtype, out, ldo, m, x, ldx, k, dev, m, 1, alpha, beta, n, stream);
}

int rl_residual_dev(int dtype, void* w, int64_t ldw, const void* ax, int64_t ldax, const void* x, int64_t ldx,
                    int64_t m, int64_t n, const double* lmd, void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == RL_F64) {
        const int vec = host_aligned16(w) && host_aligned16(ax) && host_aligned16(x) && ldw % 2 == 0 && ldax % 2 == 0 && ldx % 2 == 0;
        residual_kernel<double><<<stream_grid(m, n, 8), 256, 0, st>>>((double*)w, ldw, (const double*)ax, ldax, (const double*)x, ldx, n, lmd, vec);
    } else if (dtype == RL_F32) {
        const int vec = host_aligned16(w) && host_aligned16(ax) && host_aligned16(x) && ldw % 4 == 0 && ldax % 4 == 0 && ldx % 4 == 0;
        residual_kernel<float><<<stream_grid(m, n, 16), 256, 0, st>>>((float*)w, ldw, (const float*)ax, ldax, (const float*)x, ldx, n, lmd, vec);
    } else return RL_E_DTYPE;
    return check_launch();
}

int rl_scale_rsqrt_dev(int dtype, void* y, int64_t ldy, int64_t m, int64_t n, const double* s2, void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == RL_F64) {
        const int vec = host_aligned16(y) && ldy % 2 == 0;
        scale_rsqrt_kernel<double><<<stream_grid(m, n, 8), 256, 0, st>>>((double*)y, ldy, n, s2, vec);
    } else if (dtype == RL_F32) {
        const int vec = host_aligned16(y) && ldy % 4 == 0;
        scale_rsqrt_kernel<float><<<stream_grid(m, n, 16), 256, 0, st>>>((float*)y, ldy, n, s2, vec);
    } else return RL_E_DTYPE;
    return check_launch();
}

int rl_rr_ritz_check(const double* xax, const double* xbx, int64_t ld, int64_t nx, const double* lmdx,
                     double* lmd, double* stats, void* stream) {
    if (nx < 0) return RL_E_ARG;
    ritz_check_kernel<<<1, 1024, 0, as_stream(stream)>>>(xax, xbx, ld, (int)nx, lmdx, lmd, stats);
    return check_launch();
}

int rl_rr_conjugation(const double* zay, const double* zby, double* beta, int64_t ld, int64_t nz, int64_t ny,
                      const double* lmd, const double* lmdz, const double* sy2, const double* sz2, void* stream) {
    if (nz < 0 || ny < 0) return RL_E_ARG;
    if (nz == 0 || ny == 0) return 0;
    conjugation_kernel<<<small_grid((int)nz, (int)ny), 128, 0, as_stream(stream)>>>(zay, zby, beta, ld, (int)nz, (int)ny, lmd, lmdz, sy2, sz2);
    return check_launch();
}

/* Pivoted Cholesky of the leading n x n block of `a` in place (upper factor), first k columns
 * unpivoted; `a0` is an n x n (same ld) scratch copy; ind (n ints) the permutation;
 * info[0] = dropped, info[1] = status (1: not positive definite), info[2] = drop case, info[3] = last check. */
int rl_rr_piv_chol(double* a, double* a0, int64_t ld, int64_t n, int64_t k, double eps, int* ind, int* info,
                   void* stream) {
    if (n < 0 || k < 0 || k > n || n > 4096) return RL_E_ARG;
    cudaStream_t st = as_stream(stream);
    if (n == 0) return (int)cudaMemsetAsync(info, 0, 4 * sizeof(int), st);
    Span span(PK_PIV_CHOL, st, 2.0 * n * n * 8, 1.0 * n * n * n / 3.0);
    const int64_t ny = n - k;
    if (k <= CH_SMEM_MAX && ny <= CH_SMEM_MAX && g_knob[KNOB_CHOL_GLOBAL] == 0) {
        static bool configured = false;
        if (!configured) {
            const int cap = (int)((size_t)CH_SMEM_MAX * (CH_SMEM_MAX + 1) * sizeof(double) + 4096 * sizeof(double));
            RL_CUDA(cudaFuncSetAttribute(chol_lead_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            RL_CUDA(cudaFuncSetAttribute(chol_tail_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            configured = true;
        }
        chol_lead_smem_kernel<<<1, CH_THREADS, (size_t)(k > 0 ? k * (k + 1) : 1) * sizeof(double), st>>>(a, a0, ld, (int)n, (int)k, ind, info);
        int rc = check_launch();
        if (rc) return rc;
        if (ny == 0) return 0;
        if (k > 0) {
            rc = rl_small_trsm(0, a, ld, k, a + k, ld, ny, stream);                                   // U12 = U11^-T A12
            if (rc) return rc;
            rc = rl_small_gemm(1, 0, ny, ny, k, -1.0, a + k, ld, a + k, ld, 1.0, a + k * ld + k, ld, stream);   // A22 -= U12^T U12
            if (rc) return rc;
        }
        const size_t smem = (size_t)(((n + 1) & ~int64_t(1)) + ny * (ny + 1)) * sizeof(double);
        chol_tail_smem_kernel<<<1, CH_THREADS, smem, st>>>(a, a0, ld, (int)n, (int)k, eps, ind, info, g_knob[KNOB_CHOL_NOEST] == 1);
        return check_launch();
    }
    piv_chol_kernel<<<1, CH_THREADS, (size_t)n * sizeof(double), st>>>(a, a0, ld, (int)n, (int)k, eps, ind, info);
    return check_launch();
}

int rl_rr_estimates(const double* q, int64_t ldq, const double* w, int64_t nx, int64_t ny, int64_t leftX,
                    int64_t rightX, double* dX, double* dlmd, void* stream) {
    if (nx < 0 || ny < 0) return RL_E_ARG;
    if (nx == 0) return 0;
    rr_estimates_kernel<<<(unsigned)((nx + 127) / 128), 128, 0, as_stream(stream)>>>(q, ldq, w, (int)nx, (int)ny, (int)leftX, (int)rightX, dX, dlmd);
    return check_launch();
}

int rl_rr_select(const double* q, int64_t ldq, const double* w, int64_t nxy, int64_t leftXn, int64_t rightXn,
                 double* cx, int64_t ldcx, double* cz, int64_t ldcz, double* lmdx, double* lmdz, void* stream) {
    if (nxy < 0 || leftXn < 0 || rightXn < 0 || leftXn + rightXn > nxy) return RL_E_ARG;
    if (nxy == 0) return 0;
    rr_select_kernel<<<small_grid((int)nxy, (int)nxy), 128, 0, as_stream(stream)>>>(q, ldq, w, (int)nxy, (int)leftXn, (int)rightXn, cx, ldcx, cz, ldcz, lmdx, lmdz);
    return check_launch();
}


int rl_small_to_block(int dtype, const double* src, int64_t lds, int64_t rows, int64_t cols, int trans, void* dst,
                      int64_t ldd, void* stream) {
    if (rows < 0 || cols < 0) return RL_E_ARG;
    if (rows == 0 || cols == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == RL_F32) small_to_block_kernel<float><<<small_grid((int)rows, (int)cols), 128, 0, st>>>(src, lds, (int)rows, (int)cols, trans, (float*)dst, ldd);
    else if (dtype == RL_F64) small_to_block_kernel<double><<<small_grid((int)rows, (int)cols), 128, 0, st>>>(src, lds, (int)rows, (int)cols, trans, (double*)dst, ldd);
    else return RL_E_DTYPE;
    return check_launch();
}

int rl_block_to_small(int dtype, const void* src, int64_t lds, int64_t rows, int64_t cols, double* dst, int64_t ldd,
                      void* stream) {
    if (rows < 0 || cols < 0) return RL_E_ARG;
    if (rows == 0 || cols == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (dtype == RL_F32) block_to_small_kernel<float><<<small_grid((int)rows, (int)cols), 128, 0, st>>>((const float*)src, lds, (int)rows, (int)cols, dst, ldd);
    else if (dtype == RL_F64) block_to_small_kernel<double><<<small_grid((int)rows, (int)cols), 128, 0, st>>>((const double*)src, lds, (int)rows, (int)cols, dst, ldd);
    else return RL_E_DTYPE;
    return check_launch();
}

int rl_psvd_gershgorin(const double* g, int64_t ld, int64_t n, double* out3, void* stream) {
    if (n <= 0) return RL_E_ARG;
    psvd_gershgorin_kernel<<<1, 1024, 0, as_stream(stream)>>>(g, ld, (int)n, out3);
    return check_launch();
}

int rl_psvd_scale(const double* g, int64_t ld, int64_t n, double* s, int64_t lds, void* stream) {
    if (n <= 0) return RL_E_ARG;
    psvd_scale_kernel<<<small_grid((int)n, (int)n), 128, 0, as_stream(stream)>>>(g, ld, (int)n, s, lds);
    return check_launch();
}

int rl_psvd_coeffs(const double* qin, int64_t ldq, const double* w, int64_t n, double* q, double* cs, int64_t ldo,
                   double* sigma, void* stream) {
    if (n <= 0) return RL_E_ARG;
    psvd_coeffs_kernel<<<small_grid((int)n, (int)n), 128, 0, as_stream(stream)>>>(qin, ldq, w, (int)n, q, cs, ldo, sigma);
    return check_launch();
}


int rl_small_potrf(double* a, int64_t ld, int64_t n, int* info_d, void* stream) {
    if (n < 0 || !info_d) return RL_E_ARG;
    cudaStream_t st = as_stream(stream);
    RL_CUDA(cudaMemsetAsync(info_d, 0, sizeof(int), st));
    if (n == 0) return 0;
    Span span(PK_SMALL, st, 2.0 * n * n * 8, 1.0 * n * n * n / 3.0);
    const int64_t NB = 128;
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * (NB + 1) * sizeof(double))));
        configured = true;
    }
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
        const int64_t nb = n - j0 < NB ? n - j0 : NB;
        const int64_t rest = n - j0 - nb;
        potrf_diag_kernel<<<1, CH_THREADS, (size_t)(nb * (nb + 1)) * sizeof(double), st>>>(a, ld, (int)j0, (int)nb, info_d);
        int rc = check_launch();
        if (rc) return rc;
        if (rest > 0) {
            double* d = a + j0 * ld + j0;
            rc = rl_small_trsm(0, d, ld, nb, d + nb, ld, rest, stream);                                  // U12 = U11^-T A12
            if (rc) return rc;
            rc = rl_small_gemm(1, 0, rest, rest, nb, -1.0, d + nb, ld, d + nb, ld, 1.0, d + nb * ld + nb, ld, stream);
            if (rc) return rc;
        }
    }
    zero_lower_kernel<<<small_grid((int)n, (int)n), 128, 0, st>>>(a, ld, (int)n);
    return check_launch();
}


int rl_small_set_identity(double* a, int64_t ld, int64_t n, void* stream) {
    if (n < 0) return RL_E_ARG;
    if (n == 0) return 0;
    small_identity_kernel<<<small_grid((int)n, (int)n), 128, 0, as_stream(stream)>>>(a, ld, (int)n);
    return check_launch();
}

int rl_psvd_invbound(const double* uinv, int64_t ldu, const double* g, int64_t ldg, int64_t n, double* out,
                     void* stream) {
    if (n <= 0) return RL_E_ARG;
    psvd_invbound_kernel<<<1, 1024, 0, as_stream(stream)>>>(uinv, ldu, g, ldg, (int)n, out);
    return check_launch();
}


int rl_small_scale_cols(const double* src, int64_t lds, int64_t rows, int64_t cols, const double* s, double* dst,
                        int64_t ldd, void* stream) {
    if (rows < 0 || cols < 0) return RL_E_ARG;
    if (rows == 0 || cols == 0) return 0;
    small_scale_cols_kernel<<<small_grid((int)rows, (int)cols), 128, 0, as_stream(stream)>>>(src, lds, (int)rows, (int)cols, s, dst, ldd);
    return check_launch();
}

}  // extern "C"
