// Small dense symmetric eigensolver on the device (no host LAPACK): parallel
// cyclic Jacobi with round-robin pair ordering, one CTA, fp64.
//
// Used by Vectors.svd() (reference: cusolverDn?gesvd, dense_cublas.py:537-591)
// through the Gram route  S S^T = V diag(lambda) V^T, and by the Rayleigh-Ritz
// step of the device-resident driver.  p is the number of vectors in a block
// (<= a few hundred in the solver, up to ~1000 in PCA post-processing), so the
// matrix lives in L2; the kernel is latency-, not bandwidth-bound.
#include "common.cuh"

namespace rl {

constexpr int EIG_THREADS = 1024;
constexpr int EIG_MAX_SWEEPS = 60;

// workspace layout: V (p*p doubles) | cs (P doubles: c,s per pair) | top/bot (P ints) | flags
__global__ void __launch_bounds__(EIG_THREADS)
syevj_kernel(double* __restrict__ A, int p, double* __restrict__ w, double* __restrict__ V,
             double* __restrict__ cs, int* __restrict__ order, int* __restrict__ sweeps_out) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int P = (p + 1) & ~1;          // padded to even; index >= p is a dummy player
    const int half = P / 2;
    int* top = order;
    int* bot = order + half;
    __shared__ double s_off, s_diag;
    __shared__ double s_red[EIG_THREADS / 32];

    for (int e = tid; e < p * p; e += nt) V[e] = (e / p == e % p) ? 1.0 : 0.0;
    for (int i = tid; i < half; i += nt) { top[i] = i; bot[i] = P - 1 - i; }
    __syncthreads();

    int sweep = 0;
    for (; sweep < EIG_MAX_SWEEPS; ++sweep) {
        // convergence: off-diagonal Frobenius norm against the diagonal's
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < p * p; e += nt) {
            double v = A[e];
            if (e / p == e % p) dg += v * v; else off += v * v;
        }
        off = warp_sum(off); dg = warp_sum(dg);
        if ((tid & 31) == 0) s_red[tid >> 5] = off;
        __syncthreads();
        if (tid == 0) { double t = 0; for (int i = 0; i < nt / 32; ++i) t += s_red[i]; s_off = t; }
        __syncthreads();
        if ((tid & 31) == 0) s_red[tid >> 5] = dg;
        __syncthreads();
        if (tid == 0) { double t = 0; for (int i = 0; i < nt / 32; ++i) t += s_red[i]; s_diag = t; }
        __syncthreads();
        // backward-stable stop: ||off||_F <= p * eps * ||A||_F (rounding in the rotations
        // re-creates off-diagonal noise of that size, so a tighter test never passes)
        const double tol = fmax(1e-15, 2.2e-16 * p);
        if (s_off <= tol * tol * (s_diag + s_off) || s_off == 0.0) break;
        if (!(s_off == s_off) || !(s_diag == s_diag)) { sweep = -1; break; }   // NaN input

        for (int round = 0; round < P - 1; ++round) {
            // rotation angles for the disjoint pairs of this round
            for (int i = tid; i < half; i += nt) {
                int a = top[i], b = bot[i];
                int pp = a < b ? a : b, qq = a < b ? b : a;
                double c = 1.0, s = 0.0;
                if (qq < p) {
                    double apq = A[pp * p + qq];
                    double app = A[pp * p + pp], aqq = A[qq * p + qq];
                    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq))) {
                        double tau = (aqq - app) / (2.0 * apq);
                        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                cs[2 * i] = c; cs[2 * i + 1] = s;
            }
            __syncthreads();
            // row rotations: rows pp, qq of A  (A <- J^T A), coalesced along columns
            for (int e = tid; e < half * p; e += nt) {
                int i = e / p, col = e % p;
                int a = top[i], b = bot[i];
                int pp = a < b ? a : b, qq = a < b ? b : a;
                if (qq >= p) continue;
                double c = cs[2 * i], s = cs[2 * i + 1];
                if (s == 0.0) continue;
                double x = A[pp * p + col], y = A[qq * p + col];
                A[pp * p + col] = c * x - s * y;
                A[qq * p + col] = s * x + c * y;
            }
            __syncthreads();
            // column rotations: columns pp, qq of A (A <- A J) and of V (V <- V J)
            for (int e = tid; e < half * p; e += nt) {
                int i = e % half, row = e / half;
                int a = top[i], b = bot[i];
                int pp = a < b ? a : b, qq = a < b ? b : a;
                if (qq >= p) continue;
                double c = cs[2 * i], s = cs[2 * i + 1];
                if (s == 0.0) continue;
                double x = A[row * p + pp], y = A[row * p + qq];
                A[row * p + pp] = c * x - s * y;
                A[row * p + qq] = s * x + c * y;
                x = V[row * p + pp]; y = V[row * p + qq];
                V[row * p + pp] = c * x - s * y;
                V[row * p + qq] = s * x + c * y;
            }
            __syncthreads();
            // rotate the tournament: top[0] fixed, others move round-robin
            if (tid == 0 && half > 1) {
                int last_top = top[half - 1];
                int first_bot = bot[0];
                for (int i = half - 1; i > 1; --i) top[i] = top[i - 1];
                top[1] = first_bot;
                for (int i = 0; i < half - 1; ++i) bot[i] = bot[i + 1];
                bot[half - 1] = last_top;
            }
            __syncthreads();
        }
    }
    // eigenvalues = diagonal; rank them ascending (ties by index) and scatter
    for (int i = tid; i < p; i += nt) {
        double v = A[i * p + i];
        int rank = 0;
        for (int j = 0; j < p; ++j) {
            double u = A[j * p + j];
            rank += (u < v) || (u == v && j < i);
        }
        w[rank] = v;
        order[P + i] = rank;     // reuse tail of `order` for the permutation
    }
    __syncthreads();
    // A <- V with columns permuted into ascending order
    for (int e = tid; e < p * p; e += nt) {
        int row = e / p, col = e % p;
        A[row * p + order[P + col]] = V[e];
    }
    if (tid == 0) *sweeps_out = sweep;
}

}  // namespace rl

using namespace rl;

extern "C" {

size_t rl_syevj_ws_bytes(int64_t p) {
    if (p <= 0) return 0;
    int64_t P = (p + 1) & ~int64_t(1);
    return (size_t)(p * p + P + 8) * sizeof(double) + (size_t)(2 * P + 8 + 2) * sizeof(int);
}

int rl_syevj(double* a, int64_t p, double* w, void* ws, size_t ws_bytes, int* sweeps_out_h, void* stream) {
    if (p < 0 || p > 4096) return RL_E_ARG;
    if (p == 0) return 0;
    if (ws_bytes < rl_syevj_ws_bytes(p)) return RL_E_WORKSPACE;
    int64_t P = (p + 1) & ~int64_t(1);
    double* V = (double*)ws;
    double* cs = V + p * p;
    int* order = (int*)(cs + P + 8);
    int* sweeps_d = order + 2 * P + 8;
    cudaStream_t st = as_stream(stream);
    int rc;
    {
        Span span(PK_SYEVJ, st, 2.0 * p * p * 8, 0.0);
        syevj_kernel<<<1, EIG_THREADS, 0, st>>>(a, (int)p, w, V, cs, order, sweeps_d);
        rc = check_launch();
    }
    if (rc) return rc;
    if (sweeps_out_h) {
        RL_CUDA(cudaMemcpyAsync(sweeps_out_h, sweeps_d, sizeof(int), cudaMemcpyDeviceToHost, st));
        RL_CUDA(cudaStreamSynchronize(st));
        if (*sweeps_out_h < 0) return RL_E_NOTCONV;      // NaN/Inf in the input
    }
    return 0;
}

}  // extern "C"
