// Small dense symmetric eigensolver on the device (no host LAPACK): parallel
// two-sided Jacobi, fp64.
//
// Used by Vectors.svd() (reference: cusolverDn?gesvd, dense_cublas.py:537-591)
// through the Gram route  S S^T = V diag(lambda) V^T, and by the Rayleigh-Ritz
// step of the device-resident driver.  p is the number of vectors in a block
// (2 x block size in the solver, up to ~1000 in PCA post-processing).
//
// One round of the round-robin ("circle") tournament rotates p/2 disjoint index
// pairs at once.  Because rotations on disjoint pairs commute, the whole
// similarity transform of a round factorises over 2x2 blocks:
//     A'[P_i, P_j] = R_i^T A[P_i, P_j] R_j        for every pair of pairs (P_i, P_j),
// so a round is: (1) rotation angles from the 2x2 diagonal blocks, (2) ONE pass in
// which every 2x2 block (and every 1x2 piece of the eigenvector matrix) is updated
// independently -- two barriers per round, no separate row and column passes.
//   p <= 64  : one CTA, A in shared memory, __syncthreads barriers
//   p  > 64  : cooperative multi-CTA launch, A in L2, grid-wide barriers
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace rl {

constexpr int EIG_THREADS = 1024;
constexpr int EIG_MAX_SWEEPS = 60;
constexpr int EIG_SMEM_MAX_P = 64;     // measured: p = 128 takes 12.6 ms in one CTA, less with the cooperative grid

// Round `t` of the circle method on P (even) players: pair i of P/2.
__device__ __forceinline__ void rr_pair(int P, int t, int i, int& a, int& b) {
    if (i == 0) { a = P - 1; b = t; }
    else { a = (t + i) % (P - 1); b = (t - i + P - 1) % (P - 1); }
    if (a > b) { int x = a; a = b; b = x; }
}

__device__ __forceinline__ void jacobi_cs(double app, double aqq, double apq, double& c, double& s) {
    c = 1.0; s = 0.0;
    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * sqrt(fabs(app * aqq))) {
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        c = 1.0 / sqrt(1.0 + t * t);
        s = t * c;
    }
}

// B <- R_i^T B R_j for the 2x2 block (rows pi,qi; cols pj,qj); R = [[c, s], [-s, c]]
__device__ __forceinline__ void rot_block(double& a, double& b, double& c_, double& d, double ci, double si,
                                          double cj, double sj) {
    const double t1 = a * cj - b * sj, t2 = a * sj + b * cj;
    const double t3 = c_ * cj - d * sj, t4 = c_ * sj + d * cj;
    a = ci * t1 - si * t3; c_ = si * t1 + ci * t3;
    b = ci * t2 - si * t4; d = si * t2 + ci * t4;
}

// ---- single-CTA kernel, A in shared memory -------------------------------------------
__global__ void __launch_bounds__(EIG_THREADS)
syevj_smem_kernel(double* __restrict__ Ag, int p, double* __restrict__ w, double* __restrict__ V,
                  int* __restrict__ perm, int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int P = (p + 1) & ~1, half = P / 2, lda = p | 1;      // odd leading dimension: conflict-free columns
    double* A = sm;                       // [p][lda]
    double* cs = sm + (size_t)p * lda;    // [half][2]
    __shared__ double s_red[2][EIG_THREADS / 32];
    __shared__ double s_off, s_diag;

    for (int e = tid; e < p * p; e += nt) { A[(e / p) * lda + e % p] = Ag[e]; V[e] = (e / p == e % p) ? 1.0 : 0.0; }
    __syncthreads();
    int sweep = 0;
    for (; sweep < EIG_MAX_SWEEPS; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < p * p; e += nt) {
            const double v = A[(e / p) * lda + e % p];
            if (e / p == e % p) dg += v * v; else off += v * v;
        }
        off = warp_sum(off); dg = warp_sum(dg);
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = off; s_red[1][tid >> 5] = dg; }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0;
            for (int i = 0; i < nt / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; }
            s_off = a; s_diag = b;
        }
        __syncthreads();
        // backward-stable stop: ||off||_F <= p * eps * ||A||_F (rounding in the rotations
        // re-creates off-diagonal noise of that size, so a tighter test never passes)
        const double tol = fmax(1e-15, 2.2e-16 * p);
        if (s_off <= tol * tol * (s_diag + s_off) || s_off == 0.0) break;
        if (!(s_off == s_off) || !(s_diag == s_diag)) { sweep = -1; break; }   // NaN input

        for (int t = 0; t < P - 1; ++t) {
            for (int i = tid; i < half; i += nt) {
                int a, b; rr_pair(P, t, i, a, b);
                double c = 1.0, s = 0.0;
                if (b < p) jacobi_cs(A[a * lda + a], A[b * lda + b], A[a * lda + b], c, s);
                cs[2 * i] = c; cs[2 * i + 1] = s;
            }
            __syncthreads();
            for (int e = tid; e < half * half; e += nt) {
                const int i = e / half, j = e - i * half;
                int pi, qi, pj, qj; rr_pair(P, t, i, pi, qi); rr_pair(P, t, j, pj, qj);
                const double ci = cs[2 * i], si = cs[2 * i + 1], cj = cs[2 * j], sj = cs[2 * j + 1];
                if (si == 0.0 && sj == 0.0) continue;
                const bool ri = qi < p, rj = qj < p;          // dummy player of an odd p
                double a = A[pi * lda + pj], b = rj ? A[pi * lda + qj] : 0.0;
                double c_ = ri ? A[qi * lda + pj] : 0.0, d = (ri && rj) ? A[qi * lda + qj] : 0.0;
                rot_block(a, b, c_, d, ci, si, cj, sj);
                A[pi * lda + pj] = a;
                if (rj) A[pi * lda + qj] = b;
                if (ri) A[qi * lda + pj] = c_;
                if (ri && rj) A[qi * lda + qj] = d;
            }
            for (int e = tid; e < half * p; e += nt) {        // V <- V J  (columns)
                const int j = e % half, row = e / half;
                int pj, qj; rr_pair(P, t, j, pj, qj);
                const double cj = cs[2 * j], sj = cs[2 * j + 1];
                if (sj == 0.0 || qj >= p) continue;
                const double x = V[row * p + pj], y = V[row * p + qj];
                V[row * p + pj] = cj * x - sj * y;
                V[row * p + qj] = sj * x + cj * y;
            }
            __syncthreads();
        }
    }
    // eigenvalues ascending (ties by index); eigenvectors as columns of Ag
    for (int i = tid; i < p; i += nt) {
        const double v = A[i * lda + i];
        int rank = 0;
        for (int j = 0; j < p; ++j) { const double u = A[j * lda + j]; rank += (u < v) || (u == v && j < i); }
        w[rank] = v;
        perm[i] = rank;
    }
    __syncthreads();
    for (int e = tid; e < p * p; e += nt) Ag[(e / p) * p + perm[e % p]] = V[e];
    if (tid == 0) *sweeps_out = sweep >= EIG_MAX_SWEEPS ? -2 : sweep;     // -2: sweeps exhausted without convergence
}

// ---- cooperative multi-CTA kernel, A and V in global memory (L2) ----------------------
__global__ void __launch_bounds__(256)
syevj_coop_kernel(double* __restrict__ A, int p, double* __restrict__ w, double* __restrict__ V,
                  double* __restrict__ cs, double* __restrict__ red, int* __restrict__ perm,
                  int* __restrict__ sweeps_out) {
    cg::grid_group grid = cg::this_grid();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gnt = (int64_t)gridDim.x * blockDim.x;
    const int P = (p + 1) & ~1, half = P / 2;
    __shared__ double s_a[8], s_b[8];

    for (int64_t e = gtid; e < (int64_t)p * p; e += gnt) V[e] = (e / p == e % p) ? 1.0 : 0.0;
    int sweep = 0;
    for (; sweep < EIG_MAX_SWEEPS; ++sweep) {
        // off-diagonal / diagonal norms: per-block partials in fixed slots, summed by everybody
        double off = 0.0, dg = 0.0;
        for (int64_t e = gtid; e < (int64_t)p * p; e += gnt) {
            const double v = __ldcg(A + e);
            if (e / p == e % p) dg += v * v; else off += v * v;
        }
        off = warp_sum(off); dg = warp_sum(dg);
        if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = off; s_b[threadIdx.x >> 5] = dg; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b = 0;
            for (int i = 0; i < (int)blockDim.x / 32; ++i) { a += s_a[i]; b += s_b[i]; }
            __stcg(red + 2 * blockIdx.x, a); __stcg(red + 2 * blockIdx.x + 1, b);
        }
        grid.sync();
        double toff = 0.0, tdg = 0.0;
        for (int i = 0; i < (int)gridDim.x; ++i) { toff += __ldcg(red + 2 * i); tdg += __ldcg(red + 2 * i + 1); }
        const double tol = fmax(1e-15, 2.2e-16 * p);
        if (toff <= tol * tol * (tdg + toff) || toff == 0.0) break;        // uniform: same sums everywhere
        if (!(toff == toff) || !(tdg == tdg)) { sweep = -1; break; }

        for (int t = 0; t < P - 1; ++t) {
            for (int64_t i = gtid; i < half; i += gnt) {
                int a, b; rr_pair(P, t, (int)i, a, b);
                double c = 1.0, s = 0.0;
                if (b < p) jacobi_cs(__ldcg(A + (int64_t)a * p + a), __ldcg(A + (int64_t)b * p + b),
                                     __ldcg(A + (int64_t)a * p + b), c, s);
                __stcg(cs + 2 * i, c); __stcg(cs + 2 * i + 1, s);
            }
            grid.sync();
            for (int64_t e = gtid; e < (int64_t)half * half; e += gnt) {
                const int i = (int)(e / half), j = (int)(e - (int64_t)i * half);
                int pi, qi, pj, qj; rr_pair(P, t, i, pi, qi); rr_pair(P, t, j, pj, qj);
                const double ci = __ldcg(cs + 2 * i), si = __ldcg(cs + 2 * i + 1);
                const double cj = __ldcg(cs + 2 * j), sj = __ldcg(cs + 2 * j + 1);
                if (si == 0.0 && sj == 0.0) continue;
                const bool ri = qi < p, rj = qj < p;
                double a = __ldcg(A + (int64_t)pi * p + pj), b = rj ? __ldcg(A + (int64_t)pi * p + qj) : 0.0;
                double c_ = ri ? __ldcg(A + (int64_t)qi * p + pj) : 0.0;
                double d = (ri && rj) ? __ldcg(A + (int64_t)qi * p + qj) : 0.0;
                rot_block(a, b, c_, d, ci, si, cj, sj);
                __stcg(A + (int64_t)pi * p + pj, a);
                if (rj) __stcg(A + (int64_t)pi * p + qj, b);
                if (ri) __stcg(A + (int64_t)qi * p + pj, c_);
                if (ri && rj) __stcg(A + (int64_t)qi * p + qj, d);
            }
            for (int64_t e = gtid; e < (int64_t)half * p; e += gnt) {
                const int j = (int)(e % half), row = (int)(e / half);
                int pj, qj; rr_pair(P, t, j, pj, qj);
                const double cj = __ldcg(cs + 2 * j), sj = __ldcg(cs + 2 * j + 1);
                if (sj == 0.0 || qj >= p) continue;
                const double x = __ldcg(V + (int64_t)row * p + pj), y = __ldcg(V + (int64_t)row * p + qj);
                __stcg(V + (int64_t)row * p + pj, cj * x - sj * y);
                __stcg(V + (int64_t)row * p + qj, sj * x + cj * y);
            }
            grid.sync();
        }
    }
    grid.sync();
    for (int64_t i = gtid; i < p; i += gnt) {
        const double v = __ldcg(A + i * p + i);
        int rank = 0;
        for (int j = 0; j < p; ++j) { const double u = __ldcg(A + (int64_t)j * p + j); rank += (u < v) || (u == v && j < i); }
        w[rank] = v;
        __stcg(perm + i, rank);
    }
    grid.sync();
    for (int64_t e = gtid; e < (int64_t)p * p; e += gnt) A[(e / p) * p + __ldcg(perm + e % p)] = __ldcg(V + e);
    if (gtid == 0) *sweeps_out = sweep >= EIG_MAX_SWEEPS ? -2 : sweep;
}

}  // namespace rl

using namespace rl;

extern "C" {

size_t rl_syevj_ws_bytes(int64_t p) {
    if (p <= 0) return 0;
    int64_t P = (p + 1) & ~int64_t(1);
    // V (p*p) | cs (P) | partial norms (2 * max grid 1024) | perm (p ints) | sweeps
    return (size_t)(p * p + P + 2 * 1024 + 8) * sizeof(double) + (size_t)(p + 16) * sizeof(int);
}

int rl_syevj(double* a, int64_t p, double* w, void* ws, size_t ws_bytes, int* sweeps_out_h, void* stream) {
    if (p < 0 || p > 4096) return RL_E_ARG;
    if (p == 0) return 0;
    if (ws_bytes < rl_syevj_ws_bytes(p)) return RL_E_WORKSPACE;
    int64_t P = (p + 1) & ~int64_t(1);
    double* V = (double*)ws;
    double* cs = V + p * p;
    double* red = cs + P;
    int* perm = (int*)(red + 2 * 1024 + 8);
    int* sweeps_d = perm + p + 4;
    cudaStream_t st = as_stream(stream);
    int rc;
    {
        Span span(PK_SYEVJ, st, 2.0 * p * p * 8, 0.0);
        if (p <= EIG_SMEM_MAX_P) {
            const int lda = (int)p | 1;
            const size_t smem = ((size_t)p * lda + P) * sizeof(double);
            static size_t configured = 0;
            if (smem > 48 * 1024 && smem > configured) {
                RL_CUDA(cudaFuncSetAttribute(syevj_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured = smem;
            }
            syevj_smem_kernel<<<1, EIG_THREADS, smem, st>>>(a, (int)p, w, V, perm, sweeps_d);
            rc = check_launch();
        } else {
            // as many CTAs as there are 2x2 blocks to update per round, capped by co-residency
            int per_sm = 0;
            RL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, syevj_coop_kernel, 256, 0));
            int64_t want = ((P / 2) * (P / 2) + 255) / 256;
            int64_t cap = (int64_t)per_sm * sm_count();
            if (cap > 1024) cap = 1024;
            int grid = (int)(want < cap ? want : cap);
            if (grid < 1) grid = 1;
            int pi = (int)p;
            void* args[] = {&a, &pi, &w, &V, &cs, &red, &perm, &sweeps_d};
            rc = (int)cudaLaunchCooperativeKernel((void*)syevj_coop_kernel, dim3(grid), dim3(256), args, 0, st);
            ++g_launches;
        }
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
