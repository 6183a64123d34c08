// Rayleigh-Ritz step of the block Jacobi-CG iteration, on the device end to end
// (reference: host LAPACK, raleigh/core/solver.py:1456-1493 and :1589-1607).
//
//   G      = U^-T GA U^-1                      two triangular solves (_transform, :1685-1688)
//   Qy     = eigenvectors of G[nx:, nx:]       pre-rotation of the search directions (:1458-1463)
//   w, Q   = eigh(G)                           (:1470)
//   dX, dlmd                                   change estimates (:1475-1493)
//   Q      = U^-1 diag(I, Qy) Q                back-transformation (:1591-1592)
//   CX, CZ = column blocks of Q                coefficients of the new X and Z (:1593-1607)
//
// Host code here only sequences kernels of rr.cu / jacobi.cu / small.cu on the caller's stream.
#include "common.cuh"

extern "C" {
int rl_small_copy(const double*, int64_t, double*, int64_t, int64_t, int64_t, void*);
int rl_small_transpose(const double*, int64_t, double*, int64_t, int64_t, int64_t, void*);
int rl_small_mirror(double*, int64_t, int64_t, int64_t, void*);
int rl_small_gemm(int, int, int64_t, int64_t, int64_t, double, const double*, int64_t, const double*, int64_t,
                  double, double*, int64_t, void*);
int rl_small_trsm(int, const double*, int64_t, int64_t, double*, int64_t, int64_t, void*);
int rl_rr_estimates(const double*, int64_t, const double*, int64_t, int64_t, int64_t, int64_t, double*, double*,
                    void*);
int rl_rr_select(const double*, int64_t, const double*, int64_t, int64_t, int64_t, double*, int64_t, double*,
                 int64_t, double*, double*, void*);
int rl_syevj_cluster_max_n(void);
size_t rl_syevj_cluster_ws_bytes(int64_t);
int rl_syevj_cluster(const double*, int64_t, int64_t, int, double, double*, double*, int64_t, void*, size_t, int*, void*);
int rl_syevj_grid_max_n(void);
}

namespace rl {

// symmetric part into a contiguous n x n buffer
__global__ void symmetrize_kernel(const double* __restrict__ g, int64_t ld, int n, double* __restrict__ out) {
    for (int r = blockIdx.y; r < n; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
            out[(int64_t)r * n + c] = 0.5 * (g[(int64_t)r * ld + c] + g[(int64_t)c * ld + r]);
}

static size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// eigen-decomposition of sym(G) (n x n, ldg): w ascending, eigenvectors as columns of Q (ldq)
static int small_eigh(const double* G, int64_t ldg, int64_t n, double* w, double* Q, int64_t ldq, void* ws,
                      size_t ws_bytes, int* info_d, cudaStream_t st, int factor_mode = 0, double tol = 0.0) {
    if (n == 0) return 0;
    if (n <= rl_syevj_grid_max_n() && (g_knob[KNOB_EIG_LEGACY] == 0 || factor_mode))
        return rl_syevj_cluster(G, ldg, n, factor_mode, tol, w, Q, ldq, ws, ws_bytes, info_d, st);
    if (factor_mode) return RL_E_ARG;
    // large blocks: cooperative-grid two-sided kernel (small.cu) on a contiguous copy
    double* a = (double*)ws;
    const size_t off = align256((size_t)n * n * sizeof(double));
    if (ws_bytes < off + rl_syevj_ws_bytes(n)) return RL_E_WORKSPACE;
    int gx = (int)((n + 127) / 128);
    symmetrize_kernel<<<dim3((unsigned)gx, (unsigned)(n > 65535 ? 65535 : n)), 128, 0, st>>>(G, ldg, (int)n, a);
    int rc = check_launch();
    if (rc) return rc;
    rc = rl_syevj(a, n, w, (char*)ws + off, ws_bytes - off, nullptr, st);
    if (rc) return rc;
    return rl_small_copy(a, n, Q, ldq, n, n, st);
}

static size_t eigh_ws_bytes(int64_t n) {
    size_t a = rl_syevj_cluster_ws_bytes(n <= rl_syevj_grid_max_n() ? n : rl_syevj_grid_max_n());
    size_t b = align256((size_t)n * n * sizeof(double)) + rl_syevj_ws_bytes(n);
    return a > b ? a : b;
}

}  // namespace rl

using namespace rl;

extern "C" {

size_t rl_small_eigh_ws_bytes(int64_t n) { return n > 0 ? eigh_ws_bytes(n) : 0; }

/* w, Q = eigh(sym(g)); device pointers; info_d (2 ints: sweeps, converged) may be NULL */
int rl_small_eigh(const double* g, int64_t ldg, int64_t n, double tol, double* w, double* q, int64_t ldq, void* ws,
                  size_t ws_bytes, int* info_d, void* stream) {
    if (n < 0) return RL_E_ARG;
    if (ws_bytes < rl_small_eigh_ws_bytes(n)) return RL_E_WORKSPACE;
    return small_eigh(g, ldg, n, w, q, ldq, ws, ws_bytes, info_d, as_stream(stream), 0, tol);
}

/* w, Q = eigh(u^T u) from the upper Cholesky factor u (Jacobi on the factor: relative accuracy for every
 * eigenvalue); n <= rl_syevj_grid_max_n() */
int rl_small_eigh_factor(const double* u, int64_t ldu, int64_t n, double tol, double* w, double* q, int64_t ldq,
                         void* ws, size_t ws_bytes, int* info_d, void* stream) {
    if (n < 0 || n > rl_syevj_grid_max_n()) return RL_E_ARG;
    if (ws_bytes < rl_small_eigh_ws_bytes(n)) return RL_E_WORKSPACE;
    return small_eigh(u, ldu, n, w, q, ldq, ws, ws_bytes, info_d, as_stream(stream), 1, tol);
}

size_t rl_rr_solve_ws_bytes(int64_t nmax) {
    if (nmax <= 0) return 0;
    // W1, G, Q, Qy, T (n x n each) | w, wy (n each) | eigensolver workspace
    return 5 * align256((size_t)nmax * nmax * sizeof(double)) + 2 * align256((size_t)nmax * sizeof(double)) +
           eigh_ws_bytes(nmax) + 256;
}

/* ga: (nxy x nxy) A-Gram matrix of (X, Y), full symmetric; u: upper Cholesky factor of their B-Gram
 * matrix (rl_rr_piv_chol), both with leading dimension ld.  Outputs: cx (nxy x nxn), cz (nxy x nz),
 * lmdx (nxn), lmdz (nz), est = dX (nx) followed by dlmd at est + nmax.  All device memory.
 * eig_tol: stopping tolerance of the Jacobi eigensolver (<= 0: working precision). */
int rl_rr_solve(const double* ga, const double* u, int64_t ld, int64_t nx, int64_t ny, int64_t leftX,
                int64_t rightX, int64_t leftXn, int64_t rightXn, double* cx, int64_t ldcx, double* cz,
                int64_t ldcz, double* lmdx, double* lmdz, double* est, int64_t nmax, double eig_tol, void* ws,
                size_t ws_bytes, int* info_d, void* stream) {
    const int64_t n = nx + ny;
    if (nx < 0 || ny < 0 || n > nmax || leftX + rightX != nx || leftXn + rightXn > n) return RL_E_ARG;
    if (n == 0) return 0;
    if (ws_bytes < rl_rr_solve_ws_bytes(nmax)) return RL_E_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    Span span(PK_RR_SOLVE, st, 5.0 * n * n * 8, 8.0 * n * n * n);
    const size_t mat = align256((size_t)nmax * nmax * sizeof(double)), vecb = align256((size_t)nmax * sizeof(double));
    char* p = (char*)ws;
    double* W1 = (double*)p; p += mat;
    double* G = (double*)p; p += mat;
    double* Q = (double*)p; p += mat;
    double* Qy = (double*)p; p += mat;
    double* T = (double*)p; p += mat;
    double* w = (double*)p; p += vecb;
    double* wy = (double*)p; p += vecb;
    void* ews = p;
    const size_t ews_bytes = ws_bytes - (size_t)(p - (char*)ws);
    int rc;
#define RL_TRY(expr) do { rc = (expr); if (rc) return rc; } while (0)
    // G = U^-T GA U^-1
    RL_TRY(rl_small_copy(ga, ld, W1, n, n, n, st));
    RL_TRY(rl_small_trsm(0, u, ld, n, W1, n, n, st));
    RL_TRY(rl_small_transpose(W1, n, G, n, n, n, st));
    RL_TRY(rl_small_trsm(0, u, ld, n, G, n, n, st));
    if (ny > 0 && nx > 0) {
        // rotate the Y block to the eigenbasis of its own Rayleigh-Ritz problem
        RL_TRY(small_eigh(G + nx * n + nx, n, ny, wy, Qy, ny, ews, ews_bytes, nullptr, st, 0, eig_tol));
        RL_TRY(rl_small_gemm(0, 0, nx, ny, ny, 1.0, G + nx, n, Qy, ny, 0.0, T, ny, st));          // G[:nx, nx:] Qy
        RL_TRY(rl_small_copy(T, ny, G + nx, n, nx, ny, st));
        RL_TRY(rl_small_gemm(0, 0, ny, ny, ny, 1.0, G + nx * n + nx, n, Qy, ny, 0.0, T, ny, st)); // Gyy Qy
        RL_TRY(rl_small_gemm(1, 0, ny, ny, ny, 1.0, Qy, ny, T, ny, 0.0, G + nx * n + nx, n, st)); // Qy^T (Gyy Qy)
        RL_TRY(rl_small_mirror(G, n, nx, ny, st));
    } else if (ny > 0) {
        RL_TRY(small_eigh(G, n, ny, wy, Qy, ny, ews, ews_bytes, nullptr, st, 0, eig_tol));
        RL_TRY(rl_small_gemm(0, 0, ny, ny, ny, 1.0, G, n, Qy, ny, 0.0, T, ny, st));
        RL_TRY(rl_small_gemm(1, 0, ny, ny, ny, 1.0, Qy, ny, T, ny, 0.0, G, n, st));
    }
    RL_TRY(small_eigh(G, n, n, w, Q, n, ews, ews_bytes, info_d, st, 0, eig_tol));
    if (nx > 0) RL_TRY(rl_rr_estimates(Q, n, w, nx, ny, leftX, rightX, est, est + nmax, st));
    if (ny > 0) {
        RL_TRY(rl_small_gemm(0, 0, ny, n, ny, 1.0, Qy, ny, Q + nx * n, n, 0.0, T, n, st));        // Qy Q[nx:, :]
        RL_TRY(rl_small_copy(T, n, Q + nx * n, n, ny, n, st));
    }
    RL_TRY(rl_small_trsm(1, u, ld, n, Q, n, n, st));
    RL_TRY(rl_rr_select(Q, n, w, n, leftXn, rightXn, cx, ldcx, cz, ldcz, lmdx, lmdz, st));
#undef RL_TRY
    return 0;
}

}  // extern "C"
