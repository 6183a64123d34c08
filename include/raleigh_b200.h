/*
 * raleigh_b200 -- C ABI of the B200 (sm_100a) block-vector algebra library.
 *
 * This is the drop-in boundary for the hot path of RALEIGH's block
 * Jacobi-conjugated-gradient eigensolver: the abstract `Vectors` algebra, the
 * dense `Matrix` operator and the sparse symmetric operator.  Each entry point
 * replaces one vendor-library call site of the reference (cuBLAS / cuSOLVER in
 * raleigh/algebra/dense_cublas.py, MKL in raleigh/algebra/mkl_wrap.py); the
 * reference line it stands in for is cited next to every declaration.
 *
 * Conventions
 *   - plain C types only; every function returns int: 0 = ok, >0 = cudaError_t,
 *     <0 = RL_E_* argument error.  rl_error_string() decodes both.
 *   - block vectors are VECTOR-MAJOR: vector j, component r lives at
 *     base[j*ld + r] (the reference's C-ordered (nvec, n) array, ld >= n).
 *     "m" counts vectors of `self`/output, "k" vectors of `other`/input,
 *     "n" is the vector dimension (rows of the eigenproblem).
 *   - all pointers are DEVICE pointers unless the name ends in _h (host).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     Work is enqueued asynchronously; *_h entry points that return data to
 *     the host synchronise the stream before returning.
 *   - dtype: RL_F32 / RL_F64 (complex types of the reference are not built).
 */
#ifndef RALEIGH_B200_H
#define RALEIGH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RL_F32 = 0, RL_F64 = 1 };

enum {
    RL_E_DTYPE = -1,      /* unsupported dtype code */
    RL_E_ARG = -2,        /* negative size / null pointer / bad stride */
    RL_E_WORKSPACE = -3,  /* workspace too small */
    RL_E_ALIAS = -4,      /* output aliases an input where that is not allowed */
    RL_E_NOTCONV = -5     /* small dense eigensolver did not converge */
};

/* ---- library / device -------------------------------------------------- */
int rl_version(void);
const char* rl_error_string(int rc);
/* cuda_wrap.py:139-141 (getDeviceCount / getDeviceProperties / synchronize) */
int rl_device_count(int* count);
int rl_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                   size_t* l2_bytes, size_t* total_mem);
int rl_sync_device(void);
int rl_sync_stream(void* stream);
/* counts kernels this library has launched since load (bench.py gpu_launches) */
int64_t rl_launch_count(void);

/* Optional per-entry-point device timing (CUDA events on the caller's stream
 * around the device work of each call).  The reference only has wall-clock
 * timers (partial_svd.py:261,290-291); bench.py reads these for the roofline.
 * kinds 0..rl_profile_kinds()-1 are named by rl_profile_name(). */
void rl_profile_enable(int on);
void rl_profile_reset(void);
int rl_profile_kinds(void);
const char* rl_profile_name(int kind);
int rl_profile_get(int kind, int64_t* count, double* ms, double* bytes, double* flops);

/* ---- raw memory (cuda_wrap.py:142-153: malloc/free/memset/memcpy/memcpy2D;
 *      size_t sizes instead of the reference's c_int) ---------------------- */
int rl_malloc(void** ptr, size_t bytes);
int rl_free(void* ptr);
int rl_memset(void* ptr, int value, size_t bytes, void* stream);
int rl_h2d(void* dst, const void* src_h, size_t bytes, void* stream);
int rl_d2h(void* dst_h, const void* src, size_t bytes, void* stream);
/* strided host<->device block copies; widths/pitches in BYTES */
int rl_h2d_2d(void* dst, size_t dpitch, const void* src_h, size_t spitch,
              size_t width, size_t height, void* stream);
int rl_d2h_2d(void* dst_h, size_t dpitch, const void* src, size_t spitch,
              size_t width, size_t height, void* stream);

/* ---- Vectors: data movement --------------------------------------------- */
/* Vectors.copy(other)  dense_cublas.py:133-145  (cublas?copy over m*n) */
int rl_copy(int dtype, void* dst, int64_t ld_dst, const void* src, int64_t ld_src,
            int64_t m, int64_t n, void* stream);
/* Vectors.copy(other, ind)  dense_cublas.py:146-153 (one cublas?copy per index):
 * dst[t] <- src_all[ind_h[t]], t < count; ind_h are absolute vector indices. */
int rl_gather(int dtype, void* dst, int64_t ld_dst, const void* src_all,
              int64_t ld_src, const int64_t* ind_h, int64_t count, int64_t n,
              void* stream);
/* Vectors.fill_random device mode (reference: host numpy.random.rand + H2D,
 * dense_cublas.py:119-131).  Counter-based: element (j0+j, r0+r) depends only on
 * (seed, global vector index, global row index), so shards reproduce the
 * unsharded fill.  Values uniform in (-1, 1). */
int rl_fill_uniform(int dtype, void* x, int64_t ld, int64_t m, int64_t n,
                    uint64_t seed, int64_t j0, int64_t r0, void* stream);

/* ---- Vectors: BLAS-1 like ---------------------------------------------- */
/* Vectors.add(other, s) scalar  dense_cublas.py:311-316 (cublas?axpy) : Y += alpha*X */
int rl_axpy(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx,
            int64_t m, int64_t n, double alpha, void* stream);
/* Vectors.add(other, s[]) dense_cublas.py:343-350 (m cublas?axpy calls): Y[i] += s[i]*X[i] */
int rl_axpy_diag(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx,
                 int64_t m, int64_t n, const void* s, void* stream);
int rl_axpy_diag_h(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx,
                   int64_t m, int64_t n, const void* s_h, void* stream);
/* Vectors.scale(s, multiply) dense_cublas.py:155-172 (m cublas?scal calls):
 * multiply != 0: Y[i] *= s[i]; else Y[i] /= s[i] where s[i] != 0 */
int rl_scale(int dtype, void* y, int64_t ldy, int64_t m, int64_t n,
             const void* s, int multiply, void* stream);
int rl_scale_h(int dtype, void* y, int64_t ldy, int64_t m, int64_t n,
               const void* s_h, int multiply, void* stream);
/* Vectors.dots(other) dense_cublas.py:222-243 (m cublas?dot calls):
 * w[i] = sum_r O[i,r]*S[i,r].  Deterministic two-phase reduction. */
size_t rl_dots_ws_bytes(int dtype, int64_t m, int64_t n);
int rl_dots(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo,
            int64_t m, int64_t n, void* w, void* ws, size_t ws_bytes, void* stream);
int rl_dots_h(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo,
              int64_t m, int64_t n, void* w_h, void* stream);
/* Vectors.dots(other, transp=True) dense_cublas.py:175-221 (gemmBatched of n 1x1):
 * w[r] = sum_i O[i,r]*S[i,r], length n */
int rl_dots_t(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo,
              int64_t m, int64_t n, void* w, void* stream);
/* Jacobi preconditioner through Operator.apply (sparse_mkl.py:143-154):
 * Y[i,r] = X[i,r]*d[r]  (d = 1/diag(A), length n) */
int rl_diag_mul(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx,
                int64_t m, int64_t n, const void* d, void* stream);

/* ---- Vectors: tall-skinny BLAS-3 like ------------------------------------ */
/* Vectors.dot(other) dense_cublas.py:245-269 (cublas?gemm T,N + D2H):
 * G[i*m + j] = sum_r O[i,r]*S[j,r]  -- (k, m) row-major, exactly the ndarray
 * the reference returns.  Deterministic: per-CTA partials in fixed slots of
 * `ws`, then a fixed-order tree; no floating-point atomics. */
size_t rl_gram_ws_bytes(int dtype, int64_t m, int64_t k, int64_t n);
int rl_gram(int dtype, const void* s, int64_t lds, int64_t m, const void* o,
            int64_t ldo, int64_t k, int64_t n, void* g, void* ws, size_t ws_bytes,
            void* stream);
int rl_gram_h(int dtype, const void* s, int64_t lds, int64_t m, const void* o,
              int64_t ldo, int64_t k, int64_t n, void* g_h, void* stream);
/* Same product with fp64 accumulation and an fp64 (k, m) result whatever the
 * data dtype: the Gram matrix the on-device SVD / orthonormalisation factorises
 * (squaring the condition number in fp32 would lose the small singular values). */
size_t rl_gram_acc64_ws_bytes(int dtype, int64_t m, int64_t k, int64_t n);
int rl_gram_acc64(int dtype, const void* s, int64_t lds, int64_t m, const void* o,
                  int64_t ldo, int64_t k, int64_t n, double* g, void* ws,
                  size_t ws_bytes, void* stream);
/* test hooks for A/B runs: bit 0 forces the FMA-pipe Gram kernel for fp64, bit 1 the
 * two-warps-per-tile DMMA variant; rl_debug_set_update_fma forces the FMA-pipe update */
void rl_debug_set_gram_simt(int on);
void rl_debug_set_update_fma(int on);
void rl_debug_set_spmm_warps(int warps);
/* generic A/B knobs for measurements (value 0 = library default).  knob 0: fp64 Gram kernel (-1 register
 * fragments, 1 / 2 TMA ring with one / two CTAs per SM, 3 persistent with dynamic row chunks); 1: SpMM
 * shared-memory carve-out in percent; 2: SpMM register budget (16 or 24 resident warps per SM); 3: SpMM L2
 * prefetch mode (-1 off, bit mask); 4: interleaved row steps in the register Gram kernel; 5: Gram CTAs per SM
 * slot / chunks per CTA.  tools/sweep_r1e.py drives them; results in profiles/r1e_gram_spmm.md. */
void rl_debug_set_knob(int knob, int value);
int rl_debug_get_knob(int knob);
/* Vectors.multiply(q, out) dense_cublas.py:271-299 (gemm, beta=0) and
 * Vectors.add(other, s, q) dense_cublas.py:317-342 (gemm, beta=1):
 * Out[j,:] = beta*Out[j,:] + alpha * sum_{i<k} Q[i*q_rs + j*q_cs] * X[i,:], j < m.
 * Q strides are in elements (C-, F-ordered or strided views all map here).
 * Out must not overlap X. */
int rl_update(int dtype, void* out, int64_t ldo, int64_t m, const void* x,
              int64_t ldx, int64_t k, const void* q, int64_t q_rs, int64_t q_cs,
              double alpha, double beta, int64_t n, void* stream);
int rl_update_h(int dtype, void* out, int64_t ldo, int64_t m, const void* x,
                int64_t ldx, int64_t k, const void* q_h, int64_t q_rs, int64_t q_cs,
                double alpha, double beta, int64_t n, void* stream);

/* ---- dense Matrix operator ---------------------------------------------- */
/* Matrix.apply(x, y, transp) dense_cublas.py:732-776 (cublas?gemm):
 * A is (M, N) row-major with leading dimension lda;
 *   transp == 0: Y[v,i] = sum_j X[v,j]*A[i,j]   (x dim N -> y dim M)
 *   transp != 0: Y[v,j] = sum_i X[v,i]*A[i,j]   (x dim M -> y dim N)
 * Y = alpha*result + beta*Y.  The fp32 path uses tcgen05 tensor cores with a
 * 3xTF32 split when the shape allows, else an fp32 FMA kernel. */
int rl_dense_apply(int dtype, const void* a, int64_t lda, int64_t M, int64_t N,
                   const void* x, int64_t ldx, void* y, int64_t ldy, int64_t k,
                   int transp, double alpha, double beta, void* stream);

/* Tensor-core path of the same product for fp32 (tcgen05.mma kind::tf32, TMEM
 * accumulators, TMA-fed), fp32-accurate through the 3xTF32 split
 * X.A ~= Xhi.Ahi + Xhi.Alo + Xlo.Ahi.  a_lo == NULL (default): the low parts of
 * both operands are computed in shared memory by four extra warps, HBM traffic
 * is the data matrix once.  a_lo != NULL: materialised low part of the data
 * matrix from rl_split_tf32 (same shape and lda as `a`; A/B variant, streams the
 * matrix twice); the block's low part is then built per call inside `ws`.
 * Requirements: 16-byte aligned bases, lda and ldx multiples of 4
 * (rl_dense_apply_tc_supported); any M, N, k. */
int rl_dense_apply_tc_supported(const void* a, int64_t lda, const void* x, int64_t ldx);
size_t rl_dense_apply_tc_ws_bytes(int64_t M, int64_t N, int64_t k, int transp);
int rl_split_tf32(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                  int64_t cols, void* stream);
int rl_dense_apply_tc(const void* a, const void* a_lo, int64_t lda, int64_t M, int64_t N,
                      const void* x, int64_t ldx, void* y, int64_t ldy, int64_t k, int transp,
                      double alpha, double beta, void* ws, size_t ws_bytes, void* stream);

/* AMatrix.__init__ dense_matrix.py:32-34 scans the host array with numpy.amin / numpy.amax for
 * scale(); the same two numbers from the device copy in one HBM-bound pass (host scalars of the
 * block's dtype; NaNs are ignored where NumPy would propagate them). */
int rl_minmax_h(int dtype, const void* x, int64_t ld, int64_t m, int64_t n, void* min_h, void* max_h,
                void* stream);

/* ---- sparse symmetric operator ------------------------------------------- */
/* SparseSymmetricMatrix.apply sparse_mkl.py:42-48 -> mkl_?csrmm mkl_wrap.py:274-276
 * (and mkl_?csrsymv :261-262 for m == 1).  The device holds the FULL symmetric
 * matrix as 0-based CSR (indptr int64, indices int32).
 * Y[v, r] = sum_{p in row r} val[p] * X[v, col[p]],  rows [row0, row0+nrows) of
 * the operator write Y[v, 0..nrows) (row-sharded use: X is the gathered vector
 * of dimension ncols). */
int rl_csr_spmm(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr,
                const int32_t* indices, const void* values, const void* x, int64_t ldx,
                void* y, int64_t ldy, int64_t m, void* stream);

/* Extended form: `run_order` (device, ceil(nrows/32) entries, or NULL = identity) maps CTA slots to
 * 32-row runs -- see rl_spmm_cluster_runs; `warps` = runs per CTA (4, 8 or 16; 0 = default);
 * ncols_local/halo as in rl_csr_spmm_halo (0/NULL when the operator is not sharded). */
int rl_csr_spmm_ex(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                   const void* values, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m,
                   int64_t ncols_local, const void* halo, const int32_t* run_order, int warps,
                   void* stream);
/* HOST-side set-up (no device work; all pointers are host arrays): a permutation of the 32-row
 * runs of the CSR matrix such that every `group` consecutive entries (the runs one CTA of
 * rl_csr_spmm_ex processes together) share as much of their column footprint as possible, so the
 * X lines gathered for one warp are L1 hits for the others.  footprint_ratio_out (optional):
 * distinct 32-column segments gathered per CTA, summed, divided by the number of runs (1 = every
 * X line enters an SM once). */
int rl_spmm_cluster_runs(int64_t nrows, const int64_t* indptr_h, const int32_t* indices_h, int group,
                         int32_t* order_out_h, double* footprint_ratio_out);

/* Same product from the SELL-32 layout (sliced ELLPACK, 32-row slices): entry j
 * of row 32*s + l is stored at slice_ptr[s] + 32*j + l; padding entries have
 * val = 0 and any valid column.  `nnz` is the number of true entries (profiling
 * only).  Coalesced matrix reads without staging; preferred when the padding
 * overhead is small (SparseSymmetricMatrix decides at construction). */
int rl_sell_spmm(int dtype, int64_t nrows, int64_t nnz, int64_t nslices, const int64_t* slice_ptr,
                 const int32_t* cols, const void* vals, const void* x, int64_t ldx, void* y,
                 int64_t ldy, int64_t m, void* stream);

/* Row-sharded operator (one process per GPU): the local CSR/SELL holds this rank's
 * rows with columns renumbered [owned 0..ncols_local) | halo ncols_local..); owned columns
 * are read from the local block x, halo columns from `halo`, the buffer received from the
 * peers, ROW-INTERLEAVED: halo[(c - ncols_local)*m + v].  rl_pack_rows builds the matching
 * send buffer out[t*m + v] = x[v, idx[t]] (idx on the device).  The exchange itself is an
 * NCCL all-to-all driven from the host side (raleigh_b200/sparse.py). */
int rl_pack_rows(int dtype, const void* x, int64_t ldx, int64_t m, const int64_t* idx, int64_t count,
                 void* out, void* stream);
int rl_csr_spmm_halo(int dtype, int64_t nrows, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                     const void* values, const void* x, int64_t ldx, void* y, int64_t ldy, int64_t m,
                     int64_t ncols_local, const void* halo, void* stream);
int rl_sell_spmm_halo(int dtype, int64_t nrows, int64_t nnz, int64_t nslices, const int64_t* slice_ptr,
                      const int32_t* cols, const void* vals, const void* x, int64_t ldx, void* y,
                      int64_t ldy, int64_t m, int64_t ncols_local, const void* halo, void* stream);

/* ---- small dense algebra on device (no host LAPACK) ----------------------- */
/* Symmetric eigen-decomposition of the (p, p) row-major matrix `a` (fp64) by
 * cyclic Jacobi rotations: on return w[0..p) ascending eigenvalues and a holds
 * the eigenvectors as COLUMNS (a[i*p + j] = component i of eigenvector j).
 * Used by Vectors.svd (dense_cublas.py:537-591 cusolverDn?gesvd) via the Gram
 * route and by the device Rayleigh-Ritz.  ws >= rl_syevj_ws_bytes(p). */
size_t rl_syevj_ws_bytes(int64_t p);
int rl_syevj(double* a, int64_t p, double* w, void* ws, size_t ws_bytes,
             int* sweeps_out_h, void* stream);


/* ---- device-resident Rayleigh-Ritz (no host LAPACK) ------------
 * Small matrices are fp64, row-major with a leading dimension, and live in device memory for
 * the whole solve; nothing below synchronises.  The reference does all of this on the host
 * with NumPy / SciPy between two block-vector operations (raleigh/core/solver.py:838-1663). */
/* Vectors.dot without the D2H copy (dense_cublas.py:245-269): out[i*ldout + j] = <o_i, s_j>,
 * fp64 whatever the data type (fp32 data accumulate in fp64) */
int rl_gram_dev(int dtype, const void* s, int64_t lds, int64_t m, const void* o, int64_t ldo,
                int64_t k, int64_t n, double* out, int64_t ldout, void* stream);
/* Vectors.dots into a device fp64 vector (dense_cublas.py:222-243) */
int rl_dots_dev(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m,
                int64_t n, double* out, void* stream);
/* Vectors.multiply / add(other, s, q) with device-resident fp64 coefficients q (k x m, ldq)
 * (dense_cublas.py:271-342 without the H2D copy): Out = beta Out + alpha q^T X */
int rl_update_dev(int dtype, void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx,
                  int64_t k, const double* q, int64_t ldq, double alpha, double beta, int64_t n,
                  void* stream);
/* residuals W[j] = AX[j] - lmd[j] X[j] (solver.py:946-952: copy + per-vector axpy), lmd on the device */
int rl_residual_dev(int dtype, void* w, int64_t ldw, const void* ax, int64_t ldax, const void* x,
                    int64_t ldx, int64_t m, int64_t n, const double* lmd, void* stream);
/* Y[j] /= sqrt(|s2[j]|) unless zero (solver.py:1377-1378: sqrt on the host + Vectors.scale) */
int rl_scale_rsqrt_dev(int dtype, void* y, int64_t ldy, int64_t m, int64_t n, const double* s2,
                       void* stream);
/* small-matrix utilities: dst = src; dst = src^T; G[nx+j][i] = G[i][nx+j]; C = alpha op(A) op(B) + beta C */
int rl_small_copy(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows,
                  int64_t cols, void* stream);
int rl_small_transpose(const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows,
                       int64_t cols, void* stream);
int rl_small_mirror(double* g, int64_t ld, int64_t nx, int64_t ny, void* stream);
int rl_small_gemm(int transa, int transb, int64_t M, int64_t N, int64_t K, double alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                  int64_t ldc, void* stream);
/* mode 0: solve U^T X = B, mode 1: solve U X = B; U upper triangular n x n, B n x r overwritten
 * (scipy.linalg.solve_triangular at solver.py:1592, 1686-1687) */
int rl_small_trsm(int mode, const double* U, int64_t ldu, int64_t n, double* B, int64_t ldb,
                  int64_t r, void* stream);
/* Rayleigh quotients and restart indicators (solver.py:859-873): lmd[i] = xax[i][i] / xbx[i][i],
 * stats[0] = max|lmd - lmdx| / max|lmdx|, stats[1] = max|xbx - I| */
int rl_rr_ritz_check(const double* xax, const double* xbx, int64_t ld, int64_t nx,
                     const double* lmdx, double* lmd, double* stats, void* stream);
/* conjugation coefficients (solver.py:1331-1347): beta = num / den unless |num| >= 100 s |den| */
int rl_rr_conjugation(const double* zay, const double* zby, double* beta, int64_t ld, int64_t nz,
                      int64_t ny, const double* lmd, const double* lmdz, const double* sy2,
                      const double* sz2, void* stream);
/* _piv_chol (solver.py:1749-1845): in-place upper Cholesky factor of the leading n x n block,
 * first k columns unpivoted, max-diagonal pivoting and the reference's drop rule on the rest.
 * a0: scratch of the same shape; ind: n ints (permutation); info: 4 ints (dropped, status,
 * drop case, last condition check) -- all device memory. */
int rl_rr_piv_chol(double* a, double* a0, int64_t ld, int64_t n, int64_t k, double eps, int* ind,
                   int* info, void* stream);
/* change estimates (solver.py:1475-1493) and coefficient blocks (solver.py:1593-1607) */
int rl_rr_estimates(const double* q, int64_t ldq, const double* w, int64_t nx, int64_t ny,
                    int64_t leftX, int64_t rightX, double* dX, double* dlmd, void* stream);
int rl_rr_select(const double* q, int64_t ldq, const double* w, int64_t nxy, int64_t leftXn,
                 int64_t rightXn, double* cx, int64_t ldcx, double* cz, int64_t ldcz, double* lmdx,
                 double* lmdz, void* stream);
/* scipy.linalg.eigh (solver.py:1459, 1470) on the device.  n <= rl_syevj_cluster_max_n():
 * one-sided Jacobi on a thread-block cluster, columns in distributed shared memory
 * (csrc/jacobi.cu); larger: the cooperative-grid kernel behind rl_syevj.  w ascending,
 * q[i*ldq + j] = component i of eigenvector j; g is read only (both triangles, averaged). */
int rl_syevj_cluster_max_n(void);
/* orders up to rl_syevj_grid_max_n() run the same method on the whole GPU (cooperative launch,
 * columns in L2, one warp per pair); factor_mode != 0: g is the upper Cholesky factor U of the
 * matrix to decompose and the rotations act on its rows (no shift, relative accuracy);
 * tol: pairs rotate while |<p, q>| > tol |p| |q| (<= 0: sqrt(n) eps) */
int rl_syevj_grid_max_n(void);
size_t rl_syevj_cluster_ws_bytes(int64_t n);
int rl_syevj_cluster(const double* g, int64_t ldg, int64_t n, int factor_mode, double tol, double* w,
                     double* q, int64_t ldq, void* ws, size_t ws_bytes, int* info_d, void* stream);
size_t rl_small_eigh_ws_bytes(int64_t n);
int rl_small_eigh(const double* g, int64_t ldg, int64_t n, double tol, double* w, double* q,
                  int64_t ldq, void* ws, size_t ws_bytes, int* info_d, void* stream);
int rl_small_eigh_factor(const double* u, int64_t ldu, int64_t n, double tol, double* w, double* q,
                         int64_t ldq, void* ws, size_t ws_bytes, int* info_d, void* stream);
/* unpivoted Cholesky factorisation G = U^T U of an n x n fp64 matrix in place (upper factor, zeros
 * below; numpy.linalg.cholesky at partial_svd.py:192), blocked: diagonal blocks in shared memory,
 * panels by rl_small_trsm, trailing update by rl_small_gemm.  info_d[0] = 0 or 1 + index of the
 * first non-positive pivot. */
int rl_small_potrf(double* a, int64_t ld, int64_t n, int* info_d, void* stream);
/* the whole Rayleigh-Ritz step (solver.py:1456-1493, 1589-1607): transform, pre-rotation, eigh,
 * estimates, back-transformation, coefficient blocks.  est: dX at est[0..nx), dlmd at est[nmax..);
 * eig_tol: stopping tolerance of the eigensolver (<= 0: working precision) */
size_t rl_rr_solve_ws_bytes(int64_t nmax);
int rl_rr_solve(const double* ga, const double* u, int64_t ld, int64_t nx, int64_t ny,
                int64_t leftX, int64_t rightX, int64_t leftXn, int64_t rightXn, double* cx,
                int64_t ldcx, double* cz, int64_t ldcz, double* lmdx, double* lmdz, double* est,
                int64_t nmax, double eig_tol, void* ws, size_t ws_bytes, int* info_d, void* stream);


/* ---- device-resident post-processing of the partial SVD ---------
 * PartialSVD._finalize_svd (interfaces/partial_svd.py:163-235) factorises nsv x nsv matrices
 * with scipy.linalg eigh / cholesky / svd / inv on the host.  Here: Gram matrix on the device,
 * a Gershgorin certificate (or the eigenvalues) of its diagonally scaled form for the
 * conditioning test (:171-182), ONE symmetric eigen-decomposition on the device for
 * U = chol(G), svd(U), inv(U) (:190-197: A v = (Av q Sigma^-1) Sigma q^T), coefficient blocks. */
/* fp64 small matrix -> typed block (optionally transposed), and back */
int rl_small_to_block(int dtype, const double* src, int64_t lds, int64_t rows, int64_t cols,
                      int trans, void* dst, int64_t ldd, void* stream);
int rl_block_to_small(int dtype, const void* src, int64_t lds, int64_t rows, int64_t cols,
                      double* dst, int64_t ldd, void* stream);
/* out3 = { max_i sum_{j != i} |S_ij|, min_i G_ii, max_i G_ii },  S = D^-1/2 G D^-1/2 */
int rl_psvd_gershgorin(const double* g, int64_t ld, int64_t n, double* out3, void* stream);
int rl_psvd_scale(const double* g, int64_t ld, int64_t n, double* s, int64_t lds, void* stream);
/* (w ascending, Q) -> sigma descending, q = Q reordered, cs = q diag(1 / sigma) */
int rl_psvd_coeffs(const double* qin, int64_t ldq, const double* w, int64_t n, double* q,
                   double* cs, int64_t ldo, double* sigma, void* stream);

/* pieces of the conditioning test of _finalize_svd (partial_svd.py:171-182) without eigenvalues:
 * with G = U^T U and S = D^-1/2 G D^-1/2,  lambda_min(S) >= 1 / sum_ij d_i (U^-1)_ij^2.
 * rl_small_set_identity: a = I;  rl_psvd_invbound: out[0] = sum_ij g_ii uinv_ij^2 */
int rl_small_set_identity(double* a, int64_t ld, int64_t n, void* stream);
/* dst[r][c] = src[r][c] * s[c] */
int rl_small_scale_cols(const double* src, int64_t lds, int64_t rows, int64_t cols, const double* s,
                        double* dst, int64_t ldd, void* stream);
int rl_psvd_invbound(const double* uinv, int64_t ldu, const double* g, int64_t ldg, int64_t n,
                     double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RALEIGH_B200_H */
