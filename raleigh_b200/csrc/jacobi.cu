// Symmetric eigensolver for the Rayleigh-Ritz step (solver.py:1459, 1470: scipy.linalg.eigh on
// the host in the reference), fp64, order n <= 320: ONE-SIDED Jacobi on a thread-block CLUSTER.
//
// Method.  B = G + sigma I with sigma from the Gershgorin discs so that B is positive definite.
// Plane rotations applied to the COLUMNS of B make them mutually orthogonal: B V = Q D with
// orthonormal Q and D = diag(lambda + sigma), hence G = Q diag(lambda) Q^T.  A rotation needs three
// dot products of the two columns and touches nothing else -- no row pass, no eigenvector
// accumulation -- and it preserves high relative accuracy of the columns (Demmel-Veselic).
//
// Mapping.  A round of the round-robin tournament handles n/2 disjoint column pairs; one WARP owns
// one pair, the columns live in shared memory, 2 KB each at n = 256.  The whole matrix (512 KB)
// does not fit one SM, so the pairs are spread over a cluster of up to 8 CTAs and the tournament is
// done by MOVING columns: after its rotation a warp stores its two columns into the slots of the
// neighbouring pairs (next round's partners), double-buffered; only the two columns at the ends of
// a CTA cross to the neighbouring CTA through distributed shared memory.  One cluster barrier per
// round, 2n-1 rounds per sweep; everything else is warp-local (shuffle reductions).
//
// Cost model (B200): per round and warp 7*(n/32) DFMA instructions + a 5-step shuffle tree +
// one sqrt/div chain + the cluster barrier (~380 cycles) ~ 0.5 us at n = 256; 4-6 sweeps on the
// nearly diagonal matrices of the Rayleigh-Ritz step => ~0.6 ms, against 16.7 ms for the
// cooperative-grid two-sided kernel in small.cu (measured r1e) and ~6 ms for LAPACK on the host.
#include <cooperative_groups.h>
#include "common.cuh"
#include "tma.cuh"

namespace cg = cooperative_groups;

namespace rl {

constexpr int JC_MAX_SWEEPS = 48;
constexpr int JC_MAX_CLUSTER = 8;

// sigma from the Gershgorin discs of sym(G); one CTA, one warp per row (coalesced row reads)
__global__ void __launch_bounds__(1024)
jacobi_shift_kernel(const double* __restrict__ G, int64_t ld, int n, double* __restrict__ par) {
    __shared__ double rlo[32], rhi[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double lo = 1.0e308, hi = 0.0;
    for (int i = warp; i < n; i += nw) {
        double rad = 0.0;
        for (int j = lane; j < n; j += 32)
            if (j != i) rad += fabs(0.5 * (G[(int64_t)i * ld + j] + G[(int64_t)j * ld + i]));
        rad = warp_sum(rad);
        const double d = G[(int64_t)i * ld + i];
        lo = fmin(lo, d - rad);
        hi = fmax(hi, fabs(d) + rad);
    }
    if (lane == 0) { rlo[warp] = lo; rhi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) { lo = fmin(lo, rlo[w]); hi = fmax(hi, rhi[w]); }
        double sigma = 0.0;
        if (lo <= 1e-3 * hi) sigma = -lo + 1e-2 * hi;     // not safely positive definite: shift
        if (!(hi > 0.0)) sigma = 1.0;                      // zero matrix: B = I
        par[0] = sigma;
        par[1] = hi;
    }
}

// Rotation of a column pair from its three dot products, |theta| <= pi/4.  Measured on B200
// (tools/ubench/fp64_lat.cu): DFMA 8 cycles dependent, rsqrt 62, division 119 -- and since every lane of the
// warp computes the same scalars, the chain also costs FP64 issue slots (2 cycles per warp instruction and
// scheduler).  So: no division and no tangent.  With r = 1/sqrt(delta^2 + 4 gamma^2):
//   cos 2theta = |delta| r,  c^2 = (1 + cos 2theta)/2,  c = c^2 rsqrt(c^2),  s = sign(delta) gamma r / c
// (c^2 + s^2 = h r^2 = 1 to rounding, like any computed rotation): two rsqrt and ~10 DFMA, 165 cycles.
// dn = change of |p|^2 = minus the change of |q|^2 under the rotation, for the callers that track norms.
__device__ __forceinline__ bool jacobi_rotation(double alpha, double beta, double gamma, double tol2, double& c,
                                                double& s, double& dn) {
    c = 1.0; s = 0.0; dn = 0.0;
    if (!(gamma * gamma > tol2 * alpha * beta)) return false;      // also false for zero (padding) columns
    const double delta = beta - alpha;
    const double h = fma(delta, delta, 4.0 * gamma * gamma);
    const double r = rsqrt(h);
    const double c2 = fma(0.5 * fabs(delta), r, 0.5);
    const double rc = rsqrt(c2);
    c = c2 * rc;
    s = (delta >= 0.0 ? gamma : -gamma) * r * rc;
    dn = s * fma(s, delta, -2.0 * c * gamma);
    return true;
}

// NJ = (padded column length) / 64: a lane holds NJ 16-byte chunks (2 doubles) of each column,
// chunk index = lane + 32 * j -- consecutive lanes read consecutive 16-byte words (no bank conflicts).
//
// Two-level tournament.  Every CTA of the cluster holds two BLOCKS of W columns (A and B).  A sweep is
//   (1) W-1 rounds among the columns of each block (circle method on indices, in place),
//   (2) 2C-1 outer rounds; in each, W rounds pair A_w with B_(w+r) -- A_w stays in the registers of
//       warp w, only B columns go through shared memory -- and then the blocks move to the
//       neighbouring CTAs (circle method on blocks) through distributed shared memory, double buffered.
// All (W-1) + (2C-1) W = n-1 rounds of a sweep cost one __syncthreads each; the cluster barrier
// (~400 cycles + DSMEM traffic) is paid 2C-1 times per sweep instead of n-1 times.
// (at most 16 warps per CTA up to n = 256, 20 at n = 320: see rl_syevj_cluster)
template <int NJ>
__global__ void __launch_bounds__(NJ == 5 ? 640 : 512)
jacobi_cluster_kernel(const double* __restrict__ G, int64_t ldg, int n, int W, int factor_mode, double tol,
                      const double* __restrict__ par, double* __restrict__ wtmp, double* __restrict__ qtmp,
                      int* __restrict__ info) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char jc_smem[];
    constexpr int LEN = 64 * NJ;                                 // padded column length (doubles)
    double* buf = reinterpret_cast<double*>(jc_smem);            // [2][2W][LEN]: A block = columns 0..W-1, B block = W..2W-1
    __shared__ int s_rot[32];
    __shared__ double s_bn[32];                                  // squared norms of the B columns
    __shared__ int s_flag[JC_MAX_SWEEPS][JC_MAX_CLUSTER];        // meaningful in CTA 0 only
    __shared__ int s_any;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rank = (int)cluster.block_rank();
    const int C = (int)cluster.num_blocks();
    const double sigma = factor_mode ? 0.0 : par[0];
    const size_t bufstride = (size_t)2 * W * LEN;

    for (int i = threadIdx.x; i < JC_MAX_SWEEPS * JC_MAX_CLUSTER; i += blockDim.x) (&s_flag[0][0])[i] = 0;
    // initial columns of B = sym(G) + sigma I: CTA `rank` holds global columns [rank 2W, (rank+1) 2W)
    for (int side = 0; side < 2; ++side) {
        const int c = rank * 2 * W + side * W + w;
        double2* col = reinterpret_cast<double2*>(buf + (size_t)(side * W + w) * LEN);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int r0 = 2 * (lane + 32 * j);
            double2 v = make_double2(0.0, 0.0);
            if (c < n && factor_mode) {             // column c of L = row c of the upper factor U
                if (r0 < n) v.x = G[(int64_t)c * ldg + r0];
                if (r0 + 1 < n) v.y = G[(int64_t)c * ldg + r0 + 1];
            } else if (c < n) {
                if (r0 < n) v.x = 0.5 * (G[(int64_t)r0 * ldg + c] + G[(int64_t)c * ldg + r0]) + (r0 == c ? sigma : 0.0);
                if (r0 + 1 < n) v.y = 0.5 * (G[(int64_t)(r0 + 1) * ldg + c] + G[(int64_t)c * ldg + r0 + 1]) + (r0 + 1 == c ? sigma : 0.0);
            }
            col[lane + 32 * j] = v;
        }
    }
    cluster.sync();

    const double tol2 = tol * tol;
    // block movement of the outer circle tournament (the same every round):
    // A: CTA 0 keeps it; CTA c sends it to the A slot of c+1; the last CTA moves it to its own B slot.
    // B: CTA c sends it to the B slot of c-1; CTA 0 sends it to the A slot of CTA 1.
    int a_rank = rank, a_side = 0, b_rank = rank, b_side = 1;
    if (C > 1) {
        if (rank == 0) { a_rank = 0; a_side = 0; b_rank = 1; b_side = 0; }
        else {
            if (rank == C - 1) { a_rank = rank; a_side = 1; } else { a_rank = rank + 1; a_side = 0; }
            b_rank = rank - 1; b_side = 1;
        }
    }
    // destination columns in both buffers (resolved once: mapa + address arithmetic stay out of the loop)
    double2* dstA[2];
    double2* dstB[2];
    for (int b = 0; b < 2; ++b) {
        double* la = buf + b * bufstride + (size_t)(a_side * W + w) * LEN;
        double* lb = buf + b * bufstride + (size_t)(b_side * W + w) * LEN;
        dstA[b] = reinterpret_cast<double2*>(C > 1 && a_rank != rank ? cluster.map_shared_rank(la, a_rank) : la);
        dstB[b] = reinterpret_cast<double2*>(C > 1 && b_rank != rank ? cluster.map_shared_rank(lb, b_rank) : lb);
    }
    const int P = (W + 1) & ~1;                       // players of the intra-block tournament (one dummy if W is odd)
    const int nouter = C > 1 ? 2 * C - 1 : 1;
    int cur = 0, sweep = 0, converged = 0;
    // cycle counts of CTA 0 by phase (intra-block rounds | norms | cross-block rounds | block exchange), read by
    // tools/time_rr.py from the workspace: four clock reads per outer round
    long long cyc_intra = 0, cyc_norms = 0, cyc_cross = 0, cyc_xchg = 0;
    for (; sweep < JC_MAX_SWEEPS; ++sweep) {
        int rot = 0;
        long long tk = clock64();
        // (1) pairs inside each block
        for (int t = 0; t < P - 1; ++t) {
            for (int task = w; task < P; task += W) {
                const int blk = task >= P / 2 ? 1 : 0;
                const int i = task - blk * (P / 2);
                int a, b;
                if (i == 0) { a = P - 1; b = t; } else { a = (t + i) % (P - 1); b = (t - i + P - 1) % (P - 1); }
                if (a > b) { const int x = a; a = b; b = x; }
                if (b >= W) continue;                                   // dummy player
                double2* cp = reinterpret_cast<double2*>(buf + cur * bufstride + (size_t)(blk * W + a) * LEN);
                double2* cq = reinterpret_cast<double2*>(buf + cur * bufstride + (size_t)(blk * W + b) * LEN);
                double2 p[NJ], q[NJ];
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    p[j] = cp[lane + 32 * j];
                    q[j] = cq[lane + 32 * j];
                    alpha = fma(p[j].x, p[j].x, alpha); alpha = fma(p[j].y, p[j].y, alpha);
                    beta = fma(q[j].x, q[j].x, beta); beta = fma(q[j].y, q[j].y, beta);
                    gamma = fma(p[j].x, q[j].x, gamma); gamma = fma(p[j].y, q[j].y, gamma);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
                    beta += __shfl_xor_sync(0xffffffffu, beta, o);
                    gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
                }
                double c, s, t;
                if (jacobi_rotation(alpha, beta, gamma, tol2, c, s, t)) {
                    rot = 1;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        double2 x, y;
                        x.x = c * p[j].x - s * q[j].x; x.y = c * p[j].y - s * q[j].y;
                        y.x = s * p[j].x + c * q[j].x; y.y = s * p[j].y + c * q[j].y;
                        cp[lane + 32 * j] = x;
                        cq[lane + 32 * j] = y;
                    }
                }
            }
            __syncthreads();
        }
        { const long long t1 = clock64(); cyc_intra += t1 - tk; tk = t1; }
        // (2) pairs across blocks: W local rounds per outer round, then the blocks move on.  Squared norms are
        // computed once per outer round and then updated by the rotations (at most W updates apart): a round needs
        // ONE dot product and one shuffle tree -- with 16 warps the 64-bit shuffles of three trees alone cost
        // 480 cycles per round (32 lanes per clock and SM), one tree 176.
        for (int o = 0; o < nouter; ++o) {
            double2* ca = reinterpret_cast<double2*>(buf + cur * bufstride + (size_t)w * LEN);
            double2 p[NJ];
            double alpha = 0.0;
            {
                const double2* cbw = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(W + w) * LEN);
                double b0 = 0.0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    p[j] = ca[lane + 32 * j];
                    const double2 qv = cbw[lane + 32 * j];
                    alpha = fma(p[j].x, p[j].x, alpha); alpha = fma(p[j].y, p[j].y, alpha);
                    b0 = fma(qv.x, qv.x, b0); b0 = fma(qv.y, qv.y, b0);
                }
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) {
                    alpha += __shfl_xor_sync(0xffffffffu, alpha, o2);
                    b0 += __shfl_xor_sync(0xffffffffu, b0, o2);
                }
                if (lane == 0) s_bn[w] = b0;
            }
            __syncthreads();
            { const long long t1 = clock64(); cyc_norms += t1 - tk; tk = t1; }
            for (int r = 0; r < W; ++r) {
                int jb = w + r; if (jb >= W) jb -= W;
                double2* cb = reinterpret_cast<double2*>(buf + cur * bufstride + (size_t)(W + jb) * LEN);
                const double beta = s_bn[jb];
                double2 q[NJ];
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    q[j] = cb[lane + 32 * j];
                    g0 = fma(p[j].x, q[j].x, g0); g1 = fma(p[j].y, q[j].y, g1);
                }
                double gamma = g0 + g1;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) gamma += __shfl_xor_sync(0xffffffffu, gamma, o2);
                double c, s, dn;
                if (jacobi_rotation(alpha, beta, gamma, tol2, c, s, dn)) {
                    rot = 1;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        double2 x, y;
                        x.x = c * p[j].x - s * q[j].x; x.y = c * p[j].y - s * q[j].y;
                        y.x = s * p[j].x + c * q[j].x; y.y = s * p[j].y + c * q[j].y;
                        p[j] = x;
                        cb[lane + 32 * j] = y;
                    }
                    alpha += dn;
                    if (lane == 0) s_bn[jb] = beta - dn;
                }
                __syncthreads();
            }
            { const long long t1 = clock64(); cyc_cross += t1 - tk; tk = t1; }
            if (C > 1) {
                // A_w (registers) and B_w (shared memory) go to their next CTA, other buffer
                const double2* cbw = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(W + w) * LEN);
                double2* da = dstA[cur ^ 1];
                double2* db = dstB[cur ^ 1];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    da[lane + 32 * j] = p[j];
                    db[lane + 32 * j] = cbw[lane + 32 * j];
                }
                cluster.sync();
                cur ^= 1;
            } else {
#pragma unroll
                for (int j = 0; j < NJ; ++j) ca[lane + 32 * j] = p[j];
                __syncthreads();
            }
            { const long long t1 = clock64(); cyc_xchg += t1 - tk; tk = t1; }
        }
        // did anybody rotate in this sweep?
        if (lane == 0) s_rot[w] = rot;
        __syncthreads();
        int any = 0;
        if (C > 1) {
            if (threadIdx.x == 0) {
                int a = 0;
                for (int i = 0; i < W; ++i) a |= s_rot[i];
                *cluster.map_shared_rank(&s_flag[sweep][rank], 0) = a;
            }
            cluster.sync();
            const int* f = cluster.map_shared_rank(&s_flag[sweep][0], 0);
            for (int r = 0; r < C; ++r) any |= f[r];
        } else {
            if (threadIdx.x == 0) {
                int a = 0;
                for (int i = 0; i < W; ++i) a |= s_rot[i];
                s_any = a;
            }
            __syncthreads();
            any = s_any;
            __syncthreads();
        }
        if (!any) { converged = 1; ++sweep; break; }
    }
    // nobody may leave (and release its shared memory) while others still read the flags of CTA 0
    cluster.sync();
    // eigenvalue = column norm - sigma, eigenvector = column / norm; slots in arbitrary order, sorted later
    for (int side = 0; side < 2; ++side) {
        const int slot = rank * 2 * W + side * W + w;
        const double2* col = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(side * W + w) * LEN);
        double2 v[NJ];
        double nrm = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) { v[j] = col[lane + 32 * j]; nrm = fma(v[j].x, v[j].x, nrm); nrm = fma(v[j].y, v[j].y, nrm); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        nrm = sqrt(nrm);
        const bool real_col = nrm > 0.0;
        if (lane == 0) wtmp[slot] = real_col ? (factor_mode ? nrm * nrm : nrm - sigma) : (double)INFINITY;   // resolved by the sort kernel
        const double inv = real_col ? 1.0 / nrm : 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int r0 = 2 * (lane + 32 * j);
            if (r0 < n) qtmp[(size_t)slot * LEN + r0] = v[j].x * inv;
            if (r0 + 1 < n) qtmp[(size_t)slot * LEN + r0 + 1] = v[j].y * inv;
        }
    }
    if (rank == 0 && threadIdx.x == 0) {
        info[0] = sweep; info[1] = converged;
        long long* prof = reinterpret_cast<long long*>(const_cast<double*>(par) + 4);
        prof[0] = cyc_intra; prof[1] = cyc_norms; prof[2] = cyc_cross; prof[3] = cyc_xchg;
    }
}

// ascending order: w[rank] = value, Q[r][rank] = qtmp[slot][r]; every CTA ranks all slots, then scatters its share.
// Columns of zero norm arrive marked +inf: the first (slots - n) of them in slot order are the padding columns,
// any further ones are genuine null directions of a singular matrix and get `zero_value` (their eigenvector
// columns stay zero: the caller completes the basis).
__global__ void __launch_bounds__(1024)
jacobi_sort_kernel(const double* __restrict__ wtmp, const double* __restrict__ qtmp, int slots, int len, int n,
                   double zero_value, double* __restrict__ w, double* __restrict__ Q, int64_t ldq) {
    extern __shared__ __align__(8) unsigned char js_smem[];
    double* vals = reinterpret_cast<double*>(js_smem);               // [slots]
    int* s_rank = reinterpret_cast<int*>(vals + slots);              // [slots]
    const int npad = slots - n;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        double v = wtmp[i];
        if (isinf(v)) {
            int before = 0;
            for (int j = 0; j < i; ++j) before += isinf(wtmp[j]) ? 1 : 0;
            v = before < npad ? 1.0e308 : zero_value;
        }
        vals[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        const double v = vals[i];
        int r = 0;
        for (int j = 0; j < slots; ++j) { const double u = vals[j]; r += (u < v) || (u == v && j < i); }
        s_rank[i] = r;
        if (r < n && blockIdx.x == 0) w[r] = v;
    }
    __syncthreads();
    // 32 x 32 tiles through shared memory so that both the reads (along r) and the writes (along c) coalesce
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_r = (n + 31) / 32, tiles_s = (slots + 31) / 32;
    for (int tIdx = blockIdx.x; tIdx < tiles_r * tiles_s; tIdx += gridDim.x) {
        const int s0 = (tIdx / tiles_r) * 32, r0 = (tIdx % tiles_r) * 32;
        __syncthreads();
        if (s0 + ty < slots && r0 + tx < n) tile[ty][tx] = qtmp[(size_t)(s0 + ty) * len + r0 + tx];
        __syncthreads();
        if (s0 + tx < slots && r0 + ty < n) {
            const int c = s_rank[s0 + tx];
            if (c < n) Q[(int64_t)(r0 + ty) * ldq + c] = tile[tx][ty];
        }
    }
}

// ---- orders beyond the shared memory of a cluster (320 < n <= 1024): the whole GPU, cooperative launch ------
// Same one-sided method; the columns stay in global memory (8 MB at n = 1000: L2 resident), one WARP per pair
// with both columns in registers (NQ 16-byte chunks per lane and column), pairs chosen by the circle method on
// INDICES (nothing moves), one grid barrier per round.  Per round every column is read and written once
// (16 MB of L2 traffic at n = 1000) -- the barrier and the L2 round trip, ~3 us, are the cost.
// bt: n x ldb, row c = column c of B (contiguous), rotated in place.
template <int NQ>
__global__ void __launch_bounds__(128)
jacobi_grid_kernel(double* __restrict__ bt, int64_t ldb, int n, double tol, int* __restrict__ flags,
                   int* __restrict__ info) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nw = (int)((gridDim.x * blockDim.x) >> 5);
    const int P = (n + 1) & ~1, half = P / 2;
    const double tol2 = tol * tol;
    int sweep = 0, converged = 0;
    for (; sweep < JC_MAX_SWEEPS; ++sweep) {
        int rot = 0;
        for (int t = 0; t < P - 1; ++t) {
            for (int i = gw; i < half; i += nw) {
                int a, b;
                if (i == 0) { a = P - 1; b = t; } else { a = (t + i) % (P - 1); b = (t - i + P - 1) % (P - 1); }
                if (a > b) { const int x = a; a = b; b = x; }
                if (b >= n) continue;
                double2* cp = reinterpret_cast<double2*>(bt + (int64_t)a * ldb);
                double2* cq = reinterpret_cast<double2*>(bt + (int64_t)b * ldb);
                double2 p[NQ], q[NQ];
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int ch = lane + 32 * j;
                    p[j] = make_double2(0.0, 0.0); q[j] = p[j];
                    if (2 * ch < n) { p[j] = __ldcg(cp + ch); q[j] = __ldcg(cq + ch); }     // L2: written by other SMs last round
                    alpha = fma(p[j].x, p[j].x, alpha); alpha = fma(p[j].y, p[j].y, alpha);
                    beta = fma(q[j].x, q[j].x, beta); beta = fma(q[j].y, q[j].y, beta);
                    gamma = fma(p[j].x, q[j].x, gamma); gamma = fma(p[j].y, q[j].y, gamma);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
                    beta += __shfl_xor_sync(0xffffffffu, beta, o);
                    gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
                }
                double c, s, t;
                if (jacobi_rotation(alpha, beta, gamma, tol2, c, s, t)) {
                    rot = 1;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const int ch = lane + 32 * j;
                        if (2 * ch < n) {
                            double2 x, y;
                            x.x = c * p[j].x - s * q[j].x; x.y = c * p[j].y - s * q[j].y;
                            y.x = s * p[j].x + c * q[j].x; y.y = s * p[j].y + c * q[j].y;
                            __stcg(cp + ch, x);
                            __stcg(cq + ch, y);
                        }
                    }
                }
            }
            if (t == P - 2 && rot && lane == 0) atomicOr(flags + sweep, 1);
            grid.sync();
        }
        const int any = *reinterpret_cast<volatile int*>(flags + sweep);
        if (!any) { converged = 1; ++sweep; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = sweep; info[1] = converged; }
}

// bt (n x ldb, ldb even, padding zeroed) <- columns of sym(G) + sigma I (mode 0) or rows of the factor U (mode 1)
__global__ void jacobi_grid_init_kernel(const double* __restrict__ G, int64_t ldg, int n, int factor_mode,
                                        const double* __restrict__ par, double* __restrict__ bt, int64_t ldb,
                                        int* __restrict__ flags) {
    const double sigma = factor_mode ? 0.0 : par[0];
    if (blockIdx.x == 0 && blockIdx.y == 0)
        for (int i = threadIdx.x; i < JC_MAX_SWEEPS; i += blockDim.x) flags[i] = 0;
    for (int c = blockIdx.y; c < n; c += gridDim.y)
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < ldb; r += gridDim.x * blockDim.x) {
            double v = 0.0;
            if (r < n) v = factor_mode ? G[(int64_t)c * ldg + r]
                                       : 0.5 * (G[(int64_t)r * ldg + c] + G[(int64_t)c * ldg + r]) + (r == c ? sigma : 0.0);
            bt[(int64_t)c * ldb + r] = v;
        }
}

// column norms -> eigenvalues, normalised columns in place; one warp per column
__global__ void jacobi_grid_finish_kernel(double* __restrict__ bt, int64_t ldb, int n, int factor_mode,
                                          const double* __restrict__ par, double* __restrict__ wtmp) {
    const int lane = threadIdx.x & 31;
    const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (c >= n) return;
    const double sigma = factor_mode ? 0.0 : par[0];
    double* col = bt + (int64_t)c * ldb;
    double nrm = 0.0;
    for (int r = lane; r < n; r += 32) nrm = fma(col[r], col[r], nrm);
    nrm = sqrt(warp_sum(nrm));
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int r = lane; r < n; r += 32) col[r] *= inv;
    if (lane == 0) wtmp[c] = nrm > 0.0 ? (factor_mode ? nrm * nrm : nrm - sigma) : (factor_mode ? 0.0 : -sigma);
}

// ---- 320 < n <= 1024, two-level tournament over the whole GPU (default) ------------------------------------------
// The scheme of the cluster kernel with global memory (L2 resident, 16 MB at n = 1000) in place of distributed
// shared memory: CTA c holds two blocks of RW = 8 columns; an outer round = (load both blocks into shared memory)
// + RW inner rounds pairing A_w (registers of warp w) with B_(w+r) + (send A to CTA c+1, B to CTA c-1, into the
// other half of a double buffer).  n-1 rotation rounds per sweep as before, but only 2C-1 ~ n/8 of them end
// with an exchange, and the exchange is POINT-TO-POINT: a CTA waits for the two blocks it receives (one
// release/acquire flag per block, written by the block's unique sender) instead of for the whole grid.  Every
// CTA sends to exactly the CTAs it receives from, so a sender that has started round g has seen its
// destinations' round g-1 messages, i.e. they are past their loads of the buffer half it overwrites.
// One counting barrier per sweep carries the "anybody rotated" flag.  The launch is cooperative for the
// co-residency guarantee only; every spin is bounded and watches a global abort word (info[1] = -1).
constexpr int RW = 8;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 only; false = gave up (somebody aborted or the bound was hit)
__device__ __forceinline__ bool ring_wait(const unsigned* flag, unsigned want, unsigned* abort_word) {
    for (unsigned spin = 0; ld_acquire(flag) < want; ++spin) {
        if ((spin & 1023u) == 1023u) {
            if (ld_acquire(abort_word) != 0u) return false;
            if (spin > (1u << 27)) { atomicExch(abort_word, 1u); return false; }
        }
    }
    return true;
}
// both channels at once: the two polls travel together, ONE fence orders everything after them
__device__ __forceinline__ bool ring_wait2(const unsigned* fa, const unsigned* fb, unsigned want, unsigned* abort_word) {
    for (unsigned spin = 0;; ++spin) {
        const unsigned a = ld_relaxed(fa), b = ld_relaxed(fb);
        if (a >= want && b >= want) break;
        if ((spin & 1023u) == 1023u) {
            if (ld_relaxed(abort_word) != 0u) return false;
            if (spin > (1u << 27)) { atomicExch(abort_word, 1u); return false; }
        }
    }
    __threadfence();
    return true;
}

// bulk (TMA, non-tensor) copies: one thread moves a whole block of columns between shared memory and L2
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// sync words: flagA[C] | flagB[C] | barrier | abort | rotated[JC_MAX_SWEEPS]
template <int NQ>
__global__ void __launch_bounds__(RW * 32, 1)
jacobi_ring_kernel(double* __restrict__ gbuf, int factor_mode, double tol, const double* __restrict__ par,
                   unsigned* __restrict__ sync_words, double* __restrict__ wtmp, int* __restrict__ info, int dbg) {
    extern __shared__ __align__(16) unsigned char jr_smem[];
    constexpr int LEN = 64 * NQ;
    constexpr int CH = LEN / 2;                                  // 16-byte chunks per column
    double2* sm = reinterpret_cast<double2*>(jr_smem);           // [2 RW][CH]: A block, then B block
    __shared__ int s_rot[RW];
    __shared__ int s_ok, s_any;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double s_bn[RW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rank = blockIdx.x, C = gridDim.x;
    const size_t half_stride = (size_t)2 * RW * C * CH;          // double2 per half of the double buffer
    double2* g2 = reinterpret_cast<double2*>(gbuf);
    unsigned* flagA = sync_words;
    unsigned* flagB = sync_words + C;
    unsigned* bar = sync_words + 2 * C;
    unsigned* abort_word = bar + 1;
    unsigned* rotated = bar + 2;
    const double tol2 = tol * tol;
    const double sigma = factor_mode ? 0.0 : par[0];
    constexpr uint32_t BLOCK_BYTES = RW * LEN * sizeof(double);  // one block of columns: 32 or 64 KB

    int a_rank, a_side, b_rank, b_side;
    if (rank == 0) { a_rank = 0; a_side = 0; b_rank = 1; b_side = 0; }
    else {
        if (rank == C - 1) { a_rank = rank; a_side = 1; } else { a_rank = rank + 1; a_side = 0; }
        b_rank = rank - 1; b_side = 1;
    }
    if (dbg & 2) { a_rank = rank; a_side = 0; b_rank = rank; b_side = 1; }      // timing experiment: nothing moves
    const size_t own = (size_t)rank * 2 * RW * CH;
    const size_t dst_a = ((size_t)a_rank * 2 * RW + a_side * RW) * CH;
    const size_t dst_b = ((size_t)b_rank * 2 * RW + b_side * RW) * CH;
    unsigned* flag_a = (a_side ? flagB : flagA) + a_rank;
    unsigned* flag_b = (b_side ? flagB : flagA) + b_rank;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nouter = 2 * C - 1;
    unsigned g = 0;                                              // outer rounds done = messages sent per channel
    int sweep = 0, converged = 0, aborted = 0;
    for (; sweep < JC_MAX_SWEEPS && !aborted; ++sweep) {
        int rot = 0;
        for (int o = 0; o < nouter; ++o) {
            // thread 0: wait for this round's two blocks, then pull them (contiguous: A block, B block) into shared memory
            if (threadIdx.x == 0) {
                int ok = 1;
                if (g > 0) ok = ring_wait2(flagA + rank, flagB + rank, g, abort_word);
                s_ok = ok;
                if (ok) {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    const double2* src = g2 + (g & 1u) * half_stride + own;
                    mbar_expect_tx(&s_bar, 2 * BLOCK_BYTES);
#pragma unroll
                    for (int part = 0; part < 4; ++part)
                        bulk_g2s(reinterpret_cast<unsigned char*>(sm) + part * (BLOCK_BYTES / 2),
                                 reinterpret_cast<const unsigned char*>(src) + part * (BLOCK_BYTES / 2), BLOCK_BYTES / 2, &s_bar);
                }
            }
            __syncthreads();
            if (!s_ok) { aborted = 1; break; }
            mbar_wait(&s_bar, g & 1u);
            if (o == 0 && !(dbg & 1)) {
                // pairs inside each block: circle method on RW players, 4 pairs per block = one per warp
                for (int t = 0; t < RW - 1; ++t) {
                    const int blk = w >> 2, i = w & 3;
                    int a, b;
                    if (i == 0) { a = RW - 1; b = t; } else { a = (t + i) % (RW - 1); b = (t - i + RW - 1) % (RW - 1); }
                    double2* cp = sm + (size_t)(blk * RW + a) * CH;
                    double2* cq = sm + (size_t)(blk * RW + b) * CH;
                    double2 p[NQ], q[NQ];
                    double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        p[j] = cp[lane + 32 * j];
                        q[j] = cq[lane + 32 * j];
                        alpha = fma(p[j].x, p[j].x, alpha); alpha = fma(p[j].y, p[j].y, alpha);
                        beta = fma(q[j].x, q[j].x, beta); beta = fma(q[j].y, q[j].y, beta);
                        gamma = fma(p[j].x, q[j].x, gamma); gamma = fma(p[j].y, q[j].y, gamma);
                    }
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) {
                        alpha += __shfl_xor_sync(0xffffffffu, alpha, o2);
                        beta += __shfl_xor_sync(0xffffffffu, beta, o2);
                        gamma += __shfl_xor_sync(0xffffffffu, gamma, o2);
                    }
                    double c, s, tt;
                    if (jacobi_rotation(alpha, beta, gamma, tol2, c, s, tt)) {
                        rot = 1;
#pragma unroll
                        for (int j = 0; j < NQ; ++j) {
                            double2 x, y;
                            x.x = c * p[j].x - s * q[j].x; x.y = c * p[j].y - s * q[j].y;
                            y.x = s * p[j].x + c * q[j].x; y.y = s * p[j].y + c * q[j].y;
                            cp[lane + 32 * j] = x;
                            cq[lane + 32 * j] = y;
                        }
                    }
                    __syncthreads();
                }
            }
            // pairs across the blocks: A_w in registers, B columns through shared memory.  The squared norms of all
            // columns are computed once per outer round and then UPDATED by the rotations (|p'|^2 = |p|^2 - t <p, q>,
            // |q'|^2 = |q|^2 + t <p, q>, exact for the exact angle; at most RW updates apart, so they cannot drift):
            // a round needs one dot product and one shuffle tree, not three.
            double2 p[NQ];
            double alpha;
            {
                const double2* ca = sm + (size_t)w * CH;
                const double2* cbw = sm + (size_t)(RW + w) * CH;
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    p[j] = ca[lane + 32 * j];
                    const double2 qv = cbw[lane + 32 * j];
                    a0 = fma(p[j].x, p[j].x, a0); a1 = fma(p[j].y, p[j].y, a1);
                    b0 = fma(qv.x, qv.x, b0); b1 = fma(qv.y, qv.y, b1);
                }
                a0 += a1; b0 += b1;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o2);
                    b0 += __shfl_xor_sync(0xffffffffu, b0, o2);
                }
                alpha = a0;
                if (lane == 0) s_bn[w] = b0;
            }
            __syncthreads();
            for (int r = 0; r < ((dbg & 1) ? 0 : RW); ++r) {
                const int jb = (w + r) & (RW - 1);
                double2* cb = sm + (size_t)(RW + jb) * CH;
                const double beta = s_bn[jb];
                double2 q[NQ];
                double g0 = 0.0, g1 = 0.0, g2a = 0.0, g3 = 0.0;
#pragma unroll
                for (int j = 0; j < NQ; j += 2) {
                    q[j] = cb[lane + 32 * j];
                    q[j + 1] = cb[lane + 32 * (j + 1)];
                    g0 = fma(p[j].x, q[j].x, g0); g1 = fma(p[j].y, q[j].y, g1);
                    g2a = fma(p[j + 1].x, q[j + 1].x, g2a); g3 = fma(p[j + 1].y, q[j + 1].y, g3);
                }
                double gamma = (g0 + g1) + (g2a + g3);
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) gamma += __shfl_xor_sync(0xffffffffu, gamma, o2);
                double c, s, tt;
                if (jacobi_rotation(alpha, beta, gamma, tol2, c, s, tt)) {
                    rot = 1;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        double2 x, y;
                        x.x = c * p[j].x - s * q[j].x; x.y = c * p[j].y - s * q[j].y;
                        y.x = s * p[j].x + c * q[j].x; y.y = s * p[j].y + c * q[j].y;
                        p[j] = x;
                        cb[lane + 32 * j] = y;
                    }
                    alpha += tt;
                    if (lane == 0) s_bn[jb] = beta - tt;
                }
                __syncthreads();
            }
            // the blocks move on: A_w back to shared memory, then two bulk stores into the other half of the double
            // buffer and one flag per block
            {
                double2* ca = sm + (size_t)w * CH;
#pragma unroll
                for (int j = 0; j < NQ; ++j) ca[lane + 32 * j] = p[j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            ++g;
            if (threadIdx.x == 0) {
                double2* da = g2 + (g & 1u) * half_stride + dst_a;
                double2* db = g2 + (g & 1u) * half_stride + dst_b;
                bulk_s2g(da, sm, BLOCK_BYTES);
                bulk_s2g(db, sm + (size_t)RW * CH, BLOCK_BYTES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();                   // one fence, then both flags (fence + relaxed store = release)
                st_relaxed(flag_a, g);
                st_relaxed(flag_b, g);
            }
        }
        if (aborted) break;
        // did anybody rotate in this sweep?  one counting barrier per sweep
        if (lane == 0) s_rot[w] = rot;
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = 0;
            for (int i = 0; i < RW; ++i) a |= s_rot[i];
            if (a) atomicOr(rotated + sweep, 1u);
            __threadfence();
            atomicAdd(bar, 1u);
            const bool ok = ring_wait(bar, (unsigned)C * (unsigned)(sweep + 1), abort_word);
            s_ok = ok;
            s_any = ok ? (int)ld_acquire(rotated + sweep) : 1;
        }
        __syncthreads();
        if (!s_ok) { aborted = 1; break; }
        if (dbg ? sweep == 5 : !s_any) { converged = 1; ++sweep; break; }
    }
    if (aborted) {
        if (rank == 0 && threadIdx.x == 0) { info[0] = sweep; info[1] = -1; }
        return;
    }
    // column norms -> eigenvalues, normalised columns into half 0 (slots in arbitrary order, sorted later).
    // The last messages of the final round are still in flight to their receivers: wait for them like a round would.
    if (threadIdx.x == 0) s_ok = ring_wait2(flagA + rank, flagB + rank, g, abort_word);
    __syncthreads();
    if (!s_ok) {
        if (rank == 0 && threadIdx.x == 0) { info[0] = sweep; info[1] = -1; }
        return;
    }
    for (int side = 0; side < 2; ++side) {
        const size_t slot = (size_t)rank * 2 * RW + side * RW + w;
        const double2* col = g2 + (g & 1u) * half_stride + slot * CH;
        double2 v[NQ];
        double nrm = 0.0;
#pragma unroll
        for (int j = 0; j < NQ; ++j) { v[j] = __ldcg(col + lane + 32 * j); nrm = fma(v[j].x, v[j].x, nrm); nrm = fma(v[j].y, v[j].y, nrm); }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o2);
        nrm = sqrt(nrm);
        const bool real_col = nrm > 0.0;
        if (lane == 0) wtmp[slot] = real_col ? (factor_mode ? nrm * nrm : nrm - sigma) : (double)INFINITY;
        const double inv = real_col ? 1.0 / nrm : 0.0;
        double2* out = g2 + slot * CH;                            // half 0; same place when (g & 1) == 0
#pragma unroll
        for (int j = 0; j < NQ; ++j) out[lane + 32 * j] = make_double2(v[j].x * inv, v[j].y * inv);
    }
    if (rank == 0 && threadIdx.x == 0) { info[0] = sweep; info[1] = converged; }
}

// gbuf half 0 (slots x len, padding slots / rows zeroed) <- columns of sym(G) + sigma I or rows of the factor U
__global__ void jacobi_ring_init_kernel(const double* __restrict__ G, int64_t ldg, int n, int factor_mode,
                                        const double* __restrict__ par, double* __restrict__ gbuf, int len, int slots,
                                        unsigned* __restrict__ sync_words, int nsync) {
    const double sigma = factor_mode ? 0.0 : par[0];
    if (blockIdx.x == 0 && blockIdx.y == 0)
        for (int i = threadIdx.x; i < nsync; i += blockDim.x) sync_words[i] = 0u;
    for (int c = blockIdx.y; c < slots; c += gridDim.y)
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < len; r += gridDim.x * blockDim.x) {
            double v = 0.0;
            if (r < n && c < n) v = factor_mode ? G[(int64_t)c * ldg + r]
                                                : 0.5 * (G[(int64_t)r * ldg + c] + G[(int64_t)c * ldg + r]) + (r == c ? sigma : 0.0);
            gbuf[(size_t)c * len + r] = v;
        }
}

template <int NJ>
static int launch_cluster(const double* G, int64_t ldg, int n, int W, int csize, int factor_mode, double tol, const double* par,
                          double* wtmp, double* qtmp, int* info, cudaStream_t st) {
    const size_t smem = (size_t)2 * 2 * W * 64 * NJ * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        RL_CUDA(cudaFuncSetAttribute(jacobi_cluster_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csize);
    cfg.blockDim = dim3((unsigned)(W * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++g_launches;
    return (int)cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<NJ>, G, ldg, n, W, factor_mode, tol, par, wtmp, qtmp, info);
}

template <int NQ>
static int launch_ring(double* gbuf, int C, int factor_mode, double tol, const double* par, unsigned* sync_words,
                       double* wtmp, int* info, cudaStream_t st) {
    const size_t smem = (size_t)2 * RW * 64 * NQ * sizeof(double);
    static bool configured = false;
    if (!configured) {
        RL_CUDA(cudaFuncSetAttribute(jacobi_ring_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int dbg = g_knob[KNOB_EIG_RING_DEBUG];
    void* args[] = {&gbuf, &factor_mode, &tol, &par, &sync_words, &wtmp, &info, &dbg};
    ++g_launches;
    return (int)cudaLaunchCooperativeKernel((void*)jacobi_ring_kernel<NQ>, dim3((unsigned)C), dim3(RW * 32), args, smem, st);
}

static int syevj_ring(const double* g, int64_t ldg, int64_t n, int factor_mode, double tol, double* w, double* q,
                      int64_t ldq, void* ws, int* info, cudaStream_t st) {
    const int64_t len = n <= 512 ? 512 : 1024;          // the two instantiations of the kernel
    const int C = (int)((n + 2 * RW - 1) / (2 * RW));
    const int slots = 2 * RW * C;
    if (C < 2 || C > sm_count()) return RL_E_ARG;
    double* par = (double*)ws;
    double* wtmp = par + 8;
    double* gbuf = wtmp + (2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER);
    unsigned* sync_words = (unsigned*)(gbuf + (size_t)2 * slots * len);
    const int nsync = 2 * C + 2 + JC_MAX_SWEEPS;
    int rc = 0;
    if (!factor_mode) {
        jacobi_shift_kernel<<<1, 1024, 0, st>>>(g, ldg, (int)n, par);
        rc = check_launch();
        if (rc) return rc;
    }
    jacobi_ring_init_kernel<<<dim3((unsigned)((len + 127) / 128), (unsigned)slots), 128, 0, st>>>(
        g, ldg, (int)n, factor_mode, par, gbuf, (int)len, slots, sync_words, nsync);
    rc = check_launch();
    if (rc) return rc;
    if (len <= 512) rc = launch_ring<8>(gbuf, C, factor_mode, tol, par, sync_words, wtmp, info, st);
    else rc = launch_ring<16>(gbuf, C, factor_mode, tol, par, sync_words, wtmp, info, st);
    if (rc) return rc;
    jacobi_sort_kernel<<<64, 1024, (size_t)slots * 12, st>>>(wtmp, gbuf, slots, (int)len, (int)n, 0.0, w, q, ldq);
    return check_launch();
}

}  // namespace rl

using namespace rl;

extern "C" {

int rl_syevj_cluster_max_n(void) { return 320; }
int rl_syevj_grid_max_n(void) { return 1024; }

/* workspace: par (8 doubles) | wtmp | qtmp (slots * len) | info (4 ints); grid path: bt (n * ldb) + flags */
size_t rl_syevj_cluster_ws_bytes(int64_t n) {
    if (n <= 0) return 0;
    const int64_t len = (n + 63) / 64 * 64;
    const int64_t slots = 2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER;
    if (n > rl_syevj_cluster_max_n())                  // ring kernel: double-buffered column store + sync words
        return (size_t)(8 + slots + 2 * (n + 2 * RW) * (n <= 512 ? 512 : 1024)) * sizeof(double) + 4096;
    return (size_t)(8 + slots + slots * len) * sizeof(double) + 1024;
}

static int syevj_grid(const double* g, int64_t ldg, int64_t n, int factor_mode, double tol, double* w, double* q,
                      int64_t ldq, void* ws, int* info, cudaStream_t st) {
    const int64_t len = (n + 63) / 64 * 64;
    double* par = (double*)ws;
    double* wtmp = par + 8;
    double* bt = wtmp + (2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER);
    int* flags = (int*)(bt + (size_t)n * len);
    if (!factor_mode) {
        jacobi_shift_kernel<<<1, 1024, 0, st>>>(g, ldg, (int)n, par);
        int rc = check_launch();
        if (rc) return rc;
    }
    jacobi_grid_init_kernel<<<dim3((unsigned)((len + 127) / 128), (unsigned)(n > 65535 ? 65535 : n)), 128, 0, st>>>(
        g, ldg, (int)n, factor_mode, par, bt, len, flags);
    int rc = check_launch();
    if (rc) return rc;
    const int NQ = (int)(len / 64);
    const int half = (int)((n + 1) / 2);
    int want = (half + 3) / 4;                       // one warp per pair, 4 warps per CTA
    void* kern = nullptr;
    if (NQ <= 8) kern = (void*)jacobi_grid_kernel<8>; else kern = (void*)jacobi_grid_kernel<16>;
    int per_sm = 0;
    RL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
    const int cap = per_sm * sm_count();
    int grid = want < cap ? want : cap;
    if (grid < 1) grid = 1;
    int ni = (int)n;
    int64_t ldb = len;
    void* args[] = {&bt, &ldb, &ni, &tol, &flags, &info};
    rc = (int)cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(128), args, 0, st);
    ++g_launches;
    if (rc) return rc;
    jacobi_grid_finish_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(bt, len, (int)n, factor_mode, par, wtmp);
    rc = check_launch();
    if (rc) return rc;
    jacobi_sort_kernel<<<(unsigned)(n >= 512 ? 64 : 8), 1024, (size_t)n * 12, st>>>(wtmp, bt, (int)n, (int)len, (int)n, 0.0, w, q, ldq);
    return check_launch();
}

/* Eigen-decomposition of the symmetric n x n fp64 matrix g (row-major, ldg; both triangles are
 * read and averaged): w[0..n) ascending, q[i*ldq + j] = component i of eigenvector j.
 * factor_mode != 0: g holds instead the UPPER Cholesky factor U of the matrix to decompose
 * (U^T U = q diag(w) q^T): the rotations act on the rows of U, no shift is needed and every
 * eigenvalue keeps its relative accuracy (Jacobi on the Cholesky factor, Veselic-Hari).
 * tol: a column pair is rotated while |<p, q>| > tol |p| |q| (<= 0: sqrt(n) eps, working precision; the
 * Rayleigh-Ritz step of fp32 problems stops at 1e-9).
 * g is not modified.  info_d (device, 2 ints): sweeps, converged.  n <= rl_syevj_grid_max_n();
 * orders up to rl_syevj_cluster_max_n() run on one thread-block cluster, larger ones on the whole GPU. */
int rl_syevj_cluster(const double* g, int64_t ldg, int64_t n, int factor_mode, double tol, double* w, double* q,
                     int64_t ldq, void* ws, size_t ws_bytes, int* info_d, void* stream) {
    if (n < 0 || n > rl_syevj_grid_max_n()) return RL_E_ARG;
    if (n == 0) return 0;
    if (!(tol > 0.0)) tol = sqrt((double)n) * 2.220446049250313e-16;      // default: working precision
    if (ws_bytes < rl_syevj_cluster_ws_bytes(n)) return RL_E_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int npairs = (int)((n + 1) / 2);
    const int NJ = (int)((n + 63) / 64);
    const int len = 64 * NJ;
    double* par = (double*)ws;
    double* wtmp = par + 8;
    double* qtmp = wtmp + (size_t)(2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER);
    int* info_ws = (int*)(qtmp + (size_t)(2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER) * len);
    int* info = info_d ? info_d : info_ws;
    Span span(PK_SYEVJ, st, 2.0 * n * n * 8, 0.0);
    if (n > rl_syevj_cluster_max_n())
        return g_knob[KNOB_EIG_GRID_FLAT] ? syevj_grid(g, ldg, n, factor_mode, tol, w, q, ldq, ws, info, st)
                                          : syevj_ring(g, ldg, n, factor_mode, tol, w, q, ldq, ws, info, st);
    // smallest cluster whose CTAs hold their pairs within 32 warps and the shared-memory budget
    // (<= 16 warps per CTA when possible: the FP64 pipe of one SM issues 2 warp-DFMAs per clock)
    int csize = 1, W = npairs;
    for (;;) {
        W = (npairs + csize - 1) / csize;
        const size_t smem = (size_t)4 * W * len * sizeof(double);
        if ((W <= 16 || (csize == JC_MAX_CLUSTER && W <= 32)) && smem <= 200 * 1024) break;
        if (csize == JC_MAX_CLUSTER) return RL_E_ARG;
        csize *= 2;
    }
    if (W < 1) W = 1;
    const int slots = 2 * csize * W;
    int rc = 0;
    if (!factor_mode) {
        jacobi_shift_kernel<<<1, 1024, 0, st>>>(g, ldg, (int)n, par);
        rc = check_launch();
        if (rc) return rc;
    }
    switch (NJ) {
        case 1: rc = launch_cluster<1>(g, ldg, (int)n, W, csize, factor_mode, tol, par, wtmp, qtmp, info, st); break;
        case 2: rc = launch_cluster<2>(g, ldg, (int)n, W, csize, factor_mode, tol, par, wtmp, qtmp, info, st); break;
        case 3: rc = launch_cluster<3>(g, ldg, (int)n, W, csize, factor_mode, tol, par, wtmp, qtmp, info, st); break;
        case 4: rc = launch_cluster<4>(g, ldg, (int)n, W, csize, factor_mode, tol, par, wtmp, qtmp, info, st); break;
        case 5: rc = launch_cluster<5>(g, ldg, (int)n, W, csize, factor_mode, tol, par, wtmp, qtmp, info, st); break;
        default: return RL_E_ARG;
    }
    if (rc) return rc;
    jacobi_sort_kernel<<<(unsigned)(n >= 128 ? 8 : 1), 1024, (size_t)slots * 12, st>>>(wtmp, qtmp, slots, len, (int)n, 0.0, w, q, ldq);
    return check_launch();
}

}  // extern "C"
