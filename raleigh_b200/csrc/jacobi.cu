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

namespace cg = cooperative_groups;

namespace rl {

constexpr int JC_MAX_SWEEPS = 48;
constexpr int JC_MAX_CLUSTER = 8;

// sigma and the convergence tolerance; one CTA
__global__ void __launch_bounds__(1024)
jacobi_shift_kernel(const double* __restrict__ G, int64_t ld, int n, double* __restrict__ par) {
    __shared__ double rlo[32], rhi[32];
    double lo = 1.0e308, hi = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double rad = 0.0;
        for (int j = 0; j < n; ++j)
            if (j != i) rad += fabs(0.5 * (G[(int64_t)i * ld + j] + G[(int64_t)j * ld + i]));
        const double d = G[(int64_t)i * ld + i];
        lo = fmin(lo, d - rad);
        hi = fmax(hi, fabs(d) + rad);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { rlo[threadIdx.x >> 5] = lo; rhi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, rlo[w]); hi = fmax(hi, rhi[w]); }
        double sigma = 0.0;
        if (lo <= 1e-3 * hi) sigma = -lo + 1e-2 * hi;     // not safely positive definite: shift
        if (!(hi > 0.0)) sigma = 1.0;                      // zero matrix: B = I
        par[0] = sigma;
        par[1] = hi;
    }
}

// NJ = (padded column length) / 64: a lane holds NJ 16-byte chunks (2 doubles) of each column,
// chunk index = lane + 32 * j -- consecutive lanes read consecutive 16-byte words (no bank conflicts).
// (at most 16 warps per CTA up to n = 256, 20 at n = 320: see rl_syevj_cluster)
template <int NJ>
__global__ void __launch_bounds__(NJ == 5 ? 640 : 512)
jacobi_cluster_kernel(const double* __restrict__ G, int64_t ldg, int n, int npairs, int W,
                      const double* __restrict__ par, double* __restrict__ wtmp, double* __restrict__ qtmp,
                      int* __restrict__ info) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char jc_smem[];
    constexpr int LEN = 64 * NJ;                                 // padded column length (doubles)
    double* buf = reinterpret_cast<double*>(jc_smem);            // [2][2W][LEN]
    __shared__ int s_rot[32];
    __shared__ int s_flag[JC_MAX_SWEEPS][JC_MAX_CLUSTER];        // meaningful in CTA 0 only
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks();
    const int k = rank * W + w;                                  // global pair index
    const bool active = w < W && k < npairs;
    const double sigma = par[0];
    const size_t bufstride = (size_t)2 * W * LEN;

    for (int i = threadIdx.x; i < JC_MAX_SWEEPS * JC_MAX_CLUSTER; i += blockDim.x) (&s_flag[0][0])[i] = 0;
    // initial columns: pair k holds columns 2k (top) and 2k+1 (bottom) of B = sym(G) + sigma I
    if (w < W) {
        for (int side = 0; side < 2; ++side) {
            const int c = 2 * k + side;
            double2* col = reinterpret_cast<double2*>(buf + ((size_t)(2 * w + side)) * LEN);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r0 = 2 * (lane + 32 * j);
                double2 v = make_double2(0.0, 0.0);
                if (k < npairs && c < n) {
                    if (r0 < n) v.x = 0.5 * (G[(int64_t)r0 * ldg + c] + G[(int64_t)c * ldg + r0]) + (r0 == c ? sigma : 0.0);
                    if (r0 + 1 < n) v.y = 0.5 * (G[(int64_t)(r0 + 1) * ldg + c] + G[(int64_t)c * ldg + r0 + 1]) + (r0 + 1 == c ? sigma : 0.0);
                }
                col[lane + 32 * j] = v;
            }
        }
    }
    cluster.sync();

    const double tol = sqrt((double)n) * 2.220446049250313e-16;
    // destination slots of the rotated columns (the same permutation every round)
    int dtk = 0, dts = 0, dbk = 0, dbs = 0;
    if (npairs > 1) {
        if (k == 0) { dtk = 0; dts = 0; dbk = 1; dbs = 0; }
        else {
            if (k == npairs - 1) { dtk = npairs - 1; dts = 1; } else { dtk = k + 1; dts = 0; }
            dbk = k - 1; dbs = 1;
        }
    } else { dtk = 0; dts = 0; dbk = 0; dbs = 1; }
    int cur = 0, sweep = 0, converged = 0;
    const int rounds = npairs > 1 ? 2 * npairs - 1 : 1;
    for (; sweep < JC_MAX_SWEEPS; ++sweep) {
        int rot = 0;
        for (int t = 0; t < rounds; ++t) {
            if (active) {
                const double2* ct = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(2 * w) * LEN);
                const double2* cb = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(2 * w + 1) * LEN);
                double2 p[NJ], q[NJ];
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    p[j] = ct[lane + 32 * j];
                    q[j] = cb[lane + 32 * j];
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
                double c = 1.0, s = 0.0;
                if (fabs(gamma) > tol * sqrt(alpha * beta) && gamma != 0.0) {
                    rot = 1;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    c = 1.0 / sqrt(1.0 + tt * tt);
                    s = c * tt;
                }
                double* dt = buf + (cur ^ 1) * bufstride + (size_t)(2 * (dtk % W) + dts) * LEN;
                double* db = buf + (cur ^ 1) * bufstride + (size_t)(2 * (dbk % W) + dbs) * LEN;
                const int rt = dtk / W, rb = dbk / W;
                double2* pt = reinterpret_cast<double2*>(rt == rank ? dt : cluster.map_shared_rank(dt, rt));
                double2* pb = reinterpret_cast<double2*>(rb == rank ? db : cluster.map_shared_rank(db, rb));
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    double2 a, b;
                    a.x = c * p[j].x - s * q[j].x; a.y = c * p[j].y - s * q[j].y;
                    b.x = s * p[j].x + c * q[j].x; b.y = s * p[j].y + c * q[j].y;
                    pt[lane + 32 * j] = a;
                    pb[lane + 32 * j] = b;
                }
            }
            if (t == rounds - 1) {
                // sweep ends: did anybody rotate?  CTA-level OR, then one plain store per CTA into CTA 0
                if (lane == 0) s_rot[w] = rot;
                __syncthreads();
                if (threadIdx.x == 0) {
                    int any = 0;
                    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) any |= s_rot[i];
                    int* f = cluster.map_shared_rank(&s_flag[sweep][rank], 0);
                    *f = any;
                }
            }
            cluster.sync();
            cur ^= 1;
        }
        int any = 0;
        {
            const int* f = cluster.map_shared_rank(&s_flag[sweep][0], 0);
            for (int r = 0; r < csize; ++r) any |= f[r];
        }
        if (!any) { converged = 1; ++sweep; break; }
    }
    // nobody may leave (and release its shared memory) while others still read the flags of CTA 0
    cluster.sync();
    // eigenvalue = column norm - sigma, eigenvector = column / norm; slots in arbitrary order, sorted later
    if (w < W) {
        for (int side = 0; side < 2; ++side) {
            const int slot = 2 * (rank * W + w) + side;
            const double2* col = reinterpret_cast<const double2*>(buf + cur * bufstride + (size_t)(2 * w + side) * LEN);
            double2 v[NJ];
            double nrm = 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { v[j] = col[lane + 32 * j]; nrm = fma(v[j].x, v[j].x, nrm); nrm = fma(v[j].y, v[j].y, nrm); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
            nrm = sqrt(nrm);
            const bool real_col = k < npairs && nrm > 0.0;
            if (lane == 0) wtmp[slot] = real_col ? nrm - sigma : 1.0e308;      // padding columns sort last
            const double inv = real_col ? 1.0 / nrm : 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r0 = 2 * (lane + 32 * j);
                if (r0 < n) qtmp[(size_t)slot * LEN + r0] = v[j].x * inv;
                if (r0 + 1 < n) qtmp[(size_t)slot * LEN + r0 + 1] = v[j].y * inv;
            }
        }
    }
    if (rank == 0 && threadIdx.x == 0) { info[0] = sweep; info[1] = converged; }
}

// ascending order: w[rank] = value, Q[r][rank] = qtmp[slot][r]
__global__ void __launch_bounds__(1024)
jacobi_sort_kernel(const double* __restrict__ wtmp, const double* __restrict__ qtmp, int slots, int len, int n,
                   double* __restrict__ w, double* __restrict__ Q, int64_t ldq) {
    extern __shared__ int s_rank[];
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        const double v = wtmp[i];
        int r = 0;
        for (int j = 0; j < slots; ++j) { const double u = wtmp[j]; r += (u < v) || (u == v && j < i); }
        s_rank[i] = r;
        if (r < n) w[r] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < slots * n; e += blockDim.x) {
        const int slot = e / n, r = e - slot * n;
        const int c = s_rank[slot];
        if (c < n) Q[(int64_t)r * ldq + c] = qtmp[(size_t)slot * len + r];
    }
}

template <int NJ>
static int launch_cluster(const double* G, int64_t ldg, int n, int npairs, int W, int csize, const double* par,
                          double* wtmp, double* qtmp, int* info, cudaStream_t st) {
    const size_t smem = (size_t)2 * 2 * W * 64 * NJ * sizeof(double);
    static size_t configured = 0;
    static bool nonportable = false;
    if (smem > configured) {
        RL_CUDA(cudaFuncSetAttribute(jacobi_cluster_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    (void)nonportable;
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
    return (int)cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<NJ>, G, ldg, n, npairs, W, par, wtmp, qtmp, info);
}

}  // namespace rl

using namespace rl;

extern "C" {

int rl_syevj_cluster_max_n(void) { return 320; }

/* workspace: par (8 doubles) | wtmp (2*npairs_padded) | qtmp (slots * len) | info (4 ints) */
size_t rl_syevj_cluster_ws_bytes(int64_t n) {
    if (n <= 0) return 0;
    const int64_t len = (n + 63) / 64 * 64;
    const int64_t slots = 2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER;
    return (size_t)(8 + slots + slots * len) * sizeof(double) + 64;
}

/* Eigen-decomposition of the symmetric n x n fp64 matrix g (row-major, ldg; both triangles are
 * read and averaged): w[0..n) ascending, q[i*ldq + j] = component i of eigenvector j.
 * g is not modified.  info_d (device, 2 ints): sweeps, converged.  n <= rl_syevj_cluster_max_n(). */
int rl_syevj_cluster(const double* g, int64_t ldg, int64_t n, double* w, double* q, int64_t ldq, void* ws,
                     size_t ws_bytes, int* info_d, void* stream) {
    if (n < 0 || n > rl_syevj_cluster_max_n()) return RL_E_ARG;
    if (n == 0) return 0;
    if (ws_bytes < rl_syevj_cluster_ws_bytes(n)) return RL_E_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int npairs = (int)((n + 1) / 2);
    const int NJ = (int)((n + 63) / 64);
    const int len = 64 * NJ;
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
    double* par = (double*)ws;
    double* wtmp = par + 8;
    const int slots = 2 * csize * W;
    double* qtmp = wtmp + (size_t)(2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER);
    int* info_ws = (int*)(qtmp + (size_t)(2 * ((n + 1) / 2) + 2 * 32 * JC_MAX_CLUSTER) * len);
    int* info = info_d ? info_d : info_ws;
    Span span(PK_SYEVJ, st, 2.0 * n * n * 8, 0.0);
    jacobi_shift_kernel<<<1, 1024, 0, st>>>(g, ldg, (int)n, par);
    int rc = check_launch();
    if (rc) return rc;
    switch (NJ) {
        case 1: rc = launch_cluster<1>(g, ldg, (int)n, npairs, W, csize, par, wtmp, qtmp, info, st); break;
        case 2: rc = launch_cluster<2>(g, ldg, (int)n, npairs, W, csize, par, wtmp, qtmp, info, st); break;
        case 3: rc = launch_cluster<3>(g, ldg, (int)n, npairs, W, csize, par, wtmp, qtmp, info, st); break;
        case 4: rc = launch_cluster<4>(g, ldg, (int)n, npairs, W, csize, par, wtmp, qtmp, info, st); break;
        case 5: rc = launch_cluster<5>(g, ldg, (int)n, npairs, W, csize, par, wtmp, qtmp, info, st); break;
        default: return RL_E_ARG;
    }
    if (rc) return rc;
    jacobi_sort_kernel<<<1, 1024, (size_t)slots * sizeof(int), st>>>(wtmp, qtmp, slots, len, (int)n, w, q, ldq);
    return check_launch();
}

}  // extern "C"
