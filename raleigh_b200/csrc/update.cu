// Block update  Out[j,:] = beta*Out[j,:] + alpha * sum_{i<k} Q[i,j] * X[i,:]
// (Vectors.multiply: beta = 0, dense_cublas.py:271-299;  Vectors.add(other, s, q):
// beta = 1, dense_cublas.py:317-342).  X is (k, n), Out is (m, n), vector-major;
// Q is a small (k, m) coefficient matrix with arbitrary element strides.
//
// Algorithmic traffic: n*(k + m)*w bytes (beta = 0) or n*(k + 2m)*w (beta != 0);
// 2*n*k*m flops.  HBM-bound up to m,k ~ 32 in fp64, FP64-pipe-bound beyond.
//
// Mapping: a thread owns VEC consecutive rows (one 128-bit load per X vector)
// and TJ output vectors; the Q tile sits in shared memory and is read with
// warp-broadcast 128-bit loads, so per X element loaded a thread issues TJ*VEC
// FMAs and TJ*w/16 shared loads.  The CTA's j-groups share the same rows, so X
// is fetched from HBM once and re-served by L1 to the other j-groups.
#include "common.cuh"

namespace rl {

constexpr int UPD_THREADS = 256;
constexpr int UPD_TJ = 16;          // outputs per thread
constexpr int UPD_JB = 64;          // outputs per CTA (4 j-groups)
constexpr int UPD_KT = 32;          // Q rows staged per shared-memory tile

template <typename T>
__global__ void __launch_bounds__(UPD_THREADS, 2)
update_kernel(T* __restrict__ Out, int64_t ldo, int m, const T* __restrict__ X, int64_t ldx, int k,
              const T* __restrict__ Q, int64_t q_rs, int64_t q_cs, T alpha, T beta, int64_t n, int jgroups,
              int fast) {
    constexpr int V = Vec128<T>::N;
    using VT = typename Vec128<T>::type;
    __shared__ __align__(16) T Qs[UPD_KT][UPD_JB];

    const int j_base = blockIdx.y * UPD_JB;
    const int rthreads = UPD_THREADS / jgroups;            // threads along rows
    const int jg = threadIdx.x / rthreads;                 // this thread's j-group
    const int rt = threadIdx.x - jg * rthreads;
    const int64_t r = ((int64_t)blockIdx.x * rthreads + rt) * V;
    const int j0 = j_base + jg * UPD_TJ;
    const bool row_ok = r < n;
    const bool full = fast && (r + V <= n);

    T acc[UPD_TJ][V];
#pragma unroll
    for (int j = 0; j < UPD_TJ; ++j)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[j][v] = T(0);

    for (int i0 = 0; i0 < k; i0 += UPD_KT) {
        const int kt = k - i0 < UPD_KT ? k - i0 : UPD_KT;
        __syncthreads();
        for (int e = threadIdx.x; e < UPD_KT * UPD_JB; e += UPD_THREADS) {
            int ii = e / UPD_JB, jj = e % UPD_JB;
            int j = j_base + jj;
            Qs[ii][jj] = (ii < kt && j < m) ? __ldg(Q + (int64_t)(i0 + ii) * q_rs + (int64_t)j * q_cs) : T(0);
        }
        __syncthreads();
        if (!row_ok) continue;
        const T* px = X + (int64_t)i0 * ldx + r;
        if (full) {
            int ii = 0;
            for (; ii + 4 <= kt; ii += 4) {
                VT xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[u] = __ldg(reinterpret_cast<const VT*>(px + (int64_t)(ii + u) * ldx));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const T* xe = reinterpret_cast<const T*>(&xv[u]);
#pragma unroll
                    for (int j = 0; j < UPD_TJ; j += V) {
                        VT qv = *reinterpret_cast<const VT*>(&Qs[ii + u][jg * UPD_TJ + j]);
                        const T* qe = reinterpret_cast<const T*>(&qv);
#pragma unroll
                        for (int t = 0; t < V; ++t)
#pragma unroll
                            for (int v = 0; v < V; ++v) acc[j + t][v] = fma(qe[t], xe[v], acc[j + t][v]);
                    }
                }
            }
            for (; ii < kt; ++ii) {
                VT xv = __ldg(reinterpret_cast<const VT*>(px + (int64_t)ii * ldx));
                const T* xe = reinterpret_cast<const T*>(&xv);
#pragma unroll
                for (int j = 0; j < UPD_TJ; ++j) {
                    T qv = Qs[ii][jg * UPD_TJ + j];
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[j][v] = fma(qv, xe[v], acc[j][v]);
                }
            }
        } else {
            for (int ii = 0; ii < kt; ++ii) {
                T xe[V];
#pragma unroll
                for (int v = 0; v < V; ++v) xe[v] = (r + v < n) ? __ldg(px + (int64_t)ii * ldx + v) : T(0);
#pragma unroll
                for (int j = 0; j < UPD_TJ; ++j) {
                    T qv = Qs[ii][jg * UPD_TJ + j];
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[j][v] = fma(qv, xe[v], acc[j][v]);
                }
            }
        }
    }
    if (!row_ok) return;
#pragma unroll
    for (int j = 0; j < UPD_TJ; ++j) {
        int jj = j0 + j;
        if (jj >= m) break;
        T* po = Out + (int64_t)jj * ldo + r;
        if (full) {
            VT res;
            T* re = reinterpret_cast<T*>(&res);
            if (beta != T(0)) {
                VT old = *reinterpret_cast<const VT*>(po);
                const T* oe = reinterpret_cast<const T*>(&old);
#pragma unroll
                for (int v = 0; v < V; ++v) re[v] = fma(alpha, acc[j][v], beta * oe[v]);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) re[v] = alpha * acc[j][v];
            }
            *reinterpret_cast<VT*>(po) = res;
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (r + v < n) po[v] = beta != T(0) ? fma(alpha, acc[j][v], beta * po[v]) : alpha * acc[j][v];
        }
    }
}

// ---- fp64 tensor-pipe variant ----------------------------------------------------
// D(j, r) = sum_i Q[i, j] X[i, r] as DMMA m8n8k4 with A = Q^T (from shared memory) and
// B = X straight from global memory: lane (g, c) reads the four consecutive rows
// R+4g..R+4g+3 of vector i0+c with ONE 256-bit load and feeds them to four n-tiles, so a
// warp step covers 32 rows x 4 input vectors with 4 fully used 256-byte requests.  The
// C fragments of the four n-tiles give each lane 8 consecutive rows of one output
// vector: two 256-bit stores.  No per-FMA shared-memory traffic (the FMA-pipe kernel
// above is co-limited by LDS, fp64 pipe and DRAM at m = k = 32, profiles/r1b_gram_ncu.md).
constexpr int UD_WARPS = 4;
constexpr int UD_JB = 32;            // outputs per CTA
constexpr int UD_KMAX = 512;

__device__ __forceinline__ void ud_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void ld256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld256_rw(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

template <int NJ>
__global__ void __launch_bounds__(UD_WARPS * 32)
update_dmma_kernel(double* __restrict__ Out, int64_t ldo, int m, const double* __restrict__ X, int64_t ldx, int k,
                   const double* __restrict__ Q, int64_t q_rs, int64_t q_cs, double alpha, double beta, int64_t n,
                   int kp, int64_t blocks_per_cta) {
    extern __shared__ __align__(16) double Qs[];          // [UD_JB][kp], kp = 4 mod 16: i contiguous
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int j_base = blockIdx.y * UD_JB;
    const int ksteps = (k + 3) >> 2;
    for (int e = threadIdx.x; e < UD_JB * kp; e += UD_WARPS * 32) {
        const int jj = e / kp, ii = e - jj * kp;
        const int j = j_base + jj;
        Qs[e] = (ii < k && j < m) ? alpha * __ldg(Q + (int64_t)ii * q_rs + (int64_t)j * q_cs) : 0.0;
    }
    __syncthreads();

    const int64_t nblocks = (n + 31) >> 5;                // 32-row blocks
    const int64_t b_begin = (int64_t)blockIdx.x * blocks_per_cta;
    const int64_t b_end = b_begin + blocks_per_cta < nblocks ? b_begin + blocks_per_cta : nblocks;
    for (int64_t blk = b_begin + warp; blk < b_end; blk += UD_WARPS) {
        const int64_t R = blk << 5;
        const bool full = R + 32 <= n;
        double acc[NJ][4][2];
#pragma unroll
        for (int t = 0; t < NJ; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
        const double* px = X + R + 4 * g;
        if (full) {
            int ks = 0;
            for (; ks + 4 <= ksteps; ks += 4) {            // 4 k-steps = 16 input vectors, 4 loads in flight
                double xb[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = 4 * (ks + q) + c;
                    if (i < k) ld256(px + (int64_t)i * ldx, xb[q]);
                    else xb[q][0] = xb[q][1] = xb[q][2] = xb[q][3] = 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int t = 0; t < NJ; ++t) {
                        const double a = Qs[(8 * t + g) * kp + 4 * (ks + q) + c];
#pragma unroll
                        for (int u = 0; u < 4; ++u) ud_dmma(acc[t][u][0], acc[t][u][1], a, xb[q][u]);
                    }
                }
            }
            for (; ks < ksteps; ++ks) {
                double xb[4];
                const int i = 4 * ks + c;
                if (i < k) ld256(px + (int64_t)i * ldx, xb);
                else xb[0] = xb[1] = xb[2] = xb[3] = 0.0;
#pragma unroll
                for (int t = 0; t < NJ; ++t) {
                    const double a = Qs[(8 * t + g) * kp + 4 * ks + c];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ud_dmma(acc[t][u][0], acc[t][u][1], a, xb[u]);
                }
            }
        } else {
            for (int ks = 0; ks < ksteps; ++ks) {
                double xb[4];
                const int i = 4 * ks + c;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t r = R + 4 * g + u;
                    xb[u] = (i < k && r < n) ? __ldg(X + (int64_t)i * ldx + r) : 0.0;
                }
#pragma unroll
                for (int t = 0; t < NJ; ++t) {
                    const double a = Qs[(8 * t + g) * kp + 4 * ks + c];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ud_dmma(acc[t][u][0], acc[t][u][1], a, xb[u]);
                }
            }
        }
        // n-tile u, C fragment (g, c): output vector j = 8t+g, rows R + 4*(2c) + u and R + 4*(2c+1) + u
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            const int j = j_base + 8 * t + g;
            if (j >= m) continue;
            double* po = Out + (int64_t)j * ldo + R + 8 * c;
            double lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { lo[u] = acc[t][u][0]; hi[u] = acc[t][u][1]; }
            if (full) {
                if (beta != 0.0) {
                    double ol[4], oh[4];
                    ld256_rw(po, ol);
                    ld256_rw(po + 4, oh);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { lo[u] = fma(beta, ol[u], lo[u]); hi[u] = fma(beta, oh[u], hi[u]); }
                }
                st256(po, lo);
                st256(po + 4, hi);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t r0 = R + 8 * c + u, r1 = r0 + 4;
                    if (r0 < n) po[u] = beta != 0.0 ? fma(beta, po[u], lo[u]) : lo[u];
                    if (r1 < n) po[u + 4] = beta != 0.0 ? fma(beta, po[u + 4], hi[u]) : hi[u];
                }
            }
        }
    }
}

static bool update_dmma_ok(const void* out, int64_t ldo, const void* x, int64_t ldx, int64_t k) {
    return k <= UD_KMAX && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x)) & 31) == 0 &&
           (ldo % 4 == 0) && (ldx % 4 == 0);
}

static int update_dmma(void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx, int64_t k, const void* q,
                       int64_t q_rs, int64_t q_cs, double alpha, double beta, int64_t n, cudaStream_t st) {
    int kp = (int)((k + 3) / 4 * 4);
    while (kp % 16 != 4) kp += 4;                          // bank spread for the 64-bit fragment reads
    const size_t smem = (size_t)UD_JB * kp * sizeof(double);
    const int jblocks = (int)((m + UD_JB - 1) / UD_JB);
    const int64_t nblocks = (n + 31) / 32;
    // ~8 CTAs per SM in total, at least 4 row blocks per warp
    int64_t want = ((int64_t)sm_count() * 8 + jblocks - 1) / jblocks;
    int64_t maxc = (nblocks + UD_WARPS * 4 - 1) / (UD_WARPS * 4);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    const int64_t bpc = (nblocks + want - 1) / want;
    const unsigned gx = (unsigned)((nblocks + bpc - 1) / bpc);
    const int nj = (int)(((m < UD_JB ? m : UD_JB) + 7) / 8);
    static size_t configured[5] = {0, 0, 0, 0, 0};
#define RL_UD_LAUNCH(NJ_)                                                                                        \
    do {                                                                                                         \
        if (smem > 48 * 1024 && smem > configured[NJ_]) {                                                        \
            RL_CUDA(cudaFuncSetAttribute(update_dmma_kernel<NJ_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int)smem));                                                            \
            configured[NJ_] = smem;                                                                              \
        }                                                                                                        \
        update_dmma_kernel<NJ_><<<dim3(gx, (unsigned)jblocks), UD_WARPS * 32, smem, st>>>(                       \
            (double*)out, ldo, (int)m, (const double*)x, ldx, (int)k, (const double*)q, q_rs, q_cs, alpha, beta, \
            n, kp, bpc);                                                                                         \
    } while (0)
    if (jblocks > 1 || nj == 4) RL_UD_LAUNCH(4);
    else if (nj == 3) RL_UD_LAUNCH(3);
    else if (nj == 2) RL_UD_LAUNCH(2);
    else RL_UD_LAUNCH(1);
#undef RL_UD_LAUNCH
    return check_launch();
}

template <typename T>
static int update_impl(void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx, int64_t k, const void* q,
                       int64_t q_rs, int64_t q_cs, double alpha, double beta, int64_t n, cudaStream_t st) {
    constexpr int V = Vec128<T>::N;
    int jb = (int)((m + UPD_JB - 1) / UPD_JB);
    int jgroups = (int)((((m < UPD_JB) ? m : UPD_JB) + UPD_TJ - 1) / UPD_TJ);   // 1..4
    if (jgroups == 3) jgroups = 4;                                              // keep 256 % jgroups == 0
    int rthreads = UPD_THREADS / jgroups;
    int64_t rows_per_cta = (int64_t)rthreads * V;
    int64_t gx = (n + rows_per_cta - 1) / rows_per_cta;
    int fast = host_aligned16(out) && host_aligned16(x) && (ldo % V == 0) && (ldx % V == 0);
    dim3 grid((unsigned)gx, (unsigned)jb);
    update_kernel<T><<<grid, UPD_THREADS, 0, st>>>((T*)out, ldo, (int)m, (const T*)x, ldx, (int)k, (const T*)q,
                                                  q_rs, q_cs, (T)alpha, (T)beta, n, jgroups, fast);
    return check_launch();
}

// beta-only pass (k == 0): Out = beta*Out
template <typename T>
__global__ void scale_all_kernel(T* out, int64_t ldo, int64_t m, int64_t n, T beta) {
    for (int64_t j = blockIdx.y; j < m; j += gridDim.y)
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
            out[j * ldo + r] = beta == T(0) ? T(0) : beta * out[j * ldo + r];
}

}  // namespace rl

using namespace rl;

extern "C" {

static int g_update_force_fma = 0;
void rl_debug_set_update_fma(int on) { g_update_force_fma = on; }

int rl_update(int dtype, void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx, int64_t k, const void* q,
              int64_t q_rs, int64_t q_cs, double alpha, double beta, int64_t n, void* stream) {
    if (m < 0 || k < 0 || n < 0 || m > INT32_MAX || k > INT32_MAX) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    if (dtype != RL_F32 && dtype != RL_F64) return RL_E_DTYPE;
    cudaStream_t st = as_stream(stream);
    if (k == 0 || alpha == 0.0) {
        if (beta == 1.0) return 0;
        int64_t gx = (n + 255) / 256; if (gx > 1024) gx = 1024;
        dim3 g((unsigned)gx, (unsigned)(m < 65535 ? m : 65535));
        if (dtype == RL_F32) scale_all_kernel<float><<<g, 256, 0, st>>>((float*)out, ldo, m, n, (float)beta);
        else scale_all_kernel<double><<<g, 256, 0, st>>>((double*)out, ldo, m, n, beta);
        return check_launch();
    }
    {   // Out must not overlap X: compare the byte ranges of the two windows, not just the base pointers
        // (select() sub-blocks and shallow references of one buffer; NumPy materialises a temporary)
        const size_t w = dtype == RL_F32 ? 4 : 8;
        const char* o0 = (const char*)out; const char* o1 = o0 + ((size_t)(m - 1) * ldo + n) * w;
        const char* x0 = (const char*)x;   const char* x1 = x0 + ((size_t)(k - 1) * ldx + n) * w;
        if (o0 < x1 && x0 < o1) return RL_E_ALIAS;
    }
    Span span(PK_UPDATE, st, (1.0 * k + (beta != 0.0 ? 2.0 : 1.0) * m) * n * (dtype == RL_F32 ? 4 : 8),
              2.0 * n * k * m);
    if (dtype == RL_F32) return update_impl<float>(out, ldo, m, x, ldx, k, q, q_rs, q_cs, alpha, beta, n, st);
    // measured on B200 (profiles/r1c_kernel_tuning.md): DMMA wins everywhere (m = 32: 84 % vs 47 % of HBM
    // peak, m = 120: 28 vs 14 TFLOP/s) except the smallest beta = 0 blocks (m, k <= 16: 61 % vs 72 %)
    const bool small_overwrite = beta == 0.0 && m <= 16 && k <= 16 && m > 8;
    if (!g_update_force_fma && !small_overwrite && update_dmma_ok(out, ldo, x, ldx, k))
        return update_dmma(out, ldo, m, x, ldx, k, q, q_rs, q_cs, alpha, beta, n, st);
    return update_impl<double>(out, ldo, m, x, ldx, k, q, q_rs, q_cs, alpha, beta, n, st);
}

int rl_update_h(int dtype, void* out, int64_t ldo, int64_t m, const void* x, int64_t ldx, int64_t k,
                const void* q_h, int64_t q_rs, int64_t q_cs, double alpha, double beta, int64_t n, void* stream) {
    if (m < 0 || k < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    if (k == 0) return rl_update(dtype, out, ldo, m, x, ldx, 0, nullptr, 0, 0, alpha, beta, n, stream);
    if (q_rs < 0 || q_cs < 0) return RL_E_ARG;
    // repack the (k, m) coefficients densely (row-major) into the staging ring
    void *pinned = nullptr, *dev = nullptr;
    int rc = staging_acquire((size_t)k * m * w, &pinned, &dev);
    if (rc) return rc;
    if (w == 8) {
        const double* src = (const double*)q_h; double* dst = (double*)pinned;
        for (int64_t i = 0; i < k; ++i)
            for (int64_t j = 0; j < m; ++j) dst[i * m + j] = src[i * q_rs + j * q_cs];
    } else {
        const float* src = (const float*)q_h; float* dst = (float*)pinned;
        for (int64_t i = 0; i < k; ++i)
            for (int64_t j = 0; j < m; ++j) dst[i * m + j] = src[i * q_rs + j * q_cs];
    }
    RL_CUDA(cudaMemcpyAsync(dev, pinned, (size_t)k * m * w, cudaMemcpyHostToDevice, as_stream(stream)));
    return rl_update(dtype, out, ldo, m, x, ldx, k, dev, m, 1, alpha, beta, n, stream);
}

}  // extern "C"
