// Streaming (HBM-bound) block-vector kernels: copy/gather, axpy, per-vector
// axpy, scale, row/column dot products, diagonal scaling, counter-based fill.
// One launch per Vectors method, where the reference issues one cuBLAS call
// PER VECTOR (dense_cublas.py:149-153, 163-172, 233-243, 344-350).
//
// Roofline: every kernel here moves each element once: 2*n*m*w bytes (scale,
// copy, dots) or 3*n*m*w (axpy variants).  128-bit accesses, 4 independent
// vector items per thread in flight.
#include "common.cuh"

namespace rl {

constexpr int EW_THREADS = 256;
constexpr int EW_ITEMS = 4;   // independent 128-bit items per thread

template <typename T>
static inline bool vec_ok(const void* a, int64_t lda, const void* b = nullptr, int64_t ldb = 0) {
    constexpr int V = Vec128<T>::N;
    bool ok = host_aligned16(a) && (lda % V == 0);
    if (b) ok = ok && host_aligned16(b) && (ldb % V == 0);
    return ok;
}

static inline dim3 ew_grid(int64_t m, int64_t items_per_row) {
    int64_t per_block = (int64_t)EW_THREADS * EW_ITEMS;
    int64_t gx = (items_per_row + per_block - 1) / per_block;
    if (gx < 1) gx = 1;
    int64_t gy = m < 65535 ? m : 65535;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

// Generic 2-D elementwise driver.  Op::vec(j, r) handles Vec128<T>::N elements
// starting at component r of vector j; Op::one(j, r) a single element.
template <typename T, bool VEC, typename Op>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(Op op, int64_t m, int64_t n) {
    constexpr int V = VEC ? Vec128<T>::N : 1;
    const int64_t nfull = n / V;                 // full vector items per row
    const int64_t base = ((int64_t)blockIdx.x * EW_THREADS * EW_ITEMS) + threadIdx.x;
    for (int64_t j = blockIdx.y; j < m; j += gridDim.y) {
#pragma unroll
        for (int u = 0; u < EW_ITEMS; ++u) {
            int64_t it = base + (int64_t)u * EW_THREADS;
            if (it < nfull) {
                if (VEC) op.vec(j, it * V); else op.one(j, it);
            }
        }
        if (VEC) {  // ragged tail (n % V elements), done by block 0 of the row
            if (blockIdx.x == 0 && threadIdx.x < (n - nfull * V)) op.one(j, nfull * V + threadIdx.x);
        }
    }
}

template <typename T, typename Op>
static int launch_ew(Op op, int64_t m, int64_t n, bool vec, cudaStream_t st) {
    if (m <= 0 || n <= 0) return 0;
    constexpr int V = Vec128<T>::N;
    if (vec) {
        dim3 g = ew_grid(m, n / V > 0 ? n / V : 1);
        ew_kernel<T, true, Op><<<g, EW_THREADS, 0, st>>>(op, m, n);
    } else {
        dim3 g = ew_grid(m, n);
        ew_kernel<T, false, Op><<<g, EW_THREADS, 0, st>>>(op, m, n);
    }
    return check_launch();
}

template <typename T> using V128 = typename Vec128<T>::type;

__device__ __forceinline__ double2 fma2(double a, double2 x, double2 y) {
    return make_double2(fma(a, x.x, y.x), fma(a, x.y, y.y));
}
__device__ __forceinline__ float4 fma2(float a, float4 x, float4 y) {
    return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ double2 mul2(double a, double2 x) { return make_double2(a * x.x, a * x.y); }
__device__ __forceinline__ float4 mul2(float a, float4 x) {
    return make_float4(a * x.x, a * x.y, a * x.z, a * x.w);
}
__device__ __forceinline__ double2 mulv(double2 a, double2 x) { return make_double2(a.x * x.x, a.y * x.y); }
__device__ __forceinline__ float4 mulv(float4 a, float4 x) {
    return make_float4(a.x * x.x, a.y * x.y, a.z * x.z, a.w * x.w);
}

template <typename T>
struct CopyOp {
    T* y; const T* x; int64_t ldy, ldx;
    __device__ void vec(int64_t j, int64_t r) const {
        *reinterpret_cast<V128<T>*>(y + j * ldy + r) = ldg_stream(reinterpret_cast<const V128<T>*>(x + j * ldx + r));
    }
    __device__ void one(int64_t j, int64_t r) const { y[j * ldy + r] = x[j * ldx + r]; }
};

template <typename T>
struct AxpyOp {   // y += a*x
    T* y; const T* x; int64_t ldy, ldx; T a;
    __device__ void vec(int64_t j, int64_t r) const {
        V128<T>* py = reinterpret_cast<V128<T>*>(y + j * ldy + r);
        V128<T> xv = ldg_stream(reinterpret_cast<const V128<T>*>(x + j * ldx + r));
        *py = fma2(a, xv, *py);
    }
    __device__ void one(int64_t j, int64_t r) const { y[j * ldy + r] += a * x[j * ldx + r]; }
};

template <typename T>
struct AxpyDiagOp {   // y[j] += s[j]*x[j]
    T* y; const T* x; int64_t ldy, ldx; const T* s;
    __device__ void vec(int64_t j, int64_t r) const {
        T a = __ldg(s + j);
        V128<T>* py = reinterpret_cast<V128<T>*>(y + j * ldy + r);
        V128<T> xv = ldg_stream(reinterpret_cast<const V128<T>*>(x + j * ldx + r));
        *py = fma2(a, xv, *py);
    }
    __device__ void one(int64_t j, int64_t r) const { y[j * ldy + r] += __ldg(s + j) * x[j * ldx + r]; }
};

template <typename T>
struct ScaleOp {      // y[j] *= s[j]   or   y[j] /= s[j] unless s[j] == 0
    T* y; int64_t ldy; const T* s; int multiply;
    __device__ T factor(int64_t j, bool& skip) const {
        T a = __ldg(s + j);
        skip = false;
        if (!multiply) {
            if (a == T(0)) { skip = true; return T(1); }
        }
        return a;
    }
    __device__ void vec(int64_t j, int64_t r) const {
        bool skip; T a = factor(j, skip);
        if (skip) return;
        V128<T>* py = reinterpret_cast<V128<T>*>(y + j * ldy + r);
        V128<T> v = *py;
        if (multiply) { *py = mul2(a, v); }
        else {
            // true division (not multiplication by a reciprocal) to match the
            // NumPy oracle bit-for-bit: dense_numpy.py:51-52
            T* e = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int t = 0; t < Vec128<T>::N; ++t) e[t] = e[t] / a;
            *py = v;
        }
    }
    __device__ void one(int64_t j, int64_t r) const {
        bool skip; T a = factor(j, skip);
        if (skip) return;
        T& e = y[j * ldy + r];
        e = multiply ? e * a : e / a;
    }
};

template <typename T>
struct DiagMulOp {    // y[j, r] = x[j, r] * d[r]
    T* y; const T* x; int64_t ldy, ldx; const T* d;
    __device__ void vec(int64_t j, int64_t r) const {
        V128<T> dv = *reinterpret_cast<const V128<T>*>(d + r);
        V128<T> xv = ldg_stream(reinterpret_cast<const V128<T>*>(x + j * ldx + r));
        *reinterpret_cast<V128<T>*>(y + j * ldy + r) = mulv(dv, xv);
    }
    __device__ void one(int64_t j, int64_t r) const { y[j * ldy + r] = x[j * ldx + r] * __ldg(d + r); }
};

// gather: dst[t] <- src[ind[t]]; indices travel as kernel parameters
constexpr int GATHER_MAX = 128;
struct GatherIdx { int64_t v[GATHER_MAX]; };
template <typename T>
struct GatherOp {
    T* y; const T* x; int64_t ldy, ldx; GatherIdx idx;
    __device__ void vec(int64_t j, int64_t r) const {
        *reinterpret_cast<V128<T>*>(y + j * ldy + r) =
            ldg_stream(reinterpret_cast<const V128<T>*>(x + idx.v[j] * ldx + r));
    }
    __device__ void one(int64_t j, int64_t r) const { y[j * ldy + r] = x[idx.v[j] * ldx + r]; }
};

// ---- Philox4x32-10 counter RNG ------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ void philox(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int i = 0; i < 10; ++i) philox_round(c, k);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// element (J, R) of the global block gets word (R % 4) of philox(seed, R / 4, J)
// (fp32: one word; fp64: the pair (R%2)*2, +1 of philox(seed, R / 2, J)).
template <typename T>
__global__ void __launch_bounds__(256) fill_uniform_kernel(T* x, int64_t ld, int64_t m, int64_t n,
                                                           uint64_t seed, int64_t j0, int64_t r0) {
    constexpr int PER = sizeof(T) == 4 ? 4 : 2;
    const int64_t R0 = r0 / PER * PER;                      // aligned start of the quad/pair grid
    const int64_t groups = (r0 + n - R0 + PER - 1) / PER;
    for (int64_t j = blockIdx.y; j < m; j += gridDim.y) {
        for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
             g += (int64_t)gridDim.x * blockDim.x) {
            int64_t Rg = R0 + g * PER;
            uint32_t w[4];
            philox(seed, (uint64_t)(Rg / PER), (uint64_t)(j0 + j), w);
#pragma unroll
            for (int t = 0; t < PER; ++t) {
                int64_t R = Rg + t;
                if (R < r0 || R >= r0 + n) continue;
                T u;
                if (sizeof(T) == 4) {
                    u = (T)((w[t] >> 8) * (1.0f / 16777216.0f));           // 24 bits
                } else {
                    uint64_t bits = ((uint64_t)w[2 * t + 1] << 32) | w[2 * t];
                    u = (T)((bits >> 11) * (1.0 / 9007199254740992.0));    // 53 bits
                }
                x[j * ld + (R - r0)] = T(2) * u - T(1);
            }
        }
    }
}

// ---- row dots: w[i] = sum_r o[i,r]*s[i,r] ---------------------------------------
constexpr int DOT_THREADS = 256;

template <typename T>
__device__ __forceinline__ T block_sum(T v) {
    __shared__ T red[DOT_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T total = T(0);
    if (threadIdx.x < 32) {
        T x = threadIdx.x < DOT_THREADS / 32 ? red[threadIdx.x] : T(0);
        total = warp_sum(x);
    }
    __syncthreads();
    return total;   // valid in warp 0
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(DOT_THREADS) dots_partial_kernel(
    const T* __restrict__ s, int64_t lds, const T* __restrict__ o, int64_t ldo, int64_t m, int64_t n,
    int64_t chunk, int chunks, T* __restrict__ out) {
    constexpr int V = VEC ? Vec128<T>::N : 1;
    const int c = blockIdx.x;
    const int64_t lo = (int64_t)c * chunk;
    const int64_t hi = lo + chunk < n ? lo + chunk : n;
    for (int64_t i = blockIdx.y; i < m; i += gridDim.y) {
        const T* ps = s + i * lds;
        const T* po = o + i * ldo;
        T acc[4] = {T(0), T(0), T(0), T(0)};
        if (VEC) {
            const int64_t vlo = lo / V, vhi = hi / V;      // chunk is a multiple of V
            int64_t it = vlo + threadIdx.x;
            for (; it + 3 * DOT_THREADS < vhi; it += 4 * DOT_THREADS) {
                V128<T> a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] = ldg_stream(reinterpret_cast<const V128<T>*>(ps) + it + u * DOT_THREADS);
                    b[u] = ldg_stream(reinterpret_cast<const V128<T>*>(po) + it + u * DOT_THREADS);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const T* ea = reinterpret_cast<const T*>(&a[u]);
                    const T* eb = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
                    for (int t = 0; t < V; ++t) acc[u] = fma(ea[t], eb[t], acc[u]);
                }
            }
            for (; it < vhi; it += DOT_THREADS) {
                V128<T> a = ldg_stream(reinterpret_cast<const V128<T>*>(ps) + it);
                V128<T> b = ldg_stream(reinterpret_cast<const V128<T>*>(po) + it);
                const T* ea = reinterpret_cast<const T*>(&a);
                const T* eb = reinterpret_cast<const T*>(&b);
#pragma unroll
                for (int t = 0; t < V; ++t) acc[0] = fma(ea[t], eb[t], acc[0]);
            }
            if (hi == n) {   // ragged tail of the row
                int64_t r = vhi * V + threadIdx.x;
                if (r < n) acc[1] = fma(ps[r], po[r], acc[1]);
            }
        } else {
            for (int64_t r = lo + threadIdx.x; r < hi; r += DOT_THREADS) acc[0] = fma(ps[r], po[r], acc[0]);
        }
        T v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        v = block_sum(v);
        if (threadIdx.x == 0) out[i * chunks + c] = v;
    }
}

// fixed-order final sum over the chunk partials of each vector (one warp each)
template <typename T>
__global__ void __launch_bounds__(256) dots_final_kernel(const T* __restrict__ part, int64_t m, int chunks,
                                                         T* __restrict__ w) {
    int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m) return;
    int lane = threadIdx.x & 31;
    T acc = T(0);
    for (int c = lane; c < chunks; c += 32) acc += part[i * chunks + c];
    acc = warp_sum(acc);
    if (lane == 0) w[i] = acc;
}

static void dots_plan(int64_t m, int64_t n, int vec, int64_t* chunk, int* chunks) {
    // aim for ~8 CTAs per SM in total, but never chunks shorter than 8192 elements
    int64_t target = (int64_t)sm_count() * 8;
    int64_t c = (target + m - 1) / m;
    int64_t maxc = (n + 8191) / 8192;
    if (c > maxc) c = maxc;
    if (c < 1) c = 1;
    if (c > 1024) c = 1024;
    int64_t ch = (n + c - 1) / c;
    ch = (ch + 4 * vec - 1) / (4 * vec) * (4 * vec);
    c = (n + ch - 1) / ch;
    *chunk = ch;
    *chunks = (int)c;
}

template <typename T>
static int dots_impl(const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n, void* w,
                     void* ws, size_t ws_bytes, cudaStream_t st) {
    if (m <= 0) return 0;
    if (n <= 0) return (int)cudaMemsetAsync(w, 0, m * sizeof(T), st);
    bool vec = vec_ok<T>(s, lds, o, ldo);
    int64_t chunk; int chunks;
    dots_plan(m, n, Vec128<T>::N, &chunk, &chunks);
    T* part = reinterpret_cast<T*>(w);
    if (chunks > 1) {
        if (ws_bytes < (size_t)m * chunks * sizeof(T)) return RL_E_WORKSPACE;
        part = reinterpret_cast<T*>(ws);
    }
    dim3 g((unsigned)chunks, (unsigned)(m < 65535 ? m : 65535));
    if (vec)
        dots_partial_kernel<T, true><<<g, DOT_THREADS, 0, st>>>((const T*)s, lds, (const T*)o, ldo, m, n, chunk, chunks, part);
    else
        dots_partial_kernel<T, false><<<g, DOT_THREADS, 0, st>>>((const T*)s, lds, (const T*)o, ldo, m, n, chunk, chunks, part);
    int rc = check_launch();
    if (rc) return rc;
    if (chunks > 1) {
        int64_t blocks = (m + 7) / 8;
        dots_final_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(part, m, chunks, (T*)w);
        rc = check_launch();
    }
    return rc;
}

// ---- column dots: w[r] = sum_i o[i,r]*s[i,r] -------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dots_t_kernel(const T* __restrict__ s, int64_t lds,
                                                     const T* __restrict__ o, int64_t ldo, int64_t m, int64_t n,
                                                     int64_t ichunk, T* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int64_t i0 = (int64_t)blockIdx.y * ichunk;
    int64_t i1 = i0 + ichunk < m ? i0 + ichunk : m;
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    int64_t i = i0;
    for (; i + 3 < i1; i += 4) {
        a0 = fma(__ldg(o + (i + 0) * ldo + r), __ldg(s + (i + 0) * lds + r), a0);
        a1 = fma(__ldg(o + (i + 1) * ldo + r), __ldg(s + (i + 1) * lds + r), a1);
        a2 = fma(__ldg(o + (i + 2) * ldo + r), __ldg(s + (i + 2) * lds + r), a2);
        a3 = fma(__ldg(o + (i + 3) * ldo + r), __ldg(s + (i + 3) * lds + r), a3);
    }
    for (; i < i1; ++i) a0 = fma(__ldg(o + i * ldo + r), __ldg(s + i * lds + r), a0);
    out[(int64_t)blockIdx.y * n + r] = (a0 + a1) + (a2 + a3);
}
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ part, int chunks, int64_t n,
                                                     T* __restrict__ w) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    T acc = T(0);
    for (int c = 0; c < chunks; ++c) acc += part[(int64_t)c * n + r];
    w[r] = acc;
}

template <typename T>
static int dots_t_impl(const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n, void* w,
                       cudaStream_t st) {
    if (n <= 0) return 0;
    if (m <= 0) return (int)cudaMemsetAsync(w, 0, n * sizeof(T), st);
    int64_t bx = (n + 255) / 256;
    int64_t want = ((int64_t)sm_count() * 8 + bx - 1) / bx;      // CTAs along the vector axis
    int64_t chunks = want < 1 ? 1 : want;
    if (chunks > (m + 63) / 64) chunks = (m + 63) / 64;          // at least 64 vectors per chunk
    if (chunks < 1) chunks = 1;
    int64_t ichunk = (m + chunks - 1) / chunks;
    chunks = (m + ichunk - 1) / ichunk;
    T* out = (T*)w;
    if (chunks > 1) {
        void* scratch = nullptr;
        int rc = scratch_acquire((size_t)chunks * n * sizeof(T), &scratch);
        if (rc) return rc;
        out = (T*)scratch;
    }
    dots_t_kernel<T><<<dim3((unsigned)bx, (unsigned)chunks), 256, 0, st>>>((const T*)s, lds, (const T*)o, ldo, m, n, ichunk, out);
    int rc = check_launch();
    if (rc || chunks == 1) return rc;
    colsum_kernel<T><<<(unsigned)bx, 256, 0, st>>>(out, (int)chunks, n, (T*)w);
    return check_launch();
}

// ---- min / max of a block (AMatrix.scale(), dense_matrix.py:32-34) ------------------------
// The reference scans the HOST array with numpy.amin / numpy.amax (two passes over 1.9 GB at
// config 2, ~0.2 s); the block is already on the device, where one pass is HBM-bound.
// Partials per CTA in fixed slots, one CTA folds them (min/max are exact in any order).
template <typename T>
__device__ __forceinline__ void minmax_block(T& lo, T& hi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
    }
    __shared__ T slo[8], shi[8];
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < 8; ++w) { lo = slo[w] < lo ? slo[w] : lo; hi = shi[w] > hi ? shi[w] : hi; }
}
template <typename T>
__global__ void __launch_bounds__(256) minmax_partial_kernel(const T* __restrict__ x, int64_t ld, int64_t m,
                                                             int64_t n, T* __restrict__ part) {
    T lo = __ldg(x), hi = lo;
    const int64_t per_row = (n + 255) / 256;
    for (int64_t j = blockIdx.y; j < m; j += gridDim.y) {
        const T* row = x + j * ld;
        for (int64_t c = (int64_t)blockIdx.x; c < per_row; c += gridDim.x) {
            const int64_t r = c * 256 + threadIdx.x;
            if (r < n) { const T v = __ldg(row + r); lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
        }
    }
    minmax_block(lo, hi);
    if (threadIdx.x == 0) {
        const int64_t slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        part[2 * slot] = lo; part[2 * slot + 1] = hi;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) minmax_final_kernel(const T* __restrict__ part, int64_t slots, T* __restrict__ out) {
    T lo = part[0], hi = part[1];
    for (int64_t s = threadIdx.x; s < slots; s += 256) {
        const T a = part[2 * s], b = part[2 * s + 1];
        lo = a < lo ? a : lo; hi = b > hi ? b : hi;
    }
    minmax_block(lo, hi);
    if (threadIdx.x == 0) { out[0] = lo; out[1] = hi; }
}

static void minmax_grid(int64_t m, int64_t n, unsigned* gx, unsigned* gy) {
    const int64_t per_row = (n + 255) / 256;
    *gy = (unsigned)(m < 64 ? m : 64);
    int64_t g = ((int64_t)sm_count() * 16 + *gy - 1) / *gy;
    if (g > per_row) g = per_row;
    if (g < 1) g = 1;
    *gx = (unsigned)g;
}

template <typename T>
static int minmax_impl(const void* x, int64_t ld, int64_t m, int64_t n, void* out2, void* ws, cudaStream_t st) {
    unsigned gx, gy;
    minmax_grid(m, n, &gx, &gy);
    minmax_partial_kernel<T><<<dim3(gx, gy), 256, 0, st>>>((const T*)x, ld, m, n, (T*)ws);
    int rc = check_launch();
    if (rc) return rc;
    minmax_final_kernel<T><<<1, 256, 0, st>>>((const T*)ws, (int64_t)gx * gy, (T*)out2);
    return check_launch();
}

}  // namespace rl

using namespace rl;

#define RL_DISPATCH(dtype, ...)                          \
    switch (dtype) {                                     \
        case RL_F32: { using T = float; __VA_ARGS__; }   \
        case RL_F64: { using T = double; __VA_ARGS__; }  \
        default: return RL_E_DTYPE;                      \
    }

extern "C" {

int rl_copy(int dtype, void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t m, int64_t n,
            void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    if (m == 0 || n == 0 || dst == src) return 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    // the copy engine is not faster than SMs for D2D on B200, and a kernel keeps
    // stream ordering + launch accounting uniform
    Span span(PK_COPY, as_stream(stream), 2.0 * m * n * w, 0.0);
    RL_DISPATCH(dtype, {
        CopyOp<T> op{(T*)dst, (const T*)src, ld_dst, ld_src};
        return launch_ew<T>(op, m, n, vec_ok<T>(dst, ld_dst, src, ld_src), as_stream(stream));
    })
}

int rl_gather(int dtype, void* dst, int64_t ld_dst, const void* src_all, int64_t ld_src, const int64_t* ind_h,
              int64_t count, int64_t n, void* stream) {
    if (count < 0 || n < 0 || (count > 0 && !ind_h)) return RL_E_ARG;
    if (count == 0 || n == 0) return 0;
    Span span(PK_GATHER, as_stream(stream), 2.0 * count * n * (dtype == RL_F32 ? 4 : 8), 0.0);
    RL_DISPATCH(dtype, {
        for (int64_t t0 = 0; t0 < count; t0 += GATHER_MAX) {
            int64_t c = count - t0 < GATHER_MAX ? count - t0 : GATHER_MAX;
            GatherOp<T> op;
            op.y = (T*)dst + t0 * ld_dst; op.x = (const T*)src_all; op.ldy = ld_dst; op.ldx = ld_src;
            for (int64_t t = 0; t < c; ++t) { if (ind_h[t0 + t] < 0) return RL_E_ARG; op.idx.v[t] = ind_h[t0 + t]; }
            int rc = launch_ew<T>(op, c, n, vec_ok<T>(dst, ld_dst, src_all, ld_src), as_stream(stream));
            if (rc) return rc;
        }
        return 0;
    })
}

int rl_fill_uniform(int dtype, void* x, int64_t ld, int64_t m, int64_t n, uint64_t seed, int64_t j0, int64_t r0,
                    void* stream) {
    if (m < 0 || n < 0 || r0 < 0 || j0 < 0) return RL_E_ARG;
    if (m == 0 || n == 0) return 0;
    Span span(PK_FILL, as_stream(stream), 1.0 * m * n * (dtype == RL_F32 ? 4 : 8), 0.0);
    RL_DISPATCH(dtype, {
        int64_t groups = n / (sizeof(T) == 4 ? 4 : 2) + 2;
        int64_t gx = (groups + 255) / 256;
        if (gx > 4096) gx = 4096;
        dim3 g((unsigned)gx, (unsigned)(m < 65535 ? m : 65535));
        fill_uniform_kernel<T><<<g, 256, 0, as_stream(stream)>>>((T*)x, ld, m, n, seed, j0, r0);
        return check_launch();
    })
}

int rl_axpy(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t m, int64_t n, double alpha,
            void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_AXPY, as_stream(stream), 3.0 * m * n * (dtype == RL_F32 ? 4 : 8), 2.0 * m * n);
    RL_DISPATCH(dtype, {
        AxpyOp<T> op{(T*)y, (const T*)x, ldy, ldx, (T)alpha};
        return launch_ew<T>(op, m, n, vec_ok<T>(y, ldy, x, ldx), as_stream(stream));
    })
}

int rl_axpy_diag(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t m, int64_t n, const void* s,
                 void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_AXPY_DIAG, as_stream(stream), 3.0 * m * n * (dtype == RL_F32 ? 4 : 8), 2.0 * m * n);
    RL_DISPATCH(dtype, {
        AxpyDiagOp<T> op{(T*)y, (const T*)x, ldy, ldx, (const T*)s};
        return launch_ew<T>(op, m, n, vec_ok<T>(y, ldy, x, ldx), as_stream(stream));
    })
}

int rl_scale(int dtype, void* y, int64_t ldy, int64_t m, int64_t n, const void* s, int multiply, void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_SCALE, as_stream(stream), 2.0 * m * n * (dtype == RL_F32 ? 4 : 8), 1.0 * m * n);
    RL_DISPATCH(dtype, {
        ScaleOp<T> op{(T*)y, ldy, (const T*)s, multiply};
        return launch_ew<T>(op, m, n, vec_ok<T>(y, ldy), as_stream(stream));
    })
}

int rl_diag_mul(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t m, int64_t n, const void* d,
                void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_DIAG_MUL, as_stream(stream), (2.0 * m + 1.0) * n * (dtype == RL_F32 ? 4 : 8), 1.0 * m * n);
    RL_DISPATCH(dtype, {
        DiagMulOp<T> op{(T*)y, (const T*)x, ldy, ldx, (const T*)d};
        return launch_ew<T>(op, m, n, vec_ok<T>(y, ldy, x, ldx) && host_aligned16(d), as_stream(stream));
    })
}

size_t rl_dots_ws_bytes(int dtype, int64_t m, int64_t n) {
    if (m <= 0 || n <= 0) return 0;
    int64_t chunk; int chunks;
    dots_plan(m, n, dtype == RL_F32 ? 4 : 2, &chunk, &chunks);
    return chunks > 1 ? (size_t)m * chunks * 8 : 0;
}

int rl_dots(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n, void* w,
            void* ws, size_t ws_bytes, void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_DOTS, as_stream(stream), (s == o ? 1.0 : 2.0) * m * n * (dtype == RL_F32 ? 4 : 8), 2.0 * m * n);
    RL_DISPATCH(dtype, { return dots_impl<T>(s, lds, o, ldo, m, n, w, ws, ws_bytes, as_stream(stream)); })
}

int rl_dots_t(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n, void* w,
              void* stream) {
    if (m < 0 || n < 0) return RL_E_ARG;
    Span span(PK_DOTS_T, as_stream(stream), 2.0 * m * n * (dtype == RL_F32 ? 4 : 8), 2.0 * m * n);
    RL_DISPATCH(dtype, { return dots_t_impl<T>(s, lds, o, ldo, m, n, w, as_stream(stream)); })
}

// ---- host-array conveniences ---------------------------------------------------
static int stage_in(const void* src_h, size_t bytes, void** dev, cudaStream_t st) {
    void* pinned = nullptr;
    int rc = staging_acquire(bytes, &pinned, dev);
    if (rc) return rc;
    memcpy(pinned, src_h, bytes);
    return (int)cudaMemcpyAsync(*dev, pinned, bytes, cudaMemcpyHostToDevice, st);
}

int rl_axpy_diag_h(int dtype, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t m, int64_t n,
                   const void* s_h, void* stream) {
    if (m <= 0 || n <= 0) return m < 0 || n < 0 ? RL_E_ARG : 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    void* sd = nullptr;
    int rc = stage_in(s_h, (size_t)m * w, &sd, as_stream(stream));
    if (rc) return rc;
    return rl_axpy_diag(dtype, y, ldy, x, ldx, m, n, sd, stream);
}

int rl_scale_h(int dtype, void* y, int64_t ldy, int64_t m, int64_t n, const void* s_h, int multiply,
               void* stream) {
    if (m <= 0 || n <= 0) return m < 0 || n < 0 ? RL_E_ARG : 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    void* sd = nullptr;
    int rc = stage_in(s_h, (size_t)m * w, &sd, as_stream(stream));
    if (rc) return rc;
    return rl_scale(dtype, y, ldy, m, n, sd, multiply, stream);
}

int rl_dots_h(int dtype, const void* s, int64_t lds, const void* o, int64_t ldo, int64_t m, int64_t n, void* w_h,
              void* stream) {
    if (m <= 0) return m < 0 ? RL_E_ARG : 0;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    void *pinned = nullptr, *dev = nullptr, *ws = nullptr;
    int rc = staging_acquire((size_t)m * w, &pinned, &dev);
    if (rc) return rc;
    size_t wsb = rl_dots_ws_bytes(dtype, m, n);
    if (wsb) { rc = scratch_acquire(wsb, &ws); if (rc) return rc; }
    rc = rl_dots(dtype, s, lds, o, ldo, m, n, dev, ws, wsb, stream);
    if (rc) return rc;
    RL_CUDA(cudaMemcpyAsync(pinned, dev, (size_t)m * w, cudaMemcpyDeviceToHost, as_stream(stream)));
    RL_CUDA(cudaStreamSynchronize(as_stream(stream)));
    memcpy(w_h, pinned, (size_t)m * w);
    return 0;
}

/* min and max over an (m, n) block into two host scalars of the block's dtype */
int rl_minmax_h(int dtype, const void* x, int64_t ld, int64_t m, int64_t n, void* min_h, void* max_h, void* stream) {
    if (m <= 0 || n <= 0) return RL_E_ARG;
    size_t w = dtype == RL_F32 ? 4 : dtype == RL_F64 ? 8 : 0;
    if (!w) return RL_E_DTYPE;
    unsigned gx, gy;
    minmax_grid(m, n, &gx, &gy);
    void *pinned = nullptr, *dev = nullptr, *ws = nullptr;
    int rc = staging_acquire(2 * w, &pinned, &dev);
    if (rc) return rc;
    rc = scratch_acquire((size_t)gx * gy * 2 * w, &ws);
    if (rc) return rc;
    Span span(PK_COPY, as_stream(stream), (double)m * n * w, 0.0);
    RL_DISPATCH(dtype, { rc = minmax_impl<T>(x, ld, m, n, dev, ws, as_stream(stream)); break; })
    if (rc) return rc;
    RL_CUDA(cudaMemcpyAsync(pinned, dev, 2 * w, cudaMemcpyDeviceToHost, as_stream(stream)));
    RL_CUDA(cudaStreamSynchronize(as_stream(stream)));
    memcpy(min_h, pinned, w);
    memcpy(max_h, (char*)pinned + w, w);
    return 0;
}

}  // extern "C"
