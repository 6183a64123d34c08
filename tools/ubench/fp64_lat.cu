// Micro-benchmarks behind the Jacobi kernels' cost model: latency / throughput of the FP64 pipe, of the
// double-precision special functions, of a 64-bit warp reduction and of the CTA barrier on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_lat(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
template <int ILP>
__global__ void k_dfma_tp(double* out, long long* cyc, int iters) {
    double a[ILP];
    double b = out[1];
    for (int j = 0; j < ILP; ++j) a[j] = out[0] + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], b, b);
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; ++j) s += a[j];
    out[threadIdx.x + 2] = s;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_rsqrt_lat(double* out, long long* cyc, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = rsqrt(a) + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_div_lat(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = b / a + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_sqrt_lat(double* out, long long* cyc, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = sqrt(a) + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_rcp_approx_lat(double* out, long long* cyc, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_rsqrt_approx_lat(double* out, long long* cyc, int iters) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + 2] = a;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
template <int NV>
__global__ void k_shfl_reduce(double* out, long long* cyc, int iters) {
    double v[NV];
    for (int j = 0; j < NV; ++j) v[j] = out[0] + threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < NV; ++j) s += v[j];
    out[threadIdx.x + 2] = s;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_barrier(double* out, long long* cyc, int iters) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
// the rotation's scalar chain as written in jacobi.cu
__device__ __forceinline__ bool rot(double alpha, double beta, double gamma, double tol2, double& c, double& s, double& t) {
    c = 1.0; s = 0.0; t = 0.0;
    if (!(gamma * gamma > tol2 * alpha * beta)) return false;
    const double delta = beta - alpha;
    const double h = fma(delta, delta, 4.0 * gamma * gamma);
    const double den = fabs(delta) + h * rsqrt(h);
    t = (delta >= 0.0 ? 2.0 : -2.0) * gamma / den;
    c = rsqrt(fma(t, t, 1.0));
    s = c * t;
    return true;
}
__global__ void k_rot_lat(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1], g = out[2];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double c, s, t;
        rot(a, b, g, 1e-30, c, s, t);
        a = a + s; b = b + c; g = g * 0.999 + t * 1e-3;
    }
    long long t1 = clock64();
    out[threadIdx.x + 3] = a + b + g;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
// LDS.128 -> DFMA round trip on a column of 64*NJ doubles, like a Jacobi round without the rotation
__global__ void k_lds(double* out, long long* cyc, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(1.0, 2.0);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double2 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = sm[(w * 128 + lane + 32 * j + (int)acc) & 4095];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc = fma(q[j].x, 1e-30, acc); acc = fma(q[j].y, 1e-30, acc); }
    }
    long long t1 = clock64();
    out[threadIdx.x + 3] = acc;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
    double h[3] = {1.000001, 0.999999, 0.3};
    cudaMemcpy(out, h, 24, cudaMemcpyHostToDevice);
    long long c;
    const int it = 4096;
#define RUN(name, kern, threads, per, ...) kern<<<1, threads, ##__VA_ARGS__>>>(out, cyc, it); cudaDeviceSynchronize(); \
    kern<<<1, threads, ##__VA_ARGS__>>>(out, cyc, it); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
    cudaMemcpy(out, h, 24, cudaMemcpyHostToDevice); \
    printf("{\"bench\": \"%s\", \"threads\": %d, \"cycles_per_op\": %.2f}\n", name, threads, (double)c / it / (per));
    RUN("dfma_dependent_latency", k_dfma_lat, 32, 4)
    for (int th = 128; th <= 1024; th *= 2) {
        RUN("dfma_ilp8_cycles_per_warp_instr_per_thread_slot", k_dfma_tp<8>, th, 8)
    }
    RUN("rsqrt_f64_latency(+1 dadd)", k_rsqrt_lat, 32, 1)
    RUN("sqrt_f64_latency(+1 dadd)", k_sqrt_lat, 32, 1)
    RUN("div_f64_latency(+1 dadd)", k_div_lat, 32, 1)
    RUN("rcp_approx_f64_latency(+1 dadd)", k_rcp_approx_lat, 32, 1)
    RUN("rsqrt_approx_f64_latency(+1 dadd)", k_rsqrt_approx_lat, 32, 1)
    RUN("rotation_chain_latency", k_rot_lat, 32, 1)
    RUN("rotation_chain_16warps", k_rot_lat, 512, 1)
    RUN("shfl_tree_1x_f64", k_shfl_reduce<1>, 32, 1)
    RUN("shfl_tree_3x_f64", k_shfl_reduce<3>, 32, 1)
    RUN("shfl_tree_3x_f64_16warps", k_shfl_reduce<3>, 512, 1)
    RUN("shfl_tree_1x_f64_16warps", k_shfl_reduce<1>, 512, 1)
    RUN("syncthreads_8warps", k_barrier, 256, 1)
    RUN("syncthreads_16warps", k_barrier, 512, 1)
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    RUN("lds128x4_dfma8_1warp", k_lds, 32, 1, 65536)
    RUN("lds128x4_dfma8_16warps", k_lds, 512, 1, 65536)
    return 0;
}
