// Shared helpers for the raleigh_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/raleigh_b200.h"

#define RL_SM_COUNT_DEFAULT 148

namespace rl {

extern int64_t g_launches;          // kernels launched by this library (api.cu)
// A/B knobs set through rl_debug_set_knob (api.cu); 0 = library default everywhere:
//  GRAM_TMA: -1 register-fragment kernel, 1 / 2 TMA ring with 1 / 2 CTAs per SM, 3 persistent with dynamic
//            chunks (+4: interleaved stages); GRAM_WAVES: CTAs per SM slot (modes 1, 2) / chunks per CTA (3);
//  GRAM_INTERLEAVE: interleaved row steps in the register-fragment kernel;
//  SPMM_CARVEOUT: shared-memory carve-out in percent; SPMM_WPS: 16 = 127-register budget + 4-entry batches,
//            24 = 80 registers + 2-entry batches; SPMM_PREFETCH: -1 off, bit 0 largest-column lines,
//            bit 1 CSR entries of the run one resident window ahead, bit 2 X lines shifted by that window;
//  (knobs 6, 7 belonged to the r1 band-window SpMM experiment, measured 0.60 vs 0.33 ms on the 128^3 stencil in r2a
//   and removed, profiles/r2a_sweep_spmm_window.jsonl; the numbers were reused:)
//  GRAM_CHUNK_MAJOR: 1 = multi-tile Gram products keep the chunk-major 3-D grid (default: tile-major 1-D grid);
//  COPY_DIRECT: 1 = big pageable host<->device copies go straight to cudaMemcpy (default: pinned staging + threads);
//  GEMM_INSPLIT: in-kernel lo split of the tcgen05 dense apply (gemm_tc.cu), the default since r2
//  EIG_GRID_FLAT: 1 = orders 321..1024 on the flat one-grid-barrier-per-round Jacobi kernel (default: ring kernel);
//  EIG_RING_DEBUG: timing experiments on the ring kernel (bit 0: no rotations, bit 1: no block movement; six sweeps);
//  GEMM_DMMA: -1 = fp64 dense apply / big small-matrix products on the FMA-pipe kernels instead of gemm_dmma.cu;
//  CHOL_NOEST: 1 = (measurements only) pivoted Cholesky without the condition estimates of the drop rule;
//  GEMM_SKINNY: -1 = dense applies with k <= 8 vectors stay on the tiled FMA kernel (gemm_simt.cu);
//  BLOCK_TC: -1 = fp32 Gram / block update of the device-resident driver never on the tensor cores;
//  CHOL_GLOBAL: 1 = pivoted Cholesky always in global memory (the shared-memory split is the default);
//  EIG_LEGACY: 1 = always use the cooperative-grid two-sided Jacobi kernel (small.cu) instead of the cluster kernel
enum Knob { KNOB_GRAM_TMA = 0, KNOB_SPMM_CARVEOUT = 1, KNOB_SPMM_WPS = 2, KNOB_SPMM_PREFETCH = 3, KNOB_GRAM_INTERLEAVE = 4, KNOB_GRAM_WAVES = 5, KNOB_GRAM_CHUNK_MAJOR = 6, KNOB_COPY_DIRECT = 7, KNOB_GEMM_INSPLIT = 8,
            KNOB_EIG_LEGACY = 9, KNOB_CHOL_GLOBAL = 10, KNOB_BLOCK_TC = 11, KNOB_GEMM_SKINNY = 12, KNOB_CHOL_NOEST = 13, KNOB_EIG_GRID_FLAT = 14, KNOB_EIG_RING_DEBUG = 15, KNOB_GEMM_DMMA = 16, KNOB_COUNT = 24 };
extern int g_knob[KNOB_COUNT];
int sm_count();                     // cached cudaDevAttrMultiProcessorCount

inline int check_launch() {
    ++g_launches;
    return (int)cudaGetLastError();
}

#define RL_CUDA(expr)                         \
    do {                                      \
        cudaError_t _e = (expr);              \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T> struct Vec128;          // 16-byte vector of T
template <> struct Vec128<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec128<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ bool aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}
inline bool host_aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// streaming (read-once) 128-bit load: keep it out of L1
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ double ldg_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int32_t ldg_stream(const int32_t* p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- optional per-entry-point device timing (rl_profile_*) -------------------
// A Span brackets the device work of one C-ABI call with two CUDA events on the
// caller's stream when profiling is on (bench.py: roofline.achieved is measured
// live from these); it costs one branch when profiling is off.
enum ProfKind {
    PK_GRAM = 0, PK_UPDATE, PK_AXPY, PK_AXPY_DIAG, PK_SCALE, PK_DOTS, PK_DOTS_T, PK_COPY, PK_GATHER,
    PK_DIAG_MUL, PK_SPMM, PK_DENSE_APPLY, PK_DENSE_APPLY_TC, PK_SYEVJ, PK_FILL, PK_PIV_CHOL, PK_RR_SOLVE, PK_SMALL,
    PK_COUNT
};
extern int g_profile_on;
void prof_begin(int kind, cudaStream_t st, double bytes, double flops, void** token);
void prof_end(void* token, cudaStream_t st);
extern int g_span_depth;            // only the outermost span of a call records (rl_rr_solve contains eigensolver calls)
struct Span {
    void* token = nullptr;
    cudaStream_t st;
    bool counted = false;
    Span(int kind, cudaStream_t s, double bytes, double flops) : st(s) {
        if (g_profile_on) {
            counted = true;
            if (g_span_depth++ == 0) prof_begin(kind, s, bytes, flops, &token);
        }
    }
    ~Span() {
        if (counted) --g_span_depth;
        if (token) prof_end(token, st);
    }
};

// Library-owned scratch: a pinned host ring + a device ring used by the *_h
// entry points to move small coefficient / result arrays (api.cu).
struct Staging {
    void* pinned = nullptr;   // cudaHostAlloc
    void* dev = nullptr;      // cudaMalloc
    size_t bytes = 0;
};
int staging_acquire(size_t bytes, void** pinned, void** dev);   // grows on demand
int scratch_acquire(size_t bytes, void** dev);                  // second device scratch (workspaces)

}  // namespace rl
