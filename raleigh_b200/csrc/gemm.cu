// rl_dense_apply: dispatch of the dense operator application (Matrix.apply,
// dense_cublas.py:732-776) between the tensor-core kernel and the FMA kernel.
#include "common.cuh"

namespace rl {
bool gemm_tc_supported(const void* a, int64_t lda, const void* x, int64_t ldx);
int gemm_tc(const float* a_hi, const float* a_lo, int64_t lda, int64_t M, int64_t N, const float* x, int64_t ldx,
            float* y, int64_t ldy, int64_t k, int transp, double alpha, double beta, void* ws, size_t ws_bytes,
            cudaStream_t st);
size_t gemm_tc_ws_bytes(int64_t M, int64_t N, int64_t k, int transp);
int split_tf32(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
               cudaStream_t st);
template <typename T>
int gemm_simt(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
              int64_t k, int transp, double alpha, double beta, cudaStream_t st);
template <typename T>
int gemm_skinny(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
                int64_t k, int transp, double alpha, double beta, cudaStream_t st);
bool gemm_dmma_supported(const void* a, int64_t lda, const void* x, int64_t ldx, const void* y, int64_t ldy);
int gemm_dmma(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
              int64_t k, int transp, double alpha, double beta, cudaStream_t st);
}

using namespace rl;

extern "C" {

int rl_dense_apply(int dtype, const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y,
                   int64_t ldy, int64_t k, int transp, double alpha, double beta, void* stream) {
    if (M < 0 || N < 0 || k < 0) return RL_E_ARG;
    if (k == 0 || (transp ? N : M) == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const double w = dtype == RL_F32 ? 4.0 : 8.0;
    Span span(PK_DENSE_APPLY, st, (1.0 * M * N + 1.0 * k * (M + N)) * w, 2.0 * M * N * k);
    // a handful of vectors: HBM-bound matrix-vector sweep (gemm_skinny.cu); knob GEMM_SKINNY = -1 keeps the tiled kernel
    const bool skinny = k <= 8 && M * N >= 4096 && g_knob[KNOB_GEMM_SKINNY] >= 0 && y != x;
    if (dtype == RL_F32)
        return skinny ? gemm_skinny<float>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st)
                      : gemm_simt<float>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
    if (dtype == RL_F64) {
        if (skinny) return gemm_skinny<double>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
        // FP64 tensor pipe (gemm_dmma.cu) whenever the operands allow 16-byte accesses; knob GEMM_DMMA = -1: FMA pipe
        if (g_knob[KNOB_GEMM_DMMA] >= 0 && y != x && gemm_dmma_supported(a, lda, x, ldx, y, ldy))
            return gemm_dmma(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
        return gemm_simt<double>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
    }
    return RL_E_DTYPE;
}

int rl_dense_apply_tc_supported(const void* a, int64_t lda, const void* x, int64_t ldx) {
    return gemm_tc_supported(a, lda, x, ldx) ? 1 : 0;
}

size_t rl_dense_apply_tc_ws_bytes(int64_t M, int64_t N, int64_t k, int transp) {
    if (M <= 0 || N <= 0 || k <= 0) return 0;
    return gemm_tc_ws_bytes(M, N, k, transp);
}

int rl_split_tf32(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                  void* stream) {
    if (rows < 0 || cols < 0 || (ld_src % 4) || (ld_dst % 4)) return RL_E_ARG;
    return split_tf32((const float*)src, ld_src, (float*)dst, ld_dst, rows, cols, as_stream(stream));
}

int rl_dense_apply_tc(const void* a, const void* a_lo, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx,
                      void* y, int64_t ldy, int64_t k, int transp, double alpha, double beta, void* ws,
                      size_t ws_bytes, void* stream) {
    if (M < 0 || N < 0 || k < 0) return RL_E_ARG;
    if (k == 0 || (transp ? N : M) == 0) return 0;
    if (!gemm_tc_supported(a, lda, x, ldx) || !host_aligned16(a_lo)) return RL_E_ARG;
    cudaStream_t st = as_stream(stream);
    // algorithmic traffic: the data matrix once, the blocks in and out
    Span span(PK_DENSE_APPLY_TC, st, (1.0 * M * N + 1.0 * k * (M + N)) * 4.0, 2.0 * M * N * k);
    return gemm_tc((const float*)a, (const float*)a_lo, lda, M, N, (const float*)x, ldx, (float*)y, ldy, k, transp,
                   alpha, beta, ws, ws_bytes, st);
}

}  // extern "C"
