// rl_dense_apply: dispatch of the dense operator application (Matrix.apply,
// dense_cublas.py:732-776) between the tensor-core kernel and the FMA kernel.
#include "common.cuh"

namespace rl {
template <typename T>
int gemm_simt(const void* a, int64_t lda, int64_t M, int64_t N, const void* x, int64_t ldx, void* y, int64_t ldy,
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
    if (dtype == RL_F32) return gemm_simt<float>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
    if (dtype == RL_F64) return gemm_simt<double>(a, lda, M, N, x, ldx, y, ldy, k, transp, alpha, beta, st);
    return RL_E_DTYPE;
}

}  // extern "C"
