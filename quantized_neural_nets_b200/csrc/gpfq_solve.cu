// gpfq_solve_f32 / gpfq_workspace_bytes: argument validation and solver dispatch.
#include "gpfq_common.cuh"

namespace gpfq {
size_t direct_workspace_bytes(int n_rows, int d, int m);
int direct_solve(const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int d, int m, int n_rows,
                 const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q, int64_t ldq,
                 int8_t* levels, double* row_err2, float* U_out, int64_t ldu, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream);
size_t gram_workspace_bytes(int solver, int n_rows, int d, int m);
size_t gram_matrices_workspace_bytes(int solver, int d, int m);
int gram_solve(int solver, const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int d, int m,
               int n_rows, const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q,
               int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2, void* workspace, size_t workspace_bytes,
               cudaStream_t stream);
int gram_path(const float* W, int64_t ldw, int d, int n_rows, const double* GT, const double* H, const double* A,
              int64_t ldg, const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q,
              int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2, cudaStream_t stream);
int gram_matrices(int solver, const float* X, const float* Xq, int64_t ldx, int d, int m, double* GT, double* H,
                  double* A, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace gpfq

using namespace gpfq;

// K = 2^(bits-1) (quantize_neural_net.py:87-88); the reference has no bound.  The kernels carry K as an fp32 value
// and the level as an int, so any K that fp32 counts exactly works; 2^15 (16-bit alphabets) is what is tested.
constexpr int kMaxK = 32768;

extern "C" {

size_t gpfq_workspace_bytes(int32_t solver, int32_t n_rows, int32_t d, int32_t m) {
    if (n_rows <= 0 || d <= 0 || m <= 0) return 256;
    if (solver == GPFQ_SOLVER_DIRECT) return direct_workspace_bytes(n_rows, d, m);
    if (solver == GPFQ_SOLVER_GRAM || solver == GPFQ_SOLVER_GRAM_F64) return gram_workspace_bytes(solver, n_rows, d, m);
    return 0;
}

size_t gpfq_gram_workspace_bytes(int32_t solver, int32_t d, int32_t m) {
    if (d <= 0 || m <= 0 || (solver != GPFQ_SOLVER_GRAM && solver != GPFQ_SOLVER_GRAM_F64)) return 0;
    return gram_matrices_workspace_bytes(solver, d, m);
}

int gpfq_gram_f32(int32_t solver, const float* X, const float* Xq, int64_t ldx, int32_t d, int32_t m, double* GT,
                  double* H, double* A, void* workspace, size_t workspace_bytes, void* stream) {
    GPFQ_REQUIRE(solver == GPFQ_SOLVER_GRAM || solver == GPFQ_SOLVER_GRAM_F64, "gpfq_gram_f32: solver must be a Gram variant");
    GPFQ_REQUIRE(d > 0 && m > 0 && ldx >= m && (ldx % 4) == 0, "gpfq_gram_f32: bad shape");
    GPFQ_REQUIRE(X && Xq && GT && H && A && workspace, "gpfq_gram_f32: null pointer");
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0, "gpfq_gram_f32: workspace must be 256-byte aligned");
    return gram_matrices(solver, X, Xq, ldx, d, m, GT, H, A, workspace, workspace_bytes, (cudaStream_t)stream);
}

int gpfq_gram_path_f32(const float* W, int64_t ldw, int32_t N, int32_t d, int32_t n0, int32_t n1, const double* GT,
                       const double* H, const double* A, int64_t ldg, const float* delta, int32_t K, int32_t mode, float lam,
                       uint64_t seed, float* Q, int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2,
                       void* stream) {
    GPFQ_REQUIRE(N >= 0 && d > 0 && 0 <= n0 && n0 <= n1 && n1 <= N, "gpfq_gram_path_f32: bad shape or neuron range");
    GPFQ_REQUIRE(mode >= GPFQ_MODE_MSQ && mode <= GPFQ_MODE_STOCHASTIC && K >= 1 && K <= kMaxK, "gpfq_gram_path_f32: bad mode or K");
    GPFQ_REQUIRE(levels == nullptr || K + (mode == GPFQ_MODE_HARD ? 1 : 0) <= 127,
                 "gpfq_gram_path_f32: the int8 `levels` output needs K <= 127 (126 for the L0 alphabet); got K=%d", K);
    GPFQ_REQUIRE(ldw >= d && ldq >= d && ldg >= (d + 31) / 32 * 32, "gpfq_gram_path_f32: leading dimension too small");
    if (n0 == n1) return 0;
    GPFQ_REQUIRE(W && GT && H && A && delta && Q, "gpfq_gram_path_f32: null pointer");
    return gram_path(W + (int64_t)n0 * ldw, ldw, d, n1 - n0, GT, H, A, ldg, delta, K, mode, lam, seed, n0,
                     Q + (int64_t)n0 * ldq, ldq, levels ? levels + (int64_t)n0 * d : nullptr, row_err2, row_ref2,
                     (cudaStream_t)stream);
}

int gpfq_solve_f32(int32_t solver, const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int32_t N,
                   int32_t d, int32_t m, int32_t n0, int32_t n1, const float* delta, int32_t K, int32_t mode, float lam,
                   uint64_t seed, float* Q, int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2, float* U_out, int64_t ldu,
                   void* workspace, size_t workspace_bytes, void* stream) {
    GPFQ_REQUIRE(N >= 0 && d >= 0 && m >= 0, "gpfq_solve_f32: negative dimension");
    GPFQ_REQUIRE(0 <= n0 && n0 <= n1 && n1 <= N, "gpfq_solve_f32: bad neuron range [%d, %d) of %d", n0, n1, N);
    GPFQ_REQUIRE(mode >= GPFQ_MODE_MSQ && mode <= GPFQ_MODE_STOCHASTIC, "gpfq_solve_f32: bad mode %d", mode);
    GPFQ_REQUIRE(K >= 1 && K <= kMaxK, "gpfq_solve_f32: boundary index K=%d outside [1,%d]", K, kMaxK);
    GPFQ_REQUIRE(levels == nullptr || K + (mode == GPFQ_MODE_HARD ? 1 : 0) <= 127,
                 "gpfq_solve_f32: the int8 `levels` output needs K <= 127 (126 for the L0 alphabet); got K=%d", K);
    GPFQ_REQUIRE(ldw >= d && ldq >= d, "gpfq_solve_f32: ldw/ldq smaller than d");
    GPFQ_REQUIRE(ldx >= m && (ldx % 4) == 0, "gpfq_solve_f32: ldx must be >= m and a multiple of 4");
    GPFQ_REQUIRE(U_out == nullptr || ldu >= m, "gpfq_solve_f32: ldu smaller than m");
    const int n_rows = n1 - n0;
    if (n_rows == 0 || d == 0) return 0;
    GPFQ_REQUIRE(m > 0, "gpfq_solve_f32: no calibration rows");
    GPFQ_REQUIRE(W && X && Xq && delta && Q && workspace, "gpfq_solve_f32: null pointer");
    const float* Ws = W + (int64_t)n0 * ldw;
    float* Qs = Q + (int64_t)n0 * ldq;
    int8_t* Ls = levels ? levels + (int64_t)n0 * d : nullptr;
    if (solver == GPFQ_SOLVER_GRAM || solver == GPFQ_SOLVER_GRAM_F64) {
        GPFQ_REQUIRE(U_out == nullptr, "gpfq_solve_f32: the Gram solvers do not materialise U (U_out must be NULL)");
        return gram_solve(solver, Ws, ldw, X, Xq, ldx, d, m, n_rows, delta, K, mode, lam, seed, n0, Qs, ldq, Ls, row_err2,
                          row_ref2, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    GPFQ_REQUIRE(row_ref2 == nullptr || solver != GPFQ_SOLVER_DIRECT,
                 "gpfq_solve_f32: row_ref2 is produced by the Gram solvers only");
    if (solver == GPFQ_SOLVER_DIRECT)
        return direct_solve(Ws, ldw, X, Xq, ldx, d, m, n_rows, delta, K, mode, lam, seed, n0, Qs, ldq, Ls, row_err2, U_out, ldu,
                            workspace, workspace_bytes, (cudaStream_t)stream);
    GPFQ_REQUIRE(false, "gpfq_solve_f32: unknown solver %d", solver);
}

}  // extern "C"
