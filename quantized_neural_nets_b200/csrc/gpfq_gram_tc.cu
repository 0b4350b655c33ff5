// tcgen05 split-TF32 Gram formation (placeholder until the tensor-core kernel lands).
#include "gpfq_common.cuh"

namespace gpfq {

size_t gram_tc_scratch_bytes(int d, int m) { return 0; }

int gram_tc_form(const float* X, const float* Xq, int64_t ldx, int d, int m, double* GT, double* H, double* A,
                 int64_t ldg, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    GPFQ_REQUIRE(false, "GPFQ_SOLVER_GRAM (tcgen05) is not built yet; use GPFQ_SOLVER_GRAM_F64");
}

}  // namespace gpfq
