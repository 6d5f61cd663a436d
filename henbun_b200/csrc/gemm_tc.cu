// tcgen05 3xTF32 GEMM engine (placeholder until the tensor-core path is validated on hardware).
#include "gemm.cuh"
namespace hb {
bool gemm_tc_eligible(const GemmParams&) { return false; }
int gemm_tc(const GemmParams&, cudaStream_t) { return HB_ERR_ARG; }
}  // namespace hb
