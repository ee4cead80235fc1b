// q100_gemm_tc.cu -- K3b: tcgen05 dequant-GEMM for prefill shapes (placeholder until bring-up).
#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

bool gemm_tc_supported(int64_t, int64_t, int64_t, int, int) { return false; }

int gemm_tc(const void*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int, int,
            int, int, cudaStream_t) {
  set_error("tcgen05 dequant-GEMM not built yet");
  return Q100_EUNSUPPORTED;
}

}  // namespace q100
