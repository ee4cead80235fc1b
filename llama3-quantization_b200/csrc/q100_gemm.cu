// q100_gemm.cu -- C-ABI front of the fused dequant-GEMM (K3): argument checks, path choice
// (weight-streaming mma.sync kernel for decode shapes, tcgen05 kernel for prefill shapes),
// and the host-side TMA descriptor encoder.
#include <cstdlib>
#include <mutex>

#include <cudaTypedefs.h>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

int gemm_skinny(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int exact_weights, int pdl, int x_gate_up, cudaStream_t st,
                const void* x_minmax = nullptr, int a_bits = 16, int a_cuda_scale = 0, int no_rows = 0);
size_t skinny_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gemm_stream(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int pdl, cudaStream_t st);
size_t stream_workspace_bytes(int64_t M, int64_t N, int64_t K);
bool gemm_stream_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits);

int gemm_tc(const void* x, const void* packed, const void* meta, const void* bias, void* y,
            int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
            cudaStream_t st);
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits);
int gemm_tc_dense(const void* x, const void* w, void* y, int64_t M, int64_t N, int64_t K, int dtype,
                  cudaStream_t st);
int gemm_tca(const void* x, const void* packed, const void* meta, const void* bias, void* y,
             int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
             int cluster, cudaStream_t st, void* const* peer_stage = nullptr, int rows_per = 0, int rank = 0,
             int world = 1, void* workspace = nullptr, size_t workspace_bytes = 0);
bool gemm_tca_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits);

// ---------------------------------------------------------------- tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  std::call_once(g_encode_once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    else
      cudaGetLastError();
  });
  Q100_REQUIRE(g_encode != nullptr, Q100_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                          : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                            : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                               : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  Q100_REQUIRE(r == CUDA_SUCCESS, Q100_ECUDA,
               "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_pitch_bytes,
               box_inner, box_outer);
  return Q100_OK;
}

}  // namespace q100

using namespace q100;

extern "C" size_t q100_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int n_bits, int flags) {
  (void)n_bits; (void)flags;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int64_t mm = M < 32 ? M : 32;
  const size_t a = skinny_workspace_bytes(mm, N, K);
  const size_t b = stream_workspace_bytes(mm < 16 ? mm : 16, N, K);
  size_t need = a > b ? a : b;
  if (M > 32) {
    // prefill (tcgen05) kernel, stream-K: fp32 partials of the tiles shared by neighbouring clusters, one
    // [128 x 256] slots per SM, behind the first 4 MB (the decode kernels' share); optional -- without it the kernel
    // schedules whole tiles
    const size_t sk = (static_cast<size_t>(1) << 22) + 65536 + static_cast<size_t>(sm_count() + 8) * 2 * 128 * 256 * 4;
    if (sk > need) need = sk;
  }
  return need;
}

extern "C" int q100_gemm_wna16(const void* x, const void* packed, const void* meta, const void* bias,
                               void* y, int64_t M, int64_t N, int64_t K, int group, int n_bits,
                               int disable_zp, int dtype, void* workspace, size_t workspace_bytes,
                               int flags, void* stream) {
  Q100_REQUIRE(x && packed && meta && y, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED,
               "fused dequant-GEMM computes in fp16 or bf16 (got dtype %d)", dtype);
  Q100_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED,
               "n_bits in {2,3,4,8} (got %d)", n_bits);
  Q100_REQUIRE(M >= 0 && N > 0 && K > 0, Q100_EINVAL, "bad shape M=%lld N=%lld K=%lld", (long long)M,
               (long long)N, (long long)K);
  if (M == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool skinny_ok = (K % 128 == 0) && (N % 8 == 0);
  const bool gate_up = (flags & Q100_GEMM_X_GATE_UP) != 0;
  if (gate_up)
    Q100_REQUIRE(M <= 32 && skinny_ok && !(flags & Q100_GEMM_FORCE_TC), Q100_EUNSUPPORTED,
                 "Q100_GEMM_X_GATE_UP is a decode-kernel option (M <= 32, K %% 128 == 0); got M=%lld", (long long)M);
  // prefill: K3b' (weights as the TMEM A operand); K3b (dequantised weights through shared memory) on request
  const bool tca_ok = gemm_tca_supported(M, N, K, group, n_bits) && !(flags & Q100_GEMM_FORCE_TC_SMEM);
  const bool tc_ok = tca_ok || gemm_tc_supported(M, N, K, group, n_bits);
  const bool want_tc = !gate_up && ((flags & Q100_GEMM_FORCE_TC) ||
                       (!(flags & Q100_GEMM_FORCE_SKINNY) && tc_ok && (M > 32 || !skinny_ok)));
  if (want_tc) {
    Q100_REQUIRE(tc_ok, Q100_EUNSUPPORTED, "tcgen05 path does not support this shape/bit-width");
    if (tca_ok) {
      // cluster size for the x multicast: Q100_TCA_CLUSTER in the environment (1, 2 or 4) overrides the default
      static const int env_cluster = []() { const char* e = getenv("Q100_TCA_CLUSTER"); return e ? atoi(e) : 0; }();
      // measured on up_proj, M = 2048 (profiles/r01): 1193 / 1221 / 1236 TFLOP/s for clusters of 1 / 2 / 4
      const int64_t row_tiles = (N + 127) / 128;
      const int cluster = env_cluster > 0 ? env_cluster : (row_tiles % 4 == 0 ? 4 : row_tiles % 2 == 0 ? 2 : 1);
      return gemm_tca(x, packed, meta, bias, y, M, N, K, group, n_bits, disable_zp, dtype, cluster, st, nullptr, 0, 0, 1,
                      workspace, workspace_bytes);
    }
    return gemm_tc(x, packed, meta, bias, y, M, N, K, group, n_bits, disable_zp, dtype, st);
  }
  // decode shapes: tcgen05 weight-streaming kernel (weights as the TMEM A operand)
  // (measured, profiles/r01: it is M-independent -- 107 us at M = 16 vs 161 us for the mma.sync kernel on
  // 57344 x 8192 -- but its small-N MMAs and its per-block latency chain lose to mma.sync up to M = 8)
  const bool want_stream = (flags & Q100_GEMM_FORCE_STREAM_TC) || M > 8;
  if (want_stream && !gate_up && !(flags & (Q100_GEMM_FORCE_MMA_SYNC | Q100_GEMM_EXACT_WEIGHTS)) &&
      gemm_stream_supported(M, N, K, group, n_bits))
    return gemm_stream(x, packed, meta, bias, y, M, N, K, group, n_bits, disable_zp, dtype, workspace,
                       workspace_bytes, (flags & Q100_GEMM_PDL) ? 1 : 0, st);
  // mma.sync weight-streaming kernel, 32 rows of x per pass
  const size_t esz = 2;
  for (int64_t m0 = 0; m0 < M; m0 += 32) {
    const int64_t mc = (M - m0) < 32 ? (M - m0) : 32;
    rc = gemm_skinny(reinterpret_cast<const uint8_t*>(x) + m0 * K * esz * (gate_up ? 2 : 1), packed, meta, bias,
                     reinterpret_cast<uint8_t*>(y) + m0 * N * esz, mc, N, K, group, n_bits, disable_zp,
                     dtype, workspace, workspace_bytes,
                     (flags & Q100_GEMM_EXACT_WEIGHTS) ? 1 : (flags & Q100_GEMM_NO_IMMA) ? 2 : 0,
                     (flags & Q100_GEMM_PDL) ? 1 : 0, gate_up ? 1 : 0, st, nullptr, 16, 0,
                     (flags & Q100_GEMM_NO_ROWS) ? 1 : 0);
    if (rc != Q100_OK) return rc;
  }
  return Q100_OK;
}

namespace q100 {
bool gemm_rows_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits, int mode);
int gemm_rows(const void* x, const void* packed, const void* meta, const void* bias, void* y, int64_t M, int64_t N,
              int64_t K, int disable_zp, int dtype, int mode, int pdl, int x_gate_up, cudaStream_t st,
              const void* norm_w, float eps, const void* res);
}

extern "C" int q100_gemm_block_fusable(int64_t M, int64_t N, int64_t K, int group, int n_bits) {
  return (M > 0 && N > 0 && K > 0 && q100::gemm_rows_supported(M, N, K, group == 0 ? static_cast<int>(K) : group, n_bits, 2)) ? 1 : 0;
}

extern "C" int q100_gemm_wna16_block(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                                     int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                                     const void* norm_weight, float eps, const void* residual, int flags, void* stream) {
  Q100_REQUIRE(x && packed && meta && y, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  const bool gate_up = (flags & Q100_GEMM_X_GATE_UP) != 0;
  Q100_REQUIRE(!(gate_up && norm_weight), Q100_EINVAL, "Q100_GEMM_X_GATE_UP and norm_weight exclude each other");
  Q100_REQUIRE(q100_gemm_block_fusable(M, N, K, group, n_bits), Q100_EUNSUPPORTED,
               "q100_gemm_wna16_block: M=%lld N=%lld K=%lld group=%d W%d is not a row-owner (K3r) case: see "
               "q100_gemm_block_fusable", (long long)M, (long long)N, (long long)K, group, n_bits);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  return gemm_rows(x, packed, meta, bias, y, M, N, K, disable_zp, dtype, 2, (flags & Q100_GEMM_PDL) ? 1 : 0,
                   gate_up ? 1 : 0, reinterpret_cast<cudaStream_t>(stream), norm_weight, eps, residual);
}

extern "C" int q100_gemm_wna8(const void* x, const void* row_minmax, const void* packed, const void* meta,
                              const void* bias, void* y, int64_t M, int64_t N, int64_t K, int group, int n_bits,
                              int a_bits, int disable_zp, int dtype, void* workspace, size_t workspace_bytes,
                              int flags, void* stream) {
  Q100_REQUIRE(x && row_minmax && packed && meta && y, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED, "integer W x A path: W4 or W8 (got W%d)", n_bits);
  Q100_REQUIRE(M >= 0 && N > 0 && K > 0 && K % 128 == 0 && N % 8 == 0, Q100_EINVAL, "bad shape M=%lld N=%lld K=%lld",
               (long long)M, (long long)N, (long long)K);
  Q100_REQUIRE(M <= 15, Q100_EUNSUPPORTED,
               "q100_gemm_wna8 covers decode shapes (M <= 15); prefill runs q100_fakequant_rows + q100_gemm_wna16");
  if (M == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  return gemm_skinny(x, packed, meta, bias, y, M, N, K, group, n_bits, disable_zp, dtype, workspace, workspace_bytes, 0,
                     (flags & Q100_GEMM_PDL) ? 1 : 0, 0, reinterpret_cast<cudaStream_t>(stream), row_minmax, a_bits,
                     (flags & Q100_FLAG_CUDA_SCALE) ? 1 : 0);
}

extern "C" int q100_gemm_wna16_scatter(const void* x, const void* packed, const void* meta, int64_t M, int64_t N,
                                       int64_t K, int group, int n_bits, int disable_zp, int dtype,
                                       const void* peer_stage, int rank, int world, void* stream) {
  Q100_REQUIRE(x && packed && meta && peer_stage, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, Q100_EINVAL, "bad rank %d / world %d", rank, world);
  Q100_REQUIRE(M > 0 && M % (static_cast<int64_t>(world) * 256) == 0, Q100_EUNSUPPORTED,
               "scatter epilogue: M=%lld must be a multiple of world * 256 token rows", (long long)M);
  Q100_REQUIRE(gemm_tca_supported(M, N, K, group, n_bits), Q100_EUNSUPPORTED, "tcgen05 path does not support this shape");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  const int64_t row_tiles = (N + 127) / 128;
  const int cluster = row_tiles % 4 == 0 ? 4 : row_tiles % 2 == 0 ? 2 : 1;
  return gemm_tca(x, packed, meta, nullptr, nullptr, M, N, K, group, n_bits, disable_zp, dtype, cluster,
                  reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<void* const*>(peer_stage),
                  static_cast<int>(M / world), rank, world);
}

#ifdef TA_TIMING
namespace q100 { int tca_timing(long long* out, int n); }
extern "C" __attribute__((visibility("default"))) int q100_debug_tca_timing(long long* out, int n) {
  return q100::tca_timing(out, n);
}
#endif

#ifdef SK_TIMING
namespace q100 { int skinny_timing(long long* out, int n); }
extern "C" __attribute__((visibility("default"))) int q100_debug_skinny_timing(long long* out, int n) {
  return q100::skinny_timing(out, n);
}
#endif

extern "C" int q100_debug_gemm_dense_tc(const void* x, const void* w, void* y, int64_t M, int64_t N, int64_t K,
                                        int dtype, void* stream) {
  Q100_REQUIRE(x && w && y, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  return gemm_tc_dense(x, w, y, M, N, K, dtype, reinterpret_cast<cudaStream_t>(stream));
}
