// q100_gemm_tc.cu -- K3b: tcgen05 dequant-GEMM for prefill shapes (M > 32).
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// Tensor-core bound.  One CTA computes a [128 x BN] output tile with the accumulator in TMEM:
//   warp 8      TMA producer: x tile [128 x 64] (SWIZZLE_128B), packed codes [BN x 64 k] and
//               the (scale, zero) side band into a STAGES-deep shared-memory ring;
//   warps 0-7   dequantise the packed tile in registers (r(r(q - zp) * scale), bit-identical
//               to the reference's fake-quantised weight) and write it as the K-major
//               SWIZZLE_128B B operand; fence.proxy.async; arrive on the stage's mbarrier;
//               after the main loop the same warps run the epilogue (tcgen05.ld -> bf16/fp16
//               -> global);
//   warp 9      allocates TMEM, issues tcgen05.mma (kind::f16, M=128, N=BN, K=16) from one
//               elected thread, tcgen05.commit frees the stage / publishes the accumulator.
// The DENSE instantiation (B tile loaded by TMA from a 16-bit weight) exists to validate the
// TMA / descriptor / TMEM plumbing independently of the dequant stage.
#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_DQ_THREADS = 256;
constexpr int TC_THREADS = TC_DQ_THREADS + 64;

struct TcArgs {
  const void* meta;     // [K/group, N, 2] T
  const void* bias;     // [N] T or null
  void* y;              // [M, N] T
  int M, N, K, group;
  int has_zp;
  float qmin;
};

template <int BN, int NBITS, bool DENSE>
struct TcSmem {
  static constexpr int kA = TC_BM * TC_BK * 2;                       // 16 KB
  static constexpr int kB = BN * TC_BK * 2;                          // dequantised / dense B tile
  static constexpr int kP = DENSE ? 0 : BN * TC_BK * NBITS / 8;      // packed codes
  static constexpr int kMeta = DENSE ? 0 : ((BN * 4 + 1023) & ~1023);
  static constexpr int kStage = kA + kB + kP + kMeta;                // multiples of 1024
  static constexpr int kStages = (kStage * 4 + 2048 <= 227 * 1024) ? 4 : 3;
  static constexpr int kTotal = kStages * kStage + 1024;
};

template <typename T> struct DQT;
template <> struct DQT<__half> {
  using T2 = __half2;
  static constexpr uint32_t kMagic = 0x64006400u;
  __device__ static __forceinline__ T2 magic_val() { return __float2half2_rn(1024.0f); }
};
template <> struct DQT<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr uint32_t kMagic = 0x43004300u;
  __device__ static __forceinline__ T2 magic_val() { return __float2bfloat162_rn(128.0f); }
};

template <typename T>
__device__ __forceinline__ uint32_t tc_dq_pair(uint32_t pair, typename DQT<T>::T2 qmin2,
                                               typename DQT<T>::T2 zp2, typename DQT<T>::T2 s2, bool has_zp) {
  using T2 = typename DQT<T>::T2;
  const uint32_t m = pair | DQT<T>::kMagic;
  T2 c = __hsub2(*reinterpret_cast<const T2*>(&m), DQT<T>::magic_val());
  if (has_zp) c = __hsub2(c, zp2);
  else c = __hadd2(c, qmin2);
  c = __hmul2(c, s2);
  return *reinterpret_cast<uint32_t*>(&c);
}

template <typename T, int BN, int NBITS, bool DENSE>
__global__ void __launch_bounds__(TC_THREADS, 1)
k3_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const TcArgs a) {
  using SM = TcSmem<BN, NBITS, DENSE>;
  using T2 = typename DQT<T>::T2;
  constexpr int STAGES = SM::kStages;
  constexpr bool kBf16 = DT<T>::kCode == Q100_BF16;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u32 - smem_u32(smem_raw));
  __shared__ uint64_t full_tma[STAGES], full_b[STAGES], empty_bar[STAGES], tmem_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * TC_BM;
  const int num_kb = a.K / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_tma[s], 1);
      mbar_init(&full_b[s], TC_DQ_THREADS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&tmem_base_s, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_s;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      tma_prefetch_desc(&tm_w);
      const uint64_t pol_w = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = base + s * SM::kStage;
        uint32_t bytes = SM::kA + (DENSE ? SM::kB : SM::kP + BN * 4);
        mbar_expect_tx(&full_tma[s], bytes);
        tma_load_2d_hint(st, &tm_x, &full_tma[s], kb * TC_BK, m0, pol_x);
        if (DENSE) {
          tma_load_2d_hint(st + SM::kA, &tm_w, &full_tma[s], kb * TC_BK, n0, pol_w);
        } else {
          tma_load_2d_hint(st + SM::kA + SM::kB, &tm_w, &full_tma[s], kb * (TC_BK * NBITS / 32), n0, pol_w);
          const int kg = (kb * TC_BK) / a.group;
          const T* src = reinterpret_cast<const T*>(a.meta) + (static_cast<int64_t>(kg) * a.N + n0) * 2;
          bulk_load_1d(st + SM::kA + SM::kB + SM::kP, src, BN * 4, &full_tma[s]);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(kBf16 ? 1 : 0, TC_BM, BN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_tma[s], ph);
      if (!DENSE) mbar_wait(&full_b[s], ph);
      tc_fence_after();
      if (lane == 0) {
        uint8_t* st = base + s * SM::kStage;
        const uint64_t da = umma_desc_sw128(st);
        const uint64_t db = umma_desc_sw128(st + SM::kA);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // +32 bytes per K=16 step inside the 128-byte swizzled row (start-address field is >> 4)
          umma_f16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                 // frees the stage when these MMAs retire
        if (kb == num_kb - 1) umma_commit(&tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== dequant warps (0..7) =====================
    if (!DENSE) {
      const uint32_t qb = DT<T>::pack2(a.qmin, a.qmin);
      const T2 qmin2 = *reinterpret_cast<const T2*>(&qb);
      const bool has_zp = a.has_zp != 0;
      constexpr int CODES_PER_CHUNK = 128 / NBITS;                 // codes in one 16-byte chunk
      constexpr int CHUNKS_PER_ROW = TC_BK / CODES_PER_CHUNK;      // W4: 2
      constexpr int N_CHUNKS = BN * CHUNKS_PER_ROW;
      static_assert(NBITS == 4, "tcgen05 dequant path: W4 only for now");
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_tma[s], ph);
        const uint32_t st = base_u32 + s * SM::kStage;
        const uint32_t bt = st + SM::kA;
        const uint32_t pt = st + SM::kA + SM::kB;
        const uint32_t mt = pt + SM::kP;
        for (int c = tid; c < N_CHUNKS; c += TC_DQ_THREADS) {
          const int row = c / CHUNKS_PER_ROW, hf = c % CHUNKS_PER_ROW;
          uint4 w;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(pt + c * 16));
          uint32_t m;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(mt + row * 4));
          const uint32_t sb = __byte_perm(m, 0, 0x1010), zb = __byte_perm(m, 0, 0x3232);
          const T2 s2 = *reinterpret_cast<const T2*>(&sb), z2 = *reinterpret_cast<const T2*>(&zb);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {        // word j: 8 consecutive k -> one 16-byte chunk of the B row
            uint4 o;
            o.x = tc_dq_pair<T>(ww[j] & 0x000f000fu, qmin2, z2, s2, has_zp);
            o.y = tc_dq_pair<T>((ww[j] >> 4) & 0x000f000fu, qmin2, z2, s2, has_zp);
            o.z = tc_dq_pair<T>((ww[j] >> 8) & 0x000f000fu, qmin2, z2, s2, has_zp);
            o.w = tc_dq_pair<T>((ww[j] >> 12) & 0x000f000fu, qmin2, z2, s2, has_zp);
            const int chunk = hf * 4 + j;                                  // 16-byte chunk index in the 128-byte row
            const uint32_t dst = bt + row * 128 + ((chunk ^ (row & 7)) << 4);   // SWIZZLE_128B
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};"
                         :: "r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_b[s]);
      }
    }
    // ===================== epilogue =====================
    mbar_wait(&tmem_full, 0);
    tc_fence_after();
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
    const int col_half = warp >> 2;                    // warps 0-3: columns [0, BN/2), 4-7: the rest
    const int m = m0 + quad * 32 + lane;
    T* yrow = reinterpret_cast<T*>(a.y) + static_cast<int64_t>(m) * a.N + n0;
    const T* bias = reinterpret_cast<const T*>(a.bias);
#pragma unroll 1
    for (int c0 = col_half * (BN / 2); c0 < (col_half + 1) * (BN / 2); c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (m < a.M) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
            if (bias) {
              f0 += DT<T>::to_f(bias[n0 + c0 + v * 8 + 2 * e]);
              f1 += DT<T>::to_f(bias[n0 + c0 + v * 8 + 2 * e + 1]);
            }
            o[e] = DT<T>::pack2(f0, f1);
          }
          *reinterpret_cast<uint4*>(yrow + c0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
  }
}

// ---------------------------------------------------------------- host
template <typename T, int BN, int NBITS, bool DENSE>
static int launch_tc(const CUtensorMap& tx, const CUtensorMap& tw, const TcArgs& a, cudaStream_t st) {
  using SM = TcSmem<BN, NBITS, DENSE>;
  auto kern = k3_tc<T, BN, NBITS, DENSE>;
  static bool attr_done = false;
  if (!attr_done) {
    Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_done = true;
  }
  dim3 grid(a.N / BN, (a.M + TC_BM - 1) / TC_BM);
  kern<<<grid, TC_THREADS, SM::kTotal, st>>>(tx, tw, a);
  return check_cuda(cudaGetLastError(), "k3_tc launch");
}

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits) {
  const int64_t g = group == 0 ? K : group;
  return n_bits == 4 && M >= 1 && N % 128 == 0 && K % TC_BK == 0 && g % TC_BK == 0 && K % g == 0;
}

int gemm_tc(const void* x, const void* packed, const void* meta, const void* bias, void* y,
            int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
            cudaStream_t st) {
  Q100_REQUIRE(gemm_tc_supported(M, N, K, group, n_bits), Q100_EUNSUPPORTED, "tcgen05 path: unsupported shape");
  const int g = group == 0 ? static_cast<int>(K) : group;
  const int bn = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, x, 2, K, M, static_cast<uint64_t>(K) * 2, TC_BK, TC_BM, 128);
  if (rc != Q100_OK) return rc;
  const int64_t wpr = K * n_bits / 32;
  rc = make_tmap_2d(&tw, packed, 4, wpr, N, static_cast<uint64_t>(wpr) * 4, TC_BK * n_bits / 32, bn, 0);
  if (rc != Q100_OK) return rc;
  TcArgs a{};
  a.meta = meta; a.bias = bias; a.y = y;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = g;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  if (dtype == Q100_F16) {
    return bn == 256 ? launch_tc<__half, 256, 4, false>(tx, tw, a, st) : launch_tc<__half, 128, 4, false>(tx, tw, a, st);
  }
  return bn == 256 ? launch_tc<__nv_bfloat16, 256, 4, false>(tx, tw, a, st)
                   : launch_tc<__nv_bfloat16, 128, 4, false>(tx, tw, a, st);
}

// dense 16-bit weight through the same TMA/UMMA/TMEM plumbing (bring-up / regression hook)
int gemm_tc_dense(const void* x, const void* w, void* y, int64_t M, int64_t N, int64_t K, int dtype,
                  cudaStream_t st) {
  Q100_REQUIRE(N % 128 == 0 && K % TC_BK == 0 && M >= 1, Q100_EUNSUPPORTED, "dense tc: N %% 128, K %% 64");
  const int bn = (N % 256 == 0) ? 256 : 128;
  CUtensorMap tx, tw;
  int rc = make_tmap_2d(&tx, x, 2, K, M, static_cast<uint64_t>(K) * 2, TC_BK, TC_BM, 128);
  if (rc != Q100_OK) return rc;
  rc = make_tmap_2d(&tw, w, 2, K, N, static_cast<uint64_t>(K) * 2, TC_BK, bn, 128);
  if (rc != Q100_OK) return rc;
  TcArgs a{};
  a.y = y; a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = static_cast<int>(K);
  if (dtype == Q100_F16) {
    return bn == 256 ? launch_tc<__half, 256, 4, true>(tx, tw, a, st) : launch_tc<__half, 128, 4, true>(tx, tw, a, st);
  }
  return bn == 256 ? launch_tc<__nv_bfloat16, 256, 4, true>(tx, tw, a, st)
                   : launch_tc<__nv_bfloat16, 128, 4, true>(tx, tw, a, st);
}

}  // namespace q100
