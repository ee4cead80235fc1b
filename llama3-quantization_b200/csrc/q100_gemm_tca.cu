// q100_gemm_tca.cu -- K3b': persistent tcgen05 dequant-GEMM for prefill shapes, weights as the A operand
// in TENSOR MEMORY.
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// Why a second prefill kernel: ncu on K3b (q100_gemm_tc.cu; profiles/r01f) shows the tensor pipe 45 % busy
// with the shared-memory pipe as the limiter -- per 64-wide k block it moves 32 KB of x in, 32 KB of x out
// to the tensor core, 16 KB of dequantised weights in and 32 KB of them out again (two M halves), 236 bytes
// per tensor-core cycle against 128 B/clk of shared-memory bandwidth.  Here the dequantised weights never
// touch shared memory: the dequant warps write them straight into tensor memory (tcgen05.st) and the MMA
// reads its A operand from there, so shared memory only carries the x tile (TMA in, tensor core out) and
// the packed codes: 143 bytes per tensor-core cycle at a [128 rows x 256 tokens] tile.
//
// Per CTA (one per SM, persistent over [128 weight rows x NT tokens] output tiles, token tiles fastest so
// that CTAs running at the same time share weight tiles in L2), BK = 64:
//   warp 12     TMA producer of the x ring: [NT tokens x 64 k] boxes, SWIZZLE_128B = the K-major B operand;
//   warp 14     bulk-copy producer of the packed ring: the tile's [128 x 64] code tile (one contiguous run in
//               the tile-major layout) + its (scale, zero) row;
//   warps 0-7   two dequant groups of four warps alternating k blocks; thread = weight row: 32 bytes of
//               codes -> registers -> r(r(q - zp) * scale) in packed 16-bit arithmetic (bit-identical to the
//               reference's fake-quantised weight) -> tcgen05.st into one of four A stages (32 columns);
//   warp 13     allocates TMEM (512 columns: D = NT fp32 columns, A ring = 4 x 32); one elected thread issues
//               4 x tcgen05.mma (M = 128, N = NT, K = 16, A from TMEM, B from shared memory) per k block,
//               tcgen05.commit frees the A stage and the x stage / publishes the accumulator;
//   warps 8-11  epilogue: tcgen05.ld (lane = weight row, column = token) -> (+bias) -> bf16/fp16 -> global.
#include <cstdlib>

#include <mutex>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

constexpr int TA_ROWS = 128;          // weight rows per tile (UMMA M)
constexpr int TA_BK = 64;
constexpr int TA_SA = 4;              // A stages in tensor memory
constexpr int TA_SP = 8;              // packed-code stages
constexpr int TA_DQ_WARPS = 8;
constexpr int TA_EPI_WARP0 = 8;
constexpr int TA_TMAX_WARP = 12, TA_MMA_WARP = 13, TA_TMAP_WARP = 14;
constexpr int TA_THREADS = 15 * 32;
constexpr uint32_t TA_TMEM_COLS = 512;
constexpr uint32_t TA_COL_A = 256;    // A ring starts after the (up to) 256 accumulator columns

struct TcaArgs {
  const void* packed;
  const void* meta;
  const void* bias;
  void* y;
  int M, N, K, group;
  int has_zp;
  float qmin;
  int m_tiles, n_tiles;
  // row-parallel tensor parallelism, reduce-scatter as the epilogue (q100_gemm_wna16_scatter): the token rows are
  // cut into `world` blocks of rows_per; the tile of tokens [m0, m0 + NT) belongs to rank m0 / rows_per and is
  // stored over NVLink into slot `rank` of THAT rank's staging buffer [world][rows_per][N] instead of into y
  void* const* peer_stage;     // device array of `world` pointers (peer-mapped symmetric memory); null = plain y
  int rows_per, rank, world;
  // stream-K for the tiles left over after the last full round (see SegIter): the clusters that share such a tile hand
  // fp32 partials to the one that owns its last k blocks through `ws`
  int stream_k;                // 0: whole tiles only
  float* ws_part;              // [cluster][2][CL][NT columns][128 rows] fp32 partials (a range crosses <= 1 tile boundary)
  uint32_t* ws_flag;           // [cluster][2][CL] arrival flags, zero between launches
};

// The (cluster tile, k range) segments of one cluster.  Whole tiles ct0, ct0 + C, ... in full rounds (the CTAs running
// at the same time then work on neighbouring tiles and share x / weight tiles in L2); with stream_k the tiles left over
// after the last full round (L = n_ct mod C of them) are cut along K instead: their L * num_kb k blocks are dealt out
// evenly, cluster c takes the contiguous unit range [c UL / C, (c + 1) UL / C), split at the tile boundary it may
// cross.  (Cutting ALL the work stream-K style was measured first: 967 instead of 1195 TFLOP/s on up_proj -- the
// round-robin tile order matters more than the balance.)
struct SegIter {
  int ct, n_full_end, step, num_kb;      // full rounds: ct = ct0, ct0 + step, ... < n_full_end
  int u, u_end, tile0;                   // leftover: units [u, u_end) of the tiles tile0, tile0 + 1, ...
  __device__ __forceinline__ SegIter(const TcaArgs& a, int n_ct, int num_kb_, int ct0, int ct_step)
      : ct(ct0), n_full_end(n_ct), step(ct_step), num_kb(num_kb_), u(0), u_end(0), tile0(n_ct) {
    if (a.stream_k) {
      const int rounds = n_ct / ct_step;
      n_full_end = rounds * ct_step;
      tile0 = n_full_end;
      const long long UL = static_cast<long long>(n_ct - n_full_end) * num_kb_;
      u = static_cast<int>(UL * ct0 / ct_step);
      u_end = static_cast<int>(UL * (ct0 + 1) / ct_step);
    }
  }
  __device__ __forceinline__ bool next(int& ct_out, int& kb0, int& kb1) {
    if (ct < n_full_end) { ct_out = ct; kb0 = 0; kb1 = num_kb; ct += step; return true; }
    if (u >= u_end) return false;
    const int t = u / num_kb;
    ct_out = tile0 + t;
    kb0 = u - t * num_kb;
    kb1 = min(num_kb, kb0 + (u_end - u));
    u += kb1 - kb0;
    return true;
  }
};
// cluster that owns leftover unit u when C clusters split UL units as [floor(c UL / C), floor((c + 1) UL / C))
__device__ __forceinline__ int sk_owner(int u, int C, long long UL) {
  return static_cast<int>((static_cast<long long>(u + 1) * C + UL - 1) / UL) - 1;
}

template <int NBITS, int NT>
struct TcaSmem {
  static constexpr int kX = NT * TA_BK * 2;                           // x stage
  static constexpr int kP = TA_ROWS * TA_BK * NBITS / 8;              // packed codes of a stage
  static constexpr int kPStage = kP + 512;
  static constexpr int kEpi = 4 * 2 * 32 * 64;                       // 4 warps x 2 buffers x [32 tokens x 32 rows] 16-bit
  static constexpr int kBudget = 225 * 1024 - kEpi;
  static constexpr int kSX = (kBudget - TA_SP * kPStage) / kX > 6 ? 6 : (kBudget - TA_SP * kPStage) / kX;
  static constexpr int kOffP = kSX * kX;
  static constexpr int kOffE = kOffP + TA_SP * kPStage;
  static constexpr int kTotal = kOffE + kEpi + 1024;
  static_assert(kSX >= 3, "x ring too shallow");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <typename T> struct TaDQ;
template <> struct TaDQ<__half> {
  using T2 = __half2;
  static constexpr uint32_t kMagic = 0x64006400u;
  static constexpr float kMagicF = 1024.0f;
  static constexpr float kFastZp = 1000.0f;            // |zp| below this: magic + zp is exact
  __device__ static __forceinline__ T2 bcast(float v) { return __float2half2_rn(v); }
};
template <> struct TaDQ<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr uint32_t kMagic = 0x43004300u;
  static constexpr float kMagicF = 128.0f;
  static constexpr float kFastZp = 112.0f;
  __device__ static __forceinline__ T2 bcast(float v) { return __float2bfloat162_rn(v); }
};

__device__ __forceinline__ uint32_t ta_reg(uint32_t v) {
  uint32_t d;
  asm volatile("mov.b32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
template <int SH>
__device__ __forceinline__ uint32_t ta_shr(uint32_t w) {              // w >> SH on the FMA pipe
  if (SH == 0) return w;
  uint32_t d;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(w), "n"(SH == 0 ? 1u : (1u << (32 - SH))));
  return d;
}
template <uint32_t MASK>
__device__ __forceinline__ uint32_t ta_and_or(uint32_t w, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "n"(MASK), "r"(orv));
  return d;
}
__device__ __forceinline__ void ta_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ta_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// column c (pair of k = 2c, 2c+1 inside the 64-k block) of a row's codes, as (MAGIC | q) or raw pair
template <int NBITS>
__device__ __forceinline__ uint32_t ta_pair_magic(const uint32_t* w, int c, uint32_t magic) {
  if (NBITS == 4) {
    const int i = c >> 2, j = c & 3;
    return j == 0 ? ta_and_or<0x000f000fu>(w[i], magic)
         : j == 1 ? ta_and_or<0x000f000fu>(ta_shr<4>(w[i]), magic)
         : j == 2 ? ta_and_or<0x000f000fu>(ta_shr<8>(w[i]), magic)
                  : ta_and_or<0x000f000fu>(ta_shr<12>(w[i]), magic);
  } else if (NBITS == 2) {
    const int i = c >> 3, j = c & 7;
    const uint32_t x = j == 0 ? w[i] : j == 1 ? ta_shr<2>(w[i]) : j == 2 ? ta_shr<4>(w[i]) : j == 3 ? ta_shr<6>(w[i])
                     : j == 4 ? ta_shr<8>(w[i]) : j == 5 ? ta_shr<10>(w[i]) : j == 6 ? ta_shr<12>(w[i]) : ta_shr<14>(w[i]);
    return ta_and_or<0x00030003u>(x, magic);
  } else if (NBITS == 3) {
    // w[0..3]: the row's 2-bit plane (8 pairs per word), w[4..5]: its 1-bit plane (16 pairs per word)
    const int i = c >> 3, j = c & 7;
    const uint32_t lo = (w[i] >> (2 * j)) & 0x00030003u;
    const uint32_t hi = ((w[4 + (c >> 4)] >> (c & 15)) & 0x00010001u) << 2;
    return lo | hi | magic;
  } else {   // 8 bits: pair = (w >> 8j) & 0x00ff00ff
    const int i = c >> 1, j = c & 1;
    return __byte_perm(w[i], magic, j ? 0x7371 : 0x7270);
  }
}

// CL = cluster size along N: the CL CTAs of a cluster work on CL consecutive weight-row tiles of the SAME token
// tile; each loads 1/CL of the x box and multicasts it to all of them, which divides the L2 -> SM traffic of
// the x re-reads (the limiter of the single-CTA version: 1.9 GB per up_proj GEMM, 8.8 TB/s) by CL.
#ifdef TA_TIMING
__device__ long long g_ta_t[256][8];
#define TA_T(v) const long long v = clock64()
#else
#define TA_T(v)
#endif

template <typename T, int NBITS, int NT, int CL>
__global__ void __launch_bounds__(TA_THREADS, 1)
k3_tca(const __grid_constant__ CUtensorMap tm_x, const TcaArgs a) {
  using SM = TcaSmem<NBITS, NT>;
  using DQ = TaDQ<T>;
  using T2 = typename DQ::T2;
  constexpr bool kBf16 = DT<T>::kCode == Q100_BF16;
  constexpr int SX = SM::kSX;
  constexpr int NW = SM::kP / TA_ROWS / 4;             // code words per row per k block (W4: 8)

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[SX + SX + TA_SP + TA_SP + TA_SA + TA_SA + 2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t b0 = smem_u32(&bars[0]);
  const uint32_t full_x = b0, empty_x = full_x + 8 * SX, full_p = empty_x + 8 * SX, empty_p = full_p + 8 * TA_SP;
  const uint32_t a_full = empty_p + 8 * TA_SP, a_empty = a_full + 8 * TA_SA, d_full = a_empty + 8 * TA_SA, d_empty = d_full + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int num_kb = a.K / TA_BK;
  // cluster tiles: (token tile, group of CL consecutive weight-row tiles); this CTA takes row tile `rank`
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int n_groups = (a.n_tiles + CL - 1) / CL;
  const int n_ct = a.m_tiles * n_groups;
  const int ct0 = blockIdx.x / CL, ct_step = gridDim.x / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (tid == 0) {
    for (int s = 0; s < SX; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[SX + s], CL); }
    for (int s = 0; s < TA_SP; ++s) { mbar_init(&bars[2 * SX + s], 1); mbar_init(&bars[2 * SX + TA_SP + s], 4); }
    for (int s = 0; s < TA_SA; ++s) { mbar_init(&bars[2 * SX + 2 * TA_SP + s], 4); mbar_init(&bars[2 * SX + 2 * TA_SP + TA_SA + s], 1); }
    mbar_init(&bars[2 * SX + 2 * TA_SP + 2 * TA_SA], 1);        // d_full
    mbar_init(&bars[2 * SX + 2 * TA_SP + 2 * TA_SA + 1], 4);    // d_empty
    fence_mbar_init();
  }
  if (warp == TA_MMA_WARP) tmem_alloc(&tmem_base_s, TA_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                    // every CTA's barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == TA_TMAX_WARP) {
    // ===================== TMA producer: x tiles =====================
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      const uint64_t pol_x = policy_evict_last();
      uint32_t it = 0;
      SegIter seg(a, n_ct, num_kb, ct0, ct_step);
      int ct, kb0, kb1;
      while (seg.next(ct, kb0, kb1)) {
        const int m0 = (ct % a.m_tiles) * NT;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % SX;
          ta_wait(empty_x + 8 * s, ((it / SX) & 1) ^ 1);     // all CL CTAs have retired their MMAs on this stage
#ifdef TA_EXP_NOX
          if (it >= SX) { mbar_arrive(&bars[s]); continue; }
#endif
          mbar_expect_tx(&bars[s], SM::kX);                  // the whole box: own slice + the peers' slices
          if (CL == 1) {
            tma_load_2d_hint(reinterpret_cast<void*>(smem_raw + (base_u32 - smem_u32(smem_raw)) + s * SM::kX), &tm_x,
                             &bars[s], kb * TA_BK, m0, pol_x);
          } else {
            tma_load_2d_mc(base_u32 + s * SM::kX + rank * (NT / CL) * 128, &tm_x, full_x + 8 * s, kb * TA_BK,
                           m0 + rank * (NT / CL), kMask, pol_x);
          }
        }
      }
    }
  } else if (warp == TA_TMAP_WARP) {
    // ===================== producer: packed codes + side band =====================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();
      uint8_t* pbase = smem_raw + (base_u32 - smem_u32(smem_raw)) + SM::kOffP;
      uint32_t it = 0;
      SegIter seg(a, n_ct, num_kb, ct0, ct_step);
      int ct, kb0, kb1;
      while (seg.next(ct, kb0, kb1)) {
        const int nt = min((ct / a.m_tiles) * CL + rank, a.n_tiles - 1);   // ragged last group: redo the last tile
        const int n0 = nt * TA_ROWS;
        const int mbytes = min(TA_ROWS, a.N - n0) * 4;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % TA_SP;
          ta_wait(empty_p + 8 * s, ((it / TA_SP) & 1) ^ 1);
          uint8_t* st = pbase + s * SM::kPStage;
          mbar_expect_tx(&bars[2 * SX + s], SM::kP + mbytes);
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.packed) + (static_cast<int64_t>(nt) * num_kb + kb) * SM::kP;
          bulk_load_1d_hint(st, wsrc, SM::kP, &bars[2 * SX + s], pol_w);
          const int kg = (kb * TA_BK) / a.group;
          const T* src = reinterpret_cast<const T*>(a.meta) + (static_cast<int64_t>(kg) * a.N + n0) * 2;
          bulk_load_1d(st + SM::kP, src, mbytes, &bars[2 * SX + s]);
        }
      }
    }
  } else if (warp == TA_MMA_WARP) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(kBf16 ? 1 : 0, TA_ROWS, NT);
    uint32_t it = 0, tcount = 0;
#ifdef TA_TIMING
    long long t_de = 0, t_af = 0, t_fx = 0, t_is = 0; const long long t_begin = clock64();
#endif
    SegIter seg(a, n_ct, num_kb, ct0, ct_step);
    int ct, kb0, kb1;
    for (; seg.next(ct, kb0, kb1); ++tcount) {
      TA_T(q0);
      ta_wait(d_empty, (tcount & 1) ^ 1);            // the epilogue has drained the accumulator
      tc_fence_after();
#ifdef TA_TIMING
      t_de += clock64() - q0;
#endif
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t sx = it % SX, sa = it % TA_SA;
        TA_T(q1);
        ta_wait(a_full + 8 * sa, (it / TA_SA) & 1);
        TA_T(q2);
        ta_wait(full_x + 8 * sx, (it / SX) & 1);
        TA_T(q3);
#ifdef TA_TIMING
        t_af += q2 - q1; t_fx += q3 - q2;
#endif
        tc_fence_after();
        if (lane == 0) {
          const uint64_t bd = umma_desc_sw128_u32(base_u32 + sx * SM::kX);
          const uint32_t at = tmem + TA_COL_A + sa * 32;
#pragma unroll
          for (int k = 0; k < TA_BK / 16; ++k)
            umma_f16_ts(tmem, at + 8 * k, bd + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);   // a segment starts from zero
          umma_commit_u32(a_empty + 8 * sa);
          if (CL == 1) umma_commit_u32(empty_x + 8 * sx);
          else umma_commit_mc(empty_x + 8 * sx, kMask);      // frees the stage in every CTA of the cluster
          if (kb == kb1 - 1) umma_commit_u32(d_full);
        }
        __syncwarp();
#ifdef TA_TIMING
        t_is += clock64() - q3;
#endif
      }
    }
#ifdef TA_TIMING
    if (lane == 0 && blockIdx.x < 256) { g_ta_t[blockIdx.x][0] = clock64() - t_begin; g_ta_t[blockIdx.x][1] = t_de; g_ta_t[blockIdx.x][2] = t_af; g_ta_t[blockIdx.x][3] = t_fx; g_ta_t[blockIdx.x][4] = t_is; g_ta_t[blockIdx.x][5] = it; }
#endif
  } else if (warp < TA_DQ_WARPS) {
    // ===================== dequant warps: group g handles iterations it % 2 == g; thread = weight row =====================
    const int grp = warp >> 2;
    const int row = tid & 127;
    const uint32_t lane_t = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + TA_COL_A;
    const uint32_t qb = DT<T>::pack2(a.qmin, a.qmin);
    const T2 qmin2 = *reinterpret_cast<const T2*>(&qb);
    const T2 magic2 = DQ::bcast(DQ::kMagicF);
    const uint32_t magic = ta_reg(DQ::kMagic);
    const bool has_zp = a.has_zp != 0;
    // W3: a tile is its 2-bit plane (128 rows x 16 bytes) followed by its 1-bit plane (128 rows x 8 bytes)
    const uint32_t code_off = ta_reg(base_u32 + SM::kOffP + row * (NBITS == 3 ? 16 : NW * 4));
    const uint32_t code_off_hi = ta_reg(base_u32 + SM::kOffP + 128 * 16 + row * 8);
    const uint32_t meta_off = ta_reg(base_u32 + SM::kOffP + SM::kP + row * 4);
    uint32_t it = 0;
    SegIter seg(a, n_ct, num_kb, ct0, ct_step);
    int ct, kb0, kb1;
    while (seg.next(ct, kb0, kb1)) {
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        if (static_cast<int>(it & 1) != grp) continue;
        const uint32_t sp = it % TA_SP, sa = it % TA_SA;
        TA_T(d0);
        ta_wait(full_p + 8 * sp, (it / TA_SP) & 1);
        TA_T(d1);
        uint32_t w[NW];
        if constexpr (NBITS == 3) {
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(code_off + sp * SM::kPStage));
          w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[4]), "=r"(w[5]) : "r"(code_off_hi + sp * SM::kPStage));
        } else {
#pragma unroll
          for (int i = 0; i < NW / 4; ++i) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(code_off + sp * SM::kPStage + i * 16));
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
        }
        uint32_t m;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(meta_off + sp * SM::kPStage));
        __syncwarp();
        if (lane == 0) ta_arrive(empty_p + 8 * sp);          // the codes are in registers
        const uint32_t sb2 = __byte_perm(m, 0, 0x1010), zb = __byte_perm(m, 0, 0x3232);
        const T2 s2 = *reinterpret_cast<const T2*>(&sb2), z2 = *reinterpret_cast<const T2*>(&zb);
        const float zf = DT<T>::to_f(*reinterpret_cast<const T*>(&zb));
        // fast path: magic + zp (or magic - qmin) is an exactly representable integer; warp-uniform choice
        const bool fast = __all_sync(0xffffffffu, !has_zp || fabsf(zf) <= DQ::kFastZp);
        const T2 sub2 = DQ::bcast(DQ::kMagicF + (has_zp ? zf : -a.qmin));
        uint32_t av[32];
        if constexpr (NBITS == 8 && kBf16) {
          // 8-bit codes do not fit bf16's OR-magic trick (128 + q needs 9 significant bits): go through fp32
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint32_t pr = (w[c >> 1] >> (8 * (c & 1))) & 0x00ff00ffu;
            const uint32_t cb = DT<T>::pack2(static_cast<float>(pr & 0xffffu), static_cast<float>(pr >> 16));
            T2 v = *reinterpret_cast<const T2*>(&cb);
            v = has_zp ? __hsub2(v, z2) : __hadd2(v, qmin2);
            v = __hmul2(v, s2);
            av[c] = *reinterpret_cast<uint32_t*>(&v);
          }
        } else if (fast) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint32_t pm = ta_pair_magic<NBITS>(w, c, magic);
            T2 v = __hsub2(*reinterpret_cast<const T2*>(&pm), sub2);       // (magic + q) - (magic + zp): exact
            v = __hmul2(v, s2);                                            // r(. * scale)
            av[c] = *reinterpret_cast<uint32_t*>(&v);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint32_t pm = ta_pair_magic<NBITS>(w, c, magic);
            T2 v = __hsub2(*reinterpret_cast<const T2*>(&pm), magic2);     // exact
            v = has_zp ? __hsub2(v, z2) : __hadd2(v, qmin2);               // r(q - zp)
            v = __hmul2(v, s2);
            av[c] = *reinterpret_cast<uint32_t*>(&v);
          }
        }
        TA_T(d2);
        ta_wait(a_empty + 8 * sa, ((it / TA_SA) & 1) ^ 1);   // the tensor core is done with this A stage
        TA_T(d3);
        tc_fence_after();
#ifndef TA_EXP_NOST
        tmem_st_32x32b_x32(lane_t + sa * 32, av);
        tmem_st_wait();
#else
        if (av[0] == 0x12345678u && av[31] == 0x9abcdef0u) tmem_st_32x32b_x32(lane_t + sa * 32, av);
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) ta_arrive(a_full + 8 * sa);
#ifdef TA_TIMING
        if (tid == 0 && blockIdx.x < 256) {
          g_ta_t[blockIdx.x][6] += (d1 - d0) + (d3 - d2);       // waiting (full_p + a_empty)
          g_ta_t[blockIdx.x][7] += (d2 - d1) + (clock64() - d3); // working (loads + math, store + publish)
        }
#endif
      }
    }
  } else {
    // ===================== epilogue warps (8..11): lane = weight row, columns = tokens =====================
    const int quad = warp & 3;
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t tcount = 0;
    SegIter seg(a, n_ct, num_kb, ct0, ct_step);
    int ct, kb0, kb1;
    for (; seg.next(ct, kb0, kb1); ++tcount) {
      const int m0 = (ct % a.m_tiles) * NT;
      const int nt = (ct / a.m_tiles) * CL + rank;      // >= n_tiles in a ragged last group: nothing to store
      const int n = nt < a.n_tiles ? nt * TA_ROWS + quad * 32 + lane : a.N;
      // stream-K (leftover tiles only): this cluster holds k blocks [kb0, kb1) of a tile that clusters c_first .. c_last
      // share.  c_first (it owns the tile's first k block) is the finalizer: it adds the others' fp32 partials to its
      // accumulator and writes y; the others leave their accumulator in their slot.  A cluster whose range crosses a
      // tile boundary works on the END of one tile first (as a contributor) and on the START of the next one second
      // (as its finalizer), so a finalizer only ever waits for segments that are FIRST in their cluster's range -- with
      // the roles the other way round every cluster waited for its predecessor's second segment, a serial chain over
      // all clusters (measured: 751 instead of 201 us on up_proj).
      const bool split = (kb0 > 0 || kb1 < num_kb);
      int c_first = 0, c_last = 0;
      bool contributor = false;
      const int C = ct_step, me = ct0;
      const int n_full_end = (n_ct / C) * C;
      const long long UL = static_cast<long long>(n_ct - n_full_end) * num_kb;
      if (split) {
        const int tl = ct - n_full_end;
        c_first = sk_owner(tl * num_kb, C, UL);
        c_last = sk_owner((tl + 1) * num_kb - 1, C, UL);
        contributor = me != c_first;
      }
      auto slot_of = [&](int c) {                          // cluster c's slot for THIS tile: 0 = its range starts in it
        const int first_tile = n_full_end + static_cast<int>((UL * c / C) / num_kb);
        return (c * 2 + (ct == first_tile ? 0 : 1)) * CL + rank;
      };
      float* my_part = a.ws_part + static_cast<int64_t>(slot_of(me)) * (TA_ROWS * NT) + quad * 32 + lane;
      if (split && !contributor && nt < a.n_tiles) {
        if (quad == 0 && lane == 0) {
          for (int c = c_first + 1; c <= c_last; ++c) {
            uint32_t spins = 0;
            while (ld_acquire_gpu(a.ws_flag + slot_of(c)) == 0u) { if (++spins > (1u << 26)) __trap(); }
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");   // the four epilogue warps
      }
      ta_wait(d_full, tcount & 1);
      tc_fence_after();
      const float bv = (bias != nullptr && n < a.N) ? DT<T>::to_f(bias[n]) : 0.f;
      // y is [tokens, rows]: the accumulator has a row per lane, so a chunk of [32 tokens x 32 rows] is
      // transposed through shared memory and leaves as 16-byte stores (8 rows of one token each) -- the
      // direct form, one 2-byte store per value, kept the tensor core idle for ~6 us per tile
      const int n_warp = nt * TA_ROWS + quad * 32;                      // first row of this warp
      // destination of this token tile: y, or the owner rank's staging slot (rows_per % NT == 0: one owner per tile)
      T* ydst = reinterpret_cast<T*>(a.y);
      int m_shift = 0;                                                  // row index in the destination = m - m_shift
      if (a.peer_stage != nullptr) {
        const int owner = m0 / a.rows_per;
        ydst = reinterpret_cast<T*>(a.peer_stage[owner]) + static_cast<int64_t>(a.rank) * a.rows_per * a.N;
        m_shift = owner * a.rows_per;
      }
      const uint32_t ebase = base_u32 + SM::kOffE + quad * (2 * 2048);
      const int tj = lane >> 2, rq = (lane & 3) * 8;                    // read side: token tj (+8i), rows rq..rq+7
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_t + c0, r);
        tmem_ld_wait();
        if (c0 + 32 >= NT) {                         // last chunk read: the accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) ta_arrive(d_empty);
        }
        if (contributor) {
          // fp32 partial, [column][row]: a warp's store of one column is 128 contiguous bytes
          if (nt < a.n_tiles) {
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(my_part + (c0 + j) * TA_ROWS, __uint_as_float(r[j]));
          }
          continue;
        }
        if (split && nt < a.n_tiles) {
          for (int c = c_first + 1; c <= c_last; ++c) {    // fixed order: deterministic
            float* pc = a.ws_part + static_cast<int64_t>(slot_of(c)) * (TA_ROWS * NT) + quad * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldcg(pc + (c0 + j) * TA_ROWS));
            // the workspace is shared with the decode kernels' fix-up slots, which expect zeros: leave it clean
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(pc + (c0 + j) * TA_ROWS, 0.f);
          }
        }
        const uint32_t eb = ebase + ((c0 >> 5) & 1) * 2048;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const T v = DT<T>::from_f(__uint_as_float(r[j]) + bv);
          asm volatile("st.shared.u16 [%0], %1;" :: "r"(eb + j * 64 + lane * 2), "h"(*reinterpret_cast<const uint16_t*>(&v)) : "memory");
        }
        __syncwarp();
        if (nt < a.n_tiles) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = m0 + c0 + tj + 8 * i;
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(eb + (tj + 8 * i) * 64 + rq * 2));
            if (m < a.M && n_warp + rq < a.N)
              *reinterpret_cast<uint4*>(ydst + static_cast<int64_t>(m - m_shift) * a.N + n_warp + rq) = v;
          }
        }
        // the buffer is reused two chunks later: the __syncwarp of the next chunk orders these reads before
        // the writes after it
      }
      if (nt < a.n_tiles && split) {
        if (contributor) __threadfence();                  // this thread's partial stores are visible gpu-wide ...
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (quad == 0 && lane == 0) {
          if (contributor) st_release_gpu(a.ws_flag + slot_of(me), 1u);      // ... before the flag
          else for (int c = c_first + 1; c <= c_last; ++c) a.ws_flag[slot_of(c)] = 0u;   // consumed: clean for the next launch
        }
      }
    }
    if (a.peer_stage != nullptr) __threadfence_system();   // the peer stores are performed before the grid completes
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                    // nobody leaves while a peer may still multicast into it
  if (warp == TA_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, TA_TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host
#ifdef TA_TIMING
int tca_timing(long long* out, int n) {
  int rc = check_cuda(cudaMemcpyFromSymbol(out, g_ta_t, sizeof(long long) * 8 * n), "timing readback");
  static long long zeros[256][8] = {};
  cudaMemcpyToSymbol(g_ta_t, zeros, sizeof(zeros));
  return rc;
}
#endif

bool gemm_tca_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits) {
  const int64_t g = group == 0 ? K : group;
  return (n_bits == 4 || n_bits == 2 || n_bits == 3 || n_bits == 8) && M >= 1 && N % 8 == 0 && K % TA_BK == 0 &&
         g % TA_BK == 0 && K % g == 0;
}

static thread_local size_t g_tca_ws_bytes = 0;     // size of the workspace behind TcaArgs::ws_flag for the current call

template <typename T, int NBITS, int NT, int CL>
static int launch_tca(const void* x, TcaArgs a, cudaStream_t st) {
  using SM = TcaSmem<NBITS, NT>;
  auto kern = k3_tca<T, NBITS, NT, CL>;
  static std::mutex mu;                          // one per instantiation
  static bool attr_done[kMaxDevices] = {};       // the attribute is a per-device property
  static int max_clusters[kMaxDevices] = {};
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_done[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
      attr_done[dev] = true;
    }
  }
  CUtensorMap tx;
  int rc = make_tmap_2d(&tx, x, 2, a.K, a.M, static_cast<uint64_t>(a.K) * 2, TA_BK, NT / CL, 128);
  if (rc != Q100_OK) return rc;
  a.m_tiles = (a.M + NT - 1) / NT;
  a.n_tiles = (a.N + TA_ROWS - 1) / TA_ROWS;
  const int n_ct = a.m_tiles * ((a.n_tiles + CL - 1) / CL);
  int clusters = sm_count() / CL;
  if (clusters > n_ct) clusters = n_ct;
  const bool want_sk = a.stream_k != 0;           // the caller has a workspace; decided below once the grid is known
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(TA_THREADS);
  cfg.dynamicSmemBytes = SM::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  if (CL > 1) {
    // persistent kernel: never launch more clusters than can be resident at once (GPC packing: clusters of
    // four do not tile all 148 SMs)
    std::lock_guard<std::mutex> lk(mu);
    if (max_clusters[dev] == 0) {
      cfg.gridDim = dim3(sm_count() / CL * CL);
      Q100_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters[dev], kern, &cfg));
      if (max_clusters[dev] < 1) max_clusters[dev] = 1;
    }
    if (clusters > max_clusters[dev]) clusters = max_clusters[dev];
    cfg.gridDim = dim3(clusters * CL);
  }
  // stream-K for the leftover tiles when the tiles do not divide evenly over the clusters (896 tiles of up_proj at
  // M = 2048 on 37 clusters of 4: 6.05 rounds that cost 7) and there is at least one full round
  const int num_kb = a.K / TA_BK;
  a.stream_k = 0;
  if (want_sk && n_ct % clusters != 0 && n_ct > clusters && num_kb >= 8) {
    const size_t flag_bytes = (static_cast<size_t>(clusters) * 2 * CL * 4 + 255) & ~static_cast<size_t>(255);
    const size_t need = flag_bytes + static_cast<size_t>(clusters) * 2 * CL * TA_ROWS * NT * 4;
    if (a.ws_flag != nullptr && a.ws_part == nullptr && need <= g_tca_ws_bytes) {
      a.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.ws_flag) + flag_bytes);
      a.stream_k = 1;
    }
  }
  if (!a.stream_k) { a.ws_flag = nullptr; a.ws_part = nullptr; }
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tx, a), "k3_tca launch");
}

template <typename T, int NBITS>
static int gemm_tca_nt(const void* x, const TcaArgs& a, int cluster, cudaStream_t st) {
  const int n_tiles = (a.N + TA_ROWS - 1) / TA_ROWS;
  // 128-token tiles only when there are no more tokens: measured (profiles/r01 notes) a 128-token tile takes
  // as long as a 256-token one (the per-k-block pipeline, not the MMA, sets the pace), so it never pays
  static const int env_nt = []() { const char* e = getenv("Q100_TCA_NT"); return e ? atoi(e) : 0; }();   // experiments
  if (a.M <= 128 || env_nt == 128) {
    if (cluster >= 2 && n_tiles >= 2) return launch_tca<T, NBITS, 128, 2>(x, a, st);
    return launch_tca<T, NBITS, 128, 1>(x, a, st);
  }
  if (cluster >= 4 && n_tiles >= 4) return launch_tca<T, NBITS, 256, 4>(x, a, st);
  if (cluster >= 2 && n_tiles >= 2) return launch_tca<T, NBITS, 256, 2>(x, a, st);
  return launch_tca<T, NBITS, 256, 1>(x, a, st);
}

template <typename T>
static int gemm_tca_t(const void* x, const TcaArgs& a, int n_bits, int cluster, cudaStream_t st) {
  switch (n_bits) {
    case 2: return gemm_tca_nt<T, 2>(x, a, cluster, st);
    case 3: return gemm_tca_nt<T, 3>(x, a, cluster, st);
    case 8: return gemm_tca_nt<T, 8>(x, a, cluster, st);
    default: return gemm_tca_nt<T, 4>(x, a, cluster, st);
  }
}

int gemm_tca(const void* x, const void* packed, const void* meta, const void* bias, void* y,
             int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
             int cluster, cudaStream_t st, void* const* peer_stage, int rows_per, int rank, int world,
             void* workspace, size_t workspace_bytes) {
  Q100_REQUIRE(gemm_tca_supported(M, N, K, group, n_bits), Q100_EUNSUPPORTED, "tcgen05 (TMEM-A) path: unsupported shape");
  TcaArgs a{};
  // Off unless Q100_TCA_STREAMK=1.  Measured on B200 at M = 2048 (tools/exp_streamk_tca.py, profiles/r02_streamk_tca.txt):
  // cutting the leftover tiles along K balances the last round (6.05 instead of 7 tile times on up_proj) but the fp32
  // partials (write, read, re-zero: ~70 MB) and the short segments' pipeline fill cost more than that: up_proj 213.6
  // vs 201.3 us, q_proj 83.0 vs 67.5, 70B down_proj 771 vs 791 (the one shape that gains).
  static const bool no_sk = []() { const char* e = getenv("Q100_TCA_STREAMK"); return !(e && e[0] == '1'); }();
  // the stream-K slots live behind the decode kernels' share of the workspace (all zero between launches)
  constexpr size_t kSkOff = 1 << 22;
  if (workspace != nullptr && workspace_bytes > kSkOff && peer_stage == nullptr && !no_sk) {
    a.stream_k = 1;
    a.ws_flag = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + kSkOff);
    g_tca_ws_bytes = workspace_bytes - kSkOff;
  } else {
    g_tca_ws_bytes = 0;
  }
  a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  a.peer_stage = peer_stage; a.rows_per = rows_per; a.rank = rank; a.world = world;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K);
  a.group = group == 0 ? static_cast<int>(K) : group;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  if (dtype == Q100_F16) return gemm_tca_t<__half>(x, a, n_bits, cluster, st);
  return gemm_tca_t<__nv_bfloat16>(x, a, n_bits, cluster, st);
}

}  // namespace q100
