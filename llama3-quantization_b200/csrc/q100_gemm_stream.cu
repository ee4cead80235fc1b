// q100_gemm_stream.cu -- K3a': weight-streaming dequant-GEMM on tcgen05 for decode shapes (M <= 16).
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// HBM-bound like K3a (q100_gemm_skinny.cu), but the contraction runs on the 5th-generation tensor
// cores with the WEIGHTS as the A operand in tensor memory, which removes what made the mma.sync
// kernel issue-bound (profiles/r01e): no per-warp HMMA issue, no x fragments through LDS, no staging
// pass over x, and the cost per output token is a tcgen05.ld column instead of more MMAs.
//
// Persistent CTAs, 2 per SM, 7 warps:
//   warp 4      weight producer: one cp.async.bulk per stage for the [128 rows x 128 k] code block (8 KB
//               contiguous in the tile-major packed layout) + one for its (scale, zero) row; static data,
//               so it runs ahead of griddepcontrol.wait (PDL);
//   warp 5      x producer: two TMA boxes [16 tokens x 64 k] (SWIZZLE_128B, rows >= M zero-filled by TMA)
//               = the K-major B operand of the block;
//   warps 0-3   thread = output row.  Per block: 64 bytes of codes -> registers -> "magic form"
//               (MAGIC + q), one LOP3 per pair, shifts on the FMA pipe -> tcgen05.st into one of two
//               A buffers in tensor memory (64 columns = 128 k).  One block later the same thread reads
//               its row of the two accumulators back (tcgen05.ld) and does the fp32 group step
//                   acc[m] += scale * (D[m] - (MAGIC + zp) * S[m]);
//   warp 6      allocates TMEM; one elected thread issues, per block, 8 x (M=128, N=16, K=16)
//               tcgen05.mma for D = A . x^T and 8 more with a constant all-ones A for S[m] = sum_k x[m, k]
//               (every row gets the same S; the tensor core is idle otherwise), then tcgen05.commit
//               publishes the accumulators and frees the stage.
// Work is split stream-K style in units of one block exactly like K3a; the fix-up protocol (per-warp
// arrival counters, last arriver sums the slots in slot order) is the same with a row-major slot layout.
#include <mutex>
#include <unordered_map>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

constexpr int ST_ROWS = 128;
constexpr int ST_KB = 128;
constexpr int ST_NT = 16;               // tokens per MMA (N of the UMMA); M <= 16
constexpr int ST_DQ_WARPS = 4;
constexpr int ST_WPROD_WARP = 4, ST_XPROD_WARP = 5, ST_MMA_WARP = 6;
constexpr int ST_THREADS = 7 * 32;
constexpr int ST_MAX_STAGES = 12;
constexpr int ST_OCC = 2;
constexpr int ST_XBYTES = 2 * ST_NT * 128;      // two [16 x 64] bf16 boxes
constexpr size_t ST_COUNTER_BYTES = 1 << 18;    // same workspace prefix as K3a
// tensor memory columns
constexpr uint32_t TM_A0 = 0, TM_A1 = 64, TM_ONES = 128, TM_DM0 = 144, TM_DM1 = 160, TM_DS0 = 176, TM_DS1 = 192;
constexpr uint32_t TM_COLS = 256;

struct StreamArgs {
  const void* packed;     // tile-major packed codes (layout v2)
  const void* meta;       // [K/group, N, 2] T
  const void* bias;       // [N] T or null
  void* y;                // [M, N] T
  float* ws_partials;
  int* ws_counters;
  int M, N, K, group;
  int kblocks, n_tiles, total_units, max_slots;
  int has_zp;
  float qmin;
  int stages;
  int pdl;
};

template <int NBITS> struct SGeom {
  static constexpr int kTileBytes = 128 * 8 * NBITS;       // one [128 x 64] tile
  static constexpr int kBlkBytes = 2 * kTileBytes;
  static constexpr int kStagePitch = (ST_XBYTES + kBlkBytes + 512 + 1023) & ~1023;
};

template <typename T> struct SMagic;
template <> struct SMagic<__half> {
  static constexpr uint32_t kMagic = 0x64006400u;
  static constexpr float kMagicF = 1024.0f;
  static constexpr uint32_t kOnes = 0x3c003c00u;
};
template <> struct SMagic<__nv_bfloat16> {
  static constexpr uint32_t kMagic = 0x43004300u;
  static constexpr float kMagicF = 128.0f;
  static constexpr uint32_t kOnes = 0x3f803f80u;
};

__device__ __forceinline__ uint4 st_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 st_lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t st_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <uint32_t MASK>
__device__ __forceinline__ uint32_t st_and_or(uint32_t w, uint32_t orv) {       // (w & MASK) | orv: one LOP3
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "n"(MASK), "r"(orv));
  return d;
}
template <int SH>
__device__ __forceinline__ uint32_t st_shr(uint32_t w) {                        // w >> SH on the FMA pipe
  uint32_t d;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(w), "n"(1u << (32 - SH)));
  return d;
}
__device__ __forceinline__ uint32_t st_reg(uint32_t v) {
  uint32_t d;
  asm volatile("mov.b32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
__device__ __forceinline__ void st_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void st_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_tma_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ int st_unit_owner(int u, int G, int U) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * G + U - 1) / U) - 1;
}

// 64 codes (one row, 64 k = 8 words... here: `NW` words) -> (MAGIC + q) pairs in column order
template <int NBITS>
__device__ __forceinline__ void st_extract(const uint32_t* w, uint32_t magic, uint32_t* out);
// W4: word covers 8 k; pair j = (w >> 4j) & 0x000f000f
template <>
__device__ __forceinline__ void st_extract<4>(const uint32_t* w, uint32_t magic, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    out[4 * i + 0] = st_and_or<0x000f000fu>(w[i], magic);
    out[4 * i + 1] = st_and_or<0x000f000fu>(st_shr<4>(w[i]), magic);
    out[4 * i + 2] = st_and_or<0x000f000fu>(st_shr<8>(w[i]), magic);
    out[4 * i + 3] = st_and_or<0x000f000fu>(st_shr<12>(w[i]), magic);
  }
}
// W2: word covers 16 k; pair j = (w >> 2j) & 0x00030003 (4 words per 64 k)
template <>
__device__ __forceinline__ void st_extract<2>(const uint32_t* w, uint32_t magic, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[8 * i + 0] = st_and_or<0x00030003u>(w[i], magic);
    out[8 * i + 1] = st_and_or<0x00030003u>(st_shr<2>(w[i]), magic);
    out[8 * i + 2] = st_and_or<0x00030003u>(st_shr<4>(w[i]), magic);
    out[8 * i + 3] = st_and_or<0x00030003u>(st_shr<6>(w[i]), magic);
    out[8 * i + 4] = st_and_or<0x00030003u>(st_shr<8>(w[i]), magic);
    out[8 * i + 5] = st_and_or<0x00030003u>(st_shr<10>(w[i]), magic);
    out[8 * i + 6] = st_and_or<0x00030003u>(st_shr<12>(w[i]), magic);
    out[8 * i + 7] = st_and_or<0x00030003u>(st_shr<14>(w[i]), magic);
  }
}

// ---------------------------------------------------------------- kernel
template <typename T, int NBITS, int MB>
__global__ void __launch_bounds__(ST_THREADS, ST_OCC)
k3_stream(const __grid_constant__ CUtensorMap tm_x, const StreamArgs a) {
  using G = SGeom<NBITS>;
  using MG = SMagic<T>;
  constexpr bool kBf16 = DT<T>::kCode == Q100_BF16;
  constexpr int kTileB = G::kTileBytes, kBlkB = G::kBlkBytes, kPitch = G::kStagePitch;
  constexpr int kRowB = kTileB / 128;                  // bytes of one row inside a [128 x 64] tile
  constexpr int NW = kRowB / 4;                        // words per row per tile (W4: 8, W2: 4)
  constexpr int FRAG = ST_ROWS * MB;                   // floats per partial tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * ST_MAX_STAGES + 6];     // full[s], empty[s], a_full[2], d_full[2], d_empty[2]
  __shared__ uint32_t tmem_base_s;
  const uint32_t ring_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = a.stages;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = full0 + 8 * ST_MAX_STAGES;
  const uint32_t afull0 = empty0 + 8 * ST_MAX_STAGES, dfull0 = afull0 + 16, dempty0 = dfull0 + 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G_ = gridDim.x, U = a.total_units, cta = blockIdx.x;
  const int q_begin = static_cast<int>(static_cast<int64_t>(cta) * U / G_);
  const int q_end = static_cast<int>(static_cast<int64_t>(cta + 1) * U / G_);
  const int kblocks = a.kblocks;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars[s], 2);                                   // weight producer + x producer (expect_tx each)
      mbar_init(&bars[ST_MAX_STAGES + s], ST_DQ_WARPS + 1);     // 4 dequant warps + the MMA commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[2 * ST_MAX_STAGES + b], ST_DQ_WARPS);     // a_full
      mbar_init(&bars[2 * ST_MAX_STAGES + 2 + b], 1);           // d_full (commit)
      mbar_init(&bars[2 * ST_MAX_STAGES + 4 + b], ST_DQ_WARPS); // d_empty
    }
    fence_mbar_init();
  }
  if (warp == ST_MMA_WARP) tmem_alloc(&tmem_base_s, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_launch_dependents();

  if (warp == ST_WPROD_WARP) {
    // ===================== weight producer (static data: runs ahead of the PDL wait) =====================
    if (lane < 2) {
      const uint64_t pol = policy_evict_first();
      const bool is_code = lane == 0;
      const int esz2 = 2 * static_cast<int>(sizeof(T));
      const int64_t mrow_pitch = static_cast<int64_t>(a.N) * esz2;
      const int gdiv = min(a.group / ST_KB, kblocks);            // k blocks per side-band row
      int tile = q_begin / kblocks, kblk = q_begin - tile * kblocks;
      const uint8_t* mtile = reinterpret_cast<const uint8_t*>(a.meta) + static_cast<int64_t>(tile) * ST_ROWS * esz2;
      int mbytes = min(ST_ROWS, a.N - tile * ST_ROWS) * esz2;
      const uint8_t* src = is_code ? reinterpret_cast<const uint8_t*>(a.packed) + static_cast<int64_t>(q_begin) * kBlkB
                                   : mtile + (kblk / gdiv) * mrow_pitch;
      uint32_t bytes = is_code ? kBlkB : mbytes;
      const uint32_t dst0 = ring_u32 + ST_XBYTES + (is_code ? 0 : kBlkB);
      uint32_t tx = kBlkB + mbytes;
      int gleft = gdiv - kblk % gdiv, kleft = kblocks - kblk;
      uint32_t soff = 0, boff = 0, ph = 1;
      int sleft = stages;
#pragma unroll 1
      for (int q = q_begin; q < q_end; ++q) {
        st_wait(empty0 + boff, ph);
        if (is_code) st_expect_tx(full0 + boff, tx);
        st_bulk(dst0 + soff, src, bytes, full0 + boff, pol);
        soff += kPitch; boff += 8;
        if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
        if (is_code) src += kBlkB;
        else if (--gleft == 0) { gleft = gdiv; src += mrow_pitch; }
        if (--kleft == 0) {
          kleft = kblocks; ++tile;
          mtile += ST_ROWS * esz2;
          mbytes = min(ST_ROWS, a.N - tile * ST_ROWS) * esz2;
          tx = kBlkB + mbytes;
          if (!is_code) { src = mtile; bytes = mbytes; gleft = gdiv; }
        }
      }
    }
  } else if (warp == ST_XPROD_WARP) {
    // ===================== x producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      pdl_wait();                                    // x is written by the previous kernel of the stream
      int kblk = q_begin % kblocks;
      uint32_t soff = 0, boff = 0, ph = 1;
      int sleft = stages;
#pragma unroll 1
      for (int q = q_begin; q < q_end; ++q) {
        st_wait(empty0 + boff, ph);
        st_expect_tx(full0 + boff, ST_XBYTES);
        st_tma_2d(ring_u32 + soff, &tm_x, full0 + boff, kblk * ST_KB, 0);
        st_tma_2d(ring_u32 + soff + ST_XBYTES / 2, &tm_x, full0 + boff, kblk * ST_KB + 64, 0);
        soff += kPitch; boff += 8;
        if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
        if (++kblk == kblocks) kblk = 0;
      }
    }
  } else if (warp == ST_MMA_WARP) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(kBf16 ? 1 : 0, 128, ST_NT);
    uint32_t soff = 0, boff = 0, ph = 0;
    int sleft = stages;
    uint32_t it = 0;
#pragma unroll 1
    for (int q = q_begin; q < q_end; ++q, ++it) {
      const uint32_t b = it & 1, bph = (it >> 1) & 1;
      st_wait(afull0 + 8 * b, bph);                  // the block's A operand is in tensor memory
      st_wait(full0 + boff, ph);                     // ... and its x slice in shared memory
      st_wait(dempty0 + 8 * b, bph ^ 1);             // the epilogue has drained this accumulator pair
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_t = tmem + (b ? TM_A1 : TM_A0);
        const uint32_t dm = tmem + (b ? TM_DM1 : TM_DM0), ds = tmem + (b ? TM_DS1 : TM_DS0);
        const uint64_t bx0 = umma_desc_sw128_u32(ring_u32 + soff);
        const uint64_t bx1 = umma_desc_sw128_u32(ring_u32 + soff + ST_XBYTES / 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t bd = (k < 4 ? bx0 : bx1) + 2 * (k & 3);      // +32 bytes per K = 16 step
#ifndef ST_EXP_NO_MMA
          umma_f16_ts(dm, a_t + 8 * k, bd, idesc, k != 0);
#ifndef ST_EXP_NO_ONES
          umma_f16_ts(ds, tmem + TM_ONES, bd, idesc, k != 0);
#endif
#endif
        }
        umma_commit_u32(dfull0 + 8 * b);
        umma_commit_u32(empty0 + boff);
      }
      __syncwarp();
      soff += kPitch; boff += 8;
      if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
    }
  } else {
    // ===================== dequant + epilogue warps: thread = output row =====================
    const int row = tid;                                          // 0..127 = TMEM lane
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const bool has_zp = a.has_zp != 0;
    const float zp_fix = -a.qmin;
    const uint32_t magic = st_reg(MG::kMagic);
    // constant all-ones A operand (one K = 16 step)
    {
      uint32_t ones[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ones[i] = MG::kOnes;
      tmem_st_32x32b_x8(lane_t + TM_ONES, ones);
      tmem_st_wait();
    }
    const uint32_t code_off = st_reg(ring_u32 + ST_XBYTES + row * kRowB);
    const uint32_t meta_off = st_reg(ring_u32 + ST_XBYTES + kBlkB + row * 4);

    float acc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc[m] = 0.f;

    auto write_rows = [&](const float (&v)[MB], int tile) {
      const int n = tile * ST_ROWS + row;
      if (n < a.N) {
        T* y = reinterpret_cast<T*>(a.y);
        const float bv = a.bias ? DT<T>::to_f(reinterpret_cast<const T*>(a.bias)[n]) : 0.f;
#pragma unroll
        for (int m = 0; m < MB; ++m)
          if (m < a.M) y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(v[m] + bv);
      }
    };
    // stream-K fix-up, same protocol as K3a: slot layout [slot][m][row]
    int pend_tile = -1, pend_old = 0, pend_n = 0;
    auto slot_base = [&](int tile) { return a.ws_partials + static_cast<int64_t>(tile) * a.max_slots * FRAG; };
    auto resolve_pending = [&]() {
      const int old = __shfl_sync(0xffffffffu, pend_old, 0);
      __syncwarp();
      if (old == pend_n - 1) {
        const float* wst = slot_base(pend_tile);
        float v[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) v[m] = 0.f;
        for (int sl = 0; sl < pend_n; ++sl) {
#pragma unroll
          for (int m = 0; m < MB; ++m) {
            float* pp = const_cast<float*>(wst) + static_cast<int64_t>(sl) * FRAG + m * ST_ROWS + row;
            v[m] += __ldcg(pp);
            *pp = 0.f;           // the workspace is shared with K3a, whose {value, flag} slots must find zeros
          }
        }
        write_rows(v, pend_tile);
        if (lane == 0) a.ws_counters[pend_tile * 8 + warp] = 0;
      }
      pend_tile = -1;
    };
    auto finish_segment = [&](int q, int seg_begin) {
      const int tile = q / kblocks;
      const int tile_q0 = tile * kblocks, tile_q1 = tile_q0 + kblocks;
      if (seg_begin == tile_q0 && q + 1 == tile_q1) {
        write_rows(acc, tile);
      } else {
        const int c_first = st_unit_owner(tile_q0, G_, U), c_last = st_unit_owner(tile_q1 - 1, G_, U);
        float* wsp = slot_base(tile) + static_cast<int64_t>(cta - c_first) * FRAG;
#pragma unroll
        for (int m = 0; m < MB; ++m) wsp[m * ST_ROWS + row] = acc[m];
        __syncwarp();
        int* cnt = a.ws_counters + tile * 8 + warp;
        if (lane == 0) asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(pend_old) : "l"(cnt) : "memory");
        pend_tile = tile;
        pend_n = c_last - c_first + 1;
      }
#pragma unroll
      for (int m = 0; m < MB; ++m) acc[m] = 0.f;
    };

    // epilogue of block `it`: acc[m] += scale * (D[m] - (MAGIC + zp) * S[m])
    auto epilogue = [&](uint32_t it, uint32_t meta) {
      const uint32_t b = it & 1, bph = (it >> 1) & 1;
      st_wait(dfull0 + 8 * b, bph);
      tc_fence_after();
      uint32_t dm[MB], ds[MB];
      tmem_ld_32x32b<MB>(lane_t + (b ? TM_DM1 : TM_DM0), dm);
      tmem_ld_32x32b<MB>(lane_t + (b ? TM_DS1 : TM_DS0), ds);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) st_arrive(dempty0 + 8 * b);
      const float2 sz = DT<T>::to_f2(meta);
      const float off = -(MG::kMagicF + (has_zp ? sz.y : zp_fix));
#pragma unroll
      for (int m = 0; m < MB; ++m)
        acc[m] = fmaf(sz.x, fmaf(off, __uint_as_float(ds[m]), __uint_as_float(dm[m])), acc[m]);
    };

    pdl_wait();                                      // y and the workspace belong to the stream order
    uint32_t soff = 0, boff = 0, ph = 0;
    int sleft = stages;
    uint32_t it = 0, meta_prev = 0;
    int seg_begin = q_begin, kb = q_begin % kblocks;
    bool prev_end = false;
    int prev_q = 0;
#pragma unroll 1
    for (int q = q_begin; q < q_end; ++q, ++it) {
      // ---- dequant block `it` into A buffer it & 1 (free: its last reader's accumulator was drained in
      //      the epilogue two iterations ago)
      st_wait(full0 + boff, ph);
      uint32_t w[2 * NW];
      if constexpr (NW == 8) {
        const uint4 v0 = st_lds128(code_off + soff), v1 = st_lds128(code_off + soff + 16);
        const uint4 v2 = st_lds128(code_off + soff + kTileB), v3 = st_lds128(code_off + soff + kTileB + 16);
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w; w[12] = v3.x; w[13] = v3.y; w[14] = v3.z; w[15] = v3.w;
      } else {
        const uint4 v0 = st_lds128(code_off + soff), v2 = st_lds128(code_off + soff + kTileB);
        w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v2.x; w[5] = v2.y; w[6] = v2.z; w[7] = v2.w;
      }
      const uint32_t meta = st_lds32(meta_off + soff);
      __syncwarp();
      if (lane == 0) st_arrive(empty0 + boff);       // codes are in registers (x stays until the MMAs retire)
      soff += kPitch; boff += 8;
      if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
      {
        uint32_t av[32];
        const uint32_t a_t = lane_t + ((it & 1) ? TM_A1 : TM_A0);
        st_extract<NBITS>(w, magic, av);                         // k 0..63
        tmem_st_32x32b_x32(a_t, av);
        st_extract<NBITS>(w + NW, magic, av);                    // k 64..127
        tmem_st_32x32b_x32(a_t + 32, av);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) st_arrive(afull0 + 8 * (it & 1));
      }
      // ---- epilogue of the previous block (its MMAs ran while we dequantised this one)
      if (q > q_begin) {
        epilogue(it - 1, meta_prev);
        if (pend_tile >= 0) resolve_pending();
        if (prev_end) { finish_segment(prev_q, seg_begin); seg_begin = prev_q + 1; }
      }
      meta_prev = meta;
      prev_q = q;
      prev_end = (++kb == kblocks);
      if (prev_end) kb = 0;
    }
    if (q_end > q_begin) {
      epilogue(it - 1, meta_prev);
      if (pend_tile >= 0) resolve_pending();
      finish_segment(q_end - 1, seg_begin);
      if (pend_tile >= 0) resolve_pending();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ST_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

// ---------------------------------------------------------------- host
struct StreamPlan { int n_tiles, total_units, max_slots, grid_cap; };

static StreamPlan plan_stream(int64_t N, int64_t K) {
  StreamPlan p;
  const int kblocks = static_cast<int>(K / ST_KB);
  p.n_tiles = static_cast<int>((N + ST_ROWS - 1) / ST_ROWS);
  p.total_units = p.n_tiles * kblocks;
  int grid = sm_count() * ST_OCC;
  if (grid > p.total_units) grid = p.total_units;
  if (grid < 1) grid = 1;
  p.grid_cap = grid;
  p.max_slots = static_cast<int>((static_cast<int64_t>(kblocks) * grid + p.total_units - 1) / p.total_units) + 2;
  return p;
}

size_t stream_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const StreamPlan p = plan_stream(N, K);
  const int mb = M <= 1 ? 1 : M <= 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16;
  return static_cast<size_t>(p.n_tiles) * p.max_slots * ST_ROWS * mb * sizeof(float) + ST_COUNTER_BYTES;
}

bool gemm_stream_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits) {
  const int64_t g = group == 0 ? K : group;
  return M >= 1 && M <= ST_NT && (n_bits == 4 || n_bits == 2) && K % ST_KB == 0 && N % 8 == 0 &&
         g % ST_KB == 0 && K % g == 0;
}

template <typename T, int NBITS, int MB>
static int launch_stream(const CUtensorMap& tx, StreamArgs a, cudaStream_t st) {
  auto kern = k3_stream<T, NBITS, MB>;
  constexpr int pitch = SGeom<NBITS>::kStagePitch;
  const int budget = (228 * 1024) / ST_OCC - 1024 /* driver reserve */ - 1024 /* static + alignment */ - 1024;
  int stages = budget / pitch;
  if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
  a.stages = stages;
  const int smem = stages * pitch + 1024;
  static std::mutex mu;
  static bool attr_done[kMaxDevices] = {};       // the attribute is a per-device property
  {
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (!attr_done[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_done[dev] = true;
    }
  }
  int grid = sm_count() * ST_OCC;
  if (grid > a.total_units) grid = a.total_units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(ST_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tx, a), "k3_stream launch");
}

template <typename T, int NBITS>
static int launch_stream_mb(const CUtensorMap& tx, const StreamArgs& a, cudaStream_t st) {
  if (a.M <= 1) return launch_stream<T, NBITS, 1>(tx, a, st);
  if (a.M <= 2) return launch_stream<T, NBITS, 2>(tx, a, st);
  if (a.M <= 4) return launch_stream<T, NBITS, 4>(tx, a, st);
  if (a.M <= 8) return launch_stream<T, NBITS, 8>(tx, a, st);
  return launch_stream<T, NBITS, 16>(tx, a, st);
}

int gemm_stream(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int pdl, cudaStream_t st) {
  Q100_REQUIRE(gemm_stream_supported(M, N, K, group, n_bits), Q100_EUNSUPPORTED, "tcgen05 decode path: unsupported shape");
  const int g = group == 0 ? static_cast<int>(K) : group;
  const StreamPlan p = plan_stream(N, K);
  const size_t need = stream_workspace_bytes(M, N, K);
  Q100_REQUIRE(workspace != nullptr && workspace_bytes >= need, Q100_EWORKSPACE,
               "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Q100_REQUIRE(static_cast<size_t>(p.n_tiles) * 8 * sizeof(int) <= ST_COUNTER_BYTES, Q100_EUNSUPPORTED,
               "N=%lld too large for the stream-K counter area", (long long)N);
  CUtensorMap tx;
  int rc = make_tmap_2d(&tx, x, 2, K, M, static_cast<uint64_t>(K) * 2, 64, ST_NT, 128);
  if (rc != Q100_OK) return rc;
  StreamArgs a{};
  a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  a.ws_counters = reinterpret_cast<int*>(workspace);
  a.ws_partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ST_COUNTER_BYTES);
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = g;
  a.kblocks = static_cast<int>(K / ST_KB);
  a.n_tiles = p.n_tiles; a.total_units = p.total_units; a.max_slots = p.max_slots;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  a.pdl = pdl;
  if (dtype == Q100_F16) {
    if (n_bits == 2) return launch_stream_mb<__half, 2>(tx, a, st);
    return launch_stream_mb<__half, 4>(tx, a, st);
  }
  if (n_bits == 2) return launch_stream_mb<__nv_bfloat16, 2>(tx, a, st);
  return launch_stream_mb<__nv_bfloat16, 4>(tx, a, st);
}

}  // namespace q100
