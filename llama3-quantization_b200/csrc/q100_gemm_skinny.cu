// q100_gemm_skinny.cu -- K3a: weight-streaming dequant-GEMM for decode shapes (M <= 32).
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// HBM-bound: the packed weights (n_bits/8 B per weight) are the only large stream and are read
// exactly once.  Structure (persistent CTAs, 3 per SM, 8 consumer warps + 1 producer warp):
//   * the producer warp streams the CTA's contiguous run of [128 rows x 128 k] code blocks (8 KB
//     at W4; the tile-major packed layout makes the run linear in HBM) plus the blocks'
//     (scale, zero) side band into a deep shared-memory ring with cp.async.bulk (TMA) and mbarrier
//     full/empty pairs: ~170 KB are in flight per SM from the first cycle, independent of how fast
//     the math runs (tools/micro/stream_bench3.cu: this ring alone sustains 6.37 TB/s = 97 % of the
//     measured copy bandwidth; a register ring tops out near 5.2 TB/s).  The stream does not
//     depend on the previous kernel, so with programmatic dependent launch (Q100_GEMM_PDL) the
//     ring fills while the previous kernel drains;
//   * consumer warp w owns 16 output rows of the tile: it pulls its slice of a landed block into
//     registers (conflict-free LDS.128: the two 64-k halves are 64 bytes apart modulo 128), frees
//     the slot, and feeds the codes in "magic form" -- (MAGIC + q) built with one LOP3 per pair --
//     to mma.sync.m16n8k16 as the A operand (x^T is B); the block's contribution is
//     scale * (acc_q - (MAGIC + zp) * S) in fp32, S = sum of x over the block.  fp16/W4 uses the
//     word's nibbles in place ((w & 0x00f000f0) | MAGIC = 1024 + 16 q goes to a second accumulator
//     that is scaled by 1/16), five ALU ops per eight codes.  With Q100_GEMM_EXACT_WEIGHTS (or
//     group < 128) the 16-bit fake-quantised weight r(r(q-zp)*scale) is rebuilt bit-exactly.
//     tcgen05 is deliberately not used here: at M <= 32 the operand would have to be written
//     back to shared memory (2 B/weight) and that traffic alone exceeds the time budget of the
//     0.5 B/weight HBM stream (DESIGN.md "K3a");
//   * x is tiny: the consumers stage the slab of x they need into shared memory themselves, in a
//     bank-conflict-free permuted order, together with the per-block sums S[m];
//   * work is split stream-K style over 148*3 CTAs in units of one block; a tile shared by several
//     CTAs is finished warp-asynchronously: each warp stores its fp32 fragment to its CTA's slot,
//     fences, bumps the counter of ITS 16-row fragment, and the last arriver sums that fragment's
//     slots in slot order (deterministic) and writes y.  No CTA-wide barrier there.
//
// Lane <-> k mapping: inside a 128-k block lane t (= lane & 3) owns k in [32t, 32t+32) for
// both operands, i.e. the mma "slot" (s, 2t+{0,1,8,9}) of k16-step s carries
// k = 32t + 4s + {0,1,2,3}.  The contraction is invariant under that permutation because
// A and B fragments use the same one.
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cmath>

#include "q100_common.cuh"
#include "q100_ptx.cuh"
#include "q100_chain.cuh"
#include "q100_quant.cuh"

namespace q100 {

constexpr int SK_ROWS = 128;       // output rows per tile (8 warps x 16)
constexpr int SK_KB = 128;         // k per block
constexpr int SK_CWARPS = 8;       // consumer warps
constexpr int SK_THREADS = (SK_CWARPS + 1) * 32;
constexpr int SK_CTHREADS = SK_CWARPS * 32;
constexpr int SK_MAX_STAGES = 16;
constexpr int SK_CHUNK_MAX = 128;   // k blocks of x staged in shared memory at a time
#ifndef SK_OCC
#define SK_OCC 3                   // resident CTAs per SM the kernel is sized for (2 for M > 16)
#endif
constexpr size_t SK_COUNTER_BYTES = 1 << 18;   // workspace prefix: per-fragment arrival counters (N <= 1 M rows)
constexpr int SK_MAX_SPLIT = 4;    // contributors per tile the stream-K split aims for on small problems
constexpr int SK_SPLIT_UNITS = 8;  // ... as long as that keeps a CTA at or below this many blocks
constexpr int SK_XPITCH = 320;     // bytes per (m, 128-k) x row in smem: 256 + 64 pad (conflict-free LDS.128)


// One [128 rows x 64 k] tile of codes; a block = two tiles.  In shared memory the second tile is placed
// kPad bytes past a multiple of 128 so that the eight lanes of an LDS phase (two rows x four k-quarters)
// hit eight different 16-byte bank groups.
template <int NBITS> struct WGeom {
  static constexpr int kTileBytes = NBITS == 3 ? 128 * 24 : 128 * 8 * NBITS;
  static constexpr int kPad = NBITS == 8 ? 16 : 64;
  static constexpr int kTileStride = kTileBytes + kPad;
  static constexpr int kCodeBytes = 2 * kTileStride;       // smem footprint of a block's codes (16-B multiple)
};
template <int NBITS> __host__ __device__ constexpr int stage_bytes(int nsub) {
  return WGeom<NBITS>::kCodeBytes + nsub * 512;
}
constexpr int SK_ICOL = 144;       // IMMA path: bytes per (block, column) of x limbs: 128 + 16 (S, coefficient; conflict-free LDS.128)
constexpr int SK_IPITCH = 3 * SK_ICOL;   // ... per (block, token): three signed 8-bit limbs
__host__ __device__ constexpr int slab_bytes(int m_rows, int chunk, int pitch) { return (chunk * m_rows + 1) * pitch; }

// ---------------------------------------------------------------- dequant helpers
template <typename T> struct DQ;
template <> struct DQ<__half> {
  using T2 = __half2;
  static constexpr uint32_t kMagic = 0x64006400u;      // 1024.0 in both halves
  static constexpr float kMagicF = 1024.0f;
  __device__ static __forceinline__ T2 magic_val() { return __float2half2_rn(1024.0f); }
};
template <> struct DQ<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr uint32_t kMagic = 0x43004300u;      // 128.0 in both halves (codes < 128)
  static constexpr float kMagicF = 128.0f;
  __device__ static __forceinline__ T2 magic_val() { return __float2bfloat162_rn(128.0f); }
};

// pair: two unsigned codes in the low bits of each 16-bit half.  Returns
// r(r(code + qmin - zp) * scale) per half -- the reference's dequant, quantizer.py:107-110.
template <typename T, int NBITS>
__device__ __forceinline__ uint32_t dq_pair(uint32_t pair, typename DQ<T>::T2 qmin2,
                                            typename DQ<T>::T2 zp2, typename DQ<T>::T2 s2, bool has_zp) {
  using T2 = typename DQ<T>::T2;
  T2 c;
  if (NBITS == 8 && DT<T>::kCode == Q100_BF16) {
    // bf16 has 8 significant bits: the OR-magic trick only covers codes < 128
    const float lo = static_cast<float>(pair & 0xffffu), hi = static_cast<float>(pair >> 16);
    const uint32_t b = DT<T>::pack2(lo, hi);
    c = *reinterpret_cast<const T2*>(&b);
  } else {
    const uint32_t m = pair | DQ<T>::kMagic;
    c = __hsub2(*reinterpret_cast<const T2*>(&m), DQ<T>::magic_val());   // exact
  }
  if (has_zp) c = __hsub2(c, zp2);             // r(q - zp)           (qmin == 0)
  else c = __hadd2(c, qmin2);                  // stored code + qmin  (exact small integers)
  c = __hmul2(c, s2);                          // r(. * scale)
  return *reinterpret_cast<uint32_t*>(&c);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// (w & MASK) | orv as ONE LOP3: the mask is the immediate, the magic constant stays in a register
// (left to itself the compiler emits two immediate-form LOP3s)
template <uint32_t MASK>
__device__ __forceinline__ uint32_t lop_and_or(uint32_t w, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "n"(MASK), "r"(orv));
  return d;
}
// opaque register copy of a constant (keeps ptxas from folding it back into an immediate)
__device__ __forceinline__ uint32_t reg_const(uint32_t v) {
  uint32_t d;
  asm volatile("mov.b32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}

// A lane's 32 codes (k in [32t, 32t+32) of one row's 128-k block) as 16 pairs, read from the two
// [128 x 64] tiles of a landed block: lane t reads k-tile (t >> 1), half (t & 1).
//   lane_off(row, t): byte offset of the lane's codes of `row` inside the block's smem image
//   kRow8:            byte distance to the lane's second row (row + 8)
template <int NBITS> struct LaneCodes;
template <> struct LaneCodes<4> {
  uint4 w;
  __device__ __forceinline__ uint32_t word(int i) const { return i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w; }
  __device__ static __forceinline__ int lane_off(int row, int t) {
    return (t >> 1) * WGeom<4>::kTileStride + row * 32 + (t & 1) * 16;
  }
  static constexpr int kRow8 = 8 * 32;
  __device__ __forceinline__ void load(uint32_t p, uint32_t) { w = lds128(p); }
  __device__ __forceinline__ uint32_t pair(int i) const {
    return (word(i >> 2) >> (4 * (i & 3))) & 0x000f000fu;
  }
  __device__ __forceinline__ uint32_t mpair(int i, uint32_t magic) const {     // pair(i) | magic
    return lop_and_or<0x000f000fu>(word(i >> 2) >> (4 * (i & 3)), magic);
  }
};
template <> struct LaneCodes<2> {
  uint2 w;
  __device__ static __forceinline__ int lane_off(int row, int t) {
    return (t >> 1) * WGeom<2>::kTileStride + row * 16 + (t & 1) * 8;
  }
  static constexpr int kRow8 = 8 * 16;
  __device__ __forceinline__ void load(uint32_t p, uint32_t) { w = lds64(p); }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? w.x : w.y;
    return (x >> (2 * (i & 7))) & 0x00030003u;
  }
  __device__ __forceinline__ uint32_t mpair(int i, uint32_t magic) const {
    const uint32_t x = (i >> 3) == 0 ? w.x : w.y;
    return lop_and_or<0x00030003u>(x >> (2 * (i & 7)), magic);
  }
};
template <> struct LaneCodes<8> {
  uint4 w0, w1;
  __device__ static __forceinline__ int lane_off(int row, int t) {
    return (t >> 1) * WGeom<8>::kTileStride + row * 64 + (t & 1) * 32;
  }
  static constexpr int kRow8 = 8 * 64;
  __device__ __forceinline__ void load(uint32_t p, uint32_t) {
    w0 = lds128(p);
    w1 = lds128(p + 16);
  }
  __device__ __forceinline__ uint32_t wsel(int wi) const {
    return wi == 0 ? w0.x : wi == 1 ? w0.y : wi == 2 ? w0.z : wi == 3 ? w0.w
         : wi == 4 ? w1.x : wi == 5 ? w1.y : wi == 6 ? w1.z : w1.w;
  }
  __device__ __forceinline__ uint32_t pair(int i) const { return (wsel(i >> 1) >> (8 * (i & 1))) & 0x00ff00ffu; }
  __device__ __forceinline__ uint32_t mpair(int i, uint32_t magic) const {
    return __byte_perm(wsel(i >> 1), magic, (i & 1) ? 0x7371 : 0x7270);      // [b, magic_hi, b', magic_hi]: one PRMT
  }
};
template <> struct LaneCodes<3> {
  uint2 lo;
  uint32_t hi;
  // 2-bit plane: 128 rows x 16 B at the tile start; 1-bit plane: 128 rows x 8 B right after it
  __device__ static __forceinline__ int lane_off(int row, int t) {
    return (t >> 1) * WGeom<3>::kTileStride + row * 16 + (t & 1) * 8;
  }
  __device__ static __forceinline__ int lane_off_hi(int row, int t) {
    return (t >> 1) * WGeom<3>::kTileStride + 128 * 16 + row * 8 + (t & 1) * 4;
  }
  static constexpr int kRow8 = 8 * 16;
  static constexpr int kRow8Hi = 8 * 8;
  __device__ __forceinline__ void load(uint32_t p_lo, uint32_t p_hi) {
    lo = lds64(p_lo);
    hi = lds32(p_hi);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? lo.x : lo.y;
    return ((x >> (2 * (i & 7))) & 0x00030003u) | (((hi >> i) & 0x00010001u) << 2);
  }
  __device__ __forceinline__ uint32_t mpair(int i, uint32_t magic) const { return pair(i) | magic; }
};

// everything a lane needs for one 128-k block: codes of its two rows + their (scale, zero) pairs
template <int NBITS>
struct BlockRegs {
  LaneCodes<NBITS> c0, c1;
  uint32_t m0, m1;
};

// owner CTA of stream-K unit u when G CTAs split U units as [floor(c*U/G), floor((c+1)*U/G))
__device__ __forceinline__ int unit_owner(int u, int G, int U) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * G + U - 1) / U) - 1;
}

__device__ __forceinline__ void bar_sync_consumers() {
  asm volatile("bar.sync 1, %0;" :: "n"(SK_CTHREADS) : "memory");
}

// POST kernels always contract in fp16 (bf16 activations are converted exactly, see stage_x), so the
// W4 nibbles can be used in place for both dtypes, see file header
template <int NBITS> struct UseLoHi {
  static constexpr bool value = (NBITS == 4);
};
// w >> 8 on the FMA pipe (IMAD.HI): the ALU pipe (LOP3/SHF, 64 lanes/clk/SM) is the scarce unit here
__device__ __forceinline__ uint32_t shr8_fma(uint32_t w) {
  uint32_t d;
  asm("mul.hi.u32 %0, %1, 16777216;" : "=r"(d) : "r"(w));
  return d;
}
__device__ __forceinline__ uint32_t shr4_fma(uint32_t w) {
  uint32_t d;
  asm("mul.hi.u32 %0, %1, 268435456;" : "=r"(d) : "r"(w));
  return d;
}
// D += A(16 x 32, unsigned bytes, row) * B(32 x 8, signed bytes, col), int32 accumulate
__device__ __forceinline__ void imma_16832(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A(16 x 32, unsigned bytes, row) * B(32 x 8, UNSIGNED bytes, col): weight codes x activation codes (MODE 3)
__device__ __forceinline__ void imma_16832_uu(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// act_fn(gate) * up for eight elements: F.silu in fp32 stored as T, then the product rounded to T -- the arithmetic
// of k6_silu_mul (q100_block.cu), i.e. of models/int_llama_layer.py:45 on 16-bit tensors
template <typename T>
__device__ __forceinline__ uint4 silu_mul8_inl(const uint4& g, const uint4& u) {
  using D = DT<T>;
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
  uint32_t ow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = D::to_f2(gw[j]), b = D::to_f2(uw[j]);
    const float sx = D::rd(__fdiv_rn(a.x, __fadd_rn(1.0f, expf(-a.x))));
    const float sy = D::rd(__fdiv_rn(a.y, __fadd_rn(1.0f, expf(-a.y))));
    ow[j] = D::pack2(__fmul_rn(sx, b.x), __fmul_rn(sy, b.y));
  }
  return make_uint4(ow[0], ow[1], ow[2], ow[3]);
}
template <typename T>
__device__ __noinline__ uint4 silu_mul8(const uint4& g, const uint4& u) {   // out of line: keeps its temporaries out of the kernel's 72-register budget
  return silu_mul8_inl<T>(g, u);
}
// weight * (h * rstd) for eight elements: both products in fp32, one rounding to T (k4_add_rmsnorm, q100_block.cu)
template <typename T>
__device__ __forceinline__ uint4 norm8(const uint4& h, const uint4& w, float rstd) {
  using D = DT<T>;
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, ww[4] = {w.x, w.y, w.z, w.w};
  uint32_t ow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = D::to_f2(hw[j]), g = D::to_f2(ww[j]);
    ow[j] = D::pack2(__fmul_rn(g.x, __fmul_rn(f.x, rstd)), __fmul_rn(g.y, __fmul_rn(f.y, rstd)));
  }
  return make_uint4(ow[0], ow[1], ow[2], ow[3]);
}
// x_attn staging: elements 8 ch .. 8 ch + 7 of the attention output of (sequence m, head) -- the flash-decoding partials
// of the position chunks up to the new token merged (fp32) and rounded once to T
template <typename T>
__device__ __forceinline__ uint4 attn_merge8(const SkinnyArgs& a, const int64_t p, int m, int head, int ch) {
  using D = DT<T>;
  int n_parts = (p < 0) ? 0 : static_cast<int>((p + a.att_chunk_pos) / a.att_chunk_pos);
  if (n_parts > a.att_max_parts) n_parts = 0;               // position outside the cache: the attention phase raised err
  const int H = a.K / AT_D;
  const float* base = a.att_part + (static_cast<int64_t>(m) * H + head) * a.att_max_parts * AT_PART_STRIDE;
  float M = neg_inf_f(), L = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i0 = 0; i0 < n_parts; i0 += 4) {
    float2 ml[4];
    float4 oa[4], ob[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u < n_parts ? i0 + u : n_parts - 1;
      const float* src = base + i * AT_PART_STRIDE;
      ml[u] = __ldcg(reinterpret_cast<const float2*>(src));
      oa[u] = __ldcg(reinterpret_cast<const float4*>(src + 4 + 8 * ch));
      ob[u] = __ldcg(reinterpret_cast<const float4*>(src + 8 + 8 * ch));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u < n_parts && ml[u].y > 0.f) {
        const float Mn = fmaxf(M, ml[u].x);
        const float c0 = expf(M - Mn), c1 = expf(ml[u].x - Mn);
        L = fmaf(L, c0, ml[u].y * c1);
        acc[0] = fmaf(acc[0], c0, oa[u].x * c1); acc[1] = fmaf(acc[1], c0, oa[u].y * c1);
        acc[2] = fmaf(acc[2], c0, oa[u].z * c1); acc[3] = fmaf(acc[3], c0, oa[u].w * c1);
        acc[4] = fmaf(acc[4], c0, ob[u].x * c1); acc[5] = fmaf(acc[5], c0, ob[u].y * c1);
        acc[6] = fmaf(acc[6], c0, ob[u].z * c1); acc[7] = fmaf(acc[7], c0, ob[u].w * c1);
        M = Mn;
      }
    }
  }
  const float inv = L > 0.f ? 1.0f / L : 0.f;
  return make_uint4(D::pack2(acc[0] * inv, acc[1] * inv), D::pack2(acc[2] * inv, acc[3] * inv),
                    D::pack2(acc[4] * inv, acc[5] * inv), D::pack2(acc[6] * inv, acc[7] * inv));
}
// mbarrier wait without bookkeeping instructions in the retry path (try_wait suspends in hardware)
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_1d_hint_u32(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_tight_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}

// 8-byte accesses of the stream-K fix-up slots: .volatile keeps them out of L1 and un-merged; an aligned
// 8-byte access is single-copy atomic, so {value, flag} is observed whole or not at all
#ifdef SK_LL_GPU_SCOPE
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u2(uint2* p, uint2 v) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
#else
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
#endif

#ifdef SK_TIMING
// debug build: per-CTA cycle breakdown (consumer warp 0 / producer), read back by q100_debug_skinny_timing
__device__ long long g_sk_t[2048][8];
// per-launch wall-clock stamps (globaltimer, ns): [launch % 16][cta][entry, after pdl_wait, before resolve, end]
__device__ long long g_sk_w[16][512][4];
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SK_T(var) const long long var = clock64()
#else
#define SK_T(var)
#endif

// ---------------------------------------------------------------- kernel
// MODE 0: 16-bit fake-quantised weights rebuilt bit-exactly (any group size);  MODE 1 ("POST"): fp16 magic-form
// operand + fp32 group step;  MODE 2 ("IMMA", W4, M <= 8): the codes go to the tensor core as unsigned bytes
// (two LOP3 + one shift per eight codes instead of five ALU ops) against x in block fixed point -- three signed
// 8-bit limbs per (token, 128-k block), one mma column each, MB = ceil(3 M / 8) column groups.  The integer
// contraction is exact; x is rounded once to 2^-22 of its block maximum (below fp32 accumulation noise).
struct RingPos { uint32_t soff, boff; int sleft; uint32_t ph; };   // consumer position in the weight ring
__device__ __forceinline__ void chain_stamp(unsigned long long* row, int i) {
  if (row != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    row[i] = t;
  }
}

// chain mode: fixed places at the start of the dynamic shared memory, so that the GEMM share of a phase needs no
// arguments at all (see consumer_gemm_call): the phase being executed, 1 / rms of each token (x_norm), the consumers'
// ring position (uniform over the consumer threads; carried from one GEMM phase to the next) and the ring's mbarriers
constexpr size_t CH_TRACE_OFF = 8192;                             // descriptor header: barrier stamps [phases][3]
constexpr size_t CH_CTA_TRACE_OFF = 8192 + 24 * CH_MAX_PHASES;    // ... and [512 CTAs][8] stamps of the traced phase
constexpr int CH_SM_PHASE = 0;
constexpr int CH_SM_RSTD = 512;
constexpr int CH_SM_RPOS = 640;
constexpr int CH_SM_POS = 656;                       // int64: the device-side position, read once per launch
constexpr int CH_SM_TRACE = 704;                     // development: global pointer to this CTA's stamp row, or null
constexpr int CH_SM_FULL = 768;
constexpr int CH_SM_EMPTY = 768 + 8 * SK_MAX_STAGES;
constexpr int CH_SM_ATT = 1024;                      // attention phase: 8 warps x 4 heads x AT_PART_STRIDE floats
constexpr int CH_SM_NEXT = 1024 + 8 * 4 * AT_PART_STRIDE * 4;     // the next phase's arguments, prefetched with cp.async
constexpr int CH_SM_RING = CH_SM_NEXT + CH_PHASE_BYTES;
static_assert(CH_SM_RING % 128 == 0, "ring must stay 128-byte aligned");
constexpr size_t CH_POS_SLOT_OFF = 4128;                          // descriptor header: pointer to the position scalar

// The consumer warps' share of ONE dequant-GEMM: stage x, pull landed blocks out of the ring, contract, write y or
// hand the fragment to the stream-K fix-up.  Inlined into k3_skinny.  k3_chain runs it once per GEMM phase; with three
// CTAs per SM (72 registers) it is a function of its own there (consumer_gemm_call) that finds everything it needs in
// shared memory: inlined next to the phase loop and the attention code, ptxas kept the ring position and the per-lane
// operand addresses in local memory inside the block loop (26 LDL/STL per block; the loop ran at half speed).
template <typename T, int NBITS, int MB, int MODE, bool CHAIN>
__device__ __forceinline__ void consumer_gemm(const SkinnyArgs& a_in, const uint32_t full_in, const uint32_t empty_in) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const SkinnyArgs* ap = &a_in;
  if constexpr (CHAIN) ap = &reinterpret_cast<const ChainPhase*>(smem_raw + CH_SM_PHASE)->g;
  const SkinnyArgs& a = *ap;
  const int G = CHAIN ? reinterpret_cast<const ChainPhase*>(smem_raw + CH_SM_PHASE)->grid_p : static_cast<int>(gridDim.x);
  const uint32_t ring_u32 = CHAIN ? smem_u32(smem_raw) + CH_SM_RING : ((smem_u32(smem_raw) + 127u) & ~127u);
  uint8_t* ring_ptr = smem_raw + (ring_u32 - smem_u32(smem_raw));
  const uint32_t full_base = CHAIN ? smem_u32(smem_raw) + CH_SM_FULL : full_in;
  const uint32_t empty_base = CHAIN ? smem_u32(smem_raw) + CH_SM_EMPTY : empty_in;
  const float* s_rstd = reinterpret_cast<const float*>(smem_raw + CH_SM_RSTD);
  RingPos rp{0u, 0u, a.stages, 0u};
  if constexpr (CHAIN) rp = *reinterpret_cast<const RingPos*>(smem_raw + CH_SM_RPOS);
  using T2 = typename DQ<T>::T2;
  constexpr bool POST = MODE == 1;
  constexpr bool IMMA = MODE == 2;
  constexpr bool A8 = MODE == 3;                                 // activation codes x weight codes (q100_gemm_wna8)
  constexpr bool kBf16 = MODE == 0 && DT<T>::kCode == Q100_BF16;  // MMA operand type (POST: always fp16)
  constexpr bool kConvX = POST && DT<T>::kCode == Q100_BF16;    // bf16 activations -> scaled fp16 at staging
  constexpr bool LOHI = POST && UseLoHi<NBITS>::value;
  using DQM = DQ<__half>;                                        // magic constants of the POST path
  constexpr int FRAG = SK_ROWS * MB * 8;              // floats per partial tile
  const int stage_pitch = (stage_bytes<NBITS>(a.nsub) + 127) & ~127;
  const int stages = a.stages;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int U = a.total_units;
  const int q_begin = static_cast<int>(static_cast<int64_t>(cta) * U / G);
  const int q_end = static_cast<int>(static_cast<int64_t>(cta + 1) * U / G);
  const int kblocks = a.kblocks;
  // ===================== consumers =====================
  const uint32_t xs_u32 = ring_u32 + stages * stage_pitch;
  uint8_t* xs_ptr = ring_ptr + stages * stage_pitch;
  const int mrows = a.m_rows;
  constexpr int kXPitch = IMMA ? SK_IPITCH : A8 ? SK_ICOL : SK_XPITCH;
  const int blk_pitch = (A8 ? mrows + 1 : mrows) * kXPitch;   // bytes of x per staged k block (A8: + the all-ones column)
  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const bool has_zp = a.has_zp != 0;

  // per-lane offsets inside a stage
  const int coff = LaneCodes<NBITS>::lane_off(r0, t);
  int coff_hi = 0;                                    // W3 only: 1-bit plane
  if constexpr (NBITS == 3) coff_hi = LaneCodes<3>::lane_off_hi(r0, t);
  const int msub = a.group >= SK_KB ? 0 : (a.group == 32 ? t : (t >> 1));
  const int moff = WGeom<NBITS>::kCodeBytes + msub * 512 + r0 * 4;
  // opaque copies: keeps these per-lane addresses in registers (ptxas otherwise recomputes them from
  // %tid in every iteration, on the ALU pipe that the extraction needs)
  const uint32_t cbase = reg_const(ring_u32 + coff), chbase = reg_const(ring_u32 + coff_hi);
  const uint32_t mbase_l = reg_const(ring_u32 + moff);
  const uint32_t full_u32 = reg_const(full_base), empty_u32 = reg_const(empty_base);

  // x slab of local blocks [q0, q0 + n): rows m < M only.  The 16-byte chunk c = 4t + j of a block is
  // stored at offset 16t + 64j (conflict-free LDS.128 for the mma B fragments); LOHI additionally swaps
  // the two middle words so that one LDS.128 yields (k0,k1 | k4,k5 | k2,k3 | k6,k7).  The block sums
  // (sum of the "plain" pairs, sum of the "x16" pairs) of row (bl, m) go to the row's padding.
  auto stage_x = [&](int q0, int n) {
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int per_blk = (A8 ? mrows + 1 : mrows) * 16;
    const int total = n * per_blk;                         // 16-byte chunks
    const int kblk0 = q0 % kblocks;
    // XU > 1 (loads of several rounds in flight) and a cp.async variant (raw chunks global -> shared, transformed in
    // place) were both measured: no gain -- the M dependence of this kernel (gate|up 16.7 us at M = 1, 26.5 us at
    // M = 8) is shared-memory bandwidth of the x fragments (every warp re-reads the block's x), not this loop.
    constexpr int XU = 1;
    for (int c0 = tid; c0 < total; c0 += SK_CTHREADS * XU) {
      uint4 vq[XU];
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        const int c = c0 + u * SK_CTHREADS;
        if (c < total) {
          const int bl = c / per_blk;
          const int rem = c - bl * per_blk;
          const int kblk = (kblk0 + bl) % kblocks;
          if constexpr (A8) {
            if ((rem >> 4) < mrows)
              vq[u] = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(rem >> 4) * a.K + kblk * SK_KB + (rem & 15) * 8);
          } else if constexpr (CHAIN) {
            // x was written by other SMs earlier in this launch: read it through L2 (ld.global.cg), never L1
            if (a.x_gate_up) {
              const T* gp = xg + static_cast<int64_t>(rem >> 4) * (2 * a.K) + kblk * SK_KB + (rem & 15) * 8;
              vq[u] = silu_mul8_inl<T>(__ldcg(reinterpret_cast<const uint4*>(gp)), __ldcg(reinterpret_cast<const uint4*>(gp + a.K)));
            } else if (a.x_attn) {
              vq[u] = attn_merge8<T>(a, *reinterpret_cast<const int64_t*>(smem_raw + CH_SM_POS), rem >> 4, kblk,
                                     rem & 15);                                 // k block == head (head_dim 128)
            } else {
              const int kofs = kblk * SK_KB + (rem & 15) * 8;
              vq[u] = __ldcg(reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(rem >> 4) * a.K + kofs));
              if (a.x_norm)       // weight * (h * rstd), both products in fp32, one rounding: K4 / omni_norm.py:52-63
                vq[u] = norm8<T>(vq[u], *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.norm_w) + kofs),
                                 s_rstd[rem >> 4]);
            }
          } else if (a.x_gate_up) {
            const T* gp = xg + static_cast<int64_t>(rem >> 4) * (2 * a.K) + kblk * SK_KB + (rem & 15) * 8;
            vq[u] = silu_mul8<T>(*reinterpret_cast<const uint4*>(gp), *reinterpret_cast<const uint4*>(gp + a.K));
          } else {
            vq[u] = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(rem >> 4) * a.K + kblk * SK_KB + (rem & 15) * 8);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < XU; ++u) {
      const int c = c0 + u * SK_CTHREADS;
      if (c >= total) break;                                // uniform per half-warp: total is a multiple of 16
      const int bl = c / per_blk;
      const int rem = c - bl * per_blk;
      const int m = rem >> 4, ch = rem & 15;
      uint4 v = vq[u];
      if constexpr (A8) {
        // per-token activation quantisation (quantizer.py:140-159 on the row's min / max, then :103-106) right here:
        // the codes are the reference's, bit for bit; they go to the tensor core as unsigned bytes in the byte order
        // in which the weight codes come out of a packed word.  Column m = mrows is all ones: its product with the
        // weight codes is sum_k q_w, the term the activation zero-point multiplies.
        uint8_t* colp = xs_ptr + bl * blk_pitch + m * SK_ICOL;
        const int boff = (ch >> 2) * 32 + (ch & 3) * 8;
        if (m == mrows) {
          *reinterpret_cast<uint2*>(colp + boff) = make_uint2(0x01010101u, 0x01010101u);
          continue;
        }
        QuantArgs qa{};
        qa.denom = a.a_denom; qa.rcp_denom = 1.0f / a.a_denom; qa.has_zp = 1; qa.cuda_scale = a.a_cuda_scale;
        const T* mmp = reinterpret_cast<const T*>(a.x_minmax) + 2 * m;
        const GroupQ<T> q = group_params<T>(DT<T>::to_f(mmp[0]), DT<T>::to_f(mmp[1]), qa, 0);
        const float2 p0 = DT<T>::to_f2(v.x), p1 = DT<T>::to_f2(v.y), p2 = DT<T>::to_f2(v.z), p3 = DT<T>::to_f2(v.w);
        const float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
        uint32_t cq[8];
        float sx = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          using D = DT<T>;
          const float vv = D::rd(__fdiv_rn(f[i], q.scale));
          float xi = D::rd(D::rd(rintf(vv) - vv) + vv);
          xi = D::rd(xi + q.zp);
          const float c = clamp_nan(xi, 0.f, a.a_qmax);
          cq[i] = static_cast<uint32_t>(__float2int_rn(c));
          sx += static_cast<float>(cq[i]);
        }
        const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // the 16 chunks of one (block, m)
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) sx += __shfl_xor_sync(half_mask, sx, d);
        auto pk = [](uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) { return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24); };
        if constexpr (NBITS == 4)
          *reinterpret_cast<uint2*>(colp + boff) = make_uint2(pk(cq[0], cq[4], cq[1], cq[5]), pk(cq[2], cq[6], cq[3], cq[7]));
        else
          *reinterpret_cast<uint2*>(colp + boff) = make_uint2(pk(cq[0], cq[2], cq[1], cq[3]), pk(cq[4], cq[6], cq[5], cq[7]));
        if (ch == 0) *reinterpret_cast<float4*>(colp + 128) = make_float4(sx, q.scale, q.zp, 0.f);
        continue;
      }
      if constexpr (IMMA) {
        // block fixed point: xi = rint(x * 2^(21 - e)) with e the exponent of the (token, block) maximum, so
        // |xi| < 2^22, split into balanced base-256 digits l0 + 256 l1 + 65536 l2 (each in [-128, 127]).
        // Chunk ch holds k = 8 ch .. 8 ch + 7 = lane quarter t = ch >> 2, word j = ch & 3; its bytes are stored
        // in the order (e0, e4, e1, e5 | e2, e6, e3, e7) in which the weight nibbles come out of a packed word.
        const float2 p0 = DT<T>::to_f2(v.x), p1 = DT<T>::to_f2(v.y), p2 = DT<T>::to_f2(v.z), p3 = DT<T>::to_f2(v.w);
        const float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
        const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // 16 chunks of one (block, m)
        uint32_t ab = 0;                                   // max |x| as bits: finite < inf < NaN
#pragma unroll
        for (int i = 0; i < 8; ++i) ab = max(ab, __float_as_uint(f[i]) & 0x7fffffffu);
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) ab = max(ab, __shfl_xor_sync(half_mask, ab, d));
        const int e = static_cast<int>(ab >> 23);
        const int fe = min(max(275 - e, 1), 250);          // scale = 2^(fe - 127) = 2^(21 - (e - 127))
        const float scale = __uint_as_float(static_cast<uint32_t>(fe) << 23);
        int l0[8], l1[8], l2[8];
        int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int xi = __float2int_rn(f[i] * scale);
          l0[i] = (xi << 24) >> 24;
          const int x1 = (xi - l0[i]) >> 8;
          l1[i] = (x1 << 24) >> 24;
          l2[i] = (x1 - l1[i]) >> 8;
          s0 += l0[i]; s1 += l1[i]; s2 += l2[i];
        }
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          s0 += __shfl_xor_sync(half_mask, s0, d);
          s1 += __shfl_xor_sync(half_mask, s1, d);
          s2 += __shfl_xor_sync(half_mask, s2, d);
        }
        auto pk = [](int b0, int b1, int b2, int b3) {
          return (static_cast<uint32_t>(b0) & 0xffu) | ((static_cast<uint32_t>(b1) & 0xffu) << 8) |
                 ((static_cast<uint32_t>(b2) & 0xffu) << 16) | (static_cast<uint32_t>(b3) << 24);
        };
        uint8_t* colp = xs_ptr + bl * blk_pitch + (3 * m) * SK_ICOL;
        if constexpr (NBITS == 4) {
          const int boff = (ch >> 2) * 32 + (ch & 3) * 8;
          *reinterpret_cast<uint2*>(colp + boff) = make_uint2(pk(l0[0], l0[4], l0[1], l0[5]), pk(l0[2], l0[6], l0[3], l0[7]));
          *reinterpret_cast<uint2*>(colp + SK_ICOL + boff) = make_uint2(pk(l1[0], l1[4], l1[1], l1[5]), pk(l1[2], l1[6], l1[3], l1[7]));
          *reinterpret_cast<uint2*>(colp + 2 * SK_ICOL + boff) = make_uint2(pk(l2[0], l2[4], l2[1], l2[5]), pk(l2[2], l2[6], l2[3], l2[7]));
        } else {
          // W2: a packed word holds 16 codes; (w >> 2 jr) & 0x03030303 yields elements (2jr, 2jr+8, 2jr+1, 2jr+9) of
          // the word, registers jr = 0..3; mma m = 2 * word + (jr >> 1) takes register jr even as a0 and jr odd as a2.
          // Element i of chunk ch is element e = 8 (ch & 1) + i of word (ch & 3) >> 1 of lane quarter t = ch >> 2.
          const int t4 = (ch >> 2) * 32, wi = (ch & 3) >> 1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = 8 * (ch & 1) + i, e7 = e & 7, jr = e7 >> 1;
            const int off = t4 + (wi * 2 + (jr >> 1)) * 8 + (jr & 1) * 4 + (e >> 3) + (e7 & 1) * 2;
            colp[off] = static_cast<uint8_t>(l0[i]);
            colp[SK_ICOL + off] = static_cast<uint8_t>(l1[i]);
            colp[2 * SK_ICOL + off] = static_cast<uint8_t>(l2[i]);
          }
        }
        if (ch == 0) {
          // column coefficient 256^l / scale; inf / NaN anywhere in the block poisons the token's result
          const uint32_t cexp = static_cast<uint32_t>(254 - fe);
          const bool bad = e == 255;
          const float c0 = bad ? nan_f() : __uint_as_float(cexp << 23);
          const float c1 = bad ? nan_f() : __uint_as_float((cexp + 8) << 23);
          const float c2 = bad ? nan_f() : __uint_as_float((cexp + 16) << 23);
          *reinterpret_cast<float2*>(colp + 128) = make_float2(static_cast<float>(s0), c0);
          *reinterpret_cast<float2*>(colp + SK_ICOL + 128) = make_float2(static_cast<float>(s1), c1);
          *reinterpret_cast<float2*>(colp + 2 * SK_ICOL + 128) = make_float2(static_cast<float>(s2), c2);
        }
        continue;
      }
      if constexpr (POST) {
        float2 f0 = DT<T>::to_f2(v.x), f1 = DT<T>::to_f2(v.y), f2 = DT<T>::to_f2(v.z), f3 = DT<T>::to_f2(v.w);
        // the 16 chunks of one (block, m) sit in 16 consecutive threads
        const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // total is a multiple of 16
        float inv = 1.0f;
        if constexpr (kConvX) {
          // bf16 -> fp16 without loss: scale the (token, block) slice by a power of two that puts its
          // largest magnitude in [2^14, 2^15); every element down to 2^-31 of that maximum is then an
          // exact fp16 number (8 significant bits), smaller ones round at 2^-40 of it -- far below what
          // the fp32 accumulation resolves.  The inverse power goes into the fp32 post step.
          float am = fmaxf(fmaxf(fmaxf(fabsf(f0.x), fabsf(f0.y)), fmaxf(fabsf(f1.x), fabsf(f1.y))),
                           fmaxf(fmaxf(fabsf(f2.x), fabsf(f2.y)), fmaxf(fabsf(f3.x), fabsf(f3.y))));
#pragma unroll
          for (int d = 1; d < 16; d <<= 1) am = fmaxf(am, __shfl_xor_sync(half_mask, am, d));
          const int e = static_cast<int>((__float_as_uint(am) >> 23) & 0xffu);       // biased exponent
          int fe = 268 - e;                                                          // 2^(14 - (e - 127))
          fe = (e == 0 || e == 255) ? 127 : min(max(fe, 1), 253);                    // zero / inf / nan: 1.0
          const float f = __uint_as_float(static_cast<uint32_t>(fe) << 23);
          inv = __uint_as_float(static_cast<uint32_t>(254 - fe) << 23);
          f0.x *= f; f0.y *= f; f1.x *= f; f1.y *= f; f2.x *= f; f2.y *= f; f3.x *= f; f3.y *= f;
          v.x = DT<__half>::pack2(f0.x, f0.y); v.y = DT<__half>::pack2(f1.x, f1.y);
          v.z = DT<__half>::pack2(f2.x, f2.y); v.w = DT<__half>::pack2(f3.x, f3.y);
          if (e != 0 && e != 255) {      // sums of what the tensor core will see (exact conversions)
            f0 = DT<__half>::to_f2(v.x); f1 = DT<__half>::to_f2(v.y); f2 = DT<__half>::to_f2(v.z); f3 = DT<__half>::to_f2(v.w);
          }
        }
        float s_a, s_b;
        if (LOHI) { s_a = (f0.x + f0.y) + (f2.x + f2.y); s_b = (f1.x + f1.y) + (f3.x + f3.y); }
        else { s_a = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y)); s_b = 0.f; }
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          s_a += __shfl_xor_sync(half_mask, s_a, d);
          if (LOHI) s_b += __shfl_xor_sync(half_mask, s_b, d);
        }
        if (ch == 0)
          // slot (m >> 1) & 3 of the row's 64 bytes of padding: the four lane quarters of a warp read the sums of rows
          // 2t / 2t + 1, which would otherwise sit 640 bytes apart = in the same 16-byte bank group (4-way conflict,
          // measured: 10x the bank conflicts of M = 1 at M = 8)
          *reinterpret_cast<float4*>(xs_ptr + bl * blk_pitch + m * SK_XPITCH + 256 + ((m >> 1) & 3) * 16) =
              make_float4(s_a, s_b, inv, 0.f);
      }
      if (LOHI) { const uint32_t y = v.y; v.y = v.z; v.z = y; }
      *reinterpret_cast<uint4*>(xs_ptr + bl * blk_pitch + m * SK_XPITCH + (ch >> 2) * 16 + (ch & 3) * 64) = v;
      }
    }
  };

  // per-lane x addresses: column n = mb*8+g of the mma B operand is token m; columns >= M read the last
  // valid row instead (their results are never stored), so no zero fill and no branch is needed
  uint32_t xrow[MB], srow[MB], srow1[(IMMA || A8) ? MB : 1];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    if constexpr (IMMA || A8) {
      // IMMA: mma column n = mb*8 + g is (token, limb) = divmod(n, 3); A8: column n is token n, column mrows the all-ones
      // column.  This lane's output columns are n = mb*8 + 2t, +1
      const int ncols = A8 ? mrows + 1 : 3 * mrows;
      xrow[mb] = xs_u32 + min(mb * 8 + g, ncols - 1) * SK_ICOL + t * 32;
      srow[mb] = xs_u32 + min(mb * 8 + 2 * t, ncols - 1) * SK_ICOL + 128;
      srow1[mb] = xs_u32 + min(mb * 8 + 2 * t + 1, ncols - 1) * SK_ICOL + 128;
    } else {
      xrow[mb] = xs_u32 + min(mb * 8 + g, mrows - 1) * SK_XPITCH + t * 16;
      // this lane's two output columns need the block sums of tokens m = mb*8 + 2t, 2t+1 (clamped like x)
      const int srow_m = min(mb * 8 + 2 * t, mrows - 1);
      srow[mb] = xs_u32 + srow_m * SK_XPITCH + 256 + ((srow_m >> 1) & 3) * 16;
    }
  }

  float acc[MB][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }

  const uint32_t qb = DT<T>::pack2(a.qmin, a.qmin);
  const T2 qmin2 = *reinterpret_cast<const T2*>(&qb);
  const float zp_fix = -a.qmin;                       // effective zero-point when the quantizer has none
  const uint32_t magic = reg_const(DQM::kMagic);

  auto compute_block = [&](const BlockRegs<NBITS>& b, uint32_t xoff) {
    if constexpr (A8) {
      if constexpr (NBITS == 4 || NBITS == 8) {
        int di[MB][4];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) { di[mb][0] = di[mb][1] = di[mb][2] = di[mb][3] = 0; }
        uint4 xb[MB][2];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) { xb[mb][0] = lds128(xrow[mb] + xoff); xb[mb][1] = lds128(xrow[mb] + xoff + 16); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t af[4];
          if constexpr (NBITS == 4) {
            const uint32_t w0 = b.c0.word(j), w1 = b.c1.word(j);
            af[0] = w0 & 0x0f0f0f0fu;
            af[1] = w1 & 0x0f0f0f0fu;
            af[2] = shr4_fma(w0) & 0x0f0f0f0fu;
            af[3] = shr4_fma(w1) & 0x0f0f0f0fu;
          } else {                                          // W8: the packed words are the byte operands
            af[0] = b.c0.wsel(2 * j); af[1] = b.c1.wsel(2 * j);
            af[2] = b.c0.wsel(2 * j + 1); af[3] = b.c1.wsel(2 * j + 1);
          }
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint4& q4 = xb[mb][j >> 1];
            imma_16832_uu(di[mb], af, (j & 1) ? q4.z : q4.x, (j & 1) ? q4.w : q4.y);
          }
        }
        // sum_k q_w of rows r0 / r1: the all-ones column mrows, held by lane quarter (mrows & 7) >> 1
        const int cs = mrows;
        int sw0i = (cs & 1) ? di[0][1] : di[0][0], sw1i = (cs & 1) ? di[0][3] : di[0][2];
        if constexpr (MB > 1) {
          if ((cs >> 3) == 1) { sw0i = (cs & 1) ? di[1][1] : di[1][0]; sw1i = (cs & 1) ? di[1][3] : di[1][2]; }
        }
        const int src = (lane & ~3) | ((cs & 7) >> 1);
        const float sw0 = static_cast<float>(__shfl_sync(0xffffffffu, sw0i, src));
        const float sw1 = static_cast<float>(__shfl_sync(0xffffffffu, sw1i, src));
        const float2 sz0 = DT<T>::to_f2(b.m0), sz1 = DT<T>::to_f2(b.m1);     // (scale, zero) of rows r0, r1
        const float zw0 = has_zp ? sz0.y : zp_fix, zw1 = has_zp ? sz1.y : zp_fix;
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const uint4 u0 = lds128(srow[mb] + xoff), u1 = lds128(srow1[mb] + xoff);   // (sum of codes, scale, zero-point)
          const float sx0 = __uint_as_float(u0.x), ax0 = __uint_as_float(u0.y), zx0 = __uint_as_float(u0.z);
          const float sx1 = __uint_as_float(u1.x), ax1 = __uint_as_float(u1.y), zx1 = __uint_as_float(u1.z);
          // sum_k (q_w - z_w)(q_x - z_x) = D - z_x sum q_w - z_w sum q_x + 128 z_w z_x: integers, exact in fp32
          const float v0 = fmaf(zw0, fmaf(128.f, zx0, -sx0), fmaf(-zx0, sw0, static_cast<float>(di[mb][0])));
          const float v1 = fmaf(zw0, fmaf(128.f, zx1, -sx1), fmaf(-zx1, sw0, static_cast<float>(di[mb][1])));
          const float v2 = fmaf(zw1, fmaf(128.f, zx0, -sx0), fmaf(-zx0, sw1, static_cast<float>(di[mb][2])));
          const float v3 = fmaf(zw1, fmaf(128.f, zx1, -sx1), fmaf(-zx1, sw1, static_cast<float>(di[mb][3])));
          acc[mb][0] = fmaf(sz0.x * ax0, v0, acc[mb][0]);
          acc[mb][1] = fmaf(sz0.x * ax1, v1, acc[mb][1]);
          acc[mb][2] = fmaf(sz1.x * ax0, v2, acc[mb][2]);
          acc[mb][3] = fmaf(sz1.x * ax1, v3, acc[mb][3]);
        }
      }
    } else if constexpr (IMMA) {
      int di[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) { di[mb][0] = di[mb][1] = di[mb][2] = di[mb][3] = 0; }
      uint4 xb[MB][2];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) { xb[mb][0] = lds128(xrow[mb] + xoff); xb[mb][1] = lds128(xrow[mb] + xoff + 16); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t af[4];
        if constexpr (NBITS == 4) {
          const uint32_t w0 = b.c0.word(j), w1 = b.c1.word(j);
          af[0] = w0 & 0x0f0f0f0fu;                       // (e0, e4, e1, e5) of row r0
          af[1] = w1 & 0x0f0f0f0fu;                       //                  of row r1
          af[2] = shr4_fma(w0) & 0x0f0f0f0fu;             // (e2, e6, e3, e7)
          af[3] = shr4_fma(w1) & 0x0f0f0f0fu;
        } else {                                          // W2: word j >> 1, registers jr = 2 (j & 1), + 1
          const uint32_t w0 = (j >> 1) ? b.c0.w.y : b.c0.w.x, w1 = (j >> 1) ? b.c1.w.y : b.c1.w.x;
          constexpr uint32_t m2 = 0x03030303u;
          af[0] = ((j & 1) ? (w0 >> 4) : w0) & m2;
          af[1] = ((j & 1) ? (w1 >> 4) : w1) & m2;
          af[2] = ((j & 1) ? (w0 >> 6) : (w0 >> 2)) & m2;
          af[3] = ((j & 1) ? (w1 >> 6) : (w1 >> 2)) & m2;
        }
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const uint4& q4 = xb[mb][j >> 1];
          imma_16832(di[mb], af, (j & 1) ? q4.z : q4.x, (j & 1) ? q4.w : q4.y);
        }
      }
      const float2 sz0 = DT<T>::to_f2(b.m0), sz1 = DT<T>::to_f2(b.m1);     // (scale, zero) of rows r0, r1
      const float nz0 = has_zp ? -sz0.y : -zp_fix, nz1 = has_zp ? -sz1.y : -zp_fix;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const float2 u0 = lds64f(srow[mb] + xoff), u1 = lds64f(srow1[mb] + xoff);   // (sum of limbs, coefficient)
        const float v0 = fmaf(nz0, u0.x, static_cast<float>(di[mb][0]));
        const float v1 = fmaf(nz0, u1.x, static_cast<float>(di[mb][1]));
        const float v2 = fmaf(nz1, u0.x, static_cast<float>(di[mb][2]));
        const float v3 = fmaf(nz1, u1.x, static_cast<float>(di[mb][3]));
        acc[mb][0] = fmaf(sz0.x * u0.y, v0, acc[mb][0]);
        acc[mb][1] = fmaf(sz0.x * u1.y, v1, acc[mb][1]);
        acc[mb][2] = fmaf(sz1.x * u0.y, v2, acc[mb][2]);
        acc[mb][3] = fmaf(sz1.x * u1.y, v3, acc[mb][3]);
      }
    } else if constexpr (POST) {
      float ag[MB][4], agh[LOHI ? MB : 1][4];
      if constexpr (LOHI) {
        // fp16 W4: words in place, pairs {0,2} (value 1024+q) and {1,3} (value 1024+16q) of each word
        const uint32_t w0[4] = {b.c0.word(0), b.c0.word(1), b.c0.word(2), b.c0.word(3)};
        const uint32_t w1[4] = {b.c1.word(0), b.c1.word(1), b.c1.word(2), b.c1.word(3)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 xb[MB];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) xb[mb] = lds128(xrow[mb] + xoff + i * 64);
          const uint32_t s0 = shr8_fma(w0[i]), s1 = shr8_fma(w1[i]);
          uint32_t al[4], ah[4];
#ifdef SK_EXP_NOEXTRACT
          al[0] = w0[i]; al[1] = w1[i]; al[2] = s0; al[3] = s1; ah[0] = w0[i] ^ magic; ah[1] = w1[i] ^ magic; ah[2] = s0 ^ magic; ah[3] = s1 ^ magic;
#else
          al[0] = lop_and_or<0x000f000fu>(w0[i], magic);
          al[1] = lop_and_or<0x000f000fu>(w1[i], magic);
          al[2] = lop_and_or<0x000f000fu>(s0, magic);
          al[3] = lop_and_or<0x000f000fu>(s1, magic);
          ah[0] = lop_and_or<0x00f000f0u>(w0[i], magic);
          ah[1] = lop_and_or<0x00f000f0u>(w1[i], magic);
          ah[2] = lop_and_or<0x00f000f0u>(s0, magic);
          ah[3] = lop_and_or<0x00f000f0u>(s1, magic);
#endif
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t bl2[2] = {xb[mb].x, xb[mb].y}, bh2[2] = {xb[mb].z, xb[mb].w};
#ifdef SK_EXP_NOMMA
            if (i == 0) { for (int r = 0; r < 4; ++r) { ag[mb][r] = __uint_as_float(al[r] ^ bl2[r & 1]); agh[mb][r] = __uint_as_float(ah[r] ^ bh2[r & 1]); } }
            else { for (int r = 0; r < 4; ++r) { ag[mb][r] = __uint_as_float(__float_as_uint(ag[mb][r]) ^ al[r] ^ bl2[r & 1]); agh[mb][r] = __uint_as_float(__float_as_uint(agh[mb][r]) ^ ah[r] ^ bh2[r & 1]); } }
#else
            if (i == 0) { mma_16816_zero(ag[mb], al, bl2, kBf16); mma_16816_zero(agh[mb], ah, bh2, kBf16); }
            else { mma_16816(ag[mb], al, bl2, kBf16); mma_16816(agh[mb], ah, bh2, kBf16); }
#endif
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 xb[MB];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) xb[mb] = lds128(xrow[mb] + xoff + j * 64);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int p = 4 * j + 2 * h;            // pair index of k = 32t + 8j + 4h + {0,1}
            uint32_t af[4];
            af[0] = b.c0.mpair(p, magic);
            af[1] = b.c1.mpair(p, magic);
            af[2] = b.c0.mpair(p + 1, magic);
            af[3] = b.c1.mpair(p + 1, magic);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const uint32_t bf[2] = {h == 0 ? xb[mb].x : xb[mb].z, h == 0 ? xb[mb].y : xb[mb].w};
              if (j == 0 && h == 0) mma_16816_zero(ag[mb], af, bf, kBf16);
              else mma_16816(ag[mb], af, bf, kBf16);
            }
          }
        }
      }
      const float2 sz0 = DT<T>::to_f2(b.m0), sz1 = DT<T>::to_f2(b.m1);     // (scale, zero) of rows r0, r1
      const float z0 = has_zp ? sz0.y : zp_fix, z1 = has_zp ? sz1.y : zp_fix;
      const float off0 = -(DQM::kMagicF + z0), off1 = -(DQM::kMagicF + z1);
      const float offh0 = -(DQM::kMagicF * 0.0625f + z0), offh1 = -(DQM::kMagicF * 0.0625f + z1);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        // (S_plain, S_x16, 2^-e, -) of tokens m = mb*8 + 2t and 2t + 1
        const uint4 u0 = lds128(srow[mb] + xoff), u1 = lds128(srow[mb] + xoff + SK_XPITCH);
        const float sa0 = __uint_as_float(u0.x), sb0 = __uint_as_float(u0.y);
        const float sa1 = __uint_as_float(u1.x), sb1 = __uint_as_float(u1.y);
        float v0 = ag[mb][0], v1 = ag[mb][1], v2 = ag[mb][2], v3 = ag[mb][3];
        if constexpr (LOHI) {
          v0 = fmaf(agh[mb][0], 0.0625f, v0); v1 = fmaf(agh[mb][1], 0.0625f, v1);
          v2 = fmaf(agh[mb][2], 0.0625f, v2); v3 = fmaf(agh[mb][3], 0.0625f, v3);
          v0 = fmaf(offh0, sb0, v0); v1 = fmaf(offh0, sb1, v1);
          v2 = fmaf(offh1, sb0, v2); v3 = fmaf(offh1, sb1, v3);
        }
        float w00 = sz0.x, w01 = sz0.x, w10 = sz1.x, w11 = sz1.x;
        if constexpr (kConvX) {
          const float i0 = __uint_as_float(u0.z), i1 = __uint_as_float(u1.z);
          w00 *= i0; w01 *= i1; w10 *= i0; w11 *= i1;
        }
        acc[mb][0] = fmaf(w00, fmaf(off0, sa0, v0), acc[mb][0]);
        acc[mb][1] = fmaf(w01, fmaf(off0, sa1, v1), acc[mb][1]);
        acc[mb][2] = fmaf(w10, fmaf(off1, sa0, v2), acc[mb][2]);
        acc[mb][3] = fmaf(w11, fmaf(off1, sa1, v3), acc[mb][3]);
      }
    } else {
      const uint32_t s0b = __byte_perm(b.m0, 0, 0x1010), z0b = __byte_perm(b.m0, 0, 0x3232);
      const uint32_t s1b = __byte_perm(b.m1, 0, 0x1010), z1b = __byte_perm(b.m1, 0, 0x3232);
      const T2 s0 = *reinterpret_cast<const T2*>(&s0b), z0 = *reinterpret_cast<const T2*>(&z0b);
      const T2 s1 = *reinterpret_cast<const T2*>(&s1b), z1 = *reinterpret_cast<const T2*>(&z1b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 xb[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) xb[mb] = lds128(xrow[mb] + xoff + j * 64);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = 4 * j + 2 * h;
          uint32_t af[4];
          af[0] = dq_pair<T, NBITS>(b.c0.pair(p), qmin2, z0, s0, has_zp);
          af[1] = dq_pair<T, NBITS>(b.c1.pair(p), qmin2, z1, s1, has_zp);
          af[2] = dq_pair<T, NBITS>(b.c0.pair(p + 1), qmin2, z0, s0, has_zp);
          af[3] = dq_pair<T, NBITS>(b.c1.pair(p + 1), qmin2, z1, s1, has_zp);
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t bf[2] = {h == 0 ? xb[mb].x : xb[mb].z, h == 0 ? xb[mb].y : xb[mb].w};
            mma_16816(acc[mb], af, bf, kBf16);
          }
        }
      }
    }
  };

  // y <- fragment `v` of this warp's 16 rows of tile n0 (mma D layout), + bias, converted to T
  auto write_fragment = [&](const float (&v)[MB][4], int n0) {
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    // chain epilogue (epi_add): h = r(r(acc) + res) is what gets stored, and the warp's 16 rows contribute one
    // partial sum of h^2 per token to the RMSNorm of the phase that reads h (fixed summation order: deterministic)
    bool epi_add = false;
    if constexpr (CHAIN) epi_add = a.epi_add != 0;
    auto add_res = [&](float o, int m, int n) {
      const float r = DT<T>::to_f(__ldcg(reinterpret_cast<const T*>(a.res) + static_cast<int64_t>(m) * a.N + n));
      return DT<T>::rd(__fadd_rn(r, DT<T>::rd(o)));
    };
    auto put_ssq = [&](float sq, int m, int n0_) {            // sq: this lane's share for token m; lanes with equal t hold the rows
      sq += __shfl_xor_sync(0xffffffffu, sq, 4);
      sq += __shfl_xor_sync(0xffffffffu, sq, 8);
      sq += __shfl_xor_sync(0xffffffffu, sq, 16);
      if (g == 0 && m < a.M) a.ssq_out[static_cast<int64_t>((n0_ >> 7) * SK_CWARPS + warp) * a.M + m] = sq;
    };
    if constexpr (IMMA) {
      // token m = sum of columns 3m, 3m+1, 3m+2 (the limb weights are already in the coefficients); column c
      // lives in lane quarter (c & 7) >> 1, register (c & 1) (+2 for row r1) of column group c >> 3
      float s0[(MB * 8 + 2) / 3], s1[(MB * 8 + 2) / 3];
#pragma unroll
      for (int i = 0; i < (MB * 8 + 2) / 3; ++i) s0[i] = s1[i] = 0.f;
#pragma unroll
      for (int c = 0; c < MB * 8; ++c) {
        const int src = (lane & ~3) | ((c & 7) >> 1);
        s0[c / 3] += __shfl_sync(0xffffffffu, v[c >> 3][c & 1], src);
        s1[c / 3] += __shfl_sync(0xffffffffu, v[c >> 3][2 + (c & 1)], src);
      }
#pragma unroll
      for (int m = 0; m < (MB * 8) / 3; ++m) {
        float sq = 0.f;
        if ((m & 3) == t && m < a.M) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = n0 + (h ? r1 : r0);
            if (n < a.N) {
              float o = h ? s1[m] : s0[m];
              if (bias) o += DT<T>::to_f(bias[n]);
              if constexpr (CHAIN) {
                if (epi_add) { o = add_res(o, m, n); sq = fmaf(o, o, sq); }
              }
              y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(o);
            }
          }
        }
        if constexpr (CHAIN) {
          if (epi_add) {                                     // token m lives in lane quarter m & 3: sum over the 8 row lanes
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            sq += __shfl_xor_sync(0xffffffffu, sq, 8);
            sq += __shfl_xor_sync(0xffffffffu, sq, 16);
            if (g == 0 && (m & 3) == t && m < a.M)
              a.ssq_out[static_cast<int64_t>((n0 >> 7) * SK_CWARPS + warp) * a.M + m] = sq;
          }
        }
      }
      return;
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      float sq[2] = {0.f, 0.f};                               // tokens mb*8 + 2t, + 1
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = n0 + ((r & 2) ? r1 : r0);
        const int m = mb * 8 + 2 * t + (r & 1);
        if (m < a.M && n < a.N) {
          float o = v[mb][r];
          if (bias) o += DT<T>::to_f(bias[n]);
          if constexpr (CHAIN) {
            if (epi_add) { o = add_res(o, m, n); sq[r & 1] = fmaf(o, o, sq[r & 1]); }
          }
          y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(o);
        }
      }
      if constexpr (CHAIN) {
        if (epi_add) { put_ssq(sq[0], mb * 8 + 2 * t, n0); put_ssq(sq[1], mb * 8 + 2 * t + 1, n0); }
      }
    }
  };

  // end of a tile segment [seg_begin, q]: write y or hand the fragment to the stream-K fix-up.
  // Fix-up protocol ("LL": data and flag travel in one 8-byte store, which is single-copy atomic, so no
  // fence, no atomic and no counter is needed): every contributor of a shared tile stores its fp32
  // fragment as {value, 1} pairs into its slot and moves on.  The CTA that owns the tile's LAST unit is the
  // reducer.  Units are processed in increasing order, so that CTA meets the shared tile at the START of
  // its run: it stores its own fragment like everybody else and defers the reduction to the END of its run,
  // when the other contributors (lower-numbered CTAs, which reach the tile at the end of THEIR runs) are
  // finishing too.  It then polls the slots in slot order (deterministic sum), writes y and resets the
  // flags, leaving the workspace zeroed for the next launch.  The reducer only ever waits for
  // lower-numbered CTAs, which the hardware dispatches first, so the wait cannot deadlock a partially
  // resident grid.
  int pend_tile = -1, pend_n = 0;
  auto slot_base = [&](int tile) {
    // slot layout: [slot][warp][mb][reg][lane] of {value, flag}; this warp only ever touches its own [warp] part
    return a.ws_partials + static_cast<int64_t>(tile) * a.max_slots * FRAG + warp * (MB * 4 * 32);
  };
  auto resolve_pending = [&]() {
    uint2* wst = slot_base(pend_tile);
    float v[MB][4];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int r = 0; r < 4; ++r) v[mb][r] = 0.f;
    auto fetch = [&](uint2 (&dst)[MB * 4], int sl) {
#pragma unroll
      for (int i = 0; i < MB * 4; ++i) dst[i] = ld_volatile_u2(wst + static_cast<int64_t>(sl) * FRAG + i * 32 + lane);
    };
    // the loads of up to RB slots are in flight together, and when some of them have not landed yet ALL the missing
    // ones of the batch are requested again together: the reducer pays about one L2 round trip per round, not one per
    // late contributor (measured with per-CTA time stamps, profiles/r02_chain_cta_trace.txt: re-fetching one slot at
    // a time kept the reducer of a 10-way split tile 8 us past the last contributor -- they all finish within a
    // microsecond of each other, i.e. every slot was late once).  The sum still runs in slot order: deterministic.
    constexpr int RB = MB == 1 ? 6 : MB == 2 ? 3 : 2;
    for (int base = 0; base < pend_n; base += RB) {
      uint2 buf[RB][MB * 4];
      uint32_t ready = 0;                                     // warp-uniform: slots of the batch that are complete
      const int nb = min(RB, pend_n - base);
      const uint32_t all = (1u << nb) - 1u;
      for (;;) {
        if constexpr (CHAIN) {                                // development: count the polling rounds of warp 0
          unsigned long long* row = *reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE);
          if (tid == 0 && row != nullptr) row[7] = (base == 0 && ready == 0 ? 0ull : row[7]) + 1ull;
        }
#pragma unroll
        for (int s = 0; s < RB; ++s)
          if (s < nb && !((ready >> s) & 1u)) fetch(buf[s], base + s);
#pragma unroll
        for (int s = 0; s < RB; ++s) {
          if (s < nb && !((ready >> s) & 1u)) {
            uint32_t ok = 1;
#pragma unroll
            for (int i = 0; i < MB * 4; ++i) ok &= buf[s][i].y;
            if (__all_sync(0xffffffffu, ok != 0)) ready |= 1u << s;
          }
        }
        if (ready == all) break;
      }
#pragma unroll
      for (int s = 0; s < RB; ++s) {
        if (s < nb) {
#pragma unroll
          for (int i = 0; i < MB * 4; ++i) {
            v[i >> 2][i & 3] += __uint_as_float(buf[s][i].x);
            wst[static_cast<int64_t>(base + s) * FRAG + i * 32 + lane] = make_uint2(0u, 0u);   // leave the workspace clean
          }
        }
      }
    }
    if constexpr (CHAIN) {
      if (tid == 0) chain_stamp(*reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE), 1);   // polls done
    }
    write_fragment(v, pend_tile * SK_ROWS);
    pend_tile = -1;
  };
  auto finish_segment = [&](int q, int seg_begin) {
    const int tile = q / kblocks;
    const int tile_q0 = tile * kblocks;
    const int tile_q1 = tile_q0 + kblocks;                    // exclusive
    if (seg_begin == tile_q0 && q + 1 == tile_q1) {
      write_fragment(acc, tile * SK_ROWS);                    // this CTA owns the whole tile
    } else {
      const int c_first = unit_owner(tile_q0, G, U);
      const int c_last = unit_owner(tile_q1 - 1, G, U);
      uint2* wsp = slot_base(tile) + static_cast<int64_t>(cta - c_first) * FRAG;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          st_volatile_u2(wsp + (mb * 4 + r) * 32 + lane, make_uint2(__float_as_uint(acc[mb][r]), 1u));
      if (cta == c_last) {                                    // reduce at the end of this CTA's run
        pend_tile = tile;
        pend_n = c_last - c_first + 1;
      }
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }
  };

  if constexpr (!CHAIN) pdl_wait();             // x, y and the workspace belong to the stream order
#ifdef SK_TIMING
  const long long c_t0 = clock64(); long long c_stage = 0, c_wait = 0, c_fin = 0, c_comp = 0;
  if (tid == 0 && cta < 512) g_sk_w[a.seq & 15][cta][1] = gtime();
#endif

  uint32_t soff = rp.soff, boff = rp.boff;      // byte offsets of the current stage / its barriers
  int sleft = rp.sleft;                         // stages until the ring wraps
  uint32_t ph = rp.ph;
  int seg_begin = q_begin;                      // first block of the current tile segment
  int kb = q_begin % kblocks;                   // block index inside its tile
  for (int qc = q_begin; qc < q_end; qc += a.chunk) {
    const int nblk = min(a.chunk, q_end - qc);
    SK_T(ts0);
    bar_sync_consumers();                       // everyone is done with the previous slab
    stage_x(qc, nblk);
    bar_sync_consumers();
    if constexpr (CHAIN) {
      if (tid == 0 && qc == q_begin) chain_stamp(*reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE), 3);
    }
#ifdef SK_TIMING
    c_stage += clock64() - ts0;
#endif
    uint32_t xoff = 0;
#pragma unroll 1
    for (int bl = 0; bl < nblk; ++bl) {
      SK_T(tw0);
      mbar_wait_tight_u32(full_u32 + boff, ph);
#ifdef SK_TIMING
      const long long tw1 = clock64(); c_wait += tw1 - tw0;
#endif
      BlockRegs<NBITS> b;
      b.c0.load(cbase + soff, chbase + soff);
      b.c1.load(cbase + soff + LaneCodes<NBITS>::kRow8, chbase + soff + 64);
      b.m0 = lds32(mbase_l + soff);
      b.m1 = lds32(mbase_l + soff + 32);
      __syncwarp();
      if (lane == 0) mbar_arrive_u32(empty_u32 + boff);     // slot may be refilled: the block is in registers
      soff += stage_pitch; boff += 8;
      if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
      compute_block(b, xoff);
      xoff += blk_pitch;
#ifdef SK_TIMING
      const long long tc1 = clock64(); c_comp += tc1 - tw1;
#endif
      const int q = qc + bl;
      if (++kb == kblocks || q == q_end - 1) {
        finish_segment(q, seg_begin);
        seg_begin = q + 1;
        kb = 0;
#ifdef SK_TIMING
        c_fin += clock64() - tc1;
#endif
      }
    }
  }
#ifdef SK_TIMING
  if (tid == 0 && cta < 512) g_sk_w[a.seq & 15][cta][2] = pend_tile >= 0 ? gtime() : -gtime();
#endif
  if constexpr (CHAIN) {
    if (tid == 0) chain_stamp(*reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE), 4);
  }
  if (pend_tile >= 0) resolve_pending();
  if constexpr (CHAIN) {
    if (tid == 0) {
      *reinterpret_cast<RingPos*>(smem_raw + CH_SM_RPOS) = RingPos{soff, boff, sleft, ph};
      chain_stamp(*reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE), 5);
    }
  }
#ifdef SK_TIMING
  if (tid == 0 && cta < 512) g_sk_w[a.seq & 15][cta][3] = gtime();
  if (tid == 0) {
    g_sk_t[cta][0] = clock64() - c_t0; g_sk_t[cta][1] = c_stage; g_sk_t[cta][2] = c_wait; g_sk_t[cta][3] = c_comp;
    g_sk_t[cta][4] = c_fin; g_sk_t[cta][5] = q_end - q_begin;
  }
#endif
}

template <typename T, int NBITS, int MB, int MODE>
__device__ __noinline__ void consumer_gemm_call() {
  consumer_gemm<T, NBITS, MB, MODE, true>(SkinnyArgs{}, 0u, 0u);
}

// CHAIN: the kernel body run over a list of phases (q100_chain.cuh) instead of one GEMM; `a_in` then carries what
// is common to all GEMM phases (M, group, ring geometry, workspace) and the per-phase arguments come from `ch`.
template <typename T, int NBITS, int MB, int MODE, bool CHAIN, bool INLINE_GEMM = true>
__device__ __forceinline__ void skinny_body(const SkinnyArgs& a_in, const ChainArgs& ch) {
  using T2 = typename DQ<T>::T2;
  constexpr bool POST = MODE == 1;
  constexpr bool IMMA = MODE == 2;
  constexpr bool kBf16 = MODE == 0 && DT<T>::kCode == Q100_BF16;  // MMA operand type (POST: always fp16)
  constexpr bool kConvX = POST && DT<T>::kCode == Q100_BF16;    // bf16 activations -> scaled fp16 at staging
  constexpr bool LOHI = POST && UseLoHi<NBITS>::value;
  using DQM = DQ<__half>;                                        // magic constants of the POST path
  constexpr int kTileB = WGeom<NBITS>::kTileBytes;
  constexpr int kBlkB = 2 * kTileB;                   // one [128 rows x 128 k] block of codes in HBM
  constexpr int FRAG = SK_ROWS * MB * 8;              // floats per partial tile

  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t full_bar_s[CHAIN ? 1 : SK_MAX_STAGES], empty_bar_s[CHAIN ? 1 : SK_MAX_STAGES];
  uint64_t* full_bar = full_bar_s;
  uint64_t* empty_bar = empty_bar_s;
  if constexpr (CHAIN) {                              // chain mode: everything at fixed places of the dynamic region
    full_bar = reinterpret_cast<uint64_t*>(smem_raw + CH_SM_FULL);
    empty_bar = reinterpret_cast<uint64_t*>(smem_raw + CH_SM_EMPTY);
  }
  ChainPhase& s_ph = *reinterpret_cast<ChainPhase*>(smem_raw + CH_SM_PHASE);
  float* s_rstd = reinterpret_cast<float*>(smem_raw + CH_SM_RSTD);
  const SkinnyArgs* ap = &a_in;
  if constexpr (CHAIN) ap = &s_ph.g;
  const SkinnyArgs& a = *ap;                          // per-phase arguments (== a_in without CHAIN)
  const uint32_t ring_u32 = CHAIN ? smem_u32(smem_raw) + CH_SM_RING : ((smem_u32(smem_raw) + 127u) & ~127u);
  const int stage_pitch = (stage_bytes<NBITS>(a_in.nsub) + 127) & ~127;
  const int stages = a_in.stages;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  int G = gridDim.x, U = a_in.total_units;            // chain: set per phase (G = the phase's share of the grid)
  const int cta = blockIdx.x;
#ifdef SK_TIMING
  if (tid == 0 && cta < 512) g_sk_w[a_in.seq & 15][cta][0] = gtime();
#endif
  int q_begin = static_cast<int>(static_cast<int64_t>(cta) * U / G);
  int q_end = static_cast<int>(static_cast<int64_t>(cta + 1) * U / G);
  int kblocks = a_in.kblocks;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], SK_CWARPS); }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  if (warp == SK_CWARPS) {
    // ===================== producer: codes + side band of blocks [q_begin, q_end) =====================
    // (static data: no griddepcontrol.wait on this path).  One lane per bulk copy of a stage -- lane 0/1
    // the two code tiles, lanes 2.. the (scale, zero) rows -- each with its own incrementally updated
    // source pointer, so that issuing a stage costs the warp a couple of dozen instructions.
    if (lane < 2 + a_in.nsub) {
      const uint64_t pol = policy_evict_first();
      const bool is_code = lane < 2;
      const int sub = lane - 2;
      const int esz2 = 2 * static_cast<int>(sizeof(T));
      uint32_t soff = 0, boff = 0;                    // ring position: carried across the phases of a chain
      int sleft = stages;
      uint32_t ph = 1;                                // parity the producer waits on (slots start free)
      const uint32_t full_u32 = smem_u32(&full_bar[0]), empty_u32 = smem_u32(&empty_bar[0]);
#ifdef SK_TIMING
      long long p_wait = 0; const long long p_t0 = clock64();
#endif
      // codes + side band of blocks [qb, qe) of one GEMM (pa: its arguments, kblk_n = its K / 128)
      auto produce = [&](const SkinnyArgs& pa, const int kblocks, const int q_begin, const int q_end) {
      const int64_t mrow_pitch = static_cast<int64_t>(pa.N) * esz2;
      const int gdiv = pa.group >= SK_KB ? min(pa.group / SK_KB, kblocks) : 1;     // k blocks per side-band row
      const int64_t mstep = pa.group >= SK_KB ? mrow_pitch : pa.nsub * mrow_pitch;    // advance every gdiv blocks
      int tile = q_begin / kblocks, kblk = q_begin - tile * kblocks;
      const uint8_t* mtile = reinterpret_cast<const uint8_t*>(pa.meta) + static_cast<int64_t>(tile) * SK_ROWS * esz2;
      const uint8_t* src;
      uint32_t dst_off, bytes;
      int gleft = gdiv - kblk % gdiv;                 // blocks until the side-band row changes
      int mbytes = min(SK_ROWS, pa.N - tile * SK_ROWS) * esz2;
      if (is_code) {
        src = reinterpret_cast<const uint8_t*>(pa.packed) + static_cast<int64_t>(q_begin) * kBlkB + lane * kTileB;
        dst_off = lane * WGeom<NBITS>::kTileStride;
        bytes = kTileB;
      } else {
        const int kg = pa.group >= SK_KB ? kblk / gdiv : kblk * pa.nsub + sub;
        src = mtile + kg * mrow_pitch;
        dst_off = WGeom<NBITS>::kCodeBytes + sub * 512;
        bytes = mbytes;
      }
      uint32_t tx = kBlkB + pa.nsub * mbytes;
      int kleft = kblocks - kblk;                     // blocks until the tile changes
      const uint32_t ring_base = ring_u32 + dst_off;
#pragma unroll 1
      for (int q = q_begin; q < q_end; ++q) {
        SK_T(pw0);
        mbar_wait_tight_u32(empty_u32 + boff, ph);
#ifdef SK_TIMING
        p_wait += clock64() - pw0;
#endif
        if (lane == 0) mbar_expect_tx_u32(full_u32 + boff, tx);
        bulk_load_1d_hint_u32(ring_base + soff, src, bytes, full_u32 + boff, pol);
        soff += stage_pitch; boff += 8;
        if (--sleft == 0) { sleft = stages; soff = 0; boff = 0; ph ^= 1; }
        if (is_code) src += kBlkB;
        else if (--gleft == 0) { gleft = gdiv; src += mstep; }
        if (--kleft == 0) {                           // next tile
          kleft = kblocks; ++tile;
          mtile += SK_ROWS * esz2;
          mbytes = min(SK_ROWS, pa.N - tile * SK_ROWS) * esz2;
          tx = kBlkB + pa.nsub * mbytes;
          if (!is_code) { src = mtile + (pa.group >= SK_KB ? 0 : sub) * mrow_pitch; bytes = mbytes; gleft = gdiv; }
        }
      }
      };
      if constexpr (CHAIN) {
        // the weights of every phase are static: the producer never waits at a phase boundary, it is only ever
        // held back by the depth of the ring
        for (int p = 0; p < ch.n_phases; ++p) {
          const ChainPhase* php = ch.phases + p;
          if (php->type != PH_GEMM) continue;
          const int gp = php->grid_p;
          if (cta >= gp) continue;
          const int u = php->g.total_units;
          produce(php->g, php->g.kblocks, static_cast<int>(static_cast<int64_t>(cta) * u / gp),
                  static_cast<int>(static_cast<int64_t>(cta + 1) * u / gp));
        }
      } else {
        produce(a_in, kblocks, q_begin, q_end);
      }
#ifdef SK_TIMING
      if (lane == 0) { g_sk_t[cta][6] = p_wait; g_sk_t[cta][7] = clock64() - p_t0; }
#endif
    }
    return;
  }

  if constexpr (CHAIN) {
    pdl_wait();
    if (tid == 0) {
      *reinterpret_cast<RingPos*>(smem_raw + CH_SM_RPOS) = RingPos{0u, 0u, stages, 0u};
      // the position does not change during a launch: one read for all attention phases
      const int64_t* pp = *ch.pos_pp;
      *reinterpret_cast<int64_t*>(smem_raw + CH_SM_POS) = pp != nullptr ? __ldcg(pp) : -1;
    }
    // phase arguments: the first set straight from global memory, every later one prefetched with cp.async into a
    // second slot while its predecessor runs (static data: never waits for anything)
    auto prefetch_phase = [&](int p) {
      if (tid < CH_PHASE_BYTES / 16) {
        const uint32_t dst = smem_u32(smem_raw) + CH_SM_NEXT + tid * 16;
        const void* src = reinterpret_cast<const uint8_t*>(ch.phases + p) + tid * 16;
        if (tid * 16 < static_cast<int>(sizeof(ChainPhase)))
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
      }
    };
    prefetch_phase(0);
    // grid barrier between phases: every CTA of the (fully resident) grid arrives once per phase.  bar.sync orders the
    // CTA's y / ssq stores before thread 0's gpu-scope fence + arrival; the acquire on the way out (+ reading x
    // through L2) makes the other CTAs' stores visible.  A bounded spin: a protocol bug sets err instead of hanging.
    auto grid_barrier = [&](int p) {
      bar_sync_consumers();
      if (tid == 0) {
        __threadfence();
        if (ch.trace != nullptr) {
          unsigned long long t0;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          atomicMin(ch.trace + 3 * p, t0);
          atomicMax(ch.trace + 3 * p + 1, t0);
        }
        atomicAdd(ch.bars + p, 1u);
        uint32_t spins = 0;
        // relaxed polls with a back-off (hundreds of CTAs wait here for the phases only a part of the grid works on),
        // one fence on the way out
        while (ld_relaxed_gpu_u32(ch.bars + p) < gridDim.x) {
          __nanosleep(100);
          ++spins;
          // give up after ~2 s, and at once if another barrier already did: the launch then ends quickly instead of
          // serialising one time-out per phase
          if (spins > (1u << 22) || ((spins & 1023u) == 0 && ld_acquire_gpu_u32(ch.err) != 0u)) { *ch.err = 1u; break; }
        }
        __threadfence();
        if (ch.trace != nullptr) {
          unsigned long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          atomicMax(ch.trace + 3 * p + 2, t1);
        }
      }
      bar_sync_consumers();
    };
    for (int p = 0; p < ch.n_phases; ++p) {
      bar_sync_consumers();                                   // everyone is done with the previous phase's arguments
      unsigned long long* tr_row = nullptr;
      if (tid == 0) {
        if (ch.trace != nullptr && cta == 0 && (p == 0 || p == ch.n_phases - 1)) {
          // SM clock over the launch: cycles and nanoseconds at the first and the last phase (last trace row)
          unsigned long long tns;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
          unsigned long long* r = ch.trace + 3 * (CH_MAX_PHASES - 2) + (p == 0 ? 0 : 2);
          r[0] = static_cast<unsigned long long>(clock64());
          r[1] = tns;
        }
        if (ch.trace != nullptr && p == ch.trace_phase && cta < 512)
          tr_row = ch.trace + (CH_CTA_TRACE_OFF - CH_TRACE_OFF) / 8 + cta * 8;
        *reinterpret_cast<unsigned long long**>(smem_raw + CH_SM_TRACE) = tr_row;
        chain_stamp(tr_row, 0);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      bar_sync_consumers();                                   // the prefetched arguments of phase p have landed
      if (tid < static_cast<int>(sizeof(ChainPhase) / 16))
        reinterpret_cast<uint4*>(&s_ph)[tid] = reinterpret_cast<const uint4*>(smem_raw + CH_SM_NEXT)[tid];
      bar_sync_consumers();
      if (p + 1 < ch.n_phases) prefetch_phase(p + 1);
      if (tid == 0) chain_stamp(tr_row, 1);
      if (s_ph.type == PH_GEMM) {
        G = s_ph.grid_p;
        if (cta < G) {
          if (a.x_norm) {
            const int mrows = a.m_rows;
            // 1 / rms of every token of the residual stream from the partial sums the producing phase left
            for (int m = warp; m < mrows; m += SK_CWARPS) {
              float tot = 0.f;
              for (int i0 = 0; i0 < a.n_part_in; i0 += 256) {     // eight independent loads per lane in flight
                float pv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int i = i0 + u * 32 + lane;
                  pv[u] = i < a.n_part_in ? __ldcg(a.ssq_in + static_cast<int64_t>(i) * mrows + m) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) tot += pv[u];
              }
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
              if (lane == 0) s_rstd[m] = rsqrtf(__fadd_rn(__fdiv_rn(tot, static_cast<float>(a.K)), a.eps));
            }
          }
          // (the first bar_sync inside publishes s_rstd)
          if (tid == 0) chain_stamp(tr_row, 2);
          if constexpr (INLINE_GEMM) consumer_gemm<T, NBITS, MB, MODE, true>(a_in, 0u, 0u);
          else consumer_gemm_call<T, NBITS, MB, MODE>();
        }
      } else if (s_ph.type == PH_ATTN) {
        chain_attn<T>(s_ph.at, *reinterpret_cast<const int64_t*>(smem_raw + CH_SM_POS),
                      reinterpret_cast<float*>(smem_raw + CH_SM_ATT), cta, static_cast<int>(gridDim.x), warp, lane);
      } else {
        chain_combine<T>(s_ph.at, *reinterpret_cast<const int64_t*>(smem_raw + CH_SM_POS), cta,
                         static_cast<int>(gridDim.x), warp, lane);
      }
      if (tid == 0) chain_stamp(tr_row, 6);
      if (p + 1 < ch.n_phases) grid_barrier(p);
    }
    // the last CTA out leaves the arrival counters zeroed for the next launch
    bar_sync_consumers();
    if (tid == 0) {
      const uint32_t old = atomicAdd(ch.bars + ch.n_phases, 1u);
      if (old == gridDim.x - 1) {
        for (int i = 0; i <= ch.n_phases; ++i) ch.bars[i] = 0u;
        __threadfence();
      }
    }
  } else {
    consumer_gemm<T, NBITS, MB, MODE, false>(a_in, smem_u32(&full_bar[0]), smem_u32(&empty_bar[0]));
  }
}

template <typename T, int NBITS, int MB, int MODE>
__global__ void __launch_bounds__(SK_THREADS, MB <= 1 ? SK_OCC : 2)
k3_skinny(const SkinnyArgs a) {
  skinny_body<T, NBITS, MB, MODE, false>(a, ChainArgs{});
}

// the chain form: same body, phases from device memory (q100_chain.cuh)
#ifndef CHAIN_LB
#define CHAIN_LB(occ) (occ)
#endif
// OCC = resident CTAs per SM the instantiation is compiled for: 3 leaves 72 registers per thread (the GEMM share is then
// a separate function, see consumer_gemm_call), 2 leaves 112 (everything inlined, no spills)
template <typename T, int NBITS, int MB, int MODE, int OCC>
__global__ void __launch_bounds__(SK_THREADS, CHAIN_LB(OCC))
k3_chain(const SkinnyArgs a, const ChainArgs ch) {
  skinny_body<T, NBITS, MB, MODE, true, (OCC < 3)>(a, ch);
}

// ---------------------------------------------------------------- host
struct SkinnyPlan {
  int n_tiles, total_units, max_slots, grid_cap;
};

static SkinnyPlan plan_skinny(int64_t N, int64_t K) {
  SkinnyPlan p;
  const int kblocks = static_cast<int>(K / SK_KB);
  p.n_tiles = static_cast<int>((N + SK_ROWS - 1) / SK_ROWS);
  p.total_units = p.n_tiles * kblocks;
  int grid = sm_count() * SK_OCC;             // upper bound of the persistent grid
  if (grid > p.total_units) grid = p.total_units;
  // small problems: do not cut a tile more than SK_MAX_SPLIT ways unless a CTA would get more than
  // SK_SPLIT_UNITS blocks.  Measured with globaltimer stamps (profiles/r02_decode_timeline.txt): o_proj
  // (32 tiles x 32 blocks) cut 14 ways over 444 CTAs spends 4.6 us in the reducer's wait; 128 CTAs of 8
  // blocks finish in 8.5 us instead of 10.9 us.
  int min_units = (kblocks + SK_MAX_SPLIT - 1) / SK_MAX_SPLIT;
  if (min_units > SK_SPLIT_UNITS) min_units = SK_SPLIT_UNITS;
  const int cap = (p.total_units + min_units - 1) / min_units;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  p.grid_cap = grid;
  // a tile of `kblocks` units can be touched by at most this many CTAs
  p.max_slots = static_cast<int>((static_cast<int64_t>(kblocks) * grid + p.total_units - 1) / p.total_units) + 2;
  return p;
}

template <typename T, int NBITS, int MB, int MODE>
static int launch_skinny(SkinnyArgs a, cudaStream_t st) {
  auto kern = k3_skinny<T, NBITS, MB, MODE>;
  constexpr int xpitch = MODE == 2 ? SK_IPITCH : MODE == 3 ? SK_ICOL : SK_XPITCH;
  constexpr int occ_target = MB <= 1 ? SK_OCC : 2;
  const int m_rows_real = a.m_rows;
  if (MODE == 3) a.m_rows += 1;                   // sizing only: the all-ones column rides with the tokens
  // per-CTA shared memory: ring + x slab within the budget that keeps occ_target CTAs resident
  const int budget = (228 * 1024) / occ_target - 1024 /* driver reserve */ - 512 /* static */ - 128 /* align */;
  const int stage_pitch = (stage_bytes<NBITS>(a.nsub) + 127) & ~127;
  const int row_b = a.m_rows * xpitch;
  // reserve ~2/3 of the budget (at most 6 stages) for the ring; then as much x as fits up to SK_CHUNK_MAX
  // blocks; what is left goes back to the ring
  int min_stages = (2 * budget / 3) / stage_pitch;
  if (min_stages > 6) min_stages = 6;
  if (min_stages < 2) min_stages = 2;
  int chunk = (budget - min_stages * stage_pitch - xpitch) / row_b;
  if (chunk > SK_CHUNK_MAX) chunk = SK_CHUNK_MAX;
  Q100_REQUIRE(chunk >= 1, Q100_EUNSUPPORTED, "k3_skinny: M=%d rows of x do not fit in shared memory", a.m_rows);
  // a CTA never needs more x than its share of the blocks: keep the slab small so that the ring is deep
  const int grid_est = plan_skinny(a.N, a.K).grid_cap;
  const int share = (a.total_units + grid_est - 1) / grid_est + 1;
  if (chunk > share) chunk = share;
  int stages = (budget - slab_bytes(a.m_rows, chunk, xpitch)) / stage_pitch;
  if (stages > SK_MAX_STAGES) stages = SK_MAX_STAGES;
  a.chunk = chunk;
  a.stages = stages;
  const int smem = stages * stage_pitch + slab_bytes(a.m_rows, chunk, xpitch) + 128;
  a.m_rows = m_rows_real;
  static std::mutex mu;           // one per instantiation
  static int max_set[kMaxDevices] = {};            // the attribute is a per-device property
  static std::unordered_map<int, int> occ_cache;   // all devices are sm_100: occupancy is the same on each
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (smem > max_set[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      max_set[dev] = smem;
    }
    auto itc = occ_cache.find(smem);
    if (itc == occ_cache.end()) {
      Q100_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SK_THREADS, smem));
      occ_cache[smem] = occ;
    } else {
      occ = itc->second;
    }
  }
  Q100_REQUIRE(occ >= 1, Q100_ECUDA, "k3_skinny does not fit on an SM (%d bytes smem)", smem);
  // persistent grid: every CTA resident at once, units split evenly (stream-K)
  int grid = sm_count() * (occ > SK_OCC ? SK_OCC : occ);
  if (grid > a.total_units) grid = a.total_units;
  if (grid > grid_est) grid = grid_est;         // split cap, see plan_skinny
  Q100_REQUIRE(static_cast<int64_t>(a.kblocks) * grid / a.total_units + 2 <= a.max_slots, Q100_EWORKSPACE,
               "internal: slot bound");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, a), "k3_skinny launch");
}

template <typename T, int NBITS, int MB>
static int launch_skinny_post(const SkinnyArgs& a, cudaStream_t st) {
  if (a.post) return launch_skinny<T, NBITS, MB, 1>(a, st);
  return launch_skinny<T, NBITS, MB, 0>(a, st);
}

template <typename T, int NBITS>
static int launch_skinny_mb(const SkinnyArgs& a, cudaStream_t st) {
  if constexpr (NBITS == 4 || NBITS == 8) {
    // activation codes x weight codes on the integer tensor cores (q100_gemm_wna8); one mma column is the all-ones column
    if (a.post == 3) return a.M <= 7 ? launch_skinny<T, NBITS, 1, 3>(a, st) : launch_skinny<T, NBITS, 2, 3>(a, st);
  }
  if constexpr (NBITS == 4 || NBITS == 2) {
    // integer tensor-core path.  Measured (profiles/r01k): with one column group (M <= 2) it beats the fp16 form
    // (gate|up 16.6 vs 19.3 us); with two groups (M = 4) the extra MMAs, int->float conversions and limb staging
    // lose (25.3 vs 21.6 us), so larger M stay on the fp16 form.
    if (a.post == 2) return launch_skinny<T, NBITS, 1, 2>(a, st);
  }
  if (a.M <= 8) return launch_skinny_post<T, NBITS, 1>(a, st);
  if (a.M <= 16) return launch_skinny_post<T, NBITS, 2>(a, st);
  return launch_skinny_post<T, NBITS, 4>(a, st);
}

template <typename T>
static int launch_skinny_bits(int n_bits, const SkinnyArgs& a, cudaStream_t st) {
  switch (n_bits) {
    case 2: return launch_skinny_mb<T, 2>(a, st);
    case 3: return launch_skinny_mb<T, 3>(a, st);
    case 4: return launch_skinny_mb<T, 4>(a, st);
    default: return launch_skinny_mb<T, 8>(a, st);
  }
}

#ifdef SK_TIMING
int skinny_timing(long long* out, int n) {
  if (n < 0) return check_cuda(cudaMemcpyFromSymbol(out, g_sk_w, sizeof(g_sk_w)), "wall timing readback");
  return check_cuda(cudaMemcpyFromSymbol(out, g_sk_t, sizeof(long long) * 8 * n), "timing readback");
}
#endif

bool gemm_rows_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits, int mode);
int gemm_rows(const void* x, const void* packed, const void* meta, const void* bias, void* y, int64_t M, int64_t N,
              int64_t K, int disable_zp, int dtype, int mode, int pdl, int x_gate_up, cudaStream_t st,
              const void* norm_w = nullptr, float eps = 0.f, const void* res = nullptr);

size_t skinny_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const SkinnyPlan p = plan_skinny(N, K);
  const int mb = M <= 8 ? 1 : M <= 16 ? 2 : 4;
  const size_t partials = static_cast<size_t>(p.n_tiles) * p.max_slots * SK_ROWS * mb * 8 * sizeof(uint2);
  return partials + SK_COUNTER_BYTES;
}

int gemm_skinny(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int exact_weights, int pdl, int x_gate_up, cudaStream_t st,
                const void* x_minmax, int a_bits, int a_cuda_scale, int no_rows) {
  Q100_REQUIRE(M >= 1 && M <= 32, Q100_EINVAL, "skinny GEMM handles 1 <= M <= 32 (got %lld)", (long long)M);
  Q100_REQUIRE(K % SK_KB == 0, Q100_EUNSUPPORTED, "K=%lld must be a multiple of 128", (long long)K);
  Q100_REQUIRE(N % 8 == 0, Q100_EUNSUPPORTED, "N=%lld must be a multiple of 8", (long long)N);
  const int g = group == 0 ? static_cast<int>(K) : group;
  Q100_REQUIRE(g == 32 || g == 64 || g % 128 == 0, Q100_EUNSUPPORTED,
               "group=%d: supported 32, 64 or a multiple of 128", g);
  Q100_REQUIRE(K % g == 0, Q100_EINVAL, "K %% group != 0");
  const SkinnyPlan p = plan_skinny(N, K);
  const size_t need = skinny_workspace_bytes(M, N, K);
  Q100_REQUIRE(workspace != nullptr && workspace_bytes >= need, Q100_EWORKSPACE,
               "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);

  SkinnyArgs a{};
  a.x = x; a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  // the first SK_COUNTER_BYTES of the workspace belong to the arrival counters of the tcgen05 streaming
  // kernel (q100_gemm_stream.cu); this kernel's {value, flag} slots start behind them and are left zeroed
  a.ws_partials = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(workspace) + SK_COUNTER_BYTES);
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = g;
  a.kblocks = static_cast<int>(K / SK_KB);
  a.n_tiles = p.n_tiles;
  a.total_units = p.total_units; a.max_slots = p.max_slots;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  // exact_weights: 0 = default, 1 = bit-exact 16-bit weights (MODE 0), 2 = fp16 magic-form path even where
  // the integer path applies
  a.post = (g % SK_KB == 0 && exact_weights != 1) ? 1 : 0;
  if (a.post && (n_bits == 4 || n_bits == 2) && M <= 2 && exact_weights == 0) a.post = 2;
  if (x_minmax != nullptr) {
    Q100_REQUIRE(a.post != 0 && (n_bits == 4 || n_bits == 8) && M <= 15 && a_bits >= 2 && a_bits <= 8 && !x_gate_up,
                 Q100_EUNSUPPORTED, "integer W x A path: W4 / W8, group %% 128 == 0, M <= 15, abits 2..8");
    a.post = 3;
    a.x_minmax = x_minmax;
    a.a_qmax = static_cast<float>((1 << a_bits) - 1);
    a.a_denom = a.a_qmax;
    a.a_cuda_scale = a_cuda_scale;
  }
  // K3r (q100_gemm_rows.cu): every CTA owns whole rows, no fix-up between CTAs -- the default where it applies
  if (!no_rows && (a.post == 1 || a.post == 2) && gemm_rows_supported(M, N, K, g, n_bits, a.post))
    return gemm_rows(x, packed, meta, bias, y, M, N, K, disable_zp, dtype, a.post, pdl, x_gate_up, st);
  a.m_rows = static_cast<int>(M);
  a.nsub = g >= SK_KB ? 1 : SK_KB / g;
  a.pdl = pdl;
  a.x_gate_up = x_gate_up;
  static int launch_seq = 0;
  a.seq = launch_seq++;
  if (dtype == Q100_F16) return launch_skinny_bits<__half>(n_bits, a, st);
  return launch_skinny_bits<__nv_bfloat16>(n_bits, a, st);
}


// ---------------------------------------------------------------- chain: host side
constexpr size_t CH_HEADER_BYTES = 8192 + 24 * CH_MAX_PHASES + 512 * 64;   // arrival counters [CH_MAX_PHASES + 1], error flag (8 KB); trace stamps; then ChainPhase[n]

constexpr int CH_ERR_WORD = CH_MAX_PHASES + 4;
constexpr int CH_CHUNK_CAP = 16;                  // k blocks of x staged per round (gate|up: 16 blocks per CTA at 444 CTAs)

struct ChainGeom { int smem, stages, chunk, grid, mb, mode; };

static int chain_occ_env() {
  static const int v = []() { const char* e = getenv("Q100_CHAIN_OCC"); return e ? atoi(e) : 0; }();
  return v;
}

template <typename T, int NBITS, int MB, int MODE, int OCC>
static int chain_geometry_t(int m_rows, ChainGeom& g) {
  auto kern = k3_chain<T, NBITS, MB, MODE, OCC>;
  constexpr int xpitch = MODE == 2 ? SK_IPITCH : SK_XPITCH;
  constexpr int occ_target = OCC;
  const int budget = (228 * 1024) / occ_target - 1024 - 128 - CH_SM_RING - 256;
  const int stage_pitch = (stage_bytes<NBITS>(1) + 127) & ~127;
  const int row_b = m_rows * xpitch;
  int min_stages = (2 * budget / 3) / stage_pitch;
  if (min_stages > 6) min_stages = 6;
  if (min_stages < 2) min_stages = 2;
  int chunk = (budget - min_stages * stage_pitch - xpitch) / row_b;
  if (chunk > CH_CHUNK_CAP) chunk = CH_CHUNK_CAP;
  Q100_REQUIRE(chunk >= 1, Q100_EUNSUPPORTED, "k3_chain: M=%d rows of x do not fit in shared memory", m_rows);
  int stages = (budget - slab_bytes(m_rows, chunk, xpitch)) / stage_pitch;
  if (stages > SK_MAX_STAGES) stages = SK_MAX_STAGES;
  const int smem = CH_SM_RING + stages * stage_pitch + slab_bytes(m_rows, chunk, xpitch) + 128;
  static std::mutex mu;
  static int max_set[kMaxDevices] = {};
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (smem > max_set[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      max_set[dev] = smem;
    }
    Q100_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SK_THREADS, smem));
  }
  Q100_REQUIRE(occ >= 1, Q100_ECUDA, "k3_chain does not fit on an SM (%d bytes smem)", smem);
  g.smem = smem; g.stages = stages; g.chunk = chunk;
  g.grid = sm_count() * (occ > occ_target ? occ_target : occ);   // every CTA resident: the grid barrier relies on it
  g.mb = MB; g.mode = MODE;
  return Q100_OK;
}

template <typename T, int NBITS, int MB, int MODE, int OCC>
static int chain_launch_t(const SkinnyArgs& a, const ChainArgs& ch, const ChainGeom& g, cudaStream_t st) {
  auto kern = k3_chain<T, NBITS, MB, MODE, OCC>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.grid);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, a, ch), "k3_chain launch");
}

// the one place that maps (dtype, bits, M) to an instantiation: op 0 = geometry, op 1 = launch
template <typename T, int NBITS, int MB, int MODE>
static int chain_dispatch_occ(int op, int M, const SkinnyArgs* a, const ChainArgs* ch, ChainGeom& g, cudaStream_t st) {
  const int want = chain_occ_env();
  if (MB <= 1 && want == 3)
    return op == 0 ? chain_geometry_t<T, NBITS, MB, MODE, (MB <= 1 ? 3 : 2)>(M, g)
                   : chain_launch_t<T, NBITS, MB, MODE, (MB <= 1 ? 3 : 2)>(*a, *ch, g, st);
  return op == 0 ? chain_geometry_t<T, NBITS, MB, MODE, 2>(M, g) : chain_launch_t<T, NBITS, MB, MODE, 2>(*a, *ch, g, st);
}

template <typename T, int NBITS>
static int chain_dispatch_bits(int op, int M, const SkinnyArgs* a, const ChainArgs* ch, ChainGeom& g, cudaStream_t st) {
  const bool imma = (NBITS == 4 || NBITS == 2) && M <= 2;
  if constexpr (NBITS == 4 || NBITS == 2) {
    if (imma) return chain_dispatch_occ<T, NBITS, 1, 2>(op, M, a, ch, g, st);
  }
  if (M <= 8) return chain_dispatch_occ<T, NBITS, 1, 1>(op, M, a, ch, g, st);
  return chain_dispatch_occ<T, NBITS, 2, 1>(op, M, a, ch, g, st);
}
static int chain_dispatch(int op, int dtype, int n_bits, int M, const SkinnyArgs* a, const ChainArgs* ch, ChainGeom& g,
                          cudaStream_t st) {
  if (dtype == Q100_F16) {
    switch (n_bits) {
      case 2: return chain_dispatch_bits<__half, 2>(op, M, a, ch, g, st);
      case 3: return chain_dispatch_bits<__half, 3>(op, M, a, ch, g, st);
      case 4: return chain_dispatch_bits<__half, 4>(op, M, a, ch, g, st);
      default: return chain_dispatch_bits<__half, 8>(op, M, a, ch, g, st);
    }
  }
  switch (n_bits) {
    case 2: return chain_dispatch_bits<__nv_bfloat16, 2>(op, M, a, ch, g, st);
    case 3: return chain_dispatch_bits<__nv_bfloat16, 3>(op, M, a, ch, g, st);
    case 4: return chain_dispatch_bits<__nv_bfloat16, 4>(op, M, a, ch, g, st);
    default: return chain_dispatch_bits<__nv_bfloat16, 8>(op, M, a, ch, g, st);
  }
}

static void chain_common_args(SkinnyArgs& a, int M, int group, int n_bits, int disable_zp, const ChainGeom& g,
                              void* workspace) {
  a.ws_partials = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(workspace) + SK_COUNTER_BYTES);
  a.M = M; a.group = group; a.m_rows = M;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  a.post = g.mode;
  a.chunk = g.chunk; a.stages = g.stages; a.nsub = 1;
  a.pdl = 0; a.seq = 0;
}

static int chain_attn_chunk(int64_t max_len) { return 64 * static_cast<int>((max_len + 2047) / 2048); }

}  // namespace q100

using namespace q100;

extern "C" size_t q100_chain_desc_bytes(int n_phases) {
  if (n_phases < 0) n_phases = 0;
  return CH_HEADER_BYTES + static_cast<size_t>(n_phases) * sizeof(ChainPhase);
}

extern "C" int64_t q100_chain_ssq_parts(int64_t N) { return ((N + SK_ROWS - 1) / SK_ROWS) * SK_CWARPS; }

extern "C" int64_t q100_chain_attn_partial_floats(int batch, int heads, int64_t max_len) {
  if (batch <= 0 || heads <= 0 || max_len <= 0) return 0;
  const int cp = chain_attn_chunk(max_len);
  return static_cast<int64_t>(batch) * heads * ((max_len + cp - 1) / cp) * AT_PART_STRIDE;
}

extern "C" int q100_chain_build(const q100_phase* phases, int n_phases, int M, int group, int n_bits, int disable_zp,
                                int dtype, void* desc, size_t desc_bytes, void* workspace, size_t workspace_bytes,
                                size_t* workspace_needed, void* stream) {
  Q100_REQUIRE(phases != nullptr && n_phases >= 1 && n_phases <= CH_MAX_PHASES, Q100_EINVAL,
               "1 <= n_phases <= %d (got %d)", CH_MAX_PHASES, n_phases);
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED, "n_bits in {2,3,4,8}");
  Q100_REQUIRE(M >= 1 && M <= 16, Q100_EUNSUPPORTED, "chain handles 1 <= M <= 16 tokens (got %d)", M);
  Q100_REQUIRE(group > 0 && group % SK_KB == 0, Q100_EUNSUPPORTED, "chain needs group %% 128 == 0 (got %d)", group);
  Q100_REQUIRE(desc != nullptr && desc_bytes >= q100_chain_desc_bytes(n_phases), Q100_EINVAL,
               "descriptor buffer too small: need %zu bytes", q100_chain_desc_bytes(n_phases));
  Q100_REQUIRE((reinterpret_cast<uintptr_t>(desc) & 15) == 0, Q100_EINVAL, "descriptor buffer must be 16-byte aligned");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ChainGeom geo{};
  rc = chain_dispatch(0, dtype, n_bits, M, nullptr, nullptr, geo, st);
  if (rc != Q100_OK) return rc;
  std::vector<ChainPhase> host(n_phases);
  size_t need = 0;
  const void* pos_ptr = nullptr;                          // the device-side position: one per chain
  for (int i = 0; i < n_phases; ++i) {
    const bool uses = phases[i].type != Q100_PH_GEMM || phases[i].x_mode == Q100_X_ATTN;
    if (!uses) continue;
    if (pos_ptr == nullptr) pos_ptr = phases[i].position;
    Q100_REQUIRE(phases[i].position == pos_ptr, Q100_EINVAL,
                 "phase %d: every attention phase of a chain must share one position scalar", i);
  }
  for (int i = 0; i < n_phases; ++i) {
    const q100_phase& p = phases[i];
    ChainPhase& c = host[i];
    memset(&c, 0, sizeof(c));
    c.type = p.type;
    if (p.type == Q100_PH_GEMM) {
      Q100_REQUIRE(p.x && p.packed && p.meta && p.y, Q100_EINVAL, "phase %d: NULL pointer argument", i);
      Q100_REQUIRE(p.N > 0 && p.K > 0 && p.K % SK_KB == 0 && p.N % 8 == 0 && p.K % group == 0, Q100_EUNSUPPORTED,
                   "phase %d: N=%lld (%% 8), K=%lld (%% 128, %% group)", i, (long long)p.N, (long long)p.K);
      Q100_REQUIRE(p.x_mode == Q100_X_PLAIN || p.x_mode == Q100_X_GATE_UP || p.x_mode == Q100_X_NORM ||
                   p.x_mode == Q100_X_ATTN, Q100_EINVAL, "phase %d: bad x_mode %d", i, p.x_mode);
      if (p.x_mode == Q100_X_ATTN)
        Q100_REQUIRE(p.att_partials && p.position && p.max_len > 0 && p.heads > 0 && p.K == static_cast<int64_t>(p.heads) * AT_D,
                     Q100_EINVAL, "phase %d: x_mode ATTN needs att_partials / position / max_len and K == heads * 128", i);
      Q100_REQUIRE(p.epi_mode == Q100_EPI_STORE || p.epi_mode == Q100_EPI_ADD_SSQ, Q100_EINVAL, "phase %d: bad epi_mode", i);
      if (p.x_mode == Q100_X_NORM)
        Q100_REQUIRE(p.norm_weight && p.ssq_in && p.ssq_in_parts > 0, Q100_EINVAL, "phase %d: x_mode NORM needs norm_weight / ssq_in", i);
      if (p.epi_mode == Q100_EPI_ADD_SSQ)
        Q100_REQUIRE(p.residual && p.ssq_out && p.N % SK_ROWS == 0, Q100_EINVAL,
                     "phase %d: epi_mode ADD_SSQ needs residual / ssq_out and N %% 128 == 0", i);
      const SkinnyPlan pl = plan_skinny(p.N, p.K);
      SkinnyArgs& a = c.g;
      chain_common_args(a, M, group, n_bits, disable_zp, geo, workspace);
      a.x = p.x; a.packed = p.packed; a.meta = p.meta; a.bias = p.bias; a.y = p.y;
      a.N = static_cast<int>(p.N); a.K = static_cast<int>(p.K);
      a.kblocks = static_cast<int>(p.K / SK_KB);
      a.n_tiles = pl.n_tiles; a.total_units = pl.total_units;
      int gp = pl.grid_cap < geo.grid ? pl.grid_cap : geo.grid;
      // plan_skinny sizes its cap for SK_OCC CTAs per SM; with fewer resident CTAs keep the same units-per-CTA floor
      if (gp > a.total_units) gp = a.total_units;
      c.grid_p = gp;
      a.max_slots = static_cast<int>((static_cast<int64_t>(a.kblocks) * gp + a.total_units - 1) / a.total_units) + 2;
      a.x_gate_up = p.x_mode == Q100_X_GATE_UP;
      a.x_norm = p.x_mode == Q100_X_NORM;
      a.x_attn = p.x_mode == Q100_X_ATTN;
      if (a.x_attn) {
        a.att_part = reinterpret_cast<const float*>(p.att_partials);
        a.att_pos = reinterpret_cast<const int64_t*>(p.position);
        a.att_chunk_pos = chain_attn_chunk(p.max_len);
        a.att_max_parts = static_cast<int>((p.max_len + a.att_chunk_pos - 1) / a.att_chunk_pos);
      }
      a.norm_w = p.norm_weight; a.ssq_in = reinterpret_cast<const float*>(p.ssq_in); a.n_part_in = p.ssq_in_parts;
      a.eps = p.eps;
      a.epi_add = p.epi_mode == Q100_EPI_ADD_SSQ;
      a.res = p.residual; a.ssq_out = reinterpret_cast<float*>(p.ssq_out);
      const size_t ws = static_cast<size_t>(a.n_tiles) * a.max_slots * SK_ROWS * geo.mb * 8 * sizeof(uint2) + SK_COUNTER_BYTES;
      if (ws > need) need = ws;
    } else if (p.type == Q100_PH_ATTN || p.type == Q100_PH_COMBINE) {
      Q100_REQUIRE(p.qkv && p.cos_cache && p.sin_cache && p.position && p.k_cache && p.v_cache && p.att_partials && p.att_out,
                   Q100_EINVAL, "phase %d: NULL pointer argument", i);
      Q100_REQUIRE(p.head_dim == AT_D, Q100_EUNSUPPORTED, "phase %d: head_dim %d (the attention phase handles 128)", i, p.head_dim);
      Q100_REQUIRE(p.batch == M && p.heads > 0 && p.kv_heads > 0 && p.heads % p.kv_heads == 0, Q100_EINVAL,
                   "phase %d: batch must equal M, heads %% kv_heads == 0", i);
      const int rep = p.heads / p.kv_heads;
      Q100_REQUIRE(rep == 1 || rep == 2 || rep % 4 == 0, Q100_EUNSUPPORTED, "phase %d: heads / kv_heads = %d", i, rep);
      Q100_REQUIRE(p.max_len > 0 && p.rope_len > 0, Q100_EINVAL, "phase %d: bad cache length", i);
      AttnArgs& t = c.at;
      t.qkv = p.qkv; t.cosb = p.cos_cache; t.sinb = p.sin_cache; t.pos = reinterpret_cast<const int64_t*>(p.position);
      t.k_cache = p.k_cache; t.v_cache = p.v_cache; t.part = reinterpret_cast<float*>(p.att_partials); t.out = p.att_out;
      t.B = p.batch; t.H = p.heads; t.KVH = p.kv_heads;
      t.chunk_pos = chain_attn_chunk(p.max_len);
      t.max_parts = static_cast<int>((p.max_len + t.chunk_pos - 1) / t.chunk_pos);
      t.max_len = p.max_len; t.rope_len = p.rope_len;
      t.scale = 1.0f / sqrtf(static_cast<float>(AT_D));
      t.err = reinterpret_cast<uint32_t*>(desc) + CH_ERR_WORD;
    } else {
      Q100_REQUIRE(false, Q100_EINVAL, "phase %d: unknown type %d", i, p.type);
    }
  }
  if (workspace_needed) *workspace_needed = need;
  Q100_REQUIRE(workspace != nullptr && workspace_bytes >= need, Q100_EWORKSPACE,
               "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Q100_CUDA(cudaMemsetAsync(desc, 0, CH_HEADER_BYTES, st));
  Q100_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(desc) + CH_POS_SLOT_OFF, &pos_ptr, sizeof(pos_ptr),
                            cudaMemcpyHostToDevice, st));
  Q100_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(desc) + CH_HEADER_BYTES, host.data(),
                            static_cast<size_t>(n_phases) * sizeof(ChainPhase), cudaMemcpyHostToDevice, st));
  Q100_CUDA(cudaStreamSynchronize(st));                    // `host` goes out of scope
  return Q100_OK;
}

extern "C" int q100_chain_run(void* desc, int n_phases, int M, int group, int n_bits, int disable_zp, int dtype,
                              void* workspace, void* stream) {
  Q100_REQUIRE(desc != nullptr && workspace != nullptr && n_phases >= 1 && n_phases <= CH_MAX_PHASES, Q100_EINVAL,
               "bad chain arguments");
  Q100_REQUIRE(M >= 1 && M <= 16 && group > 0 && group % SK_KB == 0, Q100_EUNSUPPORTED, "chain: 1 <= M <= 16, group %% 128 == 0");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ChainGeom geo{};
  rc = chain_dispatch(0, dtype, n_bits, M, nullptr, nullptr, geo, st);
  if (rc != Q100_OK) return rc;
  SkinnyArgs a{};
  chain_common_args(a, M, group, n_bits, disable_zp, geo, workspace);
  a.total_units = 1; a.kblocks = 1;
  ChainArgs ch{};
  ch.phases = reinterpret_cast<const ChainPhase*>(reinterpret_cast<uint8_t*>(desc) + CH_HEADER_BYTES);
  ch.n_phases = n_phases;
  ch.bars = reinterpret_cast<uint32_t*>(desc);
  ch.err = reinterpret_cast<uint32_t*>(desc) + CH_ERR_WORD;
  ch.pos_pp = reinterpret_cast<const int64_t* const*>(reinterpret_cast<uint8_t*>(desc) + CH_POS_SLOT_OFF);
  static const bool want_trace = []() { const char* e = getenv("Q100_CHAIN_TRACE"); return e && e[0] == '1'; }();
  ch.trace = want_trace ? reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(desc) + CH_TRACE_OFF) : nullptr;
  static const int trace_phase = []() { const char* e = getenv("Q100_CHAIN_TRACE_PHASE"); return e ? atoi(e) : -1; }();
  ch.trace_phase = trace_phase;
  return chain_dispatch(1, dtype, n_bits, M, &a, &ch, geo, st);
}

// development hook: the per-CTA stamps of the traced phase, [512][8] uint64
extern "C" __attribute__((visibility("default"))) int q100_debug_chain_cta_trace(void* desc, unsigned long long* out, void* stream) {
  Q100_REQUIRE(desc && out, Q100_EINVAL, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Q100_CUDA(cudaMemcpyAsync(out, reinterpret_cast<uint8_t*>(desc) + CH_CTA_TRACE_OFF, 512 * 64, cudaMemcpyDeviceToHost, st));
  Q100_CUDA(cudaStreamSynchronize(st));
  return Q100_OK;
}

// development hook (not part of include/q100.h): barrier time stamps of the last traced launch, [n_phases][3] uint64;
// re-arms the trace (first-arrival slots back to UINT64_MAX, the others to 0)
extern "C" __attribute__((visibility("default"))) int q100_debug_chain_trace(void* desc, int n_phases, unsigned long long* out,
                                                                             void* stream) {
  Q100_REQUIRE(desc && out && n_phases >= 1 && n_phases <= CH_MAX_PHASES, Q100_EINVAL, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* tr = reinterpret_cast<uint8_t*>(desc) + CH_TRACE_OFF;
  Q100_CUDA(cudaMemcpyAsync(out, tr, 24ull * n_phases, cudaMemcpyDeviceToHost, st));
  Q100_CUDA(cudaStreamSynchronize(st));
  {
    unsigned long long c[4] = {0, 0, 0, 0};               // (cycles, ns) at the first and the last phase
    Q100_CUDA(cudaMemcpy(c, tr + 24ull * (CH_MAX_PHASES - 2), sizeof(c), cudaMemcpyDeviceToHost));
    if (c[3] > c[1]) fprintf(stderr, "[q100 chain trace] SM clock over the traced launch: %.0f MHz (%llu cycles in %llu ns)\n",
                             1e3 * static_cast<double>(c[2] - c[0]) / static_cast<double>(c[3] - c[1]), c[2] - c[0], c[3] - c[1]);
  }
  std::vector<unsigned long long> init(3 * static_cast<size_t>(n_phases), 0ull);
  for (int i = 0; i < n_phases; ++i) init[3 * i] = ~0ull;
  Q100_CUDA(cudaMemcpyAsync(tr, init.data(), 24ull * n_phases, cudaMemcpyHostToDevice, st));
  Q100_CUDA(cudaStreamSynchronize(st));
  return Q100_OK;
}

extern "C" int q100_chain_error(const void* desc, void* stream) {
  Q100_REQUIRE(desc != nullptr, Q100_EINVAL, "NULL descriptor");
  uint32_t v = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Q100_CUDA(cudaMemcpyAsync(&v, reinterpret_cast<const uint32_t*>(desc) + CH_ERR_WORD, 4, cudaMemcpyDeviceToHost, st));
  Q100_CUDA(cudaStreamSynchronize(st));
  Q100_REQUIRE(v == 0, Q100_ECUDA, "a chain launch reported an error (grid barrier time-out or a position outside the KV cache)");
  return Q100_OK;
}
