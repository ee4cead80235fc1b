// q100_gemm_rows.cu -- K3r: decode-shape dequant-GEMM (M <= 8, W4 g128) in which no output row is ever shared by
// two CTAs.
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// Why a second decode kernel: K3a (q100_gemm_skinny.cu) balances any N x K over the grid by cutting the work stream-K
// style into [128 rows x 128 k] units, and pays for that with a fix-up through global memory: per launch the reducers
// finish 3-4 us after the median CTA (DESIGN.md "K3a", time stamps in profiles/), which is more than the weights of a
// decoder block's smaller linears need at the HBM roofline (o_proj: 1.4 us).  Here the BYTES are balanced instead:
//   * CTA c of a grid of G = SMs x occupancy owns the contiguous rows [4 floor(c Q / G), 4 floor((c+1) Q / G)),
//     Q = N / 4: every CTA streams N / G rows (+- 4) of ALL of K, so a row's sum is finished inside one CTA and the
//     launch ends when the last byte has arrived -- no partials, no flags, no second pass;
//   * the weights go global -> registers directly (ld.global.nc.L1::no_allocate.v4): the unit of work of a warp is a
//     block of 16 rows x 128 k -- per lane the 32 codes of rows g and g + 8 it feeds to the tensor core anyway, two
//     16-byte loads that each cover whole 32-byte sectors of the tile-major layout, plus the rows' (scale, zero) pairs --
//     and every warp keeps a ring of D blocks in flight in registers.  Bulk copies cannot do this: a row range of a
//     [128 x 64] tile is <= 1 KB, and cp.async.bulk costs ~100 cycles per operation whatever its size
//     (tools/micro/stream_bench5.cu: 1.9 TB/s).  The weights are static, so the first D blocks of every warp are
//     requested BEFORE griddepcontrol.wait: with programmatic dependent launch they travel while the previous kernel drains;
//   * the rows are walked in passes of <= 32 rows; the eight warps split a pass by (16-row slab, 128-k group): warp w
//     takes slab w & 1 and groups (w >> 1) + 4 j (passes of <= 16 rows: slab 0, groups w + 8 j).  The inner step is
//     K3a's: codes -> tensor core (M <= 2: unsigned bytes x three signed 8-bit limbs of x in block fixed point on
//     mma.sync.m16n8k32, exact integer contraction; M <= 8: fp16 "magic form" 1024 + q on mma.sync.m16n8k16), one fp32
//     group step acc += scale (D - zp S) per 128 k;
//   * at the end of a pass the warps' fragments meet in shared memory and are summed in a fixed order (deterministic),
//     bias is added and y is written with coalesced stores.
// x is staged in shared memory by the CTA (all of K at once when it fits, else in chunks inside every pass).
#include <mutex>
#include <type_traits>
#include <cstdlib>
#include <cstdio>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

namespace {

constexpr int RW_TROW = 32;                       // bytes per row of a 64-k tile at W4
constexpr int RW_TILE_BYTES = 128 * RW_TROW;      // one [128 x 64] tile in HBM
constexpr int RW_XPITCH = 320;                    // fp16 form: bytes per (token, 128-k block) of x: 256 + 64 (sums)
constexpr int RW_SLOT = 1280;                     // a warp's ring slot: 32 lanes x (16 + 16 bytes of codes, 4 + 4 of side band)
constexpr int RW_ICOL = 144;                      // integer form: bytes per (block, limb column): 128 + 16 (sum, coefficient)

struct RowsArgs {
  const void* x;          // [M, K] T   (x_gate_up: [M, 2K] = gate | up)
  const void* packed;     // tile-major packed codes (layout v2)
  const void* meta;       // [K / 128, N, 2] T
  const void* bias;       // [N] T or null
  void* y;                // [M, N] T
  int M, N, K;
  int kblocks;            // K / 128 (a multiple of 8)
  int has_zp;
  float qmin;
  int x_gate_up;
  // block glue folded into the GEMM (q100_gemm_wna16_block): x = OmniLlamaRMSNorm(x) while it is staged (K4's
  // arithmetic, quant/omni_norm.py:52-63), y = res + y in the epilogue (models/int_llama_layer.py:246, 251)
  const void* norm_w;     // [K] T or null
  float eps;
  const void* res;        // [M, N] T or null
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// 16 / 4 bytes global -> shared, asynchronous
// (no L2 cache hint: with the policy operand ptxas 12.9 put the descriptor of the LDGSTS inside the main loop into an
// odd uniform register pair, desc[UR1], and the launch died with "illegal instruction" as soon as that path ran)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
template <uint32_t MASK>
__device__ __forceinline__ uint32_t lop_and_or(uint32_t w, uint32_t orv) {   // (w & MASK) | orv as ONE LOP3
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "n"(MASK), "r"(orv));
  return d;
}
__device__ __forceinline__ uint32_t reg_const(uint32_t v) {
  uint32_t d;
  asm volatile("mov.b32 %0, %1;" : "=r"(d) : "r"(v));
  return d;
}
// shifts on the FMA pipe (IMAD.HI): LOP3 / SHF share the ALU pipe with the extraction
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
__device__ __forceinline__ void imma_16832(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// act_fn(gate) * up for eight elements, K6's arithmetic (models/int_llama_layer.py:45 on 16-bit tensors)
template <typename T>
__device__ __forceinline__ uint4 silu_mul8(const uint4& g, const uint4& u) {
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

// ---------------------------------------------------------------- the kernel
// Shared memory: red[2][32 warps][4][32] fp32 | (1 KB spare) | the warps' rings [32][D][RW_SLOT] |
// stab[kblocks][8][2] fp32 (zero-point sum, weight per limb column) | x limbs [kblocks][4 M][RW_ICOL]
constexpr int RW_MISC_BYTES = 1024;
constexpr int RW_XU = 2;                          // loads of x in flight per thread while it is staged

// mma row i of a slab is slab row 2 i (i < 8) or 2 (i - 8) + 1: the two rows a lane feeds (mma rows g, g + 8) are the
// ADJACENT rows 2 g, 2 g + 1 -- one code pointer (the rows are 32 bytes apart inside a tile) and one 8-byte side-band read
// LOGW: log2 of the warps per CTA -- 5: one CTA per SM; 4: two (the next launch's CTAs are co-resident with this one's tail)
template <typename T, int D, bool HAS_ZP, int LOGW>
__global__ void __launch_bounds__(32 << LOGW, LOGW == 5 ? 1 : 2) k3_rows(const RowsArgs a) {
  constexpr int RW_WARPS = 1 << LOGW, RW_THREADS = 32 << LOGW;
  constexpr int RW_RED_BYTES = 2 * RW_WARPS * 4 * 32 * 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  float* red = reinterpret_cast<float*>(smem_raw);
  const uint32_t wring = smem_u32(smem_raw) + RW_RED_BYTES + RW_MISC_BYTES + warp * (D * RW_SLOT) + lane * 16;
  uint8_t* stab_ptr = smem_raw + RW_RED_BYTES + RW_MISC_BYTES + RW_WARPS * D * RW_SLOT;
  const int kblocks = a.kblocks;
  uint8_t* xs_ptr = stab_ptr + kblocks * 64;
  const int mrows = a.M;
  const int blk_pitch = 4 * mrows * RW_ICOL;      // bytes of x limbs per k block

  const int G = gridDim.x, cta = blockIdx.x;
  const int Q = a.N >> 2;
  const int R0 = 4 * static_cast<int>(static_cast<uint32_t>(cta) * static_cast<uint32_t>(Q) / static_cast<uint32_t>(G));
  const int nr = 4 * static_cast<int>(static_cast<uint32_t>(cta + 1) * static_cast<uint32_t>(Q) / static_cast<uint32_t>(G)) - R0;
  pdl_launch_dependents();
  if (nr == 0) return;

  // ---- geometry of this CTA: its ns = ceil(nr / 16) slabs of 16 rows are walked in passes of S = W / 4 slabs and one
  // last pass of W / 4, W / 8 or W / 16; in a pass of S slabs the W warps split the work by (slab, k group mod KW),
  // KW = W / S
  constexpr int SHM = LOGW - 2;                   // log2 of the slabs of a full pass
  const int ns = (nr + 15) >> 4;
  const int n8 = ns >> SHM, rem = ns & ((1 << SHM) - 1);
  const int np = n8 + (rem ? 1 : 0);
  const int sh_last = rem > (1 << (SHM - 1)) ? SHM : rem > (1 << (SHM - 2)) ? SHM - 1 : SHM - 2;
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(a.packed);
  const uint8_t* meta = reinterpret_cast<const uint8_t*>(a.meta);
  const int64_t ktile_row = static_cast<int64_t>(kblocks) * 2 * RW_TILE_BYTES;   // bytes of one 128-row tile row

  // ---- the load iterator: row pairs past the CTA's range are clamped to its last pair (same sectors, results unused)
  const uint8_t* pc0 = packed;                    // this lane's 16 bytes of row 2g of the slab, current k group
  const uint8_t* pm0 = meta;                      // (scale, zero) of rows 2g, 2g + 1
  int cstep = 0;                                  // bytes to this warp's next k group (< 2^31)
  int64_t mstep = 0;
  int lj = 0, lp = 0, lbpp = 0;
  auto setup_pass = [&](int p) {
    const int sh_ = p < n8 ? SHM : sh_last;
    const int kw = RW_WARPS >> sh_, kq = warp >> sh_;
    const int rel = (((p << SHM) + (warp & ((1 << sh_) - 1))) << 4) + 2 * g;
    const int r0 = R0 + min(rel, nr - 2);
    pc0 = packed + (r0 >> 7) * ktile_row + (r0 & 127) * RW_TROW + (t >> 1) * RW_TILE_BYTES + (t & 1) * 16 +
          kq * 2 * RW_TILE_BYTES;
    pm0 = meta + (static_cast<int64_t>(kq) * a.N + r0) * 4;
    cstep = kw * 2 * RW_TILE_BYTES;
    mstep = static_cast<int64_t>(kw) * a.N * 4;
    lbpp = kblocks >> (LOGW - sh_);
  };
  auto issue = [&](uint32_t slot_addr) {          // one cp.async group: the block under the iterator (or nothing)
    if (lp < np) {
      cp_async16(slot_addr, pc0);
      cp_async16(slot_addr + 512, pc0 + RW_TROW);
      cp_async8(slot_addr + 1024 - lane * 8, pm0);
      pc0 += cstep; pm0 += mstep;
      if (__builtin_expect(++lj == lbpp, 0)) { lj = 0; ++lp; if (lp < np) setup_pass(lp); }
    }
    cp_async_commit();
  };
  setup_pass(0);
#pragma unroll 1
  for (int d = 0; d < D - 1; ++d) issue(wring + d * RW_SLOT);   // static data: in flight before the previous kernel ends

  // ===================== x: block fixed point, four limbs =====================
  // Per (token, 128-k block): xi = rint(x 2^(25 - e)), e = exponent of the block's largest magnitude, as four balanced
  // base-256 digits, one mma column each.  The elements that meet the HIGH nibbles of the packed words (e2, e3, e6, e7
  // of every 8) are scaled by another 1/16 -- rint(x 2^(21 - e)), the precision K3a gives every element -- because the
  // weight operand for them is (w & 0xf0f0f0f0) = 16 q: no shift in the inner loop.  stab holds, per k block and
  // column, (sum_lo l + 16 sum_hi l, 256^l 2^(e - 25)): the factor of the weight zero-point and the column's weight.
  pdl_wait();                                     // x and y belong to the stream order
  {
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int cpt = a.K >> 3;                     // 16-byte chunks per token
    const int total_c = mrows * cpt;
    for (int i = tid; i < kblocks * 16; i += RW_THREADS)   // columns past 4 M: zero sums and weights
      if ((i & 15) >= 8 * mrows) reinterpret_cast<float*>(stab_ptr)[i] = 0.f;
    const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // the 16 chunks of one (token, block)
    // one chunk: block maximum -> exponent -> four limbs per element -> shared memory
    auto convert = [&](const uint4& v, int c) {
      const int m = c >= cpt ? 1 : 0;
      const int kc = c - m * cpt;                 // chunk inside the token
      const int bl = kc >> 4, ch = kc & 15;
      const float2 p0 = DT<T>::to_f2(v.x), p1 = DT<T>::to_f2(v.y), p2 = DT<T>::to_f2(v.z), p3 = DT<T>::to_f2(v.w);
      uint32_t ab = max(max(__float_as_uint(p0.x) & 0x7fffffffu, __float_as_uint(p0.y) & 0x7fffffffu),
                        max(__float_as_uint(p1.x) & 0x7fffffffu, __float_as_uint(p1.y) & 0x7fffffffu));
      ab = max(ab, max(max(__float_as_uint(p2.x) & 0x7fffffffu, __float_as_uint(p2.y) & 0x7fffffffu),
                       max(__float_as_uint(p3.x) & 0x7fffffffu, __float_as_uint(p3.y) & 0x7fffffffu)));
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) ab = max(ab, __shfl_xor_sync(half_mask, ab, d));   // finite < inf < NaN
      const int e = static_cast<int>(ab >> 23);
      const int fe = min(max(279 - e, 5), 250);            // scale = 2^(fe - 127) = 2^(25 - (e - 127))
      const float sc_lo = __uint_as_float(static_cast<uint32_t>(fe) << 23);
      const float sc_hi = __uint_as_float(static_cast<uint32_t>(fe - 4) << 23);
      const float f[8] = {p0.x * sc_lo, p0.y * sc_lo, p1.x * sc_hi, p1.y * sc_hi,
                          p2.x * sc_lo, p2.y * sc_lo, p3.x * sc_hi, p3.y * sc_hi};
      int l0[8], l1[8], l2[8], l3[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int xi = __float2int_rn(f[i]);
        l0[i] = (xi << 24) >> 24;
        const int x1 = (xi - l0[i]) >> 8;
        l1[i] = (x1 << 24) >> 24;
        const int x2 = (x1 - l1[i]) >> 8;
        l2[i] = (x2 << 24) >> 24;
        l3[i] = (x2 - l2[i]) >> 8;
      }
      int s0 = (l0[0] + l0[1] + l0[4] + l0[5]) + 16 * (l0[2] + l0[3] + l0[6] + l0[7]);
      int s1 = (l1[0] + l1[1] + l1[4] + l1[5]) + 16 * (l1[2] + l1[3] + l1[6] + l1[7]);
      int s2 = (l2[0] + l2[1] + l2[4] + l2[5]) + 16 * (l2[2] + l2[3] + l2[6] + l2[7]);
      int s3 = (l3[0] + l3[1] + l3[4] + l3[5]) + 16 * (l3[2] + l3[3] + l3[6] + l3[7]);
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        s0 += __shfl_xor_sync(half_mask, s0, d);
        s1 += __shfl_xor_sync(half_mask, s1, d);
        s2 += __shfl_xor_sync(half_mask, s2, d);
        s3 += __shfl_xor_sync(half_mask, s3, d);
      }
      auto pk = [](int b0, int b1, int b2, int b3) {
        return (static_cast<uint32_t>(b0) & 0xffu) | ((static_cast<uint32_t>(b1) & 0xffu) << 8) |
               ((static_cast<uint32_t>(b2) & 0xffu) << 16) | (static_cast<uint32_t>(b3) << 24);
      };
      // bytes in the order in which the nibbles come out of a packed word: (e0, e4, e1, e5 | e2, e6, e3, e7)
      uint8_t* colp = xs_ptr + bl * blk_pitch + (4 * m) * RW_ICOL + (ch >> 2) * 32 + (ch & 3) * 8;
      *reinterpret_cast<uint2*>(colp) = make_uint2(pk(l0[0], l0[4], l0[1], l0[5]), pk(l0[2], l0[6], l0[3], l0[7]));
      *reinterpret_cast<uint2*>(colp + RW_ICOL) = make_uint2(pk(l1[0], l1[4], l1[1], l1[5]), pk(l1[2], l1[6], l1[3], l1[7]));
      *reinterpret_cast<uint2*>(colp + 2 * RW_ICOL) = make_uint2(pk(l2[0], l2[4], l2[1], l2[5]), pk(l2[2], l2[6], l2[3], l2[7]));
      *reinterpret_cast<uint2*>(colp + 3 * RW_ICOL) = make_uint2(pk(l3[0], l3[4], l3[1], l3[5]), pk(l3[2], l3[6], l3[3], l3[7]));
      if (ch == 0) {
        // column weight 256^l / scale; inf / NaN anywhere in the block poisons the token's result
        const uint32_t cexp = static_cast<uint32_t>(254 - fe);
        const bool bad = e == 255;
        float4* sp = reinterpret_cast<float4*>(stab_ptr + bl * 64 + m * 32);
        sp[0] = make_float4(static_cast<float>(s0), bad ? nan_f() : __uint_as_float(cexp << 23),
                            static_cast<float>(s1), bad ? nan_f() : __uint_as_float((cexp + 8) << 23));
        sp[1] = make_float4(static_cast<float>(s2), bad ? nan_f() : __uint_as_float((cexp + 16) << 23),
                            static_cast<float>(s3), bad ? nan_f() : __uint_as_float((cexp + 24) << 23));
      }
    };
    if (a.norm_w != nullptr) {
      // x_norm (host: at most two chunks per thread): the chunks stay in registers between the sum of squares and the
      // normalisation -- one sweep, one L2 round trip
      uint4 vq[2], vw[2];
      float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = tid + u * RW_THREADS;
        const int m = c >= cpt ? 1 : 0;
        if (c < total_c) {
          vq[u] = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(m) * a.K + (c - m * cpt) * 8);
          vw[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.norm_w) + (c - m * cpt) * 8);
        } else {
          vq[u] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = tid + u * RW_THREADS;
        const uint32_t w[4] = {vq[u].x, vq[u].y, vq[u].z, vq[u].w};
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = DT<T>::to_f2(w[i]); ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); }
        if (c >= cpt) ss1 += ss; else ss0 += ss;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, d);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, d);
      }
      float* ssq = reinterpret_cast<float*>(smem_raw + RW_RED_BYTES);          // [warps][2] in the spare kilobyte
      if (lane == 0) { ssq[warp * 2] = ss0; ssq[warp * 2 + 1] = ss1; }
      __syncthreads();
      ss0 = 0.f; ss1 = 0.f;
#pragma unroll
      for (int w = 0; w < RW_WARPS; ++w) { ss0 += ssq[w * 2]; ss1 += ssq[w * 2 + 1]; }     // fixed order: every thread the same bits
      const float rstd0 = rsqrtf(__fadd_rn(__fdiv_rn(ss0, static_cast<float>(a.K)), a.eps));
      const float rstd1 = rsqrtf(__fadd_rn(__fdiv_rn(ss1, static_cast<float>(a.K)), a.eps));
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = tid + u * RW_THREADS;
        if (c < total_c) convert(norm8<T>(vq[u], vw[u], c >= cpt ? rstd1 : rstd0), c);   // uniform per half-warp
      }
    } else {
#pragma unroll 1
      for (int cb = tid; cb < total_c; cb += RW_XU * RW_THREADS) {
        uint4 vq[RW_XU], vu[RW_XU];
#pragma unroll
        for (int u = 0; u < RW_XU; ++u) {         // the loads of a batch in flight together: one L2 round trip
          const int c = cb + u * RW_THREADS;
          if (c < total_c) {
            const int m = c >= cpt ? 1 : 0;
            const int kofs = (c - m * cpt) * 8;
            if (a.x_gate_up) {
              const T* gp = xg + static_cast<int64_t>(m) * (2 * a.K) + kofs;
              vq[u] = *reinterpret_cast<const uint4*>(gp);
              vu[u] = *reinterpret_cast<const uint4*>(gp + a.K);
            } else {
              vq[u] = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(m) * a.K + kofs);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < RW_XU; ++u) {
          const int c = cb + u * RW_THREADS;
          if (c >= total_c) break;                // uniform per half-warp: total_c is a multiple of 16
          convert(a.x_gate_up ? silu_mul8<T>(vq[u], vu[u]) : vq[u], c);
        }
      }
    }
    __syncthreads();
  }

  // ===================== the stream =====================
  const int ncols = 4 * mrows;
  const float nz_fix = a.qmin;
  // ring bookkeeping: D - 1 blocks are outstanding between steps; a step reads the oldest block and refills the slot
  // that was read one step ago
  const uint8_t* wring_g = smem_raw + (wring - smem_u32(smem_raw));       // generic address of this lane's ring
  uint32_t coff = 0, roff = (D - 1) * RW_SLOT;    // byte offsets of the read position / the refill position
  int parity = 0;
#pragma unroll 1
  for (int p = 0; p < np; ++p) {
    const int sh = p < n8 ? SHM : sh_last;
    const int S = 1 << sh, KW = RW_WARPS >> sh, kq = warp >> sh;
    const int bpp = kblocks >> (LOGW - sh);       // blocks per warp in this pass
    // mma column g is (token, limb) = divmod(g, 4); columns past the last one re-read it (never stored)
    uint32_t xo = static_cast<uint32_t>(xs_ptr - smem_raw) + min(g, ncols - 1) * RW_ICOL + t * 32 + kq * blk_pitch;
    uint32_t so = static_cast<uint32_t>(stab_ptr - smem_raw) + t * 16 + kq * 64;
    const uint32_t xstep = KW * blk_pitch, sstep = KW * 64;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    // this thread's element of the residual (it finishes output (row tid mod 16 S, token tid / 16 S) of the pass):
    // requested now, needed after the pass
    float resv = 0.f;
    if (a.res != nullptr && tid < (16 << sh) * mrows) {
      const int rel_o = (p << (SHM + 4)) + (tid & ((16 << sh) - 1));
      if (rel_o < nr)
        resv = DT<T>::to_f(reinterpret_cast<const T*>(a.res)[static_cast<int64_t>(tid >> (4 + sh)) * a.N + R0 + rel_o]);
    }
#pragma unroll 1
    for (int j = 0; j < bpp; ++j) {
      cp_async_wait<D - 2>();                     // this lane's share of the oldest block has landed (it reads only that)
      const uint8_t* sp = wring_g + coff;
      const uint4 c0 = *reinterpret_cast<const uint4*>(sp);
      const uint4 c1 = *reinterpret_cast<const uint4*>(sp + 512);
      const uint2 mm = *reinterpret_cast<const uint2*>(sp + 1024 - lane * 8);
      const uint4 xb0 = *reinterpret_cast<const uint4*>(smem_raw + xo);
      const uint4 xb1 = *reinterpret_cast<const uint4*>(smem_raw + xo + 16);
      const float4 sv = *reinterpret_cast<const float4*>(smem_raw + so);      // (zero-point sum, weight) of columns 2t, 2t + 1
      issue(wring + roff);
      roff = coff;
      coff += RW_SLOT;
      if (coff == D * RW_SLOT) coff = 0;
      int di[4] = {0, 0, 0, 0};
      uint32_t af[4];
      af[0] = c0.x & 0x0f0f0f0fu; af[1] = c1.x & 0x0f0f0f0fu; af[2] = c0.x & 0xf0f0f0f0u; af[3] = c1.x & 0xf0f0f0f0u;
      imma_16832(di, af, xb0.x, xb0.y);
      af[0] = c0.y & 0x0f0f0f0fu; af[1] = c1.y & 0x0f0f0f0fu; af[2] = c0.y & 0xf0f0f0f0u; af[3] = c1.y & 0xf0f0f0f0u;
      imma_16832(di, af, xb0.z, xb0.w);
      af[0] = c0.z & 0x0f0f0f0fu; af[1] = c1.z & 0x0f0f0f0fu; af[2] = c0.z & 0xf0f0f0f0u; af[3] = c1.z & 0xf0f0f0f0u;
      imma_16832(di, af, xb1.x, xb1.y);
      af[0] = c0.w & 0x0f0f0f0fu; af[1] = c1.w & 0x0f0f0f0fu; af[2] = c0.w & 0xf0f0f0f0u; af[3] = c1.w & 0xf0f0f0f0u;
      imma_16832(di, af, xb1.z, xb1.w);
      const float2 sz0 = DT<T>::to_f2(mm.x), sz1 = DT<T>::to_f2(mm.y);     // (scale, zero) of rows 2g, 2g + 1
      const float nz0 = HAS_ZP ? -sz0.y : nz_fix, nz1 = HAS_ZP ? -sz1.y : nz_fix;
      acc0 = fmaf(sz0.x * sv.y, fmaf(nz0, sv.x, static_cast<float>(di[0])), acc0);
      acc1 = fmaf(sz0.x * sv.w, fmaf(nz0, sv.z, static_cast<float>(di[1])), acc1);
      acc2 = fmaf(sz1.x * sv.y, fmaf(nz1, sv.x, static_cast<float>(di[2])), acc2);
      acc3 = fmaf(sz1.x * sv.w, fmaf(nz1, sv.z, static_cast<float>(di[3])), acc3);
      xo += xstep; so += sstep;
    }
    // ---- end of the pass: the warps' fragments meet in shared memory, fixed summation order
    float* rp = red + parity * (RW_WARPS * 4 * 32);
    parity ^= 1;
    rp[(warp * 4 + 0) * 32 + lane] = acc0; rp[(warp * 4 + 1) * 32 + lane] = acc1;
    rp[(warp * 4 + 2) * 32 + lane] = acc2; rp[(warp * 4 + 3) * 32 + lane] = acc3;
    __syncthreads();
    const int PRW = 16 * S;
    if (tid < PRW * mrows) {
      const int rr = tid & (PRW - 1), m = tid >> (4 + sh);
      const int rel = (p << (SHM + 4)) + rr;
      if (rel < nr) {
        const int sl = rr >> 4, r16 = rr & 15;
        float sum = 0.f;
        for (int kw = 0; kw < KW; ++kw) {
          // slab row r16 = mma row (r16 >> 1) + 8 (r16 & 1): lane quarter-group g = r16 >> 1, registers 2 (r16 & 1), + 1
          const float* wp = rp + ((sl + S * kw) * 4 + ((r16 & 1) << 1)) * 32 + (r16 >> 1) * 4;
#pragma unroll
          for (int l = 0; l < 4; ++l) {           // the four limb columns of token m: 4 m + l
            const int c = 4 * m + l;
            sum += wp[(c & 1) * 32 + ((c & 7) >> 1)];
          }
        }
        const int n = R0 + rel;
        if (a.bias) sum += DT<T>::to_f(reinterpret_cast<const T*>(a.bias)[n]);
        if (a.res != nullptr) sum = __fadd_rn(resv, DT<T>::rd(sum));   // residual + r(y), rounded to T: K4's add on two 16-bit tensors
        reinterpret_cast<T*>(a.y)[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(sum);
      }
    }
  }
}

struct RowsPlan { int logw, depth, smem; };

// two CTAs of 16 warps per SM when everything fits twice (the next launch's CTAs then start -- and request their first
// blocks -- while this launch is still running), else one CTA of 32 warps
bool plan_rows(int M, int K, RowsPlan& p) {
  const int kblocks = K / 128;
  const int fixed0 = RW_MISC_BYTES + kblocks * 64 + kblocks * 4 * M * RW_ICOL + 128;
  static const int env_depth = []() { const char* e = getenv("Q100_ROWS_DEPTH"); return e ? atoi(e) : 0; }();
  static const int env_logw = []() { const char* e = getenv("Q100_ROWS_LOGW"); return e ? atoi(e) : 0; }();
  for (int logw = 4; logw <= 5; ++logw) {
    if (env_logw && logw != env_logw) continue;
    const int budget = logw == 4 ? (233472 - 2 * 1024) / 2 : 232448 - 1024;
    const int fixed = fixed0 + 2 * (1 << logw) * 4 * 32 * 4;
    const int depths[3] = {5, 4, 3};
    for (int di = 0; di < 3; ++di) {
      const int depth = depths[di];
      if (env_depth > 0 && depth > env_depth) continue;
      if (logw == 4 && depth < 4) continue;
      const int smem = fixed + (1 << logw) * depth * RW_SLOT;
      if (smem <= budget) { p = RowsPlan{logw, depth, smem}; return true; }
    }
  }
  return false;
}

template <typename T, int D, int LOGW>
int launch_rows_t(const RowsArgs& a, const RowsPlan& p, int pdl, cudaStream_t st) {
  auto kern = a.has_zp ? k3_rows<T, D, true, LOGW> : k3_rows<T, D, false, LOGW>;
  static std::mutex mu;
  static int max_set[2][kMaxDevices] = {};
  {
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (p.smem > max_set[a.has_zp ? 1 : 0][dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));
      max_set[a.has_zp ? 1 : 0][dev] = p.smem;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sm_count() * (LOGW == 4 ? 2 : 1));
  cfg.blockDim = dim3(32 << LOGW);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, a), "k3_rows launch");
}

template <typename T>
int launch_rows(const RowsArgs& a, const RowsPlan& p, int pdl, cudaStream_t st) {
  if (p.logw == 4) return p.depth == 5 ? launch_rows_t<T, 5, 4>(a, p, pdl, st) : launch_rows_t<T, 4, 4>(a, p, pdl, st);
  switch (p.depth) {
    case 5: return launch_rows_t<T, 5, 5>(a, p, pdl, st);
    case 4: return launch_rows_t<T, 4, 5>(a, p, pdl, st);
    default: return launch_rows_t<T, 3, 5>(a, p, pdl, st);
  }
}

}  // namespace

// mode: 2 = integer form (M <= 2); the fp16 form (mode 1, M <= 8) stays on K3a
bool gemm_rows_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits, int mode) {
  static const bool off = []() { const char* e = getenv("Q100_GEMM_ROWS"); return e && atoi(e) == 0; }();
  if (off) return false;
  if (mode != 2 || n_bits != 4 || group != 128 || K % 2048 != 0 || N % 8 != 0 || M < 1 || M > 2) return false;
  if (N > (1 << 24) || K > (1 << 20)) return false;
  RowsPlan p;
  return plan_rows(static_cast<int>(M), static_cast<int>(K), p);
}

int gemm_rows(const void* x, const void* packed, const void* meta, const void* bias, void* y, int64_t M, int64_t N,
              int64_t K, int disable_zp, int dtype, int mode, int pdl, int x_gate_up, cudaStream_t st,
              const void* norm_w, float eps, const void* res) {
  (void)mode;
  RowsPlan p;
  Q100_REQUIRE(plan_rows(static_cast<int>(M), static_cast<int>(K), p), Q100_EUNSUPPORTED,
               "k3_rows: no shared-memory plan for M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  Q100_REQUIRE(norm_w == nullptr || M * K / 8 <= 2 * (32 << p.logw), Q100_EUNSUPPORTED,
               "k3_rows: RMSNorm in the x staging holds x in registers: M * K <= %d (got M=%lld K=%lld)",
               16 * (32 << p.logw), (long long)M, (long long)K);
  RowsArgs a{};
  a.x = x; a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K);
  a.kblocks = static_cast<int>(K / 128);
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -8.0f : 0.0f;
  a.x_gate_up = x_gate_up;
  a.norm_w = norm_w; a.eps = eps; a.res = res;
  if (dtype == Q100_F16) return launch_rows<__half>(a, p, pdl, st);
  return launch_rows<__nv_bfloat16>(a, p, pdl, st);
}

}  // namespace q100
