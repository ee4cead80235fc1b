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
#include <cstdlib>
#include <cstdio>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

namespace {

constexpr int RW_WARPS = 8;
constexpr int RW_THREADS = RW_WARPS * 32;
constexpr int RW_PR = 32;                         // rows per pass
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
  int kblocks;            // K / 128
  int chunk;              // k blocks of x staged at a time (a multiple of 8; >= kblocks: everything, once)
  int has_zp;
  float qmin;
  int x_gate_up;
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
// 16 / 4 bytes global -> shared, asynchronous; src_bytes = 0 writes zeros without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;"
               :: "r"(dst), "l"(src), "r"(src_bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
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

struct BlockRegs { uint4 c0, c1; uint32_t m0, m1; };   // codes of rows g / g + 8 (32 k each) and their (scale, zero)

// MODE 1: fp16 magic-form operand (M <= 8);  MODE 2: integer operand, three limbs per token (M <= 2)
// D: blocks a warp keeps in flight;  OCC: CTAs per SM the instantiation is compiled for
template <typename T, int MODE, int D, int OCC>
__global__ void __launch_bounds__(RW_THREADS, OCC) k3_rows(const RowsArgs a) {
  constexpr bool IMMA = MODE == 2;
  constexpr bool kConvX = !IMMA && DT<T>::kCode == Q100_BF16;    // bf16 activations -> scaled fp16 at staging
  constexpr int RED_FLOATS = RW_WARPS * 4 * 32;                  // one parity of the cross-warp reduction buffer
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw);
  const uint32_t wring = smem_u32(smem_raw) + 2 * RED_FLOATS * 4 + (threadIdx.x >> 5) * (D * RW_SLOT);   // this warp's ring
  uint8_t* xs_ptr = smem_raw + 2 * RED_FLOATS * 4 + RW_WARPS * D * RW_SLOT;
  const uint32_t xs_u32 = smem_u32(xs_ptr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int G = gridDim.x, cta = blockIdx.x;
  const int Q = a.N >> 2;
  const int q0 = static_cast<int>(static_cast<int64_t>(cta) * Q / G);
  const int q1 = static_cast<int>(static_cast<int64_t>(cta + 1) * Q / G);
  const int R0 = 4 * q0, nr = 4 * (q1 - q0);
  const int np = (nr + RW_PR - 1) / RW_PR;
  const int kblocks = a.kblocks;
  pdl_launch_dependents();
  if (nr == 0) return;

  // ---- this warp's sequence of blocks: pass p (rows R0 + 32 p ...), j-th k group of the warp inside the pass
  const int last_pr = nr - (np - 1) * RW_PR;
  const bool last_two = last_pr > 16;
  const int bpp2 = (kblocks + 3) >> 2, bpp1 = (kblocks + 7) >> 3;      // blocks per pass and warp (two slabs / one slab)
  const int ktiles = kblocks * 2;
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(a.packed);
  const uint8_t* meta = reinterpret_cast<const uint8_t*>(a.meta);
  const int64_t mrow_pitch = static_cast<int64_t>(a.N) * 4;           // bytes between the side-band rows of two groups

  // the block the load iterator points at
  int lp = 0, lj = 0;                             // pass, index inside the pass
  const uint8_t* pc0 = nullptr;                   // codes of row g of the slab: this lane's 16 bytes of k group kb
  const uint8_t* pc1 = nullptr;                   // ... of row g + 8
  const uint8_t* pm0 = nullptr;                   // (scale, zero) of row g, group kb
  bool v0 = false, v1 = false;                    // the rows exist (inside this CTA's range)
  int lstride = 4, lbpp = bpp2, lkb = 0;
  auto setup_pass = [&](int p) {
    const bool two = (p < np - 1) || last_two;
    const int slab = two ? (warp & 1) : 0;
    lkb = two ? (warp >> 1) : warp;
    lstride = two ? 4 : 8;
    lbpp = two ? bpp2 : bpp1;
    const int rel = p * RW_PR + slab * 16 + g;    // row inside the CTA's range
    v0 = rel < nr; v1 = rel + 8 < nr;
    const int r0 = R0 + rel, r1 = r0 + 8;
    const int lo = (t >> 1) * RW_TILE_BYTES + (t & 1) * 16;
    pc0 = packed + (static_cast<int64_t>(r0 >> 7) * ktiles + 2 * lkb) * RW_TILE_BYTES + (r0 & 127) * RW_TROW + lo;
    pc1 = packed + (static_cast<int64_t>(r1 >> 7) * ktiles + 2 * lkb) * RW_TILE_BYTES + (r1 & 127) * RW_TROW + lo;
    pm0 = meta + lkb * mrow_pitch + static_cast<int64_t>(r0) * 4;
  };
  const uint64_t pol = policy_evict_first();
  uint32_t islot = 0;                             // ring slot the next request lands in
  auto issue_next = [&]() {                       // request the block under the load iterator (one cp.async group), advance it
    if (lp < np) {
      const bool kv = lkb < kblocks;
      const uint32_t dst = wring + islot * RW_SLOT + lane * 16;
      const uint32_t dm = wring + islot * RW_SLOT + 1024 + lane * 4;
      const bool a0 = kv && v0, a1 = kv && v1;
      cp_async16(dst, a0 ? pc0 : packed, a0 ? 16u : 0u, pol);
      cp_async16(dst + 512, a1 ? pc1 : packed, a1 ? 16u : 0u, pol);
      cp_async4(dm, a0 ? pm0 : meta, a0 ? 4u : 0u);
      cp_async4(dm + 128, a1 ? pm0 + 32 : meta, a1 ? 4u : 0u);
      lkb += lstride;
      pc0 += lstride * 2 * RW_TILE_BYTES;
      pc1 += lstride * 2 * RW_TILE_BYTES;
      pm0 += lstride * mrow_pitch;
      if (++lj == lbpp) { lj = 0; ++lp; if (lp < np) setup_pass(lp); }
    }
    cp_async_commit();
    if (++islot == D) islot = 0;
  };

  setup_pass(0);
#pragma unroll 1
  for (int d = 0; d < D - 1; ++d) issue_next();   // static data: in flight before the previous kernel has finished

  // ===================== x =====================
  const int mrows = a.M;
  constexpr int kXPitch = IMMA ? 3 * RW_ICOL : RW_XPITCH;
  const int blk_pitch = mrows * kXPitch;          // bytes of x per staged k block
  const bool has_zp = a.has_zp != 0;
  const float zp_fix = -a.qmin;

  // x of blocks [kb0, kb0 + n): rows m < M.  Layouts as in K3a (q100_gemm_skinny.cu: stage_x).
  auto stage_x = [&](int kb0, int n) {
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int per_blk = mrows * 16;
    const int total_c = n * per_blk;                       // 16-byte chunks
    constexpr int XU = 4;                                  // loads of four rounds in flight (one L2 round trip, not four)
    for (int cb = tid; cb < total_c; cb += RW_THREADS * XU) {
      uint4 vq[XU], vu[XU];
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        const int c = cb + u * RW_THREADS;
        if (c < total_c) {
          const int bl = c / per_blk;
          const int rem = c - bl * per_blk;
          const int kofs = (kb0 + bl) * 128 + (rem & 15) * 8;
          if (a.x_gate_up) {
            const T* gp = xg + static_cast<int64_t>(rem >> 4) * (2 * a.K) + kofs;
            vq[u] = *reinterpret_cast<const uint4*>(gp);
            vu[u] = *reinterpret_cast<const uint4*>(gp + a.K);
          } else {
            vq[u] = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(rem >> 4) * a.K + kofs);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < XU; ++u) {
      const int c = cb + u * RW_THREADS;
      if (c >= total_c) break;                             // uniform per half-warp: total_c is a multiple of 16
      const int bl = c / per_blk;
      const int rem = c - bl * per_blk;
      const int m = rem >> 4, ch = rem & 15;
      uint4 v = a.x_gate_up ? silu_mul8<T>(vq[u], vu[u]) : vq[u];
      const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // the 16 chunks of one (block, m)
      if constexpr (IMMA) {
        // block fixed point: xi = rint(x * 2^(21 - e)), e = exponent of the (token, block) maximum, |xi| < 2^22,
        // split into balanced base-256 digits.  Bytes are stored in the order (e0, e4, e1, e5 | e2, e6, e3, e7) in
        // which the weight nibbles come out of a packed word.
        const float2 p0 = DT<T>::to_f2(v.x), p1 = DT<T>::to_f2(v.y), p2 = DT<T>::to_f2(v.z), p3 = DT<T>::to_f2(v.w);
        const float f[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
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
        uint8_t* colp = xs_ptr + bl * blk_pitch + (3 * m) * RW_ICOL;
        const int boff = (ch >> 2) * 32 + (ch & 3) * 8;
        *reinterpret_cast<uint2*>(colp + boff) = make_uint2(pk(l0[0], l0[4], l0[1], l0[5]), pk(l0[2], l0[6], l0[3], l0[7]));
        *reinterpret_cast<uint2*>(colp + RW_ICOL + boff) = make_uint2(pk(l1[0], l1[4], l1[1], l1[5]), pk(l1[2], l1[6], l1[3], l1[7]));
        *reinterpret_cast<uint2*>(colp + 2 * RW_ICOL + boff) = make_uint2(pk(l2[0], l2[4], l2[1], l2[5]), pk(l2[2], l2[6], l2[3], l2[7]));
        if (ch == 0) {
          // column coefficient 256^l / scale; inf / NaN anywhere in the block poisons the token's result
          const uint32_t cexp = static_cast<uint32_t>(254 - fe);
          const bool bad = e == 255;
          const float c0 = bad ? nan_f() : __uint_as_float(cexp << 23);
          const float c1 = bad ? nan_f() : __uint_as_float((cexp + 8) << 23);
          const float c2 = bad ? nan_f() : __uint_as_float((cexp + 16) << 23);
          *reinterpret_cast<float2*>(colp + 128) = make_float2(static_cast<float>(s0), c0);
          *reinterpret_cast<float2*>(colp + RW_ICOL + 128) = make_float2(static_cast<float>(s1), c1);
          *reinterpret_cast<float2*>(colp + 2 * RW_ICOL + 128) = make_float2(static_cast<float>(s2), c2);
        }
      } else {
        float2 f0 = DT<T>::to_f2(v.x), f1 = DT<T>::to_f2(v.y), f2 = DT<T>::to_f2(v.z), f3 = DT<T>::to_f2(v.w);
        float inv = 1.0f;
        if constexpr (kConvX) {
          // bf16 -> fp16 without loss: scale the (token, block) slice by the power of two that puts its largest
          // magnitude in [2^14, 2^15); the inverse power goes into the fp32 group step
          float am = fmaxf(fmaxf(fmaxf(fabsf(f0.x), fabsf(f0.y)), fmaxf(fabsf(f1.x), fabsf(f1.y))),
                           fmaxf(fmaxf(fabsf(f2.x), fabsf(f2.y)), fmaxf(fabsf(f3.x), fabsf(f3.y))));
#pragma unroll
          for (int d = 1; d < 16; d <<= 1) am = fmaxf(am, __shfl_xor_sync(half_mask, am, d));
          const int e = static_cast<int>((__float_as_uint(am) >> 23) & 0xffu);
          int fe = 268 - e;
          fe = (e == 0 || e == 255) ? 127 : min(max(fe, 1), 253);
          const float f = __uint_as_float(static_cast<uint32_t>(fe) << 23);
          inv = __uint_as_float(static_cast<uint32_t>(254 - fe) << 23);
          f0.x *= f; f0.y *= f; f1.x *= f; f1.y *= f; f2.x *= f; f2.y *= f; f3.x *= f; f3.y *= f;
          v.x = DT<__half>::pack2(f0.x, f0.y); v.y = DT<__half>::pack2(f1.x, f1.y);
          v.z = DT<__half>::pack2(f2.x, f2.y); v.w = DT<__half>::pack2(f3.x, f3.y);
          if (e != 0 && e != 255) {      // sums of what the tensor core will see (exact conversions)
            f0 = DT<__half>::to_f2(v.x); f1 = DT<__half>::to_f2(v.y); f2 = DT<__half>::to_f2(v.z); f3 = DT<__half>::to_f2(v.w);
          }
        }
        // sums of the "plain" pairs (value 1024 + q) and of the "x16" pairs (value 1024 + 16 q)
        float s_a = (f0.x + f0.y) + (f2.x + f2.y), s_b = (f1.x + f1.y) + (f3.x + f3.y);
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          s_a += __shfl_xor_sync(half_mask, s_a, d);
          s_b += __shfl_xor_sync(half_mask, s_b, d);
        }
        if (ch == 0)
          *reinterpret_cast<float4*>(xs_ptr + bl * blk_pitch + m * RW_XPITCH + 256 + ((m >> 1) & 3) * 16) =
              make_float4(s_a, s_b, inv, 0.f);
        { const uint32_t yy = v.y; v.y = v.z; v.z = yy; }   // one LDS.128 yields (k0,k1 | k4,k5 | k2,k3 | k6,k7)
        *reinterpret_cast<uint4*>(xs_ptr + bl * blk_pitch + m * RW_XPITCH + (ch >> 2) * 16 + (ch & 3) * 64) = v;
      }
      }
    }
  };

  // per-lane x addresses: mma column n = g is token g (integer form: (token, limb) = divmod(g, 3)); columns past the
  // last valid one re-read it (their results are never stored)
  uint32_t xrow, srow, srow1;
  if constexpr (IMMA) {
    const int ncols = 3 * mrows;
    xrow = xs_u32 + min(g, ncols - 1) * RW_ICOL + t * 32;
    srow = xs_u32 + min(2 * t, ncols - 1) * RW_ICOL + 128;
    srow1 = xs_u32 + min(2 * t + 1, ncols - 1) * RW_ICOL + 128;
  } else {
    xrow = xs_u32 + min(g, mrows - 1) * RW_XPITCH + t * 16;
    const int m0 = min(2 * t, mrows - 1), m1 = min(2 * t + 1, mrows - 1);
    srow = xs_u32 + m0 * RW_XPITCH + 256 + ((m0 >> 1) & 3) * 16;
    srow1 = xs_u32 + m1 * RW_XPITCH + 256 + ((m1 >> 1) & 3) * 16;
  }

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t magic = reg_const(0x64006400u);
  constexpr float kMagicF = 1024.0f;

  auto compute_block = [&](const BlockRegs& b, uint32_t xoff) {
    const uint32_t w0[4] = {b.c0.x, b.c0.y, b.c0.z, b.c0.w}, w1[4] = {b.c1.x, b.c1.y, b.c1.z, b.c1.w};
    const float2 sz0 = DT<T>::to_f2(b.m0), sz1 = DT<T>::to_f2(b.m1);     // (scale, zero) of rows g, g + 8
    if constexpr (IMMA) {
      int di[4] = {0, 0, 0, 0};
      const uint4 xb0 = lds128(xrow + xoff), xb1 = lds128(xrow + xoff + 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t af[4];
        af[0] = w0[j] & 0x0f0f0f0fu;                       // (e0, e4, e1, e5) of row g
        af[1] = w1[j] & 0x0f0f0f0fu;                       //                  of row g + 8
        af[2] = shr4_fma(w0[j]) & 0x0f0f0f0fu;             // (e2, e6, e3, e7)
        af[3] = shr4_fma(w1[j]) & 0x0f0f0f0fu;
        const uint4& q4 = (j >> 1) ? xb1 : xb0;
        imma_16832(di, af, (j & 1) ? q4.z : q4.x, (j & 1) ? q4.w : q4.y);
      }
      const float nz0 = has_zp ? -sz0.y : -zp_fix, nz1 = has_zp ? -sz1.y : -zp_fix;
      const float2 u0 = lds64f(srow + xoff), u1 = lds64f(srow1 + xoff);   // (sum of limbs, coefficient)
      const float r0 = fmaf(nz0, u0.x, static_cast<float>(di[0]));
      const float r1 = fmaf(nz0, u1.x, static_cast<float>(di[1]));
      const float r2 = fmaf(nz1, u0.x, static_cast<float>(di[2]));
      const float r3 = fmaf(nz1, u1.x, static_cast<float>(di[3]));
      acc[0] = fmaf(sz0.x * u0.y, r0, acc[0]);
      acc[1] = fmaf(sz0.x * u1.y, r1, acc[1]);
      acc[2] = fmaf(sz1.x * u0.y, r2, acc[2]);
      acc[3] = fmaf(sz1.x * u1.y, r3, acc[3]);
    } else {
      float ag[4], agh[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 xb = lds128(xrow + xoff + i * 64);
        const uint32_t s0 = shr8_fma(w0[i]), s1 = shr8_fma(w1[i]);
        uint32_t al[4], ah[4];
        al[0] = lop_and_or<0x000f000fu>(w0[i], magic);
        al[1] = lop_and_or<0x000f000fu>(w1[i], magic);
        al[2] = lop_and_or<0x000f000fu>(s0, magic);
        al[3] = lop_and_or<0x000f000fu>(s1, magic);
        ah[0] = lop_and_or<0x00f000f0u>(w0[i], magic);
        ah[1] = lop_and_or<0x00f000f0u>(w1[i], magic);
        ah[2] = lop_and_or<0x00f000f0u>(s0, magic);
        ah[3] = lop_and_or<0x00f000f0u>(s1, magic);
        const uint32_t bl2[2] = {xb.x, xb.y}, bh2[2] = {xb.z, xb.w};
        if (i == 0) { mma_16816_zero(ag, al, bl2, false); mma_16816_zero(agh, ah, bh2, false); }
        else { mma_16816(ag, al, bl2, false); mma_16816(agh, ah, bh2, false); }
      }
      const float z0 = has_zp ? sz0.y : zp_fix, z1 = has_zp ? sz1.y : zp_fix;
      const float off0 = -(kMagicF + z0), off1 = -(kMagicF + z1);
      const float offh0 = -(kMagicF * 0.0625f + z0), offh1 = -(kMagicF * 0.0625f + z1);
      // (S_plain, S_x16, 2^-e, -) of tokens m = 2t and 2t + 1
      const uint4 u0 = lds128(srow + xoff), u1 = lds128(srow1 + xoff);
      const float sa0 = __uint_as_float(u0.x), sb0 = __uint_as_float(u0.y);
      const float sa1 = __uint_as_float(u1.x), sb1 = __uint_as_float(u1.y);
      float r0 = fmaf(agh[0], 0.0625f, ag[0]), r1 = fmaf(agh[1], 0.0625f, ag[1]);
      float r2 = fmaf(agh[2], 0.0625f, ag[2]), r3 = fmaf(agh[3], 0.0625f, ag[3]);
      r0 = fmaf(offh0, sb0, r0); r1 = fmaf(offh0, sb1, r1);
      r2 = fmaf(offh1, sb0, r2); r3 = fmaf(offh1, sb1, r3);
      float w00 = sz0.x, w01 = sz0.x, w10 = sz1.x, w11 = sz1.x;
      if constexpr (kConvX) {
        const float i0 = __uint_as_float(u0.z), i1 = __uint_as_float(u1.z);
        w00 *= i0; w01 *= i1; w10 *= i0; w11 *= i1;
      }
      acc[0] = fmaf(w00, fmaf(off0, sa0, r0), acc[0]);
      acc[1] = fmaf(w01, fmaf(off0, sa1, r1), acc[1]);
      acc[2] = fmaf(w10, fmaf(off1, sa0, r2), acc[2]);
      acc[3] = fmaf(w11, fmaf(off1, sa1, r3), acc[3]);
    }
  };

  // end of pass p: the warps' fragments meet in shared memory, fixed summation order
  int pass_parity = 0;
  auto finish_pass = [&](int p) {
    float* rp = red + pass_parity * RED_FLOATS;
    pass_parity ^= 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      rp[(warp * 4 + r) * 32 + lane] = acc[r];
      acc[r] = 0.f;
    }
    __syncthreads();
    const int pr = min(RW_PR, nr - p * RW_PR);
    const bool two = pr > 16;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const int nw = two ? 4 : 8;                   // warps that share a slab
    for (int o = tid; o < RW_PR * mrows; o += RW_THREADS) {
      const int rr32 = o & 31, m = o >> 5;
      if (rr32 < pr) {
        const int sl = rr32 >> 4, rr = rr32 & 15;
        float sum = 0.f;
        for (int j = 0; j < nw; ++j) {
          const int w = two ? sl + 2 * j : j;
          if constexpr (IMMA) {
#pragma unroll
            for (int l = 0; l < 3; ++l) {
              const int c = 3 * m + l;
              sum += rp[(w * 4 + ((c & 1) | ((rr >> 3) << 1))) * 32 + (rr & 7) * 4 + ((c & 7) >> 1)];
            }
          } else {
            sum += rp[(w * 4 + ((m & 1) | ((rr >> 3) << 1))) * 32 + (rr & 7) * 4 + ((m & 7) >> 1)];
          }
        }
        const int n = R0 + p * RW_PR + rr32;
        if (bias) sum += DT<T>::to_f(bias[n]);
        y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(sum);
      }
    }
  };

  pdl_wait();                                     // x and y belong to the stream order

  const int chunk = a.chunk;
  const bool x_once = chunk >= kblocks;
  if (x_once) {
    stage_x(0, kblocks);
    __syncthreads();
  }
  // the compute side walks the same sequence as the load iterator, D - 1 blocks behind
  uint32_t cslot = 0;
#pragma unroll 1
  for (int p = 0; p < np; ++p) {
    const bool two = (p < np - 1) || last_two;
    int ckb = two ? (warp >> 1) : warp;
    const int cstride = two ? 4 : 8;
    const int cbpp = two ? bpp2 : bpp1;
    int ckc = 0;                                  // first block of the staged chunk
    if (!x_once) {
      __syncthreads();                            // everyone is done with the previous slab
      stage_x(0, min(chunk, kblocks));
      __syncthreads();
    }
#pragma unroll 1
    for (int j = 0; j < cbpp; ++j) {
      if (!x_once) {
        // the chunk that holds the j-th blocks of every warp: a chunk is a multiple of 8 blocks, and the warps' k
        // groups of one j lie inside one aligned run of 4 or 8
        const int lead = (j * cstride / chunk) * chunk;
        if (lead != ckc) {
          ckc = lead;
          __syncthreads();
          stage_x(ckc, min(chunk, kblocks - ckc));
          __syncthreads();
        }
      }
      cp_async_wait<D - 2>();                     // this lane's share of the oldest block has landed (it reads only that)
      BlockRegs cur;
      const uint32_t sb = wring + cslot * RW_SLOT + lane * 16;
      cur.c0 = lds128(sb);
      cur.c1 = lds128(sb + 512);
      cur.m0 = lds32(wring + cslot * RW_SLOT + 1024 + lane * 4);
      cur.m1 = lds32(wring + cslot * RW_SLOT + 1152 + lane * 4);
      if (++cslot == D) cslot = 0;
      issue_next();                               // into the slot that was read one step ago
      if (ckb < kblocks) compute_block(cur, static_cast<uint32_t>((ckb - ckc) * blk_pitch));
      ckb += cstride;
    }
    finish_pass(p);
  }
}

struct RowsPlan { int occ, depth, chunk, smem; };

// shared-memory plan: two CTAs per SM (one on request); all of x next to the warps' rings when that leaves a ring of
// >= 6 blocks per warp, else x in chunks
bool plan_rows(int M, int K, int mode, RowsPlan& p) {
  const int kblocks = K / 128;
  const int blk_pitch = M * (mode == 2 ? 3 * RW_ICOL : RW_XPITCH);
  const int fixed = 2 * RW_WARPS * 4 * 32 * 4 + 256;
  const int x_all = kblocks * blk_pitch;
  static const int env_occ = []() { const char* e = getenv("Q100_ROWS_OCC"); return e ? atoi(e) : 0; }();
  static const int env_depth = []() { const char* e = getenv("Q100_ROWS_DEPTH"); return e ? atoi(e) : 0; }();
  const int occ = env_occ == 1 ? 1 : 2;
  const int budget = occ == 1 ? 232448 : (233472 - 2 * 1024) / 2;
  const int depths[3] = {8, 6, 4};
  for (int di = 0; di < 3; ++di) {
    const int depth = depths[di];
    if (env_depth > 0 && depth != env_depth) continue;
    const int ring = RW_WARPS * depth * RW_SLOT;
    if (fixed + ring + x_all <= budget) {
      p = RowsPlan{occ, depth, ((kblocks + 7) / 8) * 8, fixed + ring + x_all};
      return true;
    }
    if (depth == 6 || env_depth > 0) {            // chunked x with a ring of 6 rather than everything with a ring of 4
      const int chunk = ((budget - fixed - ring) / blk_pitch) / 8 * 8;
      if (chunk >= 16) {
        p = RowsPlan{occ, depth, chunk, fixed + ring + chunk * blk_pitch};
        return true;
      }
    }
  }
  return false;
}

template <typename T, int MODE, int D, int OCC>
int launch_rows_t(RowsArgs a, const RowsPlan& p, int pdl, cudaStream_t st) {
  auto kern = k3_rows<T, MODE, D, OCC>;
  static std::mutex mu;
  static int max_set[kMaxDevices] = {};
  {
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (p.smem > max_set[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));
      max_set[dev] = p.smem;
    }
  }
  a.chunk = p.chunk;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sm_count() * p.occ);
  cfg.blockDim = dim3(RW_THREADS);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, a), "k3_rows launch");
}

template <typename T, int MODE>
int launch_rows(const RowsArgs& a, const RowsPlan& p, int pdl, cudaStream_t st) {
  if (p.occ == 1) return p.depth == 8 ? launch_rows_t<T, MODE, 8, 1>(a, p, pdl, st)
                       : p.depth == 6 ? launch_rows_t<T, MODE, 6, 1>(a, p, pdl, st) : launch_rows_t<T, MODE, 4, 1>(a, p, pdl, st);
  return p.depth == 8 ? launch_rows_t<T, MODE, 8, 2>(a, p, pdl, st)
       : p.depth == 6 ? launch_rows_t<T, MODE, 6, 2>(a, p, pdl, st) : launch_rows_t<T, MODE, 4, 2>(a, p, pdl, st);
}

}  // namespace

// mode: 1 = fp16 magic form (M <= 8), 2 = integer form (M <= 2)
bool gemm_rows_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits, int mode) {
  static const bool off = []() { const char* e = getenv("Q100_GEMM_ROWS"); return e && atoi(e) == 0; }();
  if (off) return false;
  if (n_bits != 4 || group != 128 || K % 128 != 0 || N % 8 != 0 || M < 1) return false;
  if (mode == 2 ? M > 2 : M > 8) return false;
  if (mode != 1 && mode != 2) return false;
  if (N > (1 << 24) || K > (1 << 20)) return false;
  RowsPlan p;
  return plan_rows(static_cast<int>(M), static_cast<int>(K), mode, p);
}

int gemm_rows(const void* x, const void* packed, const void* meta, const void* bias, void* y, int64_t M, int64_t N,
              int64_t K, int disable_zp, int dtype, int mode, int pdl, int x_gate_up, cudaStream_t st) {
  RowsPlan p;
  Q100_REQUIRE(plan_rows(static_cast<int>(M), static_cast<int>(K), mode, p), Q100_EUNSUPPORTED,
               "k3_rows: no shared-memory plan for M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  RowsArgs a{};
  a.x = x; a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K);
  a.kblocks = static_cast<int>(K / 128);
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -8.0f : 0.0f;
  a.x_gate_up = x_gate_up;
  if (dtype == Q100_F16) return mode == 2 ? launch_rows<__half, 2>(a, p, pdl, st) : launch_rows<__half, 1>(a, p, pdl, st);
  return mode == 2 ? launch_rows<__nv_bfloat16, 2>(a, p, pdl, st) : launch_rows<__nv_bfloat16, 1>(a, p, pdl, st);
}

}  // namespace q100
