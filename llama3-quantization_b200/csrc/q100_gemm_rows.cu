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

constexpr int RW_WARPS = 16;
constexpr int RW_THREADS = RW_WARPS * 32;
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
  int pair;               // two blocks per step (development switch)
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
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
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

// ---------------------------------------------------------------- the kernel
// Shared memory: red[2][16 warps][4][32] fp32 | misc (token maxima, coefficients) | the warps' rings [16][D][RW_SLOT] |
// stab[kblocks][8] fp32 (zero-point sums per limb column) | x limbs [kblocks][3 M][RW_ICOL]
constexpr int RW_RED_BYTES = 2 * RW_WARPS * 4 * 32 * 4;
constexpr int RW_MISC_BYTES = 1024;
constexpr int RW_XR = 8;                          // 16-byte chunks of x a thread holds between the two staging phases

template <typename T, int D, bool HAS_ZP>
__global__ void __launch_bounds__(RW_THREADS, 1) k3_rows(const RowsArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  float* red = reinterpret_cast<float*>(smem_raw);
  uint32_t* tokmax = reinterpret_cast<uint32_t*>(smem_raw + RW_RED_BYTES);               // [16 warps][2]
  float* coef = reinterpret_cast<float*>(smem_raw + RW_RED_BYTES + 256);                 // [2 tokens][4]
  const uint32_t wring = smem_u32(smem_raw) + RW_RED_BYTES + RW_MISC_BYTES + warp * (D * RW_SLOT) + lane * 16;
  uint8_t* stab_ptr = smem_raw + RW_RED_BYTES + RW_MISC_BYTES + RW_WARPS * D * RW_SLOT;
  const int kblocks = a.kblocks;
  uint8_t* xs_ptr = stab_ptr + kblocks * 32;
  const int mrows = a.M;
  const int blk_pitch = 3 * mrows * RW_ICOL;      // bytes of x limbs per k block

  const int G = gridDim.x, cta = blockIdx.x;
  const int Q = a.N >> 2;
  const int q0 = static_cast<int>(static_cast<int64_t>(cta) * Q / G);
  const int q1 = static_cast<int>(static_cast<int64_t>(cta + 1) * Q / G);
  const int R0 = 4 * q0, nr = 4 * (q1 - q0);
  pdl_launch_dependents();
  if (nr == 0) return;

  // ---- geometry of this CTA: its ns = ceil(nr / 16) slabs of 16 rows are walked in passes of S = 4 slabs, then what
  // is left as 2 + 1, 2 or 1; in a pass of S slabs the 16 warps split the work by (slab, k group mod KW), KW = 16 / S
  const int ns = (nr + 15) >> 4;
  const int n4 = ns >> 2, rem = ns & 3;
  const int np = n4 + (rem == 3 ? 2 : rem ? 1 : 0);
  auto pass_geo = [&](int p, int& sh_, int& slab0_) {   // log2 S and first slab of pass p
    if (p < n4) { sh_ = 2; slab0_ = 4 * p; }
    else if (rem == 3) { sh_ = p == n4 ? 1 : 0; slab0_ = 4 * n4 + (p == n4 ? 0 : 2); }
    else { sh_ = rem == 2 ? 1 : 0; slab0_ = 4 * n4; }
  };
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(a.packed);
  const uint8_t* meta = reinterpret_cast<const uint8_t*>(a.meta);
  const int64_t ktile_row = static_cast<int64_t>(kblocks) * 2 * RW_TILE_BYTES;   // bytes of one 128-row tile row

  // ---- the load iterator: rows past the CTA's range are clamped to its last row (same sectors, results unused)
  const uint8_t* pc0; const uint8_t* pc1; const uint8_t* pm0; const uint8_t* pm1;
  int64_t cstep = 0, mstep = 0;                   // to this warp's next k group
  int lj = 0, lp = 0, lbpp = 0;
  auto setup_pass = [&](int p) {
    int sh_, slab0_;
    pass_geo(p, sh_, slab0_);
    const int kw = RW_WARPS >> sh_, kq = warp >> sh_;
    const int rel = (slab0_ + (warp & ((1 << sh_) - 1))) * 16 + g;
    const int r0 = R0 + min(rel, nr - 1), r1 = R0 + min(rel + 8, nr - 1);
    const int lo = (t >> 1) * RW_TILE_BYTES + (t & 1) * 16 + kq * 2 * RW_TILE_BYTES;
    pc0 = packed + (r0 >> 7) * ktile_row + (r0 & 127) * RW_TROW + lo;
    pc1 = packed + (r1 >> 7) * ktile_row + (r1 & 127) * RW_TROW + lo;
    pm0 = meta + (static_cast<int64_t>(kq) * a.N + r0) * 4;
    pm1 = meta + (static_cast<int64_t>(kq) * a.N + r1) * 4;
    cstep = static_cast<int64_t>(kw) * 2 * RW_TILE_BYTES;
    mstep = static_cast<int64_t>(kw) * a.N * 4;
    lbpp = kblocks >> (4 - sh_);
  };
  auto issue = [&](uint32_t slot_addr) {          // one cp.async group: the block under the iterator (or nothing)
    if (lp < np) {
      cp_async16(slot_addr, pc0);
      cp_async16(slot_addr + 512, pc1);
      cp_async4(slot_addr + 1024 - lane * 12, pm0);
      cp_async4(slot_addr + 1152 - lane * 12, pm1);
      pc0 += cstep; pc1 += cstep; pm0 += mstep; pm1 += mstep;
      if (++lj == lbpp) { lj = 0; ++lp; if (lp < np) setup_pass(lp); }
    }
    cp_async_commit();
  };
  setup_pass(0);
#pragma unroll 1
  for (int d = 0; d < D - 2; ++d) issue(wring + d * RW_SLOT);   // static data: in flight before the previous kernel ends

  // ===================== x: block fixed point with ONE exponent per token =====================
  // xi = rint(x 2^(21 - e)), e = exponent of the token's largest magnitude, as three balanced base-256 digits, one mma
  // column each.  The elements that meet the HIGH nibbles of the packed words (e2, e3, e6, e7 of every 8) are scaled
  // by another 1/16: the weight operand for them is (w & 0xf0f0f0f0) = 16 q, no shift.  stab holds, per k block and
  // column, sum_lo l + 16 sum_hi l: the factor of the weight zero-point.
  pdl_wait();                                     // x and y belong to the stream order
  {
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int cpt = a.K >> 3;                     // 16-byte chunks per token
    const int total_c = mrows * cpt;
    uint4 vq[RW_XR], vu[RW_XR];
#pragma unroll
    for (int u = 0; u < RW_XR; ++u) {
      const int c = tid + u * RW_THREADS;
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
    for (int i = tid; i < kblocks * 8; i += RW_THREADS) reinterpret_cast<float*>(stab_ptr)[i] = 0.f;
    uint32_t mx0 = 0, mx1 = 0;                    // max |x| as bits per token: finite < inf < NaN
#pragma unroll
    for (int u = 0; u < RW_XR; ++u) {
      const int c = tid + u * RW_THREADS;
      if (c < total_c) {
        if (a.x_gate_up) vq[u] = silu_mul8<T>(vq[u], vu[u]);
        const float2 p0 = DT<T>::to_f2(vq[u].x), p1 = DT<T>::to_f2(vq[u].y), p2 = DT<T>::to_f2(vq[u].z), p3 = DT<T>::to_f2(vq[u].w);
        uint32_t ab = max(max(__float_as_uint(p0.x) & 0x7fffffffu, __float_as_uint(p0.y) & 0x7fffffffu),
                          max(__float_as_uint(p1.x) & 0x7fffffffu, __float_as_uint(p1.y) & 0x7fffffffu));
        ab = max(ab, max(max(__float_as_uint(p2.x) & 0x7fffffffu, __float_as_uint(p2.y) & 0x7fffffffu),
                         max(__float_as_uint(p3.x) & 0x7fffffffu, __float_as_uint(p3.y) & 0x7fffffffu)));
        if (c >= cpt) mx1 = max(mx1, ab); else mx0 = max(mx0, ab);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, d));
      mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, d));
    }
    if (lane == 0) { tokmax[warp * 2] = mx0; tokmax[warp * 2 + 1] = mx1; }
    __syncthreads();
    mx0 = 0; mx1 = 0;
#pragma unroll
    for (int w = 0; w < RW_WARPS; ++w) { mx0 = max(mx0, tokmax[w * 2]); mx1 = max(mx1, tokmax[w * 2 + 1]); }
    const int e0 = static_cast<int>(mx0 >> 23), e1 = static_cast<int>(mx1 >> 23);
    const int fe0 = min(max(275 - e0, 5), 250), fe1 = min(max(275 - e1, 5), 250);   // scale = 2^(fe - 127)
    if (tid < 2) {
      const int fe = tid ? fe1 : fe0;
      const bool bad = (tid ? e1 : e0) == 255;    // inf / NaN anywhere poisons the token's result
      const uint32_t cexp = static_cast<uint32_t>(254 - fe);
      coef[tid * 4 + 0] = bad ? nan_f() : __uint_as_float(cexp << 23);
      coef[tid * 4 + 1] = bad ? nan_f() : __uint_as_float((cexp + 8) << 23);
      coef[tid * 4 + 2] = bad ? nan_f() : __uint_as_float((cexp + 16) << 23);
    }
    const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // the 16 chunks of one (token, block)
#pragma unroll
    for (int u = 0; u < RW_XR; ++u) {
      const int c = tid + u * RW_THREADS;
      if (c < total_c) {                          // uniform per half-warp: total_c is a multiple of 16
        const int m = c >= cpt ? 1 : 0;
        const int kc = c - m * cpt;               // chunk inside the token
        const int bl = kc >> 4, ch = kc & 15;
        const int fe = m ? fe1 : fe0;
        const float sc_lo = __uint_as_float(static_cast<uint32_t>(fe) << 23);
        const float sc_hi = __uint_as_float(static_cast<uint32_t>(fe - 4) << 23);
        const float2 p0 = DT<T>::to_f2(vq[u].x), p1 = DT<T>::to_f2(vq[u].y), p2 = DT<T>::to_f2(vq[u].z), p3 = DT<T>::to_f2(vq[u].w);
        const float f[8] = {p0.x * sc_lo, p0.y * sc_lo, p1.x * sc_hi, p1.y * sc_hi,
                            p2.x * sc_lo, p2.y * sc_lo, p3.x * sc_hi, p3.y * sc_hi};
        int l0[8], l1[8], l2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int xi = __float2int_rn(f[i]);
          l0[i] = (xi << 24) >> 24;
          const int x1 = (xi - l0[i]) >> 8;
          l1[i] = (x1 << 24) >> 24;
          l2[i] = (x1 - l1[i]) >> 8;
        }
        int s0 = (l0[0] + l0[1] + l0[4] + l0[5]) + 16 * (l0[2] + l0[3] + l0[6] + l0[7]);
        int s1 = (l1[0] + l1[1] + l1[4] + l1[5]) + 16 * (l1[2] + l1[3] + l1[6] + l1[7]);
        int s2 = (l2[0] + l2[1] + l2[4] + l2[5]) + 16 * (l2[2] + l2[3] + l2[6] + l2[7]);
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
        // bytes in the order in which the nibbles come out of a packed word: (e0, e4, e1, e5 | e2, e6, e3, e7)
        uint8_t* colp = xs_ptr + bl * blk_pitch + (3 * m) * RW_ICOL + (ch >> 2) * 32 + (ch & 3) * 8;
        *reinterpret_cast<uint2*>(colp) = make_uint2(pk(l0[0], l0[4], l0[1], l0[5]), pk(l0[2], l0[6], l0[3], l0[7]));
        *reinterpret_cast<uint2*>(colp + RW_ICOL) = make_uint2(pk(l1[0], l1[4], l1[1], l1[5]), pk(l1[2], l1[6], l1[3], l1[7]));
        *reinterpret_cast<uint2*>(colp + 2 * RW_ICOL) = make_uint2(pk(l2[0], l2[4], l2[1], l2[5]), pk(l2[2], l2[6], l2[3], l2[7]));
        if (ch == 0) {
          float* sp = reinterpret_cast<float*>(stab_ptr) + bl * 8 + 3 * m;
          sp[0] = static_cast<float>(s0); sp[1] = static_cast<float>(s1); sp[2] = static_cast<float>(s2);
        }
      }
    }
    __syncthreads();
  }

  // ===================== the stream =====================
  const int ncols = 3 * mrows;
  const float nz_fix = a.qmin;
  // ring bookkeeping: D - 2 blocks are outstanding between steps, the two slots behind the read position are free;
  // a step reads NB slots and refills NB slots that were read at least one step ago
  const uint8_t* wring_g = smem_raw + (wring - smem_u32(smem_raw));       // generic address of this lane's ring
  uint32_t coff = 0, roff = (D - 2) * RW_SLOT;    // byte offsets of the read position / the refill position
  auto next_off = [&](uint32_t o) { o += RW_SLOT; return o == D * RW_SLOT ? 0u : o; };
  auto step = [&](auto nb_tag, uint32_t xo, uint32_t so, uint32_t xstep, uint32_t sstep, float (&acc)[2][4]) {
    constexpr int NB = decltype(nb_tag)::value;
    cp_async_wait<D - 2 - NB>();                  // this lane's share of the NB oldest blocks has landed (it reads only that)
    uint4 c0[NB], c1[NB], xb0[NB], xb1[NB];
    uint32_t m0[NB], m1[NB];
    float2 sv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint8_t* sp = wring_g + coff;
      c0[b] = *reinterpret_cast<const uint4*>(sp);
      c1[b] = *reinterpret_cast<const uint4*>(sp + 512);
      m0[b] = *reinterpret_cast<const uint32_t*>(sp + 1024 - lane * 12);
      m1[b] = *reinterpret_cast<const uint32_t*>(sp + 1152 - lane * 12);
      coff = next_off(coff);
      xb0[b] = *reinterpret_cast<const uint4*>(smem_raw + xo + b * xstep);
      xb1[b] = *reinterpret_cast<const uint4*>(smem_raw + xo + b * xstep + 16);
      sv[b] = *reinterpret_cast<const float2*>(smem_raw + so + b * sstep);      // zero-point sums of columns 2t, 2t + 1
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) { issue(wring + roff); roff = next_off(roff); }
    int di[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b) { di[b][0] = di[b][1] = di[b][2] = di[b][3] = 0; }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t w0 = w == 0 ? c0[b].x : w == 1 ? c0[b].y : w == 2 ? c0[b].z : c0[b].w;
        const uint32_t w1 = w == 0 ? c1[b].x : w == 1 ? c1[b].y : w == 2 ? c1[b].z : c1[b].w;
        const uint4& xq = (w >> 1) ? xb1[b] : xb0[b];
        uint32_t af[4];
        af[0] = w0 & 0x0f0f0f0fu; af[1] = w1 & 0x0f0f0f0fu; af[2] = w0 & 0xf0f0f0f0u; af[3] = w1 & 0xf0f0f0f0u;
        imma_16832(di[b], af, (w & 1) ? xq.z : xq.x, (w & 1) ? xq.w : xq.y);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float2 sz0 = DT<T>::to_f2(m0[b]), sz1 = DT<T>::to_f2(m1[b]);     // (scale, zero) of rows g, g + 8
      const float nz0 = HAS_ZP ? -sz0.y : nz_fix, nz1 = HAS_ZP ? -sz1.y : nz_fix;
      acc[b][0] = fmaf(sz0.x, fmaf(nz0, sv[b].x, static_cast<float>(di[b][0])), acc[b][0]);
      acc[b][1] = fmaf(sz0.x, fmaf(nz0, sv[b].y, static_cast<float>(di[b][1])), acc[b][1]);
      acc[b][2] = fmaf(sz1.x, fmaf(nz1, sv[b].x, static_cast<float>(di[b][2])), acc[b][2]);
      acc[b][3] = fmaf(sz1.x, fmaf(nz1, sv[b].y, static_cast<float>(di[b][3])), acc[b][3]);
    }
  };
  int parity = 0;
#pragma unroll 1
  for (int p = 0; p < np; ++p) {
    int sh, slab0;
    pass_geo(p, sh, slab0);
    const int S = 1 << sh, KW = RW_WARPS >> sh, kq = warp >> sh;
    const int bpp = kblocks >> (4 - sh);          // blocks per warp in this pass
    // mma column g is (token, limb) = divmod(g, 3); columns past the last one re-read it (never stored)
    uint32_t xo = static_cast<uint32_t>(xs_ptr - smem_raw) + min(g, ncols - 1) * RW_ICOL + t * 32 + kq * blk_pitch;
    uint32_t so = static_cast<uint32_t>(stab_ptr - smem_raw) + t * 8 + kq * 32;
    const uint32_t xstep = KW * blk_pitch, sstep = KW * 32;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    int j = 0;
#pragma unroll 1
    for (; a.pair && j + 2 <= bpp; j += 2) {      // two blocks per step: independent chains the scheduler interleaves
      step(std::integral_constant<int, 2>{}, xo, so, xstep, sstep, acc);
      xo += 2 * xstep; so += 2 * sstep;
    }
#pragma unroll 1
    for (; j < bpp; ++j) {
      step(std::integral_constant<int, 1>{}, xo, so, xstep, sstep, acc);
      xo += xstep; so += sstep;
    }
    // ---- end of the pass: the warps' fragments meet in shared memory, fixed summation order
    float* rp = red + parity * (RW_WARPS * 4 * 32);
    parity ^= 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) rp[(warp * 4 + r) * 32 + lane] = acc[0][r] + acc[1][r];
    __syncthreads();
    const int PRW = 16 * S;
    if (tid < PRW * mrows) {
      const int rr = tid & (PRW - 1), m = tid >> (4 + sh);
      const int rel = slab0 * 16 + rr;
      if (rel < nr) {
        const int sl = rr >> 4, r16 = rr & 15;
        float sum = 0.f;
        for (int kw = 0; kw < KW; ++kw) {
          const float* wp = rp + ((sl + S * kw) * 4 + ((r16 >> 3) << 1)) * 32 + (r16 & 7) * 4;
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            const int c = 3 * m + l;
            sum = fmaf(coef[m * 4 + l], wp[(c & 1) * 32 + ((c & 7) >> 1)], sum);
          }
        }
        const int n = R0 + rel;
        if (a.bias) sum += DT<T>::to_f(reinterpret_cast<const T*>(a.bias)[n]);
        reinterpret_cast<T*>(a.y)[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(sum);
      }
    }
  }
}

struct RowsPlan { int depth, smem; };

bool plan_rows(int M, int K, RowsPlan& p) {
  const int kblocks = K / 128;
  const int fixed = RW_RED_BYTES + RW_MISC_BYTES + kblocks * 32 + kblocks * 3 * M * RW_ICOL + 128;
  static const int env_depth = []() { const char* e = getenv("Q100_ROWS_DEPTH"); return e ? atoi(e) : 0; }();
  const int depths[4] = {8, 6, 5, 4};              // >= 4: a step of two blocks waits for all but D - 4 groups
  for (int di = 0; di < 4; ++di) {
    const int depth = depths[di];
    if (env_depth > 0 && depth > env_depth) continue;
    const int smem = fixed + RW_WARPS * depth * RW_SLOT;
    if (smem <= 232448 - 1024) { p = RowsPlan{depth, smem}; return true; }
  }
  return false;
}

template <typename T, int D>
int launch_rows_t(const RowsArgs& a, const RowsPlan& p, int pdl, cudaStream_t st) {
  auto kern = a.has_zp ? k3_rows<T, D, true> : k3_rows<T, D, false>;
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
  cfg.gridDim = dim3(sm_count());
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

template <typename T>
int launch_rows(const RowsArgs& a, const RowsPlan& p, int pdl, cudaStream_t st) {
  switch (p.depth) {
    case 8: return launch_rows_t<T, 8>(a, p, pdl, st);
    case 6: return launch_rows_t<T, 6>(a, p, pdl, st);
    case 5: return launch_rows_t<T, 5>(a, p, pdl, st);
    default: return launch_rows_t<T, 4>(a, p, pdl, st);
  }
}

}  // namespace

// mode: 2 = integer form (M <= 2); the fp16 form (mode 1, M <= 8) stays on K3a
bool gemm_rows_supported(int64_t M, int64_t N, int64_t K, int group, int n_bits, int mode) {
  static const bool off = []() { const char* e = getenv("Q100_GEMM_ROWS"); return e && atoi(e) == 0; }();
  if (off) return false;
  if (mode != 2 || n_bits != 4 || group != 128 || K % 2048 != 0 || N % 8 != 0 || M < 1 || M > 2) return false;
  if (M * K > RW_XR * RW_THREADS * 8) return false;     // x is held in registers between the two staging phases
  if (N > (1 << 24) || K > (1 << 20)) return false;
  RowsPlan p;
  return plan_rows(static_cast<int>(M), static_cast<int>(K), p);
}

int gemm_rows(const void* x, const void* packed, const void* meta, const void* bias, void* y, int64_t M, int64_t N,
              int64_t K, int disable_zp, int dtype, int mode, int pdl, int x_gate_up, cudaStream_t st) {
  (void)mode;
  RowsPlan p;
  Q100_REQUIRE(plan_rows(static_cast<int>(M), static_cast<int>(K), p), Q100_EUNSUPPORTED,
               "k3_rows: no shared-memory plan for M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  RowsArgs a{};
  a.x = x; a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K);
  a.kblocks = static_cast<int>(K / 128);
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -8.0f : 0.0f;
  a.x_gate_up = x_gate_up;
  static const int env_pair = []() { const char* e = getenv("Q100_ROWS_PAIR"); return e ? atoi(e) : 1; }();
  a.pair = env_pair;
  if (dtype == Q100_F16) return launch_rows<__half>(a, p, pdl, st);
  return launch_rows<__nv_bfloat16>(a, p, pdl, st);
}

}  // namespace q100
