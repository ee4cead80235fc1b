// q100_quant.cu -- K1 (group-wise) and K2 (row-wise) fake-quantize / pack kernels.
//
// One pass over the tensor replaces the reference's per_token_dynamic_calibration +
// fake_quant (quant/quantizer.py:94-159): per-group min/max -> scale / zero-point ->
// round / clamp -> dequant, emitting any of {dequantised tensor, packed codes, scales,
// zeros, transposed (scale, zero) side band}.  HBM-bound: every input byte is read once
// with 128-bit streaming loads, every output byte written once.
//
// Bit-exactness contract: the reference executes each elementwise op in
// the storage dtype; torch computes in fp32 and rounds to the storage dtype after every
// op.  Two equivalent implementations live here:
//   MATH_SCALAR : literal restatement, fp32 op + round-trip through T after each op.
//   MATH_PACKED : (fp16/bf16 only) native packed half2/bfloat162 add/sub/mul, which are
//                 correctly rounded and therefore equal to "fp32 op then round" (double
//                 rounding is innocuous because 24 >= 2*11 + 2); division uses
//                 q = x*rcp(s); q += fma(-q, s, x)*rcp(s), which rounds to the same
//                 16-bit value as div.rn.f32 (checked exhaustively by q100_selftest_div).
// Compile this file with -fmad=false: no implicit contraction may merge the per-op steps.
#include <type_traits>

#include "q100_common.cuh"
#include "q100_quant.cuh"

namespace q100 {

// ------------------------------------------------------------------ one 16-byte vector
// Returns the dequantised vector; pairs[j] = code[2j] | code[2j+1] << 16 (unsigned codes,
// i.e. x_int - qmin) for the packers.
template <typename T, int MATH>
struct VecQuant;

// literal fp32 restatement, any T
template <typename T, int MATH>
__device__ __forceinline__ float quant_scalar(float x, const GroupQ<T>& q, const QuantArgs& a,
                                              uint32_t& code) {
  using D = DT<T>;
  const float v = D::rd(div_x_s<(DT<T>::kCode == Q100_F32) ? MATH_SCALAR_IEEE : MATH>(x, q.scale, q.rcp));
  float xi = D::rd(D::rd(rintf(v) - v) + v);            // round_ste forward value, quantizer.py:19
  if (a.has_zp) xi = D::rd(xi + q.zp);
  const float c = clamp_nan(xi, a.qmin, a.qmax);
  code = static_cast<uint32_t>(__float2int_rn(c - a.qmin));   // NaN -> 0
  const float d = a.has_zp ? D::rd(c - q.zp) : c;
  return D::rd(d * q.scale);
}

template <int MATH>
struct VecQuant<float, MATH> {
  static constexpr int E = 4;
  template <bool WANT_DEQ, bool WANT_CODES, bool FAST = false>
  __device__ static __forceinline__ uint4 run(const uint4& in, const GroupQ<float>& q,
                                              const QuantArgs& a, uint32_t (&pairs)[2]) {
    uint32_t c0, c1, c2, c3;
    uint4 o;
    o.x = __float_as_uint(quant_scalar<float, MATH>(__uint_as_float(in.x), q, a, c0));
    o.y = __float_as_uint(quant_scalar<float, MATH>(__uint_as_float(in.y), q, a, c1));
    o.z = __float_as_uint(quant_scalar<float, MATH>(__uint_as_float(in.z), q, a, c2));
    o.w = __float_as_uint(quant_scalar<float, MATH>(__uint_as_float(in.w), q, a, c3));
    pairs[0] = c0 | (c1 << 16);
    pairs[1] = c2 | (c3 << 16);
    return o;
  }
};

template <typename T>
struct Packed16;
template <>
struct Packed16<__half> {
  using T2 = __half2;
  __device__ static __forceinline__ T2 bcast(float v) { return __float2half2_rn(v); }
  __device__ static __forceinline__ T2 rint2(T2 v) { return h2rint(v); }
  // (v + C) - C == rint(v) (ties to even, -0 -> +0) for |v| <= kFastLimit: C = 1.5 * 2^10
  static constexpr float kRintMagic = 1536.0f, kFastLimit = 500.0f;
};
template <>
struct Packed16<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  __device__ static __forceinline__ T2 bcast(float v) { return __float2bfloat162_rn(v); }
  __device__ static __forceinline__ T2 rint2(T2 v) { return h2rint(v); }
  static constexpr float kRintMagic = 192.0f, kFastLimit = 60.0f;       // C = 1.5 * 2^7
};

template <typename T> struct FastLimit { static constexpr float value = 0.0f; };
template <> struct FastLimit<__half> { static constexpr float value = 500.0f; };
template <> struct FastLimit<__nv_bfloat16> { static constexpr float value = 60.0f; };
template <typename T> struct CodeMagic;
template <> struct CodeMagic<__half> { static constexpr float kMagic = 1024.0f; static constexpr uint32_t kMask = 0x03ff03ffu; };
template <> struct CodeMagic<__nv_bfloat16> { static constexpr float kMagic = 128.0f; static constexpr uint32_t kMask = 0x007f007fu; };

// One pair of elements.  Returns the dequantised pair (only computed when WANT_DEQ) and, when
// WANT_CODES, pair = code[0] | code[1] << 16 with code = x_int - qmin (NaN -> 0).
// FAST (warp-uniform, decided per step): every |x / scale| of the step is small enough for the magic-add
// rint below, which yields exactly round_ste's forward value for finite inputs of that size.
template <typename T, int MATH, bool WANT_DEQ, bool WANT_CODES, bool FAST = false>
__device__ __forceinline__ uint32_t quant_pair(uint32_t in, const GroupQ<T>& q, const QuantArgs& a,
                                               uint32_t& pair) {
  using D = DT<T>;
  if (MATH != MATH_PACKED) {
    const float2 xf = D::to_f2(in);
    uint32_t c0, c1;
    const float d0 = quant_scalar<T, MATH>(xf.x, q, a, c0);
    const float d1 = quant_scalar<T, MATH>(xf.y, q, a, c1);
    pair = c0 | (c1 << 16);
    return D::pack2(d0, d1);
  } else {
    using P = Packed16<T>;
    using T2 = typename P::T2;
    const float2 xf = D::to_f2(in);
    const float q0 = div_x_s<MATH_PACKED>(xf.x, q.scale, q.rcp);
    const float q1 = div_x_s<MATH_PACKED>(xf.y, q.scale, q.rcp);
    uint32_t vb = D::pack2(q0, q1);                         // rd(x / scale), both lanes
    const T2 v = *reinterpret_cast<T2*>(&vb);
    // round_ste: rint(v) + (v - v): +0 for finite v (so -0 becomes +0), NaN for inf/NaN
    T2 xi;
    if (FAST) xi = __hsub2(__hadd2(v, P::bcast(P::kRintMagic)), P::bcast(P::kRintMagic));
    else xi = __hadd2(P::rint2(v), __hsub2(v, v));
    const T2 zp2 = P::bcast(q.zp);
    if (a.has_zp) xi = __hadd2(xi, zp2);
    const T2 c = __hmin2_nan(__hmax2_nan(xi, P::bcast(a.qmin)), P::bcast(a.qmax));
    if (WANT_CODES) {
      if (a.n_bits < 8 || D::kCode == Q100_F16) {
        // integer-valued c in [qmin, qmax]: c - qmin + MAGIC is exact and carries the code in its low
        // mantissa bits; the non-propagating max turns NaN into qmin, i.e. code 0
        const T2 cc = __hmax2(c, P::bcast(a.qmin));
        const T2 m = __hadd2(cc, P::bcast(CodeMagic<T>::kMagic - a.qmin));
        pair = *reinterpret_cast<const uint32_t*>(&m) & CodeMagic<T>::kMask;
      } else {
        // bf16, 8-bit codes: 128 + code is not representable; go through fp32 (exact: |c| <= 2^15)
        uint32_t cb = *reinterpret_cast<const uint32_t*>(&c);
        const float2 cf = D::to_f2(cb);
        pair = static_cast<uint32_t>(__float2int_rn(cf.x - a.qmin)) |
               (static_cast<uint32_t>(__float2int_rn(cf.y - a.qmin)) << 16);
      }
    }
    if (!WANT_DEQ) return 0u;
    T2 d = a.has_zp ? __hsub2(c, zp2) : c;
    d = __hmul2(d, P::bcast(q.scale));
    return *reinterpret_cast<uint32_t*>(&d);
  }
}

template <typename T, int MATH>
struct VecQuant {
  static constexpr int E = 8;
  template <bool WANT_DEQ, bool WANT_CODES, bool FAST = false>
  __device__ static __forceinline__ uint4 run(const uint4& in, const GroupQ<T>& q,
                                              const QuantArgs& a, uint32_t (&pairs)[4]) {
    uint4 o;
    o.x = quant_pair<T, MATH, WANT_DEQ, WANT_CODES, FAST>(in.x, q, a, pairs[0]);
    o.y = quant_pair<T, MATH, WANT_DEQ, WANT_CODES, FAST>(in.y, q, a, pairs[1]);
    o.z = quant_pair<T, MATH, WANT_DEQ, WANT_CODES, FAST>(in.z, q, a, pairs[2]);
    o.w = quant_pair<T, MATH, WANT_DEQ, WANT_CODES, FAST>(in.w, q, a, pairs[3]);
    return o;
  }
};

// ------------------------------------------------------------------ packers (q100 v1 layout)
// A uint32 word of a b-bit plane covers F = 32/b consecutive K elements; element e of the
// word sits in field (e >> 1) + (e & 1) * F/2, i.e. bit b*(e>>1) + 16*(e&1).  With
// pairs[j] = code[2j] | code[2j+1] << 16 a vector starting at (even) word-relative element
// o contributes  sum_j pairs[j] << (b * (o/2 + j)).
template <int NP>
__device__ __forceinline__ uint32_t plane_word(const uint32_t (&pairs)[NP], int b, int shift0,
                                               uint32_t pair_mask, int pre_shift) {
  uint32_t w = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) w |= ((pairs[j] >> pre_shift) & pair_mask) << (shift0 + b * j);
  return w;
}

// Emits the packed words of one vector.  `vec_in_row` = index of this 16-byte vector inside
// its row, `lane_sub` = position of the thread among consecutive-vector lanes (all lanes
// of the warp call this together: shuffles inside).
template <int E>
__device__ __forceinline__ void emit_packed(const uint32_t (&pairs)[E / 2], const QuantArgs& a,
                                            int row, int vec_in_row, bool valid) {
  const int e0 = vec_in_row * E;                     // first element (column) of the vector
  // base of the [128 x 64] tile holding (row, e0), then the row inside the tile
  uint32_t* tile = a.packed + (static_cast<int64_t>(row >> 7) * a.k_tiles + (e0 >> 6)) * a.tile_words;
  const int r = row & 127;
  const int ek = e0 & 63;                            // column inside the tile
  const int nb = a.n_bits;
  if (nb == 8) {                                     // 16 words per row, 4 codes per word
    uint32_t* prow = tile + r * 16;
    if (E == 8) {
      if (valid) {
        uint2 w;
        w.x = (pairs[0] & 0x00ff00ffu) | ((pairs[1] & 0x00ff00ffu) << 8);
        w.y = (pairs[E / 2 - 2] & 0x00ff00ffu) | ((pairs[E / 2 - 1] & 0x00ff00ffu) << 8);
        *reinterpret_cast<uint2*>(prow + (ek >> 2)) = w;
      }
    } else {
      if (valid) prow[ek >> 2] = (pairs[0] & 0x00ff00ffu) | ((pairs[1] & 0x00ff00ffu) << 8);
    }
    return;
  }
  // planes: b bits per field, F = 32 / b = 1 << lf elements per word (all powers of two: shifts only)
  const int n_planes = (nb == 3) ? 2 : 1;
#pragma unroll
  for (int pl = 0; pl < 2; ++pl) {
    if (pl >= n_planes) break;
    const int b = (nb == 3) ? (pl == 0 ? 2 : 1) : nb;
    const int pre = (nb == 3 && pl == 1) ? 2 : 0;
    const int lf = (b == 4) ? 3 : (b == 2) ? 4 : 5;          // log2(F)
    const uint32_t m1 = (1u << b) - 1u;
    const uint32_t pmask = m1 | (m1 << 16);
    const int lpw = (1 << lf) / E;                           // lanes per word (>= 1 here)
    const int o = ek & ((1 << lf) - 1);                      // word-relative first element (multiple of E)
    uint32_t w = plane_word<E / 2>(pairs, b, b * (o >> 1), pmask, pre);
    if (!valid) w = 0;
    for (int d = 1; d < lpw; d <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, d);
    // words per row inside the tile for this plane: 64 * b / 32 = 2 * b
    if (valid && o == 0) tile[(pl == 1 ? 128 * 4 : 0) + r * (2 * b) + (ek >> lf)] = w;
  }
}

// W4, 16-bit input: a vector of 8 elements is exactly one packed word (4 pairs, 4 bits apart); the
// shifts-and-adds run as IMADs on the FMA pipe.  Same word, same address as emit_packed<8>.
__device__ __forceinline__ void emit_packed_w4(const uint32_t (&pairs)[4], const QuantArgs& a,
                                               int row, int vec_in_row) {
  const int e0 = vec_in_row * 8;
  uint32_t* tile = a.packed + (static_cast<int64_t>(row >> 7) * a.k_tiles + (e0 >> 6)) * (128 * 8);
  const uint32_t w = pairs[0] + pairs[1] * 16u + pairs[2] * 256u + pairs[3] * 4096u;
  tile[(row & 127) * 8 + ((e0 & 63) >> 3)] = w;
}

template <typename T>
__device__ __forceinline__ void store_group_meta(const QuantArgs& a, int64_t g, const GroupQ<T>& q) {
  using D = DT<T>;
  if (a.scales) reinterpret_cast<T*>(a.scales)[g] = D::from_f(q.scale);
  if (a.zeros && a.has_zp) reinterpret_cast<T*>(a.zeros)[g] = D::from_f(q.zp);
  if (a.meta) {
    const int gi = static_cast<int>(g);
    const int row = gi / a.groups_per_row, kg = gi - row * a.groups_per_row;
    T* m = reinterpret_cast<T*>(a.meta) + (static_cast<int64_t>(kg) * a.rows + row) * 2;
    m[0] = D::from_f(q.scale);
    m[1] = D::from_f(q.zp);
  }
}

// ------------------------------------------------------------------ K1: L lanes x V vectors per group
// PB: 0 = no packed output, 1 = packed, any bit width (runtime), 4 = packed W4 from a 16-bit dtype
template <typename T, int L, int V, int MATH, int PB>
__global__ void __launch_bounds__(256) k1_groups(const QuantArgs a) {
  constexpr bool PACK = PB != 0;
  constexpr int E = DT<T>::kVec;
  constexpr int GPW = 32 / L;                       // groups per warp step
  const int lane = threadIdx.x & 31;
  const int l = lane % L, sub = lane / L;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const uint4* __restrict__ xin = reinterpret_cast<const uint4*>(a.x);
  uint4* __restrict__ dq = reinterpret_cast<uint4*>(a.deq);

  int64_t g = warp * GPW + sub;
  const int64_t gstep = n_warps * GPW;
  const int64_t g_end = ((a.n_groups + GPW - 1) / GPW) * GPW;   // warp-uniform trip count

  uint4 cur[V];
#pragma unroll
  for (int k = 0; k < V; ++k)
    cur[k] = (g < a.n_groups) ? ldg_stream(xin + g * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);

  for (; g - sub < g_end; g += gstep) {
    const bool valid = g < a.n_groups;
    // software prefetch of the next step (keeps >= 2V 128-bit loads in flight per lane)
    uint4 nxt[V];
    const int64_t gn = g + gstep;
#pragma unroll
    for (int k = 0; k < V; ++k)
      nxt[k] = (gn < a.n_groups) ? ldg_stream(xin + gn * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);

    float mn, mx;
    vec_minmax<T>(cur[0], mn, mx);
#pragma unroll
    for (int k = 1; k < V; ++k) {
      float m0, m1;
      vec_minmax<T>(cur[k], m0, m1);
      mn = min_nan(mn, m0);
      mx = max_nan(mx, m1);
    }
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
      mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    const GroupQ<T> q = group_params<T>(mn, mx, a, valid ? g : 0);

    int row = 0, vec0 = 0;
    if (PACK) {                                         // 32-bit: n_groups < 2^31 is checked on the host
      const int gi = static_cast<int>(g);
      row = gi / a.groups_per_row;
      vec0 = (gi - row * a.groups_per_row) * (L * V);
    }
    // magic-add rint is valid when every |x / scale| of the warp's groups is small (and finite)
    bool fast = false;
    if (MATH == MATH_PACKED && DT<T>::kCode != Q100_F32) {
      const float lim = fmaxf(fabsf(mn), fabsf(mx)) * q.rcp;
      fast = __all_sync(0xffffffffu, lim < FastLimit<T>::value);        // false for NaN
    }
    auto body = [&](auto deq_tag, auto fast_tag) {
      constexpr bool kDeq = decltype(deq_tag)::value, kFast = decltype(fast_tag)::value;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        uint32_t pairs[E / 2];
        const uint4 o = VecQuant<T, MATH>::template run<kDeq, PACK, kFast>(cur[k], q, a, pairs);
        if (kDeq && valid) stg_stream(dq + g * (L * V) + k * L + l, o);
        if (PB == 4) { if (valid && row < a.rows) emit_packed_w4(*reinterpret_cast<uint32_t(*)[4]>(&pairs), a, row, vec0 + k * L + l); }
        else if (PACK) emit_packed<E>(pairs, a, row, vec0 + k * L + l, valid && row < a.rows);
      }
    };
    using TT = std::true_type; using FF = std::false_type;
    if (dq != nullptr) { if (fast) body(TT{}, TT{}); else body(TT{}, FF{}); }
    else { if (fast) body(FF{}, TT{}); else body(FF{}, FF{}); }
    if (l == 0 && valid) store_group_meta<T>(a, g, q);
#pragma unroll
    for (int k = 0; k < V; ++k) cur[k] = nxt[k];
  }
}

// ------------------------------------------------------------------ K1 fast: L lanes x V vectors per group, 16-bit T
// Same results as k1_groups<T, 8, V, MATH_PACKED, PB>, about a third of its instructions (profiles/r01h:
// k1_groups is ALU-pipe bound at 28 thread instructions per element, most of them the NaN-aware fp32 min/max
// chains, three div.rn.f32 sequences per group and the per-step index divisions).  Differences:
//   * min/max stay packed: one T2 holds (min, -max) so a single HMNMX2.NAN per butterfly step reduces both;
//   * group parameters use the Markstein division q = x*rcp(d); q += fma(-q, d, x)*rcp(d).  For the scale
//     (x = range, d = 2^n - 1 or 2^(n-1) - 1) the 16-bit result equals rd(div.rn.f32) for every finite
//     16-bit x >= 0 (checked exhaustively, tools/check_markstein.py); for the zero-point only
//     rint(clamp(.)) is consumed, which q100_selftest_div covers;
//   * the element pass works in the shifted integer domain t = rd(v + C) = C + rint(v):
//     c' = clamp(t, C + qmin - zp, C + qmax - zp) = C + (x_int - zp), dequant = (c' - C) * scale,
//     code = c' - C + zp - qmin.  Every step is exact integer arithmetic when |v| and |zp| are small;
//   * x / scale for two elements at a time with {mul,fma}.rn.f32x2.
// A warp step whose groups do not qualify (NaN/inf, |x/scale| or |zp| too large for the magic-add rint)
// runs the generic group_params + VecQuant path instead (warp-uniform branch), so coverage is unchanged.
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long mul_f2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b,
                                                     unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// loop-invariant constants of the element pass and of the half-precision parameter clamps
template <typename T>
struct FastK {
  typename Packed16<T>::T2 magic, range, ktop, one, k_nlo, k_top, s_lo, s_hi, z_hi, sym_zp;
};
template <typename T>
struct FastQ {
  typename Packed16<T>::T2 magic, nlo, range, top, ktop, scale, one;
  unsigned long long rcp2, nscale2;
};

template <typename T2>
__device__ __forceinline__ uint32_t bits_of(const T2& v) { return *reinterpret_cast<const uint32_t*>(&v); }
template <typename T2>
__device__ __forceinline__ T2 from_bits(uint32_t v) { return *reinterpret_cast<const T2*>(&v); }

template <typename T>
__device__ __forceinline__ FastK<T> fast_invariants(const QuantArgs& a) {
  using D = DT<T>;
  using P = Packed16<T>;
  constexpr float C = P::kRintMagic;
  FastK<T> k;
  // all of these are integers of magnitude < 2^(significand bits): exact in T
  k.magic = P::bcast(C);
  k.range = P::bcast(a.qmax - a.qmin);
  k.ktop = P::bcast(CodeMagic<T>::kMagic + (a.qmax - a.qmin));
  k.one = P::bcast(1.0f);
  k.k_nlo = P::bcast(-(C + a.qmin));
  k.k_top = P::bcast(a.qmax);
  // rd(clamp(v, lo, hi)) == clamp(v, rd(lo), rd(hi)) for a value v of T (rounding is monotonic)
  k.s_lo = P::bcast(1e-5f);
  k.s_hi = P::bcast(1e4f);
  k.z_hi = P::bcast(1e4f);
  k.sym_zp = P::bcast(a.denom);
  return k;
}

// quantizer.py:143-159 for finite, well-scaled groups; `m` = (min, -max) of the group in T.  The fp32 -> T
// conversions run on the XU pipe at 16 results/clk/SM and were the kernel's limiter (profiles/r01i), so
// everything that is exact in T stays in T: the range, the clamps (pre-rounded bounds), the zero-point's rint and
// the per-group constants of the element pass.  Two conversions per group remain (scale, zero-point).
// Returns false when the step must take the generic path; q / f are then unspecified.
template <typename T>
__device__ __forceinline__ bool fast_params(typename Packed16<T>::T2 m, const QuantArgs& a, const FastK<T>& k,
                                            int64_t g, GroupQ<T>& q, FastQ<T>& f, uint32_t& sz_bits) {
  using D = DT<T>;
  using T2 = typename Packed16<T>::T2;
  const float inf = __int_as_float(0x7f800000);
  // packed T2 registers throughout (both halves carry the same value): no 16-bit scalar temporaries
  const uint32_t mb = bits_of(m);
  const float mn = __low2float(m), mx = -__high2float(m);
  float xmin = mn, num;
  if (a.lwc_up != nullptr) {
    const float xmax = D::rd(D::to_f(reinterpret_cast<const T*>(a.lwc_up)[g]) * mx);
    xmin = D::rd(D::to_f(reinterpret_cast<const T*>(a.lwc_low)[g]) * mn);
    num = a.symmetric ? fmaxf(fabsf(xmax), fabsf(xmin)) : D::rd(xmax - xmin);
  } else if (a.symmetric) {
    num = fmaxf(fabsf(mx), fabsf(mn));
  } else {
    const T2 d2 = __hadd2(m, from_bits<T2>(__byte_perm(mb, 0, 0x1032)));   // min + (-max) in both halves
    num = -__low2float(d2);                                                // rd(max - min), one op in T
  }
  float s = num * a.rcp_denom;                                     // Q100_FLAG_CUDA_SCALE: this product is the scale
  if (!a.cuda_scale) s = fmaf(fmaf(-s, a.denom, num), a.rcp_denom, s);
  const T2 s2 = __hmin2(__hmax2(from_bits<T2>(D::pack2(s, s)), k.s_lo), k.s_hi);
  q.scale = __low2float(s2);
  q.rcp = __frcp_rn(q.scale);
  T2 zp2 = from_bits<T2>(0u);
  if (a.has_zp) {
    if (a.symmetric) zp2 = k.sym_zp;
    else {
      const float nx = -xmin;
      float z = nx * q.rcp;
      z = fmaf(fmaf(-z, q.scale, nx), q.rcp, z);
      z = copysignf(z, nx);                                // the correction step turns -0 into +0
      zp2 = h2rint(__hmin2(__hmax2(from_bits<T2>(D::pack2(z, z)), __hneg2(k.z_hi)), k.z_hi));
    }
  }
  q.zp = __low2float(zp2);
  sz_bits = __byte_perm(bits_of(s2), bits_of(zp2), 0x5410);              // scale | zero-point << 16
  f.magic = k.magic; f.range = k.range; f.ktop = k.ktop; f.one = k.one;
  f.nlo = __hadd2(zp2, k.k_nlo);                           // -(C + qmin - zp)
  f.top = __hsub2(k.k_top, zp2);                           // qmax - zp
  f.scale = s2;
  f.rcp2 = pack_f2(q.rcp, q.rcp);
  f.nscale2 = pack_f2(-q.scale, -q.scale);
  const float lim = fmaxf(fabsf(mn), fabsf(mx)) * q.rcp;
  // all comparisons are false for NaN; num < inf rejects an overflowed range
  return (lim < FastLimit<T>::value) && (fabsf(q.zp) < FastLimit<T>::value) && (num < inf) && (fabsf(xmin) < inf);
}

// the same per-group constants from fp32 group parameters (K2: one set per row and thread)
template <typename T>
__device__ __forceinline__ FastQ<T> fast_from_group(const GroupQ<T>& q, const QuantArgs& a) {
  using P = Packed16<T>;
  constexpr float C = P::kRintMagic;
  FastQ<T> f;
  f.magic = P::bcast(C);
  f.nlo = P::bcast(-(C + a.qmin - q.zp));
  f.range = P::bcast(a.qmax - a.qmin);
  f.top = P::bcast(a.qmax - q.zp);
  f.ktop = P::bcast(CodeMagic<T>::kMagic + (a.qmax - a.qmin));
  f.scale = P::bcast(q.scale);
  f.one = P::bcast(1.0f);
  f.rcp2 = pack_f2(q.rcp, q.rcp);
  f.nscale2 = pack_f2(-q.scale, -q.scale);
  return f;
}

// Clamp without the ALU pipe (HMNMX2 shares it with the integer work; profiles/r01i): with t = C + rint(v),
//   u = relu(t - (C + qmin - zp)) = max(x_int, qmin) - qmin,   w = relu((qmax - qmin) - u),
//   code = (qmax - qmin) - w,   x_int - zp = (qmax - zp) - w          (exact integer arithmetic throughout).
// Returns the dequantised pair; `pair` (WANT_CODES) = bits of (kMagic + code) in both halves, NOT masked: the
// W4 packer removes the constant part arithmetically, the generic packer masks.
template <typename T, bool WANT_DEQ, bool WANT_CODES>
__device__ __forceinline__ uint32_t fast_pair(uint32_t in, const FastQ<T>& f, uint32_t& pair) {
  using D = DT<T>;
  using T2 = typename Packed16<T>::T2;
  const float2 xf = D::to_f2(in);
  const unsigned long long x2 = pack_f2(xf.x, xf.y);
  unsigned long long q2 = mul_f2(x2, f.rcp2);
  const unsigned long long r2 = fma_f2(q2, f.nscale2, x2);
  q2 = fma_f2(r2, f.rcp2, q2);
  float q0, q1;
  unpack_f2(q2, q0, q1);
  const T2 v = from_bits<T2>(D::pack2(q0, q1));                  // rd(x / scale)
  const T2 t = __hadd2(v, f.magic);                              // C + rint(v)
  const T2 u = __hfma2_relu(t, f.one, f.nlo);
  const T2 w = __hfma2_relu(__hneg2(u), f.one, f.range);
  if (WANT_CODES) pair = bits_of(__hsub2(f.ktop, w));            // kMagic + code
  if (!WANT_DEQ) return 0u;
  return bits_of(__hmul2(__hsub2(f.top, w), f.scale));
}

// W4 from fast_pair's unmasked pairs m_j = MAGICBITS + code pair j:  sum_j m_j << 4j  carries the constant
// MAGICBITS * 0x1111 (mod 2^32) on top of the packed word; IMADs on the FMA pipe, no masks.
template <typename T>
__device__ __forceinline__ uint32_t w4_word_from_magic(const uint32_t (&m)[4]) {
  constexpr uint32_t kBits = DT<T>::kCode == Q100_F16 ? 0x64006400u : 0x43004300u;
  constexpr uint32_t kConst = kBits * 0x1111u;
  uint32_t w = m[1] * 16u + m[0];
  w = m[2] * 256u + w;
  w = m[3] * 4096u + w;
  return w - kConst;
}

template <typename T, int L, int V, int PB>
__global__ void __launch_bounds__(256) k1_fast(const QuantArgs a) {
  using D = DT<T>;
  using T2 = typename Packed16<T>::T2;
  constexpr bool PACK = PB != 0;
  constexpr int E = 8, GPW = 32 / L;
  const int lane = threadIdx.x & 31;
  const int l = lane % L, sub = lane / L;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const uint4* __restrict__ xin = reinterpret_cast<const uint4*>(a.x);
  uint4* __restrict__ dq = reinterpret_cast<uint4*>(a.deq);
  const bool want_deq = dq != nullptr;

  int64_t g = warp * GPW + sub;
  const int64_t gstep = n_warps * GPW;
  const int64_t g_end = ((a.n_groups + GPW - 1) / GPW) * GPW;   // warp-uniform trip count

  // (row, k-group) of g, advanced incrementally (n_groups < 2^31 is checked on the host)
  int row = static_cast<int>(g / a.groups_per_row);
  int kg = static_cast<int>(g - static_cast<int64_t>(row) * a.groups_per_row);
  const int step_row = static_cast<int>(gstep / a.groups_per_row);
  const int step_kg = static_cast<int>(gstep - static_cast<int64_t>(step_row) * a.groups_per_row);

  const FastK<T> fk = fast_invariants<T>(a);

  // one warp step: quantise the groups held in `cur` while the next step's vectors land in `nxt`
  auto step = [&](uint4 (&cur)[V], uint4 (&nxt)[V]) {
    const bool valid = g < a.n_groups;
    const int64_t gn = g + gstep;
#pragma unroll
    for (int k = 0; k < V; ++k)
      nxt[k] = (gn < a.n_groups) ? ldg_stream(xin + gn * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);

    // ---- min / max of the lane's V vectors, kept packed
    T2 lo, hi;
    {
      const T2* h = reinterpret_cast<const T2*>(&cur[0]);
      lo = __hmin2_nan(__hmin2_nan(h[0], h[1]), __hmin2_nan(h[2], h[3]));
      hi = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
#pragma unroll
      for (int k = 1; k < V; ++k) {
        const T2* hk = reinterpret_cast<const T2*>(&cur[k]);
        lo = __hmin2_nan(lo, __hmin2_nan(__hmin2_nan(hk[0], hk[1]), __hmin2_nan(hk[2], hk[3])));
        hi = __hmax2_nan(hi, __hmax2_nan(__hmax2_nan(hk[0], hk[1]), __hmax2_nan(hk[2], hk[3])));
      }
    }
    const uint32_t lob = bits_of(lo), nhib = bits_of(__hneg2(hi));
    T2 m = __hmin2_nan(from_bits<T2>(__byte_perm(lob, nhib, 0x5410)),
                       from_bits<T2>(__byte_perm(lob, nhib, 0x7632)));       // (min, -max)
#pragma unroll
    for (int d = 1; d < L; d <<= 1)
      m = __hmin2_nan(m, from_bits<T2>(__shfl_xor_sync(0xffffffffu, bits_of(m), d)));

    GroupQ<T> q;
    FastQ<T> f;
    uint32_t sz_bits;                                            // (scale, zero-point) as two T values
    const bool ok = fast_params<T>(m, a, fk, valid ? g : 0, q, f, sz_bits);
    const int vec0 = kg * (L * V) + l;
    // W4: this lane's first word of the group inside the tile-major image (row's 32 bytes of k tile kg*V*L/8)
    uint32_t* w4row = nullptr;
    if (PB == 4)
      w4row = a.packed + (static_cast<int64_t>(row >> 7) * a.k_tiles + kg * ((L * V) >> 3)) * (128 * 8) + (row & 127) * 8 + l;
    if (__all_sync(0xffffffffu, ok)) {
      auto body = [&](auto deq_tag) {
        constexpr bool kDeq = decltype(deq_tag)::value;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          uint32_t pairs[4];
          uint4 o;
          o.x = fast_pair<T, kDeq, PACK>(cur[k].x, f, pairs[0]);
          o.y = fast_pair<T, kDeq, PACK>(cur[k].y, f, pairs[1]);
          o.z = fast_pair<T, kDeq, PACK>(cur[k].z, f, pairs[2]);
          o.w = fast_pair<T, kDeq, PACK>(cur[k].w, f, pairs[3]);
          if (kDeq && valid) stg_stream(dq + g * (L * V) + k * L + l, o);
          if (PB == 4) {
            // vector k*L + l of the group: k tile (k*L) >> 3, word ((k*L) & 7) + l of the row's 8 -- constants
            if (valid) w4row[((k * L) >> 3) * (128 * 8) + ((k * L) & 7)] = w4_word_from_magic<T>(pairs);
          } else if (PACK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pairs[j] &= CodeMagic<T>::kMask;
            emit_packed<E>(pairs, a, row, vec0 + k * L, valid);
          }
        }
      };
      if (want_deq) body(std::true_type{}); else body(std::false_type{});
    } else {
      q = group_params<T>(__low2float(m), -__high2float(m), a, valid ? g : 0);
      sz_bits = D::pack2(q.scale, q.zp);
      auto body = [&](auto deq_tag) {
        constexpr bool kDeq = decltype(deq_tag)::value;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          uint32_t pairs[4];
          const uint4 o = VecQuant<T, MATH_PACKED>::template run<kDeq, PACK, false>(cur[k], q, a, pairs);
          if (kDeq && valid) stg_stream(dq + g * (L * V) + k * L + l, o);
          if (PB == 4) { if (valid) emit_packed_w4(pairs, a, row, vec0 + k * L); }
          else if (PACK) emit_packed<E>(pairs, a, row, vec0 + k * L, valid);
        }
      };
      if (want_deq) body(std::true_type{}); else body(std::false_type{});
    }
    if (l == 0 && valid) {
      if (a.scales) reinterpret_cast<uint16_t*>(a.scales)[g] = static_cast<uint16_t>(sz_bits);
      if (a.zeros && a.has_zp) reinterpret_cast<uint16_t*>(a.zeros)[g] = static_cast<uint16_t>(sz_bits >> 16);
      if (a.meta) {
        T* mp = reinterpret_cast<T*>(a.meta) + (static_cast<int64_t>(kg) * a.rows + row) * 2;
        *reinterpret_cast<uint32_t*>(mp) = sz_bits;
      }
    }
    g = gn;
    row += step_row;
    kg += step_kg;
    if (kg >= a.groups_per_row) { kg -= a.groups_per_row; ++row; }
  };

  // two register buffers used alternately (no copies between steps)
  uint4 buf_a[V], buf_b[V];
#pragma unroll
  for (int k = 0; k < V; ++k)
    buf_a[k] = (g < a.n_groups) ? ldg_stream(xin + g * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);
  if (!(g - sub < g_end)) return;
  for (;;) {
    step(buf_a, buf_b);
    if (!(g - sub < g_end)) break;
    step(buf_b, buf_a);
    if (!(g - sub < g_end)) break;
  }
}

// ------------------------------------------------------------------ K2: one CTA per row
template <typename T, int VMAX, int MATH, bool PACK>
__global__ void __launch_bounds__(1024) k2_rows(const QuantArgs a) {
  constexpr int E = DT<T>::kVec;
  __shared__ float s_mn[32], s_mx[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int nvec = static_cast<int>(a.cols / E);
  const int vper = (nvec + nt - 1) / nt;            // uniform trip count <= VMAX
  const uint4* __restrict__ xin = reinterpret_cast<const uint4*>(a.x);
  uint4* __restrict__ dq = reinterpret_cast<uint4*>(a.deq);

  for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x) {
    uint4 v[VMAX];
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);   // +inf / -inf
    if constexpr (DT<T>::kCode != Q100_F32) {
      // 16-bit: the thread's vectors are reduced in packed arithmetic (HMNMX2.NAN), one conversion at the end
      using T2 = typename Packed16<T>::T2;
      T2 lo = Packed16<T>::bcast(mn), hi = Packed16<T>::bcast(mx);
#pragma unroll
      for (int k = 0; k < VMAX; ++k) {
        const int idx = k * nt + tid;
        if (k < vper && idx < nvec) {
          v[k] = ldg_stream(xin + row * nvec + idx);
          const T2* h = reinterpret_cast<const T2*>(&v[k]);
          lo = __hmin2_nan(lo, __hmin2_nan(__hmin2_nan(h[0], h[1]), __hmin2_nan(h[2], h[3])));
          hi = __hmax2_nan(hi, __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3])));
        }
      }
      mn = min_nan(__low2float(lo), __high2float(lo));
      mx = max_nan(__low2float(hi), __high2float(hi));
    } else {
#pragma unroll
      for (int k = 0; k < VMAX; ++k) {
        const int idx = k * nt + tid;
        if (k < vper && idx < nvec) {
          v[k] = ldg_stream(xin + row * nvec + idx);
          float m0, m1;
          vec_minmax<T>(v[k], m0, m1);
          mn = min_nan(mn, m0);
          mx = max_nan(mx, m1);
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) { s_mn[wid] = mn; s_mx[wid] = mx; }
    __syncthreads();
    mn = (lane < nw) ? s_mn[lane] : __int_as_float(0x7f800000);
    mx = (lane < nw) ? s_mx[lane] : __int_as_float(0xff800000);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    __syncthreads();                                   // s_mn/s_mx reusable next row
    if (a.minmax_out != nullptr) {                     // statistics pass (q100_rows_minmax)
      if (tid == 0) {
        T* o = reinterpret_cast<T*>(a.minmax_out) + 2 * row;
        o[0] = DT<T>::from_f(mn);
        o[1] = DT<T>::from_f(mx);
      }
      continue;
    }
    if (a.ext_minmax != nullptr) {                     // statistics of the full (sharded) row, from the caller
      const T* e = reinterpret_cast<const T*>(a.ext_minmax) + 2 * row;
      mn = DT<T>::to_f(e[0]);
      mx = DT<T>::to_f(e[1]);
    }
    const GroupQ<T> q = group_params<T>(mn, mx, a, row);
    // rows whose |x / scale| and |zp| fit the exact-integer window take k1_fast's element pass (CTA-uniform)
    if constexpr (MATH == MATH_PACKED && DT<T>::kCode != Q100_F32) {
      const float lim = fmaxf(fabsf(mn), fabsf(mx)) * q.rcp;                 // NaN / inf: false
      const bool code_ok = !PACK || DT<T>::kCode == Q100_F16 || a.n_bits <= 4;  // bf16 cannot carry 128 + an 8-bit code
      if (lim < FastLimit<T>::value && fabsf(q.zp) < FastLimit<T>::value && code_ok) {
        const FastQ<T> f = fast_from_group<T>(q, a);
#pragma unroll
        for (int k = 0; k < VMAX; ++k) {
          if (k < vper) {
            const int idx = k * nt + tid;
            const bool valid = idx < nvec;
            uint32_t pairs[4] = {0u, 0u, 0u, 0u};
            uint4 o = make_uint4(0, 0, 0, 0);
            if (valid) {
              o.x = fast_pair<T, true, PACK>(v[k].x, f, pairs[0]);
              o.y = fast_pair<T, true, PACK>(v[k].y, f, pairs[1]);
              o.z = fast_pair<T, true, PACK>(v[k].z, f, pairs[2]);
              o.w = fast_pair<T, true, PACK>(v[k].w, f, pairs[3]);
            }
            if (dq != nullptr && valid) stg_stream(dq + row * nvec + idx, o);
            if (PACK) {
#pragma unroll
              for (int j = 0; j < 4; ++j) pairs[j] &= CodeMagic<T>::kMask;
              emit_packed<8>(pairs, a, static_cast<int>(row), idx, valid);
            }
          }
        }
        if (tid == 0) store_group_meta<T>(a, row, q);
        continue;
      }
    }
#pragma unroll
    for (int k = 0; k < VMAX; ++k) {
      if (k < vper) {
        const int idx = k * nt + tid;
        const bool valid = idx < nvec;
        uint32_t pairs[E / 2];
        uint4 o = make_uint4(0, 0, 0, 0);
        if (valid) o = VecQuant<T, MATH>::template run<true, PACK>(v[k], q, a, pairs);
        else {
#pragma unroll
          for (int j = 0; j < E / 2; ++j) pairs[j] = 0;
        }
        if (dq != nullptr && valid) stg_stream(dq + row * nvec + idx, o);
        if (PACK) emit_packed<E>(pairs, a, static_cast<int>(row), idx, valid);
      }
    }
    if (tid == 0) store_group_meta<T>(a, row, q);
  }
}

// ------------------------------------------------------------------ dequantize (utility, not hot)
template <typename T>
__global__ void k_dequantize(const uint32_t* __restrict__ packed, const T* __restrict__ scales,
                             const T* __restrict__ zeros, T* __restrict__ out, int64_t rows,
                             int64_t cols, int group, int n_bits, float qmin, int k_tiles, int tile_words) {
  using D = DT<T>;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols, c = i % cols;
  const uint32_t* tile = packed + ((r >> 7) * k_tiles + (c >> 6)) * tile_words;
  const int rr = static_cast<int>(r & 127), ek = static_cast<int>(c & 63);
  auto field = [&](int b, int plane_off) -> uint32_t {
    const int F = 32 / b;
    const int e = ek % F;
    const uint32_t w = tile[plane_off + rr * (2 * b) + ek / F];
    return (w >> (b * (e >> 1) + 16 * (e & 1))) & ((1u << b) - 1u);
  };
  uint32_t u;
  if (n_bits == 3) u = field(2, 0) | (field(1, 128 * 4) << 2);
  else u = field(n_bits, 0);
  float code = static_cast<float>(u) + qmin;
  if (scales == nullptr) { out[i] = D::from_f(code); return; }
  const int64_t g = i / group;
  const float s = D::to_f(scales[g]);
  float d = code;
  if (zeros != nullptr) d = D::rd(code - D::to_f(zeros[g]));
  out[i] = D::from_f(D::rd(d * s));
}

// ------------------------------------------------------------------ division self-test
template <typename T>
__global__ void k_selftest_div(unsigned long long* mismatches) {
  using D = DT<T>;
  // blockIdx.x enumerates scale bit patterns, threads sweep x bit patterns
  const uint32_t sbits = blockIdx.x;
  uint16_t sb = static_cast<uint16_t>(sbits);
  const T st = *reinterpret_cast<T*>(&sb);
  const float s = D::to_f(st);
  if (!(s >= 1e-5f && s <= 1e4f)) return;
  const float rcp = __frcp_rn(s);
  unsigned long long bad = 0;
  for (uint32_t xb = threadIdx.x; xb < 65536u; xb += blockDim.x) {
    uint16_t xs = static_cast<uint16_t>(xb);
    const float x = D::to_f(*reinterpret_cast<T*>(&xs));
    const float ref = D::rd(__fdiv_rn(x, s));
    const float got = D::rd(div_x_s<MATH_PACKED>(x, s, rcp));
    // the quantizer only consumes round_ste(v): compare what it can observe
    const float r0 = D::rd(D::rd(rintf(ref) - ref) + ref);
    const float r1 = D::rd(D::rd(rintf(got) - got) + got);
    const bool same = (r0 == r1) || (r0 != r0 && r1 != r1);
    bad += same ? 0ull : 1ull;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------ host dispatch

template <typename T, int L, int V, int MATH>
static int launch_k1_lv(const QuantArgs& a, bool pack, cudaStream_t st) {
  constexpr int GPW = 32 / L;
  const int64_t warps_needed = (a.n_groups + GPW - 1) / GPW;
  const int sms = sm_count();
  int64_t blocks = (warps_needed + 7) / 8;
  const int64_t cap = static_cast<int64_t>(sms) * 8;          // persistent: <= 8 CTAs / SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  constexpr bool k16 = DT<T>::kCode != Q100_F32;
  if (!pack) k1_groups<T, L, V, MATH, 0><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  else if (k16 && MATH == MATH_PACKED && a.n_bits == 4) k1_groups<T, L, V, MATH, (k16 && MATH == MATH_PACKED) ? 4 : 1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  else k1_groups<T, L, V, MATH, 1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "k1_groups launch");
}

template <typename K>
static int resident_ctas(K kernel) {               // persistent grid = what the SMs can hold at once
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 2;
  return occ * sm_count();
}

template <typename T, int L, int V, int PB>
static int launch_k1_fast_pb(const QuantArgs& a, cudaStream_t st) {
  if constexpr (DT<T>::kCode != Q100_F32) {
    static const int cap = resident_ctas(k1_fast<T, L, V, PB>);
    constexpr int GPW = 32 / L;
    const int64_t warps_needed = (a.n_groups + GPW - 1) / GPW;
    int64_t blocks = (warps_needed + 7) / 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k1_fast<T, L, V, PB><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  }
  return check_cuda(cudaGetLastError(), "k1_fast launch");
}

template <typename T, int L, int V>
static int launch_k1_fast(const QuantArgs& a, bool pack, cudaStream_t st) {
  if (!pack) return launch_k1_fast_pb<T, L, V, 0>(a, st);
  if (a.n_bits == 4) return launch_k1_fast_pb<T, L, V, 4>(a, st);
  return launch_k1_fast_pb<T, L, V, 1>(a, st);
}

template <typename T, int MATH>
static int launch_k1(const QuantArgs& a, bool pack, cudaStream_t st, bool& handled, bool generic) {
  constexpr int E = DT<T>::kVec;
  const int vecs = a.group / E;               // vectors per group
  handled = true;
  if constexpr (E == 8 && MATH == MATH_PACKED) {
    // bf16 cannot carry an 8-bit code in the mantissa of 128 + code: those packs stay on k1_groups
    const bool bf16_w8_pack = pack && a.n_bits > 4 && DT<T>::kCode == Q100_BF16;
    if (!generic && !bf16_w8_pack) {
      switch (vecs) {
        // 32 elements per lane and step: the per-group work (reduction, parameters, index bookkeeping) is
        // amortised over twice as many elements as with 8 lanes x 2 vectors
        case 8:  return launch_k1_fast<T, 4, 2>(a, pack, st);       // g = 64
        case 16: return launch_k1_fast<T, 4, 4>(a, pack, st);       // g = 128
        case 32: return launch_k1_fast<T, 8, 4>(a, pack, st);       // g = 256
        default: break;
      }
    }
  }
  if (E == 8) {
    switch (vecs) {
      case 4:  return launch_k1_lv<T, 4, 1, MATH>(a, pack, st);    // g = 32
      case 8:  return launch_k1_lv<T, 8, 1, MATH>(a, pack, st);    // g = 64
      case 16: return launch_k1_lv<T, 8, 2, MATH>(a, pack, st);    // g = 128
      case 32: return launch_k1_lv<T, 8, 4, MATH>(a, pack, st);    // g = 256
      default: break;
    }
  } else {
    switch (vecs) {
      case 8:  return launch_k1_lv<T, 8, 1, MATH>(a, pack, st);    // g = 32
      case 16: return launch_k1_lv<T, 8, 2, MATH>(a, pack, st);    // g = 64
      case 32: return launch_k1_lv<T, 8, 4, MATH>(a, pack, st);    // g = 128
      case 64: return launch_k1_lv<T, 16, 4, MATH>(a, pack, st);   // g = 256
      default: break;
    }
  }
  handled = false;
  return Q100_OK;
}

template <typename T, int MATH>
static int launch_k2(const QuantArgs& a_in, bool pack, cudaStream_t st) {
  constexpr int E = DT<T>::kVec;
  constexpr int VMAX = 8;
  // rows kernel treats every group as a row of `group` elements
  QuantArgs a = a_in;
  const int64_t nvec = a.group / E;
  Q100_REQUIRE(nvec <= VMAX * 1024, Q100_EUNSUPPORTED,
               "row/group length %d exceeds the register-resident limit of %d elements",
               a.group, VMAX * 1024 * E);
  if (pack)
    Q100_REQUIRE(a.groups_per_row == 1, Q100_EUNSUPPORTED,
                 "packed output needs group in {32,64,128,256} or one group per row (got %d)", a.group);
  int nt = 64;
  while (nt < 1024 && static_cast<int64_t>(nt) * 4 < nvec) nt <<= 1;
  while (static_cast<int64_t>(nt) * VMAX < nvec) nt <<= 1;
  // present the flat [n_groups, group] view to the kernel: every group is a "row"
  a.rows = a.n_groups;
  a.cols = a.group;
  a.groups_per_row = 1;
  const QuantArgs& k = a;
  const int sms = sm_count();
  int64_t blocks = a.n_groups;
  const int64_t cap = static_cast<int64_t>(sms) * (2048 / nt);
  if (blocks > cap) blocks = cap;
  if (pack) k2_rows<T, VMAX, MATH, true><<<static_cast<unsigned>(blocks), nt, 0, st>>>(k);
  else k2_rows<T, VMAX, MATH, false><<<static_cast<unsigned>(blocks), nt, 0, st>>>(k);
  return check_cuda(cudaGetLastError(), "k2_rows launch");
}

template <typename T>
static int quantize_dispatch(const QuantArgs& a, bool pack, int flags, cudaStream_t st) {
  constexpr bool k16 = (DT<T>::kCode != Q100_F32);
  const int math = !k16 ? MATH_SCALAR_IEEE
                        : (flags & Q100_FLAG_IEEE_DIV)    ? MATH_SCALAR_IEEE
                        : (flags & Q100_FLAG_SCALAR_MATH) ? MATH_SCALAR
                                                          : MATH_PACKED;
  bool handled = false;
  int rc = Q100_OK;
  const bool generic = (flags & Q100_FLAG_K1_GENERIC) != 0;
  if (math == MATH_SCALAR_IEEE) rc = launch_k1<T, MATH_SCALAR_IEEE>(a, pack, st, handled, generic);
  else if (math == MATH_SCALAR) rc = launch_k1<T, k16 ? MATH_SCALAR : MATH_SCALAR_IEEE>(a, pack, st, handled, generic);
  else rc = launch_k1<T, k16 ? MATH_PACKED : MATH_SCALAR_IEEE>(a, pack, st, handled, generic);
  if (handled) return rc;
  // everything else goes through the one-CTA-per-group row kernel, which cannot address
  // the transposed side band (or packed rows) of a multi-group row
  Q100_REQUIRE(a.meta == nullptr || a.groups_per_row == 1, Q100_EUNSUPPORTED,
               "meta side band needs group in {32,64,128,256} or one group per row (got %d)", a.group);
  if (math == MATH_SCALAR_IEEE) return launch_k2<T, MATH_SCALAR_IEEE>(a, pack, st);
  if (math == MATH_SCALAR) return launch_k2<T, k16 ? MATH_SCALAR : MATH_SCALAR_IEEE>(a, pack, st);
  return launch_k2<T, k16 ? MATH_PACKED : MATH_SCALAR_IEEE>(a, pack, st);
}

}  // namespace q100

using namespace q100;

extern "C" int64_t q100_packed_words(int64_t rows, int64_t cols, int n_bits) {
  if (n_bits != 2 && n_bits != 3 && n_bits != 4 && n_bits != 8) return -1;
  if (rows < 0 || cols < 0 || cols % 64 != 0) return -1;
  return ((rows + 127) / 128) * (cols / 64) * tile_words_of(n_bits);
}

extern "C" int q100_quantize_groupwise(const void* x, int dtype, int64_t rows, int64_t cols, int group,
                                       int n_bits, int symmetric, int disable_zp,
                                       const void* lwc_up, const void* lwc_low,
                                       void* dequant_out, void* packed_out,
                                       void* scales_out, void* zeros_out, void* meta_out,
                                       int flags, void* stream) {
  Q100_REQUIRE(dtype >= 0 && dtype <= 2, Q100_EINVAL, "bad dtype %d", dtype);
  Q100_REQUIRE(rows >= 0 && cols >= 0, Q100_EINVAL, "negative shape");
  Q100_REQUIRE(x != nullptr || rows == 0 || cols == 0, Q100_EINVAL, "x is NULL");
  Q100_REQUIRE(n_bits >= 2 && n_bits <= 15, Q100_EINVAL,
               "n_bits=%d: the quantizer handles 2..15 (>= 16 is the identity, quantizer.py:119)", n_bits);
  // 9..15 bits: qmax is not representable in bf16 and the reference's clamp then depends on
  // how torch casts the bound; no shipped configuration uses it (wbits 2/3/4/8, abits 4/8).
  Q100_REQUIRE(n_bits <= 8, Q100_EUNSUPPORTED, "n_bits=%d > 8 is not implemented", n_bits);
  Q100_REQUIRE((lwc_up == nullptr) == (lwc_low == nullptr), Q100_EINVAL,
               "lwc_up and lwc_low must be given together");
  if (rows == 0 || cols == 0) return Q100_OK;            // empty tensor: nothing to do
  const int g = (group == 0) ? static_cast<int>(cols) : group;
  Q100_REQUIRE(group >= 0 && cols % g == 0, Q100_EINVAL,
               "cols=%lld not divisible by group=%d (pad first, quantizer.py:136-139)",
               static_cast<long long>(cols), g);
  const int esz = (dtype == Q100_F32) ? 4 : 2;
  Q100_REQUIRE((static_cast<int64_t>(g) * esz) % 16 == 0, Q100_EUNSUPPORTED,
               "group/row of %d elements is not a multiple of 16 bytes", g);
  Q100_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(dequant_out) & 15) == 0, Q100_EINVAL,
               "x / dequant_out must be 16-byte aligned");
  const bool pack = packed_out != nullptr;
  if (pack) {
    Q100_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED,
                 "packed output supports n_bits in {2,3,4,8}, got %d", n_bits);
    Q100_REQUIRE(cols % 64 == 0, Q100_EUNSUPPORTED, "packed output needs cols %% 64 == 0");
  }
  Q100_REQUIRE(meta_out == nullptr || dtype != Q100_F32, Q100_EUNSUPPORTED,
               "meta side band is fp16/bf16 only");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;

  QuantArgs a{};
  a.x = x; a.deq = dequant_out; a.packed = reinterpret_cast<uint32_t*>(packed_out);
  a.scales = scales_out; a.zeros = zeros_out; a.meta = meta_out;
  a.lwc_up = lwc_up; a.lwc_low = lwc_low;
  a.rows = rows; a.cols = cols; a.group = g;
  a.groups_per_row = static_cast<int>(cols / g);
  a.n_groups = rows * a.groups_per_row;
  Q100_REQUIRE(a.n_groups < (1ll << 31), Q100_EUNSUPPORTED, "too many groups");
  a.n_bits = n_bits; a.symmetric = symmetric ? 1 : 0; a.has_zp = disable_zp ? 0 : 1;
  if (disable_zp) { a.qmin = -static_cast<float>(1 << (n_bits - 1)); a.qmax = static_cast<float>((1 << (n_bits - 1)) - 1); }
  else { a.qmin = 0.0f; a.qmax = static_cast<float>((1 << n_bits) - 1); }
  a.denom = symmetric ? static_cast<float>((1 << (n_bits - 1)) - 1) : static_cast<float>((1 << n_bits) - 1);
  a.rcp_denom = 1.0f / a.denom;
  a.cuda_scale = (flags & Q100_FLAG_CUDA_SCALE) ? 1 : 0;
  a.k_tiles = static_cast<int>(cols / 64);
  a.tile_words = pack ? tile_words_of(n_bits) : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case Q100_F16:  return quantize_dispatch<__half>(a, pack, flags, st);
    case Q100_BF16: return quantize_dispatch<__nv_bfloat16>(a, pack, flags, st);
    default:        return quantize_dispatch<float>(a, pack, flags, st);
  }
}

namespace {
int rows_entry(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits, int symmetric, int disable_zp,
               const void* ext_minmax, void* minmax_out, const void* lwc_up, const void* lwc_low,
               void* dequant_out, void* packed_out, void* scales_out, void* zeros_out, void* meta_out,
               int flags, void* stream) {
  Q100_REQUIRE(dtype >= 0 && dtype <= 2, Q100_EINVAL, "bad dtype %d", dtype);
  Q100_REQUIRE(rows >= 0 && cols >= 0, Q100_EINVAL, "negative shape");
  Q100_REQUIRE(n_bits >= 2 && n_bits <= 8, Q100_EUNSUPPORTED, "n_bits=%d: 2..8", n_bits);
  Q100_REQUIRE((lwc_up == nullptr) == (lwc_low == nullptr), Q100_EINVAL, "lwc_up and lwc_low must be given together");
  if (rows == 0 || cols == 0) return Q100_OK;
  Q100_REQUIRE(x != nullptr, Q100_EINVAL, "x is NULL");
  const int esz = (dtype == Q100_F32) ? 4 : 2;
  Q100_REQUIRE((cols * esz) % 16 == 0, Q100_EUNSUPPORTED, "row of %lld elements is not a multiple of 16 bytes",
               static_cast<long long>(cols));
  Q100_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dequant_out) & 15) == 0,
               Q100_EINVAL, "x / dequant_out must be 16-byte aligned");
  const bool pack = packed_out != nullptr;
  if (pack) {
    Q100_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED,
                 "packed output supports n_bits in {2,3,4,8}, got %d", n_bits);
    Q100_REQUIRE(cols % 64 == 0, Q100_EUNSUPPORTED, "packed output needs cols %% 64 == 0");
  }
  Q100_REQUIRE(meta_out == nullptr || dtype != Q100_F32, Q100_EUNSUPPORTED, "meta side band is fp16/bf16 only");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  QuantArgs a{};
  a.x = x; a.deq = dequant_out; a.packed = reinterpret_cast<uint32_t*>(packed_out);
  a.scales = scales_out; a.zeros = zeros_out; a.meta = meta_out;
  a.lwc_up = lwc_up; a.lwc_low = lwc_low;
  a.ext_minmax = ext_minmax; a.minmax_out = minmax_out;
  a.rows = rows; a.cols = cols; a.group = static_cast<int>(cols);
  a.groups_per_row = 1;
  a.n_groups = rows;
  Q100_REQUIRE(a.n_groups < (1ll << 31), Q100_EUNSUPPORTED, "too many rows");
  a.n_bits = n_bits; a.symmetric = symmetric ? 1 : 0; a.has_zp = disable_zp ? 0 : 1;
  if (disable_zp) { a.qmin = -static_cast<float>(1 << (n_bits - 1)); a.qmax = static_cast<float>((1 << (n_bits - 1)) - 1); }
  else { a.qmin = 0.0f; a.qmax = static_cast<float>((1 << n_bits) - 1); }
  a.denom = symmetric ? static_cast<float>((1 << (n_bits - 1)) - 1) : static_cast<float>((1 << n_bits) - 1);
  a.rcp_denom = 1.0f / a.denom;
  a.cuda_scale = (flags & Q100_FLAG_CUDA_SCALE) ? 1 : 0;
  a.k_tiles = static_cast<int>(cols / 64);
  a.tile_words = pack ? tile_words_of(n_bits) : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool k16 = dtype != Q100_F32;
  const int math = !k16 ? MATH_SCALAR_IEEE : (flags & Q100_FLAG_IEEE_DIV) ? MATH_SCALAR_IEEE
                        : (flags & Q100_FLAG_SCALAR_MATH) ? MATH_SCALAR : MATH_PACKED;
  switch (dtype) {
    case Q100_F16:
      if (math == MATH_SCALAR_IEEE) return launch_k2<__half, MATH_SCALAR_IEEE>(a, pack, st);
      if (math == MATH_SCALAR) return launch_k2<__half, MATH_SCALAR>(a, pack, st);
      return launch_k2<__half, MATH_PACKED>(a, pack, st);
    case Q100_BF16:
      if (math == MATH_SCALAR_IEEE) return launch_k2<__nv_bfloat16, MATH_SCALAR_IEEE>(a, pack, st);
      if (math == MATH_SCALAR) return launch_k2<__nv_bfloat16, MATH_SCALAR>(a, pack, st);
      return launch_k2<__nv_bfloat16, MATH_PACKED>(a, pack, st);
    default:
      return launch_k2<float, MATH_SCALAR_IEEE>(a, pack, st);
  }
}
}  // namespace

extern "C" int q100_rows_minmax(const void* x, int dtype, int64_t rows, int64_t cols, void* minmax_out, void* stream) {
  Q100_REQUIRE(minmax_out != nullptr || rows == 0, Q100_EINVAL, "minmax_out is NULL");
  return rows_entry(x, dtype, rows, cols, 8, 0, 0, nullptr, minmax_out, nullptr, nullptr, nullptr, nullptr, nullptr,
                    nullptr, nullptr, 0, stream);
}

extern "C" int q100_quantize_rows_stats(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits,
                                        int symmetric, int disable_zp, const void* minmax,
                                        const void* lwc_up, const void* lwc_low,
                                        void* dequant_out, void* packed_out, void* scales_out, void* zeros_out,
                                        void* meta_out, int flags, void* stream) {
  Q100_REQUIRE(minmax != nullptr || rows == 0, Q100_EINVAL, "minmax is NULL");
  return rows_entry(x, dtype, rows, cols, n_bits, symmetric, disable_zp, minmax, nullptr, lwc_up, lwc_low,
                    dequant_out, packed_out, scales_out, zeros_out, meta_out, flags, stream);
}

extern "C" int q100_fakequant_rows(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits,
                                   void* out, void* scales_out, void* zeros_out, int flags,
                                   void* stream) {
  return q100_quantize_groupwise(x, dtype, rows, cols, 0, n_bits, 0, 0, nullptr, nullptr, out,
                                 nullptr, scales_out, zeros_out, nullptr, flags, stream);
}

extern "C" int q100_dequantize(const void* packed, const void* scales, const void* zeros, int dtype,
                               int64_t rows, int64_t cols, int group, int n_bits, int disable_zp,
                               void* out, void* stream) {
  Q100_REQUIRE(packed && out, Q100_EINVAL, "NULL pointer");
  Q100_REQUIRE(dtype >= 0 && dtype <= 2, Q100_EINVAL, "bad dtype %d", dtype);
  Q100_REQUIRE(n_bits == 2 || n_bits == 3 || n_bits == 4 || n_bits == 8, Q100_EUNSUPPORTED,
               "n_bits in {2,3,4,8}");
  Q100_REQUIRE(cols % 64 == 0, Q100_EUNSUPPORTED, "cols %% 64 != 0");
  if (rows == 0 || cols == 0) return Q100_OK;
  const int g = group == 0 ? static_cast<int>(cols) : group;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  const float qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  const int kt = static_cast<int>(cols / 64), tw = tile_words_of(n_bits);
  const int64_t n = rows * cols;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t* p = reinterpret_cast<const uint32_t*>(packed);
  if (dtype == Q100_F16)
    k_dequantize<__half><<<blocks, 256, 0, st>>>(p, (const __half*)scales, (const __half*)zeros, (__half*)out, rows, cols, g, n_bits, qmin, kt, tw);
  else if (dtype == Q100_BF16)
    k_dequantize<__nv_bfloat16><<<blocks, 256, 0, st>>>(p, (const __nv_bfloat16*)scales, (const __nv_bfloat16*)zeros, (__nv_bfloat16*)out, rows, cols, g, n_bits, qmin, kt, tw);
  else
    k_dequantize<float><<<blocks, 256, 0, st>>>(p, (const float*)scales, (const float*)zeros, (float*)out, rows, cols, g, n_bits, qmin, kt, tw);
  return check_cuda(cudaGetLastError(), "k_dequantize launch");
}

extern "C" int q100_selftest_div(int dtype, void* mismatches_out, void* stream) {
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EINVAL, "fp16/bf16 only");
  Q100_REQUIRE(mismatches_out != nullptr, Q100_EINVAL, "NULL output");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == Q100_F16) k_selftest_div<__half><<<65536, 256, 0, st>>>((unsigned long long*)mismatches_out);
  else k_selftest_div<__nv_bfloat16><<<65536, 256, 0, st>>>((unsigned long long*)mismatches_out);
  return check_cuda(cudaGetLastError(), "k_selftest_div launch");
}
