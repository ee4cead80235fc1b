// q100_quant_bwd.cu -- K1b: backward of the group-wise fake-quantizer, one pass.
//
// The calibration loop of the reference (quant/omniquant.py:261-293) differentiates through
// UniformAffineQuantizer.forward (quant/quantizer.py:94-159): straight-through estimator for round(), true
// gradients through clamp, the group's scale (a function of its min / max and of the LWC multipliers) and
// nothing through the zero-point (it is `.round()`-ed without STE, quantizer.py:158).  Written out:
//
//   v = x / s;  t = round(v) + zp;  inside = qmin <= t <= qmax;  c = clamp(t);  y = (c - zp) * s
//   dy/dx      = inside                                    (direct path, STE)
//   dy/ds      = (c - zp) - inside * v                     (summed over the group -> g_s)
//   s          = clamp((xmax*su - xmin*sl) / (2^n - 1))    (asymmetric; su = sl = 1 without LWC)
//              = clamp(max(|xmax*su|, |xmin*sl|) / (2^(n-1) - 1))   (symmetric)
//   xmax / xmin = amax / amin of the group: the gradient is shared equally by the tied positions
//
// which is what quant/quantizer.py::fake_quant_backward computes with a dozen elementwise / reduction launches
// (about 40 bytes of HBM traffic per element in fp32).  Here the group (g elements of x and of the incoming
// gradient) is held in registers by L lanes x V vectors like in K1: 3 * sizeof(T) bytes per element.
// The quantisation decisions (v, t, inside, c) repeat the forward's literal per-op arithmetic (group_params and
// the MATH_SCALAR_IEEE element path), so the STE mask is bit-identical to the forward's; the gradient arithmetic
// itself runs in fp32 and is rounded once (the torch restatement rounds after every op for 16-bit dtypes).
#include "q100_quant.cuh"

namespace q100 {

struct BwdArgs {
  QuantArgs q;
  const void* gout;
  void* gx;
  void* gup;
  void* glow;
};

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int E = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <typename T> struct Vec16 {
  static constexpr int E = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const float2 a = DT<T>::to_f2(v.x), b = DT<T>::to_f2(v.y), c = DT<T>::to_f2(v.z), d = DT<T>::to_f2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
    return make_uint4(DT<T>::pack2(f[0], f[1]), DT<T>::pack2(f[2], f[3]), DT<T>::pack2(f[4], f[5]), DT<T>::pack2(f[6], f[7]));
  }
};

template <typename T, int L, int V>
__global__ void __launch_bounds__(256) k1_backward(const BwdArgs b) {
  using D = DT<T>;
  constexpr int E = Vec16<T>::E;
  constexpr int GPW = 32 / L;
  static_assert(E * V <= 32, "inside mask is one 32-bit word per lane");
  const QuantArgs& a = b.q;
  const int lane = threadIdx.x & 31;
  const int l = lane % L, sub = lane / L;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const uint4* __restrict__ xin = reinterpret_cast<const uint4*>(a.x);
  const uint4* __restrict__ gin = reinterpret_cast<const uint4*>(b.gout);
  uint4* __restrict__ gxo = reinterpret_cast<uint4*>(b.gx);
  const bool has_zp = a.has_zp != 0;

  const int64_t gstep = n_warps * GPW;
  const int64_t g_end = ((a.n_groups + GPW - 1) / GPW) * GPW;   // warp-uniform trip count
  for (int64_t g = warp * GPW + sub; g - sub < g_end; g += gstep) {
    const bool valid = g < a.n_groups;
    uint4 xv[V], gv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      xv[k] = valid ? ldg_stream(xin + g * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);
      gv[k] = valid ? ldg_stream(gin + g * (L * V) + k * L + l) : make_uint4(0, 0, 0, 0);
    }
    // ---- group min / max (NaN-propagating, like the forward)
    float mn, mx;
    vec_minmax<T>(xv[0], mn, mx);
#pragma unroll
    for (int k = 1; k < V; ++k) {
      float m0, m1;
      vec_minmax<T>(xv[k], m0, m1);
      mn = min_nan(mn, m0);
      mx = max_nan(mx, m1);
    }
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
      mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    const GroupQ<T> q = group_params<T>(mn, mx, a, valid ? g : 0);

    // ---- pass 1: STE mask, d/dscale partial sum, multiplicity of the extrema
    float gs = 0.f;
    int nmax = 0, nmin = 0;
    uint32_t inside_bits = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float xf[E], gf[E];
      Vec16<T>::unpack(xv[k], xf);
      Vec16<T>::unpack(gv[k], gf);
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const float v = D::rd(__fdiv_rn(xf[i], q.scale));
        const float r = rintf(v);
        const float t = has_zp ? D::rd(r + q.zp) : r;
        const bool inside = (t >= a.qmin) && (t <= a.qmax);
        const float c = fminf(fmaxf(t, a.qmin), a.qmax);
        const float cz = has_zp ? D::rd(c - q.zp) : c;
        gs = fmaf(gf[i], cz - (inside ? v : 0.f), gs);
        inside_bits |= (inside ? 1u : 0u) << (k * E + i);
        nmax += (xf[i] == mx) ? 1 : 0;
        nmin += (xf[i] == mn) ? 1 : 0;
      }
    }
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, d);
      nmax += __shfl_xor_sync(0xffffffffu, nmax, d);
      nmin += __shfl_xor_sync(0xffffffffu, nmin, d);
    }

    // ---- per-group chain: scale -> (clipped) extrema -> LWC multipliers / extrema
    float su = 1.f, sl = 1.f, xmax_c = mx, xmin_c = mn;
    const bool lwc = a.lwc_up != nullptr;
    if (lwc) {
      su = D::to_f(reinterpret_cast<const T*>(a.lwc_up)[valid ? g : 0]);
      sl = D::to_f(reinterpret_cast<const T*>(a.lwc_low)[valid ? g : 0]);
      xmax_c = D::rd(su * mx);
      xmin_c = D::rd(sl * mn);
    }
    float g_xmax_c, g_xmin_c;
    if (a.symmetric) {
      const float a_hi = fabsf(xmax_c), a_lo = fabsf(xmin_c);
      const float s_raw = D::rd(scale_div(fmaxf(a_hi, a_lo), a));
      const float live = (s_raw >= 1e-5f && s_raw <= 1e4f) ? 1.f : 0.f;
      const float g_abs = scale_div(gs * live, a);
      const float w_hi = a_hi > a_lo ? 1.f : (a_hi == a_lo ? 0.5f : 0.f);
      const float sg_hi = xmax_c > 0.f ? 1.f : (xmax_c < 0.f ? -1.f : 0.f);
      const float sg_lo = xmin_c > 0.f ? 1.f : (xmin_c < 0.f ? -1.f : 0.f);
      g_xmax_c = g_abs * w_hi * sg_hi;
      g_xmin_c = g_abs * (1.f - w_hi) * sg_lo;
    } else {
      const float s_raw = D::rd(scale_div(D::rd(xmax_c - xmin_c), a));
      const float live = (s_raw >= 1e-5f && s_raw <= 1e4f) ? 1.f : 0.f;
      g_xmax_c = scale_div(gs * live, a);
      g_xmin_c = -g_xmax_c;
    }
    if (l == 0 && valid && lwc) {
      if (b.gup) reinterpret_cast<T*>(b.gup)[g] = D::from_f(g_xmax_c * mx);
      if (b.glow) reinterpret_cast<T*>(b.glow)[g] = D::from_f(g_xmin_c * mn);
    }

    // ---- pass 2: grad_x = gout * inside + shares of the extrema
    if (gxo != nullptr) {
      const float pmax = (g_xmax_c * su) / static_cast<float>(nmax);
      const float pmin = (g_xmin_c * sl) / static_cast<float>(nmin);
      // a NaN anywhere in the group makes amin / amax and with them the scale NaN: autograd through the reference op
      // sequence (0 / NaN in the division's backward, a 0 / 0 share of the extrema) then yields NaN for every element
      const bool nan_group = (mn != mn) || (mx != mx);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float xf[E], gf[E], o[E];
        Vec16<T>::unpack(xv[k], xf);
        Vec16<T>::unpack(gv[k], gf);
#pragma unroll
        for (int i = 0; i < E; ++i) {
          float r = ((inside_bits >> (k * E + i)) & 1u) ? gf[i] : 0.f;
          if (xf[i] == mx) r += pmax;
          if (xf[i] == mn) r += pmin;
          o[i] = nan_group ? __int_as_float(0x7fc00000) : r;
        }
        if (valid) stg_stream(gxo + g * (L * V) + k * L + l, Vec16<T>::pack(o));
      }
    }
  }
}

template <typename T, int L, int V>
static int launch_bwd_lv(const BwdArgs& b, cudaStream_t st) {
  constexpr int GPW = 32 / L;
  const int64_t warps_needed = (b.q.n_groups + GPW - 1) / GPW;
  int64_t blocks = (warps_needed + 7) / 8;
  // persistent grid = what the SMs hold at once (80 registers: 3 CTAs per SM; a fixed 6 per SM ran as two waves)
  static const int64_t cap = []() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1_backward<T, L, V>, 256, 0) != cudaSuccess || occ < 1) occ = 2;
    return static_cast<int64_t>(occ) * sm_count();
  }();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k1_backward<T, L, V><<<static_cast<unsigned>(blocks), 256, 0, st>>>(b);
  return check_cuda(cudaGetLastError(), "k1_backward launch");
}

template <typename T>
static int launch_bwd(const BwdArgs& b, cudaStream_t st) {
  constexpr int E = Vec16<T>::E;
  const int vecs = b.q.group / E;
  if (E == 8) {
    switch (vecs) {
      case 4:  return launch_bwd_lv<T, 4, 1>(b, st);    // g = 32
      case 8:  return launch_bwd_lv<T, 8, 1>(b, st);    // g = 64
      case 16: return launch_bwd_lv<T, 8, 2>(b, st);    // g = 128
      case 32: return launch_bwd_lv<T, 8, 4>(b, st);    // g = 256
      default: break;
    }
  } else {
    switch (vecs) {
      case 8:  return launch_bwd_lv<T, 8, 1>(b, st);    // g = 32
      case 16: return launch_bwd_lv<T, 8, 2>(b, st);    // g = 64
      case 32: return launch_bwd_lv<T, 8, 4>(b, st);    // g = 128
      case 64: return launch_bwd_lv<T, 16, 4>(b, st);   // g = 256
      default: break;
    }
  }
  set_error("quantize_backward: group of %d elements is not covered by the fused kernel (32, 64, 128, 256)", b.q.group);
  return Q100_EUNSUPPORTED;
}

}  // namespace q100

using namespace q100;

extern "C" int q100_quantize_backward(const void* x, const void* grad_out, int dtype, int64_t rows, int64_t cols,
                                      int group, int n_bits, int symmetric, int disable_zp,
                                      const void* lwc_up, const void* lwc_low,
                                      void* grad_x, void* grad_up, void* grad_low, int flags, void* stream) {
  Q100_REQUIRE(dtype >= 0 && dtype <= 2, Q100_EINVAL, "bad dtype %d", dtype);
  Q100_REQUIRE(rows >= 0 && cols >= 0, Q100_EINVAL, "negative shape");
  Q100_REQUIRE(n_bits >= 2 && n_bits <= 8, Q100_EUNSUPPORTED, "n_bits=%d: 2..8", n_bits);
  Q100_REQUIRE((lwc_up == nullptr) == (lwc_low == nullptr), Q100_EINVAL, "lwc_up and lwc_low must be given together");
  if (rows == 0 || cols == 0) return Q100_OK;
  Q100_REQUIRE(x != nullptr && grad_out != nullptr, Q100_EINVAL, "x / grad_out is NULL");
  Q100_REQUIRE(group > 0 && cols % group == 0, Q100_EUNSUPPORTED,
               "quantize_backward: cols=%lld, group=%d (group-wise only)", static_cast<long long>(cols), group);
  Q100_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_out) |
                 reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0, Q100_EINVAL,
               "x / grad_out / grad_x must be 16-byte aligned");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  BwdArgs b{};
  QuantArgs& a = b.q;
  a.x = x; a.lwc_up = lwc_up; a.lwc_low = lwc_low;
  a.rows = rows; a.cols = cols; a.group = group;
  a.groups_per_row = static_cast<int>(cols / group);
  a.n_groups = rows * a.groups_per_row;
  Q100_REQUIRE(a.n_groups < (1ll << 31), Q100_EUNSUPPORTED, "too many groups");
  a.n_bits = n_bits; a.symmetric = symmetric ? 1 : 0; a.has_zp = disable_zp ? 0 : 1;
  if (disable_zp) { a.qmin = -static_cast<float>(1 << (n_bits - 1)); a.qmax = static_cast<float>((1 << (n_bits - 1)) - 1); }
  else { a.qmin = 0.0f; a.qmax = static_cast<float>((1 << n_bits) - 1); }
  a.denom = symmetric ? static_cast<float>((1 << (n_bits - 1)) - 1) : static_cast<float>((1 << n_bits) - 1);
  a.rcp_denom = 1.0f / a.denom;
  a.cuda_scale = (flags & Q100_FLAG_CUDA_SCALE) ? 1 : 0;
  b.gout = grad_out; b.gx = grad_x; b.gup = grad_up; b.glow = grad_low;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case Q100_F16:  return launch_bwd<__half>(b, st);
    case Q100_BF16: return launch_bwd<__nv_bfloat16>(b, st);
    default:        return launch_bwd<float>(b, st);
  }
}
