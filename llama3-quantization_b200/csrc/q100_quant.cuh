// q100_quant.cuh -- pieces shared by the quantize kernels (q100_quant.cu) and their backward (q100_quant_bwd.cu):
// argument block, per-group parameter computation (quantizer.py:143-159) and NaN-aware vector min / max.
#pragma once
#include "q100_common.cuh"

namespace q100 {

enum { MATH_PACKED = 0, MATH_SCALAR = 1, MATH_SCALAR_IEEE = 2 };

struct QuantArgs {
  const void* x;
  void* deq;
  uint32_t* packed;
  void* scales;
  void* zeros;
  void* meta;
  const void* lwc_up;
  const void* lwc_low;
  const void* ext_minmax;   // K2 only: [rows, 2] (min, max) supplied by the caller instead of the row reduction
  void* minmax_out;         // K2 only: write the row's (min, max) here and do nothing else
  int64_t rows, cols, n_groups;
  int group;            // elements per group
  int groups_per_row;
  int n_bits;
  int symmetric;
  int has_zp;
  float qmin, qmax;     // reference code range (quantizer.py:46-51)
  float denom;          // 2^n - 1 (asymmetric) or 2^(n-1) - 1 (symmetric)
  float rcp_denom;      // RN(1 / denom), for k1_fast
  int cuda_scale;       // Q100_FLAG_CUDA_SCALE: scale = range * RN(1 / denom) (torch's CUDA kernels) instead of range / denom
  int k_tiles;          // packed layout v2: cols / 64
  int tile_words;       // uint32 words per [128 rows x 64 k] tile
};

// ---- packed layout v2 (tile-major, DESIGN.md "Packed layout"): tiles of 128 rows x 64 k stored
// contiguously in [n_tile][k_tile] order; W3 = a 2-bit plane (128 x 4 words) then a 1-bit plane.
__host__ __device__ __forceinline__ int tile_words_of(int n_bits) {
  return n_bits == 3 ? 128 * 6 : 128 * (2 * n_bits);
}

// ------------------------------------------------------------------ per-group parameters
template <typename T>
struct GroupQ {
  float scale, zp, rcp;
};

// range / (2^n - 1), quantizer.py:147 / :152.  torch's CPU kernels divide; its CUDA kernels multiply a tensor by the fp32
// reciprocal of a Python scalar divisor (BinaryDivTrueKernel.cu: "inv_b = 1 / b; a * inv_b"), which can differ by one ulp.
__device__ __forceinline__ float scale_div(float num, const QuantArgs& a) {
  return a.cuda_scale ? __fmul_rn(num, a.rcp_denom) : __fdiv_rn(num, a.denom);
}

// quantizer.py:143-159.  xmin/xmax are exact values of T.
template <typename T>
__device__ __forceinline__ GroupQ<T> group_params(float xmin, float xmax, const QuantArgs& a,
                                                  int64_t g) {
  using D = DT<T>;
  if (a.lwc_up != nullptr) {
    const float up = D::to_f(reinterpret_cast<const T*>(a.lwc_up)[g]);
    const float low = D::to_f(reinterpret_cast<const T*>(a.lwc_low)[g]);
    xmax = D::rd(up * xmax);
    xmin = D::rd(low * xmin);
  }
  GroupQ<T> q;
  if (a.symmetric) {
    const float am = max_nan(fabsf(xmax), fabsf(xmin));
    float s = D::rd(scale_div(am, a));
    q.scale = D::rd(clamp_nan(s, 1e-5f, 1e4f));
    q.zp = D::rd(a.denom);
  } else {
    const float rng = D::rd(xmax - xmin);
    float s = D::rd(scale_div(rng, a));
    q.scale = D::rd(clamp_nan(s, 1e-5f, 1e4f));
    q.zp = D::rd(__fdiv_rn(D::rd(-xmin), q.scale));
  }
  q.zp = a.has_zp ? rintf(D::rd(clamp_nan(q.zp, -1e4f, 1e4f))) : 0.0f;
  q.rcp = __frcp_rn(q.scale);
  return q;
}

// x / s rounded like div.rn.f32 as far as a 16-bit result can tell (see file header).
template <int MATH>
__device__ __forceinline__ float div_x_s(float x, float s, float rcp) {
  if (MATH == MATH_SCALAR_IEEE) return __fdiv_rn(x, s);
  float q = x * rcp;
  const float r = fmaf(-q, s, x);
  return fmaf(r, rcp, q);
}

// ------------------------------------------------------------------ min / max of a vector
template <typename T>
__device__ __forceinline__ void vec_minmax(const uint4& v, float& mn, float& mx);

template <>
__device__ __forceinline__ void vec_minmax<float>(const uint4& v, float& mn, float& mx) {
  const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z),
              d = __uint_as_float(v.w);
  mn = fminf(fminf(a, b), fminf(c, d));
  mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
  if (a != a || b != b || c != c || d != d) mn = mx = nan_f();
}
template <>
__device__ __forceinline__ void vec_minmax<__half>(const uint4& v, float& mn, float& mx) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  const __half2 lo = __hmin2_nan(__hmin2_nan(h[0], h[1]), __hmin2_nan(h[2], h[3]));
  const __half2 hi = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
  mn = min_nan(__low2float(lo), __high2float(lo));
  mx = max_nan(__low2float(hi), __high2float(hi));
}
template <>
__device__ __forceinline__ void vec_minmax<__nv_bfloat16>(const uint4& v, float& mn, float& mx) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  const __nv_bfloat162 lo = __hmin2_nan(__hmin2_nan(h[0], h[1]), __hmin2_nan(h[2], h[3]));
  const __nv_bfloat162 hi = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
  mn = min_nan(__low2float(lo), __high2float(lo));
  mx = max_nan(__low2float(hi), __high2float(hi));
}


}  // namespace q100
