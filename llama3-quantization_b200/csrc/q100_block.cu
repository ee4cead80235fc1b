// q100_block.cu -- the small memory-bound steps that sit between the quantised linears of a decoder block
// at decode shapes (QuantLlamaDecoderLayer.forward, models/int_llama_layer.py:213-262 of the reference):
//
//   K4 add_rmsnorm   h = x + residual;  n = OmniLlamaRMSNorm(h)            quant/omni_norm.py:52-63, int_llama_layer.py:246-251
//   K5 rope_kv       rotary embedding of q and k for the new token(s), K/V written into a static cache
//                                                                          int_llama_layer.py:114-133
//   K6 silu_mul      act_fn(gate) * up on the fused gate|up GEMM output    int_llama_layer.py:45
//
// In eager torch these are ~35 launches per block and token (the reference's decode loop is launch-bound on
// them); here they are four, so that a decode step of a block is 4 GEMMs + 4 small kernels + the attention
// core.  Each kernel reproduces the reference's op sequence and per-op rounding (fp32 where torch promotes,
// one rounding to the 16-bit storage type where torch stores); only the order of the fp32 sum inside the
// variance differs from torch's reduction.  All are HBM/L2-latency bound and tiny: one CTA per row / head.
#include <cstdlib>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

// The small kernels of a decoder block are launched with programmatic stream serialization (Q100_BLOCK_PDL=0 turns it
// off): each one lets its successor start at once (griddepcontrol.launch_dependents) and waits for its predecessors
// (griddepcontrol.wait) before it touches memory -- a decode GEMM that follows K4 / K5 then fetches its first weight
// blocks while the small kernel runs.  Without a PDL-aware predecessor both instructions are no-ops.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  static const bool on = []() { const char* e = getenv("Q100_BLOCK_PDL"); return !(e && atoi(e) == 0); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- K4: (residual add +) RMSNorm
// rows x cols, one CTA of 256 threads per row, the row lives in registers (cols <= 8192).
template <typename T, int VMAX>
__global__ void __launch_bounds__(256) k4_add_rmsnorm(const T* __restrict__ x, const T* __restrict__ res,
                                                      const T* __restrict__ w, float eps, T* __restrict__ sum_out,
                                                      T* __restrict__ norm_out, int cols) {
  using D = DT<T>;
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int nvec = cols >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * cols);
  const uint4* rr = res ? reinterpret_cast<const uint4*>(res + row * cols) : nullptr;
  uint4 h[VMAX];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < VMAX; ++k) {
    const int v = threadIdx.x + k * 256;
    if (v < nvec) {
      uint4 a = xr[v];
      if (rr) {
        const uint4 b = rr[v];
        uint32_t* aw = reinterpret_cast<uint32_t*>(&a);
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {                       // residual + hidden_states, rounded to T
          const float2 fa = D::to_f2(aw[j]), fb = D::to_f2(bw[j]);
          aw[j] = D::pack2(__fadd_rn(fb.x, fa.x), __fadd_rn(fb.y, fa.y));
        }
        if (sum_out) reinterpret_cast<uint4*>(sum_out + row * cols)[v] = a;
      }
      h[k] = a;
      const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = D::to_f2(aw[j]);
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = rsqrtf(__fadd_rn(__fdiv_rn(tot, static_cast<float>(cols)), eps));
  const uint4* wr = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int k = 0; k < VMAX; ++k) {
    const int v = threadIdx.x + k * 256;
    if (v < nvec) {
      const uint4 wv = wr[v];
      const uint32_t* hw = reinterpret_cast<const uint32_t*>(&h[k]);
      const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {                         // weight * (h * rsqrt(var + eps)), both products in fp32
        const float2 f = D::to_f2(hw[j]), g = D::to_f2(ww[j]);
        ow[j] = D::pack2(__fmul_rn(g.x, __fmul_rn(f.x, inv)), __fmul_rn(g.y, __fmul_rn(f.y, inv)));
      }
      reinterpret_cast<uint4*>(norm_out + row * cols)[v] = o;
    }
  }
}

// ---------------------------------------------------------------- K5: rotary + KV-cache append (decode)
// grid (heads_total, B), D/2 threads: thread i owns the pair (i, i + D/2) of one head of one token.
// q: fp32 math on the fp32-cast query (int_llama_layer.py:117, 125), one rounding to T at the end;
// k: every op rounded to T (fp16/bf16 tensor math); v: copied.
template <typename T>
__global__ void k5_rope_kv(const T* __restrict__ qkv, const T* __restrict__ cosb, const T* __restrict__ sinb,
                           const int64_t* __restrict__ pos_p, T* __restrict__ q_out, T* __restrict__ k_cache,
                           T* __restrict__ v_cache, int H, int KVH, int D, int64_t max_len, int64_t rope_len,
                           uint32_t* __restrict__ err_flag) {
  using Dt = DT<T>;
  pdl_launch_dependents();
  pdl_wait();
  const int head = blockIdx.x, b = blockIdx.y, i = threadIdx.x, half = D >> 1;
  const int64_t pos = *pos_p;
  if (pos < 0 || pos >= max_len || pos >= rope_len) {      // past the cache / the rotary table: touch nothing
    if (err_flag != nullptr && head == 0 && b == 0 && i == 0) *err_flag = 1u;
    return;
  }
  const T* src = qkv + static_cast<int64_t>(b) * (H + 2 * KVH) * D + static_cast<int64_t>(head) * D;
  const float x1 = Dt::to_f(src[i]), x2 = Dt::to_f(src[i + half]);
  if (head >= H + KVH) {                                   // value head: plain copy into the cache
    T* dst = v_cache + ((static_cast<int64_t>(b) * KVH + (head - H - KVH)) * max_len + pos) * D;
    dst[i] = src[i];
    dst[i + half] = src[i + half];
    return;
  }
  const float c1 = Dt::to_f(cosb[pos * D + i]), c2 = Dt::to_f(cosb[pos * D + i + half]);
  const float s1 = Dt::to_f(sinb[pos * D + i]), s2 = Dt::to_f(sinb[pos * D + i + half]);
  if (head < H) {
    T* dst = q_out + (static_cast<int64_t>(b) * H + head) * D;
    dst[i] = Dt::from_f(__fadd_rn(__fmul_rn(x1, c1), __fmul_rn(-x2, s1)));
    dst[i + half] = Dt::from_f(__fadd_rn(__fmul_rn(x2, c2), __fmul_rn(x1, s2)));
  } else {
    T* dst = k_cache + ((static_cast<int64_t>(b) * KVH + (head - H)) * max_len + pos) * D;
    dst[i] = Dt::from_f(__fadd_rn(Dt::rd(__fmul_rn(x1, c1)), Dt::rd(__fmul_rn(-x2, s1))));
    dst[i + half] = Dt::from_f(__fadd_rn(Dt::rd(__fmul_rn(x2, c2)), Dt::rd(__fmul_rn(x1, s2))));
  }
}

// ---------------------------------------------------------------- K5p: rotary + head-major q / k / v (prefill)
// One CTA per token; thread = (head, 8 consecutive pair indices): 16-byte loads of the two halves of a head from the
// fused q|k|v row, 16-byte stores into q_out [B, H, T, D], k_out / v_out [B, KVH, T, D] (what the attention core and
// the tuple cache read).  Same arithmetic as K5: q in fp32 with one rounding, k with every op rounded to T, v copied.
template <typename T>
__global__ void __launch_bounds__(256) k5p_rope_qkv(const T* __restrict__ qkv, const T* __restrict__ cosb,
                                                    const T* __restrict__ sinb, const int64_t* __restrict__ pos_ids,
                                                    int64_t pos_bstride, T* __restrict__ q_out, T* __restrict__ k_out,
                                                    T* __restrict__ v_out, int Tq, int H, int KVH, int D,
                                                    int64_t rope_len, uint32_t* __restrict__ err_flag) {
  using Dt = DT<T>;
  const int t = blockIdx.x, b = blockIdx.y;
  const int half = D >> 1, cph = half >> 3;                 // 16-byte chunks per half head
  const int heads = H + 2 * KVH;
  const int64_t pos = pos_ids[b * pos_bstride + t];
  if (pos < 0 || pos >= rope_len) {
    if (err_flag != nullptr && threadIdx.x == 0) *err_flag = 1u;
    return;
  }
  const T* row = qkv + (static_cast<int64_t>(b) * Tq + t) * heads * D;
  const T* cr = cosb + pos * D;
  const T* sr = sinb + pos * D;
  for (int w = threadIdx.x; w < heads * cph; w += blockDim.x) {
    const int head = w / cph, c = w - head * cph;
    const T* src = row + static_cast<int64_t>(head) * D + c * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(src);
    const uint4 bb = *reinterpret_cast<const uint4*>(src + half);
    T* dst;
    if (head < H) dst = q_out + ((static_cast<int64_t>(b) * H + head) * Tq + t) * D + c * 8;
    else if (head < H + KVH) dst = k_out + ((static_cast<int64_t>(b) * KVH + (head - H)) * Tq + t) * D + c * 8;
    else dst = v_out + ((static_cast<int64_t>(b) * KVH + (head - H - KVH)) * Tq + t) * D + c * 8;
    if (head >= H + KVH) {
      *reinterpret_cast<uint4*>(dst) = a;
      *reinterpret_cast<uint4*>(dst + half) = bb;
      continue;
    }
    const uint4 c1v = *reinterpret_cast<const uint4*>(cr + c * 8), c2v = *reinterpret_cast<const uint4*>(cr + half + c * 8);
    const uint4 s1v = *reinterpret_cast<const uint4*>(sr + c * 8), s2v = *reinterpret_cast<const uint4*>(sr + half + c * 8);
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bb);
    const uint32_t* c1w = reinterpret_cast<const uint32_t*>(&c1v);
    const uint32_t* c2w = reinterpret_cast<const uint32_t*>(&c2v);
    const uint32_t* s1w = reinterpret_cast<const uint32_t*>(&s1v);
    const uint32_t* s2w = reinterpret_cast<const uint32_t*>(&s2v);
    uint4 o1, o2;
    uint32_t* o1w = reinterpret_cast<uint32_t*>(&o1);
    uint32_t* o2w = reinterpret_cast<uint32_t*>(&o2);
    const bool is_q = head < H;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x1 = Dt::to_f2(aw[j]), x2 = Dt::to_f2(bw[j]);
      const float2 c1 = Dt::to_f2(c1w[j]), c2 = Dt::to_f2(c2w[j]), s1 = Dt::to_f2(s1w[j]), s2 = Dt::to_f2(s2w[j]);
      float lo_x, lo_y, hi_x, hi_y;
      if (is_q) {
        lo_x = __fadd_rn(__fmul_rn(x1.x, c1.x), __fmul_rn(-x2.x, s1.x));
        lo_y = __fadd_rn(__fmul_rn(x1.y, c1.y), __fmul_rn(-x2.y, s1.y));
        hi_x = __fadd_rn(__fmul_rn(x2.x, c2.x), __fmul_rn(x1.x, s2.x));
        hi_y = __fadd_rn(__fmul_rn(x2.y, c2.y), __fmul_rn(x1.y, s2.y));
      } else {
        lo_x = __fadd_rn(Dt::rd(__fmul_rn(x1.x, c1.x)), Dt::rd(__fmul_rn(-x2.x, s1.x)));
        lo_y = __fadd_rn(Dt::rd(__fmul_rn(x1.y, c1.y)), Dt::rd(__fmul_rn(-x2.y, s1.y)));
        hi_x = __fadd_rn(Dt::rd(__fmul_rn(x2.x, c2.x)), Dt::rd(__fmul_rn(x1.x, s2.x)));
        hi_y = __fadd_rn(Dt::rd(__fmul_rn(x2.y, c2.y)), Dt::rd(__fmul_rn(x1.y, s2.y)));
      }
      o1w[j] = Dt::pack2(lo_x, lo_y);
      o2w[j] = Dt::pack2(hi_x, hi_y);
    }
    *reinterpret_cast<uint4*>(dst) = o1;
    *reinterpret_cast<uint4*>(dst + half) = o2;
  }
}

// ---------------------------------------------------------------- K6: silu(gate) * up
// gate_up [M, 2I] (gate columns first), out [M, I]; 8 elements per thread.
template <typename T>
__global__ void __launch_bounds__(256) k6_silu_mul(const T* __restrict__ gu, T* __restrict__ out, int64_t M, int I) {
  using D = DT<T>;
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = I >> 3;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * vec_per_row) return;
  const int64_t m = idx / vec_per_row;
  const int v = static_cast<int>(idx - m * vec_per_row);
  const uint4 g = reinterpret_cast<const uint4*>(gu + m * 2 * I)[v];
  const uint4 u = reinterpret_cast<const uint4*>(gu + m * 2 * I + I)[v];
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
  const uint32_t* uw = reinterpret_cast<const uint32_t*>(&u);
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = D::to_f2(gw[j]), b = D::to_f2(uw[j]);
    const float sx = D::rd(__fdiv_rn(a.x, __fadd_rn(1.0f, expf(-a.x))));     // F.silu: fp32 math, stored as T
    const float sy = D::rd(__fdiv_rn(a.y, __fadd_rn(1.0f, expf(-a.y))));
    ow[j] = D::pack2(__fmul_rn(sx, b.x), __fmul_rn(sy, b.y));
  }
  reinterpret_cast<uint4*>(out + m * I)[v] = o;
}

template <typename T>
static int launch_add_rmsnorm(const void* x, const void* res, const void* w, float eps, void* sum_out,
                              void* norm_out, int64_t rows, int cols, cudaStream_t st) {
  const int nvec = cols / 8;
  const T* xp = reinterpret_cast<const T*>(x);
  const T* rp = reinterpret_cast<const T*>(res);
  const T* wp = reinterpret_cast<const T*>(w);
  T* sp = reinterpret_cast<T*>(sum_out);
  T* np_ = reinterpret_cast<T*>(norm_out);
  const unsigned grid = static_cast<unsigned>(rows);
  cudaError_t e;
  if (nvec <= 256) e = launch_pdl(k4_add_rmsnorm<T, 1>, dim3(grid), dim3(256), st, xp, rp, wp, eps, sp, np_, cols);
  else if (nvec <= 512) e = launch_pdl(k4_add_rmsnorm<T, 2>, dim3(grid), dim3(256), st, xp, rp, wp, eps, sp, np_, cols);
  else e = launch_pdl(k4_add_rmsnorm<T, 4>, dim3(grid), dim3(256), st, xp, rp, wp, eps, sp, np_, cols);
  return check_cuda(e, "k4_add_rmsnorm launch");
}

}  // namespace q100

using namespace q100;

extern "C" int q100_add_rmsnorm(const void* x, const void* residual, const void* weight, float eps,
                                void* sum_out, void* norm_out, int64_t rows, int64_t cols, int dtype,
                                void* stream) {
  Q100_REQUIRE(x && weight && norm_out, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && cols <= 8192, Q100_EUNSUPPORTED,
               "cols=%lld: multiple of 8, at most 8192", (long long)cols);
  Q100_REQUIRE(rows < (1ll << 31), Q100_EINVAL, "too many rows");
  if (rows == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == Q100_F16)
    return launch_add_rmsnorm<__half>(x, residual, weight, eps, sum_out, norm_out, rows, static_cast<int>(cols), st);
  return launch_add_rmsnorm<__nv_bfloat16>(x, residual, weight, eps, sum_out, norm_out, rows, static_cast<int>(cols), st);
}

extern "C" int q100_rope_kv_decode(const void* qkv, const void* cos_cache, const void* sin_cache,
                                   const void* position, void* q_out, void* k_cache, void* v_cache,
                                   int batch, int heads, int kv_heads, int head_dim, int64_t max_len,
                                   int64_t rope_len, void* err_flag, int dtype, void* stream) {
  Q100_REQUIRE(qkv && cos_cache && sin_cache && position && q_out && k_cache && v_cache, Q100_EINVAL,
               "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(batch >= 0 && heads > 0 && kv_heads > 0 && head_dim >= 2 && head_dim % 2 == 0 && head_dim <= 2048 &&
               max_len > 0 && rope_len > 0, Q100_EINVAL, "bad shape");
  if (batch == 0) return Q100_OK;
  uint32_t* ef = reinterpret_cast<uint32_t*>(err_flag);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(heads + 2 * kv_heads, batch);
  const int64_t* pos = reinterpret_cast<const int64_t*>(position);
  cudaError_t e;
  if (dtype == Q100_F16)
    e = launch_pdl(k5_rope_kv<__half>, grid, dim3(head_dim / 2), st,
        reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(cos_cache),
        reinterpret_cast<const __half*>(sin_cache), pos, reinterpret_cast<__half*>(q_out),
        reinterpret_cast<__half*>(k_cache), reinterpret_cast<__half*>(v_cache), heads, kv_heads, head_dim, max_len,
        rope_len, ef);
  else
    e = launch_pdl(k5_rope_kv<__nv_bfloat16>, grid, dim3(head_dim / 2), st,
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(cos_cache),
        reinterpret_cast<const __nv_bfloat16*>(sin_cache), pos, reinterpret_cast<__nv_bfloat16*>(q_out),
        reinterpret_cast<__nv_bfloat16*>(k_cache), reinterpret_cast<__nv_bfloat16*>(v_cache), heads, kv_heads,
        head_dim, max_len, rope_len, ef);
  return check_cuda(e, "k5_rope_kv launch");
}

extern "C" int q100_rope_qkv_prefill(const void* qkv, const void* cos_cache, const void* sin_cache,
                                     const void* position_ids, int64_t pos_batch_stride, void* q_out, void* k_out,
                                     void* v_out, int batch, int64_t q_len, int heads, int kv_heads, int head_dim,
                                     int64_t rope_len, void* err_flag, int dtype, void* stream) {
  Q100_REQUIRE(qkv && cos_cache && sin_cache && position_ids && q_out && k_out && v_out, Q100_EINVAL,
               "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(batch >= 0 && batch < 65536 && q_len >= 0 && q_len < (1ll << 31) && heads > 0 && kv_heads > 0 &&
               rope_len > 0 && pos_batch_stride >= 0, Q100_EINVAL, "bad shape");
  Q100_REQUIRE(head_dim >= 16 && head_dim % 16 == 0 && head_dim <= 1024, Q100_EUNSUPPORTED,
               "head_dim=%d: a multiple of 16 (16-byte accesses of both halves of a head)", head_dim);
  if (batch == 0 || q_len == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>(q_len), static_cast<unsigned>(batch));
  const int64_t* pos = reinterpret_cast<const int64_t*>(position_ids);
  uint32_t* ef = reinterpret_cast<uint32_t*>(err_flag);
  if (dtype == Q100_F16)
    k5p_rope_qkv<__half><<<grid, 256, 0, st>>>(
        reinterpret_cast<const __half*>(qkv), reinterpret_cast<const __half*>(cos_cache),
        reinterpret_cast<const __half*>(sin_cache), pos, pos_batch_stride, reinterpret_cast<__half*>(q_out),
        reinterpret_cast<__half*>(k_out), reinterpret_cast<__half*>(v_out), static_cast<int>(q_len), heads, kv_heads,
        head_dim, rope_len, ef);
  else
    k5p_rope_qkv<__nv_bfloat16><<<grid, 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(cos_cache),
        reinterpret_cast<const __nv_bfloat16*>(sin_cache), pos, pos_batch_stride,
        reinterpret_cast<__nv_bfloat16*>(q_out), reinterpret_cast<__nv_bfloat16*>(k_out),
        reinterpret_cast<__nv_bfloat16*>(v_out), static_cast<int>(q_len), heads, kv_heads, head_dim, rope_len, ef);
  return check_cuda(cudaGetLastError(), "k5p_rope_qkv launch");
}

extern "C" int q100_silu_mul(const void* gate_up, void* out, int64_t rows, int64_t inter, int dtype, void* stream) {
  Q100_REQUIRE(gate_up && out, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(rows >= 0 && inter > 0 && inter % 8 == 0 && inter < (1ll << 31), Q100_EUNSUPPORTED,
               "inter=%lld must be a positive multiple of 8", (long long)inter);
  if (rows == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = rows * (inter / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dtype == Q100_F16)
    k6_silu_mul<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(gate_up), reinterpret_cast<__half*>(out),
                                              rows, static_cast<int>(inter));
  else
    k6_silu_mul<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(gate_up),
                                                     reinterpret_cast<__nv_bfloat16*>(out), rows,
                                                     static_cast<int>(inter));
  return check_cuda(cudaGetLastError(), "k6_silu_mul launch");
}

// ---------------------------------------------------------------- K10: argmax over the vocabulary (greedy decode step)
// next_token = logits.argmax(-1) of the reference's generation loop (llama.py:409-465 measures forward + token
// choice per step).  torch's reduction runs one CTA per row: 42 us for a [1, 128256] fp16 row inside the decode graph.
// Here AM_CTAS CTAs per row each scan a slice with 16-byte loads, leave (value, index) in global memory, and the last
// one to arrive (a counter per row, left at zero) reduces the partials.  Ties: the lowest index; NaN counts as the
// maximum (torch's rule), the first NaN wins.
namespace q100 {
constexpr int AM_CTAS = 64;

__device__ __forceinline__ bool am_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

template <typename T>
__global__ void __launch_bounds__(256) k10_argmax(const T* __restrict__ x, int64_t cols, float2* __restrict__ part,
                                                  uint32_t* __restrict__ counters, int64_t* __restrict__ out) {
  using D = DT<T>;
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* xr = x + static_cast<int64_t>(row) * cols;
  const int64_t nvec = cols >> 3;
  float bv = __int_as_float(0xff800000);
  int bi = 0x7fffffff;
  for (int64_t v = static_cast<int64_t>(c) * 256 + tid; v < nvec; v += static_cast<int64_t>(AM_CTAS) * 256) {
    const uint4 q = __ldcs(reinterpret_cast<const uint4*>(xr) + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = D::to_f2(w[j]);
      const int i0 = static_cast<int>(v * 8) + 2 * j;
      if (am_better(f.x, i0, bv, bi)) { bv = f.x; bi = i0; }
      if (am_better(f.y, i0 + 1, bv, bi)) { bv = f.y; bi = i0 + 1; }
    }
  }
  if (c == 0) {                                        // the tail when cols % 8 != 0
    for (int64_t i = (nvec << 3) + tid; i < cols; i += 256) {
      const float f = D::to_f(xr[i]);
      if (am_better(f, static_cast<int>(i), bv, bi)) { bv = f; bi = static_cast<int>(i); }
    }
  }
  auto warp_red = [&]() {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, d);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
      if (am_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  };
  warp_red();
  if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = lane < 8 ? s_v[lane] : __int_as_float(0xff800000);
    bi = lane < 8 ? s_i[lane] : 0x7fffffff;
    warp_red();
    if (lane == 0) {
      part[row * AM_CTAS + c] = make_float2(bv, __int_as_float(bi));
      __threadfence();
      const uint32_t old = atomicAdd(counters + row, 1u);
      const int last = old == AM_CTAS - 1;
      if (last) counters[row] = 0u;
      __threadfence();
      s_last = last;
    }
    __syncwarp();
    if (s_last) {
      bv = __int_as_float(0xff800000);
      bi = 0x7fffffff;
      for (int k = lane; k < AM_CTAS; k += 32) {
        const float2 p = __ldcg(part + row * AM_CTAS + k);
        const int pi = __float_as_int(p.y);
        if (am_better(p.x, pi, bv, bi)) { bv = p.x; bi = pi; }
      }
      warp_red();
      if (lane == 0) out[row] = bi;
    }
  }
}
}  // namespace q100

extern "C" size_t q100_argmax_workspace_bytes(int64_t rows) {
  if (rows <= 0) return 0;
  return ((static_cast<size_t>(rows) * 4 + 255) & ~static_cast<size_t>(255)) + static_cast<size_t>(rows) * AM_CTAS * 8;
}

extern "C" int q100_argmax_rows(const void* x, int64_t rows, int64_t cols, int dtype, void* out_index,
                                void* workspace, size_t workspace_bytes, void* stream) {
  Q100_REQUIRE(x && out_index && workspace, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(rows >= 0 && rows < 65536 && cols > 0 && cols < (1ll << 31), Q100_EINVAL, "bad shape rows=%lld cols=%lld",
               (long long)rows, (long long)cols);
  Q100_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (cols % 8 == 0 || rows == 1), Q100_EUNSUPPORTED,
               "rows must start on 16-byte boundaries (cols %% 8 == 0)");
  if (rows == 0) return Q100_OK;
  Q100_REQUIRE(workspace_bytes >= q100_argmax_workspace_bytes(rows), Q100_EWORKSPACE, "workspace too small: %zu bytes",
               workspace_bytes);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint32_t* counters = reinterpret_cast<uint32_t*>(workspace);
  float2* part = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(workspace) +
                                           ((static_cast<size_t>(rows) * 4 + 255) & ~static_cast<size_t>(255)));
  const dim3 grid(AM_CTAS, static_cast<unsigned>(rows));
  cudaError_t e;
  if (dtype == Q100_F16)
    e = launch_pdl(k10_argmax<__half>, grid, dim3(256), st, reinterpret_cast<const __half*>(x), cols, part, counters,
                   reinterpret_cast<int64_t*>(out_index));
  else
    e = launch_pdl(k10_argmax<__nv_bfloat16>, grid, dim3(256), st, reinterpret_cast<const __nv_bfloat16*>(x), cols, part,
                   counters, reinterpret_cast<int64_t*>(out_index));
  return check_cuda(e, "k10_argmax launch");
}
