// q100_chain.cuh -- argument blocks of the decode-shape GEMM kernel (k3_skinny) and of its chain form (k3_chain):
// ONE persistent launch that walks a list of phases -- the dequant-GEMMs of one or many decoder blocks with the
// RMSNorm / residual / SiLU*up steps folded into their x staging and epilogues, and the single-token attention
// between q|k|v and o_proj -- separated by grid barriers, while the producer warps keep the packed-weight stream
// running across the phase boundaries (the weights are static, so they never wait).
//
// Reference code a chain replaces, per decoder block and token (models/int_llama_layer.py:213-262):
//   input_layernorm -> q/k/v_proj -> rotary + cache append -> softmax(QK^T) V -> o_proj -> + residual ->
//   post_attention_layernorm -> gate/up_proj -> act_fn(gate) * up -> down_proj -> + residual
#pragma once
#include "q100_common.cuh"

namespace q100 {

struct SkinnyArgs {
  const void* x;          // [M, K] T
  const void* packed;     // tile-major packed codes (layout v2)
  const void* meta;       // [K/group, N, 2] T
  const void* bias;       // [N] T or null
  void* y;                // [M, N] T
  uint2* ws_partials;     // workspace: fp32 partial fragments as {value, flag} pairs (all zero between launches)
  int M, N, K, group;
  int kblocks;            // K / 128 = stream-K units per tile
  int n_tiles;
  int total_units;        // n_tiles * kblocks
  int max_slots;
  int has_zp;
  float qmin;
  int post;               // 1: magic-form operand + fp32 group scaling (group % 128 == 0)
  int m_rows;             // rows of x kept in shared memory (= M)
  int chunk;              // k blocks of x staged per round
  int stages;             // ring depth
  int nsub;               // (scale, zero) rows per block: 128 / group for group < 128, else 1
  int pdl;                // launch with programmatic stream serialization
  int seq;                // launch sequence number (timing builds)
  int x_gate_up;          // x is [M, 2K] = gate | up: stage act_fn(gate) * up (K6's arithmetic)
  // ---- MODE 3 (q100_gemm_wna8): per-token activation quantisation as the GEMM's prologue, integer contraction
  const void* x_minmax;   // [M, 2] T: (min, max) of every token row over the FULL K (q100_rows_minmax, all-reduced under TP)
  float a_qmax, a_denom;  // 2^abits - 1 (codes are [0, a_qmax]; scale = range / a_denom)
  int a_cuda_scale;       // Q100_FLAG_CUDA_SCALE for the activation scale
  // ---- chain mode only (k3_chain, q100_chain.cuh): what surrounds the GEMM inside a decoder block
  int x_norm;             // x is the residual stream h [M, K]: stage weight * (h * rstd) (K4's arithmetic, omni_norm.py:52-63)
  const void* norm_w;     // [K] T
  const float* ssq_in;    // [n_part_in][M] partial sums of h^2 written by the phase that produced h
  int n_part_in;
  float eps;
  int x_attn;             // x[m, head * 128 ..] = the attention output of (sequence m, head): merged here from the
  const float* att_part;  // flash-decoding partials [M * H][att_max_parts][AT_PART_STRIDE] the attention phase left,
  const int64_t* att_pos; // for the chunks up to the device-side position (K must equal H * 128)
  int att_max_parts, att_chunk_pos;
  int epi_add;            // y = r(r(acc) + res) (the residual adds of int_llama_layer.py:246, 251) and ...
  const void* res;        // [M, N] T
  float* ssq_out;         // ... [n_tiles * 8][M] partial sums of y^2 for the next RMSNorm
};

enum { PH_GEMM = 0, PH_ATTN = 1, PH_COMBINE = 2 };
constexpr int AT_PART_STRIDE = 132;     // floats per attention partial: {m, l, -, -, o[128]}
constexpr int AT_D = 128;               // head_dim the attention phases are written for (LLaMA3 8B / 70B)
constexpr int CH_MAX_PHASES = 1024;

// single-token attention on a static KV cache (the decode form of int_llama_layer.py:114-172)
struct AttnArgs {
  const void* qkv;        // [B, (H + 2 KVH) * D] T: fused q|k|v projection of the new tokens
  const void* cosb;       // [rope_len, D] T
  const void* sinb;
  const int64_t* pos;     // device scalar: position of the new token
  void* k_cache;          // [B, KVH, max_len, D] T
  void* v_cache;
  float* part;            // [B * H][max_parts][AT_PART_STRIDE]: one flash-decoding partial per position chunk
  void* out;              // [B, H * D] T (what o_proj reads)
  int B, H, KVH;
  int chunk_pos;          // positions per work item (multiple of 64)
  int max_parts;          // ceil(max_len / chunk_pos)
  int pad_;
  int64_t max_len, rope_len;
  float scale;            // 1 / sqrt(D)
  uint32_t* err;          // set to 1 when the position is outside the cache / rotary table
};

struct alignas(16) ChainPhase {
  int type;               // PH_*
  int grid_p;             // GEMM: CTAs that take units of this phase (<= gridDim.x; small GEMMs are not cut too finely)
  SkinnyArgs g;
  AttnArgs at;
};
constexpr int CH_PHASE_BYTES = 512;     // shared-memory slot of a phase (copied in 16-byte pieces)
static_assert(sizeof(ChainPhase) <= CH_PHASE_BYTES && sizeof(ChainPhase) % 16 == 0, "ChainPhase layout");

struct ChainArgs {
  const ChainPhase* phases;   // device memory
  int n_phases;
  uint32_t* bars;             // [n_phases + 1] arrival counters, zero between launches (the last CTA out resets them)
  uint32_t* err;              // set to 1 if a grid barrier timed out (results are then invalid)
  unsigned long long* trace;  // nullable; [n_phases][3] globaltimer stamps per barrier: first arrival, last arrival,
                              // last release (Q100_CHAIN_TRACE=1; the kernel accumulates min / max with atomics)
  const int64_t* const* pos_pp; // device slot holding the pointer to the device-side position all attention phases share
  int trace_phase;            // Q100_CHAIN_TRACE_PHASE: every CTA also stamps its way through this phase (8 stamps / CTA)
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 32 per-lane partial sums -> lane L returns the warp total of element L (31 shuffles instead of 160)
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

__device__ __forceinline__ float neg_inf_f() { return __int_as_float(0xff800000); }

// ---- PH_ATTN: work item = (sequence b, kv head, chunk of chunk_pos positions), one CTA per item.  The 8 consumer
// warps split the chunk's positions; every warp handles the q heads of the GQA group four at a time with an online
// softmax over its positions (lane l owns dims 4l .. 4l+3 of every vector), the warps' results are merged through
// shared memory (`scratch`: 8 x 4 x AT_PART_STRIDE floats) and ONE partial {m, l, o[128]} per (q head, chunk) goes to
// global memory.  The phase that consumes the attention output merges the chunks (x_attn staging, or PH_COMBINE).
// Arithmetic: rotary as K5 (q in fp32 rounded once to T; k with every op rounded to T; v copied), scores and
// softmax in fp32 from the T-rounded q / k / v (what scaled_dot_product_attention does on the same tensors).
template <typename T>
__device__ __noinline__ void chain_attn(const AttnArgs& at, const int64_t p, float* scratch, int cta, int G, int warp,
                                        int lane) {
  using D = DT<T>;
  if (p < 0 || p >= at.max_len || p >= at.rope_len) {
    if (cta == 0 && warp == 0 && lane == 0 && at.err != nullptr) *at.err = 1u;
    return;
  }
  const int rep = at.H / at.KVH;
  const int HG = rep < 4 ? rep : 4;                       // q heads per pass
  const int n_hg = rep / HG;
  const int n_chunks = static_cast<int>((p + at.chunk_pos) / at.chunk_pos);
  const int ppw = at.chunk_pos >> 3;                      // positions per warp
  const int n_items = at.B * at.KVH * n_chunks;
  if (cta >= n_items) return;
  const bool lo_half = lane < 16;                         // dims < 64: partner dims are +64, else -64
  const T* cs = reinterpret_cast<const T*>(at.cosb) + p * AT_D + 4 * lane;
  const T* sn = reinterpret_cast<const T*>(at.sinb) + p * AT_D + 4 * lane;
  const uint2 cw = *reinterpret_cast<const uint2*>(cs), sw = *reinterpret_cast<const uint2*>(sn);
  const float2 c01 = D::to_f2(cw.x), c23 = D::to_f2(cw.y), s01 = D::to_f2(sw.x), s23 = D::to_f2(sw.y);
  const float cc[4] = {c01.x, c01.y, c23.x, c23.y}, ss[4] = {s01.x, s01.y, s23.x, s23.y};
  const int qkv_row = (at.H + 2 * at.KVH) * AT_D;
  for (int item = cta; item < n_items; item += G) {
    const int c = item % n_chunks, bk = item / n_chunks;
    const int kvh = bk % at.KVH, b = bk / at.KVH;
    const T* qkv_b = reinterpret_cast<const T*>(at.qkv) + static_cast<int64_t>(b) * qkv_row;
    T* kc = reinterpret_cast<T*>(at.k_cache) + (static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len * AT_D;
    T* vc = reinterpret_cast<T*>(at.v_cache) + (static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len * AT_D;
    const int64_t p0 = static_cast<int64_t>(c) * at.chunk_pos + warp * ppw;
    if (p >= p0 && p < p0 + ppw) {
      // the new token's k (rotary, every op rounded to T) and v go into the cache slot `p`; the same lanes read
      // them back below (program order), no other work item reads position p of this (b, kv head)
      const uint2 kw = __ldcg(reinterpret_cast<const uint2*>(qkv_b + (at.H + kvh) * AT_D + 4 * lane));
      const uint2 vw = __ldcg(reinterpret_cast<const uint2*>(qkv_b + (at.H + at.KVH + kvh) * AT_D + 4 * lane));
      const float2 k01 = D::to_f2(kw.x), k23 = D::to_f2(kw.y);
      const float kx[4] = {k01.x, k01.y, k23.x, k23.y};
      float ko[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xp = __shfl_xor_sync(0xffffffffu, kx[i], 16);
        const float a0 = D::rd(__fmul_rn(kx[i], cc[i]));
        const float a1 = D::rd(__fmul_rn(lo_half ? -xp : xp, ss[i]));
        ko[i] = __fadd_rn(a0, a1);
      }
      *reinterpret_cast<uint2*>(kc + p * AT_D + 4 * lane) = make_uint2(D::pack2(ko[0], ko[1]), D::pack2(ko[2], ko[3]));
      *reinterpret_cast<uint2*>(vc + p * AT_D + 4 * lane) = vw;
    }
    for (int hg = 0; hg < n_hg; ++hg) {
      const int h0 = kvh * rep + hg * HG;
      float q[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int h = h0 + (i < HG ? i : 0);
        const uint2 qw = __ldcg(reinterpret_cast<const uint2*>(qkv_b + h * AT_D + 4 * lane));
        const float2 q01 = D::to_f2(qw.x), q23 = D::to_f2(qw.y);
        const float qx[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float xp = __shfl_xor_sync(0xffffffffu, qx[d], 16);
          q[i][d] = D::rd(__fadd_rn(__fmul_rn(qx[d], cc[d]), __fmul_rn(lo_half ? -xp : xp, ss[d])));
        }
      }
      float m_run = neg_inf_f(), l_run = 0.f;             // of head (lane & 3), replicated over the other lane bits
      float o[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
      for (int r0 = 0; r0 < ppw; r0 += 8) {
        const int64_t pr = p0 + r0;
        if (pr > p) break;                                // warp-uniform: nothing left below the new token
        // the round's K and V rows are requested together: one memory round trip per round instead of two
        uint2 kw[8], vw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          kw[j] = make_uint2(0u, 0u); vw[j] = make_uint2(0u, 0u);
          if (pr + j <= p) {
            kw[j] = *reinterpret_cast<const uint2*>(kc + (pr + j) * AT_D + 4 * lane);
            vw[j] = *reinterpret_cast<const uint2*>(vc + (pr + j) * AT_D + 4 * lane);
          }
        }
        float s[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 k01 = D::to_f2(kw[j].x), k23 = D::to_f2(kw[j].y);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            s[j * 4 + i] = fmaf(q[i][3], k23.y, fmaf(q[i][2], k23.x, fmaf(q[i][1], k01.y, q[i][0] * k01.x)));
        }
        float sc = warp_reduce_scatter32(s, lane) * at.scale;    // score of (position pr + (lane >> 2), head lane & 3)
        if (pr + (lane >> 2) > p) sc = neg_inf_f();
        float mx = sc;
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float m_new = fmaxf(m_run, mx);               // finite: position pr itself is valid
        const float e = expf(sc - m_new);                   // exp(-inf) = 0 for masked positions
        const float corr = expf(m_run - m_new);
        float es = e;
        es += __shfl_xor_sync(0xffffffffu, es, 4);
        es += __shfl_xor_sync(0xffffffffu, es, 8);
        es += __shfl_xor_sync(0xffffffffu, es, 16);
        l_run = fmaf(l_run, corr, es);
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ci = __shfl_sync(0xffffffffu, corr, i);
          o[i][0] *= ci; o[i][1] *= ci; o[i][2] *= ci; o[i][3] *= ci;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 v01 = D::to_f2(vw[j].x), v23 = D::to_f2(vw[j].y);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float pj = __shfl_sync(0xffffffffu, e, j * 4 + i);
            o[i][0] = fmaf(pj, v01.x, o[i][0]); o[i][1] = fmaf(pj, v01.y, o[i][1]);
            o[i][2] = fmaf(pj, v23.x, o[i][2]); o[i][3] = fmaf(pj, v23.y, o[i][3]);
          }
        }
      }
      // merge the 8 warps: every warp leaves {m, l, o} per head in shared memory, warp i < HG merges head i
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the previous pass's readers are done with `scratch`
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* dst = scratch + (warp * 4 + i) * AT_PART_STRIDE;
        *reinterpret_cast<float4*>(dst + 4 + 4 * lane) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
        if (lane == i) *reinterpret_cast<float2*>(dst) = make_float2(m_run, l_run);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp < HG) {
        float M = neg_inf_f(), L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float* src = scratch + (w * 4 + warp) * AT_PART_STRIDE;
          const float2 ml = *reinterpret_cast<const float2*>(src);
          const float4 ov = *reinterpret_cast<const float4*>(src + 4 + 4 * lane);
          if (ml.y > 0.f) {                                  // l == 0: a warp whose positions were all masked
            const float Mn = fmaxf(M, ml.x);
            const float c0 = expf(M - Mn), c1 = expf(ml.x - Mn);
            L = fmaf(L, c0, ml.y * c1);
            acc[0] = fmaf(acc[0], c0, ov.x * c1); acc[1] = fmaf(acc[1], c0, ov.y * c1);
            acc[2] = fmaf(acc[2], c0, ov.z * c1); acc[3] = fmaf(acc[3], c0, ov.w * c1);
            M = Mn;
          }
        }
        float* dst = at.part + ((static_cast<int64_t>(b) * at.H + h0 + warp) * at.max_parts + c) * AT_PART_STRIDE;
        *reinterpret_cast<float4*>(dst + 4 + 4 * lane) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (lane == 0) *reinterpret_cast<float2*>(dst) = make_float2(M, L);
      }
    }
  }
}

// ---- PH_COMBINE: one warp per (sequence, q head): merge the partials of the chunks up to the new token
template <typename T>
__device__ __noinline__ void chain_combine(const AttnArgs& at, const int64_t p, int cta, int G, int warp, int lane) {
  using D = DT<T>;
  if (p < 0 || p >= at.max_len || p >= at.rope_len) return;
  const int n_parts = static_cast<int>((p + at.chunk_pos) / at.chunk_pos);
  const int n_items = at.B * at.H;
  for (int item = cta * 8 + warp; item < n_items; item += G * 8) {
    const float* base = at.part + static_cast<int64_t>(item) * at.max_parts * AT_PART_STRIDE;
    float M = neg_inf_f(), L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < n_parts; i0 += 4) {
      float2 ml[4];
      float4 ov[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u < n_parts ? i0 + u : n_parts - 1;
        ml[u] = __ldcg(reinterpret_cast<const float2*>(base + i * AT_PART_STRIDE));
        ov[u] = __ldcg(reinterpret_cast<const float4*>(base + i * AT_PART_STRIDE + 4 + 4 * lane));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u < n_parts && ml[u].y > 0.f) {           // l == 0: a warp whose positions were all masked
          const float Mn = fmaxf(M, ml[u].x);
          const float c0 = expf(M - Mn), c1 = expf(ml[u].x - Mn);
          L = fmaf(L, c0, ml[u].y * c1);
          acc[0] = fmaf(acc[0], c0, ov[u].x * c1); acc[1] = fmaf(acc[1], c0, ov[u].y * c1);
          acc[2] = fmaf(acc[2], c0, ov[u].z * c1); acc[3] = fmaf(acc[3], c0, ov[u].w * c1);
          M = Mn;
        }
      }
    }
    const float inv = 1.0f / L;
    T* out = reinterpret_cast<T*>(at.out) + static_cast<int64_t>(item) * AT_D + 4 * lane;
    *reinterpret_cast<uint2*>(out) = make_uint2(D::pack2(acc[0] * inv, acc[1] * inv), D::pack2(acc[2] * inv, acc[3] * inv));
  }
}
#endif  // __CUDACC__

}  // namespace q100
