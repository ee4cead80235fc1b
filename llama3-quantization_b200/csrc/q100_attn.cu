// q100_attn.cu -- K8: the attention core of a decode step as ONE launch (decode form).
//
// Reference code it replaces, per decoder block and new token (models/int_llama_layer.py:114-172 with a KV cache):
//   apply_rotary_pos_emb(q, k) -> cache append -> repeat_kv -> softmax(q k^T / sqrt(d) + mask) -> (.) v
// i.e. what K5 (q100_rope_kv_decode) + the attention library call did in two launches and ~10 us per block.
//
// Work item = (sequence, kv head, chunk of cache positions), one CTA of 4 warps (flash-decoding: the positions are split
// over CTAs so that a batch-1 step still fills the machine).  K and V rows stream through per-warp cp.async rings, the
// GQA group's query heads are the rows of mma.sync.m16n8k16 tiles (S = Q K^T, O += P V), scores / online softmax /
// accumulators in fp32; rotary and the cache append as K5 (q in fp32 rounded once, k with every op rounded to the model
// dtype).  An item that covers all positions writes the output row itself; otherwise it leaves one partial
// {m, l, o[128]} per q head and the LAST item of a (sequence, kv head) to finish -- an arrival counter in global
// memory -- merges the chunks and writes the [batch, heads * 128] row o_proj reads.  No second launch, no host-side
// knowledge of the position: the grid covers max_len, items beyond the device-side position leave at once.
// (The first version ran the decode chain's attention phase, chain_attn of q100_chain.cuh, as a kernel: fp32 FMAs and
// warp shuffles -- 1.8 TB/s at long caches and several waves of 153-register CTAs at batch 16; measured, replaced.)
#include <cstdlib>
#include <mutex>

#include "q100_chain.cuh"
#include "q100_ptx.cuh"

namespace q100 {

namespace {

constexpr int A8_WARPS = 4;                       // 16 cache positions per warp and step
constexpr int A8_STEP = A8_WARPS * 16;            // positions per CTA and step
constexpr int A8_ROW = AT_D * 2;                  // bytes of one K / V / q row
constexpr int A8_TILE = 16 * A8_ROW;              // one warp's K (or V) rows of a step: 4 KB
constexpr int A8_STAGES = 3;                      // cp.async ring depth per warp
constexpr int A8_WARP_SMEM = A8_STAGES * 2 * A8_TILE;   // stages x (K, V)
constexpr int A8_Q_OFF = A8_WARPS * A8_WARP_SMEM; // q rows [16][128] T behind the stages
constexpr int A8_SMEM = A8_Q_OFF + 16 * A8_ROW;
constexpr float A8_LOG2E = 1.4426950408889634f;

// 16-byte chunk c of row r of a [rows][128] 16-bit tile, XOR-swizzled so that ldmatrix reads are conflict-free
__device__ __forceinline__ uint32_t a8_off(int r, int c) { return static_cast<uint32_t>(r * A8_ROW + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
}

// Work item = (sequence, kv head, chunk of `chunk_pos` cache positions), one CTA of 4 warps.  Per step every warp owns 16
// positions: their K and V rows come in with cp.async (16 B per lane and request, swizzled rows, private to the warp: no
// CTA barrier in the loop, two stages), the new token's own row is written by the CTA (rotary k as K5: every op rounded to T)
// into shared memory AND the cache.  S = Q K^T and O += P V run on mma.sync.m16n8k16 with the GQA group's query heads as
// the 16 rows of the A operand (rows >= heads / kv_heads are zero): 32 MMAs per 16 positions; scores, online softmax and the
// accumulators in fp32, P rounded to T for the second MMA (the reference rounds its probabilities to T as well,
// int_llama_layer.py:152-153).  The warps merge through shared memory; an item that covers all positions 0..p writes the
// output row itself, otherwise it leaves a partial {m, l, o[128]} per head and the last item of the (sequence, kv head)
// -- an arrival counter -- merges the chunks.
template <typename T>
__global__ void __launch_bounds__(A8_WARPS * 32, 3) k8_attn_decode(const AttnArgs at, uint32_t* __restrict__ counters) {
  using D = DT<T>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c4 = lane & 3;
  const int rep = at.H / at.KVH;
  const int bk = static_cast<int>(blockIdx.x) / at.max_parts, chunk = static_cast<int>(blockIdx.x) % at.max_parts;
  const int b = bk / at.KVH, kvh = bk % at.KVH;
  if (at.pad_ == 0) pdl_launch_dependents();   // the o_proj GEMM behind may fetch its first weights now
  pdl_wait();                       // q|k|v (and the previous block's use of the partials) are complete
  // everything that does not depend on the position is requested before the position is read: thread d <-> dim d
  const int d = tid, dp = d < 64 ? d + 64 : d - 64;
  const T* qkv_b = reinterpret_cast<const T*>(at.qkv) + static_cast<int64_t>(b) * (at.H + 2 * at.KVH) * AT_D;
  T qx[8], qp[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const T* qh = qkv_b + (kvh * rep + (h < rep ? h : 0)) * AT_D;
    qx[h] = qh[d]; qp[h] = qh[dp];
  }
  const T* kh = qkv_b + (at.H + kvh) * AT_D;
  const T knx = kh[d], knp = kh[dp], vnew = qkv_b[(at.H + at.KVH + kvh) * AT_D + d];
  const int64_t p = *at.pos;
  if (p < 0 || p >= at.max_len || p >= at.rope_len) {
    if (blockIdx.x == 0 && tid == 0 && at.err != nullptr) *at.err = 1u;
    return;
  }
  const int n_chunks = static_cast<int>(p / at.chunk_pos) + 1;
  if (chunk >= n_chunks) return;
  const int64_t c0 = static_cast<int64_t>(chunk) * at.chunk_pos;          // first position of the item
  const int n_steps = static_cast<int>((min(p, c0 + at.chunk_pos - 1) - c0) / A8_STEP) + 1;
  const T* kc = reinterpret_cast<const T*>(at.k_cache) + (static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len * AT_D;
  const T* vc = reinterpret_cast<const T*>(at.v_cache) + (static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len * AT_D;
  const uint32_t sm0 = smem_u32(smem);
  const uint32_t wsm = sm0 + warp * A8_WARP_SMEM;
  const float cs = D::to_f(reinterpret_cast<const T*>(at.cosb)[p * AT_D + d]);
  const float sn = D::to_f(reinterpret_cast<const T*>(at.sinb)[p * AT_D + d]);

  // ---- the stream: rows < p by cp.async (rows > p zero-filled: 0 * garbage must not become NaN), row p is written below
  auto issue = [&](int step) {
    const int64_t p0 = c0 + static_cast<int64_t>(step) * A8_STEP + warp * 16;
    const uint32_t kst = wsm + (step % A8_STAGES) * (2 * A8_TILE), vst = kst + A8_TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 2 + (lane >> 4), ch = lane & 15;                  // 16 rows x 16 chunks per tile
      const int64_t pos = p0 + r;
      if (pos != p) {
        const bool ok = pos < p;
        const int64_t src = (ok ? pos : 0) * AT_D + ch * 8;
        cp_async16_zfill(kst + a8_off(r, ch), kc + src, ok);
        cp_async16_zfill(vst + a8_off(r, ch), vc + src, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int st = 0; st < A8_STAGES - 1; ++st) {
    if (st < n_steps) issue(st);
    else asm volatile("cp.async.commit_group;" ::: "memory");             // keep the group count uniform
  }

  // ---- rotary.  q in fp32, one rounding (int_llama_layer.py:117, 125); k with every op rounded to T
  {
    T* qs = reinterpret_cast<T*>(smem + A8_Q_OFF);
#pragma unroll
    for (int h = 0; h < 16; ++h) {
      float v = 0.f;
      if (h < 8 && h < rep) {
        const float x = D::to_f(qx[h & 7]), xp = D::to_f(qp[h & 7]);
        v = __fadd_rn(__fmul_rn(x, cs), __fmul_rn(d < 64 ? -xp : xp, sn));
      }
      *reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(qs) + a8_off(h, d >> 3) + (d & 7) * 2) = D::from_f(v);
    }
    if (p >= c0 && p < c0 + at.chunk_pos) {           // this item owns the new token's slot
      const float x = D::to_f(knx), xp = D::to_f(knp);
      const T kr = D::from_f(__fadd_rn(D::rd(__fmul_rn(x, cs)), D::rd(__fmul_rn(d < 64 ? -xp : xp, sn))));
      reinterpret_cast<T*>(at.k_cache)[((static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len + p) * AT_D + d] = kr;
      reinterpret_cast<T*>(at.v_cache)[((static_cast<int64_t>(b) * at.KVH + kvh) * at.max_len + p) * AT_D + d] = vnew;
      const int rel = static_cast<int>(p - c0), step = rel / A8_STEP, w = (rel % A8_STEP) >> 4, r = rel & 15;
      // the slots of the first A8_STAGES - 1 steps are distinct and already requested (without this row); a later step's
      // copy of the row is re-created from the cache by its warp (see the loop)
      if (step < A8_STAGES - 1) {
        uint8_t* kst = smem + w * A8_WARP_SMEM + step * (2 * A8_TILE);
        *reinterpret_cast<T*>(kst + a8_off(r, d >> 3) + (d & 7) * 2) = kr;
        *reinterpret_cast<T*>(kst + A8_TILE + a8_off(r, d >> 3) + (d & 7) * 2) = vnew;
      }
    }
  }
  __syncthreads();                                    // q rows (and the new token's rows) are in shared memory

  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};        // heads g, g + 8 (log2 domain)
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  const float sc2 = at.scale * A8_LOG2E;
  const uint32_t qsm = sm0 + A8_Q_OFF;
  for (int step = 0; step < n_steps; ++step) {
    if (step + A8_STAGES - 1 < n_steps) issue(step + A8_STAGES - 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" :: "n"(A8_STAGES - 1) : "memory");
    __syncwarp();
    const int64_t p0 = c0 + static_cast<int64_t>(step) * A8_STEP + warp * 16;
    const uint32_t kst = wsm + (step % A8_STAGES) * (2 * A8_TILE), vst = kst + A8_TILE;
    if (step >= A8_STAGES - 1 && p >= p0 && p < p0 + 16) {
      // the new token's row of a later step: this warp re-creates it from the cache (the CTA wrote it before the barrier)
      const int r = static_cast<int>(p - p0);
      if (lane < 16) {
        *reinterpret_cast<uint4*>(smem + (kst - sm0) + a8_off(r, lane)) = __ldcg(reinterpret_cast<const uint4*>(kc + p * AT_D) + lane);
      } else {
        *reinterpret_cast<uint4*>(smem + (vst - sm0) + a8_off(r, lane - 16)) = __ldcg(reinterpret_cast<const uint4*>(vc + p * AT_D) + lane - 16);
      }
      __syncwarp();
    }
    if (p0 > p) continue;                             // warp-uniform: none of these positions exists yet
    // ---- S = Q K^T: rows = query heads, two n-tiles of 8 positions
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {                     // two k-steps of 16 dims per u
      uint32_t kb0[4], kb1[4], qa[4];
      ldsm_x4(kst + a8_off(lane & 7, 4 * u + (lane >> 3)), kb0);              // positions 0-7, dims 32u .. 32u+31
      ldsm_x4(kst + a8_off(8 + (lane & 7), 4 * u + (lane >> 3)), kb1);        // positions 8-15
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        ldsm_x4(qsm + a8_off((lane & 7) + ((lane >> 3) & 1) * 8, 2 * (2 * u + h2) + (lane >> 4)), qa);
        mma16816<T>(s0, qa, kb0[2 * h2], kb0[2 * h2 + 1]);
        mma16816<T>(s1, qa, kb1[2 * h2], kb1[2 * h2 + 1]);
      }
    }
    // ---- online softmax; thread (g, c4) holds head g at positions 2 c4, 2 c4 + 1 (+ 8) in [0], [1]; head g + 8 in [2], [3]
    float sv[2][4];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      sv[hh][0] = s0[2 * hh] * sc2; sv[hh][1] = s0[2 * hh + 1] * sc2;
      sv[hh][2] = s1[2 * hh] * sc2; sv[hh][3] = s1[2 * hh + 1] * sc2;
      if (p0 + 2 * c4 > p) sv[hh][0] = -INFINITY;
      if (p0 + 2 * c4 + 1 > p) sv[hh][1] = -INFINITY;
      if (p0 + 8 + 2 * c4 > p) sv[hh][2] = -INFINITY;
      if (p0 + 9 + 2 * c4 > p) sv[hh][3] = -INFINITY;
    }
    uint32_t pa[4];
    float corr[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float mx = fmaxf(fmaxf(sv[hh][0], sv[hh][1]), fmaxf(sv[hh][2], sv[hh][3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[hh], mx);       // finite: position p0 <= p is visible
      corr[hh] = exp2f(m_run[hh] - m_new);
      const float e0 = exp2f(sv[hh][0] - m_new), e1 = exp2f(sv[hh][1] - m_new);
      const float e2 = exp2f(sv[hh][2] - m_new), e3 = exp2f(sv[hh][3] - m_new);
      // the second MMA takes P in T: the row sum uses the same rounded values
      const uint32_t p01 = D::pack2(e0, e1), p23 = D::pack2(e2, e3);
      const float2 f01 = D::to_f2(p01), f23 = D::to_f2(p23);
      float es = (f01.x + f01.y) + (f23.x + f23.y);
      es += __shfl_xor_sync(0xffffffffu, es, 1);
      es += __shfl_xor_sync(0xffffffffu, es, 2);
      l_run[hh] = fmaf(l_run[hh], corr[hh], es);
      m_run[hh] = m_new;
      pa[hh] = p01;                                   // a0 (rows g) / a1 (rows g + 8): positions 2 c4, 2 c4 + 1
      pa[2 + hh] = p23;                               // a2 / a3: positions 8 + 2 c4, 9 + 2 c4
    }
    // ---- O = O * corr + P V
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {
      uint32_t vb[4];
      ldsm_x4_t(vst + a8_off((lane & 7) + ((lane >> 3) & 1) * 8, 2 * n2 + (lane >> 4)), vb);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        float (&oo)[4] = o[2 * n2 + h2];
        oo[0] *= corr[0]; oo[1] *= corr[0]; oo[2] *= corr[1]; oo[3] *= corr[1];
        mma16816<T>(oo, pa, vb[2 * h2], vb[2 * h2 + 1]);
      }
    }
  }

  // ---- merge the warps through shared memory (the stages are free now), thread d <-> dim d
  __syncthreads();
  float* scr = reinterpret_cast<float*>(smem);        // [warp][head 8][AT_PART_STRIDE]
  if (rep <= 8) {
    if (g < rep) {
      float* dst = scr + (warp * 8 + g) * AT_PART_STRIDE;
#pragma unroll
      for (int n = 0; n < 16; ++n) *reinterpret_cast<float2*>(dst + 4 + 8 * n + 2 * c4) = make_float2(o[n][0], o[n][1]);
      if (c4 == 0) *reinterpret_cast<float2*>(dst) = make_float2(m_run[0], l_run[0]);
    }
  }
  __syncthreads();
  const bool direct = n_chunks == 1;
  for (int h = 0; h < rep; ++h) {
    float M = -INFINITY, L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < A8_WARPS; ++w) {
      const float* src = scr + (w * 8 + h) * AT_PART_STRIDE;
      const float2 ml = *reinterpret_cast<const float2*>(src);
      if (ml.y > 0.f) {                               // l == 0: a warp whose positions were all masked
        const float Mn = fmaxf(M, ml.x);
        const float a0 = exp2f(M - Mn), a1 = exp2f(ml.x - Mn);
        L = fmaf(L, a0, ml.y * a1);
        acc = fmaf(acc, a0, src[4 + tid] * a1);
        M = Mn;
      }
    }
    const int item = b * at.H + kvh * rep + h;
    if (direct) {
      reinterpret_cast<T*>(at.out)[static_cast<int64_t>(item) * AT_D + tid] = D::from_f(acc / L);
    } else {
      float* dst = at.part + (static_cast<int64_t>(item) * at.max_parts + chunk) * AT_PART_STRIDE;
      dst[4 + tid] = acc;
      if (tid == 0) *reinterpret_cast<float2*>(dst) = make_float2(M, L);
    }
  }
  if (direct) return;
  __threadfence();                  // this CTA's partials are visible before its arrival is
  __syncthreads();
  if (tid == 0) {
    const uint32_t old = atomicAdd(counters + bk, 1u);
    const int last = old == static_cast<uint32_t>(n_chunks - 1);
    if (last) counters[bk] = 0u;    // zero between launches
    __threadfence();
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  // the last item of the (sequence, kv head): warp w merges heads w, w + 4; lane = 4 dims; eight chunks per round trip
  for (int h = warp; h < rep; h += A8_WARPS) {
    const int item = b * at.H + kvh * rep + h;
    const float* base = at.part + static_cast<int64_t>(item) * at.max_parts * AT_PART_STRIDE;
    float M = -INFINITY, L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < n_chunks; i0 += 8) {
      float2 ml[8];
      float4 ov[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u < n_chunks ? i0 + u : n_chunks - 1;
        ml[u] = __ldcg(reinterpret_cast<const float2*>(base + i * AT_PART_STRIDE));
        ov[u] = __ldcg(reinterpret_cast<const float4*>(base + i * AT_PART_STRIDE + 4 + 4 * lane));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (i0 + u < n_chunks && ml[u].y > 0.f) {
          const float Mn = fmaxf(M, ml[u].x);
          const float a0 = exp2f(M - Mn), a1 = exp2f(ml[u].x - Mn);
          L = fmaf(L, a0, ml[u].y * a1);
          acc[0] = fmaf(acc[0], a0, ov[u].x * a1); acc[1] = fmaf(acc[1], a0, ov[u].y * a1);
          acc[2] = fmaf(acc[2], a0, ov[u].z * a1); acc[3] = fmaf(acc[3], a0, ov[u].w * a1);
          M = Mn;
        }
      }
    }
    const float inv = 1.0f / L;
    T* out = reinterpret_cast<T*>(at.out) + static_cast<int64_t>(item) * AT_D + 4 * lane;
    *reinterpret_cast<uint2*>(out) = make_uint2(D::pack2(acc[0] * inv, acc[1] * inv), D::pack2(acc[2] * inv, acc[3] * inv));
  }
}

// positions per item.  Measured (tools/exp_attn.py, 8B head shapes): every item pays ~2 us of fixed latency (prologue,
// merge, arrival), so as few items as fill the machine once: ~1.5 per SM over all (sequence, kv head) pairs, at most 32 per
// pair, and never more than 2048 positions per item (a long cache at a large batch still spreads over the SMs).
int attn_chunk(int64_t max_len, int batch, int kv_heads) {
  static const int env_target = []() { const char* e = getenv("Q100_K8_CHUNKS"); return e ? atoi(e) : 0; }();   // experiments
  int target = 222 / (batch * kv_heads);
  target = target < 1 ? 1 : (target > 32 ? 32 : target);
  if (env_target > 0) target = env_target;
  int chunk = 64 * static_cast<int>((max_len + 64 * target - 1) / (64 * target));
  if (chunk > 2048) chunk = 2048;
  const int min_chunk = 64 * static_cast<int>((max_len + 2047) / 2048);      // <= 32 chunks
  return chunk < min_chunk ? min_chunk : chunk;
}

}  // namespace

}  // namespace q100

using namespace q100;

extern "C" size_t q100_attn_decode_workspace_bytes(int batch, int heads, int kv_heads, int64_t max_len) {
  if (batch <= 0 || heads <= 0 || kv_heads <= 0 || max_len <= 0) return 0;
  const int cp = attn_chunk(max_len, batch, kv_heads);
  const int64_t parts = (max_len + cp - 1) / cp;
  const size_t counters = (static_cast<size_t>(batch) * kv_heads * 4 + 255) & ~static_cast<size_t>(255);
  return counters + static_cast<size_t>(batch) * heads * parts * AT_PART_STRIDE * 4;
}

extern "C" int q100_attn_decode(const void* qkv, const void* cos_cache, const void* sin_cache, const void* position,
                                void* k_cache, void* v_cache, void* out, void* workspace, size_t workspace_bytes,
                                int batch, int heads, int kv_heads, int head_dim, int64_t max_len, int64_t rope_len,
                                void* err_flag, int dtype, void* stream) {
  Q100_REQUIRE(qkv && cos_cache && sin_cache && position && k_cache && v_cache && out && workspace, Q100_EINVAL,
               "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(batch >= 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0 && max_len > 0 && rope_len > 0,
               Q100_EINVAL, "bad shape");
  Q100_REQUIRE(head_dim == AT_D && heads / kv_heads <= 8, Q100_EUNSUPPORTED,
               "q100_attn_decode: head_dim 128 and at most 8 query heads per kv head (got %d, %d)", head_dim,
               heads / kv_heads);
  if (batch == 0) return Q100_OK;
  Q100_REQUIRE(workspace_bytes >= q100_attn_decode_workspace_bytes(batch, heads, kv_heads, max_len), Q100_EWORKSPACE,
               "workspace too small: %zu bytes", workspace_bytes);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  AttnArgs at{};
  at.qkv = qkv; at.cosb = cos_cache; at.sinb = sin_cache;
  at.pos = reinterpret_cast<const int64_t*>(position);
  at.k_cache = k_cache; at.v_cache = v_cache;
  const size_t counters = (static_cast<size_t>(batch) * kv_heads * 4 + 255) & ~static_cast<size_t>(255);
  at.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + counters);
  at.out = out;
  at.B = batch; at.H = heads; at.KVH = kv_heads;
  at.chunk_pos = attn_chunk(max_len, batch, kv_heads);
  at.max_parts = static_cast<int>((max_len + at.chunk_pos - 1) / at.chunk_pos);
  at.max_len = max_len; at.rope_len = rope_len;
  at.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  at.err = reinterpret_cast<uint32_t*>(err_flag);
  // the successor (o_proj) is released at once only where that was measured to pay: behind it runs the row-owner GEMM
  // (batch <= 2) or the tcgen05 streaming GEMM (batch > 8); an early-started stream-K GEMM (batch 3-8) cost 4-6 us per block
  static const int early_env = []() { const char* e = getenv("Q100_K8_EARLY"); return e ? atoi(e) : -1; }();   // experiments
  const bool early = early_env >= 0 ? early_env != 0 : (batch <= 2 || batch > 8);
  at.pad_ = early ? 0 : 1;                          // 1: the successor is released when this grid completes
  static const bool pdl_on = []() { const char* e = getenv("Q100_BLOCK_PDL"); return !(e && atoi(e) == 0); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(batch * kv_heads * at.max_parts));
  cfg.blockDim = dim3(A8_WARPS * 32);
  cfg.dynamicSmemBytes = A8_SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_on ? 1 : 0;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(workspace);
  {
    static std::mutex mu;
    static bool attr_done[kMaxDevices] = {};
    std::lock_guard<std::mutex> lk(mu);
    const int dev = current_device();
    if (!attr_done[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(k8_attn_decode<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, A8_SMEM));
      Q100_CUDA(cudaFuncSetAttribute(k8_attn_decode<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, A8_SMEM));
      attr_done[dev] = true;
    }
  }
  cudaError_t e;
  if (dtype == Q100_F16) e = cudaLaunchKernelEx(&cfg, k8_attn_decode<__half>, at, cnt);
  else e = cudaLaunchKernelEx(&cfg, k8_attn_decode<__nv_bfloat16>, at, cnt);
  return check_cuda(e, "k8_attn_decode launch");
}
