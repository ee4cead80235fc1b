// q100_attn.cu -- K8: the attention core of a decode step as ONE launch (decode form).
//
// Reference code it replaces, per decoder block and new token (models/int_llama_layer.py:114-172 with a KV cache):
//   apply_rotary_pos_emb(q, k) -> cache append -> repeat_kv -> softmax(q k^T / sqrt(d) + mask) -> (.) v
// i.e. what K5 (q100_rope_kv_decode) + the attention library call did in two launches and ~10 us per block.
//
// Work item = (sequence, kv head, chunk of cache positions): one CTA of 8 warps each (flash-decoding: the positions
// are split over CTAs so that a batch-1 step still fills the machine; a first K8 with one CTA per q head was
// latency-bound).  The item code is chain_attn of q100_chain.cuh -- the attention phase of the decode chain
// K9 -- so both routes share one arithmetic: rotary as K5 (q in fp32 rounded once, k with every op rounded to the
// model dtype, the new k / v appended to the cache by the item that owns the position), scores / online softmax /
// P.V in fp32 from the 16-bit q, k, v.  The item leaves one partial {m, l, o[128]} per q head; the LAST item of a
// (sequence, kv head) to finish -- an arrival counter in global memory -- merges the chunks and writes the
// [batch, heads * 128] row o_proj reads.  No second launch, no host-side knowledge of the position: the grid covers
// max_len, items beyond the device-side position leave at once.
#include <cstdlib>

#include "q100_chain.cuh"
#include "q100_ptx.cuh"

namespace q100 {

namespace {

// merge the first n_parts partials of (sequence, q head) `item` and write the normalised output row (one warp)
template <typename T>
__device__ __forceinline__ void attn_merge_item(const AttnArgs& at, int n_parts, int item, int lane) {
  using D = DT<T>;
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
      if (i0 + u < n_parts && ml[u].y > 0.f) {
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

template <typename T>
__global__ void __launch_bounds__(256) k8_attn_decode(const AttnArgs at, uint32_t* __restrict__ counters) {
  __shared__ __align__(16) float scratch[8 * 4 * AT_PART_STRIDE];
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();          // the o_proj GEMM behind may fetch its first weights now
  pdl_wait();                       // q|k|v (and the previous block's use of the partials) are complete
  const int64_t p = *at.pos;
  chain_attn<T>(at, p, scratch, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), warp, lane);
  if (p < 0 || p >= at.max_len || p >= at.rope_len) return;          // chain_attn raised the error flag
  const int n_chunks = static_cast<int>((p + at.chunk_pos) / at.chunk_pos);
  if (static_cast<int>(blockIdx.x) >= at.B * at.KVH * n_chunks) return;
  const int bk = static_cast<int>(blockIdx.x) / n_chunks;              // (sequence, kv head) of this item
  __threadfence();                  // this CTA's partials are visible before its arrival is
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(counters + bk, 1u);
    const int last = old == static_cast<uint32_t>(n_chunks - 1);
    if (last) counters[bk] = 0u;    // zero between launches
    __threadfence();
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  const int rep = at.H / at.KVH;
  const int b = bk / at.KVH, kvh = bk % at.KVH;
  for (int i = warp; i < rep; i += 8) attn_merge_item<T>(at, n_chunks, b * at.H + kvh * rep + i, lane);
}

int attn_chunk(int64_t max_len) { return 64 * static_cast<int>((max_len + 2047) / 2048); }   // <= 32 chunks per sequence

}  // namespace

}  // namespace q100

using namespace q100;

extern "C" size_t q100_attn_decode_workspace_bytes(int batch, int heads, int kv_heads, int64_t max_len) {
  if (batch <= 0 || heads <= 0 || kv_heads <= 0 || max_len <= 0) return 0;
  const int cp = attn_chunk(max_len);
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
  const int rep = heads / kv_heads;
  Q100_REQUIRE(rep < 4 || rep % 4 == 0, Q100_EUNSUPPORTED, "q100_attn_decode: query heads per kv head must be 1-3, 4 or 8");
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
  at.chunk_pos = attn_chunk(max_len);
  at.max_parts = static_cast<int>((max_len + at.chunk_pos - 1) / at.chunk_pos);
  at.max_len = max_len; at.rope_len = rope_len;
  at.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  at.err = reinterpret_cast<uint32_t*>(err_flag);
  static const bool pdl_on = []() { const char* e = getenv("Q100_BLOCK_PDL"); return !(e && atoi(e) == 0); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(batch * kv_heads * at.max_parts));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_on ? 1 : 0;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(workspace);
  cudaError_t e;
  if (dtype == Q100_F16) e = cudaLaunchKernelEx(&cfg, k8_attn_decode<__half>, at, cnt);
  else e = cudaLaunchKernelEx(&cfg, k8_attn_decode<__nv_bfloat16>, at, cnt);
  return check_cuda(e, "k8_attn_decode launch");
}
