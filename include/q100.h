/*
 * q100.h -- C ABI of libq100.so: sm_100a kernels for the group-wise quantized-linear
 * hot path of SilviaUvA/LLaMA3-Quantization.
 *
 * The reference has no FFI / plugin interface for this path: it is pure PyTorch and the
 * module boundary (quant/quantizer.py, quant/int_linear.py, quant/int_matmul.py) *is*
 * the interface.  Each entry point below names the reference code it replaces; the
 * Python drop-in modules in llama3-quantization_b200/quant/ bind them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors); the
 *     library allocates nothing that outlives a call except an internal cache of TMA
 *     descriptors;
 *   - `stream` is a cudaStream_t / CUstream passed as void*; kernels are enqueued on
 *     it with no host synchronisation;
 *   - return value 0 == success, otherwise a Q100_E* code; q100_last_error() returns
 *     a thread-local human readable message.  No exceptions cross the ABI;
 *   - dtype codes: 0 = fp16, 1 = bf16, 2 = fp32.  Scales / zero-points / LWC
 *     multipliers are stored in the same dtype as the tensor they belong to, exactly
 *     like the reference's `scale` / `round_zero_point` (quant/quantizer.py:149-159).
 *   - there is no CPU path: every compute entry point fails with Q100_ENODEVICE when
 *     no sm_100 device is present.
 */
#ifndef Q100_H_
#define Q100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define Q100_ABI_VERSION 2

enum q100_dtype { Q100_F16 = 0, Q100_BF16 = 1, Q100_F32 = 2 };

enum q100_status {
  Q100_OK = 0,
  Q100_EINVAL = 1,      /* bad shape / dtype / flag combination                   */
  Q100_EUNSUPPORTED = 2,/* valid for the reference, not implemented by a kernel   */
  Q100_ECUDA = 3,       /* a CUDA runtime / driver call failed (see last_error)   */
  Q100_ENODEVICE = 4,   /* no CUDA device, or device is not sm_100                */
  Q100_EWORKSPACE = 5   /* workspace too small (see q100_gemm_workspace_bytes)    */
};

/* flags for the quantize entry points */
#define Q100_FLAG_IEEE_DIV    0x1  /* f16/bf16: use div.rn.f32 for x/scale instead of the
                                      rcp + 2-FMA exact-for-16-bit sequence (same bits, slower) */
#define Q100_FLAG_SCALAR_MATH 0x2  /* f16/bf16: per-element fp32 math instead of packed half2 ops */
#define Q100_FLAG_K1_GENERIC  0x4  /* f16/bf16, group 64/128/256: the first group kernel (k1_groups) instead of
                                      k1_fast (same bits; kept for A/B parity tests and profiling)          */
#define Q100_FLAG_CUDA_SCALE   0x8  /* scale = range * RN(1 / (2^n - 1)) in fp32 -- what the reference computes when it runs
                                      on CUDA tensors (torch's CUDA kernels turn `tensor / python_scalar`,
                                      quant/quantizer.py:147,152, into a multiplication by the reciprocal) -- instead of
                                      the true division its CPU kernels perform (the default, pinned by the CPU golden
                                      vectors).  The two differ by at most one ulp of the scale.  Honoured by the quantize
                                      entry points and by q100_quantize_backward. */
/* flags for the GEMM entry points */
#define Q100_GEMM_FORCE_SKINNY 0x1 /* always use the mma.sync weight-streaming kernel  */
#define Q100_GEMM_FORCE_TC     0x2 /* always use the tcgen05 kernel                    */
#define Q100_GEMM_EXACT_WEIGHTS 0x4 /* decode kernel: rebuild the 16-bit fake-quantised weight
                                      r(r(q-zp)*scale) bit-exactly before the MMA (slower); default
                                      keeps (q-zp) exact and applies the scale in fp32 per group   */
#define Q100_GEMM_FORCE_MMA_SYNC 0x10 /* decode shapes: always the mma.sync streaming kernel                 */
#define Q100_GEMM_FORCE_STREAM_TC 0x20 /* decode shapes: the tcgen05 streaming kernel (weights as the TMEM A
                                      operand) wherever it applies (M <= 16, W2/W4, group % 128 == 0);
                                      default: 8 < M <= 16                                           */
#define Q100_GEMM_FORCE_TC_SMEM 0x40 /* prefill: the first tcgen05 kernel (dequantised weights staged in
                                      shared memory as the B operand) instead of the TMEM-A kernel      */
#define Q100_GEMM_NO_IMMA      0x80 /* decode shapes, W4 / W2, M <= 2: the fp16 magic-form mma.sync path instead of the
                                      integer (u8 x s8 limbs of x) mma.sync path                      */
#define Q100_GEMM_X_GATE_UP    0x100 /* decode shapes (M <= 32): x is [M, 2K] = gate | up of the MLP and the kernel
                                      applies act_fn(gate) * up (SiLU, the arithmetic of q100_silu_mul, i.e. of
                                      models/int_llama_layer.py:45) while it stages x: down_proj without the
                                      separate element-wise launch.  Q100_EUNSUPPORTED for M > 32.      */
#define Q100_GEMM_NO_ROWS      0x200 /* decode shapes, W4 g128, M <= 8: the stream-K kernel (K3a, tiles shared between CTAs and
                                      fixed up through the workspace) instead of the row-owner kernel (K3r: every CTA
                                      streams whole rows, no fix-up), which is the default where it applies      */
#define Q100_GEMM_PDL          0x8 /* launch with programmatic stream serialization: the kernel starts
                                      streaming its (static) weights while the previous kernel of the
                                      stream drains, and waits (griddepcontrol.wait) before it touches
                                      x, y or the workspace.  Only valid when `packed`/`meta` were not
                                      written by the immediately preceding q100 GEMM launch.        */

const char* q100_last_error(void);
int q100_abi_version(void);
/* 0 if device `dev` exists and is compute capability 10.x, else Q100_ENODEVICE. */
int q100_device_ok(int dev);

/*
 * K1/K2 -- UniformAffineQuantizer.forward for a 2-D view [rows, cols] (contiguous).
 * Replaces per_token_dynamic_calibration + fake_quant, quant/quantizer.py:94-159
 * (17 eager elementwise/reduction launches) with one pass.
 *
 *   group       quantisation group along cols (quantizer.py:135 reshape(-1, g));
 *               0 = one group per row (group_size=None: per-channel weights,
 *               per-token activations).  cols % group must be 0 (the reference's
 *               LWC zero-padding for K % g != 0, quantizer.py:136-139, is done by
 *               the caller).
 *   n_bits      2..8 (packed output) or 2..15 (dequant output only)
 *   symmetric / disable_zp   quantizer.py:146-159
 *   lwc_up/lwc_low   nullable; POST-sigmoid multipliers [rows*cols/group] in `dtype`
 *               (quantizer.py:143-145; torch computes the sigmoid)
 *   dequant_out nullable; [rows, cols] in `dtype`      == forward()'s return value
 *   packed_out  nullable; uint32 [q100_packed_words(rows, cols, n_bits)], tile-major layout v2
 *               (DESIGN.md "Packed layout"); must be zero-filled by the caller when
 *               rows % 128 != 0 (padding rows are never written)
 *   scales_out  nullable; [rows*cols/group] in `dtype` == self.scale
 *   zeros_out   nullable; [rows*cols/group] in `dtype` == self.round_zero_point
 *               (ignored when disable_zp)
 *   meta_out    nullable; packed-weight side band for the GEMM kernels: interleaved
 *               (scale, zero) pairs in `dtype`, TRANSPOSED to [cols/group, rows, 2]
 *               so one K-group of consecutive rows is contiguous.  fp16/bf16 only.
 */
int q100_quantize_groupwise(const void* x, int dtype, int64_t rows, int64_t cols, int group,
                            int n_bits, int symmetric, int disable_zp,
                            const void* lwc_up, const void* lwc_low,
                            void* dequant_out, void* packed_out,
                            void* scales_out, void* zeros_out, void* meta_out,
                            int flags, void* stream);

/*
 * K1b -- backward of the group-wise fake-quantizer (what torch autograd computes through
 * UniformAffineQuantizer.forward, quant/quantizer.py:94-159, in the calibration loop of
 * quant/omniquant.py:261-293): straight-through estimator for round(), clamp mask, the scale's
 * dependence on the group's min / max (shared equally by tied positions) and on the LWC
 * multipliers; no gradient through the zero-point (quantizer.py:158 rounds it without STE).
 *   x, grad_out  [rows, cols] in `dtype` (x = the forward's input, grad_out = d loss / d dequant)
 *   group        32, 64, 128 or 256 elements (Q100_EUNSUPPORTED otherwise: the caller keeps its own path)
 *   lwc_up/low   nullable, [rows*cols/group] post-sigmoid multipliers as given to the forward
 *   grad_x       nullable; [rows, cols]
 *   grad_up/low  nullable; [rows*cols/group] d loss / d multiplier (only written with LWC)
 *   flags        Q100_FLAG_CUDA_SCALE must match the forward's
 */
int q100_quantize_backward(const void* x, const void* grad_out, int dtype, int64_t rows, int64_t cols,
                           int group, int n_bits, int symmetric, int disable_zp,
                           const void* lwc_up, const void* lwc_low,
                           void* grad_x, void* grad_up, void* grad_low, int flags, void* stream);

/*
 * K2 -- per-token activation fake-quant (asymmetric, one group per row), the
 * act_quantizer / x1_quantizer / x2_quantizer call of quant/int_linear.py:59-60 and
 * quant/int_matmul.py:31-39.  Thin alias of q100_quantize_groupwise(group = 0).
 */
int q100_fakequant_rows(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits,
                        void* out, void* scales_out, void* zeros_out, int flags, void* stream);

/*
 * K2 with the row statistics supplied from outside -- what per_token_dynamic_calibration
 * (quant/quantizer.py:140-141, amin / amax over the last dim) needs when that dim is sharded over
 * tensor-parallel ranks (inputs of the row-parallel o_proj / down_proj with abits < 16, and per-channel
 * weights of those layers): every rank takes the min / max of its slice with q100_rows_minmax, the
 * [rows, 2] pairs are MIN / MAX all-reduced by the caller, and q100_quantize_rows_stats quantises the
 * local slice with the global statistics -- bit-identical to slicing the single-GPU result.
 *   minmax      [rows, 2] (min, max) in `dtype`
 *   other arguments as q100_quantize_groupwise with group = 0 (one group per row)
 */
int q100_rows_minmax(const void* x, int dtype, int64_t rows, int64_t cols, void* minmax_out, void* stream);
int q100_quantize_rows_stats(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits,
                             int symmetric, int disable_zp, const void* minmax,
                             const void* lwc_up, const void* lwc_low,
                             void* dequant_out, void* packed_out, void* scales_out, void* zeros_out,
                             void* meta_out, int flags, void* stream);

/* Total uint32 words of a packed [rows, cols] code matrix (q100 v2 tile-major layout: tiles of
 * 128 rows x 64 k, ordered [n_tile][k_tile]; see DESIGN.md "Packed layout"); -1 if unsupported. */
int64_t q100_packed_words(int64_t rows, int64_t cols, int n_bits);

/*
 * Expand packed codes back to a dense tensor in `dtype`:
 *   out[r, c] = ((code - zero) * scale) with the reference's per-op rounding
 *   (quant/quantizer.py:107-110), or the raw integer code when scales == NULL.
 * scales/zeros use the reference layout [rows*cols/group].
 */
int q100_dequantize(const void* packed, const void* scales, const void* zeros, int dtype,
                    int64_t rows, int64_t cols, int group, int n_bits, int disable_zp,
                    void* out, void* stream);

/*
 * K3 -- fused dequant-GEMM  y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias[N]).
 * Replaces F.linear on the fp16 fake-quantised weight, quant/int_linear.py:62, reading
 * n_bits/8 bytes per weight instead of 2.
 *
 *   x, y, bias  `dtype` (fp16 or bf16), row-major, bias nullable
 *   packed      uint32 [q100_packed_words(N, K, n_bits)]   (from q100_quantize_groupwise)
 *   meta        [K/group, N, 2] (scale, zero) pairs in `dtype` (meta_out above)
 *   group       quantisation group along K (multiple of 64; 0 = K)
 *   workspace   device scratch of at least q100_gemm_workspace_bytes(...) bytes, zero
 *               filled once by the caller before first use (the kernels leave it
 *               zeroed); may be NULL when that function returns 0
 */
int q100_gemm_wna16(const void* x, const void* packed, const void* meta, const void* bias,
                    void* y, int64_t M, int64_t N, int64_t K, int group, int n_bits,
                    int disable_zp, int dtype, void* workspace, size_t workspace_bytes,
                    int flags, void* stream);

size_t q100_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int n_bits, int flags);

/*
 * K3r with the glue of a decoder block folded in (decode shapes; batch <= 2): the launches of
 *   input_layernorm / post_attention_layernorm   (models/int_llama_layer.py:228, 249; quant/omni_norm.py:52-63)
 *   hidden_states = residual + hidden_states     (models/int_llama_layer.py:246, 251)
 * disappear into the GEMMs on either side of them:
 *   norm_weight != NULL  x is the residual stream h [M, K]; the kernel stages weight * (h * rsqrt(mean(h^2) + eps))
 *                        -- K4's arithmetic (q100_add_rmsnorm), up to the order of the fp32 sum of squares -- instead
 *                        of x.  Excludes Q100_GEMM_X_GATE_UP.
 *   residual != NULL     y = r(residual + r(x W^T + bias)), every rounding to `dtype` as on 16-bit tensors; residual
 *                        [M, N] may be the same buffer as y.
 * flags: Q100_GEMM_X_GATE_UP, Q100_GEMM_PDL.  Needs no workspace.  Only for the cases the row-owner kernel takes
 * (q100_gemm_block_fusable(...) != 0: W4, group 128, M <= 2, K %% 2048 == 0, N %% 8 == 0); Q100_EUNSUPPORTED otherwise
 * -- the caller then runs q100_add_rmsnorm + q100_gemm_wna16.
 */
int q100_gemm_wna16_block(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                          int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                          const void* norm_weight, float eps, const void* residual, int flags, void* stream);
int q100_gemm_block_fusable(int64_t M, int64_t N, int64_t K, int group, int n_bits);

/*
 * K3c -- W8A8 / W4A4 (SmoothQuant-style configurations, main.py:326-349): per-token activation quantisation FUSED into
 * the GEMM as its prologue and an integer contraction on the tensor cores, at decode shapes (M <= 15).
 * Replaces  act_quantizer(x)  +  F.linear(x_q, w_q)  of quant/int_linear.py:59-62.
 *
 *   row_minmax  [M, 2] (min, max) of every token row over the FULL K in `dtype` (q100_rows_minmax; MIN / MAX
 *               all-reduced under tensor parallelism): the statistics of quantizer.py:140-141
 *   a_bits      activation bits (2..8): codes in [0, 2^a_bits - 1], asymmetric, per token
 * While it stages x the kernel computes scale / zero-point per token with the reference's per-op rounding and the
 * codes q_x = clamp(round_ste(x / s_x) + z_x) -- bit-identical to the reference's -- and feeds them as unsigned bytes to
 * mma.sync.m16n8k32.u8.u8 against the weight codes (W8: the packed bytes; W4: nibbles).  Per 128-k block and
 * (row n, token m):  y += s_w s_x (sum q_w q_x - z_x sum q_w - z_w sum q_x + 128 z_w z_x), the integer sums exact,
 * the scaling in fp32.  The reference multiplies the two fake-quantised 16-bit operands instead, i.e. it rounds
 * (q - z) s to 16 bits on both sides first: the results agree within the stated GEMM tolerance (tests/test_w8a8.py),
 * not bit for bit (hard part 6).
 * Weights: W4 / W8, group a multiple of 128 (or 0 = per channel).  flags: Q100_FLAG_CUDA_SCALE, Q100_GEMM_PDL.
 */
int q100_gemm_wna8(const void* x, const void* row_minmax, const void* packed, const void* meta, const void* bias,
                   void* y, int64_t M, int64_t N, int64_t K, int group, int n_bits, int a_bits, int disable_zp,
                   int dtype, void* workspace, size_t workspace_bytes, int flags, void* stream);

/*
 * K3d -- W8A8 at prefill sizes (SmoothQuant configuration of main.py:326-349: per-channel 8-bit weights, per-token 8-bit
 * activations) as ONE integer GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in
 * tensor memory).  Replaces  act_quantizer(x)  +  F.linear(x_q, w_q)  of quant/int_linear.py:59-62 for M > 15 rows.
 *
 *   x_packed, w_packed  the 8-bit codes of the activations [M, K] and of the weight [N, K] in the q100 packed
 *                       layout (q100_quantize_groupwise / q100_quantize_rows_stats with group = 0, n_bits = 8,
 *                       `packed` output): both operands are read by TMA exactly as the quantize kernels wrote them
 *   x_meta, w_meta      [M][2] / [N][2] (scale, zero-point) in `dtype`: the `meta` output of the same calls
 *   x_sums, w_sums      [M] / [N] int32: sum over k of each row's codes (q100_code_rowsums; the weight's once per model)
 *   y[m, n] = s_x[m] s_w[n] (sum_k q_x q_w - z_x[m] w_sums[n] - z_w[n] (x_sums[m] - K z_x[m])) (+ bias[n]),
 * the integer part exact (64-bit), one fp32 product and one rounding to `dtype` per output.  The reference multiplies
 * the two fake-quantised 16-bit operands instead, i.e. it rounds (q - z) s to 16 bits on both sides first: same
 * statement as K3c -- within the stated GEMM tolerance (tests/test_w8a8.py), not bit for bit.
 * K % 128 == 0, K <= 32768 (int32 accumulator), N % 8 == 0.  Under tensor parallelism the sums are over the local K.
 *
 * q100_code_rowsums: sums_out[r] = sum_k code[r, k] for a packed 8-bit [rows, cols] matrix (cols % 64 == 0).
 */
int q100_gemm_w8a8(const void* x_packed, const void* x_meta, const void* x_sums, const void* w_packed,
                   const void* w_meta, const void* w_sums, const void* bias, void* y, int64_t M, int64_t N,
                   int64_t K, int dtype, void* stream);
int q100_code_rowsums(const void* packed, int64_t rows, int64_t cols, void* sums_out, void* stream);

/*
 * K4-K6 -- the memory-bound steps between the quantised linears of a decoder block at decode shapes
 * (QuantLlamaDecoderLayer.forward, models/int_llama_layer.py:213-262): ~35 eager launches per block and
 * token in the reference, four here.  fp16 / bf16 only; each reproduces the reference's op order and
 * per-op rounding (fp32 where torch promotes, one rounding where torch stores).
 *
 * q100_add_rmsnorm: h = residual + x (rounded to dtype; skipped when residual == NULL, then h = x),
 *   norm_out = weight * (h * rsqrt(mean(h^2) + eps))        quant/omni_norm.py:52-63 (OmniLlamaRMSNorm),
 *   residual adds of int_llama_layer.py:246, 251.  sum_out (nullable) receives h.  [rows, cols],
 *   cols % 8 == 0, cols <= 8192.
 */
int q100_add_rmsnorm(const void* x, const void* residual, const void* weight, float eps,
                     void* sum_out, void* norm_out, int64_t rows, int64_t cols, int dtype, void* stream);
/*
 * q100_rope_kv_decode: one new token per sequence.  qkv [batch, (heads + 2*kv_heads) * head_dim] is the
 *   fused q|k|v projection; cos_cache / sin_cache [>= position+1, head_dim] in dtype (transformers 4.37
 *   LlamaRotaryEmbedding); position: DEVICE pointer to one int64 (so a captured graph can advance it).
 *   q_out [batch, heads, head_dim] = rotary(q) computed in fp32 on the fp32-cast query
 *   (int_llama_layer.py:117, 125) and rounded once; rotary(k) with every op rounded to dtype and v are
 *   written to k_cache / v_cache [batch, kv_heads, max_len, head_dim] at `position`
 *   (the static-cache form of int_llama_layer.py:130-133).
 *   rope_len = rows of cos_cache / sin_cache.  A position outside [0, min(max_len, rope_len)) writes
 *   nothing and sets *err_flag (device uint32, nullable) to 1 -- the reference's torch.cat cache cannot
 *   overflow, a static one can.
 */
int q100_rope_kv_decode(const void* qkv, const void* cos_cache, const void* sin_cache,
                        const void* position, void* q_out, void* k_cache, void* v_cache,
                        int batch, int heads, int kv_heads, int head_dim, int64_t max_len,
                        int64_t rope_len, void* err_flag, int dtype, void* stream);
/*
 * K8 -- q100_attn_decode: the whole attention core of a decode step in one launch (decode
 *   form): rotary on q / k of the new token, cache append (as q100_rope_kv_decode above), then
 *   softmax(q k^T / sqrt(head_dim)) v over the cache positions 0..position of each sequence -- the static-cache form
 *   of int_llama_layer.py:114-172 with the causal mask (every slot up to the new token visible, none behind it) --
 *   with scores, softmax and accumulation in fp32 from the 16-bit q, k, v on mma.sync tiles, the probabilities rounded
 *   to `dtype` for the second product like the reference's (int_llama_layer.py:152-153; what
 *   scaled_dot_product_attention computes from the same tensors, up to the order of the fp32 sums).  out [batch, heads * head_dim] is the row o_proj reads
 *   (int_llama_layer.py:170-172).  Flash-decoding: the positions are split over CTAs, the last CTA of a
 *   (sequence, kv head) merges the partials; `workspace` (q100_attn_decode_workspace_bytes; zero-filled once by the
 *   caller, left zeroed) holds the arrival counters and the partials and may be shared by all blocks of a model on
 *   one stream.  head_dim == 128, heads / kv_heads <= 8.  Position errors as q100_rope_kv_decode.
 */
size_t q100_attn_decode_workspace_bytes(int batch, int heads, int kv_heads, int64_t max_len);
int q100_attn_decode(const void* qkv, const void* cos_cache, const void* sin_cache, const void* position,
                     void* k_cache, void* v_cache, void* out, void* workspace, size_t workspace_bytes,
                     int batch, int heads, int kv_heads, int head_dim, int64_t max_len, int64_t rope_len,
                     void* err_flag, int dtype, void* stream);
/*
 * q100_rope_qkv_prefill: the same step for q_len new tokens per sequence (prefill / calibration forward):
 *   qkv [batch, q_len, (heads + 2*kv_heads) * head_dim]; position_ids int64 on the device, element
 *   [b * pos_batch_stride + t] (stride 0 = one row shared by the batch, the reference's [1, q_len] form);
 *   q_out [batch, heads, q_len, head_dim], k_out / v_out [batch, kv_heads, q_len, head_dim] -- the
 *   transposed layouts of int_llama_layer.py:114-116 that the attention core and the tuple cache
 *   (:130-135) consume, rotary applied to q (fp32, one rounding) and k (every op rounded) as above,
 *   int_llama_layer.py:117-127.  head_dim % 16 == 0.  A position outside [0, rope_len) leaves that token's
 *   outputs unwritten and sets *err_flag (nullable).
 */
int q100_rope_qkv_prefill(const void* qkv, const void* cos_cache, const void* sin_cache,
                          const void* position_ids, int64_t pos_batch_stride, void* q_out, void* k_out,
                          void* v_out, int batch, int64_t q_len, int heads, int kv_heads, int head_dim,
                          int64_t rope_len, void* err_flag, int dtype, void* stream);
/*
 * q100_silu_mul: out[rows, inter] = silu(gate) * up for gate_up [rows, 2*inter] (gate columns first), the
 *   epilogue of the fused gate|up GEMM: act_fn(gate_proj(x)) * up_proj(x), int_llama_layer.py:45.
 */
int q100_silu_mul(const void* gate_up, void* out, int64_t rows, int64_t inter, int dtype, void* stream);
/*
 * K10 -- q100_argmax_rows: out_index[r] (int64) = argmax over the last dim of x [rows, cols] -- the greedy token choice
 *   after lm_head in the reference's generation loop (llama.py:409-465).  Ties go to the lowest index, NaN counts as
 *   the maximum (torch.argmax's rule).  64 CTAs per row + the last one merges; `workspace`
 *   (q100_argmax_workspace_bytes, zero-filled once by the caller, left zeroed) holds counters and partials.
 */
size_t q100_argmax_workspace_bytes(int64_t rows);
int q100_argmax_rows(const void* x, int64_t rows, int64_t cols, int dtype, void* out_index, void* workspace,
                     size_t workspace_bytes, void* stream);

/*
 * K9 -- the decode chain: ONE persistent launch that runs a list of phases -- the dequant-GEMMs of one or many
 * decoder blocks at decode shapes (1 <= M <= 16 tokens), with the steps between them folded in -- while the
 * packed-weight stream keeps running across the phase boundaries.  Per block and token the reference issues
 * ~45 eager launches for QuantLlamaDecoderLayer.forward (models/int_llama_layer.py:213-262: input_layernorm,
 * q/k/v_proj, rotary + cache append, softmax(QK^T)V, o_proj, residual add, post_attention_layernorm, gate/up_proj,
 * act_fn(gate) * up, down_proj, residual add); a chain expresses the same sequence as phases:
 *
 *   Q100_PH_GEMM     y = x . dequant(W)^T (+ bias)       quant/int_linear.py:62 on the packed weight
 *       x_mode   Q100_X_PLAIN    x [M, K]
 *                Q100_X_GATE_UP  x [M, 2K] = gate | up, the kernel stages act_fn(gate) * up   (int_llama_layer.py:45)
 *                Q100_X_NORM     x = residual stream h [M, K], the kernel stages norm_weight * (h * rstd)
 *                                (quant/omni_norm.py:52-63) with rstd from ssq_in: [ssq_in_parts][M] fp32 partial
 *                                sums of h^2 left by the phase that produced h
 *                Q100_X_ATTN     x = the attention output [M, heads * 128], merged by the kernel while it stages x from
 *                                the flash-decoding partials a preceding Q100_PH_ATTN left in att_partials (fields
 *                                att_partials / position / max_len / heads of this phase; `x` itself is not read):
 *                                o_proj directly behind the attention, no Q100_PH_COMBINE in between
 *       epi_mode Q100_EPI_STORE  y = result
 *                Q100_EPI_ADD_SSQ y = r(r(result) + residual[M, N]) (int_llama_layer.py:246, 251) and
 *                                ssq_out[q100_chain_ssq_parts(N)][M] receives the partial sums of y^2
 *   Q100_PH_ATTN     rotary on q / k of the new token of each sequence, k / v appended to the static cache at the
 *                    DEVICE-side position, attention of the new token over positions 0..position in fp32 (the
 *                    decode form of int_llama_layer.py:114-172), flash-decoding partials into att_partials
 *   Q100_PH_COMBINE  merges the partials into att_out [batch, heads * head_dim] (for a consumer that is not a chain GEMM)
 *
 * Phases run in order; between two phases every CTA passes a grid barrier, so a phase may read what any earlier phase
 * wrote.  All buffers are the caller's and must stay valid (and at the same addresses) as long as the chain is used.
 * Same weights, same layouts (packed / meta) as q100_gemm_wna16; group %% 128 == 0; one n_bits / dtype per chain.
 */
enum q100_phase_type { Q100_PH_GEMM = 0, Q100_PH_ATTN = 1, Q100_PH_COMBINE = 2 };
enum q100_x_mode { Q100_X_PLAIN = 0, Q100_X_GATE_UP = 1, Q100_X_NORM = 2, Q100_X_ATTN = 3 };
enum q100_epi_mode { Q100_EPI_STORE = 0, Q100_EPI_ADD_SSQ = 1 };
typedef struct q100_phase {
  int type;
  /* Q100_PH_GEMM */
  const void* x;
  const void* packed;
  const void* meta;
  const void* bias;            /* nullable */
  void* y;
  int64_t N, K;
  int x_mode;
  const void* norm_weight;     /* [K], X_NORM */
  const void* ssq_in;          /* fp32 [ssq_in_parts][M], X_NORM */
  int ssq_in_parts;
  float eps;
  int epi_mode;
  const void* residual;        /* [M, N], EPI_ADD_SSQ */
  void* ssq_out;               /* fp32 [q100_chain_ssq_parts(N)][M], EPI_ADD_SSQ */
  /* Q100_PH_ATTN / Q100_PH_COMBINE */
  const void* qkv;             /* [batch, (heads + 2 kv_heads) * head_dim] */
  const void* cos_cache;       /* [rope_len, head_dim] */
  const void* sin_cache;
  const void* position;        /* device int64 */
  void* k_cache;               /* [batch, kv_heads, max_len, head_dim] */
  void* v_cache;
  void* att_partials;          /* fp32 [q100_chain_attn_partial_floats(batch, heads, max_len)] */
  void* att_out;               /* [batch, heads * head_dim] */
  int batch, heads, kv_heads, head_dim;
  int64_t max_len, rope_len;
} q100_phase;

size_t q100_chain_desc_bytes(int n_phases);
int64_t q100_chain_ssq_parts(int64_t N);
int64_t q100_chain_attn_partial_floats(int batch, int heads, int64_t max_len);
/*
 * q100_chain_build: validates the phases (HOST array), plans the stream-K split of every GEMM for the persistent
 *   grid and writes the device-side descriptor into `desc` (>= q100_chain_desc_bytes(n_phases) bytes of device
 *   memory, 16-byte aligned).  *workspace_needed receives the scratch size the chain needs; `workspace` must be
 *   that large, zero-filled once (the kernel leaves it zeroed) and must be the one passed to q100_chain_run.
 *   Synchronises the stream (one-time set-up; not for use inside a graph capture).
 * q100_chain_run: one launch; safe to capture in a CUDA graph.  Same M / group / n_bits / dtype as at build time.
 * q100_chain_error: reads the chain's error word (grid-barrier time-out, position outside the KV cache);
 *   synchronises the stream.
 */
int q100_chain_build(const q100_phase* phases, int n_phases, int M, int group, int n_bits, int disable_zp,
                     int dtype, void* desc, size_t desc_bytes, void* workspace, size_t workspace_bytes,
                     size_t* workspace_needed, void* stream);
int q100_chain_run(void* desc, int n_phases, int M, int group, int n_bits, int disable_zp, int dtype,
                   void* workspace, void* stream);
int q100_chain_error(const void* desc, void* stream);

/*
 * K7 -- one-shot all-reduce(sum) of a small tensor over peer memory (NVLink): the collective after a
 * row-parallel linear (o_proj, down_proj) at decode shapes, where NCCL's all-reduce is latency-bound
 * (the reference has no counterpart -- it places layers, parallel_utils.py:89-163).
 *   x, y        [n] in `dtype` (fp16 / bf16), n % 8 == 0; y may alias x
 *   peer_bufs   DEVICE array [world] of pointers to every rank's staging buffer of 2 * world * slot_elems
 *               elements, mapped into this process (e.g. torch symmetric memory: handle.buffer_ptrs_dev)
 *   peer_flags  DEVICE array [world] of pointers to every rank's 2 * world * 32 uint32 flag words, zeroed
 *               before the first call (all ranks, followed by a barrier)
 *   epoch       this rank's 32 uint32 call counters, zeroed before the first call; advanced by the kernel
 *   err_flag    one uint32, set to 1 if a peer did not answer within 30 s (the call then returns garbage
 *               instead of hanging the GPU)
 * Every rank must issue the same sequence of calls (same n).  Safe to capture in a CUDA graph.
 */
int q100_allreduce_oneshot(const void* x, void* y, int64_t n, int dtype, const void* peer_bufs,
                           const void* peer_flags, void* epoch, int64_t slot_elems, int rank, int world,
                           void* err_flag, void* stream);

/*
 * Row-parallel all-reduce at prefill sizes as GEMM epilogue + one kernel, over peer memory (the
 * reference has no counterpart).  The token rows are cut into `world` blocks; rank r owns block r.
 *
 * q100_gemm_wna16_scatter: the tcgen05 dequant-GEMM of q100_gemm_wna16 (M %% (world * 256) == 0, no bias) whose
 *   epilogue stores each [128 x 256-token] tile over NVLink straight into slot `rank` of the OWNER's staging buffer
 *   (peer_stage: DEVICE array [world] of pointers to every rank's [world][M / world][N] buffer, e.g. torch symmetric
 *   memory) -- the reduce-scatter's transfer overlaps the GEMM's math tile by tile; nothing is written to a local y.
 * q100_reduce_bcast: next in the stream on every rank.  Waits until every peer's GEMM has completed, sums the `world`
 *   slots of the owned rows in rank order in fp32 (all ranks compute identical bits), stores the result into every
 *   rank's y (peer_y: DEVICE array [world] of pointers to every rank's [M, N] buffer) and returns once every peer's
 *   rows have arrived.  peer_flags: DEVICE array [world] of pointers to every rank's 4 * world * 128 uint32 flag
 *   words, zeroed once; epoch: this rank's 128 uint32 counters, zeroed once.  Every rank must issue the same sequence
 *   of calls.  err_flag is set to 1 after a 30 s wait for a peer.  Safe to capture in a CUDA graph.
 */
int q100_gemm_wna16_scatter(const void* x, const void* packed, const void* meta, int64_t M, int64_t N, int64_t K,
                            int group, int n_bits, int disable_zp, int dtype, const void* peer_stage, int rank,
                            int world, void* stream);
int q100_reduce_bcast(const void* stage, const void* peer_y, const void* peer_flags, void* epoch, int64_t rows_per,
                      int64_t N, int dtype, int rank, int world, void* err_flag, void* stream);
/*
 * K7c -- q100_allreduce_multimem: in-place all-reduce(sum) of a [world * rows_per, N] tensor that every rank holds in
 * a symmetric buffer with a MULTICAST mapping (NVSwitch): rank r reduces its rows through the switch
 * (multimem.ld_reduce, fp32 accumulation) and writes them into every rank's copy (multimem.st).  `mc_buf` is the
 * multicast address of the buffer (torch symmetric memory: handle.multicast_ptr), `peer_flags` / `epoch` / `err_flag`
 * as q100_reduce_bcast.  No counterpart in the reference (parallel_utils.py only places layers); the all-reduce
 * after the row-parallel o_proj / down_proj of a tensor-parallel block at prefill sizes.
 */
int q100_allreduce_multimem(void* mc_buf, const void* peer_flags, void* epoch, int64_t rows_per, int64_t N,
                            int dtype, int rank, int world, void* err_flag, void* stream);
/*
 * K7n -- q100_allreduce_add_rmsnorm: the all-reduce after a row-parallel linear of a tensor-parallel DECODE step fused
 * with the residual add + RMSNorm that consumes it (q100_allreduce_oneshot + q100_add_rmsnorm in one launch):
 *   y = sum over ranks of x (fp32 in rank order, one rounding to dtype -- the bits q100_allreduce_oneshot produces),
 *   h = residual + y (rounded; h = y when residual == NULL), norm_out = weight * (h * rsqrt(mean(h^2) + eps)),
 * i.e. int_llama_layer.py:246 / 251 + quant/omni_norm.py:52-63 on the all-reduced output.  sum_out (nullable) receives h.
 * One CTA per token row pushes its partial row to every rank over peer memory and waits for the others' rows.
 * peer_bufs: device array of `world` pointers to staging buffers of 2 * world * 64 * cols elements; peer_flags: `world`
 * pointers to 2 * world * 64 zero-initialised words; epoch: 64 zero-initialised words (advanced by the kernel, so the
 * launch is graph-capturable).  rows <= 64, cols % 8 == 0, cols <= 8192, the same `cols` in every call on one buffer set.
 */
int q100_allreduce_add_rmsnorm(const void* x, const void* residual, const void* weight, float eps, void* sum_out,
                               void* norm_out, int64_t rows, int64_t cols, int dtype, const void* peer_bufs,
                               const void* peer_flags, void* epoch, int rank, int world, void* err_flag, void* stream);

/*
 * Development / verification hooks (used by tests only).
 * q100_selftest_div: exhaustively compares the fast 16-bit division sequence with
 * div.rn.f32 + round over all (x, scale) pairs of `dtype` with scale in [1e-5, 1e4];
 * writes the mismatch count to *mismatches_out (device uint64).
 */
int q100_selftest_div(int dtype, void* mismatches_out, void* stream);
/*
 * q100_debug_gemm_dense_tc: y[M, N] = x[M, K] . w[N, K]^T for DENSE fp16/bf16 w through the
 * tcgen05 kernel's TMA / UMMA-descriptor / TMEM plumbing (B tile loaded by TMA instead of
 * being dequantised).  Regression hook for the tensor-core path; not a product entry point.
 */
int q100_debug_gemm_dense_tc(const void* x, const void* w, void* y, int64_t M, int64_t N, int64_t K,
                             int dtype, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* Q100_H_ */
