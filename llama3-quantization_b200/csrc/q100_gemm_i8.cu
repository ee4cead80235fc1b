// q100_gemm_i8.cu -- K3d: W8A8 at prefill sizes on the integer tensor cores (tcgen05.mma kind::i8).
//
//   y[M, N] = dequant(Xq)[M, K] . dequant(Wq)[N, K]^T (+ bias)
//
// The SmoothQuant-style configuration of the reference (main.py:326-349: per-channel 8-bit weights, per-token 8-bit
// activations) as ONE integer GEMM: quant/int_linear.py:59-62 runs act_quantizer(x) and F.linear on the two
// fake-quantised 16-bit tensors; here both operands stay 8-bit codes,
//   y[m, n] = s_x[m] s_w[n] ( sum_k q_x q_w  -  z_x[m] sum_k q_w[n, k]  -  z_w[n] sum_k q_x[m, k]  +  K z_x[m] z_w[n] ),
// the contraction exact in int32 on the tensor cores, the zero-point correction exact in 64-bit integer arithmetic in
// the epilogue, one fp32 product per output.  (The reference rounds (q - z) s to 16 bits on both sides
// before its GEMM: the difference is that rounding, SURVEY.md section 7 hard part 6; same statement as K3c.)
//
// Both operands are q100 v2 packed code matrices (q100_quantize_groupwise with group = 0, n_bits = 8): tile-major
// [rows / 128][K / 64][128 rows x 64 bytes], bytes of a word in the order (e0, e2, e1, e3) -- the SAME permutation of k
// on both sides, so the contraction does not see it.  A tile is exactly one K-major operand tile of 64 bytes per row:
// TMA (SWIZZLE_64B) drops it into shared memory in the canonical layout and tcgen05.mma reads it from there.  No
// dequant warps, no 16-bit copy of anything.
//
// Per CTA (one per SM, persistent over [128 weight rows x 256 tokens] tiles, token tiles fastest), stages of 128 k:
//   warp 0      TMA producer: two weight tiles + four activation tiles (48 KB) per stage, four stages;
//   warp 1      allocates tensor memory (2 x 256 int32 columns: two accumulator stages); one elected thread issues
//               4 x tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32) per stage, tcgen05.commit frees the stage /
//               publishes the accumulator;
//   warps 4-11  epilogue (two warps per TMEM lane quarter, 128 token columns each): tcgen05.ld (lane = weight row,
//               column = token) -> integer correction -> scale -> (+ bias) -> 16 bit -> transposed through shared
//               memory -> 16-byte stores; overlaps the next tile's MMAs (with four warps the epilogue of a
//               K = 4096 tile took longer than the tile's MMAs).
#include <cstdlib>
#include <mutex>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

namespace {

constexpr int I8_ROWS = 128;          // weight rows per tile (UMMA M)
constexpr int I8_NT = 256;            // tokens per tile (UMMA N)
constexpr int I8_TILE = 128 * 64;     // bytes of one packed [128 x 64] tile
constexpr int I8_STAGE = 6 * I8_TILE; // two weight tiles, 2 x 2 activation tiles
constexpr int I8_STAGES = 4;
constexpr int I8_THREADS = 12 * 32;
constexpr int I8_EPI = 8 * 2048;      // 8 warps x [32 tokens x 32 rows] 16-bit
constexpr int I8_COLP = 2 * I8_NT * 16;   // per accumulator stage: (s_x, z_x, sum q_x, K z_x) per token
constexpr int I8_SMEM = I8_STAGES * I8_STAGE + I8_EPI + I8_COLP + 1024;
static_assert(I8_SMEM <= 227 * 1024, "shared memory budget");

struct I8Args {
  const void* w_meta;        // [N][2] T (scale, zero): per-channel weights
  const int32_t* w_sum;      // [N] sum_k q_w
  const void* x_meta;        // [M][2] T (scale, zero): per-token activations
  const int32_t* x_sum;      // [M] sum_k q_x
  const void* bias;          // [N] T or null
  void* y;                   // [M, N] T
  int M, N, K;
  int m_tiles, n_tiles;      // tiles of 256 tokens / 128 rows
  int kt;                    // K / 64
  int mt128;                 // ceil(M / 128): token tiles of the packed x
};

// K-major operand tile, SWIZZLE_64B: rows of 64 bytes, groups of 8 rows 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64_u32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);            // start address        [0,14)
  d |= static_cast<uint64_t>(1) << 16;                           // leading byte offset  [16,30) (ignored: K extent <= swizzle width)
  d |= static_cast<uint64_t>(512 >> 4) << 32;                    // stride byte offset   [32,46): 8 rows * 64 B
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version   [46,48) = 1
  d |= static_cast<uint64_t>(4) << 61;                           // layout type          [61,64) = SWIZZLE_64B
  return d;
}
// kind::i8, unsigned 8-bit A and B (K-major), int32 accumulate (cute::UMMA::InstrDescriptor)
__device__ __host__ constexpr uint32_t umma_idesc_u8(int M, int N) {
  return (2u << 4)                                  // c_format = S32
         | (0u << 7) | (0u << 10)                   // a_format, b_format = unsigned 8 bit
         | (static_cast<uint32_t>(N >> 3) << 17)
         | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void i8_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void i8_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// CL = cluster size: the CL CTAs of a cluster work on CL consecutive weight-row tiles of the SAME 256-token tile; each
// loads 1/CL of the stage's activation tiles and multicasts it to all of them (the L2 -> SM traffic of one CTA per
// stage drops from 48 KB to 16 + 32/CL KB: at the integer rate the single-CTA version asks L2 for more than it delivers).
template <typename T, int CL>
__global__ void __launch_bounds__(I8_THREADS, 1)
k3_i8(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const I8Args a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base_u32 - smem_u32(smem_raw));
  __shared__ uint64_t bars[2 * I8_STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full = smem_u32(&bars[0]), empty = full + 8 * I8_STAGES, d_full = empty + 8 * I8_STAGES, d_empty = d_full + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int n_groups = (a.n_tiles + CL - 1) / CL;
  const int n_ct = a.m_tiles * n_groups;
  const int ct0 = blockIdx.x / CL, ct_step = gridDim.x / CL;
  const int num_ks = a.K / 128;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (tid == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[I8_STAGES + s], CL); }
    for (int d = 0; d < 2; ++d) { mbar_init(&bars[2 * I8_STAGES + d], 1); mbar_init(&bars[2 * I8_STAGES + 2 + d], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();                    // every CTA's barriers exist before any remote arrive
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tm_w);
      tma_prefetch_desc(&tm_x);
      const uint64_t pol_w = policy_evict_first(), pol_x = policy_evict_last();
      uint32_t it = 0;
      for (int ct = ct0; ct < n_ct; ct += ct_step) {
        const int mt = ct % a.m_tiles;
        const int nt = min((ct / a.m_tiles) * CL + rank, a.n_tiles - 1);                // ragged last group: redo the last tile
        const int xt0 = min(2 * mt, a.mt128 - 1), xt1 = min(2 * mt + 1, a.mt128 - 1);   // a missing second half: redo the first
        for (int ks = 0; ks < num_ks; ++ks, ++it) {
          const uint32_t s = it % I8_STAGES;
          i8_wait(empty + 8 * s, ((it / I8_STAGES) & 1) ^ 1);   // all CL CTAs have retired their MMAs on this stage
          mbar_expect_tx(&bars[s], I8_STAGE);                   // own weight tiles + the whole activation part
          uint8_t* st = base_ptr + s * I8_STAGE;
          const uint32_t st32 = base_u32 + s * I8_STAGE;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int kt = 2 * ks + kk;
            tma_load_2d_hint(st + kk * I8_TILE, &tm_w, &bars[s], 0, (nt * a.kt + kt) * 128, pol_w);
            if (CL == 1) {
              tma_load_2d_hint(st + (2 + 2 * kk) * I8_TILE, &tm_x, &bars[s], 0, (xt0 * a.kt + kt) * 128, pol_x);
              tma_load_2d_hint(st + (3 + 2 * kk) * I8_TILE, &tm_x, &bars[s], 0, (xt1 * a.kt + kt) * 128, pol_x);
            } else if (CL == 2) {                               // rank = token half, both k tiles
              tma_load_2d_mc(st32 + (2 + 2 * kk + rank) * I8_TILE, &tm_x, full + 8 * s, 0,
                             ((rank ? xt1 : xt0) * a.kt + kt) * 128, kMask, pol_x);
            } else if (kk == (rank >> 1)) {                     // CL == 4: rank = (k tile, token half)
              tma_load_2d_mc(st32 + (2 + 2 * kk + (rank & 1)) * I8_TILE, &tm_x, full + 8 * s, 0,
                             (((rank & 1) ? xt1 : xt0) * a.kt + kt) * 128, kMask, pol_x);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_u8(I8_ROWS, I8_NT);
    uint32_t it = 0, tcount = 0;
    for (int ct = ct0; ct < n_ct; ct += ct_step, ++tcount) {
      const uint32_t ds = tcount & 1;
      i8_wait(d_empty + 8 * ds, ((tcount >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator stage
      tc_fence_after();
      for (int ks = 0; ks < num_ks; ++ks, ++it) {
        const uint32_t s = it % I8_STAGES;
        i8_wait(full + 8 * s, (it / I8_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = base_u32 + s * I8_STAGE;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t ad = umma_desc_sw64_u32(st + kk * I8_TILE);
            const uint64_t bd = umma_desc_sw64_u32(st + (2 + 2 * kk) * I8_TILE);      // 256 token rows: two tiles back to back
#pragma unroll
            for (int h = 0; h < 2; ++h)                                                // K = 32 bytes per instruction
              umma_i8_ss(tmem + ds * I8_NT, ad + 2 * h, bd + 2 * h, idesc, (ks != 0 || kk != 0 || h != 0) ? 1u : 0u);
          }
          if (CL == 1) umma_commit_u32(empty + 8 * s);
          else umma_commit_mc(empty + 8 * s, kMask);           // frees the stage in every CTA of the cluster
          if (ks == num_ks - 1) umma_commit_u32(d_full + 8 * ds);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (4..11): lane = weight row, columns = tokens =====================
    // two warps per TMEM lane quarter, each takes 128 of the tile's 256 token columns
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int et = tid - 128;                                           // 0..255 among the epilogue threads
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t tcount = 0;
    for (int ct = ct0; ct < n_ct; ct += ct_step, ++tcount) {
      const uint32_t ds = tcount & 1;
      const int mt = ct % a.m_tiles, nt = (ct / a.m_tiles) * CL + rank;   // >= n_tiles in a ragged last group: nothing to store
      const int m0 = mt * I8_NT;
      const int n = nt * I8_ROWS + quad * 32 + lane;
      const bool n_ok = n < a.N;
      // this lane's weight row: scale, zero-point (an integer-valued float), sum of its codes
      float sw = 0.f;
      int zw = 0, Sw = 0;
      float bv = 0.f;
      if (n_ok) {
        const float2 m = DT<T>::to_f2(reinterpret_cast<const uint32_t*>(a.w_meta)[n]);
        sw = m.x; zw = __float2int_rn(m.y); Sw = a.w_sum[n];
        if (bias != nullptr) bv = DT<T>::to_f(bias[n]);
      }
      const bool w_wide = static_cast<uint32_t>(zw) > 255u;
      // the tile's tokens: (s_x, z_x, sum q_x - K z_x, zero-point outside [0, 255]) per column
      uint4* colp = reinterpret_cast<uint4*>(base_ptr + I8_STAGES * I8_STAGE + I8_EPI + ds * (I8_NT * 16));
      {
        const int m = m0 + et;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (m < a.M) {
          const float2 xm = DT<T>::to_f2(reinterpret_cast<const uint32_t*>(a.x_meta)[m]);
          const int zx = __float2int_rn(xm.y);
          // c1 = sum q_x - K z_x fits 32 bits (|z_x| <= 1e4 by the reference's clamp, K <= 32768)
          v = make_uint4(__float_as_uint(xm.x), static_cast<uint32_t>(zx),
                         static_cast<uint32_t>(a.x_sum[m] - a.K * zx), static_cast<uint32_t>(zx) > 255u ? 1u : 0u);
        }
        colp[et] = v;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");                    // the eight epilogue warps
      i8_wait(d_full + 8 * ds, (tcount >> 1) & 1);
      tc_fence_after();
      T* ydst = reinterpret_cast<T*>(a.y);
      const int n_warp = nt * I8_ROWS + quad * 32;                      // first row of this warp
      const uint32_t eb = base_u32 + I8_STAGES * I8_STAGE + (warp - 4) * 2048;
      const int tj = lane >> 2, rq = (lane & 3) * 8;                    // read side: token tj (+8i), rows rq..rq+7
      const int c_lo = half * (I8_NT / 2), c_hi = min(c_lo + I8_NT / 2, a.M - m0);
      if (c_lo >= c_hi) {                            // no token of this half exists: only hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) i8_arrive(d_empty + 8 * ds);
      }
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_t + ds * I8_NT + c0, r);
        tmem_ld_wait();
        if (c0 + 32 >= c_hi) {                       // last chunk read: the accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) i8_arrive(d_empty + 8 * ds);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint4 cp = colp[c0 + j];                                // broadcast read
          // sum (q_x - z_x)(q_w - z_w) = D - z_x Sw - z_w (Sx - K z_x).  Zero-points in [0, 255] (a row with both signs):
          // every term and the result fit 32 bits; otherwise (a row of one sign, zero-point up to the reference's
          // clamp at 1e4) the same expression in 64-bit integers
          float ti;
          if (!(w_wide || cp.w)) {
            ti = static_cast<float>(static_cast<int32_t>(r[j]) - static_cast<int32_t>(cp.y) * Sw -
                                    zw * static_cast<int32_t>(cp.z));
          } else {
            ti = __ll2float_rn(static_cast<long long>(static_cast<int32_t>(r[j])) -
                               static_cast<long long>(static_cast<int32_t>(cp.y)) * Sw -
                               static_cast<long long>(zw) * static_cast<int32_t>(cp.z));
          }
          const float f = fmaf(ti, __uint_as_float(cp.x) * sw, bv);
          const T v = DT<T>::from_f(f);
          asm volatile("st.shared.u16 [%0], %1;" :: "r"(eb + j * 64 + lane * 2), "h"(*reinterpret_cast<const uint16_t*>(&v)) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = m0 + c0 + tj + 8 * i;
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(eb + (tj + 8 * i) * 64 + rq * 2));
          if (m < a.M && n_warp + rq < a.N)
            *reinterpret_cast<uint4*>(ydst + static_cast<int64_t>(m) * a.N + n_warp + rq) = v;
        }
        __syncwarp();                                                   // the buffer is rewritten by the next chunk
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");                    // colp[ds] is rewritten two tiles later; keep the warps together
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                    // nobody leaves while a peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// sum of the 8-bit codes of every row of a packed [rows, K] matrix (tile-major v2): out[r] += ..., out zeroed by the caller
__global__ void __launch_bounds__(128) k_code_rowsums(const uint8_t* __restrict__ packed, int rows, int kt,
                                                      int kt_per_block, int32_t* __restrict__ out) {
  const int tile = blockIdx.x, r = tile * 128 + threadIdx.x;
  const int k0 = blockIdx.y * kt_per_block, k1 = min(kt, k0 + kt_per_block);
  uint32_t acc = 0;
  for (int k = k0; k < k1; ++k) {
    const uint4* p = reinterpret_cast<const uint4*>(packed + (static_cast<int64_t>(tile) * kt + k) * I8_TILE + threadIdx.x * 64);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 v = p[i];
      acc = __dp4a(v.x, 0x01010101u, acc); acc = __dp4a(v.y, 0x01010101u, acc);
      acc = __dp4a(v.z, 0x01010101u, acc); acc = __dp4a(v.w, 0x01010101u, acc);
    }
  }
  if (r < rows) atomicAdd(out + r, static_cast<int32_t>(acc));
}

template <typename T, int CL>
int launch_i8(const CUtensorMap& tw, const CUtensorMap& tx, const I8Args& a, cudaStream_t st) {
  auto kern = k3_i8<T, CL>;
  static std::mutex mu;                          // one per instantiation
  static bool attr_done[kMaxDevices] = {};       // the attribute is a per-device property
  static int max_clusters[kMaxDevices] = {};
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_done[dev]) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
      attr_done[dev] = true;
    }
  }
  const int n_ct = a.m_tiles * ((a.n_tiles + CL - 1) / CL);
  int clusters = sm_count() / CL;
  if (clusters > n_ct) clusters = n_ct;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(I8_THREADS);
  cfg.dynamicSmemBytes = I8_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  if (CL > 1) {
    // persistent kernel: never more clusters than can be resident at once
    std::lock_guard<std::mutex> lk(mu);
    if (max_clusters[dev] == 0) {
      cfg.gridDim = dim3(sm_count() / CL * CL);
      Q100_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters[dev], kern, &cfg));
      if (max_clusters[dev] < 1) max_clusters[dev] = 1;
    }
    if (clusters > max_clusters[dev]) clusters = max_clusters[dev];
    cfg.gridDim = dim3(clusters * CL);
  }
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tw, tx, a), "k3_i8 launch");
}

template <typename T>
int launch_i8_cl(const CUtensorMap& tw, const CUtensorMap& tx, const I8Args& a, cudaStream_t st) {
  static const int env_cl = []() { const char* e = getenv("Q100_I8_CLUSTER"); return e ? atoi(e) : 0; }();   // experiments
  int cl = env_cl ? env_cl : 2;
  if (a.n_tiles < cl) cl = 1;
  if (cl == 4) return launch_i8<T, 4>(tw, tx, a, st);
  if (cl == 2) return launch_i8<T, 2>(tw, tx, a, st);
  return launch_i8<T, 1>(tw, tx, a, st);
}

}  // namespace

}  // namespace q100

using namespace q100;

extern "C" int q100_code_rowsums(const void* packed, int64_t rows, int64_t cols, void* sums_out, void* stream) {
  Q100_REQUIRE(packed && sums_out, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(rows > 0 && cols > 0 && cols % 64 == 0 && rows < (1ll << 31), Q100_EINVAL, "bad shape rows=%lld cols=%lld",
               (long long)rows, (long long)cols);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Q100_CUDA(cudaMemsetAsync(sums_out, 0, static_cast<size_t>(rows) * 4, st));
  const int tiles = static_cast<int>((rows + 127) / 128), kt = static_cast<int>(cols / 64);
  int split = (4 * sm_count() + tiles - 1) / tiles;
  if (split > kt) split = kt;
  if (split < 1) split = 1;
  const int per = (kt + split - 1) / split;
  k_code_rowsums<<<dim3(tiles, (kt + per - 1) / per), 128, 0, st>>>(reinterpret_cast<const uint8_t*>(packed),
                                                                      static_cast<int>(rows), kt, per,
                                                                      reinterpret_cast<int32_t*>(sums_out));
  return check_cuda(cudaGetLastError(), "k_code_rowsums launch");
}

extern "C" int q100_gemm_w8a8(const void* x_packed, const void* x_meta, const void* x_sums, const void* w_packed,
                              const void* w_meta, const void* w_sums, const void* bias, void* y, int64_t M, int64_t N,
                              int64_t K, int dtype, void* stream) {
  Q100_REQUIRE(x_packed && x_meta && x_sums && w_packed && w_meta && w_sums && y, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(M >= 0 && N > 0 && K > 0, Q100_EINVAL, "bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  Q100_REQUIRE(K % 128 == 0 && N % 8 == 0 && K <= 32768, Q100_EUNSUPPORTED,
               "q100_gemm_w8a8: K %% 128 == 0 (K <= 32768: the int32 accumulator), N %% 8 == 0; got N=%lld K=%lld",
               (long long)N, (long long)K);
  if (M == 0) return Q100_OK;
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  I8Args a{};
  a.w_meta = w_meta; a.w_sum = reinterpret_cast<const int32_t*>(w_sums);
  a.x_meta = x_meta; a.x_sum = reinterpret_cast<const int32_t*>(x_sums);
  a.bias = bias; a.y = y;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K);
  a.m_tiles = static_cast<int>((M + I8_NT - 1) / I8_NT);
  a.n_tiles = static_cast<int>((N + I8_ROWS - 1) / I8_ROWS);
  a.kt = static_cast<int>(K / 64);
  a.mt128 = static_cast<int>((M + 127) / 128);
  CUtensorMap tw, tx;
  // a packed code matrix is a row-major array of 64-byte rows: [row tiles][k tiles][128 rows]
  const uint64_t w_rows = static_cast<uint64_t>(a.n_tiles) * a.kt * 128, x_rows = static_cast<uint64_t>(a.mt128) * a.kt * 128;
  rc = make_tmap_2d(&tw, w_packed, 1, 64, w_rows, 64, 64, 128, 64);
  if (rc != Q100_OK) return rc;
  rc = make_tmap_2d(&tx, x_packed, 1, 64, x_rows, 64, 64, 128, 64);
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == Q100_F16) return launch_i8_cl<__half>(tw, tx, a, st);
  return launch_i8_cl<__nv_bfloat16>(tw, tx, a, st);
}
