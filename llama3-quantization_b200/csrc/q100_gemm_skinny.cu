// q100_gemm_skinny.cu -- K3a: weight-streaming dequant-GEMM for decode shapes (M <= 32).
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// HBM-bound: the packed weights (n_bits/8 B per weight) are the only large stream; they
// are read exactly once.  Structure (one persistent CTA per SM slot, 8 consumer warps +
// 1 producer warp):
//   * producer: TMA bulk copies (cp.async.bulk) of whole [128 rows x 64 k] packed tiles -- the
//     tile-major layout makes every pipeline unit one contiguous run of HBM -- plus 1-D bulk copies
//     of the transposed (scale, zero) side band and cp.async of the x slab -> 4-stage smem ring,
//     mbarrier full/empty pairs;
//   * consumers: each warp owns 16 output rows; packed codes -> registers, dequantised
//     in registers with the reference's exact per-op rounding
//     (w = r(r(q - zp) * scale), bit-identical to the fake-quantised weight) and fed to
//     mma.sync.m16n8k16 as the A operand, x^T as B, fp32 accumulators.
//     tcgen05 is deliberately not used here: at M <= 32 the operand would have to be
//     written back to shared memory (2 B/weight) and that traffic alone exceeds the time
//     budget of the 0.5 B/weight HBM stream (DESIGN.md "K3a").
//   * work is split stream-K style over 148*k CTAs in units of [128 rows x 256 k]; a tile
//     finished by several CTAs is reduced by the last arriver from fp32 partials in a
//     caller-provided workspace, in fixed slot order (deterministic).
//
// Lane <-> k mapping: inside a 128-k block lane t (= lane & 3) owns k in [32t, 32t+32) for
// both operands, i.e. the mma "slot" (s, 2t+{0,1,8,9}) of k16-step s carries
// k = 32t + 4s + {0,1,2,3}.  The contraction is invariant under that permutation because
// A and B fragments use the same one.
#include <mutex>
#include <unordered_map>

#include "q100_common.cuh"
#include "q100_ptx.cuh"

namespace q100 {

constexpr int SK_ROWS = 128;       // output rows per tile (8 warps x 16)
constexpr int SK_KB = 128;         // k per block (one TMA box)
constexpr int SK_KBU = 2;          // k blocks per pipeline unit
constexpr int SK_STAGES = 4;
constexpr int SK_CONSUMERS = 256;
constexpr int SK_THREADS = SK_CONSUMERS + 32;
constexpr int SK_XPITCH = 320;     // bytes per (m, 128-k) x row in smem: 256 + 64 pad (conflict-free LDS.128)

struct SkinnyArgs {
  const void* x;          // [M, K] T
  const void* packed;     // tile-major packed codes (layout v2)
  const void* meta;       // [K/group, N, 2] T
  const void* bias;       // [N] T or null
  void* y;                // [M, N] T
  float* ws_partials;     // workspace: fp32 partial tiles
  int* ws_counters;       // workspace: one arrival counter per tile (zero between launches)
  int M, N, K, group;
  int kblocks;            // K / 128
  int kbu;                // k blocks per unit (1 or 2)
  int units_per_tile;
  int n_tiles;
  int total_units;
  int max_slots;
  int has_zp;
  float qmin;
  int meta_stage;         // bytes of (scale, zero) side band per stage, multiple of 128
  int post;               // 1: magic-form operand + fp32 group scaling (group % 128 == 0)
  int stage_bytes;        // W + side band + x slab, multiple of 128
};

// Packed layout v2: tiles of [128 rows x 64 k] are contiguous in HBM; a 128-k block is two
// consecutive k-tiles.  Each k-tile is bulk-copied to its own padded slot in shared memory so that
// the two k-tiles of a block fall on different banks (conflict-free LDS for all four lanes t).
template <int NBITS> struct WGeom {
  static constexpr int kTileBytes = NBITS == 3 ? 128 * 24 : 128 * 8 * NBITS;   // one [128 x 64] tile
  static constexpr int kPitch = kTileBytes + (NBITS == 8 ? 16 : 64);
  static constexpr int kRowBytes = NBITS == 3 ? 16 : 8 * NBITS;                // plane 0 bytes per row per k-tile
};

template <int NBITS, int MB>
struct SkinnySmem {
  static constexpr int kWBytes = ((SK_KBU * 2 * WGeom<NBITS>::kPitch) + 127) & ~127;
  static constexpr int kXBytes = SK_KBU * MB * 8 * SK_XPITCH;
  __host__ __device__ static constexpr int stage(int meta_stage) { return kWBytes + meta_stage + kXBytes; }
  __host__ __device__ static constexpr int total(int meta_stage) { return SK_STAGES * stage(meta_stage) + 1024; }
};

// ---------------------------------------------------------------- dequant helpers
template <typename T> struct DQ;
template <> struct DQ<__half> {
  using T2 = __half2;
  static constexpr uint32_t kMagic = 0x64006400u;      // 1024.0 in both halves
  static constexpr float kMagicF = 1024.0f;
  __device__ static __forceinline__ T2 magic_val() { return __float2half2_rn(1024.0f); }
};
template <> struct DQ<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr uint32_t kMagic = 0x43004300u;      // 128.0 in both halves (codes < 128)
  static constexpr float kMagicF = 128.0f;
  __device__ static __forceinline__ T2 magic_val() { return __float2bfloat162_rn(128.0f); }
};

// pair: two unsigned codes in the low bits of each 16-bit half.  Returns
// r(r(code + qmin - zp) * scale) per half -- the reference's dequant, quantizer.py:107-110.
template <typename T, int NBITS>
__device__ __forceinline__ uint32_t dq_pair(uint32_t pair, typename DQ<T>::T2 qmin2,
                                            typename DQ<T>::T2 zp2, typename DQ<T>::T2 s2, bool has_zp) {
  using T2 = typename DQ<T>::T2;
  T2 c;
  if (NBITS == 8 && DT<T>::kCode == Q100_BF16) {
    // bf16 has 8 significant bits: the OR-magic trick only covers codes < 128
    const float lo = static_cast<float>(pair & 0xffffu), hi = static_cast<float>(pair >> 16);
    const uint32_t b = DT<T>::pack2(lo, hi);
    c = *reinterpret_cast<const T2*>(&b);
  } else {
    const uint32_t m = pair | DQ<T>::kMagic;
    c = __hsub2(*reinterpret_cast<const T2*>(&m), DQ<T>::magic_val());   // exact
  }
  if (has_zp) c = __hsub2(c, zp2);             // r(q - zp)           (qmin == 0)
  else c = __hadd2(c, qmin2);                  // stored code + qmin  (exact small integers)
  c = __hmul2(c, s2);                          // r(. * scale)
  return *reinterpret_cast<uint32_t*>(&c);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// A lane's 32 codes (k in [32t, 32t+32) of one row's 128-k block) as 16 pairs.  `blk` is the shared
// address of the block's first k-tile; lane t reads k-tile (t >> 1), half (t & 1).
template <int NBITS> struct LaneCodes;
template <> struct LaneCodes<4> {
  uint4 w;
  __device__ __forceinline__ void load(uint32_t blk, int row, int t) {
    w = lds128(blk + (t >> 1) * WGeom<4>::kPitch + row * 32 + (t & 1) * 16);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 2) == 0 ? w.x : (i >> 2) == 1 ? w.y : (i >> 2) == 2 ? w.z : w.w;
    return (x >> (4 * (i & 3))) & 0x000f000fu;
  }
};
template <> struct LaneCodes<2> {
  uint2 w;
  __device__ __forceinline__ void load(uint32_t blk, int row, int t) {
    w = lds64(blk + (t >> 1) * WGeom<2>::kPitch + row * 16 + (t & 1) * 8);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? w.x : w.y;
    return (x >> (2 * (i & 7))) & 0x00030003u;
  }
};
template <> struct LaneCodes<8> {
  uint4 w0, w1;
  __device__ __forceinline__ void load(uint32_t blk, int row, int t) {
    const uint32_t p = blk + (t >> 1) * WGeom<8>::kPitch + row * 64 + (t & 1) * 32;
    w0 = lds128(p);
    w1 = lds128(p + 16);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const int wi = i >> 1;
    const uint32_t x = wi == 0 ? w0.x : wi == 1 ? w0.y : wi == 2 ? w0.z : wi == 3 ? w0.w
                     : wi == 4 ? w1.x : wi == 5 ? w1.y : wi == 6 ? w1.z : w1.w;
    return (x >> (8 * (i & 1))) & 0x00ff00ffu;
  }
};
template <> struct LaneCodes<3> {
  uint2 lo;
  uint32_t hi;
  __device__ __forceinline__ void load(uint32_t blk, int row, int t) {
    const uint32_t p = blk + (t >> 1) * WGeom<3>::kPitch;
    lo = lds64(p + row * 16 + (t & 1) * 8);                  // 2-bit plane: 128 rows x 16 B
    hi = lds32(p + 128 * 16 + row * 8 + (t & 1) * 4);        // 1-bit plane: 128 rows x 8 B
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? lo.x : lo.y;
    return ((x >> (2 * (i & 7))) & 0x00030003u) | (((hi >> i) & 0x00010001u) << 2);
  }
};

// owner CTA of stream-K unit u when G CTAs split U units as [floor(c*U/G), floor((c+1)*U/G))
__device__ __forceinline__ int unit_owner(int u, int G, int U) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * G + U - 1) / U) - 1;
}

// ---------------------------------------------------------------- kernel
template <typename T, int NBITS, int MB, bool POST>
__global__ void __launch_bounds__(SK_THREADS)
k3_skinny(const SkinnyArgs a) {
  using SM = SkinnySmem<NBITS, MB>;
  using T2 = typename DQ<T>::T2;
  constexpr bool kBf16 = DT<T>::kCode == Q100_BF16;
  constexpr int kTileB = WGeom<NBITS>::kTileBytes;
  constexpr int kPitch = WGeom<NBITS>::kPitch;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  // 1024-aligned carve-up
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u32 - smem_u32(smem_raw));
  __shared__ uint64_t full_bar[SK_STAGES], empty_bar[SK_STAGES];
  __shared__ int s_flag;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, U = a.total_units;
  const int cta = blockIdx.x;
  const int u_begin = static_cast<int>(static_cast<int64_t>(cta) * U / G);
  const int u_end = static_cast<int>(static_cast<int64_t>(cta + 1) * U / G);

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + 32);        // expect_tx arrive + 32 cp.async noinc arrives
      mbar_init(&empty_bar[s], 8);            // one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int groups_per_block = a.group >= SK_KB ? 1 : SK_KB / a.group;   // side-band rows per 128-k block

  if (warp == 8) {
    // ============================ producer warp ============================
    const uint64_t pol = policy_evict_first();
    const T* xg = reinterpret_cast<const T*>(a.x);
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
      const int stage = it % SK_STAGES;
      const uint32_t phase = (it / SK_STAGES) & 1;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = base + stage * a.stage_bytes;
      const int tile = u / a.units_per_tile, ku = u % a.units_per_tile;
      const int n0 = tile * SK_ROWS;
      const int kb0 = ku * a.kbu;
      const int nkb = min(a.kbu, a.kblocks - kb0);
      const int rows_valid = min(SK_ROWS, a.N - n0);
      if (lane == 0) {
        const uint32_t meta_bytes = static_cast<uint32_t>(rows_valid) * 2u * sizeof(T);
        mbar_expect_tx(&full_bar[stage], nkb * (2u * kTileB + groups_per_block * meta_bytes));
        // 2*nkb consecutive [128 x 64] tiles of this row tile: one contiguous run in HBM
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.packed) +
                              (static_cast<int64_t>(tile) * (a.kblocks * 2) + kb0 * 2) * kTileB;
        for (int kt = 0; kt < 2 * nkb; ++kt)
          bulk_load_1d_hint(st + kt * kPitch, wsrc + static_cast<int64_t>(kt) * kTileB, kTileB,
                            &full_bar[stage], pol);
        for (int kb = 0; kb < nkb; ++kb) {
          const int kblk = kb0 + kb;
          for (int gi = 0; gi < groups_per_block; ++gi) {
            const int kg = (kblk * SK_KB + gi * a.group) / a.group;
            const T* src = reinterpret_cast<const T*>(a.meta) + (static_cast<int64_t>(kg) * a.N + n0) * 2;
            bulk_load_1d(st + SM::kWBytes + (kb * groups_per_block + gi) * (SK_ROWS * 2 * sizeof(T)),
                         src, meta_bytes, &full_bar[stage]);
          }
        }
      }
      // x slab: rows m < MB*8, 16 chunks of 16 B per (m, k block); chunk c = 4t + j is stored at
      // offset 16t + 64j so that consumer LDS.128 is bank-conflict free.
      {
        const uint32_t xs = smem_u32(st + SM::kWBytes + a.meta_stage);
        const int n_chunks = nkb * MB * 8 * 16;
        for (int c = lane; c < n_chunks; c += 32) {
          const int kb = c / (MB * 8 * 16);
          const int rem = c % (MB * 8 * 16);
          const int m = rem >> 4, ch = rem & 15;
          const uint32_t dst = xs + (kb * MB * 8 + m) * SK_XPITCH + (ch >> 2) * 16 + (ch & 3) * 64;
          const bool ok = m < a.M;
          const T* src = xg + static_cast<int64_t>(ok ? m : 0) * a.K + (kb0 + kb) * SK_KB + ch * 8;
          const uint32_t src_bytes = ok ? 16u : 0u;      // src-size 0 => zero fill
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                       :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];"
                     :: "r"(smem_u32(&full_bar[stage])) : "memory");
      }
    }
    // drain: do not exit while copies are in flight
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // ============================ consumer warps ============================
  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  float acc[MB][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }
  const uint32_t qb = DT<T>::pack2(a.qmin, a.qmin);
  const T2 qmin2 = *reinterpret_cast<const T2*>(&qb);
  const bool has_zp = a.has_zp != 0;

  int it = 0;
  int seg_begin = u_begin;                      // first unit of the current tile segment
  for (int u = u_begin; u < u_end; ++u, ++it) {
    const int stage = it % SK_STAGES;
    const uint32_t phase = (it / SK_STAGES) & 1;
    const int tile = u / a.units_per_tile, ku = u % a.units_per_tile;
    const int kb0 = ku * a.kbu;
    const int nkb = min(a.kbu, a.kblocks - kb0);
    mbar_wait(&full_bar[stage], phase);
    const uint32_t st = base_u32 + stage * a.stage_bytes;
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t wblk = st + kb * 2 * kPitch;
      LaneCodes<NBITS> c0, c1;
      c0.load(wblk, r0, t);
      c1.load(wblk, r1, t);
      const int gi = a.group >= SK_KB ? 0 : (32 * t) / a.group;
      const uint32_t mt = st + SM::kWBytes + (kb * groups_per_block + gi) * (SK_ROWS * 2 * sizeof(T));
      const uint32_t m0 = lds32(mt + r0 * 4), m1 = lds32(mt + r1 * 4);
      const uint32_t xs = st + SM::kWBytes + a.meta_stage + (kb * MB * 8) * SK_XPITCH;
      if (POST) {
        // Magic-form operand: the tensor core multiplies x by (MAGIC + q) straight from an AND/OR of
        // the packed word; a second MMA with an all-ones A fragment yields S[m] = sum_k x[m][k] of
        // the block, and the group's contribution is  scale * (acc_q - (MAGIC + zp) * S)  in fp32.
        // (q - zp) is exact; the weight is never rounded to 16 bits (DESIGN.md "K3a numerics").
        constexpr uint32_t kOne2 = kBf16 ? 0x3f803f80u : 0x3c003c00u;
        const uint32_t ones[4] = {kOne2, kOne2, kOne2, kOne2};
        float ag[MB][4], so[MB][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 xb[MB];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) xb[mb] = lds128(xs + (mb * 8 + g) * SK_XPITCH + t * 16 + j * 64);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int p = 4 * j + 2 * h;
            uint32_t af[4];
            af[0] = c0.pair(p) | DQ<T>::kMagic;
            af[1] = c1.pair(p) | DQ<T>::kMagic;
            af[2] = c0.pair(p + 1) | DQ<T>::kMagic;
            af[3] = c1.pair(p + 1) | DQ<T>::kMagic;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              uint32_t bf[2];
              bf[0] = h == 0 ? xb[mb].x : xb[mb].z;
              bf[1] = h == 0 ? xb[mb].y : xb[mb].w;
              if (j == 0 && h == 0) {
                mma_16816_zero(ag[mb], af, bf, kBf16);
                mma_16816_zero(so[mb], ones, bf, kBf16);
              } else {
                mma_16816(ag[mb], af, bf, kBf16);
                mma_16816(so[mb], ones, bf, kBf16);
              }
            }
          }
        }
        const float2 sz0 = DT<T>::to_f2(m0), sz1 = DT<T>::to_f2(m1);     // (scale, zero) of rows r0, r1
        const float off0 = -(DQ<T>::kMagicF + (has_zp ? sz0.y : -a.qmin));
        const float off1 = -(DQ<T>::kMagicF + (has_zp ? sz1.y : -a.qmin));
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          acc[mb][0] = fmaf(sz0.x, fmaf(off0, so[mb][0], ag[mb][0]), acc[mb][0]);
          acc[mb][1] = fmaf(sz0.x, fmaf(off0, so[mb][1], ag[mb][1]), acc[mb][1]);
          acc[mb][2] = fmaf(sz1.x, fmaf(off1, so[mb][2], ag[mb][2]), acc[mb][2]);
          acc[mb][3] = fmaf(sz1.x, fmaf(off1, so[mb][3], ag[mb][3]), acc[mb][3]);
        }
      } else {
        const uint32_t s0b = __byte_perm(m0, 0, 0x1010), z0b = __byte_perm(m0, 0, 0x3232);
        const uint32_t s1b = __byte_perm(m1, 0, 0x1010), z1b = __byte_perm(m1, 0, 0x3232);
        const T2 s0 = *reinterpret_cast<const T2*>(&s0b), z0 = *reinterpret_cast<const T2*>(&z0b);
        const T2 s1 = *reinterpret_cast<const T2*>(&s1b), z1 = *reinterpret_cast<const T2*>(&z1b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 xb[MB];
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) xb[mb] = lds128(xs + (mb * 8 + g) * SK_XPITCH + t * 16 + j * 64);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int p = 4 * j + 2 * h;            // pair index of k = 32t + 8j + 4h + {0,1}
            uint32_t af[4];
            af[0] = dq_pair<T, NBITS>(c0.pair(p), qmin2, z0, s0, has_zp);
            af[1] = dq_pair<T, NBITS>(c1.pair(p), qmin2, z1, s1, has_zp);
            af[2] = dq_pair<T, NBITS>(c0.pair(p + 1), qmin2, z0, s0, has_zp);
            af[3] = dq_pair<T, NBITS>(c1.pair(p + 1), qmin2, z1, s1, has_zp);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              uint32_t bf[2];
              bf[0] = h == 0 ? xb[mb].x : xb[mb].z;
              bf[1] = h == 0 ? xb[mb].y : xb[mb].w;
              mma_16816(acc[mb], af, bf, kBf16);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);

    // ---- end of a tile segment?
    const bool last_of_tile = (ku == a.units_per_tile - 1);
    if (last_of_tile || u == u_end - 1) {
      const int tile_u0 = tile * a.units_per_tile;
      const int tile_u1 = tile_u0 + a.units_per_tile;           // exclusive
      const int n0 = tile * SK_ROWS;
      const bool whole = (seg_begin == tile_u0) && (u + 1 == tile_u1);
      bool write_out = whole;
      if (!whole) {
        const int c_first = unit_owner(tile_u0, G, U);
        const int c_last = unit_owner(tile_u1 - 1, G, U);
        const int n_contrib = c_last - c_first + 1;
        const int slot = cta - c_first;
        float* wsp = a.ws_partials +
                     (static_cast<int64_t>(tile) * a.max_slots + slot) * (SK_ROWS * MB * 8);
        // fragment-order layout: [warp][mb][reg][lane]
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
          for (int r = 0; r < 4; ++r) wsp[((warp * MB + mb) * 4 + r) * 32 + lane] = acc[mb][r];
        __threadfence();
        asm volatile("bar.sync 1, %0;" :: "n"(SK_CONSUMERS) : "memory");
        if (tid == 0) {
          const int old = atomicAdd(&a.ws_counters[tile], 1);
          s_flag = (old == n_contrib - 1) ? 1 : 0;
          if (old == n_contrib - 1) a.ws_counters[tile] = 0;       // leave the workspace clean
        }
        asm volatile("bar.sync 1, %0;" :: "n"(SK_CONSUMERS) : "memory");
        if (s_flag) {
          __threadfence();
          const float* wst = a.ws_partials + static_cast<int64_t>(tile) * a.max_slots * (SK_ROWS * MB * 8);
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              float s = 0.f;
              for (int sl = 0; sl < n_contrib; ++sl)
                s += __ldcg(wst + static_cast<int64_t>(sl) * (SK_ROWS * MB * 8) +
                            ((warp * MB + mb) * 4 + r) * 32 + lane);
              acc[mb][r] = s;
            }
          write_out = true;
        }
        asm volatile("bar.sync 1, %0;" :: "n"(SK_CONSUMERS) : "memory");   // s_flag reusable
      }
      if (write_out) {
        T* y = reinterpret_cast<T*>(a.y);
        const T* bias = reinterpret_cast<const T*>(a.bias);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int n = n0 + ((r & 2) ? r1 : r0);
            const int m = mb * 8 + 2 * t + (r & 1);
            if (m < a.M && n < a.N) {
              float v = acc[mb][r];
              if (bias) v += DT<T>::to_f(bias[n]);
              y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(v);
            }
          }
      }
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }
      seg_begin = u + 1;
    }
  }
}

// ---------------------------------------------------------------- host
template <typename T, int NBITS, int MB, bool POST>
static int launch_skinny(SkinnyArgs a, cudaStream_t st) {
  using SM = SkinnySmem<NBITS, MB>;
  auto kern = k3_skinny<T, NBITS, MB, POST>;
  const int smem = SM::total(a.meta_stage);
  a.stage_bytes = SM::stage(a.meta_stage);
  static std::mutex mu;           // one per instantiation
  static int max_set = 0;
  static std::unordered_map<int, int> occ_cache;
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (smem > max_set) {
      Q100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      max_set = smem;
    }
    auto itc = occ_cache.find(smem);
    if (itc == occ_cache.end()) {
      Q100_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SK_THREADS, smem));
      occ_cache[smem] = occ;
    } else {
      occ = itc->second;
    }
  }
  Q100_REQUIRE(occ >= 1, Q100_ECUDA, "k3_skinny does not fit on an SM (%d bytes smem)", smem);
  if (occ > 2) occ = 2;
  // persistent grid: every CTA resident at once, units split evenly (stream-K)
  int grid = sm_count() * occ;
  if (grid > a.total_units) grid = a.total_units;
  Q100_REQUIRE(static_cast<int64_t>(a.units_per_tile) * grid / a.total_units + 2 <= a.max_slots, Q100_EWORKSPACE,
               "internal: slot bound");
  kern<<<grid, SK_THREADS, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "k3_skinny launch");
}

template <typename T, int NBITS>
static int launch_skinny_mb(const SkinnyArgs& a, cudaStream_t st) {
  // 8-bit codes do not fit bf16's 7-bit OR-magic; that combination keeps the exact path
  constexpr bool kCanPost = !(NBITS == 8 && DT<T>::kCode == Q100_BF16);
  if (a.post && kCanPost) {
    if (a.M <= 8) return launch_skinny<T, NBITS, 1, kCanPost>(a, st);
    if (a.M <= 16) return launch_skinny<T, NBITS, 2, kCanPost>(a, st);
    return launch_skinny<T, NBITS, 4, kCanPost>(a, st);
  }
  if (a.M <= 8) return launch_skinny<T, NBITS, 1, false>(a, st);
  if (a.M <= 16) return launch_skinny<T, NBITS, 2, false>(a, st);
  return launch_skinny<T, NBITS, 4, false>(a, st);
}

template <typename T>
static int launch_skinny_bits(int n_bits, const SkinnyArgs& a, cudaStream_t st) {
  switch (n_bits) {
    case 2: return launch_skinny_mb<T, 2>(a, st);
    case 3: return launch_skinny_mb<T, 3>(a, st);
    case 4: return launch_skinny_mb<T, 4>(a, st);
    default: return launch_skinny_mb<T, 8>(a, st);
  }
}

struct SkinnyPlan {
  int kbu, units_per_tile, n_tiles, total_units, grid, max_slots;
};

static SkinnyPlan plan_skinny(int64_t N, int64_t K) {
  SkinnyPlan p;
  const int kblocks = static_cast<int>(K / SK_KB);
  p.kbu = (kblocks % SK_KBU == 0) ? SK_KBU : 1;
  p.units_per_tile = kblocks / p.kbu;
  p.n_tiles = static_cast<int>((N + SK_ROWS - 1) / SK_ROWS);
  p.total_units = p.n_tiles * p.units_per_tile;
  int grid = sm_count() * 2;                  // two CTAs per SM fit in shared memory for MB <= 2
  if (grid > p.total_units) grid = p.total_units;
  if (grid < 1) grid = 1;
  p.grid = grid;
  // a tile of `units_per_tile` units can be touched by at most this many CTAs
  p.max_slots = static_cast<int>((static_cast<int64_t>(p.units_per_tile) * grid + p.total_units - 1) /
                                 p.total_units) + 2;
  return p;
}

size_t skinny_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const SkinnyPlan p = plan_skinny(N, K);
  const int mb = M <= 8 ? 1 : M <= 16 ? 2 : 4;
  const size_t partials = static_cast<size_t>(p.n_tiles) * p.max_slots * SK_ROWS * mb * 8 * sizeof(float);
  return partials + static_cast<size_t>(p.n_tiles) * sizeof(int) + 256;
}

int gemm_skinny(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int exact_weights, cudaStream_t st) {
  Q100_REQUIRE(M >= 1 && M <= 32, Q100_EINVAL, "skinny GEMM handles 1 <= M <= 32 (got %lld)", (long long)M);
  Q100_REQUIRE(K % SK_KB == 0, Q100_EUNSUPPORTED, "K=%lld must be a multiple of 128", (long long)K);
  Q100_REQUIRE(N % 8 == 0, Q100_EUNSUPPORTED, "N=%lld must be a multiple of 8", (long long)N);
  const int g = group == 0 ? static_cast<int>(K) : group;
  Q100_REQUIRE(g == 32 || g == 64 || g % 128 == 0, Q100_EUNSUPPORTED,
               "group=%d: supported 32, 64 or a multiple of 128", g);
  Q100_REQUIRE(K % g == 0, Q100_EINVAL, "K %% group != 0");
  const SkinnyPlan p = plan_skinny(N, K);
  const size_t need = skinny_workspace_bytes(M, N, K);
  Q100_REQUIRE(workspace != nullptr && workspace_bytes >= need, Q100_EWORKSPACE,
               "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);

  SkinnyArgs a{};
  a.x = x; a.packed = packed; a.meta = meta; a.bias = bias; a.y = y;
  const int mb = M <= 8 ? 1 : M <= 16 ? 2 : 4;
  const size_t partials = static_cast<size_t>(p.n_tiles) * p.max_slots * SK_ROWS * mb * 8 * sizeof(float);
  a.ws_counters = reinterpret_cast<int*>(workspace);
  a.ws_partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) +
                                           ((static_cast<size_t>(p.n_tiles) * sizeof(int) + 255) & ~size_t(255)));
  (void)partials;
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = g;
  a.kblocks = static_cast<int>(K / SK_KB);
  a.kbu = p.kbu; a.units_per_tile = p.units_per_tile; a.n_tiles = p.n_tiles;
  a.total_units = p.total_units; a.max_slots = p.max_slots;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  const int gpb = g >= SK_KB ? 1 : SK_KB / g;
  a.post = (g % SK_KB == 0 && !exact_weights) ? 1 : 0;
  a.meta_stage = ((p.kbu * gpb * SK_ROWS * 2 * 2) + 127) & ~127;
  if (dtype == Q100_F16) return launch_skinny_bits<__half>(n_bits, a, st);
  return launch_skinny_bits<__nv_bfloat16>(n_bits, a, st);
}

}  // namespace q100
