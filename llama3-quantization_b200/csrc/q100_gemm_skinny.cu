// q100_gemm_skinny.cu -- K3a: weight-streaming dequant-GEMM for decode shapes (M <= 32).
//
//   y[M, N] = x[M, K] . dequant(W)[N, K]^T (+ bias)        quant/int_linear.py:62
//
// HBM-bound: the packed weights (n_bits/8 B per weight) are the only large stream and are read
// exactly once.  Structure (persistent CTAs of 8 warps):
//   * every consumer warp owns 16 output rows of the CTA's 128-row tile and streams ITS OWN
//     slice of the packed codes straight from global memory into registers with 128-bit
//     ld.global.nc loads, four 128-k blocks ahead (software-pipelined register ring).  The
//     tile-major packed layout makes each warp-wide load two fully used 256-byte segments and
//     the CTA's whole stream one linear run of HBM.  (A TMA/shared-memory ring for the codes was
//     measured first: it topped out at 3.9 TB/s with no compute at all -- profiles/ r01 notes --
//     because the per-SM bulk-copy path cannot keep enough bytes in flight against HBM latency;
//     registers can.)
//   * x is tiny: the CTA stages the slab of x it needs (up to 16 k-blocks at a time) into shared
//     memory itself, in a bank-conflict-free permuted order, together with the per-block sums
//     S[m] = sum_k x[m][k]; no producer warp, no mbarrier ring, one __syncthreads per 2048 k;
//   * codes are used in "magic form": (MAGIC + q) is built with one AND/OR per pair and fed to
//     mma.sync.m16n8k16 as the A operand (x^T is B); the block's contribution is
//     scale * (acc_q - (MAGIC + zp) * S) in fp32.  With Q100_GEMM_EXACT_WEIGHTS (or group < 128)
//     the 16-bit fake-quantised weight r(r(q-zp)*scale) is rebuilt bit-exactly instead.
//     tcgen05 is deliberately not used here: at M <= 32 the operand would have to be written
//     back to shared memory (2 B/weight) and that traffic alone exceeds the time budget of the
//     0.5 B/weight HBM stream (DESIGN.md "K3a").
//   * work is split stream-K style over 148*k CTAs in units of [128 rows x 256 k]; a tile shared
//     by several CTAs is finished warp-asynchronously: each warp stores its fp32 fragment to its
//     CTA's slot, fences, bumps the counter of ITS 16-row fragment, and the last arriver sums that
//     fragment's slots in slot order (deterministic) and writes y.  No CTA-wide barrier there.
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
constexpr int SK_KB = 128;         // k per block
constexpr int SK_KBU = 2;          // k blocks per stream-K unit
constexpr int SK_DEPTH = 2;        // k blocks of packed codes in flight per lane (register ring); occupancy,
                                   // not per-warp depth, is what fills HBM (profiles/r01 stream microbenchmarks)
constexpr int SK_MAX_CTAS = 4;     // resident CTAs per SM the launch bounds aim for (MB = 1)
constexpr int SK_CHUNK = 16;       // k blocks of x staged in shared memory at a time (multiple of SK_DEPTH)
constexpr int SK_THREADS = 256;
constexpr size_t SK_COUNTER_BYTES = 1 << 18;   // workspace prefix: per-fragment arrival counters (N <= 1 M rows)
constexpr int SK_XPITCH = 320;     // bytes per (m, 128-k) x row in smem: 256 + 64 pad (conflict-free LDS.128)

struct SkinnyArgs {
  const void* x;          // [M, K] T
  const void* packed;     // tile-major packed codes (layout v2)
  const void* meta;       // [K/group, N, 2] T
  const void* bias;       // [N] T or null
  void* y;                // [M, N] T
  float* ws_partials;     // workspace: fp32 partial tiles
  int* ws_counters;       // workspace: one arrival counter per 16-row fragment (zero between launches)
  int M, N, K, group;
  int kblocks;            // K / 128
  int units_per_tile;     // kblocks / 2
  int n_tiles;
  int total_units;
  int max_slots;
  int has_zp;
  float qmin;
  int post;               // 1: magic-form operand + fp32 group scaling (group % 128 == 0)
  int m_rows;             // rows of x kept in shared memory (= M)
};

template <int NBITS> struct WGeom {
  static constexpr int kTileBytes = NBITS == 3 ? 128 * 24 : 128 * 8 * NBITS;   // one [128 x 64] tile
};

// shared memory: x slab [SK_CHUNK blocks][m_rows][SK_XPITCH] + block sums [SK_CHUNK][MB*8] fp32
__host__ __device__ constexpr int skinny_smem_bytes(int m_rows, int mb) {
  return SK_CHUNK * m_rows * SK_XPITCH + SK_CHUNK * mb * 8 * 4 + 128;
}

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
__device__ __forceinline__ uint2 ldg64_stream(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg32_stream(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// A lane's 32 codes (k in [32t, 32t+32) of one row's 128-k block) as 16 pairs, loaded from the
// two consecutive [128 x 64] tiles of the block: lane t reads k-tile (t >> 1), half (t & 1).
template <int NBITS> struct LaneCodes;
template <> struct LaneCodes<4> {
  uint4 w;
  __device__ __forceinline__ void load(const uint8_t* blk, int row, int t) {
    w = ldg_stream(blk + (t >> 1) * WGeom<4>::kTileBytes + row * 32 + (t & 1) * 16);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 2) == 0 ? w.x : (i >> 2) == 1 ? w.y : (i >> 2) == 2 ? w.z : w.w;
    return (x >> (4 * (i & 3))) & 0x000f000fu;
  }
};
template <> struct LaneCodes<2> {
  uint2 w;
  __device__ __forceinline__ void load(const uint8_t* blk, int row, int t) {
    w = ldg64_stream(blk + (t >> 1) * WGeom<2>::kTileBytes + row * 16 + (t & 1) * 8);
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? w.x : w.y;
    return (x >> (2 * (i & 7))) & 0x00030003u;
  }
};
template <> struct LaneCodes<8> {
  uint4 w0, w1;
  __device__ __forceinline__ void load(const uint8_t* blk, int row, int t) {
    const uint8_t* p = blk + (t >> 1) * WGeom<8>::kTileBytes + row * 64 + (t & 1) * 32;
    w0 = ldg_stream(p);
    w1 = ldg_stream(p + 16);
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
  __device__ __forceinline__ void load(const uint8_t* blk, int row, int t) {
    const uint8_t* p = blk + (t >> 1) * WGeom<3>::kTileBytes;
    lo = ldg64_stream(p + row * 16 + (t & 1) * 8);                  // 2-bit plane: 128 rows x 16 B
    hi = ldg32_stream(p + 128 * 16 + row * 8 + (t & 1) * 4);        // 1-bit plane: 128 rows x 8 B
  }
  __device__ __forceinline__ uint32_t pair(int i) const {
    const uint32_t x = (i >> 3) == 0 ? lo.x : lo.y;
    return ((x >> (2 * (i & 7))) & 0x00030003u) | (((hi >> i) & 0x00010001u) << 2);
  }
};

// everything a lane needs for one 128-k block: codes of its two rows + their (scale, zero) pairs
template <int NBITS>
struct BlockRegs {
  LaneCodes<NBITS> c0, c1;
  uint32_t m0, m1;
};

// owner CTA of stream-K unit u when G CTAs split U units as [floor(c*U/G), floor((c+1)*U/G))
__device__ __forceinline__ int unit_owner(int u, int G, int U) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * G + U - 1) / U) - 1;
}

// ---------------------------------------------------------------- kernel
template <typename T, int NBITS, int MB, bool POST>
__global__ void __launch_bounds__(SK_THREADS, MB == 1 ? 4 : MB == 2 ? 3 : 2)
k3_skinny(const SkinnyArgs a) {
  using T2 = typename DQ<T>::T2;
  constexpr bool kBf16 = DT<T>::kCode == Q100_BF16;
  constexpr int kTileB = WGeom<NBITS>::kTileBytes;
  constexpr int FRAG = SK_ROWS * MB * 8;              // floats per partial tile

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t xs_u32 = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* xs_ptr = smem_raw + (xs_u32 - smem_u32(smem_raw));
  const int mrows = a.m_rows;
  float* xsum_s = reinterpret_cast<float*>(xs_ptr + SK_CHUNK * mrows * SK_XPITCH);     // [SK_CHUNK][MB*8]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, U = a.total_units;
  const int cta = blockIdx.x;
  const int u_begin = static_cast<int>(static_cast<int64_t>(cta) * U / G);
  const int u_end = static_cast<int>(static_cast<int64_t>(cta + 1) * U / G);

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const uint32_t qb = DT<T>::pack2(a.qmin, a.qmin);
  const T2 qmin2 = *reinterpret_cast<const T2*>(&qb);
  const bool has_zp = a.has_zp != 0;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(a.packed);
  const uint8_t* mbase = reinterpret_cast<const uint8_t*>(a.meta);

  // linear k-block index q = tile * kblocks + kblk; block q starts at tile-major tile 2q.  The load
  // cursor (tile, kblk) advances with q, so there is no integer division per block.
  int ld_tile = (u_begin * SK_KBU) / a.kblocks, ld_kblk = (u_begin * SK_KBU) - ld_tile * a.kblocks;
  const int gdiv = a.group / SK_KB;                       // groups >= 128: k blocks per group
  auto load_block = [&](BlockRegs<NBITS>& b, int q) {
    const uint8_t* blk = wbase + static_cast<int64_t>(q) * (2 * kTileB);
    b.c0.load(blk, r0, t);
    b.c1.load(blk, r1, t);
    int kg;
    if (a.group >= SK_KB) kg = gdiv == 1 ? ld_kblk : (gdiv >= a.kblocks ? 0 : ld_kblk / gdiv);
    else kg = ld_kblk * (SK_KB / a.group) + (a.group == 32 ? t : (t >> 1));
    const int n0 = ld_tile * SK_ROWS;
    const int64_t row_base = static_cast<int64_t>(kg) * a.N;
    b.m0 = ldg32_stream(mbase + (row_base + min(n0 + r0, a.N - 1)) * (2 * sizeof(T)));
    b.m1 = ldg32_stream(mbase + (row_base + min(n0 + r1, a.N - 1)) * (2 * sizeof(T)));
    if (++ld_kblk == a.kblocks) { ld_kblk = 0; ++ld_tile; }
  };

  // x slab of local blocks [c0, c0 + n): rows m < M only; 16-byte chunk c = 4t + j of a block is stored
  // at offset 16t + 64j (conflict-free LDS.128 for the mma B fragments); S[m] = sum of the block.
  auto stage_x = [&](int q0, int n) {
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int total = n * mrows * 16;                      // 16-byte chunks
    for (int c = tid; c < total; c += SK_THREADS) {
      const int bl = c / (mrows * 16);
      const int rem = c - bl * (mrows * 16);
      const int m = rem >> 4, ch = rem & 15;
      const int kblk = (q0 + bl) % a.kblocks;
      const uint4 v = *reinterpret_cast<const uint4*>(xg + static_cast<int64_t>(m) * a.K + kblk * SK_KB + ch * 8);
      *reinterpret_cast<uint4*>(xs_ptr + (bl * mrows + m) * SK_XPITCH + (ch >> 2) * 16 + (ch & 3) * 64) = v;
      if (POST) {
        const float2 f0 = DT<T>::to_f2(v.x), f1 = DT<T>::to_f2(v.y), f2 = DT<T>::to_f2(v.z), f3 = DT<T>::to_f2(v.w);
        float sm = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
        // the 16 chunks of one (block, m) sit in 16 consecutive threads
        const uint32_t half_mask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // total is a multiple of 16
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) sm += __shfl_xor_sync(half_mask, sm, d);
        if (ch == 0) xsum_s[bl * (MB * 8) + m] = sm;
      }
    }
  };

  float acc[MB][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }

  auto compute_block = [&](const BlockRegs<NBITS>& b, int bl) {
    const uint32_t xs = xs_u32 + (bl * mrows) * SK_XPITCH;
    if (POST) {
      float ag[MB][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 xb[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
          xb[mb] = (mb * 8 + g < mrows) ? lds128(xs + (mb * 8 + g) * SK_XPITCH + t * 16 + j * 64) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = 4 * j + 2 * h;            // pair index of k = 32t + 8j + 4h + {0,1}
          uint32_t af[4];
          af[0] = b.c0.pair(p) | DQ<T>::kMagic;
          af[1] = b.c1.pair(p) | DQ<T>::kMagic;
          af[2] = b.c0.pair(p + 1) | DQ<T>::kMagic;
          af[3] = b.c1.pair(p + 1) | DQ<T>::kMagic;
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            uint32_t bf[2];
            bf[0] = h == 0 ? xb[mb].x : xb[mb].z;
            bf[1] = h == 0 ? xb[mb].y : xb[mb].w;
            if (j == 0 && h == 0) mma_16816_zero(ag[mb], af, bf, kBf16);
            else mma_16816(ag[mb], af, bf, kBf16);
          }
        }
      }
      const float2 sz0 = DT<T>::to_f2(b.m0), sz1 = DT<T>::to_f2(b.m1);     // (scale, zero) of rows r0, r1
      const float off0 = -(DQ<T>::kMagicF + (has_zp ? sz0.y : -a.qmin));
      const float off1 = -(DQ<T>::kMagicF + (has_zp ? sz1.y : -a.qmin));
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int m = mb * 8 + 2 * t;
        const float s_a = (m < mrows) ? xsum_s[bl * (MB * 8) + m] : 0.f;
        const float s_b = (m + 1 < mrows) ? xsum_s[bl * (MB * 8) + m + 1] : 0.f;
        acc[mb][0] = fmaf(sz0.x, fmaf(off0, s_a, ag[mb][0]), acc[mb][0]);
        acc[mb][1] = fmaf(sz0.x, fmaf(off0, s_b, ag[mb][1]), acc[mb][1]);
        acc[mb][2] = fmaf(sz1.x, fmaf(off1, s_a, ag[mb][2]), acc[mb][2]);
        acc[mb][3] = fmaf(sz1.x, fmaf(off1, s_b, ag[mb][3]), acc[mb][3]);
      }
    } else {
      const uint32_t s0b = __byte_perm(b.m0, 0, 0x1010), z0b = __byte_perm(b.m0, 0, 0x3232);
      const uint32_t s1b = __byte_perm(b.m1, 0, 0x1010), z1b = __byte_perm(b.m1, 0, 0x3232);
      const T2 s0 = *reinterpret_cast<const T2*>(&s0b), z0 = *reinterpret_cast<const T2*>(&z0b);
      const T2 s1 = *reinterpret_cast<const T2*>(&s1b), z1 = *reinterpret_cast<const T2*>(&z1b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 xb[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
          xb[mb] = (mb * 8 + g < mrows) ? lds128(xs + (mb * 8 + g) * SK_XPITCH + t * 16 + j * 64) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = 4 * j + 2 * h;
          uint32_t af[4];
          af[0] = dq_pair<T, NBITS>(b.c0.pair(p), qmin2, z0, s0, has_zp);
          af[1] = dq_pair<T, NBITS>(b.c1.pair(p), qmin2, z1, s1, has_zp);
          af[2] = dq_pair<T, NBITS>(b.c0.pair(p + 1), qmin2, z0, s0, has_zp);
          af[3] = dq_pair<T, NBITS>(b.c1.pair(p + 1), qmin2, z1, s1, has_zp);
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
  };

  // y <- fragment `v` of this warp's 16 rows of tile n0 (mma D layout), + bias, converted to T
  auto write_fragment = [&](const float (&v)[MB][4], int n0) {
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = n0 + ((r & 2) ? r1 : r0);
        const int m = mb * 8 + 2 * t + (r & 1);
        if (m < a.M && n < a.N) {
          float o = v[mb][r];
          if (bias) o += DT<T>::to_f(bias[n]);
          y[static_cast<int64_t>(m) * a.N + n] = DT<T>::from_f(o);
        }
      }
  };

  // end of a tile segment [seg_begin, u]: write y or hand the fragment to the stream-K fix-up
  auto finish_segment = [&](int u, int seg_begin) {
    const int tile = u / a.units_per_tile;
    const int tile_u0 = tile * a.units_per_tile;
    const int tile_u1 = tile_u0 + a.units_per_tile;           // exclusive
    const int n0 = tile * SK_ROWS;
    if (seg_begin == tile_u0 && u + 1 == tile_u1) {
      write_fragment(acc, n0);                                // this CTA owns the whole tile
    } else {
      const int c_first = unit_owner(tile_u0, G, U);
      const int c_last = unit_owner(tile_u1 - 1, G, U);
      const int n_contrib = c_last - c_first + 1;
      // slot layout: [slot][warp][mb][reg][lane]; this warp only ever touches its own [warp] part
      float* wst = a.ws_partials + static_cast<int64_t>(tile) * a.max_slots * FRAG + warp * (MB * 4 * 32);
      float* wsp = wst + static_cast<int64_t>(cta - c_first) * FRAG;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int r = 0; r < 4; ++r) wsp[(mb * 4 + r) * 32 + lane] = acc[mb][r];
      __threadfence();
      __syncwarp();
      int old = 0;
      int* cnt = a.ws_counters + tile * 8 + warp;
      if (lane == 0) old = atomicAdd(cnt, 1);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == n_contrib - 1) {
        // last contributor of this 16-row fragment: all slots are visible
        __threadfence();
        float v[MB][4];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
          for (int r = 0; r < 4; ++r) v[mb][r] = 0.f;
        for (int sl = 0; sl < n_contrib; ++sl) {
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int r = 0; r < 4; ++r)
              v[mb][r] += __ldcg(wst + static_cast<int64_t>(sl) * FRAG + (mb * 4 + r) * 32 + lane);
        }
        write_fragment(v, n0);
        if (lane == 0) *cnt = 0;                              // leave the workspace clean
      }
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) { acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f; }
  };

  const int q_begin = u_begin * SK_KBU, q_end = u_end * SK_KBU;
  BlockRegs<NBITS> ring[SK_DEPTH];
#pragma unroll
  for (int d = 0; d < SK_DEPTH; ++d)
    if (q_begin + d < q_end) load_block(ring[d], q_begin + d);

  int seg_begin = u_begin;                      // first unit of the current tile segment
  for (int qc = q_begin; qc < q_end; qc += SK_CHUNK) {
    const int nblk = min(SK_CHUNK, q_end - qc);
    __syncthreads();                            // everyone is done with the previous slab
    stage_x(qc, nblk);
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += SK_DEPTH) {
#pragma unroll
      for (int d = 0; d < SK_DEPTH; ++d) {
        const int q = qc + b0 + d;
        if (q < q_end) {
          compute_block(ring[d], b0 + d);
          if (q + SK_DEPTH < q_end) load_block(ring[d], q + SK_DEPTH);
          if (d & 1) {                          // a unit = 2 k blocks: unit boundary after the odd block
            const int u = q / SK_KBU;
            const int ku = u % a.units_per_tile;
            if (ku == a.units_per_tile - 1 || u == u_end - 1) {
              finish_segment(u, seg_begin);
              seg_begin = u + 1;
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- host
template <typename T, int NBITS, int MB, bool POST>
static int launch_skinny(SkinnyArgs a, cudaStream_t st) {
  auto kern = k3_skinny<T, NBITS, MB, POST>;
  const int smem = skinny_smem_bytes(a.m_rows, MB);
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
  // persistent grid: every CTA resident at once, units split evenly (stream-K)
  int grid = sm_count() * (occ > SK_MAX_CTAS ? SK_MAX_CTAS : occ);
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
  int units_per_tile, n_tiles, total_units, max_slots;
};

static SkinnyPlan plan_skinny(int64_t N, int64_t K) {
  SkinnyPlan p;
  const int kblocks = static_cast<int>(K / SK_KB);
  p.units_per_tile = kblocks / SK_KBU;
  p.n_tiles = static_cast<int>((N + SK_ROWS - 1) / SK_ROWS);
  p.total_units = p.n_tiles * p.units_per_tile;
  int grid = sm_count() * SK_MAX_CTAS;        // upper bound of the persistent grid
  if (grid > p.total_units) grid = p.total_units;
  if (grid < 1) grid = 1;
  // a tile of `units_per_tile` units can be touched by at most this many CTAs
  p.max_slots = static_cast<int>((static_cast<int64_t>(p.units_per_tile) * grid + p.total_units - 1) /
                                 p.total_units) + 2;
  return p;
}

size_t skinny_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const SkinnyPlan p = plan_skinny(N, K);
  const int mb = M <= 8 ? 1 : M <= 16 ? 2 : 4;
  const size_t partials = static_cast<size_t>(p.n_tiles) * p.max_slots * SK_ROWS * mb * 8 * sizeof(float);
  return partials + SK_COUNTER_BYTES;
}

int gemm_skinny(const void* x, const void* packed, const void* meta, const void* bias, void* y,
                int64_t M, int64_t N, int64_t K, int group, int n_bits, int disable_zp, int dtype,
                void* workspace, size_t workspace_bytes, int exact_weights, cudaStream_t st) {
  Q100_REQUIRE(M >= 1 && M <= 32, Q100_EINVAL, "skinny GEMM handles 1 <= M <= 32 (got %lld)", (long long)M);
  Q100_REQUIRE(K % (SK_KB * SK_KBU) == 0, Q100_EUNSUPPORTED, "K=%lld must be a multiple of 256", (long long)K);
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
  a.ws_counters = reinterpret_cast<int*>(workspace);
  // counters live in a fixed-size prefix so that partials of one launch can never land on what a
  // later launch (different N) uses as counters: the prefix is the only part that must stay zero
  Q100_REQUIRE(static_cast<size_t>(p.n_tiles) * 8 * sizeof(int) <= SK_COUNTER_BYTES, Q100_EUNSUPPORTED,
               "N=%lld too large for the stream-K counter area", (long long)N);
  a.ws_partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + SK_COUNTER_BYTES);
  a.M = static_cast<int>(M); a.N = static_cast<int>(N); a.K = static_cast<int>(K); a.group = g;
  a.kblocks = static_cast<int>(K / SK_KB);
  a.units_per_tile = p.units_per_tile; a.n_tiles = p.n_tiles;
  a.total_units = p.total_units; a.max_slots = p.max_slots;
  a.has_zp = disable_zp ? 0 : 1;
  a.qmin = disable_zp ? -static_cast<float>(1 << (n_bits - 1)) : 0.0f;
  a.post = (g % SK_KB == 0 && !exact_weights) ? 1 : 0;
  a.m_rows = static_cast<int>(M);
  if (dtype == Q100_F16) return launch_skinny_bits<__half>(n_bits, a, st);
  return launch_skinny_bits<__nv_bfloat16>(n_bits, a, st);
}

}  // namespace q100
