// q100_peer.cu -- K7: one-shot all-reduce(sum) of a small tensor over peer memory (NVLink / NVSwitch).
//
// The only data-path collective of the tensor-parallel block is the all-reduce after the row-parallel linears
// (o_proj, down_proj).  At decode shapes that is 8-128 KB per call: latency, not bandwidth.
// NCCL spends ~10-15 us on it; this kernel needs one NVLink store round:
//
//   push   every rank stores its partial into slot [parity][rank] of EVERY rank's staging buffer (peer-mapped
//          symmetric memory), fences, then stores the call's epoch into flag [parity][rank][cta] of every rank;
//   wait   CTA c polls its own flags [parity][r][c] (acquire) until all `world` of them carry the epoch;
//   sum    it adds the `world` local slots in rank order (fp32, one rounding): every rank computes the same bits.
//
// The protocol is chunk-local (CTA c of every rank owns elements [c*per, (c+1)*per)), so no grid-wide barrier is
// needed.  Two parity sets make slot reuse safe: a rank can only start call e+2 after every peer has flagged e+1,
// i.e. after every peer has finished summing call e.  The epoch lives in device memory (one counter per CTA,
// advanced by the kernel), so the launch has no per-call host state and can be captured in a CUDA graph.
#include "q100_common.cuh"

namespace q100 {

constexpr int K7_MAX_CTAS = 32;   // flag words per (parity, rank)
constexpr int K7_CTAS = 16;       // CTAs per call

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) k7_allreduce(const T* x, T* y /* may alias x */, int64_t n,
                                                    T* const* __restrict__ bufs, uint32_t* const* __restrict__ flags,
                                                    uint32_t* __restrict__ epoch_p, int64_t slot_elems, int rank,
                                                    int world, uint32_t* __restrict__ err) {
  using D = DT<T>;
  const int c = blockIdx.x, tid = threadIdx.x;
  const uint32_t e = epoch_p[c] + 1;
  const int par = static_cast<int>(e & 1u);
  const int64_t nvec = n >> 3;                                   // 16-byte vectors (n % 8 == 0)
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = c * per, v1 = min(nvec, v0 + per);

  // ---- push my chunk into slot [par][rank] of every rank (own buffer included), peers first
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  for (int k = 1; k <= world; ++k) {
    const int p = (rank + k) % world;
    uint4* dst = reinterpret_cast<uint4*>(bufs[p] + (static_cast<int64_t>(par) * world + rank) * slot_elems);
    for (int64_t v = v0 + tid; v < v1; v += blockDim.x) dst[v] = xs[v];
  }
  // bar.sync orders every thread's stores before the flag writers; st.release.sys (= fence.acq_rel.sys + store) is
  // cumulative over them, so no per-thread MEMBAR.SYS is needed
  __syncthreads();
  if (tid < world)
    st_release_sys(flags[tid] + (static_cast<int64_t>(par) * world + rank) * K7_MAX_CTAS + c, e);

  // ---- wait for every rank's chunk c of this call
  if (tid < world) {
    const uint32_t* f = flags[rank] + (static_cast<int64_t>(par) * world + tid) * K7_MAX_CTAS + c;
    const unsigned long long t0 = gtimer();
    while (ld_acquire_sys(f) != e) {
      if (gtimer() - t0 > 30000000000ull) { atomicExch(err, 1u); break; }     // 30 s: a peer is gone; do not hang
    }
  }
  __syncthreads();

  // ---- sum the world slots in rank order
  const T* mine = bufs[rank] + static_cast<int64_t>(par) * world * slot_elems;
  uint4* ys = reinterpret_cast<uint4*>(y);
  for (int64_t v = v0 + tid; v < v1; v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r = 0; r < world; ++r) {
      const uint4 q = reinterpret_cast<const uint4*>(mine + r * slot_elems)[v];
      const float2 a = D::to_f2(q.x), b = D::to_f2(q.y), cc = D::to_f2(q.z), d = D::to_f2(q.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
    }
    ys[v] = make_uint4(D::pack2(acc[0], acc[1]), D::pack2(acc[2], acc[3]), D::pack2(acc[4], acc[5]),
                       D::pack2(acc[6], acc[7]));
  }
  __syncthreads();
  if (tid == 0) epoch_p[c] = e;
}

// ---------------------------------------------------------------- K7b: reduce + broadcast after a scatter-epilogue GEMM
// The second half of the row-parallel all-reduce at prefill sizes.  q100_gemm_wna16_scatter has stored every rank's
// partial of the token rows a rank owns into that rank's staging buffer stage[src rank][rows_per][N] while the GEMM
// was still running; this kernel (next in the stream) waits until every peer's GEMM is complete, sums the `world`
// slots of its rows in rank order in fp32 -- the same bits on every rank as an fp32 sum of the gathered partials --
// and stores the result into EVERY rank's y (rows [rank * rows_per, ...)), then waits until every peer has done so.
// Chunk-local protocol like K7: CTA c of every rank handshakes with CTA c of the others (a fixed CTA count keeps the
// per-CTA epochs in lockstep, the epoch lives in device memory: graph-capturable).  Flags: [barrier 2][parity 2]
// [rank][K7B_CTAS] words per rank.
constexpr int K7B_CTAS = 128;

template <typename T>
__global__ void __launch_bounds__(256) k7b_reduce_bcast(const T* __restrict__ stage, T* const* __restrict__ ys,
                                                        uint32_t* const* __restrict__ flags, uint32_t* __restrict__ epoch_p,
                                                        int64_t slot_elems, int64_t n_own, int rank, int world,
                                                        uint32_t* __restrict__ err) {
  using D = DT<T>;
  const int c = blockIdx.x, tid = threadIdx.x;
  const uint32_t e = epoch_p[c] + 1;
  const int par = static_cast<int>(e & 1u);
  auto flag_of = [&](int barrier, int r) {
    return ((static_cast<int64_t>(barrier) * 2 + par) * world + r) * K7B_CTAS + c;
  };
  auto handshake = [&](int barrier) {
    __syncthreads();                                   // this CTA's stores (if any) precede the flag writers
    if (tid < world) {
      st_release_sys(flags[tid] + flag_of(barrier, rank), e);
      const uint32_t* f = flags[rank] + flag_of(barrier, tid);
      const unsigned long long t0 = gtimer();
      while (ld_acquire_sys(f) != e) {
        if (gtimer() - t0 > 30000000000ull) { atomicExch(err, 1u); break; }
      }
    }
    __syncthreads();
  };
  // (A) every peer's GEMM -- the kernel before this one in its stream -- has completed: its partials are in `stage`
  handshake(0);
  const int64_t nvec = n_own >> 3;
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = c * per, v1 = min(nvec, v0 + per);
  // four 16-byte vectors per thread and round, all `world` slots of each requested before the first sum: enough
  // bytes in flight to keep HBM (the local slot reads) and NVLink (the stores) busy from 128 CTAs
  constexpr int U = 4;
  for (int64_t vb = v0 + tid; vb < v1; vb += static_cast<int64_t>(blockDim.x) * U) {
    float acc[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[u][i] = 0.f;
    for (int r = 0; r < world; ++r) {
      uint4 q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = vb + static_cast<int64_t>(u) * blockDim.x;
        q[u] = v < v1 ? __ldcs(reinterpret_cast<const uint4*>(stage + r * slot_elems) + v) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float2 a = D::to_f2(q[u].x), b = D::to_f2(q[u].y), cc = D::to_f2(q[u].z), d = D::to_f2(q[u].w);
        acc[u][0] += a.x; acc[u][1] += a.y; acc[u][2] += b.x; acc[u][3] += b.y;
        acc[u][4] += cc.x; acc[u][5] += cc.y; acc[u][6] += d.x; acc[u][7] += d.y;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + static_cast<int64_t>(u) * blockDim.x;
      if (v < v1) {
        const uint4 o = make_uint4(D::pack2(acc[u][0], acc[u][1]), D::pack2(acc[u][2], acc[u][3]),
                                   D::pack2(acc[u][4], acc[u][5]), D::pack2(acc[u][6], acc[u][7]));
        for (int k = 0; k < world; ++k) {                // own copy last: the NVLink stores leave first
          const int p = (rank + 1 + k) % world;
          reinterpret_cast<uint4*>(ys[p] + static_cast<int64_t>(rank) * n_own)[v] = o;
        }
      }
    }
  }
  // (B) every peer has stored its rows into my y
  handshake(1);
  if (tid == 0) epoch_p[c] = e;
}

// ---------------------------------------------------------------- K7n: all-reduce + residual add + RMSNorm, one launch
// A tensor-parallel decode step ran  GEMM -> K7 (all-reduce) -> K4 (residual add + RMSNorm)  after o_proj and down_proj:
// two launches of pure latency per row-parallel linear.  Here K4's CTA (one per token row) does K7's exchange itself:
//   push   its partial row into slot [parity][rank][row] of EVERY rank's staging buffer, then the call's epoch into flag
//          [parity][rank][row] of every rank;
//   wait   for the `world` flags of its row;
//   sum    the `world` slots in rank order in fp32, ONE rounding to T -- the bits K7 produces;
//   then K4's arithmetic on that row: h = r(residual + y), norm = weight * (h * rsqrt(mean(h^2) + eps)).
// Per-row epochs in device memory (graph-capturable), two parity sets for slot reuse as in K7, 30 s time-out.
constexpr int K7N_MAX_ROWS = 64;

// One thread per 16-byte vector of the row (cols / 8 <= 1024 threads): a thread's `world` slot reads are independent
// loads in flight together (with 256 threads and four vectors per thread the rank loop was a chain of dependent round
// trips: 17 us per call at 8 ranks x 8192 columns).
template <typename T>
__global__ void __launch_bounds__(1024) k7n_allreduce_add_rmsnorm(const T* __restrict__ x, const T* __restrict__ res,
                                                                  const T* __restrict__ w, float eps,
                                                                  T* __restrict__ sum_out, T* __restrict__ norm_out,
                                                                  T* const* __restrict__ bufs,
                                                                  uint32_t* const* __restrict__ flags,
                                                                  uint32_t* __restrict__ epoch_p, int cols, int rank,
                                                                  int world, uint32_t* __restrict__ err) {
  using D = DT<T>;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int row = blockIdx.x, tid = threadIdx.x;
  const int nvec = cols >> 3;
  const bool act = tid < nvec;
  const uint32_t e = epoch_p[row] + 1;
  const int par = static_cast<int>(e & 1u);
  const int64_t slot = static_cast<int64_t>(K7N_MAX_ROWS) * cols;           // elements per (parity, rank) slot
  const int64_t roff = static_cast<int64_t>(row) * cols;
  // ---- push my partial row, peers first
  if (act) {
    const uint4 mine = reinterpret_cast<const uint4*>(x + roff)[tid];
    for (int k = 1; k <= world; ++k) {
      const int p = (rank + k) % world;
      reinterpret_cast<uint4*>(bufs[p] + (static_cast<int64_t>(par) * world + rank) * slot + roff)[tid] = mine;
    }
  }
  __syncthreads();
  if (tid < world) {
    st_release_sys(flags[tid] + (static_cast<int64_t>(par) * world + rank) * K7N_MAX_ROWS + row, e);
    const uint32_t* f = flags[rank] + (static_cast<int64_t>(par) * world + tid) * K7N_MAX_ROWS + row;
    const unsigned long long t0 = gtimer();
    while (ld_acquire_sys(f) != e) {
      if (gtimer() - t0 > 30000000000ull) { atomicExch(err, 1u); break; }     // 30 s: a peer is gone; do not hang
    }
  }
  __syncthreads();
  // ---- sum in rank order (fp32, one rounding), residual add (rounded to T), sum of squares
  uint4 h = make_uint4(0u, 0u, 0u, 0u);
  if (act) {
    const T* base = bufs[rank] + static_cast<int64_t>(par) * world * slot + roff;
    uint4 rv = make_uint4(0u, 0u, 0u, 0u);
    if (res) rv = reinterpret_cast<const uint4*>(res + roff)[tid];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int r0 = 0; r0 < world; r0 += 8) {                // eight independent loads in flight, summed in rank order
      uint4 q[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < world) q[r] = __ldcg(reinterpret_cast<const uint4*>(base + (r0 + r) * slot) + tid);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r0 + r < world) {
          const float2 a = D::to_f2(q[r].x), b = D::to_f2(q[r].y), cc = D::to_f2(q[r].z), d = D::to_f2(q[r].w);
          acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
          acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
        }
      }
    }
    h = make_uint4(D::pack2(acc[0], acc[1]), D::pack2(acc[2], acc[3]), D::pack2(acc[4], acc[5]), D::pack2(acc[6], acc[7]));
    if (res) {
      uint32_t* aw = reinterpret_cast<uint32_t*>(&h);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {                       // residual + hidden_states, rounded to T (K4)
        const float2 fa = D::to_f2(aw[j]), fb = D::to_f2(bw[j]);
        aw[j] = D::pack2(__fadd_rn(fb.x, fa.x), __fadd_rn(fb.y, fa.y));
      }
    }
    if (sum_out) reinterpret_cast<uint4*>(sum_out + roff)[tid] = h;
  }
  // the sum of squares in K4's order (bit-identical rstd): K4's thread t runs ONE fmaf chain over the vectors
  // t, t + 256, t + 512, t + 768; here those vectors belong to four threads, so they meet in shared memory
  __shared__ uint4 hs[1024];
  __shared__ float red[8];
  hs[tid] = h;
  __syncthreads();
  if (tid < 256) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = tid + 256 * k;
      if (t < nvec) {
        const uint4 a = hs[t];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = D::to_f2(aw[j]);
          ss = fmaf(f.x, f.x, ss);
          ss = fmaf(f.y, f.y, ss);
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
  }
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = rsqrtf(__fadd_rn(__fdiv_rn(tot, static_cast<float>(cols)), eps));
  if (act) {
    const uint4 wv = reinterpret_cast<const uint4*>(w)[tid];
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(&h);
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {                         // weight * (h * rsqrt(var + eps)), both products in fp32 (K4)
      const float2 f = D::to_f2(hw[j]), g = D::to_f2(ww[j]);
      ow[j] = D::pack2(__fmul_rn(g.x, __fmul_rn(f.x, inv)), __fmul_rn(g.y, __fmul_rn(f.y, inv)));
    }
    reinterpret_cast<uint4*>(norm_out + roff)[tid] = o;
  }
  if (tid == 0) epoch_p[row] = e;
}

// ---------------------------------------------------------------- K7c: all-reduce through the switch (NVLS multicast)
// The row-parallel all-reduce at prefill sizes when the symmetric buffer has a MULTICAST mapping (NVSwitch, NVLink 5):
// every rank's GEMM wrote its partial [M, N] into its own copy of the buffer; rank r then owns the token rows
// [r * rows_per, (r + 1) * rows_per):
//   multimem.ld_reduce   one 16-byte load on the multicast address -- the SWITCH fetches the `world` copies, adds them
//                        (fp32 accumulation) and returns the sum: 1/world of the tensor comes in per GPU instead of
//                        (world - 1)/world;
//   multimem.st          one 16-byte store on the multicast address -- the switch writes it into every rank's copy.
// In place: an element is read and then rewritten by exactly one rank (its owner), nobody else touches it between the
// two handshakes.  Per GPU and direction the links carry about one tensor size (K7b: the scatter and the broadcast
// each carry (world - 1)/world of it out of every GPU, and the reduce re-reads `world` slots from HBM).  The sum is
// formed in the switch: every rank receives the same bits, but the order of the fp32 additions is the switch's, not
// rank order (tools/peer_scatter_check.py states the tolerance: one rounding of the 16-bit result).
template <typename T> struct MultimemOps;
template <> struct MultimemOps<__half> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
    return v;
  }
  static __device__ __forceinline__ void st(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};
template <> struct MultimemOps<__nv_bfloat16> {
  static __device__ __forceinline__ uint4 ld_reduce(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
    return v;
  }
  static __device__ __forceinline__ void st(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};

template <typename T>
__global__ void __launch_bounds__(512) k7c_multimem_allreduce(T* __restrict__ mc, uint32_t* const* __restrict__ flags,
                                                              uint32_t* __restrict__ epoch_p, int64_t n_own, int rank,
                                                              int world, uint32_t* __restrict__ err) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const uint32_t e = epoch_p[c] + 1;
  const int par = static_cast<int>(e & 1u);
  auto flag_of = [&](int barrier, int r) {
    return ((static_cast<int64_t>(barrier) * 2 + par) * world + r) * K7B_CTAS + c;
  };
  auto handshake = [&](int barrier) {
    __syncthreads();
    if (tid < world) {
      st_release_sys(flags[tid] + flag_of(barrier, rank), e);
      const uint32_t* f = flags[rank] + flag_of(barrier, tid);
      const unsigned long long t0 = gtimer();
      while (ld_acquire_sys(f) != e) {
        if (gtimer() - t0 > 30000000000ull) { atomicExch(err, 1u); break; }
      }
    }
    __syncthreads();
  };
  // (A) every peer's GEMM -- the kernel before this one in its stream -- has completed: its partial is in its copy
  handshake(0);
  const int64_t nvec = n_own >> 3;
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = c * per, v1 = min(nvec, v0 + per);
  uint4* base = reinterpret_cast<uint4*>(mc + static_cast<int64_t>(rank) * n_own);
  constexpr int U = 8;                                  // 16-byte requests in flight per thread
  for (int64_t vb = v0 + tid; vb < v1; vb += static_cast<int64_t>(blockDim.x) * U) {
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + static_cast<int64_t>(u) * blockDim.x;
      if (v < v1) q[u] = MultimemOps<T>::ld_reduce(base + v);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = vb + static_cast<int64_t>(u) * blockDim.x;
      if (v < v1) MultimemOps<T>::st(base + v, q[u]);
    }
  }
  __threadfence_system();                               // the multicast stores are ordered before this CTA's flags
  // (B) every peer has stored its rows into my copy
  handshake(1);
  if (tid == 0) epoch_p[c] = e;
}

}  // namespace q100

using namespace q100;

extern "C" int q100_allreduce_multimem(void* mc_buf, const void* peer_flags, void* epoch, int64_t rows_per, int64_t N,
                                       int dtype, int rank, int world, void* err_flag, void* stream) {
  Q100_REQUIRE(mc_buf && peer_flags && epoch && err_flag, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, Q100_EINVAL, "bad rank %d / world %d", rank, world);
  Q100_REQUIRE(rows_per > 0 && N > 0 && (rows_per * N) % 8 == 0, Q100_EINVAL, "rows_per * N must be a positive multiple of 8");
  Q100_REQUIRE((reinterpret_cast<uintptr_t>(mc_buf) & 15) == 0, Q100_EINVAL, "the multicast address must be 16-byte aligned");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n_own = rows_per * N;
  if (dtype == Q100_F16)
    k7c_multimem_allreduce<__half><<<K7B_CTAS, 512, 0, st>>>(reinterpret_cast<__half*>(mc_buf),
        reinterpret_cast<uint32_t* const*>(peer_flags), reinterpret_cast<uint32_t*>(epoch), n_own, rank, world,
        reinterpret_cast<uint32_t*>(err_flag));
  else
    k7c_multimem_allreduce<__nv_bfloat16><<<K7B_CTAS, 512, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(mc_buf),
        reinterpret_cast<uint32_t* const*>(peer_flags), reinterpret_cast<uint32_t*>(epoch), n_own, rank, world,
        reinterpret_cast<uint32_t*>(err_flag));
  return check_cuda(cudaGetLastError(), "k7c_multimem_allreduce launch");
}

extern "C" int q100_reduce_bcast(const void* stage, const void* peer_y, const void* peer_flags, void* epoch,
                                 int64_t rows_per, int64_t N, int dtype, int rank, int world, void* err_flag,
                                 void* stream) {
  Q100_REQUIRE(stage && peer_y && peer_flags && epoch && err_flag, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, Q100_EINVAL, "bad rank %d / world %d", rank, world);
  Q100_REQUIRE(rows_per > 0 && N > 0 && (rows_per * N) % 8 == 0, Q100_EINVAL, "rows_per * N must be a positive multiple of 8");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n_own = rows_per * N;
  if (dtype == Q100_F16)
    k7b_reduce_bcast<__half><<<K7B_CTAS, 256, 0, st>>>(reinterpret_cast<const __half*>(stage),
        reinterpret_cast<__half* const*>(peer_y), reinterpret_cast<uint32_t* const*>(peer_flags),
        reinterpret_cast<uint32_t*>(epoch), n_own, n_own, rank, world, reinterpret_cast<uint32_t*>(err_flag));
  else
    k7b_reduce_bcast<__nv_bfloat16><<<K7B_CTAS, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(stage),
        reinterpret_cast<__nv_bfloat16* const*>(peer_y), reinterpret_cast<uint32_t* const*>(peer_flags),
        reinterpret_cast<uint32_t*>(epoch), n_own, n_own, rank, world, reinterpret_cast<uint32_t*>(err_flag));
  return check_cuda(cudaGetLastError(), "k7b_reduce_bcast launch");
}

extern "C" int q100_allreduce_oneshot(const void* x, void* y, int64_t n, int dtype, const void* peer_bufs,
                                      const void* peer_flags, void* epoch, int64_t slot_elems, int rank, int world,
                                      void* err_flag, void* stream) {
  Q100_REQUIRE(x && y && peer_bufs && peer_flags && epoch && err_flag, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, Q100_EINVAL, "bad rank %d / world %d", rank, world);
  Q100_REQUIRE(n > 0 && n % 8 == 0 && n <= slot_elems, Q100_EUNSUPPORTED,
               "n=%lld must be a positive multiple of 8 and fit a slot of %lld elements", (long long)n, (long long)slot_elems);
  Q100_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && slot_elems % 8 == 0,
               Q100_EINVAL, "x / y must be 16-byte aligned");
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // a FIXED number of CTAs, whatever n: all per-CTA epoch counters then advance in lockstep, so every CTA of a call
  // uses the same parity set and the slot-reuse argument (file header) holds across calls of different sizes
  const int ctas = K7_CTAS;
  if (dtype == Q100_F16)
    k7_allreduce<__half><<<ctas, 256, 0, st>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), n,
        reinterpret_cast<__half* const*>(peer_bufs), reinterpret_cast<uint32_t* const*>(peer_flags),
        reinterpret_cast<uint32_t*>(epoch), slot_elems, rank, world, reinterpret_cast<uint32_t*>(err_flag));
  else
    k7_allreduce<__nv_bfloat16><<<ctas, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x),
        reinterpret_cast<__nv_bfloat16*>(y), n, reinterpret_cast<__nv_bfloat16* const*>(peer_bufs),
        reinterpret_cast<uint32_t* const*>(peer_flags), reinterpret_cast<uint32_t*>(epoch), slot_elems, rank, world,
        reinterpret_cast<uint32_t*>(err_flag));
  return check_cuda(cudaGetLastError(), "k7_allreduce launch");
}

extern "C" int q100_allreduce_add_rmsnorm(const void* x, const void* residual, const void* weight, float eps,
                                          void* sum_out, void* norm_out, int64_t rows, int64_t cols, int dtype,
                                          const void* peer_bufs, const void* peer_flags, void* epoch, int rank,
                                          int world, void* err_flag, void* stream) {
  Q100_REQUIRE(x && weight && norm_out && peer_bufs && peer_flags && epoch && err_flag, Q100_EINVAL, "NULL pointer argument");
  Q100_REQUIRE(dtype == Q100_F16 || dtype == Q100_BF16, Q100_EUNSUPPORTED, "fp16 / bf16 only (got dtype %d)", dtype);
  Q100_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, Q100_EINVAL, "bad rank %d / world %d", rank, world);
  Q100_REQUIRE(rows > 0 && rows <= K7N_MAX_ROWS && cols > 0 && cols % 8 == 0 && cols <= 8192, Q100_EUNSUPPORTED,
               "rows in [1, %d], cols %% 8 == 0, cols <= 8192 (got %lld x %lld)", K7N_MAX_ROWS, (long long)rows, (long long)cols);
  int rc = ensure_device();
  if (rc != Q100_OK) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(rows));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int nvec = static_cast<int>(cols / 8);
  const int c = static_cast<int>(cols);
  cfg.blockDim = dim3(static_cast<unsigned>(((nvec + 255) / 256) * 256));      // one thread per 16-byte vector, whole K4 "columns"
  uint32_t* ef = reinterpret_cast<uint32_t*>(err_flag);
  uint32_t* ep = reinterpret_cast<uint32_t*>(epoch);
  uint32_t* const* fl = reinterpret_cast<uint32_t* const*>(peer_flags);
  cudaError_t e;
  if (dtype == Q100_F16)
    e = cudaLaunchKernelEx(&cfg, k7n_allreduce_add_rmsnorm<__half>, reinterpret_cast<const __half*>(x),
                           reinterpret_cast<const __half*>(residual), reinterpret_cast<const __half*>(weight), eps,
                           reinterpret_cast<__half*>(sum_out), reinterpret_cast<__half*>(norm_out),
                           reinterpret_cast<__half* const*>(peer_bufs), fl, ep, c, rank, world, ef);
  else
    e = cudaLaunchKernelEx(&cfg, k7n_allreduce_add_rmsnorm<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(x),
                           reinterpret_cast<const __nv_bfloat16*>(residual), reinterpret_cast<const __nv_bfloat16*>(weight),
                           eps, reinterpret_cast<__nv_bfloat16*>(sum_out), reinterpret_cast<__nv_bfloat16*>(norm_out),
                           reinterpret_cast<__nv_bfloat16* const*>(peer_bufs), fl, ep, c, rank, world, ef);
  return check_cuda(e, "k7n_allreduce_add_rmsnorm launch");
}
