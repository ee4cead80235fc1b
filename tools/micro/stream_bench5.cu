// Microbenchmark 5: can a decode GEMM stream its weights with NO split of a row across CTAs?
// Every CTA owns a contiguous range of rows (a multiple of 4) of a tile-major W4 matrix
// [N/128][K/64][128 rows x 32 B] and pulls (pass of <= 32 rows) x (1024 k) stages: 16 copies of
// rows*32 B (<= 1 KB) plus 8 side-band copies of rows*4 B.  Question: do such small bulk copies
// (V0: one lane issues, V1: 32 lanes issue) or 16-byte cp.async (V2) keep HBM busy, and what does
// a whole launch cost for the four decode shapes of a LLaMA3-8B block?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" :: "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_hint(uint32_t dst, const void* src, uint32_t n, uint64_t* b, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" :: "r"(dst), "l"(src), "r"(n), "r"(s32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_arrive_noinc(uint64_t* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(s32(b)) : "memory");
}

constexpr int KS = 1024;                    // k per stage
constexpr int KT = KS / 64;                 // 64-k tiles per stage
constexpr int KG = KS / 128;                // groups per stage
constexpr int PR = 32;                      // rows per pass (max)
constexpr int TPITCH = PR * 32 + 64;        // k-tile pitch in smem (odd tiles land 64 B past a 128 multiple)
constexpr int STAGE = KT * TPITCH + KG * PR * 4;   // 17408 + 1024
constexpr int SPITCH = (STAGE + 127) & ~127;

struct Geo { int N, K; };

template <int V, int STAGES>
__global__ void __launch_bounds__(288) k(const uint8_t* __restrict__ w, const uint8_t* __restrict__ meta, uint32_t* out, Geo geo) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], V == 2 ? 32 : 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int N = geo.N, K = geo.K;
  const int Q = N / 4, G = gridDim.x;
  const int qa = (int)((long long)blockIdx.x * Q / G), qb = (int)((long long)(blockIdx.x + 1) * Q / G);
  const int nq = qb - qa;
  const int np = (nq * 4 + PR - 1) / PR;
  const int ksteps = K / KS;
  const int ktiles = K / 64;
  const uint32_t ring = s32(smem);
  if (warp == 8) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t it = 0;
    for (int p = 0; p < np; ++p) {
      const int r0 = 4 * (qa + (int)((long long)p * nq / np)), r1 = 4 * (qa + (int)((long long)(p + 1) * nq / np));
      const int pr = r1 - r0;
      // the pass may straddle a 128-row tile boundary: segment A = rows [r0, ra), B = [ra, r1)
      const int ra = min(r1, (r0 / 128 + 1) * 128);
      const int na = ra - r0, nb = r1 - ra;
      for (int ks = 0; ks < ksteps; ++ks, ++it) {
        const int s = it % STAGES;
        const uint32_t st = ring + s * SPITCH;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        if (V == 2) {
          // 16-byte cp.async, all lanes: k-tile kt: granule j = (row, half): pr*2 granules
          for (int kt = 0; kt < KT; ++kt) {
            for (int j = lane; j < pr * 2; j += 32) {
              const int r = r0 + (j >> 1);
              const uint8_t* src = w + ((size_t)(r / 128) * ktiles + ks * KT + kt) * 4096 + (r % 128) * 32 + (j & 1) * 16;
              cp16(st + kt * TPITCH + j * 16, src);
            }
          }
          for (int j = lane; j < KG * pr / 4; j += 32) {
            const int gi = j / (pr / 4), q = j % (pr / 4);
            cp16(st + KT * TPITCH + gi * PR * 4 + q * 16, meta + ((size_t)(ks * KG + gi) * N + r0) * 4 + q * 16);
          }
          cp_arrive_noinc(&full[s]);
        } else {
          const uint32_t bytes = pr * 32 * KT + pr * 4 * KG;
          if (lane == 0) mbar_expect(&full[s], bytes);
          __syncwarp();
          if (V == 0) {
            if (lane == 0) {
              for (int kt = 0; kt < KT; ++kt) {
                const uint8_t* src = w + ((size_t)(r0 / 128) * ktiles + ks * KT + kt) * 4096 + (r0 % 128) * 32;
                bulk_hint(st + kt * TPITCH, src, na * 32, &full[s], pol);
                if (nb) bulk_hint(st + kt * TPITCH + na * 32, w + ((size_t)(ra / 128) * ktiles + ks * KT + kt) * 4096, nb * 32, &full[s], pol);
              }
              for (int gi = 0; gi < KG; ++gi)
                bulk_hint(st + KT * TPITCH + gi * PR * 4, meta + ((size_t)(ks * KG + gi) * N + r0) * 4, pr * 4, &full[s], pol);
            }
          } else {
            if (lane < KT) {
              const int kt = lane;
              const uint8_t* src = w + ((size_t)(r0 / 128) * ktiles + ks * KT + kt) * 4096 + (r0 % 128) * 32;
              bulk_hint(st + kt * TPITCH, src, na * 32, &full[s], pol);
              if (nb) bulk_hint(st + kt * TPITCH + na * 32, w + ((size_t)(ra / 128) * ktiles + ks * KT + kt) * 4096, nb * 32, &full[s], pol);
            } else if (lane < KT + KG) {
              const int gi = lane - KT;
              bulk_hint(st + KT * TPITCH + gi * PR * 4, meta + ((size_t)(ks * KG + gi) * N + r0) * 4, pr * 4, &full[s], pol);
            }
          }
        }
      }
    }
  } else {
    uint32_t acc = 0, it = 0;
    const int g = lane >> 2, t = lane & 3;
    const int slab = warp & 1, kq = warp >> 1;
    for (int p = 0; p < np; ++p) {
      for (int ks = 0; ks < ksteps; ++ks, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        const uint8_t* st = smem + s * SPITCH;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int gi = kq * 2 + b;
          const uint8_t* base = st + (2 * gi + (t >> 1)) * TPITCH + (slab * 16 + g) * 32 + (t & 1) * 16;
          const uint4 v0 = *reinterpret_cast<const uint4*>(base);
          const uint4 v1 = *reinterpret_cast<const uint4*>(base + 8 * 32);
          const uint32_t m0 = *reinterpret_cast<const uint32_t*>(st + KT * TPITCH + gi * PR * 4 + (slab * 16 + g) * 4);
          acc ^= v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ m0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
}

template <int V, int STAGES> void run(const uint8_t* w, const uint8_t* meta, uint32_t* out, Geo geo, int grid) {
  const size_t wbytes = (size_t)geo.N * geo.K / 2, mbytes = (size_t)geo.N * geo.K / 128 * 4;
  const int smem = STAGES * SPITCH;
  cudaFuncSetAttribute(k<V, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const size_t wstride = 160ull << 20, mstride = 16ull << 20;
  for (int i = 0; i < 3; ++i) k<V, STAGES><<<grid, 288, smem>>>(w + (i % 3) * wstride, meta + (i % 3) * mstride, out, geo);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  const int iters = 30;
  for (int i = 0; i < iters; ++i) k<V, STAGES><<<grid, 288, smem>>>(w + (i % 3) * wstride, meta + (i % 3) * mstride, out, geo);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= iters;
  cudaError_t err = cudaGetLastError();
  printf("V%d stages=%2d grid=%3d N=%5d K=%5d %6.1f MB: %6.2f us  %6.0f GB/s %s\n", V, STAGES, grid, geo.N, geo.K,
         (wbytes + mbytes) / 1e6, ms * 1e3, (wbytes + mbytes) / ms / 1e6, err ? cudaGetErrorString(err) : "");
}
int main() {
  uint8_t *w, *meta; uint32_t* out;
  cudaMalloc(&w, 3 * (160ull << 20)); cudaMalloc(&meta, 3 * (16ull << 20)); cudaMalloc(&out, 64);
  cudaMemset(w, 1, 3 * (160ull << 20)); cudaMemset(meta, 1, 3 * (16ull << 20));
  const Geo shapes[] = {{4096, 4096}, {6144, 4096}, {28672, 4096}, {4096, 14336}, {57344, 8192}};
  for (const Geo& g : shapes) {
    run<0, 10>(w, meta, out, g, 148);
    run<1, 10>(w, meta, out, g, 148);
    run<1, 6>(w, meta, out, g, 148);
    run<1, 5>(w, meta, out, g, 296);
    run<2, 10>(w, meta, out, g, 148);
    run<2, 5>(w, meta, out, g, 296);
  }
  return 0;
}
