// Microbenchmark 4: which property of the K3a ring costs bandwidth?  Variants of stream_bench3:
// split copies (4096 + 4096 + 512 side band), evict_first hint, 3 CTAs/SM, consumer compute delay.
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
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_hint(void* dst, const void* src, uint32_t n, uint64_t* b, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" :: "r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)), "l"(pol) : "memory");
}
constexpr int CHUNK = 8192, PITCH = 8832;
template <int STAGES, bool SPLIT, bool HINT>
__global__ void __launch_bounds__(288) k(const uint8_t* __restrict__ w, const uint8_t* __restrict__ meta, uint32_t* out, long long nchunks, int delay) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const long long q0 = nchunks * blockIdx.x / gridDim.x, q1 = nchunks * (blockIdx.x + 1) / gridDim.x;
  if (warp == 8) {
    if (lane == 0) {
      uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      uint32_t it = 0;
      for (long long q = q0; q < q1; ++q, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * PITCH;
        if (SPLIT) {
          mbar_expect(&full[s], CHUNK + 512);
          if (HINT) { bulk_hint(st, w + q * CHUNK, 4096, &full[s], pol); bulk_hint(st + 4160, w + q * CHUNK + 4096, 4096, &full[s], pol); }
          else { bulk(st, w + q * CHUNK, 4096, &full[s]); bulk(st + 4160, w + q * CHUNK + 4096, 4096, &full[s]); }
          bulk(st + 8320, meta + q * 512, 512, &full[s]);
        } else {
          mbar_expect(&full[s], CHUNK);
          if (HINT) bulk_hint(st, w + q * CHUNK, CHUNK, &full[s], pol); else bulk(st, w + q * CHUNK, CHUNK, &full[s]);
        }
      }
    }
  } else {
    uint32_t acc = 0, it = 0;
    for (long long q = q0; q < q1; ++q, ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      const uint8_t* base = smem + s * PITCH + warp * 512;
      const uint4 v0 = *reinterpret_cast<const uint4*>(base + lane * 16);
      const uint4 v1 = *reinterpret_cast<const uint4*>(base + 4160 + lane * 16 - (lane >= 16 ? 0 : 0));
      acc ^= v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (delay) { const long long t0 = clock64(); while (clock64() - t0 < delay) { } }
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
}
template <int STAGES, bool SPLIT, bool HINT> void run(const uint8_t* w, const uint8_t* meta, uint32_t* out, long long bytes, int grid, int delay) {
  const long long n = bytes / CHUNK;
  const int smem = STAGES * PITCH;
  cudaFuncSetAttribute(k<STAGES, SPLIT, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  for (int i = 0; i < 3; ++i) k<STAGES, SPLIT, HINT><<<grid, 288, smem>>>(w + (i % 3) * bytes, meta, out, n, delay);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  const int iters = 12;
  for (int i = 0; i < iters; ++i) k<STAGES, SPLIT, HINT><<<grid, 288, smem>>>(w + (i % 3) * bytes, meta, out, n, delay);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= iters;
  cudaError_t err = cudaGetLastError();
  printf("stages=%2d split=%d hint=%d grid=%4d delay=%4d %7.1f MB: %7.1f us  %7.0f GB/s %s\n", STAGES, SPLIT, HINT, grid, delay,
         bytes / 1e6, ms * 1e3, bytes / ms / 1e6, err ? cudaGetErrorString(err) : "");
}
int main() {
  const long long big = 256ll << 20;
  uint8_t *w, *meta; uint32_t* out; cudaMalloc(&w, 3 * big); cudaMalloc(&meta, 64 << 20); cudaMalloc(&out, 64); cudaMemset(w, 1, 3 * big); cudaMemset(meta, 1, 64 << 20);
  const long long bytes = 250085376ll;
  run<8, false, false>(w, meta, out, bytes, 444, 0);
  run<8, false, true>(w, meta, out, bytes, 444, 0);
  run<8, true, false>(w, meta, out, bytes, 444, 0);
  run<8, true, true>(w, meta, out, bytes, 444, 0);
  for (int d : {200, 400, 600, 800, 1000, 1400}) run<8, true, true>(w, meta, out, bytes, 444, d);
  for (int d : {400, 800}) run<8, true, true>(w, meta, out, bytes, 296, d);
  for (int d : {400, 800}) run<4, true, true>(w, meta, out, bytes, 592, d);
  return 0;
}
