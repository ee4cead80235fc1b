// Microbenchmark 3: cp.async.bulk (TMA 1-D) shared-memory ring, deep, 1 producer warp + 8 consumer warps.
// Question: can the bulk-copy path keep HBM saturated when the ring is deep (>= 96 KB in flight per SM)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0,1,0,P;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(288) k(const uint8_t* __restrict__ w, uint32_t* out, long long nchunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const long long q0 = nchunks * blockIdx.x / gridDim.x, q1 = nchunks * (blockIdx.x + 1) / gridDim.x;
  if (warp == 8) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long q = q0; q < q1; ++q, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_expect(&full[s], CHUNK);
        bulk(smem + s * CHUNK, w + q * CHUNK, CHUNK, &full[s]);
      }
    }
  } else {
    uint32_t acc = 0, it = 0;
    const int g = lane >> 2, t = lane & 3;
    // conflict-free-ish fragment read: each warp reads its 1/8 of the chunk, 2 x 16 B per lane per 8 KB
    for (long long q = q0; q < q1; ++q, ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      const uint8_t* base = smem + s * CHUNK + warp * (CHUNK / 8);
#pragma unroll
      for (int j = 0; j < CHUNK / 8 / 512; ++j) {
        const uint4 v = *reinterpret_cast<const uint4*>(base + j * 512 + lane * 16);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) out[0] = acc + g + t;
  }
}
template <int STAGES, int CHUNK> void run(const uint8_t* w, uint32_t* out, long long bytes, int grid) {
  const long long n = bytes / CHUNK;
  const int smem = STAGES * CHUNK;
  cudaFuncSetAttribute(k<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  for (int i = 0; i < 3; ++i) k<STAGES, CHUNK><<<grid, 288, smem>>>(w + (i % 3) * bytes, out, n);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  const int iters = 12;
  for (int i = 0; i < iters; ++i) k<STAGES, CHUNK><<<grid, 288, smem>>>(w + (i % 3) * bytes, out, n);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= iters;
  cudaError_t err = cudaGetLastError();
  printf("bulk ring stages=%2d chunk=%5d grid=%4d (%3d KB in flight/SM) %7.1f MB: %7.1f us  %7.0f GB/s %s\n", STAGES, CHUNK, grid,
         STAGES * CHUNK * (grid / 148) / 1024, bytes / 1e6, ms * 1e3, bytes / ms / 1e6, err ? cudaGetErrorString(err) : "");
}
int main() {
  const long long big = 256ll << 20;
  uint8_t* w; uint32_t* out; cudaMalloc(&w, 3 * big); cudaMalloc(&out, 64); cudaMemset(w, 1, 3 * big);
  for (long long bytes : {29360128ll, 62390272ll, 250085376ll}) {
    run<4, 8192>(w, out, bytes, 296);
    run<8, 8192>(w, out, bytes, 296);
    run<12, 8192>(w, out, bytes, 296);
    run<24, 8192>(w, out, bytes, 148);
    run<6, 16384>(w, out, bytes, 296);
    run<24, 4096>(w, out, bytes, 296);
    run<8, 8192>(w, out, bytes, 444);
    run<6, 8192>(w, out, bytes, 592);
    run<12, 4096>(w, out, bytes, 592);
  }
  return 0;
}
