// Microbenchmark 2: warp-private cp.async ring in shared memory vs register ring at higher occupancy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// ---- A: cp.async warp-private ring.  block = 8 KB; warp w: lanes' 2 x 16 B
template <int DEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ka(const uint8_t* __restrict__ w, uint32_t* out, long long nblocks) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x, cta = blockIdx.x;
  // each CTA of WARPS warps covers WARPS*16 rows of a 128-row tile: CTAs are grouped 8/WARPS per tile stream
  const int per_tile = 8 / WARPS;
  const int stream = cta / per_tile, sub = cta % per_tile, S = G / per_tile;
  long long q0 = nblocks * stream / S, q1 = nblocks * (stream + 1) / S;
  const int wrow = (sub * WARPS + warp) * 16 + g;
  const int off0 = (t >> 1) * 4096 + wrow * 32 + (t & 1) * 16, off1 = off0 + 8 * 32;
  uint8_t* ring = smem + warp * DEPTH * 1024;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  auto issue = [&](long long q, int slot) {
    const uint32_t dst = ring_u32 + slot * 1024 + lane * 16;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(w + q * 8192 + off0) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 512), "l"(w + q * 8192 + off1) : "memory");
  };
  uint32_t acc = 0;
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) {
    if (q0 + d < q1) issue(q0 + d, d);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int slot = 0;
  for (long long q = q0; q < q1; ++q) {
    const long long qn = q + DEPTH - 1;
    if (qn < q1) issue(qn, (slot + DEPTH - 1) % DEPTH);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" :: "n"(DEPTH - 1) : "memory");
    __syncwarp();
    const uint4 a = *reinterpret_cast<const uint4*>(ring + slot * 1024 + lane * 16);
    const uint4 b = *reinterpret_cast<const uint4*>(ring + slot * 1024 + 512 + lane * 16);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    __syncwarp();
    slot = (slot + 1) % DEPTH;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
// ---- B: register ring
template <int DEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) kb(const uint8_t* __restrict__ w, uint32_t* out, long long nblocks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int G = gridDim.x, cta = blockIdx.x;
  const int per_tile = 8 / WARPS;
  const int stream = cta / per_tile, sub = cta % per_tile, S = G / per_tile;
  long long q0 = nblocks * stream / S, q1 = nblocks * (stream + 1) / S;
  const int wrow = (sub * WARPS + warp) * 16 + g;
  const int off0 = (t >> 1) * 4096 + wrow * 32 + (t & 1) * 16, off1 = off0 + 8 * 32;
  uint4 ring[DEPTH][2];
  uint32_t acc = 0;
#pragma unroll
  for (int d = 0; d < DEPTH; ++d)
    if (q0 + d < q1) { ring[d][0] = ldg_stream(w + (q0 + d) * 8192 + off0); ring[d][1] = ldg_stream(w + (q0 + d) * 8192 + off1); }
  for (long long qb = q0; qb < q1; qb += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      long long q = qb + d;
      if (q < q1) {
        acc ^= ring[d][0].x ^ ring[d][0].y ^ ring[d][0].z ^ ring[d][0].w ^ ring[d][1].x ^ ring[d][1].y ^ ring[d][1].z ^ ring[d][1].w;
        long long qn = q + DEPTH;
        if (qn < q1) { ring[d][0] = ldg_stream(w + qn * 8192 + off0); ring[d][1] = ldg_stream(w + qn * 8192 + off1); }
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <typename F> void timeit(F launch, long long bytes, const char* tag) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  for (int i = 0; i < 3; ++i) launch(i);
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  const int iters = 12;
  for (int i = 0; i < iters; ++i) launch(i);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= iters;
  cudaError_t err = cudaGetLastError();
  printf("%-44s %7.1f MB: %7.1f us  %7.0f GB/s %s\n", tag, bytes / 1e6, ms * 1e3, bytes / ms / 1e6, err ? cudaGetErrorString(err) : "");
}
int main() {
  const long long big = 160ll << 20;
  uint8_t* w; uint32_t* out; cudaMalloc(&w, 3 * big); cudaMalloc(&out, 64); cudaMemset(w, 1, 3 * big);
  for (long long bytes : {29360128ll, 155189248ll}) {
    long long nb = bytes / 8192;
#define RUNA(D, W, GRID) { cudaFuncSetAttribute(ka<D, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * D * 1024); \
    timeit([&](int i) { ka<D, W><<<GRID, W * 32, W * D * 1024>>>(w + (i % 3) * bytes, out, nb); }, bytes, "cp.async ring depth=" #D " warps=" #W " grid=" #GRID); }
#define RUNB(D, W, GRID) timeit([&](int i) { kb<D, W><<<GRID, W * 32>>>(w + (i % 3) * bytes, out, nb); }, bytes, "register ring depth=" #D " warps=" #W " grid=" #GRID);
    RUNA(8, 8, 296) RUNA(16, 8, 296) RUNA(12, 8, 444) RUNA(8, 8, 592)
    RUNB(2, 8, 592) RUNB(3, 8, 592) RUNB(4, 8, 592) RUNB(2, 8, 888) RUNB(2, 8, 1184) RUNB(4, 4, 1184) RUNB(2, 4, 2368)
  }
  return 0;
}
