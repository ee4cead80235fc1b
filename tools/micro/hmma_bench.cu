// Microbenchmark: legacy mma.sync.m16n8k16 (bf16) latency and throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ void mma(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CHAINS>
__global__ void k(float* out, long long* cyc, int iters) {
  unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(d[c], a, b);
  }
  long long t1 = clock64();
  float s = 0; for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS> void run(int warps, const char* tag) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<CHAINS><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<CHAINS><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = double(h) / iters;                      // cycles per iteration of one warp
  double mma_per_cycle_sm = double(CHAINS) * warps / per;
  printf("%-28s chains=%d warps/SM=%2d: %7.1f cyc/iter  -> %.3f HMMA/cyc/SM  (%.1f cyc per dependent HMMA)\n", tag, CHAINS, warps,
         per, mma_per_cycle_sm, per / 1.0);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1, "latency (1 warp, 1 chain)");
  run<2>(1, "1 warp, 2 chains");
  run<4>(1, "1 warp, 4 chains");
  run<8>(1, "1 warp, 8 chains");
  run<4>(4, "4 warps x 4 chains");
  run<8>(4, "4 warps x 8 chains");
  run<4>(8, "8 warps x 4 chains");
  run<4>(16, "16 warps x 4 chains");
  run<2>(16, "16 warps x 2 chains");
  run<1>(16, "16 warps x 1 chain");
  run<2>(32, "32 warps x 2 chains");
  return 0;
}
