// Throughput of legacy mma.sync on sm_100a: IMMA.16832.U8.S8 vs HMMA.16816.F32 (per SM, per clock).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k_rate(int iters, int* out, long long* cyc) {
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  int c[8][4];
  float f[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int r = 0; r < 4; ++r) { c[i][r] = 0; f[i][r] = 0.f; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  const long long t1 = clock64();
  int s = 0; float fs = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int r = 0; r < 4; ++r) { s += c[i][r]; fs += f[i][r]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + static_cast<int>(fs);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int* out; long long* cyc;
  cudaMalloc(&out, 148 * 4 * 256 * sizeof(int));
  cudaMalloc(&cyc, 148 * 4 * sizeof(long long));
  const int iters = 2000;
  for (int kind = 0; kind < 2; ++kind)
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
      if (kind == 0) k_rate<0><<<148 * ctas, 256>>>(iters, out, cyc); else k_rate<1><<<148 * ctas, 256>>>(iters, out, cyc);
      cudaDeviceSynchronize();
      long long h[148 * 4];
      cudaMemcpy(h, cyc, 148 * ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148 * ctas; ++i) avg += h[i]; avg /= 148 * ctas;
      const double mma_per_sm = 8.0 * iters * 8 /*warps*/ * ctas;
      printf("%s ctas/SM=%d: %.0f cycles, %.3f MMA/clk/SM (%s)\n", kind == 0 ? "IMMA.16832.U8.S8" : "HMMA.16816.F32 ", ctas, avg,
             mma_per_sm / avg, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
