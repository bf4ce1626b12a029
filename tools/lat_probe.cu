// Dependent-issue latencies on sm_100a (one warp): DFMA, DMUL, 64-bit SHFL, MUFU.RSQ64H, LDS.64, and DFMA issue rate with
// 1/2/4/8 independent chains.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double vfma(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }
__device__ __forceinline__ double vmul(double a, double b) { double r; asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r; }
__global__ void k(double* out, long long* cyc, double x0) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x0;
  __syncthreads();
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  int n = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = vfma(x, y, 1e-9);
  t1 = clock64(); cyc[n++] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = vmul(x, y);
  t1 = clock64(); cyc[n++] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); cyc[n++] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  t1 = clock64(); cyc[n++] = t1 - t0;
  int idx = threadIdx.x & 63;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { double v = sm[idx]; idx = (int)v & 63; }
  t1 = clock64(); cyc[n++] = t1 - t0;
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { a0 = vfma(a0, y, 1e-9); a1 = vfma(a1, y, 1e-9); }
  t1 = clock64(); cyc[n++] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { a0 = vfma(a0, y, 1e-9); a1 = vfma(a1, y, 1e-9); a2 = vfma(a2, y, 1e-9); a3 = vfma(a3, y, 1e-9); }
  t1 = clock64(); cyc[n++] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { a0 = vfma(a0, y, 1e-9); a1 = vfma(a1, y, 1e-9); a2 = vfma(a2, y, 1e-9); a3 = vfma(a3, y, 1e-9);
                                  a4 = vfma(a4, y, 1e-9); a5 = vfma(a5, y, 1e-9); a6 = vfma(a6, y, 1e-9); a7 = vfma(a7, y, 1e-9); }
  t1 = clock64(); cyc[n++] = t1 - t0;
  out[threadIdx.x] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + idx;
}
int main() {
  double* o; long long* c;
  cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 64 * 8);
  for (int warps = 1; warps <= 16; warps *= 2) {
    k<<<1, 32 * warps>>>(o, c, 1.5);
    k<<<1, 32 * warps>>>(o, c, 1.5);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    printf("warps %d | cycles per op (dependent): DFMA %.1f  DMUL %.1f  SHFL64 %.1f  RSQ64H %.1f  LDS64 %.1f | per DFMA with 2 chains %.1f  4 chains %.1f  8 chains %.1f\n",
           warps, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 512.0, h[6] / 1024.0, h[7] / 2048.0);
  }
  return 0;
}
