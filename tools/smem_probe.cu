// Shared-memory wavefront cost of 64-bit warp accesses on sm_100a for the lane -> address patterns of the register
// Cholesky (tile_potrf_small.cuh): 16 warps per SM issue LDS.64 / STS.64 back to back; cycles per warp instruction x 1 SM
// = wavefronts per instruction when the shared-memory pipe is the limiter.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_probe tools/smem_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int pattern(int p, int lane) {
  switch (p) {
    case 0: return 0;                      // one address
    case 1: return lane >> 3;              // 4 distinct, groups of 8 consecutive lanes   (row operands)
    case 2: return lane & 7;               // 8 distinct, repeated 4 times                (column operands)
    case 3: return lane & 3;               // 4 distinct, repeated 8 times
    case 4: return lane >> 2;              // 8 distinct, groups of 4
    case 5: return lane >> 1;              // 16 distinct, pairs
    case 6: return lane & 15;              // 16 distinct, repeated twice
    case 7: return lane;                   // 32 distinct
    case 8: return (lane & 7) * 2;         // 8 distinct, stride 2 doubles
    case 9: return (lane & 7) * 4;         // 8 distinct, stride 4 doubles
    case 10: return ((lane & 7) >> 1) + 4 * (lane & 1) ;   // permuted 8
    case 11: return (lane & 1) + 2 * (lane >> 3);           // 8 distinct: 2 x 4
    case 14: return lane & 1;              // 2 distinct, alternating
    case 15: return (lane >> 1) & 7;       // 8 distinct in pairs, repeated twice
    case 12: return lane >> 4;             // 2 distinct, half warps
    case 13: return (lane & 7) + 8 * (lane >> 4);          // 16 distinct: 8 repeated twice per half warp
    default: return lane & 15;
  }
}

template <int MODE>   // 0: LDS.64, 1: STS.64 all lanes, 2: STS.64 4 active lanes (lane & 7 == 3), 3: LDS.128, 4: STS.64 lanes 0-15, 5: STS.64 even lanes
__global__ void k(long long* cyc, double* out, int p) {
  __shared__ __align__(16) double sm[2048 + 8];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int idx = pattern(p, lane) * (MODE == 3 ? 2 : 1);
  const unsigned a = (unsigned)__cvta_generic_to_shared(sm + idx);
  double acc = 0.0;
  unsigned long long x[16];
#pragma unroll
  for (int u = 0; u < 16; u++) x[u] = u;
  const bool act = (lane & 7) == 3;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a + u * 512 + (it & 1) * 8) : "memory"); x[u] += v; }
      else if (MODE == 1) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a + u * 512), "d"(acc)); }
      else if (MODE == 2) { if (act) asm volatile("st.shared.f64 [%0], %1;" :: "r"(a + u * 512), "d"(acc)); }
      else if (MODE == 4) { if (lane < 16) asm volatile("st.shared.f64 [%0], %1;" :: "r"(a + u * 512), "d"(acc)); }
      else if (MODE == 5) { if (!(lane & 1)) asm volatile("st.shared.f64 [%0], %1;" :: "r"(a + u * 512), "d"(acc)); }
      else { unsigned long long v, w; asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(w) : "r"(a + u * 512 + (it & 1) * 16) : "memory"); x[u] += v + w; }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
#pragma unroll
  for (int u = 0; u < 16; u++) acc += (double)x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  long long* c; double* o;
  cudaMalloc(&c, 8 * 256); cudaMalloc(&o, 8 * 512 * 256);
  const char* names[6] = {"LDS.64", "STS.64 all lanes", "STS.64 4 lanes", "LDS.128", "STS.64 lanes 0-15", "STS.64 even lanes"};
  for (int mode = 0; mode < 6; mode++)
    for (int p = 0; p < 16; p++) {
      for (int r = 0; r < 1; r++) {
        if (mode == 0) k<0><<<1, 512>>>(c, o, p);
        if (mode == 1) k<1><<<1, 512>>>(c, o, p);
        if (mode == 2) k<2><<<1, 512>>>(c, o, p);
        if (mode == 3) k<3><<<1, 512>>>(c, o, p);
        if (mode == 4) k<4><<<1, 512>>>(c, o, p);
        if (mode == 5) k<5><<<1, 512>>>(c, o, p);
      }
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("%-18s pattern %2d: %.2f cycles per warp instruction (16 warps, 1 SM)\n", names[mode], p, (double)h / (64.0 * 16 * 16));
    }
  return 0;
}
