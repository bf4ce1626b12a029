// Do the FP64 tensor pipe (DMMA) and the FP64 ALU (DFMA) share issue bandwidth on B200?  One kernel, 16 warps per SM:
// `ndmma` of them loop on DMMA, the rest on DFMA.  If the pipes are independent the two rates add up.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

__global__ void __launch_bounds__(512, 1) mixed(double* out, int iters, int ndmma_warps) {
  const int warp = threadIdx.x >> 5;
  // spread the DMMA warps over the 4 sub-partitions (warp % 4)
  const bool is_dmma = ((warp >> 2) + (warp & 3) * 4) < ndmma_warps * 1 && ((warp >> 2) < (ndmma_warps + 3 - (warp & 3)) / 4);
  double s = 0;
  if (is_dmma) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  } else {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = i;
    for (int it = 0; it < iters * 8; it++) {      // 8x more instructions: a DFMA is 64 flops/warp vs 512 for a DMMA
#pragma unroll
      for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 20000;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 512));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int nd = 0; nd <= 16; nd += 4) {
    mixed<<<sms, 512>>>(out, 100, nd);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    mixed<<<sms, 512>>>(out, iters, nd);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dmma_flops = (double)sms * nd * iters * 8 * 512.0;
    const double dfma_flops = (double)sms * (16 - nd) * iters * 8 * 8 * 64.0;
    printf("dmma warps %2d dfma warps %2d: %8.3f ms  DMMA %6.2f TF/s  DFMA %6.2f TF/s  sum %6.2f\n", nd, 16 - nd, ms,
           dmma_flops / ms / 1e9, dfma_flops / ms / 1e9, (dmma_flops + dfma_flops) / ms / 1e9);
  }
  return 0;
}
