// FP64 peak probe for B200 (sm_100a): the roofline denominators that
// MEASURED_PEAKS.json does not carry (it has HBM copy and bf16 only).
//   * DMMA m8n8k4 register-resident issue rate (FP64 tensor pipe)
//   * DFMA issue rate (FP64 ALU)
//   * FP64 exp() rate (K-build transcendental bound)
//   * cuBLAS DGEMM 8192^3 (library-achievable FP64 GEMM; measurement only)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu -lcublas
// Prints one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

template <int ILP>
__global__ void dmma_rate(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_rate(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dexp_rate(double* out, int iters) {
  double x = -1.0 - threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; it++) { s += exp(x); x -= 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dsin_rate(double* out, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; it++) { s += sin(x); x += 1e-3; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void copy_k(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024 * 4));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);

  // DMMA: warps/SM x ILP sweep, best rate reported plus table
  {
    printf(", \"dmma_tflops\": {");
    double best = 0; bool first = true;
    int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int ilp = 1; ilp <= 8; ilp *= 2) {
        dim3 g(sms), b(warps * 32);
        float ms = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaEventRecord(e0));
          if (ilp == 1) dmma_rate<1><<<g, b>>>(out, iters);
          if (ilp == 2) dmma_rate<2><<<g, b>>>(out, iters);
          if (ilp == 4) dmma_rate<4><<<g, b>>>(out, iters);
          if (ilp == 8) dmma_rate<8><<<g, b>>>(out, iters);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float t = time_ms(e0, e1); if (t < ms) ms = t;
        }
        double fl = 2.0 * 8 * 8 * 4 * (double)ilp * iters * warps * sms;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        printf("%s\"w%d_ilp%d\": %.2f", first ? "" : ", ", warps, ilp, tf); first = false;
      }
    }
    printf(", \"best\": %.2f}", best);
  }
  // DFMA
  {
    printf(", \"dfma_tflops\": {");
    double best = 0; bool first = true; int iters = 20000;
    for (int warps = 8; warps <= 32; warps *= 2) {
      dim3 g(sms), b(warps * 32);
      float ms = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0)); dfma_rate<8><<<g, b>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float t = time_ms(e0, e1); if (t < ms) ms = t;
      }
      double tf = 2.0 * 8 * iters * warps * 32.0 * sms / (ms * 1e-3) / 1e12;
      if (tf > best) best = tf;
      printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf); first = false;
    }
    printf(", \"best\": %.2f}", best);
  }
  // exp / sin
  {
    int iters = 4000; dim3 g(sms * 2), b(1024);
    float ms = 1e30f;
    for (int rep = 0; rep < 3; rep++) { CK(cudaEventRecord(e0)); dexp_rate<<<g, b>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t = time_ms(e0, e1); if (t < ms) ms = t; }
    printf(", \"dexp_gops\": %.1f", (double)iters * g.x * b.x / (ms * 1e-3) / 1e9);
    ms = 1e30f;
    for (int rep = 0; rep < 3; rep++) { CK(cudaEventRecord(e0)); dsin_rate<<<g, b>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t = time_ms(e0, e1); if (t < ms) ms = t; }
    printf(", \"dsin_gops\": %.1f", (double)iters * g.x * b.x / (ms * 1e-3) / 1e9);
  }
  // HBM copy
  {
    size_t n = (size_t)1 << 30; // 1 GiB each way
    double4 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n)); CK(cudaMemset(a, 1, n));
    float ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) { CK(cudaEventRecord(e0)); copy_k<<<sms * 16, 512>>>(a, b, n / sizeof(double4)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float t = time_ms(e0, e1); if (t < ms) ms = t; }
    printf(", \"copy_gbs\": %.1f", 2.0 * n / (ms * 1e-3) / 1e9);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  // cuBLAS DGEMM (measurement only; not used on the product path)
  {
    cublasHandle_t h; cublasCreate(&h);
    printf(", \"cublas_dgemm_tflops\": {");
    bool first = true;
    for (int n : {2048, 4096, 8192}) {
      double *A, *B, *C; size_t bytes = sizeof(double) * n * (size_t)n;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
      double one = 1.0, zero = 0.0;
      float best = 1e30f;
      for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float t = time_ms(e0, e1); if (rep > 0 && t < best) best = t;
      }
      // sustained: back-to-back for ~2 s
      int reps = (int)(2000.0f / best) + 1; if (reps > 400) reps = 400;
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float sus = time_ms(e0, e1) / reps;
      printf("%s\"n%d_burst\": %.2f, \"n%d_sustained\": %.2f", first ? "" : ", ", n, 2.0 * n * (double)n * n / (best * 1e-3) / 1e12, n, 2.0 * n * (double)n * n / (sus * 1e-3) / 1e12);
      first = false;
      CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    }
    printf("}");
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
