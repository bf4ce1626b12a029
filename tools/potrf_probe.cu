// Cycle breakdown of the blocked 128x128 tile factorisation (ms_project_b200/csrc/tile_potrf_blocked.cuh): one CTA,
// clock64() deltas per step accumulated by thread 0.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/potrf_probe tools/potrf_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

// accumulators live in shared memory (a global read-modify-write per stamp would cost more than the steps measured)
#define TB_STAMP(slot) if (threadIdx.x == 0) { extern __shared__ long long tb_st_[]; long long t_ = clock64(); tb_st_[slot] += t_ - tb_st_[15]; tb_st_[15] = t_; }
#include "../ms_project_b200/csrc/tile_potrf_blocked.cuh"
using namespace b200gp;

constexpr int LDL = NB + 1;
constexpr int SMEM = 128 + NB * LDL * 8 + (int)sizeof(TileBlockedSmem) + 64;

__global__ void __launch_bounds__(TB_THREADS, 1) probe(double* A, double* Dout, long long* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  long long* g_stamp = reinterpret_cast<long long*>(smem_raw);
  double* Lsm = reinterpret_cast<double*>(smem_raw + 128);
  TileBlockedSmem& S = *reinterpret_cast<TileBlockedSmem*>(smem_raw + 128 + NB * LDL * 8);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; i++) g_stamp[i] = 0; }
  __syncthreads();
  long long t0 = clock64();
  double a[TB_R][TB_R];
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) a[p][q] = A[(ty + 16 * p) * NB + tx + 16 * q];
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) g_stamp[15] = t1;
  tile_potrf_blocked(a, S);
  long long t2 = clock64();
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) {
      const int i = ty + 16 * p, c = tx + 16 * q;
      const double v = (i >= c) ? a[p][q] : 0.0;
      A[i * NB + c] = v;
      Lsm[i * LDL + c] = v;
    }
  __syncthreads();
  long long t3 = clock64();
#ifndef TB_EXPERIMENT_SMALL_CODE
  tile_trtri_blocked(a, Lsm, LDL, S);
#endif
  __syncthreads();
  long long t4 = clock64();
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) Dout[(ty + 16 * p) * NB + tx + 16 * q] = a[p][q];
  __syncthreads();
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4;
    for (int i = 0; i < 6; i++) out[5 + i] = g_stamp[i];
    out[11] = S.fail;
  }
}

int main() {
  std::vector<double> h(NB * NB);
  for (int i = 0; i < NB; i++)
    for (int j = 0; j < NB; j++) { double d = (i - j) * 0.05; h[i * NB + j] = exp(-0.5 * d * d) + (i == j ? 0.1 : 0.0); }
  double *dA, *dD; long long* dout;
  cudaMalloc(&dA, NB * NB * 8); cudaMalloc(&dD, NB * NB * 8); cudaMalloc(&dout, 16 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpy(dA, h.data(), NB * NB * 8, cudaMemcpyHostToDevice);
    probe<<<1, TB_THREADS, SMEM>>>(dA, dD, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long o[16];
    cudaMemcpy(o, dout, 16 * 8, cudaMemcpyDeviceToHost);
    printf("rep %d cycles: load %lld  potrf %lld  store %lld  trtri %lld  storeD %lld | potrf steps (sum over 8 block steps): publish+bar %lld  warp16 %lld  bar+solve+bar %lld  update %lld | warp16: factor %lld  writeL %lld  (T = warp16 stamp) | fail %lld\n",
           rep, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11]);
  }
  std::vector<double> L(NB * NB), D(NB * NB);
  cudaMemcpy(L.data(), dA, NB * NB * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, NB * NB * 8, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0;
  for (int i = 0; i < NB; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0, t = 0;
      for (int k = 0; k <= j; k++) s += L[i * NB + k] * L[j * NB + k];
      for (int k = j; k <= i; k++) t += L[i * NB + k] * D[k * NB + j];
      e1 = fmax(e1, fabs(s - h[i * NB + j]));
      e2 = fmax(e2, fabs(t - (i == j ? 1.0 : 0.0)));
    }
  printf("max |L L^T - A| = %.3e   max |L D - I| = %.3e\n", e1, e2);
  return 0;
}
