// 64 x 64 CTA tile GEMM on the DMMA pipe (2 x 2 warps of 32 x 32, three CTAs per SM): the shape of the blocked phases
// (chol_kernels.cu) and of the fused K^-1 + gradient kernel (elem_kernels.cu).  See chol_kernels.cu for why this shape.
#pragma once
#include "tile_gemm.cuh"

namespace b200gp {

constexpr int K64_T = 64;                                  // CTA tile edge
constexpr int K64_THREADS = 128;                           // 2 x 2 warps of 32 x 32
constexpr int K64_STAGE_DOUBLES = 2 * K64_T * TG_LDK;      // A (64 rows) + B (64 rows), 16 k each
constexpr int K64_SMEM_BYTES = TG_STAGES * K64_STAGE_DOUBLES * 8;     // 61,440 B -> three CTAs per SM

// acc (this warp's 32 x 32 part of the 64 x 64 block) += A(64 x 16 nchunks) B(64 x 16 nchunks)^T, both operands row-major with
// K contiguous from the given pointers: the skewed pipeline of tile_gemm.cuh (barrier in front of a chunk's last k-step,
// next fragments fetched behind it) on a 3-stage ring of 64 + 64 rows, no predicate anywhere.
__device__ __forceinline__ void tile64_pipeline(double (&acc)[TG_MT][TG_NT][2], const double* __restrict__ A,
                                                const double* __restrict__ B, size_t ld, int nchunks, double* smem,
                                                size_t ldb = 0) {
  if (ldb == 0) ldb = ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, tg = lane & 3;
  const int aoff = (wm * 32 + g) * TG_LDK + tg;
  const int boff = (K64_T + wn * 32 + g) * TG_LDK + tg;
  auto load = [&](int stage, int chunk) {
    double* st = smem + stage * K64_STAGE_DOUBLES;
#pragma unroll
    for (int r = 0; r < 2 * K64_T * 8 / K64_THREADS; r++) {
      const int seg = threadIdx.x + K64_THREADS * r;        // 8 16-byte segments per row, rows 0..63 = A, 64..127 = B
      const int row = seg >> 3, part = (seg & 7) * 2;
      const double* src = (row < K64_T ? A + (size_t)row * ld : B + (size_t)(row - K64_T) * ldb) + chunk * TG_KC + part;
      cp_async16(st + row * TG_LDK + part, src);
    }
  };
#pragma unroll
  for (int s = 0; s < TG_STAGES - 1; s++) {
    if (s < nchunks) load(s, s);
    cp_async_commit();
  }
  cp_async_wait<TG_STAGES - 2>();
  __syncthreads();
  double a[TG_MT], bf[TG_NT];
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++) a[mt] = smem[aoff + mt * 8 * TG_LDK];
#pragma unroll
  for (int nt = 0; nt < TG_NT; nt++) bf[nt] = smem[boff + nt * 8 * TG_LDK];
  for (int kc = 0; kc < nchunks; kc++) {
    {
      const int nk = kc + TG_STAGES - 1;
      if (nk < nchunks) load(nk % TG_STAGES, nk);
      cp_async_commit();
    }
    const double* As = smem + (kc % TG_STAGES) * K64_STAGE_DOUBLES + aoff;
    const double* Bs = smem + (kc % TG_STAGES) * K64_STAGE_DOUBLES + boff;
    const double* An = smem + ((kc + 1) % TG_STAGES) * K64_STAGE_DOUBLES + aoff;
    const double* Bn = smem + ((kc + 1) % TG_STAGES) * K64_STAGE_DOUBLES + boff;
#pragma unroll
    for (int kk = 0; kk < TG_KC; kk += 4) {
      double a2[TG_MT], b2[TG_NT];
      if (kk + 4 < TG_KC) {
#pragma unroll
        for (int mt = 0; mt < TG_MT; mt++) a2[mt] = As[mt * 8 * TG_LDK + kk + 4];
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) b2[nt] = Bs[nt * 8 * TG_LDK + kk + 4];
      } else {
        cp_async_wait<TG_STAGES - 2>();
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < TG_MT; mt++) a2[mt] = An[mt * 8 * TG_LDK];
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) b2[nt] = Bn[nt * 8 * TG_LDK];
      }
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], bf[nt]);
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++) a[mt] = a2[mt];
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) bf[nt] = b2[nt];
    }
  }
  cp_async_wait<0>();
}

// this thread's first element of the 64 x 64 block at C (fragment layout of tile_gemm.cuh on 2 x 2 warps)
__device__ __forceinline__ double* tile64_cw(double* C, size_t ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return C + (size_t)((warp >> 1) * 32 + (lane >> 2)) * ld + (warp & 1) * 32 + (lane & 3) * 2;
}

// acc <- beta C (beta = 0: zeros, C not read)
__device__ __forceinline__ void tile64_init(double (&acc)[TG_MT][TG_NT][2], const double* Cw, size_t ld, double beta) {
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) {
      double2 v = make_double2(0.0, 0.0);
      if (beta != 0.0) v = *reinterpret_cast<const double2*>(Cw + (size_t)(mt * 8) * ld + nt * 8);
      acc[mt][nt][0] = beta * v.x;
      acc[mt][nt][1] = beta * v.y;
    }
}

__device__ __forceinline__ void tile64_store(const double (&acc)[TG_MT][TG_NT][2], double* Cw, size_t ld, double alpha) {
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++)
      *reinterpret_cast<double2*>(Cw + (size_t)(mt * 8) * ld + nt * 8) = make_double2(alpha * acc[mt][nt][0], alpha * acc[mt][nt][1]);
}

}  // namespace b200gp
