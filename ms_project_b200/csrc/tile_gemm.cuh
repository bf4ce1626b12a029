// 128x128 FP64 tile GEMM on the DMMA tensor pipe:  acc(128x128) += A(128 x K) * B(128 x K)^T
//
// Both operands are row-major with K contiguous ("NT"), which is how every blocked phase of the
// Cholesky / inverse is arranged (see chol_kernels.cu).  Blackwell's tcgen05/UMMA has no FP64 kind;
// the FP64 tensor path on sm_100a is `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4, measured 36.9 TFLOP/s
// register-resident, profiles/fp64_peaks_r01.json), so this is a warp-level MMA kernel:
//   * TG_WM x TG_WN warps (default 4 x 4 = 512 threads), warp tile (128/TG_WM) x 32 DMMA m8n8 tiles
//   * operands stream global -> shared through a 4-stage cp.async (LDGSTS) ring, 16 doubles of K
//     per stage, rows padded to 20 doubles so the 64-bit fragment loads are bank-conflict free
//   * one __syncthreads per stage
//   * [kc_lo, kc_hi): per-warp range of K-chunks that can contribute (triangular operands / diagonal output tiles);
//     outside it the warp still takes part in the copies and barriers but issues no DMMA
#pragma once
#include "common.cuh"

namespace b200gp {

#ifndef TG_WARPS_M
#define TG_WARPS_M 4
#endif
constexpr int TG_WM = TG_WARPS_M;         // warps along M
constexpr int TG_WN = 4;                  // warps along N
constexpr int TG_THREADS = 32 * TG_WM * TG_WN;
constexpr int TG_MT = NB / TG_WM / 8;     // DMMA m-tiles per warp
constexpr int TG_NT = NB / TG_WN / 8;     // DMMA n-tiles per warp
constexpr int TG_KC = 16;                 // K elements per pipeline stage
constexpr int TG_LDK = TG_KC + 4;         // padded row length in shared memory (doubles)
constexpr int TG_STAGES = 4;
constexpr int TG_STAGE_DOUBLES = 2 * NB * TG_LDK;                       // A + B per stage
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_DOUBLES * 8;         // 163,840 B

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Warp -> (wm, wn) position in the TG_WM x TG_WN warp grid.  A warp lives on SM sub-partition (warp % 4); the columns
// are rotated against the rows so every sub-partition hosts one warp of EACH wn and of each wm.  The structural-zero
// skips below (which depend on wn, on wm, or on wn > wm) then shorten all four DMMA pipes evenly.
__device__ __forceinline__ void warp_coords(int warp, int& wm, int& wn) {
  wm = warp / TG_WN;
  wn = (warp - wm) & (TG_WN - 1);
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Issue the cp.async copies of one K-chunk (A rows and B rows, 128 x 16 doubles each).
__device__ __forceinline__ void tg_load_stage(double* stage, const double* __restrict__ A, size_t lda,
                                              const double* __restrict__ B, size_t ldb, int koff) {
  double* As = stage;
  double* Bs = stage + NB * TG_LDK;
#pragma unroll
  for (int r = 0; r < 1024 / TG_THREADS; r++) {
    const int seg = threadIdx.x + TG_THREADS * r;   // 1024 16-byte segments per operand
    const int row = seg >> 3, part = (seg & 7) * 2;
    cp_async16(As + row * TG_LDK + part, A + (size_t)row * lda + koff + part);
    cp_async16(Bs + row * TG_LDK + part, B + (size_t)row * ldb + koff + part);
  }
}

// acc[mt][nt][2]: C fragment of DMMA tile (mt, nt) of this warp's (8 TG_MT) x (8 TG_NT) sub-tile.
// Element (row, col) of the CTA tile held in acc[mt][nt][e]:
//   row = wm*8*TG_MT + mt*8 + (lane>>2),  col = wn*8*TG_NT + nt*8 + (lane&3)*2 + e
__device__ __forceinline__ void tile_gemm_nt(double (&acc)[TG_MT][TG_NT][2], const double* __restrict__ A, size_t lda,
                                             const double* __restrict__ B, size_t ldb, int nchunks, double* smem,
                                             int kc_lo = 0, int kc_hi = 1 << 30) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wm, wn;
  warp_coords(warp, wm, wn);
  const int g = lane >> 2, tg = lane & 3;

#pragma unroll
  for (int s = 0; s < TG_STAGES - 1; s++) {
    if (s < nchunks) tg_load_stage(smem + s * TG_STAGE_DOUBLES, A, lda, B, ldb, s * TG_KC);
    cp_async_commit();
  }
  for (int kc = 0; kc < nchunks; kc++) {
    cp_async_wait<TG_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kc + TG_STAGES - 1;
      if (nk < nchunks) tg_load_stage(smem + (nk % TG_STAGES) * TG_STAGE_DOUBLES, A, lda, B, ldb, nk * TG_KC);
      cp_async_commit();
    }
    const double* As = smem + (kc % TG_STAGES) * TG_STAGE_DOUBLES + (wm * 8 * TG_MT + g) * TG_LDK + tg;
    const double* Bs = smem + (kc % TG_STAGES) * TG_STAGE_DOUBLES + NB * TG_LDK + (wn * 8 * TG_NT + g) * TG_LDK + tg;
    if (kc < kc_lo || kc >= kc_hi) continue;     // warp-uniform: this warp's operands are structurally zero here
#pragma unroll
    for (int kk = 0; kk < TG_KC; kk += 4) {
      double a[TG_MT], b[TG_NT];
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++) a[mt] = As[mt * 8 * TG_LDK + kk];
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) b[nt] = Bs[nt * 8 * TG_LDK + kk];
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace b200gp
