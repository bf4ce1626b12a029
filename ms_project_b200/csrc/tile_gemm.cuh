// 128x64 FP64 tile GEMM on the DMMA tensor pipe:  acc(128x64) += A(128 x K) * B(64 x K)^T
//
// Both operands are row-major with K contiguous ("NT"), which is how every blocked phase of the
// Cholesky / inverse is arranged (see chol_kernels.cu).  Blackwell's tcgen05/UMMA has no FP64 kind;
// the FP64 tensor path on sm_100a is `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4, measured 36.9 TFLOP/s
// register-resident, profiles/fp64_peaks_r01.json), so this is a warp-level MMA kernel.  The shape was
// chosen from measurements (tools/gemm_probe.cu, profiles/gemm_probe_r01.txt):
//   * 4 x 2 warps (256 threads), warp tile 32 x 32 = 4 x 4 DMMA m8n8 tiles, TWO CTAs per SM (<= 128 registers,
//     90 KB shared memory each): while one CTA sits in its chunk barrier, tile prologue or epilogue, the other
//     keeps the tensor pipe fed
//   * operands stream global -> shared through a 3-stage cp.async (LDGSTS) ring, 16 doubles of K per stage,
//     rows padded to 20 doubles so the 64-bit fragment loads are bank-conflict free
//   * SKEWED chunk barrier: the one __syncthreads per chunk sits in front of the chunk's LAST k-step, whose
//     fragments are already in registers, and the first fragments of the next chunk are fetched right behind it,
//     so 16 DMMAs per warp are queued while the block re-synchronises (+15 % over a barrier at the chunk head)
//   * [kc_lo, kc_hi): per-warp range of K-chunks that can contribute (triangular operands / diagonal output
//     tiles); outside it the warp still takes part in the copies and barriers but issues no DMMA
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int TG_WM = 4;                  // warps along M
constexpr int TG_WN = 2;                  // warps along N
constexpr int TG_THREADS = 32 * TG_WM * TG_WN;
constexpr int TG_TM = 32 * TG_WM;         // 128 = NB rows of C per CTA
constexpr int TG_TN = 32 * TG_WN;         // 64 columns of C per CTA (half a tile)
constexpr int TG_MT = 4;                  // DMMA m-tiles per warp
constexpr int TG_NT = 4;                  // DMMA n-tiles per warp
constexpr int TG_KC = 16;                 // K elements per pipeline stage
constexpr int TG_LDK = TG_KC + 4;         // padded row length in shared memory (doubles)
constexpr int TG_STAGES = 3;
constexpr int TG_STAGE_DOUBLES = (TG_TM + TG_TN) * TG_LDK;              // A + B per stage
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_DOUBLES * 8;         // 92,160 B -> two CTAs per SM
static_assert(TG_TM == NB && 2 * TG_TN == NB, "a 128x128 tile is two CTA tiles side by side");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Warp -> (wm, wn) position in the 4 x 2 warp grid.  A warp lives on SM sub-partition (warp % 4); sub-partition s
// hosts the warps (wm = s, wn = 0) and (wm = 3 - s, wn = 1), so the structural-zero skips (which grow with wm, shrink
// with wn) shorten all four DMMA pipes about evenly.
__device__ __forceinline__ void warp_coords(int warp, int& wm, int& wn) {
  const int s = warp & 3;
  wn = warp >> 2;
  wm = wn ? 3 - s : s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Issue the cp.async copies of one K-chunk (128 A rows and 64 B rows, 16 doubles each).
__device__ __forceinline__ void tg_load_stage(double* stage, const double* __restrict__ A, size_t lda,
                                              const double* __restrict__ B, size_t ldb, int koff) {
  double* As = stage;
  double* Bs = stage + TG_TM * TG_LDK;
#pragma unroll
  for (int r = 0; r < TG_TM * 8 / TG_THREADS; r++) {
    const int seg = threadIdx.x + TG_THREADS * r;   // 8 16-byte segments per row
    const int row = seg >> 3, part = (seg & 7) * 2;
    cp_async16(As + row * TG_LDK + part, A + (size_t)row * lda + koff + part);
  }
#pragma unroll
  for (int r = 0; r < TG_TN * 8 / TG_THREADS; r++) {
    const int seg = threadIdx.x + TG_THREADS * r;
    const int row = seg >> 3, part = (seg & 7) * 2;
    cp_async16(Bs + row * TG_LDK + part, B + (size_t)row * ldb + koff + part);
  }
}

// acc[mt][nt][2]: C fragment of DMMA tile (mt, nt) of this warp's 32 x 32 sub-tile.
// Element (row, col) of the CTA tile held in acc[mt][nt][e]:
//   row = wm*32 + mt*8 + (lane>>2),  col = wn*32 + nt*8 + (lane&3)*2 + e
// `load(stage, chunk)` issues the cp.async copies of K-chunk `chunk` (128 A rows, 64 B rows, 16 doubles each) into `stage`.
// One K-chunk of the skewed pipeline.  PRED = true: the warp's DMMAs are issued only inside [kc_lo, kc_hi); the
// unpredicated copy exists because the (warp-uniform) test alone costs 3.5 % of the DMMA rate (tools/gemm_probe,
// "skew+pred": 35.2 vs 36.4 TFLOP/s), so chunks that cannot be skipped run without it.
template <bool PRED, typename Loader>
__device__ __forceinline__ void tg_chunk(double (&acc)[TG_MT][TG_NT][2], double (&a)[TG_MT], double (&b)[TG_NT], Loader& load,
                                         int kc, int nchunks, double* smem, int aoff, int boff, int kc_lo, int kc_hi) {
  {   // refill the stage chunk kc-1 was read from: every warp passed the barrier of the previous iteration
    const int nk = kc + TG_STAGES - 1;
    if (nk < nchunks) load(smem + (nk % TG_STAGES) * TG_STAGE_DOUBLES, nk);
    cp_async_commit();
  }
  const double* As = smem + (kc % TG_STAGES) * TG_STAGE_DOUBLES + aoff;
  const double* Bs = smem + (kc % TG_STAGES) * TG_STAGE_DOUBLES + boff;
  const double* An = smem + ((kc + 1) % TG_STAGES) * TG_STAGE_DOUBLES + aoff;
  const double* Bn = smem + ((kc + 1) % TG_STAGES) * TG_STAGE_DOUBLES + boff;
  const bool act = !PRED || ((kc >= kc_lo) && (kc < kc_hi));   // warp-uniform: otherwise the operands are structurally zero
#pragma unroll
  for (int kk = 0; kk < TG_KC; kk += 4) {
    double a2[TG_MT], b2[TG_NT];
    if (kk + 4 < TG_KC) {
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++) a2[mt] = As[mt * 8 * TG_LDK + kk + 4];
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) b2[nt] = Bs[nt * 8 * TG_LDK + kk + 4];
    } else {
      cp_async_wait<TG_STAGES - 2>();     // chunk kc+1 has landed (this thread's copies) ...
      __syncthreads();                    // ... for every thread, and every warp is done reading chunk kc
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++) a2[mt] = An[mt * 8 * TG_LDK];
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) b2[nt] = Bn[nt * 8 * TG_LDK];
    }
    if (act) {
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < TG_MT; mt++) a[mt] = a2[mt];
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) b[nt] = b2[nt];
  }
}

// `load(stage, chunk)` issues the cp.async copies of K-chunk `chunk` (128 A rows, 64 B rows, 16 doubles each) into `stage`.
// Chunks [0, n_pred) honour the per-warp activity range [kc_lo, kc_hi); chunks from n_pred on are always active.
template <typename Loader>
__device__ __forceinline__ void tile_gemm_pipeline(double (&acc)[TG_MT][TG_NT][2], Loader load, int nchunks, double* smem,
                                                   int kc_lo = 0, int kc_hi = 1 << 30, int n_pred = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wm, wn;
  warp_coords(warp, wm, wn);
  const int g = lane >> 2, tg = lane & 3;
  const int aoff = (wm * 32 + g) * TG_LDK + tg;
  const int boff = TG_TM * TG_LDK + (wn * 32 + g) * TG_LDK + tg;

#pragma unroll
  for (int s = 0; s < TG_STAGES - 1; s++) {
    if (s < nchunks) load(smem + s * TG_STAGE_DOUBLES, s);
    cp_async_commit();
  }
  cp_async_wait<TG_STAGES - 2>();
  __syncthreads();
  double a[TG_MT], b[TG_NT];
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++) a[mt] = smem[aoff + mt * 8 * TG_LDK];
#pragma unroll
  for (int nt = 0; nt < TG_NT; nt++) b[nt] = smem[boff + nt * 8 * TG_LDK];

  const int np = n_pred < nchunks ? n_pred : nchunks;
  int kc = 0;
  for (; kc < np; kc++) tg_chunk<true>(acc, a, b, load, kc, nchunks, smem, aoff, boff, kc_lo, kc_hi);
  for (; kc < nchunks; kc++) tg_chunk<false>(acc, a, b, load, kc, nchunks, smem, aoff, boff, 0, 0);
  cp_async_wait<0>();
}

// Contiguous-K form: chunk kc starts at column 16 kc of both operands.
__device__ __forceinline__ void tile_gemm_nt(double (&acc)[TG_MT][TG_NT][2], const double* __restrict__ A, size_t lda,
                                             const double* __restrict__ B, size_t ldb, int nchunks, double* smem,
                                             int kc_lo = 0, int kc_hi = 1 << 30, int n_pred = 0) {
  tile_gemm_pipeline(acc, [=](double* stage, int chunk) { tg_load_stage(stage, A, lda, B, ldb, chunk * TG_KC); }, nchunks,
                     smem, kc_lo, kc_hi, n_pred);
}

}  // namespace b200gp
