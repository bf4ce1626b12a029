// Blocked in-CTA Cholesky + triangular inverse of one 128x128 tile: the latency-critical step of every blocked phase.
//
// tile_potrf.cuh eliminates one column per __syncthreads (128 + 128 barrier steps, ~73 us per tile on B200, which is
// the critical path of small batches: 16 tiles = 1.2 of the 3.4 ms one N = 2048 evaluation takes).  Here the tile is
// processed in 8 block steps of 16 columns:
//   phase 1 (Cholesky, right-looking over 16-column panels; the trailing matrix stays in REGISTERS, 2-D cyclic)
//     1. the owners publish the current panel (rows >= 16 jb, 16 columns) in shared memory
//     2. ONE warp factors the 16x16 diagonal block with register shuffles (no block barrier inside: ~100 cycles per
//        pivot instead of ~570) and inverts it (T = L11^-1, column per lane)
//     3. all threads: L21 = A21 T^T (a 16-term dot product per element, no substitution chain)
//     4. all threads: rank-16 update of the trailing registers from the k-major panel copy
//   phase 2 (D = L^-1): forward elimination of the identity by block rows, D_jb,: = T_jb R_jb,: then a rank-16
//     update of the block rows below, again with the residual in registers.
// 3 barriers per block step, 48 per tile instead of 256.
//
// Replaces LAPACK dpotrf / dtrtri on the diagonal blocks (GPy util.linalg.jitchol / pdinv; reference call site
// src/autoks/core/gp_model.py:64-65 via GPy's exact inference).
#pragma once
#include "common.cuh"

#ifndef TB_STAMP
#define TB_STAMP(slot)      // tools/potrf_probe.cu defines this to accumulate clock64() deltas per step
#endif

namespace b200gp {

constexpr int TB_THREADS = 256;
constexpr int TB_R = 8;            // register blocks per thread per dimension (16 x 8 = 128)
constexpr int TB_W = 16;           // panel width
constexpr int TB_PLD = TB_W + 1;   // row length of the row-major panel copy
constexpr int TB_TLD = TB_W + 2;   // row length of the T blocks (even: rows stay 16-byte aligned)

struct TileBlockedSmem {
  double pan[2][NB * TB_PLD];      // phase 1: panel, row-major [row][c], double buffered over the block steps
                                   // phase 2: pan[0] = residual block row [r][col] (ld NB)
  double pk[TB_W * NB];            // phase 1: panel, k-major [c][row] (final L) phase 2: finished block row of D [r][col]
  double T[TB_R][TB_W * TB_TLD];   // inverses of the 16x16 diagonal blocks of L (explicit zeros above the diagonal)
  double piv[NB];                  // pivots (= L_jj^2) for the log-determinant
  int fail;                        // 1-based index of the first non-positive / non-finite pivot
};
static_assert(NB * TB_PLD >= TB_W * NB, "phase 2 reuses the panel buffer for a 16 x 128 block row");

__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1 / sqrt(d) to double precision: MUFU.RSQ64H seed (relative error e ~ 2^-23) + ONE third-order step
// y (1 + e/2 + 3 e^2 / 8), residual 5/16 e^3 ~ 2^-70: 4 dependent FP64 operations instead of the 6 of two Newton steps
// (this sits on the pivot-to-pivot critical path of the factorisation).
__device__ __forceinline__ double rsqrt_full(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(y * e, p, y);
}

// One warp: Cholesky of the 16x16 block at pan[row0.., 0..16) and its inverse.
//   pk[c][row0 + r] <- L11[r][c] (zeros above the diagonal),  T[r][c] <- (L11^-1)[r][c] (zeros above the diagonal),
//   piv[row0 + j] <- pivot j,  fail <- first bad pivot.
// Lane r (and its mirror r + 16, which computes the same values and writes nothing) holds row r of the block.
__device__ __forceinline__ void warp_potrf16(TileBlockedSmem& S, double* pan, int row0, double* T) {
  const int lane = threadIdx.x & 31;
  const int r = lane & 15;
  const bool writer = lane < 16;
  double a[TB_W], rs[TB_W];
  double mypiv = 1.0;
#pragma unroll
  for (int c = 0; c < TB_W; c++) a[c] = pan[(row0 + r) * TB_PLD + c];
#pragma unroll
  for (int j = 0; j < TB_W; j++) {
    const double d = shfl_f64(a[j], j);                    // pivot, final after the updates of columns < j
    mypiv = (r == j) ? d : mypiv;
    const double y = rsqrt_full(d);
    rs[j] = y;
    const double l = a[j] * y;                             // L_rj (lane j: d / sqrt d = L_jj)
    a[j] = l;
#pragma unroll
    for (int c = j + 1; c < TB_W; c++) {
      const double lc = shfl_f64(l, c);                    // L_cj
      a[c] = fma(-l, lc, a[c]);                            // entries above the diagonal (c > r) are junk, never read
    }
  }
  TB_STAMP(4)
  {   // pivots for the log-determinant; first non-positive / non-finite pivot (branch-free inside the pivot loop)
    const bool bad = !(mypiv > 0.0) || !(mypiv < 1.7e308);
    const unsigned m = __ballot_sync(0xffffffffu, bad) & 0xffffu;
    if (writer) S.piv[row0 + r] = mypiv;
    if (lane == 0 && m != 0u && S.fail == 0) S.fail = row0 + __ffs(m);
  }
  if (writer) {
#pragma unroll
    for (int c = 0; c < TB_W; c++) {
      const double v = (c <= r) ? a[c] : 0.0;
      S.pk[c * NB + row0 + r] = v;                 // k-major copy: operand of the rank-16 update
      pan[(row0 + r) * TB_PLD + c] = v;            // row-major copy: the owners read their final L values from it
    }
  }
  __syncwarp();
  TB_STAMP(5)
  // T = L11^-1, column c per lane (outer-product form of the forward substitution: the dependent chain is one
  // multiply + one FMA per row)
  const int cc = r;
  double s[TB_W];
#pragma unroll
  for (int k = 0; k < TB_W; k++) s[k] = 0.0;
#pragma unroll
  for (int k = 0; k < TB_W; k++) {
    const double tk = (k < cc) ? 0.0 : ((k == cc) ? rs[k] : -rs[k] * s[k]);
    if (writer) T[k * TB_TLD + cc] = tk;
#pragma unroll
    for (int rr = k + 1; rr < TB_W; rr++) s[rr] = fma(S.pk[k * NB + row0 + rr], tk, s[rr]);
  }
  if (writer) {             // pad columns of the T rows
    T[cc * TB_TLD + TB_W] = 0.0;
    T[cc * TB_TLD + TB_W + 1] = 0.0;
  }
}

// Steps 1 and 4 of phase 1 index the register blocks statically, so they are templates on the block step and are
// dispatched through a switch inside a ROLLED loop over the block steps: the one-warp factorisation and the panel
// solve -- most of the instructions -- then exist once and are re-executed from the instruction cache (fully unrolled,
// the kernel was 12k straight-line instructions executed once each, and the single warp of step 2 ran at
// instruction-fetch speed: 98 us per tile under ncu).
template <int JB>
__device__ __forceinline__ void tb_publish_panel(const double (&a)[TB_R][TB_R], TileBlockedSmem& S, int ty, int tx) {
  double* pan = S.pan[JB & 1];
#pragma unroll
  for (int p = JB; p < TB_R; p++) pan[(ty + 16 * p) * TB_PLD + tx] = a[p][JB];
}

template <int JB>
__device__ __forceinline__ void tb_update_trailing(double (&a)[TB_R][TB_R], const TileBlockedSmem& S, int ty, int tx) {
#pragma unroll
  for (int p = JB; p < TB_R; p++) a[p][JB] = S.pan[JB & 1][(ty + 16 * p) * TB_PLD + tx];     // final L values of block column JB
  if (JB + 1 < TB_R) {
#pragma unroll 2
    for (int k = 0; k < TB_W; k++) {
      double rowv[TB_R], colv[TB_R];
#pragma unroll
      for (int p = JB + 1; p < TB_R; p++) {
        rowv[p] = S.pk[k * NB + ty + 16 * p];
        colv[p] = S.pk[k * NB + tx + 16 * p];
      }
#pragma unroll
      for (int p = JB + 1; p < TB_R; p++)
#pragma unroll
        for (int q = JB + 1; q <= p; q++) a[p][q] = fma(-rowv[p], colv[q], a[p][q]);
    }
  }
}

// Step 3: L21 = A21 T^T for the rows below the diagonal block: two threads per row, interleaved columns
// (c = 2 cl + h) so both carry the same load.
__device__ __forceinline__ void tb_panel_solve(TileBlockedSmem& S, double* pan, int row0, const double* T) {
  const int i = row0 + TB_W + (threadIdx.x >> 1), h = threadIdx.x & 1;
  if (i >= NB) return;
  double ar[TB_W];
#pragma unroll
  for (int k = 0; k < TB_W; k++) ar[k] = pan[i * TB_PLD + k];
  __syncwarp();          // both threads of a row have read it before either overwrites it with L
#pragma unroll
  for (int cl = 0; cl < TB_W / 2; cl++) {
    const int c = 2 * cl + h;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k <= 2 * cl + 1; k++) acc = fma(ar[k], T[c * TB_TLD + k], acc);     // T[c][k] = 0 for k > c
    S.pk[c * NB + i] = acc;
    pan[i * TB_PLD + c] = acc;
  }
}

#ifdef TB_EXPERIMENT_SMALL_CODE      // tools/potrf_probe.cu only: one case for every step (wrong numerics, timing of the code-size effect)
#define TB_DISPATCH(jb, CALL) CALL(0);
#else
#define TB_DISPATCH(jb, CALL)                                                                     \
  switch (jb) {                                                                                   \
    case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
    case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; default: CALL(7); break; \
  }
#endif

// Phase 1.  a[p][q] = element (ty + 16 p, tx + 16 q) of the tile (lower part valid on entry).  On exit a holds L
// (blocks p >= q; exact zeros above the diagonal inside the diagonal blocks, blocks p < q untouched), S.T the inverses
// of L's 16x16 diagonal blocks, S.piv the pivots, S.fail != 0 => not positive definite.
__device__ __forceinline__ void tile_potrf_blocked(double (&a)[TB_R][TB_R], TileBlockedSmem& S) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  // warp index through a shuffle: provably warp-uniform for the compiler, so the shuffles of warp_potrf16 are plain
  // SHFLs instead of WARPSYNC.COLLECTIVE sequences
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) S.fail = 0;
#pragma unroll 1
  for (int jb = 0; jb < TB_R; jb++) {
    const int row0 = jb * TB_W;
#define TB_CALL(J) tb_publish_panel<J>(a, S, ty, tx)
    TB_DISPATCH(jb, TB_CALL)                   // 1. publish the panel: column tx of block column jb, rows >= row0
#undef TB_CALL
    __syncthreads();
    TB_STAMP(0)
    if (warp == 0) warp_potrf16(S, S.pan[jb & 1], row0, S.T[jb]);      // 2. diagonal block
    TB_STAMP(1)
    __syncthreads();
    if (jb + 1 < TB_R) tb_panel_solve(S, S.pan[jb & 1], row0, S.T[jb]);      // 3.
    __syncthreads();
    TB_STAMP(2)
#define TB_CALL(J) tb_update_trailing<J>(a, S, ty, tx)
    TB_DISPATCH(jb, TB_CALL)                   // 4. final L values of block column jb, rank-16 trailing update
#undef TB_CALL
    TB_STAMP(3)
    // no barrier: the next step publishes into the OTHER panel buffer, and `pk` is rewritten only after the next
    // step's first barrier
  }
  __syncthreads();     // S.fail / S.piv visible to everybody
}

template <int JB>
__device__ __forceinline__ void tb_inverse_step(double (&d)[TB_R][TB_R], const double* Lsm, int ldl, TileBlockedSmem& S,
                                                int ty, int tx) {
  double* R = S.pan[0];      // [16][NB] residual block row
  double* Dk = S.pk;         // [16][NB] finished block row
  constexpr int row0 = JB * TB_W;
#pragma unroll
  for (int q = 0; q <= JB; q++) R[ty * NB + tx + 16 * q] = d[JB][q];
  __syncthreads();
  {   // D_jb,: = T_jb R_jb,:   (row ty of T_jb: zeros beyond column ty)
    const double* T = S.T[JB] + ty * TB_TLD;
    double acc[TB_R];
#pragma unroll
    for (int q = 0; q <= JB; q++) acc[q] = 0.0;
#pragma unroll 4
    for (int k = 0; k < TB_W; k++) {
      const double t = T[k];
#pragma unroll
      for (int q = 0; q <= JB; q++) acc[q] = fma(t, R[k * NB + tx + 16 * q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q <= JB; q++) {
      d[JB][q] = acc[q];
      Dk[ty * NB + tx + 16 * q] = acc[q];
    }
  }
  __syncthreads();
  if (JB + 1 < TB_R) {   // R_p,: -= L_p,jb D_jb,:  for the block rows below
#pragma unroll 2
    for (int k = 0; k < TB_W; k++) {
      double lv[TB_R], dv[TB_R];
#pragma unroll
      for (int p = JB + 1; p < TB_R; p++) lv[p] = Lsm[(ty + 16 * p) * ldl + row0 + k];
#pragma unroll
      for (int q = 0; q <= JB; q++) dv[q] = Dk[k * NB + tx + 16 * q];
#pragma unroll
      for (int p = JB + 1; p < TB_R; p++)
#pragma unroll
        for (int q = 0; q <= JB; q++) d[p][q] = fma(-lv[p], dv[q], d[p][q]);
    }
  }
  // no barrier: the next step writes R (last read before the barrier above) and rewrites Dk only after its own
  // first barrier, which every thread reaches after finishing this update
}

// Phase 2.  Lsm: L row-major in shared memory (leading dimension ldl, zeros above the diagonal are not required).
// On exit d[p][q] holds D = L^-1 (blocks p >= q; exact zeros above the diagonal).
__device__ __forceinline__ void tile_trtri_blocked(double (&d)[TB_R][TB_R], const double* Lsm, int ldl, TileBlockedSmem& S) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) d[p][q] = (p == q && ty == tx) ? 1.0 : 0.0;
  tb_inverse_step<0>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<1>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<2>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<3>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<4>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<5>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<6>(d, Lsm, ldl, S, ty, tx);
  tb_inverse_step<7>(d, Lsm, ldl, S, ty, tx);
}

}  // namespace b200gp
