// In-CTA Cholesky + triangular inverse of one 128x128 tile held in REGISTERS.
//
// 256 threads form a 16x16 grid; thread (ty, tx) owns the 8x8 elements
// (i, c) = (ty + 16 p, tx + 16 q), p, q in [0, 8) -- a 2-D cyclic distribution, so the shrinking
// trailing matrix stays balanced and every register index is static.  Only the pivot column
// (phase 1) / the finished row of the inverse (phase 2) travels through shared memory, one
// __syncthreads per elimination step.
//
// Phase 1 is the right-looking Cholesky with deferred column scaling:
//   a_ic -= a_ij * a_cj / a_jj   (j < c <= i),   L_ic = a_ic / sqrt(a_cc)
// Phase 2 forms D = L^-1 by forward elimination of the identity.
//
// Replaces LAPACK dpotrf / dtrtri on the diagonal blocks (GPy util.linalg.jitchol / pdinv;
// reference call site src/autoks/core/gp_model.py:64-65 via GPy inference).
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int TP_THREADS = 256;
constexpr int TP_R = 8;   // elements per thread per dimension (16 * 8 = 128)

struct TilePotrfSmem {
  double colbuf[2][NB];   // double-buffered pivot column / inverse row
  double piv[NB];         // pivots a_jj (= L_jj^2)
  double ipiv[2];         // double-buffered reciprocal of the current pivot
  int fail;               // 1-based index of the first non-positive / non-finite pivot
};

// Phase 1.  On exit a[p][q] holds L (lower, i >= c); entries above the diagonal are junk.
// S.piv[j] keeps the pivots (= L_jj^2) for the caller's log-determinant.  S.fail != 0 => not PD.
// `nact` (<= NB) is the order of the active leading block: rows / columns >= nact are identity padding and
// their elimination steps are skipped (their pivots read as 1).  kScale = false leaves the columns unscaled
// (enough when only the log-determinant is wanted).
template <bool kScale>
__device__ __forceinline__ void tile_potrf(double (&a)[TP_R][TP_R], TilePotrfSmem& S, int nact) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const bool tyge = ty >= tx;          // i >= c inside a diagonal register block (p == q)
  if (threadIdx.x == 0) S.fail = 0;
  if (threadIdx.x < NB) S.piv[threadIdx.x] = 1.0;
  __syncthreads();
#pragma unroll
  for (int jb = 0; jb < TP_R; jb++) {
    if (jb * 16 >= nact) break;
    for (int jj = 0; jj < 16; jj++) {
      const int j = jb * 16 + jj;
      if (j >= nact) break;
      double* cb = S.colbuf[j & 1];
      if (tx == jj) {
#pragma unroll
        for (int p = jb; p < TP_R; p++) cb[ty + 16 * p] = a[p][jb];
        if (ty == jj) {                 // pivot owner: reciprocal off everybody else's critical path
          const double piv = a[jb][jb];
          S.piv[j] = piv;
          S.ipiv[j & 1] = 1.0 / piv;
          if (!(piv > 0.0) || !(piv < 1.7e308)) { if (S.fail == 0) S.fail = j + 1; }
        }
      }
      __syncthreads();
      const double inv = S.ipiv[j & 1];
      const bool txgt = tx > jj;        // c > j inside the boundary block column (q == jb)
      double rowv[TP_R], colv[TP_R];
#pragma unroll
      for (int p = jb; p < TP_R; p++) { rowv[p] = cb[ty + 16 * p]; colv[p] = cb[tx + 16 * p] * inv; }
      // a_ic -= a_ij a_cj / a_jj over j < c <= i.  Register blocks strictly inside the trailing lower
      // triangle (jb < q < p) need no predicate; only the block column q == jb and the diagonal
      // blocks p == q do (all block indices are compile-time after unrolling).
#pragma unroll
      for (int p = jb; p < TP_R; p++) {
#pragma unroll
        for (int q = jb; q <= p; q++) {
          if (q > jb && q < p) {
            a[p][q] = fma(-rowv[p], colv[q], a[p][q]);
          } else {
            bool ok = true;
            if (q == jb) ok = ok && txgt;
            if (p == q) ok = ok && tyge;
            if (p == jb) ok = ok && (ty > jj);    // rows of block jb at or above the pivot row are dead
            if (ok) a[p][q] = fma(-rowv[p], colv[q], a[p][q]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (!kScale) return;
#pragma unroll
  for (int q = 0; q < TP_R; q++) {
    const int c = tx + 16 * q;
    const double d = sqrt(S.piv[c]);
    const double rd = 1.0 / d;          // LAPACK dpotf2 scales the column by 1/ajj as well
#pragma unroll
    for (int p = q; p < TP_R; p++) {
      const int i = ty + 16 * p;
      if (i > c) a[p][q] = a[p][q] * rd;
      else if (i == c) a[p][q] = d;
    }
  }
}

// Phase 2.  Lsm is L in shared memory, row-major with leading dimension `ldl` (odd => the
// strided column reads are conflict-free).  On exit d[p][q] holds D = L^-1 (lower; zeros above).
__device__ __forceinline__ void tile_trtri(double (&d)[TP_R][TP_R], const double* Lsm, int ldl, TilePotrfSmem& S) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int p = 0; p < TP_R; p++)
#pragma unroll
    for (int q = 0; q < TP_R; q++) d[p][q] = (ty + 16 * p == tx + 16 * q) ? 1.0 : 0.0;
#pragma unroll
  for (int jb = 0; jb < TP_R; jb++) {
    for (int jj = 0; jj < 16; jj++) {
      const int j = jb * 16 + jj;
      double* rb = S.colbuf[j & 1];
      if (ty == jj) {   // owners of row j finish it: D_jc = R_jc / L_jj, and publish it
        const double rl = 1.0 / Lsm[j * ldl + j];
#pragma unroll
        for (int q = 0; q <= jb; q++) {
          const double v = d[jb][q] * rl;
          d[jb][q] = v;
          rb[tx + 16 * q] = v;
        }
      }
      __syncthreads();
      const bool tygt = ty > jj;        // i > j inside block row p == jb
      const bool txle = tx <= jj;       // c <= j inside block column q == jb
      double lcol[TP_R], rowv[TP_R];
#pragma unroll
      for (int p = jb; p < TP_R; p++) lcol[p] = Lsm[(ty + 16 * p) * ldl + j];
#pragma unroll
      for (int q = 0; q <= jb; q++) rowv[q] = rb[tx + 16 * q];
#pragma unroll
      for (int p = jb; p < TP_R; p++) {
#pragma unroll
        for (int q = 0; q <= jb; q++) {
          if (p > jb && q < jb) {
            d[p][q] = fma(-lcol[p], rowv[q], d[p][q]);
          } else {
            bool ok = true;
            if (p == jb) ok = ok && tygt;
            if (q == jb) ok = ok && txle;
            if (ok) d[p][q] = fma(-lcol[p], rowv[q], d[p][q]);
          }
        }
      }
    }
  }
}

}  // namespace b200gp
