// In-CTA Cholesky of one 128x128 tile held in REGISTERS, one column per barrier step: the log-determinant variant the
// one-CTA Hellinger kernels use (hellinger.cu).  The scoring path's diagonal blocks (factor + inverse) use the blocked
// version in tile_potrf_blocked.cuh.
//
// 256 threads form a 16x16 grid; thread (ty, tx) owns the 8x8 elements
// (i, c) = (ty + 16 p, tx + 16 q), p, q in [0, 8) -- a 2-D cyclic distribution, so the shrinking
// trailing matrix stays balanced and every register index is static.  Only the pivot column travels through shared
// memory, one __syncthreads per elimination step.
//
// Right-looking Cholesky with deferred column scaling, columns left unscaled (the pivots are all the callers need):
//   a_ic -= a_ij * a_cj / a_jj   (j < c <= i)
//
// Replaces np.linalg.cholesky inside chol_safe (src/autoks/distance/distance.py:300-312) for Gram matrices of up to 128
// points.
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int TP_THREADS = 256;
constexpr int TP_R = 8;   // elements per thread per dimension (16 * 8 = 128)

struct TilePotrfSmem {
  double colbuf[2][NB];   // double-buffered pivot column
  double piv[NB];         // pivots a_jj (= L_jj^2)
  double ipiv[2];         // double-buffered reciprocal of the current pivot
  int fail;               // 1-based index of the first non-positive / non-finite pivot
};

// On exit S.piv[j] keeps the pivots (= L_jj^2) for the caller's log-determinant; S.fail != 0 => not PD.
// `nact` (<= NB) is the order of the active leading block: rows / columns >= nact are identity padding and
// their elimination steps are skipped (their pivots read as 1).
__device__ __forceinline__ void tile_potrf_pivots(double (&a)[TP_R][TP_R], TilePotrfSmem& S, int nact) {
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
}

}  // namespace b200gp
