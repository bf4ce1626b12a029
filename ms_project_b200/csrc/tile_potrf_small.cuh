// Pivots of the Cholesky factorisation of ONE small SPD matrix (order <= 16 R <= 128) held in REGISTERS by a 256-thread CTA
// (16 x 16 thread grid; the 128-thread 16 x 8 grid the pair kernel uses up to order 112 follows at the end of the file):
// the log-determinant engine of the Hellinger pair kernel (hellinger.cu), where 2.5 million 100 x 100 factorisations make
// up configuration C3.
//
// Against tile_potrf.cuh (full 128 x 128 square of 8 x 8 register blocks per thread, one CTA per SM):
//   * only the block-lower triangle is stored: R (R + 1) / 2 blocks per thread (R = 7 for n <= 112: 28 doubles instead of
//     64), so THREE CTAs fit on an SM and their barrier / shared-memory latencies overlap;
//   * TWO columns are eliminated per barrier: the owners publish columns j and j + 1 as they stand, every thread derives
//     the updated column j + 1 itself (one extra FMA per operand) and applies a rank-2 update -- half the barriers and half
//     the publish -> read round trips per matrix;
//   * no predicates: entries above the diagonal inside a diagonal block and columns already eliminated are dead (never
//     read for a live result), so updating them with whatever the formulas give is harmless.
// Right-looking, unscaled columns (only the pivots d_j = L_jj^2 are wanted):  a_ic -= a_ij a_cj / d_j.
//
// Replaces np.linalg.cholesky inside chol_safe (src/autoks/distance/distance.py:300-312) for Gram matrices of up to 128 points.
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int TS_THREADS = 256;

template <int R>
struct SmallPotrfSmem {
  double col[2][2][16 * R];   // [step parity][column j / column j + 1][row]
  double piv[16 * R];         // pivots d_j
  int fail;                   // 1-based index of the first non-positive / non-finite pivot
};

__host__ __device__ constexpr int ts_idx(int p, int q) { return p * (p + 1) / 2 + q; }    // block (p, q), q <= p

// a[ts_idx(p, q)] = element (ty + 16 p, tx + 16 q) of the matrix, blocks q <= p; rows / columns >= nact must carry the
// identity.  On exit S.piv[0 .. 16 R) holds the pivots (1 on the padding) and S.fail the first bad pivot (0 = none).
// COLS = 2: two columns per barrier (rank-2 update); COLS = 1: one column per barrier (fewer live registers).
template <int R, int COLS>
__device__ __forceinline__ void small_potrf_pivots(double (&a)[R * (R + 1) / 2], SmallPotrfSmem<R>& S, int nact) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x == 0) S.fail = 0;
  if (threadIdx.x < 16 * R) S.piv[threadIdx.x] = 1.0;
  const int nsteps = (nact + COLS - 1) / COLS;  // an odd order eliminates one identity column of the padding as well
#pragma unroll
  for (int jb = 0; jb < R; jb++) {
#pragma unroll 1
    for (int jh = 0; jh < 16 / COLS; jh++) {
      const int step = jb * (16 / COLS) + jh;
      if (step >= nsteps) break;                // uniform
      const int jj = COLS * jh, j = jb * 16 + jj;
      double* cA = S.col[step & 1][0];
      double* cB = S.col[step & 1][1];
      if (tx == jj) {
#pragma unroll
        for (int p = jb; p < R; p++) cA[ty + 16 * p] = a[ts_idx(p, jb)];
      } else if (COLS == 2 && tx == jj + 1) {
#pragma unroll
        for (int p = jb; p < R; p++) cB[ty + 16 * p] = a[ts_idx(p, jb)];
      }
      __syncthreads();
      if (COLS == 1) {
        const double d0 = cA[j];
        const double i0 = 1.0 / d0;
        if (threadIdx.x == 0) {
          S.piv[j] = d0;
          if ((!(d0 > 0.0) || !(d0 < 1.7e308)) && S.fail == 0) S.fail = j + 1;
        }
        double c0[R];
#pragma unroll
        for (int q = jb; q < R; q++) c0[q] = cA[tx + 16 * q] * i0;
#pragma unroll
        for (int p = jb; p < R; p++) {
          const double r0 = cA[ty + 16 * p];
#pragma unroll
          for (int q = jb; q <= p; q++) a[ts_idx(p, q)] = fma(-r0, c0[q], a[ts_idx(p, q)]);
        }
      } else {
        // pivots and multipliers, computed by every thread (no second barrier)
        const double d0 = cA[j], e = cA[j + 1], f = cB[j + 1];
        const double i0 = 1.0 / d0;
        const double m = e * i0;
        const double d1 = fma(-e, m, f);
        const double i1 = 1.0 / d1;
        if (threadIdx.x == 0) {
          S.piv[j] = d0;
          S.piv[j + 1] = d1;
          const bool bad0 = !(d0 > 0.0) || !(d0 < 1.7e308), bad1 = !(d1 > 0.0) || !(d1 < 1.7e308);
          if ((bad0 || bad1) && S.fail == 0) S.fail = bad0 ? j + 1 : j + 2;
        }
        double c0[R], c1[R];                      // column operands, scaled: a_cj / d_j and a'_c,j+1 / d_j+1
#pragma unroll
        for (int q = jb; q < R; q++) {
          const double u = cA[tx + 16 * q], v = cB[tx + 16 * q];
          c0[q] = u * i0;
          c1[q] = fma(-u, m, v) * i1;
        }
#pragma unroll
        for (int p = jb; p < R; p++) {
          const double r0 = cA[ty + 16 * p];
          const double r1 = fma(-r0, m, cB[ty + 16 * p]);
#pragma unroll
          for (int q = jb; q <= p; q++) a[ts_idx(p, q)] = fma(-r1, c1[q], fma(-r0, c0[q], a[ts_idx(p, q)]));
        }
      }
    }
  }
  __syncthreads();
}

}  // namespace b200gp

namespace b200gp {

// ------------------------------------------------------------------------------------------------
// The same factorisation on a 16 x 8 thread grid (128 threads, four warps): thread (ty, tx) owns the elements
// (ty + 16 p, tx + 8 q), p < RP, q < 2 RP, of which the block-lower part q <= 2 p + 1 is stored -- RP (RP + 1) register
// blocks, 56 doubles for RP = 7 (n <= 112).  Per elimination step a thread loads RP row operands and 2 RP column operands
// for up to RP (RP + 1) multiply-adds: twice the arithmetic per operand load, pivot reciprocal and barrier of the 16 x 16
// grid, i.e. half the instructions per matrix, at the price of four instead of eight warps per matrix.
// ------------------------------------------------------------------------------------------------
constexpr int TR_THREADS = 128;

template <int RP>
struct RectPotrfSmem {
  double col[2][16 * RP];     // [step parity][row]
  double piv[16 * RP];
  int fail;
};

__host__ __device__ constexpr int tr_idx(int p, int q) { return p * (p + 1) + q; }    // block (p, q), q <= 2 p + 1

template <int RP>
__device__ __forceinline__ void rect_potrf_pivots(double (&a)[RP * (RP + 1)], RectPotrfSmem<RP>& S, int nact) {
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  if (threadIdx.x == 0) S.fail = 0;
  for (int t = threadIdx.x; t < 16 * RP; t += TR_THREADS) S.piv[t] = 1.0;
#pragma unroll
  for (int qb = 0; qb < 2 * RP; qb++) {         // block column of 8 matrix columns
    const int pb = qb >> 1;                     // first block row with live elements in these columns
#pragma unroll 1
    for (int jj = 0; jj < 8; jj++) {
      const int j = qb * 8 + jj;
      if (j >= nact) break;                     // uniform
      double* cA = S.col[j & 1];
      if (tx == jj) {
#pragma unroll
        for (int p = pb; p < RP; p++) cA[ty + 16 * p] = a[tr_idx(p, qb)];
      }
      __syncthreads();
      const double d0 = cA[j];
      const double i0 = 1.0 / d0;
      if (threadIdx.x == 0) {
        S.piv[j] = d0;
        if ((!(d0 > 0.0) || !(d0 < 1.7e308)) && S.fail == 0) S.fail = j + 1;
      }
      double c0[2 * RP];
#pragma unroll
      for (int q = qb; q < 2 * RP; q++) c0[q] = cA[tx + 8 * q] * i0;
#pragma unroll
      for (int p = pb; p < RP; p++) {
        const double r0 = cA[ty + 16 * p];
#pragma unroll
        for (int q = qb; q <= 2 * p + 1; q++) a[tr_idx(p, q)] = fma(-r0, c0[q], a[tr_idx(p, q)]);
      }
    }
  }
  __syncthreads();
}

}  // namespace b200gp
