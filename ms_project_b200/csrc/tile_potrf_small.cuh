// Pivots of the Cholesky factorisation of ONE small SPD matrix (order <= 128) held in REGISTERS by one CTA: the
// log-determinant engine of the Hellinger pair kernels (hellinger.cu), where 2.5 million 100 x 100 factorisations make up
// configuration C3.  Three forms, in the order they were written: a 256-thread 16 x 16 grid (below; serves 112 < n <= 128),
// the look-ahead factorisation on a 128-thread 16 x 8 grid (104 < n <= 112) and on a 64-thread 8 x 8 grid (n <= 104, the
// one C3 runs: 0.102 s against 0.224 s for the first form).
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

__host__ __device__ constexpr int tr_idx(int p, int q) { return p * (p + 1) + q; }    // block (p, q), q <= 2 p + 1

// ------------------------------------------------------------------------------------------------
// Look-ahead form of the 16 x 8 factorisation (the pair kernel runs it for 104 < n <= 112; the 8 x 8 grid at the end of the
// file serves the orders below).  Per elimination step j the chain
// barrier -> pivot -> reciprocal -> scaled operands -> multiply-adds -> publish column j + 1 -> barrier is what bounds the
// kernel (ncu: FP64 pipe 37 % busy, issue slots 44 %), so everything that can leave it does:
//   * the block column that holds column j + 1 is updated FIRST, and its owners publish column j + 1 (into the other
//     buffer) before the rest of the trailing update is issued: by the time the last multiply-add of step j retires the
//     next column has long been visible, and the barrier costs its own latency only;
//   * the owner of pivot j + 1 computes the reciprocal right after that first update -- its MUFU + Newton chain runs under
//     the remaining multiply-adds -- and publishes it with the column, so after the barrier a thread loads 1 / d_j instead
//     of computing it (one reciprocal per step and matrix instead of one per thread);
//   * the ROW operands carry the 1 / d_j factor (RP multiplies per step instead of 2 RP on the column operands);
//   * the last step of a block column skips that block column (its eight columns are all eliminated).
// Same arithmetic up to the rounding of the reciprocal (MUFU.RCP64H + two Newton steps, ~1 ulp, instead of the IEEE divide).
// ------------------------------------------------------------------------------------------------
template <int RP>
struct RectLaSmem {
  double col[2][16 * RP];     // [step parity][row]: column j as it stands when step j starts
  double inv[2];              // [step parity]: 1 / d_j
  double piv[16 * RP];        // pivots d_j (1 on the padding)
};

__device__ __forceinline__ double rcp_newton(double m) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m));     // MUFU.RCP64H
  double e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Thread -> (ty, tx) of the look-ahead kernel.  The shared-memory pipe is what bounds the factorisation (ncu: 90 % of its
// wavefront peak with the plain ty = tid / 8, tx = tid % 8 map), and tools/smem_probe.cu measured what a 64-bit warp access
// costs on sm_100a: a LOAD is one wavefront if every aligned quad of lanes touches at most two distinct 8-byte words and
// two otherwise; a STORE is one wavefront per half warp with an active lane.  So: a quad is 2 ty x 2 tx (row operands and
// column operands: two distinct words per quad each, one wavefront per load instead of two for the column operands), a
// warp is 8 ty x 4 tx (a warp-wide global load of the Gram matrices covers full 32-byte sectors), and lane bit 4 is a tx
// bit, so the eight owners of a column inside a warp sit in ONE half warp (one wavefront per published value; the plain
// map spread the 16 owners over all four warps at two wavefronts each).
__device__ __forceinline__ void rect_la_coords(int& ty, int& tx) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  tx = (lane & 1) | ((lane >> 4) << 1) | ((w & 1) << 2);
  ty = ((lane >> 1) & 7) | ((w >> 1) << 3);
}

// One elimination step: column j = 8 QB + jj, PAR = j & 1 (a template argument, so every shared-memory address of the step
// is a per-thread base plus an immediate), LAST = (jj == 7).
template <int RP, int QB, bool LAST, int PAR>
__device__ __forceinline__ void rect_la_step(double (&a)[RP * (RP + 1)], RectLaSmem<RP>& S, int jj, int ty, int tx) {
  constexpr int PB = QB >> 1;
  constexpr int QN = LAST ? QB + 1 : QB;         // block column of column j + 1
  constexpr int PN = QN >> 1;
  constexpr int QF = LAST ? QB + 2 : QB + 1;     // first block column of the remaining update
  const double* cA = S.col[PAR];
  __syncthreads();
  const double i0 = S.inv[PAR];
  double r0[RP];
#pragma unroll
  for (int p = PB; p < RP; p++) r0[p] = cA[ty + 16 * p] * i0;
  // the first LA_PRE operands of the remaining block columns are loaded in front of the publish (the compiler does not move
  // shared loads across the stores, and the multiply-adds after them should not start with a load latency)
  constexpr int LA_PRE = 2;
  double cpre[LA_PRE];
#pragma unroll
  for (int t = 0; t < LA_PRE; t++) cpre[t] = (QF + t < 2 * RP) ? cA[tx + 8 * (QF + t)] : 0.0;
  double inv_next = 0.0;
  const int txn = LAST ? 0 : jj + 1;
  if (QN < 2 * RP) {
    const double cn = cA[tx + 8 * QN];
#pragma unroll
    for (int p = PN; p < RP; p++) a[tr_idx(p, QN)] = fma(-r0[p], cn, a[tr_idx(p, QN)]);
    if (tx == txn) {                             // eight lanes of one half warp in two of the four warps (rect_la_coords)
      double* cB = &S.col[PAR ^ 1][ty];
#pragma unroll
      for (int p = PN; p < RP; p++) cB[16 * p] = a[tr_idx(p, QN)];
    }
    // every thread runs the chain on its own element (it interleaves with the multiply-adds below); only the pivot's owner
    // -- the thread with tx = txn, ty = 8 (QN & 1) + txn -- publishes the result
    inv_next = rcp_newton(a[tr_idx(PN, QN)]);
  }
#pragma unroll
  for (int q = QF; q < 2 * RP; q++) {
    const double c = (q - QF < LA_PRE) ? cpre[q - QF < LA_PRE ? q - QF : 0] : cA[tx + 8 * q];
#pragma unroll
    for (int p = (q >> 1); p < RP; p++) a[tr_idx(p, q)] = fma(-r0[p], c, a[tr_idx(p, q)]);
  }
  if (QN < 2 * RP) {
    if (tx == txn && ty == 8 * (QN & 1) + txn) {
      S.inv[PAR ^ 1] = inv_next;
      S.piv[QB * 8 + 1 + jj] = a[tr_idx(PN, QN)];   // pivot j + 1 (final: the steps after j leave column j + 1 alone)
    }
  }
}

template <int RP, int QB>
struct RectLaBlocks {
  static __device__ __forceinline__ void run(double (&a)[RP * (RP + 1)], RectLaSmem<RP>& S, int nact, int ty, int tx) {
    // steps of this block column that eliminate a live column (uniform); the pair (even, odd) is unrolled so that the
    // buffer parity is static
    const int live = nact - QB * 8;
    if (live <= 0) return;
#pragma unroll 1
    for (int jj = 0; jj < 6; jj += 2) {
      if (jj >= live) return;
      rect_la_step<RP, QB, false, 0>(a, S, jj, ty, tx);
      if (jj + 1 >= live) return;
      rect_la_step<RP, QB, false, 1>(a, S, jj + 1, ty, tx);
    }
    if (6 >= live) return;
    rect_la_step<RP, QB, false, 0>(a, S, 6, ty, tx);
    if (7 >= live) return;
    rect_la_step<RP, QB, true, 1>(a, S, 7, ty, tx);
    RectLaBlocks<RP, QB + 1>::run(a, S, nact, ty, tx);
  }
};
template <int RP>
struct RectLaBlocks<RP, 2 * RP> {
  static __device__ __forceinline__ void run(double (&)[RP * (RP + 1)], RectLaSmem<RP>&, int, int, int) {}
};

// On exit S.piv[0 .. 16 RP) holds the pivots; the caller checks them (a non-positive or non-finite pivot poisons the ones
// after it, and the pair kernel wants a flag only, so no per-step test is needed).
template <int RP>
__device__ __forceinline__ void rect_potrf_pivots_la(double (&a)[RP * (RP + 1)], RectLaSmem<RP>& S, int nact) {
  int ty, tx;
  rect_la_coords(ty, tx);
  for (int t = threadIdx.x; t < 16 * RP; t += TR_THREADS) S.piv[t] = 1.0;
  __syncthreads();
  if (threadIdx.x == 0) { S.inv[0] = rcp_newton(a[tr_idx(0, 0)]); S.piv[0] = a[tr_idx(0, 0)]; }
  if (tx == 0) {
#pragma unroll
    for (int p = 0; p < RP; p++) S.col[0][ty + 16 * p] = a[tr_idx(p, 0)];
  }
  RectLaBlocks<RP, 0>::run(a, S, nact, ty, tx);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// The look-ahead factorisation on an 8 x 8 thread grid (64 threads, two warps per matrix): thread (ty, tx) owns the
// elements (ty + 8 p, tx + 8 q), q <= p < NB8 -- NB8 (NB8 + 1) / 2 doubles, 91 for NB8 = 13 (n <= 104).  Against the 16 x 8
// grid: 8 x 8 blocks waste less on the dead part of the diagonal blocks and of the block column being eliminated (executed
// / algorithmic multiply-adds 1.40 instead of 1.88), and a thread does up to NB8 (NB8 + 1) / 2 multiply-adds per 2 NB8
// operand loads, so a matrix costs about half the instructions and half the shared-memory wavefronts.  The price is the
// register file: ~230 registers per thread, four CTAs = eight warps per SM.
// ------------------------------------------------------------------------------------------------
constexpr int TQ_THREADS = 64;

template <int NB8>
struct SquareLaSmem {
  double col[2][8 * NB8];
  double inv[2];
  double piv[8 * NB8];
};

__host__ __device__ constexpr int tq_idx(int p, int q) { return p * (p + 1) / 2 + q; }    // block (p, q), q <= p

// lane bits (low to high): tx0, ty0, ty1, ty2, tx1; warp bit: tx2 -- see rect_la_coords for the rules behind it
__device__ __forceinline__ void square_la_coords(int& ty, int& tx) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  tx = (lane & 1) | ((lane >> 4) << 1) | (w << 2);
  ty = (lane >> 1) & 7;
}

template <int NB8, int B, bool LAST, int PAR>
__device__ __forceinline__ void square_la_step(double (&a)[NB8 * (NB8 + 1) / 2], SquareLaSmem<NB8>& S, int jj, int ty, int tx) {
  constexpr int QN = LAST ? B + 1 : B;           // block column (= block row) of column j + 1
  constexpr int QF = QN + 1 > B + 1 ? QN + 1 : B + 1;   // first block column of the remaining update
  const double* cA = S.col[PAR];
  __syncthreads();
  const double i0 = S.inv[PAR];
  double r0[NB8];
#pragma unroll
  for (int p = B; p < NB8; p++) r0[p] = cA[ty + 8 * p] * i0;
  constexpr int LA_PRE = 2;
  double cpre[LA_PRE];
#pragma unroll
  for (int t = 0; t < LA_PRE; t++) cpre[t] = (QF + t < NB8) ? cA[tx + 8 * (QF + t)] : 0.0;
  double inv_next = 0.0;
  const int txn = LAST ? 0 : jj + 1;
  if (QN < NB8) {
    const double cn = cA[tx + 8 * QN];
#pragma unroll
    for (int p = QN; p < NB8; p++) a[tq_idx(p, QN)] = fma(-r0[p], cn, a[tq_idx(p, QN)]);
    if (tx == txn) {
      double* cB = &S.col[PAR ^ 1][ty];
#pragma unroll
      for (int p = QN; p < NB8; p++) cB[8 * p] = a[tq_idx(p, QN)];
    }
    inv_next = rcp_newton(a[tq_idx(QN, QN)]);
  }
#pragma unroll
  for (int q = QF; q < NB8; q++) {
    const double c = (q - QF < LA_PRE) ? cpre[q - QF < LA_PRE ? q - QF : 0] : cA[tx + 8 * q];
#pragma unroll
    for (int p = q; p < NB8; p++) a[tq_idx(p, q)] = fma(-r0[p], c, a[tq_idx(p, q)]);
  }
  if (QN < NB8) {
    if (tx == txn && ty == txn) {
      S.inv[PAR ^ 1] = inv_next;
      S.piv[B * 8 + 1 + jj] = a[tq_idx(QN, QN)];
    }
  }
}

template <int NB8, int B>
struct SquareLaBlocks {
  static __device__ __forceinline__ void run(double (&a)[NB8 * (NB8 + 1) / 2], SquareLaSmem<NB8>& S, int nact, int ty, int tx) {
    const int live = nact - B * 8;
    if (live <= 0) return;
#pragma unroll 1
    for (int jj = 0; jj < 6; jj += 2) {
      if (jj >= live) return;
      square_la_step<NB8, B, false, 0>(a, S, jj, ty, tx);
      if (jj + 1 >= live) return;
      square_la_step<NB8, B, false, 1>(a, S, jj + 1, ty, tx);
    }
    if (6 >= live) return;
    square_la_step<NB8, B, false, 0>(a, S, 6, ty, tx);
    if (7 >= live) return;
    square_la_step<NB8, B, true, 1>(a, S, 7, ty, tx);
    SquareLaBlocks<NB8, B + 1>::run(a, S, nact, ty, tx);
  }
};
template <int NB8>
struct SquareLaBlocks<NB8, NB8> {
  static __device__ __forceinline__ void run(double (&)[NB8 * (NB8 + 1) / 2], SquareLaSmem<NB8>&, int, int, int) {}
};

template <int NB8>
__device__ __forceinline__ void square_potrf_pivots_la(double (&a)[NB8 * (NB8 + 1) / 2], SquareLaSmem<NB8>& S, int nact) {
  int ty, tx;
  square_la_coords(ty, tx);
  for (int t = threadIdx.x; t < 8 * NB8; t += TQ_THREADS) S.piv[t] = 1.0;
  __syncthreads();
  if (tx == 0 && ty == 0) { S.inv[0] = rcp_newton(a[0]); S.piv[0] = a[0]; }
  if (tx == 0) {
#pragma unroll
    for (int p = 0; p < NB8; p++) S.col[0][ty + 8 * p] = a[tq_idx(p, 0)];
  }
  SquareLaBlocks<NB8, 0>::run(a, S, nact, ty, tx);
  __syncthreads();
}

}  // namespace b200gp
