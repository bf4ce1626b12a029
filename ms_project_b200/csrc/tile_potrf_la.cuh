// Look-ahead in-CTA Cholesky + triangular inverse of one 128 x 128 tile, one column per barrier, everything in registers.
//
// tile_potrf_blocked.cuh (16-column panels, one warp factoring each 16 x 16 diagonal block with shuffles, explicit panel
// solves and rank-16 updates through shared memory) takes 98 k cycles per tile, 7 of its 8 warps waiting at a barrier most
// of the time (ncu: barrier 3.2 stall cycles per issue, FP64 pipe 16 %).  The Hellinger pair kernel's look-ahead
// factorisation (tile_potrf_small.cuh) does an elimination step in a few hundred cycles instead; this file extends it to
// what the scoring path needs from a diagonal tile -- L AND D = L^-1 -- by eliminating the augmented matrix [A | I]:
//
//   right-looking, unscaled:  step j subtracts m_i = a_ij / d_j times row j from every row i > j, in A (columns c > j) and
//   in W (columns c <= j), W = I at the start.  At the end A = Lt diag(d) Lt^T with Lt unit lower triangular, column j of
//   the register matrix holds Lt_ij d_j and W = Lt^-1, so
//       L_ij = a_ij / sqrt(d_j),      D = L^-1 = diag(d)^-1/2 W:   D_ic = w_ic / sqrt(d_i).
//
// 256 threads as a 16 x 16 grid: thread (ty, tx) owns A(ty + 16 p, tx + 16 q), q <= p (36 doubles), and -- TRANSPOSED
// ownership -- W(tx + 16 p, ty + 16 q), q <= p (36 doubles).  With that, the multipliers of the W update, m_i for the rows
// i = tx + 16 p, are the column operands the A update loads anyway, and row j + 1 of W belongs to the same threads as column
// j + 1 of A (tx = (j + 1) mod 16): one group of owners publishes both, next to the pivot's reciprocal, right after the block
// column / block row that holds them has been updated and before the rest of the step's multiply-adds is issued
// (look-ahead), so a barrier costs its own latency only.  Lane map by the measured shared-memory rules
// (tools/smem_probe.cu): quad = 2 ty x 2 tx, warp = 8 ty x 4 tx, lane bit 4 = a tx bit -- every operand load is one
// wavefront and the owners of a column sit in one half warp of two warps.
//
// Columns c <= j of the block column being eliminated and rows i <= j of W must keep their final values, so the column
// operand of that block is masked to the threads with tx > j mod 16 (one select per step); everything else that gets
// "updated" without need is dead (above the diagonal inside a diagonal block, or rows i <= j at columns c > j).
//
// Replaces LAPACK dpotrf / dtrtri on the diagonal blocks (GPy util.linalg.jitchol / pdinv; reference call site
// src/autoks/core/gp_model.py:64-65 via GPy's exact inference).
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int TL_THREADS = 256;
constexpr int TL_R = 8;                       // 16-row / 16-column blocks per dimension
constexpr int TL_NE = TL_R * (TL_R + 1) / 2;  // block-lower elements per thread and matrix

struct TileLaSmem {
  double col[2][NB];      // [step parity][row i]: column j of A as it stands when step j starts (rows >= j are live)
  double wrow[2][NB];     // [step parity][column c]: row j of W when step j starts (exact zeros for c > j)
  double inv[2];          // [step parity]: 1 / d_j
  double piv[NB];         // pivots d_j
  unsigned long long bar; // split barrier (mbarrier): a thread arrives once it has published, waits when it needs the next operands
  int fail;               // 1-based index of the first non-positive / non-finite pivot (0 = none)
};

__device__ __forceinline__ void tl_bar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tl_bar_arrive(unsigned bar) {       // release: this thread's loads and stores so far come first
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tl_bar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra TL_WAIT;\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__host__ __device__ constexpr int tl_idx(int p, int q) { return p * (p + 1) / 2 + q; }    // block (p, q), q <= p

// lane bits (low to high): tx0, ty0, ty1, ty2, tx1; warp bits: tx2, tx3, ty3
__device__ __forceinline__ void tile_la_coords(int& ty, int& tx) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  tx = (lane & 1) | ((lane >> 4) << 1) | ((w & 3) << 2);
  ty = ((lane >> 1) & 7) | ((w >> 2) << 3);
}

__device__ __forceinline__ double tl_rcp(double m) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m));     // MUFU.RCP64H
  double e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// 1 / sqrt(d) to double precision: MUFU.RSQ64H seed + one third-order step (as tile_potrf_blocked.cuh: rsqrt_full)
__device__ __forceinline__ double tl_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(y * e, p, y);
}

// One elimination step: column j = 16 JB + jj, PAR = j & 1, LAST = (jj == 15).
template <int JB, bool LAST, int PAR, bool SPLIT>
__device__ __forceinline__ void tile_la_step(double (&a)[TL_NE], double (&w)[TL_NE], TileLaSmem& S, int jj, int ty, int tx) {
  constexpr int QN = LAST ? JB + 1 : JB;       // block column of column j + 1 = block row of row j + 1 of W
  const double* cA = S.col[PAR];
  const double* wA = S.wrow[PAR];
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&S.bar);
  // SPLIT: phase j of the barrier completes when every thread has arrived after its step j - 1 publish (all of them arrive,
  // owners or not) -- by the time a thread gets here the others usually have, so warps drift apart by up to a step
  // instead of meeting 128 times
  if (SPLIT) tl_bar_wait(bar, PAR);           // phase j has parity j & 1
  else __syncthreads();
  const double i0 = S.inv[PAR];
  // operands: multipliers of the rows ty + 16 p (A) and tx + 16 p (W), column values of the columns tx + 16 q, row j of W
  double r0[TL_R], cq[TL_R], wr[TL_R];
#pragma unroll
  for (int p = JB; p < TL_R; p++) r0[p] = cA[ty + 16 * p] * i0;
#pragma unroll
  for (int q = JB; q < TL_R; q++) cq[q] = cA[tx + 16 * q];
#pragma unroll
  for (int q = 0; q <= JB; q++) wr[q] = wA[ty + 16 * q];
  cq[JB] = (tx > jj) ? cq[JB] : 0.0;           // columns c <= j of this block column and rows i <= j of W are final
  const int txn = LAST ? 0 : jj + 1;
  if (QN < TL_R) {
    // the block column of A and the block row of W that hold column / row j + 1 first ...
#pragma unroll
    for (int p = QN; p < TL_R; p++) a[tl_idx(p, QN)] = fma(-r0[p], cq[QN], a[tl_idx(p, QN)]);
    {
      const double mw = cq[QN] * i0;
#pragma unroll
      for (int q = 0; q <= JB; q++) w[tl_idx(QN, q)] = fma(-mw, wr[q], w[tl_idx(QN, q)]);
    }
    // ... published by their owners (8 lanes of one half warp in two warps) together with the next pivot's reciprocal
    if (tx == txn) {
      double* cB = &S.col[PAR ^ 1][ty];
#pragma unroll
      for (int p = QN; p < TL_R; p++) cB[16 * p] = a[tl_idx(p, QN)];
      double* wB = &S.wrow[PAR ^ 1][ty];
#pragma unroll
      for (int q = 0; q <= QN; q++) wB[16 * q] = w[tl_idx(QN, q)];
    }
  }
  const double inv_next = (QN < TL_R) ? tl_rcp(a[tl_idx(QN < TL_R ? QN : 0, QN < TL_R ? QN : 0)]) : 0.0;
  if (SPLIT && QN < TL_R) {
    // the pivot's owner waits for its reciprocal here (one warp, ~60 cycles); everybody arrives and goes on
    if (tx == txn && ty == txn) {
      S.inv[PAR ^ 1] = inv_next;
      S.piv[JB * 16 + 1 + jj] = a[tl_idx(QN < TL_R ? QN : 0, QN < TL_R ? QN : 0)];
    }
    tl_bar_arrive(bar);
  }
  // the rest of the trailing update of A ...
#pragma unroll
  for (int q = JB; q < TL_R; q++) {
    if (q == QN) continue;
    if (LAST && q == JB) continue;             // block column JB is finished
#pragma unroll
    for (int p = q; p < TL_R; p++) a[tl_idx(p, q)] = fma(-r0[p], cq[q], a[tl_idx(p, q)]);
  }
  // ... and of W: rows tx + 16 p > j, columns ty + 16 q <= j
#pragma unroll
  for (int p = JB; p < TL_R; p++) {
    if (p == QN) continue;
    if (LAST && p == JB) continue;             // rows of block row JB are all <= j
    const double mw = cq[p] * i0;
#pragma unroll
    for (int q = 0; q <= JB; q++) w[tl_idx(p, q)] = fma(-mw, wr[q], w[tl_idx(p, q)]);
  }
  if (!SPLIT && QN < TL_R) {                   // (the reciprocal's chain has run under the multiply-adds above)
    if (tx == txn && ty == txn) {
      S.inv[PAR ^ 1] = inv_next;
      S.piv[JB * 16 + 1 + jj] = a[tl_idx(QN < TL_R ? QN : 0, QN < TL_R ? QN : 0)];
    }
  }
}

template <int JB, bool SPLIT>
struct TileLaBlocks {
  static __device__ __forceinline__ void run(double (&a)[TL_NE], double (&w)[TL_NE], TileLaSmem& S, int ty, int tx) {
#pragma unroll 1
    for (int jj = 0; jj < 14; jj += 2) {
      tile_la_step<JB, false, 0, SPLIT>(a, w, S, jj, ty, tx);
      tile_la_step<JB, false, 1, SPLIT>(a, w, S, jj + 1, ty, tx);
    }
    tile_la_step<JB, false, 0, SPLIT>(a, w, S, 14, ty, tx);
    tile_la_step<JB, true, 1, SPLIT>(a, w, S, 15, ty, tx);
    TileLaBlocks<JB + 1, SPLIT>::run(a, w, S, ty, tx);
  }
};
template <bool SPLIT>
struct TileLaBlocks<TL_R, SPLIT> {
  static __device__ __forceinline__ void run(double (&)[TL_NE], double (&)[TL_NE], TileLaSmem&, int, int) {}
};

// a[tl_idx(p, q)] = A(ty + 16 p, tx + 16 q), q <= p, on entry (the lower triangle of an SPD tile; identity on padding).
// On exit:  a = L in the same layout (entries above the diagonal inside the diagonal blocks are junk: the caller masks
// them), w[tl_idx(p, q)] = D(tx + 16 p, ty + 16 q) with D = L^-1 (exact zeros above the diagonal), S.piv = the pivots,
// S.fail = first bad pivot (then a and w are meaningless).
template <bool SPLIT>
__device__ __forceinline__ void tile_potrf_inverse_la(double (&a)[TL_NE], double (&w)[TL_NE], TileLaSmem& S) {
  int ty, tx;
  tile_la_coords(ty, tx);
  if (SPLIT) {
    if (threadIdx.x == 0) tl_bar_init((unsigned)__cvta_generic_to_shared(&S.bar), TL_THREADS);
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < TL_R; p++)
#pragma unroll
    for (int q = 0; q <= p; q++) w[tl_idx(p, q)] = (p == q && tx == ty) ? 1.0 : 0.0;
  if (tx == 0) {
#pragma unroll
    for (int p = 0; p < TL_R; p++) S.col[0][ty + 16 * p] = a[tl_idx(p, 0)];
    S.wrow[0][ty] = (ty == 0) ? 1.0 : 0.0;     // row 0 of W = e_0 (block column 0 only: step 0 reads nothing else)
    if (ty == 0) { S.inv[0] = tl_rcp(a[0]); S.piv[0] = a[0]; }
  }
  if (SPLIT) tl_bar_arrive((unsigned)__cvta_generic_to_shared(&S.bar));      // phase 0: column 0, row 0 of W, 1 / d_0 are published
  TileLaBlocks<0, SPLIT>::run(a, w, S, ty, tx);
  __syncthreads();
  // first bad pivot (a bad pivot poisons the later ones, the earlier ones are good by definition)
  if (threadIdx.x < 32) {
    int first = NB + 1;
#pragma unroll
    for (int r = 0; r < NB / 32; r++) {
      const int j = threadIdx.x + 32 * r;
      const double pv = S.piv[j];
      if ((!(pv > 0.0) || !(pv < 1.7e308)) && j + 1 < first) first = j + 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (threadIdx.x == 0) S.fail = (first <= NB) ? first : 0;
  }
  __syncthreads();
  if (S.fail != 0) return;                     // uniform
  // scale: column tx + 16 q of L and row tx + 16 p of D share the factors 1 / sqrt(d_(tx + 16 b))
  double rs[TL_R];
#pragma unroll
  for (int b = 0; b < TL_R; b++) rs[b] = tl_rsqrt(S.piv[tx + 16 * b]);
#pragma unroll
  for (int p = 0; p < TL_R; p++)
#pragma unroll
    for (int q = 0; q <= p; q++) {
      a[tl_idx(p, q)] *= rs[q];
      w[tl_idx(p, q)] *= rs[p];
    }
}

}  // namespace b200gp
