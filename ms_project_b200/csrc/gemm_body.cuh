// Tile-GEMM bodies of the blocked phases (Cholesky / L^-T / K^-1), shared by chol_kernels.cu and by the fused
// K^-1 + gradient kernel of elem_kernels.cu.  See chol_kernels.cu for the blocked algorithm.
#pragma once
#include "kernels.cuh"
#include "tile_gemm.cuh"

namespace b200gp {

// ------------------------------------------------------------------------------------------------
// Tile GEMM phases.  A CTA owns one 128 x 64 HALF (hf = 0 left, 1 right) of a 128 x 128 tile.
// ------------------------------------------------------------------------------------------------
struct GemmOut {
  double* Cw;        // this thread's first output element (fragment layout of tile_gemm.cuh)
  size_t ld;
  double alpha;
  bool skip;         // this warp's sub-tile lies strictly above the diagonal of a diagonal tile: nothing to write
};

// Main loop of one CTA tile: acc <- beta C + A B^T over the K range of the mode.  Returns false when the slot is
// inactive (failed factorisation): nothing was computed.
template <int MODE>
__device__ __forceinline__ bool gemm_compute(const Batch& b, int step, const UpdRange& upd, int bx, int hf, double* smem,
                                             double (&acc)[TG_MT][TG_NT][2], GemmOut& o) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return false;

  double* M1 = b.M1 + slot * b.mstride;
  double* M2 = b.M2 + slot * b.mstride;
  const size_t ld = b.Np;
  const double *A, *B;
  double* C;
  size_t ldb = ld;
  int nchunks;
  double alpha = 1.0, beta = 0.0;
  bool tri_b = false;      // B is a lower-triangular 128x128 block (D_k): column n only needs k <= n
  bool tri_a = false;      // the first K tile of A is upper triangular (V_jj): row m only needs k >= m
  bool diag_c = false;     // C is a diagonal tile: only its lower half is ever read
  switch (MODE) {
    case GM_TRSM: {        // L_ik = A_ik D_k^T   (in place: launched for hf = 1, then for hf = 0, see launch_gemm)
      const int i = step + 1 + bx;
      A = M1 + (size_t)(i * NB) * ld + step * NB;
      C = const_cast<double*>(A);
      B = b.Dinv + ((size_t)slot * b.T + step) * (NB * NB);
      ldb = NB;
      nchunks = (hf + 1) * (TG_TN / TG_KC);       // columns [64 hf, 64 hf + 64) need k < 64 (hf + 1)
      tri_b = true;
      break;
    }
    case GM_UPD: {         // A_ij -= sum_{k in [k0, k0+kw)} L_ik L_jk^T, columns j in [j0, j1)
      int i, j;
      if (upd.j1 >= b.T) {
        int r, c;
        tri_decode(bx, r, c);
        i = upd.j0 + r; j = upd.j0 + c;
      } else {
        int rem = bx;
        j = upd.j0;
        while (rem >= b.T - j) { rem -= b.T - j; j++; }
        i = j + rem;
      }
      C = M1 + (size_t)(i * NB) * ld + j * NB;
      A = M1 + (size_t)(i * NB) * ld + upd.k0 * NB;
      B = M1 + (size_t)(j * NB) * ld + upd.k0 * NB;
      nchunks = upd.kw * (NB / TG_KC);
      alpha = -1.0; beta = -1.0;
      diag_c = (i == j);
      break;
    }
    case GM_VUPD: {        // S_jm (+)= sum_{i in [max(j,k0), k0+kw)} V_ji L_mi^T
      const int J = upd.k0 + upd.kw;
      const int m = upd.j0 + bx / J, j = bx % J;
      const int kb = j > upd.k0 ? j : upd.k0;
      C = M2 + (size_t)(j * NB) * ld + m * NB;
      A = M2 + (size_t)(j * NB) * ld + kb * NB;
      B = M1 + (size_t)(m * NB) * ld + kb * NB;
      nchunks = (J - kb) * (NB / TG_KC);
      beta = (j >= upd.k0) ? 0.0 : 1.0;
      tri_a = (kb == j);
      break;
    }
    case GM_VTRSM: {       // V_ji = -S_ji D_i^T   (in place, like GM_TRSM)
      const int i = step, j = bx;
      A = M2 + (size_t)(j * NB) * ld + i * NB;
      C = const_cast<double*>(A);
      B = b.Dinv + ((size_t)slot * b.T + i) * (NB * NB);
      ldb = NB;
      nchunks = (hf + 1) * (TG_TN / TG_KC);
      alpha = -1.0;
      tri_b = true;
      break;
    }
    default: {             // GM_KINV: (K^-1)_ij = sum_{k>=i} V_ik V_jk^T
      int i, j;
      tri_decode(bx, i, j);
      C = M1 + (size_t)(i * NB) * ld + j * NB;
      A = M2 + (size_t)(i * NB) * ld + i * NB;
      B = M2 + (size_t)(j * NB) * ld + i * NB;
      nchunks = (b.T - i) * (NB / TG_KC);
      tri_a = true;
      diag_c = (i == j);
      break;
    }
  }
  C += hf * TG_TN;
  B += (size_t)(hf * TG_TN) * ldb;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  int wm, wn;
  warp_coords(warp, wm, wn);
  double* Cw = C + (size_t)(wm * 32 + g) * ld + wn * 32 + tg * 2;

  // structural zeros, resolved per warp: rows [32 wm, +32), columns [64 hf + 32 wn, +32) of the 128 x 128 tile
  const int col0 = hf * TG_TN + wn * 32;
  int kc_lo = 0, kc_hi = nchunks;
  if (tri_b) kc_hi = (col0 + 32) / TG_KC;
  if (tri_a) kc_lo = (wm * 32) / TG_KC;
  if (diag_c && col0 >= wm * 32 + 32) kc_hi = 0;     // warp tile strictly above the diagonal

  if (beta != 0.0 && kc_hi > 0) {       // accumulators start at beta*C so the C read overlaps the pipeline fill
#pragma unroll
    for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) {
        const double2 v = *reinterpret_cast<const double2*>(Cw + (size_t)(mt * 8) * ld + nt * 8);
        acc[mt][nt][0] = beta * v.x;
        acc[mt][nt][1] = beta * v.y;
      }
  } else {
#pragma unroll
    for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
  }

  // chunks that can be skipped by some warp: everything for triangular B / diagonal C, the first K tile for triangular A
  const int n_pred = (tri_b || diag_c) ? nchunks : (tri_a ? NB / TG_KC : 0);
  tile_gemm_nt(acc, A, ld, B, ldb, nchunks, smem, kc_lo, kc_hi, n_pred);
  o.Cw = Cw;
  o.ld = ld;
  o.alpha = alpha;
  o.skip = diag_c && kc_hi == 0;      // strictly-upper part of a diagonal tile: never read, left untouched
  return true;
}

__device__ __forceinline__ void gemm_store(const double (&acc)[TG_MT][TG_NT][2], const GemmOut& o) {
  if (o.skip) return;
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) {
      double2 v;
      v.x = o.alpha * acc[mt][nt][0];
      v.y = o.alpha * acc[mt][nt][1];
      *reinterpret_cast<double2*>(o.Cw + (size_t)(mt * 8) * o.ld + nt * 8) = v;
    }
}

template <int MODE>
__device__ __forceinline__ void gemm_body(const Batch& b, int step, const UpdRange& upd, int bx, int hf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double acc[TG_MT][TG_NT][2];
  GemmOut o;
  if (!gemm_compute<MODE>(b, step, upd, bx, hf, reinterpret_cast<double*>(smem_raw), acc, o)) return;
  gemm_store(acc, o);
}


}  // namespace b200gp
