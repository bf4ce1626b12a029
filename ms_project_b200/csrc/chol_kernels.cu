// Blocked FP64 Cholesky, triangular inverse and K^-1 over batches of N x N matrices.
//
// Replaces LAPACK dpotrf / dtrtri / dpotri / dpotrs as GPy's `pdinv` + `dpotrs` call them inside
// ExactGaussianInference.inference (reached from src/autoks/core/gp_model.py:64-65).
//
// Storage per slot: M1 (Np x Np row-major) holds K_y, then L in its lower tiles, finally K^-1 in its
// lower tiles; M2 holds V = L^-T in its upper tiles; Dinv holds the T inverses of L's diagonal blocks.
// Every blocked step is an "NT" tile GEMM (tile_gemm.cuh) or the in-register diagonal-block
// factorisation (tile_potrf.cuh):
//   right-looking Cholesky, step k:   POTRF(k);  L_ik = A_ik D_k^T (i>k);  A_ij -= L_ik L_jk^T (i>=j>k)
//   V = L^-T, column i:               S_ji = sum_{k=j}^{i-1} V_jk L_ik^T;  V_ji = -S_ji D_i^T   (j<i), V_ii = D_i^T
//   K^-1 (lower):                     (K^-1)_ij = sum_{k>=i} V_ik V_jk^T                            (i>=j)
#include "kernels.cuh"
#include "tile_gemm.cuh"
#include "tile_potrf_blocked.cuh"

namespace b200gp {

// ------------------------------------------------------------------------------------------------
// Diagonal block: L_kk = chol(A_kk), D_k = L_kk^-1, V_kk = D_k^T, logdet partial, PD status.
// ------------------------------------------------------------------------------------------------
constexpr int PD_LDL = NB + 1;
constexpr int PD_SMEM_BYTES = NB * PD_LDL * 8 + (int)sizeof(TileBlockedSmem) + 64;

__global__ void __launch_bounds__(TB_THREADS, 1) potrf_diag_kernel(Batch b, int k) {
  const int slot = b.slot0 + blockIdx.x;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Lsm = reinterpret_cast<double*>(smem_raw);
  TileBlockedSmem& S = *reinterpret_cast<TileBlockedSmem*>(smem_raw + NB * PD_LDL * 8);

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double* A = b.M1 + slot * b.mstride + (size_t)(k * NB) * b.Np + k * NB;
  double a[TB_R][TB_R];
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) a[p][q] = A[(size_t)(ty + 16 * p) * b.Np + tx + 16 * q];

  tile_potrf_blocked(a, S);
  if (S.fail != 0) {   // uniform: S.fail was written before the last barrier inside tile_potrf_blocked
    if (threadIdx.x == 0) b.slot_status[slot] = k * NB + S.fail;
    return;
  }
  // log-determinant partial: sum_j log(pivot_j), reduced by warp 0 in a fixed order
  if (threadIdx.x < 32) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < NB / 32; r++) s += log(S.piv[threadIdx.x + 32 * r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) b.logdet_part[slot * b.T + k] = s;
  }
  // L -> global (lower part of the tile; the strict upper part is zeroed) and -> shared for phase 2
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) {
      const int i = ty + 16 * p, c = tx + 16 * q;
      const double v = (i >= c) ? a[p][q] : 0.0;
      A[(size_t)i * b.Np + c] = v;
      Lsm[i * PD_LDL + c] = v;
    }
  __syncthreads();
  tile_trtri_blocked(a, Lsm, PD_LDL, S);
  double* Dk = b.Dinv + ((size_t)slot * b.T + k) * (NB * NB);
  __syncthreads();
  // stage D in shared memory so the transposed copy (V_kk = D^T) is written coalesced as well
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) {
      const int i = ty + 16 * p, c = tx + 16 * q;
      Dk[i * NB + c] = a[p][q];
      Lsm[i * PD_LDL + c] = a[p][q];
    }
  if (b.M2 == nullptr) return;   // factor-only use (b200gp_potrf): no inverse wanted
  double* V = b.M2 + slot * b.mstride + (size_t)(k * NB) * b.Np + k * NB;
  __syncthreads();
#pragma unroll
  for (int p = 0; p < TB_R; p++)
#pragma unroll
    for (int q = 0; q < TB_R; q++) {
      const int i = ty + 16 * p, c = tx + 16 * q;
      V[(size_t)i * b.Np + c] = Lsm[c * PD_LDL + i];
    }
}

void launch_potrf_diag(const Batch& b, int k, cudaStream_t s) {
  potrf_diag_kernel<<<b.nslots, TB_THREADS, PD_SMEM_BYTES, s>>>(b, k);
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// Tile GEMM phases.  A CTA owns one 128 x 64 HALF (hf = 0 left, 1 right) of a 128 x 128 tile.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void gemm_body(const Batch& b, int step, const UpdRange& upd, int bx, int hf) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);

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

  double acc[TG_MT][TG_NT][2];
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

  if (diag_c && kc_hi == 0) return;      // strictly-upper part of a diagonal tile: never read, left untouched
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) {
      double2 v;
      v.x = alpha * acc[mt][nt][0];
      v.y = alpha * acc[mt][nt][1];
      *reinterpret_cast<double2*>(Cw + (size_t)(mt * 8) * ld + nt * 8) = v;
    }
}

// One named kernel per blocked step, so launch lists and ncu captures tell the phases apart.
// In-place triangular solves: the right half of a tile reads all 128 columns and writes columns 64..127, the left half
// reads and writes columns 0..63 only -- so ONE CTA does the right half first, then the left half (two launches with
// sibling CTAs before: 2 x 11 us on the critical path of every tile column of a small batch).
__global__ void __launch_bounds__(TG_THREADS, 2) chol_trsm_kernel(Batch b, int step) {
  gemm_body<GM_TRSM>(b, step, UpdRange{0, 0, 0, 0}, blockIdx.x, 1);
  __syncthreads();
  gemm_body<GM_TRSM>(b, step, UpdRange{0, 0, 0, 0}, blockIdx.x, 0);
}
__global__ void __launch_bounds__(TG_THREADS, 2) chol_update_kernel(Batch b, UpdRange r) {
  gemm_body<GM_UPD>(b, r.k0, r, blockIdx.x >> 1, blockIdx.x & 1);
}
__global__ void __launch_bounds__(TG_THREADS, 2) inv_trsm_kernel(Batch b, int step) {
  gemm_body<GM_VTRSM>(b, step, UpdRange{0, 0, 0, 0}, blockIdx.x, 1);
  __syncthreads();
  gemm_body<GM_VTRSM>(b, step, UpdRange{0, 0, 0, 0}, blockIdx.x, 0);
}
__global__ void __launch_bounds__(TG_THREADS, 2) inv_update_kernel(Batch b, UpdRange r) {
  gemm_body<GM_VUPD>(b, r.k0, r, blockIdx.x >> 1, blockIdx.x & 1);
}
__global__ void __launch_bounds__(TG_THREADS, 2) kinv_kernel(Batch b) {
  gemm_body<GM_KINV>(b, 0, UpdRange{0, 0, 0, 0}, blockIdx.x >> 1, blockIdx.x & 1);
}

void launch_gemm(const Batch& b, int mode, int step, cudaStream_t s) {
  int ntiles = 0;
  switch (mode) {
    case GM_TRSM: ntiles = b.T - step - 1; break;
    case GM_VTRSM: ntiles = step; break;
    default: ntiles = tri_count(b.T); break;
  }
  if (ntiles <= 0) return;
  dim3 grid(ntiles, b.nslots);
  switch (mode) {
    case GM_TRSM:
      chol_trsm_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, step);
      break;
    case GM_VTRSM:
      inv_trsm_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, step);
      break;
    default:
      grid.x *= 2;
      kinv_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b);
      break;
  }
  g_launch_count++;
}

void launch_vupdate(const Batch& b, UpdRange r, cudaStream_t s) {
  const int j1 = r.j1 < b.T ? r.j1 : b.T;
  const int ntiles = (j1 - r.j0) * (r.k0 + r.kw);
  if (ntiles <= 0 || r.kw <= 0) return;
  dim3 grid(2 * ntiles, b.nslots);
  inv_update_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, r);
  g_launch_count++;
}

void launch_update(const Batch& b, UpdRange r, cudaStream_t s) {
  int ntiles = 0;
  for (int j = r.j0; j < r.j1 && j < b.T; j++) ntiles += b.T - j;
  if (ntiles <= 0 || r.kw <= 0) return;
  dim3 grid(2 * ntiles, b.nslots);
  chol_update_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, r);
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// z = V^T y (= L^-1 y),  alpha = V z (= L^-T z)
//
// One CTA per 128 x 128 upper tile (j, m), j <= m, of V: a single matrix already fills the GPU (136 CTAs at N = 2048),
// where a CTA per tile column left 16 CTAs walking 2048 rows each (0.5 of the 3.4 ms a lone evaluation took).  Tile
// partials go to `solve_part` and are summed in tile order, so the result does not depend on the batch around it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ztile_kernel(Batch b) {          // part[(m, j)][c] = sum_{r in tile row j} V[r][128 m + c] y[r]
  const int slot = blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  int m, j;
  tri_decode(blockIdx.x, m, j);
  __shared__ double red[256];
  __shared__ double ys[NB];
  const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
  if (threadIdx.x < NB) ys[threadIdx.x] = b.y[j * NB + threadIdx.x];
  __syncthreads();
  const double* V = b.M2 + slot * b.mstride + (size_t)(j * NB) * b.Np + m * NB + c;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 4
  for (int r = half; r < NB; r += 8) {
    s0 = fma(V[(size_t)r * b.Np], ys[r], s0);
    s1 = fma(V[(size_t)(r + 2) * b.Np], ys[r + 2], s1);
    s2 = fma(V[(size_t)(r + 4) * b.Np], ys[r + 4], s2);
    s3 = fma(V[(size_t)(r + 6) * b.Np], ys[r + 6], s3);
  }
  red[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (half == 0) b.solve_part[((size_t)slot * tri_count(b.T) + blockIdx.x) * NB + c] = red[c] + red[c + 128];
}

__global__ void __launch_bounds__(NB) zreduce_kernel(Batch b) {         // z[128 m + c] = sum_{j <= m} part[(m, j)][c]
  const int slot = blockIdx.y, m = blockIdx.x;
  if (b.slot_status[slot] != 0) return;
  const double* part = b.solve_part + ((size_t)slot * tri_count(b.T) + tri_count(m)) * NB + threadIdx.x;
  double s = 0.0;
  for (int j = 0; j <= m; j++) s += part[(size_t)j * NB];
  b.z[(size_t)slot * b.Np + m * NB + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) atile_kernel(Batch b) {          // part[(m, j)][r] = sum_c V[128 j + r][128 m + c] z[128 m + c]
  const int slot = blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  int m, j;
  tri_decode(blockIdx.x, m, j);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* z = b.z + (size_t)slot * b.Np + m * NB + lane * 4;
  const double4 zz = *reinterpret_cast<const double4*>(z);
  const double* V = b.M2 + slot * b.mstride + (size_t)(j * NB) * b.Np + m * NB + lane * 4;
  double* part = b.solve_part + ((size_t)slot * tri_count(b.T) + blockIdx.x) * NB;
#pragma unroll 4
  for (int rr = warp; rr < NB; rr += 8) {
    const double4 v = *reinterpret_cast<const double4*>(V + (size_t)rr * b.Np);
    double s = fma(v.x, zz.x, fma(v.y, zz.y, fma(v.z, zz.z, v.w * zz.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[rr] = s;
  }
}

__global__ void __launch_bounds__(NB) areduce_kernel(Batch b) {         // alpha[128 j + r] = sum_{m >= j} part[(m, j)][r]
  const int slot = blockIdx.y, j = blockIdx.x;
  if (b.slot_status[slot] != 0) return;
  const double* part = b.solve_part + (size_t)slot * tri_count(b.T) * NB + threadIdx.x;
  double s = 0.0;
  for (int m = j; m < b.T; m++) s += part[(size_t)(tri_count(m) + j) * NB];
  b.alpha[(size_t)slot * b.Np + j * NB + threadIdx.x] = s;
}

void launch_solve(const Batch& b, cudaStream_t s) {
  dim3 tiles(tri_count(b.T), b.nslots), cols(b.T, b.nslots);
  ztile_kernel<<<tiles, 256, 0, s>>>(b);
  zreduce_kernel<<<cols, NB, 0, s>>>(b);
  atile_kernel<<<tiles, 256, 0, s>>>(b);
  areduce_kernel<<<cols, NB, 0, s>>>(b);
  g_launch_count += 4;
}

int configure_chol_kernels() {
  cudaError_t e;
  e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const void* gemms[] = {(const void*)chol_trsm_kernel, (const void*)chol_update_kernel, (const void*)inv_trsm_kernel,
                         (const void*)inv_update_kernel, (const void*)kinv_kernel};
  for (const void* f : gemms) {
    e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace b200gp
