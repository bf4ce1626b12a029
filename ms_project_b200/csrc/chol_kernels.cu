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
#include "gemm_body.cuh"
#include "tile_gemm64.cuh"
#include "tile_potrf_la.cuh"
#include "tile_gemm64_tma.cuh"
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

// The same step on the look-ahead factorisation of [A | I] (tile_potrf_la.cuh): L, D = L^-1 and V_kk = D^T go from the
// registers straight to global memory (a warp covers 8 rows x 32 bytes of L and of V, 4 rows x 64 bytes of D).
template <bool SPLIT>
__global__ void __launch_bounds__(TL_THREADS, 1) potrf_diag_la_kernel(Batch b, int k) {
  const int slot = b.slot0 + blockIdx.x;
  if (b.slot_status[slot] != 0) return;
  __shared__ TileLaSmem S;
  int ty, tx;
  tile_la_coords(ty, tx);
  double* A = b.M1 + slot * b.mstride + (size_t)(k * NB) * b.Np + k * NB;
  double a[TL_NE], w[TL_NE];
#pragma unroll
  for (int p = 0; p < TL_R; p++)
#pragma unroll
    for (int q = 0; q <= p; q++) a[tl_idx(p, q)] = A[(size_t)(ty + 16 * p) * b.Np + tx + 16 * q];
  tile_potrf_inverse_la<SPLIT>(a, w, S);
  if (S.fail != 0) {   // uniform
    if (threadIdx.x == 0) b.slot_status[slot] = k * NB + S.fail;
    return;
  }
  if (threadIdx.x < 32) {      // log-determinant partial, reduced in a fixed order
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < NB / 32; r++) s += log(S.piv[threadIdx.x + 32 * r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) b.logdet_part[slot * b.T + k] = s;
  }
  double* Dk = b.Dinv + ((size_t)slot * b.T + k) * (NB * NB);
  double* V = b.M2 ? b.M2 + slot * b.mstride + (size_t)(k * NB) * b.Np + k * NB : nullptr;
#pragma unroll
  for (int p = 0; p < TL_R; p++)
#pragma unroll
    for (int q = 0; q < TL_R; q++) {
      {   // L: element (ty + 16 p, tx + 16 q); the strict upper part of the tile is zeroed
        const int i = ty + 16 * p, c = tx + 16 * q;
        A[(size_t)i * b.Np + c] = (q <= p && c <= i) ? a[tl_idx(q <= p ? p : 0, q <= p ? q : 0)] : 0.0;
      }
      {   // D: element (tx + 16 p, ty + 16 q), exact zeros above the diagonal; V_kk = D^T
        const int i = tx + 16 * p, c = ty + 16 * q;
        const double v = (q <= p) ? w[tl_idx(q <= p ? p : 0, q <= p ? q : 0)] : 0.0;
        Dk[i * NB + c] = v;
        if (V) V[(size_t)c * b.Np + i] = v;
      }
    }
}

void launch_potrf_diag(const Batch& b, int k, cudaStream_t s) {
  if (b.potrf_la == 2) potrf_diag_la_kernel<true><<<b.nslots, TL_THREADS, 0, s>>>(b, k);
  else if (b.potrf_la) potrf_diag_la_kernel<false><<<b.nslots, TL_THREADS, 0, s>>>(b, k);
  else potrf_diag_kernel<<<b.nslots, TB_THREADS, PD_SMEM_BYTES, s>>>(b, k);
  g_launch_count++;
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

// ------------------------------------------------------------------------------------------------
// K^-1 = V V^T on 64 x 64 CTA tiles.
//
// tools/gemm_probe ("kinv pattern") compares CTA tile shapes on exactly this phase's work: tile lengths from K = 64 to N in
// one launch, upper-triangular first K block, only the lower triangle wanted.  Skipping the structural zeros per WARP inside
// 128 x 64 tiles (kinv_kernel) saves 16 % of the flops but costs as much pipe efficiency, because the skipping warps wait
// at the chunk barrier for the working ones: 29.9 TF/s in the probe, the same as not skipping at all (29.6).  Skipping
// per CTA at 64 x 64 granularity -- blocks above the diagonal are not launched, block row I starts at k = 64 I -- removes
// half of that waste with no predicate in the inner loop: 32.3 TF/s in the probe with 3 CTAs (4 warps each) per SM.
// Every output element still sums its products in ascending k, so the values are bit-identical to kinv_kernel's.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K64_THREADS, 3) kinv64_kernel(Batch b) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int I, J;
  tri_decode(blockIdx.x, I, J);                            // block row ascending = K length descending
  const size_t ld = b.Np;
  const double* V = b.M2 + slot * b.mstride;
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, nullptr, ld, 0.0);
  tile64_pipeline(acc, V + (size_t)(I * K64_T) * ld + I * K64_T, V + (size_t)(J * K64_T) * ld + I * K64_T, ld,
                  (b.Np - I * K64_T) / TG_KC, reinterpret_cast<double*>(smem_raw));
  tile64_store(acc, tile64_cw(b.M1 + slot * b.mstride + (size_t)(I * K64_T) * ld + J * K64_T, ld), ld, 1.0);
}

// Cholesky trailing / in-panel update on 64 x 64 blocks: A_ij -= sum_{k in [k0, k0+kw)} L_ik L_jk^T.  blockIdx.x = 4 x (tile
// index of chol_update_kernel) + quadrant; the upper-right quadrant of a diagonal tile is never read and does no work.
__global__ void __launch_bounds__(K64_THREADS, 3) chol_update64_kernel(Batch b, UpdRange upd) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  const int bx = blockIdx.x >> 2, vi = (blockIdx.x >> 1) & 1, vj = blockIdx.x & 1;
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
  if (i == j && vj > vi) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ld = b.Np;
  double* M1 = b.M1 + slot * b.mstride;
  const int ri = i * NB + vi * K64_T, rj = j * NB + vj * K64_T;
  double* Cw = tile64_cw(M1 + (size_t)ri * ld + rj, ld);
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, Cw, ld, -1.0);                          // accumulators start at -C: the C read overlaps the pipeline fill
  tile64_pipeline(acc, M1 + (size_t)ri * ld + upd.k0 * NB, M1 + (size_t)rj * ld + upd.k0 * NB, ld, upd.kw * (NB / TG_KC),
                  reinterpret_cast<double*>(smem_raw));
  tile64_store(acc, Cw, ld, -1.0);
}

// Right-looking L^-T update on 64 x 64 blocks: S_jm (+)= sum_{i in [max(j,k0), k0+kw)} V_ji L_mi^T.  When the K range starts at
// V_jj (upper triangular) the lower 64 rows of the tile skip its first 64 columns.
__global__ void __launch_bounds__(K64_THREADS, 3) inv_update64_kernel(Batch b, UpdRange upd) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  const int bx = blockIdx.x >> 2, vi = (blockIdx.x >> 1) & 1, vj = blockIdx.x & 1;
  const int J = upd.k0 + upd.kw;
  const int m = upd.j0 + bx / J, j = bx % J;
  const int kb = j > upd.k0 ? j : upd.k0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ld = b.Np;
  double* M1 = b.M1 + slot * b.mstride;
  double* M2 = b.M2 + slot * b.mstride;
  const int rj = j * NB + vi * K64_T, rm = m * NB + vj * K64_T;
  const int kskip = (kb == j && vi == 1) ? K64_T : 0;      // structural zeros of V_jj below its diagonal
  double* Cw = tile64_cw(M2 + (size_t)rj * ld + rm, ld);
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, Cw, ld, (j >= upd.k0) ? 0.0 : 1.0);
  tile64_pipeline(acc, M2 + (size_t)rj * ld + kb * NB + kskip, M1 + (size_t)rm * ld + kb * NB + kskip, ld,
                  ((J - kb) * NB - kskip) / TG_KC, reinterpret_cast<double*>(smem_raw));
  tile64_store(acc, Cw, ld, 1.0);
}

// In-place triangular solves on 64 x 64 blocks: X D^T with D the 128 x 128 inverse of a diagonal block of L (lower
// triangular, so output columns [64 vj, 64 vj + 64) need k < 64 (vj + 1)).  A CTA owns 64 rows of one tile: the right block
// first (it reads all 128 columns and writes columns 64..127), then the left one.
//   VT = false: L_ik = A_ik D_k^T          (Cholesky, tile rows i > step of M1)
//   VT = true:  V_ji = -S_ji D_i^T         (L^-T, tile rows j < step of M2)
template <bool VT>
__global__ void __launch_bounds__(K64_THREADS, 3) trsm64_kernel(Batch b, int step) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int t = blockIdx.x >> 1, vi = blockIdx.x & 1;
  const int row_tile = VT ? t : step + 1 + t;
  const size_t ld = b.Np;
  double* X = (VT ? b.M2 : b.M1) + slot * b.mstride + (size_t)(row_tile * NB + vi * K64_T) * ld + step * NB;
  const double* D = b.Dinv + ((size_t)slot * b.T + step) * (NB * NB);
  double acc[TG_MT][TG_NT][2];
#pragma unroll
  for (int vj = 1; vj >= 0; vj--) {
    tile64_init(acc, nullptr, ld, 0.0);
    tile64_pipeline(acc, X, D + (size_t)(vj * K64_T) * NB, ld, (vj + 1) * (K64_T / TG_KC), smem, NB);
    tile64_store(acc, tile64_cw(X + vj * K64_T, ld), ld, VT ? -1.0 : 1.0);
    __syncthreads();      // every warp has left the stages (and stored its part) before the left block refills them
  }
}

// ------------------------------------------------------------------------------------------------
// The same three tile GEMMs with TMA-fed operands (tile_gemm64_tma.cuh).  The tensor maps describe M1 / M2 as one
// (slots x Np) x Np matrix each, so a slot is a row offset.
// ------------------------------------------------------------------------------------------------
struct TmaMaps { CUtensorMap m1, m2; };
size_t tma_maps_bytes() { return sizeof(TmaMaps); }

template <int SW, int ST, int MINB>
__global__ void __launch_bounds__(K64_THREADS, MINB) kinv64_tma_kernel(Batch b, const __grid_constant__ CUtensorMap mapV) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int I, J;
  tri_decode(blockIdx.x, I, J);
  const size_t ld = b.Np;
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, nullptr, ld, 0.0);
  tile64_pipeline_tma<SW, ST>(acc, &mapV, I * K64_T, slot * b.Np + I * K64_T, &mapV, I * K64_T, slot * b.Np + J * K64_T,
                          (b.Np - I * K64_T) / TG_KC, smem_raw);
  tile64_store(acc, tile64_cw(b.M1 + slot * b.mstride + (size_t)(I * K64_T) * ld + J * K64_T, ld), ld, 1.0);
}

template <int SW, int ST, int MINB>
__global__ void __launch_bounds__(K64_THREADS, MINB) chol_update64_tma_kernel(Batch b, UpdRange upd,
                                                                           const __grid_constant__ CUtensorMap mapL) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  const int bx = blockIdx.x >> 2, vi = (blockIdx.x >> 1) & 1, vj = blockIdx.x & 1;
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
  if (i == j && vj > vi) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ld = b.Np;
  double* M1 = b.M1 + slot * b.mstride;
  const int ri = i * NB + vi * K64_T, rj = j * NB + vj * K64_T;
  double* Cw = tile64_cw(M1 + (size_t)ri * ld + rj, ld);
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, Cw, ld, -1.0);
  tile64_pipeline_tma<SW, ST>(acc, &mapL, upd.k0 * NB, slot * b.Np + ri, &mapL, upd.k0 * NB, slot * b.Np + rj,
                          upd.kw * (NB / TG_KC), smem_raw);
  tile64_store(acc, Cw, ld, -1.0);
}

template <int SW, int ST, int MINB>
__global__ void __launch_bounds__(K64_THREADS, MINB) inv_update64_tma_kernel(Batch b, UpdRange upd,
                                                                          const __grid_constant__ CUtensorMap mapL,
                                                                          const __grid_constant__ CUtensorMap mapV) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  const int bx = blockIdx.x >> 2, vi = (blockIdx.x >> 1) & 1, vj = blockIdx.x & 1;
  const int J = upd.k0 + upd.kw;
  const int m = upd.j0 + bx / J, j = bx % J;
  const int kb = j > upd.k0 ? j : upd.k0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ld = b.Np;
  double* M2 = b.M2 + slot * b.mstride;
  const int rj = j * NB + vi * K64_T, rm = m * NB + vj * K64_T;
  const int kskip = (kb == j && vi == 1) ? K64_T : 0;
  double* Cw = tile64_cw(M2 + (size_t)rj * ld + rm, ld);
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, Cw, ld, (j >= upd.k0) ? 0.0 : 1.0);
  tile64_pipeline_tma<SW, ST>(acc, &mapV, kb * NB + kskip, slot * b.Np + rj, &mapL, kb * NB + kskip, slot * b.Np + rm,
                          ((J - kb) * NB - kskip) / TG_KC, smem_raw);
  tile64_store(acc, Cw, ld, 1.0);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tma_encode_maps(void* out, double* M1, double* M2, int Np, size_t rows, int sw) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess)
      return -1;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  TmaMaps* maps = static_cast<TmaMaps*>(out);
  const CUtensorMapSwizzle mode = sw == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : sw == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                                                          : CU_TENSOR_MAP_SWIZZLE_NONE;
  const cuuint64_t dims[2] = {(cuuint64_t)Np, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)Np * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)TG_KC, (cuuint32_t)K64_T};
  const cuuint32_t estr[2] = {1, 1};
  double* ptrs[2] = {M1, M2};
  CUtensorMap* dst[2] = {&maps->m1, &maps->m2};
  for (int q = 0; q < 2; q++) {
    if (!ptrs[q]) continue;
    const CUresult r = encode(dst[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ptrs[q], dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, mode, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)r;
  }
  return 0;
}

// b.tma: 1 / 2 / 3 = swizzle none / 128B / 128B_ATOM_32B with a 4-stage ring and three CTAs per SM; 4 = 128B_ATOM_32B with a
// 3-stage ring and FOUR CTAs per SM (the TMA kernels need <= 128 registers)
#define TMA_DISPATCH(b, CALL)                          \
  switch ((b).tma) {                                   \
    case 1: CALL(0, 4, 3); break;                      \
    case 2: CALL(1, 4, 3); break;                      \
    case 4: CALL(2, 3, 4); break;                      \
    default: CALL(2, 4, 3); break;                     \
  }

void launch_kinv64(const Batch& b, cudaStream_t s) {
  dim3 grid(tri_count(2 * b.T), b.nslots);
  if (b.tma && b.tmaps) {
    const TmaMaps& tm = *static_cast<const TmaMaps*>(b.tmaps);
#define CALL(SW, ST, MB) kinv64_tma_kernel<SW, ST, MB><<<grid, K64_THREADS, k64t_smem_bytes(ST), s>>>(b, tm.m2)
    TMA_DISPATCH(b, CALL)
#undef CALL
  } else {
    kinv64_kernel<<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b);
  }
  g_launch_count++;
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
    // the 64 x 64 form doubles the CTAs of these small launches: one N = 2048 evaluation 1.8 -> 1.6 ms; on large batches the
    // two forms are equal (25.0 vs 24.9 ms per 256 items), so it is used below 65 slots (and always with gemm64 = 2)
    case GM_TRSM:
      if (b.gemm64 >= 2 || (b.gemm64 == 1 && b.nslots < 65)) { grid.x *= 2; trsm64_kernel<false><<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b, step); }
      else chol_trsm_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, step);
      break;
    case GM_VTRSM:
      if (b.gemm64 >= 2 || (b.gemm64 == 1 && b.nslots < 65)) { grid.x *= 2; trsm64_kernel<true><<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b, step); }
      else inv_trsm_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, step);
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
  if (b.gemm64) {
    dim3 grid(4 * ntiles, b.nslots);
    if (b.tma && b.tmaps) {
      const TmaMaps& tm = *static_cast<const TmaMaps*>(b.tmaps);
#define CALL(SW, ST, MB) inv_update64_tma_kernel<SW, ST, MB><<<grid, K64_THREADS, k64t_smem_bytes(ST), s>>>(b, r, tm.m1, tm.m2)
      TMA_DISPATCH(b, CALL)
#undef CALL
    } else {
      inv_update64_kernel<<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b, r);
    }
  } else {
    dim3 grid(2 * ntiles, b.nslots);
    inv_update_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, r);
  }
  g_launch_count++;
}

void launch_update(const Batch& b, UpdRange r, cudaStream_t s) {
  int ntiles = 0;
  for (int j = r.j0; j < r.j1 && j < b.T; j++) ntiles += b.T - j;
  if (ntiles <= 0 || r.kw <= 0) return;
  if (b.gemm64) {
    dim3 grid(4 * ntiles, b.nslots);
    if (b.tma && b.tmaps) {
      const TmaMaps& tm = *static_cast<const TmaMaps*>(b.tmaps);
#define CALL(SW, ST, MB) chol_update64_tma_kernel<SW, ST, MB><<<grid, K64_THREADS, k64t_smem_bytes(ST), s>>>(b, r, tm.m1)
      TMA_DISPATCH(b, CALL)
#undef CALL
    } else {
      chol_update64_kernel<<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b, r);
    }
  } else {
    dim3 grid(2 * ntiles, b.nslots);
    chol_update_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b, r);
  }
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
  const void* g64[] = {(const void*)kinv64_kernel, (const void*)chol_update64_kernel, (const void*)inv_update64_kernel,
                       (const void*)trsm64_kernel<false>, (const void*)trsm64_kernel<true>};
  for (const void* f : g64) {
    e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, K64_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
  }
#define SETATTR(K, SW, ST, MB)                                                                                              \
  if ((e = cudaFuncSetAttribute((const void*)K<SW, ST, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, k64t_smem_bytes(ST))) != \
      cudaSuccess)                                                                                                              \
    return (int)e;
#define SETALL(K) SETATTR(K, 0, 4, 3) SETATTR(K, 1, 4, 3) SETATTR(K, 2, 4, 3) SETATTR(K, 2, 3, 4)
  SETALL(kinv64_tma_kernel)
  SETALL(chol_update64_tma_kernel)
  SETALL(inv_update64_tma_kernel)
#undef SETALL
#undef SETATTR
  return 0;
}

}  // namespace b200gp
