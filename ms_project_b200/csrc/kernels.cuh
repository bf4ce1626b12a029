// Internal launch interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <atomic>
#include "common.cuh"
#include "cprog.cuh"

namespace b200gp {

// Device-side view of one batch of `nslots` matrices being scored together.
struct Batch {
  // geometry
  int N, Np, T, D;
  int nslots;
  int slot0;                 // first slot of this launch group (Cholesky / inverse phases split a chunk over two streams)
  size_t mstride;            // Np * Np
  // per-slot scratch (slot-major)
  double* M1;                // K -> L (lower) -> K^-1 (lower)
  double* M2;                // V = L^-T (upper)
  double* Dinv;              // [slot][T][NB*NB] inverses of the diagonal blocks of L
  double* z;                 // [slot][Np]  L^-1 y
  double* alpha;             // [slot][Np]  K^-1 y
  double* logdet_part;       // [slot][T]
  double* diag_part;         // [slot][T]   partial sums of diag(K_y) (jitter ladder)
  double* grad_part;         // [slot][ntiles][4][pstride]: gradient partials per 64 x 64 quadrant of a tile (elem_kernels.cu GP_PER_TILE)
  double* solve_part;        // [slot][ntiles][NB] tile partials of z = V^T y, then of alpha = V z
  int pstride;
  int* slot_status;          // [slot] 0 ok, >0 = 1-based index of first bad pivot
  // shared data
  const double* Xt;          // [D][Np] dimension-major, zero padded
  const double* y;           // [Np] zero padded
  // per-item inputs / outputs (indexed through ids[slot])
  const int* ids;
  const int4* prog;
  const int* prog_off;
  const double* theta;
  const int* theta_off;
  const double* jitter;      // [slot] extra diagonal jitter (jitchol ladder), may be null
  double* lml;               // [item]
  double* dlml;              // [theta layout]
  int* status;               // [item]
  int* tries;                // [item] jitter-ladder tries used (may be null)
  int try_no;                // ladder step of this pipeline run (0 = plain)
  int quirks;
  int add_diag;              // 1: add noise + 1e-8 + jitter to the diagonal (K_y); 2: add the noise only (Hellinger Gram
                             // matrices K + lambda I, distance.py:186-190); 0: bare kern.K
  const double* lut;         // OP_LOOKUP distance matrix (lut_n x lut_n, b200gp_set_lookup) or null
  int lut_n;
  int gemm64;                // 1: Cholesky / L^-T updates on 64 x 64 CTA tiles (chol_update64 / inv_update64), 0: 128 x 64 tiles
  int potrf_la;              // 1: diagonal tiles by the look-ahead register factorisation (tile_potrf_la.cuh), 0: the blocked one
  int tma;                   // 0: the 64 x 64 tile GEMMs stage their operands with cp.async; 1 + SW: with the TMA (tile_gemm64_tma.cuh)
                             // and swizzle mode SW (0 none, 1 128B, 2 128B_ATOM_32B) of the tensor maps below
  const void* tmaps;         // host pointer: TmaMaps of this batch's M1 / M2 (launch-time only; kernels get the maps by value)
  // compiled programs of this chunk (cprog.cuh), written by launch_compile
  CProg* cprog;              // [slot]
  int* klist;                // [CK_KINDS][list_stride] slots of each kind, in slot order
  int* kcount;               // [CK_KINDS]
  int list_stride;
};

struct TmaMaps;             // chol_kernels.cu: CUtensorMap of M1 and of M2, rows = slots x Np, columns = Np
// Encodes the tensor maps of a batch's M1 / M2 (M2 may be null) for swizzle mode `sw`; returns 0 or a CUresult / -1 when the
// driver entry point is unavailable.  `out` must hold tma_maps_bytes() bytes, 64-byte aligned.
int tma_encode_maps(void* out, double* M1, double* M2, int Np, size_t rows, int sw);
size_t tma_maps_bytes();

enum GemmMode : int { GM_TRSM = 0, GM_VTRSM = 3, GM_KINV = 4, GM_UPD = 5, GM_VUPD = 6 };

// GM_UPD: A_ij -= sum_{k in [k0, k0+kw)} L_ik L_jk^T for the tiles with column j in [j0, j1), i >= j.
struct UpdRange { int k0, kw, j0, j1; };
// GM_VUPD (right-looking L^-T): S_jm (+)= sum_{i in [max(j,k0), k0+kw)} V_ji L_mi^T for columns m in [j0, j1), rows
// j < k0+kw; a tile is overwritten when its first contribution arrives (j >= k0), accumulated afterwards.

void launch_compile(const Batch& b, cudaStream_t s);                        // programs -> CProg + per-kind slot lists
void launch_build_k(const Batch& b, cudaStream_t s);
void launch_export_sym(const Batch& b, double* out, cudaStream_t s);       // out[slot][N][N] symmetric copy of M1
void launch_potrf_diag(const Batch& b, int k, cudaStream_t s);
void launch_gemm(const Batch& b, int mode, int step, cudaStream_t s);
void launch_kinv64(const Batch& b, cudaStream_t s);                        // K^-1 = V V^T on 64 x 64 CTA tiles (same values as GM_KINV)
void launch_update(const Batch& b, UpdRange r, cudaStream_t s);
void launch_vupdate(const Batch& b, UpdRange r, cudaStream_t s);
void launch_solve(const Batch& b, cudaStream_t s);                          // z = L^-1 y, alpha = L^-T z
void launch_kinv_grad(const Batch& b, cudaStream_t s);                     // K^-1 fused with the gradient of the <= 4-leaf programs
void launch_grad(const Batch& b, cudaStream_t s, bool poly4_done);          // poly4_done: launch_kinv_grad covered the <= 4-leaf programs
void launch_finalize(const Batch& b, int with_grad, cudaStream_t s);
int configure_kernels();
extern std::atomic<long long> g_launch_count;                                            // kernels launched (host-side counter)                                                    // opt-in shared memory sizes

}  // namespace b200gp
