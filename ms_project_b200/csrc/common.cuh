// Shared definitions for libb200gp (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace b200gp {

// ---- tiling -----------------------------------------------------------------------------------
constexpr int NB = 128;            // tile edge of every blocked phase (Cholesky / inverse / grad)
constexpr int MAX_DIMS = 16;       // input dimensions staged per tile
constexpr int MAX_TOKENS = 64;     // postfix tokens per kernel program
constexpr int MAX_PARAMS = 128;    // kernel hyperparameters per program (noise excluded)
constexpr int MAX_STACK = 8;       // evaluation-stack depth

// ---- kernel-program opcodes (host compiles Covariance.postfix_tokens into these) ---------------
// token = int4 {op, a, b, c}: leaf  -> {op, active_dim, theta_offset, 0}
//                             ADD/MUL -> {op, lhs_token, rhs_token, 0}
//                             LOOKUP -> {op, index_dim, theta_offset, 0}: RBF over a looked-up distance matrix
//                                       (RBFDistanceBuilderKernelKernel of the BOMS surrogate; the inputs are model indices)
enum Op : int { OP_SE = 0, OP_RQ = 1, OP_PER = 2, OP_LIN = 3, OP_CONST = 4, OP_ADD = 5, OP_MUL = 6, OP_LOOKUP = 7 };
__host__ __device__ inline bool is_leaf(int op) { return op <= OP_CONST || op == OP_LOOKUP; }

// gradient quirk bits (SURVEY.md 8c hazard 3): reproduce the notebook formulas of the fork
constexpr int QUIRK_PER_LENGTHSCALE = 1;   // dK/dl of PER 4x too small (periodic.ipynb#2)
constexpr int QUIRK_LIN_SHIFT = 2;         // dK/dshift of LIN without the variance factor (linear.ipynb#2)

// per-item status codes returned to the host (never an exception / non-zero return)
enum Status : int { ST_OK = 0, ST_JITTERED = 1, ST_NOT_PD = 2, ST_NONFINITE = 3 };

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int tri_count(int t) { return t * (t + 1) / 2; }

// lower-triangular tile index -> (row, col), row >= col
__device__ inline void tri_decode(int idx, int& r, int& c) {
  int rr = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
  while (tri_count(rr + 1) <= idx) rr++;
  while (tri_count(rr) > idx) rr--;
  r = rr;
  c = idx - tri_count(rr);
}

}  // namespace b200gp
