// Compiled kernel programs: the register-resident ("polynomial") form the fast element-wise kernels run.
//
// A candidate kernel is a +/* tree over 1-D leaves (GPy Add / Prod over RBF, RationalQuadratic, StandardPeriodic,
// LinScaleShift, Bias; reference binding src/autoks/backend/kernel.py:17-22).  Expanding the products of sums gives
//     K = sum_g prod_{l in term g} v_l,
// a polynomial in the leaf values v_l in which every leaf is still evaluated ONCE.  Leaves get fixed slots
// 0..L-1 (postfix order), terms are bit masks over the slots.  With L <= 4 or <= 8 the whole state of the
// forward AND reverse sweep (leaf values, adjoints, per-parameter accumulators) lives in registers with static
// indices -- no per-token arrays in local memory, which is what bounded the interpreter of kprog.cuh.
// Programs that do not fit (more than 8 leaves, more than 18 terms, more than 32 tokens) keep the interpreter.
#pragma once
#include "common.cuh"

namespace b200gp {

constexpr int CP_LMAX = 8;       // leaf slots
constexpr int CP_GMAX = 18;      // product terms (the most 8 leaves can expand to: 3*3*2)
constexpr int CP_MAXTOK = 32;    // tokens the device-side expander handles

enum CKind : int { CK_POLY4 = 0, CK_POLY8 = 1, CK_GENERIC = 2, CK_KINDS = 3 };

struct __align__(16) CProg {
  int kind, L, G, single;        // single: every term is one leaf (a pure sum) -> all adjoints are 1
  int4 leaf[CP_LMAX];            // {op, active_dim, theta_offset, n_params}
  unsigned char term[24];        // slot bit mask of each product term (CP_GMAX used)
  int nparam, pad;               // kernel parameters (noise excluded)
};
static_assert(sizeof(CProg) == 176, "CProg layout");

__host__ __device__ inline int leaf_param_count(int op) {
  return (op == OP_RQ || op == OP_PER) ? 3 : (op == OP_CONST ? 1 : 2);      // SE, LIN, LOOKUP: 2
}

}  // namespace b200gp
