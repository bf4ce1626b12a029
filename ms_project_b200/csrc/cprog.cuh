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
  int nparam, shape;             // kernel parameters (noise excluded); canonical polynomial shape (<= 4 leaves, see below)
};
static_assert(sizeof(CProg) == 176, "CProg layout");

// Canonical shapes of the polynomials over <= 4 distinct leaves.  Every +/* tree with at most four leaves expands, up to a
// relabelling of its leaves, to one of these 17 term sets; the compile kernel finds the relabelling (it is free: leaf slots
// are arbitrary) and the <= 4-leaf element-wise kernels then evaluate K and all adjoints dK/dv_l by ONE switch over the
// shape per row step, in factored form -- instead of a loop over terms with a 16-way switch over the term's mask per
// leaf (ncu: 28 % of the gradient kernel's instructions, 25 % of its stall samples).  Leaves a, b, c, d = slots 0..3;
// key = the sorted term masks, four bits each.
//    0  a            1  a + b         2  ab             3  a + b + c      4  a + bc         5  a (b + c)
//    6  abc          7  a + b + c + d                   8  a + b + cd     9  a + b (c + d) 10  a + bcd
//   11  a (b + c + d)                12  (a + d)(b + c) 13  ab + cd       14  a (b + cd)    15  ab (c + d)   16  abcd
constexpr int CP_NSHAPES = 17;
__host__ __device__ inline unsigned cp_shape_key(int s) {
  const unsigned short keys[CP_NSHAPES] = {0x0001, 0x0021, 0x0003, 0x0421, 0x0061, 0x0053, 0x0007, 0x8421, 0x0C21,
                                           0x0A61, 0x00E1, 0x0953, 0xCA53, 0x00C3, 0x00D3, 0x00B7, 0x000F};
  return keys[s];
}
__host__ __device__ inline int cp_shape_leaves(int s) { return s == 0 ? 1 : (s <= 2 ? 2 : (s <= 6 ? 3 : 4)); }

__host__ __device__ inline int leaf_param_count(int op) {
  return (op == OP_RQ || op == OP_PER) ? 3 : (op == OP_CONST ? 1 : 2);      // SE, LIN, LOOKUP: 2
}

}  // namespace b200gp
