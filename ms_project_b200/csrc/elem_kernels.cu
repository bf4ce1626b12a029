// Element-wise phases of the scoring path: fused kernel-matrix build, fused gradient pass, finalize.
//
// K-build replaces `kern.K(X)` + `diag.add(Ky, variance + 1e-8)` (GPy ExactGaussianInference; reference
// call sites src/autoks/backend/kernel.py:317-322, src/autoks/distance/distance.py:186-187).
// The gradient pass replaces `kern.update_gradients_full(dL_dK, X)` and the Gaussian likelihood's
// `exact_inference_gradients`: dL/dtheta_p = sum_ij 1/2 (alpha_i alpha_j - K^-1_ij) dK_ij/dtheta_p with
// dK regenerated from X in registers, so neither dL/dK nor dK ever exists in HBM.
//
// Two implementations share the work of one chunk (the compile kernel sorts the slots into lists):
//   * poly kernels  -- programs with <= 8 leaves run in the register-resident polynomial form of cprog.cuh;
//   * interp kernels -- everything else runs the generic postfix interpreter of kprog.cuh.
// All of them are persistent (grid = resident CTAs, grid-stride over (slot, tile) work items), so an empty list
// costs one wave of CTAs that exit at once.
#include "kernels.cuh"
#include "gemm_body.cuh"
#include "tile_gemm64.cuh"
#include "kprog.cuh"
#include "cprog.cuh"
#include "fastmath.cuh"
#include <cstddef>

// rows a thread of the polynomial-form kernels interleaves (independent FP64 chains)
#ifndef POLY_E4
#define POLY_E4 4              // <= 4-leaf build kernel: 4 rows per step halve the per-step op dispatch (4.7 -> 4.3 ms / 256 items)
#endif
#ifndef POLY_BUILD_CTAS
#define POLY_BUILD_CTAS 3      // resident CTAs per SM the <= 4-leaf kernels are compiled for (register cap 85: 78 used, no spills; 5.7 -> 5.1 ms per 256 items)
#endif
#ifndef POLY_GRAD_CTAS
#define POLY_GRAD_CTAS 2
#endif
#ifndef POLY_GRAD8_CTAS
#define POLY_GRAD8_CTAS 2      // 5-8 leaf gradient kernel: 128 registers with ~300 B of spills beats 212 registers at one CTA per SM (C4: 56 -> 38 ms)
#endif
#ifndef POLY_E8
#define POLY_E8 1
#endif
#ifndef POLY_GE4
#define POLY_GE4 2             // rows interleaved by the <= 4-leaf gradient kernel (4 would spill at 128 registers; 4 rows at one
                               // CTA per SM, 234 registers: 23.3 instead of 20.1 ms per 768 items)
#endif

namespace b200gp {

constexpr double EXACT_INFERENCE_JITTER = 1e-8;   // GPy exact_gaussian_inference.py: diag.add(Ky, variance + 1e-8)


static int g_sm_count = 148;
static int g_grid_poly_build[2] = {296, 148}, g_grid_poly_grad[2] = {296, 148}, g_grid_interp_build = 592, g_grid_interp_grad = 592;

// ------------------------------------------------------------------------------------------------
// Compile: postfix program -> CProg per slot, then the per-kind slot lists (slot order preserved).
// ------------------------------------------------------------------------------------------------
__device__ void compile_one(const Batch& b, int slot) {
  const int item = b.ids[slot];
  const int t0 = b.prog_off[item], nt = b.prog_off[item + 1] - t0;
  CProg& cp = b.cprog[slot];
  cp.nparam = b.theta_off[item + 1] - b.theta_off[item] - 1;
  cp.shape = -1;
  int L = 0, G = 0, single = 0, kind = CK_GENERIC;
  if (nt <= CP_MAXTOK) {
    unsigned char masks[CP_MAXTOK][CP_GMAX];
    unsigned char cnt[CP_MAXTOK];
    bool ok = true;
    for (int t = 0; t < nt && ok; t++) {
      const int4 k = b.prog[t0 + t];
      if (k.x == OP_LOOKUP) { ok = false; break; }      // table look-ups stay with the interpreter
      if (k.x <= OP_CONST) {
        if (L >= CP_LMAX) { ok = false; break; }
        cp.leaf[L] = make_int4(k.x, k.y, k.z, leaf_param_count(k.x));
        masks[t][0] = (unsigned char)(1u << L);
        cnt[t] = 1;
        L++;
      } else if (k.x == OP_ADD) {
        const int na = cnt[k.y], nb = cnt[k.z];
        if (na + nb > CP_GMAX) { ok = false; break; }
        for (int i = 0; i < na; i++) masks[t][i] = masks[k.y][i];
        for (int i = 0; i < nb; i++) masks[t][na + i] = masks[k.z][i];
        cnt[t] = (unsigned char)(na + nb);
      } else {
        const int na = cnt[k.y], nb = cnt[k.z];
        if (na * nb > CP_GMAX) { ok = false; break; }
        for (int i = 0; i < na; i++)
          for (int j = 0; j < nb; j++) masks[t][i * nb + j] = masks[k.y][i] | masks[k.z][j];
        cnt[t] = (unsigned char)(na * nb);
      }
    }
    if (ok && nt > 0) {
      G = cnt[nt - 1];
      single = 1;
      for (int i = 0; i < G; i++) {
        cp.term[i] = masks[nt - 1][i];
        if (__popc((unsigned)masks[nt - 1][i]) != 1) single = 0;
      }
      kind = (L <= 4) ? CK_POLY4 : CK_POLY8;
    }
  }
  for (int i = L; i < CP_LMAX; i++) cp.leaf[i] = make_int4(OP_CONST, 0, 0, 0);
  for (int i = G; i < 24; i++) cp.term[i] = 0;
  cp.shape = -1;
  if (kind == CK_POLY4) {       // relabel the leaves so that the term set is one of the canonical shapes (cprog.cuh)
    int found = -1, fperm = 0;
    for (int pi = 0; pi < 24 && found < 0; pi++) {
      // pi-th permutation of {0, 1, 2, 3} (factorial number system): perm[i] = new slot of leaf i
      int perm[4], pool[4] = {0, 1, 2, 3}, code = pi;
      for (int i = 0, f = 6; i < 4; i++) {
        const int d = code / f;
        code %= f;
        perm[i] = pool[d];
        for (int t = d; t < 3 - i; t++) pool[t] = pool[t + 1];
        if (i < 2) f /= (3 - i);
      }
      unsigned m[4] = {0, 0, 0, 0};
      bool inside = true;
      for (int g = 0; g < 4; g++)
        if (g < G) {
          for (int i = 0; i < 4; i++)
            if (cp.term[g] & (1u << i)) { m[g] |= 1u << perm[i]; if (perm[i] >= L) inside = false; }
        }
      if (!inside || G > 4) continue;
      for (int x = 0; x < G; x++)            // sort ascending (G <= 4)
        for (int y = x + 1; y < G; y++)
          if (m[y] < m[x]) { const unsigned t = m[x]; m[x] = m[y]; m[y] = t; }
      const unsigned key = m[0] | (m[1] << 4) | (m[2] << 8) | (m[3] << 12);
      for (int sidx = 0; sidx < CP_NSHAPES; sidx++)
        if (cp_shape_leaves(sidx) == L && cp_shape_key(sidx) == key) { found = sidx; fperm = pi; }
    }
    if (found >= 0) {
      int perm[4], pool[4] = {0, 1, 2, 3}, code = fperm;
      for (int i = 0, f = 6; i < 4; i++) {
        const int d = code / f;
        code %= f;
        perm[i] = pool[d];
        for (int t = d; t < 3 - i; t++) pool[t] = pool[t + 1];
        if (i < 2) f /= (3 - i);
      }
      int4 old[4];
      for (int i = 0; i < 4; i++) old[i] = cp.leaf[i];
      for (int i = 0; i < L; i++) cp.leaf[perm[i]] = old[i];
      const unsigned key = cp_shape_key(found);
      for (int g = 0; g < 4; g++) cp.term[g] = (unsigned char)((key >> (4 * g)) & 15u);
      cp.shape = found;
    } else {
      kind = CK_POLY8;          // not reachable for +/* trees over distinct leaves; the generic term loop handles it anyway
    }
  }
  cp.kind = kind; cp.L = L; cp.G = G; cp.single = single;
}

__global__ void __launch_bounds__(256) compile_programs_kernel(Batch b) {
  for (int slot = threadIdx.x; slot < b.nslots; slot += blockDim.x) compile_one(b, slot);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int cnt[CK_KINDS] = {0, 0, 0};
    for (int base = 0; base < b.nslots; base += 32) {
      const int slot = base + lane;
      const int kind = slot < b.nslots ? b.cprog[slot].kind : -1;
#pragma unroll
      for (int k = 0; k < CK_KINDS; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == k);
        if (kind == k) b.klist[k * b.list_stride + cnt[k] + __popc(m & ((1u << lane) - 1u))] = slot;
        cnt[k] += __popc(m);
      }
    }
    if (lane < CK_KINDS) b.kcount[lane] = cnt[lane];
  }
}

void launch_compile(const Batch& b, cudaStream_t s) {
  compile_programs_kernel<<<1, 256, 0, s>>>(b);
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// Polynomial-form kernels.  Thread (c, h) of a 256-thread CTA owns column c = tid & 127 of the 128 x 128 tile and
// the rows [64 h, 64 h + 64): its column-side point values stay in registers for the whole tile, the row-side
// values are broadcast reads of shared memory, global accesses are 256-byte row segments per warp.
// ------------------------------------------------------------------------------------------------
template <int LMAX>
struct PolySmem {
  CProg cp;
  double cst[LMAX][4];          // per leaf: SE {v, 1/l}; RQ {v, 1/l, a, 1/2a}; PER {v, pi/p, 1/l, 1/p}; LIN {v, shift}
  double ivar[LMAX];
  double Ai[LMAX][NB];          // row-side point value: x (SE, RQ), pi x / p (PER), x - shift (LIN)
  double Si[LMAX][NB];          // PER: sin / cos of Ai, so sin(a_i - a_j) needs no per-pair sincos
  double Ci[LMAX][NB];
  double ai[NB];                // alpha_i (gradient pass) / diagonal staging (build)
  double red[8][3 * LMAX + 1];
  double noise;
  double etab[64];              // copy of FM_EXPTAB (fastmath.cuh exp_tab)
};

// Explicit shared-window loads: the hot loops address PolySmem by 32-bit offsets from one base register instead of
// generic pointers (which cost a S2UR / ULEA window-base sequence per access on sm_100).
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds_v2(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
template <int E>
__device__ __forceinline__ void lds_rows(unsigned a, double (&v)[E]) {      // E consecutive doubles, 16-byte aligned for E = 2
  if (E == 2) lds_v2(a, v[0], v[E - 1]);
  else {
#pragma unroll
    for (int e = 0; e < E; e++) v[e] = lds_f64(a + 8 * e);
  }
}

__device__ __forceinline__ void leaf_constants(int op, const double* __restrict__ th, double* cst, double& ivar) {
  double c0 = th[0], c1 = 0, c2 = 0, c3 = 0;
  switch (op) {
    case OP_SE:  c1 = 1.0 / th[1]; break;
    case OP_RQ:  c1 = 1.0 / th[1]; c2 = th[2]; c3 = 1.0 / (2.0 * th[2]); break;
    case OP_PER: c1 = 3.14159265358979323846 / th[1]; c2 = 1.0 / th[2]; c3 = 1.0 / th[1]; break;
    case OP_LIN: c1 = th[1]; break;
    default: break;
  }
  cst[0] = c0; cst[1] = c1; cst[2] = c2; cst[3] = c3;
  ivar = 1.0 / c0;
}

__device__ __forceinline__ void point_values(int op, const double* cst, double x, double& A, double& S, double& C) {
  S = 0.0; C = 0.0;
  if (op == OP_PER) { A = x * cst[1]; sincos(A, &S, &C); }
  else if (op == OP_LIN) A = x - cst[1];
  else A = x;
}

// Leaf values for E rows of one column.  `cst`, `srow`, `crow` are shared-window addresses (leaf constants, PER
// sin / cos rows); aux keeps what the reverse sweep cannot cheaply recompute.  Straight-line math (fastmath.cuh) so
// the E chains interleave.
template <int E>
__device__ __forceinline__ void leaf_fwd(int op, unsigned cst, unsigned etab, const double (&Ar)[E], unsigned srow, unsigned crow,
                                         double Aj, double Sj, double Cj, double (&v)[E], double (&aux)[E]) {
  double c0, c1;
  lds_v2(cst, c0, c1);
  switch (op) {
  case OP_SE: {            // rbf.ipynb#1-2: v exp(-r^2 / 2)
#pragma unroll
    for (int e = 0; e < E; e++) { const double r = (Ar[e] - Aj) * c1; const double a = r * r; aux[e] = a; v[e] = c0 * exp_tab(-0.5 * a, etab); }
  } break;
  case OP_RQ: {     // rational-quadratic.ipynb#2: v exp(-a log1p(r^2 / 2a))
    double c2, c3;
    lds_v2(cst + 16, c2, c3);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double r = (Ar[e] - Aj) * c1;
      const double lg = log1p_pos((r * r) * c3);
      aux[e] = lg;
      v[e] = c0 * exp_tab(-c2 * lg, etab);
    }
  } break;
  case OP_PER: {    // periodic.ipynb#2: v exp(-2 sin^2(pi (x - x') / p) / l^2); sin(a_i - a_j) by the angle-difference identity
    const double il = lds_f64(cst + 16);
    double Sr[E], Cr[E];
    lds_rows<E>(srow, Sr);
    lds_rows<E>(crow, Cr);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double s = Sr[e] * Cj - Cr[e] * Sj;
      const double q = s * il;
      aux[e] = s;
      v[e] = c0 * exp_tab(-2.0 * (q * q), etab);
    }
  } break;
  case OP_LIN: {    // linear.ipynb#2: v (x - c)(x' - c)
#pragma unroll
    for (int e = 0; e < E; e++) { aux[e] = Ar[e] * Aj; v[e] = c0 * aux[e]; }
  } break;
  default: {
#pragma unroll
    for (int e = 0; e < E; e++) { aux[e] = 0.0; v[e] = c0; }
  } break;
  }
}

// acc[p] += sum_e g[e] * (the element-dependent factor of dv/dtheta_p); the per-leaf constant factors are applied
// once per tile by leaf_grad_scale (same formulas as kprog.cuh eval_grad, constants hoisted out of the N^2 loop).
template <int E>
__device__ __forceinline__ void leaf_bwd(int op, unsigned cst, const double (&Ar)[E], unsigned srow, unsigned crow,
                                         double Aj, double Sj, double Cj, const double (&g)[E], const double (&v)[E],
                                         const double (&aux)[E], double (&acc)[3]) {
  switch (op) {
  case OP_SE: {            // dK/dv = K / v ; dK/dl = K r^2 / l
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double gv = g[e] * v[e];
      acc[0] += gv;
      acc[1] = fma(gv, aux[e], acc[1]);
    }
  } break;
  case OP_RQ: {     // rational-quadratic.ipynb#2
    double c1, c3;
    c1 = lds_f64(cst + 8);
    c3 = lds_f64(cst + 24);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double r = (Ar[e] - Aj) * c1;
      const double u = (r * r) * c3;
      const double ue = u * rcp_sl(1.0 + u);
      const double gv = g[e] * v[e];
      acc[0] += gv;
      acc[1] = fma(gv, ue, acc[1]);
      acc[2] = fma(gv, ue - aux[e], acc[2]);
    }
  } break;
  case OP_PER: {    // periodic.ipynb#2
    double Sr[E], Cr[E];
    lds_rows<E>(srow, Sr);
    lds_rows<E>(crow, Cr);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const double s = aux[e];
      const double c = fma(Cr[e], Cj, Sr[e] * Sj);
      const double gv = g[e] * v[e];
      acc[0] += gv;
      acc[1] = fma(gv * s, c * (Ar[e] - Aj), acc[1]);
      acc[2] = fma(gv * s, s, acc[2]);
    }
  } break;
  case OP_LIN: {    // linear.ipynb#2
#pragma unroll
    for (int e = 0; e < E; e++) {
      acc[0] = fma(g[e], aux[e], acc[0]);
      acc[1] = fma(g[e], Ar[e] + Aj, acc[1]);
    }
  } break;
  default: {
#pragma unroll
    for (int e = 0; e < E; e++) acc[0] += g[e];
  } break;
  }
}

// constant factor of parameter p of a leaf: dL/dtheta_p = scale * accumulated sum
__device__ __forceinline__ double leaf_grad_scale(int op, const double* cst, double ivar, int p, int quirks) {
  switch (op) {
    case OP_SE:  return p == 0 ? ivar : cst[1];
    case OP_RQ:  return p == 0 ? ivar : (p == 1 ? 2.0 * cst[2] * cst[1] : 1.0);
    case OP_PER: return p == 0 ? ivar
                        : (p == 1 ? 4.0 * cst[2] * cst[2] * cst[3]
                                  : ((quirks & QUIRK_PER_LENGTHSCALE) ? 1.0 : 4.0) * cst[2] * cst[2] * cst[2]);
    case OP_LIN: return p == 0 ? 1.0 : -((quirks & QUIRK_LIN_SHIFT) ? 1.0 : cst[0]);
    default:     return 1.0;
  }
}

// Cooperative prologue of one (slot, tile) work item: program, leaf constants, row-side tables; returns the
// column-side values of this thread's column and the leaf opcodes in registers.
template <int LMAX>
__device__ __forceinline__ void poly_prologue(PolySmem<LMAX>& S, const Batch& b, int slot, int ti, int tj, int colc,
                                              double (&Aj)[LMAX], double (&Sj)[LMAX], double (&Cj)[LMAX], int (&ops)[LMAX]) {
  __syncthreads();                                           // the previous work item is done with S
  if (threadIdx.x < (int)(sizeof(CProg) / 4))
    reinterpret_cast<int*>(&S.cp)[threadIdx.x] = reinterpret_cast<const int*>(&b.cprog[slot])[threadIdx.x];
  __syncthreads();
  const int item = b.ids[slot];
  const int th0 = b.theta_off[item];
  const int L = S.cp.L;
  if (threadIdx.x < L) leaf_constants(S.cp.leaf[threadIdx.x].x, b.theta + th0 + S.cp.leaf[threadIdx.x].z, S.cst[threadIdx.x], S.ivar[threadIdx.x]);
  if (threadIdx.x == 32) S.noise = b.theta[b.theta_off[item + 1] - 1];
  if (threadIdx.x >= 64 && threadIdx.x < 128) S.etab[threadIdx.x - 64] = FM_EXPTAB[threadIdx.x - 64];
  __syncthreads();
  if (threadIdx.x < NB) {        // row-side tables of the 128 rows of tile row ti
    const int r = threadIdx.x;
#pragma unroll
    for (int l = 0; l < LMAX; l++)
      if (l < L) {
        const int4 lf = S.cp.leaf[l];
        double A, Sv, Cv;
        point_values(lf.x, S.cst[l], b.Xt[(size_t)lf.y * b.Np + ti * NB + r], A, Sv, Cv);
        S.Ai[l][r] = A; S.Si[l][r] = Sv; S.Ci[l][r] = Cv;
      }
  }
#pragma unroll
  for (int l = 0; l < LMAX; l++) {
    Aj[l] = 0.0; Sj[l] = 0.0; Cj[l] = 0.0;
    ops[l] = OP_CONST;
    if (l < L) {
      const int4 lf = S.cp.leaf[l];
      ops[l] = lf.x;
      point_values(lf.x, S.cst[l], b.Xt[(size_t)lf.y * b.Np + tj * NB + colc], Aj[l], Sj[l], Cj[l]);   // this thread's column
    }
  }
}

// p[e] = product of the leaf values selected by the slot mask `m` (1 for the empty mask): the term loop of the 5-8 leaf
// kernels (the <= 4-leaf kernels evaluate their polynomial by canonical shape, below).
template <int LMAX, int E>
__device__ __forceinline__ void mask_product(const double (&v)[LMAX][E], unsigned m, double (&p)[E]) {
#pragma unroll
  for (int e = 0; e < E; e++) p[e] = 1.0;
#pragma unroll
  for (int l = 0; l < LMAX; l++)
    if (m & (1u << l)) {
#pragma unroll
      for (int e = 0; e < E; e++) p[e] *= v[l][e];
    }
}

// K and the adjoints dK/dv_l of the canonical <= 4-leaf shapes (cprog.cuh), in factored form: one switch per row step.
#define SH_ALL(expr) _Pragma("unroll") for (int e = 0; e < E; e++) { const double a = v[0][e], b = v[1][e], c = v[2][e], d = v[3][e]; (void)a; (void)b; (void)c; (void)d; expr; }
template <int E>
__device__ __forceinline__ void shape_value(int shape, const double (&v)[4][E], double (&k)[E]) {
  switch (shape) {
    case 0:  SH_ALL(k[e] = a) break;
    case 1:  SH_ALL(k[e] = a + b) break;
    case 2:  SH_ALL(k[e] = a * b) break;
    case 3:  SH_ALL(k[e] = (a + b) + c) break;
    case 4:  SH_ALL(k[e] = fma(b, c, a)) break;
    case 5:  SH_ALL(k[e] = a * (b + c)) break;
    case 6:  SH_ALL(k[e] = (a * b) * c) break;
    case 7:  SH_ALL(k[e] = (a + b) + (c + d)) break;
    case 8:  SH_ALL(k[e] = fma(c, d, a + b)) break;
    case 9:  SH_ALL(k[e] = fma(b, c + d, a)) break;
    case 10: SH_ALL(k[e] = fma(b * c, d, a)) break;
    case 11: SH_ALL(k[e] = a * ((b + c) + d)) break;
    case 12: SH_ALL(k[e] = (a + d) * (b + c)) break;
    case 13: SH_ALL(k[e] = fma(a, b, c * d)) break;
    case 14: SH_ALL(k[e] = a * fma(c, d, b)) break;
    case 15: SH_ALL(k[e] = (a * b) * (c + d)) break;
    default: SH_ALL(k[e] = (a * b) * (c * d)) break;
  }
}

// g[l][e] = w[e] * dK/dv_l
template <int E>
__device__ __forceinline__ void shape_adjoints(int shape, const double (&v)[4][E], const double (&w)[E], double (&g)[4][E]) {
#define SH_G(g0, g1, g2, g3) SH_ALL(g[0][e] = g0; g[1][e] = g1; g[2][e] = g2; g[3][e] = g3)
  switch (shape) {
    case 0: case 1: case 3: case 7: SH_G(w[e], w[e], w[e], w[e]) break;                     // pure sums
    case 2:  SH_G(w[e] * b, w[e] * a, 0.0, 0.0) break;
    case 4:  SH_G(w[e], w[e] * c, w[e] * b, 0.0) break;
    case 5:  { SH_ALL(const double wa = w[e] * a; g[0][e] = w[e] * (b + c); g[1][e] = wa; g[2][e] = wa; g[3][e] = 0.0) } break;
    case 6:  { SH_ALL(const double wc = w[e] * c; g[0][e] = wc * b; g[1][e] = wc * a; g[2][e] = w[e] * (a * b); g[3][e] = 0.0) } break;
    case 8:  SH_G(w[e], w[e], w[e] * d, w[e] * c) break;
    case 9:  { SH_ALL(const double wb = w[e] * b; g[0][e] = w[e]; g[1][e] = w[e] * (c + d); g[2][e] = wb; g[3][e] = wb) } break;
    case 10: { SH_ALL(const double wd = w[e] * d; g[0][e] = w[e]; g[1][e] = wd * c; g[2][e] = wd * b; g[3][e] = w[e] * (b * c)) } break;
    case 11: { SH_ALL(const double wa = w[e] * a; g[0][e] = w[e] * ((b + c) + d); g[1][e] = wa; g[2][e] = wa; g[3][e] = wa) } break;
    case 12: { SH_ALL(const double wt_ = w[e] * (b + c); const double ws = w[e] * (a + d); g[0][e] = wt_; g[1][e] = ws; g[2][e] = ws; g[3][e] = wt_) } break;
    case 13: SH_G(w[e] * b, w[e] * a, w[e] * d, w[e] * c) break;
    case 14: { SH_ALL(const double wa = w[e] * a; g[0][e] = w[e] * fma(c, d, b); g[1][e] = wa; g[2][e] = wa * d; g[3][e] = wa * c) } break;
    case 15: { SH_ALL(const double wt_ = w[e] * (c + d); const double wab = w[e] * (a * b); g[0][e] = wt_ * b; g[1][e] = wt_ * a; g[2][e] = wab; g[3][e] = wab) } break;
    default: { SH_ALL(const double wab = w[e] * (a * b); const double wcd = w[e] * (c * d); g[0][e] = wcd * b; g[1][e] = wcd * a; g[2][e] = wab * d; g[3][e] = wab * c) } break;
  }
#undef SH_G
}
#undef SH_ALL

// product-term mask g of the slot's program (5-8 leaf kernels): a byte of the CProg copy in shared memory
__device__ __forceinline__ unsigned term_mask(unsigned term_addr, int g) {
  unsigned m;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(m) : "r"(term_addr + g));
  return m;
}

template <int LMAX, int E>
__global__ void __launch_bounds__(256, LMAX == 4 ? POLY_BUILD_CTAS : 2) poly_build_kernel(Batch b, int kind) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef PolySmem<LMAX> SM;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int count = b.kcount[kind];
  const int* list = b.klist + kind * b.list_stride;
  const int ntiles = tri_count(b.T);
  const long long nwork = (long long)count * ntiles;
  const int c = threadIdx.x & (NB - 1), h = threadIdx.x >> 7;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int slot = list[w / ntiles];
    int ti, tj;
    tri_decode((int)(w % ntiles), ti, tj);
    double Aj[LMAX], Sj[LMAX], Cj[LMAX];
    int ops[LMAX];
    poly_prologue<LMAX>(S, b, slot, ti, tj, c, Aj, Sj, Cj, ops);
    __syncthreads();
    const int L = S.cp.L, G = S.cp.G, shape = S.cp.shape;
    const double diag_add = b.add_diag == 1 ? S.noise + EXACT_INFERENCE_JITTER + (b.jitter ? b.jitter[slot] : 0.0)
                                             : (b.add_diag == 2 ? S.noise : 0.0);
    double* out = b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB + c;
    const int gj = tj * NB + c;
    const bool interior = ti > tj && (ti + 1) * NB <= b.N;
#pragma unroll 1
    for (int r0 = h * 64; r0 < h * 64 + 64; r0 += E) {
      double v[LMAX][E], aux[E];
#pragma unroll
      for (int l = 0; l < LMAX; l++)
        if (l < L) {
          double Ar[E];
          lds_rows<E>(sb + (unsigned)offsetof(SM, Ai) + (l * NB + r0) * 8, Ar);
          leaf_fwd<E>(ops[l], sb + (unsigned)offsetof(SM, cst) + l * 32, sb + (unsigned)offsetof(SM, etab), Ar, sb + (unsigned)offsetof(SM, Si) + (l * NB + r0) * 8,
                      sb + (unsigned)offsetof(SM, Ci) + (l * NB + r0) * 8, Aj[l], Sj[l], Cj[l], v[l], aux);
        } else if (LMAX == 4) {
#pragma unroll
          for (int e = 0; e < E; e++) v[l][e] = 0.0;       // the shape formulas name all four slots
        }
      double k[E];
      if constexpr (LMAX == 4) {
        shape_value<E>(shape, v, k);
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) k[e] = 0.0;
        for (int g = 0; g < G; g++) {
          const unsigned mask = term_mask(sb + (unsigned)offsetof(SM, cp) + (unsigned)offsetof(CProg, term), g);
          double p[E];
          mask_product<LMAX, E>(v, mask, p);
#pragma unroll
          for (int e = 0; e < E; e++) k[e] += p[e];
        }
      }
      if (interior) {        // tile strictly below the diagonal and inside N: no padding, no diagonal element
#pragma unroll
        for (int e = 0; e < E; e++) out[(size_t)(r0 + e) * b.Np] = k[e];
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) {
          const int gi = ti * NB + r0 + e;
          double val = k[e];
          if (gi >= b.N || gj >= b.N) val = (gi == gj) ? 1.0 : 0.0;
          else if (gi == gj) { val += diag_add; S.ai[c] = val; }
          out[(size_t)(r0 + e) * b.Np] = val;
        }
      }
    }
    if (ti == tj) {    // deterministic partial of trace(K_y) for the jitter ladder
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < NB; t++)
          if (ti * NB + t < b.N) s += S.ai[t];
        b.diag_part[slot * b.T + ti] = s;
      }
    }
  }
}

// Gradient contribution of the ROWS x COLS block of tile (ti, tj) whose first element is (rbase, col0), by a CTA of THREADS
// threads.  Thread (c, h) = (tid % COLS, tid / COLS) owns column col0 + c and the rows [rbase + h RPH, rbase + (h + 1) RPH),
// RPH = ROWS COLS / THREADS.  `Kt` points at (row 0 of the TILE, column col0) of K^-1, `kld` is its row stride: global memory
// (Np) for the stand-alone kernel, the staged copy in shared memory for the fused K^-1 + gradient kernels.  `out`: this work
// item's entry of grad_part.
template <int LMAX, int E, int COLS, int THREADS = 256, int ROWS = NB>
__device__ __forceinline__ void poly_grad_tile(PolySmem<LMAX>& S, unsigned sb, const Batch& b, int slot, int ti, int tj, int col0,
                                               const double* Kt, size_t kld, double* out, int rbase = 0) {
  typedef PolySmem<LMAX> SM;
  constexpr int NV = 3 * LMAX + 1;
  constexpr int RPH = ROWS * COLS / THREADS;
  const int c = threadIdx.x % COLS, h = threadIdx.x / COLS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double Aj[LMAX], Sj[LMAX], Cj[LMAX];
  int ops[LMAX];
  poly_prologue<LMAX>(S, b, slot, ti, tj, col0 + c, Aj, Sj, Cj, ops);
  if (THREADS >= 2 * NB) {
    if (threadIdx.x >= NB && threadIdx.x < 2 * NB) S.ai[threadIdx.x - NB] = b.alpha[(size_t)slot * b.Np + ti * NB + threadIdx.x - NB];
  } else {
    for (int r = threadIdx.x; r < NB; r += THREADS) S.ai[r] = b.alpha[(size_t)slot * b.Np + ti * NB + r];
  }
  const double aj = b.alpha[(size_t)slot * b.Np + tj * NB + col0 + c];
  __syncthreads();
  const int L = S.cp.L, G = S.cp.G, single = S.cp.single, shape = S.cp.shape;
  double acc[LMAX][3], accn = 0.0;
#pragma unroll
  for (int l = 0; l < LMAX; l++) { acc[l][0] = 0.0; acc[l][1] = 0.0; acc[l][2] = 0.0; }
  const double* Kinv = Kt + c;
  const int gj = tj * NB + col0 + c;
  // diagonal tiles: rows above this warp's first column hold no lower-triangle element
  int rbeg = rbase + h * RPH;
  if (ti == tj) { const int cmin = col0 + c - lane; if (cmin > rbeg) rbeg = cmin / E * E; }
  const int rend = rbase + h * RPH + RPH;
  const bool interior = ti > tj && (ti + 1) * NB <= b.N;
  double kv[E];
#pragma unroll
  for (int e = 0; e < E; e++) kv[e] = (rbeg < rend) ? Kinv[(size_t)(rbeg + e) * kld] : 0.0;
#pragma unroll 1
  for (int r0 = rbeg; r0 < rend; r0 += E) {
    double wt[E], ar[E];
    lds_rows<E>(sb + (unsigned)offsetof(SM, ai) + r0 * 8, ar);
    if (interior) {          // tile strictly below the diagonal and inside N: every element is a valid symmetric pair
#pragma unroll
      for (int e = 0; e < E; e++) wt[e] = ar[e] * aj - kv[e];          // 2 * 0.5 * (alpha_i alpha_j - K^-1_ij)
    } else {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int gi = ti * NB + r0 + e;
        double wv = 0.5 * (ar[e] * aj - kv[e]);
        const bool valid = gi < b.N && gj <= gi;       // gj <= gi < N
        if (gi == gj) { if (valid) accn += wv; } else wv *= 2.0;      // symmetric pair (i, j), (j, i)
        wt[e] = valid ? wv : 0.0;
      }
    }
    if (r0 + E < rend) {
#pragma unroll
      for (int e = 0; e < E; e++) kv[e] = Kinv[(size_t)(r0 + E + e) * kld];
    }
    double v[LMAX][E], aux[LMAX][E];
#pragma unroll
    for (int l = 0; l < LMAX; l++)
      if (l < L) {
        double Ar[E];
        lds_rows<E>(sb + (unsigned)offsetof(SM, Ai) + (l * NB + r0) * 8, Ar);
        leaf_fwd<E>(ops[l], sb + (unsigned)offsetof(SM, cst) + l * 32, sb + (unsigned)offsetof(SM, etab), Ar, sb + (unsigned)offsetof(SM, Si) + (l * NB + r0) * 8,
                    sb + (unsigned)offsetof(SM, Ci) + (l * NB + r0) * 8, Aj[l], Sj[l], Cj[l], v[l], aux[l]);
      } else if (LMAX == 4) {
#pragma unroll
        for (int e = 0; e < E; e++) v[l][e] = 0.0;         // the shape formulas name all four slots
      }
    double g4[LMAX == 4 ? 4 : 1][E];
    if constexpr (LMAX == 4) shape_adjoints<E>(shape, v, wt, g4);
#pragma unroll
    for (int l = 0; l < LMAX; l++)
      if (l < L) {
        double g[E];
        if constexpr (LMAX == 4) {
#pragma unroll
          for (int e = 0; e < E; e++) g[e] = g4[l][e];
        } else if (single) {
#pragma unroll
          for (int e = 0; e < E; e++) g[e] = wt[e];
        } else {
          double adj[E];
#pragma unroll
          for (int e = 0; e < E; e++) adj[e] = 0.0;
          {
            for (int t = 0; t < G; t++) {
              const unsigned mask = term_mask(sb + (unsigned)offsetof(SM, cp) + (unsigned)offsetof(CProg, term), t);
              if (!(mask & (1u << l))) continue;
              double p[E];
              mask_product<LMAX, E>(v, mask & ~(1u << l), p);
#pragma unroll
              for (int e = 0; e < E; e++) adj[e] += p[e];
            }
          }
#pragma unroll
          for (int e = 0; e < E; e++) g[e] = wt[e] * adj[e];
        }
        double Ar[E];
        lds_rows<E>(sb + (unsigned)offsetof(SM, Ai) + (l * NB + r0) * 8, Ar);
        leaf_bwd<E>(ops[l], sb + (unsigned)offsetof(SM, cst) + l * 32, Ar, sb + (unsigned)offsetof(SM, Si) + (l * NB + r0) * 8,
                    sb + (unsigned)offsetof(SM, Ci) + (l * NB + r0) * 8, Aj[l], Sj[l], Cj[l], g, v[l], aux[l], acc[l]);
      }
  }
  // CTA reduction in a fixed order: lanes (shuffle tree), then the 8 warps
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = (i == NV - 1) ? accn : acc[i / 3][i % 3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) S.red[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int i = threadIdx.x;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < THREADS / 32; q++) s += S.red[q][i];
    if (i == NV - 1) out[S.cp.nparam] = s;
    else {
      const int l = i / 3, p = i % 3;
      if (l < L && p < S.cp.leaf[l].w)
        out[S.cp.leaf[l].z + p] = s * leaf_grad_scale(S.cp.leaf[l].x, S.cst[l], S.ivar[l], p, b.quirks);
    }
  }
}

// grad_part holds GP_PER_TILE entries per (slot, tile): the fused kernels write one per 64 x 64 quadrant (or per 128 x 64 half
// tile), the stand-alone kernels fill entry 0; whoever writes entry 0 clears the entries nobody else writes.  finalize_kernel
// sums all of them in index order.
constexpr int GP_PER_TILE = 4;

template <int LMAX, int E>
__global__ void __launch_bounds__(256, LMAX == 4 ? POLY_GRAD_CTAS : POLY_GRAD8_CTAS) poly_grad_kernel(Batch b, int kind) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef PolySmem<LMAX> SM;
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int count = b.kcount[kind];
  const int* list = b.klist + kind * b.list_stride;
  const int ntiles = tri_count(b.T);
  const long long nwork = (long long)count * ntiles;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int slot = list[w / ntiles];
    if (b.slot_status[slot] != 0) continue;
    const int tile = (int)(w % ntiles);
    int ti, tj;
    tri_decode(tile, ti, tj);
    double* out = b.grad_part + ((size_t)slot * ntiles + tile) * GP_PER_TILE * b.pstride;
    if (threadIdx.x < b.pstride)
      for (int q = 1; q < GP_PER_TILE; q++) out[q * b.pstride + threadIdx.x] = 0.0;
    poly_grad_tile<LMAX, E, NB>(S, sb, b, slot, ti, tj, 0, b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB, b.Np, out);
  }
}

// Fused K^-1 + gradient for the <= 4-leaf programs: the K^-1 half tile a CTA has just accumulated on the DMMA pipe is
// staged in the (now idle) pipeline buffers and contracted with dK/dtheta right there.  K^-1 never goes to HBM, and the
// element-wise FP64 work of one CTA runs while the co-resident CTA of the SM is in its DMMA main loop -- the FP64 ALU
// and the DMMA path are one datapath on B200 (tools/pipe_probe), so this fills the bubbles both kernels had on their own
// (kinv_kernel: pipe 83 % busy; poly_grad_kernel: 43 %).  Slots of other kinds get K^-1 stored as before.
constexpr int KG_LD = TG_TN + 2;         // row length of the staged 128 x 64 half tile
static_assert(NB * KG_LD * 8 + sizeof(PolySmem<4>) <= TG_SMEM_BYTES, "staged tile + gradient tables fit the pipeline buffers");

__global__ void __launch_bounds__(TG_THREADS, 2) kinv_grad_kernel(Batch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int bx = blockIdx.x >> 1, hf = blockIdx.x & 1;
  const int slot = b.slot0 + blockIdx.y;
  double acc[TG_MT][TG_NT][2];
  GemmOut o;
  if (!gemm_compute<GM_KINV>(b, 0, UpdRange{0, 0, 0, 0}, bx, hf, reinterpret_cast<double*>(smem_raw), acc, o)) return;
  if (b.cprog[slot].kind != CK_POLY4) {      // uniform per CTA
    gemm_store(acc, o);
    return;
  }
  __syncthreads();                           // every warp has left the pipeline stages
  double* Kt = reinterpret_cast<double*>(smem_raw);
  if (!o.skip) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int wm, wn;
    warp_coords(warp, wm, wn);
    double* dst = Kt + (wm * 32 + (lane >> 2)) * KG_LD + wn * 32 + (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++)
        *reinterpret_cast<double2*>(dst + (mt * 8) * KG_LD + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
  }
  // (poly_grad_tile starts with a barrier: the staged tile is complete before anybody reads it)
  PolySmem<4>& S = *reinterpret_cast<PolySmem<4>*>(smem_raw + NB * KG_LD * 8);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(&S);
  int ti, tj;
  tri_decode(bx, ti, tj);
  const int ntiles = tri_count(b.T);
  double* out = b.grad_part + (((size_t)slot * ntiles + bx) * GP_PER_TILE + hf) * b.pstride;
  if (hf == 0 && threadIdx.x < b.pstride) { out[2 * b.pstride + threadIdx.x] = 0.0; out[3 * b.pstride + threadIdx.x] = 0.0; }
  poly_grad_tile<4, POLY_GE4, TG_TN>(S, sb, b, slot, ti, tj, hf * TG_TN, Kt, KG_LD, out);
}

// The same fusion on the 64 x 64 tiles of kinv64_kernel: three 4-warp CTAs per SM, so while one CTA contracts its K^-1 block
// with dK/dtheta (element-wise FP64 work, latency-bound on its own) two others keep the DMMA pipe fed.
constexpr int KG64_LD = K64_T + 2;
static_assert(K64_T * KG64_LD * 8 + sizeof(PolySmem<4>) <= K64_SMEM_BYTES, "staged block + gradient tables fit the pipeline buffers");

__global__ void __launch_bounds__(K64_THREADS, 3) kinv64_grad_kernel(Batch b) {
  const int slot = b.slot0 + blockIdx.y;
  if (b.slot_status[slot] != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int I, J;
  tri_decode(blockIdx.x, I, J);
  const size_t ld = b.Np;
  const double* V = b.M2 + slot * b.mstride;
  double acc[TG_MT][TG_NT][2];
  tile64_init(acc, nullptr, ld, 0.0);
  tile64_pipeline(acc, V + (size_t)(I * K64_T) * ld + I * K64_T, V + (size_t)(J * K64_T) * ld + I * K64_T, ld,
                  (b.Np - I * K64_T) / TG_KC, reinterpret_cast<double*>(smem_raw));
  if (b.cprog[slot].kind != CK_POLY4) {      // uniform per CTA: K^-1 goes to HBM for the 5-8 leaf / interpreter kernels
    tile64_store(acc, tile64_cw(b.M1 + slot * b.mstride + (size_t)(I * K64_T) * ld + J * K64_T, ld), ld, 1.0);
    return;
  }
  __syncthreads();                           // every warp has left the pipeline stages
  double* Kt = reinterpret_cast<double*>(smem_raw);
  tile64_store(acc, tile64_cw(Kt, KG64_LD), KG64_LD, 1.0);
  // (poly_grad_tile starts with a barrier: the staged block is complete before anybody reads it)
  PolySmem<4>& S = *reinterpret_cast<PolySmem<4>*>(smem_raw + K64_T * KG64_LD * 8);
  const unsigned sb = (unsigned)__cvta_generic_to_shared(&S);
  const int ti = I >> 1, vi = I & 1, tj = J >> 1, vj = J & 1;
  const int ntiles = tri_count(b.T);
  double* out = b.grad_part + ((size_t)slot * ntiles + tri_count(ti) + tj) * GP_PER_TILE * b.pstride;
  if (ti == tj && vi == 0 && threadIdx.x < b.pstride) out[1 * b.pstride + threadIdx.x] = 0.0;    // quadrant (0, 1) of a diagonal tile has no CTA
  poly_grad_tile<4, POLY_GE4, K64_T, K64_THREADS, K64_T>(S, sb, b, slot, ti, tj, vj * K64_T, Kt - (size_t)(vi * K64_T) * KG64_LD,
                                                         KG64_LD, out + (vi * 2 + vj) * b.pstride, vi * K64_T);
}

void launch_kinv_grad(const Batch& b, cudaStream_t s) {
  if (b.gemm64) {
    dim3 grid(tri_count(2 * b.T), b.nslots);
    kinv64_grad_kernel<<<grid, K64_THREADS, K64_SMEM_BYTES, s>>>(b);
  } else {
    dim3 grid(2 * tri_count(b.T), b.nslots);
    kinv_grad_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, s>>>(b);
  }
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// Interpreter kernels (programs the polynomial form cannot hold): one 128 x 128 tile per work item.
// ------------------------------------------------------------------------------------------------
struct ElemSmem {
  ProgSmem P;
  double xi[MAX_DIMS * NB];
  double xj[MAX_DIMS * NB];
  double ai[NB], aj[NB];
  double red[256];
};

__device__ __forceinline__ void stage_inputs(ElemSmem& S, const Batch& b, int ti, int tj) {
  for (int idx = threadIdx.x; idx < b.D * NB; idx += blockDim.x) {
    const int d = idx / NB, r = idx % NB;
    S.xi[idx] = b.Xt[(size_t)d * b.Np + ti * NB + r];
    S.xj[idx] = b.Xt[(size_t)d * b.Np + tj * NB + r];
  }
}

__global__ void __launch_bounds__(256) interp_build_kernel(Batch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemSmem& S = *reinterpret_cast<ElemSmem*>(smem_raw);
  const int count = b.kcount[CK_GENERIC];
  const int* list = b.klist + CK_GENERIC * b.list_stride;
  const int ntiles = tri_count(b.T);
  const long long nwork = (long long)count * ntiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int slot = list[w / ntiles], item = b.ids[slot];
    int ti, tj;
    tri_decode((int)(w % ntiles), ti, tj);
    __syncthreads();
    load_program(S.P, b.prog, b.prog_off, b.theta, b.theta_off, item, b.lut, b.lut_n);
    stage_inputs(S, b, ti, tj);
    __syncthreads();
    const double diag_add = b.add_diag == 1 ? S.P.noise + EXACT_INFERENCE_JITTER + (b.jitter ? b.jitter[slot] : 0.0)
                                             : (b.add_diag == 2 ? S.P.noise : 0.0);
    double* out = b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB;
    double dsum = 0.0;
    // thread -> 64 elements: rows warp + 8k (k < 16), columns lane + 32m (m < 4): every warp store is one
    // 256-byte row segment.  One copy of the interpreter in the instruction stream (no unrolling).
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
      const int r = warp + 8 * (it >> 2), c = lane + 32 * (it & 3);
      const int gi = ti * NB + r, gj = tj * NB + c;
      double v;
      if (gi >= b.N || gj >= b.N) {
        v = (gi == gj) ? 1.0 : 0.0;
      } else {
        v = eval_K(S.P, S.xi + r, S.xj + c, NB);
        if (gi == gj) { v += diag_add; dsum += v; }
      }
      out[(size_t)r * b.Np + c] = v;
    }
    if (ti == tj) {
      S.red[threadIdx.x] = dsum;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < 256; t++) s += S.red[t];
        b.diag_part[slot * b.T + ti] = s;
      }
    }
  }
}

__global__ void __launch_bounds__(256) interp_grad_kernel(Batch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemSmem& S = *reinterpret_cast<ElemSmem*>(smem_raw);
  double* wred = reinterpret_cast<double*>(smem_raw + sizeof(ElemSmem));   // [8][pstride]
  const int count = b.kcount[CK_GENERIC];
  const int* list = b.klist + CK_GENERIC * b.list_stride;
  const int ntiles = tri_count(b.T);
  const long long nwork = (long long)count * ntiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int slot = list[w / ntiles], item = b.ids[slot];
    if (b.slot_status[slot] != 0) continue;
    const int tile = (int)(w % ntiles);
    int ti, tj;
    tri_decode(tile, ti, tj);
    __syncthreads();
    load_program(S.P, b.prog, b.prog_off, b.theta, b.theta_off, item, b.lut, b.lut_n);
    stage_inputs(S, b, ti, tj);
    if (threadIdx.x < NB) {
      S.ai[threadIdx.x] = b.alpha[(size_t)slot * b.Np + ti * NB + threadIdx.x];
      S.aj[threadIdx.x] = b.alpha[(size_t)slot * b.Np + tj * NB + threadIdx.x];
    }
    __syncthreads();
    const int np = S.P.nparam;
    double acc[MAX_PARAMS + 1];
    for (int p = 0; p <= np; p++) acc[p] = 0.0;
    const double* Kinv = b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB;
    double kv_next = Kinv[(size_t)warp * b.Np + lane];
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
      const int r = warp + 8 * (it >> 2), c = lane + 32 * (it & 3);
      const double kv = kv_next;
      if (it + 1 < 64) kv_next = Kinv[(size_t)(warp + 8 * ((it + 1) >> 2)) * b.Np + lane + 32 * ((it + 1) & 3)];
      const int gi = ti * NB + r, gj = tj * NB + c;
      if (gi >= b.N || gj >= b.N || gj > gi) continue;
      double wv = 0.5 * (S.ai[r] * S.aj[c] - kv);
      if (gi == gj) acc[np] += wv;     // dL/d(noise) = trace(dL/dK)
      else wv *= 2.0;                  // symmetric pair (i, j) and (j, i)
      eval_grad(S.P, S.xi + r, S.xj + c, NB, wv, acc, b.quirks);
    }
    for (int p = 0; p <= np; p++) {
      double v = acc[p];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) wred[warp * b.pstride + p] = v;
    }
    __syncthreads();
    double* out = b.grad_part + ((size_t)slot * ntiles + tile) * GP_PER_TILE * b.pstride;
    for (int p = threadIdx.x; p <= np; p += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; q++) s += wred[q * b.pstride + p];
      out[p] = s;
      for (int q = 1; q < GP_PER_TILE; q++) out[q * b.pstride + p] = 0.0;
    }
  }
}


void launch_build_k(const Batch& b, cudaStream_t s) {
  poly_build_kernel<4, POLY_E4><<<g_grid_poly_build[0], 256, sizeof(PolySmem<4>), s>>>(b, CK_POLY4);
  poly_build_kernel<8, POLY_E8><<<g_grid_poly_build[1], 256, sizeof(PolySmem<8>), s>>>(b, CK_POLY8);
  interp_build_kernel<<<g_grid_interp_build, 256, sizeof(ElemSmem), s>>>(b);
  g_launch_count += 3;
}

void launch_grad(const Batch& b, cudaStream_t s, bool poly4_done) {
  if (!poly4_done) poly_grad_kernel<4, POLY_GE4><<<g_grid_poly_grad[0], 256, sizeof(PolySmem<4>), s>>>(b, CK_POLY4);
  poly_grad_kernel<8, POLY_E8><<<g_grid_poly_grad[1], 256, sizeof(PolySmem<8>), s>>>(b, CK_POLY8);
  const size_t smem = sizeof(ElemSmem) + 8 * (size_t)b.pstride * sizeof(double);
  interp_grad_kernel<<<g_grid_interp_grad, 256, smem, s>>>(b);
  g_launch_count += 3;
}

// Symmetric dense copy of the lower tiles of M1 (for b200gp_build_K / parity tests).
__global__ void export_sym_kernel(Batch b, double* out) {
  const int slot = blockIdx.y;
  const size_t n2 = (size_t)b.N * b.N;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n2; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / b.N), j = (int)(idx % b.N);
    const int r = i > j ? i : j, c = i > j ? j : i;
    out[slot * n2 + idx] = b.M1[slot * b.mstride + (size_t)r * b.Np + c];
  }
}

void launch_export_sym(const Batch& b, double* out, cudaStream_t s) {
  dim3 grid(592, b.nslots);
  export_sym_kernel<<<grid, 256, 0, s>>>(b, out);
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// Finalize: LML = -1/2 (N log 2pi + logdet + y^T K^-1 y); tile partials -> gradient; status.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(Batch b, int with_grad) {
  const int slot = blockIdx.x, item = b.ids[slot];
  __shared__ double red[256];
  __shared__ int bad;
  const int th0 = b.theta_off[item];
  const int np = b.theta_off[item + 1] - th0 - 1;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (b.slot_status[slot] != 0) {
    if (threadIdx.x == 0) { b.lml[item] = nan(""); b.status[item] = ST_NOT_PD; if (b.tries) b.tries[item] = b.try_no; }
    if (with_grad)
      for (int p = threadIdx.x; p <= np; p += blockDim.x) b.dlml[th0 + p] = nan("");
    return;
  }
  const double* z = b.z + (size_t)slot * b.Np;
  double s = 0.0;
  for (int i = threadIdx.x; i < b.N; i += blockDim.x) s = fma(z[i], z[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double logdet = 0.0;
    for (int k = 0; k < b.T; k++) logdet += b.logdet_part[slot * b.T + k];
    const double lml = -0.5 * (b.N * 1.8378770664093453 + logdet + red[0]);
    b.lml[item] = lml;
    if (!(fabs(lml) < 1.7e308)) bad = 1;
  }
  if (with_grad) {
    const int ntiles = tri_count(b.T) * GP_PER_TILE;
    const double* gp = b.grad_part + (size_t)slot * ntiles * b.pstride;
    // one warp per parameter: lane l sums tiles l, l + 32, ... in order, then a fixed shuffle tree (the order never
    // depends on the batch around this item)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p <= np; p += 8) {
      double g = 0.0;
      for (int t = lane; t < ntiles; t += 32) g += gp[(size_t)t * b.pstride + p];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane == 0) {
        b.dlml[th0 + p] = g;
        if (!(fabs(g) < 1.7e308)) bad = 1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    b.status[item] = bad ? ST_NONFINITE : (b.try_no > 0 ? ST_JITTERED : ST_OK);
    if (b.tries) b.tries[item] = b.try_no;
  }
}

void launch_finalize(const Batch& b, int with_grad, cudaStream_t s) {
  finalize_kernel<<<b.nslots, 256, 0, s>>>(b, with_grad);
  g_launch_count++;
}

// X (N x D row-major) -> Xt (D x Np, zero padded); y -> yp (Np, zero padded)
__global__ void pack_data_kernel(const double* X, const double* y, int N, int D, int Np, double* Xt, double* yp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  for (int d = 0; d < D; d++) Xt[(size_t)d * Np + i] = (i < N) ? X[(size_t)i * D + d] : 0.0;
  yp[i] = (i < N) ? y[i] : 0.0;
}

void launch_pack_data(const double* X, const double* y, int N, int D, int Np, double* Xt, double* yp, cudaStream_t s) {
  pack_data_kernel<<<(Np + 255) / 256, 256, 0, s>>>(X, y, N, D, Np, Xt, yp);
  g_launch_count++;
}

int configure_chol_kernels();
int configure_kernels() {
  cudaError_t e;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  struct Cfg { const void* fn; int smem; int* grid; };
  const int interp_grad_smem = (int)(sizeof(ElemSmem) + 8 * (MAX_PARAMS + 1) * sizeof(double));
  const Cfg cfgs[] = {
      {(const void*)poly_build_kernel<4, POLY_E4>, (int)sizeof(PolySmem<4>), &g_grid_poly_build[0]},
      {(const void*)poly_build_kernel<8, POLY_E8>, (int)sizeof(PolySmem<8>), &g_grid_poly_build[1]},
      {(const void*)poly_grad_kernel<4, POLY_GE4>, (int)sizeof(PolySmem<4>), &g_grid_poly_grad[0]},
      {(const void*)poly_grad_kernel<8, POLY_E8>, (int)sizeof(PolySmem<8>), &g_grid_poly_grad[1]},
      {(const void*)interp_build_kernel, (int)sizeof(ElemSmem), &g_grid_interp_build},
      {(const void*)interp_grad_kernel, interp_grad_smem, &g_grid_interp_grad},
  };
  for (const Cfg& c : cfgs) {
    e = cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c.fn, 256, c.smem);
    if (e != cudaSuccess) return (int)e;
    *c.grid = g_sm_count * (per_sm > 0 ? per_sm : 1);
  }
  e = cudaFuncSetAttribute(kinv_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(kinv64_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K64_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  return configure_chol_kernels();
}

}  // namespace b200gp
