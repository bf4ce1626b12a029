// Kernel-program interpreter: evaluates a compositional kernel (sums / products of 1-D
// SE / RQ / PER / LIN / CONST leaves) and its hyperparameter derivatives for one (x_i, x_j) pair.
//
// Replaces, on the device, the GPy-fork leaf kernels bound by the reference at
// src/autoks/backend/kernel.py:17-22 (formulas: notebooks/exploratory/test custom kernels/
// {rational-quadratic,periodic,linear}.ipynb#2, rbf.ipynb#1-2) and GPy's Add / Prod fan-out.
// Control flow is uniform across the CTA: one CTA works on one candidate's program.
#pragma once
#include "common.cuh"

namespace b200gp {

struct ProgSmem {
  int4 tok[MAX_TOKENS];
  double cst[MAX_TOKENS][4];   // per-leaf derived constants (reciprocals hoisted out of the N^2 loop)
  double ivar[MAX_TOKENS];     // 1 / variance of each leaf (dK/dvariance = K / variance)
  int ntok;
  int nparam;                  // kernel parameters (noise excluded)
  double noise;                // Gaussian_noise.variance
  const double* lut;           // OP_LOOKUP: lut_n x lut_n distance matrix (b200gp_set_lookup), else null
  int lut_n;
};

// distance between the models whose indices are stored (as doubles) in a and b; NaN outside the table
__device__ __forceinline__ double lookup_distance(const ProgSmem& P, double a, double b) {
  const int ia = (int)a, ib = (int)b;
  if (P.lut == nullptr || ia < 0 || ib < 0 || ia >= P.lut_n || ib >= P.lut_n) return nan("");
  return P.lut[(size_t)ia * P.lut_n + ib];
}

// Cooperative load of item `item`'s program + theta into shared memory; derives leaf constants.
// `th_base` points at the item's first kernel parameter; `noise` is its Gaussian noise variance.
__device__ inline void load_program_at(ProgSmem& P, const int4* __restrict__ toks, int nt,
                                       const double* __restrict__ th_base, int nparam, double noise,
                                       const double* lut = nullptr, int lut_n = 0) {
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    int4 k = toks[t];
    P.tok[t] = k;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (is_leaf(k.x)) {
      const double* th = th_base + k.z;
      switch (k.x) {
        case OP_LOOKUP:
        case OP_SE:  c0 = th[0]; c1 = 1.0 / th[1]; break;                                   // [variance, lengthscale]
        case OP_RQ:  c0 = th[0]; c1 = 1.0 / th[1]; c2 = th[2]; c3 = 1.0 / (2.0 * th[2]); break;  // [.., power]
        case OP_PER: c0 = th[0]; c1 = 3.14159265358979323846 / th[1]; c2 = 1.0 / th[2]; c3 = 1.0 / th[1]; break;  // [variance, period, lengthscale]
        case OP_LIN: c0 = th[0]; c1 = th[1]; break;                                          // [variances, shifts]
        case OP_CONST: c0 = th[0]; break;
      }
    }
    P.cst[t][0] = c0; P.cst[t][1] = c1; P.cst[t][2] = c2; P.cst[t][3] = c3;
    P.ivar[t] = is_leaf(k.x) ? 1.0 / c0 : 0.0;
  }
  if (threadIdx.x == 0) {
    P.ntok = nt;
    P.nparam = nparam;
    P.noise = noise;
    P.lut = lut;
    P.lut_n = lut_n;
  }
}

__device__ inline void load_program(ProgSmem& P, const int4* __restrict__ prog, const int* __restrict__ prog_off,
                                    const double* __restrict__ theta, const int* __restrict__ theta_off, int item,
                                    const double* lut = nullptr, int lut_n = 0) {
  const int t0 = prog_off[item], th0 = theta_off[item], th1 = theta_off[item + 1];
  load_program_at(P, prog + t0, prog_off[item + 1] - t0, theta + th0, th1 - th0 - 1, theta[th1 - 1], lut, lut_n);
}

// K(x_i, x_j).  xi / xj point at dimension 0 of the staged inputs; consecutive dimensions are
// `stride` doubles apart.
static __device__ __noinline__ double eval_K(const ProgSmem& P, const double* xi, const double* xj, int stride) {
  double val[MAX_TOKENS];
  const int nt = P.ntok;
  for (int t = 0; t < nt; t++) {
    const int4 k = P.tok[t];
    double v;
    if (is_leaf(k.x)) {
      const double a = xi[k.y * stride], b = xj[k.y * stride];
      const double c0 = P.cst[t][0], c1 = P.cst[t][1];
      switch (k.x) {
        case OP_LOOKUP: { double r = lookup_distance(P, a, b) * c1; v = c0 * exp(-0.5 * (r * r)); break; }
        case OP_SE: { double r = (a - b) * c1; v = c0 * exp(-0.5 * (r * r)); break; }
        case OP_RQ: { double r = (a - b) * c1; v = c0 * exp(-P.cst[t][2] * log1p((r * r) * P.cst[t][3])); break; }
        case OP_PER: { double q = sin((a - b) * c1) * P.cst[t][2]; v = c0 * exp(-2.0 * (q * q)); break; }
        case OP_LIN: v = c0 * ((a - c1) * (b - c1)); break;
        default: v = c0; break;
      }
    } else {
      const double l = val[k.y], r = val[k.z];
      v = (k.x == OP_ADD) ? l + r : l * r;
    }
    val[t] = v;
  }
  return val[nt - 1];
}

// Forward + reverse sweep for one pair: acc[p] += w * dK/dtheta_p for every kernel parameter.
// `w` is the (already symmetry-weighted) dL/dK entry.
static __device__ __noinline__ void eval_grad(const ProgSmem& P, const double* xi, const double* xj, int stride, double w,
                                          double* acc, int quirks) {
  double val[MAX_TOKENS], adj[MAX_TOKENS], ax0[MAX_TOKENS], ax1[MAX_TOKENS];
  const int nt = P.ntok;
  for (int t = 0; t < nt; t++) {
    const int4 k = P.tok[t];
    double v, a0 = 0, a1 = 0;
    if (is_leaf(k.x)) {
      const double a = xi[k.y * stride], b = xj[k.y * stride];
      const double c0 = P.cst[t][0], c1 = P.cst[t][1];
      switch (k.x) {
        case OP_LOOKUP: { double r = lookup_distance(P, a, b) * c1; a0 = r * r; a1 = exp(-0.5 * a0); v = c0 * a1; break; }
        case OP_SE: { double r = (a - b) * c1; a0 = r * r; a1 = exp(-0.5 * a0); v = c0 * a1; break; }
        case OP_RQ: { double r = (a - b) * c1; double u = (r * r) * P.cst[t][3]; a0 = log1p(u); a1 = u; v = c0 * exp(-P.cst[t][2] * a0); break; }
        case OP_PER: { double s, c; sincos((a - b) * c1, &s, &c); a0 = s; a1 = c; double q = s * P.cst[t][2]; v = c0 * exp(-2.0 * (q * q)); break; }
        case OP_LIN: { a0 = a - c1; a1 = b - c1; v = c0 * (a0 * a1); break; }
        default: v = c0; break;
      }
    } else {
      const double l = val[k.y], r = val[k.z];
      v = (k.x == OP_ADD) ? l + r : l * r;
    }
    val[t] = v; ax0[t] = a0; ax1[t] = a1;
  }
  adj[nt - 1] = w;
  for (int t = nt - 1; t >= 0; t--) {
    const int4 k = P.tok[t];
    const double g = adj[t];
    if (k.x == OP_ADD) {
      adj[k.y] = g; adj[k.z] = g;
    } else if (k.x == OP_MUL) {
      adj[k.y] = g * val[k.z]; adj[k.z] = g * val[k.y];
    } else {
      const double v = val[t];
      const double c0 = P.cst[t][0], c1 = P.cst[t][1];
      double* o = acc + k.z;
      switch (k.x) {
        case OP_LOOKUP: // same Stationary RBF gradients with r = D[i, j] / l   (kernel-kernel.ipynb#2)
        case OP_SE:     // dK/dv = K/v ; dK/dl = K r^2 / l   (stationary.py update_gradients_full)
          o[0] += g * ax1[t];
          o[1] += g * (v * ax0[t] * c1);
          break;
        case OP_RQ: {   // rational-quadratic.ipynb#2
          const double e1 = 1.0 / (1.0 + ax1[t]);           // (1 + r^2/2a)^-1
          const double r2 = ax1[t] * (2.0 * P.cst[t][2]);    // r^2
          o[0] += g * (v * P.ivar[t]);
          o[1] += g * (v * r2 * c1 * e1);
          o[2] += g * (v * (ax1[t] * e1 - ax0[t]));
          break;
        }
        case OP_PER: {  // periodic.ipynb#2: theta = [variance, period, lengthscale]
          const double s = ax0[t], c = ax1[t], il = P.cst[t][2], ip = P.cst[t][3];
          const double base = (xi[k.y * stride] - xj[k.y * stride]) * c1;
          const double q = s * il;
          o[0] += g * (v * P.ivar[t]);
          o[1] += g * (v * (4.0 * il * il) * s * c * (base * ip));
          o[2] += g * (v * ((quirks & QUIRK_PER_LENGTHSCALE) ? 1.0 : 4.0) * (q * q) * il);
          break;
        }
        case OP_LIN:    // linear.ipynb#2: theta = [variances, shifts]
          o[0] += g * (ax0[t] * ax1[t]);
          o[1] += g * (((quirks & QUIRK_LIN_SHIFT) ? 1.0 : c0) * (-(ax0[t] + ax1[t])));
          break;
        default:
          o[0] += g;
          break;
      }
    }
  }
}

}  // namespace b200gp
