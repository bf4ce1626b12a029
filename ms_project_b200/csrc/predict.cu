// Posterior prediction at test inputs: GP.predict as ModelSelector.predict / evaluate use it
// (src/autoks/core/model_selection/base.py:231-251) and the BOMS surrogate's predict / get_f_max
// (src/autoks/gp_regression_models.py:113-140).  GPy semantics (Posterior._raw_predict + Gaussian.predictive_values):
//     mean_s = k(x*_s, X) alpha
//     var_s  = clip(k(x*_s, x*_s) - |L^-1 k(X, x*_s)|^2, 1e-15, inf)  (+ noise variance with include_likelihood)
// The factorisation state (V = L^-T in M2, alpha) is the one the scoring pipeline leaves in slot 0.
#include "kernels.cuh"
#include "kprog.cuh"

namespace b200gp {

constexpr int PB = 32;     // test points per predict CTA

struct CrossSmem {
  ProgSmem P;
  double xi[MAX_DIMS * NB];
  double xj[MAX_DIMS * NB];
};

// Ks[pt][r] = k(x*_pt, x_r) for a chunk of test points (rows) against all training points (columns, ld = Np, zero
// in the padding); kss[pt] = k(x*_pt, x*_pt).  One CTA per 128 x 128 block; generic postfix interpreter.
__global__ void __launch_bounds__(256) cross_k_kernel(Batch b, int item, const double* __restrict__ Xs, int ns, double* __restrict__ Ks,
                                                      double* __restrict__ kss) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CrossSmem& S = *reinterpret_cast<CrossSmem*>(smem_raw);
  const int tp = blockIdx.y, tc = blockIdx.x;
  load_program(S.P, b.prog, b.prog_off, b.theta, b.theta_off, item, b.lut, b.lut_n);
  for (int idx = threadIdx.x; idx < b.D * NB; idx += blockDim.x) {
    const int d = idx / NB, r = idx % NB;
    const int pt = tp * NB + r;
    S.xi[idx] = (pt < ns) ? Xs[(size_t)pt * b.D + d] : 0.0;
    S.xj[idx] = b.Xt[(size_t)d * b.Np + tc * NB + r];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int it = 0; it < 64; it++) {
    const int r = warp + 8 * (it >> 2), c = lane + 32 * (it & 3);
    const int pt = tp * NB + r, gj = tc * NB + c;
    if (pt >= ns) continue;
    Ks[(size_t)pt * b.Np + gj] = (gj < b.N) ? eval_K(S.P, S.xi + r, S.xj + c, NB) : 0.0;
  }
  if (tc == 0 && threadIdx.x < NB) {
    const int pt = tp * NB + threadIdx.x;
    if (pt < ns) kss[pt] = eval_K(S.P, S.xi + threadIdx.x, S.xi + threadIdx.x, NB);
  }
}

// t = Ks V restricted to one 128-column tile of V (upper triangular: rows r <= c), for PB test points:
// part[ct][pt] = sum_{c in tile} t_c^2; the ct = 0 CTAs also form mean[pt] = Ks[pt] . alpha.
__global__ void __launch_bounds__(256) predict_kernel(Batch b, const double* __restrict__ Ks, int ns, double* __restrict__ part,
                                                      double* __restrict__ mean) {
  __shared__ double ks[PB][NB + 1];
  __shared__ double red[2][NB];
  const int ct = blockIdx.x, p0 = blockIdx.y * PB;
  const int c = threadIdx.x & (NB - 1), half = threadIdx.x >> 7;      // half: test points [16 half, 16 half + 16)
  const double* V = b.M2 + (size_t)ct * NB + c;                        // slot 0
  double acc[PB / 2];
#pragma unroll
  for (int p = 0; p < PB / 2; p++) acc[p] = 0.0;
  for (int rt = 0; rt <= ct; rt++) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < PB * NB; idx += blockDim.x) {
      const int p = idx / NB, r = idx % NB;
      ks[p][r] = (p0 + p < ns) ? Ks[(size_t)(p0 + p) * b.Np + rt * NB + r] : 0.0;
    }
    __syncthreads();
    for (int r = 0; r < NB; r++) {
      const double v = V[(size_t)(rt * NB + r) * b.Np];
#pragma unroll
      for (int p = 0; p < PB / 2; p++) acc[p] = fma(ks[half * (PB / 2) + p][r], v, acc[p]);
    }
  }
  // sum of squares over the tile's 128 columns, fixed order
  for (int p = 0; p < PB / 2; p++) {
    __syncthreads();
    red[half][c] = acc[p] * acc[p];
    __syncthreads();
    if (c == 0) {
      double s = 0.0;
      for (int q = 0; q < NB; q++) s += red[half][q];
      const int pt = p0 + half * (PB / 2) + p;
      if (pt < ns) part[(size_t)ct * ns + pt] = s;
    }
  }
  if (ct == 0 && threadIdx.x < PB && p0 + threadIdx.x < ns) {
    const double* row = Ks + (size_t)(p0 + threadIdx.x) * b.Np;
    double s = 0.0;
    for (int r = 0; r < b.N; r++) s = fma(row[r], b.alpha[r], s);
    mean[p0 + threadIdx.x] = s;
  }
}

__global__ void predict_finish_kernel(Batch b, int item, const double* __restrict__ part, const double* __restrict__ kss, int ns,
                                      int include_likelihood, double* __restrict__ var) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= ns) return;
  double s = 0.0;
  for (int ct = 0; ct < b.T; ct++) s += part[(size_t)ct * ns + pt];
  double v = kss[pt] - s;
  v = (v < 1e-15) ? 1e-15 : v;                       // GPy: np.clip(var, 1e-15, inf); NaN stays NaN
  if (include_likelihood) v += b.theta[b.theta_off[item + 1] - 1];
  var[pt] = v;
}

// Ks lives in M1 of slot 0 (L is no longer needed once V and alpha exist): up to Np test points per call.
void launch_predict(const Batch& b, int item, const double* Xs, int ns, int include_likelihood, double* scratch, double* mean,
                    double* var, cudaStream_t s) {
  double* Ks = b.M1;
  double* kss = scratch;                    // [ns]
  double* part = scratch + ns;              // [T][ns]
  const int tp = (ns + NB - 1) / NB;
  cross_k_kernel<<<dim3(b.T, tp), 256, sizeof(CrossSmem), s>>>(b, item, Xs, ns, Ks, kss);
  predict_kernel<<<dim3(b.T, (ns + PB - 1) / PB), 256, 0, s>>>(b, Ks, ns, part, mean);
  predict_finish_kernel<<<(ns + 255) / 256, 256, 0, s>>>(b, item, part, kss, ns, include_likelihood, var);
  g_launch_count += 3;
}

int configure_predict_kernels() {
  return (int)cudaFuncSetAttribute(cross_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CrossSmem));
}

}  // namespace b200gp
