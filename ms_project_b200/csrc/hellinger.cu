// BOMS Hellinger kernel-kernel distance on the device.
//
// Replaces HellingerDistanceBuilder.create_precomputed_info / hellinger_distance / compute_distance and the
// np.linalg.cholesky inside chol_safe (src/autoks/distance/distance.py:126-194,110-118,300-312):
//   precompute:  per (model, sample s)  G_s = K(theta_s) + lambda_s I over the n-point subset, logdet(G_s)
//   pairs:       per (pair, sample s)   if |ld_i - ld_j| > tol: logdet(1/2 (G_i,s + G_j,s)) else ld_i
//                                        h_s = 1 - exp(1/4 (ld_i + ld_j) - 1/2 logdet_pq);  d = sqrt(clip(mean_s h_s, 0, 1))
// Each small factorisation (n <= 128) is one CTA running the in-register Cholesky of tile_potrf.cuh.
#include "kernels.cuh"
#include "kprog.cuh"
#include "tile_potrf.cuh"

namespace b200gp {

struct HellSmem {
  ProgSmem P;
  double x[MAX_DIMS * NB];
  TilePotrfSmem T;
  double stage[NB * (NB + 1)];
};

// warp 0 sums log(piv[0..n)) in a fixed order; result valid in thread 0
__device__ __forceinline__ double logdet_from_pivots(const TilePotrfSmem& T, int n) {
  double s = 0.0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int r = 0; r < NB / 32; r++) {
      const int j = threadIdx.x + 32 * r;
      if (j < n) s += log(T.piv[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;
}

__global__ void __launch_bounds__(TP_THREADS, 1)
hellinger_precompute_kernel(const double* __restrict__ Xsub, int n, int D, const int4* __restrict__ prog,
                            const int* __restrict__ prog_off, const double* __restrict__ theta,
                            const int* __restrict__ theta_off, const double* __restrict__ lambda, int S,
                            double* __restrict__ logdet, double* __restrict__ G, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HellSmem& Sm = *reinterpret_cast<HellSmem*>(smem_raw);
  const int s = blockIdx.x, m = blockIdx.y;
  const int t0 = prog_off[m], np = theta_off[m + 1] - theta_off[m];
  load_program_at(Sm.P, prog + t0, prog_off[m + 1] - t0, theta + (size_t)theta_off[m] * S + (size_t)s * np, np, 0.0);
  for (int idx = threadIdx.x; idx < D * NB; idx += blockDim.x) {
    const int d = idx / NB, r = idx % NB;
    Sm.x[idx] = (r < n) ? Xsub[(size_t)r * D + d] : 0.0;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const double lam = lambda[s];
  double* Gs = G ? G + ((size_t)m * S + s) * n * n : nullptr;
  // interpreter pass (rolled: one copy of eval_K) into the shared staging tile, lower triangle only
#pragma unroll 1
  for (int it = 0; it < TP_R * TP_R; it++) {
    const int i = ty + 16 * (it / TP_R), c = tx + 16 * (it % TP_R);
    double v = (i == c) ? 1.0 : 0.0;
    if (i < n && c <= i) {
      v = eval_K(Sm.P, Sm.x + i, Sm.x + c, NB);
      if (i == c) v += lam;
      if (Gs) { Gs[(size_t)i * n + c] = v; Gs[(size_t)c * n + i] = v; }
    }
    Sm.stage[i * (NB + 1) + c] = v;
  }
  __syncthreads();
  double a[TP_R][TP_R];
#pragma unroll
  for (int p = 0; p < TP_R; p++)
#pragma unroll
    for (int q = 0; q < TP_R; q++) a[p][q] = Sm.stage[(ty + 16 * p) * (NB + 1) + tx + 16 * q];
  tile_potrf<false>(a, Sm.T, n);
  const double ld = logdet_from_pivots(Sm.T, n);
  if (threadIdx.x == 0) {
    const bool bad = Sm.T.fail != 0;
    logdet[(size_t)m * S + s] = bad ? nan("") : ld;
    status[(size_t)m * S + s] = bad ? 1 : 0;
  }
}

__global__ void __launch_bounds__(TP_THREADS, 1)
hellinger_pair_kernel(const double* __restrict__ logdet, const double* __restrict__ G, int S, int n,
                      const int* __restrict__ pair_i, const int* __restrict__ pair_j, double tol,
                      double* __restrict__ h, int* __restrict__ hstatus) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TilePotrfSmem& T = *reinterpret_cast<TilePotrfSmem*>(smem_raw);
  const int s = blockIdx.x, pr = blockIdx.y;
  const int mi = pair_i[pr], mj = pair_j[pr];
  const double ldi = logdet[(size_t)mi * S + s], ldj = logdet[(size_t)mj * S + s];
  double ldpq = ldi;
  int bad = 0;
  if (fabs(ldi - ldj) > tol) {        // uniform across the CTA
    const double* Gi = G + ((size_t)mi * S + s) * n * n;
    const double* Gj = G + ((size_t)mj * S + s) * n * n;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double a[TP_R][TP_R];
#pragma unroll
    for (int p = 0; p < TP_R; p++)
#pragma unroll
      for (int q = 0; q < TP_R; q++) {
        const int i = ty + 16 * p, c = tx + 16 * q;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < n && c < n) v = 0.5 * (Gi[(size_t)i * n + c] + Gj[(size_t)i * n + c]);
        a[p][q] = v;
      }
    tile_potrf<false>(a, T, n);
    ldpq = logdet_from_pivots(T, n);
    bad = T.fail != 0;
  }
  if (threadIdx.x == 0) {
    const double log_h = 0.25 * (ldi + ldj) - 0.5 * ldpq;
    h[(size_t)pr * S + s] = bad ? nan("") : 1.0 - exp(log_h);
    hstatus[(size_t)pr * S + s] = bad;
  }
}

__global__ void hellinger_reduce_kernel(const double* __restrict__ h, const int* __restrict__ hstatus, int S, int npairs,
                                        double* __restrict__ dist, int* __restrict__ status) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= npairs) return;
  double sum = 0.0;
  int bad = 0;
  for (int s = 0; s < S; s++) { sum += h[(size_t)pr * S + s]; bad |= hstatus[(size_t)pr * S + s]; }
  double d = sum / S;
  d = fmin(fmax(d, 0.0), 1.0);        // np.clip(distance, 0, 1) before the sqrt (distance.py:163-165)
  dist[pr] = bad ? nan("") : sqrt(d);
  status[pr] = bad;
}

// Frobenius / correlation distance between the vec(K_s + lambda_s I) samples of two models
// (FrobeniusDistanceBuilder / CorrelationDistanceBuilder, src/autoks/distance/distance.py:197-284).  The reference keeps
// the vectors in float32, so every entry is rounded to float32 first; the sums themselves run in FP64 (the reference's
// float32 accumulations are the larger error).  One CTA per pair, all S samples, fixed reduction order.
//   metric 0:  mean_s sqrt(sum_k (a_ks - b_ks)^2)
//   metric 1:  mean_s sqrt(1/2 (1 - clip(corr(a_s, b_s), 0, 1)))
__global__ void __launch_bounds__(256) vector_pair_kernel(const double* __restrict__ G, int S, int n, const int* __restrict__ pair_i,
                                                          const int* __restrict__ pair_j, int metric, double* __restrict__ dist) {
  __shared__ double red[8][5];
  __shared__ double acc_s;
  const int pr = blockIdx.x;
  const int mi = pair_i[pr], mj = pair_j[pr];
  const int n2 = n * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) acc_s = 0.0;
  for (int s = 0; s < S; s++) {
    const double* a = G + ((size_t)mi * S + s) * n2;
    const double* b = G + ((size_t)mj * S + s) * n2;
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};      // sum a, sum b, sum ab, sum aa, sum bb  |  v[0] = sum (a-b)^2
    for (int k = threadIdx.x; k < n2; k += blockDim.x) {
      const double x = (double)(float)a[k], y = (double)(float)b[k];
      if (metric == 0) {
        const double d = (double)((float)x - (float)y);      // the float32 difference the reference squares
        v[0] = fma(d, d, v[0]);
      } else {
        v[0] += x; v[1] += y;
        v[2] = fma(x, y, v[2]); v[3] = fma(x, x, v[3]); v[4] = fma(y, y, v[4]);
      }
    }
#pragma unroll
    for (int q = 0; q < 5; q++) {
      double t = v[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) red[warp][q] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[5];
      for (int q = 0; q < 5; q++) {
        t[q] = 0.0;
        for (int w = 0; w < 8; w++) t[q] += red[w][q];
      }
      double d;
      if (metric == 0) {
        d = sqrt(t[0]);
      } else {
        const double inv = 1.0 / n2;
        const double cab = t[2] - t[0] * t[1] * inv, caa = t[3] - t[0] * t[0] * inv, cbb = t[4] - t[1] * t[1] * inv;
        double corr = cab / (sqrt(caa) * sqrt(cbb));
        corr = fmin(fmax(corr, 0.0), 1.0);       // NaN (a constant vector) stays NaN like np.clip
        d = sqrt(0.5 * (1.0 - corr));
      }
      acc_s += d;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) dist[pr] = acc_s / S;
}

void launch_vector_pairs(const double* G, int S, int n, const int* pair_i, const int* pair_j, int npairs, int metric,
                         double* dist, cudaStream_t st) {
  vector_pair_kernel<<<npairs, 256, 0, st>>>(G, S, n, pair_i, pair_j, metric, dist);
  g_launch_count++;
}

int configure_hellinger_kernels() {
  cudaError_t e = cudaFuncSetAttribute(hellinger_precompute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(HellSmem));
  return e == cudaSuccess ? 0 : (int)e;
}

void launch_hellinger_precompute(const double* Xsub, int n, int D, const int4* prog, const int* prog_off,
                                 const double* theta, const int* theta_off, const double* lambda, int M, int S,
                                 double* logdet, double* G, int* status, cudaStream_t st) {
  dim3 grid(S, M);
  hellinger_precompute_kernel<<<grid, TP_THREADS, sizeof(HellSmem), st>>>(Xsub, n, D, prog, prog_off, theta, theta_off,
                                                                          lambda, S, logdet, G, status);
  g_launch_count++;
}

void launch_hellinger_pairs(const double* logdet, const double* G, int S, int n, const int* pair_i, const int* pair_j,
                            int npairs, double tol, double* h, int* hstatus, double* dist, int* status, cudaStream_t st) {
  const int chunk = 32768;            // grid.y limit is 65535
  for (int off = 0; off < npairs; off += chunk) {
    const int m = npairs - off < chunk ? npairs - off : chunk;
    dim3 grid(S, m);
    hellinger_pair_kernel<<<grid, TP_THREADS, sizeof(TilePotrfSmem), st>>>(logdet, G, S, n, pair_i + off, pair_j + off, tol,
                                                                          h + (size_t)off * S, hstatus + (size_t)off * S);
    g_launch_count++;
  }
  hellinger_reduce_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(h, hstatus, S, npairs, dist, status);
  g_launch_count++;
}

}  // namespace b200gp
