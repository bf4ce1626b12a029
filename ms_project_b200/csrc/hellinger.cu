// BOMS Hellinger kernel-kernel distance on the device.
//
// Replaces HellingerDistanceBuilder.create_precomputed_info / hellinger_distance / compute_distance and the
// np.linalg.cholesky inside chol_safe (src/autoks/distance/distance.py:126-194,110-118,300-312):
//   precompute:  per (model, sample s)  G_s = K(theta_s) + lambda_s I over the n-point subset, logdet(G_s)
//   pairs:       per (pair, sample s)   if |ld_i - ld_j| > tol: logdet(1/2 (G_i,s + G_j,s)) else ld_i
//                                        h_s = 1 - exp(1/4 (ld_i + ld_j) - 1/2 logdet_pq);  d = sqrt(clip(mean_s h_s, 0, 1))
// Each small factorisation (n <= 128) is one CTA running an in-register Cholesky: the precompute kernel the column-at-a-time
// one of tile_potrf.cuh, the pair kernels (where the time goes: 2.5 million factorisations in C3) the look-ahead forms of
// tile_potrf_small.cuh -- an 8 x 8 thread grid up to n = 104, a 16 x 8 grid up to 112, a 16 x 16 grid up to 128.
#include "kernels.cuh"
#include "kprog.cuh"
#include "tile_potrf.cuh"
#include "tile_potrf_small.cuh"

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
  tile_potrf_pivots(a, Sm.T, n);
  const double ld = logdet_from_pivots(Sm.T, n);
  if (threadIdx.x == 0) {
    const bool bad = Sm.T.fail != 0;
    logdet[(size_t)m * S + s] = bad ? nan("") : ld;
    status[(size_t)m * S + s] = bad ? 1 : 0;
  }
}

// One CTA per (pair, sample): pivots of chol(1/2 (G_i,s + G_j,s)) by the two-columns-per-barrier register factorisation of
// tile_potrf_small.cuh on a 16 x 16 thread grid (the form that serves 112 < n <= 128; the look-ahead kernels below serve the
// smaller orders).  Only the lower triangle of the two Gram matrices is read; R block rows of 16 cover the order.
// blockIdx.y = sample (all pair kernels): every CTA in flight works on the SAME sample, whose M Gram matrices (C3: 500 x
// 80 KB, half of it touched) stay resident in the 126 MB L2 while all pairs of that sample are processed.
template <int R, int MINB, int COLS>
__global__ void __launch_bounds__(TS_THREADS, MINB)
hellinger_pair_kernel(const double* __restrict__ logdet, const double* __restrict__ G, int S, int n,
                      const int* __restrict__ pair_i, const int* __restrict__ pair_j, double tol,
                      double* __restrict__ h, int* __restrict__ hstatus) {
  __shared__ SmallPotrfSmem<R> T;
  const int pr = blockIdx.x, s = blockIdx.y;
  const int mi = pair_i[pr], mj = pair_j[pr];
  const double ldi = logdet[(size_t)mi * S + s], ldj = logdet[(size_t)mj * S + s];
  double ldpq = ldi;
  int bad = 0;
  if (fabs(ldi - ldj) > tol) {        // uniform across the CTA
    const double* Gi = G + ((size_t)mi * S + s) * n * n;
    const double* Gj = G + ((size_t)mj * S + s) * n * n;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double a[R * (R + 1) / 2];
#pragma unroll
    for (int p = 0; p < R; p++)
#pragma unroll
      for (int q = 0; q <= p; q++) {
        const int i = ty + 16 * p, c = tx + 16 * q;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < n && c <= i) v = 0.5 * (__ldg(Gi + (size_t)i * n + c) + __ldg(Gj + (size_t)i * n + c));
        a[ts_idx(p, q)] = v;
      }
    small_potrf_pivots<R, COLS>(a, T, n);
    double sum = 0.0;                 // warp 0 sums log(piv[0..n)) in a fixed order
    if (threadIdx.x < 32) {
#pragma unroll
      for (int r = 0; r < (16 * R + 31) / 32; r++) {
        const int j = threadIdx.x + 32 * r;
        if (j < n) sum += log(T.piv[j]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    ldpq = sum;
    bad = T.fail != 0;
  }
  if (threadIdx.x == 0) {
    const double log_h = 0.25 * (ldi + ldj) - 0.5 * ldpq;
    h[(size_t)pr * S + s] = bad ? nan("") : 1.0 - exp(log_h);
    hstatus[(size_t)pr * S + s] = bad;
  }
}

// The same work item on the look-ahead factorisation (rect_potrf_pivots_la: next column and its pivot reciprocal published
// under the trailing update, one reciprocal per step and matrix).
template <int RP, int MINB>
__global__ void __launch_bounds__(TR_THREADS, MINB)
hellinger_pair_la_kernel(const double* __restrict__ logdet, const double* __restrict__ G, int S, int n,
                         const int* __restrict__ pair_i, const int* __restrict__ pair_j, double tol,
                         double* __restrict__ h, int* __restrict__ hstatus) {
  __shared__ RectLaSmem<RP> T;
  const int pr = blockIdx.x, s = blockIdx.y;
  const int mi = pair_i[pr], mj = pair_j[pr];
  const double ldi = logdet[(size_t)mi * S + s], ldj = logdet[(size_t)mj * S + s];
  double ldpq = ldi;
  int bad = 0;
  if (fabs(ldi - ldj) > tol) {        // uniform across the CTA
    const double* Gi = G + ((size_t)mi * S + s) * n * n;
    const double* Gj = G + ((size_t)mj * S + s) * n * n;
    int ty, tx;
    rect_la_coords(ty, tx);
    double a[RP * (RP + 1)];
#pragma unroll
    for (int p = 0; p < RP; p++)
#pragma unroll
      for (int q = 0; q <= 2 * p + 1; q++) {
        const int i = ty + 16 * p, c = tx + 8 * q;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < n && c <= i) v = 0.5 * (__ldg(Gi + (size_t)i * n + c) + __ldg(Gj + (size_t)i * n + c));
        a[tr_idx(p, q)] = v;
      }
    rect_potrf_pivots_la<RP>(a, T, n);
    double sum = 0.0;                 // warp 0 sums log(piv[0..n)) in a fixed order and checks the pivots
    if (threadIdx.x < 32) {
      int badp = 0;
#pragma unroll
      for (int r = 0; r < (16 * RP + 31) / 32; r++) {
        const int j = threadIdx.x + 32 * r;
        if (j < n) {
          const double pv = T.piv[j];
          sum += log(pv);
          badp |= (!(pv > 0.0) || !(pv < 1.7e308)) ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      bad = __any_sync(0xffffffffu, badp);
    }
    ldpq = sum;
  }
  if (threadIdx.x == 0) {
    const double log_h = 0.25 * (ldi + ldj) - 0.5 * ldpq;
    h[(size_t)pr * S + s] = bad ? nan("") : 1.0 - exp(log_h);
    hstatus[(size_t)pr * S + s] = bad;
  }
}

// The same work item on the 8 x 8 thread grid (64 threads per matrix, square_potrf_pivots_la), n <= 8 NB8.
template <int NB8, int MINB>
__global__ void __launch_bounds__(TQ_THREADS, MINB)
hellinger_pair_sq_kernel(const double* __restrict__ logdet, const double* __restrict__ G, int S, int n,
                         const int* __restrict__ pair_i, const int* __restrict__ pair_j, double tol,
                         double* __restrict__ h, int* __restrict__ hstatus) {
  __shared__ SquareLaSmem<NB8> T;
  const int pr = blockIdx.x, s = blockIdx.y;
  const int mi = pair_i[pr], mj = pair_j[pr];
  const double ldi = logdet[(size_t)mi * S + s], ldj = logdet[(size_t)mj * S + s];
  double ldpq = ldi;
  int bad = 0;
  if (fabs(ldi - ldj) > tol) {        // uniform across the CTA
    const double* Gi = G + ((size_t)mi * S + s) * n * n;
    const double* Gj = G + ((size_t)mj * S + s) * n * n;
    int ty, tx;
    square_la_coords(ty, tx);
    // (Measured on C3 and not kept: staging G_j through cp.async into per-thread shared-memory columns so that both
    // operands are in flight at once -- 0.119 s against 0.102 s; L1 prefetches of both operands in front of the loads --
    // 0.112 s.  The load phase is 25 % of a warp's life in ncu, but the other three CTAs of the SM compute meanwhile.)
    double a[NB8 * (NB8 + 1) / 2];
#pragma unroll
    for (int p = 0; p < NB8; p++)
#pragma unroll
      for (int q = 0; q <= p; q++) {
        const int i = ty + 8 * p, c = tx + 8 * q;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < n && c <= i) v = 0.5 * (__ldg(Gi + (size_t)i * n + c) + __ldg(Gj + (size_t)i * n + c));
        a[tq_idx(p, q)] = v;
      }
    square_potrf_pivots_la<NB8>(a, T, n);
    double sum = 0.0;                 // warp 0 sums log(piv[0..n)) in a fixed order and checks the pivots
    if (threadIdx.x < 32) {
      int badp = 0;
#pragma unroll
      for (int r = 0; r < (8 * NB8 + 31) / 32; r++) {
        const int j = threadIdx.x + 32 * r;
        if (j < n) {
          const double pv = T.piv[j];
          sum += log(pv);
          badp |= (!(pv > 0.0) || !(pv < 1.7e308)) ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      bad = __any_sync(0xffffffffu, badp);
    }
    ldpq = sum;
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


// ------------------------------------------------------------------------------------------------
// Subsets of more than 128 points (the reference hands the builders the WHOLE training set,
// src/autoks/core/model_selection/smbo_model_selector.py:91-97): the Gram matrices are factorised by the blocked DMMA
// Cholesky of the scoring path over the context's workspace slots instead of the one-CTA register factorisation.
// ------------------------------------------------------------------------------------------------
// logdet[item] = sum of the per-diagonal-block partials of the slot's factorisation
__global__ void slot_logdet_kernel(Batch b, double* __restrict__ logdet, int* __restrict__ status) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= b.nslots) return;
  const int item = b.ids[slot];
  const bool bad = b.slot_status[slot] != 0;
  double s = 0.0;
  for (int k = 0; k < b.T; k++) s += b.logdet_part[slot * b.T + k];
  logdet[item] = bad ? nan("") : s;
  status[item] = bad ? 1 : 0;
}

void launch_slot_logdet(const Batch& b, double* logdet, int* status, cudaStream_t st) {
  slot_logdet_kernel<<<(b.nslots + 127) / 128, 128, 0, st>>>(b, logdet, status);
  g_launch_count++;
}

// M1[slot] (lower tiles) <- 1/2 (G_i,s + G_j,s) for work item w = w0 + slot: pair work_pair[w], sample work_s[w];
// identity on the padding.  One CTA per 128 x 128 lower tile.
__global__ void __launch_bounds__(256) havg_build_kernel(Batch b, const double* __restrict__ G, int S, int n,
                                                         const int* __restrict__ pair_i, const int* __restrict__ pair_j,
                                                         const int* __restrict__ work_pair, const int* __restrict__ work_s) {
  const int slot = blockIdx.y;
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  const int pr = work_pair[slot], s = work_s[slot];
  const double* Gi = G + ((size_t)pair_i[pr] * S + s) * n * n;
  const double* Gj = G + ((size_t)pair_j[pr] * S + s) * n * n;
  double* out = b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB;
  const int c = threadIdx.x & (NB - 1), gj = tj * NB + c;
  for (int r = threadIdx.x >> 7; r < NB; r += 2) {
    const int gi = ti * NB + r;
    double v = (gi == gj) ? 1.0 : 0.0;
    if (gi < n && gj < n) v = 0.5 * (Gi[(size_t)gi * n + gj] + Gj[(size_t)gi * n + gj]);
    out[(size_t)r * b.Np + c] = v;
  }
}

void launch_havg_build(const Batch& b, const double* G, int S, int n, const int* pair_i, const int* pair_j,
                       const int* work_pair, const int* work_s, cudaStream_t st) {
  dim3 grid(tri_count(b.T), b.nslots);
  havg_build_kernel<<<grid, 256, 0, st>>>(b, G, S, n, pair_i, pair_j, work_pair, work_s);
  g_launch_count++;
}

// ldpq[pair][s] <- ld_i[s] (the value hellinger_distance keeps for samples it does not refactorise), hstatus <- 0
__global__ void hpair_init_kernel(const double* __restrict__ logdet, int S, const int* __restrict__ pair_i, size_t total,
                                  double* __restrict__ ldpq, int* __restrict__ hstatus) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const size_t pr = k / S;
    ldpq[k] = logdet[(size_t)pair_i[pr] * S + (k % S)];
    hstatus[k] = 0;
  }
}

// scatter the chunk's log-determinants to their (pair, sample) cells
__global__ void hpair_scatter_kernel(Batch b, int S, const int* __restrict__ work_pair, const int* __restrict__ work_s,
                                     double* __restrict__ ldpq, int* __restrict__ hstatus) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= b.nslots) return;
  const bool bad = b.slot_status[slot] != 0;
  double s = 0.0;
  for (int k = 0; k < b.T; k++) s += b.logdet_part[slot * b.T + k];
  const size_t cell = (size_t)work_pair[slot] * S + work_s[slot];
  ldpq[cell] = bad ? nan("") : s;
  hstatus[cell] = bad ? 1 : 0;
}

// h[pair][s] = 1 - exp(1/4 (ld_i + ld_j) - 1/2 ldpq), in place over ldpq
__global__ void hpair_combine_kernel(const double* __restrict__ logdet, int S, const int* __restrict__ pair_i,
                                     const int* __restrict__ pair_j, size_t total, double* __restrict__ ldpq_h,
                                     const int* __restrict__ hstatus) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const size_t pr = k / S;
    const int s = (int)(k % S);
    const double ldi = logdet[(size_t)pair_i[pr] * S + s], ldj = logdet[(size_t)pair_j[pr] * S + s];
    ldpq_h[k] = hstatus[k] ? nan("") : 1.0 - exp(0.25 * (ldi + ldj) - 0.5 * ldpq_h[k]);
  }
}

void launch_hpair_init(const double* logdet, int S, const int* pair_i, int npairs, double* ldpq, int* hstatus, cudaStream_t st) {
  hpair_init_kernel<<<592, 256, 0, st>>>(logdet, S, pair_i, (size_t)npairs * S, ldpq, hstatus);
  g_launch_count++;
}

void launch_hpair_scatter(const Batch& b, int S, const int* work_pair, const int* work_s, double* ldpq, int* hstatus,
                          cudaStream_t st) {
  hpair_scatter_kernel<<<(b.nslots + 127) / 128, 128, 0, st>>>(b, S, work_pair, work_s, ldpq, hstatus);
  g_launch_count++;
}

void launch_hpair_finish(const double* logdet, int S, const int* pair_i, const int* pair_j, int npairs, double* ldpq_h,
                         int* hstatus, double* dist, int* status, cudaStream_t st) {
  hpair_combine_kernel<<<592, 256, 0, st>>>(logdet, S, pair_i, pair_j, (size_t)npairs * S, ldpq_h, hstatus);
  hellinger_reduce_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(ldpq_h, hstatus, S, npairs, dist, status);
  g_launch_count += 2;
}

int configure_hellinger_kernels() {
  cudaError_t e = cudaFuncSetAttribute(hellinger_precompute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(HellSmem));
  if (e != cudaSuccess) return (int)e;
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
  dim3 grid(npairs, S);               // x = pair (fastest), y = sample: sample-major execution order, see the kernel
#define HP_LAUNCH(R, MINB, COLS) hellinger_pair_kernel<R, MINB, COLS><<<grid, TS_THREADS, 0, st>>>(logdet, G, S, n, pair_i, pair_j, tol, h, hstatus)
  // Measured on C3 (124 750 pairs x 20 samples, n = 100; round 1's kernel: 0.369 s):
  //   16 x 16 thread grid (256 threads): one column per barrier, three CTAs per SM 0.224 s; two columns per barrier 0.245 s
  //     (three CTAs, 80 B of spills) / 0.273 s (two CTAs); pivot reciprocal by its owner in front of the barrier 0.25-0.29 s
  //   16 x 8 thread grid (128 threads, 167 registers, three CTAs per SM) 0.185 s
  //     + look-ahead (next column and its pivot reciprocal published under the trailing update, scaled row operands) 0.167 s
  //     + static buffer parity, pivot test after the factorisation (91 -> 73 instructions per step) 0.148 s
  //     + lane map for one-wavefront shared-memory accesses (ncu: the shared-memory pipe was 90 % busy) 0.130 s
  //   8 x 8 thread grid (64 threads, 240 registers, four CTAs per SM), same look-ahead and lane rules 0.102 s (kept)
  //     with G_j staged by cp.async 0.119 s, with L1 prefetches of both operands 0.112 s (not kept)
  //   DMMA formulation (8 x 8 accumulator tiles over 4 or 8 warps, 13 panel steps): correct, 0.218-0.279 s, 47 % of its
  //     stall samples behind the three barriers of a panel step; removed (profiles/experiments/)
#define HL_LAUNCH(RP, MINB) hellinger_pair_la_kernel<RP, MINB><<<grid, TR_THREADS, 0, st>>>(logdet, G, S, n, pair_i, pair_j, tol, h, hstatus)
#define HQ_LAUNCH(NB8, MINB) hellinger_pair_sq_kernel<NB8, MINB><<<grid, TQ_THREADS, 0, st>>>(logdet, G, S, n, pair_i, pair_j, tol, h, hstatus)
  if (n <= 64) HQ_LAUNCH(8, 8);
  else if (n <= 80) HQ_LAUNCH(10, 6);
  else if (n <= 96) HQ_LAUNCH(12, 4);
  else if (n <= 104) HQ_LAUNCH(13, 4);
  else if (n <= 112) HL_LAUNCH(7, 3);
  else HP_LAUNCH(8, 2, 2);
#undef HL_LAUNCH
#undef HQ_LAUNCH
#undef HP_LAUNCH
  g_launch_count++;
  hellinger_reduce_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(h, hstatus, S, npairs, dist, status);
  g_launch_count++;
}

}  // namespace b200gp
