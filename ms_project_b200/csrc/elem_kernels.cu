// Element-wise phases of the scoring path: fused kernel-matrix build, fused gradient pass, finalize.
//
// K-build replaces `kern.K(X)` + `diag.add(Ky, variance + 1e-8)` (GPy ExactGaussianInference; reference
// call sites src/autoks/backend/kernel.py:317-322, src/autoks/distance/distance.py:186-187).
// The gradient pass replaces `kern.update_gradients_full(dL_dK, X)` and the Gaussian likelihood's
// `exact_inference_gradients`: dL/dtheta_p = sum_ij 1/2 (alpha_i alpha_j - K^-1_ij) dK_ij/dtheta_p with
// dK regenerated from X in registers, so neither dL/dK nor dK ever exists in HBM.
#include "kernels.cuh"
#include "kprog.cuh"

namespace b200gp {

constexpr double EXACT_INFERENCE_JITTER = 1e-8;   // GPy exact_gaussian_inference.py: diag.add(Ky, variance + 1e-8)

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

// ------------------------------------------------------------------------------------------------
// K-build: one CTA per lower tile (ti >= tj) per slot; writes the full 128x128 tile of
// K + (noise + 1e-8 + jitter) I, identity in the padding.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_k_kernel(Batch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemSmem& S = *reinterpret_cast<ElemSmem*>(smem_raw);
  const int slot = blockIdx.y, item = b.ids[slot];
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  load_program(S.P, b.prog, b.prog_off, b.theta, b.theta_off, item);
  stage_inputs(S, b, ti, tj);
  __syncthreads();
  const double diag_add = b.add_diag ? S.P.noise + EXACT_INFERENCE_JITTER + (b.jitter ? b.jitter[slot] : 0.0) : 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  if (ti == tj) {   // deterministic partial of trace(K_y) for the jitter ladder
    S.red[threadIdx.x] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int t = 0; t < 256; t++) s += S.red[t];
      b.diag_part[slot * b.T + ti] = s;
    }
  }
}

void launch_build_k(const Batch& b, cudaStream_t s) {
  dim3 grid(tri_count(b.T), b.nslots);
  build_k_kernel<<<grid, 256, sizeof(ElemSmem), s>>>(b);
  g_launch_count++;
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
// Gradient pass: one CTA per lower tile per slot.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_kernel(Batch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemSmem& S = *reinterpret_cast<ElemSmem*>(smem_raw);
  double* wred = reinterpret_cast<double*>(smem_raw + sizeof(ElemSmem));   // [8][pstride]
  const int slot = blockIdx.y, item = b.ids[slot];
  if (b.slot_status[slot] != 0) return;
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  load_program(S.P, b.prog, b.prog_off, b.theta, b.theta_off, item);
  stage_inputs(S, b, ti, tj);
  if (threadIdx.x < NB) {
    S.ai[threadIdx.x] = b.alpha[(size_t)slot * b.Np + ti * NB + threadIdx.x];
    S.aj[threadIdx.x] = b.alpha[(size_t)slot * b.Np + tj * NB + threadIdx.x];
  }
  __syncthreads();
  const int np = S.P.nparam;
  double acc[MAX_PARAMS + 1];
  for (int p = 0; p <= np; p++) acc[p] = 0.0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* Kinv = b.M1 + slot * b.mstride + (size_t)(ti * NB) * b.Np + tj * NB;
  double kv_next = Kinv[(size_t)warp * b.Np + lane];
#pragma unroll 1
  for (int it = 0; it < 64; it++) {
    const int r = warp + 8 * (it >> 2), c = lane + 32 * (it & 3);
    const double kv = kv_next;
    if (it + 1 < 64) kv_next = Kinv[(size_t)(warp + 8 * ((it + 1) >> 2)) * b.Np + lane + 32 * ((it + 1) & 3)];
    const int gi = ti * NB + r, gj = tj * NB + c;
    if (gi >= b.N || gj >= b.N || gj > gi) continue;
    double w = 0.5 * (S.ai[r] * S.aj[c] - kv);
    if (gi == gj) acc[np] += w;     // dL/d(noise) = trace(dL/dK)
    else w *= 2.0;                  // symmetric pair (i, j) and (j, i)
    eval_grad(S.P, S.xi + r, S.xj + c, NB, w, acc, b.quirks);
  }
  for (int p = 0; p <= np; p++) {
    double v = acc[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wred[warp * b.pstride + p] = v;
  }
  __syncthreads();
  double* out = b.grad_part + ((size_t)slot * gridDim.x + blockIdx.x) * b.pstride;
  for (int p = threadIdx.x; p <= np; p += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) s += wred[w * b.pstride + p];
    out[p] = s;
  }
}

void launch_grad(const Batch& b, cudaStream_t s) {
  dim3 grid(tri_count(b.T), b.nslots);
  const size_t smem = sizeof(ElemSmem) + 8 * (size_t)b.pstride * sizeof(double);
  grad_kernel<<<grid, 256, smem, s>>>(b);
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
    const int ntiles = tri_count(b.T);
    const double* gp = b.grad_part + (size_t)slot * ntiles * b.pstride;
    for (int p = threadIdx.x; p <= np; p += blockDim.x) {
      double g = 0.0;
      for (int t = 0; t < ntiles; t++) g += gp[(size_t)t * b.pstride + p];
      b.dlml[th0 + p] = g;
      if (!(fabs(g) < 1.7e308)) bad = 1;
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
  e = cudaFuncSetAttribute(build_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ElemSmem));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(ElemSmem) + 8 * (MAX_PARAMS + 1) * sizeof(double)));
  if (e != cudaSuccess) return (int)e;
  return configure_chol_kernels();
}

}  // namespace b200gp
