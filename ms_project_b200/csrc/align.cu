// Pairwise centred kernel alignment of a population of candidate kernels:
// pairwise_centered_alignments / centered_alignment / center_kernel / alignment / inner_frob,
// src/autoks/core/covariance.py:278-338, fired on every generation by the statistics callbacks.
//
// The reference centres each N x N kernel matrix with two N^3 matmuls per PAIR (O(M^2 N^3)).  With H = I - 11^T/N and
// symmetric K_a, K_b:  <H K_a H, H K_b H>_F = <K_a, K_b>_F - (2/N) rs_a . rs_b + tot_a tot_b / N^2
// (rs = row sums, tot = sum of all entries), so all pairs need is the Gram matrix of the vectorised kernel matrices --
// one M x M x N^2 contraction, run on the DMMA pipe straight off the lower tiles the K-build leaves in M1 -- plus
// the row sums.  Everything is reduced in a fixed order (split-K partials summed sequentially).
#include "kernels.cuh"
#include "tile_gemm.cuh"

namespace b200gp {

constexpr int AL_ROWS = 16;       // matrix rows per split-K CTA (all inside one 128-row tile row)

// rs[slot][i] = sum_j K[i][j] over the real N columns, from the lower tiles (tiles above the diagonal by symmetry).
__global__ void __launch_bounds__(256) align_rowsum_kernel(Batch b, double* __restrict__ rs) {
  __shared__ double tile[32][NB + 1];
  __shared__ double part[2][NB];
  const int slot = blockIdx.y, ti = blockIdx.x;
  const double* K = b.M1 + slot * b.mstride;
  const int c = threadIdx.x & (NB - 1), half = threadIdx.x >> 7;
  double sum = 0.0;                   // thread (c, half): row ti*128 + c, half of the column tiles / rows
  for (int tj = 0; tj <= ti; tj++) {  // stored tile (ti, tj): row sums, through shared memory for coalescing
    for (int h = 0; h < 4; h++) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < 32 * NB; idx += blockDim.x) {
        const int r = idx / NB, cc = idx % NB;
        tile[r][cc] = K[(size_t)(ti * NB + h * 32 + r) * b.Np + tj * NB + cc];
      }
      __syncthreads();
      if ((c >> 5) == h) {            // this thread's row is in the staged quarter
        const int r = c & 31;
        for (int cc = half * 64; cc < half * 64 + 64; cc++)
          if (tj * NB + cc < b.N) sum += tile[r][cc];
      }
    }
  }
  for (int tj = ti + 1; tj < b.T; tj++) {   // tile (ti, tj) = stored tile (tj, ti) transposed: column sums, coalesced
    for (int r = half * 64; r < half * 64 + 64; r++)
      if (tj * NB + r < b.N) sum += K[(size_t)(tj * NB + r) * b.Np + ti * NB + c];
  }
  part[half][c] = sum;
  __syncthreads();
  if (half == 0) rs[(size_t)slot * b.Np + ti * NB + c] = (ti * NB + c < b.N) ? part[0][c] + part[1][c] : 0.0;
}

// Split-K Gram: partial[split][a][b] = sum over the matrix rows [16 split, 16 split + 16) of
//   2 * sum_{j < tile_row*128} K_a[i][j] K_b[i][j]  +  sum_{j in the diagonal tile} K_a[i][j] K_b[i][j]
// for a 128 x 64 block of model pairs; operands are the models' M1 matrices viewed as rows of length Np^2.
__global__ void __launch_bounds__(TG_THREADS, 2) align_gram_kernel(Batch b, int nmodels, int mp, double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int nbt = mp / TG_TN;                           // 64-wide column blocks
  const int ta = blockIdx.x / nbt, tb = blockIdx.x % nbt;
  if (tb * TG_TN > ta * TG_TM + TG_TM - 1) return;      // block strictly above the diagonal of the pair matrix
  const int split = blockIdx.y;
  const int r0 = split * AL_ROWS;
  const int nrows = (b.N - r0 < AL_ROWS) ? b.N - r0 : AL_ROWS;
  const int tile_row = r0 / NB;
  const size_t ld = b.mstride;
  const double* M = b.M1;
  const int a0 = ta * TG_TM, b0 = tb * TG_TN, last = nmodels - 1;
  const size_t np = b.Np;

  // cp.async copies of chunk `q` of a segment family: row r0 + q / per_row, 16 doubles at col0 + 16 (q % per_row)
  auto loader = [&](int per_row, int col0) {
    return [=](double* stage, int q) {
      const size_t off = (size_t)(r0 + q / per_row) * np + col0 + (size_t)(q % per_row) * TG_KC;
      double* As = stage;
      double* Bs = stage + TG_TM * TG_LDK;
#pragma unroll
      for (int r = 0; r < TG_TM * 8 / TG_THREADS; r++) {
        const int seg = threadIdx.x + TG_THREADS * r;
        const int row = seg >> 3, part = (seg & 7) * 2;
        const int m = (a0 + row < last) ? a0 + row : last;       // padding models re-read the last one (ignored later)
        cp_async16(As + row * TG_LDK + part, M + (size_t)m * ld + off + part);
      }
#pragma unroll
      for (int r = 0; r < TG_TN * 8 / TG_THREADS; r++) {
        const int seg = threadIdx.x + TG_THREADS * r;
        const int row = seg >> 3, part = (seg & 7) * 2;
        const int m = (b0 + row < last) ? b0 + row : last;
        cp_async16(Bs + row * TG_LDK + part, M + (size_t)m * ld + off + part);
      }
    };
  };

  double acc[TG_MT][TG_NT][2];
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
  if (nrows > 0) {
    const int per_row = tile_row * (NB / TG_KC);
    if (per_row > 0) tile_gemm_pipeline(acc, loader(per_row, 0), nrows * per_row, smem);
#pragma unroll
    for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) { acc[mt][nt][0] *= 2.0; acc[mt][nt][1] *= 2.0; }
    tile_gemm_pipeline(acc, loader(NB / TG_KC, tile_row * NB), nrows * (NB / TG_KC), smem);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  int wm, wn;
  warp_coords(warp, wm, wn);
  double* C = partial + (size_t)split * mp * mp + (size_t)(a0 + wm * 32 + g) * mp + b0 + wn * 32 + tg * 2;
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
    for (int nt = 0; nt < TG_NT; nt++) {
      double2 v;
      v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
      *reinterpret_cast<double2*>(C + (size_t)(mt * 8) * mp + nt * 8) = v;
    }
}

// centred inner products and alignments: out[a][b] = C_ab / sqrt(C_aa C_bb)
__global__ void __launch_bounds__(256) align_centered_kernel(Batch b, int nmodels, int mp, int nsplit, const double* __restrict__ partial,
                                                             const double* __restrict__ rs, double* __restrict__ cmat) {
  __shared__ double red[256];
  const int a = blockIdx.x, bb = blockIdx.y;
  if (bb > a) return;
  double s = 0.0;                    // rs_a . rs_b
  for (int i = threadIdx.x; i < b.N; i += blockDim.x) s = fma(rs[(size_t)a * b.Np + i], rs[(size_t)bb * b.Np + i], s);
  double ta = 0.0, tb = 0.0;
  for (int i = threadIdx.x; i < b.N; i += blockDim.x) { ta += rs[(size_t)a * b.Np + i]; tb += rs[(size_t)bb * b.Np + i]; }
  double vals[3] = {s, ta, tb};
  double out3[3];
  for (int q = 0; q < 3; q++) {
    __syncthreads();
    red[threadIdx.x] = vals[q];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    out3[q] = red[0];
  }
  if (threadIdx.x == 0) {
    double sab = 0.0;
    for (int k = 0; k < nsplit; k++) sab += partial[(size_t)k * mp * mp + (size_t)a * mp + bb];
    const double n = (double)b.N;
    const double c = sab - 2.0 / n * out3[0] + out3[1] * out3[2] / (n * n);
    cmat[(size_t)a * nmodels + bb] = c;
    cmat[(size_t)bb * nmodels + a] = c;
  }
}

__global__ void align_normalise_kernel(int nmodels, const double* __restrict__ cmat, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nmodels * nmodels) return;
  const int a = idx / nmodels, bb = idx % nmodels;
  out[idx] = cmat[idx] / sqrt(cmat[(size_t)a * nmodels + a] * cmat[(size_t)bb * nmodels + bb]);
}

// K of the first `nmodels` slots must already be in M1 (launch_build_k with add_diag = 0).  Scratch: rs in b.z,
// split-K partials in M2, the centred inner products in b.Dinv.
int launch_alignments(const Batch& b, int nmodels, double* out, cudaStream_t s) {
  const int mp = round_up(nmodels, TG_TM);
  const int nsplit = (b.N + AL_ROWS - 1) / AL_ROWS;
  if ((size_t)nsplit * mp * mp > (size_t)b.nslots * b.mstride) return -1;           // partials do not fit in M2
  if ((size_t)nmodels * nmodels > (size_t)b.nslots * b.T * NB * NB) return -2;      // centred products do not fit in Dinv
  double* rs = b.z;
  double* partial = b.M2;
  double* cmat = b.Dinv;
  align_rowsum_kernel<<<dim3(b.T, nmodels), 256, 0, s>>>(b, rs);
  align_gram_kernel<<<dim3((mp / TG_TM) * (mp / TG_TN), nsplit), TG_THREADS, TG_SMEM_BYTES, s>>>(b, nmodels, mp, partial);
  align_centered_kernel<<<dim3(nmodels, nmodels), 256, 0, s>>>(b, nmodels, mp, nsplit, partial, rs, cmat);
  align_normalise_kernel<<<(nmodels * nmodels + 255) / 256, 256, 0, s>>>(nmodels, cmat, out);
  g_launch_count += 4;
  return 0;
}

int configure_align_kernels() {
  return (int)cudaFuncSetAttribute(align_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES);
}

}  // namespace b200gp
