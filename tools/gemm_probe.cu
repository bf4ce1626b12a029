// Microbenchmark for the FP64 DMMA tile GEMM:  C(TM x TN) = A(TM x K) * B(TN x K)^T  ("NT", both K-contiguous),
// the shape of every blocked phase in ms_project_b200/csrc/chol_kernels.cu.  Times design variants so the library
// kernel is chosen from measurements:
//   WM x WN warps of 32 x 32 warp tiles, STAGES-deep cp.async ring of KC-wide chunks, MINB CTAs / SM,
//   SKEW = 1: fragments of the next chunk are fetched before the chunk barrier's DMMAs drain (see tile_gemm.cuh).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gemm_probe tools/gemm_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// PATTERN 1 emulates kinv_kernel (K^-1 = V V^T, V upper triangular): blockIdx.x walks the (TM x TN) blocks on or below
// the diagonal of the ld x ld output (longest K first), block row ti only needs k >= ti TM, the first TM columns of its A
// operand are upper triangular (row m needs k >= m: per-warp kc_lo) and warp tiles strictly above the diagonal are
// skipped (kc_hi = 0); the per-warp range is applied with the runtime predicate of SKEW = 2.
template <int WM, int WN, int STAGES, int KC, int MINB, int SKEW, int PATTERN = 0>
__global__ void __launch_bounds__(32 * WM * WN, MINB)
gemm_variant(const double* __restrict__ M, double* __restrict__ Cout, int ld, int K, int tiles_m, int tiles_n, int kc_lo, int kc_hi) {
  constexpr int TM = 32 * WM, TN = 32 * WN, NT = 32 * WM * WN, LDK = KC + 4;
  constexpr int STAGE = (TM + TN) * LDK;
  extern __shared__ __align__(16) double smem[];
  const int slot = blockIdx.y;
  int ti = blockIdx.x / tiles_n, tj = blockIdx.x % tiles_n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WN, wn = (warp - wm) % WN;
  const int g = lane >> 2, tg = lane & 3;
  int nchunks = K / KC;
  int kbase = 0;
  if (PATTERN == 1) {
    // enumerate blocks (ti, tj) with tj TN < (ti + 1) TM, block rows in ascending order (= descending K length)
    int rem = blockIdx.x;
    ti = 0;
    while (true) {
      const int cnt = ((ti + 1) * TM + TN - 1) / TN;
      if (rem < cnt) break;
      rem -= cnt;
      ti++;
    }
    tj = rem;
    kbase = ti * TM;
    nchunks = (ld - kbase) / KC;
    kc_lo = (wm * 32) / KC;
    kc_hi = (tj * TN + wn * 32 >= ti * TM + wm * 32 + 32) ? 0 : nchunks;
  }
  if (PATTERN == 2) {      // same enumeration, skipping at CTA granularity only (no per-warp predicate: use SKEW = 1)
    int rem = blockIdx.x;
    ti = 0;
    while (true) {
      const int cnt = ((ti + 1) * TM + TN - 1) / TN;
      if (rem < cnt) break;
      rem -= cnt;
      ti++;
    }
    tj = rem;
    kbase = ti * TM;
    nchunks = (ld - kbase) / KC;
  }
  const double* A = M + (size_t)slot * ld * ld + (size_t)(ti * TM) * ld + kbase;
  const double* B = M + (size_t)slot * ld * ld + (size_t)(tj * TN) * ld + kbase;
  double* C = Cout + (size_t)slot * ld * ld + (size_t)(ti * TM) * ld + tj * TN;

  auto load_stage = [&](int s, int koff) {
    double* As = smem + s * STAGE;
    double* Bs = As + TM * LDK;
    constexpr int SEG_ROW = KC / 2;
#pragma unroll
    for (int r = 0; r < (TM * SEG_ROW + NT - 1) / NT; r++) {
      const int seg = threadIdx.x + NT * r;
      if ((TM * SEG_ROW) % NT == 0 || seg < TM * SEG_ROW) {
        const int row = seg / SEG_ROW, part = (seg % SEG_ROW) * 2;
        cp_async16(As + row * LDK + part, A + (size_t)row * ld + koff + part);
      }
    }
#pragma unroll
    for (int r = 0; r < (TN * SEG_ROW + NT - 1) / NT; r++) {
      const int seg = threadIdx.x + NT * r;
      if ((TN * SEG_ROW) % NT == 0 || seg < TN * SEG_ROW) {
        const int row = seg / SEG_ROW, part = (seg % SEG_ROW) * 2;
        cp_async16(Bs + row * LDK + part, B + (size_t)row * ld + koff + part);
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nchunks) load_stage(s, s * KC);
    cp_async_commit();
  }
  const int aoff = (wm * 32 + g) * LDK + tg, boff = TM * LDK + (wn * 32 + g) * LDK + tg;
  if (SKEW == 0) {   // SKEW: 0 barrier at chunk head, 1 skewed, 2 skewed + runtime per-warp activity predicate
    for (int kc = 0; kc < nchunks; kc++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        const int nk = kc + STAGES - 1;
        if (nk < nchunks) load_stage(nk % STAGES, nk * KC);
        cp_async_commit();
      }
      const double* As = smem + (kc % STAGES) * STAGE + aoff;
      const double* Bs = smem + (kc % STAGES) * STAGE + boff;
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
        double a[4], b[4];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = As[mt * 8 * LDK + kk];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) b[nt] = Bs[nt * 8 * LDK + kk];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
      }
    }
  } else {
    // skewed: the chunk barrier sits in front of the LAST k-step's DMMAs, and the first fragments of the next
    // chunk are fetched right behind it, so the tensor pipe has 16 DMMAs queued while the block re-synchronises
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    double a[4], b[4];
    {
      const double* As = smem + aoff;
      const double* Bs = smem + boff;
#pragma unroll
      for (int mt = 0; mt < 4; mt++) a[mt] = As[mt * 8 * LDK];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) b[nt] = Bs[nt * 8 * LDK];
    }
    for (int kc = 0; kc < nchunks; kc++) {
      {
        const int nk = kc + STAGES - 1;
        if (nk < nchunks) load_stage(nk % STAGES, nk * KC);
        cp_async_commit();
      }
      const double* As = smem + (kc % STAGES) * STAGE + aoff;
      const double* Bs = smem + (kc % STAGES) * STAGE + boff;
      const double* An = smem + ((kc + 1) % STAGES) * STAGE + aoff;
      const double* Bn = smem + ((kc + 1) % STAGES) * STAGE + boff;
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
        double a2[4], b2[4];
        if (kk + 4 < KC) {
#pragma unroll
          for (int mt = 0; mt < 4; mt++) a2[mt] = As[mt * 8 * LDK + kk + 4];
#pragma unroll
          for (int nt = 0; nt < 4; nt++) b2[nt] = Bs[nt * 8 * LDK + kk + 4];
        } else {
          cp_async_wait<STAGES - 2>();
          __syncthreads();
          if (kc + 1 < nchunks) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a2[mt] = An[mt * 8 * LDK];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b2[nt] = Bn[nt * 8 * LDK];
          }
        }
        if (SKEW != 2 || (kc >= kc_lo && kc < kc_hi)) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = a2[mt];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) b[nt] = b2[nt];
      }
    }
  }
  cp_async_wait<0>();
  double* Cw = C + (size_t)(wm * 32 + g) * ld + wn * 32 + tg * 2;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      double2 v;
      v.x = acc[mt][nt][0]; v.y = acc[mt][nt][1];
      *reinterpret_cast<double2*>(Cw + (size_t)(mt * 8) * ld + nt * 8) = v;
    }
}

template <int WM, int WN, int STAGES, int KC, int MINB, int SKEW>
void run(const char* name, const double* M, double* C, int ld, int slots, const std::vector<int>& Ks, std::vector<double>* check) {
  constexpr int TM = 32 * WM, TN = 32 * WN;
  const int smem = STAGES * (TM + TN) * (KC + 4) * 8;
  auto fn = gemm_variant<WM, WN, STAGES, KC, MINB, SKEW>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * WM * WN, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, fn));
  printf("%-34s tile %3dx%3d thr %3d smem %6d regs %3d CTAs/SM %d |", name, TM, TN, 32 * WM * WN, smem, fa.numRegs, per_sm);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int K : Ks) {
    dim3 grid((ld / TM) * (ld / TN), slots);
    fn<<<grid, 32 * WM * WN, smem>>>(M, C, ld, K, ld / TM, ld / TN, 0, 1 << 30);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) fn<<<grid, 32 * WM * WN, smem>>>(M, C, ld, K, ld / TM, ld / TN, 0, 1 << 30);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * ld * ld * (double)K * slots * reps / (ms * 1e-3) / 1e12;
    printf("  K=%4d %6.2f TF/s", K, tf);
  }
  // checksum of slot 0, first row block (compare variants)
  std::vector<double> h(ld);
  CK(cudaMemcpy(h.data(), C + 5 * ld, ld * sizeof(double), cudaMemcpyDeviceToHost));
  double s = 0;
  for (double v : h) s += v;
  printf("  chk %.10e\n", s);
  if (check) check->push_back(s);
}

template <int WM, int WN, int STAGES, int KC, int MINB, int PATTERN = 1>
void run_kinv(const char* name, const double* M, double* C, int ld, int slots) {
  constexpr int TM = 32 * WM, TN = 32 * WN;
  const int smem = STAGES * (TM + TN) * (KC + 4) * 8;
  auto fn = gemm_variant<WM, WN, STAGES, KC, MINB, PATTERN == 1 ? 2 : 1, PATTERN>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * WM * WN, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, fn));
  int nblk = 0;
  for (int ti = 0; ti < ld / TM; ti++) nblk += ((ti + 1) * TM + TN - 1) / TN;
  dim3 grid(nblk, slots);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  fn<<<grid, 32 * WM * WN, smem>>>(M, C, ld, ld, ld / TM, ld / TN, 0, 0);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  const int reps = 3;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) fn<<<grid, 32 * WM * WN, smem>>>(M, C, ld, ld, ld / TM, ld / TN, 0, 0);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double tf = (double)ld * ld * ld / 3.0 * slots * reps / (ms * 1e-3) / 1e12;
  printf("kinv pattern %-22s tile %3dx%3d thr %3d smem %6d regs %3d CTAs/SM %d blocks/slot %4d | %6.2f TF/s algorithmic (N^3/3), %.3f ms per slot\n",
         name, TM, TN, 32 * WM * WN, smem, fa.numRegs, per_sm, nblk, tf, ms / reps / slots);
}

__global__ void fill(double* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = i * 6364136223846793005ull + 1442695040888963407ull;
    x ^= x >> 29;
    p[i] = (double)(x & 0xffff) / 65536.0 - 0.5;
  }
}

int main(int argc, char** argv) {
  const int ld = 2048;
  const int slots = argc > 1 ? atoi(argv[1]) : 48;
  double *M, *C;
  CK(cudaMalloc(&M, (size_t)slots * ld * ld * 8));
  CK(cudaMalloc(&C, (size_t)slots * ld * ld * 8));
  fill<<<1184, 256>>>(M, (size_t)slots * ld * ld);
  CK(cudaDeviceSynchronize());
  std::vector<int> Ks = {128, 512, 2048};
  std::vector<double> chk;
  if (argc > 2) {      // kinv-pattern comparison of CTA tile shapes
    run_kinv<4, 2, 3, 16, 2>("4x2 S3 x2 (library)", M, C, ld, slots);
    run_kinv<2, 2, 3, 16, 3>("2x2 S3 x3", M, C, ld, slots);
    run_kinv<2, 2, 2, 16, 4>("2x2 S2 x4", M, C, ld, slots);
    run_kinv<2, 4, 3, 16, 2>("2x4 S3 x2", M, C, ld, slots);
    run_kinv<4, 4, 3, 16, 1>("4x4 S3 x1", M, C, ld, slots);
    run_kinv<2, 1, 4, 16, 6>("2x1 S4 x6", M, C, ld, slots);
    printf("CTA-granular skipping only (no per-warp predicate):\n");
    run_kinv<4, 2, 3, 16, 2, 2>("4x2 S3 x2", M, C, ld, slots);
    run_kinv<2, 2, 3, 16, 3, 2>("2x2 S3 x3", M, C, ld, slots);
    run_kinv<2, 2, 2, 16, 4, 2>("2x2 S2 x4", M, C, ld, slots);
    run_kinv<1, 2, 4, 16, 6, 2>("1x2 S4 x6", M, C, ld, slots);
    run_kinv<1, 1, 4, 16, 8, 2>("1x1 S4 x8", M, C, ld, slots);
    run_kinv<2, 2, 2, 32, 3, 2>("2x2 S2 KC32 x3", M, C, ld, slots);
    run_kinv<2, 2, 4, 8, 3, 2>("2x2 S4 KC8 x3", M, C, ld, slots);
    run_kinv<2, 2, 5, 8, 4, 2>("2x2 S5 KC8 x4", M, C, ld, slots);
    run_kinv<2, 2, 3, 16, 2, 2>("2x2 S3 x2", M, C, ld, slots);
    return 0;
  }
  run<4, 2, 3, 16, 2, 1>("skew 4x2 S3 KC16 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 3, 16, 2, 2>("skew+pred 4x2 S3 KC16 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 2, 32, 2, 1>("skew 4x2 S2 KC32 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 2, 32, 2, 2>("skew+pred 4x2 S2 KC32 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 5, 8, 2, 1>("skew 4x2 S5 KC8 x2", M, C, ld, slots, Ks, &chk);
  return 0;
  run<4, 4, 4, 16, 1, 0>("base 4x4 S4 KC16", M, C, ld, slots, Ks, &chk);
  run<4, 4, 4, 16, 1, 1>("skew 4x4 S4 KC16", M, C, ld, slots, Ks, &chk);
  run<4, 4, 3, 16, 1, 0>("base 4x4 S3 KC16", M, C, ld, slots, Ks, &chk);
  run<4, 4, 8, 8, 1, 0>("base 4x4 S8 KC8", M, C, ld, slots, Ks, &chk);
  run<4, 4, 2, 32, 1, 0>("base 4x4 S2 KC32", M, C, ld, slots, Ks, &chk);
  run<4, 2, 3, 16, 2, 0>("base 4x2 S3 KC16 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 3, 16, 2, 1>("skew 4x2 S3 KC16 x2", M, C, ld, slots, Ks, &chk);
  run<2, 4, 3, 16, 2, 0>("base 2x4 S3 KC16 x2", M, C, ld, slots, Ks, &chk);
  run<4, 2, 4, 8, 2, 0>("base 4x2 S4 KC8 x2", M, C, ld, slots, Ks, &chk);
  run<2, 2, 4, 16, 4, 0>("base 2x2 S4 KC16 x4", M, C, ld, slots, Ks, &chk);
  run<2, 2, 3, 16, 4, 1>("skew 2x2 S3 KC16 x4", M, C, ld, slots, Ks, &chk);
  return 0;
}
