// 64 x 64 CTA tile GEMM on the DMMA pipe, operands fed by the TMA: the producer side of tile_gemm64.cuh's pipeline replaced
// by `cp.async.bulk.tensor.2d` (SASS UTMALDG) into an mbarrier-guarded ring -- no CTA-wide barrier anywhere in the K loop.
//
//   * one elected lane (warp 0, lane 0) issues TWO tensor copies per K chunk (A box and B box, 64 rows x 16 doubles each)
//     instead of the 1024 16-byte LDGSTS copies per chunk of the cp.async form (8 per thread plus their address arithmetic);
//   * full[stage]  (count 1 + 16 KB of transaction bytes): the TMA completes it; every consumer warp waits on it by itself;
//   * empty[stage] (count 4): each warp arrives once it has consumed the stage; only the issuing lane ever waits on it, and
//     only for the stage consumed TWO iterations ago -- a warp may run one whole chunk ahead of its siblings, where the
//     __syncthreads of the cp.async form tied every warp to the slowest one in every chunk;
//   * 4 stages of 16 KB (the padding of the cp.async rows is gone): the same 64 KB ring, prefetch distance 2 chunks;
//   * shared-memory layout = what the tensor map's swizzle writes.  SWIZZLE_128B_ATOM_32B XORs the 32-byte atom index of a
//     128-byte row with (row & 3): a DMMA operand fetch (8 rows x 4 consecutive doubles per quarter-warp group) then touches
//     32 distinct banks per half warp -- conflict-free without padding.  SW = 1 (16-byte atoms, row & 7) and SW = 0 (none)
//     exist for comparison.
// Every output element still sums its products in ascending k: results are bit-identical to the cp.async kernels.
// Measured on B200 (256 C2 items, N = 2048; tools/tma_probe.py): Cholesky + L^-T + K^-1 70.65 ms with cp.async, 72.6 with
// unswizzled tensor maps (4-way bank conflicts), 69.5 with SWIZZLE_128B, 69.3 with SWIZZLE_128B_ATOM_32B (kept); a 3-stage
// ring with four CTAs per SM 69.3-69.5 (no gain).  Tried on top and removed: letting a warp whose first chunks are
// structurally zero (V_II / diagonal blocks) consume and release them without loading or issuing -- possible now that no
// block barrier ties the warps -- 69.3 -> 69.7 ms: the pipe is already ~95 % busy on the long tiles, the skipped 2 % of DMMAs
// do not pay for the extra control flow.
#pragma once
#include <cuda.h>
#include "tile_gemm64.cuh"

namespace b200gp {

constexpr int K64T_TILE_BYTES = K64_T * TG_KC * 8;                 // one operand box: 64 rows x 128 B = 8 KB
constexpr int K64T_STAGE_BYTES = 2 * K64T_TILE_BYTES;              // A + B
constexpr int k64t_smem_bytes(int stages) { return stages * K64T_STAGE_BYTES + 2 * stages * 8 + 1024; }   // + 1 KB alignment slack

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// byte offset, inside a 128-byte row, of the 4 consecutive doubles [4 step, 4 step + 4) -- plus this thread's element tg of
// them -- as the tensor map's swizzle mode SW left row r (r & 7 = g for every fragment row of this thread) in shared memory
template <int SW>
__device__ __forceinline__ int k64t_swz(int g, int tg, int step) {
  if (SW == 2) return (((step ^ g) & 3) << 5) + (tg << 3);                              // 32-byte atoms ^ (row & 3)
  if (SW == 1) return ((((2 * step + (tg >> 1)) ^ g) & 7) << 4) + ((tg & 1) << 3);      // 16-byte atoms ^ (row & 7)
  return (step << 5) + (tg << 3);
}

// acc (this warp's 32 x 32 part of the 64 x 64 block) += A(64 x 16 nchunks) B(64 x 16 nchunks)^T.  A = rows [ay, ay + 64),
// columns from ax of the matrix behind mapA; B likewise.  `smem_raw`: the kernel's dynamic shared memory (k64t_smem_bytes(STAGES)).
// Must be called by all 128 threads of the CTA, once per kernel (the barriers are initialised here).
template <int SW, int K64T_STAGES>
__device__ __forceinline__ void tile64_pipeline_tma(double (&acc)[TG_MT][TG_NT][2], const CUtensorMap* mapA, int ax, int ay,
                                                    const CUtensorMap* mapB, int bx, int by, int nchunks,
                                                    unsigned char* smem_raw) {
  const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // swizzled boxes want 1 KB alignment
  const unsigned bars = base + K64T_STAGES * K64T_STAGE_BYTES;        // full[0..S), empty[0..S)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, tg = lane & 3;
  const bool issuer = threadIdx.x == 0;
  if (issuer) {
#pragma unroll
    for (int s = 0; s < K64T_STAGES; s++) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (K64T_STAGES + s), K64_THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int chunk) {       // issuer only
    const int s = chunk % K64T_STAGES;
    const unsigned full = bars + 8 * s, dst = base + s * K64T_STAGE_BYTES;
    mbar_expect_tx(full, K64T_STAGE_BYTES);
    tma_load_2d(dst, mapA, ax + chunk * TG_KC, ay, full);
    tma_load_2d(dst + K64T_TILE_BYTES, mapB, bx + chunk * TG_KC, by, full);
  };
  if (issuer) {
#pragma unroll
    for (int c = 0; c < K64T_STAGES - 2; c++)
      if (c < nchunks) issue(c);
  }
  // this thread's fragment addresses inside a stage: row base + swizzled position of (k = 4 step + tg); the swizzle term
  // depends on (row & 7) = g only, so A and B share it
  const int arow = (wm * 32 + g) * 128, brow = K64T_TILE_BYTES + (wn * 32 + g) * 128;
  int swz[4];
#pragma unroll
  for (int st = 0; st < 4; st++) swz[st] = k64t_swz<SW>(g, tg, st);
  auto ld = [&](unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  };
  double a[TG_MT], bf[TG_NT];
  mbar_wait(bars, 0);                                                  // chunk 0 has landed
#pragma unroll
  for (int mt = 0; mt < TG_MT; mt++) a[mt] = ld(base + arow + swz[0] + mt * 8 * 128);
#pragma unroll
  for (int nt = 0; nt < TG_NT; nt++) bf[nt] = ld(base + brow + swz[0] + nt * 8 * 128);
  for (int kc = 0; kc < nchunks; kc++) {
    const int s = kc % K64T_STAGES;
    if (issuer) {                      // refill the stage consumed two iterations ago with chunk kc + 2
      const int nk = kc + K64T_STAGES - 2;
      if (nk < nchunks) {
        if (nk >= K64T_STAGES) mbar_wait(bars + 8 * (K64T_STAGES + nk % K64T_STAGES), ((nk / K64T_STAGES) - 1) & 1);
        issue(nk);
      }
    }
    const unsigned st_base = base + s * K64T_STAGE_BYTES;
    const unsigned nx_base = base + ((kc + 1) % K64T_STAGES) * K64T_STAGE_BYTES;
#pragma unroll
    for (int step = 0; step < 4; step++) {
      double a2[TG_MT], b2[TG_NT];
      if (step < 3) {
#pragma unroll
        for (int mt = 0; mt < TG_MT; mt++) a2[mt] = ld(st_base + arow + swz[step + 1] + mt * 8 * 128);
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) b2[nt] = ld(st_base + brow + swz[step + 1] + nt * 8 * 128);
      } else if (kc + 1 < nchunks) {    // skew: the next chunk's first fragments are fetched behind its full barrier
        mbar_wait(bars + 8 * ((kc + 1) % K64T_STAGES), ((kc + 1) / K64T_STAGES) & 1);
#pragma unroll
        for (int mt = 0; mt < TG_MT; mt++) a2[mt] = ld(nx_base + arow + swz[0] + mt * 8 * 128);
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) b2[nt] = ld(nx_base + brow + swz[0] + nt * 8 * 128);
      }
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++)
#pragma unroll
        for (int nt = 0; nt < TG_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], bf[nt]);
      if (step == 3) {                  // every fragment of stage s is in registers (the DMMAs above consumed the last ones)
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (K64T_STAGES + s));
      }
#pragma unroll
      for (int mt = 0; mt < TG_MT; mt++) a[mt] = a2[mt];
#pragma unroll
      for (int nt = 0; nt < TG_NT; nt++) bf[nt] = b2[nt];
    }
  }
}

}  // namespace b200gp
