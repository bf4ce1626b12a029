// Branch-free FP64 exp / log1p / reciprocal for the element-wise kernels.
//
// CUDA's libm exp(), log1p() and the IEEE division each carry a rarely taken slow-path BRANCH.  A branch splits the
// basic block, so the compiler cannot interleave the polynomial chains of the 2 elements a thread works on, and the
// kernels end up latency-bound on dependent DFMAs (ncu: `wait` stalls, FP64 pipe 30 % busy).  These versions are
// straight-line code (selects instead of branches); accuracy about 1 ulp on the domains the leaf kernels use:
//   exp_sl(x)     x <= 709 (x < -708 flushes to 0), NaN -> NaN
//   log1p_pos(u)  u >= 0 (|error| <~ 2e-16 absolute + 3e-16 relative), NaN -> NaN
//   rcp_sl(m)     finite normal m
#pragma once
#include <cuda_runtime.h>

namespace b200gp {

__device__ __forceinline__ double rcp_sl(double m) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m));     // MUFU.RCP64H: ~2^-23
  double e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  e = fma(-m, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Polynomial coefficients live in constant memory so the DFMAs take them as c[bank][offset] operands instead of
// re-materialising 64-bit immediates (two UMOVs each) at every inlined call site.
__constant__ double FM_EXP[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                  1.0 / 362880.0,     1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,
                                  1.0 / 120.0,        1.0 / 24.0,        1.0 / 6.0,        0.5};
__constant__ double FM_LOG[9] = {2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0};
__constant__ double FM_C[5] = {1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01,
                               -1.90821492927058770002e-10, 1.0};

__device__ __forceinline__ double exp_sl(double x) {
  const double t = fma(x, FM_C[0], FM_C[1]);            // round(x / ln 2) in the low word
  const int n = __double2loint(t);
  const double fn = t - FM_C[1];
  double r = fma(fn, FM_C[2], x);
  r = fma(fn, FM_C[3], r);
  // exp(r), |r| <= ln2 / 2: Taylor to r^13 (remainder 4e-18 relative), Estrin's scheme: dependency depth 5 instead
  // of Horner's 14, which is what a latency-bound kernel needs (3 extra multiplies)
  const double r2 = r * r;
  const double a0 = fma(r, FM_C[4], FM_C[4]);          // 1 + r
  const double a1 = fma(r, FM_EXP[10], FM_EXP[11]);    // 1/2! + r/3!
  const double a2 = fma(r, FM_EXP[8], FM_EXP[9]);
  const double a3 = fma(r, FM_EXP[6], FM_EXP[7]);
  const double a4 = fma(r, FM_EXP[4], FM_EXP[5]);
  const double a5 = fma(r, FM_EXP[2], FM_EXP[3]);
  const double a6 = fma(r, FM_EXP[0], FM_EXP[1]);      // 1/12! + r/13!
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(a3, r2, a2);
  const double b2 = fma(a5, r2, a4);
  const double r8 = r4 * r4;
  const double d0 = fma(b1, r4, b0);
  const double d1 = fma(a6, r4, b2);
  const double p = fma(d1, r8, d0);
  // p * 2^n by exponent arithmetic (normal range).  x < -708 flushes to (sub)zero -- the test is on x itself because
  // the magic-number rounding above only holds n for |x| < 1.5e15 (tiny lengthscales reach far beyond that).  A NaN x
  // gives t = r = p = NaN with n = 0, so it passes through untouched.  x > 709 is outside the domain (the leaf
  // exponents are <= 0).
  int hi = __double2hiint(p) + (n << 20);
  hi = (x < -708.0) ? 0 : hi;
  return __hiloint2double(hi, __double2loint(p));
}

__device__ __forceinline__ double log1p_pos(double u) {
  const double m = 1.0 + u;
  int hi = __double2hiint(m);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;               // f in [1, 2)
  const bool up = hi >= 0x3ff6a09f;                  // f >= sqrt 2: halve it
  hi = up ? hi - 0x00100000 : hi;
  e = up ? e + 1 : e;
  const double f = __hiloint2double(hi, __double2loint(m));
  const double s = (f - 1.0) * rcp_sl(f + 1.0);      // log f = 2 atanh s, |s| <= 0.1716
  const double z = s * s;
  // q = z (2/3 + 2/5 z + ... + 2/19 z^8), Estrin
  const double z2 = z * z;
  const double a0 = fma(z, FM_LOG[7], FM_LOG[8]);
  const double a1 = fma(z, FM_LOG[5], FM_LOG[6]);
  const double a2 = fma(z, FM_LOG[3], FM_LOG[4]);
  const double a3 = fma(z, FM_LOG[1], FM_LOG[2]);
  const double z4 = z2 * z2;
  const double b0 = fma(a1, z2, a0);
  const double b1 = fma(a3, z2, a2);
  const double z8 = z4 * z4;
  const double d0 = fma(b1, z4, b0);
  const double q = fma(FM_LOG[0], z8, d0) * z;
  const double lf = fma(s, q, s + s);
  const double fe = (double)e;
  const double res = fma(fe, -FM_C[2], fma(fe, -FM_C[3], lf));
  return res + (m - m);                              // NaN (and inf) in -> NaN out, one DADD instead of a select
}

}  // namespace b200gp
