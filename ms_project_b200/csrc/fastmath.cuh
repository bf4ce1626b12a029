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
  // exp(r), |r| <= ln2 / 2: Taylor to r^13 (remainder 4e-18 relative) as even(r^2) + r odd(r^2): two Horner chains of
  // depth 7 with ONE constant operand per DFMA (Estrin's pairs need two, i.e. an extra LDC each); with the two elements
  // a thread interleaves that is four independent chains
  const double r2 = r * r;
  double pe = FM_EXP[1];                      // 1/12!
  double po = FM_EXP[0];                      // 1/13!
  pe = fma(pe, r2, FM_EXP[3]); po = fma(po, r2, FM_EXP[2]);      // 1/10!, 1/11!
  pe = fma(pe, r2, FM_EXP[5]); po = fma(po, r2, FM_EXP[4]);      // 1/8!,  1/9!
  pe = fma(pe, r2, FM_EXP[7]); po = fma(po, r2, FM_EXP[6]);      // 1/6!,  1/7!
  pe = fma(pe, r2, FM_EXP[9]); po = fma(po, r2, FM_EXP[8]);      // 1/4!,  1/5!
  pe = fma(pe, r2, FM_EXP[11]); po = fma(po, r2, FM_EXP[10]);    // 1/2!,  1/3!
  pe = fma(pe, r2, FM_C[4]); po = fma(po, r2, FM_C[4]);          // 1,     1
  const double p = fma(po, r, pe);
  // p * 2^n by exponent arithmetic (normal range).  x < -708 flushes to (sub)zero -- the test is on x itself because
  // the magic-number rounding above only holds n for |x| < 1.5e15 (tiny lengthscales reach far beyond that).  A NaN x
  // gives t = r = p = NaN with n = 0, so it passes through untouched.  x > 709 is outside the domain (the leaf
  // exponents are <= 0).
  int hi = __double2hiint(p) + (n << 20);
  hi = (x < -708.0) ? 0 : hi;
  return __hiloint2double(hi, __double2loint(p));
}

// Table-driven exp for the polynomial-form kernels: exp(x) = 2^e * 2^(j/64) * exp(r), n = round(64 x / ln 2) = 64 e + j,
// |r| <= ln 2 / 128, exp(r) - 1 by a degree-5 polynomial (truncation 3.5e-17).  10 FP64 operations on the shared
// FP64 datapath instead of the 18 of exp_sl, plus one 8-byte shared-memory load (the table is staged per CTA: constant
// memory would serialise the 32 different indices of a warp).  Worst error 1.24 ulp over the leaf domain
// (checked against 200-bit arithmetic).  Same edge behaviour as exp_sl: x < -708 -> 0, NaN -> NaN, x <= 709.
__constant__ double FM_EXPTAB[64] = {
    1, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.1023825833078409, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.2021567314527031, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.2553807570246911, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.3396675240533029,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.5590044002378369, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.6457554781539649, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.7186192981224779, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.9784560263879509};
__constant__ double FM_T[8] = {92.332482616893656877 /* 64 / ln 2 */, 6755399441055744.0, -6.93147180369123816490e-01 / 64.0,
                              -1.90821492927058770002e-10 / 64.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};

__device__ __forceinline__ double exp_tab(double x, unsigned tab /* shared-window address of a copy of FM_EXPTAB */) {
  const double t = fma(x, FM_T[0], FM_T[1]);            // round(64 x / ln 2) in the low word
  const int n = __double2loint(t);
  const double fn = t - FM_T[1];
  double r = fma(fn, FM_T[2], x);
  r = fma(fn, FM_T[3], r);
  double T;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab + ((unsigned)(n & 63) << 3)));
  const double r2 = r * r;
  const double A = fma(r, FM_T[5], FM_T[4]);
  const double B = fma(r, FM_T[7], FM_T[6]);
  const double C = fma(r2, B, A);
  const double p = fma(r2, C, r);                       // exp(r) - 1
  const double res = fma(T, p, T);
  int hi = __double2hiint(res) + ((n >> 6) << 20);
  hi = (x < -708.0) ? 0 : hi;
  return __hiloint2double(hi, __double2loint(res));
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
  // q = z (2/3 + 2/5 z + ... + 2/19 z^8) as even(z^2) + z odd(z^2), one constant operand per DFMA
  const double z2 = z * z;
  double qe = FM_LOG[0];                      // 2/19 (z^8)
  double qo = FM_LOG[1];                      // 2/17 (z^7)
  qe = fma(qe, z2, FM_LOG[2]); qo = fma(qo, z2, FM_LOG[3]);      // 2/15 (z^6), 2/13 (z^5)
  qe = fma(qe, z2, FM_LOG[4]); qo = fma(qo, z2, FM_LOG[5]);      // 2/11 (z^4), 2/9 (z^3)
  qe = fma(qe, z2, FM_LOG[6]); qo = fma(qo, z2, FM_LOG[7]);      // 2/7 (z^2),  2/5 (z)
  qe = fma(qe, z2, FM_LOG[8]);                                   // 2/3
  const double q = fma(qo, z, qe) * z;
  const double lf = fma(s, q, s + s);
  const double fe = (double)e;
  const double res = fma(fe, -FM_C[2], fma(fe, -FM_C[3], lf));
  return res + (m - m);                              // NaN (and inf) in -> NaN out, one DADD instead of a select
}

}  // namespace b200gp
