// Table-driven exp() for the per-individual kernels (sm_100a), also compiled on the host for its accuracy test.
//
// Why: the SE kernel value exp(theta1 - exp(-theta2)/2 (t_a - t_b)^2) is evaluated n(n+1)/2 times when Psi_i is built and again
// in the gradient pass of every objective evaluation. libdevice's exp() is 46 instructions here (13 DFMA of a degree-11
// polynomial whose 64-bit immediates each cost two UMOV): 48 warp-level calls = 19 % of the 12.7 k warp instructions of one
// evaluation at n = 30 (profiles/r02s2_eval2_c3_ncu.txt). This one: x = (64 m + j) ln2/64 + r, |r| <= ln2/128,
// exp(x) = 2^m * T[j] * (1 + r + r^2/2 + ... + r^5/120): 10 FP64 instructions, one table load, the truncation error of the
// polynomial is r^6/720 <= 3.5e-17 relative. Against libm over [-708, 709]: <= 1 ulp (tests/test_exp_table.py asserts 2).
// NaN propagates, x > 709.77 gives +inf, x < -708.39 gives 0 (no gradual underflow: the
// kernels add these values to sums of O(1) terms).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define MCB_EXP_HD __host__ __device__ __forceinline__
#else
#define MCB_EXP_HD static inline
#endif

namespace mcb {

#ifdef __CUDACC__
__device__
#endif
static const double kExpTab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

#define MCB_EXP_INVL 92.332482616893656877                 /* 64 / ln 2 */
#define MCB_EXP_LHI (0x1.62e42feep-1 / 64.0)                /* ln2 / 64, upper 32 bits: k * LHI is exact for |k| < 2^20 */
#define MCB_EXP_LLO (0x1.a39ef35793c76p-33 / 64.0)          /* (ln2 - ln2_hi) / 64 */
#define MCB_EXP_MAGIC 6755399441055744.0                    /* 1.5 * 2^52: adding it rounds to the nearest integer */
#ifdef __CUDACC__
// device code reads the constants from the constant bank (a DFMA operand) instead of materialising 64-bit immediates
__constant__ static double kExpC[8] = {MCB_EXP_INVL, -MCB_EXP_LHI, -MCB_EXP_LLO, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 0.0};
#endif

MCB_EXP_HD double exp_tab(double x, const double* tab) {
#ifdef __CUDA_ARCH__
  const double kInvL = kExpC[0], nLhi = kExpC[1], nLlo = kExpC[2], c5 = kExpC[3], c4 = kExpC[4], c3 = kExpC[5], c2 = kExpC[6];
#else
  const double kInvL = MCB_EXP_INVL, nLhi = -MCB_EXP_LHI, nLlo = -MCB_EXP_LLO, c5 = 1.0 / 120.0, c4 = 1.0 / 24.0, c3 = 1.0 / 6.0, c2 = 0.5;
#endif
  const double kMagic = MCB_EXP_MAGIC;
  double t = fma(x, kInvL, kMagic);
#ifdef __CUDA_ARCH__
  const int k = __double2loint(t);
#else
  int64_t bits; memcpy(&bits, &t, 8);
  const int k = (int)(uint32_t)bits;
#endif
  const double kd = t - kMagic;
  double r = fma(kd, nLhi, x);
  r = fma(kd, nLlo, r);
  const double r2 = r * r;
  double p = fma(r, c5, c4);
  p = fma(p, r, c3);
  p = fma(p, r, c2);
  const double q = fma(p, r2, r);                           // exp(r) - 1
#ifdef __CUDA_ARCH__
  const double tj = __ldg(tab + (k & 63));
#else
  const double tj = tab[k & 63];
#endif
  double v = fma(tj, q, tj);
  const int m = k >> 6;                                     // arithmetic shift: floor
#ifdef __CUDA_ARCH__
  const double sc = __hiloint2double((m + 1023) << 20, 0);
#else
  int64_t sb = (int64_t)(m + 1023) << 52; double sc; memcpy(&sc, &sb, 8);
#endif
  v *= sc;
  // outside [-708.39, 709.77] the scale factor above is not a normal number; NaN fails both comparisons and has already
  // propagated through r
  if (x < -708.39) v = 0.0;
  if (x > 709.77) v = INFINITY;
  return v;
}

}  // namespace mcb
