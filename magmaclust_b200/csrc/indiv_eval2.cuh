// Objective + gradient of one individual's GP at a trial theta (second design; indiv_eval.cuh is the first):
//   F_i(theta)  = logL_clust_multi_GP   (CF.R:218-248)
//   dF_i/dtheta = gr_clust_multi_GP     (CF.R:477-508)
// One thread group of TT threads (a CTA) per individual, n = n_i observations, ntot = n + 2 with the borders y, corr1.
// Everything is sized by the compile-time tile count NT >= ceil(n / 8): leading dimension, k steps and loop bounds
// are constants and the zero padding makes the extra work exact zeros (no per-instruction guards).
//
// What changed against the first design, and why (tools/dbg/eval_prof.cu + ncu source page: 38 k warp instructions
// and 60 k cycles per evaluation at n = 50, 2 CTAs per SM because of four n x n shared-memory buffers):
//   * two n x n buffers instead of four (W and S = corr2): dPsi/dtheta is recomputed in the epilogue instead of being
//     mirrored to shared memory, and X = W S never leaves registers -> 4 CTAs per SM at n = 50
//   * permuted contraction index in the mma.m8n8k4 products: k step 2m takes columns 8m + 2 gc, step 2m + 1 takes
//     columns 8m + 2 gc + 1 (any bijection works as long as A and B agree). Both fragments of two consecutive k
//     steps are then ONE 16-byte shared-memory load, and the accumulator pair of X (row gr, columns 8j + 2 gc, + 1)
//     IS the A-fragment pair of H = X W: the two products chain through registers with no shuffle, no shared-memory
//     round trip and no CTA barrier. ld = 8 mod 16 doubles keeps the 16-byte loads of a quarter warp conflict free
//   * row tiles dealt to warps by cost (a row tile costs NT + ti + 1 tile products), roles rotated per individual
//   * sweep with look-ahead: the owners of the NEXT pivot column update and publish it before the rank-1 update of
//     the rest; reciprocal of the pivot by rcp.approx + 2 Newton steps instead of an IEEE division
//   * W is written row-wise with 16-byte stores (lower block triangle) and mirrored tile by tile; the first design's
//     transposed scalar stores were 8- to 13-way bank conflicted
//
// Phases:  A build (registers, 4 x 4 blocks of the lower block triangle)  B sweep  C post: W to shared memory, the
// block-layout partial sums  D per row tile: X = W S, H = X W reduced on the fly against dPsi/dth1, dPsi/dth2, I
// E block reduction -> f, g.   Flops as executed ~ n^3/2 + n^3 + n^3/2 FMAs (reference: n^3 + 4 n^3 + 4 n^3 flops).
#pragma once
#include "common.cuh"
#include "exp_tab.h"
#include "indiv_eval.cuh"  // dmma884_i, cta_sum_v, tsync, MCB_STAMP

namespace mcb {

// exp() out of line: the evaluation kernels are instruction-fetch bound (ncu: 24 % of the warp samples stall on
// no_inst at 4 CTAs per SM, I-cache hit rate 73 %), and every inlined exp is ~35 instructions at ~30 call sites
// and a table-driven one (exp_tab.h: 24 instead of 46 instructions per call; MCB_LIBM_EXP restores libdevice's for A/B runs)
#ifdef MCB_LIBM_EXP
__device__ __noinline__ double ev_exp(double x) { return exp(x); }
#else
__device__ __noinline__ double ev_exp(double x) { return exp_tab(x, kExpTab); }
#endif
// two at a time: half the calls, two independent dependency chains per call (the kernels stall mostly on fixed-latency
// dependencies: profiles/r02s2_eval2_c3_ncu.txt, "wait")
#ifdef MCB_LIBM_EXP
__device__ __noinline__ double2 ev_exp2(double2 x) { return make_double2(exp(x.x), exp(x.y)); }
#else
__device__ __noinline__ double2 ev_exp2(double2 x) { return make_double2(exp_tab(x.x, kExpTab), exp_tab(x.y, kExpTab)); }
#endif
struct Exp4 { double a, b, c, d; };   // passed and returned in registers
__device__ __noinline__ Exp4 ev_exp4(Exp4 x) {
  Exp4 r;
  r.a = exp_tab(x.a, kExpTab);
  r.b = exp_tab(x.b, kExpTab);
  r.c = exp_tab(x.c, kExpTab);
  r.d = exp_tab(x.d, kExpTab);
  return r;
}
__device__ __noinline__ double ev_log(double x) { return log(x); }

__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
}

// compile-time geometry of the shared-memory problem for NT 8-wide tiles
template <int NT>
struct Eval2Geo {
  static constexpr int NP8 = 8 * NT;                                      // rows of W and S
  static constexpr int LDM = (NP8 % 16 == 8) ? NP8 : NP8 + 8;             // = 8 mod 16
  static constexpr int NTOT4 = NP8 + 4;                                   // >= roundup4(n + 2)
  static constexpr int CS = NTOT4 + 4;                                    // stride of the two pivot-column buffers
  static constexpr int MAT = NP8 * LDM;
  static constexpr int DOUBLES = 2 * MAT + 2 * CS + NTOT4 + 2 + 3 * NP8 + 16 * 4 + 64;  // last 64: L^-1 tile of indiv_eval3
};

template <int NT>
struct EvalSm2 {
  double *W, *S;       // NP8 x LDM each (zero padded)
  double *col;         // two pivot-column buffers of CS doubles ([NTOT4] = 1 / pivot)
  double *piv;         // NTOT4
  double *corner;      // 2
  double *t, *y, *c1;  // NP8 each
  double *red;         // 16 * 4
};

template <int NT>
__device__ __forceinline__ EvalSm2<NT> eval2_carve(double* sm) {
  using G = Eval2Geo<NT>;
  EvalSm2<NT> s;
  s.W = sm;
  s.S = s.W + G::MAT;
  s.col = s.S + G::MAT;
  s.piv = s.col + 2 * G::CS;
  s.corner = s.piv + G::NTOT4;
  s.t = s.corner + 2;
  s.y = s.t + G::NP8;
  s.c1 = s.y + G::NP8;
  s.red = s.c1 + G::NP8;
  return s;
}

// One-time per individual: zero both matrices, load t / y / c1 and S = corr2 (n x n, row-major in global memory).
template <int TT, int NT>
__device__ __forceinline__ void eval2_load(const EvalSm2<NT>& s, int n, const double* __restrict__ tg,
                                           const double* __restrict__ yg, const double* __restrict__ c1g,
                                           const double* __restrict__ c2g) {
  using G = Eval2Geo<NT>;
  double2* z = reinterpret_cast<double2*>(s.W);
  for (int e = threadIdx.x; e < G::MAT; e += TT) z[e] = make_double2(0.0, 0.0);  // W and S are adjacent
  for (int e = threadIdx.x; e < 64; e += TT) s.red[64 + e] = 0.0;                 // L^-1 tile (zero above the diagonal)
  for (int a = threadIdx.x; a < n; a += TT) {
    s.t[a] = tg[a];
    s.y[a] = yg[a];
    s.c1[a] = c1g[a];
  }
  tsync<TT>();
  // S: a warp per row, lanes over the columns, RB rows in flight per thread (independent loads issued back to back, no
  // division). The element-per-thread loop this replaces (e / n, one dependent load per trip) was 12 % of the warp samples
  // of a batched evaluation at n = 30 (profiles/r02s3_eval2_c3_lines.txt).
  constexpr int NW = TT / 32;
  constexpr int CB = (8 * NT + 31) / 32;
  constexpr int RB = (NT <= 4) ? 8 : 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a0 = warp; a0 < n; a0 += RB * NW) {
    double v[RB][CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int a = a0 + i * NW;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const int b = lane + 32 * j;
        v[i][j] = (a < n && b < n) ? c2g[a * n + b] : 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int a = a0 + i * NW;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const int b = lane + 32 * j;
        if (a < n && b < n) s.S[a * G::LDM + b] = v[i][j];
      }
    }
  }
  tsync<TT>();
}

// block q of the lower block triangle -> (I, J), I >= J; -1 when q is past the last block of this individual
template <int E, int TT>
__device__ __forceinline__ void eval2_decode(int n, int (&bI)[E], int (&bJ)[E]) {
  const int nbk = (n + 2 + 3) >> 2;
  const int nblk = nbk * (nbk + 1) / 2;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int q = threadIdx.x + TT * e;
    if (q < nblk) {
      int I = (int)((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
      while ((I + 1) * (I + 2) / 2 <= q) ++I;
      while (I * (I + 1) / 2 > q) --I;
      bI[e] = I;
      bJ[e] = q - I * (I + 1) / 2;
    } else {
      bI[e] = -1;
      bJ[e] = -1;
    }
  }
}

// out[0] = f, out[1..3] = g (shared memory, written by thread 0). All TT threads must call. 8 NT >= n.
// rot: rotation of the warp roles in the tile phase (the caller varies it per individual so that the lighter
// roles do not always land on the same SM sub-partition).
template <int E, int TT, int NT>
__device__ void indiv_eval_v2(const EvalSm2<NT>& s, int n, const int (&bI)[E], const int (&bJ)[E], double th1,
                              double th2, double sg, bool want_grad, double* out, int* bad, int rot) {
  using G = Eval2Geo<NT>;
  constexpr int NW = TT / 32;
  constexpr int LDM = G::LDM;
  constexpr int NTOT4 = G::NTOT4;
  constexpr int CS = G::CS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntot = n + 2;
  const double hl = ev_exp(-th2) * 0.5;
  const double e1 = ev_exp(th1);
  const double dg = e1 + sg * sg;
#ifdef MCB_PROF
  long long prof_t_ = clock64();
#endif

  double blk[E][4][4], k0[E][4][4];

  // ---- A: build ---------------------------------------------------------------------------------
#pragma unroll
  for (int e = 0; e < E; ++e) {
    double tr[4], tc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = 4 * bI[e] + a, c = 4 * bJ[e] + a;
      tr[a] = (bI[e] >= 0 && r < n) ? s.t[r] : 0.0;
      tc[a] = (bI[e] >= 0 && c < n) ? s.t[c] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      // the four SE values of this row of the block, two per call (lanes outside the matrix compute exp(th1): harmless)
      double kr[4];
#ifndef MCB_EXP_SCALAR
      // measured on a B200 (gpurun_out/r02s3_exp*.log): four per call is the faster form at n = 50 (theta_i phase of C2
      // 4.26 -> 4.10 ms with pairs, 4.03 ms with fours), two per call at n = 30 (one evaluation of C3 2.48 -> 2.38 / 2.41 ms)
      if constexpr (NT >= 7) {
        const double d0 = tr[a] - tc[0], d1 = tr[a] - tc[1], d2 = tr[a] - tc[2], d3 = tr[a] - tc[3];
        const Exp4 k4 = ev_exp4(Exp4{th1 - hl * (d0 * d0), th1 - hl * (d1 * d1), th1 - hl * (d2 * d2), th1 - hl * (d3 * d3)});
        kr[0] = k4.a; kr[1] = k4.b; kr[2] = k4.c; kr[3] = k4.d;
      } else {
#pragma unroll
        for (int b = 0; b < 4; b += 2) {
          const double d0 = tr[a] - tc[b], d1 = tr[a] - tc[b + 1];
          const double2 k2 = ev_exp2(make_double2(th1 - hl * (d0 * d0), th1 - hl * (d1 * d1)));
          kr[b] = k2.x;
          kr[b + 1] = k2.y;
        }
      }
#endif
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double v = 0.0, kv = 0.0;
        if (bI[e] >= 0) {
          const int r = 4 * bI[e] + a, c = 4 * bJ[e] + b;
          if (r < n && c < n) {
            if (r == c) {
              v = dg;
              kv = e1;
            } else {
#ifdef MCB_EXP_SCALAR
              const double d = tr[a] - tc[b];
              kv = ev_exp(th1 - hl * (d * d));
#else
              kv = kr[b];
#endif
              v = kv;
            }
          } else if (r < ntot && c < n) {
            v = (r == n) ? s.y[c] : s.c1[c];  // border rows (lower triangle: r >= c)
          } else if (c < ntot && r < n) {
            v = (c == n) ? s.y[r] : s.c1[r];  // upper part of a diagonal block
          }
        }
        blk[e][a][b] = v;
        k0[e][a][b] = kv;
      }
    }
  }
  MCB_STAMP(0);

  // ---- B: sweep -----------------------------------------------------------------------------------
  // col + (k & 1) CS holds column k of the current symmetric matrix (all rows) and 1 / pivot at [NTOT4]
  tsync<TT>();  // previous call's readers of col / piv
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (bJ[e] == 0) {
      double* buf = s.col;
      *reinterpret_cast<double2*>(buf + 4 * bI[e]) = make_double2(blk[e][0][0], blk[e][1][0]);
      *reinterpret_cast<double2*>(buf + 4 * bI[e] + 2) = make_double2(blk[e][2][0], blk[e][3][0]);
      if (bI[e] == 0) {
        buf[NTOT4] = fast_rcp(blk[e][0][0]);
        s.piv[0] = blk[e][0][0];
      }
    }
  }
  tsync<TT>();
  const int npb = (n + 3) >> 2;  // pivot blocks
  for (int kb = 0; kb < npb; ++kb) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * kb + kk;
      if (k >= n) break;
      const int kkn = (kk + 1) & 3;
      const int kbn = (kk == 3) ? kb + 1 : kb;
      const double* buf = s.col + (kk & 1) * CS;        // k & 1 == kk & 1
      double* nbuf = s.col + ((kk + 1) & 1) * CS;
      const double inv_d = buf[NTOT4];
      const bool more = (k + 1 < n);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (bI[e] < 0) continue;
        double ci[4], tj[4];
        {
          const double2 c01 = *reinterpret_cast<const double2*>(buf + 4 * bI[e]);
          const double2 c23 = *reinterpret_cast<const double2*>(buf + 4 * bI[e] + 2);
          const double2 t01 = *reinterpret_cast<const double2*>(buf + 4 * bJ[e]);
          const double2 t23 = *reinterpret_cast<const double2*>(buf + 4 * bJ[e] + 2);
          ci[0] = c01.x; ci[1] = c01.y; ci[2] = c23.x; ci[3] = c23.y;
          tj[0] = t01.x * inv_d; tj[1] = t01.y * inv_d; tj[2] = t23.x * inv_d; tj[3] = t23.y * inv_d;
        }
        // look-ahead: my entries of pivot column k + 1, final values of this step, published before the bulk update
        if (more) {
          if (bJ[e] == kbn) {
            double v[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) v[a] = fma(-ci[a], tj[kkn], blk[e][a][kkn]);
            if (kk != 3 && bI[e] == kb) v[kk] = tj[kkn];  // diagonal block: row kk becomes the scaled pivot row
            *reinterpret_cast<double2*>(nbuf + 4 * bI[e]) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(nbuf + 4 * bI[e] + 2) = make_double2(v[2], v[3]);
            if (bI[e] == kbn) {
              const double d = v[kkn];
              nbuf[NTOT4] = fast_rcp(d);
              s.piv[k + 1] = d;
            }
          } else if (bI[e] == kbn) {
            double v[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) v[b] = fma(-ci[kkn], tj[b], blk[e][kkn][b]);
            if (kk == 3 && bJ[e] == kb) v[3] = ci[kkn] * inv_d;  // column 3 of this block is the pivot column
            *reinterpret_cast<double2*>(nbuf + 4 * bJ[e]) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(nbuf + 4 * bJ[e] + 2) = make_double2(v[2], v[3]);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) blk[e][a][b] = fma(-ci[a], tj[b], blk[e][a][b]);
        if (bJ[e] == kb) {
#pragma unroll
          for (int a = 0; a < 4; ++a) blk[e][a][kk] = ci[a] * inv_d;
        }
        if (bI[e] == kb) {
#pragma unroll
          for (int b = 0; b < 4; ++b) blk[e][kk][b] = tj[b];
          if (bJ[e] == kb) blk[e][kk][kk] = -inv_d;
        }
      }
      tsync<TT>();
    }
  }
  MCB_STAMP(1);

  // ---- C: W (lower blocks, row-wise 16-byte stores), p, q to shared memory; block-layout partial sums -----
  // borders after the sweep: A[n][c] = p_c = (W y)_c, A[n+1][c] = q_c = (W c1)_c, A[n][n] = -y'Wy, A[n+1][n] = -y'Wc1
  double* pv = s.col;       // p
  double* uv = s.col + CS;  // q, then u = p - 2 q
  double v13[13];
#pragma unroll
  for (int q = 0; q < 13; ++q) v13[q] = 0.0;
  // 0: sum W o S   1: sum W o d1   2: sum W o d2   3: tr W   4: p'd1 u   5: p'd2 u   6: p'u   7: logdet
  // 8: tr H        9: sum H o d1   10: sum H o d2
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (bI[e] < 0) continue;
    const int r0 = 4 * bI[e], c0 = 4 * bJ[e];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = r0 + a;
      if (r < n) {
        if (c0 < n) {
          double w[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) w[b] = (c0 + b < n) ? -blk[e][a][b] : 0.0;
          *reinterpret_cast<double2*>(s.W + r * LDM + c0) = make_double2(w[0], w[1]);
          *reinterpret_cast<double2*>(s.W + r * LDM + c0 + 2) = make_double2(w[2], w[3]);
        }
      } else if (r == n) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (c0 + b < n) pv[c0 + b] = blk[e][a][b];
          else if (c0 + b == n) s.corner[0] = -blk[e][a][b];
        }
      } else if (r == n + 1) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (c0 + b < n) uv[c0 + b] = blk[e][a][b];
          else if (c0 + b == n) s.corner[1] = -blk[e][a][b];
        }
      }
    }
  }
  tsync<TT>();
  {
    // mirror the strict lower triangle, one 8 x 8 tile per warp and round
    const int nt = (n + 7) >> 3;
    const int gr = lane >> 2, gc = lane & 3;
    int u = 0;
    for (int ti = 0; ti < nt; ++ti) {
      for (int tj = 0; tj <= ti; ++tj, ++u) {
        if ((u % NW) != warp) continue;
        const int r = 8 * ti + gr, c = 8 * tj + 2 * gc;
        if (r < n && c < r) {
          const double2 x = *reinterpret_cast<const double2*>(s.W + r * LDM + c);
          s.W[c * LDM + r] = x.x;
          if (c + 1 < r) s.W[(c + 1) * LDM + r] = x.y;
        }
      }
    }
    for (int a = tid; a < n; a += TT) uv[a] = pv[a] - 2.0 * uv[a];
  }
  tsync<TT>();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (bI[e] < 0) continue;
    const bool diag = (bI[e] == bJ[e]);
    const double wgt = diag ? 1.0 : 2.0;
    const int r0 = 4 * bI[e], c0 = 4 * bJ[e];
    if (c0 >= n) continue;
    double tcv[4], pc[4], uc[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const bool ok = c0 + b < n;
      tcv[b] = ok ? s.t[c0 + b] : 0.0;
      pc[b] = ok ? pv[c0 + b] : 0.0;
      uc[b] = ok ? uv[c0 + b] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = r0 + a;
      if (r >= n) continue;
      const double2 s01 = *reinterpret_cast<const double2*>(s.S + r * LDM + c0);
      const double2 s23 = *reinterpret_cast<const double2*>(s.S + r * LDM + c0 + 2);
      const double sv[4] = {s01.x, s01.y, s23.x, s23.y};
      const double trv = s.t[r], pr = pv[r], ur = uv[r];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = c0 + b;
        if (c < n) {
          const double w = -blk[e][a][b];
          v13[0] += wgt * w * sv[b];
          if (want_grad) {
            const double kv = k0[e][a][b];
            const double d = trv - tcv[b];
            const double k2 = hl * (d * d) * kv;  // dPsi/dth2 = g K0, zero on the diagonal
            v13[1] += wgt * w * kv;
            v13[2] += wgt * w * k2;
            const double pu = diag ? pr * uc[b] : pr * uc[b] + pc[b] * ur;
            v13[4] += pu * kv;
            v13[5] += pu * k2;
            if (r == c) {
              v13[3] += w;
              v13[6] += pr * ur;
            }
          }
        }
      }
    }
  }
  for (int k = tid; k < n; k += TT) {
    const double d = s.piv[k];
    v13[7] += ev_log(d);
    if (!(d > 0.0)) *bad = 1;
  }
  MCB_STAMP(2);

  if (want_grad) {
    // ---- D: per 8-row tile, X = W S then H = X W (lower tile triangle) reduced on the fly ----------------
    // fragment pair m of a row: columns 8 m + 2 gc, 8 m + 2 gc + 1 (k steps 2 m and 2 m + 1 of the permuted contraction)
    const int nt = (n + 7) >> 3;
    const int gr = lane >> 2, gc = lane & 3;
    const int role = (warp + rot) % NW;
    const double* Sf = s.S + gr * LDM + 2 * gc;
    const double* Wf = s.W + gr * LDM + 2 * gc;
    // tiles in descending cost, dealt to the roles in snake order
    for (int sidx = 0; sidx < nt; ++sidx) {
      const int rnd = sidx / NW, pos = sidx - rnd * NW;
      const int owner = (rnd & 1) ? (NW - 1 - pos) : pos;
      if (owner != role) continue;
      const int ti = nt - 1 - sidx;
      // Both products are ROLLED over the contraction / the column tile (the kernel is instruction-fetch bound: a
      // 20-instruction body that stays in the L0 instruction cache beats 150 straight-line instructions)
      double2 acc[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j] = make_double2(0.0, 0.0);
#pragma unroll 1
      for (int m = 0; m < NT; ++m) {
        const double2 af = *reinterpret_cast<const double2*>(Wf + 8 * ti * LDM + 8 * m);
        const double* sp = Sf + 8 * m;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double2 bf = *reinterpret_cast<const double2*>(sp + 8 * j * LDM);  // S symmetric
          dmma884_i(acc[j].x, acc[j].y, af.x, bf.x);
          dmma884_i(acc[j].x, acc[j].y, af.y, bf.y);
        }
      }
      // acc[j] = X[row gr][8 j + 2 gc, + 1] is fragment pair j of X: feed it straight back as the A operand
      const int r = 8 * ti + gr;
      const double trv = (r < n) ? s.t[r] : 0.0;
#pragma unroll 1
      for (int tj = 0; tj <= ti; ++tj) {
        double2 h = make_double2(0.0, 0.0);
        double2 h2 = make_double2(0.0, 0.0);
        const double* wp = Wf + 8 * tj * LDM;
#pragma unroll
        for (int m = 0; m < NT; ++m) {
          const double2 bf = *reinterpret_cast<const double2*>(wp + 8 * m);  // W symmetric
          dmma884_i(h.x, h.y, acc[m].x, bf.x);
          dmma884_i(h2.x, h2.y, acc[m].y, bf.y);
        }
        const double wgt = (tj == ti) ? 1.0 : 2.0;
        const double hq[2] = {h.x + h2.x, h.y + h2.y};
#ifndef MCB_EXP_SCALAR
        {
          const int c0 = 8 * tj + 2 * gc;
          const double d0 = trv - ((c0 < n) ? s.t[c0] : 0.0), d1 = trv - ((c0 + 1 < n) ? s.t[c0 + 1] : 0.0);
          const double gq[2] = {hl * (d0 * d0), hl * (d1 * d1)};
          const double2 k2 = ev_exp2(make_double2(th1 - gq[0], th1 - gq[1]));
          const double kq[2] = {k2.x, k2.y};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + q;
            if (r < n && c < n) {
              const double kv = (r == c) ? e1 : kq[q];
              v13[9] += wgt * hq[q] * kv;
              v13[10] += wgt * hq[q] * (gq[q] * kv);
              if (r == c) v13[8] += hq[q];
            }
          }
        }
#else
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 8 * tj + 2 * gc + q;
          if (r < n && c < n) {
            const double d = trv - s.t[c];
            const double g2 = hl * (d * d);
            const double kv = (r == c) ? e1 : ev_exp(th1 - g2);
            v13[9] += wgt * hq[q] * kv;
            v13[10] += wgt * hq[q] * (g2 * kv);
            if (r == c) v13[8] += hq[q];
          }
        }
#endif
      }
    }
  }
  MCB_STAMP(3);
  MCB_STAMP(4);
  cta_sum_v<13, TT>(v13, s.red);
  MCB_STAMP(5);
  if (tid == 0) {
    const double yWy = s.corner[0], yWc = s.corner[1];
    out[0] = 0.5 * ((double)n * kLog2Pi + v13[7] + yWy) - yWc + 0.5 * v13[0];
    if (want_grad) {
      out[1] = -0.5 * (v13[4] + v13[9] - v13[1]);
      out[2] = -0.5 * (v13[5] + v13[10] - v13[2]);
      out[3] = -sg * (v13[6] + v13[8] - v13[3]);
    }
  }
  tsync<TT>();
}

}  // namespace mcb
