// Objective + gradient of one individual's GP at a trial theta, third design: everything on DMMA 8 x 8 tiles.
//   F_i(theta)  = logL_clust_multi_GP   (CF.R:218-248)
//   dF_i/dtheta = gr_clust_multi_GP     (CF.R:477-508)
//
// The second design (indiv_eval2.cuh) still swept the bordered matrix [[Psi, y, c1], [., 0]] with one rank-1 update per
// pivot on 4 x 4 register blocks: 52 CTA barriers and ~85 warp instructions per pivot for 16 FMAs per thread, 44 % of
// all instructions of an evaluation (ncu source page of tools/dbg/eval_prof.cu). Here the sweep is BLOCKED, 8 pivots
// at a time, in the Cholesky form that the N x N slab inverse uses (dense_slab.cu): for pivot tile b with D = A[b,b]
//   D = L L'   (one warp, all 32 lanes redundantly in registers, no communication)      -> L^-1 to shared memory
//   V = A[:,b] L^-T,  A[i,j] -= V_i V_j'   (2 DMMA per 8 x 8 tile),   A[:,b] = V L^-1,   A[b,b] = -L^-T L^-1
// The lower tile triangle lives in the accumulator registers of the warp that owns the tile row, for the whole
// sweep; with the permuted contraction index of indiv_eval2.cuh the accumulator pair of V IS the operand pair of the
// update, so V never leaves registers either. Two CTA barriers per 8 pivots. A partial last pivot tile (n not a
// multiple of 8; the border rows y, c1 sit in it) is handled by identity padding of D and by keeping the regular
// update for its non-pivot rows and columns.
//
// Tile rows are dealt to the warps in snake order by cost; the same warp owns the same rows in the sweep, in the
// post phase and in the tile phase (X = W S, H = X W reduced on the fly), where dPsi/dtheta is recomputed per element and
// every gradient sum that needs it is taken (no second copy of the kernel matrix anywhere).
#pragma once
#include "common.cuh"
#include "indiv_eval2.cuh"  // Eval2Geo, EvalSm2, eval2_carve, eval2_load (same shared-memory layout), fast_rcp

namespace mcb {

// 1 / d: hardware seed (~2^-20) and one cubic (Halley) step, r (1 + e + e^2) with e = 1 - d r
__device__ __forceinline__ double ev3_rcp3(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// sqrt(r) for r = 1 / d > 0 (= 1 / sqrt(d))
__device__ __forceinline__ double ev3_rsqrt_of_rcp(double r) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r));  // ~ 1 / sqrt(r)
  const double hr = 0.5 * r;
  y = fma(y, fma(-hr * y, y, 0.5), y);
  y = fma(y, fma(-hr * y, y, 0.5), y);
  const double g = r * y;
  return fma(fma(-g, g, r), 0.5 * y, g);
}

// L^-1 of the leading p x p part of the 8 x 8 tile Dp (row stride 8 doubles), identity beyond p. Every lane of the
// calling warp works on the whole tile in registers (same scheme as factor32.cuh). Ls (8 x 8, zero above the
// diagonal on entry and on exit) receives L^-1, piv[0..p) the pivots of the LDL' elimination.
__device__ __noinline__ void ev3_factor8(const double* __restrict__ Dp, int p, double* __restrict__ Ls,
                                         double* __restrict__ piv) {
  double a[8][8], w[8][8], sk[8], dk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[i][j] = (i < p) ? Dp[i * 8 + j] : (i == j ? 1.0 : 0.0);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double d = a[k][k];
    dk[k] = d;
    const double r = ev3_rcp3(d);
    sk[k] = ev3_rsqrt_of_rcp(r);
    if (k + 1 < 8) {
#pragma unroll
      for (int i = k + 1; i < 8; ++i) a[i][k + 1] = fma(-(a[i][k] * a[k + 1][k]), r, a[i][k + 1]);
    }
#pragma unroll
    for (int i = k + 1; i < 8; ++i) {
      const double m = a[i][k] * r;
#pragma unroll
      for (int j = k + 2; j <= i; ++j) a[i][j] = fma(-m, a[j][k], a[i][j]);
      w[i][k] = -m;
#pragma unroll
      for (int c = 0; c < k; ++c) w[i][c] = fma(-m, w[k][c], w[i][c]);
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int c = 0; c < i; ++c) Ls[i * 8 + c] = w[i][c] * sk[i];
      Ls[i * 8 + i] = sk[i];
      if (i < p) piv[i] = dk[i];
    }
  }
}

// out[0] = f, out[1..3] = g (shared memory, written by thread 0). All TT threads must call. Needs 8 NT >= n + 2.
// Shared memory: the layout of EvalSm2<NT>; Ls = s.red + 64 .. (64 doubles, zero on entry: see eval3_load).
template <int TT, int NT>
__device__ void indiv_eval_v3(const EvalSm2<NT>& s, int n, double th1, double th2, double sg, bool want_grad,
                              double* out, int* bad, int rot) {
  using G = Eval2Geo<NT>;
  constexpr int NW = TT / 32;
  constexpr int ROUNDS = (NT + NW - 1) / NW;
  constexpr int LDM = G::LDM;
  constexpr int CS = G::CS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gr = lane >> 2, gc = lane & 3;
  const int ntb = (n + 2 + 7) >> 3;   // tile rows of the bordered matrix
  const int nt = (n + 7) >> 3;        // tile rows of Psi
  const int role = (warp + rot) % NW;
  const double hl = ev_exp(-th2) * 0.5;
  const double e1 = ev_exp(th1);
  const double dg = e1 + sg * sg;
#ifdef MCB_PROF
  long long prof_t_ = clock64();
#endif
  double* Pn = s.W;            // two panel buffers of NP8 x 8 (the W region is free until the post phase)
  double* Ls = s.red + 64;     // L^-1 of the current pivot tile
  double* pv = s.col;          // p = W y
  double* uv = s.col + CS;     // q = W c1, then u = p - 2 q

  // tile rows of this warp: round r -> row irow[r] (snake order by cost), tiles (irow[r], j), j <= irow[r] < NT - r NW
  int irow[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int pos = (r & 1) ? (NW - 1 - role) : role;
    irow[r] = NT - 1 - (r * NW + pos);   // may be negative in the last round: no row
  }
  double2 acc[ROUNDS][NT];

  // ---- A: build the lower tile triangle of [[Psi, y, c1], [., 0]] in accumulator layout -----------------------
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int i = irow[r];
    const int row = 8 * i + gr;
    const double trow = (i >= 0 && row < n) ? s.t[row] : 0.0;
#pragma unroll
    for (int j = 0; j < NT - r * NW; ++j) {
      double v[2] = {0.0, 0.0};
      if (i >= 0 && j <= i && i < ntb) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = 8 * j + 2 * gc + q;
          if (row < n && c < n) {
            if (row == c) v[q] = dg;
            else {
              const double d = trow - s.t[c];
              v[q] = ev_exp(th1 - hl * (d * d));
            }
          } else if (row < n + 2 && c < n) {
            v[q] = (row == n) ? s.y[c] : s.c1[c];       // border rows
          } else if (c < n + 2 && row < n) {
            v[q] = (c == n) ? s.y[row] : s.c1[row];     // border columns inside a diagonal tile
          }
        }
      }
      acc[r][j] = make_double2(v[0], v[1]);
    }
  }
  MCB_STAMP(0);

  // ---- B: blocked sweep -----------------------------------------------------------------------------------------
  tsync<TT>();  // previous call's readers of the W region / Ls
#pragma unroll
  for (int b = 0; b < NT; ++b) {
    const int pb = min(8, n - 8 * b);   // pivots of this tile
    if (pb > 0) {
      double* Pb = Pn + (b & 1) * (G::NP8 * 8);
      // publish the pivot panel A[:, 8b .. 8b + 8) with the non-pivot columns zeroed
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        const int i = irow[r];
        if (b < NT - r * NW && i >= b && i < ntb) {
          double2 v = acc[r][b];
          if (2 * gc >= pb) v.x = 0.0;
          if (2 * gc + 1 >= pb) v.y = 0.0;
          *reinterpret_cast<double2*>(Pb + (8 * i + gr) * 8 + 2 * gc) = v;
        }
        if (i == b) {
#pragma unroll
          for (int j = 0; j < b; ++j) {
            if (j < NT - r * NW) {
              const bool live = gr < pb;
              Pb[(8 * j + 2 * gc) * 8 + gr] = live ? acc[r][j].x : 0.0;
              Pb[(8 * j + 2 * gc + 1) * 8 + gr] = live ? acc[r][j].y : 0.0;
            }
          }
        }
      }
      tsync<TT>();
      if (warp == (b % NW)) ev3_factor8(Pb + (8 * b) * 8, pb, Ls, s.piv + 8 * b);
      tsync<TT>();
      const double2 Lf = *reinterpret_cast<const double2*>(Ls + gr * 8 + 2 * gc);     // L^-1[gr][2 gc, + 1]
      const double2 LTf = make_double2(Ls[(2 * gc) * 8 + gr], Ls[(2 * gc + 1) * 8 + gr]);  // L^-1[2 gc, + 1][gr]
      // V_j = A[j, b] L^-T for every tile row (operand of the columns), and for the rows of this warp
      double2 V[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        V[j] = make_double2(0.0, 0.0);
        if (j < ntb) {
          const double2 a = *reinterpret_cast<const double2*>(Pb + (8 * j + gr) * 8 + 2 * gc);
          dmma884_i(V[j].x, V[j].y, a.x, Lf.x);
          dmma884_i(V[j].x, V[j].y, a.y, Lf.y);
        }
      }
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        const int i = irow[r];
        if (i < 0 || i >= ntb) continue;
        double2 Vm = make_double2(0.0, 0.0);
        {
          const double2 a = *reinterpret_cast<const double2*>(Pb + (8 * i + gr) * 8 + 2 * gc);
          dmma884_i(Vm.x, Vm.y, a.x, Lf.x);
          dmma884_i(Vm.x, Vm.y, a.y, Lf.y);
        }
        const double nx = -Vm.x, ny = -Vm.y;
        // regular update of every owned tile: A[i, j] -= V_i V_j'
#pragma unroll
        for (int j = 0; j < NT - r * NW; ++j) {
          if (j <= i) {
            dmma884_i(acc[r][j].x, acc[r][j].y, nx, V[j].x);
            dmma884_i(acc[r][j].x, acc[r][j].y, ny, V[j].y);
          }
        }
        // pivot column tile (i, b): columns < pb become P = V_i L^-1
        if (b < NT - r * NW && i >= b) {
          double2 P = make_double2(0.0, 0.0);
          dmma884_i(P.x, P.y, Vm.x, LTf.x);
          dmma884_i(P.x, P.y, Vm.y, LTf.y);
          if (i > b) {
            if (2 * gc < pb) acc[r][b].x = P.x;
            if (2 * gc + 1 < pb) acc[r][b].y = P.y;
          } else {
            // diagonal tile: [pivot, pivot] = -D^-1 = -L^-T L^-1; [other, pivot] = P; [pivot, other] = P'; rest updated
            double2 Di = make_double2(0.0, 0.0), PT = make_double2(0.0, 0.0);
            dmma884_i(Di.x, Di.y, LTf.x, LTf.x);
            dmma884_i(Di.x, Di.y, LTf.y, LTf.y);
            dmma884_i(PT.x, PT.y, LTf.x, Vm.x);
            dmma884_i(PT.x, PT.y, LTf.y, Vm.y);
            const bool rp = gr < pb;
            if (2 * gc < pb) acc[r][b].x = rp ? -Di.x : P.x;
            else if (rp) acc[r][b].x = PT.x;
            if (2 * gc + 1 < pb) acc[r][b].y = rp ? -Di.y : P.y;
            else if (rp) acc[r][b].y = PT.y;
          }
        }
        // pivot row tiles (b, j), j < b: rows < pb become P_j' = L^-T V_j'
        if (i == b) {
#pragma unroll
          for (int j = 0; j < b; ++j) {
            if (j < NT - r * NW) {
              double2 PT = make_double2(0.0, 0.0);
              dmma884_i(PT.x, PT.y, LTf.x, V[j].x);
              dmma884_i(PT.x, PT.y, LTf.y, V[j].y);
              if (gr < pb) acc[r][j] = PT;
            }
          }
        }
      }
    }
  }
  MCB_STAMP(1);

  // ---- C: -A -> W in shared memory (both triangles), p, q, corner; sum W o S --------------------------------------
  tsync<TT>();  // the panel buffers alias the W region
  double v13[13];
#pragma unroll
  for (int q = 0; q < 13; ++q) v13[q] = 0.0;
  // 0: sum W o S   1: sum W o d1   2: sum W o d2   3: tr W   4: p'd1 u   5: p'd2 u   6: p'u   7: logdet
  // 8: tr H        9: sum H o d1   10: sum H o d2
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int i = irow[r];
    if (i < 0 || i >= ntb) continue;
    const int row = 8 * i + gr;
#pragma unroll
    for (int j = 0; j < NT - r * NW; ++j) {
      if (j > i) continue;
      const int c = 8 * j + 2 * gc;
      const double2 a = acc[r][j];
      if (row < n) {
        const double wx = (c < n) ? -a.x : 0.0, wy = (c + 1 < n) ? -a.y : 0.0;
        if (i < nt && j < nt) {
          *reinterpret_cast<double2*>(s.W + row * LDM + c) = make_double2(wx, wy);
          if (j < i) {
            if (c < n) s.W[c * LDM + row] = wx;
            if (c + 1 < n) s.W[(c + 1) * LDM + row] = wy;
          }
          const double2 sv = *reinterpret_cast<const double2*>(s.S + row * LDM + c);
          const double wgt = (j == i) ? 1.0 : 2.0;
          v13[0] += wgt * (wx * sv.x + wy * sv.y);
        }
      } else if (row == n) {
        if (c < n) pv[c] = a.x; else if (c == n) s.corner[0] = -a.x;
        if (c + 1 < n) pv[c + 1] = a.y; else if (c + 1 == n) s.corner[0] = -a.y;
      } else if (row == n + 1) {
        if (c < n) uv[c] = a.x; else if (c == n) s.corner[1] = -a.x;
        if (c + 1 < n) uv[c + 1] = a.y; else if (c + 1 == n) s.corner[1] = -a.y;
      }
    }
  }
  for (int k = tid; k < n; k += TT) {
    const double d = s.piv[k];
    v13[7] += ev_log(d);
    if (!(d > 0.0)) *bad = 1;
  }
  tsync<TT>();
  for (int a = tid; a < n; a += TT) uv[a] = pv[a] - 2.0 * uv[a];
  tsync<TT>();
  MCB_STAMP(2);

  if (want_grad) {
    // ---- D: per 8-row tile, X = W S then H = X W (lower tile triangle); every sum against dPsi/dtheta is taken here ----
    const double* Sf = s.S + gr * LDM + 2 * gc;
    const double* Wf = s.W + gr * LDM + 2 * gc;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int ti = irow[r];
      if (ti < 0 || ti >= nt) continue;
      double2 af[NT];
#pragma unroll
      for (int m = 0; m < NT; ++m) af[m] = *reinterpret_cast<const double2*>(Wf + 8 * ti * LDM + 8 * m);
      double2 x[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) x[j] = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < NT; ++m) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double2 bf = *reinterpret_cast<const double2*>(Sf + 8 * j * LDM + 8 * m);  // S symmetric
          dmma884_i(x[j].x, x[j].y, af[m].x, bf.x);
          dmma884_i(x[j].x, x[j].y, af[m].y, bf.y);
        }
      }
      // x[j] = X[row gr][8 j + 2 gc, + 1] is fragment pair j of X: feed it straight back as the A operand
      const int row = 8 * ti + gr;
      const bool rok = row < n;
      const double trv = rok ? s.t[row] : 0.0;
      const double pr = rok ? pv[row] : 0.0, ur = rok ? uv[row] : 0.0;
#pragma unroll
      for (int tj = 0; tj < NT - r * NW; ++tj) {
        if (tj > ti) continue;
        double2 h = make_double2(0.0, 0.0), h2 = make_double2(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < NT; ++m) {
          const double2 bf = *reinterpret_cast<const double2*>(Wf + 8 * tj * LDM + 8 * m);  // W symmetric
          dmma884_i(h.x, h.y, x[m].x, bf.x);
          dmma884_i(h2.x, h2.y, x[m].y, bf.y);
        }
        const bool dtile = (tj == ti);
        const double wgt = dtile ? 1.0 : 2.0;
        const int c0 = 8 * tj + 2 * gc;
        const double2 wv = af[tj];   // W[row][c0, c0 + 1]
        const double hq[2] = {h.x + h2.x, h.y + h2.y};
        const double wq[2] = {wv.x, wv.y};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = c0 + q;
          if (rok && c < n) {
            const double d = trv - s.t[c];
            const double g2 = hl * (d * d);
            const double kv = (row == c) ? e1 : ev_exp(th1 - g2);
            const double k2 = g2 * kv;            // dPsi/dth2 = g K0, zero on the diagonal
            const double pu = dtile ? pr * uv[c] : pr * uv[c] + pv[c] * ur;
            v13[9] += wgt * hq[q] * kv;
            v13[10] += wgt * hq[q] * k2;
            v13[1] += wgt * wq[q] * kv;
            v13[2] += wgt * wq[q] * k2;
            v13[4] += pu * kv;
            v13[5] += pu * k2;
            if (row == c) {
              v13[8] += hq[q];
              v13[3] += wq[q];
              v13[6] += pr * ur;
            }
          }
        }
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
