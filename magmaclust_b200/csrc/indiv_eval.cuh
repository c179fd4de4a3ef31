// Objective + gradient of one individual's GP at a trial theta, whole CTA (128 threads) cooperating:
//   F_i(theta)  = logL_clust_multi_GP   (CF.R:218-248)
//   dF_i/dtheta = gr_clust_multi_GP     (CF.R:477-508)
//
// Layout and phases (n = n_i observations, ntot = n + 2 with the borders y and corr1):
//   A  build   each thread owns 4 x 4 blocks of the LOWER block triangle of the bordered matrix
//              [[Psi, y, c1], [., 0]] in REGISTERS (E blocks per thread); the SE kernel values K0 (= dPsi/dth1)
//              are also kept per block and mirrored to shared memory for phase D
//   B  sweep   n symmetric sweep steps; per step the pivot column travels through a double-buffered shared
//              vector (one barrier per step), each thread does a 4 x 4 rank-1 update (16 DFMA per 9 loads)
//              -> blocks hold -Psi^-1 (=: -W), borders Psi^-1 y, Psi^-1 c1, corner -y'Wy; log|Psi| = sum log pivots
//   C  X = W S       DMMA m8n8k4 from shared memory (S = corr2 stays resident in shared memory across calls)
//   D  H = X W       DMMA on the lower tile triangle, never stored: its accumulators are reduced on the fly
//              against dPsi/dth1, dPsi/dth2 and I (the trace-by-product of CF.R:499-504)
//   E  block reduction of 13 partial sums -> f, g
// Flops as executed ~ n^3/2 (sweep) + n^3 (X) + n^3/2 (H) FMAs vs the reference's n^3 + 4 n^3 + 4 n^3 flops.
#pragma once
#include "common.cuh"

#ifdef MCB_PROF
// phase profile of indiv_eval_v1 (tools/dbg/eval_prof.cu): thread 0 accumulates clock64 deltas per phase
__device__ long long g_eval_prof[16];
#define MCB_STAMP(k)                                    \
  do {                                                  \
    if (threadIdx.x == 0 && blockIdx.x == 0) {          \
      const long long now_ = clock64();                 \
      g_eval_prof[k] += now_ - prof_t_;                 \
      prof_t_ = now_;                                   \
    }                                                   \
  } while (0)
#else
#define MCB_STAMP(k) do {} while (0)
#endif

namespace mcb {

constexpr int kT = 128;  // threads per CTA of the E-step kernels (indiv.cu)

// barrier over the TT threads that share one individual: a single warp needs no CTA barrier
template <int TT>
__device__ __forceinline__ void tsync() {
  if (TT == 32) __syncwarp();
  else __syncthreads();
}

__device__ __forceinline__ void dmma884_i(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct EvalSm {
  double *W, *S, *X, *Kd;  // np8 x ldm each
  double *col[2];          // pivot column broadcast, ntot4 + 2 each (last entries: 1/pivot)
  double *piv;             // ntot4
  double *t, *y, *c1;      // nmax each
  double *red;             // 16 * 4
  int np8, ldm, ntot4;
};

__host__ __device__ inline int eval_np8(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int eval_ldm(int n) { return eval_np8(n) + 4; }      // ldm mod 16 in {4, 12}
__host__ __device__ inline int eval_ntot4(int n) { return (n + 2 + 3) & ~3; }

inline size_t eval_smem_doubles(int nmax) {
  const size_t mat = (size_t)eval_np8(nmax) * eval_ldm(nmax);
  return 4 * mat + 2 * (size_t)(eval_ntot4(nmax) + 2) + eval_ntot4(nmax) + 3 * (size_t)nmax + 16 * 4 + 8;
}

__device__ __forceinline__ EvalSm eval_carve(double* sm, int nmax) {
  EvalSm s;
  s.np8 = eval_np8(nmax);
  s.ldm = eval_ldm(nmax);
  s.ntot4 = eval_ntot4(nmax);
  const size_t mat = (size_t)s.np8 * s.ldm;
  s.W = sm;
  s.S = s.W + mat;
  s.X = s.S + mat;
  s.Kd = s.X + mat;
  s.col[0] = s.Kd + mat;
  s.col[1] = s.col[0] + s.ntot4 + 2;
  s.piv = s.col[1] + s.ntot4 + 2;
  s.t = s.piv + s.ntot4;
  s.y = s.t + nmax;
  s.c1 = s.y + nmax;
  s.red = s.c1 + nmax;
  return s;
}

// number of 4 x 4 blocks of the lower block triangle for n observations (+ 2 border rows)
inline int eval_num_blocks(int nmax) {
  const int nbk = eval_ntot4(nmax) / 4;
  return nbk * (nbk + 1) / 2;
}

template <int NV, int TT>
__device__ __forceinline__ void cta_sum_v(double (&v)[NV], double* red) {
  constexpr int NW = TT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if defined(MCB_SUM_TRANSPOSE)
  // Experiment (compile-time, default off: built and counted in round 1, not yet timed — profiles/r01f_ptxas.txt): the NV
  // butterflies of the default path are 5 NV 64-bit shuffle steps, fully unrolled (951 SASS instructions at NV = 13, the
  // largest single block of the evaluator). A transpose-reduce halves the number of live slots per step instead: 8 + 4 +
  // 2 + 1 exchanges with the partners at distance 16, 8, 4, 2, then one ordinary step at distance 1: 16 shuffle steps, and
  // lane l ends up with the warp total of slot l >> 1. Only the summation order differs from the default path.
  static_assert(NV <= 16, "transpose-reduce holds 16 slots");
  double w[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) w[q] = q < NV ? v[q] : 0.0;
#pragma unroll
  for (int off = 16, half = 8; off >= 2; off >>= 1, half >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? w[i] : w[i + half];   // the half the partner keeps
      const double keep = upper ? w[i + half] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
  if (NW == 1) {
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = __shfl_sync(0xffffffffu, w[0], 2 * q);
    return;
  }
  __syncthreads();
  if (!(lane & 1) && (lane >> 1) < NV) red[(lane >> 1) * NW + warp] = w[0];
  __syncthreads();
#else
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  if (NW == 1) return;
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[q * NW + warp] = v[q];
  }
  __syncthreads();
#endif
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double t = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) t += red[q * NW + w2];
    v[q] = t;
  }
}

// One-time per individual: zero the matrix buffers, load t / y / c1 and S = corr2 (n x n, row-major in global).
template <int TT>
__device__ __forceinline__ void eval_load(const EvalSm& s, int n, const double* __restrict__ tg,
                                          const double* __restrict__ yg, const double* __restrict__ c1g,
                                          const double* __restrict__ c2g) {
  const size_t mat = (size_t)s.np8 * s.ldm;
  for (size_t e = threadIdx.x; e < 4 * mat; e += TT) s.W[e] = 0.0;
  for (int a = threadIdx.x; a < n; a += TT) {
    s.t[a] = tg[a];
    s.y[a] = yg[a];
    s.c1[a] = c1g[a];
  }
  tsync<TT>();
  for (int e = threadIdx.x; e < n * n; e += TT) {
    const int a = e / n, b = e - a * n;
    s.S[a * s.ldm + b] = c2g[e];
  }
  tsync<TT>();
}

// out[0] = f, out[1..3] = g (shared memory, written by thread 0). All 128 threads must call.
template <int E, int TT>
__device__ void indiv_eval_v1(const EvalSm& s, int n, double th1, double th2, double sg, bool want_grad,
                              double* out, int* bad) {
  constexpr int NW = TT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntot = n + 2;
  const int nbk = (ntot + 3) >> 2;
  const int nblk = nbk * (nbk + 1) / 2;
  const int ldm = s.ldm;
  const double hl = exp(-th2) / 2.0;
  const double e1 = exp(th1);
  const double dg = e1 + sg * sg;

#ifdef MCB_PROF
  long long prof_t_ = clock64();
#endif
  int bI[E], bJ[E];
  double blk[E][4][4], k0[E][4][4];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int q = tid + TT * e;
    if (q < nblk) {
      int I = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
      while ((I + 1) * (I + 2) / 2 <= q) ++I;
      while (I * (I + 1) / 2 > q) --I;
      bI[e] = I;
      bJ[e] = q - I * (I + 1) / 2;
    } else {
      bI[e] = -1;
      bJ[e] = -1;
    }
  }

  // ---- A: build ---------------------------------------------------------------------------------
#pragma unroll
  for (int e = 0; e < E; ++e) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
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
              const double d = s.t[r] - s.t[c];
              kv = exp(th1 - hl * (d * d));
              v = kv;
            }
            s.Kd[r * ldm + c] = kv;
            s.Kd[c * ldm + r] = kv;
          } else if (r < ntot && c < n) {
            v = (r == n) ? s.y[c] : s.c1[c];        // border rows (lower triangle: r >= c)
          } else if (c < ntot && r < n) {
            v = (c == n) ? s.y[r] : s.c1[r];        // upper part of a diagonal block
          }
        }
        blk[e][a][b] = v;
        k0[e][a][b] = kv;
      }
    }
  }

  MCB_STAMP(0);
  // ---- B: sweep -----------------------------------------------------------------------------------
  auto publish = [&](int kb, int kk, double* buf) {
    // column k = 4 kb + kk of the current symmetric matrix -> buf[0..ntot4), buf[ntot4] = 1 / pivot
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (bJ[e] == kb) {
        double v[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          v[a] = blk[e][a][0];
          if (kk == 1) v[a] = blk[e][a][1];
          if (kk == 2) v[a] = blk[e][a][2];
          if (kk == 3) v[a] = blk[e][a][3];
        }
        // two 16-byte stores instead of four 8-byte ones (4 I is a multiple of 4 doubles)
        *reinterpret_cast<double2*>(buf + 4 * bI[e]) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(buf + 4 * bI[e] + 2) = make_double2(v[2], v[3]);
        if (bI[e] == kb) {
          double d = blk[e][0][0];
          if (kk == 1) d = blk[e][1][1];
          if (kk == 2) d = blk[e][2][2];
          if (kk == 3) d = blk[e][3][3];
          buf[s.ntot4] = 1.0 / d;
          s.piv[4 * kb + kk] = d;
        }
      } else if (bI[e] == kb) {
        double v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          v[b] = blk[e][0][b];
          if (kk == 1) v[b] = blk[e][1][b];
          if (kk == 2) v[b] = blk[e][2][b];
          if (kk == 3) v[b] = blk[e][3][b];
        }
        *reinterpret_cast<double2*>(buf + 4 * bJ[e]) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(buf + 4 * bJ[e] + 2) = make_double2(v[2], v[3]);
      }
    }
  };

  tsync<TT>();  // Kd / previous call's readers
  publish(0, 0, s.col[0]);
  tsync<TT>();
  const int npb = (n + 3) >> 2;  // pivot blocks
  for (int kb = 0; kb < npb; ++kb) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * kb + kk;
      if (k >= n) break;
      const double* buf = s.col[k & 1];
      const double inv_d = buf[s.ntot4];
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
      if (k + 1 < n) publish(kk == 3 ? kb + 1 : kb, (kk + 1) & 3, s.col[(k + 1) & 1]);
      tsync<TT>();
    }
  }

  MCB_STAMP(1);
  // ---- W, p, q to shared memory; block-layout partial sums ----------------------------------------------
  // borders after the sweep: A[r][n] = p_r = (W y)_r, A[r][n+1] = q_r = (W c1)_r, A[n][n] = -y'Wy, A[n][n+1] = -y'Wc1
  double* pv = s.col[0];            // p
  double* uv = s.col[1];            // u = p - 2 q
  double* corner = s.piv + s.ntot4 - 2;  // not used by piv (ntot4 >= n + 2): [0] = y'Wy, [1] = y'Wc1
  // (piv[k] for k < n stay intact: corner sits at indices >= ntot4 - 2 >= n)
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (bI[e] < 0) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * bI[e] + a, c = 4 * bJ[e] + b;
        const double v = blk[e][a][b];
        if (r < n && c < n) {
          s.W[r * ldm + c] = -v;
          s.W[c * ldm + r] = -v;
        } else if (c < n && r == n) {
          pv[c] = v;
        } else if (c < n && r == n + 1) {
          s.X[c] = v;  // q parked in X's first row until u is formed below
        } else if (r == n && c == n) {
          corner[0] = -v;
        } else if (r == n + 1 && c == n) {
          corner[1] = -v;
        }
      }
    }
  }
  tsync<TT>();
  for (int a = tid; a < n; a += TT) uv[a] = pv[a] - 2.0 * s.X[a];
  tsync<TT>();
  double v13[13];
#pragma unroll
  for (int q = 0; q < 13; ++q) v13[q] = 0.0;
  // 0: sum W o S   1: sum W o d1   2: sum W o d2   3: tr W   4: p'd1 u   5: p'd2 u   6: p'u   7: logdet
  // 8: tr H        9: sum H o d1   10: sum H o d2
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (bI[e] < 0) continue;
    const bool diag = (bI[e] == bJ[e]);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * bI[e] + a, c = 4 * bJ[e] + b;
        if (r < n && c < n) {
          const double w = -blk[e][a][b];
          const double kv = k0[e][a][b];
          const double d = s.t[r] - s.t[c];
          const double k2 = hl * (d * d) * kv;  // dPsi/dth2 = g K0, zero on the diagonal
          const double wgt = diag ? 1.0 : 2.0;
          v13[0] += wgt * w * s.S[r * ldm + c];
          if (want_grad) {
            v13[1] += wgt * w * kv;
            v13[2] += wgt * w * k2;
            const double pu = diag ? pv[r] * uv[c] : pv[r] * uv[c] + pv[c] * uv[r];
            v13[4] += pu * kv;
            v13[5] += pu * k2;
            if (r == c) {
              v13[3] += w;
              v13[6] += pv[r] * uv[r];
            }
          }
        }
      }
    }
  }
  for (int k = tid; k < n; k += TT) {
    const double d = s.piv[k];
    v13[7] += log(d);
    if (!(d > 0.0)) *bad = 1;
  }

  MCB_STAMP(2);
  if (want_grad) {
    // ---- C: X = W S (np8 x np8, DMMA) -------------------------------------------------------------
    const int nt = (n + 7) >> 3;   // 8-wide tiles
    const int ks = (n + 3) >> 2;   // 4-deep k steps (pads are zero)
    const int gr = lane >> 2, gc = lane & 3;
    {
      const int ncu = (nt + 3) >> 2;             // column chunks of <= 4 tiles
      for (int u = warp; u < nt * ncu; u += NW) {
        const int ti = u / ncu, cj = (u - ti * ncu) * 4;
        double acc[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
        const double* wr = s.W + (8 * ti + gr) * ldm + gc;
        for (int k = 0; k < ks; ++k) {
          const double af = wr[4 * k];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (cj + j < nt) {
              const double bf = s.S[(8 * (cj + j) + gr) * ldm + 4 * k + gc];  // S symmetric
              dmma884_i(acc[j][0], acc[j][1], af, bf);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cj + j < nt) {
            double* xo = s.X + (8 * ti + gr) * ldm + 8 * (cj + j) + 2 * gc;
            xo[0] = acc[j][0];
            xo[1] = acc[j][1];
          }
        }
      }
    }
    tsync<TT>();
    MCB_STAMP(3);
    // ---- D: H = X W on the lower tile triangle, reduced on the fly -----------------------------------
    {
      int u = 0;
      for (int ti = 0; ti < nt; ++ti) {
        for (int cj = 0; cj <= ti; cj += 4, ++u) {
          if ((u % NW) != warp) continue;
          double acc[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
          const double* xr = s.X + (8 * ti + gr) * ldm + gc;
          for (int k = 0; k < ks; ++k) {
            const double af = xr[4 * k];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (cj + j <= ti) {
                const double bf = s.W[(8 * (cj + j) + gr) * ldm + 4 * k + gc];  // W symmetric
                dmma884_i(acc[j][0], acc[j][1], af, bf);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (cj + j <= ti) {
              const double wgt = (cj + j == ti) ? 1.0 : 2.0;
              const int r = 8 * ti + gr;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int c = 8 * (cj + j) + 2 * gc + q;
                if (r < n && c < n) {
                  const double h = acc[j][q];
                  const double kv = s.Kd[r * ldm + c];
                  const double d = s.t[r] - s.t[c];
                  v13[9] += wgt * h * kv;
                  v13[10] += wgt * h * (hl * (d * d) * kv);
                  if (r == c) v13[8] += h;
                }
              }
            }
          }
        }
      }
    }
  }
  MCB_STAMP(4);
  cta_sum_v<13, TT>(v13, s.red);
  MCB_STAMP(5);
  if (tid == 0) {
    const double yWy = corner[0], yWc = corner[1];
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
