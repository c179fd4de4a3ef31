// Latency-optimised inverse for mid-size grids (N <= 768): one persistent cooperative kernel per batch of
// matrices. The multi-launch blocked sweep of dense.cu pays ~2 launches and ~30 us per 32-wide pivot block;
// at N = 401 (BASELINE config C2) that is the whole cost of every theta_k objective evaluation. Here the matrix
// stays in SHARED memory for the entire inversion: CTA (I, q) keeps rows 32 I .. 32 I + 31 of column part q (the
// columns are cut into Q parts of whole 32-blocks: at N = 401, Q = 2 that is 26 CTAs holding 32 x 224 doubles each).
// The CTAs of one matrix exchange the 32 x 32 panels through L2 and synchronise on global-memory counters.
//
// Same algorithm as dense.cu (blocked symmetric sweep in Cholesky form). Per pivot block b (columns in part qb):
//   phase 2  CTAs (I, qb), once L^-1 of block b is published: V_I = A[I,b] L^-T, P_I = V_I L^-1 (DMMA) -> L2
//   barrier  every CTA arrives; every CTA waits, except the lookahead CTA (below), which needs no one else's V
//   phase 3  CTA (I, q), I != b, b+1: slab[:, J] -= V_I V_J' for the blocks J != b of its part (DMMA, V_J streamed
//                                     from L2), slab[:, b] = P_I;
//            CTA (b, q):              slab[:, J] = P_J', slab[:, b] = -D^-1;
//            CTA (b+1, part of block b+1) = lookahead: updates only its diagonal block, factorises it (factor32.cuh)
//                                     and publishes L^-1 for the next step; the other parts of row block b+1 idle
//                                     (their content is replaced at the next step).
// There is no barrier between phase 3 and the next phase 2: a CTA moves on as soon as the next L^-1 is published.
// The panel buffers are triple-buffered because the lookahead CTA may run one step ahead of the slowest reader.
// After the last block the slabs hold -A^-1; they are negated on the way back to global memory.
#include <cooperative_groups.h>

#include "common.cuh"
#include "factor32.cuh"

#ifdef SLAB_PROF
// tools/dbg/slab_prof.cu: cycles per segment of a pivot-block step, by role of the CTA in that step
// (0 look-ahead CTA in the pivot column part = critical path, 1 other CTAs of the pivot column part, 2 the rest)
__device__ long long g_slab_prof[3][16];
#define SP_BEGIN(cls) const int sp_cls_ = (cls); long long sp_t_ = clock64();
#define SP(k)                                                       \
  do {                                                              \
    if (threadIdx.x == 0 && blockIdx.y == 0) {                      \
      const long long now_ = clock64();                             \
      atomicAdd((unsigned long long*)&g_slab_prof[sp_cls_][k], (unsigned long long)(now_ - sp_t_)); \
      sp_t_ = now_;                                                 \
    }                                                               \
  } while (0)
#else
#define SP_BEGIN(cls)
#define SP(k) do {} while (0)
#endif

namespace mcb {

constexpr int kSlabThreads = 256;
constexpr int kLp = kNB + 4;  // 36: row stride of the 32 x 32 operand tiles (conflict-free DMMA fragments)
static_assert(kLp == kF32Ld, "factor32 works on the operand tiles in place");

__device__ __forceinline__ void dmma884_s(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Arrival on a monotonic counter: everything this CTA wrote before (bar.sync orders the other threads' writes) is
// visible to whoever later observes the arrival.
__device__ __forceinline__ void slab_arrive(unsigned* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
// Wait until the counter / flag has reached `target`; nothing after the wait is read before that is seen.
__device__ __forceinline__ void slab_wait(const unsigned* counter, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}
__device__ __forceinline__ void slab_publish(unsigned* flag, unsigned value) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// C(32x32) = A(32x32) * B'(32x32) with A, B row-major in shared memory (stride kLp); 8 warps x 2 tiles.
// UPPER / LOWER select the triangular structure of B (Linv lower): skip k-steps that are entirely zero.
__device__ __forceinline__ void mm32_nt(const double* As, const double* Bs, double* Cs, int warp, int lane) {
  const int gr = lane >> 2, gc = lane & 3;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int tile = warp * 2 + t;       // 16 tiles of 8 x 8
    const int ti = tile >> 2, tj = tile & 3;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double a = As[(8 * ti + gr) * kLp + 4 * k + gc];
      const double b = Bs[(8 * tj + gr) * kLp + 4 * k + gc];
      dmma884_s(c0, c1, a, b);
    }
    Cs[(8 * ti + gr) * kLp + 8 * tj + 2 * gc] = c0;
    Cs[(8 * ti + gr) * kLp + 8 * tj + 2 * gc + 1] = c1;
  }
}

__global__ void __launch_bounds__(kSlabThreads, 1)
spd_inverse_slab_kernel(double* __restrict__ Aall, int N, int ld, long long stride, int Q, int Wb, int lds,
                        double* __restrict__ Vbuf, double* __restrict__ Pbuf, double* __restrict__ Lbuf,
                        unsigned* __restrict__ sync, double* __restrict__ logdet, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int ncta = gridDim.x;
  const int nblk = ncta / Q;
  const int I = blockIdx.x / Q, q = blockIdx.x - I * Q, z = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gr = lane >> 2, gc = lane & 3;
  const int r0 = I * kNB;
  const int Npad = nblk * kNB;
  const int cbeg = q * Wb * kNB;                       // this CTA's columns [cbeg, cend)
  const int cend = min(ld, cbeg + Wb * kNB);
  const int cw = cend - cbeg;                          // multiple of 16
  double* slab = sm;                       // 32 x lds
  double* Ls = slab + (size_t)kNB * lds;   // L^-1            (32 x kLp)
  double* Rs = Ls + kNB * kLp;             // A[I,b] rows     (32 x kLp); the block to factorise
  double* Vs = Rs + kNB * kLp;             // V_I
  double* Ps = Vs + kNB * kLp;             // P_I; L^-1 out of factor32
  double* scratch = Ps + kNB * kLp;
  double* A = Aall + z * stride;
  unsigned* counter = sync + 2 * z;        // barrier arrivals (monotonic)
  unsigned* lflag = sync + 2 * z + 1;      // number of pivot blocks whose L^-1 has been published
  double logdet_part = 0.0;                // lanes of warp 7: sum of log pivots of the blocks this CTA factored
  const long long sPz = (long long)Npad * kNB;          // per-matrix panel size
  const long long sPall = sPz * gridDim.y;              // per-buffer size

  // ---- load the slab part (rows r0..r0+31, columns cbeg..cend incl. zero pads) -----------------------------
  for (int e = tid; e < kNB * (cw / 2); e += kSlabThreads) {
    const int r = e / (cw / 2), c2 = (e - r * (cw / 2)) * 2;
    const int gr_ = r0 + r;
    double2 v = make_double2(0.0, 0.0);
    if (gr_ < N) v = *reinterpret_cast<const double2*>(A + (size_t)gr_ * ld + cbeg + c2);
    *reinterpret_cast<double2*>(slab + (size_t)r * lds + c2) = v;
  }
  __syncthreads();

  // Factor the diagonal block bb (of this CTA's slab): D = L L', publish L^-1 (factor32.cuh: register-resident 8 x 8
  // diagonal tiles + DMMA tile products; the N sequential pivots are the critical path of the whole inverse).
  // The block is expected in Rs (identity-padded) unless from_slab; Ps receives L^-1.
  auto factor_block = [&](int bb, bool from_slab) {
    const int c0 = bb * kNB;
    const int nb = min(kNB, N - c0);
    double* Lg = Lbuf + ((size_t)(bb & 1) * gridDim.y + z) * kNB * kNB;
    if (from_slab) {
      for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
        const int r = e >> 5, c = e & 31;
        // identity padding keeps a partial last block a no-op for the extra pivots
        Rs[r * kLp + c] = (r < nb && c < nb) ? slab[(size_t)r * lds + c0 - cbeg + c] : (r == c ? 1.0 : 0.0);
      }
      __syncthreads();
    }
    factor32(Rs, Ps, scratch, &logdet_part, info);
    for (int e = tid; e < kNB * kNB; e += kSlabThreads) Lg[e] = Ps[(e >> 5) * kLp + (e & 31)];
    slab_publish(lflag, (unsigned)(bb + 1));
  };
  // own V_I (rows r0..) of the current step from L2, for CTAs that did not compute it
  auto fetch_V = [&](const double* Vg) {
    for (int e = tid; e < kNB * kNB; e += kSlabThreads) Vs[(e >> 5) * kLp + (e & 31)] = __ldcg(Vg + (size_t)r0 * kNB + e);
    __syncthreads();
  };

  if (I == 0 && q == 0) factor_block(0, true);

  for (int b = 0; b < nblk; ++b) {
    const int c0 = b * kNB;
    const int nb = min(kNB, N - c0);
    const int buf = b % 3;
    const bool own_piv = (b / Wb == q);                  // the pivot columns live in this column part
    double* Vg = Vbuf + buf * sPall + z * sPz;
    double* Pg = Pbuf + buf * sPall + z * sPz;
    const double* Lg = Lbuf + ((size_t)(b & 1) * gridDim.y + z) * kNB * kNB;
    SP_BEGIN(((b + 1 < nblk) && I == b + 1 && ((b + 1) / Wb == q) && own_piv) ? 0 : (own_piv ? 1 : 2))

    // ---- phase 2: V_I = R L^-T, P_I = V L^-1 -------------------------------------------------------------
    if (own_piv) {
      // the rows of the pivot columns are staged BEFORE waiting for L^-1: they only depend on this CTA's own slab
      for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
        const int r = e >> 5, c = e & 31;
        double v;
        if (I == b) v = (r == c && r < nb) ? 1.0 : 0.0;          // identity rows -> P = D^-1
        else v = (c < nb) ? slab[(size_t)r * lds + c0 - cbeg + c] : 0.0;
        Rs[r * kLp + c] = v;
      }
      slab_wait(lflag, (unsigned)(b + 1));
      SP(0);
      for (int e = tid; e < kNB * kNB; e += kSlabThreads) Ls[(e >> 5) * kLp + (e & 31)] = __ldcg(Lg + e);
      __syncthreads();
      SP(1);
      mm32_nt(Rs, Ls, Vs, warp, lane);     // V[r][c] = sum_b R[r][b] Linv[c][b]
      __syncthreads();
      SP(2);
      // P[r][c] = sum_b V[r][b] Linv[b][c]: B operand rows = columns of Linv
      {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int tile = warp * 2 + t;
          const int tI = tile >> 2, tJ = tile & 3;
          double c0_ = 0.0, c1_ = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double a = Vs[(8 * tI + gr) * kLp + 4 * k + gc];
            const double bb = Ls[(4 * k + gc) * kLp + 8 * tJ + gr];
            dmma884_s(c0_, c1_, a, bb);
          }
          Ps[(8 * tI + gr) * kLp + 8 * tJ + 2 * gc] = c0_;
          Ps[(8 * tI + gr) * kLp + 8 * tJ + 2 * gc + 1] = c1_;
        }
      }
      __syncthreads();
      SP(3);
      for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
        const int r = e >> 5, c = e & 31;
        Vg[(size_t)(r0 + r) * kNB + c] = Vs[r * kLp + c];
        Pg[(size_t)(r0 + r) * kNB + c] = Ps[r * kLp + c];
      }
    }
    slab_arrive(counter);
    SP(4);
    const bool look = (b + 1 < nblk) && I == b + 1 && ((b + 1) / Wb == q);
    if (!(look && own_piv)) slab_wait(counter, (unsigned)(b + 1) * (unsigned)ncta);
    SP(5);

    // ---- phase 3 --------------------------------------------------------------------------------------
    if (look) {
      // Lookahead: the next pivot CTA only needs its diagonal block up to date (everything else in its slab is
      // dead: its rows are replaced by P_J' at the next step), so it updates those 32 columns with its own V and
      // factors the block while the other CTAs do their trailing updates.
      if (!own_piv) fetch_V(Vg);
      const int c1 = (b + 1) * kNB;
      const int nb1 = min(kNB, N - c1);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int tile = warp * 2 + t;
        const int tI = tile >> 2, tJ = tile & 3;
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          dmma884_s(n0, n1, Vs[(8 * tI + gr) * kLp + 4 * k + gc], Vs[(8 * tJ + gr) * kLp + 4 * k + gc]);
        const double* cp = slab + (size_t)(8 * tI + gr) * lds + c1 - cbeg + 8 * tJ + 2 * gc;
        const int r = 8 * tI + gr, c = 8 * tJ + 2 * gc;
        Rs[r * kLp + c] = (r < nb1 && c < nb1) ? cp[0] - n0 : (r == c ? 1.0 : 0.0);
        Rs[r * kLp + c + 1] = (r < nb1 && c + 1 < nb1) ? cp[1] - n1 : (r == c + 1 ? 1.0 : 0.0);
      }
      __syncthreads();
      SP(6);
      factor_block(b + 1, false);
      SP(7);
    } else if (I == b + 1) {
      // the other column parts of the next pivot rows: dead until they are replaced at the next step
    } else if (I != b) {
      // trailing update of the slab part. Each warp owns the 8-column tiles tj = warp, warp + 8, ...; ALL its B
      // fragments (rows of V_J, straight from L2) are requested up front so that one L2 latency covers the whole
      // phase; A fragments (V_I) are re-read from shared memory per tile.
      constexpr int kMaxT = 12;              // tiles per warp: 768 / 8 / 8
      const int ntile = cw / 8;              // columns >= N only see zero V rows
      double bf[kMaxT][8];
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int tj = warp + 8 * t;
        const int col = cbeg + 8 * tj;       // global column = row of V
        const bool live = tj < ntile && !(col >= c0 && col < c0 + kNB);
        const double* vj = Vg + (size_t)(live ? col + gr : 0) * kNB + gc;
#pragma unroll
        for (int k = 0; k < 8; ++k) bf[t][k] = live ? __ldcg(vj + 4 * k) : 0.0;
      }
      if (!own_piv) fetch_V(Vg);
      SP(8);
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int tj = warp + 8 * t;
        const int col = cbeg + 8 * tj;
        if (tj < ntile && !(col >= c0 && col < c0 + kNB)) {  // pivot block column: replaced below
#pragma unroll
          for (int rt = 0; rt < 4; ++rt) {
            double* cp = slab + (size_t)(8 * rt + gr) * lds + 8 * tj + 2 * gc;
            double2 c = *reinterpret_cast<double2*>(cp);
            double n0 = 0.0, n1 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma884_s(n0, n1, Vs[(8 * rt + gr) * kLp + 4 * k + gc], bf[t][k]);
            c.x -= n0;
            c.y -= n1;
            *reinterpret_cast<double2*>(cp) = c;
          }
        }
      }
      if (own_piv) {
        for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
          const int r = e >> 5, c = e & 31;
          if (c < nb) slab[(size_t)r * lds + c0 - cbeg + c] = Ps[r * kLp + c];
        }
      }
    } else {
      // pivot rows: A[b, J] = P_J' for J != b, A[b, b] = -D^-1
      // thread = (row r of the slab, columns j = jq, jq + 8, ...): 8 independent L2 loads in flight per batch
      const int r = tid & 31, jq = tid >> 5;
      const int jend = min(N, cend);
      for (int j0 = cbeg + jq; j0 < jend; j0 += 64) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          v[u] = (j < jend) ? __ldcg(Pg + (size_t)j * kNB + r) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          if (j < jend && r < nb) slab[(size_t)r * lds + j - cbeg] = (j >= c0 && j < c0 + kNB) ? -v[u] : v[u];
        }
      }
    }
    __syncthreads();
    SP(9);
  }

  if (warp == 7) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) logdet_part += __shfl_xor_sync(0xffffffffu, logdet_part, o);
    if (lane == 0 && logdet_part != 0.0) atomicAdd(logdet + z, logdet_part);
  }
  // ---- write back, flipping the sign (-A^-1 -> A^-1) -------------------------------------------------------
  for (int e = tid; e < kNB * (cw / 2); e += kSlabThreads) {
    const int r = e / (cw / 2), c2 = (e - r * (cw / 2)) * 2;
    const int gr_ = r0 + r;
    if (gr_ >= N) continue;
    double2 v = *reinterpret_cast<double2*>(slab + (size_t)r * lds + c2);
    v.x = (cbeg + c2 < N) ? -v.x : 0.0;
    v.y = (cbeg + c2 + 1 < N) ? -v.y : 0.0;
    *reinterpret_cast<double2*>(A + (size_t)gr_ * ld + cbeg + c2) = v;
  }
}

// column parts: Q parts of Wb 32-blocks each
struct SlabCfg { int nblk, Q, Wb, lds; size_t smem; };

static SlabCfg slab_cfg(int N, int ld, int Z, int budget) {
  SlabCfg c;
  c.nblk = (N + kNB - 1) / kNB;
  static const int forced = getenv("MCB_SLAB_Q") ? atoi(getenv("MCB_SLAB_Q")) : 0;
  int Q = 1;
  // more parts shorten the trailing update (DMMA bound: 2 * 32 * 32 * N flop per CTA and step at ~250 GFLOP/s per SM)
  // until the lookahead CTA's factorisation is the longer branch; all CTAs must be co-resident
  // (measured at N = 401, one matrix: 2 parts 122 us, 3 parts 117 us, 4 parts 117 us)
  for (int cand = forced > 0 ? forced : 3; cand >= 2; --cand) {
    const int wb = (c.nblk + cand - 1) / cand;
    if ((long long)c.nblk * cand * Z <= budget && (cand - 1) * wb < c.nblk) { Q = cand; break; }
  }
  c.Q = Q;
  c.Wb = (c.nblk + Q - 1) / Q;
  const int cwmax = std::min(ld, c.Wb * kNB);
  c.lds = cwmax + 8;   // cw % 16 == 0 -> lds % 16 == 8: conflict-free 16-byte tile rows
  c.smem = sizeof(double) * ((size_t)kNB * c.lds + 4 * (size_t)kNB * kLp + kF32Scratch + 8);
  return c;
}

size_t spd_inverse_slab_smem(int ld) {
  return sizeof(double) * ((size_t)kNB * (ld + 8) + 4 * (size_t)kNB * kLp + kF32Scratch + 8);
}

bool spd_inverse_slab_ok(int N, int ld) {
  return N <= 768 && spd_inverse_slab_smem(ld) <= 227 * 1024;
}

cudaError_t spd_inverse_slab_reserve(DenseWs& ws, int N, int Z) {
  const int nblk = (N + kNB - 1) / kNB;
  const size_t panel = (size_t)nblk * kNB * kNB * Z;
  cudaError_t e;
  if ((e = ws.slabV.ensure(3 * panel)) != cudaSuccess) return e;
  if ((e = ws.slabP.ensure(3 * panel)) != cudaSuccess) return e;
  if ((e = ws.slabL.ensure(2 * (size_t)Z * kNB * kNB)) != cudaSuccess) return e;
  return ws.bar.ensure(2 * (size_t)(Z > 64 ? Z : 64));
}

unsigned* slab_sync_words(DenseWs& ws) { return ws.bar.p; }

cudaError_t spd_inverse_slab(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                             int* info, DenseWs& ws, long long* launches, bool pre_zeroed) {
  cudaError_t e;
  if ((e = spd_inverse_slab_reserve(ws, N, Z)) != cudaSuccess) return e;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(spd_inverse_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess)
      return e;
    attr_set = true;
  }
  if (!pre_zeroed && (e = cudaMemsetAsync(logdet, 0, sizeof(double) * Z, s)) != cudaSuccess) return e;
  // all CTAs of a launch must be co-resident (1 CTA per SM): batch the matrices accordingly. ws.max_ctas is the
  // number of SMs the caller knows to be free (the M-step runs this next to a persistent kernel).
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int budget = (ws.max_ctas > 0 && ws.max_ctas < sms) ? ws.max_ctas : sms;
  const int nblk = (N + kNB - 1) / kNB;
  const int zfit = budget / nblk > 0 ? budget / nblk : 1;
  const SlabCfg c = slab_cfg(N, ld, Z < zfit ? Z : zfit, budget);
  const int zmax = budget / (c.nblk * c.Q) > 0 ? budget / (c.nblk * c.Q) : 1;
  for (int z0 = 0; z0 < Z; z0 += zmax) {
    const int zc = (Z - z0 < zmax) ? Z - z0 : zmax;
    if (!(pre_zeroed && z0 == 0) && (e = cudaMemsetAsync(ws.bar.p, 0, sizeof(unsigned) * 2 * zc, s)) != cudaSuccess) return e;
    double* Az = A + (size_t)z0 * stride;
    double* ldz = logdet + z0;
    double* Vb = ws.slabV.p;
    double* Pb = ws.slabP.p;
    double* Lb = ws.slabL.p;
    unsigned* bp = ws.bar.p;
    int N_ = N, ld_ = ld, lds_ = c.lds, Q_ = c.Q, Wb_ = c.Wb;
    long long st_ = stride;
    void* args[] = {&Az, &N_, &ld_, &st_, &Q_, &Wb_, &lds_, &Vb, &Pb, &Lb, &bp, &ldz, &info};
    e = cudaLaunchCooperativeKernel((const void*)spd_inverse_slab_kernel, dim3(c.nblk * c.Q, zc), dim3(kSlabThreads),
                                    args, c.smem, s);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace mcb
