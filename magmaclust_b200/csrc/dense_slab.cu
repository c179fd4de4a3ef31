// Latency-optimised inverse for mid-size grids (N <= 768): one persistent cooperative kernel per batch of
// matrices. The multi-launch blocked sweep of dense.cu pays ~2 launches and ~30 us per 32-wide pivot block;
// at N = 401 (BASELINE config C2) that is the whole cost of every theta_k objective evaluation. Here each CTA
// keeps a 32-row slab of the matrix in SHARED memory for the entire inversion (32 x 416 doubles = 108 KB at
// N = 401, 13 CTAs per matrix), the CTAs of one matrix exchange the 32 x 32 panels through L2 and synchronise on
// a global-memory barrier (~1 us) instead of kernel boundaries.
//
// Same algorithm as dense.cu (blocked symmetric sweep in Cholesky form), per pivot block b:
//   phase 2  every CTA: V_I = A[I,b] L^-T, P_I = V_I L^-1 (DMMA), publish V_I and P_I          | grid barrier
//   phase 3  CTA I != b, b+1: slab[:, J] -= V_I V_J' for J != b (DMMA, V_J streamed from L2), slab[:, b] = P_I;
//            CTA b: slab[:, J] = P_J', slab[:, b] = -D^-1;
//            CTA b+1 (lookahead): updates only its diagonal block and eliminates it, D = Lu Delta Lu', in
//            registers (1 CTA barrier per pivot), publishing L^-1 = Delta^-1/2 Lu^-1 for the next step  | grid barrier
// After the last block the slabs hold -A^-1; they are negated on the way back to global memory.
#include <cooperative_groups.h>

#include "common.cuh"
#include "factor32.cuh"

namespace mcb {

#ifdef MCB_SLAB_PROFILE
__device__ long long g_slab_prof[64][12];
#define SLAB_T(slot) do { if (tid == 0 && z == 0 && b < 64 && I == (b + 1) % nblk) g_slab_prof[b][slot] = clock64(); } while (0)
#define SLAB_TP(slot) do { if (tid == 0 && z == 0 && b < 64 && I == b) g_slab_prof[b][slot] = clock64(); } while (0)
__device__ long long g_slab_prof2[64][12];  // a trailing-update CTA (I = b + 2)
#define SLAB_TO(slot) do { if (tid == 0 && z == 0 && b < 64 && I == (b + 2) % nblk) g_slab_prof2[b][slot] = clock64(); } while (0)
#else
#define SLAB_TO(slot)
#define SLAB_T(slot)
#define SLAB_TP(slot)
#endif

constexpr int kSlabThreads = 256;
constexpr int kLp = kNB + 4;  // 36: row stride of the 32 x 32 operand tiles (conflict-free DMMA fragments)

__device__ __forceinline__ void dmma884_s(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1/d to full double accuracy without the slow IEEE division path: 20-bit hardware seed + 2 Newton steps.
// Sits on the serial critical path of the pivot-block elimination (one per pivot).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
}

// barrier over the `nblk` CTAs of one matrix; `target` = number of arrivals expected so far
__device__ __forceinline__ void slab_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release: everything this CTA wrote (bar.sync above orders the other threads' writes before this point) becomes
    // visible before the arrival; acquire: nothing after the barrier is read before the last arrival is seen
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
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
spd_inverse_slab_kernel(double* __restrict__ Aall, int N, int ld, long long stride, int lds,
                        double* __restrict__ Vbuf, double* __restrict__ Pbuf, double* __restrict__ Lbuf,
                        unsigned* __restrict__ bar, double* __restrict__ logdet, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int nblk = gridDim.x;
  const int I = blockIdx.x, z = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gr = lane >> 2, gc = lane & 3;
  const int r0 = I * kNB;
  const int Npad = nblk * kNB;
  double* slab = sm;                       // 32 x lds
  double* Ls = slab + (size_t)kNB * lds;   // L^-1            (32 x kLp)
  double* Rs = Ls + kNB * kLp;             // A[I,b] rows     (32 x kLp)
  double* Vs = Rs + kNB * kLp;             // V_I
  double* Ps = Vs + kNB * kLp;             // P_I
  double* uvec = Ps + kNB * kLp;           // scratch of factor32
  double* A = Aall + z * stride;
  unsigned* counter = bar + z;
  unsigned arrivals = 0;
  double logdet_part = 0.0;  // lanes of warp 7: sum of log pivots of the blocks this CTA factored
  const long long sPz = (long long)Npad * kNB;          // per-matrix panel size
  const long long sPall = sPz * gridDim.y;              // per-parity size

  // ---- load the slab (rows r0..r0+31, all ld columns incl. zero pads) ----------------------------------
  for (int e = tid; e < kNB * (ld / 2); e += kSlabThreads) {
    const int r = e / (ld / 2), c2 = (e - r * (ld / 2)) * 2;
    const int gr_ = r0 + r;
    double2 v = make_double2(0.0, 0.0);
    if (gr_ < N) v = *reinterpret_cast<const double2*>(A + (size_t)gr_ * ld + c2);
    *reinterpret_cast<double2*>(slab + (size_t)r * lds + c2) = v;
  }
  __syncthreads();

  // Factor this CTA's diagonal block (pivot block bb): D = L L', publish L^-1 (factor32.cuh: register-resident 8 x 8
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
        Rs[r * kLp + c] = (r < nb && c < nb) ? slab[(size_t)r * lds + c0 + c] : (r == c ? 1.0 : 0.0);
      }
      __syncthreads();
    }
    factor32(Rs, Ps, uvec, &logdet_part, info);
    for (int e = tid; e < kNB * kNB; e += kSlabThreads) Lg[e] = Ps[(e >> 5) * kLp + (e & 31)];
  };

  if (I == 0) factor_block(0, true);
  arrivals += nblk;
  slab_barrier(counter, arrivals);

  for (int b = 0; b < nblk; ++b) {
    const int c0 = b * kNB;
    const int nb = min(kNB, N - c0);
    const int par = b & 1;
    double* Vg = Vbuf + par * sPall + z * sPz;
    double* Pg = Pbuf + par * sPall + z * sPz;
    const double* Lg = Lbuf + ((size_t)par * gridDim.y + z) * kNB * kNB;

    // ---- phase 2: V_I = R L^-T, P_I = V L^-1 -------------------------------------------------------------
    SLAB_T(1);
    SLAB_TO(1);
    for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
      const int r = e >> 5, c = e & 31;
      Ls[r * kLp + c] = __ldcg(Lg + e);
      double v;
      if (I == b) v = (r == c && r < nb) ? 1.0 : 0.0;          // identity rows -> P = D^-1
      else v = (c < nb) ? slab[(size_t)r * lds + c0 + c] : 0.0;
      Rs[r * kLp + c] = v;
    }
    __syncthreads();
    mm32_nt(Rs, Ls, Vs, warp, lane);     // V[r][c] = sum_b R[r][b] Linv[c][b]
    __syncthreads();
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
    for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
      const int r = e >> 5, c = e & 31;
      Vg[(size_t)(r0 + r) * kNB + c] = Vs[r * kLp + c];
      Pg[(size_t)(r0 + r) * kNB + c] = Ps[r * kLp + c];
    }
    SLAB_T(2);
    SLAB_TO(2);
    arrivals += nblk;
    slab_barrier(counter, arrivals);
    SLAB_T(3);
    SLAB_TO(3);

    // ---- phase 3 --------------------------------------------------------------------------------------
    if (I == b + 1) {
      // Lookahead: the next pivot CTA only needs its diagonal block up to date (everything else in its slab is
      // dead: its rows are replaced by P_J' at the next step), so it updates those 32 columns with its own V and
      // factors the block while the other CTAs do their trailing updates.
      const int c1 = (b + 1) * kNB;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int tile = warp * 2 + t;
        const int tI = tile >> 2, tJ = tile & 3;
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          dmma884_s(n0, n1, Vs[(8 * tI + gr) * kLp + 4 * k + gc], Vs[(8 * tJ + gr) * kLp + 4 * k + gc]);
        const double* cp = slab + (size_t)(8 * tI + gr) * lds + c1 + 8 * tJ + 2 * gc;
        const int r = 8 * tI + gr, c = 8 * tJ + 2 * gc;
        const int nb1 = min(kNB, N - c1);
        Rs[r * kLp + c] = (r < nb1 && c < nb1) ? cp[0] - n0 : (r == c ? 1.0 : 0.0);
        Rs[r * kLp + c + 1] = (r < nb1 && c + 1 < nb1) ? cp[1] - n1 : (r == c + 1 ? 1.0 : 0.0);
      }
      __syncthreads();
      SLAB_T(4);
      factor_block(b + 1, false);
      SLAB_T(5);
    } else if (I != b) {
      // trailing update of the slab. Each warp owns the 8-column tiles tj = warp, warp + 8, ...; ALL its B
      // fragments (rows of V_J, straight from L2) are requested up front so that one L2 latency covers the whole
      // phase; A fragments (V_I) are re-read from shared memory per tile.
      constexpr int kMaxT = 12;              // tiles per warp: ld / 8 / 8 <= 12 for N <= 768
      const int ntile = ld / 8;              // ld <= Npad; columns >= N only see zero V rows
      double bf[kMaxT][8];
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int tj = warp + 8 * t;
        const int col = 8 * tj;
        const bool live = tj < ntile && !(col >= c0 && col < c0 + kNB);
        const double* vj = Vg + (size_t)(live ? col + gr : 0) * kNB + gc;
#pragma unroll
        for (int k = 0; k < 8; ++k) bf[t][k] = live ? __ldcg(vj + 4 * k) : 0.0;
      }
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) {
        const int tj = warp + 8 * t;
        const int col = 8 * tj;
        if (tj < ntile && !(col >= c0 && col < c0 + kNB)) {  // pivot block column: replaced below
#pragma unroll
          for (int rt = 0; rt < 4; ++rt) {
            double* cp = slab + (size_t)(8 * rt + gr) * lds + col + 2 * gc;
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
      for (int e = tid; e < kNB * kNB; e += kSlabThreads) {
        const int r = e >> 5, c = e & 31;
        if (c < nb) slab[(size_t)r * lds + c0 + c] = Ps[r * kLp + c];
      }
    } else {
      // pivot rows: A[b, J] = P_J' for J != b, A[b, b] = -D^-1
      // thread = (row r of the slab, columns j = jq, jq + 8, ...): 8 independent L2 loads in flight per batch
      const int r = tid & 31, jq = tid >> 5;
      for (int j0 = jq; j0 < N; j0 += 64) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          v[u] = (j < N) ? __ldcg(Pg + (size_t)j * kNB + r) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 8 * u;
          if (j < N && r < nb) slab[(size_t)r * lds + j] = (j >= c0 && j < c0 + kNB) ? -v[u] : v[u];
        }
      }
    }
    __syncthreads();
    SLAB_T(7);
    SLAB_TO(7);
    if (b + 1 < nblk) {
      arrivals += nblk;
      slab_barrier(counter, arrivals);
    }
  }

  if (warp == 7) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) logdet_part += __shfl_xor_sync(0xffffffffu, logdet_part, o);
    if (lane == 0) atomicAdd(logdet + z, logdet_part);
  }
  // ---- write back, flipping the sign (-A^-1 -> A^-1) -------------------------------------------------------
  for (int e = tid; e < kNB * (ld / 2); e += kSlabThreads) {
    const int r = e / (ld / 2), c2 = (e - r * (ld / 2)) * 2;
    const int gr_ = r0 + r;
    if (gr_ >= N) continue;
    double2 v = *reinterpret_cast<double2*>(slab + (size_t)r * lds + c2);
    v.x = (c2 < N) ? -v.x : 0.0;
    v.y = (c2 + 1 < N) ? -v.y : 0.0;
    *reinterpret_cast<double2*>(A + (size_t)gr_ * ld + c2) = v;
  }
}

static int slab_lds(int ld) { return ld + 8; }  // ld % 16 == 0 -> lds % 16 == 8: conflict-free 16-byte tile rows

size_t spd_inverse_slab_smem(int ld) {
  return sizeof(double) * ((size_t)kNB * slab_lds(ld) + 4 * (size_t)kNB * kLp + 2 * (kNB + 4) + kNB + 8);
}

bool spd_inverse_slab_ok(int N, int ld) {
  return N <= 768 && spd_inverse_slab_smem(ld) <= 227 * 1024;
}

cudaError_t spd_inverse_slab_reserve(DenseWs& ws, int N, int Z) {
  const int nblk = (N + kNB - 1) / kNB;
  const size_t panel = (size_t)nblk * kNB * kNB * Z;
  cudaError_t e;
  if ((e = ws.slabV.ensure(2 * panel)) != cudaSuccess) return e;
  if ((e = ws.slabP.ensure(2 * panel)) != cudaSuccess) return e;
  if ((e = ws.slabL.ensure(2 * (size_t)Z * kNB * kNB)) != cudaSuccess) return e;
  return ws.bar.ensure(Z > 64 ? Z : 64);
}

cudaError_t spd_inverse_slab(cudaStream_t s, double* A, int N, int ld, long long stride, int Z, double* logdet,
                             int* info, DenseWs& ws, long long* launches) {
  const int nblk = (N + kNB - 1) / kNB;
  const int lds = slab_lds(ld);
  const size_t smem = spd_inverse_slab_smem(ld);
  cudaError_t e;
  if ((e = spd_inverse_slab_reserve(ws, N, Z)) != cudaSuccess) return e;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(spd_inverse_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess)
      return e;
    attr_set = true;
  }
  if ((e = cudaMemsetAsync(logdet, 0, sizeof(double) * Z, s)) != cudaSuccess) return e;
  // all CTAs of a launch must be co-resident (1 CTA per SM): batch the matrices accordingly
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int zmax = sms / nblk > 0 ? sms / nblk : 1;
  for (int z0 = 0; z0 < Z; z0 += zmax) {
    const int zc = (Z - z0 < zmax) ? Z - z0 : zmax;
    if ((e = cudaMemsetAsync(ws.bar.p, 0, sizeof(unsigned) * zc, s)) != cudaSuccess) return e;
    double* Az = A + (size_t)z0 * stride;
    double* ldz = logdet + z0;
    double* Vb = ws.slabV.p;
    double* Pb = ws.slabP.p;
    double* Lb = ws.slabL.p;
    unsigned* bp = ws.bar.p;
    int N_ = N, ld_ = ld, lds_ = lds;
    long long st_ = stride;
    void* args[] = {&Az, &N_, &ld_, &st_, &lds_, &Vb, &Pb, &Lb, &bp, &ldz, &info};
    e = cudaLaunchCooperativeKernel((const void*)spd_inverse_slab_kernel, dim3(nblk, zc), dim3(kSlabThreads), args,
                                    smem, s);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace mcb
