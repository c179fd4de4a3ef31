// Rank-256 symmetric update of the large-grid inverse (N > 768, config C4: N = 8192), the one DGEMM-shaped contraction of
// the path (dense.cu: spd_inverse, wide path):   C[i][j] = sign * (C[i][j] - V[i] . V[j])  for the tiles on and below the
// diagonal, mirrored, with the pivot rows / columns taken from the finished panel P.
//
// Blackwell data movement, FP64 math: FP64 has no tcgen05 path, so the math stays on mma.sync.m8n8k4.f64 (DMMA), but the
// operands travel by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 6-stage shared-memory ring guarded by mbarriers,
// issued by ONE producer thread; 8 consumer warps do nothing but LDS.128 + DMMA + the epilogue. Registers are moved from the
// producer warpgroup to the consumers with setmaxnreg (40 / 232 per thread: 128 of them are the consumer's accumulators). The kernel is persistent
// (one CTA per SM walks the lower-triangle tiles of all matrices), so the producer fills the ring with the NEXT tile's
// operands while the consumers read-modify-write C for the current one: the short contraction (K = 256 = 16 stages) no
// longer pays a pipeline fill and a serial epilogue per tile as the cp.async kernel of dense.cu does
// (profiles/r01e_large_grid_inverse.txt: 0.75 ms per pass against 0.46 ms of DMMA time).
//
// Tile 128 x 128, consumer warp (wy, wx) of 2 x 4 owns 64 x 32: 8 x 4 DMMA accumulator tiles. A stage holds 16 contraction
// columns of the A rows and of the B rows: two boxes of 128 rows x 128 bytes, each landing 128B-swizzled (16-byte chunk
// index ^= row & 7). Contraction index permuted per k-step: for the step pair t in {0, 1} lane (g = lane / 4, c = lane % 4)
// reads the ONE 16-byte chunk 2c + t of its row (physical chunk (2c + t) ^ g) and uses its two doubles for k-steps 2t and
// 2t + 1. A quarter warp (g in {0, 1}, c in 0..3) then touches the 8 distinct physical chunks of a 128-byte bank cycle:
// conflict-free with the natural row mapping, so the epilogue keeps 16-byte accesses to C.
#include <cuda.h>

#include "common.cuh"

namespace mcb {

namespace {

constexpr int kTile = 128;
constexpr int kStageCols = 16;                       // doubles per row and stage = 128 bytes
constexpr int kStages = 6;
constexpr int kBoxBytes = kTile * kStageCols * 8;    // 16 KB
constexpr int kStageBytes = 2 * kBoxBytes;           // A rows + B rows
constexpr int kConsumerWarps = 8;
constexpr int kThreads = 32 * (kConsumerWarps + 4);   // two consumer warpgroups + one producer warpgroup (setmaxnreg is per warpgroup)

struct UpdParams {
  int N, K, rows;              // matrix size, contraction length (multiple of 16), rows per matrix of V
  double* C; int ldc; long long sC;
  const double* P; long long sP; int ldp, c0, nb;   // pivot rows == pivot columns [c0, c0 + nb)
  double sign;
  int ntile, ntri, total;      // tile rows, lower-triangle tiles per matrix, tiles over all matrices
  int stagger_ns;              // start offset between CTAs (see the consumers)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ double2 lds128(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// linear tile id -> (matrix z, tile row ti >= tile column tj)
__device__ __forceinline__ void tile_of(int id, const UpdParams& p, int& z, int& ti, int& tj) {
  z = id / p.ntri;
  const int t = id - z * p.ntri;
  int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  ti = r;
  tj = t - r * (r + 1) / 2;
}

__global__ void __launch_bounds__(kThreads, 1) sweep_update_tma_kernel(const __grid_constant__ CUtensorMap vmap, UpdParams p) {
  extern __shared__ unsigned char smem_raw[];
  // the swizzle pattern is defined on absolute shared-memory addresses: 1024-byte aligned stage buffers
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(smem_raw) + 1023ull) & ~1023ull);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(base + kStages * kStageBytes);
  unsigned long long* empty = full + kStages;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nk = p.K / kStageCols;

  if (warp >= kConsumerWarps) {
    // ===== producer warpgroup: gives its registers away; one lane streams the operands of every tile through the ring =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kConsumerWarps && lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      for (int id = blockIdx.x; id < p.total; id += gridDim.x) {
        int z, ti, tj;
        tile_of(id, p, z, ti, tj);
        const int ya = z * p.rows + ti * kTile, yb = z * p.rows + tj * kTile;
        for (int k = 0; k < nk; ++k) {
          mbar_wait(empty + stage, phase ^ 1);   // a fresh barrier passes the wait on the opposite parity
          mbar_expect_tx(full + stage, kStageBytes);
          unsigned char* dst = base + stage * kStageBytes;
          tma_load_2d(dst, &vmap, k * kStageCols, ya, full + stage);
          tma_load_2d(dst + kBoxBytes, &vmap, k * kStageCols, yb, full + stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = lane >> 2, c = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  // byte offset of this lane's 16-byte chunk inside an 8-row group, for the step pairs t = 0, 1
  const int off0 = g * 128 + (((2 * c) ^ g) << 4);
  const int off1 = g * 128 + (((2 * c + 1) ^ g) << 4);
  // Experiment (default off, MCB_TMA_STAGGER_NS): every tile costs the same and every CTA starts at the same time, so all SMs
  // reach their epilogue (384 KB of C traffic per tile) together. Starting the CTAs up to 15/16 of a tile time apart did NOT
  // help (N = 8192: 34.97 ms without, 35.03 / 36.21 ms with 1 / 2.5 us steps): the start offsets are pure cost.
  if (p.stagger_ns > 0 && p.total > (int)gridDim.x) {
    const unsigned steps = (blockIdx.x * 7u) & 15u;
    for (unsigned q = 0; q < steps; ++q) __nanosleep(p.stagger_ns);
  }
  const unsigned sbase = smem_u32(base);
  int stage = 0, held = -1;
  unsigned phase = 0;
  for (int id = blockIdx.x; id < p.total; id += gridDim.x) {
    int z, ti, tj;
    tile_of(id, p, z, ti, tj);
    const int m0 = ti * kTile, n0 = tj * kTile;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    {
      // This warp's 64 x 32 patch of C (64 rows of 256 bytes) is requested into L2 now, so that the read-modify-write after the
      // contraction (33 us later at full DMMA rate) does not wait for HBM and the C reads of all SMs do not arrive in one burst
      const double* cp = p.C + z * p.sC + (size_t)(m0 + wm) * p.ldc + n0 + wn;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = 16 * q + (lane >> 1);
        if (m0 + wm + r < p.N && n0 + wn + 16 * (lane & 1) < p.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (size_t)r * p.ldc + 16 * (lane & 1)));
      }
    }
    for (int k = 0; k < nk; ++k) {
      mbar_wait(full + stage, phase);
      // Release of the PREVIOUS stage, one stage late: an mbarrier arrive does not wait for this warp's outstanding shared
      // loads (measured: with the arrive right behind the loads of its own stage, about one launch in 700 had a 16-byte
      // fragment overwritten by the producer's next TMA box before the load had returned). Behind the wait above, every DMMA
      // of the previous stage has been issued (a warp issues in order), i.e. every load of that stage has delivered.
      if (held >= 0 && lane == 0) mbar_arrive(empty + held);
      held = stage;
      const unsigned sa = sbase + stage * kStageBytes + wm * 128;
      const unsigned sb = sbase + stage * kStageBytes + kBoxBytes + wn * 128;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int off = t ? off1 : off0;
        double2 a2[8], b2[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a2[i] = lds128(sa + i * 1024 + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) b2[j] = lds128(sb + j * 1024 + off);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a2[i].x, b2[j].x);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a2[i].y, b2[j].y);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }

    // ----- epilogue (the producer is already filling the ring for the next tile) -----
    double* Cz = p.C + z * p.sC;
    const bool clean = (m0 + kTile <= p.c0 || m0 >= p.c0 + p.nb) && (n0 + kTile <= p.c0 || n0 >= p.c0 + p.nb) &&
                       m0 + kTile <= p.N && n0 + kTile <= p.N && ti != tj && (p.ldc % 2 == 0);
    if (clean) {
      // away from the pivot block, the diagonal and the edge: two batches of 16 independent 16-byte loads (they hit L2: the
      // tile was prefetched at the start of the contraction), then stores + mirror
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 cv[4][4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int gi = m0 + wm + 8 * (4 * half + ii) + g;
#pragma unroll
          for (int j = 0; j < 4; ++j) cv[ii][j] = *reinterpret_cast<const double2*>(Cz + (size_t)gi * p.ldc + n0 + wn + 8 * j + 2 * c);
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = 4 * half + ii;
          const int gi = m0 + wm + 8 * i + g;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int gj = n0 + wn + 8 * j + 2 * c;
            const double r0 = p.sign * (cv[ii][j].x - acc[i][j][0]);
            const double r1 = p.sign * (cv[ii][j].y - acc[i][j][1]);
            *reinterpret_cast<double2*>(Cz + (size_t)gi * p.ldc + gj) = make_double2(r0, r1);
            Cz[(size_t)gj * p.ldc + gi] = r0;
            Cz[(size_t)(gj + 1) * p.ldc + gi] = r1;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int gi = m0 + wm + 8 * i + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int gq = n0 + wn + 8 * j + 2 * c + q;
            if (gi >= p.N || gq >= p.N || gq > gi) continue;   // (gq, gi) above the diagonal is written by its mirror
            const bool ip = gi >= p.c0 && gi < p.c0 + p.nb;
            const bool jp = gq >= p.c0 && gq < p.c0 + p.nb;
            double r;
            if (!ip && !jp) r = Cz[(size_t)gi * p.ldc + gq] - acc[i][j][q];
            else if (jp) r = p.P[z * p.sP + (size_t)gi * p.ldp + (gq - p.c0)];
            else r = p.P[z * p.sP + (size_t)gq * p.ldp + (gi - p.c0)];
            r *= p.sign;
            Cz[(size_t)gi * p.ldc + gq] = r;
            if (gq < gi) Cz[(size_t)gq * p.ldc + gi] = r;
          }
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

}  // namespace

// V: [Z][rows][ldv] with ldv == K (the V_j of one outer block side by side), C: Z symmetric N x N matrices, P: the finished
// panel [Z][rows][ldp]. Returns false when this path does not apply (the caller then takes the cp.async kernel).
bool sweep_update_tma(cudaStream_t s, int N, int K, const double* V, int ldv, int rows, int Z, double* C, int ldc,
                      long long sC, const double* P, int ldp, long long sP, int c0, int nb, double sign, cudaError_t* err) {
  static const bool off = getenv("MCB_NO_TMA") != nullptr;
  *err = cudaSuccess;
  // 128 x 128 tiles on 148 persistent CTAs: below ~8 tiles per CTA the last, partly filled round costs more than the
  // pipelining saves (measured at N = 4096: 8.13 ms per inversion against 7.90 ms with the 64 x 64 cp.async kernel; N = 8192:
  // 33.68 against 34.77 ms, C4 E-step 253 against 263 ms; before the L2 prefetch of the C tile and the deeper epilogue
  // batches the two kernels were equal at N = 8192, 34.97 / 34.74 ms)
  static const int min_n = getenv("MCB_TMA_MIN_N") ? atoi(getenv("MCB_TMA_MIN_N")) : 6144;
  if (off || K % kStageCols != 0 || ldv != K || N < 4 * kTile || N < min_n || (reinterpret_cast<unsigned long long>(V) & 15ull))
    return false;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)ldv, (cuuint64_t)rows * (cuuint64_t)Z};
  const cuuint64_t strides[1] = {(cuuint64_t)ldv * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)kStageCols, (cuuint32_t)kTile};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(V), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  UpdParams p{};
  p.N = N; p.K = K; p.rows = rows;
  p.C = C; p.ldc = ldc; p.sC = sC;
  p.P = P; p.sP = sP; p.ldp = ldp; p.c0 = c0; p.nb = nb;
  p.sign = sign;
  p.ntile = (N + kTile - 1) / kTile;
  p.ntri = p.ntile * (p.ntile + 1) / 2;
  p.total = p.ntri * Z;
  static const int stagger = getenv("MCB_TMA_STAGGER_NS") ? atoi(getenv("MCB_TMA_STAGGER_NS")) : 0;
  p.stagger_ns = stagger;
  const size_t smem = (size_t)kStages * kStageBytes + 2 * kStages * sizeof(unsigned long long) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if ((*err = cudaFuncSetAttribute(sweep_update_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return true;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.total < sms ? p.total : sms;
  static const int dbg_sync = getenv("MCB_TMA_SYNC") ? atoi(getenv("MCB_TMA_SYNC")) : 0;   // debug: 1 before, 2 after, 3 both
  if (dbg_sync & 1) cudaStreamSynchronize(s);
  sweep_update_tma_kernel<<<grid, kThreads, smem, s>>>(map, p);
  if (dbg_sync & 2) cudaStreamSynchronize(s);
  *err = cudaGetLastError();
  return true;
}

}  // namespace mcb
