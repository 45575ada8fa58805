// Complex128 GEMM on the FP64 tensor pipe (DMMA.8x8x4) for sm_100a.
//
// This is the engine behind every pairwise contraction of the path that is not the fused stage-3 matvec:
// NDArrayData.contractWith == numpy.tensordot -> OpenBLAS zgemm in the reference (data/__init__.py:157-159);
// here: stage-1 / stage-2 environment builds, side->corner and center->side absorption, formMatrix, the
// compression normal equations, absorbMatrixAt.
//
// Layout: CTA tile 128 x 64 x 16 (complex), 8 consumer warps as 4 (M) x 2 (N), warp tile 32 x 32 = 4 x 4 DMMA tiles with
// separate real / imaginary accumulators (128 registers).  Operands are staged into K-contiguous shared tiles with an
// odd row stride (17 complex) so that the paired fragment loads below are bank-conflict free.
// Staging is warp-specialised (template WS = true, the default): a fourth warpgroup of 4 PRODUCER warps computes the
// generalised source addresses and issues the cp.async copies of a k-tile (their completion arrives on the stage's
// "full" mbarrier: cp.async.mbarrier.arrive.noinc), the 8 consumer warps wait -> DMMA -> release ("empty" mbarrier);
// `setmaxnreg` gives the consumers 216 registers and leaves the producers 72; 4 stages.  The symmetric variant
// (WS = false, CARC_ZGEMM_WS=0: every warp copies and multiplies, a CTA-wide barrier per k-tile, 3 stages) is what round 1
// measured at 27.0 TFLOP/s on 4096^3.  A complex MMA is four real DMMAs; the sign of the imaginary operand is flipped once
// per fragment load (integer XOR on the sign bit), which is also how conjugated operands are handled.
//
// Fragment mapping: one 8-wide K chunk feeds two DMMA k-steps.  k-step e in {0,1} uses the K indices
// {2c + e : c = lane % 4}, i.e. every lane reads two adjacent complex numbers (32 contiguous bytes) per
// operand row.  Any permutation of K is legal as long as A and B use the same one.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, LDK = BK + 1;
constexpr int NTHREADS = 256;
constexpr int SMEM_BYTES = STAGES * (BM + BN) * LDK * (int)sizeof(cplx);
constexpr int STAGES_WS = 4, NPRODUCERS = 128, NTHREADS_WS = NTHREADS + NPRODUCERS;
constexpr int SMEM_BYTES_WS = STAGES_WS * (BM + BN) * LDK * (int)sizeof(cplx) + 2 * STAGES_WS * 8;

struct GemmParams {
  const cplx* A;
  const cplx* B;
  cplx* C;
  int64_t M, N, K, lda, ldb;
  int64_t strideA, strideB, strideC;
  int a_kcontig, b_kcontig;   // 1: K is the unit-stride axis of the stored operand
  int64_t a_kdiv, a_ks1, b_kdiv, b_ks1;  // optional two-level K for K-contiguous operands (kdiv == 0: off):
                                         // element k lives at (k / kdiv) * ks1 + (k % kdiv) within its row
  uint32_t a_sign, b_sign;    // 0x80000000 to conjugate
  cplx alpha, beta;
  GemmOut out;
  const int64_t* rowoff;      // optional per-row / per-column output offset tables (override `out`)
  const int64_t* coloff;
  int hermitian;              // 1: C = C^H (M == N): tiles strictly below the diagonal are skipped and mirrored
  unsigned tiles_m;           // number of M tiles (the tile grid is linearised on blockIdx.x)
  unsigned tiles_n, group_m;  // the linear order walks `group_m` M tiles at a time through all N tiles, so that the ~148 CTAs
                              // running together cover a compact block of the output and share operand tiles in L2
  int splitk;                 // > 1: blockIdx.z is a K split; raw partial products go to C = workspace[split][M][N]
  int64_t kt_per_split;       // K tiles per split
};

__device__ __forceinline__ double flip(double x, uint32_t mask) {
  return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

__device__ __forceinline__ void cp_async_arrive_on(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}

template <bool WS>
__global__ void __launch_bounds__(WS ? NTHREADS_WS : NTHREADS, 1) zgemm_kernel(GemmParams p) {
  constexpr int NST = WS ? STAGES_WS : STAGES;
  constexpr int NLOAD = WS ? NPRODUCERS : NTHREADS;   // threads that issue the copies of a k-tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* As = reinterpret_cast<cplx*>(smem_raw);
  cplx* Bs = As + NST * BM * LDK;
  const uint32_t bars = smem_u32(Bs + NST * BN * LDK);   // WS: full[NST], empty[NST]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;       // 4 x 2 warps
  const int r = lane >> 2, c = lane & 3;
  // 1-D tile grid in grouped order (a wave of CTAs reads ~2 sqrt(148) distinct operand tiles per k-step instead of
  // tiles_m + 148 / tiles_m: the Gram products of the compression re-read their 4.3 GB operand 20 x from DRAM before)
  const unsigned width = p.group_m * p.tiles_n, grp = blockIdx.x / width, first_m = grp * p.group_m;
  const unsigned gsz = min(p.tiles_m - first_m, p.group_m), within = blockIdx.x - grp * width;
  const int64_t m0 = (int64_t)(first_m + within % gsz) * BM, n0 = (int64_t)(within / gsz) * BN;
  if (p.hermitian == 1 && n0 + BN - 1 < m0) return;   // strictly lower tile: filled by the mirror of its transpose
  if (p.hermitian == 2 && n0 > m0 + BM - 1) return;   // lower-triangle update: tiles strictly above the diagonal skipped
  const bool split = p.splitk > 1;
  const cplx* A = p.A + (split ? 0 : (int64_t)blockIdx.z * p.strideA);
  const cplx* B = p.B + (split ? 0 : (int64_t)blockIdx.z * p.strideB);
  cplx* C = p.C + (int64_t)blockIdx.z * p.strideC;

  CTile acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j].zero();

  const int64_t KT_all = (p.K + BK - 1) / BK;
  const int64_t kt0 = split ? (int64_t)blockIdx.z * p.kt_per_split : 0;
  const int64_t KT = split ? min(p.kt_per_split, KT_all - kt0) : KT_all;

  const int ltid = WS ? tid - NTHREADS : tid;
  auto load_tile = [&](int64_t kt, int stage) {
    const int64_t k0 = (kt0 + kt) * BK;
    cplx* as = As + stage * BM * LDK;
    cplx* bs = Bs + stage * BN * LDK;
#pragma unroll
    for (int it = 0; it < BM * BK / NLOAD; ++it) {
      int e = it * NLOAD + ltid;
      int m, k;
      if (p.a_kcontig) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
      int64_t gm = m0 + m, gk = k0 + k;
      bool ok = gm < p.M && gk < p.K;
      if (p.a_kdiv) gk = (gk / p.a_kdiv) * p.a_ks1 + (gk % p.a_kdiv);
      const cplx* src = ok ? (p.a_kcontig ? A + gm * p.lda + gk : A + gk * p.lda + gm) : A;
      cp_async16(smem_u32(as + m * LDK + k), src, ok);
    }
#pragma unroll
    for (int it = 0; it < BN * BK / NLOAD; ++it) {
      int e = it * NLOAD + ltid;
      int n, k;
      if (p.b_kcontig) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      int64_t gn = n0 + n, gk = k0 + k;
      bool ok = gn < p.N && gk < p.K;
      if (p.b_kdiv) gk = (gk / p.b_kdiv) * p.b_ks1 + (gk % p.b_kdiv);
      const cplx* src = ok ? (p.b_kcontig ? B + gn * p.ldb + gk : B + gk * p.ldb + gn) : B;
      cp_async16(smem_u32(bs + n * LDK + k), src, ok);
    }
  };

  // one k-tile of the warp's 32 x 32 block
  auto multiply_tile = [&](int stage) {
    const uint32_t as = smem_u32(As + stage * BM * LDK + (wm * 32 + r) * LDK + 2 * c);
    const uint32_t bs = smem_u32(Bs + stage * BN * LDK + (wn * 32 + r) * LDK + 2 * c);
#pragma unroll
    for (int k8 = 0; k8 < BK / 8; ++k8) {
      cplx b[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        b[j][0] = lds_c(bs + (j * 8 * LDK + k8 * 8) * 16);
        b[j][1] = lds_c(bs + (j * 8 * LDK + k8 * 8 + 1) * 16);
        b[j][0].y = flip(b[j][0].y, p.b_sign);
        b[j][1].y = flip(b[j][1].y, p.b_sign);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cplx a0 = lds_c(as + (i * 8 * LDK + k8 * 8) * 16);
        cplx a1 = lds_c(as + (i * 8 * LDK + k8 * 8 + 1) * 16);
        a0.y = flip(a0.y, p.a_sign);
        a1.y = flip(a1.y, p.a_sign);
        const double na0 = -a0.y, na1 = -a1.y;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cmma(acc[i][j], a0.x, a0.y, na0, b[j][0].x, b[j][0].y);
          cmma(acc[i][j], a1.x, a1.y, na1, b[j][1].x, b[j][1].y);
        }
      }
    }
  };

  if (WS) {
    if (tid == 0) {
      for (int s = 0; s < NST; ++s) {
        mbar_init(bars + s * 8, NPRODUCERS);          // full: one arrival per producer thread, when its copies have landed
        mbar_init(bars + (NST + s) * 8, NTHREADS / 32);   // empty: one arrival per consumer warp
      }
      fence_barrier_init();
    }
    __syncthreads();
    if (warp >= NTHREADS / 32) {
      // ---- producers (72 registers: the generalised addresses are 64-bit arithmetic)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
      int stage = 0, parity = 1;     // parity of the EMPTY barrier's phase that frees the stage (first round: free already)
      for (int64_t kt = 0; kt < KT; ++kt) {
        if (kt >= NST) mbar_wait(bars + (NST + stage) * 8, (uint32_t)parity);
        load_tile(kt, stage);
        cp_async_arrive_on(bars + stage * 8);
        if (++stage == NST) {
          stage = 0;
          parity ^= 1;
        }
      }
      cp_async_wait<0>();
      return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
    int stage = 0, parity = 0;
    for (int64_t kt = 0; kt < KT; ++kt) {
      mbar_wait(bars + stage * 8, (uint32_t)parity);
      multiply_tile(stage);
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + (NST + stage) * 8);
      if (++stage == NST) {
        stage = 0;
        parity ^= 1;
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
      if (s < KT) load_tile(s, s);
      cp_async_commit();
    }
    for (int64_t kt = 0; kt < KT; ++kt) {
      cp_async_wait<NST - 2>();
      __syncthreads();
      {
        int64_t nk = kt + NST - 1;
        if (nk < KT) load_tile(nk, (int)(nk % NST));
        cp_async_commit();
      }
      multiply_tile((int)(kt % NST));
    }
    cp_async_wait<0>();
  }

  // epilogue: C = alpha * acc + beta * C at the generalised address
  const bool use_beta = !(p.beta.x == 0.0 && p.beta.y == 0.0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm * 32 + i * 8 + r;
    if (m >= p.M) continue;
    const int64_t moff = p.rowoff ? p.rowoff[m] : (m / p.out.m_div) * p.out.m_s1 + (m % p.out.m_div) * p.out.m_s0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t n = n0 + wn * 32 + j * 8 + 2 * c + h;
        if (n >= p.N) continue;
        if (p.hermitian == 1 && n < m) continue;   // written (bit-identically conjugated) by the mirror of (n, m)
        const int64_t off =
            moff + (p.coloff ? p.coloff[n] : (n / p.out.n_div) * p.out.n_s1 + (n % p.out.n_div) * p.out.n_s0);
        const double re = h ? acc[i][j].re1 : acc[i][j].re0;
        const double im = h ? acc[i][j].im1 : acc[i][j].im0;
        cplx v;
        v.x = p.alpha.x * re - p.alpha.y * im;
        v.y = p.alpha.x * im + p.alpha.y * re;
        if (use_beta) {
          cplx o = C[off];
          v.x += p.beta.x * o.x - p.beta.y * o.y;
          v.y += p.beta.x * o.y + p.beta.y * o.x;
        }
        C[off] = v;
        if (p.hermitian == 1 && m != n) C[n * p.N + m] = make_double2(v.x, -v.y);   // plain row-major only
      }
    }
  }
}

// split-K second pass: sum the partial products in split order (deterministic) and apply the real epilogue
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const cplx* __restrict__ ws, int splits, GemmParams p) {
  const int64_t total = p.M * p.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int s2 = 0; s2 < splits; ++s2) {
      const cplx v = ws[(int64_t)s2 * total + i];
      re += v.x;
      im += v.y;
    }
    const int64_t m = i / p.N, n = i - m * p.N;
    const int64_t off = (p.rowoff ? p.rowoff[m] : (m / p.out.m_div) * p.out.m_s1 + (m % p.out.m_div) * p.out.m_s0) +
                        (p.coloff ? p.coloff[n] : (n / p.out.n_div) * p.out.n_s1 + (n % p.out.n_div) * p.out.n_s0);
    cplx v;
    v.x = p.alpha.x * re - p.alpha.y * im;
    v.y = p.alpha.x * im + p.alpha.y * re;
    if (!(p.beta.x == 0.0 && p.beta.y == 0.0)) {
      const cplx o = p.C[off];
      v.x += p.beta.x * o.x - p.beta.y * o.y;
      v.y += p.beta.x * o.y + p.beta.y * o.x;
    }
    p.C[off] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// DMMA issue-rate microbenchmark: every warp keeps 16 independent accumulator tiles in flight.
__global__ void __launch_bounds__(256, 1) dmma_peak_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c0[16], c1[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) c0[j] = c1[j] = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dmma(c0[j], c1[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += c0[j] + c1[j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// table[i] = sum_l digit_l(i) * stride_l, digits of i taken row-major over `extent` (last level fastest)
struct IndexLevels {
  int n;
  int64_t extent[CARC_MAX_RANK], stride[CARC_MAX_RANK];
};
__global__ void __launch_bounds__(256) index_table_kernel(IndexLevels lv, int64_t total, int64_t* __restrict__ table) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t rem = i, off = 0;
  for (int l = lv.n - 1; l >= 0; --l) {
    const int64_t q = rem / lv.extent[l];
    off += (rem - q * lv.extent[l]) * lv.stride[l];
    rem = q;
  }
  table[i] = off;
}

template <int CH>
__global__ void __launch_bounds__(1024, 1) dmma_rate_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c0[CH], c1[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) c0[j] = c1[j] = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CH; ++j) dmma(c0[j], c1[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += c0[j] + c1[j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same with distinct A / B operand registers per instruction (4 x 4 outer product of fragments, the register
// traffic a real GEMM inner loop has), to separate issue-rate limits from operand-fetch limits
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) dmma_rate_distinct_kernel(double* out, int iters) {
  double a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = 1.0 + (threadIdx.x + j) * 1e-9;
    b[j] = 1.0 - (threadIdx.x + 2 * j) * 1e-9;
  }
  double c0[16], c1[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) c0[j] = c1[j] = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dmma(c0[j], c1[j], a[j >> 2], b[j & 3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // keep the operands changing like fragments reloaded every k-step
      a[j] += 1e-12;
      b[j] -= 1e-12;
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += c0[j] + c1[j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA side by side: ND DMMA.8x8x4 (16 independent accumulator tiles) and NF DFMA (16 independent chains) per
// loop trip and warp -- do the FP64 tensor instruction and the FP64 FMA pipe share execution resources on this part?
template <int ND, int NF>
__global__ void __launch_bounds__(512, 1) fp64_mix_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c0[16], c1[16], f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    c0[j] = c1[j] = 0.0;
    f[j] = j * 1e-3;
  }
  for (int i = 0; i < iters; ++i) {
    constexpr int STEPS = ND > NF ? ND : NF;
#pragma unroll
    for (int j = 0; j < STEPS; ++j) {
      // interleave the two instruction streams evenly
      if ((j + 1) * ND / STEPS != j * ND / STEPS) dmma(c0[(j * ND / STEPS) & 15], c1[(j * ND / STEPS) & 15], a, b);
      if ((j + 1) * NF / STEPS != j * NF / STEPS)
        asm volatile("fma.rn.f64 %0, %1, %2, %0;\n" : "+d"(f[(j * NF / STEPS) & 15]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += c0[j] + c1[j] + f[j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// DMMA issue rate with `warps` warps per SM (one CTA per SM) and `chains` independent accumulators per warp
int dmma_rate(int iters, int warps, int chains, double* tflops_out, cudaStream_t stream) {
  double* buf = nullptr;
  CARC_CHECK_CUDA(cudaMalloc(&buf, sizeof(double) * 148 * 1024));
  cudaEvent_t e0, e1;
  CARC_CHECK_CUDA(cudaEventCreate(&e0));
  CARC_CHECK_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CARC_CHECK_CUDA(cudaEventRecord(e0, stream));
    if (chains >= 100 && warps <= 8) dmma_rate_distinct_kernel<256><<<148, warps * 32, 0, stream>>>(buf, iters);
    else if (chains >= 100) dmma_rate_distinct_kernel<512><<<148, warps * 32, 0, stream>>>(buf, iters);
    else if (chains <= 2) dmma_rate_kernel<2><<<148, warps * 32, 0, stream>>>(buf, iters);
    else if (chains <= 4) dmma_rate_kernel<4><<<148, warps * 32, 0, stream>>>(buf, iters);
    else if (chains <= 8) dmma_rate_kernel<8><<<148, warps * 32, 0, stream>>>(buf, iters);
    else dmma_rate_kernel<16><<<148, warps * 32, 0, stream>>>(buf, iters);
    CARC_CHECK_CUDA(cudaEventRecord(e1, stream));
    CARC_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CARC_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const int ch = chains >= 100 ? 16 : chains <= 2 ? 2 : chains <= 4 ? 4 : chains <= 8 ? 8 : 16;
  *tflops_out = 148.0 * warps * iters * (double)ch * 512.0 / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return CARC_OK;
}


// Rates of a DMMA + DFMA instruction mix (per loop trip and warp: `ndmma` in {0, 16} DMMA.8x8x4 and `nfma` in
// {0, 16, 32, 64, 128} DFMA), reported separately in TFLOP/s.
int fp64_mix_rate(int iters, int warps, int ndmma, int nfma, double* tflops_dmma, double* tflops_fma, cudaStream_t stream) {
  CARC_REQUIRE(warps >= 1 && warps <= 16 && (ndmma == 0 || ndmma == 16), CARC_ERR_VALUE, "fp64_mix_rate: invalid argument");
  CARC_REQUIRE(nfma == 0 || nfma == 16 || nfma == 32 || nfma == 64 || nfma == 128, CARC_ERR_VALUE,
               "fp64_mix_rate: nfma must be 0, 16, 32, 64 or 128");
  CARC_REQUIRE(ndmma + nfma > 0, CARC_ERR_VALUE, "fp64_mix_rate: empty mix");
  double* buf = nullptr;
  CARC_CHECK_CUDA(cudaMalloc(&buf, sizeof(double) * 148 * 512));
  cudaEvent_t e0, e1;
  CARC_CHECK_CUDA(cudaEventCreate(&e0));
  CARC_CHECK_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CARC_CHECK_CUDA(cudaEventRecord(e0, stream));
#define CARC_MIX(ND, NF) \
  if (ndmma == ND && nfma == NF) fp64_mix_kernel<ND, NF><<<148, warps * 32, 0, stream>>>(buf, iters);
    CARC_MIX(16, 0) CARC_MIX(16, 16) CARC_MIX(16, 32) CARC_MIX(16, 64) CARC_MIX(16, 128)
    CARC_MIX(0, 16) CARC_MIX(0, 32) CARC_MIX(0, 64) CARC_MIX(0, 128)
#undef CARC_MIX
    CARC_CHECK_CUDA(cudaEventRecord(e1, stream));
    CARC_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CARC_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CARC_CHECK_CUDA(cudaGetLastError());
  *tflops_dmma = 148.0 * warps * iters * (double)ndmma * 512.0 / (best * 1e-3) / 1e12;
  *tflops_fma = 148.0 * warps * iters * (double)nfma * 64.0 / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return CARC_OK;
}

int index_table(int nlevels, const int64_t* extents, const int64_t* strides, int64_t* table, cudaStream_t stream) {
  CARC_REQUIRE(nlevels >= 0 && nlevels <= CARC_MAX_RANK, CARC_ERR_RANK, "index_table: %d levels unsupported", nlevels);
  IndexLevels lv;
  lv.n = nlevels;
  int64_t total = 1;
  for (int l = 0; l < nlevels; ++l) {
    CARC_REQUIRE(extents[l] > 0, CARC_ERR_VALUE, "index_table: non-positive extent");
    lv.extent[l] = extents[l];
    lv.stride[l] = strides[l];
    total *= extents[l];
  }
  index_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(lv, total, table);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

static thread_local bool g_hermitian = false;
static thread_local bool g_lower_only = false;

// C (square, any output map) = alpha op(A) op(B) + beta C on the tiles that touch the lower triangle only: the
// Hermitian rank-k update of the blocked Cholesky factorisation (solver.cu), half the tensor work of the full update.
int zgemm_lower(int opA, int opB, int64_t N, int64_t K, cplx alpha, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                cplx beta, cplx* C, const GemmOut* out, cudaStream_t stream) {
  g_lower_only = true;
  int rc = zgemm(opA, opB, N, N, K, alpha, A, lda, B, ldb, beta, C, out, nullptr, 1, 0, 0, 0, stream);
  g_lower_only = false;
  return rc;
}

// C = op(A) op(B) known to be Hermitian (Gram matrices): computes the upper triangle's tiles and mirrors them.
int zgemm_hermitian(int opA, int opB, int64_t N, int64_t K, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                    cplx* C, cudaStream_t stream) {
  g_hermitian = true;
  int rc = zgemm(opA, opB, N, N, K, make_double2(1.0, 0.0), A, lda, B, ldb, make_double2(0.0, 0.0), C, nullptr, nullptr, 1,
                 0, 0, 0, stream);
  g_hermitian = false;
  return rc;
}

int zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, cplx alpha, const cplx* A, int64_t lda, const cplx* B,
          int64_t ldb, cplx beta, cplx* C, const GemmOut* out, const GemmKMap* kmap, int64_t batch, int64_t strideA,
          int64_t strideB, int64_t strideC, cudaStream_t stream, const int64_t* rowoff, const int64_t* coloff) {
  CARC_REQUIRE(opA >= 0 && opA <= 3 && opB >= 0 && opB <= 3, CARC_ERR_VALUE, "zgemm: invalid op");
  CARC_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, CARC_ERR_VALUE, "zgemm: negative dimension");
  if (M == 0 || N == 0 || batch == 0) return CARC_OK;
  static bool configured[16] = {false};
  static const bool ws = !(getenv("CARC_ZGEMM_WS") && atoi(getenv("CARC_ZGEMM_WS")) == 0);   // 0: the symmetric kernel
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(zgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CARC_CHECK_CUDA(cudaFuncSetAttribute(zgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_WS));
    configured[dev] = true;
  }
  // CARC_ZGEMM_TRACE=1 (diagnostics): every product is timed with events and logged to stderr as
  // "zgemm M N K batch opA opB splits ms" -- scripts/zgemm_trace.py sums the log by shape
  static const bool trace = getenv("CARC_ZGEMM_TRACE") && atoi(getenv("CARC_ZGEMM_TRACE")) == 1;
  auto launch = [&](dim3 grid, const GemmParams& q) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (trace) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, stream);
    }
    if (ws) zgemm_kernel<true><<<grid, NTHREADS_WS, SMEM_BYTES_WS, stream>>>(q);
    else zgemm_kernel<false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(q);
    if (trace) {
      cudaEventRecord(e1, stream);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      fprintf(stderr, "zgemm %lld %lld %lld %lld %d %d %d %.4f\n", (long long)M, (long long)N, (long long)K, (long long)batch, opA,
              opB, q.splitk, ms);
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
  };
  GemmParams p;
  p.A = A; p.B = B; p.C = C;
  p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb;
  p.strideA = strideA; p.strideB = strideB; p.strideC = strideC;
  p.a_kcontig = (opA == OP_N || opA == OP_J);
  p.b_kcontig = (opB == OP_T || opB == OP_C);
  p.a_sign = (opA == OP_C || opA == OP_J) ? 0x80000000u : 0u;
  p.b_sign = (opB == OP_C || opB == OP_J) ? 0x80000000u : 0u;
  p.alpha = alpha; p.beta = beta;
  p.rowoff = rowoff; p.coloff = coloff;
  p.hermitian = 0;
  p.a_kdiv = p.a_ks1 = p.b_kdiv = p.b_ks1 = 0;
  if (kmap) {
    CARC_REQUIRE((!kmap->a_kdiv || p.a_kcontig) && (!kmap->b_kdiv || p.b_kcontig), CARC_ERR_VALUE,
                 "zgemm: a two-level K map needs a K-contiguous operand");
    p.a_kdiv = kmap->a_kdiv; p.a_ks1 = kmap->a_ks1; p.b_kdiv = kmap->b_kdiv; p.b_ks1 = kmap->b_ks1;
  }
  if (out) {
    p.out = *out;
  } else {
    p.out.m_div = M > 0 ? M : 1; p.out.m_s1 = 0; p.out.m_s0 = N;
    p.out.n_div = N > 0 ? N : 1; p.out.n_s1 = 0; p.out.n_s0 = 1;
  }
  CARC_REQUIRE(p.out.m_div > 0 && p.out.n_div > 0, CARC_ERR_VALUE, "zgemm: invalid output map");
  int64_t gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  CARC_REQUIRE(batch < 65536 && gx * gy < (1ll << 31), CARC_ERR_VALUE, "zgemm: grid too large (%lld tiles, batch %lld)",
               (long long)(gx * gy), (long long)batch);
  p.tiles_m = (unsigned)gy;
  p.tiles_n = (unsigned)gx;
  p.group_m = (unsigned)std::min<int64_t>(gy, 12);
  p.splitk = 1;
  p.kt_per_split = 0;
  const int64_t tiles = gx * gy, KT = (K + BK - 1) / BK;
  const bool herm = g_hermitian && M == N && !out && !rowoff && !coloff && batch == 1;
  if (herm && !(tiles <= 74 && KT >= 64)) p.hermitian = 1;   // (small outputs take the split-K path instead)
  if (g_lower_only && M == N && batch == 1 && !(tiles <= 74 && KT >= 64)) p.hermitian = 2;
  if (batch == 1 && tiles <= 74 && KT >= 64) {
    // few output tiles, long K (Gram matrices, formMatrix): split K over the idle SMs
    int64_t splits = std::min<int64_t>(148 * 2 / tiles, KT / 16);
    splits = std::min<int64_t>(splits, 64);
    if (splits >= 2) {
      const int64_t kps = (KT + splits - 1) / splits;
      splits = (KT + kps - 1) / kps;
      cplx* ws = nullptr;
      CARC_CHECK_CUDA(cudaMallocAsync((void**)&ws, sizeof(cplx) * (size_t)splits * M * N, stream));
      GemmParams q = p;
      q.splitk = (int)splits;
      q.kt_per_split = kps;
      q.C = ws;
      q.strideC = M * N;
      q.alpha = make_double2(1.0, 0.0);
      q.beta = make_double2(0.0, 0.0);
      q.rowoff = q.coloff = nullptr;
      q.out.m_div = M; q.out.m_s1 = 0; q.out.m_s0 = N;
      q.out.n_div = N; q.out.n_s1 = 0; q.out.n_s0 = 1;
      dim3 grid((unsigned)(gx * gy), 1, (unsigned)splits);
      launch(grid, q);
      const int64_t total = M * N;
      splitk_reduce_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, stream>>>(ws, (int)splits, p);
      CARC_CHECK_CUDA(cudaGetLastError());
      CARC_CHECK_CUDA(cudaFreeAsync(ws, stream));
      return CARC_OK;
    }
  }
  dim3 grid((unsigned)(gx * gy), 1, (unsigned)batch);
  launch(grid, p);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int dmma_peak(int iters, double* tflops_out, cudaStream_t stream) {
  double* buf = nullptr;
  const int blocks = 148 * 2;
  CARC_CHECK_CUDA(cudaMalloc(&buf, sizeof(double) * blocks * 256));
  cudaEvent_t e0, e1;
  CARC_CHECK_CUDA(cudaEventCreate(&e0));
  CARC_CHECK_CUDA(cudaEventCreate(&e1));
  dmma_peak_kernel<<<blocks, 256, 0, stream>>>(buf, 16);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CARC_CHECK_CUDA(cudaEventRecord(e0, stream));
    dmma_peak_kernel<<<blocks, 256, 0, stream>>>(buf, iters);
    CARC_CHECK_CUDA(cudaEventRecord(e1, stream));
    CARC_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CARC_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // one DMMA.8x8x4 = 8*8*4 FMA = 512 flop per warp
  double flops = (double)blocks * 8.0 * iters * 16.0 * 512.0;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return CARC_OK;
}

}  // namespace carc
