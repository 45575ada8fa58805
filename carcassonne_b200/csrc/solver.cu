// Device-resident center-site eigen-solver: restarted Arnoldi on N^-1 H ("relaxOver", reference utils.py:805-878),
// with N^-1 applied either by a pivoted LU of the dense normalization matrix (scipy.linalg.lu_factor / lu_solve,
// utils.py:816-818) or by restarted GMRES on the normalization operator (scipy.sparse.linalg.gmres, utils.py:819-825).
//
// The state vector, the Krylov basis and every scalar of the iteration stay in HBM: the host only enqueues kernels
// and reads back one small record per restart (the Ritz value, to evaluate the stopping rule) -- the eigenvector
// never leaves the device.  Vector algebra (classical Gram-Schmidt projections, norms, the k x k projected matrix,
// the Ritz combination) runs in fused single-CTA kernels with warp-shuffle reductions: the vectors are N = D^4 d
// long (128 KiB at D = 8) and live in L2, so these kernels are latency-bound and one CTA avoids grid-wide
// synchronisation; the k x k non-Hermitian eigenproblem (scipy.linalg.eig -> zgeev in the reference) is solved
// by thread 0 of the same kernel with a shifted QR iteration.
#include <cooperative_groups.h>
#include <math.h>

#include <cstdlib>

#include <atomic>
#include <vector>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

namespace {

constexpr int VT = 1024;      // threads of the single-CTA vector kernels
constexpr int KMAX = 8;       // largest Krylov dimension supported on device
constexpr int GM_MAX = 64;    // largest GMRES restart length

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// block-wide sum of up to NV complex values per thread (fixed order -> deterministic); result valid in all threads
template <int NV>
__device__ __forceinline__ void block_csum(cplx (&v)[NV], cplx* sh /* [32 * NV] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    v[q].x = warp_sum(v[q].x);
    v[q].y = warp_sum(v[q].y);
  }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NV; ++q) sh[w * NV + q] = v[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    cplx t = make_double2(0.0, 0.0);
    for (int i = 0; i < nw; ++i) t = cadd(t, sh[i * NV + q]);
    v[q] = t;
  }
}

// ---------------------------------------------------------------------------------------------------
// dense operator: y = M x, M [rows, cols] row-major; one warp per row
__global__ void __launch_bounds__(256) zgemv_kernel(const cplx* __restrict__ M, int64_t rows, int64_t cols, int64_t ld,
                                                    const cplx* __restrict__ x, cplx* __restrict__ y, cplx alpha,
                                                    cplx beta) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const cplx* m = M + row * ld;
  cplx acc = make_double2(0.0, 0.0);
  for (int64_t c = lane; c < cols; c += 32) acc = cadd(acc, cmul(m[c], x[c]));
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  if (lane == 0) {
    cplx r = cmul(alpha, acc);
    if (beta.x != 0.0 || beta.y != 0.0) r = cadd(r, cmul(beta, y[row]));
    y[row] = r;
  }
}





// Whole-panel factorisation in ONE cooperative launch (one CTA per SM): the column steps are separated by grid-wide
// barriers instead of kernel boundaries (two per column; a launch pair costs ~14 us end to end on this stack, a grid
// barrier ~2 us).  Per column j: [reduce the pivot candidates tracked during the previous update; interchange rows
// j <-> p across the whole matrix, each CTA a slice of columns] | barrier | [scale column j, rank-1 update of the
// rest of the current 8-column sub-panel, tracking the pivot candidates of column j + 1] | barrier.  At the end of a
// sub-panel the remaining panel columns receive the U block row (tiny triangular solve) and one rank-8 update.
struct LuCand {
  double val[160];
  int idx[160];
};

__global__ void __launch_bounds__(256, 1) lu_panel_coop_kernel(cplx* __restrict__ A, int n, int j0, int nb,
                                                               int* __restrict__ piv, int* __restrict__ singular,
                                                               LuCand* __restrict__ cand) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_p;
  __shared__ cplx s_U[8][64];
  __shared__ cplx s_L[8][8];
  __shared__ cplx s_row[8];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int pe = j0 + nb;
  const int nctas = gridDim.x, NT = blockDim.x, NW = blockDim.x >> 5;

  auto publish = [&](double best, int bi) {   // block argmax of lane-0 candidates -> cand[blockIdx]
    if (tx == 0) { s_val[ty] = best; s_idx[ty] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = s_val[0];
      int p = s_idx[0];
      for (int q = 1; q < NW; ++q)
        if (s_val[q] > b || (s_val[q] == b && s_idx[q] < p)) { b = s_val[q]; p = s_idx[q]; }
      cand->val[blockIdx.x] = b;
      cand->idx[blockIdx.x] = p;
    }
    __syncthreads();
  };

  // candidates of the first column of the panel: plain scan
  {
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = j0 + blockIdx.x * NT + tid; i < n; i += nctas * NT) {
      const cplx a = A[(int64_t)i * n + j0];
      const double m = fabs(a.x) + fabs(a.y);
      if (m > best || (m == best && i < bi)) { best = m; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    publish(best, bi);
  }
  grid.sync();

  for (int s0 = j0; s0 < pe; s0 += 8) {
    const int w = pe - s0 < 8 ? pe - s0 : 8;
    const int c_sub = s0 + w;
    for (int j = s0; j < c_sub; ++j) {
      // ---- pivot of column j from the published candidates (every CTA redundantly), row interchange
      if (tid < 32) {
        double b = -1.0;
        int p = 0x7fffffff;
        for (int q = tid; q < nctas; q += 32) {
          const double v = cand->val[q];
          const int idx = cand->idx[q];
          if (v > b || (v == b && idx < p)) { b = v; p = idx; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, b, o);
          const int oi = __shfl_xor_sync(0xffffffffu, p, o);
          if (ob > b || (ob == b && oi < p)) { b = ob; p = oi; }
        }
        if (tid == 0) {
          if (p == 0x7fffffff) p = j;
          s_p = p;
          if (blockIdx.x == 0) {
            piv[j] = p;
            if (!(b > 0.0)) *singular = 1;
          }
        }
      }
      __syncthreads();
      const int p = s_p;
      if (p != j) {
        for (int c = blockIdx.x * NT + tid; c < n; c += nctas * NT) {
          const cplx a = A[(int64_t)j * n + c], b2 = A[(int64_t)p * n + c];
          A[(int64_t)j * n + c] = b2;
          A[(int64_t)p * n + c] = a;
        }
      }
      grid.sync();
      // ---- scale column j, rank-1 update inside the sub-panel, candidates of column j + 1
      const cplx d = A[(int64_t)j * n + j];
      const cplx inv = (d.x != 0.0 || d.y != 0.0) ? cdiv(make_double2(1.0, 0.0), d) : make_double2(0.0, 0.0);
      const bool track = j + 1 < pe;
      double best = -1.0;
      int bi = 0x7fffffff;
      // one thread per row: the <= 8 sub-panel entries of a row are contiguous, all rows' loads are in flight at once
      const int wcols = c_sub - j - 1;
      if (tid < wcols) s_row[tid] = A[(int64_t)j * n + j + 1 + tid];
      __syncthreads();
      for (int i = j + 1 + blockIdx.x * NT + tid; i < n; i += nctas * NT) {
        cplx* row = A + (int64_t)i * n + j;
        cplx e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q <= wcols) e[q] = row[q];
        const cplx l = cmul(e[0], inv);
        row[0] = l;
#pragma unroll
        for (int q = 1; q < 8; ++q)
          if (q <= wcols) {
            const cplx v = csub(e[q], cmul(l, s_row[q - 1]));
            row[q] = v;
            if (q == 1) {
              const double m = fabs(v.x) + fabs(v.y);
              if (m > best || (m == best && i < bi)) { best = m; bi = i; }
            }
          }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (j + 1 == c_sub && track) {
        // the next pivot column lies outside this sub-panel: it is finished by the rank-w update below, which tracks
        // the candidates itself
      } else if (track) {
        publish(best, bi);
      }
      if (j + 1 < c_sub) grid.sync();
    }
    grid.sync();
    if (c_sub < pe) {
      // ---- U block row of the remaining panel columns (unit-lower solve with the sub-panel's L11), rank-w update
      const int ncols = pe - c_sub;
      for (int e = tid; e < w * w; e += NT) s_L[e / w][e % w] = A[(int64_t)(s0 + e / w) * n + s0 + e % w];
      for (int e = tid; e < w * ncols; e += NT) s_U[e / ncols][e % ncols] = A[(int64_t)(s0 + e / ncols) * n + c_sub + e % ncols];
      __syncthreads();
      if (tid < ncols) {
        for (int r = 1; r < w; ++r) {
          cplx acc = s_U[r][tid];
          for (int k = 0; k < r; ++k) acc = csub(acc, cmul(s_L[r][k], s_U[k][tid]));
          s_U[r][tid] = acc;
        }
      }
      __syncthreads();
      grid.sync();   // every CTA holds the raw block row before CTA 0 overwrites it with the solved one
      if (blockIdx.x == 0)
        for (int e = tid; e < w * ncols; e += NT) A[(int64_t)(s0 + e / ncols) * n + c_sub + e % ncols] = s_U[e / ncols][e % ncols];
      double best = -1.0;
      int bi = 0x7fffffff;
      // one warp per row, two rows in flight per warp (their loads are independent)
      const int rstride = nctas * NW;
      for (int i0 = c_sub + blockIdx.x * NW + ty; i0 < n; i0 += 2 * rstride) {
        cplx l[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u * rstride;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            l[u][t] = (i < n && t < w) ? A[(int64_t)i * n + s0 + t] : make_double2(0.0, 0.0);
        }
        for (int c = tx; c < ncols; c += 32) {
          cplx acc[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * rstride;
            acc[u] = i < n ? A[(int64_t)i * n + c_sub + c] : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * rstride;
            if (i >= n) continue;
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (t < w) acc[u] = csub(acc[u], cmul(l[u][t], s_U[t][c]));
            A[(int64_t)i * n + c_sub + c] = acc[u];
            if (c == 0) {
              const double m = fabs(acc[u].x) + fabs(acc[u].y);
              if (m > best || (m == best && i < bi)) { best = m; bi = i; }
            }
          }
        }
      }
      publish(best, bi);
      grid.sync();
    }
  }
}


// b <- P b (apply the recorded row interchanges in order); single thread block, sequential by nature
__global__ void lu_permute_kernel(cplx* __restrict__ b, const int* __restrict__ piv, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    for (int j = 0; j < n; ++j) {
      const int p = piv[j];
      if (p != j) { const cplx t = b[j]; b[j] = b[p]; b[p] = t; }
    }
}

// diagonal block solve (in place on x[j0 .. j0+nb)): lower => unit lower forward substitution, else upper backward
__global__ void __launch_bounds__(64) tri_block_kernel(const cplx* __restrict__ A, int n, int j0, int nb, int lower,
                                                       cplx* __restrict__ x) {
  __shared__ cplx xs[64];
  const int t = threadIdx.x;
  if (t < nb) xs[t] = x[j0 + t];
  __syncthreads();
  if (lower) {
    for (int k = 0; k < nb; ++k) {
      const cplx xk = xs[k];
      if (t > k && t < nb) xs[t] = csub(xs[t], cmul(A[(int64_t)(j0 + t) * n + j0 + k], xk));
      __syncthreads();
    }
  } else {
    for (int k = nb - 1; k >= 0; --k) {
      if (t == k) xs[k] = cdiv(xs[k], A[(int64_t)(j0 + k) * n + j0 + k]);
      __syncthreads();
      const cplx xk = xs[k];
      if (t < k) xs[t] = csub(xs[t], cmul(A[(int64_t)(j0 + t) * n + j0 + k], xk));
      __syncthreads();
    }
  }
  if (t < nb) x[j0 + t] = xs[t];
}

// x[r] -= sum_{k in block} A[r][j0 + k] * x[j0 + k] for rows r in [r0, r1); one warp per row
__global__ void __launch_bounds__(256) tri_update_kernel(const cplx* __restrict__ A, int n, int j0, int nb, int r0, int r1,
                                                         cplx* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int r = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= r1) return;
  cplx acc = make_double2(0.0, 0.0);
  for (int k = lane; k < nb; k += 32) acc = cadd(acc, cmul(A[(int64_t)r * n + j0 + k], x[j0 + k]));
  acc.x = warp_sum(acc.x);
  acc.y = warp_sum(acc.y);
  if (lane == 0) x[r] = csub(x[r], acc);
}

// ---------------------------------------------------------------------------------------------------
// Fast triangular solves for many right-hand sides in sequence (every Arnoldi multiplication applies N^-1): the
// SB x SB diagonal blocks of L and U are inverted once after the factorisation, so that each block step of the
// substitution is one launch -- every CTA forms x_k = inv(T_kk) rhs_k redundantly (a 128 x 128 matvec out of L2) and
// then updates its share of the remaining right-hand side with one warp per row.
constexpr int SB = 64;

// inv[blk] = inverse of the diagonal block: blocks [0, nblk) unit-lower (L), [nblk, 2 nblk) upper (U); one thread per
// column of the inverse
__global__ void __launch_bounds__(SB) tri_invert_kernel(const cplx* __restrict__ A, int n, int nblk,
                                                        cplx* __restrict__ inv) {
  const int blk = blockIdx.x % nblk, upper = blockIdx.x / nblk;
  const int j0 = blk * SB;
  const int nb = n - j0 < SB ? n - j0 : SB;
  cplx* out = inv + (int64_t)blockIdx.x * SB * SB;
  const int j = threadIdx.x;
  for (int i = 0; i < SB; ++i) out[i * SB + j] = make_double2(0.0, 0.0);
  if (j >= nb) return;
  if (!upper) {
    for (int i = j; i < nb; ++i) {
      cplx acc = make_double2(i == j ? 1.0 : 0.0, 0.0);
      for (int k = j; k < i; ++k) acc = csub(acc, cmul(A[(int64_t)(j0 + i) * n + j0 + k], out[k * SB + j]));
      out[i * SB + j] = acc;
    }
  } else {
    for (int i = j; i >= 0; --i) {
      cplx acc = make_double2(i == j ? 1.0 : 0.0, 0.0);
      for (int k = i + 1; k <= j; ++k) acc = csub(acc, cmul(A[(int64_t)(j0 + i) * n + j0 + k], out[k * SB + j]));
      out[i * SB + j] = cdiv(acc, A[(int64_t)(j0 + i) * n + j0 + i]);
    }
  }
}

// one block step: solved[j0 ..] = inv_kk rhs[j0 ..]; rhs[r] -= sum_j A[r][j0 + j] solved[j0 + j] for r in [r0, r1)
__global__ void __launch_bounds__(256) tri_step_kernel(const cplx* __restrict__ A, int n, int j0, int nb,
                                                       const cplx* __restrict__ inv_kk, cplx* __restrict__ rhs,
                                                       cplx* __restrict__ solved, int r0, int r1) {
  __shared__ cplx bk[SB];
  __shared__ cplx xk[SB];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < SB) bk[tid] = tid < nb ? rhs[j0 + tid] : make_double2(0.0, 0.0);
  __syncthreads();
  for (int i = w; i < nb; i += 8) {
    cplx acc = make_double2(0.0, 0.0);
    for (int k = lane; k < nb; k += 32) acc = cadd(acc, cmul(inv_kk[i * SB + k], bk[k]));
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) xk[i] = acc;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < nb) solved[j0 + tid] = xk[tid];
  for (int r = r0 + blockIdx.x * 8 + w; r < r1; r += gridDim.x * 8) {
    const cplx* row = A + (int64_t)r * n + j0;
    cplx acc = make_double2(0.0, 0.0);
    for (int k = lane; k < nb; k += 32) acc = cadd(acc, cmul(row[k], xk[k]));
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) rhs[r] = csub(rhs[r], acc);
  }
}

// Wavefront substitution: ONE launch per triangular factor.  CTA c owns row block k (k = c for L, nblk-1-c for U),
// folds the solved blocks x_j into its right-hand side as soon as their flags appear (in dependency order), then
// publishes x_k.  The SB x SB matrix block of a step does not depend on x: thread (row r, quarter q) holds its 16
// entries in registers and loads those of the NEXT step before it waits for this step's flag (double-buffered register
// sets -- the first version staged each block through shared memory with synchronous loads, 8.6 us per step, and the
// last CTA's ~127 sequential steps were the whole 1.1 ms of a solve at n = 8192).  The inverse diagonal block sits in
// shared memory from the start, so what remains on the critical path of a step is a 64-element read, two small products
// and the flag hand-off.  All CTAs are co-resident (nblk <= 148 => n <= 9472); waits are bounded and raise a sticky error
// flag instead of hanging.  Out of place: the right-hand side is read as rhs[perm[i]] (perm: the row interchanges of the
// factorisation as one gather, see lu_perm_build_kernel; null = identity) and the solution is published in x, so the two
// substitutions of a solve ping-pong between the caller's vector and a scratch vector without a copy.
__global__ void __launch_bounds__(256) tri_wavefront_kernel(const cplx* __restrict__ A, int n, int nblk,
                                                            const cplx* __restrict__ inv, cplx* __restrict__ x,
                                                            unsigned long long* __restrict__ flags,
                                                            unsigned long long epoch, int upper,
                                                            const cplx* __restrict__ rhs, const int* __restrict__ perm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* invk = reinterpret_cast<cplx*>(smem_raw);   // [SB][SB + 1]
  __shared__ cplx acc[SB];
  __shared__ cplx xj[2][SB];
  __shared__ int failed;
  const int tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;              // four threads per row of the block
  const int k = upper ? nblk - 1 - (int)blockIdx.x : (int)blockIdx.x;
  const int j0 = k * SB;
  const int nb = n - j0 < SB ? n - j0 : SB;
  unsigned long long* fl = flags + (upper ? nblk : 0);
  const cplx* inv_kk = inv + (int64_t)((upper ? nblk : 0) + k) * SB * SB;
  const int nsteps = upper ? nblk - 1 - k : k;
  auto load_block = [&](int sidx, cplx (&dst)[SB / 4]) {
    const int j = upper ? nblk - 1 - sidx : sidx;
    const int c0 = j * SB;
    const int ncol = n - c0 < SB ? n - c0 : SB;
    const cplx* src = A + (int64_t)(j0 + r) * n + c0;
#pragma unroll
    for (int u = 0; u < SB / 4; ++u) {
      const int c = q + 4 * u;
      dst[u] = (r < nb && c < ncol) ? src[c] : make_double2(0.0, 0.0);
    }
  };
  cplx b0[SB / 4], b1[SB / 4];
  if (nsteps > 0) load_block(0, b0);
  for (int e = tid; e < SB * SB; e += 256) invk[(e / SB) * (SB + 1) + e % SB] = inv_kk[e];
  cplx mine = make_double2(0.0, 0.0);               // threads with q == 0: the running right-hand side of row r
  if (q == 0 && r < nb) mine = rhs[perm ? perm[j0 + r] : j0 + r];
  if (tid == 0) failed = 0;
  __syncthreads();
  // one step with the block in `cur`, prefetching the following block into `nxt`
  auto step = [&](int sidx, cplx (&cur)[SB / 4], cplx (&nxt)[SB / 4]) -> bool {
    if (sidx + 1 < nsteps) load_block(sidx + 1, nxt);
    const int j = upper ? nblk - 1 - sidx : sidx;
    const int c0 = j * SB;
    const int ncol = n - c0 < SB ? n - c0 : SB;
    cplx* xs = xj[sidx & 1];
    if (tid < 32) {
      // warp 0 waits for x_j and fetches it (two values per lane)
      if (tid == 0) {
        const long long t0 = clock64();
        while (*((volatile unsigned long long*)(fl + j)) != epoch) {
          if (clock64() - t0 > 4000000000ll) {
            failed = 1;
            break;
          }
        }
        __threadfence();
      }
      __syncwarp();
      xs[tid] = tid < ncol ? __ldcg(x + c0 + tid) : make_double2(0.0, 0.0);
      xs[tid + 32] = tid + 32 < ncol ? __ldcg(x + c0 + tid + 32) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (failed) return false;
    cplx s2 = make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < SB / 4; ++u) s2 = cadd(s2, cmul(cur[u], xs[q + 4 * u]));
    s2.x += __shfl_xor_sync(0xffffffffu, s2.x, 1);
    s2.y += __shfl_xor_sync(0xffffffffu, s2.y, 1);
    s2.x += __shfl_xor_sync(0xffffffffu, s2.x, 2);
    s2.y += __shfl_xor_sync(0xffffffffu, s2.y, 2);
    mine = csub(mine, s2);
    return true;     // (xj is double-buffered: the next step's writes go to the other half, no second barrier)
  };
  bool ok = true;
  for (int sidx = 0; sidx < nsteps && ok; sidx += 2) {
    ok = step(sidx, b0, b1);
    if (ok && sidx + 1 < nsteps) ok = step(sidx + 1, b1, b0);
  }
  if (!ok) {
    // sticky error word for the host (read by carc_relax / carc_lu_solve_blocks at their synchronisation points) and a
    // NaN block so that nothing downstream mistakes the unsolved right-hand side for a solution
    if (tid == 0) flags[2 * nblk] = 1ull;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (tid < nb) x[j0 + tid] = make_double2(nan, nan);
    return;
  }
  if (q == 0) acc[r] = mine;
  __syncthreads();
  {
    cplx s2 = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int c = q; c < SB; c += 4) s2 = cadd(s2, cmul(invk[r * (SB + 1) + c], acc[c]));
    s2.x += __shfl_xor_sync(0xffffffffu, s2.x, 1);
    s2.y += __shfl_xor_sync(0xffffffffu, s2.y, 1);
    s2.x += __shfl_xor_sync(0xffffffffu, s2.x, 2);
    s2.y += __shfl_xor_sync(0xffffffffu, s2.y, 2);
    if (q == 0 && r < nb) x[j0 + r] = s2;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *((volatile unsigned long long*)(fl + k)) = epoch;
}

// The row interchanges of a factorisation as ONE gather: perm[i] = the index of b that ends up in position i after
// "for j: swap(b[j], b[piv[j]])".  Built once per factorisation (valid[0] says so; lu_invert_diagonal_blocks clears it)
// by replaying the interchanges on an index array in shared memory; applying them to a vector one by one took 1.7 ms per
// solve at n = 8192 (a single thread's 8192 dependent global round trips), as long as both substitutions together.
__global__ void __launch_bounds__(1024) lu_perm_build_kernel(const int* __restrict__ piv, int n, int* __restrict__ valid,
                                                             int* __restrict__ perm) {
  extern __shared__ int perm_smem[];   // [n] indices, [n] pivots
  if (*((volatile int*)valid)) return;
  int* idx = perm_smem;
  int* pv = perm_smem + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    idx[i] = i;
    pv[i] = piv[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 0; j < n; ++j) {
      const int p = pv[j];
      if (p != j) {
        const int t = idx[j];
        idx[j] = idx[p];
        idx[p] = t;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = idx[i];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *valid = 1;
}

// ---------------------------------------------------------------------------------------------------
// small dense complex eigenproblem (k <= KMAX), thread 0 only.  Eigenvalues by Hessenberg reduction + shifted QR
// with deflation; the eigenvector of the selected eigenvalue by inverse iteration.
__device__ void small_eig_min_real(int k, const cplx* Min /* k x k row-major */, cplx* lambda_out, cplx* vec_out) {
  cplx H[KMAX][KMAX];
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) H[i][j] = Min[i * k + j];
  // Hessenberg form by Givens rotations (similarity)
  for (int col = 0; col + 2 < k; ++col)
    for (int row = k - 1; row > col + 1; --row) {
      const cplx a = H[row - 1][col], b = H[row][col];
      const double nb2 = cabs2(b);
      if (nb2 == 0.0) continue;
      const double r = sqrt(cabs2(a) + nb2);
      const cplx c = cscale(a, 1.0 / r), s = cscale(b, 1.0 / r);   // G = [[conj c, conj s], [-s, c]]
      for (int j = 0; j < k; ++j) {
        const cplx x = H[row - 1][j], y = H[row][j];
        H[row - 1][j] = cadd(cmulc(c, x), cmulc(s, y));
        H[row][j] = csub(cmul(c, y), cmul(s, x));
      }
      for (int i = 0; i < k; ++i) {   // right-multiply by G^H
        const cplx x = H[i][row - 1], y = H[i][row];
        H[i][row - 1] = cadd(cmul(x, c), cmul(y, s));
        H[i][row] = csub(cmulc(c, y), cmulc(s, x));   // y conj(c) - x conj(s)
      }
    }
  cplx ev[KMAX];
  int hi = k - 1;
  int iter = 0;
  double hnorm = 0.0;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) hnorm += cabs2(H[i][j]);
  hnorm = sqrt(hnorm);
  while (hi >= 0 && iter < 500) {
    if (hi == 0) { ev[0] = H[0][0]; break; }
    // deflation check
    int l = hi;
    while (l > 0) {
      const double sc = sqrt(cabs2(H[l - 1][l - 1])) + sqrt(cabs2(H[l][l]));
      if (sqrt(cabs2(H[l][l - 1])) <= 2.3e-16 * (sc > 0.0 ? sc : hnorm)) { H[l][l - 1] = make_double2(0.0, 0.0); break; }
      --l;
    }
    if (l == hi) { ev[hi] = H[hi][hi]; --hi; iter = 0; continue; }
    // Wilkinson shift from the trailing 2x2 of the active block
    const cplx a = H[hi - 1][hi - 1], b = H[hi - 1][hi], c = H[hi][hi - 1], d = H[hi][hi];
    const cplx tr = cadd(a, d), det = csub(cmul(a, d), cmul(b, c));
    cplx disc = csub(cmul(cscale(tr, 0.5), cscale(tr, 0.5)), det);
    // complex sqrt
    const double dm = sqrt(sqrt(cabs2(disc)));
    const double ang = 0.5 * atan2(disc.y, disc.x);
    const cplx sq = make_double2(dm * cos(ang), dm * sin(ang));
    const cplx e1 = cadd(cscale(tr, 0.5), sq), e2 = csub(cscale(tr, 0.5), sq);
    cplx mu = cabs2(csub(e1, d)) < cabs2(csub(e2, d)) ? e1 : e2;
    if (iter == 10 || iter == 20) mu = cadd(mu, make_double2(sqrt(cabs2(c)), 0.0));  // exceptional shift
    // QR step on the active block [l, hi] with Givens rotations
    cplx cs[KMAX], sn[KMAX];
    for (int i = l; i <= hi; ++i) H[i][i] = csub(H[i][i], mu);
    for (int i = l; i < hi; ++i) {
      const cplx x = H[i][i], y = H[i + 1][i];
      const double r = sqrt(cabs2(x) + cabs2(y));
      if (r == 0.0) { cs[i] = make_double2(1.0, 0.0); sn[i] = make_double2(0.0, 0.0); continue; }
      cs[i] = cscale(x, 1.0 / r);
      sn[i] = cscale(y, 1.0 / r);
      for (int j = 0; j < k; ++j) {
        const cplx u = H[i][j], v = H[i + 1][j];
        H[i][j] = cadd(cmulc(cs[i], u), cmulc(sn[i], v));
        H[i + 1][j] = csub(cmul(cs[i], v), cmul(sn[i], u));
      }
    }
    for (int i = l; i < hi; ++i) {
      for (int r2 = 0; r2 < k; ++r2) {
        const cplx u = H[r2][i], v = H[r2][i + 1];
        H[r2][i] = cadd(cmul(u, cs[i]), cmul(v, sn[i]));
        H[r2][i + 1] = csub(cmulc(cs[i], v), cmulc(sn[i], u));
      }
    }
    for (int i = l; i <= hi; ++i) H[i][i] = cadd(H[i][i], mu);
    ++iter;
  }
  while (hi > 0 && iter >= 500) { ev[hi] = H[hi][hi]; --hi; if (hi == 0) ev[0] = H[0][0]; }
  int best = 0;
  for (int i = 1; i < k; ++i)
    if (ev[i].x < ev[best].x) best = i;
  const cplx lam = ev[best];
  *lambda_out = lam;
  // inverse iteration on the original matrix with a slightly perturbed shift
  double mnorm = 0.0;
  for (int i = 0; i < k * k; ++i) mnorm += cabs2(Min[i]);
  mnorm = sqrt(mnorm);
  const cplx shift = cadd(lam, make_double2(1e-10 * (mnorm > 0.0 ? mnorm : 1.0), 0.0));
  cplx x[KMAX];
  for (int i = 0; i < k; ++i) x[i] = make_double2(1.0 / (1.0 + i), 0.3 / (2.0 + i));
  for (int it = 0; it < 3; ++it) {
    cplx M[KMAX][KMAX + 1];
    for (int i = 0; i < k; ++i) {
      for (int j = 0; j < k; ++j) M[i][j] = Min[i * k + j];
      M[i][i] = csub(M[i][i], shift);
      M[i][k] = x[i];
    }
    for (int col = 0; col < k; ++col) {
      int p = col;
      for (int r = col + 1; r < k; ++r)
        if (cabs2(M[r][col]) > cabs2(M[p][col])) p = r;
      if (p != col)
        for (int j = 0; j <= k; ++j) { const cplx t = M[col][j]; M[col][j] = M[p][j]; M[p][j] = t; }
      if (cabs2(M[col][col]) < 1e-300) M[col][col] = make_double2(1e-150, 0.0);
      for (int r = col + 1; r < k; ++r) {
        const cplx f = cdiv(M[r][col], M[col][col]);
        for (int j = col; j <= k; ++j) M[r][j] = csub(M[r][j], cmul(f, M[col][j]));
      }
    }
    for (int r = k - 1; r >= 0; --r) {
      cplx acc = M[r][k];
      for (int j = r + 1; j < k; ++j) acc = csub(acc, cmul(M[r][j], x[j]));
      x[r] = cdiv(acc, M[r][r]);
    }
    double nx = 0.0;
    for (int i = 0; i < k; ++i) nx += cabs2(x[i]);
    nx = 1.0 / sqrt(nx);
    for (int i = 0; i < k; ++i) x[i] = cscale(x[i], nx);
  }
  // LAPACK zgeev convention: unit Euclidean norm, component of largest modulus real (and positive).  With it a
  // converged restart returns the previous state unchanged instead of rotated by an arbitrary phase, which the
  // state-difference convergence policies rely on.
  int big = 0;
  for (int i = 1; i < k; ++i)
    if (cabs2(x[i]) > cabs2(x[big])) big = i;
  const double mod = sqrt(cabs2(x[big]));
  if (mod > 0.0) {
    const cplx ph = make_double2(x[big].x / mod, -x[big].y / mod);
    for (int i = 0; i < k; ++i) x[i] = cmul(x[i], ph);
    x[big].y = 0.0;
  }
  for (int i = 0; i < k; ++i) vec_out[i] = x[i];
}

// The same algorithm, same operations in the same order, for a compile-time dimension K <= 4 (the Krylov dimension of
// relaxOver is 3): every loop has static bounds and run-time ranges are predicates, so that H, the rotations and the
// elimination tableau live in registers instead of dynamically indexed local memory.  The generic routine above takes
// ~60 us at k = 3 (a few hundred dependent local-memory round trips), which made ritz_kernel the longest kernel of a
// restart at small bond dimensions.
template <int K>
__device__ void small_eig_min_real_fixed(const cplx* Min /* K x K row-major */, cplx* lambda_out, cplx* vec_out) {
  cplx H[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) H[i][j] = Min[i * K + j];
#pragma unroll
  for (int col = 0; col + 2 < K; ++col)
#pragma unroll
    for (int row = K - 1; row > col + 1; --row) {
      const cplx a = H[row - 1][col], b = H[row][col];
      const double nb2 = cabs2(b);
      if (nb2 != 0.0) {
        const double r = sqrt(cabs2(a) + nb2);
        const cplx c = cscale(a, 1.0 / r), s = cscale(b, 1.0 / r);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const cplx x = H[row - 1][j], y = H[row][j];
          H[row - 1][j] = cadd(cmulc(c, x), cmulc(s, y));
          H[row][j] = csub(cmul(c, y), cmul(s, x));
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const cplx x = H[i][row - 1], y = H[i][row];
          H[i][row - 1] = cadd(cmul(x, c), cmul(y, s));
          H[i][row] = csub(cmulc(c, y), cmulc(s, x));
        }
      }
    }
  cplx ev[K];
#pragma unroll
  for (int i = 0; i < K; ++i) ev[i] = make_double2(0.0, 0.0);
  int hi = K - 1;
  int iter = 0;
  double hnorm = 0.0;
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) hnorm += cabs2(H[i][j]);
  hnorm = sqrt(hnorm);
  while (hi >= 1 && iter < 500) {
    // deflation check: the largest l <= hi whose sub-diagonal entry is negligible (0 if none)
    int l = 0;
    bool found = false;
#pragma unroll
    for (int t = K - 1; t >= 1; --t) {
      if (t <= hi && !found) {
        const double sc = sqrt(cabs2(H[t - 1][t - 1])) + sqrt(cabs2(H[t][t]));
        if (sqrt(cabs2(H[t][t - 1])) <= 2.3e-16 * (sc > 0.0 ? sc : hnorm)) {
          H[t][t - 1] = make_double2(0.0, 0.0);
          l = t;
          found = true;
        }
      }
    }
    if (l == hi) {
#pragma unroll
      for (int t = 1; t < K; ++t)
        if (t == hi) ev[t] = H[t][t];
      --hi;
      iter = 0;
      continue;
    }
    cplx a = make_double2(0.0, 0.0), b = a, c = a, d = a;
#pragma unroll
    for (int t = 1; t < K; ++t)
      if (t == hi) {
        a = H[t - 1][t - 1];
        b = H[t - 1][t];
        c = H[t][t - 1];
        d = H[t][t];
      }
    const cplx tr = cadd(a, d), det = csub(cmul(a, d), cmul(b, c));
    const cplx disc = csub(cmul(cscale(tr, 0.5), cscale(tr, 0.5)), det);
    const double dm = sqrt(sqrt(cabs2(disc)));
    const double ang = 0.5 * atan2(disc.y, disc.x);
    const cplx sq = make_double2(dm * cos(ang), dm * sin(ang));
    const cplx e1 = cadd(cscale(tr, 0.5), sq), e2 = csub(cscale(tr, 0.5), sq);
    cplx mu = cabs2(csub(e1, d)) < cabs2(csub(e2, d)) ? e1 : e2;
    if (iter == 10 || iter == 20) mu = cadd(mu, make_double2(sqrt(cabs2(c)), 0.0));
    cplx cs[K], sn[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      cs[i] = make_double2(1.0, 0.0);
      sn[i] = make_double2(0.0, 0.0);
      if (i >= l && i <= hi) H[i][i] = csub(H[i][i], mu);
    }
#pragma unroll
    for (int i = 0; i + 1 < K; ++i) {
      if (i >= l && i < hi) {
        const cplx x = H[i][i], y = H[i + 1][i];
        const double r = sqrt(cabs2(x) + cabs2(y));
        if (r != 0.0) {
          cs[i] = cscale(x, 1.0 / r);
          sn[i] = cscale(y, 1.0 / r);
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const cplx u = H[i][j], v = H[i + 1][j];
            H[i][j] = cadd(cmulc(cs[i], u), cmulc(sn[i], v));
            H[i + 1][j] = csub(cmul(cs[i], v), cmul(sn[i], u));
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i + 1 < K; ++i) {
      if (i >= l && i < hi) {
#pragma unroll
        for (int r2 = 0; r2 < K; ++r2) {
          const cplx u = H[r2][i], v = H[r2][i + 1];
          H[r2][i] = cadd(cmul(u, cs[i]), cmul(v, sn[i]));
          H[r2][i + 1] = csub(cmulc(cs[i], v), cmulc(sn[i], u));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i >= l && i <= hi) H[i][i] = cadd(H[i][i], mu);
    ++iter;
  }
  if (iter >= 500) {
#pragma unroll
    for (int t = K - 1; t >= 1; --t)
      if (t <= hi) ev[t] = H[t][t];
  }
  ev[0] = H[0][0];
  int best = 0;
  cplx lam = ev[0];
#pragma unroll
  for (int i = 1; i < K; ++i)
    if (ev[i].x < lam.x) {
      best = i;
      lam = ev[i];
    }
  (void)best;
  *lambda_out = lam;
  double mnorm = 0.0;
#pragma unroll
  for (int i = 0; i < K * K; ++i) mnorm += cabs2(Min[i]);
  mnorm = sqrt(mnorm);
  const cplx shift = cadd(lam, make_double2(1e-10 * (mnorm > 0.0 ? mnorm : 1.0), 0.0));
  cplx x[K];
#pragma unroll
  for (int i = 0; i < K; ++i) x[i] = make_double2(1.0 / (1.0 + i), 0.3 / (2.0 + i));
  for (int it = 0; it < 3; ++it) {
    cplx M[K][K + 1];
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
      for (int j = 0; j < K; ++j) M[i][j] = Min[i * K + j];
      M[i][i] = csub(M[i][i], shift);
      M[i][K] = x[i];
    }
#pragma unroll
    for (int col = 0; col < K; ++col) {
      int pr = col;
      double pv = cabs2(M[col][col]);
#pragma unroll
      for (int r = col + 1; r < K; ++r)
        if (cabs2(M[r][col]) > pv) {
          pr = r;
          pv = cabs2(M[r][col]);
        }
#pragma unroll
      for (int r = col + 1; r < K; ++r)
        if (r == pr) {
#pragma unroll
          for (int j = 0; j <= K; ++j) {
            const cplx t = M[col][j];
            M[col][j] = M[r][j];
            M[r][j] = t;
          }
        }
      if (cabs2(M[col][col]) < 1e-300) M[col][col] = make_double2(1e-150, 0.0);
#pragma unroll
      for (int r = col + 1; r < K; ++r) {
        const cplx f = cdiv(M[r][col], M[col][col]);
#pragma unroll
        for (int j = col; j <= K; ++j) M[r][j] = csub(M[r][j], cmul(f, M[col][j]));
      }
    }
#pragma unroll
    for (int r = K - 1; r >= 0; --r) {
      cplx acc = M[r][K];
#pragma unroll
      for (int j = r + 1; j < K; ++j) acc = csub(acc, cmul(M[r][j], x[j]));
      x[r] = cdiv(acc, M[r][r]);
    }
    double nx = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) nx += cabs2(x[i]);
    nx = 1.0 / sqrt(nx);
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = cscale(x[i], nx);
  }
  int big = 0;
  double bigv = cabs2(x[0]);
#pragma unroll
  for (int i = 1; i < K; ++i)
    if (cabs2(x[i]) > bigv) {
      big = i;
      bigv = cabs2(x[i]);
    }
  const double mod = sqrt(bigv);
  if (mod > 0.0) {
    cplx xb = x[0];
#pragma unroll
    for (int i = 1; i < K; ++i)
      if (i == big) xb = x[i];
    const cplx ph = make_double2(xb.x / mod, -xb.y / mod);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      x[i] = cmul(x[i], ph);
      if (i == big) x[i].y = 0.0;
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i) vec_out[i] = x[i];
}

__device__ void small_eig_dispatch(int k, const cplx* Min, cplx* lambda_out, cplx* vec_out) {
  switch (k) {
    case 1: small_eig_min_real_fixed<1>(Min, lambda_out, vec_out); break;
    case 2: small_eig_min_real_fixed<2>(Min, lambda_out, vec_out); break;
    case 3: small_eig_min_real_fixed<3>(Min, lambda_out, vec_out); break;
    case 4: small_eig_min_real_fixed<4>(Min, lambda_out, vec_out); break;
    default: small_eig_min_real(k, Min, lambda_out, vec_out); break;
  }
}

// ---------------------------------------------------------------------------------------------------
// Arnoldi state shared by the kernels below (device memory)
struct RelaxState {
  cplx ritz;          // lowest Ritz value of the last restart
  cplx initial_value; // <v0, M v0>
  cplx final_value;   // <v, M v> at exit
  int kk;             // number of valid basis vectors in the current restart
  int complete;       // Krylov space exhausted (breakdown)
  int pad[2];
};

__global__ void restart_kernel(RelaxState* st) { st->kk = 1; }

// v <- v / ||v||
__global__ void __launch_bounds__(VT) normalize_kernel(cplx* __restrict__ v, int64_t n) {
  __shared__ cplx sh[32];
  cplx s[1] = {make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) s[0].x += cabs2(v[i]);
  block_csum<1>(s, sh);
  const double inv = 1.0 / sqrt(s[0].x);
  for (int64_t i = threadIdx.x; i < n; i += VT) v[i] = cscale(v[i], inv);
}

// out = <a, b> (conjugating a) written to *dst
__global__ void __launch_bounds__(VT) dot_kernel(const cplx* __restrict__ a, const cplx* __restrict__ b, int64_t n,
                                                 cplx* __restrict__ dst) {
  __shared__ cplx sh[32];
  cplx s[1] = {make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) s[0] = cadd(s[0], cmulc(a[i], b[i]));
  block_csum<1>(s, sh);
  if (threadIdx.x == 0) *dst = s[0];
}

// Classical Gram-Schmidt step i (utils.py:852-858): w = MV_i - sum_{j<=i} <V_j, MV_i> V_j ; V_{i+1} = w / ||w||, or
// flag breakdown when ||w|| <= 1e-14.  V, MV: [k, n] row-major.
__global__ void __launch_bounds__(VT) arnoldi_step_kernel(cplx* __restrict__ V, const cplx* __restrict__ MV, int64_t n,
                                                          int i, RelaxState* __restrict__ st) {
  __shared__ cplx sh[32 * KMAX];
  if (st->complete) return;
  const cplx* w = MV + (int64_t)i * n;
  cplx c[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) c[j] = make_double2(0.0, 0.0);
  for (int64_t e = threadIdx.x; e < n; e += VT) {
    const cplx we = w[e];
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j <= i) c[j] = cadd(c[j], cmulc(V[(int64_t)j * n + e], we));
  }
  block_csum<KMAX>(c, sh);
  cplx* out = V + (int64_t)(i + 1) * n;
  cplx nrm[1] = {make_double2(0.0, 0.0)};
  for (int64_t e = threadIdx.x; e < n; e += VT) {
    cplx we = w[e];
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j <= i) we = csub(we, cmul(c[j], V[(int64_t)j * n + e]));
    out[e] = we;
    nrm[0].x += cabs2(we);
  }
  block_csum<1>(nrm, sh);
  const double norm = sqrt(nrm[0].x);
  if (norm <= 1e-14) {
    if (threadIdx.x == 0) { st->complete = 1; st->kk = i + 1; }
    return;
  }
  const double inv = 1.0 / norm;
  for (int64_t e = threadIdx.x; e < n; e += VT) out[e] = cscale(out[e], inv);
  if (threadIdx.x == 0) st->kk = i + 2;
}

// projected matrix (utils.py:862), its lowest-real-part eigenpair (863-866), and the normalised Ritz vector
// written to `next` (869-870 / 875-876) and to row 0 of V, with the basis count reset: the start of the next restart.
// (256 threads: thread 0's register-resident eigen-solver needs more than the 64 registers a 1024-thread CTA allows)
constexpr int RITZ_THREADS = 256;
__global__ void __launch_bounds__(RITZ_THREADS) ritz_kernel(cplx* __restrict__ V, const cplx* __restrict__ MV, int64_t n,
                                                  int kmax, cplx* __restrict__ next, RelaxState* __restrict__ st) {
  __shared__ cplx sh[32 * KMAX];
  __shared__ cplx small[KMAX * KMAX];
  __shared__ cplx y[KMAX];
  const int kk = st->kk < kmax ? st->kk : kmax;
  for (int a = 0; a < kk; ++a) {
    cplx c[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) c[j] = make_double2(0.0, 0.0);
    for (int64_t e = threadIdx.x; e < n; e += RITZ_THREADS) {
      const cplx va = V[(int64_t)a * n + e];
#pragma unroll
      for (int j = 0; j < KMAX; ++j)
        if (j < kk) c[j] = cadd(c[j], cmulc(va, MV[(int64_t)j * n + e]));
    }
    block_csum<KMAX>(c, sh);
    if (threadIdx.x == 0)
      for (int j = 0; j < kk; ++j) small[a * kk + j] = c[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cplx lam;
    cplx vec[KMAX];
    small_eig_dispatch(kk, small, &lam, vec);
    st->ritz = lam;
    for (int j = 0; j < kk; ++j) y[j] = vec[j];
  }
  __syncthreads();
  cplx nrm[1] = {make_double2(0.0, 0.0)};
  for (int64_t e = threadIdx.x; e < n; e += RITZ_THREADS) {
    cplx acc = make_double2(0.0, 0.0);
    for (int j = 0; j < kk; ++j) acc = cadd(acc, cmul(y[j], V[(int64_t)j * n + e]));
    next[e] = acc;
    nrm[0].x += cabs2(acc);
  }
  block_csum<1>(nrm, sh);
  const double inv = 1.0 / sqrt(nrm[0].x);
  for (int64_t e = threadIdx.x; e < n; e += RITZ_THREADS) {
    const cplx val = cscale(next[e], inv);
    next[e] = val;
    V[e] = val;   // (this thread read V[j * n + e] for every j in the loop above)
  }
  if (threadIdx.x == 0) st->kk = 1;
}

// ---------------------------------------------------------------------------------------------------
// GMRES(m) pieces.  Hessenberg column, Givens rotations and the residual recurrence are kept on device.
struct GmresState {
  cplx h[(GM_MAX + 1) * GM_MAX];   // column-major: h[j * (m+1) + i]
  cplx cs[GM_MAX], sn[GM_MAX];
  cplx g[GM_MAX + 1];
  cplx y[GM_MAX];
  double resid;                    // |g[j+1]| after step j
  double bnorm;
  double rnorm0;
  int breakdown;
  int pad;
};

// r = b - (have_ax ? ax : 0); V0 = r / ||r||; g = (||r||, 0, ...)
__global__ void __launch_bounds__(VT) gmres_start_kernel(const cplx* __restrict__ b, const cplx* __restrict__ ax,
                                                         int64_t n, cplx* __restrict__ V0, GmresState* __restrict__ gs,
                                                         int first) {
  __shared__ cplx sh[32 * 2];
  cplx s[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) {
    cplx r = b[i];
    s[1].x += cabs2(r);
    if (ax) r = csub(r, ax[i]);
    V0[i] = r;
    s[0].x += cabs2(r);
  }
  block_csum<2>(s, sh);
  const double nr = sqrt(s[0].x);
  const double inv = nr > 0.0 ? 1.0 / nr : 0.0;
  for (int64_t i = threadIdx.x; i < n; i += VT) V0[i] = cscale(V0[i], inv);
  if (threadIdx.x == 0) {
    if (first) gs->bnorm = sqrt(s[1].x);
    gs->rnorm0 = nr;
    gs->resid = nr;
    gs->breakdown = 0;
    for (int i = 0; i <= GM_MAX; ++i) gs->g[i] = make_double2(0.0, 0.0);
    gs->g[0] = make_double2(nr, 0.0);
  }
}

// step j: w (= A V_j, already stored in V_{j+1}) is orthogonalised against V_0..V_j (modified Gram-Schmidt),
// normalised, the new Hessenberg column is rotated and the residual norm updated.
__global__ void __launch_bounds__(VT) gmres_step_kernel(cplx* __restrict__ V, int64_t n, int j, int m,
                                                        GmresState* __restrict__ gs) {
  __shared__ cplx sh[32];
  __shared__ cplx hcol[GM_MAX + 1];
  cplx* w = V + (int64_t)(j + 1) * n;
  for (int i = 0; i <= j; ++i) {
    const cplx* vi = V + (int64_t)i * n;
    cplx s[1] = {make_double2(0.0, 0.0)};
    for (int64_t e = threadIdx.x; e < n; e += VT) s[0] = cadd(s[0], cmulc(vi[e], w[e]));
    block_csum<1>(s, sh);
    for (int64_t e = threadIdx.x; e < n; e += VT) w[e] = csub(w[e], cmul(s[0], vi[e]));
    if (threadIdx.x == 0) hcol[i] = s[0];
    __syncthreads();
  }
  cplx s[1] = {make_double2(0.0, 0.0)};
  for (int64_t e = threadIdx.x; e < n; e += VT) s[0].x += cabs2(w[e]);
  block_csum<1>(s, sh);
  const double hn = sqrt(s[0].x);
  const double inv = hn > 0.0 ? 1.0 / hn : 0.0;
  for (int64_t e = threadIdx.x; e < n; e += VT) w[e] = cscale(w[e], inv);
  if (threadIdx.x == 0) {
    hcol[j + 1] = make_double2(hn, 0.0);
    for (int i = 0; i < j; ++i) {   // previous rotations
      const cplx a = hcol[i], b = hcol[i + 1];
      hcol[i] = cadd(cmulc(gs->cs[i], a), cmulc(gs->sn[i], b));
      hcol[i + 1] = csub(cmul(gs->cs[i], b), cmul(gs->sn[i], a));
    }
    const cplx a = hcol[j], b = hcol[j + 1];
    const double r = sqrt(cabs2(a) + cabs2(b));
    cplx c = make_double2(1.0, 0.0), sN = make_double2(0.0, 0.0);
    if (r > 0.0) { c = cscale(a, 1.0 / r); sN = cscale(b, 1.0 / r); }
    gs->cs[j] = c;
    gs->sn[j] = sN;
    hcol[j] = make_double2(r, 0.0);
    hcol[j + 1] = make_double2(0.0, 0.0);
    const cplx g0 = gs->g[j];
    gs->g[j] = cmulc(c, g0);
    gs->g[j + 1] = cscale(cmul(sN, g0), -1.0);
    gs->resid = sqrt(cabs2(gs->g[j + 1]));
    if (hn <= 1e-300) gs->breakdown = 1;
    for (int i = 0; i <= j + 1; ++i) gs->h[j * (m + 1) + i] = hcol[i];
  }
}

// x += V[:, 0..jj) y with R y = g (back substitution by thread 0)
__global__ void __launch_bounds__(VT) gmres_update_kernel(const cplx* __restrict__ V, int64_t n, int jj, int m,
                                                          cplx* __restrict__ x, GmresState* __restrict__ gs) {
  __shared__ cplx y[GM_MAX];
  if (threadIdx.x == 0) {
    for (int i = jj - 1; i >= 0; --i) {
      cplx acc = gs->g[i];
      for (int l = i + 1; l < jj; ++l) acc = csub(acc, cmul(gs->h[l * (m + 1) + i], y[l]));
      const cplx d = gs->h[i * (m + 1) + i];
      y[i] = (d.x != 0.0 || d.y != 0.0) ? cdiv(acc, d) : make_double2(0.0, 0.0);
    }
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < n; e += VT) {
    cplx acc = x[e];
    for (int l = 0; l < jj; ++l) acc = cadd(acc, cmul(y[l], V[(int64_t)l * n + e]));
    x[e] = acc;
  }
}


// ---------------------------------------------------------------------------------------------------
// Conjugate gradients for Hermitian positive semi-definite systems (the normal equations of compression.py).
struct CgState {
  double rs;      // <r, r>
  double bnorm;
  double resid;   // sqrt(<r, r>)
  double pad;
};

// x = 0, r = p = b
__global__ void __launch_bounds__(VT) cg_start_kernel(const cplx* __restrict__ b, int64_t n, cplx* __restrict__ x,
                                                      cplx* __restrict__ r, cplx* __restrict__ p,
                                                      CgState* __restrict__ cs) {
  __shared__ cplx sh[32];
  cplx s[1] = {make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) {
    const cplx v = b[i];
    x[i] = make_double2(0.0, 0.0);
    r[i] = v;
    p[i] = v;
    s[0].x += cabs2(v);
  }
  block_csum<1>(s, sh);
  if (threadIdx.x == 0) {
    cs->rs = s[0].x;
    cs->bnorm = sqrt(s[0].x);
    cs->resid = sqrt(s[0].x);
  }
}

// alpha = rs / <p, Ap>; x += alpha p; r -= alpha Ap; rs' = <r, r>; p = r + (rs'/rs) p
__global__ void __launch_bounds__(VT) cg_step_kernel(const cplx* __restrict__ Ap, int64_t n, cplx* __restrict__ x,
                                                     cplx* __restrict__ r, cplx* __restrict__ p,
                                                     CgState* __restrict__ cs) {
  __shared__ cplx sh[32];
  cplx s[1] = {make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) s[0] = cadd(s[0], cmulc(p[i], Ap[i]));
  block_csum<1>(s, sh);
  const double rs = cs->rs;
  const double pAp = s[0].x;
  if (!(pAp > 0.0) || rs == 0.0) {   // p in the null space / already converged: leave the state untouched
    return;
  }
  const double alpha = rs / pAp;
  cplx t[1] = {make_double2(0.0, 0.0)};
  for (int64_t i = threadIdx.x; i < n; i += VT) {
    x[i] = cadd(x[i], cscale(p[i], alpha));
    const cplx ri = csub(r[i], cscale(Ap[i], alpha));
    r[i] = ri;
    t[0].x += cabs2(ri);
  }
  block_csum<1>(t, sh);
  const double beta = t[0].x / rs;
  for (int64_t i = threadIdx.x; i < n; i += VT) p[i] = cadd(r[i], cscale(p[i], beta));
  __syncthreads();
  if (threadIdx.x == 0) {
    cs->rs = t[0].x;
    cs->resid = sqrt(t[0].x);
  }
}

}  // namespace

// ===================================================================================================
int dense_matvec(const cplx* M, int64_t rows, int64_t cols, int64_t ld, const cplx* x, cplx* y, cplx alpha, cplx beta,
                 cudaStream_t stream) {
  if (rows <= 0) return CARC_OK;
  zgemv_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(M, rows, cols, ld, x, y, alpha, beta);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

static int g_coop_ctas = 0;

size_t lu_scratch_bytes() { return sizeof(LuCand); }

// Two-level blocking: panels of 64 columns (the width the panel kernels factorise) inside outer blocks of 256 or 512.
// After a panel only the remaining columns of its outer block are updated (K = 64 products on few columns); the rest of
// the matrix sees ONE product per outer block with K = 256 / 512, where zgemm_kernel runs at its large-K rate -- with K = 64
// trailing updates it spent 146 of the 230 ms at n = 8192 re-reading and re-writing the trailing matrix (10 TFLOP/s).
// The U block rows are triangular solves with the unit-lower diagonal blocks (lu_trsm_unit_lower).
int lu_factor(cplx* A, int n, int* piv, int* singular_dev, cplx* scratch, cudaStream_t stream) {
  CARC_REQUIRE(n >= 1, CARC_ERR_VALUE, "lu_factor: n must be positive");
  const int NB = 64;
  static const int outer_env = getenv("CARC_LU_OUTER") ? atoi(getenv("CARC_LU_OUTER")) : 0;   // experiments
  const int OB = outer_env >= NB ? outer_env / NB * NB : n >= 6144 ? 512 : 256;   // n = 8192: 173 / 156 / 149 ms at 128 / 256 / 512
  const cplx minus_one = make_double2(-1.0, 0.0), one = make_double2(1.0, 0.0);
  CARC_CHECK_CUDA(cudaMemsetAsync(singular_dev, 0, sizeof(int), stream));
  // C[r0 : r0 + M, c0 : c0 + N] -= A[r0 : r0 + M, k0 : k0 + K] A[k0 : k0 + K, c0 : c0 + N], all blocks of the same matrix
  auto update = [&](int r0, int M, int c0, int N, int k0, int K) -> int {
    if (M <= 0 || N <= 0 || K <= 0) return CARC_OK;
    GemmOut o;
    o.m_div = n; o.m_s1 = 0; o.m_s0 = n;
    o.n_div = n; o.n_s1 = 0; o.n_s0 = 1;
    return zgemm(OP_N, OP_N, M, N, K, minus_one, A + (int64_t)r0 * n + k0, n, A + (int64_t)k0 * n + c0, n, one,
                 A + (int64_t)r0 * n + c0, &o, nullptr, 1, 0, 0, 0, stream);
  };
  for (int o0 = 0; o0 < n; o0 += OB) {
    const int o1 = n - o0 < OB ? n : o0 + OB;
    for (int j0 = o0; j0 < o1; j0 += NB) {
      const int nb = o1 - j0 < NB ? o1 - j0 : NB;
      const int pe = j0 + nb;
      // the panel: one thread-block cluster with the sub-panel in registers (lu_panel.cu); panels taller than a cluster
      // holds (n - j0 > 20480) take the cooperative kernel with grid-wide barriers
      const int prc = lu_panel_cluster(A, n, j0, nb, piv, singular_dev, stream);
      if (prc != CARC_OK && prc != CARC_ERR_UNSUPPORTED) return prc;
      if (prc == CARC_ERR_UNSUPPORTED) {
        if (g_coop_ctas == 0) {
          int dev = 0, sms = 0, per_sm = 0;
          CARC_CHECK_CUDA(cudaGetDevice(&dev));
          CARC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
          CARC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lu_panel_coop_kernel, 256, 0));
          g_coop_ctas = per_sm > 0 ? sms : 0;   // one CTA per SM (fewer, larger CTAs measured slower)
          CARC_REQUIRE(g_coop_ctas > 0, CARC_ERR_UNSUPPORTED, "lu_factor: cooperative launch unavailable");
        }
        LuCand* cand = reinterpret_cast<LuCand*>(scratch);
        int rows = n - j0;
        int ctas = (rows + 7) / 8;   // (the rank-w update uses one warp per row, 8 warps per CTA)
        if (ctas > g_coop_ctas) ctas = g_coop_ctas;
        if (ctas < 1) ctas = 1;
        int nn = n, jj = j0, nbb = nb;
        void* args[] = {(void*)&A, (void*)&nn, (void*)&jj, (void*)&nbb, (void*)&piv, (void*)&singular_dev, (void*)&cand};
        CARC_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)lu_panel_coop_kernel, dim3(ctas), dim3(256), args, 0, stream));
      }
      if (pe < o1) {
        // inside the outer block: U block row and rank-nb update of the block's remaining columns only
        int rc = lu_trsm_unit_lower(A, n, j0, nb, pe, o1 - pe, stream);
        if (rc) return rc;
        rc = update(pe, n - pe, pe, o1 - pe, j0, nb);
        if (rc) return rc;
      }
    }
    if (o1 < n) {
      // the outer block's U block row on the columns to the right (block forward substitution), then the trailing matrix
      for (int j0 = o0; j0 < o1; j0 += NB) {
        const int pe = o1 - j0 < NB ? o1 : j0 + NB;
        int rc = lu_trsm_unit_lower(A, n, j0, pe - j0, o1, n - o1, stream);
        if (rc) return rc;
        rc = update(pe, o1 - pe, o1, n - o1, j0, pe - j0);
        if (rc) return rc;
      }
      const int rc = update(o1, n - o1, o1, n - o1, o0, o1 - o0);
      if (rc) return rc;
    }
  }
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Cholesky factorisation of a Hermitian positive definite matrix, delivered in LU form.
//
// The normalization matrix N of the center-site problem is Hermitian positive definite whenever the environment is
// a proper (double-layer) one; the reference factorises it with the general lu_factor (utils.py:816-818).  For such N
// the blocked right-looking Cholesky needs no pivot search (so no grid-wide barriers), half the trailing-update work,
// and every step but a 64 x 64 diagonal block is a DMMA GEMM.  The result is rewritten as N = L' U' with L' = L D^-1
// unit lower, U' = D L^H, identity pivots -- the format carc_lu_solve / carc_lu_solve_blocks / carc_relax consume, so
// the solves are shared with the LU path.  A non-positive (or non-finite) pivot sets *status_dev = 1: the caller
// falls back to carc_lu_factor on a fresh copy.
constexpr int CB = 64;
constexpr int CLD = CB + 1;

__global__ void __launch_bounds__(256) chol_diag_kernel(cplx* __restrict__ A, int n, int j0, int nb,
                                                        cplx* __restrict__ winv, int* __restrict__ status_dev) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  cplx* Ls = reinterpret_cast<cplx*>(chol_smem);   // [CB][CLD] lower triangle of the block, then its Cholesky factor
  cplx* Wi = Ls + CB * CLD;                        // [CB][CLD] inverse of the factor
  const int tid = threadIdx.x;
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int i = idx / nb, j = idx % nb;
    Ls[i * CLD + j] = j <= i ? A[(int64_t)(j0 + i) * n + j0 + j] : make_double2(0.0, 0.0);
    Wi[i * CLD + j] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  bool bad = false;
  for (int k = 0; k < nb; ++k) {
    const double d2 = Ls[k * CLD + k].x;
    const bool ok = d2 > 0.0 && isfinite(d2);
    bad = bad || !ok;
    const double d = ok ? sqrt(d2) : 1.0;
    __syncthreads();   // every thread has read the pivot
    if (tid == 0) Ls[k * CLD + k] = make_double2(d, 0.0);
    const double inv = 1.0 / d;
    for (int i = k + 1 + tid; i < nb; i += blockDim.x) Ls[i * CLD + k] = cscale(Ls[i * CLD + k], inv);
    __syncthreads();
    const int m = nb - k - 1;
    for (int idx = tid; idx < m * m; idx += blockDim.x) {
      const int ii = idx / m, jj = idx % m;
      if (jj <= ii) {
        const int i = k + 1 + ii, j = k + 1 + jj;
        Ls[i * CLD + j] = csub(Ls[i * CLD + j], cmulc(Ls[j * CLD + k], Ls[i * CLD + k]));   // conj(L[j][k]) L[i][k]
      }
    }
    __syncthreads();
  }
  if (bad && tid == 0) *status_dev = 1;
  // Wi = L^-1: column j by four adjacent lanes (each sums a quarter of the dot product)
  {
    const int j = tid >> 2, part = tid & 3;
    for (int i = 0; i < nb; ++i) {
      cplx acc = make_double2(0.0, 0.0);
      if (j < nb && i > j)
        for (int k = j + part; k < i; k += 4) acc = cadd(acc, cmul(Ls[i * CLD + k], Wi[k * CLD + j]));
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
      if (j < nb && part == 0 && i >= j) {
        const double inv = 1.0 / Ls[i * CLD + i].x;
        Wi[i * CLD + j] = i == j ? make_double2(inv, 0.0) : make_double2(-acc.x * inv, -acc.y * inv);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int i = idx / nb, j = idx % nb;
    if (j <= i) A[(int64_t)(j0 + i) * n + j0 + j] = Ls[i * CLD + j];
    winv[i * nb + j] = Wi[i * CLD + j];
  }
}

// Rewrite the Cholesky factor (lower triangle of A) as LU factors: for i < j  U[i][j] = d_i conj(L[j][i]) and
// L'[j][i] = L[j][i] / d_i; U[i][i] = d_i^2.  32 x 32 tiles through shared memory so both accesses are coalesced.
__global__ void __launch_bounds__(256) chol_to_lu_kernel(cplx* __restrict__ A, int n) {
  __shared__ cplx tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;   // tile (bi, bj) of the LOWER triangle: bi >= bj
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // load the lower tile rows bi*32.., cols bj*32..
  for (int rr = ty; rr < 32; rr += 8) {
    const int i = bi * 32 + rr, j = bj * 32 + tx;
    tile[rr][tx] = (i < n && j < n && j <= i) ? A[(int64_t)i * n + j] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  // upper tile (bj, bi): U[j][i] = d_j conj(L[i][j]) with d_j = L[j][j]
  for (int rr = ty; rr < 32; rr += 8) {
    const int j = bj * 32 + rr, i = bi * 32 + tx;   // writing element (row j, col i), j < i
    if (i < n && j < n && j < i) {
      const double dj = A[(int64_t)j * n + j].x;
      const cplx l = tile[tx][rr];
      A[(int64_t)j * n + i] = make_double2(dj * l.x, -dj * l.y);
    }
  }
}
// second pass: scale the strictly lower part by 1 / d_column and square the diagonal (reads the diagonal of pass 1)
__global__ void __launch_bounds__(256) chol_scale_lower_kernel(cplx* __restrict__ A, int n, const double* __restrict__ diag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  if (j < i) {
    A[idx] = cscale(A[idx], 1.0 / diag[j]);
  } else if (j == i) {
    A[idx] = make_double2(diag[i] * diag[i], 0.0);
  }
}
__global__ void __launch_bounds__(256) chol_diag_extract_kernel(const cplx* __restrict__ A, int n, double* __restrict__ diag,
                                                                int* __restrict__ piv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    diag[i] = A[(int64_t)i * n + i].x;
    piv[i] = i;
  }
}

// sums[0] += sum |A[i][j] - conj(A[j][i])|^2 over i > j, sums[1] += sum |A[i][j]|^2 over all (i, j): tile pairs through
// shared memory so both reads are coalesced
__global__ void __launch_bounds__(256) hermitian_defect_kernel(const cplx* __restrict__ A, int n, double* __restrict__ sums) {
  __shared__ cplx lo[32][33], up[32][33];
  __shared__ double red[2][8];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int i = bi * 32 + rr, j = bj * 32 + tx;
    lo[rr][tx] = (i < n && j < n) ? A[(int64_t)i * n + j] : make_double2(0.0, 0.0);
    const int i2 = bj * 32 + rr, j2 = bi * 32 + tx;
    up[rr][tx] = (i2 < n && j2 < n) ? A[(int64_t)i2 * n + j2] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  double defect = 0.0, total = 0.0;
  for (int rr = ty; rr < 32; rr += 8) {
    const cplx a = lo[rr][tx], b = up[tx][rr];   // A[i][j] and A[j][i]
    const int i = bi * 32 + rr, j = bj * 32 + tx;
    if (bi != bj) {
      defect += (a.x - b.x) * (a.x - b.x) + (a.y + b.y) * (a.y + b.y);
      total += cabs2(a) + cabs2(b);
    } else if (j <= i) {
      if (j < i) {
        defect += (a.x - b.x) * (a.x - b.x) + (a.y + b.y) * (a.y + b.y);
        total += cabs2(a) + cabs2(b);
      } else {
        defect += a.y * a.y;
        total += cabs2(a);
      }
    }
  }
  defect = warp_sum(defect);
  total = warp_sum(total);
  if (tx == 0) { red[0][ty] = defect; red[1][ty] = total; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0, t = 0.0;
    for (int w = 0; w < 8; ++w) { d += red[0][w]; t += red[1][w]; }
    atomicAdd(&sums[0], d);
    atomicAdd(&sums[1], t);
  }
}

int hermitian_defect(const cplx* A, int n, double* sums_dev, cudaStream_t stream) {
  CARC_CHECK_CUDA(cudaMemsetAsync(sums_dev, 0, 2 * sizeof(double), stream));
  const int nt = (n + 31) / 32;
  hermitian_defect_kernel<<<dim3(nt, nt), 256, 0, stream>>>(A, n, sums_dev);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int cholesky_factor_as_lu(cplx* A, int n, int* piv, int* status_dev, cudaStream_t stream) {
  CARC_REQUIRE(n >= 1, CARC_ERR_VALUE, "cholesky: n must be positive");
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  const size_t smem = sizeof(cplx) * 2 * CB * CLD;
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  const cplx one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), minus_one = make_double2(-1.0, 0.0);
  CARC_CHECK_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), stream));
  cplx* winv = nullptr;
  double* diag = nullptr;
  CARC_CHECK_CUDA(cudaMallocAsync((void**)&winv, sizeof(cplx) * CB * CB, stream));
  CARC_CHECK_CUDA(cudaMallocAsync((void**)&diag, sizeof(double) * n, stream));
  GemmOut o;
  o.m_div = n; o.m_s1 = 0; o.m_s0 = n;
  o.n_div = n; o.n_s1 = 0; o.n_s0 = 1;
  for (int j0 = 0; j0 < n; j0 += CB) {
    const int nb = n - j0 < CB ? n - j0 : CB;
    const int pe = j0 + nb;
    chol_diag_kernel<<<1, 256, smem, stream>>>(A, n, j0, nb, winv, status_dev);
    if (pe < n) {
      const int rows = n - pe;
      cplx* A21 = A + (int64_t)pe * n + j0;
      // L21 = A21 L11^-H  (in place: a CTA's 128 x 64 tile reads exactly the rows it overwrites, all K = nb columns)
      int rc = zgemm(OP_N, OP_C, rows, nb, nb, one, A21, n, winv, nb, zero, A21, &o, nullptr, 1, 0, 0, 0, stream);
      if (rc) return rc;
      // A22 -= L21 L21^H on the lower triangle's tiles
      rc = zgemm_lower(OP_N, OP_C, rows, nb, minus_one, A21, n, A21, n, one, A + (int64_t)pe * n + pe, &o, stream);
      if (rc) return rc;
    }
  }
  chol_diag_extract_kernel<<<(n + 255) / 256, 256, 0, stream>>>(A, n, diag, piv);
  const int nt = (n + 31) / 32;
  chol_to_lu_kernel<<<dim3(nt, nt), 256, 0, stream>>>(A, n);
  chol_scale_lower_kernel<<<(unsigned)(((int64_t)n * n + 255) / 256), 256, 0, stream>>>(A, n, diag);
  CARC_CHECK_CUDA(cudaGetLastError());
  CARC_CHECK_CUDA(cudaFreeAsync(winv, stream));
  CARC_CHECK_CUDA(cudaFreeAsync(diag, stream));
  return CARC_OK;
}

int lu_solve(const cplx* LU, int n, const int* piv, cplx* x, cudaStream_t stream) {
  const int NB = 64;
  lu_permute_kernel<<<1, 32, 0, stream>>>(x, piv, n);
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = n - j0 < NB ? n - j0 : NB;
    tri_block_kernel<<<1, 64, 0, stream>>>(LU, n, j0, nb, 1, x);
    const int r0 = j0 + nb;
    if (r0 < n) tri_update_kernel<<<(n - r0 + 7) / 8, 256, 0, stream>>>(LU, n, j0, nb, r0, n, x);
  }
  for (int j0 = ((n - 1) / NB) * NB; j0 >= 0; j0 -= NB) {
    const int nb = n - j0 < NB ? n - j0 : NB;
    tri_block_kernel<<<1, 64, 0, stream>>>(LU, n, j0, nb, 0, x);
    if (j0 > 0) tri_update_kernel<<<(j0 + 7) / 8, 256, 0, stream>>>(LU, n, j0, nb, 0, j0, x);
  }
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}


int lu_invert_diagonal_blocks(const cplx* LU, int n, cplx* inv, cudaStream_t stream) {
  const int nblk = (n + SB - 1) / SB;
  tri_invert_kernel<<<2 * nblk, SB, 0, stream>>>(LU, n, nblk, inv);
  // flag words and the "permutation built" word behind them
  CARC_CHECK_CUDA(cudaMemsetAsync(inv + 2ll * nblk * SB * SB, 0, sizeof(cplx) * (2 * nblk + 3), stream));
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// inverse blocks followed by 2 nblk + 1 flag words (as complex slots: 16 bytes each, more than enough), one slot for the
// "permutation built" word and the gather permutation of the row interchanges (n ints)
int64_t lu_inverse_blocks_elems(int n) {
  const int64_t nblk = (n + SB - 1) / SB;
  return 2 * nblk * SB * SB + 2 * nblk + 3 + (n + 3) / 4;
}

// x <- A^-1 x with the pre-inverted diagonal blocks; tmp: n complex
int lu_solve_fast(const cplx* LU, int n, const int* piv, const cplx* inv, cplx* x, cplx* tmp, cudaStream_t stream) {
  const int nblk = (n + SB - 1) / SB;
  if (nblk <= sm_count()) {
    static std::atomic<unsigned long long> epoch_counter{0};
    cplx* tail = const_cast<cplx*>(inv) + 2ll * nblk * SB * SB;
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(tail);
    int* perm_valid = reinterpret_cast<int*>(tail + 2 * nblk + 2);
    int* perm = reinterpret_cast<int*>(tail + 2 * nblk + 3);
    const unsigned long long epoch = ++epoch_counter;
    const size_t smem = sizeof(cplx) * SB * (SB + 1);
    const size_t perm_smem = sizeof(int) * 2 * (size_t)n;
    static bool configured[16] = {false};
    int dev = 0;
    CARC_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev]) {
      CARC_CHECK_CUDA(cudaFuncSetAttribute(tri_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CARC_CHECK_CUDA(cudaFuncSetAttribute(lu_perm_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           2 * sm_count() * SB * (int)sizeof(int)));   // n <= sm_count() * SB on this path
      configured[dev] = true;
    }
    // P b as a gather: the permutation is built on the first solve after a factorisation (a no-op launch afterwards)
    lu_perm_build_kernel<<<1, 1024, perm_smem, stream>>>(piv, n, perm_valid, perm);
    // L y = P b: b = x (gathered), y -> tmp;  U x = y: y = tmp, x -> x
    tri_wavefront_kernel<<<nblk, 256, smem, stream>>>(LU, n, nblk, inv, tmp, flags, epoch, 0, x, perm);
    tri_wavefront_kernel<<<nblk, 256, smem, stream>>>(LU, n, nblk, inv, x, flags, epoch, 1, tmp, nullptr);
    CARC_CHECK_CUDA(cudaGetLastError());
    return CARC_OK;
  }
  lu_permute_kernel<<<1, 32, 0, stream>>>(x, piv, n);
  for (int k = 0; k < nblk; ++k) {   // L y = P b : running rhs in x, y into tmp
    const int j0 = k * SB, nb = n - j0 < SB ? n - j0 : SB, r0 = j0 + nb;
    int blocks = (n - r0 + 7) / 8;
    blocks = blocks < 1 ? 1 : (blocks > 74 ? 74 : blocks);   // every CTA re-reads the 256 KB inverse block from L2
    tri_step_kernel<<<blocks, 256, 0, stream>>>(LU, n, j0, nb, inv + (int64_t)k * SB * SB, x, tmp, r0, n);
  }
  for (int k = nblk - 1; k >= 0; --k) {   // U x = y : running rhs in tmp, x into x
    const int j0 = k * SB, nb = n - j0 < SB ? n - j0 : SB;
    int blocks = (j0 + 7) / 8;
    blocks = blocks < 1 ? 1 : (blocks > 74 ? 74 : blocks);
    tri_step_kernel<<<blocks, 256, 0, stream>>>(LU, n, j0, nb, inv + (int64_t)(nblk + k) * SB * SB, tmp, x, 0, j0);
  }
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// ---------------------------------------------------------------------------------------------------
int gmres(const LinOp& A, const cplx* b, cplx* x, int64_t n, double rtol, int restart, int maxiter, cplx* work,
          void* state_dev, int* iters_out, double* resid_out, cudaStream_t stream) {
  // work: (restart + 2) * n complex; state_dev: sizeof(GmresState)
  CARC_REQUIRE(restart >= 1 && restart <= GM_MAX, CARC_ERR_VALUE, "gmres: restart %d outside 1..%d", restart, GM_MAX);
  GmresState* gs = reinterpret_cast<GmresState*>(state_dev);
  cplx* V = work;
  cplx* ax = work + (int64_t)(restart + 1) * n;
  CARC_CHECK_CUDA(cudaMemsetAsync(x, 0, sizeof(cplx) * n, stream));
  int total = 0;
  double host[3];
  bool first = true;
  for (int cycle = 0; cycle < maxiter; ++cycle) {
    if (!first) {
      int rc = A.apply(x, ax, stream);
      if (rc) return rc;
    }
    gmres_start_kernel<<<1, VT, 0, stream>>>(b, first ? nullptr : ax, n, V, gs, first ? 1 : 0);
    CARC_CHECK_CUDA(cudaMemcpyAsync(host, &gs->resid, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream));
    CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
    first = false;
    const double target = rtol * host[1];
    if (host[0] <= target || host[1] == 0.0) {
      *iters_out = total;
      *resid_out = host[0];
      return CARC_OK;
    }
    int jj = 0;
    bool converged = false;
    for (int j = 0; j < restart; ++j) {
      int rc = A.apply(V + (int64_t)j * n, V + (int64_t)(j + 1) * n, stream);
      if (rc) return rc;
      gmres_step_kernel<<<1, VT, 0, stream>>>(V, n, j, restart, gs);
      ++total;
      jj = j + 1;
      CARC_CHECK_CUDA(cudaMemcpyAsync(host, &gs->resid, sizeof(double), cudaMemcpyDeviceToHost, stream));
      int bd = 0;
      CARC_CHECK_CUDA(cudaMemcpyAsync(&bd, &gs->breakdown, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
      if (host[0] <= target || bd) { converged = true; break; }
    }
    gmres_update_kernel<<<1, VT, 0, stream>>>(V, n, jj, restart, x, gs);
    CARC_CHECK_CUDA(cudaGetLastError());
    if (converged) {
      *iters_out = total;
      *resid_out = host[0];
      return CARC_OK;
    }
  }
  *iters_out = total;
  *resid_out = host[0];
  set_error("gmres: no convergence after %d iterations (residual %.3e)", total, host[0]);
  return CARC_ERR_NO_CONVERGENCE;
}

size_t gmres_state_bytes() { return sizeof(GmresState); }

// ---------------------------------------------------------------------------------------------------
int cg(const LinOp& A, const cplx* b, cplx* x, int64_t n, double rtol, int maxiter, cplx* work, void* state_dev,
       int* iters_out, double* resid_out, cudaStream_t stream) {
  // work: 3 * n complex (r, p, Ap)
  CgState* cs = reinterpret_cast<CgState*>(state_dev);
  cplx* r = work;
  cplx* p = work + n;
  cplx* Ap = work + 2 * n;
  cg_start_kernel<<<1, VT, 0, stream>>>(b, n, x, r, p, cs);
  CgState host;
  CARC_CHECK_CUDA(cudaMemcpyAsync(&host, cs, sizeof(CgState), cudaMemcpyDeviceToHost, stream));
  CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
  const double target = rtol * host.bnorm;
  int it = 0;
  double best = host.resid;
  int since_best = 0;
  const int check_every = 8;
  while (host.resid > target && it < maxiter) {
    for (int q = 0; q < check_every && it < maxiter; ++q, ++it) {
      int rc = A.apply(p, Ap, stream);
      if (rc) return rc;
      cg_step_kernel<<<1, VT, 0, stream>>>(Ap, n, x, r, p, cs);
    }
    CARC_CHECK_CUDA(cudaMemcpyAsync(&host, cs, sizeof(CgState), cudaMemcpyDeviceToHost, stream));
    CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (host.resid < 0.999 * best) {
      best = host.resid;
      since_best = 0;
    } else if (++since_best >= 64) {
      break;   // stagnated at the rounding floor of an ill-conditioned system
    }
  }
  *iters_out = it;
  *resid_out = host.resid;
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int lu_solve_status(const cplx* inv, int n, int* timed_out) {
  const int nblk = (n + SB - 1) / SB;
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(inv + 2ll * nblk * SB * SB);
  unsigned long long v = 0;
  CARC_CHECK_CUDA(cudaMemcpy(&v, flags + 2 * nblk, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = v != 0;
  return CARC_OK;
}

size_t cg_state_bytes() { return sizeof(CgState); }


// ---------------------------------------------------------------------------------------------------
int relax(const LinOp& M, cplx* v, int64_t n, int max_mults, double tol, int k, cplx* work, void* state_dev,
          RelaxInfo* info, cudaStream_t stream) {
  // work: (2k + 2) * n complex.  M applies N^-1 H.
  CARC_REQUIRE(k >= 1 && k <= KMAX, CARC_ERR_UNSUPPORTED, "relax: Krylov dimension %d outside 1..%d", k, KMAX);
  RelaxState* st = reinterpret_cast<RelaxState*>(state_dev);
  cplx* V = work;                       // [k+1, n] (one spare row for the last CGS output)
  cplx* MV = work + (int64_t)(k + 1) * n;  // [k, n]
  cplx* tmp = MV + (int64_t)k * n;      // [n]
  CARC_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(RelaxState), stream));
  normalize_kernel<<<1, VT, 0, stream>>>(v, n);
  int rc = M.apply(v, tmp, stream);
  if (rc) return rc;
  dot_kernel<<<1, VT, 0, stream>>>(v, tmp, n, &st->initial_value);
  int mults = 0, total_applies = 1;
  bool have_last = false;
  double last_re = 0.0, last_im = 0.0;
  bool complete = (int64_t)k == n;
  RelaxState host;
  CARC_CHECK_CUDA(cudaMemcpyAsync(V, v, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, stream));
  restart_kernel<<<1, 1, 0, stream>>>(st);
  for (;;) {
    // (later restarts: ritz_kernel has left the Ritz vector in row 0 of V and reset the basis count)
    for (int i = 0; i < k; ++i) {
      // after a breakdown the remaining multiplications act on stale rows and are ignored by the Ritz step; the
      // reference stops multiplying at that point (utils.py:854-858) -- so do we, at the next host check below
      rc = M.apply(V + (int64_t)i * n, MV + (int64_t)i * n, stream);
      if (rc) return rc;
      ++total_applies;
      if (i < k - 1) {
        arnoldi_step_kernel<<<1, VT, 0, stream>>>(V, MV, n, i, st);
        if ((int64_t)(i + 2) > n) {   // tiny spaces: check for breakdown before touching row i+1 again
          CARC_CHECK_CUDA(cudaMemcpyAsync(&host, st, sizeof(RelaxState), cudaMemcpyDeviceToHost, stream));
          CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
          if (host.complete) break;
        }
      }
    }
    mults += k;
    ritz_kernel<<<1, RITZ_THREADS, 0, stream>>>(V, MV, n, k, v, st);
    CARC_CHECK_CUDA(cudaMemcpyAsync(&host, st, sizeof(RelaxState), cudaMemcpyDeviceToHost, stream));
    CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
    // a timed-out peer exchange or wavefront solve hands back NaNs (comm.cu, tri_wavefront_kernel): stop here instead of
    // iterating on garbage up to the multiplication cap
    CARC_REQUIRE(isfinite(host.ritz.x) && isfinite(host.ritz.y), CARC_ERR_EXCHANGE,
                 "relax: non-finite Ritz value after %d multiplications (NaN input, or a bounded device-side wait timed out)",
                 mults);
    if (host.complete) complete = true;
    const double dre = host.ritz.x - last_re, dim = host.ritz.y - last_im;
    const bool small_change = have_last && sqrt(dre * dre + dim * dim) <= tol;
    const bool out_of_budget = max_mults > 0 && mults >= max_mults;
    if (complete || small_change || out_of_budget) break;
    last_re = host.ritz.x;
    last_im = host.ritz.y;
    have_last = true;
  }
  rc = M.apply(v, tmp, stream);
  if (rc) return rc;
  ++total_applies;
  dot_kernel<<<1, VT, 0, stream>>>(v, tmp, n, &st->final_value);
  CARC_CHECK_CUDA(cudaMemcpyAsync(&host, st, sizeof(RelaxState), cudaMemcpyDeviceToHost, stream));
  CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
  info->initial_value[0] = host.initial_value.x; info->initial_value[1] = host.initial_value.y;
  info->final_value[0] = host.final_value.x; info->final_value[1] = host.final_value.y;
  info->ritz_value[0] = host.ritz.x; info->ritz_value[1] = host.ritz.y;
  info->multiplications = mults;
  info->applications = total_applies;
  // utils.py:871: (final - initial) / (|final| + |initial|) > 1 + 1e-7 (compared on the real part)
  const double fi = host.final_value.x - host.initial_value.x;
  const double den = hypot(host.final_value.x, host.final_value.y) + hypot(host.initial_value.x, host.initial_value.y);
  if (den > 0.0 && fi / den > 1.0 + 1e-7) {
    set_error("relax: expectation rose from %.17g to %.17g", host.initial_value.x, host.final_value.x);
    return CARC_ERR_RELAX_FAILED;
  }
  return CARC_OK;
}

size_t relax_state_bytes() { return sizeof(RelaxState); }

}  // namespace carc
