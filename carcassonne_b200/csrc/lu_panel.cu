// Panel factorisation of the blocked LU (scipy.linalg.lu_factor, reference utils.py:816-818) by ONE thread-block cluster.
//
// A 64-column panel needs 64 dependent column steps (pivot search -> row interchange -> scale -> rank-1 update).  The
// cooperative kernel in solver.cu runs them on all SMs with two grid-wide barriers per column and the panel in global
// memory: 19 us per column at n = 8192, i.e. 156 of the 212 ms of the whole factorisation were barrier latency.  Here
// the panel is factorised by one cluster of 8 (or 16) CTAs:
//   * every thread keeps the 8 sub-panel entries of its row(s) in REGISTERS for the 8 column steps of a sub-panel
//     (the column loop is unrolled, so all indexing is static) -- the column steps touch no memory but the exchange
//     slots;
//   * per column ONE cluster barrier (~0.2 us): before it every CTA pushes its pivot candidate -- value, row index AND
//     that row's sub-panel entries -- into the exchange slots of all CTAs through distributed shared memory, and the
//     owner of row j pushes row j; after it every CTA knows the pivot row's contents without reading it from its owner,
//     the owner of row p overwrites its registers with the old row j, and all rows update.  The candidates of column
//     j + 1 are tracked during the update of column j, as in LAPACK-style fused panels;
//   * after the 8 columns the sub-panel is written back, CTA 0 applies the 8 interchanges to the panel columns on the
//     right (net effect computed in shared memory: one read and one write of the <= 16 rows involved), solves the 8 x 56
//     U block row, and after one more cluster barrier every thread applies the rank-8 update to its rows with the L
//     entries it still holds in registers;
//   * interchanges are NOT applied outside the panel (nor to the panel columns on the left) by this kernel:
//     lu_laswp_kernel does that afterwards for all columns at once (LAPACK's zlaswp), again as a net permutation
//     through shared memory instead of 64 dependent global round trips.
// Pivot rule and arithmetic are those of the cooperative kernel (|re| + |im| maximal, ties to the smallest row; l = a / d
// as a multiplication with the reciprocal), so both produce the same pivots -- the tests compare them with SciPy's.
#include <cooperative_groups.h>

#include <cstdlib>

#include "carc_internal.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace carc {

namespace {

constexpr int LPC_MAXC = 16;   // largest cluster
constexpr int SUBW = 8;        // columns per sub-panel (register-resident)
constexpr int PANEL = 64;      // columns per panel (lu_factor's NB)
constexpr int NOROW = 0x7fffffff;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ cplx ldcg(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return v;
}
__device__ __forceinline__ void stcg(cplx* p, cplx v) { __stcg(reinterpret_cast<double2*>(p), v); }

struct LpcSlot {   // one CTA's pivot candidate of a column: value, row index, that row's sub-panel entries
  double val;
  int idx;
  int pad;
  cplx row[SUBW];
};

struct LpcShared {
  LpcSlot slot[2][LPC_MAXC];   // [column parity][source CTA]
  cplx rowj[2][SUBW];          // row j of the column step, pushed by its owner (CTA 0)
  double w_val[32];
  int w_idx[32];
  int piv[SUBW];
  int gslot[SUBW];
  cplx T[SUBW][PANEL];         // CTA 0: top block of the right part; all: the solved U block row
  cplx G[SUBW][PANEL];         // CTA 0: rows interchanged in from below the sub-panel
  cplx L[SUBW][SUBW];
};

template <int THREADS, int RPT>
__global__ void __launch_bounds__(THREADS, 1) lu_panel_cluster_kernel(cplx* A, int n, int j0, int nb, int* __restrict__ piv,
                                                                      int* __restrict__ singular) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  __shared__ LpcShared sh;
  extern __shared__ __align__(16) unsigned char lpc_dyn[];
  cplx(*Lrows)[SUBW] = reinterpret_cast<cplx(*)[SUBW]>(lpc_dyn);   // [RPT * THREADS][8]: the L entries of this CTA's rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int pe = j0 + nb;
  const int64_t ld = n;
  int ri[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) ri[u] = j0 + (u * C + rank) * THREADS + tid;
  cplx row[RPT][SUBW];
  double cval[RPT];

  for (int s0 = j0; s0 < pe; s0 += SUBW) {
    const int w = pe - s0 < SUBW ? pe - s0 : SUBW;
    const int c_sub = s0 + w;
    // ---- this thread's rows of the sub-panel -> registers; candidates of its first column
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const bool live = ri[u] >= s0 && ri[u] < n;
#pragma unroll
      for (int q = 0; q < SUBW; ++q)
        row[u][q] = (live && q < w) ? ldcg(A + (int64_t)ri[u] * ld + s0 + q) : make_double2(0.0, 0.0);
      cval[u] = live ? fabs(row[u][0].x) + fabs(row[u][0].y) : -1.0;
    }
#pragma unroll
    for (int jc = 0; jc < SUBW; ++jc) {
      if (jc < w) {
        const int j = s0 + jc, par = jc & 1;
        // ---- this CTA's candidate: thread -> warp -> CTA
        double best = -1.0;
        int bi = NOROW;
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          if (cval[u] >= 0.0 && (cval[u] > best || (cval[u] == best && ri[u] < bi))) {
            best = cval[u];
            bi = ri[u];
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
          }
        }
        if (lane == 0) {
          sh.w_val[warp] = best;
          sh.w_idx[warp] = bi;
        }
        __syncthreads();
        double cb = -1.0;
        int ci = NOROW;
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          const double v = sh.w_val[q];
          const int idx = sh.w_idx[q];
          if (v > cb || (v == cb && idx < ci)) {
            cb = v;
            ci = idx;
          }
        }
        // ---- push it (and its row) to every CTA of the cluster; the owner of row j pushes row j
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
          if (ci != NOROW && ri[u] == ci) {
            for (int r = 0; r < C; ++r) {
              LpcSlot* dst = cluster.map_shared_rank(&sh.slot[par][rank], r);
              dst->val = cb;
              dst->idx = ci;
#pragma unroll
              for (int q = 0; q < SUBW; ++q) dst->row[q] = row[u][q];
            }
          }
        }
        if (ci == NOROW && tid == 0) {
          for (int r = 0; r < C; ++r) {
            LpcSlot* dst = cluster.map_shared_rank(&sh.slot[par][rank], r);
            dst->val = -1.0;
            dst->idx = NOROW;
          }
        }
        if (rank == 0 && tid == j - j0) {   // row j = j0 + tid is row u = 0 of this thread
          for (int r = 0; r < C; ++r) {
            cplx* dst = cluster.map_shared_rank(&sh.rowj[par][0], r);
#pragma unroll
            for (int q = 0; q < SUBW; ++q) dst[q] = row[0][q];
          }
        }
        cluster.sync();
        // ---- the pivot: every thread reduces the C candidates
        double pb = -1.0;
        int p = NOROW, win = 0;
        for (int r = 0; r < C; ++r) {
          const double v = sh.slot[par][r].val;
          const int idx = sh.slot[par][r].idx;
          if (v > pb || (v == pb && idx < p)) {
            pb = v;
            p = idx;
            win = r;
          }
        }
        if (p == NOROW) p = j;
        cplx prow[SUBW];
#pragma unroll
        for (int q = 0; q < SUBW; ++q) prow[q] = p == j ? sh.rowj[par][q] : sh.slot[par][win].row[q];
        if (tid == 0) {
          sh.piv[jc] = p;
          if (rank == 0) {
            piv[j] = p;
            if (!(pb > 0.0)) *singular = 1;
          }
        }
        // ---- interchange j <-> p inside the sub-panel (registers)
        if (p != j) {
#pragma unroll
          for (int u = 0; u < RPT; ++u)
            if (ri[u] == p) {
#pragma unroll
              for (int q = 0; q < SUBW; ++q) row[u][q] = sh.rowj[par][q];
            }
          if (rank == 0 && tid == j - j0) {
#pragma unroll
            for (int q = 0; q < SUBW; ++q) row[0][q] = prow[q];
          }
        }
        // ---- scale column j, rank-1 update of the rest of the sub-panel, candidates of column j + 1
        const cplx d = prow[jc];
        const cplx inv = (d.x != 0.0 || d.y != 0.0) ? cdiv(make_double2(1.0, 0.0), d) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
          if (ri[u] > j && ri[u] < n) {
            const cplx l = cmul(row[u][jc], inv);
            row[u][jc] = l;
#pragma unroll
            for (int q = jc + 1; q < SUBW; ++q)
              if (q < w) row[u][q] = csub(row[u][q], cmul(l, prow[q]));
            cval[u] = (jc + 1 < SUBW && jc + 1 < w) ? fabs(row[u][jc + 1 < SUBW ? jc + 1 : jc].x) + fabs(row[u][jc + 1 < SUBW ? jc + 1 : jc].y)
                                                   : -1.0;
          } else {
            cval[u] = -1.0;
          }
        }
      }
    }
    // ---- the finished sub-panel goes back to global memory (and its L entries to shared memory for the update below)
#pragma unroll
    for (int u = 0; u < RPT; ++u)
      if (ri[u] >= s0 && ri[u] < n) {
#pragma unroll
        for (int q = 0; q < SUBW; ++q) {
          if (q < w) stcg(A + (int64_t)ri[u] * ld + s0 + q, row[u][q]);
          Lrows[u * THREADS + tid][q] = row[u][q];
        }
      }
    if (c_sub >= pe) break;
    const int ncols = pe - c_sub;
    __syncthreads();
    if (rank == 0) {
      // ---- CTA 0: the w interchanges on the panel columns to the right, then the U block row
      for (int e = tid; e < w * w; e += THREADS) sh.L[e / w][e % w] = ldcg(A + (int64_t)(s0 + e / w) * ld + s0 + e % w);
      for (int e = tid; e < w * ncols; e += THREADS)
        sh.T[e / ncols][e % ncols] = ldcg(A + (int64_t)(s0 + e / ncols) * ld + c_sub + e % ncols);
      if (tid < w) {
        const int p = sh.piv[tid];
        int slot = -1;
        if (p >= c_sub) {
          slot = tid;
          for (int k = 0; k < tid; ++k)
            if (sh.piv[k] == p) {
              slot = k;
              break;
            }
        }
        sh.gslot[tid] = slot;
      }
      __syncthreads();
      for (int e = tid; e < w * ncols; e += THREADS) {
        const int k = e / ncols;
        if (sh.gslot[k] == k) sh.G[k][e % ncols] = ldcg(A + (int64_t)sh.piv[k] * ld + c_sub + e % ncols);
      }
      __syncthreads();
      if (tid < ncols) {
        for (int k = 0; k < w; ++k) {
          const int p = sh.piv[k];
          if (p == s0 + k) continue;
          cplx* other = p < c_sub ? &sh.T[p - s0][tid] : &sh.G[sh.gslot[k]][tid];
          const cplx t = sh.T[k][tid];
          sh.T[k][tid] = *other;
          *other = t;
        }
        for (int r = 1; r < w; ++r) {
          cplx acc = sh.T[r][tid];
          for (int k = 0; k < r; ++k) acc = csub(acc, cmul(sh.L[r][k], sh.T[k][tid]));
          sh.T[r][tid] = acc;
        }
      }
      __syncthreads();
      for (int e = tid; e < w * ncols; e += THREADS) {
        const int k = e / ncols;
        stcg(A + (int64_t)(s0 + k) * ld + c_sub + e % ncols, sh.T[k][e % ncols]);
        if (sh.gslot[k] == k) stcg(A + (int64_t)sh.piv[k] * ld + c_sub + e % ncols, sh.G[k][e % ncols]);
      }
    }
    cluster.sync();
    if (rank != 0) {
      for (int e = tid; e < w * ncols; e += THREADS)
        sh.T[e / ncols][e % ncols] = ldcg(A + (int64_t)(s0 + e / ncols) * ld + c_sub + e % ncols);
    }
    __syncthreads();
    // ---- rank-w update of this CTA's rows on the remaining panel columns: one warp per row, lanes along the columns
    // (coalesced 512-byte accesses; a thread updating its own row issued eight 16-byte requests per cache line and the
    // update took 45 of a sub-panel's 90 us), the row's L entries broadcast from shared memory, U in registers
    for (int h = 0; h * 32 < ncols; ++h) {
      const int c = lane + 32 * h;
      cplx Ur[SUBW];
#pragma unroll
      for (int t = 0; t < SUBW; ++t) Ur[t] = (c < ncols && t < w) ? sh.T[t][c & (PANEL - 1)] : make_double2(0.0, 0.0);
      for (int lr0 = warp; lr0 < RPT * THREADS; lr0 += 4 * NW) {
        cplx a[4];
        int64_t off[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int lr = lr0 + v * NW;
          const int i = j0 + ((lr / THREADS) * C + rank) * THREADS + lr % THREADS;
          off[v] = (lr < RPT * THREADS && i >= c_sub && i < n && c < ncols) ? (int64_t)i * ld + c_sub + c : -1;
          a[v] = off[v] >= 0 ? ldcg(A + off[v]) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (off[v] >= 0) {
            const int lr = lr0 + v * NW;
#pragma unroll
            for (int t = 0; t < SUBW; ++t) a[v] = csub(a[v], cmul(Lrows[lr][t], Ur[t]));
            stcg(A + off[v], a[v]);
          }
        }
      }
    }
    __syncthreads();   // sh.T is rewritten by the next sub-panel
  }
}

// The row interchanges of one panel, applied to every column the panel kernel did not cover: all columns outside the
// panel take all nb interchanges, a panel column of sub-panel q takes those of the later sub-panels.  One CTA per 64
// columns; the nb panel rows and the rows interchanged in from below are staged in shared memory, the interchanges are
// replayed there (one thread per column), and both sets go back: one global read and one write per element instead of
// nb dependent round trips.
constexpr int LSW_COLS = 64;
__global__ void __launch_bounds__(256) lu_laswp_kernel(cplx* A, int n, int j0, int nb, const int* __restrict__ piv) {
  extern __shared__ __align__(16) unsigned char lsw_smem[];
  cplx(*T)[LSW_COLS] = reinterpret_cast<cplx(*)[LSW_COLS]>(lsw_smem);
  cplx(*G)[LSW_COLS] = T + PANEL;
  __shared__ int s_piv[PANEL], s_slot[PANEL];
  __shared__ int s_any;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * LSW_COLS, pe = j0 + nb;
  const int64_t ld = n;
  if (tid == 0) s_any = 0;
  __syncthreads();
  if (tid < nb) {
    s_piv[tid] = piv[j0 + tid];
    if (s_piv[tid] != j0 + tid) s_any = 1;
  }
  __syncthreads();
  if (!s_any) return;
  if (tid < nb) {
    const int p = s_piv[tid];
    int slot = -1;
    if (p >= pe) {
      slot = tid;
      for (int k = 0; k < tid; ++k)
        if (s_piv[k] == p) {
          slot = k;
          break;
        }
    }
    s_slot[tid] = slot;
  }
  __syncthreads();
  for (int e = tid; e < nb * LSW_COLS; e += 256) {
    const int k = e / LSW_COLS, c = e % LSW_COLS;
    if (c0 + c < n) {
      T[k][c] = A[(int64_t)(j0 + k) * ld + c0 + c];
      if (s_slot[k] == k) G[k][c] = A[(int64_t)s_piv[k] * ld + c0 + c];
    }
  }
  __syncthreads();
  if (tid < LSW_COLS && c0 + tid < n) {
    const int col = c0 + tid;
    const int ks = (col >= j0 && col < pe) ? SUBW * ((col - j0) / SUBW + 1) : 0;
    for (int k = ks; k < nb; ++k) {
      const int p = s_piv[k];
      if (p == j0 + k) continue;
      cplx* other = p < pe ? &T[p - j0][tid] : &G[s_slot[k]][tid];
      const cplx t = T[k][tid];
      T[k][tid] = *other;
      *other = t;
    }
  }
  __syncthreads();
  for (int e = tid; e < nb * LSW_COLS; e += 256) {
    const int k = e / LSW_COLS, c = e % LSW_COLS;
    if (c0 + c < n) {
      A[(int64_t)(j0 + k) * ld + c0 + c] = T[k][c];
      if (s_slot[k] == k) A[(int64_t)s_piv[k] * ld + c0 + c] = G[k][c];
    }
  }
}

// B <- L^-1 B for the unit-lower nb x nb block L = A[j0.., j0..] and the nb x ncols block B = A[j0.., c_first..] (the U
// block row of a panel): one thread per column, L and the CTA's 64 columns of B staged in shared memory.
constexpr int TRSM_COLS = 64;
constexpr int TRSM_THREADS = 4 * TRSM_COLS;   // four threads per column share a row's dot product
__global__ void __launch_bounds__(TRSM_THREADS) lu_trsm_unit_lower_kernel(cplx* A, int n, int j0, int nb, int c_first, int ncols) {
  extern __shared__ __align__(16) unsigned char trsm_smem[];
  cplx(*Ls)[PANEL] = reinterpret_cast<cplx(*)[PANEL]>(trsm_smem);
  cplx(*Bs)[TRSM_COLS] = reinterpret_cast<cplx(*)[TRSM_COLS]>(trsm_smem + sizeof(cplx) * PANEL * PANEL);
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * TRSM_COLS;
  const int64_t ld = n;
  for (int e = tid; e < nb * nb; e += TRSM_THREADS) Ls[e / nb][e % nb] = A[(int64_t)(j0 + e / nb) * ld + j0 + e % nb];
  for (int e = tid; e < nb * TRSM_COLS; e += TRSM_THREADS) {
    const int r = e / TRSM_COLS, c = e % TRSM_COLS;
    Bs[r][c] = c0 + c < ncols ? A[(int64_t)(j0 + r) * ld + c_first + c0 + c] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  {
    const int c = tid >> 2, h = tid & 3;   // the four threads of a column are neighbouring lanes of one warp
    for (int r = 1; r < nb; ++r) {
      cplx acc = make_double2(0.0, 0.0);
      for (int k = h; k < r; k += 4) acc = csub(acc, cmul(Ls[r][k], Bs[k][c]));
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
      if (h == 0) {
        const cplx b = Bs[r][c];
        Bs[r][c] = make_double2(b.x + acc.x, b.y + acc.y);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < nb * TRSM_COLS; e += TRSM_THREADS) {
    const int r = e / TRSM_COLS, c = e % TRSM_COLS;
    if (r > 0 && c0 + c < ncols) A[(int64_t)(j0 + r) * ld + c_first + c0 + c] = Bs[r][c];
  }
}

template <int THREADS, int RPT>
int launch_panel(cplx* A, int n, int j0, int nb, int* piv, int* singular, int csize, cudaStream_t stream) {
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(lu_panel_cluster_kernel<THREADS, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CARC_CHECK_CUDA(cudaFuncSetAttribute(lu_panel_cluster_kernel<THREADS, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         RPT * THREADS * SUBW * (int)sizeof(cplx)));
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = RPT * THREADS * SUBW * sizeof(cplx);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CARC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel<THREADS, RPT>, A, n, j0, nb, piv, singular));
  return CARC_OK;
}

// largest cluster the device schedules for the panel kernel (16 if the non-portable size is granted, else 8), cached
int panel_cluster_limit() {
  static int limit[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 16) return 8;
  if (limit[dev]) return limit[dev];
  int best = 8;
  if (cudaFuncSetAttribute(lu_panel_cluster_kernel<512, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = 2 * 512 * SUBW * sizeof(cplx);
    (void)cudaFuncSetAttribute(lu_panel_cluster_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               2 * 512 * SUBW * (int)sizeof(cplx));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, lu_panel_cluster_kernel<512, 2>, &cfg) == cudaSuccess && nclusters >= 1) best = 16;
  }
  (void)cudaGetLastError();
  limit[dev] = best;
  return best;
}

}  // namespace

// Factorise the panel [j0, j0 + nb) of the n x n row-major matrix (rows j0 .. n-1) and apply its interchanges to all other
// columns.  CARC_ERR_UNSUPPORTED (nothing launched) when the panel has more rows than one cluster holds in registers.
int lu_panel_cluster(cplx* A, int n, int j0, int nb, int* piv, int* singular, cudaStream_t stream) {
  static const bool off = getenv("CARC_LU_CLUSTER") && atoi(getenv("CARC_LU_CLUSTER")) == 0;   // experiments
  if (off || nb > PANEL) return CARC_ERR_UNSUPPORTED;
  const int rows = n - j0;
  int rc;
  // up to 4096 rows: 256-thread CTAs and twice as many of them (a column step is issue-bound per SM: n = 2592 24.6 -> 22.0 ms)
  static const bool wide = !(getenv("CARC_LU_WIDE") && atoi(getenv("CARC_LU_WIDE")) == 0);
  if (wide && rows <= 16 * 256 && rows > 256 && panel_cluster_limit() >= 16)
    rc = launch_panel<256, 1>(A, n, j0, nb, piv, singular, rows <= 512 ? 2 : rows <= 1024 ? 4 : rows <= 2048 ? 8 : 16, stream);
  else if (rows <= 8 * 512) rc = launch_panel<512, 1>(A, n, j0, nb, piv, singular, rows <= 512 ? 1 : rows <= 1024 ? 2 : rows <= 2048 ? 4 : 8, stream);
  else if (rows <= 16 * 512 && panel_cluster_limit() >= 16) rc = launch_panel<512, 1>(A, n, j0, nb, piv, singular, 16, stream);
  else if (rows <= 8 * 512 * 2) rc = launch_panel<512, 2>(A, n, j0, nb, piv, singular, 8, stream);
  else if (rows <= 16 * 512 * 2 && panel_cluster_limit() >= 16) rc = launch_panel<512, 2>(A, n, j0, nb, piv, singular, 16, stream);
  else if (rows <= 16 * 256 * 5 && panel_cluster_limit() >= 16) rc = launch_panel<256, 5>(A, n, j0, nb, piv, singular, 16, stream);
  else return CARC_ERR_UNSUPPORTED;
  if (rc) return rc;
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  const int smem = 2 * PANEL * LSW_COLS * (int)sizeof(cplx);
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(lu_laswp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  lu_laswp_kernel<<<(n + LSW_COLS - 1) / LSW_COLS, 256, smem, stream>>>(A, n, j0, nb, piv);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// A[j0 : j0 + nb, c_first : c_first + ncols] <- L11^-1 (same block), L11 = the unit-lower block at (j0, j0); nb <= 64
int lu_trsm_unit_lower(cplx* A, int n, int j0, int nb, int c_first, int ncols, cudaStream_t stream) {
  if (ncols <= 0 || nb <= 1) return CARC_OK;
  CARC_REQUIRE(nb <= PANEL, CARC_ERR_VALUE, "lu_trsm_unit_lower: block of %d rows", nb);
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  const int smem = (int)sizeof(cplx) * PANEL * (PANEL + TRSM_COLS);
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(lu_trsm_unit_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  lu_trsm_unit_lower_kernel<<<(ncols + TRSM_COLS - 1) / TRSM_COLS, TRSM_THREADS, smem, stream>>>(A, n, j0, nb, c_first, ncols);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

}  // namespace carc
