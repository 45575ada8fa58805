// Elementwise, permute ("join") and reduction kernels on complex128 device tensors.
// These are the HBM-bound pieces of the path: NDArrayData.join / conj / += / norm / contractWithAlongAll
// (reference data/__init__.py:247-256, 154, 104-107, 257) and the Arnoldi vector algebra (utils.py:845-870).
#include <stdarg.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_error; }

// ---------------------------------------------------------------------------------------------------
// permute: dst (C-contiguous, axes in `perm` order) <- src.  After host-side canonicalisation (unit axes
// dropped, axes adjacent in both layouts merged) at most CARC_MAX_RANK axes remain.
struct PermuteParams {
  int nd;
  int64_t total;
  int32_t dst_dim[CARC_MAX_RANK];     // extents in dst order
  int64_t src_stride[CARC_MAX_RANK];  // source stride (elements) of each dst axis
};

template <bool CONJ, bool ACCUM>
__global__ void __launch_bounds__(256) permute_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst,
                                                      PermuteParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
    int64_t rem = i, off = 0;
#pragma unroll 1
    for (int a = p.nd - 1; a > 0; --a) {
      int64_t q = rem / p.dst_dim[a];
      int32_t r = (int32_t)(rem - q * p.dst_dim[a]);
      off += r * p.src_stride[a];
      rem = q;
    }
    off += rem * p.src_stride[0];
    cplx v = __ldg(src + off);
    if (CONJ) v.y = -v.y;
    if (ACCUM) {
      cplx o = dst[i];
      v.x += o.x;
      v.y += o.y;
    }
    dst[i] = v;
  }
}

// Tiled variant for the case where the source's unit-stride axis is not the destination's innermost axis:
// a 32x32 tile over (dst innermost axis, src innermost axis) goes through shared memory so both the global
// read and the global write are coalesced.
struct PermuteTiledParams {
  int nd;                              // number of batch axes (the two tile axes excluded)
  int32_t na, nb;                      // extents of dst-inner axis (a) and src-inner axis (b)
  int64_t a_src_stride;                // stride of axis a in the source
  int64_t b_dst_stride;                // stride of axis b in the destination
  int32_t batch_dim[CARC_MAX_RANK];
  int64_t batch_src_stride[CARC_MAX_RANK];
  int64_t batch_dst_stride[CARC_MAX_RANK];
  int32_t tiles_a, tiles_b;
};

template <bool CONJ>
__global__ void __launch_bounds__(256) permute_tiled_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst,
                                                            PermuteTiledParams p) {
  __shared__ cplx tile[32][33];
  int64_t bid = blockIdx.x;
  const int ta = (int)(bid % p.tiles_a);
  bid /= p.tiles_a;
  const int tb = (int)(bid % p.tiles_b);
  bid /= p.tiles_b;
  int64_t soff = 0, doff = 0;
  for (int a = p.nd - 1; a >= 0; --a) {
    int64_t q = bid / p.batch_dim[a];
    int32_t r = (int32_t)(bid - q * p.batch_dim[a]);
    soff += r * p.batch_src_stride[a];
    doff += r * p.batch_dst_stride[a];
    bid = q;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int a0 = ta * 32, b0 = tb * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ia = a0 + ty + 8 * j, ib = b0 + tx;
    if (ia < p.na && ib < p.nb) {
      cplx v = __ldg(src + soff + (int64_t)ia * p.a_src_stride + ib);
      if (CONJ) v.y = -v.y;
      tile[ty + 8 * j][tx] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ib = b0 + ty + 8 * j, ia = a0 + tx;
    if (ia < p.na && ib < p.nb) dst[doff + (int64_t)ib * p.b_dst_stride + ia] = tile[tx][ty + 8 * j];
  }
}

// Block variant for permutations whose trailing destination axes are a rearrangement of the trailing source axes (the
// Hermitian symmetrisation join(1,0,2,4,3,5,7,6) of system/_2d.py:42-47, every join that only swaps small state-bond
// axes): the tensor is an array of m-element blocks that are contiguous in BOTH layouts, so a warp reads a block with
// coalesced 16-byte loads, rearranges it through shared memory with a small index table and writes it coalesced.
// The outer (block) index goes through the generic stride walk once per block instead of once per element.
struct PermuteBlockParams {
  int nd;                              // outer axes
  int32_t m;                           // elements per block
  int64_t nblocks;
  int32_t outer_dim[CARC_MAX_RANK];    // extents of the outer axes in dst order
  int64_t outer_src_stride[CARC_MAX_RANK];
  const int32_t* inner_src;            // [m]: source offset inside the block of destination element j
};

template <bool CONJ>
__global__ void __launch_bounds__(256) permute_block_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst,
                                                            PermuteBlockParams p) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  cplx* buf = reinterpret_cast<cplx*>(pb_smem) + (size_t)warp * p.m;
  int32_t* table = reinterpret_cast<int32_t*>(reinterpret_cast<cplx*>(pb_smem) + (size_t)nwarps * p.m);
  for (int j = threadIdx.x; j < p.m; j += blockDim.x) table[j] = p.inner_src[j];
  __syncthreads();
  for (int64_t blk = (int64_t)blockIdx.x * nwarps + warp; blk < p.nblocks; blk += (int64_t)gridDim.x * nwarps) {
    int64_t rem = blk, off = 0;
#pragma unroll 1
    for (int a = p.nd - 1; a > 0; --a) {
      const int64_t q = rem / p.outer_dim[a];
      off += (rem - q * p.outer_dim[a]) * p.outer_src_stride[a];
      rem = q;
    }
    if (p.nd > 0) off += rem * p.outer_src_stride[0];
    const cplx* s = src + off;
    for (int j = lane; j < p.m; j += 32) buf[j] = __ldg(s + j);
    __syncwarp();
    cplx* d = dst + blk * p.m;
    for (int j = lane; j < p.m; j += 32) {
      cplx v = buf[table[j]];
      if (CONJ) v.y = -v.y;
      d[j] = v;
    }
    __syncwarp();
  }
}

// device copies of inner index tables, keyed by their content (a handful of distinct joins per run)
struct BlockTable {
  std::vector<int32_t> host;
  int32_t* dev;
  int device;
};
static std::vector<BlockTable> g_block_tables;

static int block_table(const std::vector<int32_t>& t, const int32_t** out) {
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  for (const auto& e : g_block_tables)
    if (e.device == dev && e.host == t) {
      *out = e.dev;
      return CARC_OK;
    }
  BlockTable e;
  e.host = t;
  e.device = dev;
  CARC_CHECK_CUDA(cudaMalloc(&e.dev, sizeof(int32_t) * t.size()));
  CARC_CHECK_CUDA(cudaMemcpy(e.dev, t.data(), sizeof(int32_t) * t.size(), cudaMemcpyHostToDevice));
  g_block_tables.push_back(e);
  *out = e.dev;
  return CARC_OK;
}

int permute(const cplx* src, cplx* dst, int ndim, const int64_t* shape, const int32_t* perm, int conj, int accumulate,
            cudaStream_t stream) {
  CARC_REQUIRE(ndim >= 0 && ndim <= 32, CARC_ERR_RANK, "permute: rank %d out of range", ndim);
  // source strides
  int64_t sstride[32];
  int64_t total = 1;
  for (int a = ndim - 1; a >= 0; --a) {
    sstride[a] = total;
    total *= shape[a];
  }
  bool seen[32] = {false};
  for (int a = 0; a < ndim; ++a) {
    CARC_REQUIRE(perm[a] >= 0 && perm[a] < ndim && !seen[perm[a]], CARC_ERR_VALUE, "permute: invalid permutation");
    seen[perm[a]] = true;
  }
  if (total == 0) return CARC_OK;
  // canonicalise in dst order: drop unit axes, merge axes that are adjacent in the source too
  int64_t dim[32], str[32];
  int nd = 0;
  for (int a = 0; a < ndim; ++a) {
    int s = perm[a];
    if (shape[s] == 1) continue;
    if (nd > 0 && str[nd - 1] == sstride[s] * shape[s]) {
      dim[nd - 1] *= shape[s];
      str[nd - 1] = sstride[s];
    } else {
      dim[nd] = shape[s];
      str[nd] = sstride[s];
      ++nd;
    }
  }
  if (nd == 0) {
    dim[0] = 1;
    str[0] = 1;
    nd = 1;
  }
  CARC_REQUIRE(nd <= CARC_MAX_RANK, CARC_ERR_RANK, "permute: %d non-mergeable axes exceed the supported %d", nd,
               CARC_MAX_RANK);
  for (int a = 0; a < nd; ++a)
    CARC_REQUIRE(dim[a] < (1ll << 31), CARC_ERR_VALUE, "permute: axis extent too large");

  // block path: the last k >= 2 dst axes are exactly the last k source axes in another order, 8 <= m <= 1024 elements
  if (!accumulate && nd >= 2 && str[nd - 1] != 1) {
    // source order of the canonical axes: sort by stride (descending)
    int order[32];
    for (int a = 0; a < nd; ++a) order[a] = a;
    for (int a = 0; a < nd; ++a)
      for (int b = a + 1; b < nd; ++b)
        if (str[order[b]] > str[order[a]]) { int t = order[a]; order[a] = order[b]; order[b] = t; }
    int k = 0;
    int64_t m = 1;
    for (int kk = 2; kk <= nd; ++kk) {
      // are dst axes [nd-kk, nd) the same set as the source's last kk axes?
      bool same = true;
      int64_t mm = 1;
      for (int a = nd - kk; a < nd && same; ++a) {
        bool found = false;
        for (int b = nd - kk; b < nd; ++b) found = found || order[b] == a;
        same = found;
        mm *= dim[a];
      }
      if (same) { k = kk; m = mm; break; }
    }
    if (k >= 2 && k < nd && m >= 8 && m <= 1024) {
      // inside a block the source is contiguous (its last k axes), so offsets are the strides themselves
      std::vector<int32_t> table((size_t)m);
      for (int64_t j = 0; j < m; ++j) {
        int64_t rem = j, off = 0;
        for (int a = nd - 1; a >= nd - k; --a) {
          off += (rem % dim[a]) * str[a];
          rem /= dim[a];
        }
        table[(size_t)j] = (int32_t)off;
      }
      PermuteBlockParams p;
      int rc = block_table(table, &p.inner_src);
      if (rc) return rc;
      p.nd = nd - k;
      p.m = (int32_t)m;
      p.nblocks = total / m;
      for (int a = 0; a < nd - k; ++a) {
        p.outer_dim[a] = (int32_t)dim[a];
        p.outer_src_stride[a] = str[a];
      }
      const int threads = 256, nwarps = threads / 32;
      const size_t smem = (size_t)nwarps * m * sizeof(cplx) + (size_t)m * sizeof(int32_t);
      static bool configured[16] = {false};
      int dev = 0;
      CARC_CHECK_CUDA(cudaGetDevice(&dev));
      if (dev < 16 && !configured[dev]) {
        CARC_CHECK_CUDA(cudaFuncSetAttribute(permute_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CARC_CHECK_CUDA(cudaFuncSetAttribute(permute_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured[dev] = true;
      }
      int64_t want = (p.nblocks + nwarps - 1) / nwarps;
      unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
      if (conj)
        permute_block_kernel<true><<<blocks, threads, smem, stream>>>(src, dst, p);
      else
        permute_block_kernel<false><<<blocks, threads, smem, stream>>>(src, dst, p);
      CARC_CHECK_CUDA(cudaGetLastError());
      return CARC_OK;
    }
  }

  // tiled path: source-innermost axis exists in dst at position j != nd-1 with reasonable extents
  if (!accumulate && nd >= 2 && str[nd - 1] != 1) {
    int jb = -1;
    for (int a = 0; a < nd; ++a)
      if (str[a] == 1) jb = a;
    if (jb >= 0 && dim[jb] >= 8 && dim[nd - 1] >= 8) {
      PermuteTiledParams p;
      p.na = (int32_t)dim[nd - 1];
      p.nb = (int32_t)dim[jb];
      p.a_src_stride = str[nd - 1];
      int64_t dstride[32];
      int64_t acc = 1;
      for (int a = nd - 1; a >= 0; --a) {
        dstride[a] = acc;
        acc *= dim[a];
      }
      p.b_dst_stride = dstride[jb];
      p.nd = 0;
      int64_t nbatch = 1;
      for (int a = 0; a < nd - 1; ++a) {
        if (a == jb) continue;
        p.batch_dim[p.nd] = (int32_t)dim[a];
        p.batch_src_stride[p.nd] = str[a];
        p.batch_dst_stride[p.nd] = dstride[a];
        nbatch *= dim[a];
        ++p.nd;
      }
      p.tiles_a = (p.na + 31) / 32;
      p.tiles_b = (p.nb + 31) / 32;
      int64_t blocks = nbatch * p.tiles_a * p.tiles_b;
      if (blocks < (1ll << 31)) {
        if (conj)
          permute_tiled_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(src, dst, p);
        else
          permute_tiled_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(src, dst, p);
        CARC_CHECK_CUDA(cudaGetLastError());
        return CARC_OK;
      }
    }
  }

  PermuteParams p;
  p.nd = nd;
  p.total = total;
  for (int a = 0; a < nd; ++a) {
    p.dst_dim[a] = (int32_t)dim[a];
    p.src_stride[a] = str[a];
  }
  int64_t blocks64 = (total + 255) / 256;
  unsigned blocks = (unsigned)(blocks64 < 148 * 16 ? blocks64 : 148 * 16);
  if (conj && accumulate)
    permute_kernel<true, true><<<blocks, 256, 0, stream>>>(src, dst, p);
  else if (conj)
    permute_kernel<true, false><<<blocks, 256, 0, stream>>>(src, dst, p);
  else if (accumulate)
    permute_kernel<false, true><<<blocks, 256, 0, stream>>>(src, dst, p);
  else
    permute_kernel<false, false><<<blocks, 256, 0, stream>>>(src, dst, p);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// ---------------------------------------------------------------------------------------------------
// y = alpha * x + beta * y   (alpha, beta complex host scalars); vectorised 16-byte accesses
__global__ void __launch_bounds__(256) axpby_kernel(int64_t n, cplx alpha, const cplx* __restrict__ x, cplx beta,
                                                    cplx* __restrict__ y, int conj_x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool beta_zero = (beta.x == 0.0 && beta.y == 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx a = x[i];
    if (conj_x) a.y = -a.y;
    cplx r;
    r.x = alpha.x * a.x - alpha.y * a.y;
    r.y = alpha.x * a.y + alpha.y * a.x;
    if (!beta_zero) {
      cplx b = y[i];
      r.x += beta.x * b.x - beta.y * b.y;
      r.y += beta.x * b.y + beta.y * b.x;
    }
    y[i] = r;
  }
}

static unsigned grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  return (unsigned)(b < 148 * 8 ? b : 148 * 8);
}

int axpby(int64_t n, cplx alpha, const cplx* x, cplx beta, cplx* y, int conj_x, cudaStream_t stream) {
  if (n <= 0) return CARC_OK;
  axpby_kernel<<<grid_for(n), 256, 0, stream>>>(n, alpha, x, beta, y, conj_x);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// elementwise product y *= x (NDArrayData.__imul__ / __mul__)
__global__ void __launch_bounds__(256) mul_kernel(int64_t n, const cplx* __restrict__ x, cplx* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx a = x[i], b = y[i], r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    y[i] = r;
  }
}
int mul_inplace(int64_t n, const cplx* x, cplx* y, cudaStream_t stream) {
  if (n <= 0) return CARC_OK;
  mul_kernel<<<grid_for(n), 256, 0, stream>>>(n, x, y);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Deterministic reductions.  Stage 1: one partial per block (warp shuffles, fixed order); stage 2: the
// last block to finish (ticket) sums the partials in index order.  out = sum conj(x_i) * y_i.
#define RED_MAX_BLOCKS 1184  // 148 * 8

struct ReduceWorkspace {
  double2* partials;       // RED_MAX_BLOCKS
  unsigned int* ticket;    // zero between launches
};

__device__ __forceinline__ void block_reduce2(double& a, double& b, double* sh) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    sh[2 * w] = a;
    sh[2 * w + 1] = b;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? sh[2 * lane] : 0.0;
    b = lane < nw ? sh[2 * lane + 1] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

template <int MODE>  // 0: dotc(x,y)  1: sum |x|^2 (real)  2: nan/inf count  3: dotu(x,y) = sum x*y
__global__ void __launch_bounds__(256) reduce_kernel(int64_t n, const cplx* __restrict__ x, const cplx* __restrict__ y,
                                                     double2* __restrict__ out, ReduceWorkspace ws) {
  __shared__ double sh[16];
  __shared__ bool last;
  double a = 0.0, b = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx u = x[i];
    if (MODE == 0) {
      cplx v = y[i];
      a += u.x * v.x + u.y * v.y;
      b += u.x * v.y - u.y * v.x;
    } else if (MODE == 3) {
      cplx v = y[i];
      a += u.x * v.x - u.y * v.y;
      b += u.x * v.y + u.y * v.x;
    } else if (MODE == 1) {
      a += u.x * u.x + u.y * u.y;
    } else {
      a += (isfinite(u.x) && isfinite(u.y)) ? 0.0 : 1.0;
      b += (isnan(u.x) || isnan(u.y)) ? 1.0 : 0.0;
    }
  }
  block_reduce2(a, b, sh);
  if (threadIdx.x == 0) {
    ws.partials[blockIdx.x] = make_double2(a, b);
    __threadfence();
    unsigned t = atomicAdd(ws.ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      double2 p = ws.partials[i];
      sa += p.x;
      sb += p.y;
    }
    __syncthreads();
    block_reduce2(sa, sb, sh);
    if (threadIdx.x == 0) {
      *out = make_double2(sa, sb);
      *ws.ticket = 0u;
    }
  }
}

// One workspace per (device, stream): reductions on the same stream are ordered, so they can share partials and the
// ticket; reductions on different streams (or issued by different host threads) must not interleave tickets.
static std::mutex g_ws_mutex;
static std::map<std::pair<int, cudaStream_t>, ReduceWorkspace> g_ws;

static int get_ws(ReduceWorkspace* ws, cudaStream_t stream) {
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  auto key = std::make_pair(dev, stream);
  auto it = g_ws.find(key);
  if (it == g_ws.end()) {
    ReduceWorkspace w;
    CARC_CHECK_CUDA(cudaMalloc(&w.partials, sizeof(double2) * RED_MAX_BLOCKS));
    CARC_CHECK_CUDA(cudaMalloc(&w.ticket, sizeof(unsigned int)));
    CARC_CHECK_CUDA(cudaMemset(w.ticket, 0, sizeof(unsigned int)));
    it = g_ws.emplace(key, w).first;
  }
  *ws = it->second;
  return CARC_OK;
}

int reduce(int mode, int64_t n, const cplx* x, const cplx* y, double2* out_dev, cudaStream_t stream) {
  ReduceWorkspace ws;
  int rc = get_ws(&ws, stream);
  if (rc) return rc;
  unsigned g = grid_for(n);
  if (mode == 0)
    reduce_kernel<0><<<g, 256, 0, stream>>>(n, x, y, out_dev, ws);
  else if (mode == 1)
    reduce_kernel<1><<<g, 256, 0, stream>>>(n, x, y, out_dev, ws);
  else if (mode == 3)
    reduce_kernel<3><<<g, 256, 0, stream>>>(n, x, y, out_dev, ws);
  else
    reduce_kernel<2><<<g, 256, 0, stream>>>(n, x, y, out_dev, ws);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Mode product with a short matrix: out[pre][j][post] = sum_k M[j][k] * x[pre][k][post], j <= 16 rows.
// (NDArrayData.absorbMatrixAt, data/__init__.py:148-150, when a compressor [new, old] projects an enlarged
// environment bond: new = chi is a handful of rows while pre*post runs to 10^7..10^9.)  A DMMA GEMM would pad the
// j rows to its 128-row tile; here every thread owns one `post` column, streams the k inputs once (coalesced
// 16-byte loads, four in flight) and keeps the j outputs in registers -- the kernel is bound by reading x.
template <int J>
__global__ void __launch_bounds__(256, 2) mode_product_kernel(const cplx* __restrict__ M, const cplx* __restrict__ x,
                                                           cplx* __restrict__ out, int j, int k, int64_t post,
                                                           int64_t chunks) {
  extern __shared__ __align__(16) unsigned char mp_smem[];
  cplx* Ms = reinterpret_cast<cplx*>(mp_smem);   // [k][J], rows beyond j are zero
  for (int i = threadIdx.x; i < k * J; i += blockDim.x) {
    const int kk = i / J, jj = i % J;
    Ms[i] = jj < j ? M[(int64_t)jj * k + kk] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const int64_t b = blockIdx.x;
  const int64_t pre = b / chunks, col = (b % chunks) * blockDim.x + threadIdx.x;
  if (col >= post) return;
  const cplx* xp = x + (pre * k) * post + col;
  double re[J], im[J];
#pragma unroll
  for (int jj = 0; jj < J; ++jj) re[jj] = im[jj] = 0.0;
  int kk = 0;
  for (; kk + 4 <= k; kk += 4) {
    cplx v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(xp + (int64_t)(kk + u) * post);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int jj = 0; jj < J; ++jj) {
        const cplx m = Ms[(kk + u) * J + jj];
        re[jj] += m.x * v[u].x - m.y * v[u].y;
        im[jj] += m.x * v[u].y + m.y * v[u].x;
      }
    }
  }
  for (; kk < k; ++kk) {
    const cplx v = __ldg(xp + (int64_t)kk * post);
#pragma unroll
    for (int jj = 0; jj < J; ++jj) {
      const cplx m = Ms[kk * J + jj];
      re[jj] += m.x * v.x - m.y * v.y;
      im[jj] += m.x * v.y + m.y * v.x;
    }
  }
  cplx* op = out + (pre * j) * post + col;
#pragma unroll
  for (int jj = 0; jj < J; ++jj)
    if (jj < j) op[(int64_t)jj * post] = make_double2(re[jj], im[jj]);
}

template <int J>
static int mode_product_launch(const cplx* M, const cplx* x, cplx* out, int j, int k, int64_t pre, int64_t post,
                               cudaStream_t stream) {
  const size_t smem = sizeof(cplx) * (size_t)k * J;
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(mode_product_kernel<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured[dev] = true;
  }
  const int64_t chunks = (post + 255) / 256;
  const int64_t blocks = pre * chunks;
  CARC_REQUIRE(blocks < (1ll << 31), CARC_ERR_VALUE, "mode product: tensor too large for one launch");
  mode_product_kernel<J><<<(unsigned)blocks, 256, smem, stream>>>(M, x, out, j, k, post, chunks);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int mode_product(const cplx* M, const cplx* x, cplx* out, int64_t j, int64_t k, int64_t pre, int64_t post,
                 cudaStream_t stream) {
  CARC_REQUIRE(j >= 1 && j <= 16, CARC_ERR_UNSUPPORTED, "mode product: 1 <= rows <= 16 required (given %lld)", (long long)j);
  CARC_REQUIRE(k >= 1 && k * 16 * 16 <= 160 * 1024, CARC_ERR_UNSUPPORTED, "mode product: contracted extent %lld too large",
               (long long)k);
  if (pre <= 0 || post <= 0) return CARC_OK;
  if (j <= 4) return mode_product_launch<4>(M, x, out, (int)j, (int)k, pre, post, stream);
  if (j <= 8) return mode_product_launch<8>(M, x, out, (int)j, (int)k, pre, post, stream);
  return mode_product_launch<16>(M, x, out, (int)j, (int)k, pre, post, stream);
}

}  // namespace carc
