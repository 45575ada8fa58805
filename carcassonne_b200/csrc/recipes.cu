// High-level entry points of the C ABI: one call per recipe of the reference's boundary algebra, per normalizeAxis and
// per state-bond compression (carc_absorb_side_into_corner, carc_absorb_center_into_side,
// carc_env_build's two stages, carc_normalize_axis, carc_product_compressor).
//
// Each recipe of carcassonne/tensors/_2d/dense.py is `tensordot` + `join`; here it is one (batched) DMMA GEMM whose
// epilogue scatters through two offset tables (zgemm.cu), so the joined layout is written directly.  The tables depend
// only on the tensor shapes and are cached on the device for the life of the process.  carc_product_compressor runs the
// whole alternating-least-squares fit of compression.py:26-45 in Gram form (never forming the (l r) x (old new) matrix)
// without returning to the host between rounds: before, the Python layer issued ~150 library calls per compression and
// a small sweep iteration spent two thirds of its time there.
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/carc_b200.h"
#include "carc_internal.h"
#include "common.cuh"

namespace carc {
namespace {

typedef std::vector<std::pair<int64_t, int64_t>> Levels;   // (extent, stride) per level, row-major digits

// offset table of a level list, built once per (device, levels) and kept
int offset_table(const Levels& levels, const int64_t** out, cudaStream_t stream) {
  static std::mutex mutex;
  static std::map<std::pair<int, Levels>, int64_t*> cache;
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mutex);
  auto key = std::make_pair(dev, levels);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CARC_REQUIRE((int)levels.size() <= CARC_MAX_RANK, CARC_ERR_RANK, "offset table: too many levels");
    int64_t total = 1, extents[CARC_MAX_RANK], strides[CARC_MAX_RANK];
    for (size_t i = 0; i < levels.size(); ++i) {
      extents[i] = levels[i].first;
      strides[i] = levels[i].second;
      total *= levels[i].first;
    }
    int64_t* table = nullptr;
    CARC_CHECK_CUDA(cudaMalloc(&table, sizeof(int64_t) * (size_t)std::max<int64_t>(total, 1)));
    int rc = index_table((int)levels.size(), extents, strides, table, stream);
    if (rc) return rc;
    // the table may be used on another stream later: make it visible before it is handed out
    CARC_CHECK_CUDA(cudaStreamSynchronize(stream));
    it = cache.emplace(key, table).first;
  }
  *out = it->second;
  return CARC_OK;
}

std::vector<int64_t> row_major_strides(const std::vector<int64_t>& dims) {
  std::vector<int64_t> st(dims.size());
  int64_t acc = 1;
  for (int i = (int)dims.size() - 1; i >= 0; --i) {
    st[i] = acc;
    acc *= dims[i];
  }
  return st;
}

int64_t prod(const int64_t* s, int from, int to) {
  int64_t p = 1;
  for (int i = from; i < to; ++i) p *= s[i];
  return p;
}

const cplx ONE = {1.0, 0.0}, ZERO = {0.0, 0.0};

int gemm_scatter(int opA, int opB, int64_t M, int64_t N, int64_t K, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                 cplx* C, const Levels& rows, const Levels& cols, bool accumulate, int64_t batch, int64_t strideB,
                 int64_t strideC, cudaStream_t stream) {
  if (M * N * batch == 0) return CARC_OK;
  const int64_t *rt = nullptr, *ct = nullptr;
  int rc = offset_table(rows, &rt, stream);
  if (rc) return rc;
  rc = offset_table(cols, &ct, stream);
  if (rc) return rc;
  for (int64_t done = 0; done < batch; done += 32768) {   // the batch index is a grid dimension
    const int64_t nb = std::min<int64_t>(32768, batch - done);
    rc = zgemm(opA, opB, M, N, K, ONE, A, lda, B + done * strideB, ldb, accumulate ? ONE : ZERO, C + done * strideC, nullptr,
               nullptr, nb, 0, strideB, strideC, stream, rt, ct);
    if (rc) return rc;
  }
  return CARC_OK;
}

int plain_gemm(int opA, int opB, int64_t M, int64_t N, int64_t K, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
               cplx* C, bool accumulate, cudaStream_t stream, int64_t batch = 1, int64_t strideA = 0, int64_t strideB = 0,
               int64_t strideC = 0) {
  if (M * N * batch == 0) return CARC_OK;
  return zgemm(opA, opB, M, N, K, ONE, A, lda, B, ldb, accumulate ? ONE : ZERO, C, nullptr, nullptr, batch, strideA, strideB,
               strideC, stream);
}

// scratch memory of one call: stream-ordered allocations released together
struct Scratch {
  cudaStream_t stream;
  std::vector<void*> blocks;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  ~Scratch() {
    for (void* b : blocks) cudaFreeAsync(b, stream);
  }
  int get(cplx** out, int64_t elems) {
    void* p = nullptr;
    CARC_CHECK_CUDA(cudaMallocAsync(&p, sizeof(cplx) * (size_t)std::max<int64_t>(elems, 1), stream));
    blocks.push_back(p);
    *out = static_cast<cplx*>(p);
    return CARC_OK;
  }
};
#define CARC_TRY(expr)     \
  do {                     \
    int _rc = (expr);      \
    if (_rc) return _rc;   \
  } while (0)

// out[pre][j][post] = sum_k matrix'[j][k] x[pre][k][post] with matrix' = op(matrix) (NDArrayData.absorbMatrixAt)
int absorb_matrix(const cplx* x, int64_t pre, int64_t k, int64_t post, int op, const cplx* matrix, int64_t j, int64_t ld,
                  cplx* out, cudaStream_t stream) {
  if (post == 1) {
    // last axis: one GEMM out[pre, j] = x[pre, k] . op(matrix)^T instead of `pre` matrix-vector products
    const int opB = op == OP_N ? OP_T : op == OP_J ? OP_C : op == OP_T ? OP_N : OP_J;
    return plain_gemm(OP_N, opB, pre, j, k, x, k, matrix, ld, out, false, stream);
  }
  return plain_gemm(op, OP_N, j, post, k, matrix, ld, x, post, out, false, stream, pre, 0, k * post, j * post);
}

// polar isometry of a tall matrix A [m, n] (utils.py:879-881 `unitize`): Householder QR, Jacobi SVD of R, Q (U V^H)
int unitize_tall(const cplx* A, int64_t m, int n, cplx* out, Scratch& s) {
  cplx *work, *R, *Q, *tau, *U, *S, *Vh, *pieces;
  CARC_TRY(s.get(&work, m * n));
  CARC_TRY(s.get(&R, (int64_t)n * n));
  CARC_TRY(s.get(&Q, m * n));
  CARC_TRY(s.get(&tau, n));
  CARC_TRY(s.get(&U, (int64_t)n * n));
  CARC_TRY(s.get(&S, n));
  CARC_TRY(s.get(&Vh, (int64_t)n * n));
  CARC_TRY(s.get(&pieces, (int64_t)5 * n * n));
  CARC_CHECK_CUDA(cudaMemcpyAsync(work, A, sizeof(cplx) * m * n, cudaMemcpyDeviceToDevice, s.stream));
  CARC_TRY(qr(work, m, n, R, Q, tau, s.stream));
  CARC_TRY(svd_small(R, n, U, S, Vh, s.stream));
  const int64_t nn = (int64_t)n * n;
  CARC_TRY(normalizer_matrices(U, S, Vh, n, 1e-14, pieces, pieces + nn, pieces + 2 * nn, pieces + 3 * nn, pieces + 4 * nn,
                               s.stream));
  return plain_gemm(OP_N, OP_N, m, n, n, Q, n, pieces, n, out, false, s.stream);
}

// G[i][i] += eps * trace(G) / n  (one block; the shift that keeps the normal equations factorisable, compression.py)
__global__ void __launch_bounds__(256) diagonal_shift_kernel(cplx* G, int n, double eps) {
  __shared__ double sh[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += G[(int64_t)i * n + i].x;
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double shift = eps * sh[0] / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) G[(int64_t)i * n + i].x += shift;
}

// x <- A^-1 x for a small system (n <= 118: the matrix, padded to an odd row stride, and the right-hand side fit the
// shared memory of one SM): LU with partial pivoting (LAPACK's rule: largest |re| + |im|, first on ties) and both
// substitutions in ONE launch of one CTA.  The normal equations of a state-bond compression at small bond dimension are
// this size (old new = 108 at chi = 6, D = 3), four per compression and eight compressions per sweep iteration; the
// blocked device LU spends ~0.4 ms per factorisation on grid barriers there (two cooperative panel launches), this kernel
// a few dozen microseconds.  A is left untouched.
constexpr int SMALL_LU_MAX = 118;
// Two CTA-wide barriers per column: the pivot candidates of column k + 1 are tracked while column k's update runs (one
// warp per row; lane 0 of a row's warp computes that row's new entry in column k + 1) and every warp reduces the 32
// per-warp candidates itself; the column scaling is folded into the update (the row's warp computes its multiplier).
// The back substitution runs on 128 threads only.  (The first version had five barriers per column and two per
// substitution step: 153 us at n = 108, 16 such solves per sweep iteration at D = 3.)
__global__ void __launch_bounds__(1024, 1) small_lu_solve_kernel(const cplx* __restrict__ A, int n, cplx* __restrict__ x) {
  extern __shared__ __align__(16) unsigned char small_lu_raw[];
  const int ld = n + 1;
  cplx* M = reinterpret_cast<cplx*>(small_lu_raw);
  cplx* rhs = M + (size_t)n * ld;
  __shared__ double wval[2][32];
  __shared__ int widx[2][32];
  __shared__ cplx sol[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < n * n; e += blockDim.x) M[(e / n) * ld + e % n] = A[e];
  for (int i = tid; i < n; i += blockDim.x) rhs[i] = x[i];
  __syncthreads();
  if (lane == 0) {   // candidates of column 0
    double best = -1.0;
    int arg = n;
    for (int i = warp; i < n; i += 32) {
      const cplx v = M[i * ld];
      const double a = fabs(v.x) + fabs(v.y);
      if (a > best) {
        best = a;
        arg = i;
      }
    }
    wval[0][warp] = best;
    widx[0][warp] = arg;
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const int par = k & 1;
    // the pivot of column k: largest |re| + |im|, first on ties (every warp reduces the per-warp candidates)
    double best = wval[par][lane];
    int arg = widx[par][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    const int p = arg < n ? arg : k;
    if (p != k) {
      for (int j = tid; j < n; j += blockDim.x) {
        const cplx t = M[k * ld + j];
        M[k * ld + j] = M[p * ld + j];
        M[p * ld + j] = t;
      }
      if (tid == 0) {
        const cplx t = rhs[k];
        rhs[k] = rhs[p];
        rhs[p] = t;
      }
      __syncthreads();
    }
    const cplx piv = M[k * ld + k];
    const double den = piv.x * piv.x + piv.y * piv.y;
    const cplx inv = den > 0.0 ? make_double2(piv.x / den, -piv.y / den) : make_double2(1.0, 0.0);
    // scale, trailing update, forward substitution of the right-hand side, candidates of column k + 1
    double cbest = -1.0;
    int carg = n;
    for (int i = k + 1 + warp; i < n; i += 32) {
      const cplx v = M[i * ld + k];
      const cplx l = make_double2(v.x * inv.x - v.y * inv.y, v.x * inv.y + v.y * inv.x);
      __syncwarp();
      if (lane == 0) M[i * ld + k] = l;
      for (int j = k + 1 + lane; j < n; j += 32) {
        const cplx u = M[k * ld + j];
        cplx c = M[i * ld + j];
        c.x -= l.x * u.x - l.y * u.y;
        c.y -= l.x * u.y + l.y * u.x;
        M[i * ld + j] = c;
        if (j == k + 1) {
          const double a = fabs(c.x) + fabs(c.y);
          if (a > cbest) {
            cbest = a;
            carg = i;
          }
        }
      }
      if (lane == 0) {
        const cplx u = rhs[k];
        rhs[i].x -= l.x * u.x - l.y * u.y;
        rhs[i].y -= l.x * u.y + l.y * u.x;
      }
    }
    if (lane == 0) {
      wval[par ^ 1][warp] = cbest;
      widx[par ^ 1][warp] = carg;
    }
    __syncthreads();
  }
  // back substitution with U on 128 threads (n <= 118): x_k is computed by every thread, thread i < k updates its entry
  if (tid >= 128) return;
  for (int k = n - 1; k >= 0; --k) {
    const cplx piv = M[k * ld + k], b = rhs[k];
    const double den = piv.x * piv.x + piv.y * piv.y;
    const cplx xk = make_double2((b.x * piv.x + b.y * piv.y) / den, (b.y * piv.x - b.x * piv.y) / den);
    if (tid == k) sol[k] = xk;
    if (tid < k) {
      const cplx u = M[tid * ld + k];
      rhs[tid].x -= u.x * xk.x - u.y * xk.y;
      rhs[tid].y -= u.x * xk.y + u.y * xk.x;
    }
    __syncthreads();
  }
  if (tid < n) x[tid] = sol[tid];
}

int small_lu_solve(const cplx* A, int n, cplx* x, cudaStream_t stream) {
  const size_t smem = sizeof(cplx) * ((size_t)n * (n + 1) + n);
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(small_lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    configured[dev] = true;
  }
  small_lu_solve_kernel<<<1, 1024, smem, stream>>>(A, n, x);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

inline int L4(int i) { return (i + 1) % 4; }
inline int R4(int i) { return (i + 3) % 4; }
inline int O4(int i) { return (i + 2) % 4; }

}  // namespace
}  // namespace carc

using carc::cplx;
using carc::Levels;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

// reference tensors/_2d/dense.py:11-21
int carc_absorb_side_into_corner(const void* corner, const int64_t* c, const void* side, const int64_t* s, int from_left,
                                 void* out, int accumulate, void* stream) {
  CARC_REQUIRE(corner && c && side && s && out, CARC_ERR_VALUE, "absorb_side_into_corner: null argument");
  using namespace carc;
  if (from_left) {
    for (int a = 0; a < 3; ++a)
      CARC_REQUIRE(c[a] == s[3 + a], CARC_ERR_DIMENSION_MISMATCH,
                   "tensor 0's index %d has dimension %lld, whereas tensor 1's index %d has dimension %lld", a, (long long)c[a],
                   3 + a, (long long)s[3 + a]);
    const int64_t K = c[0] * c[1] * c[2];
    const auto st = row_major_strides({c[3], s[6], c[4], s[7], c[5]});
    return gemm_scatter(OP_T, OP_N, c[3] * c[4] * c[5], s[6] * s[7], K, (const cplx*)corner, c[3] * c[4] * c[5],
                        (const cplx*)side, s[6] * s[7], (cplx*)out, Levels{{c[3], st[0]}, {c[4], st[2]}, {c[5], st[4]}},
                        Levels{{s[6], st[1]}, {s[7], st[3]}}, accumulate != 0, s[0] * s[1] * s[2], K * s[6] * s[7],
                        c[3] * s[6] * c[4] * s[7] * c[5], S(stream));
  }
  for (int a = 0; a < 3; ++a)
    CARC_REQUIRE(c[3 + a] == s[a], CARC_ERR_DIMENSION_MISMATCH,
                 "tensor 0's index %d has dimension %lld, whereas tensor 1's index %d has dimension %lld", 3 + a,
                 (long long)c[3 + a], a, (long long)s[a]);
  const int64_t K = c[3] * c[4] * c[5], t345 = s[3] * s[4] * s[5];
  const auto st = row_major_strides({c[0], s[6], c[1], s[7], c[2], t345});
  return gemm_scatter(OP_N, OP_N, c[0] * c[1] * c[2], t345 * s[6] * s[7], K, (const cplx*)corner, K, (const cplx*)side,
                      t345 * s[6] * s[7], (cplx*)out, Levels{{c[0], st[0]}, {c[1], st[2]}, {c[2], st[4]}},
                      Levels{{t345, st[5]}, {s[6], st[1]}, {s[7], st[3]}}, accumulate != 0, 1, 0, 0, S(stream));
}

// reference tensors/_2d/dense.py:23-81, first half: the double-layer center of one site operator
int carc_double_layer_center(int direction, const void* center, const int64_t* n, const void* center_conj, const int64_t* m,
                             const void* operator_dev, void* E, void* stream) {
  CARC_REQUIRE(center && center_conj && n && m && E && direction >= 0 && direction < 4, CARC_ERR_VALUE,
               "double_layer_center: invalid argument");
  CARC_REQUIRE(n[4] == m[4], CARC_ERR_DIMENSION_MISMATCH,
               "tensor 1's index 4 has dimension %lld, whereas tensor 2's index 4 has dimension %lld", (long long)n[4],
               (long long)m[4]);
  using namespace carc;
  const int64_t d = n[4], rows = n[0] * n[1] * n[2] * n[3], cols = m[0] * m[1] * m[2] * m[3];
  Scratch scratch(S(stream));
  const cplx* v = (const cplx*)center;
  if (operator_dev) {
    cplx* vo;
    CARC_TRY(scratch.get(&vo, rows * d));
    CARC_TRY(plain_gemm(OP_N, OP_T, rows, d, d, v, d, (const cplx*)operator_dev, d, vo, false, S(stream)));
    v = vo;
  }
  const int i = direction, l = L4(i), r = R4(i), o = O4(i);
  const auto st = row_major_strides({n[i], m[i], n[l], m[l], n[r], m[r], n[o], m[o]});
  int role[4];
  role[i] = 0; role[l] = 2; role[r] = 4; role[o] = 6;
  Levels rowl, coll;
  for (int a = 0; a < 4; ++a) {
    rowl.push_back({n[a], st[role[a]]});
    coll.push_back({m[a], st[role[a] + 1]});
  }
  return gemm_scatter(OP_N, OP_T, rows, cols, d, v, d, (const cplx*)center_conj, d, (cplx*)E, rowl, coll, false, 1, 0, 0,
                      S(stream));
}

// ... second half: side (x) E summed over the legs facing the center; dims = the 8 extents carc_double_layer_center used
// (n_i, m_i, n_l, m_l, n_r, m_r, n_o, m_o: n the center's, m the conjugate's bonds)
int carc_absorb_center_into_side(const void* side, const int64_t* s, const void* E, const int64_t* dims, void* out,
                                 int accumulate, void* stream) {
  CARC_REQUIRE(side && s && E && dims && out, CARC_ERR_VALUE, "absorb_center_into_side: null argument");
  using namespace carc;
  CARC_REQUIRE(s[6] == dims[0] && s[7] == dims[1], CARC_ERR_DIMENSION_MISMATCH,
               "tensor 0's index 6 has dimension %lld, whereas tensor 1's facing leg has dimension %lld", (long long)s[6],
               (long long)dims[0]);
  const int64_t nl = dims[2], ml = dims[3], nr = dims[4], mr = dims[5], no = dims[6], mo = dims[7];
  const auto st = row_major_strides({s[0], nl, s[1], ml, s[2], s[3], nr, s[4], mr, s[5], no * mo});
  const int64_t K = s[6] * s[7], N = nl * ml * nr * mr * no * mo;
  return gemm_scatter(OP_N, OP_N, prod(s, 0, 6), N, K, (const cplx*)side, K, (const cplx*)E, N, (cplx*)out,
                      Levels{{s[0], st[0]}, {s[1], st[2]}, {s[2], st[4]}, {s[3], st[5]}, {s[4], st[7]}, {s[5], st[9]}},
                      Levels{{nl, st[1]}, {ml, st[3]}, {nr, st[6]}, {mr, st[8]}, {no * mo, st[10]}}, accumulate != 0, 1, 0, 0,
                      S(stream));
}

// reference tensors/_2d/dense.py:96-99
int carc_form_stage1(const void* corner, const int64_t* c, const void* side, const int64_t* s, void* out, int accumulate,
                     void* stream) {
  CARC_REQUIRE(corner && c && side && s && out, CARC_ERR_VALUE, "form_stage1: null argument");
  for (int a = 0; a < 3; ++a)
    CARC_REQUIRE(c[3 + a] == s[a], CARC_ERR_DIMENSION_MISMATCH,
                 "tensor 0's index %d has dimension %lld, whereas tensor 1's index %d has dimension %lld", 3 + a,
                 (long long)c[3 + a], a, (long long)s[a]);
  const int64_t K = c[3] * c[4] * c[5], N = carc::prod(s, 3, 8);
  return carc::plain_gemm(carc::OP_N, carc::OP_N, c[0] * c[1] * c[2], N, K, (const cplx*)corner, K, (const cplx*)side, N,
                          (cplx*)out, accumulate != 0, S(stream));
}

// reference tensors/_2d/dense.py:102-112 (+ the pre-joins of 130-131 when half = 0 / 1; half = -1: reference layout).
// slab_world > 1: only the slab [slow * rank / world, slow * (rank + 1) / world) of the slow factor of X is built and
// `out` holds that slab (multi-GPU).
int carc_form_stage2(const void* stage1_a, const int64_t* a, const void* stage1_b, const int64_t* b, int half, int slab_rank,
                     int slab_world, void* out, int accumulate, void* stream) {
  CARC_REQUIRE(stage1_a && a && stage1_b && b && half >= -1 && half <= 1, CARC_ERR_VALUE, "form_stage2: invalid argument");
  CARC_REQUIRE(a[0] == b[1], CARC_ERR_DIMENSION_MISMATCH,
               "tensor 0's index 0 has dimension %lld, whereas tensor 1's index 1 has dimension %lld", (long long)a[0],
               (long long)b[1]);
  using namespace carc;
  const int64_t K = a[0], N = b[2] * b[3], lda = a[1] * a[2] * a[3];
  const cplx* A = (const cplx*)stage1_a;
  const cplx* B = (const cplx*)stage1_b;
  if (half < 0) {
    CARC_REQUIRE(out, CARC_ERR_VALUE, "form_stage2: invalid argument (no output buffer)");
    CARC_REQUIRE(slab_world <= 1, CARC_ERR_VALUE, "form_stage2: X slabs exist only in the stage-3 layouts (half = 0 or 1)");
    const auto st = row_major_strides({a[1], a[2], b[2], a[3], b[3]});
    return gemm_scatter(OP_T, OP_N, lda, N, K, A, lda, B, N, (cplx*)out, Levels{{a[1], st[0]}, {a[2], st[1]}, {a[3], st[3]}},
                        Levels{{b[2], st[2]}, {b[3], st[4]}}, accumulate != 0, b[0], K * N, a[1] * a[2] * b[2] * a[3] * b[3],
                        S(stream));
  }
  const int64_t slow = half == 0 ? b[0] : a[1];
  int64_t lo = 0, hi = slow;
  if (slab_world > 1) {
    CARC_REQUIRE(slab_rank >= 0 && slab_rank < slab_world, CARC_ERR_VALUE, "form_stage2: rank %d outside world of %d", slab_rank,
                 slab_world);
    lo = slow * slab_rank / slab_world;
    hi = slow * (slab_rank + 1) / slab_world;
  }
  if (hi <= lo) return CARC_OK;   // more ranks than entries of the slow factor of X: this rank's slab (and `out`) is empty
  CARC_REQUIRE(out, CARC_ERR_VALUE, "form_stage2: invalid argument (no output buffer)");
  const int64_t n_b0 = half == 0 ? hi - lo : b[0], n_a1 = half == 0 ? a[1] : hi - lo;
  const int64_t rest = a[3] * b[3] * a[2] * b[2];
  const auto st = row_major_strides({a[3], b[3], a[2], b[2]});
  const int64_t a1_stride = half == 0 ? rest : n_b0 * rest, stride_c = half == 0 ? n_a1 * rest : rest;
  if (half == 0) B += lo * K * N;
  else A += lo * a[2] * a[3];
  return gemm_scatter(OP_T, OP_N, n_a1 * a[2] * a[3], N, K, A, lda, B, N, (cplx*)out,
                      Levels{{n_a1, a1_stride}, {a[2], st[2]}, {a[3], st[0]}}, Levels{{b[2], st[3]}, {b[3], st[1]}},
                      accumulate != 0, n_b0, K * N, stride_c, S(stream));
}

namespace carc {
namespace {
// matrix[(p r s'), (q S s)] += G[(p q), (r S)] * O[s', s]: the K = 1 "outer product with the site operator" that ends
// formMatrix (reference dense.py:185-194).  One thread per G element, consecutive threads along S, so that for every
// (s', s) a warp writes one contiguous run; zero entries of O (the identity of the normalization matrix has two) are
// skipped.  As a zgemm with N = d^2 = 4 this took 11.6 ms at D = 8 (131 072 CTAs of 128 x 64 tiles, 4 columns used) for
// 1 GB of output; here it is one read-modify-write pass.
__global__ void __launch_bounds__(256) form_matrix_scatter_kernel(const cplx* __restrict__ G, const cplx* __restrict__ O,
                                                                  int64_t P, int64_t Q, int64_t R, int64_t S, int d,
                                                                  cplx* __restrict__ out) {
  const int64_t total = P * Q * R * S;
  const int64_t n_in = Q * S * d;
  cplx o[16];
  for (int i = 0; i < d * d && i < 16; ++i) o[i] = O[i];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t S_ = e % S, r = (e / S) % R, q = (e / (S * R)) % Q, p = e / (S * R * Q);
    const cplx g = G[e];
    for (int s1 = 0; s1 < d; ++s1) {
      const int64_t rowoff = ((p * R + r) * d + s1) * n_in + (q * S + S_) * d;
      for (int s0 = 0; s0 < d; ++s0) {
        const cplx w = o[s1 * d + s0];
        if (w.x == 0.0 && w.y == 0.0) continue;
        cplx v = out[rowoff + s0];
        v.x += g.x * w.x - g.y * w.y;
        v.y += g.x * w.y + g.y * w.x;
        out[rowoff + s0] = v;
      }
    }
  }
}
}  // namespace
}  // namespace carc

// formMatrix of a stage-3 multiplier (reference dense.py:176-194): out[(P R s'), (Q S s)] (+)= sum_X A[X,(P Q)] B[X,(R S)]
// O[s', s] for the pre-joined halves A = [X, P, Q], B = [X, R, S] (P = D0* D1*, Q = D0 D1, R = D2* D3*, S = D2 D3) and a
// d x d site operator on the device (row-major [s'][s]; d <= 4).  accumulate == 0 zeroes `out` first.  X == 0 (an empty
// slab of the multi-GPU mode) contributes nothing.
int carc_stage3_form_matrix(const void* A, const void* B, int64_t X, int64_t P, int64_t Q, int64_t R, int64_t S,
                            const void* operator_dev, int d, void* out, int accumulate, void* stream) {
  CARC_REQUIRE(out && operator_dev && X >= 0 && P > 0 && Q > 0 && R > 0 && S > 0 && d >= 1 && d <= 4, CARC_ERR_VALUE,
               "stage3_form_matrix: invalid argument");
  using namespace carc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_out = P * R * d, n_in = Q * S * d;
  if (!accumulate) CARC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(cplx) * (size_t)(n_out * n_in), st));
  if (X == 0) return CARC_OK;
  CARC_REQUIRE(A && B, CARC_ERR_VALUE, "stage3_form_matrix: invalid argument");
  Scratch s(st);
  cplx* G;
  CARC_TRY(s.get(&G, P * Q * R * S));
  CARC_TRY(plain_gemm(OP_T, OP_N, P * Q, R * S, X, (const cplx*)A, P * Q, (const cplx*)B, R * S, G, false, s.stream));
  const int64_t total = P * Q * R * S;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  form_matrix_scatter_kernel<<<blocks, 256, 0, s.stream>>>(G, (const cplx*)operator_dev, P, Q, R, S, d, (cplx*)out);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// NDArrayData.normalizeAxis (data/__init__.py:263-301) for shape[axis] > 1: SVD of [(other axes), axis] = Q R -> U S V^H;
// normalized = Q (U V^H) with the axis back in place [same shape as t]; normalizer = conj(V S^-1 V^H), denormalizer =
// V S V^H (S^-1 skipped where S <= dont_recip_under); sqrt_svals != 0: the square-root variants and no tensor.
// Any output pointer may be NULL.
int carc_normalize_axis(const void* t, const int64_t* shape, int ndim, int axis, int sqrt_svals, double dont_recip_under,
                        void* normalized, void* normalizer, void* denormalizer, void* stream) {
  CARC_REQUIRE(t && shape && ndim >= 1 && ndim <= CARC_MAX_RANK && axis >= 0 && axis < ndim, CARC_ERR_VALUE,
               "normalize_axis: invalid argument");
  using namespace carc;
  const int64_t n = shape[axis], post = prod(shape, axis + 1, ndim), pre = prod(shape, 0, axis), m = pre * post;
  CARC_REQUIRE(n >= 1 && n <= 80, CARC_ERR_UNSUPPORTED, "normalize_axis: axis extent %lld outside 1..80", (long long)n);
  CARC_REQUIRE(m >= n, CARC_ERR_VALUE,
               "the total number of degrees of freedom in all other axes (%lld) are not enough to normalize axis (%d) with "
               "dimension (%lld)", (long long)m, axis, (long long)n);
  Scratch s(S(stream));
  cplx *M, *R, *Q, *tau, *U, *Sv, *Vh, *pieces;
  CARC_TRY(s.get(&M, m * n));
  CARC_TRY(s.get(&R, n * n));
  CARC_TRY(s.get(&Q, m * n));
  CARC_TRY(s.get(&tau, n));
  CARC_TRY(s.get(&U, n * n));
  CARC_TRY(s.get(&Sv, n));
  CARC_TRY(s.get(&Vh, n * n));
  CARC_TRY(s.get(&pieces, 5 * n * n));
  // [(pre), axis, (post)] -> [(pre post), axis]
  const int64_t shp[3] = {pre, n, post};
  const int32_t perm[3] = {0, 2, 1};
  CARC_TRY(permute((const cplx*)t, M, 3, shp, perm, 0, 0, s.stream));
  CARC_TRY(qr(M, m, (int)n, R, Q, tau, s.stream));
  CARC_TRY(svd_small(R, (int)n, U, Sv, Vh, s.stream));
  const int64_t nn = n * n;
  CARC_TRY(normalizer_matrices(U, Sv, Vh, (int)n, dont_recip_under, pieces, pieces + nn, pieces + 2 * nn, pieces + 3 * nn,
                               pieces + 4 * nn, s.stream));
  const size_t bytes = sizeof(cplx) * nn;
  if (normalizer)
    CARC_CHECK_CUDA(cudaMemcpyAsync(normalizer, pieces + (sqrt_svals ? 3 : 1) * nn, bytes, cudaMemcpyDeviceToDevice, s.stream));
  if (denormalizer)
    CARC_CHECK_CUDA(cudaMemcpyAsync(denormalizer, pieces + (sqrt_svals ? 4 : 2) * nn, bytes, cudaMemcpyDeviceToDevice, s.stream));
  if (normalized && !sqrt_svals) {
    // rows of Q are (pre, post); column j goes back to position `axis`
    GemmOut o;
    o.m_div = std::max<int64_t>(post, 1); o.m_s1 = n * post; o.m_s0 = 1;
    o.n_div = n; o.n_s1 = 0; o.n_s0 = post;
    CARC_TRY(zgemm(OP_N, OP_N, m, n, n, ONE, Q, n, pieces, n, ZERO, (cplx*)normalized, &o, nullptr, 1, 0, 0, 0, s.stream));
  }
  return CARC_OK;
}

// computeProductCompressor (compression.py:26-45) for operator bond 1: L [l, old, old, 1], R [old, old, 1, r] on the
// device, `initial` the random [old, new] draw (host RNG stays with the caller), `sweeps` ALS rounds
// (the reference: 4), `regularization` the relative diagonal shift of the normal equations (1e-10).  left_gram /
// right_gram: optional precomputed L^H L / R R^H over the outer legs, [(old old), (old old)], else NULL.
// compressor_out: [new, old].  Asynchronous on `stream`; no host round trip between rounds.
int carc_product_compressor(const void* L, int64_t l, const void* R, int64_t r, int64_t old_dim, int64_t new_dim,
                            const void* initial, int sweeps, double regularization, const void* left_gram,
                            const void* right_gram, void* compressor_out, void* stream) {
  CARC_REQUIRE(L && R && initial && compressor_out && old_dim >= 1 && new_dim >= 1 && new_dim <= old_dim && new_dim <= 80,
               CARC_ERR_VALUE, "product_compressor: invalid argument (old %lld, new %lld)", (long long)old_dim,
               (long long)new_dim);
  using namespace carc;
  cudaStream_t st = S(stream);
  Scratch s(st);
  const int64_t old = old_dim, nw = new_dim, o2 = old * old, m = old * nw;
  cplx* c;                                            // current compressor [old, new]
  CARC_TRY(s.get(&c, m));
  CARC_TRY(unitize_tall((const cplx*)initial, old, (int)nw, c, s));
  const int64_t shape_c[2] = {old, nw};
  const int32_t transpose[2] = {1, 0};
  if (nw == old || sweeps <= 0) return permute(c, (cplx*)compressor_out, 2, shape_c, transpose, 0, 0, st);

  // Gram factors over the outer legs, once per compression
  const cplx *LL0 = (const cplx*)left_gram, *RR0 = (const cplx*)right_gram;
  if (!LL0) {
    cplx* g;
    CARC_TRY(s.get(&g, o2 * o2));
    CARC_TRY(zgemm_hermitian(OP_C, OP_N, o2, l, (const cplx*)L, o2, (const cplx*)L, o2, g, st));   // sum_l conj(L) L
    LL0 = g;
  }
  if (!RR0) {
    cplx* g;
    CARC_TRY(s.get(&g, o2 * o2));
    CARC_TRY(zgemm_hermitian(OP_J, OP_T, o2, r, (const cplx*)R, r, (const cplx*)R, r, g, st));     // sum_r conj(R) R
    RR0 = g;
  }
  cplx *T, *LLg;
  CARC_TRY(s.get(&T, o2 * o2));
  CARC_TRY(s.get(&LLg, o2 * o2));
  CARC_TRY(plain_gemm(OP_N, OP_T, o2, o2, o2, LL0, o2, RR0, o2, T, false, st));                    // T[(i j), (k q)]
  const int64_t shape4[4] = {old, old, old, old};
  const int32_t p0213[4] = {0, 2, 1, 3};
  CARC_TRY(permute(LL0, LLg, 4, shape4, p0213, 0, 0, st));                                          // [(i i'), (j j')]

  cplx *Pc, *t1, *RRc, *t2, *W, *Wg, *G4, *gram, *U, *Ug, *rhs, *scratch_lu;
  int* piv;
  CARC_TRY(s.get(&Pc, o2));
  CARC_TRY(s.get(&t1, nw * old * o2));
  CARC_TRY(s.get(&RRc, nw * old * nw * old));
  CARC_TRY(s.get(&t2, nw * old * nw * old));
  CARC_TRY(s.get(&W, nw * old * nw * old));
  CARC_TRY(s.get(&Wg, o2 * nw * nw));
  CARC_TRY(s.get(&G4, o2 * nw * nw));
  CARC_TRY(s.get(&gram, m * m));
  CARC_TRY(s.get(&U, o2 * nw * old));
  CARC_TRY(s.get(&Ug, o2 * nw * old));
  CARC_TRY(s.get(&rhs, m));
  CARC_TRY(s.get(&scratch_lu, (int64_t)(lu_scratch_bytes() + 64 + 15) / 16 + (m * (int64_t)sizeof(int) + 15) / 16 + 2));
  piv = reinterpret_cast<int*>(scratch_lu + (lu_scratch_bytes() + 64 + 15) / 16);
  int* singular_dev = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch_lu) + lu_scratch_bytes());

  for (int round = 0; round < sweeps; ++round) {
    // Pc[j, q] = sum_n conj(c)[j, n] c[q, n]
    CARC_TRY(plain_gemm(OP_J, OP_T, old, old, nw, c, nw, c, nw, Pc, false, st));
    // RRc[n, q, n', q'] = sum_{k k'} c[k, n] RR0[k, q, k', q'] conj(c)[k', n']
    CARC_TRY(absorb_matrix(RR0, 1, old, old * o2, OP_T, c, nw, nw, t1, st));                        // axis 0 by c^T
    CARC_TRY(absorb_matrix(t1, nw * old, old, old, OP_C, c, nw, nw, RRc, st));                      // axis 2 by c^H
    // W[n, j, n', j'] = sum_{q q'} conj(Pc)[j, q] RRc[n, q, n', q'] Pc[j', q']
    CARC_TRY(absorb_matrix(RRc, nw, old, nw * old, OP_J, Pc, old, old, t2, st));                    // axis 1 by conj(Pc)
    CARC_TRY(absorb_matrix(t2, nw * old * nw, old, 1, OP_N, Pc, old, old, W, st));                  // axis 3 by Pc
    const int64_t shape_w[4] = {nw, old, nw, old};
    const int32_t p1302[4] = {1, 3, 0, 2};
    CARC_TRY(permute(W, Wg, 4, shape_w, p1302, 0, 0, st));                                          // [(j j'), (n n')]
    CARC_TRY(plain_gemm(OP_N, OP_N, o2, nw * nw, o2, LLg, o2, Wg, nw * nw, G4, false, st));        // [(i i'), (n n')]
    const int64_t shape_g[4] = {old, old, nw, nw};
    CARC_TRY(permute(G4, gram, 4, shape_g, p0213, 0, 0, st));                                       // [(i n), (i' n')]
    // rhs[(i n)] = sum_{j q k} conj(Pc)[j, q] c[k, n] T[i, j, k, q]
    CARC_TRY(absorb_matrix(T, o2, old, old, OP_T, c, nw, nw, U, st));                               // [i, j, n, q]
    const int64_t shape_u[4] = {old, old, nw, old};
    CARC_TRY(permute(U, Ug, 4, shape_u, p0213, 0, 0, st));                                          // [(i n), (j q)]
    CARC_TRY(plain_gemm(OP_N, OP_J, m, 1, o2, Ug, o2, Pc, 1, rhs, false, st));
    // x = (G + eps mean(diag G) I)^-1 rhs
    diagonal_shift_kernel<<<1, 256, 0, st>>>(gram, (int)m, regularization);
    CARC_CHECK_CUDA(cudaGetLastError());
    if (m <= SMALL_LU_MAX) {
      CARC_TRY(small_lu_solve(gram, (int)m, rhs, st));
    } else {
      CARC_TRY(lu_factor(gram, (int)m, piv, singular_dev, scratch_lu, st));
      CARC_TRY(lu_solve(gram, (int)m, piv, rhs, st));
    }
    CARC_TRY(unitize_tall(rhs, old, (int)nw, c, s));
  }
  return permute(c, (cplx*)compressor_out, 2, shape_c, transpose, 0, 0, st);
}

}  // extern "C"
