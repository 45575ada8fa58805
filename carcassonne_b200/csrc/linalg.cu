// Small dense factorisations on device: Householder QR of tall-skinny matrices, one-sided Jacobi SVD of the
// n x n triangular factor, and the polar / normaliser matrices NDArrayData.normalizeAxis derives from them.
//
// Reference call sites: scipy.linalg.svd (zgesdd) in NDArrayData.normalizeAxis / svd / unitize
// (data/__init__.py:263-301, 344-346, utils.py:879-881) and scipy.linalg.qr (zgeqrf) in newEnlargener
// (data/__init__.py:43-50).  The matrices are (D^3 d) x D or (chi D) x chi -- at most a few MB -- so each
// factorisation is one persistent CTA working out of L2; the flops that matter (Q * U_R) go through the DMMA
// GEMM.  Householder conventions follow LAPACK zlarfg / zgeqr2 / zung2r so that Q agrees with SciPy's Q to
// rounding (it is not gauge-free: newEnlargener hands Q itself to the caller).
#include "carc_internal.h"
#include "common.cuh"

namespace carc {

namespace {

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

__device__ __forceinline__ cplx warp_csum(cplx v) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  return v;
}

// A [m, n] row-major, m >= n.  On exit: R [n, n] upper triangular, Q [m, n] with orthonormal columns
// (A is overwritten by the reflectors).
__global__ void __launch_bounds__(1024) qr_kernel(cplx* __restrict__ A, int m, int n, cplx* __restrict__ R,
                                                  cplx* __restrict__ Q, cplx* __restrict__ tau) {
  __shared__ double sh[32];
  __shared__ cplx s_tau, s_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int k = 0; k < n; ++k) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < m; i += blockDim.x) {
      const cplx a = A[(int64_t)i * n + k];
      part += a.x * a.x + a.y * a.y;
    }
    const double xnorm2 = block_sum(part, sh);
    if (tid == 0) {
      const cplx alpha = A[(int64_t)k * n + k];
      if (xnorm2 == 0.0 && alpha.y == 0.0) {
        s_tau = make_double2(0.0, 0.0);
        s_scale = make_double2(0.0, 0.0);
      } else {
        double beta = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
        if (alpha.x > 0.0 || (alpha.x == 0.0 && !signbit(alpha.x))) beta = -beta;
        s_tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        const cplx dlt = make_double2(alpha.x - beta, alpha.y);
        const double den = dlt.x * dlt.x + dlt.y * dlt.y;
        s_scale = make_double2(dlt.x / den, -dlt.y / den);
        A[(int64_t)k * n + k] = make_double2(beta, 0.0);
      }
      tau[k] = s_tau;
    }
    __syncthreads();
    const cplx t = s_tau, sc = s_scale;
    if (t.x != 0.0 || t.y != 0.0) {
      for (int i = k + 1 + tid; i < m; i += blockDim.x) A[(int64_t)i * n + k] = cmul(A[(int64_t)i * n + k], sc);
      __syncthreads();
      // trailing update with H^H = I - conj(tau) v v^H, one column per warp
      const cplx ct = make_double2(t.x, -t.y);
      for (int j = k + 1 + warp; j < n; j += nw) {
        cplx w = make_double2(0.0, 0.0);
        for (int i = k + 1 + lane; i < m; i += 32) w = cadd(w, cmulc(A[(int64_t)i * n + k], A[(int64_t)i * n + j]));
        w = warp_csum(w);
        w = cadd(w, A[(int64_t)k * n + j]);
        const cplx f = cmul(ct, w);
        if (lane == 0) A[(int64_t)k * n + j] = csub(A[(int64_t)k * n + j], f);
        for (int i = k + 1 + lane; i < m; i += 32)
          A[(int64_t)i * n + j] = csub(A[(int64_t)i * n + j], cmul(f, A[(int64_t)i * n + k]));
      }
    }
    __syncthreads();
  }
  // R
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    R[e] = j >= i ? A[(int64_t)i * n + j] : make_double2(0.0, 0.0);
  }
  // Q = H(0) H(1) ... H(n-1) applied to the first n columns of the identity
  for (int64_t e = tid; e < (int64_t)m * n; e += blockDim.x) {
    const int64_t i = e / n;
    const int j = (int)(e % n);
    Q[e] = (i == j) ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
  }
  __syncthreads();
  for (int k = n - 1; k >= 0; --k) {
    const cplx t = tau[k];
    if (t.x != 0.0 || t.y != 0.0) {
      for (int j = k + warp; j < n; j += nw) {
        cplx w = make_double2(0.0, 0.0);
        for (int i = k + 1 + lane; i < m; i += 32) w = cadd(w, cmulc(A[(int64_t)i * n + k], Q[(int64_t)i * n + j]));
        w = warp_csum(w);
        w = cadd(w, Q[(int64_t)k * n + j]);
        const cplx f = cmul(t, w);
        if (lane == 0) Q[(int64_t)k * n + j] = csub(Q[(int64_t)k * n + j], f);
        for (int i = k + 1 + lane; i < m; i += 32)
          Q[(int64_t)i * n + j] = csub(Q[(int64_t)i * n + j], cmul(f, A[(int64_t)i * n + k]));
      }
    }
    __syncthreads();
  }
}

// One-sided Jacobi SVD of an n x n matrix (n <= 128), one CTA.  W (columns of R V) and V live in shared memory,
// stored column-major.  Pairs of a round-robin round are rotated concurrently, one pair per warp.
// Outputs (row-major): U [n,n], S [n] (descending, as (s, 0) complex), Vh [n,n].
// Columns whose singular value is below n*eps*s_max are completed to an orthonormal basis (what LAPACK's U
// provides for a rank-deficient input).
__global__ void __launch_bounds__(256) svd_small_kernel(const cplx* __restrict__ Rin, int n, cplx* __restrict__ U,
                                                        cplx* __restrict__ S, cplx* __restrict__ Vh) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* W = reinterpret_cast<cplx*>(smem_raw);   // W[j*n + i] = column j
  cplx* V = W + n * n;
  double* sv = reinterpret_cast<double*>(V + n * n);  // n
  int* order = reinterpret_cast<int*>(sv + n);        // n
  __shared__ int rotated;
  __shared__ double sh[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    W[j * n + i] = Rin[e];
    V[j * n + i] = (i == j) ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const int np = (n + 1) & ~1;  // players (one dummy if n is odd)
  const double eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < np - 1; ++round) {
      for (int pi = warp; pi < np / 2; pi += nw) {
        int p, q;
        if (pi == 0) {
          p = round % (np - 1);
          q = np - 1;
        } else {
          p = (pi + round) % (np - 1);
          q = (np - 1 - pi + round) % (np - 1);
        }
        if (p >= n || q >= n) continue;
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        cplx* wp = W + p * n;
        cplx* wq = W + q * n;
        double al = 0.0, be = 0.0;
        cplx ga = make_double2(0.0, 0.0);
        for (int i = lane; i < n; i += 32) {
          const cplx a = wp[i], b = wq[i];
          al += a.x * a.x + a.y * a.y;
          be += b.x * b.x + b.y * b.y;
          ga = cadd(ga, cmulc(a, b));
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_csum(ga);
        const double ag = sqrt(ga.x * ga.x + ga.y * ga.y);
        if (ag <= eps * sqrt(al * be) || ag == 0.0) continue;
        if (lane == 0) rotated = 1;
        const double zeta = (be - al) / (2.0 * ag);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        const cplx ph = make_double2(ga.x / ag, ga.y / ag);       // e^{i phi}
        const cplx sp = make_double2(s * ph.x, s * ph.y);         // s e^{i phi}
        const cplx sm = make_double2(s * ph.x, -s * ph.y);        // s e^{-i phi}
        for (int i = lane; i < n; i += 32) {
          const cplx a = wp[i], b = wq[i];
          wp[i] = csub(cscale(a, c), cmul(sm, b));
          wq[i] = cadd(cmul(sp, a), cscale(b, c));
        }
        cplx* vp = V + p * n;
        cplx* vq = V + q * n;
        for (int i = lane; i < n; i += 32) {
          const cplx a = vp[i], b = vq[i];
          vp[i] = csub(cscale(a, c), cmul(sm, b));
          vq[i] = cadd(cmul(sp, a), cscale(b, c));
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  // singular values and ordering
  for (int j = warp; j < n; j += nw) {
    double a = 0.0;
    for (int i = lane; i < n; i += 32) {
      const cplx w = W[j * n + i];
      a += w.x * w.x + w.y * w.y;
    }
    a = warp_sum(a);
    if (lane == 0) sv[j] = sqrt(a);
  }
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < n; ++j) order[j] = j;
    for (int a = 0; a < n; ++a) {   // selection sort, descending
      int best = a;
      for (int b2 = a + 1; b2 < n; ++b2)
        if (sv[order[b2]] > sv[order[best]]) best = b2;
      const int tmp = order[a]; order[a] = order[best]; order[best] = tmp;
    }
  }
  __syncthreads();
  const double smax = sv[order[0]];
  const double cut = smax * eps * n;
  // normalise the well-defined left vectors in place
  int nrank = 0;
  for (int j = 0; j < n; ++j)
    if (sv[order[j]] > cut) nrank = j + 1;   // descending, so ranks form a prefix
  for (int jj = warp; jj < nrank; jj += nw) {
    const int j = order[jj];
    const double inv = 1.0 / sv[j];
    for (int i = lane; i < n; i += 32) W[j * n + i] = cscale(W[j * n + i], inv);
  }
  __syncthreads();
  // complete the null directions: Gram-Schmidt of unit vectors against what is there (twice for stability)
  int cand = 0;
  for (int jj = nrank; jj < n; ++jj) {
    const int j = order[jj];
    for (;; ++cand) {
      // start from e_cand
      for (int i = tid; i < n; i += blockDim.x) W[j * n + i] = (i == cand) ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
      __syncthreads();
      for (int pass = 0; pass < 2; ++pass) {
        for (int kk = 0; kk < jj; ++kk) {
          const int k = order[kk];
          cplx dpart = make_double2(0.0, 0.0);
          for (int i = tid; i < n; i += blockDim.x) dpart = cadd(dpart, cmulc(W[k * n + i], W[j * n + i]));
          cplx dtot;
          dtot.x = block_sum(dpart.x, sh);
          dtot.y = block_sum(dpart.y, sh);
          for (int i = tid; i < n; i += blockDim.x) W[j * n + i] = csub(W[j * n + i], cmul(dtot, W[k * n + i]));
          __syncthreads();
        }
      }
      double npart = 0.0;
      for (int i = tid; i < n; i += blockDim.x) {
        const cplx w = W[j * n + i];
        npart += w.x * w.x + w.y * w.y;
      }
      const double nrm = sqrt(block_sum(npart, sh));
      if (nrm > 0.5 || cand >= n - 1) {
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        for (int i = tid; i < n; i += blockDim.x) W[j * n + i] = cscale(W[j * n + i], inv);
        __syncthreads();
        ++cand;
        break;
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, jj = e % n;
    const int j = order[jj];
    U[i * n + jj] = W[j * n + i];
    const cplx v = V[j * n + i];           // V[i][j]
    Vh[jj * n + i] = make_double2(v.x, -v.y);
  }
  for (int jj = tid; jj < n; jj += blockDim.x) S[jj] = make_double2(sv[order[jj]], 0.0);
}

// Matrices NDArrayData.normalizeAxis builds from the SVD (data/__init__.py:280-301); all n x n, row-major.
//   polar      = U Vh                       (times Q gives the isometric tensor)
//   normalizer = Vh^T conj(SI Vh)    = conj(V SI V^H)
//   denorm     = Vh^H (S Vh)         = V S V^H
//   sqrt variants: nrm_sqrt = conj(sqrt(SI) Vh), den_sqrt = sqrt(S) Vh
__global__ void __launch_bounds__(256) normalizer_kernel(const cplx* __restrict__ U, const cplx* __restrict__ S,
                                                         const cplx* __restrict__ Vh, int n, double dont_recip_under,
                                                         cplx* __restrict__ polar, cplx* __restrict__ nrm,
                                                         cplx* __restrict__ den, cplx* __restrict__ nrm_sqrt,
                                                         cplx* __restrict__ den_sqrt) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int i = e / n, j = e % n;
    cplx p = make_double2(0.0, 0.0), a = p, b = p;
    for (int k = 0; k < n; ++k) {
      const double s = S[k].x;
      double si = s;
      if (dont_recip_under != 0.0) {
        if (fabs(s) > dont_recip_under) si = 1.0 / s;
      } else {
        si = 1.0 / s;
      }
      const cplx vki = Vh[k * n + i], vkj = Vh[k * n + j];
      p = cadd(p, cmul(U[i * n + k], vkj));
      // normalizer[i][j] = sum_k Vh[k][i] * conj(si * Vh[k][j])
      a = cadd(a, cscale(cmul(vki, make_double2(vkj.x, -vkj.y)), si));
      // denorm[i][j] = sum_k conj(Vh[k][i]) * s * Vh[k][j]
      b = cadd(b, cscale(cmulc(vki, vkj), s));
    }
    polar[e] = p;
    nrm[e] = a;
    den[e] = b;
    const double s = S[i].x;
    double si = s;
    if (dont_recip_under != 0.0) {
      if (fabs(s) > dont_recip_under) si = 1.0 / s;
    } else {
      si = 1.0 / s;
    }
    const cplx v = Vh[e];
    nrm_sqrt[e] = make_double2(sqrt(si) * v.x, -sqrt(si) * v.y);
    den_sqrt[e] = cscale(v, sqrt(s));
  }
}

}  // namespace

int qr(cplx* A, int64_t m, int n, cplx* R, cplx* Q, cplx* tau, cudaStream_t stream) {
  CARC_REQUIRE(m >= n && n >= 1, CARC_ERR_VALUE, "qr: need m >= n >= 1 (got %lld x %d)", (long long)m, n);
  CARC_REQUIRE(m < (1ll << 31), CARC_ERR_VALUE, "qr: too many rows");
  qr_kernel<<<1, 1024, 0, stream>>>(A, (int)m, n, R, Q, tau);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int svd_small(const cplx* R, int n, cplx* U, cplx* S, cplx* Vh, cudaStream_t stream) {
  CARC_REQUIRE(n >= 1 && n <= 80, CARC_ERR_UNSUPPORTED, "svd_small: n = %d outside 1..80", n);
  const size_t smem = sizeof(cplx) * 2 * n * n + sizeof(double) * n + sizeof(int) * n + 16;
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(cplx) * 2 * 80 * 80 + sizeof(double) * 80 + sizeof(int) * 80 + 16)));
    configured[dev] = true;
  }
  svd_small_kernel<<<1, 256, smem, stream>>>(R, n, U, S, Vh);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int normalizer_matrices(const cplx* U, const cplx* S, const cplx* Vh, int n, double dont_recip_under, cplx* polar,
                        cplx* nrm, cplx* den, cplx* nrm_sqrt, cplx* den_sqrt, cudaStream_t stream) {
  normalizer_kernel<<<(n * n + 255) / 256, 256, 0, stream>>>(U, S, Vh, n, dont_recip_under, polar, nrm, den, nrm_sqrt,
                                                             den_sqrt);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

}  // namespace carc
