// extern "C" surface of libcarc_b200.so (declared in include/carc_b200.h).
#include <vector>

#include "../../include/carc_b200.h"
#include "carc_internal.h"
#include "common.cuh"

using carc::cplx;

struct carc_operator {
  int P, Q, R, S, d;
  std::vector<carc::Stage3Term> terms;
  carc::Stage3Term* terms_dev = nullptr;
  cplx* workspace = nullptr;
  int64_t workspace_elems = 0;
  int force_path = 0;
  bool finalized = false;
};

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline cplx C2(const double z[2]) { return make_double2(z[0], z[1]); }

extern "C" {

int carc_version(void) { return 100; }
const char* carc_last_error(void) { return carc::get_error(); }

int carc_dmma_peak(int iters, double* tflops_out, void* stream) { return carc::dmma_peak(iters, tflops_out, S(stream)); }

int carc_malloc(void** ptr, size_t bytes) {
  CARC_CHECK_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
  return CARC_OK;
}
int carc_free(void* ptr) {
  CARC_CHECK_CUDA(cudaFree(ptr));
  return CARC_OK;
}
int carc_malloc_host(void** ptr, size_t bytes) {
  CARC_CHECK_CUDA(cudaMallocHost(ptr, bytes ? bytes : 16));
  return CARC_OK;
}
int carc_free_host(void* ptr) {
  CARC_CHECK_CUDA(cudaFreeHost(ptr));
  return CARC_OK;
}
int carc_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  CARC_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S(stream)));
  return CARC_OK;
}
int carc_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  CARC_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(stream)));
  return CARC_OK;
}
int carc_stream_synchronize(void* stream) {
  CARC_CHECK_CUDA(cudaStreamSynchronize(S(stream)));
  return CARC_OK;
}

int carc_permute(const void* src, void* dst, int ndim, const int64_t* shape, const int32_t* perm, int conj,
                 int accumulate, void* stream) {
  return carc::permute((const cplx*)src, (cplx*)dst, ndim, shape, perm, conj, accumulate, S(stream));
}

int carc_axpby(int64_t n, const double alpha[2], const void* x, const double beta[2], void* y, int conj_x,
               void* stream) {
  return carc::axpby(n, C2(alpha), (const cplx*)x, C2(beta), (cplx*)y, conj_x, S(stream));
}
int carc_mul(int64_t n, const void* x, void* y, void* stream) {
  return carc::mul_inplace(n, (const cplx*)x, (cplx*)y, S(stream));
}
int carc_dotc(int64_t n, const void* x, const void* y, void* out, void* stream) {
  return carc::reduce(0, n, (const cplx*)x, (const cplx*)y, (double2*)out, S(stream));
}
int carc_sumsq(int64_t n, const void* x, void* out, void* stream) {
  return carc::reduce(1, n, (const cplx*)x, nullptr, (double2*)out, S(stream));
}
int carc_count_nonfinite(int64_t n, const void* x, void* out, void* stream) {
  return carc::reduce(2, n, (const cplx*)x, nullptr, (double2*)out, S(stream));
}

int carc_zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, const double alpha[2], const void* A, int64_t lda,
               const void* B, int64_t ldb, const double beta[2], void* C, const int64_t* out_map, const int64_t* k_map,
               int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC, void* stream) {
  carc::GemmOut om;
  carc::GemmKMap km;
  if (out_map) {
    om.m_div = out_map[0]; om.m_s1 = out_map[1]; om.m_s0 = out_map[2];
    om.n_div = out_map[3]; om.n_s1 = out_map[4]; om.n_s0 = out_map[5];
  }
  if (k_map) {
    km.a_kdiv = k_map[0]; km.a_ks1 = k_map[1]; km.b_kdiv = k_map[2]; km.b_ks1 = k_map[3];
  }
  return carc::zgemm(opA, opB, M, N, K, C2(alpha), (const cplx*)A, lda, (const cplx*)B, ldb, C2(beta), (cplx*)C,
                     out_map ? &om : nullptr, k_map ? &km : nullptr, batch, strideA, strideB, strideC, S(stream));
}

// ---------------------------------------------------------------------------------------------------
int carc_operator_create(carc_operator** op, int P, int Q, int R, int Sd, int d) {
  CARC_REQUIRE(op != nullptr, CARC_ERR_VALUE, "operator_create: null handle pointer");
  CARC_REQUIRE(P > 0 && Q > 0 && R > 0 && Sd > 0 && d > 0 && d <= 4, CARC_ERR_VALUE,
               "operator_create: invalid dimensions P=%d Q=%d R=%d S=%d d=%d", P, Q, R, Sd, d);
  carc_operator* o = new carc_operator();
  o->P = P; o->Q = Q; o->R = R; o->S = Sd; o->d = d;
  *op = o;
  return CARC_OK;
}

int carc_operator_add_term(carc_operator* op, const void* A, const void* B, int64_t X, const double* O_host) {
  CARC_REQUIRE(op && !op->finalized, CARC_ERR_VALUE, "operator_add_term: operator missing or already finalized");
  CARC_REQUIRE(A && B && X >= 0, CARC_ERR_VALUE, "operator_add_term: invalid term");
  carc::Stage3Term t;
  t.A = (const cplx*)A;
  t.B = (const cplx*)B;
  t.X = X;
  t.has_op = O_host != nullptr;
  for (int i = 0; i < 16; ++i) t.op[i] = make_double2(0.0, 0.0);
  if (O_host)
    for (int i = 0; i < op->d * op->d; ++i) t.op[i] = make_double2(O_host[2 * i], O_host[2 * i + 1]);
  else
    for (int i = 0; i < op->d; ++i) t.op[i * op->d + i] = make_double2(1.0, 0.0);
  op->terms.push_back(t);
  return CARC_OK;
}

int carc_operator_finalize(carc_operator* op) {
  CARC_REQUIRE(op && !op->finalized, CARC_ERR_VALUE, "operator_finalize: operator missing or already finalized");
  const int nt = (int)op->terms.size();
  int64_t Xmax = 1;
  for (auto& t : op->terms) Xmax = t.X > Xmax ? t.X : Xmax;
  if (nt > 0) {
    CARC_CHECK_CUDA(cudaMalloc(&op->terms_dev, sizeof(carc::Stage3Term) * nt));
    CARC_CHECK_CUDA(cudaMemcpy(op->terms_dev, op->terms.data(), sizeof(carc::Stage3Term) * nt, cudaMemcpyHostToDevice));
  }
  op->workspace_elems = carc::stage3_workspace_elems(nt, op->P, op->Q, op->R, op->S, op->d, Xmax);
  CARC_CHECK_CUDA(cudaMalloc(&op->workspace, sizeof(cplx) * (op->workspace_elems > 0 ? op->workspace_elems : 1)));
  op->finalized = true;
  return CARC_OK;
}

int carc_operator_set_path(carc_operator* op, int force_path) {
  CARC_REQUIRE(op && force_path >= 0 && force_path <= 2, CARC_ERR_VALUE, "operator_set_path: invalid argument");
  op->force_path = force_path;
  return CARC_OK;
}

int carc_operator_num_terms(const carc_operator* op) { return op ? (int)op->terms.size() : -1; }

int64_t carc_operator_cost_of_multiply(const carc_operator* op) {
  if (!op) return -1;
  int64_t cost = 0;
  const int64_t P = op->P, Q = op->Q, R = op->R, Sd = op->S, d = op->d;
  for (auto& t : op->terms) {
    if (t.has_op) cost += d * d * Q * Sd;
    cost += t.X * P * (Sd * d) * Q;
    cost += R * (P * d) * (Sd * t.X);
  }
  return cost;
}

int carc_operator_apply(carc_operator* op, const void* v, void* out, void* stream) {
  CARC_REQUIRE(op && op->finalized, CARC_ERR_VALUE, "operator_apply: operator not finalized");
  return carc::stage3_apply(op->terms.data(), op->terms_dev, (int)op->terms.size(), op->P, op->Q, op->R, op->S, op->d,
                            (const cplx*)v, (cplx*)out, op->workspace, op->workspace_elems, op->force_path, S(stream));
}

int carc_operator_destroy(carc_operator* op) {
  if (!op) return CARC_OK;
  if (op->terms_dev) cudaFree(op->terms_dev);
  if (op->workspace) cudaFree(op->workspace);
  delete op;
  return CARC_OK;
}

int carc_stage3_matvec_host(int nterms, const void* const* A_host, const void* const* B_host, const int64_t* X,
                            const double* const* O_host, int P, int Q, int R, int Sd, int d, const void* v_host,
                            void* out_host, void* stream) {
  CARC_REQUIRE(nterms >= 0, CARC_ERR_VALUE, "stage3_matvec_host: negative term count");
  cudaStream_t st = S(stream);
  std::vector<void*> bufs;
  auto cleanup = [&]() {
    for (void* b : bufs) cudaFree(b);
  };
  carc_operator* op = nullptr;
  int rc = carc_operator_create(&op, P, Q, R, Sd, d);
  if (rc) return rc;
  for (int t = 0; t < nterms && !rc; ++t) {
    void *a = nullptr, *b = nullptr;
    size_t abytes = sizeof(cplx) * (size_t)X[t] * P * Q, bbytes = sizeof(cplx) * (size_t)X[t] * R * Sd;
    if (cudaMalloc(&a, abytes ? abytes : 16) != cudaSuccess || cudaMalloc(&b, bbytes ? bbytes : 16) != cudaSuccess) {
      carc::set_error("stage3_matvec_host: out of device memory");
      rc = CARC_ERR_CUDA;
      if (a) bufs.push_back(a);
      break;
    }
    bufs.push_back(a);
    bufs.push_back(b);
    cudaMemcpyAsync(a, A_host[t], abytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(b, B_host[t], bbytes, cudaMemcpyHostToDevice, st);
    rc = carc_operator_add_term(op, a, b, X[t], O_host ? O_host[t] : nullptr);
  }
  void *v = nullptr, *out = nullptr;
  if (!rc) {
    size_t vbytes = sizeof(cplx) * (size_t)Q * Sd * d, obytes = sizeof(cplx) * (size_t)P * R * d;
    if (cudaMalloc(&v, vbytes) != cudaSuccess || cudaMalloc(&out, obytes) != cudaSuccess) {
      carc::set_error("stage3_matvec_host: out of device memory");
      rc = CARC_ERR_CUDA;
    }
    if (v) bufs.push_back(v);
    if (out) bufs.push_back(out);
    if (!rc) {
      cudaMemcpyAsync(v, v_host, vbytes, cudaMemcpyHostToDevice, st);
      rc = carc_operator_finalize(op);
      if (!rc) rc = carc_operator_apply(op, v, out, stream);
      if (!rc) {
        cudaMemcpyAsync(out_host, out, obytes, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) {
          carc::set_error("stage3_matvec_host: %s", cudaGetErrorString(cudaGetLastError()));
          rc = CARC_ERR_CUDA;
        }
      }
    }
  }
  cudaStreamSynchronize(st);
  carc_operator_destroy(op);
  cleanup();
  return rc;
}

int carc_qr(void* A, int64_t m, int n, void* R, void* Q, void* tau, void* stream) {
  return carc::qr((cplx*)A, m, n, (cplx*)R, (cplx*)Q, (cplx*)tau, S(stream));
}
int carc_svd_small(const void* R, int n, void* U, void* Sv, void* Vh, void* stream) {
  return carc::svd_small((const cplx*)R, n, (cplx*)U, (cplx*)Sv, (cplx*)Vh, S(stream));
}
int carc_normalizer_matrices(const void* U, const void* Sv, const void* Vh, int n, double dont_recip_under,
                             void* polar, void* normalizer, void* denormalizer, void* normalizer_sqrt,
                             void* denormalizer_sqrt, void* stream) {
  return carc::normalizer_matrices((const cplx*)U, (const cplx*)Sv, (const cplx*)Vh, n, dont_recip_under,
                                   (cplx*)polar, (cplx*)normalizer, (cplx*)denormalizer, (cplx*)normalizer_sqrt,
                                   (cplx*)denormalizer_sqrt, S(stream));
}

}  // extern "C"
