// extern "C" surface of libcarc_b200.so (declared in include/carc_b200.h).
#include <vector>

#include "../../include/carc_b200.h"
#include "carc_internal.h"
#include "common.cuh"

using carc::cplx;

struct carc_operator {
  int kind = 0;                 // 0: stage-3 term list, 1: dense matrix
  const cplx* matrix = nullptr; // kind 1: [n, n] row-major (not owned)
  int64_t n = 0;
  int P, Q, R, S, d;
  std::vector<carc::Stage3Term> terms;
  carc::Stage3Plan* plan = nullptr;
  cplx* workspace = nullptr;    // partial sums: from the stream-ordered pool on the first apply (see stage3_plan_upload)
  cudaStream_t workspace_stream = nullptr;
  int64_t workspace_elems = 0;
  int force_path = 0;
  bool finalized = false;
  carc::Comm* comm = nullptr;   // X-slab sharding: sum the output over ranks inside apply
};

struct carc_comm {
  carc::Comm* impl = nullptr;
};

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline cplx C2(const double z[2]) { return make_double2(z[0], z[1]); }

// cudaMallocAsync returns freed memory to the OS at the next synchronisation unless the pool is told to keep it;
// every solver call allocates its workspace from the pool, so keep it.
static void keep_pool_memory() {
  static bool done[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 16 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done[dev] = true;
}

namespace carc {
int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;   // host-side planning without a device (CPU test suite): a B200's count
  }
  if (dev < 16 && cached[dev]) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;
  }
  if (dev < 16) cached[dev] = n;
  return n;
}
}  // namespace carc

extern "C" {

int carc_version(void) { return 200; }
const char* carc_last_error(void) { return carc::get_error(); }

int carc_dmma_peak(int iters, double* tflops_out, void* stream) { return carc::dmma_peak(iters, tflops_out, S(stream)); }
int carc_dmma_rate(int iters, int warps_per_sm, int chains, double* tflops_out, void* stream) {
  return carc::dmma_rate(iters, warps_per_sm, chains, tflops_out, S(stream));
}

int carc_fp64_mix_rate(int iters, int warps_per_sm, int ndmma, int nfma, double* tflops_dmma, double* tflops_fma,
                       void* stream) {
  return carc::fp64_mix_rate(iters, warps_per_sm, ndmma, nfma, tflops_dmma, tflops_fma, S(stream));
}

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
int carc_mode_product(const void* matrix, const void* x, void* out, int64_t rows, int64_t k, int64_t pre, int64_t post,
                      void* stream) {
  return carc::mode_product((const cplx*)matrix, (const cplx*)x, (cplx*)out, rows, k, pre, post, S(stream));
}
int carc_mul(int64_t n, const void* x, void* y, void* stream) {
  return carc::mul_inplace(n, (const cplx*)x, (cplx*)y, S(stream));
}
int carc_dotc(int64_t n, const void* x, const void* y, void* out, void* stream) {
  return carc::reduce(0, n, (const cplx*)x, (const cplx*)y, (double2*)out, S(stream));
}
int carc_dotu(int64_t n, const void* x, const void* y, void* out, void* stream) {
  return carc::reduce(3, n, (const cplx*)x, (const cplx*)y, (double2*)out, S(stream));
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
  keep_pool_memory();
  return carc::zgemm(opA, opB, M, N, K, C2(alpha), (const cplx*)A, lda, (const cplx*)B, ldb, C2(beta), (cplx*)C,
                     out_map ? &om : nullptr, k_map ? &km : nullptr, batch, strideA, strideB, strideC, S(stream));
}

int carc_zgemm_hermitian(int opA, int opB, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                         void* C, void* stream) {
  keep_pool_memory();
  return carc::zgemm_hermitian(opA, opB, N, K, (const cplx*)A, lda, (const cplx*)B, ldb, (cplx*)C, S(stream));
}

int carc_index_table(int nlevels, const int64_t* extents, const int64_t* strides, void* table_dev, void* stream) {
  return carc::index_table(nlevels, extents, strides, (int64_t*)table_dev, S(stream));
}

int carc_zgemm_tab(int opA, int opB, int64_t M, int64_t N, int64_t K, const double alpha[2], const void* A, int64_t lda,
                   const void* B, int64_t ldb, const double beta[2], void* C, const void* rowoff_dev,
                   const void* coloff_dev, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                   void* stream) {
  keep_pool_memory();
  return carc::zgemm(opA, opB, M, N, K, C2(alpha), (const cplx*)A, lda, (const cplx*)B, ldb, C2(beta), (cplx*)C, nullptr,
                     nullptr, batch, strideA, strideB, strideC, S(stream), (const int64_t*)rowoff_dev,
                     (const int64_t*)coloff_dev);
}

// ---------------------------------------------------------------------------------------------------
int carc_operator_create(carc_operator** op, int P, int Q, int R, int Sd, int d) {
  CARC_REQUIRE(op != nullptr, CARC_ERR_VALUE, "operator_create: null handle pointer");
  CARC_REQUIRE(P > 0 && Q > 0 && R > 0 && Sd > 0 && d > 0 && d <= 8, CARC_ERR_VALUE,
               "operator_create: invalid dimensions P=%d Q=%d R=%d S=%d d=%d", P, Q, R, Sd, d);
  keep_pool_memory();
  carc_operator* o = new carc_operator();
  o->P = P; o->Q = Q; o->R = R; o->S = Sd; o->d = d;
  o->n = (int64_t)P * R * d;
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
  for (int i = 0; i < 64; ++i) t.op[i] = make_double2(0.0, 0.0);
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
  int prc = carc::stage3_plan_create(op->terms.data(), nt, &op->plan);
  if (prc) return prc;
  op->workspace_elems = carc::stage3_workspace_elems(nt, op->P, op->Q, op->R, op->S, op->d, Xmax);
  op->finalized = true;
  return CARC_OK;
}

int carc_operator_set_path(carc_operator* op, int force_path) {
  CARC_REQUIRE(op && force_path >= 0 && force_path <= 3, CARC_ERR_VALUE, "operator_set_path: invalid argument");
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

int carc_operator_create_dense(carc_operator** op, const void* matrix_dev, int64_t n) {
  CARC_REQUIRE(op != nullptr && matrix_dev != nullptr && n > 0, CARC_ERR_VALUE, "operator_create_dense: invalid argument");
  keep_pool_memory();
  carc_operator* o = new carc_operator();
  o->kind = 1;
  o->matrix = (const cplx*)matrix_dev;
  o->n = n;
  o->P = o->Q = o->R = o->S = o->d = 0;
  o->finalized = true;
  *op = o;
  return CARC_OK;
}

int64_t carc_operator_dimension(const carc_operator* op) { return op ? op->n : -1; }

int carc_operator_path(const carc_operator* op);
double carc_operator_executed_flops(const carc_operator* op) {
  if (!op || op->kind != 0 || !op->plan) return -1.0;
  int column_blocks = 1;
  if (carc_operator_path(op) == 3) {
    int64_t Xmax = 0;
    for (const auto& t : op->terms) Xmax = std::max<int64_t>(Xmax, t.X);
    carc::Stage3FConfig k;
    if (carc::stage3f_configure((int)op->terms.size(), op->P, op->Q, op->R, op->S, op->d, Xmax, &k)) column_blocks = k.RB;
  }
  return carc::stage3_executed_flops(op->plan, op->P, op->Q, op->R, op->S, op->d, column_blocks);
}
int carc_operator_path(const carc_operator* op) {
  if (!op || op->kind != 0 || !op->finalized) return -1;
  int64_t Xmax = 0;
  for (const auto& t : op->terms) Xmax = std::max<int64_t>(Xmax, t.X);
  return carc::stage3_path((int)op->terms.size(), op->P, op->Q, op->R, op->S, op->d, Xmax, op->force_path);
}
int carc_stage3f_profile_read(unsigned long long* host) { return carc::stage3f_profile_read(host); }
int carc_stage3_path(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int force_path) {
  return carc::stage3_path(nterms, P, Q, R, S, d, Xmax, force_path);
}
int carc_stage3_describe_stars(int nterms, const int32_t* a_id, const int32_t* b_id, const int64_t* X, int32_t* n_groups,
                               int32_t* group_kind, int32_t* group_first, int32_t* group_count, int32_t* sorted_term) {
  CARC_REQUIRE(nterms >= 0 && a_id && b_id && X && n_groups && group_kind && group_first && group_count && sorted_term,
               CARC_ERR_VALUE, "stage3_describe_stars: invalid argument");
  // stand-in tensors: the planner only compares pointers, so distinct ids become distinct (never dereferenced) addresses
  std::vector<carc::Stage3Term> terms((size_t)nterms);
  for (int t = 0; t < nterms; ++t) {
    terms[t].A = reinterpret_cast<const cplx*>((uintptr_t)(a_id[t] + 1) * 4096);
    terms[t].B = reinterpret_cast<const cplx*>((uintptr_t)(b_id[t] + 1) * 4096 + 2048);
    terms[t].X = X[t];
    terms[t].has_op = 0;
    terms[t].op[0] = make_double2((double)t, 0.0);      // carries the original index through the sort
  }
  carc::Stage3Plan plan;
  carc::stage3_plan_host(terms.data(), nterms, &plan);
  *n_groups = (int32_t)plan.groups.size();
  for (size_t g = 0; g < plan.groups.size(); ++g) {
    group_kind[g] = plan.groups[g].kind;
    group_first[g] = plan.groups[g].first;
    group_count[g] = plan.groups[g].count;
  }
  for (int t = 0; t < nterms; ++t) sorted_term[t] = (int32_t)plan.terms[t].op[0].x;
  return CARC_OK;
}
int carc_stage3f_describe(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int32_t* out, int out_len) {
  CARC_REQUIRE(out != nullptr && out_len >= 16 + 17 + 17 + 160 + 160, CARC_ERR_VALUE, "stage3f_describe: buffer too small");
  carc::Stage3FConfig k;
  if (!carc::stage3f_configure(nterms, P, Q, R, S, d, Xmax, &k)) {
    carc::set_error("stage3f_describe: shape outside the folded kernel's envelope");
    return CARC_ERR_UNSUPPORTED;
  }
  const int32_t head[16] = {k.NPT, k.NRT, k.Q4, k.NSB, k.G, k.nstA, k.nstB, k.QS, k.BSTR, k.b_whole, k.threads, k.ctas,
                            k.slots, (int32_t)k.total, (int32_t)k.slotA, (int32_t)k.slotB};
  int o = 0;
  for (int i = 0; i < 16; ++i) out[o++] = head[i];
  for (int i = 0; i < 17; ++i) out[o++] = i <= k.NSB ? k.sb_tile0[i] : -1;
  for (int i = 0; i < 17; ++i) out[o++] = i <= k.NSB ? k.sb_cta0[i] : -1;
  for (int i = 0; i < 160; ++i) out[o++] = i < k.ctas ? k.cta_sb[i] : -1;
  for (int i = 0; i < 160; ++i) out[o++] = i < k.ctas ? k.cta_sl[i] : -1;
  if (out_len >= o + 2) {     // row / column blocks of the output (NPT, NRT above are those of the largest block)
    out[o++] = k.PB;
    out[o++] = k.RB;
  }
  if (out_len >= o + 1) out[o++] = k.ws;   // consumer-warp slots of the warp-specialised kernel (0: symmetric kernel)
  return CARC_OK;
}
int carc_operator_num_groups(const carc_operator* op) { return (op && op->plan) ? (int)op->plan->groups.size() : -1; }

int carc_operator_apply(carc_operator* op, const void* v, void* out, void* stream) {
  CARC_REQUIRE(op && op->finalized, CARC_ERR_VALUE, "operator_apply: operator not finalized");
  if (op->kind == 1)
    return carc::dense_matvec(op->matrix, op->n, op->n, op->n, (const cplx*)v, (cplx*)out, make_double2(1.0, 0.0),
                              make_double2(0.0, 0.0), S(stream));
  keep_pool_memory();
  cudaStream_t st = S(stream);
  int rc = carc::stage3_plan_upload(op->plan, st);
  if (rc) return rc;
  if (!op->workspace) {
    CARC_CHECK_CUDA(cudaMallocAsync((void**)&op->workspace, sizeof(cplx) * (op->workspace_elems > 0 ? op->workspace_elems : 1), st));
    op->workspace_stream = st;
  } else if (op->workspace_stream != st) {
    CARC_CHECK_CUDA(cudaStreamSynchronize(op->workspace_stream));
    op->workspace_stream = st;
  }
  return carc::stage3_apply(op->plan, op->P, op->Q, op->R, op->S, op->d, (const cplx*)v, (cplx*)out, op->workspace,
                            op->workspace_elems, op->force_path, st, op->comm);
}

// ---------------------------------------------------------------------------------------------------
int carc_comm_create(carc_comm** comm, int rank, int world, int64_t max_elems) {
  CARC_REQUIRE(comm != nullptr, CARC_ERR_VALUE, "comm_create: null handle pointer");
  carc_comm* c = new carc_comm();
  int rc = carc::comm_create(&c->impl, rank, world, max_elems);
  if (rc) {
    delete c;
    return rc;
  }
  *comm = c;
  return CARC_OK;
}
int carc_comm_local_handles(carc_comm* comm, void* out128) {
  CARC_REQUIRE(comm, CARC_ERR_VALUE, "comm_local_handles: null communicator");
  return carc::comm_local_handles(comm->impl, out128);
}
int carc_comm_connect(carc_comm* comm, const void* all_handles) {
  CARC_REQUIRE(comm, CARC_ERR_VALUE, "comm_connect: null communicator");
  return carc::comm_connect(comm->impl, all_handles);
}
int carc_comm_allreduce(carc_comm* comm, void* data, int64_t n, void* stream) {
  CARC_REQUIRE(comm && data, CARC_ERR_VALUE, "comm_allreduce: null argument");
  return carc::comm_allreduce(comm->impl, (const cplx*)data, 1, n, (cplx*)data, S(stream));
}
int carc_comm_status(carc_comm* comm, int* timed_out) {
  CARC_REQUIRE(comm, CARC_ERR_VALUE, "comm_status: null communicator");
  return carc::comm_status(comm->impl, timed_out);
}
int carc_comm_destroy(carc_comm* comm) {
  if (!comm) return CARC_OK;
  carc::comm_destroy(comm->impl);
  delete comm;
  return CARC_OK;
}
int carc_operator_set_comm(carc_operator* op, carc_comm* comm) {
  CARC_REQUIRE(op && op->kind == 0, CARC_ERR_VALUE, "operator_set_comm: needs a stage-3 operator");
  op->comm = comm ? comm->impl : nullptr;
  return CARC_OK;
}

int carc_operator_destroy(carc_operator* op) {
  if (!op) return CARC_OK;
  carc::stage3_plan_destroy(op->plan);
  if (op->workspace) cudaFreeAsync(op->workspace, op->workspace_stream);
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

// ---------------------------------------------------------------------------------------------------
int carc_lu_factor(void* A, int n, void* piv_dev, int* singular_out, void* stream) {
  CARC_REQUIRE(A && piv_dev && n > 0, CARC_ERR_VALUE, "lu_factor: invalid argument");
  keep_pool_memory();
  cudaStream_t st = S(stream);
  void* scratch = nullptr;
  const size_t sbytes = carc::lu_scratch_bytes() + 64;
  CARC_CHECK_CUDA(cudaMallocAsync(&scratch, sbytes, st));
  int* singular_dev = (int*)((char*)scratch + carc::lu_scratch_bytes());
  int rc = carc::lu_factor((cplx*)A, n, (int*)piv_dev, singular_dev, (cplx*)scratch, st);
  int singular = 0;
  if (!rc) {
    CARC_CHECK_CUDA(cudaMemcpyAsync(&singular, singular_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CARC_CHECK_CUDA(cudaStreamSynchronize(st));
  }
  cudaFreeAsync(scratch, st);
  if (singular_out) *singular_out = singular;
  return rc;
}

int64_t carc_lu_inverse_blocks_elems(int n) { return carc::lu_inverse_blocks_elems(n); }
int carc_lu_invert_diagonal_blocks(const void* LU, int n, void* inv_blocks, void* stream) {
  CARC_REQUIRE(LU && inv_blocks && n > 0, CARC_ERR_VALUE, "lu_invert_diagonal_blocks: invalid argument");
  return carc::lu_invert_diagonal_blocks((const cplx*)LU, n, (cplx*)inv_blocks, S(stream));
}
int carc_lu_solve_blocks(const void* LU, int n, const void* piv_dev, const void* inv_blocks, void* x, void* stream) {
  CARC_REQUIRE(LU && piv_dev && inv_blocks && x && n > 0, CARC_ERR_VALUE, "lu_solve_blocks: invalid argument");
  keep_pool_memory();
  cudaStream_t st = S(stream);
  void* tmp = nullptr;
  CARC_CHECK_CUDA(cudaMallocAsync(&tmp, sizeof(cplx) * (size_t)n, st));
  int rc = carc::lu_solve_fast((const cplx*)LU, n, (const int*)piv_dev, (const cplx*)inv_blocks, (cplx*)x, (cplx*)tmp, st);
  cudaFreeAsync(tmp, st);
  return rc;
}

int carc_cholesky_factor_as_lu(void* A, int n, void* piv_dev, double hermitian_tolerance, int* status_out, void* stream) {
  CARC_REQUIRE(A && piv_dev && n > 0, CARC_ERR_VALUE, "cholesky_factor_as_lu: invalid argument");
  keep_pool_memory();
  cudaStream_t st = S(stream);
  double* scratch = nullptr;   // [0..1] the two sums of the Hermiticity check, [2] (as int) the pivot status
  CARC_CHECK_CUDA(cudaMallocAsync((void**)&scratch, 4 * sizeof(double), st));
  int status = 0;
  int rc = carc::hermitian_defect((const cplx*)A, n, scratch, st);
  if (!rc) {
    double sums[2] = {0.0, 0.0};
    CARC_CHECK_CUDA(cudaMemcpyAsync(sums, scratch, sizeof(sums), cudaMemcpyDeviceToHost, st));
    CARC_CHECK_CUDA(cudaStreamSynchronize(st));
    // |A - A^H|_F <= tol |A|_F (NaNs fail the comparison and are rejected too)
    if (!(sums[0] <= hermitian_tolerance * hermitian_tolerance * sums[1])) status = 2;
  }
  if (!rc && status == 0) {
    int* status_dev = reinterpret_cast<int*>(scratch + 2);
    rc = carc::cholesky_factor_as_lu((cplx*)A, n, (int*)piv_dev, status_dev, st);
    if (!rc) {
      CARC_CHECK_CUDA(cudaMemcpyAsync(&status, status_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
      CARC_CHECK_CUDA(cudaStreamSynchronize(st));
    }
  }
  cudaFreeAsync(scratch, st);
  if (status_out) *status_out = status;
  return rc;
}
int carc_lu_solve(const void* LU, int n, const void* piv_dev, void* x, void* stream) {
  CARC_REQUIRE(LU && piv_dev && x && n > 0, CARC_ERR_VALUE, "lu_solve: invalid argument");
  return carc::lu_solve((const cplx*)LU, n, (const int*)piv_dev, (cplx*)x, S(stream));
}

static carc::LinOp as_linop(carc_operator* op) {
  carc::LinOp l;
  l.apply = [op](const cplx* in, cplx* out, cudaStream_t st) { return carc_operator_apply(op, in, out, (void*)st); };
  return l;
}

int carc_gmres(carc_operator* A, const void* b, void* x, double rtol, int restart, int maxiter, int* iterations_out,
               double* residual_out, void* stream) {
  CARC_REQUIRE(A && A->finalized && b && x, CARC_ERR_VALUE, "gmres: invalid argument");
  cudaStream_t st = S(stream);
  const int64_t n = A->n;
  if (restart > n) restart = (int)n;
  void *work = nullptr, *state = nullptr;
  CARC_CHECK_CUDA(cudaMallocAsync(&work, sizeof(cplx) * (size_t)(restart + 2) * n, st));
  CARC_CHECK_CUDA(cudaMallocAsync(&state, carc::gmres_state_bytes(), st));
  int iters = 0;
  double resid = 0.0;
  int rc = carc::gmres(as_linop(A), (const cplx*)b, (cplx*)x, n, rtol, restart, maxiter, (cplx*)work, state, &iters,
                       &resid, st);
  cudaFreeAsync(work, st);
  cudaFreeAsync(state, st);
  if (iterations_out) *iterations_out = iters;
  if (residual_out) *residual_out = resid;
  return rc;
}

int carc_cg(carc_operator* A, const void* b, void* x, double rtol, int maxiter, int* iterations_out, double* residual_out,
            void* stream) {
  CARC_REQUIRE(A && A->finalized && b && x, CARC_ERR_VALUE, "cg: invalid argument");
  cudaStream_t st = S(stream);
  const int64_t n = A->n;
  void *work = nullptr, *state = nullptr;
  CARC_CHECK_CUDA(cudaMallocAsync(&work, sizeof(cplx) * (size_t)3 * n, st));
  CARC_CHECK_CUDA(cudaMallocAsync(&state, carc::cg_state_bytes(), st));
  int iters = 0;
  double resid = 0.0;
  int rc = carc::cg(as_linop(A), (const cplx*)b, (cplx*)x, n, rtol, maxiter, (cplx*)work, state, &iters, &resid, st);
  cudaFreeAsync(work, st);
  cudaFreeAsync(state, st);
  if (iterations_out) *iterations_out = iters;
  if (residual_out) *residual_out = resid;
  return rc;
}

int carc_relax(carc_operator* H, carc_operator* N_op, const void* N_lu, const void* N_piv, const void* N_inv_blocks,
               void* v, int max_mults, double tolerance, int krylov_dim, double gmres_rtol, int gmres_restart,
               int gmres_maxiter, double* info_out, void* stream) {
  CARC_REQUIRE(H && H->finalized && v, CARC_ERR_VALUE, "relax: invalid argument");
  CARC_REQUIRE(!N_op || N_op->n == H->n, CARC_ERR_DIMENSION_MISMATCH, "relax: H and N act on different spaces");
  cudaStream_t st = S(stream);
  const int64_t n = H->n;
  const int k = krylov_dim > 0 ? krylov_dim : 3;
  int restart = gmres_restart > 0 ? gmres_restart : 20;
  if (restart > n) restart = (int)n;
  keep_pool_memory();
  void *work = nullptr, *state = nullptr, *hv = nullptr, *gwork = nullptr, *gstate = nullptr, *lutmp = nullptr;
  CARC_CHECK_CUDA(cudaMallocAsync(&work, sizeof(cplx) * (size_t)(2 * k + 2) * n, st));
  if (N_lu && N_piv && N_inv_blocks) CARC_CHECK_CUDA(cudaMallocAsync(&lutmp, sizeof(cplx) * (size_t)n, st));
  CARC_CHECK_CUDA(cudaMallocAsync(&state, carc::relax_state_bytes(), st));
  const bool use_lu = N_lu != nullptr && N_piv != nullptr;
  const bool use_gmres = !use_lu && N_op != nullptr;
  if (use_gmres) {
    CARC_CHECK_CUDA(cudaMallocAsync(&hv, sizeof(cplx) * n, st));
    CARC_CHECK_CUDA(cudaMallocAsync(&gwork, sizeof(cplx) * (size_t)(restart + 2) * n, st));
    CARC_CHECK_CUDA(cudaMallocAsync(&gstate, carc::gmres_state_bytes(), st));
  }
  int gm_total = 0;
  carc::LinOp M;
  carc::LinOp Nl;
  if (use_gmres) Nl = as_linop(N_op);
  M.apply = [&](const cplx* in, cplx* out, cudaStream_t s2) -> int {
    if (use_lu) {
      int rc = carc_operator_apply(H, in, out, (void*)s2);
      if (rc) return rc;
      if (lutmp)
        return carc::lu_solve_fast((const cplx*)N_lu, (int)n, (const int*)N_piv, (const cplx*)N_inv_blocks, out,
                                   (cplx*)lutmp, s2);
      return carc::lu_solve((const cplx*)N_lu, (int)n, (const int*)N_piv, out, s2);
    }
    if (use_gmres) {
      int rc = carc_operator_apply(H, in, hv, (void*)s2);
      if (rc) return rc;
      int it = 0;
      double rs = 0.0;
      rc = carc::gmres(Nl, (const cplx*)hv, out, n, gmres_rtol > 0 ? gmres_rtol : 1e-5, restart,
                       gmres_maxiter > 0 ? gmres_maxiter : 1000, (cplx*)gwork, gstate, &it, &rs, s2);
      gm_total += it;
      return rc;
    }
    return carc_operator_apply(H, in, out, (void*)s2);
  };
  carc::RelaxInfo info;
  int rc = carc::relax(M, (cplx*)v, n, max_mults, tolerance, k, (cplx*)work, state, &info, st);
  cudaFreeAsync(work, st);
  cudaFreeAsync(state, st);
  if (lutmp) cudaFreeAsync(lutmp, st);
  if (use_gmres) {
    cudaFreeAsync(hv, st);
    cudaFreeAsync(gwork, st);
    cudaFreeAsync(gstate, st);
  }
  // sticky error words of the bounded device-side waits (the stream is idle here: relax() ends with a read-back)
  if (rc == CARC_OK || rc == CARC_ERR_EXCHANGE) {
    int timed_out = 0;
    for (carc_operator* o : {H, N_op})
      if (o && o->kind == 0 && o->comm && carc::comm_status(o->comm, &timed_out) == CARC_OK && timed_out) {
        carc::set_error("relax: the peer all-reduce of a sharded operator timed out (a rank fell out of step)");
        return CARC_ERR_EXCHANGE;
      }
    if (use_lu && N_inv_blocks && carc::lu_solve_status((const cplx*)N_inv_blocks, (int)n, &timed_out) == CARC_OK && timed_out) {
      carc::set_error("relax: the wavefront triangular solve timed out");
      return CARC_ERR_EXCHANGE;
    }
  }
  if (info_out && (rc == CARC_OK || rc == CARC_ERR_RELAX_FAILED)) {
    info_out[0] = info.initial_value[0]; info_out[1] = info.initial_value[1];
    info_out[2] = info.final_value[0]; info_out[3] = info.final_value[1];
    info_out[4] = info.ritz_value[0]; info_out[5] = info.ritz_value[1];
    info_out[6] = info.multiplications; info_out[7] = info.applications; info_out[8] = gm_total;
  }
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
