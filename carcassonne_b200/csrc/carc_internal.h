// Internal C++ interface between the translation units of libcarc_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <vector>

#define CARC_MAX_RANK 12

namespace carc {
typedef double2 cplx;

const char* get_error();
// Number of SMs of the current device (cached per device): grids of co-resident CTAs are sized from it.
int sm_count();

// tensor_ops.cu
int permute(const cplx* src, cplx* dst, int ndim, const int64_t* shape, const int32_t* perm, int conj, int accumulate,
            cudaStream_t stream);
int axpby(int64_t n, cplx alpha, const cplx* x, cplx beta, cplx* y, int conj_x, cudaStream_t stream);
int mode_product(const cplx* M, const cplx* x, cplx* out, int64_t j, int64_t k, int64_t pre, int64_t post,
                 cudaStream_t stream);
int mul_inplace(int64_t n, const cplx* x, cplx* y, cudaStream_t stream);
int reduce(int mode, int64_t n, const cplx* x, const cplx* y, double2* out_dev, cudaStream_t stream);

// zgemm.cu
enum Op { OP_N = 0, OP_T = 1, OP_C = 2, OP_J = 3 };
// C[M,N] = alpha * op(A) * op(B) + beta * C, row-major storage.
//   opA: N  A stored [M,K] (lda >= K);  T  stored [K,M] (lda >= M);  C  stored [K,M], conjugated;  J  [M,K] conjugated
//   opB: N  B stored [K,N] (ldb >= N);  T  stored [N,K] (ldb >= K);  C  stored [N,K], conjugated;  J  [K,N] conjugated
// The output element (m, n) is written at C + (m / m_div) * m_s1 + (m % m_div) * m_s0 + (n / n_div) * n_s1 +
// (n % n_div) * n_s0, so a GEMM can write straight into a permuted ("joined") layout.
struct GemmOut {
  int64_t m_div, m_s1, m_s0, n_div, n_s1, n_s0;
};
struct GemmKMap {
  int64_t a_kdiv, a_ks1, b_kdiv, b_ks1;
};
int zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, cplx alpha, const cplx* A, int64_t lda, const cplx* B,
          int64_t ldb, cplx beta, cplx* C, const GemmOut* out, const GemmKMap* kmap, int64_t batch, int64_t strideA,
          int64_t strideB, int64_t strideC, cudaStream_t stream, const int64_t* rowoff = nullptr,
          const int64_t* coloff = nullptr);
int zgemm_hermitian(int opA, int opB, int64_t N, int64_t K, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                    cplx* C, cudaStream_t stream);
int zgemm_lower(int opA, int opB, int64_t N, int64_t K, cplx alpha, const cplx* A, int64_t lda, const cplx* B, int64_t ldb,
                cplx beta, cplx* C, const GemmOut* out, cudaStream_t stream);
int index_table(int nlevels, const int64_t* extents, const int64_t* strides, int64_t* table, cudaStream_t stream);
int dmma_peak(int iters, double* tflops_out, cudaStream_t stream);
int dmma_rate(int iters, int warps, int chains, double* tflops_out, cudaStream_t stream);
int fp64_mix_rate(int iters, int warps, int ndmma, int nfma, double* tflops_dmma, double* tflops_fma, cudaStream_t stream);

// comm.cu
struct Comm;
int comm_create(Comm** out, int rank, int world, int64_t max_elems);
int comm_local_handles(Comm* c, void* out128);
int comm_connect(Comm* c, const void* all_handles);
int comm_allreduce(Comm* c, const cplx* src, int slots, int64_t n, cplx* out, cudaStream_t stream);
int comm_status(Comm* c, int* timed_out);
int comm_destroy(Comm* c);

// stage3.cu
struct Stage3Term {
  const cplx* A;   // [X, P, Q]   (pre-joined stage-2 half 0: [(x y), D0*, D1*, D0, D1])
  const cplx* B;   // [X, R, S]   (pre-joined stage-2 half 1: [(y' x'), D2*, D3*, D2, D3])
  int64_t X;
  int has_op;      // 0: identity on the physical leg
  cplx op[64];     // row-major d x d site operator O[s', s], d <= 8
};
struct Stage3Group {
  const cplx* center;   // the tensor every term of the star shares: B (kind 0) or A (kind 1)
  int64_t X;
  int first, count;     // terms [first, first + count) of the sorted term table
  int kind;             // 0: B-star (first products summed, one second product); 1: A-star (one first product reused)
};
struct Stage3Plan {
  std::vector<Stage3Term> terms;     // sorted by group
  std::vector<Stage3Group> groups;
  Stage3Term* terms_dev = nullptr;   // device copies: allocated from the stream-ordered pool and uploaded by
  Stage3Group* groups_dev = nullptr; // stage3_plan_upload on the first apply, freed on the stream that used them last
  cudaStream_t dev_stream = nullptr;
};
int stage3_plan_create(const Stage3Term* terms, int nterms, Stage3Plan** out);
int stage3_plan_upload(Stage3Plan* plan, cudaStream_t stream);
void stage3_plan_host(const Stage3Term* terms, int nterms, Stage3Plan* plan);
void stage3_plan_destroy(Stage3Plan* plan);
double stage3_executed_flops(const Stage3Plan* plan, int P, int Q, int R, int S, int d, int column_blocks = 1);
int stage3_apply(const Stage3Plan* plan, int P, int Q, int R, int S, int d, const cplx* v, cplx* out, cplx* workspace,
                 int64_t workspace_elems, int force_path, cudaStream_t stream, Comm* comm = nullptr);
int64_t stage3_workspace_elems(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax);
int stage3_path(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int force_path);

// stage3f.cu: the folded tiling of the fused kernel (spin index folded into the columns of the first product)
struct Stage3FConfig {
  int NPT, NRT, Q4, NSB, G, nstA, nstB, QS, BSTR, b_whole;
  int ws = 0;           // consumer-warp slots of the warp-specialised kernel (8 or 12), 0 = the symmetric kernel
  int PB = 1, RB = 1;   // row / column blocks of the output (P > 64 or R > 64); NPT, NRT are those of the largest block
  int sb_cta0[17], sb_tile0[17];
  unsigned char cta_sb[160], cta_sl[160];
  uint32_t slotA, slotB, ops_off, hasop_off, tab_off, vt_off, vtail_off, ring_off, total;
  int threads, ctas, slots;
  double padded_work;
};
int stage3f_profile_read(unsigned long long* host);
bool stage3f_configure(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, Stage3FConfig* cfg);
bool stage3f_configure_block(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, Stage3FConfig* cfg);
int stage3f_launch(const Stage3Plan* plan, const Stage3FConfig& k, int P, int Q, int R, int S, const cplx* v, cplx* partial,
                   cudaStream_t stream);

// linalg.cu
int qr(cplx* A, int64_t m, int n, cplx* R, cplx* Q, cplx* tau, cudaStream_t stream);
int svd_small(const cplx* R, int n, cplx* U, cplx* S, cplx* Vh, cudaStream_t stream);
int normalizer_matrices(const cplx* U, const cplx* S, const cplx* Vh, int n, double dont_recip_under, cplx* polar,
                        cplx* nrm, cplx* den, cplx* nrm_sqrt, cplx* den_sqrt, cudaStream_t stream);

// solver.cu
struct LinOp {
  std::function<int(const cplx*, cplx*, cudaStream_t)> apply;   // out = A in
};
struct RelaxInfo {
  double initial_value[2], final_value[2], ritz_value[2];
  int multiplications, applications;
};
int dense_matvec(const cplx* M, int64_t rows, int64_t cols, int64_t ld, const cplx* x, cplx* y, cplx alpha, cplx beta,
                 cudaStream_t stream);
int lu_factor(cplx* A, int n, int* piv, int* singular_dev, cplx* scratch, cudaStream_t stream);
int cholesky_factor_as_lu(cplx* A, int n, int* piv, int* status_dev, cudaStream_t stream);
int hermitian_defect(const cplx* A, int n, double* sums_dev, cudaStream_t stream);
int lu_solve(const cplx* LU, int n, const int* piv, cplx* x, cudaStream_t stream);
int lu_invert_diagonal_blocks(const cplx* LU, int n, cplx* inv, cudaStream_t stream);
int64_t lu_inverse_blocks_elems(int n);
// *timed_out = 1 when a wavefront solve on these inverse blocks gave up waiting (sticky); synchronises the device word
// lu_panel.cu: one 64-column panel of lu_factor by a thread-block cluster + the interchanges on all other columns;
// CARC_ERR_UNSUPPORTED (nothing launched) when the panel is taller than a cluster holds in registers
int lu_panel_cluster(cplx* A, int n, int j0, int nb, int* piv, int* singular, cudaStream_t stream);
int lu_trsm_unit_lower(cplx* A, int n, int j0, int nb, int c_first, int ncols, cudaStream_t stream);
int lu_solve_status(const cplx* inv, int n, int* timed_out);
int lu_solve_fast(const cplx* LU, int n, const int* piv, const cplx* inv, cplx* x, cplx* tmp, cudaStream_t stream);
int gmres(const LinOp& A, const cplx* b, cplx* x, int64_t n, double rtol, int restart, int maxiter, cplx* work,
          void* state_dev, int* iters_out, double* resid_out, cudaStream_t stream);
size_t gmres_state_bytes();
size_t lu_scratch_bytes();
int cg(const LinOp& A, const cplx* b, cplx* x, int64_t n, double rtol, int maxiter, cplx* work, void* state_dev,
       int* iters_out, double* resid_out, cudaStream_t stream);
size_t cg_state_bytes();
int relax(const LinOp& M, cplx* v, int64_t n, int max_mults, double tol, int k, cplx* work, void* state_dev,
          RelaxInfo* info, cudaStream_t stream);
size_t relax_state_bytes();

}  // namespace carc
