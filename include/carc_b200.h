/* carc_b200.h -- C ABI of libcarc_b200.so: the B200 (sm_100a) implementation of Carcassonne's center-site
 * optimisation hot path.
 *
 * The reference (gcross/Carcassonne) is pure Python and has no FFI; the seam this library sits behind is the
 * duck-typed tensor class `NDArrayData` (reference carcassonne/data/__init__.py:30-365) plus the `Multiplier`
 * / `relaxOver` pair (reference carcassonne/utils.py:180-207, 805-878).  Every entry point below names the
 * reference call it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - All tensors are complex128, interleaved (re, im), C-contiguous -- the memory layout of numpy.complex128 /
 *     torch.complex128.  Pointers named `*_dev` / without suffix are DEVICE pointers unless the function name ends
 *     in `_host`.  Device pointers must be 16-byte aligned.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on that stream
 *     unless stated otherwise; functions that return scalars to the host synchronise the stream.
 *   - Every function returns 0 on success or one of the CARC_ERR_* codes; carc_last_error() gives the message.
 *     Codes map 1:1 onto the reference's exceptions (utils.py:13-41).
 *   - Single host thread per device; one stream per System.
 */
#ifndef CARC_B200_H
#define CARC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARC_OK 0
#define CARC_ERR_CUDA 1               /* CUDA runtime failure                                         */
#define CARC_ERR_DIMENSION_MISMATCH 2 /* DimensionMismatchError (utils.py:14-24)                      */
#define CARC_ERR_RANK 3               /* UnexpectedTensorRankError (utils.py:35-41)                   */
#define CARC_ERR_VALUE 4              /* ValueError                                                   */
#define CARC_ERR_RELAX_FAILED 5       /* RelaxFailed (utils.py:26-34)                                 */
#define CARC_ERR_INVARIANT 6          /* InvariantViolatedError (utils.py:25)                         */
#define CARC_ERR_NO_CONVERGENCE 7     /* the reference's `assert info == 0` after GMRES (utils.py:824) */
#define CARC_ERR_UNSUPPORTED 8
#define CARC_ERR_EXCHANGE 9        /* a bounded device-side wait (peer all-reduce, wavefront solve) timed out */

/* operand flags of carc_zgemm */
#define CARC_OP_N 0 /* as stored                     */
#define CARC_OP_T 1 /* transposed                    */
#define CARC_OP_C 2 /* conjugate-transposed          */
#define CARC_OP_J 3 /* conjugated, not transposed    */

typedef struct carc_operator carc_operator; /* device-resident expectation / normalization operator */
typedef struct carc_comm carc_comm;         /* one-node multi-GPU communicator over NVLink peer memory */

/* ---- library ------------------------------------------------------------------------------------------- */
int carc_version(void);
const char* carc_last_error(void);
/* Measured issue rate of DMMA.8x8x4 on the current device in TFLOP/s (the FP64 tensor roofline). */
int carc_dmma_peak(int iters, double* tflops_out, void* stream);
/* The same microbenchmark with `warps_per_sm` warps (one CTA per SM) and `chains` (2, 4, 8 or 16) independent
 * accumulator chains per warp: how the DMMA issue rate depends on occupancy and instruction-level parallelism. */
int carc_dmma_rate(int iters, int warps_per_sm, int chains, double* tflops_out, void* stream);
/* DMMA.8x8x4 and DFMA issued side by side (`ndmma` in {0, 16} and `nfma` in {0, 16, 32, 64, 128} instructions per loop
 * trip and warp), rates reported separately in TFLOP/s: whether the FP64 tensor instruction and the FP64 FMA pipe
 * overlap on this part decides whether tile padding / remainders can be moved off the tensor pipe. */
int carc_fp64_mix_rate(int iters, int warps_per_sm, int ndmma, int nfma, double* tflops_dmma, double* tflops_fma,
                       void* stream);

/* ---- device memory for callers that do not bring their own allocator ------------------------------------ */
int carc_malloc(void** ptr, size_t bytes);
int carc_free(void* ptr);
int carc_malloc_host(void** ptr, size_t bytes); /* pinned */
int carc_free_host(void* ptr);
int carc_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int carc_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int carc_stream_synchronize(void* stream);

/* ---- NDArrayData.join / transpose / conj (data/__init__.py:247-256, 154) --------------------------------
 * dst = transpose(src, perm) written C-contiguously; conj != 0 conjugates; accumulate != 0 adds into dst. */
int carc_permute(const void* src, void* dst, int ndim, const int64_t* shape, const int32_t* perm, int conj,
                 int accumulate, void* stream);

/* ---- NDArrayData + - * += scalar*  (data/__init__.py:95-147):  y = alpha * (conj_x ? conj(x) : x) + beta * y */
int carc_axpby(int64_t n, const double alpha[2], const void* x, const double beta[2], void* y, int conj_x,
               void* stream);
/* ---- NDArrayData.absorbMatrixAt with a short matrix (data/__init__.py:148-150; the compressor projections of
 * system/_2d.py:199-228):  out[pre][j][post] = sum_k matrix[j][k] * x[pre][k][post],  1 <= rows <= 16,
 * k <= 640.  HBM-bound streaming kernel; larger matrices go through carc_zgemm (batched). */
int carc_mode_product(const void* matrix, const void* x, void* out, int64_t rows, int64_t k, int64_t pre, int64_t post,
                      void* stream);
/* y *= x elementwise (NDArrayData.__imul__) */
int carc_mul(int64_t n, const void* x, void* y, void* stream);

/* ---- reductions (NDArrayData.norm, contractWithAlongAll, hasNaN; utils.py:845-870 Arnoldi scalars) --------
 * Results are written to DEVICE memory (two doubles) so iteration loops never synchronise. */
int carc_dotc(int64_t n, const void* x, const void* y, void* out2_dev, void* stream); /* sum conj(x) y  */
int carc_dotu(int64_t n, const void* x, const void* y, void* out2_dev, void* stream); /* sum x y: contractWithAlongAll (data/__init__.py:160-163) */
int carc_sumsq(int64_t n, const void* x, void* out2_dev, void* stream);               /* (sum |x|^2, 0) */
int carc_count_nonfinite(int64_t n, const void* x, void* out2_dev, void* stream);     /* (#nonfinite, #nan) */

/* ---- NDArrayData.contractWith == numpy.tensordot (data/__init__.py:157-159) after the operands are viewed
 * as matrices: C[M,N] = alpha * op(A) * op(B) + beta * C, row-major.  See carc_internal.h for the storage each
 * op flag implies.  out_map (6 x int64: m_div, m_s1, m_s0, n_div, n_s1, n_s0) or NULL for plain row-major with
 * ldc = N: element (m, n) goes to C[(m / m_div) * m_s1 + (m % m_div) * m_s0 + (n / n_div) * n_s1 + (n % n_div) *
 * n_s0], which lets a contraction write directly into the layout NDArrayData.join would produce.
 * k_map (4 x int64: a_kdiv, a_ks1, b_kdiv, b_ks1) or NULL: two-level K index for K-contiguous operands. */
int carc_zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, const double alpha[2], const void* A, int64_t lda,
               const void* B, int64_t ldb, const double beta[2], void* C, const int64_t* out_map, const int64_t* k_map,
               int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC, void* stream);

/* C[N,N] = op(A) op(B) for a product known to be Hermitian (the Gram matrices L^H L and R R^H of the state
 * compression, compression.py:26-45): only the tiles on or above the diagonal are computed, the rest is mirrored. */
int carc_zgemm_hermitian(int opA, int opB, int64_t N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                         void* C, void* stream);

/* Same product with arbitrary output scatter: element (m, n) goes to C[rowoff[m] + coloff[n]] (device int64 tables,
 * either may be NULL = plain row-major factor).  This is how the dense recipes (tensors/_2d/dense.py:11-112) write
 * their GEMM result straight into the layout their final `join` asks for -- up to 12 interleaved axes -- without the
 * transposing copy.  carc_index_table fills a table: table[i] = sum_l digit_l(i) * strides[l], digits row-major. */
int carc_index_table(int nlevels, const int64_t* extents, const int64_t* strides, void* table_dev, void* stream);
int carc_zgemm_tab(int opA, int opB, int64_t M, int64_t N, int64_t K, const double alpha[2], const void* A, int64_t lda,
                   const void* B, int64_t ldb, const double beta[2], void* C, const void* rowoff_dev,
                   const void* coloff_dev, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                   void* stream);

/* ---- the center-site operator: formExpectationStage3 / formNormalizationStage3 / formDenseStage3
 * (tensors/_2d/sparse.py:100-161, tensors/_2d/dense.py:115-203).
 * An operator is a list of terms (A_t, B_t, O_t): A_t = stage-2 half 0 pre-joined to [X_t, P, Q]
 * (= [(x y), D0*, D1*, D0, D1], dense.py:130), B_t = half 1 pre-joined to [X_t, R, S] (dense.py:131), O_t a d x d
 * site operator (row-major O[s', s]) or NULL for the identity.  The tensors are NOT copied: the caller keeps them
 * alive for the life of the operator.  apply computes out[P,R,d] = sum_t B_t . (A_t . (O_t v)) with v[Q,S,d]. */
int carc_operator_create(carc_operator** op, int P, int Q, int R, int S, int d);
int carc_operator_add_term(carc_operator* op, const void* A, const void* B, int64_t X, const double* O_host);
int carc_operator_finalize(carc_operator* op);
int carc_operator_apply(carc_operator* op, const void* v, void* out, void* stream);
/* force_path: 0 auto, 1 fused kernel only (error if the shape is unsupported), 2 unfused GEMM path,
 * 3 fused kernel with the folded tiling only (spin index folded into the tile columns; error if unsupported) */
int carc_operator_set_path(carc_operator* op, int force_path);
/* Which device path carc_operator_apply runs for this operator: 1 fused, 3 fused with the folded tiling, 2 unfused. */
int carc_operator_path(const carc_operator* op);
/* Diagnostics (library built with -DS3F_PROFILE only, else CARC_ERR_UNSUPPORTED): cycles every warp of the last
 * folded-tiling launch spent in its three mbarrier waits and in total, [148][12][4] values. */
int carc_stage3f_profile_read(unsigned long long* host);
/* Host-side launch plan of the folded fused kernel for one shape (no device call): out = {NPT, NRT, Q4, NSB, G, nstA,
 * nstB, QS, BSTR, b_whole, threads, ctas, slots, smem bytes, A slot bytes, B slot bytes}, then sb_tile0[17], sb_cta0[17],
 * cta_sb[160], cta_sl[160] (-1 beyond the used entries) and, when out_len >= 372, {PB, RB}: the output is computed in
 * PB row blocks x RB column blocks (P or R > 64; NPT / NRT / threads are then those of the largest block);
 * CARC_ERR_UNSUPPORTED outside the kernel's envelope.  What the CPU test suite checks the work partition with. */
/* Which device path a stage-3 operator of this shape takes (host-side decision, no device call): 1 fused kernel,
 * 3 fused kernel with the folded tiling, 2 unfused GEMMs; force_path as in carc_operator_set_path. */
int carc_stage3_path(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int force_path);
/* Host-side star decomposition of a stage-3 term list (no device call): term t joins half-0 tensor a_id[t] and half-1
 * tensor b_id[t] over X[t] environment indices.  Outputs the groups (kind 0: terms sharing a half-1 tensor, first
 * products summed before one second product; kind 1: terms sharing a half-0 tensor, one first product reused), each
 * with its range [first, first + count) in the sorted term order, and sorted_term[i] = original index of sorted term i.
 * All output arrays hold nterms entries. */
int carc_stage3_describe_stars(int nterms, const int32_t* a_id, const int32_t* b_id, const int64_t* X, int32_t* n_groups,
                               int32_t* group_kind, int32_t* group_first, int32_t* group_count, int32_t* sorted_term);
/* Host-side launch plan of the folded fused kernel for one shape (no device call; tests/test_stage3f_plan.py reads it):
 * out[0..15] = NPT, NRT, Q4, NSB, G, nstA, nstB, QS, BSTR, b_whole, consumer threads, ctas, slots, smem bytes, slotA, slotB;
 * then sb_tile0[17], sb_cta0[17], cta_sb[160], cta_sl[160]; then, if out_len allows, PB, RB (output row / column blocks)
 * and the consumer-warp slots of the warp-specialised kernel (8 or 12; 0 = symmetric kernel).  out_len >= 370. */
int carc_stage3f_describe(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int32_t* out, int out_len);
int carc_operator_num_terms(const carc_operator* op);
/* cmac count the reference's CostTracker assigns to one apply (data/cost_tracker.py:17-21) */
int64_t carc_operator_cost_of_multiply(const carc_operator* op);
/* After finalize the term list is decomposed into stars: terms sharing a half-1 tensor B have their first products
 * summed before ONE second product, terms sharing a half-0 tensor A reuse ONE first product (linearity).
 * num_groups = number of stars; executed_flops = FP64 flops the fused kernel issues per apply (the roofline
 * numerator), <= 8 x cost_of_multiply. */
int carc_operator_num_groups(const carc_operator* op);
double carc_operator_executed_flops(const carc_operator* op);
int carc_operator_destroy(carc_operator* op);
/* End-to-end convenience with HOST buffers (terms, v and out on the host; copies inside the call):
 * A_host[t], B_host[t] are host pointers to [X[t],P,Q] / [X[t],R,S]; O_host[t] NULL or d*d complex. */
int carc_stage3_matvec_host(int nterms, const void* const* A_host, const void* const* B_host, const int64_t* X,
                            const double* const* O_host, int P, int Q, int R, int S, int d, const void* v_host,
                            void* out_host, void* stream);

/* ---- multi-GPU (no counterpart in the single-process reference) ---------------------------
 * The matvec shards over the joined environment bond X: each rank adds the X-slab of every term it owns
 * (carc_operator_add_term with the slab's pointers and extent) and attaches a communicator; carc_operator_apply then
 * finishes with a one-shot all-reduce of the output vector over NVLink peer memory (the slot-sum pass of stage 3
 * reads every peer's exchange buffer directly), so every rank returns the full, bit-identical H v.
 * Set-up: carc_comm_create on every rank (one process per GPU); exchange the 128-byte records of
 * carc_comm_local_handles between ranks by any host transport (torch.distributed in the Python layer) into a
 * [world][128] array; carc_comm_connect.  max_elems bounds the vector length.  Collective calls must be issued in the
 * same order on all ranks.  Waits are bounded; carc_comm_status reports a timed-out exchange. */
int carc_comm_create(carc_comm** comm, int rank, int world, int64_t max_elems);
int carc_comm_local_handles(carc_comm* comm, void* out128);
int carc_comm_connect(carc_comm* comm, const void* all_handles);
int carc_comm_allreduce(carc_comm* comm, void* data, int64_t n, void* stream); /* in-place sum over ranks */
int carc_comm_status(carc_comm* comm, int* timed_out);
int carc_comm_destroy(carc_comm* comm);
int carc_operator_set_comm(carc_operator* op, carc_comm* comm);

/* A dense operator (the `isCheaperToFormMatrix` branches of relaxOver, utils.py:814-832): matrix [n, n] row-major,
 * not copied. */
int carc_operator_create_dense(carc_operator** op, const void* matrix_dev, int64_t n);
int64_t carc_operator_dimension(const carc_operator* op);

/* ---- the center-site eigen-solver: relaxOver (utils.py:805-878) -------------------------------------------------
 * Restarted Arnoldi of dimension krylov_dim (0 = the reference's default 3) on N^-1 H with classical Gram-Schmidt,
 * a k x k non-Hermitian eigen-solve and a restart on the Ritz vector of lowest real part, entirely on device; `v`
 * (n = carc_operator_dimension(H) complex numbers) is normalised in place, as the reference does with the caller's
 * array (utils.py:808-809), and holds the result on return.  N^-1 is applied by
 *   - N_lu / N_piv != NULL : the LU factors of the dense normalization matrix from carc_lu_factor
 *                            (scipy.linalg.lu_factor / lu_solve, utils.py:816-818); N_inv_blocks (optional, from
 *                            carc_lu_invert_diagonal_blocks) selects the one-launch-per-block substitution;
 *   - else N_op != NULL    : GMRES(gmres_restart) on the operator to relative residual gmres_rtol
 *                            (scipy.sparse.linalg.gmres defaults 20 / 1e-5, utils.py:819-825); failure to converge
 *                            returns CARC_ERR_NO_CONVERGENCE (the reference's `assert info == 0`);
 *   - else                  : identity (standard eigenproblem).
 * Stops when |ritz - last ritz| <= tolerance, after max_mults multiplications (0 = no limit; counted in blocks of
 * krylov_dim like utils.py:860) or when the Krylov space is exhausted (norm <= 1e-14).  Returns
 * CARC_ERR_RELAX_FAILED under the reference's RelaxFailed condition (utils.py:871).
 * info_out (9 doubles, host): initial <v,Mv> (re, im), final (re, im), last Ritz value (re, im), multiplications
 * counted, operator applications performed, GMRES inner iterations. */
int carc_relax(carc_operator* H, carc_operator* N_op, const void* N_lu, const void* N_piv, const void* N_inv_blocks,
               void* v, int max_mults, double tolerance, int krylov_dim, double gmres_rtol, int gmres_restart,
               int gmres_maxiter, double* info_out, void* stream);
/* x = A^-1 b by restarted GMRES from x0 = 0 (scipy.sparse.linalg.gmres call sites utils.py:823, compression.py:39). */
int carc_gmres(carc_operator* A, const void* b, void* x, double rtol, int restart, int maxiter, int* iterations_out,
               double* residual_out, void* stream);
/* x = A^+ b for a Hermitian positive semi-definite operator by conjugate gradients from x0 = 0 (the normal equations
 * A^H A x = A^H b of computeProductCompressor, compression.py:36-43, which the reference hands to GMRES; CG reaches
 * the same minimum-norm solution without GMRES(20)'s restart stagnation).  Stops at |r| <= rtol |b|, after maxiter
 * iterations, or when the residual stagnates at the rounding floor; *residual_out is the final |r|. */
int carc_cg(carc_operator* A, const void* b, void* x, double rtol, int maxiter, int* iterations_out, double* residual_out,
            void* stream);
/* In-place LU with partial pivoting of a row-major n x n matrix (LAPACK zgetrf pivoting rule); piv_dev: int32[n] on
 * device.  Synchronises to report *singular_out (1 if a zero pivot was met).  carc_lu_solve overwrites x with A^-1 x. */
int carc_lu_factor(void* A, int n, void* piv_dev, int* singular_out, void* stream);
/* The same factors for a Hermitian positive definite matrix (the normalization matrix of utils.py:816-818 when the
 * environment is a proper double layer) by a blocked Cholesky factorisation -- no pivot search, half the update work,
 * all of it DMMA GEMMs -- rewritten in place as unit-lower L' = L D^-1, upper U' = D L^H with identity pivots, so
 * carc_lu_solve / carc_lu_solve_blocks / carc_relax consume them unchanged.  Synchronises to report *status_out:
 *   0  factorised;
 *   2  |A - A^H|_F > hermitian_tolerance |A|_F (or A holds NaNs): A is left untouched;
 *   1  a pivot was <= 0 or not finite: A is garbage, factor a fresh copy with carc_lu_factor. */
int carc_cholesky_factor_as_lu(void* A, int n, void* piv_dev, double hermitian_tolerance, int* status_out, void* stream);
int carc_lu_solve(const void* LU, int n, const void* piv_dev, void* x, void* stream);
/* The same solve for many right-hand sides in sequence (one per Arnoldi multiplication): invert the 64 x 64 diagonal
 * blocks of L and U once into inv_blocks (carc_lu_inverse_blocks_elems(n) complex numbers: the inverted blocks, the
 * hand-off flags of the wavefront substitutions, and the row interchanges of the factorisation as one gather
 * permutation, built on the first solve).  Each substitution is then ONE launch; call carc_lu_invert_diagonal_blocks
 * again whenever LU / piv_dev are overwritten by a new factorisation. */
int64_t carc_lu_inverse_blocks_elems(int n);
int carc_lu_invert_diagonal_blocks(const void* LU, int n, void* inv_blocks, void* stream);
int carc_lu_solve_blocks(const void* LU, int n, const void* piv_dev, const void* inv_blocks, void* x, void* stream);

/* ---- small factorisations: scipy.linalg.qr / svd call sites of NDArrayData.qr, svd, unitize, normalizeAxis
 * (data/__init__.py:43-50, 263-301, 344-346; utils.py:879-881).
 * carc_qr: A [m,n] row-major, m >= n (overwritten with reflectors); R [n,n]; Q [m,n]; tau [n] workspace.
 *          Householder conventions of LAPACK zgeqr2 / zung2r, so Q matches SciPy's economic Q to rounding.
 * carc_svd_small: one-sided Jacobi SVD of an n x n matrix (n <= 80): U [n,n], S [n] stored as (s,0) complex,
 *          descending, Vh [n,n]; null directions of U are completed to an orthonormal basis.
 * carc_normalizer_matrices: from (U, S, Vh) the n x n matrices normalizeAxis returns: polar = U Vh,
 *          normalizer = conj(V S^-1 V^H), denormalizer = V S V^H and their sqrt_svals variants
 *          (S^-1 skipped where S <= dont_recip_under). */
int carc_qr(void* A, int64_t m, int n, void* R, void* Q, void* tau, void* stream);
int carc_svd_small(const void* R, int n, void* U, void* S, void* Vh, void* stream);
int carc_normalizer_matrices(const void* U, const void* S, const void* Vh, int n, double dont_recip_under,
                             void* polar, void* normalizer, void* denormalizer, void* normalizer_sqrt,
                             void* denormalizer_sqrt, void* stream);

/* ---- one call per recipe -------------------------------------------------------------------
 * The boundary algebra of tensors/_2d/dense.py and the two solver-side helpers of the sweep, each as ONE entry point on
 * plain device pointers and int64 shapes (corner [c0 c1 c2 | c3 c4 c5], side [s0 s1 s2 | s3 s4 s5 | s6 s7], center
 * [right, up, left, down, physical]).  `accumulate` != 0 adds into `out` (the `result[tag] += ...` of
 * contractSparseTensors, sparse.py:236).  All asynchronous on `stream`. */
/* absorbDenseSideIntoCornerFromLeft / FromRight (dense.py:11-21): out [s0,s1,s2,c3 s6,c4 s7,c5] / [c0 s6,c1 s7,c2,s3,s4,s5] */
int carc_absorb_side_into_corner(const void* corner, const int64_t* corner_shape, const void* side, const int64_t* side_shape,
                                 int from_left, void* out, int accumulate, void* stream);
/* absorbDenseCenterSS/SOSIntoSide (dense.py:23-81) in two steps, so that the double-layer center of one site operator is
 * shared by every sparse tag pair: E[(g h), (vL wL vR wR vO wO)] = sum_{s z} center[.., s] O[z, s] conj[.., z] (operator_dev
 * NULL = identity; g / h the center / conjugate legs facing side `direction`), then
 * out [s0 vL, s1 wL, s2, s3 vR, s4 wR, s5, vO, wO] = sum_{g h} side[.., g, h] E[(g h), ..];
 * dims = {n_i, m_i, n_l, m_l, n_r, m_r, n_o, m_o} for center_shape n, conj_shape m and i = direction, l / r / o its
 * left / right / opposite legs. */
int carc_double_layer_center(int direction, const void* center, const int64_t* center_shape, const void* center_conj,
                             const int64_t* conj_shape, const void* operator_dev, void* E, void* stream);
int carc_absorb_center_into_side(const void* side, const int64_t* side_shape, const void* E, const int64_t* dims, void* out,
                                 int accumulate, void* stream);
/* formNormalizationStage1 / Stage2 (dense.py:96-112): the two stages of the environment build.  half = -1: the
 * reference's layout [B0,A1,A2,B2,A3,B3]; half = 0 / 1: the layout the center-site operator streams, [(B0 A1),A3,B3,A2,B2] /
 * [(A1 B0),A3,B3,A2,B2] (the pre-joins of dense.py:130-131).  slab_world > 1 builds only rank slab_rank's slab of the joined
 * environment bond (contiguous in its slow factor; both halves of a ring use the same index range), which is how the
 * environment is sharded over GPUs. */
int carc_form_stage1(const void* corner, const int64_t* corner_shape, const void* side, const int64_t* side_shape, void* out,
                     int accumulate, void* stream);
int carc_form_stage2(const void* stage1_a, const int64_t* a_shape, const void* stage1_b, const int64_t* b_shape, int half,
                     int slab_rank, int slab_world, void* out, int accumulate, void* stream);
/* formMatrix of a stage-3 multiplier (dense.py:176-194): out[(P R s'), (Q S s)] (+)= sum_X A[X,(P Q)] B[X,(R S)] O[s', s]
 * for the pre-joined halves A = [X, P, Q], B = [X, R, S] and a d x d (d <= 4) site operator on the device, row-major
 * [s'][s].  accumulate == 0 zeroes `out` ((P R d) x (Q S d)) first; X == 0 (an empty slab) contributes nothing. */
int carc_stage3_form_matrix(const void* A, const void* B, int64_t X, int64_t P, int64_t Q, int64_t R, int64_t S,
                            const void* operator_dev, int d, void* out, int accumulate, void* stream);
/* NDArrayData.normalizeAxis (data/__init__.py:263-301) for shape[axis] in 2..80: normalized (same shape as t) =
 * Q (U V^H) of the SVD of [(other axes), axis]; normalizer = conj(V S^-1 V^H), denormalizer = V S V^H, n x n
 * (S^-1 skipped where S <= dont_recip_under); sqrt_svals != 0 returns the square-root variants and no tensor.  Output
 * pointers may be NULL. */
int carc_normalize_axis(const void* t, const int64_t* shape, int ndim, int axis, int sqrt_svals, double dont_recip_under,
                        void* normalized, void* normalizer, void* denormalizer, void* stream);
/* computeProductCompressor (compression.py:26-45), operator bond 1: L [l, old, old, 1], R [old, old, 1, r]; `initial` the
 * random [old, new] draw (the host RNG stays with the caller so that seeded runs consume the reference's stream);
 * `sweeps` alternating-least-squares rounds (reference: 4) in Gram form -- the (l r) x (old new) matrix of the reference is
 * never formed -- each solved by the device LU after a relative diagonal shift `regularization` (1e-10) and followed
 * by the polar projection; left_gram / right_gram: optional L^H L / R R^H over the outer legs, [(old old), (old old)].
 * compressor_out [new, old].  No host round trip between rounds. */
int carc_product_compressor(const void* L, int64_t l, const void* R, int64_t r, int64_t old_dim, int64_t new_dim,
                            const void* initial, int sweeps, double regularization, const void* left_gram,
                            const void* right_gram, void* compressor_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CARC_B200_H */
