// Shared device helpers for the carcassonne_b200 kernels (sm_100a only).
//
// FP64 on Blackwell has no tcgen05 path (tcgen05.mma kinds are f16/tf32/f8f6f4/i8/mx*); the FP64 tensor
// instruction is mma.sync.m8n8k4.f64, which ptxas lowers to DMMA.8x8x4 on sm_100a.  All contractions in
// this library are built on that instruction; operands are staged through shared memory by
// cp.async / cp.async.bulk (TMA) and accumulators live in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace carc {

typedef double2 cplx;  // interleaved (re, im) == numpy complex128 == torch.complex128

// ---------------------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
#define CARC_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      carc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CARC_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)
#define CARC_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      carc::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// DMMA: D(8x8) += A(8x4) * B(4x8), one warp.
//   A fragment: lane holds A[row = lane/4][k = lane%4]
//   B fragment: lane holds B[k = lane%4][col = lane/4]
//   C fragment: lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Complex 8x8 accumulator tile held as separate real / imaginary C fragments.
struct CTile {
  double re0, re1, im0, im1;
  __device__ __forceinline__ void zero() { re0 = re1 = im0 = im1 = 0.0; }
};

// acc += a * b for complex fragments: four real DMMAs.  `nai` is -a_im (sign flipped once per load).
__device__ __forceinline__ void cmma(CTile& acc, double ar, double ai, double nai, double br, double bi) {
  dmma(acc.re0, acc.re1, ar, br);
  dmma(acc.re0, acc.re1, nai, bi);
  dmma(acc.im0, acc.im1, ar, bi);
  dmma(acc.im0, acc.im1, ai, br);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte shared load of one complex number.
__device__ __forceinline__ cplx lds_c(uint32_t addr) {
  cplx v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// ---------------------------------------------------------------------------------------------------
// cp.async (LDGSTS), 16 bytes, zero-fill when !valid
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace carc

// error codes shared with include/carc_b200.h
#define CARC_OK 0
#define CARC_ERR_CUDA 1
#define CARC_ERR_DIMENSION_MISMATCH 2
#define CARC_ERR_RANK 3
#define CARC_ERR_VALUE 4
#define CARC_ERR_RELAX_FAILED 5
#define CARC_ERR_INVARIANT 6
#define CARC_ERR_NO_CONVERGENCE 7
#define CARC_ERR_UNSUPPORTED 8
#define CARC_ERR_EXCHANGE 9
