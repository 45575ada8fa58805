// Microbenchmark of the inner loops of the fused center-site matvec (csrc/stage3f.cu) at the headline shape
// (D = 8: P = Q = R = S = 64, d = 2), without TMA, barriers or the op cursor: every warp repeats
//     T[4 tiles] = A_rows(8 x 64) * Vt      (first product,  256 DMMA.8x8x4)
//     acc[8][2] += T * B^T                  (second product, 256 DMMA.8x8x4)
// on operands that sit in shared memory, 8 warps per SM, one CTA per SM.  What is varied is only the ORDER in which the
// real DMMAs of the complex products are issued and how far ahead their operands are loaded -- to find out how much of
// the gap between the fused kernel (28 TFLOP/s) and the DMMA issue rate (37 TFLOP/s) is inherent to the instruction
// stream, and which stream closes it.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_loops scripts/dmma_loops.cu && ./dmma_loops
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef double2 cplx;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
struct CTile {
  double re0, re1, im0, im1;
  __device__ __forceinline__ void zero() { re0 = re1 = im0 = im1 = 0.0; }
};
__device__ __forceinline__ void cmma(CTile& acc, double ar, double ai, double nai, double br, double bi) {
  dmma(acc.re0, acc.re1, ar, br);
  dmma(acc.re0, acc.re1, nai, bi);
  dmma(acc.im0, acc.im1, ar, bi);
  dmma(acc.im0, acc.im1, ai, br);
}
__device__ __forceinline__ cplx lds_c(uint32_t addr) {
  cplx v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int NT = 4, NRT = 8, Q = 64, QS = 65, BSTR = 20, NPAIRS = 8;

// MODE 0: the order of stage3f.cu (tile by tile, the four DMMAs of a complex product back to back)
// MODE 1: same loads, DMMAs of a k-step issued "plane by plane" over the tiles (dependent DMMAs >= NT apart)
// MODE 2: MODE 1 + the operands of the next k-step loaded before the DMMAs of the current one (register double buffer)
// MODE 3: MODE 0 without any shared-memory load (operands stay in registers): the pure issue pattern
template <int MODE>
__global__ void __launch_bounds__(256, 1) loops_kernel(double* out, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  cplx* A = reinterpret_cast<cplx*>(smem);                     // [8 warps][8 rows][Q]
  cplx* Vt = A + 8 * 8 * Q;                                    // [8 NT rows][QS]
  cplx* B = Vt + 8 * NT * QS;                                  // [NRT 8 rows][BSTR]
  for (int i = tid; i < 8 * 8 * Q + 8 * NT * QS + NRT * 8 * BSTR; i += blockDim.x)
    A[i] = make_double2(1e-3 * (i % 7), -1e-3 * (i % 5));
  __syncthreads();
  const uint32_t a_base = smem_u32(A + warp * 8 * Q) + (uint32_t)((r * Q + 2 * c) * 16);
  const uint32_t v_base = smem_u32(Vt) + (uint32_t)((r * QS + 2 * c) * 16);
  const uint32_t tile_stride = 8 * QS * 16;
  const uint32_t b_base = smem_u32(B) + (uint32_t)((r * BSTR + c) * 16);
  const uint32_t rt_stride = 8 * BSTR * 16;
  CTile acc[NRT][2];
#pragma unroll
  for (int i = 0; i < NRT; ++i) {
    acc[i][0].zero();
    acc[i][1].zero();
  }
  for (int it = 0; it < iters; ++it) {
    CTile T[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) T[j].zero();
    // ---- first product
    if (MODE == 0 || MODE == 3) {
      cplx a0 = lds_c(a_base), a1 = lds_c(a_base + 16);
      cplx b0[NT], b1[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        b0[j] = lds_c(v_base + j * tile_stride);
        b1[j] = lds_c(v_base + j * tile_stride + 16);
      }
#pragma unroll 2
      for (int kp = 0; kp < NPAIRS; ++kp) {
        if (MODE == 0) {
          a0 = lds_c(a_base + kp * 128);
          a1 = lds_c(a_base + kp * 128 + 16);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (MODE == 0) {
            b0[j] = lds_c(v_base + j * tile_stride + kp * 128);
            b1[j] = lds_c(v_base + j * tile_stride + kp * 128 + 16);
          }
          cmma(T[j], a0.x, a0.y, -a0.y, b0[j].x, b0[j].y);
          cmma(T[j], a1.x, a1.y, -a1.y, b1[j].x, b1[j].y);
        }
      }
    } else {
      cplx a0 = lds_c(a_base), a1 = lds_c(a_base + 16);
      cplx b0[NT], b1[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        b0[j] = lds_c(v_base + j * tile_stride);
        b1[j] = lds_c(v_base + j * tile_stride + 16);
      }
#pragma unroll 2
      for (int kp = 0; kp < NPAIRS; ++kp) {
        cplx na0 = a0, na1 = a1, nb0[NT], nb1[NT];
        if (MODE == 1) {
          if (kp > 0) {
            a0 = lds_c(a_base + kp * 128);
            a1 = lds_c(a_base + kp * 128 + 16);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              b0[j] = lds_c(v_base + j * tile_stride + kp * 128);
              b1[j] = lds_c(v_base + j * tile_stride + kp * 128 + 16);
            }
          }
        } else {
          const int kn = kp + 1 < NPAIRS ? kp + 1 : kp;     // next k-step's operands, loaded ahead
          na0 = lds_c(a_base + kn * 128);
          na1 = lds_c(a_base + kn * 128 + 16);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            nb0[j] = lds_c(v_base + j * tile_stride + kn * 128);
            nb1[j] = lds_c(v_base + j * tile_stride + kn * 128 + 16);
          }
        }
        const double n0 = -a0.y, n1 = -a1.y;
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].re0, T[j].re1, a0.x, b0[j].x);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].im0, T[j].im1, a0.x, b0[j].y);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].re0, T[j].re1, n0, b0[j].y);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].im0, T[j].im1, a0.y, b0[j].x);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].re0, T[j].re1, a1.x, b1[j].x);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].im0, T[j].im1, a1.x, b1[j].y);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].re0, T[j].re1, n1, b1[j].y);
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(T[j].im0, T[j].im1, a1.y, b1[j].x);
        if (MODE == 2) {
          a0 = na0;
          a1 = na1;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            b0[j] = nb0[j];
            b1[j] = nb1[j];
          }
        }
      }
    }
    // ---- second product: acc[rt][s] += T[j] (as the A fragment) * B[8 rt .., 4 j ..]^T
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const CTile W = T[j];
      const double nim0 = -W.im0, nim1 = -W.im1;
      const uint32_t bj = b_base + (uint32_t)j * 64;
      if (MODE == 0 || MODE == 3) {
        cplx b = lds_c(bj);
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          cplx nb = b;
          if (MODE == 0 && rt + 1 < NRT) nb = lds_c(bj + (rt + 1) * rt_stride);
          cmma(acc[rt][0], W.re0, W.im0, nim0, b.x, b.y);
          cmma(acc[rt][1], W.re1, W.im1, nim1, b.x, b.y);
          b = nb;
        }
      } else {
        // two row tiles at a time, plane by plane: 4 independent accumulators between dependent DMMAs
        cplx b0 = lds_c(bj), b1 = lds_c(bj + rt_stride);
#pragma unroll
        for (int rt = 0; rt < NRT; rt += 2) {
          cplx nb0 = b0, nb1 = b1;
          if (rt + 2 < NRT) {
            nb0 = lds_c(bj + (rt + 2) * rt_stride);
            nb1 = lds_c(bj + (rt + 3) * rt_stride);
          }
          dmma(acc[rt][0].re0, acc[rt][0].re1, W.re0, b0.x);
          dmma(acc[rt][1].re0, acc[rt][1].re1, W.re1, b0.x);
          dmma(acc[rt + 1][0].re0, acc[rt + 1][0].re1, W.re0, b1.x);
          dmma(acc[rt + 1][1].re0, acc[rt + 1][1].re1, W.re1, b1.x);
          dmma(acc[rt][0].im0, acc[rt][0].im1, W.re0, b0.y);
          dmma(acc[rt][1].im0, acc[rt][1].im1, W.re1, b0.y);
          dmma(acc[rt + 1][0].im0, acc[rt + 1][0].im1, W.re0, b1.y);
          dmma(acc[rt + 1][1].im0, acc[rt + 1][1].im1, W.re1, b1.y);
          dmma(acc[rt][0].re0, acc[rt][0].re1, nim0, b0.y);
          dmma(acc[rt][1].re0, acc[rt][1].re1, nim1, b0.y);
          dmma(acc[rt + 1][0].re0, acc[rt + 1][0].re1, nim0, b1.y);
          dmma(acc[rt + 1][1].re0, acc[rt + 1][1].re1, nim1, b1.y);
          dmma(acc[rt][0].im0, acc[rt][0].im1, W.im0, b0.x);
          dmma(acc[rt][1].im0, acc[rt][1].im1, W.im1, b0.x);
          dmma(acc[rt + 1][0].im0, acc[rt + 1][0].im1, W.im0, b1.x);
          dmma(acc[rt + 1][1].im0, acc[rt + 1][1].im1, W.im1, b1.x);
          b0 = nb0;
          b1 = nb1;
        }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NRT; ++i) s += acc[i][0].re0 + acc[i][0].im1 + acc[i][1].re1 + acc[i][1].im0;
  if (s == 123.456) out[blockIdx.x * blockDim.x + tid] = s;
}

template <int MODE>
static void run(const char* what, int iters) {
  double* buf;
  cudaMalloc(&buf, sizeof(double) * 148 * 256);
  const size_t smem = sizeof(cplx) * (8 * 8 * Q + 8 * NT * QS + NRT * 8 * BSTR);
  cudaFuncSetAttribute(loops_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    loops_kernel<MODE><<<148, 256, smem>>>(buf, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double dmmas = 148.0 * 8 * iters * (double)(NT * NPAIRS * 8 + NT * NRT * 8);
  printf("mode %d  %-78s %7.2f TFLOP/s  (%s)\n", MODE, what, dmmas * 512.0 / (best * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(buf);
}

int main() {
  const int iters = 20000;
  run<3>("stage3f order, operands kept in registers (no shared-memory loads)", iters);
  run<0>("stage3f order: tile by tile, 4 DMMAs of a complex product back to back", iters);
  run<1>("plane by plane over the tiles (dependent DMMAs >= 4 apart)", iters);
  run<2>("plane by plane + next k-step's operands loaded ahead", iters);
  return 0;
}
