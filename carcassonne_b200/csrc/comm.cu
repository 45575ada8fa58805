// One-shot all-reduce of the center-site vector over NVLink peer memory.
//
// The matvec shards over the joined environment bond X: every GPU holds an X-slab of every stage-2 tensor and
// the full state vector, and produces a partial output vector of N = D^4 d complex numbers (128 KiB at D = 8).  The
// sum over ranks is latency-bound, so instead of a ring it is done in ONE kernel per rank:
//
//   1. every CTA reduces its chunk of the local partials (the fixed-order slot sum the single-GPU path does anyway)
//      into this rank's exchange buffer, which the peers have mapped through CUDA IPC;
//   2. the last CTA to finish publishes the epoch to a flag on every peer (system-scope release);
//   3. all CTAs wait until every peer has published the same epoch (system-scope acquire), then sum the N ranks'
//      exchange buffers for their chunk with plain peer loads over NVLink, in rank order -- so every rank obtains
//      bit-identical results, which the replicated Arnoldi iteration relies on.
//
// Exchange buffers are double-buffered by epoch parity, which makes a second barrier unnecessary: a rank can only
// reach epoch e+2 (and overwrite buffer e & 1) after every peer has signalled e+1, i.e. finished reading epoch e.
// Waits are bounded (~2 s of SM clocks) and set a sticky error flag instead of hanging the GPU.
#include <string.h>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

struct Comm {
  int rank = 0, world = 1;
  int64_t max_elems = 0;
  cplx* xbuf = nullptr;                 // local: [2][max_elems]
  unsigned long long* flags = nullptr;  // local: [world] epochs written by the peers, [world] = ticket, [world+1] = error
  cplx* peer_xbuf[16] = {nullptr};
  unsigned long long* peer_flags[16] = {nullptr};
  unsigned long long epoch = 0;
  bool connected = false;
};

namespace {

struct CommDev {
  int rank, world;
  int64_t max_elems;
  cplx* xbuf[16];
  unsigned long long* flags[16];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ cplx ld_peer(const cplx* p) {
  cplx v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// src: [slots][n] local partials (slots == 1 for a plain vector); out[n] = sum over ranks of sum over slots.
__global__ void __launch_bounds__(256) xgpu_allreduce_kernel(const cplx* __restrict__ src, int slots, int64_t n,
                                                            cplx* __restrict__ out, CommDev c,
                                                            unsigned long long epoch) {
  __shared__ bool last;
  __shared__ bool failed;
  cplx* mine = c.xbuf[c.rank] + (epoch & 1) * c.max_elems;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double re = 0.0, im = 0.0;
    for (int s = 0; s < slots; ++s) {
      const cplx v = src[(int64_t)s * n + i];
      re += v.x;
      im += v.y;
    }
    mine[i] = make_double2(re, im);
  }
  __threadfence_system();
  __syncthreads();
  unsigned long long* my_flags = c.flags[c.rank];
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(my_flags + c.world, 1ull);
    last = t == gridDim.x - 1;
    if (last) my_flags[c.world] = 0ull;   // every CTA of this launch has taken its ticket; the next launch starts at 0
  }
  __syncthreads();
  if (last) {
    __threadfence_system();   // order every CTA's exchange-buffer writes (observed through the ticket) before the flags
    if (threadIdx.x < c.world) st_release_sys(c.flags[threadIdx.x] + c.rank, epoch);
  }
  if (threadIdx.x == 0) {
    bool bad = false;
    const long long t0 = clock64();
    for (int r = 0; r < c.world; ++r) {
      while (ld_acquire_sys(my_flags + r) < epoch) {
        if (clock64() - t0 > 4000000000ll) {
          bad = true;
          break;
        }
      }
    }
    if (bad) my_flags[c.world + 1] = 1ull;
    failed = bad;
  }
  __syncthreads();
  if (failed) {
    // never hand back uninitialised memory: a NaN result trips every downstream guard (hasNaN, the solver's
    // non-finite Ritz check) and the host reads the sticky flag at its next synchronisation point
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_double2(nan, nan);
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double re = 0.0, im = 0.0;
    for (int r = 0; r < c.world; ++r) {
      const cplx v = ld_peer(c.xbuf[r] + (epoch & 1) * c.max_elems + i);
      re += v.x;
      im += v.y;
    }
    out[i] = make_double2(re, im);
  }
}

}  // namespace

int comm_create(Comm** out, int rank, int world, int64_t max_elems) {
  CARC_REQUIRE(out && world >= 1 && world <= 16 && rank >= 0 && rank < world && max_elems > 0, CARC_ERR_VALUE,
               "comm_create: invalid arguments (rank %d, world %d)", rank, world);
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  c->max_elems = max_elems;
  CARC_CHECK_CUDA(cudaMalloc(&c->xbuf, sizeof(cplx) * 2 * (size_t)max_elems));
  CARC_CHECK_CUDA(cudaMalloc(&c->flags, sizeof(unsigned long long) * (world + 2)));
  CARC_CHECK_CUDA(cudaMemset(c->flags, 0, sizeof(unsigned long long) * (world + 2)));
  CARC_CHECK_CUDA(cudaDeviceSynchronize());
  c->peer_xbuf[rank] = c->xbuf;
  c->peer_flags[rank] = c->flags;
  c->connected = world == 1;
  *out = c;
  return CARC_OK;
}

int comm_local_handles(Comm* c, void* out128) {
  CARC_REQUIRE(c && out128, CARC_ERR_VALUE, "comm_local_handles: null argument");
  cudaIpcMemHandle_t h[2];
  CARC_CHECK_CUDA(cudaIpcGetMemHandle(&h[0], c->xbuf));
  CARC_CHECK_CUDA(cudaIpcGetMemHandle(&h[1], c->flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out128, h, 128);
  return CARC_OK;
}

int comm_connect(Comm* c, const void* all_handles) {
  CARC_REQUIRE(c && all_handles, CARC_ERR_VALUE, "comm_connect: null argument");
  const char* bytes = static_cast<const char*>(all_handles);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h[2];
    memcpy(h, bytes + (size_t)r * 128, 128);
    CARC_CHECK_CUDA(cudaIpcOpenMemHandle((void**)&c->peer_xbuf[r], h[0], cudaIpcMemLazyEnablePeerAccess));
    CARC_CHECK_CUDA(cudaIpcOpenMemHandle((void**)&c->peer_flags[r], h[1], cudaIpcMemLazyEnablePeerAccess));
  }
  c->connected = true;
  return CARC_OK;
}

int comm_allreduce(Comm* c, const cplx* src, int slots, int64_t n, cplx* out, cudaStream_t stream) {
  CARC_REQUIRE(c && c->connected, CARC_ERR_VALUE, "comm_allreduce: communicator not connected");
  CARC_REQUIRE(n <= c->max_elems, CARC_ERR_VALUE, "comm_allreduce: %lld elements exceed the exchange buffer (%lld)",
               (long long)n, (long long)c->max_elems);
  CommDev d;
  d.rank = c->rank;
  d.world = c->world;
  d.max_elems = c->max_elems;
  for (int r = 0; r < 16; ++r) {
    d.xbuf[r] = c->peer_xbuf[r];
    d.flags[r] = c->peer_flags[r];
  }
  ++c->epoch;
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count()) blocks = sm_count();   // all CTAs must be co-resident: they wait on each other through the ticket
  xgpu_allreduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, slots, n, out, d, c->epoch);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

int comm_status(Comm* c, int* timed_out) {
  CARC_REQUIRE(c && timed_out, CARC_ERR_VALUE, "comm_status: null argument");
  unsigned long long v = 0;
  CARC_CHECK_CUDA(cudaMemcpy(&v, c->flags + c->world + 1, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = v != 0;
  return CARC_OK;
}

int comm_destroy(Comm* c) {
  if (!c) return CARC_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    if (c->peer_xbuf[r]) cudaIpcCloseMemHandle(c->peer_xbuf[r]);
    if (c->peer_flags[r]) cudaIpcCloseMemHandle(c->peer_flags[r]);
  }
  cudaFree(c->xbuf);
  cudaFree(c->flags);
  delete c;
  return CARC_OK;
}

}  // namespace carc
