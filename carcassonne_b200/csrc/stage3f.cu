// Fused center-site matvec, "folded" tiling (d = 2): the variant of stage3.cu for bond dimensions whose squares are
// not multiples of 8 (D = 3, 5, 6, 7, ragged shapes).
//
//   out[P,R,s] = sum_t sum_x sum_S B_t[x,R,S] * ( sum_s' O_t[s,s'] * ( sum_Q A_t[x,P,Q] * v[Q,S,s'] ) )
//
// (reference tensors/_2d/dense.py:115-160 and tensors/_2d/sparse.py:129-133, as in stage3.cu).  Same structure as
// stage3_kernel -- TMA rings for A_x and the S block of B_x, stars of terms, C fragment of the first product reused as
// the A fragment of the second, partial results summed in a fixed order -- but the spin index is folded into the
// column dimension of the first product: a tile's 8 columns are (S = 4j .. 4j+3) x (s = 0, 1), so
//   * the first product's N extent is 2*S rounded up to 8 (72 = 9 tiles at D = 6, where stage3_kernel pads
//     2 x 40), and its K loop takes Q in steps of 4 (one trailing single DMMA step when ceil(Q/4) is odd) instead of 8;
//   * a lane's C fragment holds the two spins of ONE S column (slot 0 = spin 0, slot 1 = spin 1), so the site
//     operator acts inside the lane, and the second product needs one B fragment, B_x[R = 8 rt + r, S = 4 j + c],
//     for both spins; its K extent is S rounded up to 4 instead of 16;
//   * S blocks hold a whole number of tiles, spread as evenly as the shape allows (D = 7: 5 + 4 + 4 tiles), and the
//     CTAs are shared out between the S blocks in proportion to their tile counts.
// Warps are symmetric (no producer warp, and -- unlike stage3_kernel -- no producer role for warp 0 either: measured
// there, the warp that issued all copies never waited while every other warp spent ~30 % of its time waiting for it):
//   * a warp only ever reads its own 8 rows of A_x, so every warp owns a private TMA ring for those rows (its own
//     full barriers, no empty barriers: it refills a slot itself after reading it);
//   * the S block of B_x is shared by the warps of a group; each warp issues the row copies of its share of the rows
//     (full barrier: one expect-tx arrival per warp) after waiting for the slot's empty barrier;
//   * every warp runs the same producer cursor over the flattened op sequence, ahead of its consumer cursor.
// DMMA work relative to stage3_kernel: 0.82 (D = 5), 0.86 (D = 6), 0.78 (D = 7), 0.66 (D = 3).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

namespace {

constexpr int S3F_SMEM_LIMIT = 227 * 1024;
constexpr int S3F_MAX_SB = 16;

struct S3FParams {
  const Stage3Term* terms;  // device copy, sorted by group
  const Stage3Group* groups;
  int nterms, ngroups;
  int P, Q, R, S;                 // P, R: extents of the (row block, column block) of the output this launch computes
  int Pfull, Rfull, p0, r0;       // the whole output is Pfull x Rfull; the block starts at row p0, column r0
  int Q4, NPT, G, NSB, nstA, nstB, QS, BSTR;
  int psleep;                     // ns a producer lane sleeps between two polls of an empty barrier
  int b_whole;                    // one S block: B_x is one contiguous copy, issued by the warps in turn
  int sb_cta0[S3F_MAX_SB + 1];    // S block i has sb_cta0[i+1] - sb_cta0[i] CTAs (one X slab each)
  unsigned char cta_sb[160], cta_sl[160];   // CTA -> (S block, X slab): CTAs sharing a stretch of X are neighbours
  unsigned char cta_map[160];               // blockIdx.x of this launch -> CTA (launches are per S-block width)
  int sb_tile0[S3F_MAX_SB + 1];   // S block i covers the tiles (4 S columns each) [sb_tile0[i], sb_tile0[i+1])
  uint32_t slotA_bytes, slotB_bytes, ops_off, hasop_off, tab_off, vt_off, vtail_off, ring_off, smem_total;
  const cplx* v;
  cplx* partial;
};

// 2 / 3 warps per SM sub-partition leave 255 / 170 registers per thread (see stage3.cu)
__host__ __device__ constexpr int s3f_max_threads(int nrt) { return nrt >= 7 ? 256 : 384; }
// tiles per S block (register budget: 16 NRT accumulator registers + 8 per T tile + 16 for the U pair)
__host__ __device__ constexpr int s3f_tiles(int nrt) { return nrt <= 4 ? 4 : nrt <= 6 ? 3 : nrt == 7 ? 5 : 4; }

// Optional wait-time instrumentation (make EXTRA=-DS3F_PROFILE): cycles each warp spends in the three consumer-side
// mbarrier waits and in total, read back through stage3f_profile_read().
#ifdef S3F_PROFILE
__device__ unsigned long long s3f_prof[148 * 12 * 4];
#define S3F_WAIT(slot_, bar_, parity_)             \
  do {                                             \
    const long long t0_ = clock64();               \
    mbar_wait(bar_, parity_);                      \
    prof_wait[slot_] += clock64() - t0_;           \
  } while (0)
#else
#define S3F_WAIT(slot_, bar_, parity_) mbar_wait(bar_, parity_)
#endif

struct S3FLane {
  uint32_t a_pair, a_tail, v_pair, v_tail, tile_stride, tail_stride, a_first, a_second;
  int npairs, ntv;
  bool tail, eswap;
};

// Is tile j of the S block active?  NA > 0: the block holds exactly NA tiles, known at compile time; NA == 0: it holds
// L.ntv tiles, a run-time number.  This distinction is what the kernel's speed hangs on: a run-time test around the
// DMMAs of a tile is a (possibly divergent) branch, after which mma.sync needs a WARPSYNC, and it pins the tile's
// operand loads directly in front of their first use -- the SASS was [BRA, WARPSYNC, LDS, LDS, 8 DMMA] per tile, every
// load latency exposed, 76 % DMMA pipe utilisation at D = 8.  With the tile count a template parameter the same loop
// is straight-line code that ptxas software-pipelines (scripts/dmma_loops.cu: these loops alone reach 37.1 TFLOP/s,
// the issue-rate peak).
template <int NA>
__device__ __forceinline__ bool s3f_active(int j, const S3FLane& L) {
  return NA > 0 ? j < NA : j < L.ntv;
}

// One paired k-step (8 values of Q) of the first product for the active tiles.
template <int NJ, int NA>
__device__ __forceinline__ void s3f_first_step(CTile (&T)[NJ], int jbase, const S3FLane& L, uint32_t a_base, int kp) {
  const cplx x0 = lds_c(a_base + kp * 128 + L.a_first);
  const cplx x1 = lds_c(a_base + kp * 128 + L.a_second);
  const cplx a0 = L.eswap ? x1 : x0;
  const cplx a1 = L.eswap ? x0 : x1;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    if (s3f_active<NA>(jbase + jj, L)) {
      const uint32_t va = L.v_pair + (uint32_t)(jbase + jj) * L.tile_stride + kp * 128;
      const cplx b0 = lds_c(va);
      const cplx b1 = lds_c(va + 16);
      cmma(T[jj], a0.x, a0.y, -a0.y, b0.x, b0.y);
      cmma(T[jj], a1.x, a1.y, -a1.y, b1.x, b1.y);
    }
  }
}

// T[jj] += A_x (8 rows of this warp) * Vt (tile jbase + jj), for the active tiles.  (Measured and not adopted: a k loop
// unrolled completely for the uniform bond dimensions -- ptxas hoists more loads than the register file holds and spills:
// 27.98 vs 30.14 TFLOP/s at D = 8, 17.99 vs 19.46 at D = 7.)
template <int NJ, int NA>
__device__ __forceinline__ void s3f_first(CTile (&T)[NJ], int jbase, const S3FLane& L, uint32_t slot) {
  const uint32_t a_base = slot + L.a_pair;
#pragma unroll 2
  for (int kp = 0; kp < L.npairs; ++kp) s3f_first_step<NJ, NA>(T, jbase, L, a_base, kp);
  if (L.tail) {
    const cplx a = lds_c(slot + L.a_tail);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      if (s3f_active<NA>(jbase + jj, L)) {
        const cplx b = lds_c(L.v_tail + (uint32_t)(jbase + jj) * L.tail_stride);
        cmma(T[jj], a.x, a.y, -a.y, b.x, b.y);
      }
    }
  }
}

// out (+)= O * U on the two spins a lane holds (slot 0 = spin 0, slot 1 = spin 1); w = O row-major [s][s']
__device__ __forceinline__ void s3f_apply_op(CTile& out, const CTile& U, const cplx* w) {
  const cplx w00 = w[0], w01 = w[1], w10 = w[2], w11 = w[3];
  if (w00.x != 0.0 || w00.y != 0.0) {
    out.re0 += w00.x * U.re0 - w00.y * U.im0;
    out.im0 += w00.x * U.im0 + w00.y * U.re0;
  }
  if (w01.x != 0.0 || w01.y != 0.0) {
    out.re0 += w01.x * U.re1 - w01.y * U.im1;
    out.im0 += w01.x * U.im1 + w01.y * U.re1;
  }
  if (w10.x != 0.0 || w10.y != 0.0) {
    out.re1 += w10.x * U.re0 - w10.y * U.im0;
    out.im1 += w10.x * U.im0 + w10.y * U.re0;
  }
  if (w11.x != 0.0 || w11.y != 0.0) {
    out.re1 += w11.x * U.re1 - w11.y * U.im1;
    out.im1 += w11.x * U.im1 + w11.y * U.re1;
  }
}

// acc[rt][s] += W (tile j, spin s) * B_x[8 rt .., 4 j ..]^T : W's C fragment is the A fragment, one B fragment per rt
// (loaded one row tile ahead of the DMMAs that use it)
template <int NRT>
__device__ __forceinline__ void s3f_second(CTile (&acc)[NRT][2], const CTile& W, int j, uint32_t b_base, uint32_t rt_stride) {
  const double nim0 = -W.im0, nim1 = -W.im1;
  const uint32_t bj = b_base + (uint32_t)j * 64;
  cplx b = lds_c(bj);
#pragma unroll
  for (int rt = 0; rt < NRT; ++rt) {
    cplx nb = b;
    if (rt + 1 < NRT) nb = lds_c(bj + (rt + 1) * rt_stride);
    cmma(acc[rt][0], W.re0, W.im0, nim0, b.x, b.y);
    cmma(acc[rt][1], W.re1, W.im1, nim1, b.x, b.y);
    b = nb;
  }
}

// NA: number of tiles the S blocks of THIS launch hold (NT or NT - 1: an uneven split of S gives blocks of both widths
// and one launch per width), or 0 for "read it at run time" (see s3f_active).
// WS > 0 ("warp specialised"): the CTA has WS consumer-warp slots (8 or 12: two or three warpgroups) and one more
// warpgroup of producers that runs the op cursor and issues every TMA copy, so that the consumer warps do nothing but
// wait -> DMMA -> release.  `setmaxnreg` moves the registers to where the accumulators are: the kernel is launched with
// (WS + 4) warps (168 / 128 registers each), the producer warpgroup drops to 40 and the consumer warpgroups rise to
// 232 / 152 (8 x 32 x 232 + 4 x 32 x 40 = 64 512, 12 x 32 x 152 + 4 x 32 x 40 = 63 488 of the 65 536).  A slots then need
// empty barriers (consumer -> producer).  The 4 producer warps are shared out between the G <= 4 groups: 4 / G warps per
// group, each issuing the A copies of every (4 / G)-th consumer warp and every (4 / G)-th row of the B blocks.
// WS == 0: the symmetric kernel -- every warp is consumer and producer of its own rows (kept for CARC_S3F_WS=0).
__host__ __device__ constexpr int s3f_ws_consumer_registers(int ws) { return ws == 8 ? 232 : 152; }
__host__ __device__ constexpr int s3f_producers_per_group(int G) { return G == 1 ? 4 : G == 2 ? 2 : 1; }
template <int NRT, int NT, int NA, int WS>
__global__ void __launch_bounds__(WS ? (WS + 4) * 32 : s3f_max_threads(NRT), 1) stage3f_kernel(const S3FParams p) {
  const int cta = p.cta_map[blockIdx.x];   // position in the whole job's CTA table (slab, S block, partial slot)
  constexpr int DP = 2;
  constexpr int UW = NT < 2 ? NT : 2;   // tiles per pass of a later term's first product (register budget)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int ncw = p.G * p.NPT;
  const int nbar_g = p.NPT * p.nstA + 2 * p.nstB;   // per group: full A per (warp, stage), full B and empty B per stage
  const uint32_t bars = smem_u32(smem);
  cplx* ops = reinterpret_cast<cplx*>(smem + p.ops_off);
  cplx* Vt = reinterpret_cast<cplx*>(smem + p.vt_off);
  cplx* VtTail = reinterpret_cast<cplx*>(smem + p.vtail_off);
  const int QS = p.QS, BSTR = p.BSTR;
  const uint32_t group_bytes = p.NPT * p.nstA * p.slotA_bytes + p.nstB * p.slotB_bytes;

  // zero everything behind the barriers: padding rows / columns must read as finite zeros forever
  {
    uint4* z = reinterpret_cast<uint4*>(smem + p.ops_off);
    const uint32_t n16 = (p.smem_total - p.ops_off) / 16;
    for (uint32_t i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int g = 0; g < p.G; ++g) {
      const uint32_t b = bars + g * nbar_g * 8;
      for (int i = 0; i < p.NPT * p.nstA; ++i) mbar_init(b + i * 8, 1);   // full A, private to one warp
      for (int i = 0; i < p.nstB; ++i) {
        // full B: every warp that copies rows announces its share (WS: the four producer warps)
        mbar_init(b + (p.NPT * p.nstA + 2 * i) * 8, p.b_whole ? 1 : (WS ? s3f_producers_per_group(p.G) : p.NPT));
        mbar_init(b + (p.NPT * p.nstA + 2 * i + 1) * 8, p.NPT);   // empty B
      }
    }
    if (WS)   // empty A, one per (consumer warp, stage), behind the other barriers
      for (int i = 0; i < ncw * p.nstA; ++i) mbar_init(bars + (p.G * nbar_g + i) * 8, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int sb = p.cta_sb[cta], sl = p.cta_sl[cta];
  const int NSL = p.sb_cta0[sb + 1] - p.sb_cta0[sb];
  const int S0 = 4 * p.sb_tile0[sb];
  const int SBv = min(4 * (p.sb_tile0[sb + 1] - p.sb_tile0[sb]), p.S - S0);
  int* hasop = reinterpret_cast<int*>(smem + p.hasop_off);
  const cplx** termA = reinterpret_cast<const cplx**>(smem + p.tab_off);
  const cplx** termB = termA + p.nterms;
  const cplx** groupCenter = termB + p.nterms;
  int* groupX = reinterpret_cast<int*>(groupCenter + p.ngroups);
  int* groupFirst = groupX + p.ngroups;
  int* groupCount = groupFirst + p.ngroups;
  int* groupKind = groupCount + p.ngroups;
  for (int i = tid; i < p.nterms * DP * DP; i += blockDim.x) ops[i] = p.terms[i / (DP * DP)].op[i % (DP * DP)];
  for (int i = tid; i < p.nterms; i += blockDim.x) {
    hasop[i] = p.terms[i].has_op;
    termA[i] = p.terms[i].A;
    termB[i] = p.terms[i].B;
  }
  for (int i = tid; i < p.ngroups; i += blockDim.x) {
    groupCenter[i] = p.groups[i].center;
    groupKind[i] = p.groups[i].kind;
    groupX[i] = (int)p.groups[i].X;
    groupFirst[i] = p.groups[i].first;
    groupCount[i] = p.groups[i].count;
  }
  // Vt[f = 2 * S_local + s][q] = v[q, S0 + S_local, s]
  // (the trailing single k-step, q >= 8 * (Q4 / 2), lives in VtTail[f][4] so that its loads are conflict-free too)
  for (int i = tid; i < p.Q * SBv * DP; i += blockDim.x) {
    const int f = i % (DP * SBv), q = i / (DP * SBv);
    const cplx val = p.v[((int64_t)q * p.S + S0) * DP + f];
    const int qt = q - 8 * (p.Q4 >> 1);
    if (qt < 0) Vt[f * QS + q] = val;
    else VtTail[f * 4 + qt] = val;
  }
  fence_proxy_async();
  __syncthreads();

  if (WS && warp >= WS) {
    // ---- producer warpgroup: the flattened op sequence of one group's share of this CTA's slab, every copy of it
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    const int ppg = s3f_producers_per_group(p.G);
    const int g = (warp - WS) / ppg, pl = (warp - WS) % ppg;            // group served; which of its ppg producers
    if (g >= p.G) return;
    const uint32_t barsG = bars + g * nbar_g * 8;                       // the group's full-A barriers [warp][stage]
    const uint32_t ringA = smem_u32(smem + p.ring_off) + g * group_bytes;
    const uint32_t ringB = ringA + p.NPT * p.nstA * p.slotA_bytes;
    const uint32_t bB = barsG + (p.NPT * p.nstA) * 8;
    const uint32_t bEA = bars + (p.G * nbar_g + g * p.NPT * p.nstA) * 8;
    const uint32_t b_row_bytes = (uint32_t)(SBv * 16);
    const int my_rows = p.R > pl ? (p.R - pl + ppg - 1) / ppg : 0;      // rows pl, pl + ppg, ... of every B block
    // lane l issues the A copies of consumer warp l of the group (if that warp is this producer's)
    const int a_rows = (lane < p.NPT && lane % ppg == pl) ? min(8, p.P - 8 * lane) : 0;
    int gi = -1, j = 0, x = 0, x_hi = 0;
    int slotA = 0, parA = 1, slotB = 0, parB = 1, turn = 0;
    for (;;) {
      // advance to the next op (same order as the consumers' cursor)
      if (gi >= 0 && j < groupCount[gi]) {
        ++j;
      } else {
        j = 0;
        x += p.G;
        bool done = false;
        while (gi < 0 || x >= x_hi) {
          if (++gi >= p.ngroups) {
            done = true;
            break;
          }
          const int64_t X = groupX[gi];
          x = (int)(X * sl / NSL) + g;
          x_hi = (int)(X * (sl + 1) / NSL);
        }
        if (done) break;
      }
      const int kind = groupKind[gi], first = groupFirst[gi];
      if (kind ? (j == 0) : (j < groupCount[gi])) {
        const cplx* A = kind ? groupCenter[gi] : termA[first + j];
        if (a_rows > 0) {
          const uint32_t full = barsG + (lane * p.nstA + slotA) * 8, bytes = (uint32_t)(a_rows * p.Q * 16);
          while (!mbar_try_wait(bEA + (lane * p.nstA + slotA) * 8, (uint32_t)parA))
            if (p.psleep) __nanosleep(p.psleep);
          mbar_arrive_expect_tx(full, bytes);
          bulk_g2s(ringA + (lane * p.nstA + slotA) * p.slotA_bytes, A + ((int64_t)x * p.Pfull + p.p0 + 8 * lane) * p.Q, bytes, full);
        }
        __syncwarp();
        if (++slotA == p.nstA) {
          slotA = 0;
          parA ^= 1;
        }
      } else {
        const cplx* B = kind ? termB[first + j - 1] : groupCenter[gi];
        const uint32_t fullB = bB + (2 * slotB) * 8, dst = ringB + slotB * p.slotB_bytes;
        if (p.b_whole) {
          if (turn == pl && lane == 0) {
            while (!mbar_try_wait(fullB + 8, (uint32_t)parB))
              if (p.psleep) __nanosleep(p.psleep);
            mbar_arrive_expect_tx(fullB, (uint32_t)(p.R * p.S * 16));
            bulk_g2s(dst, B + ((int64_t)x * p.Rfull + p.r0) * p.S, (uint32_t)(p.R * p.S * 16), fullB);
          }
        } else {
          if (lane == 0) {
            while (!mbar_try_wait(fullB + 8, (uint32_t)parB))
              if (p.psleep) __nanosleep(p.psleep);
            if (my_rows > 0) mbar_arrive_expect_tx(fullB, b_row_bytes * my_rows);
            else mbar_arrive(fullB);
          }
          __syncwarp();
          const cplx* src = B + ((int64_t)x * p.Rfull + p.r0) * p.S + S0;
          for (int rr = pl + ppg * lane; rr < p.R; rr += 32 * ppg) bulk_g2s(dst + rr * BSTR * 16, src + (int64_t)rr * p.S, b_row_bytes, fullB);
        }
        if (++slotB == p.nstB) {
          slotB = 0;
          parB ^= 1;
        }
        if (++turn == ppg) turn = 0;
      }
    }
    return;
  }
  if (WS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(s3f_ws_consumer_registers(WS)));
  if (warp >= ncw) return;
  {
    const int g = warp / p.NPT, wg = warp % p.NPT;
    const uint32_t bA = bars + (g * nbar_g + wg * p.nstA) * 8;          // this warp's full-A barriers
    const uint32_t bB = bars + (g * nbar_g + p.NPT * p.nstA) * 8;       // the group's (full B, empty B) pairs
    const uint32_t ringA = smem_u32(smem + p.ring_off) + g * group_bytes + wg * p.nstA * p.slotA_bytes;
    const uint32_t ringB = smem_u32(smem + p.ring_off) + g * group_bytes + p.NPT * p.nstA * p.slotA_bytes;
    CTile acc[NRT][DP];
#pragma unroll
    for (int i = 0; i < NRT; ++i)
#pragma unroll
      for (int s = 0; s < DP; ++s) acc[i][s].zero();

#ifdef S3F_PROFILE
    long long prof_wait[3] = {0, 0, 0};
    const long long prof_t0 = clock64();
#endif
    S3FLane L;
    L.npairs = p.Q4 >> 1;
    L.tail = (p.Q4 & 1) != 0;
    L.ntv = (SBv + 3) >> 2;
    L.eswap = ((p.Q & 1) == 0) && (r & 1);
    L.a_first = L.eswap ? 16u : 0u;
    L.a_second = 16u - L.a_first;
    L.a_pair = (uint32_t)((r * p.Q + 2 * c) * 16);
    L.a_tail = (uint32_t)((r * p.Q + 8 * L.npairs + c) * 16);
    L.v_pair = smem_u32(Vt) + (uint32_t)((r * QS + 2 * c) * 16);
    L.v_tail = smem_u32(VtTail) + (uint32_t)((r * 4 + c) * 16);
    L.tile_stride = (uint32_t)(8 * QS * 16);
    L.tail_stride = (uint32_t)(8 * 4 * 16);
    const uint32_t b_lane_off = (uint32_t)((r * BSTR + c) * 16);
    const uint32_t rt_stride = (uint32_t)(8 * BSTR * 16);

    // Op cursors: identical to stage3_kernel (flattened sequence of [A-op ..., B-op] per star and x).
    struct Cursor {
      int gi, j, x, x_hi;
    };
    auto next_item = [&](Cursor& cu) -> bool {
      cu.x += p.G;
      while (cu.gi < 0 || cu.x >= cu.x_hi) {
        if (++cu.gi >= p.ngroups) return false;
        const int64_t X = groupX[cu.gi];
        cu.x = (int)(X * sl / NSL) + g;
        cu.x_hi = (int)(X * (sl + 1) / NSL);
      }
      return true;
    };
    auto advance = [&](Cursor& cu) -> bool {
      if (cu.gi >= 0 && cu.j < groupCount[cu.gi]) {
        ++cu.j;
        return true;
      }
      cu.j = 0;
      return next_item(cu);
    };
    // this warp's 8 rows of A_x (fewer in the last row tile: the rest of the slot stays zero)
    const uint32_t a_bytes = (uint32_t)(min(8, p.P - 8 * wg) * p.Q * 16);
    const uint32_t b_row_bytes = (uint32_t)(SBv * 16);
    // rows wg, wg + NPT, ... of the B block are this warp's to copy
    const int my_rows = p.R > wg ? (p.R - wg + p.NPT - 1) / p.NPT : 0;
    // ring positions are kept as (slot, parity) counters: the stage counts are run-time values and a modulo / division
    // per op costs as much as a dozen DMMAs' worth of issue slots between two ops
    int pA_slot = 0, pB_slot = 0, pB_par = 1, pB_turn = 0;      // producer side (pB_par: parity of the EMPTY barrier)
    int cA_slot = 0, cA_par = 0, cB_slot = 0, cB_par = 0;       // consumer side
    auto issueA = [&](const Cursor& cu) {
      const cplx* A = groupKind[cu.gi] ? groupCenter[cu.gi] : termA[groupFirst[cu.gi] + cu.j];
      const int sa = pA_slot;
      if (++pA_slot == p.nstA) pA_slot = 0;
      if (lane == 0) {
        // the slot was read by this warp's own lanes (all past the __syncwarp that follows every first product)
        fence_proxy_async();
        mbar_arrive_expect_tx(bA + sa * 8, a_bytes);
        bulk_g2s(ringA + sa * p.slotA_bytes, A + ((int64_t)cu.x * p.Pfull + p.p0 + 8 * wg) * p.Q, a_bytes, bA + sa * 8);
      }
    };
    auto issueB = [&](const Cursor& cu) {
      const cplx* B = groupKind[cu.gi] ? termB[groupFirst[cu.gi] + cu.j - 1] : groupCenter[cu.gi];
      const int sbq = pB_slot;
      const uint32_t empty_parity = (uint32_t)pB_par;
      const bool my_turn = pB_turn == wg;
      if (++pB_slot == p.nstB) {
        pB_slot = 0;
        pB_par ^= 1;
      }
      if (++pB_turn == p.NPT) pB_turn = 0;
      const uint32_t fullB = bB + (2 * sbq) * 8;
      const uint32_t dst = ringB + sbq * p.slotB_bytes;
      if (p.b_whole) {
        // the S block is all of S: B_x is contiguous (row stride S), one copy, issued by the group's warps in turn
        if (my_turn && lane == 0) {
          mbar_wait(fullB + 8, empty_parity);
          mbar_arrive_expect_tx(fullB, (uint32_t)(p.R * p.S * 16));
          bulk_g2s(dst, B + ((int64_t)cu.x * p.Rfull + p.r0) * p.S, (uint32_t)(p.R * p.S * 16), fullB);
        }
        return;
      }
      if (lane == 0) {
        mbar_wait(fullB + 8, empty_parity);
        if (my_rows > 0) mbar_arrive_expect_tx(fullB, b_row_bytes * my_rows);
        else mbar_arrive(fullB);
      }
      __syncwarp();
      const cplx* src = B + ((int64_t)cu.x * p.Rfull + p.r0) * p.S + S0;
      for (int rr = wg + p.NPT * lane; rr < p.R; rr += 32 * p.NPT)
        bulk_g2s(dst + rr * BSTR * 16, src + (int64_t)rr * p.S, b_row_bytes, fullB);
    };

    Cursor cc = {-1, 0, 0, 0}, pc = {-1, 0, 0, 0};
    uint32_t issuedA = 0, issuedB = 0, itA = 0, itB = 0;
    bool more = true, pending = false;
    // (Measured and not adopted: starting the warps that share an SM sub-partition a fraction of an op apart, so that one
    // warp's bookkeeping between two ops falls under the other's DMMAs -- no change at D = 4, 6, 8, before and after the
    // tile loops lost their predicates: the idle DMMA slots are not a phase-locking effect.)
    const uint32_t bEA = bars + (p.G * nbar_g + (g * p.NPT + wg) * p.nstA) * 8;     // WS: this warp's empty-A barriers
    auto run_ahead = [&]() {
      if (WS) return;                   // the producer warpgroup issues the copies
      while (more) {
        if (!pending) {
          more = advance(pc);
          pending = more;
          if (!more) break;
        }
        if (groupKind[pc.gi] ? (pc.j == 0) : (pc.j < groupCount[pc.gi])) {
          if (issuedA - itA >= (uint32_t)p.nstA) break;
          issueA(pc);
          ++issuedA;
        } else {
          if (issuedB - itB >= (uint32_t)p.nstB) break;
          issueB(pc);
          ++issuedB;
        }
        pending = false;
      }
    };

    while (next_item(cc)) {
      CTile T[NT];
      const int first = groupFirst[cc.gi], cnt = groupCount[cc.gi];
      const int kind = groupKind[cc.gi];
      {
        // ---- first term of the star: all tiles of the S block accumulate straight into T
        run_ahead();
        const bool has_op = !kind && hasop[first] != 0;   // an A-star keeps the raw product
        const int slot = cA_slot;
        S3F_WAIT(0, bA + slot * 8, cA_par);
#pragma unroll
        for (int j = 0; j < NT; ++j) T[j].zero();
        s3f_first<NT, NA>(T, 0, L, ringA + slot * p.slotA_bytes);
        if (has_op) {
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (NA == 0 || j < NA) {
              const CTile U = T[j];
              T[j].zero();
              s3f_apply_op(T[j], U, ops + first * DP * DP);
            }
          }
        }
        __syncwarp();
        if (WS && lane == 0) mbar_arrive(bEA + slot * 8);
        ++itA;
        if (++cA_slot == p.nstA) {
          cA_slot = 0;
          cA_par ^= 1;
        }
      }
      if (!kind) {
        for (int jt = 1; jt < cnt; ++jt) {
          run_ahead();
          // ---- a further term of a B-star: T += O (A_x v); without a site operator the DMMA chains simply continue
          const int term = first + jt;
          const bool has_op = hasop[term] != 0;
          const int slot = cA_slot;
          S3F_WAIT(1, bA + slot * 8, cA_par);
          const uint32_t aslot = ringA + slot * p.slotA_bytes;
          if (!has_op) {
            s3f_first<NT, NA>(T, 0, L, aslot);
          } else {
#pragma unroll
            for (int j0 = 0; j0 < NT; j0 += UW) {
              if (NA == 0 || j0 < NA) {
                CTile U[UW];
#pragma unroll
                for (int jj = 0; jj < UW; ++jj) U[jj].zero();
                s3f_first<UW, NA>(U, j0, L, aslot);
#pragma unroll
                for (int jj = 0; jj < UW; ++jj)
                  if (j0 + jj < NT && (NA == 0 || j0 + jj < NA)) s3f_apply_op(T[j0 + jj], U[jj], ops + term * DP * DP);
              }
            }
          }
          __syncwarp();
          if (WS && lane == 0) mbar_arrive(bEA + slot * 8);
          ++itA;
          if (++cA_slot == p.nstA) {
            cA_slot = 0;
            cA_par ^= 1;
          }
        }
        run_ahead();
        // ---- second product of the star: acc += T * B_x^T
        {
          const int slot = cB_slot;
          S3F_WAIT(2, bB + (2 * slot) * 8, cB_par);
          const uint32_t b_base = ringB + slot * p.slotB_bytes + b_lane_off;
#pragma unroll
          for (int j = 0; j < NT; ++j)
            if (s3f_active<NA>(j, L)) s3f_second<NRT>(acc, T[j], j, b_base, rt_stride);
          __syncwarp();
          if (lane == 0) mbar_arrive(bB + (2 * slot + 1) * 8);
          ++itB;
          if (++cB_slot == p.nstB) {
            cB_slot = 0;
            cB_par ^= 1;
          }
        }
      } else {
        // ---- A-star: T holds A_x v once; every term applies its own site operator and multiplies with its own B_x
        for (int jt = 0; jt < cnt; ++jt) {
          run_ahead();
          const int term = first + jt;
          const bool has_op = hasop[term] != 0;
          const int slot = cB_slot;
          S3F_WAIT(2, bB + (2 * slot) * 8, cB_par);
          const uint32_t b_base = ringB + slot * p.slotB_bytes + b_lane_off;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (s3f_active<NA>(j, L)) {
              CTile W;
              if (has_op) {
                W.zero();
                s3f_apply_op(W, T[j], ops + term * DP * DP);
              } else {
                W = T[j];
              }
              s3f_second<NRT>(acc, W, j, b_base, rt_stride);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bB + (2 * slot + 1) * 8);
          ++itB;
          if (++cB_slot == p.nstB) {
            cB_slot = 0;
            cB_par ^= 1;
          }
        }
      }
    }
#ifdef S3F_PROFILE
    if (lane == 0 && cta < 148 && warp < 12) {
      unsigned long long* o = s3f_prof + (cta * 12 + warp) * 4;
      o[0] = prof_wait[0];
      o[1] = prof_wait[1];
      o[2] = prof_wait[2];
      o[3] = clock64() - prof_t0;
    }
#endif
    // ---- partial result of this (CTA, group)
    cplx* part = p.partial + ((int64_t)cta * p.G + g) * ((int64_t)p.Pfull * p.Rfull * DP);
    const int row = wg * 8 + r;
    if (row < p.P) {
#pragma unroll
      for (int rt = 0; rt < NRT; ++rt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = rt * 8 + 2 * c + h;
          if (col < p.R) {
#pragma unroll
            for (int s = 0; s < DP; ++s) {
              cplx val;
              val.x = h ? acc[rt][s].re1 : acc[rt][s].re0;
              val.y = h ? acc[rt][s].im1 : acc[rt][s].im0;
              part[((int64_t)(p.p0 + row) * p.Rfull + p.r0 + col) * DP + s] = val;
            }
          }
        }
      }
    }
  }
}

template <int NRT, int NA, int WS>
int s3f_launch_one(const S3FParams& p, int ctas, int threads, cudaStream_t stream) {
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(cudaFuncSetAttribute(stage3f_kernel<NRT, s3f_tiles(NRT), NA, WS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S3F_SMEM_LIMIT));
    configured[dev] = true;
  }
  stage3f_kernel<NRT, s3f_tiles(NRT), NA, WS><<<ctas, WS ? (WS + 4) * 32 : threads, p.smem_total, stream>>>(p);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// the instantiation for a launch whose S blocks hold `tiles` tiles
template <int NRT, int WS>
int s3f_launch_width(const S3FParams& p, int tiles, int ctas, int threads, cudaStream_t stream) {
  constexpr int NT = s3f_tiles(NRT);
  if (tiles == NT) return s3f_launch_one<NRT, NT, WS>(p, ctas, threads, stream);
  if (NT > 1 && tiles == NT - 1) return s3f_launch_one<NRT, (NT > 1 ? NT - 1 : NT), WS>(p, ctas, threads, stream);
  return s3f_launch_one<NRT, 0, WS>(p, ctas, threads, stream);
}

// A second stream per device, so that the launches of one apply (one per S-block width, together one CTA per SM) run
// side by side: fork from the caller's stream, join back into it.
struct S3FSideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
int s3f_side_stream(S3FSideStream** out) {
  static S3FSideStream side[16];
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  CARC_REQUIRE(dev < 16, CARC_ERR_VALUE, "device index %d not supported", dev);
  S3FSideStream& s = side[dev];
  if (!s.stream) {
    CARC_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CARC_CHECK_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    CARC_CHECK_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return CARC_OK;
}

// One launch per S-block width: the CTAs whose block holds `tiles` tiles run the kernel instantiated for that count.
// The widest blocks go to the caller's stream, the others (an uneven split of S has two widths) to the side stream.
template <int NRT>
int s3f_launch(S3FParams& p, const Stage3FConfig& k, cudaStream_t stream) {
  int widths[S3F_MAX_SB + 1], nw = 0;
  for (int i = 0; i < k.NSB; ++i) {
    const int tiles = k.sb_tile0[i + 1] - k.sb_tile0[i];
    bool seen = false;
    for (int j = 0; j < nw; ++j) seen = seen || widths[j] == tiles;
    if (!seen) widths[nw++] = tiles;
  }
  S3FSideStream* side = nullptr;
  if (nw > 1) {
    int rc = s3f_side_stream(&side);
    if (rc) return rc;
    CARC_CHECK_CUDA(cudaEventRecord(side->fork, stream));
    CARC_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  for (int w = nw - 1; w >= 0; --w) {       // side-stream launches first, the caller's stream last
    const int tiles = widths[w];
    cudaStream_t st = w == 0 ? stream : side->stream;
    int n = 0;
    for (int c = 0; c < k.ctas; ++c)
      if (k.sb_tile0[k.cta_sb[c] + 1] - k.sb_tile0[k.cta_sb[c]] == tiles) p.cta_map[n++] = (unsigned char)c;
    const int rc = k.ws ? s3f_launch_width<NRT, s3f_max_threads(NRT) / 32>(p, tiles, n, k.threads, st)
                        : s3f_launch_width<NRT, 0>(p, tiles, n, k.threads, st);
    if (rc) return rc;
  }
  if (nw > 1) {
    CARC_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
    CARC_CHECK_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
  }
  return CARC_OK;
}

}  // namespace

// Wait-time counters of the last stage3f launch: [148 CTAs][12 warps][first-term A wait, later-term A wait, B wait,
// total] in cycles; CARC_ERR_UNSUPPORTED unless the library was built with -DS3F_PROFILE.
int stage3f_profile_read(unsigned long long* host) {
#ifdef S3F_PROFILE
  CARC_CHECK_CUDA(cudaDeviceSynchronize());
  CARC_CHECK_CUDA(cudaMemcpyFromSymbol(host, s3f_prof, sizeof(unsigned long long) * 148 * 12 * 4));
  return CARC_OK;
#else
  (void)host;
  set_error("stage3f_profile_read: library built without -DS3F_PROFILE");
  return CARC_ERR_UNSUPPORTED;
#endif
}

// Launch geometry of the folded kernel for one shape; false when the shape is outside its envelope.
bool stage3f_configure(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, Stage3FConfig* cfg) {
  if (d != 2) return false;
  if (Xmax >= (1ll << 31) || Xmax < 1) return false;
  if (P < 1 || R < 1 || Q < 1 || S < 1) return false;
  // Outputs wider than the 64 columns a warp can accumulate (8 row tiles x 2 spins of register tiles) or taller than the
  // CTA's warps and shared memory can stage are computed block by block: RB column blocks x PB row blocks, one set of
  // launches each, all writing their part of the same partial-sum slots.  Row blocks cost nothing extra in arithmetic
  // (rows are independent; B_x is streamed once per row block); every column block repeats the first products.
  // Among the row-block counts that fit, take the one that loads the four DMMA pipes of an SM most evenly: a block of
  // NPT_b row tiles runs as G groups of NPT_b warps, warp w on sub-partition w % 4, so per environment index the busiest
  // sub-partition issues ceil(G NPT_b / 4) / G tile-ops.  (D = 6: one block of 5 row tiles = 10 warps as 3 + 3 + 2 + 2
  // -> 1.5; blocks of 3 + 2 row tiles = 12 and 8 warps, both even -> 1.25: 18.3 -> 19.6 TFLOP/s.)  Each extra block
  // re-streams B_x once more.
  const int rt_all = (R + 7) / 8, pt_all = (P + 7) / 8;
  static const int forced_pb = getenv("CARC_S3F_PB") ? atoi(getenv("CARC_S3F_PB")) : 0;     // experiments
  for (int RB = (rt_all + 7) / 8; RB <= rt_all && RB <= 4; ++RB) {
    bool found = false;
    double best = 0.0;
    int feasible = 0;
    for (int PB = 1; PB <= pt_all && PB <= 8 && feasible < 3; ++PB) {
      Stage3FConfig k;
      k.RB = RB;
      k.PB = PB;
      if (!stage3f_configure_block(nterms, P, Q, R, S, d, Xmax, &k)) continue;
      ++feasible;
      // per block: warps on the busiest sub-partition / G, divided by how well that many co-resident warps keep the pipe
      // fed (measured: one warp alone ~0.6, two ~0.85, three ~0.95 of what the pipe can take)
      static const double fed[4] = {1.0, 0.6, 0.85, 0.95};
      double cost = 0.03 * PB;
      for (int pb = 0; pb < PB; ++pb) {
        const int tiles = std::min(k.NPT, pt_all - pb * k.NPT);
        if (tiles <= 0) continue;
        const int busiest = (k.G * tiles + 3) / 4;
        cost += (double)busiest / fed[std::min(busiest, 3)] / k.G;
      }
      if (forced_pb > 0) cost = PB == forced_pb ? 0.0 : 1e9 + PB;
      if (!found || cost < best - 1e-9) {
        found = true;
        best = cost;
        *cfg = k;
      }
    }
    if (found) return true;
  }
  return false;
}

// Launch geometry for the block counts k->PB, k->RB; false when a block does not fit the kernel.
bool stage3f_configure_block(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, Stage3FConfig* cfg) {
  Stage3FConfig k = *cfg;
  const int rt_all = (R + 7) / 8, pt_all = (P + 7) / 8;
  k.NPT = (pt_all + k.PB - 1) / k.PB;      // row tiles of the tallest block
  k.NRT = (rt_all + k.RB - 1) / k.RB;      // column tiles of the widest block
  if (k.NRT > 8) return false;
  k.Q4 = (Q + 3) / 4;
  const int ntt = (S + 3) / 4;
  const int ntmax = s3f_tiles(k.NRT);
  k.NSB = (ntt + ntmax - 1) / ntmax;
  if (k.NSB > S3F_MAX_SB) return false;
  const int maxwarps = s3f_max_threads(k.NRT) / 32;
  if (k.NPT > maxwarps) return false;
  // tiles as evenly as the shape allows
  k.sb_tile0[0] = 0;
  for (int i = 0; i < k.NSB; ++i) k.sb_tile0[i + 1] = k.sb_tile0[i] + ntt / k.NSB + (i < ntt % k.NSB ? 1 : 0);
  const int nt_block = k.sb_tile0[1];   // the widest block comes first
  k.QS = (k.Q4 / 2) * 8;   // the paired k-steps; the trailing single step has its own array
  while (k.QS % 8 != 1) ++k.QS;
  k.BSTR = 4 * nt_block;
  while (k.BSTR % 8 != 4) ++k.BSTR;
  // One S block: B_x is contiguous in memory, so it is staged with its own row stride by a single copy instead of R
  // row copies (the per-copy issue cost is what limits the small shapes; the 2-way bank conflict that a row stride of
  // 0 mod 8 causes on the B fragments is the smaller price)
  k.b_whole = k.NSB == 1 ? 1 : 0;
  if (k.b_whole) k.BSTR = S;
  k.slotA = (uint32_t)((8 * Q + 8) * 16);   // one warp's 8 rows (+ the trailing k-step's overrun)
  k.slotB = (uint32_t)((k.NRT * 8 * k.BSTR + 4) * 16);   // + the last row's overrun into padded columns
  const uint32_t vbytes = (uint32_t)(8 * nt_block * k.QS * 16);
  const uint32_t vtail_bytes = (uint32_t)(8 * nt_block * 4 * 16);
  const uint32_t obytes = (uint32_t)(nterms * d * d * 16);
  // CARC_S3F_WS (experiments): 0 = symmetric kernels only, 8 = warp-specialised only where a warp holds 7 or 8 column
  // tiles, otherwise (default) warp-specialised everywhere; its four producer warps serve at most four groups
  static const int ws_env = getenv("CARC_S3F_WS") ? atoi(getenv("CARC_S3F_WS")) : 1;
  k.ws = ws_env == 0 ? 0 : (ws_env == 8 && maxwarps != 8) ? 0 : maxwarps;
  int G = std::min(k.ws ? 4 : 8, maxwarps / k.NPT);
  if (Xmax < G) G = (int)std::max<int64_t>(1, Xmax);
  for (; G >= 1; --G) {
    for (int nstA = 3; nstA >= 2; --nstA) {
      for (int nstB = 4; nstB >= nstA; --nstB) {
        // (the warp-specialised kernel has an empty barrier per A slot as well)
        const uint32_t bar_bytes = (uint32_t)(((G * (k.NPT * nstA * (k.ws ? 2 : 1) + 2 * nstB) * 8) + 127) / 128 * 128);
        const uint32_t ops_off = bar_bytes;
        const uint32_t hasop_off = ops_off + obytes;
        const uint32_t tab_off = (hasop_off + (uint32_t)nterms * 4 + 15) / 16 * 16;
        const uint32_t vt_off = (tab_off + (uint32_t)nterms * 16 + (uint32_t)nterms * 24 + 127) / 128 * 128;
        const uint32_t vtail_off = (vt_off + vbytes + 127) / 128 * 128;
        const uint32_t ring_off = (vtail_off + vtail_bytes + 127) / 128 * 128;
        const uint64_t total =
            (uint64_t)ring_off + (uint64_t)G * ((uint64_t)k.NPT * nstA * k.slotA + (uint64_t)nstB * k.slotB);
        if (total > (uint64_t)S3F_SMEM_LIMIT) continue;
        k.G = G;
        k.nstA = nstA;
        k.nstB = nstB;
        k.ops_off = ops_off;
        k.hasop_off = hasop_off;
        k.tab_off = tab_off;
        k.vt_off = vt_off;
        k.vtail_off = vtail_off;
        k.ring_off = ring_off;
        k.total = (uint32_t)total;
        k.threads = G * k.NPT * 32;
        // share the CTAs (one per SM) between the S blocks in proportion to their tile counts
        const int64_t max_slabs = std::max<int64_t>(1, Xmax / G);
        int budget = std::min(sm_count(), 160), used = 0;   // one CTA per SM (the CTA tables hold 160)
        if (k.NSB > budget) budget = k.NSB;
        k.sb_cta0[0] = 0;
        for (int i = 0; i < k.NSB; ++i) {
          const int tiles = k.sb_tile0[i + 1] - k.sb_tile0[i];
          int64_t n = (int64_t)budget * tiles / ntt;
          n = std::max<int64_t>(1, std::min<int64_t>(n, max_slabs));
          k.sb_cta0[i + 1] = k.sb_cta0[i] + (int)n;
          used += (int)n;
        }
        k.ctas = used;
        k.slots = used * G;
        // CTA numbering: the CTAs of different S blocks that stream the same stretch of X read the same A_x; neighbours
        // in blockIdx are placed on neighbouring SMs (same die, same L2 partition), so order all CTAs by the position of
        // their slab (measured at D = 8: 73.6 GB of DRAM reads per launch with S-block-major numbering, 1.43 x the
        // tensors' size)
        {
          int next[S3F_MAX_SB] = {0};
          for (int c = 0; c < used; ++c) {
            int best = -1;
            double best_pos = 0.0;
            for (int i = 0; i < k.NSB; ++i) {
              const int n = k.sb_cta0[i + 1] - k.sb_cta0[i];
              if (next[i] >= n) continue;
              const double pos = (next[i] + 0.5) / n;
              if (best < 0 || pos < best_pos) { best = i; best_pos = pos; }
            }
            k.cta_sb[c] = (unsigned char)best;
            k.cta_sl[c] = (unsigned char)next[best]++;
          }
        }
        // DMMA work per x and product pair, in complex 8x8x4 steps, for the choice between the two kernels
        k.padded_work = (double)pt_all * ntt * ((double)k.Q4 * k.RB + 2.0 * k.NRT * k.RB);
        *cfg = k;
        return true;
      }
    }
  }
  return false;
}

int stage3f_launch(const Stage3Plan* plan, const Stage3FConfig& k, int P, int Q, int R, int S, const cplx* v, cplx* partial,
                   cudaStream_t stream) {
  S3FParams p;
  p.terms = plan->terms_dev;
  p.groups = plan->groups_dev;
  p.nterms = (int)plan->terms.size();
  p.ngroups = (int)plan->groups.size();
  p.P = P; p.Q = Q; p.R = R; p.S = S;
  p.Pfull = P; p.Rfull = R; p.p0 = 0; p.r0 = 0;
  p.Q4 = k.Q4; p.NPT = k.NPT; p.G = k.G; p.NSB = k.NSB; p.nstA = k.nstA; p.nstB = k.nstB; p.QS = k.QS; p.BSTR = k.BSTR;
  p.b_whole = k.b_whole;
  // a spinning producer warp takes issue slots from the consumer warp on its SM sub-partition: 34.0 -> 34.15 TFLOP/s at
  // D = 8 with a 200 ns back-off (an op lasts ~4 us there; no change at D = 4, 6)
  static const int psleep = getenv("CARC_S3F_SLEEP") ? atoi(getenv("CARC_S3F_SLEEP")) : 200;
  p.psleep = psleep;
  for (int i = 0; i <= S3F_MAX_SB; ++i) {
    p.sb_cta0[i] = i <= k.NSB ? k.sb_cta0[i] : 0;
    p.sb_tile0[i] = i <= k.NSB ? k.sb_tile0[i] : 0;
  }
  for (int c = 0; c < 160; ++c) {
    p.cta_sb[c] = c < k.ctas ? k.cta_sb[c] : 0;
    p.cta_sl[c] = c < k.ctas ? k.cta_sl[c] : 0;
  }
  p.slotA_bytes = k.slotA; p.slotB_bytes = k.slotB;
  p.ops_off = k.ops_off; p.hasop_off = k.hasop_off; p.tab_off = k.tab_off; p.vt_off = k.vt_off; p.vtail_off = k.vtail_off; p.ring_off = k.ring_off;
  p.smem_total = k.total;
  p.v = v;
  p.partial = partial;
  for (int c = 0; c < 160; ++c) p.cta_map[c] = 0;
  for (int pb = 0; pb < k.PB; ++pb) {
    for (int rb = 0; rb < k.RB; ++rb) {
      p.p0 = pb * 8 * k.NPT;
      p.r0 = rb * 8 * k.NRT;
      p.P = std::min(8 * k.NPT, P - p.p0);
      p.R = std::min(8 * k.NRT, R - p.r0);
      if (p.P <= 0 || p.R <= 0) continue;
      Stage3FConfig kb = k;            // same groups, S blocks, slabs and rings; only the warps of this block's row tiles
      kb.NPT = (p.P + 7) / 8;
      kb.threads = k.G * kb.NPT * 32;
      p.NPT = kb.NPT;
      int rc;
      switch (k.NRT) {
        case 1: rc = s3f_launch<1>(p, kb, stream); break;
        case 2: rc = s3f_launch<2>(p, kb, stream); break;
        case 3: rc = s3f_launch<3>(p, kb, stream); break;
        case 4: rc = s3f_launch<4>(p, kb, stream); break;
        case 5: rc = s3f_launch<5>(p, kb, stream); break;
        case 6: rc = s3f_launch<6>(p, kb, stream); break;
        case 7: rc = s3f_launch<7>(p, kb, stream); break;
        default: rc = s3f_launch<8>(p, kb, stream); break;
      }
      if (rc) return rc;
    }
  }
  return CARC_OK;
}

}  // namespace carc
