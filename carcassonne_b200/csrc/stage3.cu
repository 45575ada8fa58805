// Fused center-site matvec ("stage 3") for sm_100a.
//
//   out[P,R,s] = sum_t sum_x sum_S B_t[x,R,S] * ( sum_s' O_t[s,s'] * ( sum_Q A_t[x,P,Q] * v[Q,S,s'] ) )
//
// with A_t = [(x y), D0*, D1*, D0, D1] and B_t = [(y' x'), D2*, D3*, D2, D3] the pre-joined stage-2 halves,
// P = Q = D0*D1, R = S = D2*D3 (reference tensors/_2d/dense.py:115-160: t = A.(3,4)*v.(0,1);
// out = B.(3,4,0)*t.(3,4,0); join [2][3][0][1][4]; the operator variant applies O to v's physical leg first;
// tensors/_2d/sparse.py:129-133 sums the terms).
//
// The reference runs this as two ZGEMMs per term with the X*P*S*d intermediate spilled to memory plus two full
// transposing copies.  Here the intermediate never leaves registers: for every environment index x a warp that
// owns 8 rows of P computes a 8 x 8 tile of A_x v with DMMA, folds the site operator in on the accumulator
// fragments (coefficients that are exactly zero -- Pauli structure -- are skipped), and immediately uses the
// result as the A operand of the second DMMA chain against B_x: the C-fragment of m8n8k4 maps onto its A-fragment
// when the K index of the second product is taken in the order {2c+e}, so no shuffle is needed.
//
// A_x (contiguous P*Q block) and the S-block of B_x (R rows) are brought in by TMA bulk copies
// (cp.async.bulk -> UBLKCP) into a multi-stage ring guarded by mbarriers.  There is no dedicated producer warp:
// the register file is split per SM sub-partition (16K registers each), so a 9th warp would cap every thread at
// 170 registers; instead warp 0 of each consumer group runs a second cursor one item ahead and issues the copies.
// Terms that share the same B tensor (e.g. every term whose half-1 tag is Identity) are grouped: their first
// products are summed in registers, T = sum_t O_t (A_t,x v), and ONE second product with B_x follows (linearity),
// which removes a third of the DMMA work of the transverse-Ising operator (9 + 9 -> 9 + 6 products per x).
// Work decomposition: CTA = (S block of 16 columns, slab of x); inside a CTA, G independent consumer groups of
// ceil(P/8) warps take interleaved x.  Every (CTA, group) writes one partial result; a second kernel sums the
// partials in a fixed order, so the result is deterministic.
#include <algorithm>
#include <vector>

#include "carc_internal.h"
#include "common.cuh"

namespace carc {

namespace {

// S-block width = 8 * CH columns (CH = 1 or 2 chunks of 8); the staged B block has an odd row stride (SBW + 1) so that
// the paired fragment loads are conflict-free.  CH = 1 halves the T / U register tiles, which lets shapes with 5-6 row
// tiles (D = 6) run two consumer groups per CTA without spilling.
constexpr int SMEM_LIMIT = 227 * 1024;

struct S3Params {
  const Stage3Term* terms;  // device copy, sorted by group
  const Stage3Group* groups;
  int nterms, ngroups;
  int P, Q, R, S;
  int Q8, NPT, G, NSB, NSL, nstA, nstB;
  uint32_t slotA_bytes, slotB_bytes, ops_off, hasop_off, tab_off, vt_off, ring_off, smem_total;
  const cplx* v;
  cplx* partial;
};

// threads per CTA are a multiple of 128 (one warp per sub-partition): 2 / 3 warps per sub-partition leave
// 255 / 170 registers per thread (4 warps would leave 128, which spills the accumulator tiles)
__host__ __device__ constexpr int s3_max_threads(int nrt) {
  return nrt >= 7 ? 256 : 384;
}

template <int NRT, int DP, int CH>
__global__ void __launch_bounds__(s3_max_threads(NRT), 1) stage3_kernel(const S3Params p) {
  constexpr int SBW = 8 * CH, BSTR = SBW + 1;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int ncw = p.G * p.NPT;
  const int nbar_g = 2 * (p.nstA + p.nstB);
  const uint32_t bars = smem_u32(smem);
  cplx* ops = reinterpret_cast<cplx*>(smem + p.ops_off);
  cplx* Vt = reinterpret_cast<cplx*>(smem + p.vt_off);
  const int QS = p.Q8 * 8 + 1;
  const uint32_t group_bytes = p.nstA * p.slotA_bytes + p.nstB * p.slotB_bytes;

  // zero everything behind the barriers: padding rows / columns must read as finite zeros forever
  {
    uint4* z = reinterpret_cast<uint4*>(smem + p.ops_off);
    const uint32_t n16 = (p.smem_total - p.ops_off) / 16;
    for (uint32_t i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int g = 0; g < p.G; ++g) {
      const uint32_t b = bars + g * nbar_g * 8;
      for (int i = 0; i < p.nstA; ++i) {
        mbar_init(b + (2 * i) * 8, 1);           // full A
        mbar_init(b + (2 * i + 1) * 8, p.NPT);   // empty A
      }
      for (int i = 0; i < p.nstB; ++i) {
        mbar_init(b + (2 * p.nstA + 2 * i) * 8, 1);
        mbar_init(b + (2 * p.nstA + 2 * i + 1) * 8, p.NPT);
      }
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int sb = blockIdx.x % p.NSB, sl = blockIdx.x / p.NSB;
  const int S0 = sb * SBW;
  const int SBv = min(SBW, p.S - S0);
  int* hasop = reinterpret_cast<int*>(smem + p.hasop_off);
  // small per-launch tables in shared memory: they sit on the critical path of every item
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
  for (int i = tid; i < p.Q * SBv * DP; i += blockDim.x) {
    const int sp = i % DP, sc = (i / DP) % SBv, q = i / (DP * SBv);
    Vt[(sp * SBW + sc) * QS + q] = p.v[((int64_t)q * p.S + S0 + sc) * DP + sp];
  }
  fence_proxy_async();
  __syncthreads();

  if (warp >= ncw) return;
  {
    const int g = warp / p.NPT, wg = warp % p.NPT;
    const uint32_t b = bars + g * nbar_g * 8;
    const uint32_t ring = smem_u32(smem + p.ring_off) + g * group_bytes;
    CTile acc[NRT][DP];
#pragma unroll
    for (int i = 0; i < NRT; ++i)
#pragma unroll
      for (int s = 0; s < DP; ++s) acc[i][s].zero();

    const bool eswap = ((p.Q & 1) == 0) && (r & 1);
    const uint32_t a_lane_off = (uint32_t)(((wg * 8 + r) * p.Q + 2 * c) * 16);
    const uint32_t a_first = eswap ? 16u : 0u, a_second = 16u - a_first;
    const uint32_t v_base = smem_u32(Vt) + (uint32_t)((r * QS + 2 * c) * 16);
    const uint32_t b_lane_off = (uint32_t)((r * BSTR + 2 * c) * 16);

    // Op cursors.  The work of this (CTA slab, consumer group) is the flattened sequence, over term groups gi and
    // environment indices x, of [A-op for each term of the group ..., B-op of the group].  The producer cursor
    // runs ahead of the consumer cursor by as many ops as the two rings have room for.
    struct Cursor {
      int gi, j, x, x_hi;   // group, op within the group (producer only), environment index and its slab end
    };
    auto next_item = [&](Cursor& cu) -> bool {
      cu.x += p.G;
      while (cu.gi < 0 || cu.x >= cu.x_hi) {
        if (++cu.gi >= p.ngroups) return false;
        const int64_t X = groupX[cu.gi];
        cu.x = (int)(X * sl / p.NSL) + g;
        cu.x_hi = (int)(X * (sl + 1) / p.NSL);
      }
      return true;
    };
    auto advance = [&](Cursor& cu) -> bool {   // producer: next op of the flattened sequence
      if (cu.gi >= 0 && cu.j < groupCount[cu.gi]) {
        ++cu.j;
        return true;
      }
      cu.j = 0;
      return next_item(cu);
    };
    const uint32_t a_bytes = (uint32_t)(p.P * p.Q * 16);
    const uint32_t b_row_bytes = (uint32_t)(SBv * 16);
    auto issueA = [&](const Cursor& cu, uint32_t it) {
      const cplx* A = groupKind[cu.gi] ? groupCenter[cu.gi] : termA[groupFirst[cu.gi] + cu.j];
      const int sa = it % p.nstA;
      const uint32_t fullA = b + (2 * sa) * 8;
      if (lane == 0) {
        mbar_wait(fullA + 8, ((it / p.nstA) & 1) ^ 1);
        mbar_arrive_expect_tx(fullA, a_bytes);
        bulk_g2s(ring + sa * p.slotA_bytes, A + (int64_t)cu.x * p.P * p.Q, a_bytes, fullA);
      }
      __syncwarp();
    };
    auto issueB = [&](const Cursor& cu, uint32_t it) {
      const cplx* B = groupKind[cu.gi] ? termB[groupFirst[cu.gi] + cu.j - 1] : groupCenter[cu.gi];
      const int sbq = it % p.nstB;
      const uint32_t fullB = b + (2 * p.nstA + 2 * sbq) * 8;
      if (lane == 0) {
        mbar_wait(fullB + 8, ((it / p.nstB) & 1) ^ 1);
        mbar_arrive_expect_tx(fullB, b_row_bytes * p.R);
      }
      __syncwarp();
      const uint32_t dst = ring + p.nstA * p.slotA_bytes + sbq * p.slotB_bytes;
      const cplx* src = B + ((int64_t)cu.x * p.R) * p.S + S0;
      for (int rr = lane; rr < p.R; rr += 32) bulk_g2s(dst + rr * BSTR * 16, src + (int64_t)rr * p.S, b_row_bytes, fullB);
    };

    Cursor cc = {-1, 0, 0, 0}, pc = {-1, 0, 0, 0};
    uint32_t issuedA = 0, issuedB = 0, itA = 0, itB = 0;
    bool more = true, pending = false;
    auto run_ahead = [&]() {
      while (more) {
        if (!pending) {
          more = advance(pc);
          pending = more;
          if (!more) break;
        }
        if (groupKind[pc.gi] ? (pc.j == 0) : (pc.j < groupCount[pc.gi])) {
          if (issuedA - itA >= (uint32_t)p.nstA) break;
          issueA(pc, issuedA++);
        } else {
          if (issuedB - itB >= (uint32_t)p.nstB) break;
          issueB(pc, issuedB++);
        }
        pending = false;
      }
    };

    // consumer: one iteration per (term group, x): the first products of the group's terms accumulate in T, the
    // second product follows in straight-line code
    while (next_item(cc)) {
      CTile T[CH][DP];
      const int first = groupFirst[cc.gi], cnt = groupCount[cc.gi];
      const int kind = groupKind[cc.gi];
      {
        // ---- first term of the group: both S chunks accumulate straight into T (8 independent DMMA chains, the A
        // fragments are loaded once for both chunks), then the site operator is applied in place
        if (wg == 0) run_ahead();
        const int term = first;
        const bool has_op = !kind && hasop[term] != 0;   // an A-star keeps the raw product: each of its terms has its own O
        const int slot = itA % p.nstA;
        mbar_wait(b + (2 * slot) * 8, (itA / p.nstA) & 1);
        const uint32_t a_base = ring + slot * p.slotA_bytes + a_lane_off;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
          for (int s = 0; s < DP; ++s) T[ch][s].zero();
#pragma unroll 2
        for (int kp = 0; kp < p.Q8; ++kp) {
          const cplx x0 = lds_c(a_base + kp * 128 + a_first);
          const cplx x1 = lds_c(a_base + kp * 128 + a_second);
          const cplx a0 = eswap ? x1 : x0;
          const cplx a1 = eswap ? x0 : x1;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
            for (int s = 0; s < DP; ++s) {
              const uint32_t va = v_base + (uint32_t)((((s * SBW + ch * 8) * QS) + kp * 8) * 16);
              const cplx b0 = lds_c(va);
              const cplx b1 = lds_c(va + 16);
              cmma(T[ch][s], a0.x, a0.y, -a0.y, b0.x, b0.y);
              cmma(T[ch][s], a1.x, a1.y, -a1.y, b1.x, b1.y);
            }
          }
        }
        if (has_op) {
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            CTile U[DP];
#pragma unroll
            for (int s = 0; s < DP; ++s) {
              U[s] = T[ch][s];
              T[ch][s].zero();
            }
#pragma unroll
            for (int s = 0; s < DP; ++s) {
#pragma unroll
              for (int s2 = 0; s2 < DP; ++s2) {
                const cplx w = ops[term * DP * DP + s * DP + s2];
                if (w.x != 0.0 || w.y != 0.0) {
                  T[ch][s].re0 += w.x * U[s2].re0 - w.y * U[s2].im0;
                  T[ch][s].im0 += w.x * U[s2].im0 + w.y * U[s2].re0;
                  T[ch][s].re1 += w.x * U[s2].re1 - w.y * U[s2].im1;
                  T[ch][s].im1 += w.x * U[s2].im1 + w.y * U[s2].re1;
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(b + (2 * slot + 1) * 8);
        ++itA;
      }
      if (!kind) {
        for (int j = 1; j < cnt; ++j) {
          if (wg == 0) run_ahead();
          // ---- first product of a further term of the group: U = A_x v (8 rows of P, this S block); T (+)= O U
          const int term = first + j;
          const bool has_op = hasop[term] != 0;
          const int slot = itA % p.nstA;
          mbar_wait(b + (2 * slot) * 8, (itA / p.nstA) & 1);
          const uint32_t a_base = ring + slot * p.slotA_bytes + a_lane_off;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            CTile U[DP];
#pragma unroll
            for (int s = 0; s < DP; ++s) U[s].zero();
#pragma unroll 2
            for (int kp = 0; kp < p.Q8; ++kp) {
              const cplx x0 = lds_c(a_base + kp * 128 + a_first);
              const cplx x1 = lds_c(a_base + kp * 128 + a_second);
              const cplx a0 = eswap ? x1 : x0;
              const cplx a1 = eswap ? x0 : x1;
#pragma unroll
              for (int s = 0; s < DP; ++s) {
                const uint32_t va = v_base + (uint32_t)((((s * SBW + ch * 8) * QS) + kp * 8) * 16);
                const cplx b0 = lds_c(va);
                const cplx b1 = lds_c(va + 16);
                cmma(U[s], a0.x, a0.y, -a0.y, b0.x, b0.y);
                cmma(U[s], a1.x, a1.y, -a1.y, b1.x, b1.y);
              }
            }
            if (!has_op) {
#pragma unroll
              for (int s = 0; s < DP; ++s) {
                T[ch][s].re0 += U[s].re0;
                T[ch][s].im0 += U[s].im0;
                T[ch][s].re1 += U[s].re1;
                T[ch][s].im1 += U[s].im1;
              }
            } else {
#pragma unroll
              for (int s = 0; s < DP; ++s) {
#pragma unroll
                for (int s2 = 0; s2 < DP; ++s2) {
                  const cplx w = ops[term * DP * DP + s * DP + s2];
                  if (w.x != 0.0 || w.y != 0.0) {
                    T[ch][s].re0 += w.x * U[s2].re0 - w.y * U[s2].im0;
                    T[ch][s].im0 += w.x * U[s2].im0 + w.y * U[s2].re0;
                    T[ch][s].re1 += w.x * U[s2].re1 - w.y * U[s2].im1;
                    T[ch][s].im1 += w.x * U[s2].im1 + w.y * U[s2].re1;
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b + (2 * slot + 1) * 8);
          ++itA;
        }
        if (wg == 0) run_ahead();
        // ---- second product of the group: acc += T * B_x^T ; T's C fragments are the A fragments
        {
          const int slot = itB % p.nstB;
          mbar_wait(b + (2 * p.nstA + 2 * slot) * 8, (itB / p.nstB) & 1);
          const uint32_t b_base = ring + p.nstA * p.slotA_bytes + slot * p.slotB_bytes + b_lane_off;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
            for (int rt = 0; rt < NRT; ++rt) {
              const uint32_t ba = b_base + (uint32_t)(((rt * 8) * BSTR + ch * 8) * 16);
              const cplx b0 = lds_c(ba);
              const cplx b1 = lds_c(ba + 16);
#pragma unroll
              for (int s = 0; s < DP; ++s) {
                cmma(acc[rt][s], T[ch][s].re0, T[ch][s].im0, -T[ch][s].im0, b0.x, b0.y);
                cmma(acc[rt][s], T[ch][s].re1, T[ch][s].im1, -T[ch][s].im1, b1.x, b1.y);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b + (2 * p.nstA + 2 * slot + 1) * 8);
          ++itB;
        }
      } else {
        // ---- A-star: T holds U = A_x v once; every term of the group applies its own site operator to U and multiplies
        // with its own B_x
        for (int j = 0; j < cnt; ++j) {
          if (wg == 0) run_ahead();
          const int term = first + j;
          const bool has_op = hasop[term] != 0;
          const int slot = itB % p.nstB;
          mbar_wait(b + (2 * p.nstA + 2 * slot) * 8, (itB / p.nstB) & 1);
          const uint32_t b_base = ring + p.nstA * p.slotA_bytes + slot * p.slotB_bytes + b_lane_off;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            CTile W[DP];
#pragma unroll
            for (int s = 0; s < DP; ++s) {
              if (!has_op) {
                W[s] = T[ch][s];
              } else {
                W[s].zero();
#pragma unroll
                for (int s2 = 0; s2 < DP; ++s2) {
                  const cplx w = ops[term * DP * DP + s * DP + s2];
                  if (w.x != 0.0 || w.y != 0.0) {
                    W[s].re0 += w.x * T[ch][s2].re0 - w.y * T[ch][s2].im0;
                    W[s].im0 += w.x * T[ch][s2].im0 + w.y * T[ch][s2].re0;
                    W[s].re1 += w.x * T[ch][s2].re1 - w.y * T[ch][s2].im1;
                    W[s].im1 += w.x * T[ch][s2].im1 + w.y * T[ch][s2].re1;
                  }
                }
              }
            }
#pragma unroll
            for (int rt = 0; rt < NRT; ++rt) {
              const uint32_t ba = b_base + (uint32_t)(((rt * 8) * BSTR + ch * 8) * 16);
              const cplx b0 = lds_c(ba);
              const cplx b1 = lds_c(ba + 16);
#pragma unroll
              for (int s = 0; s < DP; ++s) {
                cmma(acc[rt][s], W[s].re0, W[s].im0, -W[s].im0, b0.x, b0.y);
                cmma(acc[rt][s], W[s].re1, W[s].im1, -W[s].im1, b1.x, b1.y);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(b + (2 * p.nstA + 2 * slot + 1) * 8);
          ++itB;
        }
      }
    }
    // ---- partial result of this (CTA, group)
    cplx* part = p.partial + ((int64_t)blockIdx.x * p.G + g) * ((int64_t)p.P * p.R * DP);
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
              part[((int64_t)row * p.R + col) * DP + s] = val;
            }
          }
        }
      }
    }
  }
}

// out[i] = sum_slots partial[slot][i] in a fixed order (deterministic, the same on every rank).  32 elements per CTA,
// 8 threads per element: thread (e, g) sums the slots g, g + 8, ... (independent loads, 512 contiguous bytes per warp),
// the eight partial sums are combined in order 0..7.  (One thread per element walking all 150 - 600 slots was a
// latency chain on 11 CTAs: 124 us per matvec at D = 6, chi = 8, a third of the fused kernel's own time.)
// (no __restrict__: the multi-GPU path reduces in place, out == partial)
__global__ void __launch_bounds__(256) s3_reduce_kernel(const cplx* partial, cplx* out, int64_t n, int slots, int accumulate) {
  __shared__ double sre[8][32], sim[8][32];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + e;
  double re = 0.0, im = 0.0;
  if (i < n) {
#pragma unroll 4
    for (int s = g; s < slots; s += 8) {
      const cplx v = partial[(int64_t)s * n + i];
      re += v.x;
      im += v.y;
    }
  }
  sre[g][e] = re;
  sim[g][e] = im;
  __syncthreads();
  if (g == 0 && i < n) {
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      re += sre[q][e];
      im += sim[q][e];
    }
    if (accumulate) {
      re += out[i].x;
      im += out[i].y;
    }
    out[i] = make_double2(re, im);
  }
}

struct S3Config {
  int NPT, NRT, Q8, NSB, NSL, G, nstA, nstB, CH;
  uint32_t slotA, slotB, ops_off, hasop_off, tab_off, vt_off, ring_off, total;
  int threads, slots;
};

bool s3_configure(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, S3Config* cfg) {
  if (d != 2) return false;
  if (Xmax >= (1ll << 31)) return false;
  if (P > 64 || R > 64 || P < 1 || R < 1) return false;
  S3Config k;
  k.NPT = (P + 7) / 8;
  k.NRT = (R + 7) / 8;
  k.Q8 = (Q + 7) / 8;
  // two chunks per S block unless the shape has 5-6 row tiles (then one chunk: see the kernel's note on CH)
  k.CH = (k.NRT == 5 || k.NRT == 6) ? 1 : 2;
  const int SBW = 8 * k.CH, BSTR = SBW + 1;
  k.NSB = (S + SBW - 1) / SBW;
  const int maxwarps = s3_max_threads(k.NRT) / 32;
  if (k.NPT > maxwarps) return false;
  k.slotA = (uint32_t)((k.NPT * 8 * Q + 8) * 16);
  k.slotB = (uint32_t)(k.NRT * 8 * BSTR * 16);
  const int QS = k.Q8 * 8 + 1;
  const uint32_t vbytes = (uint32_t)(d * SBW * QS * 16);
  const uint32_t obytes = (uint32_t)(nterms * d * d * 16);
  int G = std::min(8, maxwarps / k.NPT);
  if (Xmax < G) G = (int)std::max<int64_t>(1, Xmax);
  for (; G >= 1; --G) {
    for (int nstA = 3; nstA >= 2; --nstA) {
      for (int nstB = 4; nstB >= nstA; --nstB) {
        const uint32_t bar_bytes = (uint32_t)(((G * 2 * (nstA + nstB) * 8) + 127) / 128 * 128);
        const uint32_t ops_off = bar_bytes;
        const uint32_t hasop_off = ops_off + obytes;
        const uint32_t tab_off = (hasop_off + (uint32_t)nterms * 4 + 15) / 16 * 16;
        const uint32_t vt_off = (tab_off + (uint32_t)nterms * 16 + (uint32_t)nterms * 24 + 127) / 128 * 128;
        const uint32_t ring_off = (vt_off + vbytes + 127) / 128 * 128;
        const uint64_t total = (uint64_t)ring_off + (uint64_t)G * ((uint64_t)nstA * k.slotA + (uint64_t)nstB * k.slotB);
        if (total <= SMEM_LIMIT) {
          k.G = G;
          k.nstA = nstA;
          k.nstB = nstB;
          k.ops_off = ops_off;
          k.hasop_off = hasop_off;
          k.tab_off = tab_off;
          k.vt_off = vt_off;
          k.ring_off = ring_off;
          k.total = (uint32_t)total;
          k.threads = G * k.NPT * 32;
          int64_t nsl = std::max<int64_t>(1, sm_count() / k.NSB);
          nsl = std::min<int64_t>(nsl, std::max<int64_t>(1, Xmax / G));
          k.NSL = (int)nsl;
          k.slots = k.NSB * k.NSL * G;
          *cfg = k;
          return true;
        }
      }
    }
  }
  return false;
}

template <int NRT, int CH>
int s3_launch(const S3Params& p, const S3Config& k, cudaStream_t stream) {
  static bool configured[16] = {false};
  int dev = 0;
  CARC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 16 && !configured[dev]) {
    CARC_CHECK_CUDA(
        cudaFuncSetAttribute(stage3_kernel<NRT, 2, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured[dev] = true;
  }
  stage3_kernel<NRT, 2, CH><<<k.NSB * k.NSL, k.threads, k.total, stream>>>(p);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

// Unfused device path (any P, R, d): three DMMA GEMMs per term and x chunk, the intermediate written directly in
// the layout the second GEMM wants through the generalised output map -- no transposing copies.
int s3_unfused(const Stage3Term* terms, int nterms, int P, int Q, int R, int S, int d, const cplx* v, cplx* out,
               cplx* ws, int64_t ws_elems, cudaStream_t stream) {
  const cplx one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
  const int64_t n = (int64_t)P * R * d;
  const int64_t wsize = (int64_t)Q * S * d;
  CARC_REQUIRE(ws_elems >= wsize + (int64_t)P * S * d + 64, CARC_ERR_VALUE, "stage3: workspace too small");
  cplx* W = ws;
  cplx* T = ws + wsize;
  cplx* opdev = ws + (ws_elems - 64);   // the last 64 elements hold the d x d site operator of the current term
  const int64_t tcap = ws_elems - 64 - wsize;
  CARC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(cplx) * n, stream));
  for (int t = 0; t < nterms; ++t) {
    const cplx* w = v;
    if (terms[t].has_op) {
      // W[(Q S), s'] = sum_s v[(Q S), s] * O[s', s]
      CARC_CHECK_CUDA(cudaMemcpyAsync(opdev, terms[t].op, sizeof(cplx) * d * d, cudaMemcpyHostToDevice, stream));
      int rc = zgemm(OP_N, OP_T, (int64_t)Q * S, d, d, one, v, d, opdev, d, zero, W, nullptr, nullptr, 1, 0, 0, 0, stream);
      if (rc) return rc;
      w = W;
    }
    const int64_t X = terms[t].X;
    int64_t xc = std::max<int64_t>(1, tcap / ((int64_t)P * S * d));
    for (int64_t x0 = 0; x0 < X; x0 += xc) {
      const int64_t nx = std::min(xc, X - x0);
      // T'[p, s, x, S] = sum_q A[(x p), q] * w[q, (S s)]
      GemmOut o1;
      o1.m_div = P; o1.m_s1 = S; o1.m_s0 = (int64_t)d * nx * S;
      o1.n_div = d; o1.n_s1 = 1; o1.n_s0 = nx * S;
      int rc = zgemm(OP_N, OP_N, nx * P, (int64_t)S * d, Q, one, terms[t].A + x0 * P * Q, Q, w, (int64_t)S * d, zero, T,
                     &o1, nullptr, 1, 0, 0, 0, stream);
      if (rc) return rc;
      // out[p, r, s] += sum_{(x S)} T'[(p s), (x S)] * B[x, r, S]
      GemmOut o2;
      o2.m_div = d; o2.m_s1 = (int64_t)R * d; o2.m_s0 = 1;
      o2.n_div = R; o2.n_s1 = 0; o2.n_s0 = d;
      GemmKMap km;
      km.a_kdiv = 0; km.a_ks1 = 0; km.b_kdiv = S; km.b_ks1 = (int64_t)R * S;
      rc = zgemm(OP_N, OP_T, (int64_t)P * d, R, nx * S, one, T, nx * S, terms[t].B + x0 * R * S, S, one, out, &o2, &km, 1,
                 0, 0, 0, stream);
      if (rc) return rc;
    }
  }
  return CARC_OK;
}

// Which device path runs for a shape: the original fused tiling, the folded one, or neither (unfused GEMMs).
void s3_choose(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int force_path, S3Config* k, Stage3FConfig* kf,
               bool* can_fuse, bool* can_fold) {
  *can_fuse = force_path != 2 && force_path != 3 && Xmax > 0 && s3_configure(nterms, P, Q, R, S, d, Xmax, k);
  *can_fold = force_path != 2 && force_path != 1 && Xmax > 0 && stage3f_configure(nterms, P, Q, R, S, d, Xmax, kf);
  // three column blocks (R > 128) repeat every first product three times: measured slower than the unfused GEMMs
  // (D = 12: 107 vs 82 ms), so the automatic choice stops at two; force_path = 3 still runs it
  if (*can_fold && force_path == 0 && kf->RB > 2) *can_fold = false;
  if (*can_fuse && *can_fold) {
    // Both tilings fit.  Take the one that issues fewer DMMA steps for this shape (complex 8x8x4 steps per x and product
    // pair).  At equal work (P, Q, S multiples of 8 / 16: D = 4, 8) the measured winner depends on the work per op: the
    // folded kernel's symmetric warps (every warp issues its own copies) win when an A_x tile is large (D = 8: 28.2 vs
    // 27.6 TFLOP/s), stage3_kernel's single issuing warp when it is small (D = 4: 23.1 vs 21.1).
    const double work = (double)k->NPT * k->NSB * k->CH * (4.0 * k->Q8 + 4.0 * k->NRT);
    // (round 2: with its tile loops free of run-time predicates the folded kernel wins at equal work for every size --
    // D = 4, chi = 16: 25.8 vs 23.1 TFLOP/s)
    const bool fold = kf->padded_work <= 1.0001 * work;
    if (fold) *can_fuse = false;
    else *can_fold = false;
  }
}

}  // namespace

int64_t stage3_workspace_elems(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax) {
  S3Config k;
  Stage3FConfig kf;
  int64_t fused = 0;
  if (s3_configure(nterms, P, Q, R, S, d, Xmax, &k)) fused = (int64_t)k.slots * P * R * d;
  if (stage3f_configure(nterms, P, Q, R, S, d, Xmax, &kf)) fused = std::max(fused, (int64_t)kf.slots * P * R * d);
  // unfused path: W + at least 64 MiB worth of intermediate (or the whole thing if smaller)
  int64_t per_x = (int64_t)P * S * d;
  int64_t t = std::min<int64_t>(Xmax * per_x, std::max<int64_t>(per_x, (64ll << 20) / 16));
  int64_t unfused = (int64_t)Q * S * d + t;
  return std::max(fused, unfused) + 64;
}

// 1: original fused tiling, 3: folded fused tiling, 2: unfused GEMMs -- what stage3_apply runs for this shape
int stage3_path(int nterms, int P, int Q, int R, int S, int d, int64_t Xmax, int force_path) {
  S3Config k;
  Stage3FConfig kf;
  bool can_fuse = false, can_fold = false;
  s3_choose(nterms, P, Q, R, S, d, Xmax, force_path, &k, &kf, &can_fuse, &can_fold);
  return can_fold ? 3 : can_fuse ? 1 : 2;
}

// Decompose the term list -- a bipartite multigraph between half-0 tensors A and half-1 tensors B -- into stars:
// repeatedly take the tensor with the most remaining terms (ties: B first, then first appearance) together with all
// of them.  A B-star (kind 0) sums the first products of its terms before ONE second product; an A-star (kind 1)
// computes ONE first product and reuses it for the second product of each of its terms.  For the transverse-Ising
// operator (9 terms over 6 + 6 tensors) this gives 7 first + 6 second products per x instead of 9 + 9.
// host half of the plan: groups and the term table sorted by group (no device call; carc_stage3_describe_stars and the
// CPU test suite use it directly)
void stage3_plan_host(const Stage3Term* terms, int nterms, Stage3Plan* plan) {
  std::vector<int> group_of(nterms, -1);
  int remaining = nterms;
  while (remaining > 0) {
    int best_deg = 0, best_kind = 0, best_t = -1;
    for (int t = 0; t < nterms; ++t) {
      if (group_of[t] >= 0) continue;
      int degB = 0, degA = 0;
      for (int u = 0; u < nterms; ++u) {
        if (group_of[u] >= 0) continue;
        if (terms[u].B == terms[t].B && terms[u].X == terms[t].X) ++degB;
        if (terms[u].A == terms[t].A && terms[u].X == terms[t].X) ++degA;
      }
      if (degB > best_deg) { best_deg = degB; best_kind = 0; best_t = t; }
      if (degA > best_deg) { best_deg = degA; best_kind = 1; best_t = t; }
    }
    Stage3Group gr;
    gr.kind = best_kind;
    gr.center = best_kind ? terms[best_t].A : terms[best_t].B;
    gr.X = terms[best_t].X;
    gr.first = 0;
    gr.count = 0;
    const int gi = (int)plan->groups.size();
    for (int u = 0; u < nterms; ++u) {
      if (group_of[u] >= 0 || terms[u].X != gr.X) continue;
      if ((best_kind ? terms[u].A : terms[u].B) == gr.center) {
        group_of[u] = gi;
        ++gr.count;
        --remaining;
      }
    }
    plan->groups.push_back(gr);
  }
  int first = 0;
  for (auto& gr : plan->groups) {
    gr.first = first;
    first += gr.count;
  }
  plan->terms.resize(nterms);
  std::vector<int> fill(plan->groups.size(), 0);
  for (int t = 0; t < nterms; ++t) {
    const int gi = group_of[t];
    plan->terms[plan->groups[gi].first + fill[gi]++] = terms[t];
  }
  plan->terms_dev = nullptr;
  plan->groups_dev = nullptr;
}

int stage3_plan_create(const Stage3Term* terms, int nterms, Stage3Plan** out) {
  Stage3Plan* plan = new Stage3Plan();
  stage3_plan_host(terms, nterms, plan);
  *out = plan;
  return CARC_OK;
}

// Device copies of the term / star tables, made on the stream of the first apply.  Stream-ordered allocation
// (cudaMallocAsync / cudaFreeAsync from the default pool, whose release threshold the library raises): a sweep creates and
// drops two operators per minimisation, and cudaMalloc / cudaFree -- which synchronise the device and, with tens of GB
// mapped, took 0.06 - 0.8 s now and then (scripts/stall_probe.py) -- were the sporadic stalls of a sweep iteration.
int stage3_plan_upload(Stage3Plan* plan, cudaStream_t stream) {
  if (plan->terms.empty()) return CARC_OK;
  if (plan->terms_dev) {
    if (plan->dev_stream != stream) {
      // used on another stream before: order this stream after everything queued there (rare)
      CARC_CHECK_CUDA(cudaStreamSynchronize(plan->dev_stream));
      plan->dev_stream = stream;
    }
    return CARC_OK;
  }
  CARC_CHECK_CUDA(cudaMallocAsync((void**)&plan->terms_dev, sizeof(Stage3Term) * plan->terms.size(), stream));
  CARC_CHECK_CUDA(cudaMallocAsync((void**)&plan->groups_dev, sizeof(Stage3Group) * plan->groups.size(), stream));
  CARC_CHECK_CUDA(cudaMemcpyAsync(plan->terms_dev, plan->terms.data(), sizeof(Stage3Term) * plan->terms.size(),
                                  cudaMemcpyHostToDevice, stream));
  CARC_CHECK_CUDA(cudaMemcpyAsync(plan->groups_dev, plan->groups.data(), sizeof(Stage3Group) * plan->groups.size(),
                                  cudaMemcpyHostToDevice, stream));
  plan->dev_stream = stream;
  return CARC_OK;
}

void stage3_plan_destroy(Stage3Plan* plan) {
  if (!plan) return;
  if (plan->terms_dev) cudaFreeAsync(plan->terms_dev, plan->dev_stream);
  if (plan->groups_dev) cudaFreeAsync(plan->groups_dev, plan->dev_stream);
  delete plan;
}

// FP64 flops the fused kernel actually issues on the tensor pipe (first products: one per term; second products:
// one per group), for the roofline; the reference-equivalent count is 8 x cost_of_multiply.
double stage3_executed_flops(const Stage3Plan* plan, int P, int Q, int R, int S, int d, int column_blocks) {
  double f = 0.0;
  for (const auto& gr : plan->groups) {
    // every column block of the output (R > 64: stage3f.cu) repeats the first products
    const double first = (gr.kind ? 1.0 : (double)gr.count) * column_blocks, second = gr.kind ? (double)gr.count : 1.0;
    f += 8.0 * (double)gr.X * (first * P * Q * S * d + second * P * R * S * d);
  }
  return f;
}

int stage3_apply(const Stage3Plan* plan, int P, int Q, int R, int S, int d, const cplx* v, cplx* out, cplx* workspace,
                 int64_t workspace_elems, int force_path, cudaStream_t stream, Comm* comm) {
  const int nterms = (int)plan->terms.size();
  CARC_REQUIRE(P > 0 && Q > 0 && R > 0 && S > 0 && d > 0 && d <= 8, CARC_ERR_VALUE, "stage3: invalid dimensions");
  const int64_t n = (int64_t)P * R * d;
  if (nterms == 0) {
    CARC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(cplx) * n, stream));
    return comm ? comm_allreduce(comm, out, 1, n, out, stream) : CARC_OK;
  }
  int64_t Xmax = 0;
  for (const auto& t : plan->terms) Xmax = std::max(Xmax, t.X);
  S3Config k;
  Stage3FConfig kf;
  bool can_fuse = false, can_fold = false;
  s3_choose(nterms, P, Q, R, S, d, Xmax, force_path, &k, &kf, &can_fuse, &can_fold);
  CARC_REQUIRE(!(force_path == 1 && !can_fuse), CARC_ERR_UNSUPPORTED, "stage3: fused path unavailable for this shape");
  CARC_REQUIRE(!(force_path == 3 && !can_fold), CARC_ERR_UNSUPPORTED, "stage3: folded fused path unavailable for this shape");
  if (can_fold) {
    CARC_REQUIRE(workspace_elems >= (int64_t)kf.slots * n, CARC_ERR_VALUE, "stage3: workspace too small");
    int rc = stage3f_launch(plan, kf, P, Q, R, S, v, workspace, stream);
    if (rc) return rc;
    if (comm) {
      // slot sum in place (into slot 0: an element's eight readers and its one writer sit in the same CTA, behind a
      // barrier), then the sum over ranks reads ONE slot per rank
      s3_reduce_kernel<<<(unsigned)((n + 31) / 32), 256, 0, stream>>>(workspace, workspace, n, kf.slots, 0);
      CARC_CHECK_CUDA(cudaGetLastError());
      return comm_allreduce(comm, workspace, 1, n, out, stream);
    }
    s3_reduce_kernel<<<(unsigned)((n + 31) / 32), 256, 0, stream>>>(workspace, out, n, kf.slots, 0);
    CARC_CHECK_CUDA(cudaGetLastError());
    return CARC_OK;
  }
  if (!can_fuse) {
    int rc = s3_unfused(plan->terms.data(), nterms, P, Q, R, S, d, v, out, workspace, workspace_elems, stream);
    if (rc || !comm) return rc;
    return comm_allreduce(comm, out, 1, n, out, stream);
  }

  CARC_REQUIRE(workspace_elems >= (int64_t)k.slots * n, CARC_ERR_VALUE, "stage3: workspace too small");
  S3Params p;
  p.terms = plan->terms_dev;
  p.groups = plan->groups_dev;
  p.nterms = nterms;
  p.ngroups = (int)plan->groups.size();
  p.P = P; p.Q = Q; p.R = R; p.S = S;
  p.Q8 = k.Q8; p.NPT = k.NPT; p.G = k.G; p.NSB = k.NSB; p.NSL = k.NSL; p.nstA = k.nstA; p.nstB = k.nstB;
  p.slotA_bytes = k.slotA; p.slotB_bytes = k.slotB;
  p.ops_off = k.ops_off; p.hasop_off = k.hasop_off; p.tab_off = k.tab_off; p.vt_off = k.vt_off; p.ring_off = k.ring_off; p.smem_total = k.total;
  p.v = v;
  p.partial = workspace;
  int rc;
  switch (k.NRT) {
    case 1: rc = s3_launch<1, 2>(p, k, stream); break;
    case 2: rc = s3_launch<2, 2>(p, k, stream); break;
    case 3: rc = s3_launch<3, 2>(p, k, stream); break;
    case 4: rc = s3_launch<4, 2>(p, k, stream); break;
    case 5: rc = s3_launch<5, 1>(p, k, stream); break;
    case 6: rc = s3_launch<6, 1>(p, k, stream); break;
    case 7: rc = s3_launch<7, 2>(p, k, stream); break;
    default: rc = s3_launch<8, 2>(p, k, stream); break;
  }
  if (rc) return rc;
  // multi-GPU: the slot sum and the sum over ranks are one kernel reading the peers' exchange buffers over NVLink
  if (comm) {
    s3_reduce_kernel<<<(unsigned)((n + 31) / 32), 256, 0, stream>>>(workspace, workspace, n, k.slots, 0);
    CARC_CHECK_CUDA(cudaGetLastError());
    return comm_allreduce(comm, workspace, 1, n, out, stream);
  }
  s3_reduce_kernel<<<(unsigned)((n + 31) / 32), 256, 0, stream>>>(workspace, out, n, k.slots, 0);
  CARC_CHECK_CUDA(cudaGetLastError());
  return CARC_OK;
}

}  // namespace carc
