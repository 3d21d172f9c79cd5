// dcs_kernels.cuh — the hot kernels of the DCS-LM path (sm_100a, fp64, HBM/fp64-pipe bound; no
// dense contraction anywhere, so no tensor cores).
//
//   k_linearize   K1+K2  fused residual + analytic Jacobian + DCS + Huber + J^T J / J^T r
//                        assembly, row-owner form: one thread per pose walks the pose's
//                        half-edges (jagged-diagonal layout -> coalesced), keeps the diagonal
//                        block and gradient in registers and streams one 3x3 off-diagonal block
//                        per half-edge.  No atomics, bit-reproducible.
//   k_cost_rows   K6     cost-only evaluation at a candidate point (same row-owner walk, every edge booked once)
//   k_pcg_chain          chain-segment preconditioned PCG vector step (k_chain_factor: its factorisation)
//   k_expand             slot-order (both triangles) block storage for the SpMV from the compact upper blocks
//   k_edge_eval          per-edge r / J / psi / rho' dump in edge order (parity hook; same math)
//   k_spmv        K3     q = (H + Lambda) p over the same layout, fused p.q partial reduction
//   k_pcg_*       K4     fused PCG vector updates + dot products, scalars stay on the device
//   k_precond     K5     A_ii = H_ii + Lambda_i, M_i^-1 = A_ii^-1 (symmetric 3x3)
#pragma once
#include "dcs_common.cuh"

namespace dcs {

// scalar slots in the device scalar block
enum { S_COST = 0, S_GSQ = 1, S_GMAX = 2, S_PQ = 3, S_RZ = 4, S_RZ_NEXT = 5, S_RR = 6, S_RR0 = 7,
       S_WG = 8, S_WHW = 9, S_STEP_SQ = 10, S_XSQ = 11, S_CAND_COST = 12, S_TMP = 13 /* 13, 14 */, S_PCG_ITERS = 15 /* k_pcg_cluster: iterations run */, S_TRES = 16,
       S_SC = 17 /* 17..21: METHOD 2 step sums */, S_SC_COST = 22 /* 22: cost, 23: sum s^2 */, S_COUNT = 24 };

// Sliced-ELL layout of the owned rows (SELL-32 with a 1024-row sorting window).  Rows are ranked by (degree,
// owner entries) inside their window and STORED in that order; a warp task = 32 consecutive stored rows; the k-th
// half-edges of a task's rows form one TILE of 32 slots, and a task's tiles are consecutive:
//     slot(row position m, k) = (tile0[m / 32] + k) * 32 + (m & 31),      tile0 = exclusive scan of the tasks' max degree.
// Every per-half-edge array is indexed by slot, so a task streams ONE contiguous run per array (no round pointers, no
// shuffles, constant offsets inside a tile); slots beyond a row's degree are padding (3 % of the slots on the
// benchmark graph - address space only, inactive lanes never touch them).
struct RowLayout {
  int32_t nrows;             // owned poses (stored positions [0, nrows) are real rows)
  int32_t ntasks;            // 32-row warp tasks (= ldn / 32)
  int64_t ldn;               // leading dimension of per-row SoA arrays
  int64_t ldh;               // slots = 32 * (tiles + kTailTiles)
  int64_t ldu;               // compact owner-block capacity (multiple of 32)
  const uint4* rowinfo;      // [ldn] per stored row: (degree, word of round 0, word of round 1, local entries: the first
                             // rowinfo.w half-edges of the row have their column on this rank)
  const int2* task_info;     // [ntasks + 1] (first tile, first compact owner-block index)
};

// What k_linearize streams: 32 bytes per half-edge, one 256-bit load per lane, a tile = 1 KB contiguous.
struct __align__(32) HalfEdgeRec {
  double tmx, tmy;           // Rm^T (dx, dy), rotated once at upload
  double thm;                // measured rotation
  uint32_t word;             // other pose (local index) | flags
  uint32_t word_next;        // word of the same row kK1Rounds rounds later (0 past the row's end): the gather of the
                             // next pipeline stage needs no separate index stream
};

// Tile-interleaved 3x3 block arrays ([tile][8][32], T = double or float): the values of a lane's block sit at constant
// offsets from one address, a warp writes / reads 8 x 256 (128) contiguous bytes of one 2048 (1024)-byte tile.
// EIGHT values per block: the xy-xy part of every off-diagonal block of this problem is symmetric (H_ab = X0^T S Y0 has
// -(alpha I + c1 f f^T) there, and so has its transpose), so m10 is not stored:
//   plane 0..7 = m00, m01 (= m10), m02, m11, m12, m20, m21, m22.
// 64 instead of 72 bytes per block in k_linearize's store stream and in the SpMV's read stream.
constexpr int kBlockVals = 8;
__device__ __forceinline__ int64_t block_base(int64_t slot) { return (slot >> 5) * (kBlockVals * 32) + (slot & 31); }
__host__ __device__ constexpr int block_plane(int rc) { return rc < 3 ? rc : (rc == 3 ? 1 : rc - 1); }   // row-major (r, c) -> plane

__device__ __forceinline__ void ld_rec(HalfEdgeRec& r, const HalfEdgeRec* p, uint64_t pol) {
  uint64_t w;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b64 {%0, %1, %2, %3}, [%4], %5;"
               : "=d"(r.tmx), "=d"(r.tmy), "=d"(r.thm), "=l"(w) : "l"(p), "l"(pol));
  r.word = (uint32_t)w; r.word_next = (uint32_t)(w >> 32);
}
struct PoseRec { double x, y, th; };
__device__ __forceinline__ void ld_pose(PoseRec& r, const double4* p, uint64_t pol) {   // one 32-byte sector per gathered pose
  [[maybe_unused]] double pad;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
               : "=d"(r.x), "=d"(r.y), "=d"(r.th), "=d"(pad) : "l"(p), "l"(pol));
  (void)pad;
}

#ifdef DCS_CHECK
// -DDCS_CHECK (make check): device-side bounds asserts on every gathered / scattered index of the row-owner kernels
#define DCS_ASSERT(cond) do { if (!(cond)) { printf("DCS_CHECK failed: %s (%s:%d) block %d lane %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define DCS_ASSERT(cond) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// K1 + K2
// ------------------------------------------------------------------------------------------------
// The assembled product is the reference's structure: one 3x3 off-diagonal block per edge (written by the half-edge
// whose row is the smaller endpoint, or whose partner row lives on another rank: kFlagOwner), the diagonal blocks and
// the gradient.  The owner blocks go to a COMPACT tile-interleaved array Hup in (task, round, lane) order - a task
// writes one dense run, every sector / line is written whole (partially written lines cost DRAM the same as full
// ones) - always in the edge's own orientation H_ab = J_a^T J_b (row a, column b), whichever endpoint owns it: no
// per-value selects; consumers transpose where their row is the b endpoint.
//
// Pipeline: one warp per task, kK1Rounds rounds per stage, registers only.  ptxas tracks every long-latency load of
// the loop on one scoreboard, so the stage is in lockstep: next stage's records and gathered poses are issued at the
// top (the gather addresses come from the CURRENT records' word_next), the current stage's math runs on registers
// that hold no pending load, and the stage ends with the only scoreboard wait.  The two register sets swap roles by
// unrolling two stages (no rotation moves); prefetch tile indices are clamped to the task's last tile (no
// predication, a re-read of that tile hits L2).
#ifndef DCS_K1_WARPS
#define DCS_K1_WARPS 16
#endif
#ifndef DCS_K1_PF_OWN
#define DCS_K1_PF_OWN 1
#endif
struct K1Stage { HalfEdgeRec rec[kK1Rounds]; PoseRec pose[kK1Rounds]; };

__global__ void __launch_bounds__(kRowsPerBlock, DCS_K1_WARPS)
k_linearize(const double4* __restrict__ xyt, RowLayout L, const HalfEdgeRec* __restrict__ recs, Params P, int32_t n_loc,
            double* __restrict__ Hup, double* __restrict__ Hdiag, double* __restrict__ grad, double* __restrict__ task_part) {
  const L2Policy pol = make_l2_policy();
  constexpr int kR = kK1Rounds;
  const int task = blockIdx.x;
  if (task >= L.ntasks) return;
  const int lane = threadIdx.x & 31;
  const int lr = task * kSlice + lane;              // rows are stored in task order: coalesced row arrays
  // start-up: three independent coalesced loads, then the first records + gathers
  const uint4 info = L.rowinfo[lr];
  const int2 ti = L.task_info[task];
  PoseRec own;
  ld_pose(own, xyt + lr, pol.keep);
  const int deg = lr < L.nrows ? (int)info.x : 0;
  const int kmax = __reduce_max_sync(0xffffffffu, deg);
  const HalfEdgeRec* rp = recs + (int64_t)ti.x * kSlice + lane;     // this lane's record of round 0; round k: + k * 32
  int orun = ti.y;                                  // compact index of the task's next owner block
  // Bulk L2 prefetches (one instruction by one lane, no destination register, no scoreboard):
  //  * DCS_K1_PF_OWN: the task's own record run is ONE contiguous range (kmax KB) - start its DRAM fetch now so the
  //    pipeline's record loads find their lines in L2.  (Prefetching a whole task ahead - the row infos, own poses and
  //    record run of the task that starts when this one ends - was measured and changed nothing.)
  constexpr int kPrefetchTiles = 32;
#if DCS_K1_PF_OWN
  if (lane == 0 && kmax > kR)
    prefetch_l2_bulk(recs + ((int64_t)ti.x + kR) * kSlice, (uint32_t)(min(kmax - kR, kPrefetchTiles) * kSlice * (int)sizeof(HalfEdgeRec)));
#endif
  double d00 = 0, d01 = 0, d02 = 0, d11 = 0, d12 = 0, d22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;

  auto process = [&](const HalfEdgeRec& r, const PoseRec& pc, bool on) {
    const uint32_t word = r.word;
    const unsigned om = __ballot_sync(0xffffffffu, on && (word & kFlagOwner));
    if (on) {
      const EdgeFrame F = edge_frame(own.x, own.y, own.th, pc.x, pc.y, pc.th, word);
      const EdgeCore C = edge_core(F, r.tmx, r.tmy, r.thm, (word & kFlagDcs) != 0, P);
      EdgeTerms T;
      edge_terms(F, C, T);
      // diagonal block of the row pose and its gradient, accumulated in slot order
      d00 += T.U00; d01 += T.U01; d11 += T.U11;
      if (word & kFlagSideB) { d02 += T.sc0; d12 += T.sc1; d22 += T.alpha; g0 += T.bf0; g1 += T.bf1; g2 += T.gb; }
      else                   { d02 -= T.e0;  d12 -= T.e1;  d22 += T.k22;   g0 -= T.bf0; g1 -= T.bf1; g2 += T.ga; }
      if (word & kFlagCost) cost += T.cost;
      if (word & kFlagOwner) {   // H_ab, edge orientation
        const int64_t idx = (int64_t)orun + __popc(om & ((1u << lane) - 1u));
        DCS_ASSERT(idx >= 0 && idx < L.ldu);
        double* out = Hup + block_base(idx);
        st_stream(out + 0 * 32, -T.U00, pol.stream);
        st_stream(out + 1 * 32, -T.U01, pol.stream);
        st_stream(out + 2 * 32, -T.sc0, pol.stream);
        st_stream(out + 3 * 32, -T.U11, pol.stream);
        st_stream(out + 4 * 32, -T.sc1, pol.stream);
        st_stream(out + 5 * 32, T.e0, pol.stream);
        st_stream(out + 6 * 32, T.e1, pol.stream);
        st_stream(out + 7 * 32, T.o22, pol.stream);
      }
    }
    orun += __popc(om);
  };
  // loads of the kR rounds starting at round k into stage S; the gather indices are `words`
  auto fetch = [&](K1Stage& S, int k, const uint32_t (&words)[kR]) {
#pragma unroll
    for (int u = 0; u < kR; ++u) {
      const int kk = min(k + u, kmax - 1);          // clamp: past the task's end re-read its last tile (unused)
      ld_rec(S.rec[u], rp + (int64_t)kk * kSlice, pol.stream);
    }
#pragma unroll
    for (int u = 0; u < kR; ++u) {
      const uint32_t j = words[u] & kIdxMask;
      DCS_ASSERT((int32_t)j < n_loc);
      ld_pose(S.pose[u], xyt + j, pol.keep);
    }
  };
  auto next_words = [&](const K1Stage& S, uint32_t (&words)[kR]) {
#pragma unroll
    for (int u = 0; u < kR; ++u) words[u] = S.rec[u].word_next;
  };

  if (kmax > 0) {
    static_assert(kR == 2, "rowinfo carries the words of two rounds");
    K1Stage A, B;
    uint32_t w[kR] = {info.y, info.z};
    fetch(A, 0, w);
#pragma unroll 1
    for (int k = 0; k < kmax; k += 2 * kR) {
#if DCS_K1_PF_OWN
      if (lane == 0 && k > 0 && (k % kPrefetchTiles) == 0 && k + kR + kPrefetchTiles / 2 < kmax)    // long rows only
        prefetch_l2_bulk(recs + ((int64_t)ti.x + k + kR + kPrefetchTiles / 2) * kSlice,
                         (uint32_t)(min(kmax - (k + kR + kPrefetchTiles / 2), kPrefetchTiles) * kSlice * (int)sizeof(HalfEdgeRec)));
#endif
      next_words(A, w);
      fetch(B, k + kR, w);
#pragma unroll
      for (int u = 0; u < kR; ++u) process(A.rec[u], A.pose[u], k + u < deg);
      if (k + kR >= kmax) break;
      next_words(B, w);
      fetch(A, k + 2 * kR, w);
#pragma unroll
      for (int u = 0; u < kR; ++u) process(B.rec[u], B.pose[u], k + kR + u < deg);
    }
  }
  if (lr < L.nrows) {
    Hdiag[0 * L.ldn + lr] = d00; Hdiag[1 * L.ldn + lr] = d01; Hdiag[2 * L.ldn + lr] = d02;
    Hdiag[3 * L.ldn + lr] = d11; Hdiag[4 * L.ldn + lr] = d12; Hdiag[5 * L.ldn + lr] = d22;
    grad[0 * L.ldn + lr] = g0; grad[1 * L.ldn + lr] = g1; grad[2 * L.ldn + lr] = g2;
  }
  // per-task partials (no fence, no ticket: a __threadfence per 32-row task cost ~7 k cycles of its ~40 k);
  // k_fold_tasks adds them in task order
  double r0 = cost, r1 = fma(g0, g0, fma(g1, g1, g2 * g2)), r2 = fmax(fabs(g0), fmax(fabs(g1), fabs(g2)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r0 += __shfl_xor_sync(0xffffffffu, r0, o);
    r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  }
  if (lane == 0) {
    const size_t n = L.ntasks;
    task_part[task] = r0; task_part[n + task] = r1; task_part[2 * n + task] = r2;
  }
}

// Deterministic fold of per-task partials (K sums then M maxima, each n long).  Up to 32 CTAs fold contiguous
// chunks (all K+M loads of a thread are independent), the CTA that arrives last folds the per-CTA partials with a
// fixed shuffle tree, so the result does not depend on arrival order.  ws: (K+M)*32 doubles, ticket: one zeroed uint.
// `rotate_rz`: S_RZ <- S_RZ_NEXT (PCG scalar rotation, see k_pcg_direction).
constexpr int kFoldThreads = 256;
constexpr int kFoldMaxBlocks = 32;
inline int fold_blocks(int n) { return n <= 4 * kFoldThreads ? 1 : min(kFoldMaxBlocks, (n + 4 * kFoldThreads - 1) / (4 * kFoldThreads)); }
template <int K, int M>
__global__ void __launch_bounds__(kFoldThreads)
k_fold_tasks(const double* __restrict__ part, int n, double* out, double* scal, int rotate_rz, double* ws, unsigned int* ticket) {
  __shared__ double s_red[K + M][kFoldThreads / 32];
  __shared__ unsigned int s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nb = gridDim.x;
  const int chunk = (n + nb - 1) / nb, lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  double a[K + M];
#pragma unroll
  for (int k = 0; k < K + M; ++k) a[k] = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += kFoldThreads) {
#pragma unroll
    for (int k = 0; k < K + M; ++k) {
      const double y = part[(size_t)k * n + i];
      a[k] = (k < K) ? a[k] + y : fmax(a[k], y);
    }
  }
#pragma unroll
  for (int k = 0; k < K + M; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, a[k], o);
      a[k] = (k < K) ? a[k] + y : fmax(a[k], y);
    }
    if (lane == 0) s_red[k][wid] = a[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < K + M; ++k) {
      double t = lane < kFoldThreads / 32 ? s_red[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, t, o);
        t = (k < K) ? t + y : fmax(t, y);
      }
      a[k] = t;
    }
    if (nb == 1) {
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K + M; ++k) out[k] = a[k];
        if (rotate_rz) scal[S_RZ] = scal[S_RZ_NEXT];
      }
      return;
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K + M; ++k) ws[k * kFoldMaxBlocks + blockIdx.x] = a[k];
      __threadfence();
      s_last = (atomicAdd(ticket, 1u) == (unsigned)nb - 1u);
    }
    __syncwarp();
    if (s_last) {
      __threadfence();
#pragma unroll
      for (int k = 0; k < K + M; ++k) {
        double t = lane < nb ? __ldcg(ws + k * kFoldMaxBlocks + lane) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double y = __shfl_xor_sync(0xffffffffu, t, o);
          t = (k < K) ? t + y : fmax(t, y);
        }
        if (lane == 0) out[k] = t;
      }
      if (lane == 0) {
        *ticket = 0u;   // re-arm for the next launch
        if (rotate_rz) scal[S_RZ] = scal[S_RZ_NEXT];
      }
    }
  }
}

// Pattern build, one warp per task walking its tiles exactly as k_linearize does.
//   kWrite = false: own_cnt[task] = owner half-edges of the task (scanned into the compact base of every task)
//   kWrite = true : cidx[slot] = compact owner-block index in (task, round, lane) order; the records' word_next
//                   (word of the same row kK1Rounds rounds later); rowinfo (degree + the words of rounds 0 and 1)
template <bool kWrite>
__global__ void __launch_bounds__(kRowsPerBlock)
k_task_walk(int32_t ntasks, const uint32_t* __restrict__ rank_info, const int32_t* __restrict__ rank_nloc, const int32_t* __restrict__ tile0,
            const int32_t* __restrict__ obase, const uint32_t* __restrict__ cols, int32_t* own_cnt, int32_t* cidx, HalfEdgeRec* recs, uint4* rowinfo) {
  const int task = blockIdx.x, lane = threadIdx.x & 31;
  if (task >= ntasks) return;
  const int m = task * kSlice + lane;
  const int deg = (int)(rank_info[m] >> 10);
  const int kmax = __reduce_max_sync(0xffffffffu, deg);
  const int64_t s0 = (int64_t)tile0[task] * kSlice + lane;
  int run = kWrite ? obase[task] : 0;
  for (int k = 0; k < kmax; ++k) {
    const int64_t s = s0 + (int64_t)k * kSlice;
    const bool own = k < deg && (cols[s] & kFlagOwner);
    const unsigned om = __ballot_sync(0xffffffffu, own);
    if (kWrite) {
      if (own) cidx[s] = run + __popc(om & ((1u << lane) - 1u));
      if (k < deg) recs[s].word_next = (k + kK1Rounds < deg) ? cols[s + (int64_t)kK1Rounds * kSlice] : 0u;
    }
    run += __popc(om);
  }
  if (kWrite) rowinfo[m] = make_uint4((uint32_t)deg, deg > 0 ? cols[s0] : 0u, deg > 1 ? cols[s0 + kSlice] : 0u, (uint32_t)rank_nloc[m]);
  else if (lane == 0) own_cnt[task] = run;
}
__global__ void k_task_kmax(int32_t ntasks, const uint32_t* __restrict__ rank_info, int32_t* kmax) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntasks) kmax[t] = (int32_t)(rank_info[(int64_t)t * kSlice] >> 10);     // ranks are degree-sorted: the first lane has the most
}
__global__ void k_task_info(int32_t ntasks, const int32_t* __restrict__ tile0, const int32_t* __restrict__ obase, int2* task_info) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= ntasks) task_info[t] = make_int2(tile0[t], obase[t]);
}

// block_src[slot]: compact index of the slot's edge block (its own if the slot is an owner, its partner's otherwise),
// -1 when the edge has no off-diagonal block (other endpoint constant)
__global__ void k_block_src(const uint32_t* __restrict__ cols, const int32_t* __restrict__ slot, const int32_t* __restrict__ mirror_src,
                            const int32_t* __restrict__ cidx, int32_t nh, int32_t* block_src) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int32_t s = slot[i];
  int32_t src = -1;
  if (cols[s] & kFlagOwner) src = cidx[s];
  else if (mirror_src[s] >= 0) src = cidx[mirror_src[s]];
  block_src[s] = src;
}

// Linear-solver setup: the SpMV walks full rows in slot order, so every slot receives its block from the compact
// array of edge blocks H_ab: slots whose row is the a endpoint a copy, b-endpoint slots the transpose, the rest
// (padding, constant partner) zero.  T = double (the operator of the linear system) or float (its single-precision
// shadow for the mixed-precision inner iterations).
template <typename T>
__global__ void __launch_bounds__(256)
k_expand(const int32_t* __restrict__ block_src, const uint32_t* __restrict__ cols, int64_t nslots, const double* __restrict__ Hup,
         T* __restrict__ Hoff) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots) return;
  const int32_t s = block_src[i];
  double v[kBlockVals];
#pragma unroll
  for (int c = 0; c < kBlockVals; ++c) v[c] = 0.0;
  if (s >= 0) {
    const double* in = Hup + block_base(s);
#pragma unroll
    for (int c = 0; c < kBlockVals; ++c) v[c] = in[c * 32];
  }
  T* out = Hoff + block_base(i);
  if (!(cols[i] & kFlagSideB)) {
#pragma unroll
    for (int c = 0; c < kBlockVals; ++c) out[c * 32] = (T)v[c];
  } else {      // transpose: (m00, m01, m02, m11, m12, m20, m21, m22) -> (m00, m01, m20, m11, m21, m02, m12, m22)
    out[0 * 32] = (T)v[0]; out[1 * 32] = (T)v[1]; out[2 * 32] = (T)v[5]; out[3 * 32] = (T)v[3];
    out[4 * 32] = (T)v[6]; out[5 * 32] = (T)v[2]; out[6 * 32] = (T)v[4]; out[7 * 32] = (T)v[7];
  }
}

// ------------------------------------------------------------------------------------------------
// Edge-order SoA (cost-only evaluation, parity dump)
// ------------------------------------------------------------------------------------------------
struct EdgeList {
  int32_t n;
  const int32_t* a;
  const int32_t* b;
  const double* tmx;
  const double* tmy;
  const double* thm;
  const uint8_t* dcs;   // 1 when the DCS functor applies to this edge
};

constexpr int kEdgeThreads = 256;

// K6: cost at a candidate point over the rank's own rows: every edge is booked on exactly one half-edge
// (kFlagCost), so each rank needs only its own + halo poses and the ranks' partial costs add up.
__global__ void __launch_bounds__(kRowsPerBlock)
k_cost_rows(const double4* __restrict__ xyt, RowLayout L, const HalfEdgeRec* __restrict__ recs, Params P, int32_t n_loc,
            double* __restrict__ task_part) {
  const L2Policy pol = make_l2_policy();
  const int task = blockIdx.x, lane = threadIdx.x & 31;
  if (task >= L.ntasks) return;
  const int lr = task * kSlice + lane;
  double cost = 0.0;
  if (lr < L.nrows) {
    const int deg = (int)L.rowinfo[lr].x;
    PoseRec own;
    ld_pose(own, xyt + lr, pol.keep);
    const HalfEdgeRec* rp = recs + (int64_t)L.task_info[task].x * kSlice + lane;
    constexpr int U = 2;
    int k = 0;
    for (; k + U <= deg; k += U) {
      HalfEdgeRec r[U]; PoseRec pj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ld_rec(r[u], rp + (int64_t)(k + u) * kSlice, pol.stream);
#pragma unroll
      for (int u = 0; u < U; ++u) { DCS_ASSERT((int32_t)(r[u].word & kIdxMask) < n_loc); ld_pose(pj[u], xyt + (r[u].word & kIdxMask), pol.keep); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!(r[u].word & kFlagCost)) continue;
        const EdgeFrame F = edge_frame(own.x, own.y, own.th, pj[u].x, pj[u].y, pj[u].th, r[u].word);
        cost += edge_core(F, r[u].tmx, r[u].tmy, r[u].thm, (r[u].word & kFlagDcs) != 0, P).cost;
      }
    }
    for (; k < deg; ++k) {
      HalfEdgeRec r; PoseRec pj;
      ld_rec(r, rp + (int64_t)k * kSlice, pol.stream);
      if (!(r.word & kFlagCost)) continue;
      ld_pose(pj, xyt + (r.word & kIdxMask), pol.keep);
      const EdgeFrame F = edge_frame(own.x, own.y, own.th, pj.x, pj.y, pj.th, r.word);
      cost += edge_core(F, r.tmx, r.tmy, r.thm, (r.word & kFlagDcs) != 0, P).cost;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
  if (lane == 0) task_part[task] = cost;
}

__global__ void __launch_bounds__(kEdgeThreads)
k_edge_eval(const double4* __restrict__ xyt, const int32_t* __restrict__ g2l, EdgeList E, Params P, double* res, double* jac,
            double* psi, double* rho1) {
  const int64_t e = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  if (e >= E.n) return;
  const int32_t a = E.a[e], b = E.b[e];
  const double4 pa = xyt[g2l[a]], pb = xyt[g2l[b]];      // global pose id -> position in the stored pose array
  EdgeLin L;
  edge_linearize<true>(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, E.tmx[e], E.tmy[e], E.thm[e], E.dcs[e] != 0, P, L);
  if (res) { res[3 * e] = L.r0; res[3 * e + 1] = L.r1; res[3 * e + 2] = L.r2; }
  if (jac) {
    double* J = jac + 18 * e;
    J[0] = L.a00; J[1] = L.a01; J[2] = L.a02; J[3] = L.b00; J[4] = L.b01; J[5] = 0.0;
    J[6] = L.a10; J[7] = L.a11; J[8] = L.a12; J[9] = L.b10; J[10] = L.b11; J[11] = 0.0;
    J[12] = L.a20; J[13] = L.a21; J[14] = L.a22; J[15] = L.b20; J[16] = L.b21; J[17] = L.b22;
  }
  if (psi) psi[e] = L.psi;
  if (rho1) rho1[e] = L.rho1;
}

// ------------------------------------------------------------------------------------------------
// pose upload helpers
// ------------------------------------------------------------------------------------------------
// Device poses live in the rank's LOCAL index space: [own rows | halo poses in global order].  32-byte records: one
// sector per gather; the whole gathered working set (own + halo) is contiguous and stays L2 resident.
__global__ void k_pack_poses(const double* __restrict__ xyt3, int32_t n_global, int32_t row_lo, int32_t rows_per_rank,
                             const int32_t* __restrict__ halo_idx, const uint16_t* __restrict__ rank_of, int32_t n_loc, double4* xyt) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loc) return;
  const int32_t j = i < rows_per_rank ? row_lo + i : halo_idx[i - rows_per_rank];
  const int32_t dst = i < rows_per_rank ? row_pos(rank_of, i) : i;
  xyt[dst] = j < n_global ? make_double4(xyt3[3 * (int64_t)j], xyt3[3 * (int64_t)j + 1], xyt3[3 * (int64_t)j + 2], 0.0)
                        : make_double4(0, 0, 0, 0);
}
// own rows -> global N x 3 staging
__global__ void k_unpack_poses(const double4* __restrict__ xyt, const uint16_t* __restrict__ rank_of, int32_t row_lo, int32_t rows,
                               int32_t n_global, double* xyt3) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows || row_lo + i >= n_global) return;
  const double4 p = xyt[row_pos(rank_of, i)];
  const int64_t j = row_lo + i;
  xyt3[3 * j] = p.x; xyt3[3 * j + 1] = p.y; xyt3[3 * j + 2] = p.z;
}

// ------------------------------------------------------------------------------------------------
// LM diagonal / preconditioner (per owned row; natural order)
// ------------------------------------------------------------------------------------------------
constexpr int kVecThreads = 256;

// scale_c = 1 / (1 + sqrt(H_cc))          (Ceres Jacobi scaling, computed once at iteration 0)
__global__ void k_jacobi_scale(const double* __restrict__ Hdiag, int32_t nrows, int64_t ldn, double* scale, int enabled) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  scale[0 * ldn + r] = enabled ? 1.0 / (1.0 + sqrt(Hdiag[0 * ldn + r])) : 1.0;
  scale[1 * ldn + r] = enabled ? 1.0 / (1.0 + sqrt(Hdiag[3 * ldn + r])) : 1.0;
  scale[2 * ldn + r] = enabled ? 1.0 / (1.0 + sqrt(Hdiag[5 * ldn + r])) : 1.0;
}
// lmdiag_c = clamp(scale_c^2 H_cc, min, max)   (LevenbergMarquardtStrategy, !reuse_diagonal)
__global__ void k_lm_diagonal(const double* __restrict__ Hdiag, const double* __restrict__ scale, int32_t nrows, int64_t ldn,
                              double dmin, double dmax, double* lmdiag) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int hd[3] = {0, 3, 5};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double s = scale[c * ldn + r];
    lmdiag[c * ldn + r] = fmin(fmax(s * s * Hdiag[hd[c] * ldn + r], dmin), dmax);
  }
}
// A_ii = H_ii + diag(lambda), lambda_c = lmdiag_c / (radius scale_c^2)  [or explicit lambda];  Minv = A_ii^-1.
// Rows that are not parameters (constant / untouched) get identity.
__global__ void k_precond(const double* __restrict__ Hdiag, const double* __restrict__ lmdiag, const double* __restrict__ scale,
                          const uint8_t* __restrict__ is_free, int32_t nrows, int64_t ldn, double inv_radius,
                          const double* __restrict__ lambda_explicit, double* Adiag, double* Minv) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double a00 = 1, a01 = 0, a02 = 0, a11 = 1, a12 = 0, a22 = 1;
  if (is_free[r]) {
    double l0, l1, l2;
    if (lambda_explicit) { l0 = lambda_explicit[0 * ldn + r]; l1 = lambda_explicit[1 * ldn + r]; l2 = lambda_explicit[2 * ldn + r]; }
    else {
      const double s0 = scale[0 * ldn + r], s1 = scale[1 * ldn + r], s2 = scale[2 * ldn + r];
      l0 = lmdiag[0 * ldn + r] * inv_radius / (s0 * s0);
      l1 = lmdiag[1 * ldn + r] * inv_radius / (s1 * s1);
      l2 = lmdiag[2 * ldn + r] * inv_radius / (s2 * s2);
    }
    a00 = Hdiag[0 * ldn + r] + l0; a01 = Hdiag[1 * ldn + r]; a02 = Hdiag[2 * ldn + r];
    a11 = Hdiag[3 * ldn + r] + l1; a12 = Hdiag[4 * ldn + r]; a22 = Hdiag[5 * ldn + r] + l2;
  }
  Adiag[0 * ldn + r] = a00; Adiag[1 * ldn + r] = a01; Adiag[2 * ldn + r] = a02;
  Adiag[3 * ldn + r] = a11; Adiag[4 * ldn + r] = a12; Adiag[5 * ldn + r] = a22;
  const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  Minv[0 * ldn + r] = c00 * id; Minv[1 * ldn + r] = c01 * id; Minv[2 * ldn + r] = c02 * id;
  Minv[3 * ldn + r] = (a00 * a22 - a02 * a02) * id; Minv[4 * ldn + r] = (a01 * a02 - a00 * a12) * id;
  Minv[5 * ldn + r] = (a00 * a11 - a01 * a01) * id;
}

// ------------------------------------------------------------------------------------------------
// K3: q = A p (A = diagonal blocks `D` + the slot-order off-diagonal blocks), fused p.q; thread per row over the
// SELL tiles: per round a warp reads 128 B of column words and 9 x 256 B (fp64) / 9 x 128 B (fp32 shadow) of one
// block tile.  T = double: the operator itself; T = float: mixed-precision inner iterations (fp64 vectors and
// accumulation, single-precision off-diagonal blocks).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_blockval(const double* p, uint64_t pol) { return ld_stream(p, pol); }
__device__ __forceinline__ double ld_blockval(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return (double)v;
}
// Multi-rank: a row's half-edges are stored local columns first, halo columns last (kKeyHalo), so the product runs
// in two passes of the same kernel over disjoint round ranges: kSpmvLocal (diagonal block + rounds [0, nloc), needs no
// halo entry of p: it runs while the peers' pushes are in flight) and kSpmvHalo (rounds [nloc, degree) added to q, then
// the p.q partial).  Single rank: kSpmvAll.
enum { kSpmvAll = 0, kSpmvLocal = 1, kSpmvHalo = 2 };
template <typename T, int kPass>
__global__ void __launch_bounds__(kRowsPerBlock)
k_spmv(const double4* __restrict__ p4, RowLayout L, const uint32_t* __restrict__ cols, const T* __restrict__ Hoff,
       const double* __restrict__ D, int32_t n_loc, double* __restrict__ q, double* __restrict__ task_part) {
  const L2Policy pol = make_l2_policy();
  const int task = blockIdx.x, lane = threadIdx.x & 31;
  if (task >= L.ntasks) return;
  const int lr = task * kSlice + lane;
  double y0 = 0, y1 = 0, y2 = 0, dot = 0;
  if (lr < L.nrows) {
    const uint4 info = L.rowinfo[lr];
    const int deg = kPass == kSpmvLocal ? (int)info.w : (int)info.x;
    const int64_t tile0 = L.task_info[task].x;
    const double4 p = ld_keep4(p4 + lr, pol.keep);
    if (kPass == kSpmvHalo) {
      y0 = q[0 * L.ldn + lr]; y1 = q[1 * L.ldn + lr]; y2 = q[2 * L.ldn + lr];
    } else {
      const double a00 = D[0 * L.ldn + lr], a01 = D[1 * L.ldn + lr], a02 = D[2 * L.ldn + lr];
      const double a11 = D[3 * L.ldn + lr], a12 = D[4 * L.ldn + lr], a22 = D[5 * L.ldn + lr];
      y0 = fma(a00, p.x, fma(a01, p.y, a02 * p.z));
      y1 = fma(a01, p.x, fma(a11, p.y, a12 * p.z));
      y2 = fma(a02, p.x, fma(a12, p.y, a22 * p.z));
    }
    const uint32_t* cp = cols + tile0 * kSlice + lane;       // round k: + 32 k
    const T* hp = Hoff + tile0 * (kBlockVals * 32) + lane;   // round k: + 256 k, value c: + 32 c
    constexpr int U = 4;     // rounds in flight per thread: 4 x (9 block words + column + gathered p) loads
    int k = kPass == kSpmvHalo ? (int)info.w : 0;
    for (; k + U <= deg; k += U) {
      uint32_t j[U]; double h[U][kBlockVals]; double4 pj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) j[u] = ld_stream_u32(cp + (int64_t)(k + u) * kSlice, pol.stream) & kIdxMask;
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < kBlockVals; ++c) h[u][c] = ld_blockval(hp + (int64_t)(k + u) * (kBlockVals * 32) + c * 32, pol.stream);
#pragma unroll
      for (int u = 0; u < U; ++u) { DCS_ASSERT((int32_t)j[u] < n_loc); pj[u] = ld_keep4(p4 + j[u], pol.keep); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        y0 = fma(h[u][0], pj[u].x, fma(h[u][1], pj[u].y, fma(h[u][2], pj[u].z, y0)));
        y1 = fma(h[u][1], pj[u].x, fma(h[u][3], pj[u].y, fma(h[u][4], pj[u].z, y1)));
        y2 = fma(h[u][5], pj[u].x, fma(h[u][6], pj[u].y, fma(h[u][7], pj[u].z, y2)));
      }
    }
    for (; k < deg; ++k) {
      const uint32_t j = ld_stream_u32(cp + (int64_t)k * kSlice, pol.stream) & kIdxMask;
      double h[kBlockVals];
#pragma unroll
      for (int c = 0; c < kBlockVals; ++c) h[c] = ld_blockval(hp + (int64_t)k * (kBlockVals * 32) + c * 32, pol.stream);
      DCS_ASSERT((int32_t)j < n_loc);
      const double4 pj = ld_keep4(p4 + j, pol.keep);
      y0 = fma(h[0], pj.x, fma(h[1], pj.y, fma(h[2], pj.z, y0)));
      y1 = fma(h[1], pj.x, fma(h[3], pj.y, fma(h[4], pj.z, y1)));
      y2 = fma(h[5], pj.x, fma(h[6], pj.y, fma(h[7], pj.z, y2)));
    }
    q[0 * L.ldn + lr] = y0; q[1 * L.ldn + lr] = y1; q[2 * L.ldn + lr] = y2;
    dot = fma(p.x, y0, fma(p.y, y1, p.z * y2));
  }
  if constexpr (kPass != kSpmvLocal) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) task_part[task] = dot;
  }
}

// ------------------------------------------------------------------------------------------------
// K4: PCG vector kernels (thread per owned row, natural order, SoA vectors; p is double4 AoS
// because it is the gathered operand of the SpMV)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply_minv(const double* Minv, int64_t ldn, int32_t r, double r0, double r1, double r2,
                                           double& z0, double& z1, double& z2) {
  const double m00 = Minv[0 * ldn + r], m01 = Minv[1 * ldn + r], m02 = Minv[2 * ldn + r];
  const double m11 = Minv[3 * ldn + r], m12 = Minv[4 * ldn + r], m22 = Minv[5 * ldn + r];
  z0 = fma(m00, r0, fma(m01, r1, m02 * r2));
  z1 = fma(m01, r0, fma(m11, r1, m12 * r2));
  z2 = fma(m02, r0, fma(m12, r1, m22 * r2));
}

// w = 0, r = rhs (masked to parameter rows), z = Minv r, p = z; S_RZ = S_RZ_NEXT = r.z, S_RR = S_RR0 = r.r
__global__ void __launch_bounds__(kVecThreads)
k_pcg_init(const double* __restrict__ rhs, const double* __restrict__ Minv, const uint8_t* __restrict__ is_free,
           int32_t row_lo, int32_t nrows, int64_t ldn, double* w, double* r, double* z, double4* p4,
           double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  double rz = 0, rr = 0;
  if (i < nrows) {
    const bool f = is_free[i] != 0;
    const double r0 = f ? rhs[0 * ldn + i] : 0.0, r1 = f ? rhs[1 * ldn + i] : 0.0, r2 = f ? rhs[2 * ldn + i] : 0.0;
    double z0, z1, z2;
    apply_minv(Minv, ldn, i, r0, r1, r2, z0, z1, z2);
    w[0 * ldn + i] = 0; w[1 * ldn + i] = 0; w[2 * ldn + i] = 0;
    r[0 * ldn + i] = r0; r[1 * ldn + i] = r1; r[2 * ldn + i] = r2;
    z[0 * ldn + i] = z0; z[1 * ldn + i] = z1; z[2 * ldn + i] = z2;
    p4[row_lo + i] = make_double4(z0, z1, z2, 0.0);
    rz = fma(r0, z0, fma(r1, z1, r2 * z2));
    rr = fma(r0, r0, fma(r1, r1, r2 * r2));
  }
  double s[2] = {rz, rr};
  grid_reduce_sum<2, kVecThreads>(s, partials, ticket, scal + S_TMP);   // S_TMP, S_TMP+1
}
__global__ void k_pcg_init_finish(double* scal) {
  scal[S_RZ] = scal[S_TMP]; scal[S_RZ_NEXT] = scal[S_TMP];
  scal[S_RR] = scal[S_TMP + 1]; scal[S_RR0] = scal[S_TMP + 1];
}

// alpha = rz / pq;  w += alpha p;  r -= alpha q;  z = Minv r;  S_RZ_NEXT = r.z, S_RR = r.r
__global__ void __launch_bounds__(kVecThreads)
k_pcg_update(const double4* __restrict__ p4, const double* __restrict__ q, const double* __restrict__ Minv,
             int32_t row_lo, int32_t nrows, int64_t ldn, double* w, double* r, double* z,
             double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  const double pq = scal[S_PQ];
  const double alpha = (pq != 0.0) ? scal[S_RZ] / pq : 0.0;
  double rz = 0, rr = 0;
  if (i < nrows) {
    const double4 p = p4[row_lo + i];
    const double r0 = fma(-alpha, q[0 * ldn + i], r[0 * ldn + i]);
    const double r1 = fma(-alpha, q[1 * ldn + i], r[1 * ldn + i]);
    const double r2 = fma(-alpha, q[2 * ldn + i], r[2 * ldn + i]);
    w[0 * ldn + i] = fma(alpha, p.x, w[0 * ldn + i]);
    w[1 * ldn + i] = fma(alpha, p.y, w[1 * ldn + i]);
    w[2 * ldn + i] = fma(alpha, p.z, w[2 * ldn + i]);
    r[0 * ldn + i] = r0; r[1 * ldn + i] = r1; r[2 * ldn + i] = r2;
    double z0, z1, z2;
    apply_minv(Minv, ldn, i, r0, r1, r2, z0, z1, z2);
    z[0 * ldn + i] = z0; z[1 * ldn + i] = z1; z[2 * ldn + i] = z2;
    rz = fma(r0, z0, fma(r1, z1, r2 * z2));
    rr = fma(r0, r0, fma(r1, r1, r2 * r2));
  }
  double s[2] = {rz, rr};
  grid_reduce_sum<2, kVecThreads>(s, partials, ticket, scal + S_TMP);
}
// beta = rz_next / rz;  p = z + beta p.   Also publishes S_RZ_NEXT / S_RR from S_TMP.
// (S_TMP was written by the previous kernel; S_RZ is rotated by the next SpMV.)
__global__ void __launch_bounds__(kVecThreads)
k_pcg_direction(const double* __restrict__ z, int32_t row_lo, int32_t nrows, int64_t ldn, double4* p4, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  const double rz_next = scal[S_TMP], rz = scal[S_RZ];
  const double beta = (rz != 0.0) ? rz_next / rz : 0.0;
  if (i < nrows) {
    const double4 p = p4[row_lo + i];
    p4[row_lo + i] = make_double4(fma(beta, p.x, z[0 * ldn + i]), fma(beta, p.y, z[1 * ldn + i]),
                                  fma(beta, p.z, z[2 * ldn + i]), 0.0);
  }
  if (i == 0) { scal[S_RZ_NEXT] = rz_next; scal[S_RR] = scal[S_TMP + 1]; }
}

// ------------------------------------------------------------------------------------------------
// Chain-segment preconditioner (options.preconditioner = 1).  Block-Jacobi with bigger blocks: the rows are cut
// into segments of 32 consecutive poses and M keeps, per segment, the diagonal 3x3 blocks and the blocks between
// consecutive poses (the odometry chain) - a block-tridiagonal SPD matrix that is factorised exactly,
// M = L S L^T (L unit lower block-bidiagonal), once per LM iteration.  It removes the stiff along-the-chain
// coupling that 3x3 block-Jacobi cannot see: 2.3-2.6x fewer PCG iterations on the Manhattan benchmark graphs.
// Layout: one warp per tile of 32 segments (1024 poses), lane = segment; factors are stored step-major
// ([tile][step][lane]) so every sequential step is one coalesced access; the right-hand side is transposed
// through (padded) shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kChainSeg = 32;                 // poses per segment
constexpr int kChainTile = 32 * kChainSeg;    // poses per warp tile

__device__ __forceinline__ void sym3_inverse(double a00, double a01, double a02, double a11, double a12, double a22,
                                             double& i00, double& i01, double& i02, double& i11, double& i12, double& i22) {
  const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double id = 1.0 / (a00 * c00 + a01 * c01 + a02 * c02);
  i00 = c00 * id; i01 = c01 * id; i02 = c02 * id;
  i11 = (a00 * a22 - a02 * a02) * id; i12 = (a01 * a02 - a00 * a12) * id; i22 = (a00 * a11 - a01 * a01) * id;
}

// factorisation: S_0 = D_0; L_j = E_{j-1}^T S_{j-1}^-1; S_j = D_j - L_j E_{j-1}   (E_{j-1} = A[j-1, j])
template <typename T>
__global__ void __launch_bounds__(32)
k_chain_factor(const double* __restrict__ Adiag, const T* __restrict__ Hoff, const int32_t* __restrict__ slot,
               const int32_t* __restrict__ chain_idx, const int32_t* __restrict__ chain_cnt, const uint16_t* __restrict__ rank_of,
               int32_t nrows, int64_t ldn, int64_t ldh, float* __restrict__ chL, float* __restrict__ chS) {
  const int lane = threadIdx.x;
  const int64_t tile0 = (int64_t)blockIdx.x * kChainTile;
  double s00 = 1, s01 = 0, s02 = 0, s11 = 1, s12 = 0, s22 = 1;      // S_{j-1}^-1
  for (int j = 0; j < kChainSeg; ++j) {
    const int64_t row = tile0 + (int64_t)lane * kChainSeg + j;
    const int64_t tr = tile0 + (int64_t)j * 32 + lane;
    double d00 = 1, d01 = 0, d02 = 0, d11 = 1, d12 = 0, d22 = 1;      // padding rows: identity
    if (row < nrows) {            // the chain follows the natural pose order; Adiag is stored in (window, rank) order
      const int64_t m = row_pos(rank_of, (int32_t)row);
      d00 = Adiag[0 * ldn + m]; d01 = Adiag[1 * ldn + m]; d02 = Adiag[2 * ldn + m];
      d11 = Adiag[3 * ldn + m]; d12 = Adiag[4 * ldn + m]; d22 = Adiag[5 * ldn + m];
    }
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (j > 0 && row < nrows) {
      double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const int32_t ci = chain_idx[row - 1], cc = chain_cnt[row - 1];
      for (int32_t d = 0; d < cc; ++d) {
        const int64_t sl = slot[ci + d];
#pragma unroll
        for (int c = 0; c < 9; ++c) E[c] += (double)Hoff[block_base(sl) + block_plane(c) * 32];
      }
      // L = E^T Sinv   (Sinv symmetric)
      const double S[9] = {s00, s01, s02, s01, s11, s12, s02, s12, s22};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) L[3 * a + b] = E[0 + a] * S[0 + b] + E[3 + a] * S[3 + b] + E[6 + a] * S[6 + b];
      // S_j = D_j - L E   (symmetric)
      d00 -= L[0] * E[0] + L[1] * E[3] + L[2] * E[6];
      d01 -= L[0] * E[1] + L[1] * E[4] + L[2] * E[7];
      d02 -= L[0] * E[2] + L[1] * E[5] + L[2] * E[8];
      d11 -= L[3] * E[1] + L[4] * E[4] + L[5] * E[7];
      d12 -= L[3] * E[2] + L[4] * E[5] + L[5] * E[8];
      d22 -= L[6] * E[2] + L[7] * E[5] + L[8] * E[8];
    }
    sym3_inverse(d00, d01, d02, d11, d12, d22, s00, s01, s02, s11, s12, s22);
#pragma unroll
    for (int c = 0; c < 9; ++c) chL[(int64_t)c * ldn + tr] = (float)L[c];
    // the factors are stored in fp32 (half the preconditioner's traffic): M only has to be a fixed SPD operator,
    // PCG still converges to the fp64 solution of A w = g.  The recursion itself continues in fp64.
    chS[0 * ldn + tr] = (float)s00; chS[1 * ldn + tr] = (float)s01; chS[2 * ldn + tr] = (float)s02;
    chS[3 * ldn + tr] = (float)s11; chS[4 * ldn + tr] = (float)s12; chS[5 * ldn + tr] = (float)s22;
  }
}

// One PCG vector step with the chain preconditioner (replaces k_pcg_init / k_pcg_update):
//   init  : w = 0, r = rhs (masked), z = M^-1 r, p = z
//   update: alpha = rz / pq; w += alpha p; r -= alpha q; z = M^-1 r
// per-tile partials (r.z, r.r) go to task_part[0..ntiles), [ntiles..2 ntiles); k_fold_tasks<2,0> adds them.
// One CTA of kChainThreads per 1024-pose tile, four rows per thread (128 registers: all 52 vector loads of a thread in
// flight at once).  The vector update (phase 1) and the output (phase 3) are one round trip each; the tile's factors go
// to shared memory with cp.async (no registers, overlapped with phase 1), so the sequential forward / backward
// substitution (warp 0, lane = segment, 2 x 32 dependent steps) reads shared memory only.  Two CTAs per SM: while one
// tile substitutes, the other streams.
// (Round 1: one warp per tile with register batches of 8 steps - 16 dependent memory round trips per tile, every tile
// of the single wave in the same phase at the same time: 68 us at 1 M poses, 35 us per iteration on a 1228-pose graph.)
constexpr int kChainThreads = 256;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
constexpr int kChainRowsPerThread = kChainTile / kChainThreads;
constexpr size_t kChainSmemBytes = 3 * (kChainTile + 32) * sizeof(double) + 15 * kChainTile * sizeof(float) + 2 * 32 * sizeof(double);
template <bool kInit>
__global__ void __launch_bounds__(kChainThreads, 2)
k_pcg_chain(const double* __restrict__ rhs, const uint8_t* __restrict__ is_free, const double4* __restrict__ p4r,
            const double* __restrict__ q, const float* __restrict__ chL, const float* __restrict__ chS,
            const uint16_t* __restrict__ perm, int32_t row_lo,
            int32_t nrows, int64_t ldn, double* w, double* r, double* z, double4* p4w, double* task_part, const double* scal) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  double (*s_v)[kChainTile + 32] = reinterpret_cast<double (*)[kChainTile + 32]>(s_raw);      // index n + n/32: conflict-free segment walk
  float (*s_L)[kChainTile] = reinterpret_cast<float (*)[kChainTile]>(s_raw + 3 * (kChainTile + 32) * sizeof(double));   // [9][step * 32 + segment]
  float (*s_S)[kChainTile] = s_L + 9;                                                          // [6][step * 32 + segment]
  double* s_red = reinterpret_cast<double*>(s_raw + 3 * (kChainTile + 32) * sizeof(double) + 15 * kChainTile * sizeof(float));   // [2][32]
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int64_t tile0 = (int64_t)blockIdx.x * kChainTile;
  double alpha = 0.0;
  if (!kInit) { const double pq = scal[S_PQ]; alpha = (pq != 0.0) ? scal[S_RZ] / pq : 0.0; }
  // the tile's factors: stored step-major ([step][segment] = this tile's index space), copied as they lie, 16 bytes a piece
  for (int i = t; i < 15 * (kChainTile / 4); i += kChainThreads) {
    const int c = i / (kChainTile / 4), o = (i % (kChainTile / 4)) * 4;
    const float* src = (c < 9 ? chL + (int64_t)c * ldn : chS + (int64_t)(c - 9) * ldn) + tile0 + o;
    cp_async16(&s_L[c][o], src);                                // s_S = s_L + 9: one [15][1024] array
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // phase 1 (coalesced, storage order): residual update, staged into shared memory in the chain's natural order
  double rv[kChainRowsPerThread][3];
  double rr = 0.0, rz = 0.0;
  {
    double qv[kChainRowsPerThread][3], wv[kChainRowsPerThread][3];
    double4 pv[kChainRowsPerThread];
#pragma unroll
    for (int u = 0; u < kChainRowsPerThread; ++u) {
      const int64_t row = tile0 + u * kChainThreads + t;
      const bool in = row < nrows;
      if (kInit) {
        const bool f = in && is_free[row] != 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) rv[u][c] = f ? rhs[c * ldn + row] : 0.0;
      } else {
        pv[u] = in ? p4r[row_lo + row] : make_double4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          rv[u][c] = in ? r[c * ldn + row] : 0.0;
          qv[u][c] = in ? q[c * ldn + row] : 0.0;
          wv[u][c] = in ? w[c * ldn + row] : 0.0;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kChainRowsPerThread; ++u) {
      const int64_t row = tile0 + u * kChainThreads + t;
      if (row < nrows) {
        if (kInit) {
          w[0 * ldn + row] = 0; w[1 * ldn + row] = 0; w[2 * ldn + row] = 0;
        } else {
          rv[u][0] = fma(-alpha, qv[u][0], rv[u][0]); rv[u][1] = fma(-alpha, qv[u][1], rv[u][1]); rv[u][2] = fma(-alpha, qv[u][2], rv[u][2]);
          w[0 * ldn + row] = fma(alpha, pv[u].x, wv[u][0]);
          w[1 * ldn + row] = fma(alpha, pv[u].y, wv[u][1]);
          w[2 * ldn + row] = fma(alpha, pv[u].z, wv[u][2]);
        }
        r[0 * ldn + row] = rv[u][0]; r[1 * ldn + row] = rv[u][1]; r[2 * ldn + row] = rv[u][2];
        rr = fma(rv[u][0], rv[u][0], fma(rv[u][1], rv[u][1], fma(rv[u][2], rv[u][2], rr)));
      }
      const int n = perm[row];                                   // natural row inside the tile: the chain's order
      const int sn = n + (n >> 5);
      s_v[0][sn] = rv[u][0]; s_v[1][sn] = rv[u][1]; s_v[2][sn] = rv[u][2];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // phase 2 (warp 0, lane = segment): forward substitution y_j = r_j - L_j y_{j-1}, then z_j = S_j^-1 y_j - L_{j+1}^T z_{j+1}.
  // The dependent chain of a step is four fp64 operations; everything else of step j + 1 (shared-memory loads, float ->
  // double conversions) is issued BEFORE the stores of step j, by hand: the compiler keeps loads behind stores into the
  // same dynamic shared array, which put ~100 cycles of load + convert latency on the critical path of every step.
  if (wid == 0) {
    double y0 = 0, y1 = 0, y2 = 0;
    double ln[9], rn[3];
    {
      const int k = lane * 33, f = lane;
#pragma unroll
      for (int c = 0; c < 9; ++c) ln[c] = (double)s_L[c][f];
#pragma unroll
      for (int c = 0; c < 3; ++c) rn[c] = s_v[c][k];
    }
#pragma unroll 4
    for (int j = 0; j < kChainSeg; ++j) {
      const int k = lane * 33 + j;
      double lc[9], rc[3];
#pragma unroll
      for (int c = 0; c < 9; ++c) lc[c] = ln[c];
#pragma unroll
      for (int c = 0; c < 3; ++c) rc[c] = rn[c];
      if (j + 1 < kChainSeg) {
        const int f1 = (j + 1) * 32 + lane;
#pragma unroll
        for (int c = 0; c < 9; ++c) ln[c] = (double)s_L[c][f1];
#pragma unroll
        for (int c = 0; c < 3; ++c) rn[c] = s_v[c][k + 1];
      }
      // dependent depth 3: (r - l0 y0 - l2 y2) - l1 y1, the product l1 y1 beside the first fma
      const double n0 = fma(-lc[2], y2, fma(-lc[0], y0, rc[0])) - lc[1] * y1;
      const double n1 = fma(-lc[5], y2, fma(-lc[3], y0, rc[1])) - lc[4] * y1;
      const double n2 = fma(-lc[8], y2, fma(-lc[6], y0, rc[2])) - lc[7] * y1;
      y0 = n0; y1 = n1; y2 = n2;
      s_v[0][k] = y0; s_v[1][k] = y1; s_v[2][k] = y2;
    }
    double z0 = 0, z1 = 0, z2 = 0;
    double an[6], vn[3];       // S_j^-1 and y_j of the step to come; its L_{j+1} is the forward factor of the step just done
    {
      const int j = kChainSeg - 1, k = lane * 33 + j, f = j * 32 + lane;
#pragma unroll
      for (int c = 0; c < 6; ++c) an[c] = (double)s_S[c][f];
#pragma unroll
      for (int c = 0; c < 3; ++c) vn[c] = s_v[c][k];
#pragma unroll
      for (int c = 0; c < 9; ++c) ln[c] = 0.0;       // no step after the last
    }
#pragma unroll 4
    for (int j = kChainSeg - 1; j >= 0; --j) {
      const int k = lane * 33 + j, f = j * 32 + lane;
      double ac[6], vc[3], lc[9];
#pragma unroll
      for (int c = 0; c < 6; ++c) ac[c] = an[c];
#pragma unroll
      for (int c = 0; c < 3; ++c) vc[c] = vn[c];
#pragma unroll
      for (int c = 0; c < 9; ++c) lc[c] = ln[c];
      if (j > 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) an[c] = (double)s_S[c][f - 32];
#pragma unroll
        for (int c = 0; c < 3; ++c) vn[c] = s_v[c][k - 1];
#pragma unroll
        for (int c = 0; c < 9; ++c) ln[c] = (double)s_L[c][f];     // L_j: what step j - 1 needs
      }
      // S^-1 y does not depend on z: off the chain; then depth 3 as above
      const double b0 = ac[0] * vc[0] + ac[1] * vc[1] + ac[2] * vc[2];
      const double b1 = ac[1] * vc[0] + ac[3] * vc[1] + ac[4] * vc[2];
      const double b2 = ac[2] * vc[0] + ac[4] * vc[1] + ac[5] * vc[2];
      const double n0 = fma(-lc[6], z2, fma(-lc[0], z0, b0)) - lc[3] * z1;
      const double n1 = fma(-lc[7], z2, fma(-lc[1], z0, b1)) - lc[4] * z1;
      const double n2 = fma(-lc[8], z2, fma(-lc[2], z0, b2)) - lc[5] * z1;
      z0 = n0; z1 = n1; z2 = n2;
      s_v[0][k] = z0; s_v[1][k] = z1; s_v[2][k] = z2;
    }
  }
  __syncthreads();
  // phase 3 (coalesced): z out, r.z
#pragma unroll
  for (int u = 0; u < kChainRowsPerThread; ++u) {
    const int64_t row = tile0 + u * kChainThreads + t;
    if (row < nrows) {
      const int n = perm[row];
      const int sn = n + (n >> 5);
      const double z0 = s_v[0][sn], z1 = s_v[1][sn], z2 = s_v[2][sn];
      z[0 * ldn + row] = z0; z[1 * ldn + row] = z1; z[2 * ldn + row] = z2;
      rz = fma(rv[u][0], z0, fma(rv[u][1], z1, fma(rv[u][2], z2, rz)));
      if (kInit) p4w[row_lo + row] = make_double4(z0, z1, z2, 0.0);
    }
  }
  // fixed-shape block reduction: warp trees, then thread 0 adds the warp sums in index order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { rz += __shfl_xor_sync(0xffffffffu, rz, o); rr += __shfl_xor_sync(0xffffffffu, rr, o); }
  if (lane == 0) { s_red[wid] = rz; s_red[32 + wid] = rr; }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < kChainThreads / 32; ++k) { a += s_red[k]; b += s_red[32 + k]; }
    task_part[blockIdx.x] = a; task_part[(size_t)gridDim.x + blockIdx.x] = b;
  }
}

// ------------------------------------------------------------------------------------------------
// LM step helpers
// ------------------------------------------------------------------------------------------------
// p4[row] = (w, 0) for the model-cost SpMV;  S_WG = w.g
__global__ void __launch_bounds__(kVecThreads)
k_pack_step(const double* __restrict__ w, const double* __restrict__ g, int32_t row_lo, int32_t nrows, int64_t ldn,
            double4* p4, double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  double wg = 0;
  if (i < nrows) {
    const double w0 = w[0 * ldn + i], w1 = w[1 * ldn + i], w2 = w[2 * ldn + i];
    p4[row_lo + i] = make_double4(w0, w1, w2, 0.0);
    wg = fma(w0, g[0 * ldn + i], fma(w1, g[1 * ldn + i], w2 * g[2 * ldn + i]));
  }
  double s[1] = {wg};
  grid_reduce_sum<1, kVecThreads>(s, partials, ticket, scal + S_WG);
}
// True residual of the LM linear system after the solve, from the SAME products the model-cost step forms:
// q = H w (k_spmv with the undamped diagonal blocks), so r_true = g - q - Lambda w, Lambda_c = lmdiag_c / (radius scale_c^2).
// S_TRES = |r_true|^2 over the parameter rows; the caller divides by |g|^2 (S_RR0 of the solve).
__global__ void __launch_bounds__(kVecThreads)
k_true_residual(const double* __restrict__ g, const double* __restrict__ q, const double* __restrict__ w,
                const double* __restrict__ lmdiag, const double* __restrict__ scale, const uint8_t* __restrict__ is_free,
                int32_t nrows, int64_t ldn, double inv_radius, double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  double rr = 0;
  if (i < nrows && is_free[i]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double sc = scale[c * ldn + i];
      const double lam = lmdiag[c * ldn + i] * inv_radius / (sc * sc);
      const double r = g[c * ldn + i] - q[c * ldn + i] - lam * w[c * ldn + i];
      rr = fma(r, r, rr);
    }
  }
  double s[1] = {rr};
  grid_reduce_sum<1, kVecThreads>(s, partials, ticket, scal + S_TRES);
}

// candidate = x - w (delta = -w);  S_STEP_SQ = |w|^2;  S_XSQ = |candidate|^2 over parameter rows
__global__ void __launch_bounds__(kVecThreads)
k_apply_step(const double4* __restrict__ xyt, const double* __restrict__ w, const uint8_t* __restrict__ is_free,
             int32_t row_lo, int32_t nrows, int64_t ldn, double4* cand_xyt,
             double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  double ss = 0, xs = 0;
  if (i < nrows) {
    double4 p = xyt[row_lo + i];
    if (is_free[i]) {
      const double w0 = w[0 * ldn + i], w1 = w[1 * ldn + i], w2 = w[2 * ldn + i];
      p.x -= w0; p.y -= w1; p.z -= w2;
      ss = fma(w0, w0, fma(w1, w1, w2 * w2));
      xs = fma(p.x, p.x, fma(p.y, p.y, p.z * p.z));
    }
    cand_xyt[row_lo + i] = p;
  }
  double sv[2] = {ss, xs};
  grid_reduce_sum<2, kVecThreads>(sv, partials, ticket, scal + S_STEP_SQ);   // S_STEP_SQ, S_XSQ
}
// |x|^2 over parameter rows
__global__ void __launch_bounds__(kVecThreads)
k_xnorm(const double4* __restrict__ xyt, const uint8_t* __restrict__ is_free, int32_t row_lo, int32_t nrows,
        double* partials, unsigned int* ticket, double* scal) {
  const int32_t i = blockIdx.x * kVecThreads + threadIdx.x;
  double xs = 0;
  if (i < nrows && is_free[i]) { const double4 p = xyt[row_lo + i]; xs = fma(p.x, p.x, fma(p.y, p.y, p.z * p.z)); }
  double s[1] = {xs};
  grid_reduce_sum<1, kVecThreads>(s, partials, ticket, scal + S_XSQ);
}

// SoA [3][ldn] <-> AoS N x 3 (host-facing) for owned rows
// natural-order N x 3 (C-ABI side) <-> storage-order SoA (device side)
__global__ void k_soa_to_aos(const double* __restrict__ v, const uint16_t* __restrict__ rank_of, int32_t nrows, int64_t ldn, double* out3) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const int32_t m = row_pos(rank_of, i);
  out3[3 * (int64_t)i] = v[m]; out3[3 * (int64_t)i + 1] = v[ldn + m]; out3[3 * (int64_t)i + 2] = v[2 * ldn + m];
}
__global__ void k_aos_to_soa(const double* __restrict__ in3, const uint16_t* __restrict__ rank_of, int32_t nrows, int64_t ldn, double* v) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const int32_t m = row_pos(rank_of, i);
  v[m] = in3[3 * (int64_t)i]; v[ldn + m] = in3[3 * (int64_t)i + 1]; v[2 * ldn + m] = in3[3 * (int64_t)i + 2];
}

}  // namespace dcs
