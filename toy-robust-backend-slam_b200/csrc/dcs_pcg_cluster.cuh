// dcs_pcg_cluster.cuh — small graphs: the WHOLE preconditioned-CG solve in one launch of one thread-block cluster.
//
// The reference's own datasets are small (INTEL 1228 poses, M3500 3500, CSAIL 1045, FR079 989, FRH 1316, MIT 808;
// DCS-ceres/dataset/*.g2o) and so are the problems of the METHOD 3/4 clients that dcs_solve_batch serves.  On such a graph
// a PCG iteration of the general path (k_spmv, k_fold_tasks, k_pcg_chain, k_fold_tasks, k_pcg_direction inside a CUDA
// graph) is five dependent kernel boundaries: 21-23 us per iteration whatever the size, three orders of magnitude above
// its memory time.  Here ONE cluster of up to eight CTAs (one CTA per 1024-pose chain tile, the sorting window of the
// row layout) runs every iteration of the solve:
//   * the tile's preconditioner factors are copied into shared memory ONCE per solve (k_pcg_chain: once per iteration);
//   * r and the own rows of p live in registers for the whole solve (two stored rows per thread), w in shared memory;
//     only p goes to global memory, because it is the operand the other rows gather;
//   * the three dependencies of an iteration that cross CTAs - p complete before the product, p.q, (r.z, r.r) - are
//     hardware cluster barriers; the scalar sums travel through distributed shared memory: every CTA stores its partial
//     into its slot of EVERY CTA's slot array and all CTAs add the slots in rank order, so alpha, beta and the
//     convergence decision are bit-identical everywhere (no CTA can leave the loop alone) and the solve stays
//     bit-reproducible;
//   * the convergence test runs every `batch` iterations, like the host's test between graph replays of the general
//     path, so both paths stop at multiples of options.pcg_check_every.
// Same arithmetic per row as k_spmv / k_pcg_chain / k_pcg_direction (same products, same substitution); only the
// association of the dot-product sums differs.
#pragma once
#include <cooperative_groups.h>

#include "dcs_kernels.cuh"

namespace dcs {
namespace cg = cooperative_groups;

constexpr int kClThreads = 512;                        // threads per CTA
constexpr int kClRows = kChainTile / kClThreads;       // stored rows per thread
constexpr int kClWarps = kClThreads / 32;
constexpr int kClMaxTiles = 8;                         // portable cluster size: graphs up to 8192 poses
// dynamic shared memory: s_v [3][1056] doubles (the substitution's vector, chain order) | s_w [3][1024] doubles (storage
// order) | the tile's factors [15][1024] floats | warp partials [3][16] | slots A [1][8], B [2][8] | the CTA's column
// words (when they fit).  Kept small on purpose: what shared memory takes, L1 loses, and the product
// phase needs L1 lines for its loads in flight (profiles/r02_small_graphs.md: with 225 KB of shared memory the same
// product loop takes 16.8 K instead of 10.2 K cycles).
constexpr size_t kClSmemVec = 3 * (kChainTile + 32) * sizeof(double);
constexpr size_t kClSmemRow = 3 * kChainTile * sizeof(double);
constexpr size_t kClSmemFac = 15 * kChainTile * sizeof(float);
constexpr size_t kClSmemRed = (3 * kClWarps + 3 * kClMaxTiles) * sizeof(double);
constexpr size_t kClSmemBase = kClSmemVec + kClSmemRow + kClSmemFac + kClSmemRed;
constexpr size_t kClSmemMax = 163 * 1024;              // stay inside the 164-KB shared-memory configuration (92 KB of L1 remain)
constexpr int kClMaxColWords = (int)((kClSmemMax - kClSmemBase) / sizeof(uint32_t));
static_assert(kClRows * kClThreads == kChainTile && kChainTile == kWindow, "one CTA per 1024-row window");
static_assert(kClSmemVec % 16 == 0 && kClSmemRow % 16 == 0 && kClSmemFac % 16 == 0 && kClSmemBase % 16 == 0, "shared-memory carve-up alignment");

// p is written by other CTAs of the cluster during the kernel: read it through L2 (never the non-coherent path, never a
// line L1 kept from the previous iteration)
__device__ __forceinline__ void ld_cg3(const double4* p, double& x, double& y, double& z) {
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
  asm volatile("ld.global.cg.f64 %0, [%1+16];" : "=d"(z) : "l"(p) : "memory");
}

// Sum of K per-thread values over the whole cluster; the result is the same bit pattern in every thread of every CTA.
// s_warp: [K][kClWarps] of this CTA; s_slot: [K][kClMaxTiles], written by every CTA of the cluster (slot = its rank).
// A slot array must not be reused by the NEXT call (a fast CTA would overwrite slots a slow one is still adding): the
// kernel alternates two arrays, and between two uses of the same array lie two cluster barriers.
template <int K>
__device__ __forceinline__ void cluster_sum(cg::cluster_group& cl, double (&v)[K], double* s_warp, double* s_slot, int t,
                                            unsigned me, unsigned nb) {
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((t & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_warp[k * kClWarps + (t >> 5)] = v[k];
  }
  __syncthreads();
  if (t < K) {            // thread k: the CTA's partial of value k (warp sums in index order), posted to every CTA
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < kClWarps; ++i) a += s_warp[t * kClWarps + i];
    for (unsigned r = 0; r < nb; ++r) cl.map_shared_rank(s_slot, r)[t * kClMaxTiles + me] = a;
  }
  cl.sync();              // barrier.cluster arrive.release / wait.acquire: the remote shared-memory stores are visible
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double a = 0.0;
    for (unsigned r = 0; r < nb; ++r) a += s_slot[k * kClMaxTiles + r];
    v[k] = a;
  }
}

// Forward / backward substitution of one tile by one warp (lane = segment), in place in s_v (index n + n / 32):
//   y_j = r_j - L_j y_{j-1},   z_j = S_j^-1 y_j - L_{j+1}^T z_{j+1}.
// The same steps, in the same order, as phase 2 of k_pcg_chain (that kernel keeps its own inlined copy: its schedule
// is tuned at the SASS level and a shared helper changes it).
__device__ __forceinline__ void chain_substitute(double (*s_v)[kChainTile + 32], float (*s_L)[kChainTile], float (*s_S)[kChainTile], int lane) {
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

#ifdef DCS_DEV_PROBES
// development build only: cycles thread 0 of CTA 0 spent in each phase of the iteration loop, summed over the solve
// (dcs_debug_cluster_cycles): 0 barrier before the product, 1 product, 2 p.q exchange, 3 vector update + staging,
// 4 substitution, 5 z + (r.z, r.r) exchange, 6 direction, 7 iterations
__device__ unsigned long long g_cl_cycles[8];
#define DCS_CL_MARK(i) do { if (t == 0 && me == 0) { const long long c_ = clock64(); cyc[i] += c_ - tlast; tlast = c_; } } while (0)
#else
#define DCS_CL_MARK(i) do { } while (0)
#endif

// Solve (offdiag blocks + D) w = rhs (rhs masked to the parameter rows) by PCG with the chain preconditioner whose
// factors k_chain_factor left in chL / chS.  grid = cluster = ntiles CTAs (<= kClMaxTiles).  Results: w (storage order
// SoA), scal[S_RR0] = |rhs|^2, scal[S_RR] = |r|^2 at exit, scal[S_PCG_ITERS] = iterations run (a multiple of `batch`).
// max_iter must be a multiple of batch.  cols_smem_words: column words a CTA may stage (0: read `cols`).
__global__ void __launch_bounds__(kClThreads, 1)
k_pcg_cluster(const double* __restrict__ rhs, const uint8_t* __restrict__ is_free, RowLayout L, const uint32_t* __restrict__ cols,
              const double* __restrict__ Hoff, const double* __restrict__ D, const float* __restrict__ chL,
              const float* __restrict__ chS, const uint16_t* __restrict__ perm, int32_t n_loc, int32_t cols_smem_words,
              int32_t batch, int32_t max_iter, double rel_tol, double4* p4, double* __restrict__ w, double* __restrict__ scal) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char s_raw[];
  double (*s_v)[kChainTile + 32] = reinterpret_cast<double (*)[kChainTile + 32]>(s_raw);
  double (*s_w)[kChainTile] = reinterpret_cast<double (*)[kChainTile]>(s_raw + kClSmemVec);                  // w, storage order
  float (*s_L)[kChainTile] = reinterpret_cast<float (*)[kChainTile]>(s_raw + kClSmemVec + kClSmemRow);       // [9][step * 32 + segment]
  float (*s_S)[kChainTile] = s_L + 9;                                                                        // [6][step * 32 + segment]
  double* s_warp = reinterpret_cast<double*>(s_raw + kClSmemVec + kClSmemRow + kClSmemFac);                  // [3][kClWarps]
  double* s_slotA = s_warp + 3 * kClWarps;                                                                   // [1][kClMaxTiles]: p.q
  double* s_slotB = s_slotA + kClMaxTiles;                                                                   // [2][kClMaxTiles]: r.z, r.r
  uint32_t* s_cols = reinterpret_cast<uint32_t*>(s_raw + kClSmemBase);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const unsigned me = cl.block_rank(), nb = cl.num_blocks();
  const int64_t ldn = L.ldn;
  const int64_t row0 = (int64_t)blockIdx.x * kChainTile;      // first stored row of this CTA's tile

  // once per solve: the tile's factors (stored step-major = this tile's index space), copied as they lie
  for (int i = t; i < 15 * (kChainTile / 4); i += kClThreads) {
    const int c = i / (kChainTile / 4), o = (i % (kChainTile / 4)) * 4;
    const float* src = (c < 9 ? chL + (int64_t)c * ldn : chS + (int64_t)(c - 9) * ldn) + row0 + o;
    cp_async16(&s_L[c][o], src);                              // s_S = s_L + 9: one [15][1024] array
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // once per solve: the column words of the CTA's 32 tasks (one contiguous run of tiles), when they fit
  const int task_lo = blockIdx.x * (kChainTile / kSlice);
  const int64_t first_tile = L.task_info[task_lo].x;
  const int64_t n_words = ((int64_t)L.task_info[task_lo + kChainTile / kSlice].x - first_tile) * kSlice;
  const bool cols_staged = n_words <= (int64_t)cols_smem_words;       // uniform over the CTA
  if (cols_staged)
    for (int64_t i = t; i < n_words; i += kClThreads) s_cols[i] = cols[first_tile * kSlice + i];

  int32_t row[kClRows];                                        // stored rows of this thread (< 8192: 32-bit state, fewer registers)
  bool in[kClRows];
  int deg[kClRows], sn[kClRows];
  int32_t rtile[kClRows];                                      // first SELL tile of the row's task
#pragma unroll
  for (int u = 0; u < kClRows; ++u) {
    row[u] = (int32_t)row0 + u * kClThreads + t;
    in[u] = row[u] < L.nrows;
    deg[u] = in[u] ? (int)L.rowinfo[row[u]].x : 0;
    rtile[u] = L.task_info[row[u] >> 5].x;
    const int n = perm[row[u]];                                // natural row inside the tile: the chain's order
    sn[u] = n + (n >> 5);
  }

  // initialisation: w = 0, r = rhs (masked), z = M^-1 r, p = z.  An s_w entry is touched by its own thread only.
  double r[kClRows][3], p[kClRows][3];
  double rr_part = 0.0, rz_part = 0.0;
#pragma unroll
  for (int u = 0; u < kClRows; ++u) {
    const bool f = in[u] && is_free[row[u]] != 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r[u][c] = f ? rhs[c * ldn + row[u]] : 0.0;
      s_w[c][u * kClThreads + t] = 0.0;
      s_v[c][sn[u]] = r[u][c];
    }
    rr_part = fma(r[u][0], r[u][0], fma(r[u][1], r[u][1], fma(r[u][2], r[u][2], rr_part)));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (wid == 0) chain_substitute(s_v, s_L, s_S, lane);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kClRows; ++u) {
#pragma unroll
    for (int c = 0; c < 3; ++c) p[u][c] = s_v[c][sn[u]];
    rz_part = fma(r[u][0], p[u][0], fma(r[u][1], p[u][1], fma(r[u][2], p[u][2], rz_part)));
    if (in[u]) p4[row[u]] = make_double4(p[u][0], p[u][1], p[u][2], 0.0);
  }
  double v2[2] = {rz_part, rr_part};
  cluster_sum<2>(cl, v2, s_warp, s_slotB, t, me, nb);
  double rz = v2[0];
  const double rr0 = v2[1];
  double rr = rr0;
  int it = 0;
  if (rr0 > 0.0 && isfinite(rr0)) {                            // the same value in every thread of the cluster
    const double target = rel_tol * rel_tol * rr0;
#ifdef DCS_DEV_PROBES
    long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    while (it < max_iter) {
      // every row of p (written above / at the end of the previous iteration) before anyone gathers it: the barrier's
      // arrive.release / wait.acquire orders the global stores of all threads of the cluster before the gathers below
      cl.sync();
      DCS_CL_MARK(0);
      // q = A p over the thread's rows (k_spmv's row walk: diagonal block first, rounds in order), p.q.  Every access of
      // the phase goes to L2: the barrier invalidated L1, and p comes from the other CTAs.
      double y[kClRows][3];
      double dot = 0.0;
#pragma unroll
      for (int u = 0; u < kClRows; ++u) {
        y[u][0] = y[u][1] = y[u][2] = 0.0;
        if (in[u]) {
          const double a00 = D[0 * ldn + row[u]], a01 = D[1 * ldn + row[u]], a02 = D[2 * ldn + row[u]];
          const double a11 = D[3 * ldn + row[u]], a12 = D[4 * ldn + row[u]], a22 = D[5 * ldn + row[u]];
          double y0 = fma(a00, p[u][0], fma(a01, p[u][1], a02 * p[u][2]));
          double y1 = fma(a01, p[u][0], fma(a11, p[u][1], a12 * p[u][2]));
          double y2 = fma(a02, p[u][0], fma(a12, p[u][1], a22 * p[u][2]));
          const uint32_t* cp = cols_staged ? s_cols + (rtile[u] - (int32_t)first_tile) * kSlice + lane : cols + (int64_t)rtile[u] * kSlice + lane;   // round k: + 32 k
          const double* hp = Hoff + (int64_t)rtile[u] * (kBlockVals * 32) + lane;                        // round k: + 256 k, value c: + 32 c
          constexpr int U = 2;                                 // rounds in flight per row
          int k = 0;
          for (; k + U <= deg[u]; k += U) {
            uint32_t j[U]; double h[U][kBlockVals]; double px[U], py[U], pz[U];
#pragma unroll
            for (int q = 0; q < U; ++q) j[q] = cp[(k + q) * kSlice] & kIdxMask;
#pragma unroll
            for (int q = 0; q < U; ++q)
#pragma unroll
              for (int c = 0; c < kBlockVals; ++c) h[q][c] = __ldg(hp + (int64_t)(k + q) * (kBlockVals * 32) + c * 32);
#pragma unroll
            for (int q = 0; q < U; ++q) { DCS_ASSERT((int32_t)j[q] < n_loc); ld_cg3(p4 + j[q], px[q], py[q], pz[q]); }
#pragma unroll
            for (int q = 0; q < U; ++q) {
              y0 = fma(h[q][0], px[q], fma(h[q][1], py[q], fma(h[q][2], pz[q], y0)));
              y1 = fma(h[q][1], px[q], fma(h[q][3], py[q], fma(h[q][4], pz[q], y1)));
              y2 = fma(h[q][5], px[q], fma(h[q][6], py[q], fma(h[q][7], pz[q], y2)));
            }
          }
          for (; k < deg[u]; ++k) {
            const uint32_t j = cp[k * kSlice] & kIdxMask;
            double h[kBlockVals], px, py, pz;
#pragma unroll
            for (int c = 0; c < kBlockVals; ++c) h[c] = __ldg(hp + (int64_t)k * (kBlockVals * 32) + c * 32);
            DCS_ASSERT((int32_t)j < n_loc);
            ld_cg3(p4 + j, px, py, pz);
            y0 = fma(h[0], px, fma(h[1], py, fma(h[2], pz, y0)));
            y1 = fma(h[1], px, fma(h[3], py, fma(h[4], pz, y1)));
            y2 = fma(h[5], px, fma(h[6], py, fma(h[7], pz, y2)));
          }
          y[u][0] = y0; y[u][1] = y1; y[u][2] = y2;
          dot = fma(p[u][0], y0, fma(p[u][1], y1, fma(p[u][2], y2, dot)));
        }
      }
      DCS_CL_MARK(1);
      double v1[1] = {dot};
      cluster_sum<1>(cl, v1, s_warp, s_slotA, t, me, nb);
      DCS_CL_MARK(2);
      const double pq = v1[0];
      const double alpha = (pq != 0.0) ? rz / pq : 0.0;
      // w += alpha p, r -= alpha q, r staged in the chain's order
      rr_part = 0.0;
#pragma unroll
      for (int u = 0; u < kClRows; ++u) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          r[u][c] = fma(-alpha, y[u][c], r[u][c]);
          s_w[c][u * kClThreads + t] = fma(alpha, p[u][c], s_w[c][u * kClThreads + t]);
          s_v[c][sn[u]] = r[u][c];
        }
        rr_part = fma(r[u][0], r[u][0], fma(r[u][1], r[u][1], fma(r[u][2], r[u][2], rr_part)));
      }
      __syncthreads();
      DCS_CL_MARK(3);
      if (wid == 0) chain_substitute(s_v, s_L, s_S, lane);      // z = M^-1 r
      __syncthreads();
      DCS_CL_MARK(4);
      double z[kClRows][3];
      rz_part = 0.0;
#pragma unroll
      for (int u = 0; u < kClRows; ++u) {
#pragma unroll
        for (int c = 0; c < 3; ++c) z[u][c] = s_v[c][sn[u]];
        rz_part = fma(r[u][0], z[u][0], fma(r[u][1], z[u][1], fma(r[u][2], z[u][2], rz_part)));
      }
      v2[0] = rz_part; v2[1] = rr_part;
      cluster_sum<2>(cl, v2, s_warp, s_slotB, t, me, nb);
      DCS_CL_MARK(5);
      const double rz_next = v2[0];
      rr = v2[1];
      const double beta = (rz != 0.0) ? rz_next / rz : 0.0;
      rz = rz_next;
      // p = z + beta p
#pragma unroll
      for (int u = 0; u < kClRows; ++u) {
#pragma unroll
        for (int c = 0; c < 3; ++c) p[u][c] = fma(beta, p[u][c], z[u][c]);
        if (in[u]) p4[row[u]] = make_double4(p[u][0], p[u][1], p[u][2], 0.0);
      }
      ++it;
      DCS_CL_MARK(6);
      if (it % batch == 0 && !(rr > target)) break;            // converged, or NaN (the caller validates the step)
    }
#ifdef DCS_DEV_PROBES
    if (t == 0 && me == 0) {
#pragma unroll
      for (int i = 0; i < 7; ++i) g_cl_cycles[i] = (unsigned long long)cyc[i];
      g_cl_cycles[7] = (unsigned long long)it;
    }
#endif
  }
#pragma unroll
  for (int u = 0; u < kClRows; ++u)
    if (in[u]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) w[c * ldn + row[u]] = s_w[c][u * kClThreads + t];
    }
  if (me == 0 && t == 0) { scal[S_RR0] = rr0; scal[S_RR] = rr; scal[S_RZ] = rz; scal[S_PCG_ITERS] = (double)it; }
}

}  // namespace dcs
