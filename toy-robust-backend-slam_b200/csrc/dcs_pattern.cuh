// dcs_pattern.cuh — K0: one-time, integer-only construction of the block structure of J^T J.
//
// Replaces what Ceres does once per solve for SPARSE_NORMAL_CHOLESKY (reference call site
// DCS-ceres/main.cpp:156): enumerate the per-residual block pairs, sort them, build the
// compressed structure.  Here: every edge (a,b) yields two half-edges (row=a,col=b) and
// (row=b,col=a) (only for non-constant rows); a hand-written stable LSD radix sort orders them
// by (row, col); the sorted list is (i) the full-storage block CSR the SpMV walks and (ii) after
// a flag+scan unique, the upper-triangular block pattern {(i,i)} U {(min,max)} that the parity
// hook exports.  The CSR is then re-laid per window of kWindow (1024) rows in jagged-diagonal order (rows ranked by
// degree inside the window, a warp task = 32 consecutive ranks) so that a thread-per-row kernel reads it fully
// coalesced with no padding.
#pragma once
#include "dcs_common.cuh"

namespace dcs {

// ---- exclusive scan (int32), in place; recursion on block totals --------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void k_scan_tile(int32_t* data, int64_t n, int32_t* totals) {
  __shared__ int32_t s_warp[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { v[i] = (base + i < n) ? data[base + i] : 0; sum += v[i]; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int32_t w = (lane < kScanThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    if (lane < kScanThreads / 32) s_warp[lane] = w;
  }
  __syncthreads();
  int32_t excl = incl - sum + (wid > 0 ? s_warp[wid - 1] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { if (base + i < n) data[base + i] = excl; excl += v[i]; }
  if (threadIdx.x == kScanThreads - 1) totals[blockIdx.x] = excl;
}

__global__ void k_scan_add(int32_t* data, int64_t n, const int32_t* totals) {
  const int64_t i = (int64_t)blockIdx.x * kScanTile + threadIdx.x;
  const int32_t off = totals[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { const int64_t j = i + (int64_t)k * kScanThreads; if (j < n) data[j] += off; }
}

// ---- radix sort pass: 8-bit digit, stable ----------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void k_radix_hist(const uint64_t* keys, int64_t n, int shift, int32_t* hist, int nblk) {
  __shared__ int32_t s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int c = 0; c < kSortItems; ++c) {
    const int64_t i = base + (int64_t)c * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = s_h[threadIdx.x];   // digit-major
}

__global__ void k_radix_scatter(const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                                uint32_t* vals_out, int64_t n, int shift, const int32_t* hist, int nblk) {
  __shared__ int32_t s_run[256];
  __shared__ int32_t s_wcnt[kSortThreads / 32][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s_run[threadIdx.x] = hist[(int64_t)threadIdx.x * nblk + blockIdx.x];
#pragma unroll
  for (int w = 0; w < kSortThreads / 32; ++w) s_wcnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int c = 0; c < kSortItems; ++c) {
    const int64_t i = base + (int64_t)c * kSortThreads + threadIdx.x;
    const bool valid = i < n;
    uint64_t key = 0; uint32_t val = 0;
    if (valid) { key = keys_in[i]; val = vals_in[i]; }
    const unsigned int d = valid ? (unsigned int)((key >> shift) & 0xFF) : 256u + lane;  // invalid: unique class
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) s_wcnt[wid][d] = __popc(peers);
    __syncthreads();
    if (valid) {
      int off = s_run[d] + rank;
      for (int w = 0; w < wid; ++w) off += s_wcnt[w][d];
      keys_out[off] = key;
      vals_out[off] = val;
    }
    __syncthreads();
    {
      int t = 0;
#pragma unroll
      for (int w = 0; w < kSortThreads / 32; ++w) { t += s_wcnt[w][threadIdx.x]; s_wcnt[w][threadIdx.x] = 0; }
      s_run[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// ---- half-edge generation ---------------------------------------------------------------------
// key = row << 32 | nonowner << 27 | col (27 column bits); val = edge << 1 | side.  Rows outside [row_lo,row_hi) or constant
// rows are not generated.  "Owner" half-edges (col > row: the upper triangle, the blocks k_linearize stores;
// also blocks whose partner row lives on another rank) sort before a row's other half-edges, so that in the
// jagged-diagonal layout the early rounds are (almost) all owners and their block stores fill whole sectors.
__global__ void k_halfedge_count(const int32_t* ea, const int32_t* eb, int32_t E, int32_t fixed, int32_t row_lo,
                                 int32_t row_hi, int32_t* cnt) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t a = ea[e], b = eb[e];
  int c = 0;
  if (a != fixed && a >= row_lo && a < row_hi) ++c;
  if (b != fixed && b >= row_lo && b < row_hi) ++c;
  cnt[e] = c;
}
__global__ void k_halfedge_fill(const int32_t* ea, const int32_t* eb, int32_t E, int32_t fixed, int32_t row_lo,
                                int32_t row_hi, const int32_t* off, uint64_t* keys, uint32_t* vals) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t a = ea[e], b = eb[e];
  int32_t o = off[e];
  auto key = [&](int32_t row, int32_t col) {
    const bool halo = col < row_lo || col >= row_hi;
    const bool owner = col != fixed && (row < col || halo);
    return ((uint64_t)(uint32_t)row << 32) | (halo ? kKeyHalo : 0u) | (owner ? 0u : kKeyNonOwner) | (uint32_t)col;
  };
  if (a != fixed && a >= row_lo && a < row_hi) { keys[o] = key(a, b); vals[o] = ((uint32_t)e << 1); ++o; }
  if (b != fixed && b >= row_lo && b < row_hi) { keys[o] = key(b, a); vals[o] = ((uint32_t)e << 1) | 1u; }
}

// degree of every pose over ALL edges (decides which poses are parameters at all)
__global__ void k_pose_degree(const int32_t* ea, const int32_t* eb, int32_t E, int32_t* deg) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  atomicAdd(&deg[ea[e]], 1);
  atomicAdd(&deg[eb[e]], 1);
}

// row_ptr[r - row_lo] = first sorted half-edge whose row >= r   (r in [row_lo, row_hi]); binary search
__global__ void k_row_ptr(const uint64_t* keys, int32_t nh, int32_t row_lo, int32_t nrows, int32_t* row_ptr) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  const uint64_t target = (uint64_t)(uint32_t)(row_lo + r) << 32;
  int32_t lo = 0, hi = nh;
  while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
  row_ptr[r] = lo;
}

// ---- sliced-ELL re-layout (SELL-32, sorting window kWindow rows) ----------------------------------
// Rows of a window are ranked by decreasing degree (stable) and stored in rank order.  A warp task = 32 consecutive
// ranks: its lanes have near-equal degrees (97% lane efficiency at kWindow = 1024 on the 1M-pose benchmark graph,
// 86% at 128), its k-th half-edges are one 32-slot tile, its tiles are consecutive (see RowLayout).
// pass 1: rank of each row inside its window
__global__ void __launch_bounds__(kWindow)
k_jds_rank(const int32_t* row_ptr, const uint64_t* keys, int32_t nrows, int by_local, uint16_t* rank_of, uint16_t* perm,
           uint32_t* rank_info, int32_t* rank_nloc) {
  __shared__ int32_t s_deg[kWindow];
  __shared__ int32_t s_own[kWindow];
  const int32_t r = blockIdx.x * kWindow + threadIdx.x;
  int32_t d = 0, own = 0, nloc = 0;
  if (r < nrows) {
    const int32_t b = row_ptr[r], e = row_ptr[r + 1];
    d = e - b;
    // a row's entries are sorted [local owner | local non-owner | halo (owners)]: nloc = entries with the halo bit clear
    int32_t lo = b, hi = e;
    while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (keys[mid] & kKeyHalo) hi = mid; else lo = mid + 1; }
    nloc = lo - b;
    // own = how many carry the non-owner key bit clear (inside the local part the bit is monotone)
    lo = b; hi = b + nloc;
    while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (keys[mid] & kKeyNonOwner) hi = mid; else lo = mid + 1; }
    own = (lo - b) + (d - nloc);
    if (by_local) own = nloc;      // multi-rank: second ranking key = local entries, so that in every tile of a task the
                                   // local lanes are a prefix and the halo lanes a suffix (k_spmv's two passes)
  }
  s_deg[threadIdx.x] = d;
  s_own[threadIdx.x] = own;
  __syncthreads();
  // Rank by (degree, owner entries) descending, stable.  The second key makes the lanes that write an
  // off-diagonal block in round k (own > k) a contiguous run of their slice, so k_linearize's owner-only block
  // stores fill whole 32-byte sectors instead of scattering 8-byte pieces over every sector of the round.
  int rank = 0;
  for (int u = 0; u < kWindow; ++u) {
    const int32_t du = s_deg[u], ou = s_own[u];
    rank += (du > d) || (du == d && (ou > own || (ou == own && u < (int)threadIdx.x)));
  }
  rank_of[(int64_t)blockIdx.x * kWindow + threadIdx.x] = (uint16_t)rank;
  perm[(int64_t)blockIdx.x * kWindow + rank] = (uint16_t)threadIdx.x;
  // what a warp task needs to start, in one coalesced word per rank: local row (10 bits) | degree
  rank_info[(int64_t)blockIdx.x * kWindow + rank] = ((uint32_t)d << 10) | (uint32_t)threadIdx.x;
  rank_nloc[(int64_t)blockIdx.x * kWindow + rank] = nloc;
}
// pass 2: sorted CSR position -> SELL slot = (first tile of the row's task + k) * 32 + lane
__global__ void k_sell_slot(const uint64_t* keys, int32_t nh, int32_t row_lo, const int32_t* row_ptr,
                            const uint16_t* rank_of, const int32_t* tile0, int32_t* slot) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int32_t r = (int32_t)(keys[i] >> 32) - row_lo;
  const int32_t k = i - row_ptr[r];
  const int32_t m = row_pos(rank_of, r);               // stored position: task = m / 32, lane = m % 32
  slot[i] = (tile0[m >> 5] + k) * kSlice + (m & 31);
}

// ---- odometry-chain entries for the segment preconditioner --------------------------------------------
// chain_idx[r] = first (row,col)-sorted half-edge of local row r whose column is the next pose (r+1), or -1;
// chain_cnt[r] = how many (duplicate edges between the same pair add up)
__global__ void k_chain_entries(const uint64_t* keys, int32_t nh, int32_t row_lo, int32_t nrows, int32_t* chain_idx,
                                int32_t* chain_cnt) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const uint64_t target = ((uint64_t)(uint32_t)(row_lo + r) << 32) | (uint32_t)(row_lo + r + 1);   // owner entry: non-owner bit clear
  int32_t lo = 0, hi = nh;
  while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
  int32_t c = 0;
  while (lo + c < nh && keys[lo + c] == target) ++c;
  chain_idx[r] = c ? lo : -1;
  chain_cnt[r] = c;
}

// ---- halo lists (multi-rank): which poses of other ranks do my half-edges reference -----------------------
__global__ void k_halo_mark(const uint64_t* keys, int32_t nh, int32_t row_lo, int32_t row_hi, int32_t* need) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int32_t col = (int32_t)(keys[i] & kIdxMask);
  if (col < row_lo || col >= row_hi) need[col] = 1;      // benign race: everybody writes 1
}
// compaction: list[scan[j]] = j for every marked pose (scan = exclusive scan of need), and the global -> local
// index map of the rank: own rows first ([0, rows_per_rank)), then the halo in global order
__global__ void k_halo_compact(const int32_t* need, const int32_t* scan, int32_t n, int32_t row_lo, int32_t rows_per_rank,
                               const uint16_t* rank_of, int32_t* list, int32_t* g2l) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int32_t l = -1;
  if (j >= row_lo && j < row_lo + rows_per_rank) l = row_pos(rank_of, j - row_lo);     // own rows: storage position
  else if (need[j]) { list[scan[j]] = j; l = rows_per_rank + scan[j]; }
  g2l[j] = l;
}
__global__ void k_halo_pack(const double4* __restrict__ arr, const int32_t* __restrict__ idx, int32_t n, double4* buf) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = arr[idx[i]];
}

// Halo push over NVLink peer memory: entry j of the send list (own row idx[j]) goes straight into the halo
// region of the peer that asked for it (peer pointers opened with CUDA IPC; the lists have the same order on
// both sides, so the destination is dst_base[peer] + position in the peer's slice).  Replaces pack + ncclSend
// /ncclRecv: one kernel, 32-byte coalesced remote stores, no staging buffer.
constexpr int kMaxWorld = 8;   // ranks of one handle group = GPUs of one NVSwitch node (dcs_create rejects more)
struct HaloPeers {
  double4* ptr[kMaxWorld];           // peer array base (nullptr for self / unused)
  int32_t send_off[kMaxWorld + 1];   // slices of the send list per peer
  int32_t dst_base[kMaxWorld];       // first halo entry in the peer's array that belongs to this rank
  int32_t world;
};
template <int kPushPerThread>
__global__ void __launch_bounds__(256)
k_halo_push(const double4* __restrict__ arr, const int32_t* __restrict__ idx, int32_t n, HaloPeers P) {
  const int32_t j0 = blockIdx.x * (256 * kPushPerThread) + threadIdx.x;
  double4 v[kPushPerThread];
#pragma unroll
  for (int u = 0; u < kPushPerThread; ++u) {          // all gathers first (local L2), then the remote stores
    const int32_t j = j0 + u * 256;
    if (j < n) asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w) : "l"(arr + idx[j]));
  }
#pragma unroll
  for (int u = 0; u < kPushPerThread; ++u) {
    const int32_t j = j0 + u * 256;
    if (j >= n) continue;
    int r = 0;
#pragma unroll
    for (int q = 1; q < kMaxWorld; ++q) r += (q < P.world && j >= P.send_off[q]) ? 1 : 0;
    // ONE 256-bit store per entry: a warp writes 32 whole 32-byte sectors (1 KB contiguous in the peer's halo region).
    // A plain double4 assignment compiles to two 128-bit stores, i.e. two half-sector write packets per entry on NVLink.
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(P.ptr[r] + P.dst_base[r] + (j - P.send_off[r])), "d"(v[u].x), "d"(v[u].y),
                 "d"(v[u].z), "d"(v[u].w) : "memory");
  }
}

// Scalar all-reduce / barrier over NVLink peer memory (replaces ncclAllReduce of 1-4 doubles inside the PCG iteration:
// a 10-17 us collective launch for 8-32 bytes).  Every rank owns one XchgBuf (CUDA IPC, opened by all peers).  An
// exchange = one tiny kernel per rank: store my partial values into slot [epoch parity][my rank] of EVERY rank's buffer,
// fence, publish the epoch in that rank's flag word for me (release, system scope); then spin (acquire) until all
// ranks' flags in MY buffer carry the epoch and add their values in rank order - the same order on every rank, so the
// sums are bit-identical everywhere.  All ranks issue the same exchanges in the same stream order, so one running
// epoch serves every call site; a rank can run at most one exchange ahead of a peer (finishing exchange e+1 needs the
// peer's flag e+1, which the peer publishes after it has read epoch e), so two value slots by epoch parity suffice.
// count = 0 is a barrier: the peer stores of earlier kernels on the stream (k_halo_push) happen before the flag.
constexpr int kXchgVals = 4;
struct XchgBuf {
  double data[2][kMaxWorld][kXchgVals];
  unsigned long long flag[kMaxWorld];
  unsigned long long epoch;            // exchanges completed by the owner (owner-private)
};
struct XchgPeers { XchgBuf* buf[kMaxWorld]; int32_t world, rank; };
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(32)
k_xchg_sum(XchgPeers X, double* vals, int count) {
  __shared__ double s_v[kMaxWorld][kXchgVals];
  XchgBuf* mine = X.buf[X.rank];
  const int t = threadIdx.x;
  const unsigned long long e = mine->epoch + 1;      // read by every lane before lane 0 advances it (after the syncwarp)
  if (t < X.world) {
    XchgBuf* dst = X.buf[t];
    for (int k = 0; k < count; ++k) dst->data[e & 1][X.rank][k] = vals[k];
    __threadfence_system();
    st_release_sys(&dst->flag[X.rank], e);
    // a peer that never arrives (its process died, or the ranks issued different call sequences) must not hang the
    // device for good: after ~2 minutes of polling (ranks may enter a collective call seconds apart; NCCL would wait
    // too) the kernel traps, which surfaces as DCS_ERR_CUDA on the host
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->flag[t]) < e) {
      if (clock64() - t0 > 240000000000LL) __trap();
    }
    for (int k = 0; k < count; ++k) s_v[t][k] = __ldcg(&mine->data[e & 1][t][k]);
  }
  __syncwarp();
  if (t == 0) {
    for (int k = 0; k < count; ++k) {
      double a = 0.0;
      for (int r = 0; r < X.world; ++r) a += s_v[r][k];
      vals[k] = a;
    }
    mine->epoch = e;
  }
}

// ---- unique upper pattern (parity hook) ------------------------------------------------------------
// flag[i] = 1 for the first sorted half-edge of every distinct (row,col) with row < col and col not
// constant.  (Diagonal entries are added per non-empty row by the caller.)
__global__ void k_upper_flag(const uint64_t* keys, int32_t nh, int32_t fixed, int32_t* flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const uint64_t k = keys[i];
  const int32_t row = (int32_t)(k >> 32), col = (int32_t)(k & kIdxMask);
  const bool first = (i == 0) || (keys[i - 1] != k);
  flag[i] = (first && row < col && col != fixed) ? 1 : 0;
}

}  // namespace dcs
