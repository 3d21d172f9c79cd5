// dcs_api.cu — C-ABI of libdcs_b200 (include/dcs_b200.h): graph upload + one-time pattern build,
// the evaluate / linearize entry points, the block-Jacobi PCG driver (CUDA-graph batches, device
// scalars) and the host-side Levenberg–Marquardt controller that follows Ceres' default
// trust-region rules (reference call site DCS-ceres/main.cpp:154-163).
//
// No CPU fallback anywhere: every entry point that computes fails with DCS_ERR_CUDA if the CUDA
// runtime does.  Nothing here includes, links or loads anything under oracle/.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dcs_b200.h"
#include "dcs_common.cuh"
#include "dcs_kernels.cuh"
#include "dcs_nccl.h"
#include "dcs_pattern.cuh"
#include "dcs_pcg_cluster.cuh"
#include "dcs_switchable.cuh"

using namespace dcs;

namespace {

thread_local std::string g_err;
thread_local int64_t t_launches = 0;   // this thread's share (sizes a captured graph's launch count)
std::atomic<int64_t> g_launches{0};    // dcs_solve_batch drives several handles from several host threads

#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      g_err = std::string(#call) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
      return DCS_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)
#define CKN(call)                                                                                \
  do {                                                                                           \
    ncclResult_t _r = (call);                                                                    \
    if (_r != ncclSuccess) {                                                                     \
      g_err = std::string(#call) + ": " + nccl_api().GetErrorString(_r);                         \
      return DCS_ERR_NCCL;                                                                       \
    }                                                                                            \
  } while (0)
#define CKS(call)                                                                                \
  do {                                                                                           \
    int _s = (call);                                                                             \
    if (_s != DCS_OK) return _s;                                                                 \
  } while (0)
#define LAUNCH(kernel, grid, block, stream, ...)                                                 \
  do {                                                                                           \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                       \
    ++g_launches; ++t_launches;                                                                  \
  } while (0)

// dcs_solve_batch creates and destroys a handle per item from several host threads.  cudaFree synchronises the whole
// device, so with plain cudaMalloc / cudaFree every release in one thread waits for the kernels of all the others (the
// one-launch PCG of a small graph runs for milliseconds).  While tl_pool_stream is set (dcs_create called from a batch
// worker, single-rank handles only: CUDA IPC needs cudaMalloc memory) device buffers come from the stream-ordered pool
// instead: allocated and released in the order of the handle's own stream, no device-wide synchronisation.
thread_local cudaStream_t tl_pool_stream = nullptr;
thread_local bool tl_batch_worker = false;

// Page-locked staging buffers of a batch worker's handles: cudaMallocHost / cudaFreeHost cost a driver round trip each and
// cudaFreeHost synchronises the device like cudaFree, so a worker keeps the blocks of the handle it just destroyed for
// the handle it creates next (same thread) and returns them when it leaves dcs_solve_batch.
struct PinnedCache {
  std::vector<std::pair<void*, size_t>> blocks;
  void clear() { for (auto& b : blocks) cudaFreeHost(b.first); blocks.clear(); }
  ~PinnedCache() { clear(); }
};
thread_local PinnedCache tl_pinned;
cudaError_t pinned_get(void** p, size_t bytes, size_t* cap) {
  if (tl_batch_worker) {
    size_t best = tl_pinned.blocks.size();      // best fit: the three blocks of a handle differ by orders of magnitude
    for (size_t i = 0; i < tl_pinned.blocks.size(); ++i)
      if (tl_pinned.blocks[i].second >= bytes && (best == tl_pinned.blocks.size() || tl_pinned.blocks[i].second < tl_pinned.blocks[best].second)) best = i;
    if (best < tl_pinned.blocks.size()) {
      *p = tl_pinned.blocks[best].first; *cap = tl_pinned.blocks[best].second;
      tl_pinned.blocks.erase(tl_pinned.blocks.begin() + (long)best);
      return cudaSuccess;
    }
  }
  *cap = bytes;
  return cudaMallocHost(p, bytes);
}
void pinned_put(void* p, size_t cap) {
  if (!p) return;
  if (tl_batch_worker && tl_pinned.blocks.size() < 16) tl_pinned.blocks.emplace_back(p, cap);
  else cudaFreeHost(p);
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t pool_stream = nullptr;      // set when p came from cudaMallocAsync: released on the same stream
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) { if (pool_stream) cudaFreeAsync(p, pool_stream); else cudaFree(p); }
    p = nullptr; n = 0; pool_stream = nullptr;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    if (tl_pool_stream) { pool_stream = tl_pool_stream; return cudaMallocAsync(&p, count * sizeof(T), pool_stream); }
    return cudaMalloc(&p, count * sizeof(T));
  }
  // zero-fill ordered on the handle's (non-blocking) stream: a plain cudaMemset goes to the legacy default stream,
  // which that stream does not order against
  cudaError_t alloc_zero(size_t count, cudaStream_t st) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess || count == 0) return e;
    return cudaMemsetAsync(p, 0, count * sizeof(T), st);
  }
};

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// per-edge constants: (tmx,tmy) = Rm^T (dx,dy), cos/sin of the measured rotation, DCS flag
__global__ void k_edge_prep(const double* __restrict__ meas, const uint8_t* __restrict__ kind, int32_t E, int dcs_on,
                            double* tmx, double* tmy, double* thm, uint8_t* dcs_flag) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double dx = meas[3 * (int64_t)e], dy = meas[3 * (int64_t)e + 1], th = meas[3 * (int64_t)e + 2];
  double s, c;
  sincos(th, &s, &c);
  tmx[e] = fma(c, dx, s * dy);
  tmy[e] = fma(c, dy, -s * dx);
  thm[e] = th;
  dcs_flag[e] = (dcs_on && kind[e] != DCS_EDGE_ODOMETRY) ? 1 : 0;
}

// sorted half-edge i -> SELL slot: column word (for the SpMV) and the 32-byte record k_linearize streams
__global__ void k_fill_halfedges(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const int32_t* __restrict__ slot,
                                 int32_t nh, const int32_t* __restrict__ ea, int32_t fixed, const double* __restrict__ tmx,
                                 const double* __restrict__ tmy, const double* __restrict__ thm,
                                 const uint8_t* __restrict__ dcs_flag, int32_t row_lo, int32_t row_hi, const int32_t* __restrict__ g2l,
                                 uint32_t* cols, HalfEdgeRec* recs, int32_t* edge_slot) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const uint32_t v = vals[i];
  const int32_t e = (int32_t)(v >> 1);
  const bool side_b = (v & 1u) != 0;
  const int32_t other = (int32_t)(keys[i] & kIdxMask);
  const int32_t row = (int32_t)(keys[i] >> 32);
  const int32_t a = ea[e];
  uint32_t word = (uint32_t)g2l[other];      // gathered arrays are indexed locally: [own rows | halo]
  if (side_b) word |= kFlagSideB;
  if (dcs_flag[e]) word |= kFlagDcs;
  if (other == fixed) word |= kFlagOtherFixed;
  // upper-triangular owner; blocks whose partner row lives on another rank are written on both ranks
  else if (row < other || other < row_lo || other >= row_hi) word |= kFlagOwner;
  // the edge's cost is booked on its a-side half-edge, or on the b side when a is constant
  const bool a_has_row = (a != fixed);
  if ((!side_b && a_has_row) || (side_b && !a_has_row)) word |= kFlagCost;
  const int32_t s = slot[i];
  cols[s] = word;
  HalfEdgeRec r;
  r.tmx = tmx[e]; r.tmy = tmy[e]; r.thm = thm[e]; r.word = word; r.word_next = 0u;   // word_next: k_task_walk
  recs[s] = r;
  edge_slot[2 * (int64_t)e + (side_b ? 1 : 0)] = s;
}

__global__ void k_fill_value(double* p, int64_t n, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// mirror_src[slot]: slot of the partner half-edge whose (owner) block this slot mirrors, or -1
__global__ void k_mirror_src(const uint32_t* __restrict__ vals, const int32_t* __restrict__ slot, int32_t nh,
                             const uint32_t* __restrict__ cols, const int32_t* __restrict__ edge_slot, int32_t* mirror_src) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int32_t s = slot[i];
  const uint32_t word = cols[s];
  int32_t src = -1;
  if (!(word & kFlagOwner) && !(word & kFlagOtherFixed)) {
    const uint32_t v = vals[i];
    src = edge_slot[2 * (int64_t)(v >> 1) + ((v & 1u) ^ 1u)];
  }
  mirror_src[s] = src;
}

// peers ask with global ids -> storage position of the own row
__global__ void k_global_to_own(int32_t* idx, int32_t n, int32_t row_lo, const uint16_t* rank_of) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = row_pos(rank_of, idx[i] - row_lo);
}

__global__ void k_is_free(const int32_t* __restrict__ deg_all, int32_t row_lo, int32_t nrows, int32_t fixed,
                          const uint16_t* __restrict__ rank_of, uint8_t* is_free) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  is_free[row_pos(rank_of, r)] = (deg_all[row_lo + r] > 0 && (row_lo + r) != fixed) ? 1 : 0;
}

// parity hook: canonical upper pattern values (block (row, col), row < col).  One thread per flagged (first-of-run)
// sorted half-edge; duplicate edges between the same pair add up.  The compact array holds every edge's block in
// the edge's own orientation H_ab, so a half-edge whose row is the b endpoint (an a > b edge) contributes the transpose.
__global__ void k_export_upper(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const int32_t* __restrict__ flag_scan,
                               const int32_t* __restrict__ flag, const int32_t* __restrict__ slot, int32_t nh,
                               const int32_t* __restrict__ block_src, const double* __restrict__ Hup, double* out /* n_upper x 9 */) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh || !flag[i]) return;
  const uint64_t k = keys[i];
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int32_t j = i; j < nh && keys[j] == k; ++j) {
    const int32_t src = block_src[slot[j]];
    if (src < 0) continue;
    const double* in = Hup + block_base(src);
    const bool side_b = (vals[j] & 1u) != 0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[3 * r + c] += side_b ? in[block_plane(3 * c + r) * 32] : in[block_plane(3 * r + c) * 32];
  }
  double* o = out + 9 * (int64_t)flag_scan[i];
#pragma unroll
  for (int c = 0; c < 9; ++c) o[c] = acc[c];
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
struct dcs_handle {
  dcs_options opt;
  Params P;
  int32_t N = 0, E = 0, fixed = 0;
  int dev = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;          // multi-rank: the halo push runs here while the SpMV's local pass runs on `stream`
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // partition (world == 1: everything)
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  int32_t rows_per_rank = 0, row_lo = 0, nrows = 0, Npad = 0;
  int32_t e_lo = 0, e_hi = 0;
  int32_t nblk = 0, nh = 0, nwin = 0, ntasks = 0;
  int64_t ldn = 0, ldh = 0;     // ldh: SELL slots = 32 * (n_slot_tiles + kTailTiles)
  int64_t n_slot_tiles = 0;
  // graph (device)
  DevBuf<int32_t> ea, eb, deg_all;
  DevBuf<double> e_tmx, e_tmy, e_thm;
  DevBuf<uint8_t> e_dcs, is_free;
  // pattern
  DevBuf<uint64_t> keys;        // sorted (row<<32|col)
  DevBuf<uint32_t> vals;        // edge<<1|side
  DevBuf<int32_t> scan_ws;      // block totals of the K0 scans (all recursion levels)
  DevBuf<int32_t> row_ptr, slot, up_flag, up_scan, block_src, task_tile0, task_obase;
  DevBuf<uint32_t> rank_info;
  DevBuf<int32_t> rank_nloc;    // per stored row: half-edges whose column is on this rank (they come first in the row)
  DevBuf<uint4> rowinfo;        // per stored row: degree + the words of rounds 0 and 1
  DevBuf<int2> task_info;       // per task: first tile, first compact owner-block index
  int64_t ldu = 32;            // compact owner-block leading dimension (owner half-edges, padded)
  DevBuf<double> Hup;          // compact edge blocks H_ab, tile-interleaved [ldu / 32][8][32], (task, round, lane) order
  bool mirrored = false;       // Hoff (slot order, both triangles) has been filled from Hup for the current linearization
  DevBuf<uint16_t> rank_of, perm;
  int32_t n_upper = 0;
  // half-edges (SELL slot order)
  DevBuf<uint32_t> cols;        // other pose | flags: what the SpMV streams
  DevBuf<HalfEdgeRec> recs;     // 32-byte records: what k_linearize / k_cost_rows stream
  // state
  DevBuf<double4> xyt, cand_xyt, p4;
  DevBuf<double> Hoff, Hdiag, grad, scale, lmdiag, Adiag, Minv, w, r, q, z, lambda_tmp, rhs_tmp;
  DevBuf<double> partials, scal, stage3;   // stage3: N x 3 staging for host<->device AoS
  DevBuf<double> rank_scal;                // [world][4] per-rank (cost, gsq, gmax) after k_linearize
  double* h_rank_scal = nullptr;
  double lin_cost = 0, lin_gsq = 0, lin_gmax = 0;
  bool lin_scal_pending = false;           // per-rank scalars not folded into h_scal yet
  DevBuf<double> fold_ws;                  // k_fold_tasks per-CTA partials [(K+M) <= 4][32]
  DevBuf<double> task_part;                // [3][ntasks] per-task partial sums of the row-owner kernels
  DevBuf<float> chL, chS;                  // chain-segment preconditioner factors (fp32), step-major
  DevBuf<int32_t> chain_idx, chain_cnt;    // (r, r+1) entries in the sorted half-edge list
  int ntiles = 0;                          // 1024-pose tiles of the chain preconditioner
  // halo exchange (world > 1): poses of other ranks my half-edges reference / my poses other ranks reference
  DevBuf<int32_t> halo_send_idx, halo_recv_idx, g2l;   // send: LOCAL own-row indices; recv: GLOBAL indices; g2l: global -> local
  DevBuf<double4> halo_send_buf;
  int32_t n_halo = 0, n_loc = 0;                       // local index space = [rows_per_rank own | n_halo halo]
  std::vector<int32_t> halo_send_off, halo_recv_off;   // [world+1] offsets per peer
  std::vector<int32_t> halo_cnt_all;                   // [world][world]: rank r needs halo_cnt_all[r*W+q] rows of rank q
  // peer-memory halo push (CUDA IPC): peer bases of the two exchanged arrays, or NCCL send/recv when unavailable
  bool halo_push = false;
  HaloPeers peers_p4 = {}, peers_xa = {}, peers_xb = {};   // p, and the two pose buffers (xyt / cand_xyt swap on acceptance)
  const double4 *ipc_xa = nullptr, *ipc_xb = nullptr;       // the pose buffers as they were when the handles were exchanged
  std::vector<void*> ipc_opened;
  DevBuf<float> barrier_buf;
  DevBuf<XchgBuf> xchg;                    // this rank's scalar-exchange buffer (peers store into it over NVLink)
  XchgPeers xchg_peers = {};
  int push_per_thread = 4;                 // halo entries per thread of k_halo_push (DCS_PUSH_PER_THREAD = 1 | 2 | 4)
  bool overlap_halo = false;               // DCS_OVERLAP=1: SpMV in two passes (local columns, then halo columns) around the halo push
                                           // on a second stream.  Off by default: measured at 2 x 1M poses the two short passes cost
                                           // 40 us more than the one-pass product, the push they hide is 18 us (profiles/r02_multigpu.md)
  bool xchg_ok = false;                    // peer-memory scalar all-reduce / barrier available (else ncclAllReduce)
  int64_t pcg_graph_launches = 0;          // kernel launches inside one replay of the PCG graph
  DevBuf<unsigned int> tickets;
  double* h_scal = nullptr;                // pinned mirror of scal
  double* h_pin3 = nullptr;                // pinned N x 3 staging
  size_t h_scal_cap = 0, h_rank_scal_cap = 0, h_pin3_cap = 0;      // capacities of the three pinned blocks (pinned_put)
  cudaGraphExec_t pcg_graph = nullptr;
  bool cluster_pcg = false;                // small single-rank graph with the chain preconditioner: k_pcg_cluster runs the whole solve
  int32_t cluster_cols_words = 0;          // column words k_pcg_cluster may stage in shared memory (0: read them from global memory)
  int pcg_graph_iters = 0;
  const double* pcg_graph_D = nullptr;
  bool have_lin = false;
  // METHOD 2 (switchable constraints, dcs_switchable.cuh): one switch per loop edge, eliminated inside the linear solve
  bool sc = false;
  double sc_lambda = 1.0;
  DevBuf<double> sw, sw_cand, sw_scale;    // [E] switch values (accepted / candidate), Jacobi scale of the switch columns
  DevBuf<double2> sw_slot;                 // [ldh] (s, scale) of the slot's edge: what k_linearize_sc streams
  DevBuf<int32_t> edge_slot;               // [2 E] slots of an edge's two half-edges (-1: none)
  DevBuf<double> Hdiag_lm, grad_full;      // unreduced pose diagonal blocks / gradient (Jacobi scale, LM diagonal, model cost)
  double eval_ms = 0, pcg_ms = 0;
  int64_t pcg_iters_total = 0;

  RowLayout layout() const {
    RowLayout L;
    L.nrows = nrows; L.ntasks = ntasks; L.ldn = ldn; L.ldh = ldh; L.ldu = ldu;
    L.rowinfo = rowinfo.p; L.task_info = task_info.p;
    return L;
  }
  EdgeList edgelist() const {
    EdgeList L;
    L.n = E; L.a = ea.p; L.b = eb.p; L.tmx = e_tmx.p; L.tmy = e_tmy.p; L.thm = e_thm.p; L.dcs = e_dcs.p;
    return L;
  }
  int vec_grid() const { return std::max(1, cdiv(nrows, kVecThreads)); }
};

namespace {

// Exclusive scan in place.  `ws` is the handle's scan workspace (block totals of every recursion level, sized once in
// dcs_create for the largest scan of the pattern build): no allocation and no host synchronisation per call, so the
// whole K0 build is stream-ordered between the few places that need a count on the host.
size_t scan_ws_size(int64_t n) {
  size_t t = 0;
  while (n > 1) { n = (n + kScanTile - 1) / kScanTile; t += (size_t)n; }
  return t + 8;
}
int scan_exclusive(int32_t* d, int64_t n, cudaStream_t st, int32_t* ws, size_t ws_cap) {
  if (n <= 0) return DCS_OK;
  const int nb = cdiv(n, kScanTile);
  if ((size_t)nb > ws_cap) { g_err = "scan_exclusive: workspace too small"; return DCS_ERR_ARG; }
  LAUNCH(k_scan_tile, nb, kScanThreads, st, d, n, ws);
  if (nb > 1) {
    CKS(scan_exclusive(ws, nb, st, ws + nb, ws_cap - (size_t)nb));
    LAUNCH(k_scan_add, nb, kScanThreads, st, d, n, ws);
  }
  return DCS_OK;
}

// stable LSD radix sort of (key,val) on the bit range [0,bits_lo) and [32,32+bits_hi)
int radix_sort(DevBuf<uint64_t>& keys, DevBuf<uint32_t>& vals, int64_t n, int bits_lo, int bits_hi, cudaStream_t st,
               int32_t* ws, size_t ws_cap) {
  if (n <= 1) return DCS_OK;
  DevBuf<uint64_t> k2;
  DevBuf<uint32_t> v2;
  CK(k2.alloc((size_t)n));
  CK(v2.alloc((size_t)n));
  const int nblk = cdiv(n, kSortTile);
  DevBuf<int32_t> hist;
  CK(hist.alloc((size_t)256 * nblk));
  uint64_t* ki = keys.p; uint64_t* ko = k2.p;
  uint32_t* vi = vals.p; uint32_t* vo = v2.p;
  std::vector<int> shifts;
  for (int s = 0; s < bits_lo; s += 8) shifts.push_back(s);
  for (int s = 0; s < bits_hi; s += 8) shifts.push_back(32 + s);
  for (int shift : shifts) {
    LAUNCH(k_radix_hist, nblk, kSortThreads, st, ki, n, shift, hist.p, nblk);
    CKS(scan_exclusive(hist.p, (int64_t)256 * nblk, st, ws, ws_cap));
    LAUNCH(k_radix_scatter, nblk, kSortThreads, st, ki, vi, ko, vo, n, shift, hist.p, nblk);
    std::swap(ki, ko);
    std::swap(vi, vo);
  }
  CK(cudaStreamSynchronize(st));   // the scratch pair is freed on return
  if (ki != keys.p) {   // odd number of passes: result lives in the scratch pair
    std::swap(keys.p, k2.p);
    std::swap(vals.p, v2.p);
  }
  CK(cudaGetLastError());
  return DCS_OK;
}

int bits_for(int32_t n) { int b = 1; while ((1ll << b) < (long long)n) ++b; return b; }

int allreduce_sum(dcs_handle* h, double* d, int count) {
  if (h->world == 1) return DCS_OK;
  if (h->xchg_ok && count <= kXchgVals) {
    LAUNCH(k_xchg_sum, 1, 32, h->stream, h->xchg_peers, d, count);
    return DCS_OK;
  }
  CKN(nccl_api().AllReduce(d, d, (size_t)count, ncclDouble, ncclSum, h->comm, h->stream));
  return DCS_OK;
}
// Halo exchange of a locally indexed double4 array ([own rows | halo]) whose own rows just changed: every rank packs
// the own entries its peers reference; grouped ncclSend/ncclRecv over NVLink move them straight into the peers' halo
// regions (the halo is stored in global order, i.e. grouped by owner, so no unpack pass is needed).
int halo_exchange(dcs_handle* h, double4* arr, cudaStream_t st = nullptr) {
  if (h->world == 1) return DCS_OK;
  if (!st) st = h->stream;
  const int32_t ns = h->halo_send_off[h->world];
  if (h->halo_push && (arr == h->p4.p || arr == h->ipc_xa || arr == h->ipc_xb)) {
    // every rank takes the same accept/reject decisions, so "my buffer A" is "buffer A" on every peer
    const HaloPeers& P = (arr == h->p4.p) ? h->peers_p4 : (arr == h->ipc_xa ? h->peers_xa : h->peers_xb);
    if (ns > 0) {
      if (h->push_per_thread == 1) LAUNCH(k_halo_push<1>, cdiv(ns, 256), 256, st, arr, h->halo_send_idx.p, ns, P);
      else if (h->push_per_thread == 2) LAUNCH(k_halo_push<2>, cdiv(ns, 512), 256, st, arr, h->halo_send_idx.p, ns, P);
      else LAUNCH(k_halo_push<4>, cdiv(ns, 1024), 256, st, arr, h->halo_send_idx.p, ns, P);
    }
    // every rank's pushes have landed once all ranks passed this point of their streams
    if (h->xchg_ok) LAUNCH(k_xchg_sum, 1, 32, st, h->xchg_peers, (double*)nullptr, 0);
    else CKN(nccl_api().AllReduce(h->barrier_buf.p, h->barrier_buf.p, 1, ncclFloat, ncclSum, h->comm, st));
    return DCS_OK;
  }
  st = h->stream;     // the NCCL send/recv fallback stays on the main stream
  if (ns > 0) LAUNCH(k_halo_pack, cdiv(ns, 256), 256, h->stream, arr, h->halo_send_idx.p, ns, h->halo_send_buf.p);
  CKN(nccl_api().GroupStart());
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    const int32_t cs = h->halo_send_off[r + 1] - h->halo_send_off[r], cr = h->halo_recv_off[r + 1] - h->halo_recv_off[r];
    if (cs > 0) CKN(nccl_api().Send(h->halo_send_buf.p + h->halo_send_off[r], (size_t)cs * 4, ncclDouble, r, h->comm, h->stream));
    if (cr > 0) CKN(nccl_api().Recv(arr + h->rows_per_rank + h->halo_recv_off[r], (size_t)cr * 4, ncclDouble, r, h->comm, h->stream));
  }
  CKN(nccl_api().GroupEnd());
  return DCS_OK;
}

// one-time (world > 1): open every peer's p / candidate-pose arrays through CUDA IPC so the halo can be pushed
// with plain stores over NVLink.  All ranks agree (all-reduce of a flag) on push vs the NCCL send/recv path.
int setup_halo_push(dcs_handle* h) {
  const int W = h->world;
  h->halo_push = false;
  if (W == 1) return DCS_OK;
  CK(h->barrier_buf.alloc_zero(4, h->stream));
  cudaStream_t st = h->stream;
  const char* mode = std::getenv("DCS_HALO");
  int ok = (!(mode && std::strcmp(mode, "nccl") == 0)) ? 1 : 0;
  struct Pair { cudaIpcMemHandle_t p4, xa, xb, xc; };
  CK(h->xchg.alloc_zero(1, h->stream));
  const char* smode = std::getenv("DCS_SCALARS");        // DCS_SCALARS=nccl: keep ncclAllReduce for the scalars / barrier
  const bool want_xchg = !(smode && std::strcmp(smode, "nccl") == 0);
  std::vector<Pair> all((size_t)W);
  DevBuf<unsigned char> d_all;
  CK(d_all.alloc(sizeof(Pair) * (size_t)W));
  Pair mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaIpcGetMemHandle(&mine.p4, h->p4.p) != cudaSuccess || cudaIpcGetMemHandle(&mine.xa, h->xyt.p) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.xb, h->cand_xyt.p) != cudaSuccess || cudaIpcGetMemHandle(&mine.xc, h->xchg.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  CK(cudaMemcpyAsync(d_all.p + sizeof(Pair) * (size_t)h->rank, &mine, sizeof(Pair), cudaMemcpyHostToDevice, st));
  CKN(nccl_api().AllGather(d_all.p + sizeof(Pair) * (size_t)h->rank, d_all.p, sizeof(Pair), ncclChar, h->comm, st));
  CK(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Pair) * (size_t)W, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->ipc_xa = h->xyt.p; h->ipc_xb = h->cand_xyt.p;
  HaloPeers P4 = {}, PA = {}, PB = {};
  P4.world = PA.world = PB.world = W;
  XchgPeers XP = {};
  XP.world = W; XP.rank = h->rank; XP.buf[h->rank] = h->xchg.p;
  static_assert(sizeof(P4.send_off) / sizeof(P4.send_off[0]) == kMaxWorld + 1 && sizeof(P4.ptr) / sizeof(P4.ptr[0]) == kMaxWorld,
                "HaloPeers is sized for kMaxWorld ranks");
  if (W > kMaxWorld) { g_err = "setup_halo_push: world exceeds kMaxWorld"; return DCS_ERR_ARG; }   // dcs_create rejects it first
  for (int r = 0; r <= W; ++r) P4.send_off[r] = PA.send_off[r] = PB.send_off[r] = h->halo_send_off[r];
  for (int r = 0; r < W && ok; ++r) {
    if (r == h->rank) continue;
    void *a = nullptr, *b = nullptr, *c = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[r].p4, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    h->ipc_opened.push_back(a);
    if (cudaIpcOpenMemHandle(&b, all[r].xa, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    h->ipc_opened.push_back(b);
    if (cudaIpcOpenMemHandle(&c, all[r].xb, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    h->ipc_opened.push_back(c);
    void* d = nullptr;
    if (cudaIpcOpenMemHandle(&d, all[r].xc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    h->ipc_opened.push_back(d);
    XP.buf[r] = static_cast<XchgBuf*>(d);
    P4.ptr[r] = static_cast<double4*>(a);
    PA.ptr[r] = static_cast<double4*>(b);
    PB.ptr[r] = static_cast<double4*>(c);
    int32_t before = 0;     // rank r's halo is grouped by owner: my slice starts after the owners below me
    for (int q = 0; q < h->rank; ++q) before += h->halo_cnt_all[(size_t)r * W + q];
    P4.dst_base[r] = PA.dst_base[r] = PB.dst_base[r] = h->rows_per_rank + before;
  }
  // agree: push only if every rank could open every peer
  float flag = ok ? 0.f : 1.f;
  CK(cudaMemcpyAsync(h->barrier_buf.p + 1, &flag, 4, cudaMemcpyHostToDevice, st));
  CKN(nccl_api().AllReduce(h->barrier_buf.p + 1, h->barrier_buf.p + 1, 1, ncclFloat, ncclSum, h->comm, st));
  CK(cudaMemcpyAsync(&flag, h->barrier_buf.p + 1, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->halo_push = (flag == 0.f);
  h->peers_p4 = P4; h->peers_xa = PA; h->peers_xb = PB;
  h->xchg_peers = XP;
  if (const char* ppt = std::getenv("DCS_PUSH_PER_THREAD")) h->push_per_thread = std::atoi(ppt);
  if (const char* ov = std::getenv("DCS_OVERLAP")) h->overlap_halo = std::atoi(ov) != 0;
  h->xchg_ok = h->halo_push && want_xchg;     // DCS_SCALARS must be set the same on every rank
  if (h->opt.verbose && h->rank == 0)
    std::fprintf(stderr, "[dcs] halo exchange: %s; scalars: %s\n", h->halo_push ? "peer-memory push (CUDA IPC over NVLink)" : "ncclSend/ncclRecv",
                 h->xchg_ok ? "peer-memory flags" : "ncclAllReduce");
  return DCS_OK;
}

// one-time: build the halo lists from the sorted half-edge keys and swap them with the peers
int build_halo(dcs_handle* h, int32_t nh) {
  const int W = h->world;
  h->halo_send_off.assign(W + 1, 0);
  h->halo_recv_off.assign(W + 1, 0);
  cudaStream_t st = h->stream;
  const int32_t NP = h->Npad;
  CK(h->g2l.alloc((size_t)NP));
  DevBuf<int32_t> need, scan;
  CK(need.alloc_zero((size_t)NP + 1, st));
  CK(scan.alloc((size_t)NP + 1));
  if (nh > 0) LAUNCH(k_halo_mark, cdiv(nh, 256), 256, st, h->keys.p, nh, h->row_lo, h->row_lo + h->rows_per_rank, need.p);
  CK(cudaMemcpyAsync(scan.p, need.p, ((size_t)NP + 1) * 4, cudaMemcpyDeviceToDevice, st));
  CKS(scan_exclusive(scan.p, (int64_t)NP + 1, st, h->scan_ws.p, h->scan_ws.n));
  // receive offsets per owner rank = scan at the rank boundaries
  std::vector<int32_t> bnd(W + 1);
  for (int r = 0; r <= W; ++r) CK(cudaMemcpyAsync(&bnd[r], scan.p + (size_t)r * h->rows_per_rank, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int r = 0; r <= W; ++r) h->halo_recv_off[r] = bnd[r];
  const int32_t nr = bnd[W];
  h->n_halo = nr;
  h->n_loc = h->rows_per_rank + nr;
  CK(h->halo_recv_idx.alloc((size_t)std::max(nr, 1)));
  LAUNCH(k_halo_compact, cdiv(NP, 256), 256, st, need.p, scan.p, NP, h->row_lo, h->rows_per_rank, h->rank_of.p, h->halo_recv_idx.p, h->g2l.p);
  if (W == 1) { CK(cudaStreamSynchronize(st)); return DCS_OK; }
  // counts: all-gather the per-owner receive counts of every rank, read column `rank` = what I must send to whom
  DevBuf<int32_t> cnt_all;
  CK(cnt_all.alloc((size_t)W * W));
  std::vector<int32_t> my_cnt(W);
  for (int r = 0; r < W; ++r) my_cnt[r] = bnd[r + 1] - bnd[r];
  CK(cudaMemcpyAsync(cnt_all.p + (size_t)h->rank * W, my_cnt.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
  CKN(nccl_api().AllGather(cnt_all.p + (size_t)h->rank * W, cnt_all.p, (size_t)W, ncclInt32, h->comm, st));
  std::vector<int32_t> all((size_t)W * W);
  CK(cudaMemcpyAsync(all.data(), cnt_all.p, all.size() * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int r = 0; r < W; ++r) h->halo_send_off[r + 1] = h->halo_send_off[r] + all[(size_t)r * W + h->rank];   // rank r needs that many of my rows
  h->halo_cnt_all = all;
  const int32_t ns = h->halo_send_off[W];
  CK(h->halo_send_idx.alloc((size_t)std::max(ns, 1)));
  // Swap the index lists with ONE all-gather of every rank's receive list (global ids, grouped by owner, padded to the
  // longest list): what rank r needs from me is a contiguous slice of r's list.  (A grouped ncclSend/ncclRecv to every
  // peer would make NCCL set up a point-to-point connection per peer pair here: seconds at create time.)
  {
    int32_t max_nr = 1;
    for (int r = 0; r < W; ++r) {
      int32_t t = 0;
      for (int q = 0; q < W; ++q) t += all[(size_t)r * W + q];
      max_nr = std::max(max_nr, t);
    }
    DevBuf<int32_t> lists;
    CK(lists.alloc((size_t)W * max_nr));
    if (nr > 0) CK(cudaMemcpyAsync(lists.p + (size_t)h->rank * max_nr, h->halo_recv_idx.p, (size_t)nr * 4, cudaMemcpyDeviceToDevice, st));
    CKN(nccl_api().AllGather(lists.p + (size_t)h->rank * max_nr, lists.p, (size_t)max_nr, ncclInt32, h->comm, st));
    for (int r = 0; r < W; ++r) {
      if (r == h->rank) continue;
      const int32_t cs = h->halo_send_off[r + 1] - h->halo_send_off[r];
      int32_t off = 0;      // rank r's list is grouped by owner: my slice starts after the owners below me
      for (int q = 0; q < h->rank; ++q) off += all[(size_t)r * W + q];
      if (cs > 0) CK(cudaMemcpyAsync(h->halo_send_idx.p + h->halo_send_off[r], lists.p + (size_t)r * max_nr + off, (size_t)cs * 4, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaStreamSynchronize(st));     // `lists` is freed here
  }
  CK(h->halo_send_buf.alloc((size_t)std::max(ns, 1)));
  if (ns > 0) LAUNCH(k_global_to_own, cdiv(ns, 256), 256, st, h->halo_send_idx.p, ns, h->row_lo, h->rank_of.p);   // peers asked with global ids
  CK(cudaStreamSynchronize(st));
  return DCS_OK;
}

int read_scalars(dcs_handle* h) {
  CK(cudaMemcpyAsync(h->h_scal, h->scal.p, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (h->lin_scal_pending)
    CK(cudaMemcpyAsync(h->h_rank_scal, h->rank_scal.p, (size_t)h->world * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->lin_scal_pending) {
    double c = 0, g2 = 0, gm = 0;
    for (int r = 0; r < h->world; ++r) { c += h->h_rank_scal[4 * r]; g2 += h->h_rank_scal[4 * r + 1]; gm = std::max(gm, h->h_rank_scal[4 * r + 2]); }
    h->lin_cost = c; h->lin_gsq = g2; h->lin_gmax = gm;
    h->lin_scal_pending = false;
  }
  if (h->world > 1) { h->h_scal[S_COST] = h->lin_cost; h->h_scal[S_GSQ] = h->lin_gsq; h->h_scal[S_GMAX] = h->lin_gmax; }
  return DCS_OK;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// host N x 3 poses -> device packed poses (own rows + halo, local order).  With several ranks every rank uploads
// only its own rows (N / world of the array) and the halo comes from the owners over NVLink, like every other
// halo exchange; the call is collective, as dcs_linearize / dcs_cost / dcs_solve already are.
int upload_poses(dcs_handle* h, const double* pose_xyt, double4* xyt) {
  const int32_t own_lo = h->world > 1 ? h->row_lo : 0;
  const int32_t own_n = h->world > 1 ? std::max(0, std::min(h->rows_per_rank, h->N - h->row_lo)) : h->N;
  const double* src = pose_xyt + 3 * (size_t)own_lo;
  if (own_n > 0) {
    if (!is_pinned(pose_xyt)) {       // pageable caller memory: bounce through the handle's pinned buffer
      std::memcpy(h->h_pin3, src, (size_t)own_n * 3 * sizeof(double));
      src = h->h_pin3;
    }
    CK(cudaMemcpyAsync(h->stage3.p + 3 * (size_t)own_lo, src, (size_t)own_n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  const int32_t n_pack = h->world > 1 ? h->rows_per_rank : h->n_loc;
  LAUNCH(k_pack_poses, cdiv(n_pack, 256), 256, h->stream, h->stage3.p, h->N, h->row_lo, h->rows_per_rank, h->halo_recv_idx.p,
         h->rank_of.p, n_pack, xyt);
  CKS(halo_exchange(h, xyt));
  return DCS_OK;
}
// own rows of every rank -> host N x 3
int download_poses(dcs_handle* h, const double4* xyt, double* pose_xyt) {
  LAUNCH(k_unpack_poses, cdiv(h->rows_per_rank, 256), 256, h->stream, xyt, h->rank_of.p, h->row_lo, h->rows_per_rank, h->N, h->stage3.p);
  if (h->world > 1) {     // stage3 is Npad x 3: equal slices, in-place all-gather
    char* base = reinterpret_cast<char*>(h->stage3.p);
    const size_t chunk = (size_t)h->rows_per_rank * 24;
    CKN(nccl_api().AllGather(base + (size_t)h->rank * chunk, base, chunk, ncclChar, h->comm, h->stream));
  }
  CK(cudaMemcpyAsync(h->h_pin3, h->stage3.p, (size_t)h->N * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::memcpy(pose_xyt, h->h_pin3, (size_t)h->N * 3 * sizeof(double));
  return DCS_OK;
}

// K1+K2 at the given packed poses; results in Hoff / Hdiag / grad, scalars S_COST, S_GSQ, S_GMAX
int sc_args(const dcs_handle* h, double inv_radius, int reduce, ScArgs* A) {
  A->lambda = h->sc_lambda; A->inv_radius = inv_radius; A->dmin = h->opt.min_lm_diagonal; A->dmax = h->opt.max_lm_diagonal; A->reduce = reduce;
  return DCS_OK;
}
// METHOD 2: the row-owner pass at (xyt, sw).  reduce = 0: the pose blocks of J^T J and J^T r as they are (+ cost and
// the gradient norms over poses and switches); reduce = 1: the switches eliminated for the step at this radius.
int linearize_sc(dcs_handle* h, const double4* xyt, double inv_radius, int reduce) {
  ScArgs A;
  sc_args(h, inv_radius, reduce, &A);
  if (h->E > 0) LAUNCH(k_sc_to_slots, cdiv(h->E, 256), 256, h->stream, h->edge_slot.p, h->e_dcs.p, h->E, h->sw.p, h->sw_scale.p, h->sw_slot.p);
  LAUNCH(k_linearize_sc, h->nblk, kRowsPerBlock, h->stream, xyt, h->layout(), h->recs.p, h->sw_slot.p, h->P, A, h->n_loc, h->Hup.p,
         h->Hdiag.p, h->grad.p, h->task_part.p);
  if (!reduce) {
    k_fold_tasks<2, 1><<<fold_blocks(h->nblk), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->nblk, h->scal.p + S_COST, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
    ++g_launches; ++t_launches;
  }
  h->have_lin = true;
  h->mirrored = false;
  return DCS_OK;
}

int linearize(dcs_handle* h, const double4* xyt) {
  if (h->sc) return linearize_sc(h, xyt, 0.0, 0);
  LAUNCH(k_linearize, h->nblk, kRowsPerBlock, h->stream, xyt, h->layout(), h->recs.p, h->P, h->n_loc, h->Hup.p, h->Hdiag.p,
         h->grad.p, h->task_part.p);
  k_fold_tasks<2, 1><<<fold_blocks(h->nblk), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->nblk, h->scal.p + S_COST, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
  ++g_launches; ++t_launches;
  if (h->world > 1) {   // one collective: every rank's (cost, |g|^2, |g|_inf); folded on the host in rank order
    CKN(nccl_api().AllGather(h->scal.p + S_COST, h->rank_scal.p, 4, ncclDouble, h->comm, h->stream));
    h->lin_scal_pending = true;
  }
  h->have_lin = true;
  h->mirrored = false;
  return DCS_OK;
}

// linear-solver setup: fill the slot-order block storage (both triangles) the row-wise SpMV reads from the compact
// edge blocks
int ensure_mirror(dcs_handle* h) {
  if (!h->mirrored) {
    LAUNCH(k_expand<double>, cdiv(h->ldh, 256), 256, h->stream, h->block_src.p, h->cols.p, h->ldh, h->Hup.p, h->Hoff.p);
    h->mirrored = true;
  }
  return DCS_OK;
}

int cost_only(dcs_handle* h, const double4* xyt, int slot, const double* sw = nullptr) {
  if (h->sc) {     // METHOD 2: edge-order pass at (xyt, sw); results in S_SC_COST (cost) and S_SC_COST + 1 (sum s^2)
    ScArgs A;
    sc_args(h, 0.0, 0, &A);
    LAUNCH(k_sc_edges<kScCost>, std::max(1, cdiv(h->E, kEdgeThreads)), kEdgeThreads, h->stream, xyt, h->g2l.p, h->edgelist(), h->P, A, 0,
           sw ? sw : h->sw.p, (double*)nullptr, (const double*)nullptr, h->ldn, (double*)nullptr, h->partials.p, h->tickets.p + 4,
           h->scal.p + S_SC_COST);
    return DCS_OK;
  }
  LAUNCH(k_cost_rows, h->nblk, kRowsPerBlock, h->stream, xyt, h->layout(), h->recs.p, h->P, h->n_loc, h->task_part.p);
  k_fold_tasks<1, 0><<<fold_blocks(h->nblk), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->nblk, h->scal.p + slot, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
  ++g_launches; ++t_launches;
  CKS(allreduce_sum(h, h->scal.p + slot, 1));
  return DCS_OK;
}

// q = (off-diagonal blocks + D) p over the rank's rows, p.q -> scal[out_slot] (summed over ranks).  p's own rows are
// current; with several ranks the halo entries are fetched here: the peers' pushes (second stream) overlap the pass
// over the local columns, the halo columns follow once every rank's pushes have landed.  Capturable.
int spmv_product(dcs_handle* h, const double* D, int out_slot, int rotate_rz) {
  cudaStream_t st = h->stream;
  if (h->world == 1) {
    LAUNCH((k_spmv<double, kSpmvAll>), h->nblk, kRowsPerBlock, st, h->p4.p, h->layout(), h->cols.p, h->Hoff.p, D, h->n_loc, h->q.p, h->task_part.p);
  } else if (h->halo_push && h->overlap_halo) {
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    CKS(halo_exchange(h, h->p4.p, h->stream2));
    CK(cudaEventRecord(h->ev_join, h->stream2));
    LAUNCH((k_spmv<double, kSpmvLocal>), h->nblk, kRowsPerBlock, st, h->p4.p, h->layout(), h->cols.p, h->Hoff.p, D, h->n_loc, h->q.p, h->task_part.p);
    CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    LAUNCH((k_spmv<double, kSpmvHalo>), h->nblk, kRowsPerBlock, st, h->p4.p, h->layout(), h->cols.p, h->Hoff.p, D, h->n_loc, h->q.p, h->task_part.p);
  } else {
    CKS(halo_exchange(h, h->p4.p));
    LAUNCH((k_spmv<double, kSpmvAll>), h->nblk, kRowsPerBlock, st, h->p4.p, h->layout(), h->cols.p, h->Hoff.p, D, h->n_loc, h->q.p, h->task_part.p);
  }
  k_fold_tasks<1, 0><<<fold_blocks(h->nblk), kFoldThreads, 0, st>>>(h->task_part.p, h->nblk, h->scal.p + out_slot, h->scal.p, rotate_rz, h->fold_ws.p, h->tickets.p + 6);
  ++g_launches; ++t_launches;
  CKS(allreduce_sum(h, h->scal.p + out_slot, 1));
  return DCS_OK;
}

// one PCG iteration on the stream (capturable).  Starts with the halo exchange of the direction the previous
// iteration (or the initialisation) left in the own rows of p.
int pcg_iteration(dcs_handle* h, const double* D) {
  CKS(spmv_product(h, D, S_PQ, 1));
  if (h->opt.preconditioner == 1) {
    k_pcg_chain<false><<<h->ntiles, kChainThreads, kChainSmemBytes, h->stream>>>((const double*)nullptr, h->is_free.p, h->p4.p, h->q.p, h->chL.p, h->chS.p,
        h->perm.p, 0, h->nrows, h->ldn, h->w.p, h->r.p, h->z.p, h->p4.p, h->task_part.p, h->scal.p);
    k_fold_tasks<2, 0><<<fold_blocks(h->ntiles), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->ntiles, h->scal.p + S_TMP, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
    g_launches += 2; t_launches += 2;
  } else {
    LAUNCH(k_pcg_update, h->vec_grid(), kVecThreads, h->stream, h->p4.p, h->q.p, h->Minv.p, 0, h->nrows, h->ldn, h->w.p,
           h->r.p, h->z.p, h->partials.p, h->tickets.p + 4, h->scal.p);
  }
  CKS(allreduce_sum(h, h->scal.p + S_TMP, 2));
  LAUNCH(k_pcg_direction, h->vec_grid(), kVecThreads, h->stream, h->z.p, 0, h->nrows, h->ldn, h->p4.p, h->scal.p);
  return DCS_OK;
}

// Solve (H + Lambda) w = rhs.  Lambda from (lmdiag, scale, radius) or explicit.  rhs: device SoA.
int pcg_solve(dcs_handle* h, double inv_radius, const double* lambda_explicit, const double* rhs, int* iters_out,
              double* relres_out) {
  CK(cudaEventRecord(h->ev0, h->stream));
  CKS(ensure_mirror(h));
  LAUNCH(k_precond, h->vec_grid(), 256, h->stream, h->Hdiag.p, h->lmdiag.p, h->scale.p, h->is_free.p, h->nrows, h->ldn, inv_radius,
         lambda_explicit, h->Adiag.p, h->Minv.p);
  if (h->cluster_pcg) {     // small graph: factorise, then ONE cluster launch runs the solve to convergence
    LAUNCH(k_chain_factor<double>, h->ntiles, 32, h->stream, h->Adiag.p, h->Hoff.p, h->slot.p, h->chain_idx.p, h->chain_cnt.p, h->rank_of.p, h->nrows, h->ldn,
           h->ldh, h->chL.p, h->chS.p);
    const int batch = std::max(1, h->opt.pcg_check_every);
    const int max_iter = cdiv(std::max(1, h->opt.pcg_max_iter), batch) * batch;     // whole batches, like the graph replays below
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)h->ntiles); cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = kClSmemBase + (size_t)h->cluster_cols_words * 4; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)h->ntiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_pcg_cluster, rhs, (const uint8_t*)h->is_free.p, h->layout(), (const uint32_t*)h->cols.p, (const double*)h->Hoff.p,
                          (const double*)h->Adiag.p, (const float*)h->chL.p, (const float*)h->chS.p, (const uint16_t*)h->perm.p, h->n_loc,
                          h->cluster_cols_words, (int32_t)batch, (int32_t)max_iter, h->opt.pcg_rel_tol, h->p4.p, h->w.p, h->scal.p));
    ++g_launches; ++t_launches;
    CKS(read_scalars(h));
    const double rr0 = h->h_scal[S_RR0], rr = h->h_scal[S_RR];
    const int iters = (int)h->h_scal[S_PCG_ITERS];
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->pcg_ms += ms;
    h->pcg_iters_total += iters;
    if (iters_out) *iters_out = iters;
    if (relres_out) *relres_out = (rr0 > 0.0) ? std::sqrt(rr / rr0) : 0.0;
    return DCS_OK;
  }
  if (h->opt.preconditioner == 1) {
    LAUNCH(k_chain_factor<double>, h->ntiles, 32, h->stream, h->Adiag.p, h->Hoff.p, h->slot.p, h->chain_idx.p, h->chain_cnt.p, h->rank_of.p, h->nrows, h->ldn,
           h->ldh, h->chL.p, h->chS.p);
    k_pcg_chain<true><<<h->ntiles, kChainThreads, kChainSmemBytes, h->stream>>>(rhs, h->is_free.p, h->p4.p, h->q.p, h->chL.p, h->chS.p, h->perm.p, 0, h->nrows,
        h->ldn, h->w.p, h->r.p, h->z.p, h->p4.p, h->task_part.p, h->scal.p);
    ++g_launches; ++t_launches;
    k_fold_tasks<2, 0><<<fold_blocks(h->ntiles), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->ntiles, h->scal.p + S_TMP, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
    ++g_launches; ++t_launches;
  } else
  LAUNCH(k_pcg_init, h->vec_grid(), kVecThreads, h->stream, rhs, h->Minv.p, h->is_free.p, 0, h->nrows, h->ldn, h->w.p, h->r.p,
         h->z.p, h->p4.p, h->partials.p, h->tickets.p + 4, h->scal.p);
  CKS(allreduce_sum(h, h->scal.p + S_TMP, 2));
  LAUNCH(k_pcg_init_finish, 1, 1, h->stream, h->scal.p);
  CKS(read_scalars(h));       // (the halo of p is fetched by the first product)
  const double rr0 = h->h_scal[S_RR0];
  int iters = 0;
  double rr = rr0;
  if (rr0 > 0.0 && std::isfinite(rr0)) {
    const int batch = std::max(1, h->opt.pcg_check_every);
    if (!h->pcg_graph || h->pcg_graph_iters != batch || h->pcg_graph_D != h->Adiag.p) {
      if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
      cudaGraph_t g = nullptr;
      const int64_t launches_before = t_launches;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      int st = DCS_OK;
      for (int i = 0; i < batch && st == DCS_OK; ++i) st = pcg_iteration(h, h->Adiag.p);
      cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
      if (st != DCS_OK) { if (g) cudaGraphDestroy(g); return st; }
      CK(ce);
      CK(cudaGraphInstantiate(&h->pcg_graph, g, 0));
      CK(cudaGraphDestroy(g));
      h->pcg_graph_iters = batch;
      h->pcg_graph_D = h->Adiag.p;
      h->pcg_graph_launches = t_launches - launches_before;      // capture does not launch: counted per replay
      g_launches -= h->pcg_graph_launches;
    }
    const double target = h->opt.pcg_rel_tol * h->opt.pcg_rel_tol * rr0;
    while (iters < h->opt.pcg_max_iter) {
      CK(cudaGraphLaunch(h->pcg_graph, h->stream));
      g_launches += h->pcg_graph_launches;
      iters += batch;
      CKS(read_scalars(h));
      rr = h->h_scal[S_RR];
      if (!(rr > target)) break;   // converged, or NaN (caller validates the step)
    }
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->pcg_ms += ms;
  h->pcg_iters_total += iters;
  if (iters_out) *iters_out = iters;
  if (relres_out) *relres_out = (rr0 > 0.0) ? std::sqrt(rr / rr0) : 0.0;
  return DCS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

void dcs_options_default(dcs_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->dcs_on = 1;
  o->phi = 0.5;
  o->huber_delta = 0.01;
  o->max_num_iterations = 50;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->pcg_rel_tol = 1e-12;
  o->pcg_max_iter = 200000;
  o->pcg_check_every = 32;
  o->preconditioner = 1;
  o->device = 0;
  o->verbose = 0;
  o->rank = 0;
  o->world = 1;
  o->nccl_unique_id = nullptr;
  o->max_solver_time_s = 1e6;
  o->switchable_on = 0;
  o->switch_prior_lambda = 1.0;
}

void* dcs_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; cudaGetLastError(); return nullptr; }
  return p;
}
void dcs_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* dcs_version(void) { return "dcs_b200 0.1 (sm_100a)"; }
const char* dcs_last_error(void) { return g_err.c_str(); }
int64_t dcs_launch_count(int reset) { const int64_t v = g_launches.load(); if (reset) g_launches = 0; return v; }

int dcs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int dcs_partition(int32_t n_poses, int32_t n_edges, int32_t rank, int32_t world, int32_t out[5]) {
  if (!out || n_poses <= 0 || n_edges < 0 || world < 1 || rank < 0 || rank >= world) return DCS_ERR_ARG;
  const int64_t tile = (int64_t)kWindow * world;
  const int64_t npad = ((int64_t)n_poses + tile - 1) / tile * tile;
  const int32_t rpr = (int32_t)(npad / world);
  const int32_t row_lo = rank * rpr;
  out[0] = row_lo;
  out[1] = std::max(0, std::min(n_poses, row_lo + rpr) - row_lo);
  out[2] = rpr;
  const int64_t epr = ((int64_t)n_edges + world - 1) / world;
  out[3] = (int32_t)std::min<int64_t>(n_edges, epr * rank);
  out[4] = (int32_t)std::min<int64_t>(n_edges, epr * (rank + 1));
  return DCS_OK;
}

int dcs_nccl_unique_id(void* out128) {
  if (!out128) return DCS_ERR_ARG;
  if (!nccl_api().load()) { g_err = "libnccl.so.2 not found"; return DCS_ERR_NCCL; }
  ncclUniqueId id;
  CKN(nccl_api().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return DCS_OK;
}

void dcs_destroy(dcs_handle* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
  for (void* q : h->ipc_opened) cudaIpcCloseMemHandle(q);
  if (h->comm) nccl_api().CommDestroy(h->comm);
  if (h->stream) cudaStreamSynchronize(h->stream);      // nothing in flight may still write the pinned blocks a batch worker reuses
  pinned_put(h->h_pin3, h->h_pin3_cap);
  pinned_put(h->h_scal, h->h_scal_cap);
  pinned_put(h->h_rank_scal, h->h_rank_scal_cap);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaStream_t s1 = h->stream, s2 = h->stream2;
  delete h;                 // the device buffers first: pool memory is released in the order of the handle's stream
  if (s2) cudaStreamDestroy(s2);
  if (s1) cudaStreamDestroy(s1);
}

int dcs_create(const dcs_graph* g, const dcs_options* o, dcs_handle** out) {
  if (!g || !o || !out || g->n_poses <= 0 || g->n_edges < 0 || !g->pose_xyt) { g_err = "dcs_create: bad argument"; return DCS_ERR_ARG; }
  if (g->n_edges > 0 && (!g->edge_a || !g->edge_b || !g->meas_xyt || !g->kind)) { g_err = "dcs_create: null edge array"; return DCS_ERR_ARG; }
  if ((uint32_t)g->n_poses > kIdxMask) { g_err = "dcs_create: too many poses"; return DCS_ERR_ARG; }
  if (g->fixed_pose < -1 || g->fixed_pose >= g->n_poses) { g_err = "dcs_create: fixed_pose must be -1 (none) or a pose index"; return DCS_ERR_ARG; }
  if (o->switchable_on && o->dcs_on) { g_err = "dcs_create: dcs_on and switchable_on are exclusive (METHOD 1 vs METHOD 2)"; return DCS_ERR_ARG; }
  if (o->switchable_on && o->world > 1) { g_err = "dcs_create: switchable constraints (METHOD 2) run on single-rank handles only"; return DCS_ERR_ARG; }
  if (o->switchable_on && !(o->switch_prior_lambda > 0.0)) { g_err = "dcs_create: switch_prior_lambda must be positive"; return DCS_ERR_ARG; }
  if (o->world > kMaxWorld) { g_err = "dcs_create: world sizes above " + std::to_string(kMaxWorld) + " (one NVSwitch node) are not supported"; return DCS_ERR_ARG; }
  for (int32_t k = 0; k < g->n_edges; ++k) {
    const int32_t a = g->edge_a[k], b = g->edge_b[k];
    if (a < 0 || b < 0 || a >= g->n_poses || b >= g->n_poses || a == b) {
      g_err = "dcs_create: edge " + std::to_string(k) + " has an invalid endpoint pair";   // Ceres aborts on a == b
      return DCS_ERR_ARG;
    }
  }
  if (g->n_edges >= (1 << 22)) {   // rank_info packs the degree into 22 bits
    std::vector<int32_t> deg((size_t)g->n_poses, 0);
    for (int32_t k = 0; k < g->n_edges; ++k) { ++deg[g->edge_a[k]]; ++deg[g->edge_b[k]]; }
    for (int32_t d : deg)
      if (d >= (1 << 22)) { g_err = "dcs_create: a pose with 2^22 or more edges is not supported"; return DCS_ERR_ARG; }
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (o->device < 0 || o->device >= ndev) { g_err = "dcs_create: no such CUDA device"; return DCS_ERR_CUDA; }
  CK(cudaSetDevice(o->device));
  int sm_count_ = 148;
  cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, o->device);
  CK(cudaFuncSetAttribute(k_pcg_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
  CK(cudaFuncSetAttribute(k_pcg_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
  // two 85-KB tiles per SM: ask for the largest shared-memory carve-out (the default heuristic sizes it for one CTA)
  CK(cudaFuncSetAttribute(k_pcg_chain<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClSmemMax));
  CK(cudaFuncSetAttribute(k_pcg_chain<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));

  const double t_c0 = now_s();
  auto lap = [&](const char* what) { if (std::getenv("DCS_CREATE_TIMING")) { cudaDeviceSynchronize(); std::fprintf(stderr, "[dcs_create] %-28s %.3f s\n", what, now_s() - t_c0); } };
  dcs_handle* h = new dcs_handle();
  struct Guard { dcs_handle* h; ~Guard() { if (h) dcs_destroy(h); } } guard{h};
  h->opt = *o;
  h->opt.nccl_unique_id = nullptr;
  h->dev = o->device;
  h->sm_count = sm_count_;
  h->P.phi = o->phi; h->P.hub_a = o->huber_delta; h->P.hub_b = o->huber_delta * o->huber_delta;
  h->N = g->n_poses; h->E = g->n_edges; h->fixed = g->fixed_pose;
  h->rank = o->rank; h->world = std::max(1, o->world);
  h->sc = o->switchable_on != 0; h->sc_lambda = o->switch_prior_lambda;
  if (h->rank < 0 || h->rank >= h->world) { g_err = "dcs_create: bad rank"; return DCS_ERR_ARG; }
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  struct PoolScope { ~PoolScope() { tl_pool_stream = nullptr; } } pool_scope;      // every return below leaves pool mode
  if (tl_batch_worker && o->world <= 1) tl_pool_stream = h->stream;
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  cudaStream_t st = h->stream;

  if (h->world > 1) {
    if (!o->nccl_unique_id) { g_err = "dcs_create: world > 1 needs nccl_unique_id"; return DCS_ERR_ARG; }
    if (!nccl_api().load()) { g_err = "libnccl.so.2 not found"; return DCS_ERR_NCCL; }
    ncclUniqueId id;
    std::memcpy(&id, o->nccl_unique_id, sizeof(id));
    CKN(nccl_api().CommInitRank(&h->comm, h->world, id, h->rank));
  }

  // contiguous pose ranges, equal-sized (multiple of the CTA row tile) so in-place all-gathers work
  const int32_t N = h->N, E = h->E;
  int32_t part[5];
  if (dcs_partition(N, E, h->rank, h->world, part) != DCS_OK) { g_err = "dcs_create: bad partition"; return DCS_ERR_ARG; }
  h->row_lo = part[0]; h->nrows = part[1]; h->rows_per_rank = part[2];
  h->Npad = h->rows_per_rank * h->world;
  h->nwin = std::max(1, h->rows_per_rank / kWindow);
  h->ntasks = h->nwin * kSlicesPerWindow;
  h->nblk = h->ntasks;                                // CTAs of the row-owner kernels (one warp task each)
  h->ldn = (int64_t)h->nwin * kWindow;
  h->e_lo = part[3]; h->e_hi = part[4];

  lap("stream/nccl");
  // ---- upload the graph ---------------------------------------------------------------------
  DevBuf<double> d_meas;
  DevBuf<uint8_t> d_kind;
  CK(h->ea.alloc((size_t)std::max(E, 1))); CK(h->eb.alloc((size_t)std::max(E, 1)));
  CK(d_meas.alloc((size_t)std::max(E, 1) * 3)); CK(d_kind.alloc((size_t)std::max(E, 1)));
  if (E > 0) {
    CK(cudaMemcpyAsync(h->ea.p, g->edge_a, (size_t)E * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->eb.p, g->edge_b, (size_t)E * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_meas.p, g->meas_xyt, (size_t)E * 24, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_kind.p, g->kind, (size_t)E, cudaMemcpyHostToDevice, st));
  }
  const size_t EE = (size_t)std::max(E, 1);
  CK(h->e_tmx.alloc(EE)); CK(h->e_tmy.alloc(EE)); CK(h->e_thm.alloc(EE));
  CK(h->e_dcs.alloc(EE));
  if (E > 0) LAUNCH(k_edge_prep, cdiv(E, 256), 256, st, d_meas.p, d_kind.p, E, o->dcs_on || o->switchable_on, h->e_tmx.p, h->e_tmy.p, h->e_thm.p,
                    h->e_dcs.p);

  lap("upload + edge prep");
  // ---- K0: half-edges, sort, CSR, jagged-diagonal re-layout ---------------------------------------
  {
    const int64_t max_scan = std::max<int64_t>({2 * (int64_t)E + 2, (int64_t)h->Npad + 2, 256 * ((int64_t)cdiv(2 * (int64_t)E + 1, kSortTile) + 1)});
    CK(h->scan_ws.alloc(scan_ws_size(max_scan)));
  }
  CK(h->deg_all.alloc_zero((size_t)h->Npad, st));
  if (E > 0) LAUNCH(k_pose_degree, cdiv(E, 256), 256, st, h->ea.p, h->eb.p, E, h->deg_all.p);
  CK(h->is_free.alloc_zero((size_t)h->ldn, st));

  DevBuf<int32_t> he_off;
  CK(he_off.alloc_zero((size_t)E + 1, st));
  const int32_t row_hi = h->row_lo + h->nrows;
  if (E > 0) LAUNCH(k_halfedge_count, cdiv(E, 256), 256, st, h->ea.p, h->eb.p, E, h->fixed, h->row_lo, row_hi, he_off.p);
  CKS(scan_exclusive(he_off.p, (int64_t)E + 1, st, h->scan_ws.p, h->scan_ws.n));
  int32_t nh = 0;
  CK(cudaMemcpyAsync(&nh, he_off.p + E, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->nh = nh;
  h->ldh = ((int64_t)std::max(nh, 1) + 31) / 32 * 32;
  CK(h->keys.alloc((size_t)std::max(nh, 1)));
  CK(h->vals.alloc((size_t)std::max(nh, 1)));
  if (E > 0) LAUNCH(k_halfedge_fill, cdiv(E, 256), 256, st, h->ea.p, h->eb.p, E, h->fixed, h->row_lo, row_hi, he_off.p, h->keys.p, h->vals.p);
  const int nb = bits_for(std::max(N, 2));
  CKS(radix_sort(h->keys, h->vals, nh, 28, nb, st, h->scan_ws.p, h->scan_ws.n));   // column word: 27 index bits + the owner-order bit

  lap("half-edges + radix sort");
  CK(h->row_ptr.alloc((size_t)h->ldn + 1));
  {
    // rows beyond nrows (padding) get the end offset so their degree is zero
    const int32_t rows = (int32_t)h->ldn;
    LAUNCH(k_row_ptr, cdiv(rows + 1, 256), 256, st, h->keys.p, nh, h->row_lo, rows, h->row_ptr.p);
  }
  CK(h->rank_of.alloc((size_t)h->ldn)); CK(h->perm.alloc((size_t)h->ldn));
  CK(h->rank_info.alloc((size_t)h->ldn)); CK(h->rank_nloc.alloc((size_t)h->ldn));
  LAUNCH(k_jds_rank, h->nwin, kWindow, st, h->row_ptr.p, h->keys.p, (int32_t)h->ldn, h->world > 1 ? 1 : 0, h->rank_of.p, h->perm.p,
         h->rank_info.p, h->rank_nloc.p);
  // per-row arrays live in (window, rank) order from here on
  if (h->nrows > 0) LAUNCH(k_is_free, cdiv(h->nrows, 256), 256, st, h->deg_all.p, h->row_lo, h->nrows, h->fixed, h->rank_of.p, h->is_free.p);
  // SELL tiles: a task's tile count = its largest degree (first lane: ranks are degree-sorted); tile0 = exclusive scan
  CK(h->task_tile0.alloc_zero((size_t)h->ntasks + 1, st));
  LAUNCH(k_task_kmax, cdiv(h->ntasks, 256), 256, st, h->ntasks, h->rank_info.p, h->task_tile0.p);
  CKS(scan_exclusive(h->task_tile0.p, (int64_t)h->ntasks + 1, st, h->scan_ws.p, h->scan_ws.n));
  int32_t n_tiles = 0;
  CK(cudaMemcpyAsync(&n_tiles, h->task_tile0.p + h->ntasks, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if ((int64_t)n_tiles + kTailTiles >= (int64_t)1 << 26) { g_err = "dcs_create: more than 2^31 half-edge slots on one rank"; return DCS_ERR_ARG; }
  h->n_slot_tiles = n_tiles;
  h->ldh = ((int64_t)n_tiles + kTailTiles) * kSlice;
  CK(h->slot.alloc((size_t)std::max(nh, 1)));
  if (nh > 0) LAUNCH(k_sell_slot, cdiv(nh, 256), 256, st, h->keys.p, nh, h->row_lo, h->row_ptr.p, h->rank_of.p, h->task_tile0.p, h->slot.p);

  CKS(build_halo(h, nh));      // halo lists + the global -> local index map the half-edge words use
  const size_t HH = (size_t)h->ldh;
  CK(h->cols.alloc_zero(HH, st)); CK(h->recs.alloc_zero(HH, st));          // padding slots: word 0, zero record
  CK(h->block_src.alloc(HH));
  CK(cudaMemsetAsync(h->block_src.p, 0xFF, HH * 4, st));                    // -1: no block
  CK(h->rowinfo.alloc_zero((size_t)h->ldn, st));
  CK(h->task_obase.alloc_zero((size_t)h->ntasks + 1, st));
  CK(h->task_info.alloc((size_t)h->ntasks + 1));
  {
    DevBuf<int32_t> mirror_src, cidx, edge_slot;
    CK(mirror_src.alloc(HH)); CK(cidx.alloc(HH));
    CK(edge_slot.alloc((size_t)2 * EE));
    CK(cudaMemsetAsync(edge_slot.p, 0xFF, (size_t)2 * EE * 4, st));
    CK(cudaMemsetAsync(mirror_src.p, 0xFF, HH * 4, st));
    if (nh > 0) {
      LAUNCH(k_fill_halfedges, cdiv(nh, 256), 256, st, h->keys.p, h->vals.p, h->slot.p, nh, h->ea.p, h->fixed, h->e_tmx.p,
             h->e_tmy.p, h->e_thm.p, h->e_dcs.p, h->row_lo, row_hi, h->g2l.p, h->cols.p, h->recs.p, edge_slot.p);
      LAUNCH(k_mirror_src, cdiv(nh, 256), 256, st, h->vals.p, h->slot.p, nh, h->cols.p, edge_slot.p, mirror_src.p);
    }
    // compact owner-block order = the order k_linearize meets the owner half-edges in
    LAUNCH(k_task_walk<false>, h->ntasks, kRowsPerBlock, st, h->ntasks, h->rank_info.p, h->rank_nloc.p, h->task_tile0.p, (const int32_t*)nullptr,
           h->cols.p, h->task_obase.p, (int32_t*)nullptr, (HalfEdgeRec*)nullptr, (uint4*)nullptr);
    CKS(scan_exclusive(h->task_obase.p, (int64_t)h->ntasks + 1, st, h->scan_ws.p, h->scan_ws.n));
    int32_t n_own = 0;
    CK(cudaMemcpyAsync(&n_own, h->task_obase.p + h->ntasks, 4, cudaMemcpyDeviceToHost, st));
    LAUNCH(k_task_walk<true>, h->ntasks, kRowsPerBlock, st, h->ntasks, h->rank_info.p, h->rank_nloc.p, h->task_tile0.p, h->task_obase.p, h->cols.p,
           (int32_t*)nullptr, cidx.p, h->recs.p, h->rowinfo.p);
    LAUNCH(k_task_info, cdiv(h->ntasks + 1, 256), 256, st, h->ntasks, h->task_tile0.p, h->task_obase.p, h->task_info.p);
    if (nh > 0) LAUNCH(k_block_src, cdiv(nh, 256), 256, st, h->cols.p, h->slot.p, mirror_src.p, cidx.p, nh, h->block_src.p);
    CK(cudaStreamSynchronize(st));     // n_own on the host; the scratch arrays are freed here
    h->ldu = ((int64_t)std::max(n_own, 1) + 31) / 32 * 32;
    if (h->sc) { std::swap(h->edge_slot.p, edge_slot.p); std::swap(h->edge_slot.n, edge_slot.n); }
  }

  lap("sell layout + fill");
  // unique upper pattern (parity hook)
  CK(h->up_flag.alloc_zero((size_t)nh + 1, st)); CK(h->up_scan.alloc_zero((size_t)nh + 1, st));
  if (nh > 0) {
    LAUNCH(k_upper_flag, cdiv(nh, 256), 256, st, h->keys.p, nh, h->fixed, h->up_flag.p);
    CK(cudaMemcpyAsync(h->up_scan.p, h->up_flag.p, (size_t)(nh + 1) * 4, cudaMemcpyDeviceToDevice, st));
    CKS(scan_exclusive(h->up_scan.p, (int64_t)nh + 1, st, h->scan_ws.p, h->scan_ws.n));
    CK(cudaMemcpyAsync(&h->n_upper, h->up_scan.p + nh, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }

  lap("upper pattern");
  // ---- state --------------------------------------------------------------------------------------
  const size_t LN = (size_t)h->ldn;
  const size_t NL = (size_t)std::max(h->n_loc, 1);
  CK(h->xyt.alloc_zero(NL, st)); CK(h->cand_xyt.alloc_zero(NL, st)); CK(h->p4.alloc_zero(NL, st));
  CK(h->Hoff.alloc_zero(kBlockVals * HH, st)); CK(h->Hup.alloc_zero(kBlockVals * (size_t)h->ldu, st)); CK(h->Hdiag.alloc_zero(6 * LN, st)); CK(h->grad.alloc_zero(3 * LN, st));
  CK(h->scale.alloc_zero(3 * LN, st)); CK(h->lmdiag.alloc_zero(3 * LN, st)); CK(h->Adiag.alloc_zero(6 * LN, st)); CK(h->Minv.alloc_zero(6 * LN, st));
  CK(h->w.alloc_zero(3 * LN, st)); CK(h->r.alloc_zero(3 * LN, st)); CK(h->q.alloc_zero(3 * LN, st)); CK(h->z.alloc_zero(3 * LN, st));
  CK(h->lambda_tmp.alloc_zero(3 * LN, st)); CK(h->rhs_tmp.alloc_zero(3 * LN, st));
  const size_t max_grid = (size_t)std::max<int64_t>({(int64_t)h->nblk, (int64_t)h->vec_grid(), 148 * 8, 2 * (int64_t)cdiv(E, kEdgeThreads)});
  CK(h->partials.alloc_zero(4 * max_grid, st));
  CK(h->scal.alloc_zero(S_COUNT, st));
  CK(h->tickets.alloc_zero(8, st));
  CK(h->fold_ws.alloc_zero(4 * kFoldMaxBlocks, st));
  h->ntiles = (int)(h->ldn / kChainTile);
  {   // small single-rank graphs: the whole PCG solve in one cluster launch (dcs_pcg_cluster.cuh).  DCS_PCG_CLUSTER=0 keeps
      // the general path (CUDA-graph batches of five kernels per iteration)
    const char* ev = std::getenv("DCS_PCG_CLUSTER");
    const bool want = !(ev && std::atoi(ev) == 0);
    if (want && h->world == 1 && h->opt.preconditioner == 1 && h->ntiles >= 1 && h->ntiles <= kClMaxTiles) {
      h->cluster_cols_words = h->ldh <= (int64_t)kClMaxColWords ? (int32_t)h->ldh : 0;
      // Can the device co-schedule a cluster of this many CTAs?  Asked once per (device, cluster size) with the largest
      // shared-memory request a launch can make and remembered: the occupancy query costs milliseconds and is serialised
      // by the driver, and dcs_solve_batch creates a handle per item from several host threads.
      static std::mutex mu;
      static signed char known[64][kClMaxTiles + 1];        // 0: not asked yet, 1: yes, -1: no
      const int dslot = h->dev & 63;
      std::lock_guard<std::mutex> lock(mu);
      if (known[dslot][h->ntiles] == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->ntiles); cfg.blockDim = dim3(kClThreads);
        cfg.dynamicSmemBytes = kClSmemMax;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)h->ntiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, k_pcg_cluster, &cfg) == cudaSuccess && n_clusters >= 1) known[dslot][h->ntiles] = 1;
        else { (void)cudaGetLastError(); known[dslot][h->ntiles] = -1; }      // cannot co-schedule the cluster: general path
      }
      h->cluster_pcg = known[dslot][h->ntiles] == 1;
    }
  }
  CK(h->chL.alloc_zero(9 * LN, st)); CK(h->chS.alloc_zero(6 * LN, st));
  CK(h->chain_idx.alloc_zero(LN, st)); CK(h->chain_cnt.alloc_zero(LN, st));
  if (h->nrows > 0) LAUNCH(k_chain_entries, cdiv(h->nrows, 256), 256, st, h->keys.p, nh, h->row_lo, h->nrows, h->chain_idx.p, h->chain_cnt.p);
  CK(h->task_part.alloc_zero((size_t)3 * std::max(h->nblk, (int32_t)(h->ldn / kChainTile) + 1), st));
  CK(h->stage3.alloc_zero((size_t)h->Npad * 3, st));
  CK(pinned_get((void**)&h->h_scal, S_COUNT * sizeof(double), &h->h_scal_cap));
  CK(h->rank_scal.alloc_zero((size_t)h->world * 4, st));
  CK(pinned_get((void**)&h->h_rank_scal, (size_t)h->world * 4 * sizeof(double), &h->h_rank_scal_cap));
  CK(pinned_get((void**)&h->h_pin3, (size_t)N * 3 * sizeof(double), &h->h_pin3_cap));
  if (h->sc) {
    CK(h->sw.alloc(EE)); CK(h->sw_cand.alloc(EE)); CK(h->sw_scale.alloc(EE));
    CK(h->sw_slot.alloc_zero(HH, st));
    CK(h->Hdiag_lm.alloc_zero(6 * LN, st)); CK(h->grad_full.alloc_zero(3 * LN, st));
    LAUNCH(k_fill_value, cdiv((int64_t)EE, 256), 256, st, h->sw.p, (int64_t)EE, 1.0);
    LAUNCH(k_fill_value, cdiv((int64_t)EE, 256), 256, st, h->sw_cand.p, (int64_t)EE, 1.0);
    LAUNCH(k_fill_value, cdiv((int64_t)EE, 256), 256, st, h->sw_scale.p, (int64_t)EE, 1.0);
  }
  lap("state alloc");
  CKS(setup_halo_push(h));
  CKS(upload_poses(h, g->pose_xyt, h->xyt.p));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  lap("pose upload");
  guard.h = nullptr;
  *out = h;
  return DCS_OK;
}

int dcs_evaluate(dcs_handle* h, const double* pose_xyt, double* cost, double* residuals, double* jacobians, double* psi,
                 double* rho1, double* gradient) {
  if (!h) return DCS_ERR_ARG;
  CK(cudaSetDevice(h->dev));
  if (pose_xyt) CKS(upload_poses(h, pose_xyt, h->xyt.p));
  CKS(linearize(h, h->xyt.p));
  CKS(read_scalars(h));
  if (cost) *cost = h->h_scal[S_COST];
  const int32_t E = h->E;
  if ((residuals || jacobians || psi || rho1) && h->sc) {
    g_err = "dcs_evaluate: the per-edge dump evaluates the METHOD 0/1 functors; not available on switchable_on handles";
    return DCS_ERR_ARG;
  }
  if ((residuals || jacobians || psi || rho1) && h->world > 1) {
    g_err = "dcs_evaluate: per-edge outputs are available on single-rank handles only";
    return DCS_ERR_ARG;
  }
  if ((residuals || jacobians || psi || rho1) && E > 0) {
    DevBuf<double> dr, dj, dp, dq;
    if (residuals) CK(dr.alloc((size_t)E * 3));
    if (jacobians) CK(dj.alloc((size_t)E * 18));
    if (psi) CK(dp.alloc((size_t)E));
    if (rho1) CK(dq.alloc((size_t)E));
    LAUNCH(k_edge_eval, cdiv(E, kEdgeThreads), kEdgeThreads, h->stream, h->xyt.p, h->g2l.p, h->edgelist(), h->P, dr.p, dj.p, dp.p, dq.p);
    if (residuals) CK(cudaMemcpyAsync(residuals, dr.p, (size_t)E * 24, cudaMemcpyDeviceToHost, h->stream));
    if (jacobians) CK(cudaMemcpyAsync(jacobians, dj.p, (size_t)E * 144, cudaMemcpyDeviceToHost, h->stream));
    if (psi) CK(cudaMemcpyAsync(psi, dp.p, (size_t)E * 8, cudaMemcpyDeviceToHost, h->stream));
    if (rho1) CK(cudaMemcpyAsync(rho1, dq.p, (size_t)E * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (gradient) {
    std::memset(gradient, 0, (size_t)h->N * 24);
    if (h->nrows > 0) {
      LAUNCH(k_soa_to_aos, cdiv(h->nrows, 256), 256, h->stream, h->grad.p, h->rank_of.p, h->nrows, h->ldn, h->stage3.p);
      CK(cudaMemcpyAsync(gradient + 3 * (size_t)h->row_lo, h->stage3.p, (size_t)h->nrows * 24, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
  }
  CK(cudaGetLastError());
  return DCS_OK;
}

int dcs_linearize(dcs_handle* h, const double* pose_xyt, double* cost, double* gradient) {
  if (!h || !pose_xyt) return DCS_ERR_ARG;
  CK(cudaSetDevice(h->dev));
  CKS(upload_poses(h, pose_xyt, h->xyt.p));
  CKS(linearize(h, h->xyt.p));
  const bool direct = gradient && is_pinned(gradient);
  if (gradient && h->world > 1) std::memset(gradient, 0, (size_t)h->N * 24);
  if (gradient && h->nrows > 0) {
    LAUNCH(k_soa_to_aos, cdiv(h->nrows, 256), 256, h->stream, h->grad.p, h->rank_of.p, h->nrows, h->ldn, h->stage3.p);
    CK(cudaMemcpyAsync(direct ? gradient + 3 * (size_t)h->row_lo : h->h_pin3, h->stage3.p, (size_t)h->nrows * 24,
                       cudaMemcpyDeviceToHost, h->stream));
  }
  CKS(read_scalars(h));
  if (cost) *cost = h->h_scal[S_COST];
  if (gradient && !direct) std::memcpy(gradient + 3 * (size_t)h->row_lo, h->h_pin3, (size_t)h->nrows * 24);
  return DCS_OK;
}

int dcs_linearize_resident(dcs_handle* h, int32_t repeats, int32_t with_solver_setup, float* ms_total) {
  if (!h || repeats <= 0) return DCS_ERR_ARG;
  CK(cudaSetDevice(h->dev));
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < repeats; ++i) {
    CKS(linearize(h, h->xyt.p));
    if (with_solver_setup) CKS(ensure_mirror(h));   // + the row-storage expansion the PCG needs (once per LM iteration)
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (ms_total) *ms_total = ms;
  CK(cudaGetLastError());
  return DCS_OK;
}

int dcs_cost(dcs_handle* h, const double* pose_xyt, double* cost) {
  if (!h || !cost) return DCS_ERR_ARG;
  CK(cudaSetDevice(h->dev));
  const double4* x = h->xyt.p;
  if (pose_xyt) { CKS(upload_poses(h, pose_xyt, h->cand_xyt.p)); x = h->cand_xyt.p; }
  CKS(cost_only(h, x, S_CAND_COST));
  CKS(read_scalars(h));
  *cost = h->h_scal[S_CAND_COST];
  return DCS_OK;
}

// host copy of row_pos for the single-rank parity hooks: natural row r -> storage position
static int host_row_positions(dcs_handle* h, std::vector<int32_t>* pos) {
  std::vector<uint16_t> rk((size_t)h->ldn);
  CK(cudaMemcpy(rk.data(), h->rank_of.p, rk.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  pos->resize((size_t)h->N);
  for (int32_t r = 0; r < h->N; ++r) (*pos)[r] = (r & ~(kWindow - 1)) + (int32_t)rk[r];
  return DCS_OK;
}

int dcs_get_pattern(dcs_handle* h, int32_t* n_block_rows, int32_t* nnzb, int32_t* row_ptr, int32_t* col_idx) {
  if (!h) return DCS_ERR_ARG;
  if (h->world != 1) { g_err = "dcs_get_pattern: single-rank handles only"; return DCS_ERR_ARG; }
  CK(cudaSetDevice(h->dev));
  std::vector<uint8_t> is_free_st((size_t)h->ldn), is_free((size_t)h->N);
  CK(cudaMemcpy(is_free_st.data(), h->is_free.p, is_free_st.size(), cudaMemcpyDeviceToHost));
  std::vector<int32_t> rpos;
  CKS(host_row_positions(h, &rpos));
  int32_t n_diag = 0;
  for (int32_t i = 0; i < h->N; ++i) { is_free[i] = is_free_st[rpos[i]]; n_diag += is_free[i]; }
  if (n_block_rows) *n_block_rows = h->N;
  if (nnzb) *nnzb = n_diag + h->n_upper;
  if (!row_ptr || !col_idx) return DCS_OK;
  // unique upper (row,col) keys come off the device already sorted by (row, col)
  std::vector<uint64_t> keys((size_t)h->nh);
  std::vector<int32_t> flag((size_t)h->nh);
  if (h->nh > 0) {
    CK(cudaMemcpy(keys.data(), h->keys.p, (size_t)h->nh * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(flag.data(), h->up_flag.p, (size_t)h->nh * 4, cudaMemcpyDeviceToHost));
  }
  int32_t pos = 0;
  size_t i = 0;
  for (int32_t r = 0; r < h->N; ++r) {
    row_ptr[r] = pos;
    if (is_free[r]) col_idx[pos++] = r;
    while (i < keys.size() && (int32_t)(keys[i] >> 32) == r) {
      if (flag[i]) col_idx[pos++] = (int32_t)(keys[i] & kIdxMask);
      ++i;
    }
  }
  row_ptr[h->N] = pos;
  return DCS_OK;
}

int dcs_get_hessian(dcs_handle* h, double* block_values) {
  if (!h || !block_values) return DCS_ERR_ARG;
  if (h->world != 1) { g_err = "dcs_get_hessian: single-rank handles only"; return DCS_ERR_ARG; }
  if (!h->have_lin) { g_err = "dcs_get_hessian: nothing linearized yet"; return DCS_ERR_ARG; }
  CK(cudaSetDevice(h->dev));
  const int32_t N = h->N, nh = h->nh;
  std::vector<uint8_t> is_free_st((size_t)h->ldn);
  CK(cudaMemcpy(is_free_st.data(), h->is_free.p, is_free_st.size(), cudaMemcpyDeviceToHost));
  std::vector<int32_t> rpos;
  CKS(host_row_positions(h, &rpos));
  std::vector<double> hd((size_t)6 * h->ldn), up((size_t)std::max(h->n_upper, 1) * 9);
  CK(cudaMemcpy(hd.data(), h->Hdiag.p, hd.size() * 8, cudaMemcpyDeviceToHost));
  std::vector<uint64_t> keys((size_t)nh);
  std::vector<int32_t> flag((size_t)nh);
  if (nh > 0) {
    DevBuf<double> d_up;
    CK(d_up.alloc_zero(up.size(), h->stream));
    LAUNCH(k_export_upper, cdiv(nh, 256), 256, h->stream, h->keys.p, h->vals.p, h->up_scan.p, h->up_flag.p, h->slot.p, nh, h->block_src.p,
           h->Hup.p, d_up.p);
    CK(cudaMemcpyAsync(up.data(), d_up.p, up.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(keys.data(), h->keys.p, (size_t)nh * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(flag.data(), h->up_flag.p, (size_t)nh * 4, cudaMemcpyDeviceToHost));
  }
  int64_t pos = 0, u = 0;
  size_t i = 0;
  for (int32_t r = 0; r < N; ++r) {
    const int64_t m = rpos[r];          // per-row arrays are stored in (window, rank) order
    if (is_free_st[m]) {
      const double d00 = hd[0 * h->ldn + m], d01 = hd[1 * h->ldn + m], d02 = hd[2 * h->ldn + m];
      const double d11 = hd[3 * h->ldn + m], d12 = hd[4 * h->ldn + m], d22 = hd[5 * h->ldn + m];
      const double blk[9] = {d00, d01, d02, d01, d11, d12, d02, d12, d22};
      std::memcpy(block_values + 9 * pos, blk, sizeof(blk));
      ++pos;
    }
    while (i < keys.size() && (int32_t)(keys[i] >> 32) == r) {
      if (flag[i]) { std::memcpy(block_values + 9 * pos, &up[9 * u], 72); ++pos; ++u; }
      ++i;
    }
  }
  return DCS_OK;
}

int dcs_pcg_solve(dcs_handle* h, const double* lambda, const double* rhs, double* w, int32_t* iterations, double* rel_residual) {
  if (!h || !rhs || !w) return DCS_ERR_ARG;
  if (!h->have_lin) { g_err = "dcs_pcg_solve: nothing linearized yet"; return DCS_ERR_ARG; }
  CK(cudaSetDevice(h->dev));
  const int32_t nr = h->nrows;
  auto up3 = [&](const double* src, double* dst_soa) -> int {
    std::memcpy(h->h_pin3, src + 3 * (size_t)h->row_lo, (size_t)nr * 24);
    CK(cudaMemcpyAsync(h->stage3.p, h->h_pin3, (size_t)nr * 24, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_aos_to_soa, cdiv(nr, 256), 256, h->stream, h->stage3.p, h->rank_of.p, nr, h->ldn, dst_soa);
    CK(cudaStreamSynchronize(h->stream));
    return DCS_OK;
  };
  if (nr > 0) {
    if (lambda) CKS(up3(lambda, h->lambda_tmp.p)); else CK(cudaMemsetAsync(h->lambda_tmp.p, 0, 3 * (size_t)h->ldn * 8, h->stream));
    CKS(up3(rhs, h->rhs_tmp.p));
  }
  int iters = 0;
  double rel = 0;
  CKS(pcg_solve(h, 0.0, h->lambda_tmp.p, h->rhs_tmp.p, &iters, &rel));
  std::memset(w, 0, (size_t)h->N * 24);
  if (nr > 0) {
    LAUNCH(k_soa_to_aos, cdiv(nr, 256), 256, h->stream, h->w.p, h->rank_of.p, nr, h->ldn, h->stage3.p);
    CK(cudaMemcpyAsync(h->h_pin3, h->stage3.p, (size_t)nr * 24, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::memcpy(w + 3 * (size_t)h->row_lo, h->h_pin3, (size_t)nr * 24);
  }
  if (iterations) *iterations = iters;
  if (rel_residual) *rel_residual = rel;
  return DCS_OK;
}

// Ceres-default trust-region Levenberg–Marquardt (TrustRegionMinimizer + LevenbergMarquardtStrategy),
// control on the host, all arithmetic on the device.  Works in unscaled variables: with the Jacobi
// scaling S, Ceres solves (S H S + D^2) y = S g, step = -y, delta = S step; substituting w = S y gives
// (H + D^2 S^-2) w = g and delta = -w, and model_cost_change = w.g - w.H.w / 2.
int dcs_solve(dcs_handle* h, double* pose_xyt_inout, dcs_summary* sum, dcs_iteration* trace, int32_t trace_cap) {
  if (!h || !pose_xyt_inout || !sum) return DCS_ERR_ARG;
  CK(cudaSetDevice(h->dev));
  std::memset(sum, 0, sizeof(*sum));
  const dcs_options& o = h->opt;
  const double t_start = now_s();
  h->eval_ms = 0; h->pcg_ms = 0; h->pcg_iters_total = 0;
  cudaStream_t st = h->stream;
  float ms = 0;

  auto timed_linearize = [&](const double4* x) -> int {
    CK(cudaEventRecord(h->ev0, st));
    CKS(linearize(h, x));
    if (h->sc) {   // the reduced pass of every LM step overwrites Hdiag / grad: keep the pose blocks' diagonal and J_p^T r
      CK(cudaMemcpyAsync(h->Hdiag_lm.p, h->Hdiag.p, 6 * (size_t)h->ldn * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(h->grad_full.p, h->grad.p, 3 * (size_t)h->ldn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaEventRecord(h->ev1, st));
    CKS(read_scalars(h));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->eval_ms += ms;
    return DCS_OK;
  };

  CKS(upload_poses(h, pose_xyt_inout, h->xyt.p));
  LAUNCH(k_xnorm, h->vec_grid(), kVecThreads, st, h->xyt.p, h->is_free.p, 0, h->nrows, h->partials.p, h->tickets.p + 5, h->scal.p);
  CKS(allreduce_sum(h, h->scal.p + S_XSQ, 1));
  const double* Hdiag_lm = h->sc ? h->Hdiag_lm.p : h->Hdiag.p;     // pose diagonal blocks of J^T J (METHOD 2: before the elimination)
  const double* grad_full = h->sc ? h->grad_full.p : h->grad.p;
  if (h->sc) {     // every solve starts with all switches at 1 (main.cpp:117,139) and unit column scales
    const int64_t EE = std::max(h->E, 1);
    LAUNCH(k_fill_value, cdiv(EE, 256), 256, st, h->sw.p, EE, 1.0);
    LAUNCH(k_fill_value, cdiv(EE, 256), 256, st, h->sw_scale.p, EE, 1.0);
    CKS(cost_only(h, h->xyt.p, S_CAND_COST));                        // S_SC_COST + 1 = sum s^2 for |x|
  }
  CKS(timed_linearize(h->xyt.p));
  double x_cost = h->h_scal[S_COST];
  double x_norm = std::sqrt(h->h_scal[S_XSQ] + (h->sc ? h->h_scal[S_SC_COST + 1] : 0.0));
  if (!std::isfinite(x_cost)) {
    sum->termination_type = DCS_FAILURE;
    std::snprintf(sum->message, sizeof(sum->message), "Initial cost is not finite.");
    g_err = sum->message;
    return DCS_ERR_NUMERIC;
  }
  LAUNCH(k_jacobi_scale, h->vec_grid(), 256, st, Hdiag_lm, h->nrows, h->ldn, h->scale.p, o.jacobi_scaling);
  if (h->sc) {
    ScArgs A;
    sc_args(h, 0.0, 0, &A);
    LAUNCH(k_sc_edges<kScScale>, std::max(1, cdiv(h->E, kEdgeThreads)), kEdgeThreads, st, h->xyt.p, h->g2l.p, h->edgelist(), h->P, A,
           o.jacobi_scaling, h->sw.p, h->sw_scale.p, (const double*)nullptr, h->ldn, (double*)nullptr, h->partials.p, h->tickets.p + 4,
           h->scal.p + S_SC);
  }

  int n_logged = 0;
  double min_logged_cost = std::numeric_limits<double>::max();
  auto log_iter = [&](const dcs_iteration& it) {
    if (trace && n_logged < trace_cap) trace[n_logged] = it;
    ++n_logged;
    min_logged_cost = std::min(min_logged_cost, it.cost);
    if (o.verbose && h->rank == 0) {
      if (it.iteration == 0)
        std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");
      std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration, it.cost, it.cost_change,
                  it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius, it.linear_solver_iterations,
                  it.iteration_time_s, it.cumulative_time_s);
      std::fflush(stdout);
    }
  };

  dcs_iteration it;
  std::memset(&it, 0, sizeof(it));
  it.cost = x_cost;
  it.gradient_max_norm = h->h_scal[S_GMAX];
  it.gradient_norm = std::sqrt(h->h_scal[S_GSQ]);
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  it.trust_region_radius = radius;
  it.iteration_time_s = it.cumulative_time_s = now_s() - t_start;
  sum->initial_cost = x_cost;
  int invalid = 0;
  int term = DCS_NO_CONVERGENCE;
  const char* msg = "Maximum number of iterations reached.";
  log_iter(it);
  dcs_iteration prev = it;
  // x always lives in h->xyt (accepted iterate); the candidate in h->cand_xyt; accepted -> pointer swap

  if (it.gradient_max_norm <= o.gradient_tolerance) { term = DCS_CONVERGENCE; msg = "Gradient tolerance reached."; }
  else
  for (;;) {
    {   // MaxSolverTimeReached; with several ranks rank 0's clock decides for everybody
      bool out_of_time = now_s() - t_start >= o.max_solver_time_s;
      if (h->world > 1 && o.max_solver_time_s < 1e6) {
        h->h_rank_scal[0] = (h->rank == 0 && out_of_time) ? 1.0 : 0.0;
        CK(cudaMemcpyAsync(h->rank_scal.p, h->h_rank_scal, sizeof(double), cudaMemcpyHostToDevice, st));
        CKS(allreduce_sum(h, h->rank_scal.p, 1));
        CK(cudaMemcpyAsync(h->h_rank_scal, h->rank_scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out_of_time = h->h_rank_scal[0] > 0.5;
      }
      if (out_of_time) { term = DCS_NO_CONVERGENCE; msg = "Maximum solver time reached."; break; }
    }
    if (prev.iteration >= o.max_num_iterations) { term = DCS_NO_CONVERGENCE; msg = "Maximum number of iterations reached."; break; }
    if (prev.gradient_max_norm <= o.gradient_tolerance) { term = DCS_CONVERGENCE; msg = "Gradient tolerance reached."; break; }
    if (radius <= o.min_trust_region_radius) { term = DCS_CONVERGENCE; msg = "Minimum trust region radius reached."; break; }
    const double it_start = now_s();
    std::memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1;

    if (!reuse_diagonal)
      LAUNCH(k_lm_diagonal, h->vec_grid(), 256, st, Hdiag_lm, h->scale.p, h->nrows, h->ldn, o.min_lm_diagonal, o.max_lm_diagonal, h->lmdiag.p);
    reuse_diagonal = true;
    int pcg_it = 0;
    double pcg_rel = 0;
    if (h->sc) {   // the switches' Schur complement depends on the radius: re-assemble the reduced system for this step
      CK(cudaEventRecord(h->ev0, st));
      CKS(linearize_sc(h, h->xyt.p, 1.0 / radius, 1));
      CK(cudaEventRecord(h->ev1, st));
      CK(cudaEventSynchronize(h->ev1));
      CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      h->eval_ms += ms;
    }
    CKS(pcg_solve(h, 1.0 / radius, nullptr, h->grad.p, &pcg_it, &pcg_rel));
    it.linear_solver_iterations = pcg_it;
    it.linear_solver_residual = pcg_rel;

    // model_cost_change = w.g - w.H.w / 2
    LAUNCH(k_pack_step, h->vec_grid(), kVecThreads, st, h->w.p, grad_full, 0, h->nrows, h->ldn, h->p4.p, h->partials.p,
           h->tickets.p + 4, h->scal.p);
    CKS(allreduce_sum(h, h->scal.p + S_WG, 1));
    CKS(ensure_mirror(h));
    CKS(spmv_product(h, h->Hdiag.p, S_WHW, 0));
    // true residual |(H + Lambda) w - g| / |g| of this step's linear solve, from q = H w just formed
    LAUNCH(k_true_residual, h->vec_grid(), kVecThreads, st, h->grad.p, h->q.p, h->w.p, h->lmdiag.p, h->scale.p, h->is_free.p,
           h->nrows, h->ldn, 1.0 / radius, h->partials.p, h->tickets.p + 4, h->scal.p);
    CKS(allreduce_sum(h, h->scal.p + S_TRES, 1));
    // candidate = x - w
    LAUNCH(k_apply_step, h->vec_grid(), kVecThreads, st, h->xyt.p, h->w.p, h->is_free.p, 0, h->nrows, h->ldn, h->cand_xyt.p,
           h->partials.p, h->tickets.p + 5, h->scal.p);
    CKS(allreduce_sum(h, h->scal.p + S_STEP_SQ, 2));
    CKS(halo_exchange(h, h->cand_xyt.p));
    if (h->sc) {   // switch steps from the pose step (back-substitution of the elimination) + their model-cost terms
      ScArgs A;
      sc_args(h, 1.0 / radius, 1, &A);
      LAUNCH(k_sc_edges<kScStep>, std::max(1, cdiv(h->E, kEdgeThreads)), kEdgeThreads, st, h->xyt.p, h->g2l.p, h->edgelist(), h->P, A,
             o.jacobi_scaling, h->sw.p, h->sw_scale.p, h->w.p, h->ldn, h->sw_cand.p, h->partials.p, h->tickets.p + 4, h->scal.p + S_SC);
    }
    CK(cudaEventRecord(h->ev0, st));
    CKS(cost_only(h, h->cand_xyt.p, S_CAND_COST, h->sw_cand.p));
    CK(cudaEventRecord(h->ev1, st));
    CKS(read_scalars(h));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->eval_ms += ms;
    double wg = h->h_scal[S_WG], whw = h->h_scal[S_WHW];
    double step_sq = h->h_scal[S_STEP_SQ], cand_xsq = h->h_scal[S_XSQ];
    if (h->sc) {   // model cost change over poses AND switches: w.g + ws.g_s - (w.H.w + ws.H_ss.ws + 2 ws.H_sx.w) / 2,
                   // with w.H.w = w.H_reduced.w + sum (H_sx.w)^2 / den
      const double* q = h->h_scal + S_SC;
      wg += q[0];
      whw += q[2] + q[1];
      step_sq += q[3];
      cand_xsq += q[4];
      h->h_scal[S_CAND_COST] = h->h_scal[S_SC_COST];
    }
    it.linear_solver_true_residual = h->h_scal[S_RR0] > 0.0 ? std::sqrt(h->h_scal[S_TRES] / h->h_scal[S_RR0]) : 0.0;
    const double model_cost_change = wg - 0.5 * whw;
    const bool finite_step = std::isfinite(wg) && std::isfinite(whw) && std::isfinite(step_sq);
    it.step_is_valid = finite_step && model_cost_change > 0.0;
    if (!it.step_is_valid) {
      if (++invalid >= o.max_num_consecutive_invalid_steps) {
        term = DCS_FAILURE; msg = "Number of consecutive invalid steps more than max_num_consecutive_invalid_steps."; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      it.cost = x_cost; it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
      it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      log_iter(it); prev = it; sum->num_unsuccessful_steps++;
      continue;
    }
    invalid = 0;
    double cand = h->h_scal[S_CAND_COST];
    if (!std::isfinite(cand)) cand = std::numeric_limits<double>::max();
    it.step_norm = std::sqrt(step_sq);
    it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
    if (it.step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      term = DCS_CONVERGENCE; msg = "Parameter tolerance reached.";
      it.cost = x_cost; it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      break;   // Ceres returns from Minimize() here: the terminating iteration is not appended to summary.iterations
    }
    it.cost_change = x_cost - cand;
    if (std::fabs(it.cost_change) <= o.function_tolerance * x_cost) {
      term = DCS_CONVERGENCE; msg = "Function tolerance reached.";
      it.cost = x_cost; it.trust_region_radius = radius;
      it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
      break;   // Ceres returns from Minimize() here: the terminating iteration is not appended to summary.iterations
    }
    it.relative_decrease = (cand >= std::numeric_limits<double>::max()) ? std::numeric_limits<double>::lowest()
                                                                          : (x_cost - cand) / model_cost_change;
    if (it.relative_decrease > o.min_relative_decrease) {
      std::swap(h->xyt.p, h->cand_xyt.p);
      if (h->sc) std::swap(h->sw.p, h->sw_cand.p);
      x_norm = std::sqrt(cand_xsq);
      CKS(timed_linearize(h->xyt.p));
      x_cost = h->h_scal[S_COST];
      it.step_is_successful = 1;
      it.cost = x_cost;
      it.gradient_max_norm = h->h_scal[S_GMAX];
      it.gradient_norm = std::sqrt(h->h_scal[S_GSQ]);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      sum->num_successful_steps++;
    } else {
      it.cost = cand;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum->num_unsuccessful_steps++;
    }
    it.trust_region_radius = radius;
    it.iteration_time_s = now_s() - it_start; it.cumulative_time_s = now_s() - t_start;
    log_iter(it);
    prev = it;
  }

  // monotonic steps: the accepted iterate in h->xyt is the best one
  CKS(download_poses(h, h->xyt.p, pose_xyt_inout));
  sum->final_cost = std::min(sum->initial_cost, min_logged_cost);
  sum->num_iterations = n_logged;
  sum->termination_type = term;
  sum->total_pcg_iterations = h->pcg_iters_total;
  sum->total_time_s = now_s() - t_start;
  sum->eval_time_s = h->eval_ms * 1e-3;
  sum->linear_solver_time_s = h->pcg_ms * 1e-3;
  std::snprintf(sum->message, sizeof(sum->message), "%s", msg);
  CK(cudaGetLastError());
  return DCS_OK;
}

// N3: many small independent solves at once.  One handle + one stream per problem, n_threads host threads pulling
// problems from a shared counter: the kernels of several tiny solves (each far too small to fill 148 SMs, each
// latency-bound on its own host round trips) overlap on the device.
int dcs_solve_batch(dcs_batch_item* items, int32_t n_items, const dcs_options* options, int32_t n_threads) {
  if (!items || n_items < 0 || !options) { g_err = "dcs_solve_batch: bad argument"; return DCS_ERR_ARG; }
  if (options->world > 1) { g_err = "dcs_solve_batch: single-rank only"; return DCS_ERR_ARG; }
  if (n_items == 0) return DCS_OK;
  const int nt = std::max(1, std::min<int>({n_threads > 0 ? n_threads : 8, n_items, 64}));
  std::atomic<int32_t> next{0};
  std::vector<std::string> errs((size_t)n_items);
  // keep released pool memory cached between items for the duration of the batch (default: returned to the driver at
  // every synchronisation); the previous threshold is restored and the pool trimmed when the batch is done
  cudaMemPool_t pool = nullptr;
  uint64_t old_threshold = 0;
  {
    int dev = 0;
    if (cudaSetDevice(options->device) == cudaSuccess && cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &old_threshold) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else pool = nullptr;
    (void)cudaGetLastError();
  }
  auto worker = [&]() {
    struct Mode { Mode() { tl_batch_worker = true; } ~Mode() { tl_batch_worker = false; tl_pinned.clear(); } } mode;
    for (;;) {
      const int32_t i = next.fetch_add(1);
      if (i >= n_items) return;
      dcs_batch_item& it = items[i];
      std::memset(&it.summary, 0, sizeof(it.summary));
      dcs_handle* h = nullptr;
      int rc = dcs_create(&it.graph, options, &h);
      if (rc == DCS_OK) {
        std::vector<double> tmp;
        double* x = it.pose_xyt_inout;
        if (!x) { tmp.assign(it.graph.pose_xyt, it.graph.pose_xyt + 3 * (size_t)it.graph.n_poses); x = tmp.data(); }
        rc = dcs_solve(h, x, &it.summary, nullptr, 0);
      }
      if (rc != DCS_OK) errs[(size_t)i] = g_err;      // g_err is thread-local: carry the text to the caller's thread
      if (h) dcs_destroy(h);
      it.status = rc;
    }
  };
  if (nt == 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  if (pool) {
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &old_threshold);
    cudaDeviceSynchronize();                 // every stream-ordered release of the batch has happened
    cudaMemPoolTrimTo(pool, 0);
    (void)cudaGetLastError();
  }
  for (int32_t i = 0; i < n_items; ++i)
    if (items[i].status != DCS_OK) { g_err = "dcs_solve_batch: item " + std::to_string(i) + ": " + errs[(size_t)i]; return items[i].status; }
  return DCS_OK;
}

int dcs_get_switches(dcs_handle* h, double* switches) {
  if (!h || !switches) return DCS_ERR_ARG;
  if (!h->sc) { g_err = "dcs_get_switches: the handle was not created with switchable_on"; return DCS_ERR_ARG; }
  CK(cudaSetDevice(h->dev));
  if (h->E > 0) CK(cudaMemcpy(switches, h->sw.p, (size_t)h->E * sizeof(double), cudaMemcpyDeviceToHost));
  return DCS_OK;
}

}  // extern "C"

#ifdef DCS_DEV_PROBES
#include "dcs_dev_probes.cuh"
#endif
