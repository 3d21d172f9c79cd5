// dcs_switchable.cuh — METHOD 2 (switchable constraints) on the device.
//
// Reference: DCS-ceres/src/ceres_error.cpp:203-297 (SwitchableClosureResidue <3,3,3,1>: e = s * e_plain, one scalar
// switch s per loop edge, HuberLoss(0.01) on that block), :300-317 (SwitchPriorResidue <1,1>: sqrt(lambda) (1 - s),
// no loss), DCS-ceres/main.cpp:105-150 (one switch + prior per closure / bogus edge, initial value 1, lambda = 1).
//
// A switch couples only to the two poses of its edge, so it is eliminated per edge, exactly, inside the linear solve
// (what a sparse Cholesky of the full system computes too).  With the plain Jacobians A0, B0 and e the plain error,
//   rho' = Huber'(s^2 |e|^2),  alpha = rho' s^2,  u_x = X0^T e,
//   H_xy = alpha X0^T Y0,  g_x = alpha u_x,  H_xs = rho' s u_x,  H_ss = rho' |e|^2 + lambda,  g_s = rho' s |e|^2 - lambda (1 - s).
// Ceres' Jacobi scale of the switch column sg = 1 / (1 + sqrt(H_ss at iteration 0)) and its LM diagonal
// d_s = clamp(sg^2 H_ss, min, max) give, in unscaled variables, den = H_ss + d_s / (radius sg^2) and the reduced system
//   H_xy - gamma u_x u_y^T,   g_x - (rho' s g_s / den) u_x,   gamma = (rho' s)^2 / den,
// i.e. X0^T (alpha I - gamma e e^T) Y0: the same rank-one form as the DCS terms (c1 = c2 = -gamma) with
// S_thth = alpha - gamma e_th^2 and one beta for all three components.  After the solve
//   delta_s = -(g_s - H_sx . w) / den      (w = -delta_x the pose step the PCG returns)
// d_s never has to be stored: a rejected step leaves x and s where they are, so clamp(sg^2 H_ss) recomputed at the
// current point IS the reused diagonal.
//
// k_linearize_sc is the row-owner kernel of METHOD 2 (same SELL walk, same compact block order and the same
// outputs as k_linearize, plain loop instead of the register pipeline: the linear solve dominates METHOD 2);
// k_sc_edges are the edge-order passes over the switches (scale, step recovery + model-cost terms, candidate cost).
#pragma once
#include "dcs_kernels.cuh"

namespace dcs {

struct ScArgs {
  double lambda;        // prior weight (main.cpp:110: 1.0)
  double inv_radius;    // 1 / trust-region radius of the step being solved
  double dmin, dmax;    // min / max LM diagonal
  int reduce;           // 0: pose blocks J_p^T J_p, J_p^T r as they are;  1: switches eliminated (Schur complement)
};

// per-switch quantities at the current point
struct ScEdge { double alpha, a_s, hss, gs, den, cost; };
__device__ __forceinline__ ScEdge sc_edge(const EdgeCore& C, double s, double sg, const ScArgs& A) {
  ScEdge K;
  K.alpha = C.rho1 * C.psi2;                 // psi2 = s^2
  K.a_s = C.rho1 * s;
  K.hss = fma(C.rho1, C.e2, A.lambda);
  const double om = 1.0 - s;
  K.gs = fma(K.a_s, C.e2, -A.lambda * om);
  const double ds = fmin(fmax(sg * sg * K.hss, A.dmin), A.dmax);
  K.den = K.hss + ds * A.inv_radius / (sg * sg);
  K.cost = fma(0.5 * A.lambda * om, om, C.cost);
  return K;
}

// normal-equation terms for S = [[alpha I + c1 exy exy^T, c2 eth exy],[., alpha_th]], v = (beta exy, beta eth)
__device__ __forceinline__ void edge_terms_general(const EdgeFrame& F, const EdgeCore& C, double alpha, double c1, double c2,
                                                   double alpha_th, double beta, EdgeTerms& T) {
  const double f0 = fma(C.q00, C.ex, -C.q01 * C.ey);
  const double f1 = fma(C.q01, C.ex, C.q00 * C.ey);
  const double et = fma(C.ex, C.epy, -C.ey * C.epx);
  const double tt = fma(C.epx, C.epx, C.epy * C.epy);
  const double cf0 = c1 * f0, cf1 = c1 * f1;
  T.U00 = fma(cf0, f0, alpha);
  T.U01 = cf0 * f1;
  T.U11 = fma(cf1, f1, alpha);
  const double k2 = c2 * C.t;
  T.sc0 = k2 * f0; T.sc1 = k2 * f1;
  const double k1 = c1 * et;
  T.e0 = fma(k1, f0, fma(alpha, F.dyw, -T.sc0));
  T.e1 = fma(k1, f1, fma(-alpha, F.dxw, -T.sc1));
  const double sts2 = k2 * et;
  T.alpha = alpha_th;
  T.k22 = fma(k1, et, fma(alpha, tt, alpha_th - 2.0 * sts2));
  T.o22 = sts2 - alpha_th;
  T.bf0 = beta * f0; T.bf1 = beta * f1;
  const double sae = beta * C.t;
  T.ga = fma(beta, et, -sae);
  T.gb = sae;
  T.cost = C.cost;
}

// sw_slot[slot] = (s, sg) of the slot's edge (loop edges only; anything for the rest)
__global__ void __launch_bounds__(kRowsPerBlock)
k_linearize_sc(const double4* __restrict__ xyt, RowLayout L, const HalfEdgeRec* __restrict__ recs, const double2* __restrict__ sw_slot,
               Params P, ScArgs A, int32_t n_loc, double* __restrict__ Hup, double* __restrict__ Hdiag, double* __restrict__ grad,
               double* __restrict__ task_part) {
  const L2Policy pol = make_l2_policy();
  const int task = blockIdx.x;
  if (task >= L.ntasks) return;
  const int lane = threadIdx.x & 31;
  const int lr = task * kSlice + lane;
  const uint4 info = L.rowinfo[lr];
  const int2 ti = L.task_info[task];
  PoseRec own;
  ld_pose(own, xyt + lr, pol.keep);
  const int deg = lr < L.nrows ? (int)info.x : 0;
  const int kmax = __reduce_max_sync(0xffffffffu, deg);
  const int64_t s0 = (int64_t)ti.x * kSlice + lane;
  int orun = ti.y;
  double d00 = 0, d01 = 0, d02 = 0, d11 = 0, d12 = 0, d22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0, gs_sq = 0, gs_max = 0;
  for (int k = 0; k < kmax; ++k) {
    const bool on = k < deg;
    HalfEdgeRec r;
    r.tmx = r.tmy = r.thm = 0.0; r.word = 0u; r.word_next = 0u;
    double2 sw = make_double2(1.0, 1.0);
    if (on) { ld_rec(r, recs + s0 + (int64_t)k * kSlice, pol.stream); sw = sw_slot[s0 + (int64_t)k * kSlice]; }
    const uint32_t word = r.word;
    const unsigned om = __ballot_sync(0xffffffffu, on && (word & kFlagOwner));
    if (on) {
      DCS_ASSERT((int32_t)(word & kIdxMask) < n_loc);
      PoseRec pc;
      ld_pose(pc, xyt + (word & kIdxMask), pol.keep);
      const EdgeFrame F = edge_frame(own.x, own.y, own.th, pc.x, pc.y, pc.th, word);
      const bool switched = (word & kFlagDcs) != 0;           // METHOD 2: the flag marks the edges that carry a switch
      EdgeTerms T;
      double ecost;
      if (switched) {
        const EdgeCore C = edge_core_t<true>(F, r.tmx, r.tmy, r.thm, false, sw.x, P);
        const ScEdge K = sc_edge(C, sw.x, sw.y, A);
        const double gamma = A.reduce ? K.a_s * K.a_s / K.den : 0.0;
        const double bcorr = A.reduce ? K.a_s * K.gs / K.den : 0.0;
        edge_terms_general(F, C, K.alpha, -gamma, -gamma, fma(-gamma * C.t, C.t, K.alpha), K.alpha - bcorr, T);
        ecost = K.cost;
        if (word & kFlagCost) { gs_sq = fma(K.gs, K.gs, gs_sq); gs_max = fmax(gs_max, fabs(K.gs)); }
      } else {
        const EdgeCore C = edge_core_t<false>(F, r.tmx, r.tmy, r.thm, false, 1.0, P);
        edge_terms(F, C, T);
        ecost = C.cost;
      }
      d00 += T.U00; d01 += T.U01; d11 += T.U11;
      if (word & kFlagSideB) { d02 += T.sc0; d12 += T.sc1; d22 += T.alpha; g0 += T.bf0; g1 += T.bf1; g2 += T.gb; }
      else                   { d02 -= T.e0;  d12 -= T.e1;  d22 += T.k22;   g0 -= T.bf0; g1 -= T.bf1; g2 += T.ga; }
      if (word & kFlagCost) cost += ecost;
      if (word & kFlagOwner) {
        const int64_t idx = (int64_t)orun + __popc(om & ((1u << lane) - 1u));
        DCS_ASSERT(idx >= 0 && idx < L.ldu);
        double* out = Hup + block_base(idx);
        out[0 * 32] = -T.U00; out[1 * 32] = -T.U01; out[2 * 32] = -T.sc0;
        out[3 * 32] = -T.U11; out[4 * 32] = -T.sc1;
        out[5 * 32] = T.e0;   out[6 * 32] = T.e1;   out[7 * 32] = T.o22;
      }
    }
    orun += __popc(om);
  }
  if (lr < L.nrows) {
    Hdiag[0 * L.ldn + lr] = d00; Hdiag[1 * L.ldn + lr] = d01; Hdiag[2 * L.ldn + lr] = d02;
    Hdiag[3 * L.ldn + lr] = d11; Hdiag[4 * L.ldn + lr] = d12; Hdiag[5 * L.ldn + lr] = d22;
    grad[0 * L.ldn + lr] = g0; grad[1 * L.ldn + lr] = g1; grad[2 * L.ldn + lr] = g2;
  }
  // the gradient norms of METHOD 2 run over poses AND switches (every switch is a parameter block)
  double r0 = cost, r1 = fma(g0, g0, fma(g1, g1, fma(g2, g2, gs_sq))), r2 = fmax(fmax(fabs(g0), gs_max), fmax(fabs(g1), fabs(g2)));
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

// (s, sg) of every switched edge -> the (up to two) slots of the edge
__global__ void k_sc_to_slots(const int32_t* __restrict__ edge_slot, const uint8_t* __restrict__ switched, int32_t E,
                              const double* __restrict__ sw, const double* __restrict__ sw_scale, double2* sw_slot) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || !switched[e]) return;
  const double2 v = make_double2(sw[e], sw_scale[e]);
  const int32_t sa = edge_slot[2 * (int64_t)e], sb = edge_slot[2 * (int64_t)e + 1];
  if (sa >= 0) sw_slot[sa] = v;
  if (sb >= 0) sw_slot[sb] = v;
}

// Edge-order passes (one thread per edge; both poses gathered through the global -> stored-position map):
//   kScScale: sw_scale = 1 / (1 + sqrt(H_ss))  (Ceres' Jacobi scaling of the switch columns, iteration 0)
//   kScStep : ws = (g_s - H_sx . w) / den = -delta_s; sw_cand = s - ws; sums for the model cost change and the norms:
//             out[0] = sum ws g_s, out[1] = sum (H_ss ws^2 + 2 ws H_sx.w), out[2] = sum (H_sx.w)^2 / den,
//             out[3] = sum ws^2, out[4] = sum sw_cand^2
//   kScCost : cost at (xyt, sw) over ALL edges (odometry edges: the plain functor, no switch) -> out[0];
//             out[1] = sum s^2 over the switches
enum { kScScale = 0, kScStep = 1, kScCost = 2 };
template <int kMode>
__global__ void __launch_bounds__(kEdgeThreads)
k_sc_edges(const double4* __restrict__ xyt, const int32_t* __restrict__ g2l, EdgeList E, Params P, ScArgs A, int jacobi,
           const double* __restrict__ sw, double* sw_scale, const double* __restrict__ w, int64_t ldn, double* sw_cand,
           double* partials, unsigned int* ticket, double* out) {
  const int64_t e = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  constexpr int K = kMode == kScStep ? 5 : 2;
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.0;
  if (e < E.n) {
    const bool switched = E.dcs[e] != 0;
    if (switched || kMode == kScCost) {
      const int32_t la = g2l[E.a[e]], lb = g2l[E.b[e]];
      const double4 pa = xyt[la], pb = xyt[lb];
      const EdgeFrame F = edge_frame(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, 0u);
      const double s = switched ? sw[e] : 1.0;
      const EdgeCore C = edge_core_t<true>(F, E.tmx[e], E.tmy[e], E.thm[e], false, s, P);
      if (kMode == kScCost) {
        const double om = 1.0 - s;
        acc[0] = switched ? fma(0.5 * A.lambda * om, om, C.cost) : C.cost;
        acc[1] = switched ? s * s : 0.0;
      } else if (kMode == kScScale) {
        sw_scale[e] = jacobi ? 1.0 / (1.0 + sqrt(fma(C.rho1, C.e2, A.lambda))) : 1.0;
      } else {
        const ScEdge Kk = sc_edge(C, s, sw_scale[e], A);
        const double f0 = fma(C.q00, C.ex, -C.q01 * C.ey), f1 = fma(C.q01, C.ex, C.q00 * C.ey);
        const double et = fma(C.ex, C.epy, -C.ey * C.epx);
        // u_a = (-f, et - t), u_b = (f, t);  w is zero on rows that are not parameters
        const double wa0 = w[la], wa1 = w[ldn + la], wa2 = w[2 * ldn + la];
        const double wb0 = w[lb], wb1 = w[ldn + lb], wb2 = w[2 * ldn + lb];
        const double uw = fma(f0, wb0 - wa0, fma(f1, wb1 - wa1, fma(et - C.t, wa2, C.t * wb2)));
        const double hw = Kk.a_s * uw;
        const double ws = (Kk.gs - hw) / Kk.den;
        const double sc = s - ws;
        sw_cand[e] = sc;
        acc[0] = ws * Kk.gs;
        acc[1] = fma(Kk.hss * ws, ws, 2.0 * ws * hw);
        acc[2] = hw * hw / Kk.den;
        acc[3] = ws * ws;
        acc[4] = sc * sc;
      }
    }
  }
  if (kMode != kScScale) grid_reduce_sum<K, kEdgeThreads>(acc, partials, ticket, out);
}

}  // namespace dcs
