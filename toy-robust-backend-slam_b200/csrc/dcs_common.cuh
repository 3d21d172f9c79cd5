// dcs_common.cuh — shared device code of libdcs_b200: error handling, the per-edge
// SE(2)+DCS+Huber linearisation (analytic form of what Ceres autodiff produces for the
// reference functors) and deterministic reduction helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcs {

// ------------------------------------------------------------------------------------------
// layout constants
// ------------------------------------------------------------------------------------------
constexpr int kRowsPerBlock = 32;        // threads per CTA in the row-owner kernels: ONE warp task per CTA, so a
                                         // finished task frees its registers at once (no sibling warp to wait for)
constexpr int kWindow = 1024;            // rows per sorting window (rows ranked by degree inside it)
constexpr int kSlice = 32;               // rows per warp task = lanes of a tile
constexpr int kSlicesPerWindow = kWindow / kSlice;
constexpr int kK1Rounds = 2;             // rounds per pipeline stage of k_linearize; a half-edge record carries the index
                                         // word of the round kK1Rounds later (baked into the data by the pattern build)
constexpr int kTailTiles = 4;            // zero tiles after the last task (prefetch addresses are clamped, this is slack)
constexpr uint32_t kKeyNonOwner = 1u << 27;  // sort-key bit just above the 27 column bits: non-owner half-edges follow a
                                             // row's owner ones (the radix sort covers bits 0..31 of the column word)
constexpr uint32_t kKeyHalo = 1u << 28;      // multi-rank: half-edges whose column lives on another rank sort LAST in their
                                             // row, so the SpMV can do the local columns while the halo is still in flight
constexpr uint32_t kIdxMask = 0x07FFFFFFu;  // low 27 bits of a half-edge word: other pose (<= 134M poses)
constexpr uint32_t kFlagSideB = 1u << 31;   // row pose is the edge's second endpoint (Edge::b)
constexpr uint32_t kFlagDcs = 1u << 30;     // DCS functor applies (loop/bogus edge and METHOD 1)
constexpr uint32_t kFlagOtherFixed = 1u << 29;  // other endpoint is constant: no off-diagonal block
constexpr uint32_t kFlagCost = 1u << 28;    // this half-edge accounts for the edge's cost
constexpr uint32_t kFlagOwner = 1u << 27;   // this half-edge writes the edge's off-diagonal block (upper triangle, or a
                                            // block whose mirror lives on another rank)

struct Params {
  double phi;       // DCS upper bound (0.5)
  double hub_a;     // Huber delta
  double hub_b;     // delta^2
};

// ------------------------------------------------------------------------------------------
// Per-edge linearisation.
//
// Reference: DCS-ceres/src/ceres_error.cpp:42-94 (OdometryResidue) and :135-196
// (DCSClosureResidue), evaluated by AutoDiffCostFunction<.,3,3,3> (:34, :127), then
// ceres::HuberLoss(0.01) + Corrector (DCS-ceres/main.cpp:67-68).  Closed form used here:
//   diff = M^-1 (A^-1 B)  =>  (ex,ey) = Q (tb - ta) - Rm^T tm,   Q = Rm^T Ra^T = R(-(tha+thm))
//   e_th = asin(sin d), d = thb - tha - thm  (fold of d into [-pi/2, pi/2]),
//   d e_th / d d = cos d / sqrt(1 - sin^2 d) = sign(cos d)
//   DCS: psi = min(1, sqrt(2 phi / (phi + ex^2 + ey^2))), r = psi e,
//        J = psi J_e + e (grad psi)^T, grad psi = -psi/(phi+res) (ex grad ex + ey grad ey) iff psi<1
//   Huber: s = |r|^2 > delta^2 -> rho' = delta / sqrt(s); r, J scaled by sqrt(rho'); cost = rho/2.
// Measurement is pre-rotated once at upload: (tmx,tmy) = Rm^T (dx,dy); cos/sin(tha+thm) come from one sincos.
// ------------------------------------------------------------------------------------------
struct EdgeLin {
  double r0, r1, r2;                                // corrected residual
  double a00, a01, a02, a10, a11, a12, a20, a21, a22;  // d r / d (xa, ya, tha)
  double b00, b01, b10, b11, b20, b21, b22;         // d r / d (xb, yb, thb); b02 = b12 = 0
  double cost, psi, rho1;
};

__device__ __forceinline__ double fold_angle(double d, double* sigma) {
  // asin(sin d) and sign(cos d), without transcendental calls.
  const double inv2pi = 0.15915494309189533577;
  const double twopi_hi = 6.283185307179586232;     // double(2*pi)
  const double twopi_lo = 2.4492935982947064e-16;   // 2*pi - twopi_hi
  const double pi = 3.141592653589793116;
  const double k = rint(d * inv2pi);
  double t = fma(-k, twopi_hi, d);
  t = fma(-k, twopi_lo, t);            // t in [-pi, pi]
  const double half_pi = 1.5707963267948966;
  const bool hi = t > half_pi, lo = t < -half_pi;
  *sigma = (hi || lo) ? -1.0 : 1.0;
  return hi ? (pi - t) : (lo ? (-pi - t) : t);
}

template <bool kNeedJac>
__device__ __forceinline__ void edge_linearize(double xa, double ya, double tha,
                                               double xb, double yb, double thb,
                                               double tmx, double tmy, double thm,
                                               bool dcs, const Params& P, EdgeLin& L) {
  double q00, q01;                             // cos / sin (tha + thm): Q = Rm^T Ra^T = R(-(tha + thm))
  sincos(tha + thm, &q01, &q00);
  const double dxw = xb - xa, dyw = yb - ya;
  const double epx = fma(q00, dxw, q01 * dyw);    // Q d
  const double epy = fma(q00, dyw, -q01 * dxw);
  const double ex = epx - tmx, ey = epy - tmy;
  double sigma;
  const double eth = fold_angle(thb - tha - thm, &sigma);

  double psi = 1.0, kap = 0.0;
  if (dcs) {
    const double res = fma(ex, ex, ey * ey);
    if (res > P.phi) {                       // <=> psi_org < 1
      const double den = P.phi + res;
      psi = sqrt(2.0 * P.phi / den);
      kap = -psi / den;
    }
  }
  double r0 = psi * ex, r1 = psi * ey, r2 = psi * eth;
  const double s = fma(r0, r0, fma(r1, r1, r2 * r2));
  double w = 1.0, rho1 = 1.0, cost = 0.5 * s;
  if (s > P.hub_b) {
    const double rs = sqrt(s);
    rho1 = fmax(2.2250738585072014e-308, P.hub_a / rs);
    cost = 0.5 * (2.0 * P.hub_a * rs - P.hub_b);
    w = sqrt(rho1);
  }
  L.cost = cost; L.psi = psi; L.rho1 = rho1;
  L.r0 = w * r0; L.r1 = w * r1; L.r2 = w * r2;
  if (!kNeedJac) return;
  // plain Jacobian rows: d ex = (-q00, -q01, epy | q00, q01, 0), d ey = (q01, -q00, -epx | -q01, q00, 0),
  // d eth = (0, 0, -sigma | 0, 0, sigma)
  const double wp = w * psi;
  // grad psi (zero unless DCS active): kap * (ex d ex + ey d ey)
  const double m0 = kap * fma(-q00, ex, q01 * ey);
  const double m1 = kap * fma(-q01, ex, -q00 * ey);
  const double m2 = kap * fma(ex, epy, -ey * epx);
  // w * e (unscaled by psi) for the rank-one term
  const double we0 = w * ex, we1 = w * ey, we2 = w * eth;
  L.a00 = fma(we0, m0, -wp * q00); L.a01 = fma(we0, m1, -wp * q01); L.a02 = fma(we0, m2, wp * epy);
  L.a10 = fma(we1, m0, wp * q01);  L.a11 = fma(we1, m1, -wp * q00); L.a12 = fma(we1, m2, -wp * epx);
  L.a20 = we2 * m0;                L.a21 = we2 * m1;                L.a22 = fma(we2, m2, -wp * sigma);
  L.b00 = fma(-we0, m0, wp * q00); L.b01 = fma(-we0, m1, wp * q01);
  L.b10 = fma(-we1, m0, -wp * q01); L.b11 = fma(-we1, m1, wp * q00);
  L.b20 = -we2 * m0;               L.b21 = -we2 * m1;               L.b22 = wp * sigma;
}

// ------------------------------------------------------------------------------------------
// Normal-equation terms of one edge WITHOUT forming the Jacobian (hot path, k_linearize).
//
// With A0 = [[-Q, t],[0, -sigma]], B0 = [[Q, 0],[0, sigma]] the plain Jacobians (Q = Rm^T Ra^T,
// t = perp(Q d)), the corrected Jacobians are J_x = w C X0 with C = psi I + kappa e ebar^T
// (ebar = (ex,ey,0), kappa = -psi/(phi+res) iff DCS active) and w^2 = rho' (Huber).  Hence
//   J_x^T J_y = X0^T S Y0,  S = rho' C^T C = [[alpha I + c1 exy exy^T, c2 eth exy],[., alpha]],
//   J_x^T r   = X0^T v,     v = (beta exy, alpha eth),
//   alpha = rho' psi^2, c2 = -alpha/(phi+res), c1 = 2 c2 + alpha |e|^2/(phi+res)^2, beta = alpha + c2 |e|^2.
// Rotating by Q (f = Q^T exy, Q^T t = perp(d) = n) every block is a combination of
//   U = alpha I + c1 f f^T,  c = c2 eth f,  dv = alpha n + c1 (exy.t) f,  and a few scalars:
//   H_aa = [[U, sigma c - dv],[., alpha(|t|^2+1) + c1 et^2 - 2 sigma ts2]]   g_a = (-beta f, beta et - sigma alpha eth)
//   H_bb = [[U, sigma c],[., alpha]]                                         g_b = ( beta f, sigma alpha eth)
//   H_ab = [[-U, -sigma c],[(dv - sigma c)^T, sigma ts2 - alpha]],  H_ba = H_ab^T,      ts2 = c2 eth et.
// Only one reciprocal (DCS active) and one reciprocal square root (Huber active) per edge; psi and
// sqrt(rho') themselves are never needed.  ~95 fp64 instructions instead of ~330 for the explicit J^T J.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {      // x normal, > 0
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {    // x normal, > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y;
  double e = fma(-x * y, y, 1.0);
  y = fma(h, e, y);
  h = 0.5 * y;
  e = fma(-x * y, y, 1.0);
  return fma(h, e, y);
}
// ---- transcendental-free helpers of the hot path -------------------------------------------------------------------
// Numeric constants live in __constant__ memory: fp64 instructions take a constant-bank operand directly, whereas an
// immediate costs two UMOVs per use inside the loop (37 per round in the round-1 kernel).
struct MathConsts {
  double magic;            // 1.5 * 2^52: round-to-nearest-integer by addition
  double inv_pi, pi_hi, pi_lo;
  double two_over_pi, pio2_hi, pio2_lo;
  double s1, s2, s3, s4, s5, s6;   // fdlibm __kernel_sin
  double c1, c2, c3, c4, c5, c6;   // fdlibm __kernel_cos
};
__constant__ MathConsts kMath = {
    6755399441055744.0,
    0.31830988618379067154, 3.141592653589793116, 1.2246467991473532e-16,
    0.63661977236758134308, 1.5707963267948966, 6.123233995736766e-17,
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};

// t = d - rint(d / pi) * pi  in [-pi/2, pi/2].  The reference's angle row is asin(sin d) = sigma t with
// sigma = sign(cos d) = (-1)^rint(d/pi) (also its derivative); the normal-equation terms only ever need
// sigma * e_th = t and e_th^2 = t^2, so neither sigma nor the signed value is formed.
__device__ __forceinline__ double fold_pi(double d) {
  const double k = fma(d, kMath.inv_pi, kMath.magic) - kMath.magic;
  return fma(-k, kMath.pi_lo, fma(-k, kMath.pi_hi, d));
}

// sin / cos of a pose angle (|x| < ~1e6): two-term Cody-Waite reduction by pi/2 (the products are exact inside the
// fma, absolute error ~1e-16) and the fdlibm kernel polynomials (< 1 ulp on [-pi/4, pi/4]).  Leaner than the inlined
// sincos(): no Payne-Hanek slow path, no special-value handling, coefficients as constant-bank operands.
__device__ __forceinline__ void sincos_cw(double x, double& sn, double& cs) {
  const double kd = fma(x, kMath.two_over_pi, kMath.magic);
  const int q = __double2loint(kd);
  const double k = kd - kMath.magic;
  const double r = fma(-k, kMath.pio2_lo, fma(-k, kMath.pio2_hi, x));
  const double z = r * r;
  double ps = fma(z, kMath.s6, kMath.s5);
  ps = fma(z, ps, kMath.s4); ps = fma(z, ps, kMath.s3); ps = fma(z, ps, kMath.s2); ps = fma(z, ps, kMath.s1);
  const double s = fma(r * z, ps, r);
  double pc = fma(z, kMath.c6, kMath.c5);
  pc = fma(z, pc, kMath.c4); pc = fma(z, pc, kMath.c3); pc = fma(z, pc, kMath.c2); pc = fma(z, pc, kMath.c1);
  const double c = fma(z, fma(z, pc, -0.5), 1.0);
  const bool swap = (q & 1) != 0;
  const double ss = swap ? c : s, cc = swap ? s : c;
  // quadrant signs by flipping the sign bit: sin negative for q = 2, 3; cos negative for q = 1, 2
  sn = __hiloint2double(__double2hiint(ss) ^ ((q & 2) << 30), __double2loint(ss));
  cs = __hiloint2double(__double2hiint(cc) ^ (((q + 1) & 2) << 30), __double2loint(cc));
}

struct EdgeTerms {
  double U00, U01, U11;   // alpha I + c1 f f^T
  double sc0, sc1;        // sigma * c
  double e0, e1;          // dv - sigma c
  double alpha;
  double k22;             // alpha (|t|^2 + 1) + c1 et^2 - 2 sigma ts2
  double o22;             // sigma ts2 - alpha
  double bf0, bf1;        // beta f
  double ga;              // beta et - sigma alpha eth        (gradient, theta of pose a)
  double gb;              // sigma alpha eth                  (gradient, theta of pose b)
  double cost;
};

// Row-owner frame of a half-edge.  The row pose o and the other pose p give the edge frame (a = first endpoint,
// b = second) without selecting six doubles: d = tb - ta = +-(p - o) is a sign-bit flip with the word's side bit
// (bit 31 = kFlagSideB = the IEEE sign position), only the angle of pose a needs a select.
struct EdgeFrame { double dxw, dyw, sdth, tha; };
__device__ __forceinline__ double flip_sign(double v, uint32_t sign_mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)sign_mask, __double2loint(v));
}
__device__ __forceinline__ EdgeFrame edge_frame(double ox, double oy, double oth, double px, double py, double pth, uint32_t word) {
  const uint32_t sgn = word & kFlagSideB;
  EdgeFrame F;
  F.dxw = flip_sign(px - ox, sgn);
  F.dyw = flip_sign(py - oy, sgn);
  F.sdth = flip_sign(pth - oth, sgn);         // thb - tha
  F.tha = sgn ? pth : oth;
  return F;
}

// cost of one edge and the intermediates the normal-equation terms reuse
struct EdgeCore { double q00, q01, epx, epy, ex, ey, t, psi2, inv_den, e2, rho1, cost; };
// kSwitch (METHOD 2, dcs_switchable.cuh): the error is scaled by the edge's switch value sw instead of the DCS weight
template <bool kSwitch>
__device__ __forceinline__ EdgeCore edge_core_t(const EdgeFrame& F, double tmx, double tmy, double thm, bool dcs, double sw, const Params& P) {
  EdgeCore C;
  sincos_cw(F.tha + thm, C.q01, C.q00);     // Q = Rm^T Ra^T = R(-(tha + thm)); one sincos per half-edge instead of
                                            // streaming cos/sin of the measurement and gathering cos/sin of the pose
  C.epx = fma(C.q00, F.dxw, C.q01 * F.dyw);
  C.epy = fma(C.q00, F.dyw, -C.q01 * F.dxw);
  C.ex = C.epx - tmx; C.ey = C.epy - tmy;
  C.t = fold_pi(F.sdth - thm);              // sigma * e_th
  const double res = fma(C.ex, C.ex, C.ey * C.ey);
  C.e2 = fma(C.t, C.t, res);
  // DCS (psi_org < 1 <=> res > phi) and Huber, branch-free: nearly every warp has a lane on either side of both
  // thresholds, so the reciprocal and the reciprocal square root are always computed and selected afterwards
  if (kSwitch) {
    C.inv_den = 0.0;
    C.psi2 = sw * sw;
  } else {
    const bool active = dcs && res > P.phi;
    const double inv = fast_rcp(P.phi + res);
    C.inv_den = active ? inv : 0.0;
    C.psi2 = active ? 2.0 * P.phi * inv : 1.0;
  }
  const double s = C.psi2 * C.e2;
  const bool lin = s > P.hub_b;
  const double rs = fast_rsqrt(lin ? s : 1.0);
  C.rho1 = lin ? P.hub_a * rs : 1.0;          // >= DBL_MIN for every finite s
  C.cost = lin ? fma(P.hub_a, s * rs, -0.5 * P.hub_b) : 0.5 * s;
  return C;
}
__device__ __forceinline__ EdgeCore edge_core(const EdgeFrame& F, double tmx, double tmy, double thm, bool dcs, const Params& P) {
  return edge_core_t<false>(F, tmx, tmy, thm, dcs, 1.0, P);
}

__device__ __forceinline__ void edge_terms(const EdgeFrame& F, const EdgeCore& C, EdgeTerms& T) {
  const double alpha = C.rho1 * C.psi2;
  const double c2 = -alpha * C.inv_den;                       // 0 unless DCS active
  const double c1 = fma(alpha * C.e2, C.inv_den * C.inv_den, 2.0 * c2);
  const double beta = fma(c2, C.e2, alpha);
  const double f0 = fma(C.q00, C.ex, -C.q01 * C.ey);          // Q^T exy
  const double f1 = fma(C.q01, C.ex, C.q00 * C.ey);
  const double et = fma(C.ex, C.epy, -C.ey * C.epx);          // exy . t
  const double tt = fma(C.epx, C.epx, C.epy * C.epy);         // |t|^2
  const double cf0 = c1 * f0, cf1 = c1 * f1;
  T.U00 = fma(cf0, f0, alpha);
  T.U01 = cf0 * f1;
  T.U11 = fma(cf1, f1, alpha);
  const double k2 = c2 * C.t;                                 // sigma c2 eth
  T.sc0 = k2 * f0; T.sc1 = k2 * f1;
  const double k1 = c1 * et;
  T.e0 = fma(k1, f0, fma(alpha, F.dyw, -T.sc0));              // dv - sigma c, dv = alpha n + k1 f, n = (dyw, -dxw)
  T.e1 = fma(k1, f1, fma(-alpha, F.dxw, -T.sc1));
  const double sts2 = k2 * et;                                // sigma ts2
  T.alpha = alpha;
  T.k22 = fma(k1, et, fma(alpha, tt + 1.0, -2.0 * sts2));
  T.o22 = sts2 - alpha;
  T.bf0 = beta * f0; T.bf1 = beta * f1;
  const double sae = alpha * C.t;                             // sigma alpha eth
  T.ga = fma(beta, et, -sae);
  T.gb = sae;
  T.cost = C.cost;
}

// Own rows are STORED in (window, rank) order - the order the row-owner kernels walk them in - so that every
// per-row array (diagonal blocks, gradient, PCG vectors, own poses) is read and written coalesced by K1 / SpMV.
// Natural local row r <-> storage position row_pos(rank_of, r); perm is the inverse inside a window.  Real rows
// occupy the storage positions [0, nrows) (padding rows have degree 0 and rank last), so elementwise kernels
// are order-agnostic.
__device__ __forceinline__ int32_t row_pos(const uint16_t* __restrict__ rank_of, int32_t r) {
  return (r & ~(kWindow - 1)) + (int32_t)rank_of[r];
}

// ------------------------------------------------------------------------------------------
// cache-hinted accesses: matrix / half-edge streams are read once per pass (keep them out of
// L1, first to leave L2); gathered vectors stay cached.
// ------------------------------------------------------------------------------------------
struct L2Policy { uint64_t stream, keep; };
__device__ __forceinline__ L2Policy make_l2_policy() {
  L2Policy p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  return p;
}
__device__ __forceinline__ double ld_stream(const double* p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_stream(double* p, double v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
// bulk L2 prefetch of a contiguous range (address and size multiples of 16 bytes); asynchronous, no destination
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// gathered operands (poses, the PCG direction vector): small, reused by every row that references them
__device__ __forceinline__ double4 ld_keep4(const double4* p, uint64_t pol) {
  double4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2+16], %3;" : "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
  return v;
}
// Predicated forms (one @p instruction each, no branch around the load: a branch would make ptxas re-arm the
// consumer's scoreboard wait on the path that skipped it).  The destination keeps its value when `on` is false.
__device__ __forceinline__ void ld_stream_if(double& v, const double* p, uint64_t pol, bool on) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;\n\t}"
               : "+d"(v) : "l"(p), "l"(pol), "r"((int)on));
}
__device__ __forceinline__ void ld_stream_u32_if(uint32_t& v, const uint32_t* p, uint64_t pol, bool on) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;\n\t}"
               : "+r"(v) : "l"(p), "l"(pol), "r"((int)on));
}
__device__ __forceinline__ void ld_keep3_if(double& x, double& y, double& z, const double4* p, uint64_t pol, bool on) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%3], %4;\n\t"
               "@q ld.global.nc.L2::cache_hint.f64 %2, [%3+16], %4;\n\t}"
               : "+d"(x), "+d"(y), "+d"(z) : "l"(p), "l"(pol), "r"((int)on));
}
__device__ __forceinline__ double2 ld_keep2(const double2* p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
// plain (no policy) variant for the edge-order kernels
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------
// Deterministic grid reduction: fixed-shape block tree, per-block partials, the last block to
// arrive (ticket) folds the partials in index order.  K values at once.
// `partials` holds K * gridDim.x doubles, `out` K doubles, `ticket` one zero-initialised uint.
// ------------------------------------------------------------------------------------------
template <int K, int kThreads>
__device__ __forceinline__ void grid_reduce_sum(double (&v)[K], double* partials, unsigned int* ticket, double* out) {
  __shared__ double s_red[K][kThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[k][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) x += s_red[k][w];
      partials[(size_t)k * gridDim.x + blockIdx.x] = x;
    }
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block: fold partials in a fixed order (strided per thread, then fixed tree)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) x += __ldcg(&partials[(size_t)k * gridDim.x + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) s_red[k][wid] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += s_red[k][w];
      out[k] = t;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;  // re-arm for the next launch
}

// max-reduction variant (single value) used for |g|_inf
template <int kThreads>
__device__ __forceinline__ void grid_reduce_max(double v, double* partials, unsigned int* ticket, double* out) {
  __shared__ double s_red[kThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double x = v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  if (lane == 0) s_red[wid] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) m = fmax(m, s_red[w]);
    partials[blockIdx.x] = m;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double m = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) m = fmax(m, __ldcg(&partials[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t = fmax(t, s_red[w]);
    *out = t;
    *ticket = 0u;
  }
}

}  // namespace dcs
