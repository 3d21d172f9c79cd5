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
constexpr int kRowsPerBlock = 128;       // rows (poses) per CTA in the row-owner kernels
constexpr uint32_t kIdxMask = 0x0FFFFFFFu;  // low 28 bits of a half-edge word: other pose
constexpr uint32_t kFlagSideB = 1u << 31;   // row pose is the edge's second endpoint (Edge::b)
constexpr uint32_t kFlagDcs = 1u << 30;     // DCS functor applies (loop/bogus edge and METHOD 1)
constexpr uint32_t kFlagOtherFixed = 1u << 29;  // other endpoint is constant: no off-diagonal block
constexpr uint32_t kFlagCost = 1u << 28;    // this half-edge accounts for the edge's cost

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
// Measurement is pre-rotated once at upload: (tmx,tmy) = Rm^T (dx,dy).
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
__device__ __forceinline__ void edge_linearize(double xa, double ya, double tha, double ca, double sa,
                                               double xb, double yb, double thb,
                                               double tmx, double tmy, double thm, double cm, double sm,
                                               bool dcs, const Params& P, EdgeLin& L) {
  const double q00 = fma(cm, ca, -sm * sa);   // cos(tha + thm)
  const double q01 = fma(cm, sa, sm * ca);    // sin(tha + thm)
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
// cache-hinted accesses: matrix / half-edge streams are read once per pass (keep them out of
// L1, first to leave L2); gathered vectors stay cached.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
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
