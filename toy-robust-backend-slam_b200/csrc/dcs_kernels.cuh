// dcs_kernels.cuh — the hot kernels of the DCS-LM path (sm_100a, fp64, HBM/fp64-pipe bound; no
// dense contraction anywhere, so no tensor cores).
//
//   k_linearize   K1+K2  fused residual + analytic Jacobian + DCS + Huber + J^T J / J^T r
//                        assembly, row-owner form: one thread per pose walks the pose's
//                        half-edges (jagged-diagonal layout -> coalesced), keeps the diagonal
//                        block and gradient in registers and streams one 3x3 off-diagonal block
//                        per half-edge.  No atomics, bit-reproducible.
//   k_cost        K6     cost-only evaluation at a candidate point (thread per edge)
//   k_edge_eval          per-edge r / J / psi / rho' dump in edge order (parity hook; same math)
//   k_spmv        K3     q = (H + Lambda) p over the same layout, fused p.q partial reduction
//   k_pcg_*       K4     fused PCG vector updates + dot products, scalars stay on the device
//   k_precond     K5     A_ii = H_ii + Lambda_i, M_i^-1 = A_ii^-1 (symmetric 3x3)
#pragma once
#include "dcs_common.cuh"

namespace dcs {

// scalar slots in the device scalar block
enum { S_COST = 0, S_GSQ = 1, S_GMAX = 2, S_PQ = 3, S_RZ = 4, S_RZ_NEXT = 5, S_RR = 6, S_RR0 = 7,
       S_WG = 8, S_WHW = 9, S_STEP_SQ = 10, S_XSQ = 11, S_CAND_COST = 12, S_TMP = 13, S_COUNT = 16 };

struct RowLayout {           // jagged-diagonal layout of the owned rows
  int32_t row_lo;            // first owned pose
  int32_t nrows;             // owned poses
  int64_t ldn;               // leading dimension of per-row SoA arrays (>= gridDim * kRowsPerBlock)
  int64_t ldh;               // leading dimension of per-half-edge SoA arrays
  const int32_t* row_ptr;    // [nrows+1] sorted-CSR offsets (gives the degree)
  const uint8_t* perm;       // [nblk*128] rank -> local row within the CTA
  const int32_t* rp_off;     // [nblk]   offset of the CTA's rounds in round_ptr
  const int32_t* round_ptr;  // first slot of every round
};

struct HalfEdges {           // per-half-edge SoA, in JDS slot order
  const uint32_t* other;     // other pose | flags
  const double* tmx;         // Rm^T (dx,dy)
  const double* tmy;
  const double* thm;
  const double* cm;
  const double* sm;
};

// ------------------------------------------------------------------------------------------------
// K1 + K2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock)
k_linearize(const double4* __restrict__ xyt, const double2* __restrict__ cs, RowLayout L, HalfEdges H, Params P,
            double* __restrict__ Hoff, double* __restrict__ Hdiag, double* __restrict__ grad,
            double* partials, unsigned int* tickets, double* scal) {
  const int t = threadIdx.x;
  const int64_t slot0 = (int64_t)blockIdx.x * kRowsPerBlock;
  const int lr = (int)slot0 + L.perm[slot0 + t];
  int deg = 0;
  double ox = 0, oy = 0, oth = 0, oc = 1, os = 0;
  if (lr < L.nrows) {
    deg = L.row_ptr[lr + 1] - L.row_ptr[lr];
    const double4 p = xyt[L.row_lo + lr];
    const double2 q = cs[L.row_lo + lr];
    ox = p.x; oy = p.y; oth = p.z; oc = q.x; os = q.y;
  }
  const int32_t* rp = L.round_ptr + L.rp_off[blockIdx.x];
  double d00 = 0, d01 = 0, d02 = 0, d11 = 0, d12 = 0, d22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
  for (int k = 0; k < deg; ++k) {
    const int64_t idx = (int64_t)rp[k] + t;
    const uint32_t word = ld_stream_u32(H.other + idx);
    const double tmx = ld_stream(H.tmx + idx), tmy = ld_stream(H.tmy + idx), thm = ld_stream(H.thm + idx);
    const double cm = ld_stream(H.cm + idx), sm = ld_stream(H.sm + idx);
    const uint32_t j = word & kIdxMask;
    const bool side_b = (word & kFlagSideB) != 0;
    const double4 pj = xyt[j];
    double2 qj = make_double2(1.0, 0.0);
    if (side_b) qj = cs[j];
    // edge frame: a = first endpoint, b = second
    const double xa = side_b ? pj.x : ox, ya = side_b ? pj.y : oy, tha = side_b ? pj.z : oth;
    const double ca = side_b ? qj.x : oc, sa = side_b ? qj.y : os;
    const double xb = side_b ? ox : pj.x, yb = side_b ? oy : pj.y, thb = side_b ? oth : pj.z;
    EdgeLin e;
    edge_linearize<true>(xa, ya, tha, ca, sa, xb, yb, thb, tmx, tmy, thm, cm, sm, (word & kFlagDcs) != 0, P, e);
    // R = Jacobian w.r.t. the row pose, O = w.r.t. the other pose
    const double R00 = side_b ? e.b00 : e.a00, R01 = side_b ? e.b01 : e.a01, R02 = side_b ? 0.0 : e.a02;
    const double R10 = side_b ? e.b10 : e.a10, R11 = side_b ? e.b11 : e.a11, R12 = side_b ? 0.0 : e.a12;
    const double R20 = side_b ? e.b20 : e.a20, R21 = side_b ? e.b21 : e.a21, R22 = side_b ? e.b22 : e.a22;
    const double O00 = side_b ? e.a00 : e.b00, O01 = side_b ? e.a01 : e.b01, O02 = side_b ? e.a02 : 0.0;
    const double O10 = side_b ? e.a10 : e.b10, O11 = side_b ? e.a11 : e.b11, O12 = side_b ? e.a12 : 0.0;
    const double O20 = side_b ? e.a20 : e.b20, O21 = side_b ? e.a21 : e.b21, O22 = side_b ? e.a22 : e.b22;
    // diagonal block R^T R (symmetric) and gradient R^T r, accumulated in slot order
    d00 += fma(R00, R00, fma(R10, R10, R20 * R20));
    d01 += fma(R00, R01, fma(R10, R11, R20 * R21));
    d02 += fma(R00, R02, fma(R10, R12, R20 * R22));
    d11 += fma(R01, R01, fma(R11, R11, R21 * R21));
    d12 += fma(R01, R02, fma(R11, R12, R21 * R22));
    d22 += fma(R02, R02, fma(R12, R12, R22 * R22));
    g0 += fma(R00, e.r0, fma(R10, e.r1, R20 * e.r2));
    g1 += fma(R01, e.r0, fma(R11, e.r1, R21 * e.r2));
    g2 += fma(R02, e.r0, fma(R12, e.r1, R22 * e.r2));
    if (word & kFlagCost) cost += e.cost;
    // off-diagonal block R^T O, zero when the other endpoint is constant
    const double z = (word & kFlagOtherFixed) ? 0.0 : 1.0;
    double* out = Hoff + idx;
    st_stream(out + 0 * L.ldh, z * fma(R00, O00, fma(R10, O10, R20 * O20)));
    st_stream(out + 1 * L.ldh, z * fma(R00, O01, fma(R10, O11, R20 * O21)));
    st_stream(out + 2 * L.ldh, z * fma(R00, O02, fma(R10, O12, R20 * O22)));
    st_stream(out + 3 * L.ldh, z * fma(R01, O00, fma(R11, O10, R21 * O20)));
    st_stream(out + 4 * L.ldh, z * fma(R01, O01, fma(R11, O11, R21 * O21)));
    st_stream(out + 5 * L.ldh, z * fma(R01, O02, fma(R11, O12, R21 * O22)));
    st_stream(out + 6 * L.ldh, z * fma(R02, O00, fma(R12, O10, R22 * O20)));
    st_stream(out + 7 * L.ldh, z * fma(R02, O01, fma(R12, O11, R22 * O21)));
    st_stream(out + 8 * L.ldh, z * fma(R02, O02, fma(R12, O12, R22 * O22)));
  }
  if (lr < L.nrows) {
    Hdiag[0 * L.ldn + lr] = d00; Hdiag[1 * L.ldn + lr] = d01; Hdiag[2 * L.ldn + lr] = d02;
    Hdiag[3 * L.ldn + lr] = d11; Hdiag[4 * L.ldn + lr] = d12; Hdiag[5 * L.ldn + lr] = d22;
    grad[0 * L.ldn + lr] = g0; grad[1 * L.ldn + lr] = g1; grad[2 * L.ldn + lr] = g2;
  }
  double sums[2] = {cost, fma(g0, g0, fma(g1, g1, g2 * g2))};
  grid_reduce_sum<2, kRowsPerBlock>(sums, partials, tickets + 0, scal + S_COST);   // S_COST, S_GSQ
  grid_reduce_max<kRowsPerBlock>(fmax(fabs(g0), fmax(fabs(g1), fabs(g2))), partials + 2 * (size_t)gridDim.x, tickets + 1,
                                 scal + S_GMAX);
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
  const double* cm;
  const double* sm;
  const uint8_t* dcs;   // 1 when the DCS functor applies to this edge
};

constexpr int kEdgeThreads = 256;

__global__ void __launch_bounds__(kEdgeThreads)
k_cost(const double4* __restrict__ xyt, const double2* __restrict__ cs, EdgeList E, int32_t e_lo, int32_t e_hi, Params P,
       double* partials, unsigned int* ticket, double* out) {
  double cost = 0.0;
  for (int64_t e = (int64_t)e_lo + (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x; e < e_hi;
       e += (int64_t)gridDim.x * kEdgeThreads) {
    const int32_t a = E.a[e], b = E.b[e];
    const double4 pa = xyt[a], pb = xyt[b];
    const double2 qa = cs[a];
    EdgeLin L;
    edge_linearize<false>(pa.x, pa.y, pa.z, qa.x, qa.y, pb.x, pb.y, pb.z, ld_stream(E.tmx + e), ld_stream(E.tmy + e),
                          ld_stream(E.thm + e), ld_stream(E.cm + e), ld_stream(E.sm + e), E.dcs[e] != 0, P, L);
    cost += L.cost;
  }
  double s[1] = {cost};
  grid_reduce_sum<1, kEdgeThreads>(s, partials, ticket, out);
}

__global__ void __launch_bounds__(kEdgeThreads)
k_edge_eval(const double4* __restrict__ xyt, const double2* __restrict__ cs, EdgeList E, Params P, double* res, double* jac,
            double* psi, double* rho1) {
  const int64_t e = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  if (e >= E.n) return;
  const int32_t a = E.a[e], b = E.b[e];
  const double4 pa = xyt[a], pb = xyt[b];
  const double2 qa = cs[a];
  EdgeLin L;
  edge_linearize<true>(pa.x, pa.y, pa.z, qa.x, qa.y, pb.x, pb.y, pb.z, E.tmx[e], E.tmy[e], E.thm[e], E.cm[e], E.sm[e],
                       E.dcs[e] != 0, P, L);
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
__global__ void k_pack_poses(const double* __restrict__ xyt3, int32_t n, double4* xyt, double2* cs) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xyt3[3 * i], y = xyt3[3 * i + 1], th = xyt3[3 * i + 2];
  double s, c;
  sincos(th, &s, &c);
  xyt[i] = make_double4(x, y, th, 0.0);
  cs[i] = make_double2(c, s);
}
__global__ void k_unpack_poses(const double4* __restrict__ xyt, int32_t n, double* xyt3) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xyt[i];
  xyt3[3 * i] = p.x; xyt3[3 * i + 1] = p.y; xyt3[3 * i + 2] = p.z;
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
// K3: q = A p (A = diag blocks `D` + off-diagonal blocks), fused p.q; thread per row, JDS layout.
// `rotate_rz`: fold the PCG scalar rotation S_RZ <- S_RZ_NEXT into the finalising thread (this
// kernel never reads either, so there is no hazard).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerBlock)
k_spmv(const double4* __restrict__ p4, RowLayout L, const uint32_t* __restrict__ other, const double* __restrict__ Hoff,
       const double* __restrict__ D, double* __restrict__ q, double* partials, unsigned int* ticket, double* scal,
       int dot_slot, int rotate_rz) {
  const int t = threadIdx.x;
  const int64_t slot0 = (int64_t)blockIdx.x * kRowsPerBlock;
  const int lr = (int)slot0 + L.perm[slot0 + t];
  double y0 = 0, y1 = 0, y2 = 0, dot = 0;
  if (lr < L.nrows) {
    const int deg = L.row_ptr[lr + 1] - L.row_ptr[lr];
    const double4 p = p4[L.row_lo + lr];
    const double a00 = D[0 * L.ldn + lr], a01 = D[1 * L.ldn + lr], a02 = D[2 * L.ldn + lr];
    const double a11 = D[3 * L.ldn + lr], a12 = D[4 * L.ldn + lr], a22 = D[5 * L.ldn + lr];
    y0 = fma(a00, p.x, fma(a01, p.y, a02 * p.z));
    y1 = fma(a01, p.x, fma(a11, p.y, a12 * p.z));
    y2 = fma(a02, p.x, fma(a12, p.y, a22 * p.z));
    const int32_t* rp = L.round_ptr + L.rp_off[blockIdx.x];
    for (int k = 0; k < deg; ++k) {
      const int64_t idx = (int64_t)rp[k] + t;
      const uint32_t j = ld_stream_u32(other + idx) & kIdxMask;
      const double* h = Hoff + idx;
      const double h0 = ld_stream(h + 0 * L.ldh), h1 = ld_stream(h + 1 * L.ldh), h2 = ld_stream(h + 2 * L.ldh);
      const double h3 = ld_stream(h + 3 * L.ldh), h4 = ld_stream(h + 4 * L.ldh), h5 = ld_stream(h + 5 * L.ldh);
      const double h6 = ld_stream(h + 6 * L.ldh), h7 = ld_stream(h + 7 * L.ldh), h8 = ld_stream(h + 8 * L.ldh);
      const double4 pj = p4[j];
      y0 = fma(h0, pj.x, fma(h1, pj.y, fma(h2, pj.z, y0)));
      y1 = fma(h3, pj.x, fma(h4, pj.y, fma(h5, pj.z, y1)));
      y2 = fma(h6, pj.x, fma(h7, pj.y, fma(h8, pj.z, y2)));
    }
    q[0 * L.ldn + lr] = y0; q[1 * L.ldn + lr] = y1; q[2 * L.ldn + lr] = y2;
    dot = fma(p.x, y0, fma(p.y, y1, p.z * y2));
  }
  double s[1] = {dot};
  grid_reduce_sum<1, kRowsPerBlock>(s, partials, ticket, scal + dot_slot);
  // grid_reduce_sum returns in every thread; only the finalising thread sees ticket == 0 reset.
  if (rotate_rz && threadIdx.x == 0 && blockIdx.x == 0) {
    // Safe: S_RZ / S_RZ_NEXT are not read by any thread of this kernel.
    scal[S_RZ] = scal[S_RZ_NEXT];
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
// candidate = x - w (delta = -w);  S_STEP_SQ = |w|^2;  S_XSQ = |candidate|^2 over parameter rows
__global__ void __launch_bounds__(kVecThreads)
k_apply_step(const double4* __restrict__ xyt, const double* __restrict__ w, const uint8_t* __restrict__ is_free,
             int32_t row_lo, int32_t nrows, int64_t ldn, double4* cand_xyt, double2* cand_cs,
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
    double s, c;
    sincos(p.z, &s, &c);
    cand_xyt[row_lo + i] = p;
    cand_cs[row_lo + i] = make_double2(c, s);
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
__global__ void k_soa_to_aos(const double* __restrict__ v, int32_t nrows, int64_t ldn, double* out3) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  out3[3 * (int64_t)i] = v[i]; out3[3 * (int64_t)i + 1] = v[ldn + i]; out3[3 * (int64_t)i + 2] = v[2 * ldn + i];
}
__global__ void k_aos_to_soa(const double* __restrict__ in3, int32_t nrows, int64_t ldn, double* v) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  v[i] = in3[3 * (int64_t)i]; v[ldn + i] = in3[3 * (int64_t)i + 1]; v[2 * ldn + i] = in3[3 * (int64_t)i + 2];
}

}  // namespace dcs
