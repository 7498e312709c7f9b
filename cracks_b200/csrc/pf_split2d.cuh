// pf_split2d.cuh -- Miehe's tensile/compressive stress split in 2-D and its
// linearisation, evaluated per quadrature point inside the cell kernels.
//
// What is computed follows tjhei/cracks: eigen_vectors_and_values
// (cracks.cc:1691-1737) and decompose_stress (cracks.cc:1923-2120), including
// the "close to diagonal" shortcut and the unguarded divisions by E01 of the
// derivative branch.  Symmetric 2x2 tensors are held as {xx, xy, yy}.
#pragma once
#include "pf_common.cuh"

namespace pf {

struct Sym2
{
  double xx, xy, yy;
};

struct Eig2
{
  double l1, l2;   // eigenvalues
  double c1, s1;   // first eigenvector  (c1, s1)
  double c2, s2;   // second eigenvector (c2, s2)
  double q1, q2;   // (l - Exx) / Exy, the slope the eigenvectors are built from
  bool diagonal;   // shortcut branch taken (cracks.cc:1700-1710)
};

__device__ __forceinline__ Eig2
eig_sym2 (const Sym2 &E)
{
  Eig2 r;
  r.diagonal = fabs (E.xy) < 1e-10 * fabs (E.xx) || fabs (E.xy) < 1e-10 * fabs (E.yy);
  if (r.diagonal)
    {
      r.l1 = E.xx, r.c1 = 1.0, r.s1 = 0.0;
      r.l2 = E.yy, r.c2 = 0.0, r.s2 = 1.0;
      r.q1 = (r.l1 - E.xx) / E.xy;
      r.q2 = (r.l2 - E.xx) / E.xy;
      return r;
    }
  const double dd = E.xx - E.yy;
  const double sq = sqrt (dd * dd + 4.0 * E.xy * E.xy);
  r.l1 = 0.5 * ((E.xx + E.yy) + sq);
  r.l2 = 0.5 * ((E.xx + E.yy) - sq);
  r.q1 = (r.l1 - E.xx) / E.xy;
  r.q2 = (r.l2 - E.xx) / E.xy;
  const double n1 = 1.0 / sqrt (1.0 + r.q1 * r.q1), n2 = 1.0 / sqrt (1.0 + r.q2 * r.q2);
  r.c1 = n1, r.s1 = r.q1 * n1;
  r.c2 = n2, r.s2 = r.q2 * n2;
  return r;
}

// P diag(a, b) P^T for P = [v1 v2]
__device__ __forceinline__ Sym2
spectral (const Eig2 &e, double a, double b)
{
  Sym2 r;
  r.xx = a * e.c1 * e.c1 + b * e.c2 * e.c2;
  r.xy = a * e.c1 * e.s1 + b * e.c2 * e.s2;
  r.yy = a * e.s1 * e.s1 + b * e.s2 * e.s2;
  return r;
}

// sigma+ / sigma- of the state (derivative = false branch, cracks.cc:1955-1970)
__device__ __forceinline__ void
split_stress (const Sym2 &E, const Eig2 &e, double lambda, double mu, Sym2 &sp, Sym2 &sm)
{
  const Sym2 Ep = spectral (e, fmax (0.0, e.l1), fmax (0.0, e.l2));
  const double tr = E.xx + E.yy, trp = fmax (0.0, tr);
  sp.xx = lambda * trp + 2.0 * mu * Ep.xx;
  sp.yy = lambda * trp + 2.0 * mu * Ep.yy;
  sp.xy = 2.0 * mu * Ep.xy;
  sm.xx = lambda * (tr - trp) + 2.0 * mu * (E.xx - Ep.xx);
  sm.yy = lambda * (tr - trp) + 2.0 * mu * (E.yy - Ep.yy);
  sm.xy = 2.0 * mu * (E.xy - Ep.xy);
}

// directional derivative of (sigma+, sigma-) at E in direction L (cracks.cc:1971-2109).
// The result of P' L+ P^T + P L+' P^T + P L+ P'^T is symmetric; it is stored as such.
__device__ __forceinline__ void
split_stress_lin (const Sym2 &E, const Eig2 &e, const Sym2 &L, double lambda, double mu, Sym2 &sp, Sym2 &sm)
{
  const double trL = L.xx + L.yy, tr = E.xx + E.yy;
  const double dd = E.xx - E.yy;
  const double disk = sqrt (E.xy * E.xy + dd * dd / 4.0);
  const double num = (L.xy * E.xy + E.xy * L.xy + dd * (L.xx - L.yy) / 2.0);
  const double l1L = 0.5 * trL + 1.0 / (2.0 * disk) * num;
  const double l2L = 0.5 * trL - 1.0 / (2.0 * disk) * num;
  // the derivative branch always uses the general eigenvector formulas (no diagonal shortcut)
  const double q1 = (e.l1 - E.xx) / E.xy, q2 = (e.l2 - E.xx) / E.xy;
  const double n1 = 1.0 / sqrt (1.0 + q1 * q1), n2 = 1.0 / sqrt (1.0 + q2 * q2);
  const double dq1 = ((l1L - L.xx) * E.xy - (e.l1 - E.xx) * L.xy) / (E.xy * E.xy);
  const double dq2 = ((l2L - L.xx) * E.xy - (e.l2 - E.xx) * L.xy) / (E.xy * E.xy);
  const double n1L = -1.0 * (1.0 / (1.0 + q1 * q1) * 1.0 / (2.0 * sqrt (1.0 + q1 * q1)) * (2.0 * q1) * dq1);
  const double n2L = -1.0 * (1.0 / (1.0 + q2 * q2) * 1.0 / (2.0 * sqrt (1.0 + q2 * q2)) * (2.0 * q2) * dq2);
  const double c1L = n1L, s1L = n1 * dq1 + n1L * q1;
  const double c2L = n2L, s2L = n2 * dq2 + n2L * q2;
  const double a = fmax (0.0, e.l1), b = fmax (0.0, e.l2);
  const double aL = (e.l1 < 0.0) ? 0.0 : l1L, bL = (e.l2 < 0.0) ? 0.0 : l2L; // keyed on the eigenvalue of E (2080-2094)
  // E+' = P' L+ P^T + P L+' P^T + P L+ P'^T with P = [v1 v2], P' = [v1' v2']
  Sym2 EpL;
  EpL.xx = 2.0 * (a * c1L * e.c1 + b * c2L * e.c2) + aL * e.c1 * e.c1 + bL * e.c2 * e.c2;
  EpL.yy = 2.0 * (a * s1L * e.s1 + b * s2L * e.s2) + aL * e.s1 * e.s1 + bL * e.s2 * e.s2;
  EpL.xy = a * (c1L * e.s1 + e.c1 * s1L) + b * (c2L * e.s2 + e.c2 * s2L) + aL * e.c1 * e.s1 + bL * e.c2 * e.s2;
  const double trpL = (tr < 0.0) ? 0.0 : trL;
  sp.xx = lambda * trpL + 2.0 * mu * EpL.xx;
  sp.yy = lambda * trpL + 2.0 * mu * EpL.yy;
  sp.xy = 2.0 * mu * EpL.xy;
  sm.xx = lambda * (trL - trpL) + 2.0 * mu * (L.xx - EpL.xx);
  sm.yy = lambda * (trL - trpL) + 2.0 * mu * (L.yy - EpL.yy);
  sm.xy = 2.0 * mu * (L.xy - EpL.xy);
}

} // namespace pf
