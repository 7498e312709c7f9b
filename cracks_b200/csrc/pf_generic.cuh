// pf_generic.cuh -- DIM-templated cell kernels, one thread per cell, dense shape
// tables.  This is the portable-in-dimension path: it serves dim = 2 and the
// secondary 3-D kernels (residual, diagonal, functionals).  The 3-D operator
// apply has its own tiled kernel in pf_apply3d.cuh.
//
// Weak form restated from cracks.cc:2235-2432 (no stress split: sigma+ = sigma).
#pragma once
#include "pf_common.cuh"
#include "pf_split2d.cuh"

namespace pf {

// deterministic mode on box meshes: is the cell one of Grid::colour (index parities in x, y, z)?
template <int DIM>
__device__ __forceinline__ bool
cell_has_colour (const Grid &g, long long lc)
{
  if (g.colour < 0 || g.conn)
    return true;
  int par = 0;
  long long rem = lc;
  for (int d = 0; d < DIM - 1; ++d)
    {
      par |= (int) ((rem % g.n[d]) & 1) << d;
      rem /= g.n[d];
    }
  par |= (int) ((rem + g.cell_begin) & 1) << (DIM - 1);
  return par == (DIM == 3 ? g.colour : (g.colour & 3));
}

template <int DIM>
__device__ __forceinline__ void
cell_nodes (const Grid &g, long long lc, long long *node)
{
  if (g.conn) // forest mesh: explicit connectivity
    {
      for (int v = 0; v < (1 << DIM); ++v)
        node[v] = g.conn[lc * (1 << DIM) + v];
      return;
    }
  // lc = local cell index within [cell_begin, cell_end) layers, x fastest
  long long rem = lc;
  int ci[3] = {0, 0, 0};
  for (int d = 0; d < DIM - 1; ++d)
    {
      ci[d] = (int) (rem % g.n[d]);
      rem /= g.n[d];
    }
  ci[DIM - 1] = (int) rem + g.cell_begin - g.plane_begin; // local plane index
  const long long sy = g.nn[0];
  const long long sz = (long long) g.nn[0] * g.nn[1];
  for (int v = 0; v < (1 << DIM); ++v)
    {
      long long id = (ci[0] + (v & 1)) + (ci[1] + ((v >> 1) & 1)) * sy;
      if (DIM == 3)
        id += (ci[2] + ((v >> 2) & 1)) * sz;
      node[v] = id;
    }
  if (DIM == 2 && g.slit_row >= 0 && ci[1] == g.slit_row)
    for (int v = 0; v < 2; ++v) // bottom vertices of the cell row above the slit: doubled nodes
      {
        const int ix = ci[0] + (v & 1);
        if (ix >= g.slit_i0)
          node[v] = g.slit_base + (ix - g.slit_i0);
      }
}

// shape table of the cell (one per refinement level on forest meshes) and its material constants
template <int DIM>
__device__ __forceinline__ const FeTab<DIM> &
cell_table (const Grid &g, const FeTab<DIM> *tab, long long lc)
{
  return g.cell_level ? tab[g.cell_level[lc]] : *tab;
}

__device__ __forceinline__ Phys
cell_phys (const Grid &g, const Phys &p, long long lc)
{
  Phys q = p;
  if (g.cell_lame) // `test case = multiple het`: Lame coefficients per cell (cracks.cc:2207-2216)
    {
      q.lambda = g.cell_lame[2 * lc];
      q.mu = g.cell_lame[2 * lc + 1];
    }
  return q;
}

// q-point state for the no-split constitutive law
template <int DIM> struct QState
{
  double pf, pf_extra, div_u, spE;
  double gpf[DIM];
  double gu[DIM][DIM];
  double sp[DIM][DIM];
  double sm[DIM][DIM]; // sigma^- (zero without the split)
  Sym2 E2;             // 2-D: strain and its eigen-decomposition, reused by the linearisation
  Eig2 eig;
};

template <int DIM>
__device__ __forceinline__ void
eval_qstate (const FeTab<DIM> &t, int q, const Phys &p, const double (*ls)[DIM + 1],
             const double *lpt, QState<DIM> &s)
{
  constexpr int NV = 1 << DIM;
  double pf = 0, pte = 0;
  for (int a = 0; a < DIM; ++a)
    {
      s.gpf[a] = 0;
      for (int b = 0; b < DIM; ++b)
        s.gu[a][b] = 0;
    }
  for (int v = 0; v < NV; ++v)
    {
      const double N = t.N[q][v];
      pf += N * ls[v][DIM];
      pte += N * lpt[v];
      for (int e = 0; e < DIM; ++e)
        {
          const double d = t.dN[q][v][e];
          s.gpf[e] += d * ls[v][DIM];
          for (int c = 0; c < DIM; ++c)
            s.gu[c][e] += d * ls[v][c];
        }
    }
  if (p.clamp_extra)
    pte = fmin (fmax (pte, 0.0), 1.0);
  s.pf = pf;
  s.pf_extra = pte;
  double tr = 0;
  for (int a = 0; a < DIM; ++a)
    tr += s.gu[a][a];
  s.div_u = tr;
  double spE = 0;
  for (int a = 0; a < DIM; ++a)
    for (int b = 0; b < DIM; ++b)
      {
        const double E = 0.5 * (s.gu[a][b] + s.gu[b][a]);
        s.sp[a][b] = (a == b ? p.lambda * tr : 0.0) + 2 * p.mu * E;
        s.sm[a][b] = 0.0;
        spE += s.sp[a][b] * E;
      }
  if (DIM == 2 && p.split) // cracks.cc:2294-2300
    {
      s.E2.xx = s.gu[0][0];
      s.E2.yy = s.gu[1][1];
      s.E2.xy = 0.5 * (s.gu[0][1] + s.gu[1][0]);
      s.eig = eig_sym2 (s.E2);
      Sym2 sp, sm;
      split_stress (s.E2, s.eig, p.lambda, p.mu, sp, sm);
      s.sp[0][0] = sp.xx, s.sp[0][1] = s.sp[1][0] = sp.xy, s.sp[1][1] = sp.yy;
      s.sm[0][0] = sm.xx, s.sm[0][1] = s.sm[1][0] = sm.xy, s.sm[1][1] = sm.yy;
      spE = sp.xx * s.E2.xx + 2.0 * sp.xy * s.E2.xy + sp.yy * s.E2.yy;
    }
  s.spE = spE;
}

// ---- y += J x on the cells of this rank (rows/cols of constrained dofs dropped)
template <int DIM>
__global__ void __launch_bounds__ (128)
k_apply_generic (Grid g, Phys p_in, const FeTab<DIM> *__restrict__ tab,
                 const double *__restrict__ x, const double *__restrict__ sol,
                 const double *__restrict__ pt, const uint8_t *__restrict__ mask,
                 double *__restrict__ y)
{
  constexpr int NC = DIM + 1, NV = 1 << DIM, NQ = FeTab<DIM>::NQ;
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= g.n_local_cells)
    return;
  const FeTab<DIM> &t = cell_table<DIM> (g, tab, lc);
  const Phys p = cell_phys (g, p_in, lc);
  long long node[NV];
  cell_nodes<DIM> (g, lc, node);
  double lx[NV][NC], ls[NV][NC], lpt[NV], out[NV][NC];
  uint8_t lm[NV];
  for (int v = 0; v < NV; ++v)
    {
      lm[v] = mask[node[v]];
      lpt[v] = pt[node[v]];
      for (int c = 0; c < NC; ++c)
        {
          ls[v][c] = sol[node[v] * NC + c];
          lx[v][c] = is_constrained (lm[v], c) ? 0.0 : x[node[v] * NC + c];
          out[v][c] = 0;
        }
    }
  for (int q = 0; q < NQ; ++q)
    {
      QState<DIM> s;
      eval_qstate<DIM> (t, q, p, ls, lpt, s);
      double G[DIM][DIM], dphi = 0, gdphi[DIM];
      for (int a = 0; a < DIM; ++a)
        {
          gdphi[a] = 0;
          for (int b = 0; b < DIM; ++b)
            G[a][b] = 0;
        }
      for (int v = 0; v < NV; ++v)
        {
          dphi += t.N[q][v] * lx[v][DIM];
          for (int e = 0; e < DIM; ++e)
            {
              const double d = t.dN[q][v][e];
              gdphi[e] += d * lx[v][DIM];
              for (int c = 0; c < DIM; ++c)
                G[c][e] += d * lx[v][c];
            }
        }
      double trG = 0, spG = 0;
      for (int a = 0; a < DIM; ++a)
        {
          trG += G[a][a];
          for (int b = 0; b < DIM; ++b)
            spG += s.sp[a][b] * G[a][b];
        }
      const double gdeg = (1.0 - p.kappa) * s.pf_extra * s.pf_extra + p.kappa;
      double Sig[DIM][DIM];
      // sigma+'(u; du):E(u) + sigma+(u):E(du); equal for the unsplit law
      double cross = 2.0 * spG;
      if (DIM == 2 && p.split)
        {
          // the linearisation is linear in E(du), so it is applied to the interpolated
          // direction instead of once per trial function (cracks.cc:2338-2345, 2359-2364)
          Sym2 L, spL, smL;
          L.xx = G[0][0], L.yy = G[1][1], L.xy = 0.5 * (G[0][1] + G[1][0]);
          split_stress_lin (s.E2, s.eig, L, p.lambda, p.mu, spL, smL);
          Sig[0][0] = gdeg * spL.xx + p.d_mat * smL.xx;
          Sig[1][1] = gdeg * spL.yy + p.d_mat * smL.yy;
          Sig[0][1] = Sig[1][0] = gdeg * spL.xy + p.d_mat * smL.xy;
          cross = (spL.xx * s.E2.xx + 2.0 * spL.xy * s.E2.xy + spL.yy * s.E2.yy) + spG;
        }
      else
        for (int a = 0; a < DIM; ++a)
          for (int b = 0; b < DIM; ++b)
            Sig[a][b] = gdeg * ((a == b ? p.lambda * trG : 0.0) + p.mu * (G[a][b] + G[b][a]));
      // (phi,u) + (phi,phi) value coefficient, cracks.cc:2375-2382
      const double a_val = s.pf * ((1.0 - p.kappa) * cross - 2.0 * p.P1 * trG)
                           + dphi * ((1.0 - p.kappa) * s.spE + p.G_c / p.eps - 2.0 * p.P1 * s.div_u);
      const double w = t.JxW[q];
      for (int v = 0; v < NV; ++v)
        {
          double gb = 0;
          for (int e = 0; e < DIM; ++e)
            {
              const double d = t.dN[q][v][e];
              gb += gdphi[e] * d;
              for (int c = 0; c < DIM; ++c)
                out[v][c] += w * Sig[c][e] * d;
            }
          out[v][DIM] += w * (a_val * t.N[q][v] + p.G_c * p.eps * gb);
        }
    }
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < NC; ++c)
      if (!is_constrained (lm[v], c))
        atomicAdd (&y[node[v] * NC + c], out[v][c]);
}

// ---- r_total += local_rhs (cracks.cc:2393-2432); constraints applied later
template <int DIM>
__global__ void __launch_bounds__ (128)
k_residual_generic (Grid g, Phys p_in, const FeTab<DIM> *__restrict__ tab,
                    const double *__restrict__ sol, const double *__restrict__ pt,
                    double *__restrict__ r)
{
  constexpr int NC = DIM + 1, NV = 1 << DIM, NQ = FeTab<DIM>::NQ;
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= g.n_local_cells)
    return;
  const FeTab<DIM> &t = cell_table<DIM> (g, tab, lc);
  const Phys p = cell_phys (g, p_in, lc);
  long long node[NV];
  cell_nodes<DIM> (g, lc, node);
  double ls[NV][NC], lpt[NV], out[NV][NC];
  for (int v = 0; v < NV; ++v)
    {
      lpt[v] = pt[node[v]];
      for (int c = 0; c < NC; ++c)
        {
          ls[v][c] = sol[node[v] * NC + c];
          out[v][c] = 0;
        }
    }
  for (int q = 0; q < NQ; ++q)
    {
      QState<DIM> s;
      eval_qstate<DIM> (t, q, p, ls, lpt, s);
      const double gdeg = (1.0 - p.kappa) * s.pf_extra * s.pf_extra + p.kappa;
      const double w = t.JxW[q];
      const double cphi = (1.0 - p.kappa) * s.spE * s.pf - p.G_c / p.eps * (1.0 - s.pf)
                          - 2.0 * p.P1 * s.pf * s.div_u;
      const double pe2 = p.P1 * s.pf_extra * s.pf_extra;
      for (int v = 0; v < NV; ++v)
        {
          double gg = 0;
          for (int e = 0; e < DIM; ++e)
            gg += s.gpf[e] * t.dN[q][v][e];
          for (int c = 0; c < DIM; ++c)
            {
              double sc = 0;
              for (int e = 0; e < DIM; ++e)
                sc += (gdeg * s.sp[c][e] + p.d_rhs * s.sm[c][e]) * t.dN[q][v][e]; // sm = 0 without the split
              out[v][c] -= w * (sc - pe2 * t.dN[q][v][c]);
            }
          out[v][DIM] -= w * (cphi * t.N[q][v] + p.G_c * p.eps * gg);
        }
    }
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < NC; ++c)
      atomicAdd (&r[node[v] * NC + c], out[v][c]);
}

// ---- diag += sum_cells |local_matrix(i,i)| (average |diag| if zero), the value
// AffineConstraints::distribute_local_to_global leaves on constrained rows.
template <int DIM>
__global__ void __launch_bounds__ (128)
k_diag_generic (Grid g, Phys p_in, const FeTab<DIM> *__restrict__ tab,
                const double *__restrict__ sol, const double *__restrict__ pt,
                double *__restrict__ diag)
{
  constexpr int NC = DIM + 1, NV = 1 << DIM, NQ = FeTab<DIM>::NQ;
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= g.n_local_cells || !cell_has_colour<DIM> (g, lc))
    return;
  const FeTab<DIM> &t = cell_table<DIM> (g, tab, lc);
  const Phys p = cell_phys (g, p_in, lc);
  long long node[NV];
  cell_nodes<DIM> (g, lc, node);
  double ls[NV][NC], lpt[NV], out[NV][NC];
  for (int v = 0; v < NV; ++v)
    {
      lpt[v] = pt[node[v]];
      for (int c = 0; c < NC; ++c)
        {
          ls[v][c] = sol[node[v] * NC + c];
          out[v][c] = 0;
        }
    }
  for (int q = 0; q < NQ; ++q)
    {
      QState<DIM> s;
      eval_qstate<DIM> (t, q, p, ls, lpt, s);
      const double gdeg = (1.0 - p.kappa) * s.pf_extra * s.pf_extra + p.kappa;
      const double w = t.JxW[q];
      const double c0 = (1.0 - p.kappa) * s.spE + p.G_c / p.eps - 2.0 * p.P1 * s.div_u;
      for (int v = 0; v < NV; ++v)
        {
          double g2 = 0;
          for (int e = 0; e < DIM; ++e)
            g2 += t.dN[q][v][e] * t.dN[q][v][e];
          const double N = t.N[q][v];
          for (int c = 0; c < DIM; ++c)
            {
              const double d = t.dN[q][v][c];
              if (DIM == 2 && p.split)
                {
                  // trial = test = e_c N_v: E_lin = sym(e_c (x) grad N_v)
                  Sym2 L, spL, smL;
                  const double dx = t.dN[q][v][0], dy = t.dN[q][v][1];
                  L.xx = c == 0 ? dx : 0.0;
                  L.yy = c == 1 ? dy : 0.0;
                  L.xy = 0.5 * (c == 0 ? dy : dx);
                  split_stress_lin (s.E2, s.eig, L, p.lambda, p.mu, spL, smL);
                  const double sx = c == 0 ? gdeg * spL.xx + p.d_mat * smL.xx : gdeg * spL.xy + p.d_mat * smL.xy;
                  const double sy = c == 0 ? gdeg * spL.xy + p.d_mat * smL.xy : gdeg * spL.yy + p.d_mat * smL.yy;
                  out[v][c] += w * (sx * dx + sy * dy);
                  continue;
                }
              // g * (lambda d_c^2 + mu (|grad N|^2 + d_c^2))
              out[v][c] += w * gdeg * (p.lambda * d * d + p.mu * (g2 + d * d));
            }
          out[v][DIM] += w * (c0 * N * N + p.G_c * p.eps * g2);
        }
    }
  double avg = 0;
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < NC; ++c)
      avg += fabs (out[v][c]);
  avg /= (NV * NC);
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < NC; ++c)
      {
        const double d = fabs (out[v][c]);
        atomicAdd (&diag[node[v] * NC + c], d != 0.0 ? d : avg);
      }
}

// ---- functionals: bulk / crack energy (cracks.cc:3663-3681), TCV (3585-3586)
// block-reduced, one atomicAdd per block into out[0..2]
template <int DIM>
__global__ void __launch_bounds__ (128)
k_functionals_generic (Grid g, Phys p_in, const FeTab<DIM> *__restrict__ tab,
                       const double *__restrict__ sol, int owned_cells_only_begin,
                       int owned_cells_only_end, double *__restrict__ out3)
{
  constexpr int NC = DIM + 1, NV = 1 << DIM, NQ = FeTab<DIM>::NQ;
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  double eb = 0, ec = 0, tcv = 0;
  bool active = lc < g.n_local_cells;
  if (active)
    {
      // cells are owned by exactly one rank: layers [owned_cells_only_begin, end)
      long long per_layer = g.n[0];
      if (DIM == 3)
        per_layer *= g.n[1];
      const int layer = (int) (lc / per_layer) + g.cell_begin;
      active = layer >= owned_cells_only_begin && layer < owned_cells_only_end;
    }
  if (active)
    {
      const FeTab<DIM> &t = cell_table<DIM> (g, tab, lc);
      const Phys p = cell_phys (g, p_in, lc);
      long long node[NV];
      cell_nodes<DIM> (g, lc, node);
      double ls[NV][NC];
      for (int v = 0; v < NV; ++v)
        for (int c = 0; c < NC; ++c)
          ls[v][c] = sol[node[v] * NC + c];
      for (int q = 0; q < NQ; ++q)
        {
          double pf = 0, gpf[DIM], u[DIM], gu[DIM][DIM];
          for (int a = 0; a < DIM; ++a)
            {
              gpf[a] = 0;
              u[a] = 0;
              for (int b = 0; b < DIM; ++b)
                gu[a][b] = 0;
            }
          for (int v = 0; v < NV; ++v)
            {
              const double N = t.N[q][v];
              pf += N * ls[v][DIM];
              for (int e = 0; e < DIM; ++e)
                {
                  u[e] += N * ls[v][e];
                  const double d = t.dN[q][v][e];
                  gpf[e] += d * ls[v][DIM];
                  for (int c = 0; c < DIM; ++c)
                    gu[c][e] += d * ls[v][c];
                }
            }
          double trE = 0, trE2 = 0, gg = 0, ug = 0;
          for (int a = 0; a < DIM; ++a)
            {
              trE += gu[a][a];
              gg += gpf[a] * gpf[a];
              ug += u[a] * gpf[a];
              for (int b = 0; b < DIM; ++b)
                {
                  const double E = 0.5 * (gu[a][b] + gu[b][a]);
                  trE2 += E * E;
                }
            }
          const double psi = 0.5 * p.lambda * trE * trE + p.mu * trE2;
          const double w = t.JxW[q];
          eb += ((1 + p.kappa) * pf * pf + p.kappa) * psi * w;
          ec += p.G_c / 2.0 * ((pf - 1) * (pf - 1) / p.eps + p.eps * gg) * w;
          tcv += ug * w;
        }
    }
  __shared__ double red[3][4];
  for (int o = 16; o > 0; o >>= 1)
    {
      eb += __shfl_down_sync (0xffffffffu, eb, o);
      ec += __shfl_down_sync (0xffffffffu, ec, o);
      tcv += __shfl_down_sync (0xffffffffu, tcv, o);
    }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
    {
      red[0][w] = eb;
      red[1][w] = ec;
      red[2][w] = tcv;
    }
  __syncthreads ();
  if (threadIdx.x < 3)
    {
      double s = 0;
      for (int i = 0; i < (int) (blockDim.x >> 5); ++i)
        s += red[threadIdx.x][i];
      atomicAdd (&out3[threadIdx.x], s);
    }
}

// ---- crack opening displacement on the plane x = eval_line, cracks.cc:3452-3549:
// every cell face on that plane (visited from both sides; the caller halves the
// sum), QGauss<dim-1>(3), sum of 0.5 u . grad(phi) JxW.  out2[0] += value,
// out2[1] += number of faces.  Cells of layers [own_begin, own_end) only.
template <int DIM>
__global__ void __launch_bounds__ (128)
k_cod_generic (Grid g, const double *__restrict__ sol, double eval_line, int own_begin, int own_end,
               double *__restrict__ out2)
{
  constexpr int NC = DIM + 1, NV = 1 << DIM;
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  double cod = 0, faces = 0;
  bool active = lc < g.n_local_cells;
  int cx = 0;
  if (active)
    {
      long long per_layer = g.n[0];
      if (DIM == 3)
        per_layer *= g.n[1];
      const int layer = (int) (lc / per_layer) + g.cell_begin;
      active = layer >= own_begin && layer < own_end;
      cx = (int) (lc % g.n[0]);
    }
  if (active)
    {
      const double gq = 0.5 * sqrt (3.0 / 5.0);
      const double xi[3] = {0.5 - gq, 0.5, 0.5 + gq};
      const double w[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
      for (int side = 0; side < 2; ++side)
        {
          const double fx = g.origin[0] + (cx + side) * g.h[0];
          if (!(fx < eval_line + 1e-8 && fx > eval_line - 1e-8))
            continue;
          faces += 1.0;
          long long node[NV];
          cell_nodes<DIM> (g, lc, node);
          double ls[NV][NC];
          for (int v = 0; v < NV; ++v)
            for (int c = 0; c < NC; ++c)
              ls[v][c] = sol[node[v] * NC + c];
          for (int qz = 0; qz < (DIM == 3 ? 3 : 1); ++qz)
            for (int qy = 0; qy < 3; ++qy)
              {
                const double pt[3] = {(double) side, xi[qy], xi[qz]};
                double JxW = g.h[1] * w[qy];
                if (DIM == 3)
                  JxW *= g.h[2] * w[qz];
                double u[DIM], gpf[DIM];
                for (int e = 0; e < DIM; ++e)
                  u[e] = gpf[e] = 0;
                for (int v = 0; v < NV; ++v)
                  {
                    double N = 1.0;
                    for (int d = 0; d < DIM; ++d)
                      N *= ((v >> d) & 1) ? pt[d] : 1.0 - pt[d];
                    for (int e = 0; e < DIM; ++e)
                      {
                        double gr = 1.0;
                        for (int d = 0; d < DIM; ++d)
                          {
                            const int b = (v >> d) & 1;
                            gr *= (d == e) ? (b ? 1.0 : -1.0) / g.h[d] : (b ? pt[d] : 1.0 - pt[d]);
                          }
                        u[e] += N * ls[v][e];
                        gpf[e] += gr * ls[v][DIM];
                      }
                  }
                double dot = 0;
                for (int e = 0; e < DIM; ++e)
                  dot += u[e] * gpf[e];
                cod += 0.5 * dot * JxW;
              }
        }
    }
  __shared__ double red[2][4];
  for (int o = 16; o > 0; o >>= 1)
    {
      cod += __shfl_down_sync (0xffffffffu, cod, o);
      faces += __shfl_down_sync (0xffffffffu, faces, o);
    }
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0)
    {
      red[0][wi] = cod;
      red[1][wi] = faces;
    }
  __syncthreads ();
  if (threadIdx.x < 2)
    {
      double s = 0;
      for (int i = 0; i < (int) (blockDim.x >> 5); ++i)
        s += red[threadIdx.x][i];
      if (s != 0.0)
        atomicAdd (&out2[threadIdx.x], s);
    }
}

// ---- phi-block lumped mass by a cell loop (cracks.cc:2514-2562: vertex quadrature,
// vol / 2^dim per cell vertex); used where the node -> cells count is not implied
// by the node index (slit meshes)
template <int DIM>
__global__ void __launch_bounds__ (128)
k_lumped_mass_cells (Grid g, double *__restrict__ mass)
{
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= g.n_local_cells)
    return;
  long long node[1 << DIM];
  cell_nodes<DIM> (g, lc, node);
  double vol = 1;
  for (int d = 0; d < DIM; ++d)
    vol *= g.h[d];
  for (int v = 0; v < (1 << DIM); ++v)
    atomicAdd (&mass[node[v]], vol / (1 << DIM));
}

// ---- load on boundary id 3 = the top edge (cracks.cc:3728-3816): int sigma(u) n ds,
// n = (0,1), QGauss<1>(3), undegraded stress.  out2 += (sigma_xy, sigma_yy) integrals.
__global__ void __launch_bounds__ (128)
k_load_top_2d (Grid g, Phys p, const double *__restrict__ sol, double *__restrict__ out2)
{
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;
  double lx = 0, ly = 0;
  if (cx < g.n[0])
    {
      long long node[4];
      cell_nodes<2> (g, cx + (long long) (g.n[1] - 1) * g.n[0], node);
      const double gq = 0.5 * sqrt (3.0 / 5.0);
      const double xi[3] = {0.5 - gq, 0.5, 0.5 + gq};
      const double wq[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
      for (int q = 0; q < 3; ++q)
        {
          const double pt[2] = {xi[q], 1.0};
          double gu[2][2] = {{0, 0}, {0, 0}};
          for (int v = 0; v < 4; ++v)
            for (int e = 0; e < 2; ++e)
              {
                double gr = 1.0;
                for (int d = 0; d < 2; ++d)
                  {
                    const int b = (v >> d) & 1;
                    gr *= (d == e) ? (b ? 1.0 : -1.0) / g.h[d] : (b ? pt[d] : 1.0 - pt[d]);
                  }
                for (int c = 0; c < 2; ++c)
                  gu[c][e] += gr * sol[node[v] * 3 + c];
              }
          const double tr = gu[0][0] + gu[1][1];
          const double JxW = g.h[0] * wq[q];
          lx += p.mu * (gu[0][1] + gu[1][0]) * JxW;
          ly += (p.lambda * tr + 2.0 * p.mu * gu[1][1]) * JxW;
        }
    }
  for (int o = 16; o > 0; o >>= 1)
    {
      lx += __shfl_down_sync (0xffffffffu, lx, o);
      ly += __shfl_down_sync (0xffffffffu, ly, o);
    }
  if ((threadIdx.x & 31) == 0)
    {
      atomicAdd (&out2[0], lx);
      atomicAdd (&out2[1], ly);
    }
}

// ---- Dirichlet rows and values of the Miehe tests on the unit square with slit
// (cracks.cc:2584-2625 with BoundaryTensionTest 780-797 / BoundaryShearTest 845-861).
// kind 1 = tension: u_y = 0 on y = 0, u = (0, t) on y = 1;
// kind 2 = shear:   u_y = 0 on x = 0 and x = 1, u = 0 on y = 0, u = (-t, 0) on y = 1,
//                   u_y = 0 on the lower face of the slit (boundary id 4).
// set_values = 0: only (re)build the mask bits; 1: also write the values into sol.
__global__ void
k_dirichlet_miehe (Grid g, int kind, double time, int set_values, uint8_t *__restrict__ mask,
                   double *__restrict__ sol)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.n_local_nodes)
    return;
  const bool dup = g.slit_row >= 0 && n >= g.slit_base;
  const int i = dup ? (int) (n - g.slit_base) + g.slit_i0 : (int) (n % g.nn[0]);
  const int j = dup ? g.slit_row : (int) (n / g.nn[0]);
  const bool top = j == g.nn[1] - 1, bottom = j == 0, left = i == 0, right = i == g.nn[0] - 1;
  bool cx = false, cy = false;
  double vx = 0, vy = 0;
  if (kind == 1)
    {
      cy = bottom || top;
      cx = top;
      if (top)
        vy = time;
    }
  else
    {
      const bool lower_slit = !dup && g.slit_row >= 0 && j == g.slit_row && i >= g.slit_i0 - 1;
      cy = left || right || bottom || top || lower_slit;
      cx = bottom || top;
      if (top)
        vx = -time;
    }
  uint8_t m = mask[n] & (uint8_t) 4u; // keep the active-set bit of phi
  if (cx)
    m |= 1u;
  if (cy)
    m |= 2u;
  mask[n] = m;
  if (set_values)
    {
      if (cx)
        sol[n * 3 + 0] = vx;
      if (cy)
        sol[n * 3 + 1] = vy;
    }
}

} // namespace pf
