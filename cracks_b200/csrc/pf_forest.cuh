// pf_forest.cuh -- kernels for locally refined (forest) meshes with hanging nodes.
//
// Written against the CPU oracle's formulation (oracle/adaptive_oracle.py, which
// reproduces the reference's sneddon_2d_1 / miehe_shear_1 / hetero_3d_1 goldens); on the GPU the same goldens are
// reproduced by tests/test_gpu_forest.py.  Only reachable through pf_create_forest[_distributed].  The box-mesh paths do not use this file.
//
// The reference resolves hanging-node constraints inside distribute_local_to_global
// (cracks.cc:2439-2464) with the AffineConstraints built at 1630-1642.  Here they are applied
// algebraically around the unconstrained cell kernels: with H the interpolation matrix (identity on
// regular dofs, row of weights 1/n_parents on a hanging dof) the condensed operator is H^T J H, so
//   x_full = H x   (k_hanging_distribute, before the cell kernel),
//   y      = H^T y (k_hanging_fold, after it: a hanging row is added to its parents and cleared).
// Table layout: 5 entries per hanging node: node, parent 0..3 (-1 = unused).
#pragma once
#include "pf_common.cuh"

namespace pf {

// v[h] = mean of v[parents]; parents constrained in component c count as zero when
// zero_constrained != 0 (columns of Dirichlet / active dofs are dropped from the operator)
template <int NC>
__global__ void
k_hanging_distribute (long long n_hanging, const long long *__restrict__ table, const uint8_t *__restrict__ mask,
                      int zero_constrained, double *__restrict__ v)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_hanging * NC)
    return;
  const long long h = t / NC;
  const int c = (int) (t % NC);
  const long long *row = table + 5 * h;
  double sum = 0;
  int np = 0;
  for (int q = 0; q < 4; ++q)
    {
      const long long p = row[1 + q];
      if (p < 0)
        continue;
      ++np;
      if (!(zero_constrained && is_constrained (mask[p], c)))
        sum += v[p * NC + c];
    }
  v[row[0] * NC + c] = sum / np;
}

// y[parents] += y[h] / n_parents (rows of constrained parents are dropped), then
// y[h] = diag[h] * x[h] if diag != nullptr (the decoupled row of the condensed operator), else 0
template <int NC>
__global__ void
k_hanging_fold (long long n_hanging, const long long *__restrict__ table, const uint8_t *__restrict__ mask,
                const double *__restrict__ diag, const double *__restrict__ x, double *__restrict__ y)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_hanging * NC)
    return;
  const long long h = t / NC;
  const int c = (int) (t % NC);
  const long long *row = table + 5 * h;
  const long long i = row[0] * NC + c;
  int np = 0;
  for (int q = 0; q < 4; ++q)
    np += row[1 + q] >= 0;
  const double val = y[i] / np;
  for (int q = 0; q < 4; ++q)
    {
      const long long p = row[1 + q];
      if (p >= 0 && !is_constrained (mask[p], c))
        atomicAdd (&y[p * NC + c], val);
    }
  y[i] = diag ? diag[i] * x[i] : 0.0;
}

// Jacobi diagonal of the condensed operator, without the cross terms: d[parent] += d[h] / n_parents^2
template <int NC>
__global__ void
k_hanging_fold_diag (long long n_hanging, const long long *__restrict__ table, double *__restrict__ d)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_hanging * NC)
    return;
  const long long h = t / NC;
  const int c = (int) (t % NC);
  const long long *row = table + 5 * h;
  int np = 0;
  for (int q = 0; q < 4; ++q)
    np += row[1 + q] >= 0;
  const double val = d[row[0] * NC + c] / (np * np);
  for (int q = 0; q < 4; ++q)
    if (row[1 + q] >= 0)
      atomicAdd (&d[row[1 + q] * NC + c], val);
}

__global__ void
k_mark_hanging (long long n_hanging, const long long *__restrict__ table, uint8_t *__restrict__ mask)
{
  const long long h = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n_hanging)
    mask[table[5 * h]] |= (uint8_t) PF_HANGING_BIT;
}

// lumped mass on a forest: vol(cell) / 2^dim per cell vertex, vol from the level's shape table
template <int DIM>
__global__ void __launch_bounds__ (128)
k_lumped_mass_forest (Grid g, const FeTab<DIM> *__restrict__ tab, double *__restrict__ mass)
{
  const long long lc = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= g.n_local_cells)
    return;
  const FeTab<DIM> &t = tab[g.cell_level[lc]];
  double vol = 0;
  for (int q = 0; q < FeTab<DIM>::NQ; ++q)
    vol += t.JxW[q];
  for (int v = 0; v < (1 << DIM); ++v)
    atomicAdd (&mass[g.conn[lc * (1 << DIM) + v]], vol / (1 << DIM));
}

// set_initial_bc(): constraints.distribute(solution) for inhomogeneous Dirichlet data (cracks.cc:2700-2707):
// sol = vals on the displacement dofs whose Dirichlet bit is set
template <int DIM>
__global__ void
k_set_dirichlet_values (long long n_nodes, const uint8_t *__restrict__ mask, const double *__restrict__ vals,
                        double *__restrict__ sol)
{
  const long long n = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes)
    return;
  const uint8_t m = mask[n];
  for (int c = 0; c < DIM; ++c)
    if ((m >> c) & 1)
      sol[n * (DIM + 1) + c] = vals[n * (DIM + 1) + c];
}

// compute_load() (cracks.cc:3728-3816) on a forest: the listed cells have their top edge (vertices 2, 3)
// on boundary id 3; int sigma(u) n ds with n = (0, 1), QGauss<1>(3), undegraded stress.
// out2 += (integral of sigma_xy, integral of sigma_yy); one thread per listed cell, atomics.
__global__ void __launch_bounds__ (128)
k_load_top_forest (Grid g, Phys p_in, const double *__restrict__ level_h, long long n_list,
                   const long long *__restrict__ list, const double *__restrict__ sol, double *__restrict__ out2)
{
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_list)
    return;
  const long long lc = list[t];
  Phys p = p_in;
  if (g.cell_lame)
    {
      p.lambda = g.cell_lame[2 * lc];
      p.mu = g.cell_lame[2 * lc + 1];
    }
  const double hx = level_h[2 * g.cell_level[lc]], hy = level_h[2 * g.cell_level[lc] + 1];
  const double gq = 0.5 * sqrt (3.0 / 5.0);
  const double xi[3] = {0.5 - gq, 0.5, 0.5 + gq};
  const double wq[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  double lx = 0, ly = 0;
  for (int q = 0; q < 3; ++q)
    {
      double gu[2][2] = {{0, 0}, {0, 0}};
      for (int v = 0; v < 4; ++v)
        {
          const int bx = v & 1, by = (v >> 1) & 1;
          const double Nx = bx ? xi[q] : 1.0 - xi[q], Ny = by ? 1.0 : 0.0; // on the edge eta = 1
          const double gx = (bx ? 1.0 : -1.0) / hx * Ny, gy = Nx * (by ? 1.0 : -1.0) / hy;
          const long long n = g.conn[lc * 4 + v];
          for (int c = 0; c < 2; ++c)
            {
              gu[c][0] += gx * sol[n * 3 + c];
              gu[c][1] += gy * sol[n * 3 + c];
            }
        }
      const double tr = gu[0][0] + gu[1][1];
      lx += p.mu * (gu[0][1] + gu[1][0]) * hx * wq[q];
      ly += (p.lambda * tr + 2.0 * p.mu * gu[1][1]) * hx * wq[q];
    }
  atomicAdd (&out2[0], lx);
  atomicAdd (&out2[1], ly);
}

} // namespace pf
