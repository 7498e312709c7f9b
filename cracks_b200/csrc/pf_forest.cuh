// pf_forest.cuh -- kernels for locally refined (forest) meshes with hanging nodes.
//
// EXPERIMENTAL: written against the CPU oracle's formulation (oracle/adaptive_oracle.py, which
// reproduces the reference's sneddon_2d_1 / miehe_shear_1 / hetero_3d_1 goldens) but not yet run
// on a GPU; only reachable through pf_create_forest.  The box-mesh paths do not use this file.
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

} // namespace pf
