// emu_kernels.cc -- TEST INFRASTRUCTURE: runs the *actual kernel sources* of cracks_b200/csrc
// (pf_generic.cuh, pf_forest.cuh, pf_vector.cuh, pf_split2d.cuh) on the CPU, one emulated thread at a
// time, in the sequence pf_api.cu launches them, so that the kernel logic can be held against the
// oracle without a GPU (tests/test_kernel_emulation_cpu.py).  Compiled with -I tests/emu/cuda_shim so
// that <cuda_runtime.h> resolves to the shim.  This is not a CPU fallback of the product: nothing
// under cracks_b200/ links it.
#include <cuda_runtime.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

#include <algorithm>
#include <vector>

#include "../../cracks_b200/csrc/pf_common.cuh"
#include "../../cracks_b200/csrc/pf_forest.cuh"
#include "../../cracks_b200/csrc/pf_generic.cuh"
#include "../../cracks_b200/csrc/pf_vector.cuh"

using namespace pf;

template <class K, class... A>
static void
launch (K kernel, long long n_threads, int block, A... args)
{
  const long long grid = (n_threads + block - 1) / block;
  gridDim.x = (unsigned) grid;
  blockDim.x = (unsigned) block;
  for (long long b = 0; b < grid; ++b)
    for (int t = 0; t < block; ++t)
      {
        blockIdx.x = (unsigned) b;
        threadIdx.x = (unsigned) t;
        kernel (args...);
      }
}

template <int DIM>
static void
fill_tab (FeTab<DIM> &t, const double *h)
{
  // same table as fill_fetab in pf_api.cu
  const double gq = 0.5 * std::sqrt (3.0 / 5.0);
  const double xi[3] = {0.5 - gq, 0.5, 0.5 + gq};
  const double w[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  double vol = 1;
  for (int d = 0; d < DIM; ++d)
    vol *= h[d];
  for (int q = 0; q < FeTab<DIM>::NQ; ++q)
    {
      const int qi[3] = {q % 3, (q / 3) % 3, q / 9};
      double wq = vol;
      for (int d = 0; d < DIM; ++d)
        wq *= w[qi[d]];
      t.JxW[q] = wq;
      for (int v = 0; v < (1 << DIM); ++v)
        {
          double val = 1;
          for (int d = 0; d < DIM; ++d)
            val *= ((v >> d) & 1) ? xi[qi[d]] : 1.0 - xi[qi[d]];
          t.N[q][v] = val;
          for (int e = 0; e < DIM; ++e)
            {
              double gr = 1;
              for (int d = 0; d < DIM; ++d)
                {
                  const int b = (v >> d) & 1;
                  gr *= (d == e) ? (b ? 1.0 : -1.0) / h[d] : (b ? xi[qi[d]] : 1.0 - xi[qi[d]]);
                }
              t.dN[q][v][e] = gr;
            }
        }
    }
}

struct EmuForest
{
  int dim;
  long long n_cells, n_nodes, n_hanging;
  const long long *conn;
  const unsigned char *level;
  int n_levels;
  const double *level_h;
  const long long *hanging; // [n_hanging][5]
  const double *cell_lame;  // or null
};

template <int DIM>
static Grid
make_grid (const EmuForest &f)
{
  Grid g{};
  g.dim = DIM;
  g.n[0] = (int) f.n_cells;
  g.n[1] = g.n[2] = 1;
  g.plane_end = g.owned_end = g.cell_end = 1;
  g.nodes_per_plane = g.n_local_nodes = g.n_global_nodes = f.n_nodes;
  g.n_local_cells = f.n_cells;
  g.slit_row = -1;
  g.conn = f.conn;
  g.cell_level = f.level;
  g.cell_lame = f.cell_lame;
  return g;
}

// the launch sequences of pf_api.cu on a forest mesh: residual_dev, diag_and_aux, apply_forest_dev
template <int DIM>
static void
run (const EmuForest &f, const Phys &p, const double *sol, const double *pt, const uint8_t *mask_in, const double *x,
     double *r_total, double *r_pde, double *diag, double *y, double *mass)
{
  constexpr int NC = DIM + 1;
  const long long nn = f.n_nodes, nd = nn * NC;
  Grid g = make_grid<DIM> (f);
  std::vector<FeTab<DIM>> tabs ((size_t) f.n_levels);
  for (int l = 0; l < f.n_levels; ++l)
    fill_tab<DIM> (tabs[(size_t) l], f.level_h + l * DIM);
  std::vector<uint8_t> mask (mask_in, mask_in + nn), zero_mask ((size_t) nn, 0);
  launch (k_mark_hanging, f.n_hanging, 256, f.n_hanging, f.hanging, mask.data ());
  // create_forest_impl: lumped mass
  std::fill (mass, mass + nn, 0.0);
  launch (k_lumped_mass_forest<DIM>, f.n_cells, 128, g, tabs.data (), mass);
  // residual_dev
  std::fill (r_total, r_total + nd, 0.0);
  launch (k_residual_generic<DIM>, f.n_cells, 128, g, p, tabs.data (), sol, pt, r_total);
  launch (k_hanging_fold<NC>, f.n_hanging * NC, 256, f.n_hanging, f.hanging, zero_mask.data (), (const double *) nullptr,
          (const double *) nullptr, r_total);
  for (long long i = 0; i < nd; ++i) // k_residual_finish without its block reduction
    r_pde[i] = is_constrained (mask[(size_t) (i / NC)], (int) (i % NC)) ? 0.0 : r_total[i];
  // diag_and_aux
  std::fill (diag, diag + nd, 0.0);
  launch (k_diag_generic<DIM>, f.n_cells, 128, g, p, tabs.data (), sol, pt, diag);
  launch (k_hanging_fold_diag<NC>, f.n_hanging * NC, 256, f.n_hanging, f.hanging, diag);
  // apply_forest_dev
  std::vector<double> fx (x, x + nd);
  launch (k_hanging_distribute<NC>, f.n_hanging * NC, 256, f.n_hanging, f.hanging, mask.data (), 1, fx.data ());
  launch (k_apply_init<DIM>, nn, 256, nn, x, diag, mask.data (), y);
  launch (k_apply_generic<DIM>, f.n_cells, 128, g, p, tabs.data (), fx.data (), sol, pt, mask.data (), y);
  launch (k_hanging_fold<NC>, f.n_hanging * NC, 256, f.n_hanging, f.hanging, mask.data (), diag, x, y);
}

extern "C" {

// sol, x, outputs: node-major interleaved (dim + 1 doubles per node); pt: nodal phi~ (unclamped);
// mask: one byte per node, bit c = component c constrained
void
emu_forest (int dim, long long n_cells, long long n_nodes, const long long *conn, const unsigned char *level, int n_levels,
            const double *level_h, long long n_hanging, const long long *hanging, const double *cell_lame,
            const double *phys /* lambda, mu, G_c, kappa, eps, P1, clamp_extra */, const double *sol, const double *pt,
            const unsigned char *mask, const double *x, double *r_total, double *r_pde, double *diag, double *y,
            double *mass)
{
  EmuForest f{dim, n_cells, n_nodes, n_hanging, conn, level, n_levels, level_h, hanging, cell_lame};
  Phys p;
  std::memset (&p, 0, sizeof p);
  p.lambda = phys[0], p.mu = phys[1], p.G_c = phys[2], p.kappa = phys[3], p.eps = phys[4], p.P1 = phys[5];
  p.clamp_extra = (int) phys[6];
  if (dim == 2)
    run<2> (f, p, sol, pt, mask, x, r_total, r_pde, diag, y, mass);
  else
    run<3> (f, p, sol, pt, mask, x, r_total, r_pde, diag, y, mass);
}

// the generic kernels on a 2-D box / slit mesh (implicit connectivity), with or without the Miehe split,
// in the order residual_dev / diag_and_aux / apply_dev launch them
void
emu_box2d (const int *n, const double *h, int slit, const double *phys /* lambda, mu, G_c, kappa, eps, P1, clamp, split,
           d_rhs, d_mat */, const double *sol, const double *pt, const unsigned char *mask, const double *x,
           double *r_total, double *diag, double *y, double *mass)
{
  Grid g{};
  g.dim = 2;
  g.nodes_per_plane = 1;
  g.n_global_nodes = 1;
  for (int d = 0; d < 3; ++d)
    {
      g.n[d] = d < 2 ? n[d] : 1;
      g.nn[d] = d < 2 ? n[d] + 1 : 1;
      g.h[d] = d < 2 ? h[d] : 1.0;
      g.n_global_nodes *= g.nn[d];
    }
  g.nodes_per_plane = g.nn[0];
  g.plane_end = g.owned_end = g.nn[1];
  g.cell_end = n[1];
  g.n_local_nodes = g.n_global_nodes;
  g.n_local_cells = (long long) n[0] * n[1];
  g.slit_row = -1;
  if (slit)
    {
      g.slit_row = n[1] / 2;
      g.slit_i0 = n[0] / 2 + 1;
      g.slit_base = g.n_local_nodes;
      g.n_local_nodes += n[0] / 2;
      g.n_global_nodes += n[0] / 2;
    }
  Phys p;
  std::memset (&p, 0, sizeof p);
  p.lambda = phys[0], p.mu = phys[1], p.G_c = phys[2], p.kappa = phys[3], p.eps = phys[4], p.P1 = phys[5];
  p.clamp_extra = (int) phys[6];
  p.split = (int) phys[7];
  p.d_rhs = phys[8];
  p.d_mat = phys[9];
  FeTab<2> tab;
  fill_tab<2> (tab, h);
  const long long nn = g.n_local_nodes, nd = nn * 3;
  std::fill (mass, mass + nn, 0.0);
  launch (k_lumped_mass_cells<2>, g.n_local_cells, 128, g, mass);
  std::fill (r_total, r_total + nd, 0.0);
  launch (k_residual_generic<2>, g.n_local_cells, 128, g, p, (const FeTab<2> *) &tab, sol, pt, r_total);
  std::fill (diag, diag + nd, 0.0);
  launch (k_diag_generic<2>, g.n_local_cells, 128, g, p, (const FeTab<2> *) &tab, sol, pt, diag);
  launch (k_apply_init<2>, nn, 256, nn, x, (const double *) diag, mask, y);
  launch (k_apply_generic<2>, g.n_local_cells, 128, g, p, (const FeTab<2> *) &tab, x, sol, pt, mask, y);
}

// compute_load on a 2-D forest (k_load_top_forest) and the Dirichlet-value kernel
void
emu_load_forest (long long n_cells, long long n_nodes, const long long *conn, const unsigned char *level,
                 const double *level_h, double lambda, double mu, long long n_list, const long long *list,
                 const double *sol, double *out2)
{
  EmuForest f{2, n_cells, n_nodes, 0, conn, level, 0, level_h, nullptr, nullptr};
  Grid g = make_grid<2> (f);
  Phys p;
  std::memset (&p, 0, sizeof p);
  p.lambda = lambda;
  p.mu = mu;
  out2[0] = out2[1] = 0.0;
  launch (k_load_top_forest, n_list, 128, g, p, level_h, n_list, list, sol, out2);
}

void
emu_set_dirichlet_values (int dim, long long n_nodes, const unsigned char *mask, const double *vals, double *sol)
{
  if (dim == 2)
    launch (k_set_dirichlet_values<2>, n_nodes, 256, n_nodes, mask, vals, sol);
  else
    launch (k_set_dirichlet_values<3>, n_nodes, 256, n_nodes, mask, vals, sol);
}

// active-set kernel on a forest: returns counts {active, cycling, changed}
void
emu_active_set (int dim, long long n_nodes, double c_scale, const double *r_total, const double *mass, const double *old,
                double *sol, int *cycle, unsigned char *mask, unsigned long long *counts)
{
  counts[0] = counts[1] = counts[2] = 0;
  if (dim == 2)
    launch (k_active_set<2>, n_nodes, 256, n_nodes, 0ll, n_nodes, c_scale, r_total, mass, old, sol, cycle, mask, counts);
  else
    launch (k_active_set<3>, n_nodes, 256, n_nodes, 0ll, n_nodes, c_scale, r_total, mass, old, sol, cycle, mask, counts);
}
}
