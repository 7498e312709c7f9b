// forest_dump.cc -- prints the tables of cracks::Forest for a few refinement recipes as JSON,
// so that tests/test_host_forest.py can compare them with the CPU oracle's forests on the
// reference's golden meshes.  Usage: forest_dump kat2 | kat5 | slit <global> | transfer
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bitmap_function.h"
#include "forest.h"

using namespace cracks;

static void
dump (const Forest &f)
{
  std::printf ("{\"dim\": %d, \"n_cells\": %lld, \"n_nodes\": %lld, \"h_min\": %.17g,\n \"cells\": [", f.dim (),
               f.n_cells (), f.n_nodes (), f.min_cell_diameter ());
  for (long long c = 0; c < f.n_cells (); ++c)
    {
      const ForestCell &x = f.cells ()[(size_t) c];
      std::printf ("%s[%d,%d,%d,%d]", c ? "," : "", x.level, x.idx[0], x.idx[1], x.idx[2]);
    }
  std::printf ("],\n \"conn\": [");
  for (size_t i = 0; i < f.connectivity ().size (); ++i)
    std::printf ("%s%lld", i ? "," : "", f.connectivity ()[i]);
  std::printf ("],\n \"coords\": [");
  for (size_t i = 0; i < f.coordinates ().size (); ++i)
    std::printf ("%s%.17g", i ? "," : "", f.coordinates ()[i]);
  std::printf ("],\n \"hanging\": [");
  for (size_t i = 0; i < f.hanging_nodes ().size (); ++i)
    {
      const HangingNode &h = f.hanging_nodes ()[i];
      std::printf ("%s[%lld", i ? "," : "", h.node);
      for (int q = 0; q < h.n_parents; ++q)
        std::printf (",%lld", h.parents[q]);
      std::printf ("]");
    }
  std::printf ("]}\n");
}

int
main (int argc, char **argv)
{
  if (argc < 2)
    return 2;
  if (!std::strcmp (argv[1], "kat2"))
    {
      // tests/sneddon_2d_1.prm: 10 x 10 box, one `fixed preref sneddon` step (cracks.cc:3901-3923)
      const int n[3] = {10, 10, 1};
      const double lo[3] = {-10, -10, 0}, hi[3] = {10, 10, 0};
      Forest f (2, n, lo, hi);
      std::vector<char> flags ((size_t) f.n_cells (), 0);
      for (long long c = 0; c < f.n_cells (); ++c)
        for (int v = 0; v < 4; ++v)
          {
            const long long nd = f.connectivity ()[(size_t) (c * 4 + v)];
            const double x = f.coordinates ()[(size_t) (nd * 2)], y = f.coordinates ()[(size_t) (nd * 2 + 1)];
            if (x <= 2.5 && x >= -2.5 && y <= 1.25 && y >= -1.25)
              flags[(size_t) c] = 1;
          }
      f.refine (flags);
      dump (f);
    }
  else if (!std::strcmp (argv[1], "kat5"))
    {
      // tests/hetero_3d_1.prm: single-tree cube, 3 global refinements, one `phase field` pre-refinement
      // (threshold 0.4) on the interpolated InitialValuesMultipleHet
      const int n[3] = {1, 1, 1};
      const double lo[3] = {0, 0, 0}, hi[3] = {10, 10, 10};
      Forest f (3, n, lo, hi);
      f.refine_global (3);
      const double h = f.min_cell_diameter ();
      std::vector<char> flags ((size_t) f.n_cells (), 0);
      for (long long c = 0; c < f.n_cells (); ++c)
        for (int v = 0; v < 8; ++v)
          {
            const long long nd = f.connectivity ()[(size_t) (c * 8 + v)];
            if (initial_multiple_het_3d (&f.coordinates ()[(size_t) (nd * 3)], h) < 0.4)
              flags[(size_t) c] = 1;
          }
      f.refine (flags);
      dump (f);
    }
  else if (!std::strcmp (argv[1], "slit") && argc > 2)
    {
      // unit_slit.inp, global refinement, then two rounds of refinement around the crack tip region
      const int n[3] = {2, 2, 1};
      const double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 0};
      Forest f (2, n, lo, hi, true);
      f.refine_global (std::atoi (argv[2]));
      for (const double radius : {0.3, 0.12})
        {
          std::vector<char> flags ((size_t) f.n_cells (), 0);
          for (long long c = 0; c < f.n_cells (); ++c)
            {
              double x[3];
              f.cell_centre (c, x);
              flags[(size_t) c] = std::hypot (x[0] - 0.7, x[1] - 0.45) < radius;
            }
          f.refine (flags);
        }
      dump (f);
    }
  else if (!std::strcmp (argv[1], "transfer"))
    {
      // a Q1 field that is linear in x, y, z must survive SolutionTransfer exactly
      const int n[3] = {2, 3, 2};
      const double lo[3] = {0, -1, 2}, hi[3] = {1, 2, 4};
      Forest coarse (3, n, lo, hi);
      std::vector<char> flags ((size_t) coarse.n_cells (), 0);
      flags[3] = 1;
      coarse.refine (flags);
      Forest fine = coarse;
      std::vector<char> flags2 ((size_t) fine.n_cells (), 0);
      flags2[0] = flags2[5] = flags2[(size_t) fine.n_cells () - 1] = 1;
      fine.refine (flags2);
      auto field = [](const double *x, int c) { return 1.0 + (c + 1) * x[0] - 3.0 * x[1] + 0.5 * c * x[2]; };
      std::vector<double> v0 ((size_t) coarse.n_nodes () * 2), v1 ((size_t) fine.n_nodes () * 2, -1e300);
      for (long long i = 0; i < coarse.n_nodes (); ++i)
        for (int c = 0; c < 2; ++c)
          v0[(size_t) (i * 2 + c)] = field (&coarse.coordinates ()[(size_t) (i * 3)], c);
      fine.transfer (coarse, v0.data (), v1.data (), 2);
      double err = 0;
      for (long long i = 0; i < fine.n_nodes (); ++i)
        for (int c = 0; c < 2; ++c)
          err = std::fmax (err, std::fabs (v1[(size_t) (i * 2 + c)] - field (&fine.coordinates ()[(size_t) (i * 3)], c)));
      std::printf ("{\"coarse_cells\": %lld, \"fine_cells\": %lld, \"fine_hanging\": %zu, \"max_error\": %.3e}\n",
                   coarse.n_cells (), fine.n_cells (), fine.hanging_nodes ().size (), err);
    }
  else if (!std::strcmp (argv[1], "bitmap") && argc > 2)
    {
      // E-modulus field of tests/hetero_3d_1.prm at the cell centres of the KAT-5 mesh: argv[2] = PGM file
      const int n[3] = {1, 1, 1};
      const double lo[3] = {0, 0, 0}, hi[3] = {10, 10, 10};
      Forest f (3, n, lo, hi);
      f.refine_global (3);
      const double h = f.min_cell_diameter ();
      std::vector<char> flags ((size_t) f.n_cells (), 0);
      for (long long c = 0; c < f.n_cells (); ++c)
        for (int v = 0; v < 8; ++v)
          if (initial_multiple_het_3d (&f.coordinates ()[(size_t) (f.connectivity ()[(size_t) (c * 8 + v)] * 3)], h) < 0.4)
            flags[(size_t) c] = 1;
      f.refine (flags);
      const BitmapFunction E (argv[2], 0, 10, 0, 10, 1e4, 1e5); // BitmapFunction(filename,0,10,0,10,E,10 E), cracks.cc:1543
      std::printf ("[");
      for (long long c = 0; c < f.n_cells (); ++c)
        {
          double x[3];
          f.cell_centre (c, x);
          std::printf ("%s%.17g", c ? "," : "", E.value (x, 3));
        }
      std::printf ("]\n");
    }
  else
    return 2;
  return 0;
}
