// pf_common.cuh -- shared device-side definitions for the (u,phi) hot path.
//
// Physics restated from tjhei/cracks cracks.cc:2129-2498 (assemble_system).
// Layout: node-major interleaved, NC = dim+1 doubles per node, local slab of
// node planes [plane_begin, plane_end) of the slowest coordinate.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

struct Grid
{
  int dim;
  int n[3];        // global cells per direction
  int nn[3];       // global nodes per direction
  double h[3];
  double origin[3];
  int plane_begin; // first local node plane of the slowest coordinate (incl. ghost)
  int plane_end;   // one past last local node plane
  int owned_begin; // owned node planes [owned_begin, owned_end)
  int owned_end;
  int cell_begin;  // cell layers [cell_begin, cell_end) evaluated by this rank
  int cell_end;
  // tiled 3-D kernels with TZ == 1: tile layer b covers cell layer cell_begin + b * layer_stride.  A stride
  // of cell_end - 1 - cell_begin with two tile layers evaluates the two boundary layers of a slab in one launch.
  int layer_stride = 1;
  long long nodes_per_plane;
  long long n_local_nodes;
  long long n_local_cells;
  long long n_global_nodes;
  // unit_slit.inp topology (2-D Miehe tests, cracks.cc:1202-1205): the nodes on the line
  // y = origin + h * n[1]/2 with x-index >= slit_i0 are doubled; the cell row slit_row
  // (just above the line) uses the copies stored from slit_base on.  slit_row < 0: no slit.
  int slit_row;
  int slit_i0;
  long long slit_base;
  // forest (locally refined) meshes, see pf_create_forest: explicit connectivity, one shape table per
  // refinement level, optional per-cell Lame coefficients.  All null on the box meshes.
  const long long *conn;           // [n_local_cells][2^dim] node numbers
  const unsigned char *cell_level; // [n_local_cells] index into the FeTab array
  const double *cell_lame;         // [n_local_cells][2] = (lambda, mu) or null
  // deterministic mode (pf_set_deterministic): a launch covers only the tiles (or cells) whose index parities in
  // x, y, z are the three bits of `colour`.  Tiles of one colour share no node, the eight colours are launched one
  // after the other: the scatter adds of a launch never meet and every node sees its contributions in the same
  // order in every run.  -1: all tiles in one launch (order of the adds left to the scheduler).
  int colour = -1;
};

// linear block index -> tile index of the tiled 3-D kernels (all tiles, or the tiles of Grid::colour)
__host__ __device__ inline void
decode_tile (const Grid &g, int b, int tiles_x, int tiles_y, int &bx, int &by, int &bz)
{
  if (g.colour < 0)
    {
      bx = b % tiles_x;
      b /= tiles_x;
      by = b % tiles_y;
      bz = b / tiles_y;
      return;
    }
  const int px = g.colour & 1, py = (g.colour >> 1) & 1, pz = g.colour >> 2;
  const int hx = (tiles_x + 1 - px) >> 1, hy = (tiles_y + 1 - py) >> 1;
  bx = 2 * (b % hx) + px;
  b /= hx;
  by = 2 * (b % hy) + py;
  bz = 2 * (b / hy) + pz;
}

// number of tiles a launch with this colour covers
inline long long
tiles_of_colour (int colour, int tiles_x, int tiles_y, int tiles_z)
{
  if (colour < 0)
    return (long long) tiles_x * tiles_y * tiles_z;
  const int px = colour & 1, py = (colour >> 1) & 1, pz = colour >> 2;
  return (long long) ((tiles_x + 1 - px) >> 1) * ((tiles_y + 1 - py) >> 1) * ((tiles_z + 1 - pz) >> 1);
}

// quantities that change per Newton step / time step
struct Phys
{
  double lambda, mu, G_c, kappa, eps;
  double P1;        // (alpha_biot - 1) * pressure   (cracks.cc:2381, 2409, 2428)
  int clamp_extra;  // 1: pf_extra = clamp01(interp(pt)); 0: use_old_timestep_pf (cracks.cc:2276)
  int split;        // Miehe stress split active: decompose_stress_matrix > 0 && timestep_number > 0 (2294, 2338); 2-D
  double d_rhs, d_mat; // decompose_stress_rhs / decompose_stress_matrix (cracks.cc:1568-1569)
};

template <int DIM> struct FeTab
{
  static constexpr int NV = 1 << DIM;
  static constexpr int NQ = DIM == 2 ? 9 : 27;
  double N[NQ][NV];
  double dN[NQ][NV][DIM];
  double JxW[NQ];
};

// mask bits: bit c set <=> component c of the node is constrained
// (Dirichlet rows for c < dim, active set for c == dim)
__host__ __device__ inline bool is_constrained (uint8_t m, int c) { return (m >> c) & 1u; }

#define PF_ACTIVE_BIT(DIM) (1u << (DIM))
// forest meshes: the node is a hanging node (all its components are constrained to its parents)
#define PF_HANGING_BIT 0x80u

} // namespace pf
