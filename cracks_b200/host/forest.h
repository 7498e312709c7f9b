// forest.h -- host-side mesh forest for locally refined box meshes.
//
// The reference keeps its mesh in a parallel::distributed::Triangulation (p4est) and
// refines it in refine_mesh() (cracks.cc:3895-4163): flag cells, keep the level
// difference between neighbours (faces, edges, corners) at one, split, interpolate
// the three solution vectors to the new mesh (SolutionTransfer, 4137-4159) and
// rebuild the hanging-node constraints (make_hanging_node_constraints, 1630-1634).
// This class is that data structure for the meshes the reference's test cases
// start from: the box of subdivided_hyper_rectangle (cracks.cc:1248-1253), the
// single-tree cube of meshes/unit_cube_10.inp and the 2 x 2 square with a slit of
// meshes/unit_slit.inp.  It produces the flat tables a GPU context consumes:
// cell -> node connectivity, cell levels, node coordinates and the hanging-node
// table (node, parents); all cells are axis-aligned, so a level fixes the cell size.
//
// Pure host code, no CUDA.  Checked against the CPU oracle's forests on the
// reference's golden meshes (tests/test_host_forest.py).
#pragma once
#include <array>
#include <cstdint>
#include <unordered_set>
#include <vector>

namespace cracks {

struct ForestCell
{
  int level;
  int idx[3]; // position in cells of its own level
  bool operator== (const ForestCell &o) const
  {
    return level == o.level && idx[0] == o.idx[0] && idx[1] == o.idx[1] && idx[2] == o.idx[2];
  }
};

struct HangingNode
{
  long long node;
  int n_parents;        // 2: edge midpoint, 4: face centre (3-D)
  long long parents[4]; // weights 1 / n_parents
};

class Forest
{
public:
  // slit (2-D): the topology of unit_slit.inp, see pf_mesh.slit in include/cracks_b200.h
  Forest (int dim, const int *n_coarse, const double *lo, const double *hi, bool slit = false);

  void refine_global (int times);
  // flags: one entry per active cell in the order of cells(); flagged cells are split
  // once, coarser neighbours are split first where the 2:1 rule demands it
  void refine (const std::vector<char> &flags);
  // (re)number nodes and build the tables below; called by the constructor and by refine*
  void build ();

  int dim () const { return dim_; }
  int vertices_per_cell () const { return 1 << dim_; }
  long long n_cells () const { return (long long) cells_.size (); }
  long long n_nodes () const { return (long long) (coords_.size () / dim_); }
  int max_level () const { return max_level_; }
  const std::vector<ForestCell> &cells () const { return cells_; }
  const std::vector<long long> &connectivity () const { return conn_; } // [n_cells][2^dim], x fastest
  const std::vector<double> &coordinates () const { return coords_; }  // [n_nodes][dim]
  const std::vector<HangingNode> &hanging_nodes () const { return hanging_; }
  void cell_size (int level, double *h) const;
  void cell_centre (long long cell, double *x) const;
  double min_cell_diameter () const; // cell->diameter() minimum, cracks.cc:3823-3835
  // doubled node of the slit (upper side)?
  bool is_upper_slit_copy (long long node) const { return upper_copy_[node] != 0; }

  // SolutionTransfer::interpolate: nodal vectors with ncomp interleaved components
  // from the forest `from` (a coarser state of this one) to this forest
  void transfer (const Forest &from, const double *v_from, double *v_to, int ncomp) const;

private:
  static uint64_t key (int level, const int *idx);
  bool active (int level, const int *idx) const;
  bool leaf_containing (int level, const int *idx, ForestCell *out) const;
  bool connected (int level, const int *idx, const int *d) const;
  void split (const ForestCell &c);

  int dim_;
  int n_[3];
  double lo_[3], hi_[3];
  bool slit_;
  std::unordered_set<uint64_t> set_;
  std::vector<ForestCell> cells_;
  std::vector<long long> conn_;
  std::vector<double> coords_;
  std::vector<HangingNode> hanging_;
  std::vector<char> upper_copy_;
  int max_level_ = 0;
};

// InitialValuesMultipleHet<3>::value for the phase-field component (cracks.cc:595-610)
double initial_multiple_het_3d (const double *x, double min_cell_diameter);

} // namespace cracks
