#include "forest.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>
#include <tuple>

namespace cracks {

Forest::Forest (int dim, const int *n_coarse, const double *lo, const double *hi, bool slit)
  : dim_ (dim), slit_ (slit)
{
  if (dim != 2 && dim != 3)
    throw std::invalid_argument ("Forest: dim must be 2 or 3");
  for (int d = 0; d < 3; ++d)
    {
      n_[d] = d < dim ? n_coarse[d] : 1;
      lo_[d] = d < dim ? lo[d] : 0.0;
      hi_[d] = d < dim ? hi[d] : 1.0;
      if (n_[d] < 1 || !(hi_[d] > lo_[d]))
        throw std::invalid_argument ("Forest: bad box");
    }
  if (slit && (dim != 2 || n_[0] % 2 || n_[1] % 2))
    throw std::invalid_argument ("Forest: the slit topology needs a 2-D mesh with an even number of coarse cells");
  for (int k = 0; k < n_[2]; ++k)
    for (int j = 0; j < n_[1]; ++j)
      for (int i = 0; i < n_[0]; ++i)
        {
          const int idx[3] = {i, j, k};
          set_.insert (key (0, idx));
        }
  build ();
}

uint64_t
Forest::key (int level, const int *idx)
{
  // 6 bits of level, 19 bits per index: enough for 2^19 cells per direction
  return ((uint64_t) level << 57) | ((uint64_t) idx[2] << 38) | ((uint64_t) idx[1] << 19) | (uint64_t) idx[0];
}

bool
Forest::active (int level, const int *idx) const
{
  return set_.count (key (level, idx)) != 0;
}

// the active cell covering position idx of `level`, if it is that cell or a coarser one
bool
Forest::leaf_containing (int level, const int *idx, ForestCell *out) const
{
  int id[3] = {idx[0], idx[1], idx[2]};
  for (int d = 0; d < 3; ++d)
    if (id[d] < 0 || id[d] >= (n_[d] << level))
      return false;
  for (int L = level; L >= 0; --L)
    {
      if (active (L, id))
        {
          out->level = L;
          for (int d = 0; d < 3; ++d)
            out->idx[d] = id[d];
          return true;
        }
      for (int d = 0; d < 3; ++d)
        id[d] >>= 1;
    }
  return false;
}

// is position idx + d a face / edge / corner neighbour of idx in the coarse-mesh connectivity?
// Across the slit only through points with x <= mid: the tip is a shared vertex of all four trees.
bool
Forest::connected (int level, const int *idx, const int *d) const
{
  if (!slit_ || d[1] == 0)
    return true;
  const int half_y = (n_[1] << level) / 2, half_x = (n_[0] << level) / 2;
  const int j2 = idx[1] + d[1];
  if ((idx[1] < half_y) == (j2 < half_y))
    return true;
  if (d[0] == 0)
    return idx[0] <= half_x;
  return (d[0] > 0 ? idx[0] + 1 : idx[0]) <= half_x;
}

void
Forest::split (const ForestCell &c)
{
  if (!active (c.level, c.idx))
    return;
  // 2:1 balance (p4est, full connectivity): no neighbour may be coarser than c before c is split
  for (int dk = (dim_ == 3 ? -1 : 0); dk <= (dim_ == 3 ? 1 : 0); ++dk)
    for (int dj = -1; dj <= 1; ++dj)
      for (int di = -1; di <= 1; ++di)
        {
          if (di == 0 && dj == 0 && dk == 0)
            continue;
          const int d[3] = {di, dj, dk};
          if (!connected (c.level, c.idx, d))
            continue;
          const int nb[3] = {c.idx[0] + di, c.idx[1] + dj, c.idx[2] + dk};
          ForestCell leaf;
          if (leaf_containing (c.level, nb, &leaf) && leaf.level < c.level)
            split (leaf);
        }
  set_.erase (key (c.level, c.idx));
  for (int v = 0; v < (1 << dim_); ++v)
    {
      const int child[3] = {2 * c.idx[0] + (v & 1), 2 * c.idx[1] + ((v >> 1) & 1),
                            dim_ == 3 ? 2 * c.idx[2] + ((v >> 2) & 1) : 0};
      set_.insert (key (c.level + 1, child));
    }
}

void
Forest::refine_global (int times)
{
  for (int t = 0; t < times; ++t)
    refine (std::vector<char> (cells_.size (), 1));
}

void
Forest::refine (const std::vector<char> &flags)
{
  if (flags.size () != cells_.size ())
    throw std::invalid_argument ("Forest::refine: one flag per active cell");
  const std::vector<ForestCell> before = cells_;
  for (size_t c = 0; c < before.size (); ++c)
    if (flags[c])
      split (before[c]);
  build ();
}

void
Forest::cell_size (int level, double *h) const
{
  for (int d = 0; d < dim_; ++d)
    h[d] = (hi_[d] - lo_[d]) / (double) (n_[d] << level);
}

void
Forest::cell_centre (long long cell, double *x) const
{
  const ForestCell &c = cells_[(size_t) cell];
  double h[3];
  cell_size (c.level, h);
  for (int d = 0; d < dim_; ++d)
    x[d] = lo_[d] + (c.idx[d] + 0.5) * h[d];
}

double
Forest::min_cell_diameter () const
{
  double h[3], s = 0;
  cell_size (max_level_, h);
  for (int d = 0; d < dim_; ++d)
    s += h[d] * h[d];
  return std::sqrt (s);
}

void
Forest::build ()
{
  cells_.clear ();
  max_level_ = 0;
  for (const uint64_t k : set_)
    {
      ForestCell c;
      c.level = (int) (k >> 57);
      c.idx[0] = (int) (k & 0x7FFFF);
      c.idx[1] = (int) ((k >> 19) & 0x7FFFF);
      c.idx[2] = (int) ((k >> 38) & 0x7FFFF);
      cells_.push_back (c);
      max_level_ = std::max (max_level_, c.level);
    }
  const int Lm = max_level_;
  // deterministic order: by lower corner on the finest lattice (z, y, x), then level
  std::sort (cells_.begin (), cells_.end (), [Lm](const ForestCell &a, const ForestCell &b) {
    const int sa = Lm - a.level, sb = Lm - b.level;
    return std::make_tuple (a.idx[2] << sa, a.idx[1] << sa, a.idx[0] << sa, a.level)
           < std::make_tuple (b.idx[2] << sb, b.idx[1] << sb, b.idx[0] << sb, b.level);
  });

  using NodeKey = std::array<long long, 4>; // lattice x, y, z, slit side
  std::map<NodeKey, long long> node_of;
  coords_.clear ();
  conn_.clear ();
  upper_copy_.clear ();
  double h0[3];
  cell_size (Lm, h0);
  const long long xm = ((long long) n_[0] << Lm) / 2, ym = ((long long) n_[1] << Lm) / 2;
  auto node_key = [&](const long long *p, bool upper) {
    NodeKey k = {p[0], p[1], p[2], (slit_ && upper && p[1] == ym && p[0] > xm) ? 1 : 0};
    return k;
  };
  auto node = [&](const NodeKey &k) {
    auto it = node_of.find (k);
    if (it != node_of.end ())
      return it->second;
    const long long id = (long long) upper_copy_.size ();
    node_of.emplace (k, id);
    for (int d = 0; d < dim_; ++d)
      coords_.push_back (lo_[d] + (double) k[(size_t) d] * h0[d]);
    upper_copy_.push_back ((char) k[3]);
    return id;
  };
  const int nv = 1 << dim_;
  std::vector<std::array<NodeKey, 8>> lattice (cells_.size ());
  std::vector<char> upper (cells_.size (), 0);
  for (size_t c = 0; c < cells_.size (); ++c)
    {
      const ForestCell &cell = cells_[c];
      const long long s = 1ll << (Lm - cell.level);
      upper[c] = slit_ && (long long) cell.idx[1] * s >= ym;
      for (int v = 0; v < nv; ++v)
        {
          const long long p[3] = {(cell.idx[0] + (v & 1)) * s, (cell.idx[1] + ((v >> 1) & 1)) * s,
                                  dim_ == 3 ? (cell.idx[2] + ((v >> 2) & 1)) * s : 0};
          lattice[c][(size_t) v] = node_key (p, upper[c]);
          conn_.push_back (node (lattice[c][(size_t) v]));
        }
    }
  // hanging nodes: the midpoint of an edge (centre of a face) of an active cell that is a vertex of
  // the refined neighbour is constrained to the mean of the edge ends (face corners)
  std::map<long long, HangingNode> hang;
  std::vector<std::vector<int>> groups;
  for (int d = 0; d < dim_; ++d)
    for (int a = 0; a < nv; ++a)
      if (!(a & (1 << d)))
        groups.push_back ({a, a | (1 << d)});
  if (dim_ == 3)
    for (int d = 0; d < 3; ++d)
      for (int side = 0; side < 2; ++side)
        {
          std::vector<int> f;
          for (int a = 0; a < 8; ++a)
            if (((a >> d) & 1) == side)
              f.push_back (a);
          groups.push_back (f);
        }
  for (size_t c = 0; c < cells_.size (); ++c)
    for (const auto &grp : groups)
      {
        long long sum[3] = {0, 0, 0};
        for (int a : grp)
          for (int d = 0; d < 3; ++d)
            sum[d] += lattice[c][(size_t) a][(size_t) d];
        const long long np = (long long) grp.size ();
        if (sum[0] % np || sum[1] % np || sum[2] % np)
          continue;
        const long long mid[3] = {sum[0] / np, sum[1] / np, sum[2] / np};
        auto it = node_of.find (node_key (mid, upper[c]));
        if (it == node_of.end ())
          continue;
        HangingNode hnode;
        hnode.node = it->second;
        hnode.n_parents = (int) np;
        for (int q = 0; q < 4; ++q)
          hnode.parents[q] = q < np ? node_of.at (lattice[c][(size_t) grp[(size_t) q]]) : -1;
        hang[hnode.node] = hnode;
      }
  hanging_.clear ();
  for (const auto &kv : hang)
    hanging_.push_back (kv.second);
}

void
Forest::transfer (const Forest &from, const double *v_from, double *v_to, int ncomp) const
{
  std::map<uint64_t, long long> old_index;
  for (size_t c = 0; c < from.cells_.size (); ++c)
    old_index[key (from.cells_[c].level, from.cells_[c].idx)] = (long long) c;
  const int nv = 1 << dim_;
  std::vector<char> done ((size_t) n_nodes (), 0);
  for (size_t c = 0; c < cells_.size (); ++c)
    {
      const ForestCell &cell = cells_[c];
      int L = cell.level, id[3] = {cell.idx[0], cell.idx[1], cell.idx[2]};
      std::map<uint64_t, long long>::const_iterator it;
      while ((it = old_index.find (key (L, id))) == old_index.end ())
        {
          if (L == 0)
            throw std::runtime_error ("Forest::transfer: the target is not a refinement of the source");
          --L;
          for (int d = 0; d < 3; ++d)
            id[d] >>= 1;
        }
      const long long oc = it->second;
      const int s = 1 << (cell.level - L);
      for (int v = 0; v < nv; ++v)
        {
          const long long n = conn_[c * (size_t) nv + (size_t) v];
          if (done[(size_t) n])
            continue;
          double xi[3] = {0, 0, 0};
          for (int d = 0; d < dim_; ++d)
            xi[d] = (double) ((cell.idx[d] - id[d] * s) + ((v >> d) & 1)) / (double) s;
          for (int q = 0; q < ncomp; ++q)
            v_to[n * ncomp + q] = 0.0;
          for (int w = 0; w < nv; ++w)
            {
              double weight = 1.0;
              for (int d = 0; d < dim_; ++d)
                weight *= ((w >> d) & 1) ? xi[d] : 1.0 - xi[d];
              if (weight == 0.0)
                continue;
              const long long on = from.conn_[(size_t) oc * (size_t) nv + (size_t) w];
              for (int q = 0; q < ncomp; ++q)
                v_to[n * ncomp + q] += weight * v_from[on * ncomp + q];
            }
          done[(size_t) n] = 1;
        }
    }
}

double
initial_multiple_het_3d (const double *p, double min_cell_diameter)
{
  const double w = min_cell_diameter;
  if (((p[0] >= 2.6 - w / 2.0) && (p[0] <= 2.6 + w / 2.0)) && ((p[1] >= 3.8 - w / 2.0) && (p[1] <= 5.5 + w / 2.0))
      && (p[2] >= 4.0 - w / 2.0) && (p[2] <= 4.0 + w / 2.0))
    return 0.0;
  if (((p[0] >= 5.5 - w / 2.0) && (p[0] <= 7.0 + w / 2.0)) && ((p[1] >= 4.0 - w / 2.0) && (p[1] <= 4.0 + w / 2.0))
      && (p[2] >= 6.0 - w / 2.0) && (p[2] <= 6.0 + w / 2.0))
    return 0.0;
  return 1.0;
}

} // namespace cracks

// ---- C binding (ctypes mirror: cracks_b200/forest.py) ---------------------------------------------
extern "C" {

void *
pfh_forest_create (int dim, const int *n, const double *lo, const double *hi, int slit)
{
  try
    {
      return new cracks::Forest (dim, n, lo, hi, slit != 0);
    }
  catch (std::exception &)
    {
      return nullptr;
    }
}

void *
pfh_forest_clone (const void *f)
{
  return new cracks::Forest (*static_cast<const cracks::Forest *> (f));
}

void
pfh_forest_destroy (void *f)
{
  delete static_cast<cracks::Forest *> (f);
}

void
pfh_forest_refine_global (void *f, int times)
{
  static_cast<cracks::Forest *> (f)->refine_global (times);
}

int
pfh_forest_refine (void *f, const unsigned char *flags, long long n)
{
  cracks::Forest *F = static_cast<cracks::Forest *> (f);
  if (n != F->n_cells ())
    return -1;
  F->refine (std::vector<char> (flags, flags + n));
  return 0;
}

long long
pfh_forest_n_cells (const void *f)
{
  return static_cast<const cracks::Forest *> (f)->n_cells ();
}

long long
pfh_forest_n_nodes (const void *f)
{
  return static_cast<const cracks::Forest *> (f)->n_nodes ();
}

long long
pfh_forest_n_hanging (const void *f)
{
  return (long long) static_cast<const cracks::Forest *> (f)->hanging_nodes ().size ();
}

int
pfh_forest_max_level (const void *f)
{
  return static_cast<const cracks::Forest *> (f)->max_level ();
}

double
pfh_forest_min_cell_diameter (const void *f)
{
  return static_cast<const cracks::Forest *> (f)->min_cell_diameter ();
}

// conn [n_cells][2^dim], level [n_cells], coords [n_nodes][dim], hanging [n_hanging][5] (node, parents,
// -1 = unused), level_h [(max_level + 1)][dim], upper_copy [n_nodes] (doubled slit nodes); any may be null
void
pfh_forest_tables (const void *f, long long *conn, unsigned char *level, double *coords, long long *hanging,
                   double *level_h, unsigned char *upper_copy)
{
  const cracks::Forest &F = *static_cast<const cracks::Forest *> (f);
  if (conn)
    std::copy (F.connectivity ().begin (), F.connectivity ().end (), conn);
  if (level)
    for (long long c = 0; c < F.n_cells (); ++c)
      level[c] = (unsigned char) F.cells ()[(size_t) c].level;
  if (coords)
    std::copy (F.coordinates ().begin (), F.coordinates ().end (), coords);
  if (hanging)
    for (size_t h = 0; h < F.hanging_nodes ().size (); ++h)
      {
        const cracks::HangingNode &hn = F.hanging_nodes ()[h];
        hanging[5 * h] = hn.node;
        for (int q = 0; q < 4; ++q)
          hanging[5 * h + 1 + q] = q < hn.n_parents ? hn.parents[q] : -1;
      }
  if (level_h)
    for (int l = 0; l <= F.max_level (); ++l)
      F.cell_size (l, level_h + l * F.dim ());
  if (upper_copy)
    for (long long n = 0; n < F.n_nodes (); ++n)
      upper_copy[n] = F.is_upper_slit_copy (n) ? 1 : 0;
}

int
pfh_forest_transfer (const void *to, const void *from, const double *v_from, double *v_to, int ncomp)
{
  try
    {
      static_cast<const cracks::Forest *> (to)->transfer (*static_cast<const cracks::Forest *> (from), v_from, v_to, ncomp);
      return 0;
    }
  catch (std::exception &)
    {
      return -1;
    }
}

} // extern "C"
