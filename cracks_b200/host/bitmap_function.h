// bitmap_function.h -- the heterogeneous E-modulus field of `test case = multiple het`.
//
// What the reference computes in BitmapFile / BitmapFunction<dim> (cracks.cc:118-241), including the
// behaviour its goldens depend on:
//   * the PGM reader does not skip the max-value token of the header, so "255" becomes pixel 0 and every
//     pixel is shifted by one (150-155);
//   * the fractional position inside a pixel is clamped with min(max(., 1), 0), i.e. it is always 0 and the
//     "bilinear interpolation" returns the lower-left pixel (197-198);
//   * rows are stored top-down, row j of the lookup is row ny-1-j of the file (180);
//   * the 3-D field mixes three look-ups with weights 1, 1/2, 1/4 and divides by 2.25 (227-235).
// Pure host code; the product turns the field into per-cell (lambda, mu) for pf_create_forest.
#pragma once
#include <string>
#include <vector>

namespace cracks {

// grey values in [0, 1] on the unit square
class PgmImage
{
public:
  explicit PgmImage (const std::string &path);
  double sample (double x, double y) const;
  int width () const { return nx_; }
  int height () const { return ny_; }

private:
  double pixel (int i, int j) const { return grey_[(size_t) (nx_ * (ny_ - 1 - j) + i)]; }
  std::vector<double> grey_;
  int nx_ = 0, ny_ = 0;
};

// value = lo + image(x, y) * (hi - lo) on [x1,x2] x [y1,y2] (2-D); the reference's three-look-up mix in 3-D
class BitmapFunction
{
public:
  BitmapFunction (const std::string &path, double x1, double x2, double y1, double y2, double lo, double hi)
    : image_ (path), x1_ (x1), x2_ (x2), y1_ (y1), y2_ (y2), lo_ (lo), hi_ (hi)
  {}
  double value (const double *point, int dim) const;

private:
  PgmImage image_;
  double x1_, x2_, y1_, y2_, lo_, hi_;
};

} // namespace cracks
