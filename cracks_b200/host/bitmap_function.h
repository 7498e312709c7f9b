// bitmap_function.h -- the heterogeneous E-modulus field of `test case = multiple het`.
//
// Restates BitmapFile / BitmapFunction<dim> of the reference (cracks.cc:118-241) including the behaviour its
// goldens depend on: the PGM reader does not skip the max-value token, so "255" becomes the first pixel and
// every pixel is shifted by one (150-155); xi = eta = min(max(., 1), 0) = 0, i.e. the bilinear interpolation
// degenerates to the lower-left pixel (197-198); the 3-D variant mixes three look-ups (227-235).
// Pure host code; the product feeds the resulting per-cell (lambda, mu) to pf_create_forest.
#pragma once
#include <string>
#include <vector>

namespace cracks {

class BitmapFile
{
public:
  explicit BitmapFile (const std::string &name);
  double get_value (double x, double y) const;
  int width () const { return nx; }
  int height () const { return ny; }

private:
  double get_pixel_value (int i, int j) const;
  std::vector<double> image_data;
  double hx = 0, hy = 0;
  int nx = 0, ny = 0;
};

class BitmapFunction
{
public:
  BitmapFunction (const std::string &filename, double x1, double x2, double y1, double y2, double minvalue,
                  double maxvalue)
    : f (filename), x1 (x1), x2 (x2), y1 (y1), y2 (y2), minvalue (minvalue), maxvalue (maxvalue)
  {}
  // p: dim coordinates
  double value (const double *p, int dim) const;

private:
  BitmapFile f;
  double x1, x2, y1, y2, minvalue, maxvalue;
};

} // namespace cracks
