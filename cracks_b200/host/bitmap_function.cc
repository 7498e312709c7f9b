#include "bitmap_function.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cracks {

PgmImage::PgmImage (const std::string &path)
{
  std::ifstream in (path.c_str ());
  if (!in)
    throw std::runtime_error ("cannot read the bitmap <" + path + ">");
  std::string line, token;
  std::getline (in, line);          // "P2"
  in >> token;                      // either the creator comment or already the width
  if (!token.empty () && token[0] == '#')
    {
      std::getline (in, line);
      in >> nx_;
    }
  else
    nx_ = std::atoi (token.c_str ());
  in >> ny_;
  if (nx_ < 2 || ny_ < 2)
    throw std::runtime_error ("bitmap <" + path + ">: bad header");
  // nx * ny tokens follow the size; the first of them is the header's max-value (cracks.cc:150-155)
  grey_.resize ((size_t) nx_ * ny_);
  for (double &g : grey_)
    {
      unsigned v = 0;
      in >> v;
      g = v / 255.0;
    }
}

double
PgmImage::sample (double x, double y) const
{
  const double hx = 1.0 / (nx_ - 1), hy = 1.0 / (ny_ - 1);
  const int i = std::clamp ((int) (x / hx), 0, nx_ - 2);
  const int j = std::clamp ((int) (y / hy), 0, ny_ - 2);
  // the reference clamps the in-pixel offsets to min(max(t, 1), 0) = 0: no interpolation takes place
  const double tx = std::min (std::max ((x - i * hx) / hx, 1.0), 0.0);
  const double ty = std::min (std::max ((y - j * hy) / hy, 1.0), 0.0);
  const double bottom = (1 - tx) * pixel (i, j) + tx * pixel (i + 1, j);
  const double top = (1 - tx) * pixel (i, j + 1) + tx * pixel (i + 1, j + 1);
  return (1 - ty) * bottom + ty * top;
}

double
BitmapFunction::value (const double *point, int dim) const
{
  const double u = (point[0] - x1_) / (x2_ - x1_), v = (point[1] - y1_) / (y2_ - y1_);
  if (dim == 2)
    return lo_ + image_.sample (u, v) * (hi_ - lo_);
  const double w = (point[2] - y1_) / (y2_ - y1_);
  const double mix = image_.sample (u / 10.0, (v - w) / 10.0) + 0.5 * image_.sample ((u + v) / 2.0, (w + u) / 2.0)
                     + 0.25 * image_.sample (std::fmod (w + u - v, 10.0), std::fmod (v + u, 10.0));
  return lo_ + mix * (hi_ - lo_) / 2.25;
}

} // namespace cracks
