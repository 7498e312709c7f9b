#include "bitmap_function.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <stdexcept>

namespace cracks {

BitmapFile::BitmapFile (const std::string &name)
{
  std::ifstream in (name.c_str ());
  if (!in)
    throw std::runtime_error ("Can't read from file <" + name + ">!");
  std::string temp;
  std::getline (in, temp); // magic number
  in >> temp;
  if (!temp.empty () && temp[0] == '#')
    std::getline (in, temp); // comment line; otherwise the token just read was nx -- not handled by the reference either
  in >> nx >> ny;
  if (!(nx > 0 && ny > 0))
    throw std::runtime_error ("Invalid file format.");
  image_data.reserve ((size_t) nx * ny);
  for (int k = 0; k < nx * ny; ++k)
    {
      unsigned int val = 0;
      in >> val; // the first value read is the max-value token of the header, like the reference
      image_data.push_back (val / 255.0);
    }
  hx = 1.0 / (nx - 1);
  hy = 1.0 / (ny - 1);
}

double
BitmapFile::get_pixel_value (int i, int j) const
{
  return image_data[(size_t) (nx * (ny - 1 - j) + i)];
}

double
BitmapFile::get_value (double x, double y) const
{
  const int ix = std::min (std::max ((int) (x / hx), 0), nx - 2);
  const int iy = std::min (std::max ((int) (y / hy), 0), ny - 2);
  const double xi = std::min (std::max ((x - ix * hx) / hx, 1.), 0.);
  const double eta = std::min (std::max ((y - iy * hy) / hy, 1.), 0.);
  return ((1 - xi) * (1 - eta) * get_pixel_value (ix, iy) + xi * (1 - eta) * get_pixel_value (ix + 1, iy)
          + (1 - xi) * eta * get_pixel_value (ix, iy + 1) + xi * eta * get_pixel_value (ix + 1, iy + 1));
}

double
BitmapFunction::value (const double *p, int dim) const
{
  const double x = (p[0] - x1) / (x2 - x1);
  const double y = (p[1] - y1) / (y2 - y1);
  if (dim == 2)
    return minvalue + f.get_value (x, y) * (maxvalue - minvalue);
  const double z = (p[2] - y1) / (y2 - y1);
  return minvalue
         + (f.get_value (x / 10.0, (y - z) / 10.0) + 0.5 * f.get_value ((x + y) / 2.0, (z + x) / 2.0)
            + 0.25 * f.get_value (std::fmod (z + x - y, 10.0), std::fmod (y + x, 10.0)))
             * (maxvalue - minvalue) / 2.25;
}

} // namespace cracks
