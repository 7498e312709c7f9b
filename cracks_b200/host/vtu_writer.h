// vtu_writer.h -- on-disk output of output_results() (cracks.cc:3142-3258): one VTK XML UnstructuredGrid
// piece per call plus the .pvtu / .visit / .pvd records that index the pieces.  deal.II's DataOut writes
// zlib-compressed base64; here the arrays are raw appended binary (valid VTK XML, no dependency) -- the files
// are for ParaView / VisIt, no golden pins their bytes.  Pure host code.
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace cracks {

struct VtuPointField
{
  std::string name;
  int n_components; // 1 = scalar, 3 = vector (2-D vectors are padded with a zero z component)
  std::vector<double> values; // [n_points][n_components]
};

struct VtuCellField
{
  std::string name;
  std::vector<float> values; // [n_cells]
};

// cells: 2^dim vertices per cell in lexicographic order (x fastest), as the kernels use them
void write_vtu (const std::string &path, int dim, long long n_points, const double *coordinates, long long n_cells,
                const long long *connectivity, const std::vector<VtuPointField> &point_fields,
                const std::vector<VtuCellField> &cell_fields);

// DataOutInterface::write_pvtu_record / write_visit_record / write_pvd_record
void write_pvtu_record (const std::string &path, const std::vector<std::string> &piece_names,
                        const std::vector<VtuPointField> &point_fields, const std::vector<VtuCellField> &cell_fields);
void write_visit_record (const std::string &path, const std::vector<std::vector<std::string>> &pieces_by_timestep);
void write_pvd_record (const std::string &path, const std::vector<std::pair<double, std::string>> &times_and_names);

} // namespace cracks
