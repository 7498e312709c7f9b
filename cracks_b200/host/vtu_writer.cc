#include "vtu_writer.h"

#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cracks {
namespace {

struct Appended
{
  std::string bytes;
  // returns the offset of the block (counted from the byte after the underscore)
  template <class T>
  uint64_t add (const std::vector<T> &v)
  {
    const uint64_t offset = bytes.size (), n = v.size () * sizeof (T);
    bytes.append (reinterpret_cast<const char *> (&n), sizeof n);
    bytes.append (reinterpret_cast<const char *> (v.data ()), (size_t) n);
    return offset;
  }
};

std::string
array_tag (const char *type, const std::string &name, int ncomp, uint64_t offset)
{
  std::ostringstream o;
  o << "        <DataArray type=\"" << type << "\" Name=\"" << name << "\" NumberOfComponents=\"" << ncomp
    << "\" format=\"appended\" offset=\"" << offset << "\"/>\n";
  return o.str ();
}

} // namespace

void
write_vtu (const std::string &path, int dim, long long n_points, const double *coordinates, long long n_cells,
           const long long *connectivity, const std::vector<VtuPointField> &point_fields,
           const std::vector<VtuCellField> &cell_fields)
{
  if (dim != 2 && dim != 3)
    throw std::invalid_argument ("write_vtu: dim must be 2 or 3");
  const int nv = 1 << dim;
  // VTK_QUAD / VTK_HEXAHEDRON number the vertices counter-clockwise per z level
  static const int vtk_order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
  Appended app;
  std::vector<double> xyz ((size_t) n_points * 3, 0.0);
  for (long long p = 0; p < n_points; ++p)
    for (int d = 0; d < dim; ++d)
      xyz[(size_t) (3 * p + d)] = coordinates[(size_t) (dim * p + d)];
  std::vector<int64_t> conn ((size_t) n_cells * nv), offsets ((size_t) n_cells);
  for (long long c = 0; c < n_cells; ++c)
    {
      for (int v = 0; v < nv; ++v)
        conn[(size_t) (nv * c + v)] = connectivity[(size_t) (nv * c + vtk_order[v])];
      offsets[(size_t) c] = (int64_t) nv * (c + 1);
    }
  std::vector<uint8_t> types ((size_t) n_cells, (uint8_t) (dim == 2 ? 9 : 12));

  std::ostringstream head;
  head << "<?xml version=\"1.0\"?>\n"
       << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
       << "  <UnstructuredGrid>\n"
       << "    <Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">\n";
  head << "      <Points>\n" << array_tag ("Float64", "points", 3, app.add (xyz)) << "      </Points>\n";
  head << "      <Cells>\n"
       << array_tag ("Int64", "connectivity", 1, app.add (conn)) << array_tag ("Int64", "offsets", 1, app.add (offsets))
       << array_tag ("UInt8", "types", 1, app.add (types)) << "      </Cells>\n";
  head << "      <PointData>\n";
  for (const VtuPointField &f : point_fields)
    {
      if ((long long) f.values.size () != n_points * f.n_components)
        throw std::invalid_argument ("write_vtu: point field " + f.name + " has the wrong size");
      head << array_tag ("Float64", f.name, f.n_components, app.add (f.values));
    }
  head << "      </PointData>\n      <CellData>\n";
  for (const VtuCellField &f : cell_fields)
    {
      if ((long long) f.values.size () != n_cells)
        throw std::invalid_argument ("write_vtu: cell field " + f.name + " has the wrong size");
      head << array_tag ("Float32", f.name, 1, app.add (f.values));
    }
  head << "      </CellData>\n    </Piece>\n  </UnstructuredGrid>\n  <AppendedData encoding=\"raw\">\n_";
  std::ofstream out (path.c_str (), std::ios::binary);
  if (!out)
    throw std::runtime_error ("cannot write " + path);
  out << head.str ();
  out.write (app.bytes.data (), (std::streamsize) app.bytes.size ());
  out << "\n  </AppendedData>\n</VTKFile>\n";
}

void
write_pvtu_record (const std::string &path, const std::vector<std::string> &piece_names,
                   const std::vector<VtuPointField> &point_fields, const std::vector<VtuCellField> &cell_fields)
{
  std::ofstream out (path.c_str ());
  out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
      << "  <PUnstructuredGrid GhostLevel=\"0\">\n    <PPointData>\n";
  for (const VtuPointField &f : point_fields)
    out << "      <PDataArray type=\"Float64\" Name=\"" << f.name << "\" NumberOfComponents=\"" << f.n_components << "\"/>\n";
  out << "    </PPointData>\n    <PCellData>\n";
  for (const VtuCellField &f : cell_fields)
    out << "      <PDataArray type=\"Float32\" Name=\"" << f.name << "\" NumberOfComponents=\"1\"/>\n";
  out << "    </PCellData>\n    <PPoints>\n      <PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n    </PPoints>\n";
  for (const std::string &p : piece_names)
    out << "    <Piece Source=\"" << p << "\"/>\n";
  out << "  </PUnstructuredGrid>\n</VTKFile>\n";
}

void
write_visit_record (const std::string &path, const std::vector<std::vector<std::string>> &pieces_by_timestep)
{
  std::ofstream out (path.c_str ());
  if (pieces_by_timestep.empty ())
    return;
  out << "!NBLOCKS " << pieces_by_timestep[0].size () << "\n";
  for (const auto &step : pieces_by_timestep)
    for (const std::string &p : step)
      out << p << "\n";
}

void
write_pvd_record (const std::string &path, const std::vector<std::pair<double, std::string>> &times_and_names)
{
  std::ofstream out (path.c_str ());
  out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" ByteOrder=\"LittleEndian\">\n  <Collection>\n";
  out.precision (12);
  for (const auto &tn : times_and_names)
    out << "    <DataSet timestep=\"" << tn.first << "\" group=\"\" part=\"0\" file=\"" << tn.second << "\"/>\n";
  out << "  </Collection>\n</VTKFile>\n";
}

} // namespace cracks
