// VTK output of a state, one unstructured-grid file per compartment and time stamp plus a
// ParaView time-sequence file, as Model::write_vtk of the reference
// (dune/copasi/model/diffusion_reaction/model_multi_compartment.impl.hh:218-300: files
// "<path>/<stem>-<compartment>-<00000>.vtu", one vertex-data array per species of the compartment,
// conforming P1 data).  Host-side convenience next to the hot path; dune-grid's
// writer is third party, so the byte layout is ours: ASCII, Float64, the .pvd next to the .vtu files.
#include <sys/stat.h>

#include <cerrno>
#include <cstdio>
#include <fstream>
#include <map>

#include "grid.hpp"
#include "model.hpp"

namespace dcb {

namespace {
void make_dirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && cur != "." && cur != "..") {
        if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) fail("cannot create output directory '", cur, "'");
      }
    }
    if (i < path.size()) cur += path[i];
  }
}
std::string stem_of(std::string path) {
  while (path.size() > 1 && path.back() == '/') path.pop_back();
  auto slash = path.rfind('/');
  return slash == std::string::npos ? path : path.substr(slash + 1);
}
}  // namespace

// timesteps: the time stamps already written under `path` (in/out); cleared unless `append`
void write_vtk(const Grid& g, const Model& m, const double* u, double time, const std::string& path, bool append,
               std::vector<double>& timesteps) {
  if (g.elem_comp.size() != (size_t)g.ne) fail("grid is not bound to a model");
  make_dirs(path);
  if (!append) timesteps.clear();
  const int nd = g.nd(), dim = g.dim;
  // VTK_TRIANGLE / VTK_TETRA; Q1 cells as VTK_PIXEL / VTK_VOXEL, whose corner order is the bit pattern
  const int cell_type = g.elem_kind == 1 ? (dim == 2 ? 8 : 11) : (dim == 2 ? 5 : 10);
  const std::string stem = stem_of(path);
  char num[16];
  snprintf(num, sizeof num, "%05zu", timesteps.size());
  timesteps.push_back(time);
  for (int c = 0; c < m.ncomp(); ++c) {
    const std::string name = stem + "-" + m.comp_names[c];
    const auto& verts = g.comp_vertices[c];
    const int ns = m.comp_nspec[c];
    std::vector<int32_t> local(g.nv, -1);
    for (size_t k = 0; k < verts.size(); ++k) local[verts[k]] = (int32_t)k;
    int64_t ncell = 0;
    for (int64_t e = 0; e < g.ne; ++e) ncell += g.elem_comp[e] == c;
    const std::string file = path + "/" + name + "-" + num + ".vtu";
    FILE* f = fopen(file.c_str(), "w");
    if (!f) fail("cannot open '", file, "' for writing");
    fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
    fprintf(f, "<UnstructuredGrid>\n<Piece NumberOfPoints=\"%zu\" NumberOfCells=\"%lld\">\n", verts.size(), (long long)ncell);
    fprintf(f, "<PointData Scalars=\"%s\">\n", ns ? m.species[m.comp_first[c]].name.c_str() : "");
    for (int s = 0; s < ns; ++s) {
      fprintf(f, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"ascii\">\n",
              m.species[m.comp_first[c] + s].name.c_str());
      for (size_t k = 0; k < verts.size(); ++k)
        fprintf(f, "%.17g%c", u[g.comp_vdof[c][verts[k]] + s], (k % 6 == 5 || k + 1 == verts.size()) ? '\n' : ' ');
      fprintf(f, "</DataArray>\n");
    }
    fprintf(f, "</PointData>\n<Points>\n<DataArray type=\"Float64\" Name=\"Coordinates\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (size_t k = 0; k < verts.size(); ++k) {
      const double* x = &g.coords[(size_t)verts[k] * dim];
      fprintf(f, "%.17g %.17g %.17g\n", x[0], x[1], dim == 3 ? x[2] : 0.0);
    }
    fprintf(f, "</DataArray>\n</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 0; e < g.ne; ++e) {
      if (g.elem_comp[e] != c) continue;
      for (int a = 0; a < nd; ++a) fprintf(f, "%d%c", local[g.elems[e * nd + a]], a + 1 == nd ? '\n' : ' ');
    }
    fprintf(f, "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t k = 1; k <= ncell; ++k) fprintf(f, "%lld%c", (long long)(k * nd), (k % 12 == 0 || k == ncell) ? '\n' : ' ');
    fprintf(f, "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t k = 1; k <= ncell; ++k) fprintf(f, "%d%c", cell_type, (k % 24 == 0 || k == ncell) ? '\n' : ' ');
    fprintf(f, "</DataArray>\n</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
    if (fclose(f) != 0) fail("error while writing '", file, "'");
    // time sequence file, rewritten with every stamp
    const std::string pvd = path + "/" + name + ".pvd";
    FILE* p = fopen(pvd.c_str(), "w");
    if (!p) fail("cannot open '", pvd, "' for writing");
    fprintf(p, "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n<Collection>\n");
    for (size_t i = 0; i < timesteps.size(); ++i)
      fprintf(p, "<DataSet timestep=\"%.17g\" group=\"\" part=\"0\" name=\"\" file=\"%s-%05zu.vtu\"/>\n", timesteps[i], name.c_str(), i);
    fprintf(p, "</Collection>\n</VTKFile>\n");
    fclose(p);
  }
}

}  // namespace dcb
