/* TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 *
 * Plain-C restatement of md-flexible's checkpoint writer for one rank's particles
 * (examples/md-flexible/src/ParallelVtkWriter.cpp:55-201 recordParticleStates, :308-356 createParticlesPvtuFile,
 * :437-441 generateFilename), single-site mode. The reference streams doubles through a default-constructed
 * std::ofstream, i.e. "%.6g" (std::num_put: %g with precision 6), and unsigned long ids / type ids as "%lu".
 * Positions go through writeWithDynamicPrecision (:130-157): if the particle is closer than 0.1 to the upper box
 * corner in that dimension, the precision is raised until the rounded value can be told from the border
 * (autopas::utils::Math::roundFloating, src/autopas/utils/Math.cpp:37-45; isNearAbs, Math.h:318-321), at most up to
 * std::numeric_limits<double>::digits10 = 15 digits - beyond that the reference throws (here: return -2).
 * Pinned by bytes of the unmodified writer: tests/golden/fn_vtk.npz (fixture) and oracle/_ref/vtk_ref_writer (live). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* Math.cpp:37-45 */
static double round_floating(double d, int floatingPrecision) {
  if (d == 0.0) return d;
  const double factor = pow(10, floatingPrecision - ceil(log10(fabs(d))));
  return round(d * factor) / factor;
}

/* ParallelVtkWriter.cpp:130-157; returns the precision the value is written with, -1 if the reference would throw */
int vtk_oracle_position_precision(double position, double border) {
  int precision = 6; /* std::ios_base default */
  if (border - position < 0.1) {
    while (fabs(round_floating(position, precision) - border) <= pow(10, -precision)) {
      ++precision;
      if (precision > 15) return -1;
    }
  }
  return precision;
}

#define EMIT(...)                                                       \
  do {                                                                  \
    int w_ = snprintf(tmp, sizeof tmp, __VA_ARGS__);                    \
    if (out && pos + w_ <= cap) memcpy(out + pos, tmp, (size_t)w_);     \
    pos += w_;                                                          \
  } while (0)

/* The ".vtu" piece of one rank (recordParticleStates). r, v, f: n x 3 row-major. Returns the number of bytes the record
 * has (written only if it fits `cap`; call with out = NULL to measure), -2 if the reference would throw. */
int64_t vtk_oracle_particle_record(int64_t n, const double *r, const double *v, const double *f, const int64_t *id,
                                   const int64_t *type, const double *boxMax, char *out, int64_t cap) {
  char tmp[256];
  int64_t pos = 0;
  EMIT("<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n");
  EMIT("<VTKFile byte_order=\"LittleEndian\" type=\"UnstructuredGrid\" version=\"0.1\">\n");
  EMIT("  <UnstructuredGrid>\n");
  EMIT("    <Piece NumberOfCells=\"0\" NumberOfPoints=\"%lu\">\n", (unsigned long)n);
  EMIT("      <PointData>\n");
  EMIT("        <DataArray Name=\"velocities\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n");
  for (int64_t i = 0; i < n; ++i) EMIT("        %g %g %g\n", v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  EMIT("        </DataArray>\n");
  EMIT("        <DataArray Name=\"forces\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n");
  for (int64_t i = 0; i < n; ++i) EMIT("        %g %g %g\n", f[3 * i], f[3 * i + 1], f[3 * i + 2]);
  EMIT("        </DataArray>\n");
  EMIT("        <DataArray Name=\"typeIds\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n");
  for (int64_t i = 0; i < n; ++i) EMIT("        %lu\n", (unsigned long)type[i]);
  EMIT("        </DataArray>\n");
  EMIT("        <DataArray Name=\"ids\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n");
  for (int64_t i = 0; i < n; ++i) EMIT("        %lu\n", (unsigned long)id[i]);
  EMIT("        </DataArray>\n");
  EMIT("      </PointData>\n");
  EMIT("      <CellData/>\n");
  EMIT("      <Points>\n");
  EMIT("        <DataArray Name=\"positions\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n");
  for (int64_t i = 0; i < n; ++i) {
    int p[3];
    for (int d = 0; d < 3; ++d) {
      p[d] = vtk_oracle_position_precision(r[3 * i + d], boxMax[d]);
      if (p[d] < 0) return -2;
    }
    EMIT("        %.*g %.*g %.*g\n", p[0], r[3 * i], p[1], r[3 * i + 1], p[2], r[3 * i + 2]);
  }
  EMIT("        </DataArray>\n");
  EMIT("      </Points>\n");
  EMIT("      <Cells>\n");
  EMIT("        <DataArray Name=\"types\" NumberOfComponents=\"0\" format=\"ascii\" type=\"Float32\"/>\n");
  EMIT("      </Cells>\n");
  EMIT("    </Piece>\n");
  EMIT("  </UnstructuredGrid>\n");
  EMIT("</VTKFile>\n");
  return pos;
}

/* The ".pvtu" index rank 0 writes next to the pieces (createParticlesPvtuFile). */
int64_t vtk_oracle_pvtu_record(const char *session, int numRanks, uint64_t iteration, int digits, char *out, int64_t cap) {
  char tmp[1024];
  int64_t pos = 0;
  EMIT("<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n");
  EMIT("<VTKFile byte_order=\"LittleEndian\" type=\"PUnstructuredGrid\" version=\"0.1\">\n");
  EMIT("  <PUnstructuredGrid GhostLevel=\"0\">\n");
  EMIT("    <PPointData>\n");
  EMIT("      <PDataArray Name=\"velocities\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n");
  EMIT("      <PDataArray Name=\"forces\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n");
  EMIT("      <PDataArray Name=\"typeIds\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\"/>\n");
  EMIT("      <PDataArray Name=\"ids\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\"/>\n");
  EMIT("    </PPointData>\n");
  EMIT("    <PCellData/>\n");
  EMIT("    <PPoints>\n");
  EMIT("      <PDataArray Name=\"positions\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n");
  EMIT("    </PPoints>\n");
  EMIT("    <PCells>\n");
  EMIT("      <PDataArray Name=\"types\" NumberOfComponents=\"0\" format=\"ascii\" type=\"Float32\"/>\n");
  EMIT("    </PCells>\n");
  for (int i = 0; i < numRanks; ++i)
    EMIT("    <Piece Source=\"./data/%s_Particles_%d_%0*lu.vtu\"/>\n", session, i, digits, (unsigned long)iteration);
  EMIT("  </PUnstructuredGrid>\n");
  EMIT("</VTKFile>\n");
  return pos;
}
