// md-flexible's VTK checkpoint record written straight from the device SoA (SURVEY.md section 8 f4;
// examples/md-flexible/src/ParallelVtkWriter.cpp:55-201 recordParticleStates, :308-356 createParticlesPvtuFile).
//
// The reference walks the owned particles five times and streams velocities, forces, type ids, ids and positions as
// ASCII rows ("        a b c\n", doubles as "%.6g", positions near the upper box corner with raised precision). Here
// every owned slot formats its own rows with the exact decimal conversion of vtk_format.cuh, twice: a measuring pass
// gives the row lengths, one exclusive scan per data array turns them into byte offsets, and the writing pass puts the
// text where it belongs; the XML scaffolding between the arrays comes from the host. The finished record is one
// device -> host copy. Rows follow the storage order (the order the container's iterators visit), like the reference's.
#include <string>

#include "internal.cuh"
#include "vtk_format.cuh"

namespace {

constexpr int kSections = 5;  // velocities, forces, typeIds, ids, positions
constexpr int kRowMax = 96;   // 8 blanks + 3 x (sign, 17 digits, point, e-308) + separators < 96

struct VtkCols {
  const double *v[3], *f[3], *r[3];
  const int64_t *id;
  const int32_t *type;
  double boxMax[3];
};

__device__ __forceinline__ int vtkTriple(const double *const c[3], int64_t slot, const int P[3], char *row) {
  int len = 0;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) row[len++] = ' ';
#pragma unroll 1
  for (int d = 0; d < 3; ++d) {
    if (d) row[len++] = ' ';
    len += apbFormatG(c[d][slot], P[d], row + len);
  }
  row[len++] = '\n';
  return len;
}

// the row of `slot` in data array `section`; returns its length, -1 if a position cannot be told from the border
__device__ int vtkRow(const VtkCols &c, const ApbVtkTables *__restrict__ t, int64_t slot, int section, char *row) {
  const int six[3] = {6, 6, 6};
  if (section == 0) return vtkTriple(c.v, slot, six, row);
  if (section == 1) return vtkTriple(c.f, slot, six, row);
  if (section == 4) {
    int P[3];
    for (int d = 0; d < 3; ++d) {
      P[d] = apbVtkPositionPrecision(*t, c.r[d][slot], c.boxMax[d]);
      if (P[d] < 0) return -1;
    }
    return vtkTriple(c.r, slot, P, row);
  }
  int len = 0;
  for (int k = 0; k < 8; ++k) row[len++] = ' ';
  // getTypeId() / getID() return unsigned long (MoleculeLJ.h, ParticleBase.h)
  len += apbFormatU64(section == 2 ? static_cast<uint64_t>(c.type[slot]) : static_cast<uint64_t>(c.id[slot]), row + len);
  row[len++] = '\n';
  return len;
}

__global__ void kVtkSelect(int64_t n, const int32_t *__restrict__ own, int *__restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = own[i] == APB_OWN_OWNED;
}

// one thread per (slot, data array): the formatting of the five arrays of a particle is independent work
__global__ void __launch_bounds__(128) kVtkMeasure(int64_t n, int64_t m, const int *__restrict__ flag,
                                                   const int *__restrict__ rowOf, VtkCols c,
                                                   const ApbVtkTables *__restrict__ t, int *__restrict__ len,
                                                   int *__restrict__ bad) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t slot = g / kSections;
  const int section = static_cast<int>(g % kSections);
  if (slot >= n || !flag[slot]) return;
  char row[kRowMax];
  int l = vtkRow(c, t, slot, section, row);
  if (l < 0) {
    atomicExch(bad, 1);
    l = 0;
  }
  len[section * m + rowOf[slot]] = l;
}

struct VtkBase {
  long long at[kSections];  // byte offset of the first row of each data array in the record
};

__global__ void __launch_bounds__(128) kVtkWrite(int64_t n, int64_t m, const int *__restrict__ flag,
                                                 const int *__restrict__ rowOf, VtkCols c,
                                                 const ApbVtkTables *__restrict__ t, const int *__restrict__ off,
                                                 VtkBase base, char *__restrict__ out) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t slot = g / kSections;
  const int section = static_cast<int>(g % kSections);
  if (slot >= n || !flag[slot]) return;
  char row[kRowMax];
  const int l = vtkRow(c, t, slot, section, row);
  char *dst = out + base.at[section] + off[section * m + rowOf[slot]];
  for (int k = 0; k < l; ++k) dst[k] = row[k];
}

const char *const kArrayOpen[kSections] = {
    "        <DataArray Name=\"velocities\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n",
    "        <DataArray Name=\"forces\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n",
    "        <DataArray Name=\"typeIds\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n",
    "        <DataArray Name=\"ids\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n",
    "        <DataArray Name=\"positions\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n"};
const char *const kArrayClose = "        </DataArray>\n";

// the text before data array s (s = kSections: the tail of the file), ParallelVtkWriter.cpp:72-76, 84, 92, 124-128, 166-176
std::string scaffold(int s, long long numParticles) {
  std::string t;
  if (s == 0) {
    t += "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n";
    t += "<VTKFile byte_order=\"LittleEndian\" type=\"UnstructuredGrid\" version=\"0.1\">\n";
    t += "  <UnstructuredGrid>\n";
    t += "    <Piece NumberOfCells=\"0\" NumberOfPoints=\"" + std::to_string(numParticles) + "\">\n";
    t += "      <PointData>\n";
  } else {
    t += kArrayClose;
  }
  if (s == 4) {
    t += "      </PointData>\n";
    t += "      <CellData/>\n";
    t += "      <Points>\n";
  }
  if (s < kSections) {
    t += kArrayOpen[s];
  } else {
    t += "      </Points>\n";
    t += "      <Cells>\n";
    t += "        <DataArray Name=\"types\" NumberOfComponents=\"0\" format=\"ascii\" type=\"Float32\"/>\n";
    t += "      </Cells>\n";
    t += "    </Piece>\n";
    t += "  </UnstructuredGrid>\n";
    t += "</VTKFile>\n";
  }
  return t;
}

}  // namespace

extern "C" int apb_vtk_particle_record(apb_handle h, void *dst, int64_t capacityBytes, int64_t *outBytes) {
  APB_ENTRY(h);
  if (outBytes) *outBytes = 0;
  if (capacityBytes < 0 || (!dst && !outBytes)) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_particle_record: bad argument");
  if (h->cfg.particle_kind != APB_PARTICLE_LJ)
    return h->fail(APB_ERR_NOT_APPLICABLE, "the checkpoint record is the one of MoleculeLJ (md-flexible's single-site mode)");
  const int64_t n = h->nslots;
  // control words: [0..4] bytes of the rows of each data array, [5] owned particles, then the error flag
  APB_CHECK(apbEnsure(h, h->vtkCtl, 64));
  long long *totals = static_cast<long long *>(h->vtkCtl.p);
  int *bad = reinterpret_cast<int *>(totals + 6);
  APB_CUDA(cudaMemsetAsync(h->vtkCtl.p, 0, 64, h->stream));
  if (!h->vtkTablesReady) {
    static const ApbVtkTables *hostTables = [] {
      auto *t = new ApbVtkTables;
      apbVtkBuildTables(*t);
      return t;
    }();
    APB_CHECK(apbEnsure(h, h->vtkTables, sizeof(ApbVtkTables)));
    APB_CUDA(cudaMemcpyAsync(h->vtkTables.p, hostTables, sizeof(ApbVtkTables), cudaMemcpyHostToDevice, h->stream));
    h->vtkTablesReady = true;
  }
  long long m = 0;
  int *flag = nullptr, *rowOf = nullptr;
  if (n > 0) {
    APB_CHECK(apbEnsure(h, h->vtkFlag, sizeof(int) * n * 2));
    flag = static_cast<int *>(h->vtkFlag.p);
    rowOf = flag + n;
    ++h->launchCount, kVtkSelect<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->own, flag);
    APB_CHECK(apbExclusiveScan(h, flag, rowOf, n, totals + 5));
    APB_CUDA(cudaMemcpyAsync(&m, totals + 5, 8, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (m * kRowMax > 0x7fffffffLL)
    return h->fail(APB_ERR_NOT_APPLICABLE, "apb_vtk_particle_record: more than 2^31 bytes per data array (" + std::to_string(m) + " particles)");
  VtkCols c;
  for (int d = 0; d < 3; ++d) {
    c.r[d] = h->col[APB_COL_X + d];
    c.v[d] = h->col[APB_COL_VX + d];
    c.f[d] = h->col[APB_COL_FX + d];
    c.boxMax[d] = h->cfg.box_max[d];
  }
  c.id = h->id;
  c.type = h->type;
  const ApbVtkTables *tables = static_cast<const ApbVtkTables *>(h->vtkTables.p);
  long long sectionBytes[kSections] = {0, 0, 0, 0, 0};
  int *len = nullptr, *off = nullptr;
  if (m > 0) {
    APB_CHECK(apbEnsure(h, h->vtkLen, sizeof(int) * m * kSections * 2));
    len = static_cast<int *>(h->vtkLen.p);
    off = len + m * kSections;
    ++h->launchCount, kVtkMeasure<<<apbDivUp(n * kSections, 128), 128, 0, h->stream>>>(n, m, flag, rowOf, c, tables, len, bad);
    APB_CUDA(cudaGetLastError());
    for (int s = 0; s < kSections; ++s) APB_CHECK(apbExclusiveScan(h, len + s * m, off + s * m, m, totals + s));
    int hostBad = 0;
    APB_CUDA(cudaMemcpyAsync(sectionBytes, totals, 8 * kSections, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaMemcpyAsync(&hostBad, bad, 4, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    if (hostBad)  // the reference throws std::runtime_error here (ParallelVtkWriter.cpp:141-149)
      return h->fail(APB_ERR_INVALID_ARGUMENT,
                     "ParallelVtkWriter::writeWithDynamicPrecision(): a position is identical to the box border up to 15 digits of precision");
  }
  std::string text[kSections + 1];
  VtkBase base;
  long long total = 0;
  for (int s = 0; s <= kSections; ++s) {
    text[s] = scaffold(s, m);
    total += static_cast<long long>(text[s].size());
    if (s < kSections) {
      base.at[s] = total;
      total += sectionBytes[s];
    }
  }
  if (outBytes) *outBytes = total;
  if (!dst) return APB_OK;  // size query
  if (total > capacityBytes)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_particle_record: the record has " + std::to_string(total) + " bytes, the buffer " + std::to_string(capacityBytes));
  APB_CHECK(apbEnsure(h, h->vtkOut, static_cast<size_t>(total)));
  char *out = static_cast<char *>(h->vtkOut.p);
  long long at = 0;
  for (int s = 0; s <= kSections; ++s) {
    APB_CUDA(cudaMemcpyAsync(out + at, text[s].data(), text[s].size(), cudaMemcpyHostToDevice, h->stream));
    at += static_cast<long long>(text[s].size()) + (s < kSections ? sectionBytes[s] : 0);
  }
  if (m > 0) {
    ++h->launchCount, kVtkWrite<<<apbDivUp(n * kSections, 128), 128, 0, h->stream>>>(n, m, flag, rowOf, c, tables, off, base, out);
    APB_CUDA(cudaGetLastError());
  }
  APB_CUDA(cudaMemcpyAsync(dst, out, static_cast<size_t>(total), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// The ".pvtu" index that rank 0 writes next to the pieces (ParallelVtkWriter.cpp:308-356): host text only.
extern "C" int apb_vtk_pvtu_record(const char *sessionName, int32_t numRanks, uint64_t iteration, int32_t digits, char *dst,
                                   int64_t capacityBytes, int64_t *outBytes) {
  if (!sessionName || numRanks < 0 || digits < 0 || capacityBytes < 0 || (!dst && !outBytes)) return APB_ERR_INVALID_ARGUMENT;
  std::string it = std::to_string(iteration);
  if (static_cast<int>(it.size()) < digits) it.insert(0, static_cast<size_t>(digits) - it.size(), '0');  // setfill('0') << setw(digits)
  std::string t;
  t += "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n";
  t += "<VTKFile byte_order=\"LittleEndian\" type=\"PUnstructuredGrid\" version=\"0.1\">\n";
  t += "  <PUnstructuredGrid GhostLevel=\"0\">\n";
  t += "    <PPointData>\n";
  t += "      <PDataArray Name=\"velocities\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n";
  t += "      <PDataArray Name=\"forces\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n";
  t += "      <PDataArray Name=\"typeIds\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\"/>\n";
  t += "      <PDataArray Name=\"ids\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\"/>\n";
  t += "    </PPointData>\n";
  t += "    <PCellData/>\n";
  t += "    <PPoints>\n";
  t += "      <PDataArray Name=\"positions\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\"/>\n";
  t += "    </PPoints>\n";
  t += "    <PCells>\n";
  t += "      <PDataArray Name=\"types\" NumberOfComponents=\"0\" format=\"ascii\" type=\"Float32\"/>\n";
  t += "    </PCells>\n";
  for (int i = 0; i < numRanks; ++i)
    t += "    <Piece Source=\"./data/" + std::string(sessionName) + "_Particles_" + std::to_string(i) + "_" + it + ".vtu\"/>\n";
  t += "  </PUnstructuredGrid>\n";
  t += "</VTKFile>\n";
  if (outBytes) *outBytes = static_cast<int64_t>(t.size());
  if (!dst) return APB_OK;
  if (static_cast<int64_t>(t.size()) > capacityBytes) return APB_ERR_INVALID_ARGUMENT;
  std::memcpy(dst, t.data(), t.size());
  return APB_OK;
}
