// md-flexible's VTK checkpoint record written straight from the device SoA (SURVEY.md section 8 f4;
// examples/md-flexible/src/ParallelVtkWriter.cpp:55-201 recordParticleStates, :308-356 createParticlesPvtuFile).
//
// The reference walks the owned particles five times and streams velocities, forces, type ids, ids and positions as
// ASCII rows ("        a b c\n", doubles as "%.6g", positions near the upper box corner with raised precision). Here
// every owned slot formats its own rows with the exact decimal conversion of vtk_format.cuh, twice: a measuring pass
// gives the row lengths, one exclusive scan per data array turns them into byte offsets, and the writing pass puts the
// text where it belongs; the XML scaffolding between the arrays comes from the host. The finished record leaves in one
// device -> host copy (apb_vtk_particle_record) or through pinned buffers into a file (apb_vtk_write_particle_record).
// Rows follow the storage order (the order the container's iterators visit), like the reference's. The second half of
// the file reads such a piece back (apb_vtk_load_particle_record).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

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

// One thread per (slot, data array), the data array in blockIdx.y: the five arrays of a particle are independent work,
// and a warp that formats one array for 32 consecutive slots runs one code path over coalesced column reads (the first
// version interleaved the arrays within a warp: 6 of 32 lanes active per instruction, ncu).
__global__ void __launch_bounds__(128) kVtkMeasure(int64_t n, int64_t m, const int *__restrict__ flag,
                                                   const int *__restrict__ rowOf, VtkCols c,
                                                   const ApbVtkTables *__restrict__ t, int *__restrict__ len,
                                                   int *__restrict__ bad) {
  const int64_t slot = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int section = static_cast<int>(blockIdx.y);
  if (slot >= n || !flag[slot]) return;
  char row[kRowMax];
  int l = vtkRow(c, t, slot, section, row);
  if (l < 0) {
    atomicExch(bad, 1);
    l = 0;
  }
  len[section * m + rowOf[slot]] = l;
}

// The rows of a warp's owned slots are consecutive in the record (row index and byte offset are exclusive scans over the
// slots; a chunk of rows starts where the previous one ends), so the warp formats them into shared memory at their relative offsets and stores the whole stretch (~1 KB) with
// aligned 16-byte stores; per-thread byte stores wrote 1.76 x the record's bytes to DRAM (partial sectors, ncu).
__global__ void __launch_bounds__(128) kVtkWrite(int64_t n, int64_t m, const int *__restrict__ flag,
                                                 const int *__restrict__ rowOf, VtkCols c,
                                                 const ApbVtkTables *__restrict__ t, const int *__restrict__ off,
                                                 const long long *__restrict__ base, int numChunks, int chunkRows,
                                                 char *__restrict__ out) {
  __shared__ __align__(16) char stage[4][32 * kRowMax + 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t slot = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int section = static_cast<int>(blockIdx.y);
  const bool active = slot < n && flag[slot];
  long long myOff = 0;
  if (active) {  // byte offsets are 32-bit within a chunk of rows; base[] holds where each chunk of each data array starts
    const int row = rowOf[slot];
    myOff = base[section * numChunks + row / chunkRows] + off[section * m + row];
  }
  const unsigned owners = __ballot_sync(0xffffffffu, active);
  if (owners == 0u) return;
  const long long warpOff = __shfl_sync(0xffffffffu, myOff, __ffs(owners) - 1);
  // the stretch starts at the same position within a 16-byte line in shared memory as in the record
  const int skew = static_cast<int>((reinterpret_cast<uintptr_t>(out) + static_cast<uintptr_t>(warpOff)) & 15u);
  char *sh = stage[warp];
  int endMine = 0;
  if (active) {
    const int rel = static_cast<int>(myOff - warpOff) + skew;
    endMine = rel + vtkRow(c, t, slot, section, sh + rel);
  }
  const int end = __shfl_sync(0xffffffffu, endMine, 31 - __clz(owners));
  __syncwarp();
  char *dst = out + warpOff - skew;  // 16-byte aligned
  const int lineBegin = (skew + 15) & ~15, lineEnd = end & ~15;
  if (lineBegin <= lineEnd) {
    for (int i = skew + lane; i < lineBegin; i += 32) dst[i] = sh[i];
    for (int i = lineBegin + 16 * lane; i < lineEnd; i += 512) *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(sh + i);
    for (int i = lineEnd + lane; i < end; i += 32) dst[i] = sh[i];
  } else {
    for (int i = skew + lane; i < end; i += 32) dst[i] = sh[i];
  }
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

// Rows per scan: the byte offsets of a data array are 32-bit exclusive scans of the row lengths, so the rows are scanned
// in chunks of 2^24 (at most 1.6 GB of text each) whose starting offsets are added up in 64 bits on the host.
static long long vtkChunkRows() {
  static const long long rows = [] {
    const char *e = std::getenv("APB_VTK_CHUNK_ROWS");  // tests: exercise the chunked path on small containers
    const long long v = e ? std::atoll(e) : 0;
    return v > 0 ? v : (1ll << 24);
  }();
  return rows;
}

// Measures the record (measuring pass and scans) and, if it has at most `writeUpTo` bytes, formats it into h->vtkOut.
static int vtkBuild(apb_handle h, long long writeUpTo, long long *totalOut) {
  if (h->cfg.particle_kind != APB_PARTICLE_LJ)
    return h->fail(APB_ERR_NOT_APPLICABLE, "the checkpoint record is the one of MoleculeLJ (md-flexible's single-site mode)");
  const int64_t n = h->nslots;
  if (n > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "apb_vtk_particle_record: more than 2^31 slots");
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
    APB_CHECK(apbEnsure(h, h->vtkCtl, 64));
    APB_CHECK(apbEnsure(h, h->vtkFlag, sizeof(int) * n * 2));
    flag = static_cast<int *>(h->vtkFlag.p);
    rowOf = flag + n;
    ++h->launchCount, kVtkSelect<<<apbDivUp(n, 256), 256, 0, h->stream>>>(n, h->own, flag);
    APB_CHECK(apbExclusiveScan(h, flag, rowOf, n, static_cast<long long *>(h->vtkCtl.p)));
    APB_CUDA(cudaMemcpyAsync(&m, h->vtkCtl.p, 8, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
  }
  const long long chunkRows = vtkChunkRows();
  const long long numChunks = std::max(1ll, (m + chunkRows - 1) / chunkRows);
  // control words: [0] error flag, then per (data array, chunk) the bytes of its rows, then where it starts in the record
  APB_CHECK(apbEnsure(h, h->vtkCtl, 8 * static_cast<size_t>(1 + 2 * kSections * numChunks)));
  long long *ctl = static_cast<long long *>(h->vtkCtl.p);
  int *bad = reinterpret_cast<int *>(ctl);
  long long *chunkBytes = ctl + 1, *chunkBase = chunkBytes + kSections * numChunks;
  APB_CUDA(cudaMemsetAsync(ctl, 0, 8 * static_cast<size_t>(1 + 2 * kSections * numChunks), h->stream));
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
  std::vector<long long> bytesHost(static_cast<size_t>(kSections * numChunks), 0), baseHost(static_cast<size_t>(kSections * numChunks), 0);
  int *len = nullptr, *off = nullptr;
  if (m > 0) {
    APB_CHECK(apbEnsure(h, h->vtkLen, sizeof(int) * m * kSections * 2));
    len = static_cast<int *>(h->vtkLen.p);
    off = len + m * kSections;
    ++h->launchCount, kVtkMeasure<<<dim3(static_cast<unsigned>(apbDivUp(n, 128)), kSections), 128, 0, h->stream>>>(n, m, flag, rowOf, c, tables, len, bad);
    APB_CUDA(cudaGetLastError());
    for (int s = 0; s < kSections; ++s)
      for (long long k = 0; k < numChunks; ++k) {
        const long long first = s * m + k * chunkRows;
        APB_CHECK(apbExclusiveScan(h, len + first, off + first, std::min(chunkRows, m - k * chunkRows), chunkBytes + s * numChunks + k));
      }
    int hostBad = 0;
    APB_CUDA(cudaMemcpyAsync(bytesHost.data(), chunkBytes, 8 * bytesHost.size(), cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaMemcpyAsync(&hostBad, bad, 4, cudaMemcpyDeviceToHost, h->stream));
    APB_CUDA(cudaStreamSynchronize(h->stream));
    if (hostBad)  // the reference throws std::runtime_error here (ParallelVtkWriter.cpp:141-149)
      return h->fail(APB_ERR_INVALID_ARGUMENT,
                     "ParallelVtkWriter::writeWithDynamicPrecision(): a position is identical to the box border up to 15 digits of precision");
  }
  std::string text[kSections + 1];
  long long textAt[kSections + 1];
  long long total = 0;
  for (int s = 0; s <= kSections; ++s) {
    text[s] = scaffold(s, m);
    textAt[s] = total;
    total += static_cast<long long>(text[s].size());
    if (s < kSections)
      for (long long k = 0; k < numChunks; ++k) {
        baseHost[s * numChunks + k] = total;
        total += bytesHost[s * numChunks + k];
      }
  }
  *totalOut = total;
  if (total > writeUpTo) return APB_OK;
  APB_CHECK(apbEnsure(h, h->vtkOut, static_cast<size_t>(total)));
  char *out = static_cast<char *>(h->vtkOut.p);
  for (int s = 0; s <= kSections; ++s)
    APB_CUDA(cudaMemcpyAsync(out + textAt[s], text[s].data(), text[s].size(), cudaMemcpyHostToDevice, h->stream));
  if (m > 0) {
    APB_CUDA(cudaMemcpyAsync(chunkBase, baseHost.data(), 8 * baseHost.size(), cudaMemcpyHostToDevice, h->stream));
    ++h->launchCount, kVtkWrite<<<dim3(static_cast<unsigned>(apbDivUp(n, 128)), kSections), 128, 0, h->stream>>>(
        n, m, flag, rowOf, c, tables, off, chunkBase, static_cast<int>(numChunks), static_cast<int>(std::min<long long>(chunkRows, 0x7fffffffLL)), out);
    APB_CUDA(cudaGetLastError());
  }
  // (the host strings and vectors above are pageable: cudaMemcpyAsync has staged them before it returned)
  return APB_OK;
}

extern "C" int apb_vtk_particle_record(apb_handle h, void *dst, int64_t capacityBytes, int64_t *outBytes) {
  APB_ENTRY(h);
  if (outBytes) *outBytes = 0;
  if (capacityBytes < 0 || (!dst && !outBytes)) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_particle_record: bad argument");
  long long total = 0;
  APB_CHECK(vtkBuild(h, dst ? capacityBytes : -1, &total));
  if (outBytes) *outBytes = total;
  if (!dst) return APB_OK;  // size query
  if (total > capacityBytes)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_particle_record: the record has " + std::to_string(total) + " bytes, the buffer " + std::to_string(capacityBytes));
  APB_CUDA(cudaMemcpyAsync(dst, h->vtkOut.p, static_cast<size_t>(total), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  return APB_OK;
}

// The same record straight into a file, like the reference writer (ParallelVtkWriter.cpp:61-70, 200): the text crosses
// PCIe in 32 MB pieces through two pinned buffers, the copy of one piece overlapping the fwrite of the previous one.
extern "C" int apb_vtk_write_particle_record(apb_handle h, const char *path, int64_t *outBytes) {
  APB_ENTRY(h);
  if (outBytes) *outBytes = 0;
  if (!path) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_write_particle_record: no file name");
  long long total = 0;
  APB_CHECK(vtkBuild(h, std::numeric_limits<long long>::max(), &total));
  std::FILE *file = std::fopen(path, "wb");
  if (!file)  // the reference throws std::runtime_error with this text (:68-70)
    return h->fail(APB_ERR_INVALID_ARGUMENT, std::string("Simulation::writeVTKFile(): Failed to open file \"") + path + "\"");
  const long long piece = 32ll << 20;
  const long long numPieces = (total + piece - 1) / piece;
  int rc = apbEnsurePinned(h, static_cast<size_t>(2 * piece));
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaError_t ce = cudaSuccess;
  if (rc == APB_OK) {
    for (auto &e : ev)
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    const char *out = static_cast<const char *>(h->vtkOut.p);
    char *stage = static_cast<char *>(h->pinned);
    auto fetch = [&](long long k) {
      const long long len = std::min(piece, total - k * piece);
      ce = cudaMemcpyAsync(stage + (k & 1) * piece, out + k * piece, static_cast<size_t>(len), cudaMemcpyDeviceToHost, h->stream);
      if (ce == cudaSuccess) ce = cudaEventRecord(ev[k & 1], h->stream);
    };
    bool writeFailed = false;
    if (ce == cudaSuccess && numPieces > 0) fetch(0);
    for (long long k = 0; k < numPieces && ce == cudaSuccess && !writeFailed; ++k) {
      if (k + 1 < numPieces) fetch(k + 1);  // its buffer held piece k - 1, which fwrite has consumed
      if (ce == cudaSuccess) ce = cudaEventSynchronize(ev[k & 1]);
      const size_t len = static_cast<size_t>(std::min(piece, total - k * piece));
      if (ce == cudaSuccess) writeFailed = std::fwrite(stage + (k & 1) * piece, 1, len, file) != len;
    }
    cudaStreamSynchronize(h->stream);
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
    if (writeFailed) rc = h->fail(APB_ERR_INVALID_ARGUMENT, std::string("apb_vtk_write_particle_record: writing \"") + path + "\" failed");
  }
  if (std::fclose(file) != 0 && rc == APB_OK && ce == cudaSuccess)
    rc = h->fail(APB_ERR_INVALID_ARGUMENT, std::string("apb_vtk_write_particle_record: closing \"") + path + "\" failed");
  if (ce != cudaSuccess) return h->failCuda(ce, "checkpoint copy", __FILE__, __LINE__);
  if (rc == APB_OK && outBytes) *outBytes = total;
  return rc;
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

// ------------------------------------------------------------------------------------------------------------------
// Reading a piece back (loadParticlesFromRankRecord, examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180):
// NumberOfPoints, then for "velocities", "forces", "typeIds", "ids", "positions" in this order: skip to the end of the
// line that names the array and read NumberOfPoints x {3, 3, 1, 1, 3} whitespace-separated values with operator>>.
// The text is copied to the device once; the host only looks at the XML tags (their positions come from a kernel that
// lists every '<' - the payload has none). Per data array: threads count the tokens that start in their 128-byte
// segment, a scan numbers them, and a second walk converts token t into component t % k of particle t / k with the
// exact decimal -> binary conversion of vtk_format.cuh.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kSegment = 128;   // bytes of text per thread
constexpr int kTokenMax = 72;   // longest token converted (40 digits, sign, point, exponent)
constexpr int kTagCap = 256;    // '<' characters expected in a piece: 25

__device__ __forceinline__ bool vtkIsSpace(char ch) {  // std::isspace in the "C" locale, what operator>> skips
  return ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f';
}

__global__ void kVtkFindTags(const char *__restrict__ text, long long numBytes, int *__restrict__ count,
                             long long *__restrict__ where) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= numBytes || text[i] != '<') return;
  const int k = atomicAdd(count, 1);
  if (k < kTagCap) where[k] = i;
}

// tokens that START in segment g of the payload [begin, end)
__global__ void kVtkCountTokens(const char *__restrict__ text, long long begin, long long end, int *__restrict__ count) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long a = begin + g * kSegment;
  if (a >= end) return;
  const long long b = min(a + kSegment, end);
  bool prevSpace = a == begin ? true : vtkIsSpace(text[a - 1]);
  int c = 0;
  for (long long i = a; i < b; ++i) {
    const bool sp = vtkIsSpace(text[i]);
    c += prevSpace && !sp;
    prevSpace = sp;
  }
  count[g] = c;
}

struct VtkLoadTarget {
  double *d[3];      // columns receiving components 0 .. 2 (doubles), or
  int64_t *id;       // ids (k == 1), or
  int32_t *type;     // type ids (k == 1)
  int k;             // values per particle
  long long first;   // slot of particle 0
  long long numParticles;
};

__global__ void kVtkParseTokens(const char *__restrict__ text, long long begin, long long end, const int *__restrict__ base,
                                VtkLoadTarget t, int *__restrict__ bad) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long a = begin + g * kSegment;
  if (a >= end) return;
  const long long b = min(a + kSegment, end);
  bool prevSpace = a == begin ? true : vtkIsSpace(text[a - 1]);
  long long token = base[g];
  for (long long i = a; i < b; ++i) {
    const bool sp = vtkIsSpace(text[i]);
    if (prevSpace && !sp) {
      const long long particle = token / t.k;
      const int comp = static_cast<int>(token % t.k);
      ++token;
      if (particle < t.numParticles) {  // the reference reads NumberOfPoints x k values and stops
        char tok[kTokenMax];
        int len = 0;
        for (long long j = i; j < end && !vtkIsSpace(text[j]); ++j) {
          if (len == kTokenMax) {
            len = -1;
            break;
          }
          tok[len++] = text[j];
        }
        int status = len < 0 ? 1 : 0;
        if (!status) {
          if (t.id) t.id[t.first + particle] = static_cast<int64_t>(apbParseU64(tok, len, status));
          else if (t.type) t.type[t.first + particle] = static_cast<int32_t>(apbParseU64(tok, len, status));
          else t.d[comp][t.first + particle] = apbParseDouble(tok, len, status);
        }
        if (status) atomicOr(bad, 1);
      }
    }
    prevSpace = sp;
  }
}

__global__ void kVtkFinishLoaded(long long first, long long n, const double *__restrict__ x, const double *__restrict__ y,
                                 const double *__restrict__ z, int32_t *__restrict__ own, double lox, double loy, double loz,
                                 double hix, double hiy, double hiz, int checkBox, int *__restrict__ bad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  own[first + i] = APB_OWN_OWNED;
  if (checkBox) {  // utils::inBox is half-open [lo, hi) (utils/inBox.h:26-36)
    const double px = x[first + i], py = y[first + i], pz = z[first + i];
    if (!(px >= lox && px < hix && py >= loy && py < hiy && pz >= loz && pz < hiz)) atomicOr(bad, 2);
  }
}

// position just behind the first `word` at or after `from` in [text, text + n); -1 if absent
long long findBehind(const char *text, long long n, long long from, const std::string &word) {
  if (from < 0 || from >= n) return -1;
  const void *p = memmem(text + from, static_cast<size_t>(n - from), word.data(), word.size());
  return p ? static_cast<const char *>(p) - text + static_cast<long long>(word.size()) : -1;
}

}  // namespace

extern "C" int apb_vtk_load_particle_record(apb_handle h, const void *src, int64_t numBytes, int32_t checkBox, int64_t *outNum) {
  APB_ENTRY(h);
  if (outNum) *outNum = 0;
  if (!src || numBytes <= 0) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_load_particle_record: empty record");
  if (h->cfg.particle_kind != APB_PARTICLE_LJ)
    return h->fail(APB_ERR_NOT_APPLICABLE, "the checkpoint record is the one of MoleculeLJ (md-flexible's single-site mode)");
  const char *host = static_cast<const char *>(src);
  // the text on the device, and where its tags are
  APB_CHECK(apbEnsure(h, h->vtkOut, static_cast<size_t>(numBytes)));
  APB_CHECK(apbEnsure(h, h->vtkCtl, 64 + 8 * kTagCap));
  char *text = static_cast<char *>(h->vtkOut.p);
  int *ctl = static_cast<int *>(h->vtkCtl.p);  // [0] number of '<', [1] error flags; positions from byte 64 on
  long long *whereDev = reinterpret_cast<long long *>(static_cast<char *>(h->vtkCtl.p) + 64);
  APB_CUDA(cudaMemsetAsync(ctl, 0, 64, h->stream));
  APB_CUDA(cudaMemcpyAsync(text, host, static_cast<size_t>(numBytes), cudaMemcpyHostToDevice, h->stream));
  ++h->launchCount, kVtkFindTags<<<apbDivUp(numBytes, 256), 256, 0, h->stream>>>(text, numBytes, ctl, whereDev);
  APB_CUDA(cudaGetLastError());
  int numTags = 0;
  APB_CUDA(cudaMemcpyAsync(&numTags, ctl, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  if (numTags <= 0 || numTags > kTagCap) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_load_particle_record: not a .vtu piece of md-flexible");
  std::vector<long long> tags(static_cast<size_t>(numTags));
  APB_CUDA(cudaMemcpyAsync(tags.data(), whereDev, 8 * tags.size(), cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  std::sort(tags.begin(), tags.end());
  // NumberOfPoints="N" (MDFlexConfig.cpp:121-130); the words are looked for where the tags are, not in the payload
  const long long headEnd = std::min<long long>(numBytes, 4096);
  long long at = findBehind(host, headEnd, 0, "NumberOfPoints");
  if (at >= 0) at = findBehind(host, headEnd, at, "\"");
  long long numParticles = 0;
  for (; at >= 0 && at < headEnd && host[at] >= '0' && host[at] <= '9'; ++at) numParticles = numParticles * 10 + (host[at] - '0');
  if (numParticles <= 0)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "Could not determine the number of particles in the checkpoint file");
  if (h->nslots + numParticles > 0x7fffffffLL) return h->fail(APB_ERR_NOT_APPLICABLE, "apb_vtk_load_particle_record: more than 2^31 slots");
  // payload of every data array: from behind the line that names it to the next tag
  const char *names[kSections] = {"velocities", "forces", "typeIds", "ids", "positions"};
  long long payloadBegin[kSections], payloadEnd[kSections];
  size_t tagIndex = 0;
  for (int s = 0; s < kSections; ++s) {
    long long found = -1;
    for (; tagIndex < tags.size() && found < 0; ++tagIndex) {
      const long long tagEnd = tagIndex + 1 < tags.size() ? tags[tagIndex + 1] : numBytes;
      const long long limit = std::min(tagEnd, tags[tagIndex] + 256);
      const long long behind = findBehind(host, limit, tags[tagIndex], std::string("\"") + names[s] + "\"");
      if (behind >= 0) found = findBehind(host, tagEnd, behind, "\n");
    }
    if (found < 0 || tagIndex >= tags.size())
      return h->fail(APB_ERR_INVALID_ARGUMENT, std::string("apb_vtk_load_particle_record: data array \"") + names[s] + "\" not found");
    payloadBegin[s] = found;
    payloadEnd[s] = tags[tagIndex];  // the closing </DataArray>
    if (payloadEnd[s] - payloadBegin[s] > 0x7fffffffLL)  // (token numbers are 32-bit scans)
      return h->fail(APB_ERR_NOT_APPLICABLE, "apb_vtk_load_particle_record: a data array of more than 2^31 bytes");
  }
  // append the particles
  h->ownedKnown = false;
  h->ownedInsideBox = false;
  const int64_t first = h->nslots;
  APB_CHECK(apbReserveSlots(h, first + numParticles));
  for (int c = APB_COL_X; c < APB_NUM_COLUMNS; ++c)
    if (h->active[c]) APB_CUDA(cudaMemsetAsync(h->col[c] + first, 0, sizeof(double) * numParticles, h->stream));
  APB_CUDA(cudaMemsetAsync(h->id + first, 0, sizeof(int64_t) * numParticles, h->stream));
  APB_CUDA(cudaMemsetAsync(h->type + first, 0, sizeof(int32_t) * numParticles, h->stream));
  int *bad = ctl + 1;
  long long *tokenTotal = reinterpret_cast<long long *>(ctl + 2);
  long long tokensHost[kSections] = {0, 0, 0, 0, 0};
  for (int s = 0; s < kSections; ++s) {
    const long long bytes = payloadEnd[s] - payloadBegin[s];
    const long long segments = (bytes + kSegment - 1) / kSegment;
    if (segments <= 0) continue;
    APB_CHECK(apbEnsure(h, h->vtkLen, sizeof(int) * segments * 2));
    int *count = static_cast<int *>(h->vtkLen.p), *base = count + segments;
    ++h->launchCount, kVtkCountTokens<<<apbDivUp(segments, 128), 128, 0, h->stream>>>(text, payloadBegin[s], payloadEnd[s], count);
    APB_CUDA(cudaGetLastError());
    APB_CHECK(apbExclusiveScan(h, count, base, segments, tokenTotal));
    APB_CUDA(cudaMemcpyAsync(&tokensHost[s], tokenTotal, 8, cudaMemcpyDeviceToHost, h->stream));
    VtkLoadTarget t{};
    t.k = (s == 2 || s == 3) ? 1 : 3;
    t.first = first;
    t.numParticles = numParticles;
    const int colBase = s == 0 ? APB_COL_VX : s == 1 ? APB_COL_FX : APB_COL_X;
    for (int d = 0; d < 3; ++d) t.d[d] = h->col[colBase + d];
    t.id = s == 3 ? h->id : nullptr;
    t.type = s == 2 ? h->type : nullptr;
    ++h->launchCount, kVtkParseTokens<<<apbDivUp(segments, 128), 128, 0, h->stream>>>(text, payloadBegin[s], payloadEnd[s], base, t, bad);
    APB_CUDA(cudaGetLastError());
  }
  ++h->launchCount, kVtkFinishLoaded<<<apbDivUp(numParticles, 256), 256, 0, h->stream>>>(
      first, numParticles, h->col[APB_COL_X], h->col[APB_COL_Y], h->col[APB_COL_Z], h->own, h->cfg.box_min[0], h->cfg.box_min[1],
      h->cfg.box_min[2], h->cfg.box_max[0], h->cfg.box_max[1], h->cfg.box_max[2], checkBox, bad);
  APB_CUDA(cudaGetLastError());
  int hostBad = 0;
  APB_CUDA(cudaMemcpyAsync(&hostBad, bad, 4, cudaMemcpyDeviceToHost, h->stream));
  APB_CUDA(cudaStreamSynchronize(h->stream));
  for (int s = 0; s < kSections; ++s)
    if (tokensHost[s] < numParticles * ((s == 2 || s == 3) ? 1 : 3))
      return h->fail(APB_ERR_INVALID_ARGUMENT, std::string("apb_vtk_load_particle_record: data array \"") + names[s] + "\" holds fewer values than NumberOfPoints asks for");
  if (hostBad & 1)
    return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_vtk_load_particle_record: a value is not a decimal number (the reference's stream extraction fails on it)");
  if (hostBad & 2) return h->fail(APB_ERR_PARTICLE_OUTSIDE, "apb_vtk_load_particle_record: a particle of the checkpoint is outside the container box");
  h->nslots = first + numParticles;
  h->structureValid = false;
  h->prunedValid = false;
  h->countsValid = false;
  apbForgetHaloLinks(h);
  if (outNum) *outNum = numParticles;
  return APB_OK;
}
