// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Runs the UNMODIFIED md-flexible checkpoint writer (examples/md-flexible/src/ParallelVtkWriter.cpp, compiled where it
// lies, single-site mode) on a stock autopas::AutoPas<MoleculeLJ> filled with the particles of a binary input file, so
// that the bytes of its "<session>_Particles_<rank>_<iteration>.vtu" piece and of the ".pvtu" index can pin the oracle
// restatement (oracle/vtk_oracle.c) and, through it, apb_vtk_particle_record.
//
//   vtk_ref_writer <input.bin> <output folder> <session name> <iteration> <digits>
//   vtk_ref_writer --load <checkpoint.pvtu> <rank> <number of ranks> <output.bin>
//
// The second form runs the UNMODIFIED checkpoint loader, MDFlexConfig::loadParticlesFromCheckpoint
// (examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180, 648-671, compiled where it lies; yaml-cpp's headers
// come from the reference's own libs/ archive, its parser objects are never linked because nothing here parses YAML) on
// a default-constructed configuration and dumps the particles it produced: int64 n, then n records as below.
//
// input.bin: int64 n, double boxMin[3], boxMax[3], then n records of {double r[3], v[3], f[3]; int64 id, type}.
// ParallelVtkWriter::recordParticleStates is private (the public recordTimestep also wants a RegularGridDecomposition,
// i.e. md-flexible's whole YAML configuration); the access specifier is lifted for this translation unit only - the
// writer's source is untouched.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <sys/stat.h>

#include <array>
#include <sstream>
#include <string>
#include <unordered_set>

#include "autopas/AutoPasImpl.h"
#include "autopas/tuning/Configuration.h"
#include "src/TypeDefinitions.h"
#include "src/configuration/MDFlexConfig.h"
#include "src/domainDecomposition/RegularGridDecomposition.h"
// everything ParallelVtkWriter.h includes has been seen (and is guarded): the macro only reaches the writer's class
#define private public
#include "ParallelVtkWriter.h"
#undef private

template class autopas::AutoPas<ParticleType>;

struct OutRec {
  double r[3], v[3], f[3];
  int64_t id, type;
};

static int loadMode(int argc, char **argv) {
  if (argc < 6) return 2;
  MDFlexConfig config;
  config.checkpointfile.value = argv[2];
  const auto t0 = std::chrono::steady_clock::now();
  config.loadParticlesFromCheckpoint(static_cast<size_t>(std::atoll(argv[3])), static_cast<size_t>(std::atoll(argv[4])));
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::FILE *out = std::fopen(argv[5], "wb");
  if (!out) return 3;
  const int64_t n = static_cast<int64_t>(config.particles.size());
  std::fwrite(&n, 8, 1, out);
  for (const auto &p : config.particles) {
    OutRec q{{p.getR()[0], p.getR()[1], p.getR()[2]}, {p.getV()[0], p.getV()[1], p.getV()[2]}, {p.getF()[0], p.getF()[1], p.getF()[2]},
             static_cast<int64_t>(p.getID()), static_cast<int64_t>(p.getTypeId())};
    std::fwrite(&q, sizeof q, 1, out);
  }
  std::fclose(out);
  std::printf("%lld %.6f\n", static_cast<long long>(n), seconds);
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "--load") return loadMode(argc, argv);
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s input.bin folder session iteration digits\n", argv[0]);
    return 2;
  }
  std::FILE *in = std::fopen(argv[1], "rb");
  if (!in) return 3;
  int64_t n = 0;
  double box[6];
  if (std::fread(&n, 8, 1, in) != 1 || std::fread(box, 8, 6, in) != 6) return 4;
  struct Rec {
    double r[3], v[3], f[3];
    int64_t id, type;
  };
  std::vector<Rec> recs(static_cast<size_t>(n));
  if (n > 0 && std::fread(recs.data(), sizeof(Rec), recs.size(), in) != recs.size()) return 5;
  std::fclose(in);

  autopas::AutoPas<ParticleType> autoPas;
  autoPas.setBoxMin({box[0], box[1], box[2]});
  autoPas.setBoxMax({box[3], box[4], box[5]});
  autoPas.setCutoff(1.0);
  autoPas.setVerletSkin(0.2);
  autoPas.setAllowedContainers({autopas::ContainerOption::linkedCells});
  autoPas.setAllowedTraversals({autopas::TraversalOption::lc_c08});
  autoPas.setOutputSuffix("apb_vtk_ref");
  autoPas.init();
  for (const auto &q : recs) {
    ParticleType p({q.r[0], q.r[1], q.r[2]}, {q.v[0], q.v[1], q.v[2]}, static_cast<unsigned long>(q.id),
                   static_cast<unsigned long>(q.type));
    p.setF({q.f[0], q.f[1], q.f[2]});
    autoPas.addParticle(p);
  }
  ParallelVtkWriter writer(argv[3], argv[2], std::atoi(argv[5]));
  const auto t0 = std::chrono::steady_clock::now();
  writer.recordParticleStates(static_cast<size_t>(std::atoll(argv[4])), autoPas);
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // particles, seconds the writer took (bench.py quotes it next to the device-side record)
  std::printf("%zu %.6f\n", autoPas.getNumberOfParticles(autopas::IteratorBehavior::owned), seconds);
  return 0;
}
