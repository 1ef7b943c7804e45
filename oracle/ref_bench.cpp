// TEST / MEASUREMENT INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Timing driver for the reference arm of bench.py (`--impl reference` and the `cpu_baseline` leg): the UNMODIFIED
// AutoPas reference headers (compiled from /root/reference where they lie; nothing is copied), built the way the
// reference builds its own Release binaries — `-O3`, the host's vector ISA (cmake/modules/autopas_vectorization.cmake:15
// uses -march=native; oracle/Makefile builds one library per x86-64 micro-architecture level because the GPU box's host
// CPU is not the build container's, and bench.py loads the highest level the host supports), `-fno-math-errno`,
// `-fopenmp-simd`, floating-point contraction left at the compiler default. The parity library (ref_driver.cpp) keeps
// `-ffp-contract=off`; this one is never used for parity.
//
// Two of the reference's LJ kernels are timed, as md-flexible offers them (examples/md-flexible/src/TypeDefinitions.h:
// 100-130, `functor: lennard-jones` / `lennard-jones-highway`): mdLib::LJFunctor (compiler auto-vectorisation) and
// mdLib::LJFunctorHWY (explicit SIMD through Google Highway, header-only static dispatch). The loop is the force step
// of LogicHandler::computeInteractionsPipeline (LogicHandler.h:1066-1141): on rebuild iterations updateContainer(false),
// halos re-added, rebuildNeighborLists; every iteration functor.initTraversal / computeInteractions / endTraversal.
#include <omp.h>

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "autopas/containers/linkedCells/traversals/LCC18Traversal.h"
#include "autopas/containers/verletClusterLists/VerletClusterLists.h"
#include "autopas/containers/verletClusterLists/traversals/VCLC01BalancedTraversal.h"
#include "autopas/containers/verletClusterLists/traversals/VCLC06Traversal.h"
#include "autopas/containers/verletClusterLists/traversals/VCLClusterIterationTraversal.h"
#include "autopas/utils/Timer.h"
#include "autopas/utils/WrapOpenMP.h"
#include "molecularDynamicsLibrary/LJFunctor.h"
#include "molecularDynamicsLibrary/LJFunctorHWY.h"
#include "molecularDynamicsLibrary/MoleculeLJ.h"

namespace {
using Molecule = mdLib::MoleculeLJ;
using FMCell = autopas::FullParticleCell<Molecule>;
// shift, no mixing, both newton3 modes, globals, no FLOP counting
using FunctorAutoVec = mdLib::LJFunctor<Molecule, true, false, autopas::FunctorN3Modes::Both, true, false>;
using FunctorHWY = mdLib::LJFunctorHWY<Molecule, true, false, autopas::FunctorN3Modes::Both, true, false>;

template <class Functor, class Container, class MakeTraversal>
int benchLoop(Container &container, MakeTraversal makeTraversal, int64_t n, const double *x, const double *y,
              const double *z, const int64_t *own, bool n3, int warmup, int iters, int rebuildFreq, double *out) {
  Functor functor(container.getCutoff());
  functor.setParticleProperties(24., 1.);
  std::vector<Molecule> halos;
  for (int64_t i = 0; i < n; ++i) {
    Molecule m({x[i], y[i], z[i]}, {0., 0., 0.}, static_cast<unsigned long>(i), 0);
    if (own[i] == 1) {
      container.addParticle(m);
    } else if (own[i] == 2) {
      m.setOwnershipState(autopas::OwnershipState::halo);
      halos.push_back(m);
    }
  }
  auto trav = makeTraversal(functor);
  double tRebuild = 0., tCompute = 0.;
  int nRebuild = 0;
  for (int it = -warmup; it < iters; ++it) {
    const bool timed = it >= 0;
    // warm-up iterations rebuild once at their start, timed iterations every rebuildFreq
    if (it == -warmup or (timed and it % rebuildFreq == 0)) {
      autopas::utils::Timer t;
      t.start();
      auto leavers = container.updateContainer(false);
      for (const auto &hp : halos) container.addHaloParticle(hp);
      container.rebuildNeighborLists(trav.get());
      if (timed) {
        tRebuild += static_cast<double>(t.stop()) * 1e-9;
        ++nRebuild;
      }
    }
    autopas::utils::Timer t;
    t.start();
    functor.initTraversal();
    container.computeInteractions(trav.get());
    functor.endTraversal(n3);
    if (timed) tCompute += static_cast<double>(t.stop()) * 1e-9;
  }
  out[0] = tRebuild;
  out[1] = tCompute;
  out[2] = nRebuild;
  out[3] = functor.getPotentialEnergy();
  out[4] = functor.getVirial();
  return 0;
}

template <class Functor>
int run(int64_t n, const double *x, const double *y, const double *z, const int64_t *own, const double *boxMin,
        const double *boxMax, double cutoff, double skin, int container, int traversal, int64_t clusterSize, bool n3,
        int warmup, int iters, int rebuildFreq, double *out) {
  const std::array<double, 3> bmin{boxMin[0], boxMin[1], boxMin[2]}, bmax{boxMax[0], boxMax[1], boxMax[2]};
  const auto layout = autopas::DataLayoutOption::soa;
  if (container == 0) {
    autopas::LinkedCells<Molecule> c(bmin, bmax, cutoff, skin, 1.0);
    const auto info = c.getTraversalSelectorInfo();
    auto mk = [&](Functor &f) -> std::unique_ptr<autopas::TraversalInterface> {
      if (traversal == 0)
        return std::make_unique<autopas::LCC08Traversal<FMCell, Functor>>(info.cellsPerDim, f, info.interactionLength,
                                                                          info.cellLength, layout, n3);
      return std::make_unique<autopas::LCC18Traversal<FMCell, Functor>>(info.cellsPerDim, f, info.interactionLength,
                                                                        info.cellLength, layout, n3);
    };
    return benchLoop<Functor>(c, mk, n, x, y, z, own, n3, warmup, iters, rebuildFreq, out);
  }
  autopas::VerletClusterLists<Molecule> c(bmin, bmax, cutoff, skin, static_cast<size_t>(clusterSize));
  auto mk = [&](Functor &f) -> std::unique_ptr<autopas::TraversalInterface> {
    if (traversal == 0)
      return std::make_unique<autopas::VCLClusterIterationTraversal<FMCell, Functor>>(f, clusterSize, layout, n3);
    if (traversal == 1) return std::make_unique<autopas::VCLC06Traversal<FMCell, Functor>>(f, clusterSize, layout, n3);
    return std::make_unique<autopas::VCLC01BalancedTraversal<Molecule, Functor>>(f, clusterSize, layout, n3);
  };
  return benchLoop<Functor>(c, mk, n, x, y, z, own, n3, warmup, iters, rebuildFreq, out);
}
}  // namespace

extern "C" {
// torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm uses the cores the process may run on
void refb_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}
int refb_num_threads() { return autopas::autopas_get_max_threads(); }
// the vector ISA this library was compiled for (reported in the bench line)
const char *refb_isa() {
#if defined(__AVX512F__)
  return "x86-64-v4 (AVX-512)";
#elif defined(__AVX2__)
  return "x86-64-v3 (AVX2 + FMA)";
#else
  return "x86-64 baseline";
#endif
}

// functor: 0 mdLib::LJFunctor (auto-vectorised), 1 mdLib::LJFunctorHWY (Highway)
// container: 0 LinkedCells (traversal 0 lc_c08, 1 lc_c18), 1 VerletClusterLists (0 cluster_iteration, 1 c06, 2 c01_balanced)
// out: [rebuild s, compute s, number of timed rebuilds, potential energy, virial]
int refb_bench_lj(int64_t n, const double *x, const double *y, const double *z, const int64_t *own, const double *boxMin,
                  const double *boxMax, double cutoff, double skin, int functor, int container, int traversal,
                  int64_t clusterSize, int newton3, int warmup, int iters, int rebuildFreq, double *out) {
  try {
    if (functor == 1)
      return run<FunctorHWY>(n, x, y, z, own, boxMin, boxMax, cutoff, skin, container, traversal, clusterSize,
                             newton3 != 0, warmup, iters, rebuildFreq, out);
    return run<FunctorAutoVec>(n, x, y, z, own, boxMin, boxMax, cutoff, skin, container, traversal, clusterSize,
                               newton3 != 0, warmup, iters, rebuildFreq, out);
  } catch (const std::exception &e) {
    fprintf(stderr, "refb_bench_lj: %s\n", e.what());
    return 1;
  }
}
}
