// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Thin extern "C" driver around the UNMODIFIED AutoPas reference headers (compiled from /root/reference where they
// lie; nothing is copied). It follows the flow of the reference's own parity harness
// (tests/testAutopas/tests/containers/TraversalComparison.cpp:135-199): build container -> fill -> rebuildNeighborLists
// -> functor.initTraversal(); container->computeInteractions(traversal); functor.endTraversal(n3) -> read forces by id.
// Built by oracle/Makefile into oracle/_ref/libautopas_ref.so (git-ignored, travels to the GPU box).
#include <array>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "autopas/containers/linkedCells/traversals/LCC18Traversal.h"
#include "autopas/containers/verletClusterLists/VerletClusterLists.h"
#include "autopas/containers/verletClusterLists/traversals/VCLC01BalancedTraversal.h"
#include "autopas/containers/verletClusterLists/traversals/VCLC06Traversal.h"
#include "autopas/containers/verletClusterLists/traversals/VCLClusterIterationTraversal.h"
#include "autopas/utils/WrapOpenMP.h"
#include "molecularDynamicsLibrary/LJFunctor.h"
#include "molecularDynamicsLibrary/MoleculeLJ.h"
#include "molecularDynamicsLibrary/ParticlePropertiesLibrary.h"

namespace {
using Molecule = mdLib::MoleculeLJ;
using FMCell = autopas::FullParticleCell<Molecule>;

enum Flags : int { kShift = 1, kMixing = 2, kNewton3 = 4, kSoA = 8 };

struct VclDump {
  std::vector<int64_t> towerOfParticle;   // by id, -1 if not stored
  std::vector<int64_t> clusterParticles;  // numClusters * M ids, -1 for dummies
  std::vector<int64_t> clusterTower;      // numClusters
  std::vector<int64_t> pairs;             // 2 * numPairs (global cluster idx A, B)
  std::array<int64_t, 2> towersPerDim{};
  std::array<double, 2> towerSide{};
  int64_t clusterSize = 0;
} g_dump;

template <class Functor>
void collect(Functor &functor, bool n3, autopas::ParticleContainerInterface<Molecule> &container, int64_t n, double *f,
             double *globals, uint64_t *flops, double *hitRate) {
  for (auto it = container.begin(autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) {
    const auto id = static_cast<int64_t>(it->getID());
    if (id < 0 or id >= n) continue;
    const auto &F = it->getF();
    f[3 * id + 0] = F[0];
    f[3 * id + 1] = F[1];
    f[3 * id + 2] = F[2];
  }
  globals[0] = functor.getPotentialEnergy();
  globals[1] = functor.getVirial();
  *flops = functor.getNumFLOPs();
  *hitRate = functor.getHitRate();
}

template <class Container>
void fill(Container &container, int64_t n, const double *x, const double *y, const double *z, const int64_t *type,
          const int64_t *own) {
  for (int64_t i = 0; i < n; ++i) {
    Molecule m({x[i], y[i], z[i]}, {0., 0., 0.}, static_cast<unsigned long>(i),
               static_cast<unsigned long>(type ? type[i] : 0));
    if (own[i] == 1) {
      container.addParticle(m);
    } else if (own[i] == 2) {
      m.setOwnershipState(autopas::OwnershipState::halo);  // LogicHandler::addHaloParticle does this (LogicHandler.h:379-389)
      container.addHaloParticle(m);
    }
  }
}

template <bool shift, bool mixing>
struct FunctorFactory {
  using Functor = mdLib::LJFunctor<Molecule, shift, mixing, autopas::FunctorN3Modes::Both, /*globals*/ true,
                                   /*countFLOPs*/ true>;
  std::unique_ptr<ParticlePropertiesLibrary<double, size_t>> ppl;
  std::unique_ptr<Functor> make(double cutoff, int ntypes, const double *eps, const double *sigma) {
    if constexpr (mixing) {
      ppl = std::make_unique<ParticlePropertiesLibrary<double, size_t>>(cutoff);
      for (int t = 0; t < ntypes; ++t) {
        ppl->addSiteType(t, 1.0);
        ppl->addLJParametersToSite(t, eps[t], sigma[t]);
      }
      ppl->calculateMixingCoefficients();
      return std::make_unique<Functor>(cutoff, *ppl);
    } else {
      auto fun = std::make_unique<Functor>(cutoff);
      fun->setParticleProperties(24. * eps[0], sigma[0] * sigma[0]);
      return fun;
    }
  }
};

template <bool shift, bool mixing>
int runLC(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
          const double *boxMin, const double *boxMax, double cutoff, double skin, double csf, int flags, int traversal,
          int ntypes, const double *eps, const double *sigma, double *f, double *globals, uint64_t *flops,
          double *hitRate, int64_t *cellOfParticle, int64_t *cellsPerDim) {
  FunctorFactory<shift, mixing> factory;
  auto functor = factory.make(cutoff, ntypes, eps, sigma);
  using Functor = typename FunctorFactory<shift, mixing>::Functor;
  const std::array<double, 3> bmin{boxMin[0], boxMin[1], boxMin[2]}, bmax{boxMax[0], boxMax[1], boxMax[2]};
  autopas::LinkedCells<Molecule> container(bmin, bmax, cutoff, skin, csf,
                                           /*sortingThreshold*/ std::numeric_limits<size_t>::max());
  fill(container, n, x, y, z, type, own);
  const bool n3 = flags & kNewton3;
  const auto layout = (flags & kSoA) ? autopas::DataLayoutOption::soa : autopas::DataLayoutOption::aos;
  const auto info = container.getTraversalSelectorInfo();
  std::unique_ptr<autopas::TraversalInterface> trav;
  if (traversal == 0) {
    trav = std::make_unique<autopas::LCC08Traversal<FMCell, Functor>>(info.cellsPerDim, *functor,
                                                                      info.interactionLength, info.cellLength, layout, n3);
  } else {
    trav = std::make_unique<autopas::LCC18Traversal<FMCell, Functor>>(info.cellsPerDim, *functor,
                                                                      info.interactionLength, info.cellLength, layout, n3);
  }
  container.rebuildNeighborLists(trav.get());
  functor->initTraversal();
  container.computeInteractions(trav.get());
  functor->endTraversal(n3);
  std::memset(f, 0, sizeof(double) * 3 * n);
  collect(*functor, n3, container, n, f, globals, flops, hitRate);
  if (cellOfParticle) {
    for (int64_t i = 0; i < n; ++i) {
      cellOfParticle[i] = static_cast<int64_t>(container.getCellBlock().get1DIndexOfPosition({x[i], y[i], z[i]}));
    }
  }
  if (cellsPerDim) {
    for (int d = 0; d < 3; ++d) cellsPerDim[d] = static_cast<int64_t>(info.cellsPerDim[d]);
  }
  return 0;
}

template <bool shift, bool mixing>
int runVCL(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
           const double *boxMin, const double *boxMax, double cutoff, double skin, int64_t clusterSize, int flags,
           int traversal, int ntypes, const double *eps, const double *sigma, double *f, double *globals,
           uint64_t *flops, double *hitRate) {
  FunctorFactory<shift, mixing> factory;
  auto functor = factory.make(cutoff, ntypes, eps, sigma);
  using Functor = typename FunctorFactory<shift, mixing>::Functor;
  const std::array<double, 3> bmin{boxMin[0], boxMin[1], boxMin[2]}, bmax{boxMax[0], boxMax[1], boxMax[2]};
  autopas::VerletClusterLists<Molecule> container(bmin, bmax, cutoff, skin, static_cast<size_t>(clusterSize));
  fill(container, n, x, y, z, type, own);
  const bool n3 = flags & kNewton3;
  const auto layout = (flags & kSoA) ? autopas::DataLayoutOption::soa : autopas::DataLayoutOption::aos;
  std::unique_ptr<autopas::TraversalInterface> trav;
  if (traversal == 0) {
    trav = std::make_unique<autopas::VCLClusterIterationTraversal<FMCell, Functor>>(*functor, clusterSize, layout, n3);
  } else if (traversal == 1) {
    trav = std::make_unique<autopas::VCLC06Traversal<FMCell, Functor>>(*functor, clusterSize, layout, n3);
  } else {
    trav = std::make_unique<autopas::VCLC01BalancedTraversal<Molecule, Functor>>(*functor, clusterSize, layout, n3);
  }
  container.rebuildNeighborLists(trav.get());

  // dump structure (tower membership, clusters, cluster pairs) for set-level parity checks
  auto &block = container.getTowerBlock();
  g_dump = VclDump{};
  g_dump.clusterSize = clusterSize;
  g_dump.towersPerDim = {static_cast<int64_t>(block.getTowersPerDim()[0]),
                         static_cast<int64_t>(block.getTowersPerDim()[1])};
  g_dump.towerSide = {block.getTowerSideLength()[0], block.getTowerSideLength()[1]};
  g_dump.towerOfParticle.assign(n, -1);
  std::vector<const void *> clusterAddr;
  std::vector<size_t> towerFirstCluster(block.size() + 1, 0);
  for (size_t t = 0; t < block.size(); ++t) {
    auto &tower = block[t];
    towerFirstCluster[t] = g_dump.clusterTower.size();
    for (size_t c = 0; c < tower.getNumClusters(); ++c) {
      auto &cluster = tower.getCluster(c);
      clusterAddr.push_back(&cluster);
      g_dump.clusterTower.push_back(static_cast<int64_t>(t));
      for (size_t k = 0; k < static_cast<size_t>(clusterSize); ++k) {
        const auto &p = cluster[k];
        if (p.isDummy()) {
          g_dump.clusterParticles.push_back(-1);
        } else {
          g_dump.clusterParticles.push_back(static_cast<int64_t>(p.getID()));
          g_dump.towerOfParticle[p.getID()] = static_cast<int64_t>(t);
        }
      }
    }
  }
  towerFirstCluster[block.size()] = g_dump.clusterTower.size();
  {
    // map cluster address -> global index (clusters of one tower are contiguous in a std::vector)
    size_t g = 0;
    for (size_t t = 0; t < block.size(); ++t) {
      auto &tower = block[t];
      for (size_t c = 0; c < tower.getNumClusters(); ++c, ++g) {
        auto *nbrs = tower.getCluster(c).getNeighbors();
        if (not nbrs) continue;
        for (auto *nb : *nbrs) {
          if (not nb) continue;
          // find tower of nb by address range
          for (size_t t2 = 0; t2 < block.size(); ++t2) {
            auto &tw = block[t2];
            if (tw.getNumClusters() == 0) continue;
            const auto *first = &tw.getCluster(0);
            const auto *last = first + tw.getNumClusters();
            if (nb >= first and nb < last) {
              g_dump.pairs.push_back(static_cast<int64_t>(g));
              g_dump.pairs.push_back(static_cast<int64_t>(towerFirstCluster[t2] + (nb - first)));
              break;
            }
          }
        }
      }
    }
  }

  functor->initTraversal();
  container.computeInteractions(trav.get());
  functor->endTraversal(n3);
  std::memset(f, 0, sizeof(double) * 3 * n);
  collect(*functor, n3, container, n, f, globals, flops, hitRate);
  return 0;
}
}  // namespace

// Timing of the reference's own force step (bench.py --impl reference / cpu_baseline): container maintenance as
// LogicHandler does it on rebuild iterations (updateContainer(false) -> halos re-added -> rebuildNeighborLists,
// LogicHandler.h:165-218, 1095-1108) and functor.initTraversal / computeInteractions / endTraversal every iteration
// (:1092-1122), OpenMP over all host threads. LJFunctor with shift and globals, no FLOP counting.
template <class Container, class MakeTraversal>
int benchLoop(Container &container, MakeTraversal makeTraversal, int64_t n, const double *x, const double *y,
              const double *z, const int64_t *own, bool n3, int iters, int rebuildFreq, double *seconds) {
  using Functor = mdLib::LJFunctor<Molecule, true, false, autopas::FunctorN3Modes::Both, true, false>;
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
  for (int it = 0; it < iters; ++it) {
    if (it % rebuildFreq == 0) {
      autopas::utils::Timer t;
      t.start();
      auto leavers = container.updateContainer(false);
      for (const auto &hp : halos) container.addHaloParticle(hp);
      container.rebuildNeighborLists(trav.get());
      tRebuild += static_cast<double>(t.stop()) * 1e-9;
      ++nRebuild;
    }
    autopas::utils::Timer t;
    t.start();
    functor.initTraversal();
    container.computeInteractions(trav.get());
    functor.endTraversal(n3);
    tCompute += static_cast<double>(t.stop()) * 1e-9;
  }
  seconds[0] = tRebuild;
  seconds[1] = tCompute;
  seconds[2] = nRebuild;
  seconds[3] = functor.getPotentialEnergy();
  return 0;
}

#define DISPATCH(fn, ...)                                    \
  do {                                                        \
    const bool s = flags & kShift, m = flags & kMixing;       \
    if (s and m) return fn<true, true>(__VA_ARGS__);          \
    if (s and not m) return fn<true, false>(__VA_ARGS__);     \
    if (not s and m) return fn<false, true>(__VA_ARGS__);     \
    return fn<false, false>(__VA_ARGS__);                     \
  } while (0)

extern "C" {
int ref_num_threads() { return autopas::autopas_get_max_threads(); }

int ref_lj_linkedcells(int64_t n, const double *x, const double *y, const double *z, const int64_t *type,
                       const int64_t *own, const double *boxMin, const double *boxMax, double cutoff, double skin,
                       double csf, int flags, int traversal, int ntypes, const double *eps, const double *sigma,
                       double *f, double *globals, uint64_t *flops, double *hitRate, int64_t *cellOfParticle,
                       int64_t *cellsPerDim) {
  try {
    DISPATCH(runLC, n, x, y, z, type, own, boxMin, boxMax, cutoff, skin, csf, flags, traversal, ntypes, eps, sigma, f,
             globals, flops, hitRate, cellOfParticle, cellsPerDim);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_lj_linkedcells: %s\n", e.what());
    return 1;
  }
}

int ref_lj_vcl(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
               const double *boxMin, const double *boxMax, double cutoff, double skin, int64_t clusterSize, int flags,
               int traversal, int ntypes, const double *eps, const double *sigma, double *f, double *globals,
               uint64_t *flops, double *hitRate) {
  try {
    DISPATCH(runVCL, n, x, y, z, type, own, boxMin, boxMax, cutoff, skin, clusterSize, flags, traversal, ntypes, eps,
             sigma, f, globals, flops, hitRate);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_lj_vcl: %s\n", e.what());
    return 1;
  }
}

// container: 0 LinkedCells (traversal 0 lc_c08, 1 lc_c18), 1 VerletClusterLists (0 cluster_iteration, 1 c06, 2 c01_balanced)
// seconds: [total rebuild s, total compute s, number of rebuilds, potential energy of the last iteration]
int ref_bench_lj(int64_t n, const double *x, const double *y, const double *z, const int64_t *own, const double *boxMin,
                 const double *boxMax, double cutoff, double skin, int container, int traversal, int64_t clusterSize,
                 int newton3, int iters, int rebuildFreq, double *seconds) {
  try {
    using Functor = mdLib::LJFunctor<Molecule, true, false, autopas::FunctorN3Modes::Both, true, false>;
    const std::array<double, 3> bmin{boxMin[0], boxMin[1], boxMin[2]}, bmax{boxMax[0], boxMax[1], boxMax[2]};
    const bool n3 = newton3 != 0;
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
      return benchLoop(c, mk, n, x, y, z, own, n3, iters, rebuildFreq, seconds);
    }
    autopas::VerletClusterLists<Molecule> c(bmin, bmax, cutoff, skin, static_cast<size_t>(clusterSize));
    auto mk = [&](Functor &f) -> std::unique_ptr<autopas::TraversalInterface> {
      if (traversal == 0)
        return std::make_unique<autopas::VCLClusterIterationTraversal<FMCell, Functor>>(f, clusterSize, layout, n3);
      if (traversal == 1) return std::make_unique<autopas::VCLC06Traversal<FMCell, Functor>>(f, clusterSize, layout, n3);
      return std::make_unique<autopas::VCLC01BalancedTraversal<Molecule, Functor>>(f, clusterSize, layout, n3);
    };
    return benchLoop(c, mk, n, x, y, z, own, n3, iters, rebuildFreq, seconds);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_bench_lj: %s\n", e.what());
    return 1;
  }
}

// sizes: [numClusters, numPairs, towersX, towersY, clusterSize]
void ref_vcl_dump_sizes(int64_t *sizes, double *towerSide) {
  sizes[0] = static_cast<int64_t>(g_dump.clusterTower.size());
  sizes[1] = static_cast<int64_t>(g_dump.pairs.size() / 2);
  sizes[2] = g_dump.towersPerDim[0];
  sizes[3] = g_dump.towersPerDim[1];
  sizes[4] = g_dump.clusterSize;
  towerSide[0] = g_dump.towerSide[0];
  towerSide[1] = g_dump.towerSide[1];
}
void ref_vcl_dump_copy(int64_t *towerOfParticle, int64_t *clusterParticles, int64_t *clusterTower, int64_t *pairs) {
  std::memcpy(towerOfParticle, g_dump.towerOfParticle.data(), g_dump.towerOfParticle.size() * 8);
  std::memcpy(clusterParticles, g_dump.clusterParticles.data(), g_dump.clusterParticles.size() * 8);
  std::memcpy(clusterTower, g_dump.clusterTower.data(), g_dump.clusterTower.size() * 8);
  std::memcpy(pairs, g_dump.pairs.data(), g_dump.pairs.size() * 8);
}
}
