// Drop-in check of GpuLJMultisiteFunctor (autopas_b200/shim/GpuContainers.h) against the UNMODIFIED AutoPas headers in
// the reference's MULTISITE build mode (-DMD_FLEXIBLE_MODE=MULTISITE, applicationLibrary/CMakeLists.txt:20-24; compiled by
// oracle/Makefile, target `shimtest`; runs on the GPU box through tests/test_gpu_shim.py).
// Reference configuration as in TraversalComparison.cpp:210-214: LinkedCells, lc_c08, AoS, newton3, with
// mdLib::LJMultisiteFunctor<shift, mixing, globals>; GPU: gpuLinkedCells + gpulc_c08 / gpulc_c18 through the same
// container / traversal / functor interfaces, forces and torques read back through the container iterators.
#include <cmath>
#include <cstdio>
#include <map>
#include <random>

#include "GpuContainers.h"
#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "molecularDynamicsLibrary/MultisiteMoleculeLJ.h"

using Mol = mdLib::MultisiteMoleculeLJ;
using Cell = autopas::FullParticleCell<Mol>;

static int g_fail = 0;
#define CHECK(cond, ...)                               \
  do {                                                 \
    if (!(cond)) {                                     \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                        \
      std::printf("\n");                               \
      ++g_fail;                                        \
    }                                                  \
  } while (0)

template <bool n3>
static void compareMultisite(int gpuTraversal) {
  const std::array<double, 3> boxMin{0., 0., 0.}, boxMax{10., 10., 10.};
  const double cutoff = 2.5, skin = 0.3, il = cutoff + skin;
  ParticlePropertiesLibrary<double, size_t> ppl(cutoff);
  ppl.addSiteType(0, 1.0);
  ppl.addLJParametersToSite(0, 1.0, 1.0);
  ppl.addSiteType(1, 1.5);
  ppl.addLJParametersToSite(1, 0.7, 0.9);
  ppl.addMolType(0, {0, 1}, {{{0.12, 0., 0.}, {-0.12, 0., 0.}}}, {1., 1., 1.});
  ppl.addMolType(1, {0, 1, 1}, {{{0., 0.1, 0.}, {0.08, -0.06, 0.}, {-0.08, -0.06, 0.05}}}, {1., 1., 1.});
  ppl.calculateMixingCoefficients();

  std::mt19937_64 rng(123);
  std::uniform_real_distribution<double> u(-0.15, 0.15), g(-1., 1.);
  std::vector<Mol> owned, halo;
  size_t idOwned = 0, idHalo = 1000000;
  for (int iz = -3; iz < 13; ++iz)
    for (int iy = -3; iy < 13; ++iy)
      for (int ix = -3; ix < 13; ++ix) {
        const std::array<double, 3> r{ix + 0.5 + u(rng), iy + 0.5 + u(rng), iz + 0.5 + u(rng)};
        std::array<double, 4> q{g(rng), g(rng), g(rng), g(rng)};
        const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (auto &v : q) v /= qn;
        const bool inHaloBox = r[0] >= -il and r[0] < 10. + il and r[1] >= -il and r[1] < 10. + il and r[2] >= -il and r[2] < 10. + il;
        if (autopas::utils::inBox(r, boxMin, boxMax)) {
          owned.emplace_back(r, std::array<double, 3>{u(rng), u(rng), u(rng)}, q, std::array<double, 3>{g(rng), g(rng), g(rng)},
                             idOwned, idOwned % 2);
          ++idOwned;
        } else if (inHaloBox) {
          Mol m(r, {0., 0., 0.}, q, {0., 0., 0.}, idHalo, idHalo % 2);
          ++idHalo;
          m.setOwnershipState(autopas::OwnershipState::halo);
          halo.push_back(m);
        }
      }
  autopas::LinkedCells<Mol> ref(boxMin, boxMax, cutoff, skin, 1.0);
  autopas_b200::GpuParticleContainer<Mol> gpu(APB_CONTAINER_LINKED_CELLS, boxMin, boxMax, cutoff, skin, 1.0, 4);
  autopas::ParticleContainerInterface<Mol> &c = gpu;
  for (const auto &p : owned) ref.addParticle(p), c.addParticle(p);
  for (const auto &p : halo) ref.addHaloParticle(p), c.addHaloParticle(p);

  using RefFunctor = mdLib::LJMultisiteFunctor<Mol, true, true, autopas::FunctorN3Modes::Both, true>;
  RefFunctor fr(cutoff, ppl);
  const auto info = ref.getTraversalSelectorInfo();
  autopas::LCC08Traversal<Cell, RefFunctor> tr(info.cellsPerDim, fr, info.interactionLength, info.cellLength,
                                               autopas::DataLayoutOption::aos, true);
  ref.rebuildNeighborLists(&tr);
  fr.initTraversal();
  ref.computeInteractions(&tr);
  fr.endTraversal(true);

  using GpuFunctor = autopas_b200::GpuLJMultisiteFunctor<Mol, true, true, autopas::FunctorN3Modes::Both, true>;
  GpuFunctor fg(cutoff, ppl);
  autopas_b200::GpuTraversal<GpuFunctor> tg(gpuTraversal, fg, n3);
  CHECK(tg.isApplicableToDomain(), "multi-site gpulc applicable");
  autopas_b200::GpuTraversal<GpuFunctor> tv(APB_TRAVERSAL_GPUVCL_C06, fg, n3);
  CHECK(!tv.isApplicableToDomain(), "the multi-site functor has no kernel for the cluster-list traversals");
  c.rebuildNeighborLists(&tg);
  fg.initTraversal();
  c.computeInteractions(&tg);
  fg.endTraversal(n3);

  struct Out {
    std::array<double, 3> f, t, w;
  };
  std::map<size_t, Out> outRef;
  double fmax = 0., tmax = 0.;
  for (auto it = ref.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) {
    outRef[it->getID()] = {it->getF(), it->getTorque(), it->getAngularVel()};
    for (int d = 0; d < 3; ++d) {
      fmax = std::max(fmax, std::fabs(it->getF()[d]));
      tmax = std::max(tmax, std::fabs(it->getTorque()[d]));
    }
  }
  double maxF = 0., maxT = 0.;
  size_t seen = 0;
  bool angularKept = true;
  for (auto it = c.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it, ++seen) {
    const Out &o = outRef.at(it->getID());
    for (int d = 0; d < 3; ++d) {
      maxF = std::max(maxF, std::fabs(it->getF()[d] - o.f[d]) / fmax);
      maxT = std::max(maxT, std::fabs(it->getTorque()[d] - o.t[d]) / tmax);
    }
    angularKept = angularKept and it->getAngularVel() == o.w;
  }
  CHECK(seen == owned.size(), "multi-site iterator visits %zu of %zu owned", seen, owned.size());
  CHECK(maxF <= 1e-12, "multi-site force mismatch %.3e (relative to max |F| = %.3e)", maxF, fmax);
  CHECK(maxT <= 1e-12, "multi-site torque mismatch %.3e (relative to max |T| = %.3e)", maxT, tmax);
  CHECK(angularKept, "angular velocity (no device column) must survive the device round trip");
  const double u0 = fr.getPotentialEnergy(), u1 = fg.getPotentialEnergy();
  const double v0 = fr.getVirial(), v1 = fg.getVirial();
  CHECK(std::fabs(u1 - u0) <= 1e-12 * std::fabs(u0), "multi-site Upot %.17g vs %.17g", u1, u0);
  CHECK(std::fabs(v1 - v0) <= 1e-12 * std::fabs(v0), "multi-site virial %.17g vs %.17g", v1, v0);
  std::printf("gpuLinkedCells/%s LJMultisiteFunctor newton3=%d  max |dF|/max|F| = %.2e  max |dT|/max|T| = %.2e  Upot %.12e  virial %.12e\n",
              autopas_b200::traversalName(gpuTraversal), int(n3), maxF, maxT, u1, v1);
}

int main() {
  autopas::utils::ExceptionHandler::setBehavior(autopas::utils::ExceptionBehavior::throwException);
  try {
    compareMultisite<true>(APB_TRAVERSAL_GPULC_C08);
    compareMultisite<false>(APB_TRAVERSAL_GPULC_C18);
  } catch (const std::exception &e) {
    std::printf("FAIL: exception %s\n", e.what());
    ++g_fail;
  }
  std::printf(g_fail ? "SHIM TEST FAILED (%d)\n" : "SHIM TEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
