// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED AutoPas reference for mdLib::LJMultisiteFunctor, compiled in the reference's
// MULTISITE mode (-DMD_FLEXIBLE_MODE=MULTISITE, applicationLibrary/CMakeLists.txt:20-24) into its own library
// oracle/_ref/libautopas_ref_ms.so (the ParticlePropertiesLibrary differs between the two modes).
// LinkedCells + lc_c08 + AoS, like TraversalComparison.cpp:210-214.
#include <array>
#include <cstdint>
#include <vector>

#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "molecularDynamicsLibrary/LJMultisiteFunctor.h"
#include "molecularDynamicsLibrary/MultisiteMoleculeLJ.h"
#include "molecularDynamicsLibrary/ParticlePropertiesLibrary.h"

namespace {
using Mol = mdLib::MultisiteMoleculeLJ;
using Cell = autopas::FullParticleCell<Mol>;
}  // namespace

extern "C" {
// site types: eps[nSiteTypes], sigma[nSiteTypes]; molecule types: siteStart[nMolTypes + 1], sitePos[3 * nSites],
// siteType[nSites]. out: f[3n], torque[3n] by id; globals {Upot, virial}
int ref_multisite(int64_t n, const double *x, const double *y, const double *z, const double *q, const int64_t *molType,
                  const int64_t *own, const double *boxMin, const double *boxMax, double cutoff, double skin,
                  int applyShift, int newton3, int nSiteTypes, const double *eps, const double *sigma, int nMolTypes,
                  const int32_t *siteStart, const double *sitePos, const int32_t *siteType, double *f, double *torque,
                  double *globals) {
  try {
    ParticlePropertiesLibrary<double, size_t> ppl(cutoff);
    for (int t = 0; t < nSiteTypes; ++t) {
      ppl.addSiteType(t, 1.0);
      ppl.addLJParametersToSite(t, eps[t], sigma[t]);
    }
    for (int m = 0; m < nMolTypes; ++m) {
      std::vector<size_t> ids;
      std::vector<std::array<double, 3>> pos;
      for (int s = siteStart[m]; s < siteStart[m + 1]; ++s) {
        ids.push_back(static_cast<size_t>(siteType[s]));
        pos.push_back({sitePos[3 * s], sitePos[3 * s + 1], sitePos[3 * s + 2]});
      }
      ppl.addMolType(m, ids, pos, {1., 1., 1.});
    }
    ppl.calculateMixingCoefficients();
    autopas::LinkedCells<Mol> c({boxMin[0], boxMin[1], boxMin[2]}, {boxMax[0], boxMax[1], boxMax[2]}, cutoff, skin, 1.0);
    for (int64_t i = 0; i < n; ++i) {
      Mol m({x[i], y[i], z[i]}, {0., 0., 0.}, {q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]}, {0., 0., 0.},
            static_cast<unsigned long>(i), static_cast<unsigned long>(molType[i]));
      if (own[i] == 1) {
        c.addParticle(m);
      } else if (own[i] == 2) {
        m.setOwnershipState(autopas::OwnershipState::halo);
        c.addHaloParticle(m);
      }
    }
    const auto info = c.getTraversalSelectorInfo();
    auto run = [&](auto &functor) {
      autopas::LCC08Traversal<Cell, std::remove_reference_t<decltype(functor)>> t(
          info.cellsPerDim, functor, info.interactionLength, info.cellLength, autopas::DataLayoutOption::aos, newton3 != 0);
      c.rebuildNeighborLists(&t);
      functor.initTraversal();
      c.computeInteractions(&t);
      functor.endTraversal(newton3 != 0);
      globals[0] = functor.getPotentialEnergy();
      globals[1] = functor.getVirial();
    };
    if (applyShift) {
      mdLib::LJMultisiteFunctor<Mol, true, true, autopas::FunctorN3Modes::Both, true> functor(cutoff, ppl);
      run(functor);
    } else {
      mdLib::LJMultisiteFunctor<Mol, false, true, autopas::FunctorN3Modes::Both, true> functor(cutoff, ppl);
      run(functor);
    }
    for (auto it = c.begin(autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) {
      const auto id = static_cast<int64_t>(it->getID());
      const auto &F = it->getF();
      const auto &T = it->getTorque();
      for (int d = 0; d < 3; ++d) {
        f[3 * id + d] = F[d];
        torque[3 * id + d] = T[d];
      }
    }
    return 0;
  } catch (const std::exception &) {
    return -1;
  }
}
}
