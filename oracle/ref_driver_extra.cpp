// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED AutoPas reference for the functors other than single-site LJ:
// sphLib::SPHCalcDensityFunctor / SPHCalcHydroForceFunctor and mdLib::AxilrodTellerMutoFunctor, run through the
// reference's own LinkedCells container and traversals (lc_c08 AoS for pairwise, lc_c01 AoS newton3-off for triwise =
// the reference configurations of TraversalComparison.cpp:210-220). Part of oracle/_ref/libautopas_ref.so.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <vector>

#include "SPHLibrary/SPHCalcDensityFunctor.h"
#include "SPHLibrary/SPHCalcHydroForceFunctor.h"
#include "SPHLibrary/SPHParticle.h"
#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC01Traversal.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "molecularDynamicsLibrary/AxilrodTellerMutoFunctor.h"
#include "molecularDynamicsLibrary/MoleculeLJ.h"
#include "molecularDynamicsLibrary/ParticlePropertiesLibrary.h"

namespace {
using SPHP = sphLib::SPHParticle;
using SPHCell = autopas::FullParticleCell<SPHP>;
using Molecule = mdLib::MoleculeLJ;
using FMCell = autopas::FullParticleCell<Molecule>;
}  // namespace

namespace {
// timing of the traversal alone (tools/bench_functors.py reports the reference's CPU time next to the GPU kernels): with
// more than one repetition the outputs are accumulated that many times and are only good for timing
int g_timingReps = 1;
double g_lastComputeSeconds = 0.;
template <class Container, class Traversal>
void timedCompute(Container &c, Traversal *t) {
  g_lastComputeSeconds = 1e300;
  for (int r = 0; r < g_timingReps; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    c.computeInteractions(t);
    g_lastComputeSeconds = std::min(g_lastComputeSeconds, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
}
}  // namespace

extern "C" {
void ref_set_timing_reps(int reps) { g_timingReps = reps < 1 ? 1 : reps; }
double ref_last_compute_seconds() { return g_lastComputeSeconds; }
// which: 0 density, 1 hydro force. out (by id): density[n] | acc[3n], engDot[n], vsigmax[n]
int ref_sph(int64_t n, const double *x, const double *y, const double *z, const double *vx, const double *vy,
            const double *vz, const double *mass, const double *smth, const double *density, const double *pressure,
            const double *snd, const int64_t *own, const double *boxMin, const double *boxMax, double cutoff,
            double skin, int which, int newton3, double *outDensity, double *outAcc, double *outEngDot,
            double *outVsigmax) {
  try {
    autopas::LinkedCells<SPHP> c({boxMin[0], boxMin[1], boxMin[2]}, {boxMax[0], boxMax[1], boxMax[2]}, cutoff, skin, 1.0);
    for (int64_t i = 0; i < n; ++i) {
      SPHP p({x[i], y[i], z[i]}, {vx[i], vy[i], vz[i]}, static_cast<unsigned long>(i), mass[i], smth[i], snd[i]);
      p.setDensity(density[i]);
      p.setPressure(pressure[i]);
      if (own[i] == 1) {
        c.addParticle(p);
      } else if (own[i] == 2) {
        p.setOwnershipState(autopas::OwnershipState::halo);
        c.addHaloParticle(p);
      }
    }
    const auto info = c.getTraversalSelectorInfo();
    if (which == 0) {
      sphLib::SPHCalcDensityFunctor<SPHP> f;
      autopas::LCC08Traversal<SPHCell, decltype(f)> t(info.cellsPerDim, f, info.interactionLength, info.cellLength,
                                                       autopas::DataLayoutOption::aos, newton3 != 0);
      c.rebuildNeighborLists(&t);
      f.initTraversal();
      timedCompute(c, &t);
      f.endTraversal(newton3 != 0);
    } else {
      sphLib::SPHCalcHydroForceFunctor<SPHP> f;
      autopas::LCC08Traversal<SPHCell, decltype(f)> t(info.cellsPerDim, f, info.interactionLength, info.cellLength,
                                                       autopas::DataLayoutOption::aos, newton3 != 0);
      c.rebuildNeighborLists(&t);
      f.initTraversal();
      timedCompute(c, &t);
      f.endTraversal(newton3 != 0);
    }
    for (auto it = c.begin(autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) {
      const auto id = static_cast<int64_t>(it->getID());
      outDensity[id] = it->getDensity();
      const auto a = it->getAcceleration();
      outAcc[3 * id] = a[0];
      outAcc[3 * id + 1] = a[1];
      outAcc[3 * id + 2] = a[2];
      outEngDot[id] = it->getEngDot();
      outVsigmax[id] = it->getVSigMax();
    }
    return 0;
  } catch (const std::exception &) {
    return -1;
  }
}

// Axilrod-Teller-Muto through LinkedCells + lc_c01, AoS, newton3 off. nuOfType: per type (mixing) or nullptr (nu).
// globals: {Upot, virial}; counters: {numFLOPs}
int ref_atm(int64_t n, const double *x, const double *y, const double *z, const int64_t *type, const int64_t *own,
            const double *boxMin, const double *boxMax, double cutoff, double skin, double nu, int ntypes,
            const double *nuOfType, double *f, double *globals, uint64_t *flops) {
  try {
    autopas::LinkedCells<Molecule> c({boxMin[0], boxMin[1], boxMin[2]}, {boxMax[0], boxMax[1], boxMax[2]}, cutoff, skin, 1.0);
    for (int64_t i = 0; i < n; ++i) {
      Molecule m({x[i], y[i], z[i]}, {0., 0., 0.}, static_cast<unsigned long>(i), static_cast<unsigned long>(type ? type[i] : 0));
      if (own[i] == 1) {
        c.addParticle(m);
      } else if (own[i] == 2) {
        m.setOwnershipState(autopas::OwnershipState::halo);
        c.addHaloParticle(m);
      }
    }
    const auto info = c.getTraversalSelectorInfo();
    auto run = [&](auto &functor) {
      autopas::LCC01Traversal<FMCell, std::remove_reference_t<decltype(functor)>> t(
          info.cellsPerDim, functor, info.interactionLength, info.cellLength, autopas::DataLayoutOption::aos, false);
      c.rebuildNeighborLists(&t);
      functor.initTraversal();
      timedCompute(c, &t);
      functor.endTraversal(false);
      globals[0] = functor.getPotentialEnergy();
      globals[1] = functor.getVirial();
      *flops = functor.getNumFLOPs();
    };
    if (nuOfType != nullptr) {
      ParticlePropertiesLibrary<double, size_t> ppl(cutoff);
      for (int t = 0; t < ntypes; ++t) {
        ppl.addSiteType(t, 1.0);
        ppl.addATMParametersToSite(t, nuOfType[t]);
      }
      ppl.calculateMixingCoefficients();
      mdLib::AxilrodTellerMutoFunctor<Molecule, true, autopas::FunctorN3Modes::Both, true, true> functor(cutoff, ppl);
      run(functor);
    } else {
      mdLib::AxilrodTellerMutoFunctor<Molecule, false, autopas::FunctorN3Modes::Both, true, true> functor(cutoff);
      functor.setParticleProperties(nu);
      run(functor);
    }
    for (auto it = c.begin(autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) {
      const auto id = static_cast<int64_t>(it->getID());
      const auto &F = it->getF();
      f[3 * id] = F[0];
      f[3 * id + 1] = F[1];
      f[3 * id + 2] = F[2];
    }
    return 0;
  } catch (const std::exception &) {
    return -1;
  }
}
}
