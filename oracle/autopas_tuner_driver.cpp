// TEST INFRASTRUCTURE ONLY. A stock `autopas::AutoPas<MoleculeLJ>` - the reference's own facade, LogicHandler, AutoTuner and
// tuning manager, compiled from /root/reference where they lie - with the GPU options registered through the header
// overlay of tools/make_autopas_overlay.py (the additive edits of INTEGRATION.md section 2, nothing else).
//
// The allowed search space holds CPU and GPU configurations side by side,
//     {LinkedCells / lc_c08, gpuVerletClusterLists / gpuvcl_pruned, gpuLinkedCells / gpulc_c08} x SoA x newton3 {off, on},
// and the AutoTuner samples all of them through AutoPas::computeInteractions (LogicHandler::computeInteractionsPipeline,
// LogicHandler.h:1258: container switch by ContainerSelector::generateContainer, particles copied over through the
// iterators, TraversalSelector::generateTraversalFromConfig, timing of rebuild + traversal per sample), then settles on
// the fastest. The functor is autopas_b200::GpuLJFunctor (wraps mdLib::LJFunctor: CPU traversals call the reference functor).
// Particles do not move, so every iteration must reproduce the same forces and globals whichever configuration ran:
// the driver prints, per iteration, the configuration, the maximum deviation of the forces from the first CPU sample
// relative to the per-particle force scale, Upot and the virial, and at the end the configuration the tuner chose.
// tests/test_gpu_shim.py runs it on the GPU box and asserts on that report.
#include <cmath>
#include <cstdio>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "GpuContainers.h"
#include "autopas/AutoPasImpl.h"
#include "molecularDynamicsLibrary/MoleculeLJ.h"

using Molecule = mdLib::MoleculeLJ;
using GpuFunctor = autopas_b200::GpuLJFunctor<Molecule, /*shift*/ true, /*mixing*/ false, autopas::FunctorN3Modes::Both,
                                           /*globals*/ true>;

template class autopas::AutoPas<Molecule>;
template bool autopas::AutoPas<Molecule>::computeInteractions(GpuFunctor *);

int main(int argc, char **argv) {
  const int nPerDim = argc > 1 ? std::atoi(argv[1]) : 24;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 60;
  const double spacing = 1.1, cutoff = 2.5, skin = 0.3;
  const double L = nPerDim * spacing;

  autopas::AutoPas<Molecule> autoPas;
  autoPas.setBoxMin({0., 0., 0.});
  autoPas.setBoxMax({L, L, L});
  autoPas.setCutoff(cutoff);
  autoPas.setVerletSkin(skin);
  autoPas.setVerletRebuildFrequency(10);
  autoPas.setVerletClusterSize(32);
  autoPas.setAllowedContainers({autopas::ContainerOption::linkedCells, autopas::ContainerOption::gpuVerletClusterLists,
                                autopas::ContainerOption::gpuLinkedCells});
  autoPas.setAllowedTraversals({autopas::TraversalOption::lc_c08, autopas::TraversalOption::gpuvcl_pruned,
                                autopas::TraversalOption::gpulc_c08});
  autoPas.setAllowedDataLayouts({autopas::DataLayoutOption::soa});
  autoPas.setAllowedNewton3Options({autopas::Newton3Option::disabled, autopas::Newton3Option::enabled});
  autoPas.setAllowedCellSizeFactors(autopas::NumberSetFinite<double>({1.0}));
  autoPas.setTuningStrategyOption({});  // full search
  autoPas.setTuningInterval(1000);
  autoPas.setNumSamples(3);
  autoPas.setOutputSuffix("apb_tuner_test");
  autoPas.init();

  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(-0.12, 0.12);
  size_t id = 0;
  for (int z = 0; z < nPerDim; ++z)
    for (int y = 0; y < nPerDim; ++y)
      for (int x = 0; x < nPerDim; ++x) {
        Molecule m({(x + 0.5) * spacing + u(rng), (y + 0.5) * spacing + u(rng), (z + 0.5) * spacing + u(rng)}, {0., 0., 0.}, id++, 0);
        autoPas.addParticle(m);
      }
  const size_t n = id;

  GpuFunctor functor(cutoff);
  functor.setParticleProperties(24.0, 1.0);

  std::vector<std::array<double, 3>> refF(n);
  std::vector<double> scale(n, 0.);
  bool haveRef = false;
  double refUpot = 0., refVirial = 0.;
  std::set<std::string> sampled;
  std::printf("{\"particles\": %zu, \"iterations\": [\n", n);
  for (int it = 0; it < iterations; ++it) {
    auto leavers = autoPas.updateContainer();
    for (auto p = autoPas.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) p->setF({0., 0., 0.});
    const bool stillTuning = autoPas.computeInteractions(&functor);
    const autopas::Configuration cfg = autoPas.getCurrentConfigs().at(autopas::InteractionTypeOption::pairwise).get();
    const std::string name = cfg.container.to_string() + "/" + cfg.traversal.to_string() + "/" + cfg.newton3.to_string();
    sampled.insert(name);
    const bool isCpu = cfg.container == autopas::ContainerOption::linkedCells;
    double maxRel = 0.;
    if (not haveRef and isCpu) {
      for (auto p = autoPas.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) {
        refF[p->getID()] = p->getF();
        const auto &f = p->getF();
        scale[p->getID()] = std::abs(f[0]) + std::abs(f[1]) + std::abs(f[2]);
      }
      // per-particle yardstick: mean force magnitude of the system (net forces of a liquid cancel)
      double mean = 0.;
      for (double s : scale) mean += s;
      mean /= static_cast<double>(n);
      for (double &s : scale) s = std::max(s, mean) * 30.;  // ~ sum over the ~55 partners of |f_ij|
      refUpot = functor.getPotentialEnergy();
      refVirial = functor.getVirial();
      haveRef = true;
    } else if (haveRef) {
      for (auto p = autoPas.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) {
        const auto &f = p->getF();
        const auto &r = refF[p->getID()];
        const double d = std::max({std::abs(f[0] - r[0]), std::abs(f[1] - r[1]), std::abs(f[2] - r[2])});
        maxRel = std::max(maxRel, d / scale[p->getID()]);
      }
    }
    std::printf("  {\"it\": %d, \"config\": \"%s\", \"tuning\": %s, \"max_rel_force_dev\": %.3e, \"upot\": %.15e, \"virial\": %.15e}%s\n", it,
                name.c_str(), stillTuning ? "true" : "false", haveRef ? maxRel : -1., functor.getPotentialEnergy(),
                functor.getVirial(), it + 1 < iterations ? "," : "");
  }
  const autopas::Configuration chosen = autoPas.getCurrentConfigs().at(autopas::InteractionTypeOption::pairwise).get();
  std::printf("], \"ref_upot\": %.15e, \"ref_virial\": %.15e, \"num_configs_sampled\": %zu, \"chosen\": \"%s/%s/%s\"}\n", refUpot,
              refVirial, sampled.size(), chosen.container.to_string().c_str(), chosen.traversal.to_string().c_str(),
              chosen.newton3.to_string().c_str());
  return 0;
}
