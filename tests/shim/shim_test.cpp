// Drop-in check of autopas_b200/shim/GpuContainers.h against the UNMODIFIED AutoPas headers (compiled where they lie by
// oracle/Makefile, target `shimtest`; the binary runs on the GPU box through tests/test_gpu_shim.py).
// Same flow as the reference's own parity harness (tests/testAutopas/tests/containers/TraversalComparison.cpp:135-199):
// container -> fill -> rebuildNeighborLists -> functor.initTraversal(); computeInteractions(traversal);
// functor.endTraversal(n3) -> read forces through the container iterators. Reference configuration:
// LinkedCells, csf 1, lc_c08, AoS, newton3 (:210-214).
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iterator>
#include <sstream>
#include <map>
#include <random>
#include <set>

#include "GpuContainers.h"
#include "autopas/containers/linkedCells/LinkedCells.h"
#include "autopas/containers/linkedCells/traversals/LCC01Traversal.h"
#include "autopas/containers/linkedCells/traversals/LCC08Traversal.h"
#include "molecularDynamicsLibrary/LJFunctor.h"
#include "molecularDynamicsLibrary/MoleculeLJ.h"
#include "SPHLibrary/SPHParticle.h"

using Molecule = mdLib::MoleculeLJ;
using FMCell = autopas::FullParticleCell<Molecule>;

static int g_fail = 0;
#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);     \
      std::printf(__VA_ARGS__);                            \
      std::printf("\n");                                   \
      ++g_fail;                                            \
    }                                                      \
  } while (0)

struct Scenario {
  std::array<double, 3> boxMin{0., 0., 0.}, boxMax{12., 12., 12.};
  double cutoff = 2.5, skin = 0.3;
  std::vector<Molecule> owned, halo;
};

// jittered simple-cubic lattice (spacing 1, jitter 0.2) over the box and its halo shell: liquid-like distances, no
// near contacts, so that the force comparison is meaningful per particle
static Scenario makeScenario(unsigned seed) {
  Scenario s;
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> u(-0.2, 0.2);
  const double il = s.cutoff + s.skin;
  size_t idOwned = 0, idHalo = 1000000;
  for (int iz = -3; iz < 15; ++iz)
    for (int iy = -3; iy < 15; ++iy)
      for (int ix = -3; ix < 15; ++ix) {
        const std::array<double, 3> r{ix + 0.5 + u(rng), iy + 0.5 + u(rng), iz + 0.5 + u(rng)};
        const bool inHaloBox = r[0] >= -il and r[0] < 12. + il and r[1] >= -il and r[1] < 12. + il and r[2] >= -il and r[2] < 12. + il;
        if (autopas::utils::inBox(r, s.boxMin, s.boxMax)) {
          s.owned.emplace_back(r, std::array<double, 3>{u(rng), u(rng), u(rng)}, idOwned++, 0);
        } else if (inHaloBox) {
          Molecule m(r, {0., 0., 0.}, idHalo++, 0);
          m.setOwnershipState(autopas::OwnershipState::halo);
          s.halo.push_back(m);
        }
      }
  return s;
}

template <class Container>
static void fill(Container &c, const Scenario &s) {
  for (const auto &p : s.owned) c.addParticle(p);
  for (const auto &p : s.halo) c.addHaloParticle(p);
}

template <class Container>
static std::map<size_t, std::array<double, 3>> forcesOfOwned(Container &c) {
  std::map<size_t, std::array<double, 3>> f;
  for (auto it = c.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) f[it->getID()] = it->getF();
  return f;
}

template <bool n3>
static void compare(const Scenario &s, int gpuContainer, int gpuTraversal, unsigned clusterSize, const char *name) {
  constexpr bool shift = true, globals = true;
  // reference
  autopas::LinkedCells<Molecule> ref(s.boxMin, s.boxMax, s.cutoff, s.skin, 1.0);
  fill(ref, s);
  using RefFunctor = mdLib::LJFunctor<Molecule, shift, false, autopas::FunctorN3Modes::Both, globals, true>;
  RefFunctor fr(s.cutoff);
  fr.setParticleProperties(24.0, 1.0);
  const auto info = ref.getTraversalSelectorInfo();
  autopas::LCC08Traversal<FMCell, RefFunctor> tr(info.cellsPerDim, fr, info.interactionLength, info.cellLength,
                                                 autopas::DataLayoutOption::aos, true);
  ref.rebuildNeighborLists(&tr);
  fr.initTraversal();
  ref.computeInteractions(&tr);
  fr.endTraversal(true);
  // GPU behind the same interface
  autopas_b200::GpuParticleContainer<Molecule> gpu(gpuContainer, s.boxMin, s.boxMax, s.cutoff, s.skin, 1.0, clusterSize);
  autopas::ParticleContainerInterface<Molecule> &c = gpu;
  fill(c, s);
  CHECK(c.getNumberOfParticles(autopas::IteratorBehavior::owned) == s.owned.size(), "%s owned count", name);
  CHECK(c.getNumberOfParticles(autopas::IteratorBehavior::halo) == s.halo.size(), "%s halo count", name);
  using GpuFunctor = autopas_b200::GpuLJFunctor<Molecule, shift, false, autopas::FunctorN3Modes::Both, globals, true>;
  GpuFunctor fg(s.cutoff);
  fg.setParticleProperties(24.0, 1.0);
  autopas_b200::GpuTraversal<GpuFunctor> tg(gpuTraversal, fg, n3);
  CHECK(tg.isApplicableToDomain(), "%s applicable", name);
  c.rebuildNeighborLists(&tg);
  fg.initTraversal();
  c.computeInteractions(&tg);
  fg.endTraversal(n3);
  const auto fRef = forcesOfOwned(ref);
  const auto fGpu = forcesOfOwned(c);
  CHECK(fRef.size() == fGpu.size() && fGpu.size() == s.owned.size(), "%s iterator visits %zu of %zu owned", name,
        fGpu.size(), s.owned.size());
  double maxRel = 0., fmax = 0.;
  for (const auto &[id, f] : fRef) fmax = std::max({fmax, std::fabs(f[0]), std::fabs(f[1]), std::fabs(f[2])});
  for (const auto &[id, f] : fRef) {
    const auto it = fGpu.find(id);
    if (it == fGpu.end()) {
      CHECK(false, "%s id %zu missing", name, id);
      continue;
    }
    for (int d = 0; d < 3; ++d) maxRel = std::max(maxRel, std::fabs(it->second[d] - f[d]) / fmax);
  }
  CHECK(maxRel <= 1e-12, "%s force mismatch %.3e (relative to max |F| = %.3e)", name, maxRel, fmax);
  const double u0 = fr.getPotentialEnergy(), u1 = fg.getPotentialEnergy();
  const double v0 = fr.getVirial(), v1 = fg.getVirial();
  CHECK(std::fabs(u1 - u0) <= 1e-12 * std::fabs(u0), "%s Upot %.17g vs %.17g", name, u1, u0);
  CHECK(std::fabs(v1 - v0) <= 1e-12 * std::fabs(v0), "%s virial %.17g vs %.17g", name, v1, v0);
  std::printf("%-44s newton3=%d  max |dF|/max|F| = %.2e  Upot %.12e  virial %.12e\n", name, int(n3), maxRel, u1, v1);

  // region iterator: same particle set as the reference container
  const std::array<double, 3> lo{2., 3., 1.}, hi{7.5, 9., 12.5};
  std::set<size_t> a, b;
  for (auto it = ref.getRegionIterator(lo, hi, autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) a.insert(it->getID());
  for (auto it = c.getRegionIterator(lo, hi, autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) b.insert(it->getID());
  CHECK(a == b, "%s region iterator: %zu vs %zu particles", name, b.size(), a.size());

  // mutable iterators write through: move owned particles, some of them out of the box; leavers must match
  auto move = [&](auto &cont) {
    for (auto it = cont.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) {
      auto r = it->getR();
      r[0] += (it->getID() % 7 == 0) ? 1.5 : 0.01;
      it->setR(r);
    }
  };
  move(ref);
  move(c);
  auto leaveRef = ref.updateContainer(false);
  auto leaveGpu = c.updateContainer(false);
  std::set<size_t> la, lb;
  for (auto &p : leaveRef) la.insert(p.getID());
  for (auto &p : leaveGpu) lb.insert(p.getID());
  CHECK(la == lb, "%s leavers: %zu vs %zu", name, lb.size(), la.size());
  CHECK(c.getNumberOfParticles(autopas::IteratorBehavior::halo) == 0, "%s halos dropped by updateContainer", name);
  CHECK(c.getNumberOfParticles(autopas::IteratorBehavior::owned) == s.owned.size() - la.size(), "%s owned after update", name);
  // deleteParticle through an iterator
  size_t deleted = 0;
  for (auto it = c.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it)
    if (it->getID() % 5 == 1) {
      autopas::internal::deleteParticle(it);
      ++deleted;
    }
  CHECK(c.getNumberOfParticles(autopas::IteratorBehavior::owned) == s.owned.size() - la.size() - deleted, "%s delete", name);
}

// Axilrod-Teller-Muto through the triwise interfaces: reference LinkedCells + lc_c01 (AoS, newton3 off - the reference's
// own ground-truth configuration for three-body functors, TraversalComparison.cpp:216-220) against gpuLinkedCells +
// gpulc_c08 with GpuATMFunctor, newton3 off and on.
template <bool n3>
static void compareATM(const Scenario &s0) {
  Scenario s = s0;
  s.cutoff = 1.6;  // ~17 neighbours, ~60 triplets per particle: keeps the O(N n^2) reference loop short
  s.skin = 0.2;
  autopas::LinkedCells<Molecule> ref(s.boxMin, s.boxMax, s.cutoff, s.skin, 1.0);
  fill(ref, s);
  using RefFunctor = mdLib::AxilrodTellerMutoFunctor<Molecule, false, autopas::FunctorN3Modes::Both, true, true>;
  RefFunctor fr(s.cutoff);
  fr.setParticleProperties(0.073);
  const auto info = ref.getTraversalSelectorInfo();
  autopas::LCC01Traversal<FMCell, RefFunctor, false> tr(info.cellsPerDim, fr, info.interactionLength, info.cellLength,
                                                        autopas::DataLayoutOption::aos, false);
  ref.rebuildNeighborLists(&tr);
  fr.initTraversal();
  ref.computeInteractions(&tr);
  fr.endTraversal(false);
  autopas_b200::GpuParticleContainer<Molecule> gpu(APB_CONTAINER_LINKED_CELLS, s.boxMin, s.boxMax, s.cutoff, s.skin, 1.0, 4);
  autopas::ParticleContainerInterface<Molecule> &c = gpu;
  fill(c, s);
  using GpuFunctor = autopas_b200::GpuATMFunctor<Molecule, false, autopas::FunctorN3Modes::Both, true, true>;
  GpuFunctor fg(s.cutoff);
  fg.setParticleProperties(0.073);
  autopas_b200::GpuTraversal<GpuFunctor> tg(APB_TRAVERSAL_GPULC_C08, fg, n3);
  CHECK(tg.isApplicableToDomain(), "ATM gpulc_c08 applicable");
  c.rebuildNeighborLists(&tg);
  fg.initTraversal();
  c.computeInteractions(&tg);
  fg.endTraversal(n3);
  const auto fRef = forcesOfOwned(ref);
  const auto fGpu = forcesOfOwned(c);
  double maxRel = 0., fmax = 0.;
  for (const auto &[id, f] : fRef) fmax = std::max({fmax, std::fabs(f[0]), std::fabs(f[1]), std::fabs(f[2])});
  for (const auto &[id, f] : fRef) {
    const auto it = fGpu.find(id);
    if (it == fGpu.end()) {
      CHECK(false, "ATM id %zu missing", id);
      continue;
    }
    for (int d = 0; d < 3; ++d) maxRel = std::max(maxRel, std::fabs(it->second[d] - f[d]) / fmax);
  }
  CHECK(maxRel <= 1e-12, "ATM force mismatch %.3e (relative to max |F| = %.3e)", maxRel, fmax);
  const double u0 = fr.getPotentialEnergy(), u1 = fg.getPotentialEnergy();
  CHECK(std::fabs(u1 - u0) <= 1e-12 * std::fabs(u0), "ATM Upot %.17g vs %.17g", u1, u0);
  // the virial sums f_p * r_p with absolute positions (cancelling terms): yardstick = sum of |f_i| |r_i| over owned
  double vscale = 0.;
  for (auto it = ref.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it)
    for (int d = 0; d < 3; ++d) vscale += std::fabs(it->getF()[d] * it->getR()[d]);
  CHECK(std::fabs(fg.getVirial() - fr.getVirial()) <= 1e-12 * vscale, "ATM virial %.17g vs %.17g", fg.getVirial(), fr.getVirial());
  std::printf("%-44s newton3=%d  max |dF|/max|F| = %.2e  Upot %.12e  virial %.12e\n", "gpuLinkedCells/gpulc_c08 Axilrod-Teller", int(n3),
              maxRel, u1, fg.getVirial());
}

// SPH density and hydro-force functors behind the pairwise interfaces: reference LinkedCells<SPHParticle> + lc_c08 (AoS)
// against gpuLinkedCells + gpulc_c08 / gpulc_c18 with GpuSPHCalcDensityFunctor / GpuSPHCalcHydroForceFunctor. The flow is
// the one of examples/sph/main.cpp: density pass -> pressure set per particle through the container iterators -> hydro
// force pass; density, acceleration, engDot and vSigMax are read back through the iterators. Newton3 evaluates a pair
// once with the first particle's support radius, so that case uses one smoothing length for all particles.
using SPH = sphLib::SPHParticle;
using SPHCell = autopas::FullParticleCell<SPH>;
template <bool n3>
static void compareSPH(const Scenario &s0) {
  const double smth = 1.0;  // support radius 2.5 h = the container cutoff
  std::vector<SPH> owned, halo;
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(0., 1.);
  for (const auto &m : s0.owned)
    owned.emplace_back(m.getR(), m.getV(), m.getID(), 0.8 + 0.4 * u(rng), n3 ? smth : smth * (0.9 + 0.1 * u(rng)), 1.0 + u(rng));
  for (const auto &m : s0.halo) {
    SPH p(m.getR(), std::array<double, 3>{0.1 * u(rng), 0.1 * u(rng), 0.1 * u(rng)}, m.getID(), 0.8 + 0.4 * u(rng),
          n3 ? smth : smth * (0.9 + 0.1 * u(rng)), 1.0 + u(rng));
    p.setOwnershipState(autopas::OwnershipState::halo);
    halo.push_back(p);
  }
  for (auto &p : owned) {
    p.setEnergy(1.0 + u(rng));
    p.setDt(0.01 * u(rng));
  }
  const double cutoff = 2.5, skin = 0.3;
  autopas::LinkedCells<SPH> ref(s0.boxMin, s0.boxMax, cutoff, skin, 1.0);
  autopas_b200::GpuParticleContainer<SPH> gpu(APB_CONTAINER_LINKED_CELLS, s0.boxMin, s0.boxMax, cutoff, skin, 1.0, 4);
  autopas::ParticleContainerInterface<SPH> &c = gpu;
  for (const auto &p : owned) ref.addParticle(p), c.addParticle(p);
  for (const auto &p : halo) ref.addHaloParticle(p), c.addHaloParticle(p);
  const auto info = ref.getTraversalSelectorInfo();
  // ---- density
  sphLib::SPHCalcDensityFunctor<SPH> dr;
  autopas::LCC08Traversal<SPHCell, sphLib::SPHCalcDensityFunctor<SPH>> tdr(info.cellsPerDim, dr, info.interactionLength,
                                                                           info.cellLength, autopas::DataLayoutOption::aos, n3);
  ref.rebuildNeighborLists(&tdr);
  dr.initTraversal();
  ref.computeInteractions(&tdr);
  dr.endTraversal(n3);
  autopas_b200::GpuSPHCalcDensityFunctor<SPH> dg;
  autopas_b200::GpuTraversal<autopas_b200::GpuSPHCalcDensityFunctor<SPH>> tdg(APB_TRAVERSAL_GPULC_C08, dg, n3);
  CHECK(tdg.isApplicableToDomain(), "SPH density gpulc_c08 applicable");
  autopas_b200::GpuTraversal<autopas_b200::GpuSPHCalcDensityFunctor<SPH>> tdv(APB_TRAVERSAL_GPUVCL_PRUNED, dg, n3);
  CHECK(!tdv.isApplicableToDomain(), "SPH functors have no kernel for the cluster-list traversals");
  c.rebuildNeighborLists(&tdg);
  dg.initTraversal();
  c.computeInteractions(&tdg);
  dg.endTraversal(n3);
  std::map<size_t, double> rhoRef;
  for (auto it = ref.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) rhoRef[it->getID()] = it->getDensity();
  double maxRho = 0.;
  size_t seen = 0;
  for (auto it = c.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it, ++seen)
    maxRho = std::max(maxRho, std::fabs(it->getDensity() - rhoRef.at(it->getID())) / rhoRef.at(it->getID()));
  CHECK(seen == owned.size(), "SPH iterator visits %zu of %zu owned", seen, owned.size());
  CHECK(maxRho <= 1e-12, "SPH density mismatch %.3e", maxRho);
  // ---- pressure and halo densities through the mutable iterators (write-through of the SPH columns), host-only
  // attributes (energy, dt) survive the round trip
  auto prepare = [](auto &cont) {
    for (auto it = cont.begin(autopas::IteratorBehavior::ownedOrHalo); it.isValid(); ++it) {
      if (it->isHalo()) it->setDensity(1.0 + 0.001 * static_cast<double>(it->getID() % 97));
      it->setPressure(0.4 * it->getDensity() * (1.0 + 0.01 * static_cast<double>(it->getID() % 13)));
    }
  };
  prepare(ref);
  prepare(c);
  // ---- hydro force
  sphLib::SPHCalcHydroForceFunctor<SPH> hr;
  autopas::LCC08Traversal<SPHCell, sphLib::SPHCalcHydroForceFunctor<SPH>> thr(info.cellsPerDim, hr, info.interactionLength,
                                                                              info.cellLength, autopas::DataLayoutOption::aos, n3);
  hr.initTraversal();
  ref.computeInteractions(&thr);
  hr.endTraversal(n3);
  autopas_b200::GpuSPHCalcHydroForceFunctor<SPH> hg;
  autopas_b200::GpuTraversal<autopas_b200::GpuSPHCalcHydroForceFunctor<SPH>> thg(APB_TRAVERSAL_GPULC_C18, hg, n3);
  hg.initTraversal();
  c.computeInteractions(&thg);
  hg.endTraversal(n3);
  struct Out {
    std::array<double, 3> acc;
    double engDot, vsig, energy, dt;
  };
  std::map<size_t, Out> outRef;
  double amax = 0., emax = 0.;
  for (auto it = ref.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) {
    outRef[it->getID()] = {it->getAcceleration(), it->getEngDot(), it->getVSigMax(), it->getEnergy(), it->getDt()};
    for (int d = 0; d < 3; ++d) amax = std::max(amax, std::fabs(it->getAcceleration()[d]));
    emax = std::max(emax, std::fabs(it->getEngDot()));
  }
  double maxA = 0., maxE = 0., maxV = 0.;
  bool hostOnlyKept = true;
  for (auto it = c.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it) {
    const Out &o = outRef.at(it->getID());
    for (int d = 0; d < 3; ++d) maxA = std::max(maxA, std::fabs(it->getAcceleration()[d] - o.acc[d]) / amax);
    maxE = std::max(maxE, std::fabs(it->getEngDot() - o.engDot) / emax);
    maxV = std::max(maxV, std::fabs(it->getVSigMax() - o.vsig) / std::fabs(o.vsig));
    hostOnlyKept = hostOnlyKept and it->getEnergy() == o.energy and it->getDt() == o.dt;
  }
  CHECK(maxA <= 1e-12, "SPH acceleration mismatch %.3e (relative to max |a| = %.3e)", maxA, amax);
  CHECK(maxE <= 1e-12, "SPH engDot mismatch %.3e (relative to max = %.3e)", maxE, emax);
  CHECK(maxV <= 1e-14, "SPH vSigMax mismatch %.3e", maxV);
  CHECK(hostOnlyKept, "SPH energy / dt (no device column) must survive the device round trip");
  std::printf("%-44s newton3=%d  density %.2e  acceleration %.2e  engDot %.2e  vSigMax %.2e\n",
              "gpuLinkedCells/gpulc_c08+c18 SPH", int(n3), maxRho, maxA, maxE, maxV);
  // leavers are whole particles: SPH attributes and host-only attributes travel with them
  for (auto *cont : {static_cast<autopas::ParticleContainerInterface<SPH> *>(&ref), &c})
    for (auto it = cont->begin(autopas::IteratorBehavior::owned); it.isValid(); ++it)
      if (it->getID() % 11 == 0) it->setR({it->getR()[0], it->getR()[1] - 1.7, it->getR()[2]});
  auto leaveRef = ref.updateContainer(false);
  auto leaveGpu = c.updateContainer(false);
  std::map<size_t, SPH> la;
  for (auto &p : leaveRef) la.emplace(p.getID(), p);
  CHECK(leaveGpu.size() == la.size() and not la.empty(), "SPH leavers: %zu vs %zu", leaveGpu.size(), la.size());
  for (auto &p : leaveGpu) {
    const auto it = la.find(p.getID());
    if (it == la.end()) {
      CHECK(false, "SPH leaver %zu not a reference leaver", static_cast<size_t>(p.getID()));
      continue;
    }
    const SPH &q = it->second;
    const bool same[10] = {p.getR() == q.getR(), p.getV() == q.getV(), p.getMass() == q.getMass(),
                           p.getSmoothingLength() == q.getSmoothingLength(),
                           std::fabs(p.getPressure() - q.getPressure()) <= 1e-12 * q.getPressure(),  // set from the density
                           p.getSoundSpeed() == q.getSoundSpeed(), p.getEnergy() == q.getEnergy(), p.getDt() == q.getDt(),
                           std::fabs(p.getDensity() - q.getDensity()) <= 1e-12 * q.getDensity(),
                           std::fabs(p.getEngDot() - q.getEngDot()) <= 1e-12 * emax};
    bool all = true;
    for (bool b : same) all = all and b;
    CHECK(all, "SPH leaver %zu lost attributes (r v m h P c e dt rho engDot: %d %d %d %d %d %d %d %d %d %d)",
          static_cast<size_t>(p.getID()), same[0], same[1], same[2], same[3], same[4], same[5], same[6], same[7], same[8], same[9]);
  }
}

// The checkpoint piece formatted on the device against the reference writer's own statements
// (examples/md-flexible/src/ParallelVtkWriter.cpp:79-166) run over the same container's iterators: same particle order,
// same std::ostream formatting, same libm in writeWithDynamicPrecision.
static void compareVtk(const Scenario &s) {
  autopas_b200::GpuParticleContainer<Molecule> gpu(APB_CONTAINER_VERLET_CLUSTER_LISTS, s.boxMin, s.boxMax, s.cutoff, s.skin, 4);
  fill(gpu, s);
  size_t k = 0;
  for (auto it = gpu.begin(autopas::IteratorBehavior::owned); it.isValid(); ++it, ++k) {
    it->setF({1e3 * std::sin(double(k)), -1e-7 * k, k % 7 == 0 ? 0. : 1e9 / (k + 1.)});
    if (k % 50 == 0) {  // next to the upper corner: raised precision
      auto r = it->getR();
      r[k % 3] = s.boxMax[k % 3] - std::pow(10., -1. - double(k % 11));
      it->setR(r);
    }
  }
  std::ostringstream want;
  const auto n = gpu.getNumberOfParticles(autopas::IteratorBehavior::owned);
  want << "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n";
  want << "<VTKFile byte_order=\"LittleEndian\" type=\"UnstructuredGrid\" version=\"0.1\">\n";
  want << "  <UnstructuredGrid>\n";
  want << "    <Piece NumberOfCells=\"0\" NumberOfPoints=\"" << n << "\">\n";
  want << "      <PointData>\n";
  want << "        <DataArray Name=\"velocities\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n";
  for (auto p = gpu.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p)
    want << "        " << p->getV()[0] << " " << p->getV()[1] << " " << p->getV()[2] << "\n";
  want << "        </DataArray>\n";
  want << "        <DataArray Name=\"forces\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n";
  for (auto p = gpu.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p)
    want << "        " << p->getF()[0] << " " << p->getF()[1] << " " << p->getF()[2] << "\n";
  want << "        </DataArray>\n";
  want << "        <DataArray Name=\"typeIds\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n";
  for (auto p = gpu.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) want << "        " << p->getTypeId() << "\n";
  want << "        </DataArray>\n";
  want << "        <DataArray Name=\"ids\" NumberOfComponents=\"1\" format=\"ascii\" type=\"Int32\">\n";
  for (auto p = gpu.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) want << "        " << p->getID() << "\n";
  want << "        </DataArray>\n";
  want << "      </PointData>\n      <CellData/>\n      <Points>\n";
  want << "        <DataArray Name=\"positions\" NumberOfComponents=\"3\" format=\"ascii\" type=\"Float32\">\n";
  const auto boxMax = gpu.getBoxMax();
  for (auto p = gpu.begin(autopas::IteratorBehavior::owned); p.isValid(); ++p) {
    const auto dynamic = [&](double position, double border) {
      const auto initial = want.precision();
      if (border - position < 0.1)
        while (autopas::utils::Math::isNearAbs(autopas::utils::Math::roundFloating(position, want.precision()), border,
                                               std::pow(10, -want.precision())))
          want << std::setprecision(want.precision() + 1);
      want << position << std::setprecision(initial);
    };
    want << "        ";
    dynamic(p->getR()[0], boxMax[0]);
    want << " ";
    dynamic(p->getR()[1], boxMax[1]);
    want << " ";
    dynamic(p->getR()[2], boxMax[2]);
    want << "\n";
  }
  want << "        </DataArray>\n      </Points>\n      <Cells>\n";
  want << "        <DataArray Name=\"types\" NumberOfComponents=\"0\" format=\"ascii\" type=\"Float32\"/>\n";
  want << "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n";
  const std::string got = gpu.vtkParticleRecord();
  CHECK(got == want.str(), "device-side VTK record differs from the reference writer's statements (%zu vs %zu bytes)", got.size(),
        want.str().size());
  const std::string path = "/tmp/apb_shim_test_Particles_0_000001.vtu";
  const size_t written = gpu.writeVtkParticleRecord(path);
  std::ifstream back(path, std::ios::binary);
  const std::string fileText((std::istreambuf_iterator<char>(back)), std::istreambuf_iterator<char>());
  CHECK(written == got.size() && fileText == got, "the file written from the device differs from the record (%zu vs %zu bytes)",
        fileText.size(), got.size());
  std::remove(path.c_str());
  std::printf("vtk record: %zu particles, %zu bytes, %s\n", size_t(n), got.size(), got == want.str() ? "identical" : "DIFFERENT");
}

int main() {
  autopas::utils::ExceptionHandler::setBehavior(autopas::utils::ExceptionBehavior::throwException);
  const Scenario s = makeScenario(42);
  try {
    compare<true>(s, APB_CONTAINER_LINKED_CELLS, APB_TRAVERSAL_GPULC_C08, 4, "gpuLinkedCells/gpulc_c08");
    compare<false>(s, APB_CONTAINER_LINKED_CELLS, APB_TRAVERSAL_GPULC_C18, 4, "gpuLinkedCells/gpulc_c18");
    compare<true>(s, APB_CONTAINER_VERLET_CLUSTER_LISTS, APB_TRAVERSAL_GPUVCL_C06, 4, "gpuVerletClusterLists/gpuvcl_c06");
    compare<false>(s, APB_CONTAINER_VERLET_CLUSTER_LISTS, APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION, 4,
                   "gpuVerletClusterLists/gpuvcl_cluster_iteration");
    compare<false>(s, APB_CONTAINER_VERLET_CLUSTER_LISTS, APB_TRAVERSAL_GPUVCL_PRUNED, 32, "gpuVerletClusterLists/gpuvcl_pruned");
    compareATM<false>(s);
    compareATM<true>(s);
    compareSPH<false>(s);
    compareSPH<true>(s);
    compareVtk(s);
    // wrong traversal type is rejected like the reference containers do
    autopas_b200::GpuParticleContainer<Molecule> gpu(APB_CONTAINER_LINKED_CELLS, s.boxMin, s.boxMax, s.cutoff, s.skin);
    using RefFunctor = mdLib::LJFunctor<Molecule>;
    RefFunctor fr(s.cutoff);
    const auto info = gpu.getTraversalSelectorInfo();
    autopas::LCC08Traversal<FMCell, RefFunctor> tr(info.cellsPerDim, fr, info.interactionLength, info.cellLength,
                                                   autopas::DataLayoutOption::aos, true);
    bool threw = false;
    try {
      gpu.computeInteractions(&tr);
    } catch (const autopas::utils::ExceptionHandler::AutoPasException &) {
      threw = true;
    }
    CHECK(threw, "CPU traversal on the GPU container must throw");
    // a functor without a GPU kernel makes the GPU traversal inapplicable (no CPU fallback)
    autopas_b200::GpuTraversal<RefFunctor> tn(APB_TRAVERSAL_GPULC_C08, fr, true);
    CHECK(!tn.isApplicableToDomain(), "reference LJFunctor has no GPU kernel: configuration must be inapplicable");
    // adding an owned particle outside the box throws (ParticleContainerInterface.h:92-105)
    threw = false;
    try {
      gpu.addParticle(Molecule({-1., 0., 0.}, {0., 0., 0.}, 1, 0));
    } catch (const autopas::utils::ExceptionHandler::AutoPasException &) {
      threw = true;
    }
    CHECK(threw, "addParticle outside the box must throw");
  } catch (const std::exception &e) {
    std::printf("FAIL: exception %s\n", e.what());
    ++g_fail;
  }
  std::printf(g_fail ? "SHIM TEST FAILED (%d)\n" : "SHIM TEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
