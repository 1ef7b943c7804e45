/**
 * @file GpuContainers.h
 * Header-only C++20 drop-in classes that put the B200 library (include/autopas_b200.h) behind AutoPas' own container /
 * traversal / functor interfaces. Compiles against the UNMODIFIED AutoPas headers; INTEGRATION.md lists the additive
 * enum / selector edits that make the tuner generate these classes.
 *
 *   autopas_b200::GpuParticleContainer<Particle_T>  : autopas::ParticleContainerInterface<Particle_T>
 *       (src/autopas/containers/ParticleContainerInterface.h:38-410). Particles live in device SoA columns; the host
 *       sees them through a lazily synchronised AoS mirror that is downloaded when an iterator touches it and written
 *       back before the next device operation if it was handed out mutable.
 *   autopas_b200::GpuTraversal<Functor_T>           : autopas::TraversalInterface (containers/TraversalInterface.h:18-84)
 *       maps the static functor type to a kernel descriptor; functors without a GPU kernel make
 *       isApplicableToDomain() false, so the configuration is rejected instead of silently running on the CPU.
 *   autopas_b200::GpuLJFunctor<Particle_T, ...>     : autopas::PairwiseFunctor, wraps mdLib::LJFunctor (same template
 *       flags). CPU traversals run the wrapped reference functor unchanged; GPU traversals deposit the raw accumulators
 *       (LJFunctor.h:1121-1195) here and endTraversal applies the reference normalisation (LJFunctor.h:661-685).
 */
#pragma once

#include <array>
#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "autopas/baseFunctors/PairwiseFunctor.h"
#include "autopas/containers/ParticleContainerInterface.h"
#include "autopas/containers/TraversalInterface.h"
#include "autopas/iterators/ContainerIterator.h"
#include "autopas/options/DataLayoutOption.h"
#include "autopas/particles/OwnershipState.h"
#include "autopas/utils/ExceptionHandler.h"
#include "autopas/utils/WrapOpenMP.h"
#include "autopas/utils/inBox.h"
#include "autopas/utils/markParticleAsDeleted.h"
#include "autopas_b200.h"
#include "autopas/baseFunctors/TriwiseFunctor.h"
#include "molecularDynamicsLibrary/AxilrodTellerMutoFunctor.h"
#include "molecularDynamicsLibrary/LJFunctor.h"
#include "molecularDynamicsLibrary/ParticlePropertiesLibrary.h"
#if defined(MD_FLEXIBLE_MODE) && defined(MULTISITE) && MD_FLEXIBLE_MODE == MULTISITE
#define AUTOPAS_B200_MULTISITE 1
#include "molecularDynamicsLibrary/LJMultisiteFunctor.h"
#endif
#if __has_include("SPHLibrary/SPHCalcDensityFunctor.h")
#define AUTOPAS_B200_SPH 1
#include "SPHLibrary/SPHCalcDensityFunctor.h"
#include "SPHLibrary/SPHCalcHydroForceFunctor.h"
#endif

namespace autopas_b200 {

/// Which device particle kind (apb_particle_kind) and which SoA columns a particle class maps to: mdLib::MoleculeLJ
/// (MoleculeLJ.h:39-70), mdLib::MultisiteMoleculeLJ (quaternion, torque) and sphLib::SPHParticle (SPHParticle.h: mass,
/// smoothing length, density, pressure, sound speed, engDot, vSigMax; the force columns carry the acceleration).
/// Attributes without a device column (angular velocity; SPH energy and dt) stay on the host, keyed by particle id.
template <class P>
struct ParticleColumns {
  static constexpr bool sph = requires(const P &p) { p.getSmoothingLength(); };
  static constexpr bool multisite = requires(const P &p) { p.getQuaternion(); };
  static constexpr bool hasOldF = requires(const P &p) { p.getOldF(); };
  static constexpr int kind = sph ? APB_PARTICLE_SPH : (multisite ? APB_PARTICLE_MULTISITE : APB_PARTICLE_LJ);
  static constexpr int numHostOnly = sph ? 2 : (multisite ? 3 : 0);
  static constexpr int maxColumns = 24;

  /// device columns of this particle class; the first three are the position
  static const std::vector<int> &ids() {
    static const std::vector<int> v = [] {
      std::vector<int> c{APB_COL_X, APB_COL_Y, APB_COL_Z, APB_COL_VX, APB_COL_VY, APB_COL_VZ, APB_COL_FX, APB_COL_FY, APB_COL_FZ};
      if constexpr (hasOldF and not sph) c.insert(c.end(), {APB_COL_OLDFX, APB_COL_OLDFY, APB_COL_OLDFZ});
      if constexpr (multisite) c.insert(c.end(), {APB_COL_Q0, APB_COL_Q1, APB_COL_Q2, APB_COL_Q3, APB_COL_TX, APB_COL_TY, APB_COL_TZ});
      if constexpr (sph)
        c.insert(c.end(), {APB_COL_MASS, APB_COL_SMTH, APB_COL_DENSITY, APB_COL_PRESSURE, APB_COL_SNDSPEED, APB_COL_ENGDOT, APB_COL_VSIGMAX});
      return c;
    }();
    return v;
  }
  /// out[k] = value of column ids()[k]
  static void gather(const P &p, double *out) {
    int k = 0;
    for (int d = 0; d < 3; ++d) out[k++] = p.getR()[d];
    for (int d = 0; d < 3; ++d) out[k++] = p.getV()[d];
    if constexpr (sph) {
      for (int d = 0; d < 3; ++d) out[k++] = p.getAcceleration()[d];
    } else {
      for (int d = 0; d < 3; ++d) out[k++] = p.getF()[d];
    }
    if constexpr (hasOldF and not sph)
      for (int d = 0; d < 3; ++d) out[k++] = p.getOldF()[d];
    if constexpr (multisite) {
      for (int d = 0; d < 4; ++d) out[k++] = p.getQuaternion()[d];
      for (int d = 0; d < 3; ++d) out[k++] = p.getTorque()[d];
    }
    if constexpr (sph) {
      out[k++] = p.getMass();
      out[k++] = p.getSmoothingLength();
      out[k++] = p.getDensity();
      out[k++] = p.getPressure();
      out[k++] = p.getSoundSpeed();
      out[k++] = p.getEngDot();
      out[k++] = p.getVSigMax();
    }
  }
  static void scatter(P &p, const double *in) {
    int k = 0;
    p.setR({in[0], in[1], in[2]});
    p.setV({in[3], in[4], in[5]});
    if constexpr (sph) p.setAcceleration({in[6], in[7], in[8]});
    else p.setF({in[6], in[7], in[8]});
    k = 9;
    if constexpr (hasOldF and not sph) {
      p.setOldF({in[k], in[k + 1], in[k + 2]});
      k += 3;
    }
    if constexpr (multisite) {
      p.setQuaternion({in[k], in[k + 1], in[k + 2], in[k + 3]});
      p.setTorque({in[k + 4], in[k + 5], in[k + 6]});
      k += 7;
    }
    if constexpr (sph) {
      p.setMass(in[k]);
      p.setSmoothingLength(in[k + 1]);
      p.setDensity(in[k + 2]);
      p.setPressure(in[k + 3]);
      p.setSoundSpeed(in[k + 4]);
      p.setEngDot(in[k + 5]);
      p.setVSigMax(in[k + 6]);
    }
  }
  static std::array<double, 3> hostOnly(const P &p) {
    if constexpr (sph) return {p.getEnergy(), p.getDt(), 0.};
    else if constexpr (multisite) return p.getAngularVel();
    else return {0., 0., 0.};
  }
  static void setHostOnly(P &p, const std::array<double, 3> &v) {
    if constexpr (sph) {
      p.setEnergy(v[0]);
      p.setDt(v[1]);
    } else if constexpr (multisite) {
      p.setAngularVel(v);
    }
  }
};

/// Kernel descriptor handed from a GPU-capable functor to the container.
struct FunctorDescriptor {
  apb_functor functor{};
  std::vector<double> mixingTable;  // keeps functor.mixing_table alive
  std::vector<int32_t> siteStart, siteTypes;  // multi-site: keep functor.site_start / site_types alive
  std::vector<double> sitePositions;          // and functor.site_positions
  FunctorDescriptor() = default;
  FunctorDescriptor(FunctorDescriptor &&o) noexcept { *this = std::move(o); }
  FunctorDescriptor &operator=(FunctorDescriptor &&o) noexcept {
    functor = o.functor;
    mixingTable = std::move(o.mixingTable);
    siteStart = std::move(o.siteStart);
    siteTypes = std::move(o.siteTypes);
    sitePositions = std::move(o.sitePositions);
    if (functor.mixing_table) functor.mixing_table = mixingTable.data();
    if (functor.site_start) functor.site_start = siteStart.data();
    if (functor.site_types) functor.site_types = siteTypes.data();
    if (functor.site_positions) functor.site_positions = sitePositions.data();
    return *this;
  }
};

/// Interface the container sees; implemented by GpuTraversal<Functor_T>.
class GpuTraversalInterface {
 public:
  virtual ~GpuTraversalInterface() = default;
  [[nodiscard]] virtual int apbTraversal() const = 0;
  [[nodiscard]] virtual bool functorHasGpuKernel() const = 0;
  virtual FunctorDescriptor describeFunctor() = 0;
  virtual void depositResult(const apb_traversal_result &raw) = 0;
};

/// Names as they would appear in TraversalOption (unique prefixes gpulc_ / gpuvcl_, CompatibleTraversals.h:29-37).
inline const char *traversalName(int t) {
  switch (t) {
    case APB_TRAVERSAL_GPULC_C08: return "gpulc_c08";
    case APB_TRAVERSAL_GPULC_C18: return "gpulc_c18";
    case APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION: return "gpuvcl_cluster_iteration";
    case APB_TRAVERSAL_GPUVCL_C06: return "gpuvcl_c06";
    case APB_TRAVERSAL_GPUVCL_C01_BALANCED: return "gpuvcl_c01_balanced";
    default: return "gpuvcl_pruned";
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// GpuLJFunctor
// ---------------------------------------------------------------------------------------------------------------------
template <class Particle_T, bool applyShift = false, bool useMixing = false,
          autopas::FunctorN3Modes useNewton3 = autopas::FunctorN3Modes::Both, bool calculateGlobals = false,
          bool countFLOPs = false, bool relevantForTuning = true>
class GpuLJFunctor
    : public autopas::PairwiseFunctor<Particle_T, GpuLJFunctor<Particle_T, applyShift, useMixing, useNewton3,
                                                               calculateGlobals, countFLOPs, relevantForTuning>> {
  using Self = GpuLJFunctor<Particle_T, applyShift, useMixing, useNewton3, calculateGlobals, countFLOPs, relevantForTuning>;
  using Cpu = mdLib::LJFunctor<Particle_T, applyShift, useMixing, useNewton3, calculateGlobals, countFLOPs, relevantForTuning>;
  using SoAArraysType = typename Particle_T::SoAArraysType;

 public:
  static constexpr bool apbHasGpuKernel = true;

  explicit GpuLJFunctor(double cutoff) requires(not useMixing)
      : autopas::PairwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff), _cutoff(cutoff) {}
  GpuLJFunctor(double cutoff, ParticlePropertiesLibrary<double, size_t> &ppl) requires(useMixing)
      : autopas::PairwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff, ppl), _cutoff(cutoff), _ppl(&ppl) {}

  std::string getName() final { return "GpuLJFunctor"; }
  bool isRelevantForTuning() final { return relevantForTuning; }
  bool allowsNewton3() final { return _cpu.allowsNewton3(); }
  bool allowsNonNewton3() final { return _cpu.allowsNonNewton3(); }

  // ---- CPU path: the wrapped reference functor, untouched ----
  void AoSFunctor(Particle_T &i, Particle_T &j, bool newton3) final { _cpu.AoSFunctor(i, j, newton3); }
  void SoAFunctorSingle(autopas::SoAView<SoAArraysType> soa, bool newton3) final { _cpu.SoAFunctorSingle(soa, newton3); }
  void SoAFunctorPair(autopas::SoAView<SoAArraysType> soa1, autopas::SoAView<SoAArraysType> soa2, bool newton3) final {
    _cpu.SoAFunctorPair(soa1, soa2, newton3);
  }
  void SoAFunctorVerlet(autopas::SoAView<SoAArraysType> soa, const size_t indexFirst,
                        const std::vector<size_t, autopas::AlignedAllocator<size_t>> &neighborList, bool newton3) final {
    _cpu.SoAFunctorVerlet(soa, indexFirst, neighborList, newton3);
  }
  constexpr static auto getNeededAttr() { return Cpu::getNeededAttr(); }
  constexpr static auto getNeededAttr(std::false_type) { return Cpu::getNeededAttr(std::false_type()); }
  constexpr static auto getComputedAttr() { return Cpu::getComputedAttr(); }
  constexpr static bool getMixing() { return useMixing; }

  /// LJFunctor::setParticleProperties (LJFunctor.h:588-596)
  void setParticleProperties(double epsilon24, double sigmaSquared) {
    _cpu.setParticleProperties(epsilon24, sigmaSquared);
    _epsilon24 = epsilon24;
    _sigmaSquared = sigmaSquared;
  }

  void initTraversal() final {
    _cpu.initTraversal();
    _gpuRaw = {};
    _gpuUpot = _gpuVirial = 0.;
    _postProcessed = false;
  }

  /// Reference normalisation (LJFunctor.h:661-685): Upot = sum * 0.5 / 6, virial = (vx + vy + vz) * 0.5
  void endTraversal(bool newton3) final {
    if (_postProcessed) {
      autopas::utils::ExceptionHandler::exception(
          "Already postprocessed, endTraversal(bool newton3) was called twice without calling initTraversal().");
    }
    _cpu.endTraversal(newton3);
    if constexpr (calculateGlobals) {
      apb_lj_end_traversal(&_gpuRaw, &_gpuUpot, &_gpuVirial);
    }
    _postProcessed = true;
  }

  double getPotentialEnergy() {
    double cpu = 0.;
    if constexpr (calculateGlobals) cpu = _cpu.getPotentialEnergy();  // throws like the reference if misused
    else return _cpu.getPotentialEnergy();
    return cpu + _gpuUpot;
  }
  double getVirial() {
    double cpu = 0.;
    if constexpr (calculateGlobals) cpu = _cpu.getVirial();
    else return _cpu.getVirial();
    return cpu + _gpuVirial;
  }
  [[nodiscard]] size_t getNumFLOPs() const override {
    if constexpr (countFLOPs) return _cpu.getNumFLOPs() + apb_lj_num_flops(&_gpuRaw, applyShift ? 1 : 0);
    return std::numeric_limits<size_t>::max();
  }
  [[nodiscard]] double getHitRate() const override {
    if constexpr (countFLOPs) {
      const auto kernel = _gpuRaw.num_kernel_calls_n3 + _gpuRaw.num_kernel_calls_no_n3;
      if (_gpuRaw.num_dist_calls > 0) return static_cast<double>(kernel) / static_cast<double>(_gpuRaw.num_dist_calls);
      return _cpu.getHitRate();
    }
    return std::numeric_limits<double>::quiet_NaN();
  }

  // ---- GPU path hooks used by GpuTraversal ----
  FunctorDescriptor apbDescribe() {
    FunctorDescriptor d;
    d.functor.kind = APB_FUNCTOR_LJ;
    d.functor.flags = (applyShift ? APB_FUNCTOR_APPLY_SHIFT : 0) | (useMixing ? APB_FUNCTOR_USE_MIXING : 0) |
                      (calculateGlobals ? APB_FUNCTOR_CALC_GLOBALS : 0) | (countFLOPs ? APB_FUNCTOR_COUNT_FLOPS : 0) |
                      APB_FUNCTOR_VIRIAL_TRACE;  // getVirial() returns the sum of the components (LJFunctor.h:719)
    d.functor.cutoff = _cutoff;
    d.functor.epsilon24 = _epsilon24;
    d.functor.sigma_squared = _sigmaSquared;
    if constexpr (useMixing) {
      // ParticlePropertiesLibrary stores {epsilon24, sigmaSquared, shift6} per pair (ParticlePropertiesLibrary.h:324-328)
      const auto T = _ppl->getNumberRegisteredSiteTypes();
      d.mixingTable.resize(T * T * 3);
      for (size_t i = 0; i < T; ++i)
        for (size_t j = 0; j < T; ++j) {
          d.mixingTable[3 * (i * T + j) + 0] = _ppl->getMixing24Epsilon(i, j);
          d.mixingTable[3 * (i * T + j) + 1] = _ppl->getMixingSigmaSquared(i, j);
          d.mixingTable[3 * (i * T + j) + 2] = _ppl->getMixingShift6(i, j);
        }
      d.functor.num_types = static_cast<int32_t>(T);
      d.functor.mixing_table = d.mixingTable.data();
    }
    return d;
  }
  void apbDeposit(const apb_traversal_result &raw) {
    _gpuRaw.upot_sum += raw.upot_sum;
    for (int d = 0; d < 3; ++d) _gpuRaw.virial_sum[d] += raw.virial_sum[d];
    _gpuRaw.num_dist_calls += raw.num_dist_calls;
    _gpuRaw.num_kernel_calls_n3 += raw.num_kernel_calls_n3;
    _gpuRaw.num_kernel_calls_no_n3 += raw.num_kernel_calls_no_n3;
    _gpuRaw.num_global_calcs_n3 += raw.num_global_calcs_n3;
    _gpuRaw.num_global_calcs_no_n3 += raw.num_global_calcs_no_n3;
  }

 private:
  Cpu _cpu;
  double _cutoff;
  ParticlePropertiesLibrary<double, size_t> *_ppl = nullptr;
  double _epsilon24 = 0., _sigmaSquared = 0.;
  apb_traversal_result _gpuRaw{};
  double _gpuUpot = 0., _gpuVirial = 0.;
  bool _postProcessed = false;
};

// ---------------------------------------------------------------------------------------------------------------------
// GpuATMFunctor
// ---------------------------------------------------------------------------------------------------------------------
/// autopas::TriwiseFunctor around mdLib::AxilrodTellerMutoFunctor (same template flags): CPU triwise traversals call the
/// wrapped reference functor; the GPU traversal gpulc_c08 (kATMTriplets / kATMTripletsN3) deposits the raw accumulators
/// and endTraversal applies the reference normalisation (AxilrodTellerMutoFunctor.h:360-386: Upot = sum / 9, virial
/// not scaled).
template <class Particle_T, bool useMixing = false, autopas::FunctorN3Modes useNewton3 = autopas::FunctorN3Modes::Both,
          bool calculateGlobals = false, bool countFLOPs = false>
class GpuATMFunctor
    : public autopas::TriwiseFunctor<Particle_T, GpuATMFunctor<Particle_T, useMixing, useNewton3, calculateGlobals, countFLOPs>> {
  using Self = GpuATMFunctor<Particle_T, useMixing, useNewton3, calculateGlobals, countFLOPs>;
  using Cpu = mdLib::AxilrodTellerMutoFunctor<Particle_T, useMixing, useNewton3, calculateGlobals, countFLOPs>;

 public:
  static constexpr bool apbHasGpuKernel = true;
  static constexpr bool apbLinkedCellsOnly = true;

  explicit GpuATMFunctor(double cutoff) requires(not useMixing)
      : autopas::TriwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff), _cutoff(cutoff) {}
  GpuATMFunctor(double cutoff, ParticlePropertiesLibrary<double, size_t> &ppl) requires(useMixing)
      : autopas::TriwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff, ppl), _cutoff(cutoff), _ppl(&ppl) {}

  std::string getName() final { return "GpuAxilrodTellerMutoFunctor"; }
  bool isRelevantForTuning() final { return true; }
  bool allowsNewton3() final { return _cpu.allowsNewton3(); }
  bool allowsNonNewton3() final { return _cpu.allowsNonNewton3(); }
  void AoSFunctor(Particle_T &i, Particle_T &j, Particle_T &k, bool newton3) final { _cpu.AoSFunctor(i, j, k, newton3); }
  constexpr static auto getNeededAttr() { return Cpu::getNeededAttr(); }
  constexpr static auto getNeededAttr(std::false_type) { return Cpu::getNeededAttr(std::false_type()); }
  constexpr static auto getComputedAttr() { return Cpu::getComputedAttr(); }
  constexpr static bool getMixing() { return useMixing; }
  void setParticleProperties(double nu) {
    _cpu.setParticleProperties(nu);
    _nu = nu;
  }
  void initTraversal() final {
    _cpu.initTraversal();
    _gpuRaw = {};
    _gpuUpot = _gpuVirial = 0.;
    _postProcessed = false;
  }
  void endTraversal(bool newton3) final {
    if (_postProcessed) {
      autopas::utils::ExceptionHandler::exception(
          "Already postprocessed, endTraversal(bool newton3) was called twice without calling initTraversal().");
    }
    _cpu.endTraversal(newton3);
    if constexpr (calculateGlobals) apb_atm_end_traversal(&_gpuRaw, &_gpuUpot, &_gpuVirial);
    _postProcessed = true;
  }
  double getPotentialEnergy() { return _cpu.getPotentialEnergy() + _gpuUpot; }
  double getVirial() { return _cpu.getVirial() + _gpuVirial; }
  [[nodiscard]] size_t getNumFLOPs() const override {
    if constexpr (countFLOPs) return _cpu.getNumFLOPs() + apb_atm_num_flops(&_gpuRaw);
    return std::numeric_limits<size_t>::max();
  }
  [[nodiscard]] double getHitRate() const override {
    if constexpr (countFLOPs) {
      if (_gpuRaw.num_dist_calls > 0)
        return static_cast<double>(_gpuRaw.num_kernel_calls_n3 + _gpuRaw.num_kernel_calls_no_n3) /
               static_cast<double>(_gpuRaw.num_dist_calls);
      return _cpu.getHitRate();
    }
    return std::numeric_limits<double>::quiet_NaN();
  }

  FunctorDescriptor apbDescribe() {
    FunctorDescriptor d;
    d.functor.kind = APB_FUNCTOR_ATM;
    d.functor.flags = (useMixing ? APB_FUNCTOR_USE_MIXING : 0) | (calculateGlobals ? APB_FUNCTOR_CALC_GLOBALS : 0) |
                      (countFLOPs ? APB_FUNCTOR_COUNT_FLOPS : 0);
    d.functor.cutoff = _cutoff;
    d.functor.nu = _nu;
    if constexpr (useMixing) {
      // nu_ijk = cbrt(nu_i nu_j nu_k), row-major [T * T * T] (ParticlePropertiesLibrary.h:460-470)
      const auto T = _ppl->getNumberRegisteredSiteTypes();
      d.mixingTable.resize(T * T * T);
      for (size_t i = 0; i < T; ++i)
        for (size_t j = 0; j < T; ++j)
          for (size_t k = 0; k < T; ++k) d.mixingTable[(i * T + j) * T + k] = _ppl->getMixingNu(i, j, k);
      d.functor.num_types = static_cast<int32_t>(T);
      d.functor.mixing_table = d.mixingTable.data();
    }
    return d;
  }
  void apbDeposit(const apb_traversal_result &raw) {
    _gpuRaw.upot_sum += raw.upot_sum;
    for (int d = 0; d < 3; ++d) _gpuRaw.virial_sum[d] += raw.virial_sum[d];
    _gpuRaw.num_dist_calls += raw.num_dist_calls;
    _gpuRaw.num_kernel_calls_n3 += raw.num_kernel_calls_n3;
    _gpuRaw.num_kernel_calls_no_n3 += raw.num_kernel_calls_no_n3;
    _gpuRaw.num_global_calcs_n3 += raw.num_global_calcs_n3;
    _gpuRaw.num_global_calcs_no_n3 += raw.num_global_calcs_no_n3;
  }

 private:
  Cpu _cpu;
  double _cutoff;
  ParticlePropertiesLibrary<double, size_t> *_ppl = nullptr;
  double _nu = 0.;
  apb_traversal_result _gpuRaw{};
  double _gpuUpot = 0., _gpuVirial = 0.;
  bool _postProcessed = false;
};

#ifdef AUTOPAS_B200_SPH
// ---------------------------------------------------------------------------------------------------------------------
// GpuSPHCalcDensityFunctor / GpuSPHCalcHydroForceFunctor
// ---------------------------------------------------------------------------------------------------------------------
/// Common part of the two SPH wrappers: the CPU interfaces forward to the wrapped reference functor `Cpu_T`
/// (SPHCalcDensityFunctor.h:20 / SPHCalcHydroForceFunctor.h:19; no template flags, no globals, functor cutoff 0 - the
/// kernel support 2.5 h is applied per pair); the GPU traversals gpulc_c08 / gpulc_c18 run kSPHDensityLC / kSPHHydroLC
/// on gpuLinkedCells with SPHParticle storage and write density, or acceleration / engDot / vSigMax, on the device.
template <class Particle_T, class Cpu_T, class Self_T, int apbKind>
class GpuSPHFunctorBase : public autopas::PairwiseFunctor<Particle_T, Self_T> {
  using SoAArraysType = typename Particle_T::SoAArraysType;

 public:
  static constexpr bool apbHasGpuKernel = true;
  static constexpr bool apbLinkedCellsOnly = true;
  GpuSPHFunctorBase() : autopas::PairwiseFunctor<Particle_T, Self_T>(0.) {}
  bool isRelevantForTuning() final { return true; }
  bool allowsNewton3() final { return true; }
  bool allowsNonNewton3() final { return true; }
  void AoSFunctor(Particle_T &i, Particle_T &j, bool newton3) final { _cpu.AoSFunctor(i, j, newton3); }
  void SoAFunctorSingle(autopas::SoAView<SoAArraysType> soa, bool newton3) final { _cpu.SoAFunctorSingle(soa, newton3); }
  void SoAFunctorPair(autopas::SoAView<SoAArraysType> soa1, autopas::SoAView<SoAArraysType> soa2, bool newton3) final {
    _cpu.SoAFunctorPair(soa1, soa2, newton3);
  }
  void SoAFunctorVerlet(autopas::SoAView<SoAArraysType> soa, const size_t indexFirst,
                        const std::vector<size_t, autopas::AlignedAllocator<size_t>> &neighborList, bool newton3) final {
    _cpu.SoAFunctorVerlet(soa, indexFirst, neighborList, newton3);
  }
  constexpr static auto getNeededAttr() { return Cpu_T::getNeededAttr(); }
  constexpr static auto getNeededAttr(std::false_type) { return Cpu_T::getNeededAttr(std::false_type()); }
  constexpr static auto getComputedAttr() { return Cpu_T::getComputedAttr(); }

  FunctorDescriptor apbDescribe() {
    FunctorDescriptor d;
    d.functor.kind = apbKind;
    return d;
  }
  void apbDeposit(const apb_traversal_result &) {}  // no globals (SPHCalcDensityFunctor.h / SPHCalcHydroForceFunctor.h)

 protected:
  Cpu_T _cpu;
};

template <class Particle_T>
class GpuSPHCalcDensityFunctor
    : public GpuSPHFunctorBase<Particle_T, sphLib::SPHCalcDensityFunctor<Particle_T>, GpuSPHCalcDensityFunctor<Particle_T>,
                               APB_FUNCTOR_SPH_DENSITY> {
 public:
  std::string getName() final { return "GpuSPHDensityFunctor"; }
  /// SPHCalcDensityFunctor.h:65-72
  static unsigned long getNumFlopsPerKernelCall() { return sphLib::SPHCalcDensityFunctor<Particle_T>::getNumFlopsPerKernelCall(); }
};

template <class Particle_T>
class GpuSPHCalcHydroForceFunctor
    : public GpuSPHFunctorBase<Particle_T, sphLib::SPHCalcHydroForceFunctor<Particle_T>,
                               GpuSPHCalcHydroForceFunctor<Particle_T>, APB_FUNCTOR_SPH_HYDRO> {
 public:
  std::string getName() final { return "GpuSPHHydroForceFunctor"; }
};
#endif  // AUTOPAS_B200_SPH

#ifdef AUTOPAS_B200_MULTISITE
// ---------------------------------------------------------------------------------------------------------------------
// GpuLJMultisiteFunctor
// ---------------------------------------------------------------------------------------------------------------------
/// autopas::PairwiseFunctor around mdLib::LJMultisiteFunctor (LJMultisiteFunctor.h:44-47, same template flags; the
/// reference library offers it only in its MULTISITE build mode, so does this wrapper). GPU traversal: gpulc_c08 /
/// gpulc_c18 on gpuLinkedCells with MultisiteMoleculeLJ storage (kLJMultisiteLC: forces and torques on the device).
/// Site geometry and site types come from the ParticlePropertiesLibrary (getSitePositions / getSiteTypes, :186-196)
/// when mixing is used, else from setParticleProperties (:629-640).
template <class Particle_T, bool applyShift = false, bool useMixing = false,
          autopas::FunctorN3Modes useNewton3 = autopas::FunctorN3Modes::Both, bool calculateGlobals = false,
          bool relevantForTuning = true>
class GpuLJMultisiteFunctor
    : public autopas::PairwiseFunctor<Particle_T, GpuLJMultisiteFunctor<Particle_T, applyShift, useMixing, useNewton3,
                                                                        calculateGlobals, relevantForTuning>> {
  using Self = GpuLJMultisiteFunctor<Particle_T, applyShift, useMixing, useNewton3, calculateGlobals, relevantForTuning>;
  using Cpu = mdLib::LJMultisiteFunctor<Particle_T, applyShift, useMixing, useNewton3, calculateGlobals, relevantForTuning>;
  using SoAArraysType = typename Particle_T::SoAArraysType;

 public:
  static constexpr bool apbHasGpuKernel = true;
  static constexpr bool apbLinkedCellsOnly = true;

  explicit GpuLJMultisiteFunctor(double cutoff) requires(not useMixing)
      : autopas::PairwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff), _cutoff(cutoff) {}
  GpuLJMultisiteFunctor(double cutoff, ParticlePropertiesLibrary<double, size_t> &ppl) requires(useMixing)
      : autopas::PairwiseFunctor<Particle_T, Self>(cutoff), _cpu(cutoff, ppl), _cutoff(cutoff), _ppl(&ppl) {}

  std::string getName() final { return "GpuLJMultisiteFunctor"; }
  bool isRelevantForTuning() final { return relevantForTuning; }
  bool allowsNewton3() final { return _cpu.allowsNewton3(); }
  bool allowsNonNewton3() final { return _cpu.allowsNonNewton3(); }
  void AoSFunctor(Particle_T &i, Particle_T &j, bool newton3) final { _cpu.AoSFunctor(i, j, newton3); }
  void SoAFunctorSingle(autopas::SoAView<SoAArraysType> soa, bool newton3) final { _cpu.SoAFunctorSingle(soa, newton3); }
  void SoAFunctorPair(autopas::SoAView<SoAArraysType> soa1, autopas::SoAView<SoAArraysType> soa2, bool newton3) final {
    _cpu.SoAFunctorPair(soa1, soa2, newton3);
  }
  void SoAFunctorVerlet(autopas::SoAView<SoAArraysType> soa, const size_t indexFirst,
                        const std::vector<size_t, autopas::AlignedAllocator<size_t>> &neighborList, bool newton3) final {
    _cpu.SoAFunctorVerlet(soa, indexFirst, neighborList, newton3);
  }
  constexpr static auto getNeededAttr() { return Cpu::getNeededAttr(); }
  constexpr static auto getNeededAttr(std::false_type) { return Cpu::getNeededAttr(std::false_type()); }
  constexpr static auto getComputedAttr() { return Cpu::getComputedAttr(); }
  constexpr static bool getMixing() { return useMixing; }

  void setParticleProperties(double epsilon24, double sigmaSquared, std::vector<std::array<double, 3>> sitePositionsLJ) {
    _cpu.setParticleProperties(epsilon24, sigmaSquared, sitePositionsLJ);
    _epsilon24 = epsilon24;
    _sigmaSquared = sigmaSquared;
    _sitePositions = std::move(sitePositionsLJ);
  }
  void initTraversal() final {
    _cpu.initTraversal();
    _gpuRaw = {};
    _gpuUpot = _gpuVirial = 0.;
    _postProcessed = false;
  }
  /// same normalisation as LJFunctor (LJMultisiteFunctor.h:725-750)
  void endTraversal(bool newton3) final {
    if (_postProcessed) {
      autopas::utils::ExceptionHandler::exception(
          "Already postprocessed, endTraversal(bool newton3) was called twice without calling initTraversal().");
    }
    _cpu.endTraversal(newton3);
    if constexpr (calculateGlobals) apb_lj_end_traversal(&_gpuRaw, &_gpuUpot, &_gpuVirial);
    _postProcessed = true;
  }
  double getPotentialEnergy() { return _cpu.getPotentialEnergy() + _gpuUpot; }  // throws like the reference if misused
  double getVirial() { return _cpu.getVirial() + _gpuVirial; }

  FunctorDescriptor apbDescribe() {
    FunctorDescriptor d;
    d.functor.kind = APB_FUNCTOR_LJ_MULTISITE;
    d.functor.flags = (applyShift ? APB_FUNCTOR_APPLY_SHIFT : 0) | APB_FUNCTOR_USE_MIXING |
                      (calculateGlobals ? APB_FUNCTOR_CALC_GLOBALS : 0);
    d.functor.cutoff = _cutoff;
    d.siteStart.push_back(0);
    if constexpr (useMixing) {
      const auto T = _ppl->getNumberRegisteredSiteTypes();
      d.mixingTable.resize(T * T * 3);
      for (size_t i = 0; i < T; ++i)
        for (size_t j = 0; j < T; ++j) {
          d.mixingTable[3 * (i * T + j) + 0] = _ppl->getMixing24Epsilon(i, j);
          d.mixingTable[3 * (i * T + j) + 1] = _ppl->getMixingSigmaSquared(i, j);
          d.mixingTable[3 * (i * T + j) + 2] = applyShift ? _ppl->getMixingShift6(i, j) : 0.;
        }
      d.functor.num_types = static_cast<int32_t>(T);
      const size_t numMolTypes = static_cast<size_t>(_ppl->getNumberRegisteredMolTypes());
      for (size_t m = 0; m < numMolTypes; ++m) {
        const auto pos = _ppl->getSitePositions(m);
        const auto types = _ppl->getSiteTypes(m);
        for (size_t k = 0; k < pos.size(); ++k) {
          d.sitePositions.insert(d.sitePositions.end(), pos[k].begin(), pos[k].end());
          d.siteTypes.push_back(static_cast<int32_t>(types[k]));
        }
        d.siteStart.push_back(static_cast<int32_t>(d.siteTypes.size()));
      }
    } else {
      const double shift6 = applyShift ? ParticlePropertiesLibrary<double, size_t>::calcShift6(_epsilon24, _sigmaSquared, _cutoff * _cutoff) : 0.;
      d.mixingTable = {_epsilon24, _sigmaSquared, shift6};
      d.functor.num_types = 1;
      for (const auto &p : _sitePositions) {
        d.sitePositions.insert(d.sitePositions.end(), p.begin(), p.end());
        d.siteTypes.push_back(0);
      }
      d.siteStart.push_back(static_cast<int32_t>(d.siteTypes.size()));
    }
    d.functor.mixing_table = d.mixingTable.data();
    d.functor.num_mol_types = static_cast<int32_t>(d.siteStart.size() - 1);
    d.functor.site_start = d.siteStart.data();
    d.functor.site_positions = d.sitePositions.data();
    d.functor.site_types = d.siteTypes.data();
    return d;
  }
  void apbDeposit(const apb_traversal_result &raw) {
    _gpuRaw.upot_sum += raw.upot_sum;
    for (int d = 0; d < 3; ++d) _gpuRaw.virial_sum[d] += raw.virial_sum[d];
  }

 private:
  Cpu _cpu;
  double _cutoff;
  ParticlePropertiesLibrary<double, size_t> *_ppl = nullptr;
  double _epsilon24 = 0., _sigmaSquared = 0.;
  std::vector<std::array<double, 3>> _sitePositions;
  apb_traversal_result _gpuRaw{};
  double _gpuUpot = 0., _gpuVirial = 0.;
  bool _postProcessed = false;
};
#endif  // AUTOPAS_B200_MULTISITE

template <class F>
concept HasGpuKernel = requires(F &f) {
  { f.apbDescribe() } -> std::same_as<FunctorDescriptor>;
  f.apbDeposit(std::declval<const apb_traversal_result &>());
};

// ---------------------------------------------------------------------------------------------------------------------
// GpuTraversal
// ---------------------------------------------------------------------------------------------------------------------
template <class Functor_T>
class GpuTraversal : public autopas::TraversalInterface, public GpuTraversalInterface {
 public:
  /// like the reference traversals (e.g. LCC08Traversal.h:40-43): functor by reference, newton3; the data layout is SoA
  GpuTraversal(int apbTraversalOption, Functor_T &functor, bool useNewton3,
               autopas::TraversalOption reportedOption = autopas::TraversalOption::vcl_cluster_iteration)
      : autopas::TraversalInterface(autopas::DataLayoutOption::soa, useNewton3),
        _traversal(apbTraversalOption), _functor(&functor), _reported(reportedOption) {}

  /// With the additive enum edits of INTEGRATION.md this returns TraversalOption::gpulc_c08 etc.; against the
  /// unmodified reference it reports the stock option passed to the constructor.
  [[nodiscard]] autopas::TraversalOption getTraversalType() const override { return _reported; }

  /// No CPU fallback: a functor without a GPU kernel, or newton3 on a newton3-off-only traversal
  /// (CompatibleTraversals.h:142-151), makes the configuration inapplicable (TraversalSelector.h:353-356).
  [[nodiscard]] bool isApplicableToDomain() const override {
    if (not functorHasGpuKernel()) return false;
    if constexpr (requires { Functor_T::apbLinkedCellsOnly; }) {  // kernels that exist for gpuLinkedCells only
      if (_traversal != APB_TRAVERSAL_GPULC_C08 and _traversal != APB_TRAVERSAL_GPULC_C18) return false;
    }
    if (_useNewton3 and (_traversal == APB_TRAVERSAL_GPUVCL_CLUSTER_ITERATION or
                         _traversal == APB_TRAVERSAL_GPUVCL_C01_BALANCED))
      return false;
    return true;
  }
  void initTraversal() override {}
  void traverseParticles() override {
    autopas::utils::ExceptionHandler::exception(
        "GpuTraversal::traverseParticles(): GPU traversals run inside GpuParticleContainer::computeInteractions()");
  }
  void endTraversal() override {}

  [[nodiscard]] int apbTraversal() const override { return _traversal; }
  [[nodiscard]] bool functorHasGpuKernel() const override { return HasGpuKernel<Functor_T>; }
  FunctorDescriptor describeFunctor() override {
    if constexpr (HasGpuKernel<Functor_T>) return _functor->apbDescribe();
    autopas::utils::ExceptionHandler::exception("GpuTraversal: functor {} has no GPU kernel", _functor->getName());
    return {};
  }
  void depositResult(const apb_traversal_result &raw) override {
    if constexpr (HasGpuKernel<Functor_T>) _functor->apbDeposit(raw);
  }

 private:
  int _traversal;
  Functor_T *_functor;
  autopas::TraversalOption _reported;
};

// ---------------------------------------------------------------------------------------------------------------------
// GpuParticleContainer
// ---------------------------------------------------------------------------------------------------------------------
template <class Particle_T>
class GpuParticleContainer : public autopas::ParticleContainerInterface<Particle_T> {
  using Base = autopas::ParticleContainerInterface<Particle_T>;
  using Columns = ParticleColumns<Particle_T>;
  static constexpr size_t kChunk = 1024;  // mirror particles per iterator "cell" (threads stride over chunks)

 public:
  /**
   * @param container APB_CONTAINER_LINKED_CELLS or APB_CONTAINER_VERLET_CLUSTER_LISTS
   * Arguments as ContainerSelector::generateContainer passes them (tuning/selectors/ContainerSelector.h:57-104).
   */
  GpuParticleContainer(int container, const std::array<double, 3> &boxMin, const std::array<double, 3> &boxMax,
                       double cutoff, double skin, double cellSizeFactor = 1.0, unsigned int clusterSize = 4,
                       int device = 0)
      : Base(skin), _boxMin(boxMin), _boxMax(boxMax), _cutoff(cutoff), _containerKind(container),
        _pending(autopas::autopas_get_max_threads()) {
    apb_config cfg{};
    for (int d = 0; d < 3; ++d) {
      cfg.box_min[d] = boxMin[d];
      cfg.box_max[d] = boxMax[d];
    }
    cfg.cutoff = cutoff;
    cfg.skin = skin;
    cfg.cell_size_factor = cellSizeFactor;
    cfg.cluster_size = static_cast<int32_t>(clusterSize);
    cfg.container = container;
    cfg.particle_kind = Columns::kind;
    cfg.device = device;
    if (apb_create(&cfg, &_h) != APB_OK) {
      autopas::utils::ExceptionHandler::exception("GpuParticleContainer: {}", apb_last_error(nullptr));
    }
  }
  ~GpuParticleContainer() override {
    if (_h) apb_destroy(_h);
  }

  /// With the edits of INTEGRATION.md applied (tools/make_autopas_overlay.py, -DAUTOPAS_B200_INTEGRATED):
  /// ContainerOption::gpuLinkedCells / gpuVerletClusterLists; against the unmodified tree the stock look-alikes.
  [[nodiscard]] autopas::ContainerOption getContainerType() const override {
#ifdef AUTOPAS_B200_INTEGRATED
    return _containerKind == APB_CONTAINER_LINKED_CELLS ? autopas::ContainerOption::gpuLinkedCells
                                                        : autopas::ContainerOption::gpuVerletClusterLists;
#else
    return _containerKind == APB_CONTAINER_LINKED_CELLS ? autopas::ContainerOption::linkedCells
                                                        : autopas::ContainerOption::verletClusterLists;
#endif
  }
  void reserve(size_t, size_t) override {}

  // ---- storage ------------------------------------------------------------------------------------------------------
 protected:
  /// thread-safe like VerletClusterLists::addParticleImpl (VerletClusterLists.h:184-192): per-thread staging vectors
  void addParticleImpl(const Particle_T &p) override { _pending[autopas::autopas_get_thread_num()].push_back(p); }
  void addHaloParticleImpl(const Particle_T &haloParticle) override {
    Particle_T copy = haloParticle;
    copy.setOwnershipState(autopas::OwnershipState::halo);
    _pending[autopas::autopas_get_thread_num()].push_back(copy);
  }

 public:
  /// ParticleContainerInterface::updateHaloParticle (:152): position of the halo copy with the same id is replaced
  bool updateHaloParticle(const Particle_T &haloParticle) override {
    syncToDevice();
    const int64_t id = static_cast<int64_t>(haloParticle.getID());
    const auto &r = haloParticle.getR();
    int64_t notFound = 0;
    check(apb_update_halo_particles(_h, 1, &id, &r[0], &r[1], &r[2], &notFound));
    _mirrorValid = false;
    return notFound == 0;
  }
  /// bulk form (one device call for a whole halo message); returns the number of particles that were not found
  size_t updateHaloParticles(const std::vector<Particle_T> &halos) {
    syncToDevice();
    std::vector<int64_t> ids(halos.size());
    std::vector<double> x(halos.size()), y(halos.size()), z(halos.size());
    for (size_t i = 0; i < halos.size(); ++i) {
      ids[i] = static_cast<int64_t>(halos[i].getID());
      x[i] = halos[i].getR()[0];
      y[i] = halos[i].getR()[1];
      z[i] = halos[i].getR()[2];
    }
    int64_t notFound = 0;
    check(apb_update_halo_particles(_h, static_cast<int64_t>(halos.size()), ids.data(), x.data(), y.data(), z.data(),
                                    &notFound));
    _mirrorValid = false;
    return static_cast<size_t>(notFound);
  }

  void deleteHaloParticles() override {
    syncToDevice();
    check(apb_delete_halo_particles(_h));
    _mirrorValid = false;
  }
  void deleteAllParticles() override {
    for (auto &v : _pending) v.clear();
    check(apb_delete_all_particles(_h));
    _hostOnly.clear();
    _mirror.clear();
    _mirrorValid = true;
    _mirrorDirty = false;
  }

  [[nodiscard]] size_t getNumberOfParticles(autopas::IteratorBehavior behavior = autopas::IteratorBehavior::owned) const override {
    const_cast<GpuParticleContainer *>(this)->syncToDevice();
    int64_t owned = 0, halo = 0;
    check(apb_get_num_particles(_h, &owned, &halo));
    size_t n = 0;
    if (behavior & autopas::IteratorBehavior::owned) n += static_cast<size_t>(owned);
    if (behavior & autopas::IteratorBehavior::halo) n += static_cast<size_t>(halo);
    if (behavior & autopas::IteratorBehavior::dummy) {
      int64_t slots = 0;
      check(apb_get_num_slots(_h, &slots));
      n += static_cast<size_t>(slots - owned - halo);
    }
    return n;
  }
  [[nodiscard]] size_t size() const override {
    const_cast<GpuParticleContainer *>(this)->syncToDevice();
    int64_t slots = 0;
    check(apb_get_num_slots(_h, &slots));
    return static_cast<size_t>(slots);
  }

  // ---- iterators: served from the host mirror -----------------------------------------------------------------------
  [[nodiscard]] autopas::ContainerIterator<Particle_T, true, false> begin(
      autopas::IteratorBehavior behavior = autopas::IteratorBehavior::ownedOrHalo,
      autopas::utils::optRef<typename autopas::ContainerIterator<Particle_T, true, false>::ParticleVecType> additionalVectors =
          std::nullopt) override {
    acquireMirror(true);
    return autopas::ContainerIterator<Particle_T, true, false>(*this, behavior, additionalVectors);
  }
  [[nodiscard]] autopas::ContainerIterator<Particle_T, false, false> begin(
      autopas::IteratorBehavior behavior = autopas::IteratorBehavior::ownedOrHalo,
      autopas::utils::optRef<typename autopas::ContainerIterator<Particle_T, false, false>::ParticleVecType> additionalVectors =
          std::nullopt) const override {
    const_cast<GpuParticleContainer *>(this)->acquireMirror(false);
    return autopas::ContainerIterator<Particle_T, false, false>(*this, behavior, additionalVectors);
  }
  [[nodiscard]] autopas::ContainerIterator<Particle_T, true, true> getRegionIterator(
      const std::array<double, 3> &lowerCorner, const std::array<double, 3> &higherCorner,
      autopas::IteratorBehavior behavior,
      autopas::utils::optRef<typename autopas::ContainerIterator<Particle_T, true, true>::ParticleVecType> additionalVectors =
          std::nullopt) override {
    acquireMirror(true);
    return autopas::ContainerIterator<Particle_T, true, true>(*this, behavior, additionalVectors, lowerCorner, higherCorner);
  }
  [[nodiscard]] autopas::ContainerIterator<Particle_T, false, true> getRegionIterator(
      const std::array<double, 3> &lowerCorner, const std::array<double, 3> &higherCorner,
      autopas::IteratorBehavior behavior,
      autopas::utils::optRef<typename autopas::ContainerIterator<Particle_T, false, true>::ParticleVecType> additionalVectors =
          std::nullopt) const override {
    const_cast<GpuParticleContainer *>(this)->acquireMirror(false);
    return autopas::ContainerIterator<Particle_T, false, true>(*this, behavior, additionalVectors, lowerCorner, higherCorner);
  }

  /// ParticleContainerInterface::getParticle (:337-352): the mirror is cut into chunks of kChunk particles that play
  /// the role of cells; a thread starts at chunk `thread id` and strides by the number of threads.
  std::tuple<const Particle_T *, size_t, size_t> getParticle(size_t cellIndex, size_t particleIndex,
                                                             autopas::IteratorBehavior behavior) const override {
    constexpr std::array<double, 3> lo{std::numeric_limits<double>::lowest(), std::numeric_limits<double>::lowest(),
                                       std::numeric_limits<double>::lowest()};
    constexpr std::array<double, 3> hi{std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                                       std::numeric_limits<double>::max()};
    return getParticleImpl<false>(cellIndex, particleIndex, behavior, lo, hi);
  }
  std::tuple<const Particle_T *, size_t, size_t> getParticle(size_t cellIndex, size_t particleIndex,
                                                             autopas::IteratorBehavior behavior,
                                                             const std::array<double, 3> &boxMin,
                                                             const std::array<double, 3> &boxMax) const override {
    return getParticleImpl<true>(cellIndex, particleIndex, behavior, boxMin, boxMax);
  }

  /// like VerletClusterLists::deleteParticle (VerletClusterLists.h:350-360): mark as dummy, storage order is kept
  bool deleteParticle(Particle_T &particle) override {
    autopas::internal::markParticleAsDeleted(particle);
    _mirrorDirty = true;
    return false;
  }
  bool deleteParticle(size_t cellIndex, size_t particleIndex) override {
    autopas::internal::markParticleAsDeleted(_mirror[cellIndex * kChunk + particleIndex]);
    _mirrorDirty = true;
    return false;
  }

  // ---- the hot path -------------------------------------------------------------------------------------------------
  void rebuildNeighborLists(autopas::TraversalInterface *traversal) override {
    auto *gpu = asGpuTraversal(traversal);
    syncToDevice();
    check(apb_rebuild_neighbor_lists(_h, gpu->apbTraversal(), traversal->getUseNewton3() ? 1 : 0));
    _mirrorValid = false;
  }
  void computeInteractions(autopas::TraversalInterface *traversal) override {
    auto *gpu = asGpuTraversal(traversal);
    syncToDevice();
    FunctorDescriptor d = gpu->describeFunctor();
    apb_traversal_result raw{};
    check(apb_compute_interactions(_h, gpu->apbTraversal(), &d.functor, traversal->getUseNewton3() ? 1 : 0, &raw));
    gpu->depositResult(raw);
    _mirrorValid = false;
  }
  [[nodiscard]] std::vector<Particle_T> updateContainer(bool keepNeighborListsValid) override {
    syncToDevice();
    int64_t nl = 0;
    check(apb_update_container(_h, keepNeighborListsValid ? 1 : 0, &nl));
    _mirrorValid = false;
    std::vector<Particle_T> leavers(static_cast<size_t>(nl));
    if (nl > 0) {
      // leavers are whole copies of the particles (LeavingParticleCollector.h:101-110): every column travels
      const auto &colIds = Columns::ids();
      const size_t nc = colIds.size();
      std::vector<std::vector<double>> c(nc, std::vector<double>(static_cast<size_t>(nl)));
      for (size_t k = 0; k < nc; ++k) check(apb_get_leaver_column(_h, colIds[k], c[k].data()));
      std::vector<int64_t> ids(nl);
      std::vector<int32_t> types(nl);
      check(apb_get_leavers(_h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ids.data(), types.data()));
      for (int64_t i = 0; i < nl; ++i) {
        Particle_T p;
        double row[Columns::maxColumns];
        for (size_t k = 0; k < nc; ++k) row[k] = c[k][i];
        Columns::scatter(p, row);
        p.setID(static_cast<size_t>(ids[i]));
        if constexpr (requires { p.setTypeId(size_t{}); }) p.setTypeId(static_cast<size_t>(types[i]));
        p.setOwnershipState(autopas::OwnershipState::owned);
        if constexpr (Columns::numHostOnly > 0) {
          const auto it = _hostOnly.find(p.getID());
          if (it != _hostOnly.end()) {
            Columns::setHostOnly(p, it->second);
            _hostOnly.erase(it);
          }
        }
        leavers[i] = p;
      }
    }
    return leavers;
  }
  [[nodiscard]] autopas::TraversalSelectorInfo getTraversalSelectorInfo() const override {
    apb_geometry g{};
    check(apb_get_geometry(_h, &g));
    return autopas::TraversalSelectorInfo(
        {static_cast<unsigned long>(g.cells_per_dim[0]), static_cast<unsigned long>(g.cells_per_dim[1]),
         static_cast<unsigned long>(g.cells_per_dim[2])},
        g.interaction_length, {g.cell_length[0], g.cell_length[1], g.cell_length[2]},
        static_cast<unsigned int>(g.cluster_size));
  }

  [[nodiscard]] const std::array<double, 3> &getBoxMax() const override { return _boxMax; }
  [[nodiscard]] const std::array<double, 3> &getBoxMin() const override { return _boxMin; }
  [[nodiscard]] double getCutoff() const override { return _cutoff; }
  void setCutoff(double cutoff) override { _cutoff = cutoff; }
  [[nodiscard]] double getVerletSkin() const override { return this->_skin; }
  [[nodiscard]] double getInteractionLength() const override { return _cutoff + this->_skin; }

  /// non-virtual helpers required by the withStaticContainerType lambdas (AutoPasDecl.h:295-545)
  template <typename Lambda>
  void forEach(Lambda forEachLambda, autopas::IteratorBehavior behavior = autopas::IteratorBehavior::ownedOrHalo) {
    for (auto it = this->begin(behavior | autopas::IteratorBehavior::forceSequential); it.isValid(); ++it) forEachLambda(*it);
  }
  template <typename Lambda, typename A>
  void reduce(Lambda reduceLambda, A &result, autopas::IteratorBehavior behavior = autopas::IteratorBehavior::ownedOrHalo) {
    for (auto it = this->begin(behavior | autopas::IteratorBehavior::forceSequential); it.isValid(); ++it) reduceLambda(*it, result);
  }
  template <typename Lambda>
  void forEachInRegion(Lambda forEachLambda, const std::array<double, 3> &lowerCorner,
                       const std::array<double, 3> &higherCorner, autopas::IteratorBehavior behavior) {
    for (auto it = this->getRegionIterator(lowerCorner, higherCorner, behavior | autopas::IteratorBehavior::forceSequential);
         it.isValid(); ++it)
      forEachLambda(*it);
  }
  template <typename Lambda, typename A>
  void reduceInRegion(Lambda reduceLambda, A &result, const std::array<double, 3> &lowerCorner,
                      const std::array<double, 3> &higherCorner, autopas::IteratorBehavior behavior) {
    for (auto it = this->getRegionIterator(lowerCorner, higherCorner, behavior | autopas::IteratorBehavior::forceSequential);
         it.isValid(); ++it)
      reduceLambda(*it, result);
  }

  /// md-flexible's checkpoint piece of this rank, formatted on the device (ParallelVtkWriter::recordParticleStates,
  /// examples/md-flexible/src/ParallelVtkWriter.cpp:55-201): the bytes the reference writer puts into
  /// "<session>_Particles_<rank>_<iteration>.vtu" when it walks this container's owned particles.
  [[nodiscard]] std::string vtkParticleRecord() {
    syncToDevice();
    int64_t bytes = 0;
    check(apb_vtk_particle_record(_h, nullptr, 0, &bytes));
    std::string record(static_cast<size_t>(bytes), '\0');
    check(apb_vtk_particle_record(_h, record.data(), bytes, &bytes));
    return record;
  }

  /// the same record into a file, as the reference writer leaves it (ParallelVtkWriter.cpp:61-70, 200); returns the bytes
  size_t writeVtkParticleRecord(const std::string &path) {
    syncToDevice();
    int64_t bytes = 0;
    check(apb_vtk_write_particle_record(_h, path.c_str(), &bytes));
    return static_cast<size_t>(bytes);
  }

  /// md-flexible's checkpoint loader for one piece (loadParticlesFromRankRecord + addParticle,
  /// examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180): the particles of the piece's bytes are parsed on the
  /// device and appended as owned particles; returns their number. A particle outside the box throws like addParticle.
  size_t loadVtkParticleRecord(const std::string &pieceBytes, bool checkInBox = true) {
    syncToDevice();
    int64_t added = 0;
    check(apb_vtk_load_particle_record(_h, pieceBytes.data(), static_cast<int64_t>(pieceBytes.size()), checkInBox ? 1 : 0, &added));
    _mirrorValid = false;
    return static_cast<size_t>(added);
  }

  /// the C handle, for device-resident extensions (apb_run_steps, apb_exchange_halos, ...)
  [[nodiscard]] apb_handle handle() {
    syncToDevice();
    _mirrorValid = false;
    return _h;
  }

 private:
  void check(int rc) const {
    if (rc != APB_OK) autopas::utils::ExceptionHandler::exception("GpuParticleContainer: {}", apb_last_error(_h));
  }
  GpuTraversalInterface *asGpuTraversal(autopas::TraversalInterface *traversal) const {
    // the reference containers dynamic_cast to their own traversal interface and throw on mismatch
    // (LinkedCells.h:576-588, VerletClusterLists.h:154-162)
    auto *gpu = dynamic_cast<GpuTraversalInterface *>(traversal);
    if (gpu == nullptr) {
      autopas::utils::ExceptionHandler::exception(
          "GpuParticleContainer: trying to use a traversal of the wrong type (not a GpuTraversal)");
    }
    return gpu;
  }

  /// staged additions -> device (bulk apb_add_particles), mutated mirror -> device. Called before every device operation.
  void syncToDevice() {
    std::lock_guard<std::mutex> lock(_mutex);
    if (_mirrorDirty) uploadMirror();
    flushPending();
  }
  void flushPending() {
    size_t total = 0;
    for (auto &v : _pending) total += v.size();
    if (total == 0) return;
    const auto &colIds = Columns::ids();
    const size_t nc = colIds.size();
    for (int pass = 0; pass < 2; ++pass) {  // owned first, then halo: two bulk calls
      const auto want = pass == 0 ? autopas::OwnershipState::owned : autopas::OwnershipState::halo;
      std::vector<std::vector<double>> col(nc);
      std::vector<int64_t> ids;
      std::vector<int32_t> types;
      double row[Columns::maxColumns];
      for (auto &v : _pending)
        for (auto &p : v) {
          if (p.getOwnershipState() != want) continue;
          Columns::gather(p, row);
          for (size_t k = 0; k < nc; ++k) col[k].push_back(row[k]);
          ids.push_back(static_cast<int64_t>(p.getID()));
          if constexpr (requires { p.getTypeId(); }) types.push_back(static_cast<int32_t>(p.getTypeId()));
          else types.push_back(0);
          // (owned particles only: a halo copy shares the id of its owner and must not overwrite the owner's values)
          if constexpr (Columns::numHostOnly > 0)
            if (pass == 0) _hostOnly[p.getID()] = Columns::hostOnly(p);
        }
      if (ids.empty()) continue;
      int64_t before = 0;
      check(apb_get_num_slots(_h, &before));
      check(apb_add_particles(_h, static_cast<int64_t>(ids.size()), col[0].data(), col[1].data(), col[2].data(), ids.data(),
                              types.data(), pass == 0 ? APB_OWN_OWNED_VALUE : APB_OWN_HALO_VALUE, 0));
      // the other columns of the appended slots: columns are transferred whole, so patch them through the mirror path
      col.erase(col.begin(), col.begin() + 3);
      _appendedExtra.push_back({static_cast<size_t>(before), std::move(col)});
    }
    for (auto &v : _pending) v.clear();
    patchAppendedColumns();
    _mirrorValid = false;
  }
  struct Appended {
    size_t first;
    std::vector<std::vector<double>> c;  // columns ids()[3 ...]
  };
  void patchAppendedColumns() {
    if (_appendedExtra.empty()) return;
    int64_t slots = 0;
    check(apb_get_num_slots(_h, &slots));
    const auto &colIds = Columns::ids();
    std::vector<double> tmp(static_cast<size_t>(slots));
    for (size_t k = 3; k < colIds.size(); ++k) {
      bool any = false;
      for (auto &a : _appendedExtra)
        for (double v : a.c[k - 3]) any = any or v != 0.;
      if (not any) continue;  // appended slots are zero-initialised by the library
      check(apb_download_column(_h, colIds[k], tmp.data()));
      for (auto &a : _appendedExtra)
        for (size_t i = 0; i < a.c[k - 3].size(); ++i) tmp[a.first + i] = a.c[k - 3][i];
      check(apb_upload_column(_h, colIds[k], tmp.data()));
    }
    _appendedExtra.clear();
  }

  /// device -> mirror if the mirror is stale; `forWriting` marks it dirty (it is handed out through mutable iterators)
  void acquireMirror(bool forWriting) {
    std::lock_guard<std::mutex> lock(_mutex);
    size_t pending = 0;
    for (auto &v : _pending) pending += v.size();
    if (pending > 0) {  // staged additions must become visible to iterators: push them (and a mutated mirror) first
      if (_mirrorDirty) uploadMirror();
      flushPending();
    }
    if (not _mirrorValid) downloadMirror();
    if (forWriting) _mirrorDirty = true;
  }
  void downloadMirror() {
    int64_t slots = 0;
    check(apb_get_num_slots(_h, &slots));
    const size_t n = static_cast<size_t>(slots);
    const auto &colIds = Columns::ids();
    const size_t nc = colIds.size();
    std::vector<std::vector<double>> c(nc);
    for (size_t k = 0; k < nc; ++k) {
      c[k].resize(n);
      if (n) check(apb_download_column(_h, colIds[k], c[k].data()));
    }
    std::vector<int64_t> ids(n);
    std::vector<int32_t> types(n), own(n);
    if (n) check(apb_download_ids(_h, ids.data(), types.data(), own.data()));
    _mirror.resize(n);
    AUTOPAS_OPENMP(parallel for schedule(static))
    for (size_t i = 0; i < n; ++i) {
      Particle_T &p = _mirror[i];
      double row[Columns::maxColumns];
      for (size_t k = 0; k < nc; ++k) row[k] = c[k][i];
      Columns::scatter(p, row);
      p.setID(static_cast<size_t>(ids[i]));
      if constexpr (requires { p.setTypeId(size_t{}); }) p.setTypeId(static_cast<size_t>(types[i]));
      p.setOwnershipState(static_cast<autopas::OwnershipState>(own[i]));
      if constexpr (Columns::numHostOnly > 0) {
        if (own[i] == APB_OWN_OWNED_VALUE) {
          const auto it = _hostOnly.find(p.getID());
          if (it != _hostOnly.end()) Columns::setHostOnly(p, it->second);
        }
      }
    }
    _mirrorValid = true;
    _mirrorDirty = false;
  }
  void uploadMirror() {
    const size_t n = _mirror.size();
    int64_t slots = 0;
    check(apb_get_num_slots(_h, &slots));
    if (static_cast<size_t>(slots) != n) {
      autopas::utils::ExceptionHandler::exception("GpuParticleContainer: host mirror and device storage diverged");
    }
    const auto &colIds = Columns::ids();
    const size_t nc = colIds.size();
    std::vector<std::vector<double>> c(nc);
    std::vector<int32_t> own(n);
    for (auto &v : c) v.resize(n);
    AUTOPAS_OPENMP(parallel for schedule(static))
    for (size_t i = 0; i < n; ++i) {
      const Particle_T &p = _mirror[i];
      double row[Columns::maxColumns];
      Columns::gather(p, row);
      for (size_t k = 0; k < nc; ++k) c[k][i] = row[k];
      own[i] = static_cast<int32_t>(p.getOwnershipState());
    }
    if constexpr (Columns::numHostOnly > 0) {  // attributes without a device column follow the particle id
      _hostOnly.clear();
      for (size_t i = 0; i < n; ++i)
        if (own[i] == APB_OWN_OWNED_VALUE) _hostOnly[_mirror[i].getID()] = Columns::hostOnly(_mirror[i]);
    }
    if (n) {
      for (size_t k = 0; k < nc; ++k) check(apb_upload_column(_h, colIds[k], c[k].data()));
      check(apb_upload_ownership(_h, own.data()));
    }
    _mirrorDirty = false;
  }

  template <bool regionIter>
  std::tuple<const Particle_T *, size_t, size_t> getParticleImpl(size_t cellIndex, size_t particleIndex,
                                                                 autopas::IteratorBehavior behavior,
                                                                 const std::array<double, 3> &boxMin,
                                                                 const std::array<double, 3> &boxMax) const {
    const size_t n = _mirror.size();
    const size_t numChunks = (n + kChunk - 1) / kChunk;
    const bool sequential = behavior & autopas::IteratorBehavior::forceSequential;
    const size_t stride = sequential ? 1 : static_cast<size_t>(autopas::autopas_get_num_threads());
    if (cellIndex == 0 and particleIndex == 0) {
      cellIndex = sequential ? 0 : static_cast<size_t>(autopas::autopas_get_thread_num());
    }
    while (cellIndex < numChunks) {
      const size_t base = cellIndex * kChunk;
      const size_t len = std::min(kChunk, n - base);
      for (; particleIndex < len; ++particleIndex) {
        const Particle_T &p = _mirror[base + particleIndex];
        if (autopas::containerIteratorUtils::particleFulfillsIteratorRequirements<regionIter>(p, behavior, boxMin, boxMax)) {
          return {&p, cellIndex, particleIndex};
        }
      }
      cellIndex += stride;
      particleIndex = 0;
    }
    return {nullptr, 0, 0};
  }

  static constexpr int32_t APB_OWN_OWNED_VALUE = 1, APB_OWN_HALO_VALUE = 2;  // OwnershipState.h:20-29

  apb_handle _h = nullptr;
  std::array<double, 3> _boxMin, _boxMax;
  double _cutoff;
  int _containerKind;
  std::vector<std::vector<Particle_T>> _pending;
  std::vector<Appended> _appendedExtra;
  std::unordered_map<size_t, std::array<double, 3>> _hostOnly;  // attributes without a device column, by particle id
  mutable std::vector<Particle_T> _mirror;
  mutable bool _mirrorValid = true;   // mirror == device (an empty container starts coherent)
  mutable bool _mirrorDirty = false;  // mirror was handed out mutable since the last upload
  mutable std::mutex _mutex;
};

}  // namespace autopas_b200
