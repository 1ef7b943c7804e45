"""Host-side mirror of the reference's container / traversal / functor interface on top of the C ABI.

Names, argument meaning and error behaviour follow the reference so that the parity tests read like the reference's:
  * GpuParticleContainer  <-> autopas::ParticleContainerInterface (src/autopas/containers/ParticleContainerInterface.h)
  * GpuTraversal          <-> autopas::TraversalInterface (src/autopas/containers/TraversalInterface.h:18-84)
  * LJFunctor             <-> mdLib::LJFunctor (applicationLibrary/molecularDynamics/molecularDynamicsLibrary/LJFunctor.h)
  * ParticlePropertiesLibrary <-> ParticlePropertiesLibrary.h
The C++ drop-in (autopas_b200/shim/GpuContainers.h) is the same layer for a C++ host. All numerics happen in the CUDA
library; this module only marshals numpy arrays.
"""
import ctypes

import numpy as np

from . import capi
from .capi import ApbError


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ParticlePropertiesLibrary:
    """Per-type epsilon / sigma / mass and the mixing table (ParticlePropertiesLibrary.h:34-99, 444-473)."""

    def __init__(self, cutoff):
        self._cutoff = float(cutoff)
        self._eps, self._sigma, self._mass, self._nu = [], [], [], []
        self._table = None

    def addSiteType(self, siteId, mass):
        if siteId != len(self._mass):
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "site types must be registered with consecutive ids")
        self._mass.append(float(mass))
        self._eps.append(0.0)
        self._sigma.append(0.0)
        self._nu.append(0.0)

    def addLJParametersToSite(self, siteId, epsilon, sigma):
        self._eps[siteId] = float(epsilon)
        self._sigma[siteId] = float(sigma)

    def addATMParametersToSite(self, siteId, nu):
        self._nu[siteId] = float(nu)

    def getNumberRegisteredSiteTypes(self):
        return len(self._mass)

    def calculateMixingCoefficients(self):
        T = len(self._mass)
        if T == 0:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "calculateMixingCoefficients without registered site types")
        out = np.zeros(T * T * 3)
        rc = capi.load().apb_make_lj_mixing_table(T, _ptr(_f64(self._eps)), _ptr(_f64(self._sigma)), self._cutoff,
                                                  _ptr(out))
        if rc != 0:
            raise ApbError(rc, "apb_make_lj_mixing_table failed")
        self._table = out

    def getMixingTable(self):
        if self._table is None:
            self.calculateMixingCoefficients()
        return self._table

    def getMixing24Epsilon(self, i, j):
        return self.getMixingTable()[3 * (i * len(self._mass) + j)]

    def getMixingSigmaSquared(self, i, j):
        return self.getMixingTable()[3 * (i * len(self._mass) + j) + 1]

    def getMixingShift6(self, i, j):
        return self.getMixingTable()[3 * (i * len(self._mass) + j) + 2]

    def getMixingNuTable(self):
        """nu_ijk = cbrt(nu_i nu_j nu_k), row-major [T*T*T] (ParticlePropertiesLibrary.h:460-470)."""
        nu = np.asarray(self._nu)
        return np.cbrt(nu[:, None, None] * nu[None, :, None] * nu[None, None, :]).ravel()

    def addMolType(self, molId, siteIds, relSitePos, momentOfInertia=(1.0, 1.0, 1.0)):
        """ParticlePropertiesLibrary::addMolType: a molecule type = site types + unrotated relative site positions."""
        if not hasattr(self, "_mols"):
            self._mols = []
        if molId != len(self._mols):
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "molecule types must be registered with consecutive ids")
        pos = np.asarray(relSitePos, dtype=np.float64).reshape(-1, 3)
        if len(siteIds) != len(pos):
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "number of site types and site positions differ")
        self._mols.append((list(siteIds), pos, tuple(momentOfInertia)))

    def getMolTables(self):
        """(site_start[numMolTypes + 1], site_positions[numSites, 3], site_types[numSites]) for the C ABI."""
        mols = getattr(self, "_mols", [])
        starts = np.cumsum([0] + [len(m[0]) for m in mols])
        pos = np.concatenate([m[1] for m in mols]) if mols else np.zeros((0, 3))
        types = np.concatenate([np.asarray(m[0]) for m in mols]) if mols else np.zeros(0)
        return starts, pos, types


class LJFunctor:
    """mdLib::LJFunctor: template flags become constructor arguments (LJFunctor.h:39-41)."""

    def __init__(self, cutoff, particlePropertiesLibrary=None, applyShift=False, useMixing=False,
                 calculateGlobals=False, countFLOPs=False, virialTraceOnly=False):
        """virialTraceOnly: the device may accumulate only the sum of the three virial components, which is all
        getVirial() returns (LJFunctor.h:719); the C++ shim sets it, the parity tests cover both settings."""
        self.virialTraceOnly = bool(virialTraceOnly)
        if useMixing and particlePropertiesLibrary is None:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "Mixing without a ParticlePropertiesLibrary is not possible")
        if particlePropertiesLibrary is not None and not useMixing:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "Not using Mixing but using a ParticlePropertiesLibrary is not allowed")
        self._cutoff = float(cutoff)
        self._ppl = particlePropertiesLibrary
        self.applyShift, self.useMixing = bool(applyShift), bool(useMixing)
        self.calculateGlobals, self.countFLOPs = bool(calculateGlobals), bool(countFLOPs)
        self._epsilon24, self._sigmaSquared = 0.0, 0.0
        self._raw = capi.TraversalResult()
        self._postProcessed = False
        self._upot = 0.0
        self._virial = 0.0

    # --- Functor interface
    def getName(self):
        return "LJFunctorB200"

    def isRelevantForTuning(self):
        return True

    def allowsNewton3(self):
        return True

    def allowsNonNewton3(self):
        return True

    def getCutoff(self):
        return self._cutoff

    def setParticleProperties(self, epsilon24, sigmaSquared):
        self._epsilon24, self._sigmaSquared = float(epsilon24), float(sigmaSquared)

    def initTraversal(self):
        self._raw = capi.TraversalResult()
        self._postProcessed = False
        self._upot = 0.0
        self._virial = 0.0

    def _deposit(self, raw):
        """Called by the GPU traversal: adds the device accumulators, like a thread's _aosThreadData entry."""
        self._raw.upot_sum += raw.upot_sum
        for d in range(3):
            self._raw.virial_sum[d] += raw.virial_sum[d]
        for name in ("num_dist_calls", "num_kernel_calls_n3", "num_kernel_calls_no_n3", "num_global_calcs_n3",
                     "num_global_calcs_no_n3"):
            setattr(self._raw, name, getattr(self._raw, name) + getattr(raw, name))

    def endTraversal(self, newton3):
        if self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Already postprocessed, endTraversal(bool newton3) was called twice without "
                                           "calling initTraversal().")
        if self.calculateGlobals:
            u, v = ctypes.c_double(), ctypes.c_double()
            capi.load().apb_lj_end_traversal(ctypes.byref(self._raw), ctypes.byref(u), ctypes.byref(v))
            self._upot, self._virial = u.value, v.value
            self._postProcessed = True

    def getPotentialEnergy(self):
        if not self.calculateGlobals:
            raise ApbError(capi.ERR_STATE, "Trying to get potential energy even though calculateGlobals is false.")
        if not self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Cannot get potential energy, because endTraversal was not called.")
        return self._upot

    def getVirial(self):
        if not self.calculateGlobals:
            raise ApbError(capi.ERR_STATE, "Trying to get virial even though calculateGlobals is false.")
        if not self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Cannot get virial, because endTraversal was not called.")
        return self._virial

    def getNumFLOPs(self):
        if not self.countFLOPs:
            return 2 ** 64 - 1  # std::numeric_limits<size_t>::max() (LJFunctor.h:786-788)
        return capi.load().apb_lj_num_flops(ctypes.byref(self._raw), 1 if self.applyShift else 0)

    def getHitRate(self):
        if not self.countFLOPs or self._raw.num_dist_calls == 0:
            return float("nan")
        return (self._raw.num_kernel_calls_no_n3 + self._raw.num_kernel_calls_n3) / self._raw.num_dist_calls

    # --- marshalling
    def _c_functor(self):
        f = capi.Functor()
        f.kind = capi.FUNCTOR_LJ
        f.flags = ((capi.FLAG_APPLY_SHIFT if self.applyShift else 0) | (capi.FLAG_USE_MIXING if self.useMixing else 0) |
                   (capi.FLAG_CALC_GLOBALS if self.calculateGlobals else 0) |
                   (capi.FLAG_COUNT_FLOPS if self.countFLOPs else 0) |
                   (capi.FLAG_VIRIAL_TRACE if self.virialTraceOnly else 0))
        f.cutoff = self._cutoff
        f.epsilon24 = self._epsilon24
        f.sigma_squared = self._sigmaSquared
        if self.useMixing:
            self._table_keepalive = _f64(self._ppl.getMixingTable())
            f.num_types = self._ppl.getNumberRegisteredSiteTypes()
            f.mixing_table = self._table_keepalive.ctypes.data
        return f


class _FunctorBase:
    """Common part of the functor mirrors: accumulators deposited by the GPU traversal, initTraversal / endTraversal."""
    calculateGlobals = False
    countFLOPs = False

    def _init_base(self, cutoff):
        self._cutoff = float(cutoff)
        self._raw = capi.TraversalResult()
        self._postProcessed = False

    def isRelevantForTuning(self):
        return True

    def allowsNewton3(self):
        return True

    def allowsNonNewton3(self):
        return True

    def getCutoff(self):
        return self._cutoff

    def initTraversal(self):
        self._raw = capi.TraversalResult()
        self._postProcessed = False

    _deposit = LJFunctor._deposit

    def endTraversal(self, newton3):
        if self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Already postprocessed, endTraversal(bool newton3) was called twice without "
                                           "calling initTraversal().")
        self._postProcessed = True


class SPHCalcDensityFunctor(_FunctorBase):
    """sphLib::SPHCalcDensityFunctor (applicationLibrary/sph/SPHLibrary/SPHCalcDensityFunctor.h): functor cutoff is 0,
    the kernel support 2.5 h is enforced inside W."""

    def __init__(self):
        self._init_base(0.0)

    def getName(self):
        return "SPHDensityFunctor"

    @staticmethod
    def getNumFlopsPerKernelCall():
        return 3 + 2 * 19 + 2 + 2  # SPHCalcDensityFunctor.h:65-72, SPHKernels.cpp:11-21

    def _c_functor(self):
        f = capi.Functor()
        f.kind = capi.FUNCTOR_SPH_DENSITY
        return f


class SPHCalcHydroForceFunctor(_FunctorBase):
    """sphLib::SPHCalcHydroForceFunctor (applicationLibrary/sph/SPHLibrary/SPHCalcHydroForceFunctor.h)."""

    def __init__(self):
        self._init_base(0.0)

    def getName(self):
        return "SPHHydroForceFunctor"

    def _c_functor(self):
        f = capi.Functor()
        f.kind = capi.FUNCTOR_SPH_HYDRO
        return f


class AxilrodTellerMutoFunctor(_FunctorBase):
    """mdLib::AxilrodTellerMutoFunctor (AxilrodTellerMutoFunctor.h): triwise; template flags become arguments."""

    def __init__(self, cutoff, particlePropertiesLibrary=None, useMixing=False, calculateGlobals=False,
                 countFLOPs=False):
        if useMixing and particlePropertiesLibrary is None:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "Mixing without a ParticlePropertiesLibrary is not possible")
        self._init_base(cutoff)
        self._ppl = particlePropertiesLibrary
        self.useMixing, self.calculateGlobals, self.countFLOPs = bool(useMixing), bool(calculateGlobals), bool(countFLOPs)
        self._nu = 0.0
        self._upot = self._virial = 0.0

    def getName(self):
        return "AxilrodTellerMutoFunctorAutoVec"

    def setParticleProperties(self, nu):
        self._nu = float(nu)

    def endTraversal(self, newton3):
        super().endTraversal(newton3)
        if self.calculateGlobals:
            u, v = ctypes.c_double(), ctypes.c_double()
            capi.load().apb_atm_end_traversal(ctypes.byref(self._raw), ctypes.byref(u), ctypes.byref(v))
            self._upot, self._virial = u.value, v.value

    def getPotentialEnergy(self):
        if not self.calculateGlobals:
            raise ApbError(capi.ERR_STATE, "Trying to get potential energy even though calculateGlobals is false.")
        if not self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Cannot get potential energy, because endTraversal was not called.")
        return self._upot

    def getVirial(self):
        if not self.calculateGlobals:
            raise ApbError(capi.ERR_STATE, "Trying to get virial even though calculateGlobals is false.")
        if not self._postProcessed:
            raise ApbError(capi.ERR_STATE, "Cannot get virial, because endTraversal was not called.")
        return self._virial

    def getNumFLOPs(self):
        if not self.countFLOPs:
            return 2 ** 64 - 1
        return capi.load().apb_atm_num_flops(ctypes.byref(self._raw))

    def _c_functor(self):
        f = capi.Functor()
        f.kind = capi.FUNCTOR_ATM
        f.flags = ((capi.FLAG_USE_MIXING if self.useMixing else 0) | (capi.FLAG_CALC_GLOBALS if self.calculateGlobals else 0) |
                   (capi.FLAG_COUNT_FLOPS if self.countFLOPs else 0))
        f.cutoff = self._cutoff
        f.nu = self._nu
        if self.useMixing:
            self._table_keepalive = _f64(self._ppl.getMixingNuTable())
            f.num_types = self._ppl.getNumberRegisteredSiteTypes()
            f.mixing_table = self._table_keepalive.ctypes.data
        return f


class LJMultisiteFunctor(_FunctorBase):
    """mdLib::LJMultisiteFunctor (LJMultisiteFunctor.h): site geometry comes from the ParticlePropertiesLibrary
    (addMolType) when mixing is used, else from setParticleProperties(epsilon24, sigmaSquared, sitePositions)."""

    def __init__(self, cutoff, particlePropertiesLibrary=None, applyShift=False, useMixing=False, calculateGlobals=False):
        if useMixing and particlePropertiesLibrary is None:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "Mixing without a ParticlePropertiesLibrary is not possible")
        self._init_base(cutoff)
        self._ppl = particlePropertiesLibrary
        self.applyShift, self.useMixing, self.calculateGlobals = bool(applyShift), bool(useMixing), bool(calculateGlobals)
        self._epsilon24 = self._sigmaSquared = 0.0
        self._sitePositions = np.zeros((0, 3))
        self._upot = self._virial = 0.0

    def getName(self):
        return "LJMultisiteFunctor"

    def setParticleProperties(self, epsilon24, sigmaSquared, sitePositionsLJ):
        self._epsilon24, self._sigmaSquared = float(epsilon24), float(sigmaSquared)
        self._sitePositions = np.ascontiguousarray(sitePositionsLJ, dtype=np.float64).reshape(-1, 3)

    def endTraversal(self, newton3):
        super().endTraversal(newton3)
        if self.calculateGlobals:  # same normalisation as LJFunctor (LJMultisiteFunctor.h:725-750)
            u, v = ctypes.c_double(), ctypes.c_double()
            capi.load().apb_lj_end_traversal(ctypes.byref(self._raw), ctypes.byref(u), ctypes.byref(v))
            self._upot, self._virial = u.value, v.value

    getPotentialEnergy = AxilrodTellerMutoFunctor.getPotentialEnergy
    getVirial = AxilrodTellerMutoFunctor.getVirial

    def _c_functor(self):
        f = capi.Functor()
        f.kind = capi.FUNCTOR_LJ_MULTISITE
        f.flags = ((capi.FLAG_APPLY_SHIFT if self.applyShift else 0) | capi.FLAG_USE_MIXING |
                   (capi.FLAG_CALC_GLOBALS if self.calculateGlobals else 0))
        f.cutoff = self._cutoff
        if self.useMixing:
            table = _f64(self._ppl.getMixingTable())
            ntypes = self._ppl.getNumberRegisteredSiteTypes()
            starts, pos, types = self._ppl.getMolTables()
        else:
            shift6 = capi.load().apb_lj_calc_shift6(self._epsilon24, self._sigmaSquared, self._cutoff ** 2)
            table = _f64([self._epsilon24, self._sigmaSquared, shift6])
            ntypes = 1
            ns = len(self._sitePositions)
            starts, pos, types = np.array([0, ns]), self._sitePositions, np.zeros(ns)
        self._keep = (table, np.ascontiguousarray(starts, dtype=np.int32), _f64(np.asarray(pos).ravel()),
                      np.ascontiguousarray(types, dtype=np.int32))
        f.num_types = ntypes
        f.mixing_table = self._keep[0].ctypes.data
        f.num_mol_types = len(self._keep[1]) - 1
        f.site_start = self._keep[1].ctypes.data
        f.site_positions = self._keep[2].ctypes.data
        f.site_types = self._keep[3].ctypes.data
        return f


class GpuTraversal:
    """TraversalInterface: (traversal option, functor, newton3). Data layout is SoA only."""

    def __init__(self, traversalOption, functor, useNewton3):
        if traversalOption not in capi.TRAVERSAL_NAMES:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, f"unknown traversal option {traversalOption}")
        self.option = traversalOption
        self.functor = functor
        self.useNewton3 = bool(useNewton3)

    def getTraversalType(self):
        return self.option

    def getUseNewton3(self):
        return self.useNewton3

    def getDataLayout(self):
        return "soa"


class GpuParticleContainer:
    """gpuLinkedCells / gpuVerletClusterLists behind ParticleContainerInterface's method names."""

    def __init__(self, containerOption, boxMin, boxMax, cutoff, skin, cellSizeFactor=1.0, clusterSize=4,
                 particleKind=capi.PARTICLE_LJ, device=0):
        if containerOption not in capi.CONTAINER_NAMES:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, f"unknown container option {containerOption}")
        self._lib = capi.load()
        cfg = capi.Config()
        for d in range(3):
            cfg.box_min[d] = float(boxMin[d])
            cfg.box_max[d] = float(boxMax[d])
        cfg.cutoff, cfg.skin = float(cutoff), float(skin)
        cfg.cell_size_factor = float(cellSizeFactor)
        cfg.cluster_size = int(clusterSize)
        cfg.container = capi.CONTAINER_NAMES[containerOption]
        cfg.particle_kind = int(particleKind)
        cfg.device = int(device)
        self._cfg = cfg
        self.containerOption = containerOption
        self._h = capi._H()
        rc = self._lib.apb_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise ApbError(rc, self._lib.apb_last_error(None).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.apb_destroy(self._h)
            self._h = capi._H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ApbError(rc, self._lib.apb_last_error(self._h).decode())

    # --- getters
    def getContainerType(self):
        return self.containerOption

    def getBoxMin(self):
        return tuple(self._cfg.box_min)

    def getBoxMax(self):
        return tuple(self._cfg.box_max)

    def getCutoff(self):
        return self._cfg.cutoff

    def getVerletSkin(self):
        return self._cfg.skin

    def getInteractionLength(self):
        return self._cfg.cutoff + self._cfg.skin

    def getTraversalSelectorInfo(self):
        g = capi.Geometry()
        self._check(self._lib.apb_get_geometry(self._h, ctypes.byref(g)))
        return g

    # --- storage
    def addParticles(self, x, y, z, ids=None, types=None, checkInBox=True):
        self._add(x, y, z, ids, types, capi.OWN_OWNED, checkInBox)

    def addHaloParticles(self, x, y, z, ids=None, types=None):
        self._add(x, y, z, ids, types, capi.OWN_HALO, False)

    def _add(self, x, y, z, ids, types, own, check):
        x, y, z = _f64(x), _f64(y), _f64(z)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        types = None if types is None else np.ascontiguousarray(types, dtype=np.int32)
        self._check(self._lib.apb_add_particles(self._h, len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(ids), _ptr(types), own,
                                                1 if check else 0))

    def updateHaloParticles(self, ids, x, y, z):
        """Bulk ParticleContainerInterface::updateHaloParticle; returns the number of ids without a halo slot."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        x, y, z = _f64(x), _f64(y), _f64(z)
        nf = ctypes.c_int64()
        self._check(self._lib.apb_update_halo_particles(self._h, len(ids), _ptr(ids), _ptr(x), _ptr(y), _ptr(z),
                                                        ctypes.byref(nf)))
        return nf.value

    def deleteHaloParticles(self):
        self._check(self._lib.apb_delete_halo_particles(self._h))

    def reserve(self, numParticles, numParticlesHaloEstimate=0):
        """ParticleContainerInterface::reserve (containers/ParticleContainerInterface.h:95)."""
        self._check(self._lib.apb_reserve(self._h, int(numParticles), int(numParticlesHaloEstimate)))

    def deleteAllParticles(self):
        self._check(self._lib.apb_delete_all_particles(self._h))

    def getNumberOfParticles(self, behavior="owned"):
        o, h = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.apb_get_num_particles(self._h, ctypes.byref(o), ctypes.byref(h)))
        return {"owned": o.value, "halo": h.value, "ownedOrHalo": o.value + h.value}[behavior]

    def size(self):
        return self.getNumberOfParticles("ownedOrHalo")

    def numSlots(self):
        n = ctypes.c_int64()
        self._check(self._lib.apb_get_num_slots(self._h, ctypes.byref(n)))
        return n.value

    # --- host mirror (what the iterators expose), storage order
    def downloadColumn(self, name):
        out = np.zeros(self.numSlots())
        self._check(self._lib.apb_download_column(self._h, capi.COL[name], _ptr(out)))
        return out

    def uploadColumn(self, name, values):
        values = _f64(values)
        if len(values) != self.numSlots():
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "column length must equal the number of slots")
        self._check(self._lib.apb_upload_column(self._h, capi.COL[name], _ptr(values)))

    def downloadIds(self):
        n = self.numSlots()
        ids = np.zeros(n, dtype=np.int64)
        types = np.zeros(n, dtype=np.int32)
        own = np.zeros(n, dtype=np.int32)
        self._check(self._lib.apb_download_ids(self._h, _ptr(ids), _ptr(types), _ptr(own)))
        return ids, types, own

    def uploadOwnership(self, own):
        own = np.ascontiguousarray(own, dtype=np.int32)
        self._check(self._lib.apb_upload_ownership(self._h, _ptr(own)))

    def uploadPositions(self, x, y, z):
        self._check(self._lib.apb_upload_positions(self._h, _ptr(x), _ptr(y), _ptr(z)))

    def downloadForces(self, fx, fy, fz):
        self._check(self._lib.apb_download_forces(self._h, _ptr(fx), _ptr(fy), _ptr(fz)))

    def uploadPositionsById(self, x, y, z, idBegin=0):
        """Positions of the owned particles from host arrays indexed by particle id - idBegin."""
        self._check(self._lib.apb_upload_positions_by_id(self._h, int(idBegin), len(x), _ptr(x), _ptr(y), _ptr(z)))

    def downloadForcesById(self, fx, fy, fz, idBegin=0):
        self._check(self._lib.apb_download_forces_by_id(self._h, int(idBegin), len(fx), _ptr(fx), _ptr(fy), _ptr(fz)))

    def resetForces(self, fx=0.0, fy=0.0, fz=0.0):
        self._check(self._lib.apb_reset_forces(self._h, fx, fy, fz))

    def forcesById(self, n):
        """Forces of the non-dummy particles scattered to an (n, 3) array indexed by particle id (test helper)."""
        ids, _, own = self.downloadIds()
        out = np.zeros((n, 3))
        m = (own != capi.OWN_DUMMY) & (ids >= 0) & (ids < n)
        for d, name in enumerate(("FX", "FY", "FZ")):
            out[ids[m], d] = self.downloadColumn(name)[m]
        return out

    # --- maintenance and the hot path
    def updateContainer(self, keepNeighborListsValid):
        nl = ctypes.c_int64()
        self._check(self._lib.apb_update_container(self._h, 1 if keepNeighborListsValid else 0, ctypes.byref(nl)))
        n = nl.value
        cols = [np.zeros(n) for _ in range(6)]
        ids = np.zeros(n, dtype=np.int64)
        types = np.zeros(n, dtype=np.int32)
        if n:
            self._check(self._lib.apb_get_leavers(self._h, *[_ptr(c) for c in cols], _ptr(ids), _ptr(types)))
        self._numLeavers = n
        return {"x": cols[0], "y": cols[1], "z": cols[2], "vx": cols[3], "vy": cols[4], "vz": cols[5], "id": ids,
                "type": types}

    def serializeParticles(self, behavior="ownedOrHalo"):
        """ParticleSerializationTools::serializeParticle for every owned and / or halo particle, in storage order: a uint8
        array of 120-byte records (md-flexible's MPI wire format)."""
        mask = {"owned": 1, "halo": 2, "ownedOrHalo": 3}[behavior]
        cap = self.numSlots()
        buf = np.zeros(cap * capi.WIRE_RECORD_BYTES, dtype=np.uint8)
        n = ctypes.c_int64()
        self._check(self._lib.apb_serialize_particles(self._h, mask, _ptr(buf), cap, ctypes.byref(n)))
        return buf[: n.value * capi.WIRE_RECORD_BYTES]

    def deserializeParticles(self, data):
        """ParticleSerializationTools::deserializeParticles + addParticle / addHaloParticle by the record's ownership."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if len(data) % capi.WIRE_RECORD_BYTES:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "wire data is not a whole number of 120-byte records")
        self._check(self._lib.apb_deserialize_particles(self._h, _ptr(data), len(data) // capi.WIRE_RECORD_BYTES))

    def vtkParticleRecord(self):
        """ParallelVtkWriter::recordParticleStates (examples/md-flexible/src/ParallelVtkWriter.cpp:55-201): the bytes of
        this rank's `.vtu` piece, formatted on the device from the SoA columns (uint8 array)."""
        n = ctypes.c_int64()
        self._check(self._lib.apb_vtk_particle_record(self._h, None, 0, ctypes.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        self._check(self._lib.apb_vtk_particle_record(self._h, _ptr(buf), n.value, ctypes.byref(n)))
        return buf[: n.value]

    def writeVtkParticleRecord(self, path):
        """The same record written to `path` (device -> pinned pieces -> file); returns the number of bytes."""
        n = ctypes.c_int64()
        self._check(self._lib.apb_vtk_write_particle_record(self._h, str(path).encode(), ctypes.byref(n)))
        return n.value

    def loadVtkParticleRecord(self, data, checkInBox=True):
        """loadParticlesFromRankRecord (examples/md-flexible/src/configuration/MDFlexConfig.cpp:91-180) + addParticle: the
        particles of a `.vtu` piece (uint8 array / bytes) appended as owned particles, parsed on the device; returns
        their number."""
        data = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
        n = ctypes.c_int64()
        self._check(self._lib.apb_vtk_load_particle_record(self._h, _ptr(data), len(data), 1 if checkInBox else 0, ctypes.byref(n)))
        return n.value

    def leaverColumn(self, name):
        """Any other attribute of the particles the last updateContainer returned (they are whole copies in the
        reference, LeavingParticleCollector.h:101-110), in the same order."""
        out = np.zeros(getattr(self, "_numLeavers", 0))
        self._check(self._lib.apb_get_leaver_column(self._h, capi.COL[name], _ptr(out)))
        return out

    def rebuildNeighborLists(self, traversal):
        self._check(self._lib.apb_rebuild_neighbor_lists(self._h, capi.TRAVERSAL_NAMES[traversal.option],
                                                         1 if traversal.useNewton3 else 0))

    def computeInteractions(self, traversal):
        """ParticleContainerInterface::computeInteractions: runs the traversal and deposits the device accumulators
        into the functor; the caller brackets it with functor.initTraversal() / endTraversal(newton3)."""
        functor = traversal.functor
        cf = functor._c_functor()
        raw = capi.TraversalResult()
        self._check(self._lib.apb_compute_interactions(self._h, capi.TRAVERSAL_NAMES[traversal.option], ctypes.byref(cf),
                                                       1 if traversal.useNewton3 else 0, ctypes.byref(raw)))
        functor._deposit(raw)
        return raw

    def forceStepById(self, traversal, x, y, z, fx, fy, fz, rebuild, idBegin=0):
        """apb_force_step_by_id: one force step for host arrays indexed by particle id; deposits the accumulators into
        the functor like computeInteractions (bracket with functor.initTraversal() / endTraversal(newton3))."""
        functor = traversal.functor
        cf = functor._c_functor()
        raw = capi.TraversalResult()
        self._check(self._lib.apb_force_step_by_id(
            self._h, capi.TRAVERSAL_NAMES[traversal.option], ctypes.byref(cf), 1 if traversal.useNewton3 else 0,
            1 if rebuild else 0, int(idBegin), len(x), _ptr(x), _ptr(y), _ptr(z), _ptr(fx), _ptr(fy), _ptr(fz),
            ctypes.byref(raw)))
        functor._deposit(raw)
        return raw

    # --- device-resident simulation loop
    def integratePositions(self, dt, massOfType, globalForce=None):
        m = _f64(np.atleast_1d(massOfType))
        g = None if globalForce is None else _f64(globalForce)
        self._check(self._lib.apb_integrate_positions(self._h, float(dt), _ptr(m), len(m), _ptr(g)))

    def integrateVelocities(self, dt, massOfType):
        m = _f64(np.atleast_1d(massOfType))
        self._check(self._lib.apb_integrate_velocities(self._h, float(dt), _ptr(m), len(m)))

    # --- thermostat, dynamic-rebuild trigger, remainder traversal (LogicHandler / md-flexible pieces)
    def calcTemperature(self, massOfType):
        """Thermostat::calcTemperatureComponent: (temperature per type, particle count per type)."""
        m = _f64(np.atleast_1d(massOfType))
        t = np.zeros(len(m))
        c = np.zeros(len(m), dtype=np.int64)
        self._check(self._lib.apb_calc_temperature(self._h, _ptr(m), len(m), _ptr(t), _ptr(c)))
        return t, c

    def applyThermostat(self, massOfType, targetTemperature, deltaTemperature):
        m = _f64(np.atleast_1d(massOfType))
        self._check(self._lib.apb_apply_thermostat(self._h, _ptr(m), len(m), float(targetTemperature), float(deltaTemperature)))

    def setThermostat(self, enable, interval=1, targetTemperature=0.0, deltaTemperature=0.0):
        self._check(self._lib.apb_set_thermostat(self._h, 1 if enable else 0, int(interval), float(targetTemperature),
                                                 float(deltaTemperature)))

    def setDynamicRebuild(self, enable):
        self._check(self._lib.apb_set_dynamic_rebuild(self._h, 1 if enable else 0))

    def checkDynamicRebuild(self):
        r = ctypes.c_int32()
        self._check(self._lib.apb_check_dynamic_rebuild(self._h, ctypes.byref(r)))
        return bool(r.value)

    def getDynamicRebuildCount(self):
        n = ctypes.c_int64()
        self._check(self._lib.apb_get_dynamic_rebuild_count(self._h, ctypes.byref(n)))
        return n.value

    def computeRemainder(self, functor, x, y, z, ownership, types=None):
        """RemainderPairwiseInteractionHandler::computeRemainderInteractions for buffered particles; returns their forces
        (n, 3) and deposits the accumulators into the functor (in addition to what computeInteractions deposited)."""
        x, y, z = _f64(x), _f64(y), _f64(z)
        own = np.ascontiguousarray(ownership, dtype=np.int32)
        ty = None if types is None else np.ascontiguousarray(types, dtype=np.int32)
        n = len(x)
        f = [np.zeros(n) for _ in range(3)]
        cf = functor._c_functor()
        raw = capi.TraversalResult()
        self._check(self._lib.apb_compute_remainder(self._h, ctypes.byref(cf), n, _ptr(x), _ptr(y), _ptr(z), _ptr(ty),
                                                    _ptr(own), _ptr(f[0]), _ptr(f[1]), _ptr(f[2]), ctypes.byref(raw)))
        functor._deposit(raw)
        return np.stack(f, axis=1), raw

    def commInit(self, nranks, rank, uniqueId=None):
        buf = None if uniqueId is None else np.frombuffer(bytes(uniqueId), dtype=np.uint8).copy()
        self._check(self._lib.apb_comm_init(self._h, int(nranks), int(rank), _ptr(buf)))

    def setDecomposition(self, globalBoxMin, globalBoxMax, neighbours6, periodic3=(1, 1, 1)):
        gmin, gmax = _f64(globalBoxMin), _f64(globalBoxMax)
        nb = np.ascontiguousarray(neighbours6, dtype=np.int32)
        per = np.ascontiguousarray(periodic3, dtype=np.int32)
        self._check(self._lib.apb_set_decomposition(self._h, _ptr(gmin), _ptr(gmax), _ptr(nb), _ptr(per)))

    def migrate(self):
        s, r = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.apb_migrate(self._h, ctypes.byref(s), ctypes.byref(r)))
        return r.value

    def exchangeHalos(self):
        self._check(self._lib.apb_exchange_halos(self._h))

    def refreshHaloColumns(self, names):
        """sph-mpi's updateHaloParticles between the density and the hydro-force pass: the halo copies recorded by the
        last generating exchangeHalos() receive the owners' current values of the named columns."""
        cols = np.ascontiguousarray([capi.COL[n] for n in names], dtype=np.int32)
        self._check(self._lib.apb_refresh_halo_columns(self._h, len(cols), _ptr(cols)))

    def allreduceGlobals(self, raw):
        self._check(self._lib.apb_allreduce_globals(self._h, ctypes.byref(raw)))
        return raw

    def runSteps(self, traversal, numSteps, firstIteration, dt, massOfType, rebuildFrequency, globalForce=None,
                 wantResults=True):
        """Simulation::simulate for `numSteps` iterations on the device; returns the per-step raw accumulators."""
        m = _f64(np.atleast_1d(massOfType))
        g = None if globalForce is None else _f64(globalForce)
        p = capi.LoopParams()
        p.dt = float(dt)
        p.mass_of_type = m.ctypes.data
        p.num_types = len(m)
        p.global_force = None if g is None else g.ctypes.data
        p.rebuild_frequency = int(rebuildFrequency)
        p.traversal = capi.TRAVERSAL_NAMES[traversal.option]
        p.newton3 = 1 if traversal.useNewton3 else 0
        cf = traversal.functor._c_functor()
        res = (capi.TraversalResult * max(numSteps, 1))() if wantResults else None
        self._check(self._lib.apb_run_steps(self._h, ctypes.byref(cf), ctypes.byref(p), int(numSteps),
                                            int(firstIteration), res))
        return res

    def getLaunchCount(self):
        n = ctypes.c_int64()
        self._check(self._lib.apb_get_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def getAllocCount(self):
        n = ctypes.c_int64()
        self._check(self._lib.apb_get_alloc_count(self._h, ctypes.byref(n)))
        return n.value

    def enableLoopTiming(self, enable=True):
        self._check(self._lib.apb_enable_loop_timing(self._h, 1 if enable else 0))

    def getLoopTiming(self):
        ms = np.zeros(4)
        cnt = np.zeros(4, dtype=np.int64)
        self._check(self._lib.apb_get_loop_timing(self._h, _ptr(ms), _ptr(cnt)))
        names = ("force", "rebuild", "halo_refresh", "integrate")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(names)}

    def getStream(self):
        s = ctypes.c_void_p()
        self._check(self._lib.apb_get_stream(self._h, ctypes.byref(s)))
        return s.value

    # --- parity artefacts
    def debugCellOfSlot(self):
        out = np.zeros(self.numSlots(), dtype=np.int64)
        self._check(self._lib.apb_debug_cell_of_slot(self._h, _ptr(out)))
        return out

    def debugClusterPairs(self):
        g = self.getTraversalSelectorInfo()
        out = np.zeros((max(g.num_cluster_pairs, 1), 2), dtype=np.int64)
        self._check(self._lib.apb_debug_cluster_pairs(self._h, _ptr(out)))
        return out[:g.num_cluster_pairs]


class ParallelVtkWriter:
    """md-flexible's checkpoint writer (examples/md-flexible/src/ParallelVtkWriter.h:21-37) over the device-side record:
    same constructor arguments, same folders and file names (`<folder>/<session>/<session>_Particles_<iteration>.pvtu`,
    `<folder>/<session>/data/<session>_Particles_<rank>_<iteration>.vtu`, ParallelVtkWriter.cpp:285-296, 437-441), so that
    md-flexible's `--checkpoint` loader (MDFlexConfig.cpp:91-180) reads what it wrote."""

    def __init__(self, sessionName, outputFolder, maximumNumberOfDigitsInIteration, rank=0, numberOfRanks=1):
        import os
        self._session, self._digits, self._rank, self._ranks = sessionName, int(maximumNumberOfDigitsInIteration), int(rank), int(numberOfRanks)
        self._sessionFolder = os.path.join(outputFolder, sessionName) + "/"
        self._dataFolder = self._sessionFolder + "data/"
        # (the reference lets rank 0 create the folders and broadcasts their names, ParallelVtkWriter.cpp:22-41; without
        # that synchronisation point every rank makes sure they exist)
        os.makedirs(self._dataFolder, exist_ok=True)

    def pvtuRecord(self, currentIteration):
        lib = capi.load()
        n = ctypes.c_int64()
        args = (self._session.encode(), self._ranks, int(currentIteration), self._digits)
        if lib.apb_vtk_pvtu_record(*args, None, 0, ctypes.byref(n)) != capi.APB_OK:
            raise ApbError(capi.ERR_INVALID_ARGUMENT, "apb_vtk_pvtu_record: bad argument")
        buf = np.zeros(n.value, dtype=np.uint8)
        lib.apb_vtk_pvtu_record(*args, _ptr(buf), n.value, ctypes.byref(n))
        return buf

    def recordParticleStates(self, currentIteration, container):
        """Writes the `.pvtu` index (rank 0) and this rank's `.vtu` piece; returns the piece's path."""
        it = str(int(currentIteration)).zfill(self._digits)
        if self._rank == 0:
            self.pvtuRecord(currentIteration).tofile(f"{self._sessionFolder}{self._session}_Particles_{it}.pvtu")
        path = f"{self._dataFolder}{self._session}_Particles_{self._rank}_{it}.vtu"
        container.writeVtkParticleRecord(path)
        return path


def checkpointPieces(filename):
    """The piece files a `.pvtu` checkpoint refers to, named as md-flexible's loader reconstructs them
    (getNumPiecesInCheckpoint + loadParticlesFromRankRecord, MDFlexConfig.cpp:67-117): one per word "Piece" in the index,
    `<folder>/data/<scenario>_<rank>_<iteration>.vtu` with scenario / iteration split at the last '_' of the file name."""
    import os
    import re
    with open(filename) as f:
        numPieces = len([w for w in re.split(r"[ /.,?!\"'<>=:;\n\t\r]", f.read()) if w == "Piece"])
    folder, base = os.path.split(str(filename))
    scenario, rest = base.rsplit("_", 1)
    iteration = rest.rsplit(".", 1)[0]
    return [os.path.join(folder, "data", f"{scenario}_{rank}_{iteration}.vtu") for rank in range(numPieces)]


def loadParticlesFromCheckpoint(filename, rank, numRanks, container, checkInBox=True):
    """MDFlexConfig::loadParticlesFromCheckpoint (MDFlexConfig.cpp:648-671) into a GPU container: with as many ranks as
    pieces every rank loads its own piece, otherwise rank 0 loads all of them. Returns the number of particles added."""
    pieces = checkpointPieces(filename)
    mine = [pieces[rank]] if numRanks == len(pieces) else (pieces if rank == 0 else [])
    return sum(container.loadVtkParticleRecord(np.fromfile(p, dtype=np.uint8), checkInBox) for p in mine)
