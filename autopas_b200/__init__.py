"""autopas_b200 — B200-native short-range interaction path behind AutoPas' container / traversal / functor interface.

The product is the C-ABI CUDA library ``libautopas_b200.so`` (sources in ``csrc/``, header ``include/autopas_b200.h``).
This package is the Python host mirror used by tests and benchmarks; ``shim/`` holds the C++ drop-in classes.
"""
from . import capi
from .capi import ApbError
from .containers import (AxilrodTellerMutoFunctor, GpuParticleContainer, GpuTraversal, LJFunctor, LJMultisiteFunctor,
                         ParallelVtkWriter, ParticlePropertiesLibrary, SPHCalcDensityFunctor, SPHCalcHydroForceFunctor,
                         checkpointPieces, loadParticlesFromCheckpoint)

__all__ = ["capi", "ApbError", "GpuParticleContainer", "GpuTraversal", "LJFunctor", "ParticlePropertiesLibrary",
           "SPHCalcDensityFunctor", "SPHCalcHydroForceFunctor", "AxilrodTellerMutoFunctor", "LJMultisiteFunctor",
           "ParallelVtkWriter", "checkpointPieces", "loadParticlesFromCheckpoint"]
