"""B200-native ensemble Metropolis sampler with the MetropolisEngine API.

``import metropolisengine_b200 as me; me.MetropolisEngine(...)`` mirrors ``import metropolisengine as me`` of the
reference (README.md:15,33; package surface metropolisengine/__init__.py:1).
"""
from .engine import BuiltinEnergy, CudaEnergy, MetropolisEngine, adaptation_constants  # noqa: F401
from .engine_shared import SharedCovarianceEngine, SharedEnergy  # noqa: F401

__all__ = ["MetropolisEngine", "SharedCovarianceEngine", "SharedEnergy", "BuiltinEnergy", "CudaEnergy", "adaptation_constants"]
