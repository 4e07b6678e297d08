"""coddiwomple_b200 — B200-native SMC annealing hot path of choderalab/coddiwomple behind the reference's Python API.

Layout: csrc/ (CUDA kernels + C ABI, built into libcdw_b200.so), _lib (ctypes binding), system (OpenMM XML -> term
tables), engine (device-resident ensemble + fused loop), and modules named after the reference's
(particles, states, propagators, resamplers, distribution_factories, utils, smc, reporters, openmm/*); parallel (the loop
sharded over the GPUs of one box) and csmc (the host-side zeroth step of controlled SMC).
"""
from .particles import Particle, ParticleList  # noqa: F401
from .states import ParticleState, PDFState  # noqa: F401
from .propagators import Propagator  # noqa: F401
from .reporters import Reporter  # noqa: F401

__version__ = "0.1.0"
