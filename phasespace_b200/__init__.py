"""phasespace_b200 -- B200-native n-body phase-space event generator.

Drop-in for the event-generator path of zfit/phasespace (``nbody_decay``, ``GenParticle``,
``generate``, ``fromdecay.GenMultiDecay``): same public API (reference ``__init__.py:16``), with the
Raubold-Lynch / GENBOD generation fused into one hand-written sm_100a CUDA kernel per decay tree,
called through a C ABI (``include/phasespace_b200.h``).  There is no CPU fallback.
"""

__all__ = ["nbody_decay", "GenParticle", "random", "to_vectors", "numpy", "kinematics", "ForbiddenDecayError"]

import torch as numpy  # noqa: F401  (the reference exports its tensor namespace under this name, __init__.py:18)

from . import kinematics, random  # noqa: E402
from .phasespace import ForbiddenDecayError, GenParticle, nbody_decay, to_vectors  # noqa: E402

__version__ = "0.1.0"
