"""phasespace_b200 -- B200-native n-body phase-space event generator.

Drop-in for the event-generator path of zfit/phasespace (``nbody_decay``, ``GenParticle``,
``generate``, ``fromdecay.GenMultiDecay``): same public API (reference ``__init__.py:16``), with the
Raubold-Lynch / GENBOD generation fused into one hand-written sm_100a CUDA kernel per decay tree,
called through a C ABI (``include/phasespace_b200.h``).  There is no CPU fallback: the first use of
the API loads ``libphasespace_b200.so`` and fails loudly if it has not been built
(``python -m phasespace_b200.build``).
"""

import importlib

__all__ = ["nbody_decay", "GenParticle", "random", "to_vectors", "numpy", "kinematics", "ForbiddenDecayError"]
__version__ = "0.1.0"

_LAZY = {
    "GenParticle": ".phasespace", "nbody_decay": ".phasespace", "to_vectors": ".phasespace",
    "ForbiddenDecayError": ".phasespace", "CompiledDecay": ".phasespace",
}
_SUBMODULES = {"random", "kinematics", "phasespace", "fromdecay", "sharding", "unweight", "histogram", "boost", "_lib", "build"}


def __getattr__(name):
    # resolved on first access so that `python -m phasespace_b200.build` works before the library exists
    if name in _LAZY:
        return getattr(importlib.import_module(_LAZY[name], __name__), name)
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    if name == "numpy":  # the reference exports its tensor namespace under this name (__init__.py:18)
        import torch

        return torch
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(set(__all__) | set(_LAZY) | _SUBMODULES)
