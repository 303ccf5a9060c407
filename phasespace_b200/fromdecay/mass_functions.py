"""Mass functions for resonant particles (mirror of the reference's fromdecay/mass_functions.py).

Each factory returns a callable ``f(min_mass, max_mass, n_events)`` giving one mass per event inside
the per-event limits, as in the reference (mass_functions.py:16-109).  Upstream these wrap zfit /
zfit_physics PDFs; here the distributions are sampled on the device by inverse CDF
(csrc/ps_samplers.cuh).  The callables carry a ``_ps_sampler`` descriptor: when such a callable is
the mass of a ``GenParticle`` the sampling is fused into the generator kernel instead of being a
separate call.
"""

from __future__ import annotations

import ctypes

import torch

from .. import _lib
from ..random import get_rng


def _make(kind: int, name: str, mass, width):
    mass, width = float(mass), float(width)

    def sampler(min_mass, max_mass, n_events, seed=None):
        if not torch.cuda.is_available():
            raise RuntimeError("phasespace_b200 mass functions need a CUDA device: there is no CPU implementation.")
        device = torch.device("cuda", torch.cuda.current_device())
        n = int(n_events)
        lo = torch.as_tensor(min_mass, dtype=torch.float64, device=device).reshape(-1).expand(n).contiguous()
        hi = torch.as_tensor(max_mass, dtype=torch.float64, device=device).reshape(-1).expand(n).contiguous()
        out = torch.empty((n,), dtype=torch.float64, device=device)
        rng = get_rng(seed)
        first = rng.reserve(n)
        _lib.check(_lib.lib.ps_sample_mass(
            kind, mass, width, ctypes.c_void_p(lo.data_ptr()), ctypes.c_void_p(hi.data_ptr()), n, rng.seed, first,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    sampler.__name__ = sampler.__qualname__ = name
    sampler._ps_sampler = (kind, mass, width)
    return sampler


def gauss_factory(mass, width):
    """Truncated Gaussian mass distribution (mass_functions.py:16-40)."""
    return _make(_lib.PS_MASS_GAUSS, "gauss", mass, width)


def breitwigner_factory(mass, width):
    """Breit-Wigner (Cauchy, gamma = width) mass distribution (mass_functions.py:43-67)."""
    return _make(_lib.PS_MASS_BW, "bw", mass, width)


def relativistic_breitwigner_factory(mass, width):
    """Relativistic Breit-Wigner mass distribution (mass_functions.py:70-102)."""
    return _make(_lib.PS_MASS_RELBW, "relbw", mass, width)


DEFAULT_CONVERTER = {
    "gauss": gauss_factory,
    "bw": breitwigner_factory,
    "relbw": relativistic_breitwigner_factory,
}
