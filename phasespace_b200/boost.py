"""FONLL-style ``boost_to`` generator: the input-side caller of the generator in the reference's own tests.

``tests/helpers/rapidsim.py:23-51`` of the reference (``generate_fonll``) draws the transverse momentum and the
pseudorapidity of the decaying particle from binned FONLL spectra (``numpy.random.choice`` over the bin centres), the
azimuth uniformly, and builds ``[px, py, pz, E]`` rows that ``GenParticle.generate`` takes as ``boost_to``.  Here the
same recipe runs on the device (``ps_fonll_boost``), keyed on ``(seed, event index)`` like every other draw of this
package, so the momenta never exist in host memory and any sharding of the event range gives identical rows.

The reference reads the spectra from ROOT files with ``uproot``; neither is available here, so the spectra are passed as
``(bin_edges, bin_contents)`` pairs (what ``numpy.histogram`` returns, in that order reversed) or ``(bin_centres,
bin_contents)`` when the two arrays have equal length.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .phasespace import _device, _ptr
from .random import SeedLike, get_rng


def _spectrum(hist, device):
    """(bin centres, normalised cumulative sum) of one binned spectrum, as device tensors (rapidsim.py:24-30)."""
    x, values = (np.asarray(a, dtype=np.float64).reshape(-1) for a in hist)
    if x.size == values.size + 1:  # bin edges
        centers = x[:-1] + (x[1:] - x[:-1]) / 2
    elif x.size == values.size:
        centers = x
    else:
        raise ValueError("a spectrum is (bin_edges, bin_contents) or (bin_centres, bin_contents)")
    if values.size == 0 or np.any(values < 0) or not np.sum(values) > 0:
        raise ValueError("bin contents must be non-negative with a positive sum")
    cdf = np.cumsum(values / np.sum(values))
    cdf /= cdf[-1]  # numpy.random.choice does the same
    return (torch.as_tensor(centers, device=device), torch.as_tensor(cdf, device=device))


def generate_fonll(mass, pt_hist, eta_hist, n_events: int, seed: SeedLike = None, *, pt_scale: float = 1000.0):
    """``(n_events, 4)`` CUDA tensor of ``[px, py, pz, E]`` momenta for ``boost_to`` (rapidsim.py:23-51).

    ``pt_hist`` / ``eta_hist``: binned spectra, see the module docstring.  ``pt_scale`` converts the pT axis to the
    mass unit (the FONLL histograms are in GeV and the reference multiplies by 1000).  The reference returns the
    transposed ``(4, n_events)`` numpy array and transposes it at the call site.
    """
    device = _device()
    n = int(n_events)
    pt_c, pt_cdf = _spectrum(pt_hist, device)
    eta_c, eta_cdf = _spectrum(eta_hist, device)
    out = torch.empty((n, 4), dtype=torch.float64, device=device)
    rng = get_rng(seed)
    first = rng.reserve(n)
    _lib.check(_lib.lib.ps_fonll_boost(
        _ptr(pt_c), _ptr(pt_cdf), int(pt_c.numel()), _ptr(eta_c), _ptr(eta_cdf), int(eta_c.numel()), float(mass),
        float(pt_scale), rng.seed, first, n, _ptr(out), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return out
