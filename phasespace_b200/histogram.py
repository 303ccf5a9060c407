"""Fused weighted histograms: the consumer of the reference's own physics tests, without storing events.

``tests/test_physics.py:96-107`` of the reference generates events, copies them to the host and histograms every
coordinate of every particle with the event weights (``make_norm_histo``: 100 bins, momenta on (-3000, 3000), energies
on (0, 3000)), plus the weights themselves on (0, 1 + 1e-8).  ``generate_histograms`` does that inside the generator
kernel (``ps_histogram``): events are produced exactly as by ``generate`` and binned in shared memory; nothing but the
histograms is ever written to HBM, so the kernel runs at the fp64 pipe's speed rather than at the memory system's.
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib
from .phasespace import _device, _ptr, process_list_to_tensor
from .random import SeedLike, get_rng


def generate_histograms(decay, n_events: int, seed: SeedLike = None, *, bins: int = 100, p_range=(-3000.0, 3000.0),
                        e_range=(0.0, 3000.0), boost_to=None, normalize: bool = True, chunk_events: int = 1 << 30,
                        exact: bool = False, w_cap: float | None = None, first_event: int | None = None):
    """Weighted histograms of ``decay.generate(n_events, seed=seed)`` without materialising the events.

    Returns ``(particle_histos, weight_histo)``: ``particle_histos[name]`` is a ``(4, bins)`` CUDA tensor (rows px, py,
    pz, E; each event enters with its normalised weight) and ``weight_histo`` a ``(bins,)`` tensor of event counts in
    bins of the normalised weight on ``[0, 1 + 1e-8)``.  With ``normalize`` every histogram is divided by its sum,
    like the reference's ``make_norm_histo`` (tests/helpers/plotting.py:12-15).

    ``w_cap`` is the expected upper bound of the normalised weights: weights within 26 binary orders of magnitude
    below it are accumulated in exact fixed point with native shared-memory atomics, the others (still correctly)
    with slower fp64 atomics.  By default it is four times the largest weight of a 32768-event pilot sample.
    ``first_event`` (with an integer ``seed``) histograms the events ``[first_event, first_event + n_events)`` of the
    stream instead of the generator's next ones: the multi-GPU entry (``sharding.generate_histograms_sharded``).
    """
    device = _device()
    n = int(n_events)
    compiled = decay.compile()
    if compiled.n_user:
        raise ValueError("fused histograms need in-kernel mass functions (no arbitrary Python callables)")
    decay._mark_generate_called()
    rng = get_rng(seed)
    first = rng.reserve(n) if first_event is None else int(first_event)
    P = compiled.n_particles
    hist = torch.zeros((4 * P + 1, bins), dtype=torch.float64, device=device)
    err = torch.zeros((1,), dtype=torch.int32, device=device)
    boost_rows = 0
    if boost_to is not None:
        boost_to = process_list_to_tensor(boost_to, device=device).reshape(-1, 4).contiguous()
        boost_rows = boost_to.shape[0]
        if boost_rows not in (1, n):
            raise ValueError(f"The number of events requested ({n}) doesn't match the boost_to input size "
                             f"of {tuple(boost_to.shape)}")
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    flags = _lib.PS_GEN_EXACT if exact else 0
    if w_cap is None and n > 0:  # pilot: the first events of the same stream
        m = min(n, 32768)
        stats = torch.zeros((4,), dtype=torch.float64, device=device)
        pilot_boost = boost_to if boost_rows <= 1 else boost_to[:m]
        compiled.launch(m, rng.seed, first, boost_to=pilot_boost, exact=exact, stats=stats, err=err)
        w_cap = 4.0 * float(stats[0].item())
        err.zero_()  # a forbidden pilot event is reported by the real pass
    if not w_cap or not (w_cap > 0.0) or w_cap != w_cap or w_cap == float("inf"):
        w_cap = 1.0
    done = 0
    while done < n:  # chunking only bounds the launch size; the histograms accumulate
        m = min(chunk_events, n - done)
        b = boost_to[done:done + m] if boost_rows == n and n != 1 else boost_to
        _lib.check(_lib.lib.ps_histogram(
            compiled.handle, rng.seed, first + done, m, _ptr(b), (m if boost_rows == n and n != 1 else boost_rows),
            int(bins), float(p_range[0]), float(p_range[1]), float(e_range[0]), float(e_range[1]), float(w_cap), _ptr(hist),
            _ptr(err), flags, stream))
        done += m
    if n > 0 and (compiled.has_resonances or boost_to is not None) and int(err.item()):
        raise _lib.ForbiddenDecayError("Forbidden decay")
    if normalize:
        sums = hist.sum(dim=1, keepdim=True)
        hist = hist / torch.where(sums > 0, sums, torch.ones_like(sums))
    parts = {name: hist[4 * compiled.slots[name]: 4 * compiled.slots[name] + 4] for name in compiled.names}
    return parts, hist[4 * P]
