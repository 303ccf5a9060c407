"""Accept-reject unweighting fused with the generator (BASELINE config 5).

The consumer of weighted phase-space events usually unweights them (accept event ``i`` when
``u_i * w_bound < w_i / w_max``).  Storing every candidate first would cost 8 + 32 P bytes per
*generated* event; instead

1. ``ps_unweight_mask``   recomputes only the weights (no 4-vectors are formed or stored), draws the
                          acceptance uniform from Philox stream 1 and emits one bit per event,
2. ``ps_unweight_scan`` / ``ps_unweight_index``  turn that into the ordered list of accepted event indices,
3. ``ps_generate_indexed`` regenerates exactly those events (the RNG is counter based) into a dense slab.

The result is ordered by event index and independent of grid size, chunking and GPU count.
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib
from .phasespace import _device, _ptr

_CAPACITY_STEP = 1 << 16  # events


class UnweightingOverflowError(RuntimeError):
    """``w / w_max`` exceeded ``w_bound`` for some candidates: the accepted sample would be biased."""


def generate_unweighted(decay, n_events: int, seed: int, *, w_bound: float = 1.0, first_event: int = 0,
                        exact: bool = False, return_overflow: bool = False, on_overflow: str = "raise"):
    """Unweighted events among the candidates ``[first_event, first_event + n_events)`` of stream ``seed``.

    ``w_bound`` is the acceptance bound on the *normalised* weight ``w / w_max`` (1.0 is always safe; a
    smaller measured maximum raises the efficiency).  Returns ``(particles, event_index, weights)``:
    ``particles`` maps names to ``(n_accepted, 4)`` tensors, ``event_index`` the accepted global event
    indices (int64, ascending) and ``weights`` their normalised weights.

    A candidate with ``w / w_max > w_bound`` would have to be accepted with probability above one, so a bound that
    is too small biases the sample.  The mask pass counts those candidates; ``on_overflow`` says what a non-zero
    count does: ``"raise"`` (default) raises :class:`UnweightingOverflowError`, ``"warn"`` warns, ``"ignore"``
    nothing.  With ``return_overflow=True`` the count is returned as a fourth value (and never raises).
    """
    if on_overflow not in ("raise", "warn", "ignore"):
        raise ValueError("on_overflow must be 'raise', 'warn' or 'ignore'")
    device = _device()
    n = int(n_events)
    compiled = decay.compile()
    if compiled.n_user:
        raise ValueError("unweighting needs in-kernel mass functions (no arbitrary Python callables)")
    decay._mark_generate_called()
    lib = _lib.lib
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    n_tiles = (n + _lib.TILE - 1) // _lib.TILE
    mask = torch.empty(((n + 31) // 32,), dtype=torch.int32, device=device)
    offsets = torch.empty((max(n_tiles, 1),), dtype=torch.int64, device=device)
    # [accepted count (int64) | error flag (int32, low half of the word) | overflow count]: one read-back for all
    state = torch.zeros((3,), dtype=torch.int64, device=device)
    total, err, over = state[:1], state[1:2].view(torch.int32)[:1], state[2:]
    flags = _lib.PS_GEN_EXACT if exact else 0
    seed = int(seed) & (2**64 - 1)
    if n > 0:
        _lib.check(lib.ps_unweight_mask(compiled.handle, seed, int(first_event), n, float(w_bound), _ptr(mask),
                                        _ptr(over), _ptr(err), flags, stream))
        _lib.check(lib.ps_unweight_scan(_ptr(mask), n, _ptr(offsets), _ptr(total), stream))
    n_acc, failed, n_over = state.tolist()
    if failed:
        raise _lib.ForbiddenDecayError("Forbidden decay")
    if n_over and not return_overflow and on_overflow != "ignore":
        msg = (f"{n_over} of {n} candidates have w / w_max > w_bound = {w_bound!r}: the bound is too small and "
               "the accepted sample would be biased; raise w_bound (1.0 is always safe)")
        if on_overflow == "raise":
            raise UnweightingOverflowError(msg)
        import warnings

        warnings.warn(msg, RuntimeWarning, stacklevel=2)
    # Capacity rounded up to a multiple of 65536 events: consecutive chunks of a long run accept slightly different
    # numbers of events, and buffers of a size seen before come out of torch's caching allocator instead of cudaMalloc
    # (which costs more than a millisecond for the 3 GB slab of a 10-body chunk).
    cap = -(-n_acc // _CAPACITY_STEP) * _CAPACITY_STEP if n_acc > _CAPACITY_STEP else n_acc
    index = torch.empty((cap,), dtype=torch.int64, device=device)[:n_acc]
    weights = torch.empty((cap,), dtype=torch.float64, device=device)[:n_acc]
    slab = torch.empty((compiled.n_particles, cap, 4), dtype=torch.float64, device=device)
    if n_acc > 0:
        _lib.check(lib.ps_unweight_index(_ptr(mask), _ptr(offsets), int(first_event), n, _ptr(index), stream))
        _lib.check(lib.ps_generate_indexed(
            compiled.handle, seed, _ptr(index), n_acc, None, 0, None, _ptr(weights), None, _ptr(slab),
            int(4 * max(cap, 1)), None, _ptr(err), flags | _lib.PS_GEN_NORMALIZE, stream))
    slab = slab[:, :n_acc]
    parts = {name: slab[compiled.slots[name]] for name in compiled.names}
    if return_overflow:
        return parts, index, weights, int(n_over)
    return parts, index, weights
