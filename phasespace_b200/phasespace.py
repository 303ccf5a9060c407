"""Host-side mirror of the reference's public API for the event-generator path.

Same names, argument meaning and error behaviour as ``src/phasespace/phasespace.py`` of
zfit/phasespace (cited below as ps.py), but ``generate`` compiles the decay tree into a descriptor
and makes ONE call into the C ABI (``ps_generate``, include/phasespace_b200.h), which launches one
fused sm_100a kernel for the whole tree.  Returned tensors are ``torch`` CUDA fp64 tensors
(DLPack-exportable) instead of ``tf.Tensor``.
"""

from __future__ import annotations

import ctypes
import functools
import inspect
from collections.abc import Callable
from typing import NoReturn

import numpy as np
import torch

from . import _lib
from ._lib import ForbiddenDecayError
from .random import SeedLike, get_rng

__all__ = ["GenParticle", "nbody_decay", "to_vectors", "ForbiddenDecayError", "CompiledDecay"]


_CUDA_CHECKED = False
_DEVICES: dict = {}


def _device() -> torch.device:
    """The current CUDA device.  (torch.cuda.is_available() costs microseconds per call: asked once per process; after
    that the device index comes from the binding underneath torch.cuda.current_device(), without its lazy-init check.)"""
    global _CUDA_CHECKED
    if not _CUDA_CHECKED:
        if not torch.cuda.is_available():
            raise RuntimeError(
                "phasespace_b200 needs a CUDA device (B200, sm_100a): there is no CPU implementation of the generator."
            )
        torch.cuda.current_device()  # initialises torch's CUDA state
        _CUDA_CHECKED = True
    index = _current_device_index()
    dev = _DEVICES.get(index)
    if dev is None:
        dev = _DEVICES[index] = torch.device("cuda", index)
    return dev


_current_device_index = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


try:  # the raw cudaStream_t of torch's current stream without building a Stream object (what Triton's launcher uses)
    _raw_current_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    def _raw_current_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def _check_out(t, name, shape, device, dtype=torch.float64, pinned_host=False, planes=False):
    """A caller-supplied output buffer goes to the C ABI as a bare pointer: anything but the exact shape, dtype,
    device and a dense layout would make the kernel (or the D2H copies) write out of bounds."""
    if t is None:
        return
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"out[{name!r}] must be a torch tensor, not {type(t).__name__}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"out[{name!r}] must have shape {tuple(shape)}, not {tuple(t.shape)}")
    if t.dtype != dtype:
        raise ValueError(f"out[{name!r}] must have dtype {dtype}, not {t.dtype}")
    if planes and t.ndim == 3 and t.shape[0] > 1 and t.shape[1] > 0:
        # (P, n, 4): dense (n, 4) planes; the distance between planes may exceed 4 n (a slab with spare capacity)
        if t.stride(2) != 1 or t.stride(1) != 4 or t.stride(0) < 4 * t.shape[1] or t.stride(0) % 4:
            raise ValueError(f"out[{name!r}] must consist of dense (n, 4) planes at a stride that is a multiple of 4 "
                             f"and at least 4 n; strides are {tuple(t.stride())}")
    elif not t.is_contiguous():
        raise ValueError(f"out[{name!r}] must be contiguous")
    if pinned_host:
        if t.device.type != "cpu":
            raise ValueError(f"out[{name!r}] must be a CPU tensor, not on {t.device}")
    elif t.device != device:
        raise ValueError(f"out[{name!r}] must be on {device}, not on {t.device}")


def _structure(part):
    """Identity of a decay tree: which GenParticle objects hang where (masses cannot change after construction)."""
    return (id(part), tuple(_structure(c) for c in part.children))


# Bumped by every set_children anywhere: a compiled tree is re-validated (by its structure) only when some tree in the
# process has changed since it was compiled -- not on every generate() call.
_TREE_EPOCH = 0


def process_list_to_tensor(lst, device=None):
    """ps.py:37-53 -- a Python *list* is transposed (read as ``(4, N)``), anything else is taken as is."""
    if isinstance(lst, list):
        return torch.as_tensor(np.asarray(lst, dtype=np.float64), device=device).T.contiguous()
    if isinstance(lst, torch.Tensor):
        return lst.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(lst, dtype=np.float64), device=device)


def _ptr(t):
    """Device (or host) address for a ``c_void_p`` argument: ctypes converts a plain int itself."""
    return t.data_ptr() if t is not None else None


class CompiledDecay:
    """A decay tree compiled to a ``ps_tree`` handle plus the slot of every named particle.

    This is the object the benchmark and the sharding helpers drive directly; ``GenParticle.generate``
    is a thin wrapper around :meth:`launch`.
    """

    def __init__(self, top: "GenParticle", split: bool):
        self.top = top
        self.split = split  # True: every resonance is pre-sampled on the device by its Python callable
        self.particles: list[GenParticle] = []
        descs = []

        def add(part, parent):
            idx = len(self.particles)
            self.particles.append(part)
            kind, mass, width = _lib.PS_MASS_FIXED, 0.0, 0.0
            if parent == -1 and not part.has_fixed_mass:
                # ps.py:534-543: a resonance may be the generating particle when boost_to is given -- its mass is then
                # kin.mass(boost_to), evaluated per event by the BOOST kernel, and the mass function is never called
                mass = float("nan")
            elif part.has_fixed_mass:
                mass = float(np.asarray(part._mass_val, dtype=np.float64).reshape(-1)[0])
            elif split:
                kind = _lib.PS_MASS_USER
            else:
                kind, mass, width = part._mass._ps_sampler
            descs.append((parent, kind, mass, width))
            for child in part.children:
                add(child, idx)

        add(top, -1)
        arr = (_lib.NodeDesc * len(descs))(*[_lib.NodeDesc(*d) for d in descs])
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib.ps_tree_create(arr, len(descs), ctypes.byref(handle)))
        self.handle = handle
        lib = _lib.lib
        self.n_particles = lib.ps_tree_n_particles(handle)
        self.n_draws = lib.ps_tree_n_draws(handle)
        self.n_user = lib.ps_tree_n_user(handle)
        self.w_max = lib.ps_tree_w_max(handle)
        self.signature = lib.ps_tree_signature(handle).decode()
        self.slots = {p.name: lib.ps_tree_slot(handle, i) for i, p in enumerate(self.particles) if i > 0}
        self.names = sorted(self.slots, key=self.slots.get)
        self.user_columns = {i: lib.ps_tree_user_column(handle, i) for i in range(len(descs))}
        self.has_resonances = any(d[1] != _lib.PS_MASS_FIXED for d in descs)
        self.top_is_resonance = not top.has_fixed_mass
        self.structure = _structure(top)
        self.epoch = _TREE_EPOCH
        self._err_flags = {}
        self._prepared = set()
        self._static_forbidden = self._check_static()
        # the plain call of generate(): no per-event failure mode, no Python mass functions, nothing forbidden
        self.plain_ok = (not self.has_resonances and not self.top_is_resonance and self.n_user == 0
                         and self._static_forbidden is None)
        # generate_fast: masses fixed or sampled in the kernel, a top particle that can decay at rest
        self.fast_ok = not self.top_is_resonance and self.n_user == 0 and self._static_forbidden is None
        self._slot_items = tuple((name, self.slots[name]) for name in self.names)
        self._generate_c = _lib.lib.ps_generate

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib.ps_tree_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _check_static(self):
        """Forbidden decays that are known without generating (all masses fixed, ps.py:357-363)."""
        for part in self.particles:
            if part.children and part.has_fixed_mass and all(c.has_fixed_mass for c in part.children):
                total = 0.0
                for c in part.children:
                    total = total + float(c._mass_val)
                if not (float(part._mass_val) - total >= 0.0):
                    return part.name
        return None

    def prepare(self, boost: bool = False, debug: bool = False, exact: bool = False, device_index: int = -1):
        """Resolve the generator kernel ahead of the first launch (``ps_tree_prepare``): for a tree without an
        ahead-of-time kernel that is the NVRTC compilation.  ctypes releases the GIL for the call, so several trees can
        be prepared from several threads at once (``prepare_all``)."""
        key = (bool(boost), bool(debug), bool(exact), int(device_index))
        if key not in self._prepared:
            flags = _lib.PS_GEN_EXACT if exact else 0
            _lib.check(_lib.lib.ps_tree_prepare(self.handle, int(boost), int(debug), flags, int(device_index)))
            self._prepared.add(key)

    # ------------------------------------------------------------------ resonances sampled by Python callables
    def presample(self, n_events: int, top_mass, device):
        """Masses of every non-fixed particle, evaluated in the reference's order (ps.py:335-354):
        ``max_mass`` starts at parent mass minus the fixed children and shrinks by each sampled sibling."""
        out = torch.empty((n_events, self.n_user), dtype=torch.float64, device=device)
        index = {id(p): i for i, p in enumerate(self.particles)}

        def recurse_stable(part):  # ps.py:326-333
            total = 0.0
            for child in part.children:
                total = total + (float(child._mass_val) if child.has_fixed_mass else recurse_stable(child))
            return total

        def node(part, parent_mass):
            fixed = 0.0
            for c in part.children:
                if c.has_fixed_mass:
                    fixed = fixed + float(c._mass_val)
            max_mass = parent_mass - fixed
            if not isinstance(max_mass, torch.Tensor):
                max_mass = torch.full((n_events,), float(max_mass), dtype=torch.float64, device=device)
            sampled = {}
            for c in part.children:
                if c.has_fixed_mass:
                    continue
                min_mass = torch.full((n_events,), recurse_stable(c), dtype=torch.float64, device=device)
                mass = c.get_mass(min_mass, max_mass, n_events)
                mass = torch.as_tensor(mass, dtype=torch.float64, device=device).reshape(n_events)
                max_mass = max_mass - mass
                out[:, self.user_columns[index[id(c)]]] = mass
                sampled[id(c)] = mass
            for c in part.children:
                if c.children:
                    node(c, sampled[id(c)] if id(c) in sampled else float(c._mass_val))

        node(self.top, top_mass)
        return out

    def generate_plain(self, n: int, seed: int, first_event: int, device):
        """``generate(n)`` of a tree at rest with fixed masses, normalised weights: one allocation
        ``[particles (P, n, 4) | weights (n)]``, one C-ABI call, the views.  (The general path is :meth:`launch`.)"""
        n_part = self.n_particles
        buf = torch.empty(((4 * n_part + 1) * n,), dtype=torch.float64, device=device)
        base = buf.data_ptr()
        err = self._err_flags.get(device)
        if err is None:
            err = self._err_flags[device] = torch.zeros((1,), dtype=torch.int32, device=device)
        rc = self._generate_c(self.handle, seed, first_event, n, None, 0, None, None, base + 32 * n_part * n, None, base,
                              4 * n, None, err.data_ptr(), _lib.PS_GEN_NORMALIZE, _raw_current_stream(device.index), 1.0)
        if rc:
            _lib.check(rc)
        # one view per output straight off the buffer (as_strided: half the cost of slicing, reshaping and indexing)
        plane = 4 * n
        return (torch.as_strided(buf, (n,), (1,), n_part * plane),
                {name: torch.as_strided(buf, (n, 4), (4, 1), slot * plane) for name, slot in self._slot_items})

    def generate_fast(self, n: int, seed: int, first_event: int, device, err, normalize: bool = True,
                      weight_scale: float = 1.0):
        """The same single-allocation, single-call path for any tree whose masses are fixed or sampled in the kernel
        (``fast_ok``), at rest: resonances set the caller's sticky ``err`` flag instead of being checked here, the weights
        may carry ``weight_scale`` or come unnormalised with their maximum weights.  This is what ``GenMultiDecay`` issues
        per decay mode -- dozens of small launches per call, where the general path's per-launch host work (an error flag
        of its own, a stream context, indexed views) cost four times the kernels.
        Returns ``(weights, weights_max | None, {name: (n, 4)})``."""
        n_part = self.n_particles
        extra = 1 if normalize else 2
        buf = torch.empty(((4 * n_part + extra) * n,), dtype=torch.float64, device=device)
        base = buf.data_ptr()
        plane = 4 * n
        w_ptr = base + 8 * n_part * plane
        rc = self._generate_c(self.handle, seed, first_event, n, None, 0, None, None, w_ptr,
                              None if normalize else w_ptr + 8 * n, base, plane, None, err.data_ptr(),
                              _lib.PS_GEN_NORMALIZE if normalize else 0, _raw_current_stream(device.index), weight_scale)
        if rc:
            _lib.check(rc)
        return (torch.as_strided(buf, (n,), (1,), n_part * plane),
                None if normalize else torch.as_strided(buf, (n,), (1,), n_part * plane + n),
                {name: torch.as_strided(buf, (n, 4), (4, 1), slot * plane) for name, slot in self._slot_items})

    # ------------------------------------------------------------------ the one C-ABI call
    def launch(self, n_events: int, seed: int, first_event: int = 0, boost_to=None, normalize_weights=True,
               exact=False, debug_uniforms=None, user_masses=None, out=None, stats=None, err=None, stream=None,
               _device_hint=None, weight_scale=1.0):
        """Generate events ``[first_event, first_event + n_events)`` of stream ``seed`` on the current device.

        ``out`` may hold preallocated ``weights`` ``(N,)``, ``particles`` ``(P, N, 4)`` and
        ``weights_max`` ``(N,)`` tensors to reuse.  Returns ``(weights, weights_max | None, particles, err)``.
        ``weight_scale`` (normalised weights of a tree with an event-independent maximum weight only) is multiplied
        into the weights inside the kernel at no cost.
        """
        if self.top_is_resonance and boost_to is None:
            raise ValueError("Cannot use resonance as top particle")  # ps.py:543
        device = _device_hint or _device()
        n = int(n_events)
        if out:
            weights, particles = out.get("weights"), out.get("particles")
            weights_max = out.get("weights_max") if not normalize_weights else None
            _check_out(weights, "weights", (n,), device)
            _check_out(particles, "particles", (self.n_particles, n, 4), device, planes=True)
            _check_out(weights_max, "weights_max", (n,), device)
        else:
            weights = particles = weights_max = None
        if stats is not None:
            _check_out(stats, "stats", (4,), device)
        if err is not None:
            _check_out(err, "err", (1,), device, dtype=torch.int32)
        if weights is None and particles is None and weights_max is None:
            # one allocation for everything: [particles (P,n,4) | weights (n) | weights_max (n)]
            extra = 1 if normalize_weights else 2
            buf = torch.empty(((4 * self.n_particles + extra) * n,), dtype=torch.float64, device=device)
            particles = buf[: 4 * self.n_particles * n].view(self.n_particles, n, 4)
            weights = buf[4 * self.n_particles * n: (4 * self.n_particles + 1) * n]
            if not normalize_weights:
                weights_max = buf[(4 * self.n_particles + 1) * n:]
        else:
            if weights is None:
                weights = torch.empty((n,), dtype=torch.float64, device=device)
            if particles is None:
                particles = torch.empty((self.n_particles, n, 4), dtype=torch.float64, device=device)
            if not normalize_weights and weights_max is None:
                weights_max = torch.empty((n,), dtype=torch.float64, device=device)
        if err is None:
            # trees that cannot fail per event never read the flag back: share one per device
            err = self._err_flags.get(device)
            if err is None or self.has_resonances or boost_to is not None:
                err = torch.zeros((1,), dtype=torch.int32, device=device)
                if not self.has_resonances:
                    self._err_flags.setdefault(device, err)
        boost_rows = 0
        if boost_to is not None:
            boost_rows = int(boost_to.shape[0])
            if boost_to.data_ptr() % 32:
                boost_to = boost_to.clone()
        flags = (_lib.PS_GEN_NORMALIZE if normalize_weights else 0) | (_lib.PS_GEN_EXACT if exact else 0)
        if stream is None:
            stream = _raw_current_stream(device.index)
        stride = particles.stride(0) if self.n_particles > 1 and n > 0 else 4 * max(n, 1)
        rc = _lib.lib.ps_generate(
            self.handle, int(seed) & (2**64 - 1), int(first_event), n, _ptr(boost_to), boost_rows,
            _ptr(debug_uniforms), _ptr(user_masses), _ptr(weights), _ptr(weights_max), _ptr(particles),
            int(stride), _ptr(stats), _ptr(err), flags, stream, float(weight_scale),
        )
        _lib.check(rc)
        return weights, weights_max, particles, err


def prepare_all(compiled, boost=False, exact=False, device_index=None, max_workers=None):
    """Prepare the kernels of several compiled trees concurrently (one NVRTC compilation per tree that has no
    ahead-of-time kernel, one to three seconds each): the reference retraces one ``tf.function`` per decay mode one after
    the other; here the modes of a ``GenMultiDecay`` compile side by side on the host's cores."""
    if device_index is None:
        device_index = torch.cuda.current_device()
    todo = [c for c in compiled if (bool(boost), False, bool(exact), int(device_index)) not in c._prepared]
    if not todo:
        return
    if len(todo) == 1:
        todo[0].prepare(boost, False, exact, device_index)
        return
    import concurrent.futures
    import os

    workers = max_workers or min(len(todo), os.cpu_count() or 4, 16)
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        list(pool.map(lambda c: c.prepare(boost, False, exact, device_index), todo))


class GenParticle:
    """Representation of a particle (reference: ps.py:74-226).

    Instances are combined into decay chains with :meth:`set_children`; :meth:`generate` produces
    phase-space events for the chain.  ``mass`` is a number (fixed mass) or a callable
    ``mass(min_mass, max_mass, n_events)`` returning per-event masses (a resonance).  Callables made
    by ``phasespace_b200.fromdecay.mass_functions`` are sampled inside the generator kernel; any
    other callable is evaluated on the device with torch tensors before the kernel runs.
    """

    def __init__(self, name: str, mass: Callable | int | float | np.typing.ArrayLike) -> None:  # ps.py:100-119
        self.name = name
        self.children = []
        self._mass_val = mass
        if callable(mass):

            @functools.wraps(mass)
            def mass_preprocessed(*args, mass=mass, **kwargs):
                return torch.atleast_1d(torch.as_tensor(mass(*args, **kwargs), dtype=torch.float64))

            if hasattr(mass, "_ps_sampler"):
                mass_preprocessed._ps_sampler = mass._ps_sampler
        else:
            mass_preprocessed = torch.atleast_1d(torch.as_tensor(np.asarray(mass, dtype=np.float64)))
        self._mass = mass_preprocessed
        self._generate_called = False
        self._compiled = None

    def __repr__(self):  # ps.py:121-126
        return "<phasespace.GenParticle: name='{}' mass={} children=[{}]>".format(
            self.name,
            f"{self._mass_val:.2f}" if self.has_fixed_mass else "variable",
            ", ".join(child.name for child in self.children),
        )

    def _do_names_clash(self, particles):  # ps.py:128-142
        def get_list_of_names(part):
            output = [part.name]
            for child in part.children:
                output.extend(get_list_of_names(child))
            return output

        names_to_check = [self.name]
        for part in particles:
            names_to_check.extend(get_list_of_names(part))
        dup_names = {name for name in names_to_check if names_to_check.count(name) > 1}
        return dup_names or None

    def get_mass(self, min_mass=None, max_mass=None, n_events=None, seed: SeedLike = None):
        """Particle mass (ps.py:144-184): the fixed mass, or the mass function evaluated on
        ``(min_mass, max_mass, n_events[, seed=rng])`` with shape ``(n_events,)`` limits."""
        if self.has_fixed_mass:
            return self._mass
        n_events = int(n_events)
        min_mass = torch.reshape(torch.as_tensor(min_mass, dtype=torch.float64), (n_events,))
        max_mass = torch.reshape(torch.as_tensor(max_mass, dtype=torch.float64), (n_events,))
        if "seed" in inspect.signature(self._mass).parameters:
            return self._mass(min_mass, max_mass, n_events, seed=get_rng(seed))
        return self._mass(min_mass, max_mass, n_events)

    @property
    def has_fixed_mass(self):
        """bool: is the mass a number rather than a callable? (ps.py:186-189)"""
        return not callable(self._mass)

    def set_children(self, *children):
        """Assign children (ps.py:191-221); returns ``self``.

        Raises:
            ValueError: children already set, or fewer than two children given.
            KeyError: particle name clash anywhere in the tree.
            RuntimeError: ``generate`` was already called.
        """
        if self._generate_called:
            raise RuntimeError("Cannot set children after the first call to `generate`.")
        if self.children:
            raise ValueError("Children already set!")
        if len(children) <= 1:
            raise ValueError(f"Have to set at least 2 children, not {len(children)} for a particle to decay")
        if name_clash := self._do_names_clash(children):
            raise KeyError(f"Particle name {name_clash} already used")
        self.children = children
        global _TREE_EPOCH
        _TREE_EPOCH += 1
        return self

    @property
    def has_children(self):
        """bool: does the particle have children? (ps.py:223-226)"""
        return bool(self.children)

    @property
    def has_grandchildren(self):
        """bool: does the particle have grandchildren? (ps.py:228-233)"""
        if not self.children:
            return False
        return any(child.has_children for child in self.children)

    # ------------------------------------------------------------------ compilation
    def _mark_generate_called(self):  # ps.py:318: the latch is set on every decaying particle
        self._generate_called = True
        for child in self.children:
            if child.children:
                child._mark_generate_called()

    def compile(self) -> CompiledDecay:
        """Compile (once) the decay tree below this particle into its kernel descriptor."""
        # a leaf can still get children after this tree was compiled (ps.py:207-210 latches decaying particles only)
        compiled = self._compiled
        if compiled is not None:
            if compiled.epoch == _TREE_EPOCH:
                return compiled
            if compiled.structure == _structure(self):
                compiled.epoch = _TREE_EPOCH
                return compiled

        def fused(part):
            return all((c.has_fixed_mass or hasattr(c._mass, "_ps_sampler")) and fused(c) for c in part.children)

        self._compiled = CompiledDecay(self, split=not fused(self))
        return self._compiled

    # ------------------------------------------------------------------ generation
    def generate(
        self,
        n_events,
        boost_to=None,
        normalize_weights: bool = True,
        seed: SeedLike = None,
        *,
        as_vectors: bool | None = None,
        exact: bool | None = None,
        debug_uniforms=None,
        _deferred_checks: list | None = None,
        _weight_scale: float = 1.0,
    ):
        """Generate normalized n-body phase space (reference: ps.py:628-717).

        Events are generated in the rest frame of the particle unless ``boost_to`` is given.

        Args:
            n_events: number of events.
            boost_to: optional ``(n_events, 4)`` or ``(1, 4)`` momenta ``[px, py, pz, E]`` the events
                are boosted to (tensor / array; a Python list is read transposed like the reference;
                a ``vector`` momentum object is accepted when that package is installed).
            normalize_weights: divide the weights by their analytic maximum.
            seed: ``None`` (global generator), an int, or a ``phasespace_b200.random.Generator``.
            as_vectors: return ``vector`` objects instead of tensors (needs the ``vector`` package).
            exact: (extension) run the variant that keeps the reference's operation order, op for op, including its
                re-derivation of every decaying particle's mass from its boosted 4-vector (ps.py:322); ``False``: the
                product kernel (same function of the draws, restructured arithmetic).  Default: ``True`` in debug
                mode (``debug_uniforms`` given), else ``False``.
            debug_uniforms: (extension) ``(n_events, n_draws)`` uniforms in the reference's draw order,
                used instead of the in-kernel RNG -- the parity/debug mode.

        Returns:
            ``(weights, particles)`` if ``normalize_weights`` else ``(weights, weights_max, particles)``;
            ``particles`` maps names to ``(n_events, 4)`` tensors in the reference's key order.

        Raises:
            ValueError: no children; resonance as top particle; ``boost_to`` size mismatch.
            ForbiddenDecayError: the decay is kinematically forbidden.
        """
        if not self.children:
            raise ValueError("No children have been configured")  # ps.py:319-320
        if boost_to is None and not self.has_fixed_mass:
            raise ValueError("Cannot use resonance as top particle")  # ps.py:543
        rng = get_rng(seed)
        device = _device()
        if isinstance(n_events, torch.Tensor):
            n_events = int(n_events.item())
        n_events = int(n_events)
        # The plain call -- events at rest, normalised weights, tensors out, a tree that cannot fail per event -- goes
        # straight to the C ABI: a 1e6-event launch lasts ~13 us on a B200, so every microsecond of host work shows.
        if (boost_to is None and normalize_weights and debug_uniforms is None and not as_vectors and not exact
                and _weight_scale == 1.0 and n_events > 0):
            compiled = self._compiled
            if compiled is not None and compiled.epoch == _TREE_EPOCH and compiled.plain_ok and self._generate_called:
                return compiled.generate_plain(n_events, rng.seed, rng.reserve(n_events), device)
        if boost_to is not None:
            boost_to = _convert_boost(boost_to, device)
            if boost_to.ndim == 1:
                boost_to = boost_to.reshape(1, -1)
            if boost_to.ndim != 2 or boost_to.shape[1] != 4:
                raise ValueError(f"Bad shape for momentum -> {list(boost_to.shape)}")  # ps.py:257-258
            if boost_to.shape[0] not in (n_events, 1):  # ps.py:697-703
                raise ValueError(
                    f"The number of events requested ({n_events}) doesn't match the boost_to input size "
                    f"of {tuple(boost_to.shape)}"
                )
            boost_to = boost_to.contiguous()
        if exact is None:
            exact = debug_uniforms is not None
        self._mark_generate_called()
        compiled = self.compile()
        if compiled._static_forbidden is not None and boost_to is None:
            raise ForbiddenDecayError(f"Forbidden decay (particle {compiled._static_forbidden})")
        user_masses = None
        if compiled.n_user:
            top_mass = float(self._mass_val) if boost_to is None else _mass_of(boost_to).expand(n_events)
            user_masses = compiled.presample(n_events, top_mass, device)
        if debug_uniforms is not None:
            debug_uniforms = torch.as_tensor(debug_uniforms, dtype=torch.float64, device=device).contiguous()
            if tuple(debug_uniforms.shape) != (n_events, compiled.n_draws):
                raise ValueError(f"debug_uniforms must have shape ({n_events}, {compiled.n_draws})")
        first_event = rng.reserve(n_events)
        weights, weights_max, slab, err = compiled.launch(
            n_events, rng.seed, first_event, boost_to=boost_to, normalize_weights=normalize_weights, exact=exact,
            debug_uniforms=debug_uniforms, user_masses=user_masses, _device_hint=device, weight_scale=_weight_scale,
        )
        if (compiled.has_resonances or boost_to is not None) and n_events > 0:
            if _deferred_checks is not None:  # GenMultiDecay checks all modes with one synchronisation
                _deferred_checks.append(err)
            elif int(err.item()):
                raise ForbiddenDecayError("Forbidden decay")  # ps.py:357-363
        parts = {name: slab[compiled.slots[name]] for name in compiled.names}
        if as_vectors:
            parts = to_vectors(parts)
        return (weights, parts) if normalize_weights else (weights, weights_max, parts)

    def generate_host(self, n_events, boost_to=None, normalize_weights: bool = True, seed: SeedLike = None, *,
                      chunk_events: int = 1 << 23, out: dict | None = None, exact: bool = False):
        """(extension) The same events as :meth:`generate`, delivered into HOST memory.

        The events are generated on the device in chunks of ``chunk_events`` and streamed into pinned
        host buffers while the next chunk is being generated (``ps_generate_host``).  This is the
        end-to-end call for users who want numpy arrays, which is what the reference returns.

        ``out`` may hold pinned CPU tensors ``weights`` ``(N,)``, ``particles`` ``(P, N, 4)`` and
        ``weights_max`` ``(N,)`` to reuse.  Returns numpy views: ``(weights, particles)`` or
        ``(weights, weights_max, particles)``.
        """
        if not self.children:
            raise ValueError("No children have been configured")
        if boost_to is None and not self.has_fixed_mass:
            raise ValueError("Cannot use resonance as top particle")
        _device()
        rng = get_rng(seed)
        n_events = int(n_events)
        self._mark_generate_called()
        compiled = self.compile()
        if compiled.n_user:
            raise ValueError("generate_host does not support arbitrary Python mass callables; use generate()")
        if compiled._static_forbidden is not None and boost_to is None:
            raise ForbiddenDecayError(f"Forbidden decay (particle {compiled._static_forbidden})")
        boost_rows = 0
        if boost_to is not None:
            boost_to = _convert_boost(boost_to, "cpu")
            boost_to = boost_to.reshape(1, -1) if boost_to.ndim == 1 else boost_to
            if boost_to.ndim != 2 or boost_to.shape[1] != 4:
                raise ValueError(f"Bad shape for momentum -> {list(boost_to.shape)}")  # ps.py:257-258
            if boost_to.shape[0] not in (n_events, 1):
                raise ValueError(
                    f"The number of events requested ({n_events}) doesn't match the boost_to input size "
                    f"of {tuple(boost_to.shape)}"
                )
            boost_to = boost_to.contiguous()
            boost_rows = boost_to.shape[0]
        out = out or {}

        def pinned(key, shape):
            t = out.get(key)
            if t is None:
                t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            _check_out(t, key, shape, None, pinned_host=True)
            return t

        weights = pinned("weights", (n_events,))
        particles = pinned("particles", (compiled.n_particles, n_events, 4))
        weights_max = None if normalize_weights else pinned("weights_max", (n_events,))
        flags = (_lib.PS_GEN_NORMALIZE if normalize_weights else 0) | (_lib.PS_GEN_EXACT if exact else 0)
        first_event = rng.reserve(n_events)
        _lib.check(_lib.lib.ps_generate_host(
            compiled.handle, rng.seed, first_event, n_events, _ptr(boost_to), boost_rows, _ptr(weights),
            _ptr(weights_max), _ptr(particles), int(chunk_events), flags))
        parts = {name: particles[compiled.slots[name]].numpy() for name in compiled.names}
        if normalize_weights:
            return weights.numpy(), parts
        return weights.numpy(), weights_max.numpy(), parts

    def generate_tensor(self, n_events: int, boost_to=None, normalize_weights: bool = True):
        """Removed in the reference as well (ps.py:719-761)."""
        raise RuntimeError("This function is removed. Use `generate` which does not return a Tensor as well.")


def _mass_of(vec4: torch.Tensor) -> torch.Tensor:
    """kin.mass of ``(n, 4)`` rows -> ``(n,)`` (kinematics.py:93-104)."""
    sq = vec4 * vec4
    return torch.sqrt(((-sq[:, 0] - sq[:, 1]) - sq[:, 2]) + sq[:, 3])


def _convert_boost(boost_to, device):
    """ps.py:676-695: accept ``vector`` momentum objects when the package exists, else array-likes."""
    try:
        import vector
    except ImportError:
        vector = None
    if vector is not None and isinstance(boost_to, vector.Vector):
        if not (isinstance(boost_to, vector.Momentum) and isinstance(boost_to, vector.Lorentz)):
            raise ValueError("boost_to has to be a momentum Lorentz vector.")
        try:
            boost_to = boost_to.to_pxpypzenergy()
        except Exception as error:
            raise ValueError(
                "boost_to has to be a momentum Lorentz vector, failed to convert to pxpypzenergy."
            ) from error
        boost_to = np.stack([boost_to.px, boost_to.py, boost_to.pz, boost_to.energy], axis=-1)
    return process_list_to_tensor(boost_to, device=device)


class Particle:
    """Deprecated name (ps.py:765-775)."""

    def __init__(self):
        raise NameError(
            "'Particle' has been renamed to 'GenParticle'. Please update your code accordingly."
            "For more information, see: https://github.com/zfit/phasespace/issues/22"
        )


def nbody_decay(mass_top: float, masses: list, top_name: str = "", names: list = None):
    """Shortcut to build an n-body decay of a GenParticle (ps.py:778-806).

    Without names the top particle is called ``top`` and the children ``p_{i}``.

    Raises:
        ValueError: if the lengths of ``masses`` and ``names`` differ.
    """
    if not top_name:
        top_name = "top"
    if not names:
        names = [f"p_{num}" for num in range(len(masses))]
    if len(names) != len(masses):
        raise ValueError("Mismatch in length between children masses and their names.")
    return GenParticle(top_name, mass=mass_top).set_children(
        *(GenParticle(names[num], mass=mass) for num, mass in enumerate(masses))
    )


def generate_decay(*args, **kwargs):
    """Deprecated (ps.py:809-814)."""
    raise NameError(
        "'generate_decay' has been removed. A similar behavior can be accomplished with 'nbody_decay'. "
        "For more information see https://github.com/zfit/phasespace/issues/22"
    )


def to_vectors(particles: dict) -> dict:
    """Convert a dict of ``(N, 4)`` tensors to ``vector`` momentum arrays (ps.py:817-835)."""
    try:
        import vector
    except ImportError as error:
        _raise_missing_vector_package(error)
    newparticles = {}
    for name, particle in particles.items():
        arr = particle.detach().cpu().numpy() if isinstance(particle, torch.Tensor) else np.asarray(particle)
        px, py, pz, e = np.moveaxis(arr, -1, 0)
        newparticles[name] = vector.array(dict(px=px, py=py, pz=pz, energy=e))
    return newparticles


def _raise_missing_vector_package(exception: ImportError) -> NoReturn:
    raise ImportError("To use `boost_to`, the `vector` package has to be installed.") from exception
