"""``GenMultiDecay`` (mirror of the reference's fromdecay/genmultidecay.py:14-340)."""

from __future__ import annotations

import itertools
from collections.abc import Callable

import numpy as np
import torch

from ..phasespace import GenParticle
from ..random import Generator, get_rng
from . import pdg
from .mass_functions import DEFAULT_CONVERTER


class GenMultiDecay:
    MASS_WIDTH_TOLERANCE = 0.01
    DEFAULT_MASS_FUNC = "relbw"
    MAX_STREAMS = 8  # CUDA streams the per-mode kernels are spread over

    def __init__(self, gen_particles: list[tuple[float, GenParticle]]):
        """A ``GenParticle``-type container that can handle multiple decays (genmultidecay.py:18-26).

        Args:
            gen_particles: ``[(probability, GenParticle), ...]``.
        """
        self.gen_particles = gen_particles
        self._streams = {}

    def _stream_pool(self, n):
        """Side streams are kept per device: the caching allocator reuses blocks per stream, so fresh streams
        on every call would turn every output allocation into a cudaMalloc."""
        pool = self._streams.setdefault(torch.cuda.current_device(), [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream())
        return pool[:n]

    @classmethod
    def from_dict(cls, dec_dict: dict, mass_converter: dict[str, Callable] = None, tolerance: float = None,
                  particle_model_map: dict[str, str] = None):
        """Create a ``GenMultiDecay`` from a dict in the ``DecayLanguage`` format (genmultidecay.py:28-153).

        Args:
            dec_dict: ``{mother: [{"bf": float, "fs": [name | dict, ...], "zfit": mass-function name}, ...]}``.
            mass_converter: extra ``{name: factory(mass, width) -> mass function}`` entries, merged over the
                built-in ``gauss`` / ``bw`` / ``relbw``.
            tolerance: minimum width for a particle to get a mass function instead of a fixed mass
                (default: the class variable ``MASS_WIDTH_TOLERANCE``).
            particle_model_map: ``{particle name: mass-function name}`` applied to every appearance of
                that particle; a ``zfit`` key in the decay dict takes precedence.
        """
        if tolerance is None:
            tolerance = cls.MASS_WIDTH_TOLERANCE
        if particle_model_map is None:
            particle_model_map = {}
        total_mass_converter = DEFAULT_CONVERTER if mass_converter is None else {**DEFAULT_CONVERTER, **mass_converter}
        gen_particles = _recursively_traverse(
            dec_dict, total_mass_converter, particle_model_map=particle_model_map, tolerance=tolerance
        )
        return cls(gen_particles)

    def generate(self, n_events: int, normalize_weights: bool = True, **kwargs):
        """Generate four-momenta from the decay modes (genmultidecay.py:155-197).

        ``n_events`` is split over the modes by one multinomial draw with the mode probabilities; each
        populated mode generates its share with ``GenParticle.generate(n, normalize_weights=False)``.
        Returns lists with one entry per populated mode: ``(weights, events)`` when
        ``normalize_weights`` (weights divided by the maximum of all maximum weights), else
        ``(weights, max_weights, events)``.

        Differences from the reference, on purpose: the split is drawn from the ``seed`` keyword (the
        reference uses TensorFlow's unseeded global generator even when ``seed`` is given) and the
        modes are returned in declaration order (the reference: order of first appearance in the draw).
        """
        n_events = int(n_events)
        rng = get_rng(kwargs.pop("seed", None))
        probs = np.asarray([float(dm[0]) for dm in self.gen_particles], dtype=np.float64)
        probs = probs / probs.sum()
        split_rng = np.random.Generator(np.random.Philox(key=rng.seed ^ 0x5DEECE66D, counter=rng.reserve(1)))
        counts = split_rng.multinomial(n_events, probs)
        weights, max_weights, events = [], [], []
        # one kernel per populated mode, fanned out over a few CUDA streams; forbidden-decay flags are
        # collected and checked with a single synchronisation at the end
        main = torch.cuda.current_stream()
        populated = [(gen, int(n)) for (_, gen), n in zip(self.gen_particles, counts) if n > 0]
        n_streams = min(len(populated), self.MAX_STREAMS)
        streams = self._stream_pool(n_streams) if n_streams > 1 else []
        for st in streams:
            st.wait_stream(main)
        checks = []
        for k, (gen, n) in enumerate(populated):
            if streams:
                with torch.cuda.stream(streams[k % n_streams]):
                    out = gen.generate(n, normalize_weights=False, seed=rng, _deferred_checks=checks, **kwargs)
                for t in (out[0], out[1], next(iter(out[2].values()))):
                    t.record_stream(main)  # the slab was allocated on the side stream and is consumed on `main`
            else:
                out = gen.generate(n, normalize_weights=False, seed=rng, _deferred_checks=checks, **kwargs)
            weights.append(out[0])
            max_weights.append(out[1])
            events.append(out[2])
        for st in streams:
            main.wait_stream(st)
        if checks and bool(torch.stack(checks).any().item()):
            from ..phasespace import ForbiddenDecayError

            raise ForbiddenDecayError("Forbidden decay")
        if normalize_weights:
            total_max = torch.stack([mw.max() for mw in max_weights]).max()  # genmultidecay.py:193
            return [w / total_max for w in weights], events
        return weights, max_weights, events


def _unique_name(name: str, preexisting_particles: set[str]) -> str:
    """A name not in ``preexisting_particles``: ``name`` or ``"name [i]"`` (genmultidecay.py:200-221)."""
    if name not in preexisting_particles:
        preexisting_particles.add(name)
        return name
    name += " [0]"
    i = 1
    while name in preexisting_particles:
        name = name[: name.rfind("[")] + f"[{str(i)}]"
        i += 1
    preexisting_particles.add(name)
    return name


def _get_particle_mass(name: str, mass_converter: dict[str, Callable], mass_func: str, tolerance: float):
    """Fixed mass if the width is below ``tolerance``, else a mass function (genmultidecay.py:224-249)."""
    particle = pdg.lookup(name)
    if particle.width <= tolerance:
        return float(particle.mass)
    return mass_converter[mass_func](mass=particle.mass, width=particle.width)


def _recursively_traverse(decaychain: dict, mass_converter: dict[str, Callable], particle_model_map: dict[str, str],
                          tolerance: float, preexisting_particles: set[str] = None) -> list[tuple[float, GenParticle]]:
    """All possible GenParticles of a DecayLanguage dict, one per combination of sub-modes, with the
    product of their branching fractions (genmultidecay.py:252-340)."""
    if tolerance is None:
        tolerance = GenMultiDecay.MASS_WIDTH_TOLERANCE
    (original_mother_name,) = decaychain.keys()
    if preexisting_particles is None:
        preexisting_particles = set()
        is_top_particle = True
    else:
        is_top_particle = False
    decay_modes = decaychain[original_mother_name]
    mother_name = _unique_name(original_mother_name, preexisting_particles)
    all_decays = []
    for dm in decay_modes:
        dm_probability = dm["bf"]
        daughter_gens = []
        for daughter_name in dm["fs"]:
            if isinstance(daughter_name, str):  # stable particles always get a constant mass
                daughter = GenParticle(
                    _unique_name(daughter_name, preexisting_particles), float(pdg.lookup(daughter_name).mass)
                )
                daughter = [(1.0, daughter)]
            elif isinstance(daughter_name, dict):
                daughter = _recursively_traverse(
                    daughter_name, mass_converter, particle_model_map, tolerance, preexisting_particles
                )
            else:
                raise TypeError(
                    f'Expected elements in decaychain["fs"] to only be str or dict '
                    f"but found an instance of type {type(daughter_name)}"
                )
            daughter_gens.append(daughter)
        for daughter_combination in itertools.product(*daughter_gens):
            p = float(np.prod([decay[0] for decay in daughter_combination])) * dm_probability
            if is_top_particle:
                mother_mass = float(pdg.lookup(original_mother_name).mass)
            else:
                if "zfit" in dm:
                    mass_func = dm["zfit"]
                elif original_mother_name in particle_model_map:
                    mass_func = particle_model_map[original_mother_name]
                else:
                    mass_func = GenMultiDecay.DEFAULT_MASS_FUNC
                mother_mass = _get_particle_mass(
                    original_mother_name, mass_converter=mass_converter, mass_func=mass_func, tolerance=tolerance
                )
            one_decay = _clone(GenParticle(mother_name, mother_mass)).set_children(
                *(_clone(decay[1]) for decay in daughter_combination)
            )
            all_decays.append((p, one_decay))
    return all_decays


def _clone(part: GenParticle) -> GenParticle:
    """Daughter objects are shared between the combinations of ``itertools.product``; every decay tree
    gets its own copies so that compiled kernels and ``generate`` latches are per tree."""
    new = GenParticle(part.name, part._mass_val)
    if part.children:
        new.children = tuple(_clone(c) for c in part.children)
    return new
