"""``GenMultiDecay`` (mirror of the reference's fromdecay/genmultidecay.py:14-340)."""

from __future__ import annotations

import itertools
from collections.abc import Callable

import numpy as np
import torch

from ..phasespace import GenParticle, prepare_all
from ..random import Generator, get_rng
from . import pdg
from .mass_functions import DEFAULT_CONVERTER


class GenMultiDecay:
    MASS_WIDTH_TOLERANCE = 0.01
    DEFAULT_MASS_FUNC = "relbw"
    MAX_STREAMS = 8  # CUDA streams the per-mode kernels are spread over
    MIN_EVENTS_FOR_STREAMS = 1 << 23  # calls with fewer events stay on the current stream

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

    def mode_counts(self, n_events: int, rng: Generator):
        """Events per decay mode: one multinomial draw with the mode probabilities (genmultidecay.py:165-183), from the
        generator's seed and one reserved counter -- a pure function of (seed, offset), so every rank of a sharded job
        computes the same split on its host."""
        probs = np.asarray([float(dm[0]) for dm in self.gen_particles], dtype=np.float64)
        probs = probs / probs.sum()
        split_rng = np.random.Generator(np.random.Philox(key=rng.seed ^ 0x5DEECE66D, counter=rng.reserve(1)))
        return split_rng.multinomial(int(n_events), probs)

    def generate_sharded(self, n_events: int, normalize_weights: bool = True, *, rank: int | None = None,
                         world: int | None = None, group=None, **kwargs):
        """This rank's share of ``generate(n_events, ...)``: the mode counts come from the seed and
        are the same on every rank, and rank r generates the contiguous slice ``shard_range(n_mode, r, world)`` of every
        populated mode's events, with the event indices the single-GPU call would have used.  The concatenation over
        ranks of each mode's tensors is therefore bit-identical to the single-GPU result.  No collective on the data
        path; the only exchange is one MAX all-reduce of a scalar when the global maximum weight is not known beforehand
        (modes with ``boost_to`` or resonant leaves).  ``rank`` / ``world`` default to the process group's.  Returns what
        ``generate`` returns, each list entry holding this rank's events of that mode (possibly zero of them)."""
        from ..sharding import _rank_world

        if rank is None or world is None:
            rank, world = _rank_world(group)
        return self.generate(n_events, normalize_weights, _shard=(int(rank), int(world), group), **kwargs)

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
        shard = kwargs.pop("_shard", None)
        counts = self.mode_counts(n_events, rng)
        weights, max_weights, events = [], [], []
        # one kernel per populated mode, fanned out over a few CUDA streams; forbidden-decay flags are
        # collected and checked with a single synchronisation at the end
        main = torch.cuda.current_stream()
        populated = [(gen, int(n)) for (_, gen), n in zip(self.gen_particles, counts) if n > 0]
        # kernels of modes seen for the first time are compiled side by side, not one inside each generate() below
        prepare_all([gen.compile() for gen, _ in populated], boost=kwargs.get("boost_to") is not None,
                    exact=bool(kwargs.get("exact")))
        # Side streams pay when the modes' kernels are long enough to overlap their tails; a call of a few million events
        # over a dozen modes is a row of small launches, cheapest issued back to back on the current stream (where each
        # also overlaps the drain of the one before: programmatic dependent launch).
        n_streams = min(len(populated), self.MAX_STREAMS) if n_events >= self.MIN_EVENTS_FOR_STREAMS else 1
        streams = self._stream_pool(n_streams) if n_streams > 1 else []
        # nothing but n_events, normalize_weights and the seed: every mode that can goes through generate_fast, and all of
        # them share one sticky forbidden-decay flag (zeroed on `main` BEFORE the side streams start waiting on it)
        fast = not kwargs and len(populated) > 0
        device = torch.device("cuda", torch.cuda.current_device())
        checks = []
        shared_err = None
        if fast:
            shared_err = torch.zeros((1,), dtype=torch.int32, device=device)
            checks.append(shared_err)
        for st in streams:
            st.wait_stream(main)
        # genmultidecay.py:192-195 divides every mode's weights by the largest maximum weight of all modes.  When every
        # populated mode has an event-independent maximum weight (particles at rest with fixed-mass leaves: what a
        # DecayLanguage chain gives) that number is known before anything is generated, and each kernel normalises by
        # it directly (w / w_max_mode * (w_max_mode / w_max_all)) -- no weights_max arrays, no max and divide passes.
        mode_w_max = None
        if normalize_weights and populated and kwargs.get("boost_to") is None and not kwargs.get("exact"):
            mode_w_max = [gen.compile().w_max for gen, _ in populated]
            if not all(np.isfinite(w) and w > 0 for w in mode_w_max):
                mode_w_max = None
        total_w_max = max(mode_w_max) if mode_w_max else None
        if shard is not None:
            from ..sharding import shard_range

            boost = kwargs.get("boost_to")
            if boost is not None and np.shape(boost)[0] != 1 and len(populated) > 0:
                raise ValueError("sharded GenMultiDecay takes one boost_to row for all events, not one per event")
            if any(gen.compile().n_user for gen, _ in populated):
                raise ValueError("sharded generation needs in-kernel mass functions (no arbitrary Python callables)")
        for k, (gen, n) in enumerate(populated):
            if mode_w_max:
                call = dict(normalize_weights=True, _weight_scale=mode_w_max[k] / total_w_max)
            else:
                call = dict(normalize_weights=False)
            mode_rng, n_here = rng, n
            if shard is not None:  # this rank's slice of the mode's events, at the indices the unsharded call uses
                lo, n_here = shard_range(n, shard[0], shard[1])
                mode_rng = Generator(rng.seed, rng.reserve(n) + lo)
            compiled = gen.compile() if fast else None
            if fast and compiled.fast_ok and n_here > 0:
                # one allocation, one C-ABI call, strided views; all modes share one sticky forbidden-decay flag
                gen._mark_generate_called()
                first = mode_rng.reserve(n_here)
                if streams:
                    with torch.cuda.stream(streams[k % n_streams]):
                        res = compiled.generate_fast(n_here, mode_rng.seed, first, device, shared_err,
                                                     call["normalize_weights"], call.get("_weight_scale", 1.0))
                    for t in (res[0], next(iter(res[2].values()))):
                        t.record_stream(main)
                else:
                    res = compiled.generate_fast(n_here, mode_rng.seed, first, device, shared_err,
                                                 call["normalize_weights"], call.get("_weight_scale", 1.0))
                out = (res[0], res[2]) if call["normalize_weights"] else res
            elif streams:
                with torch.cuda.stream(streams[k % n_streams]):
                    out = gen.generate(n_here, seed=mode_rng, _deferred_checks=checks, **call, **kwargs)
                for t in (out[0], next(iter(out[-1].values()))):
                    t.record_stream(main)  # the slab was allocated on the side stream and is consumed on `main`
            else:
                out = gen.generate(n_here, seed=mode_rng, _deferred_checks=checks, **call, **kwargs)
            weights.append(out[0])
            max_weights.append(None if mode_w_max else out[1])
            events.append(out[-1])
        for st in streams:
            main.wait_stream(st)
        if checks and bool(torch.stack(checks).any().item()):
            from ..phasespace import ForbiddenDecayError

            raise ForbiddenDecayError("Forbidden decay")
        if normalize_weights:
            if mode_w_max:
                return weights, events
            local = [mw.max() for mw in max_weights if mw.numel()]  # genmultidecay.py:193
            total_max = torch.stack(local).max() if local else torch.zeros((), dtype=torch.float64, device=weights[0].device if weights else "cuda")
            if shard is not None and shard[1] > 1:
                from ..sharding import allreduce_max

                total_max = allreduce_max(total_max, shard[2], world=shard[1])
            return [w / total_max for w in weights], events
        return weights, max_weights, events


class _Namer:
    """Hands out particle names that are unique inside one ``from_dict`` call: the first use of ``"pi+"`` keeps
    its name, later ones become ``"pi+ [0]"``, ``"pi+ [1]"``, ... (genmultidecay.py:200-221)."""

    def __init__(self):
        self.taken: set[str] = set()

    def __call__(self, base: str) -> str:
        name, k = base, 0
        while name in self.taken:
            name = f"{base} [{k}]"
            k += 1
        self.taken.add(name)
        return name


def _unique_name(name: str, preexisting_particles: set[str]) -> str:
    """Functional form of :class:`_Namer` over a caller-owned set (same contract as the reference helper)."""
    namer = _Namer()
    namer.taken = preexisting_particles
    return namer(name)


def _resonance_mass(evtgen_name: str, model: str, converters: dict[str, Callable], tolerance: float):
    """Mass of a decaying particle: the PDG mass if its width is at most ``tolerance``, otherwise the mass
    function ``converters[model](mass, width)`` (genmultidecay.py:224-249)."""
    entry = pdg.lookup(evtgen_name)
    if entry.width <= tolerance:
        return float(entry.mass)
    return converters[model](mass=entry.mass, width=entry.width)


def _model_for(mode: dict, evtgen_name: str, particle_model_map: dict[str, str]) -> str:
    """Precedence of mass-function names: the mode's own ``zfit`` key, then ``particle_model_map``, then the
    class default (genmultidecay.py:321-326)."""
    if "zfit" in mode:
        return mode["zfit"]
    return particle_model_map.get(evtgen_name, GenMultiDecay.DEFAULT_MASS_FUNC)


def _expand(chain: dict, converters: dict[str, Callable], particle_model_map: dict[str, str], tolerance: float,
            namer: _Namer, is_top: bool) -> list[tuple[float, GenParticle]]:
    """Every way ``chain`` (one mother with its list of decay modes) can decay, as ``(probability, GenParticle)``.

    A final-state entry is either a stable particle name or a nested chain with alternatives of its own; the
    alternatives of all daughters are combined by Cartesian product, probabilities multiplied
    (genmultidecay.py:252-340)."""
    (mother,) = chain.keys()
    mother_name = namer(mother)
    out = []
    for mode in chain[mother]:
        alternatives = []  # per daughter: [(probability, template GenParticle), ...]
        for daughter in mode["fs"]:
            if isinstance(daughter, str):
                alternatives.append([(1.0, GenParticle(namer(daughter), float(pdg.lookup(daughter).mass)))])
            elif isinstance(daughter, dict):
                alternatives.append(_expand(daughter, converters, particle_model_map, tolerance, namer, is_top=False))
            else:
                raise TypeError(
                    'Expected elements in decaychain["fs"] to only be str or dict '
                    f"but found an instance of type {type(daughter)}"
                )
        for combo in itertools.product(*alternatives):
            prob = float(mode["bf"])
            for daughter_prob, _ in combo:
                prob *= float(daughter_prob)
            if is_top:
                mass = float(pdg.lookup(mother).mass)
            else:
                mass = _resonance_mass(mother, _model_for(mode, mother, particle_model_map), converters, tolerance)
            decay = GenParticle(mother_name, mass).set_children(*(_clone(template) for _, template in combo))
            out.append((prob, decay))
    return out


def _recursively_traverse(decaychain: dict, mass_converter: dict[str, Callable], particle_model_map: dict[str, str],
                          tolerance: float, preexisting_particles: set[str] = None) -> list[tuple[float, GenParticle]]:
    """Entry point kept under the reference's name."""
    if tolerance is None:
        tolerance = GenMultiDecay.MASS_WIDTH_TOLERANCE
    namer = _Namer()
    is_top = preexisting_particles is None
    if not is_top:
        namer.taken = preexisting_particles
    return _expand(decaychain, mass_converter, particle_model_map, tolerance, namer, is_top)


def _clone(part: GenParticle) -> GenParticle:
    """Templates are shared between the combinations of ``itertools.product``; every decay tree gets its own
    particle objects so that compiled kernels and ``generate`` latches are per tree."""
    new = GenParticle(part.name, part._mass_val)
    if part.children:
        new.children = tuple(_clone(c) for c in part.children)
    return new
