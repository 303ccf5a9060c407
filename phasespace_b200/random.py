"""Random number handling (mirror of the reference's ``phasespace/random.py:14-40``).

The reference hands a ``tf.random.Generator`` down to ``rng.uniform``.  Here the draws are made
inside the kernel by Philox4x32-10 keyed on ``(seed, global event index)``
(csrc/ps_philox.cuh), so a generator is just the pair *(seed, next event index)*:

* ``get_rng(None)``      -> the process-global generator (advances between calls),
* ``get_rng(int)``       -> a fresh generator: same seed => same events,
* ``get_rng(Generator)`` -> that generator, used and advanced.

Because an event depends only on ``(seed, event index)``, any contiguous range of events can be
regenerated on any GPU: that is what makes multi-GPU sharding bit-identical and is the
"checkpoint/resume" of this generator.
"""

from __future__ import annotations

import os
import threading

import torch

_MASK64 = (1 << 64) - 1


class Generator:
    """Counter-based generator state: ``seed`` and the index of the next unused event."""

    __slots__ = ("seed", "offset")
    _lock = threading.Lock()  # one lock for all generators: reserve() holds it for two additions

    def __init__(self, seed: int, offset: int = 0):
        self.seed = int(seed) & _MASK64
        self.offset = int(offset)

    @classmethod
    def from_seed(cls, seed: int) -> "Generator":
        return cls(seed)

    @classmethod
    def from_non_deterministic_state(cls) -> "Generator":
        return cls(int.from_bytes(os.urandom(8), "little"))

    def reserve(self, n_events: int) -> int:
        """Claim ``n_events`` event indices; returns the first one."""
        with self._lock:
            first = self.offset
            self.offset += int(n_events)
        return first

    def split(self, count: int = 1):
        """New independent generators; advances this one (``tf.random.Generator.split``)."""
        first = self.reserve(count)
        return [Generator(_mix(self.seed, first + i + 1)) for i in range(count)]

    # Host-visible draws, for API parity with tests/test_random.py of the reference.  Not used by the
    # event generator (its draws are made in the kernel).
    def uniform_full_int(self, shape, dtype=torch.int64, device=None):
        n = int(torch.Size(shape).numel()) if not isinstance(shape, int) else int(shape)
        first = self.reserve(n)
        words = _philox_words(self.seed, first, n, stream=3, device=device)
        out = (words[1] << 32) | words[0]
        return out.reshape(shape).to(dtype)

    def uniform(self, shape, minval=0.0, maxval=1.0, dtype=torch.float64, device=None):
        n = int(torch.Size(shape).numel()) if not isinstance(shape, int) else int(shape)
        first = self.reserve(n)
        words = _philox_words(self.seed, first, n, stream=3, device=device)
        bits = ((words[1] << 32) | words[0]) >> 12 & ((1 << 52) - 1)
        u = bits.to(torch.float64) * 2.0**-52
        return (minval + (maxval - minval) * u).reshape(shape).to(dtype)

    def __repr__(self):
        return f"<phasespace_b200.random.Generator seed={self.seed} offset={self.offset}>"


def _mix(seed: int, k: int) -> int:
    z = (seed + k * 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def _philox_words(seed, first, n, stream, device=None):
    """Philox4x32-10 outputs (4 int64 tensors holding 32-bit words) for counters first..first+n-1."""
    m32 = 0xFFFFFFFF
    ev = torch.arange(n, dtype=torch.int64, device=device)
    lo = (ev + (first & m32)) & m32
    carry = (ev + (first & m32)) >> 32
    hi = (carry + ((first >> 32) & m32)) & m32
    c = [lo, hi, torch.zeros_like(ev), torch.full_like(ev, stream)]
    k0, k1 = seed & m32, (seed >> 32) & m32

    def mulhilo(a, b):  # 32x32 -> (hi, lo) without overflowing int64: split b in 16-bit halves
        b_lo, b_hi = b & 0xFFFF, b >> 16
        p_lo, p_hi = a * b_lo, a * b_hi  # each < 2^48
        full_lo = (p_lo + ((p_hi & 0xFFFF) << 16))
        low = full_lo & m32
        high = ((p_hi >> 16) + (full_lo >> 32)) & m32
        return high, low

    for r in range(10):
        if r:
            k0 = (k0 + 0x9E3779B9) & m32
            k1 = (k1 + 0xBB67AE85) & m32
        hi0, lo0 = mulhilo(0xD2511F53, c[0])
        hi1, lo1 = mulhilo(0xCD9E8D57, c[2])
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
    return c


SeedLike = int | Generator | None

_GLOBAL = Generator.from_non_deterministic_state()


def get_global_generator() -> Generator:
    return _GLOBAL


def set_global_seed(seed: int) -> None:
    global _GLOBAL
    _GLOBAL = Generator(seed)


def get_rng(seed: SeedLike = None) -> Generator:
    """Get or create a generator (reference: random.py:14-40)."""
    if seed is None:
        return _GLOBAL
    if type(seed) is int:  # the common case of a seeded call, first: a small call's host time shows (bench config 1)
        return Generator(seed)
    if isinstance(seed, Generator):
        return seed
    if isinstance(seed, torch.Tensor):
        seed = int(seed.item())
    return Generator.from_seed(int(seed))
