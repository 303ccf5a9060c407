"""Philox4x32-10 counter RNG -- CPU restatement used ONLY as test infrastructure.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product path never does.

The reference draws its uniforms from ``tf.random.Generator`` (random.py:14-40, call sites
phasespace.py:367,444-446,451).  TensorFlow is a third-party dependency that is absent here, so
the *stream* of the reference is unpinned ("parity unpinned" for RNG values; only determinism is
tested upstream, tests/test_generate.py:66-88).  The B200 kernel replaces it by the published
Philox4x32-10 algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)
keyed on (seed, event index); this file restates that algorithm in numpy so that the in-kernel
stream can be checked bit-for-bit, and is pinned against the Random123 known-answer vectors in
tests/test_oracle_philox.py.

Stream layout shared with the kernel (phasespace_b200/csrc/ps_philox.cuh):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (event & 0xffffffff, event >> 32, block, stream)
    the 128-bit outputs of blocks 0, 1, 2, ... of one event form a bit string (word j of block b holds bits
    128 b + 32 j ... + 31, least significant first); uniform(column) = the 51-bit field at bit offset
    51 * column, times 2**-51.  (Dense packing: the three-body decay's 5 uniforms need two blocks, not three.)

``column`` is the index of the draw in the tree-wide draw order, so a
``(n_events, n_draws)`` table built here is exactly what the kernel consumes internally.
"""

from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint64(0x9E3779B9)
PHILOX_W1 = np.uint64(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)


def philox4x32_10(counter, key):
    """Vectorised Philox4x32-10.

    Args:
        counter: sequence of 4 uint32-valued arrays (broadcastable).
        key: sequence of 2 uint32-valued arrays (broadcastable).

    Returns:
        list of 4 ``np.uint64`` arrays holding 32-bit outputs.
    """
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in counter)
    k0, k1 = (np.asarray(k, dtype=np.uint64) & _MASK32 for k in key)
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    for rnd in range(10):
        if rnd:
            k0 = (k0 + PHILOX_W0) & _MASK32
            k1 = (k1 + PHILOX_W1) & _MASK32
        p0 = PHILOX_M0 * c0  # 32x32 -> 64 bit, no overflow in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> _SHIFT32, p0 & _MASK32
        hi1, lo1 = p1 >> _SHIFT32, p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return [c0, c1, c2, c3]


UNIFORM_BITS = 51


def uniform_table(seed: int, first_event: int, n_events: int, n_columns: int, stream: int = 0):
    """The ``(n_events, n_columns)`` fp64 table of U[0,1) draws the kernel uses for these events."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    events = np.arange(first_event, first_event + n_events, dtype=np.uint64)
    out = np.empty((n_events, n_columns), dtype=np.float64)
    n_blocks = (UNIFORM_BITS * n_columns + 127) // 128
    words = []  # 32-bit words of the per-event bit string, least significant first
    for blk in range(n_blocks):
        words.extend(philox4x32_10(
            (events & _MASK32, events >> _SHIFT32, np.uint64(blk), np.uint64(stream)),
            (np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)),
        ))
    words.extend([np.zeros(n_events, dtype=np.uint64)] * 2)
    mask = np.uint64((1 << UNIFORM_BITS) - 1)
    for col in range(n_columns):
        i, sh = divmod(UNIFORM_BITS * col, 32)
        lo = words[i] | (words[i + 1] << _SHIFT32)  # bits 32 i ... 32 i + 63
        field = lo >> np.uint64(sh)
        if sh + UNIFORM_BITS > 64:
            field = field | (words[i + 2] << np.uint64(64 - sh))
        out[:, col] = (field & mask).astype(np.float64) * 2.0**-UNIFORM_BITS
    return out
