"""CPU timing of the oracle for bench.py's cpu_baseline / --impl reference legs.

TEST INFRASTRUCTURE: never imported by the product path.

The reference's own CPU path (TensorFlow, benchmark/bench_phasespace.py:109-133: B -> pi pi pi,
1e6 events per call) cannot run here (TensorFlow is absent); its stand-in is the numpy restatement
(oracle/genbod_numpy.py), which has the same op-by-op array-materialising structure as the TF graph.
To use all host cores, like TF's intra-op thread pool would, the sample is split over one worker
process per core.
"""

from __future__ import annotations

import concurrent.futures
import multiprocessing
import os
import time

import numpy as np


def _worker(args):
    m_top, masses, n_events, seed, reps = args
    from oracle import genbod_numpy as orc
    from oracle.philox import uniform_table

    tree = orc.flat(m_top, masses)
    nd = orc.n_draws(tree)
    done = 0
    for r in range(reps):
        # the draws are part of the reference's work (rng.uniform inside _generate): time them too
        draws = np.random.default_rng(seed + r).random((n_events, nd))
        w, parts = orc.generate(tree, n_events, draws)
        done += len(w)
    return done


class NumpyOracleBench:
    """Pool of worker processes, created once (spawn: safe next to an initialised CUDA context)."""

    def __init__(self, m_top, masses, cores=None):
        self.m_top, self.masses = float(m_top), [float(m) for m in masses]
        self.cores = cores or os.cpu_count() or 1
        ctx = multiprocessing.get_context("spawn")
        self.pool = concurrent.futures.ProcessPoolExecutor(self.cores, mp_context=ctx)
        list(self.pool.map(_worker, [(self.m_top, self.masses, 1000, 0, 1)] * self.cores))  # warm the workers

    def step(self, events_per_worker=200_000, reps=1, seed=1):
        """One timed pass: every worker generates ``reps`` x ``events_per_worker`` events. Returns (events, seconds)."""
        t0 = time.perf_counter()
        done = sum(self.pool.map(
            _worker, [(self.m_top, self.masses, events_per_worker, seed + 1000 * i, reps) for i in range(self.cores)]
        ))
        return done, time.perf_counter() - t0

    def close(self):
        self.pool.shutdown()


def c_genbod_rate(m_top, masses, n_events=4_000_000, threads=0):
    """events/s of the compiled scalar GENBOD (oracle/genbod_ref.c) on ``threads`` (0 = all) cores."""
    from oracle import genbod_ref

    genbod_ref.generate(m_top, masses, 100_000, seed=1, n_threads=threads, store=False)
    t0 = time.perf_counter()
    _, _, used = genbod_ref.generate(m_top, masses, n_events, seed=2, n_threads=threads, store=False)
    return n_events / (time.perf_counter() - t0), used
