#!/usr/bin/env python3
"""The reference's own benchmark protocol (benchmark/bench_phasespace.py:95-133): B -> pi pi pi with B_MASS 5279.0,
PION_MASS 139.6, 1e6 events per call, one warm-up call then 10 timed calls, ms per call -- through the public API."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as phasespace

B_MASS, PION_MASS, N_EVENTS, n_runs = 5279.0, 139.6, 1000000, 10
decay = phasespace.nbody_decay(B_MASS, [PION_MASS, PION_MASS, PION_MASS])
t0 = time.perf_counter(); samples = decay.generate(N_EVENTS); torch.cuda.synchronize()
print(f"Initial run: {(time.perf_counter() - t0) * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(n_runs):
    samples = decay.generate(N_EVENTS)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n_runs
print(f"Elapsed time: {dt * 1e3:.4f} ms per call of {N_EVENTS} events = {N_EVENTS / dt:.3e} events/s")
print("nevents produced", samples[0].shape, "Shape of one particle momentum", samples[1]["p_0"].shape)
