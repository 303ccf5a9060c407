import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
decay = ps.nbody_decay(5279.0, [139.6] * 3)   # benchmark/bench_phasespace.py:95-102
for n in (1000, 1_000_000, 10_000_000):
    decay.generate(n); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        w, p = decay.generate(n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"generate({n}): {dt*1e6:.1f} us per call, {n/dt:.3e} events/s through the Python API")
