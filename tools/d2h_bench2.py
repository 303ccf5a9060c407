import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
decay = ps.nbody_decay(5279.58, [139.57018] * 3)
n = 32_000_000
host = {"weights": torch.empty((n,), dtype=torch.float64, pin_memory=True),
        "particles": torch.empty((3, n, 4), dtype=torch.float64, pin_memory=True)}
for chunk in (1 << 20, 1 << 22):
    decay.generate_host(n, seed=1, out=host, chunk_events=chunk)
    t0 = time.perf_counter()
    decay.generate_host(n, seed=1, out=host, chunk_events=chunk)
    dt = (time.perf_counter() - t0)
    print(f"generate_host chunk={chunk}: {n/dt:.3e} ev/s  {104*n/dt/1e9:.1f} GB/s total {dt:.3f}s", flush=True)
