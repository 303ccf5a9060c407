#!/usr/bin/env python3
"""GenMultiDecay.generate on the 21-mode D*+ example of the reference's tests (from_dict, relativistic Breit-Wigner D0 with
its physical width): time per call for small and large n.   python tools/multidecay_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from phasespace_b200.fromdecay import GenMultiDecay
from tests.helpers import example_decay_chains

torch.cuda.init()
if "--cold" in sys.argv:  # first call with an empty on-disk kernel cache: table builds + NVRTC of every mode
    import tempfile
    os.environ["PS_B200_CACHE_DIR"] = tempfile.mkdtemp(prefix="ps_b200_cold_")
t0 = time.perf_counter()
c = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay, tolerance=1e-10)
t1 = time.perf_counter()
out = c.generate(1_000_000, seed=0)
torch.cuda.synchronize()
print(f"modes: {len(c.gen_particles)}; from_dict {1e3 * (t1 - t0):.0f} ms; first generate(1e6) (trees, node tables, kernels of "
      f"{len(out[0])} populated modes) {time.perf_counter() - t1:.2f} s on {os.cpu_count()} cores")
del out
for n in (1000, 1_000_000, 50_000_000):
    reps = 20 if n <= 1_000_000 else 5
    for k in range(reps):  # the same seeds as the timed calls: every mode they populate is compiled before the clock starts
        out = c.generate(n, seed=k)
        del out
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        out = c.generate(n, seed=k)
        del out
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"n = {n:>9d}: {dt * 1e3:8.3f} ms per call, {n / dt:.3g} events/s")
