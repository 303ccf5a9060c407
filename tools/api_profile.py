#!/usr/bin/env python3
"""Where the host time of a small generate() call goes (the reference benchmark's protocol: 1e6 events per call).
   python tools/api_profile.py [n_events] [n_body]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
decay = ps.nbody_decay(5279.65, [139.57018, 493.677][:nb] if nb == 2 else [139.57018] * nb)
dev = torch.device("cuda")


def loop(fn, reps=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return f"issue {1e6 * (t1 - t0) / reps:7.2f} us  total {1e6 * (t2 - t0) / reps:7.2f} us"


comp = decay.compile()
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((comp.n_particles, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
print("generate(n, seed=int)        ", loop(lambda: decay.generate(n, seed=3)))
print("generate(n) global generator ", loop(lambda: decay.generate(n)))
print("launch, outputs allocated    ", loop(lambda: comp.launch(n, 1)))
print("launch, outputs reused       ", loop(lambda: comp.launch(n, 1, out=out, err=err)))
if hasattr(decay, "generator"):
    g = decay.generator(n)
    print("replay object                ", loop(lambda: g(seed=3)))
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    decay.generate(n, seed=3)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4000])
