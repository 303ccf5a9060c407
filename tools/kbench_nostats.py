#!/usr/bin/env python3
"""As kbench.py, with and without the statistics buffer: python tools/kbench_nostats.py [n_body] [n_events] [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

n_body = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
comp = ps.nbody_decay(5279.58, [139.57018] * n_body).compile()
dev = torch.device("cuda")
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((n_body, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
for label, stats in (("stats", torch.zeros((4,), dtype=torch.float64, device=dev)), ("no stats", None)):
    for _ in range(3):
        comp.launch(n, seed=1, out=out, err=err, stats=stats)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        comp.launch(n, seed=1, out=out, err=err, stats=stats)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{os.path.basename(os.environ.get('PS_B200_LIB', 'default')):20s} n_body={n_body} {label:8s} ms={ms:.4f} ev/s={n / ms * 1e3:.4g}")
