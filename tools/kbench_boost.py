#!/usr/bin/env python3
"""Kernel-only timing of a flat decay with per-event boost_to (BASELINE config 3 for n_body = 4), library from PS_B200_LIB:
   python tools/kbench_boost.py [n_body] [n_events] [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

n_body = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
B0 = 5279.58
g = torch.Generator(device=dev).manual_seed(1234)
pt = 1000.0 * torch.empty(n, dtype=torch.float64, device=dev).exponential_(1 / 5.0, generator=g)
eta = 2 + 3 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
phi = 6.283185307179586 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
boost = torch.empty((n, 4), dtype=torch.float64, device=dev)
boost[:, 0], boost[:, 1], boost[:, 2] = pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta)
boost[:, 3] = torch.sqrt((boost[:, :3] ** 2).sum(1) + B0**2)
del pt, eta, phi
comp = ps.nbody_decay(B0, [139.57018] * n_body).compile()
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((n_body, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
for _ in range(3):
    comp.launch(n, seed=1, boost_to=boost, out=out, err=err)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    comp.launch(n, seed=1, boost_to=boost, out=out, err=err)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
bpe = 8 + 32 * n_body + 32
print(f"{os.path.basename(os.environ.get('PS_B200_LIB') or 'default'):20s} n_body={n_body}+boost ms={ms:.4f} ev/s={n / ms * 1e3:.4g} GB/s={bpe * n / ms / 1e6:.0f} frac={bpe * n / ms / 1e6 / 6531.3:.3f}")
