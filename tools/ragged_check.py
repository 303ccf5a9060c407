#!/usr/bin/env python3
"""Small ragged launches of every kernel family, with their own bounds and ordering assertions (compute-sanitizer is not available on the GPU pool):
   python tools/sanitize_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
from phasespace_b200.unweight import generate_unweighted
from phasespace_b200.histogram import generate_histograms

for n_body, n in ((2, 1001), (3, 70_001), (4, 33_333), (10, 5_003)):
    w, parts = ps.nbody_decay(5279.58, [139.57018] * n_body).generate(n, seed=3)
    assert w.shape == (n,) and all(p.shape == (n, 4) for p in parts.values())
for n_body, n, bound in ((4, 100_003, 1.0), (4, 255, 1.0), (10, 70_001, 1e-6)):
    parts, idx, w = generate_unweighted(ps.nbody_decay(5279.58, [139.57018] * n_body), n, 5, w_bound=bound)
    assert idx.numel() == w.numel() and (idx.numel() == 0 or int(idx.max()) < n)
    assert bool((idx[1:] > idx[:-1]).all())
generate_histograms(ps.nbody_decay(5279.58, [139.57018] * 3), 50_001, seed=1)
torch.cuda.synchronize()
print("ragged_check done")
