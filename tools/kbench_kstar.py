#!/usr/bin/env python3
"""Kernel-only timing of B0 -> K*0(-> K pi) gamma for the library named by PS_B200_LIB:
   python tools/kbench_kstar.py [n_events] [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
from phasespace_b200.fromdecay import mass_functions as mf

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((4, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
name = os.path.basename(os.environ.get("PS_B200_LIB") or "default")
for kind, fac in (("fixed", None), ("bw", mf.breitwigner_factory), ("gauss", mf.gauss_factory), ("relbw", mf.relativistic_breitwigner_factory)):
    kst = ps.GenParticle("K*0", 895.81 if fac is None else fac(895.81, 47.4))
    comp = ps.GenParticle("B0", 5279.58).set_children(
        kst.set_children(ps.GenParticle("K+", 493.677), ps.GenParticle("pi-", 139.57018)), ps.GenParticle("gamma", 0.0)).compile()
    for _ in range(3):
        comp.launch(n, seed=1, out=out, err=err)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        comp.launch(n, seed=1, out=out, err=err)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name:20s} K* {kind:6s} ms={ms:.4f} ev/s={n / ms * 1e3:.4g} GB/s={136 * n / ms / 1e6:.0f}")
