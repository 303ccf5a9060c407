#!/usr/bin/env python3
"""Per-launch time series of B0 -> K*0(-> K pi) gamma for each K* mass model (kernel only, CUDA events around every
launch), with the SM clock and board power read through NVML before and after each series:
   python tools/kbench_kstar_series.py [n_events] [launches] [kinds, comma separated]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import pynvml
import torch
import phasespace_b200 as ps
from phasespace_b200.fromdecay import mass_functions as mf

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fixed", "bw", "gauss", "relbw"]
FAC = {"fixed": None, "bw": mf.breitwigner_factory, "gauss": mf.gauss_factory, "relbw": mf.relativistic_breitwigner_factory}
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
def state():
    return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
dev = torch.device("cuda")
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((4, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
for kind in kinds:
    fac = FAC[kind]
    kst = ps.GenParticle("K*0", 895.81 if fac is None else fac(895.81, 47.4))
    comp = ps.GenParticle("B0", 5279.58).set_children(
        kst.set_children(ps.GenParticle("K+", 493.677), ps.GenParticle("pi-", 139.57018)), ps.GenParticle("gamma", 0.0)).compile()
    comp.launch(n, seed=1, out=out, err=err)
    torch.cuda.synchronize()
    time.sleep(1.0)  # cool down: every series starts from an idle board
    before = state()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        comp.launch(n, seed=1, out=out, err=err)
        ev[k + 1].record()
    torch.cuda.synchronize()
    after = state()
    ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(reps)])
    print(f"K* {kind:6s} first={np.round(ms[:4], 3)} median={np.median(ms):.4f} last={np.round(ms[-4:], 3)} "
          f"clock/power before={before} after={after}", flush=True)
