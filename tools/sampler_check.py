#!/usr/bin/env python3
"""Accuracy of the device resonance samplers against the numpy oracle over a grid of (mass, width, lo, hi):
   prints max |x_gpu - x_oracle| / hi per case, for the stand-alone kernel (per-event limits) and for the fused
   generator (event-independent limits, host-precomputed)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import phasespace_b200 as ps
from phasespace_b200.fromdecay import mass_functions as mf
from oracle import genbod_numpy as orc
from oracle.philox import uniform_table

n = 1 << 18
cases = [(895.81, 47.4, 633.247, 5279.58), (895.81, 47.4, 633.247, 1500.0), (895.81, 47.4, 633.247, 700.0),
         (895.81, 47.4, 1000.0, 5279.58), (895.81, 47.4, 3000.0, 5279.58), (1272.0, 90.0, 1100.0, 5000.0),
         (775.0, 149.0, 280.0, 5000.0), (775.0, 149.0, 280.0, 1e5), (3096.9, 0.0929, 3000.0, 3200.0),
         (3096.9, 0.0929, 200.0, 5279.0), (895.81, 47.4, 0.0, 5279.58), (895.81, 47.4, 895.0, 896.5),
         (500.0, 500.0, 280.0, 5000.0), (134.9768, 0.5, 0.0, 1200.0)]
worst = 0.0
for kind, fac in ((orc.RELBW, mf.relativistic_breitwigner_factory), (orc.BW, mf.breitwigner_factory), (orc.GAUSS, mf.gauss_factory)):
    for (m, w, lo, hi) in cases:
        u = uniform_table(7, 0, n, 1, stream=2)[:, 0]
        ref = orc.sample_mass(kind, m, w, np.full(n, lo), np.full(n, hi), u)
        got = fac(m, w)(lo, hi, n, seed=7).cpu().numpy()
        err = np.abs(got - ref) / hi
        # per-event limits that differ from event to event
        lo_v = lo + (hi - lo) * 0.3 * uniform_table(8, 0, n, 1, stream=2)[:, 0]
        ref_v = orc.sample_mass(kind, m, w, lo_v, np.full(n, hi), u)
        got_v = fac(m, w)(torch.as_tensor(lo_v), hi, n, seed=7).cpu().numpy()
        err_v = np.abs(got_v - ref_v) / hi
        # fused, event-independent limits: X -> R(-> a b) c with lo = ma + mb, hi = M - mc
        ma = lo * 0.4
        mb = lo - ma
        mc = 10.0
        M = hi + mc
        fused = float("nan")
        if ma > 0:
            dec = ps.GenParticle("X", M).set_children(
                ps.GenParticle("R", fac(m, w)).set_children(ps.GenParticle("a", ma), ps.GenParticle("b", mb)),
                ps.GenParticle("c", mc))
            _, parts = dec.generate(n, seed=7)
            r = parts["R"].cpu().numpy()
            mass = np.sqrt(np.maximum(r[:, 3] ** 2 - (r[:, :3] ** 2).sum(1), 0))
            ug = uniform_table(7, 0, n, 5)[:, 0]
            refm = orc.sample_mass(kind, m, w, np.full(n, ma + mb), np.full(n, M - mc), ug)
            fused = np.max(np.abs(mass - refm) / M)
        worst = max(worst, err.max(), err_v.max())
        print(f"kind={kind} m={m} w={w} lo={lo} hi={hi}: standalone {err.max():.2e} (99.9%: {np.quantile(err, 0.999):.1e})  "
              f"varying-lo {err_v.max():.2e}  fused-constlim(mass from 4-vector) {fused:.2e}", flush=True)
print("worst", worst)
