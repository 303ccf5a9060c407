#!/usr/bin/env python3
"""One launch of every kernel behind the BASELINE.json configs, for an ncu metrics pass:
   python tools/profile_configs.py [n_events]
   ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,... -k regex:genbod python tools/profile_configs.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
from phasespace_b200.fromdecay import mass_functions as mf
from phasespace_b200.histogram import generate_histograms

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda")
B0, PI, K = 5279.58, 139.57018, 493.677


def once(decay, boost=None):
    comp = decay.compile()
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
           "particles": torch.empty((comp.n_particles, n, 4), dtype=torch.float64, device=dev)}
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    comp.launch(n, seed=1, boost_to=boost, out=out, err=err)
    torch.cuda.synchronize()
    assert int(err.item()) == 0


for nb in (2, 3, 4, 6, 10):
    once(ps.nbody_decay(B0, [PI] * nb))
g = torch.Generator(device=dev).manual_seed(1234)
pt = 1000.0 * torch.empty(n, dtype=torch.float64, device=dev).exponential_(1 / 5.0, generator=g)
eta = 2 + 3 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
phi = 6.283185307179586 * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
px, py, pz = pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta)
boost = torch.stack([px, py, pz, torch.sqrt(px * px + py * py + pz * pz + B0 * B0)], dim=1).contiguous()
once(ps.nbody_decay(B0, [PI] * 4), boost=boost)
for fac in (None, mf.breitwigner_factory, mf.gauss_factory, mf.relativistic_breitwigner_factory):
    kst = ps.GenParticle("K*0", 895.81 if fac is None else fac(895.81, 47.4))
    once(ps.GenParticle("B0", B0).set_children(kst.set_children(ps.GenParticle("K+", K), ps.GenParticle("pi-", PI)),
                                               ps.GenParticle("gamma", 0.0)))
# the weights-only pass of the accept-reject unweighting (BASELINE config 5), ten bodies
import ctypes
from phasespace_b200 import _lib
from phasespace_b200.phasespace import _ptr
ten = ps.nbody_decay(B0, [PI] * 10).compile()
mask = torch.empty(((n + 31) // 32,), dtype=torch.int32, device=dev)
state = torch.zeros((2,), dtype=torch.int64, device=dev)
_lib.check(_lib.lib.ps_unweight_mask(ten.handle, 1, 0, n, 1e-6, _ptr(mask), _ptr(state[:1]), _ptr(state[1:].view(torch.int32)[:1]), 0,
                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
torch.cuda.synchronize()
generate_histograms(ps.nbody_decay(B0, [PI] * 3), n, seed=1)
torch.cuda.synchronize()
print("done")
