#!/usr/bin/env python3
"""Kernel-only timing of the five BASELINE.json configs on one GPU (secondary numbers; bench.py is the headline).
   python tools/bench_configs.py [scale]   -- scale < 1 shrinks n_events"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
from phasespace_b200.fromdecay import mass_functions as mf
from phasespace_b200.unweight import generate_unweighted

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda")
B0, PI, K = 5279.58, 139.57018, 493.677


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(name, decay, n, boost=None, bytes_per_event=None):
    comp = decay.compile()
    P = comp.n_particles
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
           "particles": torch.empty((P, n, 4), dtype=torch.float64, device=dev)}
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    ms = timed(lambda: comp.launch(n, seed=1, boost_to=boost, out=out, err=err))
    assert int(err.item()) == 0
    bpe = bytes_per_event or (8 + 32 * P + (32 if boost is not None and boost.shape[0] == n else 0))
    res = {"config": name, "n_events": n, "ms": ms, "events_per_s": n / ms * 1e3, "GB_per_s": bpe * n / ms / 1e6,
           "frac_of_6531": bpe * n / ms / 1e6 / 6531.3, "bytes_per_event": bpe}
    print(json.dumps(res), flush=True)
    del out
    torch.cuda.empty_cache()


n8 = int(1e8 * scale)
run("1: B0->K pi, n=1e6", ps.nbody_decay(5279.65, [PI, K]), 1_000_000)
run("2: B0->3pi, n=1e8", ps.nbody_decay(B0, [PI] * 3), n8)
g = torch.Generator(device=dev).manual_seed(1234)
pt = 1000.0 * torch.empty(n8, dtype=torch.float64, device=dev).exponential_(1 / 5.0, generator=g)
eta = 2 + 3 * torch.rand(n8, dtype=torch.float64, device=dev, generator=g)
phi = 6.283185307179586 * torch.rand(n8, dtype=torch.float64, device=dev, generator=g)
px, py, pz = pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta)
boost = torch.stack([px, py, pz, torch.sqrt(px * px + py * py + pz * pz + B0 * B0)], dim=1).contiguous()
del pt, eta, phi, px, py, pz
run("3: B0->4pi + per-event boost_to, n=1e8", ps.nbody_decay(B0, [PI] * 4), n8, boost=boost)
del boost
torch.cuda.empty_cache()
for kind, fac in (("relbw", mf.relativistic_breitwigner_factory), ("gauss", mf.gauss_factory), ("bw", mf.breitwigner_factory),
                  ("fixed-mass", None)):
    chain = ps.GenParticle("B0", B0).set_children(
        ps.GenParticle("K*0", 895.81 if fac is None else fac(895.81, 47.4)).set_children(ps.GenParticle("K+", K), ps.GenParticle("pi-", PI)),
        ps.GenParticle("gamma", 0.0))
    run(f"4: B0->K*0(->K pi) gamma, {kind} K*0, n=1e8", chain, n8)
# config 4 as BASELINE.json words it: through GenMultiDecay (public API: allocates its outputs, one host check per call)
from phasespace_b200.fromdecay import GenMultiDecay
kst = ps.GenParticle("B0", B0).set_children(
    ps.GenParticle("K*0", mf.relativistic_breitwigner_factory(895.81, 47.4)).set_children(ps.GenParticle("K+", K), ps.GenParticle("pi-", PI)),
    ps.GenParticle("gamma", 0.0))
multi = GenMultiDecay([(1.0, kst)])
def _multi():
    w, ev = multi.generate(n8, seed=1)
    del w, ev
ms = timed(_multi, reps=5)
print(json.dumps({"config": "4m: the same chain through GenMultiDecay([(1.0, tree)]).generate (API call, outputs allocated per call)",
                  "n_events": n8, "ms": ms, "events_per_s": n8 / ms * 1e3}), flush=True)
ten = ps.nbody_decay(B0, [PI] * 10)
run("5a: 10-body, all events stored, n=5e7", ten, int(5e7 * scale))
n5 = int(2.5e8 * scale)
probe_w, _ = ten.generate(2_000_000, seed=3)
bound = float(probe_w.max()) * 1.5
ms = timed(lambda: generate_unweighted(ten, n5, 1, w_bound=bound), reps=3)
parts, idx, w = generate_unweighted(ten, n5, 1, w_bound=bound)
print(json.dumps({"config": "5b: 10-body accept-reject unweighting (mask+scan+index+regenerate)", "n_candidates": n5,
                  "w_bound": bound, "accepted": int(idx.numel()), "ms": ms, "candidates_per_s": n5 / ms * 1e3,
                  "accepted_per_s": idx.numel() / ms * 1e3, "max_w_seen": float(w.max())}), flush=True)

# fused histograms: nothing but 13 x 100 doubles leaves the chip
from phasespace_b200.histogram import generate_histograms
three = ps.nbody_decay(B0, [PI] * 3)
ms = timed(lambda: generate_histograms(three, n8, seed=1), reps=5)
print(json.dumps({"config": "2h: B0->3pi fused weighted histograms (13 x 100 bins), nothing stored", "n_events": n8, "ms": ms,
                  "events_per_s": n8 / ms * 1e3}), flush=True)
