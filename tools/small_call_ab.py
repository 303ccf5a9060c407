#!/usr/bin/env python3
"""Kernel time of small calls of the memory-bound trees (two bodies, K* gamma with a fixed K*) for the launch shape selected
by PS_B200_STRIDED_BELOW (read once per process): CUDA events over many launches rotating over four output slabs, as in
bench.py's config 1.    PS_B200_STRIDED_BELOW=16 python tools/small_call_ab.py [sizes, comma separated]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1e5,3e5,1e6,3e6,1e7".split(","))]
dev = torch.device("cuda")
trees = {
    "two_body": ps.nbody_decay(5279.65, [139.57018, 493.677]),
    "kstar_gamma_fixed": ps.GenParticle("B0", 5279.58).set_children(
        ps.GenParticle("K*0", 895.81).set_children(ps.GenParticle("K+", 493.677), ps.GenParticle("pi-", 139.57018)),
        ps.GenParticle("gamma", 0.0)),
}
tag = os.environ.get("PS_B200_STRIDED_BELOW", "default")
for name, tree in trees.items():
    comp = tree.compile()
    P = comp.n_particles
    for n in sizes:
        slabs = [{"weights": torch.empty((n,), dtype=torch.float64, device=dev),
                  "particles": torch.empty((P, n, 4), dtype=torch.float64, device=dev)} for _ in range(4)]
        err = torch.zeros((1,), dtype=torch.int32, device=dev)
        reps = max(40, min(2000, int(2e9 / n)))
        for k in range(20):
            comp.launch(n, seed=1, out=slabs[k % 4], err=err)
        torch.cuda.synchronize()
        best = 1e9
        for trial in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for k in range(reps):
                comp.launch(n, seed=1, out=slabs[k % 4], err=err)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        bpe = 8 + 32 * P
        print(f"strided_below={tag:8s} {name:18s} n={n:>9d} us={best * 1e3:8.2f} GB/s={bpe * n / best / 1e6:7.0f}", flush=True)
        del slabs
