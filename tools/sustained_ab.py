#!/usr/bin/env python3
"""Burst (10 launches) and sustained (>= 2.5 s back to back, board at its power limit) time per launch of a flat decay,
with and without the statistics buffer, for the library named by PS_B200_LIB:
   python tools/sustained_ab.py [n_body] [n_events] [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

n_body = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5
comp = ps.nbody_decay(5279.58, [139.57018] * n_body).compile()
dev = torch.device("cuda")
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((n_body, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
name = os.path.basename(os.environ.get("PS_B200_LIB") or "default")
for label, stats in (("stats", torch.zeros((4,), dtype=torch.float64, device=dev)), ("no stats", None)):
    time.sleep(1.0)  # let the power average decay: every case starts from a cool board
    for _ in range(3):
        comp.launch(n, seed=1, out=out, err=err, stats=stats)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        comp.launch(n, seed=1, out=out, err=err, stats=stats)
    b.record()
    torch.cuda.synchronize()
    burst = a.elapsed_time(b) / 10
    marks, t0, done = [], time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(25):
            comp.launch(n, seed=1, out=out, err=err, stats=stats)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        ev.synchronize()
        marks.append(ev)
        done += 25
    half = len(marks) // 2
    sustained = marks[half].elapsed_time(marks[-1]) / (25 * (len(marks) - 1 - half))
    print(f"{name:20s} n_body={n_body} {label:8s} burst {burst:.4f} ms  sustained (second half of {seconds} s) {sustained:.4f} ms")
