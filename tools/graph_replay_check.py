#!/usr/bin/env python3
"""Would a CUDA-graph replay object make small generate() calls cheaper?  Lower bound: replaying a captured single-kernel
graph with FIXED parameters (same seed, same event range, same buffers -- useless as a generator, a real replay object would
add cudaGraphExecKernelNodeSetParams per call) against the direct launch path with reused buffers.
   python tools/graph_replay_check.py [n_events]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")
comp = ps.nbody_decay(5279.65, [139.57018, 493.677]).compile()
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev),
       "particles": torch.empty((2, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)


def loop(fn, reps=3000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return f"issue {1e6 * (t1 - t0) / reps:6.2f} us per call, total {1e6 * (t2 - t0) / reps:6.2f} us per call"


side = torch.cuda.Stream()
with torch.cuda.stream(side):
    comp.launch(n, 1, out=out, err=err)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        comp.launch(n, 1, out=out, err=err)
torch.cuda.synchronize()
print(f"n_events = {n}")
print("direct launch, buffers reused (CompiledDecay.launch):", loop(lambda: comp.launch(n, 1, out=out, err=err)))
print("graph.replay(), fixed parameters                    :", loop(graph.replay))
decay = ps.nbody_decay(5279.65, [139.57018, 493.677])
print("generate(n) (allocates outputs, builds the dict)    :", loop(lambda: decay.generate(n)))
