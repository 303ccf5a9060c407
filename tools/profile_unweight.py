#!/usr/bin/env python3
"""Stage-by-stage timing (CUDA events) of one unweighting chunk of BASELINE config 5: mask -> scan -> index -> regenerate.
   python tools/profile_unweight.py [n_body] [chunk_events]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
from phasespace_b200 import _lib
from phasespace_b200.phasespace import _ptr

n_body = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 27
dev = torch.device("cuda")
decay = ps.nbody_decay(5279.58, [139.57018] * n_body)
comp = decay.compile()
stats = torch.zeros(4, dtype=torch.float64, device=dev)
comp.launch(4_000_000, 2, stats=stats)
bound = 1.5 * float(stats[0])
lib = _lib.lib
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
n_tiles = (n + _lib.TILE - 1) // _lib.TILE
mask = torch.empty(((n + 31) // 32,), dtype=torch.int32, device=dev)
offsets = torch.empty((n_tiles,), dtype=torch.int64, device=dev)
total = torch.zeros((1,), dtype=torch.int64, device=dev)
err = torch.zeros((1,), dtype=torch.int32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for rep in range(3):
    ev[0].record()
    _lib.check(lib.ps_unweight_mask(comp.handle, 1, 0, n, bound, _ptr(mask), None, _ptr(err), 0, stream))
    ev[1].record()
    _lib.check(lib.ps_unweight_scan(_ptr(mask), n, _ptr(offsets), _ptr(total), stream))
    ev[2].record()
    n_acc = int(total.item())
    scan_ms = ev[1].elapsed_time(ev[2])
    index = torch.empty((n_acc,), dtype=torch.int64, device=dev)
    weights = torch.empty((n_acc,), dtype=torch.float64, device=dev)
    slab = torch.empty((comp.n_particles, n_acc, 4), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ev[2] = torch.cuda.Event(enable_timing=True)
    ev[2].record()
    _lib.check(lib.ps_unweight_index(_ptr(mask), _ptr(offsets), 0, n, _ptr(index), stream))
    ev[3].record()
    _lib.check(lib.ps_generate_indexed(comp.handle, 1, _ptr(index), n_acc, None, 0, None, _ptr(weights), None, _ptr(slab),
                                       4 * n_acc, None, _ptr(err), _lib.PS_GEN_NORMALIZE, stream))
    ev[4].record()
    torch.cuda.synchronize()
    print(f"n_body={n_body} chunk={n} accepted={n_acc} ({n_acc / n:.4f}): mask {ev[0].elapsed_time(ev[1]):.3f} ms, "
          f"scan {scan_ms:.3f} ms, index {ev[2].elapsed_time(ev[3]):.3f} ms, regenerate {ev[3].elapsed_time(ev[4]):.3f} ms "
          f"({n_acc / ev[3].elapsed_time(ev[4]) / 1e6:.2f}e9 events/s)")
