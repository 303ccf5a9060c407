import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
dev = torch.device("cuda")
for mb in (64, 256, 1024, 4096):
    n = mb * 1024 * 1024 // 8
    d = torch.empty(n, dtype=torch.float64, device=dev)
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.copy_(d); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"D2H pinned {mb} MiB: {n*8/dt/1e9:.1f} GB/s")
    t0 = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"H2D pinned {mb} MiB: {n*8/dt/1e9:.1f} GB/s")
decay = ps.nbody_decay(5279.58, [139.57018] * 3)
n = 50_000_000
host = {"weights": torch.empty((n,), dtype=torch.float64, pin_memory=True),
        "particles": torch.empty((3, n, 4), dtype=torch.float64, pin_memory=True)}
for chunk in (1 << 20, 1 << 22, 1 << 24, 1 << 26):
    decay.generate_host(n, seed=1, out=host, chunk_events=chunk)
    t0 = time.perf_counter()
    for _ in range(3): decay.generate_host(n, seed=1, out=host, chunk_events=chunk)
    dt = (time.perf_counter() - t0) / 3
    print(f"generate_host chunk={chunk}: {n/dt:.3e} ev/s  {104*n/dt/1e9:.1f} GB/s")
os.system("nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -E 'NUMA|Model name|Socket' | head -8")
