import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
import phasespace_b200 as ps
decay = ps.nbody_decay(5279.0, [139.6] * 3)
n = 1_000_000
for _ in range(20): decay.generate(n)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        w, p = decay.generate(n)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
# reuse outputs through the compiled handle (no allocation)
comp = decay.compile()
dev = torch.device("cuda")
out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev), "particles": torch.empty((3, n, 4), dtype=torch.float64, device=dev)}
err = torch.zeros((1,), dtype=torch.int32, device=dev)
for _ in range(10): comp.launch(n, 1, out=out, err=err)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): comp.launch(n, 1, out=out, err=err)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
print(f"compiled.launch reuse: {dt*1e6:.1f} us per call  {n/dt:.3e} ev/s")
