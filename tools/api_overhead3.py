import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import phasespace_b200 as ps
dev = torch.device("cuda")
def loop(fn, reps=100):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6
for n in (1000, 1_000_000, 10_000_000):
    print(n, "empty(13n) issue/total us:", loop(lambda: torch.empty((13 * n,), dtype=torch.float64, device=dev)))
    keep = []
    def alloc_keep():
        t = torch.empty((13 * n,), dtype=torch.float64, device=dev); keep[:] = [t]; return t
    print(n, "empty(13n) holding previous:", loop(alloc_keep))
    decay = ps.nbody_decay(5279.0, [139.6] * 3)
    print(n, "generate issue/total us:", loop(lambda: decay.generate(n)))
    comp = decay.compile()
    print(n, "launch(alloc inside) issue/total us:", loop(lambda: comp.launch(n, 1)))
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=dev), "particles": torch.empty((3, n, 4), dtype=torch.float64, device=dev)}
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    print(n, "launch(reuse) issue/total us:", loop(lambda: comp.launch(n, 1, out=out, err=err)))
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
