import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from phasespace_b200.fromdecay import GenMultiDecay
from tests.helpers import example_decay_chains
c = GenMultiDecay.from_dict(example_decay_chains.dstarplus_big_decay, tolerance=1e-10)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000
for _ in range(20): c.generate(n, seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter(); c.generate(n, seed=1); torch.cuda.synchronize(); print("one call", (time.perf_counter() - t0) * 1e3, "ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): c.generate(n, seed=1)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
