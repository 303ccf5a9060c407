#!/usr/bin/env python3
"""BASELINE.json config 5: 10-body flat decay, N candidates sharded over the ranks by contiguous index ranges,
accept-reject unweighting fused with generation (only accepted events are ever stored).

    torchrun --nproc-per-node G tools/bench_config5.py [--events 1e10] [--chunk 134217728]

Each rank walks its range in chunks: weights-only mask pass -> ordered index list -> regenerate the accepted events.
The only collective is the final all-reduce of the accepted count (and of the weight maximum used as the bound)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import phasespace_b200 as ps
from phasespace_b200 import sharding
from phasespace_b200.unweight import generate_unweighted

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=float, default=1e10)
ap.add_argument("--chunk", type=int, default=1 << 27)
ap.add_argument("--seed", type=int, default=1)
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_total = int(args.events)
first, count = sharding.shard_range(n_total, rank, world)
decay = ps.nbody_decay(5279.58, [139.57018] * 10)

# acceptance bound: global maximum of the normalised weight over a pilot sample (one tiny MAX all-reduce), x1.5
stats = torch.zeros(4, dtype=torch.float64, device=dev)
decay.compile().launch(4_000_000, args.seed + 1, first_event=rank * 4_000_000, stats=stats)
bound = 1.5 * float(sharding.allreduce_stats(stats)[0])

def run():
    kept, done, wmax_seen = 0, 0, 0.0
    while done < count:
        n = min(args.chunk, count - done)
        parts, idx, w = generate_unweighted(decay, n, args.seed, w_bound=bound, first_event=first + done)
        kept += int(idx.numel())
        if idx.numel():
            wmax_seen = max(wmax_seen, float(w.max()))
        done += n
    return kept, wmax_seen

run() if count <= args.chunk else generate_unweighted(decay, args.chunk, args.seed, w_bound=bound, first_event=first)  # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
kept, wmax_seen = run()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
total_kept = sharding.allreduce_count(kept, device=dev)
if rank == 0:
    dt = float(dt)
    print(json.dumps({"config": "5: 10-body flat decay with accept-reject unweighting", "n_gpus": world, "candidates": n_total,
                      "accepted": total_kept, "acceptance": total_kept / n_total, "w_bound": bound, "max_w_seen": wmax_seen,
                      "seconds": dt, "candidates_per_s": n_total / dt, "accepted_per_s": total_kept / dt,
                      "chunk_events": args.chunk, "bytes_stored_per_accepted_event": 8 + 8 + 320}))
if world > 1:
    dist.destroy_process_group()
