#!/usr/bin/env python3
"""Headline benchmark: events/s of the fused GENBOD generator, B0 -> pi pi pi, fp64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--events E] [--impl b200|reference]

One "step" is one pass of the hot path over one batch: ONE kernel launch generating E (default 1e8,
BASELINE.json configs[1]) events per GPU into HBM-resident output buffers.  With N > 1 (torchrun,
one rank per GPU) every rank generates its own contiguous range of event indices
[rank*E, (rank+1)*E): weak scaling, no collective on the data path, bit-identical to one GPU
generating the whole range.  Prints ONE JSON line (see DESIGN.md "measurement").

--impl reference times the CPU restatement of the reference (oracle/, numpy, all host cores) on the
same workload, because the reference itself (TensorFlow) cannot be installed here.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B0_MASS, PION_MASS = 5279.58, 139.57018  # tests/helpers/decays.py:14-15 of the reference
N_BODY = 3
BYTES_PER_EVENT = 8 + 32 * N_BODY  # one fp64 weight + three (px,py,pz,E) rows: SURVEY.md 8(d)
UNIT = "events/s"
try:  # the metric string is BASELINE.json's, verbatim
    with open(os.path.join(ROOT, "BASELINE.json"), encoding="utf-8") as _fh:
        METRIC = json.load(_fh)["metric"]
except Exception:  # pragma: no cover
    METRIC = "events/sec, B0->pi pi pi fp64, 1/2/4/8 B200; % of HBM-write roofline"


def _config(n_events, world):
    return {
        "workload": f"B0->pi pi pi (nbody_decay({B0_MASS}, 3x{PION_MASS})), n_events={n_events:.3g} per GPU per step, "
                    "normalize_weights=True, one kernel launch per step",
        "events_per_gpu_per_step": int(n_events),
        "sharding": f"contiguous event-index ranges, {world} rank(s), no data-path collective",
        "l2": f"outputs {BYTES_PER_EVENT * n_events / 1e9:.2f} GB per launch >> 126 MB L2, so no L2 flush between steps",
    }


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the bench runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.sm_max = None
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        mask = 0
        for s in inside:
            mask |= s[2]
        return {
            "sm_mhz": statistics.median(s[1] for s in inside),
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit),
            "samples": len(inside),
        }


def _physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def _bind_to_gpu_numa_node(gpu_index):
    """Restrict this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end path (first touch) and the copy-engine traffic stay on the GPU's own NUMA node.  With several ranks
    writing 56 GB/s each into host memory, remote-node buffers are what limits the aggregate."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------- CPU legs
def _cpu_baseline(min_seconds=10.0):
    """Oracle (numpy port of the reference) on all host cores, bounded sample. Returns the dict for JSON."""
    from oracle.cpu_bench import NumpyOracleBench, c_genbod_rate

    bench = NumpyOracleBench(B0_MASS, [PION_MASS] * N_BODY)
    per_worker, events, seconds, passes = 200_000, 0, 0.0, 0
    while seconds < min_seconds and passes < 40:
        done, dt = bench.step(per_worker, seed=passes)
        events, seconds, passes = events + done, seconds + dt, passes + 1
    cores = bench.cores
    bench.close()
    out = {
        "value": events / seconds, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"numpy restatement of the reference (oracle/genbod_numpy.py; bit-identical to the reference's own "
                  f"source run on NumPy, tests/golden/make_reference_golden.py), {passes} passes x {cores} worker "
                  f"processes x {per_worker} events = {events} events in {seconds:.1f} s; the reference's own "
                  "TensorFlow CPU path cannot be installed here",
    }
    try:
        rate, used = c_genbod_rate(B0_MASS, [PION_MASS] * N_BODY)
        out["c_genbod"] = {"value": rate, "unit": UNIT, "cores": used,
                           "what": "scalar C GENBOD (oracle/genbod_ref.c), the bench_tgenphasespace.cxx analogue"}
    except Exception as exc:  # pragma: no cover
        out["c_genbod"] = {"error": str(exc)}
    return out


def run_reference(args, rank, world):
    """--impl reference: the oracle port on the host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle.cpu_bench import NumpyOracleBench

    bench = NumpyOracleBench(B0_MASS, [PION_MASS] * N_BODY)
    per_worker = 100_000
    for i in range(args.warmup):
        bench.step(per_worker, seed=i)
    total, t_total = 0, 0.0
    for i in range(args.steps):
        done, dt = bench.step(per_worker, seed=100 + i)
        total, t_total = total + done, t_total + dt
    cores = bench.cores
    bench.close()
    value = total / t_total
    sample = (f"each step = {cores} worker processes x {per_worker} events of the numpy restatement "
              "(oracle/genbod_numpy.py; bit-identical to the reference's own source run on NumPy); reference TensorFlow "
              "path not installable here")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(args.events, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------- B200 leg
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import phasespace_b200 as ps
    from phasespace_b200 import _lib

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = _cpu_baseline()  # before CUDA is busy; worker processes are spawned, not forked

    numa_cpus = _bind_to_gpu_numa_node(_physical_gpu_index(local_rank)) if world > 1 else None
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = int(args.events)
    decay = ps.nbody_decay(B0_MASS, [PION_MASS] * N_BODY)
    comp = decay.compile()
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=device),
           "particles": torch.empty((N_BODY, n, 4), dtype=torch.float64, device=device)}
    err = torch.zeros((1,), dtype=torch.int32, device=device)
    stats = torch.zeros((4,), dtype=torch.float64, device=device)
    first_event = rank * n
    stream = torch.cuda.current_stream(device)

    def step():
        comp.launch(n, seed=args.seed, first_event=first_event, out=out, stats=stats, err=err, stream=stream.cuda_stream)

    sampler = ClockSampler(_physical_gpu_index(local_rank))
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    stats.zero_()
    launches0 = _lib.lib.ps_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t_start.record(stream)
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    launches = _lib.lib.ps_launch_count() - launches0
    barrier()
    total_ms = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=device)
    kernel_ms = torch.tensor([statistics.fmean(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(total_ms.item()), float(kernel_ms.item())
    clocks = sampler.summary(wall0, wall1)
    gstats = stats.clone()  # statistics of exactly the timed launches
    if clocks["samples"] < 3:  # timed region shorter than the sampling period: probe the same kernel a little longer
        p0 = time.perf_counter()
        while time.perf_counter() - p0 < 0.6:
            step()
            torch.cuda.synchronize()
        clocks = sampler.summary(p0, time.perf_counter())
        clocks["note"] = "timed region too short to sample; taken from an immediately following probe of the same launches"

    assert int(err.item()) == 0, "forbidden decay flagged"
    # the only collective of the path: the tiny weight-statistics allreduce (outside the timed loop)
    if world > 1:
        mx = gstats[[0, 3]].clone()
        sm = gstats[[1, 2]].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        gstats = torch.stack([mx[0], sm[0], sm[1], mx[1]])
    n_acc = float(world) * n * args.steps
    global_stats = {"max_w": float(gstats[0]), "mean_w": float(gstats[1]) / n_acc}

    # ---- end to end through the public API with HOST output buffers (D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_n = n
        host = None
        while host is None and e2e_n >= 1 << 20:
            try:
                host = {"weights": torch.empty((e2e_n,), dtype=torch.float64, pin_memory=True),
                        "particles": torch.empty((N_BODY, e2e_n, 4), dtype=torch.float64, pin_memory=True)}
            except RuntimeError:
                host, e2e_n = None, e2e_n // 2
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        decay.generate_host(e2e_n, seed=args.seed, out=host)  # warm-up (touches the pinned pages)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            w_host, _parts = decay.generate_host(e2e_n, seed=args.seed, out=host)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_n * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": BYTES_PER_EVENT * e2e_n, "steps": e2e_steps, "events_per_gpu_per_step": e2e_n,
               "api": "GenParticle.generate_host -> ps_generate_host (chunked generate + async D2H into pinned host memory)",
               "checksum_w0": float(w_host[0])}
        if numa_cpus:
            e2e["host_buffers"] = f"each rank bound to the {numa_cpus} CPUs NVML reports as local to its GPU before allocating"
    # ---- supplementary: the GPU-consumer case through the public API -- events stay in HBM for a device-side
    # consumer (a fit on the same GPU); per step only a scalar metric (mean weight) is read back to the host
    e2e_device = None
    if not args.no_e2e:
        for _ in range(2):  # warm-up: the allocator ends up holding the output slab, torch has loaded its reduction
            w_dev, parts_dev = decay.generate(n, seed=args.seed)
            float(w_dev.mean().item())
            del w_dev, parts_dev
        barrier()
        dev_steps = max(1, min(args.steps, 10))
        t0 = time.perf_counter()
        for k in range(dev_steps):
            w_dev, parts_dev = decay.generate(n, seed=args.seed + k)
            mean_w = float(w_dev.mean().item())
            del w_dev, parts_dev  # the consumer is done with the events: the slab is reused by the next call
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_device = {"value": world * n * dev_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": 8, "steps": dev_steps, "mean_w": mean_w,
                      "api": "GenParticle.generate (allocates outputs) + weights.mean().item(); events stay in HBM"}
    # ---- supplementary: the histogram consumer of the reference's physics tests, fused into the generator
    # (generate_histograms): per step only 13 x 100 doubles cross PCIe
    e2e_hist = None
    if not args.no_e2e:
        from phasespace_b200.histogram import generate_histograms

        generate_histograms(decay, n, seed=args.seed, w_cap=1.0)
        barrier()
        hist_steps = max(1, min(args.steps, 10))
        t0 = time.perf_counter()
        for k in range(hist_steps):
            hp, hw = generate_histograms(decay, n, seed=args.seed + k, w_cap=1.0)
            hw_host = hw.cpu()
            hp_host = {name: h.cpu() for name, h in hp.items()}
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_hist = {"value": world * n * hist_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 8 * 100 * (4 * N_BODY + 1), "steps": hist_steps,
                    "sum_weight_histogram": float(hw_host.sum()),
                    "api": "histogram.generate_histograms -> ps_histogram (events binned in shared memory, never stored) "
                           "+ the normalised histograms copied to the host"}
    sampler.stop()

    if rank == 0:
        peak, peak_src = 6650.0, "of fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peak, peak_src = float(json.load(fh)["hbm_gbs"]), "of measured"
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                t = json.load(fh)
                if int(t.get("events_per_launch", -1)) == n:
                    traffic = t["dram_bytes_per_launch"]
        except Exception:
            pass
        achieved = BYTES_PER_EVENT * n / (kernel_ms * 1e-3) / 1e9
        # the other bound of the north star: fp64 instructions at the measured pipe peak (profiles/fp64_pipe.json,
        # tools/fp64_bench.cu; instruction count per event from the ncu pass in profiles/)
        fp64 = None
        try:
            with open(os.path.join(ROOT, "profiles", "fp64_pipe.json")) as fh:
                f = json.load(fh)
            bound = f["peak_fp64_inst_per_s"] / f["fp64_inst_per_event"]["flat_3body"]
            fp64 = {"fp64_inst_per_event": f["fp64_inst_per_event"]["flat_3body"], "peak_fp64_inst_per_s": f["peak_fp64_inst_per_s"],
                    "bound_events_per_s": bound, "frac": n / (kernel_ms * 1e-3) / bound,
                    "note": "HBM is the binding roofline for the 3-body decay; fp64 pipe peak measured with tools/fp64_bench.cu"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_EVENT * n, "kernel_ms": kernel_ms, "fp64_bound": fp64},
            "clocks": clocks, "gpu_launches": int(launches), "global_stats": global_stats,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if e2e_device is not None:
            line["e2e_device_consumer"] = e2e_device
        if e2e_hist is not None:
            line["e2e_histogram_consumer"] = e2e_hist
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--events", type=float, default=1e8, help="events per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.events = int(args.events)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
