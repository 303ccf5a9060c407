#!/usr/bin/env python3
"""Headline benchmark: events/s of the fused GENBOD generator, B0 -> pi pi pi, fp64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--events E] [--impl b200|reference]

One "step" is one pass of the hot path over one batch: ONE kernel launch generating E (default 1e8,
BASELINE.json configs[1]) events per GPU into HBM-resident output buffers.  With N > 1 (torchrun,
one rank per GPU) every rank generates its own contiguous range of event indices
[rank*E, (rank+1)*E): weak scaling, no collective on the data path, bit-identical to one GPU
generating the whole range.  Prints ONE JSON line (see DESIGN.md "measurement"):

    value / roofline   the headline: K launches of config 2, a BURST figure (the timed region is tens of ms)
    sustained          the same launches back to back for >= 2 s with their own clocks / power record
    configs            all five BASELINE.json configs: kernel ms, events/s, both bounds, fraction, clocks;
                       config 5 (10-body accept-reject unweighting) through generate_unweighted on N ranks
    e2e                the public host-delivery call (generate_host), D2H inside the timed region
    d2h_probe          raw pinned-memory D2H bandwidth of every rank at the same time: the ceiling of e2e
    cpu_baseline       the numpy port of the reference on the host cores (N = 1 only)

--impl reference times the CPU restatement of the reference (oracle/, numpy, all host cores) on a bounded
sample of the same workload, because the reference itself (TensorFlow) cannot be installed here.
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

B0_MASS, PION_MASS, KAON_MASS = 5279.58, 139.57018, 493.677  # tests/helpers/decays.py:14-20 of the reference
KSTAR_MASS, KSTAR_WIDTH = 895.81, 47.4
N_BODY = 3
BYTES_PER_EVENT = 8 + 32 * N_BODY  # one fp64 weight + three (px,py,pz,E) rows: SURVEY.md 8(d)
UNIT = "events/s"
WORKLOAD = (f"B0->pi pi pi (nbody_decay({B0_MASS}, 3x{PION_MASS})), normalize_weights=True, fp64, "
            "BASELINE.json configs[1]")
try:  # the metric string is BASELINE.json's, verbatim
    with open(os.path.join(ROOT, "BASELINE.json"), encoding="utf-8") as _fh:
        METRIC = json.load(_fh)["metric"]
except Exception:  # pragma: no cover
    METRIC = "events/sec, B0->pi pi pi fp64, 1/2/4/8 B200; % of HBM-write roofline"


def _config(n_events, world):
    return {
        "workload": WORKLOAD,
        "events_per_gpu_per_step": int(n_events),
        "launches_per_step": 1,
        "sharding": f"contiguous event-index ranges, {world} rank(s), no data-path collective",
        "l2": f"outputs {BYTES_PER_EVENT * n_events / 1e9:.2f} GB per launch >> 126 MB L2, so no L2 flush between steps",
        "buffers": "every step writes the same preallocated output slab with the same seed: the kernel is a pure write "
                   "stream, so neither the cache state nor the values depend on the step",
    }


def _load_json(*path):
    try:
        with open(os.path.join(ROOT, *path), encoding="utf-8") as fh:
            return json.load(fh)
    except Exception:
        return None


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock, board power and throttle reasons of one GPU through NVML while the bench runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []  # (t, sm_mhz, reasons_mask, power_w)
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
                try:
                    power = self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    power = None
                self.samples.append((time.perf_counter(), mhz, mask, power))
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
        out = {
            "sm_mhz": statistics.median(s[1] for s in inside),
            "sm_mhz_min": min(s[1] for s in inside),
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit),
            "samples": len(inside),
        }
        powers = [s[3] for s in inside if s[3] is not None]
        if powers:
            out["power_w_median"], out["power_w_max"] = statistics.median(powers), max(powers)
        return out


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
    end-to-end path (first touch) and the copy-engine traffic stay on the GPU's own NUMA node.  (A no-op on hosts that
    expose a single node, like the virtualised 8-GPU boxes of this pool: reported as such.)"""
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
_PORT_NOTE = ("numpy restatement of the reference (oracle/genbod_numpy.py; bit-identical to the reference's own source "
              "run on NumPy, tests/golden/make_reference_golden.py); the reference's own TensorFlow CPU path cannot be "
              "installed here, so this is a PORT of it, not TensorFlow")


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
        "sample": f"{_PORT_NOTE}: {passes} passes x {cores} worker processes x {per_worker} events = {events} events "
                  f"in {seconds:.1f} s",
    }
    try:
        rate, used = c_genbod_rate(B0_MASS, [PION_MASS] * N_BODY)
        out["c_genbod"] = {"value": rate, "unit": UNIT, "cores": used,
                           "what": "scalar C GENBOD (oracle/genbod_ref.c), the bench_tgenphasespace.cxx analogue"}
    except Exception as exc:  # pragma: no cover
        out["c_genbod"] = {"error": str(exc)}
    return out


def run_reference(args, rank, world):
    """--impl reference: the oracle port on the host cores; rank 0 only.  Each step is a BOUNDED SAMPLE of the
    workload (cores x 100 000 events, not the 1e8 of the GPU arm) and the config says so."""
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
    per_step = cores * per_worker
    sample = (f"{_PORT_NOTE}; each step = {cores} worker processes x {per_worker} events = {per_step} events, a bounded "
              f"sample of the {args.events:.3g}-event step of the GPU arm (a rate is a rate)")
    print(json.dumps({
        "impl": "reference", "reference_kind": "numpy port of the reference, NOT TensorFlow",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step": per_step, "cores": cores,
                   "worker_processes": cores, "events_per_worker_per_step": per_worker,
                   "sample_of": f"the GPU arm's step of {int(args.events)} events per GPU",
                   "sharding": "host cores only, rank 0"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------- B200 leg
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import phasespace_b200 as ps
    from phasespace_b200 import _lib, sharding
    from phasespace_b200.fromdecay import GenMultiDecay
    from phasespace_b200.fromdecay import mass_functions as mf
    from phasespace_b200.unweight import generate_unweighted

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

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"
    peaks = _load_json("MEASURED_PEAKS.json")
    if peaks and "hbm_gbs" in peaks:
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth, burst)"
    fp64_pipe = _load_json("profiles", "fp64_pipe.json") or {}
    fp64_peak = fp64_pipe.get("peak_fp64_inst_per_s")
    fp64_counts = fp64_pipe.get("fp64_inst_per_event", {})

    n = int(args.events)
    stream = torch.cuda.current_stream(device)
    sampler = ClockSampler(_physical_gpu_index(local_rank))
    sampler.start()

    def timed_launches(step, reps, warmup=2):
        """CUDA-event time per launch of `step` (ms, max over ranks) over `reps` back-to-back launches + clocks.
        Every config starts from an idle, cooled board (1 s) and its timed region lasts 30-50 ms, like the headline's:
        a burst figure at full clock (the board's power limit starts to bite after ~50-100 ms of load)."""
        barrier()
        time.sleep(1.0)
        for _ in range(warmup):
            step(0)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        a.record(stream)
        for k in range(reps):
            step(k)
        b.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        ms = max_over_ranks(a.elapsed_time(b) / reps)
        return ms, sampler.summary(w0, w1)

    def bounds(events_per_s, bytes_per_event, fp64_key):
        """Both rooflines of the north star: bytes at the HBM peak, fp64 instructions at the measured pipe peak."""
        out = {"bytes_per_event": bytes_per_event, "hbm_bound_events_per_s": peak * 1e9 / bytes_per_event,
               "achieved_gbs": events_per_s * bytes_per_event / 1e9}
        hbm_frac = events_per_s / out["hbm_bound_events_per_s"]
        out["frac_of_hbm_bound"] = hbm_frac
        count = fp64_counts.get(fp64_key)
        if fp64_peak and count:
            out["fp64_inst_per_event"] = count
            out["fp64_bound_events_per_s"] = fp64_peak / count
            out["frac_of_fp64_bound"] = events_per_s / out["fp64_bound_events_per_s"]
            binding = "hbm" if out["hbm_bound_events_per_s"] <= out["fp64_bound_events_per_s"] else "fp64"
        else:
            binding = "hbm"
        out["binding"] = binding
        out["frac"] = hbm_frac if binding == "hbm" else out["frac_of_fp64_bound"]
        return out

    # ================================================================== headline: config 2, K launches (burst)
    decay = ps.nbody_decay(B0_MASS, [PION_MASS] * N_BODY)
    comp = decay.compile()
    out = {"weights": torch.empty((n,), dtype=torch.float64, device=device),
           "particles": torch.empty((N_BODY, n, 4), dtype=torch.float64, device=device)}
    err = torch.zeros((1,), dtype=torch.int32, device=device)
    stats = torch.zeros((4,), dtype=torch.float64, device=device)
    first_event = rank * n

    def step(with_stats=False):
        # the timed launches are what generate() issues: no statistics buffer (the reference's generate computes none)
        comp.launch(n, seed=args.seed, first_event=first_event, out=out, stats=stats if with_stats else None, err=err,
                    stream=stream.cuda_stream)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
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
    total_ms = max_over_ranks(t_start.elapsed_time(t_end))
    kernel_ms = max_over_ranks(statistics.fmean(a.elapsed_time(b) for a, b in evs))
    clocks = sampler.summary(wall0, wall1)
    # weight statistics of the same events from the kernel's own epilogue (one more launch, outside the timed region)
    stats.zero_()
    step(with_stats=True)
    torch.cuda.synchronize()
    gstats = stats.clone()
    if clocks["samples"] < 3:  # timed region shorter than the sampling period: probe the same kernel a little longer
        p0 = time.perf_counter()
        while time.perf_counter() - p0 < 0.15:
            step()
            torch.cuda.synchronize()
        clocks = sampler.summary(p0, time.perf_counter())
        clocks["note"] = "timed region too short to sample; taken from an immediately following probe of the same launches"
    assert int(err.item()) == 0, "forbidden decay flagged"
    # the only collective of the path: the tiny weight-statistics allreduce (outside the timed loop)
    if world > 1:
        gstats = sharding.allreduce_stats(gstats)
    n_acc = float(world) * n
    global_stats = {"max_w": float(gstats[0]), "mean_w": float(gstats[1]) / n_acc,
                    "how": "kernel epilogue (warp shuffle -> shared -> one atomic per CTA) + two 16-byte all-reduces"}

    # ================================================================== all five BASELINE.json configs
    configs = {}
    if not args.no_configs:
        scale = n / 1e8
        # ---- config 1: B0 -> K pi, 1e6 events per call (README example; the reference benchmark's call size)
        n1 = 1_000_000
        d1 = ps.nbody_decay(5279.65, [PION_MASS, KAON_MASS])
        c1 = d1.compile()
        slabs = [{"weights": torch.empty((n1,), dtype=torch.float64, device=device),
                  "particles": torch.empty((2, n1, 4), dtype=torch.float64, device=device)} for _ in range(4)]
        # straight through the C ABI with the argument lists built beforehand (~3 us of host time per launch): the
        # Python-level CompiledDecay.launch costs 9-15 us per call depending on the host, which on a slow host is MORE
        # than this kernel lasts and would turn the figure into a host time
        raw_stream = stream.cuda_stream
        calls = [(c1.handle, args.seed, rank * n1, n1, None, 0, None, None, sl["weights"].data_ptr(), None,
                  sl["particles"].data_ptr(), 4 * n1, None, err.data_ptr(), _lib.PS_GEN_NORMALIZE, raw_stream, 1.0)
                 for sl in slabs]
        ps_generate = _lib.lib.ps_generate

        def launch1(k):
            rc = ps_generate(*calls[k % 4])
            if rc:
                _lib.check(rc)

        ms, clk = timed_launches(launch1, reps=2000, warmup=20)
        ev_s = world * n1 / (ms * 1e-3)
        entry = {"workload": "B0->K pi two-body nbody_decay(5279.65, [139.57018, 493.677]), n_events=1e6 per call, "
                             "normalize_weights=True", "events_per_gpu_per_launch": n1, "kernel_ms": ms,
                 "events_per_s": ev_s, **bounds(ev_s / world, 72, "flat_2body"), "clocks": clk,
                 "l2": "72 MB of outputs per launch fit in the 126 MB L2: launches rotate over four output slabs (288 MB)",
                 "note": "a 1e6-event launch writes 72 MB (10 us at full bandwidth); the rest is the gap between two kernels "
                         "of a stream and the ramp of the grid, which programmatic dependent launch (launches of at most "
                         "2^21 events) partly hides; launched straight through the C ABI (ps_generate with prebuilt argument "
                         "lists, ~3 us of host time per call) so that the figure is the kernel's, not the host's"}
        # the reference benchmark's protocol (benchmark/bench_phasespace.py:109-133): generate(1e6) per call through the
        # public API (allocation of the outputs included), 1 warm-up + 10 timed calls
        def do_run(k):  # benchmark/bench_phasespace.py:131-133: the result is returned and dropped
            return d1.generate(n1, seed=args.seed + k)

        do_run(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(10):
            do_run(k)
        torch.cuda.synchronize()
        api_ms = (time.perf_counter() - t0) / 10 * 1e3
        entry["api_generate_ms_per_call"] = max_over_ranks(api_ms)
        entry["api_events_per_s"] = world * n1 / (entry["api_generate_ms_per_call"] * 1e-3)
        entry["api_protocol"] = ("benchmark/bench_phasespace.py:109-133: 1 warm-up + 10 x generate(1_000_000) whose results are "
                                 "dropped, wall clock including the final synchronisation")
        t0 = time.perf_counter()
        for k in range(1000):
            do_run(k)
        torch.cuda.synchronize()
        entry["api_generate_ms_per_call_steady"] = max_over_ranks((time.perf_counter() - t0) / 1000 * 1e3)
        entry["api_events_per_s_steady"] = world * n1 / (entry["api_generate_ms_per_call_steady"] * 1e-3)
        del slabs
        configs["1_two_body_1e6"] = entry

        # ---- config 2: the headline above
        ev_s = world * n / (kernel_ms * 1e-3)
        configs["2_three_body_1e8"] = {"workload": WORKLOAD, "events_per_gpu_per_launch": n, "kernel_ms": kernel_ms,
                                       "events_per_s": ev_s, **bounds(ev_s / world, BYTES_PER_EVENT, "flat_3body"),
                                       "clocks": clocks, "note": "the headline kernel; sustained figure in `sustained`"}
        torch.cuda.empty_cache()

        # ---- config 3: B0 -> 4 pi with per-event boost_to (FONLL-like spectrum, synthesised on the device)
        n3 = int(1e8 * scale)
        g = torch.Generator(device=device).manual_seed(1234 + rank)
        pt = 1000.0 * torch.empty(n3, dtype=torch.float64, device=device).exponential_(1 / 5.0, generator=g)
        eta = 2 + 3 * torch.rand(n3, dtype=torch.float64, device=device, generator=g)
        phi = 6.283185307179586 * torch.rand(n3, dtype=torch.float64, device=device, generator=g)
        boost = torch.empty((n3, 4), dtype=torch.float64, device=device)
        boost[:, 0], boost[:, 1], boost[:, 2] = pt * torch.cos(phi), pt * torch.sin(phi), pt * torch.sinh(eta)
        boost[:, 3] = torch.sqrt((boost[:, :3] ** 2).sum(1) + B0_MASS**2)
        del pt, eta, phi
        d3 = ps.nbody_decay(B0_MASS, [PION_MASS] * 4)
        c3 = d3.compile()
        o3 = {"weights": torch.empty((n3,), dtype=torch.float64, device=device),
              "particles": torch.empty((4, n3, 4), dtype=torch.float64, device=device)}
        err.zero_()
        ms, clk = timed_launches(lambda k: c3.launch(n3, seed=args.seed, first_event=rank * n3, boost_to=boost, out=o3, err=err),
                                 reps=10)
        assert int(err.item()) == 0
        ev_s = world * n3 / (ms * 1e-3)
        configs["3_four_body_boost_1e8"] = {
            "workload": "B0->pi pi pi pi with per-event boost_to (n,4) [pT~Exp(5 GeV), eta~U(2,5), phi~U(0,2pi); "
                        "tests/helpers/rapidsim.py:46-51], normalize_weights=True", "events_per_gpu_per_launch": n3,
            "kernel_ms": ms, "events_per_s": ev_s, **bounds(ev_s / world, 8 + 32 * 4 + 32, "flat_4body_boost"), "clocks": clk,
            "note": "136 B written + 32 B read per event"}
        del o3
        torch.cuda.empty_cache()
        if not args.no_e2e:
            # the same config end to end with HOST buffers: boost_to rows go H2D and the events come back D2H, both
            # inside the timed region (generate_host: chunked, copies overlapped with generation)
            n3h = min(n3, 25_000_000)
            h_boost = torch.empty((n3h, 4), dtype=torch.float64, pin_memory=True)
            h_boost.copy_(boost[:n3h])
            h_out = {"weights": torch.empty((n3h,), dtype=torch.float64, pin_memory=True),
                     "particles": torch.empty((4, n3h, 4), dtype=torch.float64, pin_memory=True)}
            d3.generate_host(n3h, boost_to=h_boost, seed=args.seed, out=h_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                d3.generate_host(n3h, boost_to=h_boost, seed=args.seed, out=h_out)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            configs["3_four_body_boost_1e8"]["e2e_host"] = {
                "value": world * n3h * 2 / dt, "unit": UNIT, "events_per_gpu_per_step": n3h, "steps": 2,
                "h2d_bytes_per_step": 32 * n3h, "d2h_bytes_per_step": 136 * n3h,
                "gbs_all_ranks_both_directions": world * 168 * n3h * 2 / dt / 1e9,
                "api": "GenParticle.generate_host(n, boost_to=<pinned host rows>) -> ps_generate_host"}
            del h_boost, h_out
        del boost
        torch.cuda.empty_cache()

        # ---- config 4: B0 -> K*0(-> K pi) gamma, relativistic Breit-Wigner K*0 sampled in the kernel
        def kstar_chain():
            return ps.GenParticle("B0", B0_MASS).set_children(
                ps.GenParticle("K*0", mf.relativistic_breitwigner_factory(KSTAR_MASS, KSTAR_WIDTH)).set_children(
                    ps.GenParticle("K+", KAON_MASS), ps.GenParticle("pi-", PION_MASS)),
                ps.GenParticle("gamma", 0.0))

        n4 = int(1e8 * scale)
        d4 = kstar_chain()
        c4 = d4.compile()
        o4 = {"weights": torch.empty((n4,), dtype=torch.float64, device=device),
              "particles": torch.empty((4, n4, 4), dtype=torch.float64, device=device)}
        err.zero_()
        ms, clk = timed_launches(lambda k: c4.launch(n4, seed=args.seed, first_event=rank * n4, out=o4, err=err), reps=10)
        assert int(err.item()) == 0
        ev_s = world * n4 / (ms * 1e-3)
        entry = {"workload": "B0->K*0(->K+ pi-) gamma, K*0 ~ relativistic Breit-Wigner(895.81, 47.4) sampled in the "
                             "kernel (set_children chain), normalize_weights=True", "events_per_gpu_per_launch": n4,
                 "kernel_ms": ms, "events_per_s": ev_s, **bounds(ev_s / world, 8 + 32 * 4, "kstar_gamma_relbw"), "clocks": clk}
        del o4
        torch.cuda.empty_cache()
        multi = GenMultiDecay([(1.0, kstar_chain())])
        for _ in range(2):
            w_, e_ = multi.generate(n4, seed=args.seed)
            del w_, e_
        barrier()
        t0 = time.perf_counter()
        for k in range(5):
            w_, e_ = multi.generate(n4, seed=args.seed + k)
            del w_, e_
        torch.cuda.synchronize()
        api_ms = max_over_ranks((time.perf_counter() - t0) / 5 * 1e3)
        entry["genmultidecay_ms_per_call"] = api_ms
        entry["genmultidecay_events_per_s"] = world * n4 / (api_ms * 1e-3)
        entry["genmultidecay_api"] = ("GenMultiDecay([(1.0, chain)]).generate(n): allocates its 13.6 GB of outputs and "
                                      "synchronises once for the forbidden-decay check, wall clock")
        configs["4_kstar_gamma_relbw_1e8"] = entry
        torch.cuda.empty_cache()

        # ---- config 5: 10-body flat decay, accept-reject unweighting, 1.25e9 candidates per GPU (1e10 on 8 GPUs)
        d5 = ps.nbody_decay(B0_MASS, [PION_MASS] * 10)
        c5 = d5.compile()
        n5s = int(5e7 * scale)  # every event stored: the fp64-bound kernel on its own
        o5 = {"weights": torch.empty((n5s,), dtype=torch.float64, device=device),
              "particles": torch.empty((10, n5s, 4), dtype=torch.float64, device=device)}
        ms, clk = timed_launches(lambda k: c5.launch(n5s, seed=args.seed, first_event=rank * n5s, out=o5, err=err), reps=10)
        ev_s = world * n5s / (ms * 1e-3)
        stored = {"workload": "10-body flat decay, every event stored (the generator kernel on its own)",
                  "events_per_gpu_per_launch": n5s, "kernel_ms": ms, "events_per_s": ev_s,
                  **bounds(ev_s / world, 8 + 320, "flat_10body"), "clocks": clk}
        del o5
        torch.cuda.empty_cache()
        per_gpu = int(1.25e9 * scale)
        chunk = 1 << 27
        # acceptance bound: 1.5 x the global maximum of the normalised weight over a pilot sample of 4e6 events per
        # rank (one 16-byte MAX all-reduce); candidates above the bound are COUNTED by the mask pass (overflow)
        pstats = torch.zeros(4, dtype=torch.float64, device=device)
        c5.launch(4_000_000, args.seed + 1, first_event=rank * 4_000_000, stats=pstats)
        w_bound = 1.5 * float(sharding.allreduce_stats(pstats)[0])
        first5 = rank * per_gpu

        def unweight_pass():
            kept = over = done = 0
            while done < per_gpu:
                m = min(chunk, per_gpu - done)
                parts, idx, w, n_over = generate_unweighted(d5, m, args.seed, w_bound=w_bound, first_event=first5 + done,
                                                            return_overflow=True)
                kept, over, done = kept + int(idx.numel()), over + n_over, done + m
                del parts, idx, w
            return kept, over

        generate_unweighted(d5, min(chunk, per_gpu), args.seed, w_bound=w_bound, first_event=first5, return_overflow=True)
        barrier()
        time.sleep(1.0)
        a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        l0 = _lib.lib.ps_launch_count()
        a5.record(stream)
        kept, over = unweight_pass()
        b5.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        l5 = _lib.lib.ps_launch_count() - l0
        sec = max_over_ranks(a5.elapsed_time(b5) * 1e-3)
        kept_all, over_all = sum_over_ranks(kept), sum_over_ranks(over)  # the path's only collective: two counters
        cand = world * per_gpu
        configs["5_ten_body_unweighting"] = {
            "workload": f"10-body flat decay, accept-reject unweighting: {per_gpu:.4g} candidates per GPU "
                        f"(= 1e10 over 8 GPUs) in chunks of 2^27 through unweight.generate_unweighted (mask pass: weights "
                        "only -> scan -> index -> regenerate the accepted events); only accepted events are stored",
            "candidates_per_gpu": per_gpu, "candidates": cand, "seconds": sec, "candidates_per_s": cand / sec,
            "accepted": int(kept_all), "accepted_per_s": kept_all / sec, "acceptance": kept_all / cand,
            "w_bound": w_bound, "overflow_candidates": int(over_all),
            "overflow_note": "candidates with w / w_max > w_bound (accepted with probability 1 instead of > 1); must be 0 "
                             "for an unbiased sample",
            "bytes_stored_per_accepted_event": 8 + 8 + 320, "gpu_launches": int(l5),
            "timing": "CUDA events around the whole chunk loop (host read-back of the accepted count per chunk included), "
                      "max over ranks", "clocks": sampler.summary(w0, w1), "all_events_stored": stored}
        # (reported, not asserted: a bench line with a non-zero count says the bound must be raised, it must not get lost)
        configs["5_ten_body_unweighting"]["unbiased"] = bool(over_all == 0)
        torch.cuda.empty_cache()

    # ================================================================== sustained: the same launches for >= 2 s
    sustained = None
    if not args.no_sustained:
        barrier()
        time.sleep(0.2)
        batch = max(10, int(0.05 / max(kernel_ms * 1e-3, 1e-6)))  # ~50 ms of launches between two host checks
        marks = [torch.cuda.Event(enable_timing=True)]
        w0 = time.perf_counter()
        marks[0].record(stream)
        done = 0
        while time.perf_counter() - w0 < args.sustained_seconds:
            for _ in range(batch):
                step()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append(ev)
            done += batch
            ev.synchronize()
        w1 = time.perf_counter()
        total = marks[0].elapsed_time(marks[-1])
        q = max(1, (len(marks) - 1) // 4)
        last_q = marks[-1 - q].elapsed_time(marks[-1]) / (q * batch)
        sus_ms = max_over_ranks(total / done)
        sus_last = max_over_ranks(last_q)
        sus_clocks = sampler.summary(w0 + 0.5 * (w1 - w0), w1)  # second half: the power cap has settled by then
        sustained = {
            "what": "the headline launches issued back to back (host check every ~50 ms) until the wall clock passes "
                    "sustained_seconds; the board reaches its power cap after ~0.3 s, so this is what a long job sees",
            "seconds": total * 1e-3, "steps": done, "ms_per_step": sus_ms, "events_per_s": world * n / (sus_ms * 1e-3),
            "achieved_gbs": BYTES_PER_EVENT * n / (sus_ms * 1e-3) / 1e9, "frac": BYTES_PER_EVENT * n / (sus_ms * 1e-3) / 1e9 / peak,
            "ms_per_step_last_quarter": sus_last, "frac_last_quarter": BYTES_PER_EVENT * n / (sus_last * 1e-3) / 1e9 / peak,
            "clocks_second_half": sus_clocks,
        }
        if peaks and "bf16_tflops_sustained" in peaks:
            sustained["note"] = ("peak is the driver's burst copy figure; the driver's own sustained/burst ratio for bf16 "
                                 f"matmul on this pool is {peaks['bf16_tflops_sustained'] / peaks['bf16_tflops']:.2f}")
    del out
    torch.cuda.empty_cache()

    # ================================================================== e2e: public host-delivery call
    e2e = None
    d2h_probe = None
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
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * e2e_n * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": BYTES_PER_EVENT * e2e_n, "steps": e2e_steps, "events_per_gpu_per_step": e2e_n,
               "d2h_gbs_all_ranks": world * BYTES_PER_EVENT * e2e_n * e2e_steps / dt / 1e9,
               "api": "GenParticle.generate_host -> ps_generate_host (chunked generate + async D2H into pinned host memory)",
               "inputs": "config 2 has no inputs (H2D = 0); the step's whole result (10.4 GB per GPU) is delivered to the host",
               "checksum_w0": float(w_host[0])}
        if world > 1:
            e2e["host_buffers"] = (f"each rank bound to the {numa_cpus} CPUs NVML reports as local to its GPU before "
                                   "allocating" if numa_cpus else "no CPU-affinity information from NVML")
        # the ceiling of e2e: raw D2H copies of the same size from every rank at the same time, no kernel involved
        probe_bytes = min(BYTES_PER_EVENT * e2e_n, 2 << 30)
        src = torch.empty((probe_bytes // 8,), dtype=torch.float64, device=device)
        dst = host["particles"].view(-1)[: probe_bytes // 8]
        dst.copy_(src, non_blocking=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        b.record(stream)
        torch.cuda.synchronize()
        probe_s = max_over_ranks(a.elapsed_time(b) * 1e-3)
        d2h_probe = {"what": "cudaMemcpyAsync device -> pinned host, 4 copies per rank, all ranks concurrently, CUDA events, "
                             "max over ranks: the ceiling of e2e on this host",
                     "bytes_per_copy": probe_bytes, "gbs_all_ranks": world * 4 * probe_bytes / probe_s / 1e9,
                     "gbs_per_rank": 4 * probe_bytes / probe_s / 1e9,
                     "e2e_fraction_of_probe": e2e["d2h_gbs_all_ranks"] / (world * 4 * probe_bytes / probe_s / 1e9)}
        del src, dst, host, w_host, _parts
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
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_device = {"value": world * n * dev_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 0,
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
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_hist = {"value": world * n * hist_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 8 * 100 * (4 * N_BODY + 1), "steps": hist_steps,
                    "sum_weight_histogram": float(hw_host.sum()), "particles": len(hp_host),
                    "api": "histogram.generate_histograms -> ps_histogram (events binned in shared memory, never stored) "
                           "+ the normalised histograms copied to the host"}
    sampler.stop()

    if rank == 0:
        traffic, traffic_src = None, None
        t = _load_json("profiles", "traffic.json")
        if t and int(t.get("events_per_launch", -1)) == n:
            traffic = t["dram_bytes_per_launch"]
            traffic_src = "STORED ncu capture (profiles/traffic.json: " + str(t.get("source", "")) + "), not measured in this run"
        achieved = BYTES_PER_EVENT * n / (kernel_ms * 1e-3) / 1e9
        fp64 = None
        if fp64_peak and fp64_counts.get("flat_3body"):
            bound = fp64_peak / fp64_counts["flat_3body"]
            fp64 = {"fp64_inst_per_event": fp64_counts["flat_3body"], "peak_fp64_inst_per_s": fp64_peak,
                    "bound_events_per_s": bound, "frac": n / (kernel_ms * 1e-3) / bound,
                    "note": "HBM is the binding roofline for the 3-body decay; fp64 pipe peak measured with tools/fp64_bench.cu"}
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_EVENT * n, "kernel_ms": kernel_ms, "fp64_bound": fp64,
                         "burst": f"timed region of {total_ms:.0f} ms: a burst figure, see `sustained` for >= 2 s"},
            "clocks": clocks, "gpu_launches": int(launches), "global_stats": global_stats,
        }
        if sustained is not None:
            line["sustained"] = sustained
        if configs:
            line["configs"] = configs
        if e2e is not None:
            line["e2e"] = e2e
        if d2h_probe is not None:
            line["d2h_probe"] = d2h_probe
        if e2e_device is not None:
            line["e2e_device_consumer"] = e2e_device
        if e2e_hist is not None:
            line["e2e_histogram_consumer"] = e2e_hist
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        line["reference_note"] = ("the reference arm (--impl reference) and cpu_baseline are the numpy PORT of the reference, "
                                  "not its TensorFlow CPU path, which cannot be installed here")
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
    ap.add_argument("--sustained-seconds", type=float, default=2.5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
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
