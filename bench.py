#!/usr/bin/env python
"""bench.py — footprint-pose checks/s on BASELINE.json config 4 (DWA lattice), plus the other configs as `secondary`.

Headline workload (one "step" = one planning cycle = ingest the costmap, sweep the whole lattice): 4000 x 4000 uint8
costmap at 5 cm (sparse obstacles inflated like costmap_2d's InflationLayer), 1.0 x 0.6 m rectangular footprint,
(x, y, theta) lattice 1180 x 1180 x 72 = 100 252 800 poses, verdict only (first-lethal) — the batched
`is_free(polygon, pose, map)` of src/cover/costmap.cpp:78-85.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-secondary] [--no-cpu-baseline]

N > 1 is launched by torchrun (one rank per GPU): poses are sharded, the costmap is replicated, no data-path collective
(weak scaling on C4: every rank sweeps its own full lattice window; the 8-shape C5 sweep is additionally timed with its
poses split across the ranks = strong scaling).  The end-to-end arm replicates the costmap the cheap way at N > 1: rank 0
uploads it from pinned host memory, an NCCL broadcast fans it out over NVLink.  Prints ONE JSON line (rank 0).
DESIGN.md §6 says how each field is obtained.

Environment: COVER_BENCH_FLIGHT (planning cycles in flight in the e2e arm, default 2), COVER_BENCH_THREADS=1 (N = 1 e2e
arm with one host thread per cycle and blocking calls instead of one thread and COVER_MEM_HOST_ASYNC results), COVER_BENCH_BCAST=0 (N > 1: every
rank uploads its own copy of the costmap instead of the broadcast), COVER_BENCH_WINDOW (time another rank's lattice window
on one GPU), COVER_BENCH_NO_PIN (no CPU pinning of the ranks).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cover_b200 import workloads as wl  # noqa: E402

METRIC = "footprint_pose_checks_per_second"
UNIT = "checks/s"
MAP_SIZE = 4000
RESOLUTION = 0.05
CPU_CHUNK = 250_000  # poses per contiguous chunk of the CPU arm's sample


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "C4 DWA lattice: 4000x4000 inflated costmap @0.05 m, 1.0x0.6 m rect footprint, "
                    "1180x1180x72 (x,y,theta) lattice = 100252800 poses per GPU, is_free (first-lethal)",
        "map": [MAP_SIZE, MAP_SIZE], "resolution_m": RESOLUTION, "footprint": "rect 1.0x0.6 m",
        "poses_per_gpu": 1180 * 1180 * 72, "pose_format": "lattice", "op": "area_is_free",
        "outputs": "bit-packed verdicts (12.5 MB) + not-OK counter",
        "parallelism": f"pose-sharded x{n_gpus}, costmap replicated, no collective on the data path",
        "l2_flush_between_steps": True,
    }


def lattice_for_rank(rank: int) -> dict:
    # every rank gets its own window of the map (distinct poses), all >= 1 m from the map edge
    shift = 2.0 * (rank % 8)
    return wl.dwa_lattice(x0=20.0 + shift, y0=20.0 + 3.0 * (rank % 8))


def make_map() -> np.ndarray:
    return wl.inflated_costmap(MAP_SIZE, MAP_SIZE, RESOLUTION, seed=3)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources (oracle/_ref) when they were compiled, else the restatement
# ------------------------------------------------------------------------------------------------

def cpu_backend():
    from oracle import pyoracle as po  # the CPU arm is the one place bench.py may execute oracle/
    ref = po.reference()
    return (ref, "reference") if ref is not None else (po.default(), "port")


def cpu_sample(data: np.ndarray, desc: dict, n_sample: int, threads: int, seed: int = 11, tally: bool = True):
    """Runs the CPU implementation on `n_sample` poses of the SAME lattice (random contiguous chunks).  Returns
    (checks/s, seconds, cells read per pose from the restatement's tally or None, n, kind, [(first, verdicts, statuses)])."""
    from oracle import pyoracle as po
    backend, kind = cpu_backend()
    omap = po.Costmap(data, RESOLUTION)
    olat = po.Lattice(**desc)
    rng = np.random.default_rng(seed)
    chunk = min(CPU_CHUNK, n_sample)
    starts = np.sort(rng.integers(0, olat.count - chunk, max(1, n_sample // chunk)))
    kept = []
    t0 = time.perf_counter()
    for s in starts.tolist():
        r = backend.polygon_batch(omap, wl.RECT, olat, first=s, count=chunk, threads=threads)
        kept.append((s, r.is_free, r.status))
    dt = time.perf_counter() - t0
    n = len(starts) * chunk
    cells = None
    if tally:  # algorithmic bytes: cells the reference's sequential loop reads (the restatement counts them)
        probe = min(chunk, 100_000)
        cells = po.default().polygon_batch(omap, wl.RECT, olat, first=int(starts[0]), count=probe, threads=threads).cells_read / probe
    return n / dt, dt, cells, n, kind, kept


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    threads = po.hardware_threads()
    data = make_map()
    desc = lattice_for_rank(0)
    per_step = 2_000_000  # bounded sample of the workload per step
    for _ in range(args.warmup):
        cpu_sample(data, desc, CPU_CHUNK, threads, tally=False)
    t_total = 0.0
    kind = "port"
    for k in range(args.steps):
        _, dt, _, n, kind, _ = cpu_sample(data, desc, per_step, threads, seed=100 + k, tally=False)
        t_total += dt
    value = per_step * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{per_step} poses of the same lattice per step ({CPU_CHUNK}-pose chunks at random offsets), {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(", ") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, flag in zip(names, r[5:9]):
                    if flag.strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_counters() -> dict:
    """Counters of the committed ncu capture of the dominant kernel (profiles/roofline_traffic.json): NOT measured in
    this run - the file names the commit they were taken at."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


def device_timed(ctx, stream, fn, reps, warm=2, flush=None):
    """Average device time (ms) of fn() over `reps` calls: one CUDA event pair per call on the library's stream."""
    import torch
    for _ in range(warm):
        fn()
    ctx.synchronize()
    pairs = []
    for k in range(reps):
        with torch.cuda.stream(stream):
            if flush is not None:
                flush.fill_(k & 0xFF)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            fn()
            e.record(stream)
        pairs.append((s, e))
    ctx.synchronize()
    return sum(s.elapsed_time(e) for s, e in pairs) / reps


def compare(gpu_free, gpu_status, cpu_free, cpu_status) -> int:
    """Mismatching poses: verdicts wherever both sides finished, and the std::logic_error flag everywhere."""
    gpu_free, gpu_status = np.asarray(gpu_free), np.asarray(gpu_status)
    logic = (gpu_status == 1) != (np.asarray(cpu_status) == 1)
    ok = (gpu_status == 0) & (np.asarray(cpu_status) == 0)
    return int(logic.sum() + (gpu_free[ok] != np.asarray(cpu_free)[ok]).sum())


def secondary_configs(ctx, stream, peak, threads, with_cpu):
    """The other BASELINE.json configs on this GPU, device-resident inputs and outputs, device-timed; every line carries
    the algorithmic bytes per check of SURVEY 8(d), what that makes of the HBM roof, and a parity count against the CPU arm."""
    import torch

    import cover_b200
    from tests.fixtures import FOOTPRINTS
    rows = []
    if with_cpu:
        from oracle import pyoracle as po
        backend, kind = cpu_backend()

    def dev_result(n, cost=False):
        return cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"),
                                 torch.empty(n, dtype=torch.float64, device="cuda") if cost else None,
                                 torch.empty(n, dtype=torch.uint8, device="cuda"))

    def report(name, n, ms, bytes_per_check, parity=None, **extra):
        row = {"config": name, "poses": n, "ms": ms, "checks_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_check": bytes_per_check}
        if bytes_per_check is not None:
            row["algorithmic_gbs"] = bytes_per_check * n / (ms * 1e-3) / 1e9
            row["algorithmic_frac_of_hbm"] = row["algorithmic_gbs"] / peak
        if parity is not None:
            row["parity"] = parity
        row.update(extra)
        rows.append(row)

    def polygon_parity(data, origin, poly, poses_np, gpu_out, m, shape=2, tally=True):
        """(parity dict, cells read per pose) of the first m poses against the CPU arm."""
        if not with_cpu:
            return None, None
        omap = po.Costmap(data, RESOLUTION, *origin)
        r = backend.polygon_batch(omap, poly, poses_np[:m], shape=shape, threads=threads)
        mism = compare(gpu_out.is_free[:m].cpu().numpy(), gpu_out.status[:m].cpu().numpy(), r.is_free, r.status)
        cells = po.default().polygon_batch(omap, poly, poses_np[:min(m, 20000)], shape=shape, threads=threads).cells_read / min(m, 20000) if tally else None
        return {"poses": m, "mismatches": mism, "against": kind}, cells

    # ---- C4 with byte outputs (verdict byte + status byte per pose: 200 MB of stores per sweep)
    data4 = make_map()
    cmap4 = cover_b200.Costmap(ctx, data4, RESOLUTION)
    d4 = lattice_for_rank(0)
    n4 = d4["nx"] * d4["ny"] * d4["cs"].shape[0]
    lat4 = cover_b200.Lattice(**d4)
    out4 = dev_result(n4)
    map_dev = torch.from_numpy(data4).cuda()

    def step_bytes():
        cmap4.upload(map_dev)
        cmap4.area_is_free(wl.RECT, lat4, out=out4)
    ms = device_timed(ctx, stream, step_bytes, 5)
    report("C4 lattice, byte verdict + status outputs (200 MB of stores), map re-ingested per step", n4, ms, None,
           hbm_bytes_moved=2 * n4 + 2 * data4.size, hbm_frac=(2 * n4 + 2 * data4.size) / (ms * 1e-3) / 1e9 / peak)
    # other folds / generators on the same lattice
    for name, fn in (("C4 lattice, outline is_free", lambda: cmap4.outline_is_free(wl.RECT, lat4, out=out4)),):
        report(name, n4, device_timed(ctx, stream, fn, 3), None)
    n4s = 1180 * 1180 * 18
    lat4s = cover_b200.Lattice(**dict(d4, cs=d4["cs"][::4]))
    out4c = dev_result(n4s, cost=True)
    report("C4 lattice (18 headings), area accumulate_cost", n4s, device_timed(ctx, stream, lambda: cmap4.area_accumulate_cost(wl.RECT, lat4s, out=out4c), 3), None)
    report("C4 lattice (18 headings), area max_cost", n4s, device_timed(ctx, stream, lambda: cmap4.area_max_cost(wl.RECT, lat4s, out=out4c), 3), None)
    del out4, out4c, cmap4, map_dev

    # ---- C1: 1000^2 sparse-lethal map, rect, random poses (1e6 instead of 1e5 so that the launch is not latency-bound)
    n = 1_000_000
    data = wl.random_costmap(1000, 1000, seed=1, lethal_p=1e-3)
    poses_np = wl.random_poses(n, seed=2, lo=1.0, hi=49.0)
    poses = torch.from_numpy(poses_np).cuda()
    out = dev_result(n, cost=True)
    cmap = cover_b200.Costmap(ctx, data, RESOLUTION)
    ms = device_timed(ctx, stream, lambda: cmap.area_is_free(wl.RECT, poses, out=out), 5)
    par, cells = polygon_parity(data, (0.0, 0.0), wl.RECT, poses_np, out, 100_000)
    report("C1 rect area is_free, 1e6 random poses, 1000^2 sparse-lethal map", n, ms, None if cells is None else cells + 34, par)
    ms = device_timed(ctx, stream, lambda: cmap.area_accumulate_cost(wl.RECT, poses, out=out), 5)
    report("C1 rect area accumulate_cost", n, ms, None if cells is None else cells + 42)
    # the same check by the byte-scanning kernels (no bit planes): what the kernel moves ~ the algorithmic bytes
    os.environ["COVER_B200_NO_PLANES"] = "1"
    ctx_b = cover_b200.Context(ctx.device)
    os.environ.pop("COVER_B200_NO_PLANES")
    cmap_b = cover_b200.Costmap(ctx_b, data, RESOLUTION)
    stream_b = torch.cuda.ExternalStream(ctx_b.stream)
    ms = device_timed(ctx_b, stream_b, lambda: cmap_b.area_is_free(wl.RECT, poses, out=out), 5)
    par_b, _ = polygon_parity(data, (0.0, 0.0), wl.RECT, poses_np, out, 100_000, tally=False)
    report("C1 rect area is_free, byte scans (COVER_B200_NO_PLANES=1: k_area_pair reads the cost bytes)", n, ms, None if cells is None else cells + 34, par_b)
    del cmap_b
    ctx_b.close()

    # ---- C2: 2000^2 inflated map, outlines + rays, 1e6 poses
    n = 1_000_000
    data = wl.inflated_costmap(2000, 2000, RESOLUTION, seed=3)
    cmap = cover_b200.Costmap(ctx, data, RESOLUTION)
    poses_np = wl.random_poses(n, seed=4, lo=1.0, hi=99.0)
    poses = torch.from_numpy(poses_np).cuda()
    out = dev_result(n, cost=True)
    for name, poly in (("rect", wl.RECT), ("TORU", FOOTPRINTS[5])):
        ms = device_timed(ctx, stream, lambda: cmap.outline_is_free(poly, poses, out=out), 5)
        par, cells = polygon_parity(data, (0.0, 0.0), poly, poses_np, out, 100_000, shape=1)
        report(f"C2 {name} outline is_free, 1e6 random poses, 2000^2 inflated map", n, ms, None if cells is None else cells + 34, par)
    rng = np.random.default_rng(5)
    b = rng.integers(30, 1970, (n, 2))
    segs_np = np.concatenate([b, b + rng.integers(-25, 26, (n, 2))], 1).astype(np.int32)
    segs = torch.from_numpy(segs_np).cuda()
    ms = device_timed(ctx, stream, lambda: cmap.ray_is_free(segs, out=out), 5)
    par = None
    cells = None
    if with_cpu:
        r = backend.rays_batch(po.Costmap(data, RESOLUTION), segs_np[:100_000], threads=threads)
        par = {"poses": 100_000, "mismatches": compare(out.is_free[:100_000].cpu().numpy(), out.status[:100_000].cpu().numpy(), r.is_free, r.status), "against": kind}
        cells = po.default().rays_batch(po.Costmap(data, RESOLUTION), segs_np[:20000], threads=threads).cells_read / 20000
    report("C2 rays (<= 25 cells) is_free, 1e6 segments", n, ms, None if cells is None else cells + 18, par)

    # ---- C3: 4000^2 inflated map, 72-heading potato footprint bank x 1e7 offsets; split footprint, pose form
    n = 10_000_000
    cmap = cover_b200.Costmap(ctx, data4, RESOLUTION)
    pot = wl.potato(48)
    outlines, areas = [wl.scaled(pot, 0.45)], [pot]  # a hand-made split: inner skeleton outline + full remainder scan
    th = -math.pi + 2 * math.pi * np.arange(72) / 72
    head = np.stack([np.cos(th), np.sin(th), np.zeros(72), np.zeros(72)], 1)
    zero_map = cover_b200.Costmap(ctx, np.zeros((8, 8), dtype=np.uint8), RESOLUTION)  # expand only needs resolution and origin
    oc, os_, _ = zero_map.outline_cells(outlines[0], head)
    ac, as_, _ = zero_map.area_cells(areas[0], head)
    bank = [(oc[os_[k]:os_[k + 1]], ac[as_[k]:as_[k + 1]]) for k in range(72)]  # to_outline / to_area per heading, on the device
    fb = cover_b200.FootprintBank(ctx, bank)
    rng = np.random.default_rng(5)
    offs_np = rng.integers(40, 3960, (n, 2)).astype(np.int32)
    which_np = rng.integers(0, 72, n).astype(np.int32)
    offs, which = torch.from_numpy(offs_np).cuda(), torch.from_numpy(which_np).cuda()
    out = dev_result(n)
    ms = device_timed(ctx, stream, lambda: fb.is_free(cmap, offs, which, out=out), 3)
    par = None
    if with_cpu:
        r = backend.discrete_footprint_batch(po.Costmap(data4, RESOLUTION), bank, offs_np[:200_000], which_np[:200_000], threads=threads)
        par = {"poses": 200_000, "mismatches": compare(out.is_free[:200_000].cpu().numpy(), out.status[:200_000].cpu().numpy(), r.is_free, r.status), "against": kind}
    cells_fp = float(np.mean([len(o) + len(a) for o, a in bank]))
    report("C3 potato discrete_footprint bank (72 headings) is_free, 1e7 offsets, 4000^2 inflated map", n, ms, cells_fp + 10, par, cells_per_footprint=cells_fp)
    sf = cover_b200.SplitFootprint(ctx, outlines, areas)
    n2 = 1_000_000
    poses_np = wl.random_poses(n2, seed=6, lo=2.0, hi=198.0)
    poses = torch.from_numpy(poses_np).cuda()
    out2 = dev_result(n2)
    ms = device_timed(ctx, stream, lambda: sf.is_free(cmap, poses, out=out2), 3)
    par = None
    if with_cpu:
        r = backend.continuous_footprint_batch(po.Costmap(data4, RESOLUTION), outlines, areas, poses_np[:50_000], threads=threads)
        par = {"poses": 50_000, "mismatches": compare(out2.is_free[:50_000].cpu().numpy(), out2.status[:50_000].cpu().numpy(), r.is_free, r.status), "against": kind}
    report("C3 potato continuous_footprint is_free (re-rasterised per pose), 1e6 poses", n2, ms, None, par)
    del cmap, fb, sf, out, out2, offs, which, data4

    rows.extend(c5_sweep(ctx, stream, peak, threads, with_cpu, 0, 1))
    return rows


def c5_shapes():
    from tests.fixtures import FOOTPRINTS
    return [("rect(4)", wl.RECT), ("L(6)", FOOTPRINTS[6]), ("S(6)", FOOTPRINTS[0]), ("W(8)", FOOTPRINTS[1]), ("M(10)", FOOTPRINTS[2]),
            ("SOTO(22)", FOOTPRINTS[4]), ("TORU(32)", FOOTPRINTS[5]), ("potato(64)", wl.potato(64))]


def c5_sweep(ctx, stream, peak, threads, with_cpu, rank, world, n_total=10_000_000):
    """C5: 8 polygon shapes x 1e7 random poses each on an 8000^2 inflated map; this rank checks its contiguous block of
    every shape's poses (strong scaling: the work is fixed, the ranks split it)."""
    import torch

    import cover_b200
    from cover_b200 import sharding
    rows = []
    data = wl.inflated_costmap(8000, 8000, RESOLUTION, seed=7)
    cmap = cover_b200.Costmap(ctx, data, RESOLUTION)
    first, n = sharding.partition(n_total, world, rank)
    poses_np = wl.random_poses(n_total, seed=8, lo=8.0, hi=392.0)[first:first + n]
    poses = torch.from_numpy(np.ascontiguousarray(poses_np)).cuda()
    out = cover_b200.Result(torch.empty(n, dtype=torch.uint8, device="cuda"), None, torch.empty(n, dtype=torch.uint8, device="cuda"))
    if with_cpu:
        from oracle import pyoracle as po
        backend, kind = cpu_backend()
        omap = po.Costmap(data, RESOLUTION)
    for name, poly in c5_shapes():
        ms = device_timed(ctx, stream, lambda: cmap.area_is_free(poly, poses, out=out), 2, warm=1)
        row = {"config": f"C5 {name} area is_free, random poses, 8000^2 inflated map", "poses": n, "ms": ms, "checks_per_s": n / (ms * 1e-3),
               "not_ok": int((out.status != 0).sum())}
        if with_cpu:
            m = 20_000
            r = backend.polygon_batch(omap, poly, poses_np[:m], threads=threads)
            row["parity"] = {"poses": m, "mismatches": compare(out.is_free[:m].cpu().numpy(), out.status[:m].cpu().numpy(), r.is_free, r.status), "against": kind}
            cells = po.default().polygon_batch(omap, poly, poses_np[:5000], threads=threads).cells_read / 5000
            row["algorithmic_bytes_per_check"] = cells + 34
            row["algorithmic_gbs"] = (cells + 34) * n / (ms * 1e-3) / 1e9
            row["algorithmic_frac_of_hbm"] = row["algorithmic_gbs"] / peak
        rows.append(row)
    return rows


def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    import cover_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — cover_b200 has no CPU path to time (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # one process per GPU on a shared host: give every rank its own cores and keep torch's CPU pools out of the
        # way - a 0.15 ms step is short enough for the launching thread's scheduling jitter to show in the device time
        torch.set_num_threads(1)
        try:
            if os.environ.get("COVER_BENCH_NO_PIN"):
                raise OSError
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    data = make_map()
    window = int(os.environ.get("COVER_BENCH_WINDOW", rank))  # (override: time another rank's window on one GPU)
    desc = lattice_for_rank(window)
    n_poses = desc["nx"] * desc["ny"] * desc["cs"].shape[0]
    n_words = (n_poses + 31) // 32

    ctx = cover_b200.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    lat = cover_b200.Lattice(**desc)

    # ---- device-resident arm (`value`): map in HBM, lattice descriptor, bit-packed verdicts + not-OK counter written to HBM
    cmap = cover_b200.Costmap(ctx, data, RESOLUTION)
    out = cover_b200.Result(free_bits=torch.zeros(n_words, dtype=torch.int32, device="cuda"), not_ok=torch.zeros(1, dtype=torch.int64, device="cuda"))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    map_dev = torch.from_numpy(data).cuda()  # the costmap of the timed steps is already in HBM

    def step():
        # one planning cycle: ingest the (device-resident) costmap - which invalidates the lethal bit planes the sweep
        # reads, so they are rebuilt: nothing derived from the map's cells is carried over between steps - and sweep
        cmap.upload(map_dev)
        cmap.area_is_free(wl.RECT, lat, out=out)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    ctx.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:  # (one nvidia-smi poller per box: eight of them contend with the ranks' launch threads for the driver)
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the event pair)
            starts[k].record(stream)
            step()
            ends[k].record(stream)
    ctx.synchronize()
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    t_ms = sum(step_ms)
    t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms_max = float(t.item())
    ms_per_step = t_ms_max / args.steps
    value = world * n_poses * args.steps / (t_ms_max * 1e-3)
    assert int(out.not_ok.item()) == 0
    dev_free = np.unpackbits(out.free_bits.cpu().numpy().view(np.uint8), bitorder="little")[:n_poses]
    free_fraction = float(dev_free.mean())
    # the sweep alone (map and planes untouched since the last step): prepare + sweep + drain kernels
    sweep_ms = device_timed(ctx, stream, lambda: cmap.area_is_free(wl.RECT, lat, out=out), 10, flush=flush)

    # ---- end-to-end arm (`e2e`): host buffers through the public API, copies inside the timed region.  Two planning
    # cycles are kept in flight (two contexts = two streams, one host thread each): one cycle's results travel back
    # while the other's costmap travels in.
    n_flight = int(os.environ.get("COVER_BENCH_FLIGHT", "2"))
    e2e_steps = max(50, args.steps)
    e2e_steps += e2e_steps % n_flight
    workers = []
    for _ in range(n_flight):
        c = cover_b200.Context(local)
        m = cover_b200.Costmap(c, data, RESOLUTION)
        pinned = torch.from_numpy(data).pin_memory()
        hb = torch.empty(n_words, dtype=torch.int32).pin_memory()
        hn = torch.empty(1, dtype=torch.int64).pin_memory()
        workers.append({"ctx": c, "map": m, "image": pinned.numpy(), "pinned": pinned, "bits": hb, "not_ok": hn,
                        "out": cover_b200.Result(free_bits=hb.numpy().view(np.uint32), not_ok=hn.numpy())})

    def e2e_step(w, image=None):
        # H2D: this planning cycle's costmap (16 MB), queued as row bands; the sweep starts on the first bands while the
        # rest is in flight (cover_map_upload_async).  Lattice descriptor in; bit-packed verdicts (12.5 MB) + the not-OK
        # counter out (D2H); returns when they are on the host.
        w["map"].upload_async(w["image"] if image is None else image)
        w["map"].area_is_free(wl.RECT, lat, out=w["out"])

    # warm-up, and a check that every step really sees ITS upload: an empty map first (everything free) ...
    empty_map = torch.zeros(MAP_SIZE, MAP_SIZE, dtype=torch.uint8).pin_memory().numpy()
    for w in workers:
        e2e_step(w, empty_map)
        assert bool(w["out"].unpack_bits(n_poses).all()), "e2e: the sweep did not see the freshly uploaded (empty) costmap"
        for _ in range(2):  # ... then the real one (compared with the device-resident arm after the timed steps)
            e2e_step(w)

    def loop(w, k):
        w["ctx"].bind()  # this thread drives one context: stay on its device
        for _ in range(k):
            e2e_step(w)

    # N > 1 (COVER_BENCH_BCAST=0 switches it off: every rank then uploads its own copy, as at N = 1): the planning cycle's
    # costmap reaches the host ONCE - rank 0 uploads it and fans it out
    # over NVLink (NCCL broadcast), every rank ingests it from device memory, sweeps its own window and downloads its own
    # verdicts.  One issuing thread per rank (collectives in the same order everywhere), two cycles in flight on the two
    # contexts' streams; a slot is reused when its results have landed in pinned host memory.
    use_bcast = world > 1 and os.environ.get("COVER_BENCH_BCAST", "1") == "1"
    if use_bcast:
        empty_pinned = torch.from_numpy(empty_map)
        for w in workers:
            # torch's own stream carries the copies and the collective (its allocators track the streams their tensors were
            # used on, and the library's stream dies with its context); events order it with the library's stream
            w["torch_stream"] = torch.cuda.Stream(device=local)
            w["lib_stream"] = torch.cuda.ExternalStream(w["ctx"].stream, device=local)
            w["img_dev"] = torch.empty((MAP_SIZE, MAP_SIZE), dtype=torch.uint8, device="cuda")
            w["bits_dev"] = torch.zeros(n_words, dtype=torch.int32, device="cuda")
            w["nok_dev"] = torch.zeros(1, dtype=torch.int64, device="cuda")
            w["out_dev"] = cover_b200.Result(free_bits=w["bits_dev"], not_ok=w["nok_dev"])
            w["done"] = torch.cuda.Event()
        torch.cuda.synchronize()

        def bcast_step(w, image_pinned):
            ts, ls = w["torch_stream"], w["lib_stream"]
            with torch.cuda.stream(ts):
                if rank == 0:
                    w["img_dev"].copy_(image_pinned, non_blocking=True)  # H2D: 16 MB, on rank 0 only
                dist.broadcast(w["img_dev"], src=0)
                arrived = ts.record_event()
            ls.wait_event(arrived)
            w["map"].upload(w["img_dev"])  # (device to device, then the sweep: both on the library's stream)
            w["map"].area_is_free(wl.RECT, lat, out=w["out_dev"])
            swept = ls.record_event()
            ts.wait_event(swept)
            with torch.cuda.stream(ts):
                w["bits"].copy_(w["bits_dev"], non_blocking=True)  # D2H: 12.5 MB of verdict bits + the counter
                w["not_ok"].copy_(w["nok_dev"], non_blocking=True)
                w["done"].record(ts)

        for w in workers:
            bcast_step(w, empty_pinned)
            w["done"].synchronize()
            assert bool(w["out"].unpack_bits(n_poses).all()), "e2e: the sweep did not see the broadcast (empty) costmap"
            bcast_step(w, w["pinned"])
            w["done"].synchronize()
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            w = workers[k % n_flight]
            w["done"].synchronize()
            bcast_step(w, w["pinned"])
        for w in workers:
            w["done"].synchronize()
        barrier()
        e2e_dt = time.perf_counter() - t0
    elif os.environ.get("COVER_BENCH_THREADS", "0") == "1":
        # one host thread per cycle in flight, every call returning when its results are on the host
        barrier()
        t0 = time.perf_counter()
        threads_ = [threading.Thread(target=loop, args=(w, e2e_steps // n_flight)) for w in workers]
        for th in threads_:
            th.start()
        for th in threads_:
            th.join()
        barrier()
        e2e_dt = time.perf_counter() - t0
    else:
        # ONE host thread keeps the cycles in flight: results go to pinned host memory without waiting
        # (COVER_MEM_HOST_ASYNC); a slot is reused when cover_ctx_synchronize says its previous cycle has landed
        for w in workers:
            w["out_async"] = cover_b200.Result(free_bits=w["out"].free_bits, not_ok=w["out"].not_ok, host_async=True)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            w = workers[k % n_flight]
            w["ctx"].synchronize()
            w["map"].upload_async(w["image"])
            w["map"].area_is_free(wl.RECT, lat, out=w["out_async"])
        for w in workers:
            w["ctx"].synchronize()
        barrier()
        e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_poses * e2e_steps / float(te.item())
    clocks = sampler.stop()  # sampled through both timed regions (device-resident and end-to-end)
    for w in workers:
        assert np.array_equal(w["out"].unpack_bits(n_poses), dev_free), "e2e verdicts differ from the device-resident arm"
        assert int(w["not_ok"][0]) == 0
    h2d = int(data.size + desc["cs"].size * 8 + 8 * 8)
    d2h = int(4 * n_words + 8)
    for w in workers:
        w["map"].close()
        w["ctx"].close()
    del workers

    # ---- CPU arm on this rank's window: timing (rank 0, N = 1) and the parity of the sampled poses (every rank)
    from oracle import pyoracle as po
    host_threads = po.hardware_threads()
    cpu = None
    cells_per_pose = None
    parity = None
    if not args.no_cpu_baseline:
        n_sample = 12_000_000 if world == 1 else 1_000_000
        rate, dt, cells_per_pose, n, kind, kept = cpu_sample(data, desc, n_sample, host_threads if world == 1 else max(1, host_threads // world))
        mism = sum(compare(dev_free[s:s + len(f)], np.zeros(len(f), dtype=np.uint8), f, st) for s, f, st in kept)
        parity = {"poses": n, "mismatches": mism, "against": kind, "window": window}
        if world == 1:
            cpu = {"value": rate, "unit": UNIT, "cores": host_threads, "kind": kind,
                   "sample": f"{n} poses of the same lattice ({CPU_CHUNK}-pose chunks at random offsets) in {dt:.1f} s on {host_threads} threads"}
        if world > 1:
            tot = torch.tensor([parity["poses"], parity["mismatches"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(tot)
            parity = {"poses": int(tot[0]), "mismatches": int(tot[1]), "against": kind, "window": "every rank's own"}
        assert parity["mismatches"] == 0, f"GPU verdicts differ from the CPU {kind}: {parity}"

    peak, peak_kind = measured_peak()
    secondary = None
    c5_strong = None
    if world > 1 and not args.no_secondary:
        # C5 strong scaling: 8 shapes x 1e7 poses, the ranks split every shape's poses
        rows = c5_sweep(ctx, stream, peak, 1, False, rank, world)
        ms_local = torch.tensor([r["ms"] for r in rows], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms_local, op=dist.ReduceOp.MAX)
        c5_strong = [{"config": r["config"], "poses_total": 10_000_000, "ms": float(m), "checks_per_s": 10_000_000 / (float(m) * 1e-3)}
                     for r, m in zip(rows, ms_local.tolist())]
    if rank == 0:
        if world == 1 and not args.no_secondary:
            secondary = secondary_configs(ctx, stream, peak, host_threads, not args.no_cpu_baseline)
        # ---- roofline of the step: bytes it has to move through HBM (the L2 is flushed before every step)
        rows_lo = int((desc["y0"] - 0.6) / RESOLUTION) - 2
        rows_hi = int((desc["y0"] + desc["ny"] * desc["dy"] + 0.6) / RESOLUTION) + 2
        rows = min(MAP_SIZE, rows_hi - rows_lo + 1)  # map rows under the lattice window +- the footprint's reach
        plane_bytes = 6 * rows * (MAP_SIZE // 8)
        moved = {"costmap_ingest_read+write": 2 * data.size, "plane_build_read": rows * MAP_SIZE, "plane_build_write": plane_bytes,
                 "sweep_plane_read": plane_bytes, "verdict_bits_write": 4 * n_words}  # (no zeroing pass: the sweep writes whole words)
        moved_total = sum(moved.values())
        achieved = moved_total / (ms_per_step * 1e-3) / 1e9
        rec = recorded_counters()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": rec.get("k_lw_sweep_dram_bytes_per_launch"), "peak_kind": peak_kind, "kernel": "k_lw_sweep_simple<4>",
                    "bytes_moved_per_step": moved, "sweep_only_ms": sweep_ms,
                    "note": "achieved = bytes the step must move through HBM / step time (measured in this run): the step is bound by "
                            "instruction issue and L1 wavefronts (bit-plane walk), not by HBM; `recorded` holds the ncu counters of the "
                            "dominant kernel from the committed capture, not from this run",
                    "recorded": rec}
        if cells_per_pose is not None:
            # SURVEY 8(d)'s figure, for the record: bytes the REFERENCE's sequential loop reads + verdict + status per check.
            # The kernel answers a w-cell window from two words of precomputed bit planes, so this is not a utilisation.
            alg = (cells_per_pose + 2) * n_poses
            roofline["algorithmic_bytes_per_check"] = cells_per_pose + 2
            roofline["algorithmic_gbs"] = alg / (ms_per_step * 1e-3) / 1e9
        l2_peak = ctx.probe_read_bandwidth(16 << 20, 50)
        roofline["l2"] = {"peak": l2_peak, "peak_kind": "measured live (cover_probe_read_bandwidth, 16 MB, best of 10)",
                          "traffic": rec.get("k_lw_sweep_l2_bytes_per_launch")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "cycles_in_flight": n_flight,
                    **({"costmap_fanout": "rank 0 uploads the 16 MB image (h2d_bytes_per_step is rank 0's), NCCL broadcast to the other ranks"} if use_bcast else {})},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "free_fraction": free_fraction, "wall_s_timed_region": wall, "step_ms": step_ms,
        }
        if secondary is not None:
            line["secondary"] = secondary
        if c5_strong is not None:
            line["c5_strong_scaling"] = c5_strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline / parity leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configs (C1, C2, C3, C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
