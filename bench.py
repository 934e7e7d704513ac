#!/usr/bin/env python
"""bench.py — footprint-pose checks/s on BASELINE.json config 4 (DWA lattice).

Workload (one "step" = one pass of the hot path over the whole batch): 4000 x 4000 uint8 costmap at
5 cm (sparse obstacles inflated like costmap_2d's InflationLayer), 1.0 x 0.6 m rectangular footprint,
(x, y, theta) lattice 1180 x 1180 x 72 = 100 252 800 poses, verdict only (first-lethal early exit) —
the batched `is_free(polygon, pose, map)` of src/cover/costmap.cpp:78-85.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): poses are sharded, the costmap is replicated, no
data-path collective (weak scaling: every rank checks its own full lattice window).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
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
B_OUT = 2  # verdict byte + status byte per pose; B_pose = 0 for a lattice


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "C4 DWA lattice: 4000x4000 inflated costmap @0.05 m, 1.0x0.6 m rect footprint, "
                    "1180x1180x72 (x,y,theta) lattice = 100252800 poses per GPU, is_free (first-lethal)",
        "map": [MAP_SIZE, MAP_SIZE], "resolution_m": RESOLUTION, "footprint": "rect 1.0x0.6 m",
        "poses_per_gpu": 1180 * 1180 * 72, "pose_format": "lattice", "op": "area_is_free",
        "parallelism": f"pose-sharded x{n_gpus}, costmap replicated, no collective",
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


def cpu_sample(data: np.ndarray, desc: dict, n_sample: int, threads: int, seed: int = 11):
    """Times the CPU implementation on `n_sample` poses of the SAME lattice (random contiguous chunks) and
    returns (checks/s, seconds, cells_read_per_pose from the restatement's tally)."""
    from oracle import pyoracle as po
    backend, kind = cpu_backend()
    omap = po.Costmap(data, RESOLUTION)
    olat = po.Lattice(**desc)
    rng = np.random.default_rng(seed)
    chunk = 20_000
    starts = rng.integers(0, olat.count - chunk, max(1, n_sample // chunk))
    t0 = time.perf_counter()
    for s in starts.tolist():
        backend.polygon_batch(omap, wl.RECT, olat, first=s, count=chunk, threads=threads)
    dt = time.perf_counter() - t0
    n = len(starts) * chunk
    # algorithmic bytes: cells the reference's sequential loop reads (the restatement counts them)
    tally = 0
    probe = starts[: max(1, min(len(starts), 10))]
    for s in probe.tolist():
        tally += po.default().polygon_batch(omap, wl.RECT, olat, first=s, count=chunk, threads=threads).cells_read
    return n / dt, dt, tally / (len(probe) * chunk), n, kind


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
        cpu_sample(data, desc, 40_000, threads)
    rates = []
    t_total = 0.0
    kind = "port"
    for k in range(args.steps):
        rate, dt, _, n, kind = cpu_sample(data, desc, per_step, threads, seed=100 + k)
        rates.append(rate)
        t_total += dt
    value = per_step * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{per_step} poses of the same lattice per step (random 20000-pose chunks), {threads} threads"},
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


def recorded_traffic(key="k_area_lattice_is_free_dram_bytes_per_launch"):
    """DRAM (or L2) bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(key)
        except Exception:
            return None
    return None


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
        # way - a 0.5 ms step is short enough for the launching thread's scheduling jitter to show in the device time
        torch.set_num_threads(1)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    data = make_map()
    desc = lattice_for_rank(int(os.environ.get("COVER_BENCH_WINDOW", rank)))  # (override: time another rank's window on one GPU)
    n_poses = desc["nx"] * desc["ny"] * desc["cs"].shape[0]

    ctx = cover_b200.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    lat = cover_b200.Lattice(**desc)

    # ---- device-resident arm (`value`): map in HBM, lattice descriptor, verdicts + status written to HBM
    cmap = cover_b200.Costmap(ctx, data, RESOLUTION)
    out = cover_b200.Result(torch.empty(n_poses, dtype=torch.uint8, device="cuda"), None,
                            torch.empty(n_poses, dtype=torch.uint8, device="cuda"))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    map_dev = torch.from_numpy(data).cuda()  # the costmap of the timed steps is already in HBM

    def step():
        # one planning cycle: ingest the (device-resident) costmap - which also rebuilds the lethal bit planes the
        # sweep reads, nothing derived from the map is carried over between steps - and sweep the lattice
        cmap.upload(map_dev)
        cmap.area_is_free(wl.RECT, lat, out=out)

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()

    sampler = ClockSampler(local)
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
    free_fraction = float(out.is_free.float().mean().item())

    # ---- end-to-end arm (`e2e`): host buffers through the public API, copies inside the timed region
    pinned_map = torch.from_numpy(data).pin_memory()
    n_words = (n_poses + 31) // 32
    host_bits = torch.empty(n_words, dtype=torch.int32).pin_memory()
    host_not_ok = torch.empty(1, dtype=torch.int64).pin_memory()
    host_out = cover_b200.Result(free_bits=host_bits.numpy().view(np.uint32), not_ok=host_not_ok.numpy())
    map_np = pinned_map.numpy()

    def e2e_step(image=map_np):
        # H2D: this planning cycle's costmap (16 MB), queued as row bands; the sweep below starts on the first
        # bands while the rest is in flight (cover_map_upload_async)
        cmap.upload_async(image)
        # lattice descriptor in; bit-packed verdicts (12.5 MB) + the not-OK counter out (D2H); returns when they are on the host
        cmap.area_is_free(wl.RECT, lat, out=host_out)

    e2e_steps = max(2, min(args.steps, 5))
    # warm-up, and a check that every step really sees ITS upload: an empty map first (everything free) ...
    empty_map = torch.zeros_like(pinned_map).pin_memory().numpy()
    e2e_step(empty_map)
    assert bool(host_out.unpack_bits(n_poses).all()), "e2e: the sweep did not see the freshly uploaded (empty) costmap"
    for _ in range(2):  # ... then the real one (compared with the device-resident arm after the timed steps)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_poses * e2e_steps / float(te.item())
    clocks = sampler.stop()  # sampled through both timed regions (device-resident and end-to-end)
    e2e_free = host_out.unpack_bits(n_poses)
    assert np.array_equal(e2e_free, out.is_free.cpu().numpy()), "e2e verdicts differ from the device-resident arm"
    assert int(host_not_ok[0]) == 0
    h2d = int(data.size + desc["cs"].size * 8 + 8 * 8)
    d2h = int(4 * n_words + 8)

    if rank == 0:
        # ---- CPU baseline + algorithmic bytes (rank 0, N = 1 only runs the timing; bytes are needed always)
        cpu = None
        cells_per_pose = None
        if not args.no_cpu_baseline:
            from oracle import pyoracle as po
            threads = po.hardware_threads()
            n_sample = 12_000_000 if world == 1 else 200_000
            rate, dt, cells_per_pose, n, kind = cpu_sample(data, desc, n_sample, threads)
            if world == 1:
                cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                       "sample": f"{n} poses of the same lattice (random 20000-pose chunks) in {dt:.1f} s on {threads} threads"}
        peak, peak_kind = measured_peak()
        roofline = None
        if cells_per_pose is not None:
            # dominant kernel = k_area_lattice<OP_IS_FREE>: one launch per step (+ a 16 MB device copy, a ~13 us bit-plane
            # build and a ~3 us drain launch); its duration ~ the step's event time
            alg_bytes = (cells_per_pose + B_OUT) * n_poses
            achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": recorded_traffic(), "peak_kind": peak_kind, "kernel": "k_area_lattice<OP_IS_FREE>",
                        "algorithmic_bytes_per_check": cells_per_pose + B_OUT}
            # the costmap is L2/L1-resident, so also state the L2 roof (measured live: streaming 16-B read of 16 MB)
            l2_peak = ctx.probe_read_bandwidth(16 << 20, 50)
            l2_bytes = recorded_traffic("k_area_lattice_is_free_l2_bytes_per_launch")
            roofline["l2"] = {"peak": l2_peak, "peak_kind": "measured live (cover_probe_read_bandwidth, 16 MB, best of 10)",
                              "algorithmic_frac": achieved / l2_peak, "traffic": l2_bytes,
                              "achieved": None if l2_bytes is None else l2_bytes / (ms_per_step * 1e-3) / 1e9}
            # what actually bounds the kernel (ncu capture under profiles/): issue slots, not bytes - the bit planes
            # turn a w-cell window into two bit tests, so far fewer bytes move than the reference's loop reads
            roofline["issue_active_pct"] = recorded_traffic("k_area_lattice_is_free_issue_active_pct")
            roofline["note"] = ("frac > 1: `achieved` counts the bytes the reference's sequential loop reads (SURVEY 8d); the kernel "
                                "answers each scan-row window from two words of precomputed lethal bit planes, L1/L2-resident")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "free_fraction": free_fraction, "wall_s_timed_region": wall, "step_ms": step_ms,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
