"""Per-call wall times of the e2e step on every rank of a torchrun launch, with and without an NCCL process group."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cover_b200
from cover_b200 import workloads as wl
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
use_dist = os.environ.get("PROBE_DIST", "1") == "1" and world > 1
if use_dist:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
data = wl.inflated_costmap(4000, 4000, 0.05, seed=3)
desc = wl.dwa_lattice(x0=20.0 + 2 * rank, y0=20.0 + 3 * rank)
lat = cover_b200.Lattice(**desc); n = lat.count
ctx = cover_b200.Context(local)
cmap = cover_b200.Costmap(ctx, data, 0.05)
pinned = torch.from_numpy(data).pin_memory().numpy()
hb = torch.empty((n + 31) // 32, dtype=torch.int32).pin_memory(); hn = torch.empty(1, dtype=torch.int64).pin_memory()
out = cover_b200.Result(free_bits=hb.numpy().view(np.uint32), not_ok=hn.numpy())
ts = {"upload_async": [], "sweep": []}
for k in range(30):
    t0 = time.perf_counter(); cmap.upload_async(pinned); t1 = time.perf_counter(); cmap.area_is_free(wl.RECT, lat, out=out); t2 = time.perf_counter()
    if k >= 5:
        ts["upload_async"].append(t1 - t0); ts["sweep"].append(t2 - t1)
print(f"rank {rank} dist={use_dist} OMP={os.environ.get('OMP_NUM_THREADS')}: upload_async call med {np.median(ts['upload_async'])*1e3:.3f} ms, sweep call med {np.median(ts['sweep'])*1e3:.3f} ms max {max(ts['sweep'])*1e3:.3f}", flush=True)
if use_dist:
    dist.destroy_process_group()
