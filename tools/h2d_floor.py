"""Host-to-device copy floor of the end-to-end path, alone and with N ranks copying at once.

    python tools/h2d_floor.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_floor.py

Every rank copies the config-2 batch (742 MB of log-probs) from pinned host memory to its GPU, all
ranks at the same time (barrier before every repetition), in 1 and in 12 chunks.  Prints one JSON
line on rank 0: per-rank GB/s (min / max), the aggregate, and the time of the slowest rank -- the
floor bench.py's `e2e` (kab_plan_run_host: H2D + kernels + D2H) can reach at that N on this box.
Also prints what the box says about topology (GPU <-> CPU affinity, NUMA nodes), because the
aggregate is a property of the host, not of this library."""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_align_b200 import parallel  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
bind = "--no-bind" not in sys.argv
cpus = parallel.bind_host_to_gpu(local, world) if (world > 1 and bind) else None
rt = ctypes.CDLL("libcudart.so.12")
n = 742_319_760
d = torch.empty(n, dtype=torch.uint8, device=dev)
p = ctypes.c_void_p()
assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0)) == 0
ctypes.memset(p, 1, n)           # first touch on this rank's cores
s = torch.cuda.Stream()
res = {}
for chunks in (1, 12):
    times = []
    for rep in range(6):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        sz = n // chunks
        for c in range(chunks):
            rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr() + c * sz), ctypes.c_void_p(p.value + c * sz),
                               ctypes.c_size_t(sz), ctypes.c_int(1), ctypes.c_void_p(s.cuda_stream))
        s.synchronize()
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = [float(x.item()) for x in allt]
    else:
        allt = [best]
    res[f"chunks_{chunks}"] = {"slowest_rank_ms": max(allt) * 1e3, "per_rank_gbs_min": n / max(allt) / 1e9,
                               "per_rank_gbs_max": n / min(allt) / 1e9, "aggregate_gbs": world * n / max(allt) / 1e9}
if rank == 0:
    topo = {}
    for name, cmd in (("nvidia_smi_topo", ["nvidia-smi", "topo", "-m"]), ("lscpu_numa", ["bash", "-c", "lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)'"])):
        try:
            topo[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[:24]
        except Exception as e:  # noqa: BLE001
            topo[name] = [f"unavailable: {e}"]
    print(json.dumps({"n_gpus": world, "bytes_per_rank": n, "host_binding": (f"{len(cpus)} cores per rank" if cpus else "none"),
                      "affinity_rank0": sorted(os.sched_getaffinity(0))[:4] + ["..."], **res, "topology": topo}), flush=True)
rt.cudaFreeHost(p)
if world > 1:
    dist.destroy_process_group()
