"""Latency of the drop-in call ctc_best_path(log_probs, labels) (one plan per call: create, copy in,
align, copy out, destroy) against the device time of the same lattice.
    python tools/call_overhead.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from kokoro_align_b200 import align, synth  # noqa: E402

for name, T, L in (("segment 5 s", 431, 60), ("chapter 10 min", 52000, 7280), ("Gon gitsune", 81135, 11359)):
    lp, labels = synth.make_lattice(T, L, 39, seed=1)
    for _ in range(3):
        align.ctc_best_path(lp, labels)
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        align.ctc_best_path(lp, labels)
    call_ms = (time.perf_counter() - t0) / n * 1e3
    with align.AlignPlan([0, T], labels, [0, L], 39) as plan:
        d = torch.from_numpy(lp).cuda()
        for _ in range(3):
            plan.run_torch(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            plan.run_torch(d)
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / n
        t0 = time.perf_counter()
        for _ in range(n):
            plan.run_host(lp)
        host_ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:16s} T={T:6d}: ctc_best_path() {call_ms:8.3f} ms | plan.run_host {host_ms:8.3f} ms | device {dev_ms:8.3f} ms")
