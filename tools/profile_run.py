"""Minimal launcher for ncu: builds one workload, runs the device-resident path a few times.
    python tools/profile_run.py --workload segments|gon|chapters [--runs 3] [--lattices N]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import workload_shapes  # noqa: E402
from kokoro_align_b200 import align, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="segments")
ap.add_argument("--runs", type=int, default=3)
ap.add_argument("--lattices", type=int, default=None)
ap.add_argument("--beam", type=int, default=1000)
a = ap.parse_args()
T, L, desc = workload_shapes(a.workload, 2000, a.lattices)
lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2001)
plan = align.AlignPlan(t_off, labels, l_off, 39, beam_size=a.beam)
d_lp = torch.from_numpy(lp).cuda()
for _ in range(a.runs):
    out = plan.run_torch(d_lp)
torch.cuda.synchronize()
assert (out[4].cpu().numpy() == 0).all()
print(desc, "cells", int(plan.info.cells_eval), "ok")
