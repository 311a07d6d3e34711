"""One Gon-gitsune-shaped lattice (BASELINE config 1), device resident, a few timed runs; with a
-DKAB_BANDQ_TIMING / -DKAB_BANDP_TIMING build (KAB_LIBRARY=...) the kernel's per-warp cycle
breakdown goes to stderr.    python tools/gon_once.py [T L]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_align_b200 import align, synth  # noqa: E402

T, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (81135, 11359)
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
lp, t_off, labels, l_off = synth.make_batch_fast(np.array([T]), np.array([L]), seed=1001)
plan = align.AlignPlan(t_off, labels, l_off, 39, beam_size=W)
d = torch.from_numpy(lp).cuda()
for k in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = plan.run_torch(d)
    e1.record()
    torch.cuda.synchronize()
    print(f"run {k}: {e0.elapsed_time(e1):.3f} ms, {e0.elapsed_time(e1) * 1e6 / T:.1f} ns/frame, status {o[4].cpu().numpy()}", flush=True)
plan.close()
