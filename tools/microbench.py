"""Kernel-only timing of synthetic uniform batches (development aid).
    python tools/microbench.py B T L [B T L ...] [--beam W] [--reps N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from kokoro_align_b200 import align, synth  # noqa: E402

args = sys.argv[1:]
beam, reps = 1000, 5
if "--beam" in args:
    i = args.index("--beam"); beam = int(args[i + 1]); del args[i:i + 2]
if "--reps" in args:
    i = args.index("--reps"); reps = int(args[i + 1]); del args[i:i + 2]
vals = [int(x) for x in args]
for j in range(0, len(vals), 3):
    B, T, L = vals[j:j + 3]
    Ts, Ls = np.full(B, T), np.full(B, L)
    lp, t_off, labels, l_off = synth.make_batch_fast(Ts, Ls, seed=1)
    plan = align.AlignPlan(t_off, labels, l_off, 39, beam_size=beam)
    d_lp = torch.from_numpy(lp).cuda()
    for _ in range(2):
        out = plan.run_torch(d_lp)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for r in range(reps):
        out = plan.run_torch(d_lp)
        ev[r + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
    cells = int(plan.info.cells_eval)
    assert (out[4].cpu().numpy() == 0).all()
    print(f"B={B} T={T} L={L} S={2*L+1} W={beam}: {ms:.4f} ms  {cells/ms/1e6:.1f} Gcells/s  "
          f"{ms*1e6/ (T):.1f} ns/frame/lattice-serial  classes={list(plan.info.n_class)}", flush=True)
    plan.close()
