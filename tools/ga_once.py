import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from kokoro_align_b200 import align, synth
T, L, V = 100000, 10000, 4096
lp, t_off, labels, l_off = synth.make_batch_fast(np.array([T]), np.array([L]), V=V, seed=77)
plan = align.AlignPlan(t_off, labels, l_off, V, beam_size=1000)
d = torch.from_numpy(lp).cuda()
for k in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = plan.run_torch(d); e1.record(); torch.cuda.synchronize()
    print(f"run {k}: {e0.elapsed_time(e1):.3f} ms, {e0.elapsed_time(e1)*1e6/T:.1f} ns/frame, status {o[4].cpu().numpy()}", flush=True)
plan.close()
