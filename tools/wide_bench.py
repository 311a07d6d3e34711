import sys, numpy as np, torch, time
sys.path.insert(0, ".")
from kokoro_align_b200 import align, synth
import os
shapes = ((20000, 10000), (100000, 50000)) + (((1000000, 50000),) if os.environ.get("KAB_WIDE_HUGE") else ())
for T, L in shapes:
    S = 2 * L + 1
    lp, t_off, labels, l_off = synth.make_batch_fast(np.array([T]), np.array([L]), seed=9)
    plan = align.AlignPlan(t_off, labels, l_off, 39, beam_size=2 * S + 2)
    d = torch.from_numpy(lp).cuda()
    o = plan.run_torch(d); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(3):
        e0.record(); o = plan.run_torch(d); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    st = o[4].cpu().numpy()
    print(f"T={T} L={L} S={S} unbanded: {ms:.2f} ms (3 runs: {", ".join("%.1f" % t for t in times)}), {plan.info.cells_eval/ms/1e6:.1f} Gcells/s, bp {plan.info.backptr_bytes/1e9:.2f} GB, alg GB/s {plan.info.algorithmic_bytes/ms/1e6:.0f}, status {st}, classes {list(plan.info.n_class)}", flush=True)
    plan.close(); del d
