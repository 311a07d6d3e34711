"""Unbanded book-length lattices (BASELINE config 4(ii)) in the chain-of-warps kernel (kab_wide.cuh):
time per lattice and, on EVERY run -- the 10^11-cell one included --, the size-independent
invariants of the result (the C oracle stops being practical at ~10^9 cells):
  monotone path, moves <= 3, path ends at the highest state reached, best_labels == ext[path],
  best_scores == lp[t, labels] bit for bit, sequential fp32 sum of best_scores == final_score bit for bit.
    python tools/wide_bench.py            (KAB_WIDE_HUGE=1 adds T = 10^6, L = 5*10^4: 30.8 GB of backpointers)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_align_b200 import align, synth  # noqa: E402

shapes = ((20000, 10000), (100000, 50000)) + (((1000000, 50000),) if os.environ.get("KAB_WIDE_HUGE") else ())
for T, L in shapes:
    S = 2 * L + 1
    lp, t_off, labels, l_off = synth.make_batch_fast(np.array([T]), np.array([L]), seed=9)
    plan = align.AlignPlan(t_off, labels, l_off, 39, beam_size=2 * S + 2)
    d = torch.from_numpy(lp).cuda()
    o = plan.run_torch(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(3):
        e0.record()
        o = plan.run_torch(d)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    st = o[4].cpu().numpy()
    path, labs, scores, final = (x.cpu().numpy() for x in o[:4])
    assert st[0] == 0
    dp = np.diff(path)
    assert path[0] in (0, 1, 3) and dp.min() >= 0 and dp.max() <= 3, "path is not a monotone <= 3 walk"
    ext = np.zeros(S, np.int32)
    ext[1::2] = labels
    assert np.array_equal(labs, ext[path]), "best_labels != ext[best_path]"
    assert scores.tobytes() == lp[np.arange(T), labs].tobytes(), "best_scores is not the gather of log_probs"
    assert np.cumsum(scores, dtype=np.float32)[-1].tobytes() == final[0].tobytes(), "final_score != sequential sum"
    print(f"T={T} L={L} S={S} unbanded: {ms:.2f} ms (3 runs: {', '.join('%.1f' % t for t in times)}), "
          f"{plan.info.cells_eval / ms / 1e6:.1f} Gcells/s, bp {plan.info.backptr_bytes / 1e9:.2f} GB, "
          f"alg GB/s {plan.info.algorithmic_bytes / ms / 1e6:.0f}, status {st}, classes {list(plan.info.n_class)}, "
          f"invariants ok (monotone, moves<=3, labels, scores gather, cumsum == final_score)", flush=True)
    plan.close()
    del d
