"""BASELINE config 5: synthetic sweep T x L (x beam) -> cells/s and algorithmic GB/s per point.
    python tools/sweep.py [--quick] [--wide-v | --unbanded] > profiles/rNN_sweep.json
Each point is a batch of identical-shape lattices sized to fill the GPU (B lattices), timed
device-resident with CUDA events (best of 3).  Points with S > 2T are skipped."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from kokoro_align_b200 import align, synth  # noqa: E402

quick = "--quick" in sys.argv
Ts = [1000, 10000, 100000] if quick else [1000, 10000, 100000, 1000000]
Ls = [100, 1000, 10000] if quick else [100, 1000, 10000, 50000]
peak = 6538.6
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = []
points = [(T, L, 39) for T in Ts for L in Ls]
# wide vocabularies (whole rows are staged up to V = 512)
# (V = 4096: compact per-lattice copy of the used columns for the short lattices, kab_compact.cuh; the
# band-shaped ones -- more than 511 distinct labels from L = 1000 on -- in kab_bandr.cuh's gather mode)
wide_v = [(1000, 100, 256), (10000, 1000, 256), (1000, 100, 512), (1000, 100, 4096), (10000, 1000, 4096),
          (100000, 10000, 4096)]
points = wide_v if "--wide-v" in sys.argv else points + ([] if quick else wide_v)
unbanded = "--unbanded" in sys.argv   # beam_size covers the lattice (config 5: where T*S <= 1e11)
if unbanded:
    # (S <= 1000: the default band already covers the lattice, same numbers as the banded sweep)
    points = [(T, L, 39) for T in Ts for L in Ls if 2 * L + 1 > 1000 and T * (2 * L + 1) <= 1.001e11]
for T, L, V in points:
    if True:
        S = 2 * L + 1
        if S > 2 * T:
            continue
        frames_budget = 6_000_000 * 39 // V          # keep inputs below ~1 GB per point
        B = int(max(1, min(4096, frames_budget // T)))
        if unbanded and S > 248:                      # band / wide kernels: cells, not frames, bound the point
            B = int(max(1, min(B, 4 * 10 ** 10 // (T * S))))
        lp, t_off, labels, l_off = synth.make_batch_fast(np.full(B, T), np.full(B, L), V=V, seed=5000 + T % 97 + L % 89)
        beam = 2 * S + 2 if unbanded else 1000
        plan = align.AlignPlan(t_off, labels, l_off, V, beam_size=beam)
        d_lp = torch.from_numpy(lp).cuda()
        for _ in range(2):
            o = plan.run_torch(d_lp)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = plan.run_torch(d_lp); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        st = o[4].cpu().numpy()
        info = plan.info
        rec = dict(T=T, L=L, S=S, B=B, V=V, beam_size=beam, ms=best, status_ok=bool((st == 0).all()),
                   cells_eval=int(info.cells_eval), cells_nominal=int(info.cells_nominal),
                   cells_eval_per_s=info.cells_eval / best * 1e3, cells_nominal_per_s=info.cells_nominal / best * 1e3,
                   algorithmic_gbs=info.algorithmic_bytes / best / 1e6, hbm_frac=info.algorithmic_bytes / best / 1e6 / peak,
                   ns_per_frame=best * 1e6 / T, lp_gbs=lp.nbytes / best / 1e6, classes=list(info.n_class))
        out.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
        plan.close()
        del d_lp
print(json.dumps(out, indent=1))
