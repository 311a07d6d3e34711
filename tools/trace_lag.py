"""Per-pass trace of kab_bandr_kernel's compute warps.

A -DKAB_BANDR_TIMING build writes, for every two-group pass of every compute warp, its start and end
(SM clock), the tile wait, both message waits and the time in the frames; KAB_TRACE_FILE=path dumps
them as raw int64 [64 warps][16384 groups][2]:
    KAB_LIBRARY=libkab_timing.so KAB_TRACE_FILE=trace.bin python tools/gon_once.py
    python tools/trace_lag.py trace.bin [n_warps n_groups]
Prints, per role (passes without / with a neighbour's message), the mean duration of a pass and where
it goes, and the spread of the chain.  The first version of this trace (start time of every group,
%globaltimer) is what showed warps 4.5 groups apart with nobody waiting: the start-up lag a joining
warp was given, paid again at every change of the head."""
import sys

import numpy as np

path = sys.argv[1]
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 28
NG = int(sys.argv[3]) if len(sys.argv) > 3 else 10142
tr = np.fromfile(path, dtype=np.int64).reshape(64, 16384, 2)
g = np.arange(2, NG - 4, 2)
st, pk, en = tr[:NW][:, g, 0], tr[:NW][:, g, 1], tr[:NW][:, g + 1, 0]
ok = (st > 0) & (en > 0)
tile, w1, w2 = pk & 0xfff, (pk >> 12) & 0xfff, (pk >> 24) & 0xfff
fr, n0, n1 = (pk >> 36) & 0x3fff, (pk >> 50) & 1, (pk >> 51) & 1
dur = en - st
nxt = np.zeros_like(st)
nxt[:, :-1] = st[:, 1:]
gap = nxt - en


def stats(m, name):
    m = m & ok
    m2 = m.copy()
    m2[:, -1] = False
    if not m.any():
        return
    print(f"{name:34s} passes {m.sum():7d}: {dur[m].mean():6.0f} cycles = tile {tile[m].mean():4.0f} + message polls "
          f"{w1[m].mean():4.0f} + {w2[m].mean():4.0f} + frames {fr[m].mean():5.0f} + rest {(dur - tile - w1 - w2 - fr)[m].mean():4.0f}; "
          f"to the next pass {gap[m2].mean():4.0f}")


print(f"{ok.mean() * 100:.1f} % of the passes traced (the others took the general body)")
stats(np.ones_like(ok), "all passes")
stats((n0 == 0) & (n1 == 0), "no message (head / outside the window)")
stats((n0 == 1) & (n1 == 1), "followers")
print("(cycle counts include the clock reads of the timing build, ~20-40 each)")
# spread of the chain: time from the first to the last warp starting the same pass (same SM clock only
# within a CTA: CW consecutive warps)
CW = 4
k = len(g) // 2
for c in range(NW // CW):
    s4 = st[c * CW:(c + 1) * CW, k]
    print(f"CTA {c}: pass {int(g[k])} starts (cycles after the CTA's first warp)", (s4 - s4.min()).tolist())
