import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from kokoro_align_b200 import align, synth
T, L = synth.segment_lengths(10000, 2000)
lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2001)
import ctypes
from kokoro_align_b200 import _lib
hlp = align.pinned_like(lp) if hasattr(align, "pinned_like") else lp
for k in range(3):
    t0 = time.perf_counter()
    plan = align.AlignPlan(t_off, labels, l_off, 39)
    t1 = time.perf_counter()
    out = plan.run_host(lp)
    t2 = time.perf_counter()
    out = plan.run_host(lp)
    t3 = time.perf_counter()
    plan.close()
    t4 = time.perf_counter()
    print(f"iter {k}: create {1e3*(t1-t0):.2f} ms, first run_host {1e3*(t2-t1):.2f}, second {1e3*(t3-t2):.2f}, close {1e3*(t4-t3):.2f}", flush=True)
