"""Generate tests/golden/ from the REFERENCE's own ctc_best_path.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

For every case the reference function kokoro_align.align.ctc_best_path (align.py:43-109)
is imported unmodified and executed; its outputs (or the exception type it raises) are
stored.  "gauss" cases store their fp32 inputs too (numpy's exp/log may differ in the last
ulp between CPUs); "exact" cases are integer-derived and are regenerated from the recipe.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from kokoro_align.align import ctc_best_path as ref_ctc_best_path  # noqa: E402
from kokoro_align_b200 import synth  # noqa: E402


def cases():
    c = []

    def add(name, kind, T, L, V=39, seed=0, beam=1000, mm=4, **kw):
        c.append(dict(name=name, kind=kind, T=T, L=L, V=V, seed=seed, beam_size=beam,
                      max_move=mm, kw=kw))

    # (1,2,3) iid / planted / quantised Gaussian-softmax lattices (inputs stored)
    add("gauss_iid_small", "gauss", 400, 60, seed=101)
    add("gauss_planted", "gauss", 1200, 170, seed=102, planted=True)
    for q in (1, 2, 4):
        add(f"gauss_quant{q}", "gauss", 600, 90, seed=110 + q, quant=q)
    add("gauss_iid_band", "gauss", 1500, 700, seed=103, beam=300)
    # (4) beam sizes
    for w in (6, 16, 24, 30, 40, 1000, 2000):
        add(f"beam{w}", "exact", 3000, 420, seed=200 + w, beam=w, planted=True)
        add(f"beam{w}_iid", "exact", 1000, 140, seed=300 + w, beam=w)
    # (5) S/T ratios
    for r, L in ((0.05, 50), (0.28, 280), (1.3, 1300), (2.5, 2500)):
        add(f"ratio{r}", "exact", 2000, L, seed=400 + L, beam=200, planted=True)
        add(f"ratio{r}_w1000", "exact", 2000, L, seed=450 + L, planted=True)
    # (6) flush-interval lengths
    for T in (1, 2, 3, 7, 9999, 10000, 10001, 20500, 30011):
        add(f"T{T}", "exact", T, int(round(0.14 * T)), seed=500 + T % 97, planted=True)
    add("T10001_w200", "exact", 10001, 1400, seed=520, beam=200, planted=True)
    add("T1_L1", "exact", 1, 1, seed=521)
    add("T1_L2", "exact", 1, 2, seed=522)
    add("T2_L3", "exact", 2, 3, seed=523)
    # (7) tiny transcripts
    add("L0", "exact", 50, 0, seed=600)
    add("L1", "exact", 50, 1, seed=601)
    add("L0_T1", "exact", 1, 0, seed=602)
    # (8) repeated labels
    add("repeat", "exact", 800, 200, seed=700, repeat_labels=True, planted=True)
    add("repeat_gauss", "gauss", 500, 120, seed=701, repeat_labels=True)
    # (9) dead band
    add("dead_band", "exact", 40, 100, seed=800, beam=20)
    add("dead_band2", "exact", 10, 40, seed=801)          # S > 3T+1: end unreachable but band ok
    # (10) max_move
    for mm in (1, 2, 3, 4, 5, 6, 8):
        add(f"mm{mm}", "exact", 900, 130, seed=900 + mm, mm=mm, planted=True)
        add(f"mm{mm}_band", "exact", 1500, 600, seed=950 + mm, mm=mm, beam=120, planted=True)
        add(f"mm{mm}_dense", "exact", 300, 280, seed=970 + mm, mm=mm)
    # (11) label value 0 inside the transcript (value-based blank test), negative label
    add("zero_label", "exact", 700, 150, seed=1000, zero_every=7)
    add("neg_label", "exact", 300, 40, seed=1001, neg_every=5)
    # (12) label >= V
    add("bad_label", "exact", 100, 10, seed=1002, bad_at=3)
    # (14) other vocab sizes
    add("V256", "exact", 800, 110, V=256, seed=1100, planted=True)
    add("V5", "exact", 800, 110, V=5, seed=1101, planted=True)
    add("V64", "exact", 500, 300, V=64, seed=1102)
    # (15) S < W but bottom-clipped window (S-1 > W//2)
    add("clip_bottom", "exact", 1500, 400, seed=1200, planted=True)
    add("clip_bottom_iid", "exact", 900, 450, seed=1201)
    # short-segment shapes of config 2
    for n, T in enumerate((86, 200, 431, 640, 861)):
        add(f"seg{T}", "gauss", T, int(round(0.14 * T)), seed=1300 + n)
        add(f"seg{T}_ties", "exact", T, int(round(0.14 * T)), seed=1350 + n, levels=4, scale=1.0)
    return c


def build_inputs(case):
    kw = dict(case["kw"])
    zero_every = kw.pop("zero_every", 0)
    neg_every = kw.pop("neg_every", 0)
    bad_at = kw.pop("bad_at", None)
    fn = synth.make_lattice if case["kind"] == "gauss" else synth.make_lattice_exact
    lp, labels = fn(case["T"], case["L"], case["V"], case["seed"], **kw)
    labels = labels.copy()
    if zero_every:
        labels[::zero_every] = 0
    if neg_every:
        labels[::neg_every] = -1 - (np.arange(len(labels[::neg_every])) % case["V"])
    if bad_at is not None:
        labels[bad_at] = case["V"]
    return lp, labels


def main():
    index, arrays = [], {}
    t_all = time.time()
    for case in cases():
        lp, labels = build_inputs(case)
        t0 = time.time()
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                path, labs, scores = ref_ctc_best_path(lp, labels, beam_size=case["beam_size"],
                                                       max_move=case["max_move"])
            outcome = "ok"
        except (ValueError, IndexError) as e:
            outcome = type(e).__name__
        dt = time.time() - t0
        name = case["name"]
        entry = dict(case, outcome=outcome, ref_seconds=round(dt, 3))
        if outcome == "ok":
            assert path.dtype == np.int32 and labs.dtype == np.int32 and scores.dtype == np.float32
            arrays[f"{name}.best_path"] = path
            # best_labels / best_scores are gathers of the inputs along the path; keep them
            # for the stored-input cases so the gather itself is pinned as well.
            if case["kind"] == "gauss":
                arrays[f"{name}.best_labels"] = labs
                arrays[f"{name}.best_scores"] = scores
            # DP end score == sequential fp32 sum of the emissions along the path
            entry["final_score_hex"] = np.cumsum(scores, dtype=np.float32)[-1].tobytes().hex()
        if case["kind"] == "gauss":
            arrays[f"{name}.log_probs"] = lp
            arrays[f"{name}.labels"] = labels.astype(np.int32)
        index.append(entry)
        print(f"{name:24s} T={case['T']:6d} L={case['L']:5d} W={case['beam_size']:5d} "
              f"M={case['max_move']} -> {outcome} ({dt:.2f}s)")
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden_index.json"), "w") as f:
        json.dump(dict(reference="kaiidams/Kokoro-Align kokoro_align/align.py:43-109",
                       numpy=np.__version__, cases=index), f, indent=1)
    print(f"{len(index)} cases in {time.time() - t_all:.1f}s")


if __name__ == "__main__":
    main()
