"""Wall time per book, files in -> files out (BASELINE.json: "wall-time per book vs host-CPU ref").
    python tools/book_wall.py [--chapters 36] [--frames 2721800] [--dir /dev/shm/kab_book]
Writes a synthetic config-3 book (C chapter `*.logits.npz` + `*.voca.txt`, SURVEY.md 8d: log-normal
chapter lengths, L = 0.14 T) and times align.best_path_files on it: (a) host numpy normalisation
(bit-identical to the reference), (b) device normalisation (kab_softmax.cuh), each twice (the
first call pays cudaHostAlloc of the pinned batch buffer).  The reference's own loop
(run_example.py:247-254 -> align.py:112-124) is timed on a bounded sample with the numpy port and
scaled linearly in frames (it is one core, ~3e2 us per frame)."""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_align_b200 import align, encoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chapters", type=int, default=36)
ap.add_argument("--frames", type=int, default=2_721_800)
ap.add_argument("--dir", default="/dev/shm/kab_book")
ap.add_argument("--cpu-sample-frames", type=int, default=40_000)
args = ap.parse_args()

rng = np.random.default_rng(3000)
w = rng.lognormal(0.0, 0.5, args.chapters)
T = np.maximum(100, np.round(w / w.sum() * args.frames)).astype(np.int64)
os.makedirs(args.dir, exist_ok=True)
tokens = [t for t in encoder.VOCAB[1:]]
lf, vf, bf = [], [], []
t0 = time.perf_counter()
for c, t in enumerate(T):
    lf.append(os.path.join(args.dir, f"c{c:02d}.logits.npz"))
    vf.append(os.path.join(args.dir, f"c{c:02d}.voca.txt"))
    bf.append(os.path.join(args.dir, f"c{c:02d}.best_path.npz"))
    np.savez(lf[-1], data=(rng.standard_normal((int(t), 39)) * 3).astype(np.float32), indices=np.array([t], np.int32))
    L = int(round(0.14 * t))
    ids = rng.integers(0, len(tokens), L)
    with open(vf[-1], "w") as f:
        for a in range(0, L, 8):
            f.write("w|" + " ".join(tokens[i] for i in ids[a:a + 8]) + "\n")
print(f"wrote {args.chapters} chapters, {int(T.sum())} frames in {time.perf_counter() - t0:.1f} s", file=sys.stderr)

out = {"chapters": args.chapters, "frames": int(T.sum()), "runs": []}
for mode in ("host_numpy", "device", "host_numpy", "device"):
    for f in bf:
        if os.path.exists(f):
            os.unlink(f)
    t = {}
    t0 = time.perf_counter()
    written = align.best_path_files(lf, vf, bf, verbose=False, device_log_softmax=(mode == "device"), timings=t)
    t["wall_s"] = time.perf_counter() - t0
    t["log_softmax"] = mode
    assert len(written) == args.chapters
    out["runs"].append(t)
    print(json.dumps(t), file=sys.stderr)

# the reference's loop on a bounded sample: shortest chapters up to --cpu-sample-frames
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctc_oracle_np  # noqa: E402  (bench tooling: the CPU baseline leg)
order = np.argsort(T)
frames = 0
t0 = time.perf_counter()
for c in order:
    if frames and frames + T[c] > args.cpu_sample_frames:
        break
    with np.load(lf[c]) as f:
        lp = align.log_softmax(f["data"])
    ctc_oracle_np.ctc_best_path(lp, align.read_transcript_labels(vf[c]))
    frames += int(T[c])
dt = time.perf_counter() - t0
out["cpu_reference"] = {"kind": "port", "cores": 1, "sample_frames": frames, "sample_s": dt,
                        "us_per_frame": dt / frames * 1e6, "book_s_extrapolated": dt / frames * int(T.sum())}
print(json.dumps(out, indent=1))
shutil.rmtree(args.dir, ignore_errors=True)
