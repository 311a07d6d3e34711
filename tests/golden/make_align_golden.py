"""Generate tests/golden/align_case/ with the REFERENCE's align() (align.py:127-169) and
ctc_best_path.  Build container only (needs /root/reference; fugashi is stubbed because
transcript.py imports the G2P module at import time and the tagger is not needed here).
    python tests/golden/make_align_golden.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
stub = types.ModuleType("fugashi")
stub.Tagger = lambda *a, **k: None
sys.modules["fugashi"] = stub

from kokoro_align import align as ref_align  # noqa: E402
from kokoro_align.encoder import vocab  # noqa: E402
from kokoro_align.transcript import read_transcript  # noqa: E402
from kokoro_align_b200 import synth  # noqa: E402

out = os.path.join(HERE, "align_case")
rng = np.random.default_rng(4242)
# a transcript of 60 "words": voiced tokens of 1-6 phonemes, some punctuation-only lines,
# some tokens with out-of-vocabulary symbols (q), one empty text
lines = []
for n in range(60):
    r = rng.random()
    if r < 0.12:
        lines.append((rng.choice(["、", "。", "！"]), rng.choice([",", ".", "!", "?"])))
    elif r < 0.16:
        lines.append(("っ", "q"))
    else:
        k = int(rng.integers(1, 7))
        toks = [vocab[int(rng.integers(1, len(vocab)))] for _ in range(k)]
        if rng.random() < 0.1:
            toks.insert(int(rng.integers(0, k + 1)), "q")
        lines.append((f"語{n}" if rng.random() > 0.05 else "", " ".join(toks)))
voca_file = os.path.join(out, "case.voca.txt")
with open(voca_file, "wt") as f:
    for t, v in lines:
        f.write(f"{t}|{v}\n")
labels = read_transcript(voca_file)
T = int(len(labels) / 0.14)
lp, _ = synth.make_lattice(T, len(labels), 39, seed=4243, planted=False)
# plant the true labels along a monotone path so that decoded text is meaningful
ext = np.zeros(2 * len(labels) + 1, np.int64)
ext[1::2] = labels
pos = np.minimum((np.arange(T) * len(ext)) // T, len(ext) - 1)
logits = rng.standard_normal((T, 39)).astype(np.float32)
logits[np.arange(T), ext[pos]] += 5.0
lp = synth.log_softmax_ref(logits).astype(np.float32)
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
    path, labs, scores = ref_align.ctc_best_path(lp, labels)
np.savez(os.path.join(out, "case.best_path.npz"), best_path=path, best_labels=labs, best_scores=scores)
cuts = np.sort(rng.choice(np.arange(40, T - 40), 11, replace=False))
indices = np.concatenate([cuts, [T]]).astype(np.int32)
np.savez(os.path.join(out, "case.mfcc.npz"), indices=indices, data=np.zeros((1, 1), np.float32))
np.savez_compressed(os.path.join(out, "case.log_probs.npz"), log_probs=lp, labels=labels.astype(np.int32))
for rw in (True, False):
    ref_align.align(os.path.join(out, "case.best_path.npz"), os.path.join(out, "case.mfcc.npz"), voca_file,
                    os.path.join(out, f"case.align.wordsep{int(not rw)}.txt"), rw)
print("wrote", sorted(os.listdir(out)))
