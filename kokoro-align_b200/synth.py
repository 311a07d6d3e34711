"""Synthetic CTC lattices of the BASELINE.json shapes (SURVEY.md section 8d recipe).

numpy only; used by tests/, bench.py and tests/golden/make_golden.py so that every side
sees byte-identical inputs.  The log-softmax is the reference's two lines
(kokoro_align/align.py:116-117) executed in fp32 numpy.
"""
import numpy as np

FRAME_RATE = 22050.0 / 256.0     # preprocess.py:102-115 -> 86.13 frames/s
LABELS_PER_FRAME = 0.14          # SURVEY.md 8(d) assumption: ~12 phonemes/s / 86.13 fps
VOCAB_SIZE = 39                  # encoder.py:11


def log_softmax_ref(logits):
    """align.py:116-117, verbatim arithmetic in fp32 numpy."""
    logits = logits - np.mean(logits, axis=-1, keepdims=True)
    return logits - np.log(np.sum(np.exp(logits), axis=-1, keepdims=True))


def label_dtype(V):
    return np.int8 if V <= 127 else np.int32   # encoder.py:19 produces int8


def make_lattice(T, L, V=VOCAB_SIZE, seed=0, planted=False, quant=None, repeat_labels=False):
    """One lattice: (log_probs f32 [T,V], labels int [L]).

    planted: add +6.0 on the true label of a random monotone path ending at S-1.
    quant:   round log-probs to multiples of 1/quant (tie stress).
    repeat_labels: draw labels from 3 symbols so that adjacent labels repeat often.
    """
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((T, V)).astype(np.float32)
    if repeat_labels:
        labels = rng.integers(1, min(4, V), L).astype(label_dtype(V))
    else:
        labels = rng.integers(1, V, L).astype(label_dtype(V))
    if planted and T > 0:
        S = 2 * L + 1
        ext = np.zeros(S, dtype=np.int64)
        ext[1::2] = labels
        jitter = rng.uniform(-0.5, 0.5, T)
        pos = np.floor((np.arange(T) + 1 + jitter) * S / T).astype(np.int64)
        pos = np.maximum.accumulate(np.clip(pos, 0, S - 1))
        pos[-1] = S - 1
        logits[np.arange(T), ext[pos]] += np.float32(6.0)
    log_probs = log_softmax_ref(logits)
    if quant:
        log_probs = (np.round(log_probs * np.float32(quant)) / np.float32(quant)).astype(np.float32)
    return np.ascontiguousarray(log_probs, dtype=np.float32), labels


def segment_lengths(B, seed, t_min=86, t_max=861):
    """Config 2: B silence-split segments of 1-10 s, T_b ~ U{86..861}, L_b = round(0.14 T_b)."""
    rng = np.random.default_rng(seed)
    T = rng.integers(t_min, t_max + 1, B).astype(np.int64)
    L = np.round(LABELS_PER_FRAME * T).astype(np.int64)
    return T, L


def chapter_lengths(C, total_T, seed):
    """Config 3: C chapter lattices, lengths ~ LogNormal(0, 0.5) normalised to total_T."""
    rng = np.random.default_rng(seed)
    w = rng.lognormal(0.0, 0.5, C)
    T = np.maximum(1, np.floor(w / w.sum() * total_T)).astype(np.int64)
    T[-1] += total_T - T.sum()
    L = np.round(LABELS_PER_FRAME * T).astype(np.int64)
    return T, L


def make_batch(T_list, L_list, V=VOCAB_SIZE, seed=0, **kw):
    """Flat batch in the C-ABI layout: log_probs [sum T, V], t_off, labels i32 [sum L], l_off.

    Lattice n is ``make_lattice(T_n, L_n, V, seed + n)``.
    """
    T_list = np.asarray(T_list, dtype=np.int64)
    L_list = np.asarray(L_list, dtype=np.int64)
    t_off = np.concatenate([[0], np.cumsum(T_list)]).astype(np.int64)
    l_off = np.concatenate([[0], np.cumsum(L_list)]).astype(np.int64)
    lp = np.empty((int(t_off[-1]), V), dtype=np.float32)
    labels = np.empty(int(l_off[-1]), dtype=np.int32)
    for n, (T, L) in enumerate(zip(T_list, L_list)):
        a, b = make_lattice(int(T), int(L), V, seed + n, **kw)
        lp[t_off[n]:t_off[n + 1]] = a
        labels[l_off[n]:l_off[n + 1]] = b
    return lp, t_off, labels, l_off


def make_batch_fast(T_list, L_list, V=VOCAB_SIZE, seed=0):
    """Same layout as make_batch but one RNG stream for the whole batch (bench-sized inputs:
    ~5 M frames in a few seconds).  Lattices are still independent iid-Gaussian logits."""
    T_list = np.asarray(T_list, dtype=np.int64)
    L_list = np.asarray(L_list, dtype=np.int64)
    t_off = np.concatenate([[0], np.cumsum(T_list)]).astype(np.int64)
    l_off = np.concatenate([[0], np.cumsum(L_list)]).astype(np.int64)
    rng = np.random.default_rng(seed)
    n = int(t_off[-1])
    lp = np.empty((n, V), dtype=np.float32)
    step = 1 << 18
    for a in range(0, n, step):
        b = min(n, a + step)
        lp[a:b] = log_softmax_ref(rng.standard_normal((b - a, V), dtype=np.float32))
    labels = rng.integers(1, V, int(l_off[-1])).astype(np.int32)
    return lp, t_off, labels, l_off


def make_lattice_exact(T, L, V=VOCAB_SIZE, seed=0, levels=64, scale=4.0, planted=False,
                       repeat_labels=False):
    """Platform-independent lattice: log-probs are -(integer in [0, levels)) / scale, exactly
    representable in fp32, no exp/log involved (so golden fixtures need to store only the
    recipe).  Coarse quantisation => massive exact ties (tie-break stress).  planted: the
    true label along a random monotone path gets 0.0 (the maximum)."""
    rng = np.random.default_rng(seed)
    lp = -(rng.integers(0, levels, (T, V)).astype(np.float32) / np.float32(scale))
    if repeat_labels:
        labels = rng.integers(1, min(4, V), L).astype(label_dtype(V))
    else:
        labels = rng.integers(1, V, L).astype(label_dtype(V))
    if planted and T > 0:
        S = 2 * L + 1
        ext = np.zeros(S, dtype=np.int64)
        ext[1::2] = labels
        jitter = rng.uniform(-0.5, 0.5, T)
        pos = np.floor((np.arange(T) + 1 + jitter) * S / T).astype(np.int64)
        pos = np.maximum.accumulate(np.clip(pos, 0, S - 1))
        pos[-1] = S - 1
        lp[np.arange(T), ext[pos]] = np.float32(0.0)
    return np.ascontiguousarray(lp, dtype=np.float32), labels
