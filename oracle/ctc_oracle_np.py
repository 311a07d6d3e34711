"""Per-frame numpy restatement of the reference's CTC best-path -- TEST INFRASTRUCTURE ONLY.

Follows kaiidams/Kokoro-Align ``kokoro_align/align.py:43-109`` (``ctc_best_path``) and the
traceback of ``align.py:21-40`` frame by frame with small numpy calls, i.e. it has the same
cost structure as the reference (a Python loop over T frames, a handful of vectorised ops on
<= beam_size elements per frame).  ``bench.py`` times this as the reference's CPU path
(``cpu_baseline.kind == "port"``); the tests use it as a second oracle beside the C one.

Parity is PINNED by tests/test_oracle_golden.py against tests/golden/ (vectors produced by
the reference's own function, see tests/golden/make_golden.py).

Differences from the reference, all deliberate and documented in DESIGN.md:
  * dense per-frame window instead of a compacted active list (same results: an inactive
    state is a -inf score, valid because log-probs are required to be finite);
  * the full backpointer table is kept and walked once from the forced end state instead of
    the incremental ``flush_determined_path`` (identical result);
  * errors are the reference's exception types: ValueError (dead band, align.py:101),
    IndexError (label outside [-V, V), align.py:77); non-finite log-probs raise ValueError.
"""
import numpy as np


def band(S, i, T, W):
    """Window [lo, hi) of target states at frame i.  align.py:64-65 (Python big ints)."""
    lo = max(0, S * i // T - W // 2)
    hi = min(lo + W, S)
    return lo, max(hi, lo)


def cells_eval(T, L, beam_size=1000):
    """sum_i (hi_i - lo_i): the cells the reference actually evaluates."""
    S = 2 * L + 1
    i = np.arange(T, dtype=np.int64)
    lo = np.maximum(0, S * i // T - beam_size // 2)
    hi = np.maximum(np.minimum(lo + beam_size, S), lo)
    return int((hi - lo).sum())


def ctc_best_path(log_probs, labels, beam_size=1000, max_move=4, return_final_score=False):
    log_probs = np.asarray(log_probs)
    if log_probs.dtype != np.float32:
        log_probs = log_probs.astype(np.float32)
    T, V = log_probs.shape
    labels = np.asarray(labels)
    M, W = int(max_move), int(beam_size)
    if T == 0:
        raise IndexError("list index out of range")  # beams[-1] on an empty list, align.py:100
    if M < 1:
        raise ValueError("attempt to get argmax of an empty sequence")
    if labels.size and (labels.min() < -V or labels.max() >= V):
        raise IndexError(f"label out of bounds for axis 1 with size {V}")
    if not np.isfinite(log_probs).all():
        raise ValueError("log_probs must be finite")

    # Expand label with blanks.  align.py:46-48
    ext = np.zeros(labels.shape[0] * 2 + 1, dtype=np.int32)
    ext[1::2] = labels
    S = ext.shape[0]
    is_zero = ext == 0
    pad = M - 1

    # prev holds scores of states [plo - pad, phi); -inf == inactive.  Virtual start:
    # state 0 active with score 0 before frame 0.  align.py:57-58
    plo, phi = 0, 1
    prev = np.full(pad + 1, -np.inf, dtype=np.float32)
    prev[pad] = 0.0
    moves, los = [], np.empty(T, dtype=np.int64)
    neg_inf = np.float32(-np.inf)

    for i in range(T):
        lo, hi = band(S, i, T, W)
        n = hi - lo
        los[i] = lo
        # source scores for states [lo - pad, hi) taken from the previous frame's window
        src = np.full(n + pad, neg_inf, dtype=np.float32)
        a, b = max(lo - pad, plo - pad), min(hi, phi)
        if b > a:
            src[a - (lo - pad):b - (lo - pad)] = prev[a - (plo - pad):b - (plo - pad)]
        e = log_probs[i, ext[lo:hi]]                      # align.py:77 (emission gather)
        cand = np.empty((M, n), dtype=np.float32)
        for j in range(M):                                # align.py:70
            cand[j] = src[pad - j:pad - j + n] + e        # one fp32 add per candidate
            if j > 0 and j % 2 == 0:                      # align.py:80-81
                cand[j, is_zero[lo:hi]] = neg_inf
        k = np.argmax(cand, axis=0)                       # first max => smallest j, :83
        best = np.take_along_axis(cand, k[None, :], axis=0)[0]
        moves.append(k.astype(np.uint8))
        prev = np.full(n + pad, neg_inf, dtype=np.float32)
        prev[pad:] = best
        plo, phi = lo, hi

    active = np.nonzero(prev[pad:] > neg_inf)[0]
    if active.size == 0:                                  # align.py:101
        raise ValueError("attempt to get argmax of an empty sequence")
    v = int(active[-1]) + plo                             # highest active state, :99-101
    final_score = prev[pad + v - plo]
    best_path = np.empty(T, dtype=np.int32)
    for i in range(T - 1, -1, -1):                        # == align.py:21-40
        best_path[i] = v
        v -= int(moves[i][v - los[i]])
    best_labels = ext[best_path]                          # align.py:106
    best_scores = log_probs[np.arange(T), best_labels]    # align.py:107
    if return_final_score:
        return best_path, best_labels, best_scores, np.float32(final_score)
    return best_path, best_labels, best_scores
