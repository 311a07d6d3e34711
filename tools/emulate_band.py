"""Thread-level numpy emulation of kab_band_kernel's ring / recycle / mask / backpointer logic
(kokoro-align_b200/csrc/kab_band.cuh), checked against the C oracle.  Development aid for a
container without a GPU: validates the index arithmetic, not the CUDA code itself."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctc_oracle  # noqa: E402
from kokoro_align_b200 import synth  # noqa: E402

NINF = np.float32(-np.inf)


def blank(s0, s1, s3, e):
    a0, a1, a3 = s0 + e, s1 + e, s3 + e
    p1 = a1 > a0
    m = np.where(p1, a1, a0)
    p3 = a3 > m
    return np.where(p3, a3, m), np.where(p3, 3, np.where(p1, 1, 0))


def label(s0, s1, s2, s3, e):
    a0, a1, a2, a3 = s0 + e, s1 + e, s2 + e, s3 + e
    p01, p23 = a1 > a0, a3 > a2
    m01, m23 = np.where(p01, a1, a0), np.where(p23, a3, a2)
    ph = m23 > m01
    return np.where(ph, m23, m01), np.where(ph, np.where(p23, 3, 2), np.where(p01, 1, 0))


def band_emulate(lp, labels, W, NT):
    """Event-driven variant (kab_band.cuh v2): threads know nothing about the window except at
    their own event frames, where they recompute lo/hi exactly, recycle their chunk, refresh
    the per-cell caps (+inf inside the window, -inf outside) and schedule their next event."""
    T, V = lp.shape
    L = len(labels)
    S = 2 * L + 1
    R = 4 * NT
    assert min(W, S) + 12 <= R and S <= 3 * T
    col = np.concatenate([labels.astype(np.int64), np.zeros(R + 16, np.int64)])
    tid = np.arange(NT)
    prev = np.full(R, NINF, np.float32)
    prev[0] = 0
    vb = 4 * tid
    half = W // 2
    caps = np.full((NT, 4), NINF, np.float32)
    next_event = np.zeros(NT, np.int64)
    BIG = 1 << 62
    n_events = 0

    def cols(base):
        c1 = np.where(base + 1 < S, col[np.minimum(base >> 1, len(col) - 2)], 0)
        c3 = np.where(base + 3 < S, col[np.minimum((base >> 1) + 1, len(col) - 1)], 0)
        return c1, c3

    def handle_event(t, i):
        lo = max(0, S * i // T - half)
        hi = min(lo + W, S)
        while vb[t] + 3 < lo - 3:
            vb[t] += R
        for k in range(4):
            v = vb[t] + k
            caps[t, k] = np.float32(np.inf) if lo <= v < hi else NINF
        # next lo value at which this thread's pattern changes
        th = [vb[t] + 1, vb[t] + 2, vb[t] + 3, vb[t] + 4, vb[t] + 7]
        th += [vb[t] + k - W + 1 for k in range(4) if vb[t] + k < S]
        th = [x for x in th if x > lo]
        if not th:
            return BIG
        Lstar = min(th)
        # first frame with lo_i >= Lstar  <=>  floor(S*i/T) >= Lstar + half
        need = Lstar + half
        return -((-need * T) // S)   # ceil(need*T/S)

    c1, c3 = cols(vb)
    bp = np.zeros((T, NT), np.uint8)
    with np.errstate(invalid="ignore"):
        for i in range(T):
            for t in np.nonzero(next_event == i)[0]:
                next_event[t] = handle_event(t, i)
                assert next_event[t] > i
                n_events += 1
            c1, c3 = cols(vb)
            row = lp[i]
            eb, e1, e3 = row[0], row[c1], row[c3]
            P = prev.reshape(NT, 4)
            H = prev[((4 * tid + R - 4) % R)[:, None] + np.arange(4)[None, :]]
            n0, m0 = blank(P[:, 0], H[:, 3], H[:, 1], eb)
            n1, m1 = label(P[:, 1], P[:, 0], H[:, 3], H[:, 2], e1)
            n2, m2 = blank(P[:, 2], P[:, 1], H[:, 3], eb)
            n3, m3 = label(P[:, 3], P[:, 2], P[:, 1], P[:, 0], e3)
            N = np.stack([n0, n1, n2, n3], 1).astype(np.float32)
            N = np.minimum(N, caps)
            bp[i] = m0 | (m1 << 2) | (m2 << 4) | (m3 << 6)
            prev = N.reshape(-1).copy()
    states = vb[:, None] + np.arange(4)[None, :]
    ok = (states < S) & (prev.reshape(NT, 4) > NINF)
    if not ok.any():
        raise ValueError("dead")
    v = int(states[ok].max())
    final = prev[v % R]
    path = np.empty(T, np.int32)
    for i in range(T - 1, -1, -1):
        slot = v % R
        path[i] = v
        v -= (int(bp[i, slot >> 2]) >> (2 * (slot & 3))) & 3
    band_emulate.events = n_events
    return path, final


def main():
    rng = np.random.default_rng(0)
    n = 0
    for trial in range(400):
        NT = int(rng.choice([4, 8, 16, 32]))
        R = 4 * NT
        W = int(rng.integers(1, R - 12 + 1))
        T = int(rng.integers(1, 400))
        L = int(rng.integers(0, min(3 * T - 1, 600) // 2 + 1))
        S = 2 * L + 1
        if S > 3 * T or min(W, S) + 12 > R:
            continue
        lp, labels = synth.make_lattice_exact(T, L, 39, seed=trial, levels=int(rng.choice([2, 4, 64])),
                                              planted=bool(rng.integers(0, 2)))
        try:
            rp, _, _, rf = ctc_oracle.ctc_best_path(lp, labels, W, 4, return_final_score=True)
            ref = (rp, rf)
        except ValueError:
            ref = None
        try:
            got = band_emulate(lp, labels, W, NT)
        except ValueError:
            got = None
        assert (ref is None) == (got is None), (trial, NT, W, T, L)
        if ref is not None:
            assert np.array_equal(ref[0], got[0]), (trial, NT, W, T, L)
            assert np.float32(ref[1]).tobytes() == np.float32(got[1]).tobytes()
        n += 1
    print(f"band ring emulation (event-driven) == oracle on {n} random cases")


if __name__ == "__main__":
    main()
