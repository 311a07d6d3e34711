"""Lane-level numpy emulation of the NEXT band-kernel layout:
TWO states per lane (one blank + one label), 12 ghost lanes per warp that recompute the previous
warp's top 24 states for a group of 8 frames, 20 owned lanes = 40 owned ring slots per warp, the
3-state halo from lanes l-1 and l-2 (two SHFL.UP), one neighbour exchange per group.

Why this layout: the timing build of kab_bandp_kernel shows that a lone warp per scheduler issues
the ~400 instructions of an 8-frame group at ~2 cycles each (half-rate ALU / FMA pipes), i.e. the
frame block is issue bound for that warp while each pipe idles half of the time.  Half the states
per lane halves the instructions on every warp's chain, and two such warps per scheduler overlap
one warp's ALU phase with the other's FMA phase.  The price is more redundancy (12 of 32 lanes are
ghosts instead of 6) and a longer chain of warps.

This file checks the SCHEME -- ring slots, recycling between groups, window masks, ghost-lane
junk never reaching an owned state within a group, 2-bit backpointer bytes -- against the C
oracle, bit for bit, with adversarial junk in the halo of lane 0.  It is a development aid for a
container without a GPU; no product code depends on it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctc_oracle  # noqa: E402
from kokoro_align_b200 import synth  # noqa: E402
from tools.emulate_band import blank, label, NINF  # noqa: E402

G = 8                  # frames per group
K = 2                  # states per lane: k = 0 blank (even state), k = 1 label
GHOST = 3 * G // K     # 12 ghost lanes: junk climbs <= 3 states per frame, 24 states per group
OWN = 32 - GHOST       # 20 owned lanes
OW = OWN * K           # 40 owned ring slots per warp


def emulate(lp, labels, W, NW, junk=None):
    T, V = lp.shape
    L = len(labels)
    S = 2 * L + 1
    R = OW * NW
    assert R % 2 == 0 and min(W, S) + 32 <= R and S <= 3 * T
    col = np.concatenate([labels.astype(np.int64), np.zeros(R + 64, np.int64)])
    half = W // 2
    w_id = np.arange(NW)[:, None]
    lane = np.arange(32)[None, :]
    # ring slot of (warp, lane, k=0): owned lanes 12..31 -> 40w + 2(l-12); ghost lanes 0..11 mirror
    # the previous warp's lanes 20..31 (its top 24 slots)
    slot0 = np.where(lane >= GHOST, OW * w_id + K * (lane - GHOST), (OW * w_id - K * GHOST + K * lane) % R) % R
    vb = slot0.copy()             # alias level 0: state == slot
    s = np.full((NW, 32, K), NINF, np.float32)
    s[0, GHOST, 0] = 0.0          # virtual start: state 0 (align.py:57-58)
    rng = np.random.default_rng(1)
    bp = np.zeros((T, R // 4), np.uint8)   # 4 states (two lanes) per byte, 2 bits each

    with np.errstate(invalid="ignore"):
        for i in range(T):
            lo = max(0, S * i // T - half)
            hi = min(lo + W, S)
            if i % G == 0:     # recycling happens between groups only (32 spare ring slots)
                rec = vb + (K - 1) < lo - 3
                while rec.any():
                    vb = np.where(rec, vb + R, vb)
                    rec = vb + (K - 1) < lo - 3
            c1 = np.where(vb + 1 < S, col[np.minimum(vb >> 1, len(col) - 1)], 0)
            row = lp[i]
            eb, e1 = row[0], row[c1]
            # two SHFL.UP: lane l-1 (states -1, -2) and lane l-2 (state -3); lanes 0 / 1 keep junk
            up1 = np.concatenate([s[:, :1, :], s[:, :-1, :]], axis=1)
            up2 = np.concatenate([s[:, :2, :], s[:, :-2, :]], axis=1)
            if junk is not None:   # make the junk explicit: any finite value must be harmless
                up1[:, 0, :] = junk(rng, up1[:, 0, :].shape)
                up2[:, :2, :] = junk(rng, up2[:, :2, :].shape)
            h1, h2, h3 = up1[:, :, 1], up1[:, :, 0], up2[:, :, 1]
            n0, m0 = blank(s[:, :, 0], h1, h3, eb)                  # blank state: moves 0, 1, 3
            n1, m1 = label(s[:, :, 1], s[:, :, 0], h1, h2, e1)      # label state: moves 0 .. 3
            N = np.stack([n0, n1], 2).astype(np.float32)
            st = vb[:, :, None] + np.arange(K)[None, None, :]
            N = np.where((st < lo) | (st >= hi), NINF, N)
            code = (m0 | (m1 << 2)).astype(np.uint8)                # 4 bits per lane and frame
            own = np.broadcast_to(lane >= GHOST, code.shape)
            sl = slot0[own]
            np.bitwise_or.at(bp[i], sl >> 2, code[own] << (2 * (sl & 3)))
            s = N
            if (i + 1) % G == 0:   # group boundary: owners publish their top 12 lanes, ghosts reload
                top = s[:, 32 - GHOST:, :].copy()
                s[:, :GHOST, :] = np.roll(top, 1, axis=0)           # warp w takes warp w-1's top lanes
    own3 = np.broadcast_to((lane >= GHOST)[:, :, None], s.shape)
    states = vb[:, :, None] + np.arange(K)[None, None, :]
    ok = own3 & (states < S) & (s > NINF)
    if not ok.any():
        raise ValueError("dead")
    v = int(states[ok].max())
    final = s[ok & (states == v)][0]
    path = np.empty(T, np.int32)
    slot = v % R
    for i in range(T - 1, -1, -1):
        path[i] = v
        mv = (int(bp[i, slot >> 2]) >> (2 * (slot & 3))) & 3
        v -= mv
        slot -= mv
        if slot < 0:
            slot += R
    return path, final


def main():
    rng = np.random.default_rng(0)
    n = 0
    for trial in range(300):
        NW = int(rng.choice([2, 3, 5, 8, 26]))
        R = OW * NW
        W = int(rng.integers(1, R - 32 + 1))
        T = int(rng.integers(1, 500))
        L = int(rng.integers(0, min(3 * T - 1, 900) // 2 + 1))
        S = 2 * L + 1
        if S > 3 * T or min(W, S) + 32 > R:
            continue
        lp, labels = synth.make_lattice_exact(T, L, 39, seed=trial, levels=int(rng.choice([2, 4, 64])),
                                              planted=bool(rng.integers(0, 2)))
        try:
            rp, _, _, rf = ctc_oracle.ctc_best_path(lp, labels, W, 4, return_final_score=True)
            ref = (rp, rf)
        except ValueError:
            ref = None
        for junk in (None, lambda r, shp: r.uniform(-5, 50, shp).astype(np.float32)):
            try:
                got = emulate(lp, labels, W, NW, junk)
            except ValueError:
                got = None
            assert (ref is None) == (got is None), (trial, NW, W, T, L)
            if ref is not None:
                assert np.array_equal(ref[0], got[0]), (trial, NW, W, T, L)
                assert np.float32(ref[1]).tobytes() == np.float32(got[1]).tobytes()
        n += 1
    print(f"two-states-per-lane ghost scheme == oracle on {n} random cases (with adversarial junk halos)")


if __name__ == "__main__":
    main()
