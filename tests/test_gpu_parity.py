"""Parity of the CUDA path (through the C ABI) with the reference's golden vectors and with
the CPU oracle on seeded inputs.  Bit-exact for best_path / best_labels / best_scores;
final score within 1e-4 relative (north_star) -- asserted bit-equal in practice."""
import os

import numpy as np
import pytest

from tests.golden_util import check_case, golden_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kab():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from kokoro_align_b200 import align
    return align


@pytest.mark.parametrize("name", golden_names())
def test_golden(golden, kab, name):
    index, arrays = golden
    case = next(c for c in index if c["name"] == name)
    check_case(case, arrays, kab.ctc_best_path)


def _compare_batch(kab, lp, t_off, labels, l_off, beam_size=1000, max_move=4, V=39, threads=8):
    from oracle import ctc_oracle
    rp, rl, rs, rf, rst = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, beam_size,
                                                         max_move, n_threads=threads)
    with kab.AlignPlan(t_off, labels, l_off, V, beam_size, max_move) as plan:
        path, labs, scores, final, status = plan.run_host(lp)
        info = plan.info
    np.testing.assert_array_equal(status, rst)
    for b in range(len(t_off) - 1):
        if rst[b] != 0:
            continue
        a, e = int(t_off[b]), int(t_off[b + 1])
        np.testing.assert_array_equal(path[a:e], rp[a:e], err_msg=f"lattice {b}")
        np.testing.assert_array_equal(labs[a:e], rl[a:e], err_msg=f"lattice {b}")
        assert scores[a:e].tobytes() == rs[a:e].tobytes(), f"lattice {b}"
        assert abs(float(final[b]) - float(rf[b])) <= 1e-4 * max(1.0, abs(float(rf[b])))
        assert final[b].tobytes() == rf[b].tobytes(), f"lattice {b}"
    return info


def test_batch_short_segments(kab):
    """Config-2-shaped batch (warp kernel, all K classes), iid Gaussian log-softmax."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(300, seed=2000)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2001)
    info = _compare_batch(kab, lp, t_off, labels, l_off)
    assert info.n_class[0] == 300 and info.cells_eval == int((T * (2 * L + 1)).sum())


def test_batch_short_segments_ties(kab):
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(200, seed=2100, t_min=1, t_max=400)
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=2101)
    lp = (np.round(lp * 2) / 2).astype(np.float32)
    _compare_batch(kab, lp, t_off, labels, l_off)


def test_batch_mixed_classes(kab):
    """One batch that exercises warp, band (two ring sizes) and generic kernels at once,
    including a bad-label lattice and a dead band."""
    from kokoro_align_b200 import synth
    T = np.array([300, 2500, 50, 4000, 700, 90, 1200, 10, 640])
    L = np.array([40, 900, 5, 1500, 300, 100, 170, 40, 90])
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=3000, planted=True)
    labels = labels.copy()
    labels[l_off[4] + 7] = 0            # value-0 label -> generic kernel
    labels[l_off[6] + 3] = 39           # bad label -> status 2
    info = _compare_batch(kab, lp, t_off, labels, l_off, beam_size=600)
    assert info.n_class[0] >= 1 and info.n_class[1] >= 2 and info.n_class[2] >= 1


@pytest.mark.parametrize("beam_size,max_move", [(1000, 4), (200, 4), (64, 4), (1000, 3), (150, 6)])
def test_chapter_lattices(kab, beam_size, max_move):
    """Chapter-shaped banded lattices (band kernel for max_move 4, generic otherwise)."""
    from kokoro_align_b200 import synth
    T = np.array([12000, 7001, 3000])
    L = np.round(0.14 * T).astype(np.int64)
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=4000 + beam_size, planted=True)
    _compare_batch(kab, lp, t_off, labels, l_off, beam_size=beam_size, max_move=max_move)


def test_nonfinite_rejected(kab):
    from kokoro_align_b200 import synth
    for T, L in ((300, 40), (3000, 900)):
        lp, labels = synth.make_lattice(T, L, seed=5)
        lp[T // 2, 7] = -np.inf
        with pytest.raises(ValueError):
            kab.ctc_best_path(lp, labels)
    lp, labels = synth.make_lattice(100, 10, seed=6)
    lp[99, 38] = np.nan
    with pytest.raises(ValueError):
        kab.ctc_best_path(lp, labels, max_move=3)
    # the wide (unbanded) kernel and the single-CTA band kernel check too
    lp, labels = synth.make_lattice(900, 2000, seed=7)
    lp[899, 0] = np.inf
    with pytest.raises(ValueError):
        kab.ctc_best_path(lp, labels, beam_size=9000)
    import os
    os.environ["KAB_BAND_CLUSTER"] = "0"
    try:
        lp, labels = synth.make_lattice(3000, 900, seed=8)
        lp[17, 3] = -np.inf
        with pytest.raises(ValueError):
            kab.ctc_best_path(lp, labels)
    finally:
        del os.environ["KAB_BAND_CLUSTER"]


def test_device_resident_torch(kab):
    """Inputs already in HBM (torch tensors), asynchronous launch on the current stream."""
    import torch
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    T, L = synth.segment_lengths(64, seed=2200)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2201)
    rp, rl, rs, rf, rst = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, n_threads=8)
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        d_lp = torch.from_numpy(lp).cuda()
        path, labs, scores, final, status = plan.run_torch(d_lp)
        torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    np.testing.assert_array_equal(path.cpu().numpy(), rp)
    np.testing.assert_array_equal(labs.cpu().numpy(), rl)
    assert scores.cpu().numpy().tobytes() == rs.tobytes()
    assert final.cpu().numpy().tobytes() == rf.tobytes()


def test_full_size_properties(kab):
    """BASELINE config 1 shape (T = 81 135, L = 11 359): size-independent properties plus the
    C oracle (which finishes this shape in about a second)."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    T, L = 81135, 11359
    lp, labels = synth.make_lattice(T, L, seed=1000)
    path, labs, scores, final = kab.ctc_best_path(lp, labels, return_final_score=True)
    S = 2 * L + 1
    d = np.diff(path)
    assert path[0] in (0, 1, 3) and d.min() >= 0 and d.max() <= 3
    i = np.arange(T, dtype=np.int64)
    lo = np.maximum(0, S * i // T - 500)
    assert (path >= lo).all() and (path < np.minimum(lo + 1000, S)).all()
    ext = np.zeros(S, np.int32)
    ext[1::2] = labels
    np.testing.assert_array_equal(labs, ext[path])
    assert scores.tobytes() == lp[i, labs].tobytes()
    assert np.cumsum(scores, dtype=np.float32)[-1].tobytes() == np.float32(final).tobytes()
    rp, _, _, rf = ctc_oracle.ctc_best_path(lp, labels, return_final_score=True)
    np.testing.assert_array_equal(path, rp)
    assert np.float32(final).tobytes() == np.float32(rf).tobytes()


def test_pipelined_host_path(kab):
    """A batch large enough (>= 64 MB of log-probs) for kab_plan_run_host to cut it into
    segments and overlap H2D / kernels / D2H; results must not depend on the segmentation."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(2200, seed=2300)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2301)
    assert lp.nbytes >= 96 << 20
    labels = labels.copy()
    labels[l_off[1500] + 2] = 39            # one bad-label lattice in a later segment
    _compare_batch(kab, lp, t_off, labels, l_off)


def test_align_sharded_single_rank(kab):
    from kokoro_align_b200 import parallel, synth
    from oracle import ctc_oracle
    lps, labs = [], []
    for k, t in enumerate((90, 400, 3000, 700)):
        a, b = synth.make_lattice(t, int(round(0.14 * t)), 39, seed=2400 + k)
        lps.append(a)
        labs.append(b)
    res = parallel.align_sharded(lps, labs)
    for (path, _, _, final, status), lp, lab in zip(res, lps, labs):
        rp, _, _, rf = ctc_oracle.ctc_best_path(lp, lab, return_final_score=True)
        assert status == 0
        np.testing.assert_array_equal(path, rp)
        assert np.float32(final).tobytes() == np.float32(rf).tobytes()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_shapes_and_beams(kab, seed):
    """Randomised (T, L, beam_size) batches: every ring size of the band kernel (1..32 warps),
    class boundaries (S = 248/249, beam_size = 104k - 32 +- 1), tiny and degenerate lattices."""
    from kokoro_align_b200 import synth
    rng = np.random.default_rng(seed)
    edge_beams = [1, 2, 5, 71, 72, 73, 175, 176, 177, 1000, 3295, 3296, 3297, 5000]
    W = int(rng.choice(edge_beams)) if seed % 2 else int(rng.integers(1, 3400))
    n = 40
    T = rng.integers(1, 2500, n)
    ratio = rng.choice([0.02, 0.14, 0.5, 1.0, 1.45], n)
    L = np.minimum(np.maximum(0, np.round(ratio * T)).astype(np.int64), 3 * T)
    L[:4] = [123, 124, 0, 1]                       # S = 247 / 249 straddle the warp-class limit
    T[:4] = [900, 900, 1, 2]
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=7000 + 100 * seed, planted=bool(seed % 2))
    if seed % 3 == 0:                              # tie stress
        lp = (np.round(lp * 2) / 2).astype(np.float32)
    _compare_batch(kab, lp, t_off, labels, l_off, beam_size=W)


@pytest.mark.parametrize("beam_size,cluster,serial_bt,q,r", [
    (1000, 1, 0, 1, 1), (64, 1, 0, 1, 1), (200, 1, 0, 1, 1), (1200, 1, 0, 1, 1), (1, 1, 0, 1, 1),
    (1000, 1, 0, 1, 0), (64, 1, 0, 1, 0), (200, 1, 0, 1, 0), (2400, 1, 0, 1, 0), (1, 1, 0, 1, 0),
    (1000, 1, 0, 0, 0), (64, 1, 0, 0, 0), (200, 2, 0, 0, 0), (1000, 1, 1, 0, 0), (200, 2, 1, 0, 0),
    (1000, 0, 0, 0, 0), (64, 0, 0, 0, 0), (300, 0, 0, 0, 0)])
def test_both_band_kernels(kab, monkeypatch, beam_size, cluster, serial_bt, q, r):
    """The four band kernels against the C oracle, forced through KAB_BAND_CLUSTER / KAB_BAND_Q /
    KAB_BAND_R: kab_bandr.cuh (warp-specialised: one compute warp per scheduler, prep warps, mailboxes
    in shared memory; the default for up to 49 lattices; always the parallel traceback), kab_bandq.cuh
    (the same lanes, every warp does its own bookkeeping; KAB_BAND_R=0), kab_bandp.cuh (four states per
    lane; KAB_BAND_Q=0) with both tracebacks -- the parallel block-map composition (kab_btpar.cuh) and
    its own single-thread walker (KAB_BAND_SERIAL_BT=1) -- and the single-CTA kernel (kab_band.cuh,
    the default for larger batches, KAB_BAND_CLUSTER=0)."""
    from kokoro_align_b200 import synth
    monkeypatch.setenv("KAB_BAND_CLUSTER", str(cluster))
    monkeypatch.setenv("KAB_BAND_Q", str(q))
    monkeypatch.setenv("KAB_BAND_R", str(r))
    monkeypatch.setenv("KAB_BAND_SERIAL_BT", str(serial_bt))   # 0 forces the parallel traceback
    T = np.array([12000, 7001, 3000, 41, 5003])
    L = np.round(0.14 * T).astype(np.int64)
    L[4] = 1700  # S/T = 0.68: the walker changes warp regions often
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=4300 + beam_size, planted=True)
    info = _compare_batch(kab, lp, t_off, labels, l_off, beam_size=beam_size)
    assert info.n_class[1] >= 4


@pytest.mark.parametrize("seed,r", [(11, 0), (12, 0), (13, 0), (14, 0), (11, 1), (12, 1), (13, 1), (14, 1), (15, 1), (16, 1)])
def test_bandq_random_shapes(kab, monkeypatch, seed, r):
    """kab_bandq.cuh (r = 0) and kab_bandr.cuh (r = 1) forced on randomised (T, L, beam_size) batches:
    every cluster size 1..8, ring wrap-arounds (long T with narrow beams), S/T up to 3, tie stress,
    tiny lattices, more lattices than clusters (work queue)."""
    from kokoro_align_b200 import synth
    monkeypatch.setenv("KAB_BAND_Q", "1")
    monkeypatch.setenv("KAB_BAND_R", str(r))
    rng = np.random.default_rng(seed)
    wmax = 1248 if r else 2528          # widest band of the kernel: 40 * CW * 8 - 32 ring slots
    W = int(rng.choice([8, 9, 40, 41, 128, 129, 288, 289, 1000, wmax - 1, wmax])) if seed % 2 else int(rng.integers(1, wmax + 1))
    n = 24
    T = rng.integers(1, 4000, n)
    ratio = rng.choice([0.02, 0.14, 0.5, 1.0, 1.45], n)
    L = np.minimum(np.maximum(0, np.round(ratio * T)).astype(np.int64), (3 * T - 1) // 2)
    L[:3] = [200, 0, 1]
    T[:3] = [9000, 1, 2]
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=7700 + 100 * seed, planted=bool(seed % 2))
    if seed % 3 == 0:
        lp = (np.round(lp * 2) / 2).astype(np.float32)
    _compare_batch(kab, lp, t_off, labels, l_off, beam_size=W)


@pytest.mark.parametrize("V", [128, 256, 512, 600, 4096])
def test_wide_vocabularies(kab, V):
    """Vocabularies up to 512 columns run in the staged kernels with whole rows staged; wider ones
    (BASELINE config 5: V = 4096) on a compact copy of the columns the lattices use
    (kab_compact.cuh); all bit-exact against the C oracle, best_labels in original label values."""
    from kokoro_align_b200 import synth
    T = np.array([700, 90, 3000, 431])
    L = np.array([100, 12, 400, 60])
    lp, t_off, labels, l_off = synth.make_batch(T, L, V=V, seed=5100 + V, planted=True)
    info = _compare_batch(kab, lp, t_off, labels, l_off, beam_size=300, V=V)
    assert info.n_class[0] == 2 and info.n_class[1] == 2 and info.n_class[2] == 0


def test_wide_vocabulary_too_many_labels(kab, monkeypatch):
    """More than 511 distinct labels in one lattice: no compact copy is possible.  A band-shaped
    lattice then runs kab_bandr_kernel in its gather mode (emissions straight from the caller's
    log-probs), the short one next to it still goes through the compact copy; with the gather mode
    switched off the generic kernel takes the plan.  Bit-exact either way."""
    from kokoro_align_b200 import synth
    T = np.array([2500, 300])
    L = np.array([900, 40])
    lp, t_off, labels, l_off = synth.make_batch(T, L, V=4096, seed=5200)
    assert len(np.unique(labels[:900])) > 511
    info = _compare_batch(kab, lp, t_off, labels, l_off, V=4096)
    assert info.n_class[0] == 1 and info.n_class[1] == 1 and info.n_class[2] == 0
    monkeypatch.setenv("KAB_BAND_GATHER", "0")
    info = _compare_batch(kab, lp, t_off, labels, l_off, V=4096)
    assert info.n_class[2] == 2


@pytest.mark.parametrize("max_move", [4, 3, 2])
def test_wide_vocabulary_gather_mode_batch(kab, max_move):
    """BASELINE config 5 with a BPE-sized vocabulary: several chapter-like lattices with thousands
    of distinct labels (more lattices than resident clusters' worth of work-queue items), ties from
    quantised log-probs, edge windows at both ends, next to segment-sized lattices of the same
    vocabulary -- gather-mode band kernel + compact warp kernel in one plan."""
    from kokoro_align_b200 import synth
    T = np.array([9000, 150, 4001, 12000, 640, 2999, 8])
    L = np.array([2500, 20, 1300, 1500, 90, 1000, 1])
    lp, t_off, labels, l_off = synth.make_batch(T, L, V=4096, seed=5300 + max_move)
    lp = (np.round(lp * 4) / 4).astype(np.float32)
    assert len(np.unique(labels[:2500])) > 511
    info = _compare_batch(kab, lp, t_off, labels, l_off, beam_size=700, max_move=max_move, V=4096)
    assert info.n_class[1] == 4 and info.n_class[2] == 0


def test_wide_vocabulary_gather_mode_nonfinite_and_bad_label(kab):
    """Gather mode keeps the error contract: a non-finite value anywhere in a lattice's rows (also in
    a column it never reads) is status 3, an out-of-range label status 2, the others are aligned."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    T = np.array([3000, 2000, 2500])
    L = np.array([800, 600, 700])
    V = 1500
    lp, t_off, labels, l_off = synth.make_batch(T, L, V=V, seed=5400)
    lp[1234, 77 if 77 not in labels[:800] else 78] = np.nan
    labels = labels.copy()
    labels[800 + 10] = V + 3
    rst = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, 400, 4, n_threads=2)[4]
    with kab.AlignPlan(t_off, labels, l_off, V, 400) as plan:
        path, labs, scores, final, status = plan.run_host(lp)
        assert plan.info.n_class[1] == 2
    np.testing.assert_array_equal(status, rst)
    assert list(status) == [3, 2, 0]
    rp, rl, rs, rf, _ = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, 400, 4, n_threads=2)
    a = int(t_off[2])
    np.testing.assert_array_equal(path[a:], rp[a:])
    np.testing.assert_array_equal(labs[a:], rl[a:])
    assert scores[a:].tobytes() == rs[a:].tobytes() and final[2].tobytes() == rf[2].tobytes()


def test_hybrid_band_plan(kab, monkeypatch):
    """A multi-book job: hundreds of chapter-like lattices and a few much longer ones.  The plan gives
    the longest to kab_bandr_kernel clusters and the rest to the single-CTA kernel, side by side
    (KAB_BAND_KERNEL_HYBRID); the same batch through the single kernel alone and through a forced
    one-lattice split gives the same bits."""
    from kokoro_align_b200 import synth
    rng = np.random.default_rng(6100)
    T = np.concatenate([[30000, 28000, 26000], rng.integers(1500, 3501, 700)])
    L = np.maximum(1, np.round(0.14 * T)).astype(np.int64)
    perm = rng.permutation(len(T))                       # the long ones anywhere in the batch
    T, L = T[perm], L[perm]
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=6101)
    lp = (np.round(lp * 8) / 8).astype(np.float32)       # ties
    info = _compare_batch(kab, lp, t_off, labels, l_off)
    assert info.band_kernel == 5 and info.band_cluster == 7 and info.n_class[1] == len(T)
    # the same hybrid plan launched back to back on two streams, nothing synchronised in between (the
    # gate's counter only grows), and captured into a CUDA graph that is replayed: same results every time
    import torch
    from oracle import ctc_oracle
    rp = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, 1000, 4, n_threads=8)[0]
    d_lp = torch.from_numpy(lp).cuda()
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        outs = [plan.run_torch(d_lp, stream=(s1 if k % 2 else s2)) for k in range(4)]
        torch.cuda.synchronize()
        for o in outs:
            assert (o[4].cpu().numpy() == 0).all()
            np.testing.assert_array_equal(o[0].cpu().numpy(), rp)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s1):
            o = plan.run_torch(d_lp, stream=s1)
        for _ in range(3):
            o[0].zero_()
            g.replay()
            torch.cuda.synchronize()
            np.testing.assert_array_equal(o[0].cpu().numpy(), rp)
    monkeypatch.setenv("KAB_BAND_HYBRID", "0")
    info = _compare_batch(kab, lp, t_off, labels, l_off)
    assert info.band_kernel == 1
    monkeypatch.setenv("KAB_BAND_HYBRID", "1")
    n = 40                                               # a small plan, split by force; max_move 3
    info = _compare_batch(kab, lp[:int(t_off[n])], t_off[:n + 1], labels[:int(l_off[n])], l_off[:n + 1], max_move=3)
    assert info.band_kernel in (4, 5)


def test_config4_banded_million_frames(kab):
    """BASELINE config 4(i): ONE lattice of T = 10^6 frames, L = 50 000 labels, the default
    1000-wide band (10^9 evaluated cells), bit-exact against the C oracle; then the same batch
    entry with a second, shorter lattice behind it (offsets beyond 2^31 bytes of log-probs are
    not reached here, but T * V * 4 = 156 MB exercises the 64-bit row arithmetic)."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    T, L = 1_000_000, 50_000
    lp, t_off, labels, l_off = synth.make_batch_fast(np.array([T, 4097]), np.array([L, 600]), seed=4004)
    rp, rl, rs, rf, rst = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, n_threads=2)
    assert (rst == 0).all()
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        path, labs, scores, final, status = plan.run_host(lp)
        assert plan.info.n_class[1] == 2
    assert (status == 0).all()
    np.testing.assert_array_equal(path, rp)
    np.testing.assert_array_equal(labs, rl)
    assert scores.tobytes() == rs.tobytes() and final.tobytes() == rf.tobytes()
    d = np.diff(path[:T])
    assert d.min() >= 0 and d.max() <= 3


def test_log_probs_beyond_2_gib(kab):
    """Maximum sizes: a batch whose log-probs exceed 2^31 bytes (30 000 segments, 14 M frames,
    2.2 GB) with a band lattice BEHIND the 2 GiB mark: every byte offset of the warp / band
    kernels, the staging descriptors and the pipelined host path must be 64-bit.  Bit-exact
    against the C oracle."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(30000, seed=6100)
    T = np.concatenate([T, [20011]])
    L = np.concatenate([L, [2801]])
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=6101)
    assert int(t_off[-2]) * 39 * 4 > 2 ** 31
    info = _compare_batch(kab, lp, t_off, labels, l_off, threads=16)
    assert info.n_class[0] == 30000 and info.n_class[1] == 1


def test_best_path_files_batch(kab, tmp_path, capsys):
    """The per-book loop (run_example.py:247-254) as one batch: same npz files as the per-file
    best_path(), existing outputs skipped with the reference's message."""
    rng = np.random.default_rng(77)
    logits_files, voca_files, out_files, ref_files = [], [], [], []
    for n, (T, nl) in enumerate(((900, 9), (2600, 40), (130, 1), (5200, 70), (300, 3))):
        lf, vf = tmp_path / f"c{n}.logits.npz", tmp_path / f"c{n}.voca.txt"
        np.savez(lf, data=rng.standard_normal((T, 39)).astype(np.float32) * 3, indices=np.array([T], np.int32))
        with open(vf, "w") as f:
            for k in range(nl):
                f.write(f"text {k}|k o k o r o , w a t a sh i\\n")
        logits_files.append(str(lf)); voca_files.append(str(vf))
        out_files.append(str(tmp_path / f"c{n}.best_path.npz")); ref_files.append(str(tmp_path / f"c{n}.ref.npz"))
    for lf, vf, rf in zip(logits_files, voca_files, ref_files):
        kab.best_path(lf, vf, rf)
    np.savez(out_files[2], best_path=np.zeros(1, np.int32))  # already there: must be skipped
    t = {}
    written = kab.best_path_files(logits_files, voca_files, out_files, timings=t)
    out = capsys.readouterr().out
    want = [out_files[k] for k in (0, 1, 3, 4)]
    assert written == want and t["groups"] == [2, 2]    # the two long chapters / the two short ones
    assert f"Skip writing {out_files[2]}" in out
    assert [ln for ln in out.splitlines() if ln.startswith("Writing")] == [f"Writing {w}" for w in want]
    for k in (0, 1, 3, 4):
        with np.load(out_files[k]) as a, np.load(ref_files[k]) as b:
            assert sorted(a.files) == ["best_labels", "best_path", "best_scores"]
            for key in a.files:
                assert a[key].dtype == b[key].dtype and a[key].tobytes() == b[key].tobytes()
    # one plan for everything gives the same files
    for k in (0, 1, 3, 4):
        os.unlink(out_files[k])
    assert kab.best_path_files(logits_files, voca_files, out_files, verbose=False, pipeline=False) == want
    for k in (0, 1, 3, 4):
        with np.load(out_files[k]) as a, np.load(ref_files[k]) as b:
            for key in a.files:
                assert a[key].tobytes() == b[key].tobytes()


@pytest.mark.parametrize("T,L", [(3000, 2500), (1003, 1500), (5, 300), (2047, 4000)])
def test_wide_unbanded_lattices(kab, T, L):
    """Unbanded lattices wider than one CTA (beam_size >= 2S: BASELINE config 4(ii) shape, small):
    the chain-of-warps kernel (kab_wide.cuh) against the C oracle, bit for bit; a second lattice
    behind the first one exercises the lattice loop and the backtrack CTA's hand-over."""
    from kokoro_align_b200 import synth
    Ts = np.array([T, max(3, T // 2), 700])
    Ls = np.array([L, L + 37, 2000])
    lp, t_off, labels, l_off = synth.make_batch(Ts, Ls, seed=6000 + T, planted=(T > 100))
    beam = 2 * (2 * int(Ls.max()) + 1) + 2
    info = _compare_batch(kab, lp, t_off, labels, l_off, beam_size=beam)
    assert info.n_class[3] >= 2 and info.n_class[2] == 0


def test_memory_pool_reuse_and_trim(kab):
    """Plans built one after the other reuse the device blocks of destroyed ones (dirty memory:
    nothing may rely on zero-filled allocations), kab_pool_trim() returns the cache to the driver,
    and a plan destroyed right after an ASYNCHRONOUS run does not hand its workspace to the next
    plan while its kernels are still in flight."""
    import torch
    from kokoro_align_b200 import _lib, synth
    from oracle import ctc_oracle
    shapes = [(5000, 700), (300, 42), (5000, 700), (9000, 1260), (300, 42), (5000, 700)]
    outs = []
    for k, (T, L) in enumerate(shapes):
        lp, labels = synth.make_lattice(T, L, 39, seed=8800 + k)
        plan = kab.AlignPlan([0, T], labels, [0, L], 39)
        outs.append((plan.run_torch(torch.from_numpy(lp).cuda()), lp, labels))
        plan.close()                                   # asynchronous work may still be running
        if k == 3:
            assert _lib.lib().kab_pool_trim() == 0
    torch.cuda.synchronize()
    for (path, labs, scores, final, status), lp, labels in outs:
        rp, rl, rs, rf = ctc_oracle.ctc_best_path(lp, labels, return_final_score=True)
        assert int(status[0]) == 0
        np.testing.assert_array_equal(path.cpu().numpy(), rp)
        assert scores.cpu().numpy().tobytes() == rs.tobytes()
