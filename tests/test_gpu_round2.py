"""Round-2 GPU tests: parity at the EXACT BASELINE shapes bench.py runs (configs 2, 3, 4(ii)-small),
the one-shot C entry the reference-side binding calls, plan life-cycle hazards, the device-side
consumer (segment records) and the encoder hand-off (rank 4).
Everything goes through the C ABI; the oracle is only the checker."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "align_case")


@pytest.fixture(scope="module")
def kab():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from kokoro_align_b200 import align
    return align


def _oracle(lp, t_off, labels, l_off, beam_size=1000, max_move=4, threads=None):
    from oracle import ctc_oracle
    threads = threads or min(32, os.cpu_count() or 8)
    return ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, beam_size, max_move, n_threads=threads)


def _assert_same(got, ref, t_off):
    path, labs, scores, final, status = got
    rp, rl, rs, rf, rst = ref
    np.testing.assert_array_equal(status, rst)
    assert (rst == 0).all()
    np.testing.assert_array_equal(path, rp)
    np.testing.assert_array_equal(labs, rl)
    assert scores.tobytes() == rs.tobytes()
    assert final.tobytes() == rf.tobytes()


# ------------------------------------------------------------------ exact BASELINE shapes
def test_config2_full_10000_segments(kab):
    """BASELINE config 2 exactly as bench.py builds it (10 000 segments, seed 2000 / 2001), every
    lattice against the C oracle, bit for bit."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(10000, 2000)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2001)
    ref = _oracle(lp, t_off, labels, l_off)
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        got = plan.run_host(lp)
        assert plan.info.n_class[0] == 10000 and plan.info.cells_eval == 777417790
    _assert_same(got, ref, t_off)


@pytest.mark.parametrize("serial_bt", [0, 1])
def test_config3_book_36_chapters(kab, monkeypatch, serial_bt):
    """BASELINE config 3 exactly as bench.py builds it (36 log-normal chapter lattices, sum T =
    2 721 800): the plan-level choice 'one cluster per chapter + parallel traceback' with 36
    unequal lattices, and the in-kernel walkers, against the C oracle."""
    from kokoro_align_b200 import synth
    monkeypatch.setenv("KAB_BAND_SERIAL_BT", str(serial_bt))
    T, L = synth.chapter_lengths(36, 2721800, 2000)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=2001)
    ref = _oracle(lp, t_off, labels, l_off)
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        got = plan.run_host(lp)
        assert plan.info.n_class[1] == 36
    _assert_same(got, ref, t_off)


def test_wide_unbanded_4e8_cells(kab):
    """BASELINE config 4(ii) / config 5 unbanded point T = S ~ 2e4 (4e8 cells) against the C oracle."""
    from kokoro_align_b200 import synth
    T, L = 20000, 10000
    lp, labels = synth.make_lattice(T, L, seed=4402)
    t_off, l_off = np.array([0, T]), np.array([0, L])
    ref = _oracle(lp, t_off, labels.astype(np.int32), l_off, beam_size=2 * (2 * L + 1) + 2, threads=1)
    with kab.AlignPlan(t_off, labels, l_off, 39, beam_size=2 * (2 * L + 1) + 2) as plan:
        got = plan.run_host(lp)
        assert plan.info.n_class[3] == 1
    _assert_same(got, ref, t_off)


# ------------------------------------------------------------------ the one-shot C entry
def _c_best_path(lp, labels, beam_size=1000, max_move=4):
    """Exactly the binding INTEGRATION.md shows for kab_ctc_best_path."""
    from kokoro_align_b200 import _lib
    L = _lib.lib()
    lp = np.ascontiguousarray(lp, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    T, V = lp.shape
    path, labs = np.empty(T, np.int32), np.empty(T, np.int32)
    scores, final, status = np.empty(T, np.float32), ctypes.c_float(), ctypes.c_int32(-7)
    rc = L.kab_ctc_best_path(lp.ctypes.data_as(ctypes.c_void_p), T, V, labels.ctypes.data_as(ctypes.c_void_p),
                             len(labels), beam_size, max_move, path.ctypes.data_as(ctypes.c_void_p),
                             labs.ctypes.data_as(ctypes.c_void_p), scores.ctypes.data_as(ctypes.c_void_p),
                             ctypes.byref(final), ctypes.byref(status))
    return rc, status.value, path, labs, scores, final.value


def test_kab_ctc_best_path_entry(kab):
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    for T, L, W in ((431, 60, 1000), (3000, 420, 1000), (900, 126, 64)):
        lp, labels = synth.make_lattice(T, L, seed=9100 + T)
        rc, st, path, labs, scores, final = _c_best_path(lp, labels, W)
        rp, rl, rs, rf = ctc_oracle.ctc_best_path(lp, labels, W, 4, return_final_score=True)
        assert rc == 0 and st == 0
        np.testing.assert_array_equal(path, rp)
        np.testing.assert_array_equal(labs, rl)
        assert scores.tobytes() == rs.tobytes() and np.float32(final).tobytes() == np.float32(rf).tobytes()
    lp, labels = synth.make_lattice(40, 100, seed=9200)          # dead band: status 1
    assert _c_best_path(lp, labels, 20)[:2] == (0, 1)
    lp, labels = synth.make_lattice(300, 40, seed=9201)
    bad = labels.astype(np.int32).copy()
    bad[5] = 39                                                   # label out of range: status 2
    assert _c_best_path(lp, bad)[:2] == (0, 2)
    lp[17, 3] = np.nan                                            # non-finite: status 3
    assert _c_best_path(lp, labels)[:2] == (0, 3)
    from kokoro_align_b200 import _lib
    assert _lib.lib().kab_ctc_best_path(None, 10, 39, None, 0, 1000, 4, None, None, None, None, None) == _lib.KAB_E_BAD_ARG


def test_integration_md_binding_runs_as_written(kab):
    """The ctypes stub of INTEGRATION.md section 2, executed VERBATIM (only the library path is made
    absolute): results equal the reference's golden vectors, statuses become its exceptions."""
    import re
    from kokoro_align_b200 import _lib, synth
    from tests.golden_util import load_golden, case_inputs
    text = open(os.path.join(os.path.dirname(HERE), "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "kab_ctc_best_path.argtypes" in b)
    _lib.lib()                                                      # built / present
    ns = {}
    exec(block.replace('"libkokoro_align_b200.so"', repr(_lib.SO_PATH)), ns)
    fn = ns["ctc_best_path"]
    index, arrays = load_golden()
    done = 0
    for case in index:
        if case["outcome"] != "ok" or case["max_move"] != 4 or case["T"] > 12000:
            continue
        lp, labels = case_inputs(case, arrays)
        path, labs, sc = fn(lp, labels, beam_size=case["beam_size"], max_move=case["max_move"])
        np.testing.assert_array_equal(path, arrays[f"{case['name']}.best_path"], err_msg=case["name"])
        done += 1
    assert done >= 20
    lp, labels = synth.make_lattice(40, 100, seed=9200)
    with pytest.raises(ValueError):
        fn(lp, labels, beam_size=20)
    with pytest.raises(IndexError):
        fn(lp, np.array([1, 39, 2]))


# ------------------------------------------------------------------ plan life cycle
def test_two_live_plans_with_different_shared_memory(kab):
    """Plans are reusable and several may be alive: a later, smaller plan (narrow vocabulary, narrow
    beam) must not shrink the dynamic-shared-memory limit the earlier plan's kernels need."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    cases = []
    for V, W, T, L in ((512, 1000, 3000, 420), (64, 200, 2500, 350), (512, 1000, 700, 100), (39, 1000, 600, 80)):
        lp, labels = synth.make_lattice(T, L, V, seed=9300 + V + T)
        cases.append((lp, labels, V, W, kab.AlignPlan([0, T], labels, [0, L], V, W)))
    for lp, labels, V, W, plan in cases + cases[::-1]:            # every plan runs after all were created
        path, labs, scores, final, status = plan.run_host(lp)
        rp, rl, rs, rf = ctc_oracle.ctc_best_path(lp, labels, W, 4, return_final_score=True)
        assert status[0] == 0
        np.testing.assert_array_equal(path, rp)
        assert final.tobytes() == np.float32(rf).tobytes()
    for c in cases:
        c[4].close()


def test_child_plan_failure_leaves_plan_usable(kab, monkeypatch):
    """kab_plan_run_host builds its streams, buffers and segment plans as ONE transaction: a child
    plan that fails returns the error and leaves nothing half-built, the next call starts over."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(1200, seed=9400)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=9401)
    ref = _oracle(lp, t_off, labels, l_off)
    monkeypatch.setenv("KAB_HOST_SEGMENTS", "4")
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        monkeypatch.setenv("KAB_TEST_FAIL_CHILD", "2")
        with pytest.raises(kab.KabError):
            plan.run_host(lp)
        monkeypatch.delenv("KAB_TEST_FAIL_CHILD")
        got = plan.run_host(lp)
        _assert_same(got, ref, t_off)
        got = plan.run_host(lp)                                   # and the built pipeline is reused
        _assert_same(got, ref, t_off)


@pytest.mark.parametrize("V", [39, 600, 4096])
def test_nonfinite_in_unused_column(kab, V):
    """One contract on every path: a non-finite value ANYWHERE in the lattice's rows is status 3,
    also in a column no label of the lattice uses and also on the compact (V > 512) path."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    for T, L in ((300, 40), (3000, 420)):
        lp, labels = synth.make_lattice(T, L, V, seed=9500 + V)
        labels = np.where(labels == 7, 8, labels)                 # column 7 is unused
        lp[T // 2, 7] = np.inf
        with pytest.raises(ValueError):
            kab.ctc_best_path(lp, labels)
        assert ctc_oracle.ctc_best_path_batch(lp, np.array([0, T]), labels.astype(np.int32), np.array([0, L]),
                                              n_threads=1)[4][0] == 3


def test_current_device_is_restored(kab):
    import torch
    from kokoro_align_b200 import synth
    lp, labels = synth.make_lattice(200, 20, seed=9600)
    before = torch.cuda.current_device()
    kab.ctc_best_path(lp, labels)
    assert torch.cuda.current_device() == before


# ------------------------------------------------------------------ the consumer on the device
def _host_records(path, labs, scores, t_off, indices_list):
    rec = []
    for b, ends in enumerate(indices_list):
        p, l, s = (x[int(t_off[b]):int(t_off[b + 1])] for x in (path, labs, scores))
        for i in range(len(ends)):
            a0 = int(ends[i - 1]) if i > 0 else 0
            b0 = int(ends[i])
            voiced = l[a0:b0] != 0
            rec.append((p[a0] // 2, p[b0] // 2 if b0 < len(p) else -1, np.sum(voiced),
                        np.sum(s[a0:b0][voiced]), np.sum(s[a0:b0]), 0))
    from kokoro_align_b200 import align
    return np.array(rec, dtype=align.SEGMENT_RECORD)


def test_segment_records_match_numpy(kab):
    """Per-segment boundaries, counts and BOTH np.sum's (numpy's pairwise float32 order) from the
    device, bit for bit, on a mixed batch: chapters cut into many segments of 1..2000 frames (block
    and split boundaries of the pairwise tree), single-segment lattices, an empty segment, and a
    segment end beyond the lattice; records only (no T-length array crosses PCIe), then with the
    arrays, then from device-resident tensors."""
    import torch
    from kokoro_align_b200 import synth
    rng = np.random.default_rng(9700)
    T = np.array([9000, 431, 86, 20000, 300, 1])
    L = np.round(0.14 * T).astype(np.int64)
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=9701, planted=True)
    indices = []
    for t in T:
        if t < 500:
            indices.append(np.array([t], np.int32))
            continue
        cuts = np.sort(rng.choice(np.arange(1, t), size=min(40, t // 50), replace=False))
        cuts = np.concatenate([cuts[:5], cuts[4:5], cuts[5:], [t]])     # one empty segment
        indices.append(cuts.astype(np.int32))
    indices[1] = np.array([200, 431 + 50], np.int32)                       # end past the lattice
    with kab.AlignPlan(t_off, labels, l_off, 39) as plan:
        path, labs, scores, final, status = plan.run_host(lp)
        want = _host_records(path, labs, scores, t_off, indices)
        rec, f2, st2, ex = plan.run_host_segments(lp, indices)
        assert ex == {} and rec.tobytes() == want.tobytes()
        assert f2.tobytes() == final.tobytes() and (st2 == 0).all()
        rec, _, _, ex = plan.run_host_segments(lp, indices, arrays=True, labels_u8=True)
        assert rec.tobytes() == want.tobytes()
        np.testing.assert_array_equal(ex["best_path"], path)
        np.testing.assert_array_equal(ex["labels_u8"], labs.astype(np.uint8))
        d = plan.run_torch(torch.from_numpy(lp).cuda())
        drec, dlab8 = plan.segment_stats_torch(d[0], d[1], d[2], d[4], indices, labels_u8=True)
        torch.cuda.synchronize()
        assert drec.cpu().numpy().reshape(-1).view(kab.SEGMENT_RECORD).tobytes() == want.tobytes()
        np.testing.assert_array_equal(dlab8.cpu().numpy(), labs.astype(np.uint8))
        # decreasing segment ends are refused
        with pytest.raises(kab.KabError):
            plan.run_host_segments(lp, [np.array([10, 5])] + indices[1:])


def test_book_pipeline_run_host(kab, monkeypatch):
    """kab_plan_run_host on a book: the longest chapters' rows are copied first and run in their own
    sub-plan while the other chapters' rows arrive (forced here on a small book; a 400 MB book takes it
    by itself).  Same bits as the oracle, from log-probs, and the same records / arrays through the
    segment entry point; a chapter whose band dies keeps its status."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    T = np.array([2500, 9000, 700, 5000, 350, 3000, 4000, 6000, 1200])
    L = np.maximum(1, np.round(0.14 * T)).astype(np.int64)
    L[5] = 2900                                                  # S > 3T / 2: the 200-wide band loses the end state
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=9700)
    lp = (np.round(lp * 4) / 4).astype(np.float32)
    rp, rl, rs, rf, rst = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, 200, 4, n_threads=4)
    monkeypatch.setenv("KAB_HOST_BOOK", "1")
    with kab.AlignPlan(t_off, labels, l_off, 39, 200) as plan:
        for _ in range(2):                                       # (the second run reuses the pipeline)
            path, labs, scores, final, status = plan.run_host(lp)
            np.testing.assert_array_equal(status, rst)
            for b in np.flatnonzero(rst == 0):
                a, e = int(t_off[b]), int(t_off[b + 1])
                np.testing.assert_array_equal(path[a:e], rp[a:e])
                np.testing.assert_array_equal(labs[a:e], rl[a:e])
                assert scores[a:e].tobytes() == rs[a:e].tobytes() and final[b].tobytes() == rf[b].tobytes()
        indices = [np.array([t // 3, 2 * t // 3, t]) for t in T]
        rec, f2, st2, ex = plan.run_host_segments(lp, indices, arrays=True)
        np.testing.assert_array_equal(st2, rst)
        ok = np.repeat(rst == 0, 3)
        want = _host_records(np.where(np.repeat(rst, T) == 0, rp, 0), rl, rs, t_off, indices)
        assert rec[ok].tobytes() == want[ok].tobytes()
    monkeypatch.setenv("KAB_HOST_BOOK", "0")
    with kab.AlignPlan(t_off, labels, l_off, 39, 200) as plan:
        p2, l2, s2, f2, st2 = plan.run_host(lp)
    np.testing.assert_array_equal(st2, status)
    good = np.repeat(status == 0, T)
    np.testing.assert_array_equal(p2[good], path[good])
    assert s2[good].tobytes() == scores[good].tobytes()


def test_segment_records_long_segment_and_failed_lattice(kab):
    """A single 120 000-frame segment (more leaves than a warp keeps: the one-lane path) and a
    batch with a dead-band lattice (its records carry the status)."""
    from kokoro_align_b200 import synth
    T = np.array([120000, 40])
    L = np.array([16800, 100])
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=9800)
    with kab.AlignPlan(t_off, labels, l_off, 39, beam_size=20) as plan:   # beam 20: the short lattice's band dies
        path, labs, scores, final, status = plan.run_host(lp)
        assert status.tolist() == [0, 1]
        rec, _, st, _ = plan.run_host_segments(lp, [np.array([120000]), np.array([20, 40])])
    want = _host_records(path, labs, scores, t_off, [np.array([120000])])
    assert rec[:1].tobytes() == want.tobytes()
    assert rec["status"].tolist() == [0, 1, 1]


@pytest.mark.parametrize("remove_wordsep", [True, False])
def test_align_text_from_device_records(kab, tmp_path, remove_wordsep):
    """align()'s text (align.py:127-169) from numbers computed on the DEVICE: identical to what the
    reference wrote for the golden case."""
    with np.load(os.path.join(CASE, "case.log_probs.npz")) as f:
        lp, labels = f["log_probs"], f["labels"]
    with np.load(os.path.join(CASE, "case.mfcc.npz")) as f:
        ends = f["indices"]
    T = len(lp)
    with kab.AlignPlan([0, T], labels, [0, len(labels)], 39) as plan:
        rec, _, status, ex = plan.run_host_segments(lp, [ends], labels_u8=True)
    assert status[0] == 0
    out = tmp_path / "dev.align.txt"
    kab.align_from_records(rec, ex["labels_u8"], ends, os.path.join(CASE, "case.voca.txt"), str(out), remove_wordsep)
    assert out.read_text() == open(os.path.join(CASE, f"case.align.wordsep{int(not remove_wordsep)}.txt")).read()


def test_best_path_and_align_files(kab, tmp_path):
    """The fused file-level entry: logits.npz + mfcc.npz + voca.txt -> align.txt (+ best_path.npz),
    equal to best_path() followed by align()."""
    with np.load(os.path.join(CASE, "case.log_probs.npz")) as f:
        lp = f["log_probs"]
    logits = tmp_path / "c.logits.npz"
    np.savez(logits, data=lp * np.float32(1.5) + np.float32(0.25), indices=np.array([len(lp)], np.int32))
    voca, mfcc = os.path.join(CASE, "case.voca.txt"), os.path.join(CASE, "case.mfcc.npz")
    kab.best_path(str(logits), voca, str(tmp_path / "a.best_path.npz"))
    kab.align(str(tmp_path / "a.best_path.npz"), mfcc, voca, str(tmp_path / "a.align.txt"), True)
    kab.best_path_and_align(str(logits), mfcc, voca, str(tmp_path / "b.align.txt"), True,
                            best_path_file=str(tmp_path / "b.best_path.npz"))
    assert (tmp_path / "a.align.txt").read_text() == (tmp_path / "b.align.txt").read_text()
    with np.load(tmp_path / "a.best_path.npz") as a, np.load(tmp_path / "b.best_path.npz") as b:
        for k in ("best_path", "best_labels", "best_scores"):
            assert a[k].dtype == b[k].dtype and a[k].tobytes() == b[k].tobytes()


# ------------------------------------------------------------------ the producer hand-off
def test_encoder_handoff_stays_on_device(kab):
    """predict() -> best_path() without the host round trip (train.py:215-229, align.py:113-117):
    padded time-major encoder batches are packed + normalised on the device (ChapterLogits), the
    chapter is aligned from HBM, and the result equals the host path on the same log-probs."""
    import torch
    from kokoro_align_b200 import synth
    rng = np.random.default_rng(9900)
    lens_all = [int(x) for x in rng.integers(86, 700, 23)]
    V, T = 39, sum(lens_all)
    labels = rng.integers(1, V, int(0.14 * T)).astype(np.int8)
    chapter = kab.ChapterLogits(T, V)
    host_rows = []
    for b0 in range(0, len(lens_all), 8):                       # batches of 8 segments, as the DataLoader gives
        lens = lens_all[b0:b0 + 8]
        t_max = max(lens)
        x = (rng.standard_normal((t_max, len(lens), V)) * 3).astype(np.float32)
        chapter.append(torch.from_numpy(x).cuda(), torch.tensor(lens))
        for j, n in enumerate(lens):
            host_rows.append(x[:n, j, :])                       # what predict() writes to the npz
    torch.cuda.synchronize()
    assert chapter.rows == T and chapter.indices == list(np.cumsum(lens_all))
    logits = np.concatenate(host_rows)
    lp_dev = chapter.log_probs.cpu().numpy()
    np.testing.assert_allclose(lp_dev, synth.log_softmax_ref(logits), atol=4e-6, rtol=0)
    # the same normalisation as the flat device kernel, bit for bit
    flat = kab.log_softmax_torch(torch.from_numpy(logits).cuda()).cpu().numpy()
    assert flat.tobytes() == lp_dev.tobytes()
    out = kab.best_path_from_logits_tensor(chapter.log_probs, labels, indices=chapter.indices, normalised=True)
    torch.cuda.synchronize()
    path, labs, scores = kab.ctc_best_path(lp_dev, labels)      # host path on the device's log-probs
    np.testing.assert_array_equal(out["best_path"].cpu().numpy(), path)
    np.testing.assert_array_equal(out["best_labels"].cpu().numpy(), labs)
    assert out["best_scores"].cpu().numpy().tobytes() == scores.tobytes()
    want = _host_records(path, labs, scores, np.array([0, T]), [np.array(chapter.indices)])
    assert out["records"].cpu().numpy().reshape(-1).view(kab.SEGMENT_RECORD).tobytes() == want.tobytes()
    # raw logits in, normalised in place
    d_logits = torch.from_numpy(logits).cuda()
    out2 = kab.best_path_from_logits_tensor(d_logits, labels)
    np.testing.assert_array_equal(out2["best_path"].cpu().numpy(), path)
    assert d_logits.cpu().numpy().tobytes() == lp_dev.tobytes()


# ------------------------------------------------------------------ non-default keywords in the staged kernels
@pytest.mark.parametrize("max_move", [1, 2, 3])
@pytest.mark.parametrize("band_env", [{}, {"KAB_BAND_CLUSTER": "0"}])
def test_max_move_below_four_in_staged_kernels(kab, monkeypatch, max_move, band_env):
    """align.py:43's `max_move` keyword at 1, 2, 3: same warp / band kernels (their MM instantiations
    turn the candidates of the excluded moves into -inf), not the generic kernel; bit-exact against
    the C oracle, tie stress included; dead bands (max_move 1 cannot leave state 0) keep status 1."""
    from kokoro_align_b200 import synth
    from oracle import ctc_oracle
    for k, v in band_env.items():
        monkeypatch.setenv(k, v)
    T = np.array([300, 861, 120, 5000, 2500, 9, 700])
    L = np.array([20, 60, 3, 350, 400, 1, 0])       # sparse transcripts: the path can end at S-1 without long jumps
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=9950 + max_move, planted=True)
    lp[: int(t_off[2])] = (np.round(lp[: int(t_off[2])] * 2) / 2).astype(np.float32)   # ties in the first two lattices
    ref = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, 1000, max_move, n_threads=4)
    with kab.AlignPlan(t_off, labels, l_off, 39, 1000, max_move) as plan:
        got = plan.run_host(lp)
        assert plan.info.n_class[2] == 0, "max_move < 4 must not fall back to the generic kernel"
        assert plan.info.n_class[0] >= 4 and plan.info.n_class[1] >= 2
    path, labs, scores, final, status = got
    rp, rl, rs, rf, rst = ref
    np.testing.assert_array_equal(status, rst)
    for b in range(len(T)):
        if rst[b] != 0:
            continue
        a, e = int(t_off[b]), int(t_off[b + 1])
        np.testing.assert_array_equal(path[a:e], rp[a:e], err_msg=f"lattice {b}")
        assert scores[a:e].tobytes() == rs[a:e].tobytes() and final[b].tobytes() == rf[b].tobytes()
    assert (rst == 0).sum() >= (5 if max_move >= 2 else 1)
