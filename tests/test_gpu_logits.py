"""the normalisation align.py:116-117 on the device (kab_softmax.cuh) and the
raw-logits entry kab_plan_run_host_logits.

Parity statement (two steps, both asserted here):
  (1) the device log-probs agree with numpy's align.py:116-117 within LP_ATOL = 4e-6 absolute
      (a few fp32 ulp of values in [-20, 0]); bit equality is not attainable -- numpy's SIMD
      float32 exp is itself not correctly rounded and depends on the host CPU;
  (2) the alignment of THOSE log-probs is bit-exact against the C oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LP_ATOL = 4e-6


def _np_log_softmax(x):  # align.py:116-117, verbatim semantics
    x = x - np.mean(x, axis=-1, keepdims=True)
    return x - np.log(np.sum(np.exp(x), axis=-1, keepdims=True))


@pytest.fixture(scope="module")
def kab():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from kokoro_align_b200 import align
    return align


@pytest.mark.parametrize("V", [39, 1, 5, 8, 33, 64, 127, 128, 129, 600, 4096])
@pytest.mark.parametrize("rows", [1, 255, 1000, 4097])
def test_log_softmax_device(kab, V, rows):
    import torch
    rng = np.random.default_rng(7000 + V + rows)
    x = (rng.standard_normal((rows, V)) * 3).astype(np.float32)
    ref = _np_log_softmax(x)
    d = torch.from_numpy(x).cuda()
    out = kab.log_softmax_torch(d)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=0, atol=LP_ATOL)
    # in place, and from a pointer that is only 4-byte aligned (row 1 of an odd-V array)
    kab.log_softmax_torch(d, out=d)
    torch.cuda.synchronize()
    assert d.cpu().numpy().tobytes() == got.tobytes()
    if rows > 1:
        d2 = torch.from_numpy(x).cuda()[1:]
        got2 = kab.log_softmax_torch(d2).cpu().numpy()
        assert got2.tobytes() == got[1:].tobytes()


def test_log_softmax_large_logits(kab):
    """Peaked rows (the planted-path recipe adds +6) and a wide dynamic range."""
    import torch
    rng = np.random.default_rng(7100)
    x = (rng.standard_normal((5000, 39)) * 8).astype(np.float32)
    x[np.arange(5000), rng.integers(0, 39, 5000)] += 20
    ref = _np_log_softmax(x)
    got = kab.log_softmax_torch(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=LP_ATOL)


def _check_logits_batch(kab, logits, t_off, labels, l_off, V=39):
    from oracle import ctc_oracle
    lp_dev = np.empty_like(logits)
    with kab.AlignPlan(t_off, labels, l_off, V) as plan:
        path, labs, scores, final, status = plan.run_host(logits, logits=True, log_probs_out=lp_dev)
    np.testing.assert_allclose(lp_dev, _np_log_softmax(logits), rtol=0, atol=LP_ATOL)       # (1)
    rp, rl, rs, rf, rst = ctc_oracle.ctc_best_path_batch(lp_dev, t_off, labels, l_off, n_threads=8)
    np.testing.assert_array_equal(status, rst)                                               # (2)
    ok = np.repeat(rst == 0, np.diff(t_off))
    np.testing.assert_array_equal(path[ok], rp[ok])
    np.testing.assert_array_equal(labs[ok], rl[ok])
    assert scores[ok].tobytes() == rs[ok].tobytes()
    assert final[rst == 0].tobytes() == rf[rst == 0].tobytes()
    return path


def test_run_host_logits_small(kab):
    from kokoro_align_b200 import synth
    T = np.array([86, 300, 431, 861, 6000])
    L = np.array([12, 42, 60, 121, 1700])
    rng = np.random.default_rng(7200)
    logits = (rng.standard_normal((int(T.sum()), 39)) * 3).astype(np.float32)
    _, t_off, labels, l_off = synth.make_batch(T, L, seed=7201)
    _check_logits_batch(kab, logits, t_off, labels, l_off)


def test_run_host_logits_pipelined(kab):
    """>= 96 MB of logits: the segmented H2D / normalise / align / D2H pipeline."""
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(2200, seed=7300)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=7301)
    rng = np.random.default_rng(7302)
    logits = lp + rng.standard_normal((lp.shape[0], 1)).astype(np.float32) * 2   # un-normalised rows
    assert logits.nbytes >= 96 << 20
    _check_logits_batch(kab, logits, t_off, labels, l_off)


def test_best_path_device_log_softmax(kab, tmp_path):
    """best_path(..., device_log_softmax=True) and best_path_files(..., device_log_softmax=True):
    same npz keys / dtypes / shapes as the host-normalised run.  The few-ulp difference of the
    log-probs may move a boundary where two paths tie within rounding, so the frames are compared
    statistically (>= 99 % identical) and best_scores within the log-softmax tolerance where the
    labels agree; exactness is asserted in _check_logits_batch against the oracle."""
    rng = np.random.default_rng(7400)
    V = 39
    files = []
    for n, T in enumerate((2600, 700)):
        lf, vf = tmp_path / f"c{n}.logits.npz", tmp_path / f"c{n}.voca.txt"
        with open(vf, "w") as f:
            for k in range(T // 90):
                f.write(f"text {k}|k o k o r o , w a t a sh i\n")
        np.savez(lf, data=(rng.standard_normal((T, V)) * 3).astype(np.float32), indices=np.array([T], np.int32))
        files.append((str(lf), str(vf)))
    kab.best_path(files[0][0], files[0][1], str(tmp_path / "host0.npz"))
    kab.best_path(files[1][0], files[1][1], str(tmp_path / "host1.npz"))
    kab.best_path(files[0][0], files[0][1], str(tmp_path / "dev0.npz"), device_log_softmax=True)
    t = {}
    written = kab.best_path_files([f[0] for f in files], [f[1] for f in files],
                                  [str(tmp_path / "book0.npz"), str(tmp_path / "book1.npz")],
                                  device_log_softmax=True, verbose=False, timings=t)
    assert len(written) == 2 and t["chapters"] == 2 and t["frames"] == 3300
    with np.load(tmp_path / "dev0.npz") as a, np.load(tmp_path / "book0.npz") as b:
        for k in a.files:
            assert a[k].tobytes() == b[k].tobytes()          # per-file and per-book entry agree
    for n in (0, 1):
        with np.load(tmp_path / f"host{n}.npz") as a, np.load(tmp_path / f"book{n}.npz") as b:
            assert sorted(a.files) == sorted(b.files) == ["best_labels", "best_path", "best_scores"]
            for k in a.files:
                assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape
            same = a["best_path"] == b["best_path"]
            assert same.mean() >= 0.99
            np.testing.assert_allclose(a["best_scores"][same], b["best_scores"][same], rtol=0, atol=LP_ATOL)
    # the host-normalised per-book entry stays bit-identical to the per-file entry
    kab.best_path_files([files[1][0]], [files[1][1]], [str(tmp_path / "bookh1.npz")], verbose=False)
    with np.load(tmp_path / "host1.npz") as a, np.load(tmp_path / "bookh1.npz") as b:
        for k in a.files:
            assert a[k].tobytes() == b[k].tobytes()
