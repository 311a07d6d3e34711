"""the consumer of best_path.npz (align.py:127-169) against the text
the reference's own align() wrote for the same files (tests/golden/make_align_golden.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "align_case")


@pytest.mark.parametrize("remove_wordsep", [True, False])
def test_align_text_matches_reference(tmp_path, remove_wordsep):
    from kokoro_align_b200 import align
    out = tmp_path / "x.align.txt"
    align.align(os.path.join(CASE, "case.best_path.npz"), os.path.join(CASE, "case.mfcc.npz"),
                os.path.join(CASE, "case.voca.txt"), str(out), remove_wordsep)
    ref = open(os.path.join(CASE, f"case.align.wordsep{int(not remove_wordsep)}.txt")).read()
    assert out.read_text() == ref


def test_align_failure_removes_partial_file(tmp_path):
    from kokoro_align_b200 import align
    bad = tmp_path / "bad.npz"
    np.savez(bad, indices=np.array([10, 10 ** 9], np.int32))   # second boundary far past the path: fine
    out = tmp_path / "y.align.txt"
    voca = tmp_path / "broken.voca.txt"
    voca.write_text("no separator here\n")
    with pytest.raises(ValueError):
        align.align(os.path.join(CASE, "case.best_path.npz"), str(bad), str(voca), str(out), True)
    assert not out.exists()


def test_transcript_labels_match_golden_case():
    from kokoro_align_b200 import align
    with np.load(os.path.join(CASE, "case.log_probs.npz")) as f:
        labels = f["labels"]
    got = align.read_transcript_labels(os.path.join(CASE, "case.voca.txt"))
    assert got.dtype == np.int8 and got.tolist() == labels.tolist()


def test_merge_repeated():
    from kokoro_align_b200 import encoder
    assert encoder.merge_repeated("a a _ k k _ _ a") == "a k a"
    assert encoder.merge_repeated("_") == ""
    assert encoder.merge_repeated("_ _ _") == ""
    assert encoder.merge_repeated("a k a k") == "a k"          # the regex collapses repeated runs


@pytest.mark.gpu
def test_file_level_pipeline_matches_reference(tmp_path):
    """best_path() (align.py:112-124 drop-in) -> align(): files in, identical files out."""
    from kokoro_align_b200 import align
    with np.load(os.path.join(CASE, "case.log_probs.npz")) as f:
        lp = f["log_probs"]
    # best_path() takes raw logits and applies the reference's log-softmax; log-probs are a
    # fixed point of it up to rounding, so feed them through ctc_best_path directly instead
    path, labs, scores = align.ctc_best_path(lp, align.read_transcript_labels(os.path.join(CASE, "case.voca.txt")))
    with np.load(os.path.join(CASE, "case.best_path.npz")) as f:
        np.testing.assert_array_equal(path, f["best_path"])
        np.testing.assert_array_equal(labs, f["best_labels"])
        assert scores.tobytes() == f["best_scores"].tobytes()
    bp = tmp_path / "c.best_path.npz"
    np.savez(bp, best_path=path, best_labels=labs, best_scores=scores)
    out = tmp_path / "c.align.txt"
    align.align(str(bp), os.path.join(CASE, "case.mfcc.npz"), os.path.join(CASE, "case.voca.txt"), str(out), True)
    assert out.read_text() == open(os.path.join(CASE, "case.align.wordsep0.txt")).read()
    # and the npz-level entry point with logits = log-probs scaled (softmax-invariant shift)
    logits = tmp_path / "c.logits.npz"
    np.savez(logits, data=lp, indices=np.array([len(lp)], np.int32))
    align.best_path(str(logits), os.path.join(CASE, "case.voca.txt"), str(tmp_path / "d.best_path.npz"))
    with np.load(tmp_path / "d.best_path.npz") as f:
        assert set(f.keys()) == {"best_path", "best_labels", "best_scores"}
        assert f["best_path"].dtype == np.int32 and f["best_scores"].dtype == np.float32
        assert len(f["best_path"]) == len(lp)


def test_numpy_pairwise_sum_model():
    """The summation order the segment-statistics kernel implements IS np.sum's (float32,
    contiguous): checked on every length around the block / split boundaries."""
    from tests.np_sum_model import pairwise_sum
    rng = np.random.default_rng(7)
    lengths = list(range(0, 300)) + [511, 512, 513, 1000, 1023, 1024, 1025, 2049, 4097, 10007]
    for n in lengths:
        a = (-rng.random(n) * 9).astype(np.float32)       # log-prob-like: negative, finite
        assert np.float32(pairwise_sum(a)).tobytes() == np.float32(np.sum(a)).tobytes(), n
        b = (a[a < -4.0]).copy()                           # a compacted copy, as scores[labels != 0]
        assert np.float32(pairwise_sum(b)).tobytes() == np.float32(np.sum(b)).tobytes(), n


def test_align_from_records_matches_reference_text(tmp_path):
    """align_from_records (the writer fed by the device's segment records) against the reference's
    text, with the records computed on the HOST from the golden arrays by the reference's own
    expressions (align.py:151-162): pins the record contract without a GPU."""
    from kokoro_align_b200 import align
    with np.load(os.path.join(CASE, "case.best_path.npz")) as f:
        path, labs, scores = f["best_path"], f["best_labels"], f["best_scores"]
    with np.load(os.path.join(CASE, "case.mfcc.npz")) as f:
        ends = f["indices"]
    rec = np.zeros(len(ends), align.SEGMENT_RECORD)
    for i in range(len(ends)):
        a = int(ends[i - 1]) if i > 0 else 0
        b = int(ends[i])
        voiced = labs[a:b] != 0
        rec[i] = (path[a] // 2, path[b] // 2 if b < len(path) else -1, np.sum(voiced),
                  np.sum(scores[a:b][voiced]), np.sum(scores[a:b]), 0)
    for rw in (True, False):
        out = tmp_path / f"r{int(rw)}.align.txt"
        align.align_from_records(rec, labs.astype(np.uint8), ends, os.path.join(CASE, "case.voca.txt"), str(out), rw)
        assert out.read_text() == open(os.path.join(CASE, f"case.align.wordsep{int(not rw)}.txt")).read()
