"""The two CPU oracles (C per-cell, numpy per-frame) against the reference's golden vectors."""
import numpy as np
import pytest

from oracle import ctc_oracle, ctc_oracle_np
from tests.golden_util import check_case, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_reference(golden, name):
    index, arrays = golden
    case = next(c for c in index if c["name"] == name)
    check_case(case, arrays, ctc_oracle.ctc_best_path)


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith(("T20500", "T30011"))])
def test_numpy_oracle_matches_reference(golden, name):
    index, arrays = golden
    case = next(c for c in index if c["name"] == name)
    check_case(case, arrays, ctc_oracle_np.ctc_best_path)


def test_cells_eval_agree():
    for T, L, W in ((81135, 11359, 1000), (861, 121, 1000), (100, 300, 20), (1, 0, 1000)):
        assert ctc_oracle.cells_eval(T, L, W) == ctc_oracle_np.cells_eval(T, L, W)
    assert ctc_oracle.cells_eval(861, 121, 1000) == 861 * 243


def test_nonfinite_rejected():
    lp = np.zeros((5, 4), np.float32)
    lp[2, 1] = -np.inf
    for mod in (ctc_oracle, ctc_oracle_np):
        with pytest.raises(ValueError):
            mod.ctc_best_path(lp, np.array([1, 2], np.int32))


def test_batch_matches_single():
    from kokoro_align_b200 import synth
    T, L = synth.segment_lengths(12, seed=7, t_min=20, t_max=120)
    lp, t_off, labels, l_off = synth.make_batch(T, L, seed=50)
    p, l, s, fs, st = ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, n_threads=3)
    assert (st == 0).all()
    for b in range(len(T)):
        rp, rl, rs, rf = ctc_oracle.ctc_best_path(lp[t_off[b]:t_off[b + 1]],
                                                  labels[l_off[b]:l_off[b + 1]],
                                                  return_final_score=True)
        np.testing.assert_array_equal(p[t_off[b]:t_off[b + 1]], rp)
        np.testing.assert_array_equal(l[t_off[b]:t_off[b + 1]], rl)
        assert s[t_off[b]:t_off[b + 1]].tobytes() == rs.tobytes()
        assert fs[b].tobytes() == np.float32(rf).tobytes()
