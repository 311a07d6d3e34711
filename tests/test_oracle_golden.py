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


@pytest.mark.parametrize("seed", range(24))
def test_oracles_agree_on_random_shapes(seed):
    """Beyond the golden vectors: the two independently written restatements (C per cell, numpy per
    frame) agree bit for bit -- including the exception they raise -- on random shapes, beams,
    max_move, tie-heavy (quantised) and planted log-probs."""
    from kokoro_align_b200 import synth
    rng = np.random.default_rng(9000 + seed)
    T = int(rng.integers(1, 400))
    L = int(rng.integers(0, max(1, min(3 * T // 2, 300))))
    V = int(rng.choice([3, 5, 39, 64]))
    beam = int(rng.choice([4, 7, 16, 33, 100, 1000]))
    max_move = int(rng.choice([1, 2, 3, 4, 4, 4, 5, 7]))
    lp, labels = synth.make_lattice(T, L, V, seed=9100 + seed, planted=bool(seed & 1))
    if seed % 3 == 0:
        lp = (np.round(lp * 2) / 2).astype(np.float32)
    out = []
    for mod in (ctc_oracle, ctc_oracle_np):
        try:
            out.append(mod.ctc_best_path(lp, labels, beam, max_move, return_final_score=True))
        except (ValueError, IndexError) as e:
            out.append(type(e))
    if isinstance(out[0], type) or isinstance(out[1], type):
        assert out[0] is out[1]
        return
    (p0, l0, s0, f0), (p1, l1, s1, f1) = out
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_array_equal(l0, l1)
    assert s0.tobytes() == s1.tobytes() and np.float32(f0).tobytes() == np.float32(f1).tobytes()
