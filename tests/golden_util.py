"""Loader for tests/golden/ (vectors produced by the reference, see golden/make_golden.py)."""
import json
import os

import numpy as np

from kokoro_align_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EXC = {"ValueError": ValueError, "IndexError": IndexError}


def case_inputs(case, arrays):
    """Rebuild (log_probs, labels) of a golden case: stored for "gauss", recipe for "exact"."""
    name = case["name"]
    if case["kind"] == "gauss":
        return arrays[f"{name}.log_probs"], arrays[f"{name}.labels"]
    kw = dict(case["kw"])
    zero_every = kw.pop("zero_every", 0)
    neg_every = kw.pop("neg_every", 0)
    bad_at = kw.pop("bad_at", None)
    lp, labels = synth.make_lattice_exact(case["T"], case["L"], case["V"], case["seed"], **kw)
    labels = labels.astype(np.int32)
    if zero_every:
        labels[::zero_every] = 0
    if neg_every:
        labels[::neg_every] = -1 - (np.arange(len(labels[::neg_every])) % case["V"])
    if bad_at is not None:
        labels[bad_at] = case["V"]
    return lp, labels


def load_golden():
    with open(os.path.join(HERE, "golden", "golden_index.json")) as f:
        index = json.load(f)["cases"]
    arrays = dict(np.load(os.path.join(HERE, "golden", "golden.npz")))
    return index, arrays


def golden_names():
    with open(os.path.join(HERE, "golden", "golden_index.json")) as f:
        return [c["name"] for c in json.load(f)["cases"]]


def check_case(case, arrays, fn):
    """Run fn(log_probs, labels, beam_size=, max_move=, return_final_score=True) on a golden
    case and compare with what the reference produced: bit-exact path/labels/scores, the
    reference's exception type, final score bit-equal to the sequential fp32 sum."""
    import pytest
    lp, labels = case_inputs(case, arrays)
    name = case["name"]
    kw = dict(beam_size=case["beam_size"], max_move=case["max_move"], return_final_score=True)
    if case["outcome"] != "ok":
        with pytest.raises(EXC[case["outcome"]]):
            fn(lp, labels, **kw)
        return
    path, labs, scores, final = fn(lp, labels, **kw)
    ref_path = arrays[f"{name}.best_path"]
    assert path.dtype == np.int32 and labs.dtype == np.int32 and scores.dtype == np.float32
    np.testing.assert_array_equal(path, ref_path)
    ext = np.zeros(2 * len(labels) + 1, np.int32)
    ext[1::2] = labels
    np.testing.assert_array_equal(labs, ext[ref_path])
    ref_scores = lp[np.arange(len(ref_path)), ext[ref_path]]
    assert scores.tobytes() == ref_scores.tobytes()
    if f"{name}.best_labels" in arrays:
        np.testing.assert_array_equal(labs, arrays[f"{name}.best_labels"])
        assert scores.tobytes() == arrays[f"{name}.best_scores"].tobytes()
    ref_final = np.frombuffer(bytes.fromhex(case["final_score_hex"]), np.float32)[0]
    # tolerance from north_star: 1e-4 relative; in practice it is bit-equal
    assert abs(float(final) - float(ref_final)) <= 1e-4 * max(1.0, abs(float(ref_final)))
    assert np.float32(final).tobytes() == ref_final.tobytes()
