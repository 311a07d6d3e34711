"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol
include/kokoro_align_b200.h declares; the host mirror keeps the reference's signatures; the
product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kokoro_align_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kab_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kokoro_align_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert L.kab_version() == 200
    assert L.kab_error_string(-2) == b"bad argument"


def test_band_kernel_codes_match_the_header():
    """kab_plan_info.band_kernel: the Python mirror names exactly the KAB_BAND_KERNEL_* codes of the
    header, and the struct mirrors have the header's sizes."""
    from kokoro_align_b200 import _lib
    text = open(os.path.join(ROOT, "include", "kokoro_align_b200.h")).read()
    codes = {name: int(v) for name, v in re.findall(r"#define\s+KAB_BAND_KERNEL_(\w+)\s+(\d+)", text)}
    assert codes == {"CTA": 1, "CLUSTER": 2, "CLUSTER2": 3, "SPEC": 4, "HYBRID": 5}
    assert sorted(k for k in _lib.BAND_KERNELS if k) == sorted(codes.values())
    assert "kab_bandr_kernel" in _lib.BAND_KERNELS[codes["SPEC"]] and "kab_band_kernel" in _lib.BAND_KERNELS[codes["HYBRID"]]
    assert ctypes.sizeof(_lib.SegmentRecord) == 24


def test_reference_signatures():
    from kokoro_align_b200 import align
    sig = inspect.signature(align.ctc_best_path)
    assert list(sig.parameters)[:4] == ["log_probs", "labels", "beam_size", "max_move"]
    assert sig.parameters["beam_size"].default == 1000 and sig.parameters["max_move"].default == 4
    assert list(inspect.signature(align.best_path).parameters)[:3] == ["input_file", "voca_file", "output_file"]
    assert inspect.signature(align.best_path).parameters["device_log_softmax"].default is False


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from kokoro_align_b200 import align
    with pytest.raises(align.KabError):
        align.ctc_best_path(np.zeros((4, 5), np.float32), np.array([1], np.int8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kokoro-align_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_argument_errors_match_reference():
    from kokoro_align_b200 import align
    with pytest.raises(IndexError):
        align.ctc_best_path(np.zeros((0, 5), np.float32), np.array([1], np.int8))
    with pytest.raises(ValueError):
        align.ctc_best_path(np.zeros((3, 5), np.float32), np.array([1], np.int8), max_move=0)


def test_encoder_contract():
    from kokoro_align_b200 import encoder
    assert encoder.VOCAB_SIZE == 39 and encoder.VOCAB[0] == "_"
    ids = encoder.encode_text("a k a q . N zz")
    assert ids.dtype == np.int8 and ids.tolist() == [2, 18, 2, 1]


def test_transcript_labels(tmp_path):
    from kokoro_align_b200 import align
    p = tmp_path / "x.voca.txt"
    p.write_text("こん|k o N\nにちは|n i ch i w a .\n")
    assert align.read_transcript_labels(str(p)).tolist() == [18, 24, 1, 22, 15, 6, 15, 36, 2]


def test_npz_wire_format(tmp_path):
    """{data, indices} archives as np.savez writes them (preprocess.py:12-35,
    train.py:228) are read straight into a caller-provided buffer; compressed archives and other
    dtypes fall back to np.load."""
    from kokoro_align_b200 import align
    rng = np.random.default_rng(3)
    x = rng.standard_normal((777, 39)).astype(np.float32)
    idx = np.array([300, 777], np.int32)
    np.savez(tmp_path / "stored.npz", indices=idx, data=x)
    np.savez_compressed(tmp_path / "deflated.npz", indices=idx, data=x)
    np.savez(tmp_path / "f64.npz", indices=idx, data=x.astype(np.float64))
    shape, dtype, off = align.npz_member_info(str(tmp_path / "stored.npz"))
    assert shape == (777, 39) and dtype == np.float32 and off is not None
    with open(tmp_path / "stored.npz", "rb") as f:
        f.seek(off)
        assert f.read(x.nbytes) == x.tobytes()
    assert align.npz_member_info(str(tmp_path / "deflated.npz"))[2] is None
    for name in ("stored", "deflated", "f64"):
        dst = np.full((777, 39), np.nan, np.float32)
        align.npz_read_into(str(tmp_path / f"{name}.npz"), dst)
        assert dst.tobytes() == x.tobytes()
    assert align.npz_member_info(str(tmp_path / "stored.npz"), "indices")[0] == (2,)
    with pytest.raises(ValueError):
        align.npz_read_into(str(tmp_path / "stored.npz"), np.empty((5, 39), np.float32))


def test_transcript_scanner_matches_text_path(tmp_path):
    """kab_encode_transcript (byte scan in the C library) against the reference's control flow
    (line.rstrip('\r\n').split('|')[1], ' '.join, str.split, vocabulary lookup) on fuzzed files:
    same labels, same exception type; odd files fall back to the text path."""
    from kokoro_align_b200 import align, encoder
    rng = np.random.default_rng(0)
    toks = list(encoder.VOCAB) + [',', '.', '!', '?', 'xx', 'abc', 'q', 'N:', 'a::', '~', '|']
    n_plain = 0
    for trial in range(300):
        lines = []
        for _ in range(rng.integers(0, 12)):
            text = ''.join(rng.choice(list('あいう ab')) for _ in range(rng.integers(0, 5)))
            voca = (' ' * rng.integers(1, 3)).join(rng.choice(toks[:-1]) for _ in range(rng.integers(0, 7)))
            if rng.random() < 0.1:
                voca = ' ' + voca + '  '
            line = text + '|' + voca
            r = rng.random()
            if r < 0.10:
                line += '|extra|more a i'
            elif r < 0.13:
                line = text                              # no '|': IndexError in both paths
            elif r < 0.16:
                line = text + '|' + voca.replace(' ', '\t', 1)
            elif r < 0.18:
                line = text + '|' + voca + '　a'
            lines.append(line)
        nlc = rng.choice(['\n', '\r\n']) if rng.random() < 0.9 else '\r'
        content = (nlc.join(lines) + (nlc if rng.random() < 0.7 else '')).encode()
        p = tmp_path / f"t{trial}.txt"
        p.write_bytes(content)
        res = []
        for fn in (align._read_transcript_labels_text, align.read_transcript_labels):
            try:
                res.append(fn(str(p)))
            except Exception as e:  # noqa: BLE001
                res.append(type(e))
        if isinstance(res[0], type):
            assert res[1] is res[0]
        else:
            assert res[1].dtype == np.int8 and res[1].tolist() == res[0].tolist()
        out, n = np.empty(len(content) // 2 + 1, np.int8), ctypes.c_int64(0)
        from kokoro_align_b200 import _lib
        n_plain += _lib.lib().kab_encode_transcript(content, len(content), align._ptr(align._TOKEN_IDS),
                                                    align._ptr(out), ctypes.byref(n)) == 0
    assert 60 < n_plain < 300


def test_merge_repeated_matches_the_regular_expression():
    """encoder.merge_repeated evaluates encoder.py:28's regex in C (kab_merge_repeated): identical
    to `re` on random token runs, double spaces, sub-token matches ('k ky'), and it falls back to
    `re` for text the byte scan does not take."""
    import re
    from kokoro_align_b200 import encoder

    def ref(text):
        r = re.sub(r'(.+)( \1)+', r'\1', text).replace(' _', '').replace('_ ', '')
        return '' if r == '_' else r
    rng = np.random.default_rng(5)
    V = list(encoder.VOCAB)
    cases = ['', '_', '_ _', 'k ky', 'a a:', 'a a a', 'a a a a', 'a b a b', 'a b a b a b c', 'x  x', 'ab ab\nab ab', 'あ あ']
    for _ in range(1500):
        toks = []
        for _ in range(int(rng.integers(0, 40))):
            toks += [V[int(rng.integers(0, rng.choice([3, 6, 39])))]] * int(rng.integers(1, 6))
        text = ' '.join(toks)
        if rng.random() < 0.1:
            text = text.replace(' ', '  ', 1)
        cases.append(text)
    for text in cases:
        assert encoder.merge_repeated(text) == ref(text), repr(text)
