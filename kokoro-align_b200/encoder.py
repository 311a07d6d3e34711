"""Phoneme vocabulary of the acoustic model (the data contract of encoder.py:5-19 in the
reference): 39 symbols, index 0 ('_') is the CTC blank; tokens outside the table are dropped
when a transcript is encoded; ids are int8."""
import numpy as np

VOCAB = ('_', 'N', 'a', 'a:', 'b', 'by', 'ch', 'd', 'e', 'e:', 'f', 'g', 'gy', 'h', 'hy', 'i',
         'i:', 'j', 'k', 'ky', 'm', 'my', 'n', 'ny', 'o', 'o:', 'p', 'py', 'r', 'ry', 's', 'sh',
         't', 'ts', 'u', 'u:', 'w', 'y', 'z')
VOCAB_SIZE = len(VOCAB)
_INDEX = {tok: n for n, tok in enumerate(VOCAB)}


def encode_text(text):
    get = _INDEX.get
    ids = [i for i in map(get, text.split()) if i is not None]
    return np.array(ids, dtype=np.int8)


def decode_text(encoded):
    return ' '.join(map(VOCAB.__getitem__, np.asarray(encoded).tolist()))


def merge_repeated(text):
    """Greedy CTC collapse of a decoded token string (the contract of encoder.py:26-31):
    runs of an identical token sequence collapse to one copy, then blanks are dropped.  The
    regular expression is evaluated by kab_merge_repeated in the C library (same leftmost /
    greedy semantics; the backtracking `re` engine needs ~30 ms per segment, 90 s per book);
    text it does not take (non-ASCII, newlines) goes through `re`."""
    collapsed = None
    if text.isascii() and '\n' not in text:
        import ctypes
        from . import _lib
        raw = text.encode('ascii')
        out = ctypes.create_string_buffer(max(1, len(raw)))
        n = ctypes.c_int64(0)
        if _lib.lib().kab_merge_repeated(raw, len(raw), out, ctypes.byref(n)) == _lib.KAB_OK:
            collapsed = out.raw[:n.value].decode('ascii')
    if collapsed is None:
        import re
        collapsed = re.sub(r'(.+)( \1)+', r'\1', text)
    collapsed = collapsed.replace(' _', '').replace('_ ', '')
    return '' if collapsed == '_' else collapsed
