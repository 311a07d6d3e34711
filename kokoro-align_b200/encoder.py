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
    return ' '.join(VOCAB[n] for n in encoded)


def merge_repeated(text):
    """Greedy CTC collapse of a decoded token string (the contract of encoder.py:26-31):
    runs of an identical token sequence collapse to one copy, then blanks are dropped."""
    import re
    collapsed = re.sub(r'(.+)( \1)+', r'\1', text)
    collapsed = collapsed.replace(' _', '').replace('_ ', '')
    return '' if collapsed == '_' else collapsed
