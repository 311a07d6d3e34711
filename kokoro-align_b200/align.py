"""Host-side mirror of the reference's best-path interface (kokoro_align/align.py).

Same names, arguments, return values and exception types as the reference:

    ctc_best_path(log_probs, labels, beam_size=1000, max_move=4)      align.py:43-109
    best_path(input_file, voca_file, output_file)                     align.py:112-124

backed by the sm_100a kernels through the C ABI (include/kokoro_align_b200.h).  The batched
form (``AlignPlan``) is what a B200 pipeline uses: many independent chapter / segment
lattices per launch, inputs optionally already resident in HBM.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import ST_BAD_LABEL, ST_DEAD_BAND, ST_NONFINITE, ST_OK, KabError, PlanInfo  # noqa: F401


def raise_for_status(status, V=None):
    """Turn a per-lattice status into the exception the reference raises."""
    if status == ST_OK:
        return
    if status == ST_DEAD_BAND:   # np.argmax of an empty array, align.py:101
        raise ValueError("attempt to get argmax of an empty sequence")
    if status == ST_BAD_LABEL:   # log_probs[i, labels[v]], align.py:77
        raise IndexError(f"index out of bounds for axis 1 with size {V}")
    if status == ST_NONFINITE:
        raise ValueError("log_probs must be finite")
    raise KabError(f"unknown lattice status {status}")


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


class AlignPlan:
    """A fixed batch of transcripts + frame counts, classified and resident on one GPU.

    t_off / l_off are int64 [B+1] row / label offsets into the flat ``log_probs [sum T, V]``
    and ``labels [sum L]`` arrays (the C-ABI "flat batch" layout).
    """

    def __init__(self, t_off, labels, l_off, vocab_size, beam_size=1000, max_move=4, device=0):
        self.t_off = np.ascontiguousarray(t_off, dtype=np.int64)
        self.l_off = np.ascontiguousarray(l_off, dtype=np.int64)
        self.labels = np.ascontiguousarray(labels).astype(np.int32, copy=False)
        self.labels = np.ascontiguousarray(self.labels)
        self.B = int(self.t_off.shape[0] - 1)
        self.V = int(vocab_size)
        self.total_T = int(self.t_off[-1]) if self.B else 0
        self.device = int(device)
        self._h = ctypes.c_void_p(0)
        L = _lib.lib()
        _lib.check(L.kab_plan_create(ctypes.byref(self._h), self.device, self.B, _ptr(self.t_off),
                                     _ptr(self.labels), _ptr(self.l_off), self.V, int(beam_size),
                                     int(max_move)))
        info = PlanInfo()
        _lib.check(L.kab_plan_get_info(self._h, ctypes.byref(info)))
        self.info = info

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and _lib._lib is not None:   # (module globals vanish at exit)
            _lib._lib.kab_plan_destroy(h)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- host buffers in, host buffers out (H2D + kernels + D2H inside the call)
    def run_host(self, log_probs, out=None):
        lp = np.ascontiguousarray(log_probs, dtype=np.float32)
        if lp.shape != (self.total_T, self.V):
            raise ValueError(f"log_probs must have shape {(self.total_T, self.V)}, got {lp.shape}")
        if out is None:
            out = (np.empty(self.total_T, np.int32), np.empty(self.total_T, np.int32),
                   np.empty(self.total_T, np.float32), np.empty(self.B, np.float32),
                   np.empty(self.B, np.int32))
        path, labs, scores, final, status = out
        _lib.check(_lib.lib().kab_plan_run_host(self._h, _ptr(lp), _ptr(path), _ptr(labs),
                                                _ptr(scores), _ptr(final), _ptr(status)))
        return path, labs, scores, final, status

    # -- device pointers (e.g. torch tensors' data_ptr()), asynchronous on `stream`
    def run_device(self, d_log_probs, d_best_path, d_best_labels, d_best_scores, d_final_score,
                   d_status, stream=0):
        _lib.check(_lib.lib().kab_plan_run_device(self._h, ctypes.c_void_p(d_log_probs),
                                                  ctypes.c_void_p(d_best_path),
                                                  ctypes.c_void_p(d_best_labels),
                                                  ctypes.c_void_p(d_best_scores),
                                                  ctypes.c_void_p(d_final_score),
                                                  ctypes.c_void_p(d_status), ctypes.c_void_p(stream)))

    def run_torch(self, log_probs, stream=None):
        """log_probs: CUDA float32 tensor [sum T, V] on this plan's device.  Returns CUDA tensors
        (best_path, best_labels, best_scores, final_score, status); asynchronous."""
        import torch
        assert log_probs.is_cuda and log_probs.dtype == torch.float32 and log_probs.is_contiguous()
        dev = log_probs.device
        n = self.total_T
        path = torch.empty(n, dtype=torch.int32, device=dev)
        labs = torch.empty(n, dtype=torch.int32, device=dev)
        scores = torch.empty(n, dtype=torch.float32, device=dev)
        final = torch.empty(self.B, dtype=torch.float32, device=dev)
        status = torch.empty(self.B, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev) if stream is None else stream
        self.run_device(log_probs.data_ptr(), path.data_ptr(), labs.data_ptr(), scores.data_ptr(),
                        final.data_ptr(), status.data_ptr(), s.cuda_stream)
        return path, labs, scores, final, status


def ctc_best_path(log_probs, labels, beam_size=1000, max_move=4, return_final_score=False):
    """Drop-in for kokoro_align.align.ctc_best_path (align.py:43-109).

    log_probs: [T, V] float32 (other float dtypes are cast to float32 first -- the reference's
    own path is float32 end to end, align.py:113-117); labels: [L] any integer dtype (int8 from
    encoder.encode_text).  Returns (best_path int32 [T], best_labels int32 [T],
    best_scores float32 [T]).  Raises ValueError when no state is active in the last frame and
    IndexError for a label outside [-V, V), as the reference does.
    """
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    if lp.ndim != 2:
        raise ValueError("log_probs must be [T, V]")
    labels = np.ascontiguousarray(labels)
    T, V = lp.shape
    if T == 0:   # beams[-1] on an empty list, align.py:100
        raise IndexError("list index out of range")
    if max_move < 1 or beam_size < 0:   # argmax over an empty move axis / negative dimensions
        raise ValueError("attempt to get argmax of an empty sequence")
    with AlignPlan([0, T], labels, [0, labels.shape[0]], V, beam_size, max_move,
                   device=_current_device()) as plan:
        path, labs, scores, final, status = plan.run_host(lp)
    raise_for_status(int(status[0]), V)
    if return_final_score:
        return path, labs, scores, final[0]
    return path, labs, scores


def _current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


def log_softmax(logits):
    """The reference's normalisation, align.py:116-117, kept on the host in fp32 numpy so that
    the parity boundary (the log_probs array) is bit-identical."""
    logits = logits - np.mean(logits, axis=-1, keepdims=True)
    return logits - np.log(np.sum(np.exp(logits), axis=-1, keepdims=True))


def read_transcript_labels(voca_file):
    """Labels of a ``text|voca`` transcript: transcript.py:60-67 + encoder.py:5-19 (tokens
    outside the 39-symbol vocabulary are dropped, int8 ids)."""
    from .encoder import encode_text
    res = []
    with open(voca_file) as f:
        for line in f:
            parts = line.rstrip('\r\n').split('|')
            res.append(parts[1])
    return encode_text(' '.join(res))


def best_path(input_file, voca_file, output_file):
    """Drop-in for kokoro_align.align.best_path (align.py:112-124): raw logits npz + voca.txt
    -> best_path.npz {best_path, best_labels, best_scores}."""
    with np.load(input_file) as f:
        logits = f['data']
    log_probs = log_softmax(logits)
    labels = read_transcript_labels(voca_file)
    path, labs, scores = ctc_best_path(log_probs, labels)
    np.savez(output_file, best_path=path, best_labels=labs, best_scores=scores)


def best_path_files(logits_files, voca_files, best_path_files, skip_existing=True, beam_size=1000,
                    max_move=4, verbose=True):
    """The per-book loop of run_example.py:247-254 as ONE batch: every chapter whose output does
    not exist yet is loaded, normalised (align.py:116-117), and all of them are aligned in a single
    plan -- the chapters are independent lattices, so they run side by side on the GPU (36
    chapters of a 9-hour book take the time of the longest one).  Writes the same best_path.npz
    files, prints the reference's two messages, and raises the reference's exception for the first
    chapter that fails (after the other chapters have been written).  Returns the written paths.

    Under torch.distributed (one process per GPU) the chapters are sharded over the ranks by
    cost (parallel.align_sharded: no collective on the data path, a host-side gather of the
    results) and rank 0 writes the files."""
    import os
    todo = []
    for lf, vf, bf in zip(logits_files, voca_files, best_path_files):
        if skip_existing and os.path.exists(bf):
            if verbose:
                print(f'Skip writing {bf}')
        else:
            todo.append((lf, vf, bf))
    if not todo:
        return []
    lps, labs = [], []
    for lf, vf, _ in todo:
        with np.load(lf) as f:
            lps.append(np.ascontiguousarray(log_softmax(f['data']), dtype=np.float32))
        labs.append(np.asarray(read_transcript_labels(vf), dtype=np.int32))
    from . import parallel
    V = lps[0].shape[1]
    results = parallel.align_sharded(lps, labs, beam_size=beam_size, max_move=max_move,
                                     device=_current_device())
    if results is None:  # not the gathering rank
        return []
    written, first_bad = [], ST_OK
    for (_, _, bf), (path, lab, sc, _, st) in zip(todo, results):
        if st != ST_OK:
            first_bad = first_bad or st
            continue
        if verbose:
            print(f'Writing {bf}')
        np.savez(bf, best_path=path, best_labels=lab, best_scores=sc)
        written.append(bf)
    raise_for_status(first_bad, V)
    return written


# --------------------------------------------------------------------------- the consumer
# the step that reads best_path.npz (align.py:127-169).  Host-side text
# work; it defines the output contract of the kernels (best_path // 2 at the segment
# boundaries, per-segment sums of best_scores), so it lives beside them.

_PUNCT = (',', '.', '!', '?')


class TokenTable:
    """label index -> token index table of a ``text|voca`` transcript (the behaviour of
    transcript.py:13-57): a token owns the labels from the middle of the previous voiced token
    to its own middle; punctuation-only tokens move the boundary without owning labels."""

    def __init__(self, voca_file):
        from .encoder import encode_text
        self.texts, self.vocas, self.table = [], [], []
        labels_seen, boundary = 0, 0
        with open(voca_file) as f:
            for n, line in enumerate(f):
                text, voca = line.rstrip('\r\n').split('|')
                self.texts.append(text)
                self.vocas.append(voca)
                k = len(encode_text(voca))
                if k:
                    labels_seen += k
                    upto = labels_seen - k // 2
                    self.table.extend([boundary] * (upto - len(self.table)))
                    boundary = n + 1
                elif voca in _PUNCT:
                    boundary = n + 1

    def __len__(self):
        return len(self.table)

    def get_token(self, start, end, remove_wordsep=True):
        import re
        lo = self.table[start] if start < len(self.table) else len(self.texts)
        hi = self.table[end] if end < len(self.table) else len(self.texts)
        text = ' '.join(t for t in self.texts[lo:hi] if t)
        parts = [t for t in self.vocas[lo:hi] if t]
        if remove_wordsep:
            voca = ' '.join(parts)
        else:
            voca = ' _ '.join(parts)
            voca = re.sub(r'_ ([.,!?])', r'\1', voca)
            voca = re.sub(r'([.,!?]) _', r'\1', voca)
        return text.strip(), voca.strip()


def align(best_path_file, mfcc_file, voca_file, align_file, remove_wordsep):
    """Drop-in for kokoro_align.align.align (align.py:127-169): one line per silence-split
    segment, ``audio_end|text|voca|decoded|non_blanks|non_blanks_score|all_score``."""
    import os
    from .encoder import decode_text, merge_repeated
    with np.load(best_path_file) as f:
        label_idx = f['best_path'] // 2            # extended-state index -> label index
        best_labels = f['best_labels']
        best_scores = f['best_scores']
    with np.load(mfcc_file) as f:
        ends = f['indices']
    table = TokenTable(voca_file)
    n_frames = len(label_idx)
    try:
        with open(align_file, 'wt') as out:
            for i in range(len(ends)):
                a = ends[i - 1] if i > 0 else 0
                b = ends[i]
                t0 = min(label_idx[a], len(table))
                t1 = min(label_idx[b] if b < n_frames else len(table), len(table))
                labels = best_labels[a:b]
                scores = best_scores[a:b]
                voiced = labels != 0
                decoded = merge_repeated(decode_text(labels))
                text, voca = table.get_token(t0, t1, remove_wordsep=remove_wordsep)
                out.write(f'{b}|{text}|{voca}|{decoded}|{np.sum(voiced).item()}|'
                          f'{np.sum(scores[voiced]).item()}|{np.sum(scores).item()}\n')
    except BaseException:
        os.unlink(align_file)
        raise
