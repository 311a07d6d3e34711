"""ctypes front end of oracle/ctc_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header).

``ctc_best_path`` has the reference's signature and exceptions (align.py:43-109);
``ctc_best_path_batch`` has the flat offset-array layout of the product C-ABI.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libctc_oracle.so")
_lib = None

STATUS_OK, STATUS_DEAD_BAND, STATUS_BAD_LABEL, STATUS_NONFINITE = 0, 1, 2, 3


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "ctc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libctc_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        f32p, i32p, i64p = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32),
                            ctypes.POINTER(ctypes.c_int64))
        L.oracle_ctc_best_path.restype = ctypes.c_int
        L.oracle_ctc_best_path.argtypes = [f32p, ctypes.c_int64, ctypes.c_int32, i32p,
                                           ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                           i32p, i32p, f32p, f32p]
        L.oracle_ctc_best_path_batch.restype = ctypes.c_int
        L.oracle_ctc_best_path_batch.argtypes = [f32p, i64p, i32p, i64p, ctypes.c_int64,
                                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                 i32p, i32p, f32p, f32p, i32p, ctypes.c_int32]
        L.oracle_cells_eval.restype = ctypes.c_int64
        L.oracle_cells_eval.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int32]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def raise_for_status(status, V=None):
    """Map a status code to the exception the reference raises."""
    if status == STATUS_OK:
        return
    if status == STATUS_DEAD_BAND:
        raise ValueError("attempt to get argmax of an empty sequence")
    if status == STATUS_BAD_LABEL:
        raise IndexError(f"label out of bounds for axis 1 with size {V}")
    if status == STATUS_NONFINITE:
        raise ValueError("log_probs must be finite")
    raise RuntimeError(f"oracle error {status}")


def ctc_best_path(log_probs, labels, beam_size=1000, max_move=4, return_final_score=False):
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    lab = np.ascontiguousarray(labels).astype(np.int32)
    T, V = lp.shape
    if T == 0:
        raise IndexError("list index out of range")
    if max_move < 1:
        raise ValueError("attempt to get argmax of an empty sequence")
    path = np.empty(T, np.int32)
    labs = np.empty(T, np.int32)
    scores = np.empty(T, np.float32)
    fs = np.zeros(1, np.float32)
    st = lib().oracle_ctc_best_path(_p(lp, ctypes.c_float), T, V, _p(lab, ctypes.c_int32),
                                    lab.shape[0], int(beam_size), int(max_move),
                                    _p(path, ctypes.c_int32), _p(labs, ctypes.c_int32),
                                    _p(scores, ctypes.c_float), _p(fs, ctypes.c_float))
    raise_for_status(st, V)
    if return_final_score:
        return path, labs, scores, fs[0]
    return path, labs, scores


def ctc_best_path_batch(log_probs, t_off, labels, l_off, beam_size=1000, max_move=4,
                        n_threads=1):
    """Flat batch: log_probs [sum T, V] f32, labels [sum L] i32, offsets int64 [B+1]."""
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    t_off = np.ascontiguousarray(t_off, dtype=np.int64)
    l_off = np.ascontiguousarray(l_off, dtype=np.int64)
    B = t_off.shape[0] - 1
    n, V = lp.shape
    path = np.empty(n, np.int32)
    labs = np.empty(n, np.int32)
    scores = np.empty(n, np.float32)
    fs = np.empty(B, np.float32)
    status = np.empty(B, np.int32)
    lib().oracle_ctc_best_path_batch(_p(lp, ctypes.c_float), _p(t_off, ctypes.c_int64),
                                     _p(lab, ctypes.c_int32), _p(l_off, ctypes.c_int64), B, V,
                                     int(beam_size), int(max_move), _p(path, ctypes.c_int32),
                                     _p(labs, ctypes.c_int32), _p(scores, ctypes.c_float),
                                     _p(fs, ctypes.c_float), _p(status, ctypes.c_int32),
                                     int(n_threads))
    return path, labs, scores, fs, status


def cells_eval(T, L, beam_size=1000):
    return int(lib().oracle_cells_eval(T, L, beam_size))
