"""ctypes front-end of oracle/oracle_ref.c (plain-C CPU oracle).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmts_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def crf_viterbi(emis, lengths, trans):
    emis = np.ascontiguousarray(emis, dtype=np.float32)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    trans = np.ascontiguousarray(trans, dtype=np.float32)
    B, L, C = emis.shape
    best = np.empty(B, dtype=np.float32)
    paths = np.empty((B, L), dtype=np.int32)
    lib().mtso_crf_viterbi(_p(emis, ctypes.c_float), _p(lengths, ctypes.c_int64), _p(trans, ctypes.c_float),
                           B, L, C, _p(best, ctypes.c_float), _p(paths, ctypes.c_int32))
    return best, paths


def crf_forward(emis, lengths, trans):
    emis = np.ascontiguousarray(emis, dtype=np.float32)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    trans = np.ascontiguousarray(trans, dtype=np.float32)
    B, L, C = emis.shape
    out = np.empty(B, dtype=np.float32)
    lib().mtso_crf_forward(_p(emis, ctypes.c_float), _p(lengths, ctypes.c_int64), _p(trans, ctypes.c_float),
                           B, L, C, _p(out, ctypes.c_float))
    return out


def crf_gold(emis, tags, lengths, trans):
    emis = np.ascontiguousarray(emis, dtype=np.float32)
    tags = np.ascontiguousarray(tags, dtype=np.int64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    trans = np.ascontiguousarray(trans, dtype=np.float32)
    B, L, C = emis.shape
    out = np.empty(B, dtype=np.float32)
    lib().mtso_crf_gold(_p(emis, ctypes.c_float), _p(tags, ctypes.c_int64), _p(lengths, ctypes.c_int64),
                        _p(trans, ctypes.c_float), B, L, C, _p(out, ctypes.c_float))
    return out


def lstm_dir(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    x = np.ascontiguousarray(x, dtype=np.float32)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    w_ih, w_hh, b_ih, b_hh = (np.ascontiguousarray(a, dtype=np.float32) for a in (w_ih, w_hh, b_ih, b_hh))
    B, T, D = x.shape
    H = w_hh.shape[1]
    out = np.empty((B, T, H), dtype=np.float32)
    f = ctypes.c_float
    lib().mtso_lstm_dir(_p(x, f), _p(lengths, ctypes.c_int64), _p(w_ih, f), _p(w_hh, f), _p(b_ih, f), _p(b_hh, f),
                        B, T, D, H, int(bool(reverse)), _p(out, f))
    return out


def bilstm_stack(x, lengths, params, prefix, num_layers):
    tmax = int(np.max(lengths))
    h = np.ascontiguousarray(np.asarray(x, dtype=np.float32)[:, :tmax])
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", 0), ("_reverse", 1)):
            g = lambda n: params[f"{prefix}{n}_l{layer}{suffix}"]
            outs.append(lstm_dir(h, lengths, g("weight_ih"), g("weight_hh"), g("bias_ih"), g("bias_hh"), rev))
        h = np.concatenate(outs, axis=-1)
    return h


def pk(hyp, ref):
    """-> (numerator, denominator) of segeval-style Pk with the last unit forced to be a boundary."""
    hyp = np.ascontiguousarray(hyp, dtype=np.uint8)
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    den = ctypes.c_int(0)
    num = lib().mtso_pk(_p(hyp, ctypes.c_uint8), _p(ref, ctypes.c_uint8), len(ref), ctypes.byref(den))
    return num, den.value


def window_diff(hyp, ref):
    hyp = np.ascontiguousarray(hyp, dtype=np.uint8)
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    den = ctypes.c_int(0)
    num = lib().mtso_wd(_p(hyp, ctypes.c_uint8), _p(ref, ctypes.c_uint8), len(ref), ctypes.byref(den))
    return num, den.value
