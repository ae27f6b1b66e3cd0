"""ctypes wrapper around oracle/libem_oracle.so (C restatement of single_abundance) — TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libem_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "em_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = ctypes.CDLL(so)
        _lib.em_oracle_single_abundance.restype = ctypes.c_int
        _lib.em_oracle_iterations.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def csr_from_cmpt(cmpt_items, names=None):
    """cmpt_items: [(key, count)] in dict order.  Returns (names, off, mem, cnt); allele index = position in
    `names` (sorted names unless given)."""
    if names is None:
        s = set()
        for k, _ in cmpt_items:
            s.update(k.split("-"))
        names = sorted(s)
    idx = {n: i for i, n in enumerate(names)}
    off, mem, cnt = [0], [], []
    for k, c in cmpt_items:
        mem.extend(idx[a] for a in k.split("-"))
        off.append(len(mem))
        cnt.append(float(c))
    return names, np.asarray(off, np.int64), np.asarray(mem, np.int32), np.asarray(cnt, np.float64)


def single_abundance_csr(A, off, mem, cnt, lengths=None, remove_low=False):
    out_a = np.zeros(max(A, 1), np.int32)
    out_p = np.zeros(max(A, 1), np.float64)
    iters = ctypes.c_int32(0)
    ln = np.ascontiguousarray(lengths, np.float64) if lengths is not None else None
    n = lib().em_oracle_single_abundance(
        ctypes.c_int(len(cnt)), ctypes.c_int(A), _p(off, ctypes.c_int64), _p(mem, ctypes.c_int32),
        _p(cnt, ctypes.c_double), _p(ln, ctypes.c_double), ctypes.c_int(1 if remove_low else 0),
        _p(out_a, ctypes.c_int32), _p(out_p, ctypes.c_double), ctypes.byref(iters))
    if n == -2:
        raise KeyError("allele vanished during SQUAREM step")
    if n == -3:
        raise ZeroDivisionError("float division by zero")
    return out_a[:n].copy(), out_p[:n].copy(), iters.value


def single_abundance(cmpt, remove_low=False, lengths=None):
    """Same signature and result shape as the reference's single_abundance (common:1282-1284)."""
    items = list(cmpt.items()) if isinstance(cmpt, dict) else list(cmpt)
    names, off, mem, cnt = csr_from_cmpt(items)
    ln = np.asarray([lengths[n] for n in names], np.float64) if lengths else None
    a, p, it = single_abundance_csr(len(names), off, mem, cnt, ln, remove_low)
    return [[names[i], float(x)] for i, x in zip(a, p)], it


def time_iterations(A, off, mem, cnt, lengths, n_iters):
    chk = ctypes.c_double(0.0)
    ln = np.ascontiguousarray(lengths, np.float64) if lengths is not None else None
    rc = lib().em_oracle_iterations(ctypes.c_int(len(cnt)), ctypes.c_int(A), _p(off, ctypes.c_int64),
                                    _p(mem, ctypes.c_int32), _p(cnt, ctypes.c_double), _p(ln, ctypes.c_double),
                                    ctypes.c_int(n_iters), ctypes.byref(chk))
    return rc, chk.value
