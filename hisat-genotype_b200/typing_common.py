"""Host-side mirror of the reference's hisatgenotype_typing_common API for the hot path.

`single_abundance` keeps the reference's name, arguments, return shape and error behaviour
(reference hisatgenotype_modules/hisatgenotype_typing_common.py:1282-1410) and runs the EM on the GPU through
libhgt (csrc/em.cu).  No CPU fallback exists."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

FK_NONE = 0x7FFFFFFF


def _index_alleles(keys):
    names = set()
    for k in keys:
        names.update(k.split("-"))
    names = sorted(names)
    return names, {n: i for i, n in enumerate(names)}


def em_arrays(class_bits, class_count, n_alleles, allele_len=None, remove_low=False, device=None):
    """EM on packed inputs.  Returns (prob[A], in_result[A], first_class[A], iters)."""
    L = _lib.lib()
    C = int(class_bits.shape[0])
    wp = _lib.row_pitch(n_alleles)
    bits = np.ascontiguousarray(class_bits, np.uint64)
    assert bits.shape == (C, wp)
    cnt = np.ascontiguousarray(class_count, np.int64)
    ln = None if allele_len is None else np.ascontiguousarray(allele_len, np.float64)
    prob = np.zeros(n_alleles, np.float64)
    inres = np.zeros(n_alleles, np.uint8)
    fk = np.zeros(n_alleles, np.int32)
    iters = ctypes.c_int32(0)
    rc = L.hgt_em(_lib.ctx(device), _lib.ptr(bits), _lib.ptr(cnt), C, n_alleles, wp, _lib.ptr(ln),
                  1 if remove_low else 0, _lib.ptr(prob), _lib.ptr(inres), _lib.ptr(fk), ctypes.byref(iters))
    if rc == _lib.HGT_ERR_KEY:
        raise KeyError(_lib.last_error())
    if rc == _lib.HGT_ERR_ZERODIV:
        raise ZeroDivisionError("float division by zero")
    _lib.check(rc)
    return prob, inres, fk, iters.value


def rank_result(names, prob, in_result, first_class, max_n=None):
    """[[allele, prob], ...] sorted like `sorted(..., key=prob, reverse=True)` on the reference's dict:
    ties keep dict insertion order = (first class that touched the allele, position inside its key)."""
    idx = np.nonzero(in_result)[0]
    prob = np.asarray(prob, np.float64)
    order = idx[np.lexsort((idx, np.asarray(first_class)[idx].astype(np.int64), -prob[idx]))]  # last key is the primary one
    if max_n is not None:
        order = order[:max_n]
    return [[names[i], float(prob[i])] for i in order.tolist()]


def single_abundance(Gene_cmpt, remove_low_abundance_allele=False, Gene_length={}):
    keys = list(Gene_cmpt.keys())
    if not keys:
        return []
    names, index = _index_alleles(keys)
    A = len(names)
    bits = _lib.pack_bits([[index[a] for a in k.split("-")] for k in keys], A)
    cnt = np.asarray([Gene_cmpt[k] for k in keys], np.int64)
    ln = None
    if len(Gene_length) > 0:
        for a in names:
            assert a in Gene_length
        ln = np.asarray([Gene_length[a] for a in names], np.float64)
    prob, inres, fk, _ = em_arrays(bits, cnt, A, ln, bool(remove_low_abundance_allele))
    return rank_result(names, prob, inres, fk)
