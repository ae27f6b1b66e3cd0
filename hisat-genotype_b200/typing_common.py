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
    cnt = np.ascontiguousarray(class_count, np.float64)  # the reference does float(count): fractional counts are legal
    ln = None if allele_len is None else np.ascontiguousarray(allele_len, np.float64)
    prob = np.zeros(n_alleles, np.float64)
    inres = np.zeros(n_alleles, np.uint8)
    fk = np.zeros(n_alleles, np.int32)
    iters = ctypes.c_int32(0)
    rc = L.hgt_em_f64(_lib.ctx(device), _lib.ptr(bits), _lib.ptr(cnt), C, n_alleles, wp, _lib.ptr(ln),
                  1 if remove_low else 0, _lib.ptr(prob), _lib.ptr(inres), _lib.ptr(fk), ctypes.byref(iters))
    if rc == _lib.HGT_ERR_KEY:
        raise KeyError(_lib.last_error())
    if rc == _lib.HGT_ERR_ZERODIV:
        raise ZeroDivisionError("float division by zero")
    _lib.check(rc)
    return prob, inres, fk, iters.value


def rank_result(names, prob, in_result, first_class, max_n=None, key_pos=None):
    """[[allele, prob], ...] sorted like `sorted(..., key=prob, reverse=True)` on the reference's dict:
    ties keep dict insertion order = (first class that touched the allele, position inside its key).  key_pos[a] = that
    position when the keys are not name-sorted (default: the sorted-name index, which is the position for sorted keys)."""
    idx = np.nonzero(in_result)[0]
    prob = np.asarray(prob, np.float64)
    third = idx if key_pos is None else np.asarray(key_pos)[idx]
    order = idx[np.lexsort((third, np.asarray(first_class)[idx].astype(np.int64), -prob[idx]))]  # last key is the primary one
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
    cnt = np.asarray([Gene_cmpt[k] for k in keys], np.float64)
    ln = None
    if len(Gene_length) > 0:
        for a in names:
            assert a in Gene_length
        ln = np.asarray([Gene_length[a] for a in names], np.float64)
    prob, inres, fk, _ = em_arrays(bits, cnt, A, ln, bool(remove_low_abundance_allele))
    # ties: position of the allele inside the key of the first class that holds it (keys need not be name-sorted)
    key_pos = np.arange(A)
    for a in np.nonzero(inres)[0].tolist():
        if 0 <= fk[a] < len(keys):
            key_pos[a] = keys[fk[a]].split("-").index(names[a])
    return rank_result(names, prob, inres, fk, key_pos=key_pos)


MAX_PAIRS = 1 << 22


def joint_abundance(HLA_cmpt, HLA_length=None, device=None, return_iters=False):
    """Diploid allele-pair model: drop-in for joint_abundance(HLA_cmpt, HLA_length) of the reference's legacy typer
    (etc/hisatgenotype_hla_cyp.py:236-302; Python 2 there).  Returns [["a1-a2", prob], ...] sorted by probability
    (descending), ties by pair name (HLA_prob_cmp :167-176).  The host enumerates the pairs that survive the first
    choose_top_alleles (:259-270) from the per-allele masses - pair mass = m[a] + m[b], m[a] alone for the pair (a, a),
    :243-254 - and the GPU runs the loop (libhgt hgt_pair_em).  HLA_length is unused, as in the reference."""
    keys = list(HLA_cmpt.keys())
    if not keys:
        return ({}, 0) if return_iters else {}
    names, index = _index_alleles(keys)
    A = len(names)
    members = [[index[a] for a in k.split("-")] for k in keys]
    bits = _lib.pack_bits(members, A)
    cnt = np.asarray([HLA_cmpt[k] for k in keys], np.int64)
    # per-allele mass; a class lists an allele once per occurrence in its key (the legacy loop walks cmpt.split('-'))
    m = np.zeros(A, np.float64)
    for mem, c in zip(members, cnt.tolist()):
        share = float(c) / len(mem)
        for a in mem:
            m[a] += share
    order = np.argsort(-m, kind="stable")
    ms = m[order]
    best = ms[0] + ms[1] if A >= 2 else ms[0]
    best = max(best, ms[0])
    # kept pairs: mass * 2 > best  (the sorted loop of choose_top_alleles stops at the first mass * 2 <= best)
    pa, pb, pm = [], [], []
    total = 0
    for i in range(A):
        if (ms[i] + ms[0]) * 2 <= best and ms[i] * 2 <= best:
            break
        hi = int(np.searchsorted(-ms, -(best / 2.0 - ms[i]), side="left"))  # ms[j] > best/2 - ms[i]  <=>  j < hi
        js = np.arange(i + 1, max(hi, i + 1))
        if js.size:
            mass = ms[i] + ms[js]
            ok = ~(mass * 2 <= best)
            js, mass = js[ok], mass[ok]
            total += js.size
            if total > MAX_PAIRS:
                raise _lib.HgtError(_lib.HGT_ERR_UNSUPPORTED, "joint_abundance: more than %d allele pairs survive the first "
                                    "pruning" % MAX_PAIRS)
            pa.append(np.full(js.size, order[i]))
            pb.append(order[js])
            pm.append(mass)
        if not (ms[i] * 2 <= best):  # the pair (a, a)
            pa.append(np.asarray([order[i]]))
            pb.append(np.asarray([order[i]]))
            pm.append(np.asarray([ms[i]]))
            total += 1
    pa, pb, pm = np.concatenate(pa), np.concatenate(pb), np.concatenate(pm)
    lo, hi = np.minimum(pa, pb), np.maximum(pa, pb)  # names are sorted: index order = string order of the key "x-y"
    Q = int(lo.size)
    # alleles whose NAME occurs inside a kept allele's name (`allele in allele_pair`, :280 / :284)
    sub = {}
    for a in set(lo.tolist()) | set(hi.tolist()):
        na = names[a]
        sub[a] = [x for x in range(A) if names[x] in na]
    ubits = _lib.pack_bits([sorted(set(sub[a]) | set(sub[b])) for a, b in zip(lo.tolist(), hi.tolist())], A)
    p0 = np.ascontiguousarray(pm / pm.sum(), np.float64)
    prob = np.zeros(Q, np.float64)
    iters = ctypes.c_int32(0)
    wp = _lib.row_pitch(A)
    _lib.check(_lib.lib().hgt_pair_em(_lib.ctx(device), _lib.ptr(np.ascontiguousarray(bits)), _lib.ptr(cnt), len(keys), wp,
                                      _lib.ptr(np.ascontiguousarray(ubits)), Q, _lib.ptr(p0), _lib.ptr(prob), ctypes.byref(iters)))
    out = [["%s-%s" % (names[a], names[b]), float(p)] for a, b, p in zip(lo.tolist(), hi.tolist(), prob.tolist()) if p > 0.0]
    out.sort(key=lambda x: (-x[1], x[0]))
    return (out, iters.value) if return_iters else out
