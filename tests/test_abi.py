"""CPU-side checks: the C-ABI library loads and exports every symbol include/hgt.h declares."""
import ctypes
import os
import re

from conftest import ROOT

SO = os.path.join(ROOT, "hisat-genotype_b200", "libhgt.so")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hgt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgt_[a-z0-9_]+)\s*\(", text)))


def _ensure_built():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(SO)


def test_header_symbols_exported():
    _ensure_built()
    lib = ctypes.CDLL(SO)
    syms = declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), "libhgt.so does not export %s" % s


def test_row_pitch_and_version():
    _ensure_built()
    lib = ctypes.CDLL(SO)
    lib.hgt_row_pitch.restype = ctypes.c_int
    assert lib.hgt_abi_version() >= 1
    for a, w in [(1, 2), (64, 2), (128, 2), (129, 4), (7000, 110), (8192, 128), (8193, 130)]:
        assert lib.hgt_row_pitch(a) == w


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    _ensure_built()
    lib = ctypes.CDLL(SO)
    lib.hgt_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    rc = lib.hgt_init(0, ctypes.byref(h))
    assert rc != 0 and h.value is None
    assert lib.hgt_last_error() != b""
