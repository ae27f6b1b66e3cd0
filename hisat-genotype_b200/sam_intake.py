"""Native SAM intake: the aligner's SAM text -> per-locus, name-grouped alignment text, without samtools.

Replaces `samtools view <bam> <backbone> | sort -k1,1 -s` (reference hisatgenotype_typing_core.py:436-468) and the BAM
round trip in front of it (hisatgenotype_typing_common.py:1038-1054) for every locus of a sample in one pass (libhgt
hgt_sam_split_*, host threads)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def split_sam(sam_text, ref_names, n_threads=0, pinned=None, drop_qual=False):
    """sam_text: bytes of SAM (headers allowed, any record order).  ref_names: backbone names (RNAME), e.g. "A*BACKBONE".
    Returns {ref_name: bytes}, or with pinned=_lib.PinnedText {ref_name: (address, n_bytes)} written into it.
    drop_qual: write '*' for the QUAL column (never read by typing; 27 % fewer bytes to copy to the GPU for 2x100 bp records)."""
    if isinstance(sam_text, str):
        sam_text = sam_text.encode()
    L = _lib.lib()
    n = len(ref_names)
    enc = [r.encode() for r in ref_names]
    arr = (ctypes.c_char_p * n)(*enc)
    h = ctypes.c_void_p()
    L.hgt_sam_split_create_opts.restype = ctypes.c_int
    L.hgt_sam_split_create_opts.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p),
                                            ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]
    rc = L.hgt_sam_split_create_opts(sam_text, len(sam_text), n, arr, int(n_threads), 1 if drop_qual else 0, ctypes.byref(h))
    if rc == _lib.HGT_ERR_PARSE:
        raise AssertionError(_lib.last_error())
    _lib.check(rc)
    try:
        sizes = (ctypes.c_size_t * n)()
        lines = (ctypes.c_int64 * n)()
        _lib.check(L.hgt_sam_split_sizes(h, sizes, lines))
        out = {}
        for r, name in enumerate(ref_names):
            if pinned is not None:
                addr, nb = pinned.reserve(sizes[r])
                _lib.check(L.hgt_sam_split_write(h, r, ctypes.c_void_p(addr)))
                out[name] = (addr, nb)
            else:
                buf = np.empty(max(int(sizes[r]), 1), np.uint8)
                _lib.check(L.hgt_sam_split_write(h, r, buf.ctypes.data_as(ctypes.c_void_p)))
                out[name] = buf[:sizes[r]].tobytes()
        return out
    finally:
        L.hgt_sam_split_free(h)


def hisat2_command(simulation, index_name, base_fname, read_fname, fastq, threads):
    """The command line of the reference's align_reads() for aligner == "hisat2" on a graph index
    (hisatgenotype_typing_common.py:995-1024), without the samtools stages behind it."""
    cmd = ["hisat2", "--mm"]
    if not simulation:
        cmd += ["--no-unal"]
    cmd += ["--no-spliced-alignment", "-X", "1000", "--max-altstried", "64", "--haplotype"]
    if base_fname == "codis":
        cmd += ["--enable-codis", "--no-softclip"]
    cmd += ["-x", index_name]
    assert len(read_fname) in (1, 2)
    cmd += ["-p", str(threads)]
    if not fastq:
        cmd += ["-f"]
    if len(read_fname) == 1:
        cmd += ["-U", read_fname[0]]
    else:
        cmd += ["-1", "%s" % read_fname[0], "-2", "%s" % read_fname[1]]
    return cmd


def align_to_sam(simulation, index_name, base_fname, read_fname, fastq, threads, verbose=0):
    """HISAT2's SAM text straight from its stdout: no `samtools view -bS`, `sort`, `index` (common:1033-1056)."""
    import subprocess
    import sys
    cmd = hisat2_command(simulation, index_name, base_fname, read_fname, fastq, threads)
    if verbose >= 1:
        print(" ".join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    return proc.stdout


def read_alignment_text(path):
    """SAM text of an alignment file; None when the file is BGZF / BAM (it then goes through samtools, like the reference)."""
    with open(path, "rb") as f:
        head = f.read(2)
        if head == b"\x1f\x8b":
            return None
        return head + f.read()
