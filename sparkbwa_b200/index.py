"""FM-index construction: a thin Python wrapper of the native builder.

The work is done by `b200bwa_index_build` of libbwa.so (sparkbwa_b200/csrc/b2_index.cu + b2_host.cpp:
`bwa index` twin, bwa/bwtindex.c:210-324): the FASTA is packed exactly as bns_fasta2bntseq does
(bwa/bntseq.c:228-328, ambiguous bases replaced by a replay of lrand48() with srand48(11)), the suffixes of
forward + reverse-complement text are sorted on the GPU by a hand-written radix sort, group by group, and the
BWT / Occ / sampled-SA files are written in the layout of bwa/bwt.c:385-405 -- byte-identical to the
reference's `.pac/.ann/.amb/.bwt/.sa` (tests/test_cpu.py::test_index_builder_matches_bwa_index,
tests/test_gpu.py::test_index_builder_matches_bwa_index).

`build_index(contigs, prefix)` keeps the interface the benchmark and the tests use: contigs as
(name, uint8 codes 0..3 [4 = N]) pairs, written to a scratch FASTA next to the prefix.
"""
from __future__ import annotations

import ctypes
import os

from . import synth

_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbwa.so")


def build_index_fasta(fasta: str, prefix: str | None = None, device: int | None = None, lib_path: str | None = None) -> None:
    """`bwa index [-p prefix] fasta` on the GPU; raises RuntimeError with the library's message on failure."""
    lib = ctypes.CDLL(lib_path or os.environ.get("B200BWA_LIB", _LIB))
    lib.b200bwa_index_build.restype = ctypes.c_int
    lib.b200bwa_last_error.restype = ctypes.c_char_p
    args = ["index"] + (["-p", prefix] if prefix else []) + [fasta]
    arr = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
    old = os.environ.get("B200BWA_DEVICE")
    if device is not None:
        os.environ["B200BWA_DEVICE"] = str(device)
    try:
        rc = lib.b200bwa_index_build(len(args), arr)
    finally:
        if device is not None:
            if old is None:
                os.environ.pop("B200BWA_DEVICE", None)
            else:
                os.environ["B200BWA_DEVICE"] = old
    if rc != 0:
        raise RuntimeError("index construction failed: " + lib.b200bwa_last_error().decode())


def build_index(contigs, prefix: str, device: int | None = None, keep_fasta: bool = False) -> None:
    """contigs: list of (name, uint8 codes 0..3).  Writes prefix.{pac,ann,amb,bwt,sa}."""
    fa = prefix + ".build.fa"
    synth.write_fasta(fa, contigs)
    try:
        build_index_fasta(fa, prefix, device)
    finally:
        if not keep_fasta:
            try:
                os.remove(fa)
            except OSError:
                pass
