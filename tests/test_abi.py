"""CPU-side checks of the C-ABI boundary: the library loads, exports every declared symbol, and the
alignment entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from poasta_b200 import _lib, CudaError, POAGraph, PoastaAligner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "poasta_b200.h")).read()
    declared = set(re.findall(r"\b(poasta_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_version_string():
    assert b"sm_100a" in _lib.load().poasta_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    g = POAGraph()
    g.add_alignment_with_weights("s", b"ACGT", None)
    with pytest.raises(CudaError):
        PoastaAligner().align(g, b"ACGT")  # graphs as handles: poasta_align_batch_graphs
    from poasta_b200 import TIEBREAK_DP, align_batch, msa_batch
    with pytest.raises(CudaError):
        align_batch([g.view()], [b"ACGT"], [0])  # graphs as CSR views: poasta_align_batch
    for tiebreak in (0, TIEBREAK_DP):  # the engine and the device-resident DP rounds
        with pytest.raises(CudaError):
            msa_batch([b">a\nACGT\n>b\nACGA\n"], tiebreak=tiebreak)


def test_product_does_not_reference_oracle():
    """Nothing under poasta_b200/ may include, link or call oracle/ (the oracle is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poasta_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle/|liboracle|astar_oracle|dp_model", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_oracle_does_not_include_the_product():
    """The oracle library is built from oracle/ alone (plus the public C header for one struct): no product header or source."""
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".cpp", ".hpp", ".h")) or f == "Makefile":
            text = open(os.path.join(ROOT, "oracle", f), errors="ignore").read()
            assert not re.search(r'#include\s+"[^"]*poasta_b200/|csrc/', text), f
