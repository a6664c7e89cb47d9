"""The PRODUCT's graph, writers and readers (poasta_b200/csrc/poa_graph.hpp, poa_io.hpp, through the C ABI) against the oracle's
INDEPENDENT restatement of graphs/poa.rs, io/fasta.rs and io/graph.rs (oracle/poa_oracle.hpp shares no code with the product):
the same insertion lists must give byte-identical FASTA-MSA, GFA, GFAv1 and .poasta output."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from poasta_b200 import InvalidAlignment, POAGraph, synth
from poasta_b200.graph import NONE

u32p = C.POINTER(C.c_uint32)


def _oracle_build(inserts, fmt, graph_in=None, msa_in=None):
    """inserts: list of (name, seq, pairs or None)."""
    lib = H.oracle()
    lib.oracle_graph_build.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), u32p, u32p, C.POINTER(C.c_int64), C.c_uint32, C.c_char_p,
                                       C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_int32)]
    n = len(inserts)
    names = (C.c_char_p * max(n, 1))(*[i[0].encode() for i in inserts])
    seqs = (C.c_char_p * max(n, 1))(*[bytes(i[1]) for i in inserts])
    lens = (C.c_uint32 * max(n, 1))(*[len(i[1]) for i in inserts])
    flat = []
    for _, _, pairs in inserts:
        for p in pairs or []:
            flat += [NONE if p.rpos is None else p.rpos, NONE if p.qpos is None else p.qpos]
    pairs_arr = np.array(flat + [0], dtype=np.uint32)
    npairs = (C.c_int64 * max(n, 1))(*[-1 if i[2] is None else len(i[2]) for i in inserts])
    out, out_len = C.c_void_p(), C.c_size_t()
    status = (C.c_int32 * max(n, 1))()
    rc = lib.oracle_graph_build(names, seqs, lens, pairs_arr.ctypes.data_as(u32p), npairs, n, graph_in, len(graph_in) if graph_in else 0,
                                msa_in, len(msa_in) if msa_in else 0, fmt, C.byref(out), C.byref(out_len), status)
    data = b""
    if rc == 0:
        data = C.string_at(out, out_len.value)
        lib.oracle_free(out)
    return rc, data


def _product_outputs(g):
    return {0: g.to_fasta(), 1: g.to_gfa(), 2: g.to_gfa(v1=True), 3: g.save()}


def _family_inserts(fasta, heuristic=1):
    g, inserts = POAGraph(), []
    for name, seq in synth.parse_fasta(fasta):
        if g.is_empty():
            pairs = None
        else:
            st, _, pairs, _ = H.oracle_align(g, seq, heuristic=heuristic)
            assert st == 0
        g.add_alignment_with_weights(name, seq, pairs)
        inserts.append((name, seq, pairs))
    return g, inserts


@pytest.mark.parametrize("seed,n,length,rate", [(1, 6, 60, 0.1), (2, 10, 150, 0.2), (3, 12, 400, 0.05), (4, 5, 30, 0.4)])
def test_same_insertions_same_bytes(seed, n, length, rate):
    g, inserts = _family_inserts(synth.family(seed, n, length, rate))
    for fmt, data in _product_outputs(g).items():
        assert _oracle_build(inserts, fmt) == (0, data), f"format {fmt}"


def test_rows_longer_than_80_columns_wrap_identically():
    g, inserts = _family_inserts(synth.family(9, 4, 200, 0.05))
    fasta = g.to_fasta()
    assert max(len(line) for line in fasta.split(b"\n")) == 80  # noodles' default line width
    assert _oracle_build(inserts, 0) == (0, fasta)


def test_reference_fixture_inputs():
    for name in ("small_test", "test_from_abpoa", "test2_from_abpoa"):
        g, inserts = _family_inserts(open(f"tests/golden/{name}.fa", "rb").read())
        for fmt, data in _product_outputs(g).items():
            assert _oracle_build(inserts, fmt) == (0, data), (name, fmt)


def test_save_load_and_continue():
    recs = synth.parse_fasta(synth.family(21, 9, 120, 0.12))
    text = lambda r: b"".join(b">%s\n%s\n" % (n.encode(), s) for n, s in r)
    g, first = _family_inserts(text(recs[:5]))
    blob = g.save()
    assert _oracle_build(first, 3) == (0, blob)
    # both sides load the same file and continue: slot reuse after the reload must agree too (GFA L-line order, .poasta bytes)
    g2 = POAGraph.load(blob)
    later = []
    for name, seq in recs[5:]:
        st, _, pairs, _ = H.oracle_align(g2, seq)
        g2.add_alignment_with_weights(name, seq, pairs)
        later.append((name, seq, pairs))
    for fmt, data in _product_outputs(g2).items():
        assert _oracle_build(later, fmt, graph_in=blob) == (0, data), f"format {fmt}"


def test_fasta_msa_import():
    g, _ = _family_inserts(synth.family(33, 7, 90, 0.15))
    msa = g.to_fasta()
    h = POAGraph.from_fasta_msa(msa)
    for fmt, data in _product_outputs(h).items():
        assert _oracle_build([], fmt, msa_in=msa) == (0, data), f"format {fmt}"


def test_error_cases_agree():
    g = POAGraph()
    g.add_alignment_with_weights("a", b"ACGT", None)
    # Some(empty alignment) for a non-empty sequence: InvalidAlignment on both sides (poa.rs:210-224)
    with pytest.raises(InvalidAlignment):
        g.add_alignment_with_weights("b", b"AC", [])
    assert _oracle_build([("a", b"ACGT", None), ("b", b"AC", [])], 0)[0] == -4
    # an empty record: a sequence that starts at `start`; the FASTA row is empty, the GFA writers fail (missing map key)
    inserts = [("a", b"ACGT", None), ("e", b"", [])]
    g.add_alignment_with_weights("e", b"", [])
    assert _oracle_build(inserts, 0) == (0, g.to_fasta())
    assert _oracle_build(inserts, 1)[0] == -8
    with pytest.raises(OSError):
        g.to_gfa()
