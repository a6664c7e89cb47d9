"""The product's order-exact search core (poasta_b200/csrc/replay_core.hpp, the code the device runs) compiled
for the host through a test hook in the oracle library, against the A* oracle: identical score, pairs and
queued/visited/pruned counters. This pins the device logic without a GPU; the -m gpu tests pin the kernel."""
import pytest

import helpers as H
from poasta_b200 import AlignedPair, POAGraph, synth


def _cases(fasta, heuristic):
    g = POAGraph()
    for name, seq in synth.parse_fasta(fasta):
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        st, sc, pairs, stats = H.oracle_align(g, seq, heuristic=heuristic)
        yield g, seq, st, sc, pairs, stats
        g.add_alignment_with_weights(name, seq, pairs)


@pytest.mark.parametrize("heuristic", [0, 1, 2])
def test_search_core_equals_oracle_on_families(heuristic):
    n = 0
    for f in range(3):
        for g, seq, st, sc, pairs, stats in _cases(synth.config2_family(f, n_seqs=8, root_len=250, rate=0.06), heuristic):
            for cell16 in (False, True):  # u32 and u16 score cells
                st2, sc2, pairs2, stats2 = H.replay_host_align(g, seq, heuristic=heuristic, cell16=cell16)
                assert (st2, sc2, pairs2, stats2[:3]) == (st, sc, pairs, stats[:3])
            n += 1
    assert n == 21


@pytest.mark.parametrize("fixture", ["small_test.fa", "test_from_abpoa.fa", "test2_from_abpoa.fa"])
def test_search_core_on_reference_fixtures(fixture):
    fasta = open(f"tests/golden/{fixture}", "rb").read()
    for g, seq, st, sc, pairs, stats in _cases(fasta, 1):
        assert H.replay_host_align(g, seq)[:3] == (st, sc, pairs)


def test_search_core_ends_free_and_edge_cases():
    g = POAGraph()
    g.add_alignment_with_weights("ref", b"ATCGATCGGCTA", None)
    g.add_alignment_with_weights("alt", b"ATCGTTCGGCTA", [AlignedPair(2 + i, i) for i in range(12)])
    for q in (b"TCG", b"ATCGAA", b"GGGG", b"A", b"CGATC", b""):
        for mode in (0, 1):
            for heur in (0, 1, 2):
                a = H.oracle_align(g, q, mode=mode, heuristic=heur)
                b = H.replay_host_align(g, q, mode=mode, heuristic=heur)
                assert (b[0], b[1], b[2], b[3][:3]) == (a[0], a[1], a[2], a[3][:3]), (q, mode, heur)


def test_search_core_high_divergence():
    for g, seq, st, sc, pairs, stats in _cases(synth.family(99, 6, 200, 0.2), 1):
        assert H.replay_host_align(g, seq)[:3] == (st, sc, pairs)


def _counters():
    import ctypes as C
    buf = (C.c_uint64 * 64)()
    n = H.oracle().replay_host_counters(buf, 64)
    return list(buf)[:n]


@pytest.mark.parametrize("epoch", [2, 0x7FFF, 0xFFFE])
def test_stale_records_read_as_unvisited(epoch, monkeypatch):
    """The visited table is never reset between alignments: records carry the epoch of the alignment that wrote them and any
    other epoch reads as unvisited (replay_core.hpp Rec). The hook pre-fills the table with garbage records of OTHER epochs;
    results and counters must stay identical to the oracle, with 16-bit and 32-bit cells."""
    monkeypatch.setenv("RP_HOST_EPOCH", str(epoch))
    n = 0
    for g, seq, st, sc, pairs, stats in _cases(synth.config2_family(1, n_seqs=6, root_len=250, rate=0.06), 1):
        for cell16 in (True, False):
            st2, sc2, pairs2, stats2 = H.replay_host_align(g, seq, cell16=cell16)
            assert (st2, sc2, pairs2, stats2[:3]) == (st, sc, pairs, stats[:3])
        n += 1
    assert n == 5
