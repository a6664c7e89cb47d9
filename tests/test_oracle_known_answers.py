"""Pins the oracle (oracle/astar_oracle.hpp) against every exact-value test the reference holds
for the alignment path. Each test names the reference test it replays."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from helpers import DELETION, INSERTION, MATCH
from poasta_b200 import POAGraph

u32p = C.POINTER(C.c_uint32)


def test_rev_postorder_small_branching_graph():
    # graphs/tools.rs:54-69: expects [n1, n2, n3, n4]
    v = H.mock_view(4, [(0, 1), (0, 2), (1, 3), (2, 3)])
    out = np.zeros(4, np.uint32)
    n = H.oracle().oracle_rev_postorder(C.byref(v.c), 1, out.ctypes.data_as(u32p))
    assert n == 4 and list(out) == [0, 1, 2, 3]


def test_rev_postorder_single_node():
    # graphs/tools.rs:42-51
    v = H.mock_view(1, [])
    out = np.zeros(1, np.uint32)
    assert H.oracle().oracle_rev_postorder(C.byref(v.c), 1, out.ctypes.data_as(u32p)) == 1 and out[0] == 0


def _bubbles(v):
    out = np.zeros(2 * 64, np.uint32)
    n = H.oracle().oracle_superbubbles(C.byref(v.c), 1, out.ctypes.data_as(u32p), 64)
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]


def test_superbubble_finder():
    # bubbles/finder.rs:189-218 (values are node index + 1)
    g1 = H.test_graph1()
    end = 9
    b1 = {(s + 1, t + 1) for s, t in _bubbles(g1) if s != end and t != end}
    assert b1 == {(1, 2), (2, 3), (4, 5), (5, 6), (7, 8), (8, 9)}
    g2 = H.test_graph2()
    b2 = {(s + 1, t + 1) for s, t in _bubbles(g2)}
    assert b2 == {(8, 15), (11, 12), (5, 7), (3, 8), (1, 3)}


def _bubble_index(v):
    n = v.n
    counts = np.zeros(n, np.uint32)
    maps = np.zeros((n, 8, 3), np.uint32)
    dist = np.zeros((n, 2), np.uint32)
    H.oracle().oracle_bubble_index(C.byref(v.c), 1, 8, counts.ctypes.data_as(u32p), maps.ctypes.data_as(u32p),
                                   dist.ctypes.data_as(u32p))
    return counts, maps, dist


def test_bubble_map_builder():
    # bubbles/index.rs:232-289
    g1 = H.test_graph1()
    counts, maps, _ = _bubble_index(g1)
    truth1 = [[(1, 1)], [(2, 1), (1, 0)], [(2, 0)], [(4, 1)], [(5, 1), (4, 0)], [(5, 0)], [(7, 1)], [(8, 1), (7, 0)],
              [(8, 0)]]
    for n in range(9):
        got = [(int(maps[n, k, 0]), int(maps[n, k, 1])) for k in range(counts[n]) if maps[n, k, 0] != 9]
        assert got == truth1[n], n
    g2 = H.test_graph2()
    counts, maps, _ = _bubble_index(g2)
    truth2 = [[(2, 1)], [(2, 1)], [(7, 2), (2, 0)], [(7, 1)], [(6, 2), (7, 3)], [(7, 2), (6, 1)], [(7, 1), (6, 0)],
              [(14, 1), (7, 0)], [(7, 3), (6, 2)], [(7, 2), (6, 1)], [(11, 1), (7, 2)], [(7, 1), (11, 0)], [(14, 1)],
              [(14, 1)], [(14, 0)]]
    for n in range(15):
        got = [(int(maps[n, k, 0]), int(maps[n, k, 1])) for k in range(counts[n])]
        assert got == truth2[n], n


def test_dist_to_exit():
    # bubbles/index.rs:292-317
    _, _, dist = _bubble_index(H.test_graph2())
    assert [tuple(map(int, d)) for d in dist] == [(4, 10), (4, 9), (3, 8), (2, 4), (4, 7), (3, 6), (2, 4), (1, 3), (4, 6),
                                                  (3, 5), (3, 5), (2, 4), (1, 2), (1, 1), (0, 0)]


def test_min_gap_cost_heuristic():
    # aligner/heuristic.rs:224-247: GapAffine::new(4, 2, 6) = mismatch 4, extend 2, open 6; seq len 10; node index 1
    g = H.test_graph1()
    h = lambda off, st: H.oracle().oracle_heuristic(C.byref(g.c), 1, 1, 4, 6, 2, 10, 1, off, st, 10, 5)
    assert (h(2, MATCH), h(2, DELETION), h(2, INSERTION)) == (14, 14, 8)
    assert (h(7, MATCH), h(7, DELETION), h(7, INSERTION)) == (8, 2, 8)
    assert (h(6, MATCH), h(6, DELETION), h(6, INSERTION)) == (0, 0, 0)


def test_dijkstra_heuristic_is_zero():
    # aligner/heuristic.rs:210-221
    g = H.test_graph1()
    for st in (MATCH, DELETION, INSERTION):
        assert H.oracle().oracle_heuristic(C.byref(g.c), 1, 0, 4, 6, 2, 10, 1, 5, st, 10, 5) == 0


def test_path_aware_heuristic_properties():
    # aligner/heuristic.rs:249-292 (inequalities only in the reference)
    g = H.test_graph1()
    h0 = H.oracle().oracle_heuristic(C.byref(g.c), 1, 2, 4, 6, 2, 10, 0, 0, MATCH, 5, 3)
    h5 = H.oracle().oracle_heuristic(C.byref(g.c), 1, 2, 4, 6, 2, 10, 0, 5, MATCH, 5, 3)
    assert h5 <= h0
    hl = H.oracle().oracle_heuristic(C.byref(g.c), 1, 2, 4, 6, 2, 10, 9, 5, MATCH, 1, 3)
    assert hl <= H.oracle().oracle_gap_cost(4, 6, 2, MATCH, 5)


def test_gap_cost():
    # aligner/scoring/gap_affine.rs:1033-1047: GapAffine::new(2, 1, 5)
    gc = lambda st, n: H.oracle().oracle_gap_cost(2, 5, 1, st, n)
    assert gc(MATCH, 0) == 0 and gc(MATCH, 1) == 6 and gc(MATCH, 3) == 8
    assert gc(INSERTION, 1) == 1 and gc(DELETION, 2) == 2


def test_layered_queue():
    # aligner/queue.rs:110-136
    assert H.oracle().oracle_queue_selftest() == 0


def _dfa(seq, node, offset, prune):
    g = H.test_graph1()
    syms = bytes([ord(c) for c in "ACGTACGTAN"])
    out = np.zeros(7, np.uint32)
    H.oracle().oracle_dfa_event(C.byref(g.c), syms, seq, len(seq), node, offset, prune, out.ctypes.data_as(u32p))
    return [int(x) for x in out]


def test_dfa_events():
    # aligner/dfa.rs:420-482. kinds: 0 None, 1 RefGraphEnd, 2 QueryEnd, 3 Mismatch
    kind, pn, po, cn, co, nv, npr = _dfa(b"AA", 0, 0, 0)
    assert (kind, pn, po, cn, co, nv, npr) == (3, 0, 1, 1, 2, 1, 0)       # returns_mismatch_variant
    kind, pn, po, cn, co, nv, npr = _dfa(b"AC", 0, 0, 0)
    assert kind == 2 and po == 2 and cn == 2 and nv == 2                    # returns_query_end_variant (parent = node 1)
    kind, *_rest, nv, npr = _dfa(b"A", 5, 0, 0)
    assert kind == 1 and nv == 0                                            # returns_ref_graph_end_variant
    kind, *_rest, nv, npr = _dfa(b"AC", 0, 0, 1)
    assert kind == 0 and npr == 1                                           # counts_pruned_states


def _linear_graph(seq):
    g = POAGraph()
    g.add_alignment_with_weights("ref", seq, None)
    return g


def test_ends_free_exact_scores():
    # gap_affine.rs:1125-1178 (prefix/suffix skipping), :1320-1353 (single nucleotide), Dijkstra, costs (1, 2, 8)
    for ref, qry in ((b"ATCG", b"TCG"), (b"ATCG", b"ATCGAA"), (b"ATCGATCG", b"A")):
        st, score, pairs, _ = H.oracle_align(_linear_graph(ref), qry, x=1, o=8, e=2, mode=1, heuristic=0)
        assert st == 0 and score == 0, (ref, qry, score)


def test_ends_free_basic_alignment_bound():
    # gap_affine.rs:1050-1080: score <= 5
    st, score, _, _ = H.oracle_align(_linear_graph(b"ATCGATCG"), b"CGATC", x=1, o=8, e=2, mode=1, heuristic=0)
    assert st == 0 and score <= 5


def test_ends_free_vs_global():
    # gap_affine.rs:1083-1122
    g = _linear_graph(b"AAATTTGGGCCC")
    _, sg, _, _ = H.oracle_align(g, b"TTTGGG", x=1, o=8, e=2, mode=0, heuristic=0)
    _, se, _, _ = H.oracle_align(g, b"TTTGGG", x=1, o=8, e=2, mode=1, heuristic=0)
    assert se <= sg
    assert sg == 26  # 6 mismatches (x=1) + one 6-base deletion (8 + 6*2): cheaper than two 3-base deletions


def test_single_character_sequences():
    # tests/edge_cases.rs:240-280: match => 0; mismatch => 1, or an insertion with score 10
    g = _linear_graph(b"A")
    for nuc in b"ATCG":
        st, score, pairs, _ = H.oracle_align(g, bytes([nuc]), x=1, o=8, e=2, mode=1, heuristic=0)
        assert st == 0
        if nuc == ord("A"):
            assert score == 0
        else:
            assert score == 1 or (score == 10 and len(pairs) == 1 and pairs[0].rpos is None and pairs[0].qpos == 0)


def test_empty_graph_shortcut():
    # aligner/mod.rs:124-142
    g = POAGraph()
    st, score, pairs, _ = H.oracle_align(g, b"ACGT")
    assert (st, score, pairs) == (0, 16, [])
    st, score, pairs, _ = H.oracle_align(g, b"")
    assert (st, score, pairs) == (0, 0, [])


def _three_heuristics(graph, query):
    return [H.oracle_align(graph, query, heuristic=h)[1] for h in (0, 1, 2)]


def test_heuristic_consistency_scores():
    # tests/test_heuristics.rs:7-56, 59-109, 111-148: equal global scores for dijkstra / mingap / path
    g = POAGraph()
    for name, s in (("reference", b"ACGTACGTACGTACGTACGT"), ("variant1", b"ACGTACGTAAGTACGTACGT"),
                    ("variant2", b"ACGTACGTACGTACGT")):
        g.add_alignment_with_weights(name, s, None)
    a, b, c = _three_heuristics(g, b"ACGTACGTAAGTACGTACGT")
    assert a == b == c == 0

    g = POAGraph()
    ref = b"ACGT" * 16
    for name, s in (("reference", ref), ("deletion", b"ACGT" * 13),
                    ("insertion", b"ACGTACGTACGTACGTAAAACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"),
                    ("complex", b"ACGTACGTAAGTACCTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT")):
        g.add_alignment_with_weights(name, s, None)
    q = b"ACGTACGTAAGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
    _, b, c = _three_heuristics(g, q)
    assert b == c

    g = POAGraph()
    g.add_alignment_with_weights("reference", b"ACGTACGTACGTACGT", None)
    g.add_alignment_with_weights("variant", b"ACGTAAGTACGTACGT", None)
    a, b, c = _three_heuristics(g, b"ACGTACGTAAGTACGT")
    assert a == b == c


def test_survey_worked_example():
    # SURVEY.md A.16 (hand-derived from the reference code): CAAAAG vs CAAAG => score 8, deletion on the LAST A
    g = _linear_graph(b"CAAAAG")
    st, score, pairs, _ = H.oracle_align(g, b"CAAAG")
    assert st == 0 and score == 8
    assert [(p.rpos, p.qpos) for p in pairs] == [(2, 0), (3, 1), (4, 2), (5, 3), (6, None), (7, 4)]
    rc, fasta, scores, _ = H.oracle_msa(b">a\nCAAAAG\n>b\nCAAAG\n")
    assert rc == 0 and fasta == b">a\nCAAAAG\n>b\nCAAA-G\n" and list(scores) == [0, 8]


@pytest.mark.parametrize("name", ["small_test", "test_from_abpoa", "test2_from_abpoa"])
def test_oracle_reproduces_committed_golden_outputs(name):
    """tests/golden/*.expected.* were written by tools/make_golden.py from THIS oracle (not by the real reference binary,
    which cannot be built here): they freeze the restated behaviour, so an unintended change of the oracle, the graph
    model or the writers shows up as a byte difference."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fasta = open(os.path.join(gold, name + ".fa"), "rb").read()
    for fmt, ext in ((0, "fa"), (1, "gfa")):
        rc, data, scores, _ = H.oracle_msa(fasta, fmt=fmt)
        assert rc == 0
        assert data == open(os.path.join(gold, f"{name}.expected.{ext}"), "rb").read()
    want = [int(x) for x in open(os.path.join(gold, name + ".expected.scores")).read().split()]
    assert [int(s) for s in scores] == want
