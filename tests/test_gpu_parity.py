"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the oracle.

What is compared (DESIGN.md "Parity"):
  * default mode (tiebreak = reference-exact): score AND alignment pairs bit-exact against the A* oracle, and
    whole-MSA FASTA / GFA bytes of progressive builds identical to the oracle-driven builds;
  * tiebreak = DP: score bit-exact against the A* oracle, pairs bit-exact against the dense model of the
    kernel's recurrences (oracle/dp_model.hpp).
"""
import numpy as np
import pytest

import helpers as H
from poasta_b200 import (TIEBREAK_DP, AffineDijkstra, AffineMinGapCost, AffinePathAware, AlignedPair, AlignmentType, GapAffine,
                         POAGraph, PoastaAligner, align_batch, msa_batch, synth)

pytestmark = pytest.mark.gpu


def _progressive_cases(fasta, max_align=None, heuristic=1):
    """Oracle-driven progressive build; yields (graph, query, oracle score, oracle pairs) before each insertion."""
    recs = synth.parse_fasta(fasta)
    g = POAGraph()
    n = 0
    for name, seq in recs:
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        st, sc, pairs, _ = H.oracle_align(g, seq, heuristic=heuristic)
        assert st == 0
        yield g, seq, sc, pairs
        n += 1
        if max_align and n >= max_align:
            return
        g.add_alignment_with_weights(name, seq, pairs)


def _check_both_modes(g, seq, sc, pairs):
    exact = PoastaAligner().align(g, seq)
    assert exact.score == sc, "score differs from the A* oracle"
    assert exact.alignment == pairs, "alignment differs from the A* oracle"
    dp = PoastaAligner(tiebreak=TIEBREAK_DP).align(g, seq)
    dsc, dpairs, _ = H.dpmodel_align(g, seq)
    assert dp.score == sc == dsc
    assert dp.alignment == dpairs, "DP-mode pairs differ from the dense model of the kernel"


def test_worked_example():
    g = POAGraph()
    g.add_alignment_with_weights("a", b"CAAAAG", None)
    for tb in (0, 1):
        res = PoastaAligner(tiebreak=tb).align(g, b"CAAAG")
        assert res.score == 8
        assert [(p.rpos, p.qpos) for p in res.alignment] == [(2, 0), (3, 1), (4, 2), (5, 3), (6, None), (7, 4)]


@pytest.mark.parametrize("fixture", ["small_test.fa", "test_from_abpoa.fa", "test2_from_abpoa.fa"])
def test_reference_fixtures(fixture):
    fasta = open(f"tests/golden/{fixture}", "rb").read()
    for g, seq, sc, pairs in _progressive_cases(fasta):
        _check_both_modes(g, seq, sc, pairs)
    # whole-MSA bytes: `poasta align -O fasta` / `-O gfa` through the device path vs the oracle-driven build
    graphs, scores, _ = msa_batch([fasta])
    for fmt, out in ((0, graphs[0].to_fasta()), (1, graphs[0].to_gfa()), (2, graphs[0].to_gfa(v1=True))):
        rc, expect, osc, _ = H.oracle_msa(fasta, fmt=fmt)
        assert rc == 0 and out == expect
        assert list(scores[0]) == list(osc)


def test_small_families_both_modes():
    for f in range(6):
        for g, seq, sc, pairs in _progressive_cases(synth.config2_family(f, n_seqs=8, root_len=200, rate=0.05)):
            _check_both_modes(g, seq, sc, pairs)


@pytest.mark.parametrize("config,heur", [(AffineDijkstra(), 0), (AffineMinGapCost(), 1), (AffinePathAware(), 2)])
def test_heuristics_change_ties_like_the_reference(config, heur):
    for f in range(3):
        for g, seq, sc, pairs in _progressive_cases(synth.config2_family(50 + f, n_seqs=6, root_len=150, rate=0.08), heuristic=heur):
            res = PoastaAligner(config).align(g, seq)
            assert res.score == sc and res.alignment == pairs


def test_non_default_costs():
    cfg = AffineMinGapCost(GapAffine(cost_mismatch=3, cost_gap_extend=1, cost_gap_open=5))
    for g, seq, _, _ in _progressive_cases(synth.config2_family(70, n_seqs=5, root_len=150, rate=0.08)):
        st, sc, pairs, _ = H.oracle_align(g, seq, x=3, o=5, e=1)
        res = PoastaAligner(cfg).align(g, seq)
        assert st == 0 and res.score == sc and res.alignment == pairs
        dp = PoastaAligner(cfg, tiebreak=TIEBREAK_DP).align(g, seq)
        assert dp.score == sc


def test_ends_free_matches_reference_behaviour():
    # the reference's ends-free terminates at the first Match state with offset > 0 (SURVEY.md finding 3); reproduce exactly
    g = POAGraph()
    g.add_alignment_with_weights("ref", b"ATCGATCGGCTA", None)
    g.add_alignment_with_weights("alt", b"ATCGTTCGGCTA", [AlignedPair(2 + i, i) for i in range(12)])
    for q in (b"TCG", b"ATCGAA", b"GGGG", b"A", b"CGATC"):
        for cfg, heur in ((AffineDijkstra(), 0), (AffineMinGapCost(), 1)):
            st, sc, pairs, _ = H.oracle_align(g, q, mode=1, heuristic=heur)
            res = PoastaAligner(cfg, AlignmentType.EndsFree).align(g, q)
            assert st == 0 and res.score == sc and res.alignment == pairs


def test_empty_query_and_empty_graph():
    g = POAGraph()
    assert PoastaAligner().align(g, b"ACGT").score == 16  # aligner/mod.rs:124-142
    g.add_alignment_with_weights("ref", b"ACGT", None)
    st, sc, pairs, _ = H.oracle_align(g, b"")
    res = PoastaAligner().align(g, b"")
    assert st == 0 and res.score == sc and res.alignment == pairs == []


def test_batch_of_independent_problems():
    graphs, queries, expect = [], [], []
    for f in range(4):
        for g, seq, sc, pairs in _progressive_cases(synth.config2_family(100 + f, n_seqs=4, root_len=150, rate=0.1), max_align=2):
            g2 = POAGraph.load(g.save())
            # reloading re-links adjacency lists in slot order (petgraph deserialisation), so ask the oracle again
            st, sc2, pairs2, _ = H.oracle_align(g2, seq)
            graphs.append(g2)
            queries.append(seq)
            expect.append((sc2, pairs2))
    res = align_batch(graphs, queries)
    for (sc, pairs), r in zip(expect, res):
        assert r.score == sc and r.alignment == pairs


def test_shared_graph_many_queries():
    fa = synth.config2_family(300, n_seqs=6, root_len=200, rate=0.03)
    graphs, _, _ = msa_batch([fa])
    g = graphs[0]
    queries = [s for _, s in synth.parse_fasta(synth.config2_family(300, n_seqs=12, root_len=200, rate=0.06))][6:]
    res = align_batch([g], queries, [0] * len(queries))
    for q, r in zip(queries, res):
        st, sc, pairs, _ = H.oracle_align(g, q)
        assert r.score == sc and r.alignment == pairs


def test_wide_in_degree_uses_32bit_trace():
    # 40 single-base branches between two anchors => in-degree 40 at the merge node (u32 trace words in DP mode)
    g = POAGraph()
    g.add_alignment_with_weights("s0", b"AAAACAAAA", None)
    for i in range(40):
        sym = bytes([ord("B") + i % 20, ord("a") + i // 20])
        seq = b"AAAA" + sym + b"AAAA"
        st, sc, pairs, _ = H.oracle_align(g, seq)
        assert st == 0
        g.add_alignment_with_weights(f"s{i + 1}", seq, pairs)
    q = b"AAAAZZAAAA"
    st, sc, pairs, _ = H.oracle_align(g, q)
    _check_both_modes(g, q, sc, pairs)


def test_msa_batch_bit_exact():
    fastas = [synth.config2_family(200 + f, n_seqs=6, root_len=180, rate=0.05) for f in range(5)]
    fastas.append(synth.config3_window(3, n_reads=8, template_len=120))
    graphs, scores, cells = msa_batch(fastas)
    assert cells > 0
    for fa, g, sc in zip(fastas, graphs, scores):
        rc, expect, osc, _ = H.oracle_msa(fa)
        assert rc == 0 and list(sc) == list(osc)
        assert g.to_fasta() == expect
        rc, expect_gfa, _, _ = H.oracle_msa(fa, fmt=1)
        assert g.to_gfa() == expect_gfa


def test_msa_batch_dp_mode_properties():
    fastas = [synth.config2_family(200 + f, n_seqs=6, root_len=180, rate=0.05) for f in range(3)]
    graphs, scores, _ = msa_batch(fastas, tiebreak=TIEBREAK_DP)
    for fa, g, sc in zip(fastas, graphs, scores):
        rc, _, osc, _ = H.oracle_msa(fa)
        assert sc[0] == 0 and sc[1] == osc[1]  # later graphs may differ through ties, the first alignment cannot
        seqs = dict(synth.parse_fasta(fa))
        for name, row in synth.parse_fasta(g.to_fasta()):
            assert row.replace(b"-", b"") == seqs[name]  # every MSA row spells its input sequence


def test_full_size_config2_family():
    """Config-2 sized family (20 x 1 kb): bit-exact MSA and scores against the oracle."""
    fa = synth.config2_family(7)
    graphs, scores, _ = msa_batch([fa])
    rc, expect, osc, _ = H.oracle_msa(fa)
    assert rc == 0 and list(scores[0]) == list(osc)
    assert graphs[0].to_fasta() == expect
    msa = synth.parse_fasta(graphs[0].to_fasta())
    seqs = dict(synth.parse_fasta(fa))
    assert len(msa) == 20 and len({len(r) for _, r in msa}) <= 2  # rows differ by at most the leading-gap quirk
    for name, row in msa:
        assert row.replace(b"-", b"") == seqs[name]


def test_query_longer_than_warp_kernel_limit():
    """Queries of 1024+ bases take the CTA-per-alignment DP kernel (the warp kernel covers 32 x 32 columns): both tiebreak
    modes must stay exact there (regression: the score pass once ran that kernel without a trace store)."""
    fa = synth.config2_family(41, n_seqs=4, root_len=1100, rate=0.04)
    graphs, scores, _ = msa_batch([fa])
    rc, expect, osc, _ = H.oracle_msa(fa)
    assert rc == 0 and list(scores[0]) == list(osc) and graphs[0].to_fasta() == expect
    for g, seq, sc, pairs in _progressive_cases(fa, max_align=2):
        _check_both_modes(g, seq, sc, pairs)


@pytest.mark.gpu
def test_int_peak_probe_and_device_time():
    """The integer issue micro-benchmark bench.py reports and the device-time accessor."""
    from poasta_b200 import aligner
    add_t, min_t, addmin_t = aligner.measure_int_peak(0)
    for v in (add_t, min_t, addmin_t):
        assert 1.0 < v < 200.0, (add_t, min_t, addmin_t)  # Tera lane-operations per second on any B200
    fams = [synth.config2_family(f, n_seqs=3, root_len=120, rate=0.05) for f in range(4)]
    msa_batch(fams, device=0)
    crit = aligner.last_batch_device_ms()
    kms, launches, cells = aligner.last_batch_stats()
    assert 0.0 < crit <= kms + aligner.last_batch_search_ms() + 1e-6


@pytest.mark.parametrize("name", ["small_test", "test_from_abpoa", "test2_from_abpoa"])
def test_msa_reproduces_committed_golden_outputs(name):
    """The reference's own test inputs through the CUDA path against the committed golden bytes (tools/make_golden.py)."""
    fasta = open(f"tests/golden/{name}.fa", "rb").read()
    graphs, scores, cells = msa_batch([fasta])
    assert graphs[0].to_fasta() == open(f"tests/golden/{name}.expected.fa", "rb").read()
    assert graphs[0].to_gfa() == open(f"tests/golden/{name}.expected.gfa", "rb").read()
    assert [int(s) for s in scores[0]] == [int(x) for x in open(f"tests/golden/{name}.expected.scores").read().split()]


def test_dp_mode_many_short_queries_are_chunked_not_overallocated():
    """the DP chunk budget must count what the chosen kernel allocates (the warp kernel's bit planes cost 1 KB per
    row whatever the query length). Many short queries against a 3k-node graph with a small budget: split into chunks, all right."""
    import subprocess, sys, os
    code = '''
import sys
sys.path[:0] = ["tests", "."]
import helpers as H
from poasta_b200 import POAGraph, TIEBREAK_DP, align_batch, synth
g = POAGraph()
for n, s in synth.parse_fasta(synth.family(11, 3, 3000, 0.02)):
    if g.is_empty():
        g.add_alignment_with_weights(n, s, None)
    else:
        g.add_alignment_with_weights(n, s, H.oracle_align(g, s)[2])
qs = [synth.parse_fasta(synth.family(100 + i, 1, 40, 0.0))[0][1] for i in range(600)] + [b""]
res = align_batch([g], qs, [0] * len(qs), tiebreak=TIEBREAK_DP)
# DP mode returns the optimum and the dense model's alignment (the reference's own search may sit above the optimum for a
# 40-base query against a 3 kb graph); the empty query goes through the search kernel and equals the reference
bad = 0
for q, r in list(zip(qs, res))[:600:25]:
    sc, pairs, _ = H.dpmodel_align(g, q)
    bad += not (r.score == sc and r.alignment == pairs)
st, sc, pairs, _ = H.oracle_align(g, b"")
bad += not (res[-1].score == sc and res[-1].alignment == pairs == [])
print("BAD", bad)
'''
    env = dict(os.environ, POASTA_DP_BUDGET_MB="64")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "BAD 0" in r.stdout, r.stdout + r.stderr[-1500:]
