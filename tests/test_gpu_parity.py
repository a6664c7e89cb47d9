"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the oracle.

Layers of the check (DESIGN.md "Parity"):
  * score: bit-exact against the A* oracle (the reference's optimal score) on every case;
  * alignment pairs: bit-exact against the dense DP model of the kernel's recurrences (oracle/dp_model.hpp);
  * alignment pairs vs the A* oracle: measured and reported (ties decided by A*'s pop order).
"""
import numpy as np
import pytest

import helpers as H
from poasta_b200 import AlignedPair, POAGraph, PoastaAligner, align_batch, msa_batch, synth

pytestmark = pytest.mark.gpu


def _progressive_cases(fasta, max_align=None):
    """Oracle-driven progressive build; yields (graph, query) before each insertion."""
    recs = synth.parse_fasta(fasta)
    g = POAGraph()
    n = 0
    for name, seq in recs:
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        st, sc, pairs, _ = H.oracle_align(g, seq)
        assert st == 0
        yield g, seq, sc, pairs
        n += 1
        if max_align and n >= max_align:
            return
        g.add_alignment_with_weights(name, seq, pairs)


def _check(g, seq, sc, pairs, res):
    assert res.score == sc, "score differs from the A* oracle"
    dsc, dpairs, _ = H.dpmodel_align(g, seq)
    assert res.score == dsc
    assert res.alignment == dpairs, "pairs differ from the dense model of the kernel"
    return res.alignment == pairs


def test_worked_example():
    g = POAGraph()
    g.add_alignment_with_weights("a", b"CAAAAG", None)
    res = PoastaAligner().align(g, b"CAAAG")
    assert res.score == 8
    assert [(p.rpos, p.qpos) for p in res.alignment] == [(2, 0), (3, 1), (4, 2), (5, 3), (6, None), (7, 4)]


@pytest.mark.parametrize("fixture", ["small_test.fa", "test_from_abpoa.fa", "test2_from_abpoa.fa"])
def test_reference_fixtures(fixture):
    fasta = open(f"tests/golden/{fixture}", "rb").read()
    for g, seq, sc, pairs in _progressive_cases(fasta):
        _check(g, seq, sc, pairs, PoastaAligner().align(g, seq))


def test_small_families_score_and_model_parity():
    same = total = 0
    for f in range(6):
        for g, seq, sc, pairs in _progressive_cases(synth.config2_family(f, n_seqs=8, root_len=200, rate=0.05)):
            same += _check(g, seq, sc, pairs, PoastaAligner().align(g, seq))
            total += 1
    print(f"pairs identical to the A* oracle: {same}/{total}")


def test_batch_of_independent_problems():
    graphs, queries, expect = [], [], []
    for f in range(4):
        for g, seq, sc, pairs in _progressive_cases(synth.config2_family(100 + f, n_seqs=4, root_len=150, rate=0.1), max_align=2):
            g2 = POAGraph.load(g.save())
            graphs.append(g2)
            queries.append(seq)
            expect.append((g2, seq, sc, pairs))
    res = align_batch(graphs, queries)
    for (g, seq, sc, pairs), r in zip(expect, res):
        _check(g, seq, sc, pairs, r)


def test_wide_in_degree_uses_32bit_trace():
    # 40 single-base branches between two anchors => in-degree 40 at the merge node
    g = POAGraph()
    g.add_alignment_with_weights("s0", b"AAAACAAAA", None)
    for i in range(40):
        sym = bytes([ord("B") + i % 20, ord("a") + i // 20])
        seq = b"AAAA" + sym + b"AAAA"
        res = PoastaAligner().align(g, seq)
        st, sc, pairs, _ = H.oracle_align(g, seq)
        assert st == 0 and res.score == sc
        g.add_alignment_with_weights(f"s{i + 1}", seq, pairs)
    q = b"AAAAZZAAAA"
    res = PoastaAligner().align(g, q)
    st, sc, pairs, _ = H.oracle_align(g, q)
    dsc, dpairs, _ = H.dpmodel_align(g, q)
    assert res.score == sc == dsc and res.alignment == dpairs


def test_msa_batch_scores_match_oracle():
    fastas = [synth.config2_family(200 + f, n_seqs=6, root_len=180, rate=0.05) for f in range(5)]
    graphs, scores, cells = msa_batch(fastas)
    assert cells > 0
    for fa, g, sc in zip(fastas, graphs, scores):
        # progressive MSA: the first alignment's score is comparable even when later graphs diverge through ties
        rc, _, osc, _ = H.oracle_msa(fa)
        assert rc == 0 and sc[0] == 0 and sc[1] == osc[1]
        assert g.num_sequences() == 6
        seqs = dict(synth.parse_fasta(fa))
        for name, row in synth.parse_fasta(g.to_fasta()):
            assert row.replace(b"-", b"") == seqs[name]  # every MSA row spells its input sequence


def test_full_size_config2_family_properties():
    """Config-2 sized family (20 x 1 kb): scores vs oracle for the first alignments, round trip for all."""
    fa = synth.config2_family(7)
    graphs, scores, _ = msa_batch([fa])
    rc, _, osc, _ = H.oracle_msa(fa)
    assert scores[0][1] == osc[1]
    seqs = dict(synth.parse_fasta(fa))
    msa = synth.parse_fasta(graphs[0].to_fasta())
    assert len(msa) == 20
    for name, row in msa:
        assert row.replace(b"-", b"") == seqs[name]
