"""The persistent on-device MSA engine (poasta_b200/csrc/msa_engine.cuh) through the C ABI, bit-exact against the oracle-driven
build: whole-MSA bytes, per-sequence scores and the A* counters; the paths that only larger inputs reach are forced with
environment knobs in a subprocess (workers too small -> suspension and a second pass; 32-bit score cells)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from poasta_b200 import AffinePathAware, AlignmentType, POAGraph, aligner, align_batch, msa_batch, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expect(fa, fmt=0, **kw):
    rc, out, sc, _ = H.oracle_msa(fa, fmt=fmt, **kw)
    assert rc == 0
    return out, list(sc)


def test_ragged_families_and_counters():
    # families of 1..9 records of different lengths in one batch, incl. a single-record family, an empty record and an empty family
    fastas = [synth.family(1000 + f, 1 + f % 9, 60 + 37 * f, 0.08) for f in range(24)]
    fastas.append(b">a\nACGTACGT\n>b\n\n>c\nACGTTCGT\n")
    fastas.append(b"")
    graphs, scores, cells = msa_batch(fastas)
    q = v = p = c = 0
    for fa, g, sc in zip(fastas, graphs, scores):
        nseq = int(fa.startswith(b">")) + fa.count(b"\n>")
        rc, out, osc, ocells = H.oracle_msa(fa)
        if nseq == 0:
            assert g.is_empty()
            continue
        assert rc == 0 and list(sc) == list(osc) and g.to_fasta() == out
        assert g.save() == _expect(fa, fmt=3)[0]
        rc_gfa, gfa, _, _ = H.oracle_msa(fa, fmt=1)
        if rc_gfa == 0:
            assert g.to_gfa() == gfa
        else:  # a walk that starts at `start` (empty record): the reference's GFA writer panics on the missing map key
            with pytest.raises(OSError):
                g.to_gfa()
        c += ocells
        st = H.oracle_msa_stats(fa)
        q, v, p = q + st[0], v + st[1], p + st[2]
    assert cells == c
    assert aligner.last_batch_search_counts() == (q, v, p)  # AstarResult.num_queued / num_visited / num_pruned, summed


def test_existing_graphs_continue_on_the_device():
    # `poasta align -I graph` semantics: half of each family builds a graph, the other half is aligned into it in a second call
    fams = [synth.parse_fasta(synth.family(2000 + f, 8, 150, 0.07)) for f in range(6)]
    fa = lambda recs: b"".join(b">%s\n%s\n" % (n.encode(), s) for n, s in recs)
    first, second = [fa(r[:4]) for r in fams], [fa(r[4:]) for r in fams]
    graphs, _, _ = msa_batch(first)
    graphs, scores, _ = msa_batch(second, graphs=graphs)
    for r, g, sc in zip(fams, graphs, scores):
        out, osc = _expect(fa(r))
        assert g.to_fasta() == out and list(sc) == osc[4:]
        assert g.save() == _expect(fa(r), fmt=3)[0]


def test_all_heuristics_and_ends_free_through_the_engine():
    fastas = [synth.family(3000 + f, 6, 120, 0.1) for f in range(5)]
    for cfg, heur in ((aligner.AffineDijkstra(), 0), (AffinePathAware(), 2)):
        graphs, scores, _ = msa_batch(fastas, config=cfg)
        for fa, g, sc in zip(fastas, graphs, scores):
            out, osc = _expect(fa, heuristic=heur)
            assert g.to_fasta() == out and list(sc) == osc
    # ends-free stops at the first Match state past offset 0 (gap_affine.rs:202-207), so most builds die in the insert with
    # InvalidAlignment, in the reference as here (the call reports the first failing family)
    from poasta_b200 import InvalidAlignment
    for fa in fastas:
        rc, out, osc, _ = H.oracle_msa(fa, mode=1)
        if rc == 0:
            graphs, scores, _ = msa_batch([fa], aln_type=AlignmentType.EndsFree)
            assert graphs[0].to_fasta() == out and list(scores[0]) == list(osc)
        else:
            assert rc == -4
            with pytest.raises(InvalidAlignment):
                msa_batch([fa], aln_type=AlignmentType.EndsFree)


_SNIPPET = """
import sys, json
sys.path[:0] = ["tests", "."]
import helpers as H
from poasta_b200 import msa_batch, synth
fastas = [synth.family(4000 + f, 7, 300, 0.08) for f in range(40)]
graphs, scores, cells = msa_batch(fastas)
bad = 0
for fa, g, sc in zip(fastas, graphs, scores):
    rc, out, osc, _ = H.oracle_msa(fa)
    bad += not (rc == 0 and list(sc) == list(osc) and g.to_fasta() == out)
print(json.dumps({"bad": bad}))
"""


def _run_with_env(env):
    e = dict(os.environ)
    e.update(env)
    e["POASTA_TIMING"] = "1"
    r = subprocess.run([sys.executable, "-c", _SNIPPET], capture_output=True, text=True, cwd=ROOT, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


def test_families_that_outgrow_their_workers_are_resumed_by_a_second_pass():
    # workers sized for 330 nodes: the 300 bp families outgrow them after a few rounds, are suspended with their graphs intact and
    # finished by further passes with larger workers
    out, log = _run_with_env({"POASTA_WORKER_NODES": "330"})
    assert out["bad"] == 0
    assert "engine pass 1" in log and " 0 suspended" in log and "40 suspended" in log


def test_32_bit_score_cells():
    out, log = _run_with_env({"POASTA_CELL_BITS": "32"})
    assert out["bad"] == 0 and "32-bit cells" in log


def test_config2_scale_sample():
    # 1/10-scale slice of config 2 through the engine; every 50th family checked against the oracle (scores, FASTA, GFA)
    fastas = [synth.config2_family(f) for f in range(1000)]
    graphs, scores, cells = msa_batch(fastas)
    for f in range(0, 1000, 50):
        out, osc = _expect(fastas[f])
        assert list(scores[f]) == osc and graphs[f].to_fasta() == out
        assert graphs[f].to_gfa() == _expect(fastas[f], fmt=1)[0]
    # size-independent property over ALL families: every row of the MSA spells its input sequence
    for f in range(0, 1000, 7):
        recs = synth.parse_fasta(fastas[f])
        rows = synth.parse_fasta(graphs[f].to_fasta())
        assert [n for n, _ in rows] == [n for n, _ in recs]
        assert all(r.replace(b"-", b"") == s for (_, r), (_, s) in zip(rows, recs))


def test_config3_scale_sample():
    # 2,000 windows x 30 reads x 500 bp at 10 % error (1/100 of config 3, full window size); every 125th checked
    fastas = [synth.config3_window(w) for w in range(2000)]
    graphs, scores, cells = msa_batch(fastas)
    for w in range(0, 2000, 125):
        out, osc = _expect(fastas[w])
        assert list(scores[w]) == osc and graphs[w].to_fasta() == out


def test_config4_first_members_full_length():
    # the single 1.5 kb x 15 % family of config 4: its first 10 members (a sequential chain on ONE worker; ~1.3e7 visited states)
    fa = synth.config4_family(n_seqs=10)
    graphs, scores, _ = msa_batch([fa])
    out, osc = _expect(fa)
    assert list(scores[0]) == osc and graphs[0].to_fasta() == out
    assert graphs[0].node_count() > 2500


def test_config5_full_size_graph_reads_sample():
    # the 20 kb x 50 haplotype graph at full size, built with the product's DP-only mode (the oracle needs 13 minutes for it; any
    # graph of the haplotypes is a valid fixed graph); 200 reads of 2 kb ends-free -H path against it, scores of all and
    # alignments + A* counters of the first 20 checked against the oracle on the SAME graph
    from poasta_b200 import TIEBREAK_DP
    hap_fa, reads = synth.config5_reads(50, 20000, 0.01, 0, 200)
    g = msa_batch([hap_fa], tiebreak=TIEBREAK_DP)[0][0]
    assert g.node_count() > 20000 and g.num_sequences() == 50
    rows = synth.parse_fasta(g.to_fasta())
    assert all(r.replace(b"-", b"") == s for (_, r), (_, s) in zip(rows, synth.parse_fasta(hap_fa)))
    res = align_batch([g], reads, [0] * len(reads), AffinePathAware(), AlignmentType.EndsFree)
    from poasta_b200 import aligner
    assert aligner.last_search_table() == 2, "a 25,000 x 2,000 table per read must run as the blocked table"
    secs, cells, osc = H.oracle_align_many(g, reads, os.cpu_count() or 4, mode=1, heuristic=2, want_scores=True)
    assert [r.score for r in res] == list(osc)
    for q, r in list(zip(reads, res))[:20]:
        st, sc, pairs, stats = H.oracle_align(g, q, mode=1, heuristic=2)
        assert st == 0 and r.alignment == pairs and (r.num_queued, r.num_visited, r.num_pruned) == tuple(stats[:3])
