"""BASELINE.json's other configurations as parity cases (reduced sizes the oracle finishes in seconds), through the
public batched entry point, bit-exact against the oracle-driven build: FASTA-MSA bytes and per-sequence scores."""
import pytest

import helpers as H
from poasta_b200 import AffinePathAware, AlignmentType, POAGraph, msa_batch, synth, align_batch

pytestmark = pytest.mark.gpu


def _check_msa(fastas, **kw):
    graphs, scores, cells = msa_batch(fastas, **kw)
    for fa, g, sc in zip(fastas, graphs, scores):
        rc, expect, osc, _ = H.oracle_msa(fa)
        assert rc == 0 and list(sc) == list(osc)
        assert g.to_fasta() == expect
    return graphs


def test_config3_consensus_windows():
    # 500 bp x 30 reads at 10 % error; here 6 windows x 12 reads
    _check_msa([synth.config3_window(w, n_reads=12) for w in range(6)])


def test_config4_large_divergent_family():
    # one family of 1.5 kb sequences at 15 % divergence (queries above the warp kernel's 1023-base limit); 8 of the 1000 members
    g = _check_msa([synth.config4_family(n_seqs=8)])[0]
    rc, gfa, _, _ = H.oracle_msa(synth.config4_family(n_seqs=8), fmt=1)
    assert g.to_gfa() == gfa


def test_config5_reads_against_fixed_graph():
    # library semantics (i): every read against the FIXED haplotype graph, ends-free + path heuristic, no insertion
    hap_fa, reads_fa = synth.config5(n_haps=6, root_len=1500, n_reads=12, read_len=300)
    g = _check_msa([hap_fa])[0]
    reads = [s for _, s in synth.parse_fasta(reads_fa)]
    for cfg, heur in ((AffinePathAware(), 2),):
        res = align_batch([g], reads, [0] * len(reads), cfg, AlignmentType.EndsFree)
        for q, r in zip(reads, res):
            st, sc, pairs, _ = H.oracle_align(g, q, mode=1, heuristic=heur)
            assert st == 0 and r.score == sc and r.alignment == pairs
    # and the same reads in global mode with the default heuristic (what a mapper would want)
    res = align_batch([g], reads[:4], [0] * 4)
    for q, r in zip(reads[:4], res):
        st, sc, pairs, _ = H.oracle_align(g, q)
        assert st == 0 and r.score == sc and r.alignment == pairs


def test_blocked_visited_table_matches_the_dense_one_and_the_oracle():
    """The search's visited table as 64-offset blocks handed out on first touch (picked for large V x W; forced here on a small
    graph): alignments and A* counters equal the oracle's in ends-free and global mode, with every heuristic; a block pool that is
    too small makes the alignment rerun with the dense table and still be right."""
    import os
    from poasta_b200 import AffineDijkstra, AffineMinGapCost, aligner
    hap_fa, reads_fa = synth.config5(n_haps=6, root_len=1500, n_reads=10, read_len=300)
    g = _check_msa([hap_fa])[0]
    reads = [s for _, s in synth.parse_fasta(reads_fa)]

    def check(expect_table):
        for cfg, heur in ((AffinePathAware(), 2), (AffineMinGapCost(), 1), (AffineDijkstra(), 0)):
            for aln_type, mode in ((AlignmentType.EndsFree, 1), (AlignmentType.Global, 0)):
                res = align_batch([g], reads, [0] * len(reads), cfg, aln_type)
                # (the blocked launch sizes its queue arena smaller too: a search that outgrows it is rerun with dense full-size
                # arenas, which is what the wide-open Dijkstra searches do - the layout is asserted where no rerun can happen)
                if heur == 2 and mode == 1:
                    assert aligner.last_search_table() == expect_table
                for q, r in zip(reads, res):
                    st, sc, pairs, stats = H.oracle_align(g, q, mode=mode, heuristic=heur)
                    assert st == 0 and r.score == sc and r.alignment == pairs
                    assert (r.num_queued, r.num_visited, r.num_pruned) == tuple(stats[:3])

    os.environ["POASTA_BLOCKED_TABLE"] = "1"
    try:
        check(2)
        os.environ["POASTA_BLOCKED_POOL_BLOCKS"] = "64"  # fewer blocks than nodes: every alignment overflows and is rerun dense
        try:
            check(1)
        finally:
            del os.environ["POASTA_BLOCKED_POOL_BLOCKS"]
    finally:
        del os.environ["POASTA_BLOCKED_TABLE"]
    os.environ["POASTA_BLOCKED_TABLE"] = "0"
    try:
        check(1)
    finally:
        del os.environ["POASTA_BLOCKED_TABLE"]
