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
