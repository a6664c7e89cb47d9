"""Device-resident DP-only builds (tiebreak = DP; msa_dp_device.cuh): per round the graph is packed, filled, walked back and updated
on the device. Checked through the C ABI against a progressive build driven by the dense model of the recurrences
(oracle/dp_model.hpp): per-record scores and the bytes of the FASTA-MSA / GFA / .poasta outputs must be identical, for ragged
families, long queries (thread-block cluster kernel), existing graphs, forced sub-rounds (trace buffer smaller than a round), and
inputs the packed kernels do not cover (which must take the per-round host pipeline and still be right)."""
import os

import pytest

import helpers as H
from poasta_b200 import TIEBREAK_DP, AffineMinGapCost, GapAffine, POAGraph, aligner, msa_batch, synth

pytestmark = pytest.mark.gpu


def _expected(fa, x=4, o=6, e=2, g=None):
    """Progressive build with the dense model's alignments; returns (graph, per-record scores)."""
    g = g or POAGraph()
    scores = []
    for name, seq in synth.parse_fasta(fa):
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            scores.append(0)
            continue
        sc, pairs, _ = H.dpmodel_align(g, seq, x=x, o=o, e=e)
        scores.append(sc)
        g.add_alignment_with_weights(name, seq, pairs)
    return g, scores


def _check(fastas, x=4, o=6, e=2, path=2, kernel=None):
    cfg = AffineMinGapCost(GapAffine(cost_mismatch=x, cost_gap_extend=e, cost_gap_open=o))
    graphs, scores, cells = msa_batch(fastas, cfg, tiebreak=TIEBREAK_DP)
    assert aligner.last_msa_path() == path, f"pipeline {aligner.last_msa_path()}, expected {path}"
    if kernel is not None:
        assert aligner.last_dp_kernel() == kernel
    want_cells = 0
    for i, (fa, g, sc) in enumerate(zip(fastas, graphs, scores)):
        eg, esc = _expected(fa, x, o, e)
        assert list(sc) == esc, f"family {i}: scores {list(sc)} vs {esc}"
        assert g.to_fasta() == eg.to_fasta(), f"family {i}: FASTA-MSA differs"
        assert g.to_gfa() == eg.to_gfa(), f"family {i}: GFA differs"
        assert g.save() == eg.save(), f"family {i}: .poasta bytes differ"
    return graphs, scores, cells


def _ragged(seed0, n, lo=3, hi=9, length=(150, 420), rate=0.06):
    out = []
    for i in range(n):
        k = lo + (i * 7) % (hi - lo + 1)
        ln = length[0] + (i * 53) % (length[1] - length[0])
        out.append(synth.family(seed0 + i, k, ln, rate, prefix=f"r{i}_"))
    return out


def test_ragged_families_match_the_dense_model_build():
    _check(_ragged(100, 48), kernel=3)


def test_single_record_and_tiny_families():
    fas = [synth.family(7, 1, 50, 0.05), synth.family(8, 2, 1, 0.0), synth.family(9, 3, 2, 0.3), synth.family(10, 4, 33, 0.2)]
    _check(fas)


def test_full_size_config2_and_config3_units():
    _check([synth.config2_family(5), synth.config3_window(6), synth.config2_family(11, n_seqs=8)], kernel=3)


def test_long_queries_run_on_the_cluster_kernel():
    """Config-4 shape: 1.5 kb members at 15 % divergence; every round of the family crosses the 1,024-column segment boundary."""
    fas = [synth.config4_family(n_seqs=6, root_len=1500, rate=0.15), synth.family(3, 4, 2300, 0.05), synth.family(4, 3, 300, 0.05)]
    _check(fas, kernel=4)


def test_non_default_costs():
    _check(_ragged(300, 12), x=2, o=3, e=1)
    _check(_ragged(320, 12), x=5, o=0, e=3)


def test_sub_rounds_when_the_trace_buffer_is_smaller_than_a_round():
    fas = _ragged(500, 40, length=(300, 400))
    os.environ["POASTA_DP_TRACE_BYTES"] = str(3 << 20)  # a few alignments per sub-round
    try:
        _check(fas)
    finally:
        del os.environ["POASTA_DP_TRACE_BYTES"]


def test_pipeline_lanes_give_the_same_graphs():
    """Large batches are split into pipeline lanes (host thread + stream + buffers each); forced here on a small ragged batch."""
    fas = _ragged(600, 30)
    for lanes in ("2", "4"):
        os.environ["POASTA_DP_LANES"] = lanes
        try:
            _check(fas)
        finally:
            del os.environ["POASTA_DP_LANES"]


def test_existing_graphs_continue_on_the_device():
    fas = _ragged(700, 10, lo=6, hi=8)
    first, rest, expect = [], [], []
    for fa in fas:
        recs = fa.split(b">")[1:]
        first.append(b">" + b">".join(recs[:3]))
        rest.append(b">" + b">".join(recs[3:]))
        expect.append(_expected(fa)[0])
    graphs, _, _ = msa_batch(first, tiebreak=TIEBREAK_DP)
    graphs, scores, _ = msa_batch(rest, tiebreak=TIEBREAK_DP, graphs=graphs)
    assert aligner.last_msa_path() == 2
    for g, eg in zip(graphs, expect):
        assert g.to_fasta() == eg.to_fasta() and g.save() == eg.save()


def test_inputs_outside_the_packed_kernels_take_the_host_pipeline():
    fas = _ragged(900, 6)
    fas[2] = fas[2].replace(b"A", b"N", 3)  # a symbol outside ACGT
    _check(fas, path=3)
    _check(_ragged(950, 4), x=4, o=6, e=11, path=3)  # e * 1024 beyond the prefix-min offsets of the 16-bit kernel


def test_scale_sample_config2():
    """600 full-size config-2 families through the device rounds; every 60th family against the dense-model build."""
    fas = [synth.config2_family(20000 + f) for f in range(600)]
    graphs, scores, cells = msa_batch(fas, tiebreak=TIEBREAK_DP)
    assert aligner.last_msa_path() == 2
    for i in range(0, 600, 60):
        eg, esc = _expected(fas[i])
        assert list(scores[i]) == esc and graphs[i].to_fasta() == eg.to_fasta()
    for g in graphs:
        assert g.num_sequences() == 20
