"""dp_fill_packed_kernel (two 16-bit scores per register, DPX packed instructions; dp_packed_kernel.cuh) + traceback_rows_kernel,
through the C ABI (tiebreak = DP and the score cross-check of the exact mode): score bit-exact against the A* oracle, pairs
bit-exact against the dense model of the recurrences (oracle/dp_model.hpp), on progressive graphs with merge rows, at the lane /
pair / chunk boundaries of the query axis, at full config-2 / config-3 size, under costs that keep or leave the packed kernel,
and for graphs outside ACGT (which must fall back to the 32-bit kernel and still be right)."""
import os
import subprocess
import sys

import pytest

import helpers as H
from poasta_b200 import TIEBREAK_DP, AffineMinGapCost, GapAffine, POAGraph, PoastaAligner, align_batch, synth

pytestmark = pytest.mark.gpu


def _graph(fasta, upto=None):
    g = POAGraph()
    for i, (name, seq) in enumerate(synth.parse_fasta(fasta)):
        if upto is not None and i >= upto:
            break
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
        else:
            g.add_alignment_with_weights(name, seq, H.oracle_align(g, seq)[2])
    return g


def _check(g, queries, x=4, o=6, e=2):
    cfg = AffineMinGapCost(GapAffine(cost_mismatch=x, cost_gap_extend=e, cost_gap_open=o))
    res = align_batch([g], queries, [0] * len(queries), cfg, tiebreak=TIEBREAK_DP)
    for q, r in zip(queries, res):
        sc, pairs, _ = H.dpmodel_align(g, q, x=x, o=o, e=e)
        assert r.score == sc, (len(q), r.score, sc)
        assert r.alignment == pairs, f"pairs differ from the dense model (Q = {len(q)})"


def test_query_axis_boundaries():
    """Query lengths around the pair (2), chunk (8), lane (32) and kernel (1023) boundaries against a graph with merge rows."""
    fa = synth.config2_family(3, n_seqs=6, root_len=1000, rate=0.05)
    g = _graph(fa)
    root = synth.parse_fasta(fa)[0][1]
    lens = [1, 2, 3, 7, 8, 9, 15, 16, 17, 30, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 767, 991, 992, 993, 1000, 1015, 1016, 1022, 1023]
    _check(g, [root[:n] for n in lens] + [root[100:100 + n] for n in (5, 40, 333)])


def test_full_size_config2_family_every_round():
    fa = synth.config2_family(9)
    recs = synth.parse_fasta(fa)
    g = POAGraph()
    for i, (name, seq) in enumerate(recs):
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        st, sc, pairs, _ = H.oracle_align(g, seq)
        dp = PoastaAligner(tiebreak=TIEBREAK_DP).align(g, seq)
        dsc, dpairs, _ = H.dpmodel_align(g, seq)
        assert st == 0 and dp.score == sc == dsc, (i, dp.score, sc, dsc)
        assert dp.alignment == dpairs, f"round {i}"
        g.add_alignment_with_weights(name, seq, pairs)


def test_config3_window_batch():
    """Racon-style window (30 reads x 500 bp, 10 % error): all reads against the graph of the first 12 in ONE batch."""
    fa = synth.config3_window(5)
    g = _graph(fa, upto=12)
    _check(g, [s for _, s in synth.parse_fasta(fa)])


@pytest.mark.parametrize("x,o,e", [(3, 5, 1), (1, 0, 1), (7, 0, 3), (2, 30, 10), (5, 8, 11), (255, 255, 2)])
def test_costs_inside_and_outside_the_packed_range(x, o, e):
    """e <= 10 keeps the packed kernel (prefix-min offsets fit 16 bits), e = 11 and saturating costs leave it; o = 0 makes
    open and extend tie everywhere."""
    fa = synth.config2_family(21, n_seqs=5, root_len=300, rate=0.08)
    g = _graph(fa)
    _check(g, [s for _, s in synth.parse_fasta(fa)] + [synth.parse_fasta(fa)[0][1][:77]], x=x, o=o, e=e)


def test_symbols_outside_acgt_fall_back():
    g = POAGraph()
    g.add_alignment_with_weights("a", b"ACGTNNACGTRYACGTACGTTTGACA", None)
    st, sc, pairs, _ = H.oracle_align(g, b"ACGTNACGTYACGAACGTTGACA")
    g.add_alignment_with_weights("b", b"ACGTNACGTYACGAACGTTGACA", pairs)
    _check(g, [b"ACGTNNACGTACGTACGTTTGACA", b"NNNN", b"ACGTACGT", b"acgtacgtnnnn"])


def test_packed_equals_the_32_bit_kernel():
    """The same batch through the packed kernel and (POASTA_DP_KERNEL=warp32, a separate process) the 32-bit warp kernel."""
    code = '''
import sys
sys.path[:0] = ["tests", "."]
import helpers as H
from poasta_b200 import POAGraph, TIEBREAK_DP, align_batch, synth
fa = synth.config2_family(33, n_seqs=10, root_len=700, rate=0.06)
g = POAGraph()
for n, s in synth.parse_fasta(fa):
    if g.is_empty():
        g.add_alignment_with_weights(n, s, None)
    else:
        g.add_alignment_with_weights(n, s, H.oracle_align(g, s)[2])
qs = [s for _, s in synth.parse_fasta(synth.config2_family(34, n_seqs=10, root_len=700, rate=0.06))] + [s for _, s in synth.parse_fasta(fa)]
for r in align_batch([g], qs, [0] * len(qs), tiebreak=TIEBREAK_DP):
    print(r.score, [(p.rpos, p.qpos) for p in r.alignment])
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for kernel in ("", "warp32"):
        env = dict(os.environ, POASTA_DP_KERNEL=kernel)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and outs[0].count("\n") == 20


# ---- cluster mode: queries of 1024+ bases, one 1024-column segment per CTA of a thread-block cluster, carries through DSMEM ----
def _check_cluster(g, queries, x=4, o=6, e=2, kernel=4):
    from poasta_b200 import aligner
    cfg = AffineMinGapCost(GapAffine(cost_mismatch=x, cost_gap_extend=e, cost_gap_open=o))
    res = align_batch([g], queries, [0] * len(queries), cfg, tiebreak=TIEBREAK_DP)
    assert aligner.last_dp_kernel() == kernel, f"the batch ran on DP kernel {aligner.last_dp_kernel()}, expected {kernel}"
    for q, r in zip(queries, res):
        sc, pairs, _ = H.dpmodel_align(g, q, x=x, o=o, e=e)
        assert r.score == sc, (len(q), r.score, sc)
        assert r.alignment == pairs, f"pairs differ from the dense model (Q = {len(q)})"


def test_cluster_segment_boundaries():
    """Query lengths around the 1024-column segment boundaries, 2 and 3 active segments in one launch (cluster of 4)."""
    fa = synth.config2_family(61, n_seqs=4, root_len=2500, rate=0.05)
    g = _graph(fa)
    recs = synth.parse_fasta(fa)
    root = recs[0][1]
    _check_cluster(g, [root[:n] for n in (1024, 1025, 1055, 1056, 1500, 2046, 2047, 2048, 2049, 2400)] + [recs[3][1], recs[1][1][300:1900]])


def test_cluster_of_eight():
    """7 kb queries: 7 active segments in a cluster of 8; a short query in the same batch leaves six CTAs idle. Costs 2/3/1:
    the packed kernels need the provable bound 2o + e (Lmin + Q) + x below 0x5000, which 14 k bases allow only at e = 1."""
    fa = synth.config2_family(62, n_seqs=3, root_len=7000, rate=0.04)
    g = _graph(fa)
    recs = synth.parse_fasta(fa)
    _check_cluster(g, [recs[2][1], recs[0][1][:4100], recs[1][1][:1030]], x=2, o=3, e=1)


def test_long_query_outside_the_16_bit_bound_falls_back():
    """The same 7 kb problem at the default costs (bound 28 k > 0x5000) must leave the packed kernels and still be right."""
    fa = synth.config2_family(62, n_seqs=3, root_len=7000, rate=0.04)
    g = _graph(fa)
    recs = synth.parse_fasta(fa)
    _check_cluster(g, [recs[2][1]], kernel=1)


def test_cluster_config4_shape():
    """Config-4 shape (16S-like 1.5 kb at 15 % divergence): every member against the growing graph, DP mode."""
    from poasta_b200 import aligner
    recs = synth.parse_fasta(synth.config4_family(n_seqs=6, root_len=1500, rate=0.15))
    g = POAGraph()
    for i, (name, seq) in enumerate(recs):
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        dp = PoastaAligner(tiebreak=TIEBREAK_DP).align(g, seq)
        dsc, dpairs, _ = H.dpmodel_align(g, seq)
        assert dp.score == dsc and dp.alignment == dpairs, f"member {i}"
        assert len(seq) < 1024 or aligner.last_dp_kernel() == 4
        g.add_alignment_with_weights(name, seq, dpairs)
