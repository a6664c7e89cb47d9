"""Isolated DP-kernel throughput: N alignments of config-2 shape (graphs of round `rnd`, 1 kb queries) in ONE align_batch call in
DP mode; prints kernel GCUPS (CUDA events around fill + traceback). usage: python tools/dp_kernel_probe.py [N] [rnd] [graphs]"""
import sys
import time

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402
from poasta_b200 import POAGraph, TIEBREAK_DP, aligner, align_batch, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3552
rnd = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n_graphs = int(sys.argv[3]) if len(sys.argv) > 3 else 24
graphs, queries, gidx = [], [], []
for f in range(n_graphs):
    recs = synth.parse_fasta(synth.config2_family(f))
    g = POAGraph()
    for name, seq in recs[:rnd]:
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
        else:
            g.add_alignment_with_weights(name, seq, H.dpmodel_align(g, seq)[1])
    graphs.append(g)
for i in range(N):
    f = i % n_graphs
    recs = synth.parse_fasta(synth.config2_family(f))
    queries.append(recs[rnd + (i // n_graphs) % (20 - rnd)][1])
    gidx.append(f)
import os
for rep in range(int(os.environ.get('PROBE_REPS', '3'))):
    t0 = time.perf_counter()
    res = align_batch(graphs, queries, gidx, tiebreak=TIEBREAK_DP)
    wall = time.perf_counter() - t0
    kms, launches, cells = aligner.last_batch_stats()
    print(f"rep {rep}: {N} alignments, V ~ {graphs[0].node_count()}, cells {cells / 1e9:.2f} G, kernels {kms:.2f} ms = {cells / kms / 1e6:.1f} GCUPS, wall {wall:.2f} s", flush=True)
sc, pairs, _ = H.dpmodel_align(graphs[0], queries[0])
print("check:", "OK" if res[0].score == sc and res[0].alignment == pairs else "MISMATCH")
