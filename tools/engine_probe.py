"""GPU probe of the persistent MSA engine: F config-2 families through msa_batch (reference-exact), wall / kernel time, and a
bit-exact check of the first `check` families against the oracle. usage: python tools/engine_probe.py F [check] [seqs] [length]"""
import sys
import time

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402
from poasta_b200 import aligner, msa_batch, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 592
check = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seqs = int(sys.argv[3]) if len(sys.argv) > 3 else 20
length = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
rate = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
fastas = [synth.config2_family(f, n_seqs=seqs, root_len=length, rate=rate) for f in range(F)]
import os
for rep in range(int(os.environ.get('PROBE_REPS', '2'))):
    t0 = time.perf_counter()
    graphs, scores, cells = msa_batch(fastas, device=0)
    wall = time.perf_counter() - t0
    ms = aligner.last_batch_search_ms()
    h2d, d2h = aligner.last_batch_io()
    print(f"rep {rep}: {F} families, wall {wall:.2f} s, kernel {ms / 1e3:.2f} s, {cells / wall / 1e9:.2f} GCUPS e2e, "
          f"{cells / ms / 1e6:.2f} GCUPS kernel, h2d {h2d / 1e6:.1f} MB, d2h {d2h / 1e6:.1f} MB, counts {aligner.last_batch_search_counts()}", flush=True)
bad = 0
for f in range(min(check, F)):
    rc, expect, osc, _ = H.oracle_msa(fastas[f])
    ok = rc == 0 and list(scores[f]) == list(osc) and graphs[f].to_fasta() == expect and graphs[f].to_gfa() == H.oracle_msa(fastas[f], fmt=1)[1]
    bad += not ok
print("oracle check:", "OK" if not bad else f"{bad} families differ", f"({min(check, F)} families)")
sys.exit(1 if bad else 0)
