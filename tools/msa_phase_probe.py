"""Where the wall time of one msa_batch call goes (host phases vs device): python tools/msa_phase_probe.py [families] [dp|exact]
Run with POASTA_TIMING=1 to get the C side's phase line on stderr."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poasta_b200 import TIEBREAK_DP, TIEBREAK_EXACT, POAGraph, aligner, msa_batch, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
mode = TIEBREAK_DP if (len(sys.argv) < 3 or sys.argv[2] == "dp") else TIEBREAK_EXACT
for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
    fas = [synth.config2_family(rep * n + f) for f in range(n)]
    t0 = time.perf_counter()
    graphs, scores, cells = msa_batch(fas, tiebreak=mode)
    t1 = time.perf_counter()
    kms, launches, _ = aligner.last_batch_stats()
    dev = aligner.last_batch_device_ms()
    fill = aligner.last_batch_fill_ms()
    POAGraph.free_many(graphs)
    del graphs
    t2 = time.perf_counter()
    print(f"rep {rep}: msa_batch {t1 - t0:.3f} s (device {dev:.1f} ms, fill alone {fill[0]:.1f} ms in {fill[1]} launches, {launches} launches), "
          f"freeing the graphs {t2 - t1:.3f} s, {cells / (t1 - t0) / 1e9:.1f} GCUPS e2e, {cells / (dev / 1e3) / 1e9:.1f} GCUPS device", flush=True)
