"""Small engine workload for ncu captures: 148 config-2 families cut to their first 4 records (one launch of msa_engine_kernel)."""
import sys
sys.path[:0] = ["tests", "."]
from poasta_b200 import msa_batch, synth  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
fastas = [b">".join(synth.config2_family(f).split(b">")[:5]) for f in range(n)]
graphs, scores, cells = msa_batch(fastas, device=0)
print("cells", cells)
