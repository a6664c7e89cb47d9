"""Measure how often the dense DP model (what the CUDA kernels compute) returns the same alignment
as the A* oracle, alignment by alignment along oracle-driven progressive builds. Dev tool."""
import sys
import time

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import helpers as H
from poasta_b200 import POAGraph, synth


def study(fasta, label, heuristic=1, verbose=False):
    recs = synth.parse_fasta(fasta)
    g = POAGraph()
    n = same_score = same_path = 0
    t_or = t_dp = 0.0
    first_bad = None
    for name, seq in recs:
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        t0 = time.time()
        st, sc, pairs, stats = H.oracle_align(g, seq, heuristic=heuristic)
        t1 = time.time()
        dsc, dpairs, _ = H.dpmodel_align(g, seq)
        t2 = time.time()
        t_or += t1 - t0
        t_dp += t2 - t1
        assert st == 0
        n += 1
        same_score += sc == dsc
        ok = pairs == dpairs
        same_path += ok
        if not ok and first_bad is None:
            first_bad = (name, g.node_count(), len(seq), sc, dsc)
            if verbose:
                k = next(i for i in range(min(len(pairs), len(dpairs))) if pairs[i] != dpairs[i])
                print("  first diff at pair", k, pairs[max(0, k - 3):k + 6], dpairs[max(0, k - 3):k + 6])
        g.add_alignment_with_weights(name, seq, pairs)
    print(f"{label}: alignments={n} same_score={same_score} same_path={same_path} "
          f"oracle={t_or:.2f}s dp={t_dp:.2f}s first_bad={first_bad}")
    return n, same_score, same_path


if __name__ == "__main__":
    tot = [0, 0, 0]
    nfam = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rate = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
    for f in range(nfam):
        r = study(synth.config2_family(f, n_seqs=10, root_len=L, rate=rate), f"fam{f}", verbose=True)
        tot = [a + b for a, b in zip(tot, r)]
    print("TOTAL", tot)
