"""Random-parameter fuzz: the dense model of the DP kernels (oracle/dp_model.hpp, global mode) against the A* oracle.
The DP optimum must never exceed the reference's score; it may be lower only through the reference's offset-0 special
case. CPU only. usage: python tools/fuzz_dp_model_vs_oracle.py [n_cases] [seed]"""
import random
import sys

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402
from poasta_b200 import POAGraph, synth  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    n_aln = lower = 0
    for case in range(n_cases):
        n_seqs, length, rate = rng.randint(3, 8), rng.randint(8, 250), rng.choice([0.01, 0.05, 0.1, 0.2, 0.35])
        x, o, e = rng.randint(1, 9), rng.randint(0, 12), rng.randint(1, 4)
        g = POAGraph()
        for name, seq in synth.parse_fasta(synth.family(rng.randint(0, 1 << 30), n_seqs, length, rate)):
            if rng.random() < 0.1:
                seq = seq[: rng.randint(1, max(1, len(seq)))]
            if g.is_empty():
                g.add_alignment_with_weights(name, seq, None)
                continue
            a = H.oracle_align(g, seq, x=x, o=o, e=e, mode=0, heuristic=0)  # Dijkstra: admissible
            d = H.dpmodel_align(g, seq, x=x, o=o, e=e)
            dp_score = d[0]
            if a[0] == 0:
                if dp_score > a[1]:
                    print("DP ABOVE REFERENCE", dict(case=case, x=x, o=o, e=e, length=length, rate=rate, dp=dp_score, ref=a[1]))
                    return 1
                lower += dp_score < a[1]
                g.add_alignment_with_weights(name, seq, a[2])
            n_aln += 1
    print(f"{n_cases} cases, {n_aln} alignments: DP optimum <= reference score everywhere, strictly lower in {lower}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
