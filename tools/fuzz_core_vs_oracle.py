"""Random-parameter fuzz: the product's search core (host build) against the A* oracle — score, pairs and counters.
CPU only. usage: python tools/fuzz_core_vs_oracle.py [n_cases] [seed]"""
import random
import sys

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402
from poasta_b200 import POAGraph, synth  # noqa: E402
from poasta_b200.errors import InvalidAlignment  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    n_aln = 0
    for case in range(n_cases):
        n_seqs, length, rate = rng.randint(3, 8), rng.randint(8, 300), rng.choice([0.01, 0.05, 0.1, 0.2, 0.35])
        x, o, e = rng.randint(1, 9), rng.randint(0, 12), rng.randint(1, 4)
        mode, heur = rng.choice([0, 0, 1]), rng.choice([0, 1, 1, 2])
        fasta = synth.family(rng.randint(0, 1 << 30), n_seqs, length, rate)
        g = POAGraph()
        for name, seq in synth.parse_fasta(fasta):
            if rng.random() < 0.1:
                seq = seq[: rng.randint(0, len(seq))]  # truncated and empty queries
            if g.is_empty():
                if not seq:
                    continue
                g.add_alignment_with_weights(name, seq, None)
                continue
            a = H.oracle_align(g, seq, x=x, o=o, e=e, mode=mode, heuristic=heur)
            for cell16 in (False, True):
                b = H.replay_host_align(g, seq, x=x, o=o, e=e, mode=mode, heuristic=heur, cell16=cell16)
                # a non-zero status mirrors a panic of the reference (e.g. the NonMaxU32 underflow in its backtrace): pairs undefined
                if (b[0], b[1], b[3][:3]) != (a[0], a[1], a[3][:3]) or (a[0] == 0 and b[2] != a[2]):
                    print("MISMATCH", dict(case=case, n_seqs=n_seqs, length=length, rate=rate, x=x, o=o, e=e, mode=mode, heur=heur,
                                           cell16=cell16, oracle=(a[0], a[1], a[3][:3]), core=(b[0], b[1], b[3][:3])))
                    return 1
            n_aln += 1
            if a[0] == 0 and mode == 0 and a[2]:  # Some(empty alignment) is InvalidAlignment in the reference too
                try:
                    g.add_alignment_with_weights(name, seq, a[2])
                except InvalidAlignment:
                    # seen with the inadmissible path heuristic: "insert the whole query at start, delete the whole graph";
                    # the backtrace stops at the start node, so no pair carries a query position (the reference's insert
                    # rejects it the same way)
                    break
    print(f"{n_cases} cases, {n_aln} alignments: product search core == oracle")
    return 0


if __name__ == "__main__":
    sys.exit(main())
