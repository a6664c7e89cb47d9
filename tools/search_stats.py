"""Event counts of the order-exact search (host build of poasta_b200/csrc/replay_core.hpp, -DRP_COUNT) on config-2
families: what one alignment does, by round. CPU only; uses the oracle library's test hooks."""
import ctypes as C
import sys

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402
from poasta_b200 import POAGraph, synth  # noqa: E402

NAMES = ["pop_M", "pop_D", "pop_I", "pop_stale", "dfa_step", "dfa_match", "dfa_mismatch", "dfa_relax_fail", "prune_call", "prune_bub",
         "prune_noreach", "prune_probe", "prune_score", "push", "relax", "relax_ok", "window_move", "pop_scan",
         "next_same_stack", "next_just_pushed", "next_other"]


def counters():
    buf = (C.c_uint64 * 32)()
    n = H.oracle().replay_host_counters(buf, 32)
    return dict(zip(NAMES, list(buf)[:n]))


def main():
    fam = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    g = POAGraph()
    counters()
    for i, (name, seq) in enumerate(synth.parse_fasta(synth.config2_family(fam, n_seqs=n_seqs))):
        if g.is_empty():
            g.add_alignment_with_weights(name, seq, None)
            continue
        st, sc, pairs, stats = H.replay_host_align(g, seq, cell16=True)
        c = counters()
        print(f"round {i}: V={g.node_count()} score={sc} queued={stats[0]} visited={stats[1]} pruned={stats[2]}")
        print("   ", " ".join(f"{k}={v}" for k, v in c.items()))
        g.add_alignment_with_weights(name, seq, pairs)


if __name__ == "__main__":
    main()
