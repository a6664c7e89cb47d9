"""Writes tests/golden/*.expected.{fa,gfa,scores}: `poasta align` output of the ORACLE (the CPU restatement of the reference)
for the reference's own test inputs, default options (-n 4 -g 6 -e 2 -m global -H mingap). Since round 2 the oracle is an
independent restatement (aligner, graph and writers share no code with the product), so these files are "oracle-frozen", not
"self-frozen": the product must reproduce them on the CPU (graph / writers) and on the GPU (whole MSA). They were NOT produced
by the real reference binary (no Rust toolchain in this image, DESIGN.md section 5: whole-MSA parity stays unpinned).
Note: the reference's own tests/small_test.truth.fa differs from small_test.expected.fa by design: it has 6 leading gaps where
the current io/fasta.rs writer emits `col - 1` = 5 (saturating_sub); no reference test opens that file.
usage: python tools/make_golden.py"""
import os
import sys

sys.path[:0] = ["tests", "."]
import helpers as H  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    for name in ("small_test", "test_from_abpoa", "test2_from_abpoa"):
        fasta = open(os.path.join(gold, name + ".fa"), "rb").read()
        for fmt, ext in ((0, "fa"), (1, "gfa")):
            rc, data, scores, _ = H.oracle_msa(fasta, fmt=fmt)
            assert rc == 0, (name, rc)
            open(os.path.join(gold, f"{name}.expected.{ext}"), "wb").write(data)
        open(os.path.join(gold, name + ".expected.scores"), "w").write(" ".join(str(int(s)) for s in scores) + "\n")
        print(name, len(scores), "sequences")


if __name__ == "__main__":
    main()
