"""Writes tests/golden/*.expected.{fa,gfa,scores}: `poasta align` output of the ORACLE (the CPU restatement of the reference)
for the reference's own test inputs, default options (-n 4 -g 6 -e 2 -m global -H mingap). They freeze the restated
behaviour so that any later change of the oracle or the product shows up as a byte difference; they were NOT produced by
the real reference binary (no Rust toolchain in this image, DESIGN.md section 5: whole-MSA parity is unpinned).
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
