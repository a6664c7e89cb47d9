"""`poasta` CLI (poasta_b200/csrc/cli_main.cpp): mirrors the reference's tests/cli_poasta.rs plus byte checks."""
import os
import subprocess

import pytest

import helpers as H
from poasta_b200 import AlignedPair, POAGraph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "poasta_b200", "poasta")
SMALL = os.path.join(ROOT, "tests", "golden", "small_test.fa")


def _run(*args):
    return subprocess.run([BIN, *args], capture_output=True)


def _small_graph():
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACGT", None)
    g.add_alignment_with_weights("seq2", b"AGT", [AlignedPair(2, 0), AlignedPair(4, 1), AlignedPair(5, 2)])
    return g


def test_view_converts_without_a_gpu(tmp_path):
    g = _small_graph()
    p = tmp_path / "g.poasta"
    p.write_bytes(g.save())
    for fmt, expect in (("fasta", g.to_fasta()), ("gfa", g.to_gfa()), ("gfav1", g.to_gfa(v1=True)), ("dot", g.to_dot()),
                        ("poasta", g.save())):
        r = _run("view", "-O", fmt, str(p))
        assert r.returncode == 0 and r.stdout == expect, fmt
    out = tmp_path / "sub" / "dir" / "g.gfa"
    assert _run("view", "-O", "gfa", "-o", str(out), str(p)).returncode == 0  # -o creates parent directories
    assert out.read_bytes() == g.to_gfa()
    # a column FASTA MSA is accepted as a graph when the name ends in .fa/.fasta (bin/poasta.rs:304-311)
    msa = tmp_path / "msa.fa"
    msa.write_bytes(g.to_fasta())
    assert _run("view", "-O", "fasta", str(msa)).stdout == g.to_fasta()


def test_stats_goes_to_stderr(tmp_path):
    p = tmp_path / "g.poasta"
    p.write_bytes(_small_graph().save())
    r = _run("stats", str(p))
    assert r.returncode == 0 and r.stdout == b"" and b"node_count: 4" in r.stderr and b"edge_count: 4" in r.stderr


def test_flag_errors():
    assert _run().returncode != 0
    assert _run("view", "x.poasta").returncode == 2                       # -O is required for view
    assert _run("align", "-H", "nonsense", SMALL).returncode == 1
    assert _run("align", "-g", "6,7", SMALL).returncode == 1            # needs one value each (or two each)
    r = _run("align", "-g", "8,24", "-e", "2,1", SMALL)
    assert r.returncode == 1 and b"two-piece" in r.stderr


def test_align_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("align", "-O", "fasta", SMALL)
    assert r.returncode == 1 and b"no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_align_outputs_fasta():
    # tests/cli_poasta.rs:16-23 (stdout contains >seq1) + bytes equal to the oracle-driven build
    r = _run("align", "-O", "fasta", SMALL)
    assert r.returncode == 0 and b">seq1" in r.stdout
    rc, expect, _, _ = H.oracle_msa(open(SMALL, "rb").read())
    assert rc == 0 and r.stdout == expect


@pytest.mark.gpu
def test_align_then_view_gfa(tmp_path):
    # tests/cli_poasta.rs:25-42
    graph = tmp_path / "out.poasta"
    assert _run("align", "-o", str(graph), SMALL).returncode == 0
    r = _run("view", "-O", "gfa", str(graph))
    assert r.returncode == 0 and b"S\t" in r.stdout
    rc, expect, _, _ = H.oracle_msa(open(SMALL, "rb").read(), fmt=1)
    assert r.stdout == expect


@pytest.mark.gpu
@pytest.mark.parametrize("span,mode,hname,heur", [("global", 0, "path", 2), ("ends-free", 1, "dijkstra", 0), ("ends-free", 1, "path", 2),
                                                  ("semi-global", 1, "mingap", 1)])
def test_align_into_existing_graph(tmp_path, span, mode, hname, heur):
    # config 5 shape: -I graph, -m ends-free, -H path (every read is also inserted, bin/poasta.rs:226). The reference's
    # ends-free search can return an empty alignment, which its graph insertion rejects (InvalidAlignment): whatever the
    # oracle-driven run does — output bytes or error — the CLI must do the same.
    fa = open(os.path.join(ROOT, "tests", "golden", "test2_from_abpoa.fa"), "rb").read()
    recs = fa.split(b">")[1:]
    first, rest = b">" + b">".join(recs[:6]), b">" + b">".join(recs[6:])
    (tmp_path / "first.fa").write_bytes(first)
    (tmp_path / "rest.fa").write_bytes(rest)
    graph = tmp_path / "g.poasta"
    assert _run("align", "-o", str(graph), str(tmp_path / "first.fa")).returncode == 0
    rc, g0, _, _ = H.oracle_msa(first, fmt=3)
    assert rc == 0 and graph.read_bytes() == g0
    r = _run("align", "-I", str(graph), "-m", span, "-H", hname, "-O", "gfa", str(tmp_path / "rest.fa"))
    rc2, expect, _, _ = H.oracle_msa(rest, mode=mode, heuristic=heur, fmt=1, graph_in=g0)
    if rc2 == 0:
        assert r.returncode == 0 and r.stdout == expect
    else:
        assert rc2 == -4 and r.returncode == 1 and b"poasta_status -4" in r.stderr
