"""CPU checks of the code the device runs for graph maintenance (poasta_b200/csrc/dev_graph.hpp, dev_index.hpp), compiled for the
host by tests/hooks: progressive builds with the oracle's alignments, the arena graph compared field by field with the
std::vector graph after every insertion, the per-round derived arrays compared with search_index.hpp, and the download path
compared byte for byte (.poasta)."""
import ctypes as C
import glob
import os

import pytest

import helpers as H
from poasta_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(fa, heuristic=1, mode=0, graph_in=None, costs=(4, 6, 2)):
    lib = H.oracle()
    lib.devgraph_host_check.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t] + [C.c_int] * 5 + [C.c_char_p, C.c_size_t]
    msg = C.create_string_buffer(512)
    rc = lib.devgraph_host_check(fa, len(fa), graph_in, len(graph_in) if graph_in else 0, *costs, mode, heuristic, msg, 512)
    return rc, msg.value.decode()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.fa"))))
def test_reference_fixtures(path):
    fa = open(path, "rb").read()
    for heuristic in (0, 1, 2):
        assert _check(fa, heuristic) == (0, "")


def test_synthetic_families_incl_path_index():
    for f in range(8):
        fa = synth.config2_family(f, n_seqs=8, root_len=200, rate=0.1)
        for heuristic in (1, 2):
            assert _check(fa, heuristic) == (0, "")
    assert _check(synth.family(77, 6, 150, 0.25), 2) == (0, "")  # divergent: deep bubble nesting, many aligned columns


def test_full_size_family_and_existing_graph():
    fa = synth.config2_family(3, n_seqs=20, root_len=1000, rate=0.05)
    assert _check(fa) == (0, "")
    recs = synth.parse_fasta(synth.family(5, 8, 120, 0.08))
    text = lambda r: b"".join(b">%s\n%s\n" % (n.encode(), s) for n, s in r)
    rc, graph, _, _ = H.oracle_msa(text(recs[:4]), fmt=3)
    assert rc == 0
    assert _check(text(recs[4:]), graph_in=graph) == (0, "")


def test_degenerate_records():
    assert _check(b">a\nACGT\n>b\n\n>c\nACGA\n") == (0, "")
    assert _check(b">a\nA\n>b\nC\n>c\nA\n>d\nG\n>e\nT\n>f\nN\n") == (0, "")  # one column of five mutually aligned nodes


def test_pool_backed_lists_of_a_downloaded_graph():
    """PoolList lifetimes (poa_graph.hpp): download, growth past the pool slice, copy / assign / move, second download, continuing
    the build on the host - every graph must serialise like its std-path twin."""
    lib = H.oracle()
    lib.poollist_lifetime_check.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    for fa in (synth.config2_family(4, n_seqs=8, root_len=300, rate=0.08), synth.family(78, 6, 150, 0.25)):
        rc, data, _, _ = H.oracle_msa(fa, fmt=3)
        assert rc == 0
        msg = C.create_string_buffer(512)
        assert (lib.poollist_lifetime_check(data, len(data), msg, 512), msg.value.decode()) == (0, "")
