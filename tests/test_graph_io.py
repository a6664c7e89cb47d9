"""Host graph model + writers (poasta_b200/csrc/poa_graph.hpp, poa_io.hpp) against the reference's
own exact-value tests: tests/io_fasta.rs, src/io/fasta.rs:165-197, src/graphs/poa.rs:507-559,
tests/io_graph.rs, tests/poa_graph.rs."""
import numpy as np
import pytest

from poasta_b200 import AlignedPair, InvalidAlignment, POAGraph, WeightsUnequalSize


def test_new_graph_is_empty():
    # graphs/poa.rs:507-516
    g = POAGraph()
    assert g.is_empty() and g.node_count() == 0 and g.node_count_with_start_and_end() == 2
    assert g.edge_count() == 0 and g.num_sequences() == 0


def test_add_alignment_with_weights_basic():
    # graphs/poa.rs:518-559
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACG", None, [1, 1, 1])
    assert not g.is_empty() and g.node_count() == 3 and g.edge_count() == 2 and g.num_sequences() == 1
    a = g.view_arrays()
    topo = list(a["topo_order"])
    assert topo[0] == a["start_node"] and topo[-1] == a["end_node"]
    real = [v for v in topo if v not in (a["start_node"], a["end_node"])]
    assert bytes(a["symbol"][real]) == b"ACG"
    aln = [AlignedPair(real[0], 0), AlignedPair(real[1], 1), AlignedPair(real[2], 2)]
    g.add_alignment_with_weights("seq2", b"ACG", aln, [1, 1, 1])
    assert g.node_count() == 3 and g.edge_count() == 2 and g.num_sequences() == 2


def test_poa_graph_to_fasta_basic():
    # tests/io_fasta.rs:5-20 — pins neighbour order, DFS column order and the col-1 leading-gap quirk
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"AC", None)
    g.add_alignment_with_weights("seq2", b"ACGT", None)
    assert g.to_fasta() == b">seq1\n---AC\n>seq2\nACGT--\n"


def test_poa_graph_to_fasta_empty():
    # tests/io_fasta.rs:22-33
    g = POAGraph()
    g.add_alignment_with_weights("empty", b"", None, [])
    assert g.to_fasta() == b">empty\n"


def test_fasta_unit_test_in_io_module():
    # src/io/fasta.rs:165-197: ACG then A-G aligned => rows ACG / A-G
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACG", None)
    aln = [AlignedPair(2, 0), AlignedPair(4, 1)]
    g.add_alignment_with_weights("seq2", b"AG", aln)
    assert g.to_fasta() == b">seq1\nACG\n>seq2\nA-G\n"


def test_weights_unequal_size():
    g = POAGraph()
    with pytest.raises(WeightsUnequalSize):
        g.add_alignment_with_weights("s", b"ACG", None, [1, 1])


def test_invalid_alignment():
    # graphs/poa.rs:210-224: no valid query position for a non-empty sequence
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACG", None)
    with pytest.raises(InvalidAlignment):
        g.add_alignment_with_weights("seq2", b"ACG", [AlignedPair(2, None)])


def test_save_load_roundtrip():
    # tests/io_graph.rs:15-28: node/edge counts survive save -> load; here also the writers' bytes
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACGT", None)
    g.add_alignment_with_weights("seq2", b"AGT", [AlignedPair(2, 0), AlignedPair(4, 1), AlignedPair(5, 2)])
    g2 = POAGraph.load(g.save())
    assert g2.node_count() == g.node_count() and g2.edge_count() == g.edge_count()
    assert g2.to_fasta() == g.to_fasta()
    assert g2.save() == g.save()


def test_toposort_is_topological():
    # tests/poa_graph.rs:50-95: ranks form a valid topological order
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACGTACGT", None)
    g.add_alignment_with_weights("seq2", b"ACTTACGT", [AlignedPair(2 + i, i) for i in range(8)])
    a = g.view_arrays()
    rank = np.zeros(a["n_nodes"], np.int64)
    rank[a["topo_order"]] = np.arange(a["n_nodes"])
    for v in range(a["n_nodes"]):
        for s in a["succ"][a["succ_off"][v]:a["succ_off"][v + 1]]:
            assert rank[v] < rank[s]
    assert a["topo_order"][0] == 0 and a["topo_order"][-1] == 1


def test_gfa_small():
    g = POAGraph()
    g.add_alignment_with_weights("seq1", b"ACGT", None)
    g.add_alignment_with_weights("seq2", b"AGT", [AlignedPair(2, 0), AlignedPair(4, 1), AlignedPair(5, 2)])
    gfa = g.to_gfa().decode().splitlines()
    assert gfa[0] == "H\tVN:Z:1.1"
    segs = [l.split("\t")[2] for l in gfa if l.startswith("S")]
    assert sorted(segs) == ["A", "C", "GT"]
    assert sum(l.startswith("L") for l in gfa) == 3
    walks = [l.split("\t") for l in gfa if l.startswith("W")]
    assert [w[3] for w in walks] == ["seq1", "seq2"]
    v1 = g.to_gfa(v1=True).decode().splitlines()
    assert sum(l.startswith("P") for l in v1) == 2


def test_fasta_msa_import_roundtrip():
    # io/graph.rs:36-103
    g = POAGraph.from_fasta_msa(b">a\nACGT\n>b\nA-GT\n>c\nATGT\n")
    assert g.node_count() == 5 and g.num_sequences() == 3
    assert g.to_fasta() == b">a\nACGT\n>b\nA-GT\n>c\nATGT\n"


def _progressive_graphs():
    """Graphs of growing depth from synthetic families, alignments from the oracle (CPU)."""
    import helpers as H
    from poasta_b200 import synth
    for seed, n, length, rate in ((11, 6, 120, 0.08), (12, 5, 60, 0.2), (13, 8, 200, 0.03)):
        g = POAGraph()
        for name, seq in synth.parse_fasta(synth.family(seed, n, length, rate)):
            if g.is_empty():
                g.add_alignment_with_weights(name, seq, None)
                continue
            st, sc, pairs, _ = H.oracle_align(g, seq)
            assert st == 0
            g.add_alignment_with_weights(name, seq, pairs)
            yield g


def test_roundtrips_on_progressive_graphs():
    """Size-independent properties of the graph side of the path. `.poasta` save -> load -> save is the identity on bytes.
    Loading rebuilds the adjacency lists in edge-index order (petgraph's StableGraph deserialisation), which differs from
    the insertion order of a graph whose start/end edges were re-wired through the edge free list - so GFA / DOT / column
    order may change across the FIRST load, exactly as with the reference; from then on every writer is stable. In every
    form each MSA row spells its sequence, rows have equal length, and the stored order is topological."""
    from poasta_b200 import synth
    n = 0
    for g in _progressive_graphs():
        blob = g.save()
        g2 = POAGraph.load(blob)
        assert g2.save() == blob
        g3 = POAGraph.load(g2.save())
        for v1 in (False, True):
            assert g3.to_gfa(v1=v1) == g2.to_gfa(v1=v1)
        assert g3.to_fasta() == g2.to_fasta() and g3.to_dot() == g2.to_dot()
        assert (g2.node_count(), g2.edge_count(), g2.num_sequences()) == (g.node_count(), g.edge_count(), g.num_sequences())
        spelled = None
        for h in (g, g2):
            rows = synth.parse_fasta(h.to_fasta())
            assert len({len(r) for _, r in rows}) == 1  # equal-length rows (80-column wrapping undone by the parser)
            seqs = [(name, r.replace(b"-", b"")) for name, r in rows]
            assert spelled is None or seqs == spelled
            spelled = seqs
            segs = {}
            for line in h.to_gfa().decode().splitlines():
                f = line.split("\t")
                if f[0] == "S":
                    segs[f[1]] = f[2]
                elif f[0] == "W":  # a walk spells the same sequence as the MSA row of that name
                    walk = "".join(segs[x] for x in f[6].replace(">", " ").split())
                    assert walk.encode() == dict(seqs)[f[3]]
            a = h.view_arrays()
            rank = np.zeros(a["n_nodes"], np.int64)
            rank[a["topo_order"]] = np.arange(a["n_nodes"])
            for v in range(a["n_nodes"]):
                for s_ in a["succ"][a["succ_off"][v]:a["succ_off"][v + 1]]:
                    assert rank[v] < rank[s_]
        msa = g.to_fasta()
        g4 = POAGraph.from_fasta_msa(msa)
        assert g4.to_fasta() == msa and g4.num_sequences() == g.num_sequences()
        n += 1
    assert n == 16


def test_malformed_poasta_files_are_rejected_not_trusted():
    """`view`, `stats` and `align -I` read untrusted files: truncations and corrupted ids / counts must give an error status,
    never an out-of-bounds access, a huge allocation or an exception across the C boundary."""
    import random
    import struct
    from poasta_b200 import POAGraph, synth
    import helpers as H
    rc, blob, _, _ = H.oracle_msa(synth.family(5, 6, 80, 0.1), fmt=3)
    assert rc == 0
    g = POAGraph.load(blob)
    assert g.save() == blob
    rng = random.Random(7)
    # every truncation point in the header and a sample of the rest
    for cut in list(range(0, 64)) + [rng.randrange(len(blob)) for _ in range(300)]:
        with pytest.raises(OSError):
            POAGraph.load(blob[:cut])
    # counts blown up: node count, an aligned-nodes count, the edge count
    for off, val in ((4, 1 << 40), (4, (1 << 31) - 1), (4 + 8 + 1, 1 << 33)):
        bad = bytearray(blob)
        bad[off:off + 8] = struct.pack("<Q", val)
        with pytest.raises(OSError):
            POAGraph.load(bytes(bad))
    # random single-word corruptions: either rejected, or a graph whose every id is in range (all writers must run)
    accepted = 0
    for _ in range(400):
        bad = bytearray(blob)
        off = rng.randrange(4, len(bad) - 4)
        bad[off:off + 4] = struct.pack("<I", rng.choice([0xFFFFFFFF, 0x7FFFFFFF, len(blob), rng.randrange(1 << 32)]))
        try:
            h = POAGraph.load(bytes(bad))
        except OSError:
            continue
        accepted += 1
        h.view_arrays()
        for fn in (h.to_fasta, h.to_gfa, h.save):
            try:
                fn()
            except OSError:
                pass
    assert accepted < 400


def test_add_alignment_argument_checks():
    from poasta_b200 import AlignedPair, InvalidAlignment, POAGraph, WeightsUnequalSize
    g = POAGraph()
    g.add_alignment_with_weights("a", b"ACGT", None, weights=[1, 1, 1, 1])
    with pytest.raises(WeightsUnequalSize):  # poa.rs:178-180, checked behind the C boundary
        g.add_alignment_with_weights("b", b"ACGT", None, weights=[1, 1, 1])
    with pytest.raises(InvalidAlignment):  # node id out of range: the reference panics on graph[rpos]
        g.add_alignment_with_weights("c", b"AC", [AlignedPair(999, 0), AlignedPair(3, 1)])
    with pytest.raises(InvalidAlignment):  # query position out of range: the reference panics on sequence[qpos]
        g.add_alignment_with_weights("d", b"AC", [AlignedPair(2, 0), AlignedPair(3, 7)])
    assert g.num_sequences() == 1
    # a column of more than 64 mutually aligned nodes is reported in full
    h = POAGraph()
    h.add_alignment_with_weights("s0", bytes([33]), None)
    for k in range(1, 80):
        h.add_alignment_with_weights(f"s{k}", bytes([33 + k]), [AlignedPair(2, 0)])
    assert len(h.get_aligned_nodes(2)) == 79
