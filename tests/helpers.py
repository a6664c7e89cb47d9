"""Test-side bindings: the oracle (oracle/liboracle.so) and small graph builders.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from poasta_b200 import _lib as plib
from poasta_b200.graph import NONE, AlignedPair, POAGraph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
HOOKS_PATH = os.path.join(ROOT, "tests", "hooks", "libproduct_hooks.so")
_oracle = None
_hooks = None

MATCH, DELETION, INSERTION = 0, 1, 2
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class _Both:
    """Symbols of the oracle library first, of the product-hook library (host builds of product headers) second."""

    def __init__(self, a, b):
        self._a, self._b = a, b

    def __getattr__(self, name):
        try:
            return getattr(self._a, name)
        except AttributeError:
            return getattr(self._b, name)


def oracle():
    global _oracle
    if _oracle is None:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hooks")], check=True, capture_output=True)
        o = _Both(C.CDLL(ORACLE_PATH), C.CDLL(HOOKS_PATH))
        GV = C.POINTER(plib.GraphView)
        o.oracle_align.argtypes = [GV, C.c_char_p, C.c_uint32] + [C.c_int] * 6 + [u32p, u32p, u32p, C.c_uint32, u32p,
                                                                                 C.POINTER(C.c_uint64)]
        o.oracle_align_dump.argtypes = [GV, C.c_char_p, C.c_uint32] + [C.c_int] * 6 + [u32p, C.c_uint64]
        o.oracle_align_dump.restype = C.c_uint64
        o.oracle_rev_postorder.argtypes = [GV, C.c_int, u32p]
        o.oracle_rev_postorder.restype = C.c_uint32
        o.oracle_superbubbles.argtypes = [GV, C.c_int, u32p, C.c_uint32]
        o.oracle_superbubbles.restype = C.c_uint32
        o.oracle_bubble_index.argtypes = [GV, C.c_int, C.c_uint32, u32p, u32p, u32p]
        o.oracle_heuristic.argtypes = [GV] + [C.c_int] * 5 + [C.c_uint32] * 3 + [C.c_int] * 3
        o.oracle_heuristic.restype = C.c_uint64
        o.oracle_gap_cost.argtypes = [C.c_int] * 4 + [C.c_uint64]
        o.oracle_gap_cost.restype = C.c_uint64
        o.oracle_dfa_event.argtypes = [GV, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u32p]
        o.oracle_msa.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t] + [C.c_int] * 6 + [
            C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), u32p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        o.oracle_msa_many.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint32] + [C.c_int] * 6 + [
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        o.oracle_msa_many.restype = C.c_double
        o.oracle_align_many.argtypes = [GV, C.POINTER(C.c_char_p), u32p, C.c_uint32] + [C.c_int] * 6 + [u32p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        o.oracle_align_many.restype = C.c_double
        o.dpmodel_align.argtypes = [GV, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, u32p, u32p, C.c_uint32, u32p, u32p]
        o.dpmodel_align.restype = C.c_uint32
        o.replay_host_align.argtypes = [GV, C.c_char_p, C.c_uint32] + [C.c_int] * 5 + [u32p, u32p, u32p, C.c_uint32, u32p,
                                                                                    C.POINTER(C.c_uint64), C.c_int]
        o.oracle_free.argtypes = [C.c_void_p]
        _oracle = o
    return _oracle


class View:
    """Owns numpy arrays behind a ``poasta_graph_view``."""

    def __init__(self, n_nodes, start, end, topo, pred_lists, succ_lists, symbols):
        self.n = n_nodes
        self.topo = np.ascontiguousarray(topo, dtype=np.uint32)
        self.pred_off = np.zeros(n_nodes + 1, np.uint32)
        self.succ_off = np.zeros(n_nodes + 1, np.uint32)
        for v in range(n_nodes):
            self.pred_off[v + 1] = self.pred_off[v] + len(pred_lists[v])
            self.succ_off[v + 1] = self.succ_off[v] + len(succ_lists[v])
        self.pred = np.array([p for l in pred_lists for p in l] + [0], dtype=np.uint32)
        self.succ = np.array([s for l in succ_lists for s in l] + [0], dtype=np.uint32)
        self.symbol = np.ascontiguousarray(symbols, dtype=np.uint8)
        v = plib.GraphView()
        v.n_nodes, v.start_node, v.end_node = n_nodes, start, end
        v.topo_order = self.topo.ctypes.data_as(u32p)
        v.pred_off = self.pred_off.ctypes.data_as(u32p)
        v.pred = self.pred.ctypes.data_as(u32p)
        v.succ_off = self.succ_off.ctypes.data_as(u32p)
        v.succ = self.succ.ctypes.data_as(u32p)
        v.symbol = self.symbol.ctypes.data_as(u8p)
        self.c = v

    @classmethod
    def from_arrays(cls, a):
        n = a["n_nodes"]
        preds = [list(a["pred"][a["pred_off"][v]:a["pred_off"][v + 1]]) for v in range(n)]
        succs = [list(a["succ"][a["succ_off"][v]:a["succ_off"][v + 1]]) for v in range(n)]
        return cls(n, a["start_node"], a["end_node"], a["topo_order"], preds, succs, a["symbol"])


def mock_view(n_nodes, edges, symbols=None):
    """petgraph ``DiGraph`` as in graphs/mock.rs: adjacency newest-edge-first; start = node 0, end = last."""
    preds = [[] for _ in range(n_nodes)]
    succs = [[] for _ in range(n_nodes)]
    for s, t in edges:
        succs[s].insert(0, t)
        preds[t].insert(0, s)
    # any topological order (ranks only index the visited table)
    indeg = [len(p) for p in preds]
    order, ready = [], [v for v in range(n_nodes) if indeg[v] == 0]
    while ready:
        v = ready.pop(0)
        order.append(v)
        for s in succs[v]:
            indeg[s] -= 1
            if indeg[s] == 0:
                ready.append(s)
    assert len(order) == n_nodes
    if symbols is None:
        symbols = [ord('-')] * n_nodes
    return View(n_nodes, 0, n_nodes - 1, order, preds, succs, symbols)


def test_graph1():
    """create_test_graph1 (graphs/mock.rs:92-125); node values 1..9 are indices 0..8, mock end = index 9."""
    edges = [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (3, 7), (7, 8), (8, 9)]
    e = [(s - 1, t - 1) for s, t in edges]
    out_deg = [0] * 10
    for s, _ in e:
        out_deg[s] += 1
    for n in range(9):
        if out_deg[n] == 0:
            e.append((n, 9))
    return mock_view(10, e)


def test_graph2():
    """create_test_graph2 (graphs/mock.rs:127-165)."""
    edges = [(1, 2), (1, 3), (2, 3), (3, 4), (3, 5), (3, 11), (4, 8), (5, 6), (5, 9), (6, 7), (6, 10), (7, 8), (8, 13),
             (8, 15), (9, 10), (10, 7), (11, 12), (12, 8), (13, 14), (13, 15), (14, 15)]
    return mock_view(15, [(s - 1, t - 1) for s, t in edges])


def _cview(view):
    """Accepts a View, a POAGraph (kept alive by the caller's reference) or a raw GraphView."""
    if isinstance(view, View):
        return view.c
    if isinstance(view, POAGraph):
        return view.view()
    return view


def pairs_from_arrays(r, q, n):
    return [AlignedPair(None if r[i] == NONE else int(r[i]), None if q[i] == NONE else int(q[i])) for i in range(n)]


def oracle_align(view, seq, x=4, o=6, e=2, mode=0, heuristic=1, mock=False, pruning=True):
    """PoastaAligner::align through the A* restatement. Returns (status, score, pairs, stats)."""
    seq = bytes(seq)
    cview = _cview(view)
    cap = cview.n_nodes + len(seq) + 8
    r = np.zeros(cap, np.uint32)
    q = np.zeros(cap, np.uint32)
    score, n = C.c_uint32(), C.c_uint32()
    stats = (C.c_uint64 * 5)()
    flags = (1 if mock else 0) | (0 if pruning else 2)
    st = oracle().oracle_align(C.byref(cview), seq, len(seq), x, o, e, mode, heuristic, flags, C.byref(score),
                               r.ctypes.data_as(u32p), q.ctypes.data_as(u32p), cap, C.byref(n), stats)
    return st, score.value, pairs_from_arrays(r, q, n.value), list(stats)


def replay_host_align(view, seq, x=4, o=6, e=2, mode=0, heuristic=1, cell16=False):
    """Host build of the product's order-exact search core (test hook). Returns (status, score, pairs, stats)."""
    seq = bytes(seq)
    cview = _cview(view)
    cap = cview.n_nodes + len(seq) + 8
    r = np.zeros(cap, np.uint32)
    q = np.zeros(cap, np.uint32)
    score, n = C.c_uint32(), C.c_uint32()
    stats = (C.c_uint64 * 5)()
    st = oracle().replay_host_align(C.byref(cview), seq, len(seq), x, o, e, mode, heuristic, C.byref(score),
                                    r.ctypes.data_as(u32p), q.ctypes.data_as(u32p), cap, C.byref(n), stats, 1 if cell16 else 0)
    return st, score.value, pairs_from_arrays(r, q, n.value), list(stats)


def dpmodel_align(view, seq, x=4, o=6, e=2, tables=False):
    seq = bytes(seq)
    cview = _cview(view)
    cap = cview.n_nodes + len(seq) + 8
    r = np.zeros(cap, np.uint32)
    q = np.zeros(cap, np.uint32)
    n = C.c_uint32()
    t = None
    tp = None
    if tables:
        t = np.zeros((3, cview.n_nodes, len(seq) + 1), np.uint32)
        tp = t.ctypes.data_as(u32p)
    score = oracle().dpmodel_align(C.byref(cview), seq, len(seq), x, o, e, r.ctypes.data_as(u32p), q.ctypes.data_as(u32p),
                                   cap, C.byref(n), tp)
    return score, pairs_from_arrays(r, q, n.value), t


def oracle_dump(view, seq, x=4, o=6, e=2, mode=0, heuristic=1):
    seq = bytes(seq)
    cview = _cview(view)
    cap = 3 * cview.n_nodes * (len(seq) + 9) + 1024
    out = np.zeros((cap, 4), np.uint32)
    n = oracle().oracle_align_dump(C.byref(cview), seq, len(seq), x, o, e, mode, heuristic, 0, out.ctypes.data_as(u32p), cap)
    return out[:min(n, cap)]


def oracle_msa(fasta, x=4, o=6, e=2, mode=0, heuristic=1, fmt=0, graph_in=None):
    """`poasta align` with the oracle aligner. Returns (rc, output bytes, scores, cells)."""
    if isinstance(fasta, str):
        fasta = fasta.encode()
    nseq = fasta.count(b">")
    out, n = C.c_void_p(), C.c_size_t()
    scores = np.zeros(max(nseq, 1), np.uint32)
    cells = C.c_uint64()
    stats = (C.c_uint64 * 3)()
    gi = graph_in if graph_in else None
    rc = oracle().oracle_msa(fasta, len(fasta), gi, len(graph_in) if graph_in else 0, x, o, e, mode, heuristic, fmt,
                             C.byref(out), C.byref(n), scores.ctypes.data_as(u32p), nseq, C.byref(cells), stats)
    data = b""
    if rc == 0:
        data = C.string_at(out, n.value)
        oracle().oracle_free(out)
    return rc, data, scores[:nseq], cells.value


def oracle_msa_stats(fasta, x=4, o=6, e=2, mode=0, heuristic=1):
    """(num_queued, num_visited, num_pruned) summed over every alignment of the oracle-driven build of one family."""
    if isinstance(fasta, str):
        fasta = fasta.encode()
    nseq = max(1, fasta.count(b">"))
    scores = np.zeros(nseq, np.uint32)
    cells = C.c_uint64()
    stats = (C.c_uint64 * 3)()
    rc = oracle().oracle_msa(fasta, len(fasta), None, 0, x, o, e, mode, heuristic, 0, None, None, scores.ctypes.data_as(u32p), nseq,
                             C.byref(cells), stats)
    assert rc == 0
    return tuple(stats)


def oracle_align_many(graph, reads, threads, x=4, o=6, e=2, mode=0, heuristic=1, want_scores=False):
    """Independent reads against one fixed graph with the oracle over `threads` threads: (wall seconds, cells[, scores])."""
    n = len(reads)
    ptr = (C.c_char_p * n)(*reads)
    lens = (C.c_uint32 * n)(*[len(r) for r in reads])
    scores = np.zeros(max(n, 1), np.uint32)
    cells, err = C.c_uint64(), C.c_int()
    view = _cview(graph)
    secs = oracle().oracle_align_many(C.byref(view), ptr, lens, n, x, o, e, mode, heuristic, threads, scores.ctypes.data_as(u32p),
                                      C.byref(cells), C.byref(err))
    if err.value:
        raise RuntimeError(f"oracle panic {err.value}")
    return (secs, cells.value, scores[:n]) if want_scores else (secs, cells.value)


def graph_from_seqs(seqs, aligner):
    """Progressive build like bin/poasta.rs:191-233 with a caller-supplied ``aligner(view, seq) -> pairs``."""
    g = POAGraph()
    for i, s in enumerate(seqs):
        if g.is_empty():
            g.add_alignment_with_weights(f"seq{i + 1}", s, None)
        else:
            pairs = aligner(g.view(), s)
            g.add_alignment_with_weights(f"seq{i + 1}", s, pairs)
    return g
