"""Python mirror of the reference's aligner API (/root/reference/src/aligner/mod.rs:40-146) over the C ABI.

    aligner = PoastaAligner(AffineMinGapCost(GapAffine(4, 2, 6)), AlignmentType.Global)
    result = aligner.align(graph, b"ACGT")          # -> AstarResult(score, alignment, ...)
    graph.add_alignment_with_weights("name", b"ACGT", result.alignment)

``align`` runs on the GPU (device 0 unless given); there is no CPU path.
"""
import ctypes as C
import os
import sys
import threading
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .errors import check
from .graph import NONE, AlignedPair, POAGraph


@dataclass(frozen=True)
class GapAffine:
    """GapAffine::new(cost_mismatch, cost_gap_extend, cost_gap_open) — note the reference's argument order
    (aligner/scoring/gap_affine.rs:26-30)."""
    cost_mismatch: int = 4
    cost_gap_extend: int = 2
    cost_gap_open: int = 6


class AlignmentType:
    """aligner/scoring/mod.rs:50-62; EndsFree here is the all-Unbounded variant the CLI builds (bin/poasta.rs:287-301)."""
    Global = 0
    EndsFree = 1


@dataclass(frozen=True)
class AffineDijkstra:
    costs: GapAffine = GapAffine()
    heuristic = 0


@dataclass(frozen=True)
class AffineMinGapCost:
    costs: GapAffine = GapAffine()
    heuristic = 1


@dataclass(frozen=True)
class AffinePathAware:
    costs: GapAffine = GapAffine()
    heuristic = 2


@dataclass
class AstarResult:
    """aligner/astar.rs:81-90. The three counters are the order-exact search's (0 in DP-only mode); num_cells = real nodes x query
    length, the unit GCUPS is quoted on."""
    score: int
    alignment: list
    num_queued: int = 0
    num_visited: int = 0
    num_pruned: int = 0
    num_cells: int = 0
    rpos: np.ndarray = field(default=None, repr=False)
    qpos: np.ndarray = field(default=None, repr=False)


TIEBREAK_EXACT, TIEBREAK_DP = 0, 1

_staging = threading.local()  # align_batch's pair staging buffers

# numpy view of `poasta_result` (_lib.Result)
_RESULT_DTYPE = np.dtype([("score", "<u4"), ("n_pairs", "<u4"), ("rpos", "<u8"), ("qpos", "<u8"), ("capacity", "<u4"), ("status", "<i4"),
                          ("num_cells", "<u8"), ("num_queued", "<u8"), ("num_visited", "<u8"), ("num_pruned", "<u8")])
assert _RESULT_DTYPE.itemsize == C.sizeof(_lib.Result)


def _params(config, aln_type, tiebreak=TIEBREAK_EXACT):
    c = config.costs
    return _lib.Params(c.cost_mismatch, c.cost_gap_open, c.cost_gap_extend, aln_type, config.heuristic, tiebreak)


def align_batch(graphs, queries, graph_of_query=None, config=AffineMinGapCost(), aln_type=AlignmentType.Global, device=0,
                tiebreak=TIEBREAK_EXACT):
    """``poasta_align_batch``: independent (graph, query) problems in one device batch. Returns AstarResult list."""
    lib = _lib.load()
    t_start = time.perf_counter()
    n = len(queries)
    # POAGraph objects go in as handles (the library keeps what it derives from a graph until the graph changes); plain views
    # (poasta_graph_view structs, e.g. hand-made test graphs) through the view entry point
    by_handle = len(graphs) > 0 and all(isinstance(g, POAGraph) for g in graphs)
    views = (_lib.GraphView * len(graphs))()
    for i, g in enumerate(graphs):
        views[i] = g.view() if isinstance(g, POAGraph) else g
    if graph_of_query is None:
        graph_of_query = list(range(n))
    qs = [bytes(q) for q in queries]
    qptr = (C.c_char_p * n)(*qs)
    qlens = (C.c_uint32 * n)(*[len(q) for q in qs])
    goq = (C.c_uint32 * n)(*graph_of_query)
    # one staging area for all alignments (capacity V + Q + 2 pairs each, as the header asks) and the result records filled
    # as one numpy array: per-problem Python work would dominate batches of short reads (config 5: 20,000 reads per call)
    qlens_np = np.fromiter((len(q) for q in qs), np.uint64, n)
    goq_np = np.asarray(graph_of_query, np.int64)
    nn = np.fromiter((v.n_nodes for v in views), np.uint64, len(graphs))
    caps = (nn[goq_np] + qlens_np + 2) if n else np.zeros(0, np.uint64)
    offs = np.zeros(n + 1, np.uint64)
    np.cumsum(caps, out=offs[1:])
    # the staging area is kept across calls (grow-only, per thread): a fresh multi-GB allocation would be page-faulted in anew by
    # every call, and the results below are copied out of it anyway
    need = max(1, int(offs[-1]))
    st = getattr(_staging, "bufs", None)
    if st is None or len(st[0]) < need:
        st = _staging.bufs = (np.empty(need + need // 8, np.uint32), np.empty(need + need // 8, np.uint32))
    rbuf, qbuf = st
    rs = np.zeros(n, _RESULT_DTYPE)
    rs["rpos"] = rbuf.ctypes.data + offs[:-1] * 4
    rs["qpos"] = qbuf.ctypes.data + offs[:-1] * 4
    rs["capacity"] = caps
    results = rs.ctypes.data_as(C.POINTER(_lib.Result))
    prm = _params(config, aln_type, tiebreak)
    t_call = time.perf_counter()
    if by_handle:
        handles = (C.c_void_p * len(graphs))(*[g._h if isinstance(g._h, int) else g._h.value for g in graphs])
        check(lib.poasta_align_batch_graphs(device, handles, len(graphs), qptr, qlens, goq, n, C.byref(prm), results),
              "poasta_align_batch_graphs")
    else:
        check(lib.poasta_align_batch(device, views, len(graphs), qptr, qlens, goq, n, C.byref(prm), results), "poasta_align_batch")
    t_done = time.perf_counter()
    bad = np.flatnonzero(rs["status"])
    if len(bad):
        check(int(rs["status"][bad[0]]), f"alignment {int(bad[0])}")
    out = []
    scores, npairs = rs["score"].tolist(), rs["n_pairs"].tolist()
    nq, nv, npr, nc = rs["num_queued"].tolist(), rs["num_visited"].tolist(), rs["num_pruned"].tolist(), rs["num_cells"].tolist()
    starts = offs[:-1].tolist()
    for i in range(n):
        a, k = starts[i], npairs[i]
        r, q = rbuf[a:a + k].copy(), qbuf[a:a + k].copy()
        aln = [AlignedPair(None if x == NONE else x, None if y == NONE else y) for x, y in zip(r.tolist(), q.tolist())]
        out.append(AstarResult(scores[i], aln, nq[i], nv[i], npr[i], nc[i], r, q))
    if os.environ.get("POASTA_TIMING"):
        sys.stderr.write(f"poasta_b200 align_batch (python): set-up {t_call - t_start:.3f} s, C call {t_done - t_call:.3f} s, "
                         f"results {time.perf_counter() - t_done:.3f} s for {n} problems\n")
    return out


def last_batch_stats():
    lib = _lib.load()
    ms, nl, cells = C.c_double(), C.c_uint32(), C.c_uint64()
    lib.poasta_last_batch_stats(C.byref(ms), C.byref(nl), C.byref(cells))
    return ms.value, nl.value, cells.value


def last_dp_kernel():
    """Which DP fill kernel the last DP-path call ran: 1 CTA, 2 warp (32-bit), 3 packed s16x2, 4 packed cluster (diagnostic)."""
    return int(_lib.load().poasta_last_dp_kernel())


def last_msa_path():
    """Pipeline of the last msa_batch: 1 reference-exact engine, 2 device-resident DP rounds, 3 per-round host pipeline (diagnostic)."""
    return int(_lib.load().poasta_last_msa_path())


def last_search_table():
    """Visited-table layout of the last search launch of align_batch: 1 dense, 2 blocked (diagnostic)."""
    return int(_lib.load().poasta_last_search_table())


def last_batch_fill_ms():
    """(milliseconds, launches) of the DP fill kernel alone in the last device-resident DP-only build."""
    ms, n = C.c_double(), C.c_uint32()
    _lib.load().poasta_last_batch_fill_ms(C.byref(ms), C.byref(n))
    return ms.value, n.value


def last_batch_search_ms():
    lib = _lib.load()
    ms = C.c_double()
    lib.poasta_last_batch_search_ms(C.byref(ms))
    return ms.value


def last_batch_device_ms():
    """Device-side critical path of the last call: the largest per-lane sum of kernel time (lanes overlap on the device)."""
    lib = _lib.load()
    ms = C.c_double()
    lib.poasta_last_batch_device_ms(C.byref(ms))
    return ms.value


def last_batch_search_counts():
    """(num_queued, num_visited, num_pruned) summed over the last msa_batch in reference-exact mode."""
    lib = _lib.load()
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.poasta_last_batch_search_counts(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def last_batch_io():
    lib = _lib.load()
    a, b = C.c_uint64(), C.c_uint64()
    lib.poasta_last_batch_io(C.byref(a), C.byref(b))
    return a.value, b.value


def measure_int_peak(device=0):
    """Measured integer issue peaks of the device in Tera lane-operations/s: (add, xor+min, fused add+min)."""
    lib = _lib.load()
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib.poasta_measure_int_peak(device, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


class PoastaAligner:
    def __init__(self, config=AffineMinGapCost(), aln_type=AlignmentType.Global, device=0, tiebreak=TIEBREAK_EXACT):
        self.config, self.aln_type, self.device, self.tiebreak = config, aln_type, device, tiebreak

    def align(self, graph, seq):
        return align_batch([graph], [seq], [0], self.config, self.aln_type, self.device, self.tiebreak)[0]


def msa_batch(fastas, config=AffineMinGapCost(), aln_type=AlignmentType.Global, device=0, graphs=None,
              tiebreak=TIEBREAK_EXACT):
    """``poasta align`` for many families at once (bin/poasta.rs:163-236): returns (graphs, scores per family, cells)."""
    lib = _lib.load()
    n = len(fastas)
    fa = [f if isinstance(f, bytes) else f.encode() for f in fastas]
    ptr = (C.c_char_p * n)(*fa)
    lens = (C.c_size_t * n)(*[len(f) for f in fa])
    nseq = [int(f.startswith(b">")) + f.count(b"\n>") for f in fa]  # '>' at line starts only, like the C side
    scores = np.zeros(max(1, sum(nseq)), np.uint32)
    cells = C.c_uint64()
    prm = _params(config, aln_type, tiebreak)
    u32p = C.POINTER(C.c_uint32)
    if graphs is None:
        handles = (C.c_void_p * n)()
        check(lib.poasta_msa_batch(device, ptr, lens, n, C.byref(prm), handles, scores.ctypes.data_as(u32p), C.byref(cells)),
              "poasta_msa_batch")
        graphs = [POAGraph(_handle=C.c_void_p(h)) for h in handles]
    else:
        handles = (C.c_void_p * n)(*[g._h for g in graphs])
        check(lib.poasta_msa_batch_into(device, handles, ptr, lens, n, C.byref(prm), scores.ctypes.data_as(u32p),
                                        C.byref(cells)), "poasta_msa_batch_into")
    per_family, k = [], 0
    for c in nseq:
        per_family.append(scores[k:k + c].copy())
        k += c
    return graphs, per_family, cells.value
