"""``POAGraph`` — Python mirror of /root/reference/src/graphs/poa.rs over the C-ABI handle."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check

NONE = 0xFFFFFFFF


class AlignedPair:
    """aligner/alignment.rs:3-40. ``rpos``/``qpos`` are ``None`` for gaps."""
    __slots__ = ("rpos", "qpos")

    def __init__(self, rpos, qpos):
        self.rpos, self.qpos = rpos, qpos

    def is_aligned(self):
        return self.rpos is not None and self.qpos is not None

    def __eq__(self, o):
        return (self.rpos, self.qpos) == (o.rpos, o.qpos)

    def __repr__(self):
        return f"AlignedPair({self.rpos}, {self.qpos})"


def _pairs_to_arrays(alignment):
    n = len(alignment)
    r = np.fromiter((NONE if p.rpos is None else p.rpos for p in alignment), dtype=np.uint32, count=n)
    q = np.fromiter((NONE if p.qpos is None else p.qpos for p in alignment), dtype=np.uint32, count=n)
    return r, q


class POAGraph:
    def __init__(self, _handle=None):
        self._lib = _lib.load()
        self._h = _handle if _handle is not None else self._lib.poasta_graph_new()
        self._view_keepalive = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.poasta_graph_free(self._h)
            self._h = None

    @staticmethod
    def free_many(graphs, threads=0):
        """Frees the graphs of a batch on several host threads (``poasta_graph_free_many``); the objects become unusable."""
        hs = [g for g in graphs if getattr(g, "_h", None)]
        if not hs:
            return
        arr = (C.c_void_p * len(hs))(*[g._h if isinstance(g._h, int) else g._h.value for g in hs])
        for g in hs:
            g._h = None
        _lib.load().poasta_graph_free_many(arr, len(hs), threads)

    # -- graphs/poa.rs:114, 399-420 --
    def is_empty(self):
        return self.node_count() == 0

    def node_count(self):
        return self._lib.poasta_graph_node_count(self._h)

    def node_count_with_start_and_end(self):
        return self._lib.poasta_graph_node_count_with_start_and_end(self._h)

    def edge_count(self):
        return self._lib.poasta_graph_edge_count(self._h)

    def num_sequences(self):
        return self._lib.poasta_graph_num_sequences(self._h)

    # -- graphs/poa.rs:171-321 --
    def add_alignment_with_weights(self, name, seq, alignment=None, weights=None):
        seq = bytes(seq)
        w, wn = None, 0
        if weights is not None:
            w, wn = (C.c_uint64 * max(1, len(weights)))(*weights), len(weights)  # the C side checks the length (poa.rs:178-180)
        if alignment is None:
            rc = self._lib.poasta_graph_add_alignment(self._h, name.encode(), seq, len(seq), None, None, 0, w, wn)
        else:
            r, q = _pairs_to_arrays(alignment) if not isinstance(alignment, tuple) else alignment
            r = np.ascontiguousarray(r, dtype=np.uint32)
            q = np.ascontiguousarray(q, dtype=np.uint32)
            u32p = C.POINTER(C.c_uint32)
            # a zero-length alignment is Some(empty), not None: pass non-null pointers
            if len(r) == 0:
                r = np.zeros(1, np.uint32)
                q = np.zeros(1, np.uint32)
                n = 0
            else:
                n = len(r)
            rc = self._lib.poasta_graph_add_alignment(self._h, name.encode(), seq, len(seq),
                                                      r.ctypes.data_as(u32p), q.ctypes.data_as(u32p), n, w, wn)
        check(rc, "add_alignment_with_weights")

    def get_aligned_nodes(self, node):
        n = self._lib.poasta_graph_aligned_nodes(self._h, node, None, 0)  # the call returns the full count
        buf = (C.c_uint32 * max(1, n))()
        self._lib.poasta_graph_aligned_nodes(self._h, node, buf, n)
        return list(buf[:n])

    # -- flattened AlignableRefGraph --
    def view(self):
        v = _lib.GraphView()
        check(self._lib.poasta_graph_get_view(self._h, C.byref(v)), "graph_get_view")
        return v

    def view_arrays(self):
        """Copy of the CSR view as numpy arrays (topo_order, pred_off, pred, succ_off, succ, symbol)."""
        v = self.view()
        n = v.n_nodes
        as_np = lambda p, k, dt: np.ctypeslib.as_array(p, shape=(k,)).astype(dt, copy=True) if k else np.zeros(0, dt)
        pred_off = as_np(v.pred_off, n + 1, np.uint32)
        succ_off = as_np(v.succ_off, n + 1, np.uint32)
        return dict(n_nodes=n, start_node=v.start_node, end_node=v.end_node,
                    topo_order=as_np(v.topo_order, n, np.uint32), pred_off=pred_off,
                    pred=as_np(v.pred, int(pred_off[-1]), np.uint32), succ_off=succ_off,
                    succ=as_np(v.succ, int(succ_off[-1]), np.uint32), symbol=as_np(v.symbol, n, np.uint8))

    # -- writers (io/fasta.rs:69, io/graph.rs:24,245,374) --
    def _emit(self, fn, *args):
        out, n = C.c_void_p(), C.c_size_t()
        check(fn(self._h, *args, C.byref(out), C.byref(n)), fn.__name__)
        data = C.string_at(out, n.value)
        self._lib.poasta_free_buffer(out)
        return data

    def to_fasta(self):
        return self._emit(self._lib.poasta_graph_to_fasta)

    def to_gfa(self, v1=False):
        return self._emit(self._lib.poasta_graph_to_gfa, 1 if v1 else 0)

    def to_dot(self, petgraph_dialect=False):
        return self._emit(self._lib.poasta_graph_to_dot, 1 if petgraph_dialect else 0)

    def save(self):
        return self._emit(self._lib.poasta_graph_save)

    @classmethod
    def load(cls, data):
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.poasta_graph_load(data, len(data), C.byref(h)), "load_graph")
        return cls(_handle=h)

    @classmethod
    def from_fasta_msa(cls, text):
        lib = _lib.load()
        if isinstance(text, str):
            text = text.encode()
        h = C.c_void_p()
        check(lib.poasta_graph_from_fasta_msa(text, len(text), C.byref(h)), "load_graph_from_fasta_msa")
        return cls(_handle=h)
