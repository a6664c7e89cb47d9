"""ctypes loader for the C-ABI shared library (include/poasta_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m poasta_b200.build``.
There is no fallback: if the shared object (with its sm_100a kernels) is missing, importing the
alignment entry points raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POASTA_B200_LIB", os.path.join(_HERE, "libposta_b200.so"))


class GraphView(C.Structure):
    """``poasta_graph_view``: the AlignableRefGraph trait flattened (graphs/mod.rs:23-53)."""
    _fields_ = [
        ("n_nodes", C.c_uint32), ("start_node", C.c_uint32), ("end_node", C.c_uint32),
        ("topo_order", C.POINTER(C.c_uint32)),
        ("pred_off", C.POINTER(C.c_uint32)), ("pred", C.POINTER(C.c_uint32)),
        ("succ_off", C.POINTER(C.c_uint32)), ("succ", C.POINTER(C.c_uint32)),
        ("symbol", C.POINTER(C.c_uint8)),
    ]


class Params(C.Structure):
    """``poasta_params``: GapAffine costs + AlignmentType + heuristic."""
    _fields_ = [("mismatch", C.c_uint8), ("gap_open", C.c_uint8), ("gap_extend", C.c_uint8),
                ("mode", C.c_uint8), ("heuristic", C.c_uint8), ("tiebreak", C.c_uint8)]


class Result(C.Structure):
    """``poasta_result``: AstarResult (aligner/astar.rs:81-90)."""
    _fields_ = [("score", C.c_uint32), ("n_pairs", C.c_uint32),
                ("rpos", C.POINTER(C.c_uint32)), ("qpos", C.POINTER(C.c_uint32)),
                ("capacity", C.c_uint32), ("status", C.c_int32), ("num_cells", C.c_uint64),
                ("num_queued", C.c_uint64), ("num_visited", C.c_uint64), ("num_pruned", C.c_uint64)]


# every symbol include/poasta_b200.h declares
EXPORTS = [
    "poasta_align_batch", "poasta_align_batch_graphs", "poasta_last_batch_stats", "poasta_last_dp_kernel", "poasta_last_msa_path", "poasta_last_search_table", "poasta_last_batch_fill_ms", "poasta_last_batch_io", "poasta_last_batch_search_counts", "poasta_last_batch_search_ms", "poasta_last_batch_device_ms", "poasta_measure_int_peak",
    "poasta_graph_new", "poasta_graph_free", "poasta_graph_free_many", "poasta_graph_node_count",
    "poasta_graph_node_count_with_start_and_end", "poasta_graph_edge_count",
    "poasta_graph_num_sequences", "poasta_graph_add_alignment", "poasta_graph_get_view",
    "poasta_graph_aligned_nodes", "poasta_graph_to_fasta", "poasta_graph_to_gfa", "poasta_graph_to_dot",
    "poasta_graph_save", "poasta_graph_load", "poasta_graph_from_fasta_msa",
    "poasta_free_buffer", "poasta_msa_batch", "poasta_msa_batch_into", "poasta_version",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m poasta_b200.build` "
            "(nvcc, sm_100a). poasta_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    lib.poasta_graph_new.restype = vp
    lib.poasta_graph_free.argtypes = [vp]
    lib.poasta_graph_free_many.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32]
    lib.poasta_graph_free_many.restype = None
    for f in ("poasta_graph_node_count", "poasta_graph_node_count_with_start_and_end",
              "poasta_graph_edge_count", "poasta_graph_num_sequences"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = C.c_uint32
    lib.poasta_graph_add_alignment.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint32, u32p, u32p,
                                               C.c_uint32, C.POINTER(C.c_uint64), C.c_uint32]
    lib.poasta_graph_get_view.argtypes = [vp, C.POINTER(GraphView)]
    lib.poasta_graph_aligned_nodes.argtypes = [vp, C.c_uint32, u32p, C.c_uint32]
    lib.poasta_graph_aligned_nodes.restype = C.c_uint32
    for f in ("poasta_graph_to_fasta", "poasta_graph_save"):
        getattr(lib, f).argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.poasta_graph_to_gfa.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.poasta_graph_to_dot.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.poasta_graph_load.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    lib.poasta_graph_from_fasta_msa.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    lib.poasta_free_buffer.argtypes = [vp]
    lib.poasta_version.restype = C.c_char_p
    lib.poasta_align_batch.argtypes = [C.c_int, C.POINTER(GraphView), C.c_uint32, C.POINTER(C.c_char_p),
                                       u32p, u32p, C.c_uint32, C.POINTER(Params), C.POINTER(Result)]
    lib.poasta_align_batch_graphs.argtypes = [C.c_int, C.POINTER(vp), C.c_uint32, C.POINTER(C.c_char_p),
                                              u32p, u32p, C.c_uint32, C.POINTER(Params), C.POINTER(Result)]
    lib.poasta_last_batch_stats.argtypes = [C.POINTER(C.c_double), u32p, C.POINTER(C.c_uint64)]
    lib.poasta_last_batch_search_ms.argtypes = [C.POINTER(C.c_double)]
    lib.poasta_last_batch_io.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.poasta_last_batch_device_ms.argtypes = [C.POINTER(C.c_double)]
    lib.poasta_last_batch_search_counts.argtypes = [C.POINTER(C.c_uint64)] * 3
    lib.poasta_measure_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.poasta_msa_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_uint32,
                                     C.POINTER(Params), C.POINTER(vp), u32p, C.POINTER(C.c_uint64)]
    lib.poasta_msa_batch_into.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                          C.c_uint32, C.POINTER(Params), u32p, C.POINTER(C.c_uint64)]
    _lib = lib
    return lib
