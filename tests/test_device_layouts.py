"""The host-side packing the device kernels read (poasta_b200/csrc/device_batch.hpp: rank-ordered rows, row store
slots, RowRec; search_index.hpp: NodeRec), re-derived here from the graph view in plain Python. Runs without a GPU;
the -m gpu tests check the kernels that consume these arrays."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from poasta_b200 import POAGraph, synth

ALWAYS_OPEN, HAS_SUCC, IS_END, IS_START, STORE, SLOT_SHIFT = 1 << 8, 1 << 9, 1 << 10, 1 << 11, 1 << 12, 13
PREV_IN_REGS = 0xFFFFFFFF


def _dump(view):
    cv = H._cview(view)
    V = cv.n_nodes
    E = cv.pred_off[V]
    u32 = lambda n: np.zeros(max(1, n), np.uint32)  # noqa: E731
    meta, rows, info, nid, oslot, poff, prank, pslot = u32(4), u32(4 * V), u32(V), u32(V), u32(V), u32(V + 1), u32(E), u32(E)
    sym = np.zeros(V, np.uint8)
    recs = u32(8 * V)
    p = lambda a: a.ctypes.data_as(H.u32p)  # noqa: E731
    rc = H.oracle().packed_graph_dump(C.byref(cv), p(meta), p(rows), p(info), p(nid), p(oslot), p(poff), p(prank), p(pslot),
                                      sym.ctypes.data_as(C.POINTER(C.c_uint8)), p(recs))
    assert rc == 0
    return cv, meta, rows.reshape(V, 4), info, nid, oslot, poff, prank[:E], pslot[:E], sym, recs.reshape(V, 8)


def _graphs():
    yield H.test_graph1()
    yield H.test_graph2()
    for fam, n in ((0, 6), (3, 9)):
        g = POAGraph()
        for name, seq in synth.parse_fasta(synth.config2_family(fam, n_seqs=n, root_len=150, rate=0.08)):
            if g.is_empty():
                g.add_alignment_with_weights(name, seq, None)
                continue
            st, sc, pairs, _ = H.oracle_align(g, seq)
            assert st == 0
            g.add_alignment_with_weights(name, seq, pairs)
            yield g  # every intermediate graph of the progressive build


def test_rank_ordered_rows_and_row_store():
    n_graphs = 0
    for g in _graphs():
        cv, meta, rows, info, nid, oslot, poff, prank, pslot, sym, _ = _dump(g)
        V = cv.n_nodes
        topo = [cv.topo_order[r] for r in range(V)]
        rank = {v: r for r, v in enumerate(topo)}
        preds = lambda v: [cv.pred[k] for k in range(cv.pred_off[v], cv.pred_off[v + 1])]  # noqa: E731
        succs = lambda v: [cv.succ[k] for k in range(cv.succ_off[v], cv.succ_off[v + 1])]  # noqa: E731
        assert meta[0] == V and meta[1] == max(len(preds(v)) for v in range(V))
        # which rows some later row reads from memory: every predecessor that is not the row just before its successor
        stored = set()
        for r, v in enumerate(topo):
            for pnode in preds(v):
                if rank[pnode] != r - 1:
                    stored.add(rank[pnode])
        slot_of = {r: i for i, r in enumerate(sorted(stored))}
        assert meta[2] == len(stored)
        n_ord = 0
        for r, v in enumerate(topo):
            assert nid[r] == v and sym[r] == cv.symbol[v]
            pr = preds(v)
            # predecessor ranks in the reference's Incoming order (graphs/poa.rs:433-435)
            assert list(prank[poff[r]:poff[r + 1]]) == [rank[x] for x in pr]
            assert list(pslot[poff[r]:poff[r + 1]]) == [slot_of.get(rank[x], 0xFFFFFFFF) for x in pr]
            assert oslot[r] == (n_ord if len(pr) >= 2 else 0xFFFFFFFF)
            n_ord += len(pr) >= 2
            # openI_ok summary: an `end` successor or two different successor symbols always allow opening
            sc = succs(v)
            syms = [cv.symbol[s] for s in sc if s != cv.end_node]
            always = (cv.end_node in sc) or len(set(syms)) >= 2
            want = (HAS_SUCC if sc else 0) | (ALWAYS_OPEN if always else 0) | (syms[0] if syms else 0)
            want |= (IS_END if v == cv.end_node else 0) | (IS_START if v == cv.start_node else 0)
            if r in stored:
                want |= STORE | (slot_of[r] << SLOT_SHIFT)
            assert info[r] == want, (r, v, hex(info[r]), hex(want))
            # the one record the warp kernel loads per row
            assert rows[r][0] == info[r] and rows[r][1] == poff[r]
            assert rows[r][2] == (len(pr) | (int(cv.symbol[v]) << 24))
            if len(pr) == 1:
                assert rows[r][3] == (PREV_IN_REGS if rank[pr[0]] == r - 1 else slot_of[rank[pr[0]]])
        assert meta[3] == n_ord
        n_graphs += 1
    assert n_graphs >= 10


def test_search_node_records():
    for g in _graphs():
        cv, *_, recs = _dump(g)
        V = cv.n_nodes
        # distances to `end` by plain dynamic programming over the DAG
        topo = [cv.topo_order[r] for r in range(V)]
        dmin, dmax = {}, {}
        for v in reversed(topo):
            sc = [cv.succ[k] for k in range(cv.succ_off[v], cv.succ_off[v + 1])]
            dmin[v] = 0 if v == cv.end_node else min(dmin[s] + 1 for s in sc)
            dmax[v] = 0 if not sc else max(dmax[s] + 1 for s in sc)
        for v in range(V):
            succ_off, bub_off, counts, symflags, dist_min, dist_max, pred_off, n_pred = recs[v]
            assert succ_off == cv.succ_off[v] and (counts & 0xFFFF) == cv.succ_off[v + 1] - cv.succ_off[v]
            assert (symflags & 0xFF) == cv.symbol[v]
            assert pred_off == cv.pred_off[v] and n_pred == cv.pred_off[v + 1] - cv.pred_off[v]
            assert (dist_min, dist_max) == (dmin[v], dmax[v]), v


def test_product_search_index_equals_oracle_bubble_index():
    """search_index.hpp (product, host side of the search kernel) against the oracle's BubbleIndex: bubble lists per node in
    the same order, distances to `end`, exit flags derived from the oracle's superbubbles."""
    n_graphs = 0
    for g in _graphs():
        cv = H._cview(g)
        V, M = cv.n_nodes, 16
        mk = lambda *shape: np.zeros(shape, np.uint32)  # noqa: E731
        c1, m1, d1 = mk(V), mk(V, M, 3), mk(V, 2)
        c2, m2, d2 = mk(V), mk(V, M, 3), mk(V, 2)
        ex = np.zeros(V, np.uint8)
        p = lambda a: a.ctypes.data_as(H.u32p)  # noqa: E731
        H.oracle().oracle_bubble_index(C.byref(cv), 0, M, p(c1), p(m1), p(d1))
        H.oracle().product_bubble_index(C.byref(cv), M, p(c2), p(m2), p(d2), ex.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert c1.max() <= M
        assert (c1 == c2).all() and (m1 == m2).all() and (d1 == d2).all()
        pairs = np.zeros(2 * V, np.uint32)
        nb = H.oracle().oracle_superbubbles(C.byref(cv), 0, p(pairs), V)
        exits = {int(pairs[2 * i + 1]) for i in range(nb)}
        assert {v for v in range(V) if ex[v]} == exits
        n_graphs += 1
    assert n_graphs >= 10
