// Per-round packing of a device-resident graph (dev_graph.hpp) into the rank-ordered layout the DP fill / traceback kernels
// read (device_batch.hpp: RowRec, rank -> node id, predecessor ranks in the reference's Incoming order), done ON THE DEVICE by
// the family's warp right before the round's fill, from the graph the previous round's insertion left in HBM. It is the device
// counterpart of PackedGraph::build for the DP-only mode (tiebreak = 1) of progressive builds: with it a round needs no host
// work and no copies (msa_dp_device.cuh).
//   adjacency in petgraph order        /root/reference/src/graphs/poa.rs:433-449
//   reverse post-order from `start`    /root/reference/src/graphs/tools.rs:5-37 (any topological order gives the same DP
//                                      tables; this one puts `start` at rank 0 and `end` at rank V - 1, which the walker assumes)
// Only what the kernels read when every row is kept in the trace store (WRITE_ROWS): no row-store slots, no ordinal blocks.
// The same source compiles for the host, where tests/test_dev_graph.py compares it with PackedGraph::build field by field.
#pragma once
#include "dev_graph.hpp"
#include "device_batch.hpp"

namespace poasta {
namespace dg {

// Scratch of one packing worker (warp); capacities in elements.
struct DpPackCaps {
    uint32_t V, E;
};
struct DpPackLayout {
    uint64_t succ_off, pred_off, succ, pred, order, rank, dfs, seen, total;
    DG_HD explicit DpPackLayout(const DpPackCaps &c) {
        uint64_t o = 0;
        auto take = [&](uint64_t bytes) {
            const uint64_t at = o;
            o += (bytes + 31) & ~(uint64_t)31;
            return at;
        };
        const uint64_t V = c.V, E = c.E;
        succ_off = take((V + 1) * 4), pred_off = take((V + 1) * 4);
        succ = take(E * 4), pred = take(E * 4);
        order = take(V * 4), rank = take(V * 4), dfs = take(V * 8), seen = take(V);
        total = o;
    }
};
struct DpPackScratch {
    DpPackCaps cap;
    uint32_t *succ_off, *pred_off, *succ, *pred, *order, *rank, *dfs;
    uint8_t *seen;
    DG_HD DpPackScratch(uint8_t *base, const DpPackCaps &c) : cap(c) {
        const DpPackLayout L(c);
        succ_off = (uint32_t *)(base + L.succ_off), pred_off = (uint32_t *)(base + L.pred_off);
        succ = (uint32_t *)(base + L.succ), pred = (uint32_t *)(base + L.pred);
        order = (uint32_t *)(base + L.order), rank = (uint32_t *)(base + L.rank), dfs = (uint32_t *)(base + L.dfs);
        seen = base + L.seen;
    }
};

// Where one graph's packed arrays go (the family's region of the round's batch arrays).
struct DpPackOut {
    RowRec *rows;         // V
    uint32_t *node_id;    // V: rank -> node id
    uint8_t *node_sym;    // V: symbol per rank
    uint32_t *pred_off;   // V + 1: offsets into the batch's pred_rank array, ABSOLUTE (pred_base added)
    uint32_t *pred_rank;  // E: predecessor ranks, Incoming order
    uint32_t pred_base;   // index of pred_rank[0] in the batch array
};

// Returns DG_OK, DG_CAPACITY (scratch too small) or DG_INVALID_ALIGNMENT (a node is not reachable from start: cannot happen for
// graphs built by add_alignment + post_process). Warp-uniform result. `max_in_degree` is written by lane 0 only.
template <class Sync>
DG_HD int dp_pack(const Store &g, DpPackScratch &w, const DpPackOut &o, uint32_t *max_in_degree, uint32_t lane, uint32_t n_lanes, Sync &&sync) {
    const uint32_t V = g.f->n_nodes;
    if (V > w.cap.V || g.f->live_edges > w.cap.E) return DG_CAPACITY;
    // ---- CSR in adjacency (newest-edge-first) order, node-id space ----
    for (uint32_t n = lane; n < V; n += n_lanes) {
        uint32_t a = 0, b = 0;
        for (uint32_t e = g.head_out[n]; e != NIL; e = g.e_next_out[e]) ++a;
        for (uint32_t e = g.head_in[n]; e != NIL; e = g.e_next_in[e]) ++b;
        w.succ_off[n + 1] = a, w.pred_off[n + 1] = b;
        w.seen[n] = 0;
    }
    sync();
    if (lane == 0) {
        w.succ_off[0] = w.pred_off[0] = 0;
        for (uint32_t n = 0; n < V; ++n) w.succ_off[n + 1] += w.succ_off[n], w.pred_off[n + 1] += w.pred_off[n];
    }
    sync();
    for (uint32_t n = lane; n < V; n += n_lanes) {
        uint32_t k = w.succ_off[n];
        for (uint32_t e = g.head_out[n]; e != NIL; e = g.e_next_out[e]) w.succ[k++] = g.e_dst[e];
        k = w.pred_off[n];
        for (uint32_t e = g.head_in[n]; e != NIL; e = g.e_next_in[e]) w.pred[k++] = g.e_src[e];
    }
    sync();
    // ---- reverse post-order of the DFS from start, children in adjacency order, marked when pushed (lane 0) ----
    int rc = DG_OK;
    if (lane == 0) {
        uint32_t n_order = 0, depth = 1;
        w.dfs[0] = Store::START, w.dfs[1] = w.succ_off[Store::START];
        w.seen[Store::START] = 1;
        while (depth) {
            uint32_t *top = w.dfs + 2 * (size_t)(depth - 1);
            const uint32_t end = w.succ_off[top[0] + 1];
            uint32_t child = NIL;
            while (top[1] != end) {
                const uint32_t c = w.succ[top[1]++];
                if (!w.seen[c]) {
                    child = c;
                    break;
                }
            }
            if (child != NIL) {
                w.seen[child] = 1;
                w.dfs[2 * (size_t)depth] = child, w.dfs[2 * (size_t)depth + 1] = w.succ_off[child];
                ++depth;
            } else {
                w.order[V - 1 - n_order] = top[0];  // post-order position n_order = rank V - 1 - n_order (n_order < V: each node once)
                ++n_order;
                --depth;
            }
        }
        if (n_order != V || w.order[0] != Store::START || w.order[V - 1] != Store::END) rc = DG_INVALID_ALIGNMENT;
    }
    rc = sync.bcast(rc);
    if (rc != DG_OK) return rc;
    for (uint32_t r = lane; r < V; r += n_lanes) w.rank[w.order[r]] = r;
    sync();
    // ---- per rank: symbol, successor summary (openI_ok), in-degree; then the rank-space predecessor CSR ----
    uint32_t my_max = 0;
    for (uint32_t r = lane; r < V; r += n_lanes) {
        const uint32_t v = w.order[r];
        uint32_t info = 0;
        bool has_end = false, have_sym = false, multi = false;
        uint8_t usym = 0;
        for (uint32_t k = w.succ_off[v]; k < w.succ_off[v + 1]; ++k) {
            const uint32_t s = w.succ[k];
            info |= INFO_HAS_SUCC;
            if (s == Store::END) {
                has_end = true;
            } else if (!have_sym) {
                usym = g.sym[s];
                have_sym = true;
            } else if (g.sym[s] != usym) {
                multi = true;
            }
        }
        if (has_end || multi) info |= INFO_ALWAYS_OPEN;
        info |= usym;
        if (v == Store::END) info |= INFO_IS_END;
        if (v == Store::START) info |= INFO_IS_START;
        const uint32_t np = w.pred_off[v + 1] - w.pred_off[v];
        my_max = np > my_max ? np : my_max;
        o.node_id[r] = v;
        o.node_sym[r] = g.sym[v];
        o.rows[r].info = info;
        o.rows[r].npred_sym = np | ((uint32_t)g.sym[v] << 24);
        o.pred_off[r + 1] = np;  // per-rank in-degree, scanned below
    }
    sync();
    if (lane == 0) {
        o.pred_off[0] = o.pred_base;
        for (uint32_t r = 0; r < V; ++r) o.pred_off[r + 1] += o.pred_off[r];
    }
    sync();
    for (uint32_t r = lane; r < V; r += n_lanes) {
        const uint32_t v = w.order[r];
        const uint32_t rel = o.pred_off[r] - o.pred_base;
        uint32_t k2 = rel;
        for (uint32_t k = w.pred_off[v]; k < w.pred_off[v + 1]; ++k) o.pred_rank[k2++] = w.rank[w.pred[k]];
        const uint32_t np = k2 - rel;
        o.rows[r].pred_begin = rel;
        o.rows[r].pred0_slot = (np == 1 && o.pred_rank[rel] + 1 == r) ? ROW_PREV_IN_REGS : 0u;
    }
    // largest in-degree over the lanes (the host build's PackedGraph::max_in_degree)
#if defined(__CUDA_ARCH__)
    for (int d = 16; d > 0; d >>= 1) {
        const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, my_max, d);
        my_max = other > my_max ? other : my_max;
    }
#endif
    if (lane == 0 && max_in_degree) *max_in_degree = my_max;
    sync();
    return DG_OK;
}

}  // namespace dg
}  // namespace poasta
