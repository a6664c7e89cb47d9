// Host-side packing of (graph, query) problems into the flat, rank-ordered layout the kernels read.
//
// HBM layout (one batch):
//   graphs[]      GraphDesc   per distinct graph: V, bases into the node / pred arrays
//   node_sym[]    u8          symbol per RANK (row r of the DP = node topo_order[r])
//   node_info[]   u32         per rank: successor summary used by openI_ok
//                               bits 0-7  the unique successor symbol (valid if bit 9 set and bit 8 clear)
//                               bit  8    "always open": an `end` successor or >= 2 distinct successor symbols
//                               bit  9    has at least one successor
//                               bit  10   row is the end node, bit 11 row is the start node
//   node_id[]     u32         rank -> node id (for the emitted rpos)
//   pred_off[]    u32         per graph V+1 offsets into pred_rank (absolute)
//   pred_rank[]   u32         predecessor RANKS in the reference's Incoming order (graphs/poa.rs:433-435)
//   queries[]     u8          query bytes, each padded to 4
//   alns[]        AlnDesc     per alignment: graph id, Q, bases into queries / trace / output
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/poasta_b200.h"

namespace poasta {

struct GraphDesc {
    uint32_t V;
    uint32_t node_base;     // into node_sym / node_info / node_id
    uint32_t predoff_base;  // into pred_off (V + 1 entries)
    uint32_t max_in_degree;
    uint32_t n_stored;      // rows that later rows read from the row store (warp kernel)
    uint32_t n_ord_rows;    // rows with >= 2 predecessors: they own a block of predecessor ordinals in the warp kernel's trace
    uint32_t pred_base;     // first entry of this graph in pred_rank / pred_slot
};

struct AlnDesc {
    uint32_t graph;
    uint32_t Q;
    uint32_t W;           // padded row width (multiple of 4, >= Q + 1)
    uint32_t query_base;  // into queries
    uint64_t trace_base;  // into the trace store (elements), V * W elements (CTA kernel: one word per cell)
    uint64_t ptrace_base; // into the trace store (32-bit words): V * 256 bit-plane words + n_ord_rows * 1024 ordinals (warp kernel)
    uint32_t out_base;    // into the pair output staging area, capacity V + Q + 2
    uint32_t out_cap;
    uint64_t prow_base;   // into the trace store (16-bit halves): V rows of ceil(W / 32) * 64 halves, M then D (packed kernel)
};

constexpr uint32_t INFO_ALWAYS_OPEN = 1u << 8;
constexpr uint32_t INFO_HAS_SUCC = 1u << 9;
constexpr uint32_t INFO_IS_END = 1u << 10;
constexpr uint32_t INFO_IS_START = 1u << 11;
constexpr uint32_t INFO_STORE = 1u << 12;  // some later row reads this row from the row store (warp kernel)
constexpr uint32_t INFO_SLOT_SHIFT = 13;   // bits 13.. : index of the row inside the compact row store

// Everything the warp kernel needs to start a row, in one 128-bit load per row (fetched one row ahead):
//   x  node_info of the row
//   y  first index of the row's predecessors in pred_rank / pred_slot (relative to the graph)
//   z  in-degree | node symbol << 24
//   w  for rows with exactly one predecessor: ROW_PREV_IN_REGS when that predecessor is the row just before (the common
//      case: nothing is read from memory), else its slot in the row store; unused otherwise
struct alignas(16) RowRec {
    uint32_t info, pred_begin, npred_sym, pred0_slot;
};
constexpr uint32_t ROW_PREV_IN_REGS = 0xFFFFFFFFu;

// One graph in the kernels' rank-ordered layout (see the header comment); pred_off is relative to this graph's pred_rank.
struct PackedGraph {
    uint32_t V = 0, max_in_degree = 0, n_stored = 0, n_ord_rows = 0;
    bool acgt_only = true;  // every real node carries A, C, G or T (the packed kernel's symbol tables)
    uint32_t lmin = 0;      // edges on the shortest start -> end path (bounds the optimal global score from above)
    std::vector<uint8_t> node_sym;
    std::vector<uint32_t> node_info, node_id, ord_slot, pred_off, pred_rank, pred_slot;
    std::vector<RowRec> rows;

    void build(const poasta_graph_view &g, std::vector<uint32_t> &rank_scratch) {
        V = g.n_nodes;
        rank_scratch.resize(V);
        for (uint32_t r = 0; r < V; ++r) rank_scratch[g.topo_order[r]] = r;
        node_sym.clear(), node_info.clear(), node_id.clear(), ord_slot.clear(), pred_off.clear(), pred_rank.clear(), pred_slot.clear();
        node_sym.reserve(V), node_info.reserve(V), node_id.reserve(V), ord_slot.reserve(V), pred_off.reserve(V + 1);
        pred_rank.reserve(g.pred_off[V]);
        max_in_degree = n_stored = n_ord_rows = 0;
        acgt_only = true;
        for (uint32_t r = 0; r < V; ++r) {
            const uint32_t v = g.topo_order[r];
            node_sym.push_back(g.symbol[v]);
            if (v != g.start_node && v != g.end_node) {
                const uint8_t c = g.symbol[v];
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') acgt_only = false;
            }
            node_id.push_back(v);
            ord_slot.push_back(g.pred_off[v + 1] - g.pred_off[v] >= 2 ? n_ord_rows++ : 0xFFFFFFFFu);
            uint32_t info = 0;
            bool has_end = false, have_sym = false, multi = false;
            uint8_t usym = 0;
            for (uint32_t k = g.succ_off[v]; k < g.succ_off[v + 1]; ++k) {
                const uint32_t s = g.succ[k];
                info |= INFO_HAS_SUCC;
                if (s == g.end_node) {
                    has_end = true;
                } else if (!have_sym) {
                    usym = g.symbol[s];
                    have_sym = true;
                } else if (g.symbol[s] != usym) {
                    multi = true;
                }
            }
            if (has_end || multi) info |= INFO_ALWAYS_OPEN;
            info |= usym;
            if (v == g.end_node) info |= INFO_IS_END;
            if (v == g.start_node) info |= INFO_IS_START;
            node_info.push_back(info);
            pred_off.push_back((uint32_t)pred_rank.size());
            for (uint32_t k = g.pred_off[v]; k < g.pred_off[v + 1]; ++k) pred_rank.push_back(rank_scratch[g.pred[k]]);
            max_in_degree = std::max(max_in_degree, g.pred_off[v + 1] - g.pred_off[v]);
        }
        pred_off.push_back((uint32_t)pred_rank.size());
        {
            std::vector<uint32_t> &dist = rank_scratch;  // the rank table is no longer needed
            dist.assign(V, 0xFFFFFFFFu);
            for (uint32_t r = 0; r < V; ++r) {
                if (node_info[r] & INFO_IS_START) dist[r] = 0;
                for (uint32_t k = pred_off[r]; k < pred_off[r + 1]; ++k)
                    if (dist[pred_rank[k]] != 0xFFFFFFFFu) dist[r] = std::min(dist[r], dist[pred_rank[k]] + 1);
                if (node_info[r] & INFO_IS_END) lmin = dist[r] == 0xFFFFFFFFu ? V : dist[r];
            }
        }
        // rows read through the row store: every predecessor that is not the row just before its successor (that one is
        // still in registers when the successor is computed)
        for (uint32_t r = 0; r < V; ++r)
            for (uint32_t k = pred_off[r]; k < pred_off[r + 1]; ++k)
                if (pred_rank[k] != r - 1) node_info[pred_rank[k]] |= INFO_STORE;
        for (uint32_t r = 0; r < V; ++r)
            if (node_info[r] & INFO_STORE) node_info[r] |= (n_stored++) << INFO_SLOT_SHIFT;
        pred_slot.resize(pred_rank.size());
        for (uint32_t k = 0; k < pred_rank.size(); ++k) {
            const uint32_t pi = node_info[pred_rank[k]];
            pred_slot[k] = (pi & INFO_STORE) ? (pi >> INFO_SLOT_SHIFT) : 0xFFFFFFFFu;
        }
        rows.resize(V);
        for (uint32_t r = 0; r < V; ++r) {
            const uint32_t np = pred_off[r + 1] - pred_off[r];
            RowRec &rr = rows[r];
            rr.info = node_info[r];
            rr.pred_begin = pred_off[r];
            rr.npred_sym = np | ((uint32_t)node_sym[r] << 24);
            rr.pred0_slot = np == 1 ? (pred_rank[pred_off[r]] + 1 == r ? ROW_PREV_IN_REGS : pred_slot[pred_off[r]]) : 0;
        }
    }
};

struct PackedBatch {
    std::vector<GraphDesc> graphs;
    std::vector<AlnDesc> alns;
    std::vector<uint8_t> node_sym;
    std::vector<uint32_t> node_info, node_id, ord_slot, pred_off, pred_rank, pred_slot;
    std::vector<RowRec> rows;  // pred_begin is relative to the graph's first predecessor entry (GraphDesc::pred_base)
    std::vector<uint8_t> queries;
    uint64_t trace_elems = 0;
    uint64_t ptrace_words = 0;
    uint64_t prow_halves = 0;  // packed kernel's M / D row store
    uint32_t max_n_stored = 0; // max over alignments of the graph's n_stored
    uint32_t max_V = 0, max_Q = 0, max_lmin = 0;
    bool acgt_only = true;
    uint64_t out_elems = 0;
    uint64_t max_cells = 0;  // max over alignments of V * W
    uint64_t max_store_cells = 0;  // max over alignments of n_stored * W (warp kernel's compact row store)
    uint32_t max_W = 0;
    uint64_t total_cells = 0;  // sum of V_real * Q (the GCUPS numerator)
    uint32_t max_in_degree = 0;

    void clear() {
        graphs.clear();
        alns.clear();
        node_sym.clear();
        node_info.clear();
        node_id.clear();
        ord_slot.clear();
        pred_off.clear();
        pred_rank.clear();
        pred_slot.clear();
        rows.clear();
        queries.clear();
        trace_elems = ptrace_words = prow_halves = out_elems = max_cells = total_cells = max_store_cells = 0;
        max_in_degree = max_n_stored = max_V = max_Q = max_lmin = 0;
        acgt_only = true;
        max_W = 0;
    }

    // Appends one graph; returns its id. `rank_scratch` avoids reallocation across calls.
    uint32_t add_graph(const poasta_graph_view &g, std::vector<uint32_t> &rank_scratch) {
        PackedGraph pg;
        pg.build(g, rank_scratch);
        return add_packed(pg);
    }

    // Appends a graph packed earlier (pack once per graph on a host worker thread, concatenate here).
    uint32_t add_packed(const PackedGraph &pg) {
        GraphDesc gd;
        gd.V = pg.V;
        gd.node_base = (uint32_t)node_sym.size();
        gd.predoff_base = (uint32_t)pred_off.size();
        gd.max_in_degree = pg.max_in_degree;
        gd.n_stored = pg.n_stored;
        gd.n_ord_rows = pg.n_ord_rows;
        gd.pred_base = (uint32_t)pred_rank.size();
        rows.insert(rows.end(), pg.rows.begin(), pg.rows.end());
        node_sym.insert(node_sym.end(), pg.node_sym.begin(), pg.node_sym.end());
        node_info.insert(node_info.end(), pg.node_info.begin(), pg.node_info.end());
        node_id.insert(node_id.end(), pg.node_id.begin(), pg.node_id.end());
        ord_slot.insert(ord_slot.end(), pg.ord_slot.begin(), pg.ord_slot.end());
        const uint32_t base = (uint32_t)pred_rank.size();
        const size_t o0 = pred_off.size();
        pred_off.insert(pred_off.end(), pg.pred_off.begin(), pg.pred_off.end());
        for (size_t i = o0; i < pred_off.size(); ++i) pred_off[i] += base;
        pred_rank.insert(pred_rank.end(), pg.pred_rank.begin(), pg.pred_rank.end());
        pred_slot.insert(pred_slot.end(), pg.pred_slot.begin(), pg.pred_slot.end());
        max_in_degree = std::max(max_in_degree, gd.max_in_degree);
        acgt_only = acgt_only && pg.acgt_only;
        max_lmin = std::max(max_lmin, pg.lmin);
        graphs.push_back(gd);
        return (uint32_t)graphs.size() - 1;
    }

    void add_alignment(uint32_t graph, const uint8_t *q, uint32_t Q) {
        AlnDesc a;
        a.graph = graph;
        a.Q = Q;
        a.W = (Q + 1 + 7) & ~7u;
        a.query_base = (uint32_t)queries.size();
        queries.insert(queries.end(), q, q + Q);
        while (queries.size() & 3) queries.push_back(0);
        const uint64_t V = graphs[graph].V;
        a.trace_base = trace_elems;
        trace_elems += V * a.W;
        a.ptrace_base = ptrace_words;
        ptrace_words += V * 256 + (uint64_t)graphs[graph].n_ord_rows * 1024;
        a.prow_base = prow_halves;
        prow_halves += V * (uint64_t)((a.W + 31) / 32) * 64;
        max_n_stored = std::max(max_n_stored, graphs[graph].n_stored);
        max_V = std::max<uint32_t>(max_V, (uint32_t)V), max_Q = std::max(max_Q, Q);
        a.out_cap = (uint32_t)V + Q + 2;
        a.out_base = (uint32_t)out_elems;
        out_elems += a.out_cap;
        max_cells = std::max<uint64_t>(max_cells, V * a.W);
        max_store_cells = std::max<uint64_t>(max_store_cells, (uint64_t)std::max<uint32_t>(1, graphs[graph].n_stored) * a.W);
        max_W = std::max(max_W, a.W);
        total_cells += (V - 2) * Q;
        alns.push_back(a);
    }
};

}  // namespace poasta
