// POAGraph -> CSR view (`poasta_graph_view`, include/poasta_b200.h).
// This is the `AlignableRefGraph` trait (/root/reference/src/graphs/mod.rs:23-53) materialised:
// predecessors()/successors() in the reference's iteration order, symbols, topological order.
#pragma once
#include "../../include/poasta_b200.h"
#include "poa_graph.hpp"

namespace poasta {

struct FlatGraph {
    std::vector<uint32_t> topo_order, pred_off, pred, succ_off, succ;
    std::vector<uint8_t> symbol;
    uint32_t n_nodes = 0, start_node = 0, end_node = 1;

    void build(const POAGraph &g) {
        n_nodes = g.node_count_with_start_and_end();
        start_node = g.start_node;
        end_node = g.end_node;
        topo_order = g.topological_sorted;
        pred_off.assign(n_nodes + 1, 0);
        succ_off.assign(n_nodes + 1, 0);
        pred.clear();
        succ.clear();
        symbol.resize(n_nodes);
        for (uint32_t v = 0; v < n_nodes; ++v) {
            symbol[v] = g.nodes[v].symbol;
            g.for_predecessors(v, [&](uint32_t p) { pred.push_back(p); });
            g.for_successors(v, [&](uint32_t s) { succ.push_back(s); });
            pred_off[v + 1] = (uint32_t)pred.size();
            succ_off[v + 1] = (uint32_t)succ.size();
        }
    }

    poasta_graph_view view() const {
        poasta_graph_view v;
        v.n_nodes = n_nodes;
        v.start_node = start_node;
        v.end_node = end_node;
        v.topo_order = topo_order.data();
        v.pred_off = pred_off.data();
        v.pred = pred.data();
        v.succ_off = succ_off.data();
        v.succ = succ.data();
        v.symbol = symbol.data();
        return v;
    }
};

}  // namespace poasta
