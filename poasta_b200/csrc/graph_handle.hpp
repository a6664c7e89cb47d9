// The opaque `poasta_graph` of include/poasta_b200.h: a POAGraph<u32> plus its cached CSR view.
#pragma once
#include "device_batch.hpp"
#include "flatten.hpp"
#include "poa_io.hpp"
#include "search_index.hpp"

struct poasta_graph {
    poasta::POAGraph g;
    poasta::FlatGraph flat;
    bool flat_valid = false;
    poasta::SearchIndex sidx;  // bubbles / distances / sampled paths of the current graph (order-exact search)
    int sidx_state = 0;        // 0 = stale, 1 = built without paths, 2 = built with paths
    poasta::PackedGraph packed;  // rank-ordered layout of the current graph for the DP kernels
    bool packed_valid = false;
    std::vector<poasta::replay::NodeRec> node_recs;  // node records of the current graph for the search kernel
    bool recs_valid = false;

    poasta_graph_view flat_view() {
        if (!flat_valid) {
            flat.build(g);
            flat_valid = true;
            sidx_state = 0;
            packed_valid = recs_valid = false;
        }
        return flat.view();
    }

    const poasta::PackedGraph &packed_graph() {
        const poasta_graph_view v = flat_view();
        if (!packed_valid) {
            std::vector<uint32_t> scratch;
            packed.build(v, scratch);
            packed_valid = true;
        }
        return packed;
    }

    const poasta::SearchIndex &search_index(bool with_paths) {
        const poasta_graph_view v = flat_view();
        if (sidx_state == 0 || (with_paths && sidx_state != 2)) {
            sidx.build(v, with_paths);
            sidx_state = with_paths ? 2 : 1;
            recs_valid = false;
        }
        if (!recs_valid) {
            node_recs.clear();
            sidx.node_records(v, node_recs);
            recs_valid = true;
        }
        return sidx;
    }
};
