// The opaque `poasta_graph` of include/poasta_b200.h: a POAGraph<u32> plus its cached CSR view.
#pragma once
#include "flatten.hpp"
#include "poa_io.hpp"
#include "search_index.hpp"

struct poasta_graph {
    poasta::POAGraph g;
    poasta::FlatGraph flat;
    bool flat_valid = false;
    poasta::SearchIndex sidx;  // bubbles / distances / sampled paths of the current graph (order-exact search)
    int sidx_state = 0;        // 0 = stale, 1 = built without paths, 2 = built with paths

    poasta_graph_view flat_view() {
        if (!flat_valid) {
            flat.build(g);
            flat_valid = true;
            sidx_state = 0;
        }
        return flat.view();
    }

    const poasta::SearchIndex &search_index(bool with_paths) {
        const poasta_graph_view v = flat_view();
        if (sidx_state == 0 || (with_paths && sidx_state != 2)) {
            sidx.build(v, with_paths);
            sidx_state = with_paths ? 2 : 1;
        }
        return sidx;
    }
};
