// The opaque `poasta_graph` of include/poasta_b200.h: a POAGraph<u32> plus its cached CSR view.
#pragma once
#include "flatten.hpp"
#include "poa_io.hpp"

struct poasta_graph {
    poasta::POAGraph g;
    poasta::FlatGraph flat;
    bool flat_valid = false;

    poasta_graph_view flat_view() {
        if (!flat_valid) {
            flat.build(g);
            flat_valid = true;
        }
        return flat.view();
    }
};
