// C ABI: POAGraph handle + readers/writers (include/poasta_b200.h, host-only half).
// Reference interfaces replaced: graphs/poa.rs:96-382, io/fasta.rs:69, io/graph.rs:24-103,245-502.
#include <atomic>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <new>

#include "graph_handle.hpp"

using namespace poasta;

static int map_graph_error(GraphError e) {
    switch (e) {
        case GraphError::Ok: return POASTA_OK;
        case GraphError::WeightsUnequalSize: return POASTA_ERR_WEIGHTS_UNEQUAL;
        case GraphError::InvalidAlignment: return POASTA_ERR_INVALID_ALIGNMENT;
        case GraphError::Cycle: return POASTA_ERR_GRAPH_CYCLE;
    }
    return POASTA_ERR_INVALID_ARG;
}

// No C++ exception may cross the C boundary: allocation failures (e.g. on a crafted graph file) become a status.
template <class F> static int guarded(F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return POASTA_ERR_IO;
    } catch (...) {
        return POASTA_ERR_INTERNAL;
    }
}

static int emit(const std::string &buf, char **out, size_t *len) {
    char *p = (char *)malloc(buf.size() + 1);
    if (!p) return POASTA_ERR_IO;
    memcpy(p, buf.data(), buf.size());
    p[buf.size()] = 0;
    *out = p;
    *len = buf.size();
    return POASTA_OK;
}

extern "C" {

poasta_graph *poasta_graph_new(void) { return new (std::nothrow) poasta_graph(); }
void poasta_graph_free(poasta_graph *g) { delete g; }
void poasta_graph_free_many(poasta_graph **graphs, uint32_t n, uint32_t n_threads) {
    // a finished 20 x 1 kb graph owns a few thousand small vectors (aligned-node lists, sequence ids per edge): freeing the
    // thousands of graphs a batch returns is worth spreading over the host cores
    if (!graphs || n == 0) return;
    unsigned nt = n_threads ? n_threads : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (n < 64 || nt <= 1) {
        for (uint32_t i = 0; i < n; ++i) delete graphs[i], graphs[i] = nullptr;
        return;
    }
    std::atomic<uint32_t> cursor{0};
    auto work = [&]() {
        for (;;) {
            const uint32_t i0 = cursor.fetch_add(16);
            if (i0 >= n) break;
            for (uint32_t i = i0; i < std::min(n, i0 + 16); ++i) delete graphs[i], graphs[i] = nullptr;
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
}
uint32_t poasta_graph_node_count(const poasta_graph *g) { return g->g.node_count(); }
uint32_t poasta_graph_node_count_with_start_and_end(const poasta_graph *g) { return g->g.node_count_with_start_and_end(); }
uint32_t poasta_graph_edge_count(const poasta_graph *g) { return g->g.edge_count(); }
uint32_t poasta_graph_num_sequences(const poasta_graph *g) { return (uint32_t)g->g.sequences.size(); }

int poasta_graph_add_alignment(poasta_graph *h, const char *name, const uint8_t *seq, uint32_t len, const uint32_t *rpos,
                               const uint32_t *qpos, uint32_t n_pairs, const uint64_t *weights, uint32_t weights_len) {
    if (!h || !name || (len && !seq)) return POASTA_ERR_INVALID_ARG;
    if (weights && weights_len != len) return POASTA_ERR_WEIGHTS_UNEQUAL;  // poa.rs:178-180
    return guarded([&]() -> int {
        h->flat_valid = false;
        if (!rpos || !qpos) return map_graph_error(h->g.add_alignment_with_weights(name, seq, len, false, nullptr, 0, weights));
        // the reference indexes graph[rpos] and sequence[qpos] and panics when they are out of range; here that is a status
        const uint32_t nn = h->g.node_count_with_start_and_end();
        std::vector<AlignedPair> aln(n_pairs);
        for (uint32_t i = 0; i < n_pairs; ++i) {
            if ((rpos[i] != NIL && rpos[i] >= nn) || (qpos[i] != NIL && qpos[i] >= len)) return POASTA_ERR_INVALID_ALIGNMENT;
            aln[i] = {rpos[i], qpos[i]};
        }
        return map_graph_error(h->g.add_alignment_with_weights(name, seq, len, true, aln.data(), n_pairs, weights));
    });
}

int poasta_graph_get_view(poasta_graph *h, poasta_graph_view *view) {
    if (!h || !view) return POASTA_ERR_INVALID_ARG;
    if (h->g.topological_sorted.size() != h->g.nodes.size()) {
        GraphError e = h->g.post_process();  // a fresh graph has never been post-processed
        if (e != GraphError::Ok) return map_graph_error(e);
        h->flat_valid = false;
    }
    *view = h->flat_view();
    return POASTA_OK;
}

uint32_t poasta_graph_aligned_nodes(const poasta_graph *h, uint32_t node, uint32_t *out, uint32_t cap) {
    if (!h || node >= h->g.nodes.size()) return 0;
    const auto &a = h->g.nodes[node].aligned_nodes;
    for (uint32_t i = 0; i < a.size() && i < cap; ++i) out[i] = a[i];
    return (uint32_t)a.size();
}

int poasta_graph_to_fasta(const poasta_graph *h, char **out, size_t *len) {
    if (!h || !out || !len) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        std::string buf;
        if (!graph_to_fasta(h->g, buf)) return POASTA_ERR_IO;
        return emit(buf, out, len);
    });
}

int poasta_graph_to_gfa(const poasta_graph *h, int v1, char **out, size_t *len) {
    if (!h || !out || !len) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        std::string buf;
        if (!graph_to_gfa(h->g, buf, v1 != 0)) return POASTA_ERR_IO;
        return emit(buf, out, len);
    });
}

int poasta_graph_to_dot(const poasta_graph *h, int dialect, char **out, size_t *len) {
    if (!h || !out || !len) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        std::string buf;
        if (dialect == 0) graph_to_dot_view(h->g, buf);
        else graph_to_dot_petgraph(h->g, buf);
        return emit(buf, out, len);
    });
}

int poasta_graph_save(const poasta_graph *h, char **out, size_t *len) {
    if (!h || !out || !len) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        std::string buf;
        graph_to_poasta(h->g, buf);
        return emit(buf, out, len);
    });
}

int poasta_graph_load(const uint8_t *data, size_t len, poasta_graph **out) {
    if (!data || !out) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        poasta_graph *h = new (std::nothrow) poasta_graph();
        if (!h) return POASTA_ERR_IO;
        bool ok = false;
        try {
            ok = graph_from_poasta(data, len, h->g);
        } catch (...) {
            delete h;
            throw;
        }
        if (!ok) {
            delete h;
            return POASTA_ERR_IO;
        }
        *out = h;
        return POASTA_OK;
    });
}

int poasta_graph_from_fasta_msa(const char *fasta, size_t len, poasta_graph **out) {
    if (!fasta || !out) return POASTA_ERR_INVALID_ARG;
    return guarded([&]() -> int {
        poasta_graph *h = new (std::nothrow) poasta_graph();
        if (!h) return POASTA_ERR_IO;
        GraphError e;
        try {
            auto recs = parse_fasta(fasta, len);
            e = graph_from_fasta_msa(recs, h->g);
        } catch (...) {
            delete h;
            throw;
        }
        if (e != GraphError::Ok) {
            delete h;
            return map_graph_error(e);
        }
        *out = h;
        return POASTA_OK;
    });
}

void poasta_free_buffer(void *p) { free(p); }

const char *poasta_version(void) { return "poasta_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
