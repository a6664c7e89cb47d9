// Host side of the device-resident graph (dev_graph.hpp): writing an existing POAGraph into a family arena before a build
// (`poasta align -I graph`, bin/poasta.rs:303-317) and reading the finished graph back into the std::vector form the
// writers (poa_io.hpp) and the C API use. Both directions keep edge SLOT numbers and the free-list order, which decide the
// GFA `L` line order and the .poasta bytes.
#pragma once
#include <string>
#include <vector>

#include "dev_graph.hpp"
#include "poa_graph.hpp"

namespace poasta {
namespace dg {

// what an existing graph already occupies (for caps_for)
struct Occupancy {
    uint32_t V, E, I, A, N;
};
inline Occupancy occupancy(const POAGraph &g) {
    Occupancy o{(uint32_t)g.nodes.size(), (uint32_t)g.edges.size(), 0, 0, (uint32_t)g.sequences.size()};
    for (const auto &e : g.edges) o.I += (uint32_t)e.sequence_ids.size();
    for (const auto &n : g.nodes) o.A += (uint32_t)n.aligned_nodes.size();
    return o;
}

// Writes `g` into the (host copy of the) arena of family `f`; f->cap / arena_off must be set. Returns false if it does not fit
// or holds values the device layout cannot carry (weights or sequence ids >= 2^32).
inline bool upload(const POAGraph &g, Family *f, uint8_t *chunk_arena) {
    Store s(f, chunk_arena);
    const Occupancy o = occupancy(g);
    if (o.V > f->cap.nodes || o.E > f->cap.edges || o.I > f->cap.sids || o.A > f->cap.als || o.N > f->cap.seqs) return false;
    f->n_nodes = o.V, f->n_edge_slots = o.E, f->free_edge = g.free_edge, f->live_edges = g.live_edges;
    f->n_sids = f->n_als = 0;
    f->n_seqs_in_graph = o.N;
    for (uint32_t n = 0; n < o.V; ++n) {
        s.sym[n] = g.nodes[n].symbol;
        s.head_out[n] = g.nodes[n].head[0], s.head_in[n] = g.nodes[n].head[1];
        s.al_head[n] = NIL;
        uint32_t prev = NIL;
        for (uint32_t other : g.nodes[n].aligned_nodes) {
            const uint32_t k = f->n_als++;
            s.al_node[k] = other, s.al_next[k] = NIL;
            if (prev == NIL) s.al_head[n] = k;
            else s.al_next[prev] = k;
            prev = k;
        }
    }
    for (uint32_t e = 0; e < o.E; ++e) {
        const POAGraph::Edge &ed = g.edges[e];
        s.e_src[e] = ed.node[0], s.e_dst[e] = ed.node[1];
        s.e_next_out[e] = ed.next[0], s.e_next_in[e] = ed.next[1];
        if (ed.weight > 0xFFFFFFFFull) return false;
        s.e_weight[e] = (uint32_t)ed.weight;
        s.e_sid_head[e] = s.e_sid_tail[e] = NIL;
        for (uint64_t id : ed.sequence_ids) {
            if (id >= 0xFFFFFFFFull) return false;
            const uint32_t k = f->n_sids++;
            s.sid_val[k] = (uint32_t)id, s.sid_next[k] = NIL;
            if (s.e_sid_tail[e] == NIL) s.e_sid_head[e] = k;
            else s.sid_next[s.e_sid_tail[e]] = k;
            s.e_sid_tail[e] = k;
        }
    }
    for (uint32_t i = 0; i < o.N; ++i) s.seq_start[i] = g.sequences[i].start_node;
    return true;
}

// Reads family `f` back. `names` = names of the sequences this build appended (in order); earlier ones keep theirs from `g`.
// Ends with the toposort the reference runs inside post_process (poa.rs:360); false = cycle.
inline bool download(Family *f, uint8_t *chunk_arena, POAGraph &g, const std::vector<std::string> &names) {
    const Store s(f, chunk_arena);
    const size_t old_n = g.sequences.size();
    // every list of the graph becomes a slice of one of two pools (PoolList, poa_graph.hpp): two allocations per graph instead of
    // one per list. The pools are sized first and not touched again, so the slices stay valid; lists the graph already had are
    // re-pointed without being read.
    g.al_pool.resize(f->n_als);
    g.id_pool.resize(f->n_sids);
    uint32_t *al_pool = g.al_pool.data();
    uint64_t *id_pool = g.id_pool.data();
    size_t al_used = 0, id_used = 0;
    g.nodes.resize(f->n_nodes);
    for (uint32_t n = 0; n < f->n_nodes; ++n) {
        POAGraph::Node &nd = g.nodes[n];
        nd.symbol = s.sym[n];
        nd.head[0] = s.head_out[n], nd.head[1] = s.head_in[n];
        const size_t al0 = al_used;
        for (uint32_t k = s.al_head[n]; k != NIL; k = s.al_next[k]) {
            if (al_used >= g.al_pool.size()) return false;  // corrupt arena
            al_pool[al_used++] = s.al_node[k];
        }
        nd.aligned_nodes.set_pool_filled(al_pool + al0, (uint32_t)(al_used - al0));
    }
    g.edges.resize(f->n_edge_slots);
    for (uint32_t e = 0; e < f->n_edge_slots; ++e) {
        POAGraph::Edge &ed = g.edges[e];
        ed.node[0] = s.e_src[e], ed.node[1] = s.e_dst[e];
        ed.next[0] = s.e_next_out[e], ed.next[1] = s.e_next_in[e];
        ed.vacant = s.e_src[e] == NIL;
        ed.weight = s.e_weight[e];
        const size_t id0 = id_used;
        for (uint32_t k = s.e_sid_head[e]; k != NIL; k = s.sid_next[k]) {
            if (id_used >= g.id_pool.size()) return false;
            id_pool[id_used++] = s.sid_val[k];
        }
        ed.sequence_ids.set_pool_filled(id_pool + id0, (uint32_t)(id_used - id0));
    }
    g.free_edge = f->free_edge;
    g.live_edges = f->live_edges;
    g.sequences.resize(f->n_seqs_in_graph);
    for (uint32_t i = 0; i < f->n_seqs_in_graph; ++i) {
        if (i >= old_n) g.sequences[i].name = i - old_n < names.size() ? names[i - old_n] : std::string();
        g.sequences[i].start_node = s.seq_start[i];
    }
    return g.toposort();
}

}  // namespace dg
}  // namespace poasta
