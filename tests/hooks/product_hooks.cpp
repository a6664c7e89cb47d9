// TEST HOOKS — host builds of PRODUCT code (poasta_b200/csrc/*.hpp) so the logic the device runs can be checked on a CPU
// against the oracle. Test infrastructure: only tests/ and tools/ load this library; the product never does.
// (The oracle itself, oracle/liboracle.so, includes no product header.)
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../oracle/astar_oracle.hpp"
#include "../../poasta_b200/csrc/dev_graph_host.hpp"
#include "../../poasta_b200/csrc/dev_index.hpp"
#include "../../poasta_b200/csrc/flatten.hpp"
#include "../../poasta_b200/csrc/poa_io.hpp"
#include "../../poasta_b200/csrc/search_index.hpp"

using namespace poasta;

namespace {
struct HostSync {
    void operator()() const {}
    int bcast(int x) const { return x; }
};

#define FAIL(...)                              \
    do {                                       \
        snprintf(msg, msg_cap, __VA_ARGS__);   \
        return 1;                              \
    } while (0)

// field-by-field comparison of the arena graph with the std::vector graph
int compare_graphs(const dg::Store &s, const POAGraph &g, char *msg, size_t msg_cap) {
    const dg::Family &f = *s.f;
    if (f.n_nodes != g.nodes.size()) FAIL("n_nodes %u vs %zu", f.n_nodes, g.nodes.size());
    if (f.n_edge_slots != g.edges.size()) FAIL("edge slots %u vs %zu", f.n_edge_slots, g.edges.size());
    if (f.free_edge != g.free_edge || f.live_edges != g.live_edges) FAIL("free list head / live edges differ");
    for (uint32_t n = 0; n < f.n_nodes; ++n) {
        if (s.sym[n] != g.nodes[n].symbol) FAIL("symbol of node %u", n);
        if (s.head_out[n] != g.nodes[n].head[0] || s.head_in[n] != g.nodes[n].head[1]) FAIL("list heads of node %u", n);
        std::vector<uint32_t> al;
        for (uint32_t k = s.al_head[n]; k != dg::NIL; k = s.al_next[k]) al.push_back(s.al_node[k]);
        if (al != g.nodes[n].aligned_nodes) FAIL("aligned_nodes of node %u", n);
    }
    for (uint32_t e = 0; e < f.n_edge_slots; ++e) {
        const POAGraph::Edge &ed = g.edges[e];
        if (s.e_src[e] != ed.node[0] || s.e_dst[e] != ed.node[1]) FAIL("endpoints of edge %u", e);
        if (s.e_next_out[e] != ed.next[0] || s.e_next_in[e] != ed.next[1]) FAIL("links of edge %u", e);
        if (s.e_weight[e] != ed.weight) FAIL("weight of edge %u: %u vs %llu", e, s.e_weight[e], (unsigned long long)ed.weight);
        std::vector<uint64_t> ids;
        for (uint32_t k = s.e_sid_head[e]; k != dg::NIL; k = s.sid_next[k]) ids.push_back(s.sid_val[k]);
        if (ids != ed.sequence_ids) FAIL("sequence_ids of edge %u", e);
    }
    if (f.n_seqs_in_graph != g.sequences.size()) FAIL("sequence count");
    for (uint32_t i = 0; i < f.n_seqs_in_graph; ++i)
        if (s.seq_start[i] != g.sequences[i].start_node) FAIL("start node of sequence %u", i);
    return 0;
}
}  // namespace

extern "C" {

// Progressive build of one family with the ORACLE's alignments, inserted (a) into the std::vector POAGraph and (b) into a
// dev_graph.hpp arena by the code the device runs; after every insertion the two graphs are compared field by field and the
// per-round derived arrays of dev_index.hpp are compared with search_index.hpp (itself pinned against the oracle's
// BubbleIndex by tests/test_device_layouts.py). graph_in (optional .poasta bytes) starts both from an existing graph.
// Finally the arena is read back (dev_graph_host.hpp) and must serialise to the same .poasta bytes. 0 = all equal.
int devgraph_host_check(const char *fasta, size_t fasta_len, const uint8_t *graph_in, size_t graph_in_len, int x, int o, int e, int mode,
                        int heuristic, char *msg, size_t msg_cap) {
    auto recs = parse_fasta(fasta, fasta_len);
    POAGraph g;
    if (graph_in && graph_in_len && !graph_from_poasta(graph_in, graph_in_len, g)) FAIL("bad graph_in");
    uint64_t S = 0;
    for (auto &r : recs) S += r.seq.size();
    const dg::Occupancy occ = dg::occupancy(g);
    dg::Family fam;
    memset(&fam, 0, sizeof fam);
    fam.cap = dg::caps_for(S, (uint32_t)recs.size(), occ.V, occ.E, occ.I, occ.A, occ.N);
    fam.arena_off = 0;
    std::vector<uint8_t> arena(dg::Layout(fam.cap).total);
    dg::Store s(&fam, arena.data());
    if (occ.V > 2 || occ.N) {
        if (!dg::upload(g, &fam, arena.data())) FAIL("upload failed");
    } else {
        s.init_empty();
    }
    if (compare_graphs(s, g, msg, msg_cap)) return 1;
    oracle::AlignOptions opt;
    opt.costs = oracle::Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode, opt.heuristic = heuristic;
    FlatGraph flat;
    SearchIndex si;
    std::vector<replay::NodeRec> recs_host;
    std::vector<uint8_t> scratch;
    std::vector<std::string> names;
    for (size_t i = 0; i < recs.size(); ++i) {
        const uint8_t *q = (const uint8_t *)recs[i].seq.data();
        const uint32_t Q = (uint32_t)recs[i].seq.size();
        names.push_back(recs[i].name);
        if (g.is_empty()) {
            if (!s.is_empty()) FAIL("is_empty differs");
            g.add_alignment_with_weights(recs[i].name, q, Q, false, nullptr, 0);
            if (int rc = s.add_first_sequence(q, Q)) FAIL("add_first_sequence rc %d", rc);
        } else {
            flat.build(g);
            const poasta_graph_view view = flat.view();
            // derived arrays: device code on the host vs search_index.hpp
            si.build(view, heuristic == 2);
            recs_host.clear();
            si.node_records(view, recs_host);
            dg::IndexCaps ic{fam.n_nodes, fam.live_edges, 16 * fam.n_nodes + 16};
            scratch.assign(dg::IndexLayout(ic).total, 0xAB);
            dg::IndexBuf w(scratch.data(), ic);
            if (int rc = dg::build_index(s, w, heuristic == 2, 0, 1, HostSync{})) FAIL("build_index rc %d", rc);
            const uint32_t V = view.n_nodes;
            if (memcmp(w.succ, view.succ, (size_t)view.succ_off[V] * 4) || memcmp(w.pred, view.pred, (size_t)view.pred_off[V] * 4))
                FAIL("round %zu: adjacency differs", i);
            if (memcmp(w.nodes, recs_host.data(), (size_t)V * sizeof(replay::NodeRec))) FAIL("round %zu: node records differ", i);
            if (w.n_bub * 4 != si.bub.size() || memcmp(w.bub, si.bub.data(), si.bub.size() * 4)) FAIL("round %zu: bubble quadruples differ", i);
            if (heuristic == 2) {
                if (memcmp(w.path_off, si.path_off.data(), (size_t)(V + 1) * 4)) FAIL("round %zu: path_off differs", i);
                if (w.n_path != si.path_rem.size() || memcmp(w.path_rem, si.path_rem.data(), si.path_rem.size() * 4))
                    FAIL("round %zu: path_rem differs", i);
            }
            oracle::AstarResult r = oracle::astar_align(view, q, Q, opt);
            if (r.status) FAIL("oracle panic %d", r.status);
            std::vector<AlignedPair> aln(r.rpos.size());
            for (size_t k = 0; k < aln.size(); ++k) aln[k] = {r.rpos[k], r.qpos[k]};
            const GraphError ge = g.add_alignment_with_weights(recs[i].name, q, Q, true, aln.data(), aln.size());
            const int rc = s.add_alignment(q, Q, r.rpos.data(), r.qpos.data(), (uint32_t)r.rpos.size());
            if ((int)ge != rc) FAIL("round %zu: insert status %d vs %d", i, (int)ge, rc);
            if (rc) break;
        }
        if (compare_graphs(s, g, msg, msg_cap)) {
            const size_t n = strlen(msg);
            snprintf(msg + n, msg_cap - n, " (after record %zu)", i);
            return 1;
        }
    }
    POAGraph back;
    if (graph_in && graph_in_len) graph_from_poasta(graph_in, graph_in_len, back);
    if (!recs.empty() && !dg::download(&fam, arena.data(), back, names)) FAIL("download: cycle");
    if (!recs.empty()) {
        std::string a, b;
        graph_to_poasta(g, a);
        graph_to_poasta(back, b);
        if (a != b) FAIL("downloaded graph serialises differently (%zu vs %zu bytes)", a.size(), b.size());
    }
    return 0;
}

}  // extern "C"
