// TEST HOOKS — host builds of PRODUCT code (poasta_b200/csrc/*.hpp) so the logic the device runs can be checked on a CPU
// against the oracle. Test infrastructure: only tests/ and tools/ load this library; the product never does.
// (The oracle itself, oracle/liboracle.so, includes no product header.)
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <chrono>

#include "../../oracle/astar_oracle.hpp"
#include "../../poasta_b200/csrc/device_batch.hpp"
#include "../../poasta_b200/csrc/dev_graph_host.hpp"
#include "../../poasta_b200/csrc/dev_dp_pack.hpp"
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
            // DP packing (dev_dp_pack.hpp, the DP-only mode's per-round step): device code on the host vs PackedGraph::build.
            // The two use different topological orders, so everything is compared in node-id space.
            {
                const uint32_t E = view.pred_off[V];
                dg::DpPackCaps pc{fam.n_nodes, fam.live_edges};
                std::vector<uint8_t> pscratch(dg::DpPackLayout(pc).total, 0xCD);
                dg::DpPackScratch pw(pscratch.data(), pc);
                std::vector<RowRec> rows(V);
                std::vector<uint32_t> node_id(V), pred_off(V + 1), pred_rank(E + 1);
                std::vector<uint8_t> node_sym(V);
                const uint32_t pred_base = 7;
                dg::DpPackOut po{rows.data(), node_id.data(), node_sym.data(), pred_off.data(), pred_rank.data(), pred_base};
                uint32_t maxdeg = 0;
                if (int rc = dg::dp_pack(s, pw, po, &maxdeg, 0, 1, HostSync{})) FAIL("round %zu: dp_pack rc %d", i, rc);
                PackedGraph pg;
                std::vector<uint32_t> rs;
                pg.build(view, rs);
                if (maxdeg != pg.max_in_degree) FAIL("round %zu: dp_pack max in-degree %u vs %u", i, maxdeg, pg.max_in_degree);
                if (node_id[0] != view.start_node || node_id[V - 1] != view.end_node) FAIL("round %zu: dp_pack start / end ranks", i);
                std::vector<uint32_t> host_rank(V, 0xFFFFFFFFu), seen(V, 0);
                for (uint32_t r = 0; r < V; ++r) host_rank[pg.node_id[r]] = r;
                if (pred_off[0] != pred_base || pred_off[V] != pred_base + E) FAIL("round %zu: dp_pack pred_off ends", i);
                for (uint32_t r = 0; r < V; ++r) {
                    const uint32_t v = node_id[r];
                    if (v >= V || seen[v]++) FAIL("round %zu: dp_pack node_id is not a permutation", i);
                    const uint32_t hr = host_rank[v];
                    if (node_sym[r] != pg.node_sym[hr]) FAIL("round %zu: dp_pack symbol of node %u", i, v);
                    if ((rows[r].info & 0xFFF) != (pg.rows[hr].info & 0xFFF)) FAIL("round %zu: dp_pack info of node %u: %x vs %x", i, v, rows[r].info, pg.rows[hr].info);
                    if (rows[r].npred_sym != pg.rows[hr].npred_sym) FAIL("round %zu: dp_pack npred_sym of node %u", i, v);
                    const uint32_t np = rows[r].npred_sym & 0xFFFFFFu, b = pred_off[r] - pred_base, hb = pg.pred_off[hr];
                    if (rows[r].pred_begin != b || pred_off[r + 1] - pred_off[r] != np) FAIL("round %zu: dp_pack pred range of node %u", i, v);
                    for (uint32_t k = 0; k < np; ++k) {
                        if (pred_rank[b + k] >= r) FAIL("round %zu: dp_pack order is not topological at node %u", i, v);
                        if (node_id[pred_rank[b + k]] != pg.node_id[pg.pred_rank[hb + k]]) FAIL("round %zu: dp_pack predecessor %u of node %u", i, k, v);
                    }
                    const bool prev = np == 1 && pred_rank[b] + 1 == r;
                    if ((rows[r].pred0_slot == ROW_PREV_IN_REGS) != prev) FAIL("round %zu: dp_pack previous-row flag of node %u", i, v);
                }
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


// Host cost of graph insertion (add_alignment_with_weights + post_process) and CSR flattening for one family: the MSA is built
// once with the oracle to obtain the alignments, then the insertions are replayed `reps` times. Microseconds per sequence.
int host_insert_timing(const char *fasta, size_t fasta_len, int reps, double *us_insert, double *us_flatten) {
    using namespace poasta;
    auto recs = parse_fasta(fasta, fasta_len);
    oracle::AlignOptions opt;
    opt.costs = oracle::Costs{4, 6, 2};
    opt.mode = 0, opt.heuristic = 1;
    std::vector<std::vector<AlignedPair>> alns(recs.size());
    {
        POAGraph g;
        FlatGraph flat;
        for (size_t i = 0; i < recs.size(); ++i) {
            const uint8_t *s = (const uint8_t *)recs[i].seq.data();
            if (!g.is_empty()) {
                flat.build(g);
                oracle::AstarResult r = oracle::astar_align(flat.view(), s, recs[i].seq.size(), opt);
                if (r.status) return -1;
                alns[i].resize(r.rpos.size());
                for (size_t k = 0; k < alns[i].size(); ++k) alns[i][k] = {r.rpos[k], r.qpos[k]};
            }
            g.add_alignment_with_weights(recs[i].name, s, recs[i].seq.size(), i != 0, alns[i].data(), alns[i].size());
        }
    }
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_ins = 0, t_flat = 0;
    for (int rep = 0; rep < reps; ++rep) {
        POAGraph g;
        FlatGraph flat;
        for (size_t i = 0; i < recs.size(); ++i) {
            const uint8_t *s = (const uint8_t *)recs[i].seq.data();
            const double t0 = now();
            g.add_alignment_with_weights(recs[i].name, s, recs[i].seq.size(), i != 0, alns[i].data(), alns[i].size());
            const double t1 = now();
            flat.build(g);
            const double t2 = now();
            t_ins += t1 - t0, t_flat += t2 - t1;
        }
    }
    *us_insert = t_ins / (reps * (double)recs.size());
    *us_flatten = t_flat / (reps * (double)recs.size());
    return 0;
}

// Lifetimes of the pool-backed lists of a downloaded graph (PoolList, poa_graph.hpp; dev_graph_host.hpp download): the graph read
// back from an arena must serialise like the one it was written from; lists that grow past their pool slice, copies, assignments,
// moves, a second download into a graph that already has lists, and continuing the build on the host must all leave graphs that
// serialise like their std-path twins. 0 = all equal. (Run under ASan when changing PoolList.)
int poollist_lifetime_check(const uint8_t *graph_in, size_t graph_in_len, char *msg, size_t msg_cap) {
    POAGraph g;
    if (!graph_from_poasta(graph_in, graph_in_len, g)) FAIL("bad graph_in");
    const dg::Occupancy occ = dg::occupancy(g);
    dg::Family fam;
    memset(&fam, 0, sizeof fam);
    fam.cap = dg::Caps{occ.V, occ.E, occ.I + 1, occ.A + 1, occ.N + 1};
    std::vector<uint8_t> arena(dg::Layout(fam.cap).total);
    if (!dg::upload(g, &fam, arena.data())) FAIL("upload failed");
    std::vector<std::string> names;
    for (auto &sq : g.sequences) names.push_back(sq.name);
    auto bytes = [](const POAGraph &x) {
        std::string out;
        graph_to_poasta(x, out);
        return out;
    };
    const std::string want = bytes(g);
    auto load = [&](POAGraph &into) {
        dg::Family F = fam;
        return dg::download(&F, arena.data(), into, names);
    };
    POAGraph h;
    if (!load(h)) FAIL("download failed");
    if (bytes(h) != want) FAIL("downloaded graph serialises differently");
    // grow lists past their slices, on the pool-backed graph and on its std-path twin
    POAGraph twin = g;
    for (int i = 0; i < 10; ++i) h.nodes[2].aligned_nodes.push_back(7 + i), twin.nodes[2].aligned_nodes.push_back(7 + i);
    for (size_t e = 0; e < h.edges.size() && e < 50; ++e)
        if (!h.edges[e].vacant)
            for (int i = 0; i < 40; ++i) h.edges[e].sequence_ids.push_back(1000 + i), twin.edges[e].sequence_ids.push_back(1000 + i);
    if (bytes(h) != bytes(twin)) FAIL("lists grown past their pool slices differ");
    POAGraph c = h;  // copy: owning lists
    if (bytes(c) != bytes(twin)) FAIL("copy differs");
    h = c;           // copy-assignment over pool-backed lists
    POAGraph m = std::move(c);
    if (bytes(h) != bytes(twin) || bytes(m) != bytes(twin)) FAIL("assignment / move differs");
    if (!load(h) || !load(m)) FAIL("second download failed");  // re-points every list of graphs that already had lists
    if (bytes(h) != want || bytes(m) != want) FAIL("second download differs");
    // continue the build on the host (new nodes, new edges, ids appended to pool-backed lists, start / end edges recycled)
    const uint8_t extra[] = "ACGTTGCAACGT";
    POAGraph t2 = g;
    if (h.add_alignment_with_weights("x", extra, 12, false, nullptr, 0) != t2.add_alignment_with_weights("x", extra, 12, false, nullptr, 0))
        FAIL("insert status differs");
    std::vector<AlignedPair> aln;  // the first sequence of the graph against itself: walk its path
    std::string seq;
    for (uint32_t v = g.sequences[0].start_node; v != g.end_node && seq.size() < 400;) {
        aln.push_back({v, (uint32_t)seq.size()});
        seq.push_back((char)g.nodes[v].symbol);
        uint32_t next = NIL;
        for (uint32_t e = g.nodes[v].head[0]; e != NIL; e = g.edges[e].next[0])
            if (std::binary_search(g.edges[e].sequence_ids.begin(), g.edges[e].sequence_ids.end(), (uint64_t)0)) next = g.edges[e].node[1];
        if (next == NIL) break;
        v = next;
    }
    if (h.add_alignment_with_weights("y", (const uint8_t *)seq.data(), seq.size(), true, aln.data(), aln.size()) !=
        t2.add_alignment_with_weights("y", (const uint8_t *)seq.data(), seq.size(), true, aln.data(), aln.size()))
        FAIL("insert status differs (aligned)");
    if (bytes(h) != bytes(t2)) FAIL("graphs differ after continuing the build on the host");
    std::vector<POAGraph> v;
    v.push_back(h);
    v.push_back(std::move(m));
    v.resize(6);
    v.erase(v.begin());
    if (bytes(v[0]) != want) FAIL("moved graph differs");
    return 0;
}

// Host build of the product's order-exact search core (poasta_b200/csrc/replay_core.hpp), for CPU-side tests of the
// logic that runs on the device. TEST HOOK ONLY: the product never executes this path (it has no CPU fallback).
int replay_host_align(const poasta_graph_view *view, const uint8_t *seq, uint32_t qlen, int x, int o, int e, int mode,
                      int heuristic, uint32_t *score, uint32_t *rpos, uint32_t *qpos, uint32_t cap, uint32_t *n_pairs,
                      uint64_t *stats, int cell16) {
    namespace rp = poasta::replay;
    using rp::Item;
    using rp::Params;
    using rp::Result;
    using rp::Scratch;
    using rp::ScratchDims;
    poasta::SearchIndex si;
    si.build(*view, heuristic == 2);
    std::vector<rp::NodeRec> recs;
    si.node_records(*view, recs);
    std::vector<uint32_t> path_off_zero;
    if (si.path_off.empty()) path_off_zero.assign(view->n_nodes + 1, 0);
    rp::Graph g{view->n_nodes, view->start_node, view->end_node, recs.data(), view->succ, view->pred, reinterpret_cast<const rp::BubQ *>(si.bub.data()),
                si.path_off.empty() ? path_off_zero.data() : si.path_off.data(), si.path_rem.empty() ? nullptr : si.path_rem.data()};
    ScratchDims d(view->n_nodes, qlen, x, o, e);
    std::vector<uint32_t> qhead((size_t)d.f_max * 3, 0), qwin(rp::QWIN * 3, 0), rb((size_t)view->n_nodes * 2, 0),
        dfa((size_t)(view->n_nodes + 2) * 6, 0);
    // generous arena for the host build: every state can be queued once per successful relaxation
    const uint32_t max_items = (3 * view->n_nodes * (qlen + 2) + 65536) & ~7u;
    const uint64_t plane_elems = (uint64_t)view->n_nodes * d.W;
    // records {M, D, I, epoch}: zero-filled = "written by no alignment"; this alignment runs as epoch 1 (RP_HOST_EPOCH overrides
    // it and pre-fills the table with records of OTHER epochs, to check that stale records read as unvisited)
    std::vector<uint8_t> planes((size_t)4 * plane_elems * (cell16 ? 2 : 4), 0);
    uint32_t epoch = 1;
    if (const char *ep = getenv("RP_HOST_EPOCH")) {
        epoch = (uint32_t)atoi(ep);
        uint32_t x = 12345;
        for (size_t i = 0; i < planes.size(); i += (cell16 ? 8 : 16)) {  // garbage scores stamped with epochs != epoch
            for (int k = 0; k < (cell16 ? 8 : 16); ++k) planes[i + k] = (uint8_t)((x = x * 1664525u + 1013904223u) >> 24);
            uint32_t other = (x >> 8) % 0xFFFEu + 1;
            if (other == epoch) other = epoch == 1 ? 2 : epoch - 1;
            if (cell16) ((uint16_t *)(planes.data() + i))[3] = (uint16_t)other;
            else ((uint32_t *)(planes.data() + i))[3] = other;
        }
    }
    std::vector<Item> items(max_items);
    std::vector<uint32_t> free_slots(max_items);
    const uint32_t Wb = (d.W + 31) / 32;
    std::vector<uint32_t> rbits((size_t)view->n_nodes * Wb, 0);
    Scratch s{planes.data(), epoch, qhead.data(), qwin.data(), items.data(), free_slots.data(), rb.data(), rbits.data(), Wb, dfa.data(), d.W, max_items, d.f_max};
    // RP_HOST_BLOCKED=<pool blocks>: the blocked visited table (64-offset blocks behind a map) instead of the dense one; the map is
    // pre-filled with entries of OTHER map epochs, which must read as "no block"
    const bool blocked = getenv("RP_HOST_BLOCKED") != nullptr;
    std::vector<uint32_t> bmap;
    if (blocked) {
        s.bpr = (d.W + rp::BLK_CELLS - 1) / rp::BLK_CELLS;
        s.max_blocks = (uint32_t)std::min<uint64_t>((uint64_t)atoll(getenv("RP_HOST_BLOCKED")), rp::BMAP_BLOCK_MASK);
        planes.resize(std::max<size_t>(planes.size(), (size_t)s.max_blocks * rp::BLK_CELLS * (cell16 ? 8 : 16)), 0);
        s.planes = planes.data();
        s.bmap_epoch = 5;
        bmap.resize((size_t)view->n_nodes * s.bpr);
        uint32_t x = 999;
        for (auto &e : bmap) {
            x = x * 1664525u + 1013904223u;
            uint32_t other = (x >> 8) % rp::BMAP_EPOCH_MAX + 1;
            if (other == s.bmap_epoch) other = 6;
            e = (other << 20) | ((x >> 3) & rp::BMAP_BLOCK_MASK);
        }
        s.bmap = bmap.data();
    }
    Params p{(uint32_t)x, (uint32_t)o, (uint32_t)e, mode, heuristic};
    const uint32_t ocap = view->n_nodes + qlen + 2;
    std::vector<uint32_t> out_r(ocap), out_q(ocap);
    Result r;
    uint32_t used_blocks = 0, used_items = 0;
    auto go = [&](auto cell, auto heur) {
        // RP_HOST_PREINIT: the start-up of an ends-free alignment through prepare_initial + run<true> (what the device's lanes share)
        const bool preinit = getenv("RP_HOST_PREINIT") != nullptr;
        if (blocked) {
            rp::Search<decltype(cell), decltype(heur)::value, true> search(g, seq, qlen, p, s);
            r = (preinit && search.prepare_initial(0, 1)) ? search.template run<true>(out_r.data(), out_q.data(), ocap) : search.run(out_r.data(), out_q.data(), ocap);
            used_items = search.n_items;
            used_blocks = search.n_blocks;
        } else {
            rp::Search<decltype(cell), decltype(heur)::value> search(g, seq, qlen, p, s);
            r = (preinit && search.prepare_initial(0, 1)) ? search.template run<true>(out_r.data(), out_q.data(), ocap) : search.run(out_r.data(), out_q.data(), ocap);
            used_items = search.n_items;
        }
    };
    auto go_cell = [&](auto cell) {
        if (heuristic == 0) go(cell, std::integral_constant<int, 0>{});
        else if (heuristic == 1) go(cell, std::integral_constant<int, 1>{});
        else go(cell, std::integral_constant<int, 2>{});
    };
    if (cell16) go_cell(uint16_t{});
    else go_cell(uint32_t{});
    *score = r.score;
    *n_pairs = r.n_pairs;
    for (uint32_t i = 0; i < r.n_pairs && i < cap; ++i) {
        rpos[i] = out_r[ocap - r.n_pairs + i];
        qpos[i] = out_q[ocap - r.n_pairs + i];
    }
    if (stats) stats[0] = r.num_queued, stats[1] = r.num_visited, stats[2] = r.num_pruned, stats[3] = used_blocks, stats[4] = used_items;
    return r.status;
}

// The product's host-side packing of one graph for the DP kernels (device_batch.hpp) and its node records for the search
// kernel (search_index.hpp), dumped so CPU tests can pin the layouts the device code reads. TEST HOOK ONLY.
//   meta[4] = V, max_in_degree, n_stored, n_ord_rows; rows = 4 words per row (RowRec); recs = 8 words per node (NodeRec)
int packed_graph_dump(const poasta_graph_view *view, uint32_t *meta, uint32_t *rows, uint32_t *node_info, uint32_t *node_id,
                      uint32_t *ord_slot, uint32_t *pred_off, uint32_t *pred_rank, uint32_t *pred_slot, uint8_t *node_sym, uint32_t *recs) {
    poasta::PackedGraph pg;
    std::vector<uint32_t> scratch;
    pg.build(*view, scratch);
    meta[0] = pg.V, meta[1] = pg.max_in_degree, meta[2] = pg.n_stored, meta[3] = pg.n_ord_rows;
    static_assert(sizeof(poasta::RowRec) == 16 && sizeof(poasta::replay::NodeRec) == 32, "device record sizes");
    memcpy(rows, pg.rows.data(), pg.rows.size() * sizeof(poasta::RowRec));
    memcpy(node_info, pg.node_info.data(), pg.node_info.size() * 4);
    memcpy(node_id, pg.node_id.data(), pg.node_id.size() * 4);
    memcpy(ord_slot, pg.ord_slot.data(), pg.ord_slot.size() * 4);
    memcpy(pred_off, pg.pred_off.data(), pg.pred_off.size() * 4);
    memcpy(pred_rank, pg.pred_rank.data(), pg.pred_rank.size() * 4);
    memcpy(pred_slot, pg.pred_slot.data(), pg.pred_slot.size() * 4);
    memcpy(node_sym, pg.node_sym.data(), pg.node_sym.size());
    poasta::SearchIndex si;
    si.build(*view, false);
    std::vector<poasta::replay::NodeRec> nr;
    si.node_records(*view, nr);
    memcpy(recs, nr.data(), nr.size() * sizeof(poasta::replay::NodeRec));
    return 0;
}

// The PRODUCT's search index (poasta_b200/csrc/search_index.hpp) in the same shape as oracle_bubble_index, so tests can
// compare the two restatements entry by entry. TEST HOOK ONLY.
void product_bubble_index(const poasta_graph_view *view, uint32_t max_per_node, uint32_t *counts, uint32_t *out_maps, uint32_t *dist,
                          uint8_t *is_exit) {
    poasta::SearchIndex si;
    si.build(*view, false);
    for (uint32_t n = 0; n < view->n_nodes; ++n) {
        counts[n] = si.bub_off[n + 1] - si.bub_off[n];
        for (uint32_t k = 0; k < counts[n] && k < max_per_node; ++k) {
            const uint32_t *q = si.bub.data() + 4 * (size_t)(si.bub_off[n] + k);
            uint32_t *o = out_maps + 3 * ((size_t)n * max_per_node + k);
            o[0] = q[0], o[1] = q[1], o[2] = q[2];
            // the fourth word is derived: min_dist_to_end of the exit used by the prune test (reached.rs:57-58)
            if (q[3] != (si.dist_min[q[0]] ? si.dist_min[q[0]] - 1 : 0)) counts[n] = 0xFFFFFFFFu;
        }
        dist[2 * n] = si.dist_min[n], dist[2 * n + 1] = si.dist_max[n];
        is_exit[n] = si.is_exit[n];
    }
}

// Host cost of the per-round preparation the product does for every graph (search index, DP packing, node records):
// milliseconds per call, averaged over `reps`. TEST / MEASUREMENT HOOK ONLY.
int host_prep_timing(const poasta_graph_view *view, int reps, double *ms_index, double *ms_pack, double *ms_recs) {
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    poasta::SearchIndex si;
    poasta::PackedGraph pg;
    std::vector<uint32_t> scratch;
    std::vector<poasta::replay::NodeRec> nr;
    double t0 = now();
    for (int i = 0; i < reps; ++i) si.build(*view, false);
    double t1 = now();
    for (int i = 0; i < reps; ++i) pg.build(*view, scratch);
    double t2 = now();
    for (int i = 0; i < reps; ++i) nr.clear(), si.node_records(*view, nr);
    double t3 = now();
    *ms_index = (t1 - t0) / reps, *ms_pack = (t2 - t1) / reps, *ms_recs = (t3 - t2) / reps;
    return 0;
}

// event counters of the host build of the search core (tools/search_stats.py); reading resets them
int replay_host_counters(uint64_t *out, int cap) {
    namespace rp = poasta::replay;
    for (int i = 0; i < rp::C_N && i < cap; ++i) out[i] = rp::rp_counters[i], rp::rp_counters[i] = 0;
    return rp::C_N;
}

}  // extern "C"
