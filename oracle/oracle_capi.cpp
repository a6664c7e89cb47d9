// ORACLE — TEST INFRASTRUCTURE ONLY (see astar_oracle.hpp header).  C entry points for ctypes.
// The progressive-MSA driver below restates bin/poasta.rs:163-236 (perform_alignment) on top of
// the host graph model in poasta_b200/csrc/poa_graph.hpp (graph insertion is host code shared
// with the product and pinned separately by tests/test_graph_io.py); the ALIGNMENT inside the
// loop is the oracle's A* restatement, never the CUDA path.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../poasta_b200/csrc/flatten.hpp"
#include "../poasta_b200/csrc/poa_io.hpp"
#include "astar_oracle.hpp"
#include "dp_model.hpp"
#include "../poasta_b200/csrc/replay_core.hpp"
#include "../poasta_b200/csrc/search_index.hpp"
#include "../poasta_b200/csrc/device_batch.hpp"

using namespace oracle;

extern "C" {

// Returns 0 ok / OraclePanic code. stats = {num_queued, num_visited, num_pruned, end_node, end_offset}
int oracle_align(const poasta_graph_view *view, const uint8_t *seq, uint32_t qlen, int x, int o, int e, int mode,
                 int heuristic, int flags, uint32_t *score, uint32_t *rpos, uint32_t *qpos, uint32_t cap, uint32_t *n_pairs,
                 uint64_t *stats) {
    AlignOptions opt;
    opt.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode;
    opt.heuristic = heuristic;
    opt.mock = flags & 1;
    opt.pruning = !(flags & 2);
    AstarResult r = astar_align(*view, seq, qlen, opt);
    *score = r.score;
    *n_pairs = (uint32_t)r.rpos.size();
    for (size_t i = 0; i < r.rpos.size() && i < cap; ++i) {
        rpos[i] = r.rpos[i];
        qpos[i] = r.qpos[i];
    }
    if (stats) {
        stats[0] = r.num_queued, stats[1] = r.num_visited, stats[2] = r.num_pruned;
        stats[3] = r.end_node, stats[4] = r.end_offset;
    }
    return r.status;
}

// Dump of the A* visited table: quadruples (node, offset, state[0=M,1=D,2=I], score). Returns count of quadruples.
uint64_t oracle_align_dump(const poasta_graph_view *view, const uint8_t *seq, uint32_t qlen, int x, int o, int e, int mode,
                           int heuristic, int flags, uint32_t *out, uint64_t cap_quads) {
    AlignOptions opt;
    opt.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode;
    opt.heuristic = heuristic;
    opt.mock = flags & 1;
    opt.pruning = !(flags & 2);
    std::vector<uint32_t> dump;
    opt.dump = &dump;
    astar_align(*view, seq, qlen, opt);
    uint64_t n = dump.size() / 4;
    for (uint64_t i = 0; i < n && i < cap_quads; ++i) memcpy(out + 4 * i, dump.data() + 4 * i, 16);
    return n;
}

uint32_t oracle_rev_postorder(const poasta_graph_view *view, int mock, uint32_t *out) {
    Graph g(*view, mock);
    auto v = rev_postorder_nodes(g);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (uint32_t)v.size();
}

uint32_t oracle_superbubbles(const poasta_graph_view *view, int mock, uint32_t *out_pairs, uint32_t cap) {
    Graph g(*view, mock);
    SuperbubbleFinder f(g);
    auto b = f.bubbles(g);
    for (size_t i = 0; i < b.size() && i < cap; ++i) {
        out_pairs[2 * i] = b[i].first;
        out_pairs[2 * i + 1] = b[i].second;
    }
    return (uint32_t)b.size();
}

// out_maps: per node up to max_per_node triples (exit, min_dist, max_dist); counts[node]; dist[2*node..]
void oracle_bubble_index(const poasta_graph_view *view, int mock, uint32_t max_per_node, uint32_t *counts, uint32_t *out_maps,
                         uint32_t *dist) {
    Graph g(*view, mock);
    BubbleIndex bi(g);
    for (uint32_t n = 0; n < g.n(); ++n) {
        counts[n] = (uint32_t)bi.node_bubble_map[n].size();
        for (uint32_t k = 0; k < counts[n] && k < max_per_node; ++k) {
            const auto &m = bi.node_bubble_map[n][k];
            uint32_t *o = out_maps + 3 * ((size_t)n * max_per_node + k);
            o[0] = m.bubble_exit, o[1] = (uint32_t)m.min_dist_to_exit, o[2] = (uint32_t)m.max_dist_to_exit;
        }
        dist[2 * n] = (uint32_t)bi.dist_to_end[n].first;
        dist[2 * n + 1] = (uint32_t)bi.dist_to_end[n].second;
    }
}

uint64_t oracle_heuristic(const poasta_graph_view *view, int mock, int kind, int x, int o, int e, uint32_t seq_len,
                          uint32_t node, uint32_t offset, int state, int max_paths_per_node, int max_paths_to_consider) {
    Graph g(*view, mock);
    BubbleIndex bi(g);
    Heuristic h;
    h.kind = kind;
    h.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    h.bubbles = &bi;
    h.seq_len = seq_len;
    std::unique_ptr<PathIndex> pi;
    if (kind == 2) {
        pi.reset(new PathIndex(g, max_paths_per_node));
        h.paths = pi.get();
        h.max_paths_to_consider = max_paths_to_consider;
    }
    return h.h(node, offset, (State)state);
}

uint64_t oracle_gap_cost(int x, int o, int e, int state, uint64_t len) {
    return Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e}.gap_cost((State)state, len);
}

// aligner/queue.rs:110-136 replayed on the bucket queue; returns 0 when every assertion holds, else the failing step.
int oracle_queue_selftest() {
    struct VecLayer {
        std::vector<size_t> v;
        bool empty() const { return v.empty(); }
    };
    LayeredQueueT<VecLayer> q;
    auto queue = [&](size_t item, size_t pr) { q.queue_at(pr, [&](VecLayer &l) { l.v.push_back(item); }); };
    auto pop = [&]() -> long {
        if (q.layers.empty()) return -1;
        long r = -1;
        if (!q.layers[0].v.empty()) {
            r = (long)q.layers[0].v.back();
            q.layers[0].v.pop_back();
        }
        q.drop_empty_front();
        return r;
    };
    queue(10, 10);
    if (q.layers.size() != 1 || q.layer_min != 10) return 1;
    queue(5, 5);
    if (q.layers.size() != 6 || q.layer_min != 5) return 2;
    if (pop() != 5) return 3;
    if (q.layers.size() != 1 || q.layer_min != 10) return 4;
    queue(15, 15);
    if (q.layers.size() != 6 || q.layer_min != 10) return 5;
    queue(12, 12);
    if (q.layers.size() != 6 || q.layer_min != 10) return 6;
    if (q.layers[2].v != std::vector<size_t>{12}) return 7;
    return 0;
}

// aligner/dfa.rs:420-482: the four DFA unit tests with the reference's DummyVisited.
// out = {kind, parent_node, parent_offset, child_node, child_offset, num_visited, num_pruned}
void oracle_dfa_event(const poasta_graph_view *view, const uint8_t *symbols_override, const uint8_t *seq, uint32_t qlen,
                      uint32_t node, uint32_t offset, int prune_next, uint32_t *out) {
    struct Dummy {
        bool prune_next;
        bool update_score_if_lower(uint32_t, uint32_t, State, uint32_t) { return true; }
        void dfa_match(uint32_t, uint32_t) {}
        bool prune(uint32_t, uint32_t, uint32_t, State) { return prune_next; }
    } vis{prune_next != 0};
    // SymbolMockGraph::is_symbol_equal has no end-node wildcard; the reference test gives the end node 'N',
    // and the DFA tests `child == end_node` before any symbol comparison, so the wildcard is never consulted.
    poasta_graph_view v = *view;
    v.symbol = symbols_override;
    Graph g(v, false);
    DepthFirstGreedyAlignment<Dummy> dfa(g, seq, qlen, 0, node, offset);
    ExtendResult r = dfa.extend(vis);
    out[0] = (uint32_t)r.kind;
    out[1] = r.parent_node, out[2] = r.parent_offset, out[3] = r.child_node, out[4] = r.child_offset;
    out[5] = (uint32_t)dfa.num_visited, out[6] = (uint32_t)dfa.num_pruned;
}

// ---- progressive MSA with the oracle aligner (bin/poasta.rs:163-236) ------------------------------
// fmt: 0 fasta, 1 gfa, 2 gfav1, 3 poasta. Returns 0 or negative error; *out is malloc'ed.
int oracle_msa(const char *fasta, size_t fasta_len, const uint8_t *graph_in, size_t graph_in_len, int x, int o, int e, int mode,
               int heuristic, int fmt, char **out, size_t *out_len, uint32_t *scores, uint32_t scores_cap, uint64_t *cells,
               uint64_t *stats3) {
    using namespace poasta;
    POAGraph g;
    if (graph_in && graph_in_len) {
        if (!graph_from_poasta(graph_in, graph_in_len, g)) return -8;
    }
    auto recs = parse_fasta(fasta, fasta_len);
    AlignOptions opt;
    opt.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode;
    opt.heuristic = heuristic;
    FlatGraph flat;
    uint64_t ncells = 0;
    uint64_t st[3] = {0, 0, 0};
    for (size_t i = 0; i < recs.size(); ++i) {
        const auto &rec = recs[i];
        const uint8_t *s = (const uint8_t *)rec.seq.data();
        uint32_t sc = 0;
        GraphError ge;
        if (g.is_empty()) {
            ge = g.add_alignment_with_weights(rec.name, s, rec.seq.size(), false, nullptr, 0);
        } else {
            flat.build(g);
            ncells += (uint64_t)g.node_count() * rec.seq.size();
            AstarResult r = astar_align(flat.view(), s, rec.seq.size(), opt);
            if (r.status) return -100 - r.status;
            sc = r.score;
            st[0] += r.num_queued, st[1] += r.num_visited, st[2] += r.num_pruned;
            std::vector<AlignedPair> aln(r.rpos.size());
            for (size_t k = 0; k < aln.size(); ++k) aln[k] = {r.rpos[k], r.qpos[k]};
            ge = g.add_alignment_with_weights(rec.name, s, rec.seq.size(), true, aln.data(), aln.size());
        }
        if (ge != GraphError::Ok) return -(int)ge - 2;
        if (scores && i < scores_cap) scores[i] = sc;
    }
    if (cells) *cells = ncells;
    if (stats3) memcpy(stats3, st, sizeof st);
    if (out) {
        std::string buf;
        bool ok = true;
        if (fmt == 0) ok = graph_to_fasta(g, buf);
        else if (fmt == 1) ok = graph_to_gfa(g, buf, false);
        else if (fmt == 2) ok = graph_to_gfa(g, buf, true);
        else graph_to_poasta(g, buf);
        if (!ok) return -8;
        *out = (char *)malloc(buf.size() + 1);
        memcpy(*out, buf.data(), buf.size());
        (*out)[buf.size()] = 0;
        *out_len = buf.size();
    }
    return 0;
}

// Many families over a thread pool: the CPU baseline leg of bench.py ("port" of the reference, one family per task,
// like `xargs -P $(nproc) poasta align`). Returns wall seconds; total cells and sequences via out params.
double oracle_msa_many(const char *const *fasta, const size_t *fasta_len, uint32_t n_families, int x, int o, int e, int mode,
                       int heuristic, int n_threads, uint64_t *cells_out, uint64_t *seqs_out, int *err_out) {
    std::vector<uint64_t> cells(n_families, 0), seqs(n_families, 0);
    std::vector<int> errs(n_families, 0);
    std::atomic<uint32_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            uint32_t f = next.fetch_add(1);
            if (f >= n_families) break;
            uint32_t sc[4096];
            errs[f] = oracle_msa(fasta[f], fasta_len[f], nullptr, 0, x, o, e, mode, heuristic, 0, nullptr, nullptr, sc, 4096,
                                 &cells[f], nullptr);
            size_t n = 0;
            for (size_t k = 0; k < fasta_len[f]; ++k) n += fasta[f][k] == '>';
            seqs[f] = n;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t c = 0, s = 0;
    int err = 0;
    for (uint32_t f = 0; f < n_families; ++f) {
        c += cells[f];
        s += seqs[f];
        if (errs[f]) err = errs[f];
    }
    *cells_out = c;
    *seqs_out = s;
    *err_out = err;
    return secs;
}

// Many independent reads against ONE fixed graph over a thread pool (config 5, library semantics): PoastaAligner::align per read,
// which rebuilds the bubble index (and the path index for -H path) for every call (aligner/config.rs:113-134, 296-308).
// Returns wall seconds; scores (optional) per read, total cells.
double oracle_align_many(const poasta_graph_view *view, const char *const *seqs, const uint32_t *lens, uint32_t n, int x, int o, int e, int mode,
                         int heuristic, int n_threads, uint32_t *scores, uint64_t *cells_out, int *err_out) {
    std::atomic<uint32_t> next{0};
    std::atomic<int> err{0};
    AlignOptions opt;
    opt.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode, opt.heuristic = heuristic;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            AstarResult r = astar_align(*view, (const uint8_t *)seqs[i], lens[i], opt);
            if (r.status) err = r.status;
            if (scores) scores[i] = r.score;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t c = 0;
    for (uint32_t i = 0; i < n; ++i) c += (uint64_t)(view->n_nodes - 2) * lens[i];
    *cells_out = c;
    *err_out = err;
    return secs;
}

// Host cost of graph insertion (add_alignment_with_weights + post_process) and CSR flattening for one family: the MSA is built
// once with the oracle to obtain the alignments, then the insertions are replayed `reps` times. Microseconds per sequence.
int host_insert_timing(const char *fasta, size_t fasta_len, int reps, double *us_insert, double *us_flatten) {
    using namespace poasta;
    auto recs = parse_fasta(fasta, fasta_len);
    AlignOptions opt;
    opt.costs = Costs{4, 6, 2};
    opt.mode = 0, opt.heuristic = 1;
    std::vector<std::vector<AlignedPair>> alns(recs.size());
    {
        POAGraph g;
        FlatGraph flat;
        for (size_t i = 0; i < recs.size(); ++i) {
            const uint8_t *s = (const uint8_t *)recs[i].seq.data();
            if (!g.is_empty()) {
                flat.build(g);
                AstarResult r = astar_align(flat.view(), s, recs[i].seq.size(), opt);
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

// Dense DP model (dp_model.hpp): global mode. Returns score (0x3FFFFFFF when unreachable).
uint32_t dpmodel_align(const poasta_graph_view *view, const uint8_t *seq, uint32_t qlen, int x, int o, int e, uint32_t *rpos,
                       uint32_t *qpos, uint32_t cap, uint32_t *n_pairs, uint32_t *tables_out) {
    std::vector<uint32_t> r, q;
    dpmodel::Tables t;
    uint32_t score = dpmodel::align_global(*view, seq, qlen, x, o, e, r, q, &t);
    *n_pairs = (uint32_t)r.size();
    for (size_t i = 0; i < r.size() && i < cap; ++i) {
        rpos[i] = r[i];
        qpos[i] = q[i];
    }
    if (tables_out) {  // [3][V][Q+1] in M, D, I order (node-id major)
        size_t n = (size_t)t.V * (t.Q + 1);
        memcpy(tables_out, t.M.data(), n * 4);
        memcpy(tables_out + n, t.D.data(), n * 4);
        memcpy(tables_out + 2 * n, t.I.data(), n * 4);
    }
    return score;
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
    Scratch s{planes.data(), epoch, qhead.data(), qwin.data(), items.data(), free_slots.data(), rb.data(), dfa.data(), d.W, max_items, d.f_max};
    Params p{(uint32_t)x, (uint32_t)o, (uint32_t)e, mode, heuristic};
    const uint32_t ocap = view->n_nodes + qlen + 2;
    std::vector<uint32_t> out_r(ocap), out_q(ocap);
    Result r;
    uint32_t used_blocks = 0, used_items = 0;
    auto go = [&](auto cell, auto heur) {
        rp::Search<decltype(cell), decltype(heur)::value> search(g, seq, qlen, p, s);
        r = search.run(out_r.data(), out_q.data(), ocap);
        used_items = search.n_items;
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

void oracle_free(void *p) { free(p); }

}  // extern "C"
