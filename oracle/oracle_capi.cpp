// ORACLE — TEST INFRASTRUCTURE ONLY (see astar_oracle.hpp header).  C entry points for ctypes.
// Everything here is the oracle's own restatement: the A* path (astar_oracle.hpp), the graph / writers side (poa_oracle.hpp) and
// the dense model of the DP recurrences (dp_model.hpp). NO product header is included (only the public C header for the
// poasta_graph_view struct); host builds of product code for CPU-side checks live in tests/hooks/.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "astar_oracle.hpp"
#include "dp_model.hpp"
#include "poa_oracle.hpp"

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
    poa::POAGraph g;
    if (graph_in && graph_in_len) {
        if (!poa::load_graph(graph_in, graph_in_len, g)) return -8;
    }
    const auto recs = poa::read_fasta(fasta, fasta_len);
    AlignOptions opt;
    opt.costs = Costs{(uint8_t)x, (uint8_t)o, (uint8_t)e};
    opt.mode = mode;
    opt.heuristic = heuristic;
    uint64_t ncells = 0;
    uint64_t st[3] = {0, 0, 0};
    for (size_t i = 0; i < recs.size(); ++i) {
        const auto &rec = recs[i];
        const std::vector<uint64_t> weights(rec.seq.size(), 1);  // bin/poasta.rs:193
        uint32_t sc = 0;
        poa::Err ge;
        if (g.is_empty()) {
            ge = g.add_alignment_with_weights(rec.name, rec.seq, nullptr, weights);
        } else {
            const poa::Flat flat(g);
            ncells += (uint64_t)g.node_count() * rec.seq.size();
            AstarResult r = astar_align(flat.view, (const uint8_t *)rec.seq.data(), rec.seq.size(), opt);
            if (r.status) return -100 - r.status;
            sc = r.score;
            st[0] += r.num_queued, st[1] += r.num_visited, st[2] += r.num_pruned;
            std::vector<poa::Pair> aln(r.rpos.size());
            for (size_t k = 0; k < aln.size(); ++k) aln[k] = {r.rpos[k], r.qpos[k]};
            ge = g.add_alignment_with_weights(rec.name, rec.seq, &aln, weights);
        }
        if (ge != poa::Err::Ok) return -(int)ge - 2;
        if (scores && i < scores_cap) scores[i] = sc;
    }
    if (cells) *cells = ncells;
    if (stats3) memcpy(stats3, st, sizeof st);
    if (out) {
        std::string buf;
        bool ok = true;
        if (fmt == 0) ok = poa::graph_to_fasta(g, buf);
        else if (fmt == 1) ok = poa::graph_to_gfa(g, buf, false);
        else if (fmt == 2) ok = poa::graph_to_gfa(g, buf, true);
        else poa::save_graph(g, buf);
        if (!ok) return -8;
        *out = (char *)malloc(buf.size() + 1);
        memcpy(*out, buf.data(), buf.size());
        (*out)[buf.size()] = 0;
        *out_len = buf.size();
    }
    return 0;
}

// The oracle's graph side alone, for tests that pin the PRODUCT's graph / writers against it: replays a list of insertions
// (name, sequence, alignment or None) and returns the requested output. pairs: concatenated (rpos, qpos); n_pairs[i] < 0 = None.
int oracle_graph_build(const char *const *names, const char *const *seqs, const uint32_t *seq_lens, const uint32_t *pairs, const int64_t *n_pairs,
                       uint32_t n, const uint8_t *graph_in, size_t graph_in_len, const char *msa_in, size_t msa_in_len, int fmt, char **out,
                       size_t *out_len, int32_t *insert_status) {
    poa::POAGraph g;
    if (graph_in && graph_in_len && !poa::load_graph(graph_in, graph_in_len, g)) return -8;
    if (msa_in && msa_in_len && !poa::graph_from_fasta_msa(msa_in, msa_in_len, g)) return -5;
    size_t at = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const std::string seq(seqs[i], seq_lens[i]);
        const std::vector<uint64_t> weights(seq.size(), 1);
        poa::Err ge;
        if (n_pairs[i] < 0) {
            ge = g.add_alignment_with_weights(names[i], seq, nullptr, weights);
        } else {
            std::vector<poa::Pair> aln((size_t)n_pairs[i]);
            for (auto &p : aln) p = {pairs[2 * at], pairs[2 * at + 1]}, ++at;
            ge = g.add_alignment_with_weights(names[i], seq, &aln, weights);
        }
        if (insert_status) insert_status[i] = (int32_t)ge;
        if (ge != poa::Err::Ok) return -(int)ge - 2;
    }
    std::string buf;
    bool ok = true;
    if (fmt == 0) ok = poa::graph_to_fasta(g, buf);
    else if (fmt == 1) ok = poa::graph_to_gfa(g, buf, false);
    else if (fmt == 2) ok = poa::graph_to_gfa(g, buf, true);
    else poa::save_graph(g, buf);
    if (!ok) return -8;
    *out = (char *)malloc(buf.size() + 1);
    memcpy(*out, buf.data(), buf.size());
    (*out)[buf.size()] = 0;
    *out_len = buf.size();
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

void oracle_free(void *p) { free(p); }

}  // extern "C"
