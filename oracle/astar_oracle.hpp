// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under poasta_b200/ may
// include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it (as the checker / the timed CPU baseline).
//
// CPU restatement of POASTA's gap-affine A* alignment path, written from the reference's
// behaviour (no source copied).  Each block cites the reference file:line it follows
// (paths relative to /root/reference).
//
// PARITY STATUS: "parity unpinned" for whole-MSA output.  The reference is Rust; no cargo/rustc
// exists in this environment and its 102 crates are not vendored, so the reference cannot be
// compiled or run here.  What IS pinned: every exact-value unit test the reference holds for this
// path (bubble sets, bubble maps, dist-to-end tables, heuristic values, queue behaviour,
// rev-postorder, DFA events, FASTA strings, exact-match scores) — see tests/test_oracle_*.py.
// Behaviour that lives in un-vendored crates (petgraph 0.6.5 adjacency order / toposort,
// noodles-fasta 0.45 line wrapping, bincode 1.3.3 layout) is restated from the crates' published
// behaviour.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <tuple>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/poasta_b200.h"

namespace oracle {

constexpr uint32_t UNVISITED = 0xFFFFFFFFu;  // Score::Unvisited (aligner/scoring/mod.rs:65-70): greater than any score
constexpr uint32_t NONE = 0xFFFFFFFFu;

enum State : uint8_t { MATCH = 0, DELETION = 1, INSERTION = 2 };  // aligner/aln_graph.rs:7-14 (order M < D < I)

struct OraclePanic {
    int code;  // 1 = empty queue (astar.rs:143), 2 = no backtrace (gap_affine.rs:853), 3 = NonMax unwrap on wrapped score
};

// ---- graph access: the AlignableRefGraph trait over a CSR view (graphs/mod.rs:23-53) ----------
struct Graph {
    poasta_graph_view v;
    bool mock;  // graphs/mock.rs: is_symbol_equal == false always, start = first node, end = last node
    std::vector<uint32_t> rank;  // get_node_ordering()

    Graph(const poasta_graph_view &view, bool mock_ = false) : v(view), mock(mock_), rank(view.n_nodes, 0) {
        for (uint32_t r = 0; r < v.n_nodes; ++r) rank[v.topo_order[r]] = r;
    }
    uint32_t n() const { return v.n_nodes; }
    uint32_t start() const { return v.start_node; }
    uint32_t end() const { return v.end_node; }
    uint32_t node_count() const { return mock ? v.n_nodes : v.n_nodes - 2; }
    const uint32_t *succ_begin(uint32_t x) const { return v.succ + v.succ_off[x]; }
    const uint32_t *succ_end(uint32_t x) const { return v.succ + v.succ_off[x + 1]; }
    const uint32_t *pred_begin(uint32_t x) const { return v.pred + v.pred_off[x]; }
    const uint32_t *pred_end(uint32_t x) const { return v.pred + v.pred_off[x + 1]; }
    uint32_t out_degree(uint32_t x) const { return v.succ_off[x + 1] - v.succ_off[x]; }
    uint32_t in_degree(uint32_t x) const { return v.pred_off[x + 1] - v.pred_off[x]; }
    bool is_symbol_equal(uint32_t node, uint8_t s) const {  // graphs/poa.rs:463-465, mock.rs:77-79
        if (mock) return false;
        return node == v.end_node || v.symbol[node] == s;
    }
};

// ---- graphs/tools.rs:5-37 ----------------------------------------------------------------------
inline std::vector<uint32_t> rev_postorder_nodes(const Graph &g) {
    std::vector<uint32_t> ordered;
    ordered.reserve(g.n());
    std::vector<std::pair<uint32_t, const uint32_t *>> stack;
    stack.emplace_back(g.start(), g.succ_begin(g.start()));
    std::vector<uint8_t> visited(g.n(), 0);
    while (!stack.empty()) {
        auto &top = stack.back();
        const uint32_t *e = g.succ_end(top.first);
        uint32_t child = NONE;
        while (top.second != e) {
            uint32_t c = *top.second++;
            if (!visited[c]) {
                child = c;
                break;
            }
        }
        if (child != NONE) {
            visited[child] = 1;
            stack.emplace_back(child, g.succ_begin(child));
        } else {
            ordered.push_back(top.first);
            stack.pop_back();
        }
    }
    std::reverse(ordered.begin(), ordered.end());
    return ordered;
}

// ---- bubbles/finder.rs:30-177 --------------------------------------------------------------------
struct SuperbubbleFinder {
    std::vector<uint32_t> inv_rev_postorder;  // rank -> node
    std::vector<int64_t> rev_postorder;       // node -> rank
    std::vector<int64_t> out_parent, out_child;

    explicit SuperbubbleFinder(const Graph &g) {
        inv_rev_postorder = rev_postorder_nodes(g);
        rev_postorder.assign(g.n(), 0);
        for (size_t i = 0; i < inv_rev_postorder.size(); ++i) rev_postorder[inv_rev_postorder[i]] = (int64_t)i;
        out_parent.assign(g.n(), -1);
        out_child.assign(g.n(), INT64_MAX);
        for (uint32_t n = 0; n < g.n(); ++n) {
            int64_t mn = INT64_MAX;
            bool any = false;
            for (const uint32_t *p = g.pred_begin(n); p != g.pred_end(n); ++p) {
                mn = std::min(mn, rev_postorder[*p]);
                any = true;
            }
            out_parent[n] = any ? mn : -1;
            int64_t mx = INT64_MIN;
            any = false;
            for (const uint32_t *s = g.succ_begin(n); s != g.succ_end(n); ++s) {
                mx = std::max(mx, rev_postorder[*s]);
                any = true;
            }
            out_child[n] = any ? mx : INT64_MAX;
        }
    }

    // finder.rs:109-177: yields (entrance, exit) pairs in the iterator's order.
    std::vector<std::pair<uint32_t, uint32_t>> bubbles(const Graph &g) const {
        std::vector<std::pair<uint32_t, uint32_t>> out;
        std::vector<int64_t> opm(g.n(), 0);  // out_parent_map (every key is written before it is read)
        std::vector<uint32_t> stack;
        uint32_t candidate_exit = NONE;
        auto pop_stack = [&]() -> uint32_t {
            if (stack.empty()) return NONE;
            uint32_t c = stack.back();
            stack.pop_back();
            return c;
        };
        for (int64_t curr = (int64_t)g.n() - 1; curr >= 0; --curr) {
            if ((size_t)curr >= inv_rev_postorder.size()) continue;  // unreachable nodes cannot occur after post_process
            uint32_t n = inv_rev_postorder[curr];
            int64_t furthest_child = out_child[n];
            bool found = false;
            std::pair<uint32_t, uint32_t> ret;
            if (furthest_child == curr + 1) {
                if (candidate_exit != NONE) stack.push_back(candidate_exit);
                candidate_exit = inv_rev_postorder[curr + 1];
            } else {
                while (candidate_exit != NONE) {
                    uint32_t candidate = candidate_exit;
                    if (furthest_child <= rev_postorder[candidate]) break;
                    candidate_exit = pop_stack();
                    if (candidate_exit != NONE) opm[candidate_exit] = std::min(opm[candidate], opm[candidate_exit]);
                }
            }
            if (candidate_exit != NONE) {
                uint32_t candidate = candidate_exit;
                if ((uint64_t)opm[candidate] == (uint64_t)curr) {
                    ret = {n, candidate};
                    found = true;
                    candidate_exit = pop_stack();
                    if (candidate_exit != NONE) opm[candidate_exit] = std::min(opm[candidate], opm[candidate_exit]);
                }
            }
            opm[n] = out_parent[n];
            if (candidate_exit != NONE) opm[candidate_exit] = std::min(opm[n], opm[candidate_exit]);
            if (found) out.push_back(ret);
        }
        return out;
    }
};

// ---- bubbles/index.rs:51-156 ---------------------------------------------------------------------
struct NodeBubbleMap {
    uint32_t bubble_exit;
    size_t min_dist_to_exit, max_dist_to_exit;
};

struct BubbleIndex {
    std::vector<uint32_t> entrance_of;  // node -> exit if node is an entrance else NONE
    std::vector<uint32_t> exit_of;      // node -> entrance if node is an exit else NONE
    std::vector<std::vector<NodeBubbleMap>> node_bubble_map;
    std::vector<std::pair<size_t, size_t>> dist_to_end;

    explicit BubbleIndex(const Graph &g) {
        SuperbubbleFinder finder(g);
        const uint32_t n = g.n();
        entrance_of.assign(n, NONE);
        exit_of.assign(n, NONE);
        for (auto &b : finder.bubbles(g)) {
            entrance_of[b.first] = b.second;
            exit_of[b.second] = b.first;
        }
        node_bubble_map.assign(n, {});
        dist_to_end.assign(n, {0, 0});
        typedef std::vector<std::pair<size_t, uint32_t>> BStack;
        BStack init;
        if (exit_of[g.end()] != NONE) init.emplace_back(0, g.end());
        std::deque<std::tuple<uint32_t, size_t, BStack>> queue;
        queue.emplace_back(g.end(), 0, init);
        std::vector<uint8_t> visited(n, 0);
        visited[g.end()] = 1;
        while (!queue.empty()) {
            auto [curr, dist_from_end, bubble_stack] = std::move(queue.front());
            queue.pop_front();
            for (auto &b : bubble_stack)
                node_bubble_map[curr].push_back({b.second, dist_from_end - b.first, 0});
            dist_to_end[curr].first = dist_from_end;
            for (const uint32_t *pp = g.pred_begin(curr); pp != g.pred_end(curr); ++pp) {
                uint32_t pred = *pp;
                if (visited[pred]) continue;
                size_t nd = dist_from_end + 1;
                BStack ns = bubble_stack;
                if (entrance_of[pred] != NONE) {
                    auto top = ns.back();  // reference: pop().unwrap()
                    ns.pop_back();
                    node_bubble_map[pred].push_back({top.second, nd - top.first, 0});
                }
                if (exit_of[pred] != NONE) ns.emplace_back(nd, pred);
                visited[pred] = 1;
                queue.emplace_back(pred, nd, std::move(ns));
            }
        }
        for (size_t i = finder.inv_rev_postorder.size(); i-- > 0;) {
            uint32_t nn = finder.inv_rev_postorder[i];
            size_t mx = 0;
            for (const uint32_t *s = g.succ_begin(nn); s != g.succ_end(nn); ++s) mx = std::max(mx, dist_to_end[*s].second + 1);
            dist_to_end[nn].second = mx;
            for (auto &b : node_bubble_map[nn]) b.max_dist_to_exit = mx - dist_to_end[b.bubble_exit].second;
        }
    }
    bool is_exit(uint32_t n) const { return exit_of[n] != NONE; }
    bool is_entrance(uint32_t n) const { return entrance_of[n] != NONE; }
};

// ---- costs: aligner/scoring/gap_affine.rs:19-81 ---------------------------------------------------
struct Costs {
    uint8_t x, o, e;
    size_t gap_cost(State st, size_t len) const {
        if (len == 0) return 0;
        size_t open = (st == MATCH) ? o : 0;
        return open + len * (size_t)e;
    }
    uint8_t open_extend() const { return (uint8_t)(o + e); }  // u8 sum, wraps like the release build (gap_affine.rs:288)
};

// ---- aligner/path_index.rs:48-267 -----------------------------------------------------------------
struct PathIndex {
    std::vector<std::vector<uint32_t>> paths;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> node_to_paths;
    size_t max_paths_per_node;

    PathIndex(const Graph &g, size_t max_per_node) : node_to_paths(g.n()), max_paths_per_node(max_per_node) {
        std::set<std::pair<uint32_t, uint32_t>> visited_edges;
        std::vector<uint8_t> has_visited_out(g.n(), 0);
        std::vector<uint32_t> start_nodes{g.start()};
        for (uint32_t node = 0; node < g.n(); ++node) {
            uint32_t in = g.in_degree(node), out = g.out_degree(node);
            if (in == 0 || (out > 2 && in == 1)) start_nodes.push_back(node);
        }
        auto extract = [&](uint32_t start) {
            std::vector<uint32_t> nodes{start};
            uint32_t current = start;
            while (current != g.end()) {
                const uint32_t *b = g.succ_begin(current), *e = g.succ_end(current);
                if (b == e) break;
                // filter unvisited, max_by_key(out_degree): the LAST maximum wins; else first neighbour
                uint32_t next = NONE;
                uint32_t best = 0;
                for (const uint32_t *s = b; s != e; ++s) {
                    if (visited_edges.count({current, *s})) continue;
                    uint32_t od = g.out_degree(*s);
                    if (next == NONE || od >= best) {
                        next = *s;
                        best = od;
                    }
                }
                if (next == NONE) next = *b;
                visited_edges.insert({current, next});
                has_visited_out[current] = 1;
                nodes.push_back(next);
                current = next;
            }
            return nodes;
        };
        auto add_path = [&](std::vector<uint32_t> nodes) {
            uint32_t id = (uint32_t)paths.size();
            for (size_t pos = 0; pos < nodes.size(); ++pos) {
                auto &entry = node_to_paths[nodes[pos]];
                if (entry.size() < max_paths_per_node) entry.emplace_back(id, (uint32_t)pos);
            }
            paths.push_back(std::move(nodes));
        };
        for (uint32_t s : start_nodes) {
            if (has_visited_out[s]) continue;
            auto p = extract(s);
            if (p.size() > 1) add_path(std::move(p));
        }
        if (paths.size() < 10) {
            std::vector<std::pair<uint32_t, size_t>> cands;
            for (uint32_t node = 0; node < g.n(); ++node) {
                size_t cnt = 0;
                for (const uint32_t *s = g.succ_begin(node); s != g.succ_end(node); ++s)
                    if (!visited_edges.count({node, *s})) ++cnt;
                if (cnt > 0) cands.emplace_back(node, cnt);
            }
            std::stable_sort(cands.begin(), cands.end(), [](auto &a, auto &b) { return a.second > b.second; });
            for (size_t i = 0; i < cands.size() && i < 20; ++i) {
                auto p = extract(cands[i].first);
                if (p.size() > 3) add_path(std::move(p));
            }
        }
    }
    uint32_t distance_to_end(uint32_t path_id, uint32_t pos) const { return (uint32_t)paths[path_id].size() - 1 - pos; }
};

// ---- aligner/heuristic.rs:40-184 --------------------------------------------------------------------
struct Heuristic {
    int kind;  // 0 dijkstra, 1 mingap, 2 path
    Costs costs;
    const BubbleIndex *bubbles = nullptr;
    const PathIndex *paths = nullptr;
    size_t seq_len = 0;
    size_t max_paths_to_consider = 5;

    size_t h(uint32_t node, uint32_t offset, State st) const {
        if (kind == 0) return 0;
        if (kind == 1) {
            size_t a = bubbles->dist_to_end[node].first, b = bubbles->dist_to_end[node].second;
            size_t min_d = a ? a - 1 : 0, max_d = b ? b - 1 : 0;
            size_t tmin = offset + min_d, tmax = offset + max_d;
            size_t gap;
            if (tmin > seq_len) {
                gap = tmin - seq_len;
                if (st != DELETION) st = MATCH;
            } else if (tmax < seq_len) {
                gap = seq_len - tmax;
                if (st != INSERTION) st = MATCH;
            } else {
                gap = 0;
            }
            return costs.gap_cost(st, gap);
        }
        const auto &ps = paths->node_to_paths[node];
        size_t rem_q = seq_len > offset ? seq_len - offset : 0;
        if (ps.empty()) return costs.gap_cost(st, rem_q);
        size_t ncheck = std::min(ps.size(), max_paths_to_consider);
        size_t best = SIZE_MAX;
        for (size_t i = 0; i < ncheck; ++i) {
            size_t rem_p = paths->distance_to_end(ps[i].first, ps[i].second);
            size_t cost;
            if (rem_p > rem_q) cost = costs.gap_cost(st == DELETION ? DELETION : MATCH, rem_p - rem_q);
            else if (rem_q > rem_p) cost = costs.gap_cost(st == INSERTION ? INSERTION : MATCH, rem_q - rem_p);
            else cost = 0;
            best = std::min(best, cost);
        }
        return best;
    }
};

// ---- aligner/queue.rs:19-82 + gap_affine.rs:929-1013 ------------------------------------------------
struct QueueItem {
    uint32_t score, node, offset;
    State state;
};
struct Layer {
    std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> m, i, d;
    bool empty() const { return m.empty() && i.empty() && d.empty(); }
};
template <class L> struct LayeredQueueT {
    std::deque<L> layers;
    size_t layer_min = 0;
    template <class Push> void queue_at(size_t priority, Push &&push) {
        if (layers.empty()) {
            layers.emplace_back();
            layer_min = priority;
        } else {
            size_t layer_max = layer_min + layers.size();
            if (priority < layer_min) {
                size_t diff = layer_min - priority;
                for (size_t k = 0; k < diff; ++k) layers.emplace_front();
                layer_min = priority;
            } else if (priority >= layer_max) {
                layers.resize(priority - layer_min + 1);
            }
        }
        push(layers[priority - layer_min]);
    }
    void drop_empty_front() {
        while (!layers.empty()) {
            if (layers[0].empty()) {
                layers.pop_front();
                ++layer_min;
            } else {
                break;
            }
        }
    }
};
struct AffineQueue : LayeredQueueT<Layer> {
    void queue(uint32_t node, uint32_t offset, State st, uint32_t score, size_t h) {
        size_t priority = (size_t)score + h;
        queue_at(priority, [&](Layer &l) {
            auto item = std::make_pair(score, std::make_pair(node, offset));
            if (st == MATCH) l.m.push_back(item);
            else if (st == INSERTION) l.i.push_back(item);
            else l.d.push_back(item);
        });
    }
    bool pop(QueueItem &out) {
        if (layers.empty()) return false;
        Layer &l = layers[0];
        bool got = true;
        if (!l.m.empty()) {
            out = {l.m.back().first, l.m.back().second.first, l.m.back().second.second, MATCH};
            l.m.pop_back();
        } else if (!l.d.empty()) {
            out = {l.d.back().first, l.d.back().second.first, l.d.back().second.second, DELETION};
            l.d.pop_back();
        } else if (!l.i.empty()) {
            out = {l.i.back().first, l.i.back().second.first, l.i.back().second.second, INSERTION};
            l.i.pop_back();
        } else {
            got = false;
        }
        drop_empty_front();
        return got;
    }
};

// ---- visited storage: gap_affine.rs:442-548 (8x8 blocked sparse table keyed by rank/offset blocks) ---
struct VisitedStore {
    struct Cell {
        uint32_t m, i, d;
    };
    struct Block {
        Cell c[8][8];
    };
    // open-addressing map (rank_block, offset_block) -> block index; iteration order is never used
    std::vector<uint64_t> keys;
    std::vector<uint32_t> vals;
    std::vector<Block> blocks;
    size_t mask = 0, used = 0;
    const Graph *g = nullptr;

    explicit VisitedStore(const Graph &graph) : g(&graph) { rehash(1024); }
    void rehash(size_t cap) {
        std::vector<uint64_t> ok = std::move(keys);
        std::vector<uint32_t> ov = std::move(vals);
        keys.assign(cap, UINT64_MAX);
        vals.assign(cap, 0);
        mask = cap - 1;
        for (size_t k = 0; k < ok.size(); ++k)
            if (ok[k] != UINT64_MAX) {
                size_t h = hash(ok[k]);
                while (keys[h] != UINT64_MAX) h = (h + 1) & mask;
                keys[h] = ok[k];
                vals[h] = ov[k];
            }
    }
    size_t hash(uint64_t k) const { return (size_t)((k * 0x9E3779B97F4A7C15ull) >> 20) & mask; }
    Block *find(uint64_t key, bool create) {
        size_t h = hash(key);
        while (keys[h] != UINT64_MAX) {
            if (keys[h] == key) return &blocks[vals[h]];
            h = (h + 1) & mask;
        }
        if (!create) return nullptr;
        if ((used + 1) * 2 > mask + 1) {
            rehash((mask + 1) * 2);
            h = hash(key);
            while (keys[h] != UINT64_MAX) h = (h + 1) & mask;
        }
        keys[h] = key;
        vals[h] = (uint32_t)blocks.size();
        blocks.emplace_back();
        memset(&blocks.back(), 0xFF, sizeof(Block));
        ++used;
        return &blocks.back();
    }
    uint32_t *slot(uint32_t node, uint32_t offset, State st, bool create) {
        uint32_t r = g->rank[node];
        uint64_t key = ((uint64_t)(r >> 3) << 32) | (offset >> 3);
        Block *b = find(key, create);
        if (!b) return nullptr;
        Cell &c = b->c[r & 7][offset & 7];
        return st == MATCH ? &c.m : (st == INSERTION ? &c.i : &c.d);
    }
    uint32_t get(uint32_t node, uint32_t offset, State st) {
        uint32_t *s = slot(node, offset, st, false);
        return s ? *s : UNVISITED;
    }
    void set(uint32_t node, uint32_t offset, State st, uint32_t score) { *slot(node, offset, st, true) = score; }
    bool update_if_lower(uint32_t node, uint32_t offset, State st, uint32_t score) {  // gap_affine.rs:520-548
        uint32_t *s = slot(node, offset, st, true);
        if (score < *s) {
            *s = score;
            return true;
        }
        return false;
    }
};

inline uint32_t nonmax(uint32_t v) {  // NonMaxU32::new(v).unwrap()
    if (v == 0xFFFFFFFFu) throw OraclePanic{3};
    return v;
}
inline uint32_t score_add(uint32_t s, uint32_t d) {
    if (s == UNVISITED) throw OraclePanic{3};  // "Can't add to Score::Unvisited!"
    return nonmax(s + d);
}
inline uint32_t score_sub(uint32_t s, uint32_t d) { return nonmax(s - d); }  // wraps (release build), then NonMax check

// ---- AffineAstarData: gap_affine.rs:702-921 + bubbles/reached.rs ------------------------------------
struct AstarData {
    const Graph &g;
    Costs costs;
    size_t seq_len;
    const BubbleIndex &bubbles;
    VisitedStore visited;
    std::vector<std::set<uint32_t>> reached_m;
    std::vector<uint32_t> *pruned_log = nullptr;  // diagnostics (tools/tie_study.cpp): (node, offset, state) of every pruned state

    AstarData(const Graph &graph, Costs c, size_t qlen, const BubbleIndex &b)
        : g(graph), costs(c), seq_len(qlen), bubbles(b), visited(graph), reached_m(graph.n()) {}

    uint32_t get_score(uint32_t n, uint32_t o, State s) { return visited.get(n, o, s); }
    void set_score(uint32_t n, uint32_t o, State s, uint32_t sc) { visited.set(n, o, s, sc); }
    bool update_score_if_lower(uint32_t n, uint32_t o, State s, uint32_t sc) { return visited.update_if_lower(n, o, s, sc); }
    void mark_reached(uint32_t n, uint32_t o, State s) {  // gap_affine.rs:767-773
        if (s == MATCH && bubbles.is_exit(n)) reached_m[n].insert(o);
    }
    void dfa_match(uint32_t n, uint32_t o) { mark_reached(n, o, MATCH); }

    // reached.rs:191-255
    bool can_improve_at_offset(uint32_t exit, uint32_t off, uint32_t score, const uint32_t *left, const uint32_t *right,
                               size_t min_dist_to_end) {
        bool have = false;
        uint32_t implicit = 0;
        if (left && right) {
            uint32_t ls = get_score(exit, *left, MATCH), rs = get_score(exit, *right, MATCH);
            uint32_t gl = off - *left, gr = *right - off;
            uint32_t il = score_add(ls, (uint32_t)costs.gap_cost(MATCH, gl));
            uint32_t ir = score_add(rs, (uint32_t)costs.gap_cost(MATCH, gr));
            implicit = (gr > min_dist_to_end) ? il : std::min(il, ir);
            have = true;
        } else if (right) {
            uint32_t rs = get_score(exit, *right, MATCH);
            uint32_t gr = *right - off;
            uint32_t ir = score_add(rs, (uint32_t)costs.gap_cost(MATCH, gr));
            if (!(gr > min_dist_to_end)) {
                implicit = ir;
                have = true;
            }
        } else if (left) {
            uint32_t ls = get_score(exit, *left, MATCH);
            uint32_t gl = off - *left;
            implicit = score_add(ls, (uint32_t)costs.gap_cost(MATCH, gl));
            have = true;
        }
        return have ? (score < implicit) : true;
    }

    // reached.rs:38-189
    bool can_improve_bubble(const NodeBubbleMap &bubble, uint32_t node, uint32_t offset, State st, uint32_t score) {
        const std::set<uint32_t> &reached = reached_m[bubble.bubble_exit];
        if (reached.empty()) return true;
        if (node == bubble.bubble_exit) return true;
        uint32_t tmin = offset + (uint32_t)bubble.min_dist_to_exit;
        uint32_t tmax = offset + (uint32_t)bubble.max_dist_to_exit;
        size_t a = bubbles.dist_to_end[bubble.bubble_exit].first;
        size_t min_dist_to_end = a ? a - 1 : 0;
        if ((size_t)tmax > seq_len) return true;

        const uint32_t *prev = nullptr;
        {
            auto it = reached.lower_bound(tmin);  // range(..tmin).next_back()
            if (it != reached.begin()) prev = &*std::prev(it);
        }
        bool have_last = false;
        uint32_t last_offset = 0;
        for (auto it = reached.lower_bound(tmin); it != reached.end() && *it <= tmax; ++it) {
            const uint32_t *next = &*it;
            uint32_t offset1 = prev ? std::max(tmin, *prev + 1) : tmin;
            if (st == DELETION) {
                uint32_t c = get_score(bubble.bubble_exit, *next, MATCH);
                if (score_add(c, costs.o) > score) return true;
            }
            if (prev && st == INSERTION) {
                uint32_t c = get_score(bubble.bubble_exit, *prev, MATCH);
                if (score_add(c, costs.o) > score) return true;
            }
            if (can_improve_at_offset(bubble.bubble_exit, offset1, score, prev, next, min_dist_to_end)) return true;
            uint32_t offset2 = std::min(tmax, std::max(tmin, *next - 1u));  // u32 wrap when *next == 0 (release build)
            if (offset2 != offset1) {
                if (can_improve_at_offset(bubble.bubble_exit, offset2, score, prev, next, min_dist_to_end)) return true;
            }
            prev = next;
            last_offset = offset2;
            have_last = true;
        }
        const uint32_t *next = nullptr;
        {
            uint32_t from = (tmax == 0xFFFFFFFFu) ? tmax : tmax + 1;  // saturating_add
            auto it = reached.lower_bound(from);
            if (it != reached.end()) next = &*it;
        }
        if (!have_last && can_improve_at_offset(bubble.bubble_exit, tmin, score, prev, next, min_dist_to_end)) return true;
        if ((!have_last || last_offset < tmax) &&
            can_improve_at_offset(bubble.bubble_exit, tmax, score, prev, next, min_dist_to_end))
            return true;
        if (prev && st == INSERTION) {
            uint32_t c = get_score(bubble.bubble_exit, *prev, MATCH);
            if (score_add(c, costs.o) > score) return true;
        }
        return false;
    }

    bool prune(uint32_t score, uint32_t node, uint32_t offset, State st) {  // gap_affine.rs:780-792
        const auto &bs = bubbles.node_bubble_map[node];
        if (bs.empty()) return false;
        for (const auto &b : bs)
            if (!can_improve_bubble(b, node, offset, st, score)) {  // !all(...)
                if (pruned_log) pruned_log->insert(pruned_log->end(), {node, offset, (uint32_t)st});
                return true;
            }
        return false;
    }

    // gap_affine.rs:550-657; returns false for None.
    bool get_backtrace(const uint8_t *seq, uint32_t node, uint32_t offset, State st, uint32_t &bn, uint32_t &bo, State &bs) {
        uint32_t curr = get_score(node, offset, st);
        if (curr == UNVISITED) return false;
        if (st == MATCH) {
            if (offset > 0) {
                bool is_match_or_end = g.is_symbol_equal(node, seq[offset - 1]) || node == g.end();
                uint32_t pred_offset = (node == g.end()) ? offset : offset - 1;
                for (const uint32_t *p = g.pred_begin(node); p != g.pred_end(node); ++p) {
                    uint32_t ps = get_score(*p, pred_offset, MATCH);
                    if ((is_match_or_end && ps == curr) || (!is_match_or_end && ps == score_sub(curr, costs.x))) {
                        bn = *p, bo = pred_offset, bs = MATCH;
                        return true;
                    }
                }
            }
            if (get_score(node, offset, DELETION) == curr) {
                bn = node, bo = offset, bs = DELETION;
                return true;
            }
            if (get_score(node, offset, INSERTION) == curr) {
                bn = node, bo = offset, bs = INSERTION;
                return true;
            }
        } else if (st == DELETION) {
            for (const uint32_t *p = g.pred_begin(node); p != g.pred_end(node); ++p) {
                uint32_t ps = get_score(*p, offset, MATCH);
                if (ps == score_sub(score_sub(curr, costs.o), costs.e)) {
                    bn = *p, bo = offset, bs = MATCH;
                    return true;
                }
            }
            for (const uint32_t *p = g.pred_begin(node); p != g.pred_end(node); ++p) {
                uint32_t ps = get_score(*p, offset, DELETION);
                if (ps == score_sub(curr, costs.e)) {
                    bn = *p, bo = offset, bs = DELETION;
                    return true;
                }
            }
        } else {
            if (offset > 0) {
                uint32_t ps = get_score(node, offset - 1, MATCH);
                if (ps == score_sub(score_sub(curr, costs.o), costs.e)) {
                    bn = node, bo = offset - 1, bs = MATCH;
                    return true;
                }
                ps = get_score(node, offset - 1, INSERTION);
                if (ps == score_sub(curr, costs.e)) {
                    bn = node, bo = offset - 1, bs = MATCH;  // sic: the reference returns Match here (gap_affine.rs:647-650)
                    return true;
                }
            }
        }
        return false;
    }

    // gap_affine.rs:804-915
    void backtrace(const uint8_t *seq, size_t qlen, uint32_t end_node, uint32_t end_offset, std::vector<uint32_t> &rpos,
                   std::vector<uint32_t> &qpos) {
        rpos.clear();
        qpos.clear();
        if (qlen == 0) return;
        if (qlen == 1 && end_offset == 1 && g.is_symbol_equal(end_node, seq[0])) {
            rpos.push_back(end_node);
            qpos.push_back(0);
            return;
        }
        uint32_t cn, co;
        State cs;
        bool ok = get_backtrace(seq, end_node, end_offset, MATCH, cn, co, cs) ||
                  get_backtrace(seq, end_node, end_offset, INSERTION, cn, co, cs) ||
                  get_backtrace(seq, end_node, end_offset, DELETION, cn, co, cs);
        if (!ok) {
            if (qlen <= 3) {
                for (uint32_t i = 0; i < qlen; ++i) {
                    rpos.push_back(end_node);
                    qpos.push_back(i);
                }
                return;
            }
            throw OraclePanic{2};
        }
        uint32_t bn, bo;
        State bs;
        while (get_backtrace(seq, cn, co, cs, bn, bo, bs)) {
            if (cs == MATCH && (bs == INSERTION || bs == DELETION)) {
                cn = bn, co = bo, cs = bs;
                continue;
            }
            if (cs == MATCH) {
                rpos.push_back(cn);
                qpos.push_back(co - 1);  // usize underflow when co == 0 would be a huge index in the reference
            } else if (cs == INSERTION) {
                rpos.push_back(NONE);
                qpos.push_back(co - 1);
            } else {
                rpos.push_back(cn);
                qpos.push_back(NONE);
            }
            if (bn == g.start()) break;
            cn = bn, co = bo, cs = bs;
        }
        std::reverse(rpos.begin(), rpos.end());
        std::reverse(qpos.begin(), qpos.end());
    }
};

// ---- aligner/dfa.rs:109-263 (templated on the visited type so the reference's DummyVisited tests can be replayed) ----
enum class Ext { None, RefGraphEnd, QueryEnd, Mismatch };
struct ExtendResult {
    Ext kind = Ext::None;
    uint32_t parent_node = 0, parent_offset = 0, child_node = 0, child_offset = 0;
};

template <class V> struct DepthFirstGreedyAlignment {
    const Graph &g;
    const uint8_t *seq;
    size_t qlen;
    uint32_t score;
    size_t num_visited = 0, num_pruned = 0;
    struct Frame {
        uint32_t node, offset;
        const uint32_t *it;
    };
    std::vector<Frame> stack;

    DepthFirstGreedyAlignment(const Graph &graph, const uint8_t *s, size_t n, uint32_t sc, uint32_t node, uint32_t offset)
        : g(graph), seq(s), qlen(n), score(sc) {
        stack.push_back({node, offset, g.succ_begin(node)});
    }

    ExtendResult extend(V &vis) {
        ExtendResult r;
        if (stack.size() == 1 && qlen != 0) {  // dfa.rs:145-168 (offset-0 special case)
            Frame init = stack[0];
            if (init.offset == 0 && g.is_symbol_equal(init.node, seq[0])) {
                if (vis.update_score_if_lower(init.node, 1, MATCH, score)) {
                    stack[0] = {init.node, 1, g.succ_begin(init.node)};
                    vis.dfa_match(init.node, 1);
                    ++num_visited;
                    if (1 == qlen) {
                        r.kind = Ext::RefGraphEnd;
                        r.parent_node = init.node, r.parent_offset = 0, r.child_node = init.node, r.child_offset = 1;
                        return r;
                    }
                }
            }
        }
        while (!stack.empty()) {
            Frame &parent = stack.back();
            // next_valid_successor (dfa.rs:210-250)
            int kind = 0;  // 0 exhausted, 1 refend, 2 queryend, 3 match, 4 mismatch
            uint32_t cnode = 0, coff = 0;
            const uint32_t *e = g.succ_end(parent.node);
            while (parent.it != e) {
                uint32_t child = *parent.it++;
                if (child == g.end()) {
                    vis.update_score_if_lower(child, parent.offset, MATCH, score);
                    kind = 1, cnode = child, coff = parent.offset;
                    break;
                }
                if (parent.offset >= qlen) {
                    kind = 2, cnode = child;
                    break;
                }
                uint32_t child_offset = parent.offset + 1;
                if (g.is_symbol_equal(child, seq[child_offset - 1])) {
                    if (vis.update_score_if_lower(child, child_offset, MATCH, score)) {
                        kind = 3, cnode = child, coff = child_offset;
                        break;
                    }
                } else {
                    kind = 4, cnode = child, coff = child_offset;
                    break;
                }
            }
            r.parent_node = parent.node, r.parent_offset = parent.offset, r.child_node = cnode, r.child_offset = coff;
            if (kind == 1) {
                r.kind = Ext::RefGraphEnd;
                return r;
            } else if (kind == 2) {
                r.kind = Ext::QueryEnd;
                return r;
            } else if (kind == 3) {
                if (vis.prune(score, cnode, coff, MATCH)) {
                    ++num_pruned;
                    continue;
                }
                vis.dfa_match(cnode, coff);
                ++num_visited;
                stack.push_back({cnode, coff, g.succ_begin(cnode)});
            } else if (kind == 4) {
                r.kind = Ext::Mismatch;
                return r;
            } else {
                stack.pop_back();
            }
        }
        r.kind = Ext::None;
        return r;
    }
};

struct AstarResult {
    uint32_t score = 0;
    std::vector<uint32_t> rpos, qpos;
    uint64_t num_queued = 0, num_visited = 0, num_pruned = 0;
    int status = 0;  // 0 ok, OraclePanic code otherwise
    uint32_t end_node = 0, end_offset = 0;
};

struct AlignOptions {
    Costs costs{4, 6, 2};
    int mode = 0;       // 0 Global, 1 EndsFree all-Unbounded (bin/poasta.rs:287-301)
    int heuristic = 1;  // 0 dijkstra, 1 mingap, 2 path
    bool pruning = true;
    bool mock = false;
    // optional: dump of the visited table for diagnostics (tests only)
    std::vector<uint32_t> *dump = nullptr;  // (node, offset, state, score)*
    std::vector<uint32_t> *pruned_log = nullptr;  // (node, offset, state)* of pruned states
    std::vector<uint32_t> *pop_log = nullptr;     // (node, offset, state, score)* of expanded (popped, not stale, not pruned) states
};

// gap_affine.rs:185-248 with the four bounds Unbounded (what the CLI passes)
inline bool is_end(const Graph &g, int mode, size_t qlen, uint32_t node, uint32_t offset, State st) {
    if (mode == 0) return st == MATCH && node == g.end() && offset == qlen;
    return st == MATCH && (offset > 0 || qlen == 0);
}

// aligner/astar.rs:108-226 + aligner/mod.rs:114-145
inline AstarResult astar_align(const poasta_graph_view &view, const uint8_t *seq, size_t qlen, const AlignOptions &opt) {
    AstarResult res;
    Graph g(view, opt.mock);
    if (g.node_count() == 0) {  // mod.rs:124-142
        res.score = qlen == 0 ? 0 : (uint32_t)qlen * 4;
        return res;
    }
    try {
        BubbleIndex bubbles(g);  // rebuilt per alignment (config.rs:128)
        std::unique_ptr<PathIndex> pidx;
        Heuristic heur;
        heur.kind = opt.heuristic;
        heur.costs = opt.costs;
        heur.bubbles = &bubbles;
        heur.seq_len = qlen;
        if (opt.heuristic == 2) {
            pidx.reset(new PathIndex(g, 10));  // config.rs:299
            heur.paths = pidx.get();
        }
        AstarData vis(g, opt.costs, qlen, bubbles);
        vis.pruned_log = opt.pruned_log;
        AffineQueue queue;
        const Costs &c = opt.costs;

        // initial states: gap_affine.rs:136-183
        std::vector<uint32_t> init;
        if (opt.mode == 0) {
            init.push_back(g.start());
        } else {
            for (uint32_t n = 0; n < g.n(); ++n)
                if (n != g.start() && n != g.end()) init.push_back(n);
            std::reverse(init.begin(), init.end());
            if (init.empty()) init.push_back(g.start());
        }
        for (uint32_t n : init) {
            size_t h = heur.h(n, 0, MATCH);
            queue.queue(n, 0, MATCH, 0, h);
            vis.set_score(n, 0, MATCH, 0);
            ++res.num_queued;
        }
        auto push = [&](uint32_t score, uint8_t delta, uint32_t node, uint32_t offset, State st) {
            size_t h = heur.h(node, offset, st);
            ++res.num_queued;
            queue.queue(node, offset, st, score_add(score, delta), h);
        };
        uint32_t end_score = 0, end_node = 0, end_offset = 0;
        bool done = false;
        while (!done) {
            QueueItem it{};
            if (!queue.pop(it)) throw OraclePanic{1};
            if (it.score > vis.get_score(it.node, it.offset, it.state)) continue;
            if (is_end(g, opt.mode, qlen, it.node, it.offset, it.state)) {
                ++res.num_visited;
                end_score = it.score, end_node = it.node, end_offset = it.offset;
                break;
            }
            if (opt.pruning && vis.prune(it.score, it.node, it.offset, it.state)) {
                ++res.num_pruned;
                continue;
            }
            vis.mark_reached(it.node, it.offset, it.state);
            ++res.num_visited;
            const uint32_t score = it.score;
            if (opt.pop_log) opt.pop_log->insert(opt.pop_log->end(), {it.node, it.offset, (uint32_t)it.state, score});
            if (it.state == MATCH) {
                DepthFirstGreedyAlignment<AstarData> dfa(g, seq, qlen, score, it.node, it.offset);
                for (;;) {
                    ExtendResult ev = dfa.extend(vis);
                    if (ev.kind == Ext::None) break;
                    if (ev.kind == Ext::RefGraphEnd) {
                        if (is_end(g, opt.mode, qlen, ev.child_node, ev.child_offset, MATCH)) {
                            end_score = score, end_node = ev.child_node, end_offset = ev.child_offset;
                            done = true;
                            break;
                        }
                        // expand_ref_graph_end: gap_affine.rs:346-367
                        uint32_t ns = score_add(score_add(score, c.o), c.e);
                        if (vis.update_score_if_lower(ev.parent_node, ev.parent_offset + 1, INSERTION, ns))
                            push(score, c.open_extend(), ev.parent_node, ev.parent_offset + 1, INSERTION);
                    } else if (ev.kind == Ext::QueryEnd) {
                        // expand_query_end: gap_affine.rs:369-391
                        uint32_t ns = score_add(score_add(score, c.o), c.e);
                        if (vis.update_score_if_lower(ev.child_node, ev.parent_offset, DELETION, ns))
                            push(score, c.open_extend(), ev.child_node, ev.parent_offset, DELETION);
                    } else {
                        // expand_mismatch: gap_affine.rs:393-430
                        if (vis.update_score_if_lower(ev.child_node, ev.child_offset, MATCH, score_add(score, c.x)))
                            push(score, c.x, ev.child_node, ev.child_offset, MATCH);
                        uint32_t ns = score_add(score_add(score, c.o), c.e);
                        if (vis.update_score_if_lower(ev.parent_node, ev.parent_offset + 1, INSERTION, ns))
                            push(score, c.open_extend(), ev.parent_node, ev.parent_offset + 1, INSERTION);
                        if (vis.update_score_if_lower(ev.child_node, ev.parent_offset, DELETION, ns))
                            push(score, c.open_extend(), ev.child_node, ev.parent_offset, DELETION);
                    }
                }
                if (done) break;  // astar.rs:172 breaks out of the main loop before the counter update
                res.num_visited += dfa.num_visited;
            } else if (it.state == INSERTION) {  // expand_all: gap_affine.rs:307-322
                if (vis.update_score_if_lower(it.node, it.offset, MATCH, score)) push(score, 0, it.node, it.offset, MATCH);
                if (it.offset < qlen && vis.update_score_if_lower(it.node, it.offset + 1, INSERTION, score_add(score, c.e)))
                    push(score, c.e, it.node, it.offset + 1, INSERTION);
            } else {  // Deletion: gap_affine.rs:323-341
                if (vis.update_score_if_lower(it.node, it.offset, MATCH, score)) push(score, 0, it.node, it.offset, MATCH);
                for (const uint32_t *s = g.succ_begin(it.node); s != g.succ_end(it.node); ++s)
                    if (vis.update_score_if_lower(*s, it.offset, DELETION, score_add(score, c.e)))
                        push(score, c.e, *s, it.offset, DELETION);
            }
        }
        res.score = end_score;
        res.end_node = end_node;
        res.end_offset = end_offset;
        vis.backtrace(seq, qlen, end_node, end_offset, res.rpos, res.qpos);
        if (opt.dump) {
            std::vector<uint32_t> inv(g.n());
            for (uint32_t n = 0; n < g.n(); ++n) inv[g.rank[n]] = n;
            for (size_t k = 0; k < vis.visited.keys.size(); ++k) {
                if (vis.visited.keys[k] == UINT64_MAX) continue;
                uint32_t rb = (uint32_t)(vis.visited.keys[k] >> 32), ob = (uint32_t)vis.visited.keys[k];
                const auto &b = vis.visited.blocks[vis.visited.vals[k]];
                for (int r = 0; r < 8; ++r)
                    for (int cc = 0; cc < 8; ++cc) {
                        uint32_t rk = rb * 8 + r;
                        if (rk >= g.n()) continue;
                        const uint32_t sc[3] = {b.c[r][cc].m, b.c[r][cc].d, b.c[r][cc].i};
                        for (uint32_t s = 0; s < 3; ++s)
                            if (sc[s] != UNVISITED) {
                                opt.dump->push_back(inv[rk]);
                                opt.dump->push_back(ob * 8 + cc);
                                opt.dump->push_back(s);
                                opt.dump->push_back(sc[s]);
                            }
                    }
            }
        }
    } catch (const OraclePanic &p) {
        res.status = p.code;
    }
    return res;
}

}  // namespace oracle
