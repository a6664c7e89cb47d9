// Host-side preprocessing for the order-exact search kernel (replay_core.hpp): the per-graph
// structures POASTA derives before every alignment, flattened for the device.
//   superbubbles            /root/reference/src/bubbles/finder.rs:30-177 (Gaertner et al. 2018)
//   bubble membership, dist /root/reference/src/bubbles/index.rs:51-156
//   greedy path sample      /root/reference/src/aligner/path_index.rs:48-267 (for -H path)
// All arrays are indexed by NODE ID (the search works in node-id space like the reference).
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <set>
#include <vector>

#include "../../include/poasta_b200.h"
#include "replay_core.hpp"

namespace poasta {

struct SearchIndex {
    std::vector<uint32_t> dist_min, dist_max;  // shortest / longest path length to `end`
    std::vector<uint8_t> is_exit;
    std::vector<uint32_t> bub_off;  // V + 1
    std::vector<uint32_t> bub;      // quadruples (exit, min_dist_to_exit, max_dist_to_exit, max(dist_min(exit), 1) - 1)
    std::vector<uint32_t> path_off;  // V + 1 (only when the path heuristic is requested)
    std::vector<uint32_t> path_rem;  // remaining path length for the first <= 5 sampled paths through a node

    // 32-byte node records of the search kernel (replay_core.hpp); offsets are relative to the graph's own arrays
    void node_records(const poasta_graph_view &v, std::vector<replay::NodeRec> &out) const {
        out.reserve(out.size() + v.n_nodes);
        for (uint32_t n = 0; n < v.n_nodes; ++n) {
            replay::NodeRec r;
            r.a.succ_off = v.succ_off[n];
            r.a.bub_off = bub_off[n];
            r.a.counts = (v.succ_off[n + 1] - v.succ_off[n]) | ((bub_off[n + 1] - bub_off[n]) << 16);
            r.a.symflags = (uint32_t)v.symbol[n] | ((uint32_t)is_exit[n] << 8);
            r.b.dist_min = dist_min[n];
            r.b.dist_max = dist_max[n];
            r.b.pred_off = v.pred_off[n];
            r.b.n_pred = v.pred_off[n + 1] - v.pred_off[n];
            out.push_back(r);
        }
    }

    struct G {
        const poasta_graph_view &v;
        const uint32_t *sb(uint32_t x) const { return v.succ + v.succ_off[x]; }
        const uint32_t *se(uint32_t x) const { return v.succ + v.succ_off[x + 1]; }
        const uint32_t *pb(uint32_t x) const { return v.pred + v.pred_off[x]; }
        const uint32_t *pe(uint32_t x) const { return v.pred + v.pred_off[x + 1]; }
    };

    // Work buffers kept per host thread: the index of a family's graph is rebuilt every round (config.rs:128), ~10^5 times
    // per bench step, so build() allocates nothing once the buffers have grown (it was 350 us per 1.4k-node graph with
    // per-node vectors and copied bubble stacks; profiles/README.md).
    struct Scratch {
        std::vector<uint32_t> order, pos, out_parent, out_child, entrance_exit, stack, queue;
        std::vector<int64_t> opm;
        std::vector<uint8_t> seen, from_start;
        std::vector<std::pair<uint32_t, const uint32_t *>> dfs;
        struct Open {  // one open bubble on the persistent stack carried by the backward BFS: (distance of the exit from end, exit)
            uint32_t exit_dist, exit, parent;
        };
        std::vector<Open> open;
        std::vector<uint32_t> top_of, dist_of;  // per queued node: its stack top (index into open, NONE = empty) and BFS distance
        struct Ent {
            uint32_t node, exit, min_d;
        };
        std::vector<Ent> ents;  // bubble memberships in generation order
        std::vector<uint32_t> chain;
    };

    void build(const poasta_graph_view &view, bool with_paths) {
        static thread_local Scratch w;
        const G g{view};
        const uint32_t V = view.n_nodes;
        const uint32_t NONE = 0xFFFFFFFFu;
        // reverse post-order of the DFS from start, children in adjacency order, marked when pushed (graphs/tools.rs:5-37)
        std::vector<uint32_t> &order = w.order;
        order.clear();
        order.reserve(V);
        {
            w.seen.assign(V, 0);
            auto &st = w.dfs;
            st.clear();
            st.emplace_back(view.start_node, g.sb(view.start_node));
            while (!st.empty()) {
                auto &top = st.back();
                uint32_t child = NONE;
                for (const uint32_t *e = g.se(top.first); top.second != e;) {
                    const uint32_t c = *top.second++;
                    if (!w.seen[c]) {
                        child = c;
                        break;
                    }
                }
                if (child != NONE) {
                    w.seen[child] = 1;
                    st.emplace_back(child, g.sb(child));
                } else {
                    order.push_back(top.first);
                    st.pop_back();
                }
            }
            std::reverse(order.begin(), order.end());
            w.from_start = w.seen;
            w.from_start[view.start_node] = 1;
        }
        // positions in that order; nodes the DFS did not reach keep position 0 like the reference's zero-initialised table.
        // out_parent: smallest predecessor position (NONE = no predecessor, the reference's -1);
        // out_child: largest successor position (NONE = no successor, the reference's i64::MAX)
        w.pos.assign(V, 0);
        w.out_parent.assign(V, NONE);
        w.out_child.assign(V, NONE);
        for (size_t i = 0; i < order.size(); ++i) w.pos[order[i]] = (uint32_t)i;
        for (uint32_t n = 0; n < V; ++n) {
            if (g.pb(n) != g.pe(n)) {
                uint32_t mn = NONE;
                for (const uint32_t *p = g.pb(n); p != g.pe(n); ++p) mn = std::min(mn, w.pos[*p]);
                w.out_parent[n] = mn;
            }
            if (g.sb(n) != g.se(n)) {
                uint32_t mx = 0;
                for (const uint32_t *s = g.sb(n); s != g.se(n); ++s) mx = std::max(mx, w.pos[*s]);
                w.out_child[n] = mx;
            }
        }
        // superbubble scan, high rank to low (finder.rs:109-177); later pairs overwrite earlier ones (index.rs:61-64).
        // opm holds "outgoing parent minimum" positions as signed values of the reference: -1 (no predecessor) is kept as
        // int64 so the min() / equality tests below behave like the reference's i64 arithmetic.
        w.entrance_exit.assign(V, NONE);
        is_exit.assign(V, 0);
        {
            std::vector<int64_t> &opm = w.opm;
            opm.assign(V, 0);
            auto parent_of = [&](uint32_t n) -> int64_t { return w.out_parent[n] == NONE ? -1 : (int64_t)w.out_parent[n]; };
            auto child_of = [&](uint32_t n) -> int64_t { return w.out_child[n] == NONE ? INT64_MAX : (int64_t)w.out_child[n]; };
            auto &stack = w.stack;
            stack.clear();
            uint32_t cand = NONE;
            auto pop = [&]() {
                if (stack.empty()) return NONE;
                const uint32_t c = stack.back();
                stack.pop_back();
                return c;
            };
            for (int64_t curr = (int64_t)order.size() - 1; curr >= 0; --curr) {
                const uint32_t n = order[curr];
                const int64_t far = child_of(n);
                if (far == curr + 1) {
                    if (cand != NONE) stack.push_back(cand);
                    cand = order[curr + 1];
                } else {
                    while (cand != NONE) {
                        const uint32_t c0 = cand;
                        if (far <= (int64_t)w.pos[c0]) break;
                        cand = pop();
                        if (cand != NONE) opm[cand] = std::min(opm[c0], opm[cand]);
                    }
                }
                if (cand != NONE && (uint64_t)opm[cand] == (uint64_t)curr) {
                    const uint32_t c0 = cand;
                    w.entrance_exit[n] = c0;
                    is_exit[c0] = 1;
                    cand = pop();
                    if (cand != NONE) opm[cand] = std::min(opm[c0], opm[cand]);
                }
                opm[n] = parent_of(n);
                if (cand != NONE) opm[cand] = std::min(opm[n], opm[cand]);
            }
        }
        // backward BFS from end carrying the stack of open bubbles (index.rs:73-131). The reference clones a Vec per queued
        // node; here the stacks share structure: `open` is a pool of (exit distance, exit, parent) and a queued node only
        // remembers its top. A node's memberships are generated in the reference's order: the bubble popped for it when it
        // was discovered first, then its own stack from bottom to top when it is dequeued.
        dist_min.assign(V, 0);
        dist_max.assign(V, 0);
        w.ents.clear();
        {
            auto &open = w.open;
            open.clear();
            w.top_of.assign(V, NONE);
            w.dist_of.assign(V, 0);
            auto &queue = w.queue;
            queue.clear();
            if (is_exit[view.end_node]) {
                open.push_back({0u, view.end_node, NONE});
                w.top_of[view.end_node] = 0;
            }
            queue.push_back(view.end_node);
            w.seen.assign(V, 0);
            w.seen[view.end_node] = 1;
            for (size_t head = 0; head < queue.size(); ++head) {
                const uint32_t node = queue[head];
                const uint32_t dist = w.dist_of[node], top = w.top_of[node];
                w.chain.clear();
                for (uint32_t o = top; o != NONE; o = open[o].parent) w.chain.push_back(o);
                for (size_t i = w.chain.size(); i-- > 0;) w.ents.push_back({node, open[w.chain[i]].exit, dist - open[w.chain[i]].exit_dist});
                dist_min[node] = dist;
                for (const uint32_t *pp = g.pb(node); pp != g.pe(node); ++pp) {
                    const uint32_t p = *pp;
                    if (w.seen[p]) continue;
                    const uint32_t nd = dist + 1;
                    uint32_t ns = top;
                    if (w.entrance_exit[p] != NONE && ns != NONE) {
                        w.ents.push_back({p, open[ns].exit, nd - open[ns].exit_dist});
                        ns = open[ns].parent;
                    }
                    if (is_exit[p]) {
                        open.push_back({nd, p, ns});
                        ns = (uint32_t)open.size() - 1;
                    }
                    w.seen[p] = 1;
                    w.top_of[p] = ns;
                    w.dist_of[p] = nd;
                    queue.push_back(p);
                }
            }
        }
        for (size_t i = order.size(); i-- > 0;) {  // longest distances in post-order (index.rs:133-148)
            const uint32_t n = order[i];
            uint32_t mx = 0;
            for (const uint32_t *s = g.sb(n); s != g.se(n); ++s) mx = std::max(mx, dist_max[*s] + 1);
            dist_max[n] = mx;
        }
        // memberships grouped by node, generation order kept inside a node (stable counting sort)
        bub_off.assign(V + 1, 0);
        for (const auto &e : w.ents) ++bub_off[e.node + 1];
        for (uint32_t n = 0; n < V; ++n) bub_off[n + 1] += bub_off[n];
        bub.resize(4 * w.ents.size());
        w.chain.assign(bub_off.begin(), bub_off.end() - 1);  // write cursors
        for (const auto &e : w.ents) {
            uint32_t *q = bub.data() + 4 * (size_t)w.chain[e.node]++;
            q[0] = e.exit;
            q[1] = e.min_d;
            q[2] = w.from_start[e.node] ? dist_max[e.node] - dist_max[e.exit] : 0;  // the reference leaves 0 for nodes its DFS never saw
            q[3] = dist_min[e.exit] ? dist_min[e.exit] - 1 : 0;  // min_dist_to_end used by the prune test (reached.rs:57-58)
        }
        path_off.clear();
        path_rem.clear();
        if (with_paths) build_paths(view);
    }

    // path_index.rs:64-237 with max_paths_per_node = 10; the heuristic reads the first 5 entries (config.rs:296-308)
    void build_paths(const poasta_graph_view &view) {
        const G g{view};
        const uint32_t V = view.n_nodes;
        std::set<std::pair<uint32_t, uint32_t>> used;
        std::vector<uint8_t> has_used_out(V, 0);
        std::vector<std::vector<uint32_t>> paths;
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> node_paths(V);
        auto out_deg = [&](uint32_t n) { return view.succ_off[n + 1] - view.succ_off[n]; };
        auto in_deg = [&](uint32_t n) { return view.pred_off[n + 1] - view.pred_off[n]; };
        auto extract = [&](uint32_t start) {
            std::vector<uint32_t> nodes{start};
            uint32_t cur = start;
            while (cur != view.end_node && g.sb(cur) != g.se(cur)) {
                uint32_t next = 0xFFFFFFFFu, best = 0;
                for (const uint32_t *s = g.sb(cur); s != g.se(cur); ++s) {
                    if (used.count({cur, *s})) continue;
                    const uint32_t od = out_deg(*s);
                    if (next == 0xFFFFFFFFu || od >= best) next = *s, best = od;  // max_by_key keeps the last maximum
                }
                if (next == 0xFFFFFFFFu) next = *g.sb(cur);
                used.insert({cur, next});
                has_used_out[cur] = 1;
                nodes.push_back(next);
                cur = next;
            }
            return nodes;
        };
        auto add = [&](std::vector<uint32_t> nodes) {
            const uint32_t id = (uint32_t)paths.size();
            for (size_t p = 0; p < nodes.size(); ++p)
                if (node_paths[nodes[p]].size() < 10) node_paths[nodes[p]].emplace_back(id, (uint32_t)p);
            paths.push_back(std::move(nodes));
        };
        std::vector<uint32_t> seeds{view.start_node};
        for (uint32_t n = 0; n < V; ++n)
            if (in_deg(n) == 0 || (out_deg(n) > 2 && in_deg(n) == 1)) seeds.push_back(n);
        for (uint32_t s : seeds) {
            if (has_used_out[s]) continue;
            auto p = extract(s);
            if (p.size() > 1) add(std::move(p));
        }
        if (paths.size() < 10) {
            std::vector<std::pair<uint32_t, uint32_t>> cands;
            for (uint32_t n = 0; n < V; ++n) {
                uint32_t cnt = 0;
                for (const uint32_t *s = g.sb(n); s != g.se(n); ++s) cnt += !used.count({n, *s});
                if (cnt) cands.emplace_back(n, cnt);
            }
            std::stable_sort(cands.begin(), cands.end(), [](auto &a, auto &b) { return a.second > b.second; });
            for (size_t i = 0; i < cands.size() && i < 20; ++i) {
                auto p = extract(cands[i].first);
                if (p.size() > 3) add(std::move(p));
            }
        }
        path_off.assign(V + 1, 0);
        for (uint32_t n = 0; n < V; ++n) {
            const size_t k = std::min<size_t>(node_paths[n].size(), 5);
            for (size_t i = 0; i < k; ++i)
                path_rem.push_back((uint32_t)paths[node_paths[n][i].first].size() - 1 - node_paths[n][i].second);
            path_off[n + 1] = (uint32_t)path_rem.size();
        }
    }
};

}  // namespace poasta
