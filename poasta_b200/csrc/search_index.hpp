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

    void build(const poasta_graph_view &view, bool with_paths) {
        const G g{view};
        const uint32_t V = view.n_nodes;
        // reverse post-order of the DFS from start, children in adjacency order, marked when pushed (graphs/tools.rs:5-37)
        std::vector<uint32_t> order;
        order.reserve(V);
        {
            std::vector<uint8_t> seen(V, 0);
            std::vector<std::pair<uint32_t, const uint32_t *>> st;
            st.emplace_back(view.start_node, g.sb(view.start_node));
            while (!st.empty()) {
                auto &top = st.back();
                uint32_t child = 0xFFFFFFFFu;
                for (const uint32_t *e = g.se(top.first); top.second != e;) {
                    const uint32_t c = *top.second++;
                    if (!seen[c]) {
                        child = c;
                        break;
                    }
                }
                if (child != 0xFFFFFFFFu) {
                    seen[child] = 1;
                    st.emplace_back(child, g.sb(child));
                } else {
                    order.push_back(top.first);
                    st.pop_back();
                }
            }
            std::reverse(order.begin(), order.end());
        }
        std::vector<int64_t> pos(V, 0), out_parent(V, -1), out_child(V, INT64_MAX);
        for (size_t i = 0; i < order.size(); ++i) pos[order[i]] = (int64_t)i;
        for (uint32_t n = 0; n < V; ++n) {
            if (g.pb(n) != g.pe(n)) {
                int64_t mn = INT64_MAX;
                for (const uint32_t *p = g.pb(n); p != g.pe(n); ++p) mn = std::min(mn, pos[*p]);
                out_parent[n] = mn;
            }
            if (g.sb(n) != g.se(n)) {
                int64_t mx = INT64_MIN;
                for (const uint32_t *s = g.sb(n); s != g.se(n); ++s) mx = std::max(mx, pos[*s]);
                out_child[n] = mx;
            }
        }
        // superbubble scan, high rank to low (finder.rs:109-177); later pairs overwrite earlier ones (index.rs:61-64)
        const uint32_t NONE = 0xFFFFFFFFu;
        std::vector<uint32_t> entrance_exit(V, NONE);
        is_exit.assign(V, 0);
        {
            std::vector<int64_t> opm(V, 0);
            std::vector<uint32_t> stack;
            uint32_t cand = NONE;
            auto pop = [&]() {
                if (stack.empty()) return NONE;
                const uint32_t c = stack.back();
                stack.pop_back();
                return c;
            };
            for (int64_t curr = (int64_t)order.size() - 1; curr >= 0; --curr) {
                const uint32_t n = order[curr];
                const int64_t far = out_child[n];
                if (far == curr + 1) {
                    if (cand != NONE) stack.push_back(cand);
                    cand = order[curr + 1];
                } else {
                    while (cand != NONE) {
                        const uint32_t c0 = cand;
                        if (far <= pos[c0]) break;
                        cand = pop();
                        if (cand != NONE) opm[cand] = std::min(opm[c0], opm[cand]);
                    }
                }
                if (cand != NONE && (uint64_t)opm[cand] == (uint64_t)curr) {
                    const uint32_t c0 = cand;
                    entrance_exit[n] = c0;
                    is_exit[c0] = 1;
                    cand = pop();
                    if (cand != NONE) opm[cand] = std::min(opm[c0], opm[cand]);
                }
                opm[n] = out_parent[n];
                if (cand != NONE) opm[cand] = std::min(opm[n], opm[cand]);
            }
        }
        // backward BFS from end carrying the stack of open bubbles (index.rs:73-131)
        struct Entry {
            uint32_t exit;
            uint32_t min_d, max_d;
        };
        std::vector<std::vector<Entry>> maps(V);
        dist_min.assign(V, 0);
        dist_max.assign(V, 0);
        {
            typedef std::vector<std::pair<uint32_t, uint32_t>> BStack;  // (distance of the exit from end, exit)
            struct Item {
                uint32_t node, dist;
                BStack stack;
            };
            std::deque<Item> queue;
            BStack init;
            if (is_exit[view.end_node]) init.emplace_back(0u, view.end_node);
            queue.push_back({view.end_node, 0, init});
            std::vector<uint8_t> seen(V, 0);
            seen[view.end_node] = 1;
            while (!queue.empty()) {
                Item it = std::move(queue.front());
                queue.pop_front();
                for (auto &b : it.stack) maps[it.node].push_back({b.second, it.dist - b.first, 0});
                dist_min[it.node] = it.dist;
                for (const uint32_t *pp = g.pb(it.node); pp != g.pe(it.node); ++pp) {
                    const uint32_t p = *pp;
                    if (seen[p]) continue;
                    const uint32_t nd = it.dist + 1;
                    BStack ns = it.stack;
                    if (entrance_exit[p] != NONE && !ns.empty()) {
                        const auto top = ns.back();
                        ns.pop_back();
                        maps[p].push_back({top.second, nd - top.first, 0});
                    }
                    if (is_exit[p]) ns.emplace_back(nd, p);
                    seen[p] = 1;
                    queue.push_back({p, nd, std::move(ns)});
                }
            }
        }
        for (size_t i = order.size(); i-- > 0;) {  // longest distances in post-order (index.rs:133-148)
            const uint32_t n = order[i];
            uint32_t mx = 0;
            for (const uint32_t *s = g.sb(n); s != g.se(n); ++s) mx = std::max(mx, dist_max[*s] + 1);
            dist_max[n] = mx;
            for (auto &b : maps[n]) b.max_d = mx - dist_max[b.exit];
        }
        bub_off.assign(V + 1, 0);
        bub.clear();
        for (uint32_t n = 0; n < V; ++n) {
            for (auto &b : maps[n]) {
                bub.push_back(b.exit);
                bub.push_back(b.min_d);
                bub.push_back(b.max_d);
                bub.push_back(dist_min[b.exit] ? dist_min[b.exit] - 1 : 0);  // min_dist_to_end used by the prune test (reached.rs:57-58)
            }
            bub_off[n + 1] = (uint32_t)(bub.size() / 4);
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
