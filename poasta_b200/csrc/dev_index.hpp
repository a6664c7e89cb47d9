// Per-round preprocessing of a device-resident graph (dev_graph.hpp) for the order-exact search (replay_core.hpp):
// what the reference derives before EVERY alignment (aligner/config.rs:113-134), rebuilt here by the family's worker
// right before its search, from the graph the previous round's insertion left in HBM.
//   adjacency (CSR, petgraph order)  /root/reference/src/graphs/poa.rs:433-449
//   reverse post-order               /root/reference/src/graphs/tools.rs:5-37
//   superbubbles                     /root/reference/src/bubbles/finder.rs:30-177
//   bubble membership, distances     /root/reference/src/bubbles/index.rs:51-156
//   sampled paths (-H path)          /root/reference/src/aligner/path_index.rs:48-267
// Same algorithms as search_index.hpp (the host build used by poasta_align_batch, checked entry by entry against the oracle's
// BubbleIndex in tests/test_device_layouts.py), restated over raw arrays in the worker's scratch so the one source runs on
// the device and on the host (tests/test_dev_graph.py compares both builds array by array).
// Cost: a handful of O(V + E) passes by one lane, ~0.1 % of a round whose search visits 10^5 states.
#pragma once
#include "dev_graph.hpp"
#include "replay_core.hpp"

namespace poasta {
namespace dg {

// Worker scratch for the derived arrays of ONE graph; capacities in elements.
struct IndexCaps {
    uint32_t V, E, ents;  // nodes, edges, bubble memberships
};

struct IndexLayout {
    uint64_t nodes, succ, pred, bub, path_off, path_rem, succ_off, pred_off, order, pos, out_parent, out_child, entrance_exit, stack, queue, top_of,
        dist_of, dist_min, dist_max, bub_off, cursor, opm, open, ents, dfs, seen, from_start, is_exit, used, path_cnt, path_tmp, total;
    DG_HD explicit IndexLayout(const IndexCaps &c) {
        uint64_t o = 0;
        auto take = [&](uint64_t bytes) {
            const uint64_t at = o;
            o += (bytes + 31) & ~(uint64_t)31;
            return at;
        };
        const uint64_t V = c.V, E = c.E;
        nodes = take(V * sizeof(replay::NodeRec));
        succ = take(E * 4), pred = take(E * 4);
        bub = take((uint64_t)c.ents * sizeof(replay::BubQ));
        path_off = take((V + 1) * 4), path_rem = take(V * 5 * 4);
        succ_off = take((V + 1) * 4), pred_off = take((V + 1) * 4);
        order = take(V * 4), pos = take(V * 4), out_parent = take(V * 4), out_child = take(V * 4), entrance_exit = take(V * 4);
        stack = take(V * 4), queue = take(V * 4), top_of = take(V * 4), dist_of = take(V * 4);
        dist_min = take(V * 4), dist_max = take(V * 4), bub_off = take((V + 1) * 4), cursor = take(V * 4), opm = take(V * 4);
        open = take(V * 12), ents = take((uint64_t)c.ents * 12), dfs = take(V * 8);
        seen = take(V), from_start = take(V), is_exit = take(V);
        used = take(E), path_cnt = take(V), path_tmp = take((V + 1) * 4);
        total = o;
    }
};

struct IndexBuf {
    IndexCaps cap;
    replay::NodeRec *nodes;
    uint32_t *succ, *pred;
    replay::BubQ *bub;
    uint32_t *path_off, *path_rem;
    uint32_t *succ_off, *pred_off, *order, *pos, *out_parent, *out_child, *entrance_exit, *stack, *queue, *top_of, *dist_of, *dist_min, *dist_max,
        *bub_off, *cursor;
    int32_t *opm;
    uint32_t *open;  // triples (exit_dist, exit, parent)
    uint32_t *ents;  // triples (node, exit, min_d)
    uint32_t *dfs;   // pairs (node, successor cursor)
    uint8_t *seen, *from_start, *is_exit, *used, *path_cnt;
    uint32_t *path_tmp;
    uint32_t n_bub = 0, n_path = 0;

    DG_HD IndexBuf(uint8_t *base, const IndexCaps &c) : cap(c) {
        const IndexLayout L(c);
        nodes = (replay::NodeRec *)(base + L.nodes);
        succ = (uint32_t *)(base + L.succ), pred = (uint32_t *)(base + L.pred);
        bub = (replay::BubQ *)(base + L.bub);
        path_off = (uint32_t *)(base + L.path_off), path_rem = (uint32_t *)(base + L.path_rem);
        succ_off = (uint32_t *)(base + L.succ_off), pred_off = (uint32_t *)(base + L.pred_off);
        order = (uint32_t *)(base + L.order), pos = (uint32_t *)(base + L.pos);
        out_parent = (uint32_t *)(base + L.out_parent), out_child = (uint32_t *)(base + L.out_child);
        entrance_exit = (uint32_t *)(base + L.entrance_exit), stack = (uint32_t *)(base + L.stack), queue = (uint32_t *)(base + L.queue);
        top_of = (uint32_t *)(base + L.top_of), dist_of = (uint32_t *)(base + L.dist_of);
        dist_min = (uint32_t *)(base + L.dist_min), dist_max = (uint32_t *)(base + L.dist_max);
        bub_off = (uint32_t *)(base + L.bub_off), cursor = (uint32_t *)(base + L.cursor);
        opm = (int32_t *)(base + L.opm);
        open = (uint32_t *)(base + L.open), ents = (uint32_t *)(base + L.ents), dfs = (uint32_t *)(base + L.dfs);
        seen = base + L.seen, from_start = base + L.from_start, is_exit = base + L.is_exit, used = base + L.used, path_cnt = base + L.path_cnt;
        path_tmp = (uint32_t *)(base + L.path_tmp);
    }
};

// Returns DG_OK, or DG_CAPACITY when the graph does not fit the worker's scratch (the family is then resumed by a pass
// with larger workers). `lane` / `n_lanes`: the embarrassingly parallel loops are strided over the worker's lanes; the
// sequential passes run on lane 0 between two `sync()` calls; `sync.bcast(x)` returns lane 0's x on every lane
// (host: n_lanes = 1, both no-ops).
template <class Sync>
DG_HD int build_index(const Store &g, IndexBuf &w, bool with_paths, uint32_t lane, uint32_t n_lanes, Sync &&sync) {
    const uint32_t V = g.f->n_nodes;
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    if (V > w.cap.V || g.f->live_edges > w.cap.E) return DG_CAPACITY;
    // ---- CSR in adjacency (newest-edge-first) order: degrees, exclusive scan, fill ----
    for (uint32_t n = lane; n < V; n += n_lanes) {
        uint32_t a = 0, b = 0;
        for (uint32_t e = g.head_out[n]; e != NIL; e = g.e_next_out[e]) ++a;
        for (uint32_t e = g.head_in[n]; e != NIL; e = g.e_next_in[e]) ++b;
        w.succ_off[n + 1] = a, w.pred_off[n + 1] = b;
    }
    sync();
    if (lane == 0) {
        w.succ_off[0] = w.pred_off[0] = 0;
        for (uint32_t n = 0; n < V; ++n) w.succ_off[n + 1] += w.succ_off[n], w.pred_off[n + 1] += w.pred_off[n];
    }
    sync();
    for (uint32_t n = lane; n < V; n += n_lanes) {
        uint32_t k = w.succ_off[n];
        for (uint32_t e = g.head_out[n]; e != NIL; e = g.e_next_out[e]) w.succ[k++] = g.e_dst[e];
        k = w.pred_off[n];
        for (uint32_t e = g.head_in[n]; e != NIL; e = g.e_next_in[e]) w.pred[k++] = g.e_src[e];
        w.seen[n] = 0;
        w.pos[n] = 0;
        w.entrance_exit[n] = NONE;
        w.is_exit[n] = 0;
        w.opm[n] = 0;
        w.dist_min[n] = w.dist_max[n] = 0;
        w.top_of[n] = NONE;
        w.dist_of[n] = 0;
        w.bub_off[n + 1] = 0;
    }
    sync();
    int rc = DG_OK;
    if (lane == 0) {
        const uint32_t START = Store::START, END = Store::END;
        // reverse post-order of the DFS from start, children in adjacency order, marked when pushed (graphs/tools.rs:5-37)
        uint32_t n_order = 0;
        {
            uint32_t depth = 0;
            w.dfs[0] = START, w.dfs[1] = w.succ_off[START];
            depth = 1;
            while (depth) {
                uint32_t *top = w.dfs + 2 * (size_t)(depth - 1);
                const uint32_t end = w.succ_off[top[0] + 1];
                uint32_t child = NONE;
                while (top[1] != end) {
                    const uint32_t c = w.succ[top[1]++];
                    if (!w.seen[c]) {
                        child = c;
                        break;
                    }
                }
                if (child != NONE) {
                    w.seen[child] = 1;
                    w.dfs[2 * (size_t)depth] = child, w.dfs[2 * (size_t)depth + 1] = w.succ_off[child];
                    ++depth;
                } else {
                    w.order[n_order++] = top[0];
                    --depth;
                }
            }
            for (uint32_t i = 0; i < n_order / 2; ++i) {
                const uint32_t t = w.order[i];
                w.order[i] = w.order[n_order - 1 - i];
                w.order[n_order - 1 - i] = t;
            }
            for (uint32_t n = 0; n < V; ++n) w.from_start[n] = w.seen[n];
            w.from_start[START] = 1;
        }
        for (uint32_t i = 0; i < n_order; ++i) w.pos[w.order[i]] = i;
        // out_parent: smallest predecessor position (NONE = none); out_child: largest successor position (NONE = none)
        for (uint32_t n = 0; n < V; ++n) {
            uint32_t mn = NONE, mx = NONE;
            for (uint32_t k = w.pred_off[n]; k < w.pred_off[n + 1]; ++k) mn = w.pos[w.pred[k]] < mn ? w.pos[w.pred[k]] : mn;
            if (w.succ_off[n] != w.succ_off[n + 1]) {
                mx = 0;
                for (uint32_t k = w.succ_off[n]; k < w.succ_off[n + 1]; ++k) mx = w.pos[w.succ[k]] > mx ? w.pos[w.succ[k]] : mx;
            }
            w.out_parent[n] = mn, w.out_child[n] = mx;
        }
        // superbubble scan, high rank to low (finder.rs:109-177); later pairs overwrite earlier ones (index.rs:61-64)
        {
            uint32_t sp = 0, cand = NONE;
            auto pop = [&]() -> uint32_t { return sp ? w.stack[--sp] : NONE; };
            auto mn = [](int32_t a, int32_t b) { return a < b ? a : b; };
            for (int64_t curr = (int64_t)n_order - 1; curr >= 0; --curr) {
                const uint32_t n = w.order[curr];
                const int64_t far = w.out_child[n] == NONE ? INT64_MAX : (int64_t)w.out_child[n];
                if (far == curr + 1) {
                    if (cand != NONE) w.stack[sp++] = cand;
                    cand = w.order[curr + 1];
                } else {
                    while (cand != NONE) {
                        const uint32_t c0 = cand;
                        if (far <= (int64_t)w.pos[c0]) break;
                        cand = pop();
                        if (cand != NONE) w.opm[cand] = mn(w.opm[c0], w.opm[cand]);
                    }
                }
                if (cand != NONE && w.opm[cand] >= 0 && (int64_t)w.opm[cand] == curr) {
                    const uint32_t c0 = cand;
                    w.entrance_exit[n] = c0;
                    w.is_exit[c0] = 1;
                    cand = pop();
                    if (cand != NONE) w.opm[cand] = mn(w.opm[c0], w.opm[cand]);
                }
                w.opm[n] = w.out_parent[n] == NONE ? -1 : (int32_t)w.out_parent[n];
                if (cand != NONE) w.opm[cand] = mn(w.opm[n], w.opm[cand]);
            }
        }
        // backward BFS from end carrying the stack of open bubbles (index.rs:73-131); stacks share structure through `open`
        uint32_t n_ents = 0, n_open = 0;
        {
            for (uint32_t n = 0; n < V; ++n) w.seen[n] = 0;
            auto ent = [&](uint32_t node, uint32_t exit, uint32_t min_d) {
                if (n_ents >= w.cap.ents) {
                    rc = DG_CAPACITY;
                    return;
                }
                uint32_t *t = w.ents + 3 * (size_t)n_ents++;
                t[0] = node, t[1] = exit, t[2] = min_d;
                ++w.bub_off[node + 1];
            };
            if (w.is_exit[END]) {
                w.open[0] = 0, w.open[1] = END, w.open[2] = NONE;
                n_open = 1;
                w.top_of[END] = 0;
            }
            uint32_t qn = 0;
            w.queue[qn++] = END;
            w.seen[END] = 1;
            for (uint32_t head = 0; head < qn && rc == DG_OK; ++head) {
                const uint32_t node = w.queue[head];
                const uint32_t dist = w.dist_of[node], top = w.top_of[node];
                // the node's own stack, bottom to top: collect the chain top -> bottom in `stack`, emit reversed
                uint32_t nc = 0;
                for (uint32_t o = top; o != NONE; o = w.open[3 * (size_t)o + 2]) w.stack[nc++] = o;
                for (uint32_t i = nc; i-- > 0;) ent(node, w.open[3 * (size_t)w.stack[i] + 1], dist - w.open[3 * (size_t)w.stack[i]]);
                w.dist_min[node] = dist;
                for (uint32_t k = w.pred_off[node]; k < w.pred_off[node + 1]; ++k) {
                    const uint32_t p = w.pred[k];
                    if (w.seen[p]) continue;
                    const uint32_t nd = dist + 1;
                    uint32_t ns = top;
                    if (w.entrance_exit[p] != NONE && ns != NONE) {
                        ent(p, w.open[3 * (size_t)ns + 1], nd - w.open[3 * (size_t)ns]);
                        ns = w.open[3 * (size_t)ns + 2];
                    }
                    if (w.is_exit[p]) {
                        uint32_t *t = w.open + 3 * (size_t)n_open;
                        t[0] = nd, t[1] = p, t[2] = ns;
                        ns = n_open++;
                    }
                    w.seen[p] = 1;
                    w.top_of[p] = ns;
                    w.dist_of[p] = nd;
                    w.queue[qn++] = p;
                }
            }
        }
        if (rc == DG_OK) {
            for (uint32_t i = n_order; i-- > 0;) {  // longest distances in post-order (index.rs:133-148)
                const uint32_t n = w.order[i];
                uint32_t mx = 0;
                for (uint32_t k = w.succ_off[n]; k < w.succ_off[n + 1]; ++k) {
                    const uint32_t d = w.dist_max[w.succ[k]] + 1;
                    mx = d > mx ? d : mx;
                }
                w.dist_max[n] = mx;
            }
            // memberships grouped by node, generation order kept inside a node (stable counting sort)
            w.bub_off[0] = 0;
            for (uint32_t n = 0; n < V; ++n) {
                w.bub_off[n + 1] += w.bub_off[n];
                w.cursor[n] = w.bub_off[n];
            }
            for (uint32_t i = 0; i < n_ents; ++i) {
                const uint32_t *t = w.ents + 3 * (size_t)i;
                replay::BubQ &q = w.bub[w.cursor[t[0]]++];
                q.exit = t[1];
                q.min_d = t[2];
                q.max_d = w.from_start[t[0]] ? w.dist_max[t[0]] - w.dist_max[t[1]] : 0;
                q.mde = w.dist_min[t[1]] ? w.dist_min[t[1]] - 1 : 0;
            }
            w.n_bub = n_ents;
            w.n_path = 0;
            if (with_paths) {
                // path_index.rs:64-237 with max_paths_per_node = 10; the heuristic reads a node's first 5 entries, kept here
                // as 5 fixed slots per node (path_rem[5 n + i]) and compacted into CSR form afterwards
                const uint32_t E = w.succ_off[V];
                for (uint32_t k = 0; k < E; ++k) w.used[k] = 0;
                for (uint32_t n = 0; n < V; ++n) w.path_cnt[n] = 0, w.seen[n] = 0;  // seen = "has a used out-edge"
                uint32_t n_paths = 0;
                auto out_deg = [&](uint32_t n) { return w.succ_off[n + 1] - w.succ_off[n]; };
                auto in_deg = [&](uint32_t n) { return w.pred_off[n + 1] - w.pred_off[n]; };
                auto extract = [&](uint32_t start) -> uint32_t {  // nodes into path_tmp, returns length
                    uint32_t len = 0, cur = start;
                    w.path_tmp[len++] = cur;
                    while (cur != END && w.succ_off[cur] != w.succ_off[cur + 1]) {
                        uint32_t next_k = NONE, best = 0;
                        for (uint32_t k = w.succ_off[cur]; k < w.succ_off[cur + 1]; ++k) {
                            if (w.used[k]) continue;
                            const uint32_t od = out_deg(w.succ[k]);
                            if (next_k == NONE || od >= best) next_k = k, best = od;  // max_by_key keeps the last maximum
                        }
                        if (next_k == NONE) next_k = w.succ_off[cur];
                        // the reference marks the EDGE (cur, next): parallel edges do not exist, so marking every slot to `next` is the same
                        const uint32_t nx = w.succ[next_k];
                        for (uint32_t k = w.succ_off[cur]; k < w.succ_off[cur + 1]; ++k)
                            if (w.succ[k] == nx) w.used[k] = 1;
                        w.seen[cur] = 1;
                        if (len <= V) w.path_tmp[len] = nx;
                        ++len;
                        cur = nx;
                    }
                    return len;
                };
                auto add = [&](uint32_t len) {
                    for (uint32_t p = 0; p < len; ++p) {
                        const uint32_t n = w.path_tmp[p];
                        if (w.path_cnt[n] < 10) {
                            if (w.path_cnt[n] < 5) w.path_rem[5 * (size_t)n + w.path_cnt[n]] = len - 1 - p;
                            ++w.path_cnt[n];
                        }
                    }
                    ++n_paths;
                };
                // seeds: start, then every node in index order with in_degree == 0 || (out_degree > 2 && in_degree == 1)
                for (uint32_t si = 0; si <= V; ++si) {
                    const uint32_t s = si == 0 ? START : si - 1;
                    if (si > 0 && !(in_deg(s) == 0 || (out_deg(s) > 2 && in_deg(s) == 1))) continue;
                    if (w.seen[s]) continue;
                    const uint32_t len = extract(s);
                    if (len > 1) add(len);
                }
                if (n_paths < 10) {
                    // nodes with unused out-edges, stable-sorted by that count descending, first 20 (counts frozen before extraction)
                    uint32_t top_n[20], top_c[20], nt = 0;
                    for (uint32_t n = 0; n < V; ++n) {
                        uint32_t cnt = 0;
                        for (uint32_t k = w.succ_off[n]; k < w.succ_off[n + 1]; ++k) cnt += !w.used[k];
                        if (!cnt) continue;
                        uint32_t at = nt;  // insert after every entry with count >= cnt (stable, descending)
                        while (at > 0 && top_c[at - 1] < cnt) --at;
                        if (at >= 20) continue;
                        const uint32_t last = nt < 20 ? nt : 19;
                        for (uint32_t i = last; i > at; --i) top_n[i] = top_n[i - 1], top_c[i] = top_c[i - 1];
                        top_n[at] = n, top_c[at] = cnt;
                        if (nt < 20) ++nt;
                    }
                    for (uint32_t i = 0; i < nt; ++i) {
                        const uint32_t len = extract(top_n[i]);
                        if (len > 3) add(len);
                    }
                }
                // compact the 5-slot table into CSR form in place (slot k of node n moves left or stays)
                uint32_t np = 0;
                w.path_off[0] = 0;
                for (uint32_t n = 0; n < V; ++n) {
                    const uint32_t c = w.path_cnt[n] < 5 ? w.path_cnt[n] : 5;
                    for (uint32_t i = 0; i < c; ++i) w.path_rem[np++] = w.path_rem[5 * (size_t)n + i];
                    w.path_off[n + 1] = np;
                }
                w.n_path = np;
            }
        }
    }
    rc = sync.bcast(rc);  // lane 0's verdict, warp-uniform from here on
    if (rc != DG_OK) return rc;
    if (!with_paths)
        for (uint32_t n = lane; n <= V; n += n_lanes) w.path_off[n] = 0;
    // 32-byte node records of the search kernel
    for (uint32_t n = lane; n < V; n += n_lanes) {
        replay::NodeRec r;
        r.a.succ_off = w.succ_off[n];
        r.a.bub_off = w.bub_off[n];
        r.a.counts = (w.succ_off[n + 1] - w.succ_off[n]) | ((w.bub_off[n + 1] - w.bub_off[n]) << 16);
        r.a.symflags = (uint32_t)g.sym[n] | ((uint32_t)w.is_exit[n] << 8);
        r.b.dist_min = w.dist_min[n];
        r.b.dist_max = w.dist_max[n];
        r.b.pred_off = w.pred_off[n];
        r.b.n_pred = w.pred_off[n + 1] - w.pred_off[n];
        w.nodes[n] = r;
    }
    sync();
    return DG_OK;
}

}  // namespace dg
}  // namespace poasta
