// ORACLE — TEST INFRASTRUCTURE ONLY.
// Dense CPU model of the recurrences the CUDA kernels implement: POASTA's
// transition rules (aligner/dfa.rs:210-250 + aligner/scoring/gap_affine.rs:250-430) written in
// pull form over full M/I/D tables, followed by the reference's score-equality backtrace
// (gap_affine.rs:550-657, 804-915) on those tables.  Used by tests to (a) localise differences
// between the A* restatement and the device kernels and (b) give expected values for the
// per-cell traceback words.  Never used by the product.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../include/poasta_b200.h"

namespace dpmodel {

constexpr uint32_t INF = 0x3FFFFFFFu;

struct Tables {
    uint32_t V = 0, Q = 0;
    std::vector<uint32_t> M, I, D;  // indexed [node * (Q+1) + j] by NODE ID (not rank)
    uint32_t &m(uint32_t v, uint32_t j) { return M[(size_t)v * (Q + 1) + j]; }
    uint32_t &i(uint32_t v, uint32_t j) { return I[(size_t)v * (Q + 1) + j]; }
    uint32_t &d(uint32_t v, uint32_t j) { return D[(size_t)v * (Q + 1) + j]; }
};

inline void fill(const poasta_graph_view &g, const uint8_t *q, uint32_t Q, uint32_t x, uint32_t o, uint32_t e, Tables &t) {
    const uint32_t V = g.n_nodes, oe = o + e;
    t.V = V;
    t.Q = Q;
    t.M.assign((size_t)V * (Q + 1), INF);
    t.I.assign((size_t)V * (Q + 1), INF);
    t.D.assign((size_t)V * (Q + 1), INF);
    auto sat = [](uint32_t a, uint32_t b) { return a >= INF ? INF : a + b; };
    for (uint32_t r = 0; r < V; ++r) {
        const uint32_t v = g.topo_order[r];
        const bool is_end = v == g.end_node, is_start = v == g.start_node;
        const uint32_t *pb = g.pred + g.pred_off[v], *pe = g.pred + g.pred_off[v + 1];
        const uint32_t *sb = g.succ + g.succ_off[v], *se = g.succ + g.succ_off[v + 1];
        for (uint32_t j = 0; j <= Q; ++j) {
            // D[v][j]
            uint32_t dval = INF;
            const bool openD_ok = !is_end && (j >= Q || g.symbol[v] != q[j]);
            for (const uint32_t *p = pb; p != pe; ++p) {
                if (openD_ok) dval = std::min(dval, sat(t.m(*p, j), oe));
                dval = std::min(dval, sat(t.d(*p, j), e));
            }
            t.d(v, j) = dval;
            // diagonal / end transition
            uint32_t mval = INF;
            if (is_start && j == 0) mval = 0;
            if (is_end) {
                for (const uint32_t *p = pb; p != pe; ++p) mval = std::min(mval, t.m(*p, j));
            } else if (j > 0) {
                const uint32_t sub = (g.symbol[v] == q[j - 1]) ? 0 : x;
                for (const uint32_t *p = pb; p != pe; ++p) mval = std::min(mval, sat(t.m(*p, j - 1), sub));
            }
            // I[v][j]
            uint32_t ival = INF;
            if (j > 0 && !is_end) {
                bool openI_ok = false;
                for (const uint32_t *s = sb; s != se; ++s)
                    if (*s == g.end_node || g.symbol[*s] != q[j - 1]) openI_ok = true;
                if (openI_ok) ival = std::min(ival, sat(t.m(v, j - 1), oe));
                ival = std::min(ival, sat(t.i(v, j - 1), e));
            }
            t.i(v, j) = ival;
            t.m(v, j) = std::min(mval, std::min(dval, ival));
        }
    }
}

// One backtrace step on the dense tables (gap_affine.rs:550-657). state: 0=M,1=D,2=I. Returns false for None.
inline bool get_backtrace(const poasta_graph_view &g, const uint8_t *q, uint32_t x, uint32_t o, uint32_t e, Tables &t,
                          uint32_t node, uint32_t off, int st, uint32_t &bn, uint32_t &bo, int &bs) {
    auto val = [&](uint32_t n, uint32_t j, int s) -> uint32_t {
        uint32_t r = s == 0 ? t.m(n, j) : (s == 1 ? t.d(n, j) : t.i(n, j));
        return r >= INF ? 0xFFFFFFFFu : r;
    };
    const uint32_t cur = val(node, off, st);
    if (cur == 0xFFFFFFFFu) return false;
    auto eq = [](uint32_t ps, uint32_t target) { return ps != 0xFFFFFFFFu && ps == target; };
    const uint32_t *pb = g.pred + g.pred_off[node], *pe = g.pred + g.pred_off[node + 1];
    if (st == 0) {
        if (off > 0) {
            const bool me = node == g.end_node || g.symbol[node] == q[off - 1];
            const uint32_t po = node == g.end_node ? off : off - 1;
            for (const uint32_t *p = pb; p != pe; ++p) {
                uint32_t ps = val(*p, po, 0);
                if ((me && eq(ps, cur)) || (!me && eq(ps, cur - x))) {
                    bn = *p, bo = po, bs = 0;
                    return true;
                }
            }
        }
        if (eq(val(node, off, 1), cur)) {
            bn = node, bo = off, bs = 1;
            return true;
        }
        if (eq(val(node, off, 2), cur)) {
            bn = node, bo = off, bs = 2;
            return true;
        }
    } else if (st == 1) {
        for (const uint32_t *p = pb; p != pe; ++p)
            if (eq(val(*p, off, 0), cur - o - e)) {
                bn = *p, bo = off, bs = 0;
                return true;
            }
        for (const uint32_t *p = pb; p != pe; ++p)
            if (eq(val(*p, off, 1), cur - e)) {
                bn = *p, bo = off, bs = 1;
                return true;
            }
    } else if (off > 0) {
        if (eq(val(node, off - 1, 0), cur - o - e)) {
            bn = node, bo = off - 1, bs = 0;
            return true;
        }
        if (eq(val(node, off - 1, 2), cur - e)) {
            bn = node, bo = off - 1, bs = 0;  // sic (gap_affine.rs:647-650)
            return true;
        }
    }
    return false;
}

// Global-mode alignment: returns score (INF if unreachable) and pairs.
inline uint32_t align_global(const poasta_graph_view &g, const uint8_t *q, uint32_t Q, uint32_t x, uint32_t o, uint32_t e,
                             std::vector<uint32_t> &rpos, std::vector<uint32_t> &qpos, Tables *keep = nullptr) {
    Tables local;
    Tables &t = keep ? *keep : local;
    fill(g, q, Q, x, o, e, t);
    rpos.clear();
    qpos.clear();
    const uint32_t score = t.m(g.end_node, Q);
    if (Q == 0 || score >= INF) return score;
    if (Q == 1) {  // gap_affine.rs:812-824: the end node matches any symbol, so a 1-base global query pairs with `end`
        rpos.push_back(g.end_node);
        qpos.push_back(0);
        return score;
    }
    uint32_t cn, co, bn, bo;
    int cs, bs;
    if (!get_backtrace(g, q, x, o, e, t, g.end_node, Q, 0, cn, co, cs)) return score;
    while (get_backtrace(g, q, x, o, e, t, cn, co, cs, bn, bo, bs)) {
        if (cs == 0 && bs != 0) {
            cn = bn, co = bo, cs = bs;
            continue;
        }
        if (cs == 0) {
            rpos.push_back(cn);
            qpos.push_back(co - 1);
        } else if (cs == 2) {
            rpos.push_back(0xFFFFFFFFu);
            qpos.push_back(co - 1);
        } else {
            rpos.push_back(cn);
            qpos.push_back(0xFFFFFFFFu);
        }
        if (bn == g.start_node) break;
        cn = bn, co = bo, cs = bs;
    }
    std::reverse(rpos.begin(), rpos.end());
    std::reverse(qpos.begin(), qpos.end());
    return score;
}

}  // namespace dpmodel
