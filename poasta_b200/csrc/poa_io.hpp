// Readers / writers on either side of the alignment path.
//   FASTA reader          : behaviour of noodles-fasta 0.45 as used at bin/poasta.rs:189-195
//   FASTA-MSA writer      : /root/reference/src/io/fasta.rs:19-156   (incl. the col-1 leading-gap quirk)
//   GFA (W-lines) / GFAv1 : /root/reference/src/io/graph.rs:245-372 / :374-502
//   FASTA-MSA import      : /root/reference/src/io/graph.rs:36-103
//   .poasta (bincode 1.3) : /root/reference/src/io/graph.rs:24-34 + poa.rs:84-94,484-489 (layout unpinned)
#pragma once
#include <charconv>
#include <cmath>
#include <cstring>
#include <deque>
#include <sstream>
#include <unordered_map>

#include "poa_graph.hpp"

namespace poasta {

struct FastaRecord {
    std::string name;
    std::string seq;
};

// name = text after '>' up to first whitespace; sequence lines concatenated (trailing CR/LF stripped).
inline std::vector<FastaRecord> parse_fasta(const char *data, size_t len) {
    std::vector<FastaRecord> out;
    size_t i = 0;
    while (i < len) {
        size_t eol = i;
        while (eol < len && data[eol] != '\n') ++eol;
        size_t end = eol;
        if (end > i && data[end - 1] == '\r') --end;
        if (end > i && data[i] == '>') {
            size_t j = i + 1;
            while (j < end && data[j] != ' ' && data[j] != '\t') ++j;
            out.push_back({std::string(data + i + 1, j - i - 1), std::string()});
        } else if (end > i && !out.empty()) {
            out.back().seq.append(data + i, end - i);
        }
        i = eol + 1;
    }
    return out;
}

// ---- FASTA MSA ---------------------------------------------------------------------------
inline bool edge_has_seq(const POAGraph::Edge &e, uint64_t seq_id) {
    return std::binary_search(e.sequence_ids.begin(), e.sequence_ids.end(), seq_id);
}

// io/fasta.rs:69-137: column per node via DFS post-order where successors are consumed oldest-first.
inline std::vector<uint32_t> msa_columns(const POAGraph &g, uint32_t &n_cols) {
    const uint32_t n = g.node_count_with_start_and_end();
    std::vector<uint32_t> col(n, NIL);
    std::vector<uint8_t> visited(n, 0);
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> stack;
    auto succ_vec = [&](uint32_t v, std::vector<uint32_t> &out) { g.for_successors(v, [&](uint32_t s) { out.push_back(s); }); };
    stack.emplace_back(g.start_node, std::vector<uint32_t>());
    succ_vec(g.start_node, stack.back().second);
    std::vector<uint32_t> rev_postorder;
    while (!stack.empty()) {
        auto &frame = stack.back().second;
        uint32_t child = NIL;
        while (!frame.empty()) {
            uint32_t s = frame.back();
            frame.pop_back();
            if (!visited[s]) {
                child = s;
                break;
            }
        }
        if (child != NIL) {
            visited[child] = 1;
            std::vector<uint32_t> successors;
            succ_vec(child, successors);
            for (uint32_t a : g.nodes[child].aligned_nodes)
                if (!visited[a]) {
                    visited[a] = 1;
                    succ_vec(a, successors);
                }
            stack.emplace_back(child, std::move(successors));
        } else {
            rev_postorder.push_back(stack.back().first);
            stack.pop_back();
        }
    }
    std::reverse(rev_postorder.begin(), rev_postorder.end());
    uint32_t curr = 0;
    for (uint32_t v : rev_postorder) {
        if (v == g.start_node || v == g.end_node) continue;
        if (col[v] == NIL) {
            col[v] = curr;
            for (uint32_t a : g.nodes[v].aligned_nodes) col[a] = curr;
            ++curr;
        }
    }
    n_cols = curr;
    return col;
}

inline bool graph_to_fasta(const POAGraph &g, std::string &out, size_t line_width = 80) {
    uint32_t n_cols = 0;
    std::vector<uint32_t> col = msa_columns(g, n_cols);
    bool have_cols = false;
    uint32_t max_col = 0;
    for (uint32_t c : col)
        if (c != NIL) {
            have_cols = true;
            max_col = std::max(max_col, c);
        }
    for (size_t seq_id = 0; seq_id < g.sequences.size(); ++seq_id) {
        std::string row;
        uint32_t curr = g.sequences[seq_id].start_node;
        uint64_t last_col = 0;
        bool empty_row = false;
        while (curr != NIL) {
            if (col[curr] == NIL) {  // start node of an empty sequence (io/fasta.rs:33-41)
                empty_row = true;
                break;
            }
            uint64_t node_col = col[curr];
            uint64_t sat = node_col == 0 ? 0 : node_col - 1;
            if (sat < last_col) return false;  // reference would wrap (release) / panic (debug)
            row.append(sat - last_col, '-');
            row.push_back((char)g.nodes[curr].symbol);
            uint32_t next = NIL;
            for (uint32_t e = g.nodes[curr].head[0]; e != NIL; e = g.edges[e].next[0])
                if (edge_has_seq(g.edges[e], seq_id)) next = g.edges[e].node[1];
            curr = next;
            last_col = node_col;
        }
        if (empty_row) row.clear();
        else if (have_cols) row.append(max_col - last_col, '-');
        out.push_back('>');
        out += g.sequences[seq_id].name;
        out.push_back('\n');
        for (size_t p = 0; p < row.size(); p += line_width) {
            out.append(row, p, std::min(line_width, row.size() - p));
            out.push_back('\n');
        }
    }
    return true;
}

// ---- GFA -----------------------------------------------------------------------------------
// io/graph.rs:245-372 (v1_paths = false) and :374-502 (v1_paths = true)
inline bool graph_to_gfa(const POAGraph &g, std::string &out, bool v1_paths) {
    const uint32_t n = g.node_count_with_start_and_end();
    std::vector<uint8_t> visited(n, 0);
    std::deque<uint32_t> queue;
    queue.push_back(g.start_node);
    visited[g.start_node] = 1;
    out += "H\tVN:Z:1.1\n";
    std::vector<uint32_t> seg_of(n, NIL), pos_of(n, 0), seg_start(n, NIL), seg_end(n, NIL);
    std::vector<size_t> seg_len;
    uint32_t seg_id = 0;
    auto seg_name = [&](uint32_t s) { return v1_paths ? std::to_string(s + 1) : "s" + std::to_string(s); };
    while (!queue.empty()) {
        uint32_t front = queue.front();
        queue.pop_front();
        if (front == g.start_node) {
            g.for_successors(front, [&](uint32_t s) {
                if (!visited[s]) {
                    queue.push_back(s);
                    visited[s] = 1;
                }
            });
            continue;
        }
        std::string segment(1, (char)g.nodes[front].symbol);
        uint32_t curr = front;
        uint32_t out_deg = g.out_degree(front);
        uint32_t seg_pos = 0;
        seg_of[front] = seg_id;
        pos_of[front] = seg_pos;
        seg_start[front] = seg_id;
        while (out_deg == 1) {
            uint32_t next = g.edges[g.nodes[curr].head[0]].node[1];
            if (g.in_degree(next) == 1 && next != g.end_node) {
                segment.push_back((char)g.nodes[next].symbol);
                seg_of[next] = seg_id;
                pos_of[next] = seg_pos;
            } else {
                break;
            }
            curr = next;
            out_deg = g.out_degree(curr);
            ++seg_pos;
        }
        out += "S\t" + seg_name(seg_id) + "\t" + segment + "\n";
        seg_end[curr] = seg_id;
        seg_len.push_back(segment.size());
        visited[curr] = 1;
        g.for_successors(curr, [&](uint32_t s) {
            if (!visited[s] && s != g.end_node) {
                visited[s] = 1;
                queue.push_back(s);
            }
        });
        ++seg_id;
    }
    for (const POAGraph::Edge &e : g.edges) {
        if (e.vacant) continue;
        if (seg_end[e.node[0]] != NIL && seg_start[e.node[1]] != NIL)
            out += "L\t" + seg_name(seg_end[e.node[0]]) + "\t+\t" + seg_name(seg_start[e.node[1]]) + "\t+\t0M\n";
    }
    for (size_t seq_id = 0; seq_id < g.sequences.size(); ++seq_id) {
        uint32_t curr = g.sequences[seq_id].start_node;
        if (seg_of[curr] == NIL) return false;  // reference panics on the missing map key
        uint32_t prev_segment = seg_of[curr];
        uint32_t start_pos = pos_of[curr], last_pos = 0;
        std::vector<uint32_t> walk{prev_segment};
        size_t total = seg_len[prev_segment];
        while (curr != NIL) {
            if (seg_of[curr] == NIL) return false;
            uint32_t node_segment = seg_of[curr];
            last_pos = pos_of[curr];
            if (node_segment != prev_segment) {
                walk.push_back(node_segment);
                total += seg_len[node_segment];
            }
            uint32_t next = NIL;
            for (uint32_t e = g.nodes[curr].head[0]; e != NIL; e = g.edges[e].next[0])
                if (edge_has_seq(g.edges[e], seq_id)) next = g.edges[e].node[1];
            curr = next;
            prev_segment = node_segment;
        }
        if (v1_paths) {
            out += "P\t" + g.sequences[seq_id].name + "\t";
            for (size_t i = 0; i < walk.size(); ++i) {
                if (i) out.push_back(',');
                out += std::to_string(walk[i] + 1) + "+";
            }
            out += "\t*\n";
        } else {
            size_t end_pos = total - seg_len[prev_segment] + last_pos;
            out += "W\t*\t0\t" + g.sequences[seq_id].name + "\t" + std::to_string(start_pos) + "\t" +
                   std::to_string(end_pos) + "\t";
            for (uint32_t s : walk) out += ">s" + std::to_string(s);
            out.push_back('\n');
        }
    }
    return true;
}

// ---- DOT ----------------------------------------------------------------------------------------
// `view -O dot`: io/graph.rs:504-609 (hand-written dialect). f64 Display = shortest round-trip digits.
inline std::string fmt_f64(double v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, r.ptr);
}
inline const char *graphviz_node_color(uint8_t c) {
    switch (c) {
        case 'A': return "#80BC42";
        case 'C': return "#006DB6";
        case 'G': return "#F36C3E";
        case 'T': return "#B12028";
        default: return "#939393";
    }
}
inline void graph_to_dot_view(const POAGraph &g, std::string &out) {
    out += "# seq:\t";
    for (size_t i = 0; i < g.sequences.size(); ++i) {
        if (i) out.push_back('\t');
        out += g.sequences[i].name + ":" + std::to_string(g.sequences[i].start_node);
    }
    out += "\ndigraph {\nrankdir=\"LR\"\nnode [shape=square, style=filled, fillcolor=\"#e3e3e3\", penwidth=0]\n\n";
    const uint32_t n = g.node_count_with_start_and_end();
    for (uint32_t v = 0; v < n; ++v)
        out += std::to_string(v) + " [label=\"" + std::string(1, (char)g.nodes[v].symbol) + "\"; fontcolor=\"" +
               graphviz_node_color(g.nodes[v].symbol) + "\"]\n";
    std::vector<uint8_t> done(n, 0);
    for (uint32_t v = 0; v < n; ++v) {
        if (done[v]) continue;
        done[v] = 1;
        if (!g.nodes[v].aligned_nodes.empty()) {
            out += "{rank=same; " + std::to_string(v);
            for (uint32_t a : g.nodes[v].aligned_nodes) out += "; " + std::to_string(a);
            out += "}\n";
        }
        for (uint32_t a : g.nodes[v].aligned_nodes) done[a] = 1;
    }
    size_t max_num_seq = 0;
    bool any_edge = false;
    for (const auto &e : g.edges)
        if (!e.vacant) {
            max_num_seq = std::max(max_num_seq, e.sequence_ids.size());
            any_edge = true;
        }
    if (!any_edge) max_num_seq = 1;
    for (const auto &e : g.edges) {
        if (e.vacant) continue;
        const double frac = (double)e.sequence_ids.size() / (double)max_num_seq;
        const long long weight = (long long)std::round(1.0 + frac * (40.0 - 1.0));
        const double pen = 0.5 + frac * (3.5 - 0.5);
        out += std::to_string(e.node[0]) + " -> " + std::to_string(e.node[1]) + " [weight=" + std::to_string(weight) +
               "; penwidth=" + fmt_f64(pen) + "; label=" + std::to_string(e.sequence_ids.size()) + "; class=\"";
        for (size_t i = 0; i < e.sequence_ids.size(); ++i) {
            if (i) out.push_back(' ');
            out += "s" + std::to_string(e.sequence_ids[i]);
        }
        out += "\"]\n";
    }
    out += "}\n";
}

// `align -O dot`: petgraph 0.6.5 `Dot` of the relabelled graph (io/graph.rs:229-243, poa.rs:473-499);
// node label = format!("{:?} ({:?})", char, index), edge label = format!("{}, {:?}", weight, sequence_ids),
// labels escaped the way petgraph's Display-escaper does (quote, backslash, newline). Layout restated from the
// crate's published output format (unpinned).
inline void graph_to_dot_petgraph(const POAGraph &g, std::string &out) {
    auto esc = [](const std::string &s) {
        std::string r;
        for (char c : s) {
            if (c == '"' || c == '\\') r.push_back('\\');
            if (c == '\n') {
                r += "\\l";
                continue;
            }
            r.push_back(c);
        }
        return r;
    };
    out += "digraph {\n";
    for (uint32_t v = 0; v < g.nodes.size(); ++v) {
        std::string label = std::string("'") + (char)g.nodes[v].symbol + "' (" + std::to_string(v) + ")";
        out += "    " + std::to_string(v) + " [ label = \"" + esc(label) + "\" ]\n";
    }
    for (const auto &e : g.edges) {
        if (e.vacant) continue;
        std::string label = std::to_string(e.weight) + ", [";
        for (size_t i = 0; i < e.sequence_ids.size(); ++i) {
            if (i) label += ", ";
            label += std::to_string(e.sequence_ids[i]);
        }
        label += "]";
        out += "    " + std::to_string(e.node[0]) + " -> " + std::to_string(e.node[1]) + " [ label = \"" + esc(label) + "\" ]\n";
    }
    out += "}\n\n";
}

// ---- FASTA-MSA import (io/graph.rs:36-103) ---------------------------------------------------
inline GraphError graph_from_fasta_msa(const std::vector<FastaRecord> &recs, POAGraph &g) {
    std::vector<std::vector<uint32_t>> nodes_per_col;
    for (size_t seq_id = 0; seq_id < recs.size(); ++seq_id) {
        const std::string &s = recs[seq_id].seq;
        if (s.size() > nodes_per_col.size()) nodes_per_col.resize(s.size());
        uint32_t prev = NIL;
        for (size_t c = 0; c < s.size(); ++c) {
            if (s[c] == '-') continue;
            uint32_t node = NIL;
            for (uint32_t v : nodes_per_col[c])
                if (g.nodes[v].symbol == (uint8_t)s[c]) {
                    node = v;
                    break;
                }
            if (node == NIL) {
                node = g.add_node((uint8_t)s[c]);
                for (uint32_t other : nodes_per_col[c]) {
                    g.nodes[other].aligned_nodes.push_back(node);
                    g.nodes[node].aligned_nodes.push_back(other);
                }
                nodes_per_col[c].push_back(node);
            }
            if (prev != NIL) g.add_edge(prev, node, seq_id, 2);
            else g.sequences.push_back({recs[seq_id].name, node});
            prev = node;
        }
    }
    return g.post_process();
}

// ---- .poasta: bincode 1.3 default config (little endian, fixed-width ints, u64 lengths) ------
struct BinWriter {
    std::string &o;
    void u8(uint8_t v) { o.push_back((char)v); }
    void u32(uint32_t v) { o.append((const char *)&v, 4); }
    void u64(uint64_t v) { o.append((const char *)&v, 8); }
};
struct BinReader {
    const uint8_t *p, *end;
    bool ok = true;
    bool need(size_t n) {
        if ((size_t)(end - p) < n) ok = false;
        return ok;
    }
    uint8_t u8() {
        if (!need(1)) return 0;
        return *p++;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    uint64_t u64() {
        if (!need(8)) return 0;
        uint64_t v;
        memcpy(&v, p, 8);
        p += 8;
        return v;
    }
};

// enum POAGraphWithIx::U32 (tag 2) { graph{nodes,node_holes,edge_property,edges}, sequences, topological_sorted, start, end }
inline void graph_to_poasta(const POAGraph &g, std::string &out) {
    BinWriter w{out};
    w.u32(2);
    w.u64(g.nodes.size());
    for (const auto &nd : g.nodes) {
        w.u8(nd.symbol);
        w.u64(nd.aligned_nodes.size());
        for (uint32_t a : nd.aligned_nodes) w.u32(a);
    }
    w.u64(0);  // node_holes
    w.u32(1);  // petgraph serialization enum EdgeProperty { Undirected = 0, Directed = 1 }
    w.u64(g.edges.size());
    for (const auto &e : g.edges) {
        if (e.vacant) {
            w.u8(0);
            continue;
        }
        w.u8(1);
        w.u32(e.node[0]);
        w.u32(e.node[1]);
        w.u64(e.weight);
        w.u64(e.sequence_ids.size());
        for (uint64_t s : e.sequence_ids) w.u64(s);
    }
    w.u64(g.sequences.size());
    for (const auto &s : g.sequences) {
        w.u64(s.name.size());
        out += s.name;
        w.u32(s.start_node);
    }
    w.u64(g.topological_sorted.size());
    for (uint32_t v : g.topological_sorted) w.u32(v);
    w.u32(g.start_node);
    w.u32(g.end_node);
}

// Reads a .poasta file. The file is UNTRUSTED input (`view`, `stats`, `align -I`): every count is bounded by the bytes that
// are left before anything is allocated, every node id is checked against the node count, and the topological order is
// recomputed instead of trusted (the reference's bincode + petgraph deserialisers reject or panic on such files).
inline bool graph_from_poasta(const uint8_t *data, size_t len, POAGraph &g) {
    BinReader r{data, data + len};
    auto left = [&]() { return (uint64_t)(r.end - r.p); };
    if (r.u32() != 2) return false;  // only the U32 variant is produced by the CLI (bin/poasta.rs:316)
    g = POAGraph();
    g.nodes.clear();
    const uint64_t nn = r.u64();
    if (!r.ok || nn < 2 || nn > (1ull << 31) || nn > left() / 9) return false;  // a node is at least 9 bytes
    for (uint64_t i = 0; i < nn; ++i) {
        POAGraph::Node nd{r.u8(), {NIL, NIL}, {}};
        const uint64_t na = r.u64();
        if (!r.ok || na > nn || na > left() / 4) return false;
        for (uint64_t k = 0; k < na; ++k) {
            const uint32_t a = r.u32();
            if (a >= nn) return false;
            nd.aligned_nodes.push_back(a);
        }
        g.nodes.push_back(std::move(nd));
    }
    if (r.u64() != 0) return false;  // node holes never occur (nodes are never removed)
    r.u32();
    const uint64_t ne = r.u64();
    if (!r.ok || ne > (1ull << 32) - 2 || ne > left()) return false;  // a slot is at least 1 byte
    // Rebuild adjacency exactly as petgraph's deserializer does: edges re-added in slot order
    // (so list heads are the highest slots), vacant slots chained on the free list afterwards.
    g.edges.clear();
    g.free_edge = NIL;
    g.live_edges = 0;
    g.edges.resize(ne);
    for (uint64_t i = 0; i < ne; ++i) {
        POAGraph::Edge &e = g.edges[i];
        const uint8_t tag = r.u8();
        if (!r.ok || tag > 1) return false;
        if (tag == 0) {
            e.vacant = true;
            e.node[0] = e.node[1] = NIL;
            e.next[0] = g.free_edge;
            e.next[1] = NIL;
            e.weight = 0;
            g.free_edge = (uint32_t)i;
            continue;
        }
        e.vacant = false;
        e.node[0] = r.u32();
        e.node[1] = r.u32();
        e.weight = r.u64();
        const uint64_t ns = r.u64();
        if (!r.ok || e.node[0] >= nn || e.node[1] >= nn || ns > left() / 8) return false;
        for (uint64_t k = 0; k < ns; ++k) e.sequence_ids.push_back(r.u64());
        e.next[0] = g.nodes[e.node[0]].head[0];
        e.next[1] = g.nodes[e.node[1]].head[1];
        g.nodes[e.node[0]].head[0] = (uint32_t)i;
        g.nodes[e.node[1]].head[1] = (uint32_t)i;
        ++g.live_edges;
    }
    const uint64_t nsq = r.u64();
    if (!r.ok || nsq > left() / 12) return false;  // a sequence is at least 12 bytes
    for (uint64_t i = 0; i < nsq; ++i) {
        const uint64_t nl = r.u64();
        if (!r.ok || nl > left()) return false;
        std::string name((const char *)r.p, nl);
        r.p += nl;
        const uint32_t first = r.u32();
        if (!r.ok || first >= nn) return false;
        g.sequences.push_back({std::move(name), first});
    }
    const uint64_t nt = r.u64();
    if (!r.ok || nt > left() / 4 || (nt != 0 && nt != nn)) return false;
    std::vector<uint32_t> file_order;
    for (uint64_t i = 0; i < nt; ++i) file_order.push_back(r.u32());
    g.start_node = r.u32();
    g.end_node = r.u32();
    if (!r.ok || g.start_node >= nn || g.end_node >= nn || g.start_node == g.end_node) return false;
    if (nt) {
        // a permutation that respects every edge: anything else would index out of bounds (or misorder the DP rows) later
        std::vector<uint32_t> rank(nn, NIL);
        for (uint64_t i = 0; i < nt; ++i) {
            if (file_order[i] >= nn || rank[file_order[i]] != NIL) return false;
            rank[file_order[i]] = (uint32_t)i;
        }
        for (const auto &e : g.edges)
            if (!e.vacant && rank[e.node[0]] >= rank[e.node[1]]) return false;
        g.topological_sorted = std::move(file_order);
    }
    return true;
}

}  // namespace poasta
