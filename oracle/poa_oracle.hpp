// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under poasta_b200/ may include, link or call
// this file, and this file includes nothing from poasta_b200/ (round 1's oracle shared the product's graph and writer
// headers, so the whole-MSA byte comparisons pinned those against themselves; this is the independent restatement).
//
// Literal CPU restatement of the reference's graph side of the path, with std containers:
//   graphs/poa.rs:84-382      POAGraph over petgraph 0.6.5 StableDiGraph (add_alignment_with_weights, post_process)
//   io/fasta.rs:19-156        column FASTA MSA writer (noodles-fasta 0.45 writer: 80-column wrap)
//   io/graph.rs:24-103        .poasta (bincode 1.3 of POAGraphWithIx::U32) save / load, column-FASTA import
//   io/graph.rs:245-502       GFA (W lines) and GFAv1 (P lines) writers
//   bin/poasta.rs:174-195     FASTA input (noodles reader: id up to whitespace, multi-line sequences)
// petgraph / noodles / bincode are un-vendored crates (Cargo.lock: petgraph 0.6.5, noodles-fasta 0.45.0, bincode 1.3.3);
// their behaviour is restated from the crates' published semantics.
// PARITY STATUS: pinned by the reference's own exact-value tests (tests/io_fasta.rs:17-32, src/io/fasta.rs:165-197,
// graphs/tools.rs:54-69 neighbour order); whole-MSA output "parity unpinned" - no Rust toolchain here to run the binary.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <optional>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/poasta_b200.h"

namespace oracle {
namespace poa {

constexpr uint32_t END = 0xFFFFFFFFu;  // petgraph's EdgeIndex::end() / "no index"

// ---- petgraph::stable_graph::StableGraph<N, E, Directed, u32> ---------------------------------------------------
// Nodes and edges are slots in two Vecs; a node's `next` heads its outgoing / incoming edge lists, an edge's `next` links the
// lists. add_edge links at the list heads; remove_edge unlinks and chains the slot on `free_edge` (LIFO) through next[0];
// add_edge takes a free slot first. (Nodes are never removed on this path: no node holes.)
template <class N, class E> struct StableDiGraph {
    struct Node {
        N weight;
        uint32_t next[2];
    };
    struct Edge {
        std::optional<E> weight;
        uint32_t next[2];
        uint32_t node[2];
    };
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    uint32_t free_edge = END;
    size_t edge_count_ = 0;

    uint32_t add_node(N w) {
        nodes.push_back(Node{std::move(w), {END, END}});
        return (uint32_t)nodes.size() - 1;
    }
    uint32_t add_edge(uint32_t a, uint32_t b, E w) {
        uint32_t idx;
        if (free_edge != END) {
            idx = free_edge;
            free_edge = edges[idx].next[0];
        } else {
            idx = (uint32_t)edges.size();
            edges.push_back(Edge{});
        }
        Edge &e = edges[idx];
        e.weight = std::move(w);
        e.node[0] = a, e.node[1] = b;
        e.next[0] = nodes[a].next[0];
        e.next[1] = nodes[b].next[1];
        nodes[a].next[0] = idx;
        nodes[b].next[1] = idx;
        ++edge_count_;
        return idx;
    }
    void remove_edge(uint32_t idx) {
        Edge &e = edges[idx];
        for (int d = 0; d < 2; ++d) {  // change_edge_links
            uint32_t *cur = &nodes[e.node[d]].next[d];
            while (*cur != idx) cur = &edges[*cur].next[d];
            *cur = e.next[d];
        }
        e.next[0] = free_edge, e.next[1] = END;
        e.node[0] = e.node[1] = END;
        e.weight.reset();
        free_edge = idx;
        --edge_count_;
    }
    std::optional<uint32_t> find_edge(uint32_t a, uint32_t b) const {
        for (uint32_t e = nodes[a].next[0]; e != END; e = edges[e].next[0])
            if (edges[e].node[1] == b) return e;
        return std::nullopt;
    }
    // edges_directed(n, dir): edge ids in list order (newest first)
    std::vector<uint32_t> edges_directed(uint32_t n, int dir) const {
        std::vector<uint32_t> out;
        for (uint32_t e = nodes[n].next[dir]; e != END; e = edges[e].next[dir]) out.push_back(e);
        return out;
    }
    std::vector<uint32_t> neighbors_directed(uint32_t n, int dir) const {
        std::vector<uint32_t> out;
        for (uint32_t e = nodes[n].next[dir]; e != END; e = edges[e].next[dir]) out.push_back(edges[e].node[1 - dir]);
        return out;
    }
    size_t node_count() const { return nodes.size(); }
    size_t edge_count() const { return edge_count_; }
};

// petgraph::algo::toposort (0.6.5): DFS from every undiscovered node, highest index first; a node is pushed on first visit
// together with its undiscovered successors, appended to the finish stack on second visit; result reversed. The second pass
// (a DFS over the REVERSED graph in that order; any node reached from a later root is a cycle) is what turns a cycle into Err.
template <class G> bool toposort(const G &g, std::vector<uint32_t> &order) {
    const size_t n = g.node_count();
    std::vector<char> discovered(n, 0), finished(n, 0);
    std::vector<uint32_t> stack, finish_stack;
    for (size_t ii = n; ii-- > 0;) {
        const uint32_t i = (uint32_t)ii;
        if (discovered[i]) continue;
        stack.push_back(i);
        while (!stack.empty()) {
            const uint32_t nx = stack.back();
            if (!discovered[nx]) {
                discovered[nx] = 1;
                for (uint32_t succ : g.neighbors_directed(nx, 0)) {
                    if (succ == nx) return false;
                    if (!discovered[succ]) stack.push_back(succ);
                }
            } else {
                stack.pop_back();
                if (!finished[nx]) {
                    finished[nx] = 1;
                    finish_stack.push_back(nx);
                }
            }
        }
    }
    std::reverse(finish_stack.begin(), finish_stack.end());
    std::fill(discovered.begin(), discovered.end(), 0);
    for (uint32_t i : finish_stack) {  // dfs.move_to(i); while let Some(j) = dfs.next(Reversed(g)) { if cycle { Err } cycle = true }
        if (discovered[i]) continue;  // (dfs.next skips nodes it has discovered; the first node it yields is i itself)
        stack.assign(1, i);
        bool cycle = false;
        while (!stack.empty()) {
            const uint32_t j = stack.back();
            stack.pop_back();
            if (discovered[j]) continue;
            discovered[j] = 1;
            for (uint32_t p : g.neighbors_directed(j, 1))
                if (!discovered[p]) stack.push_back(p);
            if (cycle) return false;
            cycle = true;
        }
    }
    order = std::move(finish_stack);
    return true;
}

// ---- graphs/poa.rs ------------------------------------------------------------------------------------------------
struct NodeData {  // POANodeData (poa.rs:40-55)
    uint8_t symbol;
    std::vector<uint32_t> aligned_nodes;
};
struct EdgeData {  // POAEdgeData (poa.rs:57-80)
    uint64_t weight;
    std::vector<uint64_t> sequence_ids;
};
struct Pair {  // AlignedPair (aligner/alignment.rs:3-14); NONE = Option::None
    uint32_t rpos, qpos;
};
constexpr uint32_t NONE = 0xFFFFFFFFu;
enum class Err { Ok = 0, WeightsUnequalSize = 1, InvalidAlignment = 2, GraphError = 3 };

struct POAGraph {
    StableDiGraph<NodeData, EdgeData> graph;
    std::vector<std::pair<std::string, uint32_t>> sequences;  // Sequence(name, start_node)
    std::vector<uint32_t> topological_sorted;
    uint32_t start_node, end_node;

    POAGraph() {  // poa.rs:100-112
        start_node = graph.add_node(NodeData{'#', {}});
        end_node = graph.add_node(NodeData{'$', {}});
    }
    size_t node_count() const { return graph.node_count() - 2; }  // poa.rs:399-401
    bool is_empty() const { return node_count() == 0; }
    std::vector<uint32_t> successors(uint32_t n) const { return graph.neighbors_directed(n, 0); }
    std::vector<uint32_t> predecessors(uint32_t n) const { return graph.neighbors_directed(n, 1); }

    void add_edge(uint32_t s, uint32_t t, uint64_t sequence_id, uint64_t weight) {  // poa.rs:118-134
        if (auto e = graph.find_edge(s, t)) {
            EdgeData &d = *graph.edges[*e].weight;
            d.sequence_ids.push_back(sequence_id);
            d.weight += weight;
        } else {
            graph.add_edge(s, t, EdgeData{weight, {sequence_id}});
        }
    }

    // poa.rs:136-169
    std::optional<std::pair<uint32_t, uint32_t>> add_nodes_for_sequence(const std::string &sequence, const std::vector<uint64_t> &weights, size_t start,
                                                                      size_t end) {
        if (start == end) return std::nullopt;
        std::optional<uint32_t> first_node, prev;
        for (size_t pos = start; pos < end; ++pos) {
            const uint32_t curr_node = graph.add_node(NodeData{(uint8_t)sequence[pos], {}});
            if (!first_node) first_node = curr_node;
            if (prev) add_edge(*prev, curr_node, sequences.size(), weights[pos - 1] + weights[pos]);
            prev = curr_node;
        }
        return std::make_pair(*first_node, *prev);
    }

    // poa.rs:171-321; alignment == nullptr is None
    Err add_alignment_with_weights(const std::string &name, const std::string &sequence, const std::vector<Pair> *alignment,
                                   const std::vector<uint64_t> &weights) {
        if (sequence.size() != weights.size()) return Err::WeightsUnequalSize;
        if (!alignment) {
            if (sequence.empty()) {
                sequences.emplace_back(name, start_node);
            } else {
                auto nodes = add_nodes_for_sequence(sequence, weights, 0, sequence.size());
                sequences.emplace_back(name, nodes->first);
            }
            return post_process();
        }
        std::vector<size_t> valid_ix;
        for (const Pair &p : *alignment)
            if (p.qpos != NONE && p.qpos < sequence.size()) valid_ix.push_back(p.qpos);
        if (valid_ix.empty()) {
            if (sequence.empty()) {
                sequences.emplace_back(name, start_node);
                return post_process();
            }
            return Err::InvalidAlignment;
        }
        const size_t first = valid_ix.front(), last = valid_ix.back();
        auto nodes_unaligned_begin = add_nodes_for_sequence(sequence, weights, 0, first);
        std::optional<uint32_t> prev;
        if (nodes_unaligned_begin) prev = nodes_unaligned_begin->second;
        const auto nodes_unaligned_end = add_nodes_for_sequence(sequence, weights, last + 1, sequence.size());
        for (const Pair &p : *alignment) {
            if (p.qpos == NONE) continue;
            const size_t q = p.qpos;
            std::optional<uint32_t> curr;
            const uint8_t qsymbol = (uint8_t)sequence[q];
            if (p.rpos != NONE) {
                const uint32_t r = p.rpos;
                if (graph.nodes[r].weight.symbol == qsymbol) {
                    curr = r;
                } else {
                    for (uint32_t other_ix : graph.nodes[r].weight.aligned_nodes)
                        if (graph.nodes[other_ix].weight.symbol == qsymbol) {
                            curr = other_ix;
                            break;
                        }
                    if (!curr) {
                        const uint32_t new_node = graph.add_node(NodeData{qsymbol, {}});
                        curr = new_node;
                        const std::vector<uint32_t> other_nodes = graph.nodes[r].weight.aligned_nodes;  // .clone()
                        for (uint32_t other_ix : other_nodes) {
                            graph.nodes[other_ix].weight.aligned_nodes.push_back(new_node);
                            graph.nodes[new_node].weight.aligned_nodes.push_back(other_ix);
                        }
                        graph.nodes[r].weight.aligned_nodes.push_back(new_node);
                        graph.nodes[new_node].weight.aligned_nodes.push_back(r);
                    }
                }
            } else {
                curr = graph.add_node(NodeData{qsymbol, {}});
            }
            if (!nodes_unaligned_begin) nodes_unaligned_begin = std::make_pair(*curr, *curr);
            if (prev) add_edge(*prev, *curr, sequences.size(), weights[q - 1] + weights[q]);
            prev = curr;
        }
        if (nodes_unaligned_end) add_edge(*prev, nodes_unaligned_end->first, sequences.size(), weights[last] + weights[last + 1]);
        sequences.emplace_back(name, nodes_unaligned_begin->first);
        return post_process();
    }

    Err post_process() {  // poa.rs:323-363
        topological_sorted.clear();
        while (graph.nodes[start_node].next[0] != END) graph.remove_edge(graph.nodes[start_node].next[0]);
        while (graph.nodes[end_node].next[1] != END) graph.remove_edge(graph.nodes[end_node].next[1]);
        const uint32_t n = (uint32_t)graph.node_count();
        for (uint32_t node = 0; node < n; ++node)
            if (node != start_node && node != end_node && graph.nodes[node].next[1] == END) graph.add_edge(start_node, node, EdgeData{0, {}});
        for (uint32_t node = 0; node < n; ++node)
            if (node != end_node && node != start_node && graph.nodes[node].next[0] == END) graph.add_edge(node, end_node, EdgeData{0, {}});
        return toposort(graph, topological_sorted) ? Err::Ok : Err::GraphError;
    }
};

// ---- the AlignableRefGraph trait as the aligner's CSR view (graphs/poa.rs:384-471) ----------------------------------
struct Flat {
    std::vector<uint32_t> topo, pred_off, pred, succ_off, succ;
    std::vector<uint8_t> symbol;
    poasta_graph_view view;
    explicit Flat(const POAGraph &g) {
        const uint32_t n = (uint32_t)g.graph.node_count();
        topo = g.topological_sorted;
        pred_off.assign(1, 0), succ_off.assign(1, 0);
        for (uint32_t v = 0; v < n; ++v) {
            symbol.push_back(g.graph.nodes[v].weight.symbol);
            for (uint32_t p : g.predecessors(v)) pred.push_back(p);
            for (uint32_t s : g.successors(v)) succ.push_back(s);
            pred_off.push_back((uint32_t)pred.size()), succ_off.push_back((uint32_t)succ.size());
        }
        pred.push_back(0), succ.push_back(0);  // never empty arrays
        view.n_nodes = n, view.start_node = g.start_node, view.end_node = g.end_node;
        view.topo_order = topo.data(), view.pred_off = pred_off.data(), view.pred = pred.data();
        view.succ_off = succ_off.data(), view.succ = succ.data(), view.symbol = symbol.data();
    }
};

// ---- noodles-fasta reader / writer -----------------------------------------------------------------------------------
struct Record {
    std::string name, seq;
};
// reader: a definition line starts with '>', the name runs to the first whitespace; sequence lines are concatenated with line
// endings ('\n' or '\r\n') stripped
inline std::vector<Record> read_fasta(const char *data, size_t len) {
    std::vector<Record> out;
    size_t pos = 0;
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && data[eol] != '\n') ++eol;
        size_t stop = eol;
        if (stop > pos && data[stop - 1] == '\r') --stop;
        if (stop > pos) {
            if (data[pos] == '>') {
                size_t k = pos + 1;
                while (k < stop && data[k] != ' ' && data[k] != '\t') ++k;
                out.push_back(Record{std::string(data + pos + 1, data + k), std::string()});
            } else if (!out.empty()) {
                out.back().seq.append(data + pos, data + stop);
            }
        }
        pos = eol + 1;
    }
    return out;
}
// writer: ">name\n", then the sequence in lines of 80 bases (noodles' default line_base_count); an empty sequence writes no line
inline void write_fasta_record(std::string &out, const std::string &name, const std::string &seq) {
    out += '>';
    out += name;
    out += '\n';
    for (size_t i = 0; i < seq.size(); i += 80) {
        out.append(seq, i, 80);
        out += '\n';
    }
}

// ---- io/fasta.rs ---------------------------------------------------------------------------------------------------
// fasta_aln_for_seq (io/fasta.rs:19-67). Returns false where the reference's usize arithmetic would underflow
// (`node_col.saturating_sub(1) - last_col` with overflow checks off wraps to a huge repeat count and aborts on allocation).
inline bool fasta_aln_for_seq(const POAGraph &g, const std::unordered_map<uint32_t, size_t> &node_to_column, size_t seq_id, uint32_t start_node,
                              std::string &seq) {
    seq.clear();
    std::optional<uint32_t> curr = start_node;
    size_t last_col = 0;
    while (curr) {
        const uint32_t n = *curr;
        auto it = node_to_column.find(n);
        if (it == node_to_column.end()) {
            seq.clear();
            return true;  // "likely an empty sequence": empty row
        }
        const size_t node_col = it->second;
        const size_t sat = node_col == 0 ? 0 : node_col - 1;
        if (sat < last_col) return false;
        seq.append(sat - last_col, '-');
        seq += (char)g.graph.nodes[n].weight.symbol;
        curr.reset();
        for (uint32_t e : g.graph.edges_directed(n, 0)) {
            const auto &ids = g.graph.edges[e].weight->sequence_ids;
            if (std::binary_search(ids.begin(), ids.end(), (uint64_t)seq_id)) curr = g.graph.edges[e].node[1];
        }
        last_col = node_col;
    }
    if (!node_to_column.empty()) {
        size_t max_col = 0;
        for (auto &kv : node_to_column) max_col = std::max(max_col, kv.second);
        if (max_col < last_col) return false;
        seq.append(max_col - last_col, '-');
    }
    return true;
}

// poa_graph_to_fasta (io/fasta.rs:69-156)
inline bool graph_to_fasta(const POAGraph &g, std::string &out) {
    std::unordered_map<uint32_t, size_t> node_to_column;
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> stack;
    stack.emplace_back(g.start_node, g.successors(g.start_node));
    std::unordered_set<uint32_t> visited;
    std::vector<uint32_t> rev_postorder;
    while (!stack.empty()) {
        std::vector<uint32_t> &succ_iter = stack.back().second;
        std::optional<uint32_t> child;
        while (!succ_iter.empty()) {  // Vec::pop(): from the back
            const uint32_t s = succ_iter.back();
            succ_iter.pop_back();
            if (!visited.count(s)) {
                child = s;
                break;
            }
        }
        if (child) {
            visited.insert(*child);
            std::vector<uint32_t> successors = g.successors(*child);
            for (uint32_t aln_node : g.graph.nodes[*child].weight.aligned_nodes)
                if (!visited.count(aln_node)) {
                    visited.insert(aln_node);
                    for (uint32_t s : g.successors(aln_node)) successors.push_back(s);
                }
            stack.emplace_back(*child, std::move(successors));
        } else {
            rev_postorder.push_back(stack.back().first);
            stack.pop_back();
        }
    }
    std::reverse(rev_postorder.begin(), rev_postorder.end());
    size_t curr_col = 0;
    for (uint32_t n : rev_postorder) {
        if (n == g.start_node || n == g.end_node) continue;
        if (!node_to_column.count(n)) {
            node_to_column[n] = curr_col;
            for (uint32_t a : g.graph.nodes[n].weight.aligned_nodes) node_to_column[a] = curr_col;
            ++curr_col;
        }
    }
    std::string row;
    for (size_t seq_id = 0; seq_id < g.sequences.size(); ++seq_id) {
        if (!fasta_aln_for_seq(g, node_to_column, seq_id, g.sequences[seq_id].second, row)) return false;
        write_fasta_record(out, g.sequences[seq_id].first, row);
    }
    return true;
}

// ---- io/graph.rs:245-502 GFA / GFAv1 --------------------------------------------------------------------------------
// Returns false where the reference panics (HashMap index on a missing key, e.g. a walk that starts at `start`).
inline bool graph_to_gfa(const POAGraph &g, std::string &out, bool v1) {
    std::unordered_set<uint32_t> visited;
    std::deque<uint32_t> queue;
    queue.push_back(g.start_node);
    visited.insert(g.start_node);
    out += "H\tVN:Z:1.1\n";
    std::unordered_map<uint32_t, std::pair<size_t, size_t>> node_to_segment;
    std::unordered_map<uint32_t, size_t> segment_starts, segment_ends;
    std::unordered_map<size_t, size_t> segment_lengths;
    size_t curr_segment_id = 0;
    auto seg_name = [&](size_t id) { return v1 ? std::to_string(id + 1) : "s" + std::to_string(id); };
    while (!queue.empty()) {
        const uint32_t front = queue.front();
        queue.pop_front();
        if (front == g.start_node) {
            for (uint32_t succ : g.successors(front))
                if (!visited.count(succ)) {
                    queue.push_back(succ);
                    visited.insert(succ);
                }
            continue;
        }
        std::string segment(1, (char)g.graph.nodes[front].weight.symbol);
        uint32_t curr_node = front;
        size_t curr_out_degree = g.successors(front).size();
        size_t seg_pos = 0;
        node_to_segment[front] = {curr_segment_id, seg_pos};
        segment_starts[front] = curr_segment_id;
        while (curr_out_degree == 1) {
            const uint32_t next_node = g.successors(curr_node)[0];
            const size_t in_degree_next = g.predecessors(next_node).size();
            if (in_degree_next == 1 && next_node != g.end_node) {
                segment += (char)g.graph.nodes[next_node].weight.symbol;
                node_to_segment[next_node] = {curr_segment_id, seg_pos};
            } else {
                break;
            }
            curr_node = next_node;
            curr_out_degree = g.successors(curr_node).size();
            seg_pos += 1;
        }
        out += "S\t" + seg_name(curr_segment_id) + "\t" + segment + "\n";
        segment_ends[curr_node] = curr_segment_id;
        segment_lengths[curr_segment_id] = segment.size();
        visited.insert(curr_node);
        for (uint32_t succ : g.successors(curr_node))
            if (!visited.count(succ) && succ != g.end_node) {
                visited.insert(succ);
                queue.push_back(succ);
            }
        ++curr_segment_id;
    }
    for (size_t e = 0; e < g.graph.edges.size(); ++e) {  // edge_references(): slot order, vacant slots skipped
        const auto &ed = g.graph.edges[e];
        if (!ed.weight) continue;
        auto a = segment_ends.find(ed.node[0]);
        auto b = segment_starts.find(ed.node[1]);
        if (a != segment_ends.end() && b != segment_starts.end()) out += "L\t" + seg_name(a->second) + "\t+\t" + seg_name(b->second) + "\t+\t0M\n";
    }
    for (size_t seq_id = 0; seq_id < g.sequences.size(); ++seq_id) {
        const uint32_t first = g.sequences[seq_id].second;
        auto it0 = node_to_segment.find(first);
        if (it0 == node_to_segment.end()) return false;
        size_t prev_segment = it0->second.first;
        const size_t start_pos = it0->second.second;
        std::vector<size_t> walk_segments{prev_segment};
        size_t last_pos = 0;
        size_t total_segments_length = segment_lengths[prev_segment];
        std::optional<uint32_t> curr = first;
        while (curr) {
            const uint32_t n = *curr;
            auto it = node_to_segment.find(n);
            if (it == node_to_segment.end()) return false;
            const size_t node_segment = it->second.first;
            last_pos = it->second.second;
            if (node_segment != prev_segment) {
                walk_segments.push_back(node_segment);
                total_segments_length += segment_lengths[node_segment];
            }
            curr.reset();
            for (uint32_t e : g.graph.edges_directed(n, 0)) {
                const auto &ids = g.graph.edges[e].weight->sequence_ids;
                if (std::binary_search(ids.begin(), ids.end(), (uint64_t)seq_id)) curr = g.graph.edges[e].node[1];
            }
            prev_segment = node_segment;
        }
        if (!v1) {
            const size_t end_pos = total_segments_length - segment_lengths[prev_segment] + last_pos;
            out += "W\t*\t0\t" + g.sequences[seq_id].first + "\t" + std::to_string(start_pos) + "\t" + std::to_string(end_pos) + "\t";
            for (size_t s : walk_segments) out += ">s" + std::to_string(s);
            out += "\n";
        } else {
            out += "P\t" + g.sequences[seq_id].first + "\t";
            for (size_t k = 0; k < walk_segments.size(); ++k) out += (k ? "," : "") + std::to_string(walk_segments[k] + 1) + "+";
            out += "\t*\n";
        }
    }
    return true;
}

// ---- .poasta: bincode 1.3 (little endian, fixed-width ints, u64 lengths, u32 enum tags) of POAGraphWithIx::U32 -------------
// POAGraph { graph: StableGraph (petgraph serde form: nodes, node_holes, edge_property, edges: Vec<Option<(u32, u32, E)>>),
//            sequences: Vec<(String, u32)>, topological_sorted: Vec<u32>, start_node: u32, end_node: u32 }   (poa.rs:84-94, 484-489)
struct Bin {
    std::string &out;
    void u8(uint8_t v) { out.push_back((char)v); }
    void u32(uint32_t v) { out.append((const char *)&v, 4); }
    void u64(uint64_t v) { out.append((const char *)&v, 8); }
};
inline void save_graph(const POAGraph &g, std::string &out) {
    Bin b{out};
    b.u32(2);  // enum POAGraphWithIx { U8, U16, U32, USIZE }: the CLI builds U32 (bin/poasta.rs:316)
    b.u64(g.graph.nodes.size());
    for (const auto &n : g.graph.nodes) {
        b.u8(n.weight.symbol);
        b.u64(n.weight.aligned_nodes.size());
        for (uint32_t a : n.weight.aligned_nodes) b.u32(a);
    }
    b.u64(0);  // node_holes: none
    b.u32(1);  // EdgeProperty::Directed
    b.u64(g.graph.edges.size());
    for (const auto &e : g.graph.edges) {
        if (!e.weight) {
            b.u8(0);  // None: vacant slot
            continue;
        }
        b.u8(1);
        b.u32(e.node[0]), b.u32(e.node[1]);
        b.u64(e.weight->weight);
        b.u64(e.weight->sequence_ids.size());
        for (uint64_t s : e.weight->sequence_ids) b.u64(s);
    }
    b.u64(g.sequences.size());
    for (const auto &s : g.sequences) {
        b.u64(s.first.size());
        out += s.first;
        b.u32(s.second);
    }
    b.u64(g.topological_sorted.size());
    for (uint32_t v : g.topological_sorted) b.u32(v);
    b.u32(g.start_node), b.u32(g.end_node);
}
// load: petgraph's deserializer re-links the Some edges in slot order (so a list's head is its highest slot) and chains the
// None slots on the free list in slot order (the LAST vacant slot becomes the head)
inline bool load_graph(const uint8_t *p, size_t len, POAGraph &g) {
    const uint8_t *end = p + len;
    bool ok = true;
    auto take = [&](void *dst, size_t n) {
        if ((size_t)(end - p) < n) {
            ok = false;
            memset(dst, 0, n);
            return;
        }
        memcpy(dst, p, n);
        p += n;
    };
    auto r8 = [&]() { uint8_t v; take(&v, 1); return v; };
    auto r32 = [&]() { uint32_t v; take(&v, 4); return v; };
    auto r64 = [&]() { uint64_t v; take(&v, 8); return v; };
    if (r32() != 2) return false;
    g = POAGraph();
    g.graph.nodes.clear();
    const uint64_t nn = r64();
    if (!ok || nn > (uint64_t)(end - p)) return false;
    for (uint64_t i = 0; i < nn && ok; ++i) {
        NodeData d{r8(), {}};
        const uint64_t na = r64();
        if (!ok || na > (uint64_t)(end - p) / 4) return false;
        for (uint64_t k = 0; k < na; ++k) d.aligned_nodes.push_back(r32());
        g.graph.nodes.push_back({std::move(d), {END, END}});
    }
    if (r64() != 0) return false;
    r32();
    const uint64_t ne = r64();
    if (!ok || ne > (uint64_t)(end - p)) return false;
    for (uint64_t i = 0; i < ne && ok; ++i) {
        g.graph.edges.emplace_back();
        auto &e = g.graph.edges.back();
        if (r8() == 0) {
            e.next[0] = g.graph.free_edge, e.next[1] = END;
            e.node[0] = e.node[1] = END;
            g.graph.free_edge = (uint32_t)i;
            continue;
        }
        e.node[0] = r32(), e.node[1] = r32();
        EdgeData d{r64(), {}};
        const uint64_t ns = r64();
        if (!ok || e.node[0] >= nn || e.node[1] >= nn || ns > (uint64_t)(end - p) / 8) return false;
        for (uint64_t k = 0; k < ns; ++k) d.sequence_ids.push_back(r64());
        e.weight = std::move(d);
        e.next[0] = g.graph.nodes[e.node[0]].next[0], e.next[1] = g.graph.nodes[e.node[1]].next[1];
        g.graph.nodes[e.node[0]].next[0] = (uint32_t)i, g.graph.nodes[e.node[1]].next[1] = (uint32_t)i;
        ++g.graph.edge_count_;
    }
    const uint64_t nsq = r64();
    for (uint64_t i = 0; i < nsq && ok; ++i) {
        const uint64_t nl = r64();
        if (!ok || nl > (uint64_t)(end - p)) return false;
        std::string name((const char *)p, nl);
        p += nl;
        g.sequences.emplace_back(std::move(name), r32());
    }
    const uint64_t nt = r64();
    if (!ok || nt > (uint64_t)(end - p) / 4) return false;
    for (uint64_t i = 0; i < nt; ++i) g.topological_sorted.push_back(r32());
    g.start_node = r32(), g.end_node = r32();
    return ok;
}

// load_graph_from_fasta_msa (io/graph.rs:36-103): one node per (column, symbol), '-' skipped
inline bool graph_from_fasta_msa(const char *data, size_t len, POAGraph &g) {
    g = POAGraph();
    std::vector<std::vector<uint32_t>> nodes_per_col;
    const auto recs = read_fasta(data, len);
    for (size_t seq_id = 0; seq_id < recs.size(); ++seq_id) {
        const std::string &chars = recs[seq_id].seq;
        if (chars.size() > nodes_per_col.size()) nodes_per_col.resize(chars.size());
        std::optional<uint32_t> prev_node;
        for (size_t col = 0; col < chars.size(); ++col) {
            const uint8_t c = (uint8_t)chars[col];
            if (c == '-') continue;
            std::optional<uint32_t> node_ix;
            for (uint32_t v : nodes_per_col[col])
                if (g.graph.nodes[v].weight.symbol == c) {
                    node_ix = v;
                    break;
                }
            if (!node_ix) {
                const uint32_t new_node = g.graph.add_node(NodeData{c, {}});
                for (uint32_t other : nodes_per_col[col]) {
                    g.graph.nodes[other].weight.aligned_nodes.push_back(new_node);
                    g.graph.nodes[new_node].weight.aligned_nodes.push_back(other);
                }
                nodes_per_col[col].push_back(new_node);
                node_ix = new_node;
            }
            if (prev_node) g.add_edge(*prev_node, *node_ix, seq_id, 2);
            else g.sequences.emplace_back(recs[seq_id].name, *node_ix);
            prev_node = node_ix;
        }
    }
    return g.post_process() == Err::Ok;
}

}  // namespace poa
}  // namespace oracle
