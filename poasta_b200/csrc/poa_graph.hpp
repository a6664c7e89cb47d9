// Host-side partial-order graph for the B200 aligner.
//
// Behavioural contract (not code) follows the reference's POAGraph:
//   /root/reference/src/graphs/poa.rs:96-382  (new / add_alignment_with_weights / post_process)
//   /root/reference/src/graphs/poa.rs:384-471 (AlignableRefGraph view: adjacency order, symbols)
// The reference stores the graph in petgraph 0.6.5 `StableDiGraph` (un-vendored crate). What
// leaks from that crate into results is restated here from its published behaviour
// (SURVEY.md A.1-A.3):
//   * every node heads two singly linked edge lists (outgoing / incoming); a new edge is linked
//     at the HEAD of both lists, so neighbours enumerate newest-edge-first;
//   * removing an edge unlinks it (order of the rest unchanged) and pushes its slot on a LIFO
//     free list which `add_edge` pops before growing the edge array;
//   * `toposort` is the DFS / finish-stack variant started from the highest node index.
// The layout is index-based (struct-of-arrays friendly) because the graph is flattened to CSR
// for the device after every insertion (flatten.hpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

namespace poasta {

constexpr uint32_t NIL = 0xFFFFFFFFu;

enum class GraphError { Ok = 0, WeightsUnequalSize = 1, InvalidAlignment = 2, Cycle = 3 };

// One entry of an alignment: rpos/qpos == NIL means None (reference: aligner/alignment.rs:3-14).
struct AlignedPair {
    uint32_t rpos;
    uint32_t qpos;
};

// A small list whose storage is either its own heap block or a slice of a pool owned by the graph. The bulk loader
// (dev_graph_host.hpp: download of a finished build) points every list of a graph into two pools: a 20 x 1 kb graph holds
// ~3,000 of these lists, and one allocation per list made both the conversion and the destruction of a batch's 6,000 graphs
// allocation-bound (0.4 ms per graph). A list that grows past its slice, and every copy, owns its storage.
template <typename T>
class PoolList {
    T *p_ = nullptr;
    uint32_t n_ = 0, cap_ = 0;
    bool owned_ = false;

    void release() {
        if (owned_) free(p_);
        p_ = nullptr;
        n_ = cap_ = 0;
        owned_ = false;
    }
    void assign(const T *first, size_t n) {  // [first, first + n) never lies inside this list's own storage
        release();
        if (n) {
            p_ = static_cast<T *>(malloc(n * sizeof(T)));
            memcpy(p_, first, n * sizeof(T));
            n_ = cap_ = (uint32_t)n;
            owned_ = true;
        }
    }

  public:
    PoolList() = default;
    PoolList(std::initializer_list<T> il) { assign(il.begin(), il.size()); }
    PoolList(const std::vector<T> &v) { assign(v.data(), v.size()); }
    PoolList(const PoolList &o) { assign(o.p_, o.n_); }
    PoolList(PoolList &&o) noexcept : p_(o.p_), n_(o.n_), cap_(o.cap_), owned_(o.owned_) {
        o.p_ = nullptr;
        o.n_ = o.cap_ = 0;
        o.owned_ = false;
    }
    PoolList &operator=(const PoolList &o) {
        if (this != &o) assign(o.p_, o.n_);
        return *this;
    }
    PoolList &operator=(PoolList &&o) noexcept {
        if (this != &o) {
            release();
            p_ = o.p_, n_ = o.n_, cap_ = o.cap_, owned_ = o.owned_;
            o.p_ = nullptr;
            o.n_ = o.cap_ = 0;
            o.owned_ = false;
        }
        return *this;
    }
    PoolList &operator=(const std::vector<T> &v) {
        assign(v.data(), v.size());
        return *this;
    }
    ~PoolList() { release(); }

    // empty list over `cap` elements of a pool the owner of this list keeps alive (and does not reallocate) while the list lives
    void set_pool_slice(T *p, uint32_t cap) {
        release();
        p_ = p, cap_ = cap;
    }
    // the same with the slice already filled
    void set_pool_filled(T *p, uint32_t n) {
        release();
        p_ = p, n_ = cap_ = n;
    }
    void clear() { n_ = 0; }
    void push_back(const T &v) {
        if (n_ == cap_) {
            const uint32_t nc = cap_ < 4 ? 4 : cap_ * 2;
            T *np = static_cast<T *>(malloc((size_t)nc * sizeof(T)));
            if (n_) memcpy(np, p_, (size_t)n_ * sizeof(T));
            if (owned_) free(p_);
            p_ = np, cap_ = nc, owned_ = true;
        }
        p_[n_++] = v;
    }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
    const T &operator[](size_t i) const { return p_[i]; }
    bool operator==(const std::vector<T> &v) const { return v.size() == n_ && (n_ == 0 || memcmp(v.data(), p_, (size_t)n_ * sizeof(T)) == 0); }
    bool operator!=(const std::vector<T> &v) const { return !(*this == v); }
};
template <typename T> bool operator==(const std::vector<T> &v, const PoolList<T> &l) { return l == v; }
template <typename T> bool operator!=(const std::vector<T> &v, const PoolList<T> &l) { return !(l == v); }

struct POAGraph {
    struct Node {
        uint8_t symbol;
        uint32_t head[2];  // head[0] = first outgoing edge, head[1] = first incoming edge
        PoolList<uint32_t> aligned_nodes;
    };
    struct Edge {
        uint32_t node[2];  // source, target (NIL,NIL when the slot is vacant)
        uint32_t next[2];  // next edge in source's out list / target's in list
        bool vacant;
        uint64_t weight;
        PoolList<uint64_t> sequence_ids;
    };
    struct Sequence {
        std::string name;
        uint32_t start_node;
    };

    // pools behind the lists of a bulk-loaded graph; declared before the lists so that they outlive them
    std::vector<uint32_t> al_pool;
    std::vector<uint64_t> id_pool;
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    uint32_t free_edge = NIL;
    uint32_t live_edges = 0;
    std::vector<Sequence> sequences;
    std::vector<uint32_t> topological_sorted;
    uint32_t start_node = 0, end_node = 1;

    POAGraph() {  // poa.rs:100-112
        add_node('#');
        add_node('$');
    }

    // ---- petgraph-like primitives -------------------------------------------------------
    uint32_t add_node(uint8_t symbol) {
        nodes.push_back(Node{symbol, {NIL, NIL}, {}});
        return (uint32_t)nodes.size() - 1;
    }

    uint32_t raw_add_edge(uint32_t a, uint32_t b, uint64_t weight, std::vector<uint64_t> ids) {
        uint32_t e;
        if (free_edge != NIL) {
            e = free_edge;
            free_edge = edges[e].next[0];
        } else {
            e = (uint32_t)edges.size();
            edges.emplace_back();
        }
        Edge &ed = edges[e];
        ed.node[0] = a;
        ed.node[1] = b;
        ed.next[0] = nodes[a].head[0];
        ed.next[1] = nodes[b].head[1];
        ed.vacant = false;
        ed.weight = weight;
        ed.sequence_ids = std::move(ids);
        nodes[a].head[0] = e;
        nodes[b].head[1] = e;
        ++live_edges;
        return e;
    }

    void raw_remove_edge(uint32_t e) {
        Edge &ed = edges[e];
        for (int d = 0; d < 2; ++d) {
            uint32_t n = ed.node[d];
            uint32_t *link = &nodes[n].head[d];
            while (*link != e) link = &edges[*link].next[d];
            *link = ed.next[d];
        }
        ed.next[0] = free_edge;
        ed.next[1] = NIL;
        ed.node[0] = ed.node[1] = NIL;
        ed.vacant = true;
        ed.weight = 0;
        ed.sequence_ids.clear();
        free_edge = e;
        --live_edges;
    }

    uint32_t find_edge(uint32_t a, uint32_t b) const {
        for (uint32_t e = nodes[a].head[0]; e != NIL; e = edges[e].next[0])
            if (edges[e].node[1] == b) return e;
        return NIL;
    }

    template <class F> void for_successors(uint32_t n, F &&f) const {
        for (uint32_t e = nodes[n].head[0]; e != NIL; e = edges[e].next[0]) f(edges[e].node[1]);
    }
    template <class F> void for_predecessors(uint32_t n, F &&f) const {
        for (uint32_t e = nodes[n].head[1]; e != NIL; e = edges[e].next[1]) f(edges[e].node[0]);
    }
    uint32_t out_degree(uint32_t n) const {
        uint32_t c = 0;
        for (uint32_t e = nodes[n].head[0]; e != NIL; e = edges[e].next[0]) ++c;
        return c;
    }
    uint32_t in_degree(uint32_t n) const {
        uint32_t c = 0;
        for (uint32_t e = nodes[n].head[1]; e != NIL; e = edges[e].next[1]) ++c;
        return c;
    }

    // ---- reference-visible API -----------------------------------------------------------
    uint32_t node_count() const { return (uint32_t)nodes.size() - 2; }  // poa.rs:399-401
    uint32_t node_count_with_start_and_end() const { return (uint32_t)nodes.size(); }
    bool is_empty() const { return node_count() == 0; }
    uint32_t edge_count() const {  // poa.rs:409-420
        return live_edges - out_degree(start_node) - in_degree(end_node);
    }
    uint8_t get_symbol(uint32_t n) const { return nodes[n].symbol; }
    bool is_symbol_equal(uint32_t n, uint8_t s) const {  // poa.rs:463-465
        return n == end_node || nodes[n].symbol == s;
    }

    void add_edge(uint32_t s, uint32_t t, uint64_t sequence_id, uint64_t weight) {  // poa.rs:118-134
        uint32_t e = find_edge(s, t);
        if (e != NIL) {
            edges[e].sequence_ids.push_back(sequence_id);
            edges[e].weight += weight;
        } else {
            raw_add_edge(s, t, weight, {sequence_id});
        }
    }

    // poa.rs:136-169. Returns false when start == end (None).
    bool add_nodes_for_sequence(const uint8_t *seq, const uint64_t *weights, size_t start, size_t end,
                                uint32_t &first, uint32_t &last) {
        if (start == end) return false;
        uint32_t prev = NIL;
        first = NIL;
        for (size_t pos = start; pos < end; ++pos) {
            uint32_t curr = add_node(seq[pos]);
            if (first == NIL) first = curr;
            if (prev != NIL)
                add_edge(prev, curr, sequences.size(), (weights ? weights[pos - 1] + weights[pos] : 2));
            prev = curr;
        }
        last = prev;
        return true;
    }

    // poa.rs:171-321. `has_alignment == false` is the reference's `alignment_opt == None`; an EMPTY alignment is
    // Some(&[]) and is rejected as InvalidAlignment for a non-empty sequence (poa.rs:210-224).
    // `weights == nullptr` means all ones (what the CLI passes, bin/poasta.rs:193).
    GraphError add_alignment_with_weights(const std::string &name, const uint8_t *seq, size_t len, bool has_alignment,
                                          const AlignedPair *aln, size_t n_pairs,
                                          const uint64_t *weights = nullptr) {
        auto w = [&](size_t i) -> uint64_t { return weights ? weights[i] : 1; };
        if (!has_alignment) {
            if (len == 0) {
                sequences.push_back({name, start_node});
                return post_process();
            }
            uint32_t first, last;
            add_nodes_for_sequence(seq, weights, 0, len, first, last);
            sequences.push_back({name, first});
            return post_process();
        }
        size_t first_q = SIZE_MAX, last_q = 0;
        bool any = false;
        for (size_t i = 0; i < n_pairs; ++i) {
            if (aln[i].qpos != NIL && aln[i].qpos < len) {
                if (!any) first_q = aln[i].qpos;
                last_q = aln[i].qpos;
                any = true;
            }
        }
        if (!any) {
            if (len == 0) {
                sequences.push_back({name, start_node});
                return post_process();
            }
            return GraphError::InvalidAlignment;
        }
        uint32_t begin_first = NIL, begin_last = NIL, end_first = NIL, end_last = NIL;
        bool has_begin = add_nodes_for_sequence(seq, weights, 0, first_q, begin_first, begin_last);
        uint32_t prev = has_begin ? begin_last : NIL;
        bool has_end = add_nodes_for_sequence(seq, weights, last_q + 1, len, end_first, end_last);

        for (size_t i = 0; i < n_pairs; ++i) {
            if (aln[i].qpos == NIL) continue;
            uint32_t q = aln[i].qpos, r = aln[i].rpos;
            uint8_t qsymbol = seq[q];
            uint32_t curr = NIL;
            if (r != NIL) {
                if (nodes[r].symbol == qsymbol) {
                    curr = r;
                } else {
                    for (uint32_t other : nodes[r].aligned_nodes)
                        if (nodes[other].symbol == qsymbol) {
                            curr = other;
                            break;
                        }
                    if (curr == NIL) {
                        uint32_t nn = add_node(qsymbol);
                        curr = nn;
                        const std::vector<uint32_t> others(nodes[r].aligned_nodes.begin(), nodes[r].aligned_nodes.end());
                        for (uint32_t other : others) {
                            nodes[other].aligned_nodes.push_back(nn);
                            nodes[nn].aligned_nodes.push_back(other);
                        }
                        nodes[r].aligned_nodes.push_back(nn);
                        nodes[nn].aligned_nodes.push_back(r);
                    }
                }
            } else {
                curr = add_node(qsymbol);
            }
            if (!has_begin) {
                begin_first = curr;
                has_begin = true;
            }
            if (prev != NIL) add_edge(prev, curr, sequences.size(), w(q - 1) + w(q));
            prev = curr;
        }
        if (has_end) add_edge(prev, end_first, sequences.size(), w(last_q) + w(last_q + 1));
        sequences.push_back({name, begin_first});
        return post_process();
    }

    // poa.rs:323-363
    GraphError post_process() {
        topological_sorted.clear();
        while (nodes[start_node].head[0] != NIL) raw_remove_edge(nodes[start_node].head[0]);
        while (nodes[end_node].head[1] != NIL) raw_remove_edge(nodes[end_node].head[1]);
        const uint32_t n = (uint32_t)nodes.size();
        for (uint32_t v = 0; v < n; ++v)
            if (v != start_node && v != end_node && nodes[v].head[1] == NIL) raw_add_edge(start_node, v, 0, {});
        for (uint32_t v = 0; v < n; ++v)
            if (v != start_node && v != end_node && nodes[v].head[0] == NIL) raw_add_edge(v, end_node, 0, {});
        return toposort() ? GraphError::Ok : GraphError::Cycle;
    }

    // petgraph 0.6.5 algo::toposort restated.
    bool toposort() {
        const uint32_t n = (uint32_t)nodes.size();
        std::vector<uint8_t> discovered(n, 0), finished(n, 0);
        std::vector<uint32_t> stack, finish_stack;
        finish_stack.reserve(n);
        for (uint32_t i = n; i-- > 0;) {
            if (discovered[i]) continue;
            stack.push_back(i);
            while (!stack.empty()) {
                uint32_t nx = stack.back();
                if (!discovered[nx]) {
                    discovered[nx] = 1;
                    for (uint32_t e = nodes[nx].head[0]; e != NIL; e = edges[e].next[0]) {
                        uint32_t succ = edges[e].node[1];
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
        // cycle check: every edge must go forward in the order
        std::vector<uint32_t> rank(n);
        for (uint32_t r = 0; r < n; ++r) rank[finish_stack[r]] = r;
        for (const Edge &e : edges)
            if (!e.vacant && rank[e.node[0]] >= rank[e.node[1]]) return false;
        topological_sorted = std::move(finish_stack);
        return true;
    }

    std::vector<uint32_t> get_node_ranks() const {  // poa.rs:365-372
        std::vector<uint32_t> ranks(topological_sorted.size(), 0);
        for (uint32_t r = 0; r < topological_sorted.size(); ++r) ranks[topological_sorted[r]] = r;
        return ranks;
    }
};

}  // namespace poasta
