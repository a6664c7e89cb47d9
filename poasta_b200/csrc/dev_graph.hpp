// Device-resident POA graph: the graph of one family lives in HBM for the whole progressive build and is
// updated ON THE DEVICE after every alignment, so a round needs no host round trip.
//
// Behavioural contract = the reference's POAGraph over petgraph 0.6.5 `StableDiGraph`:
//   /root/reference/src/graphs/poa.rs:100-112  new (start '#' = node 0, end '$' = node 1)
//   /root/reference/src/graphs/poa.rs:118-134  add_edge (reuse an existing edge: push the sequence id, add the weight)
//   /root/reference/src/graphs/poa.rs:136-169  add_nodes_for_sequence
//   /root/reference/src/graphs/poa.rs:171-321  add_alignment_with_weights (all weights 1: bin/poasta.rs:193)
//   /root/reference/src/graphs/poa.rs:323-363  post_process (start / end rewiring; the toposort is redone by the host
//                                              when the graph is downloaded - ranks never influence an alignment)
// petgraph behaviour that leaks into results: an edge is linked at the HEAD of its source's out-list and
// its target's in-list (neighbours enumerate newest-edge-first); a removed edge's slot goes on a LIFO free list that
// add_edge pops before growing the edge array.
//
// Layout: struct-of-arrays in one arena per family, sized on the host from the family's total sequence length S
// (worst case: every base becomes a node). Everything is plain index arithmetic over raw pointers so the same source
// runs on the device (one lane of the family's worker warp; the update is ~0.1 % of a round) and on the host, where
// tests/test_dev_graph.py checks it field by field against the std::vector implementation (poa_graph.hpp).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define DG_HD __host__ __device__ __forceinline__
#else
#define DG_HD inline
#endif

namespace poasta {
namespace dg {

constexpr uint32_t NIL = 0xFFFFFFFFu;
enum : int { DG_OK = 0, DG_INVALID_ALIGNMENT = 2, DG_CAPACITY = 8 };  // 2 = GraphError::InvalidAlignment (poa_graph.hpp)

// Capacities of one family's arena (elements). S = sum of the family's sequence lengths, n = number of sequences,
// V0 / E0 / ... = what an uploaded existing graph already holds.
struct Caps {
    uint32_t nodes, edges, sids, als, seqs;
};

// Mutable header of one family (device memory; one cache line).
struct alignas(64) Family {
    uint64_t arena_off;   // byte offset of this family's arena inside the chunk arena
    Caps cap;
    uint32_t n_nodes, n_edge_slots, free_edge, live_edges, n_sids, n_als, n_seqs_in_graph;
    uint32_t seq_first, n_seqs;  // this family's records in the chunk's sequence table [seq_first, seq_first + n_seqs)
    uint32_t next_seq;           // next record to align / insert
    int32_t status;              // first error (poasta_status) - the family stops there
    uint32_t pad;
};

// Byte layout of a family arena for given caps (all sub-arrays 16-byte aligned).
struct Layout {
    uint64_t sym, head_out, head_in, al_head, e_src, e_dst, e_next_out, e_next_in, e_weight, e_sid_head, e_sid_tail, sid_val, sid_next, al_node,
        al_next, seq_start, total;
    DG_HD static uint64_t up(uint64_t b) { return (b + 15) & ~(uint64_t)15; }
    DG_HD explicit Layout(const Caps &c) {
        uint64_t o = 0;
        auto take = [&](uint64_t bytes) {
            const uint64_t at = o;
            o += up(bytes);
            return at;
        };
        sym = take(c.nodes);
        head_out = take((uint64_t)c.nodes * 4), head_in = take((uint64_t)c.nodes * 4), al_head = take((uint64_t)c.nodes * 4);
        e_src = take((uint64_t)c.edges * 4), e_dst = take((uint64_t)c.edges * 4);
        e_next_out = take((uint64_t)c.edges * 4), e_next_in = take((uint64_t)c.edges * 4);
        e_weight = take((uint64_t)c.edges * 4), e_sid_head = take((uint64_t)c.edges * 4), e_sid_tail = take((uint64_t)c.edges * 4);
        sid_val = take((uint64_t)c.sids * 4), sid_next = take((uint64_t)c.sids * 4);
        al_node = take((uint64_t)c.als * 4), al_next = take((uint64_t)c.als * 4);
        seq_start = take((uint64_t)c.seqs * 4);
        total = o;
    }
};

// Pointer view of one family's graph.
struct Store {
    Family *f;
    uint8_t *sym;
    uint32_t *head_out, *head_in;  // first outgoing / incoming edge of a node (NIL = none)
    uint32_t *al_head;             // aligned_nodes list of a node: chain through al_node / al_next, in push order
    uint32_t *e_src, *e_dst, *e_next_out, *e_next_in, *e_weight;  // e_src == NIL: vacant slot (e_next_out = next free slot)
    uint32_t *e_sid_head, *e_sid_tail;                             // sequence_ids of an edge, in push order
    uint32_t *sid_val, *sid_next;
    uint32_t *al_node, *al_next;
    uint32_t *seq_start;  // first node of every inserted sequence (POAGraph::sequences)

    DG_HD Store(Family *fam, uint8_t *chunk_arena) : f(fam) {
        const Layout L(fam->cap);
        uint8_t *b = chunk_arena + fam->arena_off;
        sym = b + L.sym;
        head_out = (uint32_t *)(b + L.head_out), head_in = (uint32_t *)(b + L.head_in), al_head = (uint32_t *)(b + L.al_head);
        e_src = (uint32_t *)(b + L.e_src), e_dst = (uint32_t *)(b + L.e_dst);
        e_next_out = (uint32_t *)(b + L.e_next_out), e_next_in = (uint32_t *)(b + L.e_next_in);
        e_weight = (uint32_t *)(b + L.e_weight), e_sid_head = (uint32_t *)(b + L.e_sid_head), e_sid_tail = (uint32_t *)(b + L.e_sid_tail);
        sid_val = (uint32_t *)(b + L.sid_val), sid_next = (uint32_t *)(b + L.sid_next);
        al_node = (uint32_t *)(b + L.al_node), al_next = (uint32_t *)(b + L.al_next);
        seq_start = (uint32_t *)(b + L.seq_start);
    }

    static constexpr uint32_t START = 0, END = 1;  // poa.rs:102-103

    DG_HD void init_empty() {  // POAGraph::new
        f->n_nodes = f->n_edge_slots = f->live_edges = f->n_sids = f->n_als = f->n_seqs_in_graph = 0;
        f->free_edge = NIL;
        add_node('#');
        add_node('$');
    }
    DG_HD bool is_empty() const { return f->n_nodes <= 2; }

    // ---- petgraph-like primitives ------------------------------------------------------------------------
    DG_HD uint32_t add_node(uint8_t s) {
        if (f->n_nodes >= f->cap.nodes) return NIL;
        const uint32_t n = f->n_nodes++;
        sym[n] = s;
        head_out[n] = head_in[n] = al_head[n] = NIL;
        return n;
    }
    // new edge a -> b carrying `sid` (NIL: no sequence id, the start / end edges of post_process)
    DG_HD int raw_add_edge(uint32_t a, uint32_t b, uint32_t weight, uint32_t sid) {
        uint32_t e;
        if (f->free_edge != NIL) {
            e = f->free_edge;
            f->free_edge = e_next_out[e];
        } else {
            if (f->n_edge_slots >= f->cap.edges) return DG_CAPACITY;
            e = f->n_edge_slots++;
        }
        e_src[e] = a, e_dst[e] = b;
        e_next_out[e] = head_out[a], e_next_in[e] = head_in[b];
        e_weight[e] = weight;
        e_sid_head[e] = e_sid_tail[e] = NIL;
        head_out[a] = e, head_in[b] = e;
        ++f->live_edges;
        return sid == NIL ? DG_OK : push_sid(e, sid);
    }
    DG_HD int push_sid(uint32_t e, uint32_t sid) {
        if (f->n_sids >= f->cap.sids) return DG_CAPACITY;
        const uint32_t k = f->n_sids++;
        sid_val[k] = sid, sid_next[k] = NIL;
        if (e_sid_tail[e] == NIL) e_sid_head[e] = k;
        else sid_next[e_sid_tail[e]] = k;
        e_sid_tail[e] = k;
        return DG_OK;
    }
    DG_HD void raw_remove_edge(uint32_t e) {
        {
            uint32_t *link = &head_out[e_src[e]];
            while (*link != e) link = &e_next_out[*link];
            *link = e_next_out[e];
        }
        {
            uint32_t *link = &head_in[e_dst[e]];
            while (*link != e) link = &e_next_in[*link];
            *link = e_next_in[e];
        }
        e_next_out[e] = f->free_edge;
        e_next_in[e] = NIL;
        e_src[e] = e_dst[e] = NIL;
        e_weight[e] = 0;
        e_sid_head[e] = e_sid_tail[e] = NIL;  // start / end edges carry no ids, so no pool entries are lost
        f->free_edge = e;
        --f->live_edges;
    }
    DG_HD uint32_t find_edge(uint32_t a, uint32_t b) const {
        for (uint32_t e = head_out[a]; e != NIL; e = e_next_out[e])
            if (e_dst[e] == b) return e;
        return NIL;
    }
    DG_HD int add_edge(uint32_t s, uint32_t t, uint32_t sid, uint32_t weight) {  // poa.rs:118-134
        const uint32_t e = find_edge(s, t);
        if (e != NIL) {
            e_weight[e] += weight;
            return push_sid(e, sid);
        }
        return raw_add_edge(s, t, weight, sid);
    }
    DG_HD int push_aligned(uint32_t node, uint32_t other) {  // nodes[node].aligned_nodes.push(other)
        if (f->n_als >= f->cap.als) return DG_CAPACITY;
        const uint32_t k = f->n_als++;
        al_node[k] = other, al_next[k] = NIL;
        if (al_head[node] == NIL) {
            al_head[node] = k;
        } else {
            uint32_t t = al_head[node];
            while (al_next[t] != NIL) t = al_next[t];
            al_next[t] = k;
        }
        return DG_OK;
    }

    // poa.rs:136-169; returns DG_OK and first/last (NIL when start == end)
    DG_HD int add_nodes_for_sequence(const uint8_t *seq, uint32_t start, uint32_t end, uint32_t sid, uint32_t &first, uint32_t &last) {
        first = last = NIL;
        uint32_t prev = NIL;
        for (uint32_t pos = start; pos < end; ++pos) {
            const uint32_t curr = add_node(seq[pos]);
            if (curr == NIL) return DG_CAPACITY;
            if (first == NIL) first = curr;
            if (prev != NIL) {
                const int rc = add_edge(prev, curr, sid, 2);
                if (rc) return rc;
            }
            prev = curr;
        }
        last = prev;
        return DG_OK;
    }

    DG_HD int push_sequence(uint32_t first_node) {
        if (f->n_seqs_in_graph >= f->cap.seqs) return DG_CAPACITY;
        seq_start[f->n_seqs_in_graph++] = first_node;
        return DG_OK;
    }

    // poa.rs:182-198: first sequence of an empty graph (alignment None)
    DG_HD int add_first_sequence(const uint8_t *seq, uint32_t len) {
        const uint32_t sid = f->n_seqs_in_graph;
        if (len == 0) {
            const int rc = push_sequence(START);
            return rc ? rc : post_process();
        }
        uint32_t first, last;
        int rc = add_nodes_for_sequence(seq, 0, len, sid, first, last);
        if (rc) return rc;
        rc = push_sequence(first);
        return rc ? rc : post_process();
    }

    // poa.rs:201-321 with alignment Some(pairs); rp / qp hold n_pairs entries (NIL = None)
    DG_HD int add_alignment(const uint8_t *seq, uint32_t len, const uint32_t *rp, const uint32_t *qp, uint32_t n_pairs) {
        const uint32_t sid = f->n_seqs_in_graph;
        uint32_t first_q = NIL, last_q = 0;
        bool any = false;
        for (uint32_t i = 0; i < n_pairs; ++i)
            if (qp[i] != NIL && qp[i] < len) {
                if (!any) first_q = qp[i];
                last_q = qp[i];
                any = true;
            }
        if (!any) {
            if (len == 0) {
                const int rc = push_sequence(START);
                return rc ? rc : post_process();
            }
            return DG_INVALID_ALIGNMENT;
        }
        uint32_t begin_first, begin_last, end_first, end_last;
        int rc = add_nodes_for_sequence(seq, 0, first_q, sid, begin_first, begin_last);
        if (rc) return rc;
        bool has_begin = begin_first != NIL;
        uint32_t prev = begin_last;
        rc = add_nodes_for_sequence(seq, last_q + 1, len, sid, end_first, end_last);
        if (rc) return rc;
        for (uint32_t i = 0; i < n_pairs; ++i) {
            if (qp[i] == NIL) continue;
            const uint32_t q = qp[i], r = rp[i];
            const uint8_t qs = seq[q];
            uint32_t curr = NIL;
            if (r != NIL) {
                if (sym[r] == qs) {
                    curr = r;
                } else {
                    for (uint32_t k = al_head[r]; k != NIL; k = al_next[k])
                        if (sym[al_node[k]] == qs) {
                            curr = al_node[k];
                            break;
                        }
                    if (curr == NIL) {
                        const uint32_t nn = add_node(qs);
                        if (nn == NIL) return DG_CAPACITY;
                        curr = nn;
                        // r's list as it is now (the pushes below append to OTHER nodes' lists and to nn's, r's comes last)
                        for (uint32_t k = al_head[r]; k != NIL; k = al_next[k]) {
                            const uint32_t other = al_node[k];
                            if ((rc = push_aligned(other, nn))) return rc;
                            if ((rc = push_aligned(nn, other))) return rc;
                        }
                        if ((rc = push_aligned(r, nn))) return rc;
                        if ((rc = push_aligned(nn, r))) return rc;
                    }
                }
            } else {
                curr = add_node(qs);
                if (curr == NIL) return DG_CAPACITY;
            }
            if (!has_begin) {
                begin_first = curr;
                has_begin = true;
            }
            if (prev != NIL && (rc = add_edge(prev, curr, sid, 2))) return rc;
            prev = curr;
        }
        if (end_first != NIL && (rc = add_edge(prev, end_first, sid, 2))) return rc;
        if ((rc = push_sequence(begin_first))) return rc;
        return post_process();
    }

    // poa.rs:323-363 without the toposort
    DG_HD int post_process() {
        while (head_out[START] != NIL) raw_remove_edge(head_out[START]);
        while (head_in[END] != NIL) raw_remove_edge(head_in[END]);
        const uint32_t n = f->n_nodes;
        for (uint32_t v = 2; v < n; ++v)
            if (head_in[v] == NIL) {
                const int rc = raw_add_edge(START, v, 0, NIL);
                if (rc) return rc;
            }
        for (uint32_t v = 2; v < n; ++v)
            if (head_out[v] == NIL) {
                const int rc = raw_add_edge(v, END, 0, NIL);
                if (rc) return rc;
            }
        return DG_OK;
    }
};

// Worst-case capacities for a family that starts from a graph holding (V0 nodes incl. start/end, E0 edge slots, I0 sequence
// ids, A0 aligned-list entries, N0 sequences) and inserts n sequences of total length S.
inline Caps caps_for(uint64_t S, uint32_t n, uint32_t V0 = 2, uint32_t E0 = 0, uint32_t I0 = 0, uint32_t A0 = 0, uint32_t N0 = 0) {
    Caps c;
    c.nodes = (uint32_t)(V0 + S + 2);
    // real edges <= S (one per consecutive pair of bases); start / end edges: every node may be a source AND a sink in the
    // worst case (n single-base sequences), but their slots are recycled each round, so V edges bound both kinds together
    c.edges = (uint32_t)(E0 + S + 2 * ((uint64_t)n + N0) + V0 + 8);
    c.sids = (uint32_t)(I0 + S + 2);
    // aligned lists: a column of k mutually aligned nodes holds k (k - 1) entries; k <= 256 symbols. For k <= 4 (ACGT) that is
    // <= 3 per node; 6 per node leaves room for IUPAC columns, beyond that the family fails with POASTA_ERR_UNSUPPORTED
    c.als = (uint32_t)(A0 + 6 * (S + 2));
    c.seqs = N0 + n + 1;
    return c;
}

}  // namespace dg
}  // namespace poasta
