// Persistent on-device engine for progressive MSA builds in reference-exact mode (bin/poasta.rs:163-236 for thousands of
// families at once). Included by align.cu.
//
// Round 1 ran "round k of every family" as one launch per pipeline lane with the graphs rebuilt, re-indexed, re-packed and
// re-uploaded by host threads between launches: a launch lasted as long as its slowest alignment (max / mean ~ 1.7), the
// host phases needed ~0.6 ms of a core per alignment (which capped 8-GPU scaling on a 32-core box) and every graph crossed
// PCIe every round (164 KB per alignment). Here the whole build stays on the device:
//   * every family's POA graph lives in HBM (dev_graph.hpp) and is updated there after each alignment
//     (add_alignment_with_weights + post_process, graphs/poa.rs:171-363);
//   * ONE persistent kernel per batch: a worker (warp) takes a READY family from a ticket queue, rebuilds the per-alignment
//     index of its graph (dev_index.hpp: adjacency, superbubbles, distances, sampled paths - aligner/config.rs:113-134), runs the
//     order-exact search (replay_core.hpp), inserts the alignment and queues the family again. A family's next round starts
//     when ITS alignment ends (continuous batching), not when the slowest of a launch does; no host work, no copies between
//     rounds. Host traffic per build: the sequences in, the finished graphs out.
// Ticket queue: every family produces exactly n_seqs tickets over its life (one per record), so the queue is a plain array
// of sum(n_seqs) slots; consumers take slot numbers with one atomicAdd and spin (ld.acquire + nanosleep) until the slot is
// filled (st.release after the family's graph update), which also publishes that update to the next worker's SM.
// Workers are sized for the common case (V up to a few query lengths); a family that outgrows them (graph, score range,
// queue arena) is suspended with its state intact and resumed by a further pass with fewer, larger workers.
#pragma once
#include <thread>

#include "dev_graph_host.hpp"
#include "dev_index.hpp"
#include "replay_kernel.cuh"

namespace poasta {

constexpr uint32_t TICKET_NULL = 0xFFFFFFFFu;
constexpr int SEQ_OUT_WORDS = 8;  // per record: score, status, queued, visited, pruned, real nodes at alignment time, search / whole-round time in 1024-cycle units

struct EngineDev {
    dg::Family *fams;
    uint8_t *arena;
    const uint8_t *seq_data;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    uint32_t *seq_out;
    uint32_t *tickets;
    uint32_t n_tickets;
    uint32_t *counters;   // [0] next ticket to consume, [1] next ticket to fill, [2] suspended families
    uint32_t *suspended;  // family ids
    uint8_t *workers;
    uint64_t worker_bytes, cell_bytes, off_qhead, off_items, off_free, off_rb, off_bits, off_dfa, off_pairs, off_index;
    uint32_t max_items, pair_cap, f_cap;
    dg::IndexCaps icap;
    replay::Params prm;
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct WarpSync {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
    __device__ __forceinline__ int bcast(int x) const { return __shfl_sync(0xFFFFFFFFu, x, 0); }
};

// kept out of line: graph maintenance is ~0.1 % of a round and must not shape the register allocation of the search loop
__device__ __noinline__ int engine_build_index(const dg::Store &g, dg::IndexBuf &w, bool with_paths, uint32_t lane) {
    return dg::build_index(g, w, with_paths, lane, 32u, WarpSync{});
}
__device__ __noinline__ int engine_insert(dg::Store &g, const uint8_t *q, uint32_t Q, const uint32_t *rp, const uint32_t *qp, uint32_t n_pairs) {
    return g.add_alignment(q, Q, rp, qp, n_pairs);
}
__device__ __noinline__ int engine_first(dg::Store &g, const uint8_t *q, uint32_t Q) {
    if (g.f->n_nodes == 0) g.init_empty();
    return g.is_empty() ? g.add_first_sequence(q, Q) : -1;
}

enum : int { ENG_CONTINUE = 0, ENG_FAILED = 1, ENG_SUSPEND = 2 };

// 32 resident CTAs of 1 warp per SM -> 64 registers, 32 workers per SM. Measured alternatives (profiles/README.md, round 2):
// 40 registers / 48 workers per SM: 18.2k instead of 13.5k cycles per visited state, 26 % slower at 7,104 families.
template <typename CellT, int HEUR>
__global__ void __launch_bounds__(32, 32) msa_engine_kernel(EngineDev E) {
    __shared__ uint32_t s_qwin[replay::QWIN * 3];
    __shared__ replay::Scratch s_scr;  // the worker's pointer blocks at fixed shared-memory addresses (see replay_kernel.cuh)
    __shared__ replay::Graph s_gr;
    const uint32_t lane = threadIdx.x;
    uint8_t *slab = E.workers + (uint64_t)blockIdx.x * E.worker_bytes;
    for (;;) {
        uint32_t h = 0;
        if (lane == 0) h = atomicAdd(&E.counters[0], 1u);
        h = __shfl_sync(0xFFFFFFFFu, h, 0);
        if (h >= E.n_tickets) break;
        // spin on a relaxed load; ONE acquire once the slot is filled: an acquire at gpu scope invalidates the SM's whole L1
        // (CCTL.IVALL), which a polling loop would do to the 31 other workers of the SM every few hundred nanoseconds
        while (ld_relaxed_u32(&E.tickets[h]) == 0) __nanosleep(1000);
        const uint32_t t = ld_acquire_u32(&E.tickets[h]);
        if (t == TICKET_NULL) continue;
        const uint32_t f = t - 1;
        dg::Family *F = E.fams + f;
        dg::Store g(F, E.arena);
        const uint32_t k = F->next_seq, gi = F->seq_first + k;
        const uint8_t *q = E.seq_data + E.seq_off[gi];
        const uint32_t Q = E.seq_len[gi];
        uint32_t *out = E.seq_out + (size_t)SEQ_OUT_WORDS * gi;
        int outcome = ENG_CONTINUE, fam_status = 0;
        const long long t_round = clock64();
        // first record of an empty graph: plain chain (bin/poasta.rs:209-211)
        int first_rc = -1;
        if (lane == 0) first_rc = engine_first(g, q, Q);
        first_rc = __shfl_sync(0xFFFFFFFFu, first_rc, 0);
        if (first_rc >= 0) {
            if (first_rc != dg::DG_OK) outcome = ENG_FAILED, fam_status = first_rc == dg::DG_CAPACITY ? POASTA_ERR_UNSUPPORTED : POASTA_ERR_INVALID_ALIGNMENT;
            if (lane == 0) out[0] = 0, out[1] = 0, out[2] = out[3] = out[4] = out[5] = 0;
        } else {
            const uint32_t V = F->n_nodes;
            const replay::ScratchDims dims(V, Q, E.prm.x, E.prm.o, E.prm.e);
            const uint64_t m = E.prm.x > E.prm.o + E.prm.e ? E.prm.x : E.prm.o + E.prm.e;
            const uint64_t worst = ((uint64_t)V + Q + 4) * m + E.prm.o;  // every node deleted + every base inserted
            const bool fits = V <= E.icap.V && F->live_edges <= E.icap.E && (uint64_t)V * dims.W * sizeof(replay::Rec<CellT>) <= E.cell_bytes &&
                              dims.f_max <= E.f_cap && V + Q + 2 <= E.pair_cap && (sizeof(CellT) == 4 || worst < replay::CELL16_LIMIT);
            if (!fits) {
                outcome = ENG_SUSPEND;
            } else {
                dg::IndexBuf w(slab + E.off_index, E.icap);
                if (engine_build_index(g, w, HEUR == 2, lane) != dg::DG_OK) {
                    outcome = ENG_SUSPEND;
                } else {
                    const uint32_t Wb = (dims.W + 31) / 32;
                    const uint32_t epoch = worker_begin_alignment<CellT>(slab, E.cell_bytes, E.off_qhead, E.off_rb, E.off_bits, V, Wb, s_qwin, lane);
                    if (lane == 0) {
                        replay::Graph &rg = s_gr;
                        rg.V = V, rg.start = dg::Store::START, rg.end = dg::Store::END;
                        rg.nodes = w.nodes, rg.succ = w.succ, rg.pred = w.pred, rg.bub = w.bub;
                        rg.path_off = w.path_off, rg.path_rem = w.path_rem;
                        replay::Scratch &s = s_scr;
                        s.planes = slab + SLAB_CELLS;
                        s.epoch = epoch;
                        s.qhead = reinterpret_cast<uint32_t *>(slab + E.off_qhead);
                        s.qwin = s_qwin;
                        s.items = reinterpret_cast<replay::Item *>(slab + E.off_items);
                        s.free_slots = reinterpret_cast<uint32_t *>(slab + E.off_free);
                        s.rb = reinterpret_cast<uint32_t *>(slab + E.off_rb);
                        s.rbits = reinterpret_cast<uint32_t *>(slab + E.off_bits);
                        s.Wb = Wb;
                        s.dfa = reinterpret_cast<uint32_t *>(slab + E.off_dfa);
                        s.W = dims.W, s.f_max = dims.f_max;
                        s.max_items = E.max_items;
                        uint32_t *pr = reinterpret_cast<uint32_t *>(slab + E.off_pairs), *pq = pr + E.pair_cap;
                        replay::Search<CellT, HEUR> search(rg, q, Q, E.prm, s);
                        const long long t_search = clock64();
                        const replay::Result r = search.run(pr, pq, E.pair_cap);
                        out[6] = (uint32_t)((clock64() - t_search) >> 10);
                        worker_end_alignment(slab, epoch, search.f_top);
                        if (r.status == replay::RP_OUT_OF_ITEMS || r.status == replay::RP_OUT_OF_BLOCKS || r.status == replay::RP_PRIORITY_RANGE ||
                            r.status == replay::RP_OUT_OF_PAIRS) {
                            outcome = ENG_SUSPEND;  // arena sized for the common case: a later pass has full-size arenas
                        } else {
                            out[0] = r.score, out[2] = r.num_queued, out[3] = r.num_visited, out[4] = r.num_pruned, out[5] = V - 2;
                            if (r.status != replay::RP_OK) {
                                outcome = ENG_FAILED, fam_status = POASTA_ERR_NO_ALIGNMENT;  // the reference panics here
                            } else {
                                const int rc = engine_insert(g, q, Q, pr + (E.pair_cap - r.n_pairs), pq + (E.pair_cap - r.n_pairs), r.n_pairs);
                                if (rc != dg::DG_OK)
                                    outcome = ENG_FAILED, fam_status = rc == dg::DG_CAPACITY ? POASTA_ERR_UNSUPPORTED : POASTA_ERR_INVALID_ALIGNMENT;
                            }
                            out[1] = (uint32_t)fam_status;
                        }
                    }
                    outcome = __shfl_sync(0xFFFFFFFFu, outcome, 0);
                    fam_status = __shfl_sync(0xFFFFFFFFu, fam_status, 0);
                }
            }
        }
        if (lane == 0) {
            const uint32_t n = F->n_seqs;
            out[7] = (uint32_t)((clock64() - t_round) >> 10);
            if (outcome == ENG_CONTINUE) {
                F->next_seq = k + 1;
                if (k + 1 < n) st_release_u32(&E.tickets[atomicAdd(&E.counters[1], 1u)], f + 1);
            } else {
                if (outcome == ENG_SUSPEND) E.suspended[atomicAdd(&E.counters[2], 1u)] = f;
                else F->status = fam_status, out[1] = (uint32_t)fam_status;
                const uint32_t remaining = n - 1 - k;  // tickets this family will never fill
                if (remaining) {
                    const uint32_t t0 = atomicAdd(&E.counters[1], remaining);
                    for (uint32_t i = 0; i < remaining; ++i) st_release_u32(&E.tickets[t0 + i], TICKET_NULL);
                }
            }
        }
        __syncwarp();
    }
}

// Copies the used part of every family arena into a packed buffer (same sub-array order, capacities := used counts), so the
// download moves what the graphs hold, not what they could have held. One warp per family.
__global__ void pack_graphs_kernel(const dg::Family *fams, uint32_t n, const uint8_t *arena, const uint64_t *packed_off, uint8_t *out) {
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n) return;
    const dg::Family F = fams[f];
    dg::Caps used{F.n_nodes, F.n_edge_slots, F.n_sids, F.n_als, F.n_seqs_in_graph};
    const dg::Layout a(F.cap), b(used);
    const uint8_t *src = arena + F.arena_off;
    uint8_t *dst = out + packed_off[f];
    const uint64_t so[16] = {a.sym, a.head_out, a.head_in, a.al_head, a.e_src, a.e_dst, a.e_next_out, a.e_next_in, a.e_weight, a.e_sid_head, a.e_sid_tail,
                             a.sid_val, a.sid_next, a.al_node, a.al_next, a.seq_start};
    const uint64_t dd[17] = {b.sym, b.head_out, b.head_in, b.al_head, b.e_src, b.e_dst, b.e_next_out, b.e_next_in, b.e_weight, b.e_sid_head, b.e_sid_tail,
                             b.sid_val, b.sid_next, b.al_node, b.al_next, b.seq_start, b.total};
    for (int k = 0; k < 16; ++k) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + so[k]);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + dd[k]);
        const uint64_t n16 = (dd[k + 1] - dd[k]) / 16;
        for (uint64_t i = lane; i < n16; i += 32) d4[i] = s4[i];
    }
}

}  // namespace poasta
