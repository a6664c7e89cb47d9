// Device launcher for the order-exact search (replay_core.hpp). Included by align.cu.
//
// One warp = one in-flight alignment ("worker"). The worker's V x W visited table is never reset:
// its records carry the epoch of the alignment that wrote them (replay_core.hpp Rec), so the 32 lanes
// only clear the bucket heads and reached bounds the previous alignment touched; then lane 0 runs the
// (inherently sequential) state machine. Workers pull alignments from a global counter, so a launch
// keeps every SM's warp slots busy until the batch is drained. Results go to the same staging / scan /
// compact path as the DP traceback. (Progressive builds use the persistent engine in msa_engine.cuh.)
#pragma once
#include "replay_core.hpp"
#include "search_index.hpp"

namespace poasta {

struct RGraphDesc {
    uint32_t V, start, end;
    uint32_t node_base;                                  // into nodes (NodeRec) ; path_off uses node_base + graph index slack
    uint32_t succ_base, pred_base, bub_base, path_base;  // bases of this graph's succ / pred / bub (quadruples) / path_rem
    uint32_t pathoff_base;                               // V + 1 entries
};

struct ReplayDev {
    const RGraphDesc *graphs;
    const AlnDesc *alns;
    const replay::NodeRec *nodes;
    const uint8_t *queries;
    const uint32_t *succ, *pred, *bub, *path_off, *path_rem;
    uint32_t n_alns;
    replay::Params prm;
    uint8_t *slabs;
    uint64_t slab_bytes, off_qhead, off_items, off_free, off_rb, off_bits, off_dfa;  // byte offsets inside a slab (header at 0, records at SLAB_CELLS)
    uint32_t max_items;
    // blocked visited table (replay_core.hpp BLK_SHIFT): the records area [SLAB_CELLS, off_qhead) is the block pool
    uint64_t off_bmap;
    uint32_t bpr, max_blocks;
    uint32_t *counter, *scores, *n_pairs, *out_r, *out_q;
    int32_t *status;
    uint32_t *stats;  // per alignment: queued, visited, pruned, item high-water mark
};

// ONE warp per CTA: 32 CTAs are resident per SM, exactly the 32 workers the 64-register build allows. With a single warp the
// worker's pointer blocks (replay::Scratch, replay::Graph) sit at fixed shared-memory addresses, so an access to any of the
// worker's dozen arrays takes one LDS for the base instead of the S2R / LDC / IMAD.WIDE chain the compiler kept
// re-materialising from the kernel parameters (12 % of all executed instructions in the round-2 capture, profiles/README.md).
constexpr int REPLAY_MAX_WARPS_PER_CTA = 1;

// Slab header (first 256 bytes of a worker's slab; the visited table starts at SLAB_CELLS). The slab buffer is zeroed by the
// host when it is (re)allocated or its layout changes, so epoch 0 means "never written".
constexpr uint64_t SLAB_CELLS = 256;
struct SlabHeader {
    uint32_t epoch;      // epoch of the last alignment this worker ran
    uint32_t dirty_f;    // priorities [0, dirty_f) of qhead may be non-zero
    uint32_t map_epoch;  // blocked table: epoch of the block map entries the last alignment wrote (0 = none yet)
};
// Start of an alignment on a worker with a blocked visited table: next map epoch; the map is zeroed only when the 12-bit epoch wraps.
__device__ __forceinline__ uint32_t worker_begin_blocks(uint8_t *slab, uint64_t off_bmap, uint64_t map_entries, uint32_t lane) {
    SlabHeader *hdr = reinterpret_cast<SlabHeader *>(slab);
    uint32_t mep = hdr->map_epoch + 1;
    if (mep > replay::BMAP_EPOCH_MAX) {
        uint32_t *m = reinterpret_cast<uint32_t *>(slab + off_bmap);
        for (uint64_t i = lane; i < map_entries; i += 32) m[i] = 0;
        mep = 1;
    }
    __syncwarp();
    if (lane == 0) hdr->map_epoch = mep;
    __syncwarp();
    return mep;
}
template <typename CellT> struct EpochTraits;
template <> struct EpochTraits<uint16_t> { static constexpr uint32_t MAX = 0xFFFEu; };
template <> struct EpochTraits<uint32_t> { static constexpr uint32_t MAX = 0xFFFFFFFEu; };

// Start of one alignment on a worker, all 32 lanes: next epoch (zeroing the table only when the counter wraps), bucket
// heads of the priorities the previous alignment touched, reached bounds and reached-offset bitmap of this graph's nodes,
// the queue window. Returns the epoch (warp-uniform).
template <typename CellT>
__device__ __forceinline__ uint32_t worker_begin_alignment(uint8_t *slab, uint64_t cell_bytes, uint64_t off_qhead, uint64_t off_rb, uint64_t off_bits,
                                                           uint32_t V, uint32_t Wb, uint32_t *s_qwin, uint32_t lane) {
    SlabHeader *hdr = reinterpret_cast<SlabHeader *>(slab);
    uint32_t ep = hdr->epoch + 1, dirty = hdr->dirty_f;
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (ep > EpochTraits<CellT>::MAX) {
        uint4 *p0 = reinterpret_cast<uint4 *>(slab + SLAB_CELLS);
        for (uint64_t i = lane; i < cell_bytes / 16; i += 32) p0[i] = z;
        ep = 1;
    }
    uint4 *p1 = reinterpret_cast<uint4 *>(slab + off_qhead);
    const uint32_t n1 = (uint32_t)(((uint64_t)dirty * 3 + 3) / 4);
    for (uint32_t i = lane; i < n1; i += 32) p1[i] = z;
    uint4 *p2 = reinterpret_cast<uint4 *>(slab + off_rb);
    const uint32_t n2 = (uint32_t)(((uint64_t)V * 2 + 3) / 4);
    for (uint32_t i = lane; i < n2; i += 32) p2[i] = z;
    uint4 *p3 = reinterpret_cast<uint4 *>(slab + off_bits);
    const uint32_t n3 = (uint32_t)(((uint64_t)V * Wb + 3) / 4);
    for (uint32_t i = lane; i < n3; i += 32) p3[i] = z;
    for (uint32_t i = lane; i < replay::QWIN * 3; i += 32) s_qwin[i] = 0;
    __syncwarp();
    return ep;
}
__device__ __forceinline__ void worker_end_alignment(uint8_t *slab, uint32_t epoch, uint32_t f_top) {
    SlabHeader *hdr = reinterpret_cast<SlabHeader *>(slab);
    hdr->epoch = epoch;
    hdr->dirty_f = f_top + 1;
}

// MINB = minimum resident CTAs per SM the register allocation must allow. Only lane 0 of a warp runs the state machine
// but registers are allocated for all 32 lanes, so the register count decides how many alignments are in flight per SM.
template <typename CellT, int MINB, int HEUR, bool BLOCKED = false>
__global__ void __launch_bounds__(32, MINB) replay_kernel(ReplayDev B) {
    __shared__ uint32_t s_qwin[replay::QWIN * 3];  // queue-head window of the worker
    __shared__ replay::Scratch s_scr;              // the worker's pointer blocks at fixed shared-memory addresses
    __shared__ replay::Graph s_gr;
    const uint32_t lane = threadIdx.x;
    uint8_t *slab = B.slabs + (uint64_t)blockIdx.x * B.slab_bytes;
    for (;;) {
        uint32_t a = 0;
        if (lane == 0) a = atomicAdd(B.counter, 1u);
        a = __shfl_sync(0xFFFFFFFFu, a, 0);
        if (a >= B.n_alns) break;
        const AlnDesc ad = B.alns[a];
        const RGraphDesc gd = B.graphs[ad.graph];
        const replay::ScratchDims dims(gd.V, ad.Q, B.prm.x, B.prm.o, B.prm.e);
        const uint32_t Wb = (dims.W + 31) / 32;
        const uint32_t epoch = worker_begin_alignment<CellT>(slab, B.off_qhead - SLAB_CELLS, B.off_qhead, B.off_rb, B.off_bits, gd.V, Wb, s_qwin, lane);
        uint32_t map_epoch = 0;
        if (BLOCKED) map_epoch = worker_begin_blocks(slab, B.off_bmap, (uint64_t)gd.V * B.bpr, lane);
        if (lane == 0) {
            replay::Graph &g = s_gr;
            g.V = gd.V, g.start = gd.start, g.end = gd.end;
            g.nodes = B.nodes + gd.node_base;
            g.succ = B.succ + gd.succ_base;
            g.pred = B.pred + gd.pred_base;
            g.bub = reinterpret_cast<const replay::BubQ *>(B.bub) + gd.bub_base;
            g.path_off = B.path_off + gd.pathoff_base;
            g.path_rem = B.path_rem + gd.path_base;
            replay::Scratch &s = s_scr;
            s.planes = slab + SLAB_CELLS;
            s.epoch = epoch;
            s.qhead = reinterpret_cast<uint32_t *>(slab + B.off_qhead);
            s.qwin = s_qwin;
            s.items = reinterpret_cast<replay::Item *>(slab + B.off_items);
            s.free_slots = reinterpret_cast<uint32_t *>(slab + B.off_free);
            s.rb = reinterpret_cast<uint32_t *>(slab + B.off_rb);
            s.rbits = reinterpret_cast<uint32_t *>(slab + B.off_bits);
            s.Wb = Wb;
            s.dfa = reinterpret_cast<uint32_t *>(slab + B.off_dfa);
            s.W = dims.W, s.f_max = dims.f_max;
            s.max_items = B.max_items;
            s.bmap = reinterpret_cast<uint32_t *>(slab + B.off_bmap), s.bpr = B.bpr, s.max_blocks = B.max_blocks, s.bmap_epoch = map_epoch;
        }
        __syncwarp();
        // every lane holds a view of the search (references into shared memory): the lanes share the per-node start-up work of an
        // ends-free alignment, then lane 0 runs the state machine
        replay::Search<CellT, HEUR, BLOCKED> search(s_gr, B.queries + ad.query_base, ad.Q, B.prm, s_scr);
        const bool prepared = search.prepare_initial(lane, 32u);
        __syncwarp();
        if (lane == 0) {
            const replay::Result r = prepared ? search.template run<true>(B.out_r + ad.out_base, B.out_q + ad.out_base, ad.out_cap)
                                              : search.template run<false>(B.out_r + ad.out_base, B.out_q + ad.out_base, ad.out_cap);
            worker_end_alignment(slab, epoch, search.f_top);
            B.scores[a] = r.score;
            B.n_pairs[a] = r.status == replay::RP_OK ? r.n_pairs : 0;
            B.status[a] = r.status;
            B.stats[4 * a + 0] = r.num_queued, B.stats[4 * a + 1] = r.num_visited, B.stats[4 * a + 2] = r.num_pruned;
            B.stats[4 * a + 3] = search.n_items;
        }
        __syncwarp();
    }
}

// Host-side packing of problems for the search kernel (node-id space + SearchIndex arrays).
struct PackedReplay {
    std::vector<RGraphDesc> graphs;
    std::vector<AlnDesc> alns;
    std::vector<replay::NodeRec> nodes;
    std::vector<uint8_t> queries;
    std::vector<uint32_t> succ, pred, bub, path_off, path_rem;
    uint64_t out_elems = 0, total_cells = 0;
    uint32_t max_V = 0, max_Q = 0;

    void clear() {
        graphs.clear(), alns.clear(), nodes.clear(), queries.clear();
        succ.clear(), pred.clear(), bub.clear(), path_off.clear(), path_rem.clear();
        out_elems = total_cells = 0;
        max_V = max_Q = 0;
    }

    uint32_t add_graph(const poasta_graph_view &v, const SearchIndex &si, const std::vector<replay::NodeRec> *recs = nullptr) {
        RGraphDesc gd;
        const uint32_t V = v.n_nodes;
        gd.V = V, gd.start = v.start_node, gd.end = v.end_node;
        gd.node_base = (uint32_t)nodes.size();
        gd.succ_base = (uint32_t)succ.size();
        gd.pred_base = (uint32_t)pred.size();
        gd.bub_base = (uint32_t)(bub.size() / 4);
        gd.path_base = (uint32_t)path_rem.size();
        gd.pathoff_base = (uint32_t)path_off.size();
        if (recs) nodes.insert(nodes.end(), recs->begin(), recs->end());
        else si.node_records(v, nodes);
        succ.insert(succ.end(), v.succ, v.succ + v.succ_off[V]);
        pred.insert(pred.end(), v.pred, v.pred + v.pred_off[V]);
        bub.insert(bub.end(), si.bub.begin(), si.bub.end());
        if (si.path_off.empty()) {
            path_off.insert(path_off.end(), V + 1, 0u);
        } else {
            path_off.insert(path_off.end(), si.path_off.begin(), si.path_off.end());
            path_rem.insert(path_rem.end(), si.path_rem.begin(), si.path_rem.end());
        }
        max_V = std::max(max_V, V);
        graphs.push_back(gd);
        return (uint32_t)graphs.size() - 1;
    }

    void add_alignment(uint32_t graph, const uint8_t *q, uint32_t Q) {
        AlnDesc a;
        a.graph = graph;
        a.Q = Q;
        a.W = Q + 1;
        a.query_base = (uint32_t)queries.size();
        queries.insert(queries.end(), q, q + Q);
        while (queries.size() & 3) queries.push_back(0);
        a.trace_base = 0;
        a.out_cap = graphs[graph].V + Q + 2;
        a.out_base = (uint32_t)out_elems;
        out_elems += a.out_cap;
        total_cells += (uint64_t)(graphs[graph].V - 2) * Q;
        max_Q = std::max(max_Q, Q);
        alns.push_back(a);
    }
};

}  // namespace poasta
