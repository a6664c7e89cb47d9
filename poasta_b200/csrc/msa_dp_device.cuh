// Device-resident progressive builds in DP-only mode (tiebreak = 1): the north-star pipeline end to end on the GPU.
// Included by align.cu after msa_engine_host.cuh (shares its family arenas, sequence upload and graph download).
//
// Round k of bin/poasta.rs:163-236 for every family of a chunk at once, with no host work and no copies between rounds:
//   dp_prepare_kernel   one warp per family: the first record of an empty graph becomes the graph (poasta.rs:209-211); otherwise the
//                       family's graph (dev_graph.hpp arena) is packed into the rank-ordered layout of the DP kernels
//                       (dev_dp_pack.hpp: adjacency, reverse post-order, row records, predecessor ranks), its M / D row store is
//                       carved out of the round's trace buffer with one atomicAdd, and the GraphDesc / AlnDesc of the alignment
//                       are written;
//   dp_fill_packed_kernel<true, false | true>   the packed s16x2 fill (queries <= 1023: one warp per alignment; longer: a
//                       thread-block cluster linked through DSMEM), dp_packed_kernel.cuh;
//   traceback_rows_kernel                       the reference's score-equality backtrace on the stored rows;
//   dp_insert_kernel    one warp per alignment: add_alignment_with_weights + post_process on the arena (graphs/poa.rs:171-363).
// The host reads one 32-byte header per round (alignment count, widest row, families deferred because the trace buffer was
// full - those repeat the round in a further sub-round) to size the launches. Host traffic per build: sequences in, finished
// graphs out, exactly as in the reference-exact engine.
//
// Eligibility (checked by the caller, dp_device_eligible): global mode, sequences and existing graphs over ACGT, costs and
// lengths inside the packed kernels' proven 16-bit range, queries <= 8191. Everything else takes the per-round host pipeline
// (msa_group), which can fall back to the 32-bit kernels.
#pragma once

#include "dev_dp_pack.hpp"

namespace poasta {

struct DpHdr {
    uint32_t n_alns, deferred, max_W, failed;  // n_alns: alignments whose rows fit one warp (W <= 1024), slots [0, n_alns)
    uint32_t n_long, pad0, pad1, pad2;         // alignments of longer queries (cluster kernel), slots [n_fams, n_fams + n_long)
    unsigned long long trace_cursor;  // halves of the trace buffer handed out this sub-round
    unsigned long long pad;
};

struct DpDev {
    dg::Family *fams;
    uint8_t *arena;
    const uint8_t *seq_data;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    uint32_t *seq_out;
    uint32_t n_fams, round;
    const uint32_t *node_base, *pred_base, *out_base;  // per family: first element of its region in the node / predecessor / pair pools
    GraphDesc *gdesc;
    AlnDesc *adesc;
    uint32_t *aln_family;
    RowRec *rows;
    uint32_t *node_id;
    uint8_t *node_sym;
    uint32_t *pred_off, *pred_rank;
    uint8_t *workers;
    uint64_t worker_bytes;
    dg::DpPackCaps wcap;
    DpHdr *hdr;
    unsigned long long trace_cap;  // halves
    uint32_t *work_counter;
};

struct DpBuffers {  // grow-only, one set per device
    Buf node_base, pred_base, out_base, gdesc, adesc, aln_family, rows, node_id, node_sym, pred_off, pred_rank, workers, hdr, counter, trace, scores,
        n_pairs, status, out_r, out_q, scratch;
    PinBuf h_hdr;
};

// One warp per family, persistent over a work counter.
__global__ void __launch_bounds__(32) dp_prepare_kernel(DpDev D) {
    const uint32_t lane = threadIdx.x;
    uint8_t *slab = D.workers + (uint64_t)blockIdx.x * D.worker_bytes;
    for (;;) {
        uint32_t f = 0;
        if (lane == 0) f = atomicAdd(D.work_counter, 1u);
        f = __shfl_sync(0xFFFFFFFFu, f, 0);
        if (f >= D.n_fams) break;
        dg::Family *F = D.fams + f;
        if (F->status || F->next_seq != D.round || F->next_seq >= F->n_seqs) continue;
        dg::Store g(F, D.arena);
        const uint32_t k = F->next_seq, gi = F->seq_first + k;
        const uint8_t *q = D.seq_data + D.seq_off[gi];
        const uint32_t Q = D.seq_len[gi];
        uint32_t *out = D.seq_out + (size_t)SEQ_OUT_WORDS * gi;
        int first_rc = -1;
        if (lane == 0) first_rc = engine_first(g, q, Q);
        first_rc = __shfl_sync(0xFFFFFFFFu, first_rc, 0);
        if (first_rc >= 0) {
            if (lane == 0) {
                out[0] = 0, out[1] = 0, out[5] = 0;
                if (first_rc != dg::DG_OK) {
                    F->status = first_rc == dg::DG_CAPACITY ? POASTA_ERR_UNSUPPORTED : POASTA_ERR_INVALID_ALIGNMENT;
                    out[1] = (uint32_t)F->status;
                    atomicAdd(&D.hdr->failed, 1u);
                } else {
                    F->next_seq = k + 1;
                }
            }
            __syncwarp();
            continue;
        }
        const uint32_t V = F->n_nodes, W = (Q + 1 + 7) & ~7u;
        const unsigned long long need = (unsigned long long)V * ((W + 31) / 32) * 64;
        unsigned long long off = 0;
        int ok = 1;
        if (lane == 0) {
            off = atomicAdd(&D.hdr->trace_cursor, need);
            if (off + need > D.trace_cap) {
                atomicAdd(&D.hdr->deferred, 1u);  // the round is repeated for this family once the buffer is free again
                ok = 0;
            }
        }
        ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
        if (!ok) continue;
        const uint32_t nb = D.node_base[f], pb = D.pred_base[f];
        dg::DpPackScratch w(slab, D.wcap);
        dg::DpPackOut o{D.rows + nb, D.node_id + nb, D.node_sym + nb, D.pred_off + nb + f, D.pred_rank + pb, pb};
        uint32_t max_deg = 0;
        const int rc = dg::dp_pack(g, w, o, &max_deg, lane, 32u, WarpSync{});
        if (lane == 0) {
            if (rc != dg::DG_OK || V > D.node_base[f + 1] - nb || F->live_edges > D.pred_base[f + 1] - pb) {
                F->status = POASTA_ERR_UNSUPPORTED;
                out[1] = (uint32_t)F->status;
                atomicAdd(&D.hdr->failed, 1u);
            } else {
                // queries that fit one warp and longer ones go to different kernels: two slot ranges, graph ids relative to the range
                const bool is_long = W > (uint32_t)WARP_MAX_W;
                const uint32_t rel = is_long ? atomicAdd(&D.hdr->n_long, 1u) : atomicAdd(&D.hdr->n_alns, 1u);
                const uint32_t a = is_long ? D.n_fams + rel : rel;
                if (is_long) atomicMax(&D.hdr->max_W, W);
                GraphDesc gd;
                gd.V = V, gd.node_base = nb, gd.predoff_base = nb + f, gd.max_in_degree = max_deg, gd.n_stored = 0, gd.n_ord_rows = 0, gd.pred_base = pb;
                AlnDesc ad;
                ad.graph = rel, ad.Q = Q, ad.W = W, ad.query_base = (uint32_t)D.seq_off[gi];
                ad.trace_base = 0, ad.ptrace_base = 0;
                ad.out_base = D.out_base[f], ad.out_cap = V + Q + 2;
                ad.prow_base = off;
                D.gdesc[a] = gd;
                D.adesc[a] = ad;
                D.aln_family[a] = f;
            }
        }
        __syncwarp();
    }
}

// One warp per alignment (lane 0 works: the update is a chain of dependent list operations).
__global__ void __launch_bounds__(32) dp_insert_kernel(DpDev D, BatchDev P, uint32_t slot_base) {
    const uint32_t a = blockIdx.x;
    if (a >= P.n_alns || threadIdx.x != 0) return;
    const uint32_t f = D.aln_family[slot_base + a];
    dg::Family *F = D.fams + f;
    dg::Store g(F, D.arena);
    const uint32_t k = F->next_seq, gi = F->seq_first + k;
    const uint8_t *q = D.seq_data + D.seq_off[gi];
    const uint32_t Q = D.seq_len[gi];
    uint32_t *out = D.seq_out + (size_t)SEQ_OUT_WORDS * gi;
    const AlnDesc ad = P.alns[a];
    int st = P.status[a];
    out[0] = P.scores[a];
    out[5] = P.graphs[a].V - 2;
    if (st == 0) {
        const uint32_t np = P.n_pairs[a];
        const uint32_t *rp = P.out_r + ad.out_base + (ad.out_cap - np), *qp = P.out_q + ad.out_base + (ad.out_cap - np);
        const int rc = engine_insert(g, q, Q, rp, qp, np);
        if (rc != dg::DG_OK) st = rc == dg::DG_CAPACITY ? POASTA_ERR_UNSUPPORTED : POASTA_ERR_INVALID_ALIGNMENT;
    }
    out[1] = (uint32_t)st;
    if (st) {
        F->status = st;
        atomicAdd(&D.hdr->failed, 1u);
    } else {
        F->next_seq = k + 1;
    }
}

// Can the device pipeline take this batch? (see the header comment)
static bool dp_device_eligible(const std::vector<EngineFamilyIn> &fams, const poasta_params &prm) {
    static const bool force_host = getenv("POASTA_DP_HOST") != nullptr || getenv("POASTA_DP_KERNEL") != nullptr;
    if (force_host || prm.mode != 0 || prm.tiebreak != 1) return false;
    if ((uint64_t)prm.gap_extend * 1024 > PK_KMAX) return false;
    auto acgt = [](uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; };
    for (const EngineFamilyIn &in : fams) {
        if (in.recs->empty()) continue;
        uint64_t qmax = 0, n_seqs = in.recs->size() + in.graph->g.sequences.size();
        for (const auto &r : *in.recs) {
            if (r.seq.empty()) return false;  // empty queries go through the search kernel (host pipeline)
            qmax = std::max<uint64_t>(qmax, r.seq.size());
            for (unsigned char ch : r.seq)
                if (!acgt(ch)) return false;
        }
        if (qmax > 8191 || n_seqs + 2 > 8191) return false;
        // shortest start -> end path: at most the first sequence of the graph plus one edge (paths of inserted sequences persist)
        uint64_t lmin;
        const POAGraph &g = in.graph->g;
        if (g.nodes.size() > 2) {
            lmin = g.nodes.size();
            for (size_t v = 0; v < g.nodes.size(); ++v)
                if (v != g.start_node && v != g.end_node && !acgt(g.nodes[v].symbol)) return false;
        } else {
            lmin = (*in.recs)[0].seq.size() + 1;
        }
        const uint64_t worst = 2ull * prm.gap_open + (uint64_t)prm.gap_extend * (lmin + qmax) + prm.mismatch;
        if (worst >= PK_INF) return false;
    }
    return true;
}

static void ensure_packed_kernel_attrs(int device) {
    static bool attr_set[MAX_DEVICES] = {};
    if (attr_set[device]) return;
    cudaFuncSetAttribute(dp_fill_packed_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM_BYTES);
    cudaFuncSetAttribute(dp_fill_packed_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM_BYTES);
    cudaFuncSetAttribute(dp_fill_packed_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_CLUSTER_SMEM);
    attr_set[device] = true;
}

// All rounds of one chunk (families already uploaded into eb.fams / eb.arena, sequences into eb.seq_*). `hf` = the chunk's family
// headers as uploaded (capacities, record ranges).
static int msa_dp_rounds(DeviceContext &c, EngineBuffers &eb, DpBuffers &db, const std::vector<dg::Family> &hf, uint32_t n_records, uint64_t seq_bytes,
                         const std::vector<uint32_t> &seq_len, const poasta_params &prm, EngineStats &st) {
    cudaStream_t s = c.stream;
    const uint32_t nf = (uint32_t)hf.size();
    if (seq_bytes >= 0xFFFFFFFFull) return POASTA_ERR_UNSUPPORTED;
    // ---- static regions per family: nodes / predecessors by the arena's capacities, pairs by nodes + longest record ----
    std::vector<uint32_t> node_base(nf + 1), pred_base(nf + 1), out_base(nf + 1);
    uint64_t nb = 0, pb = 0, ob = 0, upper_trace = 0, est_trace = 0;
    uint32_t capV = 0, capE = 0, max_rounds = 0;
    for (uint32_t f = 0; f < nf; ++f) {
        node_base[f] = (uint32_t)nb, pred_base[f] = (uint32_t)pb, out_base[f] = (uint32_t)ob;
        uint32_t qmax = 0;
        for (uint32_t k = 0; k < hf[f].n_seqs; ++k) qmax = std::max(qmax, seq_len[hf[f].seq_first + k]);
        nb += hf[f].cap.nodes, pb += hf[f].cap.edges, ob += (uint64_t)hf[f].cap.nodes + qmax + 2;
        capV = std::max(capV, hf[f].cap.nodes), capE = std::max(capE, hf[f].cap.edges);
        max_rounds = std::max(max_rounds, hf[f].n_seqs);
        const uint64_t row_bytes = (uint64_t)((((qmax + 1 + 7) & ~7u) + 31) / 32) * 128;
        upper_trace += (uint64_t)hf[f].cap.nodes * row_bytes;
        est_trace += std::min<uint64_t>(hf[f].cap.nodes, (uint64_t)hf[f].n_nodes + 4ull * qmax + 2) * row_bytes;
    }
    if (nb + nf >= 0xFFFFFFFFull || pb >= 0xFFFFFFFFull || ob >= 0xFFFFFFFFull) return POASTA_ERR_UNSUPPORTED;
    node_base[nf] = (uint32_t)nb, pred_base[nf] = (uint32_t)pb, out_base[nf] = (uint32_t)ob;
    const dg::DpPackCaps wcap{capV, capE};
    const uint64_t worker_bytes = (dg::DpPackLayout(wcap).total + 255) & ~(uint64_t)255;
    uint64_t n_workers = std::max<uint64_t>(1, std::min<uint64_t>(nf, (uint64_t)c.n_sms * 16));
    while (n_workers > 1 && n_workers * worker_bytes > ((uint64_t)8 << 30)) n_workers /= 2;
    if (!db.node_base.reserve((size_t)(nf + 1) * 4) || !db.pred_base.reserve((size_t)(nf + 1) * 4) || !db.out_base.reserve((size_t)(nf + 1) * 4) ||
        !db.gdesc.reserve((size_t)2 * nf * sizeof(GraphDesc)) || !db.adesc.reserve((size_t)2 * nf * sizeof(AlnDesc)) ||
        !db.aln_family.reserve((size_t)2 * nf * 4) ||
        !db.rows.reserve(nb * sizeof(RowRec) + 16) || !db.node_id.reserve(nb * 4 + 16) || !db.node_sym.reserve(nb + 16) ||
        !db.pred_off.reserve((nb + nf + 1) * 4) || !db.pred_rank.reserve(pb * 4 + 16) || !db.workers.reserve(n_workers * worker_bytes) ||
        !db.hdr.reserve(sizeof(DpHdr)) || !db.counter.reserve(64) || !db.scores.reserve((size_t)2 * nf * 4) || !db.n_pairs.reserve((size_t)2 * nf * 4) ||
        !db.status.reserve((size_t)2 * nf * 4) || !db.out_r.reserve(ob * 4 + 16) || !db.out_q.reserve(ob * 4 + 16) || !db.scratch.reserve(256) ||
        !db.h_hdr.reserve(sizeof(DpHdr)))
        return POASTA_ERR_CUDA;
    // The trace buffer (M / D rows of every alignment of a round) is sized for graphs of four query lengths per family, bounded by what
    // the device has left; a round that needs more is split into sub-rounds. It is kept across calls (the reference-exact engine gives
    // it back before it sizes its worker slabs, and its slabs are given back here if they are in the way).
    uint64_t want_trace = std::min<uint64_t>(upper_trace, est_trace) + 256;
    if (const char *e = getenv("POASTA_DP_TRACE_BYTES")) {  // test knob: forces sub-rounds
        db.trace.release();
        want_trace = std::min<uint64_t>(want_trace, strtoull(e, nullptr, 10));
    }
    if (db.trace.cap < want_trace) {
        db.trace.release();
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if (free_b < want_trace + ((size_t)4 << 30) && eb.workers.d) {
            eb.workers.release();
            eb.workers_ptr = nullptr, eb.workers_sig = 0, eb.workers_zeroed = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
        }
        static const double trace_fraction = getenv("POASTA_DP_TRACE_FRACTION") ? atof(getenv("POASTA_DP_TRACE_FRACTION")) : 0.85;
        const uint64_t take = std::min<uint64_t>(want_trace, (uint64_t)(trace_fraction * (double)free_b));
        if (cudaMalloc(&db.trace.d, take) != cudaSuccess) {
            cudaGetLastError();
            db.trace.d = nullptr;
            return POASTA_ERR_CUDA;
        }
        db.trace.cap = take;
    }
    const uint64_t trace_bytes = db.trace.cap;
    CK(cudaMemcpyAsync(db.node_base.d, node_base.data(), (size_t)(nf + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(db.pred_base.d, pred_base.data(), (size_t)(nf + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(db.out_base.d, out_base.data(), (size_t)(nf + 1) * 4, cudaMemcpyHostToDevice, s));
    tl_h2d += (uint64_t)(nf + 1) * 12;
    ensure_packed_kernel_attrs(c.device);

    DpDev D;
    D.fams = (dg::Family *)eb.fams.d, D.arena = (uint8_t *)eb.arena.d;
    D.seq_data = (const uint8_t *)eb.seq_data.d, D.seq_off = (const uint64_t *)eb.seq_off.d, D.seq_len = (const uint32_t *)eb.seq_len.d;
    D.seq_out = (uint32_t *)eb.seq_out.d;
    D.n_fams = nf, D.round = 0;
    D.node_base = (const uint32_t *)db.node_base.d, D.pred_base = (const uint32_t *)db.pred_base.d, D.out_base = (const uint32_t *)db.out_base.d;
    D.gdesc = (GraphDesc *)db.gdesc.d, D.adesc = (AlnDesc *)db.adesc.d, D.aln_family = (uint32_t *)db.aln_family.d;
    D.rows = (RowRec *)db.rows.d, D.node_id = (uint32_t *)db.node_id.d, D.node_sym = (uint8_t *)db.node_sym.d;
    D.pred_off = (uint32_t *)db.pred_off.d, D.pred_rank = (uint32_t *)db.pred_rank.d;
    D.workers = (uint8_t *)db.workers.d, D.worker_bytes = worker_bytes, D.wcap = wcap;
    D.hdr = (DpHdr *)db.hdr.d, D.trace_cap = trace_bytes / 2;
    D.work_counter = (uint32_t *)db.counter.d + 8;

    BatchDev P;
    P.graphs = D.gdesc, P.alns = D.adesc;
    P.node_sym = D.node_sym, P.node_info = nullptr, P.node_id = D.node_id, P.ord_slot = nullptr;
    P.pred_off = D.pred_off, P.pred_rank = D.pred_rank, P.pred_slot = D.pred_rank;  // pred_slot is read only by the scores-only fill
    P.rows = D.rows;
    P.queries = D.seq_data;
    P.x = (int)prm.mismatch, P.oe = (int)(prm.gap_open + prm.gap_extend), P.e = (int)prm.gap_extend;
    P.scratch = (int *)db.scratch.d, P.scratch_stride = 0;
    P.trace = db.trace.d;
    P.scores = (uint32_t *)db.scores.d;
    P.counter = (uint32_t *)db.counter.d;
    P.out_r = (uint32_t *)db.out_r.d, P.out_q = (uint32_t *)db.out_q.d;
    P.n_pairs = (uint32_t *)db.n_pairs.d, P.status = (int32_t *)db.status.d;

    int packed_ctas = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&packed_ctas, dp_fill_packed_kernel<true, false>, PK_WARPS * 32, PK_SMEM_BYTES);
    packed_ctas = std::max(1, packed_ctas);
    static const bool timing = getenv("POASTA_TIMING") != nullptr;
    tl_dp_kernel = 0;
    std::vector<cudaEvent_t> fill_events;  // pairs around every fill launch (the dominant kernel, timed alone for the roofline)
    struct EventGuard {
        std::vector<cudaEvent_t> &v;
        ~EventGuard() {
            for (cudaEvent_t e : v) cudaEventDestroy(e);
        }
    } event_guard{fill_events};
    CK(cudaEventRecord(c.ev0, s));
    DpHdr &h = *(DpHdr *)db.h_hdr.h;
    for (uint32_t round = 0; round < max_rounds; ++round) {
        D.round = round;
        for (int sub = 0;; ++sub) {
            CK(cudaMemsetAsync(db.hdr.d, 0, sizeof(DpHdr), s));
            CK(cudaMemsetAsync(db.counter.d, 0, 64, s));  // [0] / [4] fill work queues, [1] / [5] cluster abort flags, [8] prepare work queue
            dp_prepare_kernel<<<(unsigned)n_workers, 32, 0, s>>>(D);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&h, db.hdr.d, sizeof(DpHdr), cudaMemcpyDeviceToHost, s));
            CK(wait_stream(c));
            ++st.launches;
            for (int part = 0; part < 2; ++part) {  // 0: rows that fit one warp, 1: longer queries on the cluster kernel
                const uint32_t n = part ? h.n_long : h.n_alns, base = part ? nf : 0;
                if (!n) continue;
                BatchDev Q = P;
                Q.graphs = D.gdesc + base, Q.alns = D.adesc + base;
                Q.scores = P.scores + base, Q.n_pairs = P.n_pairs + base, Q.status = P.status + base;
                Q.counter = P.counter + 4 * part;  // each launch has its own work queue / abort flag
                Q.n_alns = n;
                CK(cudaMemsetAsync(Q.scores, 0x3F, (size_t)n * 4, s));
                cudaEvent_t f0 = nullptr, f1 = nullptr;
                CK(cudaEventCreate(&f0));
                fill_events.push_back(f0);
                CK(cudaEventCreate(&f1));
                fill_events.push_back(f1);
                CK(cudaEventRecord(f0, s));
                if (part) {
                    const uint32_t n_seg_max = (h.max_W + 1023) / 1024;
                    if (n_seg_max > 8) return POASTA_ERR_INTERNAL;  // dp_device_eligible bounds the queries
                    const unsigned cluster_ctas = n_seg_max <= 2 ? 2 : (n_seg_max <= 4 ? 4 : 8);
                    cudaLaunchConfig_t cfg = {};
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = cluster_ctas, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
                    cfg.blockDim = dim3(32), cfg.dynamicSmemBytes = PK_CLUSTER_SMEM, cfg.stream = s, cfg.attrs = at, cfg.numAttrs = 1;
                    cfg.gridDim = dim3(cluster_ctas);
                    int max_clusters = 0;
                    if (cudaOccupancyMaxActiveClusters(&max_clusters, dp_fill_packed_kernel<true, true>, &cfg) != cudaSuccess || max_clusters < 1) max_clusters = 1;
                    cfg.gridDim = dim3((unsigned)std::min<uint64_t>(n, (uint64_t)max_clusters) * cluster_ctas);
                    CK(cudaLaunchKernelEx(&cfg, dp_fill_packed_kernel<true, true>, Q));
                    tl_dp_kernel = 4;  // sticky for the call: some round needed the cluster kernel
                } else {
                    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c.n_sms * packed_ctas, ((uint64_t)n + PK_WARPS - 1) / PK_WARPS);
                    dp_fill_packed_kernel<true, false><<<grid, PK_WARPS * 32, PK_SMEM_BYTES, s>>>(Q);
                    if (tl_dp_kernel != 4) tl_dp_kernel = 3;
                }
                CK(cudaEventRecord(f1, s));
                traceback_rows_kernel<<<(n + 3) / 4, 128, 0, s>>>(Q);
                dp_insert_kernel<<<n, 32, 0, s>>>(D, Q, base);
                CK(cudaGetLastError());
                st.launches += 3;
            }
            if (timing)
                fprintf(stderr, "poasta_b200 dp round %u.%d: %u + %u alignments (one warp / cluster), widest row %u, %u deferred, %u failed, trace %.1f of %.1f GB\n", round, sub,
                        h.n_alns, h.n_long, h.max_W, h.deferred, h.failed, 2.0 * (double)std::min<unsigned long long>(h.trace_cursor, D.trace_cap) / 1e9,
                        (double)trace_bytes / 1e9);
            if (h.deferred == 0) break;
            if (h.n_alns + h.n_long == 0) return POASTA_ERR_UNSUPPORTED;  // one alignment's rows do not fit the device
        }
    }
    CK(cudaEventRecord(c.ev1, s));
    CK(wait_stream(c));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    st.kernel_ms += ms;
    ++st.passes;
    for (size_t i = 0; i + 1 < fill_events.size(); i += 2) {
        float fms = 0;
        if (cudaEventElapsedTime(&fms, fill_events[i], fill_events[i + 1]) == cudaSuccess) st.fill_ms += fms, ++st.fill_launches;
    }
    return 0;
}

}  // namespace poasta
