// Host driver of the persistent MSA engine (msa_engine.cuh). Included by align.cu after DeviceContext / Buf / PinBuf / CK.
#pragma once

namespace poasta {

struct EngineBuffers {  // grow-only, one set per device context
    Buf fams, arena, seq_data, seq_off, seq_len, seq_out, tickets, counters, suspended, workers, packed_off, packed;
    PinBuf h_fams, h_seq_out, h_packed, h_in;
    void *workers_ptr = nullptr;
    uint64_t workers_sig = 0, workers_zeroed = 0;
};

struct EngineFamilyIn {
    poasta_graph *graph;                    // updated in place
    const std::vector<FastaRecord> *recs;   // records to align / insert, in order
    uint32_t *scores;                       // recs->size() entries or nullptr
};

struct EngineStats {
    uint64_t cells = 0, queued = 0, visited = 0, pruned = 0;
    double kernel_ms = 0;
    double fill_ms = 0;  // DP-only mode: the fill launches alone
    uint32_t launches = 0, passes = 0, fill_launches = 0;
};

// worker scratch layout for given capacities
struct WorkerLayout {
    uint64_t cell_bytes, off_qhead, off_items, off_free, off_rb, off_bits, off_dfa, off_pairs, off_index, total;
    uint32_t max_items, pair_cap, f_cap;
    dg::IndexCaps icap;
    WorkerLayout(uint32_t Vw, uint32_t Qmax, bool cell16, bool full_items, const poasta_params &prm) {
        auto up = [](uint64_t b) { return (b + 255) & ~(uint64_t)255; };
        const replay::ScratchDims d(Vw, Qmax, prm.mismatch, prm.gap_open, prm.gap_extend);
        const uint64_t cells = (uint64_t)Vw * d.W;
        cell_bytes = up(cells * (cell16 ? 8 : 16));
        f_cap = d.f_max;
        max_items = (uint32_t)std::min<uint64_t>(0xFFFFFFF0u, full_items ? 262144 + 2 * cells : 32768 + cells / 32) & ~7u;  // whole chunks
        pair_cap = Vw + Qmax + 2;
        icap = dg::IndexCaps{Vw, 4 * Vw + 64, 8 * Vw + 64};
        off_qhead = SLAB_CELLS + cell_bytes;
        off_items = off_qhead + up((uint64_t)f_cap * 12);
        off_free = off_items + up((uint64_t)max_items * 16);
        off_rb = off_free + up((uint64_t)max_items * 4);
        off_bits = off_rb + up((uint64_t)Vw * 8);
        off_dfa = off_bits + up((uint64_t)Vw * ((d.W + 31) / 32) * 4);
        off_pairs = off_dfa + up((uint64_t)(Vw + 2) * 24);
        off_index = off_pairs + up((uint64_t)pair_cap * 8);
        total = off_index + up(dg::IndexLayout(icap).total);
    }
};

}  // namespace poasta

#include "msa_dp_device.cuh"

namespace poasta {

// Progressive builds of `n` families on one device. Families are processed in chunks that fit the arena budget.
// prm.tiebreak == 0: the reference-exact engine (msa_engine.cuh). prm.tiebreak == 1 (the caller has checked dp_device_eligible):
// the DP-only rounds of msa_dp_device.cuh on the same arenas. Upload and download are shared.
static int msa_engine_run(DeviceContext &c, EngineBuffers &eb, DpBuffers &db, std::vector<EngineFamilyIn> &fams_in, const poasta_params &prm,
                          EngineStats &st, unsigned host_threads) {
    CK(cudaSetDevice(c.device));
    cudaStream_t s = c.stream;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    static const double arena_fraction = getenv("POASTA_ARENA_FRACTION") ? atof(getenv("POASTA_ARENA_FRACTION")) : 0.25;
    static const double worker_fraction = getenv("POASTA_SLAB_FRACTION") ? atof(getenv("POASTA_SLAB_FRACTION")) : 0.55;
    const uint64_t arena_budget = (uint64_t)(arena_fraction * (double)total_b);
    const uint64_t worker_budget = (uint64_t)(worker_fraction * (double)total_b);
    const uint64_t m = std::max<uint64_t>(prm.mismatch, (uint64_t)prm.gap_open + prm.gap_extend);

    static const bool phase_timing = getenv("POASTA_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_upload = 0, t_device = 0, t_copy_back = 0, t_convert = 0;
    size_t next = 0;
    std::vector<dg::Family> hf;
    std::vector<uint64_t> seq_off;
    std::vector<uint32_t> seq_len, tickets, chunk_ids;
    while (next < fams_in.size()) {
        // ---- one chunk: families [next, end) whose arenas fit the budget ----
        const double t_chunk0 = now();
        hf.clear(), seq_off.clear(), seq_len.clear(), chunk_ids.clear();
        uint64_t arena_bytes = 0, seq_bytes = 0;
        uint32_t Qmax = 0, V0max = 2, n_records = 0;
        uint64_t Vworst = 0;
        size_t end = next;
        for (; end < fams_in.size(); ++end) {
            const EngineFamilyIn &in = fams_in[end];
            if (in.recs->empty()) continue;
            uint64_t S = 0;
            for (const auto &r : *in.recs) S += r.seq.size();
            const dg::Occupancy occ = dg::occupancy(in.graph->g);
            if (S + occ.V + 64 >= 0x7FFFFFFFull) return POASTA_ERR_UNSUPPORTED;
            dg::Family F;
            memset(&F, 0, sizeof F);
            F.cap = dg::caps_for(S, (uint32_t)in.recs->size(), occ.V, occ.E, occ.I, occ.A, occ.N);
            const uint64_t bytes = (dg::Layout(F.cap).total + 255) & ~(uint64_t)255;
            if (!hf.empty() && arena_bytes + bytes > arena_budget) break;
            F.arena_off = arena_bytes;
            arena_bytes += bytes;
            F.seq_first = n_records, F.n_seqs = (uint32_t)in.recs->size();
            F.n_nodes = 0;  // 0 = fresh: the kernel creates start / end; an existing graph is written below
            F.free_edge = dg::NIL;
            for (const auto &r : *in.recs) {
                seq_off.push_back(seq_bytes);
                seq_len.push_back((uint32_t)r.seq.size());
                seq_bytes += (r.seq.size() + 3) & ~(size_t)3;
                Qmax = std::max<uint32_t>(Qmax, (uint32_t)r.seq.size());
            }
            n_records += F.n_seqs;
            V0max = std::max(V0max, occ.V);
            Vworst = std::max<uint64_t>(Vworst, (uint64_t)occ.V + S + 2);
            hf.push_back(F);
            chunk_ids.push_back((uint32_t)end);
        }
        next = end;
        const uint32_t nf = (uint32_t)hf.size();
        if (nf == 0) continue;
        // ---- upload: family headers, sequences, existing graphs ----
        if (!eb.fams.reserve((size_t)nf * sizeof(dg::Family)) || !eb.arena.reserve(arena_bytes + 256) || !eb.seq_data.reserve(seq_bytes + 16) ||
            !eb.seq_off.reserve((size_t)n_records * 8) || !eb.seq_len.reserve((size_t)n_records * 4) ||
            !eb.seq_out.reserve((size_t)n_records * SEQ_OUT_WORDS * 4) || !eb.tickets.reserve((size_t)n_records * 4) || !eb.counters.reserve(16) ||
            !eb.suspended.reserve((size_t)nf * 4) || !eb.packed_off.reserve((size_t)nf * 8))
            return POASTA_ERR_CUDA;
        if (!eb.h_in.reserve(seq_bytes + 16)) return POASTA_ERR_CUDA;
        {
            uint8_t *hs = (uint8_t *)eb.h_in.h;
            size_t r = 0;
            for (uint32_t i = 0; i < nf; ++i)
                for (const auto &rec : *fams_in[chunk_ids[i]].recs) {
                    memcpy(hs + seq_off[r], rec.seq.data(), rec.seq.size());
                    ++r;
                }
        }
        CK(cudaMemcpyAsync(eb.seq_data.d, eb.h_in.h, seq_bytes, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(eb.seq_off.d, seq_off.data(), (size_t)n_records * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(eb.seq_len.d, seq_len.data(), (size_t)n_records * 4, cudaMemcpyHostToDevice, s));
        tl_h2d += seq_bytes + (uint64_t)n_records * 12 + (uint64_t)nf * sizeof(dg::Family);
        std::vector<uint8_t> image;
        for (uint32_t i = 0; i < nf; ++i) {
            const POAGraph &g = fams_in[chunk_ids[i]].graph->g;
            if (g.nodes.size() <= 2 && g.sequences.empty()) continue;
            image.assign(dg::Layout(hf[i].cap).total, 0);
            dg::Family tmp = hf[i];
            tmp.arena_off = 0;
            if (!dg::upload(g, &tmp, image.data())) return POASTA_ERR_UNSUPPORTED;
            tmp.arena_off = hf[i].arena_off;
            hf[i] = tmp;
            CK(cudaMemcpyAsync((uint8_t *)eb.arena.d + hf[i].arena_off, image.data(), image.size(), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));  // `image` is reused
            tl_h2d += image.size();
        }
        CK(cudaMemcpyAsync(eb.fams.d, hf.data(), (size_t)nf * sizeof(dg::Family), cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(eb.seq_out.d, 0, (size_t)n_records * SEQ_OUT_WORDS * 4, s));

        const double t_chunk1 = now();
        if (prm.tiebreak != 0) {
            const int rc = msa_dp_rounds(c, eb, db, hf, n_records, seq_bytes, seq_len, prm, st);
            if (rc) return rc;
        }
        // ---- passes: common-case workers first, larger ones for the families that were suspended ----
        std::vector<uint32_t> todo(prm.tiebreak != 0 ? 0 : nf);  // family ids (chunk-local) still to run, with the records they have left
        for (uint32_t i = 0; i < (uint32_t)todo.size(); ++i) todo[i] = i;
        std::vector<uint32_t> left(nf);
        for (uint32_t i = 0; i < nf; ++i) left[i] = hf[i].n_seqs;
        // largest V a 16-bit table can carry: (V + Q + 4) * m + o < CELL16_LIMIT
        const uint64_t v16 = replay::CELL16_LIMIT > prm.gap_open ? (replay::CELL16_LIMIT - 1 - prm.gap_open) / std::max<uint64_t>(m, 1) : 0;
        const uint64_t Vmax16 = v16 > (uint64_t)Qmax + 4 ? v16 - Qmax - 4 : 0;
        uint64_t Vw = std::min<uint64_t>(Vworst, std::max<uint64_t>((uint64_t)V0max + 2 * (uint64_t)Qmax + 64, 256));
        bool full_items = false;
        // test knob: start with workers of this many nodes (families outgrow them and are resumed by later passes)
        static const uint64_t forced_nodes = getenv("POASTA_WORKER_NODES") ? strtoull(getenv("POASTA_WORKER_NODES"), nullptr, 10) : 0;
        if (forced_nodes) Vw = std::min<uint64_t>(Vworst, forced_nodes);
        for (int pass = 0; !todo.empty(); ++pass) {
            if (pass > 12) return POASTA_ERR_UNSUPPORTED;
            bool cell16 = Vw <= Vmax16;
            if (const char *e = getenv("POASTA_CELL_BITS")) cell16 = cell16 && atoi(e) == 16;
            // as many nodes per worker as the budget allows while every warp slot (or every family) still gets a worker
            uint64_t want = std::min<uint64_t>(todo.size(), (uint64_t)c.n_sms * 32);
            if (pass == 0 && !forced_nodes) {
                uint64_t hi = std::min<uint64_t>(Vworst, cell16 ? Vmax16 : Vworst);
                while (hi > Vw && WorkerLayout((uint32_t)hi, Qmax, cell16, full_items, prm).total * want > worker_budget)
                    hi = Vw + (hi - Vw) / 2;
                Vw = std::max(Vw, hi);
            }
            WorkerLayout L((uint32_t)Vw, Qmax, cell16, full_items, prm);
            uint64_t workers = std::max<uint64_t>(1, std::min<uint64_t>(want, worker_budget / L.total));
            const unsigned wpc = 1;  // one warp per CTA (replay_kernel.cuh)
            workers = (workers + wpc - 1) / wpc * wpc;
            if (!eb.workers.reserve(workers * L.total)) {
                workers = wpc;
                if (!eb.workers.reserve(workers * L.total)) return POASTA_ERR_CUDA;
            }
            {
                const uint64_t sig = L.total * 1000003ull + L.off_qhead * 10007ull + L.off_items * 101ull + L.off_index + (cell16 ? 1ull << 62 : 0);
                if (eb.workers_ptr != eb.workers.d || eb.workers_sig != sig) eb.workers_ptr = eb.workers.d, eb.workers_sig = sig, eb.workers_zeroed = 0;
                if (workers * L.total > eb.workers_zeroed) {
                    CK(cudaMemsetAsync((uint8_t *)eb.workers.d + eb.workers_zeroed, 0, workers * L.total - eb.workers_zeroed, s));
                    eb.workers_zeroed = workers * L.total;
                }
            }
            // tickets: one slot per record still to go; the first todo.size() are filled now
            uint32_t n_tickets = 0;
            for (uint32_t f : todo) n_tickets += left[f];
            tickets.assign(n_tickets, 0);
            for (size_t i = 0; i < todo.size(); ++i) tickets[i] = todo[i] + 1;
            const uint32_t counters[4] = {0, (uint32_t)todo.size(), 0, 0};
            CK(cudaMemcpyAsync(eb.tickets.d, tickets.data(), (size_t)n_tickets * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(eb.counters.d, counters, 16, cudaMemcpyHostToDevice, s));
            EngineDev E;
            E.fams = (dg::Family *)eb.fams.d;
            E.arena = (uint8_t *)eb.arena.d;
            E.seq_data = (const uint8_t *)eb.seq_data.d;
            E.seq_off = (const uint64_t *)eb.seq_off.d;
            E.seq_len = (const uint32_t *)eb.seq_len.d;
            E.seq_out = (uint32_t *)eb.seq_out.d;
            E.tickets = (uint32_t *)eb.tickets.d;
            E.n_tickets = n_tickets;
            E.counters = (uint32_t *)eb.counters.d;
            E.suspended = (uint32_t *)eb.suspended.d;
            E.workers = (uint8_t *)eb.workers.d;
            E.worker_bytes = L.total, E.cell_bytes = L.cell_bytes;
            E.off_qhead = L.off_qhead, E.off_items = L.off_items, E.off_free = L.off_free, E.off_rb = L.off_rb, E.off_bits = L.off_bits, E.off_dfa = L.off_dfa;
            E.off_pairs = L.off_pairs, E.off_index = L.off_index;
            E.max_items = L.max_items, E.pair_cap = L.pair_cap, E.f_cap = L.f_cap;
            E.icap = L.icap;
            E.prm = replay::Params{prm.mismatch, prm.gap_open, prm.gap_extend, prm.mode, prm.heuristic};
            CK(cudaEventRecord(c.ev0, s));
            {
                const dim3 grid((unsigned)(workers / wpc)), block(wpc * 32);
                static const int carveout = getenv("POASTA_ENGINE_CARVEOUT") ? atoi(getenv("POASTA_ENGINE_CARVEOUT")) : -1;
                auto launch_k = [&](auto kernel) {
                    if (carveout >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
                    kernel<<<grid, block, 0, s>>>(E);
                };
                auto launch = [&](auto cell) {
                    using CellT = decltype(cell);
                    if (prm.heuristic == 0) launch_k(msa_engine_kernel<CellT, 0>);
                    else if (prm.heuristic == 1) launch_k(msa_engine_kernel<CellT, 1>);
                    else launch_k(msa_engine_kernel<CellT, 2>);
                };
                if (cell16) launch(uint16_t{});
                else launch(uint32_t{});
            }
            CK(cudaEventRecord(c.ev1, s));
            CK(cudaGetLastError());
            uint32_t hc[4];
            CK(cudaMemcpyAsync(hc, eb.counters.d, 16, cudaMemcpyDeviceToHost, s));
            CK(wait_stream(c));
            float ms = 0;
            cudaEventElapsedTime(&ms, c.ev0, c.ev1);
            st.kernel_ms += ms;
            ++st.launches, ++st.passes;
            static const bool timing = getenv("POASTA_TIMING") != nullptr;
            if (timing)
                fprintf(stderr, "poasta_b200 engine pass %d: %zu families, %u tickets, %llu workers x %.1f MB (V <= %llu, %s cells), %.1f ms, %u suspended\n",
                        pass, todo.size(), n_tickets, (unsigned long long)workers, L.total / 1048576.0, (unsigned long long)Vw,
                        cell16 ? "16-bit" : "32-bit", ms, hc[2]);
            todo.clear();
            if (hc[2]) {
                todo.resize(hc[2]);
                CK(cudaMemcpy(todo.data(), eb.suspended.d, (size_t)hc[2] * 4, cudaMemcpyDeviceToHost));
                std::sort(todo.begin(), todo.end());
                CK(cudaMemcpy(hf.data(), eb.fams.d, (size_t)nf * sizeof(dg::Family), cudaMemcpyDeviceToHost));
                for (uint32_t f : todo) left[f] = hf[f].n_seqs - hf[f].next_seq;
                if (Vw >= Vworst && !cell16 && full_items) return POASTA_ERR_UNSUPPORTED;  // nothing left to enlarge
                Vw = std::min<uint64_t>(Vworst, Vw * 2);
                full_items = true;
            }
        }
        // ---- download: headers, then the used parts of the arenas, packed on the device ----
        const double t_chunk2 = now();
        if (!eb.h_fams.reserve((size_t)nf * sizeof(dg::Family)) || !eb.h_seq_out.reserve((size_t)n_records * SEQ_OUT_WORDS * 4)) return POASTA_ERR_CUDA;
        CK(cudaMemcpyAsync(eb.h_fams.h, eb.fams.d, (size_t)nf * sizeof(dg::Family), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(eb.h_seq_out.h, eb.seq_out.d, (size_t)n_records * SEQ_OUT_WORDS * 4, cudaMemcpyDeviceToHost, s));
        CK(wait_stream(c));
        dg::Family *df = (dg::Family *)eb.h_fams.h;
        std::vector<uint64_t> poff(nf + 1, 0);
        for (uint32_t i = 0; i < nf; ++i) {
            const dg::Caps used{df[i].n_nodes, df[i].n_edge_slots, df[i].n_sids, df[i].n_als, df[i].n_seqs_in_graph};
            poff[i + 1] = poff[i] + dg::Layout(used).total;
        }
        if (!eb.packed.reserve(poff[nf] + 16) || !eb.h_packed.reserve(poff[nf] + 16)) return POASTA_ERR_CUDA;
        CK(cudaMemcpyAsync(eb.packed_off.d, poff.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, s));
        pack_graphs_kernel<<<(nf * 32 + 127) / 128, 128, 0, s>>>((const dg::Family *)eb.fams.d, nf, (const uint8_t *)eb.arena.d,
                                                                 (const uint64_t *)eb.packed_off.d, (uint8_t *)eb.packed.d);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(eb.h_packed.h, eb.packed.d, poff[nf], cudaMemcpyDeviceToHost, s));
        CK(wait_stream(c));
        ++st.launches;
        tl_d2h += poff[nf] + (uint64_t)nf * sizeof(dg::Family) + (uint64_t)n_records * SEQ_OUT_WORDS * 4;
        // ---- back into the std::vector graphs (host threads; the reference's toposort runs here, poa.rs:360) ----
        const double t_chunk3 = now();
        const uint32_t *so = (const uint32_t *)eb.h_seq_out.h;
        std::atomic<int> err{0};
        std::atomic<uint32_t> cursor{0};
        std::vector<EngineStats> part(std::max(1u, host_threads));
        auto work = [&](unsigned tid) {
            EngineStats &ps = part[tid];
            std::vector<std::string> names;
            for (;;) {
                const uint32_t i = cursor.fetch_add(1);
                if (i >= nf) break;
                EngineFamilyIn &in = fams_in[chunk_ids[i]];
                dg::Family F = df[i];
                const uint32_t done = F.next_seq;
                for (uint32_t k = 0; k < F.n_seqs; ++k) {
                    const uint32_t *o = so + (size_t)SEQ_OUT_WORDS * (F.seq_first + k);
                    if (in.scores) in.scores[k] = o[0];
                    if (k < done) {
                        ps.cells += (uint64_t)o[5] * (*in.recs)[k].seq.size();
                        ps.queued += o[2], ps.visited += o[3], ps.pruned += o[4];
                    }
                }
                if (F.status) {
                    err = F.status;
                    continue;
                }
                F.cap = dg::Caps{F.n_nodes, F.n_edge_slots, F.n_sids, F.n_als, F.n_seqs_in_graph};
                F.arena_off = poff[i];
                names.clear();
                for (uint32_t k = 0; k < done; ++k) names.push_back((*in.recs)[k].name);
                in.graph->flat_valid = false;
                if (!dg::download(&F, (uint8_t *)eb.h_packed.h, in.graph->g, names)) err = POASTA_ERR_GRAPH_CYCLE;
            }
        };
        if (host_threads <= 1 || nf < 32) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < host_threads; ++t) pool.emplace_back(work, t);
            for (auto &t : pool) t.join();
        }
        for (const EngineStats &ps : part) st.cells += ps.cells, st.queued += ps.queued, st.visited += ps.visited, st.pruned += ps.pruned;
        t_upload += t_chunk1 - t_chunk0, t_device += t_chunk2 - t_chunk1, t_copy_back += t_chunk3 - t_chunk2, t_convert += now() - t_chunk3;
        static const bool timing2 = getenv("POASTA_TIMING") != nullptr;
        if (timing2) {  // where the kernel's time went: busy share of the worker slots, longest family chain, search share of a round
            double busy = 0, search = 0, longest = 0;
            for (uint32_t i = 0; i < nf; ++i) {
                double chain = 0;
                for (uint32_t k = 0; k < df[i].n_seqs; ++k) {
                    const uint32_t *o = so + (size_t)SEQ_OUT_WORDS * (df[i].seq_first + k);
                    chain += o[7] * 1024.0, search += o[6] * 1024.0;
                }
                busy += chain;
                longest = std::max(longest, chain);
            }
            fprintf(stderr, "poasta_b200 engine: worker-busy %.3g cycles (search %.1f %%), mean family chain %.3g, longest %.3g cycles; %.0f cycles per visited state\n",
                    busy, 100.0 * search / std::max(busy, 1.0), busy / nf, longest, busy / std::max<double>(1.0, (double)st.visited));
        }
        if (err) return err;
    }
    if (phase_timing)
        fprintf(stderr, "poasta_b200 msa phases: upload %.3f s, device %.3f s, copy back %.3f s (%.1f MB), host graphs %.3f s (%u threads)\n", t_upload,
                t_device, t_copy_back, (double)tl_d2h / 1e6, t_convert, host_threads);
    return 0;
}

}  // namespace poasta
