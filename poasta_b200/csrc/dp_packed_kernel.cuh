// dp_fill_packed_kernel — the M/I/D fill with two 16-bit scores per register (included by align.cu after dp_warp_kernel.cuh).
//
// Same mapping as dp_fill_warp_kernel (one WARP per alignment, lane l owns the 32 query columns [32 l, 32 l + 32) of every
// row, no block barriers), but a lane's 32 columns live in 16 registers of packed s16x2 scores and every recurrence is a
// Blackwell DPX / packed instruction that updates two cells at once:
//   diagonal term       PRMT (shift the previous row by one column) + IADD (packed mismatch cost from a per-symbol table)
//   deletion            IADD, LOP3 (packed select: a deletion opens only where the next query base differs from the node
//                       symbol, dfa.rs:210-250) and VIMNMX3.S16x2  = __vimin3_s16x2(open, extend, INF)
//   min(diagonal, D)    VIMNMX.S16x2    = __vmins2
//   insertion chain     I[j] = o + min_{k<j, openI(k)} (X[k] + (j - k) e): per pair IADD, LOP3, two VIMNMX.S16x2 and one PRMT for the
//                       running prefix-min inside the lane, ONE 5-step warp-shuffle prefix-min per row across lanes
//   final M             VIMNMX3.S16x2   = __vimin3_s16x2(X, I, INF)
// = 20 instructions per PAIR of cells (the s32 kernel needs ~27 per cell when it also derives the traceback bits).
//
// Traceback: the kernel writes NO per-cell decision bits. It stores the M and D rows themselves (2 x u16 per cell, fully
// coalesced 16-byte chunks: chunk t of lane l sits at (t * lanes + l) * 16 bytes of the row), and traceback_rows_kernel applies the
// reference's score-equality backtrace (gap_affine.rs:550-657, 804-915) to those tables directly — the literal rule set, so
// predecessor ordinals, "first predecessor in order wins" and the equality quirks need no encoding. The same rows are the
// row store that rows with several predecessors (or a predecessor that is not the row just before) read back. I is never
// stored: M == I is implied when neither a diagonal predecessor nor D equals M, and an insertion state always steps to
// (node, offset - 1, Match) (gap_affine.rs:639-650).
//
// Arithmetic: "infinite" is any value >= PK_INF = 0x5000; M and D are clamped to PK_INF in the same VIMNMX3 that forms them,
// so no half ever reaches 0x8000 and plain 32-bit adds never carry between the halves. Clamping is exact for every cell
// whose true value is below PK_INF (its optimal predecessors are smaller still), and the host picks this kernel only when the
// optimal score is provably below PK_INF: 2 o + e (V + Q) + x < PK_INF, e * 1024 <= PK_KMAX (align.cu: run_chunk).
#pragma once

namespace poasta {

constexpr uint32_t PK_INF = 0x5000u, PK_INF2 = 0x50005000u;   // infinite score (per half)
constexpr uint32_t PK_BIG = 0x7E00u, PK_BIG2 = 0x7E007E00u;   // "no candidate" of the insertion prefix-min (>= any candidate)
constexpr uint32_t PK_KMAX = 11000u;                          // largest (lanes * 32) * e the prefix-min offsets may reach
constexpr int PK_WARPS = 4;
constexpr int PK_TAB_WORDS = 2 * 4 * 16 * 32;                 // per warp: E and N tables, 4 symbols, 16 pairs, 32 lanes (16 KB)
constexpr int PK_TILE_ROWS = 64;                              // row records per TMA tile (1 KB), two tiles in flight per warp
constexpr int PK_WARP_SMEM = PK_TAB_WORDS * 4 + 2 * PK_TILE_ROWS * 16 + 16;  // tables + two record tiles + two mbarriers
constexpr size_t PK_SMEM_BYTES = (size_t)PK_WARPS * PK_WARP_SMEM;

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) of a tile of row records ----
__device__ __forceinline__ uint32_t pk_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pk_mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void pk_tile_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    // order this warp's earlier (generic-proxy) reads of the buffer before the async-proxy write that reuses it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pk_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pk_smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(pk_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void pk_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra PK_WAIT;\n"
        "}\n" ::"r"(pk_smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// ---- cluster mode: queries of 1024+ bases ------------------------------------------------------------------------
// A thread-block CLUSTER of up to 8 one-warp CTAs shares one alignment: CTA s owns the columns [1024 s, 1024 s + 1024). Per row the
// only values that cross a segment boundary are the predecessor row's M at the segment's last column (the diagonal term of the
// next segment's first column) and the running insertion prefix-min; CTA s writes both into a ring in CTA s + 1's shared memory
// through DSMEM (st.release.cluster on the sequence word) and CTA s + 1, working about one row behind, spins on its own shared
// memory (ld.acquire.cluster). The consumer acknowledges through a counter in the producer's shared memory, so the ring never
// overruns. Sequence numbers run on across alignments; no cluster-wide barrier is needed.
constexpr int PK_RING = 32;
struct PkMsg {
    uint32_t left;  // packed pair whose high half is the predecessor row's M at the producer's last column
    int pm;         // inclusive insertion prefix-min of all columns up to the producer's last one, in the producer's frame
    uint32_t seq;   // row sequence number + 1 once the entry is valid
    uint32_t pad;
};
constexpr int PK_CLUSTER_SMEM = PK_WARP_SMEM + PK_RING * (int)sizeof(PkMsg) + 16;
constexpr uint32_t PK_SPIN_LIMIT = 1u << 27;  // a wait that lasts this many polls aborts the launch instead of hanging the GPU

__device__ __forceinline__ void pk_st_release_cluster(uint32_t *p, uint32_t v) {
    asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t pk_ld_acquire_cluster(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cluster.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t pk_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t pk_cluster_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t pk_n_clusters() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// generic address of `p` (a shared-memory address of this CTA) in CTA `rank` of the cluster
template <typename T> __device__ __forceinline__ T *pk_map_rank(T *p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
    return reinterpret_cast<T *>(out);
}

// node symbols index the per-symbol tables; the host picks this kernel only for graphs over A, C, G, T (start / end rows carry no
// symbol: their diagonal and deletion terms are infinite or symbol-free, so any table serves)
__device__ __forceinline__ int pk_sym_index(int s) { return s == 'C' ? 1 : (s == 'G' ? 2 : (s == 'T' ? 3 : 0)); }

// Row r of an alignment's store: lanes * 64 halves = [M: 4 chunks x lanes x 8 halves][D: the same]; as uint4: 8 * lanes per row.
// CLUSTER = false: PK_WARPS independent warps per CTA, one alignment each (queries <= 1023). CLUSTER = true: one warp per CTA, the
// CTAs of a cluster are the 1024-column segments of one alignment (always WRITE_ROWS).
template <bool WRITE_ROWS, bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? 32 : PK_WARPS * 32, CLUSTER ? 12 : 3) dp_fill_packed_kernel(BatchDev P) {
    static_assert(!CLUSTER || WRITE_ROWS, "the cluster kernel keeps every row in the trace store");
    extern __shared__ uint4 pk_smem[];
    const int lane = threadIdx.x & 31, wib = CLUSTER ? 0 : threadIdx.x >> 5;
    constexpr int NW = CLUSTER ? 1 : PK_WARPS;  // warps sharing this CTA's dynamic shared memory
    const uint32_t worker = blockIdx.x * PK_WARPS + wib;
    uint4 *const tabE = pk_smem + (size_t)wib * (PK_TAB_WORDS / 4);  // [sym][chunk][lane] uint4 (4 pairs per chunk)
    uint4 *const tabN = tabE + 4 * 4 * 32;
    // the graph's row records (16 B per row: flags, predecessor summary, symbol) are staged through shared memory in tiles of
    // PK_TILE_ROWS rows by TMA bulk copies, two tiles ahead of the row being computed; lane 0 issues, an mbarrier per buffer signals
    RowRec *const tile = reinterpret_cast<RowRec *>(reinterpret_cast<uint8_t *>(pk_smem) + (size_t)NW * PK_TAB_WORDS * 4) + (size_t)wib * 2 * PK_TILE_ROWS;
    uint64_t *const bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(pk_smem) + (size_t)NW * (PK_TAB_WORDS * 4 + 2 * PK_TILE_ROWS * 16)) + wib * 2;
    // cluster mode: the incoming ring and the acknowledgement counter of the outgoing one
    PkMsg *const ring = reinterpret_cast<PkMsg *>(reinterpret_cast<uint8_t *>(pk_smem) + PK_WARP_SMEM);
    uint32_t *const acked = reinterpret_cast<uint32_t *>(ring + PK_RING);  // rows the NEXT segment has consumed (it writes this)
    const uint32_t seg = CLUSTER ? pk_cluster_ctarank() : 0u;
    uint32_t g_in = 0, g_out = 0;  // rows received from segment seg - 1 / sent to segment seg + 1 so far
    if (CLUSTER) {
        for (int i = lane; i < PK_RING; i += 32) ring[i].seq = 0;
        if (lane == 0) *acked = 0;
        __syncwarp();
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");  // rings are zero before anyone sends
    }
    if (lane == 0) {
        pk_mbar_init(&bars[0]);
        pk_mbar_init(&bars[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase0 = 0, phase1 = 0;  // parity the next completion of each buffer's mbarrier will have
    uint4 *const slot = reinterpret_cast<uint4 *>(P.scratch + (uint64_t)worker * P.scratch_stride);  // scores-only: compact row store
    const uint32_t x2 = (uint32_t)P.x * 0x10001u, oe2 = (uint32_t)P.oe * 0x10001u, e2 = (uint32_t)P.e * 0x10001u;
    const uint32_t o2 = (uint32_t)(P.oe - P.e) * 0x10001u;
    const int j0 = (int)seg * 1024 + lane * 32;  // first column of this lane
    const int vlane = (int)seg * 32 + lane;      // lane index across the segments of a row

    for (uint32_t it = 0;; ++it) {
        uint32_t a = 0;
        if (CLUSTER) {
            a = pk_cluster_id() + it * pk_n_clusters();  // static assignment: every CTA of the cluster derives the same alignment
        } else {
            if (lane == 0) a = atomicAdd(P.counter, 1u);
            a = __shfl_sync(0xFFFFFFFFu, a, 0);
        }
        if (a >= P.n_alns) break;
        if (CLUSTER && *reinterpret_cast<volatile uint32_t *>(P.counter + 1)) break;  // another cluster gave up waiting
        const AlnDesc ad = P.alns[a];
        const GraphDesc gd = P.graphs[ad.graph];
        const int Q = (int)ad.Q, V = (int)gd.V;
        const int nl = ((int)ad.W + 31) >> 5;  // lanes that own columns, over all segments
        const int n_seg = (nl + 31) >> 5;
        if (CLUSTER && (int)seg >= n_seg) continue;  // this alignment needs fewer segments than the cluster has
        const bool lane_active = vlane < nl;
        const bool recv = CLUSTER && seg > 0, send = CLUSTER && (int)seg + 1 < n_seg;
        const uint8_t *q = P.queries + ad.query_base;
        const RowRec *rows = P.rows + gd.node_base;
        const uint32_t *prank = P.pred_rank + gd.pred_base, *pslot = P.pred_slot + gd.pred_base;
        uint4 *const store = WRITE_ROWS ? reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(P.trace) + ad.prow_base) : slot;
        const int row_u4 = 8 * nl;

        // per-symbol tables of this lane's columns
        __syncwarp();
        // Packed per-symbol tables of this lane's columns (pair k = columns j0 + 2k, j0 + 2k + 1; 0xFFFF per half where it holds):
        //   E  column j matches s:            1 <= j <= Q && q[j - 1] == s
        //   N  the next query base equals s:  j < Q && q[j] == s
        {
            uint32_t *wE = reinterpret_cast<uint32_t *>(tabE), *wN = reinterpret_cast<uint32_t *>(tabN);
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {
                const int j = j0 + 2 * k;
                const int qm = (j >= 1 && j <= Q) ? q[j - 1] : 256, q0 = j < Q ? q[j] : 256, q1 = j + 1 < Q ? q[j + 1] : 256;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int sym = s == 0 ? 'A' : (s == 1 ? 'C' : (s == 2 ? 'G' : 'T'));
                    const uint32_t E = (qm == sym ? 0xFFFFu : 0u) | (q0 == sym ? 0xFFFF0000u : 0u);  // column j + 1 <= Q matches q[j]
                    const uint32_t N = (q0 == sym ? 0xFFFFu : 0u) | (q1 == sym ? 0xFFFF0000u : 0u);
                    const int w = ((s * 4 + (k >> 2)) * 32 + lane) * 4 + (k & 3);
                    wE[w] = E, wN[w] = N;
                }
            }
        }
        __syncwarp();
        uint32_t K[16];  // (Wc - j) e per half, Wc = 32 nl: the offsets of the insertion prefix-min (always >= e)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // single warp: Wc = 32 nl; cluster: every segment uses its own frame Wc = 1024 (the carry is shifted by 1024 e per hop)
            const uint32_t lo = (uint32_t)(CLUSTER ? 1024 - (lane * 32 + 2 * k) : 32 * nl - (j0 + 2 * k)) * (uint32_t)P.e, hi = lo - (uint32_t)P.e;
            K[k] = lane_active ? (lo | (hi << 16)) : 0u;
        }

        uint32_t Pm[16], Pd[16];  // the previous row (rank r - 1), then the gathered predecessor row, then this row's M and D
#pragma unroll
        for (int k = 0; k < 16; ++k) Pm[k] = Pd[k] = PK_INF2;

        const int n_tiles = (V + PK_TILE_ROWS - 1) / PK_TILE_ROWS;
        auto issue_tile = [&](int t) {  // lane 0 only
            const int first = t * PK_TILE_ROWS, cnt = min(PK_TILE_ROWS, V - first);
            pk_tile_load(tile + (t & 1) * PK_TILE_ROWS, rows + first, (uint32_t)cnt * 16u, &bars[t & 1]);
        };
        if (lane == 0) {
            issue_tile(0);
            if (n_tiles > 1) issue_tile(1);
        }
        for (int r = 0; r < V; ++r) {
            if ((r & (PK_TILE_ROWS - 1)) == 0) {  // entering a tile: wait for its copy, refill the buffer the previous tile left
                const int t = r / PK_TILE_ROWS;
                if (t >= 1) {
                    __syncwarp();  // every lane is done with tile t - 1
                    if (lane == 0 && t + 1 < n_tiles) issue_tile(t + 1);
                }
                if (t & 1) pk_mbar_wait(&bars[1], phase1), phase1 ^= 1;
                else pk_mbar_wait(&bars[0], phase0), phase0 ^= 1;
            }
            const RowRec row = tile[((r / PK_TILE_ROWS) & 1) * PK_TILE_ROWS + (r & (PK_TILE_ROWS - 1))];
            const uint32_t inf = row.info;
            const int vs = (int)(row.npred_sym >> 24);
            const bool is_end = inf & INFO_IS_END, is_start = inf & INFO_IS_START;
            const int npred = (int)(row.npred_sym & 0xFFFFFFu);

            // ---- gather: elementwise minimum of the predecessor rows (min commutes with the column shift and the costs) ----
            if (!(npred == 1 && row.pred0_slot == ROW_PREV_IN_REGS)) {
                uint32_t Am[16], Ad[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) Am[k] = Ad[k] = PK_INF2;
                const uint32_t pb = row.pred_begin, pe = pb + (uint32_t)npred;
                for (uint32_t kk = pb; kk < pe; ++kk) {
                    const uint32_t pr = prank[kk];
                    if (!WRITE_ROWS && (int)pr == r - 1) {  // scores-only: the row just computed may have no slot in the compact store
#pragma unroll
                        for (int k = 0; k < 16; ++k) Am[k] = __vmins2(Am[k], Pm[k]), Ad[k] = __vmins2(Ad[k], Pd[k]);
                    } else if (lane_active) {
                        const uint4 *src = store + (uint64_t)(WRITE_ROWS ? pr : pslot[kk]) * row_u4 + vlane;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const uint4 m = src[t * nl], d = src[(4 + t) * nl];
                            Am[4 * t] = __vmins2(Am[4 * t], m.x), Am[4 * t + 1] = __vmins2(Am[4 * t + 1], m.y);
                            Am[4 * t + 2] = __vmins2(Am[4 * t + 2], m.z), Am[4 * t + 3] = __vmins2(Am[4 * t + 3], m.w);
                            Ad[4 * t] = __vmins2(Ad[4 * t], d.x), Ad[4 * t + 1] = __vmins2(Ad[4 * t + 1], d.y);
                            Ad[4 * t + 2] = __vmins2(Ad[4 * t + 2], d.z), Ad[4 * t + 3] = __vmins2(Ad[4 * t + 3], d.w);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) Pm[k] = Am[k], Pd[k] = Ad[k];
            }

            // ---- diagonal, deletion, X = min(diagonal, D) ----
            uint32_t left = __shfl_up_sync(0xFFFFFFFFu, Pm[15], 1);  // its high half is column j0 - 1
            if (lane == 0) left = PK_INF2;
            const uint32_t left_send = Pm[15];  // lane 31: the next segment's `left`
            int pm_in = (int)PK_BIG;
            bool abort_wait = false;
            if (recv) {  // the row's message from segment seg - 1 (it runs about one row ahead)
                if (lane == 0) {
                    PkMsg *m = &ring[g_in % PK_RING];
                    uint32_t polls = 0;
                    while (pk_ld_acquire_cluster(&m->seq) != g_in + 1)
                        if (++polls > PK_SPIN_LIMIT || (!(polls & 0xFFFF) && *reinterpret_cast<volatile uint32_t *>(P.counter + 1))) {
                            abort_wait = true;
                            break;
                        }
                    left = m->left;
                    pm_in = m->pm;
                    pk_st_release_cluster(pk_map_rank(acked, seg - 1), g_in + 1);  // the entry may be reused
                }
                pm_in = __shfl_sync(0xFFFFFFFFu, pm_in, 0);
                // the producer's frame ends 1024 columns to the left of mine
                pm_in = min(pm_in + 1024 * P.e, (int)PK_BIG);
                ++g_in;
            }
            uint32_t X[16];
            const int si = pk_sym_index(vs);
            if (is_end) {
                // M(end, j) = min over predecessors of M(p, j): same column, no symbol; no deletion is opened into `end`
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t d = __vimin3_s16x2(Pd[k] + e2, PK_INF2, PK_INF2);
                    X[k] = __vmins2(Pm[k], d);
                    Pd[k] = d;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint4 e4 = tabE[(si * 4 + t) * 32 + lane], n4 = tabN[(si * 4 + t) * 32 + lane];
                    const uint32_t E[4] = {e4.x, e4.y, e4.z, e4.w}, N[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int k = 4 * t + p;
                        const uint32_t prev = k == 0 ? left : Pm[k - 1];
                        const uint32_t diag = __byte_perm(prev, Pm[k], 0x5432);           // (M[j - 1], M[j]) of the predecessor row
                        const uint32_t mdiag = diag + (~E[p] & x2);                         // + mismatch cost where q[j - 1] != symbol
                        const uint32_t dopen = ((Pm[k] + oe2) & ~N[p]) | (PK_INF2 & N[p]);  // open only where q[j] != symbol
                        const uint32_t d = __vimin3_s16x2(dopen, Pd[k] + e2, PK_INF2);
                        X[k] = __vmins2(mdiag, d);
                        Pd[k] = d;
                    }
                }
            }
            if (is_start && vlane == 0) X[0] &= 0xFFFF0000u;  // M(start, 0) = 0

            // ---- insertion chain: prefix-min over the whole row ----
            // an insertion opens at column k only if the node has a successor that is `end` or mismatches q[k]
            const bool any_open = (inf & INFO_HAS_SUCC) && !is_end;
            const bool always_open = inf & INFO_ALWAYS_OPEN;
            const int ui = pk_sym_index((int)(inf & 0xFF));
            uint32_t FI[16];  // 0xFFFF per half where no insertion may open
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint4 n4 = make_uint4(0u, 0u, 0u, 0u);
                if (!any_open) {
                    n4 = make_uint4(~0u, ~0u, ~0u, ~0u);
                } else if (!always_open) {
                    n4 = tabN[(ui * 4 + t) * 32 + lane];
                }
                FI[4 * t] = n4.x, FI[4 * t + 1] = n4.y, FI[4 * t + 2] = n4.z, FI[4 * t + 3] = n4.w;
            }
            uint32_t rm = PK_BIG2;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t av = ((X[k] + K[k]) & ~FI[k]) | (PK_BIG2 & FI[k]);
                rm = __vmins2(rm, av);
            }
            int run = (int)min(rm & 0xFFFFu, rm >> 16);
            if (!lane_active) run = (int)PK_BIG;
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl = min(incl, o);
            }
            int pm = __shfl_up_sync(0xFFFFFFFFu, incl, 1);  // exclusive prefix-min over all columns left of mine
            if (lane == 0) pm = (int)PK_BIG;
            if (CLUSTER) pm = min(pm, pm_in);
            if (send) {
                if (lane == 31) {
                    uint32_t polls = 0;
                    while (g_out - pk_ld_acquire_cluster(acked) >= (uint32_t)PK_RING)  // the consumer is a whole ring behind
                        if (++polls > PK_SPIN_LIMIT || (!(polls & 0xFFFF) && *reinterpret_cast<volatile uint32_t *>(P.counter + 1))) {
                            abort_wait = true;
                            break;
                        }
                    PkMsg *m = pk_map_rank(&ring[g_out % PK_RING], seg + 1);
                    m->left = left_send;
                    m->pm = min(incl, pm_in);
                    pk_st_release_cluster(&m->seq, g_out + 1);
                }
                ++g_out;
            }
            if (CLUSTER && __any_sync(0xFFFFFFFFu, abort_wait)) {
                if (lane == 0) {
                    atomicExch(P.counter + 1, 1u);
                    P.scores[a] = (uint32_t)INF;
                    P.status[a] = POASTA_ERR_INTERNAL;
                }
                break;
            }
            uint32_t pmv = (uint32_t)pm * 0x10001u;

            // ---- final M ----
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t av = ((X[k] + K[k]) & ~FI[k]) | (PK_BIG2 & FI[k]);
                const uint32_t pre = __vmins2(pmv, __byte_perm(av, PK_BIG2, 0x1054));  // (pm, min(pm, av.lo)): exclusive per column
                const uint32_t iv = (pre - K[k]) + o2;                                  // I[j] = pre + o - (Wc - j) e
                Pm[k] = __vimin3_s16x2(X[k], iv, PK_INF2);
                const uint32_t t2 = __vmins2(pre, av);
                pmv = __byte_perm(t2, t2, 0x3232);                                      // min(pm, av.lo, av.hi) in both halves
            }

            // ---- store the row ----
            if (lane_active && (WRITE_ROWS || (inf & INFO_STORE))) {
                uint4 *dst = store + (uint64_t)(WRITE_ROWS ? (uint32_t)r : (inf >> INFO_SLOT_SHIFT)) * row_u4 + vlane;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    dst[t * nl] = make_uint4(Pm[4 * t], Pm[4 * t + 1], Pm[4 * t + 2], Pm[4 * t + 3]);
                    dst[(4 + t) * nl] = make_uint4(Pd[4 * t], Pd[4 * t + 1], Pd[4 * t + 2], Pd[4 * t + 3]);
                }
            }
            if (is_end && j0 <= Q && Q < j0 + 32) {
                uint32_t sc = PK_INF;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (j0 + 2 * k == Q) sc = Pm[k] & 0xFFFFu;
                    if (j0 + 2 * k + 1 == Q) sc = Pm[k] >> 16;
                }
                P.scores[a] = sc < PK_INF ? sc : (uint32_t)INF;
            }
        }
        __syncwarp();
    }
    // no CTA may leave while a neighbour can still write into its shared memory (acknowledgements, ring entries)
    if (CLUSTER) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// gap_affine.rs:550-657 + 804-915 on the M / D rows written by dp_fill_packed_kernel<true>. One WARP per alignment.
// Most of a walk is a run of diagonal steps along a chain of rows whose only predecessor is the row before (a match or a
// mismatch: M(r - 1, o - 1) == M(r, o) [- x]). Lane i tests step i of such a run from the current cell — its loads depend on
// (r - i, o - i) only — and one ballot gives the length of the run (up to 32 steps per round trip to memory, pairs written
// coalesced). Where the run stops (a gap, a row with several predecessors, the start or end row) lane 0 takes ONE step with the
// reference's full rule set and the lanes try again from there.
__global__ void traceback_rows_kernel(BatchDev P) {
    const uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= P.n_alns) return;
    const AlnDesc ad = P.alns[a];
    const GraphDesc gd = P.graphs[ad.graph];
    const int Q = (int)ad.Q, V = (int)gd.V;
    const int nl = ((int)ad.W + 31) >> 5;
    const uint32_t *node_id = P.node_id + gd.node_base;
    const uint8_t *sym = P.node_sym + gd.node_base;
    const RowRec *recs = P.rows + gd.node_base;
    const uint32_t *poff = P.pred_off + gd.predoff_base;
    const uint8_t *q = P.queries + ad.query_base;
    const uint16_t *rows = reinterpret_cast<const uint16_t *>(P.trace) + ad.prow_base;
    const int x = P.x, oe = P.oe, e = P.e;
    uint32_t *out_r = P.out_r + ad.out_base, *out_q = P.out_q + ad.out_base;
    uint32_t pos = ad.out_cap;  // fill from the back
    int32_t status = 0;
    const uint32_t score = P.scores[a];
    // score of (rank r, column j); plane 0 = M, 1 = D; -1 = unvisited
    auto val = [&](int r, int j, int plane) -> int {
        const int l = j >> 5, c = j & 31;
        const uint32_t v = __ldg(&rows[(uint64_t)r * (nl * 64) + (uint32_t)plane * (nl * 32) + (uint32_t)(((c >> 3) * nl + l) * 8 + (c & 7))]);
        return v >= PK_INF ? -1 : (int)v;
    };
    if (score >= (uint32_t)INF) {
        status = POASTA_ERR_NO_ALIGNMENT;
    } else if (Q == 1) {  // gap_affine.rs:812-824
        --pos;
        if (lane == 0) out_r[pos] = node_id[V - 1], out_q[pos] = 0;
    } else if (Q > 1) {
        auto step = [&](int r, int o, int s, int &nr, int &no, int &ns) -> bool {
            const uint32_t pb = poff[r], pe = poff[r + 1];
            if (s == 0) {
                const int cur = val(r, o, 0);
                if (cur < 0) return false;
                if (o > 0) {  // predecessors first, in the reference's Incoming order
                    const bool end_row = r == V - 1;
                    const bool me = end_row || sym[r] == q[o - 1];
                    const int po = end_row ? o : o - 1, want = me ? cur : cur - x;
                    if (want >= 0)
                        for (uint32_t k = pb; k < pe; ++k) {
                            const int p = (int)P.pred_rank[k];
                            if (val(p, po, 0) == want) {
                                nr = p, no = po, ns = 0;
                                return true;
                            }
                        }
                }
                if (val(r, o, 1) == cur) {
                    nr = r, no = o, ns = 1;
                    return true;
                }
                // M is the minimum of the diagonal candidates, D and I: neither of the first two equals it, so I does
                if (o > 0 && r != V - 1) {
                    nr = r, no = o, ns = 2;
                    return true;
                }
                return false;
            } else if (s == 1) {
                const int cur = val(r, o, 1);
                if (cur < 0) return false;
                const int w1 = cur - oe, w2 = cur - e;
                if (w1 >= 0)
                    for (uint32_t k = pb; k < pe; ++k) {
                        const int p = (int)P.pred_rank[k];
                        if (val(p, o, 0) == w1) {
                            nr = p, no = o, ns = 0;
                            return true;
                        }
                    }
                if (w2 >= 0)
                    for (uint32_t k = pb; k < pe; ++k) {
                        const int p = (int)P.pred_rank[k];
                        if (val(p, o, 1) == w2) {
                            nr = p, no = o, ns = 1;
                            return true;
                        }
                    }
                return false;
            } else {
                if (o <= 0) return false;
                nr = r, no = o - 1, ns = 0;  // open and extend both resume in Match state (gap_affine.rs:639-650)
                return true;
            }
        };
        int cr = V - 1, co = Q, cs = 0, br = 0, bo = 0, bs = 0;
        int going = 0;
        if (lane == 0 && step(cr, co, cs, br, bo, bs)) cr = br, co = bo, cs = bs, going = 1;
        going = __shfl_sync(0xFFFFFFFFu, going, 0);
        cr = __shfl_sync(0xFFFFFFFFu, cr, 0), co = __shfl_sync(0xFFFFFFFFu, co, 0), cs = __shfl_sync(0xFFFFFFFFu, cs, 0);
        while (going) {
            if (cs == 0) {
                // lane i: is step i of a diagonal run from (cr, co) the step the reference takes at (cr - i, co - i)?
                const int ri = cr - lane, oi = co - lane;
                bool ok = ri >= 2 && oi >= 1 && ri != V - 1;  // the predecessor ri - 1 is not the start row (rank 0)
                if (ok) {
                    const uint4 rv = __ldg(reinterpret_cast<const uint4 *>(recs + ri));
                    const RowRec rec{rv.x, rv.y, rv.z, rv.w};
                    ok = (rec.npred_sym & 0xFFFFFFu) == 1u && rec.pred0_slot == ROW_PREV_IN_REGS;
                    if (ok) {
                        const int cur = val(ri, oi, 0), prev = val(ri - 1, oi - 1, 0);
                        const bool me = (int)(rec.npred_sym >> 24) == (int)q[oi - 1];
                        const int want = me ? cur : cur - x;
                        ok = cur >= 0 && want >= 0 && prev == want;
                    }
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
                const int n = __ffs((int)~m) - 1;  // leading run of successes (32 when every lane succeeded: ~m == 0 -> ffs 0 -> -1)
                const int run = n < 0 ? 32 : n;
                if (run > 0) {
                    if (pos < (uint32_t)run) {
                        status = POASTA_ERR_UNSUPPORTED;
                        break;
                    }
                    if (lane < run) out_r[pos - 1 - lane] = node_id[ri], out_q[pos - 1 - lane] = (uint32_t)(oi - 1);
                    pos -= (uint32_t)run, cr -= run, co -= run;
                    continue;
                }
            }
            // one step with the full rule set (lane 0), exactly the reference's loop body
            int stop = 0;
            if (lane == 0) {
                if (!step(cr, co, cs, br, bo, bs)) {
                    stop = 1;
                } else if (cs == 0 && bs != 0) {
                    cr = br, co = bo, cs = bs;
                } else if (pos == 0) {
                    status = POASTA_ERR_UNSUPPORTED;
                    stop = 1;
                } else {
                    --pos;
                    if (cs == 0) out_r[pos] = node_id[cr], out_q[pos] = (uint32_t)(co - 1);
                    else if (cs == 2) out_r[pos] = POASTA_NONE, out_q[pos] = (uint32_t)(co - 1);
                    else out_r[pos] = node_id[cr], out_q[pos] = POASTA_NONE;
                    if (br == 0) stop = 1;  // predecessor is the start node (rank 0)
                    cr = br, co = bo, cs = bs;
                }
            }
            stop = __shfl_sync(0xFFFFFFFFu, stop, 0);
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
            status = __shfl_sync(0xFFFFFFFFu, status, 0);
            cr = __shfl_sync(0xFFFFFFFFu, cr, 0), co = __shfl_sync(0xFFFFFFFFu, co, 0), cs = __shfl_sync(0xFFFFFFFFu, cs, 0);
            if (stop) break;
        }
    }
    if (lane == 0) {
        P.n_pairs[a] = ad.out_cap - pos;
        P.status[a] = status;
    }
}

}  // namespace poasta
