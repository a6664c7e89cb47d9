// dp_fill_warp_kernel — the M/I/D fill for queries up to 1023 bases (included by align.cu after BatchDev).
//
// One WARP per alignment, no block barriers. Lane l owns the 32 contiguous query columns [32 l, 32 l + 32)
// of every row. The previous row's M and D for those columns live in REGISTERS, so a row whose only
// predecessor is the row just computed (the common case in a POA graph laid out in topological order)
// touches no memory except its traceback bits:
//   * diagonal / deletion terms straight from registers, the left neighbour's column via one shuffle;
//   * symbol tests from per-lane 32-bit equality masks of the query (one pair per symbol and alignment);
//   * the in-row insertion chain I[j] = o + j e + min_{k<j, openI(k)} (min(Mdiag, D)[k] - k e) as a sequential
//     prefix-min over the lane's own columns plus ONE 5-step warp-shuffle prefix-min per row;
//   * add / min recurrences in plain 32-bit arithmetic (ptxas fuses them into VIADDMNMX / VIMNMX3; the packed 16-bit kernel in
//     dp_packed_kernel.cuh, which replaces this one whenever scores fit, calls the DPX intrinsics directly);
//   * "infinite" is any value >= 2^28 (INF = 2^29): sums never overflow, so there is no clamping in the inner loops.
// Traceback is stored as BIT PLANES: per (row, lane) eight 32-bit words
//     [M<-diag, M<-D, M<-I, D<-open, D<-extend, I reachable, -, -]   (bit c = column 32 lane + c)
// = 1 byte per cell, written with two 128-bit stores per lane and row; predecessor ordinals exist only for rows
// with several predecessors (side store, 4 KB per such row). The meaning of the bits is the reference's
// score-equality backtrace (gap_affine.rs:550-657), as in dp_fill_kernel.
// Rows with several predecessors, or whose predecessor is not the previous row, read predecessor rows from a
// per-warp row store in HBM; only rows that some later row needs from there are written to it (host flag).
// Every row starts from ONE 128-bit record (RowRec) fetched a row ahead, so a plain row issues no dependent loads at all.
// Measured and rejected: prefetch.global.L1 of the predecessor rows the next row will read (predecessors are a median of
// 2 rows back, i.e. already cache-resident) made the fill 5 % slower (profiles/README.md).
#pragma once

namespace poasta {

constexpr int WARP_COLS = 32;      // columns per lane
constexpr int WARP_MAX_W = 1024;   // 32 lanes x 32 columns
constexpr int WARP_KERNEL_WARPS = 4;
constexpr int WINF = 0x20000000;   // "infinite" score of the warp kernel
constexpr int WFIN = 0x10000000;   // values below this are finite
enum { PL_M_DIAG = 0, PL_M_DEL = 1, PL_M_INS = 2, PL_D_OPEN = 3, PL_D_EXT = 4, PL_I = 5, PL_WORDS = 8 };

// per-lane equality masks of the query against one symbol:
//   m bit c = (1 <= j <= Q && q[j-1] == s)   "column j matches s"      (j = 32 lane + c)
//   n bit c = (j < Q && q[j] == s)           "the next query base equals s"
struct SymMask {
    uint32_t m, n;
};

__device__ __forceinline__ SymMask make_mask(const uint8_t *q, int Q, int j0, int s) {
    SymMask r{0u, 0u};
#pragma unroll 8
    for (int c = 0; c < WARP_COLS; ++c) {
        const int j = j0 + c;
        if (j >= 1 && j <= Q && q[j - 1] == s) r.m |= 1u << c;
        if (j < Q && q[j] == s) r.n |= 1u << c;
    }
    return r;
}

template <bool WRITE_TRACE>
__global__ void __launch_bounds__(WARP_KERNEL_WARPS * 32, 3) dp_fill_warp_kernel(BatchDev P) {
    constexpr int C = WARP_COLS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t worker = blockIdx.x * WARP_KERNEL_WARPS + wib;
    int *const slot = P.scratch + (uint64_t)worker * P.scratch_stride;
    const int x = P.x, oe = P.oe, e = P.e;
    const int j0 = lane * C;

    for (;;) {
        uint32_t a = 0;
        if (lane == 0) a = atomicAdd(P.counter, 1u);
        a = __shfl_sync(0xFFFFFFFFu, a, 0);
        if (a >= P.n_alns) break;
        const AlnDesc ad = P.alns[a];
        const GraphDesc gd = P.graphs[ad.graph];
        const int Q = (int)ad.Q, W = (int)ad.W, V = (int)gd.V;
        const uint8_t *q = P.queries + ad.query_base;
        const RowRec *rows = P.rows + gd.node_base;
        const uint32_t *ord_slot = P.ord_slot + gd.node_base;
        const uint32_t *prank = P.pred_rank + gd.pred_base, *pslot = P.pred_slot + gd.pred_base;
        uint32_t *planes = (uint32_t *)P.trace + ad.ptrace_base;           // V * 32 lanes * PL_WORDS words
        uint32_t *ord_store = planes + (uint64_t)V * 32 * PL_WORDS;         // n_ord_rows * 1024 words, [c][lane] per row

        const SymMask mA = make_mask(q, Q, j0, 'A'), mC = make_mask(q, Q, j0, 'C'), mG = make_mask(q, Q, j0, 'G'),
                      mT = make_mask(q, Q, j0, 'T');
        auto mask_of = [&](int s) -> SymMask {
            if (s == 'A') return mA;
            if (s == 'C') return mC;
            if (s == 'G') return mG;
            if (s == 'T') return mT;
            return make_mask(q, Q, j0, s);
        };
        uint32_t valid = 0, ltQ = 0, ge1 = 0;  // bit c: j <= Q, j < Q, j >= 1
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (j0 + c <= Q) valid |= 1u << c;
            if (j0 + c < Q) ltQ |= 1u << c;
            if (j0 + c >= 1) ge1 |= 1u << c;
        }
        const bool lane_active = j0 < W;
        const int Qj = Q - j0;  // (Q - j) = Qj - c

        int Mp[C], Dp[C];  // previous row (rank r-1), final values
#pragma unroll
        for (int c = 0; c < C; ++c) Mp[c] = WINF, Dp[c] = WINF;

        RowRec next_row = rows[0];
        for (int r = 0; r < V; ++r) {
            const RowRec row = next_row;
            if (r + 1 < V) next_row = rows[r + 1];  // in flight while this row is computed
            const uint32_t inf = row.info;
            const int vs = (int)(row.npred_sym >> 24);
            const bool is_end = inf & INFO_IS_END, is_start = inf & INFO_IS_START;
            const int npred = (int)(row.npred_sym & 0xFFFFFFu);
            const uint32_t pb = row.pred_begin, pe = pb + (uint32_t)npred;
            const SymMask ms = mask_of(vs);
            const uint32_t openD_mask = is_end ? 0u : (valid & ~ms.n);
            uint32_t openI_mask = 0;
            if ((inf & INFO_HAS_SUCC) && !is_end) openI_mask = (inf & INFO_ALWAYS_OPEN) ? ltQ : (ltQ & ~mask_of((int)(inf & 0xFF)).n);
            int Xr[C];                         // min(Mdiag, D) per column
            uint32_t diag_le = 0, d_open = 0;  // bit c: Mdiag <= D ; D equals the best "open" candidate
            const uint32_t oslot = WRITE_TRACE ? ord_slot[r] : 0xFFFFFFFFu;
            int prev_left = __shfl_up_sync(0xFFFFFFFFu, Mp[C - 1], 1);  // M[r-1][j0 - 1] (register copy of the previous row)
            if (lane == 0) prev_left = WINF;

            if (npred == 1 && !is_end) {
                // ---- single predecessor: registers (pred == r-1) or one reload from the row store ----
                const bool reload = row.pred0_slot != ROW_PREV_IN_REGS;
                if (reload) {
                    const int *Ms = slot + (uint64_t)(2 * row.pred0_slot) * W;
                    if (lane_active) {
#pragma unroll
                        for (int k = 0; k < C / 4; ++k) {
                            if (j0 + 4 * k < W) {
                                const int4 m4 = *reinterpret_cast<const int4 *>(Ms + j0 + 4 * k);
                                const int4 d4 = *reinterpret_cast<const int4 *>(Ms + W + j0 + 4 * k);
                                Mp[4 * k] = m4.x, Mp[4 * k + 1] = m4.y, Mp[4 * k + 2] = m4.z, Mp[4 * k + 3] = m4.w;
                                Dp[4 * k] = d4.x, Dp[4 * k + 1] = d4.y, Dp[4 * k + 2] = d4.z, Dp[4 * k + 3] = d4.w;
                            } else {
                                Mp[4 * k] = Mp[4 * k + 1] = Mp[4 * k + 2] = Mp[4 * k + 3] = WINF;
                                Dp[4 * k] = Dp[4 * k + 1] = Dp[4 * k + 2] = Dp[4 * k + 3] = WINF;
                            }
                        }
                    }
                }
                int prevM = prev_left;
                if (reload) {  // reloaded row: fetch the left neighbour's last column again
                    prevM = __shfl_up_sync(0xFFFFFFFFu, Mp[C - 1], 1);
                    if (lane == 0) prevM = WINF;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int mpc = Mp[c];
                    const int mdiag = prevM + (((ms.m >> c) & 1) ? 0 : x);
                    const int dopen = mpc + oe;
                    const int dext = Dp[c] + e;
                    const int d = ((openD_mask >> c) & 1) ? min(dopen, dext) : dext;
                    if (dopen == d) d_open |= 1u << c;
                    if (mdiag <= d) diag_le |= 1u << c;
                    Xr[c] = min(mdiag, d);
                    Dp[c] = d;
                    prevM = mpc;
                }
            } else {
                // ---- start row, end row, several predecessors: every predecessor row comes from the row store ----
#pragma unroll
                for (int cc = 0; cc < C / 8; ++cc) {
                    int Xa[8], dom[8], dem[8];
                    int dop01[2][8];   // the "open" candidates of the first two predecessors (see the equality test below)
                    uint32_t ords[8];  // mord | dopen_ord << 10 | dext_ord << 20
#pragma unroll
                    for (int c = 0; c < 8; ++c) Xa[c] = WINF, dom[c] = WINF, dem[c] = WINF, ords[c] = 0, dop01[0][c] = dop01[1][c] = WINF;
                    const int jb = j0 + cc * 8;
                    if (jb < W) {
                        for (uint32_t k = pb; k < pe; ++k) {
                            const uint32_t ord = k - pb;
                            int mp[9], dp[8];
                            if ((int)prank[k] == r - 1) {  // the row just computed is still in registers
                                mp[0] = cc == 0 ? prev_left : Mp[cc * 8 - (cc == 0 ? 0 : 1)];
#pragma unroll
                                for (int c = 0; c < 8; ++c) mp[c + 1] = Mp[cc * 8 + c], dp[c] = Dp[cc * 8 + c];
                            } else {
                                const int *Ms = slot + (uint64_t)(2 * pslot[k]) * W;
                                const int4 ma = *reinterpret_cast<const int4 *>(Ms + jb), mb = *reinterpret_cast<const int4 *>(Ms + jb + 4);
                                const int4 da = *reinterpret_cast<const int4 *>(Ms + W + jb), db = *reinterpret_cast<const int4 *>(Ms + W + jb + 4);
                                mp[0] = jb > 0 ? Ms[jb - 1] : WINF;
                                mp[1] = ma.x, mp[2] = ma.y, mp[3] = ma.z, mp[4] = ma.w, mp[5] = mb.x, mp[6] = mb.y, mp[7] = mb.z, mp[8] = mb.w;
                                dp[0] = da.x, dp[1] = da.y, dp[2] = da.z, dp[3] = da.w, dp[4] = db.x, dp[5] = db.y, dp[6] = db.z, dp[7] = db.w;
                            }
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int cg = cc * 8 + c;
                                const int cand = is_end ? mp[c + 1] : mp[c] + (((ms.m >> cg) & 1) ? 0 : x);
                                if (cand < Xa[c]) Xa[c] = cand, ords[c] = (ords[c] & ~0x3FFu) | ord;
                                const int dop = mp[c + 1] + oe, dex = dp[c] + e;
                                if (WRITE_TRACE && ord == 0) dop01[0][c] = dop;
                                if (WRITE_TRACE && ord == 1) dop01[1][c] = dop;
                                if (dop < dom[c]) dom[c] = dop, ords[c] = (ords[c] & ~(0x3FFu << 10)) | (ord << 10);
                                if (dex < dem[c]) dem[c] = dex, ords[c] = (ords[c] & ~(0x3FFu << 20)) | (ord << 20);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int cg = cc * 8 + c, j = jb + c;
                        const int d = ((openD_mask >> cg) & 1) ? min(dom[c], dem[c]) : dem[c];
                        int xd = Xa[c];
                        if (!is_end && j == 0) xd = WINF;  // no diagonal into column 0
                        if (is_start && j == 0) xd = 0;
                        bool is_open = dom[c] == d;
                        if (WRITE_TRACE && d < WFIN && !is_open && !((openD_mask >> cg) & 1) && npred > 1 && dom[c] < d && j <= Q) {
                            // a non-minimal "open" candidate equals D although opening is not allowed here (the reference's
                            // backtrace tests score equality only; first predecessor in order wins). About one column in eight
                            // of a merge row gets here, so the first two candidates are kept in registers.
                            if (dop01[0][c] == d) {
                                is_open = true;
                                ords[c] = ords[c] & ~(0x3FFu << 10);
                            } else if (dop01[1][c] == d) {
                                is_open = true;
                                ords[c] = (ords[c] & ~(0x3FFu << 10)) | (1u << 10);
                            } else
                            for (uint32_t k = pb + 2; k < pe; ++k) {
                                const int mpj = (int)prank[k] == r - 1 ? Mp[cg] : slot[(uint64_t)(2 * pslot[k]) * W + j];
                                if (mpj + oe == d) {
                                    is_open = true;
                                    ords[c] = (ords[c] & ~(0x3FFu << 10)) | ((k - pb) << 10);
                                    break;
                                }
                            }
                        }
                        if (is_open) d_open |= 1u << cg;
                        if (xd <= d && !(is_start && j == 0)) diag_le |= 1u << cg;
                        Xr[cg] = min(xd, d);
                        Dp[cg] = d;
                        if (WRITE_TRACE && oslot != 0xFFFFFFFFu && lane_active) ord_store[(uint64_t)oslot * 1024 + cg * 32 + lane] = ords[c];
                    }
                }
            }

            // ---- insertion chain: prefix-min over the whole row ----
            int run = 2 * WINF - 1;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int av = ((openI_mask >> c) & 1) ? Xr[c] + (Qj - c) * e : 2 * WINF - 1;
                run = min(run, av);
            }
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl = min(incl, o);
            }
            int pm = __shfl_up_sync(0xFFFFFFFFu, incl, 1);  // exclusive prefix-min over all columns left of mine
            if (lane == 0) pm = 2 * WINF - 1;

            // ---- final M and the traceback bit planes ----
            uint32_t x_le_i = 0, i_le_x = 0, m_fin = 0, i_fin = 0, d_fin = 0, d_eq_m = 0;
            const int ioff = oe - e;  // I[j] = pm + o - (Q - j) e
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int kje = (Qj - c) * e;
                const int iv = pm + ioff - kje;
                const int xr = Xr[c], dv = Dp[c];
                const int mv = min(xr, iv);
                const int av = ((openI_mask >> c) & 1) ? xr + kje : 2 * WINF - 1;
                pm = min(pm, av);
                Mp[c] = mv;
                if (WRITE_TRACE) {
                    if (xr <= iv) x_le_i |= 1u << c;
                    if (iv <= xr) i_le_x |= 1u << c;
                    if (mv < WFIN) m_fin |= 1u << c;
                    if (iv < WFIN) i_fin |= 1u << c;
                    if (dv < WFIN) d_fin |= 1u << c;
                    if (dv == mv) d_eq_m |= 1u << c;
                }
            }
            if (WRITE_TRACE && lane_active) {
                // gap_affine.rs:577-613: predecessors first (offset > 0), then close-deletion, then close-insertion
                const uint32_t m_diag = m_fin & x_le_i & diag_le & ge1 & (npred > 0 ? ~0u : 0u);
                const uint32_t m_del = m_fin & ~m_diag & d_eq_m;
                const uint32_t m_ins = m_fin & ~m_diag & ~m_del & i_le_x;
                const uint32_t pd_open = d_fin & d_open, pd_ext = d_fin & ~d_open;
                uint4 *tp = reinterpret_cast<uint4 *>(planes + ((uint64_t)r * 32 + lane) * PL_WORDS);
                tp[0] = make_uint4(m_diag & valid, m_del & valid, m_ins & valid, pd_open & valid);
                tp[1] = make_uint4(pd_ext & valid, i_fin & valid & ge1, 0u, 0u);
            }
            if ((inf & INFO_STORE) && lane_active) {
                int *Mrow = slot + (uint64_t)(2 * (inf >> INFO_SLOT_SHIFT)) * W + j0;
#pragma unroll
                for (int b = 0; b < C; b += 4) {
                    if (j0 + b < W) {
                        *reinterpret_cast<int4 *>(Mrow + b) = make_int4(Mp[b], Mp[b + 1], Mp[b + 2], Mp[b + 3]);
                        *reinterpret_cast<int4 *>(Mrow + W + b) = make_int4(Dp[b], Dp[b + 1], Dp[b + 2], Dp[b + 3]);
                    }
                }
            }
            if (inf & INFO_STORE) __syncwarp();  // row-store writes of this warp are read by other lanes of the same warp later
            if (is_end && j0 <= Q && Q < j0 + C) {
                int sc = WINF;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (j0 + c == Q) sc = Mp[c];
                P.scores[a] = sc < WFIN ? (uint32_t)sc : (uint32_t)INF;
            }
        }
        __syncwarp();
    }
}

// gap_affine.rs:804-915 on the bit planes written by dp_fill_warp_kernel. One thread per alignment.
__global__ void traceback_planes_kernel(BatchDev P) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.n_alns) return;
    const AlnDesc ad = P.alns[a];
    const GraphDesc gd = P.graphs[ad.graph];
    const int Q = (int)ad.Q, V = (int)gd.V;
    const uint32_t *node_id = P.node_id + gd.node_base;
    const uint32_t *ord_slot = P.ord_slot + gd.node_base;
    const uint32_t *poff = P.pred_off + gd.predoff_base;
    const uint32_t *planes = (const uint32_t *)P.trace + ad.ptrace_base;
    const uint32_t *ord_store = planes + (uint64_t)V * 32 * PL_WORDS;
    uint32_t *out_r = P.out_r + ad.out_base, *out_q = P.out_q + ad.out_base;
    uint32_t pos = ad.out_cap;  // fill from the back
    int32_t status = 0;
    const uint32_t score = P.scores[a];
    auto bit = [&](int r, int o, int plane) -> bool {
        return (planes[((uint64_t)r * 32 + (o >> 5)) * PL_WORDS + plane] >> (o & 31)) & 1;
    };
    auto ord_of = [&](int r, int o, int shift) -> uint32_t {
        const uint32_t s = ord_slot[r];
        if (s == 0xFFFFFFFFu) return 0;
        return (ord_store[(uint64_t)s * 1024 + (o & 31) * 32 + (o >> 5)] >> shift) & 0x3FF;
    };
    if (score >= (uint32_t)INF) {
        status = POASTA_ERR_NO_ALIGNMENT;
    } else if (Q == 1) {  // gap_affine.rs:812-824
        --pos;
        out_r[pos] = node_id[V - 1];
        out_q[pos] = 0;
    } else if (Q > 1) {
        auto step = [&](int r, int o, int s, int &nr, int &no, int &ns) -> bool {
            if (s == 0) {
                if (bit(r, o, PL_M_DIAG)) {
                    nr = (int)P.pred_rank[poff[r] + ord_of(r, o, 0)];
                    no = (r == V - 1) ? o : o - 1;
                    ns = 0;
                    return true;
                }
                if (bit(r, o, PL_M_DEL)) {
                    nr = r, no = o, ns = 1;
                    return true;
                }
                if (bit(r, o, PL_M_INS)) {
                    nr = r, no = o, ns = 2;
                    return true;
                }
                return false;
            } else if (s == 1) {
                if (bit(r, o, PL_D_OPEN)) {
                    nr = (int)P.pred_rank[poff[r] + ord_of(r, o, 10)];
                    no = o, ns = 0;
                    return true;
                }
                if (bit(r, o, PL_D_EXT)) {
                    nr = (int)P.pred_rank[poff[r] + ord_of(r, o, 20)];
                    no = o, ns = 1;
                    return true;
                }
                return false;
            } else {
                if (!bit(r, o, PL_I)) return false;
                nr = r, no = o - 1, ns = 0;  // both open and extend resume in Match state (gap_affine.rs:639-650)
                return true;
            }
        };
        int cr = V - 1, co = Q, cs = 0, br = 0, bo = 0, bs = 0;
        if (step(cr, co, cs, br, bo, bs)) {
            cr = br, co = bo, cs = bs;
            while (step(cr, co, cs, br, bo, bs)) {
                if (cs == 0 && bs != 0) {
                    cr = br, co = bo, cs = bs;
                    continue;
                }
                if (pos == 0) {
                    status = POASTA_ERR_UNSUPPORTED;
                    break;
                }
                --pos;
                if (cs == 0) out_r[pos] = node_id[cr], out_q[pos] = (uint32_t)(co - 1);
                else if (cs == 2) out_r[pos] = POASTA_NONE, out_q[pos] = (uint32_t)(co - 1);
                else out_r[pos] = node_id[cr], out_q[pos] = POASTA_NONE;
                if (br == 0) break;  // predecessor is the start node (rank 0)
                cr = br, co = bo, cs = bs;
            }
        }
    }
    P.n_pairs[a] = ad.out_cap - pos;
    P.status[a] = status;
}

}  // namespace poasta
