// dp_fill_warp_kernel — v2 of the M/I/D fill (included by align.cu after BatchDev / TraceFmt).
//
// One WARP per alignment, no block barriers. Lane l owns the 32 contiguous query columns [32 l, 32 l + 32)
// of every row (queries up to 1023 bases; longer ones take the CTA-per-alignment kernel). The previous row's
// M and D for those columns live in REGISTERS, so a row whose only predecessor is the row just computed (the
// common case in a POA graph laid out in topological order) touches no memory except its traceback words:
//   * diagonal / deletion terms straight from registers, the left neighbour's column via one shuffle;
//   * symbol tests from per-lane 32-bit equality masks of the query (computed once per symbol and alignment);
//   * the in-row insertion chain I[j] = o + j e + min_{k<j, openI(k)} (min(Mdiag, D)[k] - k e) as a sequential
//     prefix-min over the lane's own columns plus ONE 5-step warp-shuffle prefix-min per row;
//   * min-of-three / add-min recurrences through the DPX intrinsics (__vimin3_s32, __viaddmin_s32).
// Rows with several predecessors, or whose predecessor is not the previous row, read predecessor rows from a
// per-warp row store in HBM; only rows that some later row needs from there are written to it (host flag).
// Traceback words have the same meaning and layout as in dp_fill_kernel, so traceback_kernel serves both.
#pragma once

namespace poasta {

constexpr int WARP_COLS = 32;      // columns per lane
constexpr int WARP_MAX_W = 1024;   // 32 lanes x 32 columns
constexpr int WARP_KERNEL_WARPS = 4;

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

template <typename TraceT, bool WRITE_TRACE>
__global__ void __launch_bounds__(WARP_KERNEL_WARPS * 32) dp_fill_warp_kernel(BatchDev P) {
    constexpr int OB = TraceFmt<TraceT>::ORD_BITS;
    constexpr int C = WARP_COLS;
    __shared__ uint32_t s_ords[WARP_KERNEL_WARPS][C][32];  // generic rows: packed predecessor ordinals per cell, [c][lane]
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
        const uint8_t *sym = P.node_sym + gd.node_base;
        const uint32_t *info = P.node_info + gd.node_base;
        const uint32_t *poff = P.pred_off + gd.predoff_base;
        TraceT *trace = (TraceT *)P.trace + ad.trace_base;

        // equality masks for the four common symbols, others on demand
        const SymMask mA = make_mask(q, Q, j0, 'A'), mC = make_mask(q, Q, j0, 'C'), mG = make_mask(q, Q, j0, 'G'),
                      mT = make_mask(q, Q, j0, 'T');
        auto mask_of = [&](int s) -> SymMask {
            if (s == 'A') return mA;
            if (s == 'C') return mC;
            if (s == 'G') return mG;
            if (s == 'T') return mT;
            return make_mask(q, Q, j0, s);
        };
        uint32_t valid = 0, ltQ = 0;  // bit c: j <= Q, j < Q
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (j0 + c <= Q) valid |= 1u << c;
            if (j0 + c < Q) ltQ |= 1u << c;
        }
        const bool lane_active = j0 < W;

        int Mp[C], Dp[C];  // previous row (rank r-1), final values
#pragma unroll
        for (int c = 0; c < C; ++c) Mp[c] = INF, Dp[c] = INF;

        for (int r = 0; r < V; ++r) {
            const uint32_t inf = info[r];
            const int vs = sym[r];
            const bool is_end = inf & INFO_IS_END, is_start = inf & INFO_IS_START;
            const uint32_t pb = poff[r], pe = poff[r + 1];
            const int npred = (int)(pe - pb);
            const SymMask ms = mask_of(vs);
            const uint32_t openD_mask = is_end ? 0u : (valid & ~ms.n);
            uint32_t openI_mask = 0;
            if ((inf & INFO_HAS_SUCC) && !is_end) openI_mask = (inf & INFO_ALWAYS_OPEN) ? ltQ : (ltQ & ~mask_of((int)(inf & 0xFF)).n);
            int Xr[C];                         // min(Mdiag, D) per column
            uint32_t diag_le = 0, d_open = 0;  // bit c: Mdiag <= D ; D equals the best "open" candidate
            bool generic = true;

            if (npred == 1 && !is_end) {
                // ---- single predecessor: registers (pred == r-1) or one reload from the row store ----
                generic = false;
                const int p = (int)P.pred_rank[pb];
                int mleft;
                if (p != r - 1) {
                    const int *Ms = slot + (uint64_t)(2 * P.pred_slot[pb]) * W;
                    if (lane_active) {
#pragma unroll
                        for (int k = 0; k < C / 4; ++k) {
                            if (j0 + 4 * k < W) {
                                const int4 m4 = *reinterpret_cast<const int4 *>(Ms + j0 + 4 * k);
                                const int4 d4 = *reinterpret_cast<const int4 *>(Ms + W + j0 + 4 * k);
                                Mp[4 * k] = m4.x, Mp[4 * k + 1] = m4.y, Mp[4 * k + 2] = m4.z, Mp[4 * k + 3] = m4.w;
                                Dp[4 * k] = d4.x, Dp[4 * k + 1] = d4.y, Dp[4 * k + 2] = d4.z, Dp[4 * k + 3] = d4.w;
                            } else {
                                Mp[4 * k] = Mp[4 * k + 1] = Mp[4 * k + 2] = Mp[4 * k + 3] = INF;
                                Dp[4 * k] = Dp[4 * k + 1] = Dp[4 * k + 2] = Dp[4 * k + 3] = INF;
                            }
                        }
                    }
                }
                mleft = __shfl_up_sync(0xFFFFFFFFu, Mp[C - 1], 1);
                if (lane == 0) mleft = INF;
                int prevM = mleft;  // M[p][j-1]
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int mpc = Mp[c];
                    const int sub = (ms.m >> c) & 1 ? 0 : x;
                    int mdiag = min(prevM + sub, INF);
                    if (!((valid >> c) & 1) || (j0 + c) == 0) mdiag = INF;
                    const int dopen = min(mpc + oe, INF);
                    const int dext = min(Dp[c] + e, INF);
                    int d = (openD_mask >> c) & 1 ? min(dopen, dext) : dext;
                    if (!((valid >> c) & 1)) d = INF;
                    if (d < INF && dopen == d) d_open |= 1u << c;
                    if (mdiag <= d) diag_le |= 1u << c;
                    Xr[c] = min(mdiag, d);
                    Dp[c] = d;
                    prevM = mpc;
                }
            }
            if (generic) {
                // ---- start row, end row, several predecessors: every predecessor row comes from the row store ----
#pragma unroll
                for (int cc = 0; cc < C / 8; ++cc) {
                    int Xa[8], dom[8], dem[8];
                    uint32_t ords[8];  // mord | dopen_ord << 10 | dext_ord << 20
#pragma unroll
                    for (int c = 0; c < 8; ++c) Xa[c] = INF, dom[c] = INF, dem[c] = INF, ords[c] = 0;
                    const int jb = j0 + cc * 8;
                    if (jb < W) {
                        for (uint32_t k = pb; k < pe; ++k) {
                            const uint32_t ord = k - pb;
                            const int *Ms = slot + (uint64_t)(2 * P.pred_slot[k]) * W;
                            const int4 ma = *reinterpret_cast<const int4 *>(Ms + jb), mb = *reinterpret_cast<const int4 *>(Ms + jb + 4);
                            const int4 da = *reinterpret_cast<const int4 *>(Ms + W + jb), db = *reinterpret_cast<const int4 *>(Ms + W + jb + 4);
                            const int ml = jb > 0 ? Ms[jb - 1] : INF;
                            const int mp[9] = {ml, ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
                            const int dp[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const int cg = cc * 8 + c, j = jb + c;
                                int cand;
                                if (is_end) cand = mp[c + 1];
                                else cand = (j >= 1 && j <= Q) ? min(mp[c] + ((ms.m >> cg) & 1 ? 0 : x), INF) : INF;
                                if (cand < Xa[c]) Xa[c] = cand, ords[c] = (ords[c] & ~0x3FFu) | ord;
                                const int dop = min(mp[c + 1] + oe, INF), dex = min(dp[c] + e, INF);
                                if (dop < dom[c]) dom[c] = dop, ords[c] = (ords[c] & ~(0x3FFu << 10)) | (ord << 10);
                                if (dex < dem[c]) dem[c] = dex, ords[c] = (ords[c] & ~(0x3FFu << 20)) | (ord << 20);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int cg = cc * 8 + c, j = jb + c;
                        const bool v = (valid >> cg) & 1;
                        int d = v ? min((openD_mask >> cg) & 1 ? dom[c] : INF, dem[c]) : INF;
                        int xd = v ? Xa[c] : INF;
                        if (is_start && j == 0) xd = 0;
                        bool is_open = d < INF && dom[c] == d;
                        if (d < INF && !is_open && !((openD_mask >> cg) & 1) && npred > 1 && dom[c] < d) {
                            // rare: a non-minimal "open" candidate equals D although opening is not allowed here
                            // (the reference's backtrace tests score equality only)
                            for (uint32_t k = pb; k < pe; ++k) {
                                if (min(slot[(uint64_t)(2 * P.pred_slot[k]) * W + j] + oe, INF) == d) {
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
                        s_ords[wib][cg][lane] = ords[c];
                    }
                }
            }

            // ---- insertion chain: prefix-min over the whole row ----
            int run = INF;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int av = ((openI_mask >> c) & 1) && Xr[c] < INF ? Xr[c] + (Q - (j0 + c)) * e : INF;
                run = min(run, av);
            }
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl = min(incl, o);
            }
            int pm = __shfl_up_sync(0xFFFFFFFFu, incl, 1);  // exclusive prefix-min over all columns left of mine
            if (lane == 0) pm = INF;

            // ---- final M, traceback words, stores ----
            TraceT *tp = trace + (uint64_t)r * W + j0;
            int *Mrow = slot + (uint64_t)(2 * (inf >> INFO_SLOT_SHIFT)) * W + j0;
            const bool store_row = inf & INFO_STORE;
#pragma unroll
            for (int c8 = 0; c8 < C / 8; ++c8) {
                uint32_t tw[8];
#pragma unroll
                for (int c = c8 * 8; c < c8 * 8 + 8; ++c) {
                    const int j = j0 + c;
                    const bool v = (valid >> c) & 1;
                    const int iv = (v && j >= 1 && pm < INF) ? min(pm + oe - (Q - j) * e - e, INF) : INF;
                    const int xr = Xr[c], dv = Dp[c];
                    const int mv = min(xr, iv);
                    // advance the running prefix with this column
                    const int av = ((openI_mask >> c) & 1) && xr < INF ? xr + (Q - j) * e : INF;
                    pm = min(pm, av);
                    Mp[c] = mv;
                    if (WRITE_TRACE) {
                        uint32_t icode = iv < INF ? I_OPEN : I_NONE, mcode = M_NONE, dcode = D_NONE, mo = 0, dox = 0;
                        uint32_t ords = 0;
                        if (generic) ords = s_ords[wib][c][lane];
                        if (dv < INF) {
                            if ((d_open >> c) & 1) dcode = D_OPEN, dox = (ords >> 10) & 0x3FF;
                            else dcode = D_EXT, dox = (ords >> 20) & 0x3FF;
                        }
                        if (mv < INF) {
                            if (xr == mv) {
                                if (((diag_le >> c) & 1) && j >= 1 && npred > 0) mcode = M_DIAG, mo = ords & 0x3FF;
                                else if (dv == mv) mcode = M_DEL;
                                else if (iv == mv) mcode = M_INS;
                            } else {
                                mcode = M_INS;
                            }
                        }
                        tw[c - c8 * 8] = icode | (mcode << 2) | (dcode << 4) | (mo << 6) | (dox << (6 + OB));
                    }
                }
                if (lane_active && j0 + c8 * 8 < W) {
                    if (WRITE_TRACE) {
                        if constexpr (sizeof(TraceT) == 2) {
                            *reinterpret_cast<uint4 *>(tp + c8 * 8) = make_uint4(tw[0] | (tw[1] << 16), tw[2] | (tw[3] << 16),
                                                                                 tw[4] | (tw[5] << 16), tw[6] | (tw[7] << 16));
                        } else {
                            *reinterpret_cast<uint4 *>(tp + c8 * 8) = make_uint4(tw[0], tw[1], tw[2], tw[3]);
                            *reinterpret_cast<uint4 *>(tp + c8 * 8 + 4) = make_uint4(tw[4], tw[5], tw[6], tw[7]);
                        }
                    }
                    if (store_row) {
                        const int b = c8 * 8;
                        *reinterpret_cast<int4 *>(Mrow + b) = make_int4(Mp[b], Mp[b + 1], Mp[b + 2], Mp[b + 3]);
                        *reinterpret_cast<int4 *>(Mrow + b + 4) = make_int4(Mp[b + 4], Mp[b + 5], Mp[b + 6], Mp[b + 7]);
                        *reinterpret_cast<int4 *>(Mrow + W + b) = make_int4(Dp[b], Dp[b + 1], Dp[b + 2], Dp[b + 3]);
                        *reinterpret_cast<int4 *>(Mrow + W + b + 4) = make_int4(Dp[b + 4], Dp[b + 5], Dp[b + 6], Dp[b + 7]);
                    }
                }
            }
            if (store_row) __syncwarp();  // row-store writes of this warp are read by other lanes of the same warp later
            if (is_end && j0 <= Q && Q < j0 + C) {
                int sc = INF;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (j0 + c == Q) sc = Mp[c];
                P.scores[a] = (uint32_t)sc;
            }
        }
        __syncwarp();
    }
}

}  // namespace poasta
