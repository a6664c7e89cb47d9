// CUDA hot path: gap-affine M/I/D fill per graph node in topological order + on-device traceback.
//
// What is computed (reference semantics):
//   * transitions follow POASTA's rules, not textbook Gotoh: a deletion is opened into node c at
//     offset j only if c mismatches q[j] (or j == Q); an insertion is opened at (v, j) only if v has
//     a successor that mismatches q[j] or is `end`; matches are free
//     (/root/reference/src/aligner/dfa.rs:210-250, scoring/gap_affine.rs:346-430);
//   * per cell a traceback word records what the reference's score-equality backtrace
//     (gap_affine.rs:550-657) decides at that cell: the first predecessor (Incoming order) whose
//     score matches, else close-deletion, else close-insertion; deletion: open before extend;
//     insertion: open before extend (the extend case resumes in Match state, gap_affine.rs:647-650);
//   * the walk kernel applies gap_affine.rs:804-915 to those words.
//
// Kernels of the DP path (tiebreak = DP, and the score cross-check of the exact mode), chosen per batch in run_chunk:
//   dp_fill_packed_kernel  (dp_packed_kernel.cuh) one warp per alignment, two 16-bit scores per register, DPX packed
//                    instructions written as intrinsics (__vimin3_s16x2, __vmins2), row records staged by TMA bulk copies,
//                    M / D rows stored as the traceback store; queries <= 1023, graphs over ACGT, scores provably < 0x5000.
//   dp_fill_warp_kernel    (dp_warp_kernel.cuh) the 32-bit version with bit-plane traceback: same shape limits, any costs / symbols.
//   dp_fill_kernel         (below) persistent CTAs, one alignment per CTA, the query axis split over 256 threads in 4-column
//                    groups; for queries of 1024+ bases or in-degree > 1023. Its recurrences are plain min() / + in 32 bits
//                    (ptxas fuses them into VIADDMNMX / VIMNMX3; no intrinsic is called here).
//   traceback_rows_kernel / traceback_planes_kernel / traceback_kernel   the walkers matching the three fills.
//   compact_kernel   packs the pairs of all alignments densely for one D2H copy.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "device_batch.hpp"
#include "graph_handle.hpp"
#include "replay_kernel.cuh"
#include "msa_engine.cuh"

namespace poasta {

constexpr int INF = 0x3FFFFFFF;
constexpr int FILL_THREADS = 256;

struct BatchDev {
    const GraphDesc *graphs;
    const AlnDesc *alns;
    const uint8_t *node_sym;
    const uint32_t *node_info, *node_id, *ord_slot, *pred_off, *pred_rank, *pred_slot;
    const RowRec *rows;
    const uint8_t *queries;
    uint32_t n_alns;
    int x, oe, e;
    int *scratch;             // per CTA slot: 2 * max_cells ints (M rows then D rows interleaved per rank)
    uint64_t scratch_stride;  // ints per slot
    void *trace;              // u16 or u32 words
    uint32_t *scores;         // per alignment
    uint32_t *counter;        // work queue
    // traceback outputs
    uint32_t *out_r, *out_q;  // staging, per alignment region [out_base, out_base + out_cap), filled from the back
    uint32_t *n_pairs;        // per alignment
    int32_t *status;          // per alignment
};

// trace word layout: [1:0] I code, [3:2] M code, [5:4] D code, then M ordinal, then D ordinal
template <typename T> struct TraceFmt;
template <> struct TraceFmt<uint16_t> {
    static constexpr int ORD_BITS = 5;
};
template <> struct TraceFmt<uint32_t> {
    static constexpr int ORD_BITS = 13;
};
enum { I_NONE = 0, I_OPEN = 1, I_EXT = 2 };
enum { M_NONE = 0, M_DIAG = 1, M_DEL = 2, M_INS = 3 };
enum { D_NONE = 0, D_OPEN = 1, D_EXT = 2 };

__device__ __forceinline__ int sat(int v) { return min(v, INF); }

template <typename TraceT>
__global__ void __launch_bounds__(FILL_THREADS) dp_fill_kernel(BatchDev P) {
    constexpr int OB = TraceFmt<TraceT>::ORD_BITS;
    __shared__ uint32_t s_next;
    __shared__ int s_warp[FILL_THREADS / 32];
    __shared__ int s_carry[2];    // running prefix-min carried across column groups (double-buffered by group parity)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *const slot = P.scratch + (uint64_t)blockIdx.x * P.scratch_stride;
    const int x = P.x, oe = P.oe, e = P.e;

    for (;;) {
        if (tid == 0) s_next = atomicAdd(P.counter, 1u);
        __syncthreads();
        const uint32_t a = s_next;
        __syncthreads();
        if (a >= P.n_alns) break;
        const AlnDesc ad = P.alns[a];
        const GraphDesc gd = P.graphs[ad.graph];
        const int Q = (int)ad.Q, W = (int)ad.W, V = (int)gd.V;
        const uint8_t *q = P.queries + ad.query_base;
        const uint8_t *sym = P.node_sym + gd.node_base;
        const uint32_t *info = P.node_info + gd.node_base;
        const uint32_t *poff = P.pred_off + gd.predoff_base;
        TraceT *trace = (TraceT *)P.trace + ad.trace_base;
        const int n_groups = (W / 4 + FILL_THREADS - 1) / FILL_THREADS;

        for (int r = 0; r < V; ++r) {
            const uint32_t inf = info[r];
            const int vs = sym[r];
            const bool is_end = inf & INFO_IS_END, is_start = inf & INFO_IS_START;
            const uint32_t pb = poff[r], pe = poff[r + 1];
            int *Mrow = slot + (uint64_t)(2 * r) * W, *Drow = Mrow + W;
            if (tid == 0) s_carry[0] = INF;
            for (int g = 0; g < n_groups; ++g) {
                const int j0 = (g * FILL_THREADS + tid) * 4;
                const bool active = j0 < W;
                int Xv[4], Dv[4], mord[4], dopen_ord[4], dext_ord[4], dopen_min[4], dext_min[4], av[4];
                int qc[5];  // q[j0-1 .. j0+3] as ints (or -1 outside)
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int j = j0 + c - 1;
                    qc[c] = (active && j >= 0 && j < Q) ? q[j] : -1;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    Xv[c] = INF, Dv[c] = INF, mord[c] = 0, dopen_ord[c] = 0, dext_ord[c] = 0;
                    dopen_min[c] = INF, dext_min[c] = INF;
                }
                if (active) {
                    for (uint32_t k = pb; k < pe; ++k) {
                        const int p = (int)P.pred_rank[k];
                        const int ord = (int)(k - pb);
                        const int *Mp = slot + (uint64_t)(2 * p) * W;
                        const int4 m4 = *reinterpret_cast<const int4 *>(Mp + j0);
                        const int4 d4 = *reinterpret_cast<const int4 *>(Mp + W + j0);
                        const int mleft = j0 > 0 ? Mp[j0 - 1] : INF;
                        const int mp[5] = {mleft, m4.x, m4.y, m4.z, m4.w};
                        const int dp[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int j = j0 + c;
                            // match / mismatch (end node: same offset, no symbol compare; graphs/poa.rs:463-465)
                            int cand;
                            if (is_end) cand = mp[c + 1];
                            else cand = (j >= 1 && j <= Q) ? sat(mp[c] + (qc[c] == vs ? 0 : x)) : INF;
                            if (cand < Xv[c]) Xv[c] = cand, mord[c] = ord;
                            // deletion: open from M[p][j] (value only if allowed), extend from D[p][j]
                            const int dop = sat(mp[c + 1] + oe), dex = sat(dp[c] + e);
                            if (dop < dopen_min[c]) dopen_min[c] = dop, dopen_ord[c] = ord;
                            if (dex < dext_min[c]) dext_min[c] = dex, dext_ord[c] = ord;
                        }
                    }
                }
                bool openD[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    openD[c] = !is_end && j <= Q && (j >= Q || qc[c + 1] != vs);
                    Dv[c] = (active && j <= Q) ? min(openD[c] ? dopen_min[c] : INF, dext_min[c]) : INF;
                    if (is_start && j == 0 && active) Xv[c] = 0;
                    if (!active || j > Q) Xv[c] = INF;
                }
                // insertion chain: I[j] = o + j*e + min_{k<j, openI(k)} (min(Xdiag,D)[k] - k*e)
                const bool always = inf & INFO_ALWAYS_OPEN, has_succ = inf & INFO_HAS_SUCC;
                const int usym = inf & 0xFF;
                int run = INF;  // inclusive prefix-min over my 4 columns
                int pre[4];     // exclusive prefix within my 4 columns
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    const bool openI = active && j < Q && has_succ && !is_end && (always || qc[c + 1] != usym);
                    const int xm = min(Xv[c], Dv[c]);
                    av[c] = (openI && xm < INF) ? xm + (Q - j) * e : INF;
                    pre[c] = run;
                    run = min(run, av[c]);
                }
                int incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl = min(incl, o);
                }
                int excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
                if (lane == 0) excl = INF;
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                int base = s_carry[g & 1];
                for (int w = 0; w < warp; ++w) base = min(base, s_warp[w]);
                excl = min(excl, base);
                int Iv[4], Mv[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    const int pm = min(excl, pre[c]);
                    Iv[c] = (active && j >= 1 && j <= Q && pm < INF) ? sat(pm + oe - (Q - j) * e - e) : INF;
                    Mv[c] = min(Xv[c], min(Dv[c], Iv[c]));
                }
                if (tid == FILL_THREADS - 1) s_carry[(g + 1) & 1] = min(base, incl);
                if (active) {
                    // deletion trace when opening is not allowed by the rules but the score equality may still hold
                    // by coincidence for some predecessor (the reference's backtrace tests equality only)
                    TraceT tw[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = j0 + c;
                        uint32_t icode = I_NONE, mcode = M_NONE, dcode = D_NONE, mo = 0, dox = 0;
                        if (j <= Q) {
                            if (Dv[c] < INF) {
                                if (dopen_min[c] == Dv[c]) {
                                    dcode = D_OPEN, dox = dopen_ord[c];
                                } else {
                                    bool found = false;
                                    if (!openD[c] && pe - pb > 1 && dopen_min[c] < Dv[c]) {  // rare second pass: non-minimal open candidate equal to D
                                        for (uint32_t k = pb; k < pe && !found; ++k) {
                                            const int p = (int)P.pred_rank[k];
                                            const int mpj = slot[(uint64_t)(2 * p) * W + j];
                                            if (sat(mpj + oe) == Dv[c]) dcode = D_OPEN, dox = k - pb, found = true;
                                        }
                                    }
                                    if (!found && dext_min[c] == Dv[c]) dcode = D_EXT, dox = dext_ord[c];
                                }
                            }
                            // a finite I[v][j] always satisfies one of the two equalities of gap_affine.rs:636-651 and both
                            // resume at (v, j-1) in Match state, so one "has predecessor" code is enough
                            if (Iv[c] < INF) icode = I_OPEN;
                            if (Mv[c] < INF) {
                                if (j >= 1 && Xv[c] == Mv[c] && pe > pb) mcode = M_DIAG, mo = mord[c];
                                else if (Dv[c] == Mv[c]) mcode = M_DEL;
                                else if (Iv[c] == Mv[c]) mcode = M_INS;
                            }
                        }
                        tw[c] = (TraceT)(icode | (mcode << 2) | (dcode << 4) | (mo << 6) | (dox << (6 + OB)));
                    }
                    *reinterpret_cast<int4 *>(Mrow + j0) = make_int4(Mv[0], Mv[1], Mv[2], Mv[3]);
                    *reinterpret_cast<int4 *>(Drow + j0) = make_int4(Dv[0], Dv[1], Dv[2], Dv[3]);
                    TraceT *tp = trace + (uint64_t)r * W + j0;
                    if constexpr (sizeof(TraceT) == 2) {
                        *reinterpret_cast<uint2 *>(tp) =
                            make_uint2((uint32_t)tw[0] | ((uint32_t)tw[1] << 16), (uint32_t)tw[2] | ((uint32_t)tw[3] << 16));
                    } else {
                        *reinterpret_cast<uint4 *>(tp) = make_uint4(tw[0], tw[1], tw[2], tw[3]);
                    }
                    if (is_end && j0 <= Q && Q < j0 + 4) P.scores[a] = (uint32_t)Mv[Q - j0];
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace poasta
#include "dp_warp_kernel.cuh"
#include "dp_packed_kernel.cuh"
namespace poasta {

// gap_affine.rs:804-915 on the trace words. One thread per alignment.
template <typename TraceT>
__global__ void traceback_kernel(BatchDev P) {
    constexpr int OB = TraceFmt<TraceT>::ORD_BITS;
    constexpr uint32_t OM = (1u << OB) - 1;
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.n_alns) return;
    const AlnDesc ad = P.alns[a];
    const GraphDesc gd = P.graphs[ad.graph];
    const int Q = (int)ad.Q, W = (int)ad.W, V = (int)gd.V;
    const uint32_t *node_id = P.node_id + gd.node_base;
    const uint32_t *poff = P.pred_off + gd.predoff_base;
    const TraceT *trace = (const TraceT *)P.trace + ad.trace_base;
    uint32_t *out_r = P.out_r + ad.out_base, *out_q = P.out_q + ad.out_base;
    uint32_t pos = ad.out_cap;  // fill from the back
    int32_t status = 0;
    const uint32_t score = P.scores[a];
    if (score >= (uint32_t)INF) {
        status = POASTA_ERR_NO_ALIGNMENT;
    } else if (Q == 1) {  // gap_affine.rs:812-824: `end` equals any symbol, so a 1-base query pairs with the end node
        --pos;
        out_r[pos] = node_id[V - 1];
        out_q[pos] = 0;
    } else if (Q > 1) {
        // step 0: predecessor of the end state (end, Q, Match)
        int cr = V - 1, co = Q, cs = 0;  // current rank, offset, state (0 M, 1 D, 2 I)
        bool have = false;
        int br = 0, bo = 0, bs = 0;
        auto step = [&](int r, int o, int s, int &nr, int &no, int &ns) -> bool {
            const uint32_t t = trace[(uint64_t)r * W + o];
            const uint32_t ic = t & 3, mc = (t >> 2) & 3, dc = (t >> 4) & 3, mo = (t >> 6) & OM, dd = (t >> (6 + OB)) & OM;
            if (s == 0) {
                if (mc == M_DIAG) {
                    nr = (int)P.pred_rank[poff[r] + mo];
                    no = (r == V - 1) ? o : o - 1;
                    ns = 0;
                    return true;
                }
                if (mc == M_DEL) {
                    nr = r, no = o, ns = 1;
                    return true;
                }
                if (mc == M_INS) {
                    nr = r, no = o, ns = 2;
                    return true;
                }
                return false;
            } else if (s == 1) {
                if (dc == D_NONE) return false;
                nr = (int)P.pred_rank[poff[r] + dd];
                no = o;
                ns = dc == D_OPEN ? 0 : 1;
                return true;
            } else {
                if (ic == I_NONE) return false;
                nr = r, no = o - 1, ns = 0;  // both open and extend resume in Match state (gap_affine.rs:639-650)
                return true;
            }
        };
        have = step(cr, co, cs, br, bo, bs);
        if (have) {
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

// exclusive scan of n_pairs -> offsets (single block; n is at most a few 10^5)
__global__ void scan_kernel(const uint32_t *n_pairs, uint32_t *offsets, uint32_t n) {
    __shared__ uint32_t s_sum[1024];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t start = 0; start < n; start += 1024) {
        const uint32_t i = start + threadIdx.x;
        const uint32_t v = i < n ? n_pairs[i] : 0;
        s_sum[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t d = 1; d < 1024; d <<= 1) {
            const uint32_t t = threadIdx.x >= d ? s_sum[threadIdx.x - d] : 0;
            __syncthreads();
            s_sum[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) offsets[i] = s_base + s_sum[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_base += s_sum[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = s_base;
}

__global__ void compact_kernel(BatchDev P, const uint32_t *offsets, uint32_t *dense_r, uint32_t *dense_q) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P.n_alns) return;
    const AlnDesc ad = P.alns[warp];
    const uint32_t n = P.n_pairs[warp];
    const uint32_t src = ad.out_base + ad.out_cap - n, dst = offsets[warp];
    for (uint32_t i = lane; i < n; i += 32) {
        dense_r[dst + i] = P.out_r[src + i];
        dense_q[dst + i] = P.out_q[src + i];
    }
}

// ---------------------------------------------------------------------------------------------------
// Host side: grow-only device / pinned buffers, chunked launches.
struct Buf {
    void *d = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&d, want) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&d, bytes) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            want = bytes;
        }
        cap = want;
        return true;
    }
    void release() {
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
    }
};
struct PinBuf {
    void *h = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (h) cudaFreeHost(h);
        h = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        if (cudaHostAlloc(&h, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cap = want;
        return true;
    }
};

struct DeviceContext {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_done = nullptr;  // blocking-sync event: a waiting lane thread sleeps instead of spinning
    int n_sms = 0;
    Buf rows, graphs, alns, node_sym, node_info, node_id, ord_slot, pred_off, pred_rank, pred_slot, queries, scratch, trace, scores, counter, out_r, out_q,
        n_pairs, status, offsets, dense_r, dense_q;
    PinBuf h_scores, h_npairs, h_status, h_offsets, h_dense_r, h_dense_q, h_in;
    Buf r_graphs, r_nodes, r_succ, r_pred, r_bub, r_path_off, r_path_rem, r_slabs, r_stats;
    // what the search slabs were last laid out for: worker slabs keep their epoch-stamped tables across launches and are
    // zeroed only when the buffer moves or the layout / cell width changes (replay_kernel.cuh SlabHeader)
    void *slab_ptr = nullptr;
    uint64_t slab_sig = 0, slab_zeroed_bytes = 0;
    PinBuf h_stats;
    double last_kernel_ms = 0;
    uint32_t last_launches = 0;
    uint64_t last_cells = 0;

    int init(int dev) {
        if (device == dev) return 0;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
            cudaGetLastError();
            return POASTA_ERR_CUDA;
        }
        if (cudaSetDevice(dev) != cudaSuccess) return POASTA_ERR_CUDA;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return POASTA_ERR_CUDA;
        if (prop.major < 10) return POASTA_ERR_CUDA;  // sm_100a image only; no fallback
        n_sms = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return POASTA_ERR_CUDA;
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventCreateWithFlags(&ev_done, cudaEventBlockingSync | cudaEventDisableTiming);
        device = dev;
        return 0;
    }
};

// One context per (device, pipeline lane): a lane is one host thread + one stream driving its share of a batch, so that
// the host work of one lane (flattening, packing, graph insertion) overlaps the kernels of the others.
constexpr int MAX_LANES = 32;
static thread_local int tl_lane = 0;        // pipeline lane of the calling thread
static thread_local int tl_lane_count = 1;  // lanes sharing the device right now (memory budgets are divided by it)
constexpr int MAX_DEVICES = 16;
static DeviceContext &ctx_for(int device) {  // callers have checked 0 <= device < MAX_DEVICES
    static std::mutex mu;
    static DeviceContext ctxs[MAX_DEVICES][MAX_LANES];
    std::lock_guard<std::mutex> lk(mu);
    return ctxs[device][tl_lane % MAX_LANES];
}
// One public call at a time per device: the contexts (streams, grow-only device buffers, pinned staging) of a device are
// shared by every host thread that names it, so concurrent callers are serialised here for the WHOLE call. Different
// devices run concurrently. (Pipeline lanes inside one call use separate contexts.)
static std::mutex &device_call_mutex(int device) {
    static std::mutex mus[MAX_DEVICES];
    return mus[device];
}

static thread_local double tl_kernel_ms = 0, tl_search_ms = 0;
static thread_local double tl_lane_max_ms = -1;  // msa batches: the largest per-lane sum of kernel time (lanes overlap on the device)
static thread_local uint32_t tl_launches = 0;
static thread_local int tl_dp_kernel = 0;  // poasta_last_dp_kernel
static thread_local int tl_msa_path = 0;   // poasta_last_msa_path
static thread_local int tl_search_table = 0;  // poasta_last_search_table: 1 dense, 2 blocked (last search launch of poasta_align_batch)
static thread_local double tl_fill_ms = 0;  // poasta_last_batch_fill_ms
static thread_local uint32_t tl_fill_launches = 0;
static thread_local uint64_t tl_cells = 0;
static thread_local uint64_t tl_h2d = 0, tl_d2h = 0;
static thread_local uint64_t tl_engine_counts[3] = {0, 0, 0};  // queued, visited, pruned summed over the last msa batch (reference-exact mode)
static thread_local std::vector<uint32_t> tl_search_stats;  // per problem of the last align batch: queued, visited, pruned (search path only)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fprintf(stderr, "poasta_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return POASTA_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

struct ChunkResult {
    std::vector<uint32_t> scores, n_pairs, offsets;
    std::vector<int32_t> status;
    const uint32_t *dense_r = nullptr, *dense_q = nullptr;  // pinned, valid until the next run_chunk on this context
};

// Waits for everything queued on the lane's stream. Kernels of the search path run for seconds: the lane thread sleeps on a
// blocking-sync event so its core is free for the host phases of the other lanes.
static cudaError_t wait_stream(DeviceContext &c) {
    cudaError_t e = cudaEventRecord(c.ev_done, c.stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(c.ev_done);
}

}  // namespace poasta
#include "msa_engine_host.cuh"
namespace poasta {

// Runs one packed batch (must fit the memory budget) and leaves results in pinned host memory.
static int run_chunk(DeviceContext &c, const PackedBatch &pb, const poasta_params &prm, ChunkResult &res, bool scores_only) {
    const uint32_t n = (uint32_t)pb.alns.size();
    if (n == 0) return 0;
    CK(cudaSetDevice(c.device));
    const bool wide = pb.max_in_degree > 31;
    if (pb.max_in_degree > 8191) return POASTA_ERR_UNSUPPORTED;
    const size_t tsz = wide ? 4 : 2;
    // slots: resident CTAs; bounded by scratch memory
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // warp-per-alignment kernel when every query fits 32 lanes x 32 columns, else the CTA-per-alignment kernel
    static const bool force_cta = getenv("POASTA_DP_KERNEL") && !strcmp(getenv("POASTA_DP_KERNEL"), "cta");
    const bool use_warp = pb.max_W <= (uint32_t)WARP_MAX_W && pb.max_in_degree <= 1023 && !force_cta;
    // packed s16x2 kernel (dp_packed_kernel.cuh) when the optimal score provably stays below its 16-bit "infinite" and the
    // insertion prefix-min offsets fit: every global alignment costs at most "delete a whole path, insert the whole query"
    static const bool force_s32 = getenv("POASTA_DP_KERNEL") && !strcmp(getenv("POASTA_DP_KERNEL"), "warp32");
    // (max_lmin = the longest of the graphs' SHORTEST start -> end paths: that many nodes are deleted, the whole query inserted)
    const uint64_t worst_opt = 2ull * prm.gap_open + (uint64_t)prm.gap_extend * ((uint64_t)pb.max_lmin + pb.max_Q) + prm.mismatch;
    const bool packed_ok = !force_s32 && !force_cta && pb.acgt_only && worst_opt < PK_INF && (uint64_t)prm.gap_extend * 1024 <= PK_KMAX &&
                           pb.max_in_degree <= 1023;
    const bool use_packed = use_warp && packed_ok;
    // queries of 1024+ bases: a thread-block cluster of one-warp CTAs, one 1024-column segment each, linked through DSMEM
    const uint32_t n_seg_max = (pb.max_W + 1023) / 1024;
    const bool use_cluster = !use_warp && packed_ok && pb.max_W > (uint32_t)WARP_MAX_W && n_seg_max <= 8;
    const unsigned cluster_ctas = n_seg_max <= 2 ? 2 : (n_seg_max <= 4 ? 4 : 8);
    const size_t per_slot = use_cluster ? 16 : use_packed ? std::max<size_t>(16, (size_t)pb.max_n_stored * 32 * 128)
                            : use_warp ? (size_t)pb.max_store_cells * 2 * sizeof(int) : (size_t)pb.max_cells * 2 * sizeof(int);
    size_t slots = (size_t)c.n_sms * 4;  // CTA slots of the CTA-per-alignment kernel
    if (use_packed || use_cluster) {
        static bool attr_set[16] = {};
        if (!attr_set[c.device & 15]) {
            CK(cudaFuncSetAttribute(dp_fill_packed_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM_BYTES));
            CK(cudaFuncSetAttribute(dp_fill_packed_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM_BYTES));
            CK(cudaFuncSetAttribute(dp_fill_packed_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_CLUSTER_SMEM));
            attr_set[c.device & 15] = true;
        }
    }
    if (use_cluster) {
        slots = 4;  // no per-slot scratch: every row lives in the trace store
    } else if (use_packed) {
        int ctas = 0;
        if (scores_only) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_fill_packed_kernel<false, false>, PK_WARPS * 32, PK_SMEM_BYTES);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_fill_packed_kernel<true, false>, PK_WARPS * 32, PK_SMEM_BYTES);
        slots = (size_t)c.n_sms * std::max(1, ctas) * PK_WARPS;
    } else if (use_warp) {               // warp slots: what the register file lets be resident
        int ctas = 0;
        if (scores_only) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_fill_warp_kernel<false>, WARP_KERNEL_WARPS * 32, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, dp_fill_warp_kernel<true>, WARP_KERNEL_WARPS * 32, 0);
        slots = (size_t)c.n_sms * std::max(1, ctas) * WARP_KERNEL_WARPS;
    }
    slots = std::min<size_t>(slots, use_warp ? ((size_t)n + 3) / 4 * 4 : n);
    const size_t scratch_budget = std::max<size_t>(((size_t)8 << 30) / tl_lane_count, c.scratch.cap);
    while (slots > 4 && slots * per_slot > scratch_budget) slots = std::max<size_t>(4, slots / 2 / 4 * 4);
    if (!c.scratch.reserve(slots * per_slot) || !c.trace.reserve(use_cluster ? pb.prow_halves * 2 : use_packed ? (scores_only ? 16 : pb.prow_halves * 2) : use_warp ? (scores_only ? 16 : pb.ptrace_words * 4) : pb.trace_elems * tsz)) return POASTA_ERR_CUDA;  // only the warp kernel can skip the trace
    bool ok = c.graphs.reserve(pb.graphs.size() * sizeof(GraphDesc)) && c.alns.reserve(n * sizeof(AlnDesc)) &&
              c.node_sym.reserve(pb.node_sym.size()) && c.node_info.reserve(pb.node_info.size() * 4) &&
              c.node_id.reserve(pb.node_id.size() * 4) && c.ord_slot.reserve(pb.ord_slot.size() * 4) &&
              c.rows.reserve(pb.rows.size() * sizeof(RowRec)) &&
              c.pred_off.reserve(pb.pred_off.size() * 4) &&
              c.pred_rank.reserve(pb.pred_rank.size() * 4 + 4) && c.pred_slot.reserve(pb.pred_slot.size() * 4 + 4) &&
              c.queries.reserve(pb.queries.size() + 4) &&
              c.scores.reserve(n * 4) && c.counter.reserve(8) && c.out_r.reserve(pb.out_elems * 4) &&
              c.out_q.reserve(pb.out_elems * 4) && c.n_pairs.reserve(n * 4) && c.status.reserve(n * 4) &&
              c.offsets.reserve((n + 1) * 4);
    if (!ok) return POASTA_ERR_CUDA;
    cudaStream_t s = c.stream;
    CK(cudaMemcpyAsync(c.graphs.d, pb.graphs.data(), pb.graphs.size() * sizeof(GraphDesc), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.alns.d, pb.alns.data(), n * sizeof(AlnDesc), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.node_sym.d, pb.node_sym.data(), pb.node_sym.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.node_info.d, pb.node_info.data(), pb.node_info.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.node_id.d, pb.node_id.data(), pb.node_id.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.ord_slot.d, pb.ord_slot.data(), pb.ord_slot.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.rows.d, pb.rows.data(), pb.rows.size() * sizeof(RowRec), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c.pred_off.d, pb.pred_off.data(), pb.pred_off.size() * 4, cudaMemcpyHostToDevice, s));
    if (!pb.pred_rank.empty()) {
        CK(cudaMemcpyAsync(c.pred_rank.d, pb.pred_rank.data(), pb.pred_rank.size() * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(c.pred_slot.d, pb.pred_slot.data(), pb.pred_slot.size() * 4, cudaMemcpyHostToDevice, s));
    }
    if (!pb.queries.empty()) CK(cudaMemcpyAsync(c.queries.d, pb.queries.data(), pb.queries.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(c.counter.d, 0, 8, s));  // [0] work queue, [1] abort flag of the cluster kernel
    CK(cudaMemsetAsync(c.scores.d, 0x3F, n * 4, s));
    tl_h2d += pb.graphs.size() * sizeof(GraphDesc) + n * sizeof(AlnDesc) + pb.node_sym.size() + pb.node_info.size() * 4 +
              pb.node_id.size() * 4 + pb.pred_off.size() * 4 + pb.pred_rank.size() * 8 + pb.ord_slot.size() * 4 +
              pb.rows.size() * sizeof(RowRec) + pb.queries.size();

    BatchDev P;
    P.graphs = (const GraphDesc *)c.graphs.d;
    P.alns = (const AlnDesc *)c.alns.d;
    P.node_sym = (const uint8_t *)c.node_sym.d;
    P.node_info = (const uint32_t *)c.node_info.d;
    P.node_id = (const uint32_t *)c.node_id.d;
    P.ord_slot = (const uint32_t *)c.ord_slot.d;
    P.rows = (const RowRec *)c.rows.d;
    P.pred_off = (const uint32_t *)c.pred_off.d;
    P.pred_rank = (const uint32_t *)c.pred_rank.d;
    P.pred_slot = (const uint32_t *)c.pred_slot.d;
    P.queries = (const uint8_t *)c.queries.d;
    P.n_alns = n;
    P.x = prm.mismatch;
    P.oe = (int)prm.gap_open + (int)prm.gap_extend;
    P.e = prm.gap_extend;
    P.scratch = (int *)c.scratch.d;
    P.scratch_stride = per_slot / sizeof(int);
    P.trace = c.trace.d;
    P.scores = (uint32_t *)c.scores.d;
    P.counter = (uint32_t *)c.counter.d;
    P.out_r = (uint32_t *)c.out_r.d;
    P.out_q = (uint32_t *)c.out_q.d;
    P.n_pairs = (uint32_t *)c.n_pairs.d;
    P.status = (int32_t *)c.status.d;

    tl_dp_kernel = use_cluster ? 4 : use_packed ? 3 : use_warp ? 2 : 1;
    CK(cudaEventRecord(c.ev0, s));
    if (use_cluster) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster_ctas, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(32), cfg.dynamicSmemBytes = PK_CLUSTER_SMEM, cfg.stream = s, cfg.attrs = at, cfg.numAttrs = 1;
        cfg.gridDim = dim3(cluster_ctas);
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, dp_fill_packed_kernel<true, true>, &cfg) != cudaSuccess || max_clusters < 1) max_clusters = 1;
        const unsigned n_clusters = (unsigned)std::min<uint64_t>(n, (uint64_t)max_clusters);
        cfg.gridDim = dim3(n_clusters * cluster_ctas);
        CK(cudaLaunchKernelEx(&cfg, dp_fill_packed_kernel<true, true>, P));
    } else if (use_packed) {
        const unsigned grid = (unsigned)(slots / PK_WARPS), block = PK_WARPS * 32;
        if (scores_only) dp_fill_packed_kernel<false, false><<<grid, block, PK_SMEM_BYTES, s>>>(P);
        else dp_fill_packed_kernel<true, false><<<grid, block, PK_SMEM_BYTES, s>>>(P);
    } else if (use_warp) {
        const unsigned grid = (unsigned)(slots / WARP_KERNEL_WARPS), block = WARP_KERNEL_WARPS * 32;
        if (scores_only) dp_fill_warp_kernel<false><<<grid, block, 0, s>>>(P);
        else dp_fill_warp_kernel<true><<<grid, block, 0, s>>>(P);
    } else {
        if (wide) dp_fill_kernel<uint32_t><<<(unsigned)slots, FILL_THREADS, 0, s>>>(P);
        else dp_fill_kernel<uint16_t><<<(unsigned)slots, FILL_THREADS, 0, s>>>(P);
    }
    if (scores_only) {
        CK(cudaEventRecord(c.ev1, s));
        CK(cudaGetLastError());
        if (!c.h_scores.reserve(n * 4)) return POASTA_ERR_CUDA;
        CK(cudaMemcpyAsync(c.h_scores.h, c.scores.d, n * 4, cudaMemcpyDeviceToHost, s));
        CK(wait_stream(c));
        float ms0 = 0;
        cudaEventElapsedTime(&ms0, c.ev0, c.ev1);
        tl_kernel_ms += ms0;
        tl_d2h += (uint64_t)n * 4;
        tl_launches += 1;
        tl_cells += pb.total_cells;
        res.scores.assign((uint32_t *)c.h_scores.h, (uint32_t *)c.h_scores.h + n);
        res.n_pairs.assign(n, 0);
        res.offsets.assign(n + 1, 0);
        res.status.assign(n, 0);
        for (uint32_t i = 0; i < n; ++i)
            if (res.scores[i] >= (uint32_t)INF) res.status[i] = POASTA_ERR_NO_ALIGNMENT;
        res.dense_r = res.dense_q = nullptr;
        return 0;
    }
    const unsigned tb_blocks = (n + 127) / 128;
    if (use_packed || use_cluster) traceback_rows_kernel<<<(n + 3) / 4, 128, 0, s>>>(P);  // one warp per alignment
    else if (use_warp) traceback_planes_kernel<<<tb_blocks, 128, 0, s>>>(P);
    else if (wide) traceback_kernel<uint32_t><<<tb_blocks, 128, 0, s>>>(P);
    else traceback_kernel<uint16_t><<<tb_blocks, 128, 0, s>>>(P);
    scan_kernel<<<1, 1024, 0, s>>>(P.n_pairs, (uint32_t *)c.offsets.d, n);
    CK(cudaEventRecord(c.ev1, s));
    CK(cudaGetLastError());
    // first the small arrays (we need the total pair count to size the dense copy)
    if (!c.h_scores.reserve(n * 4) || !c.h_npairs.reserve(n * 4) || !c.h_status.reserve(n * 4) || !c.h_offsets.reserve((n + 1) * 4))
        return POASTA_ERR_CUDA;
    CK(cudaMemcpyAsync(c.h_offsets.h, c.offsets.d, (n + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_scores.h, c.scores.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_npairs.h, c.n_pairs.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_status.h, c.status.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(wait_stream(c));
    const uint32_t total_pairs = ((uint32_t *)c.h_offsets.h)[n];
    if (!c.dense_r.reserve((size_t)total_pairs * 4 + 4) || !c.dense_q.reserve((size_t)total_pairs * 4 + 4) ||
        !c.h_dense_r.reserve((size_t)total_pairs * 4 + 4) || !c.h_dense_q.reserve((size_t)total_pairs * 4 + 4))
        return POASTA_ERR_CUDA;
    compact_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(P, (const uint32_t *)c.offsets.d, (uint32_t *)c.dense_r.d,
                                                       (uint32_t *)c.dense_q.d);
    CK(cudaGetLastError());
    if (total_pairs) {
        CK(cudaMemcpyAsync(c.h_dense_r.h, c.dense_r.d, (size_t)total_pairs * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c.h_dense_q.h, c.dense_q.d, (size_t)total_pairs * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(wait_stream(c));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    tl_kernel_ms += ms;
    tl_d2h += (uint64_t)(4 * n + 1) * 4 + (uint64_t)total_pairs * 8;
    tl_launches += 4;
    tl_cells += pb.total_cells;
    res.scores.assign((uint32_t *)c.h_scores.h, (uint32_t *)c.h_scores.h + n);
    res.n_pairs.assign((uint32_t *)c.h_npairs.h, (uint32_t *)c.h_npairs.h + n);
    res.offsets.assign((uint32_t *)c.h_offsets.h, (uint32_t *)c.h_offsets.h + n + 1);
    res.status.assign((int32_t *)c.h_status.h, (int32_t *)c.h_status.h + n);
    res.dense_r = (const uint32_t *)c.h_dense_r.h;
    res.dense_q = (const uint32_t *)c.h_dense_q.h;
    return 0;
}

// memory budget for one chunk's trace store
static uint64_t trace_budget_bytes() {
    static const uint64_t forced_mb = getenv("POASTA_DP_BUDGET_MB") ? strtoull(getenv("POASTA_DP_BUDGET_MB"), nullptr, 10) : 0;  // test knob
    if (forced_mb) return forced_mb << 20;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (uint64_t)8 << 30;
    return std::max<uint64_t>((uint64_t)1 << 30, std::min<uint64_t>((uint64_t)48 << 30, total_b / 3)) / tl_lane_count;
}

struct Problem {
    const poasta_graph_view *graph;
    uint32_t graph_key;  // problems with the same key share one packed graph
    const uint8_t *query;
    uint32_t qlen;
    const SearchIndex *sidx;  // optional prebuilt index for the order-exact search (built on demand when null)
    const PackedGraph *packed = nullptr;               // optional prebuilt rank-ordered layout (DP kernels)
    const std::vector<replay::NodeRec> *recs = nullptr;  // optional prebuilt node records (search kernel)
};

// DP path: aligns all problems; calls sink(i, score, status, rpos*, qpos*, n_pairs) per problem in order.
template <class Sink>
static int dp_problems(int device, const std::vector<Problem> &probs, const poasta_params &prm, Sink &&sink, bool scores_only = false) {
    DeviceContext &c = ctx_for(device);
    int rc = c.init(device);
    if (rc) return rc;
    if (prm.mode != 0) return POASTA_ERR_UNSUPPORTED;  // ends-free is an order property of the reference's search, not a DP optimum
    const uint64_t budget = trace_budget_bytes();
    // kept per host thread: a lane packs ~75 MB per round, and fresh vectors would be re-grown and page-faulted in every round
    static thread_local PackedBatch pb;
    std::vector<uint32_t> rank_scratch;
    ChunkResult res;
    size_t i = 0;
    const size_t n = probs.size();
    std::vector<size_t> idx;
    while (i < n) {
        pb.clear();
        idx.clear();
        uint32_t last_key = 0xFFFFFFFFu, last_gid = 0;
        uint64_t chunk_bytes = 0, merge_rows = 0;
        uint32_t budget_key = 0;
        bool have_budget_key = false;
        while (i < n) {
            const Problem &p = probs[i];
            const uint64_t V = p.graph->n_nodes;
            if (V <= 2 || p.qlen == 0) {  // aligner/mod.rs:124-142 empty-graph shortcut; empty query => empty alignment
                idx.push_back(i | ((size_t)1 << 62));  // handled on the host below
                ++i;
                continue;
            }
            // what run_chunk will allocate for this alignment, whichever kernel it picks: the CTA kernel keeps one trace word per
            // cell (2 or 4 bytes), the warp kernel 1 KB of bit planes per row plus a 4 KB ordinal block per row with >= 2
            // predecessors, independent of the query length (so short queries are NOT cheap there)
            if (p.graph_key != budget_key || !have_budget_key) {
                merge_rows = 0;
                for (uint32_t v = 0; v < p.graph->n_nodes; ++v) merge_rows += p.graph->pred_off[v + 1] - p.graph->pred_off[v] >= 2;
                budget_key = p.graph_key, have_budget_key = true;
            }
            const uint64_t add = std::max<uint64_t>(V * ((p.qlen + 1 + 31) & ~31ull) * 4, V * 1024 + merge_rows * 4096);  // packed rows: 4 B per cell, 32-column lanes
            if (!pb.alns.empty() && (chunk_bytes + add > budget || pb.alns.size() >= (1u << 20))) break;
            chunk_bytes += add;
            if (p.graph_key != last_key || pb.graphs.empty()) {
                last_gid = p.packed ? pb.add_packed(*p.packed) : pb.add_graph(*p.graph, rank_scratch);
                last_key = p.graph_key;
            }
            pb.add_alignment(last_gid, p.query, p.qlen);
            idx.push_back(i);
            ++i;
        }
        rc = run_chunk(c, pb, prm, res, scores_only);
        if (rc) return rc;
        size_t k = 0;
        for (size_t t = 0; t < idx.size(); ++t) {
            if (idx[t] >> 62) {
                const size_t pi = idx[t] & (((size_t)1 << 62) - 1);
                const Problem &p = probs[pi];
                uint32_t score = 0;
                if (p.graph->n_nodes <= 2) score = p.qlen * 4;  // "rough cost estimate" of the reference (mod.rs:131-133)
                // non-empty graph, empty query: global score is the cheapest all-deletion path; take it from a
                // one-column DP on the host? It is a degenerate case the reference handles through A*: not batched here.
                if (p.graph->n_nodes > 2) {
                    sink(pi, 0u, (int32_t)POASTA_ERR_UNSUPPORTED, (const uint32_t *)nullptr, (const uint32_t *)nullptr, 0u);
                } else {
                    sink(pi, score, 0, (const uint32_t *)nullptr, (const uint32_t *)nullptr, 0u);
                }
                continue;
            }
            sink(idx[t], res.scores[k], res.status[k], res.dense_r ? res.dense_r + res.offsets[k] : nullptr,
                 res.dense_q ? res.dense_q + res.offsets[k] : nullptr, res.n_pairs[k]);
            ++k;
        }
    }
    return 0;
}

// ---- order-exact search path ------------------------------------------------------------------------
static int run_replay(DeviceContext &c, const PackedReplay &pr, const poasta_params &prm, ChunkResult &res,
                      std::vector<uint32_t> *stats_out, bool full_arenas) {
    const uint32_t n = (uint32_t)pr.alns.size();
    if (n == 0) return 0;
    CK(cudaSetDevice(c.device));
    const replay::ScratchDims dmax(pr.max_V, pr.max_Q, prm.mismatch, prm.gap_open, prm.gap_extend);
    auto up = [](uint64_t b) { return (b + 255) & ~(uint64_t)255; };
    const uint64_t cells_per_aln = (uint64_t)pr.max_V * dmax.W;
    // u16 score cells when no reachable score can hit 0xFFFF (all nodes deleted + all query bases inserted is the worst case)
    const uint64_t worst = ((uint64_t)pr.max_V + pr.max_Q + 4) * std::max<uint64_t>(prm.mismatch, (uint64_t)prm.gap_open + prm.gap_extend) +
                           prm.gap_open;
    const bool cell16 = worst < replay::CELL16_LIMIT;
    const uint64_t rec_bytes = cell16 ? 8 : 16;  // records {M, D, I, epoch}
    // Blocked visited table (replay_core.hpp BLK_SHIFT) when a dense V x W table would be large: reads against a big fixed graph
    // touch a sliver of it (config 5: 400 MB dense per worker = 245 workers per GPU; blocked: ~50 MB = every warp slot busy).
    // The pool holds one block per node (ends-free alignments start a state at offset 0 of every node) plus room for the search;
    // an alignment that outgrows it is rerun with the dense table (full_arenas).
    const int blocked_knob = getenv("POASTA_BLOCKED_TABLE") ? atoi(getenv("POASTA_BLOCKED_TABLE")) : -1;  // tests: 1 forces, 0 forbids
    const uint32_t bpr = (dmax.W + replay::BLK_CELLS - 1) / replay::BLK_CELLS;
    const uint64_t dense_blocks = (uint64_t)pr.max_V * bpr;
    uint64_t max_blocks = std::min<uint64_t>(replay::BMAP_BLOCK_MASK, (uint64_t)pr.max_V + std::max<uint64_t>(16384, dense_blocks / 32));
    if (const char *e = getenv("POASTA_BLOCKED_POOL_BLOCKS")) max_blocks = std::max<uint64_t>(1, std::min<uint64_t>(max_blocks, strtoull(e, nullptr, 10)));  // tests: forces the dense rerun
    const bool blocked = !full_arenas && (blocked_knob == 1 || (blocked_knob != 0 && cells_per_aln * rec_bytes > ((uint64_t)64 << 20) &&
                                                                 max_blocks * 2 < dense_blocks));
    const uint64_t plane_bytes = blocked ? up(max_blocks * replay::BLK_CELLS * rec_bytes) : up(cells_per_aln * rec_bytes);
    // the live queue is a small fraction of the table: size the item arena for that and retry the rare overflow at full size
    // measured high-water mark of the live queue: <= 1 % of the cells (16k items at V = 1770, Q = 1000)
    // (the inadmissible path heuristic on divergent input re-queues states: high-water 0.99 x cells seen in tools/fuzz_core_vs_oracle.py)
    uint64_t max_items = full_arenas ? 262144 + 2 * cells_per_aln : 32768 + cells_per_aln / (blocked ? 128 : 32);
    max_items = (max_items + 7) & ~(uint64_t)7;  // whole chunks (replay_core.hpp CHUNK)
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // in reference-exact mode the DP pass keeps no trace store, so most of HBM can hold search slabs
    static const double slab_fraction = getenv("POASTA_SLAB_FRACTION") ? atof(getenv("POASTA_SLAB_FRACTION")) : 0.6;
    const uint64_t budget = std::max<uint64_t>((uint64_t)2 << 30, (uint64_t)(slab_fraction * (double)total_b)) / tl_lane_count;
    uint64_t slab = 0, off_qhead = 0, off_items = 0, off_free = 0, off_reached = 0, off_bits = 0, off_dfa = 0, off_bmap = 0;
    off_qhead = SLAB_CELLS + plane_bytes;
    off_items = off_qhead + up((uint64_t)dmax.f_max * 12);
    off_free = off_items + up(max_items * 16);
    off_reached = off_free + up(max_items * 4);                   // reached bounds: 2 words per node
    off_bits = off_reached + up((uint64_t)pr.max_V * 8);          // reached-offset bitmap: Wb words per node
    off_dfa = off_bits + up((uint64_t)pr.max_V * ((dmax.W + 31) / 32) * 4);
    off_bmap = off_dfa + up((uint64_t)(pr.max_V + 2) * 24);       // DFA frames: 6 words
    slab = off_bmap + (blocked ? up(dense_blocks * 4) : 0);      // block map of the blocked table
    const uint64_t want_workers = std::min<uint64_t>(n, (uint64_t)c.n_sms * 32);  // 32 CTAs of one warp per SM
    uint64_t workers = std::max<uint64_t>(1, std::min<uint64_t>(want_workers, budget / slab));
    const unsigned warps_per_cta = 1;  // one warp per CTA (replay_kernel.cuh)
    workers = ((workers + warps_per_cta - 1) / warps_per_cta) * warps_per_cta;
    if (!c.r_slabs.reserve(workers * slab)) {
        workers = warps_per_cta;
        if (!c.r_slabs.reserve(workers * slab)) return POASTA_ERR_CUDA;
    }
    {
        const uint64_t sig = slab * 1000003ull + off_qhead * 10007ull + off_items * 101ull + off_dfa + (cell16 ? 1ull << 62 : 0) +
                             (blocked ? (1ull << 61) + bpr * 7919ull + max_blocks * 31ull : 0);
        if (c.slab_ptr != c.r_slabs.d || c.slab_sig != sig) c.slab_ptr = c.r_slabs.d, c.slab_sig = sig, c.slab_zeroed_bytes = 0;
        if (workers * slab > c.slab_zeroed_bytes) {
            CK(cudaMemsetAsync((uint8_t *)c.r_slabs.d + c.slab_zeroed_bytes, 0, workers * slab - c.slab_zeroed_bytes, c.stream));
            c.slab_zeroed_bytes = workers * slab;
        }
    }
    bool ok = c.r_graphs.reserve(pr.graphs.size() * sizeof(RGraphDesc)) && c.alns.reserve(n * sizeof(AlnDesc)) &&
              c.r_nodes.reserve(pr.nodes.size() * sizeof(replay::NodeRec) + 32) && c.queries.reserve(pr.queries.size() + 4) &&
              c.r_succ.reserve(pr.succ.size() * 4 + 4) && c.r_pred.reserve(pr.pred.size() * 4 + 4) &&
              c.r_bub.reserve(pr.bub.size() * 4 + 4) && c.r_path_off.reserve(pr.path_off.size() * 4 + 4) &&
              c.r_path_rem.reserve(pr.path_rem.size() * 4 + 4) && c.scores.reserve(n * 4) && c.counter.reserve(4) &&
              c.out_r.reserve(pr.out_elems * 4) && c.out_q.reserve(pr.out_elems * 4) && c.n_pairs.reserve(n * 4) &&
              c.status.reserve(n * 4) && c.offsets.reserve((n + 1) * 4) && c.r_stats.reserve((size_t)n * 16);
    if (!ok) return POASTA_ERR_CUDA;
    cudaStream_t s = c.stream;
    auto h2d = [&](Buf &b, const void *src, size_t bytes) -> cudaError_t {
        tl_h2d += bytes;
        return bytes ? cudaMemcpyAsync(b.d, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    CK(h2d(c.r_graphs, pr.graphs.data(), pr.graphs.size() * sizeof(RGraphDesc)));
    CK(h2d(c.alns, pr.alns.data(), n * sizeof(AlnDesc)));
    CK(h2d(c.r_nodes, pr.nodes.data(), pr.nodes.size() * sizeof(replay::NodeRec)));
    CK(h2d(c.queries, pr.queries.data(), pr.queries.size()));
    CK(h2d(c.r_succ, pr.succ.data(), pr.succ.size() * 4));
    CK(h2d(c.r_pred, pr.pred.data(), pr.pred.size() * 4));
    CK(h2d(c.r_bub, pr.bub.data(), pr.bub.size() * 4));
    CK(h2d(c.r_path_off, pr.path_off.data(), pr.path_off.size() * 4));
    CK(h2d(c.r_path_rem, pr.path_rem.data(), pr.path_rem.size() * 4));
    CK(cudaMemsetAsync(c.counter.d, 0, 4, s));

    ReplayDev B;
    B.graphs = (const RGraphDesc *)c.r_graphs.d;
    B.alns = (const AlnDesc *)c.alns.d;
    B.nodes = (const replay::NodeRec *)c.r_nodes.d;
    B.queries = (const uint8_t *)c.queries.d;
    B.succ = (const uint32_t *)c.r_succ.d;
    B.pred = (const uint32_t *)c.r_pred.d;
    B.bub = (const uint32_t *)c.r_bub.d;
    B.path_off = (const uint32_t *)c.r_path_off.d;
    B.path_rem = (const uint32_t *)c.r_path_rem.d;
    B.n_alns = n;
    B.prm = replay::Params{prm.mismatch, prm.gap_open, prm.gap_extend, prm.mode, prm.heuristic};
    B.slabs = (uint8_t *)c.r_slabs.d;
    B.slab_bytes = slab;
    B.off_qhead = off_qhead, B.off_items = off_items, B.off_free = off_free, B.off_rb = off_reached, B.off_bits = off_bits, B.off_dfa = off_dfa;
    B.max_items = (uint32_t)max_items;
    B.off_bmap = off_bmap, B.bpr = bpr, B.max_blocks = (uint32_t)max_blocks;
    tl_search_table = blocked ? 2 : 1;
    B.counter = (uint32_t *)c.counter.d;
    B.scores = (uint32_t *)c.scores.d;
    B.n_pairs = (uint32_t *)c.n_pairs.d;
    B.out_r = (uint32_t *)c.out_r.d;
    B.out_q = (uint32_t *)c.out_q.d;
    B.status = (int32_t *)c.status.d;
    B.stats = (uint32_t *)c.r_stats.d;

    BatchDev P;  // the fields the scan / compact kernels read
    memset(&P, 0, sizeof P);
    P.alns = B.alns, P.n_alns = n, P.out_r = B.out_r, P.out_q = B.out_q, P.n_pairs = B.n_pairs;

    CK(cudaEventRecord(c.ev0, s));
    {
        // 32 resident one-warp CTAs per SM -> 64 registers, 32 workers per SM (replay_kernel.cuh)
        const dim3 grid((unsigned)workers), block(32);
        auto launch = [&](auto cell, auto heur) {
            using CellT = decltype(cell);
            constexpr int H = decltype(heur)::value;
            if (blocked) replay_kernel<CellT, 32, H, true><<<grid, block, 0, s>>>(B);
            else replay_kernel<CellT, 32, H, false><<<grid, block, 0, s>>>(B);
        };
        auto launch_cell = [&](auto cell) {
            if (prm.heuristic == 0) launch(cell, std::integral_constant<int, 0>{});
            else if (prm.heuristic == 1) launch(cell, std::integral_constant<int, 1>{});
            else launch(cell, std::integral_constant<int, 2>{});
        };
        if (cell16) launch_cell(uint16_t{});
        else launch_cell(uint32_t{});
    }
    scan_kernel<<<1, 1024, 0, s>>>(B.n_pairs, (uint32_t *)c.offsets.d, n);
    CK(cudaEventRecord(c.ev1, s));
    CK(cudaGetLastError());
    if (!c.h_scores.reserve(n * 4) || !c.h_npairs.reserve(n * 4) || !c.h_status.reserve(n * 4) || !c.h_offsets.reserve((n + 1) * 4) ||
        !c.h_stats.reserve((size_t)n * 16))
        return POASTA_ERR_CUDA;
    CK(cudaMemcpyAsync(c.h_offsets.h, c.offsets.d, (n + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_scores.h, c.scores.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_npairs.h, c.n_pairs.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_status.h, c.status.d, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c.h_stats.h, c.r_stats.d, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    CK(wait_stream(c));
    const uint32_t total_pairs = ((uint32_t *)c.h_offsets.h)[n];
    if (!c.dense_r.reserve((size_t)total_pairs * 4 + 4) || !c.dense_q.reserve((size_t)total_pairs * 4 + 4) ||
        !c.h_dense_r.reserve((size_t)total_pairs * 4 + 4) || !c.h_dense_q.reserve((size_t)total_pairs * 4 + 4))
        return POASTA_ERR_CUDA;
    compact_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(P, (const uint32_t *)c.offsets.d, (uint32_t *)c.dense_r.d,
                                                       (uint32_t *)c.dense_q.d);
    CK(cudaGetLastError());
    if (total_pairs) {
        CK(cudaMemcpyAsync(c.h_dense_r.h, c.dense_r.d, (size_t)total_pairs * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c.h_dense_q.h, c.dense_q.d, (size_t)total_pairs * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(wait_stream(c));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    tl_search_ms += ms;
    tl_d2h += (uint64_t)(8 * n + 1) * 4 + (uint64_t)total_pairs * 8;
    tl_launches += 3;
    res.scores.assign((uint32_t *)c.h_scores.h, (uint32_t *)c.h_scores.h + n);
    res.n_pairs.assign((uint32_t *)c.h_npairs.h, (uint32_t *)c.h_npairs.h + n);
    res.offsets.assign((uint32_t *)c.h_offsets.h, (uint32_t *)c.h_offsets.h + n + 1);
    res.status.assign((int32_t *)c.h_status.h, (int32_t *)c.h_status.h + n);
    res.dense_r = (const uint32_t *)c.h_dense_r.h;
    res.dense_q = (const uint32_t *)c.h_dense_q.h;
    if (stats_out) stats_out->assign((uint32_t *)c.h_stats.h, (uint32_t *)c.h_stats.h + (size_t)n * 4);
    static const bool timing = getenv("POASTA_TIMING") != nullptr;
    if (timing) {  // a launch lasts as long as its slowest alignment: show how uneven the alignments of one launch are
        const uint32_t *st = (const uint32_t *)c.h_stats.h;
        uint64_t sum = 0, mx = 0;
        for (size_t i = 0; i < n; ++i) sum += st[4 * i + 1], mx = std::max<uint64_t>(mx, st[4 * i + 1]);
        fprintf(stderr, "poasta_b200 lane %d: search launch of %zu alignments, %u workers x %.1f MB (%s table), %.1f ms; visited states mean %.0f max %llu\n",
                tl_lane, (size_t)n, (unsigned)workers, slab / 1048576.0, blocked ? "blocked" : "dense", ms, (double)sum / (double)n, (unsigned long long)mx);
    }
    return 0;
}

static int map_replay_status(int st) {
    switch (st) {
        case replay::RP_OK: return POASTA_OK;
        case replay::RP_EMPTY_QUEUE:
        case replay::RP_NO_BACKTRACE:
        case replay::RP_SCORE_WRAP: return POASTA_ERR_NO_ALIGNMENT;  // the reference panics in these cases
        default: return POASTA_ERR_UNSUPPORTED;                    // scratch arena exhausted
    }
}

template <class Sink>
static int replay_problems(int device, const std::vector<Problem> &probs, const poasta_params &prm, Sink &&sink) {
    DeviceContext &c = ctx_for(device);
    int rc = c.init(device);
    if (rc) return rc;
    struct Out {
        uint32_t score = 0;
        int32_t status = 0;
        std::vector<uint32_t> r, q;
    };
    tl_search_stats.assign(probs.size() * 3, 0);
    std::vector<uint32_t> launch_stats;
    std::vector<Out> outs(probs.size());
    std::vector<size_t> todo;
    for (size_t i = 0; i < probs.size(); ++i) {
        if (probs[i].graph->n_nodes <= 2) outs[i].score = probs[i].qlen * 4u;  // empty-graph shortcut (aligner/mod.rs:124-142)
        else todo.push_back(i);
    }
    static thread_local PackedReplay pr;  // kept per host thread like the DP batch (~160 MB per round and lane)
    ChunkResult res;
    SearchIndex tmp;
    for (int attempt = 0; attempt < 2 && !todo.empty(); ++attempt) {
        pr.clear();
        uint32_t last_key = 0xFFFFFFFFu, last_gid = 0;
        for (size_t i : todo) {
            const Problem &p = probs[i];
            if (p.graph_key != last_key || pr.graphs.empty()) {
                const SearchIndex *si = p.sidx;
                if (!si) {
                    tmp.build(*p.graph, prm.heuristic == 2);
                    si = &tmp;
                }
                last_gid = pr.add_graph(*p.graph, *si, p.recs);
                last_key = p.graph_key;
            }
            pr.add_alignment(last_gid, p.query, p.qlen);
        }
        rc = run_replay(c, pr, prm, res, &launch_stats, attempt == 1);
        if (rc) return rc;
        std::vector<size_t> again;
        for (size_t k = 0; k < todo.size(); ++k) {
            const int st = res.status[k];
            if (attempt == 0 && (st == replay::RP_OUT_OF_BLOCKS || st == replay::RP_OUT_OF_ITEMS)) {
                again.push_back(todo[k]);  // arena sized for the common case was too small: rerun with full-size arenas
                continue;
            }
            Out &o = outs[todo[k]];
            for (int t = 0; t < 3; ++t) tl_search_stats[todo[k] * 3 + t] = launch_stats[4 * k + t];
            o.score = res.scores[k];
            o.status = map_replay_status(st);
            o.r.assign(res.dense_r + res.offsets[k], res.dense_r + res.offsets[k] + res.n_pairs[k]);
            o.q.assign(res.dense_q + res.offsets[k], res.dense_q + res.offsets[k] + res.n_pairs[k]);
        }
        todo.swap(again);
    }
    for (size_t i = 0; i < probs.size(); ++i) sink(i, outs[i].score, outs[i].status, outs[i].r.data(), outs[i].q.data(), (uint32_t)outs[i].r.size());
    return 0;
}

// Dispatch on poasta_params::tiebreak (include/poasta_b200.h).
template <class Sink>
static int align_problems(int device, const std::vector<Problem> &probs, const poasta_params &prm, Sink &&sink) {
    if (prm.tiebreak == 1) {
        // an EMPTY query against a non-empty graph has no DP table to fill; the reference aligns it (all deletions along the
        // cheapest path, empty alignment), and so does the search kernel: those problems take that path in DP mode too
        std::vector<size_t> q0, rest;
        for (size_t i = 0; i < probs.size(); ++i) (probs[i].qlen == 0 && probs[i].graph->n_nodes > 2 ? q0 : rest).push_back(i);
        if (q0.empty()) return dp_problems(device, probs, prm, sink);
        std::vector<Problem> a, b;
        for (size_t i : q0) a.push_back(probs[i]);
        for (size_t i : rest) b.push_back(probs[i]);
        struct Held {
            uint32_t score;
            int32_t status;
            std::vector<uint32_t> r, q;
        };
        std::vector<Held> held(probs.size());
        auto keep = [&](const std::vector<size_t> &map) {
            return [&held, &map](size_t i, uint32_t score, int32_t status, const uint32_t *r, const uint32_t *q, uint32_t np) {
                Held &h = held[map[i]];
                h.score = score, h.status = status;
                h.r.assign(r, r + np), h.q.assign(q, q + np);
            };
        };
        int rc = replay_problems(device, a, prm, keep(q0));
        if (rc) return rc;
        tl_search_stats.clear();
        rc = dp_problems(device, b, prm, keep(rest));
        if (rc) return rc;
        for (size_t i = 0; i < probs.size(); ++i) sink(i, held[i].score, held[i].status, held[i].r.data(), held[i].q.data(), (uint32_t)held[i].r.size());
        return 0;
    }
    if (prm.mode != 0) return replay_problems(device, probs, prm, sink);  // ends-free exists only as a search-order property
    // reference-exact, global: the DP kernels give the optimal score of every problem, the search gives the alignment;
    // the two scores must agree (device-side self check of both paths)
    std::vector<uint32_t> dp_score(probs.size(), 0);
    std::vector<int32_t> dp_status(probs.size(), 0);
    poasta_params dp_prm = prm;
    dp_prm.tiebreak = 1;
    bool any_q0 = getenv("POASTA_SKIP_DP_CHECK") != nullptr;  // debugging aid: run the search alone
    for (const Problem &p : probs) any_q0 |= (p.qlen == 0);
    int rc = 0;
    if (!any_q0) {
        rc = dp_problems(device, probs, dp_prm, [&](size_t i, uint32_t score, int32_t status, const uint32_t *, const uint32_t *, uint32_t) {
            dp_score[i] = score;
            dp_status[i] = status;
        }, /*scores_only=*/true);
        if (rc) return rc;
    }
    return replay_problems(device, probs, prm,
                           [&](size_t i, uint32_t score, int32_t status, const uint32_t *r, const uint32_t *q, uint32_t np) {
                               // -H path is not admissible (aligner/heuristic.rs:128-184): the reference itself may return a
                               // score above the optimum there, and so does the search; the DP optimum is then a lower bound
                               // The DP optimum is a lower bound on what the reference returns, not always equal to it: the
                               // reference's offset-0 special case (dfa.rs:145-168) can hide the optimal first match after a
                               // leading deletion (2 of 190,000 alignments of config 2 return optimum + 2), and -H path is not
                               // admissible. The search reproduces the reference; a search score BELOW the DP optimum is a bug.
                               if (!any_q0 && status == 0 && dp_status[i] == 0 && dp_score[i] != score && getenv("POASTA_SELFCHECK_VERBOSE"))
                                   fprintf(stderr, "poasta_b200: problem %zu (V=%u, Q=%u): reference score %u above the DP optimum %u\n", i,
                                           probs[i].graph->n_nodes, probs[i].qlen, score, dp_score[i]);
                               if (!any_q0 && status == 0 && dp_status[i] == 0 && dp_score[i] > score) {
                                   static const char *mode = getenv("POASTA_SELFCHECK");  // "warn": report and continue
                                   fprintf(stderr, "poasta_b200 self-check: problem %zu (graph key %u, V=%u, Q=%u): DP score %u, search score %u\n",
                                           i, probs[i].graph_key, probs[i].graph->n_nodes, probs[i].qlen, dp_score[i], score);
                                   if (!(mode && !strcmp(mode, "warn"))) status = POASTA_ERR_INTERNAL;
                               }
                               sink(i, score, status, r, q, np);
                           });
}

}  // namespace poasta

using namespace poasta;

// ---- integer issue peak of this device (the DP fill is bound by integer issue, and that peak is not in
// MEASURED_PEAKS.json). Each thread runs 8 independent dependent-chains so the pipes stay full; ops = lane-operations.
template <int MODE>  // 0: add, 1: min, 2: fused add+min (DPX __viaddmin_s32, counted as 2 operations)
__global__ void __launch_bounds__(256) int_peak_kernel(int *out, int iters, int seed) {
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 8 + k;
    const int b = seed | 1, c = seed + 12345;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) a[k] = a[k] + a[(k + 1) & 7];  // cross-chain adds cannot be folded into a multiply
                else if (MODE == 1) a[k] = min(a[k] ^ u, c + k);  // the xor keeps the compiler from collapsing the chain
                else a[k] = __viaddmin_s32(a[k], b, c + u);
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k];
    if (r == 0x7FFFFFFF) out[blockIdx.x * blockDim.x + threadIdx.x] = r;  // practically never: keeps the work live
}

// Worker threads that live as long as the object; run(n, fn) hands out index blocks of 16 and returns when all are done.
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<void(size_t)> fn;
    std::atomic<size_t> next{0};
    size_t n = 0, generation = 0, busy = 0;
    bool stop = false;
    explicit HostPool(unsigned n_threads) {
        for (unsigned t = 1; t < n_threads; ++t) threads.emplace_back([this]() { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
    void drain() {
        for (;;) {
            const size_t i0 = next.fetch_add(16);
            if (i0 >= n) break;
            for (size_t i = i0; i < std::min(n, i0 + 16); ++i) fn(i);
        }
    }
    void loop() {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&]() { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--busy == 0) cv_done.notify_all();
            }
        }
    }
    template <class F> void run(size_t count, F &&f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = std::ref(f);
            n = count;
            next = 0;
            busy = threads.size();
            ++generation;
        }
        cv_work.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&]() { return busy == 0; });
    }
};

extern "C" {

// shared tail of the two align-batch entry points: run the problems, fill the caller's result records
static int align_batch_results(int device, const std::vector<Problem> &probs, const poasta_params &params, poasta_result *results) {
    return align_problems(device, probs, params,
                          [&](size_t i, uint32_t score, int32_t status, const uint32_t *r, const uint32_t *q, uint32_t np) {
                              poasta_result &out = results[i];
                              out.score = score;
                              out.status = status;
                              out.num_cells = (uint64_t)(probs[i].graph->n_nodes - 2) * probs[i].qlen;
                              const bool have = tl_search_stats.size() == probs.size() * 3;  // empty with tiebreak = 1 (no search ran)
                              out.num_queued = have ? tl_search_stats[3 * i] : 0;
                              out.num_visited = have ? tl_search_stats[3 * i + 1] : 0;
                              out.num_pruned = have ? tl_search_stats[3 * i + 2] : 0;
                              out.n_pairs = np;
                              if (np > out.capacity) {
                                  out.status = POASTA_ERR_INVALID_ARG;
                                  return;
                              }
                              if (np) {
                                  memcpy(out.rpos, r, (size_t)np * 4);
                                  memcpy(out.qpos, q, (size_t)np * 4);
                              }
                          });
}

int poasta_align_batch(int device, const poasta_graph_view *graphs, uint32_t n_graphs, const uint8_t *const *queries,
                       const uint32_t *qlens, const uint32_t *graph_of_query, uint32_t n_queries, const poasta_params *params,
                       poasta_result *results) {
    if (!graphs || !queries || !qlens || !params || !results) return POASTA_ERR_INVALID_ARG;
    if (device < 0 || device >= MAX_DEVICES) return POASTA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> call_lock(device_call_mutex(device));
    tl_kernel_ms = 0, tl_search_ms = 0, tl_launches = 0, tl_cells = 0, tl_h2d = 0, tl_d2h = 0;
    tl_lane_max_ms = -1;
    tl_search_stats.clear();
    std::vector<Problem> probs(n_queries);
    for (uint32_t i = 0; i < n_queries; ++i) {
        const uint32_t gi = graph_of_query ? graph_of_query[i] : i;
        if (gi >= n_graphs) return POASTA_ERR_INVALID_ARG;
        probs[i] = {&graphs[gi], gi, queries[i], qlens[i], nullptr, nullptr, nullptr};
    }
    return align_batch_results(device, probs, *params, results);
}

int poasta_align_batch_graphs(int device, poasta_graph *const *graphs, uint32_t n_graphs, const uint8_t *const *queries,
                              const uint32_t *qlens, const uint32_t *graph_of_query, uint32_t n_queries, const poasta_params *params,
                              poasta_result *results) {
    if (!graphs || !queries || !qlens || !params || !results) return POASTA_ERR_INVALID_ARG;
    if (device < 0 || device >= MAX_DEVICES) return POASTA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> call_lock(device_call_mutex(device));
    tl_kernel_ms = 0, tl_search_ms = 0, tl_launches = 0, tl_cells = 0, tl_h2d = 0, tl_d2h = 0;
    tl_lane_max_ms = -1;
    tl_search_stats.clear();
    // the handles keep what is derived from a graph (CSR view, bubble / distance / path index, node records, rank-ordered layout)
    // until the graph changes, so repeated batches against the same graphs do not rebuild it (the reference rebuilds its index for
    // every single alignment, aligner/config.rs:128)
    const bool need_index = params->tiebreak == 0, need_pack = params->mode == 0;
    std::vector<poasta_graph_view> views(n_graphs);
    std::vector<uint8_t> used(n_graphs, 0);
    for (uint32_t i = 0; i < n_queries; ++i) {
        const uint32_t gi = graph_of_query ? graph_of_query[i] : i;
        if (gi >= n_graphs || !graphs[gi]) return POASTA_ERR_INVALID_ARG;
        used[gi] = 1;
    }
    for (uint32_t g = 0; g < n_graphs; ++g) {
        if (!used[g]) continue;
        views[g] = graphs[g]->flat_view();
        if (views[g].n_nodes <= 2) continue;  // empty graph: the shortcut of aligner/mod.rs:124-142 needs nothing derived
        if (need_index) graphs[g]->search_index(params->heuristic == 2);
        if (need_pack) graphs[g]->packed_graph();
    }
    std::vector<Problem> probs(n_queries);
    for (uint32_t i = 0; i < n_queries; ++i) {
        const uint32_t gi = graph_of_query ? graph_of_query[i] : i;
        poasta_graph *h = graphs[gi];
        const bool derived = views[gi].n_nodes > 2;
        probs[i] = {&views[gi], gi, queries[i], qlens[i], need_index && derived ? &h->sidx : nullptr, need_pack && derived ? &h->packed : nullptr,
                    need_index && derived ? &h->node_recs : nullptr};
    }
    return align_batch_results(device, probs, *params, results);
}

int poasta_last_batch_stats(double *kernel_ms, uint32_t *n_launches, uint64_t *cells) {
    if (kernel_ms) *kernel_ms = tl_kernel_ms;
    if (n_launches) *n_launches = tl_launches;
    if (cells) *cells = tl_cells;
    return 0;
}

int poasta_last_batch_search_ms(double *search_ms) {
    if (search_ms) *search_ms = tl_search_ms;
    return 0;
}

int poasta_last_batch_device_ms(double *device_ms) {
    if (device_ms) *device_ms = tl_lane_max_ms >= 0 ? tl_lane_max_ms : tl_kernel_ms + tl_search_ms;
    return 0;
}

int poasta_last_batch_search_counts(uint64_t *num_queued, uint64_t *num_visited, uint64_t *num_pruned) {
    if (num_queued) *num_queued = tl_engine_counts[0];
    if (num_visited) *num_visited = tl_engine_counts[1];
    if (num_pruned) *num_pruned = tl_engine_counts[2];
    return 0;
}

int poasta_last_dp_kernel(void) { return tl_dp_kernel; }
int poasta_last_msa_path(void) { return tl_msa_path; }
int poasta_last_search_table(void) { return tl_search_table; }
int poasta_last_batch_fill_ms(double *fill_ms, uint32_t *n_fill_launches) {
    if (fill_ms) *fill_ms = tl_fill_ms;
    if (n_fill_launches) *n_fill_launches = tl_fill_launches;
    return 0;
}

int poasta_last_batch_io(uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = tl_h2d;
    if (d2h_bytes) *d2h_bytes = tl_d2h;
    return 0;
}

int poasta_measure_int_peak(int device, double *add_tops, double *min_tops, double *addmin_tops) {
    if (device < 0 || device >= MAX_DEVICES) return POASTA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> call_lock(device_call_mutex(device));
    DeviceContext &c = ctx_for(device);
    int rc = c.init(device);
    if (rc) return rc;
    CK(cudaSetDevice(c.device));
    int *out = nullptr;
    const int blocks = c.n_sms * 8, iters = 64;
    CK(cudaMalloc(&out, (size_t)blocks * 256 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double tops[3] = {0, 0, 0};
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
            CK(cudaEventRecord(e0, c.stream));
            if (mode == 0) int_peak_kernel<0><<<blocks, 256, 0, c.stream>>>(out, iters, rep + 3);
            else if (mode == 1) int_peak_kernel<1><<<blocks, 256, 0, c.stream>>>(out, iters, rep + 3);
            else int_peak_kernel<2><<<blocks, 256, 0, c.stream>>>(out, iters, rep + 3);
            CK(cudaEventRecord(e1, c.stream));
            CK(cudaStreamSynchronize(c.stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = std::min(best, ms);
        }
        // mode 1 issues two instructions (xor, min) per step and counts one operation each; mode 2 counts add and min
        const double ops = (double)blocks * 256 * iters * 16 * 8 * (mode == 0 ? 1 : 2);
        tops[mode] = ops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (add_tops) *add_tops = tops[0];
    if (min_tops) *min_tops = tops[1];
    if (addmin_tops) *addmin_tops = tops[2];
    return 0;
}

// bin/poasta.rs:163-236 for one group of families: round k aligns record k of every family of the group against that
// family's current graph (one device batch), then inserts the alignments (host threads).
static int msa_group(int device, poasta_graph **graphs, const char *const *fasta, const size_t *fasta_len, uint32_t n_families,
                     const poasta_params *params, uint32_t *scores_out, const size_t *score_base, uint64_t *cells_out,
                     unsigned n_threads) {
    std::vector<std::vector<FastaRecord>> recs(n_families);
    size_t max_round = 0;
    for (uint32_t f = 0; f < n_families; ++f) {
        recs[f] = parse_fasta(fasta[f], fasta_len[f]);
        max_round = std::max(max_round, recs[f].size());
    }
    std::atomic<int> err{0};
    HostPool pool(n_threads);  // one set of worker threads for the whole group (round 1 spawned and joined threads 38 times per step and lane)
    auto parallel_for = [&](size_t n, auto &&fn) {
        if (n < 64 || n_threads <= 1) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        pool.run(n, fn);
    };
    uint64_t cells = 0;
    std::vector<Problem> probs;
    std::vector<uint32_t> fam_of_prob;
    std::vector<poasta_graph_view> views(n_families);
    static const bool timing = getenv("POASTA_TIMING") != nullptr;  // per-phase host timers of one lane (stderr)
    double t_index = 0, t_align = 0, t_insert = 0;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    struct Pending {
        uint32_t score;
        int32_t status;
        std::vector<uint32_t> r, q;
    };
    std::vector<Pending> pending;
    for (size_t round = 0; round < max_round; ++round) {
        probs.clear();
        fam_of_prob.clear();
        // families whose graph is still empty take the record as the first sequence (bin/poasta.rs:209-211)
        std::vector<uint32_t> to_align;
        for (uint32_t f = 0; f < n_families; ++f) {
            if (round >= recs[f].size()) continue;
            if (graphs[f]->g.is_empty()) {
                const FastaRecord &rec = recs[f][round];
                graphs[f]->flat_valid = false;
                GraphError ge = graphs[f]->g.add_alignment_with_weights(rec.name, (const uint8_t *)rec.seq.data(), rec.seq.size(), false, nullptr, 0);
                if (ge != GraphError::Ok) return POASTA_ERR_GRAPH_CYCLE;
                if (scores_out) scores_out[score_base[f] + round] = 0;
            } else {
                to_align.push_back(f);
            }
        }
        if (to_align.empty()) continue;
        const bool need_index = params->tiebreak == 0;
        const double t0 = now();
        const bool need_pack = params->mode == 0;  // the DP kernels run in both tiebreak modes of global alignment
        parallel_for(to_align.size(), [&](size_t k) {
            poasta_graph *h = graphs[to_align[k]];
            views[to_align[k]] = h->flat_view();
            if (need_index) h->search_index(params->heuristic == 2);
            if (need_pack) h->packed_graph();
        });
        for (uint32_t f : to_align) {
            const FastaRecord &rec = recs[f][round];
            probs.push_back({&views[f], f, (const uint8_t *)rec.seq.data(), (uint32_t)rec.seq.size(),
                             need_index ? &graphs[f]->sidx : nullptr, need_pack ? &graphs[f]->packed : nullptr,
                             need_index ? &graphs[f]->node_recs : nullptr});
            fam_of_prob.push_back(f);
            cells += (uint64_t)graphs[f]->g.node_count() * rec.seq.size();
        }
        pending.assign(probs.size(), Pending{});
        const double t1 = now();
        int rc = align_problems(device, probs, *params,
                                [&](size_t i, uint32_t score, int32_t status, const uint32_t *r, const uint32_t *q, uint32_t np) {
                                    Pending &p = pending[i];
                                    p.score = score;
                                    p.status = status;
                                    p.r.assign(r, r + np);
                                    p.q.assign(q, q + np);
                                });
        if (rc) return rc;
        const double t2 = now();
        parallel_for(probs.size(), [&](size_t i) {
            const uint32_t f = fam_of_prob[i];
            const FastaRecord &rec = recs[f][round];
            if (pending[i].status) {
                err = pending[i].status;
                return;
            }
            graphs[f]->flat_valid = false;
            std::vector<AlignedPair> aln(pending[i].r.size());
            for (size_t t = 0; t < aln.size(); ++t) aln[t] = {pending[i].r[t], pending[i].q[t]};
            GraphError ge = graphs[f]->g.add_alignment_with_weights(rec.name, (const uint8_t *)rec.seq.data(), rec.seq.size(), true,
                                                                    aln.data(), aln.size());
            if (ge != GraphError::Ok) err = ge == GraphError::InvalidAlignment ? POASTA_ERR_INVALID_ALIGNMENT : POASTA_ERR_GRAPH_CYCLE;
            if (scores_out) scores_out[score_base[f] + round] = pending[i].score;
        });
        if (err) return err;
        const double t3 = now();
        t_index += t1 - t0, t_align += t2 - t1, t_insert += t3 - t2;
    }
    if (timing)
        fprintf(stderr, "poasta_b200 lane %d: flatten+index %.2f s, align (pack + copies + kernels) %.2f s of which kernels %.2f s, insert %.2f s\n",
                tl_lane, t_index, t_align, (tl_kernel_ms + tl_search_ms) / 1e3, t_insert);
    *cells_out = cells;
    return 0;
}

// Families are independent, so the batch is split into pipeline lanes (host thread + stream + device buffers each):
// while one lane inserts alignments and packs its next round on the host, the kernels of the other lanes keep the GPU busy.
int poasta_msa_batch_into(int device, poasta_graph **graphs, const char *const *fasta, const size_t *fasta_len,
                          uint32_t n_families, const poasta_params *params, uint32_t *scores_out, uint64_t *cells_out) {
    if (!graphs || !fasta || !fasta_len || !params) return POASTA_ERR_INVALID_ARG;
    if (device < 0 || device >= MAX_DEVICES) return POASTA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> call_lock(device_call_mutex(device));
    tl_kernel_ms = 0, tl_search_ms = 0, tl_launches = 0, tl_cells = 0, tl_h2d = 0, tl_d2h = 0;
    static const bool dp_host_only = getenv("POASTA_DP_HOST") != nullptr || getenv("POASTA_DP_KERNEL") != nullptr;
    tl_msa_path = 0;
    if (params->tiebreak == 0 || (params->mode == 0 && !dp_host_only)) {
        // reference-exact builds run entirely on the device (msa_engine.cuh): sequences in, finished graphs out; so do DP-only
        // builds whose inputs the packed kernels cover (msa_dp_device.cuh) - the others fall through to the per-round host pipeline
        DeviceContext &c = ctx_for(device);
        int rc = c.init(device);
        if (rc) return rc;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned host_threads = std::min(16u, hw);
        if (const char *e = getenv("POASTA_HOST_THREADS")) host_threads = std::max(1, atoi(e));
        std::vector<std::vector<FastaRecord>> recs(n_families);
        {
            std::atomic<uint32_t> cursor{0};
            auto parse = [&]() {
                for (;;) {
                    const uint32_t f = cursor.fetch_add(1);
                    if (f >= n_families) break;
                    recs[f] = parse_fasta(fasta[f], fasta_len[f]);
                }
            };
            if (n_families < 64 || host_threads <= 1) {
                parse();
            } else {
                std::vector<std::thread> pool;
                for (unsigned t = 0; t < host_threads; ++t) pool.emplace_back(parse);
                for (auto &t : pool) t.join();
            }
        }
        std::vector<EngineFamilyIn> in(n_families);
        size_t base = 0;
        for (uint32_t f = 0; f < n_families; ++f) {
            in[f] = EngineFamilyIn{graphs[f], &recs[f], scores_out ? scores_out + base : nullptr};
            base += recs[f].size();
        }
        if (params->tiebreak == 0 || dp_device_eligible(in, *params)) {
            constexpr int DP_LANES_MAX = 4;
            static EngineBuffers ebs[MAX_DEVICES][DP_LANES_MAX];
            static DpBuffers dbs[MAX_DEVICES][DP_LANES_MAX];
            EngineStats st;
            if (params->tiebreak == 0) {
                for (int l = 0; l < DP_LANES_MAX; ++l) dbs[device][l].trace.release();  // the DP-only mode's row stores would take the memory the worker slabs are sized for
                rc = msa_engine_run(c, ebs[device][0], dbs[device][0], in, *params, st, host_threads);
                tl_search_ms = st.kernel_ms, tl_lane_max_ms = st.kernel_ms, tl_msa_path = 1;
            } else {
                // DP-only rounds: a round's packing and graph update are single-lane latency (a DFS, a chain of list updates) that does
                // not shrink with the batch, and the finished graphs are converted on host threads. Large batches are therefore split
                // into pipeline lanes (host thread + stream + buffers each): one lane's latency-bound kernels and host conversion
                // overlap the other lanes' fill kernels.
                // (measured, 6,000 config-2 families: 0.87 s per call with one lane, 0.78 / 0.80 / 0.80 s with 2 / 3 / 4 - the fill
                // kernel owns the register file while it runs, so the gain is the overlapped host conversion, not the device time)
                int lanes = n_families >= 2048 ? 2 : 1;
                if (const char *e = getenv("POASTA_DP_LANES")) lanes = std::max(1, std::min(DP_LANES_MAX, atoi(e)));
                if (ebs[device][0].workers.d) {  // the reference-exact engine's worker slabs would starve the lanes' row stores
                    EngineBuffers &e0 = ebs[device][0];
                    e0.workers.release();
                    e0.workers_ptr = nullptr, e0.workers_sig = 0, e0.workers_zeroed = 0;
                }
                struct LaneOut {
                    int rc = 0, dp_kernel = 0;
                    EngineStats st;
                    uint64_t h2d = 0, d2h = 0;
                };
                std::vector<LaneOut> outs(lanes);
                const unsigned lane_threads = std::max(1u, lanes > 1 ? (host_threads * 2) / (unsigned)lanes : host_threads);
                auto run_lane = [&](int lane) {
                    tl_lane = lane, tl_lane_count = lanes;
                    tl_h2d = tl_d2h = 0, tl_dp_kernel = 0;
                    LaneOut &o = outs[lane];
                    DeviceContext &lc = ctx_for(device);
                    o.rc = lc.init(device);
                    if (o.rc == 0) {
                        const size_t f0 = (size_t)n_families * lane / lanes, f1 = (size_t)n_families * (lane + 1) / lanes;
                        std::vector<EngineFamilyIn> part(in.begin() + f0, in.begin() + f1);
                        o.rc = msa_engine_run(lc, ebs[device][lane], dbs[device][lane], part, *params, o.st, lane_threads);
                    }
                    o.h2d = tl_h2d, o.d2h = tl_d2h, o.dp_kernel = tl_dp_kernel;
                };
                if (lanes == 1) {
                    const uint64_t h0 = tl_h2d, d0 = tl_d2h;
                    run_lane(0);
                    outs[0].h2d += h0, outs[0].d2h += d0;
                } else {
                    std::vector<std::thread> pool;
                    for (int l = 0; l < lanes; ++l) pool.emplace_back(run_lane, l);
                    for (auto &t : pool) t.join();
                }
                tl_lane = 0, tl_lane_count = 1;
                tl_h2d = tl_d2h = 0, tl_dp_kernel = 0;
                double max_ms = 0;
                for (const LaneOut &o : outs) {
                    if (o.rc && !rc) rc = o.rc;
                    st.cells += o.st.cells, st.launches += o.st.launches, st.fill_ms += o.st.fill_ms, st.fill_launches += o.st.fill_launches;
                    st.kernel_ms += o.st.kernel_ms;
                    max_ms = std::max(max_ms, o.st.kernel_ms);
                    tl_h2d += o.h2d, tl_d2h += o.d2h;
                    tl_dp_kernel = std::max(tl_dp_kernel, o.dp_kernel);
                }
                // lanes overlap on the device: the critical path is the longest lane, kernel_ms the sum (as in the host pipeline)
                tl_kernel_ms = st.kernel_ms, tl_lane_max_ms = max_ms, tl_msa_path = 2, tl_fill_ms = st.fill_ms, tl_fill_launches = st.fill_launches;
            }
            tl_launches = st.launches, tl_cells = st.cells;
            tl_engine_counts[0] = st.queued, tl_engine_counts[1] = st.visited, tl_engine_counts[2] = st.pruned;
            if (cells_out) *cells_out = st.cells;
            return rc;
        }
    }
    tl_msa_path = 3;
    std::vector<size_t> score_base(n_families + 1, 0);
    for (uint32_t f = 0; f < n_families; ++f) {
        size_t n = 0;
        for (size_t k = 0; k < fasta_len[f]; ++k) n += fasta[f][k] == '>' && (k == 0 || fasta[f][k - 1] == '\n');
        score_base[f + 1] = score_base[f] + n;
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int lanes = 1;
    if (const char *e = getenv("POASTA_LANES")) lanes = std::max(1, std::min(MAX_LANES, atoi(e)));
    else lanes = (int)std::max<uint32_t>(1, std::min<uint32_t>(4, n_families / 256));
    // host phases of different lanes rarely coincide (the others wait on their kernels), so lanes share the cores generously
    unsigned threads_per_lane = std::max(2u, std::min(64u, hw) / lanes);
    if (const char *e = getenv("POASTA_THREADS_PER_LANE")) threads_per_lane = std::max(1, atoi(e));
    struct LaneOut {
        int rc = 0;
        uint64_t cells = 0, h2d = 0, d2h = 0;
        double kernel_ms = 0, search_ms = 0;
        uint32_t launches = 0;
    };
    std::vector<LaneOut> outs(lanes);
    auto run_lane = [&](int lane) {
        tl_lane = lane;
        tl_lane_count = lanes;
        tl_kernel_ms = 0, tl_search_ms = 0, tl_launches = 0, tl_cells = 0, tl_h2d = 0, tl_d2h = 0;
        const uint32_t f0 = (uint32_t)((uint64_t)n_families * lane / lanes), f1 = (uint32_t)((uint64_t)n_families * (lane + 1) / lanes);
        LaneOut &o = outs[lane];
        o.rc = msa_group(device, graphs + f0, fasta + f0, fasta_len + f0, f1 - f0, params, scores_out, score_base.data() + f0, &o.cells,
                         threads_per_lane);
        o.kernel_ms = tl_kernel_ms, o.search_ms = tl_search_ms, o.launches = tl_launches, o.h2d = tl_h2d, o.d2h = tl_d2h;
    };
    if (lanes == 1) {
        run_lane(0);
    } else {
        std::vector<std::thread> pool;
        for (int l = 0; l < lanes; ++l) pool.emplace_back(run_lane, l);
        for (auto &t : pool) t.join();
        tl_lane = 0, tl_lane_count = 1;
    }
    uint64_t cells = 0;
    int rc = 0;
    tl_kernel_ms = 0, tl_search_ms = 0, tl_launches = 0, tl_h2d = 0, tl_d2h = 0;
    tl_lane_max_ms = 0;
    for (const LaneOut &o : outs) {
        if (o.rc && !rc) rc = o.rc;
        cells += o.cells;
        tl_lane_max_ms = std::max(tl_lane_max_ms, o.kernel_ms + o.search_ms);
        tl_kernel_ms += o.kernel_ms, tl_search_ms += o.search_ms, tl_launches += o.launches, tl_h2d += o.h2d, tl_d2h += o.d2h;
    }
    tl_cells = cells;
    if (cells_out) *cells_out = cells;
    return rc;
}

int poasta_msa_batch(int device, const char *const *fasta, const size_t *fasta_len, uint32_t n_families,
                     const poasta_params *params, poasta_graph **graphs_out, uint32_t *scores_out, uint64_t *cells_out) {
    std::vector<poasta_graph *> graphs(n_families);
    for (uint32_t f = 0; f < n_families; ++f) graphs[f] = new poasta_graph();
    int rc = poasta_msa_batch_into(device, graphs.data(), fasta, fasta_len, n_families, params, scores_out, cells_out);
    for (uint32_t f = 0; f < n_families; ++f) {
        if (rc == 0 && graphs_out) graphs_out[f] = graphs[f];
        else delete graphs[f];
    }
    return rc;
}

}  // extern "C"
