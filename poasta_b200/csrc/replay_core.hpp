// Order-exact search: reproduces which co-optimal alignment the reference returns.
//
// POASTA's result is not "the DP optimum + a fixed traceback priority": its backtrace tests score
// equality against the sparse table of states its A* happened to record, and which co-optimal cells
// are recorded depends on bucket pop order (Match, then Deletion, then Insertion, LIFO inside each),
// greedy match extension and strict-improvement bubble pruning (SURVEY.md findings 2-3, A.4-A.9).
// The dense DP kernels (align.cu) give the exact optimal score; this routine gives the reference's
// exact alignment by running the same state machine with the same ordering rules:
//   search loop            /root/reference/src/aligner/astar.rs:124-225
//   greedy match extension /root/reference/src/aligner/dfa.rs:138-250
//   transitions            /root/reference/src/aligner/scoring/gap_affine.rs:307-430
//   bucket queue           /root/reference/src/aligner/queue.rs:31-70 + gap_affine.rs:945-966
//   prune test             /root/reference/src/bubbles/reached.rs:38-255 + gap_affine.rs:780-792
//   heuristics             /root/reference/src/aligner/heuristic.rs:40-184
//   backtrace              /root/reference/src/aligner/scoring/gap_affine.rs:550-657, 804-915
//
// Device mapping: the state machine is sequential per alignment, so one warp owns one alignment and
// lane 0 drives it (10^4 families in flight = 10^4 warps). A worker is ONE thread walking a chain of
// ~260 dependent instructions and a handful of dependent loads per popped state (profiles/README.md:
// ~9 cycles per issued instruction, a third of it memory), so both the code and the layout are kept short:
//   node record    32 B per node: successor / bubble list ranges, symbol, is-exit, distances to `end`,
//                  predecessor range - one sector instead of seven arrays
//   visited table  dense V x W records {M, D, I, pad} (u16 cells when every reachable score fits): a popped
//                  state and the states it relaxes share a sector; the top bit of a Match cell is the
//                  "reached" mark of the prune test (the reference's BTreeSet<offset> per bubble exit), so
//                  marking costs no extra access and a prune query reads flags and scores from the same
//                  cache lines; [lo, hi] bounds per node keep the scans inside the occupied band
//   bucket queue   three LIFO stacks per priority threaded through a 16-byte item array with a free
//                  list; the heads of the 64 priorities around the current one live in shared memory
//   code           the heuristic is a template parameter and every "relax, queue if improved" step goes
//                  through one site, which halved the kernel (instruction fetch was 20-30 % of its stalls)
// The same source compiles for the host (tests only) so the logic can be checked without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RP_HD __host__ __device__ __forceinline__
#else
#define RP_HD inline
#endif

// non-binding L1 prefetch hints. Measured on B200 (profiles/README.md): prefetch.global.L1 at the points below made the
// kernel 20 % SLOWER (3.70 s vs 3.09 s per 20,000 alignments), so the hints are compiled out unless RP_USE_PREFETCH is set.
#if defined(__CUDA_ARCH__) && defined(RP_USE_PREFETCH)
#define RP_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define RP_PREFETCH(p) ((void)0)
#endif

// optional cycle accounting of the device build (-DRP_PROFILE): where a worker's time goes
#if defined(RP_PROFILE) && defined(__CUDA_ARCH__)
#define RP_T0 const long long _t0 = clock64()
#define RP_ACC(x) x += clock64() - _t0
#else
#define RP_T0
#define RP_ACC(x)
#endif

// optional event counters of the HOST build (-DRP_COUNT, tools/search_stats.py): what the state machine does per alignment
#if defined(RP_COUNT) && !defined(__CUDA_ARCH__)
namespace poasta { namespace replay {
enum { C_POP_M, C_POP_D, C_POP_I, C_POP_STALE, C_DFA_STEP, C_DFA_MATCH, C_DFA_MISMATCH, C_DFA_RELAX_FAIL, C_PRUNE_CALL, C_PRUNE_BUB,
       C_PRUNE_NOREACH, C_PRUNE_PROBE, C_PRUNE_SCORE, C_PUSH, C_RELAX, C_RELAX_OK, C_WINDOW_MOVE, C_POP_SCAN,
       C_NEXT_SAME_STACK, C_NEXT_JUST_PUSHED, C_NEXT_OTHER,
       C_N };
inline uint64_t rp_counters[C_N];
} }
#define RP_CNT(x) (++::poasta::replay::rp_counters[::poasta::replay::x])
#define RP_STAT(...) __VA_ARGS__
#else
#define RP_CNT(x) ((void)0)
#define RP_STAT(...)
#endif

namespace poasta {
namespace replay {

constexpr uint32_t UNV = 0xFFFFFFFFu;  // Score::Unvisited
enum : uint32_t { ST_M = 0, ST_D = 1, ST_I = 2 };
enum : int { RP_OK = 0, RP_EMPTY_QUEUE = 1, RP_NO_BACKTRACE = 2, RP_SCORE_WRAP = 3, RP_OUT_OF_BLOCKS = 4, RP_OUT_OF_ITEMS = 5,
             RP_PRIORITY_RANGE = 6, RP_OUT_OF_PAIRS = 7 };
// Visited-table layout: one record {Match, Deletion, Insertion, epoch} per (node, offset), read and written as ONE 64-bit
// (16-bit cells) or 128-bit (32-bit cells) access. A popped state and the states it relaxes mostly share (node, offset) or
// sit at the neighbouring offset, so they share a 32-byte sector. The fourth slot is the EPOCH of the alignment that wrote
// the record: a record whose epoch is not the current one reads as "unvisited", so a worker never resets its V x W table
// between alignments (round 1 spent 66 % of the kernel's DRAM traffic on that reset; the reference's table is sparse for the
// same reason, gap_affine.rs:442-466). The worker's epoch counter lives in its slab header; the table is zeroed when the
// counter wraps or the slab is (re)allocated.
template <typename CellT> struct alignas(4 * sizeof(CellT)) Rec {
    CellT m, d, i, ep;
};
#ifndef RP_QWIN
#define RP_QWIN 64
#endif
constexpr uint32_t QWIN = RP_QWIN;  // priorities whose stack heads are kept in the fast window (power of two)

// BLOCKED visited table (reads against a large fixed graph, SURVEY 8d config 5: V x W = 25,000 x 2,000 records = 400 MB dense, of
// which one 2 kb read touches a sliver): records live in BLOCKS of 64 consecutive offsets of one node, handed out from a pool on
// first touch; a map entry per (node, offset / 64) holds (map epoch << 20 | block number). The reference's table is blocked for the
// same reason (gap_affine.rs:442-548). Entries of another map epoch read as "no block", so the map is never reset between
// alignments either; the records inside a recycled block carry their own (older) alignment epochs and read as unvisited.
constexpr uint32_t BLK_SHIFT = 6, BLK_CELLS = 1u << BLK_SHIFT;
constexpr uint32_t BMAP_EPOCH_MAX = 0xFFFu, BMAP_BLOCK_MASK = 0xFFFFFu;

struct alignas(16) NodeA {
    uint32_t succ_off, bub_off;  // into Graph::succ / Graph::bub (quadruples)
    uint32_t counts;             // n_succ | n_bub << 16
    uint32_t symflags;           // symbol | is_exit << 8
};
struct alignas(16) BubQ {  // one enclosing superbubble of a node
    uint32_t exit, min_d, max_d;  // exit node, shortest / longest path length from the node to it
    uint32_t mde;                 // max(dist_min(exit), 1) - 1: min_dist_to_end used by the prune test (reached.rs:57-58)
};
struct alignas(16) NodeB {
    uint32_t dist_min, dist_max;  // shortest / longest path length to `end`
    uint32_t pred_off, n_pred;    // into Graph::pred
};
struct alignas(32) NodeRec {  // one sector per node: everything a visit needs except the adjacency and bubble lists
    NodeA a;
    NodeB b;
};

struct Graph {  // node-id space, adjacency in the reference's order
    uint32_t V, start, end;
    const NodeRec *nodes;
    const uint32_t *succ, *pred;
    const BubQ *bub;                      // enclosing bubbles of every node, innermost first
    const uint32_t *path_off, *path_rem;  // sampled-path remainders for -H path (path_off has V + 1 entries)
};

// Queue items live in CHUNKS of 8 (one 128-byte line): a bucket stack is a chain of chunks, its head word = top slot + 1.
// `aux` of a chunk's slot 0 = the head word the stack had before the chunk was opened (0 = none). The top items of the
// stack being popped sit in ONE cache line (two thirds of all pops take the item directly below the previous one), and a
// chunk is recycled as a whole instead of slot by slot.
constexpr uint32_t CHUNK = 8;
struct alignas(16) Item {
    uint32_t score, node, offset, aux;
};

struct Scratch {  // one slab per in-flight alignment
    void *planes;          // V * W records Rec<CellT> {M, D, I, epoch}; cell = score, all-ones = unvisited
    uint32_t epoch;        // epoch of this alignment (1 .. CellT max); records stamped otherwise read as unvisited
    uint32_t *qhead;       // f_max * 3 stack heads (0 = empty, else item index + 1) - authoritative outside the window
    uint32_t *qwin;        // QWIN * 3 heads of priorities [win_base, win_base + QWIN): shared memory on the device
    Item *items;           // max_items (multiple of CHUNK), chunk c = slots [8 c, 8 c + 8)
    uint32_t *free_slots;  // max_items / CHUNK: stack of recycled chunks (its top stays cache-hot, unlike a list threaded through items)
    uint32_t *rb;          // per node [lo, hi + 1] of the reached offsets (hi + 1 == 0: none)
    uint32_t *rbits;       // per node Wb words: bit o = offset o is a reached offset of this (exit) node - the reference's
                           // BTreeSet<offset> per bubble exit (gap_affine.rs:711, 767-773); cleared per alignment
    uint32_t Wb;           // words per bitmap row = (W + 31) / 32
    uint32_t *dfa;         // (V + 2) frames of 6 words: node, offset, next successor slot, end slot, dist_min, dist_max
    uint32_t W, max_items, f_max;
    // blocked visited table only (Search<..., BLOCKED = true>): planes = the block pool
    uint32_t *bmap = nullptr;  // V * bpr entries
    uint32_t bpr = 0;          // map entries per node = ceil(W / 64)
    uint32_t max_blocks = 0;   // pool capacity (< 2^20)
    uint32_t bmap_epoch = 0;   // 1 .. BMAP_EPOCH_MAX
};

// Read-only lookup of the record of (node, off); nullptr = no block there (blocked table only).
template <typename CellT, bool BLOCKED>
RP_HD const Rec<CellT> *table_lookup(const void *planes, const uint32_t *bmap, uint32_t bpr, uint32_t bmap_epoch, uint32_t W, uint32_t node, uint32_t off) {
    if constexpr (!BLOCKED) {
        return static_cast<const Rec<CellT> *>(planes) + (size_t)node * W + off;
    } else {
        const uint32_t e = bmap[(size_t)node * bpr + (off >> BLK_SHIFT)];
        if ((e >> 20) != bmap_epoch) return nullptr;
        return static_cast<const Rec<CellT> *>(planes) + ((size_t)(e & BMAP_BLOCK_MASK) << BLK_SHIFT) + (off & (BLK_CELLS - 1));
    }
}

struct Params {
    uint32_t x, o, e;
    int mode;       // 0 global, 1 ends-free (all bounds Unbounded)
    int heuristic;  // 0 dijkstra, 1 mingap, 2 path
};

struct Result {
    uint32_t score;
    uint32_t n_pairs;  // pairs are written BACKWARDS from out_r[cap-1]; the alignment is out_r[cap-n .. cap)
    int status;
    uint32_t num_queued, num_visited, num_pruned;
};

template <typename CellT> struct CellTraits;
// CHECK_WRAP: the reference's NonMaxU32 arithmetic panics when a score reaches u32::MAX. The host picks 16-bit cells only
// when every reachable score stays below CELL16_LIMIT, so that instantiation needs none of those tests in its hot loop.
template <> struct CellTraits<uint16_t> {
    static constexpr uint16_t CUNV = 0xFFFFu;
    static constexpr bool CHECK_WRAP = false;
};
template <> struct CellTraits<uint32_t> {
    static constexpr uint32_t CUNV = 0xFFFFFFFFu;
    static constexpr bool CHECK_WRAP = true;
};
constexpr uint32_t CELL16_LIMIT = 0xFF00u;  // 16-bit cells are picked only when no reachable score can get here

RP_HD uint32_t rp_clz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return (uint32_t)__builtin_clz(x);
#endif
}
RP_HD uint32_t rp_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

// ---- prune test (bubbles/reached.rs:38-255 + gap_affine.rs:780-792) -----------------------------------------------
// Returns 0 = can still improve (keep), 1 = prune, 2 = score arithmetic wrapped (the reference would panic).
template <bool BLOCKED> struct PruneBlocks {};  // the block map of the blocked table; nothing for the dense one
template <> struct PruneBlocks<true> {
    const uint32_t *bmap;
    uint32_t bpr, bmap_epoch;
};
template <typename CellT, bool BLOCKED>
struct PruneCtx : PruneBlocks<BLOCKED> {
    const Rec<CellT> *table;  // the visited table; record of (node, off) at node * W + off (blocked table: the block pool)
    const uint32_t *rb;       // reached bounds per node
    const uint32_t *rbits;    // reached-offset bitmap, Wb words per node
    uint32_t W, Wb, Q, o, e;
};

// can_improve_bubble for one bubble; true = the state can still improve on it (no prune from this bubble)
template <typename CellT, bool BLOCKED>
RP_HD bool prune_bubble(const PruneCtx<CellT, BLOCKED> &c, const BubQ &b, uint32_t node, uint32_t off, uint32_t st, uint32_t score,
                        bool &wrapped) {
    constexpr CellT CUNV = CellTraits<CellT>::CUNV;
    RP_CNT(C_PRUNE_BUB);
    if (node == b.exit || off + b.max_d > c.Q) return true;
    const uint32_t lo = c.rb[2 * (size_t)b.exit], hi1 = c.rb[2 * (size_t)b.exit + 1], mde = b.mde;
    if (hi1 == 0) {  // no offset reached at this exit yet
        RP_CNT(C_PRUNE_NOREACH);
        return true;
    }
    auto gap = [&](uint32_t len) -> uint32_t { return len == 0 ? 0u : c.o + len * c.e; };  // gap_cost(Match, len)
    const uint32_t tmin = off + b.min_d, tmax = off + b.max_d;
    const Rec<CellT> *mrow = c.table + (size_t)b.exit * c.W;  // dense table only
    (void)mrow;
    const uint32_t *bits = c.rbits + (size_t)b.exit * c.Wb;
    // reached offsets = set bits of the exit's bitmap row, all inside [lo, hi1): a range query is one or two word operations
    // (round 1 scanned the Match cells one by one: ~10 dependent probes per visited state)
    auto prev_of = [&](uint32_t lim) -> uint32_t {  // largest reached offset < lim
        if (lim <= lo) return UNV;
        const uint32_t i = (lim > hi1 ? hi1 : lim) - 1, wlo = lo >> 5;
        uint32_t w = i >> 5;
        RP_CNT(C_PRUNE_PROBE);
        uint32_t word = bits[w] & (0xFFFFFFFFu >> (31 - (i & 31)));
        for (;;) {
            if (word) return w * 32 + 31 - rp_clz(word);
            if (w == wlo) return UNV;
            RP_CNT(C_PRUNE_PROBE);
            word = bits[--w];
        }
    };
    auto next_of = [&](uint32_t lim) -> uint32_t {  // smallest reached offset >= lim
        if (lim >= hi1) return UNV;
        const uint32_t i = lim < lo ? lo : lim, whi = (hi1 - 1) >> 5;
        uint32_t w = i >> 5;
        RP_CNT(C_PRUNE_PROBE);
        uint32_t word = bits[w] & (0xFFFFFFFFu << (i & 31));
        for (;;) {
            if (word) return w * 32 + rp_ctz(word);
            if (w == whi) return UNV;
            RP_CNT(C_PRUNE_PROBE);
            word = bits[++w];
        }
    };
    auto S = [&](uint32_t r, uint32_t plus) -> uint32_t {  // get_score(exit, r, Match) + plus, with the NonMax checks
        // a reached offset was written by THIS alignment: no epoch test (and its block exists)
        CellT v;
        if constexpr (!BLOCKED) {
            v = mrow[r].m;
        } else {
            const Rec<CellT> *rec = table_lookup<CellT, true>(c.table, c.bmap, c.bpr, c.bmap_epoch, c.W, b.exit, r);
            v = rec ? rec->m : CUNV;
        }
        RP_CNT(C_PRUNE_SCORE);
        const uint32_t t = (uint32_t)v + plus;
        if (CellTraits<CellT>::CHECK_WRAP && (v == CUNV || t == UNV)) wrapped = true;
        return t;
    };
    auto improve_at = [&](uint32_t x, uint32_t left, uint32_t right) -> bool {  // reached.rs:191-255
        uint32_t implicit;
        if (left != UNV && right != UNV) {
            const uint32_t il = S(left, gap(x - left));
            const uint32_t gr = right - x;
            const uint32_t ir = S(right, gap(gr));
            implicit = gr > mde ? il : (il < ir ? il : ir);
        } else if (right != UNV) {
            const uint32_t gr = right - x;
            const uint32_t ir = S(right, gap(gr));
            if (gr > mde) return true;
            implicit = ir;
        } else if (left != UNV) {
            implicit = S(left, gap(x - left));
        } else {
            return true;
        }
        return score < implicit;
    };
    uint32_t prev = prev_of(tmin);
    bool have_last = false;
    uint32_t last = 0;
    for (uint32_t nx = next_of(tmin); nx != UNV && nx <= tmax; nx = next_of(nx + 1)) {
        const uint32_t o1 = prev != UNV ? (tmin > prev + 1 ? tmin : prev + 1) : tmin;
        if (st == ST_D && S(nx, c.o) > score) return true;
        if (prev != UNV && st == ST_I && S(prev, c.o) > score) return true;
        if (improve_at(o1, prev, nx)) return true;
        uint32_t o2 = nx - 1u;  // wraps for nx == 0 like the release build
        if (o2 < tmin) o2 = tmin;
        if (o2 > tmax) o2 = tmax;
        if (o2 != o1 && improve_at(o2, prev, nx)) return true;
        prev = nx;
        last = o2;
        have_last = true;
    }
    const uint32_t next = next_of(tmax == UNV ? tmax : tmax + 1);
    if (!have_last && improve_at(tmin, prev, next)) return true;
    if ((!have_last || last < tmax) && improve_at(tmax, prev, next)) return true;
    if (prev != UNV && st == ST_I && S(prev, c.o) > score) return true;
    return false;
}

// All bubbles of a node, innermost first (index.rs order); one 128-bit load per quadruple.
template <typename CellT, bool BLOCKED>
RP_HD int prune_eval(const PruneCtx<CellT, BLOCKED> &c, const BubQ *bubs, uint32_t n_bub, uint32_t node, uint32_t off, uint32_t st, uint32_t score) {
    bool wrapped = false;
    for (uint32_t k = 0; k < n_bub; ++k) {
        const BubQ b = bubs[k];
        const bool can = prune_bubble<CellT, BLOCKED>(c, b, node, off, st, score, wrapped);
        if (wrapped) return 2;
        if (!can) return 1;  // !all(can_improve)
    }
    return 0;
}

// CellT = uint16_t halves the table footprint (more alignments in flight); the host picks it only when the largest
// possible score (every node deleted + every query base inserted) is below CELL16_LIMIT.
// HEUR = Params::heuristic as a compile-time constant: every push evaluates h(), and the three heuristics inlined at every
// push site tripled the kernel's code size (instruction-cache misses were 20-30 % of its stall cycles).
// BLOCKED = the visited table is a pool of 64-offset blocks behind a map (see BLK_SHIFT); false = dense V x W (the default).
template <typename CellT, int HEUR, bool BLOCKED = false>
struct Search {
    static constexpr CellT CUNV = CellTraits<CellT>::CUNV;
    static constexpr bool CHECK_WRAP = CellTraits<CellT>::CHECK_WRAP;
    const Graph &g;
    const uint8_t *q;
    const uint32_t Q;
    const Params &p;
    Scratch &s;
    uint32_t n_items = 0, n_free = 0;  // bump pointer into the item arena (slots, multiple of CHUNK); height of the recycled-chunk stack
    uint32_t f_cur = 0, f_hi = 0, win_base = 0;
    uint32_t f_top = 0;  // highest priority queued so far: the next alignment of this worker clears qhead[0 .. f_top]
    uint32_t n_blocks = 0;  // blocked table: blocks handed out so far
    bool q_empty = true;
    int err = RP_OK;
    // counting builds: how predictable is the next pop? (the item below the popped one / an item queued since / anything else)
    RP_STAT(uint32_t stat_below = 0; uint32_t stat_pushes_at_pop = 0; uint32_t stat_pushes = 0;)
#if defined(RP_PROFILE)
    long long prof_prune = 0, prof_push = 0, prof_pop = 0;
    uint32_t prof_n_prune = 0;
#endif

    RP_HD Search(const Graph &g_, const uint8_t *q_, uint32_t Q_, const Params &p_, Scratch &s_) : g(g_), q(q_), Q(Q_), p(p_), s(s_) {}

    RP_HD NodeA node_a(uint32_t n) const { return g.nodes[n].a; }
    RP_HD NodeB node_b(uint32_t n) const { return g.nodes[n].b; }
    RP_HD bool sym_eq(uint32_t node, const NodeA &a, uint8_t c) const { return node == g.end || (a.symflags & 0xFF) == c; }

    // ---- visited table -------------------------------------------------------------------------
    using R = Rec<CellT>;
    // the record of (node, off); the blocked table hands out the block on first touch (reads included: every read of the search
    // loop is of a record it wrote earlier or writes next)
    RP_HD R *rec_ptr(uint32_t node, uint32_t off) {
        if constexpr (!BLOCKED) {
            return static_cast<R *>(s.planes) + (size_t)node * s.W + off;
        } else {
            uint32_t *me = s.bmap + (size_t)node * s.bpr + (off >> BLK_SHIFT);
            uint32_t e = *me;
            if ((e >> 20) != s.bmap_epoch) {
                if (n_blocks >= s.max_blocks) {
                    err = RP_OUT_OF_BLOCKS;  // the caller reruns the alignment with a dense table; block 0 is valid memory until the loop stops
                    e = 0;
                } else {
                    e = n_blocks++;
                    *me = (s.bmap_epoch << 20) | e;
                }
            }
            return static_cast<R *>(s.planes) + ((size_t)(e & BMAP_BLOCK_MASK) << BLK_SHIFT) + (off & (BLK_CELLS - 1));
        }
    }
    // one vector load; a record of another epoch reads as three unvisited cells (stamped with the current epoch, so that
    // storing it back after changing one cell makes the whole record current)
    RP_HD R load(const R *ptr) const {
        R r = *ptr;
        if (r.ep != (CellT)s.epoch) r = R{CUNV, CUNV, CUNV, (CellT)s.epoch};
        return r;
    }
    RP_HD static CellT field(const R &r, uint32_t st) { return st == ST_M ? r.m : (st == ST_D ? r.d : r.i); }
    RP_HD static void set_field(R &r, uint32_t st, CellT v) {
        if (st == ST_M) r.m = v;
        else if (st == ST_D) r.d = v;
        else r.i = v;
    }
    RP_HD uint32_t get(uint32_t node, uint32_t off, uint32_t st) const {  // read-only (the backtrace): never hands out a block
        const R *ptr;
        if constexpr (!BLOCKED) {
            ptr = static_cast<const R *>(s.planes) + (size_t)node * s.W + off;
        } else {
            ptr = table_lookup<CellT, true>(s.planes, s.bmap, s.bpr, s.bmap_epoch, s.W, node, off);
            if (!ptr) return UNV;
        }
        const CellT v = field(load(ptr), st);
        return v == CUNV ? UNV : (uint32_t)v;
    }
    RP_HD void set(uint32_t node, uint32_t off, uint32_t st, uint32_t v) {
        R *ptr = rec_ptr(node, off);
        R r = load(ptr);
        set_field(r, st, (CellT)v);
        *ptr = r;
    }
    // update_score_if_lower (strict) on a record loaded earlier (together with other loads) and not written since
    RP_HD bool relax_loaded(R *ptr, R r, uint32_t st, uint32_t v) {
        if (CHECK_WRAP && v >= (uint32_t)CUNV) {  // cannot happen when the host sized CellT correctly
            err = RP_SCORE_WRAP;
            return false;
        }
        const CellT curv = field(r, st);
        RP_CNT(C_RELAX);
        if (v < (uint32_t)curv) {  // unvisited = all ones: greater than any score
            RP_CNT(C_RELAX_OK);
            set_field(r, st, (CellT)v);
            *ptr = r;
            return true;
        }
        return false;
    }
    RP_HD bool relax(uint32_t node, uint32_t off, uint32_t st, uint32_t v) {
        R *ptr = rec_ptr(node, off);
        return relax_loaded(ptr, load(ptr), st, v);
    }

    // ---- score arithmetic with the reference's NonMaxU32 / wrapping behaviour -------------------------
    RP_HD uint32_t add(uint32_t a, uint32_t b) {
        const uint32_t r = a + b;
        if (CHECK_WRAP && (a == UNV || r == UNV)) err = RP_SCORE_WRAP;
        return r;
    }
    RP_HD uint32_t sub(uint32_t a, uint32_t b) {
        const uint32_t r = a - b;
        if (r == UNV) err = RP_SCORE_WRAP;
        return r;
    }
    RP_HD uint32_t gap_cost(uint32_t st, uint32_t len) const {
        if (len == 0) return 0;
        return (st == ST_M ? p.o : 0) + len * p.e;
    }

    // ---- heuristic (dist = the node's distances to `end`) -----------------------------------------------
    RP_HD uint32_t h(uint32_t node, uint32_t dist_min, uint32_t dist_max, uint32_t off, uint32_t st) const {
        if constexpr (HEUR == 0) {
            return 0;
        } else if constexpr (HEUR == 1) {
            const uint32_t tmin = off + (dist_min ? dist_min - 1 : 0), tmax = off + (dist_max ? dist_max - 1 : 0);
            if (tmin > Q) return gap_cost(st == ST_D ? ST_D : ST_M, tmin - Q);
            if (tmax < Q) return gap_cost(st == ST_I ? ST_I : ST_M, Q - tmax);
            return 0;
        } else {
            const uint32_t rem_q = Q > off ? Q - off : 0;
            const uint32_t b = g.path_off[node], e = g.path_off[node + 1];
            if (b == e) return gap_cost(st, rem_q);
            uint32_t best = UNV;
            for (uint32_t k = b; k < e; ++k) {
                const uint32_t rem_p = g.path_rem[k];
                uint32_t c = 0;
                if (rem_p > rem_q) c = gap_cost(st == ST_D ? ST_D : ST_M, rem_p - rem_q);
                else if (rem_q > rem_p) c = gap_cost(st == ST_I ? ST_I : ST_M, rem_q - rem_p);
                if (c < best) best = c;
            }
            return best;
        }
    }

    // ---- bucket queue ---------------------------------------------------------------------------------
    RP_HD uint32_t *head_ptr(uint32_t f) const {
        const uint32_t d = f - win_base;
        return d < QWIN ? s.qwin + d * 3 : s.qhead + (size_t)f * 3;
    }
    RP_HD void move_window(uint32_t f) {  // make [f & ~63, +64) the fast window
        const uint32_t nb = f & ~(QWIN - 1);
        if (nb == win_base) return;
        RP_CNT(C_WINDOW_MOVE);
        for (uint32_t i = 0; i < QWIN * 3; ++i) {
            const uint32_t fo = win_base * 3 + i, fn = nb * 3 + i;
            if (fo < s.f_max * 3) s.qhead[fo] = s.qwin[i];
            s.qwin[i] = fn < s.f_max * 3 ? s.qhead[fn] : 0;
        }
        win_base = nb;
    }
    RP_HD void push(uint32_t node, uint32_t dist_min, uint32_t dist_max, uint32_t off, uint32_t st, uint32_t score) {
        RP_T0;
        RP_CNT(C_PUSH);
        RP_STAT(++stat_pushes;)
        const uint32_t f = score + h(node, dist_min, dist_max, off, st);
        if (f >= s.f_max) {
            err = RP_PRIORITY_RANGE;
            return;
        }
        if (f > f_top) f_top = f;
        uint32_t *head = head_ptr(f) + st;
        const uint32_t hv = *head;
        uint32_t slot, aux = 0;
        if (hv == 0 || (hv & (CHUNK - 1)) == 0) {  // empty stack, or its top chunk is full: open a chunk
            if (n_free) {
                slot = s.free_slots[--n_free] * CHUNK;
            } else {
                if (n_items + CHUNK > s.max_items) {
                    err = RP_OUT_OF_ITEMS;
                    return;
                }
                slot = n_items;
                n_items += CHUNK;
            }
            aux = hv;
        } else {
            slot = hv;  // the slot above the current top (top = hv - 1)
        }
        s.items[slot] = Item{score, node, off, aux};
        *head = slot + 1;
        if (q_empty) {
            f_cur = f_hi = f;
            q_empty = false;
        } else {
            if (f < f_cur) f_cur = f;
            if (f > f_hi) f_hi = f;
        }
        RP_ACC(prof_push);
    }
    // push() with the priority computed earlier (prepare_initial)
    RP_HD void push_f(uint32_t node, uint32_t off, uint32_t st, uint32_t score, uint32_t f) {
        RP_T0;
        RP_CNT(C_PUSH);
        RP_STAT(++stat_pushes;)
        if (f >= s.f_max) {
            err = RP_PRIORITY_RANGE;
            return;
        }
        if (f > f_top) f_top = f;
        uint32_t *head = head_ptr(f) + st;
        const uint32_t hv = *head;
        uint32_t slot, aux = 0;
        if (hv == 0 || (hv & (CHUNK - 1)) == 0) {  // empty stack, or its top chunk is full: open a chunk
            if (n_free) {
                slot = s.free_slots[--n_free] * CHUNK;
            } else {
                if (n_items + CHUNK > s.max_items) {
                    err = RP_OUT_OF_ITEMS;
                    return;
                }
                slot = n_items;
                n_items += CHUNK;
            }
            aux = hv;
        } else {
            slot = hv;  // the slot above the current top (top = hv - 1)
        }
        s.items[slot] = Item{score, node, off, aux};
        *head = slot + 1;
        if (q_empty) {
            f_cur = f_hi = f;
            q_empty = false;
        } else {
            if (f < f_cur) f_cur = f;
            if (f > f_hi) f_hi = f;
        }
        RP_ACC(prof_push);
    }
    RP_HD bool pop(Item &out, uint32_t &st) {
        if (q_empty) return false;
        for (;;) {
            if (f_cur - win_base >= QWIN) move_window(f_cur);
            uint32_t *head = head_ptr(f_cur);
            RP_CNT(C_POP_SCAN);
            for (uint32_t k = 0; k < 3; ++k) {  // Match, then Deletion, then Insertion
                if (head[k]) {
                    const uint32_t slot = head[k] - 1;
                    out = s.items[slot];
                    uint32_t below;
                    if ((slot & (CHUNK - 1)) == 0) {  // chunk exhausted: back to the chunk below, recycle this one
                        below = out.aux;
                        s.free_slots[n_free++] = slot / CHUNK;
                    } else {
                        below = slot;
                    }
                    RP_STAT(if (slot + 1 == stat_below) RP_CNT(C_NEXT_SAME_STACK);
                            else if (stat_pushes != stat_pushes_at_pop) RP_CNT(C_NEXT_JUST_PUSHED);  // approximation: something was queued since
                            else RP_CNT(C_NEXT_OTHER);
                            stat_below = below; stat_pushes_at_pop = stat_pushes;)
                    head[k] = below;
                    st = k;
                    return true;
                }
            }
            if (f_cur >= f_hi) {
                q_empty = true;
                return false;
            }
            ++f_cur;
        }
    }

    // ---- reached marks (BTreeSet<offset> per exit; gap_affine.rs:767-773) -------------------------------
    RP_HD void mark_reached(uint32_t node, bool is_exit, uint32_t off, uint32_t st) {
        if (st == ST_M && is_exit) {
            s.rbits[(size_t)node * s.Wb + (off >> 5)] |= 1u << (off & 31);
            uint32_t *b = s.rb + 2 * (size_t)node;
            if (b[1] == 0) {
                b[0] = off, b[1] = off + 1;
            } else {
                if (off < b[0]) b[0] = off;
                if (off + 1 > b[1]) b[1] = off + 1;
            }
        }
    }

    RP_HD bool prune(uint32_t score, uint32_t node, const NodeA &a, uint32_t off, uint32_t st) {  // gap_affine.rs:780-792
        const uint32_t nb = a.counts >> 16;
        if (nb == 0) return false;
        RP_T0;
        RP_CNT(C_PRUNE_CALL);
        PruneCtx<CellT, BLOCKED> c;
        c.table = static_cast<const R *>(s.planes), c.rb = s.rb, c.rbits = s.rbits, c.W = s.W, c.Wb = s.Wb, c.Q = Q, c.o = p.o, c.e = p.e;
        if constexpr (BLOCKED) c.bmap = s.bmap, c.bpr = s.bpr, c.bmap_epoch = s.bmap_epoch;
        const int r = prune_eval<CellT, BLOCKED>(c, g.bub + a.bub_off, nb, node, off, st, score);
        if (r == 2) err = RP_SCORE_WRAP;
        RP_ACC(prof_prune);
#if defined(RP_PROFILE)
        ++prof_n_prune;
#endif
        return r == 1;
    }

    RP_HD bool is_end(uint32_t node, uint32_t off, uint32_t st) const {  // gap_affine.rs:185-248, Unbounded bounds
        if (p.mode == 0) return st == ST_M && node == g.end && off == Q;
        return st == ST_M && (off > 0 || Q == 0);
    }

    // ---- backtrace (gap_affine.rs:550-657) ------------------------------------------------------------------
    RP_HD bool bt(uint32_t node, uint32_t off, uint32_t st, uint32_t &bn, uint32_t &bo, uint32_t &bs) {
        const uint32_t cur = get(node, off, st);
        if (cur == UNV) return false;
        const NodeB nb = node_b(node);
        if (st == ST_M) {
            if (off > 0) {
                const bool me = sym_eq(node, node_a(node), q[off - 1]) || node == g.end;
                const uint32_t po = node == g.end ? off : off - 1;
                const uint32_t want = me ? cur : sub(cur, p.x);
                for (uint32_t k = nb.pred_off; k < nb.pred_off + nb.n_pred; ++k)
                    if (get(g.pred[k], po, ST_M) == want) {
                        bn = g.pred[k], bo = po, bs = ST_M;
                        return true;
                    }
            }
            if (get(node, off, ST_D) == cur) {
                bn = node, bo = off, bs = ST_D;
                return true;
            }
            if (get(node, off, ST_I) == cur) {
                bn = node, bo = off, bs = ST_I;
                return true;
            }
        } else if (st == ST_D) {
            const uint32_t w1 = sub(sub(cur, p.o), p.e), w2 = sub(cur, p.e);
            for (uint32_t k = nb.pred_off; k < nb.pred_off + nb.n_pred; ++k)
                if (get(g.pred[k], off, ST_M) == w1) {
                    bn = g.pred[k], bo = off, bs = ST_M;
                    return true;
                }
            for (uint32_t k = nb.pred_off; k < nb.pred_off + nb.n_pred; ++k)
                if (get(g.pred[k], off, ST_D) == w2) {
                    bn = g.pred[k], bo = off, bs = ST_D;
                    return true;
                }
        } else if (off > 0) {
            if (get(node, off - 1, ST_M) == sub(sub(cur, p.o), p.e) || get(node, off - 1, ST_I) == sub(cur, p.e)) {
                bn = node, bo = off - 1, bs = ST_M;  // both cases resume in Match state (gap_affine.rs:639-650)
                return true;
            }
        }
        return false;
    }

    // Ends-free alignments start one state at offset 0 of EVERY node. The per-node part of that - the heuristic value, the visited
    // record, (blocked table) the node's first block - is independent work, so all lanes of the worker share it here; lane 0 then
    // queues the states in the reference's order (run<true>). Priorities are parked in the DFA frame array, which is unused until
    // the first extension. Returns false when the shortcut does not apply (the caller then uses run<false>). Warp-uniform result.
    RP_HD bool prepare_initial(uint32_t lane, uint32_t n_lanes) {
        if (p.mode == 0) return false;
        if (BLOCKED && s.max_blocks < g.V) return false;
        uint32_t *ftmp = s.dfa;
        for (uint32_t n = lane; n < g.V; n += n_lanes) {
            if (n == g.start || n == g.end) {
                ftmp[n] = UNV;
                continue;
            }
            const NodeB b = node_b(n);
            ftmp[n] = h(n, b.dist_min, b.dist_max, 0, ST_M);  // score 0
            R *ptr;
            if constexpr (BLOCKED) {  // block n belongs to offsets [0, 64) of node n
                s.bmap[(size_t)n * s.bpr] = (s.bmap_epoch << 20) | n;
                ptr = static_cast<R *>(s.planes) + ((size_t)n << BLK_SHIFT);
            } else {
                ptr = static_cast<R *>(s.planes) + (size_t)n * s.W;
            }
            R r = load(ptr);  // set(n, 0, ST_M, 0)
            r.m = 0;
            *ptr = r;
        }
        if (BLOCKED) n_blocks = g.V;
        return true;
    }

    // The whole alignment (astar.rs:108-226 + mod.rs:114-145 non-empty graph). PREPARED: prepare_initial() has run.
    template <bool PREPARED = false>
    RP_HD Result run(uint32_t *out_r, uint32_t *out_q, uint32_t cap) {
        Result res{0, 0, RP_OK, 0, 0, 0};
        const uint32_t oe8 = (p.o + p.e) & 0xFF;  // the queued delta is a u8 sum in the reference (gap_affine.rs:288)
        auto push_initial = [&](uint32_t n) {
            const NodeB b = node_b(n);
            push(n, b.dist_min, b.dist_max, 0, ST_M, 0);
            set(n, 0, ST_M, 0);
            ++res.num_queued;
        };
        if (p.mode == 0) {
            push_initial(g.start);
        } else {
            bool any = false;
            for (uint32_t n = g.V; n-- > 0;)
                if (n != g.start && n != g.end) {
                    if constexpr (PREPARED) {
                        push_f(n, 0, ST_M, 0, s.dfa[n]);
                        ++res.num_queued;
                    } else {
                        push_initial(n);
                    }
                    any = true;
                }
            if (!any) push_initial(g.start);
        }
        uint32_t end_score = 0, end_node = 0, end_off = 0;
        bool done = false;
        while (!done && err == RP_OK) {
            Item it;
            uint32_t st;
            if (!pop(it, st)) {
                err = RP_EMPTY_QUEUE;
                break;
            }
            // the state's own cell (stale test) and its node record, issued as one group of loads
            const CellT own = field(load(rec_ptr(it.node, it.offset)), st);
            const NodeA na = node_a(it.node);
            const NodeB nb = node_b(it.node);
            const uint32_t n_succ = na.counts & 0xFFFF;
            {
                if (it.score > (uint32_t)own) {  // unvisited (all ones) never compares below a score
                    RP_CNT(C_POP_STALE);
                    continue;
                }
            }
            if (is_end(it.node, it.offset, st)) {
                ++res.num_visited;
                end_score = it.score, end_node = it.node, end_off = it.offset;
                break;
            }
            if (prune(it.score, it.node, na, it.offset, st)) {
                ++res.num_pruned;
                continue;
            }
            mark_reached(it.node, (na.symflags >> 8) & 1, it.offset, st);
            ++res.num_visited;
            const uint32_t score = it.score;
            if (st == ST_M) RP_CNT(C_POP_M);
            else if (st == ST_D) RP_CNT(C_POP_D);
            else RP_CNT(C_POP_I);
            // Every expansion ends in "relax, and queue if improved" steps in a fixed order. They are collected as up to three
            // transitions (plus the per-successor deletions of a popped D state) and applied at ONE place below, so the
            // queue / heuristic code exists once in the kernel instead of at a dozen inlined sites.
            //   Match: depth-first greedy match extension (dfa.rs:138-250), transitions per run end (gap_affine.rs:346-430).
            //          The top frame lives in registers; s.dfa holds the frames below it (most extensions stop after one
            //          or two nodes and never touch it).
            //   Insertion / Deletion: one pass through the same loop (gap_affine.rs:307-341).
            uint32_t depth = st == ST_M ? 1u : 0u;
            bool single_pass = st != ST_M;
            uint32_t *F = s.dfa;
            uint32_t t_node = it.node, t_off = it.offset, t_k = na.succ_off, t_end = na.succ_off + n_succ;
            uint32_t t_dmin = nb.dist_min, t_dmax = nb.dist_max;
            uint32_t dfa_visited = 0;
            const uint32_t open_score = add(add(score, p.o), p.e), open_queued = add(score, oe8);
            while ((depth > 0 || single_pass) && err == RP_OK) {
                // transitions of this pass: (node, offset, state, relaxed score); `gap-open` ones are queued with open_queued
                uint32_t nt = 0, tn0 = 0, to0 = 0, ts0 = 0, tv0 = 0, tn1 = 0, to1 = 0, ts1 = 0, tv1 = 0, tn2 = 0, to2 = 0, ts2 = 0, tv2 = 0;
                uint32_t open_mask = 0;        // bit i: transition i opens a gap
                uint32_t dk = 0, dk_end = 0;   // successor range whose Deletion states follow the listed transitions
                if (single_pass) {
                    single_pass = false;
                    tn0 = it.node, to0 = it.offset, ts0 = ST_M, tv0 = score, nt = 1;  // close the gap
                    if (st == ST_I) {
                        if (it.offset < Q) tn1 = it.node, to1 = it.offset + 1, ts1 = ST_I, tv1 = add(score, p.e), nt = 2;
                    } else {
                        dk = na.succ_off, dk_end = na.succ_off + n_succ;
                    }
                } else {
                    // offset-0 special case, evaluated on every resume while a single frame is on the stack (dfa.rs:145-168)
                    bool virtual_end = false;
                    uint32_t v_parent = 0;
                    if (depth == 1 && Q != 0 && t_off == 0 && sym_eq(t_node, na, q[0])) {
                        if (relax(t_node, 1, ST_M, score)) {
                            v_parent = t_node;
                            t_off = 1, t_k = na.succ_off;
                            mark_reached(t_node, (na.symflags >> 8) & 1, 1, ST_M);
                            ++dfa_visited;
                            if (Q == 1) virtual_end = true;
                        }
                    }
                    int kind = 0;  // 0 exhausted, 1 ref-graph end, 2 query end, 3 match, 4 mismatch
                    uint32_t pn = t_node, po = t_off, cn = 0, co = 0;
                    NodeA ca{0, 0, 0, 0};
                    NodeB cb{0, 0, 0, 0};
                    if (virtual_end) {
                        kind = 1, pn = v_parent, po = 0, cn = v_parent, co = 1;
                    } else {
                        while (t_k < t_end) {
                            const uint32_t child = g.succ[t_k++];
                            if (child == g.end) {
                                relax(child, po, ST_M, score);
                                kind = 1, cn = child, co = po;
                                break;
                            }
                            if (po >= Q) {
                                kind = 2, cn = child;
                                break;
                            }
                            // the child's cell and its record in one group of loads
                            R *cc = rec_ptr(child, po + 1);
                            const R ccur = load(cc);
                            ca = node_a(child);
                            cb = node_b(child);
                            RP_CNT(C_DFA_STEP);
                            if ((ca.symflags & 0xFF) == q[po]) {
                                if (relax_loaded(cc, ccur, ST_M, score)) {
                                    RP_CNT(C_DFA_MATCH);
                                    kind = 3, cn = child, co = po + 1;
                                    break;
                                }
                                RP_CNT(C_DFA_RELAX_FAIL);
                            } else {
                                RP_CNT(C_DFA_MISMATCH);
                                kind = 4, cn = child, co = po + 1;
                                break;
                            }
                        }
                    }
                    if (kind == 0) {
                        if (--depth > 0) {
                            const uint32_t *fr = F + (size_t)(depth - 1) * 6;
                            t_node = fr[0], t_off = fr[1], t_k = fr[2], t_end = fr[3], t_dmin = fr[4], t_dmax = fr[5];
                        }
                        continue;
                    }
                    if (kind == 3) {
                        if (prune(score, cn, ca, co, ST_M)) continue;
                        mark_reached(cn, (ca.symflags >> 8) & 1, co, ST_M);
                        ++dfa_visited;
                        uint32_t *fr = F + (size_t)(depth - 1) * 6;  // spill the frame that stops being the top
                        fr[0] = t_node, fr[1] = t_off, fr[2] = t_k, fr[3] = t_end, fr[4] = t_dmin, fr[5] = t_dmax;
                        t_node = cn, t_off = co, t_k = ca.succ_off, t_end = ca.succ_off + (ca.counts & 0xFFFF);
                        t_dmin = cb.dist_min, t_dmax = cb.dist_max;
                        ++depth;
                        continue;
                    }
                    if (kind == 1) {
                        if (is_end(cn, co, ST_M)) {
                            end_score = score, end_node = cn, end_off = co;
                            done = true;
                            break;
                        }
                        tn0 = pn, to0 = po + 1, ts0 = ST_I, tv0 = open_score, open_mask = 1, nt = 1;  // gap_affine.rs:346-367
                    } else if (kind == 2) {
                        tn0 = cn, to0 = po, ts0 = ST_D, tv0 = open_score, open_mask = 1, nt = 1;  // gap_affine.rs:369-391
                    } else {  // mismatch: gap_affine.rs:393-430
                        tn0 = cn, to0 = co, ts0 = ST_M, tv0 = add(score, p.x);
                        tn1 = pn, to1 = po + 1, ts1 = ST_I, tv1 = open_score;
                        tn2 = cn, to2 = po, ts2 = ST_D, tv2 = open_score;
                        open_mask = 6, nt = 3;
                    }
                }
                // ---- the one place where states are relaxed and queued ----
                for (uint32_t ti = 0;; ++ti) {
                    uint32_t n_, o_, s_, v_, qv_;
                    if (ti < nt) {
                        n_ = ti == 0 ? tn0 : (ti == 1 ? tn1 : tn2);
                        o_ = ti == 0 ? to0 : (ti == 1 ? to1 : to2);
                        s_ = ti == 0 ? ts0 : (ti == 1 ? ts1 : ts2);
                        v_ = ti == 0 ? tv0 : (ti == 1 ? tv1 : tv2);
                        qv_ = ((open_mask >> ti) & 1) ? open_queued : v_;
                    } else if (dk < dk_end) {
                        n_ = g.succ[dk++], o_ = it.offset, s_ = ST_D, v_ = qv_ = add(score, p.e);
                    } else {
                        break;
                    }
                    R *cc = rec_ptr(n_, o_);
                    const R ccur = load(cc);
                    const NodeB b = node_b(n_);  // loaded with the cell, used only when the relaxation succeeds
                    if (relax_loaded(cc, ccur, s_, v_)) {
                        push(n_, b.dist_min, b.dist_max, o_, s_, qv_);
                        ++res.num_queued;
                    }
                }
            }
            if (st == ST_M && !done) res.num_visited += dfa_visited;
        }
        res.score = end_score;
        if (err != RP_OK) {
            res.status = err;
            return res;
        }
        // backtrace (gap_affine.rs:804-915), pairs written backwards from out[cap-1]
        uint32_t pos = cap;
        if (Q == 0) {
        } else if (Q == 1 && end_off == 1 && sym_eq(end_node, node_a(end_node), q[0])) {
            out_r[--pos] = end_node;
            out_q[pos] = 0;
        } else {
            uint32_t cn, co, cs;
            bool ok = bt(end_node, end_off, ST_M, cn, co, cs) || bt(end_node, end_off, ST_I, cn, co, cs) ||
                      bt(end_node, end_off, ST_D, cn, co, cs);
            if (!ok) {
                if (Q <= 3) {
                    for (uint32_t i = Q; i-- > 0;) {
                        out_r[--pos] = end_node;
                        out_q[pos] = i;
                    }
                } else {
                    err = RP_NO_BACKTRACE;
                }
            } else {
                uint32_t bn, bo, bs;
                while (err == RP_OK && bt(cn, co, cs, bn, bo, bs)) {
                    if (cs == ST_M && bs != ST_M) {
                        cn = bn, co = bo, cs = bs;
                        continue;
                    }
                    if (pos == 0) {
                        err = RP_OUT_OF_PAIRS;
                        break;
                    }
                    --pos;
                    if (cs == ST_M) out_r[pos] = cn, out_q[pos] = co - 1;
                    else if (cs == ST_I) out_r[pos] = UNV, out_q[pos] = co - 1;
                    else out_r[pos] = cn, out_q[pos] = UNV;
                    if (bn == g.start) break;
                    cn = bn, co = bo, cs = bs;
                }
            }
        }
        res.n_pairs = cap - pos;
        res.status = err;
        return res;
    }
};

// scratch sizing shared by host and device code
struct ScratchDims {
    uint32_t W, f_max;
    RP_HD ScratchDims(uint32_t V, uint32_t Q, uint32_t x, uint32_t o, uint32_t e) {
        W = (Q + 2 + 7) & ~7u;  // offsets 0..Q with slack; multiple of 8 cells keeps rows 16-byte aligned
        uint32_t m = x > o + e ? x : o + e;
        if (m == 0) m = 1;
        f_max = (V + Q + 4) * m + o + 64;  // any score and any gap-cost heuristic value stays below this
        f_max = (f_max + QWIN - 1) & ~(QWIN - 1);
    }
};

}  // namespace replay
}  // namespace poasta
