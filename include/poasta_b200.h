/* poasta_b200 — C ABI of the B200-native gap-affine sequence-to-graph aligner.
 *
 * This is the drop-in boundary for POASTA's alignment path. Every entry point
 * names the reference interface it replaces (paths relative to the pangenome/poasta tree).
 * A Rust `-sys` crate binds these symbols 1:1; no torch / C++ types cross
 * this line. All functions return 0 on success or a negative `poasta_status`.
 *
 * There is NO CPU fallback behind `poasta_align_batch` / `poasta_msa_batch`: when no CUDA device
 * (or the sm_100a image) is available they fail with POASTA_ERR_CUDA.
 */
#ifndef POASTA_B200_H
#define POASTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POASTA_NONE 0xFFFFFFFFu /* Option::None for rpos / qpos (aligner/alignment.rs:3-14) */

typedef enum poasta_status {
    POASTA_OK = 0,
    POASTA_ERR_CUDA = -1,              /* no device / kernel launch failure                       */
    POASTA_ERR_INVALID_ARG = -2,
    POASTA_ERR_WEIGHTS_UNEQUAL = -3,   /* PoastaError::WeightsUnequalSize  (errors.rs:10-40)      */
    POASTA_ERR_INVALID_ALIGNMENT = -4, /* PoastaError::InvalidAlignment    (graphs/poa.rs:223)    */
    POASTA_ERR_GRAPH_CYCLE = -5,       /* PoastaError::GraphError from toposort (poa.rs:360)      */
    POASTA_ERR_NO_ALIGNMENT = -6,      /* panic!("Could not align sequence!") (aligner/astar.rs:143) */
    POASTA_ERR_UNSUPPORTED = -7,       /* mode/size outside what the device path implements       */
    POASTA_ERR_IO = -8,
    POASTA_ERR_INTERNAL = -9           /* the DP score and the search score of one problem disagree */
} poasta_status;

/* GapAffine::new(mismatch, gap_extend, gap_open)  (aligner/scoring/gap_affine.rs:19-30) plus
 * AlignmentType (aligner/scoring/mod.rs:50-62) and the -H heuristic (aligner/heuristic.rs:16-33). */
typedef struct poasta_params {
    uint8_t mismatch;   /* -n, default 4 */
    uint8_t gap_open;   /* -g, default 6 */
    uint8_t gap_extend; /* -e, default 2 */
    uint8_t mode;       /* 0 = Global, 1 = EndsFree{all Unbounded} (bin/poasta.rs:287-301) */
    uint8_t heuristic;  /* 0 = dijkstra, 1 = mingap (default), 2 = path (-H; changes which co-optimal path wins) */
    uint8_t tiebreak;   /* 0 = reference-exact: DP score + order-exact search for the alignment (default);
                           1 = DP only: optimal score, co-optimal alignment chosen by the backtrace priorities
                               on the dense tables (fastest; may differ from the reference at ties) */
} poasta_params;

/* The `AlignableRefGraph` trait (graphs/mod.rs:23-53) flattened: what the aligner reads of a graph.
 * Adjacency lists keep the reference's iteration order (newest edge first), which decides ties. */
typedef struct poasta_graph_view {
    uint32_t n_nodes;           /* node_count_with_start_and_end()                        */
    uint32_t start_node;        /* start_node()                                           */
    uint32_t end_node;          /* end_node()                                             */
    const uint32_t *topo_order; /* n_nodes node ids, start first, end last                */
    const uint32_t *pred_off;   /* n_nodes + 1                                            */
    const uint32_t *pred;       /* predecessors(), Incoming order                         */
    const uint32_t *succ_off;   /* n_nodes + 1                                            */
    const uint32_t *succ;       /* successors(), Outgoing order                           */
    const uint8_t *symbol;      /* n_nodes; '#' for start, '$' for end (poa.rs:102-103)   */
} poasta_graph_view;

/* AstarResult (aligner/astar.rs:81-90). rpos/qpos are caller-provided arrays with room for
 * `capacity` pairs (n_nodes + qlen always suffices); n_pairs is written by the call. */
typedef struct poasta_result {
    uint32_t score;
    uint32_t n_pairs;
    uint32_t *rpos;
    uint32_t *qpos;
    uint32_t capacity;
    int32_t status;     /* poasta_status of this alignment                                */
    uint64_t num_cells; /* (real nodes) x (query length): the cell count GCUPS is quoted on */
    /* AstarResult.num_queued / num_visited / num_pruned (aligner/astar.rs:81-90) as the order-exact search counted
     * them (tiebreak = 0; tests/test_heuristics.rs:53-56,142 read them). 0 with tiebreak = 1: the DP fill has no queue. */
    uint64_t num_queued, num_visited, num_pruned;
} poasta_result;

/* ---- the hot path ------------------------------------------------------------------------- */

/* Replaces PoastaAligner::align::<u32,_>(&graph, seq) (aligner/mod.rs:114-145) for a batch of
 * independent (graph, query) problems on one device: alignment i aligns queries[i] against
 * graphs[graph_of_query[i]]. Blocking. */
int poasta_align_batch(int device, const poasta_graph_view *graphs, uint32_t n_graphs,
                       const uint8_t *const *queries, const uint32_t *qlens,
                       const uint32_t *graph_of_query, uint32_t n_queries,
                       const poasta_params *params, poasta_result *results);
/* The same for graphs held as handles (poasta_graph, below): what the search and the DP kernels derive from a graph - CSR view,
 * superbubble / distance / sampled-path index (bubbles/index.rs:51-156, aligner/path_index.rs), node records, rank-ordered layout -
 * is kept in the handle until the graph changes, so batch after batch against the same graphs (reads against a fixed graph,
 * bin/poasta.rs `align -g`) does not rebuild it. Replaces PoastaAligner::align(&graph, seq), aligner/mod.rs:114-145. */
struct poasta_graph;
int poasta_align_batch_graphs(int device, struct poasta_graph *const *graphs, uint32_t n_graphs,
                              const uint8_t *const *queries, const uint32_t *qlens,
                              const uint32_t *graph_of_query, uint32_t n_queries,
                              const poasta_params *params, poasta_result *results);
/* Threading: any host thread may call; calls that name the same `device` are serialised inside the library for their
 * whole duration (they share that device's streams and buffers), calls on different devices run concurrently.
 * 0 <= device < 16. */

/* Timing of the device work of the last poasta_align_batch on this thread (CUDA events on the
 * launching stream): milliseconds of the fill+traceback kernels, number of launches, cells. */
int poasta_last_batch_stats(double *kernel_ms, uint32_t *n_launches, uint64_t *cells);
/* Milliseconds the order-exact search kernel took in the last call on this thread (0 with tiebreak = 1). */
int poasta_last_batch_search_ms(double *search_ms);
/* Device time of the last call on this thread: kernels only, CUDA events on the launching streams. For poasta_msa_batch*
 * the pipeline lanes run their kernels concurrently, so this is the LARGEST per-lane sum (the device-side critical path
 * of the call), whereas poasta_last_batch_stats / _search_ms report the sums over all lanes. */
int poasta_last_batch_device_ms(double *device_ms);
/* num_queued / num_visited / num_pruned summed over every alignment of the last poasta_msa_batch* on this thread
 * (reference-exact mode). */
int poasta_last_batch_search_counts(uint64_t *num_queued, uint64_t *num_visited, uint64_t *num_pruned);
/* Which DP fill kernel the last chunk of the last DP-path call on this thread ran (a diagnostic for tests and benches; no
 * reference counterpart): 0 none yet, 1 CTA-per-alignment (32-bit), 2 warp-per-alignment (32-bit, bit planes),
 * 3 packed s16x2 warp-per-alignment, 4 packed s16x2 thread-block cluster (queries of 1024+ bases). */
int poasta_last_dp_kernel(void);
/* Which pipeline the last poasta_msa_batch* on this thread took (diagnostic): 0 none yet, 1 the reference-exact engine (one
 * persistent kernel, everything on the device), 2 the device-resident DP-only rounds (tiebreak = 1: pack, fill, traceback and graph
 * update on the device), 3 the per-round host pipeline of the DP-only mode (inputs the packed kernels do not cover, ends-free). */
int poasta_last_msa_path(void);
/* Visited-table layout of the last search launch of poasta_align_batch on this thread (diagnostic): 0 none yet, 1 dense V x W,
 * 2 blocked (64-offset blocks handed out on first touch; picked when the dense table of one alignment would pass 64 MB). */
int poasta_last_search_table(void);
/* Device-resident DP-only rounds of the last poasta_msa_batch* on this thread: milliseconds of the dp_fill_packed_kernel launches
 * alone (CUDA events around each launch) and their number - the dominant kernel of that pipeline, for roofline figures. */
int poasta_last_batch_fill_ms(double *fill_ms, uint32_t *n_fill_launches);
/* Host<->device bytes copied by the last poasta_align_batch / poasta_msa_batch* on this thread. */
int poasta_last_batch_io(uint64_t *h2d_bytes, uint64_t *d2h_bytes);
/* Measured integer issue peaks of `device` in 10^12 lane-operations per second: 32-bit add, xor+min pairs (2 operations),
 * and the fused add+min the DP fill uses (DPX, counted as 2 operations). The DP fill's roofline is stated
 * against these; no reference counterpart. */
int poasta_measure_int_peak(int device, double *add_tops, double *min_tops, double *addmin_tops);

/* ---- POAGraph (graphs/poa.rs) -------------------------------------------------------------- */
typedef struct poasta_graph poasta_graph; /* opaque POAGraph<u32> */

poasta_graph *poasta_graph_new(void);                                  /* POAGraph::new  poa.rs:100 */
void poasta_graph_free(poasta_graph *g);
/* Frees n graphs (the handles a poasta_msa_batch call returned) on n_threads host threads (0 = as many as the host has, at most 16)
 * and nulls the entries. */
void poasta_graph_free_many(poasta_graph **graphs, uint32_t n, uint32_t n_threads);
uint32_t poasta_graph_node_count(const poasta_graph *g);               /* poa.rs:399 */
uint32_t poasta_graph_node_count_with_start_and_end(const poasta_graph *g);
uint32_t poasta_graph_edge_count(const poasta_graph *g);               /* poa.rs:409 */
uint32_t poasta_graph_num_sequences(const poasta_graph *g);
/* POAGraph::add_alignment_with_weights (poa.rs:171-321); rpos == NULL means alignment None;
 * weights == NULL means all ones (bin/poasta.rs:193), else weights_len must equal len
 * (POASTA_ERR_WEIGHTS_UNEQUAL, poa.rs:178-180). Pairs that name a node or a query position out of
 * range return POASTA_ERR_INVALID_ALIGNMENT (the reference panics on the index). */
int poasta_graph_add_alignment(poasta_graph *g, const char *name, const uint8_t *seq, uint32_t len,
                               const uint32_t *rpos, const uint32_t *qpos, uint32_t n_pairs,
                               const uint64_t *weights, uint32_t weights_len);
/* Borrowed CSR view, valid until the next mutation of `g`. */
int poasta_graph_get_view(poasta_graph *g, poasta_graph_view *view);
/* aligned_nodes of one node (poa.rs:374); returns count, writes up to cap ids. */
uint32_t poasta_graph_aligned_nodes(const poasta_graph *g, uint32_t node, uint32_t *out, uint32_t cap);

/* Writers/readers. Output buffers are malloc'ed by the library; release with poasta_free_buffer. */
int poasta_graph_to_fasta(const poasta_graph *g, char **out, size_t *len);      /* io/fasta.rs:69  */
int poasta_graph_to_gfa(const poasta_graph *g, int v1, char **out, size_t *len);/* io/graph.rs:245,374 */
/* dialect 0 = `view -O dot` (io/graph.rs:504), 1 = `align -O dot` (petgraph Dot, io/graph.rs:229) */
int poasta_graph_to_dot(const poasta_graph *g, int dialect, char **out, size_t *len);
int poasta_graph_save(const poasta_graph *g, char **out, size_t *len);          /* io/graph.rs:24  */
int poasta_graph_load(const uint8_t *data, size_t len, poasta_graph **out);     /* io/graph.rs:30  */
int poasta_graph_from_fasta_msa(const char *fasta, size_t len, poasta_graph **out); /* io/graph.rs:36 */
void poasta_free_buffer(void *p);

/* ---- progressive MSA (bin/poasta.rs:163-236 perform_alignment, batched over families) ------ */

/* Each family is one FASTA text; records are aligned and inserted in file order, family by
 * family independent. Round k of all families is one device batch. On success graphs_out[f]
 * (if non-NULL) receives the final graph of family f (caller frees), scores_out (if non-NULL,
 * n_total_seqs entries, family-major) the per-sequence alignment scores (0 for the first). */
int poasta_msa_batch(int device, const char *const *fasta, const size_t *fasta_len, uint32_t n_families,
                     const poasta_params *params, poasta_graph **graphs_out, uint32_t *scores_out,
                     uint64_t *cells_out);

/* Same loop against existing graphs (`poasta align -I graph.poasta`, bin/poasta.rs:303-317). */
int poasta_msa_batch_into(int device, poasta_graph **graphs, const char *const *fasta,
                          const size_t *fasta_len, uint32_t n_families, const poasta_params *params,
                          uint32_t *scores_out, uint64_t *cells_out);

const char *poasta_version(void);

#ifdef __cplusplus
}
#endif
#endif /* POASTA_B200_H */
