/*
 * drs_b200.h -- C-ABI of the B200-native routing-and-dispatch core for drs-abm.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point returns an int status (0 = ok, <0 = error; the message is
 * available from drs_last_error()).  Nothing throws across the boundary.
 *
 * Each function cites the reference interface it replaces (paths relative to
 * the nbailey/drs-abm checkout).  The reference reaches its native code through
 * python-igraph (C Dijkstra) and gurobipy; this library replaces the igraph
 * calls on the routing path and the Python cost loops on the dispatch path.
 *
 * Pointer conventions: parameters named *_host are host pointers (pageable or
 * pinned); parameters named *_dev are device pointers on the graph's device.
 * `stream` parameters are cudaStream_t passed as void* (NULL = default stream).
 *
 * Threading: one caller thread per handle (the reference is single-threaded).
 */
#ifndef DRS_B200_H
#define DRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct drs_graph drs_graph;

/* status codes */
#define DRS_OK 0
#define DRS_ERR_ARG (-1)      /* bad argument (maps to ValueError) */
#define DRS_ERR_CUDA (-2)     /* CUDA runtime failure (maps to RuntimeError) */
#define DRS_ERR_STATE (-3)    /* call made in the wrong state, e.g. query before build */
#define DRS_ERR_NOMEM (-4)
#define DRS_ERR_RANGE (-5)    /* node / edge index out of range */
#define DRS_ERR_CAPACITY (-6) /* caller-provided output buffer too small */

/* arithmetic mode of the travel-time matrix */
#define DRS_MODE_F32 0 /* fp32 add+min; exact for integer weights below 2^24 */
#define DRS_MODE_I32 1 /* int32 add+min; requires integer-valued weights      */

/* tile edge of the blocked Floyd-Warshall: matrices are padded to a multiple */
#define DRS_FW_TILE 128

const char* drs_last_error(void);
int drs_version(void);

/* ------------------------------------------------------------------ graph --
 * Replaces ig.Graph.Read_GML + the graph object `RoutingEngine.G`
 * (lib/RoutingEngine.py:31-35).  The host parses the GML; this call uploads the
 * edge list as device CSR (out- and in-adjacency).  For directed == 0 every
 * edge is traversable both ways (igraph undirected graph).  `tt` is the
 * `ttweight` edge attribute (default 'ttmedian', lib/RoutingEngine.py:31,41),
 * `length` the 'length' attribute (lib/RoutingEngine.py:92,203).
 * Weights must be finite and > 0.
 */
int drs_graph_create(int32_t n, int32_t m, int32_t directed, const int32_t* src_host,
                     const int32_t* dst_host, const double* tt_host, const double* length_host,
                     drs_graph** out);

/* Replaces the in-place weight mutation `G.es[idx]["ttmedian"] = x`
 * (optimizationTesting.py:221-226, generateMetadata.py:42-43).  Invalidates
 * the travel-time matrix; the next build recomputes it. */
int drs_graph_set_weights(drs_graph* g, const double* tt_host);

int drs_graph_destroy(drs_graph* g);

/* n, m, npad (matrix leading dimension), mode, apsp_valid, device */
int drs_graph_info(const drs_graph* g, int32_t* n, int32_t* m, int32_t* npad, int32_t* mode,
                   int32_t* apsp_valid, int32_t* device);

/* ------------------------------------------------------------------- APSP --
 * Builds the all-pairs travel-time matrix with the blocked min-plus
 * Floyd-Warshall and then the canonical predecessor-edge matrix.  Replaces
 * every per-query Dijkstra the reference issues
 * (lib/RoutingEngine.py:87,199,212; lib/optimUtils.py:528,531,659).
 * Single GPU; the multi-GPU build drives the drs_fw_* entry points below from
 * one process per GPU.  Blocks until the matrices are complete.
 */
int drs_apsp_build(drs_graph* g, int32_t mode);

/* Device pointers of the resident matrices (row-major, leading dimension
 * npad).  dist is float (DRS_MODE_F32) or int32 (DRS_MODE_I32); unreachable =
 * +inf / DRS_I32_INF.  pred[s*npad+t] is the id of the last edge on the
 * canonical shortest path s->t (-1 for s == t or unreachable). */
#define DRS_I32_INF 0x3fffffff
int drs_apsp_matrices(const drs_graph* g, void** dist_dev, int32_t** pred_dev, int64_t* ld);

/* Adopt externally built matrices (multi-GPU build, or a cached matrix): the
 * library takes no ownership. */
int drs_apsp_adopt(drs_graph* g, int32_t mode, void* dist_dev, int32_t* pred_dev, int64_t ld);

/* Copy rows [row0, row0+nrows) x columns [0, n) of the distance matrix to the
 * host, converted to double (inf for unreachable).  Serves
 * G.shortest_paths(src, tgt, weights=) (lib/optimUtils.py:528,531,659). */
int drs_apsp_rows_to_host(const drs_graph* g, int32_t row0, int32_t nrows, double* out_host);

/* ---------------------------------------------------- FW building blocks --
 * Used by the single-GPU build and by the row-block-sharded multi-GPU build
 * (SURVEY.md section 8e).  `local` is this rank's shard: rows
 * [row0, row0+nrows) of the padded matrix, leading dimension ld == npad; nrows
 * is a multiple of DRS_FW_TILE.  `panel` is a DRS_FW_TILE x ld buffer holding
 * the pivot row panel of round kb (the owner's own rows, or the broadcast
 * copy).
 */
/* local[i - row0][j] = (i == j ? 0 : min weight of edges i->j, else inf) */
int drs_fw_init(const drs_graph* g, int32_t mode, void* local_dev, int64_t ld, int32_t row0,
                int32_t nrows, void* stream);
/* owner: close the diagonal tile of block kb in place inside panel_dev (phase 1)
 * and update the rest of the pivot row panel (phase 2, row part). */
int drs_fw_pivot_row(int32_t mode, void* panel_dev, int64_t ld, int32_t kb, void* stream);
/* every rank: phase 2 (column part) for its rows and phase 3 for all remaining
 * tiles.  skip_local_block >= 0 names the local block-row that IS the pivot
 * panel (owner only, -1 otherwise).  block_lo/block_hi restrict the local
 * block-rows processed (look-ahead splits); pass 0 and nrows/DRS_FW_TILE for all. */
int drs_fw_update(int32_t mode, void* local_dev, int64_t ld, int32_t nrows, const void* panel_dev,
                  int32_t kb, int32_t skip_local_block, int32_t block_lo, int32_t block_hi,
                  void* stream);
/* canonical predecessor edges for rows [row0, row0+nrows) from finished
 * distances (rule: DESIGN.md "canonical predecessor"). */
int drs_pred_build(const drs_graph* g, int32_t mode, const void* local_dev, int64_t ld,
                   int32_t row0, int32_t nrows, int32_t* pred_local_dev, void* stream);

/* ---------------------------------------------------------------- queries --
 * drs_query_pairs replaces RoutingEngine.get_duration / get_distance /
 * get_distance_duration (lib/RoutingEngine.py:195-226): for each (s, t) it
 * walks the canonical path and returns the fp64 left-to-right sums of the
 * travel-time and length attributes, and the hop count (-1 = unreachable, 0 =
 * same node).  All pointers are host pointers; any output may be NULL.
 */
int drs_query_pairs(const drs_graph* g, int32_t npairs, const int32_t* s_host,
                    const int32_t* t_host, double* tt_out_host, double* len_out_host,
                    int32_t* hops_out_host);

/* Matrix look-up only (no path walk): dist[s,t] as double, inf if unreachable.
 * Replaces shortestPathTime (lib/optimUtils.py:656-664). */
int drs_query_matrix(const drs_graph* g, int32_t npairs, const int32_t* s_host,
                     const int32_t* t_host, double* tt_out_host);

/* Replaces G.get_shortest_paths(o, to=d, weights=, output='epath')[0]
 * (lib/RoutingEngine.py:87): edge ids in travel order.  *nedges: in = capacity
 * of edge_ids_host, out = path length (0 when s == t or unreachable;
 * *reachable tells which). */
int drs_path(const drs_graph* g, int32_t s, int32_t t, int32_t* edge_ids_host, int32_t* nedges,
             int32_t* reachable);

/* ------------------------------------------------------------- ALU peak ---
 * Register-resident add+min stream used as the ALU roofline denominator
 * (SURVEY.md section 8d).  Returns relaxations per second of the fp32
 * FADD+FMNMX pair stream (variant 0) or the int32 stream (variant 1). */
int drs_alu_peak(int32_t variant, int32_t iters, double* relax_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_H */
