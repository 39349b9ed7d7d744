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
#define DRS_MAX_PEERS 15 /* peers a fused pivot step can store to (one host: <= 16 GPUs) */

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

/* The library keeps up to six large device blocks (matrices, work arrays of the separator build) of destroyed graphs in a
 * process-wide cache so that the next graph of similar size does not pay cudaFree + cudaMalloc of ~10 GB; this call
 * returns them to the driver. */
int drs_pool_trim(void);

/* Host-side helper of the bindings: vertex ids in the order of a locality-preserving DEVICE numbering (order_out[k] =
 * caller's id of device vertex k; Morton curve over the ranks of the lng / lat values).  The kernels index distance rows by
 * vertex id; a caller that numbers its vertices this way before drs_graph_create (and translates ids at the boundary, as
 * routing.RoadGraph does) gets ~10 % faster builds than with row-major ids and ~40 % faster than with random ones.
 * DRS_ERR_ARG: coordinates unusable (non-finite or all equal). */
int drs_locality_order(int32_t n, const double* lng_host, const double* lat_host, int32_t* order_out_host);

/* n, m, npad (matrix leading dimension), mode, apsp_valid, device */
int drs_graph_info(const drs_graph* g, int32_t* n, int32_t* m, int32_t* npad, int32_t* mode,
                   int32_t* apsp_valid, int32_t* device);

/* ------------------------------------------------------------------- APSP --
 * Builds the all-pairs travel-time matrix with the algorithm selected by drs_apsp_set_algorithm (default AUTO: the
 * separator-partitioned min-plus build when the graph carries a partition with small separators and integer travel
 * times, else N batched delta-stepping searches when arcs/vertex <= 16, else the blocked min-plus Floyd-Warshall), and -- only in
 * DRS_PRED_MATRIX mode -- the canonical predecessor-edge matrix; by default predecessors are evaluated lazily by the
 * path queries.  Replaces every per-query Dijkstra the reference issues
 * (lib/RoutingEngine.py:87,199,212; lib/optimUtils.py:528,531,659).
 * Single GPU; the multi-GPU build drives the drs_apsp_rows_sep / drs_apsp_rows_sssp / drs_fw_* entry points below from
 * one process per GPU.  Blocks until the matrix is complete.
 */
int drs_apsp_build(drs_graph* g, int32_t mode);

/* Device pointers of the resident matrices (row-major, leading dimension
 * npad).  dist is float (DRS_MODE_F32) or int32 (DRS_MODE_I32); unreachable =
 * +inf / DRS_I32_INF.  pred[s*npad+t] is the id of the last edge on the
 * canonical shortest path s->t (-1 for s == t or unreachable); *pred_dev is NULL
 * unless DRS_PRED_MATRIX was selected. */
#define DRS_I32_INF 0x3fffffff
int drs_apsp_matrices(const drs_graph* g, void** dist_dev, int32_t** pred_dev, int64_t* ld);

/* Adopt externally built matrices (multi-GPU build, or a cached matrix): the
 * library takes no ownership.  pred_dev may be NULL (lazy predecessors). */
int drs_apsp_adopt(drs_graph* g, int32_t mode, void* dist_dev, int32_t* pred_dev, int64_t ld);

/* Copy rows [row0, row0+nrows) x columns [0, n) of the distance matrix to the
 * host, converted to double (inf for unreachable).  Serves
 * G.shortest_paths(src, tgt, weights=) (lib/optimUtils.py:528,531,659). */
int drs_apsp_rows_to_host(const drs_graph* g, int32_t row0, int32_t nrows, double* out_host);

/* Algorithm used by drs_apsp_build.  FW: blocked min-plus Floyd-Warshall (N^3 work, ALU
 * bound, independent of sparsity).  SSSP: N delta-stepping searches over the CSR (about
 * N*E work, memory-latency bound; far less work on sparse road networks).  Both give the
 * same matrix bit for bit on integer weights.  SEP: the blocked Floyd-Warshall in nested-dissection
 * order over a vertex-separator partition of the map (drs_graph_set_partition): about N^2 * |S_j|
 * relaxations, |S_j| = separators around a part, ALU bound, the matrix written once; it re-associates the
 * sums: the same matrix bit for bit on integer travel times, last-bit differences on float ones (within the
 * 1e-5 float maps are held to; DRS_SEP_FLOAT=0 in the environment keeps float maps off it).
 * AUTO picks SEP when the graph has a partition with small separators, else SSSP when arcs/vertex <= 16,
 * else FW. */
#define DRS_APSP_FW 0
#define DRS_APSP_SSSP 1
#define DRS_APSP_AUTO 2
#define DRS_APSP_SEP 3
int drs_apsp_set_algorithm(drs_graph* g, int32_t algorithm);

/* ------------------------------------------------- separator partition --
 * Host-side helper of the bindings (no device work): a part-major device numbering of the vertices.
 * order_out[k] = caller's id of device vertex k; device ids [part_ptr_out[j], part_ptr_out[j+1]) are the
 * interior of part j (j < *nparts_out), ids [part_ptr_out[*nparts_out], n) the separator vertices; no edge
 * joins the interiors of two parts.  Parts hold about `target` vertices: recursive coordinate bisection
 * when lng/lat are usable (they may be NULL), else bisection of breadth-first level structures;
 * separators = a greedy vertex cover of the cut edges.  Parts follow the bisection tree, so the numbering
 * also serves as the locality numbering of drs_locality_order.  Replaces nothing in the reference (vertex
 * ids are igraph's load order there, lib/RoutingEngine.py:48-60); the bindings translate ids at every call. */
int drs_partition_order(int32_t n, int32_t m, const int32_t* src, const int32_t* dst, const double* lng,
                        const double* lat, int32_t target, int32_t* order_out, int32_t* part_ptr_out,
                        int32_t max_parts, int32_t* nparts_out, int32_t* sep_part_ptr_out, int32_t* n_sep_parts_out);
/* Second level (sep_part_ptr_out may be NULL): the separator vertices themselves are numbered in groups that follow
 * the upper levels of the bisection tree, the "top" separators (those with a neighbour in another group) last;
 * sep_part_ptr_out[0..n_sep_parts] are the group boundaries relative to the first separator id.  (groups, top) is a
 * separator partition of the separator graph, which lets the build close that graph the way it closes the map. */
/* Declares the partition of a graph created in that numbering (part_ptr as above, nparts + 1 entries; nparts = 0
 * removes it; n_sep_parts = 0 / sep_part_ptr = NULL: no second level).  DRS_ERR_ARG if an edge joins two interiors.
 * Builds the device tables of the SEP build. */
int drs_graph_set_partition(drs_graph* g, int32_t nparts, const int32_t* part_ptr, int32_t n_sep_parts,
                            const int32_t* sep_part_ptr);
/* groups and top separators of the second level (0, 0 without one); whether the last build ran its output tiles on the
 * 16-bit path (VIADDMNMX.U16x2: fp32 matrices of maps whose distances all stay below 2^15, decided per build from the
 * largest vertex->separator and in-part distances; twice the relaxation rate, same matrix) */
int drs_sep_info2(const drs_graph* g, int32_t* n_sep_parts, int32_t* n_top, int32_t* used16);
/* parts, separator vertices, sum and maximum of the per-part separator counts (padded), and whether AUTO would use it */
int drs_sep_info(const drs_graph* g, int32_t* nparts, int32_t* nsep, int32_t* ktot, int32_t* kmax, int32_t* usable);
/* Rows [row0, row0 + nrows) (multiples of DRS_FW_TILE) of the matrix by the SEP build into local_dev (nrows x ld, ld = npad):
 * the sharded form used by the multi-GPU build (every rank closes the parts and the separator graph, then expands
 * its own rows; no collective).  Same role as drs_apsp_rows_sssp. */
int drs_apsp_rows_sep(const drs_graph* g, int32_t mode, void* local_dev, int64_t ld, int32_t row0, int32_t nrows,
                      void* stream);
/* relaxations (one add + one min each) a SEP build of rows [row0, row0 + nrows) executes (nrows = 0: the whole matrix):
 * [0] part closures, [1] separator closure, [2] vertex->separator rows, [3] output tiles as launched (padded), [4] output
 * tiles without padding = the algorithmic work of the dominant kernel: rows * sum_j n_j * |S_j|, halved for a whole
 * undirected matrix (mirrored tiles) */
int drs_sep_work(const drs_graph* g, int32_t row0, int32_t nrows, double* relax5);
/* milliseconds of the last SEP build by stage (the library's own CUDA events): [0] fill + scatter of the closed parts,
 * [1] part closures, [2] extraction (A', A_j, separator-graph blocks), [3] separator closure (second level included),
 * [4] gathers, [5] vertex->separator rows + the spread into the per-part ranges and the separator columns, [6] output
 * tiles, [7] unused (the separator columns are written by the spread) */
int drs_apsp_sep_profile(const drs_graph* g, double* ms8);

/* Predecessors.  LAZY (default): no predecessor matrix; path queries evaluate the canonical rule on the
 * spot from the distance row and the target's in-arcs (a handful of extra gathers per hop, half the HBM
 * footprint and no predecessor pass at build time).  MATRIX: materialise pred[s*ld+t] during the build. */
#define DRS_PRED_LAZY 0
#define DRS_PRED_MATRIX 1
int drs_apsp_set_pred_mode(drs_graph* g, int32_t pred_mode);

/* --------------------------------------------------------- batched SSSP --
 * Distances from nsrc sources (delta-stepping, one CTA per source).  Replaces the same
 * igraph Dijkstra calls when the N x N matrix does not fit in HBM (BASELINE config C5).
 * _dev: row k of dist_dev (float or int32 per mode, leading dimension ld >= n) receives
 * the distances from sources_host[k]; delta <= 0 selects the default bucket width (six
 * times the mean arc weight, the measured sweet spot on grid cities; DRS_SSSP_DELTA_MULT overrides).
 * mode DRS_MODE_I32 requires integer travel times below 1e6 (checked by every entry point that takes a mode).  The host variant returns nsrc x n doubles (inf = unreachable). */
int drs_sssp_batch_dev(const drs_graph* g, int32_t mode, int32_t nsrc, const int32_t* sources_host,
                       void* dist_dev, int64_t ld, double delta, void* stream);
int drs_sssp_batch(const drs_graph* g, int32_t mode, int32_t nsrc, const int32_t* sources_host,
                   double* dist_out_host);
/* Rows [row0, row0+nrows) of the padded all-pairs matrix by SSSP (sources are independent:
 * the multi-GPU build shards them across ranks with no collective). */
int drs_apsp_rows_sssp(const drs_graph* g, int32_t mode, void* local_dev, int64_t ld, int32_t row0,
                       int32_t nrows, double delta, void* stream);

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
/* ---- fused pivot step + broadcast over NVLink peer memory (multi-GPU FW) ----
 * Each rank allocates one communication block with drs_p2p_alloc (device memory + a CUDA IPC
 * handle the other ranks open with drs_p2p_open).  Layout of a block (see drs_p2p_layout):
 * signal words, acknowledgement words, a time-out word and two DRS_FW_TILE x ld panel buffers.
 * drs_fw_pivot_row_p2p is drs_fw_pivot_row whose tile epilogues ALSO store every finished tile of
 * the pivot panel straight into each peer's panel buffer (P2P stores: the transfer overlaps the
 * tile arithmetic), followed by a release of the peers' signal word for round kb.  Receivers order
 * their stream behind the signal with drs_fw_wait_word; before a panel buffer is overwritten the
 * owner waits for the peers' acknowledgements of the round that last used it.  Every wait has a
 * time-out (it sets the block's time-out word instead of hanging). */
int drs_p2p_alloc(int64_t bytes, void** dev_ptr, unsigned char* ipc_handle_64);
int drs_p2p_open(const unsigned char* ipc_handle_64, void** peer_ptr);
int drs_p2p_close(void* peer_ptr);
int drs_p2p_free(void* dev_ptr);
int drs_fw_pivot_row_p2p(int32_t mode, void* panel_dev, int64_t ld, int32_t kb, int32_t n_peers,
                         void* const* peer_panels, int32_t* const* peer_signal_words, int32_t signal_value,
                         void* stream);
/* stream-ordered: store `value` (system-scope release) into n remote or local words */
int drs_fw_post_words(int32_t n, int32_t* const* words, int32_t value, void* stream);
/* stream-ordered: wait until each of the n words is >= value; on time-out set *timed_out_dev = 1 */
int drs_fw_wait_words(int32_t n, const int32_t* const* words, int32_t value, int64_t timeout_ms,
                      int32_t* timed_out_dev, void* stream);

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

/* Scalar form of drs_query_pairs -- the reference asks one pair at a time (lib/Agents.py:898-923, 756, 1073): ONE kernel
 * that reads the pair from, and writes the answer to, pinned host memory mapped into the device; no allocation, no copy. */
int drs_query_pair(const drs_graph* g, int32_t s, int32_t t, double* tt_out, double* len_out, int32_t* hops_out);

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

/* --------------------------------------------------------------- dispatch --
 * Per-tick vehicle-by-request evaluation.  Replaces the Python loops of
 * AlonsoMora.generatePairwiseGraph / generateRTVGraph with routing_method
 * "enumerate" (lib/Optimization.py:104-357) and the functions they call:
 * findOptimalRouting, minimalDelayPath, permutePassengerOrder,
 * computeRoutingCost, shortestPathTime (lib/optimUtils.py:41-160,432-557,656-664).
 *
 * A "tick" handle owns device copies of the vehicle and request batches, so the
 * evaluation itself runs on inputs resident in HBM.  All batch arrays are HOST
 * pointers copied at call time (callers mutate Veh / Req freely between ticks).
 *
 * Stop record (one per pickup / drop-off already on a vehicle's plan, in
 * createLocations order, lib/optimUtils.py:94-108: drop-offs of on-board
 * requests first, then (pickup, drop-off) of each waiting request):
 *   node   road-graph vertex            rid   request id (opaque to the kernel)
 *   pod    +1 pickup / -1 drop-off      prev  index WITHIN THE VEHICLE'S LIST of
 *   lb,ub  absolute time window               the stop that must precede it, or -1
 */
#define DRS_MAX_STOPS 16       /* stops per evaluated trip (new + planned) */
#define DRS_MAX_TRIP_REQS 4    /* new requests per trip */
#define DRS_INF_ROUTE_COST 1e6 /* lib/optimUtils.py:156,437: (1e6, None) == infeasible */

typedef struct drs_tick drs_tick;

typedef struct {
  int32_t n_veh;
  const int32_t* start_node;  /* veh.get_next_node() vertex (lib/Optimization.py:124,130) */
  const double* start_delay;  /* its t: time to reach that vertex (veh["t"]) */
  const int32_t* capacity;    /* veh.K */
  const int32_t* onboard;     /* len(pax_reqs) */
  const int32_t* stop_off;    /* n_veh + 1 offsets into the stop arrays */
  const int32_t* stop_node;
  const int32_t* stop_rid;
  const int8_t* stop_pod;
  const int8_t* stop_prev;
  const double* stop_lb;
  const double* stop_ub;
} drs_vehicle_batch;

typedef struct {
  int32_t n_req;
  const int32_t* rid;
  const int32_t* o_node;
  const int32_t* d_node;
  const double* lb_p; /* Cep - T  (relative lower bound, lib/Optimization.py:117) */
  const double* ub_p; /* Clp */
  const double* lb_d; /* Ced */
  const double* ub_d; /* Cld */
} drs_request_batch;

/* status of one (vehicle, request) pair of the RV stage */
#define DRS_RV_FEASIBLE 0   /* edge with (cost, route)                 lib/Optimization.py:212-214 */
#define DRS_RV_SKIPPED 1    /* T + tt(veh -> origin) > Clp             lib/Optimization.py:183-186 */
#define DRS_RV_SAVED 2      /* empty-dummy-vehicle check failed        lib/Optimization.py:193-202,217-218 */
#define DRS_RV_INFEASIBLE 3 /* no feasible ordering                    lib/Optimization.py:215-216 */

int drs_tick_create(drs_graph* g, drs_tick** out);
int drs_tick_destroy(drs_tick* t);
int drs_tick_set_vehicles(drs_tick* t, const drs_vehicle_batch* vehicles);
int drs_tick_set_requests(drs_tick* t, const drs_request_batch* requests);

/* flags for the evaluation calls */
#define DRS_EVAL_PAIRWISE_CHECKS 1 /* findOptimalRouting's pairwise pre-filter when >= 3 requests are
                                      involved (lib/optimUtils.py:127-156); "enumerate" sets it, the
                                      tabu front end (lib/optimUtils.py:163-212) has it commented out */

/* RV stage for every (request, vehicle) pair, results left on the device.
 * counters_host[4] receives the number of pairs per status; *n_edges the number
 * of feasible edges. */
int drs_tick_rv_eval(drs_tick* t, double T, int32_t flags, int32_t* n_edges, int64_t* counters_host);
/* Copy the feasible edges to the host, sorted by (request index, vehicle index):
 * order[e*DRS_MAX_STOPS + k] is the k-th visited stop, numbered 0 = new pickup,
 * 1 = new drop-off, 2.. = the vehicle's planned stops in batch order. */
int drs_tick_rv_fetch(drs_tick* t, int32_t capacity_edges, int32_t* veh_idx, int32_t* req_idx,
                      double* cost, uint8_t* order, int32_t* n_stops);

/* Explicit trip list (RR edges, RTV trips of size >= 2, or single pairs):
 * task k = vehicle task_veh[k] (-1: no vehicle, route starts at the first stop at
 * time T, lib/optimUtils.py:160) with the task_nreq[k] new requests
 * task_req[k*DRS_MAX_TRIP_REQS ..] (indices into the request batch).  Implements
 * findOptimalRouting (lib/optimUtils.py:113-160).  cost_host[k] ==
 * DRS_INF_ROUTE_COST means infeasible.  Stop numbering in order_host: new
 * requests' (pickup, drop-off) pairs in task order, then the vehicle's stops.
 * Among equal-cost orderings the lexicographically smallest (by that numbering)
 * is returned.  The search is exact and, like the reference's "enumerate",
 * exponential in the stop count: a trip that needs more than 2*10^7 search steps
 * makes the call fail with DRS_ERR_CAPACITY ("budget") -- the same holds for
 * drs_tick_rv_eval.  Environment (tuning only, results never depend on them):
 * DRS_EVAL_SPLIT2 / DRS_EVAL_SPLIT3 = stop counts from which a trip's search is
 * split into parallel searches by its first two / three stops (default 5 / 8). */
int drs_tick_trip_eval(drs_tick* t, double T, int32_t flags, int32_t n_tasks, const int32_t* task_veh,
                       const int32_t* task_nreq, const int32_t* task_req, double* cost_host,
                       uint8_t* order_host, int32_t* n_stops_host);

/* One level of the RTV trip enumeration for the whole fleet, on the HOST (set logic only; the candidates are then costed
 * by drs_tick_trip_eval).  Replaces the per-vehicle loops of AlonsoMora.generateRTVGraph, lib/Optimization.py:265-288
 * (k = 2) and :291-347 (k >= 3), including the order in which candidates appear and the reference's early exit from the
 * inner loop when a (k-1)-subset of a union is not a trip of the vehicle (:325-327).
 *   key_off[n_veh + 1], key_req[n_keys * (k-1)]: every vehicle's feasible trips of size k-1 in the order the reference
 *     holds them (members ascending within a trip); veh_on[n_veh] (may be NULL = all): vehicles to enumerate for;
 *   rr_bits: the request-request graph as n_req rows of ceil(n_req / 64) 64-bit words (bit b of row a: a and b share
 *     an edge; symmetric).
 * Writes the first cap_tasks candidates (vehicle index, k members ascending) and always returns the full count in
 * *n_tasks (call again with a larger buffer if it exceeds cap_tasks); *n_saved = candidates the reference counts as
 * "skipped via incomplete subgraphs / previous trip length checks". */
int drs_rtv_level(int32_t n_veh, int32_t k, const int32_t* key_off, const int32_t* key_req, const uint8_t* veh_on,
                  int32_t n_req, const uint64_t* rr_bits, int64_t cap_tasks, int32_t* task_veh, int32_t* task_req,
                  int64_t* n_tasks, int64_t* n_saved);

/* ------------------------------------------------------ legacy insertion --
 * The (vehicle, pickup slot, drop-off slot) insertion heuristic: replaces
 * Model.insert_heuristics / test_constraints_get_cost (lib/Agents.py:1451-1572;
 * dead code at the reference's HEAD, semantics in SURVEY.md section 3.4).  For each
 * new request the vehicles are scanned in order, the request's pickup is tried
 * after `i` planned stops and its drop-off after `j-1`, and the LAST candidate
 * whose running cost never exceeds the incumbent bound wins (lib/Agents.py:1470-1476),
 * with the violation-code loop breaks of lib/Agents.py:1479-1482.
 *
 * The request table holds every request referenced by a plan plus the new ones;
 * plan stops and new requests point into it by row.
 */
typedef struct {
  int32_t n_veh;
  const int32_t* start_node;   /* vehicle position if it is a vertex, else next downstream vertex */
  const double* start_delay;   /* 0, or the remaining time of the current step (lib/Agents.py:1508-1516) */
  const int32_t* onboard;      /* veh.n */
  const double* clock;         /* veh.T */
  const int32_t* capacity;     /* veh.K */
  const double* cost;          /* veh.c */
  const int32_t* stop_off;     /* n_veh + 1 */
  const int32_t* stop_node;    /* plan stops in route order (leg.tlng, leg.tlat) */
  const int32_t* stop_req;     /* row of the request table */
  const int8_t* stop_pod;
  /* What a commit (Veh.build_route, lib/Agents.py:148-297) needs beyond the scan; each may be NULL.
   * pos_node: vertex the vehicle stands on, -1 between vertices (NULL: start_node where start_delay == 0).
   * keep_*:   the first step of the current route -- its end vertex, remaining time `t` and expected time `et` --
   *           which clear_route_but_keep_next_node keeps in front of a new route (lib/Agents.py:319-336);
   *           keep_node -1 = the vehicle has no route (NULL: start_node / start_delay for vehicles under way or
   *           with a plan, none otherwise). */
  const int32_t* pos_node;
  const int32_t* keep_node;
  const double* keep_t;
  const double* keep_et;
} drs_plan_batch;

typedef struct {
  int32_t n_rows;
  const int32_t* o_node;
  const int32_t* d_node;
  const double* Cep;
  const double* Clp;
  const double* Cld;  /* current value; only read for on-board requests */
  const double* Ts;
  const double* Tr;
  const double* pwt;
} drs_request_table;

typedef struct {
  double max_detour;      /* lib/Constants.py:144 MAX_DETOUR = 1.5 */
  double coef_inveh;      /* lib/Constants.py:163 */
  double coef_wait;       /* lib/Constants.py:164 */
  double coef_extra_wait; /* lib/Constants.py:165 */
} drs_insertion_params;

int drs_tick_set_plans(drs_tick* t, const drs_plan_batch* plans, const drs_request_table* table,
                       const drs_insertion_params* params);
/* Evaluates every (vehicle, i, j) candidate for each new request (rows into the
 * request table) and folds them in the reference's scan order.  Outputs (host,
 * one per new request): winning vehicle index (-1: rejected), i, j (as in
 * lib/Agents.py:1466-1469) and the cost increase dc.  *n_candidates = number of
 * (vehicle, i, j) triples evaluated. */
int drs_tick_insertion_eval(drs_tick* t, int32_t n_new, const int32_t* new_rows_host, int32_t* best_veh,
                            int32_t* best_i, int32_t* best_j, double* best_dc, int64_t* n_candidates);

/* The queue of Model.insertion_heuristics (lib/Agents.py:1429-1448): the new requests are inserted ONE AFTER ANOTHER in
 * the given order, each scan seeing the plans the previous winners' Veh.build_route left behind -- the route with the
 * pickup at i and the drop-off at j, veh.c (lib/Agents.py:256-270) and the predicted waiting time pwt of the requests
 * still to be picked up (lib/Agents.py:165-220).  T_now is the tick time handed to build_route.  All of it happens on
 * the device: after a request is folded, its winner's plan is patched, that vehicle is re-prepared and its column is
 * re-evaluated for the requests still queued.  Outputs as drs_tick_insertion_eval; the plans stay patched (fetch them
 * with drs_tick_fetch_plans).  Plans hold at most DRS_INS_MAX_PLAN stops; a vehicle whose plan could not take two more
 * cannot be patched (the call fails instead of picking another vehicle).
 * flags: DRS_QUEUE_TRACK_CLD also reproduces the side effect of lib/Agents.py:1550,1554 -- test_constraints_get_cost
 * assigns Req.Cld at every pickup of every candidate it walks, so after a scan each waiting request carries the value
 * of the LAST walk that reached its pickup; one extra kernel per request repeats the bounded scan of the vehicles in
 * reach and leaves those values in the request table (drs_tick_fetch_plans row_Cld).  Within a tick the values are
 * rewritten before every read, so winners do not depend on the flag; it matters for on-board riders of later ticks. */
#define DRS_QUEUE_TRACK_CLD 1
#define DRS_INS_MAX_PLAN 16
int drs_tick_insertion_queue(drs_tick* t, int32_t n_new, const int32_t* new_rows_host, double T_now, int32_t flags,
                             int32_t* best_veh, int32_t* best_i, int32_t* best_j, double* best_dc, int64_t* n_candidates);
/* One commit by hand (the building block of the queue): request row req_row joins vehicle veh with its pickup at route
 * position i and its drop-off at j (lib/Agents.py:1468-1469), then Veh.build_route's bookkeeping as above. */
int drs_tick_commit_insertion(drs_tick* t, int32_t req_row, int32_t veh, int32_t i, int32_t j, double T_now);
/* The plans as they are now (host outputs, any may be NULL): veh_cost[n_veh] = veh.c, plan_len[n_veh], stop_req / stop_pod
 * [n_veh x stops_per_vehicle] (stops_per_vehicle >= DRS_INS_MAX_PLAN), and per request-table row Cld and pwt. */
int drs_tick_fetch_plans(drs_tick* t, double* veh_cost, int32_t* plan_len, int32_t* stop_req, int8_t* stop_pod,
                         int32_t stops_per_vehicle, double* row_Cld, double* row_pwt);

/* ---------------------------------------------------------- instrumentation --
 * The CUDA stream (cudaStream_t as void*) every call on this graph launches on:
 * lets the host time kernels with events on the launching stream. */
int drs_graph_stream(const drs_graph* g, void** stream);
/* Device time (ms, CUDA events on the graph's stream) of the last evaluation kernels of a tick:
 * ms[0] rv_filter, ms[1] rv_eval (drs_tick_rv_eval); ms[2] insertion scan = plan prep + reach test + evaluation,
 * ms[3] insertion fold (drs_tick_insertion_eval) or the whole fold / commit / re-scan loop (drs_tick_insertion_queue). */
int drs_tick_last_ms(const drs_tick* t, double* ms);
/* Device times of the last insertion call by kernel (ms): [0] plan prep, [1] reach test, [2] evaluation of the surviving
 * pairs, [3] fold (drs_tick_insertion_eval), [4] fold / commit / re-scan loop (drs_tick_insertion_queue); ms[5] = number
 * of (request, vehicle) pairs that survived the reach test. */
int drs_tick_last_insertion_ms(const drs_tick* t, double* ms);
/* Number of kernels this library has launched in this process so far. */
int64_t drs_launch_count(void);
/* Number of device allocations (cudaMalloc) this library has made in this process so far: a steady-state tick or query
 * makes none (tests/test_gpu_insertion.py, tests/test_gpu_routing.py). */
int64_t drs_alloc_count(void);
/* Device time of the last drs_apsp_build by phase, from CUDA events on the
 * graph's stream: ms[0] init (fill + edge scatter), ms[1] phase 1 + row part of
 * phase 2 (all rounds), ms[2] column part of phase 2 + phase 3 (all rounds),
 * ms[3] predecessor pass; *rounds = number of pivot blocks. */
int drs_apsp_last_profile(const drs_graph* g, double* ms, int32_t* rounds);

/* ------------------------------------------------------------- ALU peak ---
 * Register-resident add+min stream used as the ALU roofline denominator
 * (SURVEY.md section 8d).  Returns relaxations per second of the stream
 * 2xFADD+FMNMX3 (variant 0), 2xVIADDMNMX int32 (variant 1) or FADD2+FMNMX3 (variant 2). */
int drs_alu_peak(int32_t variant, int32_t iters, double* relax_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_H */
