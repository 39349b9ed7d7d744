// Per-tick vehicle-by-request evaluation (K3 in SURVEY.md section 2).
//
// Replaces the Python loops of AlonsoMora.generatePairwiseGraph / generateRTVGraph
// with routing_method "enumerate" (lib/Optimization.py:104-357) and the functions
// they call (lib/optimUtils.py: findOptimalRouting :113-160, minimalDelayPath
// :432-453, permutePassengerOrder :41-81, computeRoutingCost :456-557,
// shortestPathTime :656-664).  All arithmetic that decides feasibility or cost is
// fp64 in the reference's operation order, so integer-weighted inputs give
// bit-identical costs; travel times are gathered from the resident matrix.
//
// Kernels:
//   rv_filter_kernel   one thread per (request, vehicle) pair, vehicles fastest: the
//                      reachability pre-filter (lib/Optimization.py:183-186) and the
//                      empty-dummy-vehicle check (:193-202); survivors are appended
//                      to a compact task list (warp-aggregated atomic).
//   rv_eval_kernel     eight lanes per surviving pair: findOptimalRouting of the
//                      request inserted into the vehicle's plan.
//   trip_eval_kernel   eight lanes per explicit trip (RR edges, RTV trips): gathers the
//                      (S+1)^2 travel times among the trip's vertices, runs the pairwise
//                      pre-checks (lib/optimUtils.py:127-156) and a depth-first
//                      enumeration of the precedence-respecting orderings in
//                      lexicographic order with exact feasibility / cost-bound pruning.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "drs_common.cuh"
#include "drs_tick.cuh"

namespace {

constexpr int MAXS = DRS_MAX_STOPS;
constexpr int MAXR = DRS_MAX_TRIP_REQS;

// One evaluation problem in thread-local storage.  Vertex 0 of `tt` is the start.
struct Trip {
  int ns;                 // stops
  int has_start;          // 1: vertex 0 is the vehicle's next node; 0: route starts at its first stop
  double t_start;         // T + veh["t"]  (or T without a vehicle)
  double cap;             // capacity (1e6 when unconstrained, lib/optimUtils.py:432)
  int pax0;               // passengers on board at the start
  int node[MAXS + 1];
  double lb[MAXS], ub[MAXS];
  int8_t pod[MAXS], prev[MAXS];
  uint8_t group[MAXS];    // request group of each stop (for the pairwise pre-checks)
  unsigned dropmask[MAXS];// bit j set in dropmask[i]: stop j is the drop-off whose pick-up is stop i
  double tt[(MAXS + 1) * (MAXS + 1)];   // travel times among the trip's vertices (matrix entries, exact)
};

__device__ __forceinline__ double trip_tt(const Trip& tr, int a, int b) {
  return tr.tt[a * (MAXS + 1) + b];
}

// Depth-first enumeration over the stops in `subset` (bit mask).  Mirrors
// minimalDelayPath + computeRoutingCost (lib/optimUtils.py:432-557):
//   arrival_k = t_start + (((0 + seg_0) + seg_1) + ... + seg_k)        (:536)
//   feasible  = pax <= capacity at every stop (:488-492) and lb <= arrival <= ub (:540-545)
//   cost      = sum over drop-offs of (arrival - lb)                    (:549-550)
// Orderings are visited in lexicographic stop-index order and only a strictly
// smaller cost replaces the incumbent (:447), i.e. the canonical tie rule.
// first_only: return as soon as one feasible complete ordering is found.
// The enumeration is exact and therefore exponential in the number of stops (as in the reference); a
// budget on visited search nodes turns a hopeless instance into an error instead of a kernel that runs
// for hours (*exhausted is set; the host reports it).
constexpr long long kSearchBudget = 20000000LL;   // ~20x the largest instance of a K=4 fleet (10 stops: 113 400 orderings)

//
// forced_first >= 0 restricts the first position to that stop (the lanes of a task split the search by first
// stop); shared_best (may be null) is the task's incumbent cost as unsigned-long-long bits, shared by its lanes:
// a partial ordering is dropped only when it is STRICTLY worse than the shared incumbent, so an equal-cost
// ordering with a lexicographically smaller first stop is still found by its own lane.
//
// The loop is flat -- one candidate stop (or one backtrack) per iteration, the next admissible stop found with
// bit masks (`unlocked` = stops whose pick-up precedes them or is outside the subset) -- so the lanes of a warp,
// each on its own search, stay converged at the loop head.  The ordering in progress is packed 4 bits per depth.
__device__ double trip_dfs(const Trip& tr, unsigned subset, double cap, int pax0, bool first_only,
                           uint8_t* best_order, bool* exhausted, int forced_first = -1,
                           unsigned long long* shared_best = nullptr) {
  const int total = __popc(subset);
  long long visited = 0;
  double best = DRS_INF_ROUTE_COST;
  double psum[MAXS + 1], cost[MAXS + 1];
  unsigned long long ord = 0;
  unsigned used = 0, unlocked = 0;
  for (int i = 0; i < tr.ns; ++i) {
    const int pv = tr.prev[i];
    if (pv < 0 || !((subset >> pv) & 1u)) unlocked |= 1u << i;
  }
  const unsigned first_mask = forced_first >= 0 ? (1u << forced_first) : 0xffffffffu;
  int d = 0, pax = pax0, start = 0;
  psum[0] = 0.0; cost[0] = 0.0;
  while (true) {
    if (++visited > kSearchBudget) { *exhausted = true; break; }
    unsigned candm = subset & ~used & unlocked & (0xffffffffu << start);
    if (d == 0) candm &= first_mask;
    if (!candm) {   // no stop left at this depth: backtrack
      if (d == 0) break;
      --d;
      const int j = (int)((ord >> (4 * d)) & 15ull);
      used &= ~(1u << j);
      unlocked &= ~tr.dropmask[j];
      if (d > 0 || tr.has_start) pax -= tr.pod[j];
      start = j + 1;
      continue;
    }
    const int i = __ffs(candm) - 1;   // orderings in lexicographic stop-index order
    start = i + 1;
    double np, nc;
    int npax;
    if (d == 0 && !tr.has_start) {
      // no vehicle: the first stop is the start; never capacity-counted nor window-checked
      // (lib/optimUtils.py:475-488,540 iterate over orderedLocs[1:])
      np = 0.0; nc = 0.0; npax = pax;
    } else {
      npax = pax + tr.pod[i];
      if ((double)npax > cap) continue;
      const int cur = d == 0 ? 0 : (int)((ord >> (4 * (d - 1))) & 15ull) + 1;
      np = psum[d] + trip_tt(tr, cur, i + 1);
      const double arr = tr.t_start + np;
      if (!(arr >= tr.lb[i] && arr <= tr.ub[i])) continue;
      nc = cost[d];
      if (tr.pod[i] < 0) nc += arr - tr.lb[i];
      if (!(nc < best)) continue;   // cost is non-decreasing along an ordering: exact pruning
      if (shared_best && nc > __longlong_as_double((long long)*((volatile unsigned long long*)shared_best))) continue;
    }
    ord = (ord & ~(15ull << (4 * d))) | ((unsigned long long)i << (4 * d));
    if (d + 1 == total) {   // complete ordering, strictly cheaper than the incumbent
      best = nc;
      if (shared_best) atomicMin(shared_best, (unsigned long long)__double_as_longlong(nc));   // costs are >= 0
      if (best_order)
        for (int k = 0; k < total; ++k) best_order[k] = (uint8_t)((ord >> (4 * k)) & 15ull);
      if (first_only) return best;
      continue;
    }
    used |= 1u << i;
    unlocked |= tr.dropmask[i];
    pax = npax;
    ++d;
    psum[d] = np; cost[d] = nc;
    start = 0;
  }
  return best;
}

struct TaskList {
  int32_t* veh;     // vehicle index or -1
  int32_t* nreq;
  int32_t* req;     // MAXR per task
};

constexpr int GROUP = 8;                 // lanes cooperating on one task
constexpr int TASKS_PER_BLOCK = 8;       // 64 threads per block

// Stop records of a task (lane 0 of the group) -- createLocations order (lib/optimUtils.py:84-110).
__device__ bool fill_trip(Trip& tr, const TickDev& tk, double T, int v, int nreq, const int32_t* reqs, int* n_groups) {
  int ns = 0;
  for (int i = 0; i < MAXS; ++i) tr.dropmask[i] = 0;
  for (int k = 0; k < nreq; ++k) {   // new requests first (:86-92)
    const int r = reqs[k];
    tr.node[1 + ns] = tk.r_o[r]; tr.lb[ns] = tk.r_lbp[r]; tr.ub[ns] = tk.r_ubp[r]; tr.pod[ns] = 1;
    tr.prev[ns] = -1; tr.group[ns] = k; ++ns;
    tr.node[1 + ns] = tk.r_d[r]; tr.lb[ns] = tk.r_lbd[r]; tr.ub[ns] = tk.r_ubd[r]; tr.pod[ns] = -1;
    tr.prev[ns] = (int8_t)(ns - 1); tr.group[ns] = k; tr.dropmask[ns - 1] = 1u << ns; ++ns;
  }
  int groups = nreq;
  if (v >= 0) {
    const int s0 = tk.v_off[v], s1 = tk.v_off[v + 1];
    if (ns + (s1 - s0) > MAXS) return false;
    const int base = ns;
    for (int s = s0; s < s1; ++s) {
      tr.node[1 + ns] = tk.s_node[s]; tr.lb[ns] = tk.s_lb[s]; tr.ub[ns] = tk.s_ub[s]; tr.pod[ns] = tk.s_pod[s];
      const int pv = tk.s_prev[s];
      tr.prev[ns] = (int8_t)(pv < 0 ? -1 : base + pv);
      tr.group[ns] = (uint8_t)(pv < 0 ? groups++ : tr.group[base + pv]);
      if (pv >= 0) tr.dropmask[base + pv] = 1u << ns;
      ++ns;
    }
    tr.has_start = 1;
    tr.node[0] = tk.v_node[v];
    tr.t_start = T + tk.v_delay[v];           // lib/optimUtils.py:117
    tr.cap = (double)tk.v_cap[v];
    tr.pax0 = tk.v_pax[v];
  } else {
    tr.has_start = 0;
    tr.node[0] = tr.node[1];
    tr.t_start = T;                            // lib/optimUtils.py:160
    tr.cap = 1e6;
    tr.pax0 = 0;
  }
  tr.ns = ns;
  *n_groups = groups;
  return true;
}

// findOptimalRouting (lib/optimUtils.py:113-160) for one task by a group of GROUP lanes: the trip lives in
// shared memory; the lanes gather the (S+1) x S travel times together, split the pairwise pre-checks by pair and
// the final enumeration by first stop, share the incumbent cost, and lane 0 writes the result.  The winner among
// equal costs is the lane with the smallest first stop = the lexicographically smallest ordering (each lane visits
// its own orderings in lexicographic order and keeps the first minimum).
__device__ void group_solve(Trip& tr, unsigned long long& shared_best, int* s_flag, const TickDev& tk, const DistView& dv,
                            double T, int flags, bool active, int v, int nreq, const int32_t* reqs, int lane, unsigned gmask,
                            double* cost_out, int* ns_out, uint8_t* order_out) {
  // s_flag[0]: trip fits, [1]: groups, [2]: a pairwise check failed, [3]: budget exhausted
  if (lane == 0) {
    int groups = 0;
    const bool ok = active && fill_trip(tr, tk, T, v, nreq, reqs, &groups);
    s_flag[0] = ok ? 1 : 0; s_flag[1] = groups; s_flag[2] = 0; s_flag[3] = 0;
    shared_best = (unsigned long long)__double_as_longlong(DRS_INF_ROUTE_COST);
  }
  __syncwarp(gmask);
  const bool fits = s_flag[0] != 0;
  const int ns = fits ? tr.ns : 0, groups = s_flag[1];
  for (int idx = lane; idx < (ns + 1) * ns; idx += GROUP) {   // travel times among the trip's vertices
    const int a = idx / ns, b = idx % ns + 1;
    tr.tt[a * (MAXS + 1) + b] = (tr.node[a] == tr.node[b]) ? 0.0 : dv.at(tr.node[a], tr.node[b]);
  }
  __syncwarp(gmask);
  bool exhausted = false;
  // Work is dealt to the lanes so that all of them enter trip_dfs at the same point of the program (a lane that
  // called it from inside a divergent branch would run alone while the others wait).
  if (fits && (flags & DRS_EVAL_PAIRWISE_CHECKS) && tr.has_start && groups >= 3) {
    int npairs = 0;
    for (int a = 0; a < groups; ++a)
      for (int b = a + 1; b < groups; ++b)
        if (!(a >= nreq && b >= nreq)) ++npairs;   // both already on the vehicle: skipped (:142-143)
    for (int p0 = 0; p0 < npairs; p0 += GROUP) {
      const int mine = p0 + lane;
      int pa = -1, pb = -1, pidx = 0;
      for (int a = 0; a < groups; ++a)
        for (int b = a + 1; b < groups; ++b) {
          if (a >= nreq && b >= nreq) continue;
          if (pidx++ == mine) { pa = a; pb = b; }
        }
      unsigned sub = 0;
      for (int i = 0; i < ns; ++i)
        if (tr.group[i] == pa || tr.group[i] == pb) sub |= 1u << i;
      // minimalDelayPath(comboLocs, G, startTime, startLoc): default capacity 1e6, startPax 0 (:148-149)
      const double c = sub ? trip_dfs(tr, sub, 1e6, 0, true, nullptr, &exhausted) : 0.0;
      if (!(c < DRS_INF_ROUTE_COST)) s_flag[2] = 1;
    }
  }
  __syncwarp(gmask);
  double best = DRS_INF_ROUTE_COST;
  uint8_t order[MAXS];
  int best_first = MAXS;
  if (fits && !s_flag[2]) {
    const unsigned all = (ns >= 32) ? 0xffffffffu : ((1u << ns) - 1u);
    unsigned firsts = 0;   // stops that may come first
    for (int f = 0; f < ns; ++f)
      if (tr.prev[f] < 0) firsts |= 1u << f;
    const int nf = __popc(firsts);
    for (int k0 = 0; k0 < nf; k0 += GROUP) {
      unsigned m = firsts;
      for (int q = 0; q < k0 + lane; ++q) m &= m - 1;   // drop the k lowest
      const int f = m ? __ffs(m) - 1 : -1;
      uint8_t ord[MAXS];
      const double c = f >= 0 ? trip_dfs(tr, all, tr.cap, tr.pax0, false, ord, &exhausted, f, &shared_best) : DRS_INF_ROUTE_COST;
      if (c < best) { best = c; best_first = f; for (int k = 0; k < ns; ++k) order[k] = ord[k]; }
    }
  }
  if (exhausted) s_flag[3] = 1;
  // reduce over the group: smallest cost, then smallest first stop
  for (int off = GROUP / 2; off > 0; off >>= 1) {
    const double oc = __shfl_down_sync(gmask, best, off, GROUP);
    const int of = __shfl_down_sync(gmask, best_first, off, GROUP);
    int take = (oc < best || (oc == best && of < best_first)) ? 1 : 0;
    // the winner's order travels with it
    for (int k = 0; k < MAXS; k += 4) {
      unsigned w = (unsigned)order[k] | ((unsigned)order[k + 1] << 8) | ((unsigned)order[k + 2] << 16) | ((unsigned)order[k + 3] << 24);
      const unsigned ow = __shfl_down_sync(gmask, w, off, GROUP);
      if (take) { order[k] = ow & 255; order[k + 1] = (ow >> 8) & 255; order[k + 2] = (ow >> 16) & 255; order[k + 3] = (ow >> 24) & 255; }
    }
    if (take) { best = oc; best_first = of; }
  }
  __syncwarp(gmask);
  if (lane == 0 && active) {
    int nso = fits ? ns : -1;                    // -1: too many stops for DRS_MAX_STOPS
    if (s_flag[3]) { nso = -2; best = DRS_INF_ROUTE_COST; }   // -2: search budget exhausted
    *cost_out = fits ? best : DRS_INF_ROUTE_COST;
    *ns_out = nso;
    for (int i = 0; i < MAXS; ++i) order_out[i] = (best < DRS_INF_ROUTE_COST && i < ns) ? order[i] : 0xff;
  }
}

__global__ void __launch_bounds__(GROUP * TASKS_PER_BLOCK) trip_eval_kernel(TickDev tk, DistView dv, double T, int flags, int ntasks,
                                                                            TaskList tl, double* cost_out, uint8_t* order_out,
                                                                            int32_t* nstops_out) {
  __shared__ Trip s_trip[TASKS_PER_BLOCK];
  __shared__ unsigned long long s_best[TASKS_PER_BLOCK];
  __shared__ int s_flag[TASKS_PER_BLOCK][4];
  const int slot = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;
  const int k = blockIdx.x * TASKS_PER_BLOCK + slot;
  const bool active = k < ntasks;
  const unsigned gmask = 0xffu << (((threadIdx.x & 31) / GROUP) * GROUP);
  const int kk = active ? k : 0;
  group_solve(s_trip[slot], s_best[slot], s_flag[slot], tk, dv, T, flags, active, active ? tl.veh[kk] : -1, active ? tl.nreq[kk] : 0,
              tl.req + (int64_t)kk * MAXR, lane, gmask, cost_out + kk, nstops_out + kk, order_out + (int64_t)kk * MAXS);
}

// status per pair + compact survivor list (warp-aggregated append)
__global__ void rv_filter_kernel(TickDev tk, DistView dv, double T, int8_t* status, int32_t* surv_veh,
                                 int32_t* surv_req, int* surv_count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)tk.n_veh * tk.n_req;
  bool survive = false;
  int v = 0, r = 0;
  if (idx < total) {
    // vehicles fastest (sorted by plan length): on an undirected map tt(veh -> origin) is read from the
    // ORIGIN's matrix row, which all warps working on this request share (L2 resident)
    r = (int)(idx / tk.n_veh);
    v = tk.v_perm[(int)(idx % tk.n_veh)];
    const int vn = tk.v_node[v], o = tk.r_o[r];
    const double tvo = (vn == o) ? 0.0 : (tk.directed ? dv.at(vn, o) : dv.at(o, vn));
    int st;
    if (T + tvo > tk.r_ubp[r]) {               // lib/Optimization.py:183-186
      st = DRS_RV_SKIPPED;
    } else {
      st = -1;
      if (tk.v_off[v + 1] > tk.v_off[v]) {     // busy vehicle: empty-dummy check (:193-202)
        const double t0 = T + tk.v_delay[v];
        bool ok = 1 <= tk.v_cap[v];
        const double p1 = 0.0 + tvo;
        const double a1 = t0 + p1;
        ok = ok && a1 >= tk.r_lbp[r] && a1 <= tk.r_ubp[r];
        const int dn = tk.r_d[r];
        const double p2 = p1 + ((o == dn) ? 0.0 : dv.at(o, dn));
        const double a2 = t0 + p2;
        ok = ok && a2 >= tk.r_lbd[r] && a2 <= tk.r_ubd[r];
        ok = ok && ((a2 - tk.r_lbd[r]) < DRS_INF_ROUTE_COST);
        if (!ok) st = DRS_RV_SAVED;
      }
      if (st < 0) survive = true;
    }
    if (!survive) status[(int64_t)v * tk.n_req + r] = (int8_t)st;
  }
  const unsigned m = __ballot_sync(0xffffffffu, survive);
  if (m) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(surv_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (survive) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      surv_veh[pos] = v;
      surv_req[pos] = r;
    }
  }
}

// RV stage, second half: one group of GROUP lanes per surviving pair
__global__ void __launch_bounds__(GROUP * TASKS_PER_BLOCK) rv_eval_kernel(TickDev tk, DistView dv, double T, int flags, int nsurv,
                                                                          const int32_t* surv_veh, const int32_t* surv_req,
                                                                          int8_t* status, double* cost_out, uint8_t* order_out,
                                                                          int32_t* nstops_out) {
  __shared__ Trip s_trip[TASKS_PER_BLOCK];
  __shared__ unsigned long long s_best[TASKS_PER_BLOCK];
  __shared__ int s_flag[TASKS_PER_BLOCK][4];
  __shared__ int32_t s_req[TASKS_PER_BLOCK];
  const int slot = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;
  const int k = blockIdx.x * TASKS_PER_BLOCK + slot;
  const bool active = k < nsurv;
  const unsigned gmask = 0xffu << (((threadIdx.x & 31) / GROUP) * GROUP);
  const int kk = active ? k : 0;
  const int v = active ? surv_veh[kk] : -1, r = active ? surv_req[kk] : 0;
  if (lane == 0) s_req[slot] = r;
  __syncwarp(gmask);
  group_solve(s_trip[slot], s_best[slot], s_flag[slot], tk, dv, T, flags, active, v, active ? 1 : 0, &s_req[slot], lane, gmask,
              cost_out + kk, nstops_out + kk, order_out + (int64_t)kk * MAXS);
  if (lane == 0 && active)
    status[(int64_t)v * tk.n_req + r] = (int8_t)((cost_out[kk] < DRS_INF_ROUTE_COST) ? DRS_RV_FEASIBLE : DRS_RV_INFEASIBLE);
}

__global__ void count_status_kernel(const int8_t* status, int64_t total, unsigned long long* counters) {
  unsigned long long c[4] = {0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = status[i];
    if (s >= 0 && s < 4) c[s]++;
  }
  for (int s = 0; s < 4; ++s) {
    unsigned long long x = c[s];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(&counters[s], x);
  }
}

template <typename T> int upload(T** dptr, const T* h, size_t n, cudaStream_t st) { return drs_upload(dptr, h, n, st); }

}  // namespace

static void tick_free_results(drs_tick* t) {
  void* ptrs[] = {t->d_status, t->d_surv_veh, t->d_surv_req, t->d_surv_count, t->d_cost, t->d_order, t->d_nstops,
                  t->d_counters};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  t->d_status = nullptr; t->d_surv_veh = t->d_surv_req = nullptr; t->d_surv_count = nullptr; t->d_cost = nullptr;
  t->d_order = nullptr; t->d_nstops = nullptr; t->d_counters = nullptr;
  t->surv_capacity = 0; t->rv_pairs = 0; t->rv_nsurv = 0;
}

extern "C" int drs_tick_create(drs_graph* g, drs_tick** out) {
  DRS_REQUIRE(g && out, DRS_ERR_ARG, "drs_tick_create: null argument");
  drs_tick* t = new (std::nothrow) drs_tick();
  DRS_REQUIRE(t, DRS_ERR_NOMEM, "drs_tick_create: out of host memory");
  t->g = g;
  *out = t;
  return DRS_OK;
}

extern "C" int drs_tick_destroy(drs_tick* t) {
  if (!t) return DRS_OK;
  DrsDeviceGuard guard(t->g->device);
  tick_free_results(t);
  drs_tick_free_plans(t);
  TickDev& d = t->dev;
  void* ptrs[] = {d.v_perm, d.v_node, d.v_cap, d.v_pax, d.v_off, d.v_delay, d.s_node, d.s_rid, d.s_pod, d.s_prev, d.s_lb, d.s_ub,
                  d.r_rid, d.r_o, d.r_d, d.r_lbp, d.r_ubp, d.r_lbd, d.r_ubd};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete t;
  return DRS_OK;
}

extern "C" int drs_tick_set_vehicles(drs_tick* t, const drs_vehicle_batch* b) {
  DRS_REQUIRE(t && b, DRS_ERR_ARG, "drs_tick_set_vehicles: null argument");
  DRS_REQUIRE(b->n_veh >= 0, DRS_ERR_ARG, "drs_tick_set_vehicles: negative n_veh");
  const int nv = b->n_veh;
  DRS_REQUIRE(nv == 0 || (b->start_node && b->start_delay && b->capacity && b->onboard && b->stop_off), DRS_ERR_ARG,
              "drs_tick_set_vehicles: null vehicle array");
  const int n = t->g->n;
  int nstops = nv ? b->stop_off[nv] : 0;
  DRS_REQUIRE(nv == 0 || b->stop_off[0] == 0, DRS_ERR_ARG, "drs_tick_set_vehicles: stop_off[0] must be 0");
  for (int v = 0; v < nv; ++v) {
    DRS_REQUIRE(b->start_node[v] >= 0 && b->start_node[v] < n, DRS_ERR_RANGE,
                "drs_tick_set_vehicles: vehicle %d start node %d out of range", v, b->start_node[v]);
    const int cnt = b->stop_off[v + 1] - b->stop_off[v];
    DRS_REQUIRE(cnt >= 0 && cnt <= DRS_MAX_STOPS - 2, DRS_ERR_ARG,
                "drs_tick_set_vehicles: vehicle %d has %d planned stops (max %d)", v, cnt, DRS_MAX_STOPS - 2);
    DRS_REQUIRE(b->onboard[v] <= b->capacity[v], DRS_ERR_ARG,
                "drs_tick_set_vehicles: vehicle %d starts over capacity (lib/optimUtils.py:468)", v);
    for (int s = b->stop_off[v]; s < b->stop_off[v + 1]; ++s) {
      DRS_REQUIRE(b->stop_node[s] >= 0 && b->stop_node[s] < n, DRS_ERR_RANGE,
                  "drs_tick_set_vehicles: stop %d node %d out of range", s, b->stop_node[s]);
      DRS_REQUIRE(b->stop_prev[s] < s - b->stop_off[v], DRS_ERR_ARG,
                  "drs_tick_set_vehicles: stop %d must come after its predecessor", s);
    }
  }
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  TickDev& d = t->dev;
  int rc;
  std::vector<int32_t> perm(std::max(nv, 1));
  for (int v = 0; v < nv; ++v) perm[v] = v;
  // longest plans first: their ordering searches are the long ones and should not form the tail of the launch
  std::stable_sort(perm.begin(), perm.begin() + nv, [&](int x, int y) {
    return (b->stop_off[x + 1] - b->stop_off[x]) > (b->stop_off[y + 1] - b->stop_off[y]);
  });
  if ((rc = upload(&d.v_perm, perm.data(), nv, st))) return rc;
  d.directed = t->g->directed;
  if ((rc = upload(&d.v_node, b->start_node, nv, st)) || (rc = upload(&d.v_delay, b->start_delay, nv, st)) ||
      (rc = upload(&d.v_cap, b->capacity, nv, st)) || (rc = upload(&d.v_pax, b->onboard, nv, st)) ||
      (rc = upload(&d.v_off, b->stop_off, nv ? nv + 1 : 0, st)) || (rc = upload(&d.s_node, b->stop_node, nstops, st)) ||
      (rc = upload(&d.s_rid, b->stop_rid, nstops, st)) || (rc = upload(&d.s_pod, b->stop_pod, nstops, st)) ||
      (rc = upload(&d.s_prev, b->stop_prev, nstops, st)) || (rc = upload(&d.s_lb, b->stop_lb, nstops, st)) ||
      (rc = upload(&d.s_ub, b->stop_ub, nstops, st)))
    return rc;
  DRS_CUDA(cudaStreamSynchronize(st));
  d.n_veh = nv;
  d.n_stops = nstops;
  return DRS_OK;
}

extern "C" int drs_tick_set_requests(drs_tick* t, const drs_request_batch* b) {
  DRS_REQUIRE(t && b, DRS_ERR_ARG, "drs_tick_set_requests: null argument");
  const int nr = b->n_req;
  DRS_REQUIRE(nr >= 0, DRS_ERR_ARG, "drs_tick_set_requests: negative n_req");
  DRS_REQUIRE(nr == 0 || (b->rid && b->o_node && b->d_node && b->lb_p && b->ub_p && b->lb_d && b->ub_d), DRS_ERR_ARG,
              "drs_tick_set_requests: null request array");
  for (int r = 0; r < nr; ++r)
    DRS_REQUIRE(b->o_node[r] >= 0 && b->o_node[r] < t->g->n && b->d_node[r] >= 0 && b->d_node[r] < t->g->n,
                DRS_ERR_RANGE, "drs_tick_set_requests: request %d node out of range", r);
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  TickDev& d = t->dev;
  int rc;
  if ((rc = upload(&d.r_rid, b->rid, nr, st)) || (rc = upload(&d.r_o, b->o_node, nr, st)) ||
      (rc = upload(&d.r_d, b->d_node, nr, st)) || (rc = upload(&d.r_lbp, b->lb_p, nr, st)) ||
      (rc = upload(&d.r_ubp, b->ub_p, nr, st)) || (rc = upload(&d.r_lbd, b->lb_d, nr, st)) ||
      (rc = upload(&d.r_ubd, b->ub_d, nr, st)))
    return rc;
  DRS_CUDA(cudaStreamSynchronize(st));
  d.n_req = nr;
  return DRS_OK;
}

extern "C" int drs_tick_last_ms(const drs_tick* t, double* ms) {
  DRS_REQUIRE(t && ms, DRS_ERR_ARG, "drs_tick_last_ms: null argument");
  for (int i = 0; i < 4; ++i) ms[i] = t->last_ms[i];
  return DRS_OK;
}

static int tick_ready(const drs_tick* t, const char* who) {
  DRS_REQUIRE(t, DRS_ERR_ARG, "%s: null tick", who);
  DRS_REQUIRE(t->g->apsp_valid, DRS_ERR_STATE, "%s: travel-time matrix not built (or invalidated by a weight update)", who);
  return DRS_OK;
}

extern "C" int drs_tick_rv_eval(drs_tick* t, double T, int32_t flags, int32_t* n_edges, int64_t* counters_host) {
  int rc = tick_ready(t, "drs_tick_rv_eval");
  if (rc) return rc;
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  const TickDev& d = t->dev;
  const int64_t total = (int64_t)d.n_veh * d.n_req;
  DRS_REQUIRE(total < (int64_t)1 << 31, DRS_ERR_ARG, "drs_tick_rv_eval: %lld pairs exceed 2^31", (long long)total);
  if (n_edges) *n_edges = 0;
  if (counters_host) std::memset(counters_host, 0, 4 * sizeof(int64_t));
  t->rv_pairs = total;
  t->rv_nsurv = 0;
  if (total == 0) return DRS_OK;
  if (t->surv_capacity < total) {
    tick_free_results(t);
    t->rv_pairs = total;
    DRS_CUDA(cudaMalloc((void**)&t->d_status, total));
    DRS_CUDA(cudaMalloc((void**)&t->d_surv_veh, sizeof(int32_t) * total));
    DRS_CUDA(cudaMalloc((void**)&t->d_surv_req, sizeof(int32_t) * total));
    DRS_CUDA(cudaMalloc((void**)&t->d_surv_count, sizeof(int)));
    DRS_CUDA(cudaMalloc((void**)&t->d_counters, 4 * sizeof(unsigned long long)));
    t->surv_capacity = total;
  }
  DistView dv{t->g->d_dist, t->g->ld, t->g->mode};
  DRS_CUDA(cudaMemsetAsync(t->d_surv_count, 0, sizeof(int), st));
  DRS_CUDA(cudaMemsetAsync(t->d_counters, 0, 4 * sizeof(unsigned long long), st));
  const int threads = 256;
  DrsTimer tm_filter(st), tm_eval(st);
  tm_filter.start();
  rv_filter_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(d, dv, T, t->d_status, t->d_surv_veh,
                                                                                    t->d_surv_req, t->d_surv_count);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  tm_filter.stop();
  int nsurv = 0;
  DRS_CUDA(cudaMemcpyAsync(&nsurv, t->d_surv_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  t->rv_nsurv = nsurv;
  if (t->d_cost) { cudaFree(t->d_cost); cudaFree(t->d_order); cudaFree(t->d_nstops); t->d_cost = nullptr; t->d_order = nullptr; t->d_nstops = nullptr; }
  DRS_CUDA(cudaMalloc((void**)&t->d_cost, sizeof(double) * std::max(nsurv, 1)));
  DRS_CUDA(cudaMalloc((void**)&t->d_order, (size_t)MAXS * std::max(nsurv, 1)));
  DRS_CUDA(cudaMalloc((void**)&t->d_nstops, sizeof(int32_t) * std::max(nsurv, 1)));
  tm_eval.start();
  if (nsurv > 0) {
    rv_eval_kernel<<<(nsurv + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, GROUP * TASKS_PER_BLOCK, 0, st>>>(d, dv, T, flags, nsurv, t->d_surv_veh, t->d_surv_req, t->d_status,
                                                     t->d_cost, t->d_order, t->d_nstops);
    drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  }
  tm_eval.stop();
  count_status_kernel<<<148 * 2, 256, 0, st>>>(t->d_status, total, t->d_counters);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  unsigned long long c[4];
  DRS_CUDA(cudaMemcpyAsync(c, t->d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  t->last_ms[0] = tm_filter.ms();
  t->last_ms[1] = tm_eval.ms();
  if (counters_host)
    for (int s = 0; s < 4; ++s) counters_host[s] = (int64_t)c[s];
  if (n_edges) *n_edges = (int32_t)c[DRS_RV_FEASIBLE];
  return DRS_OK;
}

extern "C" int drs_tick_rv_fetch(drs_tick* t, int32_t capacity_edges, int32_t* veh_idx, int32_t* req_idx, double* cost,
                                 uint8_t* order, int32_t* n_stops) {
  DRS_REQUIRE(t, DRS_ERR_ARG, "drs_tick_rv_fetch: null tick");
  const int ns = t->rv_nsurv;
  if (ns == 0) return DRS_OK;
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  std::vector<int32_t> v(ns), r(ns), nst(ns);
  std::vector<double> c(ns);
  std::vector<uint8_t> o((size_t)ns * MAXS);
  DRS_CUDA(cudaMemcpyAsync(v.data(), t->d_surv_veh, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(r.data(), t->d_surv_req, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(c.data(), t->d_cost, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(nst.data(), t->d_nstops, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(o.data(), t->d_order, (size_t)ns * MAXS, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < ns; ++k) {
    DRS_REQUIRE(nst[k] != -2, DRS_ERR_CAPACITY,
                "drs_tick_rv_fetch: exact ordering search for vehicle %d + request %d exceeded its budget (too many stops to enumerate)", v[k], r[k]);
    DRS_REQUIRE(nst[k] >= 0, DRS_ERR_ARG, "drs_tick_rv_fetch: vehicle %d + request %d exceed DRS_MAX_STOPS", v[k], r[k]);
  }
  std::vector<int> idx;
  idx.reserve(ns);
  for (int k = 0; k < ns; ++k)
    if (c[k] < DRS_INF_ROUTE_COST) idx.push_back(k);
  DRS_REQUIRE((int)idx.size() <= capacity_edges, DRS_ERR_CAPACITY, "drs_tick_rv_fetch: %d edges, buffer holds %d",
              (int)idx.size(), capacity_edges);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return r[a] != r[b] ? r[a] < r[b] : v[a] < v[b]; });
  for (size_t e = 0; e < idx.size(); ++e) {
    const int k = idx[e];
    if (veh_idx) veh_idx[e] = v[k];
    if (req_idx) req_idx[e] = r[k];
    if (cost) cost[e] = c[k];
    if (n_stops) n_stops[e] = nst[k];
    if (order) std::memcpy(order + e * MAXS, o.data() + (size_t)k * MAXS, MAXS);
  }
  return DRS_OK;
}

extern "C" int drs_tick_trip_eval(drs_tick* t, double T, int32_t flags, int32_t n_tasks, const int32_t* task_veh,
                                  const int32_t* task_nreq, const int32_t* task_req, double* cost_host,
                                  uint8_t* order_host, int32_t* n_stops_host) {
  int rc = tick_ready(t, "drs_tick_trip_eval");
  if (rc) return rc;
  DRS_REQUIRE(n_tasks >= 0 && (n_tasks == 0 || (task_veh && task_nreq && task_req && cost_host)), DRS_ERR_ARG,
              "drs_tick_trip_eval: bad argument");
  if (n_tasks == 0) return DRS_OK;
  const TickDev& d = t->dev;
  for (int k = 0; k < n_tasks; ++k) {
    DRS_REQUIRE(task_veh[k] >= -1 && task_veh[k] < d.n_veh, DRS_ERR_RANGE, "drs_tick_trip_eval: task %d vehicle %d out of range",
                k, task_veh[k]);
    DRS_REQUIRE(task_nreq[k] >= 1 && task_nreq[k] <= MAXR, DRS_ERR_ARG, "drs_tick_trip_eval: task %d has %d requests (1..%d)", k,
                task_nreq[k], MAXR);
    for (int j = 0; j < task_nreq[k]; ++j)
      DRS_REQUIRE(task_req[k * MAXR + j] >= 0 && task_req[k * MAXR + j] < d.n_req, DRS_ERR_RANGE,
                  "drs_tick_trip_eval: task %d request index out of range", k);
  }
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  int32_t *dv_ = nullptr, *dn = nullptr, *dr = nullptr, *dns = nullptr;
  double* dc = nullptr;
  uint8_t* dord = nullptr;
  auto cleanup = [&]() { cudaFree(dv_); cudaFree(dn); cudaFree(dr); cudaFree(dns); cudaFree(dc); cudaFree(dord); };
  if ((rc = upload(&dv_, task_veh, n_tasks, st)) || (rc = upload(&dn, task_nreq, n_tasks, st)) ||
      (rc = upload(&dr, task_req, (size_t)n_tasks * MAXR, st))) { cleanup(); return rc; }
  if (cudaMalloc((void**)&dc, sizeof(double) * n_tasks) != cudaSuccess ||
      cudaMalloc((void**)&dord, (size_t)n_tasks * MAXS) != cudaSuccess ||
      cudaMalloc((void**)&dns, sizeof(int32_t) * n_tasks) != cudaSuccess) {
    cleanup();
    drs_set_error("drs_tick_trip_eval: cudaMalloc failed");
    return DRS_ERR_NOMEM;
  }
  DistView dv{t->g->d_dist, t->g->ld, t->g->mode};
  TaskList tl{dv_, dn, dr};
  trip_eval_kernel<<<(n_tasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, GROUP * TASKS_PER_BLOCK, 0, st>>>(d, dv, T, flags, n_tasks, tl, dc, dord, dns);
  drs_count_launch();
  std::vector<int32_t> nst(n_tasks);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(cost_host, dc, sizeof(double) * n_tasks, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && order_host) e = cudaMemcpyAsync(order_host, dord, (size_t)n_tasks * MAXS, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nst.data(), dns, sizeof(int32_t) * n_tasks, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) {
    drs_set_error("drs_tick_trip_eval: %s", cudaGetErrorString(e));
    return DRS_ERR_CUDA;
  }
  for (int k = 0; k < n_tasks; ++k) {
    DRS_REQUIRE(nst[k] != -2, DRS_ERR_CAPACITY,
                "drs_tick_trip_eval: exact ordering search for task %d exceeded its budget (too many stops to enumerate)", k);
    DRS_REQUIRE(nst[k] >= 0, DRS_ERR_ARG, "drs_tick_trip_eval: task %d exceeds DRS_MAX_STOPS (%d)", k, DRS_MAX_STOPS);
  }
  if (n_stops_host) std::memcpy(n_stops_host, nst.data(), sizeof(int32_t) * n_tasks);
  return DRS_OK;
}
