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
//   rv_eval_kernel     findOptimalRouting of each surviving request inserted into the
//                      vehicle's plan: 8 trips per CTA in shared memory, the searches
//                      (pairwise pre-checks, then exact searches split by the first 1-3
//                      stops) pulled by the 128 lanes from a CTA-wide work list.
//   trip_eval_kernel   the same for explicit trips (RR edges, RTV trips): gathers the
//                      (S+1)^2 travel times among the trip's vertices, runs the pairwise
//                      pre-checks (lib/optimUtils.py:127-156) and a depth-first
//                      enumeration of the precedence-respecting orderings in
//                      lexicographic order with exact feasibility / cost-bound pruning.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "drs_common.cuh"
#include "drs_tick.cuh"

namespace {

constexpr int MAXS = DRS_MAX_STOPS;
constexpr int MAXR = DRS_MAX_TRIP_REQS;

// One evaluation problem ("trip": the stops of a vehicle's plan plus up to four new requests) in shared
// memory.  Vertex 0 of `tt` is the start.
struct Trip {
  double lb[MAXS], ub[MAXS];
  double t_start;         // T + veh["t"]  (or T without a vehicle)
  double snap;            // TickDev::snap
  double cap;             // capacity (1e6 when unconstrained, lib/optimUtils.py:432)
  int tt[(MAXS + 1) * (MAXS + 1)];   // raw matrix entries (float bits or int32) among the trip's vertices
  int node[MAXS + 1];
  unsigned short dropmask[MAXS];     // bit j of dropmask[i]: stop j is the drop-off whose pick-up is stop i
  int8_t pod[MAXS], prev[MAXS];
  uint8_t group[MAXS];    // request group of each stop (for the pairwise pre-checks)
  int ns;                 // stops
  int has_start;          // 1: vertex 0 is the vehicle's next node; 0: route starts at its first stop
  int pax0;               // passengers on board at the start
  int groups, nreq;       // distinct requests; the first nreq groups are the new ones
  int fits;               // 0: more than DRS_MAX_STOPS stops (or an inactive slot)
};

__device__ __forceinline__ double decode_tt(int raw, int mode) {
  if (mode == DRS_MODE_F32) return (double)__int_as_float(raw);
  return raw >= DRS_I32_INF ? __longlong_as_double(0x7ff0000000000000LL) : (double)raw;
}

// Depth-first enumeration over the stops in `subset` (bit mask).  Mirrors
// minimalDelayPath + computeRoutingCost (lib/optimUtils.py:432-557):
//   arrival_k = t_start + (((0 + seg_0) + seg_1) + ... + seg_k)        (:536)
//   feasible  = pax <= capacity at every stop (:488-492) and lb <= arrival <= ub (:540-545)
//   cost      = sum over drop-offs of (arrival - lb)                    (:549-550)
// Orderings are visited in lexicographic stop-index order and only a strictly
// smaller cost replaces the incumbent (:447), i.e. the canonical tie rule.
// The enumeration is exact and therefore exponential in the number of stops (as in the reference); a
// budget on visited search nodes turns a hopeless instance into an error instead of a kernel that runs
// for hours (`exhausted` is set; the host reports it).
constexpr int kSearchBudget = 20000000;   // steps per trip and work list; ~20x the largest instance of a K=4 fleet (10 stops: 113 400 orderings)
constexpr int kChargeEvery = 256;         // a search charges its trip's step counter every so many steps

// The search is a resumable state machine: dfs_step tries ONE candidate stop (or backtracks once).  Every lane of
// a CTA runs its own searches, pulled from a CTA-wide work list, inside one flat loop -- so lanes whose searches
// differ in length by orders of magnitude stay converged at the loop head and none waits for another.
//   pmask[0..2] restrict the first three positions (the searches of a trip are split by its first stops)
//   shared_best the trip's incumbent cost (fp64 bits; costs are >= 0 so the bit patterns order like the values),
//               shared by its searches: a partial ordering is dropped only when STRICTLY worse, so equal-cost
//               orderings in other sub-trees are still found
//   first_only  stop at the first complete feasible ordering
// The next admissible stop comes from bit masks (`unlocked` = stops whose pick-up precedes them or is outside
// the subset); the ordering in progress is packed 4 bits per position, FIRST position in the most significant
// nibble, so that comparing two packed orderings of a trip as integers compares them lexicographically.
struct Dfs {   // scalars only, so that it lives in registers; the per-depth sums are separate local arrays
  double best, cap;
  unsigned long long ord, best_ord;
  unsigned long long* shared_best;
  int* steps;             // the trip's step counter (shared by its searches), charged every kChargeEvery steps
  const Trip* tr;
  unsigned subset, used, unlocked, pmask0, pmask1, pmask2;
  int total, d, pax, start, visited;
  bool first_only, done, exhausted;
};

struct DfsStack {
  double psum[MAXS + 1], cost[MAXS + 1];   // partial travel-time sum and cost at each depth
};

__device__ __forceinline__ int ord_at(unsigned long long ord, int pos) { return (int)((ord >> (4 * (15 - pos))) & 15ull); }

__device__ __forceinline__ void dfs_begin(Dfs& s, DfsStack& k, const Trip* tr, unsigned subset, double cap, int pax0, bool first_only,
                                          const unsigned* pmask, unsigned long long* shared_best, int* steps) {
  s.steps = steps;
  s.tr = tr; s.subset = subset; s.cap = cap; s.pax = pax0; s.first_only = first_only;
  s.pmask0 = pmask ? pmask[0] : 0xffffffffu; s.pmask1 = pmask ? pmask[1] : 0xffffffffu; s.pmask2 = pmask ? pmask[2] : 0xffffffffu;
  s.shared_best = shared_best;
  s.total = __popc(subset);
  s.best = DRS_INF_ROUTE_COST; s.ord = 0; s.best_ord = 0; s.used = 0; s.d = 0; s.start = 0; s.visited = 0;
  s.done = (s.total == 0); s.exhausted = false;
  unsigned unlocked = 0;
  for (int i = 0; i < tr->ns; ++i) {
    const int pv = tr->prev[i];
    if (pv < 0 || !((subset >> pv) & 1u)) unlocked |= 1u << i;
  }
  s.unlocked = unlocked;
  k.psum[0] = 0.0; k.cost[0] = 0.0;
}

// lb <= arr <= ub, an arrival within the relative slack `snap` of a bound counting as on it (snap = 0: the exact test)
__device__ __forceinline__ bool in_window(double arr, double lb, double ub, double snap) {
  return arr >= lb - snap * fmax(1.0, fabs(lb)) && arr <= ub + snap * fmax(1.0, fabs(ub));
}

__device__ __forceinline__ void dfs_step(Dfs& s, DfsStack& k, int mode) {
  const Trip& tr = *s.tr;
  if ((++s.visited & (kChargeEvery - 1)) == 0 && atomicAdd(s.steps, kChargeEvery) > kSearchBudget) {
    s.exhausted = true; s.done = true;
    return;
  }
  const int d = s.d;
  unsigned candm = s.subset & ~s.used & s.unlocked & (0xffffffffu << s.start);
  if (d < 3) candm &= (d == 0 ? s.pmask0 : d == 1 ? s.pmask1 : s.pmask2);
  if (!candm) {   // no stop left at this depth: backtrack
    if (d == 0) { s.done = true; return; }
    const int j = ord_at(s.ord, d - 1);
    s.d = d - 1;
    s.used &= ~(1u << j);
    s.unlocked &= ~(unsigned)tr.dropmask[j];
    if (d - 1 > 0 || tr.has_start) s.pax -= tr.pod[j];
    s.start = j + 1;
    return;
  }
  const int i = __ffs(candm) - 1;   // orderings in lexicographic stop-index order
  s.start = i + 1;
  double np, nc;
  int npax;
  if (d == 0 && !tr.has_start) {
    // no vehicle: the first stop is the start; never capacity-counted nor window-checked
    // (lib/optimUtils.py:475-488,540 iterate over orderedLocs[1:])
    np = 0.0; nc = 0.0; npax = s.pax;
  } else {
    npax = s.pax + tr.pod[i];
    if ((double)npax > s.cap) return;
    const int cur = d == 0 ? 0 : ord_at(s.ord, d - 1) + 1;
    np = k.psum[d] + decode_tt(tr.tt[cur * (MAXS + 1) + i + 1], mode);
    const double arr = tr.t_start + np;
    if (!in_window(arr, tr.lb[i], tr.ub[i], tr.snap)) return;
    nc = k.cost[d];
    if (tr.pod[i] < 0) {
      double late = arr - tr.lb[i];
      if (tr.snap > 0.0 && fabs(late) <= tr.snap * fmax(1.0, fabs(tr.lb[i]))) late = 0.0;   // on the bound: the reference's delay is exactly 0
      nc += late;
    }
    if (!(nc < s.best)) return;   // cost is non-decreasing along an ordering: exact pruning
    if (s.shared_best && nc > __longlong_as_double((long long)*((volatile unsigned long long*)s.shared_best))) return;
  }
  const int sh4 = 4 * (15 - d);
  s.ord = (s.ord & ~(15ull << sh4)) | ((unsigned long long)i << sh4);
  if (d + 1 == s.total) {   // complete ordering, strictly cheaper than this search's incumbent
    s.best = nc;
    s.best_ord = s.ord;
    if (s.shared_best) atomicMin(s.shared_best, (unsigned long long)__double_as_longlong(nc));
    if (s.first_only) s.done = true;
    return;
  }
  s.used |= 1u << i;
  s.unlocked |= (unsigned)tr.dropmask[i];
  s.pax = npax;
  s.d = d + 1;
  k.psum[d + 1] = np; k.cost[d + 1] = nc;
  s.start = 0;
}

struct TaskList {
  const int32_t* veh;     // vehicle index or -1
  const int32_t* nreq;    // null: one request per task
  const int32_t* req;     // MAXR per task (stride 1 when nreq is null)
};

// Stop records of a task -- createLocations order (lib/optimUtils.py:84-110): the new requests first (:86-92),
// then the vehicle's planned stops.  The `lanes` lanes of the trip load one stop each (the loads of a serial loop
// would each wait for the previous one's store); number_groups then numbers the request groups in stop order.
__device__ void fill_trip(Trip& tr, const TickDev& tk, double T, int v, int nreq, const int32_t* reqs, int sub, int lanes) {
  const int s0 = v >= 0 ? tk.v_off[v] : 0, s1 = v >= 0 ? tk.v_off[v + 1] : 0;
  const int base = 2 * nreq, ns = base + (s1 - s0);
  const bool fits = ns <= MAXS;
  if (sub == 0) {
    tr.nreq = nreq;
    tr.fits = fits ? 1 : 0;
    tr.ns = fits ? ns : 0;
    if (v >= 0) {
      tr.has_start = 1;
      tr.node[0] = tk.v_node[v];
      tr.t_start = T + tk.v_delay[v];           // lib/optimUtils.py:117
      tr.snap = tk.snap;
      tr.cap = (double)tk.v_cap[v];
      tr.pax0 = tk.v_pax[v];
    } else {
      tr.has_start = 0;
      tr.node[0] = tk.r_o[reqs[0]];             // the first stop is the start (lib/optimUtils.py:160)
      tr.t_start = T;
      tr.snap = tk.snap;
      tr.cap = 1e6;
      tr.pax0 = 0;
    }
  }
  if (!fits) return;
  for (int q = sub; q < nreq; q += lanes) {
    const int r = reqs[q], i = 2 * q;
    tr.node[1 + i] = tk.r_o[r]; tr.lb[i] = tk.r_lbp[r]; tr.ub[i] = tk.r_ubp[r]; tr.pod[i] = 1; tr.prev[i] = -1;
    tr.node[2 + i] = tk.r_d[r]; tr.lb[i + 1] = tk.r_lbd[r]; tr.ub[i + 1] = tk.r_ubd[r]; tr.pod[i + 1] = -1; tr.prev[i + 1] = (int8_t)i;
  }
  for (int q = sub; q < s1 - s0; q += lanes) {
    const int s = s0 + q, i = base + q;
    tr.node[1 + i] = tk.s_node[s]; tr.lb[i] = tk.s_lb[s]; tr.ub[i] = tk.s_ub[s]; tr.pod[i] = tk.s_pod[s];
    const int pv = tk.s_prev[s];
    tr.prev[i] = (int8_t)(pv < 0 ? -1 : base + pv);
  }
}

// Request group of each stop (new request k = group k; planned stops in order of first appearance) and the
// pick-up -> drop-off masks; one lane, shared memory only.
__device__ void number_groups(Trip& tr) {
  int groups = tr.nreq;
  for (int i = 0; i < MAXS; ++i) tr.dropmask[i] = 0;
  for (int i = 0; i < tr.ns; ++i) {
    const int pv = tr.prev[i];
    if (i < 2 * tr.nreq) tr.group[i] = (uint8_t)(i >> 1);
    else tr.group[i] = (uint8_t)(pv < 0 ? groups++ : tr.group[pv]);
    if (pv >= 0) tr.dropmask[pv] = (unsigned short)(1u << i);
  }
  tr.groups = tr.fits ? groups : 0;
}

// Pairs of request groups findOptimalRouting pre-checks (lib/optimUtils.py:127-156): every pair that is not
// already on the vehicle together, when there are at least three groups and a vehicle.
__device__ __forceinline__ int pair_count(const Trip& tr, int flags) {
  if (!tr.fits || !(flags & DRS_EVAL_PAIRWISE_CHECKS) || !tr.has_start || tr.groups < 3) return 0;
  const int g = tr.groups, old = g - tr.nreq;
  return g * (g - 1) / 2 - old * (old - 1) / 2;
}

// A trip's searches are split by its first `depth` stops (1..3).  Search k of a trip stands for the prefix whose
// stops are the digits of k in base ns (most significant first), so k order is lexicographic order; prefixes that
// repeat a stop or place a drop-off before its pick-up are empty searches.  Returns false for those, else the
// one-bit masks of the first `depth` positions (all-ones beyond) and the packed key (like Dfs::ord).
__device__ __forceinline__ int prefix_count(const Trip& tr, int depth) {
  return depth == 1 ? tr.ns : depth == 2 ? tr.ns * tr.ns : tr.ns * tr.ns * tr.ns;
}

__device__ __forceinline__ bool prefix_masks(const Trip& tr, int depth, int k, unsigned* pmask, unsigned long long* key) {
  pmask[0] = pmask[1] = pmask[2] = 0xffffffffu;
  int digit[3];
  for (int p = depth - 1; p >= 0; --p) { digit[p] = k % tr.ns; k /= tr.ns; }
  unsigned used = 0;
  unsigned long long kk = 0;
  for (int p = 0; p < depth; ++p) {
    const int f = digit[p], pv = tr.prev[f];
    if (((used >> f) & 1u) || (pv >= 0 && !((used >> pv) & 1u))) return false;
    used |= 1u << f;
    pmask[p] = 1u << f;
    kk |= (unsigned long long)f << (4 * (15 - p));
  }
  *key = kk;
  return true;
}

// findOptimalRouting (lib/optimUtils.py:113-160) for TASKS trips per CTA of THREADS lanes.
//   1. the trips are laid out in shared memory (the lanes of a trip load one stop each) and all lanes gather the
//      travel times;
//   2. work list A = the pairwise pre-checks of all trips (one search each, first feasible ordering wins);
//   3. work list B = for the trips that passed, one exact search per admissible prefix of 1, 2 or 3 stops (longer
//      trips are split deeper: their sub-trees are the long tail of the launch), sharing the trip's incumbent:
//      the minimum cost;
//   4. work list C = the same searches again with the minimum as the bound, each stopping at its first complete
//      ordering (which attains the minimum, and is the lexicographically smallest that does in its sub-tree);
//      a 64-bit atomicMin over the packed orderings leaves the lexicographically smallest optimal ordering of the
//      trip = the canonical tie rule.  Searches whose prefix is already beaten are skipped.
// Lanes pull searches from the work lists with a shared-memory counter.
template <int TASKS, int THREADS>
struct EvalShared {
  Trip trip[TASKS];
  unsigned long long best[TASKS];      // minimum cost (fp64 bits)
  unsigned long long win_ord[TASKS];   // smallest packed ordering attaining it
  int off[TASKS + 1];
  int count[TASKS], depth[TASKS];
  int fail[TASKS], exhausted[TASKS], steps[TASKS];
  int next;
};

constexpr unsigned long long kNoOrder = ~0ull;

template <int TASKS, int THREADS>
__device__ void run_searches(EvalShared<TASKS, THREADS>& sh, int phase, int mode) {
  const int nitems = sh.off[TASKS];
  Dfs st;
  DfsStack stk;
  st.done = true; st.exhausted = false; st.best = DRS_INF_ROUTE_COST; st.best_ord = 0;
  int item = -1, task = 0;
  while (true) {
    if (st.done) {
      if (item >= 0) {   // retire the finished search
        if (st.exhausted) sh.exhausted[task] = 1;
        else if (phase == 0) { if (!(st.best < DRS_INF_ROUTE_COST)) sh.fail[task] = 1; }
        else if (phase == 2) { if (st.best < DRS_INF_ROUTE_COST) atomicMin(&sh.win_ord[task], st.best_ord); }
      }
      item = atomicAdd(&sh.next, 1);
      if (item >= nitems) break;
      int lo = 0, hi = TASKS;   // task = last t with off[t] <= item
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sh.off[mid] <= item) lo = mid; else hi = mid; }
      task = lo;
      const Trip& tr = sh.trip[task];
      const int k = item - sh.off[task];
      if (sh.exhausted[task]) { item = -1; continue; }   // the trip is over budget: its result is an error anyway
      if (phase == 0) {
        if (sh.fail[task]) { item = -1; continue; }   // already infeasible: skip
        int pa = -1, pb = -1, pidx = 0;
        for (int a = 0; a < tr.groups; ++a)
          for (int b = a + 1; b < tr.groups; ++b) {
            if (a >= tr.nreq && b >= tr.nreq) continue;   // both already on the vehicle (:142-143)
            if (pidx++ == k) { pa = a; pb = b; }
          }
        unsigned sub = 0;
        for (int i = 0; i < tr.ns; ++i)
          if (tr.group[i] == pa || tr.group[i] == pb) sub |= 1u << i;
        // minimalDelayPath(comboLocs, G, startTime, startLoc): default capacity 1e6, startPax 0 (:148-149)
        dfs_begin(st, stk, &tr, sub, 1e6, 0, true, nullptr, nullptr, &sh.steps[task]);
      } else {
        unsigned pm[3];
        unsigned long long key;
        const int dep = sh.depth[task];
        if (!prefix_masks(tr, dep, k, pm, &key)) { item = -1; continue; }
        if (phase == 2) {
          // a smaller ordering is already known: this sub-tree cannot win
          const int shift = 64 - 4 * dep;
          const unsigned long long w = *((volatile unsigned long long*)&sh.win_ord[task]);
          if ((w >> shift) < (key >> shift)) { item = -1; continue; }
        }
        dfs_begin(st, stk, &tr, (1u << tr.ns) - 1u, tr.cap, tr.pax0, phase == 2, pm, &sh.best[task], &sh.steps[task]);
      }
      continue;
    }
    dfs_step(st, stk, mode);
  }
}

__device__ __forceinline__ int split_depth(int ns, int split2_min, int split3_min) {
  const int d = ns >= split3_min ? 3 : ns >= split2_min ? 2 : 1;
  return d < ns ? d : (ns > 0 ? ns : 1);
}

template <int TASKS, int THREADS>
__device__ void eval_block(EvalShared<TASKS, THREADS>& sh, const TickDev& tk, const DistView& dv, double T, int flags, int ntasks,
                           const TaskList& tl, int split, double* cost_out, int32_t* nstops_out, uint8_t* order_out) {
  constexpr int LPT = THREADS / TASKS;   // lanes per trip in the cooperative steps
  const int tid = threadIdx.x;
  const int slot = tid / LPT, sub = tid % LPT;
  const int k = blockIdx.x * TASKS + slot;
  const bool active = k < ntasks;
  Trip& tr = sh.trip[slot];
  if (active) {
    const int nreq = tl.nreq ? tl.nreq[k] : 1;
    fill_trip(tr, tk, T, tl.veh[k], nreq, tl.req + (int64_t)k * (tl.nreq ? MAXR : 1), sub, LPT);
  } else if (sub == 0) {
    tr.fits = 0; tr.ns = 0; tr.nreq = 0; tr.has_start = 0;
  }
  __syncthreads();
  if (sub == 0) {
    number_groups(tr);
    sh.best[slot] = (unsigned long long)__double_as_longlong(DRS_INF_ROUTE_COST);
    sh.win_ord[slot] = kNoOrder;
    sh.fail[slot] = 0; sh.exhausted[slot] = 0;
    sh.count[slot] = pair_count(tr, flags);
  }
  __syncthreads();
  {
    const int ns = tr.ns;
    for (int idx = sub; idx < (ns + 1) * ns; idx += LPT) {   // travel times among the trip's vertices
      const int a = idx / ns, b = idx % ns + 1;
      const int na = tr.node[a], nb = tr.node[b];
      int raw;
      if (na == nb) raw = 0;   // +0.0f and integer 0 share the bit pattern
      else raw = ((const int*)dv.dist)[(int64_t)na * dv.ld + nb];
      tr.tt[a * (MAXS + 1) + b] = raw;
    }
  }
  auto publish = [&]() {   // work list = concatenation of the trips' searches
    if (tid == 0) {
      int acc = 0;
      for (int t = 0; t < TASKS; ++t) { sh.off[t] = acc; acc += sh.count[t]; sh.steps[t] = 0; }
      sh.off[TASKS] = acc;
      sh.next = 0;
    }
    __syncthreads();
  };
  publish();
  run_searches(sh, 0, dv.mode);
  __syncthreads();
  if (sub == 0) {
    const bool go = tr.fits && !sh.fail[slot] && !sh.exhausted[slot];
    sh.depth[slot] = split_depth(tr.ns, split & 0xff, (split >> 8) & 0xff);
    sh.count[slot] = go ? prefix_count(tr, sh.depth[slot]) : 0;
  }
  __syncthreads();
  publish();
  run_searches(sh, 1, dv.mode);
  __syncthreads();
  if (sub == 0 && (sh.exhausted[slot] || !(__longlong_as_double((long long)sh.best[slot]) < DRS_INF_ROUTE_COST))) sh.count[slot] = 0;
  __syncthreads();
  publish();
  run_searches(sh, 2, dv.mode);
  __syncthreads();
  if (sub == 0 && active) {
    double best = __longlong_as_double((long long)sh.best[slot]);
    int nso = tr.fits ? tr.ns : -1;                    // -1: too many stops for DRS_MAX_STOPS
    const unsigned long long ord = sh.win_ord[slot];
    bool have = tr.fits && best < DRS_INF_ROUTE_COST && ord != kNoOrder;
    if (sh.exhausted[slot]) { nso = -2; have = false; }   // -2: search budget exhausted
    if (!have) best = DRS_INF_ROUTE_COST;
    cost_out[k] = best;
    nstops_out[k] = nso;
    for (int i = 0; i < MAXS; ++i) order_out[(int64_t)k * MAXS + i] = (have && i < tr.ns) ? (uint8_t)ord_at(ord, i) : 0xff;
  }
}

template <int TASKS, int THREADS>
__global__ void __launch_bounds__(THREADS) trip_eval_kernel(TickDev tk, DistView dv, double T, int flags, int ntasks, TaskList tl,
                                                            int split, double* cost_out, uint8_t* order_out, int32_t* nstops_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto& sh = *reinterpret_cast<EvalShared<TASKS, THREADS>*>(smem_raw);
  eval_block<TASKS, THREADS>(sh, tk, dv, T, flags, ntasks, tl, split, cost_out, nstops_out, order_out);
}

// RV stage, second half: the surviving (vehicle, request) pairs are trips of one new request
template <int TASKS, int THREADS>
__global__ void __launch_bounds__(THREADS) rv_eval_kernel(TickDev tk, DistView dv, double T, int flags, int nsurv,
                                                          const int32_t* surv_veh, const int32_t* surv_req, int8_t* status,
                                                          int split, double* cost_out, uint8_t* order_out, int32_t* nstops_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto& sh = *reinterpret_cast<EvalShared<TASKS, THREADS>*>(smem_raw);
  const TaskList tl{surv_veh, nullptr, surv_req};
  eval_block<TASKS, THREADS>(sh, tk, dv, T, flags, nsurv, tl, split, cost_out, nstops_out, order_out);
  __syncthreads();
  constexpr int LPT = THREADS / TASKS;
  const int k = blockIdx.x * TASKS + threadIdx.x / LPT;
  if (threadIdx.x % LPT == 0 && k < nsurv)
    status[(int64_t)surv_veh[k] * tk.n_req + surv_req[k]] = (int8_t)((cost_out[k] < DRS_INF_ROUTE_COST) ? DRS_RV_FEASIBLE : DRS_RV_INFEASIBLE);
}

// CTA shape of the evaluation kernels: 8 trips, 128 lanes (measured best of 8..64 trips x 32..256 lanes on the C4
// tick: larger CTAs wait longer at the barriers between the work lists, fewer lanes leave searches queued)
constexpr int EV_TASKS = 8, EV_THREADS = 128;

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
    if (T + tvo > tk.r_ubp[r] + tk.snap * fmax(1.0, fabs(tk.r_ubp[r]))) {   // lib/Optimization.py:183-186
      st = DRS_RV_SKIPPED;
    } else {
      st = -1;
      if (tk.v_off[v + 1] > tk.v_off[v]) {     // busy vehicle: empty-dummy check (:193-202)
        const double t0 = T + tk.v_delay[v];
        bool ok = 1 <= tk.v_cap[v];
        const double p1 = 0.0 + tvo;
        const double a1 = t0 + p1;
        ok = ok && in_window(a1, tk.r_lbp[r], tk.r_ubp[r], tk.snap);
        const int dn = tk.r_d[r];
        const double p2 = p1 + ((o == dn) ? 0.0 : dv.at(o, dn));
        const double a2 = t0 + p2;
        ok = ok && in_window(a2, tk.r_lbd[r], tk.r_ubd[r], tk.snap);
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


// stop counts from which a trip's searches are split by its first two / three stops, packed (split3 << 8) | split2
// (DRS_EVAL_SPLIT2 / DRS_EVAL_SPLIT3 override: tuning knobs; the split never changes the result)
int eval_split() {
  const char* e2 = getenv("DRS_EVAL_SPLIT2");
  const char* e3 = getenv("DRS_EVAL_SPLIT3");
  const int s2 = e2 ? atoi(e2) : 5, s3 = e3 ? atoi(e3) : 8;
  return ((s3 & 0xff) << 8) | (s2 & 0xff);
}

int launch_rv_eval(const TickDev& d, const DistView& dv, double T, int flags, int nsurv, const int32_t* surv_veh,
                   const int32_t* surv_req, int8_t* status, double* cost, uint8_t* order, int32_t* nstops, cudaStream_t st) {
  auto k = rv_eval_kernel<EV_TASKS, EV_THREADS>;
  const size_t smem = sizeof(EvalShared<EV_TASKS, EV_THREADS>);
  k<<<(nsurv + EV_TASKS - 1) / EV_TASKS, EV_THREADS, smem, st>>>(d, dv, T, flags, nsurv, surv_veh, surv_req, status, eval_split(), cost, order, nstops);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

int launch_trip_eval(const TickDev& d, const DistView& dv, double T, int flags, int ntasks, const TaskList& tl, double* cost,
                     uint8_t* order, int32_t* nstops, cudaStream_t st) {
  auto k = trip_eval_kernel<EV_TASKS, EV_THREADS>;
  const size_t smem = sizeof(EvalShared<EV_TASKS, EV_THREADS>);
  k<<<(ntasks + EV_TASKS - 1) / EV_TASKS, EV_THREADS, smem, st>>>(d, dv, T, flags, ntasks, tl, eval_split(), cost, order, nstops);
  drs_count_launch();
  return DRS_OK;
}

static_assert(sizeof(EvalShared<EV_TASKS, EV_THREADS>) <= 48 * 1024, "evaluation CTA must fit the default dynamic shared memory");

}  // namespace

static void tick_free_results(drs_tick* t) {
  t->a_rv.release(); t->a_rvres.release(); t->a_trip.release();
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
  t->a_veh.release(); t->a_req.release();
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
  // longest plans first: their ordering searches are the long ones and should not form the tail of the launch;
  // within a plan length by start vertex, so that neighbouring lanes gather neighbouring columns of a request's
  // matrix row (vertex ids are spatially coherent under the host layer's locality numbering)
  std::stable_sort(perm.begin(), perm.begin() + nv, [&](int x, int y) {
    const int lx = b->stop_off[x + 1] - b->stop_off[x], ly = b->stop_off[y + 1] - b->stop_off[y];
    return lx != ly ? lx > ly : b->start_node[x] < b->start_node[y];
  });
  d.directed = t->g->directed;
  d.snap = t->g->weights_integral ? 0.0 : 1e-5;   // float-weighted map: fp32 matrix entries vs the fp64 sums the windows were built from
  // one host image, one copy (a steady-state tick allocates nothing)
  DrsArenaLayout lay;
  const size_t nvp = std::max(nv, 1), nsp = std::max(nstops, 1);
  const size_t o_perm = lay.add(4 * nvp), o_node = lay.add(4 * nvp), o_delay = lay.add(8 * nvp), o_cap = lay.add(4 * nvp), o_pax = lay.add(4 * nvp),
               o_off = lay.add(4 * (nvp + 1)), o_snode = lay.add(4 * nsp), o_srid = lay.add(4 * nsp), o_spod = lay.add(nsp), o_sprev = lay.add(nsp),
               o_slb = lay.add(8 * nsp), o_sub = lay.add(8 * nsp);
  if ((rc = t->a_veh.reserve(lay.total))) return rc;
  std::vector<unsigned char>& img = t->a_veh.host;
  img.assign(lay.total, 0);
  auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes) std::memcpy(img.data() + o, src, bytes); };
  put(o_perm, perm.data(), 4 * (size_t)nv); put(o_node, b->start_node, 4 * (size_t)nv); put(o_delay, b->start_delay, 8 * (size_t)nv);
  put(o_cap, b->capacity, 4 * (size_t)nv); put(o_pax, b->onboard, 4 * (size_t)nv); put(o_off, b->stop_off, nv ? 4 * (size_t)(nv + 1) : 0);
  put(o_snode, b->stop_node, 4 * (size_t)nstops); put(o_srid, b->stop_rid, 4 * (size_t)nstops); put(o_spod, b->stop_pod, (size_t)nstops);
  put(o_sprev, b->stop_prev, (size_t)nstops); put(o_slb, b->stop_lb, 8 * (size_t)nstops); put(o_sub, b->stop_ub, 8 * (size_t)nstops);
  DRS_CUDA(cudaMemcpyAsync(t->a_veh.ptr, img.data(), lay.total, cudaMemcpyHostToDevice, st));
  unsigned char* vb = (unsigned char*)t->a_veh.ptr;
  d.v_perm = (int32_t*)(vb + o_perm); d.v_node = (int32_t*)(vb + o_node); d.v_delay = (double*)(vb + o_delay); d.v_cap = (int32_t*)(vb + o_cap);
  d.v_pax = (int32_t*)(vb + o_pax); d.v_off = (int32_t*)(vb + o_off); d.s_node = (int32_t*)(vb + o_snode); d.s_rid = (int32_t*)(vb + o_srid);
  d.s_pod = (int8_t*)(vb + o_spod); d.s_prev = (int8_t*)(vb + o_sprev); d.s_lb = (double*)(vb + o_slb); d.s_ub = (double*)(vb + o_sub);
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
  {
    DrsArenaLayout lay;
    const size_t nrp = std::max(nr, 1);
    const size_t o_rid = lay.add(4 * nrp), o_o = lay.add(4 * nrp), o_d = lay.add(4 * nrp), o_lbp = lay.add(8 * nrp), o_ubp = lay.add(8 * nrp),
                 o_lbd = lay.add(8 * nrp), o_ubd = lay.add(8 * nrp);
    if ((rc = t->a_req.reserve(lay.total))) return rc;
    std::vector<unsigned char>& img = t->a_req.host;
    img.assign(lay.total, 0);
    auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes) std::memcpy(img.data() + o, src, bytes); };
    put(o_rid, b->rid, 4 * (size_t)nr); put(o_o, b->o_node, 4 * (size_t)nr); put(o_d, b->d_node, 4 * (size_t)nr); put(o_lbp, b->lb_p, 8 * (size_t)nr);
    put(o_ubp, b->ub_p, 8 * (size_t)nr); put(o_lbd, b->lb_d, 8 * (size_t)nr); put(o_ubd, b->ub_d, 8 * (size_t)nr);
    DRS_CUDA(cudaMemcpyAsync(t->a_req.ptr, img.data(), lay.total, cudaMemcpyHostToDevice, st));
    unsigned char* rb = (unsigned char*)t->a_req.ptr;
    d.r_rid = (int32_t*)(rb + o_rid); d.r_o = (int32_t*)(rb + o_o); d.r_d = (int32_t*)(rb + o_d); d.r_lbp = (double*)(rb + o_lbp);
    d.r_ubp = (double*)(rb + o_ubp); d.r_lbd = (double*)(rb + o_lbd); d.r_ubd = (double*)(rb + o_ubd);
  }
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
  {
    DrsArenaLayout lay;
    const size_t o_st = lay.add((size_t)total), o_sv = lay.add(sizeof(int32_t) * (size_t)total), o_sr = lay.add(sizeof(int32_t) * (size_t)total),
                 o_cnt = lay.add(sizeof(int)), o_ctr = lay.add(4 * sizeof(unsigned long long));
    if ((rc = t->a_rv.reserve(lay.total))) return rc;
    unsigned char* rb = (unsigned char*)t->a_rv.ptr;
    t->d_status = (int8_t*)(rb + o_st); t->d_surv_veh = (int32_t*)(rb + o_sv); t->d_surv_req = (int32_t*)(rb + o_sr);
    t->d_surv_count = (int*)(rb + o_cnt); t->d_counters = (unsigned long long*)(rb + o_ctr);
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
  {
    DrsArenaLayout lay;
    const size_t ns1 = (size_t)std::max(nsurv, 1);
    const size_t o_c = lay.add(sizeof(double) * ns1), o_o = lay.add((size_t)MAXS * ns1), o_n = lay.add(sizeof(int32_t) * ns1);
    if ((rc = t->a_rvres.reserve(lay.total))) return rc;
    unsigned char* rb = (unsigned char*)t->a_rvres.ptr;
    t->d_cost = (double*)(rb + o_c); t->d_order = (uint8_t*)(rb + o_o); t->d_nstops = (int32_t*)(rb + o_n);
  }
  tm_eval.start();
  if (nsurv > 0) {
    rc = launch_rv_eval(d, dv, T, flags, nsurv, t->d_surv_veh, t->d_surv_req, t->d_status, t->d_cost, t->d_order, t->d_nstops, st);
    if (rc) return rc;
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
  DrsArenaLayout lay;
  const size_t nt = (size_t)n_tasks;
  const size_t o_v = lay.add(4 * nt), o_n = lay.add(4 * nt), o_r = lay.add(4 * nt * MAXR);
  const size_t in_bytes = lay.total;
  const size_t o_c = lay.add(8 * nt), o_ord = lay.add(nt * MAXS), o_ns = lay.add(4 * nt);
  if ((rc = t->a_trip.reserve(lay.total))) return rc;
  std::vector<unsigned char>& img = t->a_trip.host;
  img.assign(in_bytes, 0);
  std::memcpy(img.data() + o_v, task_veh, 4 * nt); std::memcpy(img.data() + o_n, task_nreq, 4 * nt);
  std::memcpy(img.data() + o_r, task_req, 4 * nt * MAXR);
  DRS_CUDA(cudaMemcpyAsync(t->a_trip.ptr, img.data(), in_bytes, cudaMemcpyHostToDevice, st));
  unsigned char* tb = (unsigned char*)t->a_trip.ptr;
  int32_t *dv_ = (int32_t*)(tb + o_v), *dn = (int32_t*)(tb + o_n), *dr = (int32_t*)(tb + o_r), *dns = (int32_t*)(tb + o_ns);
  double* dc = (double*)(tb + o_c);
  uint8_t* dord = (uint8_t*)(tb + o_ord);
  auto cleanup = [&]() {};
  DistView dv{t->g->d_dist, t->g->ld, t->g->mode};
  TaskList tl{dv_, dn, dr};
  rc = launch_trip_eval(d, dv, T, flags, n_tasks, tl, dc, dord, dns, st);
  if (rc) { cleanup(); return rc; }
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
