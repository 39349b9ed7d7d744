// (vehicle, pickup slot, drop-off slot) insertion evaluation, argmin and commit.
//
// Replaces Model.insertion_heuristics / insert_heuristics / test_constraints_get_cost
// (lib/Agents.py:1429-1572) and, for the sequential queue, what Veh.build_route leaves behind for
// the next scan (route, veh.c, Req.pwt: lib/Agents.py:148-297).  Every candidate route is walked in
// fp64 with the reference's operation order, so costs are bit-identical on integer-weighted maps;
// travel times are gathered from the resident matrix.
//
// Plans live slot-major on the device (element [k][vs] = plan stop k of the vehicle at sorted
// position vs; vehicles sorted by plan length, then start vertex), with room for INS_MAXL stops per
// vehicle so that a winner's plan can be patched in place.
//
//   plan_prep_kernel     per vehicle: plan segment times, the walk of the unmodified plan (the common
//       prefix of every candidate) and a conservative fp32 bound per pickup position for the reach test.
//   insertion_reach_kernel<STAGED>   one persistent CTA per SM walks the new requests; for each it stages
//       the request's ORIGIN ROW of the travel-time matrix (4*ld bytes, 200 KB on the 50k-node map) in
//       shared memory with bulk asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier)
//       and tests every vehicle: can ANY pickup position be reached before the latest pickup time?  The
//       (L+1) travel times per pair are shared-memory gathers; the plans stream in coalesced.  Survivors
//       (~1 % of the pairs) go to a bit mask per request and a compact pair list.  Directed maps and rows
//       that do not fit use the same kernel without staging (gathers from L2).
//   insertion_eval_kernel   one thread per surviving pair: every (i, j) candidate WITHOUT the incumbent
//       bound; the pair's smallest successful total cost is written out (+inf if none).
//   insertion_fold_kernel   one warp per request walks the vehicles in the reference's order.  A vehicle
//       can only change the incumbent if its smallest cost is <= the bound C = veh.c + dc
//       (lib/Agents.py:1470); only then is the pair re-scanned exactly as the reference does it -- with
//       the bound, "last equal-or-better wins" and the viol-code loop breaks.
//   insertion_commit_kernel / insertion_rescan_kernel   the sequential queue: after a request is folded
//       its winner's plan is patched on the device (build_route's route, veh.c and pwt), that vehicle is
//       re-prepared and its column is re-evaluated for the requests still in the queue.
#include <algorithm>
#include <cstring>

#include "drs_paths.cuh"
#include "drs_tick.cuh"

namespace {

constexpr int INS_MAXL = 16;            // plan slots per vehicle on the device
constexpr int INS_SCAN_MAXL = 14;       // longest plan a request can still be inserted into (L + 2 <= INS_MAXL)

struct PlanView {
  int n_veh;
  // per vehicle, indexed by sorted position vs (perm[vs] = caller's vehicle index, inv[v] = vs)
  const int32_t *perm, *inv;
  const int32_t *node, *onboard, *cap;
  int32_t* len;
  const double *delay, *clock;
  double* cost;
  // what a later build_route keeps in front of a new route (lib/Agents.py:319-336)
  const int32_t* pos_node;             // vertex of the vehicle's position, -1 between vertices
  int32_t* keep_node;                  // -1: the vehicle has no route
  double *keep_t, *keep_et;
  int32_t* keep_pend;                  // >= 0: the vehicle got its first route in this tick; its kept step is the first step of the
                                       //       leg pos_node -> keep_pend, found (a path walk) only if the vehicle wins again
  // request table (Cld and pwt are updated by commits)
  const int32_t *q_o, *q_d;
  const double *q_cep, *q_clp, *q_ts, *q_tr;
  double *q_cld, *q_pwt;
  drs_insertion_params prm;
  // slot-major plan [k][vs]
  int32_t *m_node, *m_req;
  int8_t *m_pod, *m_partner;           // partner: plan index of the pickup belonging to a drop-off, -1 if on board
  // per-vehicle precomputation (plan_prep): plan segment times and the walk of the unmodified plan
  double* m_seg;                       // [k][vs]  tt(previous stop or start vertex -> stop k)
  double* m_base_t;                    // [k][vs]  t before plan stop k, k = 0..L
  double* m_base_c;                    // [k][vs]  c before plan stop k
  double* m_base_cld;                  // [k][vs]  Cld assigned when plan stop k is a pickup
  int32_t* m_base_viol;                // [vs]     first plan stop the unmodified walk fails at (L if none), -1: over capacity
  float* m_lb;                         // [k][vs]  fp32 lower bound of clock + base_t[k] (reach test), +inf past the usable slots
  uint8_t* m_nslots;                   // [vs]     pickup positions worth testing: min(L, base_viol) + 1, 0 if none
  int directed;
};

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// Travel times between a request vertex q (origin or destination) and a plan vertex x.  On an undirected map the
// matrix is symmetric and EVERY kernel reads such a time from q's row (the reach test stages that row in shared
// memory); reading one direction everywhere keeps the prefilter, the evaluation and the exact re-scan consistent
// with each other even where an fp32 matrix is not bitwise symmetric (float-weighted maps).
__device__ __forceinline__ double tt_into(const DistView& dv, int directed, int q, int x) {   // x -> q
  return x == q ? 0.0 : (directed ? dv.at(x, q) : dv.at(q, x));
}
__device__ __forceinline__ double tt_outof(const DistView& dv, int q, int x) {                 // q -> x
  return x == q ? 0.0 : dv.at(q, x);
}

// Everything one (vehicle, request) pair needs, in thread-local storage.
struct Pair {
  int L, n0, K;
  double T, t0, c0;
  int8_t pod[INS_MAXL];
  int8_t partner[INS_MAXL];
  int req[INS_MAXL];
  double seg[INS_MAXL];         // seg[k]  = tt(stop k-1 -> stop k), stop -1 = start vertex
  double to_o[INS_MAXL + 1];    // to_o[k] = tt(stop k-1 -> origin)      (k = 0: from the start vertex)
  double to_d[INS_MAXL + 1];    // to_d[k] = tt(stop k-1 -> destination)
  double from_o[INS_MAXL];      // from_o[k] = tt(origin -> stop k)
  double from_d[INS_MAXL];      // from_d[k] = tt(destination -> stop k)
  double od;                    // tt(origin -> destination)
  int new_req;
};

__device__ void load_pair(Pair& p, const PlanView& pv, const DistView& dv, int vs, int rrow) {
  const int nv = pv.n_veh;
  p.L = pv.len[vs];
  p.n0 = pv.onboard[vs]; p.K = pv.cap[vs];
  p.T = pv.clock[vs]; p.t0 = pv.delay[vs]; p.c0 = pv.cost[vs];
  p.new_req = rrow;
  const int o = pv.q_o[rrow], d = pv.q_d[rrow];
  const int start = pv.node[vs];
  p.to_o[0] = tt_into(dv, pv.directed, o, start);
  p.to_d[0] = tt_into(dv, pv.directed, d, start);
  for (int k = 0; k < p.L; ++k) {
    const int nd = pv.m_node[k * nv + vs];
    p.pod[k] = pv.m_pod[k * nv + vs];
    p.partner[k] = pv.m_partner[k * nv + vs];
    p.req[k] = pv.m_req[k * nv + vs];
    p.seg[k] = pv.m_seg[k * nv + vs];
    p.from_o[k] = tt_outof(dv, o, nd);
    p.from_d[k] = tt_outof(dv, d, nd);
    p.to_o[k + 1] = tt_into(dv, pv.directed, o, nd);
    p.to_d[k + 1] = tt_into(dv, pv.directed, d, nd);
  }
  p.od = tt_outof(dv, o, d);
}

// test_constraints_get_cost (lib/Agents.py:1493-1572) for the candidate "pickup after a
// planned stops, drop-off after b planned stops" (a <= b), i.e. (i, j) = (a, b + 1).
// Returns viol (-1 = success) and the total cost through *c_out.
// FX: also record the Req.Cld values the walk assigns at every pickup it processes (lib/Agents.py:1550,1554 mutate the
// request objects while a candidate is merely evaluated): fx_cld[k] for planned stop k, fx_cld[L] for the new request,
// bit k / L of *fx_mask set.  A walk that fails later keeps what it wrote before, like the reference.
// mpre_out (with C = +inf): the largest running cost after any stop walked BEFORE the stop that violates (all stops if
// none does).  Under a finite bound C the reference's walk returns 0 at the first stop whose running cost exceeds C, and
// a violation is tested before the cost of its own stop, so: result(C) = 1 if over capacity, else 0 if mpre > C, else
// the unbounded result -- which lets the lanes of a warp walk the candidates of a vehicle in parallel, without the
// bound, and one pass over their (viol, mpre, cost) triples replay the sequential scan exactly (scan_vehicle_warp).
template <bool FX>
__device__ int eval_candidate(const Pair& p, const PlanView& pv, int a, int b, double C, double* c_out, double* fx_cld = nullptr,
                              uint32_t* fx_mask = nullptr, double* mpre_out = nullptr) {
  // capacity pre-check over the whole candidate route (:1518-1521)
  {
    int n = p.n0;
    for (int k = 0; k <= p.L; ++k) {
      if (k == a) { n += 1; if (n > p.K) return 1; }
      if (k == b) { n -= 1; }
      if (k < p.L) { n += p.pod[k]; if (n > p.K) return 1; }
    }
  }
  double cld_local[INS_MAXL];        // Cld of requests picked up inside this candidate (:1550,1554)
  double cld_new = 0.0;
  double c = 0.0, t = 0.0 + p.t0;
  int n = p.n0;
  const double T = p.T;
  // sequence: planned stops 0..L-1 with the pickup before stop a and the drop-off before stop b
  // (after the pickup when a == b)
  int last = -1;                     // -1 start vertex, 0..L-1 planned stop, L pickup, L+1 drop-off
  for (int pos = 0; pos < p.L + 2; ++pos) {
    int kind;                        // 0 planned stop, 1 new pickup, 2 new drop-off
    int k = 0;
    // position layout: [stops 0..a-1] P [stops a..b-1] D [stops b..L-1]
    if (pos < a) { kind = 0; k = pos; }
    else if (pos == a) { kind = 1; }
    else if (pos <= b) { kind = 0; k = pos - 1; }
    else if (pos == b + 1) { kind = 2; }
    else { kind = 0; k = pos - 2; }
    double dt;
    int pod, rq;
    if (kind == 0) {
      if (last == p.L) dt = p.from_o[k];
      else if (last == p.L + 1) dt = p.from_d[k];
      else dt = p.seg[k];            // previous element is stop k-1 (or the start): the plan's own segment
      pod = p.pod[k]; rq = p.req[k];
    } else if (kind == 1) {
      dt = p.to_o[a];
      pod = 1; rq = p.new_req;
    } else {
      dt = (last == p.L) ? p.od : p.to_d[b];
      pod = -1; rq = p.new_req;
    }
    t += dt;
    if (pod == 1) {
      const double cep = pv.q_cep[rq];
      double cld;
      if (T + t < cep) {             // early: wait until Cep (:1547-1550)
        dt += cep - T - t;
        t += cep - T - t;
        cld = cep + pv.prm.max_detour * pv.q_ts[rq];
      } else if (T + t > pv.q_clp[rq]) {
        return 2;                    // late pickup (:1551-1552)
      } else {
        cld = T + t + pv.prm.max_detour * pv.q_ts[rq];
      }
      if (kind == 1) cld_new = cld; else cld_local[k] = cld;
      if (FX) { const int at = kind == 1 ? p.L : k; fx_cld[at] = cld; *fx_mask |= 1u << at; }
    } else if (pod == -1) {
      double cld;
      if (kind == 2) cld = cld_new;
      else cld = p.partner[k] >= 0 ? cld_local[p.partner[k]] : pv.q_cld[rq];   // set by its pickup earlier in this walk, else stored
      if (T + t > cld) return 3;     // late drop-off (:1555-1556)
    }
    c += n * dt * pv.prm.coef_inveh;
    n += pod;
    if (pod == 1) {
      const double pwt = pv.q_pwt[rq];
      if (pwt > 0) c += (T + t - pv.q_tr[rq] - pwt) * pv.prm.coef_extra_wait;
      else c += t * pv.prm.coef_wait;
    }
    if (c > C) return 0;             // over the incumbent bound (:1568-1569)
    if (mpre_out) *mpre_out = fmax(*mpre_out, c);
    last = (kind == 0) ? k : (kind == 1 ? p.L : p.L + 1);
  }
  *c_out = c;
  return -1;
}

// The reference's scan of ONE vehicle (lib/Agents.py:1466-1482) by a whole warp: lanes walk the (i, j) candidates in
// parallel without the bound, then every lane replays the sequential decisions from the triples (identically, so the
// warp stays converged).  dc / wi / wj are updated like dc_, route_ in the reference; returns true if the vehicle won.
__device__ bool scan_vehicle_warp(const Pair& p, const PlanView& pv, double& dc, int& wi, int& wj) {
  const int lane = threadIdx.x & 31;
  const int ncand = (p.L + 1) * (p.L + 2) / 2;
  const double c0 = p.c0;
  bool won = false, over = false;
  int skip_i = -1;                       // the j-loop of this i was left (viol > 0)
  for (int base = 0; base < ncand && !over; base += 32) {
    const int k = base + lane;
    int ci = 0, cj = 1, viol = 1;
    double F = 0.0, mpre = -d_inf();
    if (k < ncand) {
      int rem = k;                       // k-th candidate in scan order: i = 0..L, j = i+1..L+1
      while (rem >= p.L + 1 - ci) { rem -= p.L + 1 - ci; ++ci; }
      cj = ci + 1 + rem;
      viol = eval_candidate<false>(p, pv, ci, cj - 1, d_inf(), &F, nullptr, nullptr, &mpre);
    }
    const int nb = min(32, ncand - base);
    for (int q = 0; q < nb && !over; ++q) {
      const int qi = __shfl_sync(0xffffffffu, ci, q), qj = __shfl_sync(0xffffffffu, cj, q);
      const int qv = __shfl_sync(0xffffffffu, viol, q);
      const double qF = __shfl_sync(0xffffffffu, F, q), qM = __shfl_sync(0xffffffffu, mpre, q);
      if (qi == skip_i) continue;
      const int res = (qv == 1) ? 1 : ((qM > c0 + dc) ? 0 : qv);
      if (res < 0) { dc = qF - c0; wi = qi; wj = qj; won = true; }
      if (res > 0) skip_i = qi;
      if (res == 2) over = true;
    }
  }
  return won;
}

// One stop of test_constraints_get_cost's walk (lib/Agents.py:1523-1567) without the bound test.
struct Walk { double t, c; int n; };
__device__ __forceinline__ int walk_stop(Walk& w, double dt, int pod, int rq, double T, const PlanView& pv,
                                         double cld_drop, double* cld_set) {
  w.t += dt;
  if (pod == 1) {
    const double cep = pv.q_cep[rq];
    if (T + w.t < cep) {
      dt += cep - T - w.t;
      w.t += cep - T - w.t;
      *cld_set = cep + pv.prm.max_detour * pv.q_ts[rq];
    } else if (T + w.t > pv.q_clp[rq]) {
      return 2;
    } else {
      *cld_set = T + w.t + pv.prm.max_detour * pv.q_ts[rq];
    }
  } else if (pod == -1 && T + w.t > cld_drop) {
    return 3;
  }
  w.c += w.n * dt * pv.prm.coef_inveh;
  w.n += pod;
  if (pod == 1) {
    const double pwt = pv.q_pwt[rq];
    if (pwt > 0) w.c += (T + w.t - pv.q_tr[rq] - pwt) * pv.prm.coef_extra_wait;
    else w.c += w.t * pv.prm.coef_wait;
  }
  return -1;
}

// Per vehicle: plan segment times, the walk of the unmodified plan, and the reach-test bounds.
__device__ void plan_prep(const PlanView& pv, const DistView& dv, int vs) {
  const int nv = pv.n_veh;
  const int L = pv.len[vs];
  int prev = pv.node[vs];
  for (int k = 0; k < L; ++k) {
    const int nd = pv.m_node[k * nv + vs];
    pv.m_seg[k * nv + vs] = (prev == nd) ? 0.0 : dv.at(prev, nd);
    prev = nd;
  }
  int occ = pv.onboard[vs], viol_at = L;
  bool over = occ > pv.cap[vs];
  for (int k = 0; k < L; ++k) { occ += pv.m_pod[k * nv + vs]; if (occ > pv.cap[vs]) over = true; }
  Walk w{0.0 + pv.delay[vs], 0.0, pv.onboard[vs]};
  const double T = pv.clock[vs];
  for (int k = 0; k <= L; ++k) {
    pv.m_base_t[k * nv + vs] = w.t;
    pv.m_base_c[k * nv + vs] = w.c;
    if (k == L || viol_at < L) continue;
    const int pod = pv.m_pod[k * nv + vs], rq = pv.m_req[k * nv + vs];
    double cld_drop = pv.q_cld[rq];
    const int pk = pv.m_partner[k * nv + vs];
    if (pk >= 0) cld_drop = pv.m_base_cld[pk * nv + vs];
    double cld_set = 0.0;
    if (walk_stop(w, pv.m_seg[k * nv + vs], pod, rq, T, pv, cld_drop, &cld_set) >= 0) viol_at = k;
    pv.m_base_cld[k * nv + vs] = cld_set;
  }
  const int bviol = (over || L > INS_SCAN_MAXL) ? -1 : viol_at;   // plans too long to take two more stops are not insertable
  pv.m_base_viol[vs] = bviol;
  // reach test: pickup position a is worth a look iff clock + (base_t[a] + tt) can be <= the latest pickup time.
  // lb = that sum's vehicle part rounded DOWN to fp32 and one ulp further, so that an fp32 add and compare against the
  // rounded-UP limit can only err towards "reachable" (the exact fp64 test follows in insertion_eval_kernel).
  const int ns = bviol < 0 ? 0 : min(L, bviol) + 1;
  pv.m_nslots[vs] = (uint8_t)ns;
  for (int k = 0; k <= INS_MAXL; ++k) {
    float lb = __int_as_float(0x7f800000);
    if (k < ns) {
      lb = __double2float_rd(T + pv.m_base_t[k * nv + vs]);
      lb = (lb > 0.f) ? __int_as_float(__float_as_int(lb) - 1) : ((lb < 0.f) ? __int_as_float(__float_as_int(lb) + 1) : -1.0e-30f);
    }
    pv.m_lb[k * nv + vs] = lb;
  }
}

// plan_prep for ONE vehicle whose plan the caller holds in local arrays (the commit): everything it needs from memory is
// gathered up front (independent loads), the walk then runs in registers, the results are stored at the end -- the same
// values as plan_prep, without a memory round trip per stop on the queue kernel's critical path.
__device__ void plan_prep_local(const PlanView& pv, const DistView& dv, int vs, int L, const int* node, const int* req, const int8_t* pod,
                                const int8_t* partner) {
  const int nv = pv.n_veh;
  double seg[INS_MAXL], cep[INS_MAXL], clp[INS_MAXL], ts[INS_MAXL], tr[INS_MAXL], pwt[INS_MAXL], cldq[INS_MAXL];
  int prev = pv.node[vs];
  for (int k = 0; k < L; ++k) {
    seg[k] = (prev == node[k]) ? 0.0 : dv.at(prev, node[k]);
    prev = node[k];
    const int rq = req[k];
    cep[k] = pv.q_cep[rq]; clp[k] = pv.q_clp[rq]; ts[k] = pv.q_ts[rq]; tr[k] = pv.q_tr[rq]; pwt[k] = pv.q_pwt[rq]; cldq[k] = pv.q_cld[rq];
  }
  const int onboard = pv.onboard[vs], cap = pv.cap[vs];
  const double T = pv.clock[vs];
  int occ = onboard, viol_at = L;
  bool over = occ > cap;
  for (int k = 0; k < L; ++k) { occ += pod[k]; if (occ > cap) over = true; }
  double base_t[INS_MAXL + 1], base_c[INS_MAXL + 1], base_cld[INS_MAXL];
  Walk w{0.0 + pv.delay[vs], 0.0, onboard};
  for (int k = 0; k <= L; ++k) {
    base_t[k] = w.t; base_c[k] = w.c;
    if (k == L || viol_at < L) continue;
    // one stop of the walk (walk_stop with the request's fields already in registers)
    double dt = seg[k], cld_set = 0.0;
    bool bad = false;
    w.t += dt;
    if (pod[k] == 1) {
      if (T + w.t < cep[k]) { dt += cep[k] - T - w.t; w.t += cep[k] - T - w.t; cld_set = cep[k] + pv.prm.max_detour * ts[k]; }
      else if (T + w.t > clp[k]) bad = true;
      else cld_set = T + w.t + pv.prm.max_detour * ts[k];
    } else if (pod[k] == -1) {
      const double cld_drop = partner[k] >= 0 ? base_cld[partner[k]] : cldq[k];
      if (T + w.t > cld_drop) bad = true;
    }
    if (bad) {
      viol_at = k;
    } else {
      w.c += w.n * dt * pv.prm.coef_inveh;
      w.n += pod[k];
      if (pod[k] == 1) {
        if (pwt[k] > 0) w.c += (T + w.t - tr[k] - pwt[k]) * pv.prm.coef_extra_wait;
        else w.c += w.t * pv.prm.coef_wait;
      }
    }
    base_cld[k] = cld_set;
  }
  const int bviol = (over || L > INS_SCAN_MAXL) ? -1 : viol_at;
  const int ns = bviol < 0 ? 0 : min(L, bviol) + 1;
  for (int k = 0; k < L; ++k) { pv.m_seg[k * nv + vs] = seg[k]; pv.m_base_cld[k * nv + vs] = base_cld[k]; }
  for (int k = 0; k <= L; ++k) { pv.m_base_t[k * nv + vs] = base_t[k]; pv.m_base_c[k * nv + vs] = base_c[k]; }
  pv.m_base_viol[vs] = bviol;
  pv.m_nslots[vs] = (uint8_t)ns;
  for (int k = 0; k <= INS_MAXL; ++k) {
    float lb = __int_as_float(0x7f800000);
    if (k < ns) {
      lb = __double2float_rd(T + base_t[k]);
      lb = (lb > 0.f) ? __int_as_float(__float_as_int(lb) - 1) : ((lb < 0.f) ? __int_as_float(__float_as_int(lb) + 1) : -1.0e-30f);
    }
    pv.m_lb[k * nv + vs] = lb;
  }
}

__global__ void plan_prep_kernel(PlanView pv, DistView dv) {
  const int vs = blockIdx.x * blockDim.x + threadIdx.x;
  if (vs < pv.n_veh) plan_prep(pv, dv, vs);
}

// ------------------------------------------------------------------------------------------------ reach test
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// fp32 upper bound of the latest time a pickup may still happen: max(Clp, Cep) rounded up, one ulp further
__device__ __forceinline__ float reach_limit(const PlanView& pv, int rq) {
  const double lim = fmax(pv.q_clp[rq], pv.q_cep[rq]);
  float f = __double2float_ru(lim);
  if (f > 0.f) f = __int_as_float(__float_as_int(f) + 1);
  else if (f < 0.f) f = __int_as_float(__float_as_int(f) - 1);
  else f = 1.0e-30f;
  return f;
}

// STAGED: the origin row of each request is copied into shared memory by the bulk-copy engine (one elected thread
// issues cp.async.bulk in 32 KB pieces, all threads wait on the mbarrier); otherwise the same gathers go to L2.
template <bool STAGED, bool I32>
__global__ void __launch_bounds__(1024, 1) insertion_reach_kernel(PlanView pv, DistView dv, int n_new, const int32_t* __restrict__ new_rows,
                                                                    uint32_t* __restrict__ mask, int words, int2* __restrict__ list,
                                                                    int list_cap, int* __restrict__ counters /* [0] list fill, [1] next request */) {
  extern __shared__ __align__(128) unsigned char reach_smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int s_req;
  const int nv = pv.n_veh, tid = threadIdx.x, lane = tid & 31;
  const size_t row_bytes = (size_t)dv.ld * 4;
  const float* frow = reinterpret_cast<const float*>(reach_smem);
  const int* irow = reinterpret_cast<const int*>(reach_smem);
  if (STAGED && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;
  if (tid == 0) s_req = atomicAdd(&counters[1], 1);
  while (true) {
    __syncthreads();                                   // every lane is done with the previous row (and sees s_req)
    const int r = s_req;
    if (r >= n_new) break;
    const int rq = new_rows[r];
    const int o = pv.q_o[rq];
    if (STAGED) {
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)row_bytes) : "memory");
        const char* src = reinterpret_cast<const char*>(dv.dist) + (size_t)o * row_bytes;
        for (size_t off = 0; off < row_bytes; off += 32768) {
          const uint32_t bytes = (uint32_t)min((size_t)32768, row_bytes - off);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(reach_smem + off)), "l"(src + off), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
        }
      }
      // wait for the bytes (all threads; try_wait suspends the warp in hardware for a bounded time)
      uint32_t landed = 0;
      while (!landed)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(landed) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    __syncthreads();                                   // everybody has read s_req
    if (tid == 0) {
      // claim the next request now and pull its row towards L2 while this one is being tested
      const int nxt = atomicAdd(&counters[1], 1);
      s_req = nxt;
      if (STAGED && nxt < n_new) {
        const char* src = reinterpret_cast<const char*>(dv.dist) + (size_t)pv.q_o[new_rows[nxt]] * row_bytes;
        for (size_t off = 0; off < row_bytes; off += 32768) {
          const uint32_t bytes = (uint32_t)min((size_t)32768, row_bytes - off);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(bytes) : "memory");
        }
      }
    }
    const float limit = reach_limit(pv, rq);
    uint32_t* mrow = mask + (size_t)r * words;
    for (int vs0 = tid - lane; vs0 < nv; vs0 += 1024) {
      const int vs = vs0 + lane;
      bool reach = false;
      {
        const int ns = vs < nv ? pv.m_nslots[vs] : 0;
        // the vehicle's bounds and vertices for ALL its positions first (independent loads in flight together: the
        // kernel was bound by one L2 round trip per position), then the gathers and the compares.  Fully predicated on
        // purpose: a warp-uniform branch per position (skip what lies beyond the warp's longest plan) measured 13 %
        // SLOWER -- the branches keep the compiler from issuing the loads together.
        constexpr int U = 9;                              // positions handled in registers (plans up to 8 stops); the rest in the loop below
        float lb[U];
        int xs[U];
#pragma unroll
        for (int a = 0; a < U; ++a) {
          lb[a] = __int_as_float(0x7f800000);
          xs[a] = 0;
          if (a < ns) {
            lb[a] = pv.m_lb[a * nv + vs];
            xs[a] = a == 0 ? pv.node[vs] : pv.m_node[(a - 1) * nv + vs];
          }
        }
#pragma unroll
        for (int a = 0; a < U; ++a) {
          float w;
          if (STAGED) w = I32 ? (irow[xs[a]] >= DRS_I32_INF ? __int_as_float(0x7f800000) : __int2float_rd(irow[xs[a]])) : frow[xs[a]];
          else w = a < ns ? __double2float_rd(tt_into(dv, pv.directed, o, xs[a])) : 0.f;
          reach = reach || (lb[a] + w <= limit);          // lb = +inf past the usable positions
        }
        for (int a = U; a < ns && !reach; ++a) {
          const int x = pv.m_node[(a - 1) * nv + vs];
          float w;
          if (STAGED) w = I32 ? (irow[x] >= DRS_I32_INF ? __int_as_float(0x7f800000) : __int2float_rd(irow[x])) : frow[x];
          else w = __double2float_rd(tt_into(dv, pv.directed, o, x));
          reach = pv.m_lb[a * nv + vs] + w <= limit;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, reach);
      if (lane == 0) mrow[vs0 >> 5] = m;
      if (m) {
        int b0 = 0;
        if (lane == 0) b0 = atomicAdd(&counters[0], __popc(m));
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (reach) {
          const int pos = b0 + __popc(m & ((1u << lane) - 1u));
          if (pos < list_cap) list[pos] = make_int2(r, vs);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ evaluation
// Smallest successful total cost of one (request, vehicle) pair over all (i, j), WITHOUT the incumbent bound:
//   * the unmodified-plan prefix is shared (precomputed per vehicle), the state "pickup inserted after a stops, b-a
//     further stops served" is advanced incrementally over b, and only the tail after the drop-off is re-walked per
//     candidate.  Arithmetic and operation order per candidate are those of eval_candidate; only successes matter
//     here, so no viol codes.
__device__ double pair_cmin(const PlanView& pv, const DistView& dv, int rq, int vs, bool* reach_out = nullptr) {
  const int nv = pv.n_veh;
  const double inf = d_inf();
  const int L = pv.len[vs];
  double cmin = inf;
  if (reach_out) *reach_out = false;
  const int bviol = pv.m_base_viol[vs];
  if (bviol < 0) return inf;
  const int o = pv.q_o[rq], d = pv.q_d[rq];
  const int K = pv.cap[vs];
  const double T = pv.clock[vs];
  double to_o[INS_MAXL + 1], to_d[INS_MAXL + 1], from_o[INS_MAXL], from_d[INS_MAXL], cld_cur[INS_MAXL];
  // origin-side travel times first: most vehicles cannot reach the pickup in time at any position, and
  // for those the destination-side gathers (half of the traffic) are never needed
  const int start = pv.node[vs];
  to_o[0] = tt_into(dv, pv.directed, o, start);
  for (int k = 0; k < L; ++k) {
    const int nd = pv.m_node[k * nv + vs];
    from_o[k] = tt_outof(dv, o, nd);
    to_o[k + 1] = pv.directed ? tt_into(dv, 1, o, nd) : from_o[k];
  }
  bool reachable = false;
  {
    const double clp = pv.q_clp[rq], cep = pv.q_cep[rq];
    for (int a = 0; a <= L && a <= bviol; ++a) {
      const double t = pv.m_base_t[a * nv + vs] + to_o[a];      // same sum walk_stop forms for the pickup
      if (T + t < cep || !(T + t > clp)) { reachable = true; break; }
    }
  }
  if (!reachable) return inf;
  if (reach_out) *reach_out = true;
  to_d[0] = tt_into(dv, pv.directed, d, start);
  for (int k = 0; k < L; ++k) {
    const int nd = pv.m_node[k * nv + vs];
    from_d[k] = tt_outof(dv, d, nd);
    to_d[k + 1] = pv.directed ? tt_into(dv, 1, d, nd) : from_d[k];
  }
  const double od = tt_outof(dv, o, d);
  int occ_before = pv.onboard[vs];   // occupancy before plan stop a in the unmodified plan
  for (int a = 0; a <= L && a <= bviol; ++a) {
    if (a > 0) {
      cld_cur[a - 1] = pv.m_base_cld[(a - 1) * nv + vs];
      occ_before += pv.m_pod[(a - 1) * nv + vs];
    }
    if (occ_before + 1 > K) continue;
    Walk S{pv.m_base_t[a * nv + vs], pv.m_base_c[a * nv + vs], occ_before};
    double cld_new = 0.0;
    if (walk_stop(S, to_o[a], 1, rq, T, pv, 0.0, &cld_new) >= 0) continue;   // late pickup
    int occ_after = occ_before;      // occupancy of the unmodified plan after stop b-1
    for (int b = a; b <= L; ++b) {
      // ---- candidate (a, b): drop-off now, then the rest of the plan
      Walk W = S;
      double dummy;
      if (walk_stop(W, b == a ? od : to_d[b], -1, rq, T, pv, cld_new, &dummy) < 0) {
        bool ok = true;
        for (int k = b; k < L; ++k) {
          const int pod = pv.m_pod[k * nv + vs], q = pv.m_req[k * nv + vs];
          double cld_drop = 0.0;
          if (pod < 0) { const int pk = pv.m_partner[k * nv + vs]; cld_drop = pk >= 0 ? cld_cur[pk] : pv.q_cld[q]; }
          if (walk_stop(W, k == b ? from_d[k] : pv.m_seg[k * nv + vs], pod, q, T, pv, cld_drop, &cld_cur[k]) >= 0) { ok = false; break; }
        }
        if (ok && W.c < cmin) cmin = W.c;
      }
      // ---- serve plan stop b with the new rider on board
      if (b == L) break;
      const int pod = pv.m_pod[b * nv + vs], q = pv.m_req[b * nv + vs];
      occ_after += pod;
      if (occ_after + 1 > K) break;                   // over capacity for every larger b
      double cld_drop = 0.0;
      if (pod < 0) { const int pk = pv.m_partner[b * nv + vs]; cld_drop = pk >= 0 ? cld_cur[pk] : pv.q_cld[q]; }
      if (walk_stop(S, b == a ? from_o[b] : pv.m_seg[b * nv + vs], pod, q, T, pv, cld_drop, &cld_cur[b]) >= 0) break;
    }
  }
  return cmin;
}

// One thread per surviving pair.  (Eight lanes per pair, one pickup position each through pair_cmin_at and a shuffle
// minimum, measured 18 % slower on the C4 tick: every lane rebuilds the prefix state and repeats the shared gathers.)
__global__ void __launch_bounds__(128) insertion_eval_kernel(PlanView pv, DistView dv, const int32_t* __restrict__ new_rows,
                                                             const int2* __restrict__ list, const int* __restrict__ counters, int list_cap,
                                                             double* __restrict__ cmin_out) {
  const int n = min(counters[0], list_cap);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int2 e = list[idx];
    cmin_out[(int64_t)e.x * pv.n_veh + e.y] = pair_cmin(pv, dv, new_rows[e.x], e.y);
  }
}

// ------------------------------------------------------------------------------------------------ fold
// One WARP per request.  Lanes read the smallest costs of 32 consecutive vehicles IN THE CALLER'S ORDER (eight chunks
// ahead: the walk is a chain of dependent decisions, the loads are not); only flagged vehicles are re-scanned (by
// their lane) with the running bound.  r_first / r_count select the requests (the sequential queue folds one at a time).
struct FoldEvent { int v; int pad; double dc; };   // after the exact re-scan of vehicle v the incumbent cost increase is dc
constexpr int FOLD_EVENT_CAP = 8192;

__global__ void insertion_fold_dense_kernel(PlanView pv, DistView dv, int r_first, int r_count, const int32_t* new_rows, const uint32_t* mask,
                                      int words, const double* cmin, int32_t* best_veh, int32_t* best_i, int32_t* best_j, double* best_dc,
                                      FoldEvent* events /* may be null; single-request folds only */, int* n_events) {
  const int rr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (rr >= r_count) return;
  const int r = r_first + rr;
  const double inf = d_inf();
  double dc = inf;                   // dc_ = np.inf (:1452)
  int wv = -1, wi = -1, wj = -1;
  int n_ev = 0;
  const uint32_t* mrow = mask + (size_t)r * words;
  constexpr int AHEAD = 8;
  for (int base0 = 0; base0 < pv.n_veh; base0 += 32 * AHEAD) {
    double cms[AHEAD], c0s[AHEAD];
    int vss[AHEAD];
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int v = base0 + 32 * u + lane;
      cms[u] = inf; c0s[u] = 0.0; vss[u] = 0;
      if (v < pv.n_veh) {
        const int vs = pv.inv[v];
        vss[u] = vs;
        if ((mrow[vs >> 5] >> (vs & 31)) & 1u) { cms[u] = cmin[(int64_t)r * pv.n_veh + vs]; c0s[u] = pv.cost[vs]; }
      }
    }
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int v = base0 + 32 * u + lane;
      const double cm = cms[u], c0 = c0s[u];
      unsigned m = __ballot_sync(0xffffffffu, cm != inf && !(cm > c0 + dc));
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        double ndc = dc; int nv = wv, ni = wi, nj = wj;
        if (lane == l && !(cm > c0 + dc)) {              // re-check: dc may have dropped inside this chunk
          Pair p;
          load_pair(p, pv, dv, vss[u], new_rows[r]);
          int viol = -1;
          for (int i = 0; i <= p.L; ++i) {                         // (:1466-1482)
            for (int j = i + 1; j <= p.L + 1; ++j) {
              double c;
              viol = eval_candidate<false>(p, pv, i, j - 1, c0 + ndc, &c);
              if (viol < 0) { ndc = c - c0; nv = v; ni = i; nj = j; }
              if (viol > 0) break;
            }
            if (viol == 2) break;
          }
        }
        dc = __shfl_sync(0xffffffffu, ndc, l); wv = __shfl_sync(0xffffffffu, nv, l);
        wi = __shfl_sync(0xffffffffu, ni, l); wj = __shfl_sync(0xffffffffu, nj, l);
        if (events && lane == 0) {
          if (n_ev < FOLD_EVENT_CAP) events[n_ev] = FoldEvent{base0 + 32 * u + l, 0, dc};
          ++n_ev;
        }
      }
    }
  }
  if (lane == 0) { best_veh[r] = wv; best_i[r] = wi; best_j[r] = wj; best_dc[r] = dc; if (events) *n_events = n_ev; }
}

// Fold of the frozen-plan evaluation: one CTA per request.  The request's mask (n_veh bits) is scanned by all threads,
// the vehicles in reach (~1 %) go to a shared-memory list with their smallest cost, the list is ordered by the caller's
// vehicle index (rank by counting), and warp 0 walks it the way the reference walks the fleet: a vehicle is re-scanned
// (scan_vehicle_warp: candidates in parallel, decisions replayed in order) only if its smallest cost can meet the
// incumbent bound.  A request with more vehicles in reach than the list holds is flagged for the dense kernel.
constexpr int FOLD_THREADS = 128;
constexpr int FOLD_MAXC = 1024;

__global__ void __launch_bounds__(FOLD_THREADS) insertion_fold_kernel(PlanView pv, DistView dv, int n_new, const int32_t* __restrict__ new_rows,
                                                                       const uint32_t* __restrict__ mask, int words, const double* __restrict__ cmin,
                                                                       int32_t* best_veh, int32_t* best_i, int32_t* best_j, double* best_dc,
                                                                       int* n_overflow) {
  __shared__ int s_vs[FOLD_MAXC], s_v[FOLD_MAXC], s_order[FOLD_MAXC];
  __shared__ double s_cm[FOLD_MAXC];
  __shared__ int s_count;
  const int r = blockIdx.x, tid = threadIdx.x, nv = pv.n_veh;
  if (r >= n_new) return;
  const double inf = d_inf();
  if (tid == 0) s_count = 0;
  __syncthreads();
  const uint32_t* mrow = mask + (size_t)r * words;
  for (int w = tid; w < words; w += FOLD_THREADS) {
    uint32_t m = mrow[w];
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int vs = w * 32 + b;
      const double cm = cmin[(int64_t)r * nv + vs];
      if (cm == inf) continue;                           // in reach, but no candidate succeeds
      const int pos = atomicAdd(&s_count, 1);
      if (pos < FOLD_MAXC) { s_vs[pos] = vs; s_v[pos] = pv.perm[vs]; s_cm[pos] = cm; }
    }
  }
  __syncthreads();
  const int cnt = s_count;
  if (cnt > FOLD_MAXC) {
    if (tid == 0) { best_veh[r] = -2; atomicAdd(n_overflow, 1); }
    return;
  }
  for (int k = tid; k < cnt; k += FOLD_THREADS) {
    const int mine = s_v[k];
    int rank = 0;
    for (int m = 0; m < cnt; ++m) rank += s_v[m] < mine;
    s_order[rank] = k;
  }
  __syncthreads();
  if (tid >= 32) return;
  double dc = inf;                   // dc_ = np.inf (:1452)
  int wv = -1, wi = -1, wj = -1;
  const int rq = new_rows[r];
  for (int k = 0; k < cnt; ++k) {
    const int at = s_order[k];
    const int vs = s_vs[at];
    const double c0 = pv.cost[vs];
    if (s_cm[at] > c0 + dc) continue;
    Pair p;
    load_pair(p, pv, dv, vs, rq);
    if (scan_vehicle_warp(p, pv, dc, wi, wj)) wv = s_v[at];
  }
  if (tid == 0) { best_veh[r] = wv; best_i[r] = wi; best_j[r] = wj; best_dc[r] = dc; }
}

// Side effects of one request's scan on Req.Cld (lib/Agents.py:1550,1554), exactly: every vehicle whose pickup position 0
// can be reached (bit set in the request's mask; for all others the very first walk ends at the late pickup, viol 2, and
// the scan of that vehicle is over) repeats the reference's bounded scan -- the bound it starts with is the incumbent the
// fold had when it passed that vehicle (events) -- and the LAST value each pickup received wins: per planned pickup within
// its vehicle, for the new request across vehicles in the caller's order (fx_new / fx_last_v, applied by the commit).
__global__ void insertion_sidefx_kernel(PlanView pv, DistView dv, const int32_t* new_rows, int r, const uint32_t* mask, int words,
                                        const FoldEvent* events, const int* n_events, double* fx_new, int* fx_last_v, int* err) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= pv.n_veh) return;
  const int vs = pv.inv[v];
  if (!((mask[(size_t)r * words + (vs >> 5)] >> (vs & 31)) & 1u)) return;
  const int ne = *n_events;
  if (ne > FOLD_EVENT_CAP) { *err = 4; return; }
  double dc = d_inf();
  for (int e = 0; e < ne && events[e].v < v; ++e) dc = events[e].dc;
  Pair p;
  load_pair(p, pv, dv, vs, new_rows[r]);
  double fx_cld[INS_MAXL + 1];
  uint32_t fx_mask = 0;
  const double c0 = p.c0;
  int viol = -1;
  for (int i = 0; i <= p.L; ++i) {
    for (int j = i + 1; j <= p.L + 1; ++j) {
      double c;
      viol = eval_candidate<true>(p, pv, i, j - 1, c0 + dc, &c, fx_cld, &fx_mask);
      if (viol < 0) dc = c - c0;
      if (viol > 0) break;
    }
    if (viol == 2) break;
  }
  for (int k = 0; k < p.L; ++k)
    if ((fx_mask >> k) & 1u) pv.q_cld[p.req[k]] = fx_cld[k];
  if ((fx_mask >> p.L) & 1u) { fx_new[vs] = fx_cld[p.L]; atomicMax(fx_last_v, v); }
}

// ------------------------------------------------------------------------------------------------ commit
// What Veh.build_route(router, route_, reqs, T) leaves behind for the next scan (lib/Agents.py:148-297), for the
// winner of request r: the route with the pickup inserted at i and the drop-off at j (:1468-1469), veh.c (:256-270),
// pwt of the requests still to be picked up (:165-220), and -- when the vehicle had no route -- the first step of
// the new one, which a later build_route would keep (:319-336).  One thread; then the vehicle is re-prepared.
__device__ void commit_one(const PlanView& pv, const DistView& dv, const PredSource& ps, const int32_t* esrc, const int32_t* edst,
                           const double* tt64, int rq, int v, int i, int j, double T_now, int* err) {
  const int nv = pv.n_veh;
  const int vs = pv.inv[v];
  const int L = pv.len[vs];
  if (L + 2 > INS_MAXL || i < 0 || i > L || j <= i || j > L + 1) { *err = 1; return; }
  int node[INS_MAXL], req[INS_MAXL];
  int8_t pod[INS_MAXL];
  {
    int k2 = 0;
    for (int k = 0; k <= L; ++k) {                       // route.insert(i, pickup)
      if (k == i) { node[k2] = pv.q_o[rq]; req[k2] = rq; pod[k2] = 1; ++k2; }
      if (k < L) { node[k2] = pv.m_node[k * nv + vs]; req[k2] = pv.m_req[k * nv + vs]; pod[k2] = pv.m_pod[k * nv + vs]; ++k2; }
    }
    for (int k = L + 1; k > j; --k) { node[k] = node[k - 1]; req[k] = req[k - 1]; pod[k] = pod[k - 1]; }   // route.insert(j, drop-off)
    node[j] = pv.q_d[rq]; req[j] = rq; pod[j] = -1;
  }
  const int L2 = L + 2;
  int8_t partners[INS_MAXL];
  for (int k = 0; k < L2; ++k) {
    int8_t partner = -1;
    if (pod[k] < 0)
      for (int m = 0; m < k; ++m)
        if (req[m] == req[k] && pod[m] > 0) partner = (int8_t)m;
    partners[k] = partner;
    pv.m_node[k * nv + vs] = node[k]; pv.m_req[k * nv + vs] = req[k]; pv.m_pod[k * nv + vs] = pod[k]; pv.m_partner[k * nv + vs] = partner;
  }
  pv.len[vs] = L2;
  if (pv.keep_pend[vs] >= 0) {
    // second win of a vehicle that had no route at the start of the tick: NOW its kept step is needed -- the first step of
    // the leg it was given at its first win, along the canonical path (walk the predecessor chain back from that stop)
    const int from = pv.pos_node[vs], to = pv.keep_pend[vs];
    int hop_to = from;                                   // zero-length leg: the dummy stay-in-place step (lib/RoutingEngine.py:173-189)
    double hop_t = 0.0;
    if (from != to) {
      int x = to, guard = 0;
      while (x != from && guard++ < (1 << 22)) {
        const int e = ps.last_edge(from, x);
        if (e < 0) { *err = 3; break; }
        const int y = prev_node(e, x, esrc, edst);
        if (y == from) { hop_to = x; hop_t = tt64[e]; }
        x = y;
      }
    }
    pv.keep_node[vs] = hop_to; pv.keep_t[vs] = hop_t; pv.keep_et[vs] = hop_t;
    pv.keep_pend[vs] = -1;
  }
  const bool merge = pv.keep_node[vs] >= 0;              // len(self.route) > 0 (:156-160)
  int cur = merge ? pv.keep_node[vs] : pv.pos_node[vs];
  if (cur < 0) { *err = 2; return; }                     // a vehicle without a route must stand on a vertex
  const int first_from = cur;
  // all leg times and pickup windows first (independent loads in flight together), then the dependent arithmetic
  double pred[INS_MAXL], cep[INS_MAXL], pwt[INS_MAXL];
  for (int k = 0; k < L2; ++k) {
    pred[k] = (cur == node[k]) ? 0.0 : dv.at(cur, node[k]);
    cep[k] = pod[k] > 0 ? pv.q_cep[req[k]] : 0.0;
    cur = node[k];
  }
  const double keep_t = merge ? pv.keep_t[vs] : 0.0, keep_et = merge ? pv.keep_et[vs] : 0.0;
  for (int k = 0; k < L2; ++k) pwt[k] = merge ? 0.0 + keep_et : 0.0;      // pwt = 0 (+ the kept step's expected time) (:186-193)
  double vt = merge ? 0.0 + keep_t : 0.0;
  double c = 0.0, t = 0.0;
  int n = pv.onboard[vs];
  for (int k = 0; k < L2; ++k) {
    const double pred_t = pred[k];
    double t_leg = pred_t;
    if (pod[k] > 0 && T_now + vt + t_leg < cep[k]) t_leg += cep[k] - (T_now + vt + t_leg);   // add_leg waits (:363-367)
    vt += t_leg;
    for (int m = k; m < L2; ++m)
      if (pod[m] > 0) pwt[m] += pred_t;                  // every request not yet picked up, this one included (:214-220)
    const double leg_t = (k == 0 && merge) ? keep_t + t_leg : t_leg;   // the kept step joins the first leg (:222-240)
    t += leg_t;
    c += n * leg_t * pv.prm.coef_inveh;
    n += pod[k];
    if (pod[k] > 0) c += t * pv.prm.coef_wait;
  }
  for (int k = 0; k < L2; ++k)
    if (pod[k] > 0) pv.q_pwt[req[k]] = pwt[k];
  pv.cost[vs] = c;
  if (!merge) pv.keep_pend[vs] = node[0];                // the kept step of this new route is looked up only if it is ever needed
  (void)first_from;
  plan_prep_local(pv, dv, vs, L2, node, req, pod, partners);
}

__global__ void insertion_commit_kernel(PlanView pv, DistView dv, PredSource ps, const int32_t* esrc, const int32_t* edst, const double* tt64,
                                        const int32_t* new_rows, int r, const int32_t* best_veh, const int32_t* best_i,
                                        const int32_t* best_j, double T_now, int* err, const double* fx_new, int* fx_last_v) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (fx_new) {                      // the last walk that picked the new request up left its Cld (side-effect pass)
    if (*fx_last_v >= 0) pv.q_cld[new_rows[r]] = fx_new[pv.inv[*fx_last_v]];
    *fx_last_v = -1;
  }
  const int v = best_veh[r];
  if (v < 0) return;
  commit_one(pv, dv, ps, esrc, edst, tt64, new_rows[r], v, best_i[r], best_j[r], T_now, err);
}

__global__ void insertion_commit_explicit_kernel(PlanView pv, DistView dv, PredSource ps, const int32_t* esrc, const int32_t* edst,
                                                 const double* tt64, int rq, int v, int i, int j, double T_now, int* err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  commit_one(pv, dv, ps, esrc, edst, tt64, rq, v, i, j, T_now, err);
}

// the committed vehicle's column for the requests still in the queue (rows r+1 ..): exact re-evaluation, mask bit and cost
__global__ void insertion_rescan_kernel(PlanView pv, DistView dv, int n_new, const int32_t* new_rows, int r, const int32_t* best_veh,
                                        uint32_t* mask, int words, double* cmin) {
  const int v = best_veh[r];
  if (v < 0) return;
  const int vs = pv.inv[v];
  for (int r2 = r + 1 + blockIdx.x * blockDim.x + threadIdx.x; r2 < n_new; r2 += gridDim.x * blockDim.x) {
    bool reach;
    const double cm = pair_cmin(pv, dv, new_rows[r2], vs, &reach);
    uint32_t* w = mask + (size_t)r2 * words + (vs >> 5);
    // the bit means "some pickup position can be reached in time" (the fold reads the cost only where it is set; the
    // side-effect pass repeats the reference's scan only there), whether or not any candidate succeeds
    if (reach) { cmin[(int64_t)r2 * pv.n_veh + vs] = cm; atomicOr(w, 1u << (vs & 31)); }
    else atomicAnd(w, ~(1u << (vs & 31)));
  }
}


// The whole queue in ONE launch: a single persistent CTA walks the requests in order (the chain of dependent decisions
// leaves nothing for a second CTA to do); per request
//   A  all threads: the vehicles whose bit is set in the request's mask, with their smallest cost, ordered by the
//      caller's vehicle index (rank by counting in shared memory),
//   B  warp 0: the fold over that short list -- a vehicle is re-scanned (scan_vehicle_warp) only if its smallest cost
//      can meet the incumbent bound -- recording the incumbent after every re-scan,
//   C  (DRS_QUEUE_TRACK_CLD) one thread per listed vehicle: the reference's bounded scan again, for the Req.Cld values
//      it leaves behind,
//   D  thread 0: commit (build_route's route, veh.c, pwt) and re-preparation of the winner,
//   E  all threads: the winner's column for the requests still queued.
constexpr int QUEUE_THREADS = 1024;
constexpr int QUEUE_MAXC = 3072;          // vehicles in reach of one request that the shared-memory list holds

struct QueueShared {
  int vs[QUEUE_MAXC];
  int v[QUEUE_MAXC];
  double cm[QUEUE_MAXC];
  double c0[QUEUE_MAXC];                  // veh.c of the listed vehicle
  int order[QUEUE_MAXC];                  // order[rank] = list position, ascending caller's vehicle index
  int ev_v[QUEUE_MAXC];                   // fold events: after the re-scan of vehicle ev_v the incumbent is ev_dc
  double ev_dc[QUEUE_MAXC];
  double fx_new[QUEUE_MAXC];              // Cld the scan of the k-th listed vehicle left on the new request
  int count, n_ev, win_v, win_i, win_j, last_fx;
  double win_dc;
};

__global__ void __launch_bounds__(QUEUE_THREADS, 1) insertion_queue_kernel(
    PlanView pv, DistView dv, PredSource ps, const int32_t* esrc, const int32_t* edst, const double* tt64, int n_new,
    const int32_t* new_rows, uint32_t* mask, int words, double* cmin, int32_t* best_veh, int32_t* best_i, int32_t* best_j,
    double* best_dc, double T_now, int track_cld, int* err,
    unsigned long long* prof /* cycles of thread 0 by phase A..E, [5] listed vehicles, [6] re-scans */) {
  extern __shared__ __align__(16) unsigned char queue_smem[];
  unsigned long long pacc[7] = {0, 0, 0, 0, 0, 0, 0};
  long long pt = clock64();
  auto lap = [&](int i) { const long long now = clock64(); pacc[i] += (unsigned long long)(now - pt); pt = now; };
  QueueShared& sh = *reinterpret_cast<QueueShared*>(queue_smem);
  const int tid = threadIdx.x, lane = tid & 31, nv = pv.n_veh;
  const double inf = d_inf();
  for (int r = 0; r < n_new; ++r) {
    const int rq = new_rows[r];
    if (tid == 0) { sh.count = 0; sh.n_ev = 0; sh.last_fx = -1; }
    __syncthreads();
    // ---- A: list of the vehicles in reach
    const uint32_t* mrow = mask + (size_t)r * words;
    for (int w = tid; w < words; w += QUEUE_THREADS) {
      uint32_t m = mrow[w];
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int vs = w * 32 + b;
        const int pos = atomicAdd(&sh.count, 1);
        if (pos < QUEUE_MAXC) { sh.vs[pos] = vs; sh.v[pos] = pv.perm[vs]; sh.cm[pos] = cmin[(int64_t)r * nv + vs]; sh.c0[pos] = pv.cost[vs]; }
      }
    }
    __syncthreads();
    const int cnt = sh.count;
    if (cnt > QUEUE_MAXC) { if (tid == 0) *err = 5; return; }
    for (int k = tid; k < cnt; k += QUEUE_THREADS) {
      const int mine = sh.v[k];
      int rank = 0;
      for (int m = 0; m < cnt; ++m) rank += sh.v[m] < mine;
      sh.order[rank] = k;
    }
    __syncthreads();
    lap(0);
    pacc[5] += cnt;
    // ---- B: fold (warp 0)
    if (tid < 32) {
      double dc = inf;
      int wv = -1, wi = -1, wj = -1, n_ev = 0;
      for (int k = 0; k < cnt; ++k) {
        const int at = sh.order[k];
        const double cm = sh.cm[at];
        const int vs = sh.vs[at];
        const double c0 = sh.c0[at];                     // veh.c, fetched with the list (a global load here would sit on the chain)
        if (cm == inf || cm > c0 + dc) continue;
        Pair p;
        load_pair(p, pv, dv, vs, rq);
        const bool won = scan_vehicle_warp(p, pv, dc, wi, wj);
        if (won) wv = sh.v[at];
        if (lane == 0) { sh.ev_v[n_ev] = sh.v[at]; sh.ev_dc[n_ev] = dc; }
        ++n_ev;
      }
      if (lane == 0) {
        sh.win_v = wv; sh.win_i = wi; sh.win_j = wj; sh.win_dc = dc; sh.n_ev = n_ev;
        best_veh[r] = wv; best_i[r] = wi; best_j[r] = wj; best_dc[r] = dc;
      }
    }
    __syncthreads();
    lap(1);
    pacc[6] += sh.n_ev;
    // ---- C: Req.Cld side effects of this scan (lib/Agents.py:1550,1554)
    if (track_cld) {
      for (int k = tid; k < cnt; k += QUEUE_THREADS) {
        const int at = sh.order[k];
        const int v = sh.v[at], vs = sh.vs[at];
        double dc = inf;
        for (int e = 0; e < sh.n_ev && sh.ev_v[e] < v; ++e) dc = sh.ev_dc[e];
        Pair p;
        load_pair(p, pv, dv, vs, rq);
        double fx_cld[INS_MAXL + 1];
        uint32_t fx_mask = 0;
        int viol = -1;
        for (int i = 0; i <= p.L; ++i) {
          for (int j = i + 1; j <= p.L + 1; ++j) {
            double c;
            viol = eval_candidate<true>(p, pv, i, j - 1, p.c0 + dc, &c, fx_cld, &fx_mask);
            if (viol < 0) dc = c - p.c0;
            if (viol > 0) break;
          }
          if (viol == 2) break;
        }
        for (int s2 = 0; s2 < p.L; ++s2)
          if ((fx_mask >> s2) & 1u) pv.q_cld[p.req[s2]] = fx_cld[s2];
        if ((fx_mask >> p.L) & 1u) { sh.fx_new[k] = fx_cld[p.L]; atomicMax(&sh.last_fx, k); }
      }
      __syncthreads();
      if (tid == 0 && sh.last_fx >= 0) pv.q_cld[rq] = sh.fx_new[sh.last_fx];   // the last vehicle in the caller's order wins
    }
    lap(2);
    // ---- D: commit
    if (tid == 0 && sh.win_v >= 0) commit_one(pv, dv, ps, esrc, edst, tt64, rq, sh.win_v, sh.win_i, sh.win_j, T_now, err);
    __syncthreads();
    lap(3);
    // ---- E: the winner's column for the requests still queued
    if (sh.win_v >= 0) {
      const int vs = pv.inv[sh.win_v];
      for (int r2 = r + 1 + tid; r2 < n_new; r2 += QUEUE_THREADS) {
        bool reach;
        const double cm = pair_cmin(pv, dv, new_rows[r2], vs, &reach);
        uint32_t* w = mask + (size_t)r2 * words + (vs >> 5);
        if (reach) { cmin[(int64_t)r2 * nv + vs] = cm; atomicOr(w, 1u << (vs & 31)); }
        else atomicAnd(w, ~(1u << (vs & 31)));
      }
    }
    __syncthreads();
    lap(4);
  }
  if (tid == 0 && prof)
    for (int i = 0; i < 7; ++i) prof[i] = pacc[i];
}

}  // namespace

// ====================================================================================================== host side
struct LegacyDev {
  int n_veh = 0, n_rows = 0;
  DrsArena veh, tab, res;            // grow-only device allocations: vehicles + plans, request table, per-evaluation buffers
  PlanView pv{};
  std::vector<int32_t> h_perm, h_len;
  // per-evaluation buffers (slices of `res`)
  int cap_new = 0, words = 0, list_cap = 0;
  int32_t *d_rows = nullptr, *d_bv = nullptr, *d_bi = nullptr, *d_bj = nullptr;
  double *d_dc = nullptr, *d_cmin = nullptr;
  uint32_t* d_mask = nullptr;
  int2* d_list = nullptr;
  int* d_counters = nullptr;         // [0] list fill, [1] next request, [2] commit error, [3] fold events, [4] last vehicle that set the new request's Cld
  FoldEvent* d_events = nullptr;
  unsigned long long* d_prof = nullptr;
  double* d_fx_new = nullptr;        // [vs] Cld the vehicle's scan left on the new request
  int64_t base_candidates = 0;       // sum over vehicles of (L+1)(L+2)/2 for the plans as they are now
  bool prepared = false, full_list = false, list_is_full = false;
  double detail_ms[6] = {0, 0, 0, 0, 0, 0};   // last call: plan prep, reach test, evaluation, fold, queue loop; [5] = surviving pairs
  std::vector<int32_t> h_rows;       // the new-request rows of the running call
};

void drs_tick_free_plans(drs_tick* t) {
  LegacyDev* d = t->legacy;
  if (!d) return;
  d->veh.release(); d->tab.release(); d->res.release();
  delete d;
  t->legacy = nullptr;
}

namespace {

int64_t candidates_of(int L) { return (int64_t)(L + 1) * (L + 2) / 2; }

// (re)allocates the per-evaluation buffers for n_new requests; pointers stay valid while n_new does not grow
int ensure_eval_buffers(LegacyDev& d, int n_new) {
  const int nv = std::max(d.n_veh, 1);
  const int words = (nv + 31) / 32;
  if (n_new <= d.cap_new && words == d.words && d.d_rows && d.full_list == d.list_is_full) return DRS_OK;
  const int cap = std::max(n_new, d.cap_new);
  const int64_t pairs = (int64_t)cap * nv;
  // survivors of the reach test are ~1 % of the pairs; the list starts at a quarter and grows to all pairs if it ever overflows
  const int list_cap = (int)(d.full_list ? pairs : std::min<int64_t>(pairs, std::max<int64_t>(65536, pairs / 4)));
  d.list_is_full = d.full_list;
  DrsArenaLayout lay;
  const size_t o_rows = lay.add(sizeof(int32_t) * cap), o_bv = lay.add(sizeof(int32_t) * cap), o_bi = lay.add(sizeof(int32_t) * cap),
               o_bj = lay.add(sizeof(int32_t) * cap), o_dc = lay.add(sizeof(double) * cap), o_cnt = lay.add(sizeof(int) * 8), o_prof = lay.add(64), o_ev = lay.add(sizeof(FoldEvent) * FOLD_EVENT_CAP),
               o_fx = lay.add(sizeof(double) * nv),
               o_mask = lay.add(sizeof(uint32_t) * (size_t)cap * words), o_list = lay.add(sizeof(int2) * (size_t)list_cap),
               o_cmin = lay.add(sizeof(double) * (size_t)pairs);
  int rc = d.res.reserve(lay.total);
  if (rc) return rc;
  unsigned char* b = (unsigned char*)d.res.ptr;
  d.d_rows = (int32_t*)(b + o_rows); d.d_bv = (int32_t*)(b + o_bv); d.d_bi = (int32_t*)(b + o_bi); d.d_bj = (int32_t*)(b + o_bj);
  d.d_dc = (double*)(b + o_dc); d.d_counters = (int*)(b + o_cnt); d.d_mask = (uint32_t*)(b + o_mask); d.d_list = (int2*)(b + o_list);
  d.d_prof = (unsigned long long*)(b + o_prof);
  d.d_cmin = (double*)(b + o_cmin); d.d_events = (FoldEvent*)(b + o_ev); d.d_fx_new = (double*)(b + o_fx);
  d.cap_new = cap; d.words = words; d.list_cap = list_cap;
  return DRS_OK;
}

// reach test + evaluation of every (request, vehicle) pair against the plans as they are now
int scan_all(drs_tick* t, LegacyDev& d, int n_new, cudaStream_t st) {
  const drs_graph* g = t->g;
  DistView dv{g->d_dist, g->ld, g->mode};
  DrsTimer tm_prep(st), tm_reach(st), tm_eval(st);
  tm_prep.start();
  if (!d.prepared) {
    plan_prep_kernel<<<(d.n_veh + 127) / 128, 128, 0, st>>>(d.pv, dv);
    drs_count_launch();
    d.prepared = true;
  }
  tm_prep.stop();
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
  const size_t row_bytes = (size_t)g->ld * 4;
  const char* env = getenv("DRS_INS_STAGED");
  const bool staged = !g->directed && row_bytes <= 220 * 1024 && row_bytes % 16 == 0 && !(env && env[0] == '0');
  for (int attempt = 0; attempt < 2; ++attempt) {
    DRS_CUDA(cudaMemsetAsync(d.d_counters, 0, sizeof(int) * 8, st));
    tm_reach.start();
    const int grid = std::max(1, std::min(n_new, sms));
    if (staged && g->mode == DRS_MODE_I32) {
      DRS_CUDA(cudaFuncSetAttribute(insertion_reach_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
      insertion_reach_kernel<true, true><<<grid, 1024, row_bytes, st>>>(d.pv, dv, n_new, d.d_rows, d.d_mask, d.words, d.d_list, d.list_cap, d.d_counters);
    } else if (staged) {
      DRS_CUDA(cudaFuncSetAttribute(insertion_reach_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
      insertion_reach_kernel<true, false><<<grid, 1024, row_bytes, st>>>(d.pv, dv, n_new, d.d_rows, d.d_mask, d.words, d.d_list, d.list_cap, d.d_counters);
    } else {
      insertion_reach_kernel<false, false><<<grid, 1024, 0, st>>>(d.pv, dv, n_new, d.d_rows, d.d_mask, d.words, d.d_list, d.list_cap, d.d_counters);
    }
    drs_count_launch();
    tm_reach.stop();
    DRS_CUDA(cudaGetLastError());
    const int64_t pairs = (int64_t)n_new * d.n_veh;
    if (d.list_cap >= pairs) break;                    // the list cannot overflow
    int fill = 0;
    DRS_CUDA(cudaMemcpyAsync(&fill, d.d_counters, sizeof(int), cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaStreamSynchronize(st));
    if (fill <= d.list_cap) break;
    // more survivors than the list holds (dense demand around the fleet): size it for every pair and redo the reach test
    d.full_list = true;
    int rc = ensure_eval_buffers(d, n_new);
    if (rc) return rc;
    DRS_CUDA(cudaMemcpyAsync(d.d_rows, d.h_rows.data(), sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, st));
  }
  const int eval_blocks = std::max(1, std::min(sms * 8, (d.list_cap + 127) / 128));
  tm_eval.start();
  insertion_eval_kernel<<<eval_blocks, 128, 0, st>>>(d.pv, dv, d.d_rows, d.d_list, d.d_counters, d.list_cap, d.d_cmin);
  drs_count_launch();
  tm_eval.stop();
  DRS_CUDA(cudaGetLastError());
  int fill = 0;
  DRS_CUDA(cudaMemcpyAsync(&fill, d.d_counters, sizeof(int), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  d.detail_ms[0] = tm_prep.ms(); d.detail_ms[1] = tm_reach.ms(); d.detail_ms[2] = tm_eval.ms(); d.detail_ms[5] = (double)fill;
  return DRS_OK;
}

}  // namespace

extern "C" int drs_tick_set_plans(drs_tick* t, const drs_plan_batch* b, const drs_request_table* q,
                                  const drs_insertion_params* prm) {
  DRS_REQUIRE(t && b && q, DRS_ERR_ARG, "drs_tick_set_plans: null argument");
  const int nv = b->n_veh, nq = q->n_rows;
  DRS_REQUIRE(nv >= 0 && nq >= 0, DRS_ERR_ARG, "drs_tick_set_plans: negative size");
  DRS_REQUIRE(nv == 0 || (b->start_node && b->start_delay && b->onboard && b->clock && b->capacity && b->cost && b->stop_off),
              DRS_ERR_ARG, "drs_tick_set_plans: null vehicle array");
  DRS_REQUIRE(nq == 0 || (q->o_node && q->d_node && q->Cep && q->Clp && q->Cld && q->Ts && q->Tr && q->pwt), DRS_ERR_ARG,
              "drs_tick_set_plans: null request-table array");
  const int n = t->g->n;
  for (int v = 0; v < nv; ++v) {
    DRS_REQUIRE(b->start_node[v] >= 0 && b->start_node[v] < n, DRS_ERR_RANGE, "drs_tick_set_plans: vehicle %d node out of range", v);
    const int cnt = b->stop_off[v + 1] - b->stop_off[v];
    DRS_REQUIRE(cnt >= 0 && cnt <= INS_MAXL, DRS_ERR_ARG, "drs_tick_set_plans: vehicle %d has %d planned stops (max %d)", v, cnt, INS_MAXL);
    for (int s = b->stop_off[v]; s < b->stop_off[v + 1]; ++s) {
      DRS_REQUIRE(b->stop_node[s] >= 0 && b->stop_node[s] < n, DRS_ERR_RANGE, "drs_tick_set_plans: stop %d node out of range", s);
      DRS_REQUIRE(b->stop_req[s] >= 0 && b->stop_req[s] < nq, DRS_ERR_RANGE, "drs_tick_set_plans: stop %d request row out of range", s);
    }
    if (b->keep_node) DRS_REQUIRE(b->keep_node[v] >= -1 && b->keep_node[v] < n, DRS_ERR_RANGE, "drs_tick_set_plans: vehicle %d kept-step node out of range", v);
    if (b->pos_node) DRS_REQUIRE(b->pos_node[v] >= -1 && b->pos_node[v] < n, DRS_ERR_RANGE, "drs_tick_set_plans: vehicle %d position node out of range", v);
  }
  for (int r = 0; r < nq; ++r)
    DRS_REQUIRE(q->o_node[r] >= 0 && q->o_node[r] < n && q->d_node[r] >= 0 && q->d_node[r] < n, DRS_ERR_RANGE,
                "drs_tick_set_plans: request row %d node out of range", r);
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  if (!t->legacy) t->legacy = new LegacyDev();
  LegacyDev& d = *t->legacy;
  // vehicles sorted by plan length, then by start vertex (neighbouring lanes run the same trip counts and gather
  // neighbouring columns: vertex ids are spatially coherent under the host layer's locality numbering)
  const int nvp = std::max(nv, 1);
  d.h_perm.resize(nvp);
  for (int v = 0; v < nv; ++v) d.h_perm[v] = v;
  std::stable_sort(d.h_perm.begin(), d.h_perm.begin() + nv, [&](int x, int y) {
    const int lx = b->stop_off[x + 1] - b->stop_off[x], ly = b->stop_off[y + 1] - b->stop_off[y];
    return lx != ly ? lx < ly : b->start_node[x] < b->start_node[y];
  });
  // ---- one host image of the vehicle-side arrays, one copy
  DrsArenaLayout lv;
  const size_t slots = (size_t)INS_MAXL * nvp, slots1 = (size_t)(INS_MAXL + 1) * nvp;
  const size_t o_perm = lv.add(4 * nvp), o_inv = lv.add(4 * nvp), o_node = lv.add(4 * nvp), o_onb = lv.add(4 * nvp), o_cap = lv.add(4 * nvp),
               o_len = lv.add(4 * nvp), o_delay = lv.add(8 * nvp), o_clock = lv.add(8 * nvp), o_cost = lv.add(8 * nvp), o_pos = lv.add(4 * nvp),
               o_kn = lv.add(4 * nvp), o_kp = lv.add(4 * nvp), o_kt = lv.add(8 * nvp), o_ke = lv.add(8 * nvp), o_mnode = lv.add(4 * slots), o_mreq = lv.add(4 * slots),
               o_mpod = lv.add(slots), o_mpar = lv.add(slots);
  const size_t upload_bytes = lv.total;
  const size_t o_seg = lv.add(8 * slots), o_bt = lv.add(8 * slots1), o_bc = lv.add(8 * slots1), o_bcld = lv.add(8 * slots), o_bv = lv.add(4 * nvp),
               o_lb = lv.add(4 * slots1), o_ns = lv.add(nvp);
  int rc = d.veh.reserve(lv.total);
  if (rc) return rc;
  std::vector<unsigned char>& img = d.veh.host;
  img.assign(upload_bytes, 0);
  auto I32 = [&](size_t o) { return (int32_t*)(img.data() + o); };
  auto F64 = [&](size_t o) { return (double*)(img.data() + o); };
  auto I8 = [&](size_t o) { return (int8_t*)(img.data() + o); };
  d.h_len.assign(nvp, 0);
  d.base_candidates = 0;
  for (int vs = 0; vs < nv; ++vs) {
    const int v = d.h_perm[vs], s0 = b->stop_off[v], L = b->stop_off[v + 1] - s0;
    I32(o_perm)[vs] = v; I32(o_inv)[v] = vs;
    I32(o_node)[vs] = b->start_node[v]; I32(o_onb)[vs] = b->onboard[v]; I32(o_cap)[vs] = b->capacity[v]; I32(o_len)[vs] = L;
    F64(o_delay)[vs] = b->start_delay[v]; F64(o_clock)[vs] = b->clock[v]; F64(o_cost)[vs] = b->cost[v];
    // defaults: a vehicle with no delay stands on its start vertex; one under way keeps the step it is on
    const int pos = b->pos_node ? b->pos_node[v] : (b->start_delay[v] == 0.0 ? b->start_node[v] : -1);
    int kn = b->keep_node ? b->keep_node[v] : (b->start_delay[v] != 0.0 || L > 0 ? b->start_node[v] : -1);
    I32(o_pos)[vs] = pos; I32(o_kn)[vs] = kn; I32(o_kp)[vs] = -1;
    F64(o_kt)[vs] = b->keep_t ? b->keep_t[v] : b->start_delay[v];
    F64(o_ke)[vs] = b->keep_et ? b->keep_et[v] : F64(o_kt)[vs];
    d.h_len[vs] = L;
    d.base_candidates += candidates_of(L);
    for (int k = 0; k < L; ++k) {
      const size_t at = (size_t)k * nv + vs;
      I32(o_mnode)[at] = b->stop_node[s0 + k]; I32(o_mreq)[at] = b->stop_req[s0 + k]; I8(o_mpod)[at] = b->stop_pod[s0 + k];
      int8_t partner = -1;
      if (b->stop_pod[s0 + k] < 0)
        for (int m = 0; m < k; ++m)
          if (b->stop_req[s0 + m] == b->stop_req[s0 + k] && b->stop_pod[s0 + m] > 0) partner = (int8_t)m;
      I8(o_mpar)[at] = partner;
    }
  }
  DRS_CUDA(cudaMemcpyAsync(d.veh.ptr, img.data(), upload_bytes, cudaMemcpyHostToDevice, st));
  // ---- request table
  DrsArenaLayout lt;
  const int nqp = std::max(nq, 1);
  const size_t t_o = lt.add(4 * nqp), t_d = lt.add(4 * nqp), t_cep = lt.add(8 * nqp), t_clp = lt.add(8 * nqp), t_cld = lt.add(8 * nqp),
               t_ts = lt.add(8 * nqp), t_tr = lt.add(8 * nqp), t_pwt = lt.add(8 * nqp);
  if ((rc = d.tab.reserve(lt.total))) return rc;
  std::vector<unsigned char>& timg = d.tab.host;
  timg.assign(lt.total, 0);
  if (nq) {
    std::memcpy(timg.data() + t_o, q->o_node, 4 * (size_t)nq); std::memcpy(timg.data() + t_d, q->d_node, 4 * (size_t)nq);
    std::memcpy(timg.data() + t_cep, q->Cep, 8 * (size_t)nq); std::memcpy(timg.data() + t_clp, q->Clp, 8 * (size_t)nq);
    std::memcpy(timg.data() + t_cld, q->Cld, 8 * (size_t)nq); std::memcpy(timg.data() + t_ts, q->Ts, 8 * (size_t)nq);
    std::memcpy(timg.data() + t_tr, q->Tr, 8 * (size_t)nq); std::memcpy(timg.data() + t_pwt, q->pwt, 8 * (size_t)nq);
  }
  DRS_CUDA(cudaMemcpyAsync(d.tab.ptr, timg.data(), lt.total, cudaMemcpyHostToDevice, st));
  unsigned char* vb = (unsigned char*)d.veh.ptr;
  unsigned char* tb = (unsigned char*)d.tab.ptr;
  PlanView& pv = d.pv;
  pv.n_veh = nv;
  pv.perm = (int32_t*)(vb + o_perm); pv.inv = (int32_t*)(vb + o_inv); pv.node = (int32_t*)(vb + o_node); pv.onboard = (int32_t*)(vb + o_onb);
  pv.cap = (int32_t*)(vb + o_cap); pv.len = (int32_t*)(vb + o_len); pv.delay = (double*)(vb + o_delay); pv.clock = (double*)(vb + o_clock);
  pv.cost = (double*)(vb + o_cost); pv.pos_node = (int32_t*)(vb + o_pos); pv.keep_node = (int32_t*)(vb + o_kn); pv.keep_pend = (int32_t*)(vb + o_kp); pv.keep_t = (double*)(vb + o_kt);
  pv.keep_et = (double*)(vb + o_ke); pv.m_node = (int32_t*)(vb + o_mnode); pv.m_req = (int32_t*)(vb + o_mreq); pv.m_pod = (int8_t*)(vb + o_mpod);
  pv.m_partner = (int8_t*)(vb + o_mpar); pv.m_seg = (double*)(vb + o_seg); pv.m_base_t = (double*)(vb + o_bt); pv.m_base_c = (double*)(vb + o_bc);
  pv.m_base_cld = (double*)(vb + o_bcld); pv.m_base_viol = (int32_t*)(vb + o_bv); pv.m_lb = (float*)(vb + o_lb); pv.m_nslots = (uint8_t*)(vb + o_ns);
  pv.q_o = (int32_t*)(tb + t_o); pv.q_d = (int32_t*)(tb + t_d); pv.q_cep = (double*)(tb + t_cep); pv.q_clp = (double*)(tb + t_clp);
  pv.q_cld = (double*)(tb + t_cld); pv.q_ts = (double*)(tb + t_ts); pv.q_tr = (double*)(tb + t_tr); pv.q_pwt = (double*)(tb + t_pwt);
  pv.prm = prm ? *prm : drs_insertion_params{1.5, 1.0, 1.5, 3.0};
  pv.directed = t->g->directed;
  DRS_CUDA(cudaStreamSynchronize(st));
  d.n_veh = nv; d.n_rows = nq;
  d.prepared = false;
  return DRS_OK;
}

namespace {

int check_new_rows(const char* who, drs_tick* t, int32_t n_new, const int32_t* new_rows, int32_t* best_veh, int32_t* best_i, int32_t* best_j,
                   double* best_dc) {
  DRS_REQUIRE(t && t->legacy, DRS_ERR_STATE, "%s: call drs_tick_set_plans first", who);
  DRS_REQUIRE(t->g->apsp_valid, DRS_ERR_STATE, "%s: travel-time matrix not built (or invalidated)", who);
  DRS_REQUIRE(n_new >= 0 && (n_new == 0 || (new_rows && best_veh && best_i && best_j && best_dc)), DRS_ERR_ARG, "%s: bad argument", who);
  LegacyDev& d = *t->legacy;
  for (int r = 0; r < n_new; ++r)
    DRS_REQUIRE(new_rows[r] >= 0 && new_rows[r] < d.n_rows, DRS_ERR_RANGE, "%s: request row %d out of range", who, new_rows[r]);
  DRS_REQUIRE((int64_t)d.n_veh * n_new < ((int64_t)1 << 31), DRS_ERR_ARG, "%s: %lld pairs exceed 2^31", who, (long long)d.n_veh * n_new);
  return DRS_OK;
}

int fetch_results(LegacyDev& d, int n_new, int32_t* best_veh, int32_t* best_i, int32_t* best_j, double* best_dc, cudaStream_t st) {
  DRS_CUDA(cudaMemcpyAsync(best_veh, d.d_bv, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(best_i, d.d_bi, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(best_j, d.d_bj, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaMemcpyAsync(best_dc, d.d_dc, sizeof(double) * n_new, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  return DRS_OK;
}

}  // namespace

extern "C" int drs_tick_insertion_eval(drs_tick* t, int32_t n_new, const int32_t* new_rows, int32_t* best_veh,
                                       int32_t* best_i, int32_t* best_j, double* best_dc, int64_t* n_candidates) {
  int rc = check_new_rows("drs_tick_insertion_eval", t, n_new, new_rows, best_veh, best_i, best_j, best_dc);
  if (rc) return rc;
  if (n_candidates) *n_candidates = 0;
  if (n_new == 0) return DRS_OK;
  LegacyDev& d = *t->legacy;
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  if ((rc = ensure_eval_buffers(d, n_new))) return rc;
  d.h_rows.assign(new_rows, new_rows + n_new);
  DRS_CUDA(cudaMemcpyAsync(d.d_rows, d.h_rows.data(), sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, st));
  DistView dv{t->g->d_dist, t->g->ld, t->g->mode};
  DrsTimer tm_scan(st), tm_fold(st);
  tm_scan.start();
  if (d.n_veh > 0 && (rc = scan_all(t, d, n_new, st))) return rc;
  tm_scan.stop();
  tm_fold.start();
  int* d_novf = d.d_counters + 5;
  DRS_CUDA(cudaMemsetAsync(d_novf, 0, sizeof(int), st));
  insertion_fold_kernel<<<n_new, FOLD_THREADS, 0, st>>>(d.pv, dv, n_new, d.d_rows, d.d_mask, d.words, d.d_cmin, d.d_bv, d.d_bi, d.d_bj, d.d_dc, d_novf);
  drs_count_launch();
  tm_fold.stop();
  DRS_CUDA(cudaGetLastError());
  if ((rc = fetch_results(d, n_new, best_veh, best_i, best_j, best_dc, st))) return rc;
  int novf = 0;
  DRS_CUDA(cudaMemcpy(&novf, d_novf, sizeof(int), cudaMemcpyDeviceToHost));
  if (novf > 0) {   // requests with more than FOLD_MAXC vehicles in reach: the dense walk over the whole fleet
    for (int r = 0; r < n_new; ++r)
      if (best_veh[r] == -2) {
        insertion_fold_dense_kernel<<<1, 32, 0, st>>>(d.pv, dv, r, 1, d.d_rows, d.d_mask, d.words, d.d_cmin, d.d_bv, d.d_bi, d.d_bj, d.d_dc,
                                                      nullptr, nullptr);
        drs_count_launch();
      }
    DRS_CUDA(cudaGetLastError());
    if ((rc = fetch_results(d, n_new, best_veh, best_i, best_j, best_dc, st))) return rc;
  }
  t->last_ms[2] = tm_scan.ms(); t->last_ms[3] = tm_fold.ms();
  d.detail_ms[3] = t->last_ms[3]; d.detail_ms[4] = 0;
  if (n_candidates) *n_candidates = d.base_candidates * n_new;
  return DRS_OK;
}

// Sequential queue (lib/Agents.py:1429-1448): the requests are inserted one after another, each scan seeing the routes,
// costs and predicted waiting times the previous winners' build_route left behind.
extern "C" int drs_tick_insertion_queue(drs_tick* t, int32_t n_new, const int32_t* new_rows, double T_now, int32_t flags, int32_t* best_veh,
                                        int32_t* best_i, int32_t* best_j, double* best_dc, int64_t* n_candidates) {
  int rc = check_new_rows("drs_tick_insertion_queue", t, n_new, new_rows, best_veh, best_i, best_j, best_dc);
  if (rc) return rc;
  if (n_candidates) *n_candidates = 0;
  if (n_new == 0) return DRS_OK;
  LegacyDev& d = *t->legacy;
  drs_graph* g = t->g;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  if ((rc = ensure_eval_buffers(d, n_new))) return rc;
  d.h_rows.assign(new_rows, new_rows + n_new);
  DRS_CUDA(cudaMemcpyAsync(d.d_rows, d.h_rows.data(), sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, st));
  DistView dv{g->d_dist, g->ld, g->mode};
  PredSource ps = drs_pred_source(g);
  DrsTimer tm_scan(st), tm_queue(st);
  tm_scan.start();
  if (d.n_veh > 0 && (rc = scan_all(t, d, n_new, st))) return rc;
  tm_scan.stop();
  tm_queue.start();
  int* d_err = d.d_counters + 2;
  const int rescan_blocks = std::max(1, std::min(148, (n_new + 127) / 128));
  const bool track_cld = (flags & DRS_QUEUE_TRACK_CLD) != 0;
  const char* unfused_env = getenv("DRS_INS_QUEUE_UNFUSED");
  bool fused = !(unfused_env && unfused_env[0] == '1') && d.n_veh > 0;
  if (fused) {
    // one persistent CTA walks the whole queue (insertion_queue_kernel)
    DRS_CUDA(cudaFuncSetAttribute(insertion_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QueueShared)));
    insertion_queue_kernel<<<1, QUEUE_THREADS, sizeof(QueueShared), st>>>(d.pv, dv, ps, g->d_esrc, g->d_edst, g->d_tt64, n_new, d.d_rows, d.d_mask,
                                                                           d.words, d.d_cmin, d.d_bv, d.d_bi, d.d_bj, d.d_dc, T_now, track_cld ? 1 : 0,
                                                                           d_err, d.d_prof);
    drs_count_launch();
    DRS_CUDA(cudaGetLastError());
    int err5 = 0;
    DRS_CUDA(cudaMemcpyAsync(&err5, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaStreamSynchronize(st));
    if (getenv("DRS_INS_QUEUE_PROF")) {
      unsigned long long pr[8];
      cudaMemcpy(pr, d.d_prof, sizeof(pr), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[queue prof] %d requests: cycles/request list %.0f fold %.0f sidefx %.0f commit %.0f rescan %.0f | listed vehicles/request %.1f "
              "re-scans/request %.2f\n", n_new, (double)pr[0] / n_new, (double)pr[1] / n_new, (double)pr[2] / n_new, (double)pr[3] / n_new,
              (double)pr[4] / n_new, (double)pr[5] / n_new, (double)pr[6] / n_new);
    }
    DRS_REQUIRE(err5 != 5, DRS_ERR_CAPACITY, "drs_tick_insertion_queue: more than %d vehicles in reach of one request; "
                "set DRS_INS_QUEUE_UNFUSED=1 for the kernel-per-step form", QUEUE_MAXC);
  }
  int* d_nev = d.d_counters + 3;
  int* d_last_v = d.d_counters + 4;
  if (track_cld) {
    const int minus_one = -1;
    DRS_CUDA(cudaMemcpyAsync(d_last_v, &minus_one, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  for (int r = 0; r < n_new && !fused; ++r) {
    insertion_fold_dense_kernel<<<1, 32, 0, st>>>(d.pv, dv, r, 1, d.d_rows, d.d_mask, d.words, d.d_cmin, d.d_bv, d.d_bi, d.d_bj, d.d_dc,
                                                  track_cld ? d.d_events : nullptr, d_nev);
    if (track_cld)
      insertion_sidefx_kernel<<<(d.n_veh + 127) / 128, 128, 0, st>>>(d.pv, dv, d.d_rows, r, d.d_mask, d.words, d.d_events, d_nev, d.d_fx_new,
                                                                      d_last_v, d_err);
    insertion_commit_kernel<<<1, 32, 0, st>>>(d.pv, dv, ps, g->d_esrc, g->d_edst, g->d_tt64, d.d_rows, r, d.d_bv, d.d_bi, d.d_bj, T_now, d_err,
                                              track_cld ? d.d_fx_new : nullptr, d_last_v);
    if (r + 1 < n_new)
      insertion_rescan_kernel<<<rescan_blocks, 128, 0, st>>>(d.pv, dv, n_new, d.d_rows, r, d.d_bv, d.d_mask, d.words, d.d_cmin);
    drs_count_launch(2 + (track_cld ? 1 : 0) + (r + 1 < n_new ? 1 : 0));
  }
  tm_queue.stop();
  DRS_CUDA(cudaGetLastError());
  int err = 0;
  DRS_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = fetch_results(d, n_new, best_veh, best_i, best_j, best_dc, st))) return rc;
  t->last_ms[2] = tm_scan.ms(); t->last_ms[3] = tm_queue.ms();
  d.detail_ms[3] = 0; d.detail_ms[4] = t->last_ms[3];
  DRS_REQUIRE(err == 0, DRS_ERR_STATE, "drs_tick_insertion_queue: a winner's plan could not be patched (code %d)", err);
  if (n_candidates) {
    // every request scans every vehicle's plan as it is at its turn;
    // lengths of the winners grow by two per commit
    std::vector<int> len(d.h_len);
    std::vector<int32_t> inv(d.n_veh);
    for (int vs = 0; vs < d.n_veh; ++vs) inv[d.h_perm[vs]] = vs;
    int64_t per_req = d.base_candidates, total = 0;
    for (int r = 0; r < n_new; ++r) {
      total += per_req;
      if (best_veh[r] >= 0) {
        int& L = len[inv[best_veh[r]]];
        per_req += candidates_of(L + 2) - candidates_of(L);
        L += 2;
      }
    }
    for (int vs = 0; vs < d.n_veh; ++vs) d.h_len[vs] = len[vs];
    d.base_candidates = per_req;
    *n_candidates = total;
  }
  return DRS_OK;
}

// One commit by hand (the building block of the queue): request row `req_row` joins vehicle `veh` with its pickup at
// route position i and its drop-off at j (lib/Agents.py:1468-1469, then Veh.build_route).
extern "C" int drs_tick_commit_insertion(drs_tick* t, int32_t req_row, int32_t veh, int32_t i, int32_t j, double T_now) {
  DRS_REQUIRE(t && t->legacy, DRS_ERR_STATE, "drs_tick_commit_insertion: call drs_tick_set_plans first");
  LegacyDev& d = *t->legacy;
  DRS_REQUIRE(t->g->apsp_valid, DRS_ERR_STATE, "drs_tick_commit_insertion: travel-time matrix not built (or invalidated)");
  DRS_REQUIRE(req_row >= 0 && req_row < d.n_rows && veh >= 0 && veh < d.n_veh, DRS_ERR_RANGE, "drs_tick_commit_insertion: row or vehicle out of range");
  drs_graph* g = t->g;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  int rc;
  if ((rc = ensure_eval_buffers(d, 1))) return rc;
  DistView dv{g->d_dist, g->ld, g->mode};
  if (!d.prepared) {
    plan_prep_kernel<<<(d.n_veh + 127) / 128, 128, 0, st>>>(d.pv, dv);
    drs_count_launch();
    d.prepared = true;
  }
  int* d_err = d.d_counters + 2;
  DRS_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  insertion_commit_explicit_kernel<<<1, 32, 0, st>>>(d.pv, dv, drs_pred_source(g), g->d_esrc, g->d_edst, g->d_tt64, req_row, veh, i, j, T_now, d_err);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  int err = 0;
  DRS_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  DRS_REQUIRE(err == 0, DRS_ERR_ARG, "drs_tick_commit_insertion: the plan of vehicle %d cannot take (i=%d, j=%d) (code %d)", veh, i, j, err);
  int inv = 0;
  for (int vs = 0; vs < d.n_veh; ++vs) if (d.h_perm[vs] == veh) inv = vs;
  d.base_candidates += candidates_of(d.h_len[inv] + 2) - candidates_of(d.h_len[inv]);
  d.h_len[inv] += 2;
  return DRS_OK;
}

// State of the plans after commits: per vehicle its cost and plan (request row, pod per stop), per request row Cld and pwt.
extern "C" int drs_tick_fetch_plans(drs_tick* t, double* veh_cost, int32_t* plan_len, int32_t* stop_req, int8_t* stop_pod,
                                    int32_t stops_per_vehicle, double* row_Cld, double* row_pwt) {
  DRS_REQUIRE(t && t->legacy, DRS_ERR_STATE, "drs_tick_fetch_plans: call drs_tick_set_plans first");
  LegacyDev& d = *t->legacy;
  DRS_REQUIRE(!stop_req || stops_per_vehicle >= INS_MAXL, DRS_ERR_CAPACITY, "drs_tick_fetch_plans: stops_per_vehicle must be >= %d", INS_MAXL);
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  const int nv = d.n_veh;
  std::vector<double> cost(std::max(nv, 1));
  std::vector<int32_t> len(std::max(nv, 1)), mreq((size_t)INS_MAXL * std::max(nv, 1));
  std::vector<int8_t> mpod((size_t)INS_MAXL * std::max(nv, 1));
  if (nv) {
    DRS_CUDA(cudaMemcpyAsync(cost.data(), d.pv.cost, sizeof(double) * nv, cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaMemcpyAsync(len.data(), d.pv.len, sizeof(int32_t) * nv, cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaMemcpyAsync(mreq.data(), d.pv.m_req, sizeof(int32_t) * (size_t)INS_MAXL * nv, cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaMemcpyAsync(mpod.data(), d.pv.m_pod, (size_t)INS_MAXL * nv, cudaMemcpyDeviceToHost, st));
  }
  if (row_Cld && d.n_rows) DRS_CUDA(cudaMemcpyAsync(row_Cld, d.pv.q_cld, sizeof(double) * d.n_rows, cudaMemcpyDeviceToHost, st));
  if (row_pwt && d.n_rows) DRS_CUDA(cudaMemcpyAsync(row_pwt, d.pv.q_pwt, sizeof(double) * d.n_rows, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  for (int vs = 0; vs < nv; ++vs) {
    const int v = d.h_perm[vs];
    if (veh_cost) veh_cost[v] = cost[vs];
    if (plan_len) plan_len[v] = len[vs];
    if (stop_req)
      for (int k = 0; k < len[vs]; ++k) {
        stop_req[(size_t)v * stops_per_vehicle + k] = mreq[(size_t)k * nv + vs];
        if (stop_pod) stop_pod[(size_t)v * stops_per_vehicle + k] = mpod[(size_t)k * nv + vs];
      }
  }
  return DRS_OK;
}

// Device times of the last insertion call by kernel (ms): [0] plan prep, [1] reach test, [2] evaluation of the surviving
// pairs, [3] fold (drs_tick_insertion_eval), [4] fold / commit / re-scan loop (drs_tick_insertion_queue); [5] = number of
// (request, vehicle) pairs that survived the reach test.
extern "C" int drs_tick_last_insertion_ms(const drs_tick* t, double* ms) {
  DRS_REQUIRE(t && t->legacy && ms, DRS_ERR_ARG, "drs_tick_last_insertion_ms: no insertion call on this tick yet");
  for (int i = 0; i < 6; ++i) ms[i] = t->legacy->detail_ms[i];
  return DRS_OK;
}
