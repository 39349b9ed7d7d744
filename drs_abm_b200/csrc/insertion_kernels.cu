// (vehicle, pickup slot, drop-off slot) insertion evaluation and argmin.
//
// Replaces Model.insert_heuristics / test_constraints_get_cost
// (lib/Agents.py:1451-1572).  Every candidate route is walked in fp64 with the
// reference's operation order, so costs are bit-identical for integer-weighted maps;
// travel times are gathered from the resident matrix.
//
//   plan_prep_kernel        per vehicle: plan segment times and the walk of the
//       unmodified plan (the common prefix of every candidate), slot-major.
//   insertion_scan_kernel   one thread per (request, vehicle) pair, vehicles fastest:
//       every (i, j) candidate is evaluated WITHOUT the incumbent bound and the pair's
//       smallest successful total cost is written out (+inf if none).
//   insertion_fold_kernel   one warp per request walks the vehicles in the
//       reference's order.  A vehicle can only change the incumbent if its
//       smallest cost is <= the bound C = veh.c + dc (lib/Agents.py:1470); only
//       then is the pair re-scanned exactly as the reference does it -- with the
//       bound, "last equal-or-better wins" and the viol-code loop breaks
//       (load_pair / eval_candidate below are that literal restatement).
#include <algorithm>
#include <cstring>

#include "drs_tick.cuh"

namespace {

constexpr int MAXL = DRS_MAX_STOPS - 2;   // planned stops per vehicle

struct PlanView {
  int n_veh;
  const int32_t *node, *onboard, *cap, *off, *s_node, *s_req;
  const double *delay, *clock, *cost;
  const int8_t* s_pod;
  // request table
  const int32_t *q_o, *q_d;
  const double *q_cep, *q_clp, *q_cld, *q_ts, *q_tr, *q_pwt;
  drs_insertion_params prm;
  // per-vehicle precomputation (plan_prep_kernel): plan segment times and the walk of the
  // unmodified plan, which is the common prefix of every candidate
  // Slot-major copy for the scan kernel: vehicles sorted by plan length (vs = sorted index), element
  // [k * n_veh + vs] is plan stop k of that vehicle, so that a warp of consecutive vehicles loads
  // coalesced and runs the same trip counts.  perm[vs] = original vehicle index.
  const int32_t* perm;
  const int32_t* m_node;    // [k][vs] vertex of plan stop k
  const int32_t* m_req;     // [k][vs] request-table row
  const int8_t* m_pod;      // [k][vs]
  const int8_t* m_partner;  // [k][vs] plan index of the pickup belonging to a drop-off, -1 if on board
  // per-vehicle precomputation (plan_prep_kernel): plan segment times and the walk of the unmodified
  // plan, which is the common prefix of every candidate
  double* m_seg;            // [k][vs]  tt(previous stop or start vertex -> stop k)
  double* m_base_t;         // [k][vs]  t before plan stop k, k = 0..L
  double* m_base_c;         // [k][vs]  c before plan stop k
  double* m_base_cld;       // [k][vs]  Cld assigned when plan stop k is a pickup
  int32_t* m_base_viol;     // [vs]     first plan stop the unmodified walk fails at (L if none), -1: over capacity
  int directed;
};

// Everything one (vehicle, request) pair needs, in thread-local storage.
struct Pair {
  int L;                    // planned stops
  int n0, K;
  double T, t0, c0;
  int8_t pod[MAXL];
  int req[MAXL];
  double seg[MAXL];         // seg[k]  = tt(stop k-1 -> stop k), stop -1 = start vertex
  double to_o[MAXL + 1];    // to_o[k] = tt(stop k-1 -> origin)      (k = 0: from the start vertex)
  double to_d[MAXL + 1];    // to_d[k] = tt(stop k-1 -> destination)
  double from_o[MAXL];      // from_o[k] = tt(origin -> stop k)
  double from_d[MAXL];      // from_d[k] = tt(destination -> stop k)
  double od;                // tt(origin -> destination)
  int new_req;
};

__device__ void load_pair(Pair& p, const PlanView& pv, const DistView& dv, int v, int rrow) {
  const int s0 = pv.off[v];
  p.L = pv.off[v + 1] - s0;
  p.n0 = pv.onboard[v]; p.K = pv.cap[v];
  p.T = pv.clock[v]; p.t0 = pv.delay[v]; p.c0 = pv.cost[v];
  p.new_req = rrow;
  const int o = pv.q_o[rrow], d = pv.q_d[rrow];
  int prev = pv.node[v];
  p.to_o[0] = (prev == o) ? 0.0 : dv.at(prev, o);
  p.to_d[0] = (prev == d) ? 0.0 : dv.at(prev, d);
  for (int k = 0; k < p.L; ++k) {
    const int nd = pv.s_node[s0 + k];
    p.pod[k] = pv.s_pod[s0 + k];
    p.req[k] = pv.s_req[s0 + k];
    p.seg[k] = (prev == nd) ? 0.0 : dv.at(prev, nd);
    p.from_o[k] = (o == nd) ? 0.0 : dv.at(o, nd);
    p.from_d[k] = (d == nd) ? 0.0 : dv.at(d, nd);
    p.to_o[k + 1] = (nd == o) ? 0.0 : dv.at(nd, o);
    p.to_d[k + 1] = (nd == d) ? 0.0 : dv.at(nd, d);
    prev = nd;
  }
  p.od = (o == d) ? 0.0 : dv.at(o, d);
}

// test_constraints_get_cost (lib/Agents.py:1493-1572) for the candidate "pickup after a
// planned stops, drop-off after b planned stops" (a <= b), i.e. (i, j) = (a, b + 1).
// Returns viol (-1 = success) and the total cost through *c_out.
__device__ int eval_candidate(const Pair& p, const PlanView& pv, int a, int b, double C, double* c_out) {
  // capacity pre-check over the whole candidate route (:1518-1521)
  {
    int n = p.n0;
    for (int k = 0; k <= p.L; ++k) {
      if (k == a) { n += 1; if (n > p.K) return 1; }
      if (k == b) { n -= 1; }
      if (k < p.L) { n += p.pod[k]; if (n > p.K) return 1; }
    }
  }
  double cld_local[MAXL];            // Cld of requests picked up inside this candidate (:1550,1554)
  double cld_new = 0.0;
  double c = 0.0, t = 0.0 + p.t0;
  int n = p.n0;
  const double T = p.T;
  // sequence: planned stops 0..L-1 with the pickup before stop a and the drop-off before stop b
  // (after the pickup when a == b)
  int last = -1;                     // -1 start vertex, 0..L-1 planned stop, L pickup, L+1 drop-off
  for (int pos = 0; pos < p.L + 2; ++pos) {
    // which element comes next?
    int kind;                        // 0 planned stop, 1 new pickup, 2 new drop-off
    int k = 0;
    {
      // elements before planned stop index x: x planned + (x > a or (x == a ... )) handled by counting
      // position layout: [stops 0..a-1] P [stops a..b-1] D [stops b..L-1]
      if (pos < a) { kind = 0; k = pos; }
      else if (pos == a) { kind = 1; }
      else if (pos <= b) { kind = 0; k = pos - 1; }
      else if (pos == b + 1) { kind = 2; }
      else { kind = 0; k = pos - 2; }
    }
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
    } else if (pod == -1) {
      double cld;
      if (kind == 2) cld = cld_new;
      else {
        // drop-off of a planned request: Cld set by its pickup earlier in this walk, else the stored value
        cld = pv.q_cld[rq];
        for (int m = 0; m < k; ++m)
          if (p.req[m] == rq && p.pod[m] == 1) cld = cld_local[m];
      }
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
    last = (kind == 0) ? k : (kind == 1 ? p.L : p.L + 1);
  }
  *c_out = c;
  return -1;
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

// Per vehicle: plan segment times and the walk of the unmodified plan (slot-major outputs).
__global__ void plan_prep_kernel(PlanView pv, DistView dv) {
  const int vs = blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = pv.n_veh;
  if (vs >= nv) return;
  const int v = pv.perm[vs];
  const int L = pv.off[v + 1] - pv.off[v];
  int prev = pv.node[v];
  for (int k = 0; k < L; ++k) {
    const int nd = pv.m_node[k * nv + vs];
    pv.m_seg[k * nv + vs] = (prev == nd) ? 0.0 : dv.at(prev, nd);
    prev = nd;
  }
  int occ = pv.onboard[v], viol_at = L;
  bool over = occ > pv.cap[v];
  for (int k = 0; k < L; ++k) { occ += pv.m_pod[k * nv + vs]; if (occ > pv.cap[v]) over = true; }
  Walk w{0.0 + pv.delay[v], 0.0, pv.onboard[v]};
  const double T = pv.clock[v];
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
  pv.m_base_viol[vs] = over ? -1 : viol_at;
}

// Scan: one thread per (request, vehicle) pair, VEHICLES fastest and sorted by plan length.
//   * On an undirected map the matrix is symmetric, so every travel time the pair needs lives in the two
//     matrix rows of the request (origin row, destination row): all warps working on a request gather
//     from the same 2 x 4n bytes, which stay in L2 -- DRAM sees each request row once instead of one
//     32-byte sector per gathered element.
//   * The unmodified-plan prefix is shared (precomputed per vehicle), the state "pickup inserted after a
//     stops, b-a further stops served" is advanced incrementally over b, and only the tail after the
//     drop-off is re-walked per candidate.  Arithmetic and operation order per candidate are those of
//     eval_candidate; only successes matter here (the pair's smallest total cost), so no viol codes.
__global__ void __launch_bounds__(128) insertion_scan_kernel(PlanView pv, DistView dv, int n_new,
                                                             const int32_t* __restrict__ new_rows, double* __restrict__ cmin_out,
                                                             unsigned long long* n_cand) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = pv.n_veh;
  const int64_t total = (int64_t)nv * n_new;
  if (idx >= total) return;
  const int r = (int)(idx / nv), vs = (int)(idx % nv);
  const int v = pv.perm[vs];
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const int L = pv.off[v + 1] - pv.off[v];
  if (r == 0) atomicAdd(n_cand, (unsigned long long)((L + 1) * (L + 2) / 2) * (unsigned long long)n_new);
  double cmin = inf;
  const int bviol = pv.m_base_viol[vs];
  if (bviol >= 0) {
    const int rq = new_rows[r];
    const int o = pv.q_o[rq], d = pv.q_d[rq];
    const int K = pv.cap[v];
    const double T = pv.clock[v];
    double to_o[MAXL + 1], to_d[MAXL + 1], from_o[MAXL], from_d[MAXL], cld_cur[MAXL];
    // origin-side travel times first: most vehicles cannot reach the pickup in time at any position, and
    // for those the destination-side gathers (half of the traffic) are never needed
    {
      const int start = pv.node[v];
      to_o[0] = (start == o) ? 0.0 : (pv.directed ? dv.at(start, o) : dv.at(o, start));
      for (int k = 0; k < L; ++k) {
        const int nd = pv.m_node[k * nv + vs];
        from_o[k] = (o == nd) ? 0.0 : dv.at(o, nd);
        to_o[k + 1] = pv.directed ? ((nd == o) ? 0.0 : dv.at(nd, o)) : from_o[k];
      }
    }
    bool reachable = false;
    {
      const double clp = pv.q_clp[rq], cep = pv.q_cep[rq];
      for (int a = 0; a <= L && a <= bviol; ++a) {
        const double t = pv.m_base_t[a * nv + vs] + to_o[a];      // same sum walk_stop forms for the pickup
        if (T + t < cep || !(T + t > clp)) { reachable = true; break; }
      }
    }
    if (reachable) {
      const int start = pv.node[v];
      to_d[0] = (start == d) ? 0.0 : (pv.directed ? dv.at(start, d) : dv.at(d, start));
      for (int k = 0; k < L; ++k) {
        const int nd = pv.m_node[k * nv + vs];
        from_d[k] = (d == nd) ? 0.0 : dv.at(d, nd);
        to_d[k + 1] = pv.directed ? ((nd == d) ? 0.0 : dv.at(nd, d)) : from_d[k];
      }
    }
    const double od = (o == d) ? 0.0 : dv.at(o, d);
    int occ_before = pv.onboard[v];    // occupancy before plan stop a in the unmodified plan
    for (int a = 0; reachable && a <= L && a <= bviol; ++a) {
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
  }
  cmin_out[(int64_t)r * nv + v] = cmin;
}

// Fold: one WARP per request.  Lanes read the smallest costs of 32 consecutive
// vehicles (four chunks ahead: the walk is a chain of dependent decisions, the loads are not); only flagged vehicles
// are re-scanned (by their lane) in vehicle order with the running bound.
__global__ void insertion_fold_kernel(PlanView pv, DistView dv, int n_new, const int32_t* new_rows, const double* cmin,
                                       int32_t* best_veh, int32_t* best_i, int32_t* best_j, double* best_dc) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_new) return;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double dc = inf;                   // dc_ = np.inf (:1452)
  int wv = -1, wi = -1, wj = -1;
  constexpr int AHEAD = 8;   // chunks of 32 vehicles whose costs are loaded together: the loads do not depend on the incumbent
  for (int base0 = 0; base0 < pv.n_veh; base0 += 32 * AHEAD) {
    double cms[AHEAD], c0s[AHEAD];
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int v = base0 + 32 * u + lane;
      cms[u] = inf; c0s[u] = 0.0;
      if (v < pv.n_veh) { cms[u] = cmin[(int64_t)r * pv.n_veh + v]; c0s[u] = pv.cost[v]; }
    }
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int v = base0 + 32 * u + lane;
      const double cm = cms[u], c0 = c0s[u];
      unsigned mask = __ballot_sync(0xffffffffu, cm != inf && !(cm > c0 + dc));
      while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        double ndc = dc; int nv = wv, ni = wi, nj = wj;
        if (lane == l && !(cm > c0 + dc)) {              // re-check: dc may have dropped inside this chunk
          Pair p;
          load_pair(p, pv, dv, v, new_rows[r]);
          int viol = -1;
          for (int i = 0; i <= p.L; ++i) {                         // (:1466-1482)
            for (int j = i + 1; j <= p.L + 1; ++j) {
              double c;
              viol = eval_candidate(p, pv, i, j - 1, c0 + ndc, &c);
              if (viol < 0) { ndc = c - c0; nv = v; ni = i; nj = j; }
              if (viol > 0) break;
            }
            if (viol == 2) break;
          }
        }
        dc = __shfl_sync(0xffffffffu, ndc, l); wv = __shfl_sync(0xffffffffu, nv, l);
        wi = __shfl_sync(0xffffffffu, ni, l); wj = __shfl_sync(0xffffffffu, nj, l);
      }
    }
  }
  if (lane == 0) { best_veh[r] = wv; best_i[r] = wi; best_j[r] = wj; best_dc[r] = dc; }
}

}  // namespace

struct LegacyDev {
  int n_veh = 0, n_stops = 0, n_rows = 0;
  int32_t *node = nullptr, *onboard = nullptr, *cap = nullptr, *off = nullptr, *s_node = nullptr, *s_req = nullptr;
  double *delay = nullptr, *clock = nullptr, *cost = nullptr;
  int8_t* s_pod = nullptr;
  int32_t *q_o = nullptr, *q_d = nullptr;
  double *q_cep = nullptr, *q_clp = nullptr, *q_cld = nullptr, *q_ts = nullptr, *q_tr = nullptr, *q_pwt = nullptr;
  int32_t *perm = nullptr, *m_node = nullptr, *m_req = nullptr, *m_base_viol = nullptr;
  int8_t *m_pod = nullptr, *m_partner = nullptr;
  double *m_seg = nullptr, *m_base_t = nullptr, *m_base_c = nullptr, *m_base_cld = nullptr;
  drs_insertion_params prm{1.5, 1.0, 1.5, 3.0};
};

void drs_tick_free_plans(drs_tick* t) {
  LegacyDev* d = t->legacy;
  if (!d) return;
  void* ptrs[] = {d->node, d->onboard, d->cap, d->off, d->s_node, d->s_req, d->delay, d->clock, d->cost, d->s_pod,
                  d->q_o, d->q_d, d->q_cep, d->q_clp, d->q_cld, d->q_ts, d->q_tr, d->q_pwt, d->perm, d->m_node, d->m_req,
                  d->m_base_viol, d->m_pod, d->m_partner, d->m_seg, d->m_base_t, d->m_base_c, d->m_base_cld};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete d;
  t->legacy = nullptr;
}

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
  const int nstops = nv ? b->stop_off[nv] : 0;
  for (int v = 0; v < nv; ++v) {
    DRS_REQUIRE(b->start_node[v] >= 0 && b->start_node[v] < n, DRS_ERR_RANGE, "drs_tick_set_plans: vehicle %d node out of range", v);
    const int cnt = b->stop_off[v + 1] - b->stop_off[v];
    DRS_REQUIRE(cnt >= 0 && cnt <= MAXL, DRS_ERR_ARG, "drs_tick_set_plans: vehicle %d has %d planned stops (max %d)", v, cnt, MAXL);
    for (int s = b->stop_off[v]; s < b->stop_off[v + 1]; ++s) {
      DRS_REQUIRE(b->stop_node[s] >= 0 && b->stop_node[s] < n, DRS_ERR_RANGE, "drs_tick_set_plans: stop %d node out of range", s);
      DRS_REQUIRE(b->stop_req[s] >= 0 && b->stop_req[s] < nq, DRS_ERR_RANGE, "drs_tick_set_plans: stop %d request row out of range", s);
    }
  }
  for (int r = 0; r < nq; ++r)
    DRS_REQUIRE(q->o_node[r] >= 0 && q->o_node[r] < n && q->d_node[r] >= 0 && q->d_node[r] < n, DRS_ERR_RANGE,
                "drs_tick_set_plans: request row %d node out of range", r);
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  if (!t->legacy) t->legacy = new LegacyDev();
  LegacyDev& d = *t->legacy;
  int rc;
  if ((rc = drs_upload(&d.node, b->start_node, nv, st)) || (rc = drs_upload(&d.delay, b->start_delay, nv, st)) ||
      (rc = drs_upload(&d.onboard, b->onboard, nv, st)) || (rc = drs_upload(&d.clock, b->clock, nv, st)) ||
      (rc = drs_upload(&d.cap, b->capacity, nv, st)) || (rc = drs_upload(&d.cost, b->cost, nv, st)) ||
      (rc = drs_upload(&d.off, b->stop_off, nv ? nv + 1 : 0, st)) || (rc = drs_upload(&d.s_node, b->stop_node, nstops, st)) ||
      (rc = drs_upload(&d.s_req, b->stop_req, nstops, st)) || (rc = drs_upload(&d.s_pod, b->stop_pod, nstops, st)) ||
      (rc = drs_upload(&d.q_o, q->o_node, nq, st)) || (rc = drs_upload(&d.q_d, q->d_node, nq, st)) ||
      (rc = drs_upload(&d.q_cep, q->Cep, nq, st)) || (rc = drs_upload(&d.q_clp, q->Clp, nq, st)) ||
      (rc = drs_upload(&d.q_cld, q->Cld, nq, st)) || (rc = drs_upload(&d.q_ts, q->Ts, nq, st)) ||
      (rc = drs_upload(&d.q_tr, q->Tr, nq, st)) || (rc = drs_upload(&d.q_pwt, q->pwt, nq, st)))
    return rc;
  // slot-major copy, vehicles sorted by plan length, then by start vertex (neighbouring lanes gather neighbouring
  // columns of the request's rows: vertex ids are spatially coherent under the host layer's locality numbering)
  std::vector<int32_t> perm(std::max(nv, 1));
  for (int v = 0; v < nv; ++v) perm[v] = v;
  std::stable_sort(perm.begin(), perm.begin() + nv, [&](int x, int y) {
    const int lx = b->stop_off[x + 1] - b->stop_off[x], ly = b->stop_off[y + 1] - b->stop_off[y];
    return lx != ly ? lx < ly : b->start_node[x] < b->start_node[y];
  });
  const size_t slots = (size_t)(MAXL + 1) * std::max(nv, 1);
  std::vector<int32_t> m_node(slots, 0), m_req(slots, 0);
  std::vector<int8_t> m_pod(slots, 0), m_partner(slots, -1);
  for (int vs = 0; vs < nv; ++vs) {
    const int v = perm[vs], s0 = b->stop_off[v], L = b->stop_off[v + 1] - s0;
    for (int k = 0; k < L; ++k) {
      const size_t at = (size_t)k * nv + vs;
      m_node[at] = b->stop_node[s0 + k]; m_req[at] = b->stop_req[s0 + k]; m_pod[at] = b->stop_pod[s0 + k];
      if (b->stop_pod[s0 + k] < 0)   // plan index of the pickup that belongs to this drop-off (-1: rider on board)
        for (int m = 0; m < k; ++m)
          if (b->stop_req[s0 + m] == b->stop_req[s0 + k] && b->stop_pod[s0 + m] > 0) m_partner[at] = (int8_t)m;
    }
  }
  if ((rc = drs_upload(&d.perm, perm.data(), nv, st)) || (rc = drs_upload(&d.m_node, m_node.data(), slots, st)) ||
      (rc = drs_upload(&d.m_req, m_req.data(), slots, st)) || (rc = drs_upload(&d.m_pod, m_pod.data(), slots, st)) ||
      (rc = drs_upload(&d.m_partner, m_partner.data(), slots, st)))
    return rc;
  void** scratch[] = {(void**)&d.m_seg, (void**)&d.m_base_t, (void**)&d.m_base_c, (void**)&d.m_base_cld, (void**)&d.m_base_viol};
  const size_t bytes[] = {sizeof(double) * slots, sizeof(double) * slots, sizeof(double) * slots, sizeof(double) * slots,
                          sizeof(int32_t) * std::max(nv, 1)};
  for (int i = 0; i < 5; ++i) {
    if (*scratch[i]) { cudaFree(*scratch[i]); *scratch[i] = nullptr; }
    DRS_CUDA(cudaMalloc(scratch[i], bytes[i]));
  }
  DRS_CUDA(cudaStreamSynchronize(st));
  d.n_veh = nv; d.n_stops = nstops; d.n_rows = nq;
  if (prm) d.prm = *prm;
  return DRS_OK;
}

extern "C" int drs_tick_insertion_eval(drs_tick* t, int32_t n_new, const int32_t* new_rows, int32_t* best_veh,
                                       int32_t* best_i, int32_t* best_j, double* best_dc, int64_t* n_candidates) {
  DRS_REQUIRE(t && t->legacy, DRS_ERR_STATE, "drs_tick_insertion_eval: call drs_tick_set_plans first");
  DRS_REQUIRE(t->g->apsp_valid, DRS_ERR_STATE, "drs_tick_insertion_eval: travel-time matrix not built (or invalidated)");
  DRS_REQUIRE(n_new >= 0 && (n_new == 0 || (new_rows && best_veh && best_i && best_j && best_dc)), DRS_ERR_ARG,
              "drs_tick_insertion_eval: bad argument");
  if (n_candidates) *n_candidates = 0;
  if (n_new == 0) return DRS_OK;
  LegacyDev& d = *t->legacy;
  for (int r = 0; r < n_new; ++r)
    DRS_REQUIRE(new_rows[r] >= 0 && new_rows[r] < d.n_rows, DRS_ERR_RANGE, "drs_tick_insertion_eval: request row %d out of range",
                new_rows[r]);
  const int64_t total = (int64_t)d.n_veh * n_new;
  DRS_REQUIRE(total < ((int64_t)1 << 31), DRS_ERR_ARG, "drs_tick_insertion_eval: %lld pairs exceed 2^31", (long long)total);
  DrsDeviceGuard guard(t->g->device);
  cudaStream_t st = t->g->stream;
  PlanView pv{d.n_veh, d.node, d.onboard, d.cap, d.off, d.s_node, d.s_req, d.delay, d.clock, d.cost, d.s_pod,
              d.q_o, d.q_d, d.q_cep, d.q_clp, d.q_cld, d.q_ts, d.q_tr, d.q_pwt, d.prm,
              d.perm, d.m_node, d.m_req, d.m_pod, d.m_partner, d.m_seg, d.m_base_t, d.m_base_c, d.m_base_cld, d.m_base_viol,
              t->g->directed};
  DistView dv{t->g->d_dist, t->g->ld, t->g->mode};
  int32_t *d_rows = nullptr, *d_bv = nullptr, *d_bi = nullptr, *d_bj = nullptr;
  double *d_cmin = nullptr, *d_dc = nullptr;
  unsigned long long* d_ncand = nullptr;
  auto cleanup = [&]() { cudaFree(d_rows); cudaFree(d_bv); cudaFree(d_bi); cudaFree(d_bj); cudaFree(d_cmin); cudaFree(d_dc); cudaFree(d_ncand); };
  int rc = drs_upload(&d_rows, new_rows, n_new, st);
  if (rc) { cleanup(); return rc; }
  if (cudaMalloc((void**)&d_bv, sizeof(int32_t) * n_new) != cudaSuccess || cudaMalloc((void**)&d_bi, sizeof(int32_t) * n_new) != cudaSuccess ||
      cudaMalloc((void**)&d_bj, sizeof(int32_t) * n_new) != cudaSuccess || cudaMalloc((void**)&d_dc, sizeof(double) * n_new) != cudaSuccess ||
      cudaMalloc((void**)&d_cmin, sizeof(double) * std::max<int64_t>(total, 1)) != cudaSuccess ||
      cudaMalloc((void**)&d_ncand, sizeof(unsigned long long)) != cudaSuccess) {
    cleanup();
    drs_set_error("drs_tick_insertion_eval: cudaMalloc failed");
    return DRS_ERR_NOMEM;
  }
  cudaMemsetAsync(d_ncand, 0, sizeof(unsigned long long), st);
  DrsTimer tm_scan(st), tm_fold(st);
  tm_scan.start();
  if (total > 0) {
    plan_prep_kernel<<<(d.n_veh + 127) / 128, 128, 0, st>>>(pv, dv);
    insertion_scan_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(pv, dv, n_new, d_rows, d_cmin, d_ncand);
    drs_count_launch(2);
  }
  tm_scan.stop();
  tm_fold.start();
  insertion_fold_kernel<<<(n_new * 32 + 127) / 128, 128, 0, st>>>(pv, dv, n_new, d_rows, d_cmin, d_bv, d_bi, d_bj, d_dc);
  drs_count_launch();
  tm_fold.stop();
  unsigned long long ncand = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(best_veh, d_bv, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(best_i, d_bi, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(best_j, d_bj, sizeof(int32_t) * n_new, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(best_dc, d_dc, sizeof(double) * n_new, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ncand, d_ncand, sizeof(ncand), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) { t->last_ms[2] = tm_scan.ms(); t->last_ms[3] = tm_fold.ms(); }
  cleanup();
  if (e != cudaSuccess) {
    drs_set_error("drs_tick_insertion_eval: %s", cudaGetErrorString(e));
    return DRS_ERR_CUDA;
  }
  if (n_candidates) *n_candidates = (int64_t)ncand;
  return DRS_OK;
}
