"""CPU oracle for the dispatch half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (drs_abm_b200/) never does.

Restates, in plain Python (fp64, like the reference):
  * lib/optimUtils.py:12-38     StopLocation                 -> Stop
  * lib/optimUtils.py:84-110    createLocations              -> create_locations
  * lib/optimUtils.py:41-81     permutePassengerOrder        -> orderings
  * lib/optimUtils.py:456-557   computeRoutingCost           -> compute_routing_cost
  * lib/optimUtils.py:432-453   minimalDelayPath             -> minimal_delay_path
  * lib/optimUtils.py:113-160   findOptimalRouting           -> find_optimal_routing
  * lib/optimUtils.py:651-664   greatCircleDistance, shortestPathTime
  * lib/Optimization.py:104-225 AlonsoMora.generatePairwiseGraph ("enumerate") -> rv_graph
  * lib/Optimization.py:227-357 generateRTVGraph (trips of every size)       -> rtv_trips
  * lib/Optimization.py:359-408 assignFromRTVGraph (Gurobi ILP)              -> assign_ilp (HiGHS)
  * lib/Agents.py:1451-1572     Model.insert_heuristics / test_constraints_get_cost -> legacy_*
  * lib/Agents.py:1429-1448     Model.insertion_heuristics (the sequential queue)    -> legacy_insertion_queue
  * lib/Agents.py:148-297       Veh.build_route: what the insertion path reads back
                                (route, veh.c, Req.pwt)                              -> legacy_build_route

Travel times come from a callable ``tt(node_a, node_b)`` (the oracle's Dijkstra
matrix) in place of ``G.shortest_paths``.

Parity pin: the reference's own lib/optimUtils.py, lib/Optimization.py and
lib/Agents.py functions, run unmodified in the authoring container on the igraph
facade (oracle/ref_shim), generated the fixtures in tests/golden/
(oracle/make_golden.py); tests/test_oracle_golden.py checks this restatement
against them -- for the legacy insertion heuristic call by call: every
test_constraints_get_cost invocation (vehicle, i, j, bound, flag, cost, viol
code), the winner, and Req.Cld / Req.pwt / Veh.c after the winner's build_route
(tests/golden/legacy_*.json).  Two things stay "parity unpinned" because
the reference itself leaves them undefined:
  * which of several equal-cost orderings minimalDelayPath returns -- the
    reference iterates a ``set`` of tuples hashed by object address
    (lib/optimUtils.py:55,79,442).  DEFINED here: the lexicographically smallest
    stop-index sequence (stop index = position in createLocations' output).
  * which of several optimal ILP solutions Gurobi returns.  The oracle returns
    one optimum from HiGHS; tests compare objective values and feasibility.
"""
import math

import numpy as np

CST_SPEED = 9            # lib/Constants.py:148
MAX_DETOUR = 1.5         # lib/Constants.py:144
COEF_INVEH = 1.0         # lib/Constants.py:163-165
COEF_WAIT = 1.5
COEF_EXTRA_WAIT = 3.0
INF_ROUTE_COST = 1e6     # lib/optimUtils.py:156,437


class Stop(object):
    """lib/optimUtils.py:12-38.  ``node`` is the road-graph vertex of (lng, lat)."""

    def __init__(self, node, lng, lat, rid=None, order=None, lbT=None, ubT=None, pod=None):
        self.node, self.lng, self.lat = node, lng, lat
        self.rid, self.order, self.lbT, self.ubT, self.pod = rid, order, lbT, ubT, pod

    def check_time(self, T):          # checkTimeConstraints, :37-38
        return (T >= self.lbT) and (T <= self.ubT)


def create_locations(new_reqs, veh=None):
    """lib/optimUtils.py:84-110.  new_reqs: RV request-vertex dicts with keys rid,
    o_node, origin(lng,lat), d_node, destination, Cep (already Cep - T), Clp, Ced, Cld.
    veh: dict with pax_reqs / wait_reqs (request records with absolute windows)."""
    stops = []
    for rq in new_reqs:
        stops.append(Stop(rq["o_node"], rq["origin"][0], rq["origin"][1], rq["rid"], 0, rq["Cep"], rq["Clp"], 1))
        stops.append(Stop(rq["d_node"], rq["destination"][0], rq["destination"][1], rq["rid"], 1, rq["Ced"],
                          rq["Cld"], -1))
    if veh is not None:
        for rq in veh["pax_reqs"]:          # on board: drop-off only, order 0 (:95-99)
            stops.append(Stop(rq["d_node"], rq["dlng"], rq["dlat"], rq["rid"], 0, rq["Ced"], rq["Cld"], -1))
        for rq in veh["wait_reqs"]:         # assigned earlier, not yet picked up (:101-108)
            stops.append(Stop(rq["o_node"], rq["olng"], rq["olat"], rq["rid"], 0, rq["Cep"], rq["Clp"], 1))
            stops.append(Stop(rq["d_node"], rq["dlng"], rq["dlat"], rq["rid"], 1, rq["Ced"], rq["Cld"], -1))
    return stops


def orderings(stops):
    """Every precedence-respecting ordering, as tuples of stop indices, in
    lexicographic order.  Same SET as permutePassengerOrder (lib/optimUtils.py:41-81):
    distinct permutations of the rid multiset with the k-th occurrence of a rid
    mapped to that rid's order-k stop."""
    n = len(stops)
    before = [None] * n                      # index of the stop that must precede stop i
    for i, s in enumerate(stops):
        if s.order and s.order > 0:
            for j, q in enumerate(stops):
                if q.rid == s.rid and q.order == s.order - 1:
                    before[i] = j
    out = []
    used = [False] * n
    cur = []

    def rec():
        if len(cur) == n:
            out.append(tuple(cur))
            return
        for i in range(n):
            if not used[i] and (before[i] is None or used[before[i]]):
                used[i] = True
                cur.append(i)
                rec()
                cur.pop()
                used[i] = False
    rec()
    return out


def compute_routing_cost(ordered, tt, T=0, capacity=1e6, start_pax=0, inf_stop=True, inf_cost=INF_ROUTE_COST):
    """lib/optimUtils.py:456-557.  ``ordered[0]`` is the start (vehicle node or first stop)."""
    tot_cost = 0
    feasible = True
    pax = start_pax
    assert pax <= capacity                                   # :468
    seg = []
    for i in range(len(ordered) - 1):                        # :475-492
        dest = ordered[i + 1]
        pax += dest.pod
        if pax > capacity:
            feasible = False
            if inf_stop:
                break
    if feasible or not inf_stop:                             # :514-550
        seg = [tt(ordered[i].node, ordered[i + 1].node) for i in range(len(ordered) - 1)]
        cum = [T + sum(seg[:i + 1]) for i in range(len(seg))]    # :536 (left-to-right from int 0)
        for k, dest in enumerate(ordered[1:]):
            t_arr = cum[k]
            if not dest.check_time(t_arr):
                feasible = False
                if inf_stop:
                    break
            if dest.pod < 0:
                tot_cost += t_arr - dest.lbT                 # :549-550
    if feasible:
        return tot_cost
    if inf_stop:
        return inf_cost
    return inf_cost + tot_cost


def minimal_delay_path(stops, tt, T=0, start=None, capacity=1e6, start_pax=0):
    """lib/optimUtils.py:432-453 with the canonical tie rule.  Returns
    (cost, route, order) -- order is the winning stop-index tuple (None if infeasible)."""
    best_cost, best_route, best_order = INF_ROUTE_COST, None, None
    for order in orderings(stops):
        seq = ([start] if start is not None else []) + [stops[i] for i in order]
        cost = compute_routing_cost(seq, tt, T, capacity, start_pax)
        if cost < best_cost:                                  # :447 strict: first (lexicographically smallest) kept
            best_cost = cost
            best_route = [(s.rid, s.pod, s.lng, s.lat) for s in seq if s.rid is not None]
            best_order = order
    return best_cost, best_route, best_order


def find_optimal_routing(new_reqs, tt, veh=None, T=0):
    """lib/optimUtils.py:113-160."""
    stops = create_locations(new_reqs, veh)
    if veh is None:
        return minimal_delay_path(stops, tt, T)               # :160
    start_time = T + veh["t"]                                 # :117
    start = Stop(veh["node"], veh["location"][0], veh["location"][1])
    capacity = veh["capacity"]
    pax = len(veh["pax_reqs"])
    all_rids = []
    for s in stops:
        if s.rid not in all_rids:
            all_rids.append(s.rid)
    if len(all_rids) >= 3:                                    # :127-156 pairwise pre-filter, no capacity
        veh_rids = set([r["rid"] for r in veh["pax_reqs"]]) | set([r["rid"] for r in veh["wait_reqs"]])
        for a in range(len(all_rids)):
            for b in range(a + 1, len(all_rids)):
                r1, r2 = all_rids[a], all_rids[b]
                if r1 in veh_rids and r2 in veh_rids:
                    continue
                combo = [s for s in stops if s.rid == r1 or s.rid == r2]
                _, combo_route, _ = minimal_delay_path(combo, tt, start_time, start)
                if combo_route is None:
                    return INF_ROUTE_COST, None, None
    return minimal_delay_path(stops, tt, start_time, start, capacity, pax)   # :157-158


def great_circle_distance(olat, olng, dlat, dlng):
    """lib/optimUtils.py:651-653 (equirectangular metres)."""
    return (6371000 * 2 * math.pi / 360 * np.sqrt(
        (math.cos((olat + dlat) * math.pi / 360) * (olng - dlng)) ** 2 + (olat - dlat) ** 2))


def request_vertex(rq, T):
    """The RV-graph request vertex (lib/Optimization.py:111-118): note Cep - T."""
    return {"rid": rq["rid"], "o_node": rq["o_node"], "d_node": rq["d_node"],
            "origin": (rq["olng"], rq["olat"]), "destination": (rq["dlng"], rq["dlat"]),
            "Cep": rq["Cep"] - T, "Clp": rq["Clp"], "Ced": rq["Ced"], "Cld": rq["Cld"]}


def rv_graph(vehs, reqs, tt, T=0, with_rr=True):
    """lib/Optimization.py:104-225 with routing_method == "enumerate".

    vehs: list of dicts {vid, capacity, node, location(lng,lat), t, pax_reqs, wait_reqs}
    reqs: list of unassigned request records {rid, o_node, d_node, olng, olat, dlng, dlat, Cep, Clp, Ced, Cld}
    Returns (rr_edges, rv_edges, counters):
      rr_edges {(rid1, rid2): (cost, route)}  rid1 listed before rid2 in ``reqs``
      rv_edges {(vid, rid): (cost, route, order)}
    """
    verts = [request_vertex(r, T) for r in reqs]
    rr = {}
    if with_rr:
        for a, r1 in enumerate(verts):                        # :142-165
            for b, r2 in enumerate(verts):
                if a == b:
                    continue
                key = (r1["rid"], r2["rid"]) if a < b else (r2["rid"], r1["rid"])
                if key in rr:
                    continue
                d = great_circle_distance(r1["origin"][1], r1["origin"][0], r2["origin"][1], r2["origin"][0])
                tmin = d / CST_SPEED
                if T + tmin > r1["Clp"] or T + tmin > r2["Clp"]:
                    continue
                cost, route, _ = find_optimal_routing((r1, r2), tt, None, T)
                if route is not None:
                    rr[key] = (cost, route)
    rv = {}
    counters = {"feasible": 0, "infeasible": 0, "saved": 0, "skipped": 0}
    for rq in verts:                                          # :176-218
        for veh in vehs:
            if T + tt(veh["node"], rq["o_node"]) > rq["Clp"]:            # :183-186 (T, not T + t)
                counters["skipped"] += 1
                continue
            busy = len(veh["pax_reqs"]) + len(veh["wait_reqs"]) > 0
            if busy:                                          # :193-202 empty-dummy-vehicle check
                dummy = dict(veh, pax_reqs=[], wait_reqs=[])
                _, dummy_route, _ = find_optimal_routing([rq], tt, dummy, T)
                if dummy_route is None:
                    counters["saved"] += 1
                    continue
            cost, route, order = find_optimal_routing([rq], tt, veh, T)   # :204-208
            if route is not None:
                rv[(veh["vid"], rq["rid"])] = (cost, route, order)
                counters["feasible"] += 1
            else:
                counters["infeasible"] += 1
    return rr, rv, counters


def assign_ilp(veh_ids, req_ids, trips):
    """lib/Optimization.py:359-408 solved with HiGHS instead of Gurobi.

    trips: list of (vid, tuple_of_rids, cost).  Binary e per trip, x per request;
    min sum cost*e + 1e6*sum x; sum_trips-of-veh e <= 1; for each request:
    sum_{trips containing it} e + x == 1.  Returns (chosen trip indices, rejected rids, objective)."""
    from scipy.optimize import milp, LinearConstraint, Bounds
    nt, nr = len(trips), len(req_ids)
    if nt == 0:
        return [], set(req_ids), 1e6 * nr
    c = np.array([t[2] for t in trips] + [1e6] * nr, dtype=np.float64)
    rows, lo, hi = [], [], []
    vpos = {}
    for k, (vid, rids, _) in enumerate(trips):
        vpos.setdefault(vid, []).append(k)
    for vid, ks in vpos.items():
        row = np.zeros(nt + nr); row[ks] = 1
        rows.append(row); lo.append(-np.inf); hi.append(1)
    for j, rid in enumerate(req_ids):
        row = np.zeros(nt + nr)
        for k, (vid, rids, _) in enumerate(trips):
            if rid in rids:
                row[k] = 1
        row[nt + j] = 1
        rows.append(row); lo.append(1); hi.append(1)
    res = milp(c, constraints=LinearConstraint(np.array(rows), lo, hi), integrality=np.ones(nt + nr),
               bounds=Bounds(0, 1))
    assert res.status == 0, res.message
    x = np.round(res.x).astype(int)
    chosen = [k for k in range(nt) if x[k] > 0]
    rejected = set(req_ids[j] for j in range(nr) if x[nt + j] > 0)
    return chosen, rejected, float(res.fun)


def rtv_trips(vehs, reqs, rr, rv, tt, T=0):
    """lib/Optimization.py:227-357 ("enumerate").  Returns {(vid, rid-tuple): (cost, route)}.

    Iteration orders the reference leaves to igraph / dict insertion are canonical here: a
    vehicle's neighbours in request order (igraph adjacency lists are sorted by vertex id), trips
    keyed by rid tuples in request order.  The early ``break`` at lib/Optimization.py:325-327
    (leaves the inner t2 loop when a (k-1)-subset is missing) is kept."""
    pos = {r["rid"]: i for i, r in enumerate(reqs)}
    verts = {r["rid"]: request_vertex(r, T) for r in reqs}
    out = {}

    def has_rr(a, b):
        return ((a, b) if pos[a] < pos[b] else (b, a)) in rr

    for veh in vehs:
        vid = veh["vid"]
        singles = sorted([rid for (v, rid) in rv if v == vid], key=lambda r: pos[r])
        trips = {1: {}}
        for rid in singles:                                   # :249-262
            trips[1][(rid,)] = rv[(vid, rid)][:2]
        trips[2] = {}
        for a in singles:                                     # :265-288
            for b in singles:
                if a == b or (b, a) in trips[2] or (a, b) in trips[2]:
                    continue
                if pos[a] > pos[b]:
                    continue          # (a, b) with a first was already evaluated: same stop set, same result
                if has_rr(a, b):
                    cost, route, _ = find_optimal_routing([verts[a], verts[b]], tt, veh, T)
                    if route is not None:
                        trips[2][(a, b)] = (cost, route)
        for k in range(3, veh["capacity"] + 1):               # :291-347
            trips[k] = {}
            tested = []
            keys = list(trips[k - 1].keys())
            for t1 in keys:
                for t2 in keys:
                    trip = set(t1) | set(t2)
                    if len(trip) == k and trip not in tested:
                        tested.append(trip)
                        ok = []
                        for r in trip:
                            without = trip - {r}
                            ok.append(len([t for t in keys if all(q in t for q in without)]) == 1)
                        if not all(ok):
                            break
                        members = sorted(trip, key=lambda r: pos[r])
                        if all(has_rr(a, b) for i, a in enumerate(members) for b in members[i + 1:]):
                            cost, route, _ = find_optimal_routing([verts[r] for r in members], tt, veh, T)
                            if route is not None:
                                trips[k][tuple(members)] = (cost, route)
        for k in trips:
            for key, val in trips[k].items():
                out[(vid, key)] = val
    return out


def optimize_assignment(vehs, reqs, tt, T=0, max_trip_size=None):
    """lib/Optimization.py:46-76: RV graph -> RTV graph -> ILP."""
    rr, rv, counters = rv_graph(vehs, reqs, tt, T, with_rr=True)
    trips_d = rtv_trips(vehs, reqs, rr, rv, tt, T)
    if max_trip_size is not None:
        trips_d = {k: v for k, v in trips_d.items() if len(k[1]) <= max_trip_size}
    trips = [(vid, rids, cost) for (vid, rids), (cost, route) in trips_d.items()]
    meta = {"objective": None, "rv": rv, "rr": rr, "trips": trips_d, "counters": counters}
    if not trips:
        return ([], []), meta                                 # :76 (empty RTV graph)
    chosen, rejected, obj = assign_ilp([v["vid"] for v in vehs], [r["rid"] for r in reqs], trips)
    routes = {}
    for k in chosen:
        vid, rids, _ = trips[k]
        routes[vid] = trips_d[(vid, rids)][1]
    meta["objective"] = obj
    return (routes, rejected), meta


# ------------------------------------------------------------------ legacy insertion (D10)
def legacy_test_constraints_get_cost(tt, route, veh, reqs, new_rid, C):
    """lib/Agents.py:1493-1572 with PRE_DRAW_TTS False.

    route: list of (rid, pod, node); veh: dict {n, T, K, node, t0} where node/t0
    are the start vertex and the time to reach it (:1508-1516); reqs: {rid:
    record with Cep, Clp, Cld, Ts, Tr, pwt} -- Cld is MUTATED like the reference.
    Returns (flag, cost, viol)."""
    c = 0.0
    t = 0.0 + veh["t0"]
    n = veh["n"]
    T = veh["T"]
    K = veh["K"]
    cur = veh["node"]
    for (rid, pod, node) in route:                            # :1518-1521 capacity pre-check
        n += pod
        if n > K:
            return False, None, 1
    n = veh["n"]
    for (rid, pod, node) in route:                            # :1523-1571
        rq = reqs[rid]
        dt = tt(cur, node)
        t += dt
        if pod == 1:
            if T + t < rq["Cep"]:
                dt += rq["Cep"] - T - t
                t += rq["Cep"] - T - t
                rq["Cld"] = rq["Cep"] + MAX_DETOUR * rq["Ts"]
            elif T + t > rq["Clp"]:
                return False, None, 2
            else:
                rq["Cld"] = T + t + MAX_DETOUR * rq["Ts"]
        elif pod == -1 and T + t > rq["Cld"]:
            return False, None, 3
        c += n * dt * COEF_INVEH
        n += pod
        assert n <= K
        if pod == 1:
            if rq["pwt"] > 0:
                c += (T + t - rq["Tr"] - rq["pwt"]) * COEF_EXTRA_WAIT
            else:
                c += t * COEF_WAIT
        if c > C:
            return False, None, 0
        cur = node
    return True, c, -1


def legacy_insert_heuristics(tt, vehs, reqs, new_rid):
    """lib/Agents.py:1451-1490: scan vehicles, pickup slot i, drop-off slot j; the
    last candidate whose running cost never exceeds the incumbent bound wins.

    vehs: list of dicts {vid, n, T, K, c, node, t0, route:[(rid, pod, node)]}.
    Returns (vid or None, route or None, dc, evaluated) -- evaluated counts calls."""
    rq = reqs[new_rid]
    dc_best, veh_best, route_best = np.inf, None, None
    evaluated = 0
    viol = None
    for veh in vehs:
        route = list(veh["route"])
        l = len(route)
        c = veh["c"]
        for i in range(l + 1):
            for j in range(i + 1, l + 2):
                route.insert(i, (new_rid, 1, rq["o_node"]))
                route.insert(j, (new_rid, -1, rq["d_node"]))
                flag, c_new, viol = legacy_test_constraints_get_cost(tt, route, veh, reqs, new_rid, c + dc_best)
                evaluated += 1
                if flag:
                    dc_best = c_new - c
                    veh_best = veh["vid"]
                    route_best = list(route)
                route.pop(j)
                route.pop(i)
                if viol > 0:
                    break
            if viol == 2:
                break
    return veh_best, route_best, dc_best, evaluated


def legacy_build_route(tt, veh, route, reqs, T):
    """What Veh.build_route (lib/Agents.py:148-297) leaves behind for the next insertion scan: the
    vehicle's route, ``veh.c`` (:256-270), ``idle`` and the predicted waiting time ``pwt`` of the
    requests still to be picked up (:165-220).  Mutates ``veh`` and ``reqs``.

    veh carries, besides the fields of legacy_test_constraints_get_cost, ``route`` (current legs) and
    ``keep`` = None or {node, t, et}: the first step of the current route, which
    clear_route_but_keep_next_node (:319-336) keeps in front of the new legs.  Leg times are the
    path sums get_routing returns (``tt``); a pickup reached before Cep waits (add_leg, :363-367)."""
    old = veh["route"]
    pick = set(r for (r, pod, _) in old if pod == 1)
    drop = set(r for (r, pod, _) in old if pod == -1)
    in_veh = drop - pick                                        # get_passenger_requests, :604-615
    not_in = set(r for (r, _, _) in route if r not in in_veh)
    merge = veh["keep"] is not None                             # len(self.route) > 0, :156-160
    if len(route) == 0:                                         # :163-169
        veh.update(idle=True, c=0.0, route=[], keep=veh["keep"] if merge else None)
        return
    for r in in_veh:
        reqs[r]["pwt"] = 0
    for r in not_in:
        reqs[r]["pwt"] = 0
        if merge:
            reqs[r]["pwt"] += veh["keep"]["et"]
    cur = veh["keep"]["node"] if merge else veh["node_pos"]
    vt = 0.0 + (veh["keep"]["t"] if merge else 0.0)             # self.t after clear_route(_but_keep_next_node)
    leg_t = []
    first_hop = None
    for (r, pod, node) in route:
        pred_t = tt(cur, node)                                  # sum of step.et == l['duration'] without link uncertainty
        t_leg = pred_t
        if pod == 1 and T + vt + t_leg < reqs[r]["Cep"]:        # add_leg waits at an early pickup
            t_leg += reqs[r]["Cep"] - (T + vt + t_leg)
        vt += t_leg
        leg_t.append(t_leg)
        for q in not_in:
            reqs[q]["pwt"] += pred_t
        if pod == 1:
            not_in.discard(r)
        cur = node
    if merge:                                                   # :222-240 the kept step joins the first new leg
        leg_t[0] = veh["keep"]["t"] + leg_t[0]
    c, t, n = 0.0, 0.0, veh["n"]
    for (r, pod, node), lt in zip(route, leg_t):                # :256-262
        t += lt
        c += n * lt * COEF_INVEH
        n += pod
        c += t * COEF_WAIT if pod == 1 else 0
    assert n == 0
    veh.update(idle=False, c=c, route=list(route), leg_t=leg_t)


def legacy_insertion_queue(tt, vehs, reqs, queue, T, first_step=None, trace=None):
    """Model.insertion_heuristics (lib/Agents.py:1429-1448, DISTANCE_THRESHOLD = 0): the queued requests are
    inserted ONE AFTER ANOTHER, each scan seeing the plans, costs, pwt and Cld left by the previous winners.

    vehs: dicts {vid, n, T, K, c, node, t0 (scan start, :1508-1516), node_pos (vertex of the position or None),
    keep, idle, route}.  first_step(a, b) -> (next vertex, step time) of the path a -> b: needed when a vehicle that
    had no route wins twice in one tick (its second build_route keeps the first step of the first leg).
    Returns [(rid, vid or None, i, j, dc)]; ``trace`` (a list) receives every test_constraints_get_cost call."""
    out = []
    for rid in queue:
        rq = reqs[rid]
        dc_best, win = np.inf, None
        viol = None
        for veh in vehs:
            route = list(veh["route"]) if not veh["idle"] else []
            l = len(route)
            c = veh["c"]
            for i in range(l + 1):
                for j in range(i + 1, l + 2):
                    route.insert(i, (rid, 1, rq["o_node"]))
                    route.insert(j, (rid, -1, rq["d_node"]))
                    flag, c_new, viol = legacy_test_constraints_get_cost(tt, route, veh, reqs, rid, c + dc_best)
                    if trace is not None:
                        trace.append([veh["vid"], i, j, c + dc_best, flag, c_new, viol])
                    if flag:
                        dc_best = c_new - c
                        win = (veh, list(route), i, j)
                    route.pop(j)
                    route.pop(i)
                    if viol > 0:
                        break
                if viol == 2:
                    break
        if win is None:
            out.append((rid, None, -1, -1, np.inf))
            continue
        veh, route, i, j = win
        had_route = veh["keep"] is not None
        legacy_build_route(tt, veh, route, reqs, T)
        if not had_route:                                       # the new route's first step is what a later build keeps
            nxt, st = first_step(veh["node_pos"], route[0][2]) if first_step else (route[0][2], None)
            veh["keep"] = {"node": nxt, "t": st, "et": st}
        out.append((rid, veh["vid"], i, j, dc_best))
    return out
