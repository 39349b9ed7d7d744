"""Generates tests/golden/*.json by running the REFERENCE'S OWN hot-path code.

TEST INFRASTRUCTURE ONLY; runs in the authoring container (needs /root/reference,
which does not exist on the GPU box -- the fixtures it writes are committed).

lib/RoutingEngine.py, lib/optimUtils.py, lib/Optimization.py and lib/Agents.py
are imported unmodified from /root/reference with oracle/ref_shim (an igraph
facade over networkx and a gurobipy stand-in over scipy HiGHS) ahead of them on
sys.path.  Every number stored below was produced by those files.

    python oracle/make_golden.py          # rewrites tests/golden/
"""
import contextlib
import io
import itertools
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DRS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, ROOT)

from drs_abm_b200 import gml, synth  # noqa: E402  (fixture maps only)

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

with contextlib.redirect_stdout(io.StringIO()):
    from lib.RoutingEngine import RoutingEngine  # noqa: E402
    from lib import optimUtils  # noqa: E402
    from lib.Optimization import AlonsoMora  # noqa: E402
    from lib.Agents import Veh, Req  # noqa: E402
    import lib.Constants as RC  # noqa: E402

MAPS = {
    "g8_int": (8, 1.0, 8, "int"),
    "g12_float": (12, 0.9, 12, "float"),
    "g20_int": (20, 1.0, 20, "int"),         # BASELINE config C1
    "g20_uniform": (20, 1.0, 20, "uniform"),  # tie stress
    "g10_dfloat": (10, 1.0, 10, "dfloat"),    # directed, float-weighted
}


def jf(x):
    """JSON-safe float (inf -> string)."""
    x = float(x)
    if np.isinf(x):
        return "inf"
    return x


def make_maps():
    for name, (n_side, keep, seed, w) in MAPS.items():
        data = synth.grid_city(n_side, seed, keep, w)
        gml.write_gml(os.path.join(GOLD, name + ".gml"), data)


def routing_golden(name, npairs, nrouting, seed):
    path = os.path.join(GOLD, name + ".gml")
    router = RoutingEngine(graph_loc=path, cst_speed=9, seed=7)
    G = router.G
    n = len(G.vs)
    rs = np.random.RandomState(seed)
    out = {"map": name, "n": n, "m": len(G.es), "pairs": [], "routes": [], "rows": [], "random_locations": []}
    unique_paths = "float" in name
    for _ in range(npairs):
        o, d = int(rs.randint(n)), int(rs.randint(n))
        vo, vd = G.vs[o], G.vs[d]
        dur = router.get_duration(vo["lng"], vo["lat"], vd["lng"], vd["lat"])
        dist = router.get_distance(vo["lng"], vo["lat"], vd["lng"], vd["lat"])
        dd = router.get_distance_duration(vo["lng"], vo["lat"], vd["lng"], vd["lat"], euclidean=True)
        spt = optimUtils.shortestPathTime(G, vo["lat"], vo["lng"], vd["lat"], vd["lng"])
        rec = {"o": o, "d": d, "duration": jf(dur), "shortest_path_time": jf(spt),
               "euclid": [jf(dd[0]), jf(dd[1])]}
        # length along the time-shortest path depends on WHICH tied path is taken; only
        # recorded where it cannot (unique paths, or length proportional to time)
        if unique_paths or "int" in name or "uniform" in name:
            rec["distance"] = jf(dist)
        if unique_paths:
            rec["epath"] = [int(e) for e in G.get_shortest_paths(vo["name"], to=vd["name"], weights="ttmedian",
                                                                 output="epath")[0]]
        out["pairs"].append(rec)
    if unique_paths:
        for rec in out["pairs"][:nrouting]:
            vo, vd = G.vs[rec["o"]], G.vs[rec["d"]]
            r = router.get_routing(vo["lng"], vo["lat"], vd["lng"], vd["lat"])
            out["routes"].append({"o": rec["o"], "d": rec["d"], "route": r})
    for s in [0, n // 3, n - 1]:
        row = G.shortest_paths([G.vs[s]["name"]], [v["name"] for v in G.vs], weights="ttmedian")[0]
        out["rows"].append({"s": s, "dist": [jf(x) for x in row]})
    router2 = RoutingEngine(graph_loc=path, cst_speed=9, seed=11)
    for _ in range(5):
        v = router2.generateRandomLocation()
        out["random_locations"].append({"seed": 11, "index": int(v.index), "lng": v["lng"], "lat": v["lat"]})
    with open(os.path.join(GOLD, "routing_%s.json" % name), "w") as fh:
        json.dump(out, fh)
    return out


class FakeReq(object):
    """Only what createLocations reads from an already-assigned request (lib/optimUtils.py:95-108)."""

    def __init__(self, rid, o, d, Cep, Clp, Ced, Cld):
        self.id, self.olng, self.olat, self.dlng, self.dlat = rid, o[0], o[1], d[0], d[1]
        self.Cep, self.Clp, self.Ced, self.Cld = Cep, Clp, Ced, Cld

    def get_dropoff_time_window(self):
        return [self.Ced, self.Cld]


def cost_golden(name, ncases, seed):
    """computeRoutingCost / minimalDelayPath / findOptimalRouting on random stop sets."""
    path = os.path.join(GOLD, name + ".gml")
    router = RoutingEngine(graph_loc=path, cst_speed=9, seed=7)
    G = router.G
    n = len(G.vs)
    rs = np.random.RandomState(seed)
    random.seed(seed)
    cases = []

    def loc(i):
        return (G.vs[i]["lng"], G.vs[i]["lat"])

    def tt(a, b):
        return G.shortest_paths(G.vs[a]["name"], G.vs[b]["name"], weights="ttmedian")[0][0]

    for case in range(ncases):
        T = float(rs.choice([0, 30, 600]))
        start = int(rs.randint(n))
        t0 = float(rs.choice([0, 0, 7, 12.5]))
        cap = int(rs.choice([1, 2, 4]))
        n_pax = int(rs.randint(0, min(cap, 2) + 1))
        n_wait = int(rs.randint(0, 3 - n_pax))
        slack = float(rs.choice([0.6, 1.0, 3.0]))   # scales the windows: tight cases become infeasible

        def mk(rid, onboard):
            o, d = int(rs.randint(n)), int(rs.randint(n))
            while d == o:
                d = int(rs.randint(n))
            Ts = tt(o, d)
            if onboard:
                Cep = T - float(rs.randint(100, 400))
            else:
                Cep = T - float(rs.randint(0, 60))
            Clp = Cep + 600 * slack
            Ced = Cep + Ts
            Cld = Clp + 1.5 * Ts
            return {"rid": rid, "o": o, "d": d, "Cep": Cep, "Clp": Clp, "Ced": Ced, "Cld": Cld}

        pax = [mk(100 + k, True) for k in range(n_pax)]
        wait = [mk(200 + k, False) for k in range(n_wait)]
        new = mk(1, False)
        veh = {"t": t0, "location": loc(start), "capacity": cap,
               "pax_reqs": [FakeReq(r["rid"], loc(r["o"]), loc(r["d"]), r["Cep"], r["Clp"], r["Ced"], r["Cld"]) for r in pax],
               "wait_reqs": [FakeReq(r["rid"], loc(r["o"]), loc(r["d"]), r["Cep"], r["Clp"], r["Ced"], r["Cld"]) for r in wait]}
        vert = {"rid": new["rid"], "origin": loc(new["o"]), "destination": loc(new["d"]),
                "Cep": new["Cep"] - T, "Clp": new["Clp"], "Ced": new["Ced"], "Cld": new["Cld"]}
        cost, route = optimUtils.findOptimalRouting([vert], G, veh, T)
        # every ordering's cost through the reference's own enumeration + cost function
        stops = optimUtils.createLocations([vert], veh)
        startLoc = optimUtils.StopLocation(veh["location"][0], veh["location"][1])
        perms = optimUtils.permutePassengerOrder(stops, startLoc)
        all_costs = []
        for p in perms:
            c = optimUtils.computeRoutingCost(p, G, T + t0, cap, len(pax))
            c2 = optimUtils.computeRoutingCost(p, G, T + t0, cap, len(pax), infStop=False, infCost=12345.0)
            all_costs.append({"order": [stops.index(s) for s in p[1:]], "cost": jf(c), "cost_nostop": jf(c2)})
        all_costs.sort(key=lambda r: r["order"])
        feas = [r["cost"] for r in all_costs if r["cost"] < 1e6]
        unique = len(feas) > 0 and feas.count(min(feas)) == 1
        # vehicle-less variant (RR edge style): two new requests from the first stop at time T
        other = mk(2, False)
        vert2 = {"rid": other["rid"], "origin": loc(other["o"]), "destination": loc(other["d"]),
                 "Cep": other["Cep"] - T, "Clp": other["Clp"], "Ced": other["Ced"], "Cld": other["Cld"]}
        rr_cost, rr_route = optimUtils.findOptimalRouting((vert, vert2), G, T=T)
        cases.append({"T": T, "start": start, "t0": t0, "capacity": cap, "pax": pax, "wait": wait, "new": new,
                      "other": other, "cost": jf(cost),
                      "route": None if route is None else [[r[0], r[1]] for r in route],
                      "route_unique": bool(unique), "all_costs": all_costs,
                      "rr_cost": jf(rr_cost), "max_feasible_cost": jf(optimUtils.maxFeasibleCost(stops))})
    with open(os.path.join(GOLD, "cost_%s.json" % name), "w") as fh:
        json.dump({"map": name, "cases": cases}, fh)
    return cases


def sim_golden(name, n_veh, n_ticks, req_per_tick, capacity, seed, boundary=False, suffix=""):
    """A few assignment ticks driven through the reference's own Veh / Req / AlonsoMora
    ("enumerate") classes: the corrected-unpack version of Fleet.dispatch_at_time /
    Fleet.optimal_assignment (lib/Agents.py:964-1096; unpack convention of
    optimizationTesting.py:469-472)."""
    path = os.path.join(GOLD, name + ".gml")
    router = RoutingEngine(graph_loc=path, cst_speed=9, seed=7)
    G = router.G
    n = len(G.vs)
    rs = np.random.RandomState(seed)
    T0 = 600   # not 0: a zero-length pickup leg at T == 0 trips `assert reqs[rid].Tp > 0` (lib/Agents.py:176)
    vehs = []
    for vid in range(n_veh):
        v = router.generateRandomLocation()
        vehs.append(Veh(vid, None, K=capacity, T=float(T0), ini_loc=(v["lng"], v["lat"])))
    reqs = []
    optimizer = AlonsoMora({"routing_method": "enumerate"})
    ticks = []
    node_of = {(v["lng"], v["lat"]): v.index for v in G.vs}
    for tick in range(n_ticks):
        T = T0 + tick * RC.INT_ASSIGN
        for veh in vehs:                                    # Fleet.dispatch_at_time, lib/Agents.py:965-1016
            done = veh.move_to_time(T, reqs)
            for (rid, pod, t) in done:
                if pod == 1:
                    reqs[rid].Tp = t
                elif pod == -1:
                    reqs[rid].Td = t
            if len(veh.route) == 0:
                veh.build_route(router, [])
        for k in range(req_per_tick):                       # new requests of this tick
            o, d = int(rs.randint(n)), int(rs.randint(n))
            Tr = T - float(rs.randint(0, RC.INT_ASSIGN))
            if boundary and k == 0:
                # the exact-Ced boundary: a request made NOW at the vertex an idle vehicle stands on, so that the direct
                # trip arrives at Cep + Ts == Ced to the last bit in the reference (same fp64 sums on both sides)
                idle = [v for v in vehs if len(v.route) == 0 and (v.lng, v.lat) in node_of]
                if idle:
                    o = node_of[(idle[tick % len(idle)].lng, idle[tick % len(idle)].lat)]
                    Tr = float(T)
            while d == o:
                d = int(rs.randint(n))
            reqs.append(Req(router, len(reqs), Tr, G.vs[o]["lng"], G.vs[o]["lat"], G.vs[d]["lng"], G.vs[d]["lat"]))
        # snapshot of what the optimizer sees
        snap_v = []
        for veh in vehs:
            pax, wait = veh.get_passenger_requests()
            (lng, lat), t = veh.get_next_node()
            snap_v.append({"vid": veh.id, "K": veh.K, "node": node_of[(lng, lat)], "t": jf(t),
                           "pax": sorted(int(r) for r in pax), "wait": sorted(int(r) for r in wait),
                           "route": [[int(l.rid), int(l.pod)] for l in veh.route]})
        snap_r = [{"rid": r.id, "o": node_of[(r.olng, r.olat)], "d": node_of[(r.dlng, r.dlat)], "Tr": jf(r.Tr),
                   "Cep": jf(r.Cep), "Clp": jf(r.Clp), "Ced": jf(r.Ced), "Cld": jf(r.Cld), "Ts": jf(r.Ts),
                   "Ds": jf(r.Ds), "assigned": r.assigned} for r in reqs]
        with contextlib.redirect_stdout(io.StringIO()):
            rv = optimizer.generatePairwiseGraph(vehs, reqs, G, T)
            rv_edges = []
            for e in rv.es.select(etype="req_veh"):
                a, b = rv.vs[e.source], rv.vs[e.target]
                veh_v, req_v = (a, b) if a["vtype"] == "vehicle" else (b, a)
                rv_edges.append({"vid": veh_v["vid"], "rid": req_v["rid"], "cost": jf(e["cost"]),
                                 "route": [[r[0], r[1]] for r in e["route"]]})
            rr_edges = []
            for e in rv.es.select(etype="req_req"):
                a, b = rv.vs[e.source], rv.vs[e.target]
                rr_edges.append({"rid1": a["rid"], "rid2": b["rid"], "cost": jf(e["cost"])})
            (veh_routes, rejected), meta = optimizer.optimizeAssignment(vehs, reqs, G, T)
        objective = None
        if isinstance(veh_routes, dict):
            # objective of the chosen assignment in terms of the stored edge costs is recomputed by the tests
            pass
        ticks.append({"T": T, "vehs": snap_v, "reqs": snap_r, "rv_edges": rv_edges, "rr_edges": rr_edges,
                      "veh_routes": {str(k): [[r[0], r[1]] for r in v] for k, v in dict(veh_routes).items()}
                      if isinstance(veh_routes, dict) else {},
                      "rejected": sorted(int(r) for r in rejected)})
        if isinstance(veh_routes, dict):                    # Fleet.optimal_assignment, lib/Agents.py:1060-1091
            with contextlib.redirect_stdout(io.StringIO()):
                for vid, route in veh_routes.items():
                    pax, wait = vehs[vid].get_passenger_requests()
                    for rid, _, _, _ in route:
                        if rid not in (pax | wait):
                            reqs[rid].assigned = True
                            reqs[rid].assigned_veh = vid
                    vehs[vid].build_route(router, route, reqs, T)
                for rid in rejected:
                    reqs[rid].assigned = True
    with open(os.path.join(GOLD, "sim_%s_K%d%s.json" % (name, capacity, suffix)), "w") as fh:
        json.dump({"map": name, "capacity": capacity, "ticks": ticks}, fh)
    return ticks


def legacy_golden(tag, name, n_veh, n_ticks, req_per_tick, capacity, seed, tight_every=0, clone_first=0, T0=600.0):
    """The legacy insertion heuristic run by the reference's OWN code: ``Model.insertion_heuristics``,
    ``Model.insert_heuristics`` and ``Model.test_constraints_get_cost`` (lib/Agents.py:1429-1572) are called unbound on
    a namespace that carries what they read from ``self`` (vehs, reqs, queue, rejs, distance_rejs), with the reference's
    real ``Veh`` / ``Req`` objects; vehicles are moved between ticks by ``Veh.move_to_time`` and rebuilt by
    ``Veh.build_route`` exactly as ``Fleet.dispatch_at_time`` does it (lib/Agents.py:964-1016).

    Recorded per inserted request: the fleet and the request table BEFORE the scan, every call of
    test_constraints_get_cost in call order (vehicle, i, j, bound C, flag, cost, viol code), the winner, and AFTER the
    winner's build_route: Cld and pwt of every request (the side effects of lib/Agents.py:1550,1554 and :178-220) and
    c / idle / route of every vehicle."""
    from collections import deque
    from types import SimpleNamespace
    from lib.Agents import Model
    path = os.path.join(GOLD, name + ".gml")
    router = RoutingEngine(graph_loc=path, cst_speed=9, seed=7)
    G = router.G
    n = len(G.vs)
    rs = np.random.RandomState(seed)
    node_of = {(v["lng"], v["lat"]): v.index for v in G.vs}
    vehs = []
    for vid in range(n_veh):
        if clone_first and 0 < vid <= clone_first:
            v = G.vs[node_of[(vehs[0].lng, vehs[0].lat)]]       # identical idle vehicles: equal costs, the LAST one must win
        else:
            v = router.generateRandomLocation()
        vehs.append(Veh(vid, None, K=capacity, T=float(T0), ini_loc=(v["lng"], v["lat"])))
    reqs = []
    ns = SimpleNamespace(vehs=vehs, reqs=reqs, queue=deque(), rejs=[], distance_rejs=[])
    calls = []

    def tcgc(router_, route, veh, req, C):
        out = Model.test_constraints_get_cost(ns, router_, route, veh, req, C)
        i = [k for k, s in enumerate(route) if s[0] == req.id and s[1] == 1][0]
        j = [k for k, s in enumerate(route) if s[0] == req.id and s[1] == -1][0]
        calls.append([int(veh.id), i, j, jf(C), bool(out[0]), None if out[1] is None else jf(out[1]), int(out[2])])
        return out

    def veh_state(veh):
        try:                                                   # lib/Agents.py:1508-1516
            G.vs.find(name=str((veh.lng, veh.lat)))
            node, t0, et0 = node_of[(veh.lng, veh.lat)], 0.0, 0.0
        except ValueError:
            st = veh.route[0].steps[0]
            node, t0, et0 = node_of[(st.geo[1][0], st.geo[1][1])], float(st.t), float(st.et)
        keep = None                                            # what clear_route_but_keep_next_node keeps (lib/Agents.py:319-336)
        if len(veh.route) > 0:
            st = veh.route[0].steps[0]
            keep = {"node": node_of[(st.geo[1][0], st.geo[1][1])], "t": jf(st.t), "et": jf(st.et)}
        return {"vid": int(veh.id), "idle": bool(veh.idle), "n": int(veh.n), "T": jf(veh.T), "K": int(veh.K), "c": jf(veh.c),
                "node": node, "t0": jf(t0), "et0": jf(et0), "keep": keep, "lng": veh.lng, "lat": veh.lat,
                "route": [[int(l.rid), int(l.pod), node_of[(l.tlng, l.tlat)]] for l in veh.route],
                "leg_t": [jf(l.t) for l in veh.route]}

    def req_state(r):
        return {"rid": int(r.id), "o": node_of[(r.olng, r.olat)], "d": node_of[(r.dlng, r.dlat)], "Tr": jf(r.Tr), "Cep": jf(r.Cep),
                "Clp": jf(r.Clp), "Ced": jf(r.Ced), "Cld": jf(r.Cld), "Ts": jf(r.Ts), "pwt": jf(r.pwt), "Tp": jf(r.Tp)}

    insertions = []

    def insert(router_, req, T):
        before_v = [veh_state(v) for v in vehs]
        before_r = [req_state(r) for r in reqs]
        del calls[:]
        with contextlib.redirect_stdout(io.StringIO()):
            ok = Model.insert_heuristics(ns, router_, req, T)
        wins = [c for c in calls if c[4]]
        rec = {"T": jf(T), "rid": int(req.id), "assigned": bool(ok), "vehs": before_v, "reqs": before_r, "calls": [list(c) for c in calls],
               "winner": None if not wins else {"vid": wins[-1][0], "i": wins[-1][1], "j": wins[-1][2], "cost": wins[-1][5]},
               "after_Cld": [jf(r.Cld) for r in reqs], "after_pwt": [jf(r.pwt) for r in reqs],
               "after_vehs": [veh_state(v) for v in vehs]}
        if wins:
            rec["winner"]["dc"] = jf(wins[-1][5] - before_v[wins[-1][0]]["c"])
        insertions.append(rec)
        return ok

    ns.test_constraints_get_cost = tcgc
    ns.insert_heuristics = insert
    for tick in range(n_ticks):
        T = T0 + tick * RC.INT_ASSIGN
        for veh in vehs:                                        # Fleet.dispatch_at_time, lib/Agents.py:965-1016
            with contextlib.redirect_stdout(io.StringIO()):
                done = veh.move_to_time(T, reqs)
            for (rid, pod, t) in done:
                if pod == 1:
                    reqs[rid].Tp = t
                elif pod == -1:
                    reqs[rid].Td = t
            if len(veh.route) == 0:
                veh.build_route(router, [])
        for k in range(req_per_tick):
            o, d = int(rs.randint(n)), int(rs.randint(n))
            while d == o:
                d = int(rs.randint(n))
            Tr = T - float(rs.randint(0, RC.INT_ASSIGN))
            cons = None
            rid = len(reqs)
            if tight_every and rid % tight_every == tight_every - 1:    # a window nobody can meet -> rejection
                Ts = router.get_duration(G.vs[o]["lng"], G.vs[o]["lat"], G.vs[d]["lng"], G.vs[d]["lat"])
                cons = {"Cep": Tr, "Clp": Tr + 5.0, "Ced": Tr + Ts, "Cld": Tr + 5.0 + RC.MAX_DETOUR * Ts}
            req = Req(router, rid, Tr, G.vs[o]["lng"], G.vs[o]["lat"], G.vs[d]["lng"], G.vs[d]["lat"], constraints=cons)
            reqs.append(req)
            ns.queue.append(req)
        Model.insertion_heuristics(ns, router, T)               # the reference's own queue loop (lib/Agents.py:1429-1448)
        assert len(ns.queue) == 0
    with open(os.path.join(GOLD, "legacy_%s.json" % tag), "w") as fh:
        json.dump({"map": name, "capacity": capacity, "T0": T0, "n_veh": n_veh, "insertions": insertions,
                   "rejected": [int(r.id) for r in ns.rejs],
                   "constants": {"MAX_DETOUR": RC.MAX_DETOUR, "COEF_INVEH": RC.COEF_INVEH, "COEF_WAIT": RC.COEF_WAIT,
                                 "COEF_EXTRA_WAIT": RC.COEF_EXTRA_WAIT, "INT_ASSIGN": RC.INT_ASSIGN}}, fh)
    return insertions


LEGACY_CASES = [
    # tag, map, vehicles, ticks, requests per tick, capacity, seed, tight_every, clone_first
    ("g8_int_K4", "g8_int", 4, 8, 4, 4, 21, 0, 0),            # long plans (up to 8+ stops), vehicles between nodes
    ("g20_int_K4", "g20_int", 16, 6, 6, 4, 22, 7, 0),         # a few impossible windows -> rejections
    ("g20_int_K1", "g20_int", 12, 5, 5, 1, 23, 0, 0),         # tight capacity: viol code 1
    ("g8_int_K2_few", "g8_int", 2, 8, 3, 2, 24, 0, 0),        # two vehicles, late pickups / drop-offs
    ("g20_uniform_K4", "g20_uniform", 10, 5, 4, 4, 25, 0, 2), # all-ties map, three identical idle vehicles
    ("g12_float_K2", "g12_float", 5, 5, 4, 2, 26, 0, 0),      # float weights (1e-5 class)
]


def legacy_all():
    for (tag, nm, nv, nt, rpt, K, seed, tight, clone) in LEGACY_CASES:
        ins = legacy_golden(tag, nm, nv, nt, rpt, K, seed, tight_every=tight, clone_first=clone)
        viol = {}
        for x in ins:
            for c in x["calls"]:
                viol[c[6]] = viol.get(c[6], 0) + 1
        print("legacy", tag, len(ins), "insertions,", sum(1 for x in ins if x["assigned"]), "assigned, calls by viol code", sorted(viol.items()),
              "max plan", max(len(v["route"]) for x in ins for v in x["vehs"]))


if __name__ == "__main__":
    make_maps()
    for nm in MAPS:
        r = routing_golden(nm, 150 if "20" in nm else 80, 12, seed=3)
        print("routing", nm, len(r["pairs"]))
    for nm in ["g8_int", "g12_float", "g20_int"]:
        c = cost_golden(nm, 40, seed=5)
        print("cost", nm, len(c), sum(1 for x in c if x["route"] is not None), "feasible")
    for nm, nv, nt, rpt, K in [("g8_int", 4, 5, 3, 4), ("g20_int", 8, 4, 4, 4), ("g20_int", 10, 5, 6, 1),
                               ("g12_float", 5, 4, 4, 2)]:
        t = sim_golden(nm, nv, nt, rpt, K, seed=9)
        print("sim", nm, K, [len(x["rv_edges"]) for x in t], [len(x["veh_routes"]) for x in t])
    for nm, nv, nt, rpt, K in [("g12_float", 6, 5, 3, 2), ("g10_dfloat", 5, 5, 3, 2)]:
        t = sim_golden(nm, nv, nt, rpt, K, seed=13, boundary=True, suffix="b")
        print("sim (boundary)", nm, K, [len(x["rv_edges"]) for x in t], [len(x["veh_routes"]) for x in t])
    legacy_all()
