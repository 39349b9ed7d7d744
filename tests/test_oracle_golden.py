"""The oracle restatement against fixtures produced by the reference's OWN code
(oracle/make_golden.py ran lib/RoutingEngine.py, lib/optimUtils.py,
lib/Optimization.py and lib/Agents.py unmodified on the igraph facade)."""
import numpy as np
import pytest

from conftest import load_golden, oracle_graph, unjf
from oracle import dispatch_oracle as D
from oracle.routing_oracle import OracleRoutingEngine

MAPS = ["g8_int", "g12_float", "g20_int", "g20_uniform", "g10_dfloat"]


@pytest.mark.parametrize("name", MAPS)
def test_routing_matches_reference(name):
    gold = load_golden("routing_%s.json" % name)
    G = oracle_graph(name)
    assert (G.n, G.m) == (gold["n"], gold["m"])
    router = OracleRoutingEngine(G, cst_speed=9, seed=7)
    dist = G.dist_matrix()
    for rec in gold["pairs"]:
        o, d = rec["o"], rec["d"]
        args = (G.lng[o], G.lat[o], G.lng[d], G.lat[d])
        # duration: every tied path has the same time; integer maps are exact, float maps differ only by
        # summation order along a different-but-equal path (none here: paths are unique)
        assert router.get_duration(*args) == pytest.approx(unjf(rec["duration"]), rel=1e-12, abs=0)
        assert dist[o, d] == pytest.approx(unjf(rec["shortest_path_time"]), rel=1e-12, abs=0)
        if "distance" in rec:
            assert router.get_distance(*args) == pytest.approx(unjf(rec["distance"]), rel=1e-12, abs=0)
        eu = router.get_distance_duration(*args, euclidean=True)
        assert [float(eu[0]), float(eu[1])] == [unjf(rec["euclid"][0]), unjf(rec["euclid"][1])]
        if "epath" in rec:
            assert G.epath(o, d) == rec["epath"]           # unique shortest paths: must be the same edges
    for rec in gold["routes"]:
        o, d = rec["o"], rec["d"]
        got = router.get_routing(G.lng[o], G.lat[o], G.lng[d], G.lat[d])
        want = rec["route"]
        assert got["distance"] == want["distance"] and got["duration"] == want["duration"]
        assert got["weight"] == want["weight"] and len(got["steps"]) == len(want["steps"])
        for gs, ws in zip(got["steps"], want["steps"]):
            assert gs["distance"] == ws["distance"] and gs["duration"] == ws["duration"]
            assert gs["mean_duration"] == ws["mean_duration"] and gs["weight"] == ws["weight"]
            assert [list(x["location"]) for x in gs["intersections"]] == [list(x["location"]) for x in ws["intersections"]]
    for row in gold["rows"]:
        want = np.array([unjf(x) for x in row["dist"]])
        np.testing.assert_allclose(dist[row["s"]], want, rtol=1e-12, atol=0)
    r2 = OracleRoutingEngine(G, cst_speed=9, seed=11)
    for rec in gold["random_locations"]:
        v = r2.generateRandomLocation()
        assert (v.index, v["lng"], v["lat"]) == (rec["index"], rec["lng"], rec["lat"])


def _tt(G):
    dist = G.dist_matrix()
    return lambda a, b: float(dist[a, b])


def _case_inputs(G, case):
    def rec(r):
        return {"rid": r["rid"], "o_node": r["o"], "d_node": r["d"], "olng": G.lng[r["o"]], "olat": G.lat[r["o"]],
                "dlng": G.lng[r["d"]], "dlat": G.lat[r["d"]], "Cep": r["Cep"], "Clp": r["Clp"], "Ced": r["Ced"],
                "Cld": r["Cld"]}
    veh = {"vid": 0, "capacity": case["capacity"], "node": case["start"],
           "location": (G.lng[case["start"]], G.lat[case["start"]]), "t": case["t0"],
           "pax_reqs": [rec(r) for r in case["pax"]], "wait_reqs": [rec(r) for r in case["wait"]]}
    return veh, rec(case["new"]), rec(case["other"])


@pytest.mark.parametrize("name", ["g8_int", "g12_float", "g20_int"])
def test_cost_functions_match_reference(name):
    gold = load_golden("cost_%s.json" % name)
    G = oracle_graph(name)
    tt = _tt(G)
    nfeas = 0
    for case in gold["cases"]:
        T = case["T"]
        veh, new, other = _case_inputs(G, case)
        vert = D.request_vertex(new, T)
        stops = D.create_locations([vert], veh)
        start = D.Stop(veh["node"], *veh["location"])
        # computeRoutingCost on EVERY ordering, both infStop modes
        got_orders = D.orderings(stops)
        assert [list(o) for o in got_orders] == [r["order"] for r in case["all_costs"]]
        for order, ref in zip(got_orders, case["all_costs"]):
            seq = [start] + [stops[i] for i in order]
            c = D.compute_routing_cost(seq, tt, T + case["t0"], case["capacity"], len(case["pax"]))
            c2 = D.compute_routing_cost(seq, tt, T + case["t0"], case["capacity"], len(case["pax"]), inf_stop=False,
                                        inf_cost=12345.0)
            assert c == pytest.approx(unjf(ref["cost"]), rel=1e-12, abs=1e-9)
            assert c2 == pytest.approx(unjf(ref["cost_nostop"]), rel=1e-12, abs=1e-9)
        # findOptimalRouting
        cost, route, order = D.find_optimal_routing([vert], tt, veh, T)
        assert cost == pytest.approx(unjf(case["cost"]), rel=1e-12, abs=1e-9)
        assert (route is None) == (case["route"] is None)
        if route is not None:
            nfeas += 1
            if case["route_unique"]:
                assert [[r[0], r[1]] for r in route] == case["route"]
        # vehicle-less variant (RR edge)
        rr_cost, _, _ = D.find_optimal_routing((vert, D.request_vertex(other, T)), tt, None, T)
        assert rr_cost == pytest.approx(unjf(case["rr_cost"]), rel=1e-12, abs=1e-9)
    assert nfeas >= 10


def sim_tick_inputs(G, tick):
    """Golden tick snapshot -> oracle vehicle / request dicts."""
    by_rid = {r["rid"]: r for r in tick["reqs"]}

    def rec(r):
        return {"rid": r["rid"], "o_node": r["o"], "d_node": r["d"], "olng": G.lng[r["o"]], "olat": G.lat[r["o"]],
                "dlng": G.lng[r["d"]], "dlat": G.lat[r["d"]], "Cep": unjf(r["Cep"]), "Clp": unjf(r["Clp"]),
                "Ced": unjf(r["Ced"]), "Cld": unjf(r["Cld"])}
    vehs = [{"vid": v["vid"], "capacity": v["K"], "node": v["node"], "location": (G.lng[v["node"]], G.lat[v["node"]]),
             "t": unjf(v["t"]), "pax_reqs": [rec(by_rid[r]) for r in v["pax"]],
             "wait_reqs": [rec(by_rid[r]) for r in v["wait"]]} for v in tick["vehs"]]
    reqs = [rec(r) for r in tick["reqs"] if r["assigned"] is False]
    return vehs, reqs


def route_cost(G, tt, tick, vid, route):
    """Cost of a golden route [(rid, pod)] for vehicle vid, by the oracle's computeRoutingCost."""
    vehs, reqs = sim_tick_inputs(G, tick)
    veh = [v for v in vehs if v["vid"] == vid][0]
    T = tick["T"]
    new_rids = []
    for rid, pod in route:
        if rid not in new_rids and rid not in [r["rid"] for r in veh["pax_reqs"] + veh["wait_reqs"]]:
            new_rids.append(rid)
    verts = [D.request_vertex([r for r in reqs if r["rid"] == rid][0], T) for rid in new_rids]
    stops = D.create_locations(verts, veh)
    seq = [D.Stop(veh["node"], *veh["location"])]
    for rid, pod in route:
        seq.append([s for s in stops if s.rid == rid and s.pod == pod][0])
    return D.compute_routing_cost(seq, tt, T + veh["t"], veh["capacity"], len(veh["pax_reqs"]))


@pytest.mark.parametrize("fname,name", [("sim_g8_int_K4.json", "g8_int"), ("sim_g20_int_K4.json", "g20_int"),
                                        ("sim_g20_int_K1.json", "g20_int"), ("sim_g12_float_K2.json", "g12_float"),
                                        ("sim_g12_float_K2b.json", "g12_float"), ("sim_g10_dfloat_K2b.json", "g10_dfloat")])
def test_assignment_ticks_match_reference(fname, name):
    gold = load_golden(fname)
    G = oracle_graph(name)
    tt = _tt(G)
    for tick in gold["ticks"]:
        T = tick["T"]
        vehs, reqs = sim_tick_inputs(G, tick)
        (routes, rejected), meta = D.optimize_assignment(vehs, reqs, tt, T)
        # RV edges: same (vehicle, request) set and same costs as the reference's generatePairwiseGraph
        want = {(e["vid"], e["rid"]): unjf(e["cost"]) for e in tick["rv_edges"]}
        got = {k: v[0] for k, v in meta["rv"].items()}
        assert set(got) == set(want)
        for k in want:
            assert got[k] == pytest.approx(want[k], rel=1e-12, abs=1e-9)
        want_rr = {tuple(sorted((e["rid1"], e["rid2"]))): unjf(e["cost"]) for e in tick["rr_edges"]}
        got_rr = {tuple(sorted(k)): v[0] for k, v in meta["rr"].items()}
        assert set(got_rr) == set(want_rr)
        for k in want_rr:
            assert got_rr[k] == pytest.approx(want_rr[k], rel=1e-12, abs=1e-9)
        # assignment: same objective as the reference's ILP solution (ties between optima are unpinned)
        ref_obj = 1e6 * len(tick["rejected"])
        for vid, route in tick["veh_routes"].items():
            c = route_cost(G, tt, tick, int(vid), [tuple(x) for x in route])
            assert c < 1e6
            ref_obj += c
        if not tick["veh_routes"] and not tick["rejected"]:
            assert routes == [] or routes == {}
            continue
        assert meta["objective"] == pytest.approx(ref_obj, rel=1e-12, abs=1e-6)
        assert len(rejected) == len(tick["rejected"])


# ------------------------------------------------------------------ legacy insertion heuristic
LEGACY = ["g8_int_K4", "g20_int_K4", "g20_int_K1", "g8_int_K2_few", "g20_uniform_K4", "g12_float_K2"]


def legacy_tt(G, exact_sums):
    """Travel time a -> b as the reference's get_duration returns it: the fp64 left-to-right sum along the path
    (float maps: unique paths), or the Dijkstra matrix entry (integer maps: every tied path has the same exact sum)."""
    if not exact_sums:
        dist = G.dist_matrix()
        return lambda a, b: float(dist[a, b])
    memo = {}

    def tt(a, b):
        if (a, b) not in memo:
            d = 0
            for e in G.epath(a, b):
                d += float(G.tt[e])
            memo[(a, b)] = d
        return memo[(a, b)]
    return tt


def legacy_state(rec):
    """Oracle-side fleet / request dictionaries of one golden insertion record (state BEFORE the scan)."""
    vehs = []
    for v in rec["vehs"]:
        at_vertex = unjf(v["t0"]) == 0.0 and (v["keep"] is None or v["node"] != v["keep"]["node"] or unjf(v["keep"]["t"]) == 0.0)
        vehs.append({"vid": v["vid"], "n": v["n"], "T": unjf(v["T"]), "K": v["K"], "c": unjf(v["c"]), "node": v["node"],
                     "t0": unjf(v["t0"]), "idle": v["idle"], "route": [tuple(s) for s in v["route"]],
                     "node_pos": v["node"] if at_vertex else None,
                     "keep": None if v["keep"] is None else {"node": v["keep"]["node"], "t": unjf(v["keep"]["t"]),
                                                             "et": unjf(v["keep"]["et"])}})
    reqs = {r["rid"]: {"o_node": r["o"], "d_node": r["d"], "Cep": unjf(r["Cep"]), "Clp": unjf(r["Clp"]), "Cld": unjf(r["Cld"]),
                       "Ts": unjf(r["Ts"]), "Tr": unjf(r["Tr"]), "pwt": unjf(r["pwt"])} for r in rec["reqs"]}
    return vehs, reqs


@pytest.mark.parametrize("tag", LEGACY)
def test_legacy_insertion_matches_reference_call_by_call(tag):
    """Every test_constraints_get_cost call of the reference's scan (vehicle, i, j, bound, flag, cost, viol code), the
    winner, and -- after the winner's build_route -- Cld and pwt of every request and c / route of every vehicle."""
    gold = load_golden("legacy_%s.json" % tag)
    G = oracle_graph(gold["map"])
    is_float = "float" in gold["map"]
    tt = legacy_tt(G, is_float)
    eq = (lambda a, b: a == pytest.approx(b, rel=1e-12, abs=1e-9)) if is_float else (lambda a, b: a == b)
    assert gold["constants"] == {"MAX_DETOUR": D.MAX_DETOUR, "COEF_INVEH": D.COEF_INVEH, "COEF_WAIT": D.COEF_WAIT,
                                 "COEF_EXTRA_WAIT": D.COEF_EXTRA_WAIT, "INT_ASSIGN": 30}
    n_assigned = 0
    for rec in gold["insertions"]:
        vehs, reqs = legacy_state(rec)
        trace = []
        res = D.legacy_insertion_queue(tt, vehs, reqs, [rec["rid"]], unjf(rec["T"]), trace=trace)
        assert len(trace) == len(rec["calls"])
        for got, want in zip(trace, rec["calls"]):
            assert got[:3] == want[:3] and got[4] == want[4] and got[6] == want[6]
            assert eq(got[3], unjf(want[3]))
            assert (got[5] is None) == (want[5] is None) and (got[5] is None or eq(got[5], unjf(want[5])))
        rid, vid, i, j, dc = res[0]
        if rec["winner"] is None:
            assert vid is None and not rec["assigned"]
        else:
            n_assigned += 1
            assert (vid, i, j) == (rec["winner"]["vid"], rec["winner"]["i"], rec["winner"]["j"]) and eq(dc, unjf(rec["winner"]["dc"]))
        for r, want_cld, want_pwt in zip(rec["reqs"], rec["after_Cld"], rec["after_pwt"]):
            assert eq(reqs[r["rid"]]["Cld"], unjf(want_cld)), (tag, rec["rid"], r["rid"])
            assert eq(reqs[r["rid"]]["pwt"], unjf(want_pwt)), (tag, rec["rid"], r["rid"])
        for v, want in zip(vehs, rec["after_vehs"]):
            assert eq(v["c"], unjf(want["c"])) and v["idle"] == want["idle"]
            if not v["idle"] and want["keep"] is not None:
                # the reference's legs: [kept step + first new leg, ...]; same stops in the same order
                assert [tuple(s) for s in want["route"]] == v["route"]
    assert n_assigned == sum(1 for r in gold["insertions"] if r["assigned"]) > 0


@pytest.mark.parametrize("tag", LEGACY)
def test_legacy_queue_carries_state_across_a_tick(tag):
    """The sequential semantics of Model.insertion_heuristics: all requests of one tick processed from the state before
    the FIRST of them must reproduce every later winner (each scan sees the previous winners' routes, c and pwt)."""
    gold = load_golden("legacy_%s.json" % tag)
    G = oracle_graph(gold["map"])
    is_float = "float" in gold["map"]
    tt = legacy_tt(G, is_float)
    eq = (lambda a, b: a == pytest.approx(b, rel=1e-12, abs=1e-9)) if is_float else (lambda a, b: a == b)
    ticks = {}
    for rec in gold["insertions"]:
        ticks.setdefault(rec["T"], []).append(rec)
    multi = 0
    for T, recs in ticks.items():
        vehs, reqs = legacy_state(recs[0])
        for r in recs[-1]["reqs"]:                               # requests that arrived in this tick are all known up front
            if r["rid"] not in reqs:
                reqs[r["rid"]] = {"o_node": r["o"], "d_node": r["d"], "Cep": unjf(r["Cep"]), "Clp": unjf(r["Clp"]),
                                  "Cld": unjf(r["Cld"]), "Ts": unjf(r["Ts"]), "Tr": unjf(r["Tr"]), "pwt": unjf(r["pwt"])}
        keep_after = {}                                          # first step of a freshly built route, as the reference found it
        for rec in recs:
            for v in rec["after_vehs"]:
                if v["keep"] is not None:
                    keep_after.setdefault(v["vid"], (v["keep"]["node"], unjf(v["keep"]["t"])))
        by_pos = {v["node_pos"]: v["vid"] for v in vehs if v["node_pos"] is not None}

        def first_step(a, b, vehs=vehs, keep_after=keep_after):
            vid = [v["vid"] for v in vehs if v["node_pos"] == a and v["keep"] is None and v["route"] and v["route"][0][2] == b]
            return keep_after[vid[0]]
        res = D.legacy_insertion_queue(tt, vehs, reqs, [rec["rid"] for rec in recs], unjf(T), first_step=first_step)
        winners = [r[1] for r in res if r[1] is not None]
        multi += len(winners) - len(set(winners))
        for (rid, vid, i, j, dc), rec in zip(res, recs):
            if rec["winner"] is None:
                assert vid is None
            else:
                assert (vid, i, j) == (rec["winner"]["vid"], rec["winner"]["i"], rec["winner"]["j"]) and eq(dc, unjf(rec["winner"]["dc"]))
        for r, want_cld, want_pwt in zip(recs[-1]["reqs"], recs[-1]["after_Cld"], recs[-1]["after_pwt"]):
            assert eq(reqs[r["rid"]]["Cld"], unjf(want_cld)) and eq(reqs[r["rid"]]["pwt"], unjf(want_pwt))
        for v, want in zip(vehs, recs[-1]["after_vehs"]):
            assert eq(v["c"], unjf(want["c"]))
    assert multi >= 0
