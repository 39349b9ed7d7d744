"""CUDA (vehicle, pickup slot, drop-off slot) insertion evaluation against the oracle's restatement
of Model.insert_heuristics / test_constraints_get_cost (lib/Agents.py:1451-1572)."""
import copy

import numpy as np
import pytest

from conftest import golden_map_path, oracle_graph

pytestmark = pytest.mark.gpu


def _oracle_reqs(tab):
    return {r: {"o_node": int(tab["o"][r]), "d_node": int(tab["d"][r]), "Cep": float(tab["Cep"][r]),
                "Clp": float(tab["Clp"][r]), "Cld": float(tab["Cld"][r]), "Ts": float(tab["Ts"][r]),
                "Tr": float(tab["Tr"][r]), "pwt": float(tab["pwt"][r])} for r in range(len(tab["o"]))}


def _run_case(name, n_veh, n_new, seed, capacity, plan_stops, tweak=None):
    from drs_abm_b200 import synth
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    from oracle import dispatch_oracle as D
    data = read_gml(golden_map_path(name))
    G = RoadGraph(data)
    dist = G.dist_rows(0, G.n)
    assert np.array_equal(dist, oracle_graph(name).dist_matrix())
    tt = lambda a, b: float(dist[a, b])
    sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, seed, lambda s, t: dist[s, t], plan_stops=plan_stops,
                              capacity=capacity, radius=4)
    if tweak:
        tweak(sc)
    tick = InsertionTick(G, sc["vehs"], dict(sc["table"], Cld=sc["table"]["Cld"]))
    bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
    tick.close()
    assert ncand == sum((len(v["route"]) + 1) * (len(v["route"]) + 2) // 2 for v in sc["vehs"]) * n_new
    ovehs = [{"vid": k, "n": v["n"], "T": v["T"], "K": v["K"], "c": v["c"], "node": v["node"], "t0": v["t0"],
              "route": list(v["route"])} for k, v in enumerate(sc["vehs"])]
    assigned = 0
    for k, row in enumerate(sc["new_rows"]):
        reqs = _oracle_reqs(sc["table"])                       # fresh snapshot: the scan mutates Cld
        vid, route, odc, _ = D.legacy_insert_heuristics(tt, ovehs, reqs, int(row))
        if vid is None:
            assert bv[k] == -1 and np.isinf(dc[k])
            continue
        assigned += 1
        i = [x for x, s in enumerate(route) if s[0] == row and s[1] == 1][0]
        j = [x for x, s in enumerate(route) if s[0] == row and s[1] == -1][0]
        assert (int(bv[k]), int(bi[k]), int(bj[k])) == (vid, i, j)
        assert dc[k] == odc                                    # bit-exact (integer-weighted map, fp64 both sides)
    G.close()
    return assigned


def test_insertion_matches_oracle_mixed_plans():
    assert _run_case("g20_int", 60, 40, 4, 4, (0, 2, 4, 6, 8)) >= 20


def test_insertion_matches_oracle_tight_capacity_and_idle_fleet():
    _run_case("g20_int", 40, 30, 5, 2, (0, 2, 4))
    assert _run_case("g8_int", 12, 25, 6, 1, (0,)) >= 10        # idle vehicles only: one candidate each


def test_insertion_ties_last_wins_and_rejection():
    """Identical vehicles give identical costs: the reference keeps the LAST equal-or-better candidate
    (lib/Agents.py:1471-1476).  Requests nobody can reach in time are rejected."""
    def clones(sc):
        v0 = sc["vehs"][0]
        for v in sc["vehs"][1:]:
            v.update(node=v0["node"], route=list(v0["route"]), n=v0["n"], c=v0["c"])
        sc["table"]["Clp"][sc["new_rows"][:5]] = sc["table"]["Cep"][sc["new_rows"][:5]] - 1e4   # unreachable windows
    _run_case("g20_int", 8, 20, 7, 4, (2,), tweak=clones)


def test_insertion_object_front_end():
    """InsertionHeuristics over Veh / Req-like objects (reference attribute surface)."""
    from drs_abm_b200.dispatch import InsertionHeuristics
    from drs_abm_b200.routing import RoutingEngine
    router = RoutingEngine(graph_loc=golden_map_path("g8_int"), cst_speed=9, seed=1)
    G = router.G

    class R(object):
        pass

    class V(object):
        pass
    a, b, c = G.vs[0], G.vs[G.n - 1], G.vs[9]
    req = R()
    req.id, req.olng, req.olat, req.dlng, req.dlat = 0, a["lng"], a["lat"], b["lng"], b["lat"]
    req.Ts = router.get_duration(a["lng"], a["lat"], b["lng"], b["lat"])
    req.Tr = req.Cep = 0.0; req.Clp = 600.0; req.Cld = 600.0 + 1.5 * req.Ts; req.pwt = 0.0; req.assigned = False
    near, far = V(), V()
    for v, node, vid in ((far, b, 0), (near, c, 1)):
        v.id, v.idle, v.route, v.c, v.n, v.T, v.K, v.lng, v.lat = vid, True, [], 0.0, 0, 0.0, 4, node["lng"], node["lat"]
    ins = InsertionHeuristics([far, near], [req])
    ins.queue.append(req)
    ins.insertion_heuristics(router, 0)
    veh, route, dc = ins.last_choice
    assert veh is near and [(s[0], s[1]) for s in route] == [(0, 1), (0, -1)] and req.assigned and not ins.rejs
    t_pick = router.get_duration(c["lng"], c["lat"], a["lng"], a["lat"])
    # empty until the pickup (n * dt = 0), wait term t * COEF_WAIT at the pickup, then one rider for Ts seconds
    assert dc == t_pick * 1.5 + req.Ts * 1.0


def test_insertion_object_front_end_with_vehicle_between_nodes():
    """A vehicle in the middle of a link is not a graph vertex: the scan starts at the next downstream node
    with the remaining step time added (lib/Agents.py:1508-1516), and a busy vehicle's legs become its plan."""
    from drs_abm_b200.dispatch import InsertionHeuristics
    from drs_abm_b200.routing import RoutingEngine
    from oracle import dispatch_oracle as D
    router = RoutingEngine(graph_loc=golden_map_path("g20_int"), cst_speed=9, seed=1)
    G = router.G
    dist = G.dist_rows(0, G.n)
    tt = lambda a, b: float(dist[a, b])

    class Obj(object):
        pass

    def mkreq(rid, o, d, Cep, pwt=0.0):
        r = Obj()
        r.id, r.olng, r.olat, r.dlng, r.dlat = rid, G.vs[o]["lng"], G.vs[o]["lat"], G.vs[d]["lng"], G.vs[d]["lat"]
        r.Ts = tt(o, d); r.Tr = r.Cep = Cep; r.Clp = Cep + 600.0; r.Cld = r.Clp + 1.5 * r.Ts; r.pwt = pwt; r.assigned = False
        r._o, r._d = o, d
        return r
    reqs = [mkreq(0, 45, 250, 100.0, pwt=80.0), mkreq(1, 46, 66, 90.0), mkreq(2, 44, 104, 110.0)]
    veh = Obj()
    veh.id, veh.idle, veh.c, veh.n, veh.T, veh.K = 0, False, 37.5, 0, 120.0, 4
    veh.lng, veh.lat = G.vs[42]["lng"] + 1e-4, G.vs[42]["lat"]          # between vertices 42 and 43
    step = Obj(); step.geo = [[veh.lng, veh.lat], [G.vs[43]["lng"], G.vs[43]["lat"]]]; step.t = 11.5
    legs = []
    for (rid, pod, node) in [(0, 1, 45), (0, -1, 250)]:
        leg = Obj(); leg.rid, leg.pod, leg.tlng, leg.tlat, leg.steps = rid, pod, G.vs[node]["lng"], G.vs[node]["lat"], [step]
        legs.append(leg)
    veh.route = legs
    idle = Obj()
    idle.id, idle.idle, idle.c, idle.n, idle.T, idle.K, idle.route = 1, True, 0.0, 0, 120.0, 4, []
    idle.lng, idle.lat = G.vs[399]["lng"], G.vs[399]["lat"]
    ins = InsertionHeuristics([veh, idle], reqs)
    for new in (reqs[1], reqs[2]):
        assert ins.insert_heuristics(router, new, 120.0)
        wveh, wroute, wdc = ins.last_choice
        ovehs = [{"vid": 0, "n": 0, "T": 120.0, "K": 4, "c": 37.5, "node": 43, "t0": 11.5, "route": [(0, 1, 45), (0, -1, 250)]},
                 {"vid": 1, "n": 0, "T": 120.0, "K": 4, "c": 0.0, "node": 399, "t0": 0.0, "route": []}]
        oreqs = {r.id: {"o_node": r._o, "d_node": r._d, "Cep": r.Cep, "Clp": r.Clp, "Cld": r.Cld, "Ts": r.Ts, "Tr": r.Tr, "pwt": r.pwt}
                 for r in reqs}
        vid, route, odc, _ = D.legacy_insert_heuristics(tt, ovehs, oreqs, new.id)
        assert wveh.id == vid and wdc == odc
        assert [(s[0], s[1]) for s in wroute] == [(s[0], s[1]) for s in route]


def test_c4_full_size_tick_sampled_against_oracle():
    """BASELINE config C4 at full size (10 000 vehicles x 2 000 pending requests on the 50k-node C3 map, 3.9e8
    candidates): the candidate count is the closed form, every reported winner is a valid (vehicle, i < j <= L+1)
    with a finite cost, re-evaluating a request alone gives the same answer as inside the batch (requests are
    independent), and sampled requests (two with a winner, one
    without if there is any) equal the oracle's sequential scan over ALL 10 000 vehicles."""
    from drs_abm_b200 import synth
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.routing import RoadGraph
    from oracle import dispatch_oracle as D
    data = synth.config_graph("C3")
    G = RoadGraph(data).build()
    n_veh, n_new = 10000, 2000
    sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
    tick = InsertionTick(G, sc["vehs"], sc["table"])
    bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
    L = np.array([len(v["route"]) for v in sc["vehs"]], dtype=np.int64)
    assert ncand == int(((L + 1) * (L + 2) // 2).sum()) * n_new
    won = bv >= 0
    assert won.sum() > n_new // 2
    assert np.all(np.isfinite(dc[won])) and np.all(np.isinf(dc[~won]))
    assert np.all(bi[won] >= 0) and np.all(bi[won] < bj[won]) and np.all(bj[won] <= L[bv[won]] + 1)
    pick = np.random.RandomState(9).choice(n_new, 64, replace=False)
    bv2, bi2, bj2, dc2, _ = tick.evaluate(sc["new_rows"][pick])
    assert np.array_equal(bv2, bv[pick]) and np.array_equal(bi2, bi[pick]) and np.array_equal(bj2, bj[pick]) and np.array_equal(dc2, dc[pick])
    tick.close()
    rows = {}

    def tt(a, b):
        r = rows.get(a)
        if r is None:
            r = rows[a] = G.dist_rows(int(a), 1)[0]
        return float(r[b])
    ovehs = [{"vid": k, "n": v["n"], "T": v["T"], "K": v["K"], "c": v["c"], "node": v["node"], "t0": v["t0"],
              "route": list(v["route"])} for k, v in enumerate(sc["vehs"])]
    sampled = [int(np.flatnonzero(won)[7]), int(np.flatnonzero(won)[-3])] + [int(x) for x in np.flatnonzero(~won)[:1]]
    for k in sampled:
        row = int(sc["new_rows"][k])
        vid, route, odc, _ = D.legacy_insert_heuristics(tt, ovehs, _oracle_reqs(sc["table"]), row)
        if not won[k]:
            assert vid is None
            continue
        assert vid is not None
        i = [x for x, s in enumerate(route) if s[0] == row and s[1] == 1][0]
        j = [x for x, s in enumerate(route) if s[0] == row and s[1] == -1][0]
        assert (int(bv[k]), int(bi[k]), int(bj[k])) == (vid, i, j) and dc[k] == odc
    G.close()


# ------------------------------------------------------------------ fixtures produced by the reference's own lib/Agents.py
LEGACY = ["g8_int_K4", "g20_int_K4", "g20_int_K1", "g8_int_K2_few", "g20_uniform_K4", "g12_float_K2"]


def _golden_tick(G, rec):
    """InsertionTick inputs of one golden insertion record (fleet and request table BEFORE the scan)."""
    from conftest import unjf
    vehs = []
    for v in rec["vehs"]:
        vehs.append({"node": v["node"], "t0": unjf(v["t0"]), "n": v["n"], "T": unjf(v["T"]), "K": v["K"], "c": unjf(v["c"]),
                     "route": [] if v["idle"] else [(s[0], s[1], s[2]) for s in v["route"]],
                     "keep": None if v["keep"] is None else (v["keep"]["node"], unjf(v["keep"]["t"]), unjf(v["keep"]["et"])),
                     "pos": v["node"] if unjf(v["t0"]) == 0.0 else None})
    reqs = rec["reqs"]
    assert [r["rid"] for r in reqs] == list(range(len(reqs)))
    table = {"o": np.array([r["o"] for r in reqs], dtype=np.int32), "d": np.array([r["d"] for r in reqs], dtype=np.int32)}
    for k in ("Cep", "Clp", "Cld", "Ts", "Tr", "pwt"):
        table[k] = np.array([unjf(r[k]) for r in reqs], dtype=np.float64)
    return vehs, table


@pytest.mark.parametrize("tag", LEGACY)
def test_insertion_matches_the_references_own_scan(tag):
    """Every insertion the reference's Model.insert_heuristics performed in the fixture (lib/Agents.py:1451-1572 run
    unmodified, oracle/make_golden.py): same winning vehicle, same (i, j), same cost increase -- bit-exact on the
    integer-weighted maps, 1e-5 relative on the float map (fp32 matrix)."""
    from conftest import load_golden, unjf
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    gold = load_golden("legacy_%s.json" % tag)
    G = RoadGraph(read_gml(golden_map_path(gold["map"])))
    is_float = "float" in gold["map"]
    assigned = 0
    for rec in gold["insertions"]:
        vehs, table = _golden_tick(G, rec)
        tick = InsertionTick(G, vehs, table)
        bv, bi, bj, dc, _ = tick.evaluate([rec["rid"]])
        tick.close()
        w = rec["winner"]
        if w is None:
            assert bv[0] == -1 and np.isinf(dc[0]), (tag, rec["rid"])
            continue
        assigned += 1
        assert (int(bv[0]), int(bi[0]), int(bj[0])) == (w["vid"], w["i"], w["j"]), (tag, rec["rid"])
        if is_float:
            assert dc[0] == pytest.approx(unjf(w["dc"]), rel=1e-5)
        else:
            assert dc[0] == unjf(w["dc"])
    assert assigned > 0
    G.close()


@pytest.mark.parametrize("tag", LEGACY)
def test_insertion_queue_matches_the_references_own_ticks(tag):
    """Model.insertion_heuristics tick by tick (lib/Agents.py:1429-1448 run unmodified, oracle/make_golden.py): ALL requests
    of a tick go through drs_tick_insertion_queue from the state before the first of them; every winner, (i, j) and cost
    increase must be the reference's -- later requests see the earlier winners' routes, veh.c and pwt -- and after the
    tick veh.c, the routes and pwt on the device equal what the reference's build_route left behind."""
    from conftest import load_golden, unjf
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    gold = load_golden("legacy_%s.json" % tag)
    G = RoadGraph(read_gml(golden_map_path(gold["map"])))
    is_float = "float" in gold["map"]
    eq = (lambda a, b: a == pytest.approx(b, rel=1e-5, abs=1e-6)) if is_float else (lambda a, b: a == b)
    ticks = {}
    for rec in gold["insertions"]:
        ticks.setdefault(rec["T"], []).append(rec)
    tick = None
    repeats = 0
    for T, recs in ticks.items():
        vehs, table = _golden_tick(G, recs[0])      # the tick's arrivals were all created before the queue ran
        assert len(recs[0]["reqs"]) == len(recs[-1]["reqs"])
        if tick is None:
            tick = InsertionTick(G, vehs, table)
        else:
            tick.set_plans(vehs, table)
        rids = [rec["rid"] for rec in recs]
        bv, bi, bj, dc, ncand = tick.queue(rids, unjf(T))
        winners = []
        for k, rec in enumerate(recs):
            w = rec["winner"]
            if w is None:
                assert bv[k] == -1 and np.isinf(dc[k]), (tag, rec["rid"])
            else:
                assert (int(bv[k]), int(bi[k]), int(bj[k])) == (w["vid"], w["i"], w["j"]), (tag, rec["rid"])
                assert eq(dc[k], unjf(w["dc"])), (tag, rec["rid"])
                winners.append(w["vid"])
        repeats += len(winners) - len(set(winners))
        # candidates: every request scans every plan as it is at its turn
        want = sum(sum((len(v["route"]) + 1) * (len(v["route"]) + 2) // 2 if not v["idle"] else 1 for v in rec["vehs"]) for rec in recs)
        assert ncand == want
        state = tick.fetch_plans()
        after = recs[-1]
        for v, want_v in enumerate(after["after_vehs"]):
            if want_v["idle"] and not want_v["route"]:
                continue
            if not want_v["idle"]:
                assert eq(state["c"][v], unjf(want_v["c"])), (tag, T, v)
                assert state["routes"][v] == [(s[0], s[1]) for s in want_v["route"]], (tag, T, v)
        touched = set(r for v in set(winners) for (r, pod) in state["routes"][v] if pod == 1)
        for r in touched:
            assert eq(state["pwt"][r], unjf(after["after_pwt"][r])), (tag, T, r)
        # Req.Cld of EVERY request as the reference's scans left it (side effect of lib/Agents.py:1550,1554)
        for r, want_cld in enumerate(after["after_Cld"]):
            assert eq(state["Cld"][r], unjf(want_cld)), (tag, T, r)
    tick.close()
    G.close()


def test_insertion_queue_matches_oracle_queue_on_a_fleet():
    """A 60-vehicle fleet with mixed plans and a queue of 40 requests: drs_tick_insertion_queue against the oracle's
    restatement of the sequential loop (which tests/test_oracle_golden.py pins on the reference itself)."""
    from drs_abm_b200 import synth
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    from oracle import dispatch_oracle as D
    name = "g20_int"
    data = read_gml(golden_map_path(name))
    G = RoadGraph(data)
    O = oracle_graph(name)
    dist = G.dist_rows(0, G.n)
    tt = lambda a, b: float(dist[a, b])
    sc = synth.fleet_scenario(data, 60, 40, 600.0, 11, lambda s, t: dist[s, t], plan_stops=(0, 0, 2, 4, 6), capacity=4, radius=5)
    vehs = []
    for v in sc["vehs"]:
        moving = v["t0"] != 0.0 or len(v["route"]) > 0
        vehs.append(dict(v, keep=(v["node"], v["t0"], v["t0"]) if moving else None, pos=v["node"] if v["t0"] == 0.0 else None))
    tick = InsertionTick(G, vehs, sc["table"])
    bv, bi, bj, dc, _ = tick.queue(sc["new_rows"], 600.0)
    state = tick.fetch_plans()
    tick.close()
    ovehs = [{"vid": k, "n": v["n"], "T": v["T"], "K": v["K"], "c": v["c"], "node": v["node"], "t0": v["t0"], "idle": len(v["route"]) == 0,
              "route": list(v["route"]), "node_pos": v["pos"],
              "keep": None if v["keep"] is None else {"node": v["keep"][0], "t": v["keep"][1], "et": v["keep"][2]}}
             for k, v in enumerate(vehs)]
    oreqs = _oracle_reqs(sc["table"])

    def first_step(a, b):
        path = O.epath(a, b)
        if not path:
            return (a, 0.0)
        e = path[0]
        other = int(O.dst[e]) if int(O.src[e]) == a else int(O.src[e])
        return (other, float(O.tt[e]))
    res = D.legacy_insertion_queue(tt, ovehs, oreqs, [int(r) for r in sc["new_rows"]], 600.0, first_step=first_step)
    won = [r[1] for r in res if r[1] is not None]
    assert len(won) >= 15 and len(won) > len(set(won))          # some vehicle wins twice: the carried state matters
    for k, (rid, vid, i, j, odc) in enumerate(res):
        assert (int(bv[k]), int(bi[k]), int(bj[k])) == ((vid, i, j) if vid is not None else (-1, -1, -1)), k
        assert dc[k] == odc
    for v, ov in enumerate(ovehs):
        assert state["c"][v] == ov["c"] and state["routes"][v] == [(s[0], s[1]) for s in ov["route"]]
    for r, rq in oreqs.items():
        assert state["pwt"][r] == rq["pwt"]
    G.close()


@pytest.mark.parametrize("tag", ["g8_int_K4", "g20_int_K4", "g20_uniform_K4"])
def test_single_scans_leave_the_references_cld(tag):
    """One request at a time (the granularity of Model.insert_heuristics): winner and the Req.Cld value of EVERY request
    after the scan -- candidates that were merely evaluated leave their mark too (lib/Agents.py:1550,1554)."""
    from conftest import load_golden, unjf
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    gold = load_golden("legacy_%s.json" % tag)
    G = RoadGraph(read_gml(golden_map_path(gold["map"])))
    tick = None
    changed = 0
    for rec in gold["insertions"]:
        vehs, table = _golden_tick(G, rec)
        if tick is None:
            tick = InsertionTick(G, vehs, table)
        else:
            tick.set_plans(vehs, table)
        bv, bi, bj, dc, _ = tick.queue([rec["rid"]], unjf(rec["T"]), track_cld=True)
        w = rec["winner"]
        assert (int(bv[0]), int(bi[0]), int(bj[0])) == ((w["vid"], w["i"], w["j"]) if w else (-1, -1, -1))
        cld = tick.fetch_plans()["Cld"]
        want = np.array([unjf(x) for x in rec["after_Cld"]])
        assert np.array_equal(cld, want), (tag, rec["rid"], np.flatnonzero(cld != want))
        changed += int(np.sum(want != table["Cld"]))
    assert changed > 0
    tick.close()
    G.close()


def test_steady_state_ticks_allocate_nothing():
    """A tick handle reloaded with a fleet of the same size (set_plans) and evaluated again -- frozen plans and the
    sequential queue -- makes no device allocation after the first round (drs_alloc_count)."""
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    data = read_gml(golden_map_path("g20_int"))
    G = RoadGraph(data)
    dist = G.dist_rows(0, G.n)
    lib = _lib.load()
    tick = None
    first = None
    for rnd in range(4):
        sc = synth.fleet_scenario(data, 50, 30, 600.0, 20 + rnd, lambda s, t: dist[s, t], plan_stops=(0, 2, 4), radius=5)
        if tick is None:
            tick = InsertionTick(G, sc["vehs"], sc["table"])
        else:
            tick.set_plans(sc["vehs"], sc["table"])
        tick.evaluate(sc["new_rows"])
        tick.queue(sc["new_rows"], 600.0)
        tick.fetch_plans()
        if rnd == 0:
            first = lib.drs_alloc_count()
    assert lib.drs_alloc_count() == first, "a steady-state tick allocated device memory"
    tick.close()
    G.close()
