"""Edge cases the reference's semantics imply: directed maps, multi-edges, ragged / maximum-size plans,
empty batches, capacity and window boundaries."""
import ctypes as C

import numpy as np
import pytest

from conftest import OReq, OVeh

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("eval_split")]


def _directed_graph(n_side=9, seed=4):
    """Grid with one-way streets, a few parallel multi-edges and asymmetric weights."""
    from drs_abm_b200.gml import RoadGraphData
    rs = np.random.RandomState(seed)
    idx = np.arange(n_side * n_side).reshape(n_side, n_side)
    src, dst = [], []
    for a, b in list(zip(idx[:, :-1].ravel(), idx[:, 1:].ravel())) + list(zip(idx[:-1, :].ravel(), idx[1:, :].ravel())):
        r = rs.rand()
        if r < 0.6:
            src += [a, b]; dst += [b, a]           # two-way, different weights per direction
        elif r < 0.8:
            src.append(a); dst.append(b)
        else:
            src.append(b); dst.append(a)
    for _ in range(10):                            # parallel edges: the cheaper one must win
        k = rs.randint(len(src))
        src.append(src[k]); dst.append(dst[k])
    m = len(src)
    tt = rs.randint(5, 40, size=m).astype(np.float64)
    n = n_side * n_side
    lng = [float(i // n_side) for i in range(n)]; lat = [float(i % n_side) for i in range(n)]
    return RoadGraphData(n, True, src, dst, {"name": [str((a, b)) for a, b in zip(lng, lat)], "lng": lng, "lat": lat},
                         {"ttmedian": tt, "length": 7.0 * tt})


@pytest.mark.parametrize("algorithm", [0, 1], ids=["fw", "sssp"])
@pytest.mark.parametrize("mode", [0, 1])
def test_directed_map_with_multi_edges(algorithm, mode):
    from conftest import oracle_graph_from_data
    from drs_abm_b200.routing import RoadGraph
    data = _directed_graph()
    G = RoadGraph(data, mode=mode, algorithm=algorithm)
    O = oracle_graph_from_data(data)
    want = O.dist_matrix()
    got = G.dist_rows(0, G.n)
    assert np.array_equal(got, want) and not np.array_equal(got, got.T)       # really asymmetric; inf where unreachable
    rs = np.random.RandomState(0)
    for _ in range(200):
        s, t = int(rs.randint(G.n)), int(rs.randint(G.n))
        ep = G.epath(s, t)
        assert ep == O.epath(s, t)
        cur = s
        for e in ep:                                                          # a directed walk s -> t
            assert int(data.src[e]) == cur
            cur = int(data.dst[e])
        assert (cur == t) if (ep or s == t) else not np.isfinite(want[s, t])
    assert np.array_equal(G.sssp([0, 5, 80]), want[[0, 5, 80]])
    G.close()


def test_directed_map_dispatch_and_insertion():
    """The directed branches of rv_filter / trip costing / insertion scan (no symmetric-row shortcut)."""
    from conftest import oracle_graph_from_data
    from drs_abm_b200 import synth
    from drs_abm_b200.dispatch import AlonsoMora, InsertionTick
    from drs_abm_b200.routing import RoadGraph
    from oracle import dispatch_oracle as D
    data = _directed_graph(10, 9)
    G = RoadGraph(data)
    dist = G.dist_rows(0, G.n)
    assert np.array_equal(dist, oracle_graph_from_data(data).dist_matrix())
    tt = lambda a, b: float(dist[a, b])
    rs = np.random.RandomState(3)
    # --- insertion: random plans (windows wide enough that many candidates are feasible)
    nq = 60
    o = rs.randint(0, G.n, nq); d = (o + 1 + rs.randint(0, G.n - 1, nq)) % G.n
    Ts = dist[o, d]; Ts = np.where(np.isfinite(Ts), Ts, 500.0)
    Cep = 600.0 - rs.randint(0, 200, nq)
    table = {"o": o.astype(np.int32), "d": d.astype(np.int32), "Cep": Cep, "Clp": Cep + 900.0, "Cld": Cep + 900.0 + 2.5 * Ts,
             "Ts": Ts, "Tr": Cep.copy(), "pwt": np.where(rs.rand(nq) < 0.5, 0.0, 50.0)}
    vehs, row = [], 0
    for v in range(12):
        k = int(rs.choice([0, 1, 2]))
        route = []
        for _ in range(k):
            route += [(row, 1, int(o[row])), (row, -1, int(d[row]))]
            row += 1
        vehs.append({"node": int(rs.randint(G.n)), "t0": float(rs.choice([0.0, 3.5])), "n": 0, "T": 600.0, "K": 4, "c": float(rs.randint(0, 50)),
                     "route": route})
    new_rows = np.arange(row, nq, dtype=np.int32)
    tick = InsertionTick(G, vehs, table)
    bv, bi, bj, dc, _ = tick.evaluate(new_rows)
    tick.close()
    ovehs = [dict(v, vid=k, route=list(v["route"])) for k, v in enumerate(vehs)]
    hit = 0
    for k, r in enumerate(new_rows):
        reqs = {q: {c: (int(table[c][q]) if c in "od" else float(table[c][q])) for c in table} for q in range(nq)}
        for q in reqs:
            reqs[q]["o_node"], reqs[q]["d_node"] = reqs[q]["o"], reqs[q]["d"]
        vid, route, odc, _ = D.legacy_insert_heuristics(tt, ovehs, reqs, int(r))
        assert (vid if vid is not None else -1) == int(bv[k])
        if vid is not None:
            hit += 1
            assert odc == dc[k]
            assert [x for x, s in enumerate(route) if s[0] == r] == [int(bi[k]), int(bj[k])]
    assert hit >= 10
    # --- Alonso-Mora RV stage against the oracle
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    reqs = [OReq(q, lng[o[q]], lat[o[q]], lng[d[q]], lat[d[q]], float(Cep[q]), float(Cep[q] + 900), float(Cep[q] + Ts[q]),
                 float(Cep[q] + 900 + 1.5 * Ts[q]), assigned=(q < row)) for q in range(nq)]
    ov, oveh_objs = [], []
    used = 0
    for k, v in enumerate(vehs):
        wait = sorted(set(s[0] for s in v["route"]))
        oveh_objs.append(OVeh(k, 4, (lng[v["node"]], lat[v["node"]]), v["t0"], [], wait))
        rec = lambda q: {"rid": q, "o_node": int(o[q]), "d_node": int(d[q]), "olng": lng[o[q]], "olat": lat[o[q]], "dlng": lng[d[q]],
                         "dlat": lat[d[q]], "Cep": reqs[q].Cep, "Clp": reqs[q].Clp, "Ced": reqs[q].Ced, "Cld": reqs[q].Cld}
        ov.append({"vid": k, "capacity": 4, "node": v["node"], "location": (lng[v["node"]], lat[v["node"]]), "t": v["t0"],
                   "pax_reqs": [], "wait_reqs": [rec(q) for q in oveh_objs[-1].get_passenger_requests()[1]]})
    opt = AlonsoMora({"routing_method": "enumerate", "verbose": False})
    rv = opt.generatePairwiseGraph(oveh_objs, reqs, G, 600.0)
    oreqs = [{"rid": q, "o_node": int(o[q]), "d_node": int(d[q]), "olng": lng[o[q]], "olat": lat[o[q]], "dlng": lng[d[q]], "dlat": lat[d[q]],
              "Cep": reqs[q].Cep, "Clp": reqs[q].Clp, "Ced": reqs[q].Ced, "Cld": reqs[q].Cld} for q in range(row, nq)]
    _, orv, _ = D.rv_graph(ov, oreqs, tt, 600.0, with_rr=False)
    got = {(vi, rv.tick.new_reqs[ri].id): (c, [(s[0], s[1]) for s in route]) for (vi, ri), (c, route) in rv.rv.items()}
    want = {k: (c, [(s[0], s[1]) for s in route]) for k, (c, route, _) in orv.items()}
    rv.tick.close()
    assert got == want and len(got) >= 20
    G.close()


def test_ragged_and_maximum_size_trips():
    from drs_abm_b200.dispatch import MAX_STOPS, Tick
    from drs_abm_b200.routing import RoadGraph
    from oracle import dispatch_oracle as D
    data = _directed_graph(8, 2)
    G = RoadGraph(data)
    dist = G.dist_rows(0, G.n)
    tt = lambda a, b: float(dist[a, b])
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    rs = np.random.RandomState(1)
    nodes = rs.randint(0, G.n, 40)

    def req(q, assigned):
        o, d = int(nodes[2 * q]), int(nodes[2 * q + 1])
        return OReq(q, lng[o], lat[o], lng[d], lat[d], 0.0, 5000.0, 0.0, 20000.0, assigned=assigned)
    # 2 on board + 2 waiting = 6 planned stops + one new request = 8 stops (1 260 orderings)
    reqs = [req(q, True) for q in range(9)] + [req(9, False), req(10, False), req(11, False), req(12, False)]
    veh_busy = OVeh(0, 10, (lng[3], lat[3]), 2.0, [0, 1], [4, 5])
    veh_empty = OVeh(1, 4, (lng[7], lat[7]), 0.0, [], [])
    tick = Tick(G, [veh_busy, veh_empty], reqs, 0.0, pairwise_checks=False)
    idx = {r.id: i for i, r in enumerate(tick.new_reqs)}
    res = tick.trip_eval([(0, [idx[9]]), (1, [idx[9], idx[10], idx[11], idx[12]]), (1, [idx[10]])])
    assert res[0][1] is not None and len(res[0][1]) == 8 and len(res[2][1]) == 2       # ragged task sizes in one launch
    # the 4-request trip on the empty vehicle against the oracle (8 stops: 2520 orderings)
    rec = lambda r: {"rid": r.id, "o_node": G.node_of_location(r.olng, r.olat), "d_node": G.node_of_location(r.dlng, r.dlat),
                     "olng": r.olng, "olat": r.olat, "dlng": r.dlng, "dlat": r.dlat, "Cep": r.Cep, "Clp": r.Clp, "Ced": r.Ced, "Cld": r.Cld}
    oveh = {"vid": 1, "capacity": 4, "node": 7, "location": (lng[7], lat[7]), "t": 0.0, "pax_reqs": [], "wait_reqs": []}
    ocost, oroute, oorder = D.find_optimal_routing([D.request_vertex(rec(reqs[q]), 0.0) for q in (9, 10, 11, 12)], tt, oveh, 0.0)
    assert res[1][0] == ocost and res[1][1] == list(oorder)
    tick.close()
    # 14 planned stops is the maximum the batch format takes; the exact search over 16 stops with wide-open
    # windows (3e11 orderings) is hopeless -- the reference would never return either -- and must end in an
    # error, not in a kernel that runs for hours
    veh_max = OVeh(0, 10, (lng[3], lat[3]), 2.0, [0, 1, 2, 3], [4, 5, 6, 7, 8])
    tick = Tick(G, [veh_max], reqs, 0.0, pairwise_checks=False)
    with pytest.raises(ValueError, match="budget"):
        tick.trip_eval([(0, [idx[9]])])
    tick.close()
    # one more planned stop than the batch format handles: loud error, not a wrong answer
    too_many = OVeh(2, 10, (lng[3], lat[3]), 0.0, [0, 1, 2, 3], [4, 5, 6, 7, 8, 9])
    reqs[9].assigned = True
    with pytest.raises(ValueError):
        Tick(G, [too_many], reqs, 0.0)
    G.close()


def test_empty_batches_through_the_c_abi():
    from drs_abm_b200 import _lib
    from drs_abm_b200.dispatch import InsertionTick
    from drs_abm_b200.routing import RoadGraph
    G = RoadGraph(_directed_graph(6, 1))
    lib = _lib.load()
    t = C.c_void_p()
    assert lib.drs_tick_create(G.handle, C.byref(t)) == 0
    vb, rb = _lib.VehicleBatch(), _lib.RequestBatch()
    assert lib.drs_tick_set_vehicles(t, C.byref(vb)) == 0 and lib.drs_tick_set_requests(t, C.byref(rb)) == 0
    ne, cnt = C.c_int32(-1), np.ones(4, dtype=np.int64)
    assert lib.drs_tick_rv_eval(t, 0.0, 1, C.byref(ne), cnt.ctypes.data_as(_lib._i64p)) == 0
    assert ne.value == 0 and cnt.sum() == 0
    assert lib.drs_tick_trip_eval(t, 0.0, 1, 0, None, None, None, None, None, None) == 0
    assert lib.drs_tick_insertion_eval(t, 0, None, None, None, None, None, None) == _lib.ERR_STATE      # no plans set
    lib.drs_tick_destroy(t)
    tab = {k: np.zeros(1, dtype=np.int32 if k in "od" else np.float64) for k in ("o", "d", "Cep", "Clp", "Cld", "Ts", "Tr", "pwt")}
    tick = InsertionTick(G, [], tab)                                       # no vehicles: every request is rejected
    bv, bi, bj, dc, nc = tick.evaluate([0])
    assert bv[0] == -1 and np.isinf(dc[0]) and nc == 0
    tick.close()
    assert G.path_sums([], [])[0].shape == (0,) and G.matrix_lookup([], []).shape == (0,)
    G.close()
