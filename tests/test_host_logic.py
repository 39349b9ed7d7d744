"""CPU-only checks: the C-ABI library loads and exports every symbol include/drs_b200.h declares
(no compute calls), host-side formats and helpers, and that the product refuses to run without its
CUDA extension (no CPU fallback)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "drs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build()
    assert os.path.exists(path)
    from drs_abm_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "declared in include/drs_b200.h but not exported: %s" % name
        assert name in _lib.SIGNATURES, "exported but not bound in drs_abm_b200/_lib.py: %s" % name
    exported = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    for name in re.findall(r" T (drs_[a-z0-9_]+)", exported):
        assert name in declared, "exported without a declaration in include/drs_b200.h: %s" % name
    assert lib.drs_version() == 100 and lib.drs_launch_count() == 0      # nothing has run on a GPU here


def test_sass_is_sm100a_only():
    from drs_abm_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_when_library_is_missing(tmp_path, monkeypatch):
    from drs_abm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(_lib.DrsError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "drs_abm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("oracle/", ""), fn
            assert "/root/reference" not in src.replace("/root/reference lib", "").replace("/root/reference/.gitignore", ""), fn


def test_gml_roundtrip_and_independent_reader(tmp_path):
    from drs_abm_b200 import gml, synth
    from oracle.routing_oracle import OracleGraph
    data = synth.grid_city(9, 3, 0.9, "float")
    p = str(tmp_path / "g.gml")
    gml.write_gml(p, data)
    back = gml.read_gml(p)
    assert (back.n, back.m, back.directed) == (data.n, data.m, False)
    assert np.array_equal(back.src, data.src) and np.array_equal(back.dst, data.dst)
    for k in ("length", "ttmedian", "slnshift", "slnmean", "slnsd"):
        assert np.array_equal(back.eattrs[k], data.eattrs[k])          # repr() round-trips doubles exactly
    assert back.vattrs["name"] == data.vattrs["name"] and back.vattrs["lng"] == data.vattrs["lng"]
    assert all(isinstance(x, float) for x in back.vattrs["lng"])       # igraph hands numeric attrs back as floats
    O = OracleGraph.from_gml(p)                                          # networkx reader agrees
    assert (O.n, O.m) == (data.n, data.m) and O.names == data.vattrs["name"]
    assert np.array_equal(O.tt, data.eattrs["ttmedian"]) and np.array_equal(O.src, data.src)


def test_synthetic_configs_are_deterministic_and_connected():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from drs_abm_b200 import synth
    a, b = synth.config_graph("C2"), synth.config_graph("C2")
    assert a.n == b.n == 5040 and np.array_equal(a.src, b.src) and np.array_equal(a.eattrs["ttmedian"], b.eattrs["ttmedian"])
    adj = coo_matrix((np.ones(a.m), (a.src, a.dst)), shape=(a.n, a.n))
    assert connected_components(adj, directed=False)[0] == 1
    c1 = synth.config_graph("C1")
    assert (c1.n, c1.m) == (400, 760) and c1.eattrs["ttmedian"].min() >= 20 and c1.eattrs["ttmedian"].max() <= 60
    assert np.array_equal(c1.eattrs["length"], 10 * c1.eattrs["ttmedian"])
    assert len(set(c1.vattrs["name"])) == c1.n and c1.vattrs["name"][0] == str((c1.vattrs["lng"][0], c1.vattrs["lat"][0]))
    o, d = synth.random_od_pairs(400, 1000, 1)
    assert (o != d).all()


def test_node_identity_rules():
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import RoadGraph
    G = RoadGraph(synth.config_graph("C1"))
    v = G.vs[37]
    assert G.node_index(v["name"]) == 37
    assert G.node_index(str((v["lng"], v["lat"]))) == 37
    assert G.node_index(str((np.float64(v["lng"]), np.float64(v["lat"])))) == 37      # numpy-2 scalar repr
    assert G.node_of_location(np.float64(v["lng"]), v["lat"]) == 37
    assert G.vs.find(name=v["name"]).index == 37 and G.vs.find(v["name"])["lat"] == v["lat"]
    with pytest.raises(ValueError):
        G.vs.find(name="(0.0, 0.0)")
    with pytest.raises(ValueError):
        G.node_index("nonsense")
    e = G.es[5]
    assert (e.source, e.target) == (int(G._src[5]), int(G._dst[5])) and e["length"] == 10 * e["ttmedian"]
    assert len(G.es) == 760 and len(G.vs) == 400 and len(G.es["length"]) == 760
    e["ttmedian"] = 99.0
    assert G.es[5]["ttmedian"] == 99.0
    with pytest.raises(ValueError):
        G.es["ttmedian"] = [1.0, 2.0]


def test_rng_parity_of_generate_random_location():
    """lib/RoutingEngine.py:228-229 ``rs.choice(G.vs)`` consumes exactly one randint(0, n)."""
    n = 400
    a, b = np.random.RandomState(11), np.random.RandomState(11)
    seq = list(range(n))
    for _ in range(20):
        assert a.choice(seq) == int(b.randint(0, n))


def test_set_packing_solver_matches_the_ilp_objective():
    from drs_abm_b200.dispatch import solve_set_packing
    from oracle.dispatch_oracle import assign_ilp
    rs = np.random.RandomState(0)
    for case in range(25):
        nveh, nreq = int(rs.randint(1, 7)), int(rs.randint(1, 8))
        trips = []
        for v in range(nveh):
            for r in range(nreq):
                if rs.rand() < 0.5:
                    trips.append((v, (r,), float(rs.randint(0, 500))))
            if case % 2:                                        # multi-request trips -> general ILP path
                for _ in range(int(rs.randint(0, 4))):
                    k = int(rs.randint(2, min(4, nreq) + 1)) if nreq >= 2 else 1
                    members = tuple(sorted(rs.choice(nreq, size=k, replace=False).tolist()))
                    if len(members) > 1 and (v, members) not in [(t[0], t[1]) for t in trips]:
                        trips.append((v, members, float(rs.randint(0, 900))))
        chosen, ignored, obj = solve_set_packing(nveh, nreq, [(t[0], t[1]) for t in trips], np.array([t[2] for t in trips]))
        _, orej, oobj = assign_ilp(list(range(nveh)), list(range(nreq)), trips)
        assert obj == pytest.approx(oobj, abs=1e-6)
        served = [r for k in chosen for r in trips[k][1]]
        assert len(served) == len(set(served)) and sorted(served + ignored) == list(range(nreq))
        assert len(set(trips[k][0] for k in chosen)) == len(chosen)
        assert len(ignored) == len(orej)


def test_draw_link_travel_time_matches_the_oracle_stream():
    """lib/RoutingEngine.py:56-79 stays on the host (RNG-order sensitive): same seed, same draws as the restatement."""
    from conftest import golden_map_path, oracle_graph_from_data
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph, RoutingEngine
    from oracle.routing_oracle import OracleRoutingEngine
    data = read_gml(golden_map_path("g12_float"))
    ours = RoutingEngine(graph=RoadGraph(data), cst_speed=9, seed=123)
    ref = OracleRoutingEngine(oracle_graph_from_data(data), cst_speed=9, seed=123)
    for e in [0, 5, 17, 3, 3, 40]:
        assert ours.draw_link_travel_time(e) == ref.draw_link_travel_time(e)
        assert ours.draw_link_travel_time(e, use_uncertainty=False) == data.eattrs["ttmedian"][e]
    assert ours.generateRandomLocation().index == ref.generateRandomLocation().index
    d = ours.get_distance_duration(-71.1, 42.3, -71.09, 42.31, euclidean=True)
    assert d == ref.get_distance_duration(-71.1, 42.3, -71.09, 42.31, euclidean=True) and d[1] == d[0] / 9


def test_locality_numbering_is_a_permutation_with_fallback():
    """Device numbering of the vertices (routing.locality_order): a permutation in every case; a regular grid gets its
    exact Z-curve; maps without usable coordinates fall back to reverse Cuthill-McKee; the translation helpers of
    RoadGraph invert each other and reject out-of-range ids like igraph does."""
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import RoadGraph, locality_order
    data = synth.config_graph("C1")
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    order = locality_order(data.n, data.src, data.dst, lng, lat)
    assert sorted(order.tolist()) == list(range(data.n))
    assert order[:8].tolist() == [0, 20, 1, 21, 40, 60, 41, 61]          # Z-curve over the 20 x 20 grid (node = 20 * i + j)
    for bad in ((None, None), (["a"] * data.n, lat), ([0.0] * data.n, [0.0] * data.n), (lng[:5], lat[:5]),
                ([float("nan")] * data.n, lat)):
        o = locality_order(data.n, data.src, data.dst, *bad)
        assert sorted(o.tolist()) == list(range(data.n))
    assert locality_order(3, [], [], None, None).tolist() == [0, 1, 2]
    G = RoadGraph(data)
    ids = np.array([0, 5, 399, 17], dtype=np.int32)
    dev = G.dev_ids(ids)
    assert dev.dtype == np.int32 and sorted(G.dev_ids(np.arange(data.n)).tolist()) == list(range(data.n))
    rows = np.arange(2 * 512, dtype=np.float64).reshape(2, 512)          # two padded "device rows": value = device column
    assert np.array_equal(G.host_cols(rows)[0, ids], dev.astype(np.float64)) and G.host_cols(rows).shape == (2, data.n)
    for bad_ids in ([-1], [data.n], [0, 400]):
        with pytest.raises(ValueError):
            G.dev_ids(bad_ids)
    Gi = RoadGraph(data, numbering="input")
    assert np.array_equal(Gi.dev_ids(ids), ids) and np.array_equal(Gi.host_cols(rows), rows[:, :data.n])
    with pytest.raises(ValueError):
        RoadGraph(data, numbering="sorted")


def test_rr_prefilter_equals_the_references_double_loop():
    """The vectorised crow-flies pre-filter of the R x R stage against the reference's pair-by-pair expression
    (lib/Optimization.py:148-151), including pairs placed exactly on the threshold."""
    import math
    from drs_abm_b200 import dispatch

    class R(object):
        pass
    rs = np.random.RandomState(3)
    reqs = []
    for k in range(120):
        r = R()
        r.olat, r.olng = 42.3 + 0.05 * rs.rand(), -71.1 + 0.05 * rs.rand()
        r.Clp = 600.0 + 400.0 * rs.rand()
        reqs.append(r)
    T = 300.0
    # put a few pairs exactly on the boundary: Clp := T + d / CST_SPEED
    for a, b in ((0, 1), (5, 9), (17, 80)):
        d = dispatch.great_circle_distance(reqs[a].olat, reqs[a].olng, reqs[b].olat, reqs[b].olng)
        reqs[a].Clp = T + d / dispatch.CST_SPEED
        reqs[b].Clp = reqs[a].Clp + 50.0
    want = []
    for a in range(len(reqs)):
        for b in range(a + 1, len(reqs)):
            t = T + dispatch.great_circle_distance(reqs[a].olat, reqs[a].olng, reqs[b].olat, reqs[b].olng) / dispatch.CST_SPEED
            if t > reqs[a].Clp or t > reqs[b].Clp:
                continue
            want.append((a, b))
    got = dispatch.rr_prefilter(reqs, T)
    assert got == want and 0 < len(want) < 120 * 119 // 2
    assert (0, 1) in got and (5, 9) in got and (17, 80) in got
    assert dispatch.rr_prefilter(reqs[:1], T) == []


def test_native_locality_order_equals_the_numpy_one():
    """drs_locality_order (the C++ helper other bindings would call) gives the numbering routing.RoadGraph computes with numpy."""
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import locality_order
    data = synth.config_graph("C2")
    a = locality_order(data.n, data.src, data.dst, data.vattrs["lng"], data.vattrs["lat"], native=True)
    b = locality_order(data.n, data.src, data.dst, data.vattrs["lng"], data.vattrs["lat"], native=False)
    assert np.array_equal(a, b) and sorted(a.tolist()) == list(range(data.n))
    rs = np.random.RandomState(0)
    x, y = rs.rand(3000).round(2), rs.rand(3000).round(2)          # irregular coordinates with many duplicates
    assert np.array_equal(locality_order(3000, [], [], x, y, native=True), locality_order(3000, [], [], x, y, native=False))


def _reference_rtv_level(keys_by_veh, k, rr, on):
    """The loops of lib/Optimization.py:265-288 (k = 2) and :291-347 (k >= 3), restated literally for one level."""
    tasks, saved = [], 0

    def has_rr(a, b):
        return ((a, b) if a < b else (b, a)) in rr
    for vi, keys in enumerate(keys_by_veh):
        if not keys or not on[vi]:
            continue
        if k == 2:
            singles = [t[0] for t in keys]
            for x in range(len(singles)):
                for y in range(x + 1, len(singles)):
                    if has_rr(singles[x], singles[y]):
                        tasks.append((vi, (singles[x], singles[y])))
                    else:
                        saved += 1
            continue
        keysets = [frozenset(t) for t in keys]
        tested = []
        for t1 in keys:
            for t2 in keys:
                trip = frozenset(t1) | frozenset(t2)
                if len(trip) == k and trip not in tested:
                    tested.append(trip)
                    if not all(sum(1 for ks in keysets if (trip - {r}) <= ks) == 1 for r in trip):
                        saved += 1
                        break
                    members = sorted(trip)
                    if all(has_rr(a, b) for i, a in enumerate(members) for b in members[i + 1:]):
                        tasks.append((vi, tuple(members)))
                    else:
                        saved += 1
    return tasks, saved


@pytest.mark.parametrize("k", [2, 3, 4])
def test_native_rtv_level_equals_the_references_loops(k):
    """drs_rtv_level (host-only entry point of the library) against the reference's per-vehicle set loops, including the
    order of the candidates and the early exit of the inner loop (lib/Optimization.py:325-327)."""
    import ctypes as C
    import itertools
    from drs_abm_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(100 + k)
    for trial in range(12):
        nr, nv = int(rs.randint(3, 70)), int(rs.randint(1, 9))
        dense = rs.rand() < 0.5
        rr = set()
        for a in range(nr):
            for b in range(a + 1, nr):
                if rs.rand() < (0.8 if dense else 0.3):
                    rr.add((a, b))
        keys_by_veh = []
        for v in range(nv):
            pool = sorted(rs.choice(nr, size=min(nr, int(rs.randint(0, 9))), replace=False).tolist())
            combos = [c for c in itertools.combinations(pool, k - 1)]
            if k > 2:
                combos = [c for c in combos if rs.rand() < 0.75]
                order = rs.permutation(len(combos))          # higher levels are NOT held in sorted order by the reference
                combos = [combos[i] for i in order]
            keys_by_veh.append(combos)
        on = rs.rand(nv) < 0.85
        want, want_saved = _reference_rtv_level(keys_by_veh, k, rr, on)
        key_off = np.zeros(nv + 1, dtype=np.int32)
        key_off[1:] = np.cumsum([len(x) for x in keys_by_veh])
        key_req = np.array([m for keys in keys_by_veh for t in keys for m in t], dtype=np.int32).reshape(-1, k - 1)
        words = (nr + 63) // 64
        bits = np.zeros((nr, words * 64), dtype=bool)
        for a, b in rr:
            bits[a, b] = bits[b, a] = True
        rr_bits = np.ascontiguousarray(np.packbits(bits, axis=1, bitorder="little")).view(np.uint64)
        on8 = on.astype(np.uint8)
        for cap in (0, max(1, len(want) // 2), len(want) + 3):
            tv = np.full(max(cap, 1), -7, dtype=np.int32)
            tr = np.full((max(cap, 1), k), -7, dtype=np.int32)
            n_tasks, n_saved = C.c_int64(-1), C.c_int64(-1)
            rc = lib.drs_rtv_level(nv, k, _lib.ptr_i32(key_off), _lib.ptr_i32(key_req), on8.ctypes.data_as(_lib._u8p), nr,
                                   rr_bits.ctypes.data_as(_lib._u64p), cap, _lib.ptr_i32(tv), _lib.ptr_i32(tr),
                                   C.byref(n_tasks), C.byref(n_saved))
            assert rc == 0
            assert n_tasks.value == len(want) and n_saved.value == want_saved
            got = [(int(tv[i]), tuple(int(x) for x in tr[i])) for i in range(min(cap, len(want)))]
            assert got == want[:min(cap, len(want))]
    # more members than a trip may hold
    n_tasks, n_saved = C.c_int64(0), C.c_int64(0)
    z = np.zeros(2, dtype=np.int32)
    assert lib.drs_rtv_level(1, 5, _lib.ptr_i32(z), _lib.ptr_i32(z), None, 0, None, 0, None, None,
                             C.byref(n_tasks), C.byref(n_saved)) == _lib.ERR_CAPACITY


def test_partition_order_gives_a_valid_separator_partition():
    """drs_partition_order (host code of the library, no GPU): a permutation; part interiors are contiguous id ranges; no edge
    joins two interiors; separators are a small share of a grid city; works without coordinates (breadth-first bisection),
    on disconnected maps, on maps smaller than one part and on edgeless ones."""
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import default_part_target, partition_order

    def check(n, src, dst, lng, lat, target):
        order, pp = partition_order(n, src, dst, lng, lat, target)
        assert sorted(order.tolist()) == list(range(n)) and pp[0] == 0 and np.all(np.diff(pp) > 0) and pp[-1] <= n
        new_of_old = np.empty(n, dtype=np.int64); new_of_old[order] = np.arange(n)
        part = np.searchsorted(pp, np.arange(n), side="right") - 1
        part[pp[-1]:] = -1
        ps, pt = part[new_of_old[np.asarray(src, dtype=np.int64)]], part[new_of_old[np.asarray(dst, dtype=np.int64)]]
        assert not np.any((ps >= 0) & (pt >= 0) & (ps != pt))
        assert np.diff(pp).max() <= target
        return len(pp) - 1, n - int(pp[-1])

    d = synth.config_graph("C2")
    lng, lat = d.vattrs["lng"], d.vattrs["lat"]
    parts, nsep = check(d.n, d.src, d.dst, lng, lat, 200)
    assert parts >= 20 and nsep < d.n // 6
    parts_b, nsep_b = check(d.n, d.src, d.dst, None, None, 200)          # no coordinates
    assert parts_b >= 20 and nsep_b < d.n // 4
    assert check(d.n, d.src, d.dst, lng, lat, 10 ** 6) == (1, 0)         # one part, no separator
    two = np.concatenate([d.src, d.src + d.n]), np.concatenate([d.dst, d.dst + d.n])   # two disjoint copies
    check(2 * d.n, two[0], two[1], None, None, 300)
    check(2 * d.n, two[0], two[1], list(lng) * 2, list(lat) * 2, 300)    # coincident coordinates
    assert check(5, [], [], None, None, 2)[1] == 0                       # no edges: no separators
    assert 192 <= default_part_target(50171) <= 2048


def test_attribute_writes_stay_private_to_the_graph():
    """Attribute columns are shared with the loaded map until the first write (copy on write): writes through one graph
    change neither the RoadGraphData nor another graph made from it (igraph semantics: every Graph owns its attributes)."""
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import RoadGraph
    data = synth.config_graph("C1")
    lng0, tt0 = data.vattrs["lng"][0], float(data.eattrs["ttmedian"][3])
    G1, G2 = RoadGraph(data), RoadGraph(data)
    assert G1._vattrs["lng"] is data.vattrs["lng"]                 # shared until written
    G1.vs[0]["lng"] = 99.5
    G1.es[3]["ttmedian"] = 123.0
    G1.vs[1]["colour"] = "red"                                      # a new attribute
    assert G1.vs[0]["lng"] == 99.5 and G1.es[3]["ttmedian"] == 123.0 and G1.vs[1]["colour"] == "red" and G1.vs[2]["colour"] is None
    assert data.vattrs["lng"][0] == lng0 and float(data.eattrs["ttmedian"][3]) == tt0 and "colour" not in data.vattrs
    assert G2.vs[0]["lng"] == lng0 and G2.es[3]["ttmedian"] == tt0
    assert G1._coords is None and G2._coords is not None           # the numeric coordinates of the load no longer describe G1
    G1.es["ttmedian"] = [7.0] * G1.m                               # whole-column assignment
    assert G1.es[0]["ttmedian"] == 7.0 and float(data.eattrs["ttmedian"][0]) != 7.0 or tt0 == 7.0
