"""Drop-in assignment optimizers: the reference's ``AssignmentMethod`` family over
the B200 per-tick evaluation kernels.

Mirrors /root/reference lib/Optimization.py:13-408 (``AssignmentMethod``,
``VehReqMatchingAssignment``, ``AlonsoMora``): same class names, constructor
parameters, ``optimizeAssignment(vehs, reqs, G, T=0)`` signature and return
convention ``((veh_routes, rejected_rids), metadata)``.

What runs where
  * RV stage (every request x vehicle pair: reachability pre-filter, dummy-vehicle
    check, optimal ordering of the request inserted into the vehicle's plan) and
    every RR / RTV trip costing: CUDA (drs_tick_rv_eval / drs_tick_trip_eval),
    exact enumeration with the reference's fp64 arithmetic.
  * RTV clique growth (set logic, lib/Optimization.py:265-347) and the final
    set-packing ILP (lib/Optimization.py:359-408; Gurobi in the reference, a
    third-party solver): host.  The ILP goes to scipy's HiGHS, or to the
    rectangular assignment solver when every trip has one request.

``routing_method`` "tabu" (the reference default, lib/Optimization.py:96-100) is a
non-reproducible heuristic (unseeded random(), wall-clock limits:
lib/optimUtils.py:312,339-343); here both methods use the exact enumeration, so a
"tabu" run returns a cost <= the reference's.  "enumerate" additionally applies
findOptimalRouting's pairwise pre-filter (lib/optimUtils.py:127-156), which the
tabu front end has commented out (lib/optimUtils.py:172-200).

Tie rules that the reference leaves to set-iteration order / Gurobi are defined
in DESIGN.md ("canonical ties").
"""
import ctypes as C
import math
import time

import numpy as np

from . import _lib
from .routing import RoadGraph

def _reference_constant(name, default):
    """The reference star-imports lib/Constants.py into every module (lib/Optimization.py:9, lib/Agents.py:15); when this
    package runs inside the reference's tree that module wins, so an edited Constants.py keeps steering the drop-in."""
    try:
        from lib import Constants as _RC          # the reference's own module, if it is importable here
        return getattr(_RC, name, default)
    except Exception:                             # noqa: BLE001 (stand-alone use: the reference's published defaults)
        return default


CST_SPEED = _reference_constant("CST_SPEED", 9)   # lib/Constants.py:148, used by the RR pre-filter (lib/Optimization.py:148-151)
COST_IGNORED = 1e6         # lib/Optimization.py:373
INF_ROUTE_COST = 1e6       # lib/optimUtils.py:156,437
MAX_STOPS = 16
MAX_TRIP_REQS = 4
EVAL_PAIRWISE_CHECKS = 1

_i8p, _u8p, _i64p = _lib._i8p, _lib._u8p, _lib._i64p
VehicleBatch, RequestBatch = _lib.VehicleBatch, _lib.RequestBatch


def great_circle_distance(olat, olng, dlat, dlng):
    """lib/optimUtils.py:651-653."""
    return (6371000 * 2 * math.pi / 360 * np.sqrt(
        (math.cos((olat + dlat) * math.pi / 360) * (olng - dlng)) ** 2 + (olat - dlat) ** 2))


def rr_prefilter_pairs(new_reqs, T):
    """Pairs (a < b) of new requests that pass the crow-flies pre-filter of lib/Optimization.py:148-151:
    ``T + greatCircleDistance(o1, o2) / CST_SPEED <= Clp`` of BOTH requests, as an int32 array [m, 2] in lexicographic order.

    The reference evaluates the formula pair by pair in Python floats; here all R^2 pairs are screened at once with numpy
    and only the pairs within a relative 1e-9 of the threshold (where a last-bit difference between numpy's and libm's
    cosine could matter) are decided by the reference's own scalar expression -- same result, no Python double loop
    (2*10^6 calls at BASELINE config C4)."""
    nr = len(new_reqs)
    if nr < 2:
        return np.zeros((0, 2), dtype=np.int32)
    olat = np.array([float(r.olat) for r in new_reqs], dtype=np.float64)
    olng = np.array([float(r.olng) for r in new_reqs], dtype=np.float64)
    clp = np.array([float(r.Clp) for r in new_reqs], dtype=np.float64)
    out_a, out_b = [], []
    step = max(1, (1 << 22) // nr)                          # ~4M pairs per block
    for a0 in range(0, nr, step):
        a1 = min(nr, a0 + step)
        la, lo = olat[a0:a1, None], olng[a0:a1, None]
        dist = 6371000 * 2 * math.pi / 360 * np.sqrt((np.cos((la + olat[None, :]) * math.pi / 360) * (lo - olng[None, :])) ** 2
                                                     + (la - olat[None, :]) ** 2)
        tmin = T + dist / CST_SPEED
        lim = np.minimum(clp[a0:a1, None], clp[None, :])
        margin = tmin - lim                                  # > 0: filtered out
        upper = np.triu(np.ones((a1 - a0, nr), dtype=bool), k=a0 + 1)
        sure = (margin <= -1e-9 * np.maximum(1.0, np.abs(lim))) & upper
        close = (np.abs(margin) < 1e-9 * np.maximum(1.0, np.abs(lim))) & upper
        aa, bb = np.nonzero(sure)
        out_a.append(aa + a0)
        out_b.append(bb)
        ca, cb = [], []
        for a, b in zip(*np.nonzero(close)):
            r1, r2 = new_reqs[a + a0], new_reqs[b]
            t = T + great_circle_distance(r1.olat, r1.olng, r2.olat, r2.olng) / CST_SPEED
            if not (t > r1.Clp or t > r2.Clp):
                ca.append(int(a + a0))
                cb.append(int(b))
        out_a.append(np.array(ca, dtype=np.int64))
        out_b.append(np.array(cb, dtype=np.int64))
    a = np.concatenate(out_a)
    b = np.concatenate(out_b)
    o = np.lexsort((b, a))
    return np.stack([a[o], b[o]], axis=1).astype(np.int32)


def rr_prefilter(new_reqs, T):
    """rr_prefilter_pairs as a list of (a, b) tuples."""
    return [(int(a), int(b)) for a, b in rr_prefilter_pairs(new_reqs, T)]


class Tick(object):
    """Host view of one assignment tick: marshals Veh / Req objects into the SoA
    batches of include/drs_b200.h and runs the evaluation kernels."""

    def __init__(self, G, vehs, reqs, T, pairwise_checks=True):
        if not isinstance(G, RoadGraph):
            raise TypeError("G must be the RoadGraph proxy (router.G) of drs_abm_b200.routing.RoutingEngine")
        self.G = G
        self.T = T
        self.flags = EVAL_PAIRWISE_CHECKS if pairwise_checks else 0
        self.lib = _lib.load()
        self.vehs = list(vehs)
        # ---- request vertices: lib/Optimization.py:106-118
        self.new_reqs = [r for r in reqs if r.assigned is not None and not r.assigned]
        node = G.node_of_location
        nr = len(self.new_reqs)
        self.r_rid = np.array([r.id for r in self.new_reqs], dtype=np.int32)
        self.r_o = G.dev_ids(np.array([node(r.olng, r.olat) for r in self.new_reqs], dtype=np.int32))
        self.r_d = G.dev_ids(np.array([node(r.dlng, r.dlat) for r in self.new_reqs], dtype=np.int32))
        self.r_lbp = np.array([r.Cep - T for r in self.new_reqs], dtype=np.float64)       # Cep - T (:117)
        self.r_ubp = np.array([r.Clp for r in self.new_reqs], dtype=np.float64)
        dw = [r.get_dropoff_time_window() for r in self.new_reqs]                          # (:111)
        self.r_lbd = np.array([w[0] for w in dw], dtype=np.float64)
        self.r_ubd = np.array([w[1] for w in dw], dtype=np.float64)
        assert len(self.r_rid) == nr
        # ---- vehicle vertices: lib/Optimization.py:120-134 ; stop list in createLocations order
        #      (lib/optimUtils.py:94-108): on-board drop-offs, then (pickup, drop-off) per waiting request
        v_node, v_delay, v_cap, v_pax, off = [], [], [], [], [0]
        s_node, s_rid, s_pod, s_prev, s_lb, s_ub = [], [], [], [], [], []
        self.veh_stops = []       # per vehicle: [(rid, pod, lng, lat)] in batch order
        for veh in self.vehs:
            pax_rids, wait_rids = veh.get_passenger_requests()
            (vlng, vlat), t = veh.get_next_node()
            v_node.append(node(vlng, vlat)); v_delay.append(float(t)); v_cap.append(int(veh.K))
            pax_list, wait_list = list(pax_rids), list(wait_rids)
            v_pax.append(len(pax_list))
            stops = []
            for rid in pax_list:
                rq = reqs[rid]
                Ced, Cld = rq.get_dropoff_time_window()
                s_node.append(node(rq.dlng, rq.dlat)); s_rid.append(rq.id); s_pod.append(-1); s_prev.append(-1)
                s_lb.append(Ced); s_ub.append(Cld)
                stops.append((rq.id, -1, rq.dlng, rq.dlat))
            for rid in wait_list:
                rq = reqs[rid]
                Ced, Cld = rq.get_dropoff_time_window()
                s_node.append(node(rq.olng, rq.olat)); s_rid.append(rq.id); s_pod.append(1); s_prev.append(-1)
                s_lb.append(rq.Cep); s_ub.append(rq.Clp)
                stops.append((rq.id, 1, rq.olng, rq.olat))
                s_node.append(node(rq.dlng, rq.dlat)); s_rid.append(rq.id); s_pod.append(-1)
                s_prev.append(len(stops) - 1)
                s_lb.append(Ced); s_ub.append(Cld)
                stops.append((rq.id, -1, rq.dlng, rq.dlat))
            if len(stops) > MAX_STOPS - 2:
                raise ValueError("vehicle %r has %d planned stops; the kernel handles %d" % (veh.id, len(stops), MAX_STOPS - 2))
            self.veh_stops.append(stops)
            off.append(len(s_node))
        self._v = dict(node=G.dev_ids(np.array(v_node, dtype=np.int32)), delay=np.array(v_delay, dtype=np.float64),
                       cap=np.array(v_cap, dtype=np.int32), pax=np.array(v_pax, dtype=np.int32),
                       off=np.array(off, dtype=np.int32), s_node=G.dev_ids(np.array(s_node, dtype=np.int32)),
                       s_rid=np.array(s_rid, dtype=np.int32), s_pod=np.array(s_pod, dtype=np.int8),
                       s_prev=np.array(s_prev, dtype=np.int8), s_lb=np.array(s_lb, dtype=np.float64),
                       s_ub=np.array(s_ub, dtype=np.float64))
        self.handle = C.c_void_p()
        _lib.check(self.lib.drs_tick_create(G.handle, C.byref(self.handle)), self.lib)
        self._upload()

    def _upload(self):
        v = self._v

        def p8(a):
            return a.ctypes.data_as(_i8p)
        vb = VehicleBatch(len(self.vehs), _lib.ptr_i32(v["node"]), _lib.ptr_f64(v["delay"]), _lib.ptr_i32(v["cap"]),
                          _lib.ptr_i32(v["pax"]), _lib.ptr_i32(v["off"]), _lib.ptr_i32(v["s_node"]),
                          _lib.ptr_i32(v["s_rid"]), p8(v["s_pod"]), p8(v["s_prev"]), _lib.ptr_f64(v["s_lb"]),
                          _lib.ptr_f64(v["s_ub"]))
        _lib.check(self.lib.drs_tick_set_vehicles(self.handle, C.byref(vb)), self.lib)
        rb = RequestBatch(len(self.new_reqs), _lib.ptr_i32(self.r_rid), _lib.ptr_i32(self.r_o), _lib.ptr_i32(self.r_d),
                          _lib.ptr_f64(self.r_lbp), _lib.ptr_f64(self.r_ubp), _lib.ptr_f64(self.r_lbd),
                          _lib.ptr_f64(self.r_ubd))
        _lib.check(self.lib.drs_tick_set_requests(self.handle, C.byref(rb)), self.lib)

    def close(self):
        if self.handle:
            self.lib.drs_tick_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ kernels
    def rv_eval_arrays(self):
        """RV stage.  Returns (veh_idx, req_idx, cost, order [n, MAX_STOPS], n_stops, counters): the feasible
        (vehicle, request) edges as arrays, sorted by (request, vehicle)."""
        n_edges = C.c_int32(0)
        counters = np.zeros(4, dtype=np.int64)
        _lib.check(self.lib.drs_tick_rv_eval(self.handle, float(self.T), self.flags, C.byref(n_edges),
                                             counters.ctypes.data_as(_i64p)), self.lib)
        ne = n_edges.value
        vi, ri = np.empty(ne, dtype=np.int32), np.empty(ne, dtype=np.int32)
        cost = np.empty(ne, dtype=np.float64)
        order = np.empty((ne, MAX_STOPS), dtype=np.uint8)
        ns = np.empty(ne, dtype=np.int32)
        _lib.check(self.lib.drs_tick_rv_fetch(self.handle, ne, _lib.ptr_i32(vi), _lib.ptr_i32(ri), _lib.ptr_f64(cost),
                                              order.ctypes.data_as(_u8p), _lib.ptr_i32(ns)), self.lib)
        return vi, ri, cost, order, ns, {"feasible": int(counters[0]), "skipped": int(counters[1]),
                                         "saved": int(counters[2]), "infeasible": int(counters[3])}

    def rv_eval(self):
        """RV stage.  Returns (edges, counters): edges = list of (veh_index, req_index, cost, order)."""
        vi, ri, cost, order, ns, counters = self.rv_eval_arrays()
        edges = [(int(vi[e]), int(ri[e]), float(cost[e]), [int(x) for x in order[e, :ns[e]]]) for e in range(len(vi))]
        return edges, counters

    def trip_eval_arrays(self, task_veh, task_req):
        """task_veh [n] (vehicle index or -1), task_req [n, k] (request indices; k <= MAX_TRIP_REQS, every task of one
        size).  Returns (cost [n], order [n, MAX_STOPS] uint8, n_stops [n]); cost == INF_ROUTE_COST: infeasible."""
        tv = _lib.as_i32(task_veh)
        nt = len(tv)
        members = np.asarray(task_req, dtype=np.int32)
        members = members.reshape(nt, -1) if nt else members.reshape(0, members.shape[1] if members.ndim == 2 else 1)
        k = members.shape[1]
        if k > MAX_TRIP_REQS:
            raise ValueError("a trip holds at most %d new requests" % MAX_TRIP_REQS)
        cost = np.empty(nt, dtype=np.float64)
        order = np.empty((nt, MAX_STOPS), dtype=np.uint8)
        ns = np.empty(nt, dtype=np.int32)
        if nt == 0:
            return cost, order, ns
        tr = np.zeros((nt, MAX_TRIP_REQS), dtype=np.int32)
        tr[:, :k] = members
        tn = np.full(nt, k, dtype=np.int32)
        _lib.check(self.lib.drs_tick_trip_eval(self.handle, float(self.T), self.flags, nt, _lib.ptr_i32(tv),
                                               _lib.ptr_i32(tn), _lib.ptr_i32(tr), _lib.ptr_f64(cost),
                                               order.ctypes.data_as(_u8p), _lib.ptr_i32(ns)), self.lib)
        return cost, order, ns

    def trip_eval(self, tasks):
        """tasks: list of (veh_index or -1, [request indices]).  Returns list of (cost, order or None)."""
        nt = len(tasks)
        if nt == 0:
            return []
        tv = np.array([t[0] for t in tasks], dtype=np.int32)
        tn = np.array([len(t[1]) for t in tasks], dtype=np.int32)
        tr = np.zeros((nt, MAX_TRIP_REQS), dtype=np.int32)
        for k, t in enumerate(tasks):
            if len(t[1]) > MAX_TRIP_REQS:
                raise ValueError("a trip holds at most %d new requests" % MAX_TRIP_REQS)
            tr[k, :len(t[1])] = t[1]
        cost = np.empty(nt, dtype=np.float64)
        order = np.empty((nt, MAX_STOPS), dtype=np.uint8)
        ns = np.empty(nt, dtype=np.int32)
        _lib.check(self.lib.drs_tick_trip_eval(self.handle, float(self.T), self.flags, nt, _lib.ptr_i32(tv),
                                               _lib.ptr_i32(tn), _lib.ptr_i32(tr), _lib.ptr_f64(cost),
                                               order.ctypes.data_as(_u8p), _lib.ptr_i32(ns)), self.lib)
        out = []
        for k in range(nt):
            if cost[k] < INF_ROUTE_COST:
                out.append((float(cost[k]), [int(x) for x in order[k, :ns[k]]]))
            else:
                out.append((INF_ROUTE_COST, None))
        return out

    # ------------------------------------------------------------------ routes
    def route_of(self, veh_index, req_indices, order):
        """Kernel stop numbering -> the reference's route format [(rid, pod, lng, lat), ...]
        (lib/optimUtils.py:358,449)."""
        stops = []
        for ri in req_indices:
            rq = self.new_reqs[ri]
            stops.append((rq.id, 1, rq.olng, rq.olat))
            stops.append((rq.id, -1, rq.dlng, rq.dlat))
        if veh_index >= 0:
            stops.extend(self.veh_stops[veh_index])
        return [stops[i] for i in order]


class _Sized(object):
    """len()-able stand-in for an igraph vertex / edge sequence that nobody iterates."""

    def __init__(self, n):
        self.n = int(n)

    def __len__(self):
        return self.n


class _Trips(object):
    """Feasible trips of one size as arrays: veh [n] (vehicle index, -1 for request-request edges), members [n, k]
    (request indices, ascending), cost [n], order [n, MAX_STOPS] / ns [n] (the optimal stop sequence in the kernel's
    numbering, see Tick.route_of)."""

    def __init__(self, veh, members, cost, order, ns):
        self.veh, self.members, self.cost, self.order, self.ns = veh, members, cost, order, ns
        self.k = members.shape[1]

    def __len__(self):
        return len(self.veh)

    def take(self, sel):
        return _Trips(self.veh[sel], self.members[sel], self.cost[sel], self.order[sel], self.ns[sel])

    def route(self, tick, e):
        return tick.route_of(int(self.veh[e]), [int(x) for x in self.members[e]], [int(x) for x in self.order[e, :self.ns[e]]])


class _RVGraph(object):
    """What generatePairwiseGraph hands to generateRTVGraph (an igraph Graph in the reference): .vs / .es for len(),
    the edges as arrays (rr_edges, rv_edges) and, built on first use, as the dicts .rr {(a, b): (cost, route)} and
    .rv {(veh index, request index): (cost, route)}."""

    def __init__(self, tick, rr_edges, rv_edges, counters):
        self.tick, self.rr_edges, self.rv_edges, self.counters = tick, rr_edges, rv_edges, counters
        self.vs = _Sized(len(tick.new_reqs) + len(tick.vehs))
        self.es = _Sized(len(rr_edges) + len(rv_edges))
        nr = len(tick.new_reqs)
        self.rr_ok = np.zeros((nr, nr), dtype=bool)
        if len(rr_edges):
            a, b = rr_edges.members[:, 0], rr_edges.members[:, 1]
            self.rr_ok[a, b] = True
            self.rr_ok[b, a] = True
        self._rr = self._rv = None

    @property
    def rr(self):
        if self._rr is None:
            E = self.rr_edges
            self._rr = {(int(E.members[e, 0]), int(E.members[e, 1])): (float(E.cost[e]), E.route(self.tick, e)) for e in range(len(E))}
        return self._rr

    @property
    def rv(self):
        if self._rv is None:
            E = self.rv_edges
            self._rv = {(int(E.veh[e]), int(E.members[e, 0])): (float(E.cost[e]), E.route(self.tick, e)) for e in range(len(E))}
        return self._rv


class _RTVGraph(object):
    """What generateRTVGraph hands to assignFromRTVGraph: the feasible (vehicle, trip) edges level by level (size 1, 2,
    ...), each level ordered by vehicle and, within a vehicle, in the order the reference creates the trips."""

    def __init__(self, rv_graph, levels):
        self.tick, self.levels, self.objective = rv_graph.tick, levels, None
        nr = len(self.tick.new_reqs)
        n_trips = 0
        for L in levels:                                     # distinct request sets = the reference's trip vertices
            if len(L) == 0:
                continue
            if nr < 65536:
                key = np.zeros(len(L), dtype=np.int64)
                for i in range(L.k):
                    key |= L.members[:, i].astype(np.int64) << (16 * i)
                n_trips += len(np.unique(key))
            else:
                n_trips += len(np.unique(L.members, axis=0))
        self.vs = _Sized(len(rv_graph.vs) + n_trips)
        self.es = _Sized(sum(len(L) for L in levels))
        self._veh_trips = None

    @property
    def veh_trips(self):
        """{(veh index, trip tuple): (cost, route)} (built on first use; the assignment works on the arrays)."""
        if self._veh_trips is None:
            self._veh_trips = {}
            for L in self.levels:
                for e in range(len(L)):
                    self._veh_trips[(int(L.veh[e]), tuple(int(x) for x in L.members[e]))] = (float(L.cost[e]), L.route(self.tick, e))
        return self._veh_trips


class AssignmentMethod():
    """lib/Optimization.py:13-27."""

    def __init__(self, params):
        self.parameters = params

    def optimizeAssignment(self, vehs, reqs, G, T=0):
        pass


class VehReqMatchingAssignment(AssignmentMethod):
    """lib/Optimization.py:29-85: RV graph -> RTV graph -> assignment."""

    verbose = True

    def _print(self, *a):
        if self.verbose:
            print(*a)

    def optimizeAssignment(self, vehs, reqs, G, T=0):
        self._print("  Generating R-V graph for T={}".format(T))
        t = time.time()
        rvGraph = self.generatePairwiseGraph(vehs, reqs, G, T)
        rvTime = time.time() - t
        self._print("  R-V Graph successfully generated  in {:.2f} seconds! |V|={}, |E|={}".format(
            rvTime, len(rvGraph.vs), len(rvGraph.es)))
        self._print("  Generating RTV graph for T={}".format(T))
        t = time.time()
        rtvGraph = self.generateRTVGraph(rvGraph, G, T)
        rtvTime = time.time() - t
        self._print("  RTV Graph successfully generated in {:.2f} seconds! |V|={}, |E|={}".format(
            rtvTime, len(rtvGraph.vs), len(rtvGraph.es)))
        try:
            if len(rtvGraph.vs) > 0 and len(rtvGraph.es) > 0:
                t = time.time()
                assignment = self.assignFromRTVGraph(rtvGraph)
                assTime = time.time() - t
                self._print("  Assignment completed in {:.2f} seconds!".format(assTime))
                metadata = {"benders_iters": 0, "total_feas_cuts": 0, "total_opt_cuts": 0, "solve_time": assTime,
                            "rv_time": rvTime, "rtv_time": rtvTime, "objective": rtvGraph.objective}
                return assignment, metadata
            return (list(), list()), {"solve_time": 0}      # lib/Optimization.py:75-76
        finally:
            rvGraph.tick.close()

    def generatePairwiseGraph(self, vehs, reqs, G, T=0):
        pass

    def generateRTVGraph(self, rvGraph, G, T=0):
        pass

    def assignFromRTVGraph(self, rtvGraph):
        pass


class AlonsoMora(VehReqMatchingAssignment):
    """lib/Optimization.py:88-408."""

    def __init__(self, params):
        self.parameters = params
        if "routing_method" not in params.keys():
            self.parameters["routing_method"] = "tabu"
        if self.parameters["routing_method"] == "tabu" and "tabu_max_time" not in params.keys():
            self.parameters["tabu_max_time"] = 3
        if self.parameters["routing_method"] not in ("tabu", "enumerate"):
            raise ValueError("routing_method must be 'tabu' or 'enumerate'")
        self.title = "Alonso-Mora"
        self.verbose = bool(params.get("verbose", True))
        self.max_trip_size = params.get("max_trip_size", None)   # extension: cap RTV trip size (None = veh capacity)

    # ------------------------------------------------------------------ RV graph
    def generatePairwiseGraph(self, vehs, reqs, G, T=0):
        """lib/Optimization.py:104-225.  Returns an object with .vs / .es (for len()) carrying the
        request-request and request-vehicle edges (arrays; dict views .rr / .rv on demand)."""
        tick = Tick(G, vehs, reqs, T, pairwise_checks=(self.parameters["routing_method"] == "enumerate"))
        # ---- R x R (lib/Optimization.py:142-165): crow-flies pre-filter on the host, costing on the GPU
        pairs = rr_prefilter_pairs(tick.new_reqs, T)
        cost, order, ns = tick.trip_eval_arrays(np.full(len(pairs), -1, dtype=np.int32), pairs)
        ok = cost < INF_ROUTE_COST
        rr_edges = _Trips(np.full(int(ok.sum()), -1, dtype=np.int32), pairs[ok], cost[ok], order[ok], ns[ok])
        self._print("    Feasible pairings found amongst requests: {}".format(int(ok.sum())))
        self._print("    Infeasible pairings evaluated: {}".format(int(len(ok) - ok.sum())))
        # ---- R x V (lib/Optimization.py:176-218)
        vi, ri, cost, order, ns, counters = tick.rv_eval_arrays()
        rv_edges = _Trips(vi, ri.reshape(-1, 1), cost, order, ns)
        self._print("    Feasible pairings found between requests and vehicles: {}".format(counters["feasible"]))
        self._print("    Infeasible pairings evaluated: {}".format(counters["infeasible"]))
        self._print("    Infeasible pairings skipped via dummy evaluation: {}".format(counters["saved"]))
        self._print("    Pairings skipped because vehicle was too far away: {}".format(counters["skipped"]))
        return _RVGraph(tick, rr_edges, rv_edges, counters)

    # ------------------------------------------------------------------ RTV graph
    def _next_level(self, tick, prev, k, veh_on, rr_bits):
        """Candidate trips of size k for every vehicle (drs_rtv_level: lib/Optimization.py:265-347 on the host, native)."""
        nv, nr = len(tick.vehs), len(tick.new_reqs)
        key_off = np.zeros(nv + 1, dtype=np.int32)
        np.cumsum(np.bincount(prev.veh, minlength=nv), out=key_off[1:])
        key_req = np.ascontiguousarray(prev.members, dtype=np.int32)
        on = None if veh_on is None else np.ascontiguousarray(veh_on, dtype=np.uint8)
        cap = max(1024, 4 * len(prev))
        while True:
            tv = np.empty(cap, dtype=np.int32)
            tr = np.empty((cap, k), dtype=np.int32)
            n_tasks, n_saved = C.c_int64(0), C.c_int64(0)
            _lib.check(tick.lib.drs_rtv_level(nv, k, _lib.ptr_i32(key_off), _lib.ptr_i32(key_req),
                                              None if on is None else on.ctypes.data_as(_u8p), nr,
                                              rr_bits.ctypes.data_as(_lib._u64p), cap, _lib.ptr_i32(tv), _lib.ptr_i32(tr),
                                              C.byref(n_tasks), C.byref(n_saved)), tick.lib)
            if n_tasks.value <= cap:
                return tv[:n_tasks.value], tr[:n_tasks.value], n_saved.value
            cap = n_tasks.value

    def generateRTVGraph(self, rvGraph, G, T=0):
        """lib/Optimization.py:227-357.  A level = the feasible trips of one size for the whole fleet: its candidates
        come from one native enumeration pass (drs_rtv_level) and are costed in one kernel launch."""
        tick = rvGraph.tick
        nr = len(tick.new_reqs)
        rv = rvGraph.rv_edges
        # veh.neighbors(): ascending vertex id (:243-262)
        levels = [rv.take(np.lexsort((rv.members[:, 0], rv.veh)))]
        cap = tick._v["cap"]
        max_size = cap if self.max_trip_size is None else np.minimum(cap, int(self.max_trip_size))
        words = (nr + 63) // 64
        padded = np.zeros((nr, words * 64), dtype=bool)
        padded[:, :nr] = rvGraph.rr_ok
        rr_bits = np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little")).view(np.uint64)
        feas = inf = saved = 0
        k = 2
        while len(levels[-1]):
            prev = levels[-1]
            if k == 2:                                        # trips of size two (:265-288)
                if self.max_trip_size is not None and int(self.max_trip_size) < 2:
                    break
                veh_on = None
            else:                                             # trips of size three and more (:291-347)
                veh_on = max_size >= k
                if not veh_on[prev.veh].any():
                    break
            tv, tr, sv = self._next_level(tick, prev, k, veh_on, rr_bits)
            saved += sv
            cost, order, ns = tick.trip_eval_arrays(tv, tr)
            ok = cost < INF_ROUTE_COST
            feas += int(ok.sum())
            inf += int(len(ok) - ok.sum())
            levels.append(_Trips(tv[ok], tr[ok], cost[ok], order[ok], ns[ok]))
            k += 1
        self._print("    Feasible trip assignments evaluated: {}".format(feas))
        self._print("    Infeasible trip assignments evaluated: {}".format(inf))
        self._print("    Infeasible trip evaluations skipped via incomplete subgraphs / previous trip length checks: {}".format(saved))
        return _RTVGraph(rvGraph, [L for L in levels if len(L)])

    # ------------------------------------------------------------------ assignment
    def assignFromRTVGraph(self, rtvGraph):
        """lib/Optimization.py:359-408: min sum cost*e + 1e6*sum x; one trip per vehicle; every request in
        exactly one chosen trip or ignored.  Returns (veh_routes {vid: route}, rejected {rid})."""
        tick = rtvGraph.tick
        nreq = len(tick.new_reqs)
        L = rtvGraph.levels
        veh = np.concatenate([l.veh for l in L]) if L else np.zeros(0, dtype=np.int32)
        size = np.concatenate([np.full(len(l), l.k, dtype=np.int32) for l in L]) if L else np.zeros(0, dtype=np.int32)
        members = np.full((len(veh), MAX_TRIP_REQS), -1, dtype=np.int32)
        level_of = np.concatenate([np.full(len(l), i, dtype=np.int32) for i, l in enumerate(L)]) if L else np.zeros(0, dtype=np.int32)
        pos_in = np.concatenate([np.arange(len(l), dtype=np.int64) for l in L]) if L else np.zeros(0, dtype=np.int64)
        at = 0
        for l in L:
            members[at:at + len(l), :l.k] = l.members
            at += len(l)
        costs = np.concatenate([l.cost for l in L]) if L else np.zeros(0)
        # the column order of the reference's model: by vehicle, then trip size, then members
        o = np.lexsort(tuple(members[:, i] for i in range(MAX_TRIP_REQS - 1, -1, -1)) + (size, veh))
        chosen, ignored, objective = _solve_packing_arrays(len(tick.vehs), nreq, veh[o], size[o], members[o], costs[o])
        if chosen is None:
            return None, None                                 # lib/Optimization.py:405-406
        rtvGraph.objective = objective
        veh_routes = dict()
        for idx in chosen:
            e = o[idx]
            lv = L[level_of[e]]
            veh_routes[tick.vehs[int(veh[e])].id] = lv.route(tick, int(pos_in[e]))
        rejected = set(int(tick.new_reqs[ri].id) for ri in ignored)
        return veh_routes, rejected


def solve_set_packing(nveh, nreq, trips, costs):
    """The ILP of lib/Optimization.py:359-408.  trips: list of (veh_index, tuple of request
    indices).  Returns (chosen trip indices, ignored request indices, objective)."""
    nt = len(trips)
    veh = np.array([t[0] for t in trips], dtype=np.int32)
    size = np.array([len(t[1]) for t in trips], dtype=np.int32)
    members = np.full((nt, max([1] + [len(t[1]) for t in trips])), -1, dtype=np.int32)
    for k, t in enumerate(trips):
        members[k, :len(t[1])] = t[1]
    return _solve_packing_arrays(nveh, nreq, veh, size, members, np.asarray(costs, dtype=np.float64))


def _solve_packing_arrays(nveh, nreq, veh, size, members, costs):
    """solve_set_packing on arrays: veh [nt], size [nt], members [nt, kmax] (-1 padded), costs [nt]."""
    nt = len(veh)
    if nt == 0:
        return [], list(range(nreq)), COST_IGNORED * nreq
    if (size == 1).all():
        # rectangular assignment: serve request r by vehicle v at cost c, or ignore it at 1e6
        from scipy.optimize import linear_sum_assignment
        big = 4.0 * COST_IGNORED
        M = np.full((nreq, nveh + nreq), big)
        tid = -np.ones((nreq, nveh), dtype=np.int64)
        r = members[:, 0]
        first = np.lexsort((np.arange(nt), costs, veh, r))   # of several trips for one (r, v): the cheapest, then the first
        keep = np.ones(nt, dtype=bool)
        keep[1:] = (r[first][1:] != r[first][:-1]) | (veh[first][1:] != veh[first][:-1])
        sel = first[keep]
        sel = sel[costs[sel] < big]
        M[r[sel], veh[sel]] = costs[sel]
        tid[r[sel], veh[sel]] = sel
        M[np.arange(nreq), nveh + np.arange(nreq)] = COST_IGNORED
        rows, cols = linear_sum_assignment(M)
        chosen, ignored = [], []
        for rr_, c in zip(rows, cols):
            if c < nveh and tid[rr_, c] >= 0:
                chosen.append(int(tid[rr_, c]))
            else:
                ignored.append(int(rr_))
        obj = float(sum(costs[k] for k in chosen) + COST_IGNORED * len(ignored))
        return sorted(chosen), sorted(ignored), obj
    from scipy.optimize import milp, LinearConstraint, Bounds
    from scipy.sparse import csr_matrix
    nvar = nt + nreq
    c = np.concatenate([costs, np.full(nreq, COST_IGNORED)])
    lo = np.concatenate([np.full(nveh, -np.inf), np.ones(nreq)])
    hi = np.ones(nveh + nreq)
    # rows 0..nveh-1: one trip per vehicle; rows nveh..: each request in exactly one chosen trip or ignored
    tk, mk = np.nonzero(members >= 0)
    rows = np.concatenate([veh.astype(np.int64), nveh + members[tk, mk].astype(np.int64), nveh + np.arange(nreq, dtype=np.int64)])
    cols = np.concatenate([np.arange(nt, dtype=np.int64), tk.astype(np.int64), nt + np.arange(nreq, dtype=np.int64)])
    A = csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(nveh + nreq, nvar))
    res = milp(c, constraints=LinearConstraint(A, lo, hi), integrality=np.ones(nvar), bounds=Bounds(0, 1))
    if res.status != 0:
        return None, None, None
    x = np.round(res.x).astype(np.int64)
    chosen = [int(k) for k in np.nonzero(x[:nt] > 0)[0]]
    ignored = [int(r) for r in np.nonzero(x[nt:] > 0)[0]]
    return chosen, ignored, float(res.fun)


# ---------------------------------------------------------------------------------------------
# Legacy insertion heuristic (lib/Agents.py:1429-1572)
# ---------------------------------------------------------------------------------------------
MAX_DETOUR = _reference_constant("MAX_DETOUR", 1.5)             # lib/Constants.py:144
COEF_INVEH = _reference_constant("COEF_INVEH", 1.0)             # lib/Constants.py:163-165
COEF_WAIT = _reference_constant("COEF_WAIT", 1.5)
COEF_EXTRA_WAIT = _reference_constant("COEF_EXTRA_WAIT", 3.0)


class InsertionTick(object):
    """Array-level front end of drs_tick_set_plans / drs_tick_insertion_eval / drs_tick_insertion_queue.

    vehs: list of dicts {node, t0, n, T, K, c, route: [(request row, pod, node)]} and, for commits, optionally
    ``pos`` (vertex the vehicle stands on, None between vertices) and ``keep`` = (node, t, et) of the first step of
    its current route (None: no route) -- what Veh.build_route keeps in front of a new route (lib/Agents.py:319-336);
    table: dict of request-table columns o, d, Cep, Clp, Cld, Ts, Tr, pwt (numpy arrays)."""

    MAX_PLAN = 16          # DRS_INS_MAX_PLAN

    def __init__(self, G, vehs, table, params=None):
        if not isinstance(G, RoadGraph):
            raise TypeError("G must be the RoadGraph proxy (router.G)")
        self.lib = _lib.load()
        self.G = G
        self.handle = C.c_void_p()
        _lib.check(self.lib.drs_tick_create(G.handle, C.byref(self.handle)), self.lib)
        self.set_plans(vehs, table, params)

    def set_plans(self, vehs, table, params=None):
        """(Re)loads the fleet and the request table; device buffers are reused from tick to tick."""
        G = self.G
        nv = len(vehs)
        self.n_veh = nv
        off = np.zeros(nv + 1, dtype=np.int32)
        for v, veh in enumerate(vehs):
            off[v + 1] = off[v] + len(veh["route"])

        def nodes_or_minus_one(vals):
            a = np.array([-1 if x is None else int(x) for x in vals], dtype=np.int32)
            ok = a >= 0
            if ok.any():
                a[ok] = G.dev_ids(a[ok])
            return a
        have_keep = any("keep" in v for v in vehs)
        have_pos = any("pos" in v for v in vehs)
        self._a = dict(
            node=G.dev_ids(np.array([v["node"] for v in vehs], dtype=np.int32)),
            delay=np.array([v["t0"] for v in vehs], dtype=np.float64),
            onboard=np.array([v["n"] for v in vehs], dtype=np.int32),
            clock=np.array([v["T"] for v in vehs], dtype=np.float64),
            cap=np.array([v["K"] for v in vehs], dtype=np.int32),
            cost=np.array([v["c"] for v in vehs], dtype=np.float64),
            off=off,
            s_node=G.dev_ids(np.array([s[2] for v in vehs for s in v["route"]], dtype=np.int32)),
            s_req=np.array([s[0] for v in vehs for s in v["route"]], dtype=np.int32),
            s_pod=np.array([s[1] for v in vehs for s in v["route"]], dtype=np.int8))
        a = self._a
        if have_pos:
            a["pos"] = nodes_or_minus_one([v.get("pos") for v in vehs])
        if have_keep:
            keeps = [v.get("keep") for v in vehs]
            a["keep_node"] = nodes_or_minus_one([None if k is None else k[0] for k in keeps])
            a["keep_t"] = np.array([0.0 if k is None else k[1] for k in keeps], dtype=np.float64)
            a["keep_et"] = np.array([0.0 if k is None else k[2] for k in keeps], dtype=np.float64)
        self._t = {k: (G.dev_ids(table[k]) if k in ("o", "d") else _lib.as_f64(table[k]))
                   for k in ("o", "d", "Cep", "Clp", "Cld", "Ts", "Tr", "pwt")}
        t = self._t
        self.n_rows = len(t["o"])
        pb = _lib.PlanBatch(nv, _lib.ptr_i32(a["node"]), _lib.ptr_f64(a["delay"]), _lib.ptr_i32(a["onboard"]),
                            _lib.ptr_f64(a["clock"]), _lib.ptr_i32(a["cap"]), _lib.ptr_f64(a["cost"]),
                            _lib.ptr_i32(a["off"]), _lib.ptr_i32(a["s_node"]), _lib.ptr_i32(a["s_req"]),
                            a["s_pod"].ctypes.data_as(_i8p),
                            _lib.ptr_i32(a["pos"]) if have_pos else None,
                            _lib.ptr_i32(a["keep_node"]) if have_keep else None,
                            _lib.ptr_f64(a["keep_t"]) if have_keep else None,
                            _lib.ptr_f64(a["keep_et"]) if have_keep else None)
        rt = _lib.RequestTable(len(t["o"]), _lib.ptr_i32(t["o"]), _lib.ptr_i32(t["d"]), _lib.ptr_f64(t["Cep"]),
                               _lib.ptr_f64(t["Clp"]), _lib.ptr_f64(t["Cld"]), _lib.ptr_f64(t["Ts"]),
                               _lib.ptr_f64(t["Tr"]), _lib.ptr_f64(t["pwt"]))
        p = params or {}
        prm = _lib.InsertionParams(p.get("max_detour", MAX_DETOUR), p.get("coef_inveh", COEF_INVEH),
                                   p.get("coef_wait", COEF_WAIT), p.get("coef_extra_wait", COEF_EXTRA_WAIT))
        _lib.check(self.lib.drs_tick_set_plans(self.handle, C.byref(pb), C.byref(rt), C.byref(prm)), self.lib)

    def _run(self, fn, new_rows, *extra):
        rows = _lib.as_i32(new_rows)
        n = len(rows)
        bv, bi, bj = (np.empty(n, dtype=np.int32) for _ in range(3))
        dc = np.empty(n, dtype=np.float64)
        nc = C.c_int64(0)
        _lib.check(fn(self.handle, n, _lib.ptr_i32(rows), *extra, _lib.ptr_i32(bv), _lib.ptr_i32(bi), _lib.ptr_i32(bj),
                      _lib.ptr_f64(dc), C.byref(nc)), self.lib)
        return bv, bi, bj, dc, nc.value

    def evaluate(self, new_rows):
        """Every new request against the plans AS THEY ARE (requests do not see each other).
        Returns (best_veh, best_i, best_j, best_dc, n_candidates) for the new-request rows."""
        return self._run(self.lib.drs_tick_insertion_eval, new_rows)

    def queue(self, new_rows, T_now, track_cld=True):
        """Model.insertion_heuristics (lib/Agents.py:1429-1448): the requests one after another, each seeing the plans
        the previous winners' build_route left behind; the plans stay patched on the device.  track_cld also leaves in
        the request table (fetch_plans()["Cld"]) the Req.Cld values the reference's scans assign as a side effect."""
        return self._run(self.lib.drs_tick_insertion_queue, new_rows, C.c_double(float(T_now)),
                         C.c_int32(_lib.QUEUE_TRACK_CLD if track_cld else 0))

    def commit(self, row, veh, i, j, T_now):
        _lib.check(self.lib.drs_tick_commit_insertion(self.handle, int(row), int(veh), int(i), int(j), float(T_now)), self.lib)

    def fetch_plans(self):
        """The plans as they are now: dict(c [n_veh], routes [list of (row, pod) per vehicle], Cld [rows], pwt [rows])."""
        nv, M = self.n_veh, self.MAX_PLAN
        c = np.zeros(nv, dtype=np.float64)
        ln = np.zeros(nv, dtype=np.int32)
        sreq = np.zeros((max(nv, 1), M), dtype=np.int32)
        spod = np.zeros((max(nv, 1), M), dtype=np.int8)
        cld = np.zeros(self.n_rows, dtype=np.float64)
        pwt = np.zeros(self.n_rows, dtype=np.float64)
        _lib.check(self.lib.drs_tick_fetch_plans(self.handle, _lib.ptr_f64(c), _lib.ptr_i32(ln), _lib.ptr_i32(sreq),
                                                 spod.ctypes.data_as(_i8p), M, _lib.ptr_f64(cld), _lib.ptr_f64(pwt)), self.lib)
        routes = [[(int(sreq[v, k]), int(spod[v, k])) for k in range(ln[v])] for v in range(nv)]
        return {"c": c, "routes": routes, "Cld": cld, "pwt": pwt}

    def last_ms(self):
        """Device times (ms) of the last evaluate() / queue() by kernel, and the pairs that survived the reach test."""
        ms = np.zeros(6, dtype=np.float64)
        _lib.check(self.lib.drs_tick_last_insertion_ms(self.handle, _lib.ptr_f64(ms)), self.lib)
        return {"prep": float(ms[0]), "reach": float(ms[1]), "eval": float(ms[2]), "fold": float(ms[3]), "queue": float(ms[4]),
                "scan": float(ms[0] + ms[1] + ms[2]), "survivors": int(ms[5])}

    def close(self):
        if self.handle:
            self.lib.drs_tick_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class InsertionHeuristics(object):
    """Object-level drop-in for ``Model.insertion_heuristics`` / ``insert_heuristics``
    (lib/Agents.py:1429-1490): holds ``vehs`` / ``reqs`` / ``queue`` / ``rejs`` like the legacy Model
    and assigns each queued request to the vehicle and (pickup, drop-off) slots the reference's
    scan would pick.  Vehicles are read through the reference's attribute surface (idle, route legs
    with rid/pod/tlng/tlat and steps[0].geo/t/et, c, n, T, K, lng, lat); requests through id, olng, olat,
    dlng, dlat, Cep, Clp, Cld, Ts, Tr, pwt.

    ``insertion_heuristics`` marshals the fleet ONCE per tick and runs the whole queue on the device
    (drs_tick_insertion_queue: every request sees the plans the previous winners' build_route left
    behind); the winners' ``build_route`` is then called on the host objects in queue order."""

    def __init__(self, vehs, reqs, params=None):
        from collections import deque
        self.vehs, self.reqs = vehs, reqs
        self.queue = deque()
        self.rejs = []
        self.params = params
        self._tick = None

    def _marshal(self, router, new_reqs):
        G = router.G
        rows, table = {}, {k: [] for k in ("o", "d", "Cep", "Clp", "Cld", "Ts", "Tr", "pwt")}
        self._row_reqs = []                                    # request object of every table row (Cld write-back)

        def row_of(rq):
            if rq.id not in rows:
                rows[rq.id] = len(table["o"])
                self._row_reqs.append(rq)
                table["o"].append(G.node_of_location(rq.olng, rq.olat))
                table["d"].append(G.node_of_location(rq.dlng, rq.dlat))
                for k in ("Cep", "Clp", "Cld", "Ts", "Tr", "pwt"):
                    table[k].append(float(getattr(rq, k)))
            return rows[rq.id]
        vehs = []
        for veh in self.vehs:
            keep = None
            if len(veh.route) > 0:                             # what a later build_route keeps (lib/Agents.py:319-336)
                st = veh.route[0].steps[0]
                keep = (G.node_of_location(st.geo[1][0], st.geo[1][1]), float(st.t), float(getattr(st, "et", st.t)))
            try:                                               # lib/Agents.py:1508-1516
                node = G.node_index(str((veh.lng, veh.lat)))
                t0, pos = 0.0, node
            except ValueError:
                assert keep is not None
                node, t0, pos = keep[0], keep[1], None
            route = []
            if not veh.idle:                                   # lib/Agents.py:1458-1461
                for leg in veh.route:
                    route.append((row_of(self.reqs[leg.rid]), int(leg.pod), G.node_of_location(leg.tlng, leg.tlat)))
            vehs.append({"node": node, "t0": t0, "n": int(veh.n), "T": float(veh.T), "K": int(veh.K), "c": float(veh.c),
                         "route": route, "keep": keep, "pos": pos})
        new_rows = [row_of(rq) for rq in new_reqs]
        return vehs, {k: np.asarray(v) for k, v in table.items()}, new_rows

    def _load(self, router, new_reqs):
        vehs, table, new_rows = self._marshal(router, new_reqs)
        if self._tick is None or self._tick.G is not router.G or not self._tick.handle:
            self._tick = InsertionTick(router.G, vehs, table, self.params)
        else:
            self._tick.set_plans(vehs, table, self.params)
        return new_rows

    def _assign(self, router, req, veh, i, j, T):
        route = []
        if not veh.idle:
            route = [(leg.rid, leg.pod, leg.tlng, leg.tlat) for leg in veh.route]
        elif getattr(veh, "_drs_route", None):
            route = list(veh._drs_route)                       # test doubles without build_route
        route.insert(int(i), (req.id, 1, req.olng, req.olat))
        route.insert(int(j), (req.id, -1, req.dlng, req.dlat))
        if hasattr(veh, "build_route"):
            veh.build_route(router, route, self.reqs, T, None)
        else:
            veh._drs_route = list(route)
        return route

    def _run_queue(self, router, reqs_q, T):
        new_rows = self._load(router, reqs_q)
        bv, bi, bj, dc, _ = self._tick.queue(new_rows, T, track_cld=True)
        # Req.Cld as the reference's scans leave it (lib/Agents.py:1550,1554): on-board riders of later ticks are checked
        # against it (lib/Agents.py:1555)
        cld = self._tick.fetch_plans()["Cld"]
        for row, rq in enumerate(self._row_reqs):
            rq.Cld = float(cld[row])
        return bv, bi, bj, dc

    def insert_heuristics(self, router, req, T):
        """lib/Agents.py:1451-1490.  Returns True when the request was assigned."""
        bv, bi, bj, dc = self._run_queue(router, [req], T)
        if bv[0] < 0:
            return False
        veh = self.vehs[int(bv[0])]
        route = self._assign(router, req, veh, bi[0], bj[0], T)
        self.last_choice = (veh, route, float(dc[0]))
        return True

    def insertion_heuristics(self, router, T):
        """lib/Agents.py:1429-1448 (DISTANCE_THRESHOLD = 0, lib/Constants.py:81): the whole queue in one device call."""
        reqs_q = [self.queue.popleft() for _ in range(len(self.queue))]
        if not reqs_q:
            return
        bv, bi, bj, dc = self._run_queue(router, reqs_q, T)
        self.last_results = []
        for k, req in enumerate(reqs_q):
            if bv[k] < 0:
                self.rejs.append(req)
                self.last_results.append((req, None, None, float("inf")))
            else:
                veh = self.vehs[int(bv[k])]
                route = self._assign(router, req, veh, bi[k], bj[k], T)
                self.last_choice = (veh, route, float(dc[k]))
                self.last_results.append((req, veh, route, float(dc[k])))
            req.assigned = True

    def close(self):
        if self._tick is not None:
            self._tick.close()
            self._tick = None
