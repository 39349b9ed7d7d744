#!/usr/bin/env python
"""Headline benchmark of the B200 routing-and-dispatch core (BASELINE.json metric:
"APSP travel-time matrix build (s) & insertion evals/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W

A step is one all-pairs travel-time-matrix build of the synthetic city named by --config (default C3, the 50k-node
map the metric's 1/2/4/8-GPU scaling is quoted on; it fits one GPU) with the CSR graph already resident in HBM.
The headline algorithm is the library default on integer-weighted road networks: the blocked min-plus Floyd-Warshall
run in nested-dissection order over a vertex-separator partition of the map (csrc/sep_apsp.cu, "sep": part closures,
separator-graph closure, every output tile one 128x128x|S_j| min-plus product written once; output rows sharded over
the ranks, no collective).  Round 1's headline, N batched delta-stepping searches, is timed beside it under "sssp",
and the plain blocked Floyd-Warshall the north star names under "fw" (N > 1: row-block shards, pivot panels
broadcast with NCCL, and the fused P2P-store pivot step under "fw_p2p"); all three write the same matrix.  Predecessors are lazy by default (--pred-matrix adds the predecessor pass to
the timed step; it is reported under phase_ms otherwise).  At N == 1 the same run also measures: scalar query
latencies through the drop-in, one insertion-stress tick (BASELINE config C4: 10k vehicles x 2k pending requests on
the same map: the (vehicle, pickup slot, drop-off slot) kernels with frozen plans AND the reference's sequential
queue), the Alonso-Mora RV stage, one assignment tick and a stretch of simulated operation with CPU legs of equal
scope; config C5 (200k-node SSSP, sources sharded) runs at every N.  Rank 0 prints ONE JSON line.

The oracle (oracle/) is used here only as the CPU baseline legs (cpu_baseline /
--impl reference); nothing measured on the GPU path touches it.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (C3 map, one
# B200, f32); reported as roofline.traffic only for the configuration the capture was taken on
NCU_TRAFFIC = {
    "sssp_batch_kernel": (9.401e9 + 22.049e9, "profiles/r02_kernels_ncu.md"),
    "insertion_reach_kernel": (0.403e9 + 0.007e9, "profiles/r02_kernels_ncu.md"),
    "rv_filter_kernel": (0.404e9 + 0.022e9, "profiles/r02_kernels_ncu.md"),
    "sep_expand16_kernel": (1.413e9 + 8.926e9, "profiles/r02_sep_ncu.md"),
    "sep_expand_kernel": (1.712e9 + 9.136e9, "profiles/r02_sep_ncu.md"),
}

WORKLOADS = {
    "C1": "C1: synthetic 20x20 grid GML (400 nodes, integer ttmedian 20..60 s)",
    "C2": "C2: synthetic 5k-node city GML (71x71 grid, keep 0.92, seed 5000)",
    "C3": "C3: synthetic 50k-node city GML (224x224 grid, keep 0.92, seed 50000), APSP sharded by source-row blocks",
}
INSERTION = {"C3": (10000, 2000), "C2": (200, 400), "C1": (10, 50)}   # vehicles x pending requests per tick


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc, self.path = None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, gpus):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        with open(self.path) as fh:
            for line in fh:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8 or not f[0].isdigit() or int(f[0]) >= gpus:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val == "Active":
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [x for x in sm if x > 0.5 * max(smax)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU legs
def _scipy_adjacency(data):
    from scipy.sparse import csr_matrix
    w = np.asarray(data.eattrs["ttmedian"], dtype=np.float64)
    return csr_matrix((np.concatenate([w, w]), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))),
                      shape=(data.n, data.n))


def _dijkstra_block(args):
    A, idx = args
    from scipy.sparse.csgraph import dijkstra
    t0 = time.time()
    d = dijkstra(A, directed=True, indices=idx)
    return time.time() - t0, float(d[np.isfinite(d)].sum())


def cpu_apsp_sample(data, n_sources, cores, seed=1):
    """Port of the reference's shortest-path work (one C Dijkstra per source, which is what igraph runs
    under every G.shortest_paths / get_shortest_paths call) on a bounded sample of sources.
    Returns (extrapolated full-APSP seconds, sources/s)."""
    A = _scipy_adjacency(data)
    idx = np.random.RandomState(seed).choice(data.n, size=min(n_sources, data.n), replace=False)
    if cores <= 1:
        dt = _dijkstra_block((A, idx))[0]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_dijkstra_block, [(A, part) for part in np.array_split(idx, cores) if len(part)])
        # the slowest worker's own compute time: process start-up and pickling are NOT charged to the reference
        dt = max(t for t, _ in res)
    rate = len(idx) / dt
    return data.n / rate, rate


def run_reference(args, rank, world):
    """--impl reference: the CPU path on the box's host cores, same metric / unit / config."""
    if rank != 0:
        return
    from drs_abm_b200 import synth
    data = synth.config_graph(args.config)
    cores = os.cpu_count() or 1
    per_step = max(cores * 128, 512)     # ~1-2 s of work per process on C3; only the workers' compute time is counted
    vals = []
    for it in range(args.warmup + args.steps):
        est, rate = cpu_apsp_sample(data, per_step, cores, seed=it)
        if it >= args.warmup:
            vals.append(est)
    v = float(np.mean(vals))
    # how good the extrapolation is: config C2 (5k nodes) is small enough to run in full
    c2 = synth.config_graph("C2")
    t_full, _ = cpu_apsp_sample(c2, c2.n, cores)
    est_c2, _ = cpu_apsp_sample(c2, max(cores * 16, 128), cores, seed=3)
    check = {"config": "C2", "n_nodes": c2.n, "full_run_s": t_full, "extrapolated_from_sample_s": est_c2,
             "note": "all sources of C2 timed in full next to the same sample-based estimate the C3 value uses"}
    sample = "%d of %d sources per step (scipy C Dijkstra, %d processes), extrapolated to all sources" % (per_step, data.n, cores)
    line = {"impl": "reference", "metric": "apsp_build_seconds", "value": v, "unit": "s", "extrapolated": True, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "n_nodes": data.n, "n_edges": data.m},
            "cpu_baseline": {"value": v, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
            "extrapolation_check": check,
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ insertion tick
def insertion_section(G, data, config, peaks, cpu_seconds=8.0):
    """One insertion-stress tick on the resident matrix (rank 0, N == 1)."""
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.dispatch import InsertionTick, Tick
    import ctypes as C
    n_veh, n_new = INSERTION[config]
    sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
    L = np.array([len(v["route"]) for v in sc["vehs"]], dtype=np.int64)
    lib = _lib.load()
    out = {"workload": "C4-style tick: %d vehicles (plans of L in {0,2,4,6,8} stops, K=4) x %d pending requests on the %s map"
           % (n_veh, n_new, config)}
    # ---- (vehicle, pickup slot, drop-off slot) kernels: device-resident timing, then end to end
    tick = InsertionTick(G, sc["vehs"], sc["table"])
    acc, ncand = [], 0
    for it in range(6):
        bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
        if it >= 3:
            acc.append(tick.last_ms())
    km = {k: float(np.mean([a[k] for a in acc])) for k in acc[0]}
    t_scan, t_fold = km["scan"] * 1e-3, km["fold"] * 1e-3
    pair_bytes = float(np.sum(16 * L + 20)) * n_new            # (16 L + 20) B per pair
    t0 = time.time()
    reps = 3
    for _ in range(reps):                                      # host arrays -> device (one arena copy) -> kernels -> winners on the host
        tick.set_plans(sc["vehs"], sc["table"])
        tick.evaluate(sc["new_rows"])
    e2e = (time.time() - t0) / reps
    tab = sc["table"]
    h2d = sum(tab[k].nbytes for k in tab) + int(L.sum()) * 9 + n_veh * 44 + n_new * 4
    out["insertion"] = {
        "metric": "insertion_evals_per_s", "value": ncand / (t_scan + t_fold), "unit": "candidates/s",
        "candidates_per_tick": int(ncand), "pairs_per_tick": int(n_veh) * int(n_new), "assigned": int((bv >= 0).sum()),
        "ms_scan_kernels": t_scan * 1e3, "ms_fold_kernel": t_fold * 1e3,
        "ms_by_kernel": {"plan_prep": km["prep"], "reach_test": km["reach"], "pair_eval": km["eval"], "fold": km["fold"]},
        "pairs_in_reach": km["survivors"],
        "roofline": {"bound": "hbm", "achieved": pair_bytes / t_scan / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": pair_bytes / t_scan / 1e9 / peaks["hbm_gbs"],
                     "traffic": NCU_TRAFFIC["insertion_reach_kernel"][0] if config == "C3" else None,
                     "traffic_source": NCU_TRAFFIC["insertion_reach_kernel"][1],
                     "kernel": "insertion_reach_kernel (+ plan_prep, insertion_eval_kernel): every (vehicle, i, j) candidate's cost",
                     "algorithmic_bytes_per_launch": pair_bytes,
                     "note": "SURVEY 8d bytes: (16 L + 20) per pair; the reach test stages each request's origin row in shared "
                             "memory by bulk async copy (UBLKCP) and settles ~99 % of the pairs from it"},
        "e2e": {"value": ncand / e2e, "unit": "candidates/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(n_new) * 20, "s_per_tick": e2e}}
    # ---- the reference's sequential queue (lib/Agents.py:1429-1448): each request sees the previous winners' plans
    seq = {}
    for label, track in (("with_cld_side_effects", True), ("winners_only", False)):
        tick.set_plans(sc["vehs"], sc["table"])
        t0 = time.time()
        qv, qi, qj, qdc, qcand = tick.queue(sc["new_rows"], sc["T"], track_cld=track)
        wall = time.time() - t0
        seq[label] = {"s_per_queue": wall, "ms_queue_kernel": tick.last_ms()["queue"], "candidates": int(qcand),
                      "candidates_per_s": qcand / wall, "assigned": int((qv >= 0).sum()),
                      "vehicles_that_won_more_than_once": int((qv >= 0).sum() - len(set(qv[qv >= 0].tolist())))}
    seq["note"] = ("%d requests inserted one after another into %d vehicles; wall time of the drs_tick_insertion_queue call with "
                   "host result buffers (plans already on the device)" % (n_new, n_veh))
    out["insertion"]["sequential_queue"] = seq
    tick.close()
    # ---- CPU baselines (1 core): the oracle's restatement of insert_heuristics on a bounded sample
    from oracle import dispatch_oracle as D
    sub_v = min(n_veh, 400)
    rows = {}                                                  # matrix rows fetched lazily

    def tt(a, b):
        r = rows.get(a)
        if r is None:
            r = rows[a] = G.dist_rows(int(a), 1)[0]
        return float(r[b])
    ovehs = [{"vid": k, "n": v["n"], "T": v["T"], "K": v["K"], "c": v["c"], "node": v["node"], "t0": v["t0"],
              "route": list(v["route"])} for k, v in enumerate(sc["vehs"][:sub_v])]
    used = sorted(set(s[0] for v in ovehs for s in v["route"]))

    def fresh(new_row):                                        # the scan mutates Cld: fresh records per request
        return {r: {"o_node": int(tab["o"][r]), "d_node": int(tab["d"][r]), "Cep": float(tab["Cep"][r]),
                    "Clp": float(tab["Clp"][r]), "Cld": float(tab["Cld"][r]), "Ts": float(tab["Ts"][r]),
                    "Tr": float(tab["Tr"][r]), "pwt": float(tab["pwt"][r])} for r in used + [new_row]}
    t0, evaluated, nreq = time.time(), 0, 0
    while time.time() - t0 < cpu_seconds and nreq < n_new:
        new_row = int(sc["new_rows"][nreq])
        _, _, _, ev = D.legacy_insert_heuristics(tt, ovehs, fresh(new_row), new_row)
        evaluated += ev
        nreq += 1
    dt = time.time() - t0
    out["insertion"]["cpu_baseline"] = {"value": evaluated / dt, "unit": "candidates/s", "cores": 1, "kind": "port",
                                        "sample": "%d requests x %d vehicles through the oracle's insert_heuristics "
                                                  "(travel times pre-fetched from the matrix, i.e. without the reference's "
                                                  "per-leg Dijkstra)" % (nreq, sub_v)}
    # the reference's own call pattern: router.get_duration is one Dijkstra per leg (lib/Agents.py:1543, lib/RoutingEngine.py:209-218)
    A = _scipy_adjacency(data)
    from scipy.sparse.csgraph import dijkstra
    ncalls = [0]

    def tt_dijkstra(a, b):
        ncalls[0] += 1
        return float(dijkstra(A, directed=True, indices=int(a))[int(b)])
    few = ovehs[:6]
    t0, evaluated2, nreq2 = time.time(), 0, 0
    while time.time() - t0 < cpu_seconds and nreq2 < 4:
        new_row = int(sc["new_rows"][nreq2])
        _, _, _, ev = D.legacy_insert_heuristics(tt_dijkstra, few, fresh(new_row), new_row)
        evaluated2 += ev
        nreq2 += 1
    dt2 = time.time() - t0
    out["insertion"]["cpu_ref_insert"] = {"value": evaluated2 / dt2, "unit": "candidates/s", "cores": 1, "kind": "port",
                                          "dijkstras": ncalls[0], "ms_per_dijkstra": dt2 / max(ncalls[0], 1) * 1e3,
                                          "sample": "%d requests x %d vehicles, one C Dijkstra (scipy, the stand-in for igraph's) per leg "
                                                    "of every candidate walk: the reference's call pattern" % (nreq2, len(few))}
    # ---- Alonso-Mora RV stage on the same state
    vehs, reqs = scenario_objects(data, sc)
    t0 = time.time()
    rvt = Tick(G, vehs, reqs, sc["T"], pairwise_checks=True)
    t_marshal = time.time() - t0
    rv_ms = []
    ms = np.zeros(4)
    for it in range(4):
        n_edges = C.c_int32(0)
        counters = np.zeros(4, dtype=np.int64)
        _lib.check(lib.drs_tick_rv_eval(rvt.handle, float(sc["T"]), 1, C.byref(n_edges), counters.ctypes.data_as(_lib._i64p)), lib)
        _lib.check(lib.drs_tick_last_ms(rvt.handle, _lib.ptr_f64(ms)), lib)
        if it >= 1:
            rv_ms.append((ms[0], ms[1]))
    rvt.close()
    f_ms, e_ms = float(np.mean([x[0] for x in rv_ms])), float(np.mean([x[1] for x in rv_ms]))
    pairs = int(n_veh) * int(n_new)
    out["rv_stage"] = {"metric": "rv_pairs_per_s", "value": pairs / ((f_ms + e_ms) * 1e-3), "unit": "pairs/s",
                       "pairs_per_tick": pairs, "ms_filter_kernel": f_ms, "ms_eval_kernel": e_ms,
                       "edges": int(n_edges.value), "skipped": int(counters[1]), "saved": int(counters[2]),
                       "infeasible": int(counters[3]), "host_marshal_s": t_marshal,
                       "roofline": {"bound": "hbm", "achieved": pairs * 12.0 / (f_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": pairs * 12.0 / (f_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "traffic": NCU_TRAFFIC["rv_filter_kernel"][0] if config == "C3" else None,
                                    "traffic_source": NCU_TRAFFIC["rv_filter_kernel"][1], "kernel": "rv_filter_kernel",
                                    "algorithmic_bytes_per_launch": pairs * 12.0}}
    return out


def query_latency_section(G, data, n_queries=2000, cpu_queries=12):
    """Scalar queries through the drop-in, one at a time, the way lib/Agents.py calls the router (get_duration at
    :898,904,1543; get_distance_duration at :756,1073; get_routing at :344; G.shortest_paths at lib/optimUtils.py:659):
    p50 / p99 wall time per call, next to the reference's call pattern on the CPU -- one C Dijkstra per query (scipy's,
    the stand-in for the igraph call behind every one of these methods; no cache, lib/RoutingEngine.py:87,199,212)."""
    from drs_abm_b200 import _lib
    from drs_abm_b200.routing import RoutingEngine
    from scipy.sparse.csgraph import dijkstra
    router = RoutingEngine(graph=G, cst_speed=9, seed=1)
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    rs = np.random.RandomState(12)
    o = rs.randint(0, data.n, n_queries)
    d = (o + 1 + rs.randint(0, data.n - 1, n_queries)) % data.n
    lib = _lib.load()

    def pct(fn, n):
        ts = np.empty(n)
        for k in range(n):
            t0 = time.perf_counter()
            fn(k)
            ts[k] = time.perf_counter() - t0
        return {"p50_us": float(np.percentile(ts, 50) * 1e6), "p99_us": float(np.percentile(ts, 99) * 1e6), "mean_us": float(ts.mean() * 1e6)}
    a0 = lib.drs_alloc_count()
    out = {"calls": n_queries,
           "get_duration": pct(lambda k: router.get_duration(lng[o[k]], lat[o[k]], lng[d[k]], lat[d[k]]), n_queries),
           "get_distance_duration": pct(lambda k: router.get_distance_duration(lng[o[k]], lat[o[k]], lng[d[k]], lat[d[k]]), n_queries),
           "get_routing": pct(lambda k: router.get_routing(lng[o[k]], lat[o[k]], lng[d[k]], lat[d[k]]), n_queries // 4),
           "G_shortest_paths": pct(lambda k: G.shortest_paths(int(o[k]), int(d[k]), weights="ttmedian"), n_queries)}
    out["device_allocations_during_the_calls"] = int(lib.drs_alloc_count() - a0)
    hops = G.path_sums(o, d)[2]
    out["mean_path_edges"] = float(hops.mean())
    A = _scipy_adjacency(data)
    t0 = time.perf_counter()
    for k in range(cpu_queries):
        dijkstra(A, directed=True, indices=int(o[k]))
    cpu = (time.perf_counter() - t0) / cpu_queries
    out["cpu_ref_query"] = {"value": cpu * 1e6, "unit": "us/query", "cores": 1, "kind": "port",
                            "sample": "%d single-source C Dijkstras (scipy) = one per get_duration / shortest_paths call in the "
                                      "reference; get_distance_duration is two" % cpu_queries}
    return {"query_latency": out}


def scenario_objects(data, sc):
    """Veh / Req objects with the reference's attribute surface from a synth.fleet_scenario dict."""
    tab = sc["table"]
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]

    class _R(object):
        def get_dropoff_time_window(self):
            return [self.Ced, self.Cld]

    class _V(object):
        def get_passenger_requests(self):
            return self._pax, self._wait

        def get_next_node(self):
            return self._loc, 0.0
    is_new = np.zeros(len(tab["o"]), dtype=bool)
    is_new[sc["new_rows"]] = True
    reqs = []
    for r in range(len(tab["o"])):
        q = _R()
        o, d = int(tab["o"][r]), int(tab["d"][r])
        q.id, q.olng, q.olat, q.dlng, q.dlat = r, lng[o], lat[o], lng[d], lat[d]
        q.Cep, q.Clp, q.Ced, q.Cld = float(tab["Cep"][r]), float(tab["Clp"][r]), float(tab["Ced"][r]), float(tab["Cld"][r])
        q.assigned = not is_new[r]
        reqs.append(q)
    vehs = []
    for k, v in enumerate(sc["vehs"]):
        o = _V()
        o.id, o.K, o._loc = k, v["K"], (lng[v["node"]], lat[v["node"]])
        o._pax, o._wait = set(v["pax"]), set(v["wait"])
        vehs.append(o)
    return vehs, reqs


def tick_e2e_section(G, data, config):
    """One assignment tick through the drop-in API (AlonsoMora.optimizeAssignment: marshalling, RV stage,
    RR / RTV trip costing on the GPU, set-packing on the host) at the fleet size of BASELINE config C3/C2,
    next to the oracle's restatement of the reference on a bounded sample of the same tick."""
    from drs_abm_b200 import synth
    from drs_abm_b200.dispatch import AlonsoMora
    from oracle import dispatch_oracle as D
    n_veh, n_new = {"C3": (2000, 70), "C2": (200, 7), "C1": (10, 5)}[config]
    sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 11, lambda s, t: G.matrix_lookup(s, t), plan_stops=(0, 0, 2, 4))
    vehs, reqs = scenario_objects(data, sc)
    opt = AlonsoMora({"routing_method": "enumerate", "verbose": False, "max_trip_size": 2})
    times = []
    for it in range(3):
        t0 = time.time()
        (routes, rejected), meta = opt.optimizeAssignment(vehs, reqs, G, sc["T"])
        times.append(time.time() - t0)
    # CPU: the oracle on the first `sub` vehicles, all new requests (pairs scale linearly with vehicles)
    sub = min(n_veh, 60)
    tab = sc["table"]
    rows = {}

    def tt(a, b):
        r = rows.get(a)
        if r is None:
            r = rows[a] = G.dist_rows(int(a), 1)[0]
        return float(r[b])

    def rec(r):
        return {"rid": r, "o_node": int(tab["o"][r]), "d_node": int(tab["d"][r]), "olng": 0.0, "olat": 0.0, "dlng": 0.0, "dlat": 0.0,
                "Cep": float(tab["Cep"][r]), "Clp": float(tab["Clp"][r]), "Ced": float(tab["Ced"][r]), "Cld": float(tab["Cld"][r])}
    ovehs = [{"vid": k, "capacity": v["K"], "node": v["node"], "location": (0.0, 0.0), "t": 0.0,
              "pax_reqs": [rec(r) for r in v["pax"]], "wait_reqs": [rec(r) for r in v["wait"]]}
             for k, v in enumerate(sc["vehs"][:sub])]
    oreqs = [rec(int(r)) for r in sc["new_rows"]]
    t0 = time.time()
    D.rv_graph(ovehs, oreqs, tt, sc["T"], with_rr=False)
    t_cpu = (time.time() - t0) * n_veh / sub
    return {"tick_e2e": {"workload": "%d vehicles x %d new requests, AlonsoMora(enumerate, trips <= 2) on the %s map" % (n_veh, n_new, config),
                         "value": float(np.median(times)), "unit": "s/tick", "assigned_vehicles": len(routes), "rejected": len(rejected),
                         "rv_time_s": meta.get("rv_time"), "rtv_time_s": meta.get("rtv_time"), "solve_time_s": meta.get("solve_time"),
                         "cpu_baseline": {"value": t_cpu, "unit": "s/tick", "cores": 1, "kind": "port",
                                          "sample": "oracle RV stage only (no RR / RTV / ILP) on %d of %d vehicles, travel times "
                                                    "pre-fetched from the matrix (no per-query Dijkstra), extrapolated" % (sub, n_veh)}}}


class _SimReq(object):
    def __init__(self, rid, o, d, lng, lat, Tr, Ts):
        self.id, self.o, self.d = rid, o, d
        self.olng, self.olat, self.dlng, self.dlat = lng[o], lat[o], lng[d], lat[d]
        self.Tr, self.Ts = Tr, Ts
        self.Cep, self.Clp = Tr, Tr + 600.0                       # lib/Agents.py:766-769 (MAX_WAIT, MAX_DETOUR)
        self.Ced, self.Cld = self.Cep + Ts, self.Clp + 1.5 * Ts
        self.assigned = False

    def get_dropoff_time_window(self):
        return [self.Ced, self.Cld]


class _SimVeh(object):
    """Vehicle state for the tick loop below: a plan of (rid, pod, node, arrival time).  Between two stops the
    vehicle is represented by the node of its next stop and the time left to reach it (the reference uses the next
    graph vertex on the path, lib/Agents.py:322-331; this coarser stand-in keeps the agents layer out of the bench)."""

    def __init__(self, vid, K, node, lng, lat):
        self.id, self.K, self.node, self.plan, self.onboard = vid, K, node, [], set()
        self._lng, self._lat = lng, lat

    def advance(self, T):
        while self.plan and self.plan[0][3] <= T:
            rid, pod, node, _ = self.plan.pop(0)
            self.node = node
            (self.onboard.add if pod > 0 else self.onboard.discard)(rid)

    def get_next_node(self):
        return (self._lng[self.node], self._lat[self.node]), 0.0

    def get_passenger_requests(self):
        return set(self.onboard), set(rid for (rid, pod, _, _) in self.plan if pod > 0)

    def position(self, T):
        """(node, delay): next stop and time to reach it, or the parked node."""
        if self.plan:
            return self.plan[0][2], max(0.0, self.plan[0][3] - T)
        return self.node, 0.0


def sim_section(G, data, config, ticks=40, cpu_legs=True):
    """A stretch of simulated operation through the drop-in API only: every 30 s new requests arrive, get their
    windows from the router, AlonsoMora.optimizeAssignment (trips <= 2) assigns them, the plans of the chosen vehicles
    are rebuilt from the matrix.  Fleet and demand of BASELINE config C3 / C2 (2 000 vehicles, 200k requests per day)."""
    from drs_abm_b200.dispatch import AlonsoMora
    n_veh, per_day = {"C3": (2000, 200000), "C2": (200, 10000), "C1": (10, 12000)}[config]
    rate = per_day * 30.0 / 86400.0
    rs = np.random.RandomState(17)
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    node_of = {}

    class V(_SimVeh):
        def get_next_node(self):
            nd, delay = self.position(self._T)
            return (lng[nd], lat[nd]), delay
    vehs = [V(k, 4, int(rs.randint(data.n)), lng, lat) for k in range(n_veh)]
    reqs = []
    opt = AlonsoMora({"routing_method": "enumerate", "verbose": False, "max_trip_size": 2})
    served = rejected = 0
    t_opt = t_all = 0.0
    T = 0.0
    # CPU leg of equal scope on a truncated horizon: the oracle's optimize_assignment on the SAME ticks (the first ones with
    # an empty fleet are cheap for both sides, so ticks spread over the run are taken)
    from oracle import dispatch_oracle as D
    cpu_ticks = set([2, ticks // 2, ticks - 1]) if cpu_legs else set()
    cpu_rows, n_tt, rows_cache = [], [0], {}

    def tt_counted(a, b):
        n_tt[0] += 1
        r = rows_cache.get(a)
        if r is None:
            r = rows_cache[a] = G.dist_rows(int(a), 1)[0]
        return float(r[b])
    for tick in range(ticks):
        t0 = time.time()
        T = 30.0 * tick
        for v in vehs:
            v._T = T
            v.advance(T)
        k = int(rs.poisson(rate))
        if k:
            o = rs.randint(0, data.n, k)
            d = (o + 1 + rs.randint(0, data.n - 1, k)) % data.n
            Ts = G.path_sums(o, d)[0]
            for i in range(k):
                if np.isfinite(Ts[i]):
                    reqs.append(_SimReq(len(reqs), int(o[i]), int(d[i]), lng, lat, T - float(rs.randint(0, 30)), float(Ts[i])))
        t1 = time.time()
        (routes, rej), meta = opt.optimizeAssignment(vehs, reqs, G, T)
        dt_gpu = time.time() - t1
        t_opt += dt_gpu
        if tick in cpu_ticks:
            # the same tick through the oracle's restatement of the reference (RV + RR + RTV + ILP, 1 core)
            def rec(rq):
                return {"rid": rq.id, "o_node": rq.o, "d_node": rq.d, "olng": rq.olng, "olat": rq.olat, "dlng": rq.dlng,
                        "dlat": rq.dlat, "Cep": rq.Cep, "Clp": rq.Clp, "Ced": rq.Ced, "Cld": rq.Cld}
            ovehs = []
            for v in vehs:
                pax, wait = v.get_passenger_requests()
                nd, delay = v.position(T)
                ovehs.append({"vid": v.id, "capacity": v.K, "node": nd, "location": (lng[nd], lat[nd]), "t": delay,
                              "pax_reqs": [rec(reqs[r]) for r in sorted(pax)], "wait_reqs": [rec(reqs[r]) for r in sorted(wait)]})
            oreqs = [rec(rq) for rq in reqs if rq.assigned is False]
            n_tt[0] = 0
            t2 = time.time()
            (oroutes, orej), ometa = D.optimize_assignment(ovehs, oreqs, tt_counted, T, max_trip_size=2)
            dt_cpu = time.time() - t2
            cpu_rows.append({"tick": tick, "new_requests": len(oreqs), "gpu_s": dt_gpu, "oracle_s": dt_cpu, "shortest_path_queries": n_tt[0],
                             "same_objective": bool(ometa["objective"] is None or meta.get("objective") is None or
                                                    abs(ometa["objective"] - meta["objective"]) <= 1e-6 * max(1.0, abs(ometa["objective"])))})
        if isinstance(routes, dict):
            by_loc = {}
            legs_s, legs_t, owners = [], [], []
            for vid, route in routes.items():
                veh = vehs[vid]
                cur, _ = veh.position(T)
                for (rid, pod, slng, slat) in route:
                    nd = G.node_of_location(slng, slat)
                    legs_s.append(cur); legs_t.append(nd); owners.append(vid); cur = nd
            tt = G.matrix_lookup(legs_s, legs_t) if legs_s else []
            pos = 0
            for vid, route in routes.items():
                veh = vehs[vid]
                _, delay = veh.position(T)
                t = T + delay
                plan = []
                for (rid, pod, slng, slat) in route:
                    t += float(tt[pos]); pos += 1
                    plan.append((rid, pod, G.node_of_location(slng, slat), t))
                    if pod > 0 and not reqs[rid].assigned:
                        reqs[rid].assigned = True
                        served += 1
                # the first stop's node stays the vehicle's anchor until it is reached
                veh.plan = plan
            for rid in rej:
                reqs[rid].assigned = True
                rejected += 1
        else:
            for r in reqs:                                   # empty RTV graph: nothing assignable this tick
                if r.assigned is False:
                    r.assigned = True
                    rejected += 1
        t_all += time.time() - t0
    cpu = None
    if cpu_rows:
        from scipy.sparse.csgraph import dijkstra
        A = _scipy_adjacency(data)
        t2 = time.time()
        for k in range(8):
            dijkstra(A, directed=True, indices=int(rs.randint(data.n)))
        t_dij = (time.time() - t2) / 8
        g_s = float(np.mean([r["gpu_s"] for r in cpu_rows]))
        o_s = float(np.mean([r["oracle_s"] for r in cpu_rows]))
        q = float(np.mean([r["shortest_path_queries"] for r in cpu_rows]))
        cpu = {"ticks_compared": cpu_rows, "cores": 1, "kind": "port",
               "same_objective_on_every_compared_tick": all(r["same_objective"] for r in cpu_rows),
               "s_per_tick_this_core": g_s, "s_per_tick_oracle_with_a_precomputed_matrix": o_s,
               "speedup_vs_oracle_with_a_precomputed_matrix": o_s / g_s,
               "shortest_path_queries_per_tick": q, "s_per_dijkstra": t_dij,
               "s_per_tick_reference_call_pattern_estimated": o_s + q * t_dij,
               "speedup_vs_reference_call_pattern_estimated": (o_s + q * t_dij) / g_s,
               "wall_seconds_per_simulated_day_reference_call_pattern_estimated": (o_s + q * t_dij) * 2880,
               "note": "the reference issues one igraph Dijkstra per shortest-path query (lib/optimUtils.py:528,531,659; no cache): its "
                       "tick = the oracle's Python time with travel times served from a precomputed matrix + (queries counted in the "
                       "oracle) x (one scipy C Dijkstra on this map, measured here); the first term alone is a CPU baseline the "
                       "reference does not have"}
    return {"sim": {"cpu_baseline": cpu, "workload": "%d ticks of 30 s, %d vehicles (K=4), %.0f requests/day, AlonsoMora(enumerate, trips <= 2) on the %s map"
                                % (ticks, n_veh, per_day, config),
                    "requests": len(reqs), "served": served, "rejected": rejected,
                    "s_per_tick": t_all / ticks, "s_per_tick_in_optimizer": t_opt / ticks,
                    "simulated_seconds_per_wall_second": 30.0 * ticks / t_all,
                    "wall_seconds_per_simulated_day": t_all / ticks * 2880,
                    "note": "agents layer replaced by a minimal plan follower (vehicles are anchored at their next stop); "
                            "routing and dispatch go through the public drop-in API only"}}


def sssp_c5_section(peaks, rank=0, world=1, dist=None, n_total=32768):
    """BASELINE config C5: 200k-node metro CSR, batched delta-stepping SSSP only (no N x N matrix).  The sources are
    sharded across the ranks (independent searches, no collective); every rank calls this, rank 0 gets the dict."""
    import ctypes as C
    import torch
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.routing import RoadGraph
    lib = _lib.load()
    data = synth.config_graph("C5")
    G = RoadGraph(data)
    h = G.upload()
    npad = (data.n + 127) // 128 * 128
    src_all = np.random.RandomState(5).choice(data.n, size=n_total, replace=False).astype(np.int32)
    src = np.array_split(src_all, world)[rank]
    n_sources = len(src)
    src_dev = G.dev_ids(src)                                   # the raw library call works in the library's vertex numbering
    buf = torch.empty((n_sources, npad), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    times = []
    for it in range(5):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.check(lib.drs_sssp_batch_dev(h, _lib.MODE_F32, n_sources, _lib.ptr_i32(src_dev), C.c_void_p(buf.data_ptr()), npad, 0.0,
                                          C.c_void_p(stream.cuda_stream)), lib)
        ev1.record(); torch.cuda.synchronize()
        if it >= 2:
            times.append(ev0.elapsed_time(ev1) * 1e-3)
    tm = torch.tensor([float(np.mean(times))], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)              # the job is done when the slowest rank is
    t = float(tm.item())
    narcs = 2 * data.m
    bytes_alg = (16.0 * narcs + 8.0 * data.n) * n_total
    out = None
    if rank == 0:
        # spot check against scipy on two sources (exact: integer weights)
        from scipy.sparse.csgraph import dijkstra
        ref = dijkstra(_scipy_adjacency(data), directed=True, indices=src[:2])
        ok = bool(np.array_equal(G.host_cols(buf[:2].double().cpu().numpy()), ref))
        t0 = time.time()
        dijkstra(_scipy_adjacency(data), directed=True, indices=src[:16])
        cpu_rate = 16 / (time.time() - t0)
        out = {"sssp_c5": {"workload": "C5: synthetic 200k-node metro CSR (448x448 grid, keep 0.90), %d sources sharded over %d GPU(s), "
                                       "no full APSP" % (n_total, world),
                           "n_nodes": data.n, "n_arcs": narcs, "n_gpus": world, "metric": "sssp_sources_per_s", "value": n_total / t,
                           "unit": "sources/s", "scaling": "strong", "arc_relaxations_per_s_lower_bound": n_total * narcs / t, "ms": t * 1e3,
                           "matches_scipy_dijkstra": ok,
                           "roofline": {"bound": "hbm", "achieved": bytes_alg / t / 1e9 / world, "peak": peaks["hbm_gbs"], "unit": "GB/s per GPU",
                                        "frac": bytes_alg / t / 1e9 / world / peaks["hbm_gbs"], "traffic": None, "kernel": "sssp_batch_kernel"},
                           "cpu_baseline": {"value": cpu_rate, "unit": "sources/s", "cores": 1, "kind": "port",
                                            "sample": "16 sources with scipy's C Dijkstra"}}}
    del buf
    G.close()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="f32", choices=["f32", "i32"])
    ap.add_argument("--algorithm", default="auto", choices=["auto", "fw", "sssp", "sep"],
                    help="APSP algorithm of the headline value; auto = the library default (SSSP on sparse road networks). "
                         "The blocked Floyd-Warshall is always measured too and reported under \"fw\".")
    ap.add_argument("--pred-matrix", action="store_true",
                    help="also materialise the predecessor matrix inside the timed build (default: lazy predecessors -- path "
                         "queries evaluate the canonical rule per hop; the matrix pass is reported separately under phase_ms)")
    ap.add_argument("--no-insertion", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="skip the fused-broadcast FW measurement at N > 1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)" % (args.gpus, args.gpus))
    import torch
    import torch.distributed as dist
    from drs_abm_b200 import _lib, apsp_dist, synth
    from drs_abm_b200.routing import RoadGraph
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the routing core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly ONE JSON line: everything native code prints meanwhile (NCCL's version banner comes from C, on
    # file descriptor 1, from every rank) goes to stderr; rank 0 gets the descriptor back for the line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    peaks, peaks_src = measured_peaks()
    mode = _lib.MODE_F32 if args.mode == "f32" else _lib.MODE_I32
    data = synth.config_graph(args.config)
    G = RoadGraph(data, mode=mode)
    handle = G.upload()                                   # CSR resident in HBM before the timed region
    npad = (data.n + 127) // 128 * 128
    tiles = apsp_dist.CudaTiles(handle, mode)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sparse = 2 * data.m <= 16 * data.n                    # the library's AUTO rule (drs_apsp_build)
    import ctypes as C
    sep_info = [C.c_int32(0) for _ in range(8)]
    _lib.check(lib.drs_sep_info(handle, *[C.byref(x) for x in sep_info[:5]]), lib)
    _lib.check(lib.drs_sep_info2(handle, C.byref(sep_info[5]), C.byref(sep_info[6]), None), lib)
    sep_usable = bool(sep_info[4].value)
    algo = args.algorithm if args.algorithm != "auto" else ("sep" if sep_usable else "sssp" if sparse else "fw")

    comm = [None]

    def one_build(algorithm=None):
        a = algorithm or algo
        if a == "fw_p2p":        # blocked FW whose pivot step stores the panel into the peers' buffers (fused broadcast)
            if comm[0] is None:
                comm[0] = apsp_dist.P2PComm(tiles, npad, rank, world, dist)
            return apsp_dist.build_sharded_p2p(tiles, npad, rank, world, dist, comm[0], lookahead=not args.no_lookahead,
                                               with_pred=args.pred_matrix)
        return apsp_dist.build_sharded(tiles, npad, rank, world, dist if world > 1 else None,
                                       lookahead=not args.no_lookahead, with_pred=args.pred_matrix, algorithm=a)

    # ALU roofline denominator, measured live (rank 0): best add+min stream of the three instruction mixes
    alu = {}
    if rank == 0:
        for v, name in ((0, "fadd_fadd_fmnmx3"), (1, "viaddmnmx_s32"), (2, "fadd2_fmnmx3"), (3, "viaddmnmx_u16x2")):
            r = __import__("ctypes").c_double()
            _lib.check(lib.drs_alu_peak(v, 8000, __import__("ctypes").byref(r)), lib)
            alu[name] = r.value
    def timed_builds(algorithm, warmup, steps):
        """W warm-up builds, then `steps` builds between barriers; device time by CUDA events, max over ranks."""
        for _ in range(warmup):
            res = one_build(algorithm)
            del res
        l0 = lib.drs_launch_count()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        keep = local = None
        for _ in range(steps):
            del keep, local                                  # never hold two matrices: the allocator reuses the block
            local, pred, row0 = one_build(algorithm)
            del pred
            keep = local
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        nl = torch.tensor([lib.drs_launch_count() - l0], device="cuda", dtype=torch.int64)
        fin = torch.isfinite(keep) if mode == _lib.MODE_F32 else (keep < _lib.I32_INF)
        cs = torch.where(fin, keep.double(), torch.zeros((), device="cuda", dtype=torch.float64)).sum().reshape(1)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(nl, op=dist.ReduceOp.SUM)
            dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        del keep, local
        torch.cuda.empty_cache()
        return float(ms.item()) * 1e-3 / steps, int(nl.item()), float(cs.item())

    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    sec, launches, checksum = timed_builds(algo, args.warmup, args.steps)
    clocks = sampler.stop(args.gpus) if rank == 0 else None
    # the other algorithm, for the record (the blocked FW is the north-star path; SSSP the sparse-graph default)
    other = "fw" if algo in ("sssp", "sep") else ("sssp" if sparse else None)
    other_res = None
    if other is not None:
        sampler2 = ClockSampler()
        if rank == 0:
            sampler2.start()
        other_res = timed_builds(other, 3 if other == "sssp" else 1, max(1, min(args.steps, 2)))
        other_clocks = sampler2.stop(args.gpus) if rank == 0 else None
    sssp_beside = timed_builds("sssp", 2, max(1, min(args.steps, 3))) if algo == "sep" else None
    p2p_res = None
    if world > 1 and not args.no_p2p:
        p2p_res = timed_builds("fw_p2p", 1, max(1, min(args.steps, 2)))

    # end to end through the public API: host arrays in, query results out (median of three fresh graphs)
    e2e_runs = []
    G2 = None
    for rep in range(3):
        if G2 is not None:
            G2.close()
        if rep == 0:
            torch.cuda.empty_cache()                  # the timed builds' shards go back to the driver once; later repetitions
        barrier()                                     # reuse the allocator's (torch) and the library's (drs_pool) cached blocks
        t0 = time.time()
        G2 = RoadGraph(data, mode=mode, algorithm={"fw": _lib.APSP_FW, "sssp": _lib.APSP_SSSP, "sep": _lib.APSP_SEP}[algo],
                       pred=_lib.PRED_MATRIX if args.pred_matrix else _lib.PRED_LAZY)
        h2 = G2.upload()
        o, d = synth.random_od_pairs(data.n, 4096, 3)
        if world == 1:
            G2.build()
            tt, ln, hops = G2.path_sums(o, d)
            d2h = tt.nbytes + ln.nbytes + hops.nbytes
        else:
            loc2, pred2, r02 = apsp_dist.build_sharded(apsp_dist.CudaTiles(h2, mode), npad, rank, world, dist,
                                                       with_pred=args.pred_matrix, algorithm=algo)
            sample = loc2[:8].cpu()
            d2h = sample.numel() * sample.element_size()
            del loc2, pred2
        barrier()
        e2e_runs.append(time.time() - t0)
    e2e_t = torch.tensor([float(np.median(e2e_runs))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    h2d = data.src.nbytes + data.dst.nbytes + 2 * data.m * 8 + (o.nbytes + d.nbytes if world == 1 else 0)

    # BASELINE config C5 (200k-node SSSP, sources sharded over the ranks): every rank takes part
    c5 = None
    if not args.no_insertion and args.config == "C3":
        c5 = sssp_c5_section(peaks, rank, world, dist if world > 1 else None)
    if rank == 0:
        import ctypes as C
        relax = float(data.n) ** 3
        best_alu = max(v for k, v in alu.items() if k != "viaddmnmx_u16x2")    # 32-bit relaxations (FW tiles, 32-bit SEP tiles)
        narcs = 2 * data.m
        sssp_bytes = (16.0 * narcs + 8.0 * data.n) * data.n          # 16 E_dir + 8 N bytes per source

        def fw_roofline(t_gpu_seconds):
            return {"bound": "alu", "achieved": 2 * relax / t_gpu_seconds / 1e12, "peak": 2 * best_alu / 1e12,
                    "unit": "Tlaneop/s", "frac": (relax / t_gpu_seconds) / best_alu, "traffic": None,
                    "kernel": "fw_minplus_tile_kernel",
                    "note": "2 lane-ops (add + min) per relaxation x n^3 relaxations over the GPU-seconds of the "
                            "tile kernel; peak = best register-resident add+min stream measured in this run (drs_alu_peak)"}

        def sssp_roofline(t_gpu_seconds):
            return {"bound": "hbm", "achieved": sssp_bytes / t_gpu_seconds / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": sssp_bytes / t_gpu_seconds / 1e9 / peaks["hbm_gbs"],
                    "traffic": NCU_TRAFFIC["sssp_batch_kernel"][0] if (args.config == "C3" and world == 1 and args.mode == "f32") else None,
                    "traffic_source": NCU_TRAFFIC["sssp_batch_kernel"][1],
                    "kernel": "sssp_batch_kernel", "algorithmic_bytes_per_launch": sssp_bytes,
                    "note": "(16 E_dir + 8 N) bytes per source, served mostly by L1/L2: ncu shows L1/TEX at 70 % and L2 at 59 % of peak "
                            "throughput, DRAM 31 GB per build (3.3x the matrix: the ~1300 rows in flight exceed the 126 MB L2) -- bound by "
                            "the number of searches an SM can keep in flight (each a chain of ~500 dependent wavefront steps), not by HBM "
                            "bandwidth; the row-in-shared-memory and the full-occupancy lane-per-arc kernels were built and measured: "
                            "profiles/r02_sssp_step_profile.md"}
        line = {"metric": "apsp_build_seconds", "value": sec, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
                "config": {"workload": WORKLOADS[args.config], "n_nodes": data.n, "n_edges": data.m, "npad": npad,
                           "predecessors": "matrix built inside the timed step" if args.pred_matrix else
                                           "lazy (canonical rule evaluated per hop by the path queries; no predecessor matrix)",
                           "apsp_algorithm": {"sssp": "batched delta-stepping SSSP, one CTA per source (library default on sparse graphs with non-integer travel times)",
                                              "fw": "blocked min-plus Floyd-Warshall",
                                              "sep": "blocked min-plus Floyd-Warshall in nested-dissection order over a vertex-separator partition "
                                                     "(%d parts, %d separator vertices, second level: %d groups + %d top separators): part closures, separator-graph "
                                                     "closure, then every output tile as ONE 128x128x|S_j| min-plus product, written once; mirrored tiles on "
                                                     "undirected maps at N = 1 (library default on integer-weighted maps: bit-identical to SSSP / FW)"
                                                     % (sep_info[0].value, sep_info[1].value, sep_info[5].value, sep_info[6].value)}[algo],
                           "l2": "inputs larger than L2: the %d MB matrix is written once (SEP, SSSP) / streamed once per round (FW); L2 is 126 MB" % (npad * npad * 4 >> 20),
                           "parallelism": ("source rows sharded x%d, no collective" % world) if algo == "sssp" else
                                          ("part and separator closures replicated on every rank, output row blocks sharded x%d, no collective" % world) if algo == "sep" else
                                          "row-block shards x%d, NCCL pivot-panel broadcast%s" % (world, "" if args.no_lookahead else " with look-ahead")},
                "gpu_launches": launches, "clocks": clocks, "checksum": checksum, "alu_peak_relax_per_s": alu,
                "peaks": peaks_src}
        prof_fw = prof_sssp = prof_sep = None
        if algo == "sep":
            # stage times of rank 0's last timed build (the library's own CUDA events) and the relaxations its shard executes
            prof_sep, work_sep = np.zeros(8), np.zeros(5)
            starts = apsp_dist.block_partition(npad // 128, world)
            _lib.check(lib.drs_apsp_sep_profile(handle, _lib.ptr_f64(prof_sep)), lib)
            _lib.check(lib.drs_sep_work(handle, starts[0] * 128, (starts[1] - starts[0]) * 128, _lib.ptr_f64(work_sep)), lib)
            u16 = C.c_int32(0)
            _lib.check(lib.drs_sep_info2(handle, None, None, C.byref(u16)), lib)
            sep_used16 = bool(u16.value)
        if world == 1:
            # per-kernel device times from the library's own CUDA events (public single-GPU build)
            def profile(algorithm):
                Gp = RoadGraph(data, mode=mode, algorithm=algorithm, pred=_lib.PRED_MATRIX)
                Gp.build(); Gp._valid = False; Gp.build()
                prof = np.zeros(4); rounds = C.c_int32(0)
                _lib.check(lib.drs_apsp_last_profile(Gp._handle, _lib.ptr_f64(prof), C.byref(rounds)), lib)
                Gp.close()
                return prof
            if algo == "sssp" or other == "sssp":
                prof_sssp = profile(_lib.APSP_SSSP)
            prof_fw = profile(_lib.APSP_FW)
        if algo == "sep":
            line["phase_ms"] = dict(zip(["fill_scatter", "part_closures", "extract", "separator_closure", "gathers",
                                         "vertex_to_separator_rows", "output_tiles", "separator_columns"], [float(x) for x in prof_sep])) \
                if prof_sep is not None else None
            if prof_sep is not None:
                t_k = prof_sep[6] * 1e-3
                tile_stages_ms = float(prof_sep[1] + prof_sep[3] + prof_sep[5] + prof_sep[6])
                sep_kernel = "sep_expand16_kernel" if sep_used16 else "sep_expand_kernel"
                sep_peak = alu["viaddmnmx_u16x2"] if sep_used16 else best_alu
                line["roofline"] = {
                    "bound": "alu", "achieved": 2 * work_sep[4] / t_k / 1e12, "peak": 2 * sep_peak / 1e12, "unit": "Tlaneop/s",
                    "frac": (work_sep[4] / t_k) / sep_peak, "executed_frac": (work_sep[3] / t_k) / sep_peak,
                    "frac_of_the_32_bit_peak": (work_sep[4] / t_k) / best_alu,
                    "kernel": sep_kernel, "kernel_ms": float(prof_sep[6]), "share_of_step": float(prof_sep[6]) / float(prof_sep.sum()),
                    "second_bound": {"bound": "hbm", "what": "the kernel writes its rows of the matrix once", "bytes": 4.0 * npad * npad / world,
                                     "achieved_GBps": 4.0 * npad * npad / world / t_k / 1e9,
                                     "frac": 4.0 * npad * npad / world / t_k / 1e9 / peaks["hbm_gbs"]},
                    "algorithmic_relaxations_per_launch": float(work_sep[4]), "executed_relaxations_per_launch": float(work_sep[3]),
                    "traffic": NCU_TRAFFIC.get(sep_kernel, (None, None))[0] if (args.config == "C3" and args.mode == "f32" and world == 1) else None,
                    "traffic_source": NCU_TRAFFIC.get(sep_kernel, (None, None))[1],
                    "hbm": {"compulsory_bytes": 4.0 * npad * npad, "GBps_over_the_whole_build": 4.0 * npad * npad / sec / 1e9,
                            "frac_of_peak": 4.0 * npad * npad / sec / 1e9 / peaks["hbm_gbs"]},
                    "all_tile_stages": {"relaxations": float(work_sep[0] + work_sep[1] + work_sep[2] + work_sep[3]), "ms": tile_stages_ms,
                                        "frac": float(work_sep[0] + work_sep[1] + work_sep[2] + work_sep[3]) / (tile_stages_ms * 1e-3) / best_alu},
                    "rank": 0, "rows": int((starts[1] - starts[0]) * 128),
                    "note": "min-plus is not a tensor-core operation (north_star): the bound is the add+min issue rate of the SMs.  2 lane-ops "
                            "per relaxation; algorithmic relaxations of the dominant kernel = n * sum_j n_j * |S_j| (halved: mirrored tiles on the "
                            "undirected map), executed = the same with |S_j| padded to 16 and part widths to 128-column tiles; peak = the "
                            "register-resident add+min stream of the instruction the tiles use, measured in this run (drs_alu_peak: "
                            "VIADDMNMX.U16x2 when the map's distances stay below 2^15 and the tiles run 16-bit, else the best 32-bit mix); "
                            "at that rate the 10 GB matrix write becomes the next bound (second_bound)"}
        elif algo == "sssp":
            t_k = prof_sssp[2] * 1e-3 if prof_sssp is not None else sec * world
            line["roofline"] = sssp_roofline(t_k)
            if prof_sssp is not None:
                line["phase_ms"] = {"sssp": prof_sssp[2], "pred_matrix_optional": prof_sssp[3]}
                # SURVEY's M1 counts the predecessor matrix in the build; it is optional here (lazy predecessors by default)
                line["value_with_pred"] = sec + prof_sssp[3] * 1e-3
        else:
            t_k = prof_fw[2] * 1e-3 if prof_fw is not None else sec * world
            line["roofline"] = fw_roofline(t_k)
        if other_res is not None:
            osec, olaunch, ocs = other_res
            sub = {"value": osec, "unit": "s", "gpu_launches": olaunch, "checksum": ocs, "clocks": other_clocks,
                   "same_matrix_as_headline": bool(ocs == checksum)}
            if other == "fw":
                t_k = prof_fw[2] * 1e-3 if prof_fw is not None else osec * world
                sub["roofline"] = fw_roofline(t_k)
                if prof_fw is not None:
                    sub["phase_ms"] = {"init": prof_fw[0], "pivot_row": prof_fw[1], "update": prof_fw[2], "pred": prof_fw[3]}
                line["fw"] = sub
            else:
                t_k = prof_sssp[2] * 1e-3 if prof_sssp is not None else osec * world
                sub["roofline"] = sssp_roofline(t_k)
                line["sssp"] = sub
        if sssp_beside is not None:
            ssec, slaunch, scs = sssp_beside
            line["sssp"] = {"value": ssec, "unit": "s", "gpu_launches": slaunch, "checksum": scs, "same_matrix_as_headline": bool(scs == checksum),
                            "roofline": sssp_roofline(ssec * world),
                            "note": "round 1's headline builder on the same device graph (it keeps float-weighted maps and C5)"}
        if p2p_res is not None:
            psec, plaunch, pcs = p2p_res
            line["fw_p2p"] = {"value": psec, "unit": "s", "gpu_launches": plaunch, "checksum": pcs,
                              "same_matrix_as_headline": bool(pcs == checksum), "roofline": fw_roofline(psec * world),
                              "transport": "pivot-panel tiles stored straight into the peers' buffers by the tile kernel's epilogue "
                                           "(CUDA IPC peer memory over NVLink), signal / acknowledgement words instead of ncclBroadcast"}
        line["e2e"] = {"value": float(e2e_t.item()), "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "runs_s": [float(x) for x in e2e_runs], "note": "median of three: fresh RoadGraph from host edge arrays -> build -> 4096 path queries on the host"}
        if world == 1 and not args.no_cpu_baseline:
            ns = 256 if data.n > 10000 else data.n
            est, rate = cpu_apsp_sample(data, ns, 1)
            line["cpu_baseline"] = {"value": est, "unit": "s", "cores": 1, "kind": "port",
                                    "sample": "%d of %d sources with scipy's C Dijkstra (the stand-in for igraph's), extrapolated" % (min(ns, data.n), data.n)}
        if world == 1 and not args.no_insertion:
            if not G2._valid:
                G2.build()
            line.update(query_latency_section(G2, data))
            line.update(insertion_section(G2, data, args.config, peaks))
            line.update(tick_e2e_section(G2, data, args.config))
            line.update(sim_section(G2, data, args.config))
        if c5 is not None:
            line.update(c5)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if comm[0] is not None:
        comm[0].close()
    G2.close()
    G.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
