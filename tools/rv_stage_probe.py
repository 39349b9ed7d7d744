"""RV-stage kernel times on the C4-style tick (10 000 vehicles x 2 000 requests on the C3 map) for settings of the
search split of the trip-evaluation kernels ("split2,split3" = stop counts from which a trip's search is split by
its first two / three stops).  Run on a GPU box: python tools/rv_stage_probe.py 5,8 7,9 99,99"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from drs_abm_b200 import _lib, synth
from drs_abm_b200.dispatch import Tick
from drs_abm_b200.routing import RoadGraph

shapes = sys.argv[1:] or ["5,8", "7,9", "99,99", "2,3"]
data = synth.config_graph("C3")
G = RoadGraph(data).build()
lib = _lib.load()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
vehs, reqs = bench.scenario_objects(data, sc)
rvt = Tick(G, vehs, reqs, sc["T"], pairwise_checks=True)
ms = np.zeros(4)
ref = None
for shape in shapes:
    os.environ["DRS_EVAL_SPLIT2"], os.environ["DRS_EVAL_SPLIT3"] = shape.split(",")
    times = []
    for it in range(4):
        n_edges = C.c_int32(0)
        counters = np.zeros(4, dtype=np.int64)
        _lib.check(lib.drs_tick_rv_eval(rvt.handle, float(sc["T"]), 1, C.byref(n_edges), counters.ctypes.data_as(_lib._i64p)), lib)
        _lib.check(lib.drs_tick_last_ms(rvt.handle, _lib.ptr_f64(ms)), lib)
        if it >= 1:
            times.append(float(ms[1]))
    ne = int(n_edges.value)
    veh = np.zeros(ne, dtype=np.int32); req = np.zeros(ne, dtype=np.int32); cost = np.zeros(ne)
    order = np.zeros((ne, 16), dtype=np.uint8); nst = np.zeros(ne, dtype=np.int32)
    sig = (ne, counters.tolist())
    print(json.dumps({"split2,split3": shape, "ms_eval": round(float(np.mean(times)), 4), "edges": ne, "counters": counters.tolist()}), flush=True)
    if ref is None:
        ref = sig
    assert sig == ref, "shapes disagree"
rvt.close()
G.close()
