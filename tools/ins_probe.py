"""Kernel times of the C4-style tick (10 000 vehicles x 2 000 requests on the C3 map): insertion scan / fold and the
RV filter / evaluation.  Run on a GPU box: python tools/ins_probe.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from drs_abm_b200 import _lib, synth
from drs_abm_b200.dispatch import InsertionTick, Tick
from drs_abm_b200.routing import RoadGraph

data = synth.config_graph("C3")
G = RoadGraph(data).build()
lib = _lib.load()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
ms = np.zeros(4)
tick = InsertionTick(G, sc["vehs"], sc["table"])
scan, fold = [], []
for it in range(6):
    bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
    _lib.check(lib.drs_tick_last_ms(tick.handle, _lib.ptr_f64(ms)), lib)
    if it >= 2:
        scan.append(ms[2]); fold.append(ms[3])
tick.close()
vehs, reqs = bench.scenario_objects(data, sc)
rvt = Tick(G, vehs, reqs, sc["T"], pairwise_checks=True)
flt, ev = [], []
for it in range(5):
    n_edges = C.c_int32(0)
    counters = np.zeros(4, dtype=np.int64)
    _lib.check(lib.drs_tick_rv_eval(rvt.handle, float(sc["T"]), 1, C.byref(n_edges), counters.ctypes.data_as(_lib._i64p)), lib)
    _lib.check(lib.drs_tick_last_ms(rvt.handle, _lib.ptr_f64(ms)), lib)
    if it >= 1:
        flt.append(ms[0]); ev.append(ms[1])
rvt.close()
print(json.dumps({"scan_ms": round(float(np.mean(scan)), 4), "fold_ms": round(float(np.mean(fold)), 4), "assigned": int((bv >= 0).sum()),
                  "winner_checksum": int(np.sum(bv.astype(np.int64) * 7 + bi * 3 + bj)), "dc_sum": float(np.sum(dc[np.isfinite(dc)])),
                  "rv_filter_ms": round(float(np.mean(flt)), 4), "rv_eval_ms": round(float(np.mean(ev)), 4), "edges": int(n_edges.value)}))
G.close()
