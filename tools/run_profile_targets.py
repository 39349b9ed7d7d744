"""The kernels the round-2 ncu captures look at, each launched on its full-size workload (C3 map; C4 tick):
one APSP build (sssp_batch_kernel), two insertion ticks (plan_prep / insertion_reach / insertion_eval / insertion_fold),
two RV stages (rv_filter / rv_eval).  Run plain, then under
    ncu --set full --clock-control none --import-source on -k regex:"sssp_batch|insertion_reach|insertion_eval|insertion_fold|rv_filter|rv_eval" ..."""
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

lib = _lib.load()
data = synth.config_graph("C3")
G = RoadGraph(data).build()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
tick = InsertionTick(G, sc["vehs"], sc["table"])
for _ in range(2):
    bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
out = dict(tick.last_ms(), candidates=int(ncand))
tick.close()
vehs, reqs = bench.scenario_objects(data, sc)
rvt = Tick(G, vehs, reqs, sc["T"], pairwise_checks=True)
ms = np.zeros(4)
for _ in range(2):
    n_edges = C.c_int32(0)
    counters = np.zeros(4, dtype=np.int64)
    _lib.check(lib.drs_tick_rv_eval(rvt.handle, float(sc["T"]), 1, C.byref(n_edges), counters.ctypes.data_as(_lib._i64p)), lib)
_lib.check(lib.drs_tick_last_ms(rvt.handle, _lib.ptr_f64(ms)), lib)
rvt.close()
out.update(rv_filter_ms=float(ms[0]), rv_eval_ms=float(ms[1]), rv_edges=int(n_edges.value))
print(json.dumps(out))
G.close()
