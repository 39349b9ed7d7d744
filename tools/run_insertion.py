"""One C4-style insertion tick (10 000 vehicles x 2 000 requests on the C3 map): frozen-plan evaluation, twice.
Used under ncu (python tools/run_insertion.py) -- prints the per-kernel device times of the second call."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drs_abm_b200 import synth
from drs_abm_b200.dispatch import InsertionTick
from drs_abm_b200.routing import RoadGraph

data = synth.config_graph("C3")
G = RoadGraph(data).build()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
tick = InsertionTick(G, sc["vehs"], sc["table"])
for _ in range(2):
    bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
print(json.dumps(dict(tick.last_ms(), candidates=int(ncand))))
tick.close()
G.close()
