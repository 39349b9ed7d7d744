import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, bench
from drs_abm_b200 import synth
from drs_abm_b200.dispatch import InsertionTick
from drs_abm_b200.routing import RoadGraph
data = synth.config_graph("C3")
G = RoadGraph(data).build()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
for track in (True, False):
    tick = InsertionTick(G, sc["vehs"], sc["table"])
    t0 = time.time(); r = tick.queue(sc["new_rows"], 600.0, track_cld=track); w = time.time() - t0
    print(json.dumps({"track_cld": track, "wall_ms": round(w * 1e3, 2), "queue_ms": round(tick.last_ms()["queue"], 2)}), flush=True)
    tick.close()
