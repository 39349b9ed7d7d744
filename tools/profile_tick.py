"""cProfile of one Alonso-Mora tick through the drop-in API (2 000 vehicles x 70 new requests on the C3 map):
where the host time of `tick_e2e` goes.  Run on a GPU box: python tools/profile_tick.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drs_abm_b200 import synth
from drs_abm_b200.dispatch import AlonsoMora
from drs_abm_b200.routing import RoadGraph

data = synth.config_graph("C3")
G = RoadGraph(data).build()
sc = synth.fleet_scenario(data, 2000, 70, 600.0, 11, lambda s, t: G.matrix_lookup(s, t), plan_stops=(0, 0, 2, 4))
vehs, reqs = bench.scenario_objects(data, sc)
opt = AlonsoMora({"routing_method": "enumerate", "verbose": False, "max_trip_size": 2})
for _ in range(2):
    t0 = time.time()
    (routes, rejected), meta = opt.optimizeAssignment(vehs, reqs, G, sc["T"])
    print("tick %.4f s" % (time.time() - t0), {k: round(v, 4) for k, v in meta.items() if isinstance(v, float)})
pr = cProfile.Profile()
pr.enable()
opt.optimizeAssignment(vehs, reqs, G, sc["T"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
G.close()
