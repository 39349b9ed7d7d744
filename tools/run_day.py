"""A FULL simulated day (2 880 ticks of 30 s) of BASELINE config C3 -- 50k-node city, 2 000 vehicles (K = 4), 200 000 requests per
day -- through the drop-in API only (bench.sim_section with ticks=2880): wall seconds per simulated day, measured, not extrapolated.
    python tools/run_day.py [--ticks 2880] > gpurun_out/r02_day.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from drs_abm_b200 import synth
from drs_abm_b200.routing import RoadGraph

ap = argparse.ArgumentParser()
ap.add_argument("--ticks", type=int, default=2880)
ap.add_argument("--config", default="C3")
args = ap.parse_args()
data = synth.config_graph(args.config)
t0 = time.time()
G = RoadGraph(data).build()
t_build = time.time() - t0
t0 = time.time()
out = bench.sim_section(G, data, args.config, ticks=args.ticks, cpu_legs=False)["sim"]
out["wall_seconds_measured"] = time.time() - t0
out["apsp_build_and_upload_s"] = t_build
out["ticks"] = args.ticks
print(json.dumps(out))
G.close()
