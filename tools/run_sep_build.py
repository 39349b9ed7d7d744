"""Three separator-partitioned builds of one synthetic config through the public drop-in (RoadGraph, AUTO), stage
times of the last one.  Run plain, then under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv ...                 (launch list)
    ncu --set full --clock-control none --import-source on -k regex:"sep_expand" --launch-skip 8 --launch-count 4 ...
(per build the kernels matching sep_expand launch in this order: vertex->separator rows and output tiles of the second
level, then of the map itself)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--builds", type=int, default=3)
args = ap.parse_args()
lib = _lib.load()
G = RoadGraph(synth.config_graph(args.config))
for _ in range(args.builds):
    G._valid = False
    G.build()
prof, work = np.zeros(8), np.zeros(5)
_lib.check(lib.drs_apsp_sep_profile(G._handle, _lib.ptr_f64(prof)), lib)
_lib.check(lib.drs_sep_work(G._handle, 0, 0, _lib.ptr_f64(work)), lib)
print(json.dumps({"config": args.config, "stage_ms": [round(float(x), 3) for x in prof], "total_ms": round(float(prof.sum()), 3),
                  "relaxations": [float(x) for x in work], "launches": int(lib.drs_launch_count())}))
G.close()
