"""Phase timing of the end-to-end call bench.py reports as `e2e` (host arrays in -> query results out).
Run on a GPU box: python tools/e2e_breakdown.py [--config C3]"""
import argparse, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from drs_abm_b200 import synth, _lib
from drs_abm_b200.routing import RoadGraph

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
args = ap.parse_args()
data = synth.config_graph(args.config)
torch.cuda.init()
out = []
for rep in range(4):
    torch.cuda.synchronize()
    t = [time.time()]
    G = RoadGraph(data, mode=_lib.MODE_F32); G._number(); t.append(time.time())
    G.upload(); torch.cuda.synchronize(); t.append(time.time())
    o, d = synth.random_od_pairs(data.n, 4096, 3); t.append(time.time())
    G.build(); torch.cuda.synchronize(); t.append(time.time())
    import ctypes as C
    pm = np.zeros(4); rounds = C.c_int32(0)
    G._lib.drs_apsp_last_profile(G._handle, _lib.ptr_f64(pm), C.byref(rounds))
    prof = [round(float(x), 3) for x in pm]
    tt, ln, hops = G.path_sums(o, d); t.append(time.time())
    G.close(); torch.cuda.synchronize(); t.append(time.time())
    names = ["ctor_and_numbering", "upload_create_set_partition", "od_pairs", "build", "path_sums", "close"]
    out.append({n: round((b - a) * 1e3, 3) for n, a, b in zip(names, t[:-1], t[1:])} | {"profile": prof})
print(json.dumps(out, indent=1))
