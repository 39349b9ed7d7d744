"""Round-2 insertion probe on a GPU box (C4: 10 000 vehicles x 2 000 requests on the C3 map): per-kernel device times of
the frozen-plan evaluation (staged and unstaged reach test) and of the sequential queue; results of the variants must agree.
    python tools/ins_probe_r02.py > gpurun_out/ins_probe_r02.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from drs_abm_b200 import _lib, synth
from drs_abm_b200.dispatch import InsertionTick
from drs_abm_b200.routing import RoadGraph

data = synth.config_graph("C3")
G = RoadGraph(data).build()
n_veh, n_new = bench.INSERTION["C3"]
sc = synth.fleet_scenario(data, n_veh, n_new, 600.0, 4, lambda s, t: G.matrix_lookup(s, t))
ref = None
for staged in ("1", "0"):
    os.environ["DRS_INS_STAGED"] = staged
    tick = InsertionTick(G, sc["vehs"], sc["table"])
    acc = []
    for it in range(6):
        t0 = time.time()
        bv, bi, bj, dc, ncand = tick.evaluate(sc["new_rows"])
        wall = time.time() - t0
        if it >= 2:
            acc.append(dict(tick.last_ms(), wall_ms=wall * 1e3))
    tick.close()
    out = {k: round(float(np.mean([a[k] for a in acc])), 4) for k in acc[0]}
    res = (bv.tolist(), bi.tolist(), bj.tolist(), dc.tolist())
    if ref is None:
        ref = res
    out.update(staged=staged, candidates=int(ncand), assigned=int((bv >= 0).sum()), same_as_staged=(res == ref),
               candidates_per_s=ncand / ((out["scan"] + out["fold"]) * 1e-3))
    print(json.dumps(out), flush=True)
os.environ["DRS_INS_STAGED"] = "1"
for nq in (200, 2000):
    tick = InsertionTick(G, sc["vehs"], sc["table"])
    t0 = time.time()
    bv, bi, bj, dc, ncand = tick.queue(sc["new_rows"][:nq], 600.0)
    wall = time.time() - t0
    ms = tick.last_ms()
    tick.close()
    print(json.dumps({"queue_requests": nq, "wall_ms": round(wall * 1e3, 2), "scan_ms": round(ms["scan"], 3), "queue_loop_ms": round(ms["queue"], 3),
                      "assigned": int((bv >= 0).sum()), "distinct_winners": int(len(set(bv[bv >= 0].tolist()))), "candidates": int(ncand)}), flush=True)
G.close()
