"""Separator-partitioned build against the SSSP build on the same device graph: torch.equal of the two matrices,
stage profile, and times (development probe; the tests in tests/test_gpu_sep.py are the gate)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drs_abm_b200 import _lib, synth
from drs_abm_b200.apsp_dist import CudaTiles, build_sharded
from drs_abm_b200.routing import RoadGraph

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="C1,C1_uniform,C2,C3")
ap.add_argument("--targets", default="0")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--repeat", type=int, default=3)
args = ap.parse_args()
lib = _lib.load()
for name in args.configs.split(","):
    data = synth.config_graph(name)
    for target in [int(x) for x in args.targets.split(",")]:
        G = RoadGraph(data, mode=args.mode, numbering="partition", part_target=target or None)
        t0 = time.time()
        h = G.upload()
        t_up = time.time() - t0
        info = [C.c_int32(0) for _ in range(5)]
        _lib.check(lib.drs_sep_info(h, *[C.byref(x) for x in info]), lib)
        tiles = CudaTiles(h, args.mode)
        npad = (data.n + 127) // 128 * 128
        out = {"config": name, "target": target, "n": data.n, "upload_s": round(t_up, 4), "parts": info[0].value, "nsep": info[1].value,
               "ktot": info[2].value, "kmax": info[3].value, "usable": info[4].value}
        i2 = [C.c_int32(0), C.c_int32(0), C.c_int32(0)]
        _lib.check(lib.drs_sep_info2(h, C.byref(i2[0]), C.byref(i2[1]), None), lib)
        out["groups"], out["top"] = i2[0].value, i2[1].value
        res = {}
        for algo in ("sep", "sssp"):
            ms = []
            for r in range(args.repeat):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                local, _, _ = build_sharded(tiles, npad, algorithm=algo, with_pred=False)
                e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
                if r < args.repeat - 1:
                    del local
            res[algo] = local
            out[algo + "_ms"] = [round(x, 3) for x in ms]
            if algo == "sep":
                prof = np.zeros(8)
                _lib.check(lib.drs_apsp_sep_profile(h, _lib.ptr_f64(prof)), lib)
                out["sep_stage_ms"] = [round(float(x), 3) for x in prof]
                _lib.check(lib.drs_sep_info2(h, None, None, C.byref(i2[2])), lib)
                out["used16"] = i2[2].value
        out["equal"] = bool(torch.equal(res["sep"], res["sssp"]))
        if not out["equal"]:
            diff = (res["sep"] != res["sssp"])
            out["n_diff"] = int(diff.sum().item())
            idx = diff.nonzero()[:5].tolist()
            out["first_diff"] = [(i, j, float(res["sep"][i, j]), float(res["sssp"][i, j])) for i, j in idx]
        print(json.dumps(out), flush=True)
        del res, local
        G.close()
        torch.cuda.empty_cache()
