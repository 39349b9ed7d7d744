"""Round-2 SSSP probe on a GPU box: shared-memory-resident rows against the L2-resident kernel.
All sources of one config, every (variant, CTA size, bucket width) setting; each result is compared
bit for bit with the L2 kernel's matrix.   python tools/sssp_probe_r02.py [C3] > gpurun_out/sssp_probe_r02.jsonl"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
lib = _lib.load()
data = synth.config_graph(cfg)
G = RoadGraph(data)
h = G.upload()
npad = (data.n + 127) // 128 * 128
base = torch.empty((npad, npad), dtype=torch.float32, device="cuda")
buf = torch.empty((npad, npad), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream()


def build(out, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.drs_apsp_rows_sssp(h, 0, C.c_void_p(out.data_ptr()), npad, 0, npad, 0.0, C.c_void_p(stream.cuda_stream)), lib)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


os.environ["DRS_SSSP_IMPL"] = "l2"
os.environ.pop("DRS_SSSP_DELTA_MULT", None)
ms = build(base)
print(json.dumps({"config": cfg, "n": data.n, "impl": "l2", "ms": round(ms, 3)}), flush=True)
os.environ["DRS_SSSP_IMPL"] = "smem"
settings = [(th, mult) for th in (1024, 512, 256) for mult in (6, 12, 24, 48, 0)]
if len(sys.argv) > 2:
    settings = [tuple(int(x) for x in s.split(":")) for s in sys.argv[2].split(",")]
for th, mult in settings:
    if th > 2000:                       # a shape of the row-in-L2 lane-per-arc kernel: CTA size * 100 + CTAs per SM
        os.environ["DRS_SSSP_IMPL"] = "arc"
        os.environ["DRS_SSSP_ARC_SHAPE"] = str(th)
    else:
        os.environ["DRS_SSSP_IMPL"] = "smem"
        os.environ["DRS_SSSP_SMEM_THREADS"] = str(th)
    os.environ["DRS_SSSP_DELTA_MULT"] = str(mult)
    buf.fill_(-1.0)
    try:
        ms = build(buf)
        same = bool(torch.equal(buf, base))
    except Exception as ex:  # noqa: BLE001
        print(json.dumps({"impl": "smem", "threads": th, "delta_mult": mult, "error": str(ex)}), flush=True)
        continue
    print(json.dumps({"config": cfg, "impl": os.environ["DRS_SSSP_IMPL"], "threads_or_shape": th, "delta_mult": mult, "ms": round(ms, 3), "bit_equal_to_l2": same}), flush=True)
G.close()
