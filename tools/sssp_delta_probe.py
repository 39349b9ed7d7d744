"""Bucket-width sweep of the SSSP build on C3 under the default (locality) numbering, all 50 171 sources.
Run on a GPU box: python tools/sssp_delta_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph

lib = _lib.load()
data = synth.config_graph("C3")
G = RoadGraph(data)
h = G.upload()
npad = (data.n + 127) // 128 * 128
buf = torch.empty((npad, npad), dtype=torch.float32, device="cuda")
mean_w = float(np.mean(data.eattrs["ttmedian"]))
stream = torch.cuda.current_stream()
for mult in (6, 3, 4, 5, 6, 7, 8, 10, 12):
    best = 1e9
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.drs_apsp_rows_sssp(h, 0, C.c_void_p(buf.data_ptr()), npad, 0, npad, mult * mean_w, C.c_void_p(stream.cuda_stream)), lib)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("delta = %4.1f x mean weight: %.2f ms" % (mult, best), flush=True)
G.close()
