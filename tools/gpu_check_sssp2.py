"""SSSP CTA-size / unroll / delta sweep on the C3 all-pairs rows (one process per knob setting)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph
lib = _lib.load()
data = synth.config_graph(sys.argv[1] if len(sys.argv) > 1 else "C3")
mults = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6.0]
G = RoadGraph(data, numbering="input")
h = G.upload()
npad = (data.n + 127) // 128 * 128
buf = torch.empty((npad, npad), dtype=torch.float32, device="cuda")
mean_w = float(np.mean(data.eattrs["ttmedian"]))
for mult in mults:
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        _lib.check(lib.drs_apsp_rows_sssp(h, 0, C.c_void_p(buf.data_ptr()), npad, 0, npad, mult * mean_w, None))
        torch.cuda.synchronize(); ts.append(time.time() - t0)
    print("threads=%s S=%s delta=%4.1fx mean: %.4f s (min of 3) checksum %.1f" % (
        os.environ.get("DRS_SSSP_THREADS", "dflt"), os.environ.get("DRS_SSSP_S", "dflt"), mult, min(ts),
        float(buf[:data.n, :data.n].double().sum())), flush=True)
