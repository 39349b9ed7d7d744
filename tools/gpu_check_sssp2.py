"""SSSP CTA-size / delta sweep on C3 rows (run once per DRS_SSSP_THREADS value)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph
lib = _lib.load()
data = synth.config_graph(sys.argv[1] if len(sys.argv) > 1 else "C3")
G = RoadGraph(data)
h = G.upload()
npad = (data.n + 127) // 128 * 128
buf = torch.empty((npad, npad), dtype=torch.float32, device="cuda")
mean_w = float(np.mean(data.eattrs["ttmedian"]))
for mult in (3, 6, 12, 24):
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        _lib.check(lib.drs_apsp_rows_sssp(h, 0, C.c_void_p(buf.data_ptr()), npad, 0, npad, mult * mean_w, None))
        torch.cuda.synchronize(); ts.append(time.time() - t0)
    print("threads=%s delta=%4.1fx mean: %.4f s (min of 3; first %.4f) checksum %.1f" % (os.environ.get("DRS_SSSP_THREADS", "default"), mult, min(ts), ts[0], float(buf[:data.n, :data.n].double().sum())), flush=True)
